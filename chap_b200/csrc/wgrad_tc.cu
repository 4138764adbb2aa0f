// tcgen05 weight-gradient kernel: stride-1 convolutions (k=3 pad 1, k=1; 2D and 3D) and the 2D transposed k2 s2 convolution.
//
//     dW[tap][co][ci] = sum_p dy[p, co] * x[p + tap, ci]
// is, per tap, a GEMM whose reduction dimension is the pixel index p.  Both operands are channels-last, i.e. the M and N
// dimensions are the contiguous ones: "MN-major" UMMA operands.  A spatial box of P pixels (64 / 128 / 256) is one K block:
// the dy box [channel chunk x P] and the x box shifted by the tap offset are fetched by tiled TMA loads (out-of-bounds pixels
// zero-filled = the conv padding) as rows of 32 channels (128 B, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B = UMMA layout type 1, the
// only layout tcgen05 accepts for MN-major 32-bit operands; 16- and 4-channel tensors are zero-filled to 32).  K atom = 4 rows
// (SBO = 512 B), LBO = distance between 32-channel chunks; one tcgen05.mma.kind::tf32 (K = 8) consumes 8 pixel rows.
//
// Modes (chosen on the host, tc_wgrad):
//   * row reuse (k3, <= 128 input channels per CTA): the x box carries an h-halo and serves the three ky taps at row offsets;
//   * swap (row reuse and cin <= 32): x is the M operand, its LBO is tw rows so that chunk g IS tap ky = g (one MMA = 3 taps),
//     dy is the N operand with N = cout (>= 16; the 4-channel heads are zero-filled);
//   * taps-in-N (row reuse and cout <= 64; supersedes swap): dW[ky, kx] = sum_q x[q + (ky - 1) W] dy[q - (kx - 1)], so the kx shift can sit on
//     dy: x is the M operand as in swap mode (chunk g = tap ky) and dy is loaded three times, shifted by kx, as ONE N operand
//     [3 kx][cout chunks] (LBO = box stride).  One MMA per K step covers all nine taps at N = 96 / 192 where swap needed three (N = cout)
//     and the plain mode nine -- these kernels are bound by the NUMBER of K = 8 MMAs (~40 ns each), not by their width;
//   * transposed conv: M = x (tap independent), N = dy gathered per tap through its [2C, W, 2, H, N] view.
// Work split: blockIdx.z = (M tile, N tile), blockIdx.y = tap group (as many taps as fit the TMEM columns), blockIdx.x = slice
// of the pixel blocks.  Each CTA accumulates its slice in TMEM (fp32) and adds it either straight into the torch-layout
// gradient with red.global.add.f32 (thin layers) or, with 128-bit red.global.add.v4.f32, into a [tap][M][N] scratch that a
// small kernel transposes into the torch layout (deep layers: the scalar reductions were ~50 % of their time).
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include "common.cuh"
#include "conv_plan.cuh"
#include "tc_common.cuh"

namespace chap {

struct WgParams {
    int nd, ksz, pad, taps;
    int W, H, D, n_img;
    int tw, th, td, tiles_w, tiles_h, tiles_d;
    int P;                          // pixels per K block padded to a multiple of 8 (smem rows per channel chunk)
    int p_box;                      // pixels actually loaded per block (= tw*th*td <= P); rows p_box..P stay zero
    int cout, cin;                  // full channel counts (gradient strides)
    int m_tile, n_tile;             // channels of dy / x handled by this CTA
    int mma_m;                      // 64 or 128
    int a_cpg, a_groups;            // channels per TMA chunk (32|16) and chunks per m_tile
    int b_cpg, b_groups;
    int tg;                         // taps per CTA (reuse mode: (kz, kx) pairs per CTA, each pair = 3 ky taps)
    int k2s2;                       // 1: k2 s2 convolution (strided or transposed): the N operand is the high-resolution tensor gathered per
                                    // tap through one tensor map per tap; the pixel blocks tile the LOW-resolution grid
    int swap;                       // 1 (reuse mode, cin <= 32): x is the M operand (M = 4 "ky" chunks x 32 ci, chunk stride tw rows), dy the N operand
    int tapn;                       // 1 (row-reuse geometry, cout <= 64): ALL NINE in-plane taps per MMA -- x is the M operand (4 "ky" chunks of one
                                    // 32-channel group, chunk stride tw rows of the h-haloed box), dy the N operand loaded THREE times shifted by
                                    // kx (N = 3 kx x cout rounded up to 32); one MMA per K step and 32-channel group of x
    int n_cols;                     // tapn: N of the MMA = 3 * a_groups * 32 = accumulator columns per (unit, x channel group)
    int acc2;                       // experiment (CHAP_WG_ACC2): alternate K steps between two TMEM accumulators, summed in the epilogue
    int n_mma;                      // swap mode: MMA N = cout of this CTA rounded up to 16
    int reuse;                      // 1: the x box carries an h-halo (th + 2 rows) and serves the 3 ky taps at row offsets ky * tw
    int b_rows;                     // rows (pixels) per x channel chunk in smem
    int debug;                      // CHAP_WG_DEBUG bit 0: skip the MMAs, bit 1: skip the TMA loads (timing experiments only)
    int stages, tmem_cols;          // stages: ring slots of the per-tap N-operand (x) boxes
    int a_stages;                   // ring slots of the M-operand (dy) box, which is loaded ONCE per pixel block and shared by all taps
    int blocks_total, blocks_per_cta;
    uint32_t a_stage_bytes, b_stage_bytes;
    int64_t s_co, s_ci;             // gradient strides in floats (torch layout), tap stride is 1
    float* dw;
    float* acc;                     // nullable: [tap][cout (M)][cin (N)] accumulation scratch for 128-bit vector reductions
};

// red.global.add.v4.f32 (sm_90+): one L2 reduction for four consecutive floats
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// acc [tap][M][N] -> torch-layout gradient dw[m * s_co + n * s_ci + tap]
__global__ void __launch_bounds__(256)
wgrad_unpack_kernel(float* __restrict__ acc, float* __restrict__ dw, int taps, int M, int N, int64_t s_co, int64_t s_ci, int accumulate) {
    pdl_enter();
    const int64_t total = (int64_t)taps * M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // consecutive threads walk the torch layout's fastest index (tap) so that the writes coalesce
        const int t = (int)(i % taps);
        const int64_t r = i / taps;
        const int n = (int)(r % N), m = (int)(r / N);
        float* src = acc + ((int64_t)t * M + m) * N + n;
        float* dst = dw + (int64_t)m * s_co + (int64_t)n * s_ci + t;
        if (accumulate) { *dst += *src; *src = 0.f; }        // gradient-sink mode: add into the arena, hand the scratch back zeroed
        else *dst = *src;
    }
}

constexpr int kWgThreads = 192;

// MN-major TF32 operand descriptor (cute::UMMA::SmemDescriptor), layout type 1 = SWIZZLE_128B_BASE32B, the only smem
// layout tcgen05 accepts for MN-major 32-bit operands: rows of 128 B (32 channels), K atom = 4 rows, so
// SBO = 4 rows = 512 B (stride between K atoms), LBO = stride between 32-channel chunks.
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t saddr, uint32_t row_bytes, uint32_t lbo_bytes) {
    const uint64_t sbo = (4u * row_bytes) >> 4;
    const uint64_t layout = 1ull;
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

__device__ __forceinline__ void wgrad_tc_kernel_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const TmTaps* tmTp, const WgParams& p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // The dy box of a pixel block is the same for every tap: it has its own small ring (a_stages slots, one fill per BLOCK),
    // the x boxes shifted by the tap keep the per-tap ring.  Round 1 re-loaded dy with every tap: 3x (2D) / 9x (3D) its bytes
    // through the L2 -> shared-memory path that bounds this kernel (679 MB for a 100 MB layer, profiles/r02_conv_ncu_full_summary.md).
    uint8_t* a_base = smem;
    uint8_t* b_base = smem + (size_t)p.a_stages * p.a_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + (size_t)p.stages * p.b_stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + p.stages;
    uint64_t* tmem_full = bars + 2 * p.stages;
    uint64_t* a_full = tmem_full + 1;
    uint64_t* a_empty = a_full + p.a_stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + p.a_stages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = p.cout / p.m_tile;
    const int m0 = (blockIdx.z % m_tiles) * p.m_tile, n0 = (blockIdx.z / m_tiles) * p.n_tile;      // n0 > 0 only when cin > 256
    const int tap0 = blockIdx.y * p.tg;
    const int units = p.tapn ? p.taps / 9 : (p.reuse ? p.taps / 3 : p.taps);   // scheduling units: taps, (kz, kx) pairs, or kz planes
    const int ntaps = min(p.tg, units - tap0);
    const int nky = p.reuse ? 3 : 1;
    const int blk0 = blockIdx.x * p.blocks_per_cta;
    const int nblk = min(p.blocks_per_cta, p.blocks_total - blk0);

    pdl_trigger();          // prologue (barriers, shared-memory zero fill, TMEM) overlaps the previous kernel's tail
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < p.a_stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (p.p_box != p.P) {
        // the TMA boxes cover p_box < P rows of every channel chunk: the remaining rows of the last K atom must read as zero
        float4* z = reinterpret_cast<float4*>(smem);
        const int n16 = (int)(((size_t)p.a_stages * p.a_stage_bytes + (size_t)p.stages * p.b_stage_bytes) >> 4);
        for (int i = threadIdx.x; i < n16; i += kWgThreads) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_row = (uint32_t)p.a_cpg * 4u, b_row = (uint32_t)p.b_cpg * 4u;
    const uint32_t a_chunk = (uint32_t)p.P * a_row, b_chunk = (uint32_t)p.b_rows * b_row;     // bytes of one channel chunk

    if (nblk > 0 && warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            int as = 0; uint32_t aph = 0;
            for (int b = 0; b < nblk; ++b) {
                int t = blk0 + b;
                const int tx = t % p.tiles_w; t /= p.tiles_w;
                const int ty = t % p.tiles_h; t /= p.tiles_h;
                const int tz = t % p.tiles_d;
                const int img = t / p.tiles_d;
                const int w0 = tx * p.tw, h0 = ty * p.th, d0 = tz * p.td;
                {   // the dy box of this block: once, for all taps
                    mbar_wait(&a_empty[as], aph ^ 1);
                    if (p.debug & 2) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&a_full[as])) : "memory");
                    else if (p.tapn) {
                        // three copies of the dy box shifted by kx: chunk (kx, g) holds dy[q - (kx - 1)] for the block's pixels q
                        mbar_expect_tx(&a_full[as], 3u * (uint32_t)p.a_groups * (uint32_t)p.p_box * 128u);
                        uint8_t* a_dst = a_base + (size_t)as * p.a_stage_bytes;
                        for (int kx = 0; kx < 3; ++kx)
                            for (int g = 0; g < p.a_groups; ++g) {
                                uint8_t* dst = a_dst + (size_t)(kx * p.a_groups + g) * a_chunk;
                                if (p.nd == 2) tma_load_4d(dst, &tmA, &a_full[as], m0 + g * p.a_cpg, w0 + 1 - kx, h0, img);
                                else tma_load_5d(dst, &tmA, &a_full[as], m0 + g * p.a_cpg, w0 + 1 - kx, h0, d0, img);
                            }
                    } else {
                        mbar_expect_tx(&a_full[as], (uint32_t)p.a_groups * (uint32_t)p.p_box * 128u);
                        uint8_t* a_dst = a_base + (size_t)as * p.a_stage_bytes;
                        for (int g = 0; g < p.a_groups; ++g) {
                            if (p.nd == 2) tma_load_4d(a_dst + (size_t)g * a_chunk, &tmA, &a_full[as], m0 + g * p.a_cpg, w0, h0, img);
                            else tma_load_5d(a_dst + (size_t)g * a_chunk, &tmA, &a_full[as], m0 + g * p.a_cpg, w0, h0, d0, img);
                        }
                    }
                    if (++as == p.a_stages) { as = 0; aph ^= 1; }
                }
                for (int ti = 0; ti < ntaps; ++ti) {
                    const int tap = tap0 + ti;
                    int kx, ky, kz;
                    if (p.tapn) { kx = 1; kz = tap; ky = 0; }                          // unit = kz; no kx shift on x (it sits on dy)
                    else if (p.reuse) { kx = tap % 3; kz = tap / 3; ky = 0; }          // box origin one row above the tile
                    else if (p.ksz == 3) { kx = tap % 3; ky = (tap / 3) % 3; kz = tap / 9; } else { kx = ky = kz = 0; }
                    mbar_wait(&empty[s], ph ^ 1);
                    if (p.debug & 2) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[s])) : "memory"); if (++s == p.stages) { s = 0; ph ^= 1; } continue; }
                    mbar_expect_tx(&full[s], (uint32_t)p.b_groups * (uint32_t)(p.reuse ? p.b_rows : p.p_box) * 128u);
                    uint8_t* b_dst = b_base + (size_t)s * p.b_stage_bytes;
                    for (int g = 0; g < p.b_groups; ++g) {
                        if (p.k2s2 && p.nd == 2) tma_load_4d(b_dst + (size_t)g * b_chunk, &tmTp->m[tap], &full[s], n0 + g * p.b_cpg, w0, h0, img);
                        else if (p.k2s2) tma_load_5d(b_dst + (size_t)g * b_chunk, &tmTp->m[tap], &full[s], n0 + g * p.b_cpg, w0, h0, d0, img);
                        else if (p.nd == 2) tma_load_4d(b_dst + (size_t)g * b_chunk, &tmB, &full[s], n0 + g * p.b_cpg, w0 + kx - p.pad, h0 + ky - p.pad, img);
                        else tma_load_5d(b_dst + (size_t)g * b_chunk, &tmB, &full[s], n0 + g * p.b_cpg, w0 + kx - p.pad, h0 + ky - p.pad, d0 + kz - p.pad, img);
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (nblk > 0 && warp == 1) {
        // The whole warp runs the (warp-uniform) loop and one elected lane issues, so descriptors live in uniform registers
        // and consecutive tcgen05.mma are a couple of integer adds apart: the issuing thread's own instruction stream is
        // on the critical path of these small-N MMAs.
        // D = F32, A = B = TF32, both MN-major (bits 15, 16), N >> 3 at 17, M >> 4 at 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(p.mma_m >> 4) << 24);
        // Swap mode (thin layers, cin <= 32): one tcgen05.mma costs time ~ N and M is 128 rows whatever is used, so x goes
        // on the M side: the "channel chunk" stride (LBO) of its descriptor is tw rows of the h-haloed box, i.e. chunk g IS
        // tap ky = g (chunk 3 is junk that is never stored), and dy is the N operand with N = cout.  One MMA per K step
        // covers three taps at N = cout (16..128) instead of N = 3 * 32 with 16 of 128 rows in use.
        const bool swap = p.swap != 0 || p.tapn != 0;
        // the extra M chunks alias chunk 0 (one chunk, LBO = 0) or read whatever follows in shared memory: never stored
        const uint32_t a_lbo = (p.a_groups >= 2 || p.tapn) ? a_chunk : 0u;
        const uint32_t b_lbo = swap ? (uint32_t)p.tw * b_row : b_chunk;
        // descriptor words: lo = start >> 4 | (LBO >> 4) << 16; hi = SBO >> 4 (4 rows) | version 1 (bit 46) | layout 1 (bit 61)
        const uint32_t a_hi = ((4u * a_row) >> 4) | (1u << 14) | (1u << 29);
        const uint32_t b_hi = ((4u * b_row) >> 4) | (1u << 14) | (1u << 29);
        const uint32_t a_lo0 = ((smem_u32(a_base) >> 4) & 0x3FFFu) | (((a_lbo >> 4) & 0x3FFFu) << 16);
        const uint32_t b_lo0 = ((smem_u32(b_base) >> 4) & 0x3FFFu) | (((b_lbo >> 4) & 0x3FFFu) << 16);
        const uint32_t a_stage = p.a_stage_bytes >> 4, b_stage = p.b_stage_bytes >> 4;
        const uint32_t a_k = (8u * a_row) >> 4, b_k = (8u * b_row) >> 4;        // one K step = 8 pixel rows
        const uint32_t b_ky = ((uint32_t)p.tw * b_row) >> 4;                     // reuse mode: tap ky starts ky * tw rows further down
        const int ksteps = (p.debug & 1) ? 0 : p.P / 8;
        const int n_sub = swap ? 1 : nky;
        const uint32_t n_mma = p.tapn ? (uint32_t)p.n_cols : (uint32_t)p.n_mma;
        const uint32_t id = swap ? ((idesc & ~(0x3Fu << 17)) | ((n_mma >> 3) << 17)) : idesc;
        const uint32_t d_stride = p.tapn ? (uint32_t)(p.b_groups * p.n_cols) : (swap ? (uint32_t)p.n_mma : (uint32_t)(nky * p.n_tile));
        int s = 0; uint32_t ph = 0;
        int as = 0; uint32_t aph = 0;
        uint32_t a_lo = a_lo0, b_lo = b_lo0;
        for (int b = 0; b < nblk; ++b) {
            uint32_t d_tap = tmem_base;
            mbar_wait(&a_full[as], aph);                         // this block's dy box (shared by all its taps)
            for (int ti = 0; ti < ntaps; ++ti) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (p.tapn) {
                    if (elect_one()) {
                        // per 32-channel group of x: ONE MMA per K step, M = (ky, ci), N = (kx, co)
                        for (int g = 0; g < p.b_groups; ++g) {
                            const uint32_t d_tmem = d_tap + (uint32_t)(g * p.n_cols);
                            uint32_t m_d = b_lo + (uint32_t)g * (b_chunk >> 4), n_d = a_lo;
                            if ((ksteps & 7) == 0) {
                                for (int k0 = 0; k0 < ksteps; k0 += 8) {
#pragma unroll
                                    for (int k = 0; k < 8; ++k)
                                        tc_mma_tf32_lh(d_tmem, m_d + (uint32_t)k * b_k, b_hi, n_d + (uint32_t)k * a_k, a_hi, id, (uint32_t)((b | k0 | k) != 0));
                                    m_d += 8u * b_k; n_d += 8u * a_k;
                                }
                            } else {
                                for (int k = 0; k < ksteps; ++k)
                                    tc_mma_tf32_lh(d_tmem, m_d + (uint32_t)k * b_k, b_hi, n_d + (uint32_t)k * a_k, a_hi, id, (uint32_t)((b | k) != 0));
                            }
                        }
                        tc_commit(&empty[s]);
                        if (ti == ntaps - 1) tc_commit(&a_empty[as]);
                    }
                } else
                if (elect_one()) {
                    for (int ky = 0; ky < n_sub; ++ky) {
                        const uint32_t d_tmem = d_tap + (uint32_t)(ky * p.n_tile);
                        const uint32_t b_tap = b_lo + (uint32_t)ky * b_ky;
                        // MMA operands: (dy, x) or, swapped, (x, dy)
                        uint32_t m_d = swap ? b_tap : a_lo, n_d = swap ? a_lo : b_tap;
                        const uint32_t m_hi = swap ? b_hi : a_hi, n_hi = swap ? a_hi : b_hi;
                        const uint32_t m_k = swap ? b_k : a_k, n_k = swap ? a_k : b_k;
                        if ((ksteps & 7) == 0 && p.acc2) {
                            const uint32_t acc_stride = (uint32_t)p.tmem_cols >> 1;
                            for (int k0 = 0; k0 < ksteps; k0 += 8) {
#pragma unroll
                                for (int k = 0; k < 8; ++k)
                                    tc_mma_tf32_lh(d_tmem + ((k & 1) ? acc_stride : 0u), m_d + (uint32_t)k * m_k, m_hi, n_d + (uint32_t)k * n_k, n_hi, id,
                                                   (uint32_t)((b | k0 | (k >> 1)) != 0));
                                m_d += 8u * m_k; n_d += 8u * n_k;
                            }
                        } else if ((ksteps & 7) == 0) {
                            for (int k0 = 0; k0 < ksteps; k0 += 8) {
#pragma unroll
                                for (int k = 0; k < 8; ++k)
                                    tc_mma_tf32_lh(d_tmem, m_d + (uint32_t)k * m_k, m_hi, n_d + (uint32_t)k * n_k, n_hi, id, (uint32_t)((b | k0 | k) != 0));
                                m_d += 8u * m_k; n_d += 8u * n_k;
                            }
                        } else {
                            for (int k = 0; k < ksteps; ++k)
                                tc_mma_tf32_lh(d_tmem, m_d + (uint32_t)k * m_k, m_hi, n_d + (uint32_t)k * n_k, n_hi, id, (uint32_t)((b | k) != 0));
                        }
                    }
                    tc_commit(&empty[s]);
                    if (ti == ntaps - 1) tc_commit(&a_empty[as]);          // all MMAs that read this dy box are in flight behind this commit
                }
                __syncwarp();
                d_tap += d_stride;
                b_lo += b_stage;
                if (++s == p.stages) { s = 0; ph ^= 1; b_lo = b_lo0; }
            }
            a_lo += a_stage;
            if (++as == p.a_stages) { as = 0; aph ^= 1; a_lo = a_lo0; }
        }
        if (elect_one()) tc_commit(tmem_full);
        __syncwarp();
    } else if (nblk > 0 && warp >= 2) {
        const int lg = warp & 3;
        const int row = lg * 32 + lane;                  // accumulator row = output channel inside the tile
        const bool valid = row < p.m_tile && row < p.mma_m;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        if (p.tapn) {
            // accumulator row = (ky = lane quarter, ci = lane of x channel group g); columns = [unit kz][g][kx][co chunk][32 co]
            const int ky = lg;
            if (ky < 3) {
                for (int ti = 0; ti < ntaps; ++ti) {
                    const int kz = tap0 + ti;
                    for (int g = 0; g < p.b_groups; ++g) {
                        const int ci = n0 + g * 32 + lane;
                        for (int kx = 0; kx < 3; ++kx) {
                            const int tap = (kz * 3 + ky) * 3 + kx;
                            for (int c0 = 0; c0 < p.a_groups * 32; c0 += 16) {
                                float v[16];
                                tc_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((ti * p.b_groups + g) * p.n_cols + kx * p.a_groups * 32 + c0), v);
                                if (ci < p.cin && c0 < p.m_tile) {                  // else: zero-filled channels of x / dy
                                    if (p.acc) {                                    // [tap][ci][co] scratch, 128-bit reductions (cout % 16 == 0)
                                        float* a = p.acc + ((int64_t)tap * p.cin + ci) * p.cout + m0 + c0;
#pragma unroll
                                        for (int j = 0; j < 16; j += 4) red_add_v4(a + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
                                    } else {
                                        float* dst = p.dw + (int64_t)ci * p.s_ci + tap;
#pragma unroll
                                        for (int j = 0; j < 16; ++j)
                                            if (c0 + j < p.m_tile) atomicAdd(dst + (int64_t)(m0 + c0 + j) * p.s_co, v[j]);
                                    }
                                }
                            }
                        }
                    }
                }
            }
        } else if (p.swap) {
            // accumulator row = (ky = lane quarter, ci = lane); columns = [unit (kz, kx)][co]
            const int ky = lg, ci = lane;
            if (ky < 3) {
                for (int ti = 0; ti < ntaps; ++ti) {
                    const int q = tap0 + ti;
                    const int tap = ((q / 3) * 3 + ky) * 3 + (q % 3);
                    float* dst = p.dw + (int64_t)ci * p.s_ci + tap;
                    for (int c0 = 0; c0 < p.n_mma; c0 += 16) {
                        float v[16];
                        tc_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(ti * p.n_mma + c0), v);
                        if (p.acc2) {
                            float u[16];
                            tc_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((p.tmem_cols >> 1) + ti * p.n_mma + c0), u);
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] += u[j];
                        }
                        if (ci < p.cin) {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (c0 + j < p.m_tile) atomicAdd(dst + (int64_t)(m0 + c0 + j) * p.s_co, v[j]);
                        }
                    }
                }
            }
        } else if (lg * 32 < p.mma_m) {                  // M = 64: lane groups 2,3 hold nothing
            float* dst_row = p.dw + (int64_t)(m0 + row) * p.s_co;
            for (int ti = 0; ti < ntaps * nky; ++ti) {
                int tap = tap0 + ti;
                if (p.reuse) { const int q = tap0 + ti / 3, ky = ti % 3; tap = ((q / 3) * 3 + ky) * 3 + (q % 3); }
                for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
                    float v[16];
                    tc_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(ti * p.n_tile + c0), v);
                    if (valid && n0 + c0 < p.cin) {                 // columns beyond cin are the zero-filled channels
                        if (p.acc) {
                            float* a = p.acc + ((int64_t)tap * p.cout + m0 + row) * p.cin + n0 + c0;
#pragma unroll
                            for (int j = 0; j < 16; j += 4) red_add_v4(a + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                atomicAdd(dst_row + (int64_t)(n0 + c0 + j) * p.s_ci + tap, v[j]);
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// Two entry points: the per-tap tensor maps (1 KB of kernel parameters) are only passed for the k2 s2 gathers.
__global__ void __launch_bounds__(kWgThreads)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgParams p) {
    wgrad_tc_kernel_body(tmA, tmB, nullptr, p);
}
__global__ void __launch_bounds__(kWgThreads)
wgrad_tc_kernel_taps(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ TmTaps tmT,
                     const WgParams p) {
    wgrad_tc_kernel_body(tmA, tmB, &tmT, p);
}

bool tc_wgrad_supports(const Geom& g) {
    auto ok = [](int c) { return c == 16 || (c % 32 == 0 && c <= 1024); };
    if (g.kind == CHAP_CONV_UP2 || g.kind == CHAP_CONV_DOWN2) return ok(g.cin) && ok(g.cout) && getenv("CHAP_NO_UP2_TC") == nullptr;
    if (g.kind != CHAP_CONV_K3 && g.kind != CHAP_CONV_K1) return false;
    // swap mode (row-reuse geometry, cin <= 32) also takes the 4- / 8-channel heads: dy is the N operand, zero-filled to 16
    const bool head = (g.cout == 4 || g.cout == 8) && g.kind == CHAP_CONV_K3 && (g.cin == 16 || g.cin == 32) && g.iW >= 8 && g.iH >= 10 &&
                      getenv("CHAP_NO_ROW_REUSE") == nullptr && getenv("CHAP_WG_NO_SWAP") == nullptr;
    return ok(g.cin) && (ok(g.cout) || head);
}

// spatial box inside the image (w <= W, h <= H, d <= D), at most 64 pixels, minimising the padded reduction length
static void choose_box8(int W, int H, int D, int& tw, int& th, int& td, int max_px = 64) {
    long best = -1;
    for (int w = 1; w <= W && w <= max_px; ++w)
        for (int h = 1; h <= H && w * h <= max_px; ++h)
            for (int d = 1; d <= D && w * h * d <= max_px; ++d) {
                const int P = (w * h * d + 7) / 8 * 8;
                long tiles = (long)((W + w - 1) / w) * ((H + h - 1) / h) * ((D + d - 1) / d);
                long score = tiles * P * 64 - w * h * d;    // least padded work, then the larger block
                if (best < 0 || score < best) { best = score; tw = w; th = h; td = d; }
            }
}

int tc_wgrad(const Geom& g, const float* x, const float* dy, float* dw, cudaStream_t st, float* acc_ws, bool accumulate) {
    if (!tc_wgrad_supports(g)) return 0;
    CHAP_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dw), CHAP_ERR_ALIGNMENT, "tc_wgrad: buffers must be 16-byte aligned");
    // k2 s2 convolutions: dW = sum_p S[p, .] * B[2p + tap, .] with S the low-resolution tensor (tap independent: the M operand) and
    // B the high-resolution one gathered per tap (the N operand).  Strided conv: S = dy, B = x (the usual orientation);
    // transposed conv: S = x, B = dy, i.e. the roles exchanged and the gradient strides those of the [ci][co][tap] layout.
    const bool up2 = g.kind == CHAP_CONV_UP2, down2 = g.kind == CHAP_CONV_DOWN2, k2 = up2 || down2;
    Geom v = g;                                  // v.cout = channels of the M operand, v.cin = channels of the N operand
    const float* m_src = dy; const float* n_src = x;
    if (up2) { v.cin = g.cout; v.cout = g.cin; m_src = x; n_src = dy; }
    WgParams p{};
    p.k2s2 = k2 ? 1 : 0;
    p.nd = g.nd; p.ksz = g.kind == CHAP_CONV_K3 ? 3 : 1; p.pad = g.kind == CHAP_CONV_K3 ? 1 : 0; p.taps = g.taps;
    p.W = g.iW; p.H = g.iH; p.D = g.iD; p.n_img = g.n;
    if (down2) { p.W = g.oW; p.H = g.oH; p.D = g.oD; }
    choose_box8(p.W, p.H, p.D, p.tw, p.th, p.td);
    // taps-in-N mode (see the header): row-reuse geometry, dy has <= 64 channels.  x is split into 32-channel groups; a CTA takes as many
    // groups as fit the 512 TMEM columns (n_cols accumulator columns per group and kz plane), the rest goes to blockIdx.z.
    const bool tapn = g.kind == CHAP_CONV_K3 && v.cout <= 64 && p.W >= 8 && p.H >= 10 && getenv("CHAP_NO_ROW_REUSE") == nullptr &&
                      getenv("CHAP_WG_NO_TAPN") == nullptr;
    const int cin_groups = (v.cin + 31) / 32;
    const int tapn_cols = 3 * ((v.cout + 31) / 32) * 32;
    int tapn_bg = cin_groups;
    while (tapn && (tapn_bg * tapn_cols > 512 || cin_groups % tapn_bg != 0)) --tapn_bg;
    const int a_mult = tapn ? 3 : 1;                               // the dy stage holds three kx-shifted copies
    const int n_tile_pre = tapn ? 32 * tapn_bg : (v.cin > 256 ? 256 : (v.cin < 32 ? 32 : v.cin));
    {
        // 128-pixel blocks (16 MMAs per pipeline stage instead of 8) when three stages still fit and the blocks fill the GPU
        int tw, th, td;
        choose_box8(p.W, p.H, p.D, tw, th, td, 128);
        const int m_pre = v.cout > 128 ? 128 : v.cout;
        const int P = (tw * th * td + 7) / 8 * 8;
        const size_t a_bytes = (size_t)a_mult * (((m_pre + 31) / 32) * 32) * P * 4, b_bytes = (size_t)n_tile_pre * P * 4;
        const long blocks = (long)g.n * ((p.W + tw - 1) / tw) * ((p.H + th - 1) / th) * ((p.D + td - 1) / td);
        if (a_bytes * 2 + b_bytes * 3 <= 198 * 1024 && blocks >= 4 * kNumSMs && getenv("CHAP_WG_BOX") == nullptr) { p.tw = tw; p.th = th; p.td = td; }
    }
    // Row-reuse mode (large images, <= 128 input channels per CTA): in-plane 8 x 8 pixel block; the x box is loaded once
    // per (kz, kx) with an h-halo of one row above and below and serves the three ky taps -> 3 (9) x-boxes of 10 rows
    // instead of 9 (27) boxes of 8 rows per block.
    p.reuse = 0;
    if (g.kind == CHAP_CONV_K3 && (n_tile_pre <= 128 || tapn) && p.W >= 8 && p.H >= 10 && getenv("CHAP_NO_ROW_REUSE") == nullptr) {
        p.reuse = 1; p.tw = 8; p.th = 8; p.td = 1;
        // The MMA-issuing warp pays a fixed ~0.4 us of scalar work per pipeline stage (measured), so large images use
        // larger pixel blocks (16 x 8 or 16 x 16: 16 / 32 MMAs per stage instead of 8) as long as >= 3 stages still fit
        // and there are enough blocks to fill the GPU.
        const int m_tile_pre = v.cout > 128 ? 128 : v.cout;
        const int force = getenv("CHAP_WG_BOX") ? atoi(getenv("CHAP_WG_BOX")) : 0;      // 64 / 128 / 256 pixels (experiments)
        const int cand[2][2] = {{16, 16}, {16, 8}};
        for (const auto& c : cand) {
            const int tw = c[0], th = c[1], P = tw * th;
            if (force && P != force) continue;
            if (p.W < tw || p.H < th + 2) continue;
            const size_t a_bytes = (size_t)a_mult * ((m_tile_pre + 31) / 32) * 32 * P * 4, b_bytes = (size_t)n_tile_pre * tw * (th + 2) * 4;
            const long blocks = (long)g.n * p.D * ((p.W + tw - 1) / tw) * ((p.H + th - 1) / th);
            // (taps-in-N: dy and x both advance once per block and plane, so two x slots pair with the two dy slots -- 64 -> 32 @ 12x128^2:
            //  47.9 us with 8 x 8 blocks and three slots, 40.2 us with 16 x 8 blocks and two)
            if (a_bytes * 2 + b_bytes * ((force || tapn) ? 2 : 3) > 198 * 1024 || blocks < 4 * kNumSMs) continue;
            p.tw = tw; p.th = th;
            break;
        }
    }
    p.p_box = p.tw * p.th * p.td;
    p.P = (p.p_box + 7) / 8 * 8;
    p.b_rows = p.reuse ? p.tw * (p.th + 2) : p.P;
    p.debug = getenv("CHAP_WG_DEBUG") ? atoi(getenv("CHAP_WG_DEBUG")) : 0;
    p.tiles_w = (p.W + p.tw - 1) / p.tw; p.tiles_h = (p.H + p.th - 1) / p.th; p.tiles_d = (p.D + p.td - 1) / p.td;
    p.cout = v.cout; p.cin = v.cin;
    p.m_tile = v.cout > 128 ? 128 : v.cout;
    p.n_tile = n_tile_pre;                                        // 16-channel tensors: the TMA box is 32 wide, channels 16..31 zero-filled
    p.tapn = (tapn && p.reuse) ? 1 : 0;
    p.n_cols = tapn_cols;
    p.mma_m = 128;           // M = 64 has a different TMEM lane mapping; M = 128 costs the same tensor time
    p.a_cpg = 32; p.a_groups = (p.m_tile + 31) / 32;
    p.b_cpg = 32; p.b_groups = p.n_tile / 32;
    p.a_stage_bytes = ((uint32_t)a_mult * p.a_groups * 32u * p.P * 4u + 1023u) & ~1023u;
    p.b_stage_bytes = ((uint32_t)p.n_tile * p.b_rows * 4u + 1023u) & ~1023u;
    p.a_stages = 2;
    const size_t a_ring = (size_t)p.a_stages * p.a_stage_bytes, stage = p.b_stage_bytes;      // `stage` = one per-tap ring slot (x box)
    p.blocks_total = g.n * p.tiles_d * p.tiles_h * p.tiles_w;
    const int n_tiles = p.tapn ? cin_groups / tapn_bg : (v.cin > 256 ? v.cin / 256 : 1);
    const int zdim = (v.cout / p.m_tile) * n_tiles;
    // Work split.  Parallelism comes from (channel tiles) x (tap groups) x (pixel splits).  Every pixel split adds one
    // fp32 atomic per (padded) weight element in the epilogue, so splits are capped by an atomic budget (measured: a
    // 256x256x9 layer with 24 splits spent >80% of its 106 us in 14 M atomics); tap groups are made smaller instead.
    // Two CTAs share an SM when a CTA needs <= 256 TMEM columns and <= 100 KB of smem.
    const long weights_pad = (long)v.cout * p.n_tile * n_tiles * g.taps;
    const int units = p.tapn ? g.taps / 9 : (p.reuse ? g.taps / 3 : g.taps);          // what a CTA's tap group is made of
    p.swap = !p.tapn && p.reuse && p.b_groups == 1 && getenv("CHAP_WG_NO_SWAP") == nullptr;
    p.n_mma = p.m_tile < 16 ? 16 : p.m_tile;
    const int cols_per_unit = p.tapn ? p.b_groups * p.n_cols : (p.swap ? p.n_mma : (p.reuse ? 3 : 1) * p.n_tile);
    // (measured with the 128-bit reduction path: 4 M elements is still the best budget, 8 M / 16 M are 5-15 % slower)
    const long budget = getenv("CHAP_WG_BUDGET") ? atol(getenv("CHAP_WG_BUDGET")) : 4000000L;
    long max_splits = budget / weights_pad;
    if (max_splits < 1) max_splits = 1;
    if (max_splits > p.blocks_total) max_splits = p.blocks_total;
    int best_tg = 1, best_ctas = -1, best_splits = 1, best_groups = units;
    for (int tg = (512 / cols_per_unit < units ? 512 / cols_per_unit : units); tg >= 1; --tg) {
        const int groups = (units + tg - 1) / tg;
        if ((units + groups - 1) / groups != tg) continue;                     // keep the groups balanced
        const bool two = tg * cols_per_unit <= 256 && a_ring + 3 * stage <= 100 * 1024;
        const int target = two ? 2 * kNumSMs : kNumSMs;
        long splits = (target + groups * zdim - 1) / (groups * zdim);
        if (splits > max_splits) splits = max_splits;
        const int ctas = (int)(splits * groups * zdim);
        const int score = ctas >= (target * 3) / 4 ? 1000000 + tg : ctas;      // enough CTAs: prefer the largest tap group
        if (score > best_ctas) { best_ctas = score; best_tg = tg; best_splits = (int)splits; best_groups = groups; }
    }
    const int tg = best_tg, groups = best_groups;
    int splits = best_splits;
    p.tg = tg;
    p.tmem_cols = 32; while (p.tmem_cols < tg * cols_per_unit) p.tmem_cols *= 2;
    p.acc2 = p.swap && p.P % 64 == 0 && p.tmem_cols <= 256 && getenv("CHAP_WG_ACC2") != nullptr;
    if (p.acc2) p.tmem_cols *= 2;
    const bool two_per_sm = p.tmem_cols <= 256 && a_ring + 3 * stage <= 100 * 1024;
    int stages = (int)(((two_per_sm ? 100 : 200) * 1024 - 2048 - a_ring) / stage);
    if (stages > 8) stages = 8;
    CHAP_REQUIRE(stages >= 2, CHAP_ERR_BAD_ARG, "tc_wgrad: tile does not fit shared memory");
    p.stages = stages;
    p.blocks_per_cta = (p.blocks_total + splits - 1) / splits;
    splits = (p.blocks_total + p.blocks_per_cta - 1) / p.blocks_per_cta;
    const PackSpec ps = fwd_pack(g);          // torch [co][ci][tap]: s_ci = T, s_co = Cin*T; transposed [ci][co][tap]: M = ci
    p.s_ci = up2 ? ps.sn : ps.sk; p.s_co = up2 ? ps.sk : ps.sn;
    p.dw = dw;

    CUtensorMap tmA, tmB;
    static thread_local TmTaps tmT;                 // filled (and read by the kernel) only for the k2 s2 convolutions
    for (int which = 0; which < 2; ++which) {
        const float* base = which == 0 ? m_src : n_src;
        const uint64_t C = which == 0 ? v.cout : v.cin;
        const int cpg = which == 0 ? p.a_cpg : p.b_cpg;
        uint64_t dims[5], str[4]; uint32_t box[5];
        if (k2 && which == 1) {
            const uint64_t bw = 2 * (uint64_t)p.W, bh = 2 * (uint64_t)p.H;
            for (int t = 0; t < g.taps; ++t) {
                const int kw = t & 1, kh = (t >> 1) & 1, kd = t >> 2;
                const float* tb = base + (((uint64_t)kd * bh + kh) * bw + kw) * C;
                if (g.nd == 2) {
                    dims[0] = C; dims[1] = p.W; dims[2] = p.H; dims[3] = g.n;
                    str[0] = 2 * C * 4; str[1] = 2 * bw * C * 4; str[2] = bh * bw * C * 4;
                    box[0] = cpg; box[1] = p.tw; box[2] = p.th; box[3] = 1;
                    CHAP_TRY(make_tensor_map(&tmT.m[t], tb, 4, dims, str, box, cpg, true));
                } else {
                    dims[0] = C; dims[1] = p.W; dims[2] = p.H; dims[3] = p.D; dims[4] = g.n;
                    str[0] = 2 * C * 4; str[1] = 2 * bw * C * 4; str[2] = 2 * bh * bw * C * 4; str[3] = 2 * (uint64_t)p.D * bh * bw * C * 4;
                    box[0] = cpg; box[1] = p.tw; box[2] = p.th; box[3] = p.td; box[4] = 1;
                    CHAP_TRY(make_tensor_map(&tmT.m[t], tb, 5, dims, str, box, cpg, true));
                }
            }
            tmB = tmT.m[0];
        } else if (g.nd == 2) {
            dims[0] = C; dims[1] = p.W; dims[2] = p.H; dims[3] = g.n;
            str[0] = C * 4; str[1] = str[0] * p.W; str[2] = str[1] * p.H;
            box[0] = cpg; box[1] = p.tw; box[2] = (which == 1 && p.reuse) ? p.th + 2 : p.th; box[3] = 1;
            CHAP_TRY(make_tensor_map(which == 0 ? &tmA : &tmB, base, 4, dims, str, box, cpg, true));
        } else {
            dims[0] = C; dims[1] = p.W; dims[2] = p.H; dims[3] = p.D; dims[4] = g.n;
            str[0] = C * 4; str[1] = str[0] * p.W; str[2] = str[1] * p.H; str[3] = str[2] * p.D;
            box[0] = cpg; box[1] = p.tw; box[2] = (which == 1 && p.reuse) ? p.th + 2 : p.th; box[3] = p.td; box[4] = 1;
            CHAP_TRY(make_tensor_map(which == 0 ? &tmA : &tmB, base, 5, dims, str, box, cpg, true));
        }
    }
    static std::once_flag attr_once;
    std::call_once(attr_once, [] { cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(wgrad_tc_kernel_taps, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); });
    const size_t smem = 1024 + a_ring + (size_t)stages * stage + (2 * stages + 1 + 2 * p.a_stages) * sizeof(uint64_t) + 16 +
                        ((p.swap || p.tapn) ? (size_t)p.tw * 128 + 1024 : 0);      // x on M: the junk 4th M chunk reads tw rows past the last x box
    // non-swap mode with scratch: vector reductions into [tap][M][N], then one transposing copy into the torch layout
    // (taps-in-N: the scratch is [tap][cin][cout] -- the accumulator columns run along cout)
    p.acc = (!p.swap && acc_ws && aligned16(acc_ws) && (p.tapn ? v.cout : v.cin) % 16 == 0 && getenv("CHAP_WG_NO_V4") == nullptr) ? acc_ws : nullptr;
    // the kernel ADDS (red.global.add) into its target: a fresh gradient needs it zeroed; in accumulate mode dw keeps its content
    // and the scratch arrives zeroed (the unpack kernel of its previous use cleared it)
    if (!accumulate) CHAP_TRY(zero_async(p.acc ? p.acc : dw, (size_t)g.taps * v.cin * v.cout * sizeof(float), st));
    const double rows = (double)(up2 ? g.in_rows : g.out_rows);     // = pixels of the low-resolution grid for the k2 s2 kinds
    KernelTimer timer(timer_name("conv_tc_wgrad", g.taps, v.cin, v.cout, g.iW, g.iH, g.iD, g.in_rows), 2.0 * rows * v.cin * v.cout * g.taps,
                      4.0 * (rows * v.cin + rows * v.cout + (double)g.taps * v.cin * v.cout), st);
    dim3 grid((unsigned)splits, (unsigned)groups, (unsigned)zdim);
    if (p.k2s2) launch_k(wgrad_tc_kernel_taps, grid, kWgThreads, smem, st, tmA, tmB, tmT, p);
    else launch_k(wgrad_tc_kernel, grid, kWgThreads, smem, st, tmA, tmB, p);
    CHAP_TRY(launched("wgrad_tc_kernel"));
    if (p.acc) {
        const int64_t total = (int64_t)g.taps * v.cin * v.cout;
        if (p.tapn) launch_k(wgrad_unpack_kernel, grid_for(total, 256 * 2, kNumSMs * 4), 256, 0, st, p.acc, dw, g.taps, v.cin, v.cout, p.s_ci, p.s_co, accumulate ? 1 : 0);
        else launch_k(wgrad_unpack_kernel, grid_for(total, 256 * 2, kNumSMs * 4), 256, 0, st, p.acc, dw, g.taps, v.cout, v.cin, p.s_co, p.s_ci, accumulate ? 1 : 0);
        CHAP_TRY(launched("wgrad_unpack_kernel"));
    }
    return 1;
}

}  // namespace chap
