// tcgen05 convolution for THIN k=3 layers (Cin, Cout in {16, 32}; forward and data gradient; 2D and 3D).
//
// Measured on B200: a tcgen05.mma (M = 128, kind::tf32, K = 8) costs ~130 cycles whatever N is.  The ordinary implicit
// GEMM (conv_tc.cu) needs taps * Cin / 8 = 18 MMAs with N = 16 per 128 output pixels for a 16 -> 16 layer and is purely
// MMA-issue bound there (6 % of the tensor datapath width used).  This kernel moves the TAPS into the N dimension:
//
//   P[q, (ky, kx, co)] = sum_ci X[q, ci] * W[ky, kx, co, ci]          one GEMM, M = 128 INPUT pixels q of a 16 x 8 tile
//                                                                     (with halo), N = 9 * Cout (144 / 288), K = Cin
//   out[p, co]         = sum_{ky,kx} P[p + (ky-1, kx-1), (ky, kx, co)]  spatial shift-add over the 9 partials
//
// The tensor core does the channel contraction for all nine taps in Cin/8 * ceil(9 Cout / 256) = 2..8 MMAs; the partial
// products go TMEM -> shared memory ([tap][pixel][16 channels]) and the epilogue warps gather the 9 shifted partials of
// every interior pixel (the 14 x 6 outputs of the tile), add bias, store 128-bit and reduce the BatchNorm statistics.
// 3D: the same per kz plane (input plane d + kz - 1, weights of that kz), accumulated in registers over the 3 planes.
// Zero padding = TMA out-of-bounds fill, as in conv_tc.cu.  The weights operand is the same packed [tap][Cout][Cin]
// buffer: its rows [kz*9*Cout, (kz+1)*9*Cout) ARE the B matrix.
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include "common.cuh"
#include "conv_plan.cuh"
#include "tc_common.cuh"

namespace chap {

constexpr int kIW = 16, kIH = 8;            // input tile (with halo) = 128 pixels = one M tile
constexpr int kOW = kIW - 2, kOH = kIH - 2; // output tile 14 x 6

struct ThinParams {
    int nd, W, H, D, n_img;
    int tiles_w, tiles_h;
    int cin, cout;                 // K and output channels of THIS op (dgrad: swapped by the caller)
    int nkz;                       // 1 (2D) or 3 (3D)
    int stages, tmem_cols;
    uint32_t a_stage_bytes, b_stage_bytes;
    float* out;
    const float* bias;
    double* stats;
};

constexpr int kThinThreads = 192;

__device__ __forceinline__ uint64_t thin_kmajor_desc(uint32_t saddr, uint32_t row_bytes) {
    const uint64_t sbo = (8u * row_bytes) >> 4;
    const uint64_t layout = row_bytes == 128 ? 2ull : 4ull;
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

__global__ void __launch_bounds__(kThinThreads)
conv_thin_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ThinParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_base = smem;
    uint8_t* b_base = a_base + (size_t)p.stages * p.a_stage_bytes;
    float* S = reinterpret_cast<float*>(b_base + (size_t)p.stages * p.b_stage_bytes);      // [9 taps][4 quads][128 px][4] partials
    uint64_t* bars = reinterpret_cast<uint64_t*>(S + 9 * 128 * 16);
    uint64_t* full = bars;
    uint64_t* empty = bars + p.stages;
    uint64_t* tmem_full = bars + 2 * p.stages;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
    float* red = reinterpret_cast<float*>(tmem_slot + 2);                                   // [4 warps][2][cout]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x;
    const int tx = t % p.tiles_w; t /= p.tiles_w;
    const int ty = t % p.tiles_h; t /= p.tiles_h;
    const int od = t % p.D;
    const int img = t / p.D;
    const int ow0 = tx * kOW, oh0 = ty * kOH;
    const uint32_t row_bytes = (uint32_t)p.cin * 4u;
    const int nrows_b = 9 * p.cout;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int kz = 0; kz < p.nkz; ++kz) {
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], 128u * row_bytes + (uint32_t)nrows_b * row_bytes);
                uint8_t* a_dst = a_base + (size_t)s * p.a_stage_bytes;
                uint8_t* b_dst = b_base + (size_t)s * p.b_stage_bytes;
                if (p.nd == 2) tma_load_4d(a_dst, &tmA, &full[s], 0, ow0 - 1, oh0 - 1, img);
                else tma_load_5d(a_dst, &tmA, &full[s], 0, ow0 - 1, oh0 - 1, od + kz - 1, img);
                for (int r0 = 0; r0 < nrows_b; r0 += 144)            // boxes of 144 rows (<= 256, multiple of 8)
                    tma_load_2d(b_dst + (size_t)r0 * row_bytes, &tmB, &full[s], 0, kz * nrows_b + r0);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // D = F32, A = B = TF32, K-major both, N = 144, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(144 >> 3) << 17) | ((128u >> 4) << 24);
            const int ksteps = p.cin / 8;
            int s = 0; uint32_t ph = 0;
            for (int kz = 0; kz < p.nkz; ++kz) {
                mbar_wait(&full[s], ph);
                if (kz > 0) mbar_wait(tmem_empty, (uint32_t)((kz - 1) & 1));       // epilogue has drained the previous plane
                tc_fence_after();
                const uint32_t a_addr = smem_u32(a_base + (size_t)s * p.a_stage_bytes);
                const uint32_t b_addr = smem_u32(b_base + (size_t)s * p.b_stage_bytes);
                for (int n0 = 0; n0 < nrows_b; n0 += 144) {
                    const uint64_t a_desc = thin_kmajor_desc(a_addr, row_bytes);
                    const uint64_t b_desc = thin_kmajor_desc(b_addr + (uint32_t)n0 * row_bytes, row_bytes);
                    for (int k = 0; k < ksteps; ++k)
                        tc_mma_tf32(tmem_base + (uint32_t)n0, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (uint32_t)(k != 0));
                }
                tc_commit(&empty[s]);
                tc_commit(tmem_full);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: TMEM -> smem -> shift-add
        const int lg = warp & 3;
        const int q = lg * 32 + lane;                       // input pixel of the tile handled by this thread
        const int ix = q % kIW, iy = q / kIW;
        const int ox = ow0 + ix - 1, oy = oh0 + iy - 1;
        const bool interior = ix >= 1 && ix <= kOW && iy >= 1 && iy <= kOH;
        const bool valid = interior && ox < p.W && oy < p.H;
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.f;
        const int rounds = p.cout / 16;
        for (int kz = 0; kz < p.nkz; ++kz) {
            mbar_wait(tmem_full, (uint32_t)(kz & 1));
            tc_fence_after();
            for (int r = 0; r < rounds; ++r) {
                // phase 1: this pixel's partials for 16 output channels of all 9 taps -> S[tap][q][16]
#pragma unroll 1
                for (int tp = 0; tp < 9; ++tp) {
                    float v[16];
                    tc_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(tp * p.cout + r * 16), v);
                    // layout S[tap][4 channel quads][128 pixels][4]: consecutive pixels (threads) are 16 B apart -> conflict-free
                    float4* dst = reinterpret_cast<float4*>(S) + (size_t)tp * 4 * 128 + q;
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j * 128] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                if (r == rounds - 1) {                       // TMEM fully read: the MMA warp may overwrite it with the next plane
                    tc_fence_before();
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(tmem_empty)) : "memory");
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                // phase 2: gather the 9 shifted partials
                if (interior) {
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const int qq = q + (ky - 1) * kIW + (kx - 1);
                            const float4* src = reinterpret_cast<const float4*>(S) + (size_t)(ky * 3 + kx) * 4 * 128 + qq;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 u = src[j * 128];
                                if (r == 0) { acc[4 * j] += u.x; acc[4 * j + 1] += u.y; acc[4 * j + 2] += u.z; acc[4 * j + 3] += u.w; }
                                else { acc[16 + 4 * j] += u.x; acc[16 + 4 * j + 1] += u.y; acc[16 + 4 * j + 2] += u.z; acc[16 + 4 * j + 3] += u.w; }
                            }
                        }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");      // S is rewritten by the next round / plane
            }
        }
        // bias, store, BatchNorm statistics
        float* dst = p.out + ((((int64_t)img * p.D + od) * p.H + oy) * p.W + ox) * (int64_t)p.cout;
        float* red_s = red + (size_t)(lg * 2 + 0) * p.cout;
        float* red_q = red + (size_t)(lg * 2 + 1) * p.cout;
        for (int r = 0; r < rounds; ++r) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (r == 0 ? acc[j] : acc[16 + j]) + (p.bias ? __ldg(p.bias + r * 16 + j) : 0.f);
            if (valid) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(dst + r * 16 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            if (p.stats) {
                float s16[16], q16[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { float x = valid ? v[j] : 0.f; s16[j] = x; q16[j] = x * x; }
#pragma unroll
                for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
                    const bool upper = (lane & bit) != 0;
#pragma unroll
                    for (int j = 0; j < half; ++j) {
                        float keep_s = upper ? s16[j + half] : s16[j], send_s = upper ? s16[j] : s16[j + half];
                        float keep_q = upper ? q16[j + half] : q16[j], send_q = upper ? q16[j] : q16[j + half];
                        s16[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, bit);
                        q16[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, bit);
                    }
                }
                float cs = s16[0] + __shfl_xor_sync(0xffffffffu, s16[0], 1);
                float cq = q16[0] + __shfl_xor_sync(0xffffffffu, q16[0], 1);
                if ((lane & 1) == 0) {
                    const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                    red_s[r * 16 + col] = cs; red_q[r * 16 + col] = cq;
                }
            }
        }
        if (p.stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int e = threadIdx.x - 64;
            double* slot = p.stats + (size_t)(blockIdx.x % CHAP_STAT_SLOTS) * 2 * p.cout;
            for (int c = e; c < p.cout; c += 128) {
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) { a += red[(size_t)(w * 2) * p.cout + c]; b += red[(size_t)(w * 2 + 1) * p.cout + c]; }
                atomicAdd(slot + c, (double)a);
                atomicAdd(slot + p.cout + c, (double)b);
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

bool thin_tc_supports(const Geom& g, bool dgrad) {
    if (g.kind != CHAP_CONV_K3) return false;
    const int K = dgrad ? g.cout : g.cin, N = dgrad ? g.cin : g.cout;
    if (!((K == 16 || K == 32) && (N == 16 || N == 32))) return false;
    if (g.iW < kIW || g.iH < kIH) return false;                 // the TMA box must fit inside the tensor extents
    // Measured (16 -> 16 @ 256^2, b12): 147 us vs 73 us for conv_tc.cu -- 9x fewer MMAs, but one 84-output tile per CTA
    // pays the TMEM->smem->register shift-add (73 KB through a 64 B/clk TMEM read port) and the per-CTA latency chain
    // with only 2 CTAs/SM.  Needs a persistent tile loop with resident weights to pay off: opt-in until then.
    static const bool on = getenv("CHAP_THIN_TC") != nullptr;
    return on;
}

int thin_tc_conv(const Geom& g, bool dgrad, const float* in, const float* wp, const float* bias, float* out,
                 double* ch_sums, cudaStream_t st) {
    if (!thin_tc_supports(g, dgrad)) return 0;
    CHAP_REQUIRE(aligned16(in) && aligned16(wp) && aligned16(out), CHAP_ERR_ALIGNMENT, "thin_tc_conv: buffers must be 16-byte aligned");
    const int K = dgrad ? g.cout : g.cin, N = dgrad ? g.cin : g.cout;
    ThinParams p{};
    p.nd = g.nd; p.W = g.iW; p.H = g.iH; p.D = g.iD; p.n_img = g.n;
    p.tiles_w = (p.W + kOW - 1) / kOW; p.tiles_h = (p.H + kOH - 1) / kOH;
    p.cin = K; p.cout = N; p.nkz = g.nd == 3 ? 3 : 1;
    p.tmem_cols = 9 * N <= 256 ? 256 : 512;
    p.a_stage_bytes = 128u * K * 4u;
    p.b_stage_bytes = ((uint32_t)(9 * N) * K * 4u + 1023u) & ~1023u;
    p.stages = g.nd == 3 ? 2 : 1;
    p.out = out; p.bias = bias; p.stats = ch_sums;
    const size_t smem = 1024 + (size_t)p.stages * (p.a_stage_bytes + p.b_stage_bytes) + (size_t)9 * 128 * 16 * 4 +
                        (2 * p.stages + 2) * sizeof(uint64_t) + 16 + (size_t)8 * N * sizeof(float);

    CUtensorMap tmA, tmB;
    {
        uint64_t dims[5], str[4]; uint32_t box[5];
        const uint64_t C = (uint64_t)K;
        if (g.nd == 2) {
            dims[0] = C; dims[1] = p.W; dims[2] = p.H; dims[3] = g.n;
            str[0] = C * 4; str[1] = str[0] * p.W; str[2] = str[1] * p.H;
            box[0] = K; box[1] = kIW; box[2] = kIH; box[3] = 1;
            CHAP_TRY(make_tensor_map(&tmA, in, 4, dims, str, box, K));
        } else {
            dims[0] = C; dims[1] = p.W; dims[2] = p.H; dims[3] = p.D; dims[4] = g.n;
            str[0] = C * 4; str[1] = str[0] * p.W; str[2] = str[1] * p.H; str[3] = str[2] * p.D;
            box[0] = K; box[1] = kIW; box[2] = kIH; box[3] = 1; box[4] = 1;
            CHAP_TRY(make_tensor_map(&tmA, in, 5, dims, str, box, K));
        }
        uint64_t wd[2] = {(uint64_t)K, (uint64_t)g.taps * N};
        uint64_t ws[1] = {(uint64_t)K * 4};
        uint32_t wb[2] = {(uint32_t)K, 144u};
        CHAP_TRY(make_tensor_map(&tmB, wp, 2, wd, ws, wb, K));
    }
    static std::once_flag attr_once;
    std::call_once(attr_once, [] { cudaFuncSetAttribute(conv_thin_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); });
    if (ch_sums) CHAP_TRY(zero_async(ch_sums, (size_t)CHAP_STAT_SLOTS * 2 * N * sizeof(double), st));
    const double rows = (double)g.out_rows;
    KernelTimer timer(dgrad ? "conv_thin_tc_dgrad" : "conv_thin_tc_fwd", 2.0 * rows * K * N * g.taps,
                      4.0 * (rows * K + rows * N + (double)g.taps * K * N), st);
    dim3 grid((unsigned)(g.n * p.D * p.tiles_h * p.tiles_w));
    conv_thin_tc_kernel<<<grid, kThinThreads, smem, st>>>(tmA, tmB, p);
    CHAP_TRY(launched("conv_thin_tc_kernel"));
    return 1;
}

}  // namespace chap
