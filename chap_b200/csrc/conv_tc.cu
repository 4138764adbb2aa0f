// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.
//
// Forward and data-gradient of the stride-1 convolutions (k=3 pad 1, k=1; 2D and 3D) as
//     D[m, n] = sum_{tap} sum_{k} A_tap[m, k] * B_tap[n, k]
//   m : 128 output positions of one spatial box (tw x th x td) of one image      -> TMEM lanes
//   n : up to 256 output channels                                                -> TMEM columns (fp32 accumulators)
//   k : input channels in chunks of 32 (128-byte rows, SWIZZLE_128B) or 16 (64-byte rows, SWIZZLE_64B)
// A_tap is the channels-last activation box shifted by the tap offset: ONE tiled TMA load per (tap, k-chunk) with
// the box origin moved by (kx-1, ky-1, kz-1); out-of-bounds coordinates are zero-filled by the TMA unit, which IS the
// convolution's zero padding -- no im2col buffer, no halo handling in the kernel.  B_tap is the packed weight
// [tap][n][k] (K-major).  Operands are fp32 in memory, converted to TF32 (round-to-nearest) by the TMA unit
// (CU_TENSOR_MAP_DATA_TYPE_TFLOAT32) and multiplied by tcgen05.mma.kind::tf32 with fp32 accumulation in TMEM.
//
// Also: the transposed k2 s2 convolution forward (one GEMM with N = taps * Cout and a scatter epilogue, 2D and 3D) and its 2D
// data gradient (4-tap gather of dy through a [2C, W, 2, H, N] tensor map); forward of the 4- / 8-channel heads (N zero-padded
// to 16 in the packed weight); the data gradient of conv(cat(a, b)) with the two parts written by the epilogue.
//
// Persistent CTA (<= 2 per SM) = 10 warps: warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator (both run their loops
// warp-uniformly, one elect.sync lane issues), two groups of 4 epilogue warps (tcgen05.ld pipelined against the processing of
// the previous 16-column chunk -> +bias -> 256-bit stores; per-channel sum / sum-of-squares of the following BatchNorm through
// a padded shared-memory transpose, accumulated per CTA and added with one double atomic per channel into one of
// CHAP_STAT_SLOTS slots).  The smem ring of {A, B} stages (full/empty mbarriers; 1, 2 or 4 k-chunks per stage) and TWO TMEM
// accumulators (tmem_full / tmem_empty mbarriers) run across tile boundaries: loads and MMAs of tile j + 1 overlap the
// epilogue of tile j.  Weights stay resident in shared memory when they fit (<= 40 KB).  Row-reuse mode (N <= 64, k3): one
// h-haloed A box per (kz, kx) serves the three ky taps through descriptor row offsets.
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include <unordered_map>
#include "common.cuh"
#include "conv_plan.cuh"
#include "tc_common.cuh"

namespace chap {

struct TcParams {
    int nd, ksz, pad, taps;
    int W, H, D;                  // spatial size (output == input for these kinds)
    int tw, th, td;               // spatial box of one M tile (tw*th*td <= 128)
    int tiles_w, tiles_h, tiles_d, tiles_total;
    int kc, kchunks;              // channels per k chunk (32 or 16), chunks per tap
    int n_total, nt;              // output channels, channels per CTA
    int stages, tmem_cols;
    int n_buf, n_buf_lg;          // TMEM accumulators (1, 2 or 4: the epilogue of tile j overlaps the MMAs of tiles j + 1 ..) and log2 of it
    int reuse;                    // 1: one h-haloed A box per (kz, kx) serves the three ky taps (row-offset descriptors)
    int n_real;                   // output channels that exist (< nt = 16 for the zero-padded 4- / 8-channel heads)
    int mode;                     // 0: stride-1 conv; 1: k2 s2 scatter (transposed conv forward, strided conv data gradient): one GEMM over
                                  // the low-resolution grid, N = taps * C_big, scatter epilogue; 2: 2D k2 s2 gather through the
                                  // [2C, W, 2, H, N] view (transposed conv data gradient); 3: k2 s2 gather with one tensor map per tap
                                  // (strided conv forward 2D / 3D, transposed conv data gradient 3D).  W, H, D = low-resolution grid.
    int up_c;                     // mode 1: channels of the high-resolution tensor (columns per tap); mode 2: channels of dy
    int epi_groups;               // 1 or 2 groups of 4 epilogue warps (blockDim = 64 + 128 * groups)
    int colsplit;                 // two groups: 1 = both work on every tile, half the columns each; 0 = they alternate tiles
    int cps;                      // k chunks per pipeline stage (1, 2 or 4): fewer, fatter stages for the deep layers
    int thin2d;                   // 1: row-reuse layer (2D or 3D) with one k chunk and resident weights -> the lean issue / producer loops
    int sup;                      // 1 (thin2d only): SUPER TILES -- one haloed box of 2 th + 2 rows per (kz, kx) serves TWO vertically adjacent
                                  // M tiles (rows 0.. and th..), one per TMEM accumulator: half the TMA issues / ring round trips per tile and
                                  // 1.25x instead of 1.5x halo overhead.  tiles_h / tiles_total then count super tiles.
    int kxn;                      // 1 (thin2d only): the three kx taps sit in the N dimension -- ONE haloed box (th + 2 rows x 32 pixels) per (kz)
                                  // and tile, 3 x KSTEPS MMAs with N = 3 nt (weight boxes of taps (ky, 0..2) are contiguous rows), and the
                                  // epilogue adds the three column groups shifted by one pixel (warp = one image row of the tile, lane = x:
                                  // out[x] = D0[x - 1] + D1[x] + D2[x + 1] through two shuffles).  Lanes 0 and 31 are halo: a tile yields
                                  // step_w = 30 output pixels per row.  A third of the tcgen05.mma count and of the TMA boxes.
    int step_w;                   // output pixels per tile row (tw, or 30 with kxn)
    int debug;                    // CHAP_TC_DEBUG bit mask (only with -DCHAP_TC_DEBUG_HOOKS)
    int b_resident;               // 1: all weight boxes [tap][kchunk] are loaded once per CTA and stay in shared memory
    int precise;                  // 1: split-operand 3xTF32 (kernel template PRECISE): 4 extra warps split every A stage into TF32 hi / lo halves
    int b_lo_rows;                // precise: row offset of the lo half inside the weight tensor map
    uint32_t a_lo_off, b_lo_off;  // precise: byte offsets of the lo copies of the A ring / the B area (or B ring) in shared memory
    uint32_t a_stage_bytes, b_stage_bytes, a_chunk_bytes, b_chunk_bytes, a_box_bytes, b_box_bytes, b_area_bytes;
    float* out;                   // channels [0, ca), row stride ca
    float* out_b;                 // channels [ca, n_total), row stride n_total - ca (dgrad of a channel concat), or nullptr
    int ca;
    const float* bias;
    const float* epi_ss;          // inference epilogue (eval-mode BatchNorm folded in): y = act(scale[c] * (conv + bias) + shift[c]) + residual;
    const float* epi_res;         //   epi_ss = [scale[C], shift[C]] or nullptr (off); epi_res = tensor of the output's shape or nullptr
    float epi_slope;              //   negative slope of the (Leaky)ReLU
    int epi_c;                    //   C = real output channels (offset of the shift half)
    double* stats;                // [CHAP_STAT_SLOTS][2 * n_total] or nullptr
    BnFold bn;                    // bn.mi != nullptr: the last CTA turns the statistics into BatchNorm scale / shift
};

constexpr int kTcThreadsMax = 320;         // TMA warp, MMA warp, 1 or 2 groups of 4 epilogue warps
constexpr int kTcThreadsPrecise = 448;     // + 4 operand-splitting warps (split-operand 3xTF32 mode)
// Profiling experiments (CHAP_TC_DEBUG bit mask: 1 no MMAs, 2 no A loads, 4 no stores / statistics, 8 no tcgen05.ld,
// 16 polling waits, 32 plain arrives instead of commits) are compiled in only with -DCHAP_TC_DEBUG_HOOKS: the
// single-warp issue loops are latency-bound and every extra branch costs.
#ifdef CHAP_TC_DEBUG_HOOKS
// timeline of CTA 0 (SM clock): 0 entry, 1 prologue done, 2 first smem stage full, 3 first accumulator full, 4 first tile stored,
// 5 last tile stored, 6 statistics written, 7 before exit; read back with chap_debug_tc_trace()
__device__ long long g_tc_trace[8];
#define TC_TRACE(i) do { if (blockIdx.x == 0 && blockIdx.y == 0) g_tc_trace[i] = clock64(); } while (0)
#define TC_DBG(bit) (p.debug & (bit))
#define TC_WAIT(bar, par) do { if (p.debug & 16) mbar_poll(bar, par); else mbar_wait(bar, par); } while (0)
#define TC_COMMIT(bar) do { if (p.debug & 32) mbar_arrive(bar); else tc_commit(bar); } while (0)
#define TC_DBG_HOST_OFF (p.debug == 0)
#else
#define TC_DBG_HOST_OFF true
#define TC_TRACE(i) do { } while (0)
#define TC_DBG(bit) false
#define TC_WAIT(bar, par) mbar_wait(bar, par)
#define TC_COMMIT(bar) tc_commit(bar)
#endif

// Column sums (and sums of squares) of a 32-row x 16-column register tile (row = lane) through a warp-private padded
// shared-memory scratch: 16 conflict-free stores and 16 conflict-free loads per lane instead of a 62-shuffle butterfly
// (the epilogue is one latency-bound instruction stream per warp, so instruction count is what matters).  Lane l sums
// column l % 16 over rows 16 * (l / 16) ..; the two halves meet in one shuffle and lanes 0..15 add into dst (one owner
// per column and warp: no race).
// The scratch is addressed through explicit shared-space instructions: behind the generic `red` pointer the compiler emitted generic
// LD.E / ST.E (ncu source view of round 2: 42 % of the epilogue warps' samples sat on this function, mostly on those loads).
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ void warp_column_sums(const float (&v)[16], const bool valid, const int lane, float* scratch,
                                                 float* dst_s, float* dst_q) {
    const uint32_t sc = smem_u32(scratch);
    const uint32_t wr = sc + (uint32_t)(lane * 17) * 4u;
#pragma unroll
    for (int j = 0; j < 16; ++j) sts_f32(wr + 4u * j, valid ? v[j] : 0.f);
    __syncwarp();
    const int col = lane & 15, r0 = (lane >> 4) * 16;
    const uint32_t rd = sc + (uint32_t)(r0 * 17 + col) * 4u;
    float x[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = lds_f32(rd + (uint32_t)(r * 17) * 4u);
    // two independent chains per statistic instead of one 16-deep dependent chain
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int r = 0; r < 16; r += 2) { s0 += x[r]; q0 = fmaf(x[r], x[r], q0); s1 += x[r + 1]; q1 = fmaf(x[r + 1], x[r + 1], q1); }
    float s = s0 + s1, q = q0 + q1;
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    q += __shfl_xor_sync(0xffffffffu, q, 16);
    if (lane < 16) {
        const uint32_t ds = smem_u32(dst_s + col), dq = smem_u32(dst_q + col);
        sts_f32(ds, lds_f32(ds) + s); sts_f32(dq, lds_f32(dq) + q);
    }
    __syncwarp();
}

// Walks tiles first, first + step, ... of the (w, h, d, image) tile grid without a division per tile: the step is
// decomposed once into mixed-radix digits and added with carries.
struct TileIter {
    int tx, ty, tz, img, tile;
    int sx, sy, sz, simg, step;
    __device__ __forceinline__ void init(const int first, const int step_, const int tiles_w, const int tiles_h, const int tiles_d) {
        tile = first; step = step_;
        int t = first;
        tx = t % tiles_w; t /= tiles_w; ty = t % tiles_h; t /= tiles_h; tz = t % tiles_d; img = t / tiles_d;
        t = step_;
        sx = t % tiles_w; t /= tiles_w; sy = t % tiles_h; t /= tiles_h; sz = t % tiles_d; simg = t / tiles_d;
    }
    __device__ __forceinline__ void next(const int tiles_w, const int tiles_h, const int tiles_d) {
        tile += step;
        tx += sx; int c = tx >= tiles_w; tx -= c ? tiles_w : 0;
        ty += sy + c; c = ty >= tiles_h; ty -= c ? tiles_h : 0;
        tz += sz + c; c = tz >= tiles_d; tz -= c ? tiles_d : 0;
        img += simg + c;
    }
};

// The MMA-issuing warp.  Its instruction stream is on the critical path of these small-N MMAs (a runtime modulo per MMA
// cost 30 % of the kernel), so the whole warp runs the warp-uniform loop (descriptors live in uniform registers), one
// elected lane issues, and the body is straight-line per stage: KSTEPS x NKY tcgen05.mma whose descriptors differ by
// integer adds on the lo word (+2 per 8 tf32 along K inside the swizzled row, + tw rows / one weight box per ky tap).
// Thin layers (row-reuse mode, K = 16 / 32 in one chunk, weights resident; NG = 3 (kx) ring slots per tile in 2D, 9 (kz, kx) in
// 3D): the generic loop spends ~90 SASS instructions per ring slot on index arithmetic, and that scalar stream of the single
// issuing warp IS the bound of these layers.  Here the tap structure is compile time (unrolled over the NG slots of a tile), the
// weight descriptors are constants and only the ring position (slot, phase, A address) is carried: per slot wait / fence /
// elect / 3 x KSTEPS MMAs / commit.
// One product term of the implicit GEMM.  PRECISE: x = x_hi + x_lo, w = w_hi + w_lo (TF32 halves; the x halves are written by
// the splitter warps, the w halves come packed from global memory) and x_hi w_hi + x_lo w_hi + x_hi w_lo goes into the same
// fp32 accumulator (the dropped x_lo w_lo term is 2^-22 relative).
template <bool PRECISE>
__device__ __forceinline__ void mma_term(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate,
                                         uint32_t a_lo_off, uint32_t b_lo_off) {
    tc_mma_tf32_lh(d_tmem, a_lo, hi, b_lo, hi, idesc, accumulate);
    if (PRECISE) {
        tc_mma_tf32_lh(d_tmem, a_lo + a_lo_off, hi, b_lo, hi, idesc, 1u);
        tc_mma_tf32_lh(d_tmem, a_lo, hi, b_lo + b_lo_off, hi, idesc, 1u);
    }
}

template <int KSTEPS, int NG, bool PRECISE, bool KXN = false>
__device__ __forceinline__ void issue_mmas_thin(const TcParams& p, uint8_t* a_base, uint8_t* b_base, uint64_t* full, uint64_t* empty,
                                                uint64_t* tmem_full, uint64_t* tmem_empty, uint64_t* b_full, uint32_t tmem_base) {
    const uint32_t n_mma = KXN ? 3u * (uint32_t)p.nt : (uint32_t)p.nt;          // KXN: columns = [kx][nt]
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((n_mma >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t row_bytes = KSTEPS * 32u;
    constexpr uint32_t hi = ((8u * row_bytes) >> 4) | (1u << 14) | ((row_bytes == 128 ? 2u : 4u) << 29);
    const uint32_t a_lo_off = p.a_lo_off >> 4, b_lo_off = p.b_lo_off >> 4;
    const uint32_t a0 = ((smem_u32(a_base) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t b0 = ((smem_u32(b_base) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t a_stage = p.a_stage_bytes >> 4, b_box = p.b_box_bytes >> 4;
    const uint32_t a_ky = ((uint32_t)p.tw * row_bytes) >> 4;
    mbar_wait(b_full, 0);
    int s = 0; uint32_t ph = 0, a_s = a0;
    int j = 0;
    if (p.sup) {
        // super tiles: unit k = tiles j = 2k (accumulator 0, box rows 0..) and j = 2k + 1 (accumulator 1, box rows th * tw..)
        const uint32_t a_sub = ((uint32_t)(p.th * p.tw) * row_bytes) >> 4;
        for (int unit = blockIdx.x; unit < p.tiles_total; unit += gridDim.x, ++j) {
            mbar_wait(&tmem_empty[0], (uint32_t)(j & 1) ^ 1u);
            mbar_wait(&tmem_empty[1], (uint32_t)(j & 1) ^ 1u);
            tc_fence_after();
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int kx = g % 3, kz = g / 3;                               // compile time after unrolling
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int sb = 0; sb < 2; ++sb) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)(sb * p.nt);
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                            for (int k = 0; k < KSTEPS; ++k)
                                mma_term<PRECISE>(d_tmem, a_s + (uint32_t)sb * a_sub + (uint32_t)ky * a_ky + 2u * k,
                                                  b0 + (uint32_t)((kz * 3 + ky) * 3 + kx) * b_box + 2u * k, hi, idesc,
                                                  (g | ky | k) == 0 ? 0u : 1u, a_lo_off, b_lo_off);
                        }
                    }
                    tc_commit(&empty[s]);
                }
                __syncwarp();
                a_s += a_stage;
                if (++s == p.stages) { s = 0; ph ^= 1u; a_s = a0; }
            }
            if (elect_one()) { tc_commit(&tmem_full[0]); tc_commit(&tmem_full[1]); }
            __syncwarp();
        }
        return;
    }
    for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x, ++j) {
        const int buf = j & (p.n_buf - 1);
        mbar_wait(&tmem_empty[buf], (uint32_t)((j >> p.n_buf_lg) & 1) ^ 1u);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * n_mma;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int kx = KXN ? 0 : g % 3, kz = KXN ? g : g / 3;            // compile time after unrolling
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k)
                        mma_term<PRECISE>(d_tmem, a_s + (uint32_t)ky * a_ky + 2u * k, b0 + (uint32_t)((kz * 3 + ky) * 3 + kx) * b_box + 2u * k, hi, idesc,
                                          (g | ky | k) == 0 ? 0u : 1u, a_lo_off, b_lo_off);
                }
                tc_commit(&empty[s]);
            }
            __syncwarp();
            a_s += a_stage;
            if (++s == p.stages) { s = 0; ph ^= 1u; a_s = a0; }
        }
        if (elect_one()) tc_commit(&tmem_full[buf]);
        __syncwarp();
    }
}

template <int KSTEPS, int NKY, bool PRECISE>
__device__ __forceinline__ void issue_mmas(const TcParams& p, uint8_t* a_base, uint8_t* b_base, uint64_t* full, uint64_t* empty,
                                           uint64_t* tmem_full, uint64_t* tmem_empty, uint64_t* b_full, uint32_t tmem_base, int groups) {
    // instruction descriptor: D = F32 (bit 4), A = B = TF32 (2 << 7, 2 << 10), K-major both, N >> 3 at 17, M >> 4 at 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.nt >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t row_bytes = KSTEPS * 32u;
    // descriptor hi word: SBO >> 4 (8 rows) | version 1 (bit 46) | layout (bit 61: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B)
    constexpr uint32_t hi = ((8u * row_bytes) >> 4) | (1u << 14) | ((row_bytes == 128 ? 2u : 4u) << 29);
    // descriptor lo word: start address >> 4 | LBO (= 1, unused for swizzled K-major) << 16
    const uint32_t a_lo0 = ((smem_u32(a_base) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t b_lo0 = ((smem_u32(b_base) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t a_stage = p.a_stage_bytes >> 4, b_stage = p.b_stage_bytes >> 4, b_box = p.b_box_bytes >> 4;
    const uint32_t a_chunk = p.a_chunk_bytes >> 4, b_chunk = p.b_chunk_bytes >> 4;   // one k chunk inside a stage
    const uint32_t a_lo_off = p.a_lo_off >> 4, b_lo_off = p.b_lo_off >> 4;
    const uint32_t a_ky = ((uint32_t)p.tw * row_bytes) >> 4;
    const uint32_t b_ky = p.b_resident ? 3u * (uint32_t)p.kchunks * b_box : b_box;     // resident layout is [tap][kchunk]
    const uint32_t b_step = p.b_resident ? b_box : b_chunk;                            // next k chunk of the same tap
    const int stage_iters = p.kchunks / p.cps;
    if (p.b_resident) TC_WAIT(b_full, 0);
    int s = 0; uint32_t ph = 0;
    uint32_t a_lo = a_lo0, b_lo_s = b_lo0;
    int j = 0;
    for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x, ++j) {
        const int buf = j & (p.n_buf - 1);
        const uint32_t use = (uint32_t)(j >> p.n_buf_lg);
        TC_WAIT(&tmem_empty[buf], (use & 1u) ^ 1u);          // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.nt);
        uint32_t accum = 0;
        for (int grp = 0; grp < groups; ++grp) {
            // reuse mode: grp = (kz, kx) and the stage serves taps ((kz * 3 + ky) * 3 + kx), ky = 0..2
            const int tap0 = NKY == 3 ? (grp / 3) * 9 + (grp % 3) : grp;
            uint32_t b_res = b_lo0 + (uint32_t)(tap0 * p.kchunks) * b_box;             // resident weights of this tap, k chunk 0
            for (int it = 0; it < stage_iters; ++it) {
                TC_WAIT(&full[s], ph);
                if (j == 0 && grp == 0 && it == 0 && (threadIdx.x & 31) == 0) TC_TRACE(2);
                tc_fence_after();
                if (elect_one()) {
                    uint32_t a_c = a_lo, b_c = p.b_resident ? b_res : b_lo_s;
                    for (int c = 0; c < p.cps && !TC_DBG(1); ++c) {
#pragma unroll
                        for (int ky = 0; ky < NKY; ++ky) {
#pragma unroll
                            for (int k = 0; k < KSTEPS; ++k)
                                mma_term<PRECISE>(d_tmem, a_c + (uint32_t)ky * a_ky + 2u * k, b_c + (uint32_t)ky * b_ky + 2u * k, hi, idesc,
                                                  (ky | k) == 0 ? accum : 1u, a_lo_off, b_lo_off);
                        }
                        accum = 1;
                        a_c += a_chunk; b_c += b_step;
                    }
                    TC_COMMIT(&empty[s]);
                }
                __syncwarp();
                accum = 1;
                b_res += (uint32_t)p.cps * b_box;
                a_lo += a_stage; b_lo_s += b_stage;
                if (++s == p.stages) { s = 0; ph ^= 1; a_lo = a_lo0; b_lo_s = b_lo0; }
            }
        }
        if (elect_one()) TC_COMMIT(&tmem_full[buf]);
        __syncwarp();
    }
}

// Persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ... of output-channel chunk blockIdx.y.  The smem ring and the
// two TMEM accumulators run across tile boundaries, so the TMA loads of tile j + 1 and its MMAs overlap the epilogue of
// tile j, and the per-CTA setup (barriers, TMEM allocation, resident weights) is paid once per SM instead of once per tile.
template <bool PRECISE, bool EVAL>
__device__ __forceinline__ void conv_tc_kernel_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const TmTaps* tmTp, const TcParams& p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_base = smem;
    // precise: [A hi ring][A lo ring][B hi][B lo]; the lo copies sit at constant offsets (p.a_lo_off, p.b_lo_off)
    const int dup = PRECISE ? 2 : 1;
    uint8_t* b_base = smem + (size_t)dup * p.stages * p.a_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + (size_t)dup * (p.b_resident ? (size_t)p.b_area_bytes : (size_t)p.stages * p.b_stage_bytes));
    uint64_t* full = bars;                              // TMA bytes of a ring slot have landed
    uint64_t* empty = bars + p.stages;
    uint64_t* tmem_full = bars + 2 * p.stages;          // [4]
    uint64_t* tmem_empty = tmem_full + 4;               // [4]
    uint64_t* b_full = tmem_empty + 4;
    uint64_t* split = b_full + 1;                       // [stages] precise: the slot's A data has been split into hi / lo
    uint64_t* ready = PRECISE ? split : full;           // what the MMA warp waits for
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(split + (PRECISE ? p.stages : 0));
    float* red = reinterpret_cast<float*>(tmem_slot + 2);          // [8 warps][2][nt] epilogue statistics

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * p.nt;
    if (threadIdx.x == 0) TC_TRACE(0);
    pdl_trigger();          // the prologue below (barriers, TMEM, statistics scratch) touches no global memory: it overlaps the previous kernel's tail

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); if (PRECISE) mbar_init(&split[s], 4); }
        for (int b = 0; b < 4; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], (p.epi_groups == 2 && p.colsplit) ? 8 : 4); }
        mbar_init(b_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < p.epi_groups * 8 * p.nt; i += p.epi_groups * 128) red[i] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();             // the previous kernel has completed: activations, weights, statistics slots may be touched from here on
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TC_TRACE(1);
    // reuse mode: one stage = (kz, kx, k-chunk) and covers 3 taps (ky = 0..2)
    const int groups = p.reuse ? p.taps / 3 : p.taps;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (warp-uniform, one elected lane issues)
        if (p.b_resident) {
            if (elect_one()) {
                mbar_expect_tx(b_full, (uint32_t)dup * p.b_area_bytes);
                for (int half = 0; half < dup; ++half)
                    for (int tap = 0; tap < p.taps; ++tap)
                        for (int kci = 0; kci < p.kchunks; ++kci)
                            tma_load_2d(b_base + (size_t)half * p.b_lo_off + (size_t)(tap * p.kchunks + kci) * p.b_box_bytes, &tmB, b_full,
                                        kci * p.kc, half * p.b_lo_rows + tap * p.n_total + n0);
            }
            __syncwarp();
        }
        if (p.thin2d) {
            // twin of issue_mmas_thin: one haloed box per (kz, kx) slot, only the ring position is carried
            int s = 0; uint32_t ph = 0;
            uint8_t* a_dst = a_base;
            const int ng = p.kxn ? (p.nd == 2 ? 1 : 3) : (p.nd == 2 ? 3 : 9);
            TileIter tj;
            for (tj.init(blockIdx.x, gridDim.x, p.tiles_w, p.tiles_h, p.tiles_d); tj.tile < p.tiles_total; tj.next(p.tiles_w, p.tiles_h, p.tiles_d)) {
                const int w0 = tj.tx * p.step_w - 1, h0 = tj.ty * (p.sup ? 2 * p.th : p.th) - 1, d0 = tj.tz * p.td - 1;
                for (int g = 0; g < ng; ++g) {
                    const int kx = p.kxn ? 0 : g % 3, kz = p.kxn ? g : g / 3;
                    mbar_wait(&empty[s], ph ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(&full[s], p.a_box_bytes);
                        if (p.nd == 2) tma_load_4d(a_dst, &tmA, &full[s], 0, w0 + kx, h0, tj.img);
                        else tma_load_5d(a_dst, &tmA, &full[s], 0, w0 + kx, h0, d0 + kz, tj.img);
                    }
                    __syncwarp();
                    a_dst += p.a_stage_bytes;
                    if (++s == p.stages) { s = 0; ph ^= 1u; a_dst = a_base; }
                }
            }
        } else {
        int s = 0; uint32_t ph = 0;
        TileIter ti;
        for (ti.init(blockIdx.x, gridDim.x, p.tiles_w, p.tiles_h, p.tiles_d); ti.tile < p.tiles_total; ti.next(p.tiles_w, p.tiles_h, p.tiles_d)) {
            const int img = ti.img;
            const int w0 = ti.tx * p.tw, h0 = ti.ty * p.th, d0 = ti.tz * p.td;
            for (int grp = 0; grp < groups; ++grp) {
                int kx, ky, kz;
                if (p.reuse) { kx = grp % 3; kz = grp / 3; ky = 0; }                 // box origin one row above the tile
                else if (p.ksz == 3) { kx = grp % 3; ky = (grp / 3) % 3; kz = grp / 9; }
                else { kx = ky = kz = p.pad; }                                        // 1x1: pad = 0
                for (int kc0 = 0; kc0 < p.kchunks; kc0 += p.cps) {
                    TC_WAIT(&empty[s], ph ^ 1);
                    if (elect_one()) {
                        const uint32_t nb = p.b_resident ? 0u : (p.reuse ? 3u : 1u);
                        mbar_expect_tx(&full[s], (uint32_t)p.cps * ((TC_DBG(2) ? 0u : p.a_box_bytes) + (uint32_t)dup * nb * p.b_box_bytes));
                        for (int c = 0; c < p.cps; ++c) {
                            const int kci = kc0 + c;
                            uint8_t* a_dst = a_base + (size_t)s * p.a_stage_bytes + (size_t)c * p.a_chunk_bytes;
                            if (TC_DBG(2)) {}
                            else if (p.mode == 2) tma_load_5d(a_dst, &tmA, &full[s], (grp & 1) * p.up_c + kci * p.kc, w0, grp >> 1, h0, img);
                            else if (p.mode == 3 && p.nd == 2) tma_load_4d(a_dst, &tmTp->m[grp], &full[s], kci * p.kc, w0, h0, img);
                            else if (p.mode == 3) tma_load_5d(a_dst, &tmTp->m[grp], &full[s], kci * p.kc, w0, h0, d0, img);
                            else if (p.nd == 2) tma_load_4d(a_dst, &tmA, &full[s], kci * p.kc, w0 + kx - p.pad, h0 + ky - p.pad, img);
                            else tma_load_5d(a_dst, &tmA, &full[s], kci * p.kc, w0 + kx - p.pad, h0 + ky - p.pad, d0 + kz - p.pad, img);
                            if (!p.b_resident) {
                                uint8_t* b_dst = b_base + (size_t)s * p.b_stage_bytes + (size_t)c * p.b_chunk_bytes;
                                for (int half = 0; half < dup; ++half) {
                                    uint8_t* bd = b_dst + (size_t)half * p.b_lo_off;
                                    const int r0 = half * p.b_lo_rows + n0;
                                    if (p.reuse) {
                                        for (int q = 0; q < 3; ++q)
                                            tma_load_2d(bd + (size_t)q * p.b_box_bytes, &tmB, &full[s], kci * p.kc, ((kz * 3 + q) * 3 + kx) * p.n_total + r0);
                                    } else {
                                        tma_load_2d(bd, &tmB, &full[s], kci * p.kc, grp * p.n_total + r0);
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
        }   // generic producer
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (p.thin2d && p.kxn) {
            if (p.nd == 2) { if (p.kc == 32) issue_mmas_thin<4, 1, PRECISE, true>(p, a_base, b_base, ready, empty, tmem_full, tmem_empty, b_full, tmem_base);
                             else issue_mmas_thin<2, 1, PRECISE, true>(p, a_base, b_base, ready, empty, tmem_full, tmem_empty, b_full, tmem_base); }
            else           { if (p.kc == 32) issue_mmas_thin<4, 3, PRECISE, true>(p, a_base, b_base, ready, empty, tmem_full, tmem_empty, b_full, tmem_base);
                             else issue_mmas_thin<2, 3, PRECISE, true>(p, a_base, b_base, ready, empty, tmem_full, tmem_empty, b_full, tmem_base); }
        } else
        if (p.thin2d) {
            if (p.nd == 2) { if (p.kc == 32) issue_mmas_thin<4, 3, PRECISE>(p, a_base, b_base, ready, empty, tmem_full, tmem_empty, b_full, tmem_base);
                             else issue_mmas_thin<2, 3, PRECISE>(p, a_base, b_base, ready, empty, tmem_full, tmem_empty, b_full, tmem_base); }
            else           { if (p.kc == 32) issue_mmas_thin<4, 9, PRECISE>(p, a_base, b_base, ready, empty, tmem_full, tmem_empty, b_full, tmem_base);
                             else issue_mmas_thin<2, 9, PRECISE>(p, a_base, b_base, ready, empty, tmem_full, tmem_empty, b_full, tmem_base); }
        } else
        if (p.kc == 32) { if (p.reuse) issue_mmas<4, 3, PRECISE>(p, a_base, b_base, ready, empty, tmem_full, tmem_empty, b_full, tmem_base, groups);
                          else issue_mmas<4, 1, PRECISE>(p, a_base, b_base, ready, empty, tmem_full, tmem_empty, b_full, tmem_base, groups); }
        else            { if (p.reuse) issue_mmas<2, 3, PRECISE>(p, a_base, b_base, ready, empty, tmem_full, tmem_empty, b_full, tmem_base, groups);
                          else issue_mmas<2, 1, PRECISE>(p, a_base, b_base, ready, empty, tmem_full, tmem_empty, b_full, tmem_base, groups); }
    } else if (PRECISE && warp >= 2 + 4 * p.epi_groups) {
        // ------------------------------------------------------------------ operand splitter (precise mode): 4 warps
        // Every ring slot, once its TMA bytes have landed (fp32, NOT rounded: the A tensor maps of this mode are FLOAT32):
        // hi = rna_tf32(x) in place, lo = x - hi (exact in fp32) into the lo ring at the same swizzled offset.  Elementwise, so
        // the swizzle is irrelevant.  Generic-proxy writes -> fence.proxy.async -> the MMA warp's tcgen05.mma may read them.
        const int t = (int)threadIdx.x - (64 + 128 * p.epi_groups);
        const int slot_uses = groups * (p.kchunks / p.cps);
        const uint32_t n16 = p.a_stage_bytes >> 4;
        int s = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
            for (int u = 0; u < slot_uses; ++u) {
                mbar_wait(&full[s], ph);
                float4* hi4 = reinterpret_cast<float4*>(a_base + (size_t)s * p.a_stage_bytes);
                float4* lo4 = reinterpret_cast<float4*>(a_base + p.a_lo_off + (size_t)s * p.a_stage_bytes);
                for (uint32_t i = (uint32_t)t; i < n16; i += 128u) {
                    const float4 x = hi4[i];
                    float4 h, l;
                    h.x = tf32_rn(x.x, 1); h.y = tf32_rn(x.y, 1); h.z = tf32_rn(x.z, 1); h.w = tf32_rn(x.w, 1);
                    l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
                    hi4[i] = h; lo4[i] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&split[s]);
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: 2 groups of 4 warps
        // The epilogue of one tile is a single latency-bound instruction stream per warp (measured: ~1.5 us per tile, the
        // bottleneck of the thin layers), so two warp groups work concurrently: with two accumulators group e owns
        // accumulator e (tiles j = e, e + 2, ...); with one accumulator (nt = 256) the groups split its columns.
        const int eg = (warp - 2) >> 2;
        const int lg = warp & 3;                                       // TMEM lane quarter this warp may read
        const int m = lg * 32 + lane;                                  // accumulator row == position inside the box
        const int dx = m % p.tw, dy = (m / p.tw) % p.th, dz = m / (p.tw * p.th);
        float* red_s = red + (size_t)((warp - 2) * 2 + 0) * p.nt;
        float* red_q = red + (size_t)((warp - 2) * 2 + 1) * p.nt;
        const int cb = p.n_total - p.ca;
        // two groups: with two accumulators group e owns accumulator e (tiles j = e, e + 2, ...), with one they split its columns;
        // one group: all tiles, accumulators alternating
        const bool two = p.epi_groups == 2;
        // column split whenever the halves are whole 16-column chunks: both groups work on EVERY tile, which halves the
        // epilogue latency of a tile (measured 0.9 us per 16-column chunk and warp: 7.2 us for nt = 128, on the critical path
        // of the deep layers that have one tile per CTA); nt = 16 alternates tiles between the groups instead
        const bool colsplit = two && p.colsplit;
        const bool alternate = two && !colsplit;
        const int c_begin = colsplit ? eg * (p.nt >> 1) : 0, c_end = colsplit ? c_begin + (p.nt >> 1) : p.nt;
        float* scratch = red + (size_t)p.epi_groups * 8 * p.nt + (size_t)(warp - 2) * (32 * 17);   // warp-private transpose scratch
        const bool bias_vec = (reinterpret_cast<uintptr_t>(p.bias) & 15u) == 0;     // parameters may sit at any 4-byte offset of an arena
        int jj0 = alternate ? eg : 0;                                  // index of the tile among this CTA's tiles
        // super tiles: the iteration units are pairs of tiles (j = 2k, 2k + 1 <-> accumulators 0, 1); with alternating groups, group e
        // takes sub-tile e of EVERY unit, otherwise both sub-tiles of a unit are processed one after the other
        const bool sup = p.sup != 0;
        const bool split_units = alternate && !sup;                    // no super tiles: the groups take alternate tiles
        const int sub_lo = sup ? (alternate ? eg : 0) : 0, sub_hi = sup ? (alternate ? eg + 1 : 2) : 1;
        TileIter ti;
        for (ti.init(blockIdx.x + (split_units ? eg * (int)gridDim.x : 0), (split_units ? 2 : 1) * (int)gridDim.x, p.tiles_w, p.tiles_h, p.tiles_d);
             ti.tile < p.tiles_total; ti.next(p.tiles_w, p.tiles_h, p.tiles_d), jj0 += (sup || alternate) ? 2 : 1)
        for (int sb = sub_lo; sb < sub_hi; ++sb) {
            const int jj = jj0 + ((sup && !alternate) ? sb : 0);
            const int buf = jj & (p.n_buf - 1);
            const uint32_t use = (uint32_t)(jj >> p.n_buf_lg);
            const int ow = p.kxn ? ti.tx * p.step_w + dx - 1 : ti.tx * p.tw + dx, oh = (sup ? 2 * ti.ty + sb : ti.ty) * p.th + dy, od = ti.tz * p.td + dz;
            const bool valid = (dz < p.td) && ow < p.W && oh < p.H && od < p.D && (!p.kxn || (dx >= 1 && dx <= p.step_w));
            const int64_t row = (((int64_t)ti.img * p.D + od) * p.H + oh) * p.W + ow;
            TC_WAIT(&tmem_full[buf], use & 1u);
            if (jj == 0 && threadIdx.x == 64) TC_TRACE(3);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * (p.kxn ? 3 * p.nt : p.nt));
            // one 16-column chunk: +bias, store, BatchNorm statistics
            auto process = [&](const uint32_t (&raw)[16], const int c0) {
                float v[16];
#pragma unroll
                for (int j4 = 0; j4 < 16; ++j4) v[j4] = TC_DBG(8) ? 0.f : __uint_as_float(raw[j4]);
                // channel of this chunk's first column: the taps of a transposed conv share the Cout channels
                const int ch0 = p.mode == 1 ? (n0 + c0) % p.up_c : c0;          // statistics slot (CTA-relative)
                const int bias0 = p.mode == 1 ? ch0 : n0 + c0;
                if (p.bias) {
                    if (p.n_real < 16) {                                // padded head: only n_real bias entries exist
#pragma unroll
                        for (int j4 = 0; j4 < 16; ++j4) if (j4 < p.n_real) v[j4] += __ldg(p.bias + j4);
                    } else if (bias_vec) {
                        const float4* b4 = reinterpret_cast<const float4*>(p.bias + bias0);
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 bb = __ldg(b4 + j4);
                            v[4 * j4] += bb.x; v[4 * j4 + 1] += bb.y; v[4 * j4 + 2] += bb.z; v[4 * j4 + 3] += bb.w;
                        }
                    } else {
#pragma unroll
                        for (int j4 = 0; j4 < 16; ++j4) v[j4] += __ldg(p.bias + bias0 + j4);
                    }
                }
                if (EVAL && p.epi_ss) {                                 // eval-mode BatchNorm + activation folded into the epilogue
#pragma unroll
                    for (int j4 = 0; j4 < 16; ++j4) {
                        const float sc = __ldg(p.epi_ss + bias0 + j4), sh = __ldg(p.epi_ss + p.epi_c + bias0 + j4);
                        const float t = fmaf(v[j4], sc, sh);
                        v[j4] = t > 0.f ? t : t * p.epi_slope;
                    }
                }
                if (valid && !TC_DBG(4) && !TC_DBG(128)) {
                    const int gc = n0 + c0;
                    float* dst;
                    if (p.mode == 1) {
                        // column = (tap, co); tap = ((kd * 2) + kh) * 2 + kw writes output pixel (2 d + kd, 2 h + kh, 2 w + kw)
                        const int tap = gc / p.up_c, co = gc - tap * p.up_c;
                        const int kw = tap & 1, kh = (tap >> 1) & 1, kd = tap >> 2;
                        const int64_t orow = p.nd == 2 ? ((int64_t)ti.img * (2 * p.H) + 2 * oh + kh) * (2 * p.W) + 2 * ow + kw
                                                       : (((int64_t)ti.img * (2 * p.D) + 2 * od + kd) * (2 * p.H) + 2 * oh + kh) * (2 * p.W) + 2 * ow + kw;
                        dst = p.out + orow * p.up_c + co;
                    } else {
                        dst = gc < p.ca ? p.out + row * p.ca + gc : p.out_b + row * cb + (gc - p.ca);
                    }
                    if (EVAL && p.epi_res) {                            // additive skip (vnet.py:202-215): same shape / offset as the output
                        const float4* r4 = reinterpret_cast<const float4*>(p.epi_res + (dst - p.out));
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 rr = __ldg(r4 + j4);
                            v[4 * j4] += rr.x; v[4 * j4 + 1] += rr.y; v[4 * j4 + 2] += rr.z; v[4 * j4 + 3] += rr.w;
                        }
                    }
                    if (p.n_real >= 16 && (reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
                        // two 256-bit stores (STG.256, sm_100) per 16-column chunk
                        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                                     "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
                        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 8), "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]),
                                     "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]) : "memory");
                    } else {
#pragma unroll
                        for (int j4 = 0; j4 < 16; j4 += 4)
                            if (j4 < p.n_real) *reinterpret_cast<float4*>(dst + j4) = make_float4(v[j4], v[j4 + 1], v[j4 + 2], v[j4 + 3]);
                    }
                }
                if (p.stats && !TC_DBG(4) && !TC_DBG(64)) warp_column_sums(v, valid, lane, scratch, red_s + ch0, red_q + ch0);
            };
            // ping-pong: the tcgen05.ld of chunk c + 1 is in flight while chunk c is processed; after the fence of the last
            // chunk every column of this warp's lanes is in registers and the accumulator goes back to the MMA warp
            auto release = [&]() {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[buf]);
            };
            uint32_t ra[16], rb[16];
            int c0 = c_begin;
            if (p.kxn) {
                // columns [kx][nt]: out[x] = D0[x - 1] + D1[x] + D2[x + 1]; this warp is one image row of the tile, lane = x
                if (p.kxn == 2) {
                    // one round trip: all three column groups of a 16-column chunk in flight together (48 registers)
                    uint32_t rc[16];
                    for (; c0 < c_end; c0 += 16) {
                        tc_ld16_issue(t_addr + (uint32_t)c0, ra);
                        tc_ld16_issue(t_addr + (uint32_t)(p.nt + c0), rb);
                        tc_ld16_issue(t_addr + (uint32_t)(2 * p.nt + c0), rc);
                        tc_ld16_fence(ra); tc_ld16_fence(rb); tc_ld16_fence(rc);
                        if (c0 + 16 >= c_end) release();
#pragma unroll
                        for (int j4 = 0; j4 < 16; ++j4)
                            rb[j4] = __float_as_uint(__uint_as_float(__shfl_up_sync(0xffffffffu, ra[j4], 1)) + __uint_as_float(rb[j4]) +
                                                     __uint_as_float(__shfl_down_sync(0xffffffffu, rc[j4], 1)));
                        process(rb, c0);
                    }
                    if (threadIdx.x == 64) { if (jj == 0) TC_TRACE(4); TC_TRACE(5); }
                    continue;
                }
                for (; c0 < c_end; c0 += 16) {
                    tc_ld16_issue(t_addr + (uint32_t)c0, ra);
                    tc_ld16_issue(t_addr + (uint32_t)(p.nt + c0), rb);
                    tc_ld16_fence(ra); tc_ld16_fence(rb);
#pragma unroll
                    for (int j4 = 0; j4 < 16; ++j4)
                        rb[j4] = __float_as_uint(__uint_as_float(__shfl_up_sync(0xffffffffu, ra[j4], 1)) + __uint_as_float(rb[j4]));
                    tc_ld16_issue(t_addr + (uint32_t)(2 * p.nt + c0), ra);
                    tc_ld16_fence(ra);
                    if (c0 + 16 >= c_end) release();
#pragma unroll
                    for (int j4 = 0; j4 < 16; ++j4)
                        rb[j4] = __float_as_uint(__uint_as_float(rb[j4]) + __uint_as_float(__shfl_down_sync(0xffffffffu, ra[j4], 1)));
                    process(rb, c0);
                }
                if (threadIdx.x == 64) { if (jj == 0) TC_TRACE(4); TC_TRACE(5); }
                continue;
            }
            tc_ld16_issue(t_addr + (uint32_t)c0, ra);
            while (true) {
                tc_ld16_fence(ra);
                if (c0 + 16 < c_end) tc_ld16_issue(t_addr + (uint32_t)(c0 + 16), rb); else release();
                process(ra, c0);
                c0 += 16;
                if (c0 >= c_end) break;
                tc_ld16_fence(rb);
                if (c0 + 16 < c_end) tc_ld16_issue(t_addr + (uint32_t)(c0 + 16), ra); else release();
                process(rb, c0);
                c0 += 16;
                if (c0 >= c_end) break;
            }
            if (threadIdx.x == 64) { if (jj == 0) TC_TRACE(4); TC_TRACE(5); }
        }
        if (p.stats) {
            asm volatile("bar.sync 1, %0;" ::"r"(p.epi_groups * 128) : "memory");   // the epilogue warps only
            const int e = threadIdx.x - 64;
            const int n_ch = p.mode == 1 ? p.up_c : (p.n_real < p.n_total ? p.n_real : p.n_total), ch_base = p.mode == 1 ? 0 : n0;
            const int n_mine = p.mode == 1 ? p.up_c : (p.n_real < p.nt ? p.n_real : p.nt);
            double* slot = p.stats + (size_t)(blockIdx.x % CHAP_STAT_SLOTS) * 2 * n_ch;
            for (int c = e; c < n_mine; c += p.epi_groups * 128) {
                float a = 0.f, b = 0.f;
                for (int w = 0; w < p.epi_groups * 4; ++w) { a += red[(size_t)(w * 2) * p.nt + c]; b += red[(size_t)(w * 2 + 1) * p.nt + c]; }
                atomicAdd(slot + ch_base + c, (double)a);
                atomicAdd(slot + n_ch + ch_base + c, (double)b);
            }
            if (p.bn.mi) {
                // BatchNorm finalize by the last CTA of the grid: ticket counter after the statistics slots
                __threadfence();                                                  // this thread's atomics are visible device-wide
                asm volatile("bar.sync 1, %0;" ::"r"(p.epi_groups * 128) : "memory");
                if (e == 0) {
                    const unsigned ticket = atomicAdd(p.bn.counter, 1u);
                    tmem_slot[1] = ticket == gridDim.x * gridDim.y - 1 ? 1u : 0u;
                }
                asm volatile("bar.sync 1, %0;" ::"r"(p.epi_groups * 128) : "memory");
                if (tmem_slot[1]) {
                    __threadfence();
                    for (int c = e; c < n_ch; c += p.epi_groups * 128) {
                        double s1 = 0.0, s2 = 0.0;
                        for (int sl = 0; sl < CHAP_STAT_SLOTS; ++sl) {
                            s1 += __ldcg(p.stats + (size_t)sl * 2 * n_ch + c);
                            s2 += __ldcg(p.stats + (size_t)sl * 2 * n_ch + n_ch + c);
                        }
                        const double mean = s1 / p.bn.count;
                        double var = s2 / p.bn.count - mean * mean;
                        if (var < 0.0) var = 0.0;
                        const float invstd = (float)(1.0 / sqrt(var + (double)p.bn.eps));
                        const float sc = p.bn.gamma[c] * invstd;
                        p.bn.mi[c] = (float)mean; p.bn.mi[n_ch + c] = invstd;
                        p.bn.ss[c] = sc; p.bn.ss[n_ch + c] = p.bn.beta[c] - (float)mean * sc;
                        if (p.bn.rmean) {
                            const double unbiased = p.bn.count > 1.0 ? var * p.bn.count / (p.bn.count - 1.0) : var;
                            p.bn.rmean[c] = (1.f - p.bn.momentum) * p.bn.rmean[c] + p.bn.momentum * (float)mean;
                            p.bn.rvar[c] = (1.f - p.bn.momentum) * p.bn.rvar[c] + p.bn.momentum * (float)unbiased;
                        }
                        if (p.bn.rezero) {                       // persistent statistics buffer: hand it back zeroed (no zero-fill launch per conv)
                            for (int sl = 0; sl < CHAP_STAT_SLOTS; ++sl) {
                                p.stats[(size_t)sl * 2 * n_ch + c] = 0.0;
                                p.stats[(size_t)sl * 2 * n_ch + n_ch + c] = 0.0;
                            }
                        }
                    }
                    if (e == 0 && p.bn.rezero) *p.bn.counter = 0u;
                    if (e == 0 && p.bn.nbt) *p.bn.nbt += 1;
                }
            }
        }
        if (threadIdx.x == 64) TC_TRACE(6);
        tc_fence_before();
    }
    __syncthreads();
    if (threadIdx.x == 0) TC_TRACE(7);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// Two entry points: the per-tap tensor maps (1 KB of kernel parameters) are only passed for the k2 s2 gathers.
// Entry points.  The per-tap tensor maps (1 KB of kernel parameters) are only passed for the k2 s2 gathers (`_taps`); PRECISE adds the
// operand-splitting warps (split-operand 3xTF32, chap_set_conv_precision); EVAL compiles the inference epilogue in (eval-mode
// BatchNorm + activation + skip add, chap_conv_bn_act_fwd) -- kept out of the training kernels, whose 96-register budget is tight.
template <bool PRECISE, bool EVAL>
__global__ void __launch_bounds__(PRECISE ? kTcThreadsPrecise : kTcThreadsMax, 2)
conv_tc_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    conv_tc_kernel_body<PRECISE, EVAL>(tmA, tmB, nullptr, p);
}
template <bool PRECISE, bool EVAL>
__global__ void __launch_bounds__(PRECISE ? kTcThreadsPrecise : kTcThreadsMax, 2)
conv_tc_k_taps(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ TmTaps tmT,
               const TcParams p) {
    conv_tc_kernel_body<PRECISE, EVAL>(tmA, tmB, &tmT, p);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

int make_tensor_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int kc, bool mn_major_tf32, bool plain_f32) {
    // cuTensorMapEncodeTiled is a DRIVER call: it needs the primary context current in this thread.  autograd's
    // backward thread may not have touched the runtime yet -> bind it once per thread with a no-op runtime call.
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
    EncodeTiledFn fn = encode_fn();
    CHAP_REQUIRE(fn != nullptr, CHAP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[5]; cuuint64_t gstr[4]; cuuint32_t bdim[5]; cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    // TFLOAT32: the TMA unit rounds the fp32 data to nearest TF32 on the way into shared memory (measured, common.cuh);
    // FLOAT32 (split-operand mode): bits arrive untouched and the splitter warps make the TF32 hi / lo halves
    CUresult r = fn(map, plain_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    mn_major_tf32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B),
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CHAP_REQUIRE(r == CUDA_SUCCESS, CHAP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
    return CHAP_OK;
}

static int tc_channels(const Geom& g, bool dgrad, int& K, int& N) {
    K = dgrad ? g.cout : g.cin;
    N = dgrad ? g.cin : g.cout;
    return 0;
}

bool tc_supports(const Geom& g, bool dgrad) {
    if (g.kind == CHAP_CONV_DOWN2) {
        // forward: 4- / 8-tap gather (one tensor map per tap); data gradient: one GEMM with N = taps * Cin and a scatter epilogue
        if (getenv("CHAP_NO_UP2_TC")) return false;
        const bool chan_ok = (g.cin == 16 || g.cin % 32 == 0) && (g.cout == 16 || g.cout % 32 == 0) && g.cin <= 1024 && g.cout <= 1024;
        if (!dgrad) return chan_ok && (g.cout <= 256 || g.cout % 256 == 0);
        return chan_ok && ((g.taps * g.cin) <= 256 || (g.taps * g.cin) % 256 == 0);
    }
    if (g.kind == CHAP_CONV_UP2) {
        // forward: one GEMM [pixels, Cin] x [Cin, taps * Cout] (2D and 3D); data gradient: gather of dy, 2D through its
        // [2C, W, 2, H, N] view, 3D with one tensor map per tap
        if (getenv("CHAP_NO_UP2_TC")) return false;
        const bool chan_ok = (g.cin == 16 || g.cin % 32 == 0) && g.cout % 16 == 0 && g.cout >= 16 && g.cout <= 128 && g.cin <= 1024;
        if (!dgrad) return chan_ok && ((g.taps * g.cout) <= 256 || (g.taps * g.cout) % 256 == 0);
        return chan_ok && (g.cout == 16 || g.cout % 32 == 0) && (g.cin <= 256 || g.cin % 256 == 0);
    }
    if (g.kind != CHAP_CONV_K3 && g.kind != CHAP_CONV_K1) return false;
    int K, N;
    tc_channels(g, dgrad, K, N);
    // thin heads (Cout = 4 / 8), forward only: N zero-padded to 16 in the packed weight, 4 / 8 columns stored (16 -> 4 @ 12x256^2:
    // 45 us vs 93 us on the CUDA cores).  The data gradient (K = 4 padded to 16: 16-byte TMA rows) measured 84 us vs 77 us on
    // the CUDA cores and stays there.
    static const bool no_heads = getenv("CHAP_NO_HEAD_TC") != nullptr;
    if (g.kind == CHAP_CONV_K3 && !dgrad && !no_heads && (g.cout == 4 || g.cout == 8) && (g.cin == 16 || (g.cin % 32 == 0 && g.cin <= 256))) return true;
    if (!(K == 16 || K % 32 == 0)) return false;
    if (N < 16 || N % 16 != 0) return false;
    if (N > 256 && N % 256 != 0) return false;
    return true;
}

// choose the spatial box (tw, th, td), tw*th*td <= 128, that covers the volume with the fewest tiles
static void choose_box(int W, int H, int D, int& tw, int& th, int& td) {
    long best = -1;
    for (int w = 1; w <= W && w <= 128; ++w)
        for (int h = 1; h <= H && w * h <= 128; ++h) {
            int d = 128 / (w * h);
            if (d > D) d = D;
            if (d < 1) continue;
            long tiles = (long)((W + w - 1) / w) * ((H + h - 1) / h) * ((D + d - 1) / d);
            long score = tiles * 1024 - w;                 // fewest tiles, then the widest contiguous run
            if (best < 0 || score < best) { best = score; tw = w; th = h; td = d; }
        }
}

int tc_conv(const Geom& g, bool dgrad, const float* in, const float* wp, const float* bias, float* out,
            double* ch_sums, cudaStream_t st, float* out_b, int ca, const BnFold* bn, const EvalEpilogue* epi) {
    if (!tc_supports(g, dgrad)) return 0;
    if (epi && (dgrad || out_b || ch_sums || g.cout < 16 || g.cout % 16 != 0 || !aligned16(epi->residual))) return 0;
    CHAP_REQUIRE(aligned16(in) && aligned16(wp) && aligned16(out), CHAP_ERR_ALIGNMENT, "tc_conv: buffers must be 16-byte aligned");
    int K, N;
    tc_channels(g, dgrad, K, N);
    TcParams p{};
    p.nd = g.nd; p.ksz = g.kind == CHAP_CONV_K3 ? 3 : 1; p.pad = g.kind == CHAP_CONV_K3 ? 1 : 0; p.taps = g.taps;
    p.W = g.iW; p.H = g.iH; p.D = g.iD;
    if (g.kind == CHAP_CONV_DOWN2) { p.W = g.oW; p.H = g.oH; p.D = g.oD; }      // the k2 s2 modes tile the LOW-resolution grid
    const int k_real = K, n_real = N;                    // thin heads: the MMA runs on operands zero-padded to 16
    K = tc_pad16(K); N = tc_pad16(N);
    int w_taps = g.taps;                                  // taps in the packed weight operand [tap][N][K]
    if (g.kind == CHAP_CONV_UP2 && !dgrad) { p.mode = 1; p.up_c = g.cout; N = g.taps * g.cout; p.taps = 1; w_taps = 1; }
    if (g.kind == CHAP_CONV_UP2 && dgrad) { p.mode = g.nd == 2 ? 2 : 3; p.up_c = g.cout; p.ksz = 2; }
    if (g.kind == CHAP_CONV_DOWN2 && !dgrad) { p.mode = 3; p.up_c = g.cin; p.ksz = 2; }
    if (g.kind == CHAP_CONV_DOWN2 && dgrad) { p.mode = 1; p.up_c = g.cin; N = g.taps * g.cin; p.taps = 1; w_taps = 1; }
    CHAP_REQUIRE(!(p.mode && out_b), CHAP_ERR_BAD_ARG, "tc_conv: split output is not available for the k2 s2 convolutions");
    p.n_real = p.mode == 1 ? N : n_real;
    CHAP_REQUIRE(p.n_real == N || !out_b, CHAP_ERR_BAD_ARG, "tc_conv: split output is not available for padded heads");
    choose_box(p.W, p.H, p.D, p.tw, p.th, p.td);
    // Row-reuse mode for the activation-bound layers (few output channels): in-plane 128-pixel tile (tw x th, tw % 8 == 0)
    // and ONE TMA box with an h-halo (th + 2 rows) per (kz, kx); the three ky taps read it at row offsets ky * tw.
    // L2 -> smem traffic for A drops from 9 (27) boxes of th rows to 3 (9) boxes of th + 2 rows.
    p.reuse = 0;
    if (g.kind == CHAP_CONV_K3 && N <= 64 && getenv("CHAP_NO_ROW_REUSE") == nullptr) {
        int best_tw = 0; long best_tiles = -1;
        for (int tw : {8, 16, 32}) {
            const int th = 128 / tw;
            if (p.W < tw || p.H < th + 2) continue;                 // the TMA box must fit inside the tensor extents
            long tiles = (long)((p.W + tw - 1) / tw) * ((p.H + th - 1) / th);
            if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && th > 128 / best_tw)) { best_tiles = tiles; best_tw = tw; }
        }
        if (best_tw) { p.reuse = 1; p.tw = best_tw; p.th = 128 / best_tw; p.td = 1; }
    }
    // kx-in-N (see TcParams::kxn): thin layers (one k chunk, N <= 32), plain output, training and inference kernels.  MEASURED on what bounds those
    // layers: the tensor pipe spends ~36 cycles per M128 x K8 TF32 MMA whatever N is (the ncu source view shows the issuing warp
    // stalled on the MMA queue 54 % of its time), so 18 (2D) / 54 (3D) MMAs of N = 16 per tile become 6 / 18 of N = 48.
    // Result (tools/conv_bench.py): 3D 16 -> 16 @ 2x112x112x80 forward 149 -> 105 us, data gradient 148 -> 88 us.  In 2D the 18 MMAs
    // were not the only bound: with 12 % more tiles (30 of 32 columns useful) and three TMEM reads + 32 shuffles per tile in the
    // epilogue the layer gets SLOWER (16 -> 16 @ 12x256^2: 40.3 -> 43.5 us), so the 16-channel 2D layers keep the three-box tile (CHAP_TC_KXN2D).
    // resident weights up to 112 KB for these layers (3D 32 -> 32: 27 taps x 4 KB = 108 KB): one CTA per SM then, which the shorter
    // MMA stream of the kx-in-N tile more than pays for (CHAP_TC_KXN_RES overrides the cap in KB)
    // 2D: 0 off, 1 all thin layers, 32 (default): the layers with a 32-channel operand -- with the one-round-trip epilogue those gain 1-2 us
    // per launch (32 -> 32 @ 12x128^2: 33.3 -> 31.1 / 27.3 -> 26.1 us), the 16 -> 16 layers stay on the three-box tile (41.9 vs 43.0 us);
    // replayed 2D iteration, same box: 12.553 (off) / 12.489 (32) / 12.48-12.53 (all) ms
    static const int kxn2d = getenv("CHAP_TC_KXN2D") ? atoi(getenv("CHAP_TC_KXN2D")) : 32;
    static const size_t kxn_res_cap = (getenv("CHAP_TC_KXN_RES") ? (size_t)atoi(getenv("CHAP_TC_KXN_RES")) : 112u) * 1024u;
    p.step_w = p.tw;
    p.kxn = (p.reuse && p.mode == 0 && K <= 32 && N <= 32 && !out_b && p.W >= 32 && p.H >= 6 && TC_DBG_HOST_OFF &&
             (g.nd == 3 || kxn2d == 1 || (kxn2d == 32 && (K == 32 || N == 32))) &&
             // weights resident (thin2d loops).  Inference epilogue: only while two CTAs per SM still fit (32 -> 32 @ 4x56x56x40 with
             // one CTA per SM: 116 us against 101 us for the generic two-CTA path; 16 -> 16 @ 4x112x112x80: 213 against 284 us)
             (size_t)p.taps * N * K * 4 <= (epi ? 40u * 1024u : kxn_res_cap) && getenv("CHAP_NO_RESIDENT_B") == nullptr &&
             g_precise_max_c.load(std::memory_order_relaxed) == 0 && getenv("CHAP_TC_NO_KXN") == nullptr && getenv("CHAP_TC_NO_THIN2D") == nullptr &&
             getenv("CHAP_TC_SUPER") == nullptr) ? 1 : 0;
    if (p.kxn) { p.tw = 32; p.th = 4; p.td = 1; p.step_w = 30; }
    // epilogue form: 2 = all three column groups of a chunk in ONE TMEM round trip (48 registers; default: 1-4 % faster than the two
    // round trips of form 1 on every layer measured, CHAP_TC_KXN_2RT=1 restores those)
    if (p.kxn && getenv("CHAP_TC_KXN_2RT") == nullptr) p.kxn = 2;
    p.tiles_w = (p.W + p.step_w - 1) / p.step_w; p.tiles_h = (p.H + p.th - 1) / p.th; p.tiles_d = (p.D + p.td - 1) / p.td;
    p.kc = K == 16 ? 16 : 32; p.kchunks = K / p.kc;
    p.n_total = N; p.nt = N > 256 ? 256 : N;
    p.tiles_total = g.n * p.tiles_d * p.tiles_h * p.tiles_w;
    // Small-M layers (deep levels: 3D 128 -> 128 @ 10x14x14 has 31 M tiles, 2D 256 -> 256 @ 16^2 has 24) cannot fill 148 SMs with
    // one CTA per M tile: split N across CTAs (blockIdx.y) until the grid covers half the machine (measured: 96 tiles are better
    // left alone -- 128 -> 128 @ 12x32^2: 29 us whole, 33 us split; 24 tiles gain -- 256 -> 256 @ 12x16^2: 45 -> 38 us).  Every CTA re-reads its A boxes
    // (L2 hits: these tensors are a few MB) and owns nt output channels end to end, so bias, stores and the BatchNorm
    // statistics need no cross-CTA reduction (unlike split-K).  CHAP_TC_NT forces nt for experiments.
    if (p.mode != 1) {
        static const int nt_force = getenv("CHAP_TC_NT") ? atoi(getenv("CHAP_TC_NT")) : 0;
        static const int fill = getenv("CHAP_TC_FILL") ? atoi(getenv("CHAP_TC_FILL")) : kNumSMs / 2;
        while (p.nt >= 64 && p.nt % 32 == 0 && (long)p.tiles_total * (N / p.nt) < fill && !nt_force) p.nt /= 2;
        if (nt_force >= 16 && nt_force % 16 == 0 && N % nt_force == 0 && nt_force <= p.nt) p.nt = nt_force;
    }
    // Accumulators: two while 2 CTAs/SM still fit in 512 columns; FOUR for nt <= 64 -- the ncu source view of the thin layers shows the
    // MMA warp waiting 14 % of its time for an accumulator (the two epilogue groups are ~75 % busy, so their jitter reaches it)
    static const int nbuf_force = getenv("CHAP_TC_NBUF") ? atoi(getenv("CHAP_TC_NBUF")) : 0;
    p.n_buf = p.nt <= 64 ? 4 : (p.nt <= 128 ? 2 : 1);
    if (p.kxn && p.nt > 16) p.n_buf = 2;                                  // three column groups per accumulator: 2 x 96 columns for nt = 32
    if ((nbuf_force == 1 || nbuf_force == 2 || nbuf_force == 4) && nbuf_force * p.nt <= 256) p.n_buf = nbuf_force;
    p.n_buf_lg = p.n_buf == 4 ? 2 : (p.n_buf == 2 ? 1 : 0);
    p.tmem_cols = 32; while (p.tmem_cols < p.n_buf * (p.kxn ? 3 : 1) * p.nt) p.tmem_cols *= 2;
    p.b_box_bytes = (uint32_t)p.nt * p.kc * 4u;
    p.b_area_bytes = (uint32_t)p.taps * p.kchunks * p.b_box_bytes;
    // split-operand 3xTF32 (chap_set_conv_precision): every A stage and every weight box exists twice (hi / lo halves)
    const int pmc = g_precise_max_c.load(std::memory_order_relaxed);
    p.precise = (pmc > 0 && (g.cin > g.cout ? g.cin : g.cout) <= pmc) ? 1 : 0;
    const uint32_t dup = p.precise ? 2u : 1u;
    p.b_resident = (dup * p.b_area_bytes <= 40u * 1024u || p.kxn) && getenv("CHAP_NO_RESIDENT_B") == nullptr;
    // Super tiles for the thin row-reuse layers (the conditions of the lean thin2d loops, known at this point).  MEASURED (round 2,
    // tools/conv_bench.py, same box): correct, but NOT faster -- 16 -> 16 @ 12x256^2 forward 44.0 us with, 42.1 us without; 3D 16 -> 16 @
    // 2x112x112x80 173 vs 150 us (4 ring stages of 20 KB instead of 6 of 13 KB).  Halving the TMA issues and ring round trips per tile
    // does not move these kernels, nor do polling mbarrier waits (-DCHAP_MBAR_POLL: 44.9 vs 43.2 us) -- so neither the TMA issue rate
    // nor barrier wake-up latency is what bounds the thin layers.  Opt-in (CHAP_TC_SUPER=1) experiment, off by default.
    p.sup = (p.mode == 0 && p.reuse && p.kchunks == 1 && p.b_resident && p.n_buf >= 2 && p.H >= 2 * p.th + 2 && TC_DBG_HOST_OFF &&
             getenv("CHAP_TC_NO_THIN2D") == nullptr && getenv("CHAP_TC_SUPER") != nullptr) ? 1 : 0;
    if (p.sup) {
        p.n_buf = 2; p.n_buf_lg = 1;                                     // a super tile IS the pair of accumulators
        p.tiles_h = (p.tiles_h + 1) / 2;                                 // rows of super tiles (an odd last row gets a phantom second tile: all rows masked)
        p.tiles_total = g.n * p.tiles_d * p.tiles_h * p.tiles_w;
    }
    if (p.reuse) {
        p.a_box_bytes = (uint32_t)(p.tw * ((p.sup ? 2 * p.th : p.th) + 2)) * p.kc * 4u;
        p.a_chunk_bytes = (p.a_box_bytes + 1023u) & ~1023u;              // 128 + 2 tw rows: every ky view of 128 rows stays inside
        p.b_chunk_bytes = 3u * p.b_box_bytes;
    } else {
        p.a_chunk_bytes = 128u * p.kc * 4u;                              // always room for 128 rows
        p.a_box_bytes = (uint32_t)(p.tw * p.th * p.td) * p.kc * 4u;
        p.b_chunk_bytes = (p.b_box_bytes + 1023u) & ~1023u;
    }
    // CTAs per SM.  Measured (16 -> 16 @ 256^2, b12): 2 CTAs x 6 stages 45 us, 3 CTAs 50 us, 4 CTAs x 3 stages 61 us -- a tile costs
    // ~1 us of scalar work in the MMA warp AND ~2 us of load -> MMA -> commit round trip per ring slot, so more CTAs with
    // shallower rings gain nothing.  CHAP_TC_CTAS / CHAP_TC_EPI override for experiments.
    int ctas_per_sm = 2;
    if (getenv("CHAP_TC_CTAS")) ctas_per_sm = atoi(getenv("CHAP_TC_CTAS"));
    if (ctas_per_sm > 2 && p.nt > 64) ctas_per_sm = 2;
    const size_t fixed = p.b_resident ? dup * p.b_area_bytes : 0u;
    const size_t chunk_bytes = dup * ((size_t)p.a_chunk_bytes + (p.b_resident ? 0u : p.b_chunk_bytes));
    int grid_x = 0;
    size_t extras = 0, budget = 0;
    for (;; --ctas_per_sm) {
        p.epi_groups = ctas_per_sm > 2 ? 1 : 2;
        if (getenv("CHAP_TC_EPI")) p.epi_groups = atoi(getenv("CHAP_TC_EPI")) == 1 ? 1 : 2;
        if (ctas_per_sm > 3 && p.epi_groups == 2) ctas_per_sm = 3;      // register file: 3 x 320 threads x 64 registers
        grid_x = p.tiles_total < kNumSMs * ctas_per_sm ? p.tiles_total : kNumSMs * ctas_per_sm;
        if (getenv("CHAP_NO_PERSIST")) grid_x = p.tiles_total;
        if (getenv("CHAP_TC_GRID")) grid_x = atoi(getenv("CHAP_TC_GRID")) < p.tiles_total ? atoi(getenv("CHAP_TC_GRID")) : p.tiles_total;
        extras = 1024 + 256 + (size_t)p.epi_groups * (8 * p.nt + 4 * 32 * 17) * sizeof(float);   // alignment slack, barriers, epilogue scratch
        // with <= 148 CTAs in the grid a CTA may use a whole SM's shared memory
        const size_t per_cta = (long)grid_x * (N / p.nt) <= kNumSMs ? 224 * 1024 : (size_t)(226 * 1024) / ctas_per_sm;
        budget = per_cta > extras + fixed ? per_cta - extras - fixed : 0;
        // at least three pipeline stages, else fewer CTAs per SM; one CTA per SM only when two cannot hold two stages each
        // (split-operand mode of the wide layers: every stage exists twice)
        if (budget >= 3 * chunk_bytes || (ctas_per_sm == 2 && budget >= 2 * chunk_bytes) || ctas_per_sm <= 1) break;
    }
    CHAP_REQUIRE(budget >= chunk_bytes, CHAP_ERR_BAD_ARG, "tc_conv: tile does not fit shared memory");
    p.colsplit = (p.nt >= 32 || p.n_buf == 1) && getenv("CHAP_TC_ALTERNATE") == nullptr ? 1 : 0;
    if (p.n_buf == 1) p.colsplit = 1;
    // k chunks per pipeline stage.  The MMA warp pays ~0.4 us of scalar work per stage (measured), which dominates the deep
    // layers (K = 9 x 256 in 32-channel chunks = 72 stages of 4 MMAs per tile): fuse 2 or 4 chunks into one stage when at
    // least three such stages fit.
    p.cps = 1;
    const int cps_max = getenv("CHAP_TC_CPS") ? atoi(getenv("CHAP_TC_CPS")) : 4;
    for (int c : {4, 2}) {
        if (c <= cps_max && p.kchunks % c == 0 && 3 * c * chunk_bytes <= budget) { p.cps = c; break; }
    }
    p.a_stage_bytes = p.cps * p.a_chunk_bytes; p.b_stage_bytes = p.cps * p.b_chunk_bytes;
    const size_t stage = p.cps * chunk_bytes;
    int stages = (int)(budget / stage);
    const int iters = (p.reuse ? p.taps / 3 : p.taps) * (p.kchunks / p.cps);
    const long stage_uses = (long)iters * ((p.tiles_total + grid_x - 1) / grid_x);     // ring slots one CTA ever fills
    const int stage_cap = getenv("CHAP_TC_STAGES") ? atoi(getenv("CHAP_TC_STAGES")) : 6;
    if (stages > stage_cap) stages = stage_cap;
    if (stages > (p.precise ? 7 : 10)) stages = p.precise ? 7 : 10;          // barrier block: 3 (2) barriers per stage + 9 in 256 bytes
    if (stages > stage_uses) stages = (int)stage_uses;
    if (stages < 2) stages = stage_uses < 2 ? 1 : 2;
    // lean issue / producer loops for the thin row-reuse layers (see issue_mmas_thin)
    p.thin2d = (p.mode == 0 && p.reuse && p.kchunks == 1 && p.cps == 1 && p.b_resident && p.n_buf >= 2 && stages >= 2 &&
                TC_DBG_HOST_OFF && getenv("CHAP_TC_NO_THIN2D") == nullptr) ? 1 : 0;
    CHAP_REQUIRE(!p.sup || p.thin2d, CHAP_ERR_BAD_ARG, "tc_conv: super tiles need the thin-layer loops (stages %d, cps %d)", stages, p.cps);
    CHAP_REQUIRE(!p.kxn || p.thin2d, CHAP_ERR_BAD_ARG, "tc_conv: the kx-in-N tile needs the thin-layer loops (stages %d, cps %d, resident %d)", stages, p.cps, p.b_resident);
    p.stages = stages;
    p.a_lo_off = (uint32_t)stages * p.a_stage_bytes;
    p.b_lo_off = p.b_resident ? p.b_area_bytes : (uint32_t)stages * p.b_stage_bytes;
    p.debug = getenv("CHAP_TC_DEBUG") ? atoi(getenv("CHAP_TC_DEBUG")) : 0;
    p.out = out; p.out_b = out_b; p.ca = out_b ? ca : p.n_real; p.bias = bias; p.stats = ch_sums;
    if (epi) { p.epi_ss = epi->scale_shift; p.epi_res = epi->residual; p.epi_slope = epi->slope; p.epi_c = g.cout; }
    CHAP_REQUIRE(!out_b || (ca > 0 && ca < N && ca % 16 == 0 && (N - ca) % 16 == 0 && aligned16(out_b)), CHAP_ERR_BAD_ARG,
                 "tc_conv: split output needs 16-channel aligned parts (ca %d of %d)", ca, N);
    const size_t smem = (size_t)stages * stage + fixed + extras;
    static_assert(2 * 10 + 9 <= 256 / 8 - 2 && 3 * 7 + 9 <= 256 / 8 - 2, "barrier block fits the 256-byte slot");

    // tensor maps: activations [C, W, H, (D,) N] (channels-last), weights [K, taps * N]
    CUtensorMap tmA, tmB;
    static thread_local TmTaps tmT;                     // only filled (and read by the kernel) in mode 3
    {
        uint64_t dims[5], str[4]; uint32_t box[5];
        const uint64_t C = (uint64_t)k_real;                 // a box wider than the tensor zero-fills (padded heads)
        if (p.mode == 3) {
            // high-resolution tensor [N, 2D, 2H, 2W, C]: tap (kd, kh, kw) of low-resolution position (d, h, w) is element
            // (2d + kd, 2h + kh, 2w + kw) -> per tap a map with the base moved by the tap and every spatial stride doubled
            const uint64_t bw = 2 * (uint64_t)p.W, bh = 2 * (uint64_t)p.H;
            for (int t = 0; t < g.taps; ++t) {
                const int kw = t & 1, kh = (t >> 1) & 1, kd = t >> 2;
                const float* base = in + (((uint64_t)kd * bh + kh) * bw + kw) * C;
                if (g.nd == 2) {
                    dims[0] = C; dims[1] = p.W; dims[2] = p.H; dims[3] = g.n;
                    str[0] = 2 * C * 4; str[1] = 2 * bw * C * 4; str[2] = bh * bw * C * 4;
                    box[0] = p.kc; box[1] = p.tw; box[2] = p.th; box[3] = 1;
                    CHAP_TRY(make_tensor_map(&tmT.m[t], base, 4, dims, str, box, p.kc, false, p.precise));
                } else {
                    dims[0] = C; dims[1] = p.W; dims[2] = p.H; dims[3] = p.D; dims[4] = g.n;
                    str[0] = 2 * C * 4; str[1] = 2 * bw * C * 4; str[2] = 2 * bh * bw * C * 4; str[3] = 2 * (uint64_t)p.D * bh * bw * C * 4;
                    box[0] = p.kc; box[1] = p.tw; box[2] = p.th; box[3] = p.td; box[4] = 1;
                    CHAP_TRY(make_tensor_map(&tmT.m[t], base, 5, dims, str, box, p.kc, false, p.precise));
                }
            }
            tmA = tmT.m[0];
        } else if (p.mode == 2) {
            // dy [N, 2H, 2W, C] seen as [2C (kw, c), W, 2 (kh), H, N]: tap (kh, kw) of input pixel (h, w) is one box row
            dims[0] = 2 * C; dims[1] = p.W; dims[2] = 2; dims[3] = p.H; dims[4] = g.n;
            str[0] = 2 * C * 4; str[1] = str[0] * p.W; str[2] = 2 * str[1]; str[3] = str[2] * p.H;
            box[0] = p.kc; box[1] = p.tw; box[2] = 1; box[3] = p.th; box[4] = 1;
            CHAP_TRY(make_tensor_map(&tmA, in, 5, dims, str, box, p.kc, false, p.precise));
        } else if (g.nd == 2) {
            dims[0] = C; dims[1] = p.W; dims[2] = p.H; dims[3] = g.n;
            str[0] = C * 4; str[1] = str[0] * p.W; str[2] = str[1] * p.H;
            box[0] = p.kc; box[1] = p.tw; box[2] = p.reuse ? (p.sup ? 2 * p.th : p.th) + 2 : p.th; box[3] = 1;
            CHAP_TRY(make_tensor_map(&tmA, in, 4, dims, str, box, p.kc, false, p.precise));
        } else {
            dims[0] = C; dims[1] = p.W; dims[2] = p.H; dims[3] = p.D; dims[4] = g.n;
            str[0] = C * 4; str[1] = str[0] * p.W; str[2] = str[1] * p.H; str[3] = str[2] * p.D;
            box[0] = p.kc; box[1] = p.tw; box[2] = p.reuse ? (p.sup ? 2 * p.th : p.th) + 2 : p.th; box[3] = p.td; box[4] = 1;
            CHAP_TRY(make_tensor_map(&tmA, in, 5, dims, str, box, p.kc, false, p.precise));
        }
        // packed weights: [hi half | lo half], each [w_taps * N rows][K]; the plain TF32 kernel only ever addresses the hi rows
        p.b_lo_rows = w_taps * N;
        uint64_t wd[2] = {(uint64_t)K, 2 * (uint64_t)w_taps * N};
        uint64_t ws[1] = {(uint64_t)K * 4};
        uint32_t wb[2] = {(uint32_t)p.kc, (uint32_t)p.nt};
        CHAP_TRY(make_tensor_map(&tmB, wp, 2, wd, ws, wb, p.kc));
    }
    static std::once_flag attr_once;
    std::call_once(attr_once, [] {
        const int big = 227 * 1024;
        cudaFuncSetAttribute(conv_tc_k<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(conv_tc_k<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(conv_tc_k<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(conv_tc_k<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(conv_tc_k_taps<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(conv_tc_k_taps<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(conv_tc_k_taps<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(conv_tc_k_taps<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big); });
    const size_t stat_doubles = (size_t)CHAP_STAT_SLOTS * 2 * (p.mode == 1 ? g.cout : p.n_real);
    if (bn) {
        CHAP_REQUIRE(ch_sums != nullptr, CHAP_ERR_BAD_ARG, "tc_conv: the folded BatchNorm finalize needs the statistics buffer");
        p.bn = *bn;
        p.bn.count = (double)g.out_rows;
        p.bn.counter = reinterpret_cast<unsigned*>(ch_sums + stat_doubles);
    }
    if (ch_sums && !(bn && bn->rezero)) CHAP_TRY(zero_async(ch_sums, (stat_doubles + (bn ? 1 : 0)) * sizeof(double), st));
    const double rows = (double)(g.kind == CHAP_CONV_UP2 ? g.in_rows : g.out_rows);
    KernelTimer timer(timer_name(p.precise ? (dgrad ? "conv_tc3x_dgrad" : "conv_tc3x_fwd") : (dgrad ? "conv_tc_dgrad" : "conv_tc_fwd"), g.taps,
                                 dgrad ? g.cout : g.cin, dgrad ? g.cin : g.cout, g.iW, g.iH, g.iD, g.in_rows),       // real channel counts
                      2.0 * rows * g.cin * g.cout * g.taps,
                      4.0 * ((double)g.in_rows * g.cin + (double)g.out_rows * g.cout + (double)g.taps * g.cin * g.cout), st);
    dim3 grid((unsigned)grid_x, (unsigned)(N / p.nt));
    const unsigned threads = 64 + 128 * p.epi_groups + (p.precise ? 128 : 0);
    const bool ev = epi != nullptr;
#define CHAP_TC_LAUNCH(PR, EV) do { if (p.mode == 3) launch_k(conv_tc_k_taps<PR, EV>, grid, threads, smem, st, tmA, tmB, tmT, p); \
                                    else launch_k(conv_tc_k<PR, EV>, grid, threads, smem, st, tmA, tmB, p); } while (0)
    if (p.precise) { if (ev) CHAP_TC_LAUNCH(true, true); else CHAP_TC_LAUNCH(true, false); }
    else           { if (ev) CHAP_TC_LAUNCH(false, true); else CHAP_TC_LAUNCH(false, false); }
#undef CHAP_TC_LAUNCH
    CHAP_TRY(launched("conv_tc_kernel"));
    return 1;
}

}  // namespace chap

#ifdef CHAP_TC_DEBUG_HOOKS
extern "C" int chap_debug_tc_trace(long long* out8) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out8, chap::g_tc_trace, 8 * sizeof(long long)) == cudaSuccess ? 0 : -1;
}
#endif
