// The CHAP perturbation generator: per-sample channel-wise and spatial-wise L2 normalisation of the
// gradient of the consistency loss w.r.t. each encoder level, eps-scaled injection into the
// features (BASELINE.json north_star; frozen spec: oracle/chap_losses.py perturbation()).
//
// Layout: g, f, out are channels-last [n, rows, c]; the spatial-wise norm (over c) of a position is
// local to one row (contiguous words), the channel-wise norm (over rows) and the per-sample norm are
// grid-wide reductions.  Round 1 ran three streaming passes per level over g:
//   P1  chan_sq[n, c]   = sum_rows g^2                                  (skipped for SAMPLE / SPATIAL)
//   P2  samp_sq[n]      = sum u^2,  u = combine(g / (||g||_chan + e), g / (||g||_row + e))
//   P3  out             = f + eps * u / (sqrt(samp_sq) + e)
// Round 2 (default): P2 is gone.  With S_c = sum_r g_rc^2, the row norms nr_r = sqrt(sum_c g_rc^2) + e being LOCAL to a row, and
//   Q_c = sum_r g_rc^2 / nr_r,   Z = sum_r (sum_c g_rc^2) / nr_r^2
// the per-sample norm of the combined field is a function of these per-channel sums alone:
//   ||u||^2 = 1/4 [ sum_c S_c / nc_c^2  +  Z  +  2 sum_c Q_c / nc_c ],     nc_c = sqrt(S_c) + e
// so ONE statistics pass (S, Q, Z) and ONE apply pass remain -- the minimum for a computation with one dependent grid-wide
// reduction: 16 B / element touched (g twice, f, out) for 12 B algorithmic, and every pass runs for ALL levels and samples in ONE
// launch through a descriptor table (round 1: 3 launches + a zero-fill per level = 16 per call; now 2 + 1).
// Measured, all five 2D levels of 12 samples (97 MB of g; tools/perturb_bench.py): three-phase 137.4 us, two-pass 129.0 us (121.3 us once the
// statistics pass got its own grid of ~16 rows per thread)
// (statistics 45 us + apply 69 us = 4.2 TB/s over its 12 B / element); 3D b2: 222 -> 199 us.  The statistics pass keeps <= 16
// channel accumulators x (S, Q) per thread and folds them with warp shuffles (a shared-memory fold over the row lanes made it
// SLOWER than three-phase: 165 us).  CHAP_PERTURB_3PHASE=1 selects the three-phase path (also the fallback for c > 256).
#include "common.cuh"

namespace chap {

constexpr float kEps = 1e-8f;

// P1: per-(sample, channel) sum of squares.  grid = (blocks_per_sample, n)
// one level of one call, as the batched kernels see it
struct PLevel {
    const float* g; const float* f; float* out;
    double* chan_sq; double* samp_sq;       // S[n][c]; (three-phase path) ||u||^2 per sample / (batched l2n) ||d||^2
    double* chan_q; double* row_z;          // Q[n][c], Z[n] of the two-pass scheme
    int64_t rows;
    int c, tpr;
    int bps_chan, bps_rows;          // blocks per sample in the channel-norm phase / the row phases
    int blk0_chan, blk0_rows;        // first block of this level in the flattened grids
    int bps_stat, blk0_stat;         // the same for the statistics pass of the two-pass scheme (~16 rows per thread: its fold + 2c double
                                     // atomics per block are a fixed cost that 2 rows per thread did not amortise)
};
constexpr int kMaxLevels = 8;
struct PBatch { PLevel lv[kMaxLevels]; int n_levels, n; float eps, gs; };

template <int VEC>
__device__ __forceinline__ void chan_sq_body(const float* __restrict__ g, int64_t rows, int c, float gs, double* __restrict__ chan_sq,
                                             int bx, int nbx, int sample, float* part) {
    const int cg = c / VEC, rpb = 256 / cg;
    const int gi = threadIdx.x % cg, rl = threadIdx.x / cg;
    const float* base = g + (int64_t)sample * rows * c;
    float q[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) q[v] = 0.f;
    if (rl < rpb)
        for (int64_t r = (int64_t)bx * rpb + rl; r < rows; r += (int64_t)nbx * rpb) {
            if (VEC == 4) {
                float4 t = __ldg(reinterpret_cast<const float4*>(base + r * c) + gi);
                t.x *= gs; t.y *= gs; t.z *= gs; t.w *= gs;
                q[0] += t.x * t.x; q[1 % VEC] += t.y * t.y; q[2 % VEC] += t.z * t.z; q[3 % VEC] += t.w * t.w;
            } else { float t = base[r * c + gi] * gs; q[0] += t * t; }
        }
#pragma unroll
    for (int v = 0; v < VEC; ++v) part[threadIdx.x * VEC + v] = q[v];
    __syncthreads();
    if (threadIdx.x < cg) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            double a = 0.0;
            for (int l = 0; l < rpb; ++l) a += (double)part[(l * cg + threadIdx.x) * VEC + v];
            atomicAdd(chan_sq + (int64_t)sample * c + threadIdx.x * VEC + v, a);
        }
    }
}
// P1 for all levels: flattened grid, block -> (level, sample, block in sample)
__global__ void __launch_bounds__(256)
chan_sq_all_kernel(const __grid_constant__ PBatch b) {
    pdl_enter();
    __shared__ float part[256 * 4];
    int l = 0;
    while (l + 1 < b.n_levels && (int)blockIdx.x >= b.lv[l + 1].blk0_chan) ++l;
    const PLevel& L = b.lv[l];
    const int rel = (int)blockIdx.x - L.blk0_chan;
    chan_sq_body<4>(L.g, L.rows, L.c, b.gs, L.chan_sq, rel % L.bps_chan, L.bps_chan, rel / L.bps_chan, part);
}

// u for one row.  One WARP handles one row when c >= 128 elements would not fit a thread; here a
// thread group of `tpr` lanes (power of two <= 32) covers one row, each lane c / tpr channels.
template <int MODE>
__device__ __forceinline__ float unit_value(float gv, float inv_chan, float inv_row) {
    if (MODE == CHAP_PERTURB_SAMPLE) return gv;
    if (MODE == CHAP_PERTURB_CHANNEL) return gv * inv_chan;
    if (MODE == CHAP_PERTURB_SPATIAL) return gv * inv_row;
    return 0.5f * (gv * inv_chan + gv * inv_row);
}

// P2 / P3 share the row traversal.  grid = (blocks_per_sample, n); block = 256 threads = 256/tpr rows.
// smem: inv_chan[c]
template <int MODE, bool APPLY>
__device__ __forceinline__ void perturb_rows_body(const float* __restrict__ g, const float* __restrict__ f, float* __restrict__ out,
                                                  int64_t rows, int c, int tpr, float eps, float gs, int rt,
                                                  const double* __restrict__ chan_sq, double* __restrict__ samp_sq, int samp_idx,
                                                  int bx, int nbx, int n, float* inv_chan, float* red) {
    if (MODE == CHAP_PERTURB_CHANNEL || MODE == CHAP_PERTURB_CHANNEL_SPATIAL) {
        for (int ch = threadIdx.x; ch < c; ch += 256)
            inv_chan[ch] = 1.f / (sqrtf((float)chan_sq[(int64_t)n * c + ch]) + kEps);
        __syncthreads();
    }
    float scale = 0.f;
    if (APPLY) scale = eps / (sqrtf((float)samp_sq[samp_idx]) + kEps);
    const int lane = threadIdx.x % tpr, rl = threadIdx.x / tpr, rpb = 256 / tpr;
    const int cpl = c / tpr;                       // channels per lane (multiple of 4 when c % (4 tpr) == 0)
    const float* gb = g + (int64_t)n * rows * c;
    const float* fb = f ? f + (int64_t)n * rows * c : nullptr;
    float* ob = out + (int64_t)n * rows * c;
    float acc = 0.f;
    for (int64_t r0 = (int64_t)bx * rpb; r0 < rows; r0 += (int64_t)nbx * rpb) {
        const int64_t r = r0 + rl;
        const bool live = r < rows;
        float inv_row = 0.f;
        // the row's slice of g is loaded ONCE into registers (<= 16 channels per lane) and serves both the row norm and u
        // (round 1 loaded it twice: the second, dependent load made the reduce phase latency-bound at 28 % of HBM under ncu)
        const bool in_regs = cpl <= 16;
        float4 tg[4];
        if (live && in_regs) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
                if (4 * k4 < cpl) {
                    float4 t = __ldg(reinterpret_cast<const float4*>(gb + r * c + lane * cpl) + k4);
                    t.x *= gs; t.y *= gs; t.z *= gs; t.w *= gs;
                    tg[k4] = t;
                }
        }
        if (MODE == CHAP_PERTURB_SPATIAL || MODE == CHAP_PERTURB_CHANNEL_SPATIAL) {
            float q = 0.f;
            if (live) {
                if (in_regs) {
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        if (4 * k4 < cpl) q += tg[k4].x * tg[k4].x + tg[k4].y * tg[k4].y + tg[k4].z * tg[k4].z + tg[k4].w * tg[k4].w;
                } else {
                    for (int k = 0; k < cpl; k += 4) {
                        float4 t = __ldg(reinterpret_cast<const float4*>(gb + r * c + lane * cpl + k));
                        t.x *= gs; t.y *= gs; t.z *= gs; t.w *= gs;
                        q += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
                    }
                }
            }
            for (int o = tpr >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            inv_row = 1.f / (sqrtf(q) + kEps);
        }
        if (!live) continue;
        auto emit = [&](float4 t, const int k) {            // u for 4 channels of this row: accumulate ||u||^2 or write f + scale * u
            const int ch = lane * cpl + k;
            float ic0 = 0.f, ic1 = 0.f, ic2 = 0.f, ic3 = 0.f;
            if (MODE == CHAP_PERTURB_CHANNEL || MODE == CHAP_PERTURB_CHANNEL_SPATIAL) {
                ic0 = inv_chan[ch]; ic1 = inv_chan[ch + 1]; ic2 = inv_chan[ch + 2]; ic3 = inv_chan[ch + 3];
            }
            float u0 = unit_value<MODE>(t.x, ic0, inv_row), u1 = unit_value<MODE>(t.y, ic1, inv_row);
            float u2 = unit_value<MODE>(t.z, ic2, inv_row), u3 = unit_value<MODE>(t.w, ic3, inv_row);
            if (APPLY) {
                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                if (fb) b = ldg_stream(reinterpret_cast<const float4*>(fb + r * c + ch));
                b.x = fmaf(scale, u0, b.x); b.y = fmaf(scale, u1, b.y);
                b.z = fmaf(scale, u2, b.z); b.w = fmaf(scale, u3, b.w);
                b.x = tf32_rn(b.x, rt); b.y = tf32_rn(b.y, rt); b.z = tf32_rn(b.z, rt); b.w = tf32_rn(b.w, rt);
                *reinterpret_cast<float4*>(ob + r * c + ch) = b;
            } else {
                acc += u0 * u0 + u1 * u1 + u2 * u2 + u3 * u3;
            }
        };
        if (in_regs) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)                  // static indices: tg stays in registers
                if (4 * k4 < cpl) emit(tg[k4], 4 * k4);
        } else {
            for (int k = 0; k < cpl; k += 4) {
                float4 t = __ldg(reinterpret_cast<const float4*>(gb + r * c + lane * cpl + k));
                t.x *= gs; t.y *= gs; t.z *= gs; t.w *= gs;
                emit(t, k);
            }
        }
    }
    if (!APPLY) {
        float t = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0;
            for (int w = 0; w < 8; ++w) a += (double)red[w];
            atomicAdd(samp_sq + samp_idx, a);
        }
    }
}
// P2 (APPLY = false) / P3 (APPLY = true) for all levels in one launch
template <int MODE, bool APPLY>
__global__ void __launch_bounds__(256)
perturb_rows_all_kernel(const __grid_constant__ PBatch b) {
    pdl_enter();
    __shared__ float inv_chan[1024];
    __shared__ float red[8];
    int l = 0;
    while (l + 1 < b.n_levels && (int)blockIdx.x >= b.lv[l + 1].blk0_rows) ++l;
    const PLevel& L = b.lv[l];
    const int rel = (int)blockIdx.x - L.blk0_rows;
    perturb_rows_body<MODE, APPLY>(L.g, APPLY ? L.f : nullptr, APPLY ? L.out : nullptr, L.rows, L.c, L.tpr, b.eps, b.gs, 0, L.chan_sq, L.samp_sq,
                                   rel / L.bps_rows, rel % L.bps_rows, L.bps_rows, rel / L.bps_rows, inv_chan, red);
}

// Statistics pass of the two-pass scheme, all levels in one launch: per sample S_c, Q_c (double atomics, one per channel and block) and Z.
// Row layout of the apply pass (tpr lanes per row, <= 16 channels per lane); a thread walks many rows and keeps its 2 x cpl partial
// sums in registers, the block folds them through shared memory once at the end.
template <int MODE>
__global__ void __launch_bounds__(256)
perturb_stats_all_kernel(const __grid_constant__ PBatch b) {
    pdl_enter();
    __shared__ float fold[2][8][256];              // [S | Q][warp][channel]: warp totals (c <= 256 on this path)
    __shared__ float redz[8];
    int l = 0;
    while (l + 1 < b.n_levels && (int)blockIdx.x >= b.lv[l + 1].blk0_stat) ++l;
    const PLevel& L = b.lv[l];
    const int rel = (int)blockIdx.x - L.blk0_stat, bx = rel % L.bps_stat, nbx = L.bps_stat, n = rel / L.bps_stat;
    const int c = L.c, tpr = L.tpr, cpl = c / tpr, rpb = 256 / tpr;
    const int lane = threadIdx.x % tpr, rl = threadIdx.x / tpr;
    const float* gb = L.g + (int64_t)n * L.rows * c;
    float sacc[16], qacc[16], zacc = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { sacc[j] = 0.f; qacc[j] = 0.f; }
    for (int64_t r0 = (int64_t)bx * rpb; r0 < L.rows; r0 += (int64_t)nbx * rpb) {
        const int64_t r = r0 + rl;
        const bool live = r < L.rows;
        float4 tg[4];
        float q = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
            if (4 * k4 < cpl) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live) { t = ldg_stream(reinterpret_cast<const float4*>(gb + r * c + lane * cpl) + k4); t.x *= b.gs; t.y *= b.gs; t.z *= b.gs; t.w *= b.gs; }
                t.x *= t.x; t.y *= t.y; t.z *= t.z; t.w *= t.w;          // squares from here on
                tg[k4] = t;
                q += t.x + t.y + t.z + t.w;
            }
        for (int o = tpr >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);      // R_r: the row's sum of squares
        const float inv = 1.f / (sqrtf(q) + kEps);                                           // 1 / nr_r
        if (live && lane == 0) zacc = fmaf(q * inv, inv, zacc);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
            if (4 * k4 < cpl) {
                sacc[4 * k4] += tg[k4].x; sacc[4 * k4 + 1] += tg[k4].y; sacc[4 * k4 + 2] += tg[k4].z; sacc[4 * k4 + 3] += tg[k4].w;
                qacc[4 * k4] = fmaf(tg[k4].x, inv, qacc[4 * k4]); qacc[4 * k4 + 1] = fmaf(tg[k4].y, inv, qacc[4 * k4 + 1]);
                qacc[4 * k4 + 2] = fmaf(tg[k4].z, inv, qacc[4 * k4 + 2]); qacc[4 * k4 + 3] = fmaf(tg[k4].w, inv, qacc[4 * k4 + 3]);
            }
    }
    // fold over the block's row lanes: inside a warp the threads that own the same channel slice are tpr apart -> xor shuffles over
    // the offsets >= tpr; the 8 warp totals meet in shared memory
    (void)rl; (void)rpb;
    const int wl = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 16; ++j)
        if (j < cpl) {
            float s1 = sacc[j], q1 = qacc[j];
            for (int o = 16; o >= tpr; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o); }
            if (wl < tpr) { fold[0][wid][lane * cpl + j] = s1; fold[1][wid][lane * cpl + j] = q1; }
        }
    float z = warp_sum(zacc);
    if ((threadIdx.x & 31) == 0) redz[threadIdx.x >> 5] = z;
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += 256) {
        double s1 = 0.0, q1 = 0.0;
        for (int w = 0; w < 8; ++w) { s1 += (double)fold[0][w][ch]; q1 += (double)fold[1][w][ch]; }
        atomicAdd(L.chan_sq + (int64_t)n * c + ch, s1);
        if (MODE == CHAP_PERTURB_CHANNEL_SPATIAL) atomicAdd(L.chan_q + (int64_t)n * c + ch, q1);
    }
    if (threadIdx.x == 0 && (MODE == CHAP_PERTURB_SPATIAL || MODE == CHAP_PERTURB_CHANNEL_SPATIAL)) {
        double zz = 0.0;
        for (int w = 0; w < 8; ++w) zz += (double)redz[w];
        atomicAdd(L.row_z + n, zz);
    }
}

// Apply pass of the two-pass scheme: every block first derives ||u||^2 of its sample from S, Q, Z (c <= 1024 channels: a block
// reduction over at most 4 values per thread), then streams g and f once.
template <int MODE>
__global__ void __launch_bounds__(256)
perturb_apply_all_kernel(const __grid_constant__ PBatch b) {
    pdl_enter();
    __shared__ float inv_chan[1024];
    __shared__ float red[8];
    __shared__ double part[8];
    __shared__ double samp_local;
    int l = 0;
    while (l + 1 < b.n_levels && (int)blockIdx.x >= b.lv[l + 1].blk0_rows) ++l;
    const PLevel& L = b.lv[l];
    const int rel = (int)blockIdx.x - L.blk0_rows, n = rel / L.bps_rows;
    double acc = 0.0;
    for (int ch = threadIdx.x; ch < L.c; ch += 256) {
        const double S = L.chan_sq[(int64_t)n * L.c + ch];
        const float ic = 1.f / (sqrtf((float)S) + kEps);
        if (MODE == CHAP_PERTURB_SAMPLE) acc += S;
        else if (MODE == CHAP_PERTURB_CHANNEL) acc += S * (double)ic * (double)ic;
        else if (MODE == CHAP_PERTURB_CHANNEL_SPATIAL) acc += S * (double)ic * (double)ic + 2.0 * L.chan_q[(int64_t)n * L.c + ch] * (double)ic;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += part[w];
        if (MODE == CHAP_PERTURB_SPATIAL) t = L.row_z[n];
        else if (MODE == CHAP_PERTURB_CHANNEL_SPATIAL) t = 0.25 * (t + L.row_z[n]);
        samp_local = t;
    }
    __syncthreads();
    perturb_rows_body<MODE, true>(L.g, L.f, L.out, L.rows, L.c, L.tpr, b.eps, b.gs, 0, L.chan_sq, &samp_local, 0, rel % L.bps_rows, L.bps_rows, n, inv_chan, red);
}

template <int MODE>
static int run_two_pass(const PBatch& b, int blocks_stats, int blocks_rows, double alg_bytes, cudaStream_t st) {
    KernelTimer timer("perturb_level", 0.0, alg_bytes, st);       // algorithmic: read g, read f, write out (all levels of the call)
    launch_k(perturb_stats_all_kernel<MODE>, blocks_stats, 256, 0, st, b);
    CHAP_TRY(launched("perturb_stats_all_kernel"));
    launch_k(perturb_apply_all_kernel<MODE>, blocks_rows, 256, 0, st, b);
    return launched("perturb_apply_all_kernel");
}

template <int MODE>
static int run_all(const PBatch& b, int blocks_chan, int blocks_rows, double alg_bytes, cudaStream_t st) {
    KernelTimer timer("perturb_level", 0.0, alg_bytes, st);       // algorithmic: read g, read f, write out (all levels of the call)
    if (MODE == CHAP_PERTURB_CHANNEL || MODE == CHAP_PERTURB_CHANNEL_SPATIAL) {
        launch_k(chan_sq_all_kernel, blocks_chan, 256, 0, st, b);
        CHAP_TRY(launched("chan_sq_all_kernel"));
    }
    launch_k(perturb_rows_all_kernel<MODE, false>, blocks_rows, 256, 0, st, b);
    CHAP_TRY(launched("perturb_rows_all_kernel<reduce>"));
    launch_k(perturb_rows_all_kernel<MODE, true>, blocks_rows, 256, 0, st, b);
    return launched("perturb_rows_all_kernel<apply>");
}

// out = base + xi * d / (||d|| + 1e-8) per sample
__global__ void __launch_bounds__(256)
sample_sq_kernel(const float* __restrict__ d, int64_t eps_, double* __restrict__ norms) {
    pdl_enter();
    const float* b = d + (int64_t)blockIdx.y * eps_;
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < eps_; i += (int64_t)gridDim.x * blockDim.x) {
        float v = b[i]; acc += v * v;
    }
    __shared__ float red[8];
    float t = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int w = 0; w < 8; ++w) a += (double)red[w];
        atomicAdd(norms + blockIdx.y, a);
    }
}
__global__ void __launch_bounds__(256)
l2n_axpy_kernel(const float* __restrict__ d, const float* __restrict__ base, float xi, int rt, int64_t eps_,
                const double* __restrict__ norms, float* __restrict__ out) {
    pdl_enter();
    const int64_t off = (int64_t)blockIdx.y * eps_;
    const float s = xi / (sqrtf((float)norms[blockIdx.y]) + kEps);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < eps_; i += (int64_t)gridDim.x * blockDim.x)
        out[off + i] = tf32_rn(fmaf(s, d[off + i], base ? base[off + i] : 0.f), rt);
}

// the same two phases for several tensors at once (the five levels of the VAT probe): PLevel.g = d, .f = base, .samp_sq = norms
__global__ void __launch_bounds__(256)
sample_sq_all_kernel(const __grid_constant__ PBatch b) {
    pdl_enter();
    int l = 0;
    while (l + 1 < b.n_levels && (int)blockIdx.x >= b.lv[l + 1].blk0_rows) ++l;
    const PLevel& L = b.lv[l];
    const int rel = (int)blockIdx.x - L.blk0_rows, bx = rel % L.bps_rows, n = rel / L.bps_rows;
    const int64_t eps_ = L.rows * L.c;
    const float4* d4 = reinterpret_cast<const float4*>(L.g + (int64_t)n * eps_);
    float acc = 0.f;
    for (int64_t i = (int64_t)bx * 256 + threadIdx.x; i < eps_ / 4; i += (int64_t)L.bps_rows * 256) {
        const float4 v = __ldg(d4 + i);
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    __shared__ float red[8];
    float t = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int w = 0; w < 8; ++w) a += (double)red[w];
        atomicAdd(L.samp_sq + n, a);
    }
}
__global__ void __launch_bounds__(256)
l2n_axpy_all_kernel(const __grid_constant__ PBatch b) {
    pdl_enter();
    int l = 0;
    while (l + 1 < b.n_levels && (int)blockIdx.x >= b.lv[l + 1].blk0_rows) ++l;
    const PLevel& L = b.lv[l];
    const int rel = (int)blockIdx.x - L.blk0_rows, bx = rel % L.bps_rows, n = rel / L.bps_rows;
    const int64_t eps_ = L.rows * L.c, off = (int64_t)n * eps_;
    const float s = b.eps / (sqrtf((float)L.samp_sq[n]) + kEps);              // b.eps carries xi here
    const float4* d4 = reinterpret_cast<const float4*>(L.g + off);
    const float4* f4 = L.f ? reinterpret_cast<const float4*>(L.f + off) : nullptr;
    float4* o4 = reinterpret_cast<float4*>(L.out + off);
    for (int64_t i = (int64_t)bx * 256 + threadIdx.x; i < eps_ / 4; i += (int64_t)L.bps_rows * 256) {
        const float4 v = ldg_stream(d4 + i);
        float4 o = f4 ? ldg_stream(f4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        o.x = fmaf(s, v.x, o.x); o.y = fmaf(s, v.y, o.y); o.z = fmaf(s, v.z, o.z); o.w = fmaf(s, v.w, o.w);
        o4[i] = o;
    }
}

}  // namespace chap

using namespace chap;

extern "C" int chap_l2n_sample_axpy_batched(const chap_level* levels, int32_t n_levels, int32_t n, float xi, double* norms, void* stream) {
    CHAP_REQUIRE(levels && n_levels > 0 && n_levels <= kMaxLevels && n > 0 && norms, CHAP_ERR_BAD_ARG, "l2n_sample_axpy_batched: bad argument");
    cudaStream_t st = S(stream);
    static thread_local PBatch b;
    b.n_levels = n_levels; b.n = n; b.eps = xi; b.gs = 1.f;
    double total = 0.0;
    for (int l = 0; l < n_levels; ++l) total += (double)levels[l].rows * levels[l].c;
    int blocks = 0;
    for (int l = 0; l < n_levels; ++l) {
        const chap_level& L = levels[l];
        CHAP_REQUIRE(L.g && L.out && L.rows > 0 && L.c > 0 && (L.rows * L.c) % 4 == 0, CHAP_ERR_BAD_ARG, "l2n_sample_axpy_batched: level %d: bad shape / pointer", l);
        CHAP_REQUIRE(aligned16(L.g) && aligned16(L.out) && (!L.f || aligned16(L.f)), CHAP_ERR_ALIGNMENT, "l2n_sample_axpy_batched: level %d misaligned", l);
        PLevel& P = b.lv[l];
        P.g = L.g; P.f = L.f; P.out = L.out; P.rows = L.rows; P.c = L.c; P.tpr = 1; P.chan_sq = nullptr;
        P.samp_sq = norms + (size_t)l * n;
        const int share = (int)(kNumSMs * 16 * ((double)L.rows * L.c / total) / n) + 1;
        int bps = (int)((L.rows * L.c / 4 + 256 * 4 - 1) / (256 * 4));
        if (bps > share) bps = share;
        if (bps < 1) bps = 1;
        P.bps_rows = bps; P.bps_chan = 0; P.blk0_rows = blocks; P.blk0_chan = 0;
        blocks += bps * n;
    }
    KernelTimer timer_("l2n_sample_axpy", 0.0, 12.0 * total * n, st);        // algorithmic: read d, read base, write out
    CHAP_TRY(zero_async(norms, (size_t)n_levels * n * sizeof(double), st));
    launch_k(sample_sq_all_kernel, blocks, 256, 0, st, b);
    CHAP_TRY(launched("sample_sq_all_kernel"));
    launch_k(l2n_axpy_all_kernel, blocks, 256, 0, st, b);
    return launched("l2n_axpy_all_kernel");
}

extern "C" size_t chap_perturb_workspace_elems(const chap_level* levels, int32_t n_levels, int32_t n) {
    size_t total = 0;
    for (int l = 0; l < n_levels; ++l) total += 2 * ((size_t)n * levels[l].c + (size_t)n);      // S, ||u||^2, Q, Z per level
    return total;
}

extern "C" int chap_perturb_fwd(const chap_level* levels, int32_t n_levels, int32_t n, int32_t mode, float eps, float g_scale,
                                double* workspace, size_t ws_elems, void* stream) {
    CHAP_REQUIRE(levels && n_levels > 0 && n > 0 && workspace, CHAP_ERR_BAD_ARG, "perturb_fwd: bad argument");
    CHAP_REQUIRE(mode >= CHAP_PERTURB_SAMPLE && mode <= CHAP_PERTURB_CHANNEL_SPATIAL, CHAP_ERR_BAD_ARG, "perturb_fwd: unknown mode %d", mode);
    CHAP_REQUIRE(ws_elems >= chap_perturb_workspace_elems(levels, n_levels, n), CHAP_ERR_WORKSPACE, "perturb_fwd: workspace too small");
    CHAP_REQUIRE(n_levels <= kMaxLevels, CHAP_ERR_BAD_ARG, "perturb_fwd: at most %d levels per call (got %d)", kMaxLevels, n_levels);
    cudaStream_t st = S(stream);
    CHAP_TRY(zero_async(workspace, chap_perturb_workspace_elems(levels, n_levels, n) * sizeof(double), st));
    static thread_local PBatch b;
    b.n_levels = n_levels; b.n = n; b.eps = eps; b.gs = g_scale;
    double* ws = workspace;
    int blocks_chan = 0, blocks_rows = 0, blocks_stat = 0;
    double alg_bytes = 0.0;
    bool two_pass_ok = true;
    // blocks per (level, sample): enough to fill the machine a few times over across ALL levels of the launch
    double total_elems = 0.0;
    for (int l = 0; l < n_levels; ++l) total_elems += (double)levels[l].rows * levels[l].c;
    const int block_budget = kNumSMs * 16;
    for (int l = 0; l < n_levels; ++l) {
        const chap_level& L = levels[l];
        CHAP_REQUIRE(L.g && L.out && L.rows > 0 && L.c > 0, CHAP_ERR_BAD_ARG, "perturb_fwd: level %d has a NULL pointer or empty shape", l);
        CHAP_REQUIRE(L.c % 4 == 0 && L.c <= 1024, CHAP_ERR_BAD_ARG, "perturb_fwd: level %d channel count %d must be a multiple of 4", l, L.c);
        CHAP_REQUIRE(aligned16(L.g) && aligned16(L.out) && (!L.f || aligned16(L.f)), CHAP_ERR_ALIGNMENT, "perturb_fwd: level %d misaligned", l);
        PLevel& P = b.lv[l];
        P.g = L.g; P.f = L.f; P.out = L.out; P.rows = L.rows; P.c = L.c;
        P.chan_sq = ws; ws += (size_t)n * L.c;
        P.samp_sq = ws; ws += n;
        P.chan_q = ws; ws += (size_t)n * L.c;
        P.row_z = ws; ws += n;
        int tpr = 1;                                   // lanes per row: keep <= 16 channels (4 float4) per lane
        while (tpr < 32 && L.c / tpr > 16 && (L.c / (tpr * 2)) % 4 == 0) tpr *= 2;
        P.tpr = tpr;
        if (L.c / tpr > 16 || L.c > 256) two_pass_ok = false;            // the statistics pass keeps <= 16 channels per lane in registers
        const int share = (int)(block_budget * ((double)L.rows * L.c / total_elems) / n) + 1;      // this level's share of the grid, per sample
        const int rpb = 256 / tpr;
        int bps = (int)((L.rows + rpb * 4 - 1) / (rpb * 4));
        if (bps > share) bps = share;
        if (bps < 1) bps = 1;
        const int cg = L.c / 4, rpb1 = cg <= 256 ? 256 / cg : 1;
        int b1 = (int)((L.rows + rpb1 * 8 - 1) / (rpb1 * 8));
        if (b1 > share) b1 = share;
        if (b1 < 1) b1 = 1;
        static const int stat_rows = getenv("CHAP_PERTURB_STAT_ROWS") ? atoi(getenv("CHAP_PERTURB_STAT_ROWS")) : 16;      // measured: 2 rows 130.6, 8 rows 124.9, 16 rows 121.3, 32 rows 142.9 us per call
        int b2 = (int)((L.rows + (int64_t)rpb * stat_rows - 1) / ((int64_t)rpb * stat_rows));
        if (b2 > share) b2 = share;
        if (b2 < 1) b2 = 1;
        P.bps_rows = bps; P.bps_chan = b1; P.bps_stat = b2;
        P.blk0_rows = blocks_rows; P.blk0_chan = blocks_chan; P.blk0_stat = blocks_stat;
        blocks_rows += bps * n; blocks_chan += b1 * n; blocks_stat += b2 * n;
        alg_bytes += 12.0 * (double)n * L.rows * L.c;
    }
    static const bool three_phase = getenv("CHAP_PERTURB_3PHASE") != nullptr;
    if (two_pass_ok && !three_phase) {
        switch (mode) {
            case CHAP_PERTURB_SAMPLE: return run_two_pass<CHAP_PERTURB_SAMPLE>(b, blocks_stat, blocks_rows, alg_bytes, st);
            case CHAP_PERTURB_CHANNEL: return run_two_pass<CHAP_PERTURB_CHANNEL>(b, blocks_stat, blocks_rows, alg_bytes, st);
            case CHAP_PERTURB_SPATIAL: return run_two_pass<CHAP_PERTURB_SPATIAL>(b, blocks_stat, blocks_rows, alg_bytes, st);
            default: return run_two_pass<CHAP_PERTURB_CHANNEL_SPATIAL>(b, blocks_stat, blocks_rows, alg_bytes, st);
        }
    }
    switch (mode) {
        case CHAP_PERTURB_SAMPLE: return run_all<CHAP_PERTURB_SAMPLE>(b, blocks_chan, blocks_rows, alg_bytes, st);
        case CHAP_PERTURB_CHANNEL: return run_all<CHAP_PERTURB_CHANNEL>(b, blocks_chan, blocks_rows, alg_bytes, st);
        case CHAP_PERTURB_SPATIAL: return run_all<CHAP_PERTURB_SPATIAL>(b, blocks_chan, blocks_rows, alg_bytes, st);
        default: return run_all<CHAP_PERTURB_CHANNEL_SPATIAL>(b, blocks_chan, blocks_rows, alg_bytes, st);
    }
}

extern "C" int chap_l2n_sample_axpy(const float* d, const float* base, float xi, int32_t n, int64_t elems_per_sample,
                                    double* norms, float* out, void* stream) {
    KernelTimer timer_("l2n_sample_axpy", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(d && norms && out && n > 0 && elems_per_sample > 0, CHAP_ERR_BAD_ARG, "l2n_sample_axpy: bad argument");
    cudaStream_t st = S(stream);
    CHAP_TRY(zero_async(norms, (size_t)n * sizeof(double), st));
    int bps = (int)((elems_per_sample + 256 * 8 - 1) / (256 * 8));
    int cap = (kNumSMs * 8 + n - 1) / n;
    if (bps > cap) bps = cap;
    if (bps < 1) bps = 1;
    dim3 grid((unsigned)bps, (unsigned)n);
    launch_k(sample_sq_kernel, grid, 256, 0, st, d, elems_per_sample, norms);
    CHAP_TRY(launched("sample_sq_kernel"));
    launch_k(l2n_axpy_kernel, grid, 256, 0, st, d, base, xi, round_tf32_on(), elems_per_sample, norms, out);
    return launched("l2n_axpy_kernel");
}
