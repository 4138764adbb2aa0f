// The CHAP perturbation generator: per-sample channel-wise and spatial-wise L2 normalisation of the
// gradient of the consistency loss w.r.t. each encoder level, eps-scaled injection into the
// features (BASELINE.json north_star; frozen spec: oracle/chap_losses.py perturbation()).
//
// Layout: g, f, out are channels-last [n, rows, c]; the spatial-wise norm (over c) of a position is
// local to one row (contiguous words), the channel-wise norm (over rows) and the per-sample norm are
// grid-wide reductions -> three streaming passes per level over g:
//   P1  chan_sq[n, c]   = sum_rows g^2                                  (skipped for SAMPLE / SPATIAL)
//   P2  samp_sq[n]      = sum u^2,  u = combine(g / (||g||_chan + e), g / (||g||_row + e))
//   P3  out             = f + eps * u / (sqrt(samp_sq) + e)
// Algorithmic traffic is 12 B / element (read g, read f, write out); this first version re-reads g in
// P2/P3 (20 B / element, the re-reads mostly hit the 126 MB L2 for 2D levels).
#include "common.cuh"

namespace chap {

constexpr float kEps = 1e-8f;

// P1: per-(sample, channel) sum of squares.  grid = (blocks_per_sample, n)
template <int VEC>
__global__ void __launch_bounds__(256)
chan_sq_kernel(const float* __restrict__ g, int64_t rows, int c, float gs, double* __restrict__ chan_sq) {
    __shared__ float part[256 * 4];
    const int cg = c / VEC, rpb = 256 / cg;
    const int gi = threadIdx.x % cg, rl = threadIdx.x / cg;
    const float* base = g + (int64_t)blockIdx.y * rows * c;
    float q[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) q[v] = 0.f;
    if (rl < rpb)
        for (int64_t r = (int64_t)blockIdx.x * rpb + rl; r < rows; r += (int64_t)gridDim.x * rpb) {
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
            atomicAdd(chan_sq + (int64_t)blockIdx.y * c + threadIdx.x * VEC + v, a);
        }
    }
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
__global__ void __launch_bounds__(256)
perturb_rows_kernel(const float* __restrict__ g, const float* __restrict__ f, float* __restrict__ out,
                    int64_t rows, int c, int tpr, float eps, float gs, int rt, const double* __restrict__ chan_sq,
                    double* __restrict__ samp_sq) {
    extern __shared__ float inv_chan[];
    const int n = blockIdx.y;
    if (MODE == CHAP_PERTURB_CHANNEL || MODE == CHAP_PERTURB_CHANNEL_SPATIAL) {
        for (int ch = threadIdx.x; ch < c; ch += 256)
            inv_chan[ch] = 1.f / (sqrtf((float)chan_sq[(int64_t)n * c + ch]) + kEps);
        __syncthreads();
    }
    float scale = 0.f;
    if (APPLY) scale = eps / (sqrtf((float)samp_sq[n]) + kEps);
    const int lane = threadIdx.x % tpr, rl = threadIdx.x / tpr, rpb = 256 / tpr;
    const int cpl = c / tpr;                       // channels per lane (multiple of 4 when c % (4 tpr) == 0)
    const float* gb = g + (int64_t)n * rows * c;
    const float* fb = f ? f + (int64_t)n * rows * c : nullptr;
    float* ob = out + (int64_t)n * rows * c;
    float acc = 0.f;
    for (int64_t r0 = (int64_t)blockIdx.x * rpb; r0 < rows; r0 += (int64_t)gridDim.x * rpb) {
        const int64_t r = r0 + rl;
        const bool live = r < rows;
        float inv_row = 0.f;
        if (MODE == CHAP_PERTURB_SPATIAL || MODE == CHAP_PERTURB_CHANNEL_SPATIAL) {
            float q = 0.f;
            if (live)
                for (int k = 0; k < cpl; k += 4) {
                    float4 t = __ldg(reinterpret_cast<const float4*>(gb + r * c + lane * cpl + k));
                    t.x *= gs; t.y *= gs; t.z *= gs; t.w *= gs;
                    q += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
                }
            for (int o = tpr >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            inv_row = 1.f / (sqrtf(q) + kEps);
        }
        if (!live) continue;
        for (int k = 0; k < cpl; k += 4) {
            const int ch = lane * cpl + k;
            float4 t = __ldg(reinterpret_cast<const float4*>(gb + r * c + ch));
            t.x *= gs; t.y *= gs; t.z *= gs; t.w *= gs;
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
        }
    }
    if (!APPLY) {
        __shared__ float red[8];
        float t = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0;
            for (int w = 0; w < 8; ++w) a += (double)red[w];
            atomicAdd(samp_sq + n, a);
        }
    }
}

template <int MODE>
static int run_level(const chap_level& L, int n, float eps, float gs, double* chan_sq, double* samp_sq, cudaStream_t st) {
    const int c = L.c;
    KernelTimer timer("perturb_level", 0.0, 12.0 * (double)n * L.rows * c, st);   // algorithmic: read g, read f, write out
    int tpr = 1;                                   // lanes per row: keep <= 16 channels (4 float4) per lane
    while (tpr < 32 && c / tpr > 16 && (c / (tpr * 2)) % 4 == 0) tpr *= 2;
    const int rpb = 256 / tpr;
    int bps = (int)((L.rows + rpb * 4 - 1) / (rpb * 4));
    int cap = (kNumSMs * 8 + n - 1) / n;
    if (bps > cap) bps = cap;
    if (bps < 1) bps = 1;
    dim3 grid((unsigned)bps, (unsigned)n);
    const size_t smem = (size_t)c * sizeof(float);
    if (MODE == CHAP_PERTURB_CHANNEL || MODE == CHAP_PERTURB_CHANNEL_SPATIAL) {
        const int cg = c / 4, rpb1 = 256 / cg;
        int b1 = (int)((L.rows + rpb1 * 8 - 1) / (rpb1 * 8));
        if (b1 > cap) b1 = cap;
        if (b1 < 1) b1 = 1;
        chan_sq_kernel<4><<<dim3((unsigned)b1, (unsigned)n), 256, 0, st>>>(L.g, L.rows, c, gs, chan_sq);
        CHAP_TRY(launched("chan_sq_kernel"));
    }
    perturb_rows_kernel<MODE, false><<<grid, 256, smem, st>>>(L.g, nullptr, nullptr, L.rows, c, tpr, eps, gs, 0, chan_sq, samp_sq);
    CHAP_TRY(launched("perturb_rows_kernel<reduce>"));
    perturb_rows_kernel<MODE, true><<<grid, 256, smem, st>>>(L.g, L.f, L.out, L.rows, c, tpr, eps, gs, round_tf32_on(), chan_sq, samp_sq);
    return launched("perturb_rows_kernel<apply>");
}

// out = base + xi * d / (||d|| + 1e-8) per sample
__global__ void __launch_bounds__(256)
sample_sq_kernel(const float* __restrict__ d, int64_t eps_, double* __restrict__ norms) {
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
    const int64_t off = (int64_t)blockIdx.y * eps_;
    const float s = xi / (sqrtf((float)norms[blockIdx.y]) + kEps);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < eps_; i += (int64_t)gridDim.x * blockDim.x)
        out[off + i] = tf32_rn(fmaf(s, d[off + i], base ? base[off + i] : 0.f), rt);
}

}  // namespace chap

using namespace chap;

extern "C" size_t chap_perturb_workspace_elems(const chap_level* levels, int32_t n_levels, int32_t n) {
    size_t total = 0;
    for (int l = 0; l < n_levels; ++l) total += (size_t)n * levels[l].c + (size_t)n;
    return total;
}

extern "C" int chap_perturb_fwd(const chap_level* levels, int32_t n_levels, int32_t n, int32_t mode, float eps, float g_scale,
                                double* workspace, size_t ws_elems, void* stream) {
    CHAP_REQUIRE(levels && n_levels > 0 && n > 0 && workspace, CHAP_ERR_BAD_ARG, "perturb_fwd: bad argument");
    CHAP_REQUIRE(mode >= CHAP_PERTURB_SAMPLE && mode <= CHAP_PERTURB_CHANNEL_SPATIAL, CHAP_ERR_BAD_ARG, "perturb_fwd: unknown mode %d", mode);
    CHAP_REQUIRE(ws_elems >= chap_perturb_workspace_elems(levels, n_levels, n), CHAP_ERR_WORKSPACE, "perturb_fwd: workspace too small");
    cudaStream_t st = S(stream);
    CHAP_TRY(zero_async(workspace, chap_perturb_workspace_elems(levels, n_levels, n) * sizeof(double), st));
    double* ws = workspace;
    for (int l = 0; l < n_levels; ++l) {
        const chap_level& L = levels[l];
        CHAP_REQUIRE(L.g && L.out && L.rows > 0 && L.c > 0, CHAP_ERR_BAD_ARG, "perturb_fwd: level %d has a NULL pointer or empty shape", l);
        CHAP_REQUIRE(L.c % 4 == 0 && L.c <= 1024, CHAP_ERR_BAD_ARG, "perturb_fwd: level %d channel count %d must be a multiple of 4", l, L.c);
        CHAP_REQUIRE(aligned16(L.g) && aligned16(L.out) && (!L.f || aligned16(L.f)), CHAP_ERR_ALIGNMENT, "perturb_fwd: level %d misaligned", l);
        double* chan_sq = ws; ws += (size_t)n * L.c;
        double* samp_sq = ws; ws += n;
        int rc;
        switch (mode) {
            case CHAP_PERTURB_SAMPLE: rc = run_level<CHAP_PERTURB_SAMPLE>(L, n, eps, g_scale, chan_sq, samp_sq, st); break;
            case CHAP_PERTURB_CHANNEL: rc = run_level<CHAP_PERTURB_CHANNEL>(L, n, eps, g_scale, chan_sq, samp_sq, st); break;
            case CHAP_PERTURB_SPATIAL: rc = run_level<CHAP_PERTURB_SPATIAL>(L, n, eps, g_scale, chan_sq, samp_sq, st); break;
            default: rc = run_level<CHAP_PERTURB_CHANNEL_SPATIAL>(L, n, eps, g_scale, chan_sq, samp_sq, st); break;
        }
        CHAP_TRY(rc);
    }
    return CHAP_OK;
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
    sample_sq_kernel<<<grid, 256, 0, st>>>(d, elems_per_sample, norms);
    CHAP_TRY(launched("sample_sq_kernel"));
    l2n_axpy_kernel<<<grid, 256, 0, st>>>(d, base, xi, round_tf32_on(), elems_per_sample, norms, out);
    return launched("l2n_axpy_kernel");
}
