// Bandwidth-bound kernels around the convolutions: BatchNorm statistics / apply / backward fused
// with (Leaky)ReLU, dropout factors and the additive skip; 2x2 max pooling; align_corners x2
// upsampling; channel concat / split; small elementwise helpers.
// All tensors are channels-last rows [rows, C]; 128-bit accesses whenever C % 4 == 0.
#include "common.cuh"
#include "conv_plan.cuh"

namespace chap {

// ------------------------------------------------------------------ per-channel sums
// sums[0:c] += sum_r f(r,c), sums[c:2c] += sum_r g(r,c); thread = (row lane, channel group)
template <int VEC, typename F>
__device__ __forceinline__ void channel_reduce(int64_t rows, int c, double* sums, F&& load2) {
    __shared__ float part[256 * 2 * 4];
    const int cg = c / VEC;                    // channel groups
    const int rpb = 256 / cg;                  // row lanes per block
    const int g = threadIdx.x % cg, rl = threadIdx.x / cg;
    float s[VEC], q[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) s[v] = q[v] = 0.f;
    if (rl < rpb) {
        for (int64_t r = (int64_t)blockIdx.x * rpb + rl; r < rows; r += (int64_t)gridDim.x * rpb)
            load2(r, g, s, q);
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        part[(threadIdx.x * 2 + 0) * VEC + v] = s[v];
        part[(threadIdx.x * 2 + 1) * VEC + v] = q[v];
    }
    __syncthreads();
    if (threadIdx.x < cg) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            double a = 0.0, b = 0.0;
            for (int l = 0; l < rpb; ++l) {
                int t = l * cg + threadIdx.x;
                a += (double)part[(t * 2 + 0) * VEC + v];
                b += (double)part[(t * 2 + 1) * VEC + v];
            }
            atomicAdd(sums + threadIdx.x * VEC + v, a);
            atomicAdd(sums + c + threadIdx.x * VEC + v, b);
        }
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) channel_stats_kernel(const float* __restrict__ y, int64_t rows, int c, double* sums) {
    pdl_enter();
    channel_reduce<VEC>(rows, c, sums, [&](int64_t r, int g, float* s, float* q) {
        if (VEC == 4) {
            float4 v = ldg_stream(reinterpret_cast<const float4*>(y + r * c) + g);
            s[0] += v.x; s[1 % VEC] += v.y; s[2 % VEC] += v.z; s[3 % VEC] += v.w;
            q[0] += v.x * v.x; q[1 % VEC] += v.y * v.y; q[2 % VEC] += v.z * v.z; q[3 % VEC] += v.w * v.w;
        } else {
            float v = y[r * c + g];
            s[0] += v; q[0] += v * v;
        }
    });
}

constexpr int kEwUnroll = 4;
__global__ void __launch_bounds__(256) channel_stats_fixed_kernel(const float* __restrict__ y, int64_t total_vec, int c, double* sums);

int channel_stats(const float* y, int64_t rows, int c, double* sums, cudaStream_t st) {
    CHAP_REQUIRE(y && sums && rows > 0 && c > 0, CHAP_ERR_BAD_ARG, "channel_stats: bad argument");
    CHAP_TRY(zero_async(sums, (size_t)2 * c * sizeof(double), st));
    KernelTimer timer("channel_stats", 0.0, 4.0 * (double)rows * c, st);
    const bool v4 = (c % 4 == 0) && aligned16(y);
    const int cg = v4 ? c / 4 : c;
    CHAP_REQUIRE(cg <= 256, CHAP_ERR_BAD_ARG, "channel_stats: too many channels (%d)", c);
    const int rpb = 256 / cg;
    int grid = grid_for(rows, rpb * 8, kNumSMs * 8);
    if (v4 && 256 % cg == 0) launch_k(channel_stats_fixed_kernel, grid_for(rows * cg, 256 * kEwUnroll * 2, kNumSMs * 8), 256, 0, st, y, rows * cg, c, sums);
    else if (v4) launch_k(channel_stats_kernel<4>, grid, 256, 0, st, y, rows, c, sums);
    else launch_k(channel_stats_kernel<1>, grid, 256, 0, st, y, rows, c, sums);
    return launched("channel_stats_kernel");
}

__global__ void sums_to_float_kernel(const double* sums, float* out, int c, int accumulate) {
    pdl_enter();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c) out[i] = accumulate ? out[i] + (float)sums[i] : (float)sums[i];
}
int sums_to_float(const double* sums, float* out, int c, cudaStream_t st, bool accumulate) {
    launch_k(sums_to_float_kernel, (c + 127) / 128, 128, 0, st, sums, out, c, accumulate ? 1 : 0);
    return launched("sums_to_float_kernel");
}

// ------------------------------------------------------------------ BN finalize
__global__ void bn_finalize_kernel(const double* sums, int slots, int64_t count, const float* gamma, const float* beta,
                                   float eps, float momentum, float* rmean, float* rvar, int64_t* nbt,
                                   float* mean_invstd, float* scale_shift, int c) {
    pdl_enter();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && nbt) *nbt += 1;
    if (i >= c) return;
    double s1 = 0.0, s2 = 0.0;
    for (int s = 0; s < slots; ++s) { s1 += sums[(size_t)s * 2 * c + i]; s2 += sums[(size_t)s * 2 * c + c + i]; }
    double mean = s1 / (double)count;
    double var = s2 / (double)count - mean * mean;
    if (var < 0.0) var = 0.0;
    float invstd = (float)(1.0 / sqrt(var + (double)eps));
    float sc = gamma[i] * invstd;
    mean_invstd[i] = (float)mean; mean_invstd[c + i] = invstd;
    scale_shift[i] = sc; scale_shift[c + i] = beta[i] - (float)mean * sc;
    if (rmean) {
        double unbiased = count > 1 ? var * (double)count / (double)(count - 1) : var;
        rmean[i] = (1.f - momentum) * rmean[i] + momentum * (float)mean;
        rvar[i] = (1.f - momentum) * rvar[i] + momentum * (float)unbiased;
    }
}

__global__ void bn_eval_params_kernel(const float* gamma, const float* beta, const float* rmean, const float* rvar,
                                      float eps, float* mean_invstd, float* scale_shift, int c) {
    pdl_enter();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    float invstd = 1.f / sqrtf(rvar[i] + eps);
    float sc = gamma[i] * invstd;
    mean_invstd[i] = rmean[i]; mean_invstd[c + i] = invstd;
    scale_shift[i] = sc; scale_shift[c + i] = beta[i] - rmean[i] * sc;
}

// ------------------------------------------------------------------ BN apply + act (+drop, +residual)
template <int VEC>
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const float* __restrict__ y, const float* __restrict__ ss, float slope,
                  const float* __restrict__ drop_nc, const float* __restrict__ drop_el,
                  const float* __restrict__ res, int64_t rows_per_sample, int c, int64_t total_vec,
                  int rt, float* __restrict__ out) {
    pdl_enter();
    const int cg = c / VEC;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % cg);
        const int64_t n = i / (rows_per_sample * cg);
        float v[VEC], o[VEC];
        if (VEC == 4) { float4 t = ldg_stream(reinterpret_cast<const float4*>(y) + i); v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w; }
        else v[0] = y[i];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            int ch = g * VEC + k;
            float z = fmaf(v[k], ss[ch], ss[c + ch]);
            float a = z > 0.f ? z : slope * z;
            if (drop_nc) a *= drop_nc[n * c + ch];
            o[k] = a;
        }
        if (drop_el) {
            if (VEC == 4) { float4 t = ldg_stream(reinterpret_cast<const float4*>(drop_el) + i); o[0] *= t.x; o[1 % VEC] *= t.y; o[2 % VEC] *= t.z; o[3 % VEC] *= t.w; }
            else o[0] *= drop_el[i];
        }
        if (res) {
            if (VEC == 4) { float4 t = ldg_stream(reinterpret_cast<const float4*>(res) + i); o[0] += t.x; o[1 % VEC] += t.y; o[2 % VEC] += t.z; o[3 % VEC] += t.w; }
            else o[0] += res[i];
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) o[k] = tf32_rn(o[k], rt);
        if (VEC == 4) reinterpret_cast<float4*>(out)[i] = make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]);
        else out[i] = o[0];
    }
}

// dz for one element
__device__ __forceinline__ float bn_dz(float dout, float yv, float sc, float sh, float slope) {
    float z = fmaf(yv, sc, sh);
    return z > 0.f ? dout : slope * dout;
}

template <int VEC>
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_kernel(const float* __restrict__ dout, const float* __restrict__ y, const float* __restrict__ ss,
                         const float* __restrict__ mi, float slope, const float* __restrict__ drop_nc,
                         const float* __restrict__ drop_el, int64_t rows_per_sample, int64_t rows, int c, double* sums) {
    pdl_enter();
    channel_reduce<VEC>(rows, c, sums, [&](int64_t r, int g, float* s, float* q) {
        const int64_t n = r / rows_per_sample;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            int ch = g * VEC + k;
            int64_t e = r * c + ch;
            float d = dout[e];
            if (drop_nc) d *= drop_nc[n * c + ch];
            if (drop_el) d *= drop_el[e];
            float yv = y[e];
            float dz = bn_dz(d, yv, ss[ch], ss[c + ch], slope);
            float xh = (yv - mi[ch]) * mi[c + ch];
            s[k] += dz; q[k] += dz * xh;
        }
    });
}

template <int VEC>
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ y, const float* __restrict__ ss,
                        const float* __restrict__ mi, float slope,
                        const float* __restrict__ drop_nc, const float* __restrict__ drop_el,
                        int64_t rows_per_sample, int c, int64_t total_vec, int train, int rt, double inv_count,
                        const double* __restrict__ sums, float* __restrict__ dy,
                        float* __restrict__ dgamma, float* __restrict__ dbeta) {
    pdl_enter();
    const int cg = c / VEC;
    const int acc_pg = (train >> 1) & 1;       // bit 1 of `train`: ADD the parameter gradients into dgamma / dbeta (gradient-sink mode)
    train &= 1;
    if (blockIdx.x == 0 && dgamma) {
        for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
            const float db = (float)sums[ch], dg = (float)sums[c + ch];
            dbeta[ch] = acc_pg ? dbeta[ch] + db : db;
            dgamma[ch] = acc_pg ? dgamma[ch] + dg : dg;
        }
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % cg);
        const int64_t n = i / (rows_per_sample * cg);
        float d[VEC], yv[VEC], o[VEC];
        if (VEC == 4) {
            float4 t = ldg_stream(reinterpret_cast<const float4*>(dout) + i); d[0] = t.x; d[1 % VEC] = t.y; d[2 % VEC] = t.z; d[3 % VEC] = t.w;
            float4 u = ldg_stream(reinterpret_cast<const float4*>(y) + i); yv[0] = u.x; yv[1 % VEC] = u.y; yv[2 % VEC] = u.z; yv[3 % VEC] = u.w;
        } else { d[0] = dout[i]; yv[0] = y[i]; }
        if (drop_el) {
            if (VEC == 4) { float4 t = ldg_stream(reinterpret_cast<const float4*>(drop_el) + i); d[0] *= t.x; d[1 % VEC] *= t.y; d[2 % VEC] *= t.z; d[3 % VEC] *= t.w; }
            else d[0] *= drop_el[i];
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            int ch = g * VEC + k;
            float dd = d[k];
            if (drop_nc) dd *= drop_nc[n * c + ch];
            float sc = ss[ch];
            float dz = bn_dz(dd, yv[k], sc, ss[c + ch], slope);
            if (train) {
                float xh = (yv[k] - mi[ch]) * mi[c + ch];
                float mean_dz = (float)(sums[ch] * inv_count), mean_dzx = (float)(sums[c + ch] * inv_count);
                o[k] = sc * (dz - mean_dz - xh * mean_dzx);
            } else {
                o[k] = sc * dz;
            }
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) o[k] = tf32_rn(o[k], rt);
        if (VEC == 4) reinterpret_cast<float4*>(dy)[i] = make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]);
        else dy[i] = o[0];
    }
}

// ------------------------------------------------------------------ fixed-channel-group variants (C % 4 == 0, C/4 | 256)
// The grid stride is a multiple of 256 and C/4 divides 256, so a thread sees the SAME four channels in every iteration:
// the per-channel constants live in registers, there is no division in the loop, and four independent 128-bit loads per
// tensor are in flight per thread.
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// Traversal direction of the streaming kernels.  A kernel that re-reads what its predecessor just streamed (BatchNorm apply after
// the conv wrote y; BatchNorm-backward apply after the reduce pass read dout / y) finds the END of those tensors in the 126 MB L2,
// not the beginning: walking the 256-vector chunks backwards turns the first part of the pass into L2 hits.  Chunks are reversed
// whole, so a thread keeps its channel group (i % cg == threadIdx.x % cg).  MEASURED on the 2D iteration (CHAP_EW_REVERSE = 0 / 1 / 2 / 3,
// same box, back to back): 13.51 / 13.55 / 13.54 / 13.58 ms, BatchNorm families unchanged within noise -- the re-read tensors already hit
// L2 (50 MB level-0 tensors in a 126 MB cache).  Off by default; kept as an experiment knob.
__device__ __forceinline__ int64_t ew_index(int64_t i, int64_t nchunks, int reverse) {
    return reverse ? (((nchunks - 1 - (i >> 8)) << 8) | (i & 255)) : i;
}

// Elementwise dropout without a mask tensor (nn.Dropout(p) inside ConvBlock, code/networks/unet.py:53): the 1/(1-p)-scaled keep
// mask of vector i is recomputed wherever it is needed -- forward, backward reduce, backward apply -- from Philox4x32-10 keyed by
// (seed, device-side epoch) with counter (i, subsequence).  The epoch word is read from device memory (the trainer's iteration
// counter), so a replayed CUDA graph draws fresh masks every iteration; the subsequence separates the layers / passes of one iteration.
// Replaces bernoulli_ + div_ (two torch launches, three passes over a tensor of the activation's size) and the mask reads.
struct DropRng {
    uint32_t seed_lo, seed_hi, sub_lo, sub_hi;
    uint32_t thresh;            // keep iff random word < thresh  (thresh = keep * 2^32)
    float scale;                // 1 / keep
    const long long* epoch;     // nullable
    int on;
};
__device__ __forceinline__ uint2 drop_key(const DropRng& r) {
    const unsigned long long e = r.epoch ? (unsigned long long)*r.epoch : 0ull;
    return make_uint2(r.seed_lo ^ (uint32_t)(e * 0x9E3779B97F4A7C15ull >> 32), r.seed_hi ^ (uint32_t)e);
}
__device__ __forceinline__ float4 drop_mask4(const DropRng& r, uint2 key, int64_t i) {
    uint4 c = make_uint4((uint32_t)i, (uint32_t)((uint64_t)i >> 32), r.sub_lo, r.sub_hi);
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ key.x, lo1, hi0 ^ c.w ^ key.y, lo0);
        key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
    }
    return make_float4(c.x < r.thresh ? r.scale : 0.f, c.y < r.thresh ? r.scale : 0.f, c.z < r.thresh ? r.scale : 0.f, c.w < r.thresh ? r.scale : 0.f);
}

template <bool EL, bool RES>
__global__ void __launch_bounds__(256)
bn_act_fwd_fixed_kernel(const float* __restrict__ y, const float* __restrict__ ss, float slope,
                        const float* __restrict__ drop_nc, const float* __restrict__ drop_el,
                        const float* __restrict__ res, int64_t rows_per_sample, int c, int64_t total_vec, int reverse,
                        float* __restrict__ out, const DropRng rng) {
    pdl_enter();
    const int cg = c >> 2, g = threadIdx.x % cg;
    const uint2 key = (EL && rng.on) ? drop_key(rng) : make_uint2(0u, 0u);
    const float4 sc = ld4(ss + 4 * g), sh = ld4(ss + c + 4 * g);
    const int64_t per_sample = rows_per_sample * cg;
    const int64_t stride = (int64_t)gridDim.x * 256;
    const int64_t nchunks = (total_vec + 255) >> 8, span = nchunks << 8;
    const float4* y4 = reinterpret_cast<const float4*>(y);
    const float4* e4 = reinterpret_cast<const float4*>(drop_el);
    const float4* r4 = reinterpret_cast<const float4*>(res);
    float4* o4 = reinterpret_cast<float4*>(out);
    auto act = [&](float v, float s, float h) { float z = fmaf(v, s, h); return z > 0.f ? z : slope * z; };
    for (int64_t i0 = (int64_t)blockIdx.x * 256 + threadIdx.x; i0 < span; i0 += kEwUnroll * stride) {
        float4 v[kEwUnroll], e[kEwUnroll], r[kEwUnroll];
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const int64_t iv = i0 + u * stride;
            const int64_t i = ew_index(iv, nchunks, reverse);
            if (iv < span && i < total_vec) {
                v[u] = ldg_stream(y4 + i);
                if (EL) e[u] = rng.on ? drop_mask4(rng, key, i) : ldg_stream(e4 + i);
                if (RES) r[u] = ldg_stream(r4 + i);
            }
        }
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const int64_t iv = i0 + u * stride;
            if (iv >= span) break;
            const int64_t i = ew_index(iv, nchunks, reverse);
            if (i >= total_vec) continue;
            float4 o = make_float4(act(v[u].x, sc.x, sh.x), act(v[u].y, sc.y, sh.y), act(v[u].z, sc.z, sh.z), act(v[u].w, sc.w, sh.w));
            if (drop_nc) {
                const float4 f = ld4(drop_nc + (i / per_sample) * c + 4 * g);
                o.x *= f.x; o.y *= f.y; o.z *= f.z; o.w *= f.w;
            }
            if (EL) { o.x *= e[u].x; o.y *= e[u].y; o.z *= e[u].z; o.w *= e[u].w; }
            if (RES) { o.x += r[u].x; o.y += r[u].y; o.z += r[u].z; o.w += r[u].w; }
            o4[i] = o;
        }
    }
}

// block-level finish of a per-channel reduction: lanes of a warp that share a channel group are combined by shuffles,
// warps / row lanes through shared memory, one double atomic per channel and block
__device__ __forceinline__ void channel_reduce_finish(float4 s, float4 q, int c, int cg, int g, double* sums) {
    __shared__ float4 part[2][256];
    int lanes;                                    // partial sums per channel group left after the shuffle stage
    if (cg < 32) {
        for (int o = 16; o >= cg; o >>= 1) {
            s.x += __shfl_xor_sync(0xffffffffu, s.x, o); s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
            s.z += __shfl_xor_sync(0xffffffffu, s.z, o); s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
            q.x += __shfl_xor_sync(0xffffffffu, q.x, o); q.y += __shfl_xor_sync(0xffffffffu, q.y, o);
            q.z += __shfl_xor_sync(0xffffffffu, q.z, o); q.w += __shfl_xor_sync(0xffffffffu, q.w, o);
        }
        lanes = 8;
        if ((threadIdx.x & 31) < cg) { part[0][(threadIdx.x >> 5) * cg + g] = s; part[1][(threadIdx.x >> 5) * cg + g] = q; }
    } else {
        lanes = 256 / cg;
        part[0][threadIdx.x] = s; part[1][threadIdx.x] = q;        // thread t = lane * cg + g
    }
    __syncthreads();
    // 2 * c (channel, statistic) outputs; thread t sums output t over the lanes in double
    for (int t = threadIdx.x; t < 2 * c; t += 256) {
        const int which = t / c, ch = t - which * c;
        double a = 0.0;
        for (int l = 0; l < lanes; ++l) a += (double)reinterpret_cast<const float*>(&part[which][l * cg + (ch >> 2)])[ch & 3];
        atomicAdd(sums + t, a);
    }
}

__global__ void __launch_bounds__(256)
channel_stats_fixed_kernel(const float* __restrict__ y, int64_t total_vec, int c, double* sums) {
    pdl_enter();
    const int cg = c >> 2, g = threadIdx.x % cg;
    const int64_t stride = (int64_t)gridDim.x * 256;
    const float4* y4 = reinterpret_cast<const float4*>(y);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    for (int64_t i0 = (int64_t)blockIdx.x * 256 + threadIdx.x; i0 < total_vec; i0 += kEwUnroll * stride) {
        float4 v[kEwUnroll];
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const int64_t i = i0 + u * stride;
            v[u] = i < total_vec ? ldg_stream(y4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w;
            q.x = fmaf(v[u].x, v[u].x, q.x); q.y = fmaf(v[u].y, v[u].y, q.y); q.z = fmaf(v[u].z, v[u].z, q.z); q.w = fmaf(v[u].w, v[u].w, q.w);
        }
    }
    channel_reduce_finish(s, q, c, cg, g, sums);
}

template <bool EL>
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_fixed_kernel(const float* __restrict__ dout, const float* __restrict__ y, const float* __restrict__ ss,
                               const float* __restrict__ mi, float slope, const float* __restrict__ drop_nc,
                               const float* __restrict__ drop_el, int64_t rows_per_sample, int64_t total_vec, int c, double* sums,
                               const DropRng rng) {
    pdl_enter();
    const int cg = c >> 2, g = threadIdx.x % cg;
    const uint2 key = (EL && rng.on) ? drop_key(rng) : make_uint2(0u, 0u);
    const float4 sc = ld4(ss + 4 * g), sh = ld4(ss + c + 4 * g), mean = ld4(mi + 4 * g), istd = ld4(mi + c + 4 * g);
    const int64_t per_sample = rows_per_sample * cg;
    const int64_t stride = (int64_t)gridDim.x * 256;
    const float4* d4 = reinterpret_cast<const float4*>(dout);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    const float4* e4 = reinterpret_cast<const float4*>(drop_el);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    for (int64_t i0 = (int64_t)blockIdx.x * 256 + threadIdx.x; i0 < total_vec; i0 += kEwUnroll * stride) {
        float4 d[kEwUnroll], v[kEwUnroll], e[kEwUnroll];
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < total_vec) {
                d[u] = ldg_stream(d4 + i); v[u] = ldg_stream(y4 + i);
                if (EL) e[u] = rng.on ? drop_mask4(rng, key, i) : ldg_stream(e4 + i);
            }
        }
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const int64_t i = i0 + u * stride;
            if (i >= total_vec) break;
            float4 dd = d[u];
            if (drop_nc) { const float4 f = ld4(drop_nc + (i / per_sample) * c + 4 * g); dd.x *= f.x; dd.y *= f.y; dd.z *= f.z; dd.w *= f.w; }
            if (EL) { dd.x *= e[u].x; dd.y *= e[u].y; dd.z *= e[u].z; dd.w *= e[u].w; }
            const float z0 = bn_dz(dd.x, v[u].x, sc.x, sh.x, slope), z1 = bn_dz(dd.y, v[u].y, sc.y, sh.y, slope);
            const float z2 = bn_dz(dd.z, v[u].z, sc.z, sh.z, slope), z3 = bn_dz(dd.w, v[u].w, sc.w, sh.w, slope);
            s.x += z0; s.y += z1; s.z += z2; s.w += z3;
            q.x = fmaf(z0, (v[u].x - mean.x) * istd.x, q.x); q.y = fmaf(z1, (v[u].y - mean.y) * istd.y, q.y);
            q.z = fmaf(z2, (v[u].z - mean.z) * istd.z, q.z); q.w = fmaf(z3, (v[u].w - mean.w) * istd.w, q.w);
        }
    }
    channel_reduce_finish(s, q, c, cg, g, sums);
}

template <bool EL>
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_fixed_kernel(const float* __restrict__ dout, const float* __restrict__ y, const float* __restrict__ ss,
                              const float* __restrict__ mi, float slope, const float* __restrict__ drop_nc,
                              const float* __restrict__ drop_el, int64_t rows_per_sample, int c, int64_t total_vec, int train,
                              double inv_count, double* __restrict__ sums, float* __restrict__ dy,
                              float* __restrict__ dgamma, float* __restrict__ dbeta, const DropRng rng) {
    pdl_enter();
    const int cg = c >> 2, g = threadIdx.x % cg;
    const uint2 key = (EL && rng.on) ? drop_key(rng) : make_uint2(0u, 0u);
    const int train_bits = train;
    const int acc_pg = (train >> 1) & 1;       // bit 1 of `train`: ADD the parameter gradients into dgamma / dbeta (gradient-sink mode)
    const int persist = (train >> 2) & 1;      // bit 2: `sums` is a persistent buffer (2c sums + a ticket word) that must be handed back zeroed
    train &= 1;
    if (blockIdx.x == 0 && dgamma) {
        for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
            const float db = (float)sums[ch], dg = (float)sums[c + ch];
            dbeta[ch] = acc_pg ? dbeta[ch] + db : db;
            dgamma[ch] = acc_pg ? dgamma[ch] + dg : dg;
        }
    }
    const float4 sc = ld4(ss + 4 * g), sh = ld4(ss + c + 4 * g), mean = ld4(mi + 4 * g), istd = ld4(mi + c + 4 * g);
    float4 mz = make_float4(0.f, 0.f, 0.f, 0.f), mzx = mz;            // batch means of dz and dz * xhat (0 in eval mode)
    if (train) {
        mz = make_float4((float)(sums[4 * g] * inv_count), (float)(sums[4 * g + 1] * inv_count),
                         (float)(sums[4 * g + 2] * inv_count), (float)(sums[4 * g + 3] * inv_count));
        mzx = make_float4((float)(sums[c + 4 * g] * inv_count), (float)(sums[c + 4 * g + 1] * inv_count),
                          (float)(sums[c + 4 * g + 2] * inv_count), (float)(sums[c + 4 * g + 3] * inv_count));
    }
    if (persist) {
        // every block has read what it needs from `sums` above; the block that draws the last ticket zeroes the buffer (and the
        // ticket) for the next use of this BatchNorm layer -- no zero-fill launch per backward call
        __shared__ int s_last;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            s_last = atomicAdd(reinterpret_cast<unsigned*>(sums + 2 * c), 1u) == gridDim.x - 1 ? 1 : 0;
        }
        __syncthreads();
        if (s_last) {
            for (int ch = threadIdx.x; ch < 2 * c; ch += blockDim.x) sums[ch] = 0.0;
            if (threadIdx.x == 0) *reinterpret_cast<unsigned*>(sums + 2 * c) = 0u;
        }
    }
    const int64_t per_sample = rows_per_sample * cg;
    const int64_t stride = (int64_t)gridDim.x * 256;
    const float4* d4 = reinterpret_cast<const float4*>(dout);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    const float4* e4 = reinterpret_cast<const float4*>(drop_el);
    float4* o4 = reinterpret_cast<float4*>(dy);
    auto one = [&](float dd, float yv, float s_, float h_, float m_, float is_, float mz_, float mzx_) {
        const float dz = bn_dz(dd, yv, s_, h_, slope);
        return train ? s_ * (dz - mz_ - (yv - m_) * is_ * mzx_) : s_ * dz;
    };
    const int reverse = (train_bits >> 3) & 1;     // bit 3: walk the chunks backwards (the reduce pass just streamed dout / y forwards)
    const int64_t nchunks = (total_vec + 255) >> 8, span = nchunks << 8;
    for (int64_t i0 = (int64_t)blockIdx.x * 256 + threadIdx.x; i0 < span; i0 += kEwUnroll * stride) {
        float4 d[kEwUnroll], v[kEwUnroll], e[kEwUnroll];
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const int64_t iv = i0 + u * stride;
            const int64_t i = ew_index(iv, nchunks, reverse);
            if (iv < span && i < total_vec) {
                d[u] = ldg_stream(d4 + i); v[u] = ldg_stream(y4 + i);
                if (EL) e[u] = rng.on ? drop_mask4(rng, key, i) : ldg_stream(e4 + i);
            }
        }
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const int64_t iv = i0 + u * stride;
            if (iv >= span) break;
            const int64_t i = ew_index(iv, nchunks, reverse);
            if (i >= total_vec) continue;
            float4 dd = d[u];
            if (EL) { dd.x *= e[u].x; dd.y *= e[u].y; dd.z *= e[u].z; dd.w *= e[u].w; }
            if (drop_nc) { const float4 f = ld4(drop_nc + (i / per_sample) * c + 4 * g); dd.x *= f.x; dd.y *= f.y; dd.z *= f.z; dd.w *= f.w; }
            o4[i] = make_float4(one(dd.x, v[u].x, sc.x, sh.x, mean.x, istd.x, mz.x, mzx.x), one(dd.y, v[u].y, sc.y, sh.y, mean.y, istd.y, mz.y, mzx.y),
                                one(dd.z, v[u].z, sc.z, sh.z, mean.z, istd.z, mz.z, mzx.z), one(dd.w, v[u].w, sc.w, sh.w, mean.w, istd.w, mz.w, mzx.w));
        }
    }
}

// ------------------------------------------------------------------ max pool 2x2 (2D)
template <int VEC>
__global__ void __launch_bounds__(256)
maxpool2_fwd_kernel(const float* __restrict__ x, int h, int w, int c, int64_t total_vec, float* __restrict__ y) {
    pdl_enter();
    const int cg = c / VEC, oh = h / 2, ow = w / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
        int g = (int)(i % cg); int64_t r = i / cg;
        int xo = (int)(r % ow); r /= ow; int yo = (int)(r % oh); int64_t n = r / oh;
        const float* p = x + (((n * h + 2 * yo) * w + 2 * xo) * (int64_t)c) + g * VEC;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            float m = p[k];
            m = fmaxf(m, p[c + k]); m = fmaxf(m, p[(int64_t)w * c + k]); m = fmaxf(m, p[(int64_t)w * c + c + k]);
            y[i * VEC + k] = m;
        }
    }
}
template <int VEC>
__global__ void __launch_bounds__(256)
maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int h, int w, int c,
                    int64_t total_vec, float* __restrict__ dx) {
    pdl_enter();
    const int cg = c / VEC, oh = h / 2, ow = w / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
        int g = (int)(i % cg); int64_t r = i / cg;
        int xo = (int)(r % ow); r /= ow; int yo = (int)(r % oh); int64_t n = r / oh;
        const int64_t base = (((n * h + 2 * yo) * w + 2 * xo) * (int64_t)c) + g * VEC;
        const int64_t off[4] = {0, (int64_t)c, (int64_t)w * c, (int64_t)w * c + c};
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            int best = 0; float m = x[base + k];
#pragma unroll
            for (int j = 1; j < 4; ++j) { float v = x[base + off[j] + k]; if (v > m) { m = v; best = j; } }
            float gy = dy[i * VEC + k];
#pragma unroll
            for (int j = 0; j < 4; ++j) dx[base + off[j] + k] = (j == best) ? gy : 0.f;
        }
    }
}

// Vector forms (c % 4 == 0, 16-byte aligned, < 2^31 vectors): 128-bit loads / stores and 32-bit index arithmetic.  The generic kernels above issue
// four 32-bit accesses per tap and a 64-bit div / mod chain per vector -- more instructions than a streaming kernel of this size can hide.
__global__ void __launch_bounds__(256)
maxpool2_fwd_v4_kernel(const float* __restrict__ x, int h, int w, int cg, uint32_t total_vec, float* __restrict__ y) {
    pdl_enter();
    const uint32_t oh = (uint32_t)h / 2, ow = (uint32_t)w / 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += gridDim.x * blockDim.x) {
        const uint32_t g = i % (uint32_t)cg; uint32_t r = i / (uint32_t)cg;
        const uint32_t xo = r % ow; r /= ow; const uint32_t yo = r % oh; const uint32_t n = r / oh;
        const float4* p = x4 + (((int64_t)n * h + 2 * yo) * w + 2 * xo) * cg + g;
        const float4 a = ldg_stream(p), b = ldg_stream(p + cg), c_ = ldg_stream(p + (int64_t)w * cg), d = ldg_stream(p + (int64_t)w * cg + cg);
        y4[i] = make_float4(fmaxf(fmaxf(fmaxf(a.x, b.x), c_.x), d.x), fmaxf(fmaxf(fmaxf(a.y, b.y), c_.y), d.y),
                            fmaxf(fmaxf(fmaxf(a.z, b.z), c_.z), d.z), fmaxf(fmaxf(fmaxf(a.w, b.w), c_.w), d.w));
    }
}
__global__ void __launch_bounds__(256)
maxpool2_bwd_v4_kernel(const float* __restrict__ x, const float* __restrict__ dy, int h, int w, int cg, uint32_t total_vec, float* __restrict__ dx) {
    pdl_enter();
    const uint32_t oh = (uint32_t)h / 2, ow = (uint32_t)w / 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* g4 = reinterpret_cast<const float4*>(dy);
    float4* o4 = reinterpret_cast<float4*>(dx);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += gridDim.x * blockDim.x) {
        const uint32_t g = i % (uint32_t)cg; uint32_t r = i / (uint32_t)cg;
        const uint32_t xo = r % ow; r /= ow; const uint32_t yo = r % oh; const uint32_t n = r / oh;
        const int64_t base = (((int64_t)n * h + 2 * yo) * w + 2 * xo) * cg + g;
        const int64_t off[4] = {0, (int64_t)cg, (int64_t)w * cg, (int64_t)w * cg + cg};
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = ldg_stream(x4 + base + off[j]);
        const float4 gy = ldg_stream(g4 + i);
        // first maximum wins (strict >), like the scalar kernel and torch's max_pool2d backward
        auto pick = [](float a, float b, float c_, float d) { int best = 0; float m = a; if (b > m) { m = b; best = 1; } if (c_ > m) { m = c_; best = 2; } if (d > m) best = 3; return best; };
        const int bx = pick(v[0].x, v[1].x, v[2].x, v[3].x), by = pick(v[0].y, v[1].y, v[2].y, v[3].y);
        const int bz = pick(v[0].z, v[1].z, v[2].z, v[3].z), bw = pick(v[0].w, v[1].w, v[2].w, v[3].w);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            o4[base + off[j]] = make_float4(bx == j ? gy.x : 0.f, by == j ? gy.y : 0.f, bz == j ? gy.z : 0.f, bw == j ? gy.w : 0.f);
    }
}

// ------------------------------------------------------------------ x2 upsample, align_corners=True
__device__ __forceinline__ void lerp_src(int o, int in, int out, int& i0, int& i1, float& l1) {
    float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    float src = scale * (float)o;
    i0 = (int)src;                        // src >= 0
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = src - (float)i0;
}

template <int VEC>
__global__ void __launch_bounds__(256)
upsample2x_fwd_kernel(const float* __restrict__ x, int d, int h, int w, int c, int nd, int64_t total_vec, int rt, float* __restrict__ y) {
    pdl_enter();
    const int cg = c / VEC, od = nd == 3 ? 2 * d : 1, oh = 2 * h, ow = 2 * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
        int g = (int)(i % cg); int64_t r = i / cg;
        int xo = (int)(r % ow); r /= ow; int yo = (int)(r % oh); r /= oh; int zo = (int)(r % od); int64_t n = r / od;
        int z0 = 0, z1 = 0, y0, y1, x0, x1; float lz = 0.f, ly, lx;
        if (nd == 3) lerp_src(zo, d, od, z0, z1, lz);
        lerp_src(yo, h, oh, y0, y1, ly);
        lerp_src(xo, w, ow, x0, x1, lx);
        const float* b = x + n * (int64_t)d * h * w * c + g * VEC;
        auto at = [&](int zz, int yy, int xx, int k) { return b[(((int64_t)zz * h + yy) * w + xx) * c + k]; };
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            float a00 = at(z0, y0, x0, k) * (1.f - lx) + at(z0, y0, x1, k) * lx;
            float a01 = at(z0, y1, x0, k) * (1.f - lx) + at(z0, y1, x1, k) * lx;
            float v0 = a00 * (1.f - ly) + a01 * ly;
            if (nd == 3) {
                float a10 = at(z1, y0, x0, k) * (1.f - lx) + at(z1, y0, x1, k) * lx;
                float a11 = at(z1, y1, x0, k) * (1.f - lx) + at(z1, y1, x1, k) * lx;
                float v1 = a10 * (1.f - ly) + a11 * ly;
                v0 = v0 * (1.f - lz) + v1 * lz;
            }
            y[i * VEC + k] = tf32_rn(v0, rt);
        }
    }
}

// Vector form of the forward (c % 4 == 0): one output float4 per thread from 4 (2D) / 8 (3D) 128-bit loads -- the scalar kernel above
// issued four 32-bit loads per tap and ran at 17-27 % of the HBM rate (3D: 80 us for a 128 MB output).  out_stride4 = float4 per
// output pixel row (>= c / 4): the result can be written straight into the second half of a channel-concat buffer.
// A block of 256 threads = (c / 4 vectors) x (a COMPACT tile of 256 / (c / 4) pixels, 8 x 8 in 2D, 4 x 4 x 4 in 3D for 16 channels): a
// row-major assignment gave every block one 64-pixel row segment, whose 2 (x2) source rows nobody else on the SM re-used (L2 -> L1
// traffic 2x the OUTPUT size; 3D 16 ch: 107 us for 145 MB).  lx / ly / lz = log2 of the tile edges.
struct UpTile { int lx, ly, lz, ntx, nty, ntz; int64_t tiles; };
__device__ __forceinline__ bool up_tile_pixel(const UpTile& t, uint32_t tile, int p, int W, int H, int D, int& xo, int& yo, int& zo, int64_t& n) {
    uint32_t r = tile;
    const int bx = (int)(r % (uint32_t)t.ntx); r /= (uint32_t)t.ntx;
    const int by = (int)(r % (uint32_t)t.nty); r /= (uint32_t)t.nty;
    const int bz = (int)(r % (uint32_t)t.ntz); n = r / (uint32_t)t.ntz;
    xo = (bx << t.lx) + (p & ((1 << t.lx) - 1));
    yo = (by << t.ly) + ((p >> t.lx) & ((1 << t.ly) - 1));
    zo = (bz << t.lz) + (p >> (t.lx + t.ly));
    return xo < W && yo < H && zo < D;
}
__global__ void __launch_bounds__(256)
upsample2x_fwd_v4_kernel(const float* __restrict__ x, int d, int h, int w, int cg, int nd, const UpTile tl, int out_stride4, float* __restrict__ y) {
    pdl_enter();
    const int od = nd == 3 ? 2 * d : 1, oh = 2 * h, ow = 2 * w;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    const int g = threadIdx.x % cg, p = threadIdx.x / cg;
    for (uint32_t tile = blockIdx.x; tile < (uint32_t)tl.tiles; tile += gridDim.x) {
        int xo, yo, zo; int64_t n;
        if (!up_tile_pixel(tl, tile, p, ow, oh, od, xo, yo, zo, n)) continue;
        int z0 = 0, z1 = 0, y0, y1, x0, x1; float lz = 0.f, ly, lx;
        if (nd == 3) lerp_src(zo, d, od, z0, z1, lz);
        lerp_src(yo, h, oh, y0, y1, ly);
        lerp_src(xo, w, ow, x0, x1, lx);
        const float4* b = x4 + n * (int64_t)d * h * w * cg + g;
        auto at = [&](int zz, int yy, int xx) { return __ldg(b + (((int64_t)zz * h + yy) * w + xx) * cg); };
        auto mix = [](const float4& p_, const float4& q, float l) {      // p * (1 - l) + q * l, the scalar kernel's order of operations
            return make_float4(p_.x * (1.f - l) + q.x * l, p_.y * (1.f - l) + q.y * l, p_.z * (1.f - l) + q.z * l, p_.w * (1.f - l) + q.w * l);
        };
        float4 v0 = mix(mix(at(z0, y0, x0), at(z0, y0, x1), lx), mix(at(z0, y1, x0), at(z0, y1, x1), lx), ly);
        if (nd == 3) {
            const float4 v1 = mix(mix(at(z1, y0, x0), at(z1, y0, x1), lx), mix(at(z1, y1, x0), at(z1, y1, x1), lx), ly);
            v0 = mix(v0, v1, lz);
        }
        const int64_t row = ((n * od + zo) * oh + yo) * (int64_t)ow + xo;
        y4[row * out_stride4 + g] = v0;
    }
}

// contributions of the outputs along one dim to input index i: list of (o, weight)
__device__ __forceinline__ int contrib(int i, int in, int out, int* oi, float* wt) {
    int cnt = 0;
    int lo = 2 * i - 4 < 0 ? 0 : 2 * i - 4, hi = 2 * i + 5 > out - 1 ? out - 1 : 2 * i + 5;
    if (in == 1) { lo = 0; hi = out - 1; }
    for (int o = lo; o <= hi && cnt < 8; ++o) {
        int i0, i1; float l1;
        lerp_src(o, in, out, i0, i1, l1);
        float wgt = 0.f;
        if (i0 == i) wgt += 1.f - l1;
        if (i1 == i) wgt += l1;
        if (i0 == i || i1 == i) { oi[cnt] = o; wt[cnt] = wgt; ++cnt; }
    }
    return cnt;
}

template <int VEC>
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const float* __restrict__ dy, int d, int h, int w, int c, int nd, int64_t total_vec, int rt, float* __restrict__ dx) {
    pdl_enter();
    const int cg = c / VEC, od = nd == 3 ? 2 * d : 1, oh = 2 * h, ow = 2 * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
        int g = (int)(i % cg); int64_t r = i / cg;
        int xi = (int)(r % w); r /= w; int yi = (int)(r % h); r /= h; int zi = (int)(r % d); int64_t n = r / d;
        int oz[8], oy[8], ox[8]; float wz[8], wy[8], wx[8];
        int nz = 1; oz[0] = 0; wz[0] = 1.f;
        if (nd == 3) nz = contrib(zi, d, od, oz, wz);
        int ny = contrib(yi, h, oh, oy, wy);
        int nx = contrib(xi, w, ow, ox, wx);
        float acc[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
        const float* b = dy + n * (int64_t)od * oh * ow * c + g * VEC;
        for (int a = 0; a < nz; ++a)
            for (int bb = 0; bb < ny; ++bb) {
                float wzy = wz[a] * wy[bb];
                const float* row = b + (((int64_t)oz[a] * oh + oy[bb]) * ow) * c;
                for (int cc = 0; cc < nx; ++cc) {
                    float wt = wzy * wx[cc];
                    const float* p = row + (int64_t)ox[cc] * c;
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[k] = fmaf(wt, p[k], acc[k]);
                }
            }
#pragma unroll
        for (int k = 0; k < VEC; ++k) dx[i * VEC + k] = tf32_rn(acc[k], rt);
    }
}

// Vector form of the backward (c % 4 == 0).  Input position i gathers the ~4 outputs per dimension that interpolate from it (16 taps in
// 2D, ~66 in 3D).  The per-dimension (output index, weight) lists are built ONCE per block into shared memory (the scalar kernel
// re-derived them per thread with a 10-step search per dimension) and every tap is one 128-bit load; blocks are persistent.
constexpr int kUpMaxDim = 256, kUpTaps = 6;
__global__ void __launch_bounds__(256)
upsample2x_bwd_v4_kernel(const float* __restrict__ dy, int d, int h, int w, int cg, int nd, const UpTile tl, int in_stride4, float* __restrict__ dx) {
    pdl_enter();
    // tables sized by the actual extents (dynamic shared memory: (d + h + w) * (kUpTaps * 6 + 4) bytes) -- a static worst-case table
    // (28 KB per block) took the L1 away from the dy taps, which are re-read ~8x
    extern __shared__ __align__(16) unsigned char up_smem[];
    const int ext = d + h + w;
    float* s_w = reinterpret_cast<float*>(up_smem);                          // [ext][kUpTaps]
    short* s_o = reinterpret_cast<short*>(s_w + (size_t)ext * kUpTaps);      // [ext][kUpTaps]
    int* s_n = reinterpret_cast<int*>(s_o + (size_t)ext * kUpTaps);          // [ext]   (ext * kUpTaps is even: 4-byte aligned)
    const int od = nd == 3 ? 2 * d : 1, oh = 2 * h, ow = 2 * w;
    const int dims[3] = {d, h, w}, odims[3] = {od, oh, ow};
    for (int t = threadIdx.x; t < ext; t += blockDim.x) {
        const int which = t < d ? 0 : (t < d + h ? 1 : 2);
        const int i = which == 0 ? t : (which == 1 ? t - d : t - d - h);
        int oi[8]; float wt[8];
        int cnt = 1; oi[0] = 0; wt[0] = 1.f;
        if (!(which == 0 && nd == 2)) cnt = contrib(i, dims[which], odims[which], oi, wt);
        if (cnt > kUpTaps) cnt = kUpTaps;                      // x2 with align_corners: at most 5
        s_n[t] = cnt;
        for (int k = 0; k < cnt; ++k) { s_o[t * kUpTaps + k] = (short)oi[k]; s_w[t * kUpTaps + k] = wt[k]; }
    }
    __syncthreads();
    const float4* dy4 = reinterpret_cast<const float4*>(dy);
    float4* dx4 = reinterpret_cast<float4*>(dx);
    const int g = threadIdx.x % cg, p = threadIdx.x / cg;
    for (uint32_t tile = blockIdx.x; tile < (uint32_t)tl.tiles; tile += gridDim.x) {
        int xi, yi, zi; int64_t n;
        if (!up_tile_pixel(tl, tile, p, w, h, d, xi, yi, zi, n)) continue;
        const int64_t i = (((n * d + zi) * h + yi) * (int64_t)w + xi) * cg + g;
        const int ez = zi, ey = d + yi, ex = d + h + xi;                    // table rows of this position
        const int nz = s_n[ez], ny = s_n[ey], nx = s_n[ex];
        float wx[kUpTaps]; int ox[kUpTaps];
#pragma unroll
        for (int cc = 0; cc < kUpTaps; ++cc) { wx[cc] = s_w[ex * kUpTaps + cc]; ox[cc] = s_o[ex * kUpTaps + cc] * in_stride4; }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* b = dy4 + n * (int64_t)od * oh * ow * in_stride4 + g;
        for (int a = 0; a < nz; ++a) {
            const float wza = s_w[ez * kUpTaps + a];
            const int64_t zoff = (int64_t)s_o[ez * kUpTaps + a] * oh;
            for (int bb = 0; bb < ny; ++bb) {
                const float wzy = wza * s_w[ey * kUpTaps + bb];
                const float4* rowp = b + ((zoff + s_o[ey * kUpTaps + bb]) * ow) * in_stride4;
#pragma unroll
                for (int cc = 0; cc < kUpTaps; ++cc) {
                    if (cc < nx) {
                        const float wt = wzy * wx[cc];
                        const float4 p = __ldg(rowp + ox[cc]);
                        acc.x = fmaf(wt, p.x, acc.x); acc.y = fmaf(wt, p.y, acc.y); acc.z = fmaf(wt, p.z, acc.z); acc.w = fmaf(wt, p.w, acc.w);
                    }
                }
            }
        }
        dx4[i] = acc;
    }
}


// ------------------------------------------------------------------ concat / split / misc
__global__ void __launch_bounds__(256)
concat_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t rows, int ca, int cb, float* __restrict__ out) {
    pdl_enter();
    const int ct = ca + cb;
    const int64_t total = rows * ct;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / ct; int ch = (int)(i % ct);
        out[i] = ch < ca ? a[r * ca + ch] : b[r * cb + ch - ca];
    }
}
template <typename I>         // I = uint32_t when rows * (ca + cb) / 4 < 2^31 (one 32-bit division per vector instead of a 64-bit div + mod)
__global__ void __launch_bounds__(256)
concat4_kernel(const float4* __restrict__ a, const float4* __restrict__ b, int64_t rows, int ca4, int cb4, float4* __restrict__ out) {
    pdl_enter();
    const I ct = (I)(ca4 + cb4);
    const I total = (I)(rows * (ca4 + cb4));
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
        const I r = i / ct; const int ch = (int)(i - r * ct);
        out[i] = ch < ca4 ? ldg_stream(a + (int64_t)r * ca4 + ch) : ldg_stream(b + (int64_t)r * cb4 + ch - ca4);
    }
}
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ in, int64_t rows, int ca, int cb, float* __restrict__ a, float* __restrict__ b) {
    pdl_enter();
    const int ct = ca + cb;
    const int64_t total = rows * ct;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / ct; int ch = (int)(i % ct);
        float v = in[i];
        if (ch < ca) { if (a) a[r * ca + ch] = v; } else { if (b) b[r * cb + ch - ca] = v; }
    }
}
__global__ void __launch_bounds__(256)
channel_scale_kernel(const float* __restrict__ x, const float* __restrict__ s, int64_t rows_per_sample, int c, int64_t total, float* __restrict__ out) {
    pdl_enter();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int ch = (int)(i % c); int64_t n = i / (rows_per_sample * c);
        out[i] = x[i] * s[n * c + ch];
    }
}
// perform_dropout (code/networks/FilterDropout.py:45-89) for one pyramid level in ONE pass: feat [N, rps, C] is read once and
// both decoder inputs out1 / out2 [N + nu, rps, C] = cat(feat, feat[nl:] * m{1,2}[nu, C]) are written (nl = N - nu labelled rows).
// The reference makes two masked copies and two concatenations (4 passes, 8N moved); this moves 4N.  m == nullptr: factor 1.
template <int VEC>
__global__ void __launch_bounds__(256)
feature_dropout_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ m1, const float* __restrict__ m2,
                           int n, int nu, int64_t rps, int c, float* __restrict__ out1, float* __restrict__ out2) {
    pdl_enter();
    const int cv = c / VEC;
    const int64_t per_sample = rps * cv, total = (int64_t)n * per_sample;
    const int nl = n - nu;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = i / per_sample;
        const int ch = (int)(i % cv) * VEC;
        float v[VEC];
        if (VEC == 4) { const float4 t = ldg_stream(reinterpret_cast<const float4*>(feat) + i); v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w; }
        else v[0] = feat[i];
        if (VEC == 4) {
            reinterpret_cast<float4*>(out1)[i] = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
            reinterpret_cast<float4*>(out2)[i] = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
        } else { out1[i] = v[0]; out2[i] = v[0]; }
        if (s >= nl) {
            const int64_t u = s - nl;
            const int64_t o = i + (int64_t)nu * per_sample;                 // row N + u of the outputs
            float a[VEC], b[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                a[k] = m1 ? v[k] * m1[u * c + ch + k] : v[k];
                b[k] = m2 ? v[k] * m2[u * c + ch + k] : v[k];
            }
            if (VEC == 4) {
                reinterpret_cast<float4*>(out1)[o] = make_float4(a[0], a[1 % VEC], a[2 % VEC], a[3 % VEC]);
                reinterpret_cast<float4*>(out2)[o] = make_float4(b[0], b[1 % VEC], b[2 % VEC], b[3 % VEC]);
            } else { out1[o] = a[0]; out2[o] = b[0]; }
        }
    }
}
// backward: dfeat[s] = d1[s] + d2[s] (+ d1[N + u] * m1[u] + d2[N + u] * m2[u] for the unlabelled rows s = nl + u); d1 / d2 nullable
template <int VEC>
__global__ void __launch_bounds__(256)
feature_dropout_bwd_kernel(const float* __restrict__ d1, const float* __restrict__ d2, const float* __restrict__ m1, const float* __restrict__ m2,
                           int n, int nu, int64_t rps, int c, float* __restrict__ dfeat) {
    pdl_enter();
    const int cv = c / VEC;
    const int64_t per_sample = rps * cv, total = (int64_t)n * per_sample;
    const int nl = n - nu;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = i / per_sample;
        const int ch = (int)(i % cv) * VEC;
        float acc[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
        auto add = [&](const float* src, int64_t idx, const float* m, int64_t u) {
            float v[VEC];
            if (VEC == 4) { const float4 t = ldg_stream(reinterpret_cast<const float4*>(src) + idx); v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w; }
            else v[0] = src[idx];
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[k] += m ? v[k] * m[u * c + ch + k] : v[k];
        };
        if (d1) add(d1, i, nullptr, 0);
        if (d2) add(d2, i, nullptr, 0);
        if (s >= nl) {
            const int64_t u = s - nl, o = i + (int64_t)nu * per_sample;
            if (d1) add(d1, o, m1, u);
            if (d2) add(d2, o, m2, u);
        }
        if (VEC == 4) reinterpret_cast<float4*>(dfeat)[i] = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
        else dfeat[i] = acc[0];
    }
}

__global__ void __launch_bounds__(256)
axpy_kernel(const float* __restrict__ a, const float* __restrict__ b, float alpha, int64_t total, float* __restrict__ out) {
    pdl_enter();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = fmaf(alpha, b[i], a[i]);
}
__global__ void __launch_bounds__(256)
mask_mix_kernel(const float* __restrict__ a, const float* __restrict__ b, const int64_t* __restrict__ m,
                int64_t rows_per_sample, int c, int64_t total, float* __restrict__ out) {
    pdl_enter();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = (i / c) % rows_per_sample;
        float mm = (float)m[r];
        out[i] = a[i] * mm + b[i] * (1.f - mm);
    }
}

}  // namespace chap

using namespace chap;

extern "C" int chap_channel_stats(const float* y, int64_t rows, int32_t c, double* sums, void* stream) {
    return channel_stats(y, rows, c, sums, S(stream));
}

extern "C" int chap_bn_finalize(const double* sums, int32_t slots, int64_t count, const float* gamma, const float* beta, float eps,
                                float momentum, float* running_mean, float* running_var, int64_t* nbt,
                                float* mean_invstd, float* scale_shift, int32_t c, void* stream) {
    CHAP_REQUIRE(sums && gamma && beta && mean_invstd && scale_shift && c > 0 && count > 0 && slots >= 1, CHAP_ERR_BAD_ARG, "bn_finalize: bad argument");
    CHAP_REQUIRE((running_mean == nullptr) == (running_var == nullptr), CHAP_ERR_BAD_ARG, "bn_finalize: running stats must both be set or both NULL");
    launch_k(bn_finalize_kernel, (c + 127) / 128, 128, 0, S(stream), sums, slots, count, gamma, beta, eps, momentum, running_mean,
                                                               running_var, running_mean ? nbt : nullptr, mean_invstd, scale_shift, c);
    return launched("bn_finalize_kernel");
}

extern "C" int chap_bn_eval_params(const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                                   float* mean_invstd, float* scale_shift, int32_t c, void* stream) {
    CHAP_REQUIRE(gamma && beta && rm && rv && mean_invstd && scale_shift && c > 0, CHAP_ERR_BAD_ARG, "bn_eval_params: bad argument");
    launch_k(bn_eval_params_kernel, (c + 127) / 128, 128, 0, S(stream), gamma, beta, rm, rv, eps, mean_invstd, scale_shift, c);
    return launched("bn_eval_params_kernel");
}

static bool all16(std::initializer_list<const void*> ps) {
    for (const void* p : ps) if (p && !aligned16(p)) return false;
    return true;
}

// host side of DropRng: the C ABI struct -> kernel argument (p in (0, 1), validated by the callers)
static DropRng make_rng(const chap_dropout_rng* r) {
    DropRng d{};
    if (!r) return d;
    const double keep = 1.0 - (double)r->p;
    const double t = keep * 4294967296.0;
    d.seed_lo = (uint32_t)r->seed; d.seed_hi = (uint32_t)(r->seed >> 32);
    d.sub_lo = (uint32_t)r->subsequence; d.sub_hi = (uint32_t)(r->subsequence >> 32);
    d.thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
    d.scale = (float)(1.0 / keep);
    d.epoch = reinterpret_cast<const long long*>(r->epoch_dev);
    d.on = 1;
    return d;
}
static bool rng_ok(const chap_dropout_rng* r) { return !r || (r->p > 0.f && r->p < 1.f); }

static int bn_act_fwd_impl(const float* y, const float* ss, float slope, const float* drop_nc, const float* drop_el, const chap_dropout_rng* rng_in,
                           const float* residual, int32_t n, int64_t rps, int32_t c, float* out, void* stream) {
    CHAP_REQUIRE(y && ss && out && n > 0 && rps > 0 && c > 0, CHAP_ERR_BAD_ARG, "bn_act_fwd: bad argument");
    CHAP_REQUIRE(rng_ok(rng_in) && !(rng_in && drop_el), CHAP_ERR_BAD_ARG, "bn_act_fwd: dropout needs 0 < p < 1 and either a mask or a generator");
    const DropRng rng = make_rng(rng_in);
    const int64_t total = (int64_t)n * rps * c;
    KernelTimer timer("bn_act_fwd", 0.0, 4.0 * total * (2 + (drop_el ? 1 : 0) + (residual ? 1 : 0)), S(stream));
    if (c % 4 == 0 && 256 % (c / 4) == 0 && all16({y, drop_el, residual, out, ss, drop_nc}) && round_tf32_on() == 0) {
        const int grid = grid_for(total / 4, 256 * kEwUnroll);
        static const int rev = getenv("CHAP_EW_REVERSE") ? atoi(getenv("CHAP_EW_REVERSE")) : 0;
#define CHAP_FWD_FIXED(EL, RES) launch_k(bn_act_fwd_fixed_kernel<EL, RES>, grid, 256, 0, S(stream), y, ss, slope, drop_nc, drop_el, residual, rps, c, total / 4, rev & 1, out, rng)
        if (drop_el || rng.on) { if (residual) CHAP_FWD_FIXED(true, true); else CHAP_FWD_FIXED(true, false); }
        else         { if (residual) CHAP_FWD_FIXED(false, true); else CHAP_FWD_FIXED(false, false); }
#undef CHAP_FWD_FIXED
    } else if (rng.on) {
        return fail(CHAP_ERR_BAD_ARG, "bn_act_fwd: generated dropout needs c %% 4 == 0, 256 %% (c / 4) == 0 and 16-byte aligned buffers (c = %d)", c);
    } else if (c % 4 == 0 && all16({y, drop_el, residual, out})) {
        launch_k(bn_act_fwd_kernel<4>, grid_for(total / 4, 256 * 4), 256, 0, S(stream), y, ss, slope, drop_nc, drop_el, residual, rps, c, total / 4, round_tf32_on(), out);
    } else {
        launch_k(bn_act_fwd_kernel<1>, grid_for(total, 256 * 4), 256, 0, S(stream), y, ss, slope, drop_nc, drop_el, residual, rps, c, total, round_tf32_on(), out);
    }
    return launched("bn_act_fwd_kernel");
}

extern "C" int chap_bn_act_fwd(const float* y, const float* ss, float slope, const float* drop_nc, const float* drop_el,
                               const float* residual, int32_t n, int64_t rps, int32_t c, float* out, void* stream) {
    return bn_act_fwd_impl(y, ss, slope, drop_nc, drop_el, nullptr, residual, n, rps, c, out, stream);
}
extern "C" int chap_bn_act_fwd_rng(const float* y, const float* ss, float slope, const float* drop_nc, const chap_dropout_rng* rng,
                                   const float* residual, int32_t n, int64_t rps, int32_t c, float* out, void* stream) {
    CHAP_REQUIRE(rng != nullptr, CHAP_ERR_BAD_ARG, "bn_act_fwd_rng: the generator description is required");
    return bn_act_fwd_impl(y, ss, slope, drop_nc, nullptr, rng, residual, n, rps, c, out, stream);
}

static int bn_act_bwd_impl(const float* dout, const float* y, const float* ss, const float* mi, float slope, const float* drop_nc,
                           const float* drop_el, int32_t n, int64_t rps, int32_t c, int32_t train, double* sums, float* dy, float* dgamma,
                           float* dbeta, void* stream, const chap_dropout_rng* rng_in = nullptr);

extern "C" int chap_bn_act_bwd_rng(const float* dout, const float* y, const float* ss, const float* mi, float slope, const float* drop_nc,
                                   const chap_dropout_rng* rng, int32_t n, int64_t rps, int32_t c, int32_t train, double* sums,
                                   int32_t sums_persistent, float* dy, float* dgamma, float* dbeta, int32_t accumulate, void* stream) {
    CHAP_REQUIRE(rng != nullptr, CHAP_ERR_BAD_ARG, "bn_act_bwd_rng: the generator description is required");
    return bn_act_bwd_impl(dout, y, ss, mi, slope, drop_nc, nullptr, n, rps, c, (train ? 1 : 0) | (accumulate ? 2 : 0) | (sums_persistent ? 4 : 0),
                           sums, dy, dgamma, dbeta, stream, rng);
}

extern "C" int chap_bn_act_bwd(const float* dout, const float* y, const float* ss, const float* mi, const float* gamma,
                               float slope, const float* drop_nc, const float* drop_el, int32_t n, int64_t rps, int32_t c,
                               int32_t train, double* sums, float* dy, float* dgamma, float* dbeta, void* stream) {
    (void)gamma;
    return bn_act_bwd_impl(dout, y, ss, mi, slope, drop_nc, drop_el, n, rps, c, train ? 1 : 0, sums, dy, dgamma, dbeta, stream);
}

extern "C" int chap_bn_act_bwd_acc(const float* dout, const float* y, const float* ss, const float* mi, float slope, const float* drop_nc,
                                   const float* drop_el, int32_t n, int64_t rps, int32_t c, int32_t train, double* sums,
                                   int32_t sums_persistent, float* dy, float* dgamma_acc, float* dbeta_acc, void* stream) {
    return bn_act_bwd_impl(dout, y, ss, mi, slope, drop_nc, drop_el, n, rps, c, (train ? 1 : 0) | 2 | (sums_persistent ? 4 : 0), sums, dy,
                           dgamma_acc, dbeta_acc, stream);
}

static int bn_act_bwd_impl(const float* dout, const float* y, const float* ss, const float* mi, float slope, const float* drop_nc,
                           const float* drop_el, int32_t n, int64_t rps, int32_t c, int32_t train, double* sums, float* dy, float* dgamma,
                           float* dbeta, void* stream, const chap_dropout_rng* rng_in) {
    CHAP_REQUIRE(dout && y && ss && mi && sums && dy && n > 0 && rps > 0 && c > 0, CHAP_ERR_BAD_ARG, "bn_act_bwd: bad argument");
    CHAP_REQUIRE(rng_ok(rng_in) && !(rng_in && drop_el), CHAP_ERR_BAD_ARG, "bn_act_bwd: dropout needs 0 < p < 1 and either a mask or a generator");
    const DropRng rng = make_rng(rng_in);
    CHAP_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), CHAP_ERR_BAD_ARG, "bn_act_bwd: dgamma/dbeta must both be set or both NULL");
    const int64_t rows = (int64_t)n * rps, total = rows * c;
    cudaStream_t st = S(stream);
    const bool v4 = c % 4 == 0 && all16({dout, y, drop_el, dy});
    const int cg = v4 ? c / 4 : c;
    CHAP_REQUIRE(cg <= 256, CHAP_ERR_BAD_ARG, "bn_act_bwd: too many channels (%d)", c);
    const bool fixed = v4 && 256 % cg == 0 && all16({ss, mi, drop_nc}) && round_tf32_on() == 0;
    CHAP_REQUIRE(fixed || !rng.on, CHAP_ERR_BAD_ARG, "bn_act_bwd: generated dropout needs c %% 4 == 0, 256 %% (c / 4) == 0 and 16-byte aligned buffers (c = %d)", c);
    // persistent sums (bit 2): zero on entry, re-zeroed by the last block of the apply kernel -- only the fixed-group kernels do that;
    // on the generic path the buffer is zeroed before AND after by launches (it must come back zeroed either way)
    const bool persist = (train & 4) != 0;
    if (!persist) CHAP_TRY(zero_async(sums, (size_t)2 * c * sizeof(double), st));
    if (!fixed) train &= ~4;
    KernelTimer timer("bn_act_bwd", 0.0, 4.0 * total * (3 + (drop_el ? 1 : 0)), st);   // algorithmic: read dout, y; write dy
    const double inv_count = 1.0 / (double)rows;
    if (fixed) {
        if ((train & 1) || dgamma) {
            const int rgrid = grid_for(total / 4, 256 * kEwUnroll * 2, kNumSMs * 8);
            if (drop_el || rng.on) launch_k(bn_act_bwd_reduce_fixed_kernel<true>, rgrid, 256, 0, st, dout, y, ss, mi, slope, drop_nc, drop_el, rps, total / 4, c, sums, rng);
            else launch_k(bn_act_bwd_reduce_fixed_kernel<false>, rgrid, 256, 0, st, dout, y, ss, mi, slope, drop_nc, drop_el, rps, total / 4, c, sums, rng);
            CHAP_TRY(launched("bn_act_bwd_reduce_fixed_kernel"));
        }
        const int agrid = grid_for(total / 4, 256 * kEwUnroll);
        static const int rev = getenv("CHAP_EW_REVERSE") ? atoi(getenv("CHAP_EW_REVERSE")) : 0;
        const int tb = train | ((rev & 2) ? 8 : 0);
        if (drop_el || rng.on) launch_k(bn_act_bwd_apply_fixed_kernel<true>, agrid, 256, 0, st, dout, y, ss, mi, slope, drop_nc, drop_el, rps, c, total / 4, tb, inv_count, sums, dy, dgamma, dbeta, rng);
        else launch_k(bn_act_bwd_apply_fixed_kernel<false>, agrid, 256, 0, st, dout, y, ss, mi, slope, drop_nc, drop_el, rps, c, total / 4, tb, inv_count, sums, dy, dgamma, dbeta, rng);
        return launched("bn_act_bwd_apply_fixed_kernel");
    }
    const int rpb = 256 / cg;
    int grid = grid_for(rows, rpb * 8, kNumSMs * 8);
    if (v4) launch_k(bn_act_bwd_reduce_kernel<4>, grid, 256, 0, st, dout, y, ss, mi, slope, drop_nc, drop_el, rps, rows, c, sums);
    else launch_k(bn_act_bwd_reduce_kernel<1>, grid, 256, 0, st, dout, y, ss, mi, slope, drop_nc, drop_el, rps, rows, c, sums);
    CHAP_TRY(launched("bn_act_bwd_reduce_kernel"));
    if (v4) launch_k(bn_act_bwd_apply_kernel<4>, grid_for(total / 4, 256 * 4), 256, 0, st, dout, y, ss, mi, slope, drop_nc, drop_el, rps, c, total / 4, train, round_tf32_on(), inv_count, sums, dy, dgamma, dbeta);
    else launch_k(bn_act_bwd_apply_kernel<1>, grid_for(total, 256 * 4), 256, 0, st, dout, y, ss, mi, slope, drop_nc, drop_el, rps, c, total, train, round_tf32_on(), inv_count, sums, dy, dgamma, dbeta);
    CHAP_TRY(launched("bn_act_bwd_apply_kernel"));
    if (persist) CHAP_TRY(zero_async(sums, (size_t)2 * c * sizeof(double), st));      // hand the persistent buffer back zeroed
    return CHAP_OK;
}

extern "C" int chap_maxpool2_fwd(const float* x, int32_t n, int32_t h, int32_t w, int32_t c, float* y, void* stream) {
    KernelTimer timer_("maxpool2_fwd", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(x && y && n > 0 && h > 0 && w > 0 && c > 0 && h % 2 == 0 && w % 2 == 0, CHAP_ERR_BAD_ARG, "maxpool2_fwd: bad argument (h, w must be even)");
    const int64_t total = (int64_t)n * (h / 2) * (w / 2) * c;
    if (c % 4 == 0 && aligned16(x) && aligned16(y) && total / 4 < 0x7FFFFFFFll)
        launch_k(maxpool2_fwd_v4_kernel, grid_for(total / 4, 256 * 2), 256, 0, S(stream), x, h, w, c / 4, (uint32_t)(total / 4), y);
    else if (c % 4 == 0) launch_k(maxpool2_fwd_kernel<4>, grid_for(total / 4, 256 * 2), 256, 0, S(stream), x, h, w, c, total / 4, y);
    else launch_k(maxpool2_fwd_kernel<1>, grid_for(total, 256 * 2), 256, 0, S(stream), x, h, w, c, total, y);
    return launched("maxpool2_fwd_kernel");
}

extern "C" int chap_maxpool2_bwd(const float* x, const float* dy, int32_t n, int32_t h, int32_t w, int32_t c, float* dx, void* stream) {
    KernelTimer timer_("maxpool2_bwd", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(x && dy && dx && n > 0 && h > 0 && w > 0 && c > 0 && h % 2 == 0 && w % 2 == 0, CHAP_ERR_BAD_ARG, "maxpool2_bwd: bad argument (h, w must be even)");
    const int64_t total = (int64_t)n * (h / 2) * (w / 2) * c;
    if (c % 4 == 0 && aligned16(x) && aligned16(dy) && aligned16(dx) && total / 4 < 0x7FFFFFFFll)
        launch_k(maxpool2_bwd_v4_kernel, grid_for(total / 4, 256 * 2), 256, 0, S(stream), x, dy, h, w, c / 4, (uint32_t)(total / 4), dx);
    else if (c % 4 == 0) launch_k(maxpool2_bwd_kernel<4>, grid_for(total / 4, 256 * 2), 256, 0, S(stream), x, dy, h, w, c, total / 4, dx);
    else launch_k(maxpool2_bwd_kernel<1>, grid_for(total, 256 * 2), 256, 0, S(stream), x, dy, h, w, c, total, dx);
    return launched("maxpool2_bwd_kernel");
}

// compact pixel tile of 256 / cg pixels for the vector upsampling kernels: the edge bits go to x, y (, z) in turn
static UpTile up_tile(int n, int D, int H, int W, int cg, int nd) {
    UpTile t{};
    int bits = 0;
    for (int px = 256 / cg; px > 1; px >>= 1) ++bits;
    for (int k = 0; bits > 0; ++k, --bits) {
        const int dim = k % nd;                       // 0: x, 1: y, 2: z
        if (dim == 0) ++t.lx; else if (dim == 1) ++t.ly; else ++t.lz;
    }
    t.ntx = (W + (1 << t.lx) - 1) >> t.lx; t.nty = (H + (1 << t.ly) - 1) >> t.ly; t.ntz = (D + (1 << t.lz) - 1) >> t.lz;
    t.tiles = (int64_t)n * t.ntx * t.nty * t.ntz;
    return t;
}

extern "C" int chap_upsample2x_fwd(const float* x, int32_t nd, int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, float* y, void* stream) {
    KernelTimer timer_("upsample2x_fwd", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(x && y && (nd == 2 || nd == 3) && n > 0 && d > 0 && h > 0 && w > 0 && c > 0 && (nd == 3 || d == 1), CHAP_ERR_BAD_ARG, "upsample2x_fwd: bad argument");
    const int64_t total = (int64_t)n * (nd == 3 ? 2 * d : 1) * 2 * h * 2 * w * c;
    if (c % 4 == 0 && 256 % (c / 4) == 0 && all16({x, y}) && round_tf32_on() == 0 && total / 4 < 0x7FFFFFFFll) {
        const UpTile tl = up_tile(n, nd == 3 ? 2 * d : 1, 2 * h, 2 * w, c / 4, nd);
        launch_k(upsample2x_fwd_v4_kernel, (int)(tl.tiles < kNumSMs * 16 ? tl.tiles : kNumSMs * 16), 256, 0, S(stream), x, d, h, w, c / 4, nd, tl, c / 4, y);
    }
    else if (c % 4 == 0) launch_k(upsample2x_fwd_kernel<4>, grid_for(total / 4, 256 * 2), 256, 0, S(stream), x, d, h, w, c, nd, total / 4, round_tf32_on(), y);
    else launch_k(upsample2x_fwd_kernel<1>, grid_for(total, 256 * 2), 256, 0, S(stream), x, d, h, w, c, nd, total, round_tf32_on(), y);
    return launched("upsample2x_fwd_kernel");
}

extern "C" int chap_upsample2x_bwd(const float* dy, int32_t nd, int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, float* dx, void* stream) {
    KernelTimer timer_("upsample2x_bwd", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(dy && dx && (nd == 2 || nd == 3) && n > 0 && d > 0 && h > 0 && w > 0 && c > 0 && (nd == 3 || d == 1), CHAP_ERR_BAD_ARG, "upsample2x_bwd: bad argument");
    const int64_t total = (int64_t)n * d * h * w * c;
    if (c % 4 == 0 && 256 % (c / 4) == 0 && all16({dy, dx}) && round_tf32_on() == 0 && d <= kUpMaxDim && h <= kUpMaxDim && w <= kUpMaxDim && total * 2 < 0x7FFFFFFFll) {
        const UpTile tl = up_tile(n, d, h, w, c / 4, nd);
        const size_t tab = (size_t)(d + h + w) * (kUpTaps * 6 + 4);
        launch_k(upsample2x_bwd_v4_kernel, (int)(tl.tiles < kNumSMs * 8 ? tl.tiles : kNumSMs * 8), 256, tab, S(stream), dy, d, h, w, c / 4, nd, tl, c / 4, dx);
    }
    else if (c % 4 == 0) launch_k(upsample2x_bwd_kernel<4>, grid_for(total / 4, 256), 256, 0, S(stream), dy, d, h, w, c, nd, total / 4, round_tf32_on(), dx);
    else launch_k(upsample2x_bwd_kernel<1>, grid_for(total, 256), 256, 0, S(stream), dy, d, h, w, c, nd, total, round_tf32_on(), dx);
    return launched("upsample2x_bwd_kernel");
}

extern "C" int chap_concat_channels(const float* a, const float* b, int64_t rows, int32_t ca, int32_t cb, float* out, void* stream) {
    KernelTimer timer_("concat_channels", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(a && b && out && rows > 0 && ca > 0 && cb > 0, CHAP_ERR_BAD_ARG, "concat_channels: bad argument");
    const int64_t total = rows * (ca + cb);
    if (ca % 4 == 0 && cb % 4 == 0 && all16({a, b, out}))
        if (total / 4 < 0x7FFFFFFFll) launch_k(concat4_kernel<uint32_t>, grid_for(total / 4, 256 * 4), 256, 0, S(stream), (const float4*)a, (const float4*)b, rows, ca / 4, cb / 4, (float4*)out);
        else launch_k(concat4_kernel<int64_t>, grid_for(total / 4, 256 * 4), 256, 0, S(stream), (const float4*)a, (const float4*)b, rows, ca / 4, cb / 4, (float4*)out);
    else
        launch_k(concat_kernel, grid_for(total, 256 * 4), 256, 0, S(stream), a, b, rows, ca, cb, out);
    return launched("concat_kernel");
}

extern "C" int chap_split_channels(const float* in, int64_t rows, int32_t ca, int32_t cb, float* a, float* b, void* stream) {
    KernelTimer timer_("split_channels", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(in && rows > 0 && ca > 0 && cb > 0 && (a || b), CHAP_ERR_BAD_ARG, "split_channels: bad argument");
    launch_k(split_kernel, grid_for(rows * (ca + cb), 256 * 4), 256, 0, S(stream), in, rows, ca, cb, a, b);
    return launched("split_kernel");
}

extern "C" int chap_channel_scale(const float* x, const float* s, int32_t n, int64_t rps, int32_t c, float* out, void* stream) {
    KernelTimer timer_("channel_scale", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(x && s && out && n > 0 && rps > 0 && c > 0, CHAP_ERR_BAD_ARG, "channel_scale: bad argument");
    const int64_t total = (int64_t)n * rps * c;
    launch_k(channel_scale_kernel, grid_for(total, 256 * 4), 256, 0, S(stream), x, s, rps, c, total, out);
    return launched("channel_scale_kernel");
}

extern "C" int chap_feature_dropout_fwd(const float* feat, const float* m1, const float* m2, int32_t n, int32_t nu, int64_t rps, int32_t c,
                                        float* out1, float* out2, void* stream) {
    CHAP_REQUIRE(feat && out1 && out2 && n > 0 && nu >= 0 && nu <= n && rps > 0 && c > 0, CHAP_ERR_BAD_ARG, "feature_dropout_fwd: bad argument");
    const double elems = (double)n * rps * c, extra = (double)nu * rps * c;
    KernelTimer timer_("feature_dropout_fwd", 0.0, 4.0 * (elems + 2.0 * (elems + extra)), S(stream));       // read feat once, write both outputs
    const bool v4 = c % 4 == 0 && aligned16(feat) && aligned16(out1) && aligned16(out2);
    const int64_t total = (int64_t)n * rps * c;
    if (v4) launch_k(feature_dropout_fwd_kernel<4>, grid_for(total / 4, 256 * 4), 256, 0, S(stream), feat, m1, m2, n, nu, rps, c, out1, out2);
    else launch_k(feature_dropout_fwd_kernel<1>, grid_for(total, 256 * 4), 256, 0, S(stream), feat, m1, m2, n, nu, rps, c, out1, out2);
    return launched("feature_dropout_fwd_kernel");
}

extern "C" int chap_feature_dropout_bwd(const float* d1, const float* d2, const float* m1, const float* m2, int32_t n, int32_t nu, int64_t rps,
                                        int32_t c, float* dfeat, void* stream) {
    CHAP_REQUIRE((d1 || d2) && dfeat && n > 0 && nu >= 0 && nu <= n && rps > 0 && c > 0, CHAP_ERR_BAD_ARG, "feature_dropout_bwd: bad argument");
    const double elems = (double)n * rps * c, extra = (double)nu * rps * c;
    KernelTimer timer_("feature_dropout_bwd", 0.0, 4.0 * (elems + ((d1 ? 1.0 : 0.0) + (d2 ? 1.0 : 0.0)) * (elems + extra)), S(stream));
    const bool v4 = c % 4 == 0 && aligned16(dfeat) && (!d1 || aligned16(d1)) && (!d2 || aligned16(d2));
    const int64_t total = (int64_t)n * rps * c;
    if (v4) launch_k(feature_dropout_bwd_kernel<4>, grid_for(total / 4, 256 * 4), 256, 0, S(stream), d1, d2, m1, m2, n, nu, rps, c, dfeat);
    else launch_k(feature_dropout_bwd_kernel<1>, grid_for(total, 256 * 4), 256, 0, S(stream), d1, d2, m1, m2, n, nu, rps, c, dfeat);
    return launched("feature_dropout_bwd_kernel");
}

extern "C" int chap_axpy(const float* a, const float* b, float alpha, int64_t elems, float* out, void* stream) {
    KernelTimer timer_("axpy", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(a && b && out && elems > 0, CHAP_ERR_BAD_ARG, "axpy: bad argument");
    launch_k(axpy_kernel, grid_for(elems, 256 * 4), 256, 0, S(stream), a, b, alpha, elems, out);
    return launched("axpy_kernel");
}

extern "C" int chap_mask_mix(const float* a, const float* b, const int64_t* mask, int32_t n, int64_t rps, int32_t c, float* out, void* stream) {
    KernelTimer timer_("mask_mix", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(a && b && mask && out && n > 0 && rps > 0 && c > 0, CHAP_ERR_BAD_ARG, "mask_mix: bad argument");
    const int64_t total = (int64_t)n * rps * c;
    launch_k(mask_mix_kernel, grid_for(total, 256 * 4), 256, 0, S(stream), a, b, mask, rps, c, total, out);
    return launched("mask_mix_kernel");
}
