// Softmax-family loss kernels on channels-last logits [rows, C] (C <= 8): one pass per tensor,
// the C logits of a position sit in adjacent words (128-bit load for C = 4), partial sums are
// reduced warp-shuffle -> shared memory -> one double atomic per block and quantity.
#include "common.cuh"

namespace chap {


template <int C>
__device__ __forceinline__ void load_row(const float* __restrict__ p, int64_t r, float* x) {
    if (C == 4) {
        float4 v = __ldg(reinterpret_cast<const float4*>(p) + r);
        x[0] = v.x; x[1 % C] = v.y; x[2 % C] = v.z; x[3 % C] = v.w;
    } else if (C == 2) {
        float2 v = __ldg(reinterpret_cast<const float2*>(p) + r);
        x[0] = v.x; x[1 % C] = v.y;
    } else {
#pragma unroll
        for (int k = 0; k < C; ++k) x[k] = __ldg(p + r * C + k);
    }
}
template <int C>
__device__ __forceinline__ void store_row(float* __restrict__ p, int64_t r, const float* x) {
    if (C == 4) reinterpret_cast<float4*>(p)[r] = make_float4(x[0], x[1 % C], x[2 % C], x[3 % C]);
    else if (C == 2) reinterpret_cast<float2*>(p)[r] = make_float2(x[0], x[1 % C]);
    else {
#pragma unroll
        for (int k = 0; k < C; ++k) p[r * C + k] = x[k];
    }
}

// softmax of x[C] -> p[C]; returns log-sum-exp
template <int C>
__device__ __forceinline__ float softmax_row(const float* x, float* p) {
    float m = x[0];
#pragma unroll
    for (int k = 1; k < C; ++k) m = fmaxf(m, x[k]);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k) { p[k] = expf(x[k] - m); s += p[k]; }
    float inv = 1.f / s;
#pragma unroll
    for (int k = 0; k < C; ++k) p[k] *= inv;
    return m + logf(s);
}
template <int C>
__device__ __forceinline__ int argmax_row(const float* p) {
    int best = 0; float m = p[0];
#pragma unroll
    for (int k = 1; k < C; ++k) if (p[k] > m) { m = p[k]; best = k; }
    return best;
}

// block-level reduction of NV per-thread partials into double atomics
template <int NV>
__device__ __forceinline__ void block_reduce_to(float* v, double* sums) {
    __shared__ float red[8][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        float t = warp_sum(v[k]);
        if (lane == 0) red[warp][k] = t;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double a = 0.0;
        const int nw = blockDim.x >> 5;
        for (int w = 0; w < nw; ++w) a += (double)red[w][threadIdx.x];
        atomicAdd(sums + threadIdx.x, a);
    }
}

// ------------------------------------------------------------------ pseudo-label block
template <int C>
__global__ void __launch_bounds__(256)
pseudo_label_kernel(const float* __restrict__ pre1, const float* __restrict__ pre2, int64_t rows,
                    float* soft1, float* soft2, int64_t* arg1, int64_t* arg2, float* knowledge) {
    pdl_enter();
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        float x1[C], x2[C], p1[C], p2[C];
        load_row<C>(pre1, r, x1); load_row<C>(pre2, r, x2);
        float l1 = softmax_row<C>(x1, p1), l2 = softmax_row<C>(x2, p2);
        int a1 = argmax_row<C>(p1), a2 = argmax_row<C>(p2);
        if (soft1) store_row<C>(soft1, r, p1);
        if (soft2) store_row<C>(soft2, r, p2);
        if (arg1) arg1[r] = a1;
        if (arg2) arg2[r] = a2;
        if (knowledge) {
            float v1 = x1[0], v2 = x2[0];
#pragma unroll
            for (int k = 1; k < C; ++k) { if (k == a2) v1 = x1[k]; if (k == a1) v2 = x2[k]; }
            knowledge[r] = (l1 - v1) + (l2 - v2);
        }
    }
}

template <int C>
__global__ void __launch_bounds__(256)
softmax_kernel(const float* __restrict__ logits, int64_t rows, float* __restrict__ out) {
    pdl_enter();
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        float x[C], p[C];
        load_row<C>(logits, r, x);
        softmax_row<C>(x, p);
        store_row<C>(out, r, p);
    }
}

template <int C>
__global__ void __launch_bounds__(256)
argmax_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t rows, int64_t* __restrict__ out) {
    pdl_enter();
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        float x[C], p[C];
        load_row<C>(a, r, x);
        if (b) {
            float y[C];
            load_row<C>(b, r, y);
#pragma unroll
            for (int k = 0; k < C; ++k) x[k] = (x[k] + y[k]) / 2.0f;
        }
        softmax_row<C>(x, p);
        out[r] = argmax_row<C>(p);
    }
}

// ------------------------------------------------------------------ masked Dice + CE vs hard labels
__device__ __forceinline__ int read_label(const void* labels, int dtype, int64_t r) {
    return dtype == CHAP_LABEL_I64 ? (int)reinterpret_cast<const int64_t*>(labels)[r]
                                   : (int)reinterpret_cast<const float*>(labels)[r];
}

template <int C>
__global__ void __launch_bounds__(256)
dice_ce_fwd_kernel(const float* __restrict__ logits, const void* __restrict__ labels, int dtype,
                   const int64_t* __restrict__ mask, int invert, int64_t rps, int64_t rows, double* sums) {
    pdl_enter();
    float v[3 * C + 2];
#pragma unroll
    for (int k = 0; k < 3 * C + 2; ++k) v[k] = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        float x[C], p[C];
        load_row<C>(logits, r, x);
        float lse = softmax_row<C>(x, p);
        int t = read_label(labels, dtype, r);
        float m = (float)mask[r % rps];
        if (invert) m = 1.f - m;
        float xt = 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            float hit = (k == t) ? 1.f : 0.f;
            v[k] += p[k] * hit * m;
            v[C + k] += p[k] * p[k] * m;
            v[2 * C + k] += hit * m;
            if (k == t) xt = x[k];
        }
        v[3 * C] += (lse - xt) * m;
        v[3 * C + 1] += m;
    }
    block_reduce_to<3 * C + 2>(v, sums);
}

template <int C>
__global__ void __launch_bounds__(256)
dice_ce_bwd_kernel(const float* __restrict__ logits, const void* __restrict__ labels, int dtype,
                   const int64_t* __restrict__ mask, int invert, int64_t rps, int64_t rows,
                   const float* __restrict__ coef, int accumulate, float* __restrict__ dlogits) {
    pdl_enter();
    float ci[C], cs[C];
#pragma unroll
    for (int k = 0; k < C; ++k) { ci[k] = coef[k]; cs[k] = coef[C + k]; }
    const float cce = coef[2 * C];
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        float x[C], p[C], g[C];
        load_row<C>(logits, r, x);
        softmax_row<C>(x, p);
        int t = read_label(labels, dtype, r);
        float m = (float)mask[r % rps];
        if (invert) m = 1.f - m;
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            float hit = (k == t) ? 1.f : 0.f;
            g[k] = (ci[k] * hit + cs[k] * 2.f * p[k]) * m;
            dot += g[k] * p[k];
        }
        float o[C];
#pragma unroll
        for (int k = 0; k < C; ++k) {
            float hit = (k == t) ? 1.f : 0.f;
            o[k] = p[k] * (g[k] - dot) + cce * m * (p[k] - hit);
        }
        if (accumulate) {
            float old[C];
            load_row<C>(dlogits, r, old);
#pragma unroll
            for (int k = 0; k < C; ++k) o[k] += old[k];
        }
        store_row<C>(dlogits, r, o);
    }
}

// ------------------------------------------------------------------ consistency distance vs soft target
template <int C>
__global__ void __launch_bounds__(256)
consistency_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                       const float* __restrict__ mask, int dist, int64_t rows, double* sums) {
    pdl_enter();
    float v[3 * C + 1];
#pragma unroll
    for (int k = 0; k < 3 * C + 1; ++k) v[k] = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        float x[C], p[C], t[C];
        load_row<C>(logits, r, x);
        load_row<C>(target, r, t);
        float lse = softmax_row<C>(x, p);
        float m = mask ? mask[r] : 1.f;
        if (dist == CHAP_DIST_KL) {
            float kl = 0.f;
#pragma unroll
            for (int k = 0; k < C; ++k) {
                float tl = t[k] > 0.f ? t[k] * logf(t[k]) : 0.f;       // xlogy(t, t)
                kl += tl - t[k] * (x[k] - lse);
            }
            v[0] += kl * m;
        } else {
#pragma unroll
            for (int k = 0; k < C; ++k) {
                v[k] += p[k] * t[k] * m;
                v[C + k] += p[k] * p[k] * m;
                v[2 * C + k] += t[k] * t[k] * m;
            }
        }
    }
    block_reduce_to<3 * C + 1>(v, sums);
}

template <int C>
__global__ void __launch_bounds__(256)
consistency_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                       const float* __restrict__ mask, int dist, int64_t rows, const float* __restrict__ coef,
                       float* __restrict__ dlogits) {
    pdl_enter();
    float c0[C], c1[C];
#pragma unroll
    for (int k = 0; k < C; ++k) { c0[k] = coef[k]; c1[k] = dist == CHAP_DIST_KL ? 0.f : coef[C + k]; }
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        float x[C], p[C], t[C], o[C];
        load_row<C>(logits, r, x);
        load_row<C>(target, r, t);
        softmax_row<C>(x, p);
        float m = mask ? mask[r] : 1.f;
        if (dist == CHAP_DIST_KL) {
            float ts = 0.f;
#pragma unroll
            for (int k = 0; k < C; ++k) ts += t[k];
#pragma unroll
            for (int k = 0; k < C; ++k) o[k] = c0[0] * m * (p[k] * ts - t[k]);
        } else {
            float g[C], dot = 0.f;
#pragma unroll
            for (int k = 0; k < C; ++k) { g[k] = (c0[k] * t[k] + c1[k] * 2.f * p[k]) * m; dot += g[k] * p[k]; }
#pragma unroll
            for (int k = 0; k < C; ++k) o[k] = p[k] * (g[k] - dot);
        }
        store_row<C>(dlogits, r, o);
    }
}

// ------------------------------------------------------------------ create_maskV1 pieces
__global__ void __launch_bounds__(256)
patch_score_kernel(const float* __restrict__ know, const int64_t* __restrict__ a1, const int64_t* __restrict__ a2,
                   int nd, int d, int h, int w, int s, int64_t total, float* __restrict__ score) {
    pdl_enter();
    const int pd = nd == 3 ? d / s : 1, ph = h / s, pw = w / s;
    const int sd = nd == 3 ? s : 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        int px = (int)(r % pw); r /= pw; int py = (int)(r % ph); r /= ph; int pz = (int)(r % pd); int64_t n = r / pd;
        float ks = 0.f, ds = 0.f;
        for (int z = 0; z < sd; ++z)
            for (int y = 0; y < s; ++y)
                for (int x = 0; x < s; ++x) {
                    int64_t e = ((n * d + pz * sd + z) * h + py * s + y) * (int64_t)w + px * s + x;
                    ks += know[e];
                    ds += (a1[e] != a2[e]) ? 1.f : 0.f;
                }
        float cnt = (float)(sd * s * s);
        score[i] = ks / cnt + ds / cnt;
    }
}
__global__ void __launch_bounds__(256)
patch_mask_kernel(const float* __restrict__ score, const float* __restrict__ kth, int nd, int d, int h, int w, int s,
                  int64_t total, float* __restrict__ mask) {
    pdl_enter();
    const int pd = nd == 3 ? d / s : 1, ph = h / s, pw = w / s;
    const int sd = nd == 3 ? s : 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        int x = (int)(r % w); r /= w; int y = (int)(r % h); r /= h; int z = (int)(r % d); int64_t n = r / d;
        int64_t p = ((n * pd + z / sd) * ph + y / s) * pw + x / s;
        mask[i] = score[p] >= kth[n] ? 1.f : 0.f;
    }
}

}  // namespace chap

using namespace chap;

#define DISPATCH_C(c, CALL)                                                             \
    switch (c) {                                                                        \
        case 2: { constexpr int C = 2; CALL; break; }                                   \
        case 3: { constexpr int C = 3; CALL; break; }                                   \
        case 4: { constexpr int C = 4; CALL; break; }                                   \
        case 8: { constexpr int C = 8; CALL; break; }                                   \
        default: return fail(CHAP_ERR_BAD_ARG, "unsupported class count %d (2, 3, 4, 8)", (int)(c)); \
    }

static bool row_aligned(const void* p, int c) { return c == 4 ? aligned16(p) : (c == 2 ? ((uintptr_t)p & 7u) == 0 : true); }

extern "C" int chap_pseudo_label(const float* pre1, const float* pre2, int64_t rows, int32_t c, float* soft1, float* soft2,
                                 int64_t* arg1, int64_t* arg2, float* knowledge, void* stream) {
    KernelTimer timer_("pseudo_label", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(pre1 && pre2 && rows > 0, CHAP_ERR_BAD_ARG, "pseudo_label: bad argument");
    CHAP_REQUIRE(row_aligned(pre1, c) && row_aligned(pre2, c) && row_aligned(soft1, c) && row_aligned(soft2, c), CHAP_ERR_ALIGNMENT, "pseudo_label: misaligned");
    int grid = grid_for(rows, 256 * 2);
    DISPATCH_C(c, (launch_k(pseudo_label_kernel<C>, grid, 256, 0, S(stream), pre1, pre2, rows, soft1, soft2, arg1, arg2, knowledge)));
    return launched("pseudo_label_kernel");
}

extern "C" int chap_softmax(const float* logits, int64_t rows, int32_t c, float* out, void* stream) {
    KernelTimer timer_("softmax", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(logits && out && rows > 0, CHAP_ERR_BAD_ARG, "softmax: bad argument");
    CHAP_REQUIRE(row_aligned(logits, c) && row_aligned(out, c), CHAP_ERR_ALIGNMENT, "softmax: misaligned");
    int grid = grid_for(rows, 256 * 2);
    DISPATCH_C(c, (launch_k(softmax_kernel<C>, grid, 256, 0, S(stream), logits, rows, out)));
    return launched("softmax_kernel");
}

extern "C" int chap_argmax(const float* a, const float* b, int64_t rows, int32_t c, int64_t* out, void* stream) {
    KernelTimer timer_("argmax", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(a && out && rows > 0, CHAP_ERR_BAD_ARG, "argmax: bad argument");
    CHAP_REQUIRE(row_aligned(a, c) && row_aligned(b, c), CHAP_ERR_ALIGNMENT, "argmax: misaligned");
    int grid = grid_for(rows, 256 * 2);
    DISPATCH_C(c, (launch_k(argmax_kernel<C>, grid, 256, 0, S(stream), a, b, rows, out)));
    return launched("argmax_kernel");
}

extern "C" int chap_dice_ce_fwd(const float* logits, const void* labels, int32_t dtype, const int64_t* mask, int32_t invert,
                                int32_t n, int64_t rps, int32_t c, double* sums, void* stream) {
    KernelTimer timer_("dice_ce_fwd", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(logits && labels && mask && sums && n > 0 && rps > 0, CHAP_ERR_BAD_ARG, "dice_ce_fwd: bad argument");
    CHAP_REQUIRE(row_aligned(logits, c), CHAP_ERR_ALIGNMENT, "dice_ce_fwd: misaligned logits");
    CHAP_TRY(zero_async(sums, (size_t)(3 * c + 2) * sizeof(double), S(stream)));
    const int64_t rows = (int64_t)n * rps;
    int grid = grid_for(rows, 256 * 4, kNumSMs * 4);
    DISPATCH_C(c, (launch_k(dice_ce_fwd_kernel<C>, grid, 256, 0, S(stream), logits, labels, dtype, mask, invert, rps, rows, sums)));
    return launched("dice_ce_fwd_kernel");
}

extern "C" int chap_dice_ce_bwd(const float* logits, const void* labels, int32_t dtype, const int64_t* mask, int32_t invert,
                                int32_t n, int64_t rps, int32_t c, const float* coef, int32_t accumulate, float* dlogits, void* stream) {
    KernelTimer timer_("dice_ce_bwd", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(logits && labels && mask && coef && dlogits && n > 0 && rps > 0, CHAP_ERR_BAD_ARG, "dice_ce_bwd: bad argument");
    CHAP_REQUIRE(row_aligned(logits, c) && row_aligned(dlogits, c), CHAP_ERR_ALIGNMENT, "dice_ce_bwd: misaligned");
    const int64_t rows = (int64_t)n * rps;
    int grid = grid_for(rows, 256 * 2);
    DISPATCH_C(c, (launch_k(dice_ce_bwd_kernel<C>, grid, 256, 0, S(stream), logits, labels, dtype, mask, invert, rps, rows, coef, accumulate, dlogits)));
    return launched("dice_ce_bwd_kernel");
}

// ------------------------------------------------------------------ mix_loss scalar tail (code/train_ours_2D.py:198-216)
// One thread block turns the two sets of masked sums into (loss_image, loss_patch, total); a second tiny kernel turns the
// upstream gradient of those three scalars into the coefficient vectors of chap_dice_ce_bwd.  Replaces ~85 scalar torch
// kernels per mix_loss call (each a ~2 us graph node).
namespace chap {
__device__ __forceinline__ void dice_ce_from_sums(const double* s, int c, double& dice, double& ce) {
    double acc = 0.0;
    for (int k = 0; k < c; ++k) acc += 1.0 - (2.0 * s[k] + 1e-10) / (s[c + k] + s[2 * c + k] + 1e-10);
    dice = acc / c;
    ce = s[3 * c] / (s[3 * c + 1] + 1e-16);
}
__global__ void mix_loss_finalize_kernel(const double* s_img, const double* s_patch, int c, float w_img, float w_patch, float* out3) {
    pdl_enter();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double d1, c1, d2, c2;
    dice_ce_from_sums(s_img, c, d1, c1);
    dice_ce_from_sums(s_patch, c, d2, c2);
    d1 *= w_img; c1 *= w_img; d2 *= w_patch; c2 *= w_patch;
    out3[0] = (float)((d1 + c1) / 2.0);
    out3[1] = (float)((d2 + c2) / 2.0);
    out3[2] = (float)(((d1 + d2) + (c1 + c2)) / 2.0);
}
// coef = (d loss / d inter[c], d loss / d (sum s^2 m)[c], d loss / d (sum CE m)) for one pass, given dL/d(dice) = dL/d(ce) = up
__device__ __forceinline__ void dice_ce_coef(const double* s, int c, double up, float* coef, int k) {
    if (k < c) {
        const double den = s[c + k] + s[2 * c + k] + 1e-10;
        coef[k] = (float)(up * (-2.0 / (c * den)));
        coef[c + k] = (float)(up * ((2.0 * s[k] + 1e-10) / (c * den * den)));
    }
    if (k == 0) coef[2 * c] = (float)(up / (s[3 * c + 1] + 1e-16));
}
__global__ void mix_loss_coef_kernel(const double* s_img, const double* s_patch, int c, float w_img, float w_patch,
                                     const float* g3, float* coef_img, float* coef_patch) {
    pdl_enter();
    const int k = threadIdx.x;
    // loss_image = w_img (dice1 + ce1) / 2, loss_patch = w_patch (dice2 + ce2) / 2, total = their sum
    dice_ce_coef(s_img, c, 0.5 * w_img * ((double)g3[0] + (double)g3[2]), coef_img, k);
    dice_ce_coef(s_patch, c, 0.5 * w_patch * ((double)g3[1] + (double)g3[2]), coef_patch, k);
}
}  // namespace chap

extern "C" int chap_mix_loss_finalize(const double* sums_img, const double* sums_patch, int32_t c, float w_img, float w_patch,
                                      float* out3, void* stream) {
    CHAP_REQUIRE(sums_img && sums_patch && out3 && c > 0 && c <= 64, CHAP_ERR_BAD_ARG, "mix_loss_finalize: bad argument");
    launch_k(mix_loss_finalize_kernel, 1, 32, 0, S(stream), sums_img, sums_patch, c, w_img, w_patch, out3);
    return launched("mix_loss_finalize_kernel");
}

extern "C" int chap_mix_loss_coef(const double* sums_img, const double* sums_patch, int32_t c, float w_img, float w_patch,
                                  const float* grad_out3, float* coef_img, float* coef_patch, void* stream) {
    CHAP_REQUIRE(sums_img && sums_patch && grad_out3 && coef_img && coef_patch && c > 0 && c <= 64, CHAP_ERR_BAD_ARG,
                 "mix_loss_coef: bad argument");
    launch_k(mix_loss_coef_kernel, 1, 64, 0, S(stream), sums_img, sums_patch, c, w_img, w_patch, grad_out3, coef_img, coef_patch);
    return launched("mix_loss_coef_kernel");
}

extern "C" int chap_consistency_fwd(const float* logits, const float* target, const float* mask, int32_t dist, int64_t rows,
                                    int32_t c, double* sums, void* stream) {
    KernelTimer timer_("consistency_fwd", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(logits && target && sums && rows > 0 && (dist == CHAP_DIST_KL || dist == CHAP_DIST_DICE), CHAP_ERR_BAD_ARG, "consistency_fwd: bad argument");
    CHAP_REQUIRE(row_aligned(logits, c) && row_aligned(target, c), CHAP_ERR_ALIGNMENT, "consistency_fwd: misaligned");
    CHAP_TRY(zero_async(sums, (size_t)(3 * c + 1) * sizeof(double), S(stream)));
    int grid = grid_for(rows, 256 * 4, kNumSMs * 4);
    DISPATCH_C(c, (launch_k(consistency_fwd_kernel<C>, grid, 256, 0, S(stream), logits, target, mask, dist, rows, sums)));
    return launched("consistency_fwd_kernel");
}

extern "C" int chap_consistency_bwd(const float* logits, const float* target, const float* mask, int32_t dist, int64_t rows,
                                    int32_t c, const float* coef, float* dlogits, void* stream) {
    KernelTimer timer_("consistency_bwd", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(logits && target && coef && dlogits && rows > 0 && (dist == CHAP_DIST_KL || dist == CHAP_DIST_DICE), CHAP_ERR_BAD_ARG, "consistency_bwd: bad argument");
    CHAP_REQUIRE(row_aligned(logits, c) && row_aligned(target, c) && row_aligned(dlogits, c), CHAP_ERR_ALIGNMENT, "consistency_bwd: misaligned");
    int grid = grid_for(rows, 256 * 2);
    DISPATCH_C(c, (launch_k(consistency_bwd_kernel<C>, grid, 256, 0, S(stream), logits, target, mask, dist, rows, coef, dlogits)));
    return launched("consistency_bwd_kernel");
}

extern "C" int chap_patch_score(const float* knowledge, const int64_t* arg1, const int64_t* arg2, int32_t nd, int32_t n,
                                int32_t d, int32_t h, int32_t w, int32_t s, float* score, void* stream) {
    KernelTimer timer_("patch_score", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(knowledge && arg1 && arg2 && score && (nd == 2 || nd == 3) && n > 0 && s > 0, CHAP_ERR_BAD_ARG, "patch_score: bad argument");
    CHAP_REQUIRE(h % s == 0 && w % s == 0 && (nd == 2 ? d == 1 : d % s == 0), CHAP_ERR_BAD_ARG, "patch_score: size not divisible by scale_factor");
    const int64_t total = (int64_t)n * (nd == 3 ? d / s : 1) * (h / s) * (w / s);
    launch_k(patch_score_kernel, grid_for(total, 256), 256, 0, S(stream), knowledge, arg1, arg2, nd, d, h, w, s, total, score);
    return launched("patch_score_kernel");
}

extern "C" int chap_patch_mask(const float* score, const float* kth, int32_t nd, int32_t n, int32_t d, int32_t h, int32_t w,
                               int32_t s, float* mask, void* stream) {
    KernelTimer timer_("patch_mask", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(score && kth && mask && (nd == 2 || nd == 3) && n > 0 && s > 0, CHAP_ERR_BAD_ARG, "patch_mask: bad argument");
    const int64_t total = (int64_t)n * d * h * w;
    launch_k(patch_mask_kernel, grid_for(total, 256 * 4), 256, 0, S(stream), score, kth, nd, d, h, w, s, total, mask);
    return launched("patch_mask_kernel");
}
