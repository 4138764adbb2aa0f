// extern "C" convolution entry points: shape validation + dispatch (tcgen05 path / CUDA-core path).
#include <stdlib.h>
#include "common.cuh"
#include "conv_plan.cuh"

using namespace chap;

static bool use_tc(const Geom& g, bool dgrad) {
    return g_force_simt.load() == 0 && tc_supports(g, dgrad);
}

extern "C" size_t chap_conv_packed_elems(const chap_conv_desc* d) {
    Geom g{};
    if (resolve(d, g) != CHAP_OK) return 0;
    // room for the zero-padded tensor-core operand of thin heads, twice: TF32 hi half, then lo half (split-operand mode)
    return 2 * (size_t)g.taps * tc_pad16(g.cin) * tc_pad16(g.cout);
}

extern "C" int chap_conv_pack_weights(const chap_conv_desc* d, const float* w, float* w_fwd, float* w_dgrad, void* stream) {
    Geom g{};
    CHAP_TRY(resolve(d, g));
    CHAP_REQUIRE(w != nullptr, CHAP_ERR_BAD_ARG, "pack_weights: w is NULL");
    if (w_fwd) {
        PackSpec p = fwd_pack(g);
        const bool tc = use_tc(g, false);
        CHAP_TRY(launch_pack(w, w_fwd, p.taps, p.K, p.N, p.sk, p.sn, p.flip, tc ? 0 : 1, S(stream), tc ? tc_pad16(p.K) : 0, tc ? tc_pad16(p.N) : 0));
    }
    if (w_dgrad) {
        PackSpec p = dgrad_pack(g);
        const bool tc = use_tc(g, true);
        CHAP_TRY(launch_pack(w, w_dgrad, p.taps, p.K, p.N, p.sk, p.sn, p.flip, tc ? 0 : 1, S(stream), tc ? tc_pad16(p.K) : 0, tc ? tc_pad16(p.N) : 0));
    }
    return CHAP_OK;
}

extern "C" int chap_conv_pack_weights_batched(const chap_pack_item* items, int32_t n, void* stream) {
    CHAP_REQUIRE(items != nullptr && n >= 0, CHAP_ERR_BAD_ARG, "pack_weights_batched: bad argument");
    static thread_local PackBatch batch;
    batch.n = 0;
    auto push = [&](const float* w, float* out, const PackSpec& p, bool tc) -> int {
        PackSub& s = batch.sub[batch.n++];
        s.w = w; s.out = out; s.sk = p.sk; s.sn = p.sn; s.taps = p.taps; s.K = p.K; s.N = p.N;
        s.Kp = tc ? tc_pad16(p.K) : p.K; s.Np = tc ? tc_pad16(p.N) : p.N; s.flip = p.flip; s.kn_order = tc ? 0 : 1; s.pad_ = 0;
        if (batch.n == kPackBatch) { CHAP_TRY(launch_pack_batch(batch, S(stream))); batch.n = 0; }
        return CHAP_OK;
    };
    for (int i = 0; i < n; ++i) {
        Geom g{};
        CHAP_TRY(resolve(&items[i].desc, g));
        CHAP_REQUIRE(items[i].w != nullptr, CHAP_ERR_BAD_ARG, "pack_weights_batched: item %d has no weight", i);
        if (items[i].w_fwd) CHAP_TRY(push(items[i].w, items[i].w_fwd, fwd_pack(g), use_tc(g, false)));
        if (items[i].w_dgrad) CHAP_TRY(push(items[i].w, items[i].w_dgrad, dgrad_pack(g), use_tc(g, true)));
    }
    return launch_pack_batch(batch, S(stream));
}

extern "C" int chap_conv_fwd(const chap_conv_desc* d, const float* x, const float* w_fwd, const float* bias,
                             float* y, double* ch_sums, void* stream) {
    Geom g{};
    CHAP_TRY(resolve(d, g));
    CHAP_REQUIRE(x && w_fwd && y, CHAP_ERR_BAD_ARG, "conv_fwd: NULL pointer");
    if (use_tc(g, false)) {
        int rc = tc_conv(g, false, x, w_fwd, bias, y, ch_sums, S(stream));
        if (rc < 0) return rc;
        if (rc == 1) return CHAP_OK;
    }
    CHAP_TRY(simt_conv(fwd_op(g), x, w_fwd, bias, y, S(stream)));
    if (ch_sums) {      // CUDA-core path: statistics by a separate pass into slot 0, the other slots stay zero
        CHAP_TRY(zero_async(ch_sums, (size_t)CHAP_STAT_SLOTS * 2 * g.cout * sizeof(double), S(stream)));
        CHAP_TRY(channel_stats(y, g.out_rows, g.cout, ch_sums, S(stream)));
    }
    return CHAP_OK;
}

extern "C" int chap_conv_bn_fwd(const chap_conv_desc* d, const float* x, const float* w_fwd, const float* bias, float* y,
                                double* ch_sums, const chap_bn_train_args* bn, void* stream) {
    Geom g{};
    CHAP_TRY(resolve(d, g));
    CHAP_REQUIRE(x && w_fwd && y && ch_sums && bn && bn->gamma && bn->beta && bn->mean_invstd && bn->scale_shift, CHAP_ERR_BAD_ARG,
                 "conv_bn_fwd: NULL pointer");
    CHAP_REQUIRE((bn->running_mean == nullptr) == (bn->running_var == nullptr), CHAP_ERR_BAD_ARG, "conv_bn_fwd: running_mean / running_var must both be set or both NULL");
    if (use_tc(g, false) && getenv("CHAP_NO_BN_FOLD") == nullptr) {
        BnFold f{bn->gamma, bn->beta, bn->eps, bn->momentum, bn->running_mean, bn->running_var,
                 reinterpret_cast<long long*>(bn->num_batches_tracked), bn->mean_invstd, bn->scale_shift, 0.0, nullptr,
                 bn->stats_persistent ? 1 : 0};
        int rc = tc_conv(g, false, x, w_fwd, bias, y, ch_sums, S(stream), nullptr, 0, &f);
        if (rc < 0) return rc;
        if (rc == 1) return CHAP_OK;
    }
    CHAP_TRY(chap_conv_fwd(d, x, w_fwd, bias, y, ch_sums, stream));
    CHAP_TRY(chap_bn_finalize(ch_sums, CHAP_STAT_SLOTS, g.out_rows, bn->gamma, bn->beta, bn->eps, bn->momentum, bn->running_mean,
                              bn->running_var, bn->num_batches_tracked, bn->mean_invstd, bn->scale_shift, g.cout, stream));
    if (bn->stats_persistent)          // the three-launch path of unsupported shapes must also hand a persistent buffer back zeroed
        CHAP_TRY(zero_async(ch_sums, ((size_t)CHAP_STAT_SLOTS * 2 * g.cout + 1) * sizeof(double), S(stream)));
    return CHAP_OK;
}

extern "C" int chap_conv_bn_act_fwd(const chap_conv_desc* d, const float* x, const float* w_fwd, const float* bias,
                                    const float* scale_shift, float slope, const float* residual, float* y, void* stream) {
    Geom g{};
    CHAP_TRY(resolve(d, g));
    CHAP_REQUIRE(x && w_fwd && scale_shift && y, CHAP_ERR_BAD_ARG, "conv_bn_act_fwd: NULL pointer");
    if (use_tc(g, false)) {
        EvalEpilogue e{scale_shift, slope, residual};
        int rc = tc_conv(g, false, x, w_fwd, bias, y, nullptr, S(stream), nullptr, 0, nullptr, &e);
        if (rc < 0) return rc;
        if (rc == 1) return CHAP_OK;
    }
    // shapes outside the fused path (Cin = 1 stems, CUDA-core mode): convolution, then the BatchNorm/activation kernel in place
    CHAP_TRY(chap_conv_fwd(d, x, w_fwd, bias, y, nullptr, stream));
    const int64_t rps = (int64_t)g.oD * g.oH * g.oW;
    return chap_bn_act_fwd(y, scale_shift, slope, nullptr, nullptr, residual, g.n, rps, g.cout, y, stream);
}

extern "C" int chap_conv_dgrad(const chap_conv_desc* d, const float* dy, const float* w_dgrad, float* dx, void* stream) {
    Geom g{};
    CHAP_TRY(resolve(d, g));
    CHAP_REQUIRE(dy && w_dgrad && dx, CHAP_ERR_BAD_ARG, "conv_dgrad: NULL pointer");
    if (use_tc(g, true)) {
        int rc = tc_conv(g, true, dy, w_dgrad, nullptr, dx, nullptr, S(stream));
        if (rc < 0) return rc;
        if (rc == 1) return CHAP_OK;
    }
    return simt_conv(dgrad_op(g), dy, w_dgrad, nullptr, dx, S(stream));
}

extern "C" int chap_conv_dgrad_split_supported(const chap_conv_desc* d, int32_t ca) {
    Geom g{};
    if (resolve(d, g) != CHAP_OK) return 0;
    if (g.kind != CHAP_CONV_K3 && g.kind != CHAP_CONV_K1) return 0;
    return use_tc(g, true) && ca > 0 && ca < g.cin && ca % 16 == 0 && (g.cin - ca) % 16 == 0 ? 1 : 0;
}

extern "C" int chap_conv_dgrad_split(const chap_conv_desc* d, const float* dy, const float* w_dgrad, float* dx_a, int32_t ca,
                                     float* dx_b, void* stream) {
    Geom g{};
    CHAP_TRY(resolve(d, g));
    CHAP_REQUIRE(dy && w_dgrad && dx_a && dx_b, CHAP_ERR_BAD_ARG, "conv_dgrad_split: NULL pointer");
    CHAP_REQUIRE(chap_conv_dgrad_split_supported(d, ca), CHAP_ERR_BAD_ARG,
                 "conv_dgrad_split: shape not on the tensor-core split path (use chap_conv_dgrad + chap_split_channels)");
    int rc = tc_conv(g, true, dy, w_dgrad, nullptr, dx_a, nullptr, S(stream), dx_b, ca);
    if (rc < 0) return rc;
    CHAP_REQUIRE(rc == 1, CHAP_ERR_BAD_ARG, "conv_dgrad_split: tensor-core path declined the shape");
    return CHAP_OK;
}

extern "C" size_t chap_conv_wgrad_workspace_bytes(const chap_conv_desc* d) {
    Geom g{};
    if (resolve(d, g) != CHAP_OK) return 0;
    // [2 * cout doubles for the bias gradient][pad to 256][taps * cin * cout floats: the tensor-core kernel accumulates the
    // gradient in a [tap][M][N] layout with vector reductions, then a small kernel writes the torch layout]
    // pair-packed weight gradient of the 16 -> 16 layers: [+ 4 x that for the gradient of the equivalent 32 -> 32 convolution]
    const size_t base = (((size_t)2 * g.cout * sizeof(double) + 255) & ~(size_t)255) + (size_t)g.taps * g.cin * g.cout * sizeof(float);
    return ((base + 255) & ~(size_t)255) + (size_t)4 * g.taps * g.cin * g.cout * sizeof(float);
}

// Pair-packed weight gradient (round 2).  The tensor-core weight gradient of the thin layers is bound by the NUMBER of
// tcgen05.mma instructions the single issuing thread can emit (K = 8 pixels per MMA whatever M and N are: 295 k MMAs per launch for
// 16 -> 16 @ 12x256^2 = 2,000 per CTA x ~40 ns = the measured 85 us), and a 16-channel tensor wastes half of every 128-byte operand
// row on zero fill.  A channels-last [.., W, 16] tensor IS a [.., W/2, 32] tensor (two pixels per row), so the layer is handed to the same
// kernel as a 32 -> 32 convolution on the half-width image: half the MMAs (N = 32 costs the same issue slot), no zero fill.  Its gradient
// dW'[(p', co), (p, ci), ky, e] = sum_j dy[2j + p', co] x[2(j + e) + p, ci] holds every product of pixels q = 2j + p' and q + s with
// s = 2e + p - p'; the 3x3 gradient is the fold  dW[.., kx = s + 1] = sum over {(e, p, p') : 2e + p - p' = s}:
//   s =  0: (0,0,0) + (0,1,1)      s = +1: (0,1,0) + (+1,0,1)      s = -1: (0,0,1) + (-1,1,0)
// Image-row ends are exact because W is even: pixel -1 / W live in pairs -1 / W/2, which the TMA zero-fills as a whole.
static bool pair_wgrad_ok(const Geom& g) {
    static const bool off = getenv("CHAP_WG_NO_PAIR") != nullptr;
    // cout = 4: the class heads (dy pairs are 8 floats, zero-filled to a 128-byte row by the TMA like the 4-channel rows were before)
    return !off && g.kind == CHAP_CONV_K3 && g.cin == 16 && (g.cout == 16 || g.cout == 4) && g.iW % 2 == 0 && g.iW / 2 >= 8 && g.iH >= 10;
}

__global__ void __launch_bounds__(256)
pair_fold_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int cout, int planes, int accumulate) {
    pdl_enter();
    // dwp: torch layout [2 cout (p', co)][32 (p, ci)][planes (kz, ky)][3 (e + 1)]; dw: [cout][16 ci][planes][3 kx]
    const int total = cout * 16 * planes * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int kx = i % 3, pl = (i / 3) % planes, ci = (i / (3 * planes)) % 16, co = i / (3 * planes * 16);
        auto at = [&](int pq, int pp, int e1) {               // pq = p' (dy parity), pp = p (x parity), e1 = e + 1
            return dwp[((size_t)((pq * cout + co) * 32 + (pp * 16 + ci)) * planes + pl) * 3 + e1];
        };
        float v;
        if (kx == 1) v = at(0, 0, 1) + at(1, 1, 1);
        else if (kx == 2) v = at(0, 1, 1) + at(1, 0, 2);
        else v = at(1, 0, 1) + at(0, 1, 0);
        dw[i] = accumulate ? dw[i] + v : v;
    }
}

static int conv_wgrad_impl(const chap_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias,
                           void* workspace, size_t workspace_bytes, float* scratch, bool accumulate, void* stream) {
    Geom g{};
    CHAP_TRY(resolve(d, g));
    CHAP_REQUIRE(x && dy && dw, CHAP_ERR_BAD_ARG, "conv_wgrad: NULL pointer");
    SimtOp op = fwd_op(g);
    PackSpec p = fwd_pack(g);          // dw has the torch layout: same strides as reading w
    int handled = 0;
    // thin layers (Cin * Cout small): the tensor core cannot be filled (output 16 x 144) and every K = 8 MMA costs
    // ~130 cycles, so the register-tiled CUDA-core kernel wins (measured; threshold overridable for experiments)
    static const int thin_max = getenv("CHAP_THIN_MAX") ? atoi(getenv("CHAP_THIN_MAX")) : 0;
    const bool thin = g.kind == CHAP_CONV_K3 && g.cin % 4 == 0 && g.cout % 4 == 0 && g.cin * g.cout <= thin_max;
    const size_t pair_off = ((((size_t)2 * g.cout * sizeof(double) + 255) & ~(size_t)255) + (size_t)g.taps * g.cin * g.cout * sizeof(float) + 255) & ~(size_t)255;
    if (g_force_simt.load() == 0 && !thin && pair_wgrad_ok(g) && workspace &&
        workspace_bytes >= pair_off + (size_t)4 * g.taps * g.cin * g.cout * sizeof(float)) {
        Geom g2 = g;                                   // the same memory seen as a 32 -> 32 convolution on the half-width image
        g2.cin = 32; g2.cout = 2 * g.cout; g2.iW = g.iW / 2; g2.oW = g.oW / 2; g2.in_rows = g.in_rows / 2; g2.out_rows = g.out_rows / 2;
        if (tc_wgrad_supports(g2)) {
            float* dwp = reinterpret_cast<float*>(static_cast<char*>(workspace) + pair_off);
            handled = tc_wgrad(g2, x, dy, dwp, S(stream), nullptr, false);
            if (handled < 0) return handled;
            if (handled) {
                launch_k(pair_fold_kernel, (g.cout * 16 * g.taps + 255) / 256, 256, 0, S(stream), dwp, dw, g.cout, g.taps / 3, accumulate ? 1 : 0);
                CHAP_TRY(launched("pair_fold_kernel"));
            }
        }
    }
    if (!handled && g_force_simt.load() == 0 && tc_wgrad_supports(g) && !thin) {
        handled = tc_wgrad(g, x, dy, dw, S(stream), scratch, accumulate);
        if (handled < 0) return handled;
    }
    if (!handled) CHAP_TRY(simt_wgrad(op, x, dy, dw, (int64_t)g.taps * g.cin * g.cout, p.sk, p.sn, S(stream), accumulate));
    if (dbias) {
        CHAP_REQUIRE(workspace && workspace_bytes >= (size_t)2 * g.cout * sizeof(double), CHAP_ERR_WORKSPACE,
                     "conv_wgrad: workspace too small (%zu bytes)", workspace_bytes);
        CHAP_TRY(channel_stats(dy, g.out_rows, g.cout, (double*)workspace, S(stream)));
        CHAP_TRY(sums_to_float((const double*)workspace, dbias, g.cout, S(stream), accumulate));
    }
    return CHAP_OK;
}

extern "C" int chap_conv_wgrad(const chap_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias,
                               void* workspace, size_t workspace_bytes, void* stream) {
    Geom g{};
    CHAP_TRY(resolve(d, g));
    const size_t off = ((size_t)2 * g.cout * sizeof(double) + 255) & ~(size_t)255;
    float* scratch = (workspace && workspace_bytes >= off + (size_t)g.taps * g.cin * g.cout * sizeof(float))
                         ? reinterpret_cast<float*>(static_cast<char*>(workspace) + off) : nullptr;
    return conv_wgrad_impl(d, x, dy, dw, dbias, workspace, workspace_bytes, scratch, false, stream);
}

extern "C" int chap_conv_wgrad_acc(const chap_conv_desc* d, const float* x, const float* dy, float* dw_acc, float* dbias_acc,
                                   void* workspace, size_t workspace_bytes, float* zeroed_scratch, void* stream) {
    return conv_wgrad_impl(d, x, dy, dw_acc, dbias_acc, workspace, workspace_bytes, zeroed_scratch, true, stream);
}
