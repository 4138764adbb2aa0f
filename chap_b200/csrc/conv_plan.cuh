// Mapping of the four public conv kinds x {fwd, dgrad, wgrad} onto the generic kernels.
#pragma once
#include "common.cuh"

namespace chap {

struct SimtOp {
    int up2;             // 0: gather form, 1: up2 (scatter) form
    int nd;
    int iD, iH, iW;      // spatial size of the op's INPUT rows
    int oD, oH, oW;      // spatial size of the op's OUTPUT rows
    int K, N;            // reduction channels, output channels
    int ksz, stride, pad, taps;
    int64_t in_rows, out_rows;
};

// how to read the torch-layout weight for a packed [tap][K][N] operand
struct PackSpec { int taps, K, N; int64_t sk, sn; int flip; };

inline SimtOp fwd_op(const Geom& g) {
    SimtOp o{};
    o.nd = g.nd; o.iD = g.iD; o.iH = g.iH; o.iW = g.iW; o.oD = g.oD; o.oH = g.oH; o.oW = g.oW;
    o.K = g.cin; o.N = g.cout; o.taps = g.taps; o.in_rows = g.in_rows; o.out_rows = g.out_rows;
    switch (g.kind) {
        case CHAP_CONV_K3: o.ksz = 3; o.stride = 1; o.pad = 1; break;
        case CHAP_CONV_K1: o.ksz = 1; o.stride = 1; o.pad = 0; break;
        case CHAP_CONV_DOWN2: o.ksz = 2; o.stride = 2; o.pad = 0; break;
        default: o.up2 = 1; o.ksz = 2; o.stride = 2; o.pad = 0; break;
    }
    return o;
}

// the data gradient is itself a convolution of dy (output geometry) producing dx (input geometry)
inline SimtOp dgrad_op(const Geom& g) {
    SimtOp o{};
    o.nd = g.nd; o.iD = g.oD; o.iH = g.oH; o.iW = g.oW; o.oD = g.iD; o.oH = g.iH; o.oW = g.iW;
    o.K = g.cout; o.N = g.cin; o.taps = g.taps; o.in_rows = g.out_rows; o.out_rows = g.in_rows;
    switch (g.kind) {
        case CHAP_CONV_K3: o.ksz = 3; o.stride = 1; o.pad = 1; break;
        case CHAP_CONV_K1: o.ksz = 1; o.stride = 1; o.pad = 0; break;
        case CHAP_CONV_DOWN2: o.up2 = 1; o.ksz = 2; o.stride = 2; o.pad = 0; break;   // scatter dy back
        default: o.ksz = 2; o.stride = 2; o.pad = 0; break;                            // UP2: gather k2 s2
    }
    return o;
}

inline PackSpec fwd_pack(const Geom& g) {
    const int64_t T = g.taps;
    if (g.kind == CHAP_CONV_UP2) return {g.taps, g.cin, g.cout, (int64_t)g.cout * T, T, 0};   // w[ci][co][t]
    return {g.taps, g.cin, g.cout, T, (int64_t)g.cin * T, 0};                                   // w[co][ci][t]
}
inline PackSpec dgrad_pack(const Geom& g) {
    const int64_t T = g.taps;
    if (g.kind == CHAP_CONV_UP2) return {g.taps, g.cout, g.cin, T, (int64_t)g.cout * T, 0};
    return {g.taps, g.cout, g.cin, (int64_t)g.cin * T, T, g.kind == CHAP_CONV_K3 ? 1 : 0};
}

// one packed operand of one layer inside a batched pack launch (kernel parameter: 56 B x 64 = 3.5 KB)
struct PackSub { const float* w; float* out; int64_t sk, sn; int taps, K, N, Kp, Np, flip, kn_order, pad_; };
constexpr int kPackBatch = 64;
struct PackBatch { PackSub sub[kPackBatch]; int n; };
int launch_pack_batch(const PackBatch& b, cudaStream_t st);
int launch_pack(const float* w, float* out, int taps, int K, int N, int64_t sk, int64_t sn, int flip,
                int kn_order, cudaStream_t st, int Kp = 0, int Np = 0);
// tensor-core operands need K, N >= 16: the 4- / 8-channel heads are zero-padded in the packed weight
inline int tc_pad16(int c) { return c < 16 ? 16 : c; }
int simt_conv(const SimtOp& op, const float* in, const float* wp, const float* bias, float* out, cudaStream_t st);
int simt_wgrad(const SimtOp& op, const float* a, const float* b, float* dw, int64_t dw_elems,
               int64_t sk, int64_t sn, cudaStream_t st, bool accumulate = false);

// tensor-core path (conv_tc.cu): returns 1 if it handled the op, 0 if the shape is not supported
// (caller falls through to the CUDA-core kernels), <0 on error.
// BatchNorm finalize folded into the convolution kernel (done by its last thread block); mi == nullptr: not requested
struct BnFold {
    const float* gamma; const float* beta; float eps, momentum;
    float* rmean; float* rvar; long long* nbt;
    float* mi; float* ss;
    double count;                // elements per channel (N * spatial of the output)
    unsigned* counter;           // zero-initialised ticket counter (the double after the statistics slots)
    int rezero;                  // 1: the statistics buffer is persistent -- zero on entry, zeroed again by the finalizing block
};
// out_b != nullptr: output channels [0, ca) go to `out` (row stride ca), [ca, N) to `out_b` (row stride N - ca)
// inference epilogue: y = act(scale[c] * (conv + bias) + shift[c]) (+ residual of the output's shape); forward only
struct EvalEpilogue { const float* scale_shift; float slope; const float* residual; };
int tc_conv(const Geom& g, bool dgrad, const float* in, const float* wp, const float* bias, float* out,
            double* ch_sums, cudaStream_t st, float* out_b = nullptr, int ca = 0, const BnFold* bn = nullptr,
            const EvalEpilogue* epi = nullptr);
bool tc_supports(const Geom& g, bool dgrad);
// tensor-core weight gradient (wgrad_tc.cu): 1 handled, 0 unsupported shape, <0 error.  dw is overwritten.
// acc_ws (nullable): taps * cin * cout floats of scratch; when given, the (non-swap) kernel reduces into a [tap][M][N] layout
// with 128-bit vector reductions and a small kernel writes the torch layout afterwards.
// accumulate: dw += dW (dw is NOT zeroed first) and acc_ws, when given, must be all-zero on entry and is left all-zero on return
int tc_wgrad(const Geom& g, const float* x, const float* dy, float* dw, cudaStream_t st, float* acc_ws = nullptr, bool accumulate = false);
bool tc_wgrad_supports(const Geom& g);

int channel_stats(const float* y, int64_t rows, int c, double* sums, cudaStream_t st);
int sums_to_float(const double* sums, float* out, int c, cudaStream_t st, bool accumulate = false);

}  // namespace chap
