// CUDA-core (fp32 FMA) convolution kernels: the permanent path for the layers that cannot
// feed a tensor-core tile (Cin = 1 stems: K = 9/27; Cout = 4/2 heads) and the debug path
// (chap_set_force_simt) for every other layer.  All four conv kinds of chap_b200.h are expressed
// as one of two generic forms over channels-last rows:
//   gather : out[m, n] = sum_t sum_k in[src(m, t), k] * Wp[t][k][n]      (K3, K1, DOWN2 and their
//            stride-1 data gradients; dgrad of UP2)
//   up2    : out[dst(p, t), n] = sum_k in[p, k] * Wp[t][k][n]            (UP2; dgrad of DOWN2)
// Wp is the packed weight [tap][K][N] (N contiguous) produced by pack_weights_kernel.
#include "common.cuh"
#include "conv_plan.cuh"

namespace chap {

__device__ __forceinline__ void decode_row(int64_t m, int D, int H, int W, int& n, int& d, int& h, int& w) {
    w = (int)(m % W); m /= W;
    h = (int)(m % H); m /= H;
    d = (int)(m % D); n = (int)(m / D);
}

// ---------------------------------------------------------------- weight packing
// out[t][k][n] (kn_order = 1) or out[t][n][k] (kn_order = 0) = w[k*sk + n*sn + tmap(t)]; the packed operand may be
// zero-padded to Kp x Np (tensor-core path of the 4- / 8-channel heads: the MMA needs K, N >= 16)
__global__ void pack_weights_kernel(const float* __restrict__ w, float* __restrict__ out,
                                    int taps, int K, int N, int Kp, int Np, int64_t sk, int64_t sn, int flip, int kn_order, int rt) {
    pdl_enter();
    int64_t total = (int64_t)taps * Kp * Np;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int t = (int)(i / ((int64_t)Kp * Np));
        int r = (int)(i % ((int64_t)Kp * Np));
        int k, n;
        if (kn_order) { k = r / Np; n = r % Np; } else { n = r / Kp; k = r % Kp; }
        int ts = flip ? (taps - 1 - t) : t;
        const float v = (k < K && n < N) ? w[k * sk + n * sn + ts] : 0.f;
        if (rt) {       // tensor-core operand: TF32 halves w = hi + lo for the split-operand mode (plain TF32 reads hi only)
            const float hi = tf32_rn(v, 1);
            out[i] = hi;
            out[total + i] = tf32_rn(v - hi, 1);
        } else {
            out[i] = v;
        }
    }
}

// All weights of a model in one launch (blockIdx.y = sub-item: one packed operand of one layer).  The trainer re-packs every
// conv weight once per iteration (the optimiser just changed them): 72 launches of the single-layer kernel became 2 of this one.
__global__ void pack_weights_batched_kernel(const __grid_constant__ PackBatch b) {
    pdl_enter();
    const PackSub& s = b.sub[blockIdx.y];
    const int64_t total = (int64_t)s.taps * s.Kp * s.Np;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int t = (int)(i / ((int64_t)s.Kp * s.Np));
        int r = (int)(i % ((int64_t)s.Kp * s.Np));
        int k, n;
        if (s.kn_order) { k = r / s.Np; n = r % s.Np; } else { n = r / s.Kp; k = r % s.Kp; }
        int ts = s.flip ? (s.taps - 1 - t) : t;
        const float v = (k < s.K && n < s.N) ? s.w[k * s.sk + n * s.sn + ts] : 0.f;
        if (s.kn_order == 0) {                  // tensor-core operand: TF32 hi / lo halves (see pack_weights_kernel)
            const float hi = tf32_rn(v, 1);
            s.out[i] = hi;
            s.out[total + i] = tf32_rn(v - hi, 1);
        } else {
            s.out[i] = v;
        }
    }
}

int launch_pack_batch(const PackBatch& b, cudaStream_t st) {
    if (b.n <= 0) return CHAP_OK;
    launch_k(pack_weights_batched_kernel, dim3(24, (unsigned)b.n), 256, 0, st, b);
    return launched("pack_weights_batched_kernel");
}

int launch_pack(const float* w, float* out, int taps, int K, int N, int64_t sk, int64_t sn, int flip,
                int kn_order, cudaStream_t st, int Kp, int Np) {
    if (Kp < K) Kp = K;
    if (Np < N) Np = N;
    int64_t total = (int64_t)taps * Kp * Np;
    launch_k(pack_weights_kernel, grid_for(total, 256, 1024), 256, 0, st, w, out, taps, K, N, Kp, Np, sk, sn, flip, kn_order,
                                                                    kn_order == 0 ? 1 : 0);   // kn_order 0 = tensor-core operand
    return launched("pack_weights_kernel");
}

// ---------------------------------------------------------------- gather conv
template <int CO_T, bool VEC4>
__global__ void __launch_bounds__(128)
conv_gather_kernel(SimtOp op, int all_w, const float* __restrict__ in, const float* __restrict__ wp,
                   const float* __restrict__ bias, float* __restrict__ out) {
    pdl_enter();
    extern __shared__ float ws_all[];              // [K][CO_T] per tap, or [taps][K][CO_T] when everything fits (all_w)
    const int n0 = blockIdx.y * CO_T;
    const int64_t m = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const bool live = m < op.out_rows;
    int bn = 0, od = 0, oh = 0, ow = 0;
    if (live) decode_row(m, op.oD, op.oH, op.oW, bn, od, oh, ow);
    float acc[CO_T];
#pragma unroll
    for (int j = 0; j < CO_T; ++j) acc[j] = 0.f;

    const int kd_n = op.nd == 3 ? op.ksz : 1;
    if (all_w) {
        for (int i = threadIdx.x; i < op.taps * op.K * CO_T; i += 128) {
            int tk = i / CO_T, j = i % CO_T;
            ws_all[i] = (n0 + j < op.N) ? wp[(int64_t)tk * op.N + n0 + j] : 0.f;
        }
        __syncthreads();
    }
    int t = 0;
    for (int kz = 0; kz < kd_n; ++kz)
        for (int ky = 0; ky < op.ksz; ++ky)
            for (int kx = 0; kx < op.ksz; ++kx, ++t) {
                const float* ws = all_w ? ws_all + (size_t)t * op.K * CO_T : ws_all;
                if (!all_w) {
                    __syncthreads();
                    for (int i = threadIdx.x; i < op.K * CO_T; i += 128) {
                        int k = i / CO_T, j = i % CO_T;
                        ws_all[i] = (n0 + j < op.N) ? wp[((int64_t)t * op.K + k) * op.N + n0 + j] : 0.f;
                    }
                    __syncthreads();
                }
                int id = od * op.stride + kz - (op.nd == 3 ? op.pad : 0);
                int ih = oh * op.stride + ky - op.pad;
                int iw = ow * op.stride + kx - op.pad;
                bool ok = live && id >= 0 && id < op.iD && ih >= 0 && ih < op.iH && iw >= 0 && iw < op.iW;
                if (!ok) continue;
                const float* src = in + ((((int64_t)bn * op.iD + id) * op.iH + ih) * op.iW + iw) * op.K;
                if (VEC4) {
                    for (int k = 0; k < op.K; k += 4) {
                        float4 x = __ldg(reinterpret_cast<const float4*>(src + k));
                        const float* wr = ws + k * CO_T;
#pragma unroll
                        for (int j = 0; j < CO_T; ++j) {
                            acc[j] = fmaf(x.x, wr[j], acc[j]);
                            acc[j] = fmaf(x.y, wr[CO_T + j], acc[j]);
                            acc[j] = fmaf(x.z, wr[2 * CO_T + j], acc[j]);
                            acc[j] = fmaf(x.w, wr[3 * CO_T + j], acc[j]);
                        }
                    }
                } else {
                    for (int k = 0; k < op.K; ++k) {
                        float x = __ldg(src + k);
                        const float* wr = ws + k * CO_T;
#pragma unroll
                        for (int j = 0; j < CO_T; ++j) acc[j] = fmaf(x, wr[j], acc[j]);
                    }
                }
            }
    if (!live) return;
    float* dst = out + m * op.N + n0;
    if (CO_T % 4 == 0 && op.N % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
        // 128-bit stores: a thread's CO_T outputs are contiguous (scalar stores at a 4 * N byte lane stride wasted 7/8 of every sector)
#pragma unroll
        for (int j = 0; j < CO_T; j += 4) {
            if (n0 + j < op.N) {
                float4 o = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                if (bias) { o.x += bias[n0 + j]; o.y += bias[n0 + j + 1]; o.z += bias[n0 + j + 2]; o.w += bias[n0 + j + 3]; }
                *reinterpret_cast<float4*>(dst + j) = o;
            }
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < CO_T; ++j)
        if (n0 + j < op.N) dst[j] = acc[j] + (bias ? bias[n0 + j] : 0.f);
}

// ---------------------------------------------------------------- up2 (transposed k2 s2) conv
template <int CO_T, bool VEC4>
__global__ void __launch_bounds__(128)
conv_up2_kernel(SimtOp op, int blocks_per_tap, const float* __restrict__ in, const float* __restrict__ wp,
                const float* __restrict__ bias, float* __restrict__ out) {
    pdl_enter();
    extern __shared__ float ws[];                  // [K][CO_T]
    const int t = blockIdx.x / blocks_per_tap;
    const int n0 = blockIdx.y * CO_T;
    const int64_t p = (int64_t)(blockIdx.x % blocks_per_tap) * 128 + threadIdx.x;
    for (int i = threadIdx.x; i < op.K * CO_T; i += 128) {
        int k = i / CO_T, j = i % CO_T;
        ws[i] = (n0 + j < op.N) ? wp[((int64_t)t * op.K + k) * op.N + n0 + j] : 0.f;
    }
    __syncthreads();
    if (p >= op.in_rows) return;
    int bn, id, ih, iw;
    decode_row(p, op.iD, op.iH, op.iW, bn, id, ih, iw);
    int kz = op.nd == 3 ? (t >> 2) : 0, ky = (t >> 1) & 1, kx = t & 1;
    float acc[CO_T];
#pragma unroll
    for (int j = 0; j < CO_T; ++j) acc[j] = 0.f;
    const float* src = in + p * op.K;
    if (VEC4) {
        for (int k = 0; k < op.K; k += 4) {
            float4 x = __ldg(reinterpret_cast<const float4*>(src + k));
            const float* wr = ws + k * CO_T;
#pragma unroll
            for (int j = 0; j < CO_T; ++j) {
                acc[j] = fmaf(x.x, wr[j], acc[j]);
                acc[j] = fmaf(x.y, wr[CO_T + j], acc[j]);
                acc[j] = fmaf(x.z, wr[2 * CO_T + j], acc[j]);
                acc[j] = fmaf(x.w, wr[3 * CO_T + j], acc[j]);
            }
        }
    } else {
        for (int k = 0; k < op.K; ++k) {
            float x = __ldg(src + k);
            const float* wr = ws + k * CO_T;
#pragma unroll
            for (int j = 0; j < CO_T; ++j) acc[j] = fmaf(x, wr[j], acc[j]);
        }
    }
    int od = op.nd == 3 ? 2 * id + kz : 0, oh = 2 * ih + ky, ow = 2 * iw + kx;
    float* dst = out + ((((int64_t)bn * op.oD + od) * op.oH + oh) * op.oW + ow) * op.N + n0;
#pragma unroll
    for (int j = 0; j < CO_T; ++j)
        if (n0 + j < op.N) dst[j] = acc[j] + (bias ? bias[n0 + j] : 0.f);
}

// ---------------------------------------------------------------- Cin = 1 stems (k3, pad 1, 16 output channels)
// One thread computes PX consecutive pixels of a row x 16 channels: the PX + 2 inputs of a (kz, ky) line are loaded once and every
// weight read from shared memory feeds PX FMAs (the generic gather kernel is bound by its one shared-memory read per FMA:
// 432 per voxel in 3D).  Output: PX x 64 contiguous bytes per thread.
template <int PX>
__global__ void __launch_bounds__(128)
stem_conv_kernel(SimtOp op, const float* __restrict__ in, const float* __restrict__ wp, const float* __restrict__ bias, float* __restrict__ out) {
    pdl_enter();
    __shared__ float ws[27 * 16];
    for (int i = threadIdx.x; i < op.taps * 16; i += 128) ws[i] = wp[i];            // packed [tap][K = 1][N = 16]
    __syncthreads();
    const int wq = (op.oW + PX - 1) / PX;                                            // thread columns per row
    const int64_t total = (int64_t)op.out_rows / op.oW * wq;
    const int64_t id = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (id >= total) return;
    const int q = (int)(id % wq);
    int64_t r = id / wq;                                                             // (n, d, h) row index
    const int oh = (int)(r % op.oH); r /= op.oH;
    const int od = (int)(r % op.oD); const int bn = (int)(r / op.oD);
    const int w0 = q * PX;
    float acc[PX][16];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[p][j] = 0.f;
    const int kd_n = op.nd == 3 ? 3 : 1;
    for (int kz = 0; kz < kd_n; ++kz) {
        const int iz = od + kz - (op.nd == 3 ? 1 : 0);
        if (iz < 0 || iz >= op.iD) continue;
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = oh + ky - 1;
            if (iy < 0 || iy >= op.iH) continue;
            const float* row = in + (((int64_t)bn * op.iD + iz) * op.iH + iy) * op.iW;
            float xv[PX + 2];
#pragma unroll
            for (int p = 0; p < PX + 2; ++p) {
                const int ix = w0 + p - 1;
                xv[p] = (ix >= 0 && ix < op.iW) ? __ldg(row + ix) : 0.f;
            }
            const float* wt = ws + ((kz * 3 + ky) * 3) * 16;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float wv = wt[kx * 16 + j];
#pragma unroll
                    for (int p = 0; p < PX; ++p) acc[p][j] = fmaf(xv[p + kx], wv, acc[p][j]);
                }
        }
    }
    float b[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) b[j] = bias ? bias[j] : 0.f;
    float* dst = out + ((((int64_t)bn * op.oD + od) * op.oH + oh) * op.oW + w0) * 16;
#pragma unroll
    for (int p = 0; p < PX; ++p) {
        if (w0 + p >= op.oW) break;
#pragma unroll
        for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(dst + p * 16 + j) = make_float4(acc[p][j] + b[j], acc[p][j + 1] + b[j + 1], acc[p][j + 2] + b[j + 2], acc[p][j + 3] + b[j + 3]);
    }
}

int simt_conv(const SimtOp& op, const float* in, const float* wp, const float* bias, float* out, cudaStream_t st) {
    const double rows_mac = (double)(op.up2 ? op.in_rows : op.out_rows);
    KernelTimer timer(timer_name(op.up2 ? "conv_simt_up2" : "conv_simt_gather", op.taps, op.K, op.N, op.oW, op.oH, op.oD, op.out_rows), 2.0 * rows_mac * op.K * op.N * op.taps,
                      4.0 * ((double)op.in_rows * op.K + (double)op.out_rows * op.N + (double)op.taps * op.K * op.N), st);
    if (!op.up2 && op.K == 1 && op.N == 16 && op.ksz == 3 && op.stride == 1 && op.pad == 1 && aligned16(out)) {
        constexpr int PX = 4;
        const int64_t threads = op.out_rows / op.oW * ((op.oW + PX - 1) / PX);
        launch_k(stem_conv_kernel<PX>, (unsigned)((threads + 127) / 128), 128, 0, st, op, in, wp, bias, out);
        return launched("stem_conv_kernel");
    }
    const bool vec4 = (op.K % 4 == 0) && aligned16(in);
    const bool small = op.N <= 4;
    const int co_t = small ? 4 : 16;
    const size_t smem = (size_t)op.K * co_t * sizeof(float);
    dim3 block(128);
    if (op.up2) {
        int bpt = (int)((op.in_rows + 127) / 128);
        dim3 grid((unsigned)(bpt * op.taps), (unsigned)((op.N + co_t - 1) / co_t));
        if (small) {
            if (vec4) launch_k(conv_up2_kernel<4, true>, grid, block, smem, st, op, bpt, in, wp, bias, out);
            else launch_k(conv_up2_kernel<4, false>, grid, block, smem, st, op, bpt, in, wp, bias, out);
        } else {
            if (vec4) launch_k(conv_up2_kernel<16, true>, grid, block, smem, st, op, bpt, in, wp, bias, out);
            else launch_k(conv_up2_kernel<16, false>, grid, block, smem, st, op, bpt, in, wp, bias, out);
        }
        return launched("conv_up2_kernel");
    }
    dim3 grid((unsigned)((op.out_rows + 127) / 128), (unsigned)((op.N + co_t - 1) / co_t));
    const int all_w = (size_t)op.taps * smem <= 40 * 1024 ? 1 : 0;      // all taps' weights resident: no per-tap barrier
    const size_t gsmem = all_w ? (size_t)op.taps * smem : smem;
    if (small) {
        if (vec4) launch_k(conv_gather_kernel<4, true>, grid, block, gsmem, st, op, all_w, in, wp, bias, out);
        else launch_k(conv_gather_kernel<4, false>, grid, block, gsmem, st, op, all_w, in, wp, bias, out);
    } else {
        if (vec4) launch_k(conv_gather_kernel<16, true>, grid, block, gsmem, st, op, all_w, in, wp, bias, out);
        else launch_k(conv_gather_kernel<16, false>, grid, block, gsmem, st, op, all_w, in, wp, bias, out);
    }
    return launched("conv_gather_kernel");
}

// ---------------------------------------------------------------- weight gradient
// dW[t][k][n] = sum_rows A[rowA(r, t), k] * B[rowB(r, t), n]; written with torch-layout strides.
//   gather form: r = output row of the forward op, rowA = src(r, t) (may be out of range -> 0), rowB = r
//   up2 form   : r = input row p,                   rowA = p,                          rowB = dst(p, t)
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(SimtOp op, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ dw,
                  int64_t sk, int64_t sn, int64_t rows_per_split) {
    pdl_enter();
    __shared__ float As[64][16];
    __shared__ float Bs[64][16];
    __shared__ int64_t rowA[64], rowB[64];
    const int k_tiles = (op.K + 15) / 16;
    const int k0 = (blockIdx.x % k_tiles) * 16, n0 = (blockIdx.x / k_tiles) * 16;
    const int t = blockIdx.y;
    const int64_t total_rows = op.up2 ? op.in_rows : op.out_rows;
    const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
    const int64_t r_end = min(total_rows, r_begin + rows_per_split);
    const int kk = threadIdx.x & 15, nn = threadIdx.x >> 4;
    int kz, ky, kx;
    if (op.up2 || op.ksz == 2) { kz = op.nd == 3 ? (t >> 2) : 0; ky = (t >> 1) & 1; kx = t & 1; }
    else if (op.ksz == 3) { kz = op.nd == 3 ? t / 9 : 0; ky = (t / 3) % 3; kx = t % 3; }
    else { kz = ky = kx = 0; }
    float acc = 0.f;
    for (int64_t r0 = r_begin; r0 < r_end; r0 += 64) {
        __syncthreads();
        if (threadIdx.x < 64) {
            int64_t r = r0 + threadIdx.x, ra = -1, rb = -1;
            if (r < r_end) {
                int bn, d, h, w;
                if (op.up2) {
                    decode_row(r, op.iD, op.iH, op.iW, bn, d, h, w);
                    ra = r;
                    int od = op.nd == 3 ? 2 * d + kz : 0;
                    rb = (((int64_t)bn * op.oD + od) * op.oH + 2 * h + ky) * op.oW + 2 * w + kx;
                } else {
                    decode_row(r, op.oD, op.oH, op.oW, bn, d, h, w);
                    rb = r;
                    int id = d * op.stride + kz - (op.nd == 3 ? op.pad : 0);
                    int ih = h * op.stride + ky - op.pad;
                    int iw = w * op.stride + kx - op.pad;
                    if (id >= 0 && id < op.iD && ih >= 0 && ih < op.iH && iw >= 0 && iw < op.iW)
                        ra = (((int64_t)bn * op.iD + id) * op.iH + ih) * op.iW + iw;
                }
            }
            rowA[threadIdx.x] = ra; rowB[threadIdx.x] = rb;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int idx = threadIdx.x + i * 256, r = idx >> 4, c = idx & 15;
            int64_t ra = rowA[r], rb = rowB[r];
            As[r][c] = (ra >= 0 && k0 + c < op.K) ? __ldg(a + ra * op.K + k0 + c) : 0.f;
            Bs[r][c] = (rb >= 0 && n0 + c < op.N) ? __ldg(b + rb * op.N + n0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll 16
        for (int r = 0; r < 64; ++r) acc = fmaf(As[r][kk], Bs[r][nn], acc);
    }
    if (k0 + kk < op.K && n0 + nn < op.N)
        atomicAdd(dw + (k0 + kk) * sk + (n0 + nn) * sn + t, acc);
}

// Weight gradient of the thin k = 3 layers: the Cin = 1 stems (KC = 1, NC = 16) and the Cout = 4 heads (KC = 4, NC = 4):
//   dW[co][ci][tap] = sum_p x[p + tap, ci] * dy[p, co].
// The generic kernel would re-read dy once per tap; here one thread owns a pixel, reads its NC-channel dy vector ONCE
// and the 9 neighbouring KC-channel inputs of one kz-plane, and keeps 9 x KC x NC partial sums in registers across a
// grid-stride loop; then warp shuffle -> smem -> one atomic per weight and block.
// grid = (pixel slices, kz planes, (Cin / KC) * (Cout / NC)).
template <int KC, int NC>
__global__ void __launch_bounds__(128)
thin_wgrad_kernel(SimtOp op, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int64_t sk, int64_t sn) {
    pdl_enter();
    constexpr int NA = 9 * KC * NC;
    __shared__ float red[NA];
    const int kz = blockIdx.y;
    const int kchunks = op.K / KC;
    const int k0 = (blockIdx.z % kchunks) * KC, c0 = (blockIdx.z / kchunks) * NC;
    float acc[9][KC][NC];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int k = 0; k < KC; ++k)
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[t][k][c] = 0.f;
    for (int i = threadIdx.x; i < NA; i += 128) red[i] = 0.f;
    __syncthreads();
    for (int64_t r = (int64_t)blockIdx.x * 128 + threadIdx.x; r < op.out_rows; r += (int64_t)gridDim.x * 128) {
        int bn, d, h, w;
        decode_row(r, op.oD, op.oH, op.oW, bn, d, h, w);
        const int id = op.nd == 3 ? d + kz - 1 : 0;
        if (id < 0 || id >= op.iD) continue;
        float g[NC];
        const float4* gp = reinterpret_cast<const float4*>(dy + r * op.N + c0);
#pragma unroll
        for (int q = 0; q < NC / 4; ++q) { float4 v = __ldg(gp + q); g[4 * q] = v.x; g[4 * q + 1] = v.y; g[4 * q + 2] = v.z; g[4 * q + 3] = v.w; }
        const float* xb = x + ((((int64_t)bn * op.iD + id) * op.iH) * op.iW) * op.K + k0;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int ih = h + ky - 1;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int iw = w + kx - 1;
                const bool ok = ih >= 0 && ih < op.iH && iw >= 0 && iw < op.iW;
                float xv[KC];
                if (KC == 4) {
                    float4 v = ok ? __ldg(reinterpret_cast<const float4*>(xb + ((int64_t)ih * op.iW + iw) * op.K)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    xv[0] = v.x; xv[1 % KC] = v.y; xv[2 % KC] = v.z; xv[3 % KC] = v.w;
                } else {
                    xv[0] = ok ? __ldg(xb + ((int64_t)ih * op.iW + iw) * op.K) : 0.f;
                }
#pragma unroll
                for (int k = 0; k < KC; ++k)
#pragma unroll
                    for (int c = 0; c < NC; ++c) acc[ky * 3 + kx][k][c] = fmaf(xv[k], g[c], acc[ky * 3 + kx][k][c]);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int k = 0; k < KC; ++k)
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                float v = warp_sum(acc[t][k][c]);
                if ((threadIdx.x & 31) == 0) atomicAdd(&red[(t * KC + k) * NC + c], v);
            }
    __syncthreads();
    for (int i = threadIdx.x; i < NA; i += 128) {
        const int t = i / (KC * NC), k = (i / NC) % KC, c = i % NC;
        atomicAdd(dw + (int64_t)(c0 + c) * sn + (int64_t)(k0 + k) * sk + (kz * 9 + t), red[i]);
    }
}

// Weight gradient of the Cin = 1 stems (1 -> 16): dW[co][tap] = sum_p x[p + tap] * dy[p, co].  Four lanes share a pixel, each owning four
// output channels: the dy row of a pixel is ONE coalesced 64-byte read of the quad, the nine x taps are broadcast loads, and a thread keeps
// 9 x 4 partial sums (thin_wgrad_kernel<1, 16> kept 144 per thread at 128 threads per block: 100 us for a 100 MB dy in 2D, 261 us in 3D).
// grid = (pixel slices, kz planes, Cout / 16); 32-bit pixel arithmetic (rows < 2^31, checked by the host).
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(SimtOp op, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int64_t sn) {
    pdl_enter();
    __shared__ float red[9 * 16];
    const int kz = blockIdx.y, c0 = blockIdx.z * 16;
    const int q = threadIdx.x & 3;                         // channel quad of this lane
    float acc[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t) { acc[t][0] = 0.f; acc[t][1] = 0.f; acc[t][2] = 0.f; acc[t][3] = 0.f; }
    if (threadIdx.x < 144) red[threadIdx.x] = 0.f;
    __syncthreads();
    const uint32_t rows = (uint32_t)op.out_rows, W = (uint32_t)op.oW, H = (uint32_t)op.oH, D = (uint32_t)op.oD;
    // four pixels per trip with all four dy reads issued first: one 64-byte row in flight per quad is far too little memory-level
    // parallelism (91 us for a 100 MB dy); the loop is otherwise a chain of dependent DRAM round trips
    constexpr int U = 4;
    const uint32_t stride = gridDim.x * 64u;
    for (uint32_t r0 = blockIdx.x * 64u + (threadIdx.x >> 2); r0 < rows; r0 += U * stride) {
        float4 g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t r = r0 + u * stride;
            g[u] = r < rows ? __ldg(reinterpret_cast<const float4*>(dy + (int64_t)r * op.N + c0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t r = r0 + u * stride;
            if (r >= rows) break;
            uint32_t m = r;
            const int w = (int)(m % W); m /= W;
            const int h = (int)(m % H); m /= H;
            const int d = (int)(m % D); const uint32_t bn = m / D;
            const int id = op.nd == 3 ? d + kz - 1 : 0;
            if (id < 0 || id >= op.iD) continue;
            const float* xb = x + ((int64_t)bn * op.iD + id) * op.iH * op.iW;
            float xv[9];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int ih = h + ky - 1, iw = w + kx - 1;
                    const bool ok = ih >= 0 && ih < op.iH && iw >= 0 && iw < op.iW;
                    xv[ky * 3 + kx] = ok ? __ldg(xb + ih * op.iW + iw) : 0.f;
                }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                acc[t][0] = fmaf(xv[t], g[u].x, acc[t][0]); acc[t][1] = fmaf(xv[t], g[u].y, acc[t][1]);
                acc[t][2] = fmaf(xv[t], g[u].z, acc[t][2]); acc[t][3] = fmaf(xv[t], g[u].w, acc[t][3]);
            }
        }
    }
    // lanes with the same channel quad (lane & 3) are combined over the xor offsets 4, 8, 16; lanes 0..3 then hold the warp's sums
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float v = acc[t][c];
            v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
            if ((threadIdx.x & 31) < 4) atomicAdd(&red[t * 16 + q * 4 + c], v);
        }
    __syncthreads();
    if (threadIdx.x < 144) {
        const int t = threadIdx.x / 16, c = threadIdx.x % 16;
        atomicAdd(dw + (int64_t)(c0 + c) * sn + (kz * 9 + t), red[threadIdx.x]);
    }
}

// Weight gradient of the 1x1(x1) class heads (Cin = 16, Cout <= 4, e.g. the V-Net's 16 -> 2 output conv):
//   dW[co][ci] = sum_p dy[p, co] * x[p, ci]  -- K * N <= 64 sums kept in registers over a grid-stride loop, one pass over x and dy
// (the generic kernel re-reads the operands per 16 x 16 weight tile and spent 465 us on a 4 x 112 x 112 x 80 batch).
template <int K, int N>
__global__ void __launch_bounds__(256)
k1_head_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int64_t rows, int64_t sk, int64_t sn) {
    pdl_enter();
    float acc[N][K];
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
        for (int k = 0; k < K; ++k) acc[c][k] = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x; r < rows; r += (int64_t)gridDim.x * 256) {
        float xv[K], g[N];
#pragma unroll
        for (int k = 0; k < K; k += 4) {
            const float4 v = ldg_stream(reinterpret_cast<const float4*>(x + r * K + k));
            xv[k] = v.x; xv[k + 1] = v.y; xv[k + 2] = v.z; xv[k + 3] = v.w;
        }
#pragma unroll
        for (int c = 0; c < N; ++c) g[c] = __ldg(dy + r * N + c);
#pragma unroll
        for (int c = 0; c < N; ++c)
#pragma unroll
            for (int k = 0; k < K; ++k) acc[c][k] = fmaf(g[c], xv[k], acc[c][k]);
    }
    __shared__ float red[8][N * K];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float v = warp_sum(acc[c][k]);
            if (lane == 0) red[warp][c * K + k] = v;
        }
    __syncthreads();
    for (int i = threadIdx.x; i < N * K; i += 256) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][i];
        atomicAdd(dw + (int64_t)(i % K) * sk + (int64_t)(i / K) * sn, v);          // dw[k * sk + n * sn] (one tap)
    }
}

int simt_wgrad(const SimtOp& op, const float* a, const float* b, float* dw, int64_t dw_elems,
               int64_t sk, int64_t sn, cudaStream_t st, bool accumulate) {
    if (!accumulate) CHAP_TRY(zero_async(dw, dw_elems * sizeof(float), st));      // every kernel below ADDS into dw with atomics
    if (!op.up2 && op.ksz == 1 && op.K == 16 && (op.N == 2 || op.N == 4) && aligned16(a)) {
        KernelTimer timer(timer_name("conv_thin_wgrad", op.taps, op.K, op.N, op.oW, op.oH, op.oD, op.out_rows), 2.0 * (double)op.out_rows * op.K * op.N,
                          4.0 * ((double)op.in_rows * op.K + (double)op.out_rows * op.N), st);
        const int grid = grid_for(op.out_rows, 256 * 8, kNumSMs * 4);
        if (op.N == 2) launch_k(k1_head_wgrad_kernel<16, 2>, grid, 256, 0, st, a, b, dw, op.out_rows, sk, sn);
        else launch_k(k1_head_wgrad_kernel<16, 4>, grid, 256, 0, st, a, b, dw, op.out_rows, sk, sn);
        return launched("k1_head_wgrad_kernel");
    }
    const bool stem = op.K == 1 && op.N % 16 == 0;
    const bool head = op.N % 4 == 0 && op.K % 4 == 0 && op.K * op.N <= 1024 && aligned16(a);     // heads and other thin layers
    if (!op.up2 && op.ksz == 3 && op.stride == 1 && (stem || head) && aligned16(b)) {
        KernelTimer timer(timer_name("conv_thin_wgrad", op.taps, op.K, op.N, op.oW, op.oH, op.oD, op.out_rows), 2.0 * (double)op.out_rows * op.K * op.N * op.taps,
                          4.0 * ((double)op.in_rows * op.K + (double)op.out_rows * op.N), st);
        const int planes = op.nd == 3 ? 3 : 1;
        const int zdim = stem ? op.N / 16 : (op.K / 4) * (op.N / 4);
        int slices = (int)((op.out_rows + 128 * 8 - 1) / (128 * 8));
        const int cap = (kNumSMs * 6) / (planes * zdim);
        if (slices > cap) slices = cap < 1 ? 1 : cap;
        dim3 grid((unsigned)slices, (unsigned)planes, (unsigned)zdim);
        if (stem && op.out_rows < 0x7FFFFFFFll && getenv("CHAP_STEM_WGRAD_OLD") == nullptr) {
            int sl = (int)((op.out_rows + 64 * 16 - 1) / (64 * 16));
            // every block ends with 144 atomics on the SAME 144 addresses: the block count is what the tail costs (1184 blocks: ~60 us)
            static const int per_sm = getenv("CHAP_STEM_WG_BLOCKS") ? atoi(getenv("CHAP_STEM_WG_BLOCKS")) : 2;
            const int cap2 = (kNumSMs * per_sm) / (planes * zdim);
            if (sl > cap2) sl = cap2 < 1 ? 1 : cap2;
            launch_k(stem_wgrad_kernel, dim3((unsigned)sl, (unsigned)planes, (unsigned)zdim), 256, 0, st, op, a, b, dw, sn);
        } else if (stem) launch_k(thin_wgrad_kernel<1, 16>, grid, 128, 0, st, op, a, b, dw, sk, sn);
        else launch_k(thin_wgrad_kernel<4, 4>, grid, 128, 0, st, op, a, b, dw, sk, sn);
        return launched("thin_wgrad_kernel");
    }
    const int64_t rows = op.up2 ? op.in_rows : op.out_rows;
    KernelTimer timer(timer_name("conv_simt_wgrad", op.taps, op.K, op.N, op.oW, op.oH, op.oD, op.out_rows), 2.0 * (double)rows * op.K * op.N * op.taps,
                      4.0 * ((double)op.in_rows * op.K + (double)op.out_rows * op.N + (double)op.taps * op.K * op.N), st);
    const int tiles = ((op.K + 15) / 16) * ((op.N + 15) / 16);
    int64_t want_splits = (kNumSMs * 4 + (int64_t)tiles * op.taps - 1) / ((int64_t)tiles * op.taps);
    int64_t max_splits = (rows + 511) / 512;
    int64_t splits = want_splits < 1 ? 1 : (want_splits > max_splits ? max_splits : want_splits);
    if (splits > 65535) splits = 65535;
    int64_t rps = ((rows + splits - 1) / splits + 63) / 64 * 64;
    splits = (rows + rps - 1) / rps;
    dim3 grid((unsigned)tiles, (unsigned)op.taps, (unsigned)splits);
    launch_k(conv_wgrad_kernel, grid, 256, 0, st, op, a, b, dw, sk, sn, rps);
    return launched("conv_wgrad_kernel");
}

}  // namespace chap
