// Shared helpers for libchap_b200 (sm_100a).  Internal header.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <utility>
#include "../../include/chap_b200.h"

namespace chap {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;
extern std::atomic<int> g_force_simt;
extern std::atomic<int> g_pdl;            // 1: kernels are launched with programmatic stream serialization (see launch_k)
extern std::atomic<int> g_precise_max_c;  // 3xTF32 split-operand convolutions for layers with max(K, N) <= this (0: plain TF32)

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// call after every kernel launch: counts it and converts launch errors (capture safe)
inline int launched(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(CHAP_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    }
    return CHAP_OK;
}

// Programmatic dependent launch.  EVERY kernel of this library starts with pdl_enter() (or pdl_trigger() ... pdl_wait() around a
// prologue that touches no global memory) and is launched through launch_k() with cudaLaunchAttributeProgrammaticStreamSerialization:
//   griddepcontrol.launch_dependents  -- the NEXT kernel of the stream may be scheduled as soon as every CTA of this one has started
//                                        (its CTAs become resident on whatever the tail of this kernel leaves free and run their prologue)
//   griddepcontrol.wait               -- returns when the PREVIOUS kernel has completed and its memory is visible
// Nothing reads or writes global memory before the wait, so the stream order semantics are unchanged (158 GPU tests pass either way).
// MEASURED (tools/pdl_chain_bench.py, B200): a replayed CUDA graph already runs a dependent chain at 1.2-1.4 us per small kernel node;
// the attribute changes that by -0.14 us (16 KB tensors) to +0.4 us (1 MB tensors, the early-resident CTAs get in the way), and the
// whole 2D iteration by +0.2 ms (13.71 vs 13.49 ms).  griddepcontrol.wait waits for the COMPLETION of the previous grid, so only a
// prologue can overlap, never the tails -- not worth it here.  Worse: the two instructions are NOT free in a kernel launched without
// the attribute -- same box, same step, library built with and without them: 13.49 vs 13.26 ms (0.23 us for each of 980 kernels).
// So the default build compiles them out (pdl_* are empty, the attribute is never set); `make -C chap_b200/csrc pdl` builds
// lib/libchap_b200_pdl.so with them (-DCHAP_PDL_INSN), where CHAP_PDL=1 / chap_set_pdl(1) turn the attribute on.
#ifdef __CUDACC__
#ifndef CHAP_PDL_INSN          // default build: no griddepcontrol instructions at all (see above: their presence alone costs 0.23 ms per iteration)
__device__ __forceinline__ void pdl_trigger() {}
__device__ __forceinline__ void pdl_wait() {}
#else
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
__device__ __forceinline__ void pdl_enter() { pdl_trigger(); pdl_wait(); }

template <typename... P, typename... A>
inline void launch_k(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
#ifdef CHAP_PDL_INSN
    cfg.numAttrs = g_pdl.load(std::memory_order_relaxed) ? 1u : 0u;
#else
    cfg.numAttrs = 0u;            // kernels without griddepcontrol.wait must never be launched with the attribute
#endif
    (void)cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);      // errors are picked up by launched()
}
#endif

#define CHAP_REQUIRE(cond, code, ...) \
    do { if (!(cond)) return ::chap::fail(code, __VA_ARGS__); } while (0)
#define CHAP_CUDA(call) \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
        return ::chap::fail(CHAP_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)
#define CHAP_TRY(call) do { int rc_ = (call); if (rc_ != CHAP_OK) return rc_; } while (0)

// Optional live timing of kernel families (bench.py roofline): CUDA events recorded on the launching
// stream around a launch while chap_timing_enable(1) is in effect; aggregated by chap_timing_report().
struct KernelTimer {
    int slot;
    cudaStream_t st;
    KernelTimer(const char* name, double flops, double bytes, cudaStream_t stream);
    ~KernelTimer();
};

// family name, or (CHAP_TIMING_DETAIL set) family + shape, interned
const char* timer_name(const char* family, int taps, int k, int n, int w, int h, int d, int64_t rows);

// Zero `bytes` bytes (multiple of 4, 4-byte aligned) on the stream.  A memset node costs ~4 us inside a replayed CUDA graph
// (measured: 94 of them = 0.4 ms per iteration), a small kernel node far less -> zero-fill kernel by default
// (CHAP_ZERO_MEMSET=1 restores cudaMemsetAsync).
int zero_async(void* ptr, size_t bytes, cudaStream_t st);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;   // B200

inline int grid_for(int64_t work_items, int per_block, int max_blocks = kNumSMs * 16) {
    int64_t b = (work_items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Round-to-nearest (ties away) conversion to TF32, kept in an fp32 container (low 13 mantissa bits zero).
//
// What the tensor-core path does with fp32 operands -- MEASURED on B200 (tools/tf32_probe.py, profiles/r02_tf32_probe.md):
// one convolution layer run with raw / host-pre-rounded (cvt.rna emulation) / host-truncated operands gives bit-identical
// results for "raw" and "pre-rounded" on BOTH operands and matches an exact-arithmetic evaluation of rna(x) * rna(w) (and
// cuDNN's TF32 kernels) to 1e-7, while pre-truncated operands show the -3.5e-4 per-operand shrink that truncation must give.
// So operands loaded through CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 tensor maps ARE rounded to nearest (round 1's comment here
// claimed truncation; that was wrong), producer-side rounding changes nothing (`rt` below stays 0), and a TF32 layer of
// this library has exactly cuDNN-TF32's error.  The network-level gap to torch-eager "TF32" (1.8x in 2D) comes from cuDNN
// silently running its fp32 kernels on the 16-channel layers (layer c16: cuDNN allow_tf32 error 2e-7); the answer to that
// is the split-operand ("3xTF32") mode of conv_tc.cu, not a rounding switch.
__device__ __forceinline__ float tf32_rn(float x, int on) {
    if (!on) return x;
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
inline int round_tf32_on() { return 0; }

// streaming 128-bit load (read once) / store
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

struct Geom {           // resolved convolution geometry (D = 1 for 2D)
    int kind, nd, n;
    int iD, iH, iW, oD, oH, oW;
    int cin, cout;
    int taps;           // 3^nd, 1, 2^nd, 2^nd
    int64_t in_rows, out_rows;   // N * spatial
};

inline int resolve(const chap_conv_desc* d, Geom& g) {
    CHAP_REQUIRE(d != nullptr, CHAP_ERR_BAD_ARG, "conv desc is NULL");
    CHAP_REQUIRE(d->nd == 2 || d->nd == 3, CHAP_ERR_BAD_ARG, "conv nd must be 2 or 3 (got %d)", d->nd);
    CHAP_REQUIRE(d->n > 0 && d->in_h > 0 && d->in_w > 0 && d->in_d > 0 && d->cin > 0 && d->cout > 0,
                 CHAP_ERR_BAD_ARG, "conv desc has a non-positive size");
    CHAP_REQUIRE(d->nd == 3 || d->in_d == 1, CHAP_ERR_BAD_ARG, "2D conv needs in_d == 1");
    g.kind = d->kind; g.nd = d->nd; g.n = d->n;
    g.iD = d->in_d; g.iH = d->in_h; g.iW = d->in_w;
    g.cin = d->cin; g.cout = d->cout;
    const int p2 = d->nd == 2 ? 4 : 8, p3 = d->nd == 2 ? 9 : 27;
    switch (d->kind) {
        case CHAP_CONV_K3: g.oD = g.iD; g.oH = g.iH; g.oW = g.iW; g.taps = p3; break;
        case CHAP_CONV_K1: g.oD = g.iD; g.oH = g.iH; g.oW = g.iW; g.taps = 1; break;
        case CHAP_CONV_DOWN2:
            CHAP_REQUIRE(g.iH % 2 == 0 && g.iW % 2 == 0 && (d->nd == 2 || g.iD % 2 == 0), CHAP_ERR_BAD_ARG,
                         "k2s2 conv needs even input size");
            g.oD = d->nd == 3 ? g.iD / 2 : 1; g.oH = g.iH / 2; g.oW = g.iW / 2; g.taps = p2; break;
        case CHAP_CONV_UP2:
            g.oD = d->nd == 3 ? g.iD * 2 : 1; g.oH = g.iH * 2; g.oW = g.iW * 2; g.taps = p2; break;
        default: return fail(CHAP_ERR_BAD_ARG, "unknown conv kind %d", d->kind);
    }
    g.in_rows = (int64_t)g.n * g.iD * g.iH * g.iW;
    g.out_rows = (int64_t)g.n * g.oD * g.oH * g.oW;
    return CHAP_OK;
}

}  // namespace chap
