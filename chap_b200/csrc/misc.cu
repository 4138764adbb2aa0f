// Library state (errors, launch counter), fused SGD-momentum, sliding-window aggregation.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>
#include <map>
#include <mutex>
#include <string>

namespace chap {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_force_simt{0};
static int pdl_default() { const char* e = getenv("CHAP_PDL"); return (e && e[0] == '1') ? 1 : 0; }
std::atomic<int> g_pdl{pdl_default()};
std::atomic<int> g_precise_max_c{getenv("CHAP_PRECISE_MAX_C") ? atoi(getenv("CHAP_PRECISE_MAX_C")) : 0};

// ------------------------------------------------------------------ live kernel timing
namespace {
constexpr int kMaxTimed = 1 << 15;
struct TimedLaunch { cudaEvent_t a, b; const char* name; double flops, bytes; };
TimedLaunch* g_timed = nullptr;
std::atomic<int> g_timed_n{0};
std::atomic<int> g_timing_on{0};
}
KernelTimer::KernelTimer(const char* name, double flops, double bytes, cudaStream_t stream) : slot(-1), st(stream) {
    if (!g_timing_on.load(std::memory_order_relaxed) || !g_timed) return;
    int i = g_timed_n.fetch_add(1);
    if (i >= kMaxTimed) return;
    TimedLaunch& t = g_timed[i];
    if (!t.a) { cudaEventCreate(&t.a); cudaEventCreate(&t.b); }
    t.name = name; t.flops = flops; t.bytes = bytes;
    cudaEventRecord(t.a, st);
    slot = i;
}
// CHAP_TIMING_DETAIL=1: per-shape timer names ("conv_tc_fwd:k16n16:256x256x1"), interned for the life of the process
const char* timer_name(const char* family, int taps, int k, int n, int w, int h, int d, int64_t rows) {
    static const bool detail = getenv("CHAP_TIMING_DETAIL") != nullptr;
    if (!detail || !g_timing_on.load(std::memory_order_relaxed)) return family;
    static std::mutex mu;
    static std::map<std::string, std::string*> names;
    char buf[160];
    snprintf(buf, sizeof buf, "%s:t%d:k%d:n%d:%dx%dx%d:r%lld", family, taps, k, n, w, h, d, (long long)rows);
    std::lock_guard<std::mutex> lk(mu);
    auto it = names.find(buf);
    if (it == names.end()) it = names.emplace(buf, new std::string(buf)).first;
    return it->second->c_str();
}
KernelTimer::~KernelTimer() {
    if (slot >= 0) cudaEventRecord(g_timed[slot].b, st);
}

// ------------------------------------------------------------------ zero fill
__global__ void __launch_bounds__(256) zero_kernel(uint32_t* __restrict__ p, int64_t n4, int64_t n) {
    pdl_enter();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = i0; i < n4; i += stride) reinterpret_cast<uint4*>(p)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int64_t i = n4 * 4 + i0; i < n; i += stride) p[i] = 0u;
}
int zero_async(void* ptr, size_t bytes, cudaStream_t st) {
    static const bool use_memset = getenv("CHAP_ZERO_MEMSET") != nullptr;
    if (bytes == 0) return CHAP_OK;
    static const bool trace = getenv("CHAP_ZERO_TRACE") != nullptr;      // developer aid: which callers still need a zero-fill launch
    if (trace) fprintf(stderr, "zero_async %zu\n", bytes);
    if (use_memset || (bytes & 3u) || (reinterpret_cast<uintptr_t>(ptr) & 3u)) { CHAP_CUDA(cudaMemsetAsync(ptr, 0, bytes, st)); return CHAP_OK; }
    const int64_t n = (int64_t)(bytes >> 2);
    const bool vec = (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0;
    const int64_t n4 = vec ? n / 4 : 0;
    launch_k(zero_kernel, grid_for(vec ? n4 + 1 : n, 256 * 4, kNumSMs * 4), 256, 0, st, reinterpret_cast<uint32_t*>(ptr), n4, n);
    return launched("zero_kernel");
}

// ------------------------------------------------------------------ SGD momentum on a flat arena
__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, int64_t n4, int64_t n,
           float lr, const float* __restrict__ lr_dev, float mom, float wd, float gs, int first) {
    pdl_enter();
    if (lr_dev) lr = *lr_dev;          // learning rate read from device memory: the launch is CUDA-graph replayable
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pv = reinterpret_cast<float4*>(p)[i];
        float4 gv = ldg_stream(reinterpret_cast<const float4*>(g) + i);
        float4 bv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(buf)[i];
        float g0 = fmaf(wd, pv.x, gs * gv.x), g1 = fmaf(wd, pv.y, gs * gv.y);
        float g2 = fmaf(wd, pv.z, gs * gv.z), g3 = fmaf(wd, pv.w, gs * gv.w);
        bv.x = first ? g0 : fmaf(mom, bv.x, g0); bv.y = first ? g1 : fmaf(mom, bv.y, g1);
        bv.z = first ? g2 : fmaf(mom, bv.z, g2); bv.w = first ? g3 : fmaf(mom, bv.w, g3);
        pv.x -= lr * bv.x; pv.y -= lr * bv.y; pv.z -= lr * bv.z; pv.w -= lr * bv.w;
        reinterpret_cast<float4*>(buf)[i] = bv;
        reinterpret_cast<float4*>(p)[i] = pv;
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float gg = fmaf(wd, p[i], gs * g[i]);
        float b = first ? gg : fmaf(mom, buf[i], gg);
        buf[i] = b; p[i] -= lr * b;
    }
}

// ------------------------------------------------------------------ per-iteration schedule on the device
// Poly learning rate (code/train_ours_2D.py:387) and consistency weight (:34-36,356) from an iteration counter that lives in
// device memory and is advanced here: nothing per-iteration is written from the host, so a replayed CUDA graph cannot race
// with a host that runs iterations ahead.  One thread; double math (the host oracle computes these in Python floats).
__global__ void schedule_kernel(long long* iter, double base_lr, double max_it, double consistency, double rampup, long long ramp_div,
                                float* lr, float* cw) {
    pdl_enter();
    const long long it = *iter;
    double frac = 1.0 - (double)it / max_it;
    if (frac < 0.0) frac = 0.0;
    *lr = (float)(base_lr * pow(frac, 0.9));
    double w = 1.0;
    if (rampup != 0.0) {
        double cur = (double)(it / ramp_div);
        cur = cur < 0.0 ? 0.0 : (cur > rampup ? rampup : cur);
        const double ph = 1.0 - cur / rampup;
        w = exp(-5.0 * ph * ph);
    }
    *cw = (float)(consistency * w);
    *iter = it + 1;
}

// ------------------------------------------------------------------ validation helpers (code/val_2D.py:57-92)
// Nearest-neighbour resampling of a stack of slices through per-axis index tables: out[s, Y, X] = in[s, iy[Y], ix[X]].  The
// tables are scipy.ndimage.zoom(order=0)'s own index map (floor(o * (in - 1) / (out - 1) + 0.5) in double), computed on the
// host, so the label maps are bit-identical to the reference's zoom -> net -> zoom back -- including scipy's quirk that an output
// coordinate which overshoots the last sample by one ulp (o * ratio > in - 1 in double) takes the constant fill value 0.
template <typename T>
__global__ void __launch_bounds__(256)
gather2d_kernel(const T* __restrict__ in, const int* __restrict__ iy, const int* __restrict__ ix, int h, int w, int H, int W,
                int64_t total, T* __restrict__ out) {
    pdl_enter();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int X = (int)(i % W); const int Y = (int)((i / W) % H); const int64_t s = i / ((int64_t)W * H);
        const int sy = iy[Y], sx = ix[X];                 // -1: scipy's mode='constant' fill (coordinate beyond the last sample)
        out[i] = (sy < 0 || sx < 0) ? T(0) : in[(s * h + sy) * w + sx];
    }
}
// per-class overlap counts of two label volumes: counts[c] = {|pred == c & gt == c|, |pred == c|, |gt == c|} (Dice numerators /
// denominators of code/val_2D.py:43-51 for every class in one pass); block-level shared-memory histogram, then one atomic per bin
__global__ void __launch_bounds__(256)
label_overlap_kernel(const int64_t* __restrict__ pred, const int64_t* __restrict__ gt, int64_t elems, int classes,
                     unsigned long long* __restrict__ counts) {
    pdl_enter();
    __shared__ unsigned int h[3 * 16];
    for (int i = threadIdx.x; i < 3 * classes; i += blockDim.x) h[i] = 0u;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < elems; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = pred[i], g = gt[i];
        if (p >= 0 && p < classes) atomicAdd(&h[3 * p + 1], 1u);
        if (g >= 0 && g < classes) { atomicAdd(&h[3 * g + 2], 1u); if (p == g) atomicAdd(&h[3 * g], 1u); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * classes; i += blockDim.x) if (h[i]) atomicAdd(&counts[i], (unsigned long long)h[i]);
}

// ------------------------------------------------------------------ sliding window
__device__ __forceinline__ int win_start(int i, int stride, int vol, int patch) {
    int s = stride * i;
    int lim = vol - patch;
    return s < lim ? s : lim;
}

__global__ void __launch_bounds__(256)
sw_extract_kernel(chap_sw_desc d, const float* __restrict__ vol, int first, int64_t total, float* __restrict__ patches) {
    pdl_enter();
    const int64_t pvol = (int64_t)d.patch[0] * d.patch[1] * d.patch[2];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int wi = first + (int)(i / pvol);
        int64_t r = i % pvol;
        int z = (int)(r % d.patch[2]); r /= d.patch[2]; int y = (int)(r % d.patch[1]); int x = (int)(r / d.patch[1]);
        int wz = wi % d.nwin[2], wy = (wi / d.nwin[2]) % d.nwin[1], wx = wi / (d.nwin[2] * d.nwin[1]);
        int xs = win_start(wx, d.stride[0], d.vol[0], d.patch[0]);
        int ys = win_start(wy, d.stride[1], d.vol[1], d.patch[1]);
        int zs = win_start(wz, d.stride[2], d.vol[2], d.patch[2]);
        patches[i] = vol[((int64_t)(xs + x) * d.vol[1] + ys + y) * d.vol[2] + zs + z];
    }
}

// Tile-owned, atomic-free accumulate: one thread owns one output voxel and visits the windows that
// cover it in the reference's x -> y -> z loop order, so the fp32 sums are bit-identical to the
// host `score_map[...] += y` of code/test_3D_util.py:67-70; then score/cnt and first-max argmax.
template <int C>
__global__ void __launch_bounds__(256)
sw_aggregate_kernel(chap_sw_desc d, const float* __restrict__ win, int is_prob, float* __restrict__ score,
                    float* __restrict__ cnt, int64_t* __restrict__ label) {
    pdl_enter();
    const int64_t nvox = (int64_t)d.vol[0] * d.vol[1] * d.vol[2];
    const int64_t pvol = (int64_t)d.patch[0] * d.patch[1] * d.patch[2];
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (int64_t)gridDim.x * blockDim.x) {
        int z = (int)(v % d.vol[2]); int y = (int)((v / d.vol[2]) % d.vol[1]); int x = (int)(v / ((int64_t)d.vol[2] * d.vol[1]));
        float acc[C];
#pragma unroll
        for (int k = 0; k < C; ++k) acc[k] = 0.f;
        float n = 0.f;
        for (int wx = 0; wx < d.nwin[0]; ++wx) {
            int xs = win_start(wx, d.stride[0], d.vol[0], d.patch[0]);
            if (x < xs || x >= xs + d.patch[0]) continue;
            for (int wy = 0; wy < d.nwin[1]; ++wy) {
                int ys = win_start(wy, d.stride[1], d.vol[1], d.patch[1]);
                if (y < ys || y >= ys + d.patch[1]) continue;
                for (int wz = 0; wz < d.nwin[2]; ++wz) {
                    int zs = win_start(wz, d.stride[2], d.vol[2], d.patch[2]);
                    if (z < zs || z >= zs + d.patch[2]) continue;
                    int64_t wi = ((int64_t)wx * d.nwin[1] + wy) * d.nwin[2] + wz;
                    int64_t e = wi * pvol + ((int64_t)(x - xs) * d.patch[1] + (y - ys)) * d.patch[2] + (z - zs);
                    float p[C];
                    if (C == 2) { float2 t = __ldg(reinterpret_cast<const float2*>(win) + e); p[0] = t.x; p[1 % C] = t.y; }
                    else if (C == 4) { float4 t = __ldg(reinterpret_cast<const float4*>(win) + e); p[0] = t.x; p[1 % C] = t.y; p[2 % C] = t.z; p[3 % C] = t.w; }
                    else {
#pragma unroll
                        for (int k = 0; k < C; ++k) p[k] = __ldg(win + e * C + k);
                    }
                    if (!is_prob) {
                        float m = p[0];
#pragma unroll
                        for (int k = 1; k < C; ++k) m = fmaxf(m, p[k]);
                        float s = 0.f;
#pragma unroll
                        for (int k = 0; k < C; ++k) { p[k] = expf(p[k] - m); s += p[k]; }
                        float inv = 1.f / s;
#pragma unroll
                        for (int k = 0; k < C; ++k) p[k] *= inv;
                    }
#pragma unroll
                    for (int k = 0; k < C; ++k) acc[k] = acc[k] + p[k];
                    n = n + 1.f;
                }
            }
        }
        int best = 0; float bm = 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            float q = acc[k] / n;                       // division BEFORE argmax (test_3D_util.py:71-72)
            if (score) score[(int64_t)k * nvox + v] = q;
            if (k == 0 || q > bm) { bm = q; best = k; }
        }
        if (cnt) cnt[v] = n;
        label[v] = best;
    }
}

}  // namespace chap

using namespace chap;

extern "C" const char* chap_last_error(void) { return g_err; }
extern "C" int chap_abi_version(void) { return CHAP_ABI_VERSION; }
extern "C" uint64_t chap_launch_count(void) { return g_launches.load(); }
extern "C" void chap_reset_launch_count(void) { g_launches.store(0); }
extern "C" void chap_set_force_simt(int flag) { g_force_simt.store(flag ? 1 : 0); }
extern "C" int chap_get_force_simt(void) { return g_force_simt.load(); }
extern "C" void chap_set_pdl(int flag) { g_pdl.store(flag ? 1 : 0); }
extern "C" int chap_get_pdl(void) { return g_pdl.load(); }
extern "C" void chap_set_conv_precision(int max_channels) { g_precise_max_c.store(max_channels < 0 ? 0 : max_channels); }
extern "C" int chap_get_conv_precision(void) { return g_precise_max_c.load(); }

extern "C" void chap_timing_enable(int on) {
    if (on && !g_timed) g_timed = new TimedLaunch[kMaxTimed]();
    if (on) g_timed_n.store(0);
    g_timing_on.store(on ? 1 : 0);
}

// Synchronises the device and writes one line per kernel family:
//   name launches total_ms flops bytes\n     (flops / bytes are sums of the algorithmic figures passed at launch)
extern "C" int chap_timing_report(char* buf, size_t cap) {
    CHAP_REQUIRE(buf && cap > 0, CHAP_ERR_BAD_ARG, "timing_report: bad buffer");
    buf[0] = 0;
    if (!g_timed) return CHAP_OK;
    CHAP_CUDA(cudaDeviceSynchronize());
    int n = g_timed_n.load();
    if (n > kMaxTimed) n = kMaxTimed;
    struct Agg { const char* name; int count; double ms, flops, bytes; };
    constexpr int kMaxAgg = 512;
    static Agg agg[kMaxAgg]; int na = 0;
    for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_timed[i].a, g_timed[i].b) != cudaSuccess) { cudaGetLastError(); continue; }
        int j = 0;
        for (; j < na; ++j) if (agg[j].name == g_timed[i].name || strcmp(agg[j].name, g_timed[i].name) == 0) break;
        if (j == na) { if (na == kMaxAgg) continue; agg[na++] = Agg{g_timed[i].name, 0, 0.0, 0.0, 0.0}; }
        agg[j].count++; agg[j].ms += ms; agg[j].flops += g_timed[i].flops; agg[j].bytes += g_timed[i].bytes;
    }
    size_t off = 0;
    for (int j = 0; j < na; ++j) {
        int w = snprintf(buf + off, cap - off, "%s %d %.6f %.6e %.6e\n", agg[j].name, agg[j].count, agg[j].ms, agg[j].flops, agg[j].bytes);
        if (w < 0 || (size_t)w >= cap - off) break;
        off += (size_t)w;
    }
    return CHAP_OK;
}

extern "C" int chap_check_device(void) {
    int dev = 0;
    CHAP_CUDA(cudaGetDevice(&dev));
    int major = 0, minor = 0;
    CHAP_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    CHAP_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    CHAP_REQUIRE(major == 10, CHAP_ERR_ARCH, "libchap_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return CHAP_OK;
}

extern "C" int chap_sgd_momentum(float* p, const float* g, float* buf, int64_t elems, float lr, float momentum,
                                 float weight_decay, float grad_scale, int32_t first_step, void* stream) {
    KernelTimer timer_("sgd_momentum", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(p && g && buf && elems > 0, CHAP_ERR_BAD_ARG, "sgd_momentum: bad argument");
    CHAP_REQUIRE(aligned16(p) && aligned16(g) && aligned16(buf), CHAP_ERR_ALIGNMENT, "sgd_momentum: buffers must be 16-byte aligned");
    launch_k(sgd_kernel, grid_for(elems / 4 + 1, 256 * 2), 256, 0, S(stream), p, g, buf, elems / 4, elems, lr, nullptr, momentum, weight_decay, grad_scale, first_step);
    return launched("sgd_kernel");
}

extern "C" int chap_sgd_momentum_lrdev(float* p, const float* g, float* buf, int64_t elems, const float* lr_dev, float momentum,
                                       float weight_decay, float grad_scale, void* stream) {
    KernelTimer timer_("sgd_momentum_lrdev", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(p && g && buf && lr_dev && elems > 0, CHAP_ERR_BAD_ARG, "sgd_momentum_lrdev: bad argument");
    CHAP_REQUIRE(aligned16(p) && aligned16(g) && aligned16(buf), CHAP_ERR_ALIGNMENT, "sgd_momentum_lrdev: buffers must be 16-byte aligned");
    launch_k(sgd_kernel, grid_for(elems / 4 + 1, 256 * 2), 256, 0, S(stream), p, g, buf, elems / 4, elems, 0.f, lr_dev, momentum, weight_decay, grad_scale, 0);
    return launched("sgd_kernel");
}

extern "C" int chap_schedule_step(int64_t* iter_dev, double base_lr, double max_iterations, double consistency, double rampup,
                                  int64_t ramp_div, float* lr_dev, float* cw_dev, void* stream) {
    CHAP_REQUIRE(iter_dev && lr_dev && cw_dev && max_iterations > 0 && ramp_div > 0, CHAP_ERR_BAD_ARG, "schedule_step: bad argument");
    launch_k(schedule_kernel, 1, 1, 0, S(stream), reinterpret_cast<long long*>(iter_dev), base_lr, max_iterations, consistency, rampup,
                                            (long long)ramp_div, lr_dev, cw_dev);
    return launched("schedule_kernel");
}

extern "C" int chap_gather2d(const void* in, int32_t elem_bytes, const int32_t* iy, const int32_t* ix, int32_t n, int32_t h, int32_t w,
                             int32_t out_h, int32_t out_w, void* out, void* stream) {
    CHAP_REQUIRE(in && iy && ix && out && n > 0 && h > 0 && w > 0 && out_h > 0 && out_w > 0, CHAP_ERR_BAD_ARG, "gather2d: bad argument");
    CHAP_REQUIRE(elem_bytes == 4 || elem_bytes == 8, CHAP_ERR_BAD_ARG, "gather2d: element size must be 4 or 8 bytes (got %d)", elem_bytes);
    const int64_t total = (int64_t)n * out_h * out_w;
    KernelTimer timer_("gather2d", 0.0, 2.0 * elem_bytes * (double)total, S(stream));
    if (elem_bytes == 4) launch_k(gather2d_kernel<float>, grid_for(total, 256 * 4), 256, 0, S(stream), (const float*)in, iy, ix, h, w, out_h, out_w, total, (float*)out);
    else launch_k(gather2d_kernel<int64_t>, grid_for(total, 256 * 4), 256, 0, S(stream), (const int64_t*)in, iy, ix, h, w, out_h, out_w, total, (int64_t*)out);
    return launched("gather2d_kernel");
}

extern "C" int chap_label_overlap(const int64_t* pred, const int64_t* gt, int64_t elems, int32_t classes, uint64_t* counts, void* stream) {
    CHAP_REQUIRE(pred && gt && counts && elems > 0 && classes >= 1 && classes <= 16, CHAP_ERR_BAD_ARG, "label_overlap: bad argument (classes %d)", classes);
    KernelTimer timer_("label_overlap", 0.0, 16.0 * (double)elems, S(stream));
    CHAP_TRY(zero_async(counts, (size_t)3 * classes * sizeof(uint64_t), S(stream)));
    launch_k(label_overlap_kernel, grid_for(elems, 256 * 8), 256, 0, S(stream), pred, gt, elems, classes, reinterpret_cast<unsigned long long*>(counts));
    return launched("label_overlap_kernel");
}

static int check_sw(const chap_sw_desc* d) {
    CHAP_REQUIRE(d != nullptr, CHAP_ERR_BAD_ARG, "sliding window desc is NULL");
    for (int a = 0; a < 3; ++a) {
        CHAP_REQUIRE(d->vol[a] >= d->patch[a] && d->patch[a] > 0 && d->nwin[a] > 0 && d->stride[a] > 0, CHAP_ERR_BAD_ARG,
                     "sliding window axis %d: vol %d patch %d nwin %d stride %d", a, d->vol[a], d->patch[a], d->nwin[a], d->stride[a]);
    }
    CHAP_REQUIRE(d->c >= 1, CHAP_ERR_BAD_ARG, "sliding window: classes %d", d->c);
    return CHAP_OK;
}

extern "C" int chap_sw_extract(const chap_sw_desc* d, const float* volume, int32_t first, int32_t count, float* patches, void* stream) {
    CHAP_TRY(check_sw(d));
    const int total_win = d->nwin[0] * d->nwin[1] * d->nwin[2];
    CHAP_REQUIRE(volume && patches && first >= 0 && count > 0 && first + count <= total_win, CHAP_ERR_BAD_ARG, "sw_extract: bad window range");
    const int64_t total = (int64_t)count * d->patch[0] * d->patch[1] * d->patch[2];
    launch_k(sw_extract_kernel, grid_for(total, 256 * 4), 256, 0, S(stream), *d, volume, first, total, patches);
    return launched("sw_extract_kernel");
}

extern "C" int chap_sw_aggregate(const chap_sw_desc* d, const float* win, int32_t is_prob, float* score, float* cnt,
                                 int64_t* label, void* stream) {
    CHAP_TRY(check_sw(d));
    CHAP_REQUIRE(win && label, CHAP_ERR_BAD_ARG, "sw_aggregate: NULL pointer");
    const int64_t nvox = (int64_t)d->vol[0] * d->vol[1] * d->vol[2];
    const double n_win = (double)d->nwin[0] * d->nwin[1] * d->nwin[2];
    // algorithmic bytes: every window's logits once + score (C floats), cnt, label (8 B) per voxel once
    KernelTimer timer_("sw_aggregate", 0.0, 4.0 * n_win * d->patch[0] * d->patch[1] * d->patch[2] * d->c +
                                                (double)nvox * ((score ? 4.0 * d->c : 0.0) + (cnt ? 4.0 : 0.0) + 8.0), S(stream));
    int grid = grid_for(nvox, 256);
    switch (d->c) {
        case 1: launch_k(sw_aggregate_kernel<1>, grid, 256, 0, S(stream), *d, win, is_prob, score, cnt, label); break;   // the reference's default num_classes=1
        case 2: CHAP_REQUIRE(((uintptr_t)win & 7u) == 0, CHAP_ERR_ALIGNMENT, "sw_aggregate: misaligned");
                launch_k(sw_aggregate_kernel<2>, grid, 256, 0, S(stream), *d, win, is_prob, score, cnt, label); break;
        case 3: launch_k(sw_aggregate_kernel<3>, grid, 256, 0, S(stream), *d, win, is_prob, score, cnt, label); break;
        case 4: CHAP_REQUIRE(aligned16(win), CHAP_ERR_ALIGNMENT, "sw_aggregate: misaligned");
                launch_k(sw_aggregate_kernel<4>, grid, 256, 0, S(stream), *d, win, is_prob, score, cnt, label); break;
        default: return fail(CHAP_ERR_BAD_ARG, "sw_aggregate: unsupported class count %d", d->c);
    }
    return launched("sw_aggregate_kernel");
}
