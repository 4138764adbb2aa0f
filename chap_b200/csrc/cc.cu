// Largest-connected-component filter on the device.
//
// Replaces get_ACDC_2DLargestCC of the reference (code/train_ours_2D.py:123-144): per sample and per foreground
// class keep the largest connected component (skimage.measure.label default = full connectivity: 8-neighbourhood in
// 2D, 26 in 3D); among components of equal size the one labelled first wins (np.argmax over np.bincount), i.e. the
// component whose first voxel comes first in scan order.  The reference does this on the host with one D2H/H2D
// round trip and 72 skimage calls per iteration; here it is a lock-free union-find on the label map itself:
//   1. parent[v] = first voxel of v's horizontal run inside its warp (ballot, no atomics)
//   2. run boundaries unite with the SAME-class neighbours of the previous row / plane (and across warp boundaries);
//      roots are always linked towards the smaller index, so a component's root is its first voxel in scan order
//   3. size[root] += 1 for every foreground voxel; best[n][c] = max over roots of (size << 32 | ~root)
//      -> largest size, ties to the smallest root = first labelled component
//   4. out[v] = c if root(v) == best root of (n, c) else 0
// Classes are mutually exclusive per voxel, so ONE union-find pass handles all classes of all samples.
#include "common.cuh"

namespace chap {

__device__ __forceinline__ int uf_find(int* parent, int i) {
    while (true) {
        int p = *reinterpret_cast<volatile int*>(parent + i);
        if (p == i) return i;
        int gp = *reinterpret_cast<volatile int*>(parent + p);
        if (gp != p) parent[i] = gp;               // path halving (benign race)
        i = p;
    }
}
// read-only find (no path halving): for kernels that must not disturb what other threads wrote
__device__ __forceinline__ int uf_find_ro(const int* parent, int i) {
    while (true) {
        const int p = __ldcg(parent + i);
        if (p == i) return i;
        i = p;
    }
}
__device__ __forceinline__ void uf_unite(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a > b) { int t = a; a = b; b = t; }     // a < b: hang b under a
        int old = atomicMin(parent + b, a);
        if (old == b) return;
        b = old;
    }
}

// Step 1+: parent[v] = first voxel of v's run inside its warp (32 consecutive voxels of a row): horizontal connectivity is
// resolved by a ballot instead of atomics, so the union phase only touches run boundaries.
__global__ void __launch_bounds__(256)
cc_init_kernel(const int64_t* __restrict__ seg, int* parent, int* size, unsigned long long* best, int W, int64_t total, int nbest) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (total + stride - 1) / stride;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t r = 0; r < rounds; ++r, i += stride) {            // whole warps iterate together (ballot)
        const bool in = i < total;
        const int64_t c = in ? seg[i] : 0;
        const bool left_same = in && c != 0 && lane > 0 && (i % W) != 0 && seg[i - 1] == c;
        const unsigned conn = __ballot_sync(0xffffffffu, left_same);
        if (in) {
            const unsigned starts = ~conn & (0xffffffffu >> (31 - lane));      // run starts at or before this lane (bit 0 always set)
            const int s = 31 - __clz(starts);
            parent[i] = (int)i - (lane - s);
            size[i] = 0;
            if (i < nbest) best[i] = 0ull;
        }
    }
}

// Step 2: unions with the previous row / plane, only where the connection is not already implied by a horizontal neighbour:
//   * up (x, y-1) unless the left voxel is in my run AND up-left is the same class (then the left voxel carries the link);
//   * the diagonals only when up is a different class, and only from the end of the run that touches them.
__global__ void __launch_bounds__(256)
cc_unite_kernel(const int64_t* __restrict__ seg, int* parent, int nd, int D, int H, int W, int64_t total) {
    pdl_enter();
    const int64_t vol = (int64_t)D * H * W;
    const int lane = threadIdx.x & 31;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = seg[i];
        if (c == 0) continue;
        const int64_t r = i % vol;
        const int x = (int)(r % W), y = (int)((r / W) % H), z = (int)(r / ((int64_t)W * H));
        const bool left_same = x > 0 && seg[i - 1] == c;
        const bool right_same = x < W - 1 && seg[i + 1] == c;
        if (left_same && lane == 0) uf_unite(parent, (int)i, (int)(i - 1));      // runs that continue across a warp boundary
        if (y > 0) {
            const int64_t u = i - W;
            const bool up = seg[u] == c, ul = x > 0 && seg[u - 1] == c, ur = x < W - 1 && seg[u + 1] == c;
            if (up) { if (!(left_same && ul)) uf_unite(parent, (int)i, (int)u); }
            else {
                if (ul && !left_same) uf_unite(parent, (int)i, (int)(u - 1));
                if (ur && !right_same) uf_unite(parent, (int)i, (int)(u + 1));
            }
        }
        if (nd == 3 && z > 0) {
            const int64_t b = i - (int64_t)W * H;
            const bool below = seg[b] == c;
            if (below) { if (!(left_same && seg[b - 1] == c)) uf_unite(parent, (int)i, (int)b); }
            else {
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        if (dx == 0 && dy == 0) continue;
                        const int xx = x + dx, yy = y + dy;
                        if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
                        const int64_t j = b + (int64_t)dy * W + dx;
                        if (seg[j] == c) uf_unite(parent, (int)i, (int)j);
                    }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
cc_count_kernel(const int64_t* __restrict__ seg, int* parent, int* size, int64_t total) {
    pdl_enter();
    // Warp-aggregated: the lanes of a warp that found the same root add their count with ONE atomic (a blob's voxels share one
    // root: per-voxel atomics on that single address serialised at ~1 ns each and dominated the 3D filter).
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (total + stride - 1) / stride;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t r = 0; r < rounds; ++r, i += stride) {            // whole warps iterate together (match_any)
        const bool fg = i < total && seg[i] != 0;
        int root = -1;
        // parent[i] = root is only a shortcut for the kernels that follow: another thread's path halving (parent[i] = its
        // grandparent, read before this store) may still land AFTER it and leave an intermediate ancestor here.  Round 1's
        // cc_write_kernel compared parent[i] with the winning root directly and lost a few voxels per launch that way
        // (found by tests/test_gpu_stress.py: 40 launches, 40 different outputs); it now walks to the root itself.
        if (fg) { root = uf_find(parent, (int)i); parent[i] = root; }
        const unsigned same = __match_any_sync(0xffffffffu, root);
        if (fg && lane == __ffs(same) - 1) atomicAdd(size + root, __popc(same));
    }
}

__global__ void __launch_bounds__(256)
cc_best_kernel(const int64_t* __restrict__ seg, const int* __restrict__ parent, const int* __restrict__ size,
               unsigned long long* best, int64_t vol, int n_classes, int64_t total) {
    pdl_enter();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = seg[i];
        if (c == 0 || parent[i] != (int)i) continue;               // roots only
        const unsigned long long key = ((unsigned long long)(unsigned)size[i] << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
        atomicMax(best + (i / vol) * n_classes + c, key);
    }
}

__global__ void __launch_bounds__(256)
cc_write_kernel(const int64_t* __restrict__ seg, const int* __restrict__ parent, const unsigned long long* __restrict__ best,
                int64_t vol, int n_classes, int64_t total, float* __restrict__ out) {
    pdl_enter();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = seg[i];
        float v = 0.f;
        if (c != 0) {
            const unsigned long long key = best[(i / vol) * n_classes + c];
            const unsigned root = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
            if ((unsigned)uf_find_ro(parent, (int)i) == root) v = (float)c;      // chains are 1-2 links long after cc_count_kernel
        }
        out[i] = v;
    }
}

}  // namespace chap

using namespace chap;

extern "C" size_t chap_largest_cc_workspace_bytes(int32_t n, int32_t d, int32_t h, int32_t w, int32_t n_classes) {
    const size_t total = (size_t)n * d * h * w;
    return total * 2 * sizeof(int) + (size_t)n * n_classes * sizeof(unsigned long long) + 16;
}

extern "C" int chap_largest_cc(const int64_t* seg, int32_t nd, int32_t n, int32_t d, int32_t h, int32_t w, int32_t n_classes,
                               float* out, void* workspace, size_t workspace_bytes, void* stream) {
    KernelTimer timer_("largest_cc", 0.0, 0.0, S(stream));
    CHAP_REQUIRE(seg && out && workspace && (nd == 2 || nd == 3) && n > 0 && d > 0 && h > 0 && w > 0 && n_classes >= 2,
                 CHAP_ERR_BAD_ARG, "largest_cc: bad argument");
    CHAP_REQUIRE(nd == 3 || d == 1, CHAP_ERR_BAD_ARG, "largest_cc: 2D needs d == 1");
    const int64_t vol = (int64_t)d * h * w, total = (int64_t)n * vol;
    CHAP_REQUIRE(total < 0x7FFFFFFFll, CHAP_ERR_BAD_ARG, "largest_cc: too many voxels for 32-bit union-find");
    CHAP_REQUIRE(workspace_bytes >= chap_largest_cc_workspace_bytes(n, d, h, w, n_classes), CHAP_ERR_WORKSPACE, "largest_cc: workspace too small");
    unsigned long long* best = reinterpret_cast<unsigned long long*>(workspace);          // 8-byte aligned first
    int* parent = reinterpret_cast<int*>(best + (size_t)n * n_classes);
    int* size = parent + total;
    cudaStream_t st = S(stream);
    const int grid = grid_for(total, 256 * 2);
    launch_k(cc_init_kernel, grid, 256, 0, st, seg, parent, size, best, w, total, n * n_classes);
    CHAP_TRY(launched("cc_init_kernel"));
    launch_k(cc_unite_kernel, grid, 256, 0, st, seg, parent, nd, d, h, w, total);
    CHAP_TRY(launched("cc_unite_kernel"));
    launch_k(cc_count_kernel, grid, 256, 0, st, seg, parent, size, total);
    CHAP_TRY(launched("cc_count_kernel"));
    launch_k(cc_best_kernel, grid, 256, 0, st, seg, parent, size, best, vol, n_classes, total);
    CHAP_TRY(launched("cc_best_kernel"));
    launch_k(cc_write_kernel, grid, 256, 0, st, seg, parent, best, vol, n_classes, total, out);
    return launched("cc_write_kernel");
}
