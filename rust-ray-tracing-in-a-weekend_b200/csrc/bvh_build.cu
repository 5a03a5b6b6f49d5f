// bvh_build.cu — the BVH built ON THE DEVICE: Morton LBVH -> bottom-up refit -> collapse to the 8-wide
// compressed layout -> primitive records written in leaf order.  The host only uploads what the constructors were given
// (40 bytes per rtw_sphere_batch sphere) instead of building and uploading nodes + records (128 bytes per sphere).
//
// Replaces the reference's builder new_bvh_node (src/hittable.rs:77-130: a recursive median split that clones the whole
// list at every node, :78) — and the host SAH builder of scene_host.cpp for scenes beyond the cache-resident range
// (16 M spheres: 5.1 s of host build + upload, measured round 2; here a few tens of ms of kernels).
//
// Steps (all on one stream, no host round trip except one count per collapse level):
//   prim_boxes      f32 boxes rounded outward (bulk spheres from their f64 centre / radius; other primitives arrive with
//                   the box the host flattener computed) + scene bounds
//   morton          63-bit Morton code of the box centre (21 bits per axis)
//   sort            cub::DeviceRadixSort (library code, outside the render hot path)
//   karras          binary radix tree over the sorted codes (Karras 2012), ties broken by position
//   refit           leaf boxes up to the root, second arrival at a node continues (atomic flags)
//   collapse        rtww::collapse_one per wide node, level by level (bvh_wide.h — the code the host path and the CPU
//                   tests run)
//   emit            DPrim records in final (leaf slot) order
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "bvh_build.hpp"

namespace rtwb {

namespace {

#define BCUDA(x)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (x);                                                            \
        if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); return e_ == cudaErrorMemoryAllocation ? -4 : -3; } \
    } while (0)

__device__ __forceinline__ float f_down(double v) { float f = __double2float_rd(v); return nextafterf(f, -INFINITY); }
__device__ __forceinline__ float f_up(double v) { float f = __double2float_ru(v); return nextafterf(f, INFINITY); }

// order-preserving float <-> uint (atomicMin / atomicMax on floats of either sign)
__device__ __forceinline__ unsigned f2o(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__host__ __device__ __forceinline__ float o2f(unsigned u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

// boxes[i] for every BVH primitive: [0, n_host) were flattened on the host (box given), the rest are bulk spheres
__global__ void prim_boxes_kernel(int n, int n_host, const float* __restrict__ host_boxes, const BulkSphereD* __restrict__ bulk,
                                  float* __restrict__ boxes, unsigned* __restrict__ bounds /* [6] ordered-uint min xyz, max xyz */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float b[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    if (i < n) {
        if (i < n_host) { for (int a = 0; a < 6; ++a) b[a] = host_boxes[(size_t)i * 6 + a]; }
        else {
            const BulkSphereD s = bulk[i - n_host];
            const double ar = fabs(s.r);
            b[0] = f_down(s.c[0] - ar); b[1] = f_down(s.c[1] - ar); b[2] = f_down(s.c[2] - ar);
            b[3] = f_up(s.c[0] + ar); b[4] = f_up(s.c[1] + ar); b[5] = f_up(s.c[2] + ar);
        }
        for (int a = 0; a < 6; ++a) boxes[(size_t)i * 6 + a] = b[a];
    }
    // scene bounds of the box CENTRES (Morton grid): warp reduce, then one atomic per warp and plane
    float c[3] = {0.5f * (b[0] + b[3]), 0.5f * (b[1] + b[4]), 0.5f * (b[2] + b[5])};
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) { lo[a] = i < n ? c[a] : INFINITY; hi[a] = i < n ? c[a] : -INFINITY; }
    for (int o = 16; o; o >>= 1)
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o)); hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o)); }
    if ((threadIdx.x & 31) == 0)
        for (int a = 0; a < 3; ++a) { atomicMin(bounds + a, f2o(lo[a])); atomicMax(bounds + 3 + a, f2o(hi[a])); }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {      // 21 bits -> every third bit
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void morton_kernel(int n, const float* __restrict__ boxes, const unsigned* __restrict__ bounds, unsigned long long* __restrict__ keys,
                              unsigned* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long code = 0;
    for (int a = 0; a < 3; ++a) {
        const float lo = o2f(bounds[a]), hi = o2f(bounds[3 + a]);
        const float c = 0.5f * (boxes[(size_t)i * 6 + a] + boxes[(size_t)i * 6 + 3 + a]);
        const float ext = hi - lo;
        float t = ext > 0.0f ? (c - lo) / ext : 0.0f;
        t = fminf(fmaxf(t, 0.0f), 1.0f);
        const unsigned long long q = (unsigned long long)fminf(t * 2097152.0f, 2097151.0f);
        code |= spread21(q) << a;
    }
    keys[i] = code; vals[i] = (unsigned)i;
}

// Karras 2012, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees": one thread per inner node.
// delta = length of the common prefix of two sorted keys; equal keys fall back to their positions (always distinct).
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long x = keys[i] ^ keys[j];
    return x ? __clzll((long long)x) : 64 + __clz(i ^ j);
}
__global__ void karras_kernel(int n, const unsigned long long* __restrict__ keys, int* __restrict__ left, int* __restrict__ right, int* __restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_inner = n - 1;
    if (i >= n_inner) return;
    const int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1) if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t <= 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lref = min(i, j) == gamma ? n_inner + gamma : gamma;
    const int rref = max(i, j) == gamma + 1 ? n_inner + gamma + 1 : gamma + 1;
    left[i] = lref; right[i] = rref;
    parent[lref] = i; parent[rref] = i;
    if (i == 0) parent[0] = -1;
}

// leaf boxes, then up: the second thread to reach an inner node finds both children complete
__global__ void refit_kernel(int n, const unsigned* __restrict__ sorted_idx, const float* __restrict__ prim_boxes, const int* __restrict__ left,
                             const int* __restrict__ right, const int* __restrict__ parent, float* __restrict__ box, int* __restrict__ flags,
                             int* __restrict__ count /* leaves under each inner node (the wide collapse absorbs small subtrees whole) */) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int n_inner = n - 1;
    const size_t src = (size_t)sorted_idx[j] * 6;
    float b[6];
    for (int a = 0; a < 6; ++a) { b[a] = prim_boxes[src + a]; box[(size_t)(n_inner + j) * 6 + a] = b[a]; }
    int cur = parent[n_inner + j];
    while (cur >= 0) {
        __threadfence();
        if (atomicAdd(flags + cur, 1) == 0) return;
        __threadfence();
        const volatile float* lb = box + (size_t)left[cur] * 6;
        const volatile float* rb = box + (size_t)right[cur] * 6;
        for (int a = 0; a < 3; ++a) { b[a] = fminf(lb[a], rb[a]); b[3 + a] = fmaxf(lb[3 + a], rb[3 + a]); }
        for (int a = 0; a < 6; ++a) box[(size_t)cur * 6 + a] = b[a];
        {
            const int l = left[cur], r = right[cur];
            count[cur] = (l < n_inner ? ((const volatile int*)count)[l] : 1) + (r < n_inner ? ((const volatile int*)count)[r] : 1);
        }
        cur = parent[cur];
    }
}

struct DevAlloc {
    int* c;            // [0] wide nodes, [1] leaf prims, [2] next queue, [3] max depth
    __device__ int nodes(int k) { return k ? atomicAdd(c, k) : 0; }
    __device__ int prims(int k) { return k ? atomicAdd(c + 1, k) : 0; }
    __device__ int queue(int k) { return k ? atomicAdd(c + 2, k) : 0; }
    __device__ void depth(int d) { atomicMax(c + 3, d); }
};
__global__ void collapse_kernel(rtww::B2View v, const rtww::WideItem* __restrict__ cur, int n_cur, DWNode* __restrict__ nodes, int* __restrict__ leaf_order,
                                rtww::WideItem* __restrict__ next, int* __restrict__ counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cur) return;
    DevAlloc alloc; alloc.c = counters;
    rtww::collapse_one(v, cur[i], nodes, leaf_order, next, alloc);
}

// final primitive records: position k holds the primitive of binary leaf leaf_order[k] = original index sorted_idx[...]
__global__ void emit_prims_kernel(int n, int n_host, const int* __restrict__ leaf_order, const unsigned* __restrict__ sorted_idx,
                                  const DPrim* __restrict__ host_prims, const BulkSphereD* __restrict__ bulk, DPrim* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int src = (int)sorted_idx[leaf_order[k]];
    uint4* dst = reinterpret_cast<uint4*>(out + k);
    if (src < n_host) {
        const uint4* s4 = reinterpret_cast<const uint4*>(host_prims + src);
        dst[0] = s4[0]; dst[1] = s4[1]; dst[2] = s4[2]; dst[3] = s4[3];
        return;
    }
    const BulkSphereD s = bulk[src - n_host];
    DPrim p;
    p.s.cx = s.c[0]; p.s.cy = s.c[1]; p.s.cz = s.c[2]; p.s.r = s.r;
    p.dcx = 0.f; p.dcy = 0.f; p.dcz = 0.f; p.t0 = 0.f;
    p.type = PRIM_SPHERE; p.mat = s.mat - 1; p.xform = 0; p.inv_dt = 0.f;
    const uint4* s4 = reinterpret_cast<const uint4*>(&p);
    dst[0] = s4[0]; dst[1] = s4[1]; dst[2] = s4[2]; dst[3] = s4[3];
}

// width 2: the radix tree itself as two-box nodes (rtw_types.h DNode); leaf j holds the single primitive at position j of
// the sorted order, so the primitive records are emitted in sorted order (leaf_order = identity)
__global__ void emit_binary_kernel(int n_inner, const int* __restrict__ left, const int* __restrict__ right, const float* __restrict__ box, DNode* __restrict__ nodes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inner) return;
    const int l = left[i], r = right[i];
    const float* lb = box + (size_t)l * 6; const float* rb = box + (size_t)r * 6;
    DNode d;
    d.c0minx = lb[0]; d.c0maxx = lb[3]; d.c0miny = lb[1]; d.c0maxy = lb[4]; d.c0minz = lb[2]; d.c0maxz = lb[5];
    d.c1minx = rb[0]; d.c1maxx = rb[3]; d.c1miny = rb[1]; d.c1maxy = rb[4]; d.c1minz = rb[2]; d.c1maxz = rb[5];
    d.child0 = l < n_inner ? l : ~((l - n_inner) << 3);
    d.child1 = r < n_inner ? r : ~((r - n_inner) << 3);
    d.pad0 = 0; d.pad1 = 0;
    nodes[i] = d;
}
__global__ void iota_kernel(int n, int* __restrict__ a) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = i; }
// depth of the binary tree = longest leaf-to-root parent chain (bounds the traversal stack of bvh_closest)
__global__ void binary_depth_kernel(int n, const int* __restrict__ parent, int* __restrict__ max_depth) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    int d = 0;
    if (j < n) for (int cur = parent[n - 1 + j]; cur >= 0; cur = parent[cur]) ++d;
    for (int o = 16; o; o >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_depth, d);
}

__global__ void single_leaf_root_kernel(const float* __restrict__ boxes, DWNode* __restrict__ nodes, int* __restrict__ leaf_order) {
    DWNode w; memset(&w, 0, sizeof(w));
    float p[3], step[3];
    w.ex = (uint8_t)rtww::wide_axis_grid(boxes[0], boxes[3], p[0], step[0]);
    w.ey = (uint8_t)rtww::wide_axis_grid(boxes[1], boxes[4], p[1], step[1]);
    w.ez = (uint8_t)rtww::wide_axis_grid(boxes[2], boxes[5], p[2], step[2]);
    w.px = p[0]; w.py = p[1]; w.pz = p[2];
    for (int s = 0; s < 8; ++s) { w.qlox[s] = w.qloy[s] = w.qloz[s] = 255; w.qhix[s] = w.qhiy[s] = w.qhiz[s] = 0; }
    w.qlox[0] = rtww::wide_qlo(boxes[0], p[0], step[0]); w.qhix[0] = rtww::wide_qhi(boxes[3], p[0], step[0]);
    w.qloy[0] = rtww::wide_qlo(boxes[1], p[1], step[1]); w.qhiy[0] = rtww::wide_qhi(boxes[4], p[1], step[1]);
    w.qloz[0] = rtww::wide_qlo(boxes[2], p[2], step[2]); w.qhiz[0] = rtww::wide_qhi(boxes[5], p[2], step[2]);
    w.lmask = 1;
    nodes[0] = w; leaf_order[0] = 0;
}

// ---- staged upload (SURVEY 8f-2 "commit overlap") ------------------------------------------------------------------
// The spheres of a big scene live in pageable host memory (the caller's constructors filled a std::vector).  One
// cudaMemcpyAsync from pageable memory is staged by the driver through one pinned buffer on the calling thread: 16 M
// spheres = 640 MB took 60-190 ms, most of the commit.  Here a few host threads each copy their chunks into their own
// pinned buffers (allocated once per process) and hand them to their own stream, so the host-side copy of one chunk
// overlaps the DMA of the others and several cores share the host-side copy.  Returns with the data on the device.
constexpr int kStageThreads = 4, kStageSlots = 2;
constexpr size_t kStageChunk = (size_t)8 << 20;
struct StagePool { std::mutex mu; bool tried = false, ok = false; void* buf[kStageThreads][kStageSlots] = {}; };
StagePool g_stage;

cudaError_t staged_upload(cudaStream_t st, void* dst, const void* src, size_t bytes) {
    const char* sw = getenv("RTW_STAGED_UPLOAD");
    std::unique_lock<std::mutex> lk(g_stage.mu, std::try_to_lock);      // a second concurrent commit takes the plain copy
    bool staged = bytes >= 4 * kStageChunk && lk.owns_lock() && !(sw && sw[0] == '0');
    if (staged && !g_stage.tried) {
        g_stage.tried = true; g_stage.ok = true;
        for (int t = 0; t < kStageThreads && g_stage.ok; ++t)
            for (int k = 0; k < kStageSlots && g_stage.ok; ++k)
                if (cudaHostAlloc(&g_stage.buf[t][k], kStageChunk, cudaHostAllocPortable) != cudaSuccess) { g_stage.ok = false; cudaGetLastError(); }
    }
    if (!staged || !g_stage.ok) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
    int dev = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess) return e0;
    const size_t n_chunks = (bytes + kStageChunk - 1) / kStageChunk;
    cudaError_t errs[kStageThreads];
    std::thread th[kStageThreads];
    for (int t = 0; t < kStageThreads; ++t) errs[t] = cudaSuccess;
    auto work = [&](int t) {
        cudaError_t e = cudaSetDevice(dev);
        cudaStream_t s = nullptr; cudaEvent_t ev[kStageSlots] = {};
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        for (int k = 0; k < kStageSlots && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
        size_t use = 0;
        for (size_t c = t; c < n_chunks && e == cudaSuccess; c += kStageThreads, ++use) {
            const int k = (int)(use % kStageSlots);
            if (use >= (size_t)kStageSlots) e = cudaEventSynchronize(ev[k]);      // the buffer's previous chunk has left
            if (e != cudaSuccess) break;
            const size_t off = c * kStageChunk, len = bytes - off < kStageChunk ? bytes - off : kStageChunk;
            memcpy(g_stage.buf[t][k], (const char*)src + off, len);
            e = cudaMemcpyAsync((char*)dst + off, g_stage.buf[t][k], len, cudaMemcpyHostToDevice, s);
            if (e == cudaSuccess) e = cudaEventRecord(ev[k], s);
        }
        if (s) { const cudaError_t e2 = cudaStreamSynchronize(s); if (e == cudaSuccess) e = e2; }
        for (int k = 0; k < kStageSlots; ++k) if (ev[k]) cudaEventDestroy(ev[k]);
        if (s) cudaStreamDestroy(s);
        errs[t] = e;
    };
    for (int t = 1; t < kStageThreads; ++t) {
        try { th[t] = std::thread(work, t); } catch (...) { work(t); }     // (no thread to be had: this one does the share; nothing unwinds across the C ABI)
    }
    work(0);
    for (int t = 1; t < kStageThreads; ++t) if (th[t].joinable()) th[t].join();
    (void)st;                                                           // every chunk has arrived: kernels launched from here on see the data
    for (int t = 0; t < kStageThreads; ++t) if (errs[t] != cudaSuccess) return errs[t];
    return cudaSuccess;
}

// ---- scratch pool ------------------------------------------------------------------------------------------------
// The builder's scratch (16 M spheres: 26 buffers, 4.2 GB) comes from a stream-ordered memory pool of its own, kept across
// commits and trimmed when a scene is freed (trim_scratch): measured on the 16 M-sphere commit, cudaFree of that scratch
// took 100 ms and cudaMalloc 10-60 ms of a 200-ms commit (profiles/r2_au2_commit_phases.log).  RTW_SCRATCH_POOL=0: plain
// cudaMalloc / cudaFree (also the fallback when the pool cannot be created).
constexpr int kMaxDev = 32;
struct ScratchPools { std::mutex mu; cudaMemPool_t pool[kMaxDev] = {}; bool tried[kMaxDev] = {}; };
ScratchPools g_pools;

cudaMemPool_t scratch_pool(int dev) {
    if (dev < 0 || dev >= kMaxDev) return nullptr;
    std::lock_guard<std::mutex> lk(g_pools.mu);
    if (!g_pools.tried[dev]) {
        g_pools.tried[dev] = true;
        const char* sw = getenv("RTW_SCRATCH_POOL");
        if (!(sw && sw[0] == '0')) {
            cudaMemPoolProps pr; memset(&pr, 0, sizeof(pr));
            pr.allocType = cudaMemAllocationTypePinned; pr.handleTypes = cudaMemHandleTypeNone;
            pr.location.type = cudaMemLocationTypeDevice; pr.location.id = dev;
            cudaMemPool_t mp = nullptr;
            if (cudaMemPoolCreate(&mp, &pr) == cudaSuccess) {
                unsigned long long keep = ~0ull;            // nothing goes back to the driver before trim_scratch
                cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
                g_pools.pool[dev] = mp;
            } else cudaGetLastError();
        }
    }
    return g_pools.pool[dev];
}

struct Buf {            // scratch allocation freed on scope exit (stream-ordered when it came from the pool)
    void* p = nullptr;
    cudaStream_t st = nullptr; bool pooled = false;
    ~Buf() { if (p) { if (pooled) cudaFreeAsync(p, st); else cudaFree(p); } }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

}  // namespace

int build_on_device(cudaStream_t st, const BuildInput& in, BuildOutput& out, std::string& err) {
    const int n = in.n_host + in.n_bulk;
    out = BuildOutput();
    if (n <= 0) return 0;
    const int n_inner = n - 1;
    const bool timing = getenv("RTW_TIMING") != nullptr;
    cudaEvent_t ev[8]; int n_ev = 0;
    const auto host_t0 = std::chrono::steady_clock::now();
    double malloc_ms = 0;                                       // host time inside cudaMalloc (RTW_TIMING)
    auto dmalloc = [&](auto** p, size_t bytes) {
        if (!timing) return cudaMalloc(p, bytes);
        const auto a = std::chrono::steady_clock::now();
        const cudaError_t e = cudaMalloc(p, bytes);
        malloc_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
        return e;
    };
    int dev = 0;
    BCUDA(cudaGetDevice(&dev));
    cudaMemPool_t mpool = scratch_pool(dev);
    auto salloc = [&](Buf& b, size_t bytes) {                   // scratch: from the pool, in stream order on `st`
        if (!mpool) return dmalloc(&b.p, bytes);
        const auto a = std::chrono::steady_clock::now();
        const cudaError_t e = cudaMallocFromPoolAsync(&b.p, bytes ? bytes : 16, mpool, st);
        if (timing) malloc_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
        if (e == cudaSuccess) { b.st = st; b.pooled = true; }
        return e;
    };
    auto mark = [&]() { if (timing && n_ev < 8) { cudaEventCreate(&ev[n_ev]); cudaEventRecord(ev[n_ev], st); ++n_ev; } };
    mark();
    // ---- inputs
    Buf d_bulk, d_hprims, d_hboxes;
    if (in.n_bulk) {
        BCUDA(salloc(d_bulk, (size_t)in.n_bulk * sizeof(BulkSphereD)));
        if (d_bulk.pooled) BCUDA(cudaStreamSynchronize(st));     // the upload's own streams may touch the buffer from here on
        BCUDA(staged_upload(st, d_bulk.p, in.bulk, (size_t)in.n_bulk * sizeof(BulkSphereD)));
    }
    if (in.n_host) {
        BCUDA(salloc(d_hprims, (size_t)in.n_host * sizeof(DPrim)));
        BCUDA(salloc(d_hboxes, (size_t)in.n_host * 24));
        BCUDA(cudaMemcpyAsync(d_hprims.p, in.host_prims, (size_t)in.n_host * sizeof(DPrim), cudaMemcpyHostToDevice, st));
        BCUDA(cudaMemcpyAsync(d_hboxes.p, in.host_boxes, (size_t)in.n_host * 24, cudaMemcpyHostToDevice, st));
    }
    out.h2d_bytes = (size_t)in.n_bulk * sizeof(BulkSphereD) + (size_t)in.n_host * (sizeof(DPrim) + 24);
    mark();
    // ---- boxes, Morton codes, sort
    Buf d_boxes, d_bounds, d_keys, d_vals, d_keys2, d_vals2, d_tmp;
    BCUDA(salloc(d_boxes, (size_t)n * 24));
    BCUDA(salloc(d_bounds, 64));
    {
        unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
        BCUDA(cudaMemcpyAsync(d_bounds.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    const int T = 256, G = (n + T - 1) / T;
    prim_boxes_kernel<<<G, T, 0, st>>>(n, in.n_host, d_hboxes.as<float>(), d_bulk.as<BulkSphereD>(), d_boxes.as<float>(), d_bounds.as<unsigned>());
    BCUDA(cudaGetLastError());
    Buf d_wnodes_scratch, d_order, d_q0, d_q1, d_counters, d_left, d_right, d_parent, d_box2, d_flags, d_count;
    BCUDA(salloc(d_order, (size_t)n * 4));
    BCUDA(salloc(d_counters, 64));
    DWNode* wn = nullptr;
    int n_wide = 0, depth = 0;
    DNode* bn = nullptr; int n_bin = 0;
    Buf d_bnodes;
    if (n == 1 && in.width == 2) {          // one primitive: a root whose two slots point at it (like the host flattener)
        BCUDA(salloc(d_bnodes, sizeof(DNode)));
        float hb[6];
        BCUDA(cudaMemcpyAsync(hb, d_boxes.p, 24, cudaMemcpyDeviceToHost, st));
        BCUDA(cudaStreamSynchronize(st));
        DNode d; memset(&d, 0, sizeof(d));
        d.c0minx = d.c1minx = hb[0]; d.c0maxx = d.c1maxx = hb[3]; d.c0miny = d.c1miny = hb[1]; d.c0maxy = d.c1maxy = hb[4];
        d.c0minz = d.c1minz = hb[2]; d.c0maxz = d.c1maxz = hb[5]; d.child0 = d.child1 = ~0;
        BCUDA(cudaMemcpyAsync(d_bnodes.p, &d, sizeof(d), cudaMemcpyHostToDevice, st));
        BCUDA(cudaStreamSynchronize(st));
        bn = d_bnodes.as<DNode>(); n_bin = 1; depth = 1;
        BCUDA(salloc(d_vals, 4));
        BCUDA(cudaMemsetAsync(d_vals.p, 0, 4, st));
        BCUDA(cudaMemsetAsync(d_order.p, 0, 4, st));
    } else if (n == 1) {
        BCUDA(salloc(d_wnodes_scratch, sizeof(DWNode)));
        wn = d_wnodes_scratch.as<DWNode>();
        single_leaf_root_kernel<<<1, 1, 0, st>>>(d_boxes.as<float>(), wn, d_order.as<int>());
        BCUDA(cudaGetLastError());
        BCUDA(salloc(d_vals, 4));
        BCUDA(cudaMemsetAsync(d_vals.p, 0, 4, st));
        n_wide = 1; depth = 1;
    } else {
        BCUDA(salloc(d_keys, (size_t)n * 8)); BCUDA(salloc(d_vals, (size_t)n * 4));
        BCUDA(salloc(d_keys2, (size_t)n * 8)); BCUDA(salloc(d_vals2, (size_t)n * 4));
        morton_kernel<<<G, T, 0, st>>>(n, d_boxes.as<float>(), d_bounds.as<unsigned>(), d_keys.as<unsigned long long>(), d_vals.as<unsigned>());
        BCUDA(cudaGetLastError());
        cub::DoubleBuffer<unsigned long long> kb(d_keys.as<unsigned long long>(), d_keys2.as<unsigned long long>());
        cub::DoubleBuffer<unsigned> vb(d_vals.as<unsigned>(), d_vals2.as<unsigned>());
        size_t tmp_bytes = 0;
        BCUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, n, 0, 63, st));
        BCUDA(salloc(d_tmp, tmp_bytes ? tmp_bytes : 16));
        BCUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, kb, vb, n, 0, 63, st));
        const unsigned long long* keys = kb.Current();
        const unsigned* sorted_idx = vb.Current();
        mark();
        // ---- binary radix tree + refit
        BCUDA(salloc(d_left, (size_t)n_inner * 4)); BCUDA(salloc(d_right, (size_t)n_inner * 4));
        BCUDA(salloc(d_parent, (size_t)(n_inner + n) * 4));
        BCUDA(salloc(d_box2, (size_t)(n_inner + n) * 24));
        BCUDA(salloc(d_flags, (size_t)n_inner * 4));
        BCUDA(cudaMemsetAsync(d_flags.p, 0, (size_t)n_inner * 4, st));
        BCUDA(salloc(d_count, (size_t)n_inner * 4));
        karras_kernel<<<(n_inner + T - 1) / T, T, 0, st>>>(n, keys, d_left.as<int>(), d_right.as<int>(), d_parent.as<int>());
        BCUDA(cudaGetLastError());
        refit_kernel<<<G, T, 0, st>>>(n, sorted_idx, d_boxes.as<float>(), d_left.as<int>(), d_right.as<int>(), d_parent.as<int>(), d_box2.as<float>(), d_flags.as<int>(), d_count.as<int>());
        BCUDA(cudaGetLastError());
        mark();
        if (in.width == 2) {
            BCUDA(salloc(d_bnodes, (size_t)n_inner * sizeof(DNode)));
            bn = d_bnodes.as<DNode>(); n_bin = n_inner;
            emit_binary_kernel<<<(n_inner + T - 1) / T, T, 0, st>>>(n_inner, d_left.as<int>(), d_right.as<int>(), d_box2.as<float>(), bn);
            BCUDA(cudaGetLastError());
            iota_kernel<<<G, T, 0, st>>>(n, d_order.as<int>());
            BCUDA(cudaMemsetAsync(d_counters.p, 0, 16, st));
            binary_depth_kernel<<<G, T, 0, st>>>(n, d_parent.as<int>(), d_counters.as<int>() + 3);
            BCUDA(cudaGetLastError());
            int h[4];
            BCUDA(cudaMemcpyAsync(h, d_counters.p, sizeof(h), cudaMemcpyDeviceToHost, st));
            BCUDA(cudaStreamSynchronize(st));
            depth = h[3];
            if (sorted_idx != d_vals.as<unsigned>()) std::swap(d_vals.p, d_vals2.p);
        } else {
        // ---- collapse, level by level (every wide node consumes at least one binary inner node: n_inner bounds everything)
        BCUDA(salloc(d_wnodes_scratch, (size_t)n_inner * sizeof(DWNode)));
        BCUDA(salloc(d_q0, (size_t)n_inner * sizeof(rtww::WideItem)));
        BCUDA(salloc(d_q1, (size_t)n_inner * sizeof(rtww::WideItem)));
        wn = d_wnodes_scratch.as<DWNode>();
        {
            int init[4] = {1, 0, 0, 0};
            rtww::WideItem root{0, 0, 0};
            BCUDA(cudaMemcpyAsync(d_counters.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
            BCUDA(cudaMemcpyAsync(d_q0.p, &root, sizeof(root), cudaMemcpyHostToDevice, st));
        }
        static const bool fill = !(getenv("RTW_WIDE_FILL") && atoi(getenv("RTW_WIDE_FILL")) == 0);
        rtww::B2View view{d_box2.as<float>(), d_left.as<int>(), d_right.as<int>(), n_inner, n, fill ? d_count.as<int>() : nullptr};
        rtww::WideItem* cur = d_q0.as<rtww::WideItem>();
        rtww::WideItem* next = d_q1.as<rtww::WideItem>();
        int n_cur = 1, levels = 0;
        while (n_cur > 0) {
            collapse_kernel<<<(n_cur + 127) / 128, 128, 0, st>>>(view, cur, n_cur, wn, d_order.as<int>(), next, d_counters.as<int>());
            BCUDA(cudaGetLastError());
            int h[4];
            BCUDA(cudaMemcpyAsync(h, d_counters.p, sizeof(h), cudaMemcpyDeviceToHost, st));
            BCUDA(cudaStreamSynchronize(st));
            n_cur = h[2]; n_wide = h[0]; depth = h[3];
            BCUDA(cudaMemsetAsync(d_counters.as<int>() + 2, 0, 4, st));
            std::swap(cur, next);
            if (++levels > 128) { err = "device BVH collapse did not terminate"; return -3; }
        }
        // the sorted index array must outlive the double buffer juggling: keep whichever buffer holds it
        if (sorted_idx != d_vals.as<unsigned>()) std::swap(d_vals.p, d_vals2.p);
        }
    }
    mark();
    if (in.width == 2 ? depth > 60 : depth > RTW_WIDE_STACK) { err = "device-built BVH deeper than the traversal stack (degenerate primitive distribution)"; return -2; }
    // ---- outputs: compact node array + primitive records in leaf order (+ room for the medium boundary records)
    if (in.width == 2) {
        BCUDA(dmalloc(&out.nodes, (size_t)n_bin * sizeof(DNode)));
        BCUDA(cudaMemcpyAsync(out.nodes, bn, (size_t)n_bin * sizeof(DNode), cudaMemcpyDeviceToDevice, st));
    } else {
        BCUDA(dmalloc(&out.wnodes, (size_t)n_wide * sizeof(DWNode)));
        BCUDA(cudaMemcpyAsync(out.wnodes, wn, (size_t)n_wide * sizeof(DWNode), cudaMemcpyDeviceToDevice, st));
    }
    BCUDA(dmalloc(&out.prims, (size_t)(n + in.n_boundary) * sizeof(DPrim)));
    emit_prims_kernel<<<G, T, 0, st>>>(n, in.n_host, d_order.as<int>(), d_vals.as<unsigned>(), d_hprims.as<DPrim>(), d_bulk.as<BulkSphereD>(), out.prims);
    BCUDA(cudaGetLastError());
    if (in.n_boundary)
        BCUDA(cudaMemcpyAsync(out.prims + n, in.boundary_prims, (size_t)in.n_boundary * sizeof(DPrim), cudaMemcpyHostToDevice, st));
    mark();
    BCUDA(cudaStreamSynchronize(st));
    out.n_wnodes = n_wide; out.n_nodes = n_bin; out.depth = depth; out.n_prims = n + in.n_boundary;
    out.h2d_bytes += (size_t)in.n_boundary * sizeof(DPrim);
    if (timing && n_ev >= 2) {
        const char* names[] = {"upload", "boxes+morton+sort", "radix tree+refit", "collapse", "emit"};
        for (int i = 1; i < n_ev; ++i) { float ms = 0; cudaEventElapsedTime(&ms, ev[i - 1], ev[i]); fprintf(stderr, "[device build] %s %.2f ms\n", names[i - 1], ms); }
        fprintf(stderr, "[device build] %d prims -> %d wide nodes, depth %d; host: %.2f ms in all, %.2f ms of it inside cudaMalloc (scratch is freed after this line)\n", n, n_wide, depth,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(), malloc_ms);
    }
    for (int i = 0; i < n_ev; ++i) cudaEventDestroy(ev[i]);
    return 0;
}

void trim_scratch(int dev) {
    if (dev < 0 || dev >= kMaxDev) return;
    std::lock_guard<std::mutex> lk(g_pools.mu);
    if (g_pools.pool[dev]) cudaMemPoolTrimTo(g_pools.pool[dev], 0);
}

void free_output(BuildOutput& o) {
    if (o.wnodes) cudaFree(o.wnodes);
    if (o.nodes) cudaFree(o.nodes);
    if (o.prims) cudaFree(o.prims);
    o = BuildOutput();
}

}  // namespace rtwb
