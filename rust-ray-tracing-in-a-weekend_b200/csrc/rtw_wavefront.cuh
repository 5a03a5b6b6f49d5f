// rtw_wavefront.cuh — the wavefront pipeline of the north star, for scenes that do not fit the caches.
//
// The megakernel (rtw_api.cu render_kernel) keeps one path per lane and shades between two traversals; on the 1 M - 16 M
// sphere sweep its lanes idle (9.6 of 32 per instruction: a warp waits for its longest traversal) on top of a chain of
// dependent node fetches.  Here the four stages of src/main.rs:19-38 run as two kernels over a POOL of paths in HBM:
//
//   wf_logic  : one thread per pool slot.  A slot whose ray has been traced is shaded — media, hit record, emitted +
//               scatter (path_finish: the megakernel's device functions, same Philox coordinates, so the image is the
//               megakernel's up to f32 summation order) — and either keeps going (next ray written in place) or ends
//               (radiance added to the framebuffer).  Free slots REGENERATE: they take the next camera path from a
//               counter (src/main.rs:517-520), so the pool stays full until the image runs out of samples.
//   wf_trace  : persistent warps with DYNAMIC RAY FETCH: a lane that finishes its ray stores (t, primitive) and takes the
//               next slot from a cursor, so all 32 lanes keep traversing whatever the spread of traversal lengths; the
//               queue is millions of rays long, there is no per-stage tail.  Default: the LOCKSTEP loop over 8-wide nodes
//               (wf_trace2w_kernel); the speculative while-while loop (wf_trace_kernel<F, W>) and the lockstep loop over
//               binary nodes (wf_trace2_kernel) are the measured alternatives.
//
// Queue traffic per ray segment: 48 B read + 8 B written by the trace kernel, 72 B read + 64 B written by the logic kernel,
// all float4 / uint2, slot-indexed (coalesced).  Paths are numbered so that 32 consecutive ones are the 32 pixels of a tile at
// one sample; path numbers are handed out in chunks of 2^20 from ONE counter that all GPUs share (dynamic balance, ~2 000
// remote atomics per frame); finished paths are added to the framebuffer — directly on the GPU that owns it, through a local
// framebuffer merged once per frame on the others (wf_merge_kernel).
#ifndef RTW_WAVEFRONT_CUH
#define RTW_WAVEFRONT_CUH

#include "rtw_device.cuh"

namespace rtwd {

#define RTW_WF_CHUNK_LOG2 20
#define RTW_WF_BLOCK 256
enum { WF_EMPTY = 0, WF_TRACE = 1 };

struct WfPool {                 // device pointers, P slots each
    float4* od0;                // ox, oy, oz, time
    float4* od1;                // dx, dy, dz, last_prim (int bits)
    float4* thr;                // Tx, Ty, Tz, segment (int bits)
    float4* rad;                // Lx, Ly, Lz, status (int bits)
    uint2* id;                  // pixel (y * W + x, y bottom-up), sample
    uint2* hit;                 // t (float bits), primitive (-1: miss)
    // [0] local path numbers handed out, [1] slots that hold a ray after the last logic pass, [2] trace cursor,
    // [3] rays traced, [4] chunks reserved, [5] paths started, [6] global counter exhausted (flag)
    unsigned long long* ctr;
    unsigned long long* chunk_base;   // chunk table: local chunk c covers global paths [chunk_base[c], + 2^20)
    unsigned long long* dbg;          // instrumented build: per-warp (start, end, rays) of the last trace launch, 3 x 8192
    int P, max_chunks;
};

// global path number -> (pixel, sample): 32 consecutive paths = the 32 pixels of one 8x4 tile at one sample, then the
// next sample of that tile, then the next tile (ragged edge tiles hold fewer pixels; their missing pixels are skipped)
RTW_DEV bool wf_decode_path(const DParams& prm, unsigned long long p, int& x, int& y, int& sample) {
    const unsigned long long per_tile = 32ull * (unsigned long long)prm.spp;
    const unsigned long long tile = p / per_tile;
    const unsigned r = (unsigned)(p - tile * per_tile);
    const int s = (int)(r >> 5), pl = (int)(r & 31u);
    const int tx = (int)(tile % (unsigned long long)prm.tiles_x), ty = (int)(tile / (unsigned long long)prm.tiles_x);
    x = tx * 8 + (pl & 7); y = ty * 4 + (pl >> 3);
    sample = prm.first_sample + s;
    return x < prm.width && y < prm.height;
}

// tops up the chunk table so that the logic pass that follows can hand out up to P new paths (one thread)
__global__ void wf_reserve_kernel(WfPool pool, unsigned long long* __restrict__ global_counter, unsigned long long total_paths) {
    unsigned long long handed = pool.ctr[0], chunks = pool.ctr[4];
    while ((chunks << RTW_WF_CHUNK_LOG2) < handed + (unsigned long long)pool.P && chunks < (unsigned long long)pool.max_chunks && !pool.ctr[6]) {
        const unsigned long long base = atomicAdd_system(global_counter, 1ull << RTW_WF_CHUNK_LOG2);
        if (base >= total_paths) { pool.ctr[6] = 1ull; break; }
        pool.chunk_base[chunks++] = base;
    }
    pool.ctr[4] = chunks;
    pool.ctr[1] = 0ull; pool.ctr[2] = 0ull;          // per-iteration counters: active slots, trace cursor
}

// (spheres-only variant: 64 registers without spills = 4 CTAs per SM instead of 3: sweep 1 M / 4 M +0.7 % / +1.1 %, profiles/r2_ag_logic.log)
template <int F>
__global__ void __launch_bounds__(RTW_WF_BLOCK, F == 0 ? 4 : 1)
wf_logic_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam, const __grid_constant__ DParams prm, WfPool pool,
                unsigned long long total_paths, float* __restrict__ fb) {
    __shared__ unsigned s_need[RTW_WF_BLOCK / 32];
    __shared__ unsigned long long s_base;
    const int i = blockIdx.x * RTW_WF_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool in = i < pool.P;
    bool need_new = in;
    bool active = false;
    if (in) {
        const float4 rad = pool.rad[i];
        if (__float_as_int(rad.w) == WF_TRACE) {
            // ---- one level of ray_color (src/main.rs:25-37) for the segment the trace kernel just closed
            const float4 a = pool.od0[i], b = pool.od1[i], t = pool.thr[i];
            const uint2 id = pool.id[i], h = pool.hit[i];
            PathState ps;
            ps.ray.o = mk(a.x, a.y, a.z); ps.ray.time = a.w; ps.ray.d = mk(b.x, b.y, b.z);
            ps.last_prim = __float_as_int(b.w);
            ps.T = mk(t.x, t.y, t.z); ps.segment = __float_as_int(t.w);
            ps.rng.bind(prm); ps.rng.start(id.x, id.y); ps.rng.set_bounce((uint32_t)ps.segment);
            const TRay tr = make_tray(ps.ray);
            V3 add;
            const bool cont = path_finish<F, false>(sc, prm, ps, tr, __uint_as_float(h.x), (int)h.y, add);
            V3 L = mk(rad.x + add.x, rad.y + add.y, rad.z + add.z);
            if (cont && ps.segment < prm.max_depth) {                     // src/main.rs:21-23: the next level still has depth left
                pool.od0[i] = make_float4(ps.ray.o.x, ps.ray.o.y, ps.ray.o.z, ps.ray.time);
                pool.od1[i] = make_float4(ps.ray.d.x, ps.ray.d.y, ps.ray.d.z, __int_as_float(ps.last_prim));
                pool.thr[i] = make_float4(ps.T.x, ps.T.y, ps.T.z, __int_as_float(ps.segment + 1));
                pool.rad[i] = make_float4(L.x, L.y, L.z, __int_as_float(WF_TRACE));
                need_new = false; active = true;
            } else {
                // path done: its radiance joins the pixel's sum (row 0 of the image = top, src/main.rs:591)
                const int px = (int)(id.x % (unsigned)prm.width), py = (int)(id.x / (unsigned)prm.width);
                float* dst = fb + ((size_t)(prm.height - 1 - py) * prm.width + px) * 3;
                if (L.x != 0.f) atomicAdd_system(dst, L.x);
                if (L.y != 0.f) atomicAdd_system(dst + 1, L.y);
                if (L.z != 0.f) atomicAdd_system(dst + 2, L.z);
            }
        }
    }
    // ---- regeneration: free slots take the next camera paths (one atomic per block)
    const unsigned m = __ballot_sync(0xffffffffu, need_new);
    if (lane == 0) s_need[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int w = 0; w < RTW_WF_BLOCK / 32; ++w) { const unsigned c = s_need[w]; s_need[w] = tot; tot += c; }
        s_base = tot ? atomicAdd(&pool.ctr[0], (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    if (need_new) {
        const unsigned long long n = s_base + s_need[warp] + __popc(m & ((1u << lane) - 1u));     // local path number
        const unsigned long long c = n >> RTW_WF_CHUNK_LOG2;
        bool started = false;
        if (c < pool.ctr[4]) {
            const unsigned long long p = pool.chunk_base[c] + (n & ((1ull << RTW_WF_CHUNK_LOG2) - 1ull));
            int x, y, s;
            if (p < total_paths && prm.max_depth >= 1 && wf_decode_path(prm, p, x, y, s)) {
                PathState ps;
                path_begin(cam, prm, x, y, s, ps);
                pool.od0[i] = make_float4(ps.ray.o.x, ps.ray.o.y, ps.ray.o.z, ps.ray.time);
                pool.od1[i] = make_float4(ps.ray.d.x, ps.ray.d.y, ps.ray.d.z, __int_as_float(-1));
                pool.thr[i] = make_float4(1.f, 1.f, 1.f, __int_as_float(1));
                pool.rad[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(WF_TRACE));
                pool.id[i] = make_uint2((unsigned)(y * prm.width + x), (unsigned)s);
                started = true; active = true;
            }
        }
        if (!started) pool.rad[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(WF_EMPTY));
        const unsigned ms = __ballot_sync(__activemask(), started);
        if (started && lane == __ffs(ms) - 1) atomicAdd(&pool.ctr[5], (unsigned long long)__popc(ms));
        // (a number beyond the reserved chunks is simply not used: wf_reserve tops the table up before the next pass; a
        // number that decodes to a pixel outside a ragged edge tile is skipped)
    }
    const unsigned ma = __ballot_sync(0xffffffffu, active);
    if (lane == 0 && ma) atomicAdd(&pool.ctr[1], (unsigned long long)__popc(ma));
}

// Multi-GPU: a GPU that does not own the framebuffer accumulates its paths in a LOCAL framebuffer and adds it to the owner's
// once, at the end of the frame (the north star's "final framebuffer gathered to GPU 0 over NVLink"; src/main.rs:542-547) —
// instead of three remote atomics per finished path.  Only pixels this GPU touched travel; the local buffer is left zeroed.
// Four floats per red (red.global.add.v4.f32, sm_90+).  The whole merge of a 3840x2160 image takes 0.1 ms (profiles/r2_aq_merge.log);
// what it replaces — three remote atomics per finished path from the logic kernel — slowed a non-owner rank down by 8 %.
__global__ void wf_merge_kernel(float* __restrict__ local, float* __restrict__ remote, size_t n) {
    const size_t n4 = ((reinterpret_cast<uintptr_t>(local) | reinterpret_cast<uintptr_t>(remote)) & 15) == 0 ? n / 4 : 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(local)[i];
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
            asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(remote + 4 * i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            reinterpret_cast<float4*>(local)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    for (size_t i = 4 * n4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = local[i];
        if (v != 0.f) { atomicAdd_system(remote + i, v); local[i] = 0.f; }
    }
}

// Persistent trace kernel: closest surface hit (hit_hittables src/hittable.rs:43-55 over the whole world) for every slot
// that holds a ray.  Lanes fetch slots dynamically; W = 0 binary nodes, W = 1 8-wide compressed nodes.
#define RTW_WF_FETCH_THRESHOLD 20
template <int F, int W>
__global__ void __launch_bounds__(128, 6)
wf_trace_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DParams prm, WfPool pool) {
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned long long rays = 0;
    bool active = false, exhausted = false;
    int slot = 0;
    TRay r; V3 inv = mk(0, 0, 0), cmn = mk(0, 0, 0), cmx = mk(0, 0, 0);
    float t_best = 0.f; int prim_best = -1, skip = -1;
    uint32_t stack[RTW_STACK]; int sp = 1, node = RTW_SENTINEL, leaf = 0;
    uint32_t wbase = 0, wgrp = 0; rtww::WRay wr; wr.one = 0x3F800000u | ((uint32_t)sc.n_nodes >> 31); wr.k = 0; wr.sx = wr.sy = wr.sz = 0.f;
    wr.ix = wr.iy = wr.iz = wr.oix = wr.oiy = wr.oiz = 0.f;
    bool done = true;
    stack[0] = (uint32_t)RTW_SENTINEL;
    for (;;) {
        // ---- refill idle lanes from the cursor (one atomic per warp and refill)
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle && !exhausted) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&pool.ctr[2], (unsigned long long)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= (unsigned long long)pool.P) exhausted = true;
            if (!active) {
                const unsigned long long k = base + __popc(idle & lt_mask);
                if (k < (unsigned long long)pool.P) {
                    const float4 rd = pool.rad[k];
                    if (__float_as_int(rd.w) == WF_TRACE) {
                        slot = (int)k;
                        const float4 a = pool.od0[k], b = pool.od1[k];
                        Ray ray; ray.o = mk(a.x, a.y, a.z); ray.d = mk(b.x, b.y, b.z); ray.time = a.w;
                        r = make_tray(ray);
                        slab_setup(r.o, r.d, inv, cmn, cmx);
                        t_best = CUDART_INF_F; prim_best = -1; skip = __float_as_int(b.w);
                        if (W) {
                            slab_setup_wide(r.o, r.d, wr);
                            sp = 0; wbase = 0; wgrp = sc.n_bvh_prims ? ((1u << 8) | (1u << wr.k)) : 0u; done = false;
                        } else { sp = 1; leaf = 0; node = sc.n_bvh_prims ? 0 : RTW_SENTINEL; }
                        active = true; ++rays;
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, active)) { if (exhausted) break; else continue; }
        if (active) {
            if (W) {
                while (!done) {
                    if (!(wgrp & 0xffu)) {
                        if (sp == 0) { done = true; break; }
                        sp -= 2; wbase = stack[sp]; wgrp = stack[sp + 1];
                        continue;
                    }
                    const int j = 31 - __clz(wgrp & 0xffu);
                    wgrp ^= 1u << j;
                    const uint32_t sl = (uint32_t)j ^ wr.k;
                    const uint32_t nd = wbase + __popc((wgrp >> 8) & ((1u << sl) - 1u));
                    if (wgrp & 0xffu) { stack[sp] = wbase; stack[sp + 1] = wgrp; sp += 2; }
                    const uint4* np = reinterpret_cast<const uint4*>(sc.wnodes + nd);
                    const uint4 h = __ldg(np), m = __ldg(np + 1), qa = __ldg(np + 2), qb = __ldg(np + 3), qc = __ldg(np + 4);
                    const uint32_t imask = h.w >> 24, lmask = m.z & 0xffu;
                    uint32_t hits = rtww::wide_node_hits(*reinterpret_cast<const rtww::W4*>(&h), *reinterpret_cast<const rtww::W4*>(&qa),
                                                         *reinterpret_cast<const rtww::W4*>(&qb), *reinterpret_cast<const rtww::W4*>(&qc), wr, prm.t_min, t_best);
                    hits &= imask | lmask;
                    const uint32_t m16 = rtww::wide_perm16((hits & imask) | ((hits & lmask) << 8), wr.k);
                    uint32_t pl = m16 >> 8;
                    while (pl) {
                        const int jj = 31 - __clz(pl);
                        pl ^= 1u << jj;
                        const uint32_t s2 = (uint32_t)jj ^ wr.k;
                        const int pi = (int)(m.y + __popc(lmask & ((1u << s2) - 1u)));
                        int hp;
                        const float t = prim_root<F>(sc, pi, r, prm.t_min, t_best, skip, hp);
                        if (t == t) { t_best = t; prim_best = hp; }
                    }
                    wbase = m.x; wgrp = (imask << 8) | (m16 & 0xffu);
                    if (!exhausted && __popc(__activemask()) < RTW_WF_FETCH_THRESHOLD) break;       // let the idle lanes refill
                }
                node = done ? RTW_SENTINEL : 0;
            } else {
                while (node != RTW_SENTINEL) {
                    bool searching = true;
                    while (node >= 0 && node != RTW_SENTINEL) {
                        const float4* np = reinterpret_cast<const float4*>(sc.nodes + node);
                        const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
                        const int2 ch = __ldg(reinterpret_cast<const int2*>(np + 3));
                        float e0, e1;
                        const bool h0 = slab(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, inv, cmn, cmx, prm.t_min, t_best, e0);
                        const bool h1 = slab(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, inv, cmn, cmx, prm.t_min, t_best, e1);
                        if (!h0 && !h1) node = (int)stack[--sp];
                        else {
                            node = h0 ? ch.x : ch.y;
                            if (h0 && h1) {
                                int farc = ch.y;
                                if (e1 < e0) { farc = node; node = ch.y; }
                                stack[sp++] = (uint32_t)farc;
                            }
                        }
                        if (node < 0 && leaf >= 0) { searching = false; leaf = node; node = (int)stack[--sp]; }
                        if (!__any_sync(__activemask(), searching)) break;
                    }
                    while (leaf < 0) {
                        const int code = ~leaf, first = code >> 3, count = (code & 7) + 1;
                        for (int q = 0; q < count; ++q) {
                            int hp;
                            const float t = prim_root<F>(sc, first + q, r, prm.t_min, t_best, skip, hp);
                            if (t == t) { t_best = t; prim_best = hp; }
                        }
                        leaf = node;
                        if (node < 0) node = (int)stack[--sp];
                    }
                    if (!exhausted && __popc(__activemask()) < RTW_WF_FETCH_THRESHOLD) break;       // let the idle lanes refill
                }
            }
            if (node == RTW_SENTINEL) {
                pool.hit[slot] = make_uint2(__float_as_uint(t_best), (unsigned)prim_best);
                active = false;
            }
        }
    }
    for (int o = 16; o; o >>= 1) rays += __shfl_xor_sync(0xffffffffu, rays, o);
    if (lane == 0 && rays) atomicAdd(&pool.ctr[3], rays);
}

// The same stage as a LOCKSTEP loop over binary nodes: every iteration all lanes that hold an inner node visit it (one
// two-box test), lanes that reached a leaf keep it pending (and go on speculatively until they reach a second one), and the
// pending leaves are intersected together once LEAF_T lanes wait or nobody can descend any more.  With the speculative
// while-while loop above a lane that holds two leaves sits out the rest of the node phase: ncu shows 9 of 32 lanes in the node
// test although dynamic fetch keeps 20-32 rays per warp in flight.  Here the node test runs with every lane that has a node.
#ifndef RTW_WF_LEAF_T
#define RTW_WF_LEAF_T 8
#endif
#ifndef RTW_WF_REFILL_T
#define RTW_WF_REFILL_T 8
#endif
#ifndef RTW_WF_MIN_BLOCKS
#define RTW_WF_MIN_BLOCKS 6
#endif
template <int F>
__global__ void __launch_bounds__(128, 6)
wf_trace2_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DParams prm, WfPool pool) {
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned long long rays = 0;
    bool active = false, exhausted = false;
    int slot = 0;
    TRay r; V3 inv = mk(0, 0, 0), cmn = mk(0, 0, 0), cmx = mk(0, 0, 0);
    float t_best = 0.f; int prim_best = -1, skip = -1;
    int stack[RTW_STACK]; int sp = 1, node = RTW_SENTINEL, leaf = 0;
    stack[0] = RTW_SENTINEL;
    for (;;) {
        // ---- (1) refill
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (!exhausted && (__popc(idle) >= RTW_WF_REFILL_T)) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&pool.ctr[2], (unsigned long long)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= (unsigned long long)pool.P) exhausted = true;
            if (!active) {
                const unsigned long long k = base + __popc(idle & lt_mask);
                if (k < (unsigned long long)pool.P) {
                    const float4 rd = pool.rad[k];
                    if (__float_as_int(rd.w) == WF_TRACE) {
                        slot = (int)k;
                        const float4 a = pool.od0[k], b = pool.od1[k];
                        Ray ray; ray.o = mk(a.x, a.y, a.z); ray.d = mk(b.x, b.y, b.z); ray.time = a.w;
                        r = make_tray(ray);
                        slab_setup(r.o, r.d, inv, cmn, cmx);
                        t_best = CUDART_INF_F; prim_best = -1; skip = __float_as_int(b.w);
                        sp = 1; leaf = 0; node = sc.n_bvh_prims ? 0 : RTW_SENTINEL;
                        active = true; ++rays;
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, active)) { if (exhausted) break; else continue; }
        // ---- (2) node step
        const bool can_node = active && node >= 0 && node != RTW_SENTINEL;
        if (can_node) {
            const float4* np = reinterpret_cast<const float4*>(sc.nodes + node);
            const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
            const int2 ch = __ldg(reinterpret_cast<const int2*>(np + 3));
            float e0, e1;
            const bool h0 = slab(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, inv, cmn, cmx, prm.t_min, t_best, e0);
            const bool h1 = slab(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, inv, cmn, cmx, prm.t_min, t_best, e1);
            const bool second = h1 & (!h0 | (e1 < e0));
            const int nearc = second ? ch.y : ch.x, farc = second ? ch.x : ch.y;
            node = nearc;
            if (h0 && h1) { stack[sp] = farc; ++sp; }
            if (!(h0 || h1)) { --sp; node = stack[sp]; }
            if (node < 0 && leaf == 0) { leaf = node; --sp; node = stack[sp]; }     // first leaf: keep it pending, go on
        }
        // ---- (3) leaf step: when enough lanes wait, or nobody can descend
        const bool want_leaf = active && leaf < 0;
        const unsigned ml = __ballot_sync(0xffffffffu, want_leaf);
        const unsigned mn = __ballot_sync(0xffffffffu, active && node >= 0 && node != RTW_SENTINEL);
        if (ml && (__popc(ml) >= RTW_WF_LEAF_T || mn == 0u)) {
            if (want_leaf) {
                const int code = ~leaf, first = code >> 3, count = (code & 7) + 1;
#pragma unroll 1
                for (int q = 0; q < count; ++q) {
                    int hp;
                    const float t = prim_root<F>(sc, first + q, r, prm.t_min, t_best, skip, hp);
                    if (t == t) { t_best = t; prim_best = hp; }
                }
                leaf = 0;
                if (node < 0) { leaf = node; --sp; node = stack[sp]; }                // the second leaf reached meanwhile
            }
        }
        // ---- (4) rays that are done
        if (active && node == RTW_SENTINEL && leaf == 0) {
            pool.hit[slot] = make_uint2(__float_as_uint(t_best), (unsigned)prim_best);
            active = false;
        }
    }
    for (int o = 16; o; o >>= 1) rays += __shfl_xor_sync(0xffffffffu, rays, o);
    if (lane == 0 && rays) atomicAdd(&pool.ctr[3], rays);
}

// Lockstep loop over the 8-WIDE compressed nodes: the node test (the long, ALU-heavy part: ~245 instructions) runs with
// every lane that has a node group to descend; the leaves a visit yields stay pending as a group until LEAF_T lanes hold one.
template <int F>
__global__ void __launch_bounds__(128, RTW_WF_MIN_BLOCKS)
wf_trace2w_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DParams prm, WfPool pool) {
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned long long rays = 0;
    bool active = false, exhausted = false;
    int slot = 0;
    TRay r;
    float t_best = 0.f; int prim_best = -1, skip = -1;
    uint32_t stack[2 * RTW_WIDE_STACK]; int sp = 0;
    uint32_t wbase = 0, wgrp = 0, lbase = 0, lgrp = 0;       // node group; pending leaf group (lmask << 8 | permuted hit bits)
    rtww::WRay wr; wr.one = 0x3F800000u | ((uint32_t)sc.n_nodes >> 31); wr.k = 0; wr.sx = wr.sy = wr.sz = 0.f;
    wr.ix = wr.iy = wr.iz = wr.oix = wr.oiy = wr.oiz = 0.f;
#ifdef RTW_INSTRUMENT
    unsigned long long dbg_t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
#endif
    for (;;) {
        // ---- (1) refill
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (!exhausted && (__popc(idle) >= RTW_WF_REFILL_T)) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&pool.ctr[2], (unsigned long long)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= (unsigned long long)pool.P) exhausted = true;
            if (!active) {
                const unsigned long long k = base + __popc(idle & lt_mask);
                if (k < (unsigned long long)pool.P) {
                    const float4 rd = pool.rad[k];
                    if (__float_as_int(rd.w) == WF_TRACE) {
                        slot = (int)k;
                        const float4 a = pool.od0[k], b = pool.od1[k];
                        Ray ray; ray.o = mk(a.x, a.y, a.z); ray.d = mk(b.x, b.y, b.z); ray.time = a.w;
                        r = make_tray(ray);
                        slab_setup_wide(r.o, r.d, wr);
                        t_best = CUDART_INF_F; prim_best = -1; skip = __float_as_int(b.w);
                        sp = 0; wbase = 0; wgrp = sc.n_bvh_prims ? ((1u << 8) | (1u << wr.k)) : 0u; lgrp = 0;
                        active = true; ++rays;
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, active)) { if (exhausted) break; else continue; }
        // ---- (2) node step: one wide node per lane that has a group to descend and no leaves pending
        const bool can_node = active && !(lgrp & 0xffu) && ((wgrp & 0xffu) || sp > 0);
        if (can_node) {
            if (!(wgrp & 0xffu)) { sp -= 2; wbase = stack[sp]; wgrp = stack[sp + 1]; }
            const int j = 31 - __clz(wgrp & 0xffu);
            wgrp ^= 1u << j;
            const uint32_t sl = (uint32_t)j ^ wr.k;
            const uint32_t nd = wbase + __popc((wgrp >> 8) & ((1u << sl) - 1u));
            if (wgrp & 0xffu) { stack[sp] = wbase; stack[sp + 1] = wgrp; sp += 2; }
            const uint4* np = reinterpret_cast<const uint4*>(sc.wnodes + nd);
            const uint4 h = __ldg(np), m = __ldg(np + 1), qa = __ldg(np + 2), qb = __ldg(np + 3), qc = __ldg(np + 4);
            const uint32_t imask = h.w >> 24, lmask = m.z & 0xffu;
            uint32_t hits = rtww::wide_node_hits(*reinterpret_cast<const rtww::W4*>(&h), *reinterpret_cast<const rtww::W4*>(&qa),
                                                 *reinterpret_cast<const rtww::W4*>(&qb), *reinterpret_cast<const rtww::W4*>(&qc), wr, prm.t_min, t_best);
            hits &= imask | lmask;
            const uint32_t m16 = rtww::wide_perm16((hits & imask) | ((hits & lmask) << 8), wr.k);
            wbase = m.x; wgrp = (imask << 8) | (m16 & 0xffu);
            lbase = m.y; lgrp = (lmask << 8) | (m16 >> 8);
        }
        // ---- (3) leaf step
        const bool want_leaf = active && (lgrp & 0xffu);
        const unsigned ml = __ballot_sync(0xffffffffu, want_leaf);
        const unsigned mn = __ballot_sync(0xffffffffu, active && !(lgrp & 0xffu) && ((wgrp & 0xffu) || sp > 0));
        if (ml && (__popc(ml) >= RTW_WF_LEAF_T || mn == 0u)) {
            while (__any_sync(0xffffffffu, active && (lgrp & 0xffu))) {
                if (active && (lgrp & 0xffu)) {
                    const int jj = 31 - __clz(lgrp & 0xffu);
                    lgrp ^= 1u << jj;
                    const uint32_t s2 = (uint32_t)jj ^ wr.k;
                    const int pi = (int)(lbase + __popc((lgrp >> 8) & ((1u << s2) - 1u)));
                    int hp;
                    const float t = prim_root<F>(sc, pi, r, prm.t_min, t_best, skip, hp);
                    if (t == t) { t_best = t; prim_best = hp; }
                }
            }
        }
        // ---- (4) rays that are done
        if (active && !(wgrp & 0xffu) && sp == 0 && !(lgrp & 0xffu)) {
            pool.hit[slot] = make_uint2(__float_as_uint(t_best), (unsigned)prim_best);
            active = false;
        }
    }
    for (int o = 16; o; o >>= 1) rays += __shfl_xor_sync(0xffffffffu, rays, o);
#ifdef RTW_INSTRUMENT
    if (lane == 0) {
        unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        const unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if (w < 8192u) { pool.dbg[w] = dbg_t0; pool.dbg[8192 + w] = t1; pool.dbg[16384 + w] = rays; }
    }
#endif
    if (lane == 0 && rays) atomicAdd(&pool.ctr[3], rays);
}

}  // namespace rtwd

#endif
