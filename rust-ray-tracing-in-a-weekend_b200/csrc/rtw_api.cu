// rtw_api.cu — kernels + the C ABI of include/rtw.h.
//
// Render = ONE persistent megakernel per GPU (render_kernel<F>, F = the scene's feature set): every warp pulls work
// units (an 8x4-pixel tile x a chunk of samples) from a single atomic counter that all GPUs share.  Per unit it culls
// the BVH against the tile's primary-ray bundle, then alternates dense primary batches (32 camera paths tested against
// the tile's candidate list) with secondary steps (one BVH-traversed segment per live lane, rays queued in a
// shared-memory ring), accumulates radiance per tile in shared memory and adds the finished tile straight into the
// framebuffer on the first GPU (peer atomics over NVLink — the "gather" of src/main.rs:542-547 is fused into the
// kernel).  Replaces src/main.rs:497-589.  DESIGN.md 4.1 has the measurements behind each choice.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rtw.h"
#include "bvh_build.hpp"
#include "rtw_device.cuh"
#include "rtw_pool.cuh"
#include "rtw_wavefront.cuh"
#include "scene_host.hpp"

using namespace rtwd;

// =================================================================================================
// kernels
// =================================================================================================
#ifdef RTW_INSTRUMENT
__device__ unsigned long long g_dbg_counters[8];     // [0..5] secondary steps (see below), [6] primary-ray primitive tests, [7] primary rays
// warp timeline of the last launch (ns, %globaltimer): [0] first warp start, [1] last warp start, [2] sum of warp end times,
// [3] last warp end, [4] warps, [5] first warp end
__device__ unsigned long long g_dbg_time[6] = {~0ull, 0, 0, 0, 0, ~0ull};
extern "C" int rtw_debug_counters(unsigned long long out[8], int reset) {
    if (cudaMemcpyFromSymbol(out, g_dbg_counters, 64) != cudaSuccess) return -3;
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_dbg_counters, z, 64); }
    return 0;
}
extern "C" int rtw_debug_timeline(unsigned long long out[6]) {
    if (cudaMemcpyFromSymbol(out, g_dbg_time, 48) != cudaSuccess) return -3;
    unsigned long long z[6] = {~0ull, 0, 0, 0, 0, ~0ull}; cudaMemcpyToSymbol(g_dbg_time, z, 48);
    return 0;
}
__device__ __forceinline__ unsigned long long dbg_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif
#define RTW_BLOCK 128
#define RTW_WARPS (RTW_BLOCK / 32)
#ifndef RTW_DEFAULT_MODE
#define RTW_DEFAULT_MODE 0
#endif
#ifndef RTW_WF_DEFAULT_POOLS
#define RTW_WF_DEFAULT_POOLS 1
#endif
#ifndef RTW_DEFAULT_POOL_MODE
#define RTW_DEFAULT_POOL_MODE 2
#endif
// 7 CTAs = 28 warps per SM at 72 registers: measured best of 6 / 7 / 8 on every config once the kernel had been shrunk
// (8 CTAs = 64 registers spill ~30 words per thread; 6 CTAs lose more latency hiding than the registers buy)
#ifndef RTW_NOTILE_MEDIA
#define RTW_NOTILE_MEDIA 1
#endif
#ifndef RTW_MIN_BLOCKS
#define RTW_MIN_BLOCKS 7
#endif
// Carry-over of the last rays of a work unit into the next one (1): when a unit has no more paths to start and its ring can no
// longer fill the idle lanes, the warp does not run the rest of the unit under-occupied (the "drain": 3.4 % of C1 at the
// 8-GPU unit sizes, profiles/r2_ab_units_tile_list_share.log) — it flushes the tile, keeps the rays that are still in flight
// (lanes + ring) as ORPHANS of the old tile and starts the next unit's primary batches around them.  Orphans carry one bit in
// their pixel field and add their radiance straight to the framebuffer (global atomics: a few dozen paths per unit) instead of
// the tile accumulator; at most one generation of orphans is in flight per warp (a counter in shared memory).
#ifndef RTW_CARRY
#define RTW_CARRY 1
#endif
#ifndef RTW_COOP
#define RTW_COOP 1          // unit-ball samples drawn by the whole warp (coop_unit_sphere); 0 = every lane loops on its own
#endif

// S = 1: the variant for scenes of a handful of primitives (DParams::list_max): every ray scans ALL primitives instead of
// walking the BVH — no traversal code in the kernel at all.  S = 0: everything else (no scan overhead: C1 pays 0.55 % for a
// run-time switch, profiles/r2_ak_c1_scan_ab.log).
template <int F, int W, int S>
__global__ void __launch_bounds__(RTW_BLOCK, RTW_MIN_BLOCKS)
render_kernel(DScene sc, DCamera cam, const __grid_constant__ DParams prm, unsigned int* __restrict__ unit_counter, float* __restrict__ fb,
              unsigned long long* __restrict__ stats /* [0] rays, [1] units */) {
    __shared__ __align__(16) float acc[RTW_WARPS][96];    // per-warp tile accumulator (32 px x rgb)
    __shared__ float ring[RTW_WARPS][12][64];             // secondary-ray ring: o(3) d(3) time T(3) last_prim meta, 64 per warp
    __shared__ int tlist[RTW_WARPS][RTW_TILE_LIST];       // primitives the tile's primary rays can touch
    __shared__ int coop_scr[RTW_WARPS][32];               // coop_unit_sphere: failed-lane directory
#if RTW_CARRY
    __shared__ int carry_s[RTW_WARPS][4];                 // orphans in flight, tile x / y of their unit
    __shared__ int tl_scratch[RTW_WARPS][136];            // build_tile_list scratch (the ring is in use across units)
#endif
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned long long rays = 0, units = 0;
#ifdef RTW_INSTRUMENT
    unsigned long long dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    PathState ps;
    ps.rng.bind(prm);                                          // key schedule: a constant-bank address, set once
#ifdef RTW_INSTRUMENT
    if (lane == 0) { const unsigned long long t = dbg_now(); atomicMin(&g_dbg_time[0], t); atomicMax(&g_dbg_time[1], t); }
#endif
    // (measured per variant, profiles/r2_ac_carry.log: it pays where the kernel has registers to spare — C1 at the 8-GPU unit
    // sizes -2 %, at full-frame small units 105.7 -> 102.1 ms, two_spheres -5 %, two_perlin_spheres -3 %, cornell_box -1 % — and
    // costs in the variants that carry the media code: final_scene +2.7 %, cornell_box_smoke +7 %: off there)
    constexpr bool kCarry = RTW_CARRY != 0 && !(F & FEAT_MEDIA) && F != FEAT_IMAGE;       // (earth, 1.3 rays per path, 2.3 ms: +5 % with it)
    int pix = 0, ring_head = 0, ring_count = 0;                // (carry: these survive from unit to unit)
    bool alive = false, no_more = false;
#if RTW_CARRY
    if (lane == 0) carry_s[warp][0] = 0;
    __syncwarp();
#endif
    for (;;) {
        unsigned unit = 0;
        int tile = 0, s0 = 0, s1 = 0;
        if (!no_more) {
            if (lane == 0) unit = atomicAdd_system(unit_counter, 1u);
            unit = __shfl_sync(0xffffffffu, unit, 0) * prm.unit_stride;
            no_more = unit >= prm.n_units;
        }
        if (no_more && (!kCarry || (ring_count == 0 && !__any_sync(0xffffffffu, alive)))) break;     // no unit left and nothing carried over
        const bool have_unit = !no_more;                       // (false: a last pass that only finishes the orphans)
        if (have_unit) { ++units; decode_unit(prm, unit, tile, s0, s1); }
        if (!kCarry) { pix = 0; ring_head = 0; ring_count = 0; alive = false; }
        const int tx = tile % prm.tiles_x, ty = tile / prm.tiles_x;
        const int tw = min(8, prm.width - tx * 8), th = min(4, prm.height - ty * 4), npix = tw * th;     // ragged edge tiles
        const int n_items = (have_unit && prm.max_depth >= 1) ? npix * (s1 - s0) : 0;
        acc[warp][lane] = 0.f; acc[warp][lane + 32] = 0.f; acc[warp][lane + 64] = 0.f;
#if RTW_CARRY
        int* const tl_scr = kCarry ? tl_scratch[warp] : reinterpret_cast<int*>(ring[warp][0]);
#else
        int* const tl_scr = reinterpret_cast<int*>(ring[warp][0]);
#endif
        // every primitive a primary ray of this tile can touch (-1: too many, traverse instead)
        // (the small-scene kernels keep their tile lists: without them two_spheres -4.5 % but simple_light +3.8 %, two_perlin_spheres
        // +5.7 %, cornell_box +0.8 %: profiles/r2_al_small_notile.log)
        // (not in the big variants that carry the media code — final_scene: their 69 KB are past the reach of the instruction caches,
        // DESIGN 10-3; without the list code the variant is 736 instructions shorter and final_scene renders in 436.6 instead of
        // 448.3 ms, profiles/r2_ax_final_scene_code_size.log; primary rays then walk the BVH like every other ray)
        constexpr bool kNoTile = RTW_NOTILE_MEDIA != 0 && (F & FEAT_MEDIA) && S == 0;
        const int list_n = (kNoTile || prm.no_tile_cull || !have_unit) ? -1
                           : W ? build_tile_list_wide(sc, tile_ray_bounds(cam, prm, tx * 8, ty * 4, tw, th), prm.t_min, tlist[warp], tl_scr, lane)
                               : build_tile_list(sc, tile_ray_bounds(cam, prm, tx * 8, ty * 4, tw, th), prm.t_min, tlist[warp], tl_scr, lane);
        __syncwarp();
        if (lane == 0 && have_unit) { if (list_n >= 0) atomicAdd(stats + 2, (unsigned long long)list_n); else atomicAdd(stats + 3, 1ull); }
        int next = 0;
        // Each iteration of this loop is EITHER a primary batch or a secondary step; both run the SAME copy of the
        // closest-hit and shading code (two inlined copies doubled the kernel to 145 KB and cost 40 % in i-cache misses).
        //  primary batch : taken when the ring cannot serve the lanes that need a path.  All 32 lanes queue their
        //                  in-flight path in the ring, start a new camera path each (main.rs:517-520), intersect it with the tile's
        //                  candidate list (dense: same trip count in every lane) and shade it.  Paths that end on their
        //                  first hit are accumulated right away; the survivors' SECONDARY rays go into the ring.
        //  secondary step: dead lanes take a ray from the ring, every live lane traces + shades one more segment.
        // The divergent BVH traversal therefore only ever sees secondary rays (63 % of the rays in C1).
        for (;;) {
            const bool need = !alive;
            const unsigned mask = __ballot_sync(0xffffffffu, need);
            const int n_need = __popc(mask);
            const bool primary = ring_count < n_need && next < n_items;      // warp-uniform
#if RTW_CARRY
            if (kCarry && ring_count < n_need && !primary && have_unit) {
                // Nothing left to start in this unit and the ring cannot fill the idle lanes: hand what is still in flight
                // over to the next unit (unless an older generation of orphans is still alive: then drain as usual).
                const int live = (32 - n_need) + ring_count;
                const int older = __shfl_sync(0xffffffffu, carry_s[warp][0], 0);      // (one lane's view: the decision is warp-uniform)
                if (live > 0 && older == 0) {
                    if (alive) pix |= 32;
                    for (int k = lane; k < ring_count; k += 32) {
                        float* rg = ring[warp][0] + ((ring_head + k) & 63);
                        rg[704] = __int_as_float(__float_as_int(rg[704]) | 32);
                    }
                    if (lane == 0) { carry_s[warp][0] = live; carry_s[warp][1] = tx; carry_s[warp][2] = ty; }
                    __syncwarp();
                    break;
                }
            }
#endif
            bool work = false;
            int wpix = 0;
            if (primary) {
                // in-flight paths go to the back of the ring (their segment count travels in the entry); after the batch
                // every lane is free and refills from the ring in FIFO order
                const unsigned ma = __ballot_sync(0xffffffffu, alive);
                if (alive) {
                    float* rg = ring[warp][0] + ((ring_head + ring_count + __popc(ma & lt_mask)) & 63);
                    rg[0] = ps.ray.o.x; rg[64] = ps.ray.o.y; rg[128] = ps.ray.o.z;
                    rg[192] = ps.ray.d.x; rg[256] = ps.ray.d.y; rg[320] = ps.ray.d.z; rg[384] = ps.ray.time;
                    rg[448] = ps.T.x; rg[512] = ps.T.y; rg[576] = ps.T.z;
                    rg[640] = __int_as_float(ps.last_prim);
                    rg[704] = __int_as_float(pix | (ps.segment << 6) | ((int)ps.rng.sample << 12));      // pixel(5) orphan(1) segment(6) sample(20)
                    alive = false;
                }
                ring_count += __popc(ma);
                const int idx = next + lane;
                if (idx < n_items) {
                    const int pl = idx % npix, sample = s0 + idx / npix;
                    const int px = pl % tw, py = pl / tw;
                    wpix = py * 8 + px;
                    path_begin(cam, prm, tx * 8 + px, ty * 4 + py, sample, ps, true);
                    ps.segment = 1;
                    ps.rng.set_bounce(1u);
                    work = true;
                }
            } else {
                if (need) {
                    const int rnk = __popc(mask & lt_mask);
                    if (rnk < ring_count) {
                        const float* rg = ring[warp][0] + ((ring_head + rnk) & 63);
                        ps.ray.o = mk(rg[0], rg[64], rg[128]); ps.ray.d = mk(rg[192], rg[256], rg[320]); ps.ray.time = rg[384];
                        ps.T = mk(rg[448], rg[512], rg[576]);
                        ps.last_prim = __float_as_int(rg[640]);
                        const int meta = __float_as_int(rg[704]);
                        pix = meta & 63;
#if RTW_CARRY
                        const int rtx = (kCarry && (pix & 32)) ? carry_s[warp][1] : tx, rty = (kCarry && (pix & 32)) ? carry_s[warp][2] : ty;     // an orphan belongs to the previous unit's tile
#else
                        const int rtx = tx, rty = ty;
#endif
                        ps.rng.start((uint32_t)((rty * 4 + ((pix >> 3) & 3)) * prm.width + rtx * 8 + (pix & 7)), (uint32_t)meta >> 12);
                        ps.segment = (meta >> 6) & 63;
                        alive = true;
                    }
                }
                const int taken = min(n_need, ring_count);
                ring_head = (ring_head + taken) & 63; ring_count -= taken;
                if (!__any_sync(0xffffffffu, alive)) break;     // ring empty, nothing in flight, no items left
                if (alive) {
                    if (ps.segment >= prm.max_depth) {                                       // main.rs:21-23
                        alive = false;
#if RTW_CARRY
                        if (kCarry && (pix & 32)) atomicSub(&carry_s[warp][0], 1);
#endif
                    } else { ps.segment++; ps.rng.set_bounce((uint32_t)ps.segment); work = true; wpix = pix; }
                }
            }
#ifdef RTW_INSTRUMENT
            int dbg_v = 0, dbg_p = 0;
#endif
            // ---- closest surface hit (main.rs:25)
            TRay tr; float t_best = CUDART_INF_F; int prim_best = -1;
            if (work) {
                tr = make_tray(ps.ray);
                // Candidate scan instead of a BVH walk: primary rays over their tile's list; in the small-scene kernels (S = 1) EVERY
                // ray over all primitives (cornell_box: 8 leaves — the same trip count in every lane against ~6 divergent node visits
                // + 2-3 leaf tests).  One copy of the loop serves both.  Measured (profiles/r2_ah_list_scan.log): cornell_box 145.7 ->
                // 121.2 ms, two_spheres / earth -7 %, the other small scenes -3 %.
                constexpr bool kScan = S != 0;
                const bool use_tile = primary && list_n >= 0;
                if (use_tile || kScan) {
                    const int n_scan = (!kScan || use_tile) ? list_n : sc.n_bvh_prims;
                    const int skip = (!kScan || primary) ? -1 : ps.last_prim;
#pragma unroll 1
                    for (int i = 0; i < n_scan; ++i) {
                        const int pi = (!kScan || use_tile) ? tlist[warp][i] : i;
                        int hp;
                        const float t = prim_root<F>(sc, pi, tr, prm.t_min, t_best, skip, hp);
                        if (t == t) { t_best = t; prim_best = hp; }
                    }
                } else {
#ifdef RTW_INSTRUMENT
                    closest_hit<F, W>(sc, tr, prm.t_min, t_best, prim_best, ps.last_prim, dbg_v, dbg_p);
#else
                    closest_hit<F, W>(sc, tr, prm.t_min, t_best, prim_best, ps.last_prim);
#endif
                }
            }
            // ---- media, miss, hit record, emitted + scatter (main.rs:26-37)
            bool cont = false;
            V3 add;
            // unit-ball samples by the whole warp — except in the variants that carry the Perlin / image texture code, where
            // the extra live state across the sampling loop costs more than the loop saves (measured: C1 -2.0 %, cornell_box
            // -1.2 %, but final_scene +11 %, two_perlin_spheres +6 %: profiles/r2_d_ab_coop.log)
            constexpr bool kCoop = RTW_COOP != 0 && !(F & (FEAT_NOISE | FEAT_IMAGE));
            cont = path_finish<F, kCoop>(sc, prm, ps, tr, t_best, prim_best, add, work, lane, coop_scr[warp]);    // all 32 lanes
            if (work) {
                ++rays;
                if (add.x != 0.f || add.y != 0.f || add.z != 0.f) {          // miss / emitter: T*background, T*emitted
#if RTW_CARRY
                    if (kCarry && (wpix & 32)) {                             // orphan of the previous unit: its tile was flushed, add to the image itself
                        const int ox = carry_s[warp][1] * 8 + (wpix & 7), oy = carry_s[warp][2] * 4 + ((wpix >> 3) & 3);
                        float* dst = fb + ((size_t)(prm.height - 1 - oy) * prm.width + ox) * 3;
                        if (add.x != 0.f) atomicAdd_system(dst, add.x);
                        if (add.y != 0.f) atomicAdd_system(dst + 1, add.y);
                        if (add.z != 0.f) atomicAdd_system(dst + 2, add.z);
                    } else
#endif
                    {
                        atomicAdd(&acc[warp][wpix * 3 + 0], add.x);
                        atomicAdd(&acc[warp][wpix * 3 + 1], add.y);
                        atomicAdd(&acc[warp][wpix * 3 + 2], add.z);
                    }
                }
            }
#ifdef RTW_INSTRUMENT
            if (!primary) {   // per secondary step: lanes at work, sum and max of node visits / prim tests over the lanes
                int v = dbg_v, p = dbg_p, a = work ? 1 : 0;
                int vs = v, vm = v, pss = p, pm = p, as = a;
                for (int o = 16; o; o >>= 1) {
                    vs += __shfl_xor_sync(0xffffffffu, vs, o); vm = max(vm, __shfl_xor_sync(0xffffffffu, vm, o));
                    pss += __shfl_xor_sync(0xffffffffu, pss, o); pm = max(pm, __shfl_xor_sync(0xffffffffu, pm, o));
                    as += __shfl_xor_sync(0xffffffffu, as, o);
                }
                if (lane == 0) { dbg[0] += 1; dbg[1] += as; dbg[2] += vs; dbg[3] += vm; dbg[4] += pss; dbg[5] += pm; }
            } else {          // primary batch: every working lane tests the tile's candidate list (or traverses: counted in dbg_p)
                const unsigned wm = __ballot_sync(0xffffffffu, work);
                int p = dbg_p; for (int o = 16; o; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
                if (lane == 0) { dbg[6] += (unsigned long long)__popc(wm) * (unsigned long long)(list_n >= 0 ? list_n : 0) + (unsigned long long)p; dbg[7] += __popc(wm); }
            }
#endif
            if (primary) {
                const unsigned mc = __ballot_sync(0xffffffffu, cont);
                if (cont) {
                    float* rg = ring[warp][0] + ((ring_head + ring_count + __popc(mc & lt_mask)) & 63);
                    rg[0] = ps.ray.o.x; rg[64] = ps.ray.o.y; rg[128] = ps.ray.o.z;
                    rg[192] = ps.ray.d.x; rg[256] = ps.ray.d.y; rg[320] = ps.ray.d.z; rg[384] = ps.ray.time;
                    rg[448] = ps.T.x; rg[512] = ps.T.y; rg[576] = ps.T.z;
                    rg[640] = __int_as_float(ps.last_prim);
                    rg[704] = __int_as_float(wpix | (ps.segment << 6) | ((int)ps.rng.sample << 12));
                }
                ring_count += __popc(mc);
                next += min(32, n_items - next);
                __syncwarp();
            } else if (work) {
                alive = cont;
#if RTW_CARRY
                if (kCarry && !cont && (pix & 32)) atomicSub(&carry_s[warp][0], 1);
#endif
            }
        }
        __syncwarp();
        if (have_unit) flush_tile(prm, fb, acc[warp], tx, ty, tw, lane);
        __syncwarp();
    }
    // ray statistics: one atomic per warp
    for (int o = 16; o; o >>= 1) rays += __shfl_xor_sync(0xffffffffu, rays, o);
    if (lane == 0) { atomicAdd(stats, rays); atomicAdd(stats + 1, units); }
#ifdef RTW_INSTRUMENT
    if (lane == 0) for (int k = 0; k < 8; ++k) atomicAdd(&g_dbg_counters[k], dbg[k]);
    if (lane == 0) { const unsigned long long t = dbg_now(); atomicAdd(&g_dbg_time[2], t); atomicMax(&g_dbg_time[3], t); atomicAdd(&g_dbg_time[4], 1ull); atomicMin(&g_dbg_time[5], t); }
#endif
}

// The warp-pool kernel (see rtw_pool.cuh): same work units, same Philox keys, same framebuffer protocol as
// render_kernel — only the scheduling of the four stages differs.
template <int POOL, int W>
__global__ void __launch_bounds__(RTW_BLOCK)
render_pool_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam, const __grid_constant__ DParams prm, unsigned int* __restrict__ unit_counter, float* __restrict__ fb,
                   unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(16) unsigned char pool_raw[];
    PoolSmem<POOL>* pools = reinterpret_cast<PoolSmem<POOL>*>(pool_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    PoolSmem<POOL>& P = pools[warp];
    unsigned long long rays = 0, units = 0;
    for (;;) {
        unsigned unit = 0;
        if (lane == 0) unit = atomicAdd_system(unit_counter, 1u);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= prm.n_units) break;
        ++units;
        int tile, s0, s1;
        decode_unit(prm, unit, tile, s0, s1);
        const int tile_x0 = (tile % prm.tiles_x) * 8, tile_y0 = (tile / prm.tiles_x) * 4;
        const int tw = min(8, prm.width - tile_x0), th = min(4, prm.height - tile_y0), npix = tw * th;   // ragged edge tiles
        const int n_items = prm.max_depth >= 1 ? npix * (s1 - s0) : 0;
        for (int i = lane; i < POOL; i += 32) { P.meta[i] = POOL_FREE; P.freel[i] = (unsigned char)i; }
        P.acc[lane] = 0.f; P.acc[lane + 32] = 0.f; P.acc[lane + 64] = 0.f;
        if (lane < 8) P.cnt[lane] = lane == C_FREE ? POOL : 0;
        __syncwarp();
        int next_item = 0;
        for (;;) {
            // ---- regenerate: free slots <- camera rays (main.rs:517-520).  freel is a stack, trav a list.
            const int n_free = P.cnt[C_FREE];
            const int n_cont = P.cnt[C_TRAV];                  // paths continuing from the last shade stage
            const int n_new = min(n_free, n_items - next_item);
            __syncwarp();
            for (int i = lane; i < n_new; i += 32) {
                const int slot = P.freel[n_free - 1 - i];
                const int idx = next_item + i, pl = idx % npix, sample = s0 + idx / npix;
                const int px = pl % tw, py = pl / tw, pix = py * 8 + px;
                PathState ps;
                path_begin(cam, prm, tile_x0 + px, tile_y0 + py, sample, ps);
                P.ox[slot] = ps.ray.o.x; P.oy[slot] = ps.ray.o.y; P.oz[slot] = ps.ray.o.z;
                P.dx[slot] = ps.ray.d.x; P.dy[slot] = ps.ray.d.y; P.dz[slot] = ps.ray.d.z; P.tm[slot] = ps.ray.time;
                P.tr[slot] = 1.f; P.tg[slot] = 1.f; P.tb[slot] = 1.f;
                P.last[slot] = -1;
                P.meta[slot] = meta_pack(pix, 1, 0, sample);
                P.trav[n_cont + i] = (unsigned char)slot;
            }
            next_item += n_new;
            const int n_trav = n_cont + n_new;
            __syncwarp();
            if (lane < 8) P.cnt[lane] = lane == C_FREE ? n_free - n_new : 0;
            __syncwarp();
            if (n_trav == 0) break;                            // nothing alive and nothing left to start
            // ---- traverse (hittable.rs:43-55) + classify
            rays += traverse_stage<POOL, W>(P, lane, lt_mask, n_trav, sc, prm, tile_x0, tile_y0);
            __syncwarp();
            // ---- shade, one dense loop per material kind (material.rs:15-94)
            shade_list<POOL, K_MISS>(P, lane, sc, prm, tile_x0, tile_y0);
            shade_list<POOL, K_LAMB>(P, lane, sc, prm, tile_x0, tile_y0);
            shade_list<POOL, K_METAL>(P, lane, sc, prm, tile_x0, tile_y0);
            shade_list<POOL, K_DIEL>(P, lane, sc, prm, tile_x0, tile_y0);
            shade_list<POOL, K_LIGHT>(P, lane, sc, prm, tile_x0, tile_y0);
            shade_list<POOL, K_ISO>(P, lane, sc, prm, tile_x0, tile_y0);
            __syncwarp();
        }
        flush_tile(prm, fb, P.acc, tile_x0 >> 3, tile_y0 >> 2, tw, lane);
        __syncwarp();
    }
    for (int o = 16; o; o >>= 1) rays += __shfl_xor_sync(0xffffffffu, rays, o);
    if (lane == 0) { atomicAdd(stats, rays); atomicAdd(stats + 1, units); }
}

// per-path radiance with the render's Philox keys (parity hook rtw_trace_paths)
__global__ void trace_paths_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam, const __grid_constant__ DParams prm, int n, const int* __restrict__ px, const int* __restrict__ py,
                                   const int* __restrict__ smp, double* __restrict__ out_rgb, int* __restrict__ out_seg) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PathState ps;
    path_begin(cam, prm, px[i], py[i], smp[i], ps);
    V3 L = mk(0.f, 0.f, 0.f), add;
    bool go = true;
    while (go) { go = path_step(sc, prm, ps, add); L = L + add; }
    out_rgb[3 * i] = L.x; out_rgb[3 * i + 1] = L.y; out_rgb[3 * i + 2] = L.z;
    out_seg[i] = ps.segment;
}

__global__ void philox_kernel(int n, const uint32_t* __restrict__ ctr, const uint32_t* __restrict__ key, uint32_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1],
                  out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
}

__global__ void get_ray_kernel(DCamera cam, int n, const double* __restrict__ s, const double* __restrict__ t, const double* __restrict__ xi,
                               int stride, double* __restrict__ oo, double* __restrict__ od, double* __restrict__ ot, int* __restrict__ nd) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    StreamRng g; g.xi = xi + (size_t)i * stride; g.n = stride; g.draw = 0;
    Ray r = camera_get_ray(cam, (float)s[i], (float)t[i], g);
    oo[3 * i] = r.o.x; oo[3 * i + 1] = r.o.y; oo[3 * i + 2] = r.o.z;
    od[3 * i] = r.d.x; od[3 * i + 1] = r.d.y; od[3 * i + 2] = r.d.z;
    ot[i] = r.time; nd[i] = g.draw > stride ? -1 : g.draw;
}

__global__ void hit_kernel(DScene sc, int n, const double* __restrict__ o, const double* __restrict__ d, const double* __restrict__ tm,
                           float t_min, float t_max, const double* __restrict__ xi, int stride, int* __restrict__ hit,
                           double* __restrict__ ot, double* __restrict__ op, double* __restrict__ on, int* __restrict__ front,
                           double* __restrict__ ou, double* __restrict__ ov, int* __restrict__ mat, int* __restrict__ nd) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    StreamRng g; g.xi = xi + (size_t)i * stride; g.n = stride; g.draw = 0;
    Ray r; r.o = mk((float)o[3 * i], (float)o[3 * i + 1], (float)o[3 * i + 2]);
    r.d = mk((float)d[3 * i], (float)d[3 * i + 1], (float)d[3 * i + 2]); r.time = (float)tm[i];
    TRay tr = make_tray(r);
    HitRec rec; rec.t = 0; rec.p = mk(0, 0, 0); rec.normal = mk(0, 0, 0); rec.front = 0; rec.mat = -1; rec.u = 0; rec.v = 0;
    bool h = world_hit(sc, tr, t_min, t_max, g, true, rec);
    hit[i] = h ? 1 : 0;
    ot[i] = rec.t; op[3 * i] = rec.p.x; op[3 * i + 1] = rec.p.y; op[3 * i + 2] = rec.p.z;
    on[3 * i] = rec.normal.x; on[3 * i + 1] = rec.normal.y; on[3 * i + 2] = rec.normal.z;
    front[i] = rec.front; ou[i] = rec.u; ov[i] = rec.v; mat[i] = rec.mat + 1;   // back to the 1-based handle
    nd[i] = g.draw > stride ? -1 : g.draw;
}

__global__ void aabb_kernel(int n, const double* __restrict__ bmin, const double* __restrict__ bmax, const double* __restrict__ o,
                            const double* __restrict__ d, float t_min, float t_max, int* __restrict__ hit) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    V3 ro = mk((float)o[3 * i], (float)o[3 * i + 1], (float)o[3 * i + 2]);
    V3 rd = mk((float)d[3 * i], (float)d[3 * i + 1], (float)d[3 * i + 2]);
    V3 inv, cmn, cmx; float e;
    slab_setup(ro, rd, inv, cmn, cmx);
    hit[i] = slab((float)bmin[3 * i], (float)bmax[3 * i], (float)bmin[3 * i + 1], (float)bmax[3 * i + 1], (float)bmin[3 * i + 2],
                  (float)bmax[3 * i + 2], inv, cmn, cmx, t_min, t_max, e) ? 1 : 0;
}

__global__ void scatter_kernel(DScene sc, int mat, int n, const double* __restrict__ ro, const double* __restrict__ rd, const double* __restrict__ rt,
                               const double* __restrict__ p, const double* __restrict__ nrm, const int* __restrict__ front,
                               const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ xi, int stride,
                               int* __restrict__ osc, double* __restrict__ oo, double* __restrict__ od, double* __restrict__ ot,
                               double* __restrict__ oatt, double* __restrict__ oem, int* __restrict__ nd) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    StreamRng g; g.xi = xi + (size_t)i * stride; g.n = stride; g.draw = 0;
    Ray r; r.o = mk((float)ro[3 * i], (float)ro[3 * i + 1], (float)ro[3 * i + 2]);
    r.d = mk((float)rd[3 * i], (float)rd[3 * i + 1], (float)rd[3 * i + 2]); r.time = (float)rt[i];
    HitRec rec; rec.p = mk((float)p[3 * i], (float)p[3 * i + 1], (float)p[3 * i + 2]);
    rec.normal = mk((float)nrm[3 * i], (float)nrm[3 * i + 1], (float)nrm[3 * i + 2]);
    rec.front = front[i]; rec.u = (float)u[i]; rec.v = (float)v[i]; rec.t = 0; rec.mat = mat;
    DMatRec m = load_mat(sc, mat);
    Ray s; s.o = mk(0, 0, 0); s.d = mk(0, 0, 0); s.time = 0; V3 att = mk(0, 0, 0), em = mk(0, 0, 0);
    bool ok = scatter(sc, m, r, rec, g, s, att, em);
    if (!ok) { s.o = mk(0, 0, 0); s.d = mk(0, 0, 0); s.time = 0; att = mk(0, 0, 0); }
    osc[i] = ok ? 1 : 0;
    oo[3 * i] = s.o.x; oo[3 * i + 1] = s.o.y; oo[3 * i + 2] = s.o.z;
    od[3 * i] = s.d.x; od[3 * i + 1] = s.d.y; od[3 * i + 2] = s.d.z; ot[i] = s.time;
    oatt[3 * i] = att.x; oatt[3 * i + 1] = att.y; oatt[3 * i + 2] = att.z;
    oem[3 * i] = em.x; oem[3 * i + 1] = em.y; oem[3 * i + 2] = em.z;
    nd[i] = g.draw > stride ? -1 : g.draw;
}

__global__ void texture_kernel(DScene sc, int tex, int n, const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ p,
                               double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    V3 c = texture_value(sc, tex, (float)u[i], (float)v[i], mk((float)p[3 * i], (float)p[3 * i + 1], (float)p[3 * i + 2]));
    out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
}

// write_color (src/math.rs:119-132): sqrt(c * (1/spp)), clamp [0, 0.999], *256 truncated; NaN -> 0.
// Evaluated in f64 like the reference so the bytes are bit-exact for identical sums.
__global__ void write_color_kernel(const float* __restrict__ sum, int n, double scale, uint8_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r = sqrt((double)sum[i] * scale);
    double c = r < 0.0 ? 0.0 : (r > 0.999 ? 0.999 : r);
    c *= 256.0;
    out[i] = (c == c) ? (uint8_t)(int)c : 0;
}

// =================================================================================================
// host side
// =================================================================================================
namespace {

thread_local std::string g_err;
int fail(int code, const std::string& m) { g_err = m; return code; }
#define CUDA_TRY(x)                                                                                  \
    do {                                                                                             \
        cudaError_t e_ = (x);                                                                        \
        if (e_ != cudaSuccess) return fail(e_ == cudaErrorMemoryAllocation ? RTW_ERR_OOM : RTW_ERR_CUDA, \
                                           std::string(#x) + ": " + cudaGetErrorString(e_));        \
    } while (0)

#define TRY(x) do { int rc_ = (x); if (rc_ < 0) return rc_; } while (0)

// The reference computes with whatever it is given and renders NaN pixels (or panics); across the ABI the contract is
// a status code: every constructor rejects non-finite numbers and the degenerate values that divide by zero later.
bool fin(double v) { return std::isfinite(v); }
bool fin3(const double* a) { return a && fin(a[0]) && fin(a[1]) && fin(a[2]); }
template <class... T> bool fin_all(T... v) { bool ok = true; for (double x : {(double)v...}) ok = ok && fin(x); return ok; }

struct Replica {
    int device = -1;
    uint8_t* blob = nullptr;
    size_t blob_bytes = 0;
    DScene ds{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned long long* stats = nullptr;    // [0] rays [1] units
    int grid = 0;                           // megakernel grid (also sizes the work units)
    int pool_grid[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // warp-pool kernel grids for POOL = 64/128/192/256, binary then wide (lazy)
    int sms = 0;
    rtwb::BuildOutput built;                // device-built BVH + primitive records (own allocations, outside the blob)
    WfPool wf[4] = {};                      // wavefront path pools (lazy; rtw_wavefront.cuh): wf_pools independent pools, one stream each
    int wf_pools = 0; long long wf_slots = 0;
    cudaStream_t wf_stream[4] = {nullptr, nullptr, nullptr, nullptr};   // [0] = stream
    cudaEvent_t wf_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    uint8_t* wf_mem = nullptr;
    float* wf_fb = nullptr; size_t wf_fb_floats = 0;      // local framebuffer of a GPU that does not own the shared one (merged once per frame)
    int wf_grid[4] = {0, 0, 0, 0};          // trace kernel grids: [F != 0][W]
};

struct SharedFb {          // framebuffer + unit counter reachable by every GPU / rank
    uint8_t* base = nullptr;     // [counter 0 (128 B)] [counter 1 (128 B)] [fb 0 floats] ([fb 1 floats]: cross-process sharing only)
    size_t bytes = 0;
    int width = 0, height = 0;
    bool owner = false, ipc_mapped = false;
    cudaStream_t side = nullptr; // owner: zeroes the idle half while the other one is rendered into (rtw_render_shared_epoch)
    size_t fb_bytes() const { return (((size_t)width * height * 3 * sizeof(float)) + 255) / 256 * 256; }
    unsigned int* counter(int half = 0) const { return reinterpret_cast<unsigned int*>(base + 128 * half); }
    float* fb(int half = 0) const { return reinterpret_cast<float*>(base + 256 + fb_bytes() * half); }
};

}  // namespace

struct rtw_scene {
    rtw::SceneGraph g;
    rtw::FlatScene flat;
    bool committed = false;
    int n_prims_dev = 0, n_nodes_dev = 0;   // sizes when the BVH was built on the device (flat.prims then holds only its input)
    double ms_build_dev = 0;
    std::vector<Replica> reps;
    SharedFb local;        // used by rtw_render (in-process, on reps[0].device)
    SharedFb shared;       // used by rtw_render_shared (cross-process)
    double ms_commit = 0;
    uint64_t h2d_commit = 0;
};

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void free_replicas(rtw_scene* s) {
    for (Replica& r : s->reps) {
        cudaSetDevice(r.device);
        if (r.blob) cudaFree(r.blob);
        if (r.built.prims) rtwb::trim_scratch(r.device);        // (the device builder's scratch pool, bvh_build.cu)
        rtwb::free_output(r.built);
        if (r.wf_mem) cudaFree(r.wf_mem);
        if (r.wf_fb) cudaFree(r.wf_fb);
        for (int k = 1; k < 4; ++k) if (r.wf_stream[k]) cudaStreamDestroy(r.wf_stream[k]);
        for (int k = 0; k < 4; ++k) if (r.wf_ev[k]) cudaEventDestroy(r.wf_ev[k]);
        if (r.stats) cudaFree(r.stats);
        if (r.ev0) cudaEventDestroy(r.ev0);
        if (r.ev1) cudaEventDestroy(r.ev1);
        if (r.stream) cudaStreamDestroy(r.stream);
    }
    s->reps.clear();
    if (s->local.base) { cudaFree(s->local.base); s->local = SharedFb(); }
}

// pack FlatScene into one host blob; offsets -> DScene (device pointers filled per replica)
struct Packed { std::vector<uint8_t> bytes; size_t total = 0, o_nodes, o_wnodes, o_prims, o_xf, o_media, o_mats, o_texs, o_perlin, o_image; };
// Up to this size the scene is staged in one host blob and uploaded with ONE copy (C1: 100 KB, latency matters);
// beyond it (the 1 M - 16 M sphere sweep: 0.15 - 2.4 GB) zero-filling and filling a second host copy costs ~1 s.
static const size_t kBlobStageLimit = 64u << 20;
void pack(const rtw::FlatScene& f, Packed& p) {
    size_t off = 0;
    auto place = [&](size_t n) { size_t o = off; off = align_up(off + std::max<size_t>(n, 16), 256); return o; };
    p.o_nodes = place(f.nodes.size() * sizeof(DNode));
    p.o_wnodes = place(f.wnodes.size() * sizeof(DWNode));
    p.o_prims = place(f.prims.size() * sizeof(DPrim));
    p.o_xf = place(f.xforms.size() * sizeof(DXform));
    p.o_media = place(f.media.size() * sizeof(DMedium));
    p.o_mats = place(f.mats.size() * sizeof(DMat));
    p.o_texs = place(f.texs.size() * sizeof(DTex));
    p.o_perlin = place(f.perlin.size());
    p.o_image = place(f.image.size());
    p.total = off;
    if (off > kBlobStageLimit) return;        // big scene: sections are uploaded straight from their vectors (upload())
    p.bytes.assign(off, 0);
    auto cp = [&](size_t o, const void* src, size_t n) { if (n) std::memcpy(p.bytes.data() + o, src, n); };
    cp(p.o_nodes, f.nodes.data(), f.nodes.size() * sizeof(DNode));
    cp(p.o_wnodes, f.wnodes.data(), f.wnodes.size() * sizeof(DWNode));
    cp(p.o_prims, f.prims.data(), f.prims.size() * sizeof(DPrim));
    cp(p.o_xf, f.xforms.data(), f.xforms.size() * sizeof(DXform));
    cp(p.o_media, f.media.data(), f.media.size() * sizeof(DMedium));
    cp(p.o_mats, f.mats.data(), f.mats.size() * sizeof(DMat));
    cp(p.o_texs, f.texs.data(), f.texs.size() * sizeof(DTex));
    cp(p.o_perlin, f.perlin.data(), f.perlin.size());
    cp(p.o_image, f.image.data(), f.image.size());
}
cudaError_t upload(const rtw::FlatScene& f, const Packed& p, uint8_t* dst, cudaStream_t st) {
    if (!p.bytes.empty()) return cudaMemcpyAsync(dst, p.bytes.data(), p.bytes.size(), cudaMemcpyHostToDevice, st);
    cudaError_t e = cudaSuccess;
    auto cp = [&](size_t o, const void* src, size_t n) { if (n && e == cudaSuccess) e = cudaMemcpyAsync(dst + o, src, n, cudaMemcpyHostToDevice, st); };
    cp(p.o_nodes, f.nodes.data(), f.nodes.size() * sizeof(DNode));
    cp(p.o_wnodes, f.wnodes.data(), f.wnodes.size() * sizeof(DWNode));
    cp(p.o_prims, f.prims.data(), f.prims.size() * sizeof(DPrim));
    cp(p.o_xf, f.xforms.data(), f.xforms.size() * sizeof(DXform));
    cp(p.o_media, f.media.data(), f.media.size() * sizeof(DMedium));
    cp(p.o_mats, f.mats.data(), f.mats.size() * sizeof(DMat));
    cp(p.o_texs, f.texs.data(), f.texs.size() * sizeof(DTex));
    cp(p.o_perlin, f.perlin.data(), f.perlin.size());
    cp(p.o_image, f.image.data(), f.image.size());
    return e;
}
DScene bind(const rtw::FlatScene& f, const Packed& p, uint8_t* base) {
    DScene d;
    d.nodes = f.wide ? nullptr : reinterpret_cast<const DNode*>(base + p.o_nodes);
    d.wnodes = f.wide ? reinterpret_cast<const DWNode*>(base + p.o_wnodes) : nullptr;
    d.prims = reinterpret_cast<const DPrim*>(base + p.o_prims);
    d.xforms = reinterpret_cast<const DXform*>(base + p.o_xf);
    d.media = reinterpret_cast<const DMedium*>(base + p.o_media);
    d.mats = reinterpret_cast<const DMat*>(base + p.o_mats);
    d.texs = reinterpret_cast<const DTex*>(base + p.o_texs);
    d.perlin = base + p.o_perlin;
    d.image = base + p.o_image;
    d.n_nodes = (int)f.nodes.size(); d.n_prims = (int)f.prims.size(); d.n_bvh_prims = f.n_bvh_prims;
    d.n_xforms = (int)f.xforms.size(); d.n_media = (int)f.media.size(); d.n_mats = (int)f.mats.size();
    d.n_texs = (int)f.texs.size(); d.n_perlin = (int)(f.perlin.size() / RTW_PERLIN_BYTES);
    return d;
}

DCamera to_dcamera(const rtw_camera& c) {
    DCamera d; std::memset(&d, 0, sizeof(d));
    d.ox = (float)c.origin[0]; d.oy = (float)c.origin[1]; d.oz = (float)c.origin[2]; d.lens_radius = (float)c.lens_radius;
    d.lx = (float)(c.lower_left_corner[0] - c.origin[0]); d.ly = (float)(c.lower_left_corner[1] - c.origin[1]);
    d.lz = (float)(c.lower_left_corner[2] - c.origin[2]);
    d.hx = (float)c.horizontal[0]; d.hy = (float)c.horizontal[1]; d.hz = (float)c.horizontal[2];
    d.vx = (float)c.vertical[0]; d.vy = (float)c.vertical[1]; d.vz = (float)c.vertical[2];
    d.ux = (float)c.u[0]; d.uy = (float)c.u[1]; d.uz = (float)c.u[2];
    d.wx = (float)c.v[0]; d.wy = (float)c.v[1]; d.wz = (float)c.v[2];
    d.time0 = (float)c.time0; d.time1 = (float)c.time1;
    return d;
}

int make_params(const rtw_render_params& p, int total_warps, DParams& d) {
    if (p.width < 2 || p.height < 2 || p.spp < 1 || p.max_depth < 0) return fail(RTW_ERR_INVALID_ARG, "bad render params");
    // a queued ray packs pixel(5) | orphan(1) | segment(6) | sample(20) into one word (the pool kernel: sample(17), checked at launch)
    if (p.spp > (1 << 20) || p.max_depth > 63) return fail(RTW_ERR_INVALID_ARG, "spp <= 1048576 and max_depth <= 63");
    if (!fin_all(p.background[0], p.background[1], p.background[2], p.t_min)) return fail(RTW_ERR_INVALID_ARG, "non-finite background / t_min");
    std::memset(&d, 0, sizeof(d));
    d.width = p.width; d.height = p.height; d.spp = p.spp; d.max_depth = p.max_depth;
    d.bg_r = (float)p.background[0]; d.bg_g = (float)p.background[1]; d.bg_b = (float)p.background[2];
    d.t_min = (float)p.t_min;
    d.seed_lo = (uint32_t)p.seed; d.seed_hi = (uint32_t)(p.seed >> 32);
    for (int k = 0; k < 10; ++k) { d.philox_rk[2 * k] = d.seed_lo + 0x9E3779B9u * (uint32_t)k; d.philox_rk[2 * k + 1] = d.seed_hi + 0xBB67AE85u * (uint32_t)k; }
    d.tiles_x = (p.width + 7) / 8; d.tiles_y = (p.height + 3) / 4;
    long long tiles = (long long)d.tiles_x * d.tiles_y;
    int chunk = p.samples_per_unit;
    d.spp_a = p.spp;
    if (chunk <= 0) {
        // Guided self-scheduling in two phases.  A: 80 % of the samples in units sized for ~24 per resident warp
        // (>= 32 samples: per-unit cost = tile list + drain of the ring; much bigger units lose to the spread of tile
        // costs, sky vs glass: 6 per warp measured 13 % slower).  B: the rest in units of 32 (16) samples, so the frame
        // ends on SHORT units.  Measured C1 / final_scene: one size, 16 per warp 110.1 / 572 ms; this 106.7 / 535 ms.
        // RTW_UNITS_PER_WARP / RTW_ONE_PHASE: tuning switches for tools/ab2.sh.
        static const bool one_phase = std::getenv("RTW_ONE_PHASE") != nullptr;
        static const char* wu = std::getenv("RTW_UNITS_PER_WARP");
        static const char* bspp = getenv("RTW_B_SPP");          // tuning (tools/ab_env.sh): phase-B unit size, share in %, phase-A minimum
        static const char* bshare = getenv("RTW_B_SHARE");
        static const char* amin = getenv("RTW_A_MIN");
        int spp_b = (p.spp >= 160 && !one_phase) ? (int)((long long)p.spp * (bshare ? atoi(bshare) : 20) / 100) : 0;
        d.spp_a = p.spp - spp_b;
        long long want_units = (wu ? atoll(wu) : 24LL) * total_warps;
        if (const char* e = getenv("RTW_EMULATE_RANKS")) want_units *= std::max(1, atoi(e));
        long long chunks = (want_units + tiles - 1) / tiles;
        if (chunks < 1) chunks = 1;
        chunk = (int)((d.spp_a + chunks - 1) / chunks);
        const int a_min = amin ? std::max(1, atoi(amin)) : 32;
        if (chunk < a_min) {
            chunk = a_min;
            // phase A at its floor = the many-GPU regime (a frame is ~14 ms): with the carry-over short units are cheap, and 30 %
            // of the samples in them ends the frame 0.2 ms earlier than 20 % (emulated 8 ranks: 13.24 vs 13.47 ms, r2_ad_carry_units.log)
            // (only when several GPUs share the frame: a SMALL frame on one GPU reaches the floor too, and there the extra short
            // units cost — two_perlin_spheres 800x450x200 on one GPU 12.6 -> 13.8 ms)
            const bool many = p.n_gpus > 1 || getenv("RTW_EMULATE_RANKS") != nullptr;
            if (spp_b > 0 && !bshare && many) { spp_b = (int)((long long)p.spp * 30 / 100); d.spp_a = p.spp - spp_b; }
        }
        if (spp_b > 0) {
            // (8 when phase A already sits at its floor — the 8-GPU regime: the frame then ends ~0.3 ms earlier, measured with
            // RTW_EMULATE_RANKS=8, profiles/r2_f_units.log; a warp needs ~1 ms of wall time per 32-sample unit at full residency)
            d.chunk_spp_b = bspp ? std::max(1, atoi(bspp)) : (chunk >= 64 ? 32 : (chunk > a_min ? 16 : 8));
            d.chunks_b = (spp_b + d.chunk_spp_b - 1) / d.chunk_spp_b;
        }
    }
    if (chunk > d.spp_a) chunk = d.spp_a;
    d.chunk_spp = chunk;
    d.chunks = (d.spp_a + chunk - 1) / chunk;
    d.n_units_a = (uint32_t)(tiles * d.chunks);
    long long n_units = tiles * ((long long)d.chunks + d.chunks_b);
    if (n_units >= 0xffffffffLL) return fail(RTW_ERR_INVALID_ARG, "too many work units");
    d.n_units = (uint32_t)n_units;
    d.accumulate = 1;
    d.no_tile_cull = (p.flags & RTW_FLAG_NO_TILE_CULL) ? 1 : 0;
    d.list_max = 8;      // measured: cornell_box (8 leaves, closed room) -17 % (r2_ah_list_scan.log); sparse spheres break even at 6-8, lose 8 % at 12 (r2_am_listmax_crossover.log)
    if (const char* e = getenv("RTW_LIST_MAX")) d.list_max = std::max(0, atoi(e));
    d.unit_stride = 1;
    if (const char* e = getenv("RTW_EMULATE_RANKS")) d.unit_stride = (uint32_t)std::max(1, atoi(e));   // tuning aid: image incomplete
    return 0;
}

int ensure_fb(SharedFb& fb, int device, int w, int h) {
    size_t need = 256 + (size_t)w * h * 3 * sizeof(float);
    if (fb.base && fb.bytes >= need) { fb.width = w; fb.height = h; return 0; }
    CUDA_TRY(cudaSetDevice(device));
    if (fb.base) { cudaFree(fb.base); fb.base = nullptr; }
    CUDA_TRY(cudaMalloc(&fb.base, need));
    fb.bytes = need; fb.width = w; fb.height = h; fb.owner = true;
    return 0;
}

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// kernel choice: 0 = megakernel (one path per lane), 1..4 = warp-pool kernel with 64/128/192/256 slots per warp, 5 = wavefront
// pipeline (path pool in HBM).  Chosen by measurement: the megakernel wins wherever the scene
// sits in L1 / L2 (C1 - C4: the wavefront pipeline is 2x slower there), the wavefront pipeline over the 8-wide tree wins on
// scenes beyond the caches (256 Ki / 512 Ki / 1 M / 4 M / 16 M spheres: 1.24x / 1.24x / 1.30x / 1.38x / 1.52x); the switch
// sits at 256 Ki primitives, the smallest size measured.  Flags and RTW_KERNEL override.
long long big_scene_min() {
    static long long v = -1;
    if (v < 0) { v = 1ll << 18; if (const char* e = getenv("RTW_BIG_MIN")) v = std::max(1ll, atoll(e)); }
    return v;
}
int kernel_mode(const rtw_scene* s, int flags) {
    static int env_mode = -2;
    if (env_mode == -2) {
        env_mode = -1;
        if (const char* e = getenv("RTW_KERNEL")) {
            if (!strcmp(e, "mega")) env_mode = 0;
            else if (!strcmp(e, "pool64")) env_mode = 1;
            else if (!strcmp(e, "pool128")) env_mode = 2;
            else if (!strcmp(e, "pool192")) env_mode = 3;
            else if (!strcmp(e, "pool256")) env_mode = 4;
            else if (!strcmp(e, "wavefront")) env_mode = 5;
        }
    }
    if (flags & RTW_FLAG_KERNEL_WAVEFRONT) return 5;
    if (flags & RTW_FLAG_KERNEL_MEGA) return 0;
    if (flags & RTW_FLAG_KERNEL_POOL) return RTW_DEFAULT_POOL_MODE;
    if (env_mode >= 0) return env_mode;
    if (s && s->flat.wide && std::max<long long>(s->flat.n_bvh_prims, s->n_prims_dev) >= big_scene_min()) return 5;
    return RTW_DEFAULT_MODE;
}

template <int POOL, int W>
int launch_pool_w(Replica& r, int slot, const DCamera& dc, const DParams& dp, unsigned int* counter, float* fb) {
    const size_t smem = RTW_WARPS * sizeof(PoolSmem<POOL>);
    if (!r.pool_grid[slot]) {
        CUDA_TRY(cudaFuncSetAttribute(render_pool_kernel<POOL, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_pool_kernel<POOL, W>, RTW_BLOCK, smem));
        if (per_sm < 1) return fail(RTW_ERR_CUDA, "warp-pool kernel does not fit on an SM");
        r.pool_grid[slot] = r.sms * per_sm;
    }
    render_pool_kernel<POOL, W><<<r.pool_grid[slot], RTW_BLOCK, smem, r.stream>>>(r.ds, dc, dp, counter, fb, r.stats);
    return 0;
}
template <int POOL>
int launch_pool(Replica& r, int slot, bool wide, const DCamera& dc, const DParams& dp, unsigned int* counter, float* fb) {
    return wide ? launch_pool_w<POOL, 1>(r, slot + 4, dc, dp, counter, fb) : launch_pool_w<POOL, 0>(r, slot, dc, dp, counter, fb);
}

// ---- wavefront pipeline (rtw_wavefront.cuh): pool allocation + the iteration loop -------------------------------------
int wf_ensure_pool(Replica& r, long long want_slots) {
    long long P = 8ll << 20;                                         // slots in flight per GPU (RTW_WF_POOL overrides)
    if (const char* e = getenv("RTW_WF_POOL")) P = std::max(1024ll, atoll(e));
    P = std::min(P, std::max(1024ll, want_slots));
    // K independent pools, each with its own stream: the logic pass and the end of the trace pass of one pool overlap with
    // the trace pass of another (the trace kernel is persistent: its last rays leave most of the GPU idle)
    int K = RTW_WF_DEFAULT_POOLS;
    if (const char* e = getenv("RTW_WF_POOLS")) K = std::min(4, std::max(1, atoi(e)));
    if (P < (1ll << 20)) K = 1;
    r.wf_stream[0] = r.stream;
    const long long per = ((P + K - 1) / K + RTW_WF_BLOCK - 1) / RTW_WF_BLOCK * RTW_WF_BLOCK;
    if (r.wf_mem && r.wf_slots == per && r.wf_pools == K) return 0;
    if (r.wf_mem) { CUDA_TRY(cudaFree(r.wf_mem)); r.wf_mem = nullptr; }
    const int max_chunks = 1 << 16;                                  // 2^16 chunks x 2^20 paths: more than spp <= 2^20 at 8K needs
    const size_t per_slot = 4 * sizeof(float4) + 2 * sizeof(uint2);
    const size_t per_pool = align_up((size_t)per * per_slot + 256 + (size_t)max_chunks * 8 + 3 * 8192 * 8, 256);
    CUDA_TRY(cudaMalloc(&r.wf_mem, per_pool * K));
    for (int k = 0; k < K; ++k) {
        uint8_t* p = r.wf_mem + per_pool * k;
        WfPool& w = r.wf[k];
        w.od0 = reinterpret_cast<float4*>(p); p += (size_t)per * 16;
        w.od1 = reinterpret_cast<float4*>(p); p += (size_t)per * 16;
        w.thr = reinterpret_cast<float4*>(p); p += (size_t)per * 16;
        w.rad = reinterpret_cast<float4*>(p); p += (size_t)per * 16;
        w.id = reinterpret_cast<uint2*>(p); p += (size_t)per * 8;
        w.hit = reinterpret_cast<uint2*>(p); p += (size_t)per * 8;
        w.ctr = reinterpret_cast<unsigned long long*>(p); p += 256;
        w.chunk_base = reinterpret_cast<unsigned long long*>(p); p += (size_t)max_chunks * 8;
        w.dbg = reinterpret_cast<unsigned long long*>(p);
        w.P = (int)per; w.max_chunks = max_chunks;
        if (k && !r.wf_stream[k]) CUDA_TRY(cudaStreamCreateWithFlags(&r.wf_stream[k], cudaStreamNonBlocking));
        if (!r.wf_ev[k]) CUDA_TRY(cudaEventCreateWithFlags(&r.wf_ev[k], cudaEventDisableTiming));
    }
    r.wf_pools = K; r.wf_slots = per;
    return 0;
}

template <int F, int W>
int wf_run(rtw_scene* s, Replica& r, const DCamera& dc, const DParams& dp, unsigned long long* path_counter, unsigned long long total_paths, float* fb) {
    static const bool lockstep = !(getenv("RTW_WF_TRACE") && !strcmp(getenv("RTW_WF_TRACE"), "whilewhile"));
    // the trace stage: lockstep loop (default) or the speculative while-while loop, over binary or 8-wide nodes
    void (*trace)(const DScene, const DParams, WfPool) = lockstep ? (W ? wf_trace2w_kernel<F> : wf_trace2_kernel<F>) : wf_trace_kernel<F, W>;
    int& grid = r.wf_grid[(F ? 2 : 0) + W];
    if (!grid) {
        int per_sm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace, 128, 0));
        grid = r.sms * std::max(per_sm, 1);
    }
    const int K = r.wf_pools;
    const bool timing = getenv("RTW_TIMING") != nullptr;
    CUDA_TRY(cudaEventRecord(r.wf_ev[0], r.stream));                                        // the pools start after whatever precedes on the replica's stream
    for (int k = 0; k < K; ++k) {
        const WfPool& pool = r.wf[k];
        if (k) CUDA_TRY(cudaStreamWaitEvent(r.wf_stream[k], r.wf_ev[0], 0));
        CUDA_TRY(cudaMemsetAsync(pool.rad, 0, (size_t)pool.P * sizeof(float4), r.wf_stream[k]));     // every slot EMPTY
        CUDA_TRY(cudaMemsetAsync(pool.ctr, 0, 256, r.wf_stream[k]));
    }
    unsigned long long h[4][8];
    bool done[4] = {false, false, false, false};
    const double t_begin = now_ms();
    for (long long it = 0;; ++it) {
        // a batch of iterations per pool, then one look at the counters (the only host round trip)
        for (int j = 0; j < 8; ++j)
            for (int k = 0; k < K; ++k) {
                if (done[k]) continue;
                const WfPool& pool = r.wf[k];
                wf_reserve_kernel<<<1, 1, 0, r.wf_stream[k]>>>(pool, path_counter, total_paths);
                wf_logic_kernel<F><<<pool.P / RTW_WF_BLOCK, RTW_WF_BLOCK, 0, r.wf_stream[k]>>>(r.ds, dc, dp, pool, total_paths, fb);
                trace<<<grid, 128, 0, r.wf_stream[k]>>>(r.ds, dp, pool);
            }
        CUDA_TRY(cudaGetLastError());
        for (int k = 0; k < K; ++k) if (!done[k]) CUDA_TRY(cudaMemcpyAsync(h[k], r.wf[k].ctr, sizeof(h[k]), cudaMemcpyDeviceToHost, r.wf_stream[k]));
        for (int k = 0; k < K; ++k) if (!done[k]) CUDA_TRY(cudaStreamSynchronize(r.wf_stream[k]));
#ifdef RTW_INSTRUMENT
        if (getenv("RTW_WF_TIMELINE") && it < 4) {
            const WfPool& pool = r.wf[0];
            std::vector<unsigned long long> d(3 * 8192);
            cudaMemcpy(d.data(), pool.dbg, d.size() * 8, cudaMemcpyDeviceToHost);
            const int nw = std::min(8192, grid * 4);
            unsigned long long t0 = ~0ull, t1 = 0; for (int w = 0; w < nw; ++w) { t0 = std::min(t0, d[w]); t1 = std::max(t1, d[8192 + w]); }
            std::vector<double> ends; unsigned long long rmin = ~0ull, rmax = 0, rsum = 0;
            for (int w = 0; w < nw; ++w) { ends.push_back((d[8192 + w] - t0) * 1e-6); rmin = std::min(rmin, d[16384 + w]); rmax = std::max(rmax, d[16384 + w]); rsum += d[16384 + w]; }
            std::sort(ends.begin(), ends.end());
            fprintf(stderr, "[wf timeline] launch after %lld iterations: %d warps, end times ms: min %.3f p10 %.3f p50 %.3f p90 %.3f p99 %.3f max %.3f; rays per warp min %llu mean %.0f max %llu\n",
                    (it + 1) * 8, nw, ends.front(), ends[nw / 10], ends[nw / 2], ends[nw * 9 / 10], ends[nw * 99 / 100], ends.back(), rmin, (double)rsum / nw, rmax);
        }
#endif
        // a pool is done when none of its slots holds a ray, the shared counter is exhausted and every reserved path number was handed out
        bool all = true;
        for (int k = 0; k < K; ++k) {
            if (!done[k] && h[k][1] == 0 && h[k][6] && h[k][0] >= (h[k][4] << RTW_WF_CHUNK_LOG2)) done[k] = true;
            all = all && done[k];
        }
        if (timing && (it < 3 || all)) {
            unsigned long long rays = 0, paths = 0; for (int k = 0; k < K; ++k) { rays += h[k][3]; paths += h[k][5]; }
            fprintf(stderr, "[wavefront] after %lld iterations (%d pools of %d): %.2f ms, %llu rays, %llu paths%s\n", (it + 1) * 8, K, r.wf[0].P, now_ms() - t_begin, rays, paths, all ? " (done)" : "");
        }
        if (all) break;
        if (it > (1ll << 24)) return fail(RTW_ERR_CUDA, "wavefront loop did not terminate");
    }
    for (int k = 1; k < K; ++k) {                                                            // whatever follows on the replica's stream follows every pool
        CUDA_TRY(cudaEventRecord(r.wf_ev[k], r.wf_stream[k]));
        CUDA_TRY(cudaStreamWaitEvent(r.stream, r.wf_ev[k], 0));
    }
    (void)s;
    return 0;
}

int wf_render_replica(rtw_scene* s, Replica& r, const DCamera& dc, const DParams& dp, unsigned int* counter, float* fb, double& ms, uint64_t& rays, uint64_t& paths) {
    CUDA_TRY(cudaSetDevice(r.device));
    const unsigned long long total = (unsigned long long)dp.tiles_x * dp.tiles_y * 32ull * (unsigned long long)dp.spp;
    TRY(wf_ensure_pool(r, (long long)std::min<unsigned long long>(total, 1ull << 40)));
    unsigned long long* path_counter = reinterpret_cast<unsigned long long*>(counter) + 1;     // second word of the zeroed header
    // the framebuffer of another GPU (in-process peer or IPC mapping): accumulate locally, merge once at the end of the frame
    float* target = fb;
    const size_t n_floats = (size_t)dp.width * dp.height * 3;
    {
        cudaPointerAttributes at{};
        bool remote = cudaPointerGetAttributes(&at, fb) == cudaSuccess && at.type == cudaMemoryTypeDevice && at.device != r.device;
        cudaGetLastError();
        // (a framebuffer opened through CUDA IPC reports the IMPORTING device: one process per GPU is told apart by the mapping itself)
        const uint8_t* fbb = reinterpret_cast<const uint8_t*>(fb);
        if (s->shared.ipc_mapped && s->shared.base && fbb >= s->shared.base && fbb < s->shared.base + s->shared.bytes) remote = true;
        if (remote) {
            if (r.wf_fb_floats < n_floats) {
                if (r.wf_fb) { CUDA_TRY(cudaFree(r.wf_fb)); r.wf_fb = nullptr; }
                CUDA_TRY(cudaMalloc(&r.wf_fb, n_floats * sizeof(float))); r.wf_fb_floats = n_floats;
                CUDA_TRY(cudaMemsetAsync(r.wf_fb, 0, n_floats * sizeof(float), r.stream));
            }
            target = r.wf_fb;
        }
    }
    CUDA_TRY(cudaEventRecord(r.ev0, r.stream));
    const bool plain = s->flat.features == 0;
    int rc;
    if (s->flat.wide) rc = plain ? wf_run<0, 1>(s, r, dc, dp, path_counter, total, target) : wf_run<FEAT_ALL, 1>(s, r, dc, dp, path_counter, total, target);
    else rc = plain ? wf_run<0, 0>(s, r, dc, dp, path_counter, total, target) : wf_run<FEAT_ALL, 0>(s, r, dc, dp, path_counter, total, target);
    if (rc < 0) { if (target != fb) cudaMemsetAsync(target, 0, n_floats * sizeof(float), r.stream); return rc; }     // (the local framebuffer is always left zeroed)
    const double t_merge0 = now_ms();
    if (target != fb) {
        if (getenv("RTW_TIMING")) CUDA_TRY(cudaStreamSynchronize(r.stream));
        wf_merge_kernel<<<r.sms * 8, 256, 0, r.stream>>>(target, fb, n_floats);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(r.ev1, r.stream));
    CUDA_TRY(cudaStreamSynchronize(r.stream));
    if (target != fb && getenv("RTW_TIMING")) fprintf(stderr, "[wavefront] merge of the local framebuffer: %.2f ms\n", now_ms() - t_merge0);
    float e = 0; CUDA_TRY(cudaEventElapsedTime(&e, r.ev0, r.ev1));
    ms = e; rays = 0; paths = 0;
    for (int k = 0; k < r.wf_pools; ++k) {
        unsigned long long h[8];
        CUDA_TRY(cudaMemcpy(h, r.wf[k].ctr, sizeof(h), cudaMemcpyDeviceToHost));
        rays += h[3]; paths += h[5];
    }
    return 0;
}

int launch_wavefront(rtw_scene* s, int n_rep, const rtw_camera* cam, const DParams& dp, unsigned int* counter, float* fb, rtw_stats* st) {
    const DCamera dc = to_dcamera(*cam);
    std::vector<double> ms(n_rep, 0.0); std::vector<uint64_t> rays(n_rep, 0), paths(n_rep, 0); std::vector<int> rc(n_rep, 0);
    std::vector<std::string> errs(n_rep);
    if (n_rep == 1) rc[0] = wf_render_replica(s, s->reps[0], dc, dp, counter, fb, ms[0], rays[0], paths[0]);
    else {                                       // one host thread per GPU: each drives its own iteration loop
        std::vector<std::thread> th;
        for (int i = 0; i < n_rep; ++i)
            th.emplace_back([&, i]() { rc[i] = wf_render_replica(s, s->reps[i], dc, dp, counter, fb, ms[i], rays[i], paths[i]); if (rc[i] < 0) errs[i] = g_err; });
        for (auto& t : th) t.join();
    }
    for (int i = 0; i < n_rep; ++i) if (rc[i] < 0) return fail(rc[i], errs[i].empty() ? g_err : errs[i]);
    if (st) {
        double m = 0; uint64_t r = 0;
        for (int i = 0; i < n_rep; ++i) { m = std::max(m, ms[i]); r += rays[i]; if (i < 8) st->units_per_device[i] = paths[i]; }
        st->ms_render = m; st->rays = r; st->n_devices = n_rep;
        st->kernel_launches += n_rep;       // (three small kernels per iteration; counted as one render per device)
    }
    return 0;
}

// Launch the render kernel on replicas [0, n) against (counter, fb); sync; fill stats.
// Everything a launch needs (scene pointers, camera, the Philox key schedule) travels in its kernel parameters: renders
// on different scene handles share no mutable state and may run concurrently from different host threads.
int launch_all(rtw_scene* s, int n_rep, const rtw_camera* cam, const DParams& dp, unsigned int* counter, float* fb, rtw_stats* st, int mode) {
    // MovingSphere boxes in the BVH cover the spheres' own [time0, time1] (as the reference's BvhNode boxes do); a ray
    // time outside that interval could be culled where the reference's flat world list would still hit: refuse it.
    if (std::min(cam->time0, cam->time1) < s->flat.mov_t0 || std::max(cam->time0, cam->time1) > s->flat.mov_t1)
        return fail(RTW_ERR_INVALID_ARG, "camera shutter [time0, time1] exceeds the [time0, time1] of a MovingSphere in the scene");
    if (mode == 5) return launch_wavefront(s, n_rep, cam, dp, counter, fb, st);
    DCamera dc = to_dcamera(*cam);
    for (int i = 0; i < n_rep; ++i) {
        Replica& r = s->reps[i];
        CUDA_TRY(cudaSetDevice(r.device));
        CUDA_TRY(cudaMemsetAsync(r.stats, 0, 32, r.stream));
        CUDA_TRY(cudaEventRecord(r.ev0, r.stream));
        if (mode >= 1 && dp.first_sample + dp.spp > (1 << 17)) return fail(RTW_ERR_INVALID_ARG, "the pool kernel packs the sample index into 17 bits: spp <= 131072");
        if (mode >= 1 && s->flat.media.size() > 15) return fail(RTW_ERR_INVALID_ARG, "the pool kernel packs the media draw count into 4 bits: at most 15 media");
        switch (mode) {
        case 1: TRY(launch_pool<64>(r, 0, s->flat.wide, dc, dp, counter, fb)); break;
        case 2: TRY(launch_pool<128>(r, 1, s->flat.wide, dc, dp, counter, fb)); break;
        case 3: TRY(launch_pool<192>(r, 2, s->flat.wide, dc, dp, counter, fb)); break;
        case 4: TRY(launch_pool<256>(r, 3, s->flat.wide, dc, dp, counter, fb)); break;
        default: {
            // smallest kernel variant that covers the scene's features (code size = instruction-cache pressure)
            const int f = s->flat.features;
            const bool small = !s->flat.wide && s->flat.n_bvh_prims <= dp.list_max;       // a handful of primitives: scan, no BVH walk
#define RTW_TRY_VARIANT(V) if ((f & ~(V)) == 0) { \
                if (s->flat.wide) render_kernel<V, 1, 0><<<r.grid, RTW_BLOCK, 0, r.stream>>>(r.ds, dc, dp, counter, fb, r.stats); \
                else if (small) render_kernel<V, 0, 1><<<r.grid, RTW_BLOCK, 0, r.stream>>>(r.ds, dc, dp, counter, fb, r.stats); \
                else render_kernel<V, 0, 0><<<r.grid, RTW_BLOCK, 0, r.stream>>>(r.ds, dc, dp, counter, fb, r.stats); \
                break; }
            RTW_TRY_VARIANT(0)                                         // spheres, solid / checker         (C1, two_spheres)
            RTW_TRY_VARIANT(FEAT_NOISE)                                // + Perlin                         (two_perlin_spheres)
            RTW_TRY_VARIANT(FEAT_IMAGE)                                // + image                          (earth)
            RTW_TRY_VARIANT(FEAT_RECT | FEAT_XFORM | FEAT_RXFORM)                // rects, boxes, instances (cornell_box)
            RTW_TRY_VARIANT(FEAT_RECT | FEAT_XFORM | FEAT_RXFORM | FEAT_NOISE)   //                         (simple_light)
            RTW_TRY_VARIANT(FEAT_RECT | FEAT_XFORM | FEAT_RXFORM | FEAT_MEDIA)   // + media                 (cornell_box_smoke)
            RTW_TRY_VARIANT(FEAT_ALL & ~FEAT_RXFORM)                             // all, rects not instanced (final_scene)
            RTW_TRY_VARIANT(FEAT_ALL)
#undef RTW_TRY_VARIANT
        }
        }
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(r.ev1, r.stream));
    }
    double ms_max = 0; uint64_t rays = 0;
    for (int i = 0; i < n_rep; ++i) {
        Replica& r = s->reps[i];
        CUDA_TRY(cudaSetDevice(r.device));
        CUDA_TRY(cudaStreamSynchronize(r.stream));
        float ms = 0; CUDA_TRY(cudaEventElapsedTime(&ms, r.ev0, r.ev1));
        if (ms > ms_max) ms_max = ms;
        unsigned long long h[4];
        CUDA_TRY(cudaMemcpy(h, r.stats, 32, cudaMemcpyDeviceToHost));
        rays += h[0];
        if (getenv("RTW_DEBUG_STATS") && h[1]) fprintf(stderr, "[rtw] device %d: units %llu, mean tile-list length %.2f, list overflows %llu\n", r.device, h[1], (double)h[2] / (double)h[1], h[3]);
        if (st && i < 8) st->units_per_device[i] = h[1];
    }
    if (st) { st->ms_render = ms_max; st->rays = rays; st->kernel_launches += n_rep; st->n_devices = n_rep; }
    return 0;
}

void fill_scene_stats(rtw_scene* s, rtw_stats* st) {
    st->n_prims = (int)s->flat.prims.size(); st->n_nodes = (int)(s->flat.wide ? s->flat.wnodes.size() : s->flat.nodes.size());
    if (s->flat.emit_only) { st->n_prims = s->n_prims_dev; st->n_nodes = s->n_nodes_dev; }
    st->n_materials = (int)s->flat.mats.size(); st->n_media = (int)s->flat.media.size();
    st->ms_commit = s->ms_commit;
}

// scratch device buffers for the parity hooks
struct Scratch {
    std::vector<void*> ptrs;
    ~Scratch() { for (void* p : ptrs) cudaFree(p); }
    template <class T> int up(const T* host, size_t n, T*& dev) {
        dev = nullptr;
        cudaError_t e = cudaMalloc(&dev, std::max<size_t>(n, 1) * sizeof(T));
        if (e != cudaSuccess) return fail(RTW_ERR_OOM, cudaGetErrorString(e));
        ptrs.push_back(dev);
        if (host && n) { e = cudaMemcpy(dev, host, n * sizeof(T), cudaMemcpyHostToDevice); if (e != cudaSuccess) return fail(RTW_ERR_CUDA, cudaGetErrorString(e)); }
        return 0;
    }
    template <class T> int down(T* host, const T* dev, size_t n) {
        cudaError_t e = cudaMemcpy(host, dev, n * sizeof(T), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return fail(RTW_ERR_CUDA, cudaGetErrorString(e));
        return 0;
    }
};

int need_device() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return fail(RTW_ERR_NO_DEVICE, "no CUDA device (librtw has no CPU fallback)"); }
    return 0;
}

bool tex_ok(rtw_scene* s, int t) { return s && t >= 0 && t < (int)s->g.textures.size(); }
bool mat_ok(rtw_scene* s, int m) { return s && m >= 1 && m <= (int)s->g.materials.size(); }
bool node_ok(rtw_scene* s, int id) { return s && id >= 0 && id < (int)s->g.nodes.size(); }
int push_node(rtw_scene* s, const rtw::HNode& h) { s->g.nodes.push_back(h); s->committed = false; return (int)s->g.nodes.size() - 1; }
int push_mat(rtw_scene* s, const rtw::HMaterial& m) { s->g.materials.push_back(m); s->committed = false; return (int)s->g.materials.size(); }
rtw::V3d v3(const double a[3]) { rtw::V3d v; v.x = a[0]; v.y = a[1]; v.z = a[2]; return v; }

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* rtw_last_error(void) { return g_err.c_str(); }
const char* rtw_version(void) { return "rtw-b200 0.1 (sm_100a)"; }
int rtw_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

rtw_scene* rtw_scene_new(void) { return new (std::nothrow) rtw_scene(); }
void rtw_scene_free(rtw_scene* s) {
    if (!s) return;
    if (s->shared.base) rtw_shared_close(s);
    free_replicas(s);
    delete s;
}

int rtw_tex_solid(rtw_scene* s, const double rgb[3]) {
    if (!s || !rgb) return fail(RTW_ERR_INVALID_ARG, "null argument");
    if (!fin3(rgb)) return fail(RTW_ERR_INVALID_ARG, "non-finite colour");
    rtw::HTexture t; t.kind = TEX_SOLID; std::memcpy(t.c0, rgb, 24); s->g.textures.push_back(t); return (int)s->g.textures.size() - 1;
}
int rtw_tex_checker(rtw_scene* s, const double even[3], const double odd[3]) {
    if (!s || !even || !odd) return fail(RTW_ERR_INVALID_ARG, "null argument");
    if (!fin3(even) || !fin3(odd)) return fail(RTW_ERR_INVALID_ARG, "non-finite colour");
    rtw::HTexture t; t.kind = TEX_CHECKER; std::memcpy(t.c0, even, 24); std::memcpy(t.c1, odd, 24);
    s->g.textures.push_back(t); return (int)s->g.textures.size() - 1;
}
int rtw_tex_noise(rtw_scene* s, const double* ranvec, const int32_t* px, const int32_t* py, const int32_t* pz, double scale) {
    if (!s || !ranvec || !px || !py || !pz) return fail(RTW_ERR_INVALID_ARG, "null argument");
    if (!fin(scale)) return fail(RTW_ERR_INVALID_ARG, "non-finite noise scale");
    for (int i = 0; i < 768; ++i) if (!fin(ranvec[i])) return fail(RTW_ERR_INVALID_ARG, "non-finite perlin gradient");
    rtw::HTexture t; t.kind = TEX_NOISE; t.scale = scale;
    t.ranvec.assign(ranvec, ranvec + 768);
    t.perm.assign(px, px + 256); t.perm.insert(t.perm.end(), py, py + 256); t.perm.insert(t.perm.end(), pz, pz + 256);
    for (int v : t.perm) if (v < 0 || v > 255) return fail(RTW_ERR_INVALID_ARG, "perlin permutation entry outside 0..255");
    s->g.textures.push_back(std::move(t)); return (int)s->g.textures.size() - 1;
}
int rtw_tex_image(rtw_scene* s, int32_t w, int32_t h, int32_t bps, const uint8_t* data) {
    if (!s || !data || w <= 0 || h <= 0 || bps < 3 * w) return fail(RTW_ERR_INVALID_ARG, "bad image texture");
    rtw::HTexture t; t.kind = TEX_IMAGE; t.w = w; t.h = h; t.bps = bps; t.data.assign(data, data + (size_t)bps * h);
    s->g.textures.push_back(std::move(t)); return (int)s->g.textures.size() - 1;
}

int rtw_mat_lambertian(rtw_scene* s, int tex) {
    if (!tex_ok(s, tex)) return fail(RTW_ERR_INVALID_ARG, "bad texture id");
    rtw::HMaterial m; m.kind = MAT_LAMBERTIAN; m.tex = tex; return push_mat(s, m);
}
int rtw_mat_metal(rtw_scene* s, const double albedo[3], double fuzz) {
    if (!s || !albedo) return fail(RTW_ERR_INVALID_ARG, "null argument");
    if (!fin3(albedo) || !fin(fuzz)) return fail(RTW_ERR_INVALID_ARG, "non-finite metal parameters");
    rtw::HMaterial m; m.kind = MAT_METAL; std::memcpy(m.albedo, albedo, 24); m.fuzz = fuzz; return push_mat(s, m);
}
int rtw_mat_dielectric(rtw_scene* s, double ir) {
    if (!s) return fail(RTW_ERR_INVALID_ARG, "null argument");
    if (!fin(ir) || ir == 0.0) return fail(RTW_ERR_INVALID_ARG, "index of refraction must be finite and non-zero");
    rtw::HMaterial m; m.kind = MAT_DIELECTRIC; m.ir = ir; return push_mat(s, m);
}
int rtw_mat_diffuse_light(rtw_scene* s, int tex) {
    if (!tex_ok(s, tex)) return fail(RTW_ERR_INVALID_ARG, "bad texture id");
    rtw::HMaterial m; m.kind = MAT_DIFFUSE_LIGHT; m.tex = tex; return push_mat(s, m);
}
int rtw_mat_isotropic(rtw_scene* s, int tex) {
    if (!tex_ok(s, tex)) return fail(RTW_ERR_INVALID_ARG, "bad texture id");
    rtw::HMaterial m; m.kind = MAT_ISOTROPIC; m.tex = tex; return push_mat(s, m);
}

int rtw_sphere(rtw_scene* s, int mat, const double c[3], double r) {
    if (!mat_ok(s, mat) || !c) return fail(RTW_ERR_INVALID_ARG, "bad material handle");
    if (!fin3(c) || !fin(r) || r == 0.0) return fail(RTW_ERR_INVALID_ARG, "sphere centre / radius must be finite, radius non-zero");
    rtw::HNode h; h.kind = rtw::H_SPHERE; h.mat = mat; h.c0 = v3(c); h.radius = r; return push_node(s, h);
}
int rtw_sphere_batch(rtw_scene* s, int32_t n, const int32_t* mats, const double* centers, const double* radii) {
    if (!s || n <= 0 || !mats || !centers || !radii) return fail(RTW_ERR_INVALID_ARG, "bad argument");
    for (int i = 0; i < n; ++i) if (!mat_ok(s, mats[i])) return fail(RTW_ERR_INVALID_ARG, "bad material handle");
    for (int i = 0; i < n; ++i)
        if (!fin_all(centers[3 * i], centers[3 * i + 1], centers[3 * i + 2], radii[i]) || radii[i] == 0.0)
            return fail(RTW_ERR_INVALID_ARG, "sphere centre / radius must be finite, radius non-zero");
    size_t base = s->g.bulk.size();
    s->g.bulk.resize(base + (size_t)n);
    for (int i = 0; i < n; ++i) {
        rtw::BulkSphere& b = s->g.bulk[base + i];
        b.c[0] = centers[3 * i]; b.c[1] = centers[3 * i + 1]; b.c[2] = centers[3 * i + 2]; b.r = radii[i]; b.mat = mats[i];
    }
    s->committed = false;
    return RTW_OK;
}
int rtw_moving_sphere(rtw_scene* s, int mat, const double c0[3], const double c1[3], double t0, double t1, double r) {
    if (!mat_ok(s, mat) || !c0 || !c1) return fail(RTW_ERR_INVALID_ARG, "bad material handle");
    if (!fin3(c0) || !fin3(c1) || !fin_all(t0, t1, r) || r == 0.0) return fail(RTW_ERR_INVALID_ARG, "moving sphere parameters must be finite, radius non-zero");
    if (t0 == t1) return fail(RTW_ERR_INVALID_ARG, "moving sphere needs time0 != time1 (centre(t) divides by time1 - time0)");
    rtw::HNode h; h.kind = rtw::H_MOVING_SPHERE; h.mat = mat; h.c0 = v3(c0); h.c1 = v3(c1); h.time0 = t0; h.time1 = t1; h.radius = r;
    return push_node(s, h);
}
static int push_rect(rtw_scene* s, int kind, int mat, double a0, double a1, double b0, double b1, double k) {
    if (!mat_ok(s, mat)) return fail(RTW_ERR_INVALID_ARG, "bad material handle");
    if (!fin_all(a0, a1, b0, b1, k)) return fail(RTW_ERR_INVALID_ARG, "non-finite rect coordinates");
    rtw::HNode h; h.kind = kind; h.mat = mat; h.a0 = a0; h.a1 = a1; h.b0 = b0; h.b1 = b1; h.k = k; return push_node(s, h);
}
int rtw_xy_rect(rtw_scene* s, int mat, double x0, double x1, double y0, double y1, double k) { return push_rect(s, rtw::H_XY, mat, x0, x1, y0, y1, k); }
int rtw_xz_rect(rtw_scene* s, int mat, double x0, double x1, double z0, double z1, double k) { return push_rect(s, rtw::H_XZ, mat, x0, x1, z0, z1, k); }
int rtw_yz_rect(rtw_scene* s, int mat, double y0, double y1, double z0, double z1, double k) { return push_rect(s, rtw::H_YZ, mat, y0, y1, z0, z1, k); }
int rtw_box(rtw_scene* s, const double mn[3], const double mx[3], int mat) {           // new_box src/hittable.rs:132-145
    if (!mat_ok(s, mat) || !mn || !mx) return fail(RTW_ERR_INVALID_ARG, "bad material handle");
    if (!fin3(mn) || !fin3(mx)) return fail(RTW_ERR_INVALID_ARG, "non-finite box corners");
    rtw::HNode h; h.kind = rtw::H_BOX; h.mat = mat; h.bmin = v3(mn); h.bmax = v3(mx);
    h.children.push_back(rtw_xy_rect(s, mat, mn[0], mx[0], mn[1], mx[1], mx[2]));
    h.children.push_back(rtw_xy_rect(s, mat, mn[0], mx[0], mn[1], mx[1], mn[2]));
    h.children.push_back(rtw_xz_rect(s, mat, mn[0], mx[0], mn[2], mx[2], mx[1]));
    h.children.push_back(rtw_xz_rect(s, mat, mn[0], mx[0], mn[2], mx[2], mn[1]));
    h.children.push_back(rtw_yz_rect(s, mat, mn[1], mx[1], mn[2], mx[2], mx[0]));
    h.children.push_back(rtw_yz_rect(s, mat, mn[1], mx[1], mn[2], mx[2], mn[0]));
    return push_node(s, h);
}
int rtw_translate(rtw_scene* s, int child, const double offset[3]) {
    if (!node_ok(s, child) || !offset) return fail(RTW_ERR_INVALID_ARG, "bad child id");
    if (!fin3(offset)) return fail(RTW_ERR_INVALID_ARG, "non-finite offset");
    rtw::HNode h; h.kind = rtw::H_TRANSLATE; h.child = child; h.offset = v3(offset); return push_node(s, h);
}
int rtw_rotate_y(rtw_scene* s, double angle_deg, int child) {                           // new_rotate_y src/hittable.rs:147-152
    if (!node_ok(s, child)) return fail(RTW_ERR_INVALID_ARG, "bad child id");
    if (!fin(angle_deg)) return fail(RTW_ERR_INVALID_ARG, "non-finite angle");
    rtw::HNode h; h.kind = rtw::H_ROTATE_Y; h.child = child; h.angle_deg = angle_deg;
    double radians = angle_deg * 3.1415926535897932385 / 180.0;
    h.sin_theta = std::sin(radians); h.cos_theta = std::cos(radians);
    return push_node(s, h);
}
// A reference RotateY stores sin_theta / cos_theta, not the angle (src/hittable.rs:39): a host that walks an existing
// Hittable tree hands those two numbers over unchanged instead of round-tripping them through atan2.
int rtw_rotate_y_sincos(rtw_scene* s, double sin_theta, double cos_theta, int child) {
    if (!node_ok(s, child)) return fail(RTW_ERR_INVALID_ARG, "bad child id");
    if (!fin_all(sin_theta, cos_theta)) return fail(RTW_ERR_INVALID_ARG, "non-finite sin / cos");
    rtw::HNode h; h.kind = rtw::H_ROTATE_Y; h.child = child; h.sin_theta = sin_theta; h.cos_theta = cos_theta;
    h.angle_deg = std::atan2(sin_theta, cos_theta) * 180.0 / 3.1415926535897932385;
    return push_node(s, h);
}
int rtw_constant_medium(rtw_scene* s, int child, double density, int phase_mat) {      // src/hittable.rs:201-207
    if (!node_ok(s, child)) return fail(RTW_ERR_INVALID_ARG, "bad child id");
    if (!mat_ok(s, phase_mat)) return fail(RTW_ERR_INVALID_ARG, "bad material handle");
    if (!fin(density) || density == 0.0) return fail(RTW_ERR_INVALID_ARG, "medium density must be finite and non-zero (-1/density, src/hittable.rs:205)");
    rtw::HNode h; h.kind = rtw::H_MEDIUM; h.child = child; h.mat = phase_mat; h.density = density; return push_node(s, h);
}
int rtw_bvh_node(rtw_scene* s, const int32_t* children, int32_t n, double t0, double t1) {
    if (!s || !children || n <= 0) return fail(RTW_ERR_INVALID_ARG, "empty BvhNode");
    if (!fin_all(t0, t1)) return fail(RTW_ERR_INVALID_ARG, "non-finite time range");
    rtw::HNode h; h.kind = rtw::H_BVH_NODE; h.time0 = t0; h.time1 = t1;
    for (int i = 0; i < n; ++i) { if (!node_ok(s, children[i])) return fail(RTW_ERR_INVALID_ARG, "bad child id"); h.children.push_back(children[i]); }
    return push_node(s, h);
}
int rtw_world_push(rtw_scene* s, int id) {
    if (!node_ok(s, id)) return fail(RTW_ERR_INVALID_ARG, "bad hittable id");
    s->g.world.push_back(id); s->committed = false; return RTW_OK;
}

int rtw_camera_new(const double look_from[3], const double look_at[3], const double vup[3], double vfov, double aspect,
                   double aperture, double focus_dist, double time0, double time1, rtw_camera* out) {   // src/camera.rs:18-56
    if (!look_from || !look_at || !vup || !out) return fail(RTW_ERR_INVALID_ARG, "null argument");
    if (!fin3(look_from) || !fin3(look_at) || !fin3(vup) || !fin_all(vfov, aspect, aperture, focus_dist, time0, time1))
        return fail(RTW_ERR_INVALID_ARG, "non-finite camera parameters");
    auto sub = [](const double* a, const double* b, double* r) { for (int i = 0; i < 3; ++i) r[i] = a[i] - b[i]; };
    auto norm = [](double* v) { double inv = 1.0 / std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); for (int i = 0; i < 3; ++i) v[i] = inv * v[i]; };
    auto cross = [](const double* u, const double* v, double* r) { r[0] = u[1] * v[2] - u[2] * v[1]; r[1] = u[2] * v[0] - u[0] * v[2]; r[2] = u[0] * v[1] - u[1] * v[0]; };
    double theta = vfov * 3.1415926535897932385 / 180.0;
    double h = std::tan(theta / 2.0);
    double viewport_height = 2.0 * h, viewport_width = aspect * viewport_height;
    double w[3], u[3], v[3];
    sub(look_from, look_at, w); norm(w);
    cross(vup, w, u); norm(u);
    cross(w, u, v);
    for (int i = 0; i < 3; ++i) {
        out->origin[i] = look_from[i];
        out->horizontal[i] = focus_dist * viewport_width * u[i];
        out->vertical[i] = focus_dist * viewport_height * v[i];
        out->lower_left_corner[i] = out->origin[i] - out->horizontal[i] * 0.5 - out->vertical[i] * 0.5 - focus_dist * w[i];
        out->u[i] = u[i]; out->v[i] = v[i]; out->w[i] = w[i];
    }
    out->lens_radius = aperture * 0.5; out->time0 = time0; out->time1 = time1;
    for (int i = 0; i < 3; ++i)      // look_from == look_at, vup parallel to the view direction, vfov = 180: the reference divides by 0
        if (!fin_all(out->horizontal[i], out->vertical[i], out->lower_left_corner[i], out->u[i], out->v[i], out->w[i]))
            return fail(RTW_ERR_INVALID_ARG, "degenerate camera (look_from == look_at, vup parallel to the view axis, or vfov out of range)");
    return RTW_OK;
}

static int commit_impl(rtw_scene* s, int32_t n_gpus, int32_t first_device);
int rtw_scene_commit(rtw_scene* s, int32_t n_gpus, int32_t first_device) {
    if (!s) return fail(RTW_ERR_INVALID_ARG, "null scene");
    s->committed = false;          // a commit that fails part-way leaves the scene uncommitted and without replicas:
    const int rc = commit_impl(s, n_gpus, first_device);      // no later render can launch on a half-built replica
    if (rc < 0) { std::string keep = g_err; free_replicas(s); cudaGetLastError(); g_err = keep; }
    return rc;
}
static int commit_impl(rtw_scene* s, int32_t n_gpus, int32_t first_device) {
    TRY(need_device());
    int ndev = rtw_device_count();
    if (n_gpus <= 0) n_gpus = ndev - first_device;
    if (first_device < 0 || n_gpus < 1 || first_device + n_gpus > ndev || n_gpus > 8) return fail(RTW_ERR_INVALID_ARG, "device range not available");
    double t0 = now_ms();
    std::string err;
    // Where the BVH is built.  Big scenes (RTW_BIG_MIN primitives, default 256 Ki; RTW_DEVICE_BUILD_MIN overrides) are built ON THE
    // DEVICE from what the constructors were given (bvh_build.cu): 16 M spheres commit in 0.24 s instead of 5.1 s.  As 8-wide
    // nodes under the wavefront pipeline the LBVH renders as fast as the host's binned-SAH tree (1 M: 378 vs 382, 4 M: 306 vs
    // 304 Mpaths/s, profiles/r2_v_wavefront_variants.log; as binary nodes under the megakernel it was ~6 % slower).  Small
    // scenes keep the host SAH builder (C1: 1 ms, better trees).  RTW_DEVICE_BUILD=0 / 1 forces either.
    rtw::FlattenOptions fo;
    {
        long long n_guess = (long long)s->g.bulk.size() + (long long)s->g.nodes.size();
        long long dev_min = big_scene_min();
        if (const char* e = getenv("RTW_DEVICE_BUILD_MIN")) dev_min = atoll(e);
        fo.emit_only = n_guess >= dev_min;
        if (const char* e = getenv("RTW_DEVICE_BUILD")) fo.emit_only = atoi(e) != 0;
    }
    int rc = rtw::flatten(s->g, s->g.world, s->flat, err, fo);
    if (rc) return fail(rc, err);
    const double t_flat = now_ms();
    if (s->flat.emit_only && s->flat.n_bvh_prims == 0) {         // nothing to build: the plain path handles the empty world
        fo.emit_only = false;
        rc = rtw::flatten(s->g, s->g.world, s->flat, err, fo);
        if (rc) return fail(rc, err);
    }
    Packed pk; pack(s->flat, pk);
    // replicas (stream, events, grid size, blob allocation) are reused across commits on the same devices:
    // a re-commit is then one flatten + one H2D copy per GPU
    bool reuse = (int)s->reps.size() == n_gpus;
    for (int i = 0; reuse && i < n_gpus; ++i) reuse = s->reps[i].device == first_device + i;
    if (!reuse) { free_replicas(s); s->reps.resize(n_gpus); }
    s->h2d_commit = 0;
    for (int i = 0; i < n_gpus; ++i) {
        Replica& r = s->reps[i];
        CUDA_TRY(cudaSetDevice(first_device + i));
        if (!reuse) {
            r.device = first_device + i;
            CUDA_TRY(cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreate(&r.ev0)); CUDA_TRY(cudaEventCreate(&r.ev1));
            CUDA_TRY(cudaMalloc(&r.stats, 32));
            int sms = 0, per_sm = 0;
            CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, r.device));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_kernel<FEAT_ALL, 0, 0>, RTW_BLOCK, 0));
            if (per_sm < 1) per_sm = 1;
            r.grid = sms * per_sm; r.sms = sms;
            for (int k = 0; k < 8; ++k) r.pool_grid[k] = 0;
            if (i > 0) {   // peers write the framebuffer / counter that live on the first device
                int can = 0; CUDA_TRY(cudaDeviceCanAccessPeer(&can, r.device, first_device));
                if (!can) return fail(RTW_ERR_CUDA, "peer access to the first device is not available");
                cudaError_t e = cudaDeviceEnablePeerAccess(first_device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(RTW_ERR_CUDA, cudaGetErrorString(e));
                cudaGetLastError();
            }
        }
        if (r.blob_bytes < pk.total) {
            if (r.blob) { CUDA_TRY(cudaFree(r.blob)); r.blob = nullptr; }
            CUDA_TRY(cudaMalloc(&r.blob, pk.total)); r.blob_bytes = pk.total;
        }
        rtwb::free_output(r.built);
        if (!s->flat.emit_only) {
            CUDA_TRY(upload(s->flat, pk, r.blob, r.stream));
            s->h2d_commit += pk.total;
            r.ds = bind(s->flat, pk, r.blob);
            continue;
        }
        // ---- device-side build: the blob carries the small tables only; nodes and primitive records are built in place
        const size_t small = pk.o_xf;                         // [o_xf, total): xforms, media, materials, textures, perlin, image
        {
            const rtw::FlatScene& f = s->flat;
            auto cp = [&](size_t o, const void* src, size_t n) { return n ? cudaMemcpyAsync(r.blob + o, src, n, cudaMemcpyHostToDevice, r.stream) : cudaSuccess; };
            CUDA_TRY(cp(pk.o_xf, f.xforms.data(), f.xforms.size() * sizeof(DXform)));
            CUDA_TRY(cp(pk.o_media, f.media.data(), f.media.size() * sizeof(DMedium)));
            CUDA_TRY(cp(pk.o_mats, f.mats.data(), f.mats.size() * sizeof(DMat)));
            CUDA_TRY(cp(pk.o_texs, f.texs.data(), f.texs.size() * sizeof(DTex)));
            CUDA_TRY(cp(pk.o_perlin, f.perlin.data(), f.perlin.size()));
            CUDA_TRY(cp(pk.o_image, f.image.data(), f.image.size()));
            s->h2d_commit += pk.total - small;
        }
        r.ds = bind(s->flat, pk, r.blob);
        if (i == 0) {
            rtwb::BuildInput in;
            in.host_prims = s->flat.prims.data(); in.host_boxes = s->flat.prim_boxes.data(); in.n_host = (int)s->flat.prims.size();
            in.bulk = reinterpret_cast<const rtwb::BulkSphereD*>(s->g.bulk.data()); in.n_bulk = s->flat.n_bvh_prims - in.n_host;
            in.boundary_prims = s->flat.boundary.data(); in.n_boundary = (int)s->flat.boundary.size();
            in.width = s->flat.wide ? 8 : 2;
            const double tb = now_ms();
            const int brc = rtwb::build_on_device(r.stream, in, r.built, err);
            if (brc) return fail(brc, "device BVH build: " + err);
            s->ms_build_dev = now_ms() - tb;
            s->h2d_commit += r.built.h2d_bytes;
            s->n_prims_dev = r.built.n_prims; s->n_nodes_dev = s->flat.wide ? r.built.n_wnodes : r.built.n_nodes;
            s->flat.wide_depth = r.built.depth;
        } else {                                              // other replicas: copy the finished arrays over NVLink
            const rtwb::BuildOutput& b0 = s->reps[0].built;
            r.built.n_wnodes = b0.n_wnodes; r.built.n_nodes = b0.n_nodes; r.built.n_prims = b0.n_prims; r.built.depth = b0.depth;
            if (b0.n_wnodes) { CUDA_TRY(cudaMalloc(&r.built.wnodes, (size_t)b0.n_wnodes * sizeof(DWNode))); CUDA_TRY(cudaMemcpyPeerAsync(r.built.wnodes, r.device, b0.wnodes, s->reps[0].device, (size_t)b0.n_wnodes * sizeof(DWNode), r.stream)); }
            if (b0.n_nodes) { CUDA_TRY(cudaMalloc(&r.built.nodes, (size_t)b0.n_nodes * sizeof(DNode))); CUDA_TRY(cudaMemcpyPeerAsync(r.built.nodes, r.device, b0.nodes, s->reps[0].device, (size_t)b0.n_nodes * sizeof(DNode), r.stream)); }
            CUDA_TRY(cudaMalloc(&r.built.prims, (size_t)b0.n_prims * sizeof(DPrim)));
            CUDA_TRY(cudaMemcpyPeerAsync(r.built.prims, r.device, b0.prims, s->reps[0].device, (size_t)b0.n_prims * sizeof(DPrim), r.stream));
        }
        r.ds.nodes = r.built.nodes; r.ds.wnodes = r.built.wnodes; r.ds.prims = r.built.prims;
        r.ds.n_nodes = s->flat.wide ? r.built.n_wnodes : r.built.n_nodes; r.ds.n_prims = r.built.n_prims; r.ds.n_bvh_prims = s->flat.n_bvh_prims;
    }
    for (Replica& r : s->reps) { CUDA_TRY(cudaSetDevice(r.device)); CUDA_TRY(cudaStreamSynchronize(r.stream)); }
    s->committed = true;
    s->ms_commit = now_ms() - t0;
    if (getenv("RTW_TIMING")) fprintf(stderr, "[commit] flatten %.2f ms, device build (rank-0 replica, with its scratch frees) %.2f ms, all %.2f ms\n", t_flat - t0, s->flat.emit_only ? s->ms_build_dev : 0.0, s->ms_commit);
    return RTW_OK;
}

int rtw_render(rtw_scene* s, const rtw_camera* cam, const rtw_render_params* p, float* out, rtw_stats* st) {
    if (!s || !cam || !p || !out) return fail(RTW_ERR_INVALID_ARG, "null argument");
    if (!s->committed) return fail(RTW_ERR_NOT_COMMITTED, "rtw_scene_commit has not been called since the last edit");
    double t0 = now_ms();
    int n_rep = p->n_gpus > 0 ? p->n_gpus : (int)s->reps.size();
    if (n_rep > (int)s->reps.size()) return fail(RTW_ERR_INVALID_ARG, "n_gpus exceeds the committed replicas");
    int total_warps = 0;
    for (int i = 0; i < n_rep; ++i) total_warps += s->reps[i].grid * RTW_WARPS;
    rtw_render_params pu = *p; pu.n_gpus = n_rep;             // (the unit sizing wants to know how many GPUs share the frame)
    DParams dp; TRY(make_params(pu, total_warps, dp));
    if (st) { std::memset(st, 0, sizeof(*st)); fill_scene_stats(s, st); }
    Replica& r0 = s->reps[0];
    const bool dev_out = (p->flags & RTW_FLAG_DEVICE_OUT) != 0;
    const size_t fb_bytes = (size_t)p->width * p->height * 3 * sizeof(float);
    TRY(ensure_fb(s->local, r0.device, dev_out ? 1 : p->width, dev_out ? 1 : p->height));
    float* fb = dev_out ? out : s->local.fb();
    CUDA_TRY(cudaSetDevice(r0.device));
    dp.accumulate = (RTW_CARRY || n_rep > 1 || dp.chunks + dp.chunks_b > 1 || kernel_mode(s, p->flags) == 5) ? 1 : 0;   // (wavefront, orphans of a carried unit: finished paths are ADDED)
    CUDA_TRY(cudaMemsetAsync(s->local.counter(), 0, 256, r0.stream));
    if (dp.accumulate) CUDA_TRY(cudaMemsetAsync(fb, 0, fb_bytes, r0.stream));
    if (n_rep > 1) CUDA_TRY(cudaStreamSynchronize(r0.stream));   // peers must see the zeroed buffers
    TRY(launch_all(s, n_rep, cam, dp, s->local.counter(), fb, st, kernel_mode(s, p->flags)));
    if (!dev_out) {
        CUDA_TRY(cudaSetDevice(r0.device));
        CUDA_TRY(cudaMemcpy(out, fb, fb_bytes, cudaMemcpyDeviceToHost));
        if (st) st->d2h_bytes = fb_bytes;
    }
    if (st) { st->paths = (uint64_t)p->width * p->height * p->spp; st->ms_total = now_ms() - t0; st->h2d_bytes = sizeof(DCamera) + sizeof(DParams) + sizeof(DScene); }
    return RTW_OK;
}

int rtw_render_progressive(rtw_scene* s, const rtw_camera* cam, const rtw_render_params* p, int32_t first_sample, int32_t samples_per_pass,
                           float* inout, rtw_progress_fn progress, void* user, rtw_stats* st) {
    if (!s || !cam || !p || !inout) return fail(RTW_ERR_INVALID_ARG, "null argument");
    if (!s->committed) return fail(RTW_ERR_NOT_COMMITTED, "rtw_scene_commit has not been called since the last edit");
    if (p->flags & RTW_FLAG_DEVICE_OUT) return fail(RTW_ERR_INVALID_ARG, "rtw_render_progressive works on a host buffer");
    if (first_sample < 0 || first_sample > p->spp || samples_per_pass < 0) return fail(RTW_ERR_INVALID_ARG, "bad sample range");
    // the ABSOLUTE sample index is a Philox counter coordinate and travels in 21 bits of a queued ray (17 in the pool kernel)
    if (p->spp > (1 << 20)) return fail(RTW_ERR_INVALID_ARG, "spp <= 1048576");
    double t0 = now_ms();
    int n_rep = p->n_gpus > 0 ? p->n_gpus : (int)s->reps.size();
    if (n_rep > (int)s->reps.size()) return fail(RTW_ERR_INVALID_ARG, "n_gpus exceeds the committed replicas");
    int total_warps = 0;
    for (int i = 0; i < n_rep; ++i) total_warps += s->reps[i].grid * RTW_WARPS;
    if (st) { std::memset(st, 0, sizeof(*st)); fill_scene_stats(s, st); }
    const int per_pass = samples_per_pass > 0 ? samples_per_pass : (p->spp + 9) / 10 > 0 ? (p->spp + 9) / 10 : 1;
    Replica& r0 = s->reps[0];
    const size_t fb_bytes = (size_t)p->width * p->height * 3 * sizeof(float);
    TRY(ensure_fb(s->local, r0.device, p->width, p->height));
    float* fb = s->local.fb();
    CUDA_TRY(cudaSetDevice(r0.device));
    // the device framebuffer carries the running sums across passes; a resume starts from the caller's checkpoint
    if (first_sample > 0) { CUDA_TRY(cudaMemcpyAsync(fb, inout, fb_bytes, cudaMemcpyHostToDevice, r0.stream)); if (st) st->h2d_bytes += fb_bytes; }
    else CUDA_TRY(cudaMemsetAsync(fb, 0, fb_bytes, r0.stream));
    int done = first_sample;
    double ms_render = 0; uint64_t rays = 0; uint64_t units[8] = {0, 0, 0, 0, 0, 0, 0, 0}; int launches = 0;
    if (done >= p->spp) { CUDA_TRY(cudaStreamSynchronize(r0.stream)); if (first_sample == 0) std::memset(inout, 0, fb_bytes); }
    while (done < p->spp) {
        rtw_render_params pass = *p;
        pass.spp = p->spp - done < per_pass ? p->spp - done : per_pass;
        DParams dp; TRY(make_params(pass, total_warps, dp));
        dp.first_sample = done;
        dp.accumulate = 1;
        CUDA_TRY(cudaSetDevice(r0.device));
        CUDA_TRY(cudaMemsetAsync(s->local.counter(), 0, 256, r0.stream));
        CUDA_TRY(cudaStreamSynchronize(r0.stream));               // peers (and the next pass) must see counter + sums
        rtw_stats ps; std::memset(&ps, 0, sizeof(ps));
        TRY(launch_all(s, n_rep, cam, dp, s->local.counter(), fb, &ps, kernel_mode(s, p->flags)));
        ms_render += ps.ms_render; rays += ps.rays; launches += ps.kernel_launches;
        for (int i = 0; i < 8; ++i) units[i] += ps.units_per_device[i];
        done += pass.spp;
        CUDA_TRY(cudaSetDevice(r0.device));
        CUDA_TRY(cudaMemcpy(inout, fb, fb_bytes, cudaMemcpyDeviceToHost));
        if (st) st->d2h_bytes += fb_bytes;
        if (progress && progress(done, p->spp, inout, user) != 0) break;
    }
    if (st) {
        st->ms_render = ms_render; st->rays = rays; st->kernel_launches = launches; st->n_devices = n_rep;
        for (int i = 0; i < 8; ++i) st->units_per_device[i] = units[i];
        st->paths = (uint64_t)p->width * p->height * (uint64_t)(done - first_sample);
        st->ms_total = now_ms() - t0;
    }
    return RTW_OK;
}

void* rtw_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); fail(RTW_ERR_OOM, "cudaMallocHost failed"); return nullptr; }
    return p;
}
void rtw_host_free(void* p) { if (p) cudaFreeHost(p); }

int rtw_write_color(const float* rgb_sum, int32_t n_pixels, int32_t spp, uint8_t* out_rgb8) {
    if (!rgb_sum || !out_rgb8 || n_pixels <= 0 || spp <= 0) return fail(RTW_ERR_INVALID_ARG, "bad argument");
    TRY(need_device());
    Scratch sc; float* d_in; uint8_t* d_out; size_t n = (size_t)n_pixels * 3;
    TRY(sc.up(rgb_sum, n, d_in)); TRY(sc.up((const uint8_t*)nullptr, n, d_out));
    write_color_kernel<<<(unsigned)((n + 255) / 256), 256>>>(d_in, (int)n, 1.0 / (double)spp, d_out);
    CUDA_TRY(cudaGetLastError());
    TRY(sc.down(out_rgb8, d_out, n));
    return RTW_OK;
}

// ---- cross-process sharing (one rank per GPU): CUDA IPC ----------------------------------------
struct IpcBlob { cudaIpcMemHandle_t h; uint64_t bytes; int32_t width, height; uint32_t magic; };
static_assert(sizeof(IpcBlob) <= RTW_IPC_HANDLE_BYTES, "IPC blob too large");

int rtw_shared_create(rtw_scene* s, int32_t w, int32_t h, uint8_t handle[RTW_IPC_HANDLE_BYTES]) {
    if (!s || !handle || w < 2 || h < 2) return fail(RTW_ERR_INVALID_ARG, "bad argument");
    if (!s->committed) return fail(RTW_ERR_NOT_COMMITTED, "commit first");
    if (s->shared.base) rtw_shared_close(s);
    CUDA_TRY(cudaSetDevice(s->reps[0].device));
    s->shared.width = w; s->shared.height = h;
    size_t need = 256 + 2 * s->shared.fb_bytes();        // two halves: epochs alternate (rtw_render_shared_epoch)
    CUDA_TRY(cudaMalloc(&s->shared.base, need));
    s->shared.bytes = need; s->shared.owner = true;
    CUDA_TRY(cudaMemset(s->shared.base, 0, need));
    CUDA_TRY(cudaStreamCreateWithFlags(&s->shared.side, cudaStreamNonBlocking));
    IpcBlob b; std::memset(&b, 0, sizeof(b));
    CUDA_TRY(cudaIpcGetMemHandle(&b.h, s->shared.base));
    b.bytes = need; b.width = w; b.height = h; b.magic = 0x52545721u;
    std::memset(handle, 0, RTW_IPC_HANDLE_BYTES); std::memcpy(handle, &b, sizeof(b));
    return RTW_OK;
}
int rtw_shared_open(rtw_scene* s, int32_t w, int32_t h, const uint8_t handle[RTW_IPC_HANDLE_BYTES]) {
    if (!s || !handle) return fail(RTW_ERR_INVALID_ARG, "bad argument");
    if (!s->committed) return fail(RTW_ERR_NOT_COMMITTED, "commit first");
    IpcBlob b; std::memcpy(&b, handle, sizeof(b));
    if (b.magic != 0x52545721u || b.width != w || b.height != h) return fail(RTW_ERR_INVALID_ARG, "IPC handle does not match");
    if (s->shared.base) rtw_shared_close(s);
    CUDA_TRY(cudaSetDevice(s->reps[0].device));
    void* p = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&p, b.h, cudaIpcMemLazyEnablePeerAccess));
    s->shared.base = (uint8_t*)p; s->shared.bytes = b.bytes; s->shared.width = w; s->shared.height = h;
    s->shared.owner = false; s->shared.ipc_mapped = true;
    return RTW_OK;
}
int rtw_shared_reset(rtw_scene* s) {
    if (!s || !s->shared.base || !s->shared.owner) return fail(RTW_ERR_INVALID_ARG, "not the owner of a shared framebuffer");
    CUDA_TRY(cudaSetDevice(s->reps[0].device));
    CUDA_TRY(cudaMemset(s->shared.base, 0, s->shared.bytes));
    CUDA_TRY(cudaDeviceSynchronize());
    return RTW_OK;
}
// One whole-image render per EPOCH without a reset call or a barrier before the launch: epoch e renders into half e & 1 of
// the shared allocation (its own unit counter and framebuffer) while the owner zeroes the other half on a side stream for
// epoch e + 1.  Contract: every rank has returned from epoch e - 1 before any rank calls epoch e (one barrier per step, which
// a timed loop needs anyway), and the owner's device is synchronised before epoch e + 1 starts (the barrier's cuda
// synchronize).  Replaces reset + barrier + render + barrier: one barrier per step instead of three.
int rtw_render_shared_epoch(rtw_scene* s, const rtw_camera* cam, const rtw_render_params* p, int32_t epoch, rtw_stats* st) {
    if (!s || !cam || !p || !s->shared.base || epoch < 0) return fail(RTW_ERR_INVALID_ARG, "no shared framebuffer");
    if (!s->committed) return fail(RTW_ERR_NOT_COMMITTED, "rtw_scene_commit has not been called since the last edit");
    if (p->width != s->shared.width || p->height != s->shared.height) return fail(RTW_ERR_INVALID_ARG, "size mismatch");
    double t0 = now_ms();
    const int half = epoch & 1;
    int world = p->n_gpus > 0 ? p->n_gpus : 1;
    DParams dp; TRY(make_params(*p, world * s->reps[0].grid * RTW_WARPS, dp));
    dp.accumulate = 1;
    if (st) { std::memset(st, 0, sizeof(*st)); fill_scene_stats(s, st); }
    if (s->shared.owner) {                       // the idle half, for the next epoch (nobody touches it during this one)
        CUDA_TRY(cudaSetDevice(s->reps[0].device));
        CUDA_TRY(cudaMemsetAsync(s->shared.counter(half ^ 1), 0, 128, s->shared.side));
        CUDA_TRY(cudaMemsetAsync(s->shared.fb(half ^ 1), 0, s->shared.fb_bytes(), s->shared.side));
    }
    TRY(launch_all(s, 1, cam, dp, s->shared.counter(half), s->shared.fb(half), st, kernel_mode(s, p->flags)));
    if (st) { st->paths = (uint64_t)p->width * p->height * p->spp; st->ms_total = now_ms() - t0; }
    return RTW_OK;
}
int rtw_shared_read_epoch(rtw_scene* s, int32_t epoch, float* out) {
    if (!s || !out || !s->shared.base || epoch < 0) return fail(RTW_ERR_INVALID_ARG, "no shared framebuffer");
    CUDA_TRY(cudaSetDevice(s->reps[0].device));
    CUDA_TRY(cudaMemcpy(out, s->shared.fb(epoch & 1), (size_t)s->shared.width * s->shared.height * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    return RTW_OK;
}
int rtw_render_shared(rtw_scene* s, const rtw_camera* cam, const rtw_render_params* p, rtw_stats* st) {
    if (!s || !cam || !p || !s->shared.base) return fail(RTW_ERR_INVALID_ARG, "no shared framebuffer");
    if (!s->committed) return fail(RTW_ERR_NOT_COMMITTED, "rtw_scene_commit has not been called since the last edit");
    if (p->width != s->shared.width || p->height != s->shared.height) return fail(RTW_ERR_INVALID_ARG, "size mismatch");
    double t0 = now_ms();
    int world = p->n_gpus > 0 ? p->n_gpus : 1;    // ranks taking part (sizes the work units)
    DParams dp; TRY(make_params(*p, world * s->reps[0].grid * RTW_WARPS, dp));
    dp.accumulate = 1;
    if (st) { std::memset(st, 0, sizeof(*st)); fill_scene_stats(s, st); }
    TRY(launch_all(s, 1, cam, dp, s->shared.counter(), s->shared.fb(), st, kernel_mode(s, p->flags)));
    if (st) { st->paths = (uint64_t)p->width * p->height * p->spp; st->ms_total = now_ms() - t0; }
    return RTW_OK;
}
int rtw_shared_read(rtw_scene* s, float* out) {
    if (!s || !out || !s->shared.base) return fail(RTW_ERR_INVALID_ARG, "no shared framebuffer");
    CUDA_TRY(cudaSetDevice(s->reps[0].device));
    CUDA_TRY(cudaMemcpy(out, s->shared.fb(), (size_t)s->shared.width * s->shared.height * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    return RTW_OK;
}
int rtw_shared_close(rtw_scene* s) {
    if (!s || !s->shared.base) return RTW_OK;
    if (!s->reps.empty()) cudaSetDevice(s->reps[0].device);
    if (s->shared.side) { cudaStreamSynchronize(s->shared.side); cudaStreamDestroy(s->shared.side); }
    if (s->shared.ipc_mapped) cudaIpcCloseMemHandle(s->shared.base); else cudaFree(s->shared.base);
    s->shared = SharedFb();
    return RTW_OK;
}

// ---- parity hooks ---------------------------------------------------------------------------------
int rtw_test_philox(int32_t n, const uint32_t* counter, const uint32_t* key, uint32_t* out) {
    if (n <= 0 || !counter || !key || !out) return fail(RTW_ERR_INVALID_ARG, "bad argument");
    TRY(need_device());
    Scratch sc; uint32_t *dc, *dk, *dout;
    TRY(sc.up(counter, 4 * (size_t)n, dc)); TRY(sc.up(key, 2 * (size_t)n, dk)); TRY(sc.up((uint32_t*)nullptr, 4 * (size_t)n, dout));
    philox_kernel<<<(n + 127) / 128, 128>>>(n, dc, dk, dout);
    CUDA_TRY(cudaGetLastError());
    return sc.down(out, dout, 4 * (size_t)n);
}

int rtw_test_get_ray(const rtw_camera* cam, int32_t n, const double* s, const double* t, const double* xi, int32_t stride,
                     double* oo, double* od, double* ot, int32_t* nd) {
    if (!cam || n <= 0 || !s || !t || !xi || stride <= 0) return fail(RTW_ERR_INVALID_ARG, "bad argument");
    TRY(need_device());
    Scratch sc; double *ds, *dt, *dxi, *doo, *dod, *dot_; int* dnd;
    TRY(sc.up(s, n, ds)); TRY(sc.up(t, n, dt)); TRY(sc.up(xi, (size_t)n * stride, dxi));
    TRY(sc.up((double*)nullptr, 3 * (size_t)n, doo)); TRY(sc.up((double*)nullptr, 3 * (size_t)n, dod));
    TRY(sc.up((double*)nullptr, n, dot_)); TRY(sc.up((int*)nullptr, n, dnd));
    get_ray_kernel<<<(n + 127) / 128, 128>>>(to_dcamera(*cam), n, ds, dt, dxi, stride, doo, dod, dot_, dnd);
    CUDA_TRY(cudaGetLastError());
    TRY(sc.down(oo, doo, 3 * (size_t)n)); TRY(sc.down(od, dod, 3 * (size_t)n)); TRY(sc.down(ot, dot_, n)); TRY(sc.down(nd, dnd, n));
    return RTW_OK;
}

// scene view for the hooks: the committed world, or one hittable flattened on the fly
struct View {
    uint8_t* blob = nullptr; DScene ds{}; bool owned = false;
    ~View() { if (owned && blob) cudaFree(blob); }
};
static int make_view(rtw_scene* s, int target, View& v) {
    if (target < 0) {
        if (!s->committed) return fail(RTW_ERR_NOT_COMMITTED, "commit first");
        CUDA_TRY(cudaSetDevice(s->reps[0].device));
        v.ds = s->reps[0].ds; return 0;
    }
    if (!node_ok(s, target)) return fail(RTW_ERR_INVALID_ARG, "bad target");
    rtw::FlatScene f; std::string err;
    int rc = rtw::flatten(s->g, std::vector<int>{target}, f, err);
    if (rc) return fail(rc, err);
    Packed pk; pack(f, pk);
    CUDA_TRY(cudaMalloc(&v.blob, pk.total)); v.owned = true;
    CUDA_TRY(upload(f, pk, v.blob, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    v.ds = bind(f, pk, v.blob);
    return 0;
}

int rtw_test_hit(rtw_scene* s, int32_t target, int32_t n, const double* o, const double* d, const double* tm, double t_min,
                 double t_max, const double* xi, int32_t stride, int32_t* hit, double* ot, double* op, double* on,
                 int32_t* front, double* ou, double* ov, int32_t* mat, int32_t* nd) {
    if (!s || n <= 0 || !o || !d || !tm || !xi || stride <= 0) return fail(RTW_ERR_INVALID_ARG, "bad argument");
    TRY(need_device());
    View v; TRY(make_view(s, target, v));
    Scratch sc; double *d_o, *d_d, *d_tm, *d_xi, *d_t, *d_p, *d_n, *d_u, *d_v; int *d_hit, *d_front, *d_mat, *d_nd;
    TRY(sc.up(o, 3 * (size_t)n, d_o)); TRY(sc.up(d, 3 * (size_t)n, d_d)); TRY(sc.up(tm, n, d_tm)); TRY(sc.up(xi, (size_t)n * stride, d_xi));
    TRY(sc.up((double*)nullptr, n, d_t)); TRY(sc.up((double*)nullptr, 3 * (size_t)n, d_p)); TRY(sc.up((double*)nullptr, 3 * (size_t)n, d_n));
    TRY(sc.up((double*)nullptr, n, d_u)); TRY(sc.up((double*)nullptr, n, d_v));
    TRY(sc.up((int*)nullptr, n, d_hit)); TRY(sc.up((int*)nullptr, n, d_front)); TRY(sc.up((int*)nullptr, n, d_mat)); TRY(sc.up((int*)nullptr, n, d_nd));
    hit_kernel<<<(n + 127) / 128, 128>>>(v.ds, n, d_o, d_d, d_tm, (float)t_min, (float)t_max, d_xi, stride, d_hit, d_t, d_p, d_n, d_front, d_u, d_v, d_mat, d_nd);
    CUDA_TRY(cudaGetLastError());
    TRY(sc.down(hit, d_hit, n)); TRY(sc.down(ot, d_t, n)); TRY(sc.down(op, d_p, 3 * (size_t)n)); TRY(sc.down(on, d_n, 3 * (size_t)n));
    TRY(sc.down(front, d_front, n)); TRY(sc.down(ou, d_u, n)); TRY(sc.down(ov, d_v, n)); TRY(sc.down(mat, d_mat, n)); TRY(sc.down(nd, d_nd, n));
    return RTW_OK;
}

int rtw_test_aabb(int32_t n, const double* bmin, const double* bmax, const double* o, const double* d, double t_min, double t_max, int32_t* hit) {
    if (n <= 0 || !bmin || !bmax || !o || !d || !hit) return fail(RTW_ERR_INVALID_ARG, "bad argument");
    TRY(need_device());
    Scratch sc; double *d_mn, *d_mx, *d_o, *d_d; int* d_hit;
    TRY(sc.up(bmin, 3 * (size_t)n, d_mn)); TRY(sc.up(bmax, 3 * (size_t)n, d_mx)); TRY(sc.up(o, 3 * (size_t)n, d_o)); TRY(sc.up(d, 3 * (size_t)n, d_d));
    TRY(sc.up((int*)nullptr, n, d_hit));
    aabb_kernel<<<(n + 127) / 128, 128>>>(n, d_mn, d_mx, d_o, d_d, (float)t_min, (float)t_max, d_hit);
    CUDA_TRY(cudaGetLastError());
    return sc.down(hit, d_hit, n);
}

int rtw_test_scatter(rtw_scene* s, int32_t mat, int32_t n, const double* ro, const double* rd, const double* rt, const double* p,
                     const double* nrm, const int32_t* front, const double* u, const double* v, const double* xi, int32_t stride,
                     int32_t* osc, double* oo, double* od, double* ot, double* oatt, double* oem, int32_t* nd) {
    if (!s || n <= 0 || !ro || !rd || !rt || !p || !nrm || !front || !u || !v || !xi || stride <= 0) return fail(RTW_ERR_INVALID_ARG, "bad argument");
    if (!mat_ok(s, mat)) return fail(RTW_ERR_INVALID_ARG, "bad material handle");
    TRY(need_device());
    View vw; TRY(make_view(s, -1, vw));
    Scratch sc; double *d_ro, *d_rd, *d_rt, *d_p, *d_n, *d_u, *d_v, *d_xi, *d_oo, *d_od, *d_ot, *d_att, *d_em; int *d_front, *d_sc, *d_nd;
    TRY(sc.up(ro, 3 * (size_t)n, d_ro)); TRY(sc.up(rd, 3 * (size_t)n, d_rd)); TRY(sc.up(rt, n, d_rt)); TRY(sc.up(p, 3 * (size_t)n, d_p));
    TRY(sc.up(nrm, 3 * (size_t)n, d_n)); TRY(sc.up(front, n, d_front)); TRY(sc.up(u, n, d_u)); TRY(sc.up(v, n, d_v)); TRY(sc.up(xi, (size_t)n * stride, d_xi));
    TRY(sc.up((int*)nullptr, n, d_sc)); TRY(sc.up((double*)nullptr, 3 * (size_t)n, d_oo)); TRY(sc.up((double*)nullptr, 3 * (size_t)n, d_od));
    TRY(sc.up((double*)nullptr, n, d_ot)); TRY(sc.up((double*)nullptr, 3 * (size_t)n, d_att)); TRY(sc.up((double*)nullptr, 3 * (size_t)n, d_em));
    TRY(sc.up((int*)nullptr, n, d_nd));
    scatter_kernel<<<(n + 127) / 128, 128>>>(vw.ds, mat - 1, n, d_ro, d_rd, d_rt, d_p, d_n, d_front, d_u, d_v, d_xi, stride, d_sc, d_oo, d_od, d_ot, d_att, d_em, d_nd);
    CUDA_TRY(cudaGetLastError());
    TRY(sc.down(osc, d_sc, n)); TRY(sc.down(oo, d_oo, 3 * (size_t)n)); TRY(sc.down(od, d_od, 3 * (size_t)n)); TRY(sc.down(ot, d_ot, n));
    TRY(sc.down(oatt, d_att, 3 * (size_t)n)); TRY(sc.down(oem, d_em, 3 * (size_t)n)); TRY(sc.down(nd, d_nd, n));
    return RTW_OK;
}

int rtw_test_texture(rtw_scene* s, int32_t tex, int32_t n, const double* u, const double* v, const double* p, double* out) {
    if (!s || n <= 0 || !u || !v || !p || !out) return fail(RTW_ERR_INVALID_ARG, "bad argument");
    if (!tex_ok(s, tex)) return fail(RTW_ERR_INVALID_ARG, "bad texture id");
    TRY(need_device());
    View vw; TRY(make_view(s, -1, vw));
    Scratch sc; double *d_u, *d_v, *d_p, *d_out;
    TRY(sc.up(u, n, d_u)); TRY(sc.up(v, n, d_v)); TRY(sc.up(p, 3 * (size_t)n, d_p)); TRY(sc.up((double*)nullptr, 3 * (size_t)n, d_out));
    texture_kernel<<<(n + 127) / 128, 128>>>(vw.ds, tex, n, d_u, d_v, d_p, d_out);
    CUDA_TRY(cudaGetLastError());
    return sc.down(out, d_out, 3 * (size_t)n);
}

int rtw_trace_paths(rtw_scene* s, const rtw_camera* cam, const rtw_render_params* p, int32_t n, const int32_t* px, const int32_t* py,
                    const int32_t* smp, double* out_rgb, int32_t* out_seg) {
    if (!s || !cam || !p || n <= 0 || !px || !py || !smp || !out_rgb || !out_seg) return fail(RTW_ERR_INVALID_ARG, "bad argument");
    TRY(need_device());
    View vw; TRY(make_view(s, -1, vw));
    if (std::min(cam->time0, cam->time1) < s->flat.mov_t0 || std::max(cam->time0, cam->time1) > s->flat.mov_t1)
        return fail(RTW_ERR_INVALID_ARG, "camera shutter [time0, time1] exceeds the [time0, time1] of a MovingSphere in the scene");
    DParams dp; TRY(make_params(*p, 1, dp));
    Scratch sc; int *d_x, *d_y, *d_s, *d_seg; double* d_rgb;
    TRY(sc.up(px, n, d_x)); TRY(sc.up(py, n, d_y)); TRY(sc.up(smp, n, d_s)); TRY(sc.up((double*)nullptr, 3 * (size_t)n, d_rgb)); TRY(sc.up((int*)nullptr, n, d_seg));
    trace_paths_kernel<<<(n + 127) / 128, 128>>>(vw.ds, to_dcamera(*cam), dp, n, d_x, d_y, d_s, d_rgb, d_seg);
    CUDA_TRY(cudaGetLastError());
    TRY(sc.down(out_rgb, d_rgb, 3 * (size_t)n)); TRY(sc.down(out_seg, d_seg, n));
    return RTW_OK;
}

// host-only introspection used by the CPU tests (no device needed): flatten + structural BVH check
int rtw_debug_flatten(rtw_scene* s, int32_t* out_counts /* prims, bvh_prims, nodes, xforms, media, mats, texs, depth */, double* out_sah) {
    if (!s) return fail(RTW_ERR_INVALID_ARG, "null scene");
    rtw::FlatScene f; std::string err;
    int rc = rtw::flatten(s->g, s->g.world, f, err);
    if (rc) return fail(rc, err);
    if (!(f.wide ? rtw::validate_wide(f, err) : rtw::validate_bvh(f, err))) return fail(RTW_ERR_INVALID_ARG, "invalid BVH: " + err);
    if (out_counts) {
        out_counts[0] = (int)f.prims.size(); out_counts[1] = f.n_bvh_prims; out_counts[2] = (int)(f.wide ? f.wnodes.size() : f.nodes.size()); out_counts[3] = (int)f.xforms.size();
        out_counts[4] = (int)f.media.size(); out_counts[5] = (int)f.mats.size(); out_counts[6] = (int)f.texs.size(); out_counts[7] = f.max_depth;
    }
    if (out_sah) *out_sah = f.sah_cost;
    return RTW_OK;
}
// Host-only check of the 8-wide compressed BVH (no device needed): flatten with wide nodes, validate the structure, then
// run `n_rays` seeded rays through the quantised traversal ON THE CPU with the device's arithmetic (bvh_wide.h) and require
// that it reaches every primitive whose box a ray really crosses.
// out[8]: rays, node visits, leaves reached, boxes really crossed, boxes MISSED (0), wide nodes, wide depth, bvh prims
int rtw_debug_wide(rtw_scene* s, int32_t n_rays, uint64_t seed, uint64_t* out) {
    if (!s || !out || n_rays < 0) return fail(RTW_ERR_INVALID_ARG, "bad argument");
    rtw::FlatScene f; std::string err;
    rtw::FlattenOptions opt; opt.bvh_width = 8; opt.keep_boxes = true;
    int rc = rtw::flatten(s->g, s->g.world, f, err, opt);
    if (rc) return fail(rc, err);
    if (!rtw::validate_wide(f, err)) return fail(RTW_ERR_INVALID_ARG, "invalid wide BVH: " + err);
    uint64_t st[5];
    const bool ok = rtw::check_wide_traversal(f, n_rays, seed, st, err);
    for (int i = 0; i < 5; ++i) out[i] = st[i];
    out[5] = f.wnodes.size(); out[6] = (uint64_t)f.wide_depth; out[7] = (uint64_t)f.n_bvh_prims;
    if (!ok) return fail(RTW_ERR_INVALID_ARG, err);
    return RTW_OK;
}
// tuning aid (host-only): cost of secondary-like rays through the wide tree; out[8]: rays, node visits, primitive tests,
// occupied slots of the visited nodes, hits, wide nodes, wide depth, bvh prims
int rtw_debug_wide_cost(rtw_scene* s, int32_t n_rays, uint64_t seed, uint64_t* out) {
    if (!s || !out || n_rays < 0) return fail(RTW_ERR_INVALID_ARG, "bad argument");
    rtw::FlatScene f; std::string err;
    rtw::FlattenOptions opt; opt.bvh_width = 8;
    int rc = rtw::flatten(s->g, s->g.world, f, err, opt);
    if (rc) return fail(rc, err);
    uint64_t st[5];
    rtw::wide_cost_probe(f, n_rays, seed, st);
    for (int i = 0; i < 5; ++i) out[i] = st[i];
    out[5] = f.wnodes.size(); out[6] = (uint64_t)f.wide_depth; out[7] = (uint64_t)f.n_bvh_prims;
    return RTW_OK;
}
// same, with room to grow: out_counts[16] = the 8 above, [8] BvhNode members dropped as clones of an earlier member
int rtw_debug_flatten2(rtw_scene* s, int32_t* out_counts, double* out_sah) {
    if (!s || !out_counts) return fail(RTW_ERR_INVALID_ARG, "null argument");
    rtw::FlatScene f; std::string err;
    int rc = rtw::flatten(s->g, s->g.world, f, err);
    if (rc) return fail(rc, err);
    if (!(f.wide ? rtw::validate_wide(f, err) : rtw::validate_bvh(f, err))) return fail(RTW_ERR_INVALID_ARG, "invalid BVH: " + err);
    std::memset(out_counts, 0, 16 * sizeof(int32_t));
    out_counts[0] = (int)f.prims.size(); out_counts[1] = f.n_bvh_prims; out_counts[2] = (int)(f.wide ? f.wnodes.size() : f.nodes.size()); out_counts[3] = (int)f.xforms.size();
    out_counts[4] = (int)f.media.size(); out_counts[5] = (int)f.mats.size(); out_counts[6] = (int)f.texs.size(); out_counts[7] = f.max_depth;
    out_counts[8] = f.n_dedup; out_counts[9] = f.wide ? 8 : 2; out_counts[10] = f.wide_depth;
    if (out_sah) *out_sah = f.sah_cost;
    return RTW_OK;
}

}  // extern "C"
