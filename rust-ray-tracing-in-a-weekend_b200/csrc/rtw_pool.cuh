// rtw_pool.cuh — the "warp-pool" render kernel: a wavefront pipeline that never leaves the SM.
//
// The plain megakernel (render_kernel) keeps one path per lane; ncu shows 14.5 of 32 lanes active per instruction:
// rays of one warp need very different traversal lengths, 37 % of the lanes regenerate every iteration while the
// rest wait, each material runs its own divergent branch, and the rejection loops of src/math.rs:51-76 finish when
// the unluckiest lane does.  Here every WARP owns a pool of POOL path slots in shared memory (B200: 227 KB/SM) and
// iterates over the four stages of the north star as dense loops over compacted slot lists:
//
//   regenerate : free slots <- new (pixel, sample) camera rays           (src/main.rs:517-520, camera.rs:58-66)
//   traverse   : persistent while-while BVH traversal; a lane that finishes its ray writes (t, prim), classifies the
//                hit material and immediately fetches the next slot of the list (Aila-Laine dynamic fetch), so the
//                lanes stay busy regardless of per-ray traversal length                     (hittable.rs:43-55)
//   shade      : one dense loop per material kind over its slot list (lists are filled at traversal end with
//                shared-memory atomics = material sort), scatter + texture, next ray written back    (material.rs)
//   accumulate : miss / emission contributions go straight into the warp's tile accumulator         (main.rs:28-37)
//
// Queues never touch HBM: the only global traffic is the read-only scene (L1-resident for C1) and one tile write.
#ifndef RTW_POOL_CUH
#define RTW_POOL_CUH

#include "rtw_device.cuh"

namespace rtwd {

#define POOL_FREE 0xffffffffu
enum { K_LAMB = 0, K_METAL = 1, K_DIEL = 2, K_LIGHT = 3, K_ISO = 4, K_MISS = 5, K_NKIND = 6 };
enum { C_TRAV = 6, C_FREE = 7 };   // counter slots after the kinds

template <int POOL>
struct PoolSmem {
    float ox[POOL], oy[POOL], oz[POOL], dx[POOL], dy[POOL], dz[POOL], tm[POOL];
    float tr[POOL], tg[POOL], tb[POOL];          // throughput
    float t_hit[POOL];
    int prim[POOL];                              // hit prim, -1 = miss, -2 - mat = medium scatter with material mat
    int last[POOL];                              // primitive the ray starts on
    unsigned meta[POOL];                         // pix(5) | segment(6) << 5 | media draws(4) << 11 | sample(17) << 15
    unsigned char list[K_NKIND][POOL];
    unsigned char trav[POOL], freel[POOL];
    __align__(16) float acc[96];
    int cnt[8];
};

RTW_DEV unsigned meta_pack(int pix, int seg, int ndraw, int sample) { return (unsigned)pix | ((unsigned)seg << 5) | ((unsigned)ndraw << 11) | ((unsigned)sample << 15); }

// One accepted closest hit -> HitRec + material, then Material::scatter.  Writes the continuing ray back into the
// slot (and queues it for traversal) or frees the slot.  KIND is a compile-time constant: each list runs its own
// straight-line code.
template <int POOL, int KIND>
RTW_DEV void shade_slot(PoolSmem<POOL>& P, int slot, const DScene& sc, const DParams& prm, int tile_x0, int tile_y0) {
    const unsigned meta = P.meta[slot];
    const int pix = meta & 31, seg = (meta >> 5) & 63, ndraw = (meta >> 11) & 15, sample = meta >> 15;
    V3 T = mk(P.tr[slot], P.tg[slot], P.tb[slot]);
    if (KIND == K_MISS) {                                                                  // main.rs:37
        atomicAdd(&P.acc[pix * 3 + 0], T.x * prm.bg_r);
        atomicAdd(&P.acc[pix * 3 + 1], T.y * prm.bg_g);
        atomicAdd(&P.acc[pix * 3 + 2], T.z * prm.bg_b);
        P.meta[slot] = POOL_FREE;
        P.freel[atomicAdd(&P.cnt[C_FREE], 1)] = (unsigned char)slot;
        return;
    }
    Ray ray; ray.o = mk(P.ox[slot], P.oy[slot], P.oz[slot]); ray.d = mk(P.dx[slot], P.dy[slot], P.dz[slot]); ray.time = P.tm[slot];
    const int prim = P.prim[slot];
    const float t = P.t_hit[slot];
    HitRec rec; DMatRec m;
    if (prim <= -2) {                                                                      // medium scatter hittable.rs:460-464
        rec.t = t; rec.p = ray_at(ray, t); rec.normal = mk(1.f, 0.f, 0.f); rec.front = 1; rec.u = 0.f; rec.v = 0.f; rec.mat = -2 - prim;
        m = load_mat(sc, rec.mat);
    } else {
        m = load_mat(sc, __ldg(&sc.prims[prim].mat));
        TRay tr = make_tray(ray);
        finalize_hit(sc, prim, t, tr, mat_needs_uv(sc, m), rec);
    }
    PhiloxRng g;
    const int x = tile_x0 + (pix & 7), y = tile_y0 + (pix >> 3);
    g.init(prm, (uint32_t)(y * prm.width + x), (uint32_t)sample);
    g.bounce = (uint32_t)seg; g.draw = (uint32_t)ndraw;
    if (ndraw & 3) philox4x32_10_rk(g.rk, (uint32_t)ndraw >> 2, g.bounce, g.pixel, g.sample, g.w0, g.w1, g.w2, g.w3);
    Ray sc_ray; V3 att, em;
    m.kind = KIND;                         // list membership == material kind: lets the compiler drop the other branches
    bool cont = scatter(sc, m, ray, rec, g, sc_ray, att, em);
    if (KIND == K_LIGHT) {                                                                 // main.rs:28, :32-33
        atomicAdd(&P.acc[pix * 3 + 0], T.x * em.x);
        atomicAdd(&P.acc[pix * 3 + 1], T.y * em.y);
        atomicAdd(&P.acc[pix * 3 + 2], T.z * em.z);
    }
    if (!cont || seg >= prm.max_depth) {                                                   // absorbed / light / depth exhausted (main.rs:21-23)
        P.meta[slot] = POOL_FREE;
        P.freel[atomicAdd(&P.cnt[C_FREE], 1)] = (unsigned char)slot;
        return;
    }
    P.ox[slot] = sc_ray.o.x; P.oy[slot] = sc_ray.o.y; P.oz[slot] = sc_ray.o.z;
    P.dx[slot] = sc_ray.d.x; P.dy[slot] = sc_ray.d.y; P.dz[slot] = sc_ray.d.z;
    P.tr[slot] = T.x * att.x; P.tg[slot] = T.y * att.y; P.tb[slot] = T.z * att.z;
    P.last[slot] = prim >= 0 ? prim : -1;
    P.meta[slot] = meta_pack(pix, seg + 1, 0, sample);
    P.trav[atomicAdd(&P.cnt[C_TRAV], 1)] = (unsigned char)slot;
}

template <int POOL, int KIND>
RTW_DEV void shade_list(PoolSmem<POOL>& P, int lane, const DScene& sc, const DParams& prm, int tile_x0, int tile_y0) {
    const int n = P.cnt[KIND];
    for (int i = lane; i < n; i += 32) shade_slot<POOL, KIND>(P, P.list[KIND][i], sc, prm, tile_x0, tile_y0);
}

#define RTW_FETCH_THRESHOLD 24      // leave the traversal loop to refill idle lanes once fewer lanes than this are busy

// The traverse stage over P.trav[0..n_trav).  Dynamic fetch: lanes pull the next slot as soon as their ray is done.
// W = 0: binary nodes (speculative while-while); W = 1: the 8-wide compressed BVH (bvh_wide.h) — its visits are long
// (8 quantised slab tests) and few, which is where keeping every lane supplied with rays pays most.
template <int POOL, int W>
RTW_DEV unsigned long long traverse_stage(PoolSmem<POOL>& P, int lane, unsigned lt_mask, int n_trav, const DScene& sc, const DParams& prm,
                                          int tile_x0, int tile_y0) {
    unsigned long long rays = 0;
    int q_next = 0;
    bool active = false;
    int slot = 0;
    TRay r; V3 inv = mk(0, 0, 0), cmn = mk(0, 0, 0), cmx = mk(0, 0, 0);
    float t_best = 0.f; int prim_best = -1, skip = -1;
    // binary: stack of node ids (sentinel at the bottom); wide: stack of (base, group) pairs
    uint32_t stack[RTW_STACK]; int sp = 1, node = RTW_SENTINEL, leaf = 0;
    uint32_t wbase = 0, wgrp = 0; rtww::WRay wr; wr.one = 0x3F800000u | ((uint32_t)sc.n_nodes >> 31); wr.k = 0; wr.sx = wr.sy = wr.sz = 0.f;
    wr.ix = wr.iy = wr.iz = wr.oix = wr.oiy = wr.oiz = 0.f;
    bool done = true;
    stack[0] = (uint32_t)RTW_SENTINEL;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle && q_next < n_trav) {
            if (!active) {
                const int k = q_next + __popc(idle & lt_mask);
                if (k < n_trav) {
                    slot = P.trav[k];
                    Ray ray; ray.o = mk(P.ox[slot], P.oy[slot], P.oz[slot]); ray.d = mk(P.dx[slot], P.dy[slot], P.dz[slot]); ray.time = P.tm[slot];
                    r = make_tray(ray);
                    slab_setup(r.o, r.d, inv, cmn, cmx);
                    t_best = CUDART_INF_F; prim_best = -1; skip = P.last[slot];
                    if (W) {
                        slab_setup_wide(r.o, r.d, wr);
                        sp = 0; wbase = 0; wgrp = sc.n_bvh_prims ? ((1u << 8) | (1u << wr.k)) : 0u; done = false;
                    } else { sp = 1; leaf = 0; node = sc.n_bvh_prims ? 0 : RTW_SENTINEL; }
                    active = true; ++rays;
                }
            }
            q_next += __popc(idle);
        }
        if (!__any_sync(0xffffffffu, active)) break;
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
                        const float t = prim_root(sc, pi, r, prm.t_min, t_best, skip, hp);
                        if (t == t) { t_best = t; prim_best = hp; }
                    }
                    wbase = m.x; wgrp = (imask << 8) | (m16 & 0xffu);
                    if (q_next < n_trav && __popc(__activemask()) < RTW_FETCH_THRESHOLD) break;   // refill idle lanes
                }
                node = done ? RTW_SENTINEL : 0;                                                    // (the result code below keys on node)
            } else {
            while (node != RTW_SENTINEL) {
                bool searching = true;
                while (node >= 0 && node != RTW_SENTINEL) {
                    const float4* np = reinterpret_cast<const float4*>(sc.nodes + node);
                    float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
                    int4 n3 = __ldg(reinterpret_cast<const int4*>(np + 3));
                    float e0, e1;
                    bool h0 = slab(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, inv, cmn, cmx, prm.t_min, t_best, e0);
                    bool h1 = slab(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, inv, cmn, cmx, prm.t_min, t_best, e1);
                    if (!h0 && !h1) node = (int)stack[--sp];
                    else {
                        node = h0 ? n3.x : n3.y;
                        if (h0 && h1) {
                            int farc = n3.y;
                            if (e1 < e0) { farc = node; node = n3.y; }
                            stack[sp++] = (uint32_t)farc;
                        }
                    }
                    if (node < 0 && leaf >= 0) { searching = false; leaf = node; node = (int)stack[--sp]; }
                    if (!__any_sync(__activemask(), searching)) break;
                }
                while (leaf < 0) {
                    int code = ~leaf, first = code >> 3, count = (code & 7) + 1;
                    for (int i = 0; i < count; ++i) {
                        int hp;
                        float t = prim_root(sc, first + i, r, prm.t_min, t_best, skip, hp);
                        if (t == t) { t_best = t; prim_best = hp; }
                    }
                    leaf = node;
                    if (node < 0) node = (int)stack[--sp];
                }
                if (q_next < n_trav && __popc(__activemask()) < RTW_FETCH_THRESHOLD) break;   // refill idle lanes
            }
            }
            if (node == RTW_SENTINEL) {
                // ray done: media (list order, after the surfaces), classification, result
                const unsigned meta = P.meta[slot];
                int ndraw = 0, med_mat = -1;
                if (sc.n_media) {
                    const int pix = meta & 31, seg = (meta >> 5) & 63, sample = meta >> 15;
                    PhiloxRng g;
                    g.init(prm, (uint32_t)((tile_y0 + (pix >> 3)) * prm.width + tile_x0 + (pix & 7)), (uint32_t)sample);
                    g.set_bounce((uint32_t)seg);
                    for (int mi = 0; mi < sc.n_media; ++mi) {
                        float t; int mat;
                        if (medium_hit(sc, mi, r, prm.t_min, t_best, g, t, mat)) { t_best = t; med_mat = mat; prim_best = -2; }
                    }
                    ndraw = (int)g.draw;
                }
                int kind, prim_out = prim_best;
                if (prim_best == -1) kind = K_MISS;
                else {
                    int mat = prim_best == -2 ? med_mat : __ldg(&sc.prims[prim_best].mat);
                    kind = __ldg(&sc.mats[mat].kind);
                    if (prim_best == -2) prim_out = -2 - med_mat;
                }
                P.t_hit[slot] = t_best; P.prim[slot] = prim_out;
                P.meta[slot] = (meta & ~(15u << 11)) | ((unsigned)ndraw << 11);
                P.list[kind][atomicAdd(&P.cnt[kind], 1)] = (unsigned char)slot;
                active = false;
            }
        }
    }
    return rays;
}

}  // namespace rtwd
#endif
