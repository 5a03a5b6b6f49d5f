// bvh_wide.h — the 8-wide compressed BVH: node layout, BVH2 -> BVH8 collapse, quantisation.  Shared by the host
// flattener (scene_host.cpp), the device builder (bvh_build.cu) and the traversal (rtw_device.cuh); every function
// here compiles for both sides, so the CPU tests exercise the code the GPU runs.
//
// Replaces, for scenes that do not fit the caches, the 64-byte two-box BVH2 node (rtw_types.h DNode): a ray of the
// 1 M - 16 M sphere sweep chased ~50 dependent 64-byte node fetches (7 % of HBM bandwidth, latency bound).
// The reference's own structure is the binary BvhNode of src/hittable.rs:77-130 / :290-306 (and no acceleration at all
// over the world list, src/main.rs:25).
//
// Layout after Ylitie, Karras, Laine, "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs" (HPG
// 2017): 80 bytes for up to 8 children, child boxes quantised to 8 bits per plane on a per-node power-of-two grid
// anchored at `p`, inner children stored contiguously from `child_base`, leaf primitives contiguously from
// `prim_base` (one primitive per leaf: in a warp only ~5 lanes reach a leaf together).  Children sit in
// the slot whose index bits say on which side of the node they lie (bit 0: +x, bit 1: +y, bit 2: +z), so a ray visits
// slots in the order  slot ^ octant  without sorting anything.
#ifndef RTW_BVH_WIDE_H
#define RTW_BVH_WIDE_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RTW_HD __host__ __device__ __forceinline__
#else
#define RTW_HD inline
#endif

// 80 B = 5 x 16 B.  Child c's box on axis a: [p[a] + qlo[a][c] * 2^(e[a]-127), p[a] + qhi[a][c] * 2^(e[a]-127)].
struct
#if defined(__CUDACC__)
    __align__(16)
#else
    alignas(16)
#endif
    DWNode {
    float px, py, pz;           // grid anchor (a little below the node's min corner: every child plane has padding)
    uint8_t ex, ey, ez;         // biased float exponents of the grid step per axis
    uint8_t imask;              // bit s: slot s holds an inner node
    uint32_t child_base;        // index of the first inner child; child in slot s = child_base + popc(imask & below(s))
    uint32_t prim_base;         // index of the first leaf primitive; leaf in slot s = prim_base + popc(lmask & below(s))
    uint8_t lmask;              // bit s: slot s holds a leaf (one primitive)
    uint8_t pad[7];
    uint8_t qlox[8], qloy[8], qloz[8], qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(DWNode) == 80, "DWNode is five 16-byte words");

#define RTW_WIDE 8
#define RTW_WIDE_STACK 32           // traversal stack entries = levels of the wide tree a ray can be inside of
// Outward padding of every quantised plane, in grid steps.  The traversal evaluates a plane distance as one FFMA on
// a byte placed in a float's mantissa (rtw_device.cuh wnode_hits): its rounding error is < 2^-9 step; 2^-6 covers it
// eight times over and costs 1.6 % of a step in tightness (the quantisation itself costs half a step on average).
#define RTW_WIDE_PADQ 0.015625f

namespace rtww {

// Binary BVH the collapse reads.  Refs < n_inner are inner nodes; ref >= n_inner is leaf (ref - n_inner), one primitive,
// numbered in the order the builder left them.  box: 6 floats per ref (min xyz, max xyz), already rounded outward.
struct B2View {
    const float* box;
    const int* left;
    const int* right;
    int n_inner, n_leaves;
    const int* count;       // leaves under each inner ref (nullptr: plain greedy collapse, no absorption of small subtrees)
};

struct WideItem { int b2; int wide; int depth; };

RTW_HD float box_area(const float* b) {
    const float dx = b[3] - b[0], dy = b[4] - b[1], dz = b[5] - b[2];
    return dx * dy + dy * dz + dz * dx;
}

// Per-axis grid for a node box [mn, mx]: smallest power-of-two step with  extent / step + 3 PADQ <= 255,  anchor
// p = mn - 2 PADQ step (so that floor(x - PADQ) >= 0 for every child plane and even the child that defines the node's
// min keeps PADQ of padding).  Returns the biased exponent.
RTW_HD int wide_axis_grid(float mn, float mx, float& p, float& step) {
    const float extent = fmaxf(mx - mn, 0.0f);
    int e = -100;                                   // degenerate (flat) node: any tiny normal step
    if (extent > 0.0f) {
        int ex; const float fr = frexpf(extent * (1.0f / 254.0f), &ex);     // extent / 254 = fr * 2^ex, fr in [0.5, 1)
        e = (fr == 0.5f) ? ex - 1 : ex;
        if (e < -100) e = -100;
    }
    for (;;) {
        step = ldexpf(1.0f, e);
        p = mn - 2.0f * RTW_WIDE_PADQ * step;
        if (p > mn) p = nextafterf(p, -INFINITY);
        if ((mx - p) / step + RTW_WIDE_PADQ <= 254.5f) break;
        ++e;
    }
    return e + 127;
}
RTW_HD uint8_t wide_qlo(float v, float p, float step) {
    float q = floorf((v - p) / step - RTW_WIDE_PADQ);
    return (uint8_t)(q < 0.0f ? 0.0f : (q > 255.0f ? 255.0f : q));
}
RTW_HD uint8_t wide_qhi(float v, float p, float step) {
    float q = ceilf((v - p) / step + RTW_WIDE_PADQ);
    return (uint8_t)(q < 0.0f ? 0.0f : (q > 255.0f ? 255.0f : q));
}

// Decode (tests, tile culling): child s of node n -> float box, rounded outward.
RTW_HD void wide_child_box(const DWNode& n, int s, float out[6]) {
    const float sx = ldexpf(1.0f, (int)n.ex - 127), sy = ldexpf(1.0f, (int)n.ey - 127), sz = ldexpf(1.0f, (int)n.ez - 127);
    out[0] = nextafterf(n.px + (float)n.qlox[s] * sx, -INFINITY); out[3] = nextafterf(n.px + (float)n.qhix[s] * sx, INFINITY);
    out[1] = nextafterf(n.py + (float)n.qloy[s] * sy, -INFINITY); out[4] = nextafterf(n.py + (float)n.qhiy[s] * sy, INFINITY);
    out[2] = nextafterf(n.pz + (float)n.qloz[s] * sz, -INFINITY); out[5] = nextafterf(n.pz + (float)n.qhiz[s] * sz, INFINITY);
}

// One wide node: open the binary subtree under `it.b2` into up to 8 children (always the child with the largest surface
// area next — the greedy form of the SAH-optimal collapse), place them in octant slots, quantise, and hand the inner
// children on.  `node_count` / `prim_count` allocate the children's contiguous ranges (atomics on the device, plain
// counters on the host); `leaf_order[k]` receives the binary tree's leaf number of the primitive that ends up at
// position k of the final primitive order; `next[...]` receives the work items of the inner children.
template <class Alloc>
RTW_HD void collapse_one(const B2View& v, const WideItem& it, DWNode* nodes, int* leaf_order, WideItem* next, Alloc& alloc) {
    int ch[RTW_WIDE];
    int n = 2;
    ch[0] = v.left[it.b2]; ch[1] = v.right[it.b2];
    // Which child to open next.  BIG subtrees (more than 8 leaves) by surface area, largest first — the greedy form of the
    // SAH-optimal collapse.  A SMALL subtree is never opened half-way: it is either absorbed completely — all its leaves
    // become slots of this node — when the free slots suffice, or left whole as one child node.  (Plain area-greedy opening
    // left the bottom of the tree full of nodes with 2-3 leaves: 1 M spheres gave 300 k nodes with 4.6 of 8 slots used; a
    // leaf slot costs the parent the same box test as the small node did, and the node visit itself disappears.)
    while (n < RTW_WIDE) {
        int best = -1; float best_area = -1.0f;
        for (int i = 0; i < n; ++i)
            if (ch[i] < v.n_inner && (!v.count || v.count[ch[i]] > RTW_WIDE)) {
                const float a = box_area(v.box + 6 * (size_t)ch[i]);
                if (a > best_area) { best_area = a; best = i; }
            }
        if (best >= 0) {
            const int open = ch[best];
            ch[best] = v.left[open];
            ch[n++] = v.right[open];
            continue;
        }
        if (!v.count) break;
        for (int i = 0; i < n; ++i)
            if (ch[i] < v.n_inner && v.count[ch[i]] - 1 <= RTW_WIDE - n) {
                const float a = box_area(v.box + 6 * (size_t)ch[i]);
                if (a > best_area) { best_area = a; best = i; }
            }
        if (best < 0) break;
        int stk[RTW_WIDE + 1]; int ns = 0; bool first = true;
        stk[ns++] = ch[best];
        while (ns) {
            const int r = stk[--ns];
            if (r >= v.n_inner) { if (first) { ch[best] = r; first = false; } else ch[n++] = r; }
            else { stk[ns++] = v.right[r]; stk[ns++] = v.left[r]; }
        }
    }
    // octant slots: child i prefers the slot whose sign pattern matches where its centre lies relative to the node's
    // centre; greedy assignment by the largest remaining (child, slot) score
    const float* nb = v.box + 6 * (size_t)it.b2;
    const float cx = 0.5f * (nb[0] + nb[3]), cy = 0.5f * (nb[1] + nb[4]), cz = 0.5f * (nb[2] + nb[5]);
    float dx[RTW_WIDE], dy[RTW_WIDE], dz[RTW_WIDE];
    for (int i = 0; i < n; ++i) {
        const float* b = v.box + 6 * (size_t)ch[i];
        dx[i] = 0.5f * (b[0] + b[3]) - cx; dy[i] = 0.5f * (b[1] + b[4]) - cy; dz[i] = 0.5f * (b[2] + b[5]) - cz;
    }
    int slot_of[RTW_WIDE]; int child_in[RTW_WIDE];
    for (int s = 0; s < RTW_WIDE; ++s) child_in[s] = -1;
    for (int i = 0; i < n; ++i) slot_of[i] = -1;
    for (int round = 0; round < n; ++round) {
        int bi = -1, bs = -1; float bscore = -INFINITY;
        for (int i = 0; i < n; ++i) {
            if (slot_of[i] >= 0) continue;
            for (int s = 0; s < RTW_WIDE; ++s) {
                if (child_in[s] >= 0) continue;
                const float score = ((s & 1) ? dx[i] : -dx[i]) + ((s & 2) ? dy[i] : -dy[i]) + ((s & 4) ? dz[i] : -dz[i]);
                if (score > bscore) { bscore = score; bi = i; bs = s; }
            }
        }
        slot_of[bi] = bs; child_in[bs] = bi;
    }
    DWNode w;
    float p[3], step[3];
    w.ex = (uint8_t)wide_axis_grid(nb[0], nb[3], p[0], step[0]);
    w.ey = (uint8_t)wide_axis_grid(nb[1], nb[4], p[1], step[1]);
    w.ez = (uint8_t)wide_axis_grid(nb[2], nb[5], p[2], step[2]);
    w.px = p[0]; w.py = p[1]; w.pz = p[2];
    int n_inner = 0, n_leaf = 0;
    unsigned imask = 0, lmask = 0;
    for (int s = 0; s < RTW_WIDE; ++s) {
        const int i = child_in[s];
        if (i < 0) continue;
        if (ch[i] < v.n_inner) { imask |= 1u << s; ++n_inner; } else { lmask |= 1u << s; ++n_leaf; }
    }
    w.imask = (uint8_t)imask; w.lmask = (uint8_t)lmask;
    for (int k = 0; k < 7; ++k) w.pad[k] = 0;
    const int child_base = alloc.nodes(n_inner), prim_base = alloc.prims(n_leaf), next_base = alloc.queue(n_inner);
    w.child_base = (uint32_t)child_base; w.prim_base = (uint32_t)prim_base;
    int ki = 0, kl = 0;
    for (int s = 0; s < RTW_WIDE; ++s) {
        const int i = child_in[s];
        if (i < 0) {            // empty slot: inverted box (and its bit is in neither mask)
            w.qlox[s] = w.qloy[s] = w.qloz[s] = 255; w.qhix[s] = w.qhiy[s] = w.qhiz[s] = 0;
            continue;
        }
        const float* b = v.box + 6 * (size_t)ch[i];
        w.qlox[s] = wide_qlo(b[0], p[0], step[0]); w.qhix[s] = wide_qhi(b[3], p[0], step[0]);
        w.qloy[s] = wide_qlo(b[1], p[1], step[1]); w.qhiy[s] = wide_qhi(b[4], p[1], step[1]);
        w.qloz[s] = wide_qlo(b[2], p[2], step[2]); w.qhiz[s] = wide_qhi(b[5], p[2], step[2]);
        if (ch[i] < v.n_inner) {
            WideItem c; c.b2 = ch[i]; c.wide = child_base + ki; c.depth = it.depth + 1;
            next[next_base + ki] = c; ++ki;
        } else { leaf_order[prim_base + kl] = ch[i] - v.n_inner; ++kl; }
    }
    nodes[it.wide] = w;
    alloc.depth(it.depth + 1);
}

// ---- ray x wide node -----------------------------------------------------------------------------------------------
struct W4 { uint32_t x, y, z, w; };              // one 16-byte word of a node (same layout as CUDA's uint4)
struct WRay {
    float ix, iy, iz;                            // 1 / d (zero components replaced, rtw_device.cuh slab_dir)
    float oix, oiy, oiz;                         // o / d
    float sx, sy, sz;                            // per-axis absolute slack for the rounding of o / d: 4 ulp of |o/d| (slab_setup_wide)
    uint32_t k;                                  // 7 ^ octant, octant bit a = (d[a] < 0)
    uint32_t one;                                // 0x3F800000 held in a REGISTER: PRMT takes one immediate, and it must be
                                                 // the byte selector (with the constant as immediate every PRMT needs a MOV)
};

RTW_HD float wide_bits_to_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
// byte c of `word` placed in mantissa bits 8..15 of 1.0f:  1 + q * 2^-15  (one PRMT on the device)
template <int C> RTW_HD float wide_plane_f(uint32_t word, uint32_t one) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__byte_perm(word, one, 0x7604u | (C << 4)));
#else
    return wide_bits_to_float(one | (((word >> (8 * C)) & 0xffu) << 8));
#endif
}

// Slab test of the ray against the 8 quantised child boxes -> bit s set when slot s may be hit in [t_lo, t_hi].
// A plane at p + q step is at ray parameter  t = (p - o) / d + q step / d.  With f = 1 + q 2^-15 taken straight from
// the byte (wide_plane_f), t = f A + B,  A = 2^15 step / d,  B = (p - o) / d - A: ONE FFMA per plane, no integer ->
// float conversion.  Rounding: B carries 2^-24 |A| = 2^-9 step / d (covered by RTW_WIDE_PADQ at build time) plus errors
// relative to t (covered by the 10-ulp exit padding) and to o / d (covered per axis by sx, sy, sz, like the BVH2 slab test).
// Near / far planes are picked per AXIS for four children at a time (the sign of d selects the lo or the hi word).
template <int C> RTW_HD uint32_t wide_child_hit(uint32_t nx, uint32_t ny, uint32_t nz, uint32_t fx, uint32_t fy, uint32_t fz, float Ax, float Ay,
                                               float Az, float Bnx, float Bny, float Bnz, float Bfx, float Bfy, float Bfz, float t_lo, float t_hi, uint32_t one) {
    const float tnx = fmaf(wide_plane_f<C>(nx, one), Ax, Bnx), tny = fmaf(wide_plane_f<C>(ny, one), Ay, Bny), tnz = fmaf(wide_plane_f<C>(nz, one), Az, Bnz);
    const float tfx = fmaf(wide_plane_f<C>(fx, one), Ax, Bfx), tfy = fmaf(wide_plane_f<C>(fy, one), Ay, Bfy), tfz = fmaf(wide_plane_f<C>(fz, one), Az, Bfz);
    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, t_lo));
    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, t_hi)) * 1.0000012f;
    return tn <= tf ? 1u : 0u;
}
RTW_HD uint32_t wide_node_hits(const W4& h, const W4& qa, const W4& qb, const W4& qc, const WRay& r, float t_lo, float t_hi) {
    const float px = wide_bits_to_float(h.x), py = wide_bits_to_float(h.y), pz = wide_bits_to_float(h.z);
    const float Ax = r.ix * wide_bits_to_float(((h.w & 0xffu) + 15u) << 23);
    const float Ay = r.iy * wide_bits_to_float((((h.w >> 8) & 0xffu) + 15u) << 23);
    const float Az = r.iz * wide_bits_to_float((((h.w >> 16) & 0xffu) + 15u) << 23);
    const float Bx = fmaf(px, r.ix, -r.oix) - Ax, By = fmaf(py, r.iy, -r.oiy) - Ay, Bz = fmaf(pz, r.iz, -r.oiz) - Az;
    const float Bnx = Bx - r.sx, Bny = By - r.sy, Bnz = Bz - r.sz, Bfx = Bx + r.sx, Bfy = By + r.sy, Bfz = Bz + r.sz;   // near planes earlier, far planes later
    const bool gx = r.ix < 0.0f, gy = r.iy < 0.0f, gz = r.iz < 0.0f;
    // words: qa = lox[0-3] lox[4-7] loy[0-3] loy[4-7];  qb = loz[0-3] loz[4-7] hix[0-3] hix[4-7];  qc = hiy hiy hiz hiz
    const uint32_t nx0 = gx ? qb.z : qa.x, nx1 = gx ? qb.w : qa.y, fx0 = gx ? qa.x : qb.z, fx1 = gx ? qa.y : qb.w;
    const uint32_t ny0 = gy ? qc.x : qa.z, ny1 = gy ? qc.y : qa.w, fy0 = gy ? qa.z : qc.x, fy1 = gy ? qa.w : qc.y;
    const uint32_t nz0 = gz ? qc.z : qb.x, nz1 = gz ? qc.w : qb.y, fz0 = gz ? qb.x : qc.z, fz1 = gz ? qb.y : qc.w;
    uint32_t hits = 0;
    hits |= wide_child_hit<0>(nx0, ny0, nz0, fx0, fy0, fz0, Ax, Ay, Az, Bnx, Bny, Bnz, Bfx, Bfy, Bfz, t_lo, t_hi, r.one);
    hits |= wide_child_hit<1>(nx0, ny0, nz0, fx0, fy0, fz0, Ax, Ay, Az, Bnx, Bny, Bnz, Bfx, Bfy, Bfz, t_lo, t_hi, r.one) << 1;
    hits |= wide_child_hit<2>(nx0, ny0, nz0, fx0, fy0, fz0, Ax, Ay, Az, Bnx, Bny, Bnz, Bfx, Bfy, Bfz, t_lo, t_hi, r.one) << 2;
    hits |= wide_child_hit<3>(nx0, ny0, nz0, fx0, fy0, fz0, Ax, Ay, Az, Bnx, Bny, Bnz, Bfx, Bfy, Bfz, t_lo, t_hi, r.one) << 3;
    hits |= wide_child_hit<0>(nx1, ny1, nz1, fx1, fy1, fz1, Ax, Ay, Az, Bnx, Bny, Bnz, Bfx, Bfy, Bfz, t_lo, t_hi, r.one) << 4;
    hits |= wide_child_hit<1>(nx1, ny1, nz1, fx1, fy1, fz1, Ax, Ay, Az, Bnx, Bny, Bnz, Bfx, Bfy, Bfz, t_lo, t_hi, r.one) << 5;
    hits |= wide_child_hit<2>(nx1, ny1, nz1, fx1, fy1, fz1, Ax, Ay, Az, Bnx, Bny, Bnz, Bfx, Bfy, Bfz, t_lo, t_hi, r.one) << 6;
    hits |= wide_child_hit<3>(nx1, ny1, nz1, fx1, fy1, fz1, Ax, Ay, Az, Bnx, Bny, Bnz, Bfx, Bfy, Bfz, t_lo, t_hi, r.one) << 7;
    return hits;
}

// Bit s of each byte -> bit s ^ k (both bytes of a 16-bit pair at once): slot order -> this ray's visiting priority
// (highest bit first).  Three conditional butterfly stages.
RTW_HD uint32_t wide_perm16(uint32_t m, uint32_t k) {
    if (k & 1u) m = ((m & 0x5555u) << 1) | ((m >> 1) & 0x5555u);
    if (k & 2u) m = ((m & 0x3333u) << 2) | ((m >> 2) & 0x3333u);
    if (k & 4u) m = ((m & 0x0f0fu) << 4) | ((m >> 4) & 0x0f0fu);
    return m;
}

}  // namespace rtww

#endif
