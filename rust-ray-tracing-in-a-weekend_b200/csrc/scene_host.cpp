// scene_host.cpp — flatten the reference-shaped scene graph into SoA device records + a linearised BVH.
//
// What is mirrored (semantics) and what is deliberately not (structure):
//  * Box = its 6 rects in new_box order (src/hittable.rs:132-145); closest-of-six == ONE slab test that names the entry / exit
//    face (PRIM_BOX leaf; the six rect records stay behind the BVH primitives and describe the hit).  Flat boxes keep six rect leaves.
//  * Translate / RotateY wrap a subtree (src/hittable.rs:232-247, :386-415): the chain of wrappers above a
//    primitive is composed into one DXform; spheres are baked to world space (a rigid motion of a sphere is a
//    sphere), rects keep object coordinates and the ray is moved into object space at test time.
//  * BvhNode (src/hittable.rs:77-130): membership only.  The reference builder clones the list at every node
//    (:78, O(N^2)) and leaves the world list itself un-accelerated (src/main.rs:25); any valid BVH returns the
//    same closest hit, so ONE binned-SAH BVH is built over every surface primitive of the world.
//  * ConstantMedium (src/hittable.rs:417-473): boundary subtree flattened into a private prim range scanned
//    linearly; media are evaluated after the surface closest hit (same distribution).
#include "scene_host.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <thread>
#include <unordered_map>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "../../include/rtw.h"

namespace rtw {
namespace {

const double PI = 3.1415926535897932385;   // src/math.rs:5

struct Op { bool rot; double s, c; V3d off; };
typedef std::vector<Op> Chain;

struct Box3 {
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    void grow(const double p[3]) { for (int i = 0; i < 3; ++i) { mn[i] = std::min(mn[i], p[i]); mx[i] = std::max(mx[i], p[i]); } }
    void grow(const Box3& b) { for (int i = 0; i < 3; ++i) { mn[i] = std::min(mn[i], b.mn[i]); mx[i] = std::max(mx[i], b.mx[i]); } }
    double area() const {
        double dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0;
        return 2.0 * (dx * dy + dy * dz + dz * dx);
    }
};

struct Composed { double C = 1, S = 0, b[3] = {0, 0, 0}; };

Composed compose(const Chain& ch, std::vector<Composed>* prefix = nullptr) {
    Composed m;
    for (const Op& op : ch) {
        if (!op.rot) { m.b[0] -= op.off.x; m.b[1] -= op.off.y; m.b[2] -= op.off.z; }
        else {
            double C = op.c * m.C - op.s * m.S, S = op.s * m.C + op.c * m.S;
            double bx = op.c * m.b[0] - op.s * m.b[2], bz = op.s * m.b[0] + op.c * m.b[2];
            m.C = C; m.S = S; m.b[0] = bx; m.b[2] = bz;
        }
        if (prefix) prefix->push_back(m);
    }
    return m;
}
// object -> world point / vector
inline void to_world_point(const Composed& m, const double p[3], double out[3]) {
    double x = p[0] - m.b[0], y = p[1] - m.b[1], z = p[2] - m.b[2];
    out[0] = m.C * x + m.S * z; out[1] = y; out[2] = -m.S * x + m.C * z;
}
inline void to_world_vec(const Composed& m, const double v[3], double out[3]) {
    out[0] = m.C * v[0] + m.S * v[2]; out[1] = v[1]; out[2] = -m.S * v[0] + m.C * v[2];
}

struct Flattener {
    const SceneGraph& g;
    FlatScene& out;
    std::string& err;
    std::vector<DPrim> bvh_prims;
    std::vector<Box3> bvh_boxes;
    std::vector<DPrim> boundary_prims;
    std::vector<Chain> chains;      // index = xform id - 1

    Flattener(const SceneGraph& g_, FlatScene& o, std::string& e) : g(g_), out(o), err(e) {}

    int fail(int code, const std::string& m) { err = m; return code; }

    static bool same(const Chain& a, const Chain& b) {
        if (a.size() != b.size()) return false;
        for (size_t i = 0; i < a.size(); ++i)
            if (a[i].rot != b[i].rot || a[i].s != b[i].s || a[i].c != b[i].c || a[i].off.x != b[i].off.x ||
                a[i].off.y != b[i].off.y || a[i].off.z != b[i].off.z) return false;
        return true;
    }

    int intern(const Chain& ch) {
        if (ch.empty()) return 0;
        for (size_t i = 0; i < chains.size(); ++i) if (same(chains[i], ch)) return (int)i + 1;
        chains.push_back(ch);
        std::vector<Composed> prefix;
        Composed m = compose(ch, &prefix);
        DXform x; std::memset(&x, 0, sizeof(x));
        x.m_cos = (float)m.C; x.m_sin = (float)m.S; x.bx = (float)m.b[0]; x.by = (float)m.b[1]; x.bz = (float)m.b[2];
        x.n_ops = (int)ch.size();
        x.d_cos = m.C; x.d_sin = m.S; x.d_bx = m.b[0]; x.d_by = m.b[1]; x.d_bz = m.b[2];
        for (size_t j = 0; j < ch.size(); ++j) {
            x.ops[j].op_cos = (float)(ch[j].rot ? ch[j].c : 1.0);
            x.ops[j].op_sin = (float)(ch[j].rot ? ch[j].s : 0.0);
            x.ops[j].cum_cos = (float)prefix[j].C; x.ops[j].cum_sin = (float)prefix[j].S;
            x.is_rot[j] = ch[j].rot ? 1 : 0;
        }
        out.xforms.push_back(x);
        return (int)chains.size();
    }

    // ---- structural identity of hittables ------------------------------------------------------------------------
    // new_bvh_node puts a single-object span into BOTH children as separate clones (src/hittable.rs:96-98): a host that
    // walks such a tree hands every odd-span leaf over twice.  Two coincident primitives would be tested twice and, worse,
    // defeat the exact self-intersection rule (the twin of the primitive a ray starts on is not `skip`).  Members of one
    // BvhNode that are field-by-field equal (what #[derive(Clone)] produces) are therefore emitted once; the closest hit
    // is the same.  A ConstantMedium is never merged: two copies draw twice (src/hittable.rs:446).
    std::vector<uint64_t> hash_memo;
    std::vector<uint8_t> hash_done;
    static uint64_t mix(uint64_t h, uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); return h * 0xBF58476D1CE4E5B9ull; }
    static uint64_t bits(double d) { uint64_t u; std::memcpy(&u, &d, 8); return u; }
    uint64_t hash_of(int id, int depth = 0) {
        if (id < 0 || id >= (int)g.nodes.size() || depth > 64) return 0;
        if (hash_done.empty()) { hash_done.assign(g.nodes.size(), 0); hash_memo.assign(g.nodes.size(), 0); }
        if (hash_done[id]) return hash_memo[id];
        const HNode& h = g.nodes[id];
        uint64_t v = mix(0x1234567, (uint64_t)h.kind * 131 + (uint64_t)h.mat);
        for (double d : {h.c0.x, h.c0.y, h.c0.z, h.c1.x, h.c1.y, h.c1.z, h.radius, h.time0, h.time1, h.a0, h.a1, h.b0, h.b1, h.k,
                         h.bmin.x, h.bmin.y, h.bmin.z, h.bmax.x, h.bmax.y, h.bmax.z, h.offset.x, h.offset.y, h.offset.z,
                         h.sin_theta, h.cos_theta, h.density}) v = mix(v, bits(d));
        if (h.child >= 0) v = mix(v, hash_of(h.child, depth + 1));
        for (int c : h.children) v = mix(v, hash_of(c, depth + 1));
        hash_done[id] = 1; hash_memo[id] = v;
        return v;
    }
    bool same_hittable(int a, int b, int depth = 0) const {
        if (a == b) return true;
        if (depth > 64) return false;
        const HNode& x = g.nodes[a]; const HNode& y = g.nodes[b];
        auto eq = [](double p, double q) { return bits(p) == bits(q); };
        auto eq3 = [&](const V3d& p, const V3d& q) { return eq(p.x, q.x) && eq(p.y, q.y) && eq(p.z, q.z); };
        if (x.kind != y.kind || x.mat != y.mat || !eq3(x.c0, y.c0) || !eq3(x.c1, y.c1) || !eq(x.radius, y.radius) || !eq(x.time0, y.time0) ||
            !eq(x.time1, y.time1) || !eq(x.a0, y.a0) || !eq(x.a1, y.a1) || !eq(x.b0, y.b0) || !eq(x.b1, y.b1) || !eq(x.k, y.k) ||
            !eq3(x.bmin, y.bmin) || !eq3(x.bmax, y.bmax) || !eq3(x.offset, y.offset) || !eq(x.sin_theta, y.sin_theta) ||
            !eq(x.cos_theta, y.cos_theta) || !eq(x.density, y.density) || x.children.size() != y.children.size()) return false;
        if ((x.child >= 0) != (y.child >= 0)) return false;
        if (x.child >= 0 && !same_hittable(x.child, y.child, depth + 1)) return false;
        for (size_t i = 0; i < x.children.size(); ++i) if (!same_hittable(x.children[i], y.children[i], depth + 1)) return false;
        return true;
    }
    bool has_medium(int id, int depth = 0) const {
        if (id < 0 || id >= (int)g.nodes.size() || depth > 64) return false;
        const HNode& h = g.nodes[id];
        if (h.kind == H_MEDIUM) return true;
        if (h.child >= 0 && has_medium(h.child, depth + 1)) return true;
        for (int c : h.children) if (has_medium(c, depth + 1)) return true;
        return false;
    }

    void push_prim(const DPrim& p, const Box3& b, bool boundary) {
        if (boundary) boundary_prims.push_back(p);
        else { bvh_prims.push_back(p); bvh_boxes.push_back(b); }
    }

    int emit_sphere(const HNode& h, const Chain& ch, bool boundary) {
        DPrim p; std::memset(&p, 0, sizeof(p));
        Composed m = compose(ch);
        double c0[3] = {h.c0.x, h.c0.y, h.c0.z}, c0w[3];
        to_world_point(m, c0, c0w);
        p.s.cx = c0w[0]; p.s.cy = c0w[1]; p.s.cz = c0w[2]; p.s.r = h.radius;
        double ar = std::fabs(h.radius);
        Box3 b;
        double lo[3] = {c0w[0] - ar, c0w[1] - ar, c0w[2] - ar}, hi[3] = {c0w[0] + ar, c0w[1] + ar, c0w[2] + ar};
        b.grow(lo); b.grow(hi);
        if (h.kind == H_MOVING_SPHERE) {
            double dc[3] = {h.c1.x - h.c0.x, h.c1.y - h.c0.y, h.c1.z - h.c0.z}, dcw[3];
            to_world_vec(m, dc, dcw);
            p.dcx = (float)dcw[0]; p.dcy = (float)dcw[1]; p.dcz = (float)dcw[2];
            p.t0 = (float)h.time0; p.inv_dt = (float)(1.0 / (h.time1 - h.time0));
            p.type = PRIM_MOVING_SPHERE;
            out.mov_t0 = std::max(out.mov_t0, std::min(h.time0, h.time1)); out.mov_t1 = std::min(out.mov_t1, std::max(h.time0, h.time1));
            // box over the sphere's own [time0, time1] (the MovingSphere arm shadows the arguments, hittable.rs:480-482),
            // spanned by BOTH the f64 motion vector and the f32 one the device moves the centre with
            const double dcf[3] = {(double)p.dcx, (double)p.dcy, (double)p.dcz};
            double lo1[3], hi1[3];
            for (int a = 0; a < 3; ++a) { lo1[a] = lo[a] + std::min(dcw[a], dcf[a]); hi1[a] = hi[a] + std::max(dcw[a], dcf[a]); }
            b.grow(lo1); b.grow(hi1);
        } else p.type = PRIM_SPHERE;
        p.mat = h.mat - 1;
        p.xform = intern(ch);
        push_prim(p, b, boundary);
        return 0;
    }

    int emit_rect(const HNode& h, const Chain& ch, bool boundary) {
        DPrim p; std::memset(&p, 0, sizeof(p));
        p.q.a0 = (float)h.a0; p.q.a1 = (float)h.a1; p.q.b0 = (float)h.b0; p.q.b1 = (float)h.b1; p.q.k = (float)h.k;
        p.type = h.kind == H_XY ? PRIM_XY : (h.kind == H_XZ ? PRIM_XZ : PRIM_YZ);
        p.mat = h.mat - 1;
        p.xform = intern(ch);
        Composed m = compose(ch);
        Box3 b;
        const double pad = 0.0001;    // hittable.rs:486-503
        for (int i = 0; i < 8; ++i) {
            double a = (i & 1) ? h.a1 : h.a0, bb = (i & 2) ? h.b1 : h.b0, kk = (i & 4) ? h.k + pad : h.k - pad;
            double q[3], w[3];
            if (h.kind == H_XY) { q[0] = a; q[1] = bb; q[2] = kk; }
            else if (h.kind == H_XZ) { q[0] = a; q[1] = kk; q[2] = bb; }
            else { q[0] = kk; q[1] = a; q[2] = bb; }
            to_world_point(m, q, w);
            b.grow(w);
        }
        push_prim(p, b, boundary);
        return 0;
    }

    int emit(int id, Chain& ch, bool boundary, int depth) {
        if (id < 0 || id >= (int)g.nodes.size()) return fail(RTW_ERR_INVALID_ARG, "hittable id out of range");
        if (depth > 64) return fail(RTW_ERR_UNSUPPORTED_NESTING, "hittable nesting deeper than 64");
        const HNode& h = g.nodes[id];
        switch (h.kind) {
        case H_SPHERE: case H_MOVING_SPHERE: return emit_sphere(h, ch, boundary);
        case H_XY: case H_XZ: case H_YZ: return emit_rect(h, ch, boundary);
        case H_BOX: {
            // As a surface: ONE BVH leaf (PRIM_BOX: slab test -> entry / exit face) in front of the six rect records, which
            // stay behind the BVH primitives and describe the hit.  As a ConstantMedium boundary: ONE record, the slab test
            // yields both crossings.  When the box is flat (a slab needs an extent), or with RTW_BOX_PRIM=0: the six rects
            // of new_box themselves.
            const char* ebp = getenv("RTW_BOX_PRIM");                      // (read per flatten: the tests compare both forms in one process)
            const bool box_prim = !(ebp && atoi(ebp) == 0);
            const bool solid = h.bmax.x > h.bmin.x && h.bmax.y > h.bmin.y && h.bmax.z > h.bmin.z && h.children.size() == 6;
            if (!box_prim || !solid) {
                for (int c : h.children) { int rc = emit(c, ch, boundary, depth + 1); if (rc) return rc; }
                return 0;
            }
            DPrim p; std::memset(&p, 0, sizeof(p));
            p.b.lox = (float)h.bmin.x; p.b.hix = (float)h.bmax.x; p.b.loy = (float)h.bmin.y; p.b.hiy = (float)h.bmax.y; p.b.loz = (float)h.bmin.z; p.b.hiz = (float)h.bmax.z;
            p.b.first_face = (int)boundary_prims.size();          // relative to the section behind the BVH primitives (flatten adds its offset)
            p.type = PRIM_BOX; p.mat = h.mat - 1; p.xform = intern(ch);
            if (boundary) {                  // a medium only needs the two crossings (hit_constant_medium :417-473): no face records
                p.b.first_face = -(1 << 30);
                boundary_prims.push_back(p);
                return 0;
            }
            Box3 bb;
            const size_t nb0 = bvh_prims.size();
            for (int c : h.children) {                           // the faces: records to the back section, their boxes joined
                if (c < 0 || c >= (int)g.nodes.size()) return fail(RTW_ERR_INVALID_ARG, "hittable id out of range");
                const HNode& f = g.nodes[c];
                if (f.kind != H_XY && f.kind != H_XZ && f.kind != H_YZ) return fail(RTW_ERR_INVALID_ARG, "a Box side is not a rect");
                int rc = emit_rect(f, ch, false); if (rc) return rc;
                boundary_prims.push_back(bvh_prims.back()); bb.grow(bvh_boxes.back().mn); bb.grow(bvh_boxes.back().mx);
                bvh_prims.pop_back(); bvh_boxes.pop_back();
            }
            (void)nb0;
            push_prim(p, bb, false);
            return 0;
        }
        case H_BVH_NODE: {
            std::unordered_map<uint64_t, std::vector<int>> seen;       // structural hash -> members already emitted
            for (int c : h.children) {
                if (c < 0 || c >= (int)g.nodes.size()) return fail(RTW_ERR_INVALID_ARG, "hittable id out of range");
                if (!has_medium(c)) {
                    std::vector<int>& bucket = seen[hash_of(c)];
                    bool dup = false;
                    for (int o : bucket) if (same_hittable(o, c)) { dup = true; break; }
                    if (dup) { ++out.n_dedup; continue; }
                    bucket.push_back(c);
                }
                int rc = emit(c, ch, boundary, depth + 1); if (rc) return rc;
            }
            return 0;
        }
        case H_TRANSLATE: case H_ROTATE_Y: {
            if ((int)ch.size() >= RTW_MAX_CHAIN)
                return fail(RTW_ERR_UNSUPPORTED_NESTING, "more than 4 nested Translate/RotateY wrappers");
            Op op; op.rot = h.kind == H_ROTATE_Y; op.s = h.sin_theta; op.c = h.cos_theta; op.off = h.offset;
            ch.push_back(op);
            int rc = emit(h.child, ch, boundary, depth + 1);
            ch.pop_back();
            return rc;
        }
        default: {   // H_MEDIUM
            if (boundary) return fail(RTW_ERR_UNSUPPORTED_NESTING, "ConstantMedium used as the boundary of a ConstantMedium");
            if (!ch.empty()) return fail(RTW_ERR_UNSUPPORTED_NESTING, "ConstantMedium inside Translate/RotateY");
            DMedium m; m.first = (int)boundary_prims.size();
            int rc = emit(h.child, ch, true, depth + 1);
            if (rc) return rc;
            m.count = (int)boundary_prims.size() - m.first;
            if (m.count == 0) return fail(RTW_ERR_INVALID_ARG, "ConstantMedium with an empty boundary");
            m.neg_inv_density = (float)(-1.0 / h.density);
            m.mat = h.mat - 1;
            out.media.push_back(m);
            return 0;
        }
        }
    }
};

// ---- binned-SAH BVH ----------------------------------------------------------------------------
// Works on a compact, physically partitioned array of float boxes (rounded OUTWARD from the f64 primitive boxes, so
// every union below is exact and conservative): each level streams its range sequentially instead of chasing indices.
inline float down(double v) { float f = (float)v; if ((double)f > v) f = std::nextafterf(f, -INFINITY); return std::nextafterf(f, -INFINITY); }
inline float up(double v) { float f = (float)v; if ((double)f < v) f = std::nextafterf(f, INFINITY); return std::nextafterf(f, INFINITY); }

// ---- fork-join over index ranges (the 1 M - 16 M sphere sweep: every O(n) pass of flatten runs on all cores) -------
static int worker_count() {             // RTW_BUILD_THREADS=1 forces the serial builder (tests compare the two)
    if (const char* e = getenv("RTW_BUILD_THREADS")) return std::max(1, std::min(64, atoi(e)));
    unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min<unsigned>(hw ? hw : 1, 32);
}
// fn(chunk, begin, end) on up to worker_count() threads; returns the number of chunks used (1 = ran inline)
template <class Fn> static int parallel_chunks(size_t n, size_t min_per_chunk, Fn fn) {
    const int t = (int)std::min<size_t>((size_t)worker_count(), std::max<size_t>(1, n / std::max<size_t>(1, min_per_chunk)));
    if (t <= 1) { fn(0, (size_t)0, n); return 1; }
    std::vector<std::thread> th;
    for (int i = 1; i < t; ++i) th.emplace_back([&fn, i, t, n]() { fn(i, n * (size_t)i / t, n * (size_t)(i + 1) / t); });
    fn(0, (size_t)0, n / t);
    for (auto& x : th) x.join();
    return t;
}

struct FBox {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    void grow(const FBox& b) { for (int i = 0; i < 3; ++i) { mn[i] = std::min(mn[i], b.mn[i]); mx[i] = std::max(mx[i], b.mx[i]); } }
    void grow_pt(const float p[3]) { for (int i = 0; i < 3; ++i) { mn[i] = std::min(mn[i], p[i]); mx[i] = std::max(mx[i], p[i]); } }
    double area() const {
        double dx = (double)mx[0] - mn[0], dy = (double)mx[1] - mn[1], dz = (double)mx[2] - mn[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0;
        return 2.0 * (dx * dy + dy * dz + dz * dx);
    }
};
struct Item { FBox b; int id; };
struct BuildNode { FBox box; int left = -1, right = -1; int first = 0, count = 0; };
struct BuildTask { int node, first, count, depth; };

struct Builder {
    std::vector<Item> items, tmp;          // tmp: scratch of the out-of-place parallel partition (top levels only)
    std::vector<BuildNode> nodes;
    int max_depth = 0;
    static constexpr int NBINS = 16;
    static constexpr int PAR_MIN = 1 << 18;               // nodes with at least this many items are processed by all threads
    static int MAX_LEAF;
    static double C_TRAV, C_ISECT;

    explicit Builder(const std::vector<Box3>& b) {
        items.resize(b.size());
        parallel_chunks(b.size(), 1 << 16, [&](int, size_t i0, size_t i1) {
            for (size_t i = i0; i < i1; ++i) {
                for (int a = 0; a < 3; ++a) { items[i].b.mn[a] = down(b[i].mn[a]); items[i].b.mx[a] = up(b[i].mx[a]); }
                items[i].id = (int)i;
            }
        });
    }
    static void centroid(const Item& it, float c[3]) { for (int a = 0; a < 3; ++a) c[a] = 0.5f * (it.b.mn[a] + it.b.mx[a]); }
    struct Bounds { FBox box, cbox; void add(const Item& it) { float c[3]; centroid(it, c); box.grow(it.b); cbox.grow_pt(c); }
                    void merge(const Bounds& o) { box.grow(o.box); cbox.grow(o.cbox); } };
    struct Bins { FBox bb[3][NBINS]; int bc[3][NBINS] = {}; };

    Bounds range_bounds(int first, int count, bool par) {
        Bounds r;
        if (!par) { for (int i = first; i < first + count; ++i) r.add(items[i]); return r; }
        std::vector<Bounds> part(worker_count());
        const int used = parallel_chunks((size_t)count, 1 << 16, [&](int c, size_t i0, size_t i1) { Bounds loc; for (size_t i = i0; i < i1; ++i) loc.add(items[first + i]); part[c] = loc; });   // locals: no false sharing
        for (int c = 0; c < used; ++c) r.merge(part[c]);
        return r;
    }

    // Top-down binned SAH.  `bd` = bounds of the range (handed down by the parent, which measured both halves while it
    // partitioned them: one pass over the items per level instead of two).  With `tasks` set, subtrees of at most
    // `cutoff` prims are not built but recorded, so that worker threads can build them afterwards (disjoint item
    // ranges, private node vectors); the levels above them stream their items with all threads (bins per thread,
    // out-of-place partition by prefix sums).
    int build(std::vector<BuildNode>& out, int first, int count, int depth, int& depth_max, std::vector<BuildTask>* tasks, int cutoff, const Bounds& bd) {
        depth_max = std::max(depth_max, depth);
        int me = (int)out.size();
        out.emplace_back();
        out[me].first = first; out[me].count = count; out[me].box = bd.box;
        if (tasks && depth > 0 && count <= cutoff && count > MAX_LEAF) { tasks->push_back(BuildTask{me, first, count, depth}); return me; }
        if (count <= 1) return me;
        const bool par = tasks != nullptr && count >= PAR_MIN;
        const FBox& box = bd.box; const FBox& cbox = bd.cbox;
        float lo[3], scale[3];
        for (int a = 0; a < 3; ++a) { lo[a] = cbox.mn[a]; float ext = cbox.mx[a] - cbox.mn[a]; scale[a] = ext > 0 ? NBINS / ext : 0.f; }
        auto bin_of = [&](const Item& it, int a) { float c = 0.5f * (it.b.mn[a] + it.b.mx[a]); return std::min(NBINS - 1, std::max(0, (int)((c - lo[a]) * scale[a]))); };
        // one pass fills the bins of all three axes
        Bins bins;
        auto fill = [&](Bins& B, size_t i0, size_t i1) {
            for (size_t i = i0; i < i1; ++i) { const Item& it = items[first + i]; for (int a = 0; a < 3; ++a) { int k = bin_of(it, a); B.bb[a][k].grow(it.b); B.bc[a][k]++; } }
        };
        if (!par) fill(bins, 0, (size_t)count);
        else {
            std::vector<Bins> part(worker_count());
            const int used = parallel_chunks((size_t)count, 1 << 16, [&](int c, size_t i0, size_t i1) { Bins loc; fill(loc, i0, i1); part[c] = loc; });
            for (int c = 0; c < used; ++c) for (int a = 0; a < 3; ++a) for (int k = 0; k < NBINS; ++k) if (part[c].bc[a][k]) { bins.bb[a][k].grow(part[c].bb[a][k]); bins.bc[a][k] += part[c].bc[a][k]; }
        }
        double best_cost = 1e300; int best_axis = -1, best_bin = -1;
        double parent_area = std::max(box.area(), 1e-300);
        for (int a = 0; a < 3; ++a) {
            if (!(scale[a] > 0)) continue;
            double la[NBINS], ra[NBINS]; int lc[NBINS], rc[NBINS];
            FBox acc; int n = 0;
            for (int k = 0; k < NBINS; ++k) { if (bins.bc[a][k]) acc.grow(bins.bb[a][k]); n += bins.bc[a][k]; la[k] = acc.area(); lc[k] = n; }
            acc = FBox(); n = 0;
            for (int k = NBINS - 1; k >= 0; --k) { if (bins.bc[a][k]) acc.grow(bins.bb[a][k]); n += bins.bc[a][k]; ra[k] = acc.area(); rc[k] = n; }
            for (int k = 0; k < NBINS - 1; ++k) {
                if (lc[k] == 0 || rc[k + 1] == 0) continue;
                double cost = C_TRAV + C_ISECT * (la[k] * lc[k] + ra[k + 1] * rc[k + 1]) / parent_area;
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = k; }
            }
        }
        double leaf_cost = C_ISECT * count;
        if (count <= MAX_LEAF && (best_axis < 0 || leaf_cost <= best_cost)) return me;
        int mid = first + count / 2;
        Bounds bl, br;
        bool measured = false;
        if (best_axis >= 0) {
            const int ax = best_axis;
            if (!par) {
                // Hoare partition that measures both halves on the way (each item classified exactly once)
                int i = first, j = first + count - 1;
                for (;;) {
                    while (i <= j && bin_of(items[i], ax) <= best_bin) { bl.add(items[i]); ++i; }
                    while (i <= j && bin_of(items[j], ax) > best_bin) { br.add(items[j]); --j; }
                    if (i >= j) break;
                    std::swap(items[i], items[j]);
                    bl.add(items[i]); br.add(items[j]);
                    ++i; --j;
                }
                mid = i; measured = true;
            } else {
                // out of place: per-chunk counts -> prefix sums -> scatter -> copy back (stable, all threads)
                const int T = worker_count();
                std::vector<int> nl(T, 0), nr(T, 0);
                std::vector<Bounds> pl(T), pr(T);
                const int used = parallel_chunks((size_t)count, 1 << 16, [&](int c, size_t i0, size_t i1) {
                    int a = 0, b = 0; Bounds la, lb;
                    for (size_t i = i0; i < i1; ++i) { const Item& it = items[first + i]; if (bin_of(it, ax) <= best_bin) { ++a; la.add(it); } else { ++b; lb.add(it); } }
                    nl[c] = a; nr[c] = b; pl[c] = la; pr[c] = lb;
                });
                int n_left = 0; for (int c = 0; c < used; ++c) n_left += nl[c];
                std::vector<int> ol(used, 0), orr(used, 0);
                for (int c = 1; c < used; ++c) { ol[c] = ol[c - 1] + nl[c - 1]; orr[c] = orr[c - 1] + nr[c - 1]; }
                if (tmp.size() < items.size()) tmp.resize(items.size());
                parallel_chunks((size_t)count, 1 << 16, [&](int c, size_t i0, size_t i1) {   // same chunking as above (same n, same threshold)
                    int wl = first + ol[c], wr = first + n_left + orr[c];
                    for (size_t i = i0; i < i1; ++i) { const Item& it = items[first + i]; if (bin_of(it, ax) <= best_bin) tmp[wl++] = it; else tmp[wr++] = it; }
                });
                parallel_chunks((size_t)count, 1 << 16, [&](int, size_t i0, size_t i1) { std::memcpy(&items[first + i0], &tmp[first + i0], (i1 - i0) * sizeof(Item)); });
                for (int c = 0; c < used; ++c) { bl.merge(pl[c]); br.merge(pr[c]); }
                mid = first + n_left; measured = true;
            }
            if (mid == first || mid == first + count) { mid = first + count / 2; measured = false; }
        }
        if (!measured) {                   // coincident centroids / degenerate split: halves by index, measured separately
            bl = range_bounds(first, mid - first, par); br = range_bounds(mid, first + count - mid, par);
        }
        int l = build(out, first, mid - first, depth + 1, depth_max, tasks, cutoff, bl);
        int r = build(out, mid, first + count - mid, depth + 1, depth_max, tasks, cutoff, br);
        out[me].left = l; out[me].right = r;
        return me;
    }

    void build_all(int n) {
        nodes.reserve(2 * (size_t)n);
        const int n_threads = worker_count();
        const Bounds root = range_bounds(0, n, n >= PAR_MIN && n_threads > 1);
        if (n < (1 << 16) || n_threads < 2) { build(nodes, 0, n, 0, max_depth, nullptr, 0, root); return; }
        std::vector<BuildTask> tasks;
        const bool timing = getenv("RTW_TIMING") != nullptr;
        auto tp0 = std::chrono::steady_clock::now();
        build(nodes, 0, n, 0, max_depth, &tasks, std::max(1024, n / (8 * n_threads)), root);
        if (timing) fprintf(stderr, "[build] top levels (%zu tasks, %d threads) %.3f s\n", tasks.size(), n_threads, std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count());
        std::vector<Item>().swap(tmp);
        std::vector<std::vector<BuildNode>> sub(tasks.size());
        std::vector<int> sub_depth(tasks.size(), 0);
        std::atomic<size_t> next(0);
        auto worker = [&]() {
            for (;;) {
                size_t t = next.fetch_add(1);
                if (t >= tasks.size()) break;
                sub[t].reserve(2 * (size_t)tasks[t].count);
                // the placeholder kept the range's box; the centroid bounds are re-measured here (one pass per task)
                Bounds m; for (int i = tasks[t].first; i < tasks[t].first + tasks[t].count; ++i) m.add(items[i]);
                build(sub[t], tasks[t].first, tasks[t].count, tasks[t].depth, sub_depth[t], nullptr, 0, m);
            }
        };
        std::vector<std::thread> th;
        for (int i = 1; i < n_threads; ++i) th.emplace_back(worker);
        worker();
        for (auto& t : th) t.join();
        if (timing) fprintf(stderr, "[build] subtrees %.3f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count());
        // splice: every task's local root replaces its placeholder, its other nodes go to a precomputed slice of `nodes`
        std::vector<int> off(tasks.size());
        size_t total = nodes.size();
        for (size_t t = 0; t < tasks.size(); ++t) { max_depth = std::max(max_depth, sub_depth[t]); off[t] = (int)total - 1; total += sub[t].size() - 1; }   // local index c >= 1 -> off + c
        nodes.resize(total);
        next = 0;
        auto splice = [&]() {
            for (;;) {
                size_t t = next.fetch_add(1);
                if (t >= tasks.size()) break;
                const int o = off[t];
                auto remap = [o](BuildNode b) { if (b.left >= 0) { b.left += o; b.right += o; } return b; };
                nodes[tasks[t].node] = remap(sub[t][0]);
                for (size_t c = 1; c < sub[t].size(); ++c) nodes[(size_t)o + c] = remap(sub[t][c]);
                std::vector<BuildNode>().swap(sub[t]);
            }
        };
        th.clear();
        for (int i = 1; i < n_threads; ++i) th.emplace_back(splice);
        splice();
        for (auto& t : th) t.join();
    }
};

// One primitive per leaf: in a warp only ~5 lanes reach a leaf together, so every extra primitive of a leaf is tested at
// 5 of 32 lanes (measured: leaves of up to 4 cost cornell_box 11 % — a box kept its 6 faces in two leaves — and
// final_scene 1.6 %; C1 and the sphere sweep already split down to single spheres).
int Builder::MAX_LEAF = 1;
double Builder::C_TRAV = 1.0;
double Builder::C_ISECT = 1.5;

void set_child_box(DNode& n, int which, const FBox* b) {
    const float inf = std::numeric_limits<float>::infinity();
    float mnx = inf, mxx = inf, mny = inf, mxy = inf, mnz = inf, mxz = inf;      // unused slot: never hit (see DNode)
    if (b) { mnx = b->mn[0]; mxx = b->mx[0]; mny = b->mn[1]; mxy = b->mx[1]; mnz = b->mn[2]; mxz = b->mx[2]; }
    if (which == 0) { n.c0minx = mnx; n.c0maxx = mxx; n.c0miny = mny; n.c0maxy = mxy; n.c0minz = mnz; n.c0maxz = mxz; }
    else { n.c1minx = mnx; n.c1maxx = mxx; n.c1miny = mny; n.c1maxy = mxy; n.c1minz = mnz; n.c1maxz = mxz; }
}
inline int leaf_code(int first, int count) { return ~((first << 3) | (count - 1)); }

void pack_materials(const SceneGraph& g, FlatScene& out) {
    for (const HTexture& t : g.textures) {
        DTex d; std::memset(&d, 0, sizeof(d));
        d.kind = t.kind; d.r0 = (float)t.c0[0]; d.g0 = (float)t.c0[1]; d.b0 = (float)t.c0[2];
        d.r1 = (float)t.c1[0]; d.g1 = (float)t.c1[1]; d.b1 = (float)t.c1[2]; d.scale = (float)t.scale;
        if (t.kind == TEX_NOISE) {
            d.a = (int)(out.perlin.size() / RTW_PERLIN_BYTES);
            size_t base = out.perlin.size();
            out.perlin.resize(base + RTW_PERLIN_BYTES);
            float* rv = reinterpret_cast<float*>(out.perlin.data() + base);
            for (int i = 0; i < 256; ++i) {
                rv[4 * i] = (float)t.ranvec[3 * i]; rv[4 * i + 1] = (float)t.ranvec[3 * i + 1];
                rv[4 * i + 2] = (float)t.ranvec[3 * i + 2]; rv[4 * i + 3] = 0.f;
            }
            uint8_t* pm = out.perlin.data() + base + 256 * 16;
            for (int i = 0; i < 768; ++i) pm[i] = (uint8_t)(t.perm[i] & 255);
        } else if (t.kind == TEX_IMAGE) {
            size_t base = (out.image.size() + 15) & ~(size_t)15;
            out.image.resize(base + t.data.size());
            std::memcpy(out.image.data() + base, t.data.data(), t.data.size());
            d.a = (int)base; d.w = t.w; d.h = t.h; d.bps = t.bps;
        }
        out.texs.push_back(d);
    }
    for (const HMaterial& m : g.materials) {
        DMat d; std::memset(&d, 0, sizeof(d));
        d.kind = m.kind; d.tex = -1;
        if (m.kind == MAT_METAL) { d.r = (float)m.albedo[0]; d.g = (float)m.albedo[1]; d.b = (float)m.albedo[2]; d.param = (float)m.fuzz; }
        else if (m.kind == MAT_DIELECTRIC) { d.r = d.g = d.b = 1.f; d.param = (float)m.ir; }
        else {
            const HTexture& t = g.textures[m.tex];
            if (t.kind == TEX_SOLID) { d.r = (float)t.c0[0]; d.g = (float)t.c0[1]; d.b = (float)t.c0[2]; }
            else d.tex = m.tex;
        }
        out.mats.push_back(d);
    }
}

}  // namespace

// ---- BVH2 -> BVH8 on the host (small scenes, CPU tests); the device builder runs the same collapse_one per node ----
namespace {
struct HostAlloc {
    int n_nodes = 1, n_prims = 0, n_queue = 0, max_depth = 0;     // node 0 = root, allocated up front
    int nodes(int k) { int b = n_nodes; n_nodes += k; return b; }
    int prims(int k) { int b = n_prims; n_prims += k; return b; }
    int queue(int k) { int b = n_queue; n_queue += k; return b; }
    void depth(int d) { if (d > max_depth) max_depth = d; }
};
}  // namespace

int collapse_to_wide(const rtww::B2View& v, std::vector<DWNode>& wnodes, std::vector<int>& leaf_order) {
    leaf_order.assign((size_t)v.n_leaves, -1);
    wnodes.clear();
    if (v.n_leaves == 0) return 0;
    if (v.n_inner == 0) {                   // a single primitive: a root with one leaf slot
        DWNode w; std::memset(&w, 0, sizeof(w));
        const float* b = v.box;             // leaf 0 = ref 0
        float p[3], step[3];
        w.ex = (uint8_t)rtww::wide_axis_grid(b[0], b[3], p[0], step[0]);
        w.ey = (uint8_t)rtww::wide_axis_grid(b[1], b[4], p[1], step[1]);
        w.ez = (uint8_t)rtww::wide_axis_grid(b[2], b[5], p[2], step[2]);
        w.px = p[0]; w.py = p[1]; w.pz = p[2];
        for (int s = 0; s < 8; ++s) { w.qlox[s] = w.qloy[s] = w.qloz[s] = 255; w.qhix[s] = w.qhiy[s] = w.qhiz[s] = 0; }
        w.qlox[0] = rtww::wide_qlo(b[0], p[0], step[0]); w.qhix[0] = rtww::wide_qhi(b[3], p[0], step[0]);
        w.qloy[0] = rtww::wide_qlo(b[1], p[1], step[1]); w.qhiy[0] = rtww::wide_qhi(b[4], p[1], step[1]);
        w.qloz[0] = rtww::wide_qlo(b[2], p[2], step[2]); w.qhiz[0] = rtww::wide_qhi(b[5], p[2], step[2]);
        w.lmask = 1;
        wnodes.push_back(w); leaf_order[0] = 0;
        return 1;
    }
    wnodes.resize((size_t)v.n_inner);       // upper bound: every wide node consumes at least one binary inner node
    HostAlloc alloc;
    std::vector<rtww::WideItem> cur{rtww::WideItem{0, 0, 0}}, next;
    while (!cur.empty()) {
        next.assign(cur.size() * 8, rtww::WideItem{0, 0, 0});
        alloc.n_queue = 0;
        for (const rtww::WideItem& it : cur) rtww::collapse_one(v, it, wnodes.data(), leaf_order.data(), next.data(), alloc);
        next.resize((size_t)alloc.n_queue);
        cur.swap(next);
    }
    wnodes.resize((size_t)alloc.n_nodes);
    return alloc.max_depth;
}

int kWideMinPrims = 1 << 18;             // 8-wide nodes from 256 Ki primitives (the wavefront pipeline's range); RTW_BIG_MIN / RTW_WIDE_MIN / RTW_BVH override
int choose_bvh_width(long long n, int requested) {
    int width = requested;
    if (width == 0) { if (const char* e = getenv("RTW_BVH")) width = atoi(e); }
    if (width != 2 && width != 8) {
        long long wide_min = kWideMinPrims;
        if (const char* e = getenv("RTW_BIG_MIN")) wide_min = std::max(1ll, atoll(e));
        if (const char* e = getenv("RTW_WIDE_MIN")) wide_min = std::max(1, atoi(e));
        width = n >= wide_min ? 8 : 2;
    }
    return width;
}

int flatten(const SceneGraph& g, const std::vector<int>& roots, FlatScene& out, std::string& err, const FlattenOptions& opt) {
    const bool timing = getenv("RTW_TIMING") != nullptr;
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (timing) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "[flatten] %s %.3f s\n", what, std::chrono::duration<double>(t - T0).count()); T0 = t; } };
    out = FlatScene();
    DXform ident; std::memset(&ident, 0, sizeof(ident)); ident.m_cos = 1.f; ident.d_cos = 1.0;
    out.xforms.push_back(ident);
    Flattener fl(g, out, err);
    Chain ch;
    for (int id : roots) {
        int rc = fl.emit(id, ch, false, 0);
        if (rc) return rc;
    }
    const bool with_bulk = &roots == &g.world;
    if (opt.emit_only) {
        // input of the device builder: scene-graph primitives + boxes as they are, bulk spheres untouched
        pack_materials(g, out);
        for (const DPrim& p : fl.bvh_prims) if (p.mat < 0 || p.mat >= (int)out.mats.size()) { err = "material handle out of range"; return RTW_ERR_INVALID_ARG; }
        for (const DMedium& m : out.media) if (m.mat < 0 || m.mat >= (int)out.mats.size()) { err = "phase material handle out of range"; return RTW_ERR_INVALID_ARG; }
        if (with_bulk) {                                   // (16 M spheres: 640 MB to read — 58 ms on one thread, most of this flatten)
            std::atomic<int> bad{0};
            const int n_mats = (int)out.mats.size();
            parallel_chunks(g.bulk.size(), 1 << 16, [&](int, size_t i0, size_t i1) {
                int b = 0;
                for (size_t i = i0; i < i1; ++i) b |= (g.bulk[i].mat < 1 || g.bulk[i].mat > n_mats) ? 1 : 0;
                if (b) bad.store(1, std::memory_order_relaxed);
            });
            if (bad.load()) { err = "material handle out of range"; return RTW_ERR_INVALID_ARG; }
        }
        const long long n_all = (long long)fl.bvh_prims.size() + (with_bulk ? (long long)g.bulk.size() : 0);
        if (n_all >= (1 << 28)) { err = "too many primitives"; return RTW_ERR_INVALID_ARG; }
        out.emit_only = true;
        out.wide = choose_bvh_width(n_all, opt.bvh_width) == 8;
        out.n_bvh_prims = (int)n_all;
        out.prim_boxes.resize(fl.bvh_prims.size() * 6);
        for (size_t i = 0; i < fl.bvh_prims.size(); ++i) for (int a = 0; a < 3; ++a) { out.prim_boxes[i * 6 + a] = down(fl.bvh_boxes[i].mn[a]); out.prim_boxes[i * 6 + 3 + a] = up(fl.bvh_boxes[i].mx[a]); }
        for (const DPrim& p : fl.bvh_prims) { if (p.type >= PRIM_XY) out.features |= 1; if (p.xform) out.features |= p.type >= PRIM_XY ? 2 | 32 : 2; }
        for (const DPrim& p : fl.boundary_prims) { if (p.type >= PRIM_XY) out.features |= 1; if (p.xform) out.features |= p.type >= PRIM_XY ? 2 | 32 : 2; }
        if (!out.media.empty()) out.features |= 4;
        for (const DTex& t : out.texs) { if (t.kind == TEX_NOISE) out.features |= 8; if (t.kind == TEX_IMAGE) out.features |= 16; }
        for (DMedium& m : out.media) m.first += (int)n_all;
        for (DPrim& p : fl.bvh_prims) if (p.type == PRIM_BOX) p.b.first_face += (int)n_all;       // (boundary boxes have no faces)
        out.prims.swap(fl.bvh_prims);
        out.boundary.swap(fl.boundary_prims);
        return 0;
    }
    if (with_bulk) {           // bulk spheres belong to the world, not to a single-hittable view
        const size_t base = fl.bvh_prims.size();
        fl.bvh_prims.resize(base + g.bulk.size());
        fl.bvh_boxes.resize(base + g.bulk.size());
        parallel_chunks(g.bulk.size(), 1 << 16, [&](int, size_t i0, size_t i1) {
            for (size_t i = i0; i < i1; ++i) {
                const BulkSphere& b = g.bulk[i];
                DPrim p; std::memset(&p, 0, sizeof(p));
                p.s.cx = b.c[0]; p.s.cy = b.c[1]; p.s.cz = b.c[2]; p.s.r = b.r; p.type = PRIM_SPHERE; p.mat = b.mat - 1;
                double ar = std::fabs(b.r);
                Box3 bx; double lo[3] = {b.c[0] - ar, b.c[1] - ar, b.c[2] - ar}, hi[3] = {b.c[0] + ar, b.c[1] + ar, b.c[2] + ar};
                bx.grow(lo); bx.grow(hi);
                fl.bvh_prims[base + i] = p; fl.bvh_boxes[base + i] = bx;
            }
        });
    }
    lap("emit");
    pack_materials(g, out);
    for (const DPrim& p : fl.bvh_prims) if (p.mat < 0 || p.mat >= (int)out.mats.size()) { err = "material handle out of range"; return RTW_ERR_INVALID_ARG; }
    for (const DMedium& m : out.media) if (m.mat < 0 || m.mat >= (int)out.mats.size()) { err = "phase material handle out of range"; return RTW_ERR_INVALID_ARG; }
    const int n = (int)fl.bvh_prims.size();
    if (n >= (1 << 28)) { err = "too many primitives"; return RTW_ERR_INVALID_ARG; }
    // node format: binary or 8-wide compressed (FlattenOptions / RTW_BVH / RTW_WIDE_MIN)
    out.wide = choose_bvh_width(n, opt.bvh_width) == 8;
    // BVH (RTW_BVH_LEAF / RTW_BVH_CI: tuning overrides for kernel experiments)
    Builder::MAX_LEAF = 1;
    if (const char* e = getenv("RTW_BVH_LEAF")) Builder::MAX_LEAF = std::max(1, std::min(8, atoi(e)));
    if (out.wide) Builder::MAX_LEAF = 1;              // a wide leaf slot holds exactly one primitive
    if (const char* e = getenv("RTW_BVH_CI")) Builder::C_ISECT = atof(e);
    Builder b(fl.bvh_boxes);
    DNode root; std::memset(&root, 0, sizeof(root));
    if (n == 0) {
        set_child_box(root, 0, nullptr); set_child_box(root, 1, nullptr);
        root.child0 = root.child1 = leaf_code(0, 1);
        out.nodes.push_back(root);
    } else {
        lap("builder init");
        b.build_all(n);
        lap("build");
        // leaf order
        out.prims.resize(n);
        parallel_chunks((size_t)n, 1 << 16, [&](int, size_t i0, size_t i1) { for (size_t i = i0; i < i1; ++i) out.prims[i] = fl.bvh_prims[b.items[i].id]; });
        if (b.nodes[0].left < 0) {     // the whole scene is one leaf
            // one leaf: both slots point at it (tested twice, like the reference's duplicated single-object
            // leaves, src/hittable.rs:96-98); an "empty" slot cannot be encoded with min/max slabs
            set_child_box(root, 0, &b.nodes[0].box); set_child_box(root, 1, &b.nodes[0].box);
            root.child0 = leaf_code(0, n); root.child1 = leaf_code(0, n);
            out.nodes.push_back(root);
        } else {
            // inner build nodes -> DNode indices (pre-order)
            const size_t nb = b.nodes.size();
            std::vector<int> dn(nb, -1);
            std::vector<int> part_cnt(worker_count() + 1, 0);
            const int used = parallel_chunks(nb, 1 << 16, [&](int c, size_t i0, size_t i1) { int k = 0; for (size_t i = i0; i < i1; ++i) k += b.nodes[i].left >= 0; part_cnt[c + 1] = k; });
            for (int c = 0; c < used; ++c) part_cnt[c + 1] += part_cnt[c];
            const int cnt = part_cnt[used];
            parallel_chunks(nb, 1 << 16, [&](int c, size_t i0, size_t i1) { int k = part_cnt[c]; for (size_t i = i0; i < i1; ++i) if (b.nodes[i].left >= 0) dn[i] = k++; });
            out.nodes.resize(cnt);
            const double root_area = std::max(b.nodes[0].box.area(), 1e-300);
            std::vector<double> part_sah(worker_count(), 0.0);
            parallel_chunks(nb, 1 << 16, [&](int c, size_t i0, size_t i1) {
                double sah = 0;
                for (size_t i = i0; i < i1; ++i) {
                    const BuildNode& bn = b.nodes[i];
                    if (bn.left < 0) { sah += Builder::C_ISECT * bn.count * bn.box.area() / root_area; continue; }
                    sah += Builder::C_TRAV * bn.box.area() / root_area;
                    DNode d; std::memset(&d, 0, sizeof(d));
                    const BuildNode& l = b.nodes[bn.left]; const BuildNode& r = b.nodes[bn.right];
                    set_child_box(d, 0, &l.box); set_child_box(d, 1, &r.box);
                    d.child0 = l.left >= 0 ? dn[bn.left] : leaf_code(l.first, l.count);
                    d.child1 = r.left >= 0 ? dn[bn.right] : leaf_code(r.first, r.count);
                    out.nodes[dn[i]] = d;
                }
                part_sah[c] = sah;
            });
            double sah = 0; for (double v : part_sah) sah += v;
            out.sah_cost = sah;
        }
    }
    lap("emit nodes");
    out.max_depth = b.max_depth;
    out.n_bvh_prims = n;
    if (opt.keep_boxes && n > 0) {
        out.prim_boxes.resize((size_t)n * 6);
        for (int i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) { out.prim_boxes[(size_t)i * 6 + a] = b.items[i].b.mn[a]; out.prim_boxes[(size_t)i * 6 + 3 + a] = b.items[i].b.mx[a]; }
    }
    if (out.wide && n > 0) {
        // the binary tree as a B2View: inner nodes keep their DNode index, leaf j = position j of the leaf order
        const size_t nb = b.nodes.size();
        std::vector<int> inner_of(nb, -1);
        int n_inner = 0;
        for (size_t i = 0; i < nb; ++i) if (b.nodes[i].left >= 0) inner_of[i] = n_inner++;
        std::vector<float> box((size_t)(n_inner + n) * 6);
        std::vector<int> left((size_t)std::max(n_inner, 1)), right((size_t)std::max(n_inner, 1)), count((size_t)std::max(n_inner, 1));
        auto put_box = [&](size_t ref, const FBox& fb) { for (int a = 0; a < 3; ++a) { box[ref * 6 + a] = fb.mn[a]; box[ref * 6 + 3 + a] = fb.mx[a]; } };
        auto ref_of = [&](int bi) { const BuildNode& c = b.nodes[bi]; return c.left >= 0 ? inner_of[bi] : n_inner + c.first; };
        parallel_chunks(nb, 1 << 16, [&](int, size_t i0, size_t i1) {
            for (size_t i = i0; i < i1; ++i) {
                const BuildNode& bn = b.nodes[i];
                if (bn.left >= 0) { put_box((size_t)inner_of[i], bn.box); left[inner_of[i]] = ref_of(bn.left); right[inner_of[i]] = ref_of(bn.right); count[inner_of[i]] = bn.count; }
                else put_box((size_t)(n_inner + bn.first), bn.box);
            }
        });
        const char* efill = getenv("RTW_WIDE_FILL");                              // A/B switch, read per flatten
        const bool fill = !(efill && atoi(efill) == 0);
        rtww::B2View view{box.data(), left.data(), right.data(), n_inner, n, fill ? count.data() : nullptr};
        std::vector<int> order;
        out.wide_depth = collapse_to_wide(view, out.wnodes, order);
        std::vector<DPrim> reordered((size_t)n);
        parallel_chunks((size_t)n, 1 << 16, [&](int, size_t i0, size_t i1) { for (size_t i = i0; i < i1; ++i) reordered[i] = out.prims[order[i]]; });
        out.prims.swap(reordered);
        if (!out.prim_boxes.empty()) {
            std::vector<float> pb((size_t)n * 6);
            for (int i = 0; i < n; ++i) std::memcpy(&pb[(size_t)i * 6], &out.prim_boxes[(size_t)order[i] * 6], 24);
            out.prim_boxes.swap(pb);
        }
        std::vector<DNode>().swap(out.nodes);
        lap("collapse to 8-wide");
        if (out.wide_depth > RTW_WIDE_STACK) { err = "wide BVH deeper than the traversal stack"; return RTW_ERR_UNSUPPORTED_NESTING; }
    }
    // feature bits (must match the FEAT_* enum of rtw_device.cuh): RECT 1, XFORM 2, MEDIA 4, NOISE 8, IMAGE 16, RXFORM 32 (an instanced RECT)
    for (const DPrim& p : fl.bvh_prims) { if (p.type >= PRIM_XY) out.features |= 1; if (p.xform) out.features |= p.type >= PRIM_XY ? 2 | 32 : 2; }
    for (const DPrim& p : fl.boundary_prims) { if (p.type >= PRIM_XY) out.features |= 1; if (p.xform) out.features |= p.type >= PRIM_XY ? 2 | 32 : 2; }
    if (!out.media.empty()) out.features |= 4;
    for (const DTex& t : out.texs) { if (t.kind == TEX_NOISE) out.features |= 8; if (t.kind == TEX_IMAGE) out.features |= 16; }
    // boundary prims (and the face rects of boxes) follow; shift media ranges and the boxes' face indices
    for (DMedium& m : out.media) m.first += n;
    for (int i = 0; i < n; ++i) if (out.prims[i].type == PRIM_BOX) out.prims[i].b.first_face += n;
    out.prims.insert(out.prims.end(), fl.boundary_prims.begin(), fl.boundary_prims.end());
    if (out.max_depth > 60) { err = "BVH deeper than the traversal stack"; return RTW_ERR_UNSUPPORTED_NESTING; }
    return 0;
}

bool validate_bvh(const FlatScene& f, std::string& err) {
    std::vector<int> seen(f.n_bvh_prims, 0);
    struct Item { int node; };
    if (f.nodes.empty()) { err = "no nodes"; return false; }
    std::vector<int> stack{0};
    size_t visited = 0;
    while (!stack.empty()) {
        int ni = stack.back(); stack.pop_back();
        if (ni < 0 || ni >= (int)f.nodes.size()) { err = "node index out of range"; return false; }
        if (++visited > f.nodes.size()) { err = "cycle"; return false; }
        const DNode& n = f.nodes[ni];
        for (int c = 0; c < 2; ++c) {
            int ch = c ? n.child1 : n.child0;
            float mnx = c ? n.c1minx : n.c0minx, mxx = c ? n.c1maxx : n.c0maxx;
            if (ch >= 0) { stack.push_back(ch); continue; }
            if (mnx > mxx || std::isinf(mnx)) continue;   // unused slot (empty world)
            int code = ~ch, first = code >> 3, count = (code & 7) + 1;
            if (first < 0 || first + count > f.n_bvh_prims) { err = "leaf range out of bounds"; return false; }
            for (int i = first; i < first + count; ++i) seen[i]++;
        }
    }
    const int want = (f.nodes.size() == 1 && f.nodes[0].child0 == f.nodes[0].child1 && f.nodes[0].child0 < 0) ? 2 : 1;
    for (int i = 0; i < f.n_bvh_prims; ++i) if (seen[i] != want) { err = "prim not referenced exactly once"; return false; }
    return true;
}

bool validate_wide(const FlatScene& f, std::string& err) {
    const int n = f.n_bvh_prims;
    if (n == 0) return true;
    if (f.wnodes.empty()) { err = "no wide nodes"; return false; }
    std::vector<int> seen((size_t)n, 0);
    std::vector<uint8_t> node_seen(f.wnodes.size(), 0);
    const bool boxes = f.prim_boxes.size() == (size_t)n * 6;
    struct Frame { int node; float box[6]; bool has_box; };
    std::vector<Frame> stack;
    Frame root; root.node = 0; root.has_box = false;
    stack.push_back(root);
    while (!stack.empty()) {
        Frame fr = stack.back(); stack.pop_back();
        if (fr.node < 0 || fr.node >= (int)f.wnodes.size()) { err = "wide node index out of range"; return false; }
        if (node_seen[fr.node]++) { err = "wide node referenced twice"; return false; }
        const DWNode& w = f.wnodes[fr.node];
        if (w.imask & w.lmask) { err = "slot is both inner and leaf"; return false; }
        if (!(w.imask | w.lmask)) { err = "empty wide node"; return false; }
        // union of the child boxes must stay inside what the parent promised for this node
        for (int s = 0; s < 8; ++s) {
            const bool in = (w.imask >> s) & 1, lf = (w.lmask >> s) & 1;
            if (!in && !lf) continue;
            float cb[6]; rtww::wide_child_box(w, s, cb);
            if (fr.has_box) for (int a = 0; a < 3; ++a) {
                // a child box may stick out of the parent's quantised box by its own grid padding only
                const float tol = 4.0f * std::ldexp(1.0f, (int)(a == 0 ? w.ex : a == 1 ? w.ey : w.ez) - 127) + 1e-6f * (std::fabs(cb[a]) + std::fabs(cb[a + 3]));
                if (cb[a] < fr.box[a] - tol || cb[a + 3] > fr.box[a + 3] + tol) { err = "child box outside its parent's box"; return false; }
            }
            const int below = (1 << s) - 1;
            if (in) {
                Frame c; c.node = (int)w.child_base + __builtin_popcount(w.imask & below); c.has_box = true; std::memcpy(c.box, cb, sizeof(cb));
                stack.push_back(c);
            } else {
                const int pi = (int)w.prim_base + __builtin_popcount(w.lmask & below);
                if (pi < 0 || pi >= n) { err = "leaf primitive index out of range"; return false; }
                seen[pi]++;
                if (boxes) for (int a = 0; a < 3; ++a)
                    if (f.prim_boxes[(size_t)pi * 6 + a] < cb[a] || f.prim_boxes[(size_t)pi * 6 + 3 + a] > cb[a + 3]) { err = "leaf box does not contain its primitive"; return false; }
            }
        }
    }
    for (int i = 0; i < n; ++i) if (seen[i] != 1) { err = "prim not referenced exactly once by the wide BVH"; return false; }
    for (size_t i = 0; i < f.wnodes.size(); ++i) if (!node_seen[i]) { err = "unreachable wide node"; return false; }
    return true;
}

bool check_wide_traversal(const FlatScene& f, int n_rays, uint64_t seed, uint64_t out[5], std::string& err) {
    for (int i = 0; i < 5; ++i) out[i] = 0;
    const int n = f.n_bvh_prims;
    if (n == 0 || f.wnodes.empty()) return true;
    if (f.prim_boxes.size() != (size_t)n * 6) { err = "prim boxes were not kept"; return false; }
    // scene bounds
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], (double)f.prim_boxes[(size_t)i * 6 + a]); hi[a] = std::max(hi[a], (double)f.prim_boxes[(size_t)i * 6 + 3 + a]); }
    // a huge primitive (the r = 1000 ground sphere) would put every origin far away from everything else: aim at the
    // median-sized part of the scene as well
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 12345;
    auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (double)((st >> 11) & ((1ull << 53) - 1)) / (double)(1ull << 53); };
    const float t_min = 0.001f;
    std::vector<uint8_t> reached((size_t)n);
    for (int ri = 0; ri < n_rays; ++ri) {
        // origin: around a random primitive (near or far), direction: towards another random primitive, or axis-parallel
        const int pa = (int)(rnd() * n) % n, pb = (int)(rnd() * n) % n;
        float o[3], d[3];
        const double spread = rnd() < 0.5 ? 2.0 : 50.0;
        for (int a = 0; a < 3; ++a) {
            const double ca = 0.5 * (f.prim_boxes[(size_t)pa * 6 + a] + f.prim_boxes[(size_t)pa * 6 + 3 + a]);
            const double ea = std::min(1e3, 0.5 * (f.prim_boxes[(size_t)pa * 6 + 3 + a] - f.prim_boxes[(size_t)pa * 6 + a]) + 1e-3);
            o[a] = (float)(ca + (rnd() * 2 - 1) * spread * ea);
            const double cb = f.prim_boxes[(size_t)pb * 6 + a] + rnd() * (f.prim_boxes[(size_t)pb * 6 + 3 + a] - f.prim_boxes[(size_t)pb * 6 + a]);
            d[a] = (float)((cb - o[a]) * (0.05 + rnd()));
        }
        const int mode = ri % 8;
        if (mode == 5) { d[0] = 0.0f; }                                   // axis-parallel cases (zero components, both signs of zero)
        if (mode == 6) { d[1] = -0.0f; d[2] = 0.0f; }
        if (mode == 7) { d[0] = 0.0f; d[2] = -0.0f; }
        if (d[0] == 0.0f && d[1] == 0.0f && d[2] == 0.0f) d[1] = 1.0f;
        // per-ray constants exactly as rtw_device.cuh slab_setup / bvh8_closest derive them (1/d: IEEE here, MUFU there)
        rtww::WRay r;
        auto sdir = [](float v) { return std::fabs(v) < 1e-20f ? std::copysign(1e-20f, v) : v; };
        r.ix = 1.0f / sdir(d[0]); r.iy = 1.0f / sdir(d[1]); r.iz = 1.0f / sdir(d[2]);
        r.oix = o[0] * r.ix; r.oiy = o[1] * r.iy; r.oiz = o[2] * r.iz;
        r.sx = 2.384185791015625e-07f * std::fabs(r.oix); r.sy = 2.384185791015625e-07f * std::fabs(r.oiy); r.sz = 2.384185791015625e-07f * std::fabs(r.oiz);
        const uint32_t oct = (r.ix < 0 ? 1u : 0u) | (r.iy < 0 ? 2u : 0u) | (r.iz < 0 ? 4u : 0u);
        r.k = oct ^ 7u; r.one = 0x3F800000u;
        std::fill(reached.begin(), reached.end(), 0);
        // traversal with t_best = inf (every leaf whose box passes is reached), the device's group / stack scheme
        struct Grp { uint32_t base, g; };
        std::vector<Grp> stack;
        uint32_t base = 0, grp = (1u << 8) | (1u << (0u ^ r.k));
        for (;;) {
            if (!(grp & 0xffu)) { if (stack.empty()) break; base = stack.back().base; grp = stack.back().g; stack.pop_back(); continue; }
            const int j = 31 - __builtin_clz(grp & 0xffu);
            grp ^= 1u << j;
            const uint32_t slot = (uint32_t)j ^ r.k, imask = grp >> 8;
            const uint32_t node = base + (uint32_t)__builtin_popcount(imask & ((1u << slot) - 1u));
            if (grp & 0xffu) stack.push_back(Grp{base, grp});
            if (stack.size() > RTW_WIDE_STACK) { err = "traversal stack overflow"; return false; }
            const DWNode& w = f.wnodes[node];
            rtww::W4 q[5]; std::memcpy(q, &w, 80);
            uint32_t hits = rtww::wide_node_hits(q[0], q[2], q[3], q[4], r, t_min, INFINITY);
            hits &= (uint32_t)(w.imask | w.lmask);
            out[1]++;
            const uint32_t m16 = rtww::wide_perm16((hits & w.imask) | ((hits & w.lmask) << 8), r.k);
            uint32_t pl = m16 >> 8;
            while (pl) {
                const int jj = 31 - __builtin_clz(pl); pl ^= 1u << jj;
                const uint32_t s = (uint32_t)jj ^ r.k;
                reached[w.prim_base + (uint32_t)__builtin_popcount(w.lmask & ((1u << s) - 1u))] = 1;
                out[2]++;
            }
            base = w.child_base; grp = ((uint32_t)w.imask << 8) | (m16 & 0xffu);
        }
        // exact slab test in double against every primitive box
        for (int i = 0; i < n; ++i) {
            double tn = t_min, tf = 1e300; bool miss = false;
            for (int a = 0; a < 3 && !miss; ++a) {
                const double mn = f.prim_boxes[(size_t)i * 6 + a], mx = f.prim_boxes[(size_t)i * 6 + 3 + a], oo = o[a], dd = d[a];
                if (dd == 0.0) {
                    // origin within rounding distance of a face and no motion along this axis: "inside" is undecidable in
                    // f32 (o / d is rounded at 1e20 |o|; the reference itself computes 0 * inf = NaN there) — not counted
                    const double eps = 1e-6 * (std::fabs(oo) + 1.0);
                    if (oo < mn + eps || oo > mx - eps) miss = true;
                    continue;
                }
                double t0 = (mn - oo) / dd, t1 = (mx - oo) / dd;
                if (t0 > t1) std::swap(t0, t1);
                tn = std::max(tn, t0); tf = std::min(tf, t1);
                if (tn > tf) miss = true;
            }
            if (!miss) {
                out[3]++;
                if (!reached[i]) {
                    out[4]++;
                    if (getenv("RTW_DEBUG_WIDE")) fprintf(stderr, "[wide] ray %d o (%.9g %.9g %.9g) d (%.9g %.9g %.9g) misses prim %d box (%.9g %.9g %.9g)-(%.9g %.9g %.9g) tn %.9g tf %.9g\n", ri,
                        o[0], o[1], o[2], d[0], d[1], d[2], i, f.prim_boxes[(size_t)i * 6], f.prim_boxes[(size_t)i * 6 + 1], f.prim_boxes[(size_t)i * 6 + 2],
                        f.prim_boxes[(size_t)i * 6 + 3], f.prim_boxes[(size_t)i * 6 + 4], f.prim_boxes[(size_t)i * 6 + 5], tn, tf);
                }
            }
        }
        out[0]++;
    }
    if (out[4]) { err = "the quantised traversal missed a primitive box the ray crosses"; return false; }
    return true;
}

// Cost probe for the wide tree (host-only, tuning aid for the collapse heuristic): closest-hit traversal ON THE CPU of
// Lambertian-like secondary rays (origin on a random sphere, direction = normal + unit vector) through the quantised
// nodes; spheres are intersected in f64.  out: rays, node visits, primitive tests, occupied slots of the visited nodes, hits.
bool wide_cost_probe(const FlatScene& f, int n_rays, uint64_t seed, uint64_t out[5]) {
    for (int i = 0; i < 5; ++i) out[i] = 0;
    const int n = f.n_bvh_prims;
    if (n == 0 || f.wnodes.empty()) return true;
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 777;
    auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (double)((st >> 11) & ((1ull << 53) - 1)) / (double)(1ull << 53); };
    auto unit = [&](double v[3]) { for (;;) { double l = 0; for (int a = 0; a < 3; ++a) { v[a] = 2 * rnd() - 1; l += v[a] * v[a]; } if (l < 1 && l > 1e-6) { l = std::sqrt(l); for (int a = 0; a < 3; ++a) v[a] /= l; return; } } };
    for (int ri = 0; ri < n_rays; ++ri) {
        int pa; do { pa = (int)(rnd() * n) % n; } while (f.prims[pa].type != PRIM_SPHERE || f.prims[pa].s.r > 100.0);
        const DPrim& sp = f.prims[pa];
        double nrm[3], u[3]; unit(nrm); if (nrm[1] < 0) nrm[1] = -nrm[1]; unit(u);
        float o[3], d[3];
        const double c[3] = {sp.s.cx, sp.s.cy, sp.s.cz};
        for (int a = 0; a < 3; ++a) { o[a] = (float)(c[a] + sp.s.r * nrm[a]); d[a] = (float)(nrm[a] + u[a]); }
        rtww::WRay r;
        auto sdir = [](float v) { return std::fabs(v) < 1e-20f ? std::copysign(1e-20f, v) : v; };
        r.ix = 1.0f / sdir(d[0]); r.iy = 1.0f / sdir(d[1]); r.iz = 1.0f / sdir(d[2]);
        r.oix = o[0] * r.ix; r.oiy = o[1] * r.iy; r.oiz = o[2] * r.iz;
        r.sx = 2.384185791015625e-07f * std::fabs(r.oix); r.sy = 2.384185791015625e-07f * std::fabs(r.oiy); r.sz = 2.384185791015625e-07f * std::fabs(r.oiz);
        const uint32_t oct = (r.ix < 0 ? 1u : 0u) | (r.iy < 0 ? 2u : 0u) | (r.iz < 0 ? 4u : 0u);
        r.k = oct ^ 7u; r.one = 0x3F800000u;
        float t_best = INFINITY;
        struct Grp { uint32_t base, g; };
        Grp stack[RTW_WIDE_STACK + 1]; int sp_ = 0;
        uint32_t base = 0, grp = (1u << 8) | (1u << (0u ^ r.k));
        for (;;) {
            if (!(grp & 0xffu)) { if (!sp_) break; --sp_; base = stack[sp_].base; grp = stack[sp_].g; continue; }
            const int j = 31 - __builtin_clz(grp & 0xffu);
            grp ^= 1u << j;
            const uint32_t slot = (uint32_t)j ^ r.k, imask = grp >> 8;
            const uint32_t node = base + (uint32_t)__builtin_popcount(imask & ((1u << slot) - 1u));
            if (grp & 0xffu) { stack[sp_].base = base; stack[sp_].g = grp; ++sp_; }
            const DWNode& w = f.wnodes[node];
            rtww::W4 q[5]; std::memcpy(q, &w, 80);
            uint32_t hits = rtww::wide_node_hits(q[0], q[2], q[3], q[4], r, 0.001f, t_best);
            hits &= (uint32_t)(w.imask | w.lmask);
            out[1]++; out[3] += (uint64_t)__builtin_popcount(w.imask | w.lmask);
            const uint32_t m16 = rtww::wide_perm16((hits & w.imask) | ((hits & w.lmask) << 8), r.k);
            uint32_t pl = m16 >> 8;
            while (pl) {
                const int jj = 31 - __builtin_clz(pl); pl ^= 1u << jj;
                const uint32_t s2 = (uint32_t)jj ^ r.k;
                const int pi = (int)(w.prim_base + (uint32_t)__builtin_popcount(w.lmask & ((1u << s2) - 1u)));
                out[2]++;
                const DPrim& q2 = f.prims[pi];
                if (q2.type != PRIM_SPHERE || pi == pa) continue;
                const double oc[3] = {o[0] - q2.s.cx, o[1] - q2.s.cy, o[2] - q2.s.cz};
                const double a_ = (double)d[0] * d[0] + (double)d[1] * d[1] + (double)d[2] * d[2];
                const double hb = oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2];
                const double cc = oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - q2.s.r * q2.s.r;
                const double disc = hb * hb - a_ * cc;
                if (disc < 0) continue;
                const double sq = std::sqrt(disc);
                double t = (-hb - sq) / a_;
                if (t < 0.001 || t > t_best) { t = (-hb + sq) / a_; if (t < 0.001 || t > t_best) continue; }
                t_best = (float)t;
            }
            base = w.child_base; grp = ((uint32_t)w.imask << 8) | (m16 & 0xffu);
        }
        out[0]++; if (t_best < INFINITY) out[4]++;
    }
    return true;
}

}  // namespace rtw
