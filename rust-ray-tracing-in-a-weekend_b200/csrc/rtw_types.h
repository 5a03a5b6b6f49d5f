// rtw_types.h — device-resident scene layout (shared by the host flattener and the CUDA kernels).
//
// The reference keeps the scene as a tree of tagged unions with Box<Hittable> children and by-value
// textures inside materials (src/hittable.rs:29-41, src/material.rs:6-12, src/texture.rs:4-9).  Here the
// scene is flattened once per commit into index-linked 16-byte-aligned records in ONE device blob:
//
//   nodes   : 64 B  binary BVH nodes holding BOTH children's boxes (4 x LDG.128 per visit)
//   prims   : 64 B  records (4 x 16 B).  BVH-referenced prims first (leaf order), medium boundaries after
//   xforms  : composed Translate/RotateY chains (src/hittable.rs:232-247, :386-415)
//   media   : ConstantMedium table (src/hittable.rs:417-473)
//   mats    : 32 B  materials, solid-colour textures inlined
//   texs    : 48 B  textures; perlin tables (float4 gradients + u8 perms); image texels (RGB8)
#ifndef RTW_TYPES_H
#define RTW_TYPES_H

#include <stdint.h>

#include "bvh_wide.h"

#if defined(__CUDACC__)
#define RTW_ALIGN(n) __align__(n)
#else
#define RTW_ALIGN(n) alignas(n)
#endif

enum : int32_t {
    PRIM_SPHERE = 0,         // Hittable::Sphere        src/hittable.rs:31
    PRIM_MOVING_SPHERE = 1,  // Hittable::MovingSphere  src/hittable.rs:32
    PRIM_XY = 2,             // XYRect :34  (normal +z; a = x, b = y)
    PRIM_XZ = 3,             // XZRect :35  (normal +y; a = x, b = z)
    PRIM_YZ = 4,             // YZRect :36  (normal +x; a = y, b = z)
    PRIM_BOX = 5             // Box :37 as ONE BVH leaf: slab test in object space -> entry / exit face; the six rect records of
                             // new_box (src/hittable.rs:132-145) live behind the BVH primitives and describe the hit (first_face + face)
};

enum : int32_t { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_DIFFUSE_LIGHT = 3, MAT_ISOTROPIC = 4 };
enum : int32_t { TEX_SOLID = 0, TEX_CHECKER = 1, TEX_NOISE = 2, TEX_IMAGE = 3 };

// 64 B = two 32-byte sectors, both used by every test (an 80-byte record left one sector in 2.5 unused in the L1 lines
// it occupied).  Spheres keep centre/radius in f64: the discriminant is evaluated in f64 on the device (B200 runs
// FP64 at half the FP32 rate), which removes the |oc|^2 - r^2 cancellation of the r = 1000 ground sphere.  The motion
// vector of a MovingSphere is f32 (|centre1 - centre0| < 1 in the reference's scenes: 3e-8 absolute).
struct RTW_ALIGN(16) DPrim {
    union {
        struct { double cx, cy, cz, r; } s;       // sphere: centre at time0 (world space, xform baked), radius
        struct { float a0, a1, b0, b1, k, pad0, pad1, pad2; } q;   // rect (object space)
        struct { float lox, hix, loy, hiy, loz, hiz; int32_t first_face, pad; } b;   // box (object space); faces in new_box order:
                                                                   // z max, z min, y max, y min, x max, x min
    };
    float dcx, dcy, dcz;                          // moving sphere: centre1 - centre0
    float t0;                                     // moving sphere: time0
    int32_t type;                                 // PRIM_*
    int32_t mat;                                  // 0-based material index (reference handle - 1)
    int32_t xform;                                // 0 = identity, else index into xforms
    float inv_dt;                                 // moving sphere: 1/(time1-time0)
};
static_assert(sizeof(DPrim) == 64, "DPrim is two sectors");

// Aila–Laine style node: both children's slabs live in the parent.
struct RTW_ALIGN(16) DNode {
    float c0minx, c0maxx, c0miny, c0maxy;
    float c1minx, c1maxx, c1miny, c1maxy;
    float c0minz, c0maxz, c1minz, c1maxz;
    int32_t child0, child1;     // >= 0: inner node index.  < 0: leaf, ~code, code = first << 3 | (count - 1)
    int32_t pad0, pad1;
};

#define RTW_MAX_CHAIN 4
// One Translate/RotateY chain.  World -> object: o_obj = R(m_cos, m_sin) * o_w + b, d_obj = R * d_w, with
// R(c, s) (x, z) = (c x - s z, s x + c z)  (src/hittable.rs:390-394).  ops[] (innermost LAST) replay the
// reference's nested set_face_normal calls (src/hittable.rs:238, :409): `cum` rotates d_w into the space
// inside op j.
struct RTW_ALIGN(16) DXform {
    float m_cos, m_sin, bx, by;
    float bz; int32_t n_ops; int32_t pad0, pad1;
    struct { float op_cos, op_sin, cum_cos, cum_sin; } ops[RTW_MAX_CHAIN];   // op_sin = 0 & op_cos = 1: Translate
    int32_t is_rot[RTW_MAX_CHAIN];
    double d_cos, d_sin, d_bx, d_by, d_bz, d_pad;   // the composed transform again, in f64: ray ORIGINS are moved into
                                                    // object space in f64 (|o| ~ 500 next to a plane offset ~ 1)
};

struct RTW_ALIGN(16) DMedium {
    int32_t first, count;       // boundary prims [first, first+count)
    float neg_inv_density;      // src/hittable.rs:205
    int32_t mat;                // phase function, 0-based
};

struct RTW_ALIGN(16) DMat {
    float r, g, b, param;       // albedo / inline solid colour; param = fuzz (Metal) or ir (Dielectric)
    int32_t kind;               // MAT_*
    int32_t tex;                // -1: solid colour inlined in rgb; else texture index
    int32_t pad0, pad1;
};

struct RTW_ALIGN(16) DTex {
    float r0, g0, b0, scale;    // Solid: rgb0; Checker: even; Noise: scale
    float r1, g1, b1; int32_t kind;   // Checker: odd
    int32_t a, w, h, bps;       // Noise: a = perlin table index; Image: a = byte offset into image blob
};

#define RTW_PERLIN_BYTES (256 * 16 + 768)   // float4 ranvec[256] + u8 perm_x/y/z[256]

struct DScene {
    const DNode* nodes;         // binary BVH, or nullptr when the scene uses the wide one
    const DWNode* wnodes;       // 8-wide compressed BVH (bvh_wide.h), or nullptr
    const DPrim* prims;
    const DXform* xforms;
    const DMedium* media;
    const DMat* mats;
    const DTex* texs;
    const uint8_t* perlin;      // n_perlin tables of RTW_PERLIN_BYTES
    const uint8_t* image;       // image texel blob
    int32_t n_nodes, n_prims, n_bvh_prims, n_xforms, n_media, n_mats, n_texs, n_perlin;
};

// Camera in device precision.  llc_rel = lower_left_corner - origin is folded on the host in f64
// (src/camera.rs:63 evaluates it per ray; folding avoids an f32 cancellation at |origin| ~ 800).
struct DCamera {
    float ox, oy, oz, lens_radius;
    float lx, ly, lz, time0;     // llc_rel
    float hx, hy, hz, time1;     // horizontal
    float vx, vy, vz, pad0;      // vertical
    float ux, uy, uz, pad1;      // u
    float wx, wy, wz, pad2;      // v (camera "v" basis vector)
};

struct DParams {
    int32_t width, height, spp, max_depth;
    float bg_r, bg_g, bg_b, t_min;
    uint32_t seed_lo, seed_hi;
    int32_t tiles_x, tiles_y;          // 8x4-pixel tiles
    int32_t chunks, chunk_spp;         // phase A: samples [0, spp_a) of every tile in `chunks` units of `chunk_spp`
    uint32_t n_units;                  // phase A + phase B
    uint32_t n_units_a;                // = tiles * chunks
    int32_t spp_a;                     // samples handled by phase A (= spp when there is no phase B)
    int32_t chunks_b, chunk_spp_b;     // phase B: samples [spp_a, spp) in `chunks_b` small units (short end-of-frame tail)
    int32_t accumulate;                // 1: atomicAdd into the framebuffer (several units or GPUs per pixel)
    int32_t no_tile_cull;              // 1: primary rays traverse the BVH like all others (RTW_FLAG_NO_TILE_CULL)
    int32_t first_sample;              // sample index of this launch's sample 0 (progressive passes / resume)
    uint32_t philox_rk[20];            // Philox4x32-10 key schedule of `seed` (round r: key + r * Weyl constants)
    uint32_t unit_stride;              // 1; RTW_EMULATE_RANKS=k: this GPU takes every k-th unit only (tuning the k-GPU unit sizes on one GPU)
    int32_t list_max;                  // scenes of at most this many BVH primitives: secondary rays scan ALL primitives (no BVH walk)
};

#endif
