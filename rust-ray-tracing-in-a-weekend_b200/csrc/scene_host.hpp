// scene_host.hpp — host-side scene graph (mirror of the reference `World`, src/main.rs:40-50) and its
// flattening into the device layout of rtw_types.h.  No CUDA in this file.
#ifndef RTW_SCENE_HOST_HPP
#define RTW_SCENE_HOST_HPP

#include <cstdint>
#include <string>
#include <vector>

#include "bvh_wide.h"
#include "rtw_types.h"

namespace rtw {

struct V3d { double x = 0, y = 0, z = 0; };

enum HKind { H_SPHERE, H_MOVING_SPHERE, H_BVH_NODE, H_XY, H_XZ, H_YZ, H_BOX, H_TRANSLATE, H_ROTATE_Y, H_MEDIUM };

// One reference Hittable (src/hittable.rs:29-41); children by id (ids may be shared = Rust clone()).
struct HNode {
    int kind = H_SPHERE;
    int mat = 0;                                   // 1-based handle
    V3d c0, c1; double radius = 0, time0 = 0, time1 = 1;
    double a0 = 0, a1 = 0, b0 = 0, b1 = 0, k = 0;
    V3d bmin, bmax;                                // Box
    V3d offset; double angle_deg = 0, sin_theta = 0, cos_theta = 1;
    double density = 0;
    int child = -1;
    std::vector<int> children;                     // BvhNode members / Box sides
};

struct HTexture {
    int kind = TEX_SOLID;
    double c0[3] = {0, 0, 0}, c1[3] = {0, 0, 0};
    double scale = 1;
    std::vector<double> ranvec;                    // 256*3
    std::vector<int32_t> perm;                     // 3*256
    int w = 0, h = 0, bps = 0;
    std::vector<uint8_t> data;
};

struct HMaterial {
    int kind = MAT_LAMBERTIAN;
    int tex = -1;
    double albedo[3] = {0, 0, 0};
    double fuzz = 0, ir = 1;
};

// compact record for rtw_sphere_batch (the 1M-16M sphere sweep would not fit as HNodes)
struct BulkSphere { double c[3]; double r; int mat; };

struct SceneGraph {
    std::vector<HTexture> textures;
    std::vector<HMaterial> materials;
    std::vector<HNode> nodes;
    std::vector<int> world;
    std::vector<BulkSphere> bulk;          // all of them are world members (pushed after `world`)
};

// The flattened scene, still on the host, ready to be packed into one blob.
struct FlatScene {
    std::vector<DNode> nodes;        // binary BVH (empty when the scene uses the wide one)
    std::vector<DWNode> wnodes;      // 8-wide compressed BVH (bvh_wide.h); [0] = root
    bool wide = false;
    int wide_depth = 0;              // levels of the wide tree (bounds the traversal stack)
    std::vector<float> prim_boxes;   // 6 floats per BVH prim in final order (only with FlattenOptions::keep_boxes: validation)
    std::vector<DPrim> prims;        // [0, n_bvh_prims) in leaf order, then medium boundary prims
    std::vector<DPrim> boundary;     // emit_only: the medium boundary records (they follow the BVH prims on the device)
    bool emit_only = false;
    int32_t n_bvh_prims = 0;
    std::vector<DXform> xforms;      // [0] = identity
    std::vector<DMedium> media;
    std::vector<DMat> mats;
    std::vector<DTex> texs;
    std::vector<uint8_t> perlin;
    std::vector<uint8_t> image;
    double sah_cost = 0;
    int max_depth = 0;
    int features = 0;                // FEAT_* bits (rtw_device.cuh) the scene needs: picks the kernel variant
    int n_dedup = 0;                 // BvhNode members dropped because a field-by-field equal member came before (cloned leaves)
    // Intersection of the [time0, time1] intervals of all MovingSpheres: their BVH boxes span exactly that interval (like the
    // reference's BvhNode boxes, src/hittable.rs:480-482), so a camera shutter must stay inside it (checked at render).
    double mov_t0 = -1e300, mov_t1 = 1e300;
};

// bvh_width: 2 = binary nodes, 8 = wide compressed nodes, 0 = by size (RTW_BVH=2|8 overrides; scenes of at least
// RTW_WIDE_MIN primitives, default kWideMinPrims, go wide: they no longer fit the caches).
// emit_only: stop after the primitive records — no BVH, rtw_sphere_batch spheres left alone (`prims` = the scene-graph
// primitives in emission order with their boxes in `prim_boxes`, `boundary` = the medium boundary records, n_bvh_prims =
// scene-graph primitives + bulk spheres): the input of the device-side builder (bvh_build.cu).
struct FlattenOptions { int bvh_width = 0; bool keep_boxes = false; bool emit_only = false; };
// Scenes of at least this many primitives get wide nodes.  Measured on the B200 (profiles/r2_*): ... see DESIGN.md 4.5.
extern int kWideMinPrims;
int choose_bvh_width(long long n_prims, int requested);

// Flatten `roots` (world.hittables, or a single hittable for the test hooks).  Returns 0 or a negative
// rtw_status; `err` gets a message.
int flatten(const SceneGraph& g, const std::vector<int>& roots, FlatScene& out, std::string& err, const FlattenOptions& opt = FlattenOptions());

// Host-side BVH2 -> BVH8 collapse (the device builder runs the same rtww::collapse_one per node).  `box`: 6 floats per
// ref as in rtww::B2View.  Fills wnodes / leaf_order (final position -> leaf number); returns the depth of the wide tree.
int collapse_to_wide(const rtww::B2View& v, std::vector<DWNode>& wnodes, std::vector<int>& leaf_order);

// Structural check of the wide BVH: every BVH prim referenced exactly once, every child box contains the boxes below it
// (needs prim_boxes), masks / ranges consistent.
bool validate_wide(const FlatScene& f, std::string& err);

// Conservativeness of the quantised traversal, on the CPU with the arithmetic the device uses (rtww::wide_node_hits):
// for `n_rays` seeded rays the leaves the traversal reaches must include every primitive whose box the ray really
// crosses.  out: [0] rays, [1] node visits, [2] leaves reached, [3] boxes really crossed, [4] MISSED (must be 0).
bool check_wide_traversal(const FlatScene& f, int n_rays, uint64_t seed, uint64_t out[5], std::string& err);
// tuning aid: closest-hit cost of secondary-like rays through the wide tree on the CPU (rays, node visits, primitive tests, occupied slots, hits)
bool wide_cost_probe(const FlatScene& f, int n_rays, uint64_t seed, uint64_t out[5]);

// Structural check used by the CPU tests: every BVH prim referenced exactly once and inside its ancestors' boxes.
bool validate_bvh(const FlatScene& f, std::string& err);

}  // namespace rtw

#endif
