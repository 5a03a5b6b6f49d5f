// scene_host.hpp — host-side scene graph (mirror of the reference `World`, src/main.rs:40-50) and its
// flattening into the device layout of rtw_types.h.  No CUDA in this file.
#ifndef RTW_SCENE_HOST_HPP
#define RTW_SCENE_HOST_HPP

#include <cstdint>
#include <string>
#include <vector>

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
    std::vector<DNode> nodes;
    std::vector<DPrim> prims;        // [0, n_bvh_prims) in leaf order, then medium boundary prims
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

// Flatten `roots` (world.hittables, or a single hittable for the test hooks).  Returns 0 or a negative
// rtw_status; `err` gets a message.
int flatten(const SceneGraph& g, const std::vector<int>& roots, FlatScene& out, std::string& err);

// Structural check used by the CPU tests: every BVH prim referenced exactly once and inside its ancestors' boxes.
bool validate_bvh(const FlatScene& f, std::string& err);

}  // namespace rtw

#endif
