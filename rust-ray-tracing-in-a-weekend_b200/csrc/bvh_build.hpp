// bvh_build.hpp — interface of the device-side BVH builder (bvh_build.cu).
#ifndef RTW_BVH_BUILD_HPP
#define RTW_BVH_BUILD_HPP

#include <cuda_runtime.h>

#include <cstddef>
#include <string>

#include "bvh_wide.h"
#include "rtw_types.h"

namespace rtwb {

// one rtw_sphere_batch sphere as the host keeps it (= rtw::BulkSphere of scene_host.hpp: 40 bytes)
struct BulkSphereD { double c[3]; double r; int mat; int pad; };

struct BuildInput {
    const DPrim* host_prims = nullptr;     // primitives flattened on the host (scene-graph nodes), with their f32 boxes
    const float* host_boxes = nullptr;     // 6 floats each, rounded outward
    int n_host = 0;
    const BulkSphereD* bulk = nullptr;     // spheres uploaded as the constructors got them (host pointer)
    int n_bulk = 0;
    const DPrim* boundary_prims = nullptr; // ConstantMedium boundary records, copied behind the BVH primitives
    int n_boundary = 0;
    int width = 8;                         // 8: wide compressed nodes; 2: binary two-box nodes (rtw_types.h DNode)
};

struct BuildOutput {                       // device allocations owned by the caller (free_output)
    DWNode* wnodes = nullptr;              // width 8
    DNode* nodes = nullptr;                // width 2
    DPrim* prims = nullptr;                // n_host + n_bulk records in leaf order, then the boundary records
    int n_wnodes = 0, n_nodes = 0, n_prims = 0, depth = 0;
    size_t h2d_bytes = 0;
};

// Builds on the current device, on stream `st`; returns 0 or a negative rtw_status (-2 nesting / depth, -3 CUDA, -4 OOM).
int build_on_device(cudaStream_t st, const BuildInput& in, BuildOutput& out, std::string& err);
void free_output(BuildOutput& out);
// The builder's scratch pool of device `dev` gives its unused memory back to the driver (called when a scene is freed).
void trim_scratch(int dev);

}  // namespace rtwb
#endif
