// surfel_kernels.h -- internal (non-ABI) interface between the C-ABI layer (api.cu)
// and the kernel translation units.  The public boundary is include/surfel_rasterizer.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace srf {

struct PreprocessArgs {
    int P, D, M;
    const float* means3D;        // [P,3]
    const float* scales;         // [P,2]
    const float* rotations;      // [P,4] (w,x,y,z)
    const float* opacities;      // [P]
    const float* shs;            // [P,M,3] or null
    const float* transMat_precomp;  // [P,9] or null
    const float* colors_precomp;    // [P,3] or null
    const float* viewmatrix;     // [16] column-major world->view
    const float* campos;         // [3]
    int W, H;
    float focal_x, focal_y;
    int gx, gy;                  // tile grid
    int prefiltered;
    int stage_sh;                // set by the launcher
    int raw_act;                 // inputs are LaRa's raw network outputs: opacity logits, log-scales, unnormalised quaternions
    int* radii;                  // [P] out
    float4* rec;                 // [P*6] out
    float* depths;               // [P] out
    uint2* rects;                // [P] out: (x0 | y0<<16, x1 | y1<<16)
    uint32_t* tile_count;        // [ntiles * SRF_TILE_CTR_STRIDE] in/out (zeroed by the caller)
};

struct BinArgs {
    int P;
    int ntiles;
    int gx;
    uint32_t capacity;           // number of instance slots in entries / point_list
    const float* depths;
    const uint2* rects;          // empty rect = culled
    uint32_t* tile_count;        // [ntiles * SRF_TILE_CTR_STRIDE]: word 0 count, word 1 bucket cursor
    uint2* ranges;               // [ntiles] out
    uint32_t* counters;          // [4]: num_rendered, big_count, overflow, spare
    uint32_t* big_list;          // [ntiles] scratch
    uint32_t* tile_order;        // [ntiles] out: tiles by descending instance count (launch order of the blend CTAs)
    uint64_t* entries;           // [capacity] scratch: depth_bits<<32 | gaussian idx, bucketed by tile
    uint32_t* point_list;        // [capacity] out: per-tile depth-sorted gaussian indices
};

struct RenderFwdArgs {
    int W, H, gx, gy;
    uint32_t capacity;
    const uint2* ranges;
    const uint32_t* tile_order;  // [ntiles]: CTA i renders tile tile_order[i] (heaviest first)
    const uint32_t* point_list;
    const float4* rec;
    const float* bg;             // [3] device
    float* out_color;            // [3,H,W]
    float* out_others;           // [8,H,W]
    float* accum;                // [3,H,W]: final_T, dist1, dist2
    uint32_t* n_contrib;         // [2,H,W]: last contributor, median contributor
};

struct RenderBwdArgs {
    int W, H, gx, gy;
    uint32_t capacity;
    const uint2* ranges;
    const uint32_t* tile_order;
    const uint32_t* point_list;
    const float4* rec;
    const float* bg;
    const float* accum;
    const uint32_t* n_contrib;
    const float* dL_dpix;        // [3,H,W]
    const float* dL_dothers;     // [8,H,W]
    float* ggrad;                // [P,20] zero-initialised accumulation records
};

struct PreprocessBwdArgs {
    int P, D, M;
    const float* means3D;
    const float* scales;
    const float* rotations;
    const float* shs;
    const float* viewmatrix;
    const float* campos;
    int W, H;
    float focal_x, focal_y, tan_fovx, tan_fovy;
    int has_precomp_T;           // transMat_precomp was given: only dL_dtransMat is produced
    int has_precomp_color;
    int raw_act;                 // as in PreprocessArgs; gradients are then wrt the raw parameters
    const int* radii;
    const float4* rec;
    const float* ggrad;          // [P,20]
    int accumulate;              // += into the outputs instead of overwriting (view-sharded accumulation)
    int vec_ok;                  // set by the launcher: output rows are 16/8-byte aligned -> vector stores
    float* dL_dmeans3D;          // [P,3]
    float* dL_dmeans2D;          // [P,3] or null
    float* dL_dsh;               // [P,M,3] or null
    float* dL_dcolors;           // [P,3] or null
    float* dL_dopacity;          // [P]
    float* dL_dscales;           // [P,2]
    float* dL_drotations;        // [P,4]
    float* dL_dtransMat;         // [P,9] or null
};

// Fused render_img epilogue; all images planar [C,H,W] fp32.
struct EpilogueArgs {
    int W, H;
    float depth_ratio;
    const float* color;          // [3,H,W] rasterizer colour
    const float* allmap;         // [8,H,W] rasterizer aux maps
    const float* rays;           // [H,W,6] or null (then no pseudo normals)
    const float* viewmatrix;     // [16]
    // forward outputs
    float* image;                // [3,H,W]
    float* depth;                // [1,H,W]
    float* acc;                  // [H,W]
    float* rend_normal;          // [3,H,W]
    float* depth_normal;         // [3,H,W]
    float* dist;                 // [H,W]
    // backward inputs (any may be null = zero) and outputs
    const float* g_image; const float* g_depth; const float* g_acc;
    const float* g_rend_normal; const float* g_depth_normal; const float* g_dist;
    float* scratch;              // [3,H,W]
    float* dL_dcolor;            // [3,H,W]
    float* dL_dallmap;           // [8,H,W]
};
cudaError_t launch_epilogue_fwd(const EpilogueArgs& a, cudaStream_t stream);
cudaError_t launch_epilogue_bwd(const EpilogueArgs& a, cudaStream_t stream);

// Optional per-kernel CUDA-event timing (srf_profile_begin/end in the C ABI); no-ops unless enabled.
enum KernelId { K_PREPROCESS_FWD = 0, K_TILE_SCAN, K_SCATTER, K_SORT_SMALL, K_SORT_BIG, K_RENDER_FWD,
                K_RENDER_BWD, K_PREPROCESS_BWD, K_COUNT };
void prof_start(int kernel, cudaStream_t stream);
void prof_stop(int kernel, cudaStream_t stream);

cudaError_t launch_preprocess_fwd(const PreprocessArgs& a, cudaStream_t stream);
cudaError_t launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present,
                                cudaStream_t stream);
cudaError_t launch_tile_scan(const BinArgs& a, cudaStream_t stream);
cudaError_t launch_bin_and_sort(const BinArgs& a, cudaStream_t stream);
cudaError_t launch_render_fwd(const RenderFwdArgs& a, cudaStream_t stream);
cudaError_t launch_render_bwd(const RenderBwdArgs& a, cudaStream_t stream);
cudaError_t launch_preprocess_bwd(const PreprocessBwdArgs& a, cudaStream_t stream);

}  // namespace srf
