// loss.cu -- fused loss -> dL/d(render_img outputs) producer.
//
// Restates the rasterizer-facing terms of LaRa's training loss (lightning/loss.py:33-60):
//     loss  = mean((image - tar_rgb)^2)                                   (:33-34)
//           + 1000 * mean(rend_dist)                                      (:48-51, iter > 1000)
//           + 0.2 * mean((1 - sum_c rend_normal_c * depth_normal_c) * acc_map.detach())   (:53-59)
// (the MS-SSIM term :42-46 stays a library call on `image`; its gradient simply adds to g_image).
// In torch these are ~15 elementwise / reduction kernels forward and ~20 backward, each a full-image HBM round
// trip over the [H, V*W, C] concatenation of the views.  Here:
//   loss_sums_kernel : one pass over the stacked planar outputs of the fused epilogue ([V,C,H,W]) -> three sums
//   loss_grads_kernel: one pass writing g_image / g_rend_normal / g_depth_normal / g_rend_dist, already planar
//                      and stacked, i.e. exactly what srf_views_epilogue_backward consumes -- scaled by the upstream
//                      scalar read from device memory (no host sync).
// `target` is read in the batch's own channel-last layout [V,H,W,3] (lightning/loss.py:24 permutes a view of it).
#include "surfel_common.cuh"
#include "surfel_kernels.h"

namespace srf {

__global__ void __launch_bounds__(256) loss_sums_kernel(LossArgs a) {
    const size_t npix = (size_t)a.W * a.H, total = npix * (size_t)a.nviews;
    float s_mse = 0.f, s_dist = 0.f, s_norm = 0.f;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const size_t v = i / npix, pix = i - v * npix;
        const float* img = a.image + v * 3 * npix + pix;
        const float* tar = a.target + i * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float d = img[c * npix] - __ldg(tar + c);
            s_mse = fmaf(d, d, s_mse);
        }
        if (a.with_reg) {
            s_dist += a.dist[i];
            const float* rn = a.rend_normal + v * 3 * npix + pix;
            const float* dn = a.depth_normal + v * 3 * npix + pix;
            const float dot = rn[0] * dn[0] + rn[npix] * dn[npix] + rn[2 * npix] * dn[2 * npix];
            s_norm = fmaf(1.0f - dot, a.acc[i], s_norm);
        }
    }
    // warp -> block -> one double atomic per block and term
    __shared__ float s_part[3][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_mse += __shfl_xor_sync(0xffffffffu, s_mse, o);
        s_dist += __shfl_xor_sync(0xffffffffu, s_dist, o);
        s_norm += __shfl_xor_sync(0xffffffffu, s_norm, o);
    }
    if (lane == 0) { s_part[0][wid] = s_mse; s_part[1][wid] = s_dist; s_part[2][wid] = s_norm; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += (double)s_part[threadIdx.x][w];
        atomicAdd(a.sums + threadIdx.x, t);
    }
}

__global__ void __launch_bounds__(256) loss_grads_kernel(LossArgs a) {
    const size_t npix = (size_t)a.W * a.H, total = npix * (size_t)a.nviews;
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const float up = a.gout ? __ldg(a.gout) : 1.0f;
    const size_t v = i / npix, pix = i - v * npix;
    const float* img = a.image + v * 3 * npix + pix;
    const float* tar = a.target + i * 3;
    float* gi = a.g_image + v * 3 * npix + pix;
    const float k_mse = 2.0f * a.w_mse * up;
#pragma unroll
    for (int c = 0; c < 3; ++c) gi[c * npix] = k_mse * (img[c * npix] - __ldg(tar + c));
    if (a.g_dist) a.g_dist[i] = a.with_reg ? a.w_dist * up : 0.0f;
    if (a.g_rend_normal && a.g_depth_normal) {
        float* grn = a.g_rend_normal + v * 3 * npix + pix;
        float* gdn = a.g_depth_normal + v * 3 * npix + pix;
        if (a.with_reg) {
            const float k = -a.w_normal * up * a.acc[i];
            const float* rn = a.rend_normal + v * 3 * npix + pix;
            const float* dn = a.depth_normal + v * 3 * npix + pix;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                grn[c * npix] = k * dn[c * npix];
                gdn[c * npix] = k * rn[c * npix];
            }
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) { grn[c * npix] = 0.0f; gdn[c * npix] = 0.0f; }
        }
    }
}

cudaError_t launch_loss_fused(const LossArgs& a, cudaStream_t stream) {
    const size_t total = (size_t)a.W * a.H * (size_t)a.nviews;
    if (total == 0) return cudaSuccess;
    if (a.sums != nullptr) {
        const int grid = (int)min((size_t)sm_count() * 8, (total + 255) / 256);
        loss_sums_kernel<<<grid, 256, 0, stream>>>(a);
    }
    if (a.g_image != nullptr) loss_grads_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace srf
