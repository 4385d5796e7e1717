// Diagnostic only: candidates for torch's F.normalize rounding, compared bitwise on the GPU.
#include "../../lara_b200/csrc/surfel_common.cuh"
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_approx_ftz(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__global__ void k(int n, const float4* q, const float* tn, float* out, float4* qd) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = q[i];
    const float a = fmul_(v.x, v.x), b = fmul_(v.y, v.y), c = fmul_(v.z, v.z), d = fmul_(v.w, v.w);
    const float seq = fadd_(fadd_(fadd_(a, b), c), d);
    const float pw = fadd_(fadd_(a, b), fadd_(c, d));
    const float pw2 = fadd_(fadd_(a, c), fadd_(b, d));
    const float fm = __fmaf_rn(v.w, v.w, __fmaf_rn(v.z, v.z, __fmaf_rn(v.y, v.y, a)));
    const double ds = (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    float* o = out + (size_t)i * 12;
    o[0] = __fsqrt_rn(seq); o[1] = __fsqrt_rn(pw); o[2] = __fsqrt_rn(pw2); o[3] = __fsqrt_rn(fm);
    o[4] = sqrt_approx(seq); o[5] = sqrt_approx(pw); o[6] = sqrt_approx(pw2); o[7] = sqrt_approx(fm);
    o[8] = (float)sqrt(ds); o[9] = __fsqrt_rn((float)ds); o[10] = sqrt_approx_ftz(pw); o[11] = sqrt_approx((float)ds);
    const float t = fmaxf(tn[i], 1e-12f);
    qd[i] = make_float4(__fdiv_rn(v.x, t), __fdiv_rn(v.y, t), __fdiv_rn(v.z, t), __fdiv_rn(v.w, t));
}
extern "C" int probe(int n, const float* q, const float* tn, float* out, float* qd) {
    k<<<(n + 255) / 256, 256>>>(n, (const float4*)q, tn, out, (float4*)qd);
    return (int)cudaDeviceSynchronize();
}
