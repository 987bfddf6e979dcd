// Adam step for one flat tensor + library bookkeeping (version, last error).
//
// Replaces torch.optim.Adam(betas=(0.9,0.999), eps=6e-8) as configured by MPMeshVid.get_optimizer
// (MPV.py:200-218); no weight decay, no amsgrad.  Same operation order as torch's single-tensor path:
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// (adam1 in vl3d_common.cuh: MUFU sqrt / reciprocal)
// Pure streaming kernel: 4 reads + 3 writes of 4 bytes per element, float4-vectorised.
#include <stdarg.h>
#include <string.h>

#include "vl3d_common.cuh"

namespace vl3d {

static thread_local char g_err[512] = "";

char* err_buf() { return g_err; }

int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return 0;
    return set_err((int)e, "%s: %s", what, cudaGetErrorString(e));
}

__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                   float4* __restrict__ m, float4* __restrict__ v, size_t n4,
                                                   float b1, float b2, float step_size, float inv_sqrt_bc2, float eps) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
        adam1(pp.x, gg.x, mm.x, vv.x, b1, b2, step_size, inv_sqrt_bc2, eps);
        adam1(pp.y, gg.y, mm.y, vv.y, b1, b2, step_size, inv_sqrt_bc2, eps);
        adam1(pp.z, gg.z, mm.z, vv.z, b1, b2, step_size, inv_sqrt_bc2, eps);
        adam1(pp.w, gg.w, mm.w, vv.w, b1, b2, step_size, inv_sqrt_bc2, eps);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

__global__ void adam_tail_kernel(float* p, const float* g, float* m, float* v, size_t start, size_t n, float b1,
                                 float b2, float step_size, float inv_sqrt_bc2, float eps) {
    const size_t i = start + threadIdx.x;
    if (i < n) {
        float pp = p[i], mm = m[i], vv = v[i];
        adam1(pp, g[i], mm, vv, b1, b2, step_size, inv_sqrt_bc2, eps);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

}  // namespace vl3d

using namespace vl3d;

extern "C" int vl3d_version(void) { return VL3D_VERSION; }

extern "C" const char* vl3d_last_error_string(void) { return err_buf(); }

extern "C" int vl3d_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int32_t step, float lr,
                              float beta1, float beta2, float eps, void* stream) {
    VL3D_REQUIRE(p && g && m && v, VL3D_ENULL, "adam: NULL pointer");
    VL3D_REQUIRE(n >= 0 && step >= 1, VL3D_EINVAL, "adam: n=%lld step=%d", (long long)n, step);
    VL3D_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, VL3D_EALIGN,
                 "adam: pointers must be 16-byte aligned");
    if (n == 0) return 0;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    const size_t n4 = (size_t)n / 4;
    cudaStream_t st = (cudaStream_t)stream;
    if (n4) {
        size_t blocks = (n4 + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        adam_kernel<<<(unsigned)blocks, 256, 0, st>>>((float4*)p, (const float4*)g, (float4*)m, (float4*)v, n4, beta1,
                                                      beta2, step_size, inv_sqrt_bc2, eps);
        if (int e = check_launch("adam")) return e;
    }
    if (n4 * 4 < (size_t)n) {
        adam_tail_kernel<<<1, 4, 0, st>>>(p, g, m, v, n4 * 4, (size_t)n, beta1, beta2, step_size, inv_sqrt_bc2, eps);
        return check_launch("adam_tail");
    }
    return 0;
}
