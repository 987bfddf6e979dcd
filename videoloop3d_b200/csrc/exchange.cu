// Strided box copies into (possibly peer) memory: the data movement of the T-sharded step's exchanges
// (frames -> row bands of the rendered video, dL/drgb back to the frame owners).
//
// With NVLink peer mappings (torch symmetric memory) a rank stores its rows STRAIGHT into the band / frame buffers of
// the ranks that need them: one launch replaces the pack kernels + NCCL all-to-all + unpack kernels of an exchange
// (N+1 launches and a staging copy of every byte on both sides).  A box is (frames, planes, rows, cols) with element
// strides on both sides; `src2` (optional) is added on the fly — the adjoint of the loop pad cat(rgb, rgb[:pad])
// (MPV.py:490-492) — otherwise the copy is bit-exact (int32 NN maps travel as 32-bit words).
#include "vl3d_common.cuh"

namespace vl3d {

struct BoxParams {
    vl3d_box b[VL3D_MAX_BOXES];
    int n;
};

__global__ void __launch_bounds__(256) copy_boxes_kernel(const __grid_constant__ BoxParams P) {
    const vl3d_box& B = P.b[blockIdx.y];
    const long long rows_total = (long long)B.n_frames * B.n_planes * B.n_rows;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const bool vec = (B.n_cols & 3) == 0 && ((B.src_sf | B.src_sp | B.src_sr | B.dst_sf | B.dst_sp | B.dst_sr) & 3) == 0 &&
                     (((uintptr_t)B.src | (uintptr_t)B.dst | (uintptr_t)B.src2) & 15) == 0;
    for (long long row = (long long)blockIdx.x * wpb + warp; row < rows_total; row += (long long)gridDim.x * wpb) {
        const int r = (int)(row % B.n_rows);
        const long long fp = row / B.n_rows;
        const int pl = (int)(fp % B.n_planes), f = (int)(fp / B.n_planes);
        const size_t so = (size_t)f * B.src_sf + (size_t)pl * B.src_sp + (size_t)r * B.src_sr;
        const float* s = B.src + so;
        const float* s2 = B.src2 ? B.src2 + so : nullptr;
        float* d = B.dst + (size_t)f * B.dst_sf + (size_t)pl * B.dst_sp + (size_t)r * B.dst_sr;
        if (vec) {
            for (int c = lane * 4; c < B.n_cols; c += 128) {
                float4 v = __ldcs(reinterpret_cast<const float4*>(s + c));
                if (s2) {
                    const float4 w = __ldcs(reinterpret_cast<const float4*>(s2 + c));
                    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
                }
                *reinterpret_cast<float4*>(d + c) = v;
            }
        } else {
            for (int c = lane; c < B.n_cols; c += 32) d[c] = s2 ? s[c] + s2[c] : s[c];
        }
    }
}

}  // namespace vl3d

using namespace vl3d;

extern "C" int vl3d_copy_boxes(const vl3d_box* boxes, int32_t n_boxes, void* stream) {
    VL3D_REQUIRE(n_boxes >= 0 && n_boxes <= VL3D_MAX_BOXES, VL3D_ERANGE, "copy_boxes: %d boxes (max %d)", n_boxes, VL3D_MAX_BOXES);
    if (n_boxes == 0) return 0;
    VL3D_REQUIRE(boxes != nullptr, VL3D_ENULL, "copy_boxes: boxes is NULL");
    BoxParams P{};
    long long most = 0;
    int n = 0;
    for (int i = 0; i < n_boxes; ++i) {
        const vl3d_box& b = boxes[i];
        VL3D_REQUIRE(b.n_frames >= 0 && b.n_planes >= 0 && b.n_rows >= 0 && b.n_cols >= 0, VL3D_EINVAL, "copy_boxes: negative extent");
        const long long rows = (long long)b.n_frames * b.n_planes * b.n_rows;
        if (rows == 0 || b.n_cols == 0) continue;
        VL3D_REQUIRE(b.src && b.dst, VL3D_ENULL, "copy_boxes: box %d has a NULL pointer", i);
        P.b[n++] = b;
        if (rows > most) most = rows;
    }
    if (n == 0) return 0;
    P.n = n;
    long long bx = (most + 7) / 8;
    if (bx > 148 * 8 / n + 1) bx = 148 * 8 / n + 1;                   // ~8 CTAs per SM over all boxes
    if (bx < 1) bx = 1;
    copy_boxes_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(P);
    return check_launch("copy_boxes");
}
