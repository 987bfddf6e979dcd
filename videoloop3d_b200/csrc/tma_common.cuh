// Shared pieces of the TMA-staged composite kernels: box geometry, kernel parameter block, mbarrier / bulk-tensor
// PTX wrappers and the host-side tensor-map encoder (included by composite.cu before the kernels).
#pragma once

#include "tma_ptx.cuh"

namespace vl3d {

constexpr int TMA_BW = 40, TMA_BH = 12;
constexpr int TMA_BOX_BYTES = TMA_BW * TMA_BH * 16;
constexpr int TMA_THREADS = BX * BY + 32;                          // 8 consumer warps + 1 producer warp

struct alignas(64) TmaRenderParams {
    CUtensorMap tmap;                                               // (x4 = dyn_w*4 floats, y = dyn_h, t = frames)
    CompositeParams p;
};

// tensor map of the dynamic atlas: (x4 = dyn_w*4 floats, y = dyn_h, t = frames), box = TMA_BW texels x TMA_BH rows x 1 frame
static bool make_atlas_tmap(CUtensorMap* out, const vl3d_view& view, const float* atlas_dyn, int frames) {
    vl3d_encode_tiled_fn enc = tma_encoder();
    if (enc == nullptr) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)view.dyn_w * 4, (cuuint64_t)view.dyn_h, (cuuint64_t)frames};
    const cuuint64_t gstr[2] = {(cuuint64_t)view.dyn_w * 16, (cuuint64_t)view.dyn_w * view.dyn_h * 16};
    const cuuint32_t box[3] = {TMA_BW * 4, TMA_BH, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (gstr[1] >= ((cuuint64_t)1 << 40)) return false;
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(atlas_dyn), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace vl3d
