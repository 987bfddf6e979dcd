// Shared pieces of the TMA-staged composite kernels: box geometry, kernel parameter block, mbarrier / bulk-tensor
// PTX wrappers and the host-side tensor-map encoder (included by composite.cu before the kernels).
#pragma once

#include <cuda.h>

namespace vl3d {

constexpr int TMA_BW = 40, TMA_BH = 12;
constexpr int TMA_BOX_BYTES = TMA_BW * TMA_BH * 16;
constexpr int TMA_THREADS = BX * BY + 32;                          // 8 consumer warps + 1 producer warp

struct alignas(64) TmaRenderParams {
    CUtensorMap tmap;                                               // (x4 = dyn_w*4 floats, y = dyn_h, t = frames)
    CompositeParams p;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

typedef CUresult (*vl3d_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on libcuda)
static vl3d_encode_tiled_fn tma_encoder() {
    static vl3d_encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<vl3d_encode_tiled_fn>(ptr);
        (void)cudaGetLastError();
    }
    return fn;
}

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
