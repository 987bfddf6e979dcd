// L2 residency probe for B200 (tuning aid, not part of libvl3d): which store / RED / load eviction-priority hints keep
// a "gradient-like" working set G resident in L2 while a much larger stream flows through it?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_probe scripts/l2_probe.cu
//   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv ./l2_probe
//
// Each experiment = zero(G) ; stream(S bytes) ; red(G) ; stream(S) ; read(G), every phase its own kernel so that ncu
// reports the DRAM bytes per phase: red(G) reading ~0 bytes means the zeroed lines survived the stream; read(G)
// reading ~0 bytes means the RED-dirtied lines survived.  argv: [G MB] [S MB] [persist MB (0 = leave the limit alone)]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr uint64_t EVICT_FIRST = 0x12F0000000000000ull, EVICT_LAST = 0x14F0000000000000ull;

template <int POL>   // 0 normal, 1 256-bit store with the L2::evict_last qualifier, 2 cache_hint evict_last
__global__ void k_zero(float4* g, size_t n4) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        if (POL == 0) g[i] = z;
        else if (POL == 1) { if ((i & 1) == 0) asm volatile("st.global.L2::evict_last.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(g + i), "r"(0) : "memory"); }
        else asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%1,%1,%1}, %2;" ::"l"(g + i), "f"(0.f), "l"(EVICT_LAST) : "memory");
    }
}

template <int POL>   // 0 normal, 1 .cs, 2 cache_hint evict_first + L1::no_allocate loads, 3 cache_hint evict_first
__global__ void k_stream(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v;
        if (POL == 0) v = src[i];
        else if (POL == 1) v = __ldcs(src + i);
        else if (POL == 2) asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src + i), "l"(EVICT_FIRST) : "memory");
        else asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src + i), "l"(EVICT_FIRST) : "memory");
        v.x += 1.f;
        if (POL == 0) dst[i] = v;
        else if (POL == 1) __stcs(dst + i, v);
        else if (POL == 2) asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(dst + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(EVICT_FIRST) : "memory");
        else asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(dst + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(EVICT_FIRST) : "memory");
    }
}

template <int POL>   // 0 normal red.v4, 1 red.v4 with cache_hint evict_last, 2 scalar atomicAdd x4
__global__ void k_red(float4* g, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        if (POL == 0) asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(g + i), "f"(1.f) : "memory");
        else if (POL == 1) asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1,%1,%1,%1}, %2;" ::"l"(g + i), "f"(1.f), "l"(EVICT_LAST) : "memory");
        else { float* f = reinterpret_cast<float*>(g + i); atomicAdd(f, 1.f); atomicAdd(f + 1, 1.f); atomicAdd(f + 2, 1.f); atomicAdd(f + 3, 1.f); }
    }
}

template <int POL>   // 0 ld.cg, 1 evict_first
__global__ void k_read(const float4* g, size_t n4, float* out) {
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v;
        if (POL == 0) v = __ldcg(g + i);
        else asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(g + i), "l"(EVICT_FIRST) : "memory");
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == -1.f) *out = acc;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
    const size_t g_mb = argc > 1 ? atoi(argv[1]) : 32, s_mb = argc > 2 ? atoi(argv[2]) : 512;
    const int persist_mb = argc > 3 ? atoi(argv[3]) : 0;
    int dev = 0, maxp = 0, l2 = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev));
    CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
    printf("L2 %d MB, max persisting %d MB\n", l2 >> 20, maxp >> 20);
    if (persist_mb > 0) {
        size_t want = (size_t)persist_mb << 20;
        if (want > (size_t)maxp) want = maxp;
        cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
        printf("persisting limit: asked %zu MB -> %s, now %zu MB\n", want >> 20, cudaGetErrorString(e), got >> 20);
    }
    const size_t gn4 = (g_mb << 20) / 16, sn4 = (s_mb << 20) / 16 / 2;   // stream: S/2 read + S/2 written
    float4 *G, *A, *B;
    float* out;
    CK(cudaMalloc(&G, gn4 * 16));
    CK(cudaMalloc(&A, sn4 * 16));
    CK(cudaMalloc(&B, sn4 * 16));
    CK(cudaMalloc(&out, 4));
    CK(cudaMemset(A, 0, sn4 * 16));
    const int grid = 148 * 8, blk = 256;
#define EXPERIMENT(ZP, SP, RP, DP)                                   \
    k_stream<0><<<grid, blk>>>(A, B, sn4); /* flush */               \
    k_zero<ZP><<<grid, blk>>>(G, gn4);                               \
    k_stream<SP><<<grid, blk>>>(A, B, sn4);                          \
    k_red<RP><<<grid, blk>>>(G, gn4);                                \
    k_stream<SP><<<grid, blk>>>(A, B, sn4);                          \
    k_read<DP><<<grid, blk>>>(G, gn4, out);                          \
    CK(cudaDeviceSynchronize());
    // launches per experiment: flush, zero, stream, red, stream, read  (6)
    EXPERIMENT(0, 0, 0, 0)   // everything normal
    EXPERIMENT(0, 1, 0, 0)   // stream .cs
    EXPERIMENT(0, 2, 0, 0)   // stream L2::evict_first
    EXPERIMENT(0, 3, 0, 0)   // stream cache_hint evict_first
    EXPERIMENT(1, 2, 0, 0)   // zero with L2::evict_last, stream evict_first
    EXPERIMENT(2, 2, 1, 0)   // zero + red with cache_hint evict_last, stream evict_first
    EXPERIMENT(1, 0, 1, 0)   // zero/red evict_last, stream normal
    EXPERIMENT(1, 2, 2, 1)   // zero evict_last, stream evict_first, scalar atomics, read evict_first
    printf("done\n");
    return 0;
}
