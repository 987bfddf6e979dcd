// Micro-probe (development aid): fp32 squared-difference accumulation rate per SM on sm_100a,
//   scalar:  d = a - b (FADD);   acc = fma(d, d, acc) (FFMA)            -- the search kernel's inner step today
//   packed:  d2 = a2 - b2 (FADD2); acc2 = fma(d2, d2, acc2) (FFMA2)     -- two elements per instruction
// 32 accumulators (scalar) / 32 accumulator pairs (packed) per thread, 4 x 8 operand tiles held in registers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pack_probe fp32_pack_probe.cu && ./fp32_pack_probe
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE>
__global__ void __launch_bounds__(256, 2) probe(float* out, const float4* in, int iters) {
    float4 xa[4], ya[8];
    for (int i = 0; i < 4; ++i) xa[i] = in[threadIdx.x + 256 * i];
    for (int j = 0; j < 8; ++j) ya[j] = in[threadIdx.x + 256 * (4 + j)];
    if (MODE == 0) {
        float acc[4][8];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float d;
                    d = xa[i].x - ya[j].x; acc[i][j] = fmaf(d, d, acc[i][j]);
                    d = xa[i].y - ya[j].y; acc[i][j] = fmaf(d, d, acc[i][j]);
                    d = xa[i].z - ya[j].z; acc[i][j] = fmaf(d, d, acc[i][j]);
                    d = xa[i].w - ya[j].w; acc[i][j] = fmaf(d, d, acc[i][j]);
                }
            for (int i = 0; i < 4; ++i) xa[i].x += 1e-7f;                      // keep the loop body live
        }
        float s = 0.f;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j];
        out[blockIdx.x * 256 + threadIdx.x] = s;
    } else {
        u64 acc[4][8];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) acc[i][j] = 0ull;
        u64 xl[4], xh[4], yl[8], yh[8];
        for (int i = 0; i < 4; ++i) { xl[i] = *(u64*)&xa[i].x; xh[i] = *(u64*)&xa[i].z; }
        for (int j = 0; j < 8; ++j) { yl[j] = *(u64*)&ya[j].x; yh[j] = *(u64*)&ya[j].z; }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    u64 d = sub2(xl[i], yl[j]); acc[i][j] = fma2(d, d, acc[i][j]);
                    d = sub2(xh[i], yh[j]); acc[i][j] = fma2(d, d, acc[i][j]);
                }
            for (int i = 0; i < 4; ++i) xl[i] += 1ull;
        }
        float s = 0.f;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) { float2 f = *(float2*)&acc[i][j]; s += f.x + f.y; }
        out[blockIdx.x * 256 + threadIdx.x] = s;
    }
}

int main() {
    float4* in; float* out;
    cudaMalloc(&in, 256 * 12 * sizeof(float4)); cudaMemset(in, 0, 256 * 12 * sizeof(float4));
    cudaMalloc(&out, 148 * 2 * 256 * sizeof(float));
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) probe<0><<<148 * 2, 256>>>(out, in, iters); else probe<1><<<148 * 2, 256>>>(out, in, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double pairs = 148.0 * 2 * 256 * (double)iters * 128;           // element pairs (one sub + one fma each)
        printf("%s: %.3f ms, %.1f G element-pairs/s, %.2f element-pairs / clk / SM at 1.9 GHz (%s)\n", mode ? "packed" : "scalar", ms,
               pairs / ms * 1e-6, pairs / (ms * 1e-3) / 148 / 1.9e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
