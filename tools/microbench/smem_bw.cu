// Microbenchmark: shared-memory LDS.64 / STS.64 / mixed throughput per SM (conflict-free),
// to check the real sustained rate behind ncu's "wavefronts % of peak".
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smem_bw smem_bw.cu && ./smem_bw
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 lds64(const float2* p) {
    float2 v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(float2* p, float2 v) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float lds32(const float* p) {
    float v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}

template <int MODE>   // 0: loads, 1: stores, 2: load+store alternating, 3: 32-bit loads
__global__ void k(float* out, int iters) {
    extern __shared__ float2 sm[];
    const int t = threadIdx.x;
    for (int i = t; i < 8192; i += blockDim.x) sm[i] = make_float2(i, -i);
    __syncthreads();
    float2 acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = make_float2(t, j);
    float2* base = sm + t;       // consecutive lanes -> consecutive 8-byte slots: conflict-free
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { float2 v = lds64(base + 1024 * (j & 3) + ((it + j) & 1) * 4096); acc[j].x += v.x; acc[j].y += v.y; }
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) sts64(base + 1024 * (j & 3) + ((it + j) & 1) * 4096, acc[j]);
        } else if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                float2 v = lds64(base + 1024 * (j & 3));
                acc[j].x += v.x; acc[j].y += v.y;
                sts64(base + 1024 * ((j + 1) & 3) + 4096, acc[j + 1]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { float v = lds32((float*)sm + t + 1024 * j); acc[j].x += v; }
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[j].x + acc[j].y;
    out[blockIdx.x * blockDim.x + t] = s;
}

template <int MODE>
void run(const char* name, double bytes_per_access) {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int iters = 20000;
    for (int threads : {128, 384, 1024}) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<MODE><<<148, threads, 65536>>>(out, 100);
        cudaEventRecord(e0);
        k<MODE><<<148, threads, 65536>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = 8.0 * iters * threads * bytes_per_access * 148;
        printf("%-12s %4d thr: %.3f ms  %.1f B/clk/SM @1.9GHz\n", name, threads, ms, bytes / (ms * 1e-3) / 148 / 1.9e9);
    }
    cudaFree(out);
}

int main() {
    run<0>("LDS.64", 8); run<1>("STS.64", 8); run<2>("LDS+STS.64", 8); run<3>("LDS.32", 4);
    return 0;
}
