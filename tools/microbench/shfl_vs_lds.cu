// Microbenchmark: do warp shuffles share the shared-memory data pipe?  Times a loop of
// LDS.64, a loop of SHFL.32, and both interleaved, all conflict-free, 12 warps per SM.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o shfl_vs_lds shfl_vs_lds.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 1: LDS, 2: SHFL, 3: both
__global__ void k(float* out, int iters) {
    __shared__ float2 sm[4096];
    const int t = threadIdx.x;
    for (int i = t; i < 4096; i += blockDim.x) sm[i] = make_float2(i, -i);
    __syncthreads();
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = t + j;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm + t);
    for (int it = 0; it < iters; ++it) {
        if (MODE & 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float x, y;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(base + 8u * (unsigned)(((it * 7 + j * 3) & 7) * 384)) : "memory");
                acc[j] += x + y;
            }
        }
        if (MODE & 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[(j + 1) & 7], 1 + (j & 3));
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + t] = s;
}

template <int MODE>
float run(const char* name) {
    float* out; cudaMalloc(&out, 148 * 384 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<MODE><<<148, 384>>>(out, 100);
    cudaEventRecord(e0);
    k<MODE><<<148, 384>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = 8.0 * iters * 12;      // per SM, per kind
    printf("%-6s %.3f ms  -> %.2f cycles per warp-instruction of each kind per SM @1.9 GHz\n", name, ms,
           ms * 1e-3 * 1.9e9 / warp_instr);
    cudaFree(out);
    return ms;
}

int main() {
    const float a = run<1>("LDS.64"), b = run<2>("SHFL"), c = run<3>("both");
    printf("both / (LDS + SHFL) = %.2f,  both / max = %.2f\n", c / (a + b), c / (a > b ? a : b));
    return 0;
}
