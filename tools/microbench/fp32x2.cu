// Microbenchmark: issue/throughput of FADD, FFMA, add.f32x2, fma.f32x2 on sm_100a at a few
// occupancies (warps per SM) — decides whether packed fp32x2 helps the FFT butterflies.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32x2 fp32x2.cu && ./fp32x2
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
    unsigned long long av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = x[i] + b;
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
        } else if (MODE == 2) {   // packed add: 8 instr do the work of 16 FADD
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long v;
                asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[i]), "f"(x[i + 1]));
                asm("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(bv));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(v));
            }
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long v;
                asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[i]), "f"(x[i + 1]));
                asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(av), "l"(bv));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(v));
            }
        } else if (MODE == 4) {   // FADD with two register operands (x[i] += x[i^1]-like, no immediates)
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = x[i] + x[(i + 5) & 15] * 0.0f + b;
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int flops_per_instr_x16) {
    float* out;
    cudaMalloc(&out, 148 * 1024 * sizeof(float) * 4);
    int iters = 20000;
    for (int threads : {128, 256, 384, 512, 1024}) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<MODE><<<148, threads>>>(out, 100, 1.0001f, 0.5f);
        cudaEventRecord(e0);
        k<MODE><<<148, threads>>>(out, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double lane_ops = 16.0 * iters * threads * 148;   // scalar-equivalent ops
        printf("%-10s warps/SM %2d: %.3f ms  %.1f G lane-ops/s  (%.1f lane-ops/clk/SM @1.9GHz)\n", name,
               threads / 32, ms, lane_ops / ms / 1e6, lane_ops / (ms * 1e-3) / 148 / 1.9e9);
    }
    cudaFree(out);
}

int main() {
    run<0>("FADD", 1);
    run<1>("FFMA", 2);
    run<2>("FADD2", 2);
    run<3>("FFMA2", 4);
    return 0;
}
