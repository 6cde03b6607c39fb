// dft16_tcgen05.cu — the ncu-gated experiment north_star asks for: ONE radix-16 pass of
// the 4096-point frame (256 butterflies) as a GEMM on the 5th-generation tensor cores, against the same
// pass on the FP32 pipe.
//
//   D[256 x 32] = A[256 x 32] * B^T,  row of A = (Re z_0..z_15, Im z_0..z_15) of one butterfly,
//   B[32 x 32] = the real form of the 16-point DFT matrix (scaled by 1/4: unitary, so the pass can be
//   iterated on its own output), tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = 32, K = 8 per instruction.
// fp32-equivalent accuracy needs the 3 x TF32 split  A B = Ah Bh + Al Bh + Ah Bl  (hi = top 10 mantissa bits):
// 2 M-tiles x 3 terms x 4 K-steps = 24 MMAs per frame-pass, operands in shared memory (K-major, no swizzle:
// 8 x 16-byte core matrices), accumulators in TMEM, read back with tcgen05.ld, split again and written to
// shared memory as the next pass's operands (that is the data flow the real kernel would need).
// The FP32 arm is the radix-16 butterfly of stft_r16.cuh (packed f32x2) on the same 256 butterflies per pass.
// Prints accuracy (1 x and 3 x TF32 against float64) and cycles per frame-pass per SM for both arms.
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -I em-spec_b200/csrc -I include \
//        -o tools/microbench/dft16_tcgen05 tools/microbench/dft16_tcgen05.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "stft_r16.cuh"

using namespace ems::r16;

constexpr int kRows = 256, kK = 32, kN = 32;
constexpr int kABytes = 2 * 16384;                // two M = 128 tiles, 8 K-chunks of 2048 bytes each
constexpr int kBBytes = 4096;
constexpr int kSmemT = 2 * kABytes + 2 * kBBytes + 64;

__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
    // K-major, no swizzle: (8 rows x 16 bytes) core matrices; LBO = bytes between the two 16-byte K chunks of an
    // MMA, SBO = bytes between 8-row groups; version 1 (Blackwell)
    return (unsigned long long)((saddr >> 4) & 0x3fffu) | ((unsigned long long)((lbo >> 4) & 0x3fffu) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// c = F32, a = b = TF32, both K-major, N = 32, M = 128
constexpr unsigned kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kN >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ int a_off(int row, int k) {       // byte offset of A[row][k] inside one operand buffer
    const int tile = row >> 7, r = row & 127;
    return tile * 16384 + (k >> 2) * 2048 + (r >> 3) * 128 + (r & 7) * 16 + (k & 3) * 4;
}
__device__ __forceinline__ int b_off(int n, int k) { return (k >> 2) * 512 + (n >> 3) * 128 + (n & 7) * 16 + (k & 3) * 4; }
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

#define LD32F(r, addr)                                                                                        \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                    \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, " \
                 "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                        \
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),        \
                   "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]),  \
                   "=f"(r[16]), "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), \
                   "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31]) \
                 : "r"(addr))

// terms: 1 = Ah Bh only (one TF32 pass), 3 = the fp32-equivalent split.  128 threads, one CTA per SM.
__global__ void __launch_bounds__(128, 1)
pass_tensor(const float* __restrict__ a_in, const float* __restrict__ b_in, float* __restrict__ d_out, int iters,
            int terms, long long* cyc, int* err) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned char* Ah = sm;
    unsigned char* Al = sm + kABytes;
    unsigned char* Bh = sm + 2 * kABytes;
    unsigned char* Bl = Bh + kBBytes;
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(Bl + kBBytes);
    unsigned* tslot = reinterpret_cast<unsigned*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < kRows * kK; e += 128) {
        const int row = e / kK, k = e % kK;
        const float v = a_in[e], h = tf32_hi(v);
        *reinterpret_cast<float*>(Ah + a_off(row, k)) = h;
        *reinterpret_cast<float*>(Al + a_off(row, k)) = tf32_hi(v - h);
    }
    for (int e = tid; e < kN * kK; e += 128) {
        const int n = e / kK, k = e % kK;
        const float v = b_in[e], h = tf32_hi(v);
        *reinterpret_cast<float*>(Bh + b_off(n, k)) = h;
        *reinterpret_cast<float*>(Bl + b_off(n, k)) = tf32_hi(v - h);
    }
    const unsigned mbar_s = (unsigned)__cvta_generic_to_shared(mbar);
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s));
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((unsigned)__cvta_generic_to_shared(tslot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned tbase = *tslot;
    const unsigned ah_s = (unsigned)__cvta_generic_to_shared(Ah), al_s = (unsigned)__cvta_generic_to_shared(Al);
    const unsigned bh_s = (unsigned)__cvta_generic_to_shared(Bh), bl_s = (unsigned)__cvta_generic_to_shared(Bl);

    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (tid == 0) {
            for (int tile = 0; tile < 2; ++tile) {
                for (int term = 0; term < terms; ++term) {
                    const unsigned a_s = (term == 1 ? al_s : ah_s) + tile * 16384;
                    const unsigned b_s = term == 2 ? bl_s : bh_s;
                    for (int ks = 0; ks < 4; ++ks) {
                        const unsigned long long ad = umma_desc(a_s + ks * 4096, 2048, 128);
                        const unsigned long long bd = umma_desc(b_s + ks * 1024, 512, 128);
                        const unsigned acc = (term | ks) ? 1u : 0u;
                        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0;\n"
                                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                                     ::"r"(tbase + 32u * tile), "l"(ad), "l"(bd), "r"(kIdesc), "r"(acc) : "memory");
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_s) : "memory");
        }
        {   // everybody waits for the MMAs of this pass (bounded: a descriptor mistake must not hang the GPU)
            unsigned ok = 0;
            int spins = 0;
            while (!ok) {
                asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                             : "=r"(ok) : "r"(mbar_s), "r"((unsigned)(it & 1)) : "memory");
                if (!ok && ++spins > 2000000) { if (tid == 0) *err = 1; break; }
            }
        }
        asm volatile("tcgen05.fence::after_thread_sync;");
        // epilogue: D row (32 warp + lane) of both tiles -> registers -> split -> the operand buffers of the next pass
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
            float r[32];
            LD32F(r, tbase + ((unsigned)(32 * warp) << 16) + 32u * tile);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int row = tile * 128 + 32 * warp + lane;
            if (d_out && it == iters - 1) {
#pragma unroll
                for (int n = 0; n < 32; ++n) d_out[row * kN + n] = r[n];
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 h, l;
                h.x = tf32_hi(r[4 * c]); h.y = tf32_hi(r[4 * c + 1]); h.z = tf32_hi(r[4 * c + 2]); h.w = tf32_hi(r[4 * c + 3]);
                l.x = tf32_hi(r[4 * c] - h.x); l.y = tf32_hi(r[4 * c + 1] - h.y); l.z = tf32_hi(r[4 * c + 2] - h.z); l.w = tf32_hi(r[4 * c + 3] - h.w);
                *reinterpret_cast<float4*>(Ah + a_off(row, 4 * c)) = h;
                if (terms > 1) *reinterpret_cast<float4*>(Al + a_off(row, 4 * c)) = l;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    const long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tbase));
}

// The same pass on the FP32 pipe: 128 threads, two radix-16 butterflies each, in place in shared memory
// (element j of butterfly b at 257 j + b: conflict-free), scaled by 1/4 like the tensor arm.
__global__ void __launch_bounds__(128, 1)
pass_fp32(const float* __restrict__ a_in, float* __restrict__ d_out, int iters, long long* cyc) {
    extern __shared__ __align__(16) unsigned char sm[];
    float2* Z = reinterpret_cast<float2*>(sm);
    const int tid = threadIdx.x;
    for (int e = tid; e < kRows * 16; e += 128) {
        const int b = e / 16, j = e % 16;
        Z[257 * j + b] = make_float2(a_in[b * kK + j], a_in[b * kK + 16 + j]);
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float2* z = Z + tid + 128 * u;
            float2 v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = z[257 * j];
            dft16(v);
#pragma unroll
            for (int i = 0; i < 16; ++i) z[257 * i] = mul2(v[o16(i)], make_float2(0.25f, 0.25f));
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    if (d_out)
        for (int e = tid; e < kRows * 16; e += 128) {
            const int b = e / 16, j = e % 16;
            d_out[b * kN + j] = Z[257 * j + b].x;
            d_out[b * kN + 16 + j] = Z[257 * j + b].y;
        }
}

int main() {
    std::vector<float> a(kRows * kK), b(kN * kK);
    std::vector<double> ref(kRows * kN);
    srand(7);
    for (auto& v : a) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    const double pi = 3.14159265358979323846;
    for (int n = 0; n < kN; ++n)
        for (int k = 0; k < kK; ++k) {
            const int nn = n & 15, kk = k & 15;
            const double c = cos(2 * pi * kk * nn / 16.0), s = sin(2 * pi * kk * nn / 16.0);
            // Re_n = sum re c + im s ; Im_n = sum -re s + im c
            const double v = n < 16 ? (k < 16 ? c : s) : (k < 16 ? -s : c);
            b[n * kK + k] = (float)(0.25 * v);
        }
    for (int m = 0; m < kRows; ++m)
        for (int n = 0; n < kN; ++n) {
            double acc = 0;
            for (int k = 0; k < kK; ++k) acc += (double)a[m * kK + k] * (double)b[n * kK + k];
            ref[m * kN + n] = acc;
        }
    float *da, *db, *dd; long long* dc; int* de;
    cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dd, kRows * kN * 4);
    cudaMalloc(&dc, 148 * 8); cudaMalloc(&de, 4);
    cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(pass_tensor, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemT);
    cudaFuncSetAttribute(pass_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 257 * 8);
    std::vector<float> d(kRows * kN);
    auto accuracy = [&](const char* what) {
        cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
        double num = 0, den = 0, mx = 0;
        for (size_t i = 0; i < d.size(); ++i) {
            num += (d[i] - ref[i]) * (d[i] - ref[i]); den += ref[i] * ref[i];
            mx = fmax(mx, fabs(d[i] - ref[i]));
        }
        printf("%-34s rel-L2 error %.3e, max abs error %.3e\n", what, sqrt(num / den), mx);
    };
    auto check = [&](const char* what) {
        cudaError_t e = cudaDeviceSynchronize();
        int herr = 0;
        cudaMemcpy(&herr, de, 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess || herr) { printf("%s: %s%s\n", what, cudaGetErrorString(e), herr ? " (mbarrier wait timed out)" : ""); exit(1); }
    };
    cudaMemset(de, 0, 4);
    pass_tensor<<<1, 128, kSmemT>>>(da, db, dd, 1, 1, dc, de); check("tensor 1x"); accuracy("one pass, 1 x TF32 (tcgen05):");
    pass_tensor<<<1, 128, kSmemT>>>(da, db, dd, 1, 3, dc, de); check("tensor 3x"); accuracy("one pass, 3 x TF32 (tcgen05):");
    pass_fp32<<<1, 128, 16 * 257 * 8>>>(da, dd, 1, dc); check("fp32"); accuracy("one pass, FP32 pipe (f32x2):");
    const int iters = 20000;
    long long h[148];
    auto timing = [&](const char* what) {
        cudaMemcpy(h, dc, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0;
        for (long long v : h) s += (double)v;
        printf("%-52s %8.1f cycles per frame-pass per SM\n", what, s / 148 / iters);
    };
    pass_tensor<<<148, 128, kSmemT>>>(da, db, nullptr, iters, 3, dc, de); check("tensor 3x timing");
    timing("tcgen05 3 x TF32: 24 MMAs + TMEM read-back + re-split:");
    pass_tensor<<<148, 128, kSmemT>>>(da, db, nullptr, iters, 1, dc, de); check("tensor 1x timing");
    timing("tcgen05 1 x TF32:  8 MMAs + TMEM read-back:");
    pass_fp32<<<148, 128, 16 * 257 * 8>>>(da, nullptr, iters, dc); check("fp32 timing");
    timing("FP32 pipe, 128 threads (one worker of the kernel):");
    return 0;
}
