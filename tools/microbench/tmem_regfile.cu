// tmem_regfile.cu — can TMEM serve as per-thread storage for frame-independent constants and parked
// values (a register-file extension) without touching the shared-memory pipe?
// Each warp owns 32 TMEM lanes (its quarter, warp id % 4) x a column range; a thread stores N 32-bit
// values to its lane with tcgen05.st.32x32b and reads them back with tcgen05.ld.32x32b.
// Measures: correctness, cycles per x32 load / store with 8 and 12 warps per SM doing it at once, and
// the same with a shared-memory-bound loop running in the other warps (does TMEM traffic share that pipe?).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_regfile tmem_regfile.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define LD32(r, addr)                                                                                         \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                    \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, " \
                 "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                        \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),        \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),  \
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
                 : "r"(addr))
#define ST32(r, addr)                                                                                         \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "                                             \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, " \
                 "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"                               \
                 :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),             \
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),       \
                    "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),     \
                    "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]),     \
                    "r"(addr) : "memory")

// mode 0: every warp loads; mode 1: only the first `tm_warps` warps load, the others hammer shared memory
__global__ void __launch_bounds__(384, 1)
probe(unsigned* bad, long long* cyc, int iters, int tm_warps, int do_store) {
    __shared__ unsigned tbase;
    __shared__ __align__(16) float buf[8192];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) buf[i] = (float)i;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(&tbase)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned base = tbase;
    // warp w: lanes 32 (w % 4) .., columns 128 (w / 4) .. (three warps share a lane quarter at 12 warps)
    const unsigned addr = base + ((unsigned)(32 * (warp & 3)) << 16) + 128u * (warp >> 2);
    unsigned r[32], s[32];
    unsigned errors = 0;
    long long t0 = 0, t1 = 0;
    float acc = 0.f;
    if (warp < tm_warps) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0x9e3779b9u * (blockIdx.x * 1024 + threadIdx.x) + 0x85ebca6bu * j;
        ST32(r, addr);
#pragma unroll
        for (int j = 0; j < 32; ++j) s[j] = r[j] ^ 0xdeadbeefu;
        ST32(s, addr + 32);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        __syncwarp();
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            unsigned a[32], b[32];
            LD32(a, addr);
            LD32(b, addr + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) errors += (a[j] != r[j]) + (b[j] != (r[j] ^ 0xdeadbeefu));
            if (do_store) {       // park / restore cycle: write the second half back, shifted
                ST32(b, addr + 64);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
        }
        t1 = clock64();
    } else {
        // shared-memory-bound filler: conflict-free 64-bit loads, the pattern of the FFT exchanges
        t0 = clock64();
        const float2* b2 = reinterpret_cast<const float2*>(buf);
        for (int it = 0; it < iters * 8; ++it) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float2 v = b2[(lane + 32 * j + 7 * it) & 4095];
                acc += v.x + v.y;
            }
        }
        t1 = clock64();
    }
    if (lane == 0) cyc[blockIdx.x * nw + warp] = t1 - t0;
    if (errors || acc == 12345.678f) atomicAdd(bad, errors + 1);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
}

int main() {
    unsigned* bad; long long* cyc;
    cudaMalloc(&bad, 4); cudaMalloc(&cyc, 148 * 12 * 8);
    long long h[148 * 12];
    const int iters = 20000;
    struct { int threads, tm_warps, do_store; const char* what; } cfg[] = {
        {256, 8, 0, "8 warps, all load 2 x x32 per iteration"},
        {384, 12, 0, "12 warps, all load"},
        {256, 8, 1, "8 warps, load 2 x x32 + store x32 per iteration"},
        {256, 4, 0, "4 warps load, 4 warps run the shared-memory loop"},
        {256, 0, 0, "8 warps run the shared-memory loop (reference for the line above)"},
    };
    for (auto& c : cfg) {
        cudaMemset(bad, 0, 4);
        probe<<<148, c.threads>>>(bad, cyc, iters, c.tm_warps, c.do_store);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        unsigned hb; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
        const int nw = c.threads / 32;
        cudaMemcpy(h, cyc, 148 * nw * 8, cudaMemcpyDeviceToHost);
        double tm = 0, sm = 0; int ntm = 0, nsm = 0;
        for (int b = 0; b < 148; ++b)
            for (int w = 0; w < nw; ++w) {
                if (w < c.tm_warps) { tm += (double)h[b * nw + w]; ++ntm; } else { sm += (double)h[b * nw + w]; ++nsm; }
            }
        printf("%-72s mismatches %u |", c.what, hb);
        if (ntm) printf(" TMEM warps: %.1f cycles / iteration (%.1f B/clk/SM read)", tm / ntm / iters, c.tm_warps * 2 * 4096.0 / (tm / ntm / iters));
        if (nsm) printf(" | smem warps: %.1f cycles / 16 LDS.64 (%.1f B/clk/SM)", sm / nsm / (iters * 8), (nw - c.tm_warps) * 16 * 256.0 / (sm / nsm / (iters * 8)));
        printf("\n");
    }
    return 0;
}
