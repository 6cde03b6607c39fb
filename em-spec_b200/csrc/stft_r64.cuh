// stft_r64.cuh — single-exchange variant of the fused frame gather + STFT + reassignment kernel for
// n_fft = 4096 ("next round" item 1a).
//
// Same mathematics as stft_r16.cuh (one complex FFT Z = FFT_N(x + j x th'), untangle, Hann / dh
// stencils, Auger-Flandrin epilogue) but Z is computed as 64 x 64 in TWO passes instead of
// 16 x 16 x 16 in three, so a frame crosses shared memory once between passes, not twice:
//   * a worker is 64 threads (2 warps), 4 workers per CTA (256 threads, <= 255 registers), one CTA per SM;
//   * pass 1: thread n2 takes its 64 samples n = n2 + 64 n1 straight from the tile (th' from a
//     shared-memory table, 128-bit loads), does the radix-64 butterfly over n1 in registers
//     (4 x 16: 16 DFT-4, W_64 twiddles, 4 DFT-16), multiplies by W_N^{n2 k1} (14 exact twiddles in
//     registers, the other 49 one product deep) and stores row n2 of the 64 x 65 Z buffer;
//   * pass 2: thread k1 reads column k1, does the radix-64 butterfly over n2: Z[k1 + 64 k2];
//   * untangle: Z[N - k] of residue k1 lives in the thread of residue 64 - k1; the two sit 16 lanes
//     apart in one warp and swap the upper halves of their outputs with 64 shuffles;
//   * 2 X goes back into the thread's own column of the Z buffer (rows 0..31: nobody else reads
//     that column), so there is no separate X buffer; the epilogue reads its two neighbours from the
//     adjacent columns.  3 worker barriers per frame (2 warps each).
// Shared-memory wavefronts per frame: 128 tile + 128 th' + 256 + 256 exchange + 128 shuffles +
// 128 X + 256 neighbours = 1,280 by design, 1,424 measured (ncu), against 1,931 of the three-pass kernel.
//
// RESULT (B200, profiles/r02_ncu_r64_v2.txt, r02_r64_*_times.txt): parity green, 5,018 warp instructions per
// frame (three-pass: 5,560), but 64 values per thread means 255 registers and 8 warps per SM: the kernel
// is latency-bound (issue slots 39 %, FMA pipe 47 %, l1tex 49 %) and runs at 94-99 M frames/s against 103-105 M
// for the three-pass kernel.  More warps are not to be had: five workers get the register budget of six (168,
// registers are handed out per four warps) and spill, and tensor memory as a register-file extension
// (EMS_R64_TMEM / EMS_R64_PARK: th' and 2 X_th' parked in TMEM with tcgen05.st / tcgen05.ld) works but its
// ~300-cycle loads cost more than the shared-memory table they replace.  Selected only with
// EMS_KERNEL_VARIANT=64 (A/B runs, tools/variant_check.py); the product path stays stft_r16.cuh.
#pragma once
#include "stft_r16.cuh"

// experiment switches (tools/variant_check.py builds the alternatives side by side)
#ifndef EMS_R64_PREFETCH
#define EMS_R64_PREFETCH 0      // store-mode epilogue: neighbour loads of the next four bins before the vote
#endif
#ifndef EMS_R64_BASES
#define EMS_R64_BASES 14        // exact inter-pass twiddles kept in registers: 14 (products one deep) or 6 (three deep)
#endif
#ifndef EMS_R64_LOOP
#define EMS_R64_LOOP 0          // one copy of the radix-64 butterfly code, looped over the two passes
#endif
#ifndef EMS_R64_WORKERS
#define EMS_R64_WORKERS 4       // frames in flight per CTA (64 threads each): 4 at <= 255 registers, 5 at <= 204
#endif
#ifndef EMS_R64_PARK
#define EMS_R64_PARK 0          // (with EMS_R64_TMEM) 2 X_th' of the 32 bins waits in tensor memory between untangle and epilogue
#endif
#ifndef EMS_R64_TMEM
#define EMS_R64_TMEM 0          // th' of a thread's 64 samples lives in tensor memory (tcgen05.ld) instead of a shared-memory table
#endif

namespace ems {
namespace r64 {

using namespace r16;

constexpr int kN = 4096;
constexpr int kWorkers = EMS_R64_WORKERS;
constexpr int kWT = 64;                      // threads per worker = per frame
constexpr int kThreads64 = kWorkers * kWT;
constexpr int kRow = 65;                     // Z[n2][k1] at n2 * 65 + k1: both exchanges conflict-free
constexpr int kZ = 64 * kRow;
constexpr int kSlot = 2 + kZ + 6;            // 2 X[-1] (at Zb[-2]), Z, scratch: 2 X[2049], 2 X_th'[2048]
constexpr int kThwBytes = EMS_R64_TMEM ? 0 : kN * 4;
constexpr int kFixedBytes = kThwBytes + kWorkers * kSlot * 8;
constexpr int kTileFloats = ((kMaxSmem - kFixedBytes - kSyncBytes) / 8) & ~3;
constexpr int kMaxTile = 12 * kWorkers;
static_assert(kFixedBytes % 16 == 0 && kTileFloats >= kN, "tile buffers");
__host__ __device__ constexpr int tile_frames(int hop) {
    int t = (kTileFloats - kN) / hop + 1;
    if (t > kWorkers) t -= t % kWorkers;
    return t > kMaxTile ? kMaxTile : t;
}

// cos, sin(2 pi q / 64) for compile-time q
__host__ __device__ constexpr float c64(int q) {
    constexpr float t[17] = {1.0f, 0.99518472667219693f, 0.98078528040323043f, 0.95694033573220882f,
                             0.92387953251128674f, 0.88192126434835505f, 0.83146961230254524f, 0.77301045336273699f,
                             0.70710678118654757f, 0.63439328416364549f, 0.55557023301960229f, 0.47139673682599781f,
                             0.38268343236508984f, 0.29028467725446233f, 0.19509032201612833f, 0.09801714032956077f, 0.0f};
    const int e = q & 63;
    return e <= 16 ? t[e] : e <= 32 ? -t[32 - e] : e <= 48 ? -t[e - 32] : t[64 - e];
}
__host__ __device__ constexpr float s64(int q) { return c64(q - 16); }

// v * W_64^E
template <int E>
__device__ __forceinline__ float2 mul_w64(float2 v) {
    constexpr int e = E & 63;
    if constexpr (e == 0) return v;
    else if constexpr (e == 16) return mulmj(v);
    else if constexpr (e == 32) return make_float2(-v.x, -v.y);
    else if constexpr (e == 48) return mulpj(v);
    else {
        constexpr float c = c64(e), s = s64(e);       // v (c - j s) = v c + (-j v) s
        return fma2(mulmj(v), make_float2(s, s), mul2(v, make_float2(c, c)));
    }
}

// forward DFT-16 in place on v[O .. O+15]; output i sits in v[O + o16(i)]
template <int O>
__device__ __forceinline__ void dft16_at(float2 (&v)[64]) {
#pragma unroll
    for (int j0 = 0; j0 < 4; ++j0) dft4(v[O + j0], v[O + j0 + 4], v[O + j0 + 8], v[O + j0 + 12]);
    v[O + 5] = mul_w16<1>(v[O + 5]);   v[O + 9] = mul_w16<2>(v[O + 9]);   v[O + 13] = mul_w16<3>(v[O + 13]);
    v[O + 6] = mul_w16<2>(v[O + 6]);   v[O + 10] = mul_w16<4>(v[O + 10]); v[O + 14] = mul_w16<6>(v[O + 14]);
    v[O + 7] = mul_w16<3>(v[O + 7]);   v[O + 11] = mul_w16<6>(v[O + 11]); v[O + 15] = mul_w16<9>(v[O + 15]);
#pragma unroll
    for (int i0 = 0; i0 < 4; ++i0) dft4(v[O + 4 * i0], v[O + 4 * i0 + 1], v[O + 4 * i0 + 2], v[O + 4 * i0 + 3]);
}

// forward DFT-64 in registers, input n at v[n]; output k sits in v[o64(k)].
// n = b + 16 a, k = kl + 4 kh:  W_64^{nk} = W_64^{b kl} W_16^{b kh} W_4^{a kl}
__host__ __device__ constexpr int o64(int k) { return 16 * (k & 3) + o16(k >> 2); }
__device__ __forceinline__ void dft64(float2 (&v)[64]) {
#pragma unroll
    for (int b = 0; b < 16; ++b) dft4(v[b], v[b + 16], v[b + 32], v[b + 48]);
    static_for<16>([&](auto bc) {
        constexpr int b = decltype(bc)::value;
        v[b + 16] = mul_w64<b>(v[b + 16]);
        v[b + 32] = mul_w64<2 * b>(v[b + 32]);
        v[b + 48] = mul_w64<3 * b>(v[b + 48]);
    });
    dft16_at<0>(v); dft16_at<16>(v); dft16_at<32>(v); dft16_at<48>(v);
}

// ---- tensor memory as per-thread storage: a warp owns the 32 lanes of its quarter (warp id % 4) and a
// column range; tcgen05.st / tcgen05.ld 32x32b move 16 values per thread to / from the thread's own lane.
// No shared-memory wavefronts (tools/microbench/tmem_regfile.cu).
#define EMS_TMEM_LD16(r, addr)                                                                               \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                   \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"           \
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),       \
                   "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]) \
                 : "r"(addr))
#define EMS_TMEM_LD8(r, addr)                                                                                \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"            \
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "r"(addr))
#define EMS_TMEM_ST8(r, addr)                                                                                \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};"            \
                 :: "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]), "r"(addr) : "memory")
#define EMS_TMEM_ST16(r, addr)                                                                               \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], "                                            \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"                  \
                 :: "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]),            \
                    "f"(r[8]), "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]),      \
                    "r"(addr) : "memory")

__device__ __forceinline__ float2 shfl2(float2 a, int src) {
    return make_float2(__shfl_sync(0xffffffffu, a.x, src), __shfl_sync(0xffffffffu, a.y, src));
}

// (five workers: registers are handed out per four warps, so 320 threads get the budget of 384 — 168
// registers; a 200-register build, __maxnreg__(200), fails to launch: "too many resources requested")
template <int MODE>
__global__ void __launch_bounds__(kThreads64, 1)
stft_reassign_r64(const StftArgs a_in, const int tile_T) {
    constexpr int N = kN, B = N / 2 + 1;
    StftArgs a = a_in;
    if (!stream_decode(a)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* thwT = reinterpret_cast<float4*>(smem_raw);                        // [16][64]: th'[t + 64 (4 q + r)] at [q][t].r
    float2* wbuf = reinterpret_cast<float2*>(smem_raw + kThwBytes);
    float* tile0 = reinterpret_cast<float*>(wbuf + kWorkers * kSlot);          // 2 x kTileFloats

    const int tid = threadIdx.x;
    const int w = tid / kWT;                   // worker
    const int tl = tid - w * kWT;              // thread of the worker = n2 in pass 1
    const int lane = tid & 31, wv = tl >> 5;
    // pass 2 / epilogue role: residue k1.  Residues k1 and 64 - k1 sit 16 lanes apart in one warp;
    // every half-warp holds 16 residues that differ mod 16 (conflict-free 64-bit column accesses).
    const int j16 = lane & 15;
    const int k1 = wv == 0 ? (lane < 16 ? lane : (j16 == 0 ? 32 : 64 - j16))
                           : (lane < 16 ? 16 + lane : 48 - j16);
    const bool self0 = k1 == 0;
    const int pl = (wv == 0 && j16 == 0) ? lane : (lane ^ 16);          // lane that holds residue 64 - k1
    float2* Zb = wbuf + w * kSlot + 2;
    float2* Sc = Zb + kZ;                      // [0] 2 X[2049], [1] 2 X_th'[2048]

#if EMS_R64_TMEM
    // th'[tl + 64 n1], n1 = 0..63, into 64 columns of this thread's TMEM lane
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(tile0 + 2 * kTileFloats) + 40;      // spare word of the sync block
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned tmem_base = *tmem_slot;
    const unsigned tmem_thw = tmem_base + ((unsigned)(32 * ((tid >> 5) & 3)) << 16) + 128u * (unsigned)(tid >> 7);
    const unsigned tmem_T = tmem_thw + 64;       // 2 X_th' of the thread's 32 bins, parked between untangle and epilogue
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float t16[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) t16[j] = __ldg(&a.thw[tl + 64 * (16 * q + j)]);
        EMS_TMEM_ST16(t16, tmem_thw + 16 * q);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#else
    for (int e = tid; e < N; e += kThreads64) {
        const int t = e & 63, n1 = e >> 6;
        reinterpret_cast<float*>(thwT)[((n1 >> 2) * 64 + t) * 4 + (n1 & 3)] = __ldg(&a.thw[e]);
    }
#endif
    // W_N^{n2 j}, W_N^{8 n2 j}, j = 1..7: exact, frame-independent, in registers
    float2 wlo[8], whi[8];
#pragma unroll
    for (int j = 1; j < 8; ++j) {
        if (EMS_R64_BASES == 14 || j == 1 || j == 2 || j == 4) {
            wlo[j] = __ldg(&a.tw[(tl * j) & (N - 1)]);
            whi[j] = __ldg(&a.tw[(tl * 8 * j) & (N - 1)]);
        }
    }

    const long long per_ch = a.f_end - a.f_begin;
    const long long tiles_per_ch = (per_ch + tile_T - 1) / tile_T;
    const long long n_tiles = tiles_per_ch * a.channels;

    // tile hand-over: as in stft_reassign_r16 (mbarrier full[b], release counter done[b])
    unsigned long long* full = reinterpret_cast<unsigned long long*>(tile0 + 2 * kTileFloats);   // [2]
    unsigned* done = reinterpret_cast<unsigned*>(full + 2);                                     // [2]
    int* refill = reinterpret_cast<int*>(done + 2);                                             // [kWorkers]
    const unsigned full_sm = (unsigned)__cvta_generic_to_shared(full);
    auto tile_geom = [&](long long tl_, int& ch, long long& f0, int& nf) {
        ch = (int)(tl_ / tiles_per_ch);
        f0 = a.f_begin + (tl_ - (long long)ch * tiles_per_ch) * tile_T;
        nf = (int)min((long long)tile_T, a.f_end - f0);
    };
    auto copy_tile = [&](long long tl_, float* dst, int t0, int nth) {
        int ch, nf; long long f0;
        tile_geom(tl_, ch, f0, nf);
        const int n_samp = (nf - 1) * a.hop + N;
        const float* src = a.pcm + (long long)ch * a.S + f0 * a.hop + a.samp_off;
        const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst);
        for (int s = t0; s < n_samp; s += nth)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + 4u * s), "l"(src + s) : "memory");
    };
    const long long tile_step = gridDim.x;
    if ((long long)blockIdx.x < n_tiles) copy_tile(blockIdx.x, tile0, tid, kThreads64);
    if ((long long)blockIdx.x + tile_step < n_tiles) copy_tile(blockIdx.x + tile_step, tile0 + kTileFloats, tid, kThreads64);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full_sm), "r"(kWT));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full_sm + 8), "r"(kWT));
        done[0] = 0; done[1] = 0;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    auto release_tile = [&](int b) {
        __threadfence_block();
        const unsigned old = atomicAdd(&done[b], 1u);
        const bool last = old == kWorkers - 1;
        if (last) done[b] = 0;
        __threadfence_block();
        refill[w] = last ? 1 : 0;
    };
    auto refill_tile = [&](int b, long long ti_next) {
        const long long tl2 = blockIdx.x + ti_next * tile_step;
        if (tl2 >= n_tiles) return;
        int ch2, nf2; long long f02;
        tile_geom(tl2, ch2, f02, nf2);
        const unsigned bytes = 4u * (unsigned)((nf2 - 1) * a.hop + N);
        const float* src = a.pcm + (long long)ch2 * a.S + f02 * a.hop + a.samp_off;
        if ((((unsigned long long)src | (unsigned long long)(unsigned)a.hop * 4ull) & 15ull) == 0) {
            // one TMA bulk copy by a single thread (a 64-thread worker is too few threads for the 4-byte loop)
            const unsigned bar = full_sm + 8u * b;
            if (tl == 0) {
                const unsigned d0 = (unsigned)__cvta_generic_to_shared(tile0 + b * kTileFloats);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(d0), "l"(src), "r"(bytes), "r"(bar) : "memory");
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
            }
            return;
        }
        copy_tile(tl2, tile0 + b * kTileFloats, tl, kWT);
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full_sm + 8u * b) : "memory");
    };

    const float gate_p = gate_power<N>(a);
    const float k1f = (float)k1;
    // neighbour columns of this residue: X[k - 1] at Zb[65 i + offM], X[k + 1] at Zb[65 i + offP]
    // (k1 = 0: the last column of the previous row, row -1 being the mirror slot Zb[-2]; k1 = 63: the next row)
    const float2* Xm = Zb + (k1 > 0 ? k1 - 1 : -2);
    const float2* Xp = Zb + (k1 < 63 ? k1 + 1 : kRow);

    for (long long ti = 0;; ++ti) {
        const long long tl_ = blockIdx.x + ti * tile_step;
        if (tl_ >= n_tiles) break;
        const int buf = (int)(ti & 1);
        if (ti >= 2) {
            const unsigned parity = (unsigned)(((ti >> 1) - 1) & 1);
            unsigned ok = 0;
            while (!ok)
                asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                             : "=r"(ok) : "r"(full_sm + 8u * buf), "r"(parity) : "memory");
        }
        int ch, nf; long long f0;
        tile_geom(tl_, ch, f0, nf);
        const float* tile = tile0 + buf * kTileFloats;
        if (w >= nf) {
            if (tl == 0) release_tile(buf);
            worker_bar<kWT>(w);
            if (refill[w]) refill_tile(buf, ti + 2);
            worker_bar<kWT>(w);
            continue;
        }

        for (int fi = w; fi < nf; fi += kWorkers) {
            const float* xs = tile + fi * a.hop;
            const long long f = f0 + fi;
            const bool last_frame = fi + kWorkers >= nf;

            float2 v[64];
            auto load_pass1 = [&]() {
                // ================= pass 1: radix-64 over n1 of z[n] = x[n] (1 + j th'[n]), n = tl + 64 n1
#if EMS_R64_TMEM
                float ta_[16], tb_[16];
                EMS_TMEM_LD16(ta_, tmem_thw);
#pragma unroll
                for (int n1 = 0; n1 < 64; ++n1) v[n1].x = xs[tl + 64 * n1];
                static_for<4>([&](auto qc) {
                    constexpr int q = decltype(qc)::value;
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if constexpr (q + 1 < 4) {            // the next 16 fly while these are consumed
                        if constexpr (q & 1) EMS_TMEM_LD16(ta_, tmem_thw + 16 * (q + 1));
                        else EMS_TMEM_LD16(tb_, tmem_thw + 16 * (q + 1));
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[16 * q + j].y = v[16 * q + j].x * ((q & 1) ? tb_[j] : ta_[j]);
                });
                return;
#endif
                // (quads in the order the first DFT-4 stage consumes them: n1 = b, b + 16, b + 32, b + 48)
#pragma unroll
                for (int qq = 0; qq < 16; ++qq) {
                    const int q = 4 * (qq & 3) + (qq >> 2);
                    const float4 t4 = thwT[q * 64 + tl];
                    const float x0 = xs[tl + 64 * (4 * q)], x1 = xs[tl + 64 * (4 * q + 1)],
                                x2 = xs[tl + 64 * (4 * q + 2)], x3 = xs[tl + 64 * (4 * q + 3)];
                    v[4 * q] = make_float2(x0, x0 * t4.x);
                    v[4 * q + 1] = make_float2(x1, x1 * t4.y);
                    v[4 * q + 2] = make_float2(x2, x2 * t4.z);
                    v[4 * q + 3] = make_float2(x3, x3 * t4.w);
                }
            };
            auto store_pass1 = [&]() {
                // the Z buffer still holds the previous frame's X until every thread of the worker has
                // left its epilogue
                worker_bar<kWT>(w);
                if (last_frame && tl == 0) release_tile(buf);       // every sample of the tile has been read
                float2 lo[8], hi[8];
#pragma unroll
                for (int j = 1; j < 8; ++j) { lo[j] = wlo[j]; hi[j] = whi[j]; }
                if (EMS_R64_BASES != 14) {
                    lo[3] = cmul2(lo[1], lo[2]); lo[5] = cmul2(lo[1], lo[4]); lo[6] = cmul2(lo[2], lo[4]); lo[7] = cmul2(lo[3], lo[4]);
                    hi[3] = cmul2(hi[1], hi[2]); hi[5] = cmul2(hi[1], hi[4]); hi[6] = cmul2(hi[2], hi[4]); hi[7] = cmul2(hi[3], hi[4]);
                }
                float2* zr = Zb + tl * kRow;
                static_for<64>([&](auto kc) {
                    constexpr int k = decltype(kc)::value, l = k & 7, h = k >> 3;
                    const float2 o = v[o64(k)];
                    if constexpr (k == 0) zr[0] = o;
                    else if constexpr (h == 0) zr[k] = cmul2(o, lo[l]);
                    else if constexpr (l == 0) zr[k] = cmul2(o, hi[h]);
                    else zr[k] = cmul2(o, cmul2(lo[l], hi[h]));
                });
                worker_bar<kWT>(w);
                if (last_frame && refill[w]) refill_tile(buf, ti + 2);
            };
            auto load_pass2 = [&]() {
                // ================= pass 2: radix-64 over n2 of column k1: Z[k1 + 64 k2] = v[o64(k2)]
#pragma unroll
                for (int b = 0; b < 16; ++b)
#pragma unroll
                    for (int a4 = 0; a4 < 4; ++a4) v[b + 16 * a4] = Zb[(b + 16 * a4) * kRow + k1];
            };
#if EMS_R64_LOOP
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                if (pass == 0) load_pass1(); else load_pass2();
                dft64(v);
                if (pass == 0) store_pass1();
            }
#else
            load_pass1();
            dft64(v);
            store_pass1();
            load_pass2();
            dft64(v);
#endif

            // ================= untangle: 2 X[k] = Z[k] + conj Z[N-k], 2 X_th'[k] = (Z[k] - conj Z[N-k]) / j
            // for k = k1 + 64 i, i < 32; Z[N-k] is output 63 - i of residue 64 - k1 (lane pl), or output
            // (64 - i) mod 64 of residue 0 itself
            float2 X[32];
#if EMS_R64_TMEM && EMS_R64_PARK
            static_for<8>([&](auto gc) {
                constexpr int g4 = 4 * decltype(gc)::value;
                float t8[8];
                static_for<4>([&](auto jc) {
                    constexpr int i = g4 + decltype(jc)::value;
                    const float2 got = shfl2(v[o64(63 - i)], pl);
                    const float2 zn = cj(self0 ? v[o64((64 - i) & 63)] : got);
                    const float2 zk = v[o64(i)];
                    X[i] = zk + zn;
                    const float2 tt = mulmj(zk - zn);
                    t8[2 * (i - g4)] = tt.x; t8[2 * (i - g4) + 1] = tt.y;
                    Zb[i * kRow + k1] = X[i];
                });
                EMS_TMEM_ST8(t8, tmem_T + 2 * g4);
            });
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#else
            float2 T[32];
            static_for<32>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                const float2 got = shfl2(v[o64(63 - i)], pl);
                const float2 zn = cj(self0 ? v[o64((64 - i) & 63)] : got);
                const float2 zk = v[o64(i)];
                X[i] = zk + zn;
                T[i] = mulmj(zk - zn);
                Zb[i * kRow + k1] = X[i];
            });
#endif
            if (k1 == 1) Zb[-2] = cj(X[0]);                     // X[-1] = conj X[1]
            if (k1 == 63) Sc[0] = cj(X[31]);                    // X[2049] = conj X[2047]
            if (self0) {                                        // bin N/2 pairs with itself
                const float2 z = v[o64(32)], zn = cj(z);
                Zb[32 * kRow] = z + zn;
                Sc[1] = mulmj(z - zn);
            }
            worker_bar<kWT>(w);

            // ================= epilogue on the thread's 32 bins, four per vote
            FrameCtx fc;
            fc.lo = (float)max(-f, -1048576LL);
            fc.hi = (float)min(a.F - 1 - f, 1048576LL);
            fc.f = f; fc.ch = ch;
            const long long row0 = ((a.ring ? 0 : (long long)ch * a.F) + f) * B;
            fc.pd = a.dt_cols + row0; fc.pk = a.dk_bins + row0; fc.pe = a.energy + row0;
            FrameCtx fk = fc;
            if (MODE == kStorePoints) {
                fk.pd += k1; fk.pk += k1; fk.pe += k1;
                asm volatile("" : "+l"(fk.pd), "+l"(fk.pk), "+l"(fk.pe));
            }
            // store mode: the neighbour loads of the next four bins are issued before the vote and the
            // stores of the current four; the slow path reads its two neighbours again (as in stft_r16.cuh)
            constexpr bool kPrefetch = EMS_R64_PREFETCH && MODE == kStorePoints;
            float2 nm[4], np_[4];
            auto load_group = [&](int i0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { nm[j] = Xm[(i0 + j) * kRow]; np_[j] = Xp[(i0 + j) * kRow]; }
            };
            if (kPrefetch) load_group(0);
            static_for<8>([&](auto gc) {
                constexpr int i0 = 4 * decltype(gc)::value;
                float2 A4[4], cm[4], cp[4];
                bool lv[4], any = false;
                if (!kPrefetch) load_group(i0);
#pragma unroll
                for (int j = 0; j < 4; ++j) { cm[j] = nm[j]; cp[j] = np_[j]; }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    A4[j] = hann_stencil(X[i0 + j], nm[j], np_[j]);
                    lv[j] = bin_power(A4[j]) > gate_p;
                    any = any || lv[j];
                }
                if constexpr (kPrefetch && i0 + 4 < 32) load_group(i0 + 4);
                if (!__any_sync(0xffffffffu, any)) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) bin_dead<MODE>(fk, true, 64 * (i0 + j));
                } else {
#if EMS_R64_TMEM && EMS_R64_PARK
                    float t8[8];                       // only a group that keeps something fetches its 2 X_th'
                    EMS_TMEM_LD8(t8, tmem_T + 2 * i0);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const float2 T[4] = {make_float2(t8[0], t8[1]), make_float2(t8[2], t8[3]),
                                         make_float2(t8[4], t8[5]), make_float2(t8[6], t8[7])};
                    constexpr int tb0 = i0;
#else
                    constexpr int tb0 = 0;
#endif
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (MODE == kStorePoints || __any_sync(0xffffffffu, lv[j]))
                            bin_tail<N, MODE>(a, fk, true, lv[j], k1 + 64 * (i0 + j), 64 * (i0 + j),
                                              k1f + (float)(64 * (i0 + j)), A4[j],
                                              kPrefetch ? Xm[(i0 + j) * kRow] : cm[j], kPrefetch ? Xp[(i0 + j) * kRow] : cp[j],
                                              T[i0 + j - tb0]);
                        else
                            bin_dead<MODE>(fk, true, 64 * (i0 + j));
                    }
                }
            });
            if (wv == 0)      // bin N/2: the thread of residue 0 (its warp tags along)
                bin_emit<N, MODE>(a, fc, self0, N / 2, (float)(N / 2), Zb[32 * kRow], Zb[31 * kRow + 63], Sc[0], Sc[1]);
        }
    }
#if EMS_R64_TMEM
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
#endif
}

}  // namespace r64
}  // namespace ems
