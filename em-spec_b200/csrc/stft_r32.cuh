// stft_r32.cuh — experiment for n_fft = 8192: the layout of the 4096 kernel
// carried over to 8192 = 32 x 16 x 16.
//
//   * three workers of 128 threads per CTA (384 threads, 168 registers) instead of two of 256 at 128;
//   * pass 1 is ONE radix-32 butterfly per column in registers (two columns per thread, one after the
//     other), so a frame crosses shared memory three times instead of four;
//   * every thread plays two residue roles (p and p + 128) in passes 2 and 3;
//   * 2 X is written in place into the Z rows its role has just read (the XZ path of pass3_untangle /
//     epilogue): no X buffer, three 66 KB Z buffers fit one SM;
//   * samples come straight from global memory / L2 as in stft_reassign_r16_large, software-pipelined in
//     registers: a thread's second column is loaded before the first one's butterfly, and the first column
//     of the worker's NEXT frame before the epilogue of the thread's second role.
//
// Measured (profiles/r02_8192_w3.txt, r02_ncu_8192_w3.txt): 3,461 shared-memory wavefronts per frame
// against 4,499 and 6.5 % fewer instructions, as designed — and no faster.  Without the pipelining pass 1
// is 59 % of the samples (long scoreboard: 32 loads, then nothing to do, with 12 warps per SM): 31.4 M
// frames/s against 33.0 M.  With it 31.9 M (hop 2048) / 36.9 M (hop 256) against 33.0 / 36.9 M: the 96
// values kept across the epilogue spill (240 bytes of stack) and the spill reloads stall it instead.
// Staging the next frame's column with cp.async in the free slots 9..15 of the thread's own Z rows
// removes that pressure but costs 13 % more instructions (addresses) and 256 wavefronts: 31.5 / 34.6 M.
// Three warps per scheduler issue 39 % of the time whatever the variant; the two-worker kernel stays.
//
// Selected at run time with EMS_KERNEL_VARIANT=32 (engine.cu); results are the same points as the
// default kernel up to fp32 rounding of a different (but equally exact-twiddle) factorisation.
#pragma once
#include "stft_r16.cuh"

namespace ems {
namespace r16 {

struct Cfg8kW3 {
    static constexpr int R = 32, N = 8192, kWT = 128, kWorkers = 3, kThreads = kWT * kWorkers;
    static constexpr int kSI = 257, kS16 = 16;
    static constexpr int kZBuf = R * kSI;                       // 8224 float2
    static constexpr int kSlot = kZBuf + kScratch;
    static constexpr int kZtab = 5 * 256;                       // W_N^{b 2^l}, l = 0..4, b < 256
    static constexpr int kSmemBytes = (kWorkers * kSlot + kT2 + kZtab) * 8;
    static_assert(kSmemBytes <= kMaxSmem, "shared memory");
};

// Column b of pass 1: z[n] = x[n] (1 + j th'[n]) at n = b + 256 j, j = 0..31; radix-32 butterfly in registers
// (even / odd j: two DFT-16, W_32 twiddles, radix 2); output i times W_N^{b i} becomes element b of
// sub-FFT i, Zb[kSI i + b].  th'[n] = (n - N/2)(2/N)(0.5 - 0.5 cos(theta_b + j pi/16)) by angle addition.
__device__ __forceinline__ void load_column32(float (&xr)[32], const float* __restrict__ xs, int b) {
#pragma unroll
    for (int j = 0; j < 32; ++j) xr[j] = __ldg(xs + b + 256 * j);
}
__device__ __forceinline__ void pass1_radix32(const float (&xr)[32], const float2* Ztab, int b, float2* Zb) {
    constexpr int N = 8192, kSI = 257;
    float2 wb[5];                                                // W_N^{b 2^l}: exact table values
#pragma unroll
    for (int l = 0; l < 5; ++l) wb[l] = Ztab[256 * l + b];
    float2 ev[16], od[16];
    {
        const float cb = wb[0].x, sb = -wb[0].y, rb = (float)(b - N / 2) * (2.0f / N);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float cs = fmaf(cb, c32(j), -(sb * s32(j)));
            const float th = (rb + (float)j * (256.0f * 2.0f / N)) * fmaf(-0.5f, cs, 0.5f);
            const float2 z = make_float2(xr[j], xr[j] * th);
            if (j & 1) od[j >> 1] = z; else ev[j >> 1] = z;
        }
    }
    dft16(ev); dft16(od);
    float2 lo[8], hi[4];
    lo[1] = wb[0]; lo[2] = wb[1]; lo[4] = wb[2];
    lo[3] = cmul2(lo[1], lo[2]); lo[5] = cmul2(lo[1], lo[4]); lo[6] = cmul2(lo[2], lo[4]); lo[7] = cmul2(lo[3], lo[4]);
    hi[1] = wb[3]; hi[2] = wb[4]; hi[3] = cmul2(hi[1], hi[2]);
    float2* zo = Zb + b;
    static_for<16>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        const float2 e = ev[o16(i)];
        const float2 o = i == 0 ? od[o16(0)] : cmul2(od[o16(i)], make_float2(c32(i), -s32(i)));   // W_32^i
        const float2 x0 = e + o, x1 = e - o;                     // outputs i and i + 16
        auto tw = [&](auto kc) {
            constexpr int k = decltype(kc)::value, l = k & 7, h = k >> 3;
            if constexpr (l == 0) return hi[h];
            else if constexpr (h == 0) return lo[l];
            else return cmul2(lo[l], hi[h]);
        };
        if constexpr (i == 0) zo[0] = x0;
        else zo[kSI * i] = cmul2(x0, tw(std::integral_constant<int, i>{}));
        zo[kSI * (i + 16)] = cmul2(x1, tw(std::integral_constant<int, i + 16>{}));
    });
}

template <int MODE>
__global__ void __launch_bounds__(Cfg8kW3::kThreads, 1)
stft_reassign_8192_w3(const StftArgs a_in) {
    using C = Cfg8kW3;
    constexpr int R = C::R, kWT = C::kWT, kWorkers = C::kWorkers, kSI = C::kSI, kS16 = C::kS16, kThreads = C::kThreads;
    StftArgs a = a_in;
    if (!stream_decode(a)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* T2 = reinterpret_cast<float2*>(smem_raw);           // [16][16]
    float2* Ztab = T2 + kT2;                                    // [5][256]: W_N^{b 2^l}
    const int tid = threadIdx.x;
    const int w = tid / kWT;
    const int p = (tid + 32 * w) & (kWT - 1);                   // warp roles rotate across the workers
    float2* Zb = Ztab + C::kZtab + w * C::kSlot;
    float2* Sc = Zb + C::kZBuf;
    for (int e = tid; e < 256; e += kThreads) { const int q = e / 16, i = e % 16; T2[kT2S * q + i] = __ldg(&a.tw[R * q * i]); }
    for (int e = tid; e < C::kZtab; e += kThreads) { const int l = e / 256, b = e % 256; Ztab[e] = __ldg(&a.tw[(b << l) & (C::N - 1)]); }
    __syncthreads();
    const Geom g0 = make_geom<R, kSI, kS16>(p), g1 = make_geom<R, kSI, kS16>(p + kWT);

    const long long per_ch = a.f_end - a.f_begin;
    const long long total = per_ch * a.channels;
    auto frame_ptr = [&](long long it) -> const float* {
        const int ch = (int)(it / per_ch);
        const long long f = a.f_begin + (it - (long long)ch * per_ch);
        return a.pcm + (long long)ch * a.S + f * a.hop + a.samp_off;
    };
    const long long it0 = blockIdx.x + (long long)gridDim.x * w, it_step = (long long)gridDim.x * kWorkers;
    float xr0[32];                                              // column p of the frame about to start
    if (it0 < total) load_column32(xr0, frame_ptr(it0), p);
    for (long long it = it0; it < total; it += it_step) {
        const int ch = (int)(it / per_ch);
        const long long f = a.f_begin + (it - (long long)ch * per_ch);
        const float* xs = a.pcm + (long long)ch * a.S + f * a.hop + a.samp_off;

        {
            float xr1[32];
            load_column32(xr1, xs, p + kWT);                    // in flight during the first column's butterfly
            pass1_radix32(xr0, Ztab, p, Zb);
            pass1_radix32(xr1, Ztab, p + kWT, Zb);
        }
        worker_bar<kWT>(w);

        pass2<kSI, kS16>(Zb, T2, g0);
        pass2<kSI, kS16>(Zb, T2, g1);
        worker_bar<kWT>(w);

        // pass 3 of both roles; role 0 keeps only 2 X_th' in registers (its 2 X is read back from the rows)
        float2 xa[8], xb[8], ta[8], tb[8], ta0[8], tb0[8];
        pass3_untangle<R, true>(Zb, nullptr, Sc, g0, xa, xb, ta0, tb0);
        pass3_untangle<R, true>(Zb, nullptr, Sc, g1, xa, xb, ta, tb);
        worker_bar<kWT>(w);      // 2 X of every role visible

        epilogue<R, MODE, true>(a, ch, f, Zb, Sc, g1, xa, xb, ta, tb);
        if (it + it_step < total) load_column32(xr0, frame_ptr(it + it_step), p);     // in flight during role 0's epilogue
#pragma unroll
        for (int c = 0; c < 8; ++c) { xa[c] = Zb[g0.zA + c]; xb[c] = Zb[g0.zB + c]; }
        epilogue<R, MODE, true>(a, ch, f, Zb, Sc, g0, xa, xb, ta0, tb0);
        worker_bar<kWT>(w);      // 2 X is dead: the next frame's pass 1 may overwrite the rows
    }
}

}  // namespace r16
}  // namespace ems
