// stft_generic.cuh — generic-N fused frame gather + 3-window STFT + reassignment.
//
// One CTA analyses one frame at a time (persistent, frame-strided grid).  Per frame ONE FFT:
//   Z = FFT_N( x + j*x*th' ),  th' = th*(2/N)      (rectangular-window X and X_th' packed)
// in shared memory as in-place radix-4 decimation-in-frequency passes (one radix-2 pass if
// log2 is odd); results sit in digit-reversed order.  The epilogue untangles
//   2 X[k] = Z[k] + conj Z[N-k],   2 X_th'[k] = (Z[k] - conj Z[N-k]) / j
// and applies Hann and its derivative as three-tap stencils in frequency
//   X_h[k] = X[k]/2 - (X[k-1] + X[k+1])/4,   X_dh'[k] = (X[k-1] - X[k+1]) / (2j)
// (see stft_r16.cuh for the accuracy argument), then the Auger-Flandrin operators
// (oracle/reassign_oracle.py::reassign_operators).  X_h, X_th, X_dh never leave the SM.
// Used for every n_fft without a tuned kernel.
#pragma once
#include "common.cuh"

namespace ems {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-place DIF FFT of length 2^LOG2LEN in shared memory, forward (e^{-j}) sign.
// tw is the W_Ntab^j table; tw_shift0 = log2(Ntab / LEN).
template <int LOG2LEN>
__device__ __forceinline__ void fft_inplace_dif(float2* __restrict__ buf,
                                                const float2* __restrict__ tw,
                                                int tw_shift0, int tid, int nthreads) {
    constexpr int LEN = 1 << LOG2LEN;
#pragma unroll
    for (int l = LOG2LEN; l >= 2; l -= 2) {
        const int n = 1 << l, m = n >> 2;
        const int sh = tw_shift0 + LOG2LEN - l;
        for (int b = tid; b < LEN / 4; b += nthreads) {
            const int p = b & (m - 1);
            const int base = (b >> (l - 2)) * n + p;
            const float2 a0 = buf[base], a1 = buf[base + m], a2 = buf[base + 2 * m],
                         a3 = buf[base + 3 * m];
            const float2 s02 = make_float2(a0.x + a2.x, a0.y + a2.y);
            const float2 d02 = make_float2(a0.x - a2.x, a0.y - a2.y);
            const float2 s13 = make_float2(a1.x + a3.x, a1.y + a3.y);
            const float2 d13 = make_float2(a1.x - a3.x, a1.y - a3.y);
            // A1 = d02 - j d13, A3 = d02 + j d13
            float2 A0 = make_float2(s02.x + s13.x, s02.y + s13.y);
            float2 A2 = make_float2(s02.x - s13.x, s02.y - s13.y);
            float2 A1 = make_float2(d02.x + d13.y, d02.y - d13.x);
            float2 A3 = make_float2(d02.x - d13.y, d02.y + d13.x);
            if (m > 1) {
                A1 = cmul(A1, __ldg(&tw[(p) << sh]));
                A2 = cmul(A2, __ldg(&tw[(2 * p) << sh]));
                A3 = cmul(A3, __ldg(&tw[(3 * p) << sh]));
            }
            buf[base] = A0; buf[base + m] = A1; buf[base + 2 * m] = A2; buf[base + 3 * m] = A3;
        }
        __syncthreads();
    }
    if (LOG2LEN & 1) {
        for (int b = tid; b < LEN / 2; b += nthreads) {
            const float2 a0 = buf[2 * b], a1 = buf[2 * b + 1];
            buf[2 * b] = make_float2(a0.x + a1.x, a0.y + a1.y);
            buf[2 * b + 1] = make_float2(a0.x - a1.x, a0.y - a1.y);
        }
        __syncthreads();
    }
}

// Position of output index k after fft_inplace_dif (mixed-radix digit reversal).
template <int LOG2LEN>
__device__ __forceinline__ int dif_pos(int k) {
    int pos = 0, rem = k;
#pragma unroll
    for (int l = LOG2LEN; l >= 2; l -= 2) {
        pos += (rem & 3) << (l - 2);
        rem >>= 2;
    }
    if (LOG2LEN & 1) pos += rem & 1;
    return pos;
}

// The reassignment epilogue for one bin, shared by every STFT kernel.
//   A2 = 2 X_h, B2 = 2 X_th', D2 = 2 X_dh'  (the common factor 2 cancels in the ratios)
// Emits (dt_cols, dk_bins, energy) or deposits the energy; `f` = frame in its channel.
template <int N>
__device__ __forceinline__ void reassign_emit(const StftArgs& a, int ch,
                                              long long f, int k, float2 A2, float2 B2,
                                              float2 D2) {
    constexpr int B = N / 2 + 1;
    const float p2 = A2.x * A2.x + A2.y * A2.y;
    const float e = p2 * (float)(4.0 / ((double)N * (double)N));   // |X_h|^2 (4/N)^2, A2 = 2 X_h
    float dtc = 0.f, dk = 0.f;
    bool ok = e > a.gate_lin;
    long long col = f;
    float wh = (float)k;
    if (a.reassign && ok) {
        const float inv = 1.0f / p2;
        const float dts = (B2.x * A2.x + B2.y * A2.y) * inv * (float)(N / 2);   // samples
        dk = (D2.y * A2.x - D2.x * A2.y) * inv * -0.5f;                          // bins
        dtc = dts * a.inv_hop;
        wh = (float)k + dk;
        const float rc = rintf(dtc);
        col = f + (long long)rc;
        const float rowf = (float)k + rintf(dk);     // the row the point lands in decides (exact in fp32)
        ok = (fabsf(dts) <= (float)(N / 2)) && (rowf >= 0.f) && (rowf <= (float)(N / 2)) &&
             (col >= 0) && (col < a.F);
        if (!ok) { dtc = 0.f; dk = 0.f; }
    }
    if (a.mode == kStorePoints) {
        const long long o = ((long long)ch * a.F + f) * B + k;
        a.dt_cols[o] = dtc;
        a.dk_bins[o] = dk;
        a.energy[o] = ok ? e : 0.f;
    } else if (ok) {
        const int row = out_row(a.warp_mode, a.warp_a, a.warp_c, a.inv_half, k, dk, wh);
        const long long o = acc_cell(a, ch, col, row);
        if (a.mode == kDepositU64)
            red_add_u64(reinterpret_cast<unsigned long long*>(a.acc) + o, fix_energy(e));
        else
            red_add_f32(reinterpret_cast<float*>(a.acc) + o, e);
        if (a.flags) flag_set(a.flags + acc_flag(a, ch, col, row));
    }
}

// 2 X[k] for any k in [-1, N/2+1] from the digit-reversed packed spectrum (Hermitian wrap)
template <int LOG2N>
__device__ __forceinline__ float2 rect2(const float2* __restrict__ Z, int k) {
    constexpr int N = 1 << LOG2N;
    const int kk = k < 0 ? -k : (k > N / 2 ? N - k : k);          // X[-k] = X[N-k] = conj X[k]
    const float2 zk = Z[dif_pos<LOG2N>(kk)];
    const float2 zn = Z[dif_pos<LOG2N>((N - kk) & (N - 1))];
    const float2 x = make_float2(zk.x + zn.x, zk.y - zn.y);
    return kk == k ? x : make_float2(x.x, -x.y);
}

template <int LOG2N, int THREADS>
__global__ void __launch_bounds__(THREADS)
stft_reassign_generic(const StftArgs a_in) {
    constexpr int N = 1 << LOG2N;
    StftArgs a = a_in;
    if (!stream_decode(a)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* Z = reinterpret_cast<float2*>(smem_raw);          // [N]
    const int tid = threadIdx.x;
    const long long per_ch = a.f_end - a.f_begin;
    const long long total = per_ch * a.channels;

    for (long long it = blockIdx.x; it < total; it += gridDim.x) {
        const int ch = (int)(it / per_ch);
        const long long f = a.f_begin + (it - (long long)ch * per_ch);
        const float* x = a.pcm + (long long)ch * a.S + f * a.hop + a.samp_off;
        for (int n = tid; n < N; n += THREADS) {
            const float v = __ldg(x + n);
            Z[n] = make_float2(v, v * __ldg(&a.thw[n]));
        }
        __syncthreads();
        fft_inplace_dif<LOG2N>(Z, a.tw, 0, tid, THREADS);

        for (int k = tid; k <= N / 2; k += THREADS) {
            const float2 zk = Z[dif_pos<LOG2N>(k)];
            const float2 zn = Z[dif_pos<LOG2N>((N - k) & (N - 1))];
            const float2 xk = make_float2(zk.x + zn.x, zk.y - zn.y);          // 2 X[k]
            const float2 B2 = make_float2(zk.y + zn.y, zn.x - zk.x);          // 2 X_th'[k]
            const float2 xm = rect2<LOG2N>(Z, k - 1), xp = rect2<LOG2N>(Z, k + 1);
            const float2 A2 = make_float2(0.5f * xk.x - 0.25f * (xm.x + xp.x),
                                          0.5f * xk.y - 0.25f * (xm.y + xp.y));   // 2 X_h
            const float2 D2 = make_float2(0.5f * (xm.y - xp.y), -0.5f * (xm.x - xp.x));   // 2 X_dh'
            reassign_emit<N>(a, ch, f, k, A2, B2, D2);
        }
        __syncthreads();
    }
}

// n_fft whose packed spectrum does not fit the SM (32768: 256 KB): the two real FFTs (x and
// x*th') run one after another as half-size complex FFTs (N/2 points, 4N bytes of shared
// memory); 2 X waits in an L2-resident per-CTA scratch until X_th' is ready.  Same
// arithmetic and decisions as the kernel above; only used where that one cannot be launched.
template <int LOG2N, int THREADS>
__global__ void __launch_bounds__(THREADS)
stft_reassign_big(const StftArgs a_in, float2* __restrict__ scratch_all) {
    constexpr int N = 1 << LOG2N, H2 = N / 2, B = N / 2 + 1;
    StftArgs a = a_in;
    if (!stream_decode(a)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* Y = reinterpret_cast<float2*>(smem_raw);          // [N/2]
    float* Yf = reinterpret_cast<float*>(Y);
    float2* X2 = scratch_all + (size_t)blockIdx.x * (B + 2) + 1;   // 2 X[k], k = -1 .. N/2+1
    const int tid = threadIdx.x;
    const long long per_ch = a.f_end - a.f_begin;
    const long long total = per_ch * a.channels;

    for (long long it = blockIdx.x; it < total; it += gridDim.x) {
        const int ch = (int)(it / per_ch);
        const long long f = a.f_begin + (it - (long long)ch * per_ch);
        const float* x = a.pcm + (long long)ch * a.S + f * a.hop + a.samp_off;
#pragma unroll 1
        for (int wi = 0; wi < 2; ++wi) {
            for (int n = tid; n < N; n += THREADS)
                Yf[n] = wi == 0 ? __ldg(x + n) : __ldg(x + n) * __ldg(&a.thw[n]);
            __syncthreads();
            fft_inplace_dif<LOG2N - 1>(Y, a.tw, 1, tid, THREADS);
            // real FFT of length N from the N/2-point FFT of its even/odd samples
            for (int k = tid; k <= H2; k += THREADS) {
                const float2 yk = Y[dif_pos<LOG2N - 1>(k & (H2 - 1))];
                const float2 yn = Y[dif_pos<LOG2N - 1>((H2 - k) & (H2 - 1))];
                const float2 E2 = make_float2(yk.x + yn.x, yk.y - yn.y);
                const float2 O2 = make_float2(yk.y + yn.y, yn.x - yk.x);
                const float2 w = __ldg(&a.tw[k]);
                const float2 R2 = make_float2(E2.x + (w.x * O2.x - w.y * O2.y),
                                              E2.y + (w.x * O2.y + w.y * O2.x));
                if (wi == 0) {
                    X2[k] = R2;
                    if (k == 1) X2[-1] = make_float2(R2.x, -R2.y);
                    if (k == H2 - 1) X2[H2 + 1] = make_float2(R2.x, -R2.y);
                } else {
                    const float2 xk = X2[k], xm = X2[k - 1], xp = X2[k + 1];
                    const float2 A2 = make_float2(0.5f * xk.x - 0.25f * (xm.x + xp.x),
                                                  0.5f * xk.y - 0.25f * (xm.y + xp.y));
                    const float2 D2 = make_float2(0.5f * (xm.y - xp.y), -0.5f * (xm.x - xp.x));
                    reassign_emit<N>(a, ch, f, k, A2, R2, D2);
                }
            }
            __syncthreads();     // also orders the scratch writes of pass 0 before pass 1 reads them
        }
    }
}

}  // namespace ems
