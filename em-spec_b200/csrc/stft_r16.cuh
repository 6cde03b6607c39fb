// stft_r16.cuh — tuned fused frame gather + 3-window STFT + reassignment for
// n_fft = 256 R, R = 1 .. 16 (256 .. 4096); 8192, 16384 and 32768 at the end of the file.
//
// ONE complex FFT per frame.  Z = FFT_N(x + j*x*th'), th' = th*(2/N): the real part is the
// *unwindowed* frame, so the untangle gives the rectangular-window spectrum X and X_th'.
// Hann and its derivative are three-tap stencils of X in frequency
//     X_h[k]   = X[k]/2 - (X[k-1] + X[k+1])/4          (h  = 0.5 - 0.5 cos(2 pi n/N))
//     X_dh'[k] = (X[k-1] - X[k+1]) / (2j)              (dh' = sin(2 pi n/N))
// which replaces the second FFT of the x*dh window (a third of the flops and 40 % of the
// shared-memory traffic).  fp32 accuracy is that of the three-FFT form: only t*h needs a time
// window (an unwindowed ramp is what breaks the all-stencil variant SURVEY.md §7 rejected);
// probed against the float64 oracle: <= 4.3e-5 col / 5.2e-6 bin within 40 dB of the peak.
//
// Layout of the work (DESIGN.md "K1-K3"):
//   * persistent CTAs, one per SM, 48/R workers x 8R threads (3 x 128 at n_fft = 4096); a CTA
//     walks tiles of T consecutive frames whose samples ((T-1)*hop + N floats) sit once in
//     shared memory, double-buffered with cp.async (one TMA bulk copy per tile for large aligned
//     hops); buffers are handed over through an mbarrier and a release counter, never a CTA-wide
//     barrier, so the workers run out of phase;
//   * a worker analyses one frame at a time, entirely on-chip: Z as R x 16 x 16 — pass 1 is
//     32/R radix-R butterflies per thread straight from the tile, passes 2 and 3 two radix-16
//     butterflies per thread — in packed fp32x2 arithmetic, in-place decimation-in-frequency
//     in the worker's private buffer; the padding (Cfg::kSI, Cfg::kS16) makes every exchange
//     bank-conflict-free (tools/bank_sim.py);
//   * th' of a thread's 32 pass-1 samples is frame-independent and lives in registers;
//   * the last pass gives thread p the output residues t = p and 16R - p (mod 16R), so Z[k] and
//     Z[N-k] of its 16 bins are in its own registers: untangle in registers, X (N/2+1 values)
//     goes to shared memory once so that each bin can read its two neighbours, then the
//     Auger-Flandrin epilogue runs on the thread's own bins.  3 worker barriers per frame.
//     Residues 0 and 8R pair with themselves: thread 0 selects its partners inside the common
//     instruction stream and finishes the one extra bin (N/2) on the side.
// Decisions (gate, drop rule, deposit) are those of stft_generic.cuh::reassign_emit.
#pragma once
#include "common.cuh"
#include "stft_generic.cuh"

#include <type_traits>
#include <utility>

#ifndef EMS_LARGE_PREFETCH
#define EMS_LARGE_PREFETCH 1    // samples and pass-A twiddles of a worker's NEXT frame are fetched with cp.async during the epilogue, into
                                // the Z slots the fetching warp itself overwrites in pass A.  1: n_fft = 16384 only (one worker per SM: +6.5 %);
                                // 2: 8192 too (two workers already hide each other's loads: -1..-4 %); 0: off
#endif
#ifndef EMS_LARGE_R32
#define EMS_LARGE_R32 0         // experiment, n_fft = 8192: pass 1 as one radix-32 butterfly per thread (one exchange and one
                                // barrier fewer than 2 x 16): correct, 30.9 against 33.2 M frames/s (32 loads and 64 live
                                // registers per thread, spills at 128) — off
#endif
#ifndef EMS_R16_W4
#define EMS_R16_W4 0            // experiment, n_fft = 4096: a fourth worker (16 warps at <= 128 registers):
                                // 2 X written in place into the Z rows the thread read (no X buffer), th' from a shared-memory
                                // table instead of 32 registers, one more worker barrier per frame
#endif
#ifndef EMS_FUSED_POST
#define EMS_FUSED_POST 0        // experiment: in-kernel post-pass of the deposit kernels (see fused_post_block): correct, 5 x
                                // slower, and its mere presence costs the deposit kernels 5 % — compiled out by default
#endif
#ifndef EMS_DEPOSIT_AGG
#define EMS_DEPOSIT_AGG 0       // experiment: warp-aggregated deposits (match.any + segmented shuffle sum)
#endif

namespace ems {
namespace r16 {

constexpr int kMaxSmem = 232448;       // 227 KB
constexpr int kSyncBytes = 192;        // 2 mbarriers, 2 release counters, kWorkers refill flags; fused mode: 4 finish counters, kWorkers go flags
constexpr int kT2S = 18;               // row stride of T2: 144 bytes, so that the 8 rows a quarter-warp of the narrow kernels
                                       // (n_fft <= 1024: up to 8 different p2 among 8 lanes) reads with one LDS.128 sit on different banks
constexpr int kT2 = 16 * kT2S;         // W_256^{p2 i} as [p2][i]: a butterfly's 16 twiddles are contiguous
constexpr int kScratch = 20;           // 2 X_th' of bin N/2 (and padding: the slot size stays 8 mod 16)
constexpr int kThreads = 384;

template <int R_>
struct Cfg {
    static_assert(R_ == 1 || R_ == 2 || R_ == 4 || R_ == 8 || R_ == 16, "pass-1 radix");
    static constexpr int R = R_;                    // radix of pass 1; passes 2 and 3 are radix 16
    static constexpr int kLogR = R == 1 ? 0 : R == 2 ? 1 : R == 4 ? 2 : R == 8 ? 3 : 4;
    static constexpr int N = 256 * R;
    static constexpr int kWT = 8 * R;               // threads per frame
    // a worker is kWT threads, or one warp analysing kG frames side by side when kWT < 32
    static constexpr int kG = kWT < 32 ? 32 / kWT : 1;
    static constexpr int kWorkerThreads = kWT * kG;
    static constexpr bool kW4 = EMS_R16_W4 && R == 16;           // four workers, X in place, th' table in shared memory
    static constexpr int kCta = kW4 ? 512 : kThreads;            // threads per CTA
    static constexpr int kWorkers = kCta / kWorkerThreads;
    static constexpr int kU = 32 / R;               // pass-1 butterflies per thread
    static constexpr int kRes = 16 * R;             // output residues of the last pass: k = t + kRes c
    // Z buffer: element b of sub-FFT i sits at kSI i + b + (b >> 4) (kS16 - 16).  With R = 16 the
    // classic 257 stride does it; narrower pass-1 radices put several sub-FFT rows of the same i
    // into one half-warp, so every 16 elements get one slot of padding and kSI = 16/R (mod 16).
    static constexpr int kS16 = R >= 16 ? 16 : 17;
    static constexpr int kSI = R >= 16 ? 257 : 272 + 16 / R;
    static constexpr int kZBuf = R >= 16 ? 4112 : R * kSI;    // (R = 1: 288 makes the slot size 8 mod 16, so the
                                                              //  frames of one half-warp sit in different banks)
    static constexpr int kXBuf = kW4 ? 0 : N / 2 + 4;   // float2: 2 X[k] at index k + 1, mirrors at 0 and N/2 + 2
    static constexpr int kZtab = kLogR * 256;       // W_N^{b i}, i = 1, 2, 4, .., b < 256 (the other powers are products)
    static constexpr int kTabFloat2 = kZtab + kT2;
    static constexpr int kSlot = kZBuf + kXBuf + kScratch;      // per frame in flight: Z, X, scratch
    static constexpr int kThwBytes = kW4 ? N * 4 : 0;            // th'[n] table
    static constexpr int kFixedBytes = (kWorkers * kG * kSlot + kTabFloat2) * 8 + kThwBytes;
    static constexpr int kTileFloats = ((kMaxSmem - kFixedBytes - kSyncBytes) / 8) & ~3;   // per buffer, two buffers
    static constexpr int kMaxTile = 12 * kWorkers * kG;
    // frames per tile: both tile buffers must hold (T-1)*hop + N samples
    __host__ __device__ static constexpr int tile_frames(int hop) {
        int t = (kTileFloats - N) / hop + 1;
        if (t > kWorkers * kG) t -= t % (kWorkers * kG);
        return t > kMaxTile ? kMaxTile : t;
    }
    __host__ __device__ static constexpr int zpos(int i, int b) { return kSI * i + b + (b >> 4) * (kS16 - 16); }
    static_assert(kCta % kWorkerThreads == 0, "workers tile the CTA");
    static_assert(kZBuf >= zpos(R - 1, 255) + 1, "Z buffer holds every sub-FFT");
    static_assert(kTileFloats >= N && kTileFloats % 4 == 0, "each tile buffer holds at least one frame, 16-byte granular");
    static_assert(kFixedBytes % 16 == 0, "tile buffers (TMA / cp.async destinations) start on a 16-byte boundary");
    static_assert(kFixedBytes + 2 * kTileFloats * 4 + kSyncBytes <= kMaxSmem, "shared memory");
    static_assert(kG == 1 || kSlot % 16 == 8, "frames of one half-warp sit in different banks");
    // X in place (kW4): 2 X[k], k = t + 256 c, sits at xslot(k) = zrow(t) + c, c <= 8 — slots of the row the owner
    // of residue t read in pass 3; the mirrors 2 X[-1], 2 X[N/2+1] sit in the scratch (Sc[2], Sc[3])
    __host__ __device__ static constexpr int zrow(int t) { return kSI * (t & (R - 1)) + kS16 * (t / R); }
    static_assert(16 + 8 + 4 * kWorkers + 16 + 4 * kWorkers <= kSyncBytes, "mbarriers, release counters, refill flags, finish counters and go flags fit their block");
};

template <int... Is, class Fn>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, Fn&& fn) {
    (fn(std::integral_constant<int, Is>{}), ...);
}
// fn(integral_constant<int, 0>) ... fn(integral_constant<int, Count - 1>): loop indices usable
// as template arguments and constexpr table indices
template <int Count, class Fn>
__device__ __forceinline__ void static_for(Fn&& fn) {
    static_for_impl(std::make_integer_sequence<int, Count>{}, fn);
}

// ---- packed fp32x2 arithmetic (sm_100 FADD2 / FMUL2 / FFMA2).  A complex value lives in an
// aligned register pair; one packed instruction does the work of two scalar ones at the same
// lane throughput, i.e. half the issue slots (tools/microbench/fp32x2.cu).  ptxas folds the
// half swaps and per-half negations written below into operand modifiers (.LO_HI, .NP) and
// single-register broadcasts (.F32), so rotations by +-j and complex multiplies cost no moves.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float x, float y) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ u64 pk(float2 a) { return pk(a.x, a.y); }
__device__ __forceinline__ float2 up(u64 a) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(a)); return r; }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b))); return up(r); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b))); return up(r); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b))); return up(r); }
// a * b + c, element-wise
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c))); return up(r); }
__device__ __forceinline__ float2 operator+(float2 a, float2 b) { return add2(a, b); }
__device__ __forceinline__ float2 operator-(float2 a, float2 b) { return sub2(a, b); }
// -j * a and +j * a
__device__ __forceinline__ float2 mulmj(float2 a) { return make_float2(a.y, -a.x); }
__device__ __forceinline__ float2 mulpj(float2 a) { return make_float2(-a.y, a.x); }
// complex product v * w in two packed instructions
__device__ __forceinline__ float2 cmul2(float2 v, float2 w) {
    const float2 t = mul2(v, make_float2(w.x, w.x));
    return fma2(make_float2(-v.y, v.x), make_float2(w.y, w.y), t);
}

// forward DFT-4 in place: (x0, x1, x2, x3) <- outputs 0..3
__device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 s02 = x0 + x2, d02 = x0 - x2, s13 = x1 + x3, d13 = x1 - x3;
    x0 = s02 + s13;
    x2 = s02 - s13;
    x1 = d02 + mulmj(d13);
    x3 = d02 - mulmj(d13);
}

// multiply by W16^E = exp(-2 pi j E / 16), E compile-time
template <int E>
__device__ __forceinline__ float2 mul_w16(float2 v) {
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R = 0.70710678118654752f;
    constexpr int e = E & 15;
    if constexpr (e == 0) return v;
    else if constexpr (e == 4) return mulmj(v);
    else if constexpr (e == 8) return make_float2(-v.x, -v.y);
    else if constexpr (e == 12) return mulpj(v);
    else {
        // (c, -s): v * W = v * c + (-j v) * s
        constexpr float c = (e == 1 || e == 15) ? C1 : (e == 3 || e == 13) ? S1
                          : (e == 5 || e == 11) ? -S1 : (e == 7 || e == 9) ? -C1
                          : (e == 2 || e == 14) ? R : -R;                  // e == 6, 10
        constexpr float s = (e == 1 || e == 7) ? S1 : (e == 3 || e == 5) ? C1
                          : (e == 9 || e == 15) ? -S1 : (e == 11 || e == 13) ? -C1
                          : (e == 2 || e == 6) ? R : -R;                   // e == 10, 14
        return fma2(mulmj(v), make_float2(s, s), mul2(v, make_float2(c, c)));
    }
}

// forward DFT-16 in registers.  Output i sits in a[o16(i)].
__host__ __device__ constexpr int o16(int i) { return 4 * (i & 3) + (i >> 2); }

__device__ __forceinline__ void dft16(float2 (&a)[16]) {
#pragma unroll
    for (int j0 = 0; j0 < 4; ++j0) dft4(a[j0], a[j0 + 4], a[j0 + 8], a[j0 + 12]);
    // a[j0 + 4 i0] *= W16^{i0 j0}
    a[5] = mul_w16<1>(a[5]);   a[9] = mul_w16<2>(a[9]);   a[13] = mul_w16<3>(a[13]);
    a[6] = mul_w16<2>(a[6]);   a[10] = mul_w16<4>(a[10]); a[14] = mul_w16<6>(a[14]);
    a[7] = mul_w16<3>(a[7]);   a[11] = mul_w16<6>(a[11]); a[15] = mul_w16<9>(a[15]);
#pragma unroll
    for (int i0 = 0; i0 < 4; ++i0) dft4(a[4 * i0], a[4 * i0 + 1], a[4 * i0 + 2], a[4 * i0 + 3]);
}

// forward DFT-8 in registers.  Output i sits in a[o8(i)].
__host__ __device__ constexpr int o8(int i) { return 2 * (i & 3) + (i >> 2); }

__device__ __forceinline__ void dft8(float2 (&a)[8]) {
    dft4(a[0], a[2], a[4], a[6]);
    dft4(a[1], a[3], a[5], a[7]);
    // a[1 + 2 i0] *= W8^{i0}
    a[3] = mul_w16<2>(a[3]);   a[5] = mul_w16<4>(a[5]);   a[7] = mul_w16<6>(a[7]);
#pragma unroll
    for (int i0 = 0; i0 < 4; ++i0) {
        const float2 s = a[2 * i0] + a[2 * i0 + 1], d = a[2 * i0] - a[2 * i0 + 1];
        a[2 * i0] = s; a[2 * i0 + 1] = d;
    }
}

// radix-R butterfly of pass 1; output i sits in a[oR<R>(i)]
template <int R> __host__ __device__ constexpr int oR(int i) { return R == 16 ? o16(i) : R == 8 ? o8(i) : i; }
__device__ __forceinline__ void dftR(float2 (&a)[1]) {}
__device__ __forceinline__ void dftR(float2 (&a)[2]) { const float2 s = a[0] + a[1], d = a[0] - a[1]; a[0] = s; a[1] = d; }
__device__ __forceinline__ void dftR(float2 (&a)[16]) { dft16(a); }
__device__ __forceinline__ void dftR(float2 (&a)[8]) { dft8(a); }
__device__ __forceinline__ void dftR(float2 (&a)[4]) { dft4(a[0], a[1], a[2], a[3]); }

// barrier of one worker (a single warp when n_fft = 1024)
template <int WT>
__device__ __forceinline__ void worker_bar(int w) {
    if constexpr (WT <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(w + 1), "r"(WT) : "memory");
}

// cos(q pi/16), sin(q pi/16) for compile-time q
__host__ __device__ constexpr float c32(int q) {
    constexpr float t[32] = {
        1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f,
        0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f, 0.0f, -0.19509032201612825f,
        -0.38268343236508977f, -0.55557023301960218f, -0.70710678118654752f, -0.83146961230254524f,
        -0.92387953251128674f, -0.98078528040323043f, -1.0f, -0.98078528040323043f,
        -0.92387953251128674f, -0.83146961230254524f, -0.70710678118654752f, -0.55557023301960218f,
        -0.38268343236508977f, -0.19509032201612825f, 0.0f, 0.19509032201612825f, 0.38268343236508977f,
        0.55557023301960218f, 0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f,
        0.98078528040323043f};
    return t[q & 31];
}
__host__ __device__ constexpr float s32(int q) { return c32(q - 8); }   // sin(q pi/16) = cos(q pi/16 - pi/2)

// Per-frame constants of the epilogue.
struct FrameCtx {
    float lo, hi;        // bounds on rint(dt_cols) keeping the column inside [0, F-1]
    float* pd;           // row (chan, f) of dt_cols / dk_bins / energy (store mode)
    float* pk;
    float* pe;
    long long f;         // frame index in its channel
    int ch;
};

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// a4 for one kept point.  Deliberately not inlined: the row warp (log1pf), the 64-bit cell and
// flag indices and the atomics would otherwise sit 17 times in the frame loop and push it out
// of the instruction cache (measured: no_instruction stalls 1.0 per issue, -8 % throughput).
struct DepositCtx {      // passed by value (registers): taking the address of StftArgs would
    void* acc;           // push the whole argument block into local memory
    unsigned char* flags;
    long long F;
    int ring, rows, warp_mode;
    float warp_a, warp_c, inv_half;
    int vring, NB;       // fused mode: one ring of vring virtual columns (ch * F + col), flags [vring][NB]
};
template <int MODE>
__device__ __noinline__ void deposit_point(const DepositCtx d, int ch, long long col, int k, float dk,
                                           float wh, float e) {
    const int row = out_row(d.warp_mode, d.warp_a, d.warp_c, d.inv_half, k, dk, wh);
    if (EMS_FUSED_POST && d.vring) {
        const long long slot = ((long long)ch * d.F + col) & (long long)(d.vring - 1);
        if (MODE == kDepositU64)
            red_add_u64(reinterpret_cast<unsigned long long*>(d.acc) + slot * d.rows + row, fix_energy(e));
        else
            red_add_f32(reinterpret_cast<float*>(d.acc) + slot * d.rows + row, e);
        flag_set(d.flags + slot * d.NB + (row >> kFlagShift));
        return;
    }
    const long long ncols = d.ring ? d.ring : d.F, slot = d.ring ? (col & (d.ring - 1)) : col;
    const long long o = ((long long)ch * ncols + slot) * d.rows + row;
    if (MODE == kDepositU64)
        red_add_u64(reinterpret_cast<unsigned long long*>(d.acc) + o, fix_energy(e));
    else
        red_add_f32(reinterpret_cast<float*>(d.acc) + o, e);
    if (d.flags) flag_set(d.flags + flag_index(ch, ncols, d.rows, slot, row));
}

// One bin of the epilogue: Hann / Hann-derivative stencils of the rectangular spectrum, the
// Auger-Flandrin operators and the emit — same decisions as reassign_emit (stft_generic.cuh).
//   xk, xm, xp = 2 X[k], 2 X[k-1], 2 X[k+1];  t2 = 2 X_th'[k]
// Split in three so that the epilogue can run the gate of all its bins before any branch:
//   hann_stencil  4 X_h from the three rectangular bins (one add, one fused multiply-add: every
//                 scale from here on is a power of two, folded into the constants, and exact)
//   bin_power     |4 X_h|^2; the gate compares it with gate * N^2 (energy = |4 X_h|^2 / N^2)
//   bin_dead      a bin nobody in the warp keeps: zeros in store mode, nothing otherwise
//   bin_tail      reassignment operators, drop rule, store / deposit (predicated, no divergence)
__device__ __forceinline__ float2 hann_stencil(float2 xk, float2 xm, float2 xp) {
    return fma2(xm + xp, make_float2(-0.5f, -0.5f), xk);
}
__device__ __forceinline__ float bin_power(float2 A4) {
    return __fmaf_rn(A4.x, A4.x, __fmul_rn(A4.y, A4.y));
}
template <int N>
__device__ __forceinline__ float gate_power(const StftArgs& a) { return a.gate_lin * (float)N * (float)N; }
// `off` is the bin's index relative to the row pointers of fc (k itself, or a compile-time
// multiple of the residue stride when fc points at the thread's residue: the stores then take
// immediate offsets from six live base pointers instead of a 64-bit address computation each).
#ifndef EMS_STORE_POLICY
#define EMS_STORE_POLICY 0      // trial: 1 = st.global.cs (evict-first) for the point stores, 2 = st.global.wt
#endif
__device__ __forceinline__ void st_point(float* p, float v) {
#if EMS_STORE_POLICY == 1
    __stcs(p, v);
#elif EMS_STORE_POLICY == 2
    __stwt(p, v);
#else
    __stwb(p, v);
#endif
}
template <int MODE>
__device__ __forceinline__ void bin_dead(const FrameCtx& fc, bool owner, int off) {
    // __stwb = st.global.wb, the default policy spelled out: the row pointers went through an
    // opaque asm and would otherwise be stored through as generic addresses
    if (MODE == kStorePoints && owner) { st_point(fc.pd + off, 0.f); st_point(fc.pk + off, 0.f); st_point(fc.pe + off, 0.f); }
}
template <int N, int MODE>
__device__ __forceinline__ void bin_tail(const StftArgs& a, const FrameCtx& fc, bool owner, bool live, int k, int off,
                                         float kf, float2 A4, float2 xm, float2 xp, float2 t2) {
    // explicit fused multiply-adds: every instantiation (store / deposit, every n_fft) rounds alike,
    // so a grid deposited by the fused kernel equals the scatter of the stored points bit for bit
    const float p4 = bin_power(A4);
    const float e = p4 * (float)(1.0 / ((double)N * (double)N));
    bool ok = live;
    float dtc = 0.f, dk = 0.f, rc = 0.f, wh = kf;
    if (a.reassign) {
        const float2 d = xm - xp;                                        // 2 X_dh' = d / (2j)
        const float2 D2 = make_float2(0.5f * d.y, -0.5f * d.x);
        const float inv = rcp_approx(p4);
        const float dts = __fmaf_rn(t2.x, A4.x, __fmul_rn(t2.y, A4.y)) * inv * (float)N;          // samples
        dk = __fmaf_rn(D2.y, A4.x, -__fmul_rn(D2.x, A4.y)) * inv * -1.0f;                          // bins
        dtc = dts * a.inv_hop;
        rc = rintf(dtc);
        wh = kf + dk;
        // the row the point lands in, k + rint(dk), decides (exact: two small integers in fp32); a test on
        // wh = k + dk would round in fp32 and let a point at N/2 + 0.5 through to row N/2 + 1
        const float rowf = kf + rintf(dk);
        ok = live && (fabsf(dts) <= (float)(N / 2)) && (rowf >= 0.f) && (rowf <= (float)(N / 2)) &&
             (rc >= fc.lo) && (rc <= fc.hi);
        dtc = ok ? dtc : 0.f;
        dk = ok ? dk : 0.f;
    }
    if (MODE == kStorePoints) {
        if (owner) { st_point(fc.pd + off, dtc); st_point(fc.pk + off, dk); st_point(fc.pe + off, ok ? e : 0.f); }
    } else {
#if EMS_DEPOSIT_AGG
        // Experiment: warp-level pre-aggregation.  Lanes hold adjacent bins of one frame, so
        // the bins of a main lobe land in the same cell: lanes with equal (column, row) that sit next to
        // each other sum their (fixed-point) energies with a segmented shuffle reduction and the head of
        // each run issues one reduction.  Integer sums: the grid stays bit-exact.  All 32 lanes are here.
        if (a.warp_mode == 0 && !a.fp.vring) {
            const bool mine = ok && owner;
            const int row = k + (int)rintf(dk);
            const unsigned key = mine ? (((unsigned)((int)rc + 32768) << 16) | (unsigned)row) : (0xffffff00u | (threadIdx.x & 31u));
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            const unsigned lane = threadIdx.x & 31u;
            const int n = __ffs(~(peers >> lane)) - 1;                        // lanes of my run from me upwards
            const bool head = lane == 0 || !((peers >> (lane - 1)) & 1u);
            const long long col = fc.f + (long long)rc;
            const long long ncols = a.ring ? a.ring : a.F, slot = a.ring ? (col & (a.ring - 1)) : col;
            const long long o = ((long long)fc.ch * ncols + slot) * a.rows + row;
            if (MODE == kDepositU64) {
                unsigned long long v = mine ? fix_energy(e) : 0ull;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned long long t = __shfl_down_sync(0xffffffffu, v, d);
                    if (d < n) v += t;
                }
                if (mine && head) red_add_u64(reinterpret_cast<unsigned long long*>(a.acc) + o, v);
            } else {
                float v = mine ? e : 0.f;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const float t = __shfl_down_sync(0xffffffffu, v, d);
                    if (d < n) v += t;
                }
                if (mine && head) red_add_f32(reinterpret_cast<float*>(a.acc) + o, v);
            }
            if (mine && head && a.flags) flag_set(a.flags + flag_index(fc.ch, ncols, a.rows, slot, row));
            return;
        }
#endif
        if (ok && owner) {
            const DepositCtx d{a.acc, a.flags, a.F, a.ring, a.rows, a.warp_mode, a.warp_a, a.warp_c, a.inv_half, a.fp.vring, a.fp.NB};
            deposit_point<MODE>(d, fc.ch, fc.f + (long long)rc, k, dk, wh, e);
        }
    }
}
// One bin start to end.  A whole warp under the gate leaves after the stencil.  Must be reached by
// all 32 lanes of the warp (`owner` masks lanes that only tag along).
template <int N, int MODE>
__device__ __forceinline__ void bin_emit(const StftArgs& a, const FrameCtx& fc, bool owner, int k,
                                         float kf, float2 xk, float2 xm, float2 xp, float2 t2) {
    const float2 A2 = hann_stencil(xk, xm, xp);                          // 4 X_h
    const bool live = bin_power(A2) > gate_power<N>(a);
    if (!__any_sync(0xffffffffu, live)) { bin_dead<MODE>(fc, owner, k); return; }
    bin_tail<N, MODE>(a, fc, owner, live, k, k, kf, A2, xm, xp, t2);
}

// conj(a)
__device__ __forceinline__ float2 cj(float2 a) { return make_float2(a.x, -a.y); }

// ---- the pieces of one frame, shared by the tiled kernel (n_fft <= 4096) and the large one

// Roles of one thread inside its worker (frame-independent).
struct Geom {
    int p;              // role index, 0 .. kWT - 1
    int tA, tB;         // output residues of the last pass: bins tA + kRes c and tB + kRes c
    int zA, zB;         // where their 16 inputs start in the Z buffer
    int i1, q2;         // pass-2 butterflies (i1, q2) and (i1, q2 + 8)
    float tAf, tBf;
};
// X in place (XZ): start of the Z row that holds the pass-3 inputs, and afterwards 2 X[t + 16 R c] at slot c, of residue t
template <int R>
__host__ __device__ constexpr int zrow_xz(int t) {
    static_assert(R >= 16, "classic 257 / 16 strides");
    return 257 * (t & (R - 1)) + 16 * (t / R);
}
template <int R, int kSI, int kS16>
__device__ __forceinline__ Geom make_geom(int p) {
    constexpr int kRes = 16 * R;
    Geom g;
    g.p = p;
    g.tA = p; g.tB = p ? kRes - p : kRes / 2;
    // residue t = i + R i' holds the outputs of row i' of sub-FFT i
    g.zA = kSI * (g.tA & (R - 1)) + kS16 * (g.tA / R);
    g.zB = kSI * (g.tB & (R - 1)) + kS16 * (g.tB / R);
    g.i1 = p & (R - 1); g.q2 = p / R;
    g.tAf = (float)g.tA; g.tBf = (float)g.tB;
    return g;
}

// Twiddle the outputs of a radix-RR butterfly by W^{b i} and store output i at zo[STRIDE i].
// log2 RR exact values come from the table rows tab[l][b] = W^{b 2^l}, the others are products
// (<= 3 multiplies deep): two fewer shared-memory wavefronts per product.
template <int RR, int STRIDE>
__device__ __forceinline__ void twiddle_store_rows(float2 (&v)[RR], const float2* tab, int b, float2* zo) {
    constexpr int kLog = RR == 1 ? 0 : RR == 2 ? 1 : RR == 4 ? 2 : RR == 8 ? 3 : 4;
    float2 tw[RR];
    static_for<kLog>([&](auto lc) {
        constexpr int l = decltype(lc)::value;
        tw[1 << l] = tab[l * 256 + b];
    });
    static_for<RR>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int hb = i >= 8 ? 8 : i >= 4 ? 4 : i >= 2 ? 2 : 1;     // highest set bit
        if constexpr (i > hb && i >= 3) tw[i] = cmul2(tw[i - hb], tw[hb]);
    });
    zo[0] = v[oR<RR>(0)];
#pragma unroll
    for (int i = 1; i < RR; ++i) zo[STRIDE * i] = cmul2(v[oR<RR>(i)], tw[i]);
}

// Pass 2: sub-FFTs of length 256, butterflies (i1, q2), (i1, q2 + 8), software-pipelined by
// hand: the second butterfly's loads fly while the first computes.
// n_fft = 256 has no pass 1 (R = 1): its two butterflies take their inputs straight from the tile,
// z[n] = x[n] (1 + j th'[n]) at n = q2 + 8 u, u even for the first butterfly, odd for the second
// (xs, thw given), and only their outputs go through the Z buffer.
template <int kSI, int kS16>
__device__ __forceinline__ void pass2(float2* Zb, const float2* T2, const Geom& g, const float* xs = nullptr,
                                      const float (*thw)[1] = nullptr) {
    float2* bz0 = Zb + kSI * g.i1 + g.q2;
    float2* bz1 = bz0 + 8;
    float2 v0[16], v1[16];
    if (xs) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float x0 = xs[g.q2 + 16 * j], x1 = xs[g.q2 + 8 + 16 * j];
            v0[j] = make_float2(x0, x0 * thw[2 * j][0]);
            v1[j] = make_float2(x1, x1 * thw[2 * j + 1][0]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v0[j] = bz0[kS16 * j];
#pragma unroll
        for (int j = 0; j < 16; ++j) v1[j] = bz1[kS16 * j];
    }
    // twiddles W_256^{p2 i}: the 16 of a butterfly are contiguous, two per 128-bit load (the whole
    // warp reads two addresses: one wavefront per load); loaded before the butterfly they belong
    // to, so that their latency hides behind its arithmetic
    auto twiddle_load = [&](float4 (&t)[8], int p2) {
        const float4* t4 = reinterpret_cast<const float4*>(T2 + kT2S * p2);
#pragma unroll
        for (int h = 0; h < 8; ++h) t[h] = t4[h];
    };
    auto twiddle_store = [&](float2 (&v)[16], const float4 (&t)[8], float2* base) {
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            if (h == 0) base[0] = v[o16(0)];
            else base[kS16 * (2 * h)] = cmul2(v[o16(2 * h)], make_float2(t[h].x, t[h].y));
            base[kS16 * (2 * h + 1)] = cmul2(v[o16(2 * h + 1)], make_float2(t[h].z, t[h].w));
        }
    };
    float4 tw[8];
    dft16(v0);
    twiddle_load(tw, g.q2);
    twiddle_store(v0, tw, bz0);
    dft16(v1);
    twiddle_load(tw, g.q2 + 8);
    twiddle_store(v1, tw, bz1);
}

// Pass 3: residues tA, tB; untangle in registers; 2 X to shared memory (Xs[k + 1]).
//   2 X[k] = Z[k] + conj Z[N-k],  2 X_th'[k] = (Z[k] - conj Z[N-k]) / j
// Thread 0's residues (0 and 8R) pair with themselves; it picks its conjugate partners with
// selects inside the common instruction stream (a divergent branch here would put both paths
// on the critical warp of the worker, and the other warps wait for it at the next barrier).
// Residue 0 has a ninth bin, N/2: thread 0 untangles it on the side (Xs, Sc[0]).
template <int R, bool XZ = false>
__device__ __forceinline__ void pass3_untangle(float2* Zb, float2* Xs, float2* Sc, const Geom& g,
                                               float2 (&xa)[8], float2 (&xb)[8], float2 (&ta)[8], float2 (&tb)[8]) {
    constexpr int kRes = 16 * R, N = 256 * R;
    const int tA = g.tA, tB = g.tB;
    float2 za[16], zb[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { za[j] = Zb[g.zA + j]; zb[j] = Zb[g.zB + j]; }
    dft16(za); dft16(zb);
    const bool self = g.p == 0;
    static_for<8>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        const float2 za_c = za[o16(c)], zb_c = zb[o16(c)];
        const float2 pa = self ? za[o16((16 - c) & 15)] : zb[o16(15 - c)];      // Z[N - (tA + kRes c)]
        const float2 pb = self ? zb[o16(15 - c)] : za[o16(15 - c)];              // Z[N - (tB + kRes c)]
        const float2 zb_n = cj(pa), za_n = cj(pb);
        xa[c] = za_c + zb_n; ta[c] = mulmj(za_c - zb_n);
        xb[c] = zb_c + za_n; tb[c] = mulmj(zb_c - za_n);
        if constexpr (XZ) { Zb[g.zA + c] = xa[c]; Zb[g.zB + c] = xb[c]; }       // in place: rows this thread has just read
        else { Xs[1 + tA + kRes * c] = xa[c]; Xs[1 + tB + kRes * c] = xb[c]; }
    });
    if (g.p == 1) {      // Hermitian mirrors: X[-1] = conj X[1], X[N/2+1] = conj X[N/2-1]
        if constexpr (XZ) { Sc[2] = cj(xa[0]); Sc[3] = cj(xb[7]); }       // (not in the Z rows: other threads still read them)
        else { Xs[0] = cj(xa[0]); Xs[N / 2 + 2] = cj(xb[7]); }
    }
    if (self) {
        const float2 z = za[o16(8)], zn = cj(z);
        if constexpr (XZ) Zb[g.zA + 8] = z + zn; else Xs[1 + N / 2] = z + zn;
        Sc[0] = mulmj(z - zn);
    }
}

// Epilogue of frame f of channel ch on the thread's own bins (after the barrier that makes X visible).
// Bins go four at a time (two of each residue): their neighbour loads and short dependent chains
// interleave, and one vote decides whether the warp keeps anything of the four (+5.6 % over a vote
// per bin; gating all 16 first costs registers and is slower).
// `active` is false for the lanes of a sub-warp worker whose frame slot is past the end of the tile:
// they tag along (the warp votes) and emit nothing.
template <int R, int MODE, bool XZ = false>
__device__ __forceinline__ void epilogue(const StftArgs& a, int ch, long long f, const float2* Xs, const float2* Sc,
                                         const Geom& g, const float2 (&xa)[8], const float2 (&xb)[8],
                                         const float2 (&ta)[8], const float2 (&tb)[8], bool active = true) {
    constexpr int kRes = 16 * R, N = 256 * R, B = N / 2 + 1, kWT = 8 * R;
    // XZ: Xs is the Z buffer; the neighbours of bin t + 256 c sit at zrow(t -+ 1) + c (residue 0's lower
    // neighbour is residue 255 one c down, residue 255's upper neighbour is residue 0 one c up)
    const float2* nAm = Xs; const float2* nAp = Xs; const float2* nBm = Xs; const float2* nBp = Xs;
    if constexpr (XZ) {
        nAm = Xs + (g.tA ? zrow_xz<R>(g.tA - 1) : zrow_xz<R>(kRes - 1) - 1);
        nAp = Xs + zrow_xz<R>(g.tA + 1);
        nBm = Xs + zrow_xz<R>(g.tB - 1);
        nBp = Xs + (g.tB < kRes - 1 ? zrow_xz<R>(g.tB + 1) : zrow_xz<R>(0) + 1);
    }
    const bool owner = active;
    FrameCtx fc;
    fc.lo = (float)max(-f, -1048576LL);
    fc.hi = (float)min(a.F - 1 - f, 1048576LL);
    fc.f = f; fc.ch = ch;
    const long long row0 = ((a.ring ? 0 : (long long)ch * a.F) + f) * B;
    fc.pd = a.dt_cols + row0; fc.pk = a.dk_bins + row0; fc.pe = a.energy + row0;
    const int tA = g.tA, tB = g.tB;
    // row pointers at the thread's two residues, opaque to the compiler so that it keeps them
    // in registers (it otherwise rebuilds a 64-bit address for every store: 7 IADD3 per bin)
    FrameCtx fA = fc, fB = fc;
    if (MODE == kStorePoints) {
        fA.pd += tA; fA.pk += tA; fA.pe += tA; fB.pd += tB; fB.pk += tB; fB.pe += tB;
        asm volatile("" : "+l"(fA.pd), "+l"(fA.pk), "+l"(fA.pe), "+l"(fB.pd), "+l"(fB.pk), "+l"(fB.pe));
    }
    constexpr int GC = 2;     // residue pairs per vote: 4 bins, 8 neighbour loads in flight (8 bins spill: measured -10 %)
    // Store mode, software-pipelined by hand: the neighbour loads of the next four bins are issued
    // before the vote and the stores of the current four.  The slow path then reads its two
    // neighbours again (X is still in shared memory), so the prefetch costs no registers (+1.2 %).
    // The deposit modes have no stores to hide the loads behind and keep them in program order.
    constexpr bool kPrefetch = MODE == kStorePoints;
    const float gate_p = gate_power<N>(a);
    float2 nm[2 * GC], np_[2 * GC];
    auto load_group = [&](int c0) {
#pragma unroll
        for (int i = 0; i < GC; ++i) {
            const int kA = tA + kRes * (c0 + i), kB = tB + kRes * (c0 + i);
            if constexpr (XZ) {
                nm[2 * i] = (c0 + i == 0 && g.tA == 0) ? Sc[2] : nAm[c0 + i]; np_[2 * i] = nAp[c0 + i];
                nm[2 * i + 1] = nBm[c0 + i]; np_[2 * i + 1] = nBp[c0 + i];
            } else {
                nm[2 * i] = Xs[kA]; np_[2 * i] = Xs[kA + 2];
                nm[2 * i + 1] = Xs[kB]; np_[2 * i + 1] = Xs[kB + 2];
            }
        }
    };
    if (kPrefetch) load_group(0);
    static_for<8 / GC>([&](auto cc) {
        constexpr int c0 = GC * decltype(cc)::value;
        float2 A2[2 * GC];
        bool lv[2 * GC];
        bool any = false;
        if (!kPrefetch) load_group(c0);
        float2 cm[2 * GC], cp[2 * GC];          // the current four's neighbours for the slow path (deposit modes)
#pragma unroll
        for (int i = 0; i < 2 * GC; ++i) { cm[i] = nm[i]; cp[i] = np_[i]; }
#pragma unroll
        for (int i = 0; i < GC; ++i) {
            A2[2 * i] = hann_stencil(xa[c0 + i], nm[2 * i], np_[2 * i]);                  // 4 X_h
            A2[2 * i + 1] = hann_stencil(xb[c0 + i], nm[2 * i + 1], np_[2 * i + 1]);
            lv[2 * i] = bin_power(A2[2 * i]) > gate_p;
            lv[2 * i + 1] = bin_power(A2[2 * i + 1]) > gate_p;
            any = any || lv[2 * i] || lv[2 * i + 1];
        }
        if constexpr (kPrefetch && c0 + GC < 8) load_group(c0 + GC);
        if (!__any_sync(0xffffffffu, any)) {
#pragma unroll
            for (int i = 0; i < GC; ++i) {
                bin_dead<MODE>(fA, owner, kRes * (c0 + i));
                bin_dead<MODE>(fB, owner, kRes * (c0 + i));
            }
        } else {
            // something of the four is kept: the slow path, bin by bin; in the deposit modes only
            // for the bins some lane keeps (the warps of a worker meet at the next barrier: what
            // one of them spends on atomics here, the other three wait; +3.9 % on the image path)
#pragma unroll
            for (int i = 0; i < GC; ++i) {
                const int kA = tA + kRes * (c0 + i), kB = tB + kRes * (c0 + i);
                if (MODE == kStorePoints || __any_sync(0xffffffffu, lv[2 * i]))
                    bin_tail<N, MODE>(a, fA, owner, lv[2 * i], kA, kRes * (c0 + i), g.tAf + (float)(kRes * (c0 + i)),
                                      A2[2 * i], kPrefetch ? (XZ ? ((c0 + i == 0 && g.tA == 0) ? Sc[2] : nAm[c0 + i]) : Xs[kA]) : cm[2 * i],
                                      kPrefetch ? (XZ ? nAp[c0 + i] : Xs[kA + 2]) : cp[2 * i], ta[c0 + i]);
                else
                    bin_dead<MODE>(fA, owner, kRes * (c0 + i));
                if (MODE == kStorePoints || __any_sync(0xffffffffu, lv[2 * i + 1]))
                    bin_tail<N, MODE>(a, fB, owner, lv[2 * i + 1], kB, kRes * (c0 + i), g.tBf + (float)(kRes * (c0 + i)),
                                      A2[2 * i + 1], kPrefetch ? (XZ ? nBm[c0 + i] : Xs[kB]) : cm[2 * i + 1],
                                      kPrefetch ? (XZ ? nBp[c0 + i] : Xs[kB + 2]) : cp[2 * i + 1], tb[c0 + i]);
                else
                    bin_dead<MODE>(fB, owner, kRes * (c0 + i));
            }
        }
    });
    // bin N/2, the ninth of residue 0: thread 0 of the frame (its warp tags along)
    if (kWT < 32 || g.p < 32) {
        if constexpr (XZ)
            bin_emit<N, MODE>(a, fc, owner && g.p == 0, N / 2, (float)(N / 2), Xs[zrow_xz<R>(0) + 8],
                              Xs[zrow_xz<R>(kRes - 1) + 7], Sc[3], Sc[0]);
        else
            bin_emit<N, MODE>(a, fc, owner && g.p == 0, N / 2, (float)(N / 2), Xs[N / 2 + 1], Xs[N / 2], Xs[N / 2 + 2], Sc[0]);
    }
}

// ---------------------------------------------------------------- in-kernel post-pass (fused mode)
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint8_t colour_index_of(float E, float db_floor, float inv_range, float gate_db) {
    if (!(E > 0.f)) return 0;                                     // same arithmetic as scatter_post.cuh::colour_index
    const float db = 10.0f * log10f(E);
    if (db < gate_db) return 0;
    const float v = rintf((db - db_floor) * inv_range);
    return (uint8_t)fminf(fmaxf(v, 0.f), 255.f);
}

// Shapes columns [c0, c1) of channel ch from the ring into index / grid and clears them.  Called by
// all `nth` threads of one worker (whole warps); a warp takes a column at a time.  Cells are read
// past L1 (other SMs deposited them with reds at L2).  Clean 64-row blocks are neither read nor
// cleared; their outputs are zeros.  Lanes own four consecutive rows, shifted so that the four
// index bytes of a lane are one aligned 32-bit store.
template <int MODE>
__device__ __noinline__ void fused_post_block(void* acc, unsigned char* flags, const FusedPost fp, long long F,
                                              int rows, int ch, long long c0, long long c1, int t, int nth) {
    const int lane = t & 31, wp = t >> 5, nwp = nth >> 5;
    if (fp.debug & 4) {          // timing experiment: zero-fill only, the smallest possible loop
        for (long long c = c0 + wp; c < c1; c += nwp) {
            uint8_t* ix = fp.index + ((long long)ch * F + c) * rows;
            for (int r = lane; r < rows; r += 32) ix[r] = 0;
        }
        return;
    }
    for (long long c = c0 + wp; c < c1; c += nwp) {
        const long long slot = ((long long)ch * F + c) & (long long)(fp.vring - 1);
        unsigned char* fl = flags + slot * fp.NB;
        // dirty mask of the column's (<= 64) row blocks
        const unsigned m0 = __ballot_sync(0xffffffffu, lane < fp.NB && __ldcg(fl + lane) != 0);
        const unsigned m1 = __ballot_sync(0xffffffffu, lane + 32 < fp.NB && __ldcg(fl + lane + 32) != 0);
        const unsigned long long dirty = (unsigned long long)m0 | ((unsigned long long)m1 << 32);
        const long long orow = ((long long)ch * F + c) * rows;
        uint8_t* ix = fp.index ? fp.index + orow : nullptr;
        float* gr = fp.grid ? fp.grid + orow : nullptr;
        const int a0 = ix ? (int)((4 - ((unsigned long long)ix & 3ull)) & 3ull) : 0;      // rows before the first aligned word
        for (int r0 = a0 + 4 * lane - 128; r0 < rows; r0 += 128) {
            // first iteration: the lanes whose word would start below row 0 take the a0 head rows one byte each
            float G[4];
            bool in[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int r = r0 + q;
                if (r0 < 0) r = (r0 + 128 == a0 + 4 * lane && q == 0 && lane < a0) ? lane : -1;   // head rows 0 .. a0-1
                in[q] = r >= 0 && r < rows;
                G[q] = 0.f;
                if (in[q] && ((dirty >> (r >> kFlagShift)) & 1ull)) {
                    if (MODE == kDepositU64) {
                        unsigned long long* p = reinterpret_cast<unsigned long long*>(acc) + slot * rows + r;
                        const unsigned long long v = __ldcg(p);
                        if (v) { __stcg(p, 0ull); G[q] = __double2float_rn(__ull2double_rn(v) * kFixScaleInv); }
                    } else {
                        float* p = reinterpret_cast<float*>(acc) + slot * rows + r;
                        const float v = __ldcg(p);
                        if (v != 0.f) { __stcg(p, 0.f); G[q] = v; }
                    }
                }
            }
            if (r0 < 0) {
                if (in[0]) {
                    if (ix) ix[lane] = colour_index_of(G[0] * __ldg(fp.weight + lane), fp.db_floor, fp.inv_range, fp.gate_db);
                    if (gr) gr[lane] = G[0];
                }
                continue;
            }
            if (ix) {
                unsigned word = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (in[q] && G[q] > 0.f)
                        word |= (unsigned)colour_index_of(G[q] * __ldg(fp.weight + r0 + q), fp.db_floor, fp.inv_range, fp.gate_db) << (8 * q);
                if (in[3]) *reinterpret_cast<unsigned*>(ix + r0) = word;
                else {
#pragma unroll
                    for (int q = 0; q < 3; ++q) if (in[q]) ix[r0 + q] = (uint8_t)(word >> (8 * q));
                }
            }
            if (gr) {
#pragma unroll
                for (int q = 0; q < 4; ++q) if (in[q]) gr[r0 + q] = G[q];
            }
        }
        __syncwarp();
        if (lane < fp.NB && ((dirty >> lane) & 1ull)) __stcg(fl + lane, (unsigned char)0);
        if (lane + 32 < fp.NB && ((dirty >> (lane + 32)) & 1ull)) __stcg(fl + lane + 32, (unsigned char)0);
    }
}

// ---------------------------------------------------------------- n_fft = 256 .. 4096
template <int R, int MODE>
__global__ void __launch_bounds__(Cfg<R>::kCta, 1)
stft_reassign_r16(const StftArgs a_in, const int tile_T) {
    using C = Cfg<R>;
    constexpr int kThreads = C::kCta;          // (shadows the namespace constant: 512 in the four-worker build of R = 16)
    constexpr bool kW4 = C::kW4;
    constexpr int N = C::N, kWT = C::kWT, kWorkers = C::kWorkers, kZBuf = C::kZBuf,
                  kXBuf = C::kXBuf, kZtab = C::kZtab, kTileFloats = C::kTileFloats, kSI = C::kSI,
                  kS16 = C::kS16, kU = C::kU, kG = C::kG, kSlot = C::kSlot, kWTh = C::kWorkerThreads;
    StftArgs a = a_in;
    if (!stream_decode(a)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    float2* Ztab = sm;                         // [log2 R][256]: rows i = 1, 2, 4, ..
    float2* T2 = Ztab + kZtab;                 // [16][16]
    float2* wbuf = T2 + kT2;                   // per worker: Z, X, scratch
    float* thwS = reinterpret_cast<float*>(wbuf + kWorkers * kG * kSlot);    // kW4: th'[n] table
    float* tile0 = thwS + (kW4 ? N : 0);                                     // 2 x kTileFloats

    const int tid = threadIdx.x;
    const int w = tid / kWTh;                  // worker
    const int wl = tid - w * kWTh;             // thread of the worker
    const int gs = kG > 1 ? wl / kWT : 0;      // frame slot of this thread (sub-warp workers)
    // role of this thread for its frame; rotated by one warp per worker so that the warp
    // carrying the self-paired bins lands on a different scheduler in each worker
    const int p = kG > 1 ? wl % kWT : (tid + 32 * w) & (kWT - 1);
    float2* Zb = wbuf + (w * kG + gs) * kSlot;
    float2* Xs = kW4 ? Zb : Zb + kZBuf;        // 2 X[k] at Xs[k + 1] (kW4: in place in the Z buffer, at zrow(k & 255) + (k >> 8))
    float2* Sc = Zb + kZBuf + kXBuf;

    // ---- twiddle tables (once per CTA)
    for (int e = tid; e < kZtab; e += kThreads) { const int i = 1 << (e / 256), b = e % 256; Ztab[e] = __ldg(&a.tw[b * i]); }
    for (int e = tid; e < 256; e += kThreads) { const int q = e / 16, i = e % 16; T2[kT2S * q + i] = __ldg(&a.tw[R * q * i]); }

    // ---- per-thread constants
    // th'[n] at this thread's 32 pass-1 samples n = p + kWT u + 256 j: frame-independent, so they
    // live in registers for the whole persistent loop
    float thw[kU][R];
    if constexpr (kW4) {
        for (int e = tid; e < N; e += kThreads) thwS[e] = __ldg(&a.thw[e]);
    } else {
#pragma unroll
        for (int u = 0; u < kU; ++u)
#pragma unroll
            for (int j = 0; j < R; ++j) thw[u][j] = __ldg(&a.thw[p + kWT * u + 256 * j]);
    }
    const Geom g = make_geom<R, kSI, kS16>(p);

    const long long per_ch = a.f_end - a.f_begin;
    const long long tiles_per_ch = (per_ch + tile_T - 1) / tile_T;
    const long long n_tiles = tiles_per_ch * a.channels;

    // Tiles are double-buffered and handed over without any CTA-wide barrier, so the
    // workers drift apart and their FP-heavy and shared-memory-heavy phases interleave (in
    // lockstep they collide: measured +13 %).  Protocol per buffer b:
    //   * a worker that has read its last sample of the tile in b bumps done[b]; the last one
    //     to do so refills b with the tile after next (4-byte cp.async from its own threads: no
    //     alignment demands; one TMA bulk copy for large aligned hops) and the copies arrive on
    //     the mbarrier full[b];
    //   * a worker waits on full[b] before it reads a refilled buffer.
    unsigned long long* full = reinterpret_cast<unsigned long long*>(tile0 + 2 * kTileFloats);   // [2]
    unsigned* done = reinterpret_cast<unsigned*>(full + 2);                                     // [2]
    int* refill = reinterpret_cast<int*>(done + 2);                                             // [kWorkers]
    unsigned* fin = reinterpret_cast<unsigned*>(refill + kWorkers);                             // [4] fused mode: workers done with tile ti & 3
    int* go = reinterpret_cast<int*>(fin + 4);                                                  // [kWorkers] fused mode: this worker shapes a block
    const unsigned full_sm = (unsigned)__cvta_generic_to_shared(full);
    auto tile_geom = [&](long long tl, int& ch, long long& f0, int& nf) {
        ch = (int)(tl / tiles_per_ch);
        f0 = a.f_begin + (tl - (long long)ch * tiles_per_ch) * tile_T;
        nf = (int)min((long long)tile_T, a.f_end - f0);
    };
    auto copy_tile = [&](long long tl, float* dst, int t0, int nth) {
        int ch, nf; long long f0;
        tile_geom(tl, ch, f0, nf);
        const int n_samp = (nf - 1) * a.hop + N;
        const float* src = a.pcm + (long long)ch * a.S + f0 * a.hop + a.samp_off;
        const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst);
        for (int s = t0; s < n_samp; s += nth)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + 4u * s), "l"(src + s) : "memory");
    };
    // prologue: the first two tiles by all threads
    const long long tile_step = gridDim.x;
    if ((long long)blockIdx.x < n_tiles) copy_tile(blockIdx.x, tile0, tid, kThreads);
    if ((long long)blockIdx.x + tile_step < n_tiles) copy_tile(blockIdx.x + tile_step, tile0 + kTileFloats, tid, kThreads);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full_sm), "r"(kWTh));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full_sm + 8), "r"(kWTh));
        done[0] = 0; done[1] = 0;
        fin[0] = 0; fin[1] = 0; fin[2] = 0; fin[3] = 0;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // (no initial stagger: measured, it makes no difference)

    // ---- fused mode (a.fp.vring > 0, deposit modes): ring space before a tile, completion after it
    const bool fused = EMS_FUSED_POST && MODE != kStorePoints && a.fp.vring > 0;
    const long long Rcols = (N / 2 + a.hop - 1) / a.hop;            // a deposit lands within +-Rcols columns of its frame
    // the ring slots tile tl is about to deposit into were last used Rg columns earlier: wait until the
    // blocks that owned them have been shaped and cleared (rare: the ring is several rounds of tiles long)
    auto ring_wait = [&](long long tl, int ch, long long f0, int nf) {
        if (wl == 0 && !(a.fp.debug & 2)) {
            const long long lo = max(f0 - Rcols, 0LL), hi = min(f0 + nf - 1 + Rcols, a.F - 1);
            const long long vlo = (long long)ch * a.F + lo - a.fp.vring, vhi = (long long)ch * a.F + hi - a.fp.vring;
            if (vhi >= 0) {
                const long long v0 = max(vlo, 0LL);
                const long long b0 = (v0 / a.F) * tiles_per_ch + (v0 % a.F) / tile_T;
                const long long b1 = (vhi / a.F) * tiles_per_ch + (vhi % a.F) / tile_T;
                for (long long b = b0; b <= b1; ++b)
                    while (ld_acquire(a.fp.done + b) != a.fp.epoch) __nanosleep(200);
            }
        }
        worker_bar<kWT>(w);
    };
    // this worker has deposited its last point of tile tl (tile number ti of this CTA).  The last worker
    // of the CTA to get here tells the blocks the tile can reach, and shapes those that were waiting
    // for this tile only.
    auto tile_done = [&](long long tl, long long ti, int ch) {
        __threadfence();                              // this thread's reds are performed
        worker_bar<kWT>(w);
        if (wl == 0) {
            const unsigned old = atomicAdd(&fin[ti & 3], 1u);
            const bool last = old == kWorkers - 1;
            if (last) fin[ti & 3] = 0;
            go[w] = last ? 1 : 0;
        }
        worker_bar<kWT>(w);
        if (!go[w]) return;
        const long long j = tl - (long long)ch * tiles_per_ch;           // tile of its channel
        const long long jlo = max(j - a.fp.nb, 0LL), jhi = min(j + a.fp.nb, tiles_per_ch - 1);
        for (long long b = jlo; b <= jhi; ++b) {
            worker_bar<kWT>(w);                       // go[w] of the previous round has been read
            if (wl == 0) {
                const int need = (int)(min(b + a.fp.nb, tiles_per_ch - 1) - max(b - a.fp.nb, 0LL) + 1);
                int* rp = a.fp.ready + (long long)ch * tiles_per_ch + b;
                const int got = atomicAdd(rp, 1) + 1;
                if (got == need) *rp = 0;
                go[w] = got == need ? 1 : 0;
            }
            worker_bar<kWT>(w);
            if (go[w]) {
                __threadfence();                      // acquire: the other tiles' reds, fenced before their arrival
                const long long c0 = a.f_begin + b * tile_T, c1 = min(c0 + tile_T, a.f_end);
                if (!(a.fp.debug & 1)) fused_post_block<MODE>(a.acc, a.flags, a.fp, a.F, a.rows, ch, c0, c1, wl, kWTh);
                __threadfence();
                worker_bar<kWT>(w);
                if (wl == 0)
                    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.fp.done + (long long)ch * tiles_per_ch + b), "r"(a.fp.epoch) : "memory");
            }
        }
    };

    // this worker is done reading buffer b (tile number ti of this CTA): called by its thread
    // p == 0 after a worker barrier; tells the worker whether it has to refill the buffer
    auto release_tile = [&](int b) {
        __threadfence_block();
        const unsigned old = atomicAdd(&done[b], 1u);
        const bool last = old == kWorkers - 1;
        if (last) done[b] = 0;
        __threadfence_block();
        refill[w] = last ? 1 : 0;
    };
    // all threads of the refilling worker
    auto refill_tile = [&](int b, long long ti_next) {
        const long long tl2 = blockIdx.x + ti_next * tile_step;
        if (tl2 >= n_tiles) return;
        // Tiles that turn over quickly (hop >= N/8) and start on a 16-byte boundary come in as one
        // TMA bulk copy issued by a single thread (+3..6 % over per-thread copies at hop = N/4).
        // With a small hop the copy volume is negligible and the slower 4-byte loop is kept on
        // purpose: the delay it costs the refilling worker keeps the workers out of phase
        // (4096/128: 97.6 M frames/s, 94.2 M with a bulk copy, 97.7 M with a bulk copy and a 1 us
        // sleep — measured with tools/ab_kernel.py).
        if (a.hop * 8 >= N) {
            int ch2, nf2; long long f02;
            tile_geom(tl2, ch2, f02, nf2);
            const unsigned bytes = 4u * (unsigned)((nf2 - 1) * a.hop + N);
            const float* src = a.pcm + (long long)ch2 * a.S + f02 * a.hop + a.samp_off;
            if ((((unsigned long long)src | (unsigned long long)(unsigned)a.hop * 4ull) & 15ull) == 0) {
                const unsigned bar = full_sm + 8u * b;
                if (wl == 0) {
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
        }
        copy_tile(tl2, tile0 + b * kTileFloats, wl, kWTh);
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full_sm + 8u * b) : "memory");
    };

    for (long long ti = 0;; ++ti) {
        const long long tl = blockIdx.x + ti * tile_step;
        if (tl >= n_tiles) break;
        const int buf = (int)(ti & 1);
        if (ti >= 2) {                      // refilled buffer: wait until the copies have landed
            const unsigned parity = (unsigned)(((ti >> 1) - 1) & 1);
            unsigned ok = 0;
            while (!ok)
                asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                             : "=r"(ok) : "r"(full_sm + 8u * buf), "r"(parity) : "memory");
        }
        int ch, nf; long long f0;
        tile_geom(tl, ch, f0, nf);
        const float* tile = tile0 + buf * kTileFloats;
        if (w * kG >= nf) {                 // no frame for this worker in a short tile: just release it
            if (wl == 0) release_tile(buf);
            worker_bar<kWT>(w);
            if (refill[w]) refill_tile(buf, ti + 2);
            worker_bar<kWT>(w);
            if (fused) tile_done(tl, ti, ch);
            continue;
        }
        if (fused) ring_wait(tl, ch, f0, nf);

        for (int fi0 = w * kG; fi0 < nf; fi0 += kWorkers * kG) {
            // sub-warp workers: slots past the end of the tile redo its last frame and emit nothing
            const bool active = kG == 1 || fi0 + gs < nf;
            const int fi = kG == 1 ? fi0 : min(fi0 + gs, nf - 1);
            const float* xs = tile + fi * a.hop;
            const long long f = f0 + fi;
            const bool last_frame = fi0 + kWorkers * kG >= nf;

            if constexpr (R == 1) {
                // no pass 1: pass 2 reads the tile; the tile is released after it
                pass2<kSI, kS16>(Zb, T2, g, xs, thw);
                worker_bar<kWT>(w);
                if (last_frame && wl == 0) release_tile(buf);
                float2 xa[8], xb[8], ta[8], tb[8];
                pass3_untangle<R>(Zb, Xs, Sc, g, xa, xb, ta, tb);
                worker_bar<kWT>(w);
                if (last_frame && refill[w]) refill_tile(buf, ti + 2);
                epilogue<R, MODE>(a, ch, f, Xs, Sc, g, xa, xb, ta, tb, active);
            } else {
            // ================= pass 1: radix-R butterflies b = p + kWT u on z[n] = x[n] (1 + j th'[n])
            // (all 32 samples of the thread are requested before the first butterfly starts)
            float xr[kU][R];
#pragma unroll
            for (int u = 0; u < kU; ++u)
#pragma unroll
                for (int j = 0; j < R; ++j) xr[u][j] = xs[p + kWT * u + 256 * j];   // n = b + 256 j
            // kW4: the Z buffer still holds the previous frame's X until every warp of the worker has left its epilogue
            if constexpr (kW4) worker_bar<kWT>(w);
            static_for<kU>([&](auto uc) {
                constexpr int u = decltype(uc)::value;
                const int b = p + kWT * u;
                float2 v[R];
#pragma unroll
                for (int j = 0; j < R; ++j)
                    v[j] = make_float2(xr[u][j], xr[u][j] * (kW4 ? thwS[p + kWT * u + 256 * j] : thw[u][j]));
                dftR(v);
                twiddle_store_rows<R, kSI>(v, Ztab, b, Zb + b + (b >> 4) * (kS16 - 16));
            });
            worker_bar<kWT>(w);
            if (last_frame && wl == 0) release_tile(buf);  // every sample of the tile has been read

            pass2<kSI, kS16>(Zb, T2, g);
            worker_bar<kWT>(w);
            if (last_frame && refill[w]) refill_tile(buf, ti + 2);

            float2 xa[8], xb[8], ta[8], tb[8];     // bins tA + kRes c and tB + kRes c, c = 0..7
            pass3_untangle<R, kW4>(Zb, Xs, Sc, g, xa, xb, ta, tb);
            worker_bar<kWT>(w);      // X visible; the Z buffer is free for the next frame's pass 1

            epilogue<R, MODE, kW4>(a, ch, f, Xs, Sc, g, xa, xb, ta, tb, active);
            }
        }
        if (fused) tile_done(tl, ti, ch);
    }
}

// ---------------------------------------------------------------- n_fft = 8192, 16384
// Same radix-16 passes and epilogue, but pass 1 (radix R = 16 R0) is split in two — a radix-R0
// pass over samples 4096 apart, read straight from global memory / L2 (with hop = n_fft/4 and
// frames this long a shared-memory tile has no room and little reuse to offer), then a radix-16
// pass — so a thread never holds more than 16 values.  4 worker barriers per frame.
// One CTA per SM: 2 workers x 256 threads (8192) or 1 x 512 (16384); frames are strided over
// CTAs first, then workers, so short launches (streaming: one frame per channel) spread over SMs.
template <int R0_>
struct CfgL {
    static_assert(R0_ == 2 || R0_ == 4, "radix of the pre-pass");
    static constexpr int R0 = R0_;
    static constexpr int R = 16 * R0;
    static constexpr int N = 256 * R;
    static constexpr int kWT = 8 * R;
    static constexpr int kThreads = 512;
    static constexpr int kWorkers = kThreads / kWT;
    static constexpr int kUA = 32 / R0;             // radix-R0 butterflies per thread
    static constexpr int kSI = 257, kS16 = 16;
    static constexpr int kZBuf = R * kSI;
    static constexpr int kXBuf = N / 2 + 4;
    static constexpr int kZtab = 4 * 256;           // W_4096^{b i}, i = 1, 2, 4, 8
    static constexpr int kSmemBytes = (kWorkers * (kZBuf + kXBuf + kScratch) + kZtab + kT2) * 8;
    static_assert(kSmemBytes <= kMaxSmem, "shared memory");
};

template <int R0, int MODE>
__global__ void __launch_bounds__(CfgL<R0>::kThreads, 1)
stft_reassign_r16_large(const StftArgs a_in) {
    using C = CfgL<R0>;
    constexpr int R = C::R, kWT = C::kWT, kWorkers = C::kWorkers, kZBuf = C::kZBuf, kXBuf = C::kXBuf,
                  kZtab = C::kZtab, kSI = C::kSI, kS16 = C::kS16, kUA = C::kUA, kThreads = C::kThreads;
    StftArgs a = a_in;
    if (!stream_decode(a)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    float2* Ztab = sm;                         // [4][256]: W_4096^{b 2^l}
    float2* T2 = Ztab + kZtab;                 // [16][16]
    float2* wbuf = T2 + kT2;

    const int tid = threadIdx.x;
    const int w = tid / kWT;
    const int p = (tid + 32 * w) & (kWT - 1);
    float2* Zb = wbuf + w * (kZBuf + kXBuf + kScratch);
    float2* Xs = Zb + kZBuf;
    float2* Sc = Xs + kXBuf;

    for (int e = tid; e < kZtab; e += kThreads) { const int i = 1 << (e / 256), b = e % 256; Ztab[e] = __ldg(&a.tw[R0 * b * i]); }
    for (int e = tid; e < 256; e += kThreads) { const int q = e / 16, i = e % 16; T2[kT2S * q + i] = __ldg(&a.tw[R * q * i]); }
    __syncthreads();
    const Geom g = make_geom<R, kSI, kS16>(p);
    constexpr bool kR32 = R0 == 2 && EMS_LARGE_R32;
    // radix-32 pass 1 (n_fft = 8192): W_N^{p 2^l}, l = 0..4, exact and frame-independent, in registers
    float2 wb[5];
    if constexpr (kR32) {
#pragma unroll
        for (int l = 0; l < 5; ++l) wb[l] = __ldg(&a.tw[(p << l) & (C::N - 1)]);
    }

    const long long per_ch = a.f_end - a.f_begin;
    const long long total = per_ch * a.channels;
    constexpr bool kPF = (EMS_LARGE_PREFETCH == 2 || (EMS_LARGE_PREFETCH == 1 && R0 == 4)) && !kR32;
    // Prefetch (kPF): the warp that owns butterflies b0..b0+31 of step u owns the 32 consecutive Z slots of rows
    // R0 (b0 >> 8) + j, j < R0 (256 bytes each), and is the only one to write them in pass A.  The next frame's
    // samples (4 bytes per lane and j, packed: two j per row block) and W_N^b (8 bytes per lane, row block R0/2)
    // are copied there while the epilogue runs; every lane reads back what it copied, and after a __syncwarp
    // the butterflies overwrite the block.  No barrier is added and pass A no longer waits on L2.
    const int lane = p & 31;
    auto stage = [&](int u) -> float* {
        const int b0 = (p & ~31) + kWT * u;
        return reinterpret_cast<float*>(Zb + kSI * (R0 * (b0 >> 8)) + (b0 & 255));
    };
    auto frame_ptr = [&](long long it) -> const float* {
        const int ch = (int)(it / per_ch);
        const long long f = a.f_begin + (it - (long long)ch * per_ch);
        return a.pcm + (long long)ch * a.S + f * a.hop + a.samp_off;
    };
    auto prefetch = [&](const float* xs) {
#pragma unroll
        for (int u = 0; u < kUA; ++u) {
            const unsigned sb = (unsigned)__cvta_generic_to_shared(stage(u));
#pragma unroll
            for (int j = 0; j < R0; ++j)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sb + 4u * (2 * kSI * (j >> 1) + lane + 32 * (j & 1))),
                             "l"(xs + p + kWT * u + 4096 * j) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sb + 4u * (2 * kSI * (R0 / 2) + 2 * lane)),
                         "l"(a.tw + p + kWT * u) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const long long it0 = blockIdx.x + (long long)gridDim.x * w, it_step = (long long)gridDim.x * kWorkers;
    if constexpr (kPF) { if (it0 < total) prefetch(frame_ptr(it0)); }
    for (long long it = it0; it < total; it += it_step) {
        const int ch = (int)(it / per_ch);
        const long long f = a.f_begin + (it - (long long)ch * per_ch);
        const float* xs = a.pcm + (long long)ch * a.S + f * a.hop + a.samp_off;

        if constexpr (kR32) {
            // ================= pass 1, n_fft = 8192 = 32 x 256: thread b = p takes z[n], n = b + 256 j, j = 0..31,
            // does the radix-32 butterfly in registers (even / odd j: two DFT-16, W_32 twiddles, radix 2),
            // multiplies output i by W_N^{b i} and stores it as element b of sub-FFT i: Zb[kSI i + b].
            // One exchange and one barrier fewer than radix 2 followed by radix 16.
            constexpr int N = C::N;
            float xr[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) xr[j] = __ldg(xs + p + 256 * j);
            float2 ev[16], od[16];
            {   // th'[n] = (n - N/2)(2/N)(0.5 - 0.5 cos(theta_b + j pi/16)): angle addition with constants
                const float cb = wb[0].x, sb = -wb[0].y, rb = (float)(p - N / 2) * (2.0f / N);
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
            float2* zo = Zb + p;
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
            worker_bar<kWT>(w);
        } else {
        // ================= pass A: radix-R0 butterflies b = p + kWT u over n = b + 4096 j0,
        // z[n] = x[n] (1 + j th'[n]); output i0 of butterfly b = b1 + 256 j1 goes to the slot
        // pass B reads it from: Zb[kSI (i0 + R0 j1) + b1]
        {
            float xr[kUA][R0];
            float2 w1s[kUA];                                   // W_N^b = (cos, -sin)(2 pi b / N)
            if constexpr (kPF) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
                for (int u = 0; u < kUA; ++u) {
                    const float* sb = stage(u);
#pragma unroll
                    for (int j = 0; j < R0; ++j) xr[u][j] = sb[2 * kSI * (j >> 1) + lane + 32 * (j & 1)];
                    w1s[u] = *reinterpret_cast<const float2*>(sb + 2 * kSI * (R0 / 2) + 2 * lane);
                }
                __syncwarp();                                  // every lane has its values before the block is overwritten
            } else {
#pragma unroll
            for (int u = 0; u < kUA; ++u) {
#pragma unroll
                for (int j = 0; j < R0; ++j) xr[u][j] = __ldg(xs + p + kWT * u + 4096 * j);
                w1s[u] = __ldg(&a.tw[p + kWT * u]);
            }
            }
            static_for<kUA>([&](auto uc) {
                constexpr int u = decltype(uc)::value;
                constexpr int N = C::N;
                const int b = p + kWT * u;
                const float2 w1 = w1s[u];
                // th'[n] = (n - N/2)(2/N)(0.5 - 0.5 cos(2 pi n/N)) at n = b + 4096 j without a table read:
                // cos(2 pi n/N) = cos(theta_b + 2 pi j/R0) is +-cos / +-sin of the twiddle already in hand,
                // and (n - N/2)(2/N) is exact in fp32
                const float bf = (float)b * (2.0f / N);
                float2 v[R0];
#pragma unroll
                for (int j = 0; j < R0; ++j) {
                    const float cs = R0 == 2 ? (j ? -w1.x : w1.x)
                                             : (j == 0 ? w1.x : j == 1 ? w1.y : j == 2 ? -w1.x : -w1.y);
                    const float t = (bf + (float)(4096 * j - N / 2) * (2.0f / N)) * fmaf(-0.5f, cs, 0.5f);
                    v[j] = make_float2(xr[u][j], xr[u][j] * t);
                }
                float2* zo = Zb + kSI * (R0 * (b >> 8)) + (b & 255);
                if constexpr (R0 == 2) {
                    zo[0] = v[0] + v[1];
                    zo[kSI] = cmul2(v[0] - v[1], w1);
                } else {
                    dft4(v[0], v[1], v[2], v[3]);
                    const float2 w2 = cmul2(w1, w1);
                    zo[0] = v[0];
                    zo[kSI] = cmul2(v[1], w1);
                    zo[2 * kSI] = cmul2(v[2], w2);
                    zo[3 * kSI] = cmul2(v[3], cmul2(w1, w2));
                }
            });
        }
        worker_bar<kWT>(w);

        // ================= pass B: R0 x 256 radix-16 butterflies (i0, b1) over j1, twiddle W_4096^{b1 i1}
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = p + kWT * u, b1 = e & 255, i0 = e >> 8;
            float2* zb = Zb + kSI * i0 + b1;
            float2 v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = zb[kSI * R0 * j];
            dft16(v);
            twiddle_store_rows<16, kSI * R0>(v, Ztab, b1, zb);
        }
        worker_bar<kWT>(w);
        }

        pass2<kSI, kS16>(Zb, T2, g);
        worker_bar<kWT>(w);

        float2 xa[8], xb[8], ta[8], tb[8];
        pass3_untangle<R>(Zb, Xs, Sc, g, xa, xb, ta, tb);
        worker_bar<kWT>(w);      // X visible; nobody reads Z any more
        if constexpr (kPF) { if (it + it_step < total) prefetch(frame_ptr(it + it_step)); }

        epilogue<R, MODE>(a, ch, f, Xs, Sc, g, xa, xb, ta, tb);
    }
}

// ---------------------------------------------------------------- n_fft = 32768
// The packed spectrum (256 KB) does not fit an SM, so the two real FFTs — x, then x th' — run one
// after the other as 16384-point complex FFTs of y[m] = r[2m] + j r[2m+1] (the passes of the
// 16384 kernel above), each followed by the even/odd split
//     2 R[k]     = 2E + W_N^k 2O,   2E = Y[k] + conj Y[M-k],  2O = -j (Y[k] - conj Y[M-k])
//     2 R[M - k] = conj(2E - W_N^k 2O)                                        (M = N/2)
// in the registers of the thread that holds residues t and 1024 - t.  2 X waits in a per-CTA
// scratch that stays in L2 (131 KB) and comes back into the (by then free) Z buffer for the
// neighbour reads of the Hann stencils; 2 X_th' never leaves the registers.
constexpr int k32kScratch = 16384 + 4;      // float2 per CTA: 2 X[k] at [k + 1], mirrors at [0], [M + 2]

template <int MODE>
__global__ void __launch_bounds__(512, 1)
stft_reassign_r16_32k(const StftArgs a_in, float2* __restrict__ scratch_all) {
    using C = CfgL<4>;
    constexpr int N = 32768, M = 16384, R0 = 4, R = 64, kWT = 512, kSI = C::kSI, kS16 = C::kS16,
                  kRes = 1024, kZtab = C::kZtab, kThreads = 512;
    StftArgs a = a_in;
    if (!stream_decode(a)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    float2* Ztab = sm;                         // [4][256]: W_4096^{b 2^l}
    float2* T2 = Ztab + kZtab;                 // [16][16]
    float2* Zb = T2 + kT2;                     // R * kSI = 16448
    float2* Xs = Zb;                           // 2 X[k] at Xs[k + 1], staged after the last pass of x th'
    float2* Sc = Zb + k32kScratch;             // 2 X_th' of the 33 self-paired bins (the Z buffer has 60 spare slots)
    float2* X2 = scratch_all + (size_t)blockIdx.x * k32kScratch;

    const int tid = threadIdx.x, p = tid;
    for (int e = tid; e < kZtab; e += kThreads) { const int i = 1 << (e / 256), b = e % 256; Ztab[e] = __ldg(&a.tw[8 * b * i]); }
    for (int e = tid; e < 256; e += kThreads) { const int q = e / 16, i = e % 16; T2[kT2S * q + i] = __ldg(&a.tw[128 * q * i]); }
    __syncthreads();
    const Geom g = make_geom<R, kSI, kS16>(p);
    const int tA = g.tA, tB = g.tB;
    const float2 wA = __ldg(&a.tw[tA]);          // W_N^tA (1 for thread 0)
    const float2 w512 = __ldg(&a.tw[512]);

    // even/odd split of the pair (Y[k], Y[M-k]) with W_N^k = w * W_32^c
    auto split = [](float2 y, float2 yn_conj, float2 wk, float2& rk, float2& rmk) {
        const float2 e2 = y + yn_conj, wo = cmul2(mulmj(y - yn_conj), wk);
        rk = e2 + wo;
        rmk = cj(e2 - wo);
    };

    const long long per_ch = a.f_end - a.f_begin;
    const long long total = per_ch * a.channels;
    for (long long it = blockIdx.x; it < total; it += gridDim.x) {
        const int ch = (int)(it / per_ch);
        const long long f = a.f_begin + (it - (long long)ch * per_ch);
        const float* xs = a.pcm + (long long)ch * a.S + f * a.hop + a.samp_off;

#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
            // ================= pass A: radix-4 butterflies b = p + 512 u over m = b + 4096 j0 on
            // y[m] = r[2m] + j r[2m+1], in two batches of four (registers)
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                float2 v[4][4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int n = 2 * (p + kWT * (4 * half + u) + 4096 * j);
                        v[u][j] = make_float2(__ldg(xs + n), __ldg(xs + n + 1));
                    }
                float2 w1s[4];                                          // W_M^b = W_N^{2b}
#pragma unroll
                for (int u = 0; u < 4; ++u) w1s[u] = __ldg(&a.tw[2 * (p + kWT * (4 * half + u))]);
                if (phase) {
                    // x th' without a table read: th'[n] = (n - N/2)(2/N)(0.5 - 0.5 cos(2 pi n/N)) at
                    // n = 2 (b + 4096 j) and n + 1; cos(2 pi n/N) = cos(theta_b + j pi/2) comes from the
                    // twiddle in hand, the odd sample is one rotation by 2 pi/N further
                    constexpr float Cd = 0.99999998161642933f, Sd = 1.9174759731070331e-4f;   // cos, sin(2 pi / 32768)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int b = p + kWT * (4 * half + u);
                        const float c = w1s[u].x, sn = -w1s[u].y;      // cos, sin(theta_b)
                        const float bf = (float)(2 * b) * (2.0f / N);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float cj_ = j == 0 ? c : j == 1 ? -sn : j == 2 ? -c : sn;     // cos(theta_b + j pi/2)
                            const float sj_ = j == 0 ? sn : j == 1 ? c : j == 2 ? -sn : -c;     // sin(theta_b + j pi/2)
                            const float nf = bf + (float)(8192 * j - N / 2) * (2.0f / N);
                            const float t0 = nf * fmaf(-0.5f, cj_, 0.5f);
                            const float t1 = (nf + 2.0f / N) * fmaf(-0.5f, fmaf(cj_, Cd, -sj_ * Sd), 0.5f);
                            v[u][j] = mul2(v[u][j], make_float2(t0, t1));
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int b = p + kWT * (4 * half + u);
                    float2* zo = Zb + kSI * (R0 * (b >> 8)) + (b & 255);
                    const float2 w1 = w1s[u];                          // W_M^b
                    dft4(v[u][0], v[u][1], v[u][2], v[u][3]);
                    const float2 w2 = cmul2(w1, w1);
                    zo[0] = v[u][0];
                    zo[kSI] = cmul2(v[u][1], w1);
                    zo[2 * kSI] = cmul2(v[u][2], w2);
                    zo[3 * kSI] = cmul2(v[u][3], cmul2(w1, w2));
                }
            }
            __syncthreads();

            // ================= pass B, pass 2: as in the 16384 kernel
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int e = p + kWT * u, b1 = e & 255, i0 = e >> 8;
                float2* zb = Zb + kSI * i0 + b1;
                float2 v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = zb[kSI * R0 * j];
                dft16(v);
                twiddle_store_rows<16, kSI * R0>(v, Ztab, b1, zb);
            }
            __syncthreads();
            pass2<kSI, kS16>(Zb, T2, g);
            __syncthreads();

            // ================= pass 3: Y[tA + 1024 c] = za[o16(c)], Y[tB + 1024 c] = zb[o16(c)]
            float2 za[16], zb[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { za[j] = Zb[g.zA + j]; zb[j] = Zb[g.zB + j]; }
            dft16(za); dft16(zb);

            if (phase == 0) {
                // ---- 2 X to the scratch
                if (p != 0) {
                    static_for<16>([&](auto cc) {
                        constexpr int c = decltype(cc)::value;
                        float2 rk, rmk;
                        split(za[o16(c)], cj(zb[o16(15 - c)]), cmul2(wA, make_float2(c32(c), -s32(c))), rk, rmk);
                        X2[1 + tA + kRes * c] = rk;
                        X2[1 + tB + kRes * (15 - c)] = rmk;
                        if (c == 0 && p == 1) {        // Hermitian mirrors X[-1], X[M+1]
                            X2[0] = cj(rk);
                            X2[M + 2] = cj(rmk);
                        }
                    });
                } else {
                    static_for<9>([&](auto cc) {       // residue 0: bins 1024 c and 1024 (16 - c)
                        constexpr int c = decltype(cc)::value;
                        float2 rk, rmk;
                        split(za[o16(c)], cj(za[o16((16 - c) & 15)]), make_float2(c32(c), -s32(c)), rk, rmk);
                        X2[1 + kRes * c] = rk;
                        if (c != 8) X2[1 + kRes * (16 - c)] = rmk;
                    });
                    static_for<8>([&](auto cc) {       // residue 512: bins 512 + 1024 c and 512 + 1024 (15 - c)
                        constexpr int c = decltype(cc)::value;
                        float2 rk, rmk;
                        split(zb[o16(c)], cj(zb[o16(15 - c)]), cmul2(w512, make_float2(c32(c), -s32(c))), rk, rmk);
                        X2[1 + 512 + kRes * c] = rk;
                        X2[1 + 512 + kRes * (15 - c)] = rmk;
                    });
                }
                __syncthreads();      // Z is free for the passes of x th'; the scratch writes are ordered before the reads below
            } else {
                __syncthreads();      // every thread holds its Y values: Z becomes the X buffer
                {
                    const unsigned d0 = (unsigned)__cvta_generic_to_shared(Xs);
                    for (int e = tid; e < k32kScratch / 2; e += kThreads)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + 16u * e), "l"(X2 + 2 * e) : "memory");
                    asm volatile("cp.async.commit_group;" ::: "memory");
                }
                if (p == 0) {         // 2 X_th' of the self-paired residues go through shared memory
                    static_for<9>([&](auto cc) {
                        constexpr int c = decltype(cc)::value;
                        float2 rk, rmk;
                        split(za[o16(c)], cj(za[o16((16 - c) & 15)]), make_float2(c32(c), -s32(c)), rk, rmk);
                        Sc[c] = rk;
                        if (c != 8) Sc[16 - c] = rmk;
                    });
                    static_for<8>([&](auto cc) {
                        constexpr int c = decltype(cc)::value;
                        float2 rk, rmk;
                        split(zb[o16(c)], cj(zb[o16(15 - c)]), cmul2(w512, make_float2(c32(c), -s32(c))), rk, rmk);
                        Sc[17 + c] = rk;
                        Sc[17 + 15 - c] = rmk;
                    });
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncthreads();

                // ---- epilogue: 2 X_th' of the thread's pairs, then the shared bin code
                constexpr int B = N / 2 + 1;
                FrameCtx fc;
                fc.lo = (float)max(-f, -1048576LL);
                fc.hi = (float)min(a.F - 1 - f, 1048576LL);
                fc.f = f; fc.ch = ch;
                const long long row0 = ((a.ring ? 0 : (long long)ch * a.F) + f) * B;
                fc.pd = a.dt_cols + row0; fc.pk = a.dk_bins + row0; fc.pe = a.energy + row0;
                static_for<8>([&](auto cc) {        // four bins per vote, as in epilogue()
                    constexpr int c = 2 * decltype(cc)::value;
                    float2 t2[4], xk[4], xm[4], xp[4], A2[4];
                    int kk[4];
                    bool lv[4], any = false;
                    split(za[o16(c)], cj(zb[o16(15 - c)]), cmul2(wA, make_float2(c32(c), -s32(c))), t2[0], t2[1]);
                    split(za[o16(c + 1)], cj(zb[o16(14 - c)]), cmul2(wA, make_float2(c32(c + 1), -s32(c + 1))), t2[2], t2[3]);
                    kk[0] = tA + kRes * c; kk[1] = tB + kRes * (15 - c);
                    kk[2] = tA + kRes * (c + 1); kk[3] = tB + kRes * (14 - c);
#pragma unroll
                    for (int i = 0; i < 4; ++i) { xm[i] = Xs[kk[i]]; xk[i] = Xs[kk[i] + 1]; xp[i] = Xs[kk[i] + 2]; }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        A2[i] = hann_stencil(xk[i], xm[i], xp[i]);
                        lv[i] = bin_power(A2[i]) > gate_power<N>(a);
                        any = any || lv[i];
                    }
                    if (!__any_sync(0xffffffffu, any)) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) bin_dead<MODE>(fc, p != 0, kk[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            bin_tail<N, MODE>(a, fc, p != 0, lv[i], kk[i], kk[i], (float)kk[i], A2[i], xm[i], xp[i], t2[i]);
                    }
                });
                if (p < 32) {         // the 33 self-paired bins: lane l takes bins l and (lane 0) 32 of the list
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int ls = min(p + 32 * r, 32);
                        const int ks = ls <= 16 ? kRes * ls : 512 + kRes * (ls - 17);
                        bin_emit<N, MODE>(a, fc, p + 32 * r <= 32, ks, (float)ks, Xs[ks + 1], Xs[ks], Xs[ks + 2], Sc[ls]);
                    }
                }
                __syncthreads();      // the X buffer becomes Z again
            }
        }
    }
}

}  // namespace r16
}  // namespace ems
