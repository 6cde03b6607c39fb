// common.cuh — shared types of the engine (host + device).
// Stand-in engine for the "reassignment method" of /root/reference/README.md:3,11;
// conventions are SURVEY.md §7 / oracle/reassign_oracle.py header.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ems {

// Fixed-point scale of the deterministic accumulator: energy * 2^44 as u64.
// Integer adds are associative, so the grid is bit-exact in any deposit order.
// Range per cell 2^20 (a full-scale sine is 1.0), resolution 5.7e-14 (-132 dB).
constexpr float  kFixScale    = 17592186044416.0f;        // 2^44
constexpr double kFixScaleInv = 1.0 / 17592186044416.0;
// A single point contributes at most 2^12 (36 dB over a full-scale sine): 2^8 such points still
// fit a cell, so PCM far outside [-1, 1] saturates instead of wrapping the 64-bit sum.
constexpr float  kFixMaxEnergy = 4096.0f;
__device__ __forceinline__ unsigned long long fix_energy(float e) {
    return __float2ull_rn(fminf(e, kFixMaxEnergy) * kFixScale);
}

// Deposits are reductions into global memory: spelled as red.global so that pointers that
// reach the call as generic addresses (kernel-argument structs passed through a non-inlined
// function) do not compile to a returning ATOM with a shared-memory fallback loop.
__device__ __forceinline__ void red_add_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void flag_set(unsigned char* p) {
    asm volatile("st.global.u8 [%0], %1;" ::"l"(p), "r"(1) : "memory");
}

enum DepositMode : int {
    kStorePoints = 0,   // write (dt_cols, dk_bins, energy) triples
    kDepositU64  = 1,   // red.global.add.u64 into the fixed-point accumulator
    kDepositF32  = 2    // red.global.add.f32 fast mode
};

// EXPERIMENT, compiled in only with -DEMS_FUSED_POST=1 (measured 5 x slower than the two-kernel path:
// a worker that runs the shaping code next to workers running the FFT thrashes the instruction cache;
// profiles/r02_fused_post_*.txt).
// In-kernel post-pass ("fused" mode of the deposit kernels, stft_r16.cuh): the accumulator is one ring
// of `vring` columns shared by all channels (virtual column = ch * F + col), small enough to stay in
// L2; a CTA that completes the last tile a column block was waiting for shapes that block straight
// from L2 into the colour-index image and clears it.  Nothing the size of the stream exists besides
// the outputs.
struct FusedPost {
    int*           ready;     // [n_tiles] finished tiles among the 2 nb + 1 a block waits for (the finisher resets it)
    int*           done;      // [n_tiles] = epoch once the block has been shaped and cleared
    uint8_t*       index;     // [channels][F][rows] or null
    float*         grid;      // [channels][F][rows] or null
    const float*   weight;    // [rows]
    float          db_floor, inv_range, gate_db;
    int            epoch;     // this launch's stamp for done[]
    int            nb;        // tiles on each side whose deposits can reach a block: ceil(R / tile_T)
    int            vring;     // columns of the ring (a power of two); 0: fused mode off
    int            NB;        // flag_blocks(rows)
    int            debug;     // EMS_FUSED_DEBUG (timing experiments only): 1 = skip the shaping, 2 = skip the ring wait
};

// Arguments of the fused frame-gather + 3-window STFT + reassignment kernels.
struct StftArgs {
    const float*  pcm;        // planar [channels][S]
    long long     S;          // samples per channel
    long long     F;          // frames per channel
    long long     f_begin;    // frame range of this launch (per channel)
    long long     f_end;
    int           channels;
    int           hop;
    const float*  thw;        // [N] th'[n] = (n - N/2) (2/N) (0.5 - 0.5 cos(2 pi n/N))
    const float2* tw;         // [N] W_N^j = (cos, -sin)(2 pi j / N)
    float*        dt_cols;    // [channels][F][B] or null
    float*        dk_bins;
    float*        energy;
    void*         acc;        // [channels][F][B] u64 or f32 accumulator, or null
    unsigned char* flags;     // [channels][NB][F] "this 64-bin block of this column holds energy", or null
    float         gate_lin;   // drop points with energy <= gate
    float         inv_hop;
    int           mode;       // DepositMode
    int           reassign;   // 0: plain spectrogram columns
    int           rows;       // output rows per column R (B, or display_rows)
    int           warp_mode;  // 0: row = k + rint(dk); 1: linear axis resample; 2: log1p warp
    float         warp_a;     // a = 10^(2 freq_scale) - 1
    float         warp_c;     // (R-1)/log1p(a)  (mode 2)  or  (R-1)  (mode 1)
    float         inv_half;   // 2 / n_fft
    long long     samp_off;   // frame f starts at sample f*hop + samp_off of its channel
    int           ring;       // > 0: accumulator and flags are rings of `ring` columns per channel (a power of two;
                              //      streaming, and the O(chunk) host path); 0: linear, F columns per channel
    int           stream_M;   // streaming: pushes per ring lap = ceil(n_fft / hop)
    const long long* sstate;  // streaming: device counter of completed pushes (frame range decoded on device)
    FusedPost     fp;         // in-kernel post-pass (fp.vring > 0)
};

// Streaming launches sit in a CUDA graph, so the frame they analyse is derived on the device
// from the push counter: push i (0-based) completes frame f = i + 1 - M, whose first sample
// sits at ((i mod M) + 1) * hop of the doubled sample ring.  Returns false while the ring fills.
__device__ __forceinline__ bool stream_decode(StftArgs& a) {
    if (!a.sstate) return true;
    const long long i = a.sstate[0];
    const long long f = i + 1 - a.stream_M;
    if (f < 0) return false;
    a.f_begin = f;
    a.f_end = f + 1;
    a.samp_off = ((i % a.stream_M) + 1) * a.hop - f * a.hop;
    return true;
}

constexpr int kFlagShift = 6;                              // one dirty flag per 64 bins of a column
__host__ __device__ constexpr int flag_blocks(int B) { return (B + 63) >> kFlagShift; }
// Flags are column-contiguous per bin block ([channels][NB][F]) so the post-pass reads the
// flags of 16 consecutive columns from one cache line.
// `ncols` columns per channel (F, or the ring size), `slot` = the column's place among them.
__device__ __forceinline__ long long flag_index(int ch, long long ncols, int rows, long long slot, int row) {
    return ((long long)ch * flag_blocks(rows) + (row >> kFlagShift)) * ncols + slot;
}

// Accumulator cell of (channel, column, row): linear [channels][F][R] or a ring of 2^n columns.
__device__ __forceinline__ long long acc_cell(const StftArgs& a, int ch, long long col, int row) {
    return a.ring ? ((long long)ch * a.ring + (col & (a.ring - 1))) * a.rows + row
                  : ((long long)ch * a.F + col) * a.rows + row;
}
__device__ __forceinline__ long long acc_flag(const StftArgs& a, int ch, long long col, int row) {
    return a.ring ? flag_index(ch, a.ring, a.rows, col & (a.ring - 1), row) : flag_index(ch, a.F, a.rows, col, row);
}

// Output row of a point at reassigned frequency wh = k + dk [bins]
// ("Frequency Scale", /root/reference/README.md:48; oracle/reassign_oracle.py::output_row).
__device__ __forceinline__ int out_row(int warp_mode, float warp_a, float warp_c, float inv_half,
                                       int k, float dk, float wh) {
    if (warp_mode == 0) return k + (int)rintf(dk);
    const float x = fminf(fmaxf(wh * inv_half, 0.f), 1.f);
    return (int)rintf((warp_mode == 2 ? log1pf(warp_a * x) : x) * warp_c);
}

struct PostArgs {
    void*         acc;        // u64 or f32 [channels][F][B]; cells read by the emit pass are cleared
    unsigned char* flags;     // [channels][NB][F] dirty flags written by the deposits
    int           NB;
    int           acc_is_u64;
    float*        grid;       // fp32 [channels][F][B] or null
    uint8_t*      index;      // u8   [channels][F][B] or null
    const float*  weight;     // [B] gain^2 * w_low(k)
    float*        carry;      // [channels][B] EMA state entering col_begin (updated)
    long long     F;          // columns per channel of the stream (colscale is [channels][F])
    long long     col_begin;
    long long     col_end;
    long long     acc_cols;   // columns per channel of acc / flags: F, or the ring size
    long long     acc_mask;   // slot of column c = c & acc_mask (-1: linear)
    long long     out_cols;   // columns per channel of grid / index: F, or the staging buffer's
    long long     out_col0;   // stream column held by output column 0
    int           B;          // output rows per column R
    int           channels;
    float         smoothing;
    float         db_floor;   // TOP_DB - range
    float         inv_range;  // 255 / range
    float         gate_db;
    float*        colscale;   // AGC: [channels][F] level^-strength per column (peaks while measuring), or null
};

constexpr float kAgcReleaseSeconds = 1.0f;

}  // namespace ems
