// scatter_post.cuh — a4 energy scatter from stored points, and the a5 display post-pass
// (gain, low-end boost, smoothing, dB, noise gate, colour index;
// /root/reference/README.md:46-51, oracle/reassign_oracle.py::scatter_grid / postpass).
#pragma once
#include "common.cuh"

namespace ems {

// a4 from stored points: G[f + rint(dt), k + rint(dk)] += e for e > 0.
// One thread per point, grid-stride; reads are coalesced, deposits are reds at L2.
__global__ void __launch_bounds__(256)
scatter_points_kernel(const float* __restrict__ dt_cols, const float* __restrict__ dk_bins,
                      const float* __restrict__ energy, void* __restrict__ acc, int acc_is_u64,
                      unsigned char* __restrict__ flags, long long F, int B, int channels,
                      int rows, int warp_mode, float warp_a, float warp_c, float inv_half) {
    const long long total = (long long)channels * F * B;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const float e = __ldg(energy + i);
        if (!(e > 0.f)) continue;
        const long long row_id = i / B;                 // ch*F + f
        const int k = (int)(i - row_id * B);
        const long long ch = row_id / F;
        const long long f = row_id - ch * F;
        const long long col = f + (long long)rintf(__ldg(dt_cols + i));
        const float dk = __ldg(dk_bins + i);
        const int row = out_row(warp_mode, warp_a, warp_c, inv_half, k, dk, (float)k + dk);
        if (col < 0 || col >= F || row < 0 || row >= rows) continue;   // caller-made points
        const long long o = (ch * F + col) * rows + row;
        if (acc_is_u64) red_add_u64(reinterpret_cast<unsigned long long*>(acc) + o, fix_energy(e));
        else red_add_f32(reinterpret_cast<float*>(acc) + o, e);
        flags[flag_index((int)ch, F, rows, col, row)] = 1;
    }
}

__device__ __forceinline__ float acc_load(const void* acc, int is_u64, long long o) {
    if (is_u64) {
        const unsigned long long v = reinterpret_cast<const unsigned long long*>(acc)[o];
        return __double2float_rn(__ull2double_rn(v) * kFixScaleInv);
    }
    return reinterpret_cast<const float*>(acc)[o];
}

// E: shaped energy (the noise gate sees it); scale: AGC factor level^-strength of the column
__device__ __forceinline__ uint8_t colour_index(float E, float scale, const PostArgs& a) {
    if (!(E > 0.f)) return 0;
    if (10.0f * log10f(E) < a.gate_db) return 0;
    const float db = 10.0f * log10f(E * scale);
    const float v = rintf((db - a.db_floor) * a.inv_range);
    return (uint8_t)fminf(fmaxf(v, 0.f), 255.f);
}
__device__ __forceinline__ uint8_t colour_index(float E, const PostArgs& a) {
    if (!(E > 0.f)) return 0;
    const float db = 10.0f * log10f(E);
    if (db < a.gate_db) return 0;
    const float v = rintf((db - a.db_floor) * a.inv_range);
    return (uint8_t)fminf(fmaxf(v, 0.f), 255.f);
}

// Column peak for the AGC: E >= 0, so float order = integer order of the bit patterns.
__device__ __forceinline__ void peak_max(float* colpeak, long long i, float E) {
    atomicMax(reinterpret_cast<int*>(colpeak) + i, __float_as_int(E));
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}

// Load a cell for the emit pass and leave it zero: the accumulator is clean again when the
// post-pass has covered every column, so no call ever memsets it (22 GB at configs[2]).
__device__ __forceinline__ float acc_take(void* acc, int is_u64, long long o) {
    if (is_u64) {
        unsigned long long* p = reinterpret_cast<unsigned long long*>(acc) + o;
        const unsigned long long v = *p;
        *p = 0ull;
        return __double2float_rn(__ull2double_rn(v) * kFixScaleInv);
    }
    float* p = reinterpret_cast<float*>(acc) + o;
    const float v = *p;
    *p = 0.f;
    return v;
}

__device__ __forceinline__ void acc_zero(void* acc, int is_u64, long long o) {
    if (is_u64) reinterpret_cast<unsigned long long*>(acc)[o] = 0ull;
    else reinterpret_cast<float*>(acc)[o] = 0.f;
}

// Stream column c of channel ch: its accumulator row, its dirty flag (of bin block blk), its output row.
__device__ __forceinline__ long long acc_row(const PostArgs& a, int ch, long long c) {
    return ((long long)ch * a.acc_cols + (c & a.acc_mask)) * a.B;
}
__device__ __forceinline__ long long flag_at(const PostArgs& a, int ch, int blk, long long c) {
    return ((long long)ch * a.NB + blk) * a.acc_cols + (c & a.acc_mask);
}
__device__ __forceinline__ long long out_row_of(const PostArgs& a, int ch, long long c) {
    return ((long long)ch * a.out_cols + (c - a.out_col0)) * a.B;
}

constexpr int kPostChunk = 256;   // columns per thread in the EMA scan
constexpr int kPostTile = 8;      // columns per thread when there is no recurrence along time

// Smoothing pass A: per (channel, chunk, bin) the EMA of the chunk from a zero carry;
// only the chunk's last value is kept.  local_end: [channels][n_chunks][B].
__global__ void __launch_bounds__(128)
post_ema_local_kernel(const PostArgs a, float* __restrict__ local_end, int n_chunks) {
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    const int chunk = blockIdx.x, ch = blockIdx.z;
    if (k >= a.B) return;
    const long long c0 = a.col_begin + (long long)chunk * kPostChunk;
    const long long c1 = min(c0 + (long long)kPostChunk, a.col_end);
    const float w = a.weight[k], s = a.smoothing, oms = 1.0f - a.smoothing;
    float y = 0.f;
    const int blk = k >> kFlagShift;
    for (long long c = c0; c < c1; ++c) {
        const float E = a.flags[flag_at(a, ch, blk, c)] ? acc_load(a.acc, a.acc_is_u64, acc_row(a, ch, c) + k) * w : 0.f;
        y = s * y + oms * E;
    }
    local_end[((long long)ch * n_chunks + chunk) * a.B + k] = y;
}

// Smoothing pass B: carries entering each chunk, sequential over chunks (short loop).
// carry_in: [channels][n_chunks][B]; a.carry enters chunk 0 and is advanced to col_end.
__global__ void __launch_bounds__(128)
post_ema_carry_kernel(const PostArgs a, const float* __restrict__ local_end,
                      float* __restrict__ carry_in, int n_chunks, float s_pow_chunk,
                      float s_pow_last) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int ch = blockIdx.y;
    if (k >= a.B) return;
    float c = a.carry[(long long)ch * a.B + k];
    for (int j = 0; j < n_chunks; ++j) {
        const long long o = ((long long)ch * n_chunks + j) * a.B + k;
        carry_in[o] = c;
        c = local_end[o] + (j == n_chunks - 1 ? s_pow_last : s_pow_chunk) * c;
    }
    a.carry[(long long)ch * a.B + k] = c;
}

// Pass C (the only pass when smoothing == 0): grid fp32 and colour index.
// measure = 1 (AGC first pass): nothing is written or cleared, the per-column maximum of the
// shaped, smoothed energy goes to a.colscale.
__global__ void __launch_bounds__(128, 8)
post_emit_kernel(const PostArgs a, const float* __restrict__ carry_in, int n_chunks, int chunk_cols,
                 int measure) {
    const int k = min(blockIdx.y * blockDim.x + threadIdx.x, a.B - 1);   // tail lanes shadow the last row (warp_max)
    const bool tail = blockIdx.y * blockDim.x + threadIdx.x >= a.B;
    const int chunk = blockIdx.x, ch = blockIdx.z;
    if (tail && !measure) return;
    const long long c0 = a.col_begin + (long long)chunk * chunk_cols;
    const long long c1 = min(c0 + (long long)chunk_cols, a.col_end);
    const float w = a.weight[k], s = a.smoothing, oms = 1.0f - a.smoothing;
    float y = carry_in ? carry_in[((long long)ch * n_chunks + chunk) * a.B + k] : 0.f;
    const int blk = k >> kFlagShift;
    constexpr int kTile = kPostTile;   // columns whose flags and cells are fetched together
    for (long long cb = c0; cb < c1; cb += kTile) {
        const int nt = (int)min((long long)kTile, c1 - cb);
        unsigned char f[kTile];
#pragma unroll
        for (int i = 0; i < kTile; ++i) f[i] = (i < nt) ? a.flags[flag_at(a, ch, blk, cb + i)] : 0;
        float G[kTile];
        // clean 64-bin blocks (no deposit since the last post-pass) are neither read nor cleared
#pragma unroll
        for (int i = 0; i < kTile; ++i) {
            const long long o = acc_row(a, ch, cb + i) + k;
            G[i] = !f[i] ? 0.f : measure ? acc_load(a.acc, a.acc_is_u64, o) : acc_take(a.acc, a.acc_is_u64, o);
        }
#pragma unroll
        for (int i = 0; i < kTile; ++i) {
            if (i < nt) {
                const long long o = out_row_of(a, ch, cb + i) + k;
                float E = G[i] * w;
                if (s > 0.f) { y = s * y + oms * E; E = y; }
                if (measure) {
                    const float m = warp_max(E);           // 32 rows of this column
                    if ((threadIdx.x & 31) == 0 && m > 0.f) peak_max(a.colscale, (long long)ch * a.F + cb + i, m);
                    continue;
                }
                if (a.grid) a.grid[o] = G[i];
                if (a.index)
                    a.index[o] = a.colscale ? colour_index(E, a.colscale[(long long)ch * a.F + cb + i], a)
                                            : colour_index(E, a);
            }
        }
    }
}

// Post-pass without smoothing: every cell is independent and the image is mostly empty, so
// the outputs are zero-filled by memset and this kernel visits only the dirty 64-bin blocks.
// One warp reads 32 consecutive column flags of a (channel, bin block) row; each flagged
// block is then shaped by the whole warp (2 bins per lane, coalesced), its accumulator cells
// and its flag are cleared.  grid: (channels*NB, ceil(ncols/256)), 256 threads.
// measure = 1 (AGC first pass): only the per-column peak of the shaped energy is produced.
__global__ void __launch_bounds__(256)
post_sparse_kernel(const PostArgs a, int measure, const int* __restrict__ mode) {
    if (mode && mode[0] == 1) return;                       // the dense kernel takes this range
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x;                               // ch * NB + blk (x: up to 65535 channels x NB blocks)
    const int ch = r / a.NB, blk = r - ch * a.NB;
    const long long cw = a.col_begin + ((long long)blockIdx.y * 8 + (threadIdx.x >> 5)) * 32;
    if (cw >= a.col_end) return;
    const long long c = cw + lane;
    unsigned mask = __ballot_sync(0xffffffffu, c < a.col_end && a.flags[flag_at(a, ch, blk, c)] != 0);
    const int k0 = (blk << kFlagShift) + lane, k1 = k0 + 32;
    const float w0 = k0 < a.B ? a.weight[k0] : 0.f, w1 = k1 < a.B ? a.weight[k1] : 0.f;
    constexpr int kBatch = 4;     // flagged blocks in flight per warp: 8 independent loads per lane
    while (mask) {
        int js[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            js[b] = mask ? __ffs(mask) - 1 : -1;
            mask &= mask - 1;
        }
        float G0[kBatch], G1[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            const long long row = acc_row(a, ch, cw + max(js[b], 0));
            const bool on = js[b] >= 0;
            G0[b] = (on && k0 < a.B) ? acc_load(a.acc, a.acc_is_u64, row + k0) : 0.f;
            G1[b] = (on && k1 < a.B) ? acc_load(a.acc, a.acc_is_u64, row + k1) : 0.f;
        }
        if (measure) {
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                if (js[b] < 0) continue;
                const float m = warp_max(fmaxf(G0[b] * w0, G1[b] * w1));
                if (lane == 0 && m > 0.f) peak_max(a.colscale, (long long)ch * a.F + cw + js[b], m);
            }
            continue;
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            if (js[b] < 0) continue;
            const long long row = acc_row(a, ch, cw + js[b]), orow = out_row_of(a, ch, cw + js[b]);
            const float sc = a.colscale ? a.colscale[(long long)ch * a.F + cw + js[b]] : 1.0f;
            if (k0 < a.B) {
                acc_zero(a.acc, a.acc_is_u64, row + k0);
                if (a.grid) a.grid[orow + k0] = G0[b];
                if (a.index) a.index[orow + k0] = a.colscale ? colour_index(G0[b] * w0, sc, a) : colour_index(G0[b] * w0, a);
            }
            if (k1 < a.B) {
                acc_zero(a.acc, a.acc_is_u64, row + k1);
                if (a.grid) a.grid[orow + k1] = G1[b];
                if (a.index) a.index[orow + k1] = a.colscale ? colour_index(G1[b] * w1, sc, a) : colour_index(G1[b] * w1, a);
            }
            if (lane == 0) a.flags[flag_at(a, ch, blk, cw + js[b])] = 0;
        }
    }
}

// How dirty is the image?  Counts the set flags of columns [col_begin, col_end) and leaves
// mode[0] = 1 when more than half of the 64-row blocks hold energy: the post-pass then streams whole
// columns (post_dense_kernel) instead of visiting flagged blocks one by one (post_sparse_kernel).
// Both kernels are launched; the one that is not selected exits at once.
__global__ void __launch_bounds__(256)
post_density_kernel(const PostArgs a, unsigned long long* __restrict__ count, int* __restrict__ mode, int pass) {
    const long long n = a.col_end - a.col_begin, rows = (long long)a.channels * a.NB;
    if (pass == 1) {                 // single thread: decide, reset the counter for the next call
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            // (every fourth flag was counted; the dense kernel stages a column of at most 4096 rows)
            mode[0] = (a.B <= 4096 && 8 * count[0] > (unsigned long long)(n * rows)) ? 1 : 0;
            count[0] = 0;
        }
        return;
    }
    unsigned c = 0;
    for (long long i = 4 * ((long long)blockIdx.x * blockDim.x + threadIdx.x); i < n * rows; i += 4 * (long long)gridDim.x * blockDim.x) {
        const long long r = i / n;
        c += a.flags[r * a.acc_cols + ((a.col_begin + (i - r * n)) & a.acc_mask)] != 0;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, (unsigned long long)c);
}

// Dense post-pass (no smoothing, no AGC): one block per column, every cell of the column is read,
// shaped and cleared, clean or not — 16 KB of contiguous accumulator per column, so the whole column
// is in flight at once.  A thread owns four consecutive rows, shifted so that its four index bytes are
// one aligned 32-bit store.  The caller clears the flags of the range afterwards.
template <int kIter>                                      // kIter x 128 cells in flight per block: lanes read consecutive cells
__global__ void __launch_bounds__(128)
post_dense_kernel(const PostArgs a, const int* __restrict__ mode) {
    if (mode[0] != 1) return;
    __shared__ __align__(16) uint8_t col[4352];              // the column's index bytes (rows <= 4096 in this mode), 4 bytes of slack each end
    const long long ncols = a.col_end - a.col_begin;
    const int t = threadIdx.x;
    for (long long job = blockIdx.x; job < ncols * a.channels; job += gridDim.x) {
    const int ch = (int)(job / ncols);
    const long long c = a.col_begin + (job - (long long)ch * ncols);
    __syncthreads();                                          // the previous column's bytes have been written out
    const long long arow = acc_row(a, ch, c), orow = out_row_of(a, ch, c);
    uint8_t* ix = a.index ? a.index + orow : nullptr;
    float* gr = a.grid ? a.grid + orow : nullptr;
    for (int base = 0; base < a.B; base += 128 * kIter) {
        float G[kIter];
#pragma unroll
        for (int i = 0; i < kIter; ++i) {
            const int r = base + 128 * i + t;
            G[i] = r < a.B ? acc_load(a.acc, a.acc_is_u64, arow + r) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < kIter; ++i) {
            const int r = base + 128 * i + t;
            if (r < a.B) {
                acc_zero(a.acc, a.acc_is_u64, arow + r);
                if (gr) gr[r] = G[i];
                if (ix) col[4 + r] = G[i] > 0.f ? colour_index(G[i] * a.weight[r], a) : (uint8_t)0;
            }
        }
    }
    if (!ix) continue;
    __syncthreads();
    // aligned 32-bit words of the output row from the staged bytes; the ragged ends byte by byte
    const int a0 = (int)((4 - ((unsigned long long)ix & 3ull)) & 3ull);
    const int nw = (a.B - a0) >> 2;
    for (int w = t; w < nw; w += 128) {
        const int r = a0 + 4 * w;
        const unsigned v = (unsigned)col[4 + r] | ((unsigned)col[5 + r] << 8) | ((unsigned)col[6 + r] << 16) | ((unsigned)col[7 + r] << 24);
        *reinterpret_cast<unsigned*>(ix + r) = v;
    }
    if (t < a0) ix[t] = col[4 + t];
    const int tail0 = a0 + 4 * nw;
    if (tail0 + t < a.B && t < 4) ix[tail0 + t] = col[4 + tail0 + t];
    }
}

// Capture-side formats: integer PCM, interleaved [n][channels] (one chunk of the
// stream, starting at the buffer's first byte) -> fp32 planar [channels][S] samples [0, n).
// int16: full scale 32768.
__global__ void pcm_i16_to_planar_kernel(const int16_t* __restrict__ in, float* __restrict__ out,
                                         long long S, int channels, long long n) {
    const long long total = n * channels;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long smp = e / channels;
        const int ch = (int)(e - smp * channels);
        out[(long long)ch * S + smp] = (float)in[e] * (1.0f / 32768.0f);
    }
}
// int24: three bytes per sample, little-endian, packed; full scale 2^23 (exact in fp32).  A thread
// unpacks four samples from three aligned 32-bit words.
__global__ void pcm_i24_to_planar_kernel(const uint32_t* __restrict__ in, float* __restrict__ out,
                                         long long S, int channels, long long n) {
    const long long total = n * channels, quads = (total + 3) / 4;
    const unsigned char* in8 = reinterpret_cast<const unsigned char*>(in);
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads;
         q += (long long)gridDim.x * blockDim.x) {
        int v[4];
        if (4 * q + 4 <= total) {
            const uint32_t w0 = in[3 * q], w1 = in[3 * q + 1], w2 = in[3 * q + 2];
            v[0] = (int)(w0 << 8) >> 8;
            v[1] = (int)(((w0 >> 24) | (w1 << 8)) << 8) >> 8;
            v[2] = (int)(((w1 >> 16) | (w2 << 16)) << 8) >> 8;
            v[3] = (int)w2 >> 8;
        } else {       // the last, partial quad: byte loads
            for (int i = 0; i < 4; ++i) {
                const long long e = 4 * q + i;
                v[i] = e < total ? (int)(((uint32_t)in8[3 * e] | ((uint32_t)in8[3 * e + 1] << 8) | ((uint32_t)in8[3 * e + 2] << 16)) << 8) >> 8 : 0;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long e = 4 * q + i;
            if (e < total) {
                const long long smp = e / channels;
                const int ch = (int)(e - smp * channels);
                out[(long long)ch * S + smp] = (float)v[i] * (1.0f / 8388608.0f);
            }
        }
    }
}

// AGC level scan over columns [c0, c1) of one channel per block: colscale holds the column
// peaks on entry and target * level^-strength on exit (target: the "Brightness" control).  level[m] = max(peak[m], lambda * level[m-1]);
// `level_carry[ch]` enters c0 and leaves with the level after c1 - 1.  The operator is
// associative (max-plus with decay), so each thread scans a chunk and thread 0 chains them.
__global__ void __launch_bounds__(1024)
agc_scan_kernel(float* __restrict__ colscale, float* __restrict__ level_carry, long long F,
                long long c0, long long c1, float lambda, float strength, float target) {
    __shared__ float s_end[1024];
    __shared__ float s_in[1024];
    const int ch = blockIdx.x, t = threadIdx.x;
    float* cs = colscale + (long long)ch * F;
    const long long n = c1 - c0, per = (n + 1023) / 1024;
    const long long a0 = c0 + (long long)t * per, a1 = min(a0 + per, c1);
    float lv = 0.f;
    for (long long m = a0; m < a1; ++m) lv = fmaxf(cs[m], lambda * lv);
    s_end[t] = lv;
    __syncthreads();
    if (t == 0) {
        float c = level_carry[ch];
        for (int j = 0; j < 1024; ++j) {
            s_in[j] = c;
            const long long b0 = c0 + (long long)j * per, len = max(0LL, min(b0 + per, c1) - b0);
            c = fmaxf(s_end[j], c * powf(lambda, (float)len));
        }
        level_carry[ch] = c;
    }
    __syncthreads();
    lv = s_in[t];
    for (long long m = a0; m < a1; ++m) {
        lv = fmaxf(cs[m], lambda * lv);
        cs[m] = lv > 0.f ? target * powf(lv, -strength) : 1.0f;
    }
}

// Colour map: rgba[i] = lut[index[i]].  HBM-bound (1 B read, 4 B written per cell): the body
// moves 16 indices -> four 16-byte stores per thread; head/tail cells keep any alignment legal.
__global__ void __launch_bounds__(256)
colorize_kernel(const uint8_t* __restrict__ index, uint32_t* __restrict__ rgba, size_t n,
                const uint32_t* __restrict__ lut_dev) {
    __shared__ uint32_t lut[256];
    lut[threadIdx.x] = lut_dev[threadIdx.x];
    __syncthreads();
    const size_t head = min(n, (size_t)((16 - ((uintptr_t)index & 15)) & 15));   // to 16-byte alignment of index
    const size_t nvec = (n - head) / 16;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < head; i += nth) rgba[i] = lut[index[i]];
    const uint4* iv = reinterpret_cast<const uint4*>(index + head);
    for (size_t v = tid; v < nvec; v += nth) {
        const uint4 q = __ldg(iv + v);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint32_t* o = rgba + head + v * 16;          // 4-byte aligned, stores as 32-bit words
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            o[4 * j + 0] = lut[w[j] & 255];
            o[4 * j + 1] = lut[(w[j] >> 8) & 255];
            o[4 * j + 2] = lut[(w[j] >> 16) & 255];
            o[4 * j + 3] = lut[w[j] >> 24];
        }
    }
    for (size_t i = head + nvec * 16 + tid; i < n; i += nth) rgba[i] = lut[index[i]];
}

// Per-channel summary of a u8 image of `n` bytes per channel: out[ch] = (sum of bytes, sum of
// byte * (1 + pos mod 65521)), 64-bit wrapping.  grid (blocks, channels); 16-byte loads on the aligned
// body, byte loads at the ends; one 64-bit red pair per block.
__global__ void __launch_bounds__(256)
image_summary_kernel(const uint8_t* __restrict__ img, size_t n, unsigned long long* __restrict__ out) {
    const uint8_t* p = img + (size_t)blockIdx.y * n;
    const size_t head = min(n, (size_t)((16 - ((uintptr_t)p & 15)) & 15));
    const size_t nvec = (n - head) / 16;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    unsigned long long s0 = 0, s1 = 0;
    auto add = [&](size_t pos, unsigned b) { s0 += b; s1 += (unsigned long long)b * (1ull + pos % 65521ull); };
    for (size_t i = tid; i < head; i += nth) add(i, p[i]);
    const uint4* v = reinterpret_cast<const uint4*>(p + head);
    for (size_t j = tid; j < nvec; j += nth) {
        const uint4 q = __ldg(v + j);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        const size_t pos0 = head + 16 * j;
        unsigned m = (unsigned)(pos0 % 65521ull);            // position modulus, advanced incrementally
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const unsigned b = (w[e >> 2] >> (8 * (e & 3))) & 255u;
            s0 += b;
            s1 += (unsigned long long)(b * (m + 1u));
            m = m + 1u == 65521u ? 0u : m + 1u;
        }
    }
    for (size_t i = head + 16 * nvec + tid; i < n; i += nth) add(i, p[i]);
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, d);
        s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    }
    if ((threadIdx.x & 31) == 0) {
        red_add_u64(out + 2 * blockIdx.y, s0);
        red_add_u64(out + 2 * blockIdx.y + 1, s1);
    }
}

// Clears the dirty flags of columns [c0, c1) of every (channel, bin block) row.
// `ncols` columns per row (F, or the ring size), column c sits at slot c & mask.
__global__ void clear_flags_kernel(unsigned char* __restrict__ flags, long long ncols, long long mask,
                                   long long c0, long long c1, int rows) {
    const long long n = c1 - c0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * rows;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / n;
        flags[r * ncols + ((c0 + (i - r * n)) & mask)] = 0;
    }
}

// The same, only when the dense post-pass ran (it does not clear flags block by block).
__global__ void clear_flags_if_kernel(unsigned char* __restrict__ flags, long long ncols, long long mask,
                                      long long c0, long long c1, int rows, const int* __restrict__ mode) {
    if (mode[0] != 1) return;
    const long long n = c1 - c0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * rows;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / n;
        flags[r * ncols + ((c0 + (i - r * n)) & mask)] = 0;
    }
}

}  // namespace ems
