// scatter_sorted.cuh — a4 energy scatter as sort-by-cell + segmented reduce (EMS_FLAG_SORTED_SCATTER).
//
// BASELINE.json's north_star names this as the deterministic scatter; the engine's default
// deterministic mode is 64-bit fixed-point reductions instead (common.cuh), which is faster
// (DESIGN.md K4).  This file is the named design, built so that the two can be measured side by side:
//   1. every point gets a 64-bit key = its destination cell ((ch F + col) R + row; dropped points get
//      the sentinel `cells`, which sorts behind every real cell);
//   2. a STABLE partition moves the kept points to the front (in order; their count stays on the device),
//      then a STABLE least-significant-digit radix sort (8 bits per pass, only the bits `cells` needs)
//      orders the kept (key, energy) pairs by cell — points of one cell stay in ascending point order;
//   3. the first point of each run adds the run's energies left to right in fp32 and adds the sum to
//      the cell (no atomics: a cell has one run per chunk, chunks follow each other on the stream).
// The result is defined by the point order alone: bit-exact across runs, without the fixed-point
// quantisation or saturation.  Integer / ordering work throughout; HBM-bound.
#pragma once
#include "common.cuh"

namespace ems {
namespace sorted {

constexpr int kThreads = 256;               // 8 warps
constexpr int kWarps = kThreads / 32;
constexpr int kItems = 8;                   // elements per thread per sub-tile
constexpr int kTile = kThreads * kItems;    // 2,048 elements per sub-tile
constexpr int kRadix = 256;

// Step 1: destination cell of every point of [i0, i0 + n) (same rule as scatter_points_kernel).
__global__ void __launch_bounds__(256)
keys_kernel(const float* __restrict__ dt_cols, const float* __restrict__ dk_bins,
            const float* __restrict__ energy, long long i0, long long n,
            unsigned long long* __restrict__ keys, float* __restrict__ vals,
            long long F, int B, int rows, unsigned long long cells,
            int warp_mode, float warp_a, float warp_c, float inv_half) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
         j += (long long)gridDim.x * blockDim.x) {
        const long long i = i0 + j;
        const float e = __ldg(energy + i);
        unsigned long long key = cells;
        if (e > 0.f) {
            const long long row_id = i / B;                 // ch*F + f
            const int k = (int)(i - row_id * B);
            const long long ch = row_id / F;
            const long long f = row_id - ch * F;
            const long long col = f + (long long)rintf(__ldg(dt_cols + i));
            const float dk = __ldg(dk_bins + i);
            const int row = out_row(warp_mode, warp_a, warp_c, inv_half, k, dk, (float)k + dk);
            if (col >= 0 && col < F && row >= 0 && row < rows)
                key = (unsigned long long)((ch * F + col) * rows + row);
        }
        keys[j] = key;
        vals[j] = e;
    }
}

// Elements of a pass and the span each block walks.  The first pass of a chunk (the partition pass:
// digit = 1 for the sentinel keys of dropped points, 0 otherwise, so that the kept points come first, in
// order) works on all n points of the chunk; the radix passes after it only on the kept ones, whose
// number the partition pass left on the device (its total of digit 0) — no host round trip.
struct PassGeom { long long n, span; };
__device__ __forceinline__ PassGeom pass_geom(long long n_host, const unsigned* n_dev, int blocks) {
    PassGeom g;
    g.n = n_dev ? (long long)*n_dev : n_host;
    g.span = (g.n + blocks - 1) / blocks;
    g.span = (g.span + kTile - 1) / kTile * kTile;
    return g;
}
// part = the sentinel (partition pass) or 0 (radix pass on bits [shift, shift + 8))
__device__ __forceinline__ unsigned digit_of(unsigned long long key, int shift, unsigned long long part) {
    return part ? (key >= part ? 1u : 0u) : ((unsigned)(key >> shift) & 255u);
}

// Radix pass, part 1: digit counts of each block's span; counts[d * gridDim.x + block].
__global__ void __launch_bounds__(kThreads)
histogram_kernel(const unsigned long long* __restrict__ keys, long long n_host, const unsigned* __restrict__ n_dev,
                 int shift, unsigned long long part, unsigned* __restrict__ counts) {
    __shared__ unsigned hist[kRadix];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const PassGeom pg = pass_geom(n_host, n_dev, gridDim.x);
    const long long b0 = (long long)blockIdx.x * pg.span, b1 = min(pg.n, b0 + pg.span);
    if (part) {     // two digits only: count in registers, one atomic per warp
        unsigned c1 = 0, c = 0;
        for (long long i = b0 + threadIdx.x; i < b1; i += kThreads) { c1 += keys[i] >= part; ++c; }
        c1 = __reduce_add_sync(0xffffffffu, c1);
        c = __reduce_add_sync(0xffffffffu, c);
        if ((threadIdx.x & 31) == 0) { atomicAdd(&hist[1], c1); atomicAdd(&hist[0], c - c1); }
    } else
    for (long long i = b0 + threadIdx.x; i < b1; i += kThreads)
        atomicAdd(&hist[(unsigned)(keys[i] >> shift) & 255u], 1u);
    __syncthreads();
    counts[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = hist[threadIdx.x];
}

// Radix pass, part 2: one block per digit turns its row of the digit-major count table into exclusive
// offsets within the digit and writes the digit's total; the scatter kernel adds the totals of the
// smaller digits itself (256 values).
__global__ void __launch_bounds__(256)
scan_kernel(unsigned* __restrict__ counts, unsigned* __restrict__ totals, int G, unsigned* __restrict__ keep) {
    __shared__ unsigned part[256];
    unsigned* row = counts + (size_t)blockIdx.x * G;
    unsigned carry = 0;
    for (int i0 = 0; i0 < G; i0 += 256) {
        const int i = i0 + threadIdx.x;
        const unsigned c = i < G ? row[i] : 0u;
        part[threadIdx.x] = c;
        __syncthreads();
        for (int d = 1; d < 256; d <<= 1) {
            const unsigned v = threadIdx.x >= (unsigned)d ? part[threadIdx.x - d] : 0u;
            __syncthreads();
            part[threadIdx.x] += v;
            __syncthreads();
        }
        if (i < G) row[i] = carry + part[threadIdx.x] - c;
        carry += part[255];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        totals[blockIdx.x] = carry;
        if (keep && blockIdx.x == 0) *keep = carry;      // partition pass: digit 0 = the kept points of the chunk
    }
}

// Radix pass, part 3: stable scatter.  A block walks its span in sub-tiles; inside a sub-tile the
// order is (warp, item, lane), the same order the element indices have, so equal digits keep their
// relative order: rank inside the warp from __match_any_sync, running per-warp digit counters in shared
// memory, then an exclusive prefix over the warps on top of the block's running digit bases.
__global__ void __launch_bounds__(kThreads)
scatter_kernel(const unsigned long long* __restrict__ keys_in, const float* __restrict__ vals_in,
               unsigned long long* __restrict__ keys_out, float* __restrict__ vals_out,
               long long n_host, const unsigned* __restrict__ n_dev, int shift, unsigned long long part,
               const unsigned* __restrict__ offsets, const unsigned* __restrict__ totals) {
    __shared__ unsigned wcount[kWarps][kRadix];
    __shared__ unsigned base[kRadix];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {   // exclusive prefix of the digit totals (256 values) + this block's offset within its digit
        const unsigned t = totals[threadIdx.x];
        base[threadIdx.x] = t;
        __syncthreads();
        for (int d = 1; d < kRadix; d <<= 1) {
            const unsigned v = threadIdx.x >= (unsigned)d ? base[threadIdx.x - d] : 0u;
            __syncthreads();
            base[threadIdx.x] += v;
            __syncthreads();
        }
        base[threadIdx.x] += offsets[(size_t)threadIdx.x * gridDim.x + blockIdx.x] - t;
    }
    const PassGeom pg = pass_geom(n_host, n_dev, gridDim.x);
    const long long b0 = (long long)blockIdx.x * pg.span, b1 = min(pg.n, b0 + pg.span);
    for (long long t0 = b0; t0 < b1; t0 += kTile) {
#pragma unroll
        for (int w = 0; w < kWarps; ++w) wcount[w][threadIdx.x] = 0;
        __syncthreads();
        unsigned long long key[kItems];
        float val[kItems];
        unsigned rank[kItems];
        // all loads of the sub-tile are in flight before the first rank is computed
#pragma unroll
        for (int j = 0; j < kItems; ++j) {
            const long long i = t0 + (long long)warp * (32 * kItems) + j * 32 + lane;
            if (i < b1) { key[j] = keys_in[i]; val[j] = vals_in[i]; }
        }
#pragma unroll
        for (int j = 0; j < kItems; ++j) {
            const long long i = t0 + (long long)warp * (32 * kItems) + j * 32 + lane;
            const bool on = i < b1;
            const unsigned active = __ballot_sync(0xffffffffu, on);
            if (on) {
                const unsigned d = digit_of(key[j], shift, part);
                const unsigned peers = __match_any_sync(active, d);
                const int leader = __ffs(peers) - 1;
                unsigned old = 0;
                if (lane == leader) { old = wcount[warp][d]; wcount[warp][d] = old + __popc(peers); }
                old = __shfl_sync(peers, old, leader);
                rank[j] = old + __popc(peers & ((1u << lane) - 1u));
            }
            __syncwarp();
        }
        __syncthreads();
        {   // digit d = threadIdx.x: per-warp counts -> destinations, block base moves on
            unsigned run = base[threadIdx.x];
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const unsigned c = wcount[w][threadIdx.x];
                wcount[w][threadIdx.x] = run;
                run += c;
            }
            base[threadIdx.x] = run;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kItems; ++j) {
            const long long i = t0 + (long long)warp * (32 * kItems) + j * 32 + lane;
            if (i < b1) {
                const unsigned d = digit_of(key[j], shift, part);
                if (part && d) continue;                       // partition pass: dropped points are counted, not moved
                const unsigned dst = wcount[warp][d] + rank[j];
                keys_out[dst] = key[j];
                vals_out[dst] = val[j];
            }
        }
        __syncthreads();
    }
}

// Step 3: segmented reduce.  The first element of a run of equal keys adds the run left to right
// (ascending point order) and adds the sum to the fp32 accumulator cell; one run per cell per chunk.
__global__ void __launch_bounds__(256)
reduce_kernel(const unsigned long long* __restrict__ keys, const float* __restrict__ vals,
              const unsigned* __restrict__ n_dev, unsigned long long cells, float* __restrict__ acc,
              unsigned char* __restrict__ flags, long long F, int rows) {
    const long long n = (long long)*n_dev;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        if (key >= cells) continue;
        if (i > 0 && keys[i - 1] == key) continue;
        float s = vals[i];
        for (long long j = i + 1; j < n && keys[j] == key; ++j) s += vals[j];
        acc[key] += s;
        const long long row_id = (long long)(key / (unsigned long long)rows);    // ch*F + col
        const int row = (int)(key - (unsigned long long)row_id * rows);
        const long long ch = row_id / F;
        flags[flag_index((int)ch, F, rows, row_id - ch * F, row)] = 1;
    }
}

}  // namespace sorted
}  // namespace ems
