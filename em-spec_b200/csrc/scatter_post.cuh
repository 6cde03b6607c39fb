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
                      long long F, int B, int channels) {
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
        const int row = k + (int)rintf(__ldg(dk_bins + i));
        if (col < 0 || col >= F || row < 0 || row >= B) continue;   // caller-made points
        const long long o = (ch * F + col) * B + row;
        if (acc_is_u64)
            atomicAdd(reinterpret_cast<unsigned long long*>(acc) + o,
                      __float2ull_rn(e * kFixScale));
        else
            atomicAdd(reinterpret_cast<float*>(acc) + o, e);
    }
}

__device__ __forceinline__ float acc_load(const void* acc, int is_u64, long long o) {
    if (is_u64) {
        const unsigned long long v = reinterpret_cast<const unsigned long long*>(acc)[o];
        return __double2float_rn(__ull2double_rn(v) * kFixScaleInv);
    }
    return reinterpret_cast<const float*>(acc)[o];
}

__device__ __forceinline__ uint8_t colour_index(float E, const PostArgs& a) {
    if (!(E > 0.f)) return 0;
    const float db = 10.0f * log10f(E);
    if (db < a.gate_db) return 0;
    const float v = rintf((db - a.db_floor) * a.inv_range);
    return (uint8_t)fminf(fmaxf(v, 0.f), 255.f);
}

constexpr int kPostChunk = 256;   // columns per thread in the EMA scan

// Smoothing pass A: per (channel, chunk, bin) the EMA of the chunk from a zero carry;
// only the chunk's last value is kept.  local_end: [channels][n_chunks][B].
__global__ void __launch_bounds__(128)
post_ema_local_kernel(const PostArgs a, float* __restrict__ local_end, int n_chunks) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = blockIdx.y, ch = blockIdx.z;
    if (k >= a.B) return;
    const long long c0 = a.col_begin + (long long)chunk * kPostChunk;
    const long long c1 = min(c0 + (long long)kPostChunk, a.col_end);
    const float w = a.weight[k], s = a.smoothing, oms = 1.0f - a.smoothing;
    float y = 0.f;
    for (long long c = c0; c < c1; ++c) {
        const float E = acc_load(a.acc, a.acc_is_u64, ((long long)ch * a.F + c) * a.B + k) * w;
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
__global__ void __launch_bounds__(128)
post_emit_kernel(const PostArgs a, const float* __restrict__ carry_in, int n_chunks) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = blockIdx.y, ch = blockIdx.z;
    if (k >= a.B) return;
    const long long c0 = a.col_begin + (long long)chunk * kPostChunk;
    const long long c1 = min(c0 + (long long)kPostChunk, a.col_end);
    const float w = a.weight[k], s = a.smoothing, oms = 1.0f - a.smoothing;
    float y = carry_in ? carry_in[((long long)ch * n_chunks + chunk) * a.B + k] : 0.f;
    for (long long c = c0; c < c1; ++c) {
        const long long o = ((long long)ch * a.F + c) * a.B + k;
        const float G = acc_load(a.acc, a.acc_is_u64, o);
        if (a.grid) a.grid[o] = G;
        if (a.index) {
            float E = G * w;
            if (s > 0.f) { y = s * y + oms * E; E = y; }
            a.index[o] = colour_index(E, a);
        }
    }
}

}  // namespace ems
