// stream.cuh — device side of streaming mode (ems_stream_push): sample ring ingest, the
// post-pass of the column that just became final, and the push counter.  All three read the
// push counter from device memory so the whole push is one static CUDA graph.
// Stands in for "start visualizing your system audio" (/root/reference/README.md:36).
#pragma once
#include "common.cuh"
#include "scatter_post.cuh"

namespace ems {

struct StreamArgs {
    long long*   sstate;     // [0] pushes completed
    const float* in;         // [hop][channels] interleaved, this push
    float*       ring;       // [channels][2*Lr] doubled sample ring
    void*        acc;        // [channels][ring_cols][B] accumulator ring
    float*       carry;      // [channels][B] EMA state
    const float* weight;     // [B]
    uint8_t*     out;        // [channels][B] colour index of the final column
    int hop, channels, M, Lr, R, ring_cols, B, acc_is_u64;
    float smoothing, db_floor, inv_range, gate_db;
};

// De-interleaves the hop and writes it twice (pos and pos + Lr) so that any n_fft-long
// window of the ring is contiguous.
__global__ void stream_ingest_kernel(const StreamArgs s) {
    const long long i = s.sstate[0];
    const int wp = (int)(i % s.M) * s.hop;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < s.hop * s.channels;
         e += gridDim.x * blockDim.x) {
        const int smp = e / s.channels, ch = e - smp * s.channels;
        const float v = s.in[e];
        float* r = s.ring + (long long)ch * 2 * s.Lr + wp + smp;
        r[0] = v;
        r[s.Lr] = v;
    }
}

// Column cf = f - R can no longer receive energy once frame f is in: shape it, emit the
// colour index, and clear its slot for column cf + ring_cols.
__global__ void stream_post_kernel(const StreamArgs s) {
    const long long f = s.sstate[0] + 1 - s.M;
    const long long cf = f - s.R;
    if (cf < 0) return;
    const int slot = (int)(cf % s.ring_cols);
    PostArgs pa{};
    pa.db_floor = s.db_floor; pa.inv_range = s.inv_range; pa.gate_db = s.gate_db;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < s.B * s.channels;
         e += gridDim.x * blockDim.x) {
        const int ch = e / s.B, k = e - ch * s.B;
        const long long o = ((long long)ch * s.ring_cols + slot) * s.B + k;
        float E = acc_load(s.acc, s.acc_is_u64, o) * s.weight[k];
        if (s.smoothing > 0.f) {
            E = s.smoothing * s.carry[e] + (1.0f - s.smoothing) * E;
            s.carry[e] = E;
        }
        s.out[e] = colour_index(E, pa);
        if (s.acc_is_u64) reinterpret_cast<unsigned long long*>(s.acc)[o] = 0ull;
        else reinterpret_cast<float*>(s.acc)[o] = 0.f;
    }
}

__global__ void stream_advance_kernel(long long* sstate) { sstate[0] += 1; }

}  // namespace ems
