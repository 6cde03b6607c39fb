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
    float*       etmp;       // [channels][B] shaped energy of the final column
    float*       agc;        // [2][channels]: column peak, running level
    int hop, channels, M, Lr, R, ring_cols, B, acc_is_u64;
    float smoothing, db_floor, inv_range, gate_db, agc_strength, agc_lambda;
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

// Column cf = f - R can no longer receive energy once frame f is in: shape it (weights, EMA),
// keep its peak for the AGC, and clear its slot for column cf + ring_cols.
__global__ void stream_shape_kernel(const StreamArgs s) {
    const long long f = s.sstate[0] + 1 - s.M;
    const long long cf = f - s.R;
    if (cf < 0) return;
    const int slot = (int)(cf % s.ring_cols);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < s.B * s.channels;
         e += gridDim.x * blockDim.x) {
        const int ch = e / s.B, k = e - ch * s.B;
        const long long o = ((long long)ch * s.ring_cols + slot) * s.B + k;
        float E = acc_load(s.acc, s.acc_is_u64, o) * s.weight[k];
        if (s.smoothing > 0.f) {
            E = s.smoothing * s.carry[e] + (1.0f - s.smoothing) * E;
            s.carry[e] = E;
        }
        s.etmp[e] = E;
        if (s.agc_strength > 0.f && E > 0.f) peak_max(s.agc, ch, E);
        acc_zero(s.acc, s.acc_is_u64, o);
    }
}

// Colour index of the final column, drawn at E / level^strength when the AGC is on.
__global__ void stream_emit_kernel(const StreamArgs s) {
    if (s.sstate[0] + 1 - s.M - s.R < 0) return;
    PostArgs pa{};
    pa.db_floor = s.db_floor; pa.inv_range = s.inv_range; pa.gate_db = s.gate_db;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < s.B * s.channels;
         e += gridDim.x * blockDim.x) {
        const int ch = e / s.B;
        if (s.agc_strength > 0.f) {
            const float lv = fmaxf(s.agc[ch], s.agc_lambda * s.agc[s.channels + ch]);
            s.out[e] = colour_index(s.etmp[e], lv > 0.f ? powf(lv, -s.agc_strength) : 1.0f, pa);
        } else {
            s.out[e] = colour_index(s.etmp[e], pa);
        }
    }
}

// End of a push: AGC level recurrence, then the counter.
__global__ void stream_advance_kernel(const StreamArgs s) {
    if (s.sstate[0] + 1 - s.M - s.R >= 0)
        for (int ch = 0; ch < s.channels; ++ch) {
            s.agc[s.channels + ch] = fmaxf(s.agc[ch], s.agc_lambda * s.agc[s.channels + ch]);
            s.agc[ch] = 0.f;
        }
    s.sstate[0] += 1;
}

}  // namespace ems
