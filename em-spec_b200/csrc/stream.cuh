// stream.cuh — device side of streaming mode (ems_stream_push): sample ring ingest and the
// post-pass of the column that just became final.  Both read the push counter from device
// memory, so the whole push is one static CUDA graph of three kernels (ingest, STFT, finish);
// the hop is read from, and the column written to, mapped pinned host memory directly.
// Stands in for "start visualizing your system audio" (/root/reference/README.md:36).
#pragma once
#include "common.cuh"
#include "scatter_post.cuh"

namespace ems {

struct StreamArgs {
    long long*   sstate;     // [0] pushes completed, [1] blocks of the finish kernel that are done
    const float* in;         // [hop][channels] interleaved, this push (mapped pinned host memory)
    float*       ring;       // [channels][2*Lr] doubled sample ring
    void*        acc;        // [channels][ring_cols][B] accumulator ring
    float*       carry;      // [channels][B] EMA state
    const float* weight;     // [B]
    uint8_t*     out;        // [channels][B] colour index of the final column (mapped pinned host memory)
    float*       etmp;       // [channels][B] shaped energy of the final column
    float*       agc;        // [2][channels]: (unused), running level
    int hop, channels, M, Lr, R, ring_cols, B, acc_is_u64;
    float smoothing, db_floor, inv_range, gate_db, agc_strength, agc_lambda, agc_target;
    int in_i16;              // format of the hop: 0 fp32, 1 int16 (full scale 32768), 2 packed little-endian int24 (full scale 2^23)
    const uint32_t* lut;     // [256] RGBA colour map (ems_stream_set_colormap), or null
    uint32_t*    out_rgba;   // [channels][B] pixels of the final column (mapped pinned host memory)
};

// De-interleaves the hop and writes it twice (pos and pos + Lr) so that any n_fft-long
// window of the ring is contiguous.
__global__ void stream_ingest_kernel(const StreamArgs s) {
    const long long i = s.sstate[0];
    const int wp = (int)(i % s.M) * s.hop;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < s.hop * s.channels;
         e += gridDim.x * blockDim.x) {
        const int smp = e / s.channels, ch = e - smp * s.channels;
        float v;
        if (s.in_i16 == 2) {
            const unsigned char* b = reinterpret_cast<const unsigned char*>(s.in) + 3 * e;
            v = (float)((int)(((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16)) << 8) >> 8) * (1.0f / 8388608.0f);
        } else {
            v = s.in_i16 ? (float)reinterpret_cast<const int16_t*>(s.in)[e] * (1.0f / 32768.0f) : s.in[e];
        }
        float* r = s.ring + (long long)ch * 2 * s.Lr + wp + smp;
        r[0] = v;
        r[s.Lr] = v;
    }
}

// Finishes the push: grid (channels, slices).  With the AGC on a channel is one block (its peak is a
// block reduction); without it the bins of a channel are sliced over gridDim.y blocks, which cuts
// the latency of this last kernel of the push.  Column cf = f - R can no longer receive energy once
// frame f is in: shape it (weights, EMA), take its peak for the AGC, emit the colour index
// straight into the caller-visible pinned column (mapped host memory: no copy node), clear
// the slot for column cf + ring_cols.  The last block to finish advances the AGC level and the
// push counter (every block has read the counter before it arrives there).
__global__ void __launch_bounds__(1024)
stream_finish_kernel(const StreamArgs s) {
    __shared__ float s_red[32];
    __shared__ float s_scale;
    const long long i = s.sstate[0];
    const long long cf = i + 1 - s.M - s.R;
    const int ch = blockIdx.x, t = threadIdx.x;
    const int k0 = blockIdx.y * blockDim.x + t, kstep = gridDim.y * blockDim.x;
    if (cf >= 0) {
        const int slot = (int)(cf % s.ring_cols);
        float peak = 0.f;
        for (int k = k0; k < s.B; k += kstep) {
            const int e = ch * s.B + k;
            const long long o = ((long long)ch * s.ring_cols + slot) * s.B + k;
            float E = acc_load(s.acc, s.acc_is_u64, o) * s.weight[k];
            if (s.smoothing > 0.f) {
                E = s.smoothing * s.carry[e] + (1.0f - s.smoothing) * E;
                s.carry[e] = E;
            }
            s.etmp[e] = E;
            peak = fmaxf(peak, E);
            acc_zero(s.acc, s.acc_is_u64, o);
        }
        float scale = 1.0f;
        if (s.agc_strength > 0.f) {
            peak = warp_max(peak);
            if ((t & 31) == 0) s_red[t >> 5] = peak;
            __syncthreads();
            if (t < 32) {
                float v = t < (int)(blockDim.x >> 5) ? s_red[t] : 0.f;
                v = warp_max(v);
                if (t == 0) {
                    const float lv = fmaxf(v, s.agc_lambda * s.agc[s.channels + ch]);
                    s.agc[s.channels + ch] = lv;                       // level recurrence
                    s_scale = lv > 0.f ? s.agc_target * powf(lv, -s.agc_strength) : 1.0f;
                }
            }
            __syncthreads();
            scale = s_scale;
        } else {
            __syncthreads();          // etmp written by this block is re-read below
        }
        PostArgs pa{};
        pa.db_floor = s.db_floor; pa.inv_range = s.inv_range; pa.gate_db = s.gate_db;
        for (int k = k0; k < s.B; k += kstep) {
            const int e = ch * s.B + k;
            const uint8_t ci = s.agc_strength > 0.f ? colour_index(s.etmp[e], scale, pa) : colour_index(s.etmp[e], pa);
            s.out[e] = ci;
            if (s.lut) s.out_rgba[e] = __ldg(s.lut + ci);
        }
    }
    __syncthreads();
    if (t == 0) {
        __threadfence();
        if (atomicAdd(reinterpret_cast<unsigned long long*>(s.sstate + 1), 1ull) == (unsigned long long)(gridDim.x * gridDim.y) - 1) {
            s.sstate[1] = 0;
            s.sstate[0] = i + 1;
        }
    }
}

}  // namespace ems
