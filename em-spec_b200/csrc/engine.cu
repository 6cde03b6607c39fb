// engine.cu — host side of the engine and the C-ABI of include/emspec.h.
// C++ host, no torch types; one handle = one CUDA stream.
#include "../../include/emspec.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"
#include "scatter_post.cuh"
#include "scatter_sorted.cuh"
#include "stft_generic.cuh"
#include "stft_r16.cuh"
#include "stft_r64.cuh"
#include "stft_r32.cuh"
#include "stream.cuh"

namespace ems {

constexpr double kPi = 3.14159265358979323846;
constexpr double kLowEndCornerHz = 200.0;   // oracle/reassign_oracle.py LOW_END_CORNER_HZ
constexpr double kTopDb = 0.0;

struct DevBuf {
    void*  p = nullptr;
    size_t bytes = 0;
};

}  // namespace ems

struct ems_handle {
    ems_params prm{};
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;      // the stream calls run on
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    float*  thw = nullptr;              // [N] th' window
    float2* tw = nullptr;               // [N]
    float*  weight = nullptr;           // [B]
    ems::DevBuf acc, flags, carry, ema_local, ema_carry, big_scratch, lut, colscale, agc_level, post_mode, sort_buf;
    bool post_mode_init = false;
    struct HostPipe {                   // ems_process_host*: two of everything, chunk c uses set c & 1
        ems::DevBuf pcm[2], raw[2], idx[2], grid[2];   // fp32 planar chunk, raw int16/int24 chunk, staging images
        cudaEvent_t ev_in[2]{}, ev_done[2]{}, ev_out[2]{}, ev_start{};
        bool events = false;
    } hp;
    bool acc_clean = false;             // accumulator and dirty flags are all zero (kept so by the post-pass)
    cudaEvent_t ev[EMS_STAGE_COUNT][2]{};
    bool ev_valid[EMS_STAGE_COUNT]{};
    uint64_t launches = 0;
    // streaming state (allocated on the first push)
    struct Stream {
        bool ready = false;
        cudaGraphExec_t graph = nullptr;
        long long* sstate = nullptr;     // device push counter
        float* in_dev = nullptr;         // device alias of in_pin
        float* ring = nullptr;           // [channels][2*Lr]
        void* acc = nullptr;             // [channels][ring_cols][B]
        float* carry = nullptr;          // [channels][B]
        uint8_t* out_dev = nullptr;      // device alias of out_pin
        float* etmp = nullptr;           // [channels][B]
        float* agc = nullptr;            // [2][channels]
        float* in_pin = nullptr;         // pinned staging
        uint8_t* out_pin = nullptr;
        uint32_t* rgba_pin = nullptr;    // pinned [channels][B] pixels of the final column, and its device alias
        uint32_t* rgba_dev = nullptr;
        bool last_ready = false;         // the last push delivered a column
        int M = 0, Lr = 0, R = 0, ring_cols = 0;
        int in_i16 = 0;                  // format of the hop the captured graph reads (0: fp32, 1: int16, 2: packed int24)
        long long pushes = 0;            // host mirror of the device counter
        size_t acc_bytes = 0;
    } st;
    ems::DevBuf stream_lut;             // ems_stream_set_colormap: configuration, outlives the stream state
    bool stream_lut_on = false;
    struct FusedState {                 // in-kernel post-pass of ems_process_grid (common.cuh FusedPost)
        ems::DevBuf ready, done;        // per-tile arrival counters (self-resetting) and epoch stamps
        int epoch = 0;
        bool counters_clean = false;    // ready[] is all zero (a failed launch leaves it unknown)
        int mode = 0;                   // env EMS_FUSED_POST=1 switches it on in a -DEMS_FUSED_POST=1 build (experiment: 5x slower)
        int ring = 8192;                // EMS_FUSED_RING: columns of the accumulator ring
    } fz;
    bool force_generic = false;         // EMS_FORCE_GENERIC=1: bypass the tuned kernels (A/B tests)
    int kernel_variant = 0;             // EMS_KERNEL_VARIANT: experimental kernel selection (A/B runs); 0 = default
    int max_ctas = 0;                   // EMS_MAX_CTAS: cap on the persistent grids (schedule-perturbation tests); 0 = one per SM
    char err[256] = "";
};

namespace ems {

static ems_status fail(ems_handle* h, ems_status s, const char* fmt, ...) {
    if (h) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(h->err, sizeof(h->err), fmt, ap);
        va_end(ap);
    }
    return s;
}

#define EMS_CUDA(h, call)                                                              \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess)                                                         \
            return ems::fail((h), e_ == cudaErrorMemoryAllocation ? EMS_ERR_NOMEM      \
                                                                  : EMS_ERR_CUDA,     \
                             "%s: %s", #call, cudaGetErrorString(e_));                 \
    } while (0)

static ems_status ensure(ems_handle* h, DevBuf& b, size_t bytes) {
    if (b.bytes >= bytes) return EMS_OK;
    if (b.p) { EMS_CUDA(h, cudaStreamSynchronize(h->stream)); EMS_CUDA(h, cudaFree(b.p)); b = DevBuf{}; }
    EMS_CUDA(h, cudaMalloc(&b.p, bytes));
    b.bytes = bytes;
    return EMS_OK;
}

static bool valid_params(const ems_params& p) {
    if (p.n_fft < 256 || p.n_fft > 32768 || (p.n_fft & (p.n_fft - 1))) return false;
    if (p.hop < 1 || p.hop > p.n_fft) return false;
    if (!(p.sample_rate > 0.f) || p.channels < 1 || p.channels > 65535) return false;
    if (!(p.db_range > 0.f) || !(p.gain >= 0.f) || !(p.low_end_boost > 0.f)) return false;
    if (!(p.smoothing >= 0.f) || !(p.smoothing < 1.f)) return false;
    if (!std::isfinite(p.noise_gate_db)) return false;
    if (p.display_rows < 0 || p.display_rows == 1 || p.display_rows > 65536) return false;
    if (!(p.freq_scale >= 0.f) || !(p.freq_scale <= 4.f)) return false;
    if (!(p.agc_strength >= 0.f) || !(p.agc_strength <= 1.f)) return false;
    if (!(p.brightness > 0.f) || !(p.brightness <= 1.f)) return false;
    return true;
}

static int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }

// "Brightness": the automatic gain draws the running level at T = 10^(-(1 - brightness) range / 10)
static float agc_target(const ems_params& p) {
    return (float)std::pow(10.0, -(1.0 - (double)p.brightness) * (double)p.db_range / 10.0);
}

static int rows_of(const ems_params& p);
static double warp_a_of(const ems_params& p);

// weight[r] = gain^2 * w_low(f_r), f_r = centre frequency of output row r
static ems_status upload_display(ems_handle* h) {
    const int R = rows_of(h->prm);
    std::vector<float> w(R);
    const double g2 = (double)h->prm.gain * (double)h->prm.gain;
    const double nyq = 0.5 * (double)h->prm.sample_rate, a = warp_a_of(h->prm);
    for (int r = 0; r < R; ++r) {
        double f;
        if (h->prm.display_rows > 0) {
            const double u = (double)r / (double)(R - 1);
            f = nyq * (a > 1e-6 ? std::expm1(u * std::log1p(a)) / a : u);
        } else {
            f = (double)r * (double)h->prm.sample_rate / (double)h->prm.n_fft;
        }
        const double q = f / kLowEndCornerHz;
        w[r] = (float)(g2 * (1.0 + ((double)h->prm.low_end_boost - 1.0) / (1.0 + q * q)));
    }
    EMS_CUDA(h, cudaMemcpyAsync(h->weight, w.data(), R * sizeof(float), cudaMemcpyHostToDevice,
                                h->stream));
    EMS_CUDA(h, cudaStreamSynchronize(h->stream));
    return EMS_OK;
}

// Output rows per column: one per bin, or display_rows on the warped frequency axis.
static int rows_of(const ems_params& p) { return p.display_rows > 0 ? p.display_rows : p.n_fft / 2 + 1; }
static double warp_a_of(const ems_params& p) { return std::pow(10.0, 2.0 * (double)p.freq_scale) - 1.0; }

static long long frames_of(const ems_params& p, size_t S) {
    return S < (size_t)p.n_fft ? 0 : 1 + (long long)((S - (size_t)p.n_fft) / (size_t)p.hop);
}

// ---------------------------------------------------------------- kernel dispatch
template <int LOG2N>
static ems_status launch_generic(ems_handle* h, const StftArgs& a) {
    constexpr int N = 1 << LOG2N;
    constexpr int THREADS = 256;
    const size_t smem = (size_t)N * 8;
    auto kern = stft_reassign_generic<LOG2N, THREADS>;
    EMS_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    EMS_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    if (occ < 1) return fail(h, EMS_ERR_UNSUPPORTED, "n_fft=%d does not fit shared memory", N);
    const long long total = (a.f_end - a.f_begin) * a.channels;
    long long grid = (long long)h->sm_count * occ;
    if (grid > total) grid = total;
    if (grid < 1) return EMS_OK;
    kern<<<(unsigned)grid, THREADS, smem, h->stream>>>(a);
    ++h->launches;
    EMS_CUDA(h, cudaGetLastError());
    return EMS_OK;
}

template <int LOG2N>
static ems_status launch_big(ems_handle* h, const StftArgs& a) {
    constexpr int N = 1 << LOG2N;
    constexpr int THREADS = 1024;
    const size_t smem = (size_t)N * 4;
    auto kern = stft_reassign_big<LOG2N, THREADS>;
    EMS_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long total = (a.f_end - a.f_begin) * a.channels;
    long long grid = h->max_ctas > 0 ? std::min(h->max_ctas, h->sm_count) : h->sm_count;
    if (grid > total) grid = total;
    if (grid < 1) return EMS_OK;
    const size_t need = (size_t)h->sm_count * (N / 2 + 3) * sizeof(float2);
    if (h->big_scratch.bytes < need) {     // never reallocated inside a stream capture: sized for all SMs
        ems_status s = ensure(h, h->big_scratch, need);
        if (s != EMS_OK) return s;
    }
    kern<<<(unsigned)grid, THREADS, smem, h->stream>>>(a, (float2*)h->big_scratch.p);
    ++h->launches;
    EMS_CUDA(h, cudaGetLastError());
    return EMS_OK;
}

// Tuned kernels for n_fft = 256 R (R = 4, 8, 16): persistent, one CTA per SM, 48/R workers x 8R threads.
template <int R>
static ems_status launch_r16(ems_handle* h, const StftArgs& a) {
    using C = r16::Cfg<R>;
    int tile_T = C::tile_frames(a.hop);
    if (tile_T < 1) return EMS_ERR_UNSUPPORTED;       // not a failure: the caller falls back to the generic kernel
    // short inputs: smaller tiles so that every SM gets one
    const long long per_ch = a.f_end - a.f_begin;
    const long long want = (per_ch * a.channels + h->sm_count - 1) / h->sm_count;
    constexpr int kPer = C::kWorkers * C::kG;          // frames in flight per CTA
    const long long t = ((want + kPer - 1) / kPer) * kPer;
    if (t < tile_T) tile_T = (int)std::max<long long>(t, 1);
    const size_t smem = (size_t)C::kFixedBytes + 2 * (size_t)C::kTileFloats * sizeof(float) + r16::kSyncBytes;
    void (*kern)(const StftArgs, const int) =
        a.mode == kStorePoints ? r16::stft_reassign_r16<R, kStorePoints>
        : a.mode == kDepositU64 ? r16::stft_reassign_r16<R, kDepositU64>
                                : r16::stft_reassign_r16<R, kDepositF32>;
    EMS_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, r16::kMaxSmem));
    const long long n_tiles = ((per_ch + tile_T - 1) / tile_T) * a.channels;
    long long grid = h->max_ctas > 0 ? std::min(h->max_ctas, h->sm_count) : h->sm_count;
    if (grid > n_tiles) grid = n_tiles;
    if (grid < 1) return EMS_OK;
    if (a.fp.vring > 0) {
        // in-kernel post-pass: blocks of tile_T columns, shaped by the CTA that completes the last of the
        // 2 nb + 1 tiles whose deposits can reach them.  CTAs may wait for one another (ring space), so the
        // launch is cooperative: every CTA is resident or the launch fails.
        StftArgs af = a;
        const long long Rc = (C::N / 2 + a.hop - 1) / a.hop;
        af.fp.nb = (int)((Rc + tile_T - 1) / tile_T);
        if ((long long)af.fp.vring < 4 * ((2LL * af.fp.nb + 1) * tile_T + 2 * Rc))
            return fail(h, EMS_ERR_STATE, "accumulator ring of %d columns is too short for hop %d", af.fp.vring, a.hop);
        const size_t need = (size_t)n_tiles * sizeof(int);
        if (h->fz.ready.bytes < need || h->fz.done.bytes < need) h->fz.counters_clean = false;
        ems_status s;
        if ((s = ensure(h, h->fz.ready, need)) != EMS_OK || (s = ensure(h, h->fz.done, need)) != EMS_OK) return s;
        if (!h->fz.counters_clean) {
            EMS_CUDA(h, cudaMemsetAsync(h->fz.ready.p, 0, h->fz.ready.bytes, h->stream));
            EMS_CUDA(h, cudaMemsetAsync(h->fz.done.p, 0, h->fz.done.bytes, h->stream));
            h->fz.epoch = 0;
        }
        h->fz.counters_clean = false;                  // until the caller has seen the launch succeed
        af.fp.ready = (int*)h->fz.ready.p;
        af.fp.done = (int*)h->fz.done.p;
        af.fp.epoch = ++h->fz.epoch;
        int tt = tile_T;
        void* args[] = {(void*)&af, (void*)&tt};
        EMS_CUDA(h, cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)grid), dim3(C::kCta), args, smem, h->stream));
        ++h->launches;
        return EMS_OK;
    }
    kern<<<(unsigned)grid, C::kCta, smem, h->stream>>>(a, tile_T);
    ++h->launches;
    EMS_CUDA(h, cudaGetLastError());
    return EMS_OK;
}

// n_fft = 4096, single-exchange variant (64 x 64 in two passes, 4 workers x 64 threads).
static ems_status launch_r64(ems_handle* h, const StftArgs& a) {
    int tile_T = r64::tile_frames(a.hop);
    if (tile_T < 1) return EMS_ERR_UNSUPPORTED;
    const long long per_ch = a.f_end - a.f_begin;
    const long long want = (per_ch * a.channels + h->sm_count - 1) / h->sm_count;
    const long long t = ((want + r64::kWorkers - 1) / r64::kWorkers) * r64::kWorkers;
    if (t < tile_T) tile_T = (int)std::max<long long>(t, 1);
    const size_t smem = (size_t)r64::kFixedBytes + 2 * (size_t)r64::kTileFloats * sizeof(float) + r16::kSyncBytes;
    void (*kern)(const StftArgs, const int) =
        a.mode == kStorePoints ? r64::stft_reassign_r64<kStorePoints>
        : a.mode == kDepositU64 ? r64::stft_reassign_r64<kDepositU64>
                                : r64::stft_reassign_r64<kDepositF32>;
    EMS_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, r16::kMaxSmem));
    const long long n_tiles = ((per_ch + tile_T - 1) / tile_T) * a.channels;
    long long grid = h->max_ctas > 0 ? std::min(h->max_ctas, h->sm_count) : h->sm_count;
    if (grid > n_tiles) grid = n_tiles;
    if (grid < 1) return EMS_OK;
    kern<<<(unsigned)grid, r64::kThreads64, smem, h->stream>>>(a, tile_T);
    ++h->launches;
    EMS_CUDA(h, cudaGetLastError());
    return EMS_OK;
}

// Tuned kernels for n_fft = 4096 R0 (R0 = 2, 4): frames read straight from global memory.
template <int R0>
static ems_status launch_r16_large(ems_handle* h, const StftArgs& a) {
    using C = r16::CfgL<R0>;
    void (*kern)(const StftArgs) =
        a.mode == kStorePoints ? r16::stft_reassign_r16_large<R0, kStorePoints>
        : a.mode == kDepositU64 ? r16::stft_reassign_r16_large<R0, kDepositU64>
                                : r16::stft_reassign_r16_large<R0, kDepositF32>;
    EMS_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    const long long total = (a.f_end - a.f_begin) * a.channels;
    long long grid = h->max_ctas > 0 ? std::min(h->max_ctas, h->sm_count) : h->sm_count;
    if (grid > total) grid = total;
    if (grid < 1) return EMS_OK;
    kern<<<(unsigned)grid, C::kThreads, C::kSmemBytes, h->stream>>>(a);
    ++h->launches;
    EMS_CUDA(h, cudaGetLastError());
    return EMS_OK;
}

// Experiment (EMS_KERNEL_VARIANT=32): n_fft = 8192 with three workers of 128 threads, radix-32 first pass, X in place.
static ems_status launch_8192_w3(ems_handle* h, const StftArgs& a) {
    using C = r16::Cfg8kW3;
    void (*kern)(const StftArgs) =
        a.mode == kStorePoints ? r16::stft_reassign_8192_w3<kStorePoints>
        : a.mode == kDepositU64 ? r16::stft_reassign_8192_w3<kDepositU64>
                                : r16::stft_reassign_8192_w3<kDepositF32>;
    EMS_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    const long long total = (a.f_end - a.f_begin) * a.channels;
    long long grid = h->max_ctas > 0 ? std::min(h->max_ctas, h->sm_count) : h->sm_count;
    if (grid > total) grid = total;
    if (grid < 1) return EMS_OK;
    kern<<<(unsigned)grid, C::kThreads, C::kSmemBytes, h->stream>>>(a);
    ++h->launches;
    EMS_CUDA(h, cudaGetLastError());
    return EMS_OK;
}

// n_fft = 32768: two sequential 16384-point complex FFTs per frame, 2 X parked in an L2-resident scratch.
static ems_status launch_r16_32k(ems_handle* h, const StftArgs& a) {
    using C = r16::CfgL<4>;
    void (*kern)(const StftArgs, float2*) =
        a.mode == kStorePoints ? r16::stft_reassign_r16_32k<kStorePoints>
        : a.mode == kDepositU64 ? r16::stft_reassign_r16_32k<kDepositU64>
                                : r16::stft_reassign_r16_32k<kDepositF32>;
    constexpr int smem = (C::kZBuf + C::kZtab + r16::kT2) * 8;
    EMS_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const long long total = (a.f_end - a.f_begin) * a.channels;
    long long grid = h->max_ctas > 0 ? std::min(h->max_ctas, h->sm_count) : h->sm_count;
    if (grid > total) grid = total;
    if (grid < 1) return EMS_OK;
    const size_t need = (size_t)h->sm_count * r16::k32kScratch * sizeof(float2);
    if (h->big_scratch.bytes < need) {     // never reallocated inside a stream capture: stream_init sizes it
        ems_status s = ensure(h, h->big_scratch, need);
        if (s != EMS_OK) return s;
    }
    kern<<<(unsigned)grid, 512, smem, h->stream>>>(a, (float2*)h->big_scratch.p);
    ++h->launches;
    EMS_CUDA(h, cudaGetLastError());
    return EMS_OK;
}

static ems_status launch_stft(ems_handle* h, const StftArgs& a) {
    if (!h->force_generic) {
        ems_status s = EMS_ERR_UNSUPPORTED;
        switch (h->prm.n_fft) {
            case 256: s = launch_r16<1>(h, a); break;
            case 512: s = launch_r16<2>(h, a); break;
            case 1024: s = launch_r16<4>(h, a); break;
            case 2048: s = launch_r16<8>(h, a); break;
            case 4096: s = h->kernel_variant == 64 ? launch_r64(h, a) : launch_r16<16>(h, a); break;
            case 8192: s = h->kernel_variant == 32 ? launch_8192_w3(h, a) : launch_r16_large<2>(h, a); break;
            case 16384: s = launch_r16_large<4>(h, a); break;
            case 32768: s = launch_r16_32k(h, a); break;
            default: break;
        }
        if (s != EMS_ERR_UNSUPPORTED) return s;
    }
    switch (ilog2(h->prm.n_fft)) {
        case 8:  return launch_generic<8>(h, a);
        case 9:  return launch_generic<9>(h, a);
        case 10: return launch_generic<10>(h, a);
        case 11: return launch_generic<11>(h, a);
        case 12: return launch_generic<12>(h, a);
        case 13: return launch_generic<13>(h, a);
        case 14: return launch_generic<14>(h, a);
        case 15: return launch_big<15>(h, a);
        default: return fail(h, EMS_ERR_UNSUPPORTED, "n_fft=%d has no kernel in this build",
                             h->prm.n_fft);
    }
}

static StftArgs make_args(ems_handle* h, const float* pcm, size_t S, long long F) {
    StftArgs a{};
    a.pcm = pcm; a.S = (long long)S; a.F = F; a.f_begin = 0; a.f_end = F;
    a.channels = h->prm.channels; a.hop = h->prm.hop;
    a.thw = h->thw; a.tw = h->tw;
    a.gate_lin = (float)std::pow(10.0, (double)h->prm.noise_gate_db / 10.0);
    a.inv_hop = 1.0f / (float)h->prm.hop;
    a.reassign = (h->prm.flags & EMS_FLAG_REASSIGN) ? 1 : 0;
    a.rows = rows_of(h->prm);
    a.inv_half = 2.0f / (float)h->prm.n_fft;
    if (h->prm.display_rows > 0) {
        const double wa = warp_a_of(h->prm);
        a.warp_mode = wa > 1e-6 ? 2 : 1;
        a.warp_a = (float)wa;
        a.warp_c = (float)(wa > 1e-6 ? (a.rows - 1) / std::log1p(wa) : (double)(a.rows - 1));
    }
    return a;
}

static PostArgs make_post(ems_handle* h, long long F, float* grid, uint8_t* index) {
    PostArgs p{};
    p.acc = h->acc.p;
    p.flags = (unsigned char*)h->flags.p;
    p.NB = flag_blocks(rows_of(h->prm));
    p.acc_is_u64 = (h->prm.flags & EMS_FLAG_DETERMINISTIC) ? 1 : 0;
    p.grid = grid; p.index = index; p.weight = h->weight;
    p.carry = (float*)h->carry.p;
    p.F = F; p.col_begin = 0; p.col_end = F;
    p.acc_cols = F; p.acc_mask = -1; p.out_cols = F; p.out_col0 = 0;
    p.B = rows_of(h->prm); p.channels = h->prm.channels;
    p.smoothing = h->prm.smoothing;
    p.db_floor = (float)(kTopDb - (double)h->prm.db_range);
    p.inv_range = 255.0f / h->prm.db_range;
    p.gate_db = h->prm.noise_gate_db;
    return p;
}

static ems_status clear_flags(ems_handle* h, const PostArgs& p);

// Zero-fills columns [col_begin, col_end) of every channel of a [channels][F][col_bytes] array:
// one memset when the range is the whole stream, one pitched memset otherwise (never one call
// per channel: a batch maps clips to channels, up to 65535 of them).
// `cols` = columns per channel of the array (p.F for colscale, p.out_cols for grid / index),
// `col0` = the stream column its column 0 holds.
static cudaError_t zero_cols(ems_handle* h, void* base, size_t col_bytes, const PostArgs& p, long long cols,
                             long long col0) {
    const size_t ncols = (size_t)(p.col_end - p.col_begin);
    char* b0 = (char*)base + (size_t)(p.col_begin - col0) * col_bytes;
    if (p.channels == 1 || ncols == (size_t)cols)
        return cudaMemsetAsync(b0, 0, (p.channels - 1) * (size_t)cols * col_bytes + ncols * col_bytes, h->stream);
    return cudaMemset2DAsync(b0, (size_t)cols * col_bytes, 0, ncols * col_bytes, (size_t)p.channels, h->stream);
}

// Post-pass over columns [col_begin, col_end) of every channel; h->carry holds the EMA
// state entering col_begin and leaves with the state after col_end - 1.
static ems_status run_post(ems_handle* h, PostArgs p) {
    const long long ncols = p.col_end - p.col_begin;
    if (ncols <= 0) return EMS_OK;
    const bool ema = p.smoothing > 0.f && p.index;
    const bool agc = h->prm.agc_strength > 0.f && p.index;
    const float lambda = std::exp(-(float)h->prm.hop / (h->prm.sample_rate * kAgcReleaseSeconds));
    p.colscale = nullptr;
    if (agc) {      // column peaks are accumulated into colscale, then turned into level^-strength
        ems_status s = ensure(h, h->colscale, (size_t)p.channels * p.F * sizeof(float));
        if (s != EMS_OK) return s;
        p.colscale = (float*)h->colscale.p;
        EMS_CUDA(h, zero_cols(h, p.colscale, sizeof(float), p, p.F, 0));
    }
    auto agc_scan = [&]() {
        agc_scan_kernel<<<p.channels, 1024, 0, h->stream>>>(p.colscale, (float*)h->agc_level.p, p.F,
                                                            p.col_begin, p.col_end, lambda,
                                                            h->prm.agc_strength, agc_target(h->prm));
        ++h->launches;
    };
    if (!ema) {
        // no recurrence along time: zero-fill the outputs, then visit only the dirty blocks
        if (p.index) EMS_CUDA(h, zero_cols(h, p.index, (size_t)p.B, p, p.out_cols, p.out_col0));
        if (p.grid) EMS_CUDA(h, zero_cols(h, p.grid, (size_t)p.B * sizeof(float), p, p.out_cols, p.out_col0));
        if ((ncols + 255) / 256 > 65535) return fail(h, EMS_ERR_UNSUPPORTED, "more than 16.7 M columns in one post-pass");
        const dim3 g((unsigned)(p.channels * p.NB), (unsigned)((ncols + 255) / 256));
        if (agc) {
            post_sparse_kernel<<<g, 256, 0, h->stream>>>(p, 1, nullptr);
            ++h->launches;
            agc_scan();
            post_sparse_kernel<<<g, 256, 0, h->stream>>>(p, 0, nullptr);
            ++h->launches;
            EMS_CUDA(h, cudaGetLastError());
            return EMS_OK;
        }
        // a mostly empty image is visited block by block; a mostly full one is streamed column by column.
        // The choice is made on the device from the dirty flags; the kernel not chosen exits at once.
        ems_status s = ensure(h, h->post_mode, 16);
        if (s != EMS_OK) return s;
        if (!h->post_mode_init) {
            EMS_CUDA(h, cudaMemsetAsync(h->post_mode.p, 0, 16, h->stream));
            h->post_mode_init = true;
        }
        unsigned long long* cnt = (unsigned long long*)h->post_mode.p;
        int* mode = (int*)(cnt + 1);
        const long long nflags = ncols * p.channels * p.NB;
        const unsigned gb = (unsigned)std::min<long long>((nflags / 4 + 255) / 256 + 1, (long long)h->sm_count * 8);
        post_density_kernel<<<gb, 256, 0, h->stream>>>(p, cnt, mode, 0);
        post_density_kernel<<<1, 32, 0, h->stream>>>(p, cnt, mode, 1);
        post_sparse_kernel<<<g, 256, 0, h->stream>>>(p, 0, mode);
        if (p.B <= 4096) {
            const unsigned gd = (unsigned)std::min<long long>(ncols * p.channels, (long long)h->sm_count * 16);
            if (p.acc_is_u64) post_dense_kernel<8><<<gd, 128, 0, h->stream>>>(p, mode);
            else post_dense_kernel<16><<<gd, 128, 0, h->stream>>>(p, mode);     // 4-byte cells: twice the loads in flight
            ++h->launches;
        }
        clear_flags_if_kernel<<<(unsigned)std::min<long long>((nflags + 255) / 256, 4096LL), 256, 0, h->stream>>>(
            p.flags, p.acc_cols, p.acc_mask, p.col_begin, p.col_end, p.channels * p.NB, mode);
        h->launches += 4;
        EMS_CUDA(h, cudaGetLastError());
        return EMS_OK;
    }
    // With smoothing the EMA runs along time: 256-column chunks, local pass + carries + emit,
    // then the flags of the covered columns are cleared.
    const int chunk_cols = kPostChunk;
    const long long n_chunks = (ncols + chunk_cols - 1) / chunk_cols;
    const dim3 blk(128);
    const dim3 grd((unsigned)n_chunks, (p.B + 127) / 128, p.channels);
    const size_t bytes = (size_t)p.channels * n_chunks * p.B * sizeof(float);
    ems_status s;
    if ((s = ensure(h, h->ema_local, bytes)) != EMS_OK) return s;
    if ((s = ensure(h, h->ema_carry, bytes)) != EMS_OK) return s;
    p.carry = (float*)h->carry.p;
    post_ema_local_kernel<<<grd, blk, 0, h->stream>>>(p, (float*)h->ema_local.p, (int)n_chunks);
    const long long last = ncols - (n_chunks - 1) * kPostChunk;
    post_ema_carry_kernel<<<dim3((p.B + 127) / 128, p.channels), blk, 0, h->stream>>>(
        p, (const float*)h->ema_local.p, (float*)h->ema_carry.p, (int)n_chunks,
        (float)std::pow((double)p.smoothing, (double)kPostChunk),
        (float)std::pow((double)p.smoothing, (double)last));
    h->launches += 2;
    const float* carry_in = (const float*)h->ema_carry.p;
    if (agc) {
        post_emit_kernel<<<grd, blk, 0, h->stream>>>(p, carry_in, (int)n_chunks, chunk_cols, 1);
        ++h->launches;
        agc_scan();
    }
    post_emit_kernel<<<grd, blk, 0, h->stream>>>(p, carry_in, (int)n_chunks, chunk_cols, 0);
    ++h->launches;
    if ((s = clear_flags(h, p)) != EMS_OK) return s;
    EMS_CUDA(h, cudaGetLastError());
    return EMS_OK;
}

// The accumulator [channels][F][B] and its dirty flags [channels][F][NB].  Both are all zero
// between calls: deposits flag the 64-bin blocks they touch and the emit pass clears exactly
// those, so no call pays a 22 GB memset.  A call that fails midway leaves acc_clean false
// and the next one starts with a full clear.
static ems_status prepare_acc(ems_handle* h, size_t cells, size_t rows, int B) {   // rows = channels * columns
    const bool det = h->prm.flags & EMS_FLAG_DETERMINISTIC;
    const size_t need = cells * (det ? 8 : 4), need_f = rows * flag_blocks(B);
    if (h->acc.bytes < need || h->flags.bytes < need_f) h->acc_clean = false;
    ems_status s;
    if ((s = ensure(h, h->acc, need)) != EMS_OK) return s;
    if ((s = ensure(h, h->flags, need_f)) != EMS_OK) return s;
    if (!h->acc_clean) {
        EMS_CUDA(h, cudaMemsetAsync(h->acc.p, 0, h->acc.bytes, h->stream));
        EMS_CUDA(h, cudaMemsetAsync(h->flags.p, 0, h->flags.bytes, h->stream));
    }
    h->acc_clean = false;   // until the post-pass has covered every column of this call
    return EMS_OK;
}

// Clears the dirty flags of columns [c0, c1) of every channel after their emit pass.
static ems_status clear_flags(ems_handle* h, const PostArgs& p) {
    const int rows = p.NB * p.channels;
    const long long c0 = p.col_begin, c1 = p.col_end;
    if (c1 - c0 >= p.acc_cols) {       // every column (or every slot of the ring)
        EMS_CUDA(h, cudaMemsetAsync(p.flags, 0, (size_t)rows * p.acc_cols, h->stream));
        return EMS_OK;
    }
    long long blocks = ((c1 - c0) * rows + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    clear_flags_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(p.flags, p.acc_cols, p.acc_mask, c0, c1, rows);
    ++h->launches;
    EMS_CUDA(h, cudaGetLastError());
    return EMS_OK;
}

static ems_status reset_carry(ems_handle* h) {
    const size_t bytes = (size_t)h->prm.channels * rows_of(h->prm) * sizeof(float);
    ems_status s = ensure(h, h->carry, bytes);
    if (s != EMS_OK) return s;
    EMS_CUDA(h, cudaMemsetAsync(h->carry.p, 0, bytes, h->stream));
    if ((s = ensure(h, h->agc_level, (size_t)h->prm.channels * sizeof(float))) != EMS_OK) return s;
    EMS_CUDA(h, cudaMemsetAsync(h->agc_level.p, 0, (size_t)h->prm.channels * sizeof(float), h->stream));
    return EMS_OK;
}

// Stage brackets: CUDA events for ems_stage_ms and an NVTX range (header-only NVTX3: a no-op
// unless a profiler is attached) so that timelines show a1-a3 / a4 / a5 by name.
static const char* const kStageName[EMS_STAGE_COUNT] = {"ems:stft+reassign", "ems:scatter", "ems:post"};
static void stage_begin(ems_handle* h, int st) {
    nvtxRangePushA(kStageName[st]);
    cudaEventRecord(h->ev[st][0], h->stream);
}
static void stage_abort() { nvtxRangePop(); }      // error return between stage_begin and stage_end
static void stage_end(ems_handle* h, int st) {
    cudaEventRecord(h->ev[st][1], h->stream);
    h->ev_valid[st] = true;
    nvtxRangePop();
}

static ems_status finish(ems_handle* h) {
    if (h->prm.flags & EMS_FLAG_SYNC) EMS_CUDA(h, cudaStreamSynchronize(h->stream));
    return EMS_OK;
}


// ---------------------------------------------------------------- streaming
static void stream_free(ems_handle* h) {
    auto& st = h->st;
    if (st.graph) cudaGraphExecDestroy(st.graph);
    for (void* p : {(void*)st.sstate, (void*)st.ring, st.acc, (void*)st.carry, (void*)st.etmp, (void*)st.agc})
        if (p) cudaFree(p);
    if (st.in_pin) cudaFreeHost(st.in_pin);
    if (st.out_pin) cudaFreeHost(st.out_pin);
    if (st.rgba_pin) cudaFreeHost(st.rgba_pin);
    st = ems_handle::Stream{};
}

static ems_status stream_zero(ems_handle* h) {
    auto& st = h->st;
    const int C = h->prm.channels, B = rows_of(h->prm);
    EMS_CUDA(h, cudaMemsetAsync(st.sstate, 0, 2 * sizeof(long long), h->stream));
    EMS_CUDA(h, cudaMemsetAsync(st.ring, 0, sizeof(float) * C * 2 * st.Lr, h->stream));
    EMS_CUDA(h, cudaMemsetAsync(st.acc, 0, st.acc_bytes, h->stream));
    EMS_CUDA(h, cudaMemsetAsync(st.carry, 0, sizeof(float) * C * B, h->stream));
    EMS_CUDA(h, cudaMemsetAsync(st.agc, 0, sizeof(float) * 2 * C, h->stream));
    EMS_CUDA(h, cudaStreamSynchronize(h->stream));
    st.pushes = 0;
    st.last_ready = false;
    return EMS_OK;
}

static ems_status stream_init(ems_handle* h) {
    auto& st = h->st;
    const int N = h->prm.n_fft, H = h->prm.hop, C = h->prm.channels, B = rows_of(h->prm);
    const bool det = h->prm.flags & EMS_FLAG_DETERMINISTIC;
    st.M = (N + H - 1) / H;
    st.Lr = st.M * H;
    st.R = (N / 2 + H - 1) / H;
    st.ring_cols = 1;                       // >= 2 R + 1 columns, a power of two (deposits mask the column)
    while (st.ring_cols < 2 * st.R + 1) st.ring_cols *= 2;
    st.acc_bytes = (size_t)C * st.ring_cols * B * (det ? 8 : 4);
    EMS_CUDA(h, cudaMalloc(&st.sstate, 2 * sizeof(long long)));
    EMS_CUDA(h, cudaMalloc(&st.ring, sizeof(float) * C * 2 * st.Lr));
    EMS_CUDA(h, cudaMalloc(&st.acc, st.acc_bytes));
    EMS_CUDA(h, cudaMalloc(&st.carry, sizeof(float) * C * B));
    EMS_CUDA(h, cudaMalloc(&st.etmp, sizeof(float) * C * B));
    EMS_CUDA(h, cudaMalloc(&st.agc, sizeof(float) * 2 * C));
    // staging the device reads / writes in place (mapped pinned memory): no copy nodes in the graph
    EMS_CUDA(h, cudaHostAlloc(&st.in_pin, sizeof(float) * H * C, cudaHostAllocMapped));
    EMS_CUDA(h, cudaHostAlloc(&st.out_pin, (size_t)C * B, cudaHostAllocMapped));
    EMS_CUDA(h, cudaHostGetDevicePointer((void**)&st.in_dev, st.in_pin, 0));
    EMS_CUDA(h, cudaHostGetDevicePointer((void**)&st.out_dev, st.out_pin, 0));
    ems_status s = stream_zero(h);
    if (s != EMS_OK) return s;
    if (N == 32768 && (s = ensure(h, h->big_scratch, (size_t)h->sm_count * r16::k32kScratch * sizeof(float2))) != EMS_OK) return s;
    st.ready = true;
    return EMS_OK;
}

// Records one push as a graph of three kernels: ring ingest (reads the hop from mapped pinned
// memory), fused STFT + reassignment + deposit of the frame the hop completes, and the finish
// kernel (post-pass of the column that became final, written to mapped pinned memory, AGC level
// and push counter advanced).
static ems_status stream_capture(ems_handle* h) {
    auto& st = h->st;
    const int H = h->prm.hop, C = h->prm.channels, B = rows_of(h->prm);
    const bool det = h->prm.flags & EMS_FLAG_DETERMINISTIC;
    StreamArgs sa{};
    sa.sstate = st.sstate; sa.in = st.in_dev; sa.ring = st.ring; sa.acc = st.acc;
    sa.carry = st.carry; sa.weight = h->weight; sa.out = st.out_dev;
    sa.hop = H; sa.channels = C; sa.M = st.M; sa.Lr = st.Lr; sa.R = st.R;
    sa.ring_cols = st.ring_cols; sa.B = B; sa.acc_is_u64 = det;
    sa.smoothing = h->prm.smoothing;
    sa.db_floor = (float)(kTopDb - (double)h->prm.db_range);
    sa.inv_range = 255.0f / h->prm.db_range;
    sa.gate_db = h->prm.noise_gate_db;
    sa.etmp = st.etmp; sa.agc = st.agc; sa.agc_strength = h->prm.agc_strength; sa.agc_target = agc_target(h->prm);
    sa.in_i16 = st.in_i16;
    if (h->stream_lut_on && !st.rgba_pin) {
        EMS_CUDA(h, cudaHostAlloc(&st.rgba_pin, (size_t)C * B * sizeof(uint32_t), cudaHostAllocMapped));
        EMS_CUDA(h, cudaHostGetDevicePointer((void**)&st.rgba_dev, st.rgba_pin, 0));
    }
    sa.lut = h->stream_lut_on ? (const uint32_t*)h->stream_lut.p : nullptr; sa.out_rgba = st.rgba_dev;
    sa.agc_lambda = std::exp(-(float)H / (h->prm.sample_rate * kAgcReleaseSeconds));
    StftArgs a = make_args(h, st.ring, (size_t)2 * st.Lr, /*F=*/(long long)1 << 60);
    a.f_begin = 0; a.f_end = 1;                       // grid sizing; the kernel decodes the real frame
    a.acc = st.acc; a.mode = det ? kDepositU64 : kDepositF32;
    a.ring = st.ring_cols; a.stream_M = st.M; a.sstate = st.sstate;

    // Recorded on the handle's own stream (a caller-provided stream may be the legacy default
    // stream, which cannot be captured); the graph is launched on whatever stream is current.
    cudaGraph_t g = nullptr;
    cudaStream_t user = h->stream;
    EMS_CUDA(h, cudaStreamSynchronize(user));
    h->stream = h->own_stream;
    cudaError_t be = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
    if (be != cudaSuccess) {
        h->stream = user;
        return fail(h, EMS_ERR_CUDA, "stream capture: %s", cudaGetErrorString(be));
    }
    stream_ingest_kernel<<<(H * C + 255) / 256, 256, 0, h->stream>>>(sa);
    ems_status ls = launch_stft(h, a);
    if (sa.agc_strength > 0.f)
        stream_finish_kernel<<<dim3(C, 1), 1024, 0, h->stream>>>(sa);          // channel peak = one block reduction
    else
        stream_finish_kernel<<<dim3(C, (B + 255) / 256), 256, 0, h->stream>>>(sa);
    cudaError_t ce = cudaStreamEndCapture(h->stream, &g);
    h->stream = user;
    h->launches -= 1;                                  // counted per push below, not at capture
    if (ls != EMS_OK) { if (g) cudaGraphDestroy(g); return ls; }
    if (ce != cudaSuccess) return fail(h, EMS_ERR_CUDA, "stream capture: %s", cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(&st.graph, g, 0);
    cudaGraphDestroy(g);
    if (ce != cudaSuccess) return fail(h, EMS_ERR_CUDA, "graph instantiate: %s", cudaGetErrorString(ce));
    return EMS_OK;
}

}  // namespace ems

using namespace ems;

// ============================================================================ C-ABI
extern "C" {

int ems_abi_version(void) { return EMS_ABI_VERSION; }

const char* ems_status_str(ems_status s) {
    switch (s) {
        case EMS_OK: return "ok";
        case EMS_ERR_INVALID_ARG: return "invalid argument";
        case EMS_ERR_UNSUPPORTED: return "unsupported";
        case EMS_ERR_CUDA: return "CUDA error";
        case EMS_ERR_NOMEM: return "out of memory";
        case EMS_ERR_STATE: return "invalid state";
    }
    return "unknown status";
}

const char* ems_last_error(const ems_handle* h) { return h ? h->err : "null handle"; }

ems_status ems_default_params(ems_params* p) {
    if (!p) return EMS_ERR_INVALID_ARG;
    p->n_fft = 4096; p->hop = 128; p->sample_rate = 48000.f; p->channels = 1;
    p->db_range = 58.f; p->gain = 3.5f; p->low_end_boost = 3.9f; p->smoothing = 0.f;
    p->noise_gate_db = -65.f;
    p->flags = EMS_FLAG_REASSIGN | EMS_FLAG_DETERMINISTIC;
    p->display_rows = 0; p->freq_scale = 1.0f; p->agc_strength = 0.0f; p->brightness = 0.44f;
    return EMS_OK;
}

ems_status ems_create(const ems_params* params, ems_handle** out) {
    if (!params || !out) return EMS_ERR_INVALID_ARG;
    *out = nullptr;
    if (!valid_params(*params)) return EMS_ERR_INVALID_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return EMS_ERR_CUDA;  // no CPU fallback
    ems_handle* h = new (std::nothrow) ems_handle();
    if (!h) return EMS_ERR_NOMEM;
    h->prm = *params;
    { const char* fg = getenv("EMS_FORCE_GENERIC"); h->force_generic = fg && fg[0] == '1'; }
    { const char* kv = getenv("EMS_KERNEL_VARIANT"); h->kernel_variant = kv ? atoi(kv) : 0; }
    { const char* mc = getenv("EMS_MAX_CTAS"); h->max_ctas = mc ? atoi(mc) : 0; }
    { const char* fv = getenv("EMS_FUSED_POST"); if (fv) h->fz.mode = atoi(fv); }
    { const char* fr = getenv("EMS_FUSED_RING"); if (fr && atoi(fr) >= 256) { int r = 256; while (r < atoi(fr)) r *= 2; h->fz.ring = r; } }
    auto bail = [&](ems_status s) { ems_destroy(h); return s; };
    if (cudaGetDevice(&h->device) != cudaSuccess) return bail(EMS_ERR_CUDA);
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, h->device) != cudaSuccess) return bail(EMS_ERR_CUDA);
    h->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking) != cudaSuccess)
        return bail(EMS_ERR_CUDA);
    h->stream = h->own_stream;
    for (auto& e : h->ev)
        for (auto& x : e)
            if (cudaEventCreate(&x) != cudaSuccess) return bail(EMS_ERR_CUDA);

    const int N = params->n_fft;
    if (cudaMalloc(&h->thw, sizeof(float) * N) != cudaSuccess ||
        cudaMalloc(&h->tw, sizeof(float2) * N) != cudaSuccess ||
        cudaMalloc(&h->weight, sizeof(float) * rows_of(*params)) != cudaSuccess)
        return bail(EMS_ERR_NOMEM);
    std::vector<float> thw(N);
    std::vector<float2> tw(N);
    for (int n = 0; n < N; ++n) {
        const double ang = 2.0 * kPi * (double)n / (double)N;
        const double c = std::cos(ang), s = std::sin(ang);
        const double hn = 0.5 - 0.5 * c;
        thw[n] = (float)(((double)n - N / 2) * hn * (2.0 / N));
        tw[n] = make_float2((float)c, (float)(-s));
    }
    if (cudaMemcpy(h->thw, thw.data(), sizeof(float) * N, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->tw, tw.data(), sizeof(float2) * N, cudaMemcpyHostToDevice) != cudaSuccess)
        return bail(EMS_ERR_CUDA);
    ems_status s = upload_display(h);
    if (s != EMS_OK) return bail(s);
    *out = h;
    return EMS_OK;
}

ems_status ems_destroy(ems_handle* h) {
    if (!h) return EMS_ERR_INVALID_ARG;
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (DevBuf* b : {&h->acc, &h->flags, &h->carry, &h->ema_local, &h->ema_carry, &h->lut, &h->stream_lut, &h->colscale, &h->agc_level,
                      &h->big_scratch, &h->post_mode, &h->sort_buf, &h->fz.ready, &h->fz.done, &h->hp.pcm[0], &h->hp.pcm[1], &h->hp.raw[0], &h->hp.raw[1], &h->hp.idx[0], &h->hp.idx[1],
                      &h->hp.grid[0], &h->hp.grid[1]})
        if (b->p) cudaFree(b->p);
    if (h->hp.events) {
        for (auto* evs : {h->hp.ev_in, h->hp.ev_done, h->hp.ev_out})
            for (int b = 0; b < 2; ++b) if (evs[b]) cudaEventDestroy(evs[b]);
        if (h->hp.ev_start) cudaEventDestroy(h->hp.ev_start);
    }
    stream_free(h);
    if (h->thw) cudaFree(h->thw);
    if (h->tw) cudaFree(h->tw);
    if (h->weight) cudaFree(h->weight);
    for (auto& e : h->ev)
        for (auto& x : e)
            if (x) cudaEventDestroy(x);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->copy_in) cudaStreamDestroy(h->copy_in);
    if (h->copy_out) cudaStreamDestroy(h->copy_out);
    delete h;
    return EMS_OK;
}

ems_status ems_update_display(ems_handle* h, const ems_params* p) {
    if (!h || !p) return EMS_ERR_INVALID_ARG;
    if (!valid_params(*p) || p->n_fft != h->prm.n_fft || p->hop != h->prm.hop ||
        p->channels != h->prm.channels || p->sample_rate != h->prm.sample_rate ||
        p->display_rows != h->prm.display_rows || p->freq_scale != h->prm.freq_scale)
        return fail(h, EMS_ERR_INVALID_ARG, "n_fft/hop/channels/sample_rate/display_rows/freq_scale need a new handle");
    const bool acc_type_changed = (p->flags ^ h->prm.flags) & EMS_FLAG_DETERMINISTIC;
    h->prm = *p;
    if (h->st.graph) { cudaGraphExecDestroy(h->st.graph); h->st.graph = nullptr; }   // scalars are baked into the graph
    if (acc_type_changed && h->st.ready) {   // the rolling accumulator changes element type: start over
        EMS_CUDA(h, cudaStreamSynchronize(h->stream));
        stream_free(h);
    }
    return upload_display(h);
}

ems_status ems_set_stream(ems_handle* h, void* s) {
    if (!h) return EMS_ERR_INVALID_ARG;
    // work queued on the old stream may still use scratch that the next call on the new stream
    // reallocates or overwrites: drain it first
    if (h->stream != (cudaStream_t)s) EMS_CUDA(h, cudaStreamSynchronize(h->stream));
    h->stream = (cudaStream_t)s;   // NULL is the CUDA default stream, taken literally
    return EMS_OK;
}

ems_status ems_get_stream(ems_handle* h, void** s) {
    if (!h || !s) return EMS_ERR_INVALID_ARG;
    *s = (void*)h->stream;
    return EMS_OK;
}

ems_status ems_synchronize(ems_handle* h) {
    if (!h) return EMS_ERR_INVALID_ARG;
    EMS_CUDA(h, cudaStreamSynchronize(h->stream));
    return EMS_OK;
}

ems_status ems_output_rows(const ems_handle* h, size_t* rows) {
    if (!h || !rows) return EMS_ERR_INVALID_ARG;
    *rows = (size_t)rows_of(h->prm);
    return EMS_OK;
}

ems_status ems_frame_count(const ems_handle* h, size_t S, size_t* F) {
    if (!h || !F) return EMS_ERR_INVALID_ARG;
    *F = (size_t)frames_of(h->prm, S);
    return EMS_OK;
}

// Frequency [Hz] of a (fractional) output row: the inverse of out_row (common.cuh) / upload_display's
// row centres.
static double row_to_hz(const ems_params& p, double row) {
    const int R = rows_of(p);
    row = std::min(std::max(row, 0.0), (double)(R - 1));
    if (p.display_rows <= 0) return row * (double)p.sample_rate / (double)p.n_fft;
    const double nyq = 0.5 * (double)p.sample_rate, a = warp_a_of(p);
    const double u = R > 1 ? row / (double)(R - 1) : 0.0;
    return nyq * (a > 1e-6 ? std::expm1(u * std::log1p(a)) / a : u);
}

ems_status ems_hz_to_row(const ems_handle* h, double freq_hz, double* row) {
    if (!h || !row || !(freq_hz == freq_hz)) return EMS_ERR_INVALID_ARG;
    const ems_params& p = h->prm;
    const int R = rows_of(p);
    double r;
    if (p.display_rows <= 0) {
        r = freq_hz * (double)p.n_fft / (double)p.sample_rate;
    } else {
        const double x = std::min(std::max(freq_hz / (0.5 * (double)p.sample_rate), 0.0), 1.0), a = warp_a_of(p);
        r = (double)(R - 1) * (a > 1e-6 ? std::log1p(a * x) / std::log1p(a) : x);
    }
    *row = std::min(std::max(r, 0.0), (double)(R - 1));
    return EMS_OK;
}

// "note and frequency information" under the cursor (/root/reference/README.md:39).
ems_status ems_cursor_info(const ems_handle* h, double column, double row, ems_cursor* out) {
    if (!h || !out || !(column == column) || !(row == row)) return EMS_ERR_INVALID_ARG;
    static const char* const kNames[12] = {"C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"};
    const ems_params& p = h->prm;
    std::memset(out, 0, sizeof(*out));
    out->time_s = (column * (double)p.hop + 0.5 * (double)p.n_fft) / (double)p.sample_rate;
    out->freq_hz = row_to_hz(p, row);
    out->midi_note = -1;
    if (out->freq_hz >= 1.0) {
        const double pitch = 69.0 + 12.0 * std::log2(out->freq_hz / 440.0);
        const double note = std::floor(pitch + 0.5);
        out->midi_note = (int32_t)note;
        out->cents = (float)(100.0 * (pitch - note));
        const int pc = ((out->midi_note % 12) + 12) % 12;
        const int octave = (out->midi_note - pc) / 12 - 1;
        std::snprintf(out->name, sizeof(out->name), "%s%d", kNames[pc], octave);
    }
    return EMS_OK;
}

// Built-in colour maps ("Multiple Color Maps", /root/reference/README.md:15,45): stand-in tables,
// piecewise linear through 8-bit control colours in integer arithmetic (oracle: builtin_colormap).
namespace {
struct ColourStop { int pos; int r, g, b; };
struct ColourMap { const char* name; int n; ColourStop stop[5]; };
const ColourMap kColourMaps[] = {
    {"inferno", 5, {{0, 0, 0, 4}, {64, 87, 16, 110}, {128, 188, 55, 84}, {192, 249, 142, 9}, {255, 252, 255, 164}}},   // settings.png "Default" preset
    {"gray",    2, {{0, 0, 0, 0}, {255, 255, 255, 255}}},
    {"heat",    4, {{0, 0, 0, 0}, {85, 200, 0, 0}, {170, 255, 200, 0}, {255, 255, 255, 255}}},
    {"magma",   5, {{0, 0, 0, 4}, {64, 81, 18, 124}, {128, 183, 55, 121}, {192, 252, 137, 97}, {255, 252, 253, 191}}},
    {"viridis", 5, {{0, 68, 1, 84}, {64, 59, 82, 139}, {128, 33, 145, 140}, {192, 94, 201, 98}, {255, 253, 231, 37}}},
    {"ice",     4, {{0, 0, 0, 0}, {96, 0, 60, 160}, {192, 80, 200, 255}, {255, 255, 255, 255}}},
};
constexpr int kColourMapCount = (int)(sizeof(kColourMaps) / sizeof(kColourMaps[0]));
}  // namespace

int ems_colormap_count(void) { return kColourMapCount; }
const char* ems_colormap_name(int id) { return id >= 0 && id < kColourMapCount ? kColourMaps[id].name : nullptr; }

ems_status ems_colormap_builtin(int id, uint32_t lut[256]) {
    if (!lut || id < 0 || id >= kColourMapCount) return EMS_ERR_INVALID_ARG;
    const ColourMap& m = kColourMaps[id];
    for (int s = 0; s + 1 < m.n; ++s) {
        const ColourStop &a = m.stop[s], &b = m.stop[s + 1];
        const int d = b.pos - a.pos;
        for (int i = a.pos; i <= b.pos; ++i) {
            const int t = i - a.pos;
            const uint32_t r = (uint32_t)((a.r * (d - t) + b.r * t + d / 2) / d);
            const uint32_t g = (uint32_t)((a.g * (d - t) + b.g * t + d / 2) / d);
            const uint32_t bl = (uint32_t)((a.b * (d - t) + b.b * t + d / 2) / d);
            lut[i] = 0xFF000000u | (bl << 16) | (g << 8) | r;
        }
    }
    return EMS_OK;
}

ems_status ems_process_points(ems_handle* h, const float* pcm, size_t S, float* dt_cols,
                              float* dk_bins, float* energy, size_t* n_frames) {
    if (!h) return EMS_ERR_INVALID_ARG;
    const long long F = frames_of(h->prm, S);
    if (n_frames) *n_frames = (size_t)F;
    for (bool& v : h->ev_valid) v = false;
    if (F == 0) return EMS_OK;
    if (!pcm || !dt_cols || !dk_bins || !energy) return fail(h, EMS_ERR_INVALID_ARG, "null buffer");
    StftArgs a = make_args(h, pcm, S, F);
    a.dt_cols = dt_cols; a.dk_bins = dk_bins; a.energy = energy; a.mode = kStorePoints;
    stage_begin(h, EMS_STAGE_POINTS);
    ems_status s = launch_stft(h, a);
    if (s != EMS_OK) { stage_abort(); return s; }
    stage_end(h, EMS_STAGE_POINTS);
    return finish(h);
}

// a4 as sort-by-cell + segmented reduce (scatter_sorted.cuh), in chunks of at most 2^25 points so that
// the scratch (two key and two value buffers, 24 bytes per point) stays below 1 GB.
static ems_status scatter_sorted(ems_handle* h, const float* dt_cols, const float* dk_bins, const float* energy,
                                 size_t points, size_t cells, long long F, int B, int R, const StftArgs& wa) {
    namespace so = ems::sorted;
    size_t chunk_max = (size_t)1 << 25;
    if (const char* ev = getenv("EMS_SORT_CHUNK")) { const long long v = atoll(ev); if (v > 0) chunk_max = std::min(chunk_max, (size_t)v); }   // tests: chunk seams on small inputs
    const size_t chunk = std::min(points, chunk_max);
    const int G = h->sm_count * 4;
    const size_t table = ((size_t)so::kRadix * (G + 1) + 4) * sizeof(unsigned);   // digit-major counts + 256 digit totals + kept count
    ems_status s = ensure(h, h->sort_buf, chunk * 24 + table);
    if (s != EMS_OK) return s;
    unsigned long long* keys[2] = {(unsigned long long*)h->sort_buf.p, (unsigned long long*)h->sort_buf.p + chunk};
    float* vals[2] = {(float*)(keys[1] + chunk), (float*)(keys[1] + chunk) + chunk};
    unsigned* counts = (unsigned*)(vals[1] + chunk);
    unsigned* totals = counts + (size_t)so::kRadix * G;
    unsigned* nkept = totals + so::kRadix;
    int bits = 0;
    while (bits < 64 && ((unsigned long long)cells >> bits) != 0ull) ++bits;     // the sentinel key is `cells` itself
    const int passes = (bits + 7) / 8;
    const int kb = (int)std::min<size_t>((chunk + 255) / 256, (size_t)h->sm_count * 16);
    for (size_t i0 = 0; i0 < points; i0 += chunk) {
        const long long n = (long long)std::min(chunk, points - i0);
        so::keys_kernel<<<kb, 256, 0, h->stream>>>(dt_cols, dk_bins, energy, (long long)i0, n, keys[0], vals[0], F, B, R,
                                                   (unsigned long long)cells, wa.warp_mode, wa.warp_a, wa.warp_c, wa.inv_half);
        // pass -1: stable partition kept | dropped over all n points; passes 0..: radix passes over the kept ones
        int cur = 0;
        for (int p = -1; p < passes; ++p, cur ^= 1) {
            const unsigned* n_dev = p < 0 ? nullptr : nkept;
            const unsigned long long part = p < 0 ? (unsigned long long)cells : 0ull;
            const int shift = p < 0 ? 0 : 8 * p;
            so::histogram_kernel<<<G, so::kThreads, 0, h->stream>>>(keys[cur], n, n_dev, shift, part, counts);
            so::scan_kernel<<<so::kRadix, 256, 0, h->stream>>>(counts, totals, G, p < 0 ? nkept : nullptr);
            so::scatter_kernel<<<G, so::kThreads, 0, h->stream>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, n_dev,
                                                                    shift, part, counts, totals);
            h->launches += 3;
        }
        so::reduce_kernel<<<kb, 256, 0, h->stream>>>(keys[cur], vals[cur], nkept, (unsigned long long)cells, (float*)h->acc.p,
                                                     (unsigned char*)h->flags.p, F, R);
        h->launches += 2;
        EMS_CUDA(h, cudaGetLastError());
    }
    return EMS_OK;
}

ems_status ems_scatter_points(ems_handle* h, const float* dt_cols, const float* dk_bins,
                              const float* energy, size_t n_frames, float* grid, uint8_t* index) {
    if (!h) return EMS_ERR_INVALID_ARG;
    for (bool& v : h->ev_valid) v = false;
    if (n_frames == 0) return EMS_OK;
    if (!dt_cols || !dk_bins || !energy || (!grid && !index))
        return fail(h, EMS_ERR_INVALID_ARG, "null buffer");
    const long long F = (long long)n_frames;
    const int B = h->prm.n_fft / 2 + 1, C = h->prm.channels, R = rows_of(h->prm);
    const bool det = h->prm.flags & EMS_FLAG_DETERMINISTIC;
    const size_t cells = (size_t)C * F * R, points = (size_t)C * F * B;
    ems_status s = prepare_acc(h, cells, (size_t)C * F, R);
    if (s != EMS_OK) return s;
    if ((s = reset_carry(h)) != EMS_OK) return s;
    stage_begin(h, EMS_STAGE_SCATTER);
    long long blocks = (long long)((points + 255) / 256);
    const long long cap = (long long)h->sm_count * 32;
    if (blocks > cap) blocks = cap;
    const StftArgs wa = make_args(h, nullptr, 0, F);   // carries the frequency-axis warp
    const bool sorted = h->prm.flags & EMS_FLAG_SORTED_SCATTER;
    if (sorted) {
        if ((s = scatter_sorted(h, dt_cols, dk_bins, energy, points, cells, F, B, R, wa)) != EMS_OK) { stage_abort(); return s; }
    } else {
        scatter_points_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(
            dt_cols, dk_bins, energy, h->acc.p, det, (unsigned char*)h->flags.p, F, B, C, R,
            wa.warp_mode, wa.warp_a, wa.warp_c, wa.inv_half);
        ++h->launches;
    }
    if (cudaError_t le = cudaGetLastError(); le != cudaSuccess) {
        stage_abort();
        return fail(h, EMS_ERR_CUDA, "scatter_points_kernel: %s", cudaGetErrorString(le));
    }
    stage_end(h, EMS_STAGE_SCATTER);
    stage_begin(h, EMS_STAGE_POST);
    PostArgs pa = make_post(h, F, grid, index);
    if (sorted) pa.acc_is_u64 = 0;                      // the sorted scatter sums in fp32
    if ((s = run_post(h, pa)) != EMS_OK) { stage_abort(); return s; }
    h->acc_clean = true;
    stage_end(h, EMS_STAGE_POST);
    return finish(h);
}

ems_status ems_process_grid(ems_handle* h, const float* pcm, size_t S, float* grid,
                            uint8_t* index, size_t* n_frames) {
    if (!h) return EMS_ERR_INVALID_ARG;
    const long long F = frames_of(h->prm, S);
    if (n_frames) *n_frames = (size_t)F;
    for (bool& v : h->ev_valid) v = false;
    if (F == 0) return EMS_OK;
    if (!pcm || (!grid && !index)) return fail(h, EMS_ERR_INVALID_ARG, "null buffer");
    const int C = h->prm.channels, R = rows_of(h->prm);
    const bool det = h->prm.flags & EMS_FLAG_DETERMINISTIC;
    // Without smoothing and AGC a column depends on nothing but its own cells: the deposit kernel shapes
    // finished column blocks itself, out of an accumulator ring that stays in L2 (tiled kernels only).
    const bool fused = EMS_FUSED_POST && h->fz.mode && h->prm.smoothing == 0.f && h->prm.agc_strength == 0.f && !h->force_generic &&
                       h->kernel_variant == 0 && h->prm.n_fft <= 4096 && flag_blocks(R) <= 64 &&
                       (size_t)C * (size_t)F > (size_t)0;
    if (fused) {
        const size_t ring = (size_t)h->fz.ring;
        ems_status s = prepare_acc(h, ring * R, ring, R);
        if (s != EMS_OK) return s;
        StftArgs a = make_args(h, pcm, S, F);
        a.acc = h->acc.p; a.flags = (unsigned char*)h->flags.p; a.mode = det ? kDepositU64 : kDepositF32;
        a.fp.vring = (int)ring; a.fp.NB = flag_blocks(R);
        a.fp.index = index; a.fp.grid = grid; a.fp.weight = h->weight;
        a.fp.db_floor = (float)(kTopDb - (double)h->prm.db_range);
        a.fp.inv_range = 255.0f / h->prm.db_range;
        a.fp.gate_db = h->prm.noise_gate_db;
        { const char* dv = getenv("EMS_FUSED_DEBUG"); a.fp.debug = dv ? atoi(dv) : 0; }
        stage_begin(h, EMS_STAGE_POINTS);
        s = launch_stft(h, a);
        if (s == EMS_ERR_STATE) {                      // ring too short for this geometry: the two-kernel path below
            stage_abort();
            h->acc_clean = true;                       // nothing was launched
        } else {
            if (s != EMS_OK) { stage_abort(); return s; }
            stage_end(h, EMS_STAGE_POINTS);
            // the post-pass ran inside the kernel: report it as a zero-length stage
            stage_begin(h, EMS_STAGE_POST);
            stage_end(h, EMS_STAGE_POST);
            h->acc_clean = true;
            h->fz.counters_clean = true;
            return finish(h);
        }
    }
    if (h->prm.flags & EMS_FLAG_BOUNDED_SCRATCH) {
        // The same result in O(chunk) device memory: frame chunks deposit into an accumulator ring of
        // 2^n >= chunk + 2R columns and the post-pass shapes, into the caller's buffers, the columns no
        // later frame can reach.  A chunk is a whole number of rounds of the persistent grid
        // (sm_count x 36 frames), so the per-chunk tail is small: about 2-6 % slower than one launch.
        const int N = h->prm.n_fft, H = h->prm.hop;
        const long long Rc = (N / 2 + H - 1) / H;
        long long chunk = 24LL * h->sm_count * 36;
        if (chunk * C > (1LL << 20)) chunk = std::max<long long>((1LL << 20) / C, 4 * Rc + 1024);
        if (chunk > F) chunk = F;
        long long ring = 1;
        while (ring < chunk + 2 * Rc + 2) ring *= 2;
        ems_status s = prepare_acc(h, (size_t)C * ring * R, (size_t)C * ring, R);
        if (s != EMS_OK) return s;
        if ((s = reset_carry(h)) != EMS_OK) return s;
        stage_begin(h, EMS_STAGE_POINTS);      // (the stages interleave: both brackets span the whole call)
        long long cols_done = 0;
        for (long long f0 = 0; f0 < F; f0 += chunk) {
            const long long f1 = std::min(F, f0 + chunk);
            StftArgs a = make_args(h, pcm, S, F);
            a.f_begin = f0; a.f_end = f1;
            a.acc = h->acc.p; a.flags = (unsigned char*)h->flags.p; a.mode = det ? kDepositU64 : kDepositF32;
            a.ring = (int)ring;
            if ((s = launch_stft(h, a)) != EMS_OK) { stage_abort(); return s; }
            const long long col_end = (f1 == F) ? F : std::max(cols_done, f1 - Rc);
            PostArgs p = make_post(h, F, grid, index);
            p.col_begin = cols_done; p.col_end = col_end;
            p.acc_cols = ring; p.acc_mask = ring - 1;
            if ((s = run_post(h, p)) != EMS_OK) { stage_abort(); return s; }
            cols_done = col_end;
        }
        stage_end(h, EMS_STAGE_POINTS);
        h->acc_clean = true;
        return finish(h);
    }
    const size_t cells = (size_t)C * F * R;
    ems_status s = prepare_acc(h, cells, (size_t)C * F, R);
    if (s != EMS_OK) return s;
    if ((s = reset_carry(h)) != EMS_OK) return s;
    StftArgs a = make_args(h, pcm, S, F);
    a.acc = h->acc.p; a.flags = (unsigned char*)h->flags.p; a.mode = det ? kDepositU64 : kDepositF32;
    stage_begin(h, EMS_STAGE_POINTS);
    if ((s = launch_stft(h, a)) != EMS_OK) { stage_abort(); return s; }
    stage_end(h, EMS_STAGE_POINTS);
    stage_begin(h, EMS_STAGE_POST);
    if ((s = run_post(h, make_post(h, F, grid, index))) != EMS_OK) { stage_abort(); return s; }
    h->acc_clean = true;
    stage_end(h, EMS_STAGE_POST);
    return finish(h);
}

// ---- ems_process_host*: the HOST-buffer call, O(chunk) device memory.
// The stream is cut into frame chunks.  Per chunk: its samples ((n-1) hop + n_fft per channel, the
// n_fft - hop halo is uploaded again) go into one of two device PCM buffers, the fused kernel
// deposits into an accumulator RING of 2^n >= chunk + 2R columns, the post-pass shapes the columns
// no later frame can reach (col < f_end - R) into one of two staging images, and those go back to
// the caller while the next chunk runs.  Nothing on the device grows with the stream length
// except the AGC's per-column scale (4 bytes per column).
enum HostFmt : int { kFmtF32Planar = 0, kFmtI16 = 1, kFmtI24 = 2 };
static size_t fmt_bytes(int fmt) { return fmt == kFmtI16 ? 2 : fmt == kFmtI24 ? 3 : 4; }

static ems_status process_host_impl(ems_handle* h, const void* pcm_host_v, int fmt, size_t S,
                                    float* grid_host, uint8_t* index_host, size_t* n_frames) {
    if (!h) return EMS_ERR_INVALID_ARG;
    const long long F = frames_of(h->prm, S);
    if (n_frames) *n_frames = (size_t)F;
    for (bool& v : h->ev_valid) v = false;
    if (F == 0) return EMS_OK;
    if (!pcm_host_v || (!grid_host && !index_host)) return fail(h, EMS_ERR_INVALID_ARG, "null buffer");
    const int N = h->prm.n_fft, H = h->prm.hop, B = rows_of(h->prm), C = h->prm.channels;   // B: output rows
    const bool det = h->prm.flags & EMS_FLAG_DETERMINISTIC;
    const long long R = (N / 2 + H - 1) / H;
    // Chunk = about 64 MiB of output image: the pipeline's tail is the D2H of the last chunk
    // (~1.2 ms at PCIe Gen5 rates) whatever the row count, and launches stay large.
    long long chunk = (long long)(((size_t)64 << 20) / ((size_t)B * C));
    if (const char* ev = getenv("EMS_HOST_CHUNK_FRAMES")) {   // test hook: force small chunks
        const long long v = atoll(ev);
        if (v > 0) chunk = v;
    }
    if (chunk > 262144) chunk = 262144;
    if (chunk < 4 * R + 1024) chunk = 4 * R + 1024;
    if (chunk > F) chunk = F;
    const int n_chunks = (int)((F + chunk - 1) / chunk);
    long long ring = 1;
    while (ring < chunk + 2 * R + 2) ring *= 2;             // columns [cols_done, f_end + R) are live
    const long long Lc = (chunk - 1) * H + N;               // samples per channel of one chunk
    const long long out_cols = std::min(F, chunk + R);      // columns one chunk can finish
    const bool is_int = fmt != kFmtF32Planar;
    const size_t bps = fmt_bytes(fmt);

    ems_status s;
    auto& hp = h->hp;
    if (!hp.events) {
        for (auto* evs : {hp.ev_in, hp.ev_done, hp.ev_out})
            for (int b = 0; b < 2; ++b) EMS_CUDA(h, cudaEventCreateWithFlags(&evs[b], cudaEventDisableTiming));
        EMS_CUDA(h, cudaEventCreateWithFlags(&hp.ev_start, cudaEventDisableTiming));
        hp.events = true;
    }
    for (int b = 0; b < 2; ++b) {
        if ((s = ensure(h, hp.pcm[b], (size_t)C * Lc * sizeof(float))) != EMS_OK) return s;
        if (is_int && (s = ensure(h, hp.raw[b], (size_t)C * Lc * bps + 16)) != EMS_OK) return s;
        if (index_host && (s = ensure(h, hp.idx[b], (size_t)C * out_cols * B)) != EMS_OK) return s;
        if (grid_host && (s = ensure(h, hp.grid[b], (size_t)C * out_cols * B * sizeof(float))) != EMS_OK) return s;
    }
    if ((s = prepare_acc(h, (size_t)C * ring * B, (size_t)C * ring, B)) != EMS_OK) return s;
    if ((s = reset_carry(h)) != EMS_OK) return s;

    int max_pitch = 0;
    EMS_CUDA(h, cudaDeviceGetAttribute(&max_pitch, cudaDevAttrMaxPitch, h->device));
    // rows of a [channels][...] array: one pitched copy, or one copy per channel when the caller's
    // row pitch is beyond what a pitched copy takes (an hour of image is 2.8 GB per channel)
    auto copy_rows = [&](void* dst, size_t dpitch, const void* src, size_t spitch, size_t width,
                         cudaMemcpyKind kind, cudaStream_t st) -> cudaError_t {
        if (C == 1) return cudaMemcpyAsync(dst, src, width, kind, st);
        if (dpitch <= (size_t)max_pitch && spitch <= (size_t)max_pitch)
            return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, (size_t)C, kind, st);
        for (int ch = 0; ch < C; ++ch) {
            cudaError_t e = cudaMemcpyAsync((char*)dst + ch * dpitch, (const char*)src + ch * spitch, width, kind, st);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };

    ems_status rs = EMS_OK;
    cudaError_t ce = cudaSuccess;
    const char* what = "";
#define HP_CUDA(call)                                                        \
    if ((ce = (call)) != cudaSuccess) { what = #call; rs = EMS_ERR_CUDA; break; }

    long long cols_done = 0;   // columns already post-processed
    do {
        HP_CUDA(cudaEventRecord(hp.ev_start, h->stream));
        HP_CUDA(cudaStreamWaitEvent(h->copy_in, hp.ev_start, 0));      // the memsets and earlier work come first
        HP_CUDA(cudaStreamWaitEvent(h->copy_out, hp.ev_start, 0));
    } while (0);
    for (int c = 0; c < n_chunks && rs == EMS_OK; ++c) {
        const int b = c & 1;
        const long long f0 = (long long)c * chunk, f1 = std::min(F, f0 + chunk);
        const long long s0 = f0 * H, n = (f1 - 1) * H + N - s0;          // samples [s0, s0 + n) of every channel
        float* pcm_dev = (float*)hp.pcm[b].p;
        if (c >= 2) HP_CUDA(cudaStreamWaitEvent(h->copy_in, hp.ev_done[b], 0));   // chunk c-2 has read this buffer
        if (!is_int) {
            HP_CUDA(copy_rows(pcm_dev, (size_t)Lc * sizeof(float), (const float*)pcm_host_v + s0, S * sizeof(float),
                              (size_t)n * sizeof(float), cudaMemcpyHostToDevice, h->copy_in));
        } else {      // interleaved: one contiguous run of n * C samples
            HP_CUDA(cudaMemcpyAsync(hp.raw[b].p, (const char*)pcm_host_v + (size_t)s0 * C * bps, (size_t)n * C * bps,
                                    cudaMemcpyHostToDevice, h->copy_in));
        }
        HP_CUDA(cudaEventRecord(hp.ev_in[b], h->copy_in));
        HP_CUDA(cudaStreamWaitEvent(h->stream, hp.ev_in[b], 0));
        if (is_int) {
            const long long work = fmt == kFmtI24 ? (n * C + 3) / 4 : n * C;
            long long blocks = (work + 255) / 256;
            if (blocks > 8192) blocks = 8192;
            if (fmt == kFmtI16)
                pcm_i16_to_planar_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>((const int16_t*)hp.raw[b].p, pcm_dev, Lc, C, n);
            else
                pcm_i24_to_planar_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>((const uint32_t*)hp.raw[b].p, pcm_dev, Lc, C, n);
            ++h->launches;
            HP_CUDA(cudaGetLastError());
        }
        StftArgs a = make_args(h, pcm_dev, (size_t)Lc, F);
        a.samp_off = -s0;                                                 // frame f starts at sample f hop - s0 of the buffer
        a.f_begin = f0; a.f_end = f1;
        a.acc = h->acc.p; a.flags = (unsigned char*)h->flags.p; a.mode = det ? kDepositU64 : kDepositF32;
        a.ring = (int)ring;
        if ((rs = launch_stft(h, a)) != EMS_OK) break;
        const long long col_end = (f1 == F) ? F : std::max(cols_done, f1 - R);
        if (c >= 2) HP_CUDA(cudaStreamWaitEvent(h->stream, hp.ev_out[b], 0));    // staging image b has gone to the host
        PostArgs p = make_post(h, F, grid_host ? (float*)hp.grid[b].p : nullptr, index_host ? (uint8_t*)hp.idx[b].p : nullptr);
        p.col_begin = cols_done; p.col_end = col_end;
        p.acc_cols = ring; p.acc_mask = ring - 1; p.out_cols = out_cols; p.out_col0 = cols_done;
        if ((rs = run_post(h, p)) != EMS_OK) break;
        HP_CUDA(cudaEventRecord(hp.ev_done[b], h->stream));
        HP_CUDA(cudaStreamWaitEvent(h->copy_out, hp.ev_done[b], 0));
        if (col_end > cols_done) {
            const size_t nc = (size_t)(col_end - cols_done);
            if (index_host)
                HP_CUDA(copy_rows(index_host + (size_t)cols_done * B, (size_t)F * B, hp.idx[b].p, (size_t)out_cols * B,
                                  nc * B, cudaMemcpyDeviceToHost, h->copy_out));
            if (grid_host)
                HP_CUDA(copy_rows(grid_host + (size_t)cols_done * B, (size_t)F * B * sizeof(float), hp.grid[b].p,
                                  (size_t)out_cols * B * sizeof(float), nc * B * sizeof(float), cudaMemcpyDeviceToHost,
                                  h->copy_out));
        }
        HP_CUDA(cudaEventRecord(hp.ev_out[b], h->copy_out));
        cols_done = col_end;
    }
#undef HP_CUDA
    // whatever happened, no copy may still be targeting the caller's buffers when this returns
    const cudaError_t e1 = cudaStreamSynchronize(h->copy_out);
    const cudaError_t e2 = cudaStreamSynchronize(h->stream);
    const cudaError_t e3 = cudaStreamSynchronize(h->copy_in);
    if (rs == EMS_ERR_CUDA && ce != cudaSuccess) return fail(h, EMS_ERR_CUDA, "process_host: %s: %s", what, cudaGetErrorString(ce));
    if (rs != EMS_OK) return rs;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return fail(h, EMS_ERR_CUDA, "process_host: %s",
                    cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
    h->acc_clean = true;
    return EMS_OK;
}

ems_status ems_process_host(ems_handle* h, const float* pcm_host, size_t S, float* grid_host,
                            uint8_t* index_host, size_t* n_frames) {
    return process_host_impl(h, pcm_host, kFmtF32Planar, S, grid_host, index_host, n_frames);
}

ems_status ems_process_host_i16(ems_handle* h, const int16_t* pcm_host, size_t S, float* grid_host,
                                uint8_t* index_host, size_t* n_frames) {
    return process_host_impl(h, pcm_host, kFmtI16, S, grid_host, index_host, n_frames);
}

ems_status ems_process_host_i24(ems_handle* h, const uint8_t* pcm_host, size_t S, float* grid_host,
                                uint8_t* index_host, size_t* n_frames) {
    return process_host_impl(h, pcm_host, kFmtI24, S, grid_host, index_host, n_frames);
}

ems_status ems_scratch_bytes(const ems_handle* h, size_t* bytes) {
    if (!h || !bytes) return EMS_ERR_INVALID_ARG;
    size_t n = 0;
    for (const DevBuf* b : {&h->acc, &h->flags, &h->carry, &h->ema_local, &h->ema_carry, &h->big_scratch, &h->sort_buf, &h->lut, &h->stream_lut,
                            &h->colscale, &h->agc_level, &h->hp.pcm[0], &h->hp.pcm[1], &h->hp.raw[0], &h->hp.raw[1],
                            &h->hp.idx[0], &h->hp.idx[1], &h->hp.grid[0], &h->hp.grid[1]})
        n += b->bytes;
    *bytes = n;
    return EMS_OK;
}

ems_status ems_colorize(ems_handle* h, const uint8_t* index_dev, size_t n_cells,
                        const uint32_t* lut_host, uint32_t* rgba_dev) {
    if (!h) return EMS_ERR_INVALID_ARG;
    if (n_cells == 0) return EMS_OK;
    if (!index_dev || !lut_host || !rgba_dev) return fail(h, EMS_ERR_INVALID_ARG, "null buffer");
    ems_status s = ensure(h, h->lut, 256 * sizeof(uint32_t));
    if (s != EMS_OK) return s;
    EMS_CUDA(h, cudaMemcpyAsync(h->lut.p, lut_host, 256 * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    size_t blocks = (n_cells / 16 + 255) / 256 + 1;
    const size_t cap = (size_t)h->sm_count * 16;
    if (blocks > cap) blocks = cap;
    colorize_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(index_dev, rgba_dev, n_cells, (const uint32_t*)h->lut.p);
    ++h->launches;
    EMS_CUDA(h, cudaGetLastError());
    return finish(h);
}

ems_status ems_image_summary(ems_handle* h, const uint8_t* index_dev, size_t n_frames, uint64_t* summary_dev) {
    if (!h) return EMS_ERR_INVALID_ARG;
    if (!index_dev || !summary_dev) return fail(h, EMS_ERR_INVALID_ARG, "null buffer");
    const int C = h->prm.channels;
    EMS_CUDA(h, cudaMemsetAsync(summary_dev, 0, (size_t)C * 2 * sizeof(uint64_t), h->stream));
    const size_t n = n_frames * (size_t)rows_of(h->prm);
    if (n == 0) return finish(h);
    unsigned bx = (unsigned)std::min<size_t>((n / 16 + 255) / 256 + 1, (size_t)std::max(1, h->sm_count * 8 / C));
    if (bx < 1) bx = 1;
    image_summary_kernel<<<dim3(bx, C), 256, 0, h->stream>>>(index_dev, n, (unsigned long long*)summary_dev);
    ++h->launches;
    EMS_CUDA(h, cudaGetLastError());
    return finish(h);
}

ems_status ems_stage_ms(ems_handle* h, int stage, float* ms) {
    if (!h || !ms || stage < 0 || stage >= EMS_STAGE_COUNT) return EMS_ERR_INVALID_ARG;
    if (!h->ev_valid[stage]) return fail(h, EMS_ERR_STATE, "stage %d did not run", stage);
    EMS_CUDA(h, cudaEventSynchronize(h->ev[stage][1]));
    EMS_CUDA(h, cudaEventElapsedTime(ms, h->ev[stage][0], h->ev[stage][1]));
    return EMS_OK;
}

ems_status ems_launch_count(const ems_handle* h, uint64_t* n) {
    if (!h || !n) return EMS_ERR_INVALID_ARG;
    *n = h->launches;
    return EMS_OK;
}

static ems_status stream_push_impl(ems_handle* h, const void* pcm_host, int is_i16, uint8_t* column_host,
                                   int* column_ready, int64_t* column_index) {
    if (!h || !pcm_host || !column_host || !column_ready) return EMS_ERR_INVALID_ARG;
    auto& st = h->st;
    ems_status s;
    if (!st.ready && (s = stream_init(h)) != EMS_OK) return s;
    if (st.graph && st.in_i16 != is_i16) {          // the ingest kernel of the graph reads the other format
        EMS_CUDA(h, cudaStreamSynchronize(h->stream));
        cudaGraphExecDestroy(st.graph);
        st.graph = nullptr;
    }
    st.in_i16 = is_i16;
    if (!st.graph && (s = stream_capture(h)) != EMS_OK) return s;
    const int H = h->prm.hop, C = h->prm.channels, B = rows_of(h->prm);
    memcpy(st.in_pin, pcm_host, (is_i16 == 2 ? 3 : is_i16 ? sizeof(int16_t) : sizeof(float)) * (size_t)H * C);
    EMS_CUDA(h, cudaGraphLaunch(st.graph, h->stream));
    h->launches += 3;
    EMS_CUDA(h, cudaStreamSynchronize(h->stream));
    const long long cf = st.pushes + 1 - st.M - st.R;   // column finalised by this push
    ++st.pushes;
    *column_ready = cf >= 0;
    st.last_ready = cf >= 0;
    if (column_index) *column_index = cf;
    if (cf >= 0) memcpy(column_host, st.out_pin, (size_t)C * B);
    return EMS_OK;
}

ems_status ems_stream_push_i24(ems_handle* h, const uint8_t* pcm_host, uint8_t* column_host,
                               int* column_ready, int64_t* column_index) {
    return stream_push_impl(h, pcm_host, 2, column_host, column_ready, column_index);
}

ems_status ems_stream_set_colormap(ems_handle* h, const uint32_t* lut_rgba_host) {
    if (!h) return EMS_ERR_INVALID_ARG;
    auto& st = h->st;
    EMS_CUDA(h, cudaStreamSynchronize(h->stream));
    if (lut_rgba_host) {
        ems_status s = ensure(h, h->stream_lut, 256 * sizeof(uint32_t));
        if (s != EMS_OK) return s;
        EMS_CUDA(h, cudaMemcpy(h->stream_lut.p, lut_rgba_host, 256 * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    const bool on = lut_rgba_host != nullptr;
    if (on != h->stream_lut_on && st.graph) {       // the finish kernel of the graph has the table pointer baked in
        cudaGraphExecDestroy(st.graph);
        st.graph = nullptr;
    }
    h->stream_lut_on = on;
    st.last_ready = false;
    return EMS_OK;
}

ems_status ems_stream_column_rgba(ems_handle* h, uint32_t* rgba_host) {
    if (!h || !rgba_host) return EMS_ERR_INVALID_ARG;
    auto& st = h->st;
    if (!h->stream_lut_on) return fail(h, EMS_ERR_STATE, "no colour map set (ems_stream_set_colormap)");
    if (!st.ready || !st.last_ready || !st.rgba_pin) return fail(h, EMS_ERR_STATE, "the last push delivered no column");
    memcpy(rgba_host, st.rgba_pin, (size_t)h->prm.channels * rows_of(h->prm) * sizeof(uint32_t));
    return EMS_OK;
}

ems_status ems_stream_push(ems_handle* h, const float* pcm_host, uint8_t* column_host,
                           int* column_ready, int64_t* column_index) {
    return stream_push_impl(h, pcm_host, 0, column_host, column_ready, column_index);
}

ems_status ems_stream_push_i16(ems_handle* h, const int16_t* pcm_host, uint8_t* column_host,
                               int* column_ready, int64_t* column_index) {
    return stream_push_impl(h, pcm_host, 1, column_host, column_ready, column_index);
}

// ---- stream checkpoint: header + [pushes][ring][acc][carry][agc]
namespace {
struct StreamBlobHeader {
    uint32_t magic, abi;
    int32_t n_fft, hop, channels, rows, det;
    int64_t pushes;
    uint64_t ring_bytes, acc_bytes, carry_bytes, agc_bytes;
};
constexpr uint32_t kBlobMagic = 0x53534d45u;   // "EMSS"

StreamBlobHeader blob_header(const ems_handle* h) {
    const auto& st = h->st;
    StreamBlobHeader b{};
    b.magic = kBlobMagic; b.abi = EMS_ABI_VERSION;
    b.n_fft = h->prm.n_fft; b.hop = h->prm.hop; b.channels = h->prm.channels;
    b.rows = rows_of(h->prm); b.det = (h->prm.flags & EMS_FLAG_DETERMINISTIC) ? 1 : 0;
    b.pushes = st.pushes;
    b.ring_bytes = sizeof(float) * (size_t)b.channels * 2 * st.Lr;
    b.acc_bytes = st.acc_bytes;
    b.carry_bytes = sizeof(float) * (size_t)b.channels * b.rows;
    b.agc_bytes = sizeof(float) * 2 * (size_t)b.channels;
    return b;
}
size_t blob_size(const StreamBlobHeader& b) {
    return sizeof(b) + b.ring_bytes + b.acc_bytes + b.carry_bytes + b.agc_bytes;
}
}  // namespace

ems_status ems_stream_state_size(ems_handle* h, size_t* bytes) {
    if (!h || !bytes) return EMS_ERR_INVALID_ARG;
    ems_status s;
    if (!h->st.ready && (s = stream_init(h)) != EMS_OK) return s;
    *bytes = blob_size(blob_header(h));
    return EMS_OK;
}

ems_status ems_stream_save(ems_handle* h, void* blob, size_t bytes) {
    if (!h || !blob) return EMS_ERR_INVALID_ARG;
    ems_status s;
    if (!h->st.ready && (s = stream_init(h)) != EMS_OK) return s;
    const StreamBlobHeader b = blob_header(h);
    if (bytes < blob_size(b)) return fail(h, EMS_ERR_INVALID_ARG, "blob too small");
    auto& st = h->st;
    EMS_CUDA(h, cudaStreamSynchronize(h->stream));
    char* p = (char*)blob;
    memcpy(p, &b, sizeof(b)); p += sizeof(b);
    EMS_CUDA(h, cudaMemcpy(p, st.ring, b.ring_bytes, cudaMemcpyDeviceToHost)); p += b.ring_bytes;
    EMS_CUDA(h, cudaMemcpy(p, st.acc, b.acc_bytes, cudaMemcpyDeviceToHost)); p += b.acc_bytes;
    EMS_CUDA(h, cudaMemcpy(p, st.carry, b.carry_bytes, cudaMemcpyDeviceToHost)); p += b.carry_bytes;
    EMS_CUDA(h, cudaMemcpy(p, st.agc, b.agc_bytes, cudaMemcpyDeviceToHost));
    return EMS_OK;
}

ems_status ems_stream_load(ems_handle* h, const void* blob, size_t bytes) {
    if (!h || !blob || bytes < sizeof(StreamBlobHeader)) return EMS_ERR_INVALID_ARG;
    ems_status s;
    if (!h->st.ready && (s = stream_init(h)) != EMS_OK) return s;
    StreamBlobHeader b;
    memcpy(&b, blob, sizeof(b));
    const StreamBlobHeader mine = blob_header(h);
    if (b.magic != kBlobMagic || b.abi != mine.abi || b.n_fft != mine.n_fft || b.hop != mine.hop ||
        b.channels != mine.channels || b.rows != mine.rows || b.det != mine.det ||
        b.ring_bytes != mine.ring_bytes || b.acc_bytes != mine.acc_bytes ||
        b.carry_bytes != mine.carry_bytes || b.agc_bytes != mine.agc_bytes || bytes < blob_size(mine) ||
        b.pushes < 0)
        return fail(h, EMS_ERR_INVALID_ARG, "stream blob does not match this handle");
    auto& st = h->st;
    EMS_CUDA(h, cudaStreamSynchronize(h->stream));
    const char* p = (const char*)blob + sizeof(b);
    const long long pushes = b.pushes;
    EMS_CUDA(h, cudaMemcpy(st.sstate, &pushes, sizeof(long long), cudaMemcpyHostToDevice));
    EMS_CUDA(h, cudaMemcpy(st.ring, p, b.ring_bytes, cudaMemcpyHostToDevice)); p += b.ring_bytes;
    EMS_CUDA(h, cudaMemcpy(st.acc, p, b.acc_bytes, cudaMemcpyHostToDevice)); p += b.acc_bytes;
    EMS_CUDA(h, cudaMemcpy(st.carry, p, b.carry_bytes, cudaMemcpyHostToDevice)); p += b.carry_bytes;
    EMS_CUDA(h, cudaMemcpy(st.agc, p, b.agc_bytes, cudaMemcpyHostToDevice));
    st.pushes = pushes;
    st.last_ready = false;
    return EMS_OK;
}

ems_status ems_stream_reset(ems_handle* h) {
    if (!h) return EMS_ERR_INVALID_ARG;
    if (!h->st.ready) return EMS_OK;
    return stream_zero(h);
}

}  // extern "C"
