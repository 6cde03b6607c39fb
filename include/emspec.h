/*
 * emspec.h — C-ABI of the B200-native reassigned-spectrogram engine.
 *
 * STAND-IN BOUNDARY.  effree/EM-Spec exposes no plugin / operator / FFI interface:
 * it is a GUI application whose source is private (/root/reference/README.md:73) and
 * whose only documented control surface is the settings panel
 * (/root/reference/README.md:41-51).  Each entry point below therefore cites the
 * README control or feature it stands in for, not a reference function.
 * SURVEY.md §8b is the contract this header implements.
 *
 * Conventions
 *   - plain C symbols, POD structs, caller owns every buffer it passes in;
 *   - every call returns ems_status (0 = OK); nothing throws or aborts across the ABI;
 *   - one handle = one CUDA stream; a handle is not thread-safe, handles are independent;
 *   - offline calls are asynchronous on the handle's stream unless EMS_FLAG_SYNC is set;
 *   - "dev" pointers are CUDA device pointers on the handle's device, "host" pointers
 *     are host memory (pinned memory makes the copies asynchronous);
 *   - there is no CPU fallback: without a CUDA device ems_create fails with
 *     EMS_ERR_CUDA.
 *
 * Geometry: N = n_fft, H = hop, B = N/2+1 bins, R = output rows per column
 *   (B, or display_rows when that is set),
 *   F = 0 if S < N else 1 + (S-N)/H frames per channel for S samples per channel.
 *   Frame f covers samples [f*H, f*H+N); column f of every output is centred there.
 */
#ifndef EMSPEC_H_
#define EMSPEC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMS_ABI_VERSION 5

typedef enum ems_status {
    EMS_OK = 0,
    EMS_ERR_INVALID_ARG = 1, /* null pointer, n_fft not a power of two in range, hop <= 0, ... */
    EMS_ERR_UNSUPPORTED = 2, /* valid request this build cannot serve */
    EMS_ERR_CUDA = 3,        /* CUDA runtime error (ems_last_error has the text) */
    EMS_ERR_NOMEM = 4,       /* device or host allocation failed */
    EMS_ERR_STATE = 5        /* call not valid in the handle's current state */
} ems_status;

/* flags */
#define EMS_FLAG_REASSIGN      1u /* 1: reassigned ("Enhanced"), 0: plain |X_h|^2 columns ("Natural");
                                     assets/settings.png buttons */
#define EMS_FLAG_DETERMINISTIC 2u /* scatter accumulates in 64-bit fixed point (2^-44 steps, a point
                                     saturates at 2^12 = +36 dB re full scale): order-independent,
                                     bit-exact across runs; 0: fp32 red.global.add fast mode.
                                     PCM is expected in [-1, 1] (full scale 1.0). */
#define EMS_FLAG_SYNC          4u /* offline calls synchronise the stream before returning */
#define EMS_FLAG_BOUNDED_SCRATCH 8u /* ems_process_grid works in frame chunks on an accumulator ring instead of an
                                     accumulator the size of the call (22 GB for an hour of mono at 4096/128
                                     becomes 2 GB); same bits out, a few per cent slower.  ems_process_host*
                                     always work that way. */

#define EMS_FLAG_SORTED_SCATTER 16u /* ems_scatter_points deposits by sort-by-cell + segmented reduce (the scatter
                                     BASELINE.json's north_star names): a stable radix sort orders the points by
                                     destination cell, each cell's energies are added in fp32 in ascending point
                                     order.  Defined by the point order alone (bit-exact across runs), no fixed-point
                                     quantisation; up to 1 GB of extra scratch; slower than the default
                                     deterministic mode.  Ignored by the fused ems_process_* paths. */

/* Parameter surface = the README settings glossary (/root/reference/README.md:41-51);
 * defaults in comments are the "Default" preset of assets/settings.png. */
typedef struct ems_params {
    int32_t  n_fft;         /* "FFT Size" README.md:43; power of two, 256..32768      (4096) */
    int32_t  hop;           /* "Scroll Speed" README.md:44 maps to hop, 1..n_fft      (128)  */
    float    sample_rate;   /* Hz                                                     (48000)*/
    int32_t  channels;      /* planar channels (or equal-length clips of a batch) per call,
                               1..65535                                              (1)    */
    float    db_range;      /* "dB Range" README.md:46; floor = 0 dB - range          (58)   */
    float    gain;          /* "Gain" README.md:47; linear amplitude                  (3.5)  */
    float    low_end_boost; /* "Low-End Boost" README.md:49; weight at DC             (3.9)  */
    float    smoothing;     /* "Smoothing" README.md:50; EMA coefficient in [0,1)     (0.0)  */
    float    noise_gate_db; /* "Noise Gate" README.md:51; dB re full-scale sine       (-65)  */
    uint32_t flags;         /* EMS_FLAG_*                                                     */
    int32_t  display_rows;  /* 0: one output row per bin (R = n_fft/2+1).  > 0: energy is
                               scattered straight onto R = display_rows rows of a warped
                               frequency axis (assets/spectrogram.png is 546 px high)  (0)    */
    float    freq_scale;    /* "Frequency Scale" README.md:48, used when display_rows > 0:
                               row = (R-1) * log1p(a x) / log1p(a), x = f / Nyquist,
                               a = 10^(2*freq_scale) - 1; 0 = linear axis             (1.0)  */
    float    agc_strength;  /* "AGC Strength" / "Auto Gain" README.md:14, settings.png; 0 = off.
                               level[m] = max(peak[m], lambda*level[m-1]), peak = loudest shaped
                               cell of column m, lambda = exp(-hop/(sample_rate * 1 s)); cells are
                               drawn at E / level^strength (the gate still sees E)      (0.0)  */
    float    brightness;    /* "Brightness" (assets/settings.png, 44 %), used when agc_strength > 0:
                               where on the colour scale the automatic gain places the running
                               level — cells are drawn at E * T / level^strength with
                               T = 10^(-(1 - brightness) * db_range / 10), so at full strength the
                               loudest cell gets colour index 255 * brightness; in (0, 1]   (0.44) */
} ems_params;

typedef struct ems_handle ems_handle;

/* Stage ids for ems_stage_ms (device time of the last offline call, CUDA events). */
#define EMS_STAGE_POINTS  0 /* fused frame gather + 3-window STFT + reassignment (a1-a3) */
#define EMS_STAGE_SCATTER 1 /* energy scatter onto the grid (a4) */
#define EMS_STAGE_POST    2 /* dB / gate / boost / smoothing / colour index (a5) */
#define EMS_STAGE_COUNT   3

int         ems_abi_version(void);
const char* ems_status_str(ems_status s);
/* Text of the last error on this handle (never NULL). */
const char* ems_last_error(const ems_handle* h);

/* Fills *p with the settings.png "Default" preset. */
ems_status ems_default_params(ems_params* p);

/* Creates an engine on the current CUDA device.  Stands in for launching the app
 * with a preset (/root/reference/README.md:35-38). */
ems_status ems_create(const ems_params* params, ems_handle** out);
ems_status ems_destroy(ems_handle* h);

/* Live display controls (README.md:41 "changes are applied in real-time"): updates
 * db_range, gain, low_end_boost, smoothing, noise_gate_db, agc_strength, brightness and flags; n_fft / hop /
 * channels changes require a new handle (EMS_ERR_INVALID_ARG). */
ems_status ems_update_display(ems_handle* h, const ems_params* params);

/* Run on a caller-provided cudaStream_t instead of the handle's own stream
 * (NULL = the CUDA default stream). */
ems_status ems_set_stream(ems_handle* h, void* cuda_stream);
ems_status ems_get_stream(ems_handle* h, void** cuda_stream);
ems_status ems_synchronize(ems_handle* h);

/* Output rows per column R of this handle (n_fft/2+1, or display_rows). */
ems_status ems_output_rows(const ems_handle* h, size_t* rows);

/* F for n_samples_per_ch samples with this handle's n_fft / hop. */
ems_status ems_frame_count(const ems_handle* h, size_t n_samples_per_ch, size_t* n_frames);

/* Cursor readout (/root/reference/README.md:39 "Hold Shift and hover over the spectrogram to see note
 * and frequency information"): what lies under output cell (column, row) of this handle's pictures.
 * Pure host arithmetic on the handle's geometry — the inverse of the row mapping the scatter uses
 * (row -> frequency; fractional rows follow the axis between row centres, rows are clamped to
 * [0, R-1]) and the frame clock (column f is centred on sample f*H + N/2). */
typedef struct ems_cursor {
    double  time_s;    /* centre of the column [s] */
    double  freq_hz;   /* frequency of the row [Hz] */
    int32_t midi_note; /* nearest equal-tempered note, A4 = 440 Hz = 69; -1 when freq_hz < 1 Hz */
    float   cents;     /* freq_hz relative to that note, in [-50, 50] */
    char    name[8];   /* "A4", "C#3", "D-1"; "" when midi_note = -1 */
} ems_cursor;
ems_status ems_cursor_info(const ems_handle* h, double column, double row, ems_cursor* out);
/* The other direction, for axis ticks and note grid lines over the picture: the (fractional) output row
 * of a frequency — exactly the mapping the scatter rounds to place a point (bin axis: f n_fft / sample_rate;
 * warped display axis: (R-1) log1p(a x) / log1p(a)), clamped to [0, R-1]. */
ems_status ems_hz_to_row(const ems_handle* h, double freq_hz, double* row);

/* Built-in colour maps ("Multiple Color Maps", /root/reference/README.md:15,45): 256 packed
 * 0xAABBGGRR pixels for ems_colorize / ems_stream_set_colormap.  The maps are stand-ins (EM-Spec's own
 * tables are not published): piecewise-linear ramps through a few 8-bit control colours, evaluated in
 * integer arithmetic (exactly reproducible).  id 0 is "inferno", the map of the settings.png "Default"
 * preset.  ids 0..ems_colormap_count()-1; ems_colormap_name
 * returns NULL for any other id. */
int         ems_colormap_count(void);
const char* ems_colormap_name(int id);
ems_status  ems_colormap_builtin(int id, uint32_t lut_rgba_host[256]);

/* a1-a3, "reassignment method" (/root/reference/README.md:3,11).
 * pcm_dev: fp32 planar [channels][n_samples_per_ch].
 * dt_cols, dk_bins, energy: fp32 [channels][F][B] each (device).  A point of frame f,
 * bin k sits at column f + dt_cols, bin k + dk_bins with energy |X_h|^2 (4/N)^2.
 * Dropped points (SURVEY.md §7 "Out-of-support points") carry energy 0, dt = dk = 0. */
ems_status ems_process_points(ems_handle* h, const float* pcm_dev, size_t n_samples_per_ch,
                              float* dt_cols, float* dk_bins, float* energy,
                              size_t* n_frames);

/* a1-a5: the picture the app draws (/root/reference/assets/spectrogram.png).
 * grid_dev : fp32 [channels][F][R] accumulated energy, or NULL;
 * index_dev: u8   [channels][F][R] colour index 0..255, or NULL (not both NULL). */
ems_status ems_process_grid(ems_handle* h, const float* pcm_dev, size_t n_samples_per_ch,
                            float* grid_dev, uint8_t* index_dev, size_t* n_frames);

/* a4 + a5 from caller-held points (as written by ems_process_points, possibly edited):
 * deposits energy > 0 points at (f + rint(dt_cols), k + rint(dk_bins)), then the
 * post-pass.  All pointers are device pointers; grid_dev / index_dev as above. */
ems_status ems_scatter_points(ems_handle* h, const float* dt_cols, const float* dk_bins,
                              const float* energy, size_t n_frames,
                              float* grid_dev, uint8_t* index_dev);

/* Same as ems_process_grid with HOST buffers: chunks the stream, overlaps H2D, compute
 * and D2H on two streams, and returns when index_host (and grid_host if not NULL) are
 * complete.  Pinned buffers recommended.  This is the call an application makes. */
ems_status ems_process_host(ems_handle* h, const float* pcm_host, size_t n_samples_per_ch,
                            float* grid_host, uint8_t* index_host, size_t* n_frames);

/* Capture-side input format (/root/reference/README.md:36 "system audio"):
 * int16 PCM, interleaved [n_samples_per_ch][channels], full scale 32768; otherwise as
 * ems_process_host (half the host-to-device bytes). */
ems_status ems_process_host_i16(ems_handle* h, const int16_t* pcm_host, size_t n_samples_per_ch,
                                float* grid_host, uint8_t* index_host, size_t* n_frames);

/* int24 PCM, packed little-endian, interleaved [n_samples_per_ch][channels][3 bytes], full scale
 * 2^23 (the other common capture / file format; 3/4 of the fp32 bytes); otherwise as above. */
ems_status ems_process_host_i24(ems_handle* h, const uint8_t* pcm_host, size_t n_samples_per_ch,
                                float* grid_host, uint8_t* index_host, size_t* n_frames);

/* Device memory this handle holds for scratch right now (accumulator, flags, staging, tables of
 * the smoothing / AGC scans) — ems_process_host* keep it independent of the stream length. */
ems_status ems_scratch_bytes(const ems_handle* h, size_t* bytes);

/* Per-clip summary of a colour-index image for batch jobs (clips are independent, the
 * only exchange is a final gather of per-clip descriptors).  index_dev: u8 [channels][n_frames][R];
 * summary_dev: uint64 [channels][2] (device): [0] = sum of the bytes, [1] = sum of byte * (1 + position
 * mod 65521) in wrapping 64-bit arithmetic (order-dependent).  Integer work: exact, any alignment. */
ems_status ems_image_summary(ems_handle* h, const uint8_t* index_dev, size_t n_frames,
                             uint64_t* summary_dev);

/* Colour map ("Color Map", /root/reference/README.md:15,45): applies a
 * 256-entry RGBA table (HOST pointer, packed 0xAABBGGRR like a byte-wise R,G,B,A store) to
 * n_cells colour indices on the device; rgba_dev receives n_cells 32-bit pixels in the same
 * [channels][F][R] order, ready to blit as columns. */
ems_status ems_colorize(ems_handle* h, const uint8_t* index_dev, size_t n_cells,
                        const uint32_t* lut_rgba_host, uint32_t* rgba_dev);

/* Device milliseconds of a stage of the last offline call on this handle (after the
 * stream has been synchronised); EMS_ERR_STATE if that stage did not run. */
ems_status ems_stage_ms(ems_handle* h, int stage, float* ms);
/* Kernel launches issued by this handle since creation. */
ems_status ems_launch_count(const ems_handle* h, uint64_t* launches);

/* Streaming mode ("start visualizing your system audio", /root/reference/README.md:36).
 * pcm_host: hop*channels fp32 samples, interleaved.  Once the ring holds n_fft samples
 * every push analyses one new frame per channel.  A column is final ceil(n_fft/(2 hop))
 * pushes after its own frame; then *column_ready = 1, column_host (u8 [channels][R], pinned
 * recommended) holds it and *column_index (nullable) its frame index.  Synchronous. */
ems_status ems_stream_push(ems_handle* h, const float* pcm_host, uint8_t* column_host,
                           int* column_ready, int64_t* column_index);
/* The same push with the hop in capture format: int16, interleaved [hop][channels], full scale
 * 32768.  A stream may switch formats between pushes. */
ems_status ems_stream_push_i16(ems_handle* h, const int16_t* pcm_host, uint8_t* column_host,
                               int* column_ready, int64_t* column_index);
/* ... or packed little-endian int24, three bytes per sample, interleaved [hop][channels], full scale 2^23
 * (as ems_process_host_i24). */
ems_status ems_stream_push_i24(ems_handle* h, const uint8_t* pcm_host, uint8_t* column_host,
                               int* column_ready, int64_t* column_index);
/* Colour map on the streaming path ("ready-to-blit columns for a scrolling
 * renderer", /root/reference/README.md:15,45).  lut_rgba_host: 256 packed pixels as in
 * ems_colorize, copied; NULL switches the map off.  While a map is set the kernel that finishes a
 * push also writes the final column as pixels, and ems_stream_column_rgba copies the pixels
 * (u32 [channels][R]) of the column the last push delivered; EMS_ERR_STATE when no map is set or
 * that push had column_ready = 0.  The map is configuration, not stream state: it is not part
 * of a checkpoint and survives ems_stream_reset. */
ems_status ems_stream_set_colormap(ems_handle* h, const uint32_t* lut_rgba_host);
ems_status ems_stream_column_rgba(ems_handle* h, uint32_t* rgba_host);
/* Clears the ring, the rolling grid and the smoothing state. */
ems_status ems_stream_reset(ems_handle* h);

/* Checkpoint / resume of a stream: the state is the push counter, the sample
 * ring, the rolling accumulator columns, the smoothing and AGC state.  A blob saved from one
 * handle can be loaded into another handle created with the same parameters; the columns
 * that follow are bit-identical.  ems_stream_state_size is valid after the first push (or
 * after a load); load returns EMS_ERR_INVALID_ARG when the blob does not match the handle. */
ems_status ems_stream_state_size(ems_handle* h, size_t* bytes);
ems_status ems_stream_save(ems_handle* h, void* blob_host, size_t bytes);
ems_status ems_stream_load(ems_handle* h, const void* blob_host, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* EMSPEC_H_ */
