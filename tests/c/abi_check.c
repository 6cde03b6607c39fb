/* abi_check.c — the C-ABI of include/emspec.h seen from plain C (gcc, no CUDA headers).
 * Compile-time: the header is valid C99 and ems_params has the layout the ctypes mirror assumes.
 * `abi_check layout` prints sizeof / offsetof for the Python test to compare with the mirror.
 * `abi_check run` (needs a GPU) makes one ems_process_host call on a 1 kHz tone and checks that
 * the brightest row of the image is the tone's bin.  Test infrastructure, not product. */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "emspec.h"

_Static_assert(sizeof(ems_params) == 56, "ems_params is 14 four-byte fields");
_Static_assert(offsetof(ems_params, n_fft) == 0, "n_fft");
_Static_assert(offsetof(ems_params, hop) == 4, "hop");
_Static_assert(offsetof(ems_params, sample_rate) == 8, "sample_rate");
_Static_assert(offsetof(ems_params, channels) == 12, "channels");
_Static_assert(offsetof(ems_params, db_range) == 16, "db_range");
_Static_assert(offsetof(ems_params, gain) == 20, "gain");
_Static_assert(offsetof(ems_params, low_end_boost) == 24, "low_end_boost");
_Static_assert(offsetof(ems_params, smoothing) == 28, "smoothing");
_Static_assert(offsetof(ems_params, noise_gate_db) == 32, "noise_gate_db");
_Static_assert(offsetof(ems_params, flags) == 36, "flags");
_Static_assert(offsetof(ems_params, display_rows) == 40, "display_rows");
_Static_assert(offsetof(ems_params, freq_scale) == 44, "freq_scale");
_Static_assert(offsetof(ems_params, agc_strength) == 48, "agc_strength");
_Static_assert(offsetof(ems_params, brightness) == 52, "brightness");
_Static_assert(sizeof(ems_status) == sizeof(int), "status codes travel as int");
_Static_assert(sizeof(ems_cursor) == 32 && offsetof(ems_cursor, freq_hz) == 8 && offsetof(ems_cursor, midi_note) == 16 &&
               offsetof(ems_cursor, cents) == 20 && offsetof(ems_cursor, name) == 24, "ems_cursor layout");

#define OFF(f) printf(#f " %zu\n", offsetof(ems_params, f))

static int layout(void) {
    printf("sizeof %zu\n", sizeof(ems_params));
    OFF(n_fft); OFF(hop); OFF(sample_rate); OFF(channels); OFF(db_range); OFF(gain); OFF(low_end_boost);
    OFF(smoothing); OFF(noise_gate_db); OFF(flags); OFF(display_rows); OFF(freq_scale); OFF(agc_strength);
    OFF(brightness);
    printf("abi %d\n", EMS_ABI_VERSION);
    return 0;
}

static int run(void) {
    ems_params p;
    ems_handle* h = NULL;
    if (ems_abi_version() != EMS_ABI_VERSION) { fprintf(stderr, "library ABI %d, header %d\n", ems_abi_version(), EMS_ABI_VERSION); return 2; }
    if (ems_default_params(&p) != EMS_OK) return 3;
    p.n_fft = 2048; p.hop = 256; p.flags |= EMS_FLAG_SYNC;
    ems_status s = ems_create(&p, &h);
    if (s != EMS_OK) { fprintf(stderr, "ems_create: %s\n", ems_status_str(s)); return 4; }
    const size_t S = 48000;
    size_t F = 0, R = 0;
    ems_frame_count(h, S, &F);
    ems_output_rows(h, &R);
    float* x = (float*)malloc(S * sizeof(float));
    uint8_t* img = (uint8_t*)calloc(F * R, 1);
    const int k0 = 43;                                   /* tone exactly on bin 43: 43 * 48000 / 2048 Hz */
    for (size_t i = 0; i < S; ++i) x[i] = 0.5f * (float)sin(2.0 * 3.14159265358979323846 * k0 * (double)i / 2048.0);
    size_t nf = 0;
    s = ems_process_host(h, x, S, NULL, img, &nf);       /* pageable host buffers: allowed, just slower */
    if (s != EMS_OK) { fprintf(stderr, "ems_process_host: %s: %s\n", ems_status_str(s), ems_last_error(h)); return 5; }
    if (nf != F || F != 1 + (S - 2048) / 256 || R != 1025) { fprintf(stderr, "geometry %zu %zu %zu\n", nf, F, R); return 6; }
    int bad = 0;
    for (size_t f = 8; f + 8 < F; ++f) {
        size_t best = 0;
        for (size_t r = 1; r < R; ++r) if (img[f * R + r] > img[f * R + best]) best = r;
        if ((int)best != k0 || img[f * R + best] == 0) ++bad;
    }
    ems_cursor cur;                                      /* the tone's row reads back as its frequency and note */
    if (ems_cursor_info(h, 8.0, (double)k0, &cur) != EMS_OK) return 9;
    if (fabs(cur.freq_hz - k0 * 48000.0 / 2048.0) > 1e-9 || strcmp(cur.name, "B5") != 0 || cur.midi_note != 83 ||
        fabs(cur.time_s - (8 * 256 + 1024) / 48000.0) > 1e-12) {
        fprintf(stderr, "cursor %g Hz %s %d %+.1f cents %g s\n", cur.freq_hz, cur.name, cur.midi_note, cur.cents, cur.time_s);
        return 10;
    }
    double row = -1.0;
    if (ems_hz_to_row(h, cur.freq_hz, &row) != EMS_OK || fabs(row - k0) > 1e-9) return 12;
    uint32_t lut[256];
    if (ems_colormap_count() < 2 || strcmp(ems_colormap_name(0), "inferno") != 0 || ems_colormap_builtin(1, lut) != EMS_OK || lut[255] != 0xFFFFFFFFu ||
        lut[0] != 0xFF000000u || ems_colormap_name(ems_colormap_count()) != NULL) return 11;   /* id 1 = gray: black to white */
    size_t scratch = 0;
    ems_scratch_bytes(h, &scratch);
    printf("frames %zu rows %zu bad_columns %d scratch_bytes %zu\n", F, R, bad, scratch);
    free(x); free(img);
    if (ems_destroy(h) != EMS_OK) return 7;
    return bad ? 8 : 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && strcmp(argv[1], "layout") == 0) return layout();
    if (argc > 1 && strcmp(argv[1], "run") == 0) return run();
    fprintf(stderr, "usage: abi_check layout|run\n");
    return 1;
}
