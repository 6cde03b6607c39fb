/* reassign_oracle.c — STAND-IN ORACLE in plain C (float64, pthreads).  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: effree/EM-Spec ships no source, no tests and no golden vectors
 * (/root/reference/README.md:73 "The source code is maintained in a private repository";
 * SURVEY.md §0, §8c), so nothing here is EM-Spec's own output.  Like oracle/reassign_oracle.py —
 * whose conventions, drop rule and display shaping this file restates function by function — it
 * follows the published reassignment method the README names (/root/reference/README.md:3,11):
 * Auger & Flandrin, IEEE TSP 43(5), 1995; Fulop & Fitz, JASA 119(1), 2006.  It is pinned by the
 * analytic known-answer tests and by agreement with the NumPy restatement (tests/test_c_oracle.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build,
 * load or call this; the product (em-spec_b200/) never does.  It exists next to the NumPy oracle for
 * two reasons: it is the CPU arm of bench.py with every host thread busy (a fairer baseline than
 * NumPy), and it is fast enough to check the GPU against the oracle on minutes of audio.
 *
 * Arithmetic: everything in double.  The three spectra of a frame come from complex radix-2 FFTs of
 * packed real signals — x h + i x th per frame, x dh of two consecutive frames per FFT — split by
 * Hermitian symmetry; in float64 the packing changes results at the 1e-15 level only.
 *
 * Conventions (oracle/reassign_oracle.py header):
 *   frame m covers samples [mH, mH + N); periodic Hann h; th = (n - N/2) h; dh = (pi/N) sin(2 pi n/N);
 *   e = |X_h|^2 (4/N)^2; dt = Re(X_th conj X_h)/|X_h|^2 [samples]; dk = -Im(X_dh conj X_h)/|X_h|^2 N/(2 pi);
 *   drop rule on the rounded cell; nearest-cell deposit, round-half-even (rint).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_AGC_RELEASE_SECONDS 1.0
#define ORC_LOW_END_CORNER_HZ 200.0
#define ORC_TOP_DB 0.0
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct {
    int32_t n_fft, hop;
    double sample_rate, db_range, gain, low_end_boost, smoothing, noise_gate_db;
    int32_t reassign, display_rows;
    double freq_scale, agc_strength, brightness;
} orc_params;

int orc_abi(void) { return 1; }

/* reassign_oracle.py::frame_count */
int64_t orc_frame_count(int64_t n_samples, int32_t n_fft, int32_t hop) {
    return n_samples < n_fft ? 0 : 1 + (n_samples - n_fft) / hop;
}

static int rows_of(const orc_params* p) { return p->display_rows > 0 ? p->display_rows : p->n_fft / 2 + 1; }

/* ------------------------------------------------------------------------------------------ FFT */
typedef struct {
    int n, stages;
    int* rev;        /* bit reversal */
    double** wr;     /* per stage with half-size h = 2^s: cos(pi j / h), j < h */
    double** wi;     /*                                  -sin(pi j / h)         */
    double *h, *th, *dh;   /* reassign_oracle.py::windows, stored in bit-reversed order: w[rev[n]] = window(n) */
} plan_t;

static void plan_free(plan_t* p) {
    if (!p) return;
    if (p->wr) for (int s = 0; s < p->stages; ++s) free(p->wr[s]);
    if (p->wi) for (int s = 0; s < p->stages; ++s) free(p->wi[s]);
    free(p->wr); free(p->wi); free(p->rev); free(p->h); free(p->th); free(p->dh);
    free(p);
}

static plan_t* plan_make(int n) {
    plan_t* p = calloc(1, sizeof *p);
    if (!p) return NULL;
    p->n = n;
    while ((1 << p->stages) < n) ++p->stages;
    p->rev = malloc(sizeof(int) * n);
    p->wr = calloc(p->stages, sizeof(double*));
    p->wi = calloc(p->stages, sizeof(double*));
    p->h = malloc(sizeof(double) * n); p->th = malloc(sizeof(double) * n); p->dh = malloc(sizeof(double) * n);
    if (!p->rev || !p->wr || !p->wi || !p->h || !p->th || !p->dh) { plan_free(p); return NULL; }
    for (int i = 0; i < n; ++i) {
        int r = 0;
        for (int b = 0; b < p->stages; ++b) r |= ((i >> b) & 1) << (p->stages - 1 - b);
        p->rev[i] = r;
        const double ang = 2.0 * M_PI * (double)i / (double)n;
        p->h[r] = 0.5 - 0.5 * cos(ang);
        p->th[r] = ((double)i - n / 2) * p->h[r];
        p->dh[r] = (M_PI / n) * sin(ang);
    }
    for (int s = 0; s < p->stages; ++s) {
        const int h = 1 << s;
        p->wr[s] = malloc(sizeof(double) * h); p->wi[s] = malloc(sizeof(double) * h);
        if (!p->wr[s] || !p->wi[s]) { plan_free(p); return NULL; }
        for (int j = 0; j < h; ++j) {
            p->wr[s][j] = cos(M_PI * j / h);
            p->wi[s][j] = -sin(M_PI * j / h);
        }
    }
    return p;
}

/* One radix-2 stage (half-size h) of the decimation-in-time FFT, split re / im arrays */
static void stage2(const plan_t* p, int s, double* restrict re, double* restrict im) {
    const int n = p->n, h = 1 << s;
    const double* restrict wr = p->wr[s];
    const double* restrict wi = p->wi[s];
    for (int g = 0; g < n; g += 2 * h) {
        double* restrict ar = re + g; double* restrict ai = im + g;
        double* restrict br = re + g + h; double* restrict bi = im + g + h;
        for (int j = 0; j < h; ++j) {
            const double tr = br[j] * wr[j] - bi[j] * wi[j];
            const double ti = br[j] * wi[j] + bi[j] * wr[j];
            const double xr = ar[j], xi = ai[j];
            ar[j] = xr + tr; ai[j] = xi + ti;
            br[j] = xr - tr; bi[j] = xi - ti;
        }
    }
}

/* Stages s and s + 1 in one pass over the data (the same butterflies and table entries as two
 * calls of stage2, so the same bits; half the memory traffic): one group of 4 h elements */
static inline void group4(int h, double* restrict r0, double* restrict i0, double* restrict r1, double* restrict i1,
                          double* restrict r2, double* restrict i2, double* restrict r3, double* restrict i3,
                          const double* restrict w1r, const double* restrict w1i, const double* restrict w2r,
                          const double* restrict w2i, const double* restrict w3r, const double* restrict w3i) {
    for (int j = 0; j < h; ++j) {
        /* stage s: (0, 1) and (2, 3) with W_{2h}^j */
        const double t1r = r1[j] * w1r[j] - i1[j] * w1i[j], t1i = r1[j] * w1i[j] + i1[j] * w1r[j];
        const double t3r = r3[j] * w1r[j] - i3[j] * w1i[j], t3i = r3[j] * w1i[j] + i3[j] * w1r[j];
        const double b0r = r0[j] + t1r, b0i = i0[j] + t1i, b1r = r0[j] - t1r, b1i = i0[j] - t1i;
        const double b2r = r2[j] + t3r, b2i = i2[j] + t3i, b3r = r2[j] - t3r, b3i = i2[j] - t3i;
        /* stage s + 1: (0, 2) with W_{4h}^j, (1, 3) with W_{4h}^{j+h} */
        const double u2r = b2r * w2r[j] - b2i * w2i[j], u2i = b2r * w2i[j] + b2i * w2r[j];
        const double u3r = b3r * w3r[j] - b3i * w3i[j], u3i = b3r * w3i[j] + b3i * w3r[j];
        r0[j] = b0r + u2r; i0[j] = b0i + u2i; r2[j] = b0r - u2r; i2[j] = b0i - u2i;
        r1[j] = b1r + u3r; i1[j] = b1i + u3i; r3[j] = b1r - u3r; i3[j] = b1i - u3i;
    }
}

static void stage4(const plan_t* p, int s, double* re, double* im) {
    const int n = p->n, h = 1 << s;
    for (int g = 0; g < n; g += 4 * h)
        group4(h, re + g, im + g, re + g + h, im + g + h, re + g + 2 * h, im + g + 2 * h, re + g + 3 * h, im + g + 3 * h,
               p->wr[s], p->wi[s], p->wr[s + 1], p->wi[s + 1], p->wr[s + 1] + h, p->wi[s + 1] + h);
}

/* in-place decimation-in-time FFT of bit-reversed input */
static void fft_dit(const plan_t* p, double* re, double* im) {
    int s = 0;
    if (p->stages & 1) stage2(p, s++, re, im);
    for (; s < p->stages; s += 2) stage4(p, s, re, im);
}

/* ------------------------------------------------------------------------------------------ threads */
typedef void (*range_fn)(void* ctx, int64_t lo, int64_t hi, int tid);
typedef struct { range_fn fn; void* ctx; int64_t lo, hi; int tid; } job_t;
static void* job_run(void* a) { job_t* j = a; j->fn(j->ctx, j->lo, j->hi, j->tid); return NULL; }

/* splits [0, n) into `threads` contiguous ranges whose starts are multiples of `align` */
static int parallel_for(int64_t n, int threads, int64_t align, range_fn fn, void* ctx) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    int64_t per = (n + threads - 1) / threads;
    per = (per + align - 1) / align * align;
    if (per < align) per = align;
    pthread_t th[256]; job_t jobs[256];
    int started = 0, rc = 0;
    for (int t = 0; t < threads; ++t) {
        const int64_t lo = (int64_t)t * per, hi = lo + per < n ? lo + per : n;
        if (lo >= n) break;
        jobs[t] = (job_t){fn, ctx, lo, hi, t};
        if (t == threads - 1 || lo + per >= n) { job_run(&jobs[t]); continue; }    /* the caller takes the last range */
        if (pthread_create(&th[started], NULL, job_run, &jobs[t]) != 0) { job_run(&jobs[t]); continue; }
        ++started;
    }
    for (int t = 0; t < started; ++t) rc |= pthread_join(th[t], NULL);
    return rc;
}

/* ------------------------------------------------------------------------------------------ a1-a3 */
typedef struct {
    const float* x; int64_t F; const orc_params* p; const plan_t* plan;
    double *dcol, *dbin, *en, *raw;
    int failed;
} points_ctx;

/* reassign_oracle.py::reassign_operators + keep_mask + reassign_points for one frame */
static void frame_points(const points_ctx* c, int64_t m, const double* zr, const double* zi,
                         const double* dr, const double* di, int dh_is_imag) {
    const orc_params* p = c->p;
    const int N = p->n_fft, B = N / 2 + 1, H = p->hop;
    const double gate = pow(10.0, p->noise_gate_db / 10.0), norm = (4.0 / N) * (4.0 / N);
    double* dcol = c->dcol + m * B; double* dbin = c->dbin + m * B; double* en = c->en + m * B;
    for (int k = 0; k < B; ++k) {
        const int nk = (N - k) & (N - 1);
        /* Z = FFT(x h + i x th):  X_h = (Z[k] + conj Z[N-k]) / 2,  X_th = (Z[k] - conj Z[N-k]) / 2i */
        const double hr = 0.5 * (zr[k] + zr[nk]), hi = 0.5 * (zi[k] - zi[nk]);
        const double tr = 0.5 * (zi[k] + zi[nk]), ti = -0.5 * (zr[k] - zr[nk]);
        /* D = FFT(x_m dh + i x_{m+1} dh): the same split, real or imaginary part */
        double gr, gi;
        if (!dh_is_imag) { gr = 0.5 * (dr[k] + dr[nk]); gi = 0.5 * (di[k] - di[nk]); }
        else             { gr = 0.5 * (di[k] + di[nk]); gi = -0.5 * (dr[k] - dr[nk]); }
        const double pw = hr * hr + hi * hi;
        double dt = 0.0, dk = 0.0;
        if (pw > 0.0) {
            dt = (tr * hr + ti * hi) / pw;
            dk = -(gi * hr - gr * hi) / pw * (N / (2.0 * M_PI));
        }   /* two divisions, as reassign_oracle.py::reassign_operators */
        const double e = pw * norm;
        if (c->raw) c->raw[m * B + k] = e;
        int ok;
        if (p->reassign) {
            const double col = (double)m + rint(dt / H), row = (double)k + rint(dk);
            ok = e > gate && fabs(dt) <= N / 2 && row >= 0.0 && row <= N / 2 && col >= 0.0 && col <= (double)(c->F - 1);
            dcol[k] = ok ? dt / H : 0.0;
            dbin[k] = ok ? dk : 0.0;
        } else {
            ok = e > gate;
            dcol[k] = 0.0; dbin[k] = 0.0;
        }
        en[k] = ok ? e : 0.0;
    }
}

static void points_range(void* vctx, int64_t lo, int64_t hi, int tid) {
    (void)tid;
    points_ctx* c = vctx;
    const plan_t* pl = c->plan;
    const int N = pl->n, H = c->p->hop;
    const int ld = N + 24;                       /* keeps the eight arrays off each other's cache sets */
    double* buf = malloc(sizeof(double) * 8 * (size_t)ld);
    if (!buf) { __atomic_store_n(&c->failed, 1, __ATOMIC_RELAXED); return; }
    double *ar = buf, *ai = buf + ld, *br = buf + 2 * ld, *bi = buf + 3 * ld, *dr = buf + 4 * ld, *di = buf + 5 * ld,
           *s0 = buf + 6 * ld, *s1 = buf + 7 * ld;
    for (int64_t m = lo; m < hi; m += 2) {       /* frames in pairs: their x dh share one FFT */
        const int two = m + 1 < hi;
        const float* x0 = c->x + m * H;
        const float* x1 = x0 + H;
        for (int n = 0; n < N; ++n) s0[pl->rev[n]] = (double)x0[n];          /* bit-reversed samples */
        if (two) for (int n = 0; n < N; ++n) s1[pl->rev[n]] = (double)x1[n];
        else memset(s1, 0, sizeof(double) * N);
        for (int r = 0; r < N; ++r) {
            ar[r] = s0[r] * pl->h[r]; ai[r] = s0[r] * pl->th[r]; dr[r] = s0[r] * pl->dh[r];
            br[r] = s1[r] * pl->h[r]; bi[r] = s1[r] * pl->th[r]; di[r] = s1[r] * pl->dh[r];
        }
        fft_dit(pl, ar, ai);
        fft_dit(pl, dr, di);
        frame_points(c, m, ar, ai, dr, di, 0);
        if (two) {
            fft_dit(pl, br, bi);
            frame_points(c, m + 1, br, bi, dr, di, 1);
        }
    }
    free(buf);
}

/* reassign_oracle.py::reassign_points — x: float32 samples of one channel; outputs [F][B] doubles
 * (dt in columns, dk in bins, energy); raw (nullable) receives the un-gated energy.  0 on success. */
int orc_points(const float* x, int64_t n_samples, const orc_params* p, double* dcol, double* dbin,
               double* energy, double* raw, int threads) {
    if (!x || !p || !dcol || !dbin || !energy) return 1;
    if (p->n_fft < 2 || (p->n_fft & (p->n_fft - 1)) || p->hop < 1) return 1;
    const int64_t F = orc_frame_count(n_samples, p->n_fft, p->hop);
    if (F == 0) return 0;
    plan_t* pl = plan_make(p->n_fft);
    if (!pl) return 2;
    points_ctx c = {x, F, p, pl, dcol, dbin, energy, raw, 0};
    int rc = parallel_for(F, threads, 2, points_range, &c);
    plan_free(pl);
    return rc || c.failed ? 2 : 0;
}

/* ------------------------------------------------------------------------------------------ a4 */
/* reassign_oracle.py::output_row */
static int64_t output_row(int64_t k, double dk, const orc_params* p) {
    if (p->display_rows <= 0) return k + (int64_t)rint(dk);
    double x = ((double)k + dk) / (p->n_fft / 2);
    x = x < 0.0 ? 0.0 : x > 1.0 ? 1.0 : x;
    const double a = pow(10.0, 2.0 * p->freq_scale) - 1.0;
    const double u = a > 1e-6 ? log1p(a * x) / log1p(a) : x;
    return (int64_t)rint(u * (p->display_rows - 1));
}

typedef struct {
    const double *dcol, *dbin, *en; int64_t F; const orc_params* p; double* grid; int64_t margin; int violated;
} scatter_ctx;

/* destination columns [lo, hi): deposits arrive in (frame, bin) order whatever the thread count */
static void scatter_range(void* vctx, int64_t lo, int64_t hi, int tid) {
    (void)tid;
    scatter_ctx* c = vctx;
    const int B = c->p->n_fft / 2 + 1, R = rows_of(c->p);
    const int64_t f0 = lo - c->margin > 0 ? lo - c->margin : 0;
    const int64_t f1 = hi + c->margin < c->F ? hi + c->margin : c->F;
    for (int64_t f = f0; f < f1; ++f) {
        const double* e = c->en + f * B; const double* dc = c->dcol + f * B; const double* db = c->dbin + f * B;
        const int own = f >= lo && f < hi;
        for (int k = 0; k < B; ++k) {
            if (!(e[k] > 0.0)) continue;
            const double sh = rint(dc[k]);
            if (own && fabs(sh) > (double)c->margin) __atomic_store_n(&c->violated, 1, __ATOMIC_RELAXED);    /* caller-made points: fall back to one thread */
            const int64_t col = f + (int64_t)sh;
            if (col < lo || col >= hi) continue;
            const int64_t row = output_row(k, db[k], c->p);
            if (row < 0 || row >= R) continue;
            c->grid[col * R + row] += e[k];
        }
    }
}

/* reassign_oracle.py::scatter_grid — grid [F][R] doubles, overwritten */
int orc_scatter(const double* dcol, const double* dbin, const double* energy, int64_t F, const orc_params* p,
                double* grid, int threads) {
    if (!dcol || !dbin || !energy || !p || !grid) return 1;
    const int R = rows_of(p);
    memset(grid, 0, sizeof(double) * (size_t)F * R);
    if (F == 0) return 0;
    scatter_ctx c = {dcol, dbin, energy, F, p, grid, p->n_fft / (2 * p->hop) + 2, 0};
    if (threads > 1) {
        if (parallel_for(F, threads, 1, scatter_range, &c)) return 2;
        if (!c.violated) return 0;
        memset(grid, 0, sizeof(double) * (size_t)F * R);
    }
    c.margin = F;
    c.violated = 0;
    scatter_range(&c, 0, F, 0);
    return 0;
}

/* ------------------------------------------------------------------------------------------ a5 */
typedef struct {
    const double* grid; int64_t F; const orc_params* p; double *E, *w, *scale; uint8_t* idx;
} post_ctx;

/* reassign_oracle.py::shaped_energy for rows [lo, hi): E = G gain^2 w_low, EMA over time */
static void shape_range(void* vctx, int64_t lo, int64_t hi, int tid) {
    (void)tid;
    post_ctx* c = vctx;
    const int R = rows_of(c->p);
    const double s = c->p->smoothing, g2 = c->p->gain * c->p->gain;
    for (int64_t m = 0; m < c->F; ++m) {
        const double* g = c->grid + m * R; double* e = c->E + m * R;
        const double* prev = m ? e - R : NULL;
        for (int64_t r = lo; r < hi; ++r) {
            const double v = g[r] * g2 * c->w[r];
            e[r] = s > 0.0 ? s * (prev ? prev[r] : 0.0) + (1.0 - s) * v : v;
        }
    }
}

/* reassign_oracle.py::postpass for columns [lo, hi) */
static void index_range(void* vctx, int64_t lo, int64_t hi, int tid) {
    (void)tid;
    post_ctx* c = vctx;
    const int R = rows_of(c->p);
    const double floor_db = ORC_TOP_DB - c->p->db_range;
    for (int64_t m = lo; m < hi; ++m) {
        const double* e = c->E + m * R; uint8_t* o = c->idx + m * R;
        const double sc = c->scale ? c->scale[m] : 1.0;
        for (int r = 0; r < R; ++r) {
            const double E = e[r];
            const double db_gate = E > 0.0 ? 10.0 * log10(E) : -INFINITY;
            if (!(E > 0.0) || db_gate < c->p->noise_gate_db) { o[r] = 0; continue; }
            const double db = c->scale ? 10.0 * log10(E * sc) : db_gate;
            double v = rint(255.0 * (db - floor_db) / c->p->db_range);
            v = v < 0.0 ? 0.0 : v > 255.0 ? 255.0 : v;
            o[r] = (uint8_t)v;
        }
    }
}

/* reassign_oracle.py::postpass (with low_end_weight, row_frequencies, shaped_energy, agc_scale) —
 * grid [F][R] doubles -> colour index [F][R] */
static int postpass_impl(const double* grid, int64_t F, const orc_params* p, uint8_t* index, int threads, double* E_scratch) {
    if (!grid || !p || !index) return 1;
    if (F == 0) return 0;
    const int R = rows_of(p);
    post_ctx c = {grid, F, p, E_scratch ? E_scratch : malloc(sizeof(double) * (size_t)F * R), malloc(sizeof(double) * R), NULL, index};
    if (!c.E || !c.w) { if (!E_scratch) free(c.E); free(c.w); return 2; }
    const double a = pow(10.0, 2.0 * p->freq_scale) - 1.0;
    for (int r = 0; r < R; ++r) {
        double f;
        if (p->display_rows <= 0) {
            f = (double)r * p->sample_rate / p->n_fft;
        } else {
            const double u = (double)r / (p->display_rows - 1);
            f = (a > 1e-6 ? expm1(u * log1p(a)) / a : u) * p->sample_rate / 2;
        }
        const double q = f / ORC_LOW_END_CORNER_HZ;
        c.w[r] = 1.0 + (p->low_end_boost - 1.0) / (1.0 + q * q);
    }
    int rc = parallel_for(R, threads, 8, shape_range, &c);
    if (!rc && p->agc_strength > 0.0) {          /* reassign_oracle.py::agc_scale */
        c.scale = malloc(sizeof(double) * (size_t)F);
        if (!c.scale) { if (!E_scratch) free(c.E); free(c.w); return 2; }
        const double lam = exp(-p->hop / (p->sample_rate * ORC_AGC_RELEASE_SECONDS));
        const double target = pow(10.0, -(1.0 - p->brightness) * p->db_range / 10.0);
        double lv = 0.0;
        for (int64_t m = 0; m < F; ++m) {
            double peak = 0.0;
            const double* e = c.E + m * R;
            for (int r = 0; r < R; ++r) peak = e[r] > peak ? e[r] : peak;
            lv = peak > lam * lv ? peak : lam * lv;
            c.scale[m] = lv > 0.0 ? target * pow(lv, -p->agc_strength) : 1.0;
        }
    }
    if (!rc) rc = parallel_for(F, threads, 1, index_range, &c);
    if (!E_scratch) free(c.E);
    free(c.w); free(c.scale);
    return rc ? 2 : 0;
}

int orc_postpass(const double* grid, int64_t F, const orc_params* p, uint8_t* index, int threads) {
    return postpass_impl(grid, F, p, index, threads, NULL);
}

/* reassign_oracle.py::process — one channel: grid [F][R] doubles and colour index [F][R].
 * The float64 intermediates (points, shaped energy) live in a grow-only scratch kept between calls, so a
 * caller that walks a long stream in slices does not page-fault a gigabyte per slice; calls must not overlap. */
static double* g_scratch = NULL;
static size_t g_scratch_n = 0;

void orc_release_scratch(void) { free(g_scratch); g_scratch = NULL; g_scratch_n = 0; }

int orc_process(const float* x, int64_t n_samples, const orc_params* p, double* grid, uint8_t* index, int threads) {
    if (!x || !p || !grid || !index) return 1;
    const int64_t F = orc_frame_count(n_samples, p->n_fft, p->hop);
    if (F == 0) return 0;
    const size_t n = (size_t)F * (p->n_fft / 2 + 1), ne = (size_t)F * rows_of(p);
    if (g_scratch_n < 3 * n + ne) {
        free(g_scratch);
        g_scratch_n = 3 * n + ne;
        g_scratch = malloc(sizeof(double) * g_scratch_n);
        if (!g_scratch) { g_scratch_n = 0; return 2; }
    }
    double* pts = g_scratch;
    int rc = orc_points(x, n_samples, p, pts, pts + n, pts + 2 * n, NULL, threads);
    if (!rc) rc = orc_scatter(pts, pts + n, pts + 2 * n, F, p, grid, threads);
    if (!rc) rc = postpass_impl(grid, F, p, index, threads, pts + 3 * n);
    return rc;
}
