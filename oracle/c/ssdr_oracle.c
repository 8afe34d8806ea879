/*
 * TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or called from the product path.
 *
 * Scalar C restatement of the waterfall hot path as specified in DESIGN.md section 4:
 *
 *   IQ frame -> Hann window -> mixed-radix DIF FFT (float32, fixed operation order)
 *            -> |X|^2 -> Kiwi byte (dBm + 255, utils_supersdr.py:789) by threshold counting
 *            -> time-binning mean of n byte lines (utils_supersdr.py:881-886)
 *            -> dB cal + percentile auto-scale + colour row (utils_supersdr.py:787-813)
 *            -> uint8 pixel row.
 *
 * The FFT / log-magnitude stage is ABSENT from the reference (it runs on the KiwiSDR server):
 * PARITY UNPINNED for that stage; this file is the builder-defined, bit-exact statement of it.
 * The stages after the byte line restate reference arithmetic and are pinned through
 * oracle/tier_p.py (same results, checked in tests/test_oracle_*.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no implicit FMA contraction; every fused
 * multiply-add in the spec is an explicit fmaf()).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SO_FS 32768.0
#define SO_PI 3.14159265358979323846
#define SO_MAX_PASSES 8

typedef struct { float re, im; } cpx;

/* ---- spec constants ----------------------------------------------------- */
static const float K_A = 0.92387953251128674f;   /* cos(pi/8)  */
static const float K_B = 0.70710678118654752f;   /* sqrt(1/2)  */
static const float K_C = 0.38268343236508977f;   /* sin(pi/8)  */

/* cos(2*pi*m/32), sin(2*pi*m/32) built from the first-octant values by symmetry (exact negations) */
static const float K_Q32[9] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                               0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f,
                               0.19509032201612825f, 0.0f};
static void unit32(int m, float* c, float* s) {
    int mm = m & 31, quad = mm >> 3, r = mm & 7;
    float cr = K_Q32[r], sr = K_Q32[8 - r];
    switch (quad) {
    case 0: *c = cr;  *s = sr;  break;
    case 1: *c = -sr; *s = cr;  break;
    case 2: *c = -cr; *s = -sr; break;
    default: *c = sr; *s = -cr; break;
    }
}

/* ---- plan: radix list, product == N ---------------------------------------- */
int so_fft_plan(int N, int* radices) {
    int lg = 0;
    while ((1 << lg) < N) lg++;
    if ((1 << lg) != N || lg < 6 || lg > 16) return -1;
    /* lg <= 14: lg = r + 5k: one first pass of radix 2^r (r = 1..4; a radix-32 pass when r == 0) that reads
     * the input and applies the window, followed by k radix-32 passes.
     * lg = 15, 16: a front pass of radix 2^(lg - 14) (window, chain twiddles) followed by the 16384-point plan
     * 16 x 32 x 32 -- the frame no longer fits one SM's shared memory, the front pass is its own kernel. */
    int np = 0;
    if (lg > 14) {
        radices[np++] = 1 << (lg - 14);
        lg = 14;
    }
    int k = lg / 5, r = lg - 5 * k;
    if (r == 0) { r = 5; k -= 1; }
    radices[np++] = 1 << r;
    for (int i = 0; i < k; ++i) radices[np++] = 32;
    return np;
}

void so_twiddle_table(int N, float* tab) {
    for (int k = 0; k < N; ++k) {
        double a = 2.0 * SO_PI * (double)k / (double)N;
        tab[2 * k] = (float)cos(a);
        tab[2 * k + 1] = (float)(-sin(a));
    }
}

/* First half of the periodic Hann window, w[n] = float(0.5 - 0.5 cos(2 pi n / N)), n < N/2 */
void so_window_table(int N, float* win) {
    for (int n = 0; n < N / 2; ++n)
        win[n] = (float)(0.5 - 0.5 * cos(2.0 * SO_PI * (double)n / (double)N));
}

/* byte >= k  <=>  P >= T[k], k = 1..255;  T[0] = 0 */
void so_thresholds(int N, double cal_db, float* T) {
    double ref = ((double)N * SO_FS * 0.5);
    ref = ref * ref;
    T[0] = 0.0f;
    for (int k = 1; k < 256; ++k)
        T[k] = (float)(ref * pow(10.0, ((double)k - 0.5 - 255.0 - cal_db) / 10.0));
}

/* ---- arithmetic primitives (each line is one IEEE float32 operation) ----------------------- */
static inline cpx cadd(cpx a, cpx b) { cpx r = {a.re + b.re, a.im + b.im}; return r; }
static inline cpx csub(cpx a, cpx b) { cpx r = {a.re - b.re, a.im - b.im}; return r; }
static inline cpx cmul(cpx u, cpx w) {
    cpx r;
    float t0 = u.im * w.im;
    float t1 = u.re * w.im;
    r.re = fmaf(u.re, w.re, -t0);
    r.im = fmaf(u.im, w.re, t1);
    return r;
}
static inline cpx mul_mi(cpx u) { cpx r = {u.im, -u.re}; return r; }                 /* * (-i)      */
/* rotations by odd eighth turns are ordinary complex multiplies by the rounded constants */
static inline cpx mul_w8(cpx u) { cpx w = {K_B, -K_B}; return cmul(u, w); }      /* * B(1-i)  = W8^1 */
static inline cpx mul_w83(cpx u) { cpx w = {-K_B, -K_B}; return cmul(u, w); }    /* * -B(1+i) = W8^3 */

/* ---- butterflies, DESIGN.md 4.2 -------------------------------------------------------------
 * Every radix-r butterfly (natural order in and out) starts with the same first level,
 *     (x[m], x[m + r/2]) <- (x[m] + x[m + r/2], x[m] - x[m + r/2]),   m < r/2,
 * followed by dft_rest(r).  The first pass of the transform replaces that level by the windowed
 * form l1_window(), which fuses the Hann multiply into it. */
static void l1(cpx* x, int r) {
    for (int m = 0; m < r / 2; ++m) {
        cpx a = x[m], b = x[m + r / 2];
        x[m] = cadd(a, b);
        x[m + r / 2] = csub(a, b);
    }
}

/* Windowed first level (first pass of the transform only).  The pair (x[m], x[m+h]) are the samples
 * n and n + N/2 of the frame; the periodic Hann window satisfies w[n + N/2] = 1 - w[n], so with w = w[n]
 *     x[m]   <- x[m] w + x[m+h] (1 - w) = fma(x[m] - x[m+h], w,  x[m+h])
 *     x[m+h] <- x[m] w - x[m+h] (1 - w) = fma(x[m] + x[m+h], w, -x[m+h])
 * (one rounded add / subtract and one fused multiply-add per output). */
static void l1_window(cpx* x, int r, const float* w) {
    int h = r / 2;
    for (int m = 0; m < h; ++m) {
        cpx a = x[m], b = x[m + h];
        float dr = a.re - b.re, di = a.im - b.im, sr = a.re + b.re, si = a.im + b.im;
        x[m].re = fmaf(dr, w[m], b.re);      x[m].im = fmaf(di, w[m], b.im);
        x[m + h].re = fmaf(sr, w[m], -b.re); x[m + h].im = fmaf(si, w[m], -b.im);
    }
}

/* second level of a radix-4: in (a, c, b, e) = (x0+x2, x1+x3, x0-x2, x1-x3) at x[0], x[1], x[2], x[3] */
static inline void dft4_rest(cpx* x0, cpx* x1, cpx* x2, cpx* x3) {
    cpx a = *x0, c = *x1, b = *x2, d = mul_mi(*x3);
    *x0 = cadd(a, c); *x2 = csub(a, c);
    *x1 = cadd(b, d); *x3 = csub(b, d);
}

static inline void dft4(cpx* x) {
    l1(x, 4);
    dft4_rest(&x[0], &x[1], &x[2], &x[3]);
}

static void dft8_rest(cpx* x) {              /* after l1: x[0..3] sums, x[4..7] differences */
    cpx u[2][4];
    for (int m0 = 0; m0 < 4; ++m0) { u[0][m0] = x[m0]; u[1][m0] = x[m0 + 4]; }
    u[1][1] = mul_w8(u[1][1]);
    u[1][2] = mul_mi(u[1][2]);
    u[1][3] = mul_w83(u[1][3]);
    for (int p = 0; p < 2; ++p) {
        dft4(u[p]);
        for (int s = 0; s < 4; ++s) x[p + 2 * s] = u[p][s];
    }
}

static void dft16_rest(cpx* x) {             /* after l1 on pairs (m, m + 8) */
    cpx u[4][4];   /* u[p][m0] */
    const cpx w1 = {K_A, -K_C}, w3 = {K_C, -K_A}, w9 = {-K_A, K_C};
    for (int m0 = 0; m0 < 4; ++m0) {
        /* radix-4 across (m0, m0+4, m0+8, m0+12): level 1 is done, pairs (m0, m0+8) and (m0+4, m0+12) */
        cpx t[4] = {x[m0], x[m0 + 4], x[m0 + 8], x[m0 + 12]};
        dft4_rest(&t[0], &t[1], &t[2], &t[3]);
        for (int p = 0; p < 4; ++p) u[p][m0] = t[p];
    }
    /* internal twiddles W16^(m0*p) */
    u[1][1] = cmul(u[1][1], w1);  u[1][2] = mul_w8(u[1][2]);  u[1][3] = cmul(u[1][3], w3);
    u[2][1] = mul_w8(u[2][1]);    u[2][2] = mul_mi(u[2][2]);  u[2][3] = mul_w83(u[2][3]);
    u[3][1] = cmul(u[3][1], w3);  u[3][2] = mul_w83(u[3][2]); u[3][3] = cmul(u[3][3], w9);
    for (int p = 0; p < 4; ++p) {
        dft4(u[p]);
        for (int s = 0; s < 4; ++s) x[p + 4 * s] = u[p][s];
    }
}

/* 32 = 4 x 8: radix-4 across m1 (m = m0 + 8 m1), twiddle W32^(m0 p), radix-8 across m0; q = p + 4 s.
 * W32^e is applied as W32^(e mod 8) (unit32 constants) followed by (e div 8) exact quarter turns. */
static void dft32_rest(cpx* x) {             /* after l1 on pairs (m, m + 16) */
    cpx u[4][8];   /* u[p][m0] */
    for (int m0 = 0; m0 < 8; ++m0) {
        cpx t[4] = {x[m0], x[m0 + 8], x[m0 + 16], x[m0 + 24]};
        dft4_rest(&t[0], &t[1], &t[2], &t[3]);
        for (int p = 0; p < 4; ++p) u[p][m0] = t[p];
    }
    for (int p = 1; p < 4; ++p)
        for (int m0 = 1; m0 < 8; ++m0) {
            /* W32^e = (-i)^(e div 8) W32^(e mod 8): a complex multiply by the first-quadrant constant
             * (skipped when e mod 8 == 0) followed by exact quarter turns. */
            int e = m0 * p, r = e & 7, k = e >> 3;
            cpx v = u[p][m0];
            if (r) {
                float c, s;
                unit32(r, &c, &s);
                cpx w = {c, -s};
                v = cmul(v, w);
            }
            if (k == 1) v = mul_mi(v);
            else if (k == 2) { v.re = -v.re; v.im = -v.im; }
            u[p][m0] = v;
        }
    for (int p = 0; p < 4; ++p) {
        l1(u[p], 8);
        dft8_rest(u[p]);
        for (int s = 0; s < 8; ++s) x[p + 4 * s] = u[p][s];
    }
}

static void dft_rest(cpx* x, int r) {
    if (r == 32) dft32_rest(x);
    else if (r == 16) dft16_rest(x);
    else if (r == 8) dft8_rest(x);
    else if (r == 4) dft4_rest(&x[0], &x[1], &x[2], &x[3]);
    /* r == 2: the first level is the whole butterfly */
}

/* Twiddles of a "chain" pass: output q = 4a + b of a radix-r butterfly is multiplied
 * by W^(j q) as (x * A[a]) * B[b] with B[1] = w1 = W^j, B[2] = w1*w1, B[3] = B[2]*w1, A[1] = B[2]*B[2],
 * A[2] = A[1]*A[1], A[3] = A[2]*A[1] -- six live twiddles instead of r - 1. */
static void tw_two_level(cpx* x, int r, cpx w1) {
    cpx B[4], A[4];
    B[1] = w1;
    if (r == 2) { x[1] = cmul(x[1], B[1]); return; }
    B[2] = cmul(w1, w1);
    B[3] = cmul(B[2], w1);
    if (r > 4) A[1] = cmul(B[2], B[2]);
    if (r > 8) { A[2] = cmul(A[1], A[1]); A[3] = cmul(A[2], A[1]); }
    for (int q = 1; q < r; ++q) {
        int a = q >> 2, b = q & 3;
        if (a) x[q] = cmul(x[q], A[a]);
        if (b) x[q] = cmul(x[q], B[b]);
    }
}

#define SO_TABLE_PASS_MAX 1024   /* a pass uses exact table twiddles iff (L/r)*(r-1) <= this */

/* In-place mixed-radix DIF FFT on d[0..N); output in digit-reversed order. */
static void fft_dif(cpx* d, int N, const int* radices, int np, const float* tab, const float* win) {
    int window = (win != NULL);
    int L = N;
    for (int p = 0; p < np; ++p) {
        int r = radices[p], M = L / r, step = N / L;
        /* table twiddles for the small passes and (round 2) for the radix-16 pass of the 16384-point plan, whose 30
         * per-thread twiddles the kernel keeps in tensor memory; the remaining chain passes: 4096 / 8192 first pass,
         * large-N front pass */
        int use_table = (M * (r - 1) <= SO_TABLE_PASS_MAX) || (L == 16384 && r == 16);
        for (int base = 0; base < N; base += L) {
            for (int j = 0; j < M; ++j) {
                cpx x[32], w[32];
                for (int m = 0; m < r; ++m) x[m] = d[base + j + m * M];
                if (p == 0 && window) {
                    /* Hann values of the first r/2 samples j + m M (all < N/2) from the window table */
                    float wv[16];
                    for (int m = 0; m < r / 2; ++m) wv[m] = win[j + m * M];
                    l1_window(x, r, wv);
                } else {
                    l1(x, r);
                }
                dft_rest(x, r);
                if (M > 1) {
                    if (use_table) {
                        for (int q = 1; q < r; ++q) {
                            int k = j * q * step;
                            w[q].re = tab[2 * k]; w[q].im = tab[2 * k + 1];
                        }
                        for (int q = 1; q < r; ++q) x[q] = cmul(x[q], w[q]);
                    } else {
                        int k = j * step;
                        cpx w1 = {tab[2 * k], tab[2 * k + 1]};
                        if (r > 16) return;                       /* the plan keeps radix-32 passes on tables */
                        tw_two_level(x, r, w1);
                    }
                }
                for (int q = 0; q < r; ++q) d[base + j + q * M] = x[q];
            }
        }
        L = M;
    }
}

/* position -> bin (mixed-radix digit reversal) */
static int pos_to_bin(int pos, int N, const int* radices, int np) {
    int M = N, k = 0, mult = 1;
    for (int p = 0; p < np; ++p) {
        M /= radices[p];
        int q = pos / M; pos -= q * M;
        k += q * mult; mult *= radices[p];
    }
    return k;
}

static inline uint8_t quantise(float P, const float* T) {
    int lo = 0, hi = 255;            /* largest k with P >= T[k]; T[0] = 0, T increasing */
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (P >= T[mid]) lo = mid; else hi = mid - 1;
    }
    return (uint8_t)lo;
}

/* One frame: interleaved float32 IQ[N] -> Kiwi bytes[N], fftshifted (bin 0 = lowest frequency).
 * If spec_out != NULL also returns the raw FFT (natural order) for diagnostics. */
int so_wf_frame_bytes(const float* iq, int N, int window, double cal_db, uint8_t* bytes, float* spec_out) {
    int radices[SO_MAX_PASSES];
    int np = so_fft_plan(N, radices);
    if (np < 0) return -1;
    float* tab = (float*)malloc(sizeof(float) * 2 * (size_t)N);
    cpx* d = (cpx*)malloc(sizeof(cpx) * (size_t)N);
    float* win = (float*)malloc(sizeof(float) * (size_t)(N / 2));
    float T[256];
    so_twiddle_table(N, tab);
    so_window_table(N, win);
    so_thresholds(N, cal_db, T);
    memcpy(d, iq, sizeof(cpx) * (size_t)N);
    fft_dif(d, N, radices, np, tab, window ? win : NULL);
    for (int pos = 0; pos < N; ++pos) {
        int k = pos_to_bin(pos, N, radices, np);
        float t = d[pos].im * d[pos].im;
        float P = fmaf(d[pos].re, d[pos].re, t);
        bytes[(k + N / 2) & (N - 1)] = quantise(P, T);
        if (spec_out) { spec_out[2 * k] = d[pos].re; spec_out[2 * k + 1] = d[pos].im; }
    }
    free(d); free(tab); free(win);
    return 0;
}

/* The kernel divides by loop-invariant divisors with Markstein's sequence (r = RN(1/b); q0 = RN(a r);
 * e = fma(-q0, b, a); q = fma(e, r, q0)); this restates it so that tests can compare it with IEEE division.
 * Returns the number of (a[i], b) pairs whose result differs from a[i] / b. */
int so_markstein_mismatches(const float* a, int n, float b) {
    volatile float r = 1.0f / b;
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        float q0 = a[i] * r;
        float e = fmaf(-q0, b, a[i]);
        float q = fmaf(e, r, q0);
        float ref = a[i] / b;
        if (memcmp(&q, &ref, sizeof(float)) != 0) ++bad;
    }
    return bad;
}

/* ---- Tier-P tail in strict float32 (restates oracle/tier_p.py; utils_supersdr.py:787-813) --- */
typedef struct {
    int zoom, auto_scale, delta_low_db, delta_high_db;
    int p_lo;            /* percentile lower index and lerp weight, computed by the caller with */
    float p_gamma;       /* numpy's own float32 expression (tier_p.percentile_virtual_index)     */
    float low_clip_db, dynamic_range;   /* in: values kept when auto_scale == 0; out: updated  */
    float high_clip_db, wf_min_db, wf_max_db;
} so_colour_t;

static int cmp_u16(const void* a, const void* b) {
    return (int)*(const uint16_t*)a - (int)*(const uint16_t*)b;
}

/* sums: per-bin integer sum of n byte lines. */
void so_colour_row(const uint16_t* sums, int W, int n, so_colour_t* st, float* spectrum,
                   float* colour, uint8_t* pixels) {
    float fn = (float)n, z3 = (float)(3 * st->zoom);
    uint16_t* srt = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)W);
#define WFDB(s) ((((float)(s) / fn - 255.0f) - 13.0f) + z3)
    for (int i = 0; i < W; ++i) {
        if (spectrum) spectrum[i] = (float)sums[i] / fn;
        srt[i] = sums[i ? i : 1];                        /* wf_db[0] = wf_db[1] */
    }
    if (st->auto_scale) {
        qsort(srt, (size_t)W, sizeof(uint16_t), cmp_u16);
        int lo = st->p_lo, hi = lo + 1 < W ? lo + 1 : W - 1;
        float a = WFDB(srt[lo]), b = WFDB(srt[hi]), g = st->p_gamma, d = b - a, p;
        if (g >= 0.5f) { float t = 1.0f - g; t = d * t; p = b - t; }
        else { float t = d * g; p = a + t; }
        st->low_clip_db = p;
        st->high_clip_db = WFDB(srt[W - 1]);
        float dyn = st->high_clip_db - st->low_clip_db;
        st->dynamic_range = dyn > 40.0f ? dyn : 40.0f;
    }
    float low = st->low_clip_db + (float)st->delta_low_db;
    float nf = st->dynamic_range + (float)st->delta_high_db;
    float den = nf - (float)st->delta_low_db;
    st->wf_min_db = low - z3;
    st->wf_max_db = (st->low_clip_db + nf) - z3;
    for (int i = 0; i < W; ++i) {
        float w = WFDB(sums[i ? i : 1]);
        float c = (w - low) / den;
        c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);     /* NaN (den == 0) propagates like np.clip */
        c = c * 254.0f;
        c = c < 0.0f ? 0.0f : (c > 255.0f ? 255.0f : c);
        if (colour) colour[i] = c;
        if (pixels) pixels[i] = (uint8_t)rintf(c);
    }
#undef WFDB
    free(srt);
}

/* Full chain for a batch: iq[B][n][N] float32 interleaved -> pixels[B][N] (+ optional outputs).
 * Channels are independent; callers parallelise by calling this on channel slices from several
 * threads (ctypes releases the GIL). */
int so_wf_rows(const float* iq, int B, int n, int N, int window, double cal_db, const so_colour_t* st_in,
               uint8_t* pixels, float* colour, float* spectrum, uint16_t* sums_out, float* scalars) {
    int radices[SO_MAX_PASSES];
    if (so_fft_plan(N, radices) < 0) return -1;
    int rc = 0;
    for (int b = 0; b < B; ++b) {
        uint16_t* sums = (uint16_t*)calloc((size_t)N, sizeof(uint16_t));
        uint8_t* line = (uint8_t*)malloc((size_t)N);
        for (int f = 0; f < n; ++f) {
            if (so_wf_frame_bytes(iq + 2 * ((size_t)b * n + f) * N, N, window, cal_db, line, NULL)) rc = -1;
            for (int i = 0; i < N; ++i) sums[i] = (uint16_t)(sums[i] + line[i]);
        }
        so_colour_t st = *st_in;
        so_colour_row(sums, N, n, &st, spectrum ? spectrum + (size_t)b * N : NULL,
                      colour ? colour + (size_t)b * N : NULL, pixels ? pixels + (size_t)b * N : NULL);
        if (sums_out) memcpy(sums_out + (size_t)b * N, sums, sizeof(uint16_t) * (size_t)N);
        if (scalars) {
            float* s = scalars + 5 * (size_t)b;
            s[0] = st.low_clip_db; s[1] = st.high_clip_db; s[2] = st.dynamic_range;
            s[3] = st.wf_min_db; s[4] = st.wf_max_db;
        }
        free(sums); free(line);
    }
    return rc;
}

/* ---- Tier-P audio interpolator (utils_supersdr.py:1121-1138) in float64 ------------------- */
/* x: int16[n] at 12 kHz; hist: float64[n_tap-1] (updated); h: float64[n_tap]; out: int16[4n][2] */
void so_play_buffer(const int16_t* x, int n, double volume, double balance, const double* h,
                    int n_tap, int ratio, double* hist, int16_t* out, double* mono_out) {
    int L = ratio * n, H = n_tap - 1;
    double* z = (double*)calloc((size_t)(H + L), sizeof(double));
    memcpy(z, hist, sizeof(double) * (size_t)H);
    for (int k = 0; k < n; ++k) z[H + ratio * k] = (double)x[k] * (volume / 100.0);
    memcpy(hist, z + L, sizeof(double) * (size_t)H);
    double lv = 1.0 - balance < 1.0 ? 1.0 - balance : 1.0;
    double rv = 1.0 + balance < 1.0 ? 1.0 + balance : 1.0;
    lv *= lv; rv *= rv;
    for (int j = 0; j < L; ++j) {
        double acc = 0.0;
        for (int t = 0; t < n_tap; ++t) acc += h[t] * z[j + H - t];
        acc *= (double)ratio;
        if (mono_out) mono_out[j] = acc;
        out[2 * j] = (int16_t)(int32_t)(acc * lv);
        out[2 * j + 1] = (int16_t)(int32_t)(acc * rv);
    }
    free(z);
}
