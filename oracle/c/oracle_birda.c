/*
 * oracle_birda.c — plain-C, f32 restatement of the reference's CPU front end and post step.
 *
 * TEST INFRASTRUCTURE ONLY (checker for tests/ and the cpu_baseline / --impl reference legs
 * of bench.py).  Nothing under birda_b200/ links or calls this.  Independent of the numpy
 * oracle (own FFT, f32 throughout as the reference computes) so the two cross-check.
 *
 * Restates (paths relative to /root/reference):
 *   append_samples            src/audio/decode.rs:353-411
 *   next_segment              src/audio/decode.rs:150-202
 *   decode_and_stream         src/pipeline/processor.rs:49-108
 *   resample / rubato::Fft    src/audio/resample.rs:10-105 (+ rubato 4.0.0, third-party, not
 *                             vendored: algorithm restated, SAMPLE-LEVEL PARITY UNPINNED)
 *   top-k / threshold         birdnet-onnx 2.0.0-rc.16 (third-party, PARITY UNPINNED)
 *   filter_predictions        src/inference/geomodel_filter.rs:45-79
 *   second threshold          src/pipeline/processor.rs:374
 * The resampler (filter, twiddles, plans) is rebuilt for EVERY window, as the reference does
 * (src/audio/resample.rs:19 is inside the per-chunk call).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float r, i; } cpx;

/* ------------------------------------------------------------------ mixed-radix complex FFT */
typedef struct {
    int n, nfac, fac[64];
    cpx* tw;            /* tw[k] = exp(sign * 2 pi i k / n), built in double, stored f32 (as rustfft) */
} fft_plan;

static int fft_plan_init(fft_plan* p, int n, int inverse) {
    p->n = n; p->nfac = 0;
    int m = n;
    while (m % 4 == 0) { p->fac[p->nfac++] = 4; m /= 4; }
    while (m % 2 == 0) { p->fac[p->nfac++] = 2; m /= 2; }
    for (int f = 3; f * f <= m; f += 2) while (m % f == 0) { p->fac[p->nfac++] = f; m /= f; }
    if (m > 1) p->fac[p->nfac++] = m;
    for (int f = 0; f < p->nfac; ++f) if (p->fac[f] > 63) return -1;   /* generic butterfly scratch limit */
    p->tw = (cpx*)malloc(sizeof(cpx) * (size_t)n);
    if (!p->tw) return -1;
    const double s = inverse ? 2.0 : -2.0;
    for (int k = 0; k < n; ++k) {
        double a = s * 3.14159265358979323846 * (double)k / (double)n;
        p->tw[k].r = (float)cos(a); p->tw[k].i = (float)sin(a);
    }
    return 0;
}
static void fft_plan_free(fft_plan* p) { free(p->tw); p->tw = NULL; }

static inline cpx cmul(cpx a, cpx b) { cpx c = { a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r }; return c; }
static inline cpx cadd(cpx a, cpx b) { cpx c = { a.r + b.r, a.i + b.i }; return c; }
static inline cpx csub(cpx a, cpx b) { cpx c = { a.r - b.r, a.i - b.i }; return c; }

/* decimation in time: out[0..n) = DFT(in[0], in[stride], ...); tws = N / n */
static void fft_rec(const fft_plan* P, cpx* out, const cpx* in, int n, int stride, int tws, int level, cpx* scratch) {
    if (n == 1) { out[0] = in[0]; return; }
    const int p = P->fac[level], m = n / p, N = P->n;
    for (int q = 0; q < p; ++q) fft_rec(P, out + (size_t)q * m, in + (size_t)q * stride, m, stride * p, tws * p, level + 1, scratch);
    const cpx* tw = P->tw;
    if (p == 2) {
        for (int k = 0; k < m; ++k) {
            cpx t = cmul(out[m + k], tw[(size_t)k * tws]);
            cpx a = out[k];
            out[k] = cadd(a, t); out[m + k] = csub(a, t);
        }
    } else if (p == 4) {
        const int inv = tw[N / 4 % N].i > 0;   /* sign of the transform */
        for (int k = 0; k < m; ++k) {
            cpx a0 = out[k];
            cpx a1 = cmul(out[m + k], tw[(size_t)k * tws]);
            cpx a2 = cmul(out[2 * m + k], tw[(size_t)2 * k * tws]);
            cpx a3 = cmul(out[3 * m + k], tw[(size_t)3 * k * tws]);
            cpx t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), d = csub(a1, a3);
            cpx t3 = inv ? (cpx){ -d.i, d.r } : (cpx){ d.i, -d.r };
            out[k] = cadd(t0, t2); out[2 * m + k] = csub(t0, t2);
            out[m + k] = cadd(t1, t3); out[3 * m + k] = csub(t1, t3);
        }
    } else {
        /* generic radix p: t[q] = out[q m + k] * W_n^(q k); out[k + r m] = sum_q t[q] W_p^(q r) */
        cpx* t = scratch + (size_t)level * 64;
        const int wp = N / p;
        for (int k = 0; k < m; ++k) {
            for (int q = 0; q < p; ++q) t[q] = q ? cmul(out[(size_t)q * m + k], tw[((size_t)q * k * tws) % N]) : out[k];
            for (int r = 0; r < p; ++r) {
                cpx acc = t[0];
                int idx = 0;
                for (int q = 1; q < p; ++q) {
                    idx += r * wp; if (idx >= N) idx %= N;
                    acc = cadd(acc, cmul(t[q], tw[idx]));
                }
                out[k + (size_t)r * m] = acc;
            }
        }
    }
}
static void fft_exec(const fft_plan* P, cpx* out, const cpx* in, cpx* scratch) { fft_rec(P, out, in, P->n, 1, 1, 0, scratch); }

/* ------------------------------------------------------------------ rubato-style resampler */
typedef struct {
    uint32_t from, to; int n_in, n_out, n_keep;
    float cutoff;
    float* taps;          /* [n_in] */
    cpx* filt;            /* [n_in + 1] */
    cpx *wf, *wi;         /* split twiddles exp(-i pi k/n_in) [n_in+1], exp(+i pi k/n_out) [n_out+1] */
    fft_plan fwd, inv;
    cpx *z, *Z, *Y, *zp, *scratch;
    float *ybuf, *carry;
} resampler;

static uint32_t gcd_u32(uint32_t a, uint32_t b) { while (b) { uint32_t t = a % b; a = b; b = t; } return a; }

static void real_fft_2n(resampler* R, const float* x /* [n_in], zero padded to 2 n_in implicitly */, cpx* X /* [n_in+1] */) {
    const int N = R->n_in;
    for (int n = 0; n < N; ++n) {
        float re = (2 * n < N) ? x[2 * n] : 0.0f, im = (2 * n + 1 < N) ? x[2 * n + 1] : 0.0f;
        R->z[n].r = re; R->z[n].i = im;
    }
    fft_exec(&R->fwd, R->Z, R->z, R->scratch);
    for (int k = 0; k <= N; ++k) {
        cpx zk = R->Z[k == N ? 0 : k], zn = R->Z[k == 0 ? 0 : N - k];
        zn.i = -zn.i;
        cpx fe = { 0.5f * (zk.r + zn.r), 0.5f * (zk.i + zn.i) };
        cpx d  = { 0.5f * (zk.r - zn.r), 0.5f * (zk.i - zn.i) };
        cpx fo = { d.i, -d.r };
        X[k] = cadd(fe, cmul(R->wf[k], fo));
    }
}

static void real_ifft_2m(resampler* R, const cpx* Y /* [n_out+1] */, float* y /* [2 n_out] */) {
    const int M = R->n_out;
    for (int k = 0; k < M; ++k) {
        cpx a = Y[k], b = Y[M - k];
        if (k == 0) { a.i = 0.0f; b.i = 0.0f; }          /* realfft ignores imag of DC / Nyquist */
        b.i = -b.i;
        cpx e = cadd(a, b), o = cmul(R->wi[k], csub(a, b));
        R->zp[k].r = e.r - o.i; R->zp[k].i = e.i + o.r;
    }
    fft_exec(&R->inv, (cpx*)y, R->zp, R->scratch);       /* (y[2n], y[2n+1]) = z'[n] */
}

static int resampler_init(resampler* R, uint32_t from, uint32_t to) {
    memset(R, 0, sizeof(*R));
    R->from = from; R->to = to;
    uint32_t g = gcd_u32(from, to), min_in = from / g;
    uint32_t k = (1024 + min_in - 1) / min_in;
    R->n_in = (int)(k * (from / g)); R->n_out = (int)(k * (to / g));
    R->n_keep = R->n_in < R->n_out ? R->n_in + 1 : R->n_out;
    const int N = R->n_in, M = R->n_out;
    float cutoff = powf(0.4f, 16.0f / (float)N);
    if (N > M) cutoff = cutoff * (float)M / (float)N;
    R->cutoff = cutoff;
    R->taps = (float*)calloc((size_t)2 * N, sizeof(float));
    R->filt = (cpx*)malloc(sizeof(cpx) * (size_t)(N + 1));
    R->wf = (cpx*)malloc(sizeof(cpx) * (size_t)(N + 1));
    R->wi = (cpx*)malloc(sizeof(cpx) * (size_t)(M + 1));
    const int L = N > M ? N : M;
    R->z = (cpx*)malloc(sizeof(cpx) * (size_t)(L + 1)); R->Z = (cpx*)malloc(sizeof(cpx) * (size_t)(L + 1));
    R->Y = (cpx*)calloc((size_t)(M + 1), sizeof(cpx)); R->zp = (cpx*)malloc(sizeof(cpx) * (size_t)(L + 1));
    R->scratch = (cpx*)malloc(sizeof(cpx) * 64 * 64);
    R->ybuf = (float*)malloc(sizeof(float) * (size_t)2 * M); R->carry = (float*)calloc((size_t)M, sizeof(float));
    if (fft_plan_init(&R->fwd, N, 0) || fft_plan_init(&R->inv, M, 1)) return -1;
    for (int q = 0; q <= N; ++q) { double a = -3.14159265358979323846 * q / N; R->wf[q].r = (float)cos(a); R->wf[q].i = (float)sin(a); }
    for (int q = 0; q <= M; ++q) { double a = 3.14159265358979323846 * q / M; R->wi[q].r = (float)cos(a); R->wi[q].i = (float)sin(a); }
    /* windowed sinc in f32: BlackmanHarris^2 (periodic) * sinc((x - N/2) cutoff), unit sum, / 2N */
    const float pi = 3.14159265358979323846f, npf = (float)N;
    float sum = 0.0f;
    for (int x = 0; x < N; ++x) {
        float xf = (float)x;
        float w = 0.35875f - 0.48829f * cosf(2.0f * pi * xf / npf) + 0.14128f * cosf(4.0f * pi * xf / npf) - 0.01168f * cosf(6.0f * pi * xf / npf);
        float arg = (xf - (float)(N / 2)) * cutoff;
        float s = arg == 0.0f ? 1.0f : sinf(arg * pi) / (arg * pi);
        float y = w * w * s;
        R->taps[x] = y; sum += y;
    }
    for (int x = 0; x < N; ++x) R->taps[x] = R->taps[x] / sum / (float)(2 * N);
    real_fft_2n(R, R->taps, R->filt);                    /* filter spectrum through the same f32 FFT */
    return 0;
}

static void resampler_free(resampler* R) {
    free(R->taps); free(R->filt); free(R->wf); free(R->wi); free(R->z); free(R->Z); free(R->Y); free(R->zp);
    free(R->scratch); free(R->ybuf); free(R->carry); fft_plan_free(&R->fwd); fft_plan_free(&R->inv);
}

/* resample(): src/audio/resample.rs:10-105.  Returns the number of samples written to out. */
static size_t resample_window(const float* x, size_t len, uint32_t from, uint32_t to, float* out, size_t out_cap) {
    resampler R;
    if (resampler_init(&R, from, to)) return 0;
    const int N = R.n_in, M = R.n_out;
    cpx* X = (cpx*)malloc(sizeof(cpx) * (size_t)(N + 1));
    float* blk = (float*)malloc(sizeof(float) * (size_t)N);
    size_t pos = 0, nout = 0;
    int more = 1;
    while (more) {
        size_t want_out;
        const float* src;
        if (pos + (size_t)N <= len) { src = x + pos; want_out = (size_t)M; pos += (size_t)N; }
        else if (pos < len) {
            size_t rem = len - pos;
            memset(blk, 0, sizeof(float) * (size_t)N); memcpy(blk, x + pos, sizeof(float) * rem);
            src = blk;
            size_t of = (size_t)ceil((double)rem * (double)to / (double)from);
            want_out = of < (size_t)M ? of : (size_t)M;
            pos = len; more = 0;
        } else break;
        real_fft_2n(&R, src, X);
        memset(R.Y, 0, sizeof(cpx) * (size_t)(M + 1));
        for (int k = 0; k < R.n_keep; ++k) R.Y[k] = cmul(X[k], R.filt[k]);
        real_ifft_2m(&R, R.Y, R.ybuf);
        for (size_t j = 0; j < want_out && nout < out_cap; ++j, ++nout) out[nout] = R.ybuf[j] + R.carry[j];
        memcpy(R.carry, R.ybuf + M, sizeof(float) * (size_t)M);
        if (pos >= len) more = 0;
    }
    free(X); free(blk); resampler_free(&R);
    return nout;
}

/* ------------------------------------------------------------------ front end over a file */
enum { ORC_S16 = 1, ORC_S32 = 2, ORC_F32 = 3 };

static float mono_at(const void* pcm, int fmt, uint32_t ch, uint64_t frame) {
    float sum = 0.0f;
    if (fmt == ORC_S16) {
        const int16_t* p = (const int16_t*)pcm + frame * ch;
        if (ch == 1) return (float)p[0] / 32768.0f;
        for (uint32_t c = 0; c < ch; ++c) sum += (float)p[c] / 32768.0f;
    } else if (fmt == ORC_S32) {
        const int32_t* p = (const int32_t*)pcm + frame * ch;
        if (ch == 1) return (float)p[0] / 2147483648.0f;
        for (uint32_t c = 0; c < ch; ++c) sum += (float)p[c] / 2147483648.0f;
    } else {
        const float* p = (const float*)pcm + frame * ch;
        if (ch == 1) return p[0];
        for (uint32_t c = 0; c < ch; ++c) sum += p[c];
    }
    return sum / (float)ch;
}

typedef struct {
    const void* pcm; int fmt; uint32_t ch; uint64_t frames; uint32_t sr, tr; uint64_t seg, src_seg;
    const uint64_t* start; const uint64_t* take; uint64_t nseg; float* out;
    uint64_t next; pthread_mutex_t mu;
} fe_job;

static void* fe_worker(void* arg) {
    fe_job* J = (fe_job*)arg;
    float* raw = (float*)malloc(sizeof(float) * J->src_seg);
    float* res = (float*)malloc(sizeof(float) * (J->seg + 65536));
    for (;;) {
        pthread_mutex_lock(&J->mu);
        uint64_t i = J->next++;
        pthread_mutex_unlock(&J->mu);
        if (i >= J->nseg) break;
        /* next_segment: copy `take` samples, zero-pad to src_seg (decode.rs:175-181) */
        for (uint64_t j = 0; j < J->take[i]; ++j) raw[j] = mono_at(J->pcm, J->fmt, J->ch, J->start[i] + j);
        for (uint64_t j = J->take[i]; j < J->src_seg; ++j) raw[j] = 0.0f;
        float* dst = J->out + i * J->seg;
        if (J->sr == J->tr) {
            memcpy(dst, raw, sizeof(float) * (J->seg < J->src_seg ? J->seg : J->src_seg));
        } else {
            size_t n = resample_window(raw, J->src_seg, J->sr, J->tr, res, J->seg + 65536);
            size_t c = n < J->seg ? n : J->seg;          /* samples.resize(seg, 0.0) (processor.rs:87) */
            memcpy(dst, res, sizeof(float) * c);
            memset(dst + c, 0, sizeof(float) * (J->seg - c));
        }
    }
    free(raw); free(res);
    return NULL;
}

/* Returns nseg (or -1).  out may be NULL to only count; tables may be NULL. */
int64_t orc_frontend(const void* pcm, int fmt, uint32_t channels, uint64_t frames, uint32_t sr, uint32_t tr,
                     uint64_t seg, uint64_t ovl, float* out, uint64_t cap_rows,
                     uint64_t* start_sample, float* start_time, float* end_time, int threads) {
    uint64_t src_seg = seg, src_ovl = ovl;
    if (sr != tr) {
        src_seg = (uint64_t)ceil((double)seg * (double)sr / (double)tr);
        src_ovl = (uint64_t)ceil((double)ovl * (double)sr / (double)tr);
    }
    if (src_ovl >= src_seg) return -2;
    /* literal next_segment bookkeeping over a fully buffered stream (decode.rs:150-202) */
    uint64_t cap = 16, n = 0;
    uint64_t* st = (uint64_t*)malloc(sizeof(uint64_t) * cap); uint64_t* tk = (uint64_t*)malloc(sizeof(uint64_t) * cap);
    uint64_t buffered = frames, emitted = 0;
    while (buffered > 0) {
        uint64_t take = buffered < src_seg ? buffered : src_seg;
        if (n == cap) { cap *= 2; st = (uint64_t*)realloc(st, sizeof(uint64_t) * cap); tk = (uint64_t*)realloc(tk, sizeof(uint64_t) * cap); }
        st[n] = emitted; tk[n] = take; ++n;
        uint64_t adv = take > src_ovl ? take - src_ovl : 0;
        if (adv > 0) { buffered -= adv; emitted += adv; } else { emitted += take; buffered = 0; }
    }
    for (uint64_t i = 0; i < n && i < cap_rows; ++i) {
        if (start_sample) start_sample[i] = st[i];
        float s = (float)st[i] / (float)sr, d = (float)seg / (float)tr;
        if (start_time) start_time[i] = s;
        if (end_time) end_time[i] = s + d;
    }
    if (out) {
        if (cap_rows < n) { free(st); free(tk); return -9; }
        fe_job J; memset(&J, 0, sizeof(J));
        J.pcm = pcm; J.fmt = fmt; J.ch = channels; J.frames = frames; J.sr = sr; J.tr = tr; J.seg = seg; J.src_seg = src_seg;
        J.start = st; J.take = tk; J.nseg = n; J.out = out; J.next = 0;
        pthread_mutex_init(&J.mu, NULL);
        if (threads < 1) threads = 1;
        if (threads > 256) threads = 256;
        pthread_t th[256];
        for (int t = 1; t < threads; ++t) pthread_create(&th[t], NULL, fe_worker, &J);
        fe_worker(&J);
        for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
        pthread_mutex_destroy(&J.mu);
    }
    free(st); free(tk);
    return (int64_t)n;
}

/* single window resample for cross-checks */
int64_t orc_resample(const float* x, uint64_t len, uint32_t from, uint32_t to, float* out, uint64_t cap) {
    if (from == to) { uint64_t c = len < cap ? len : cap; memcpy(out, x, sizeof(float) * c); return (int64_t)c; }
    return (int64_t)resample_window(x, len, from, to, out, cap);
}

int32_t orc_resampler_info(uint32_t from, uint32_t to, int32_t* n_in, int32_t* n_out, float* cutoff, float* taps, int32_t ntaps) {
    resampler R;
    if (resampler_init(&R, from, to)) return -1;
    *n_in = R.n_in; *n_out = R.n_out; *cutoff = R.cutoff;
    if (taps) memcpy(taps, R.taps, sizeof(float) * (size_t)(ntaps < R.n_in ? ntaps : R.n_in));
    resampler_free(&R);
    return 0;
}

/* ------------------------------------------------------------------ post step */
typedef struct {
    const float* scores; uint32_t C, valid; int act; float min_conf; uint32_t topk;
    const float* mask; const uint8_t* keep; float thr; int keep_unmatched, rerank;
    uint32_t* idx; float* conf; uint32_t* cnt; uint32_t next; pthread_mutex_t mu;
} post_job;

static void post_row(const post_job* J, uint32_t r, float* tmp) {
    const float* x = J->scores + (size_t)r * J->C;
    const uint32_t C = J->C, K = J->topk;
    if (J->act == 1) for (uint32_t i = 0; i < C; ++i) tmp[i] = 1.0f / (1.0f + expf(-x[i]));
    else if (J->act == 2) {
        float m = x[0]; for (uint32_t i = 1; i < C; ++i) if (x[i] > m) m = x[i];
        double s = 0; for (uint32_t i = 0; i < C; ++i) { tmp[i] = expf(x[i] - m); s += tmp[i]; }
        for (uint32_t i = 0; i < C; ++i) tmp[i] = (float)(tmp[i] / s);
    } else memcpy(tmp, x, sizeof(float) * C);
    uint32_t bi[16]; float bc[16]; uint32_t n = 0;
    for (uint32_t i = 0; i < C; ++i) {               /* conf desc, ties to the lower index, top-k */
        float c = tmp[i];
        if (!(c >= J->min_conf)) continue;
        if (n == K && !(c > bc[K - 1])) continue;
        uint32_t p = n < K ? n : K - 1;
        while (p > 0 && c > bc[p - 1]) { if (p < K) { bc[p] = bc[p - 1]; bi[p] = bi[p - 1]; } --p; }
        bc[p] = c; bi[p] = i; if (n < K) ++n;
    }
    uint32_t oi[16]; float oc[16]; uint32_t m = 0;
    for (uint32_t a = 0; a < n; ++a) {               /* geomodel_filter.rs:54-71 / classifier.rs:616-641 */
        float c = bc[a];
        if (J->mask) {
            float s = J->mask[bi[a]];
            if (isnan(s)) { if (!(J->keep_unmatched && !J->rerank)) continue; }
            else if (s >= J->thr) { if (J->rerank) c = c * s; }
            else continue;
        } else if (J->keep) { if (!J->keep[bi[a]]) continue; }
        oi[m] = bi[a]; oc[m] = c; ++m;
    }
    if (J->mask && J->rerank)
        for (uint32_t a = 1; a < m; ++a) { uint32_t ti = oi[a]; float tc = oc[a]; int b = (int)a - 1;
            while (b >= 0 && oc[b] < tc) { oc[b + 1] = oc[b]; oi[b + 1] = oi[b]; --b; } oc[b + 1] = tc; oi[b + 1] = ti; }
    uint32_t w = 0;
    for (uint32_t a = 0; a < m; ++a) if (oc[a] >= J->min_conf) { J->idx[(size_t)r * K + w] = oi[a]; J->conf[(size_t)r * K + w] = oc[a]; ++w; }
    J->cnt[r] = w;
    for (; w < K; ++w) { J->idx[(size_t)r * K + w] = 0xFFFFFFFFu; J->conf[(size_t)r * K + w] = 0.0f; }
}

static void* post_worker(void* arg) {
    post_job* J = (post_job*)arg;
    float* tmp = (float*)malloc(sizeof(float) * J->C);
    for (;;) {
        pthread_mutex_lock(&J->mu); uint32_t r = J->next++; pthread_mutex_unlock(&J->mu);
        if (r >= J->valid) break;
        post_row(J, r, tmp);
    }
    free(tmp);
    return NULL;
}

int32_t orc_post(const float* scores, uint32_t C, uint32_t valid, int32_t act, float min_conf, uint32_t topk,
                 const float* mask, const uint8_t* keep, float thr, int32_t keep_unmatched, int32_t rerank,
                 uint32_t* idx, float* conf, uint32_t* cnt, int32_t threads) {
    if (topk < 1 || topk > 16) return -1;
    post_job J; memset(&J, 0, sizeof(J));
    J.scores = scores; J.C = C; J.valid = valid; J.act = act; J.min_conf = min_conf; J.topk = topk; J.mask = mask; J.keep = keep;
    J.thr = thr; J.keep_unmatched = keep_unmatched; J.rerank = rerank; J.idx = idx; J.conf = conf; J.cnt = cnt;
    pthread_mutex_init(&J.mu, NULL);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (int t = 1; t < threads; ++t) pthread_create(&th[t], NULL, post_worker, &J);
    post_worker(&J);
    for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
    pthread_mutex_destroy(&J.mu);
    return 0;
}
