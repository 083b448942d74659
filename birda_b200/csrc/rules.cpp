// Host rules (product code).  Pure arithmetic, no CUDA.  Compiled with strict IEEE float
// semantics (no -ffast-math): several of these are f32 computations in the reference and the
// results must be bit-identical to Rust's.
#include "rules.hpp"
#include <cmath>
#include <numeric>

namespace bb {

uint64_t trunc_f32_to_u64(float x) {           // Rust `f32 as usize`: saturating, NaN -> 0
    if (!(x > 0.0f)) return 0;
    if (x >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)x;
}

void segment_samples(float segment_duration, float overlap, uint32_t target_rate, bool bat,
                     uint64_t* seg, uint64_t* ovl) {
    if (bat) { *seg = 144000; *ovl = 144000 / 4; return; }        // constants.rs:531, processor.rs:506
    volatile float a = segment_duration * (float)target_rate;     // f32 product (processor.rs:514)
    volatile float b = overlap * (float)target_rate;              // (processor.rs:520)
    *seg = trunc_f32_to_u64(a);
    *ovl = trunc_f32_to_u64(b);
}

void source_window(uint64_t seg, uint64_t ovl, uint32_t sr, uint32_t tr, uint64_t* sseg, uint64_t* sovl) {
    if (sr == tr) { *sseg = seg; *sovl = ovl; return; }
    *sseg = (uint64_t)std::ceil((double)seg * (double)sr / (double)tr);
    *sovl = (uint64_t)std::ceil((double)ovl * (double)sr / (double)tr);
}

bool make_window_seq(uint64_t total, uint64_t seg, uint64_t ovl, bool full_only, WindowSeq* w) {
    if (ovl >= seg) return false;
    w->total = total; w->seg = seg; w->ovl = ovl; w->hop = seg - ovl;
    w->nseg = 0; w->last_start = 0;
    if (total == 0) return true;
    // full windows: start i*hop while total - i*hop >= seg
    uint64_t nfull = total >= seg ? (total - seg) / w->hop + 1 : 0;
    if (full_only) {
        w->nseg = nfull;
        w->last_start = nfull ? (nfull - 1) * w->hop : 0;
        return true;
    }
    uint64_t pos = nfull * w->hop;                 // buffer head after the full windows
    uint64_t rem = total - pos;                    // < seg
    uint64_t n = nfull;
    uint64_t last = nfull ? (nfull - 1) * w->hop : 0;
    if (rem > 0) {
        last = pos; ++n;                           // partial (or pure-overlap) window at i*hop
        if (rem > ovl) {                           // advance = rem - ovl > 0 leaves `ovl` samples
            if (ovl > 0) { last = pos + (rem - ovl); ++n; }
        }
    }
    w->nseg = n; w->last_start = last;
    return true;
}

void chunk_times(uint64_t start_sample, uint32_t sr, uint64_t seg, uint32_t tr, float* st, float* et) {
    volatile float s = (float)start_sample / (float)sr;           // usize as f32: RNE
    volatile float d = (float)seg / (float)tr;
    volatile float e = s + d;
    *st = s; *et = e;
}

int64_t estimate_segment_count(double duration, bool has_duration, float seg_dur, float overlap) {
    if (!has_duration) return -1;
    volatile float step = seg_dur - overlap;
    if (!(step > 0.0f)) return -1;
    double v = std::ceil(duration / (double)step);
    if (!(v > 0.0)) return 0;
    return (int64_t)v;
}

uint32_t effective_batch_size(uint32_t batch, int64_t est) {
    if (est <= 0) return batch;
    return (uint64_t)batch > (uint64_t)est ? (uint32_t)est : batch;
}

static const uint32_t kDaysInMonth[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};

uint32_t date_to_week(uint32_t month, uint32_t day) {
    uint32_t doy = day;
    for (uint32_t m = 0; m + 1 < month && m < 12; ++m) doy += kDaysInMonth[m];
    volatile float q = (float)(doy - 1) / 7.6f;
    uint32_t week = (uint32_t)std::floor(q) + 1;
    return week < 48 ? week : 48;
}

uint32_t week_to_start_day(uint32_t week) {
    return (uint32_t)std::fmaf((float)(week - 1), 7.6f, 1.0f);
}

void day_of_year_to_date(uint32_t doy, uint32_t* month, uint32_t* day) {
    uint32_t rem = doy;
    for (uint32_t m = 0; m < 12; ++m) {
        if (rem <= kDaysInMonth[m]) { *month = m + 1; *day = rem; return; }
        rem -= kDaysInMonth[m];
    }
    *month = 12; *day = 31;
}

bool factorize(uint32_t n, std::vector<int>* radices) {
    // odd radices first (their stride-r stores in the first Stockham stages are bank-conflict
    // free), powers of two last (by then the store runs are long and contiguous)
    radices->clear();
    std::vector<int> odd, even;
    for (int p : {7, 5, 3, 11, 13, 17, 19, 23, 29, 31})
        while (n % p == 0) { odd.push_back(p); n /= p; }
    while (n % 8 == 0) { even.push_back(8); n /= 8; }
    while (n % 4 == 0) { even.push_back(4); n /= 4; }
    while (n % 2 == 0) { even.push_back(2); n /= 2; }
    // what is left has only prime factors > 31 (rates like 12 345 Hz -> blocks of 2 * 823): they become stages of
    // their own, evaluated as direct O(p^2) DFTs by the fallback kernel (k2_resample.cu: run_stage_generic) — slow
    // next to the butterflies, but the reference resamples every pair rubato accepts (src/audio/resample.rs:19-28)
    for (uint32_t p = 37; n > 1 && (uint64_t)p * p <= n; p += 2)
        while (n % p == 0) { odd.push_back((int)p); n /= p; }
    if (n > 1) { if (n > 32768) return false; odd.push_back((int)n); n = 1; }
    // largest odd radix first
    for (size_t i = 0; i < odd.size(); ++i)
        for (size_t j = i + 1; j < odd.size(); ++j)
            if (odd[j] > odd[i]) std::swap(odd[i], odd[j]);
    *radices = odd;
    radices->insert(radices->end(), even.begin(), even.end());
    return true;
}

bool make_resampler_spec(uint32_t from, uint32_t to, bool want_spectrum, ResamplerSpec* s, std::string* err) {
    if (from == 0 || to == 0) { if (err) *err = "sample rate must be > 0"; return false; }
    s->from = from; s->to = to;
    uint32_t g = std::gcd(from, to);
    uint32_t min_in = from / g;
    uint64_t k = (1024 + (uint64_t)min_in - 1) / min_in;          // ceil(1024 / (from/g))
    uint64_t n_in = k * (from / g), n_out = k * (to / g);
    if (n_in > (1u << 20) || n_out > (1u << 20)) {
        if (err) *err = "resampler block too large for rates " + std::to_string(from) + " -> " + std::to_string(to);
        return false;
    }
    s->n_in = (uint32_t)n_in; s->n_out = (uint32_t)n_out;
    s->n_keep = n_in < n_out ? s->n_in + 1 : s->n_out;
    // cutoff = 0.4^(16/n_in) [* n_out/n_in when down-sampling], all in f32
    volatile float e = 16.0f / (float)s->n_in;
    volatile float c = std::pow(0.4f, (float)e);
    if (n_in > n_out) { volatile float t = c * (float)s->n_out; c = t / (float)s->n_in; }
    s->cutoff = c;
    // windowed sinc, f32: BlackmanHarris^2 (periodic) * sinc((x - n/2) * cutoff); unit sum; / 2n
    const uint32_t n = s->n_in;
    s->taps.assign(n, 0.f);
    const float pi = 3.14159265358979323846f;
    const float npf = (float)n;
    volatile float sum = 0.f;
    for (uint32_t x = 0; x < n; ++x) {
        float xf = (float)x;
        volatile float a1 = 2.0f * pi * xf / npf, a2 = 4.0f * pi * xf / npf, a3 = 6.0f * pi * xf / npf;
        volatile float t1 = 0.48829f * std::cos((float)a1);
        volatile float t2 = 0.14128f * std::cos((float)a2);
        volatile float t3 = 0.01168f * std::cos((float)a3);
        volatile float w = 0.35875f - t1; w = w + t2; w = w - t3;
        volatile float w2 = w * w;
        volatile float arg = (xf - (float)(n / 2)) * (float)s->cutoff;
        volatile float v;
        if (arg == 0.0f) v = 1.0f;
        else { volatile float ap = arg * pi; v = std::sin((float)ap) / ap; }
        volatile float y = w2 * v;
        s->taps[x] = y;
        sum = sum + y;
    }
    for (uint32_t x = 0; x < n; ++x) {
        volatile float y = s->taps[x] / sum;
        s->taps[x] = y / (float)(2 * n);
    }
    if (!factorize(s->n_in, &s->radix_fwd) || !factorize(s->n_out, &s->radix_inv)) {
        if (err) *err = "resampler block sizes " + std::to_string(n_in) + "/" + std::to_string(n_out) +
                        " have a prime factor > 32768 (rates " + std::to_string(from) + " -> " + std::to_string(to) + ")";
        return false;
    }
    if (want_spectrum) {
        // spectrum of the f32 taps zero-padded to 2n, bins [0, n_keep); direct O(n*n_keep) DFT
        // in double with a recurrence-free angle (exact index reduction mod 2n)
        const uint32_t L = 2 * n;
        std::vector<double> cs(L), sn(L);
        for (uint32_t j = 0; j < L; ++j) {
            double a = -2.0 * 3.14159265358979323846 * (double)j / (double)L;
            cs[j] = std::cos(a); sn[j] = std::sin(a);
        }
        s->filt_re.assign(s->n_keep, 0.f); s->filt_im.assign(s->n_keep, 0.f);
        for (uint32_t kbin = 0; kbin < s->n_keep; ++kbin) {
            double re = 0, im = 0;
            uint32_t idx = 0;
            for (uint32_t x = 0; x < n; ++x) {
                re += (double)s->taps[x] * cs[idx];
                im += (double)s->taps[x] * sn[idx];
                idx += kbin; if (idx >= L) idx -= L;
            }
            s->filt_re[kbin] = (float)re; s->filt_im[kbin] = (float)im;
        }
    }
    return true;
}

uint64_t resampled_len(uint64_t src_len, const ResamplerSpec& s) {
    uint64_t nfull = src_len / s.n_in, rem = src_len % s.n_in;
    uint64_t n = nfull * s.n_out;
    if (rem) {
        uint64_t of = (uint64_t)std::ceil((double)rem * (double)s.to / (double)s.from);
        n += of < s.n_out ? of : s.n_out;
    }
    return n;
}

}  // namespace bb
