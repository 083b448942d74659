// Host rules shared by the C ABI and the host pipeline (product code, NOT the oracle).
// Each function restates an in-tree rule of the reference; see include/birda_b200.h for
// the file:line citations.
#pragma once
#include <cstdint>
#include <vector>
#include <string>

namespace bb {

// bytes of one sample of a bb_sample_fmt (BB_S16 = 1, BB_S32 = 2, BB_F32 = 3, BB_S24 = 4: 3-byte packed)
inline uint32_t sample_bytes(int fmt) { return fmt == 1 ? 2u : fmt == 4 ? 3u : 4u; }

struct Window { uint64_t start; uint64_t take; };

// StreamingDecoder::next_segment as a closed form over a fully buffered stream
// (src/audio/decode.rs:150-202): every window starts at i*hop except possibly the last.
struct WindowSeq {
    uint64_t total = 0, seg = 0, ovl = 0, hop = 0;
    uint64_t nseg = 0;
    uint64_t last_start = 0;       // start of window nseg-1
    Window at(uint64_t i) const {
        uint64_t st = (i + 1 == nseg) ? last_start : i * hop;
        uint64_t tk = total - st < seg ? total - st : seg;
        return Window{st, tk};
    }
};
// full_only: stop after the last FULL window (streaming a piece that is not EOF).
bool make_window_seq(uint64_t total, uint64_t seg, uint64_t ovl, bool full_only, WindowSeq* out);

uint64_t trunc_f32_to_u64(float x);
void     segment_samples(float segment_duration, float overlap, uint32_t target_rate, bool bat,
                         uint64_t* seg, uint64_t* ovl);
void     source_window(uint64_t seg, uint64_t ovl, uint32_t sr, uint32_t tr, uint64_t* sseg, uint64_t* sovl);
void     chunk_times(uint64_t start_sample, uint32_t sr, uint64_t seg, uint32_t tr, float* st, float* et);
int64_t  estimate_segment_count(double duration, bool has_duration, float seg_dur, float overlap);
uint32_t effective_batch_size(uint32_t batch, int64_t estimate);
uint32_t date_to_week(uint32_t month, uint32_t day);
uint32_t week_to_start_day(uint32_t week);
void     day_of_year_to_date(uint32_t doy, uint32_t* month, uint32_t* day);

// rubato Fft::<f32>::new(from, to, 1024, 1, FixedSync::Both) — block sizes, cutoff, taps.
struct ResamplerSpec {
    uint32_t from = 0, to = 0;
    uint32_t n_in = 0, n_out = 0, n_keep = 0;
    float cutoff = 0.f;
    std::vector<float> taps;          // [n_in], / (2*n_in) applied
    std::vector<float> filt_re, filt_im;   // [n_keep] spectrum of taps zero-padded to 2*n_in
    // FFT factorisations (complex transforms of length n_in and n_out), odd radices first
    std::vector<int> radix_fwd, radix_inv;
};
bool make_resampler_spec(uint32_t from, uint32_t to, bool want_spectrum, ResamplerSpec* out, std::string* err);
uint64_t resampled_len(uint64_t src_len, const ResamplerSpec& s);
bool factorize(uint32_t n, std::vector<int>* radices);   // false if a prime factor > 31 remains

}  // namespace bb
