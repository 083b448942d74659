// K6 — FLAC ingest (product code; SURVEY.md 8f-1: "PCM ingest: WAV/RF64 (then FLAC) parse + decode feeding K1").
//
// The reference decodes FLAC on its decode thread through symphonia (src/audio/decode.rs:54-128, :205-245; the
// extension list src/pipeline/coordinator.rs:181-190) and hands append_samples left-justified S32 samples, i.e. the
// value converted is sample / 2^(bps-1) (decode.rs:386-402).  FLAC is lossless, so the decoded PCM is a property of
// the file, not of the decoder: here the compressed file crosses PCIe (roughly half the bytes of its PCM), the host
// only finds the frame boundaries, and the GPU decodes — one thread per frame (frames are independent, the subframes of
// a frame are not: each starts where the previous one ends; the lanes of a warp decode 32 frames in lockstep through a
// flat, loop-free per-sample path) — into the interleaved buffers K1/K2 already consume:
// 16-bit and narrower streams -> BB_S16 (left-shifted to 16 bits), 17..24-bit -> BB_S24 (packed), 32-bit -> BB_S32.
//
// Format: RFC 9639.  Every frame's CRC-16 is checked on the device; a mismatch (corrupt file, or a false sync that
// fooled the host's index) fails the decode with BB_ERR_IO.
#include "common.cuh"
#include "guard.hpp"
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <vector>

namespace bb {
namespace flac {

struct FrameRef { uint64_t offset; uint32_t bytes; uint32_t blocksize; uint64_t first_sample; };

struct Header {
    uint32_t blocksize, sample_rate, channels, bps, header_bytes; int mode;      // mode: 0 independent, 1 left/side, 2 side/right, 3 mid/side
    uint64_t number; bool variable;
};

__host__ __device__ inline uint8_t crc8_update(uint8_t c, uint8_t b) {
    c ^= b;
#pragma unroll
    for (int i = 0; i < 8; ++i) c = (c & 0x80) ? (uint8_t)((c << 1) ^ 0x07) : (uint8_t)(c << 1);
    return c;
}

// Parses a frame header at p (n bytes available).  Returns false when it is not a valid header (sync, reserved bits,
// reserved codes, UTF-8 number, CRC-8).
__host__ __device__ inline bool parse_header(const uint8_t* p, uint64_t n, uint32_t si_rate, uint32_t si_bps, Header* h) {
    if (n < 6 || p[0] != 0xFF || (p[1] & 0xFE) != 0xF8) return false;
    h->variable = (p[1] & 1) != 0;
    const uint32_t bcode = p[2] >> 4, rcode = p[2] & 15, ccode = p[3] >> 4, scode = (p[3] >> 1) & 7;
    if ((p[3] & 1) || bcode == 0 || rcode == 15 || ccode > 10 || scode == 3) return false;
    uint32_t pos = 4;
    // UTF-8 coded frame / sample number
    const uint8_t lead = p[pos];
    uint32_t extra = 0; uint64_t v = 0;
    if (lead < 0x80) { v = lead; }
    else if ((lead & 0xE0) == 0xC0) { extra = 1; v = lead & 0x1F; }
    else if ((lead & 0xF0) == 0xE0) { extra = 2; v = lead & 0x0F; }
    else if ((lead & 0xF8) == 0xF0) { extra = 3; v = lead & 0x07; }
    else if ((lead & 0xFC) == 0xF8) { extra = 4; v = lead & 0x03; }
    else if ((lead & 0xFE) == 0xFC) { extra = 5; v = lead & 0x01; }
    else if (lead == 0xFE) { extra = 6; v = 0; }
    else return false;
    ++pos;
    if (n < pos + extra + 1) return false;
    for (uint32_t i = 0; i < extra; ++i, ++pos) { if ((p[pos] & 0xC0) != 0x80) return false; v = (v << 6) | (p[pos] & 0x3F); }
    h->number = v;
    if (bcode == 1) h->blocksize = 192;
    else if (bcode <= 5) h->blocksize = 576u << (bcode - 2);
    else if (bcode == 6) { if (n < pos + 2) return false; h->blocksize = (uint32_t)p[pos] + 1; pos += 1; }
    else if (bcode == 7) { if (n < pos + 3) return false; h->blocksize = (((uint32_t)p[pos] << 8) | p[pos + 1]) + 1; pos += 2; }
    else h->blocksize = 256u << (bcode - 8);
    if (rcode < 12) {
        uint32_t r = si_rate;
        switch (rcode) { case 1: r = 88200; break; case 2: r = 176400; break; case 3: r = 192000; break; case 4: r = 8000; break;
                         case 5: r = 16000; break; case 6: r = 22050; break; case 7: r = 24000; break; case 8: r = 32000; break;
                         case 9: r = 44100; break; case 10: r = 48000; break; case 11: r = 96000; break; default: break; }
        h->sample_rate = r;
    }
    else if (rcode == 12) { if (n < pos + 2) return false; h->sample_rate = (uint32_t)p[pos] * 1000u; pos += 1; }
    else { if (n < pos + 3) return false; const uint32_t r = ((uint32_t)p[pos] << 8) | p[pos + 1]; h->sample_rate = rcode == 13 ? r : r * 10u; pos += 2; }
    if (ccode < 8) { h->channels = ccode + 1; h->mode = 0; } else { h->channels = 2; h->mode = (int)ccode - 7; }
    switch (scode) { case 1: h->bps = 8; break; case 2: h->bps = 12; break; case 4: h->bps = 16; break; case 5: h->bps = 20; break;
                     case 6: h->bps = 24; break; case 7: h->bps = 32; break; default: h->bps = si_bps; break; }
    if (n < pos + 1) return false;
    uint8_t c = 0;
    for (uint32_t i = 0; i < pos; ++i) c = crc8_update(c, p[i]);
    if (c != p[pos]) return false;
    h->header_bytes = pos + 1;
    return true;
}

// ---------------------------------------------------------------------------------------------- device decode
// MSB-first bit reader over aligned 32-bit words: the top n bits of acc are valid; a refill appends one big-endian word
// (one load + one byte permute) whenever fewer than the requested bits are left, so every request of up to 33 bits is
// one predicated refill — no loops on the common path, which is what lets the lanes of a warp (one frame each) run
// the per-sample code in lockstep.  Reads past the end of the frame see the next frame's bytes (the file buffer has
// slack after its end); the CRC-16 check has already vouched for the frame.
struct BitReader {
    const uint32_t* w; uint64_t acc; int n;
    __device__ __forceinline__ void init(const uint8_t* p) {
        acc = 0; n = 0;
        while (reinterpret_cast<uintptr_t>(p) & 3) { acc |= (uint64_t)__ldg(p++) << (56 - n); n += 8; }      // up to three head bytes
        w = reinterpret_cast<const uint32_t*>(p);
    }
    __device__ __forceinline__ void need(int bits) {                  // bits <= 32; afterwards n >= bits
        if (n < bits) { acc |= (uint64_t)__byte_perm(__ldg(w++), 0, 0x0123) << (32 - n); n += 32; }
    }
    __device__ __forceinline__ uint32_t read(int k) {                 // k <= 32
        if (k == 0) return 0;
        need(k);
        const uint32_t v = (uint32_t)(acc >> (64 - k));
        acc <<= k; n -= k;
        return v;
    }
    __device__ __forceinline__ int32_t read_signed(int k) {
        if (k == 0) return 0;
        const uint32_t v = read(k);
        return k == 32 ? (int32_t)v : (int32_t)(v << (32 - k)) >> (32 - k);
    }
    __device__ __forceinline__ uint32_t unary() {                     // zeros before the terminating one (which is consumed)
        need(32);
        uint32_t hi = (uint32_t)(acc >> 32);
        if (hi != 0) { const int z = __clz((int)hi); acc <<= (z + 1); n -= z + 1; return (uint32_t)z; }
        uint32_t q = 0;                                               // 32 or more zeros: rare
        for (int guard = 0; guard < (1 << 16); ++guard) {
            q += 32; acc <<= 32; n -= 32;
            need(32);
            hi = (uint32_t)(acc >> 32);
            if (hi != 0) { const int z = __clz((int)hi); acc <<= (z + 1); n -= z + 1; return q + (uint32_t)z; }
        }
        return 0xFFFFFFFFu;
    }
};

struct FrameMeta { int32_t mode; int32_t status; };      // status: 0 ok, else what failed

__constant__ uint16_t c_crc16[256];

// CRC-16 of every frame, one thread per frame: the same short loop for every thread of a warp (they differ only in
// the frame length), so this part runs at full SIMT width.
__global__ void __launch_bounds__(128)
flac_crc_kernel(const uint8_t* __restrict__ file, const FrameRef* __restrict__ frames, uint32_t nframes, FrameMeta* __restrict__ meta) {
    __shared__ uint16_t tab[256];                    // lanes look up different entries: shared memory, not the constant cache
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = c_crc16[i];
    __syncthreads();
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    const FrameRef fr = frames[f];
    const uint8_t* base = file + fr.offset;
    const uint32_t len = fr.bytes >= 2 ? fr.bytes - 2 : 0;
    uint32_t c = 0, i = 0;
    auto step = [&](uint32_t byte) { c = ((c << 8) & 0xFFFFu) ^ tab[((c >> 8) ^ byte) & 0xFF]; };
    for (; i < len && ((reinterpret_cast<uintptr_t>(base + i)) & 3); ++i) step(__ldg(base + i));
    for (; i + 4 <= len; i += 4) {                   // aligned words: one load per four table steps
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(base + i));
        step(w & 0xFF); step((w >> 8) & 0xFF); step((w >> 16) & 0xFF); step(w >> 24);
    }
    for (; i < len; ++i) step(__ldg(base + i));
    FrameMeta m{0, 0};
    if (fr.bytes < 2 || c != (((uint32_t)base[fr.bytes - 2] << 8) | base[fr.bytes - 1])) m.status = 2;
    meta[f] = m;
}

// The entropy decode: one LANE per frame, 32 frames per warp.  A frame's bit stream is serial, so the parallelism is
// across frames; for the lanes of a warp to stay together the per-sample code has no data-dependent loops or early
// exits: ONE flat loop over the samples of a subframe (partition boundaries are a per-lane counter, not a nested loop),
// a bit reader that refills by predicated word appends, the unary run from one clz, history in registers (fixed
// predictors) or a thread-local ring (LPC), eight samples per pair of 16-byte stores.  Lanes diverge only in the
// subframe headers (once per 4 096 samples) and on rare events (escape partitions, runs of 32+ zeros).  First version
// (nested loops, byte-wise refills): 6.2 ms for a 10-minute stereo file, issue-bound with the lanes serialised.
__global__ void __launch_bounds__(128)
flac_decode_kernel(const uint8_t* __restrict__ file, const FrameRef* __restrict__ frames, uint32_t nframes, uint32_t si_rate, uint32_t si_bps,
                   uint32_t channels, int32_t* __restrict__ scratch, FrameMeta* __restrict__ meta, uint32_t lanes_per_warp) {
    // lanes_per_warp frames per warp (the other lanes idle): a short file has too few frames to hide the latency of
    // each warp's serial chain with 32 per warp (a 10-minute stereo file is 202 warps on 148 SMs), so the launcher
    // trades SIMT width for warps until there are several per SM
    const uint32_t lane = threadIdx.x & 31u;
    if (lane >= lanes_per_warp) return;
    const uint32_t f = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * lanes_per_warp + lane;
    if (f >= nframes) return;
    const FrameRef fr = frames[f];
    const uint8_t* base = file + fr.offset;
    FrameMeta m = meta[f];
    if (m.status) return;
    Header h;
    if (!parse_header(base, fr.bytes, si_rate, si_bps, &h) || h.blocksize != fr.blocksize || h.channels != channels || fr.bytes < h.header_bytes + 2) {
        m.status = 1; meta[f] = m; return;
    }
    m.mode = h.mode;
    BitReader br; br.init(base + h.header_bytes);
    const uint32_t bs = h.blocksize;
    int32_t* __restrict__ out0 = scratch + fr.first_sample * channels;
    for (uint32_t ch = 0; ch < channels && m.status == 0; ++ch) {
        int32_t* __restrict__ s = out0 + (size_t)ch * bs;
        int bps = (int)h.bps + ((h.mode == 1 && ch == 1) || (h.mode == 2 && ch == 0) || (h.mode == 3 && ch == 1) ? 1 : 0);
        if (br.read(1) != 0) { m.status = 3; break; }
        const uint32_t type = br.read(6);
        int wasted = 0;
        if (br.read(1)) { const uint32_t u = br.unary(); if (u > 31) { m.status = 3; break; } wasted = (int)u + 1; }
        bps -= wasted;
        if (bps <= 0 || bps > 32) { m.status = 3; break; }      // a 33-bit side channel (32-bit stereo-decorrelated stream) does not fit the scratch
        if (type == 0) {                                   // CONSTANT
            const int32_t v = (int32_t)((uint32_t)br.read_signed(bps) << wasted);
            for (uint32_t i = 0; i < bs; ++i) s[i] = v;
            continue;
        }
        const bool verbatim = type == 1, lpc = (type & 0x20) != 0;
        if (!verbatim && !lpc && (type & 0x38) != 0x08) { m.status = 3; break; }
        const int order = verbatim ? 0 : lpc ? (int)(type & 31) + 1 : (int)(type & 7);
        if ((!lpc && order > 4) || (uint32_t)order > bs) { m.status = 3; break; }
        int32_t hist[32];
        for (int i = 0; i < order; ++i) { const int32_t v = br.read_signed(bps); hist[i] = v; s[i] = (int32_t)((uint32_t)v << wasted); }
        int coef[32]; int shift = 0; bool wide = false;
        if (lpc) {
            const int prec = (int)br.read(4) + 1;
            if (prec == 16) { m.status = 3; break; }
            shift = br.read_signed(5);
            if (shift < 0) { m.status = 3; break; }
            for (int j = 0; j < order; ++j) coef[j] = br.read_signed(prec);
            wide = bps + prec + (32 - __clz(order)) > 32;              // the sums may leave 32 bits (libFLAC's rule)
        }
        int32_t p1 = order >= 1 ? hist[order - 1] : 0, p2 = order >= 2 ? hist[order - 2] : 0,
                p3 = order >= 3 ? hist[order - 3] : 0, p4 = order >= 4 ? hist[order - 4] : 0;
        int c1 = 0, c2 = 0, c3 = 0, c4 = 0;                            // fixed predictors as four coefficients
        if (!lpc) {
            if (order == 1) { c1 = 1; } else if (order == 2) { c1 = 2; c2 = -1; } else if (order == 3) { c1 = 3; c2 = -3; c3 = 1; }
            else if (order == 4) { c1 = 4; c2 = -6; c3 = 4; c4 = -1; }
        }
        // residual coding: verbatim samples are read like an escape partition of `bps` bits with no predictor
        int pbits = 4; uint32_t esc = 15u, part_len = bs, next_boundary = 0;
        if (!verbatim) {
            const uint32_t method = br.read(2);
            if (method > 1) { m.status = 3; break; }
            pbits = method ? 5 : 4; esc = method ? 31u : 15u;
            const uint32_t porder = br.read(4);
            if ((bs >> porder) << porder != bs || (bs >> porder) < (uint32_t)order) { m.status = 3; break; }
            part_len = bs >> porder;
        }
        uint32_t k = 0; int raw_bits = verbatim ? bps : -1;            // raw_bits >= 0: samples of this partition are raw
        uint32_t hpos = (uint32_t)order;
        const bool vec_ok = (reinterpret_cast<uintptr_t>(s) & 15) == 0;
        int32_t g[8];                                                  // the 32-byte group being filled
#pragma unroll
        for (int t = 0; t < 8; ++t) g[t] = 0;
        if (vec_ok) for (uint32_t t = (uint32_t)order & ~7u; t < (uint32_t)order; ++t) g[t & 7u] = s[t];
        for (uint32_t i = (uint32_t)order; i < bs; ++i) {
            if (!verbatim && (i == (uint32_t)order || i == next_boundary)) {   // a partition starts here
                if (i == (uint32_t)order) next_boundary = part_len; else next_boundary += part_len;
                k = br.read(pbits);
                raw_bits = -1;
                if (k == esc) raw_bits = (int)br.read(5);
            }
            int32_t r;
            if (raw_bits >= 0) r = br.read_signed(raw_bits);
            else {
                const uint32_t q = br.unary();
                if (q == 0xFFFFFFFFu) { m.status = 4; break; }
                const uint32_t u = (q << k) | br.read((int)k);
                r = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
            }
            int32_t v;
            if (!lpc) {
                v = r + c1 * p1 + c2 * p2 + c3 * p3 + c4 * p4;
                p4 = p3; p3 = p2; p2 = p1; p1 = v;
            } else {
                if (wide) {
                    long long acc = 0;
                    for (int j = 0; j < order; ++j) acc += (long long)coef[j] * (long long)hist[(hpos - 1 - (uint32_t)j) & 31u];
                    v = r + (int32_t)(acc >> shift);
                } else {
                    int32_t acc = 0;
                    for (int j = 0; j < order; ++j) acc += coef[j] * hist[(hpos - 1 - (uint32_t)j) & 31u];
                    v = r + (acc >> shift);
                }
                hist[hpos & 31u] = v; ++hpos;
            }
            const int32_t o = (int32_t)((uint32_t)v << wasted);
            if (!vec_ok) { s[i] = o; continue; }
#pragma unroll
            for (int t = 0; t < 8; ++t) if ((i & 7u) == (uint32_t)t) g[t] = o;
            if ((i & 7u) == 7u) {
                *reinterpret_cast<int4*>(s + i - 7) = make_int4(g[0], g[1], g[2], g[3]);
                *reinterpret_cast<int4*>(s + i - 3) = make_int4(g[4], g[5], g[6], g[7]);
            }
        }
        if (m.status) break;
        if (vec_ok) for (uint32_t t = bs & ~7u; t < bs; ++t) s[t] = g[t & 7u];      // the last, partial group
    }
    meta[f] = m;
}

// One CTA per frame: undo the stereo decorrelation and write interleaved samples of the output format.
template <int FMT>
__global__ void __launch_bounds__(256)
flac_interleave_kernel(const FrameRef* __restrict__ frames, const FrameMeta* __restrict__ meta, uint32_t channels, int shift,
                       const int32_t* __restrict__ scratch, void* __restrict__ out) {
    const FrameRef fr = frames[blockIdx.x];
    const FrameMeta m = meta[blockIdx.x];
    if (m.status) return;
    const uint32_t bs = fr.blocksize;
    const int32_t* __restrict__ s = scratch + fr.first_sample * channels;
    for (uint32_t t = threadIdx.x; t < bs * channels; t += blockDim.x) {
        const uint32_t i = t / channels, c = t - i * channels;
        int32_t v;
        if (m.mode == 0) v = s[(size_t)c * bs + i];
        else {
            const int32_t a = s[i], b = s[(size_t)bs + i];
            int32_t l, r;
            if (m.mode == 1) { l = a; r = a - b; }                       // left, side
            else if (m.mode == 2) { l = a + b; r = b; }                  // side, right
            else { const int32_t mid = (int32_t)(((uint32_t)a << 1) | ((uint32_t)b & 1u)); l = (mid + b) >> 1; r = (mid - b) >> 1; }
            v = c == 0 ? l : r;
        }
        v = (int32_t)((uint32_t)v << shift);
        const uint64_t o = (fr.first_sample + i) * channels + c;
        if (FMT == BB_S16) static_cast<int16_t*>(out)[o] = (int16_t)v;
        else if (FMT == BB_S32) static_cast<int32_t*>(out)[o] = v;
        else { unsigned char* q = static_cast<unsigned char*>(out) + o * 3; q[0] = (unsigned char)v; q[1] = (unsigned char)(v >> 8); q[2] = (unsigned char)(v >> 16); }
    }
}

}  // namespace flac
}  // namespace bb

using namespace bb::flac;

struct bb_flac {
    bb_ctx* ctx = nullptr;
    uint8_t* d_file = nullptr; uint64_t d_file_bytes = 0;
    FrameRef* d_frames = nullptr; FrameMeta* d_meta = nullptr; uint64_t frames_cap = 0;
    int32_t* d_scratch = nullptr; uint64_t scratch_elems = 0;
    void* d_pcm = nullptr; uint64_t d_pcm_bytes = 0;
    std::vector<FrameRef> frames; std::vector<FrameMeta> meta;
    bool crc_table_loaded = false;
};

namespace {
int flac_fail(bb_ctx* c, int code, const std::string& m) { if (c) c->last_error = m; bb::set_tls_error(m); return code; }

// Frame boundaries: a header that parses (sync, reserved bits and codes, CRC-8) AND carries the expected next frame /
// sample number ends the previous frame.  The device checks every frame's CRC-16, so a false sync that survives both
// tests cannot yield wrong samples silently.
int index_frames(const uint8_t* p, uint64_t n, const bb_flac_info& info, std::vector<FrameRef>* out, std::string* err) {
    out->clear();
    uint64_t pos = info.first_frame_offset, sample = 0, prev_len = 0;
    Header h;
    if (pos >= n) return BB_OK;                          // no audio frames
    if (!parse_header(p + pos, n - pos, info.sample_rate, info.bits_per_sample, &h)) { *err = "no FLAC frame where the metadata ends"; return BB_ERR_IO; }
    for (;;) {
        if (h.channels != info.channels || h.bps != info.bits_per_sample) { *err = "frame with a different channel count or sample size than STREAMINFO"; return BB_ERR_UNSUPPORTED_FORMAT; }
        const uint64_t first = h.variable ? h.number : sample;
        if (first != sample) { *err = "frame numbers are not contiguous"; return BB_ERR_IO; }
        const uint64_t next_number = h.variable ? sample + h.blocksize : h.number + 1;
        // search for the header of the next frame.  STREAMINFO's minimum frame size is of little help (the short last
        // block of a file drags it down), so the search first starts three quarters of the previous frame's length in —
        // frames of one stream are similar in size — and falls back to the safe start when that finds nothing acceptable.
        const uint64_t safe = std::max<uint64_t>(pos + h.header_bytes + 2, info.min_frame_bytes > h.header_bytes + 2 ? pos + info.min_frame_bytes : 0);
        uint64_t next = n; Header hn;
        auto scan_from = [&](uint64_t q, uint64_t limit) -> bool {
            if (limit > n) limit = n;
            while (q + 1 < limit) {
                const void* m = std::memchr(p + q, 0xFF, limit - 1 - q);
                if (!m) break;
                q = (uint64_t)(static_cast<const uint8_t*>(m) - p);
                if ((p[q + 1] & 0xFE) == 0xF8 && parse_header(p + q, n - q, info.sample_rate, info.bits_per_sample, &hn) &&
                    hn.number == next_number && hn.variable == h.variable) { next = q; return true; }
                ++q;
            }
            return false;
        };
        // the guess only looks a bounded distance ahead: if this frame is much shorter than the last one the true header
        // lies before the guess, and an unbounded search would run to the end of the file before the fallback finds it
        const uint64_t guess = prev_len ? pos + prev_len - prev_len / 4 : 0;
        if (!(guess > safe && scan_from(guess, guess + 2 * prev_len + 65536))) scan_from(safe, n);
        prev_len = next - pos;
        if (next - pos > 0xFFFFFFFFull) { *err = "frame too large"; return BB_ERR_IO; }
        out->push_back({pos, (uint32_t)(next - pos), h.blocksize, sample});
        sample += h.blocksize;
        if (next >= n) break;
        pos = next; h = hn;
    }
    return BB_OK;
}
}  // namespace

extern "C" {

int32_t bb_flac_probe_bytes(const void* bytes, uint64_t n, bb_flac_info* out) {
    BB_TRY
    if (!bytes || !out) return flac_fail(nullptr, BB_ERR_INVALID_ARG, "null argument");
    std::memset(out, 0, sizeof(*out));
    const uint8_t* p = static_cast<const uint8_t*>(bytes);
    if (n < 42 || std::memcmp(p, "fLaC", 4) != 0) return flac_fail(nullptr, BB_ERR_UNSUPPORTED_FORMAT, "not a FLAC stream");
    uint64_t pos = 4; bool have_si = false;
    for (;;) {
        if (pos + 4 > n) return flac_fail(nullptr, BB_ERR_IO, "truncated FLAC metadata");
        const bool last = (p[pos] & 0x80) != 0; const uint32_t type = p[pos] & 0x7F;
        const uint32_t len = ((uint32_t)p[pos + 1] << 16) | ((uint32_t)p[pos + 2] << 8) | p[pos + 3];
        pos += 4;
        if (pos + len > n) return flac_fail(nullptr, BB_ERR_IO, "truncated FLAC metadata");
        if (type == 0 && len >= 34) {
            const uint8_t* s = p + pos;
            out->min_block = ((uint32_t)s[0] << 8) | s[1]; out->max_block = ((uint32_t)s[2] << 8) | s[3];
            out->min_frame_bytes = ((uint32_t)s[4] << 16) | ((uint32_t)s[5] << 8) | s[6];
            out->max_frame_bytes = ((uint32_t)s[7] << 16) | ((uint32_t)s[8] << 8) | s[9];
            out->sample_rate = ((uint32_t)s[10] << 12) | ((uint32_t)s[11] << 4) | (s[12] >> 4);
            out->channels = ((s[12] >> 1) & 7) + 1;
            out->bits_per_sample = (((uint32_t)(s[12] & 1) << 4) | (s[13] >> 4)) + 1;
            out->frames = ((uint64_t)(s[13] & 15) << 32) | ((uint64_t)s[14] << 24) | ((uint64_t)s[15] << 16) | ((uint64_t)s[16] << 8) | s[17];
            have_si = true;
        }
        pos += len;
        if (last) break;
    }
    if (!have_si || out->sample_rate == 0) return flac_fail(nullptr, BB_ERR_UNSUPPORTED_FORMAT, "FLAC stream without STREAMINFO");
    out->first_frame_offset = pos; out->file_bytes = n;
    out->fmt = out->bits_per_sample <= 16 ? BB_S16 : out->bits_per_sample <= 24 ? BB_S24 : BB_S32;
    return BB_OK;
    BB_CATCH(nullptr)
}

int32_t bb_flac_probe(const char* path, bb_flac_info* out) {
    BB_TRY
    if (!path || !out) return flac_fail(nullptr, BB_ERR_INVALID_ARG, "null argument");
    int fd = ::open(path, O_RDONLY);
    if (fd < 0) return flac_fail(nullptr, BB_ERR_IO, std::string("cannot open ") + path);
    struct stat st;
    if (::fstat(fd, &st) != 0) { ::close(fd); return flac_fail(nullptr, BB_ERR_IO, std::string("cannot stat ") + path); }
    // the metadata of a FLAC file sits in front of the audio: the first MB holds STREAMINFO in any sane file; a file
    // whose metadata (pictures) is longer is probed again with everything
    std::vector<uint8_t> head((size_t)std::min<uint64_t>((uint64_t)st.st_size, 1u << 20));
    ssize_t got = ::pread(fd, head.data(), head.size(), 0);
    int32_t rc = got == (ssize_t)head.size() ? bb_flac_probe_bytes(head.data(), head.size(), out) : BB_ERR_IO;
    if (rc == BB_ERR_IO && (uint64_t)st.st_size > head.size()) {
        head.resize((size_t)st.st_size);
        got = ::pread(fd, head.data(), head.size(), 0);
        rc = got == (ssize_t)head.size() ? bb_flac_probe_bytes(head.data(), head.size(), out) : BB_ERR_IO;
    }
    ::close(fd);
    if (rc == BB_OK) out->file_bytes = (uint64_t)st.st_size;
    return rc;
    BB_CATCH(nullptr)
}

// Frame index of a FLAC stream (host only): byte offset, first sample and block size of every frame, found as the
// decoder finds them.  Returns the number of frames in *n_frames; arrays may be NULL or shorter (capacity) to count only.
int32_t bb_flac_index(const void* bytes, uint64_t n, const bb_flac_info* info, uint64_t* offsets, uint64_t* first_samples,
                      uint32_t* block_sizes, uint64_t capacity, uint64_t* n_frames) {
    BB_TRY
    if (!bytes || !info || !n_frames) return flac_fail(nullptr, BB_ERR_INVALID_ARG, "null argument");
    std::vector<FrameRef> fr; std::string err;
    const int rc = index_frames(static_cast<const uint8_t*>(bytes), n, *info, &fr, &err);
    if (rc != BB_OK) return flac_fail(nullptr, rc, err);
    *n_frames = fr.size();
    for (uint64_t i = 0; i < fr.size() && i < capacity; ++i) {
        if (offsets) offsets[i] = fr[i].offset;
        if (first_samples) first_samples[i] = fr[i].first_sample;
        if (block_sizes) block_sizes[i] = fr[i].blocksize;
    }
    return BB_OK;
    BB_CATCH(nullptr)
}

int32_t bb_flac_create(bb_ctx* ctx, bb_flac** out) {
    BB_TRY
    if (!ctx || !out) return flac_fail(ctx, BB_ERR_INVALID_ARG, "null argument");
    bb_flac* f = new bb_flac();
    f->ctx = ctx;
    *out = f;
    return BB_OK;
    BB_CATCH((ctx ? &ctx->last_error : nullptr))
}

void bb_flac_destroy(bb_flac* f) {
    if (!f) return;
    bb::DeviceGuard g(f->ctx->device);
    cudaStreamSynchronize(f->ctx->stream);
    for (void* p : {(void*)f->d_file, (void*)f->d_frames, (void*)f->d_meta, (void*)f->d_scratch, f->d_pcm}) if (p) cudaFree(p);
    delete f;
}

// The whole compressed file (host memory, pinned for an asynchronous copy) -> interleaved PCM on the device in info->fmt.
// *d_pcm stays valid until the next decode on this object; *frames_out = samples per channel.  Synchronous at the end
// (the frame status words come back to the host).
int32_t bb_flac_decode(bb_flac* f, const void* file_bytes, uint64_t n, const bb_flac_info* info, void** d_pcm, uint64_t* frames_out) {
    BB_TRY
    if (!f || !file_bytes || !info || !d_pcm || !frames_out) return flac_fail(f ? f->ctx : nullptr, BB_ERR_INVALID_ARG, "null argument");
    bb_ctx* c = f->ctx;
    *d_pcm = nullptr; *frames_out = 0;
    if (info->channels == 0 || info->channels > 8 || info->bits_per_sample < 4 || info->bits_per_sample > 32)
        return flac_fail(c, BB_ERR_UNSUPPORTED_FORMAT, "unsupported FLAC channel count or sample size");
    std::string err;
    static const bool dbg = [] { const char* e = std::getenv("BIRDA_K6_DEBUG"); return e && e[0] == '1'; }();
    const auto t_idx0 = std::chrono::steady_clock::now();
    int rc = index_frames(static_cast<const uint8_t*>(file_bytes), n, *info, &f->frames, &err);
    const double ms_index = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_idx0).count();
    if (rc != BB_OK) return flac_fail(c, rc, err);
    uint64_t total = 0;
    for (const auto& fr : f->frames) total += fr.blocksize;
    if (f->frames.empty()) return BB_OK;
    BB_DEVICE(c, c->device);
    cudaStream_t st = c->stream;
    auto grow = [&](void** p, uint64_t* have, uint64_t want) -> cudaError_t {
        if (*have >= want) return cudaSuccess;
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return e;
        if (*p) cudaFree(*p);
        *p = nullptr; *have = 0;
        e = cudaMalloc(p, want);
        if (e == cudaSuccess) *have = want;
        return e;
    };
    if (!f->crc_table_loaded) {
        uint16_t t[256];
        for (int i = 0; i < 256; ++i) { uint32_t v = (uint32_t)i << 8; for (int k = 0; k < 8; ++k) v = (v & 0x8000) ? ((v << 1) ^ 0x8005) & 0xFFFF : (v << 1) & 0xFFFF; t[i] = (uint16_t)v; }
        BB_CUDA_OK(c, cudaMemcpyToSymbol(c_crc16, t, sizeof(t)));
        f->crc_table_loaded = true;
    }
    const uint64_t nfr = f->frames.size();
    const uint32_t out_bytes = info->fmt == BB_S16 ? 2 : info->fmt == BB_S24 ? 3 : 4;
    uint64_t frames_bytes = f->frames_cap * sizeof(FrameRef), meta_bytes = f->frames_cap * sizeof(FrameMeta), scratch_bytes = f->scratch_elems * 4;
    BB_CUDA_OK(c, grow((void**)&f->d_file, &f->d_file_bytes, n + 8));
    if (f->frames_cap < nfr) {
        BB_CUDA_OK(c, grow((void**)&f->d_frames, &frames_bytes, nfr * sizeof(FrameRef)));
        BB_CUDA_OK(c, grow((void**)&f->d_meta, &meta_bytes, nfr * sizeof(FrameMeta)));
        f->frames_cap = nfr;
    }
    BB_CUDA_OK(c, grow((void**)&f->d_scratch, &scratch_bytes, total * info->channels * 4));
    f->scratch_elems = scratch_bytes / 4;
    BB_CUDA_OK(c, grow(&f->d_pcm, &f->d_pcm_bytes, total * info->channels * out_bytes + 16));
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (dbg) { for (auto& e : ev) cudaEventCreate(&e); cudaEventRecord(ev[0], st); }
    BB_CUDA_OK(c, cudaMemcpyAsync(f->d_file, file_bytes, n, cudaMemcpyHostToDevice, st));
    BB_CUDA_OK(c, cudaMemcpyAsync(f->d_frames, f->frames.data(), nfr * sizeof(FrameRef), cudaMemcpyHostToDevice, st));
    if (dbg) cudaEventRecord(ev[1], st);
    flac_crc_kernel<<<(unsigned)((nfr + 127) / 128), 128, 0, st>>>(f->d_file, f->d_frames, (uint32_t)nfr, f->d_meta);
    if (dbg) cudaEventRecord(ev[2], st);
    uint32_t lpw = 32;
    while (lpw > 4 && nfr / lpw < (uint64_t)c->sm_count * 8) lpw >>= 1;          // at least ~8 warps per SM, at most 8x the issue slots
    if (const char* e = std::getenv("BIRDA_K6_LANES")) { const int v = std::atoi(e); if (v >= 1 && v <= 32) lpw = (uint32_t)v; }
    const uint64_t nwarps = (nfr + lpw - 1) / lpw;
    flac_decode_kernel<<<(unsigned)((nwarps + 3) / 4), 128, 0, st>>>(f->d_file, f->d_frames, (uint32_t)nfr, info->sample_rate, info->bits_per_sample,
                                                                      info->channels, f->d_scratch, f->d_meta, lpw);
    if (dbg) cudaEventRecord(ev[3], st);
    const int shift = info->fmt == BB_S16 ? 16 - (int)info->bits_per_sample : info->fmt == BB_S24 ? 24 - (int)info->bits_per_sample : 32 - (int)info->bits_per_sample;
    if (info->fmt == BB_S16) flac_interleave_kernel<BB_S16><<<(unsigned)nfr, 256, 0, st>>>(f->d_frames, f->d_meta, info->channels, shift, f->d_scratch, f->d_pcm);
    else if (info->fmt == BB_S24) flac_interleave_kernel<BB_S24><<<(unsigned)nfr, 256, 0, st>>>(f->d_frames, f->d_meta, info->channels, shift, f->d_scratch, f->d_pcm);
    else flac_interleave_kernel<BB_S32><<<(unsigned)nfr, 256, 0, st>>>(f->d_frames, f->d_meta, info->channels, shift, f->d_scratch, f->d_pcm);
    BB_CUDA_OK(c, cudaGetLastError());
    c->launches += 3;
    f->meta.resize(nfr);
    BB_CUDA_OK(c, cudaMemcpyAsync(f->meta.data(), f->d_meta, nfr * sizeof(FrameMeta), cudaMemcpyDeviceToHost, st));
    if (dbg) cudaEventRecord(ev[4], st);
    BB_CUDA_OK(c, bb::ctx_stream_wait(c));
    if (dbg) {
        float a = 0, b2 = 0, d = 0, e2 = 0;
        cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b2, ev[1], ev[2]); cudaEventElapsedTime(&d, ev[2], ev[3]); cudaEventElapsedTime(&e2, ev[3], ev[4]);
        std::fprintf(stderr, "[k6] %llu frames: host index %.2f ms, H2D %.2f, crc %.2f, decode %.2f, interleave + D2H %.2f\n", (unsigned long long)nfr, ms_index, a, b2, d, e2);
        for (auto& e : ev) cudaEventDestroy(e);
    }
    for (uint64_t i = 0; i < nfr; ++i)
        if (f->meta[i].status) {
            static const char* what[] = {"", "frame header does not parse", "frame CRC-16 mismatch", "invalid subframe", "bitstream ends inside a frame"};
            return flac_fail(c, BB_ERR_IO, std::string("FLAC frame ") + std::to_string(i) + ": " + what[f->meta[i].status < 5 ? f->meta[i].status : 3]);
        }
    *d_pcm = f->d_pcm; *frames_out = total;
    return BB_OK;
    BB_CATCH((f && f->ctx ? &f->ctx->last_error : nullptr))
}

}  // extern "C"
