// K6 — FLAC ingest (product code; SURVEY.md 8f-1: "PCM ingest: WAV/RF64 (then FLAC) parse + decode feeding K1").
//
// The reference decodes FLAC on its decode thread through symphonia (src/audio/decode.rs:54-128, :205-245; the
// extension list src/pipeline/coordinator.rs:181-190) and hands append_samples left-justified S32 samples, i.e. the
// value converted is sample / 2^(bps-1) (decode.rs:386-402).  FLAC is lossless, so the decoded PCM is a property of
// the file, not of the decoder: here the compressed file crosses PCIe (roughly half the bytes of its PCM), the host
// only finds the frame boundaries, and the GPU decodes — one thread per frame (frames are independent, the subframes of
// a frame are not: each starts where the previous one ends) — into the interleaved buffers K1/K2 already consume:
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
struct BitReader {
    const uint8_t* p; const uint8_t* end; uint64_t acc; int n;      // the top n bits of acc are valid
    __device__ __forceinline__ void refill() {
        while (n <= 56 && p < end) { acc |= (uint64_t)__ldg(p++) << (56 - n); n += 8; }
    }
    __device__ __forceinline__ uint32_t read(int k) {               // k <= 32
        if (k == 0) return 0;
        if (n < k) refill();
        const uint32_t v = (uint32_t)(acc >> (64 - k));
        acc <<= k; n -= k;
        return v;
    }
    __device__ __forceinline__ int32_t read_signed(int k) {          // k <= 32
        if (k == 0) return 0;
        const uint32_t v = read(k);
        return k == 32 ? (int32_t)v : (int32_t)(v << (32 - k)) >> (32 - k);
    }
    __device__ __forceinline__ int64_t read_signed_wide(int k) {     // k <= 33 (side channel of a 32-bit stream)
        if (k <= 32) return read_signed(k);
        const int64_t hi = read_signed(k - 32);
        return (hi << 32) | read(32);
    }
    __device__ __forceinline__ uint32_t unary() {                    // zeros up to and including the terminating one
        uint32_t q = 0;
        for (;;) {
            if (n == 0) { refill(); if (n == 0) return 0xFFFFFFFFu; }
            const int z = acc ? __clzll((long long)acc) : 64;
            if (z < n) { acc <<= (z + 1); n -= z + 1; return q + (uint32_t)z; }
            q += (uint32_t)n; acc = 0; n = 0;
        }
    }
    __device__ __forceinline__ bool overrun() const { return p >= end && n < 0; }
};

struct FrameMeta { int32_t mode; int32_t status; };      // status: 0 ok, else what failed

__constant__ uint16_t c_crc16[256];

// One thread per frame: header, subframes (into the planar int32 scratch of the frame), CRC-16.
__global__ void __launch_bounds__(64)
flac_decode_kernel(const uint8_t* __restrict__ file, const FrameRef* __restrict__ frames, uint32_t nframes, uint32_t si_rate, uint32_t si_bps,
                   uint32_t channels, int32_t* __restrict__ scratch, FrameMeta* __restrict__ meta) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    const FrameRef fr = frames[f];
    const uint8_t* base = file + fr.offset;
    FrameMeta m{0, 0};
    Header h;
    if (!parse_header(base, fr.bytes, si_rate, si_bps, &h) || h.blocksize != fr.blocksize || h.channels != channels || fr.bytes < h.header_bytes + 2) {
        m.status = 1; meta[f] = m; return;
    }
    m.mode = h.mode;
    {   // CRC-16 of everything but the last two bytes
        uint32_t c = 0;
        for (uint32_t i = 0; i + 2 < fr.bytes; ++i) c = ((c << 8) & 0xFFFFu) ^ c_crc16[((c >> 8) ^ __ldg(base + i)) & 0xFF];
        if (c != (((uint32_t)base[fr.bytes - 2] << 8) | base[fr.bytes - 1])) { m.status = 2; meta[f] = m; return; }
    }
    BitReader br{base + h.header_bytes, base + fr.bytes - 2, 0, 0};
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
        int order = 0;
        if (type == 0) {                                   // CONSTANT
            const int32_t v = (int32_t)br.read_signed_wide(bps);
            for (uint32_t i = 0; i < bs; ++i) s[i] = v;
        } else if (type == 1) {                            // VERBATIM
            for (uint32_t i = 0; i < bs; ++i) s[i] = (int32_t)br.read_signed_wide(bps);
        } else if ((type & 0x38) == 0x08 || (type & 0x20)) {
            const bool lpc = (type & 0x20) != 0;
            order = lpc ? (int)(type & 31) + 1 : (int)(type & 7);
            if ((!lpc && order > 4) || (uint32_t)order > bs) { m.status = 3; break; }
            for (int i = 0; i < order; ++i) s[i] = (int32_t)br.read_signed_wide(bps);
            int coef[32]; int shift = 0;
            if (lpc) {
                const int prec = (int)br.read(4) + 1;
                if (prec == 16) { m.status = 3; break; }
                shift = br.read_signed(5);
                if (shift < 0) { m.status = 3; break; }
                for (int j = 0; j < order; ++j) coef[j] = br.read_signed(prec);
            }
            // residual
            const uint32_t method = br.read(2);
            if (method > 1) { m.status = 3; break; }
            const int pbits = method ? 5 : 4; const uint32_t esc = method ? 31u : 15u;
            const uint32_t porder = br.read(4);
            if ((bs >> porder) << porder != bs || (bs >> porder) < (uint32_t)order) { m.status = 3; break; }
            uint32_t i = (uint32_t)order;
            for (uint32_t part = 0; part < (1u << porder) && m.status == 0; ++part) {
                const uint32_t cnt = (bs >> porder) - (part == 0 ? (uint32_t)order : 0u);
                const uint32_t k = br.read(pbits);
                const uint32_t stop = i + cnt;
                if (k == esc) {
                    const int nb = (int)br.read(5);
                    for (; i < stop; ++i) s[i] = br.read_signed(nb);
                } else {
                    for (; i < stop; ++i) {
                        const uint32_t q = br.unary();
                        if (q == 0xFFFFFFFFu) { m.status = 4; break; }
                        const uint32_t u = (q << k) | br.read((int)k);
                        s[i] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
                    }
                }
            }
            if (m.status) break;
            // prediction: residuals in s[order..) become samples in place
            if (!lpc) {
                switch (order) {
                    case 0: break;
                    case 1: for (uint32_t t = 1; t < bs; ++t) s[t] += s[t - 1]; break;
                    case 2: for (uint32_t t = 2; t < bs; ++t) s[t] += 2 * s[t - 1] - s[t - 2]; break;
                    case 3: for (uint32_t t = 3; t < bs; ++t) s[t] += 3 * s[t - 1] - 3 * s[t - 2] + s[t - 3]; break;
                    default: for (uint32_t t = 4; t < bs; ++t) s[t] += 4 * s[t - 1] - 6 * s[t - 2] + 4 * s[t - 3] - s[t - 4]; break;
                }
            } else {
                for (uint32_t t = (uint32_t)order; t < bs; ++t) {
                    long long acc = 0;
                    for (int j = 0; j < order; ++j) acc += (long long)coef[j] * (long long)s[t - 1 - j];
                    s[t] += (int32_t)(acc >> shift);
                }
            }
        } else { m.status = 3; break; }
        if (wasted) for (uint32_t i = 0; i < bs; ++i) s[i] = (int32_t)((uint32_t)s[i] << wasted);
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
    uint64_t pos = info.first_frame_offset, sample = 0;
    Header h;
    if (pos >= n) return BB_OK;                          // no audio frames
    if (!parse_header(p + pos, n - pos, info.sample_rate, info.bits_per_sample, &h)) { *err = "no FLAC frame where the metadata ends"; return BB_ERR_IO; }
    for (;;) {
        if (h.channels != info.channels || h.bps != info.bits_per_sample) { *err = "frame with a different channel count or sample size than STREAMINFO"; return BB_ERR_UNSUPPORTED_FORMAT; }
        const uint64_t first = h.variable ? h.number : sample;
        if (first != sample) { *err = "frame numbers are not contiguous"; return BB_ERR_IO; }
        const uint64_t next_number = h.variable ? sample + h.blocksize : h.number + 1;
        // search for the header of the next frame
        uint64_t q = pos + h.header_bytes + 2;
        if (info.min_frame_bytes > h.header_bytes + 2 && pos + info.min_frame_bytes > q) q = pos + info.min_frame_bytes;
        uint64_t next = n; Header hn;
        while (q + 1 < n) {
            const void* m = std::memchr(p + q, 0xFF, n - 1 - q);
            if (!m) break;
            q = (uint64_t)(static_cast<const uint8_t*>(m) - p);
            if ((p[q + 1] & 0xFE) == 0xF8 && parse_header(p + q, n - q, info.sample_rate, info.bits_per_sample, &hn) &&
                hn.number == next_number && hn.variable == h.variable) { next = q; break; }
            ++q;
        }
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
    int rc = index_frames(static_cast<const uint8_t*>(file_bytes), n, *info, &f->frames, &err);
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
    BB_CUDA_OK(c, cudaMemcpyAsync(f->d_file, file_bytes, n, cudaMemcpyHostToDevice, st));
    BB_CUDA_OK(c, cudaMemcpyAsync(f->d_frames, f->frames.data(), nfr * sizeof(FrameRef), cudaMemcpyHostToDevice, st));
    flac_decode_kernel<<<(unsigned)((nfr + 63) / 64), 64, 0, st>>>(f->d_file, f->d_frames, (uint32_t)nfr, info->sample_rate, info->bits_per_sample,
                                                                    info->channels, f->d_scratch, f->d_meta);
    const int shift = info->fmt == BB_S16 ? 16 - (int)info->bits_per_sample : info->fmt == BB_S24 ? 24 - (int)info->bits_per_sample : 32 - (int)info->bits_per_sample;
    if (info->fmt == BB_S16) flac_interleave_kernel<BB_S16><<<(unsigned)nfr, 256, 0, st>>>(f->d_frames, f->d_meta, info->channels, shift, f->d_scratch, f->d_pcm);
    else if (info->fmt == BB_S24) flac_interleave_kernel<BB_S24><<<(unsigned)nfr, 256, 0, st>>>(f->d_frames, f->d_meta, info->channels, shift, f->d_scratch, f->d_pcm);
    else flac_interleave_kernel<BB_S32><<<(unsigned)nfr, 256, 0, st>>>(f->d_frames, f->d_meta, info->channels, shift, f->d_scratch, f->d_pcm);
    BB_CUDA_OK(c, cudaGetLastError());
    c->launches += 2;
    f->meta.resize(nfr);
    BB_CUDA_OK(c, cudaMemcpyAsync(f->meta.data(), f->d_meta, nfr * sizeof(FrameMeta), cudaMemcpyDeviceToHost, st));
    BB_CUDA_OK(c, bb::ctx_stream_wait(c));
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
