// PCM ingest for the front end: RIFF/WAVE and RF64 header parsing + positioned reads into a caller
// buffer (ideally page-locked, see bb_host_alloc).  SURVEY.md §8f rank 1: at thousands x realtime the
// reference's symphonia decode thread and its three opens per file (src/lib.rs:726,
// src/pipeline/processor.rs:457, :59; reader src/audio/decode.rs:54-128) become the wall; WAV PCM is
// already the interleaved layout K1/K2 consume, so the file is read straight into the staging buffer.
// Sample formats follow what the reference converts (src/audio/decode.rs:353-411): 16-bit PCM -> S16,
// 24-bit PCM -> S24 (3-byte packed; symphonia's PCM decoder presents it as S32 `sample << 8`, so it takes the S32
// arm, decode.rs:386-402), 32-bit PCM -> S32, 32-bit float -> F32.  8-bit PCM and 64-bit float are reported as
// unsupported (append_samples drops U8 / F64 buffers silently, decode.rs:407-409).  Headers whose block size does
// not equal channels * bytes per sample (padded containers, EXTENSIBLE with fewer valid bits than container bits)
// are rejected rather than read at the wrong stride.
#include "../../include/birda_b200.h"
#include "rules.hpp"
#include "guard.hpp"
#include "pipeline_internal.hpp"
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>


namespace {
uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
uint64_t rd64(const unsigned char* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }
int fail(int code, const std::string& m) { bb::set_tls_error(m); return code; }
}  // namespace

namespace bb {
// bytes [off0, off0 + total) of a file into dst, cut into slices read by `threads` threads at once (pread is positional)
int32_t read_range_parallel(const char* path, uint64_t off0, uint64_t total, void* dst, uint32_t threads) {
    int fd = ::open(path, O_RDONLY);
    if (fd < 0) return fail(BB_ERR_IO, std::string("cannot open ") + path);
    auto read_range = [&](uint64_t lo, uint64_t hi) -> bool {
        char* p = static_cast<char*>(dst) + lo;
        uint64_t off = off0 + lo, left = hi - lo;
        while (left) {
            const ssize_t n = ::pread(fd, p, left > (1u << 30) ? (1u << 30) : (size_t)left, (off_t)off);
            if (n <= 0) return false;
            p += n; off += (uint64_t)n; left -= (uint64_t)n;
        }
        return true;
    };
    constexpr uint64_t kMinSlice = 4ull << 20;
    uint64_t n = threads ? threads : 1;
    if (n > total / kMinSlice) n = total / kMinSlice;
    if (n > 64) n = 64;
    bool ok = true;
    if (n <= 1) ok = read_range(0, total);
    else {
        std::vector<char> good(n, 1);
        std::vector<std::thread> th;
        const uint64_t slice = ((total + n - 1) / n + 4095) & ~4095ull;
        auto lo_of = [&](uint64_t i) { return i * slice < total ? i * slice : total; };
        try {
            for (uint64_t i = 1; i < n; ++i) th.emplace_back([&, i] { good[i] = read_range(lo_of(i), lo_of(i + 1)) ? 1 : 0; });
        } catch (...) {                                        // could not start a thread: read what is left here
            for (uint64_t i = th.size() + 1; i < n; ++i) good[i] = read_range(lo_of(i), lo_of(i + 1)) ? 1 : 0;
        }
        good[0] = read_range(0, lo_of(1)) ? 1 : 0;
        for (auto& t : th) t.join();
        for (char g : good) ok = ok && g;
    }
    ::close(fd);
    if (!ok) return fail(BB_ERR_IO, std::string("short read from ") + path);
    return BB_OK;
}
}  // namespace bb

extern "C" {

int32_t bb_wav_probe(const char* path, bb_wav_info* out) {
    BB_TRY
    if (!path || !out) return fail(BB_ERR_INVALID_ARG, "null argument");
    std::memset(out, 0, sizeof(*out));
    int fd = ::open(path, O_RDONLY);
    if (fd < 0) return fail(BB_ERR_IO, std::string("cannot open ") + path);
    struct stat st;
    if (::fstat(fd, &st) != 0) { ::close(fd); return fail(BB_ERR_IO, std::string("cannot stat ") + path); }
    const uint64_t file_size = (uint64_t)st.st_size;
    unsigned char h[12];
    if (::pread(fd, h, 12, 0) != 12 || std::memcmp(h + 8, "WAVE", 4) != 0 ||
        (std::memcmp(h, "RIFF", 4) != 0 && std::memcmp(h, "RF64", 4) != 0)) {
        ::close(fd); return fail(BB_ERR_UNSUPPORTED_FORMAT, std::string(path) + ": not a RIFF/RF64 WAVE file");
    }
    const bool rf64 = std::memcmp(h, "RF64", 4) == 0;
    uint64_t pos = 12, data_size64 = 0; bool have_ds64 = false, have_fmt = false, have_data = false;
    uint16_t tag = 0, channels = 0, bits = 0, block_align = 0, valid_bits = 0; uint32_t rate = 0;
    uint64_t data_off = 0, data_size = 0;
    while (pos + 8 <= file_size) {
        unsigned char ch[8];
        if (::pread(fd, ch, 8, (off_t)pos) != 8) break;
        uint64_t sz = rd32(ch + 4);
        const uint64_t body = pos + 8;
        if (std::memcmp(ch, "ds64", 4) == 0 && sz >= 24) {
            unsigned char d[24];
            if (::pread(fd, d, 24, (off_t)body) == 24) { data_size64 = rd64(d + 8); have_ds64 = true; }
        } else if (std::memcmp(ch, "fmt ", 4) == 0 && sz >= 16) {
            unsigned char f[40]; std::memset(f, 0, sizeof(f));
            const size_t n = sz < 40 ? (size_t)sz : 40;
            if (::pread(fd, f, n, (off_t)body) != (ssize_t)n) break;
            tag = rd16(f); channels = rd16(f + 2); rate = rd32(f + 4); block_align = rd16(f + 12); bits = rd16(f + 14);
            valid_bits = bits;
            if (tag == 0xFFFE && sz >= 40) { valid_bits = rd16(f + 18); tag = rd16(f + 24); }   // WAVE_FORMAT_EXTENSIBLE: wValidBitsPerSample, then the sub-format GUID (starts with the tag)
            have_fmt = true;
        } else if (std::memcmp(ch, "data", 4) == 0) {
            data_off = body;
            data_size = (sz == 0xFFFFFFFFu && rf64 && have_ds64) ? data_size64 : sz;
            // streamed writers leave 0 (or 0xFFFFFFFF without ds64) in the size field: the data runs to the end of the file
            if (data_size == 0 || (sz == 0xFFFFFFFFu && !(rf64 && have_ds64))) data_size = file_size > data_off ? file_size - data_off : 0;
            if (data_off > file_size) data_size = 0;
            else if (data_size > file_size - data_off) data_size = file_size - data_off;     // truncated file: decode what exists
            have_data = true;
            break;
        }
        pos = body + sz + (sz & 1);
    }
    ::close(fd);
    if (!have_fmt || !have_data || channels == 0 || rate == 0)
        return fail(BB_ERR_UNSUPPORTED_FORMAT, std::string(path) + ": missing fmt/data chunk");
    out->sample_rate = rate; out->channels = channels; out->bits_per_sample = bits;
    out->data_offset = data_off;
    const uint32_t bytes = bits / 8;
    out->fmt = 0;
    if (tag == 1 && bits == 16) out->fmt = BB_S16;
    else if (tag == 1 && bits == 24) out->fmt = BB_S24;
    else if (tag == 1 && bits == 32) out->fmt = BB_S32;
    else if (tag == 3 && bits == 32) out->fmt = BB_F32;
    const uint32_t frame_bytes = block_align ? block_align : bytes * channels;
    out->frames = frame_bytes ? data_size / frame_bytes : 0;
    if (out->fmt != 0 && (frame_bytes != bytes * channels || valid_bits != bits)) {
        out->fmt = 0;
        return fail(BB_ERR_UNSUPPORTED_FORMAT, std::string(path) + ": block size " + std::to_string(frame_bytes) + " / " +
                    std::to_string(valid_bits) + " valid bits do not match " + std::to_string(channels) + " channels of " +
                    std::to_string(bits) + "-bit samples");
    }
    if (out->fmt == 0)
        return fail(BB_ERR_UNSUPPORTED_FORMAT, std::string(path) + ": sample format tag " + std::to_string(tag) + " / " +
                    std::to_string(bits) + " bits is not converted by the reference (decode.rs:353-411)");
    return BB_OK;
    BB_CATCH(nullptr)
}

// A read from the page cache is a kernel memcpy: one thread moves 6-10 GB/s, a PCIe 5 x16 link takes ~55 GB/s.  Large
// reads are therefore cut into slices read by `threads` threads at once (pread is positional: no shared file offset).
int32_t bb_wav_read_parallel(const char* path, const bb_wav_info* info, uint64_t first_frame, uint64_t frames, void* dst, uint32_t threads) {
    BB_TRY
    if (!path || !info || (!dst && frames)) return fail(BB_ERR_INVALID_ARG, "null argument");
    if (info->fmt == 0) return fail(BB_ERR_UNSUPPORTED_FORMAT, "unsupported sample format");
    if (first_frame > info->frames || frames > info->frames - first_frame) return fail(BB_ERR_INVALID_ARG, "frame range outside the data chunk");
    const uint64_t fb = (uint64_t)info->channels * bb::sample_bytes(info->fmt);
    return bb::read_range_parallel(path, info->data_offset + first_frame * fb, frames * fb, dst, threads);
    BB_CATCH(nullptr)
}

int32_t bb_wav_read(const char* path, const bb_wav_info* info, uint64_t first_frame, uint64_t frames, void* dst) {
    return bb_wav_read_parallel(path, info, first_frame, frames, dst, 1);
}

}  // extern "C"
