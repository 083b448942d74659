// Mutation fuzzer for the host-side parsers of the library (WAV / RF64 header + reads, FLAC STREAMINFO + frame index), built
// with AddressSanitizer and UBSan by tools/fuzz/run.sh: truncated, bit-flipped, size-field-poisoned and chunk-injected files
// must be rejected or read correctly, never read out of bounds.  usage: parsers_asan a.wav a.flac [iterations]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cstdint>
#include "../../include/birda_b200.h"
static uint64_t rng_s = 88172645463325252ull;
static uint32_t rnd() { rng_s ^= rng_s << 13; rng_s ^= rng_s >> 7; rng_s ^= rng_s << 17; return (uint32_t)(rng_s >> 11); }
static std::vector<unsigned char> slurp(const char* p) { FILE* f = fopen(p, "rb"); std::vector<unsigned char> v; if (!f) return v; fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); v.resize(n); if (fread(v.data(), 1, n, f) != (size_t)n) v.clear(); fclose(f); return v; }
int main(int argc, char** argv) {
    const int iters = argc > 3 ? atoi(argv[3]) : 20000;
    std::vector<unsigned char> wav = slurp(argv[1]), flac = slurp(argv[2]);
    int ok = 0, bad = 0;
    for (int it = 0; it < iters; ++it) {
        std::vector<unsigned char> b = wav;
        switch (it % 4) {
            case 0: for (int k = 0; k < 3; ++k) b[rnd() % 44] = (unsigned char)rnd(); break;
            case 1: b.resize(rnd() % b.size()); break;
            case 2: { uint32_t v[] = {0, 1, 2, 0x7fffffffu, 0xffffffffu, 0xfffffffeu, (uint32_t)b.size()}; uint32_t x = v[rnd() % 7]; memcpy(&b[rnd() % 40], &x, 4); } break;
            default: { size_t pos = 12 + rnd() % 32; std::vector<unsigned char> ins(1 + rnd() % 64); for (auto& c : ins) c = (unsigned char)rnd(); b.insert(b.begin() + pos, ins.begin(), ins.end()); }
        }
        FILE* f = fopen("/tmp/bb_fuzz.wav", "wb"); fwrite(b.data(), 1, b.size(), f); fclose(f);
        bb_wav_info info;
        if (bb_wav_probe("/tmp/bb_fuzz.wav", &info) == 0) {
            uint64_t fr = info.frames < 1000 ? info.frames : 1000;
            std::vector<unsigned char> dst(fr * info.channels * (info.bits_per_sample / 8) + 1);   // exact size: ASan sees overruns
            if (bb_wav_read("/tmp/bb_fuzz.wav", &info, 0, fr, dst.data()) == 0) ++ok; else ++bad;
            bb_wav_read_parallel("/tmp/bb_fuzz.wav", &info, 0, fr, dst.data(), 3);
        } else ++bad;
    }
    printf("wav ok %d rejected %d\n", ok, bad);
    ok = bad = 0;
    for (int it = 0; it < iters; ++it) {
        std::vector<unsigned char> b = flac;
        switch (it % 3) {
            case 0: for (int k = 0; k < 5; ++k) b[rnd() % b.size()] = (unsigned char)rnd(); break;
            case 1: b.resize(rnd() % b.size()); break;
            default: for (int k = 0; k < 3; ++k) b[rnd() % 60] = (unsigned char)rnd();
        }
        std::vector<unsigned char> exact(b.begin(), b.end());          // heap block of exactly n bytes
        bb_flac_info info;
        if (bb_flac_probe_bytes(exact.data(), exact.size(), &info) == 0) {
            const uint64_t cap = 64 + rnd() % 64;
            std::vector<uint64_t> off(cap), fs(cap); std::vector<uint32_t> bs(cap); uint64_t n = 0;
            if (bb_flac_index(exact.data(), exact.size(), &info, off.data(), fs.data(), bs.data(), cap, &n) == 0) ++ok; else ++bad;
        } else ++bad;
    }
    printf("flac ok %d rejected %d\n", ok, bad);
    return 0;
}
