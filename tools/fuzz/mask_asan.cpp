// bb_mask_build / bb_rule_scientific_name_len on random label sets (empty, malformed UTF-8, long, colliding keys) under ASan + UBSan.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cmath>
#include "../../include/birda_b200.h"
static uint64_t s = 1234567;
static uint32_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); }
static std::string rand_label() {
    static const char* parts[] = {"Turdus merula", "turdus MERULA", "Ärla", "\xCE\x91\xCE\xB2", "\xD0\x96\xD1\x83\xD0\xBA", "x", "", "_", "a_b", "Parus major_Great Tit", "\xff\xfe", "\xC3", "Genus species_", "_Common"};
    std::string r = parts[rnd() % 14];
    if (rnd() % 3 == 0) r += "_" + std::string(parts[rnd() % 14]);
    if (rnd() % 7 == 0) r += std::string(rnd() % 300, 'q');
    if (rnd() % 11 == 0) { r.push_back((char)(0x80 | (rnd() & 0x7f))); }
    return r;
}
int main() {
    for (int it = 0; it < 20000; ++it) {
        const uint32_t nc = rnd() % 40, ng = rnd() % 40, ns = rnd() % 40;
        std::vector<std::string> cl(nc), gl(ng), sp(ns);
        for (auto& x : cl) x = rand_label();
        for (auto& x : gl) x = rand_label();
        for (auto& x : sp) x = (ng && rnd() % 2) ? gl[rnd() % ng] : rand_label();
        std::vector<const char*> pc, pg, ps;
        for (auto& x : cl) pc.push_back(x.c_str());
        for (auto& x : gl) pg.push_back(x.c_str());
        for (auto& x : sp) ps.push_back(x.c_str());
        std::vector<float> val(ns), mask(nc);          // exact sizes
        for (auto& v : val) v = (float)(rnd() % 1000) / 1000.f;
        uint32_t mapped = 0, unmatched = 0;
        const int rc = bb_mask_build(pc.data(), nc, pg.data(), ng, ps.data(), val.data(), ns, mask.data(), rnd() % 2 ? &mapped : nullptr, rnd() % 2 ? &unmatched : nullptr);
        if (rc != 0) { printf("rc %d\n", rc); return 1; }
        for (auto& x : cl) (void)bb_rule_scientific_name_len(x.c_str());
    }
    printf("mask build: 20000 random label sets ok\n");
    return 0;
}
