// Range-filter precompute, label side (product code, pure host): SpeciesMapping::build and GeomodelScores::project
// (src/inference/geomodel.rs:28-38, :58-87, :140-157) producing the dense [C] mask K3 reads (NaN = the classifier
// label has no geomodel entry).  The geomodel forward itself is bb_dense_run (K4); date -> week is bb_rule_date_to_week.
#include "guard.hpp"
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

namespace bb {

// scientific_name(): the part before the first '_' when it contains a space, else the whole label (geomodel.rs:28-33)
size_t scientific_name_len(const char* label) {
    const size_t n = std::strlen(label);
    const char* us = static_cast<const char*>(std::memchr(label, '_', n));
    if (!us) return n;
    const size_t pre = (size_t)(us - label);
    return std::memchr(label, ' ', pre) ? pre : n;
}

namespace {

void put_utf8(std::string& o, uint32_t c) {
    if (c < 0x80) o.push_back((char)c);
    else if (c < 0x800) { o.push_back((char)(0xC0 | (c >> 6))); o.push_back((char)(0x80 | (c & 0x3F))); }
    else if (c < 0x10000) { o.push_back((char)(0xE0 | (c >> 12))); o.push_back((char)(0x80 | ((c >> 6) & 0x3F))); o.push_back((char)(0x80 | (c & 0x3F))); }
    else { o.push_back((char)(0xF0 | (c >> 18))); o.push_back((char)(0x80 | ((c >> 12) & 0x3F))); o.push_back((char)(0x80 | ((c >> 6) & 0x3F))); o.push_back((char)(0x80 | (c & 0x3F))); }
}

// str::to_lowercase for the scripts label files use: ASCII, Latin-1, Latin Extended-A, Greek and Cyrillic capitals.
// Other code points (and malformed UTF-8 bytes) pass through unchanged.  Differences from Rust's full mapping that
// remain: the context rule for a final capital sigma and letters outside these blocks.
std::string to_lower_utf8(const char* s, size_t n) {
    std::string o;
    o.reserve(n);
    size_t i = 0;
    while (i < n) {
        const unsigned char b = (unsigned char)s[i];
        uint32_t c = b; size_t len = 1;
        if (b >= 0xC2 && b <= 0xDF && i + 1 < n && ((unsigned char)s[i + 1] & 0xC0) == 0x80) {
            c = ((uint32_t)(b & 0x1F) << 6) | ((unsigned char)s[i + 1] & 0x3F); len = 2;
        } else if (b >= 0xE0 && b <= 0xEF && i + 2 < n && ((unsigned char)s[i + 1] & 0xC0) == 0x80 && ((unsigned char)s[i + 2] & 0xC0) == 0x80) {
            c = ((uint32_t)(b & 0x0F) << 12) | (((uint32_t)(unsigned char)s[i + 1] & 0x3F) << 6) | ((unsigned char)s[i + 2] & 0x3F); len = 3;
        } else if (b >= 0x80) { o.push_back((char)b); ++i; continue; }      // 4-byte sequences and stray bytes: as is
        i += len;
        if (c < 0x80) { o.push_back((char)((c >= 'A' && c <= 'Z') ? c + 32 : c)); continue; }
        if (c >= 0xC0 && c <= 0xDE && c != 0xD7) c += 0x20;
        else if (c == 0x130) { put_utf8(o, 'i'); c = 0x307; }                // LATIN CAPITAL I WITH DOT ABOVE -> i + combining dot
        else if (c == 0x178) c = 0xFF;
        else if ((c >= 0x100 && c <= 0x137) || (c >= 0x14A && c <= 0x177)) { if (!(c & 1)) c += 1; }
        else if ((c >= 0x139 && c <= 0x148) || (c >= 0x179 && c <= 0x17E)) { if (c & 1) c += 1; }
        else if (c >= 0x391 && c <= 0x3A9 && c != 0x3A2) c += 0x20;
        else if (c >= 0x410 && c <= 0x42F) c += 0x20;
        else if (c >= 0x400 && c <= 0x40F) c += 0x50;
        put_utf8(o, c);
    }
    return o;
}

std::string species_key(const char* label) { return to_lower_utf8(label, scientific_name_len(label)); }   // geomodel.rs:36-38

}  // namespace
}  // namespace bb

extern "C" {

uint32_t bb_rule_scientific_name_len(const char* label) { return label ? (uint32_t)bb::scientific_name_len(label) : 0; }

int32_t bb_mask_build(const char* const* classifier_labels, uint32_t n_classifier,
                      const char* const* geomodel_labels, uint32_t n_geomodel,
                      const char* const* score_species, const float* score_values, uint32_t n_scores,
                      float* mask, uint32_t* mapped, uint32_t* unmatched) {
    BB_TRY
        using namespace bb;
        if ((n_classifier && (!classifier_labels || !mask)) || (n_geomodel && !geomodel_labels) ||
            (n_scores && (!score_species || !score_values))) { set_tls_error("null argument"); return BB_ERR_INVALID_ARG; }
        for (uint32_t i = 0; i < n_classifier; ++i) if (!classifier_labels[i]) { set_tls_error("null classifier label"); return BB_ERR_INVALID_ARG; }
        for (uint32_t i = 0; i < n_geomodel; ++i) if (!geomodel_labels[i]) { set_tls_error("null geomodel label"); return BB_ERR_INVALID_ARG; }
        for (uint32_t i = 0; i < n_scores; ++i) if (!score_species[i]) { set_tls_error("null score species"); return BB_ERR_INVALID_ARG; }
        // SpeciesMapping::build — species key -> FIRST classifier label with that key (geomodel.rs:58-75)
        std::unordered_map<std::string, uint32_t> classifier_by_key;
        classifier_by_key.reserve(n_classifier * 2u + 1u);
        for (uint32_t i = 0; i < n_classifier; ++i) classifier_by_key.emplace(species_key(classifier_labels[i]), i);
        // ... restricted to keys the geomodel has (:77-83)
        std::unordered_map<std::string, uint32_t> by_species_key;
        by_species_key.reserve(n_geomodel * 2u + 1u);
        for (uint32_t g = 0; g < n_geomodel; ++g) {
            std::string key = species_key(geomodel_labels[g]);
            auto it = classifier_by_key.find(key);
            if (it != classifier_by_key.end()) by_species_key[std::move(key)] = it->second;
        }
        // GeomodelScores::project — keyed by the classifier label STRING: every mapped label starts at 0.0, reported
        // scores overwrite in order (:140-157)
        std::unordered_map<std::string, float> by_classifier_label;
        by_classifier_label.reserve(by_species_key.size() * 2u + 1u);
        for (const auto& kv : by_species_key) by_classifier_label[classifier_labels[kv.second]] = 0.0f;
        for (uint32_t j = 0; j < n_scores; ++j) {
            auto it = by_species_key.find(species_key(score_species[j]));
            if (it != by_species_key.end()) by_classifier_label[classifier_labels[it->second]] = score_values[j];
        }
        // score_of(label) per classifier row (:161-163)
        for (uint32_t i = 0; i < n_classifier; ++i) {
            auto it = by_classifier_label.find(classifier_labels[i]);
            mask[i] = it != by_classifier_label.end() ? it->second : std::numeric_limits<float>::quiet_NaN();
        }
        const uint32_t m = (uint32_t)by_species_key.size();
        if (mapped) *mapped = m;
        if (unmatched) *unmatched = n_classifier > m ? n_classifier - m : 0;     // saturating_sub (:107-110)
        return BB_OK;
    BB_CATCH(nullptr)
}

}  // extern "C"
