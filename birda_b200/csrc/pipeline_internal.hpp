// Library-internal entry points of the per-file pipeline (product code; not part of the C ABI).
#pragma once
#include <cstdint>
#include <vector>
#include "../../include/birda_b200.h"

namespace bb {
// bytes [offset, offset + bytes) of a file into dst, read by `threads` threads at once (wav.cpp)
int32_t read_range_parallel(const char* path, uint64_t offset, uint64_t bytes, void* dst, uint32_t threads);
// bb_pipeline_process_wav with a growing detection list (bb_pool: no capacity guess, no second pass over the file).
// Detections come back sorted by (start_time asc, confidence desc).  Status codes as the C entry point.
int32_t pipeline_process_wav_into(bb_pipeline* p, const char* path, uint64_t piece_frames, std::vector<bb_detection>* out,
                                  uint64_t* n_segments, uint32_t* batch_used);
}
