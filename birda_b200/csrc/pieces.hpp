// Piece table of a streamed file (csrc/pipeline.cpp: process_wav): a file too long for one staging buffer goes through
// the front end in pieces, cut so that every piece but the last holds a whole number of batches of FULL windows — only
// the file's LAST batch is padded, as in the reference's loop (processor.rs:132-170) — and the next piece starts at the
// hop after the last window of this one.  Pure arithmetic on the header; tests/host_pieces_check.cpp checks it on the CPU.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace bb {

struct Piece { uint64_t pos, frames; bool eof; };       // first frame, frames handed to the front end, last piece of the file

// piece_frames >= 2 * src_seg, hop = src_seg - src_ovl >= 1, B = batch size >= 1
inline std::vector<Piece> plan_pieces(uint64_t total_frames, uint64_t piece_frames, uint64_t src_seg, uint64_t hop, uint32_t B) {
    std::vector<Piece> pieces;
    for (uint64_t pos = 0;;) {
        uint64_t want = std::min<uint64_t>(piece_frames, total_frames - pos);
        bool eof = pos + want >= total_frames;
        if (!eof) {                                                                // trim to k*B full windows
            uint64_t nfull = want >= src_seg ? (want - src_seg) / hop + 1 : 0;
            nfull = nfull / B * B;
            if (nfull == 0) { want = std::min<uint64_t>(total_frames - pos, src_seg + (uint64_t)(B - 1) * hop); eof = pos + want >= total_frames; }
            else want = (nfull - 1) * hop + src_seg;
        }
        pieces.push_back({pos, want, eof});
        if (eof) break;
        const uint64_t nwin = (want - src_seg) / hop + 1;                         // windows of a non-final piece: all full
        pos += nwin * hop;
    }
    return pieces;
}

}  // namespace bb
