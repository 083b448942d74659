// CPU check of the piece table of a streamed file (csrc/pieces.hpp, used by csrc/pipeline.cpp): test program, not a fallback.
#include <cstdio>
#include <cstdint>
#include "../birda_b200/csrc/pieces.hpp"

int main() {
    long cases = 0, bad = 0;
    const uint64_t segs[] = {132300, 144000, 160000, 48000, 7};
    for (uint64_t src_seg : segs) for (uint64_t ovl_num : {(uint64_t)0, (uint64_t)1, (uint64_t)2, (uint64_t)3}) for (uint32_t B : {1u, 8u, 64u}) for (uint64_t pf_mul : {(uint64_t)2, (uint64_t)3, (uint64_t)9, (uint64_t)40, (uint64_t)500})
    for (uint64_t total : {(uint64_t)0, (uint64_t)1, src_seg - 1, src_seg, src_seg + 1, 3 * src_seg, 10 * src_seg + 17, 65 * src_seg, 1000 * src_seg + 5, (uint64_t)158760000}) {
        const uint64_t src_ovl = src_seg * ovl_num / 4;       // 0, 1/4, 1/2, 3/4 of the window
        const uint64_t hop = src_seg - src_ovl;
        const uint64_t piece_frames = pf_mul * src_seg;                               // process_wav enforces >= 2 * src_seg
        const auto pieces = bb::plan_pieces(total, piece_frames, src_seg, hop, B);
        ++cases;
        bool ok = !pieces.empty() && pieces.front().pos == 0 && pieces.back().eof;
        uint64_t windows = 0;
        for (size_t k = 0; ok && k < pieces.size(); ++k) {
            const bb::Piece& p = pieces[k];
            ok = ok && p.pos == windows * hop && p.pos + p.frames <= total;
            if (k + 1 < pieces.size()) {
                // a non-final piece: whole batches of full windows, nothing left over inside it
                ok = ok && !p.eof && p.frames >= src_seg && (p.frames - src_seg) % hop == 0;
                const uint64_t n = (p.frames - src_seg) / hop + 1;
                ok = ok && n % B == 0 && n >= B;
                windows += n;
            } else {
                ok = ok && p.eof && p.pos + p.frames == total;                         // the last piece runs to the end of the file
            }
        }
        // staging buffers are sized for the largest piece: bounded by the request (or one batch of windows, if that is larger)
        for (const bb::Piece& p : pieces) {
            const uint64_t one_batch = src_seg + (uint64_t)(B - 1) * hop;
            ok = ok && p.frames <= (piece_frames > one_batch ? piece_frames : one_batch);
        }
        if (!ok) { ++bad; if (bad < 10) printf("FAILED: total %llu piece %llu seg %llu ovl %llu B %u\n", (unsigned long long)total, (unsigned long long)piece_frames, (unsigned long long)src_seg, (unsigned long long)src_ovl, B); }
    }
    printf("%ld piece tables, %ld bad\n", cases, bad);
    return bad ? 1 : 0;
}
