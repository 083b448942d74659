// Compile-time resampler plans of the K2 fast path: (source rate family -> target rate) block
// sizes as rubato derives them (src/audio/resample.rs:19-30) with their FFT factorisations.
// X(NAME, N_IN, N_OUT, Plan<RSeq<forward DIF radices>, RSeq<inverse DIT radices, last one even>>)
#pragma once
#include "k2_fast.cuh"

#define BB_K2_PLANS(X)                                                                                     \
    X(p1029_1120, 1029, 1120, bb::k2f::Plan<bb::k2f::RSeq<7, 7, 7, 3>, bb::k2f::RSeq<7, 5, 4, 8>>)   /* 44.1k -> 48k */ \
    X(p1029_2240, 1029, 2240, bb::k2f::Plan<bb::k2f::RSeq<7, 7, 7, 3>, bb::k2f::RSeq<7, 5, 8, 8>>)   /* 22.05k -> 48k */ \
    X(p1026_684, 1026, 684, bb::k2f::Plan<bb::k2f::RSeq<6, 19, 9>, bb::k2f::RSeq<19, 9, 4>>)         /* 48k -> 32k, 96k -> 64k.. */ \
    X(p1024_512, 1024, 512, bb::k2f::Plan<bb::k2f::RSeq<8, 8, 16>, bb::k2f::RSeq<8, 8, 8>>)          /* 96k -> 48k */ \
    X(p1024_1536, 1024, 1536, bb::k2f::Plan<bb::k2f::RSeq<8, 8, 16>, bb::k2f::RSeq<3, 8, 8, 8>>)     /* 32k -> 48k */ \
    X(p1024_3072, 1024, 3072, bb::k2f::Plan<bb::k2f::RSeq<8, 8, 16>, bb::k2f::RSeq<3, 8, 8, 16>>)    /* 16k -> 48k */ \
    X(p1323_960, 1323, 960, bb::k2f::Plan<bb::k2f::RSeq<7, 7, 9, 3>, bb::k2f::RSeq<5, 3, 8, 8>>)     /* 44.1k -> 32k */
