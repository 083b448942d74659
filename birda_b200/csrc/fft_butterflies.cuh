// In-register DFT butterflies shared by the resampler kernels (product code).
// Host-callable too so tests/host_k2_check.cu can run the exact device logic on the CPU
// (as a build-time self check; it is not a fallback path).
#pragma once
#include <cuda_runtime.h>
#include "dft_consts.cuh"

#ifndef BB_HD
#define BB_HD __host__ __device__ __forceinline__
#endif

namespace bb {

BB_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
BB_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
BB_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
BB_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by -i (forward) / +i (inverse)
template <bool INV> BB_HD float2 rot90(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

template <int R, bool INV> struct Dft;

template <bool INV> struct Dft<2, INV> {
    static BB_HD void run(float2 (&a)[2]) { float2 t = a[0]; a[0] = cadd(t, a[1]); a[1] = csub(t, a[1]); }
};
template <bool INV> struct Dft<4, INV> {
    static BB_HD void run(float2 (&a)[4]) {
        float2 t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]), t2 = cadd(a[1], a[3]), t3 = rot90<INV>(csub(a[1], a[3]));
        a[0] = cadd(t0, t2); a[2] = csub(t0, t2); a[1] = cadd(t1, t3); a[3] = csub(t1, t3);
    }
};
template <bool INV> struct Dft<8, INV> {
    static BB_HD void run(float2 (&a)[8]) {
        constexpr float h = 0.70710678118654752440f;
        float2 e[4] = {a[0], a[2], a[4], a[6]}, o[4] = {a[1], a[3], a[5], a[7]};
        Dft<4, INV>::run(e); Dft<4, INV>::run(o);
        // o[k] *= w8^k ; w8 = exp(-+ i pi/4)
        float2 o1 = INV ? make_float2((o[1].x - o[1].y) * h, (o[1].x + o[1].y) * h) : make_float2((o[1].x + o[1].y) * h, (o[1].y - o[1].x) * h);
        float2 o2 = rot90<INV>(o[2]);
        float2 o3 = INV ? make_float2((-o[3].x - o[3].y) * h, (o[3].x - o[3].y) * h) : make_float2((o[3].y - o[3].x) * h, (-o[3].x - o[3].y) * h);
        a[0] = cadd(e[0], o[0]); a[4] = csub(e[0], o[0]);
        a[1] = cadd(e[1], o1);   a[5] = csub(e[1], o1);
        a[2] = cadd(e[2], o2);   a[6] = csub(e[2], o2);
        a[3] = cadd(e[3], o3);   a[7] = csub(e[3], o3);
    }
};

// odd prime R: pair a[j] with a[R-j]
template <int R, bool INV> struct DftOddPrime {
    static BB_HD void run(float2 (&a)[R]) {
        constexpr int H = (R - 1) / 2;
        float2 sp[H + 1], sm[H + 1];
#pragma unroll
        for (int j = 1; j <= H; ++j) { sp[j] = cadd(a[j], a[R - j]); sm[j] = csub(a[j], a[R - j]); }
        float2 a0 = a[0], b0 = a[0];
#pragma unroll
        for (int j = 1; j <= H; ++j) b0 = cadd(b0, sp[j]);
        a[0] = b0;
#pragma unroll
        for (int k = 1; k <= H; ++k) {
            float2 u = a0, v = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 1; j <= H; ++j) {
                const float c = DftConst<R>::c((j * k) % R), s = DftConst<R>::s((j * k) % R);
                u.x = fmaf(sp[j].x, c, u.x); u.y = fmaf(sp[j].y, c, u.y);
                v.x = fmaf(sm[j].x, s, v.x); v.y = fmaf(sm[j].y, s, v.y);
            }
            // forward: b_k = u - i v, b_{R-k} = u + i v ; inverse swaps them
            float2 miv = make_float2(v.y, -v.x);
            if (INV) { a[k] = csub(u, miv); a[R - k] = cadd(u, miv); }
            else     { a[k] = cadd(u, miv); a[R - k] = csub(u, miv); }
        }
    }
};
template <bool INV> struct Dft<3, INV>  : DftOddPrime<3, INV> {};
template <bool INV> struct Dft<5, INV>  : DftOddPrime<5, INV> {};
template <bool INV> struct Dft<7, INV>  : DftOddPrime<7, INV> {};
template <bool INV> struct Dft<11, INV> : DftOddPrime<11, INV> {};
template <bool INV> struct Dft<13, INV> : DftOddPrime<13, INV> {};
template <bool INV> struct Dft<17, INV> : DftOddPrime<17, INV> {};
template <bool INV> struct Dft<19, INV> : DftOddPrime<19, INV> {};
template <bool INV> struct Dft<23, INV> : DftOddPrime<23, INV> {};
template <bool INV> struct Dft<29, INV> : DftOddPrime<29, INV> {};
template <bool INV> struct Dft<31, INV> : DftOddPrime<31, INV> {};

// composite R = R1*R2 in registers: j = R2*j1 + j2 -> k = k1 + R1*k2, constant inner twiddles W_R^(k1*j2)
template <int R1, int R2, bool INV> struct DftComposite {
    static constexpr int R = R1 * R2;
    static BB_HD void run(float2 (&a)[R]) {
        float2 t[R];
#pragma unroll
        for (int j2 = 0; j2 < R2; ++j2) {
            float2 col[R1];
#pragma unroll
            for (int j1 = 0; j1 < R1; ++j1) col[j1] = a[R2 * j1 + j2];
            Dft<R1, INV>::run(col);
#pragma unroll
            for (int k1 = 0; k1 < R1; ++k1) {
                const int e = (k1 * j2) % R;
                if (e == 0) t[k1 * R2 + j2] = col[k1];
                else {
                    const float c = DftConst<R>::c(e), s = INV ? DftConst<R>::s(e) : -DftConst<R>::s(e);
                    t[k1 * R2 + j2] = make_float2(col[k1].x * c - col[k1].y * s, col[k1].x * s + col[k1].y * c);
                }
            }
        }
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            float2 row[R2];
#pragma unroll
            for (int j2 = 0; j2 < R2; ++j2) row[j2] = t[k1 * R2 + j2];
            Dft<R2, INV>::run(row);
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) a[k1 + R1 * k2] = row[k2];
        }
    }
};
template <bool INV> struct Dft<6, INV>  : DftComposite<2, 3, INV> {};
template <bool INV> struct Dft<9, INV>  : DftComposite<3, 3, INV> {};
template <bool INV> struct Dft<16, INV> : DftComposite<4, 4, INV> {};

}  // namespace bb
