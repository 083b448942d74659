// In-register DFT butterflies shared by the resampler kernels (product code), generic over the
// complex element type:
//   float2  — one complex value (re, im);
//   cx2     — TWO complex values from two independent streams held as re = (re0, re1), im = (im0, im1),
//             so every arithmetic instruction is a packed f32x2 op (FADD2 / FMUL2 / FFMA2 on sm_100a:
//             half the issue slots per butterfly; negation and scalar immediates are free modifiers).
// Host-callable too so tests/host_k2_check.cu can run the exact device logic on the CPU
// (a build-time self check; it is not a fallback path).
#pragma once
#include <cuda_runtime.h>
#include "dft_consts.cuh"

#ifndef BB_HD
#define BB_HD __host__ __device__ __forceinline__
#endif

namespace bb {

// ---- component level: K = float (scalar) or float2 (two streams, packed)
BB_HD float kadd(float a, float b) { return a + b; }
BB_HD float ksub(float a, float b) { return a - b; }
BB_HD float kneg(float a) { return -a; }
BB_HD float kmul(float a, float b) { return a * b; }
BB_HD float kmulc(float a, float c) { return a * c; }
BB_HD float kfma(float a, float b, float c) { return fmaf(a, b, c); }
BB_HD float kfmac(float a, float c, float acc) { return fmaf(a, c, acc); }
BB_HD float kzero(float) { return 0.f; }

BB_HD float2 kneg(float2 a) { return make_float2(-a.x, -a.y); }
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
BB_HD float2 kadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
BB_HD float2 ksub(float2 a, float2 b) { return __fadd2_rn(a, kneg(b)); }
BB_HD float2 kmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
BB_HD float2 kmulc(float2 a, float c) { return __fmul2_rn(a, make_float2(c, c)); }
BB_HD float2 kfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
BB_HD float2 kfmac(float2 a, float c, float2 acc) { return __ffma2_rn(a, make_float2(c, c), acc); }
#else
BB_HD float2 kadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
BB_HD float2 ksub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
BB_HD float2 kmul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
BB_HD float2 kmulc(float2 a, float c) { return make_float2(a.x * c, a.y * c); }
BB_HD float2 kfma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
BB_HD float2 kfmac(float2 a, float c, float2 acc) { return make_float2(fmaf(a.x, c, acc.x), fmaf(a.y, c, acc.y)); }
#endif
BB_HD float2 kzero(float2) { return make_float2(0.f, 0.f); }

// ---- complex level
struct cx2 { float2 re, im; };       // two streams; stored in shared memory as float4 (re0, re1, im0, im1)

template <class C> struct Cx;
template <> struct Cx<float2> {
    using K = float;
    static BB_HD float re(float2 a) { return a.x; }
    static BB_HD float im(float2 a) { return a.y; }
    static BB_HD float2 make(float r, float i) { return make_float2(r, i); }
};
template <> struct Cx<cx2> {
    using K = float2;
    static BB_HD float2 re(const cx2& a) { return a.re; }
    static BB_HD float2 im(const cx2& a) { return a.im; }
    static BB_HD cx2 make(float2 r, float2 i) { cx2 c; c.re = r; c.im = i; return c; }
};
#define BB_RE(a) Cx<C>::re(a)
#define BB_IM(a) Cx<C>::im(a)
#define BB_MK(r, i) Cx<C>::make(r, i)

template <class C> BB_HD C czero() { typename Cx<C>::K z = kzero(typename Cx<C>::K()); return Cx<C>::make(z, z); }
template <class C> BB_HD C cadd(C a, C b) { return BB_MK(kadd(BB_RE(a), BB_RE(b)), kadd(BB_IM(a), BB_IM(b))); }
template <class C> BB_HD C csub(C a, C b) { return BB_MK(ksub(BB_RE(a), BB_RE(b)), ksub(BB_IM(a), BB_IM(b))); }
template <class C> BB_HD C cmul(C a, C b) {
    return BB_MK(kfma(BB_RE(a), BB_RE(b), kneg(kmul(BB_IM(a), BB_IM(b)))), kfma(BB_RE(a), BB_IM(b), kmul(BB_IM(a), BB_RE(b))));
}
template <class C> BB_HD C cconj(C a) { return BB_MK(BB_RE(a), kneg(BB_IM(a))); }
// multiply by -i (forward) / +i (inverse)
template <bool INV, class C> BB_HD C rot90(C a) { return INV ? BB_MK(kneg(BB_IM(a)), BB_RE(a)) : BB_MK(BB_IM(a), kneg(BB_RE(a))); }
// a * (c + i s) for literal c, s
template <class C> BB_HD C cmul_const(C a, float c, float s) {
    return BB_MK(kfmac(BB_RE(a), c, kmulc(BB_IM(a), -s)), kfmac(BB_RE(a), s, kmulc(BB_IM(a), c)));
}
// acc + a * c (real literal c)
template <class C> BB_HD C cfma_real(C a, float c, C acc) { return BB_MK(kfmac(BB_RE(a), c, BB_RE(acc)), kfmac(BB_IM(a), c, BB_IM(acc))); }
template <class C> BB_HD C cmul_real(C a, float c) { return BB_MK(kmulc(BB_RE(a), c), kmulc(BB_IM(a), c)); }

template <int R, bool INV> struct Dft;

template <bool INV> struct Dft<2, INV> {
    template <class C> static BB_HD void run(C (&a)[2]) { C t = a[0]; a[0] = cadd(t, a[1]); a[1] = csub(t, a[1]); }
};
template <bool INV> struct Dft<4, INV> {
    template <class C> static BB_HD void run(C (&a)[4]) {
        C t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]), t2 = cadd(a[1], a[3]), t3 = rot90<INV>(csub(a[1], a[3]));
        a[0] = cadd(t0, t2); a[2] = csub(t0, t2); a[1] = cadd(t1, t3); a[3] = csub(t1, t3);
    }
};
template <bool INV> struct Dft<8, INV> {
    template <class C> static BB_HD void run(C (&a)[8]) {
        constexpr float h = 0.70710678118654752440f;
        C e[4] = {a[0], a[2], a[4], a[6]}, o[4] = {a[1], a[3], a[5], a[7]};
        Dft<4, INV>::run(e); Dft<4, INV>::run(o);
        // o[k] *= w8^k ; w8 = exp(-+ i pi/4) = h (1 -+ i)
        const C o1 = INV ? BB_MK(kmulc(ksub(BB_RE(o[1]), BB_IM(o[1])), h), kmulc(kadd(BB_RE(o[1]), BB_IM(o[1])), h))
                         : BB_MK(kmulc(kadd(BB_RE(o[1]), BB_IM(o[1])), h), kmulc(ksub(BB_IM(o[1]), BB_RE(o[1])), h));
        const C o2 = rot90<INV>(o[2]);
        const C o3 = INV ? BB_MK(kmulc(kadd(BB_RE(o[3]), BB_IM(o[3])), -h), kmulc(ksub(BB_RE(o[3]), BB_IM(o[3])), h))
                         : BB_MK(kmulc(ksub(BB_IM(o[3]), BB_RE(o[3])), h), kmulc(kadd(BB_RE(o[3]), BB_IM(o[3])), -h));
        a[0] = cadd(e[0], o[0]); a[4] = csub(e[0], o[0]);
        a[1] = cadd(e[1], o1);   a[5] = csub(e[1], o1);
        a[2] = cadd(e[2], o2);   a[6] = csub(e[2], o2);
        a[3] = cadd(e[3], o3);   a[7] = csub(e[3], o3);
    }
};

// odd prime R: pair a[j] with a[R-j]
template <int R, bool INV> struct DftOddPrime {
    template <class C> static BB_HD void run(C (&a)[R]) {
        constexpr int H = (R - 1) / 2;
        C sp[H + 1], sm[H + 1];
#pragma unroll
        for (int j = 1; j <= H; ++j) { sp[j] = cadd(a[j], a[R - j]); sm[j] = csub(a[j], a[R - j]); }
        const C a0 = a[0];
        C b0 = a[0];
#pragma unroll
        for (int j = 1; j <= H; ++j) b0 = cadd(b0, sp[j]);
        a[0] = b0;
#pragma unroll
        for (int k = 1; k <= H; ++k) {
            C u = cfma_real(sp[1], DftConst<R>::c(k % R), a0);
            C v = cmul_real(sm[1], DftConst<R>::s(k % R));
#pragma unroll
            for (int j = 2; j <= H; ++j) {
                u = cfma_real(sp[j], DftConst<R>::c((j * k) % R), u);
                v = cfma_real(sm[j], DftConst<R>::s((j * k) % R), v);
            }
            // forward: b_k = u - i v, b_{R-k} = u + i v ; inverse swaps them
            const C miv = rot90<false>(v);
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

// composite R = R1*R2 in registers: j = R2*j1 + j2 -> k = k1 + R1*k2, literal inner twiddles W_R^(k1*j2)
template <int R1, int R2, bool INV> struct DftComposite {
    static constexpr int R = R1 * R2;
    template <class C> static BB_HD void run(C (&a)[R]) {
        C t[R];
#pragma unroll
        for (int j2 = 0; j2 < R2; ++j2) {
            C col[R1];
#pragma unroll
            for (int j1 = 0; j1 < R1; ++j1) col[j1] = a[R2 * j1 + j2];
            Dft<R1, INV>::run(col);
#pragma unroll
            for (int k1 = 0; k1 < R1; ++k1) {
                const int e = (k1 * j2) % R;
                if (e == 0) t[k1 * R2 + j2] = col[k1];
                else t[k1 * R2 + j2] = cmul_const(col[k1], DftConst<R>::c(e), INV ? DftConst<R>::s(e) : -DftConst<R>::s(e));
            }
        }
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            C row[R2];
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
template <bool INV> struct Dft<10, INV> : DftComposite<2, 5, INV> {};
template <bool INV> struct Dft<12, INV> : DftComposite<4, 3, INV> {};
template <bool INV> struct Dft<14, INV> : DftComposite<2, 7, INV> {};
template <bool INV> struct Dft<15, INV> : DftComposite<3, 5, INV> {};
template <bool INV> struct Dft<18, INV> : DftComposite<2, 9, INV> {};
template <bool INV> struct Dft<20, INV> : DftComposite<4, 5, INV> {};
template <bool INV> struct Dft<21, INV> : DftComposite<3, 7, INV> {};

}  // namespace bb
