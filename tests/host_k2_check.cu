// Build-time self check of the K2 fast-path logic: runs the SAME templated code the device kernel
// runs (k2_fast.cuh), lane by lane on the CPU, against a naive double-precision DFT pipeline.
// It is a test program, not a fallback: nothing in the library calls it.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <complex>
#include <cmath>
#include "../birda_b200/csrc/k2_fast.cuh"

using namespace bb;
using namespace bb::k2f;
typedef std::complex<double> cd;

struct HostExec {
    static constexpr int kSlots = 32;
    template <class F> static void each(F&& f) { for (int l = 0; l < 32; ++l) f(l, l); }
};

template <class PL>
int check(const char* name) {
    constexpr int N = PL::N, M = PL::M, NKEEP = PL::NKEEP;
    const double pi = 3.14159265358979323846;
    int fwd[16], inv[16];
    for (int i = 0; i < PL::Fwd::count; ++i) fwd[i] = PL::Fwd::at(i);
    for (int i = 0; i < PL::Inv::count; ++i) inv[i] = PL::Inv::at(i);
    std::vector<uint16_t> pos_f(N), pos_i(M);
    build_pos_tables(fwd, PL::Fwd::count, inv, PL::Inv::count, N, M, pos_f.data(), pos_i.data());
    // a smooth low-pass-ish random filter of N taps, spectrum by naive DFT (double)
    srand(1234 + N);
    std::vector<double> h(N);
    for (int n = 0; n < N; ++n) h[n] = (rand() / (double)RAND_MAX - 0.5) * exp(-0.5 * pow((n - N / 2) / (N / 8.0), 2)) / N;
    std::vector<cd> H(N + 1);
    for (int k = 0; k <= N; ++k) { cd s = 0; for (int n = 0; n < N; ++n) s += h[n] * std::polar(1.0, -pi * k * n / N); H[k] = s; }
    std::vector<float> fre(NKEEP), fim(NKEEP);
    for (int k = 0; k < NKEEP; ++k) { fre[k] = (float)H[k].real(); fim[k] = (float)H[k].imag(); }
    std::vector<float2> Pt(NKEEP), Qt(NKEEP), WI(M / 2 + 1), twf(PL::TWF), twi(PL::TWI);
    build_split_tables(N, M, NKEEP, fre.data(), fim.data(), Pt.data(), Qt.data(), WI.data());
    for (int k = 0; k < PL::TWF; ++k) twf[k] = make_float2((float)cos(-2 * pi * k / N), (float)sin(-2 * pi * k / N));
    for (int k = 0; k < PL::TWI; ++k) twi[k] = make_float2((float)cos(2 * pi * k / M), (float)sin(2 * pi * k / M));
    Tables<PL> T{twf.data(), twi.data(), pos_f.data(), pos_i.data(), Pt.data(), Qt.data(), WI.data()};

    const int NB = 3;
    std::vector<float> x(NB * N);
    for (auto& v : x) v = (float)(rand() / (double)RAND_MAX - 0.5);
    const int valid_last = N - 77;                 // last block partially valid
    std::vector<float2> A(N), B(M);
    std::vector<LaneCarry<PL>> carry(32);
    for (auto& c : carry) for (int i = 0; i < PL::CARRY_ITERS; ++i) for (int j = 0; j < PL::QL; ++j) c.c[i][j] = 0.f;
    std::vector<float> out(NB * M, 0.f);
    for (int b = 0; b < NB; ++b) {
        const int valid = b + 1 == NB ? valid_last : N;
        auto loader = [&](int n) {
            float re = 2 * n < valid ? x[b * N + 2 * n] : 0.f, im = 2 * n + 1 < valid ? x[b * N + 2 * n + 1] : 0.f;
            return make_float2(re, im);
        };
        auto sinks = [&](int) { return [&, b](int n, float2 y) { out[b * M + 2 * n] = y.x; out[b * M + 2 * n + 1] = y.y; }; };
        process_block<PL, HostExec>(T, A.data(), B.data(), loader, carry.data(), sinks);
    }
    // reference
    std::vector<double> ref(NB * M + M, 0.0);
    for (int b = 0; b < NB; ++b) {
        const int valid = b + 1 == NB ? valid_last : N;
        std::vector<cd> X(N + 1), Y(M + 1, cd(0, 0));
        for (int k = 0; k <= N; ++k) { cd s = 0; for (int n = 0; n < valid; ++n) s += (double)x[b * N + n] * std::polar(1.0, -pi * k * n / N); X[k] = s; }
        for (int k = 0; k < NKEEP; ++k) Y[k] = X[k] * cd(fre[k], fim[k]);
        for (int n = 0; n < 2 * M; ++n) {
            double s = Y[0].real() + Y[M].real() * ((n & 1) ? -1.0 : 1.0);
            for (int k = 1; k < M; ++k) s += 2.0 * (Y[k] * std::polar(1.0, pi * k * n / M)).real();
            ref[b * M + n] += s;
        }
    }
    double maxerr = 0, rms = 0;
    for (int i = 0; i < NB * M; ++i) { maxerr = fmax(maxerr, fabs(out[i] - ref[i])); rms += ref[i] * ref[i]; }
    rms = sqrt(rms / (NB * M));
    printf("%-28s N=%d M=%d  max err %.3e  rms %.3e  rel %.3e\n", name, N, M, maxerr, rms, maxerr / rms);
    return maxerr / rms < 5e-6 ? 0 : 1;
}

#include "../birda_b200/csrc/k2_plans.cuh"

int main() {
    int bad = 0;
#define BB_PLAN(NAME, FROM, TO, ...) bad += check<__VA_ARGS__>(#NAME);
    BB_K2_PLANS(BB_PLAN)
#undef BB_PLAN
    printf(bad ? "FAILED\n" : "all plans ok\n");
    return bad;
}
