// Build-time self check of the K2 warp-per-block core (k2_warp.cuh): runs the SAME code the device
// kernel runs, lane by lane on the CPU, against a naive double-precision DFT pipeline.
// It is a test program, not a fallback: nothing in the library calls it.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <complex>
#include <cmath>
#include <cstring>
#include <type_traits>
#include "../birda_b200/csrc/k2_warp.cuh"

using namespace bb;
using namespace bb::k2w;
typedef std::complex<double> cd;

struct HostExec {
    int nl;     // lanes in the group that owns a block (32 = one warp, 64 = two warps, ...)
    template <class F> void each(F&& f) const { for (int l = 0; l < nl; ++l) f(l, nl); }
};

template <class CT, class C>
static int check_impl(int N, int M, int nl) {
    using MT = typename Mem<C>::T;
    constexpr int NS = std::is_same<C, cx2>::value ? 2 : 1;      // streams
    const double pi = 3.14159265358979323846;
    const int NKEEP = N < M ? N + 1 : M;
    RtPlan P; std::vector<int> fwd, inv;
    constexpr bool kCt = !std::is_same<CT, void>::value;
    if constexpr (kCt) {
        for (int i = 0; i < CT::Fwd::count; ++i) fwd.push_back(CT::Fwd::at(i));
        for (int i = 0; i < CT::Inv::count; ++i) inv.push_back(CT::Inv::at(i));
        if (!build_plan_from_radices(N, M, NKEEP, &P, &fwd, &inv)) { printf("bad ct plan\n"); return 1; }
    } else if (!build_plan(N, M, NKEEP, &P, &fwd, &inv)) { printf("N=%d M=%d: no plan (generic kernel)\n", N, M); return 0; }
    if constexpr (kCt) ct_plan_pads(&P); else rt_plan_pads(&P);
    std::vector<uint32_t> ordf, ordi;
    build_stage_orders(&P, NS == 2 ? 8 : 16, !kCt, &ordf, &ordi);
    if (kCt && NS == 1) { ordf.clear(); ordi.clear(); }      // single-stream kernels of compile-time plans use the default mapping
    std::vector<uint16_t> pos_f(N), pos_i(M);
    build_pos_tables(fwd, inv, N, M, pos_f.data(), pos_i.data());
    srand(1234 + N);
    std::vector<double> h(N);
    for (int n = 0; n < N; ++n) h[n] = (rand() / (double)RAND_MAX - 0.5) * exp(-0.5 * pow((n - N / 2) / (N / 8.0), 2)) / N;
    std::vector<cd> H(N + 1);
    for (int k = 0; k <= N; ++k) { cd s = 0; for (int n = 0; n < N; ++n) s += h[n] * std::polar(1.0, -pi * k * n / N); H[k] = s; }
    std::vector<float> fre(NKEEP), fim(NKEEP);
    for (int k = 0; k < NKEEP; ++k) { fre[k] = (float)H[k].real(); fim[k] = (float)H[k].imag(); }
    std::vector<float2> Pt(NKEEP), Qt(NKEEP), WI(M / 2 + 1), twf(P.twf_len), twi(P.twi_len);
    build_split_tables(N, M, NKEEP, fre.data(), fim.data(), Pt.data(), Qt.data(), WI.data());
    build_twiddles(P, twf.data(), twi.data());
    auto twf_e = expand_table<C>(twf), twi_e = expand_table<C>(twi);
    SplitLayout SL;
    build_split_layout(N, M, NKEEP, pos_f.data(), pos_i.data(), Pt.data(), Qt.data(), WI.data(), NS == 2 ? 8 : 16, &SL, P.pad_a, P.pad_b);
    Tables<C> T{twf_e.data(), twi_e.data(), SL.sidx.data(), SL.pq1.data(), SL.pq2.data(), SL.wi.data(),
                ordf.empty() ? nullptr : ordf.data(), ordi.empty() ? nullptr : ordi.data()};

    const int NB = 3;
    std::vector<float> x(NB * N), x2(NB * N);
    for (auto& v : x) v = (float)(rand() / (double)RAND_MAX - 0.5);
    for (auto& v : x2) v = (float)(rand() / (double)RAND_MAX - 0.5);
    const int valid_last = N - 77;                 // last block partially valid
    std::vector<MT> A(phys_len(N, P.pad_a)), B(phys_len(M, P.pad_b)), carry(M / 2);
    memset(carry.data(), 0, sizeof(MT) * carry.size());
    std::vector<float> out(NB * M, 0.f), out2(NB * M, 0.f);
    for (int b = 0; b < NB; ++b) {
        const int valid = b + 1 == NB ? valid_last : N;
        auto smp = [&](const std::vector<float>& v, int i) { return i < valid ? v[b * N + i] : 0.f; };
        auto loader = [&](int n) {
            if constexpr (NS == 1) return make_float2(smp(x, 2 * n), smp(x, 2 * n + 1));
            else { cx2 c; c.re = make_float2(smp(x, 2 * n), smp(x2, 2 * n)); c.im = make_float2(smp(x, 2 * n + 1), smp(x2, 2 * n + 1)); return c; }
        };
        auto sink = [&](int n, C y) {
            if constexpr (NS == 1) { out[b * M + 2 * n] = y.x; out[b * M + 2 * n + 1] = y.y; }
            else { out[b * M + 2 * n] = y.re.x; out[b * M + 2 * n + 1] = y.im.x; out2[b * M + 2 * n] = y.re.y; out2[b * M + 2 * n + 1] = y.im.y; }
        };
        if constexpr (kCt) process_block_ct<CT, C>(HostExec{nl}, T, A.data(), B.data(), carry.data(), loader, sink, [] {});
        else if (P.pad_a) process_block<C, MapPad8>(HostExec{nl}, P, T, A.data(), B.data(), carry.data(), loader, sink, [] {});
        else process_block<C>(HostExec{nl}, P, T, A.data(), B.data(), carry.data(), loader, sink, [] {});
    }
    double worst = 0;
    for (int stream = 0; stream < NS; ++stream) {
    const std::vector<float>& xs = stream ? x2 : x;
    const std::vector<float>& outs = stream ? out2 : out;
    std::vector<double> ref(NB * M + M, 0.0);
    for (int b = 0; b < NB; ++b) {
        const int valid = b + 1 == NB ? valid_last : N;
        std::vector<cd> X(N + 1), Y(M + 1, cd(0, 0));
        for (int k = 0; k <= N; ++k) { cd s = 0; for (int n = 0; n < valid; ++n) s += (double)xs[b * N + n] * std::polar(1.0, -pi * k * n / N); X[k] = s; }
        for (int k = 0; k < NKEEP; ++k) Y[k] = X[k] * cd(fre[k], fim[k]);
        for (int n = 0; n < 2 * M; ++n) {
            double s = Y[0].real() + Y[M].real() * ((n & 1) ? -1.0 : 1.0);
            for (int k = 1; k < M; ++k) s += 2.0 * (Y[k] * std::polar(1.0, pi * k * n / M)).real();
            ref[b * M + n] += s;
        }
    }
    double maxerr = 0, rms = 0;
    for (int i = 0; i < NB * M; ++i) { maxerr = fmax(maxerr, fabs(outs[i] - ref[i])); rms += ref[i] * ref[i]; }
    rms = sqrt(rms / (NB * M));
    worst = fmax(worst, maxerr / rms);
    }
    printf("%s x%d N=%4d M=%4d fwd[", kCt ? "ct" : "rt", NS, N, M);
    for (int r : fwd) printf(" %d", r);
    printf(" ] inv[");
    for (int r : inv) printf(" %d", r);
    printf(" ] pad %d%d orders %zu+%zu split conflicts %d  rel err %.3e\n", P.pad_a, P.pad_b, ordf.size(), ordi.size(), SL.extra_wavefronts, worst);
    // the split layout must keep the scattered accesses mostly off each other's banks: at most a third more wavefronts
    // than the conflict-free count (6 scattered accesses per lane group) for the two-stream compile-time plans
    const int ideal = 6 * ((M / 2 + 1 + (NS == 2 ? 8 : 16) - 1) / (NS == 2 ? 8 : 16));
    const bool layout_ok = !(kCt && NS == 2) || SL.extra_wavefronts * 3 <= ideal;
    if (!layout_ok) printf("split layout too conflicted: %d extra of %d\n", SL.extra_wavefronts, ideal);
    return (worst < 5e-6 && layout_ok) ? 0 : 1;
}

// every (row, block) is covered exactly once by the work items, runs are contiguous and end flags are right
static int check_work_items() {
    int bad = 0;
    const uint64_t units_l[] = {1, 2, 7, 75, 100, 360, 591, 592, 593, 1200, 1216, 5000};
    const uint64_t groups_l[] = {1, 148, 592, 888, 1776};
    const uint32_t nblk_l[] = {1, 7, 31, 32, 129, 234, 282};
    for (uint64_t units : units_l) for (uint64_t groups : groups_l) for (uint32_t nblk : nblk_l) for (int ov = 0; ov < 2; ++ov) for (int force : {0, 5}) {
        const WorkItems w = plan_work_items(units, groups, nblk, ov != 0, force);
        std::vector<uint32_t> next(units, 0); std::vector<char> ended(units, 0);
        uint64_t prev_row = 0;
        for (uint64_t item = 0; item < w.nitems; ++item) {
            uint64_t row; uint32_t b0, b1; bool last;
            decode_work_item(w, item, row, b0, b1, last);
            if (row >= units || row < prev_row || b0 != next[row] || b1 <= b0 || b1 > nblk || ended[row]) { ++bad; break; }
            next[row] = b1; prev_row = row;
            if (last != (b1 == nblk)) { ++bad; break; }
            if (last) ended[row] = 1;
        }
        for (uint64_t r = 0; r < units; ++r) if (next[r] != nblk || !ended[r]) { ++bad; break; }
        if (force == 0 && groups > 1 && units >= groups && (w.n1_rows % groups != 0 || w.n1_rows + groups <= units)) ++bad;   // phase 1 = whole rounds
    }
    printf("work items: %s\n", bad ? "FAILED" : "every (row, block) covered once");
    return bad;
}

int main() {
    int bad = check_work_items();
    const int cases[][2] = {{1029, 1120}, {1029, 2240}, {1026, 684}, {1024, 512}, {1024, 1536}, {1024, 3072},
                            {1323, 960}, {1323, 1920}, {1024, 2048}, {1026, 342}, {1029, 560}, {1125, 216}, {1024, 256}};
    for (auto& c : cases) for (int nl : {32, 64}) bad += check_impl<void, float2>(c[0], c[1], nl);
    for (auto& c : cases) bad += check_impl<void, cx2>(c[0], c[1], 96);
#define BB_CT(NAME, NI, NO, TH, ...) bad += check_impl<__VA_ARGS__, float2>(NI, NO, 64); bad += check_impl<__VA_ARGS__, cx2>(NI, NO, 96); bad += check_impl<__VA_ARGS__, cx2>(NI, NO, 160);
    BB_K2_CT_PLANS(BB_CT)
#undef BB_CT
    printf(bad ? "FAILED\n" : "all plans ok\n");
    return bad;
}
