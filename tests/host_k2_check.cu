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

// N, M: complex transform lengths; the filter has `taps` taps; a block consumes AI samples and emits AO
// (rubato's blocking: taps = AI = N, AO = M; a plan with its own blocking: AI + taps - 1 <= 2N)
template <class CT, class C>
static int check_impl(int N, int M, int nl, int AI = 0, int AO = 0, int taps = 0) {
    using MT = typename Mem<C>::T;
    if (AI == 0) AI = N;
    if (AO == 0) AO = M;
    if (taps == 0) taps = N;
    constexpr int NS = std::is_same<C, cx2>::value ? 2 : 1;      // streams
    const double pi = 3.14159265358979323846;
    const int NKEEP = N < M ? N + 1 : M;
    RtPlan P; std::vector<int> fwd, inv;
    constexpr bool kCt = !std::is_same<CT, void>::value;
    if constexpr (kCt) {
        for (int i = 0; i < CT::Fwd::count; ++i) fwd.push_back(CT::Fwd::at(i));
        for (int i = 0; i < CT::Inv::count; ++i) inv.push_back(CT::Inv::at(i));
        if (!build_plan_from_radices(N, M, NKEEP, &P, &fwd, &inv, AI, AO)) { printf("bad ct plan\n"); return 1; }
        if (P.adv_in != CT::ADV_IN || P.adv_out != CT::ADV_OUT || P.carry_slots != CT::CARRY || P.emit_k != CT::EMIT_K || P.half_in != CT::HALF_IN ||
            P.twf_len != CT::twf_len() || P.twi_len != CT::twi_len()) { printf("ct plan constants differ from the runtime plan\n"); return 1; }
    } else if (!build_plan(N, M, NKEEP, &P, &fwd, &inv)) { printf("N=%d M=%d: no plan (generic kernel)\n", N, M); return 0; }
    if constexpr (kCt) ct_plan_pads(&P); else rt_plan_pads(&P);
    std::vector<uint32_t> ordf, ordi;
    build_stage_orders(&P, NS == 2 ? 8 : 16, !kCt, &ordf, &ordi);
    if (kCt && NS == 1) { ordf.clear(); ordi.clear(); }      // single-stream kernels of compile-time plans use the default mapping
    std::vector<uint16_t> pos_f(N), pos_i(M);
    build_pos_tables(fwd, inv, N, M, pos_f.data(), pos_i.data());
    srand(1234 + N);
    std::vector<double> h(taps);
    for (int n = 0; n < taps; ++n) h[n] = (rand() / (double)RAND_MAX - 0.5) * exp(-0.5 * pow((n - taps / 2) / (taps / 8.0), 2)) / N;
    std::vector<cd> H(N + 1);
    for (int k = 0; k <= N; ++k) { cd s = 0; for (int n = 0; n < taps; ++n) s += h[n] * std::polar(1.0, -pi * k * n / N); H[k] = s; }
    std::vector<float> fre(NKEEP), fim(NKEEP);
    for (int k = 0; k < NKEEP; ++k) { fre[k] = (float)H[k].real(); fim[k] = (float)H[k].imag(); }
    std::vector<float2> Pt(NKEEP), Qt(NKEEP), WI(M / 2 + 1), twf(P.twf_len), twi(P.twi_len);
    build_split_tables(N, M, NKEEP, fre.data(), fim.data(), Pt.data(), Qt.data(), WI.data());
    build_twiddles(P, twf.data(), twi.data());
    constexpr bool kCompact = [] { if constexpr (kCt) return CT::COMPACT_TW && NS == 2; else return false; }();
    std::vector<typename Tw<C, kCompact>::E> twf_e, twi_e;
    if constexpr (kCompact) { twf_e = twf; twi_e = twi; } else { twf_e = expand_table<C>(twf); twi_e = expand_table<C>(twi); }
    SplitLayout SL;
    build_split_layout(N, M, NKEEP, pos_f.data(), pos_i.data(), Pt.data(), Qt.data(), WI.data(), NS == 2 ? 8 : 16, &SL, P.pad_a, P.pad_b);
    if (SL.n_regular != split_regular_count(M, NKEEP) || SL.n_regular != P.n_regular) { printf("regular pair count differs\n"); return 1; }
    Tables<C, kCompact> T{{twf_e.data()}, {twi_e.data()}, SL.sidx.data(), SL.pq1.data(), SL.pq2.data(), SL.wi.data(),
                ordf.empty() ? nullptr : ordf.data(), ordi.empty() ? nullptr : ordi.data()};

    if constexpr (kCt && NS == 2) if ((int)ordf.size() != CT::ordf_len() || (int)ordi.size() != CT::ordi_len()) { printf("order table sizes differ\n"); return 1; }
    const int NB = 4;
    std::vector<float> x(NB * AI), x2(NB * AI);
    for (auto& v : x) v = (float)(rand() / (double)RAND_MAX - 0.5);
    for (auto& v : x2) v = (float)(rand() / (double)RAND_MAX - 0.5);
    const int valid_last = AI - 77;                // last block partially valid
    std::vector<MT> A(phys_len(N, P.pad_a)), B(phys_len(M, P.pad_b)), carry(P.carry_slots);
    memset(carry.data(), 0, sizeof(MT) * carry.size());
    std::vector<float> out(NB * AO, 0.f), out2(NB * AO, 0.f);
    for (int b = 0; b < NB; ++b) {
        const int valid = b + 1 == NB ? valid_last : AI;
        auto smp = [&](const std::vector<float>& v, int i) { return i < valid ? v[b * AI + i] : 0.f; };
        auto loader = [&](int n) {
            if constexpr (NS == 1) return make_float2(smp(x, 2 * n), smp(x, 2 * n + 1));
            else { cx2 c; c.re = make_float2(smp(x, 2 * n), smp(x2, 2 * n)); c.im = make_float2(smp(x, 2 * n + 1), smp(x2, 2 * n + 1)); return c; }
        };
        auto sink = [&](int n, C y) {
            if constexpr (NS == 1) { out[b * AO + 2 * n] = y.x; out[b * AO + 2 * n + 1] = y.y; }
            else { out[b * AO + 2 * n] = y.re.x; out[b * AO + 2 * n + 1] = y.im.x; out2[b * AO + 2 * n] = y.re.y; out2[b * AO + 2 * n + 1] = y.im.y; }
        };
        if constexpr (kCt) process_block_ct<CT, C>(HostExec{nl}, T, A.data(), B.data(), carry.data(), loader, sink, [] {});
        else if (P.pad_a) process_block<C, MapPad8>(HostExec{nl}, P, T, A.data(), B.data(), carry.data(), loader, sink, [] {});
        else process_block<C>(HostExec{nl}, P, T, A.data(), B.data(), carry.data(), loader, sink, [] {});
    }
    double worst = 0;
    for (int stream = 0; stream < NS; ++stream) {
    const std::vector<float>& xs = stream ? x2 : x;
    const std::vector<float>& outs = stream ? out2 : out;
    std::vector<double> ref(NB * AO + 2 * M, 0.0);
    std::vector<cd> wN(2 * N), wM(2 * M);              // exact index reduction keeps the naive DFTs O(n^2) multiplies only
    for (int j = 0; j < 2 * N; ++j) wN[j] = std::polar(1.0, -pi * j / N);
    for (int j = 0; j < 2 * M; ++j) wM[j] = std::polar(1.0, pi * j / M);
    for (int b = 0; b < NB; ++b) {
        const int valid = b + 1 == NB ? valid_last : AI;
        std::vector<cd> X(N + 1), Y(M + 1, cd(0, 0));
        for (int k = 0; k <= N; ++k) { cd s = 0; for (int n = 0; n < valid; ++n) s += (double)xs[b * AI + n] * wN[(size_t)k * n % (2 * N)]; X[k] = s; }
        for (int k = 0; k < NKEEP; ++k) Y[k] = X[k] * cd(fre[k], fim[k]);
        for (int n = 0; n < 2 * M; ++n) {
            double s = Y[0].real() + Y[M].real() * ((n & 1) ? -1.0 : 1.0);
            for (int k = 1; k < M; ++k) s += 2.0 * (Y[k] * wM[(size_t)k * n % (2 * M)]).real();
            ref[b * AO + n] += s;
        }
    }
    double maxerr = 0, rms = 0;
    for (int i = 0; i < NB * AO; ++i) { maxerr = fmax(maxerr, fabs(outs[i] - ref[i])); rms += ref[i] * ref[i]; }
    rms = sqrt(rms / (NB * AO));
    worst = fmax(worst, maxerr / rms);
    }
    printf("%s x%d N=%4d M=%4d adv %4d/%4d fwd[", kCt ? "ct" : "rt", NS, N, M, AI, AO);
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
#define BB_CT(NAME, NI, NO, TH, ...)                                                                                   \
    { using PLT = __VA_ARGS__;                                                                                          \
      for (int nl : {64, 96, 160, 320}) {                                                                               \
          if (nl == 64) bad += check_impl<PLT, float2>(PLT::N, PLT::M, nl, PLT::ADV_IN, PLT::ADV_OUT, NI);             \
          else bad += check_impl<PLT, cx2>(PLT::N, PLT::M, nl, PLT::ADV_IN, PLT::ADV_OUT, NI);                         \
      } }
    BB_K2_CT_PLANS(BB_CT)
#undef BB_CT
    printf(bad ? "FAILED\n" : "all plans ok\n");
    return bad;
}
