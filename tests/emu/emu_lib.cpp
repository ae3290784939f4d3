// emu_lib.cpp — TEST INFRASTRUCTURE: runs the product's warp-level device code
// (circuitsimulator_b200/csrc/bg_device.cuh, bg_philox.cuh) on the CPU warp emulator.
#ifndef BG_EMU_FAST
#define BG_COUNT_WORK 1
#endif
#include "cpu_warp.h"
#include "bg_device.cuh"
#include "bg_warp_ops.cuh"
#include "bg_tpp.cuh"
#include "bg_shb.cuh"
#include "bg_prep.cuh"
#include <string.h>
#include <algorithm>
#include <chrono>
#include <vector>

namespace emu {
thread_local Warp* g_warp = nullptr;
static void trampoline() {
    Warp* w = g_warp;
    w->body();
    w->done[w->cur] = true;
    swapcontext(&w->fib[w->cur], &w->main_ctx);
}
void run(const std::function<void()>& body) {
    static thread_local Warp* w = nullptr;
    const size_t STK = 1 << 18;
    if (!w) { w = new Warp(); for (int i = 0; i < 32; i++) w->stacks[i] = (char*)malloc(STK); }
    g_warp = w;
    w->body = body;
    for (int i = 0; i < 32; i++) {
        w->done[i] = false; w->phase[i] = 0;
        getcontext(&w->fib[i]);
        w->fib[i].uc_stack.ss_sp = w->stacks[i];
        w->fib[i].uc_stack.ss_size = STK;
        w->fib[i].uc_link = &w->main_ctx;
        makecontext(&w->fib[i], (void (*)())trampoline, 0);
    }
    for (int j = 0; j < 32; j++) { w->kind[0][j] = w->kind[1][j] = K_NONE; }
    while (true) {
        int ndone = 0;
        for (int i = 0; i < 32; i++) {
            if (w->done[i]) { ndone++; continue; }
            w->cur = i;
            swapcontext(&w->main_ctx, &w->fib[i]);
            if (w->done[i]) ndone++;
        }
        if (ndone == 32) break;
        if (ndone != 0) { fprintf(stderr, "emu: lanes finished at different collectives\n"); abort(); }
    }
}
}  // namespace emu

#if defined(BG_EMU_FAST)
struct BgWork { unsigned long long xors, rows, dimers, monomers, basis_changes, pairs; };
#endif
BgWork g_bg_work = {0, 0, 0, 0, 0, 0};
int* g_bg_trace = nullptr; int g_bg_trace_n = 0, g_bg_trace_cap = 0;
using namespace bg;

#ifndef EMU_SHB_G
#define EMU_SHB_G 2
#endif
static double g_emu_chi_seconds = 0;     // time spent in the chi loop of emu_terms_tpp (bench.py: cpu_baseline_packed)
static int g_emu_lam_max = 4;   // checks carried as Lagrange variables (as k_pairs_tpp does; 0: every check pivoted per term)
// Same as emu_terms, but the chi loop runs through the thread-per-pair code (bg_tpp.cuh): the
// ambient form is produced by the warp-level code under emulation, each term is then a plain call.
template <int NS>
static int terms_tpp(const bg_state* theta, const bg_projector* P, int project, int exact, int t, int nterms,
                     const uint64_t* terms, int32_t* epm, int* npf_out, int* k_out, long long* zw_out) {
    typedef typename WordOf<NS>::T W;
    int alive = 1, npf = 0, k1 = 0;
    W Jrows[64], Cwrows[64]; W D1 = 0, D2 = 0, Cpend = 0, Cbeta = 0; uint32_t Q = 0;
    emu::run([&]() {
        Native<NS> st; Ambient<NS> am;
        native_load<NS>(st, theta);
        int n = 0; bool ok = true;
        make_ambient<NS>(st, am);                         // as k_prepare does: convert once, project in ambient form
        if (project) ok = project_ambient<NS>(am, P, n);
        const int lane = bg_lane();
        if (lane == 0) { alive = ok; npf = n; }
        if (ok) {
            for (int s = 0; s < NS; s++) { Jrows[lane + 32 * s] = am.f.J[s]; Cwrows[lane + 32 * s] = am.Cw[s]; }
            if (lane == 0) { D1 = am.f.D1; D2 = am.f.D2; Q = am.f.Q; Cpend = am.Cpend; Cbeta = am.Cbeta; k1 = am.k1; }
        }
    });
    *npf_out = npf; *k_out = k1;
    Zw z; z.a[0] = z.a[1] = z.a[2] = z.a[3] = 0;
    if (alive) {
        TShared<W> sh;
        W cwv[64];
        W amb[2 * 64];
        sh.J = amb; sh.D1 = D1; sh.D2 = D2; sh.Q = Q; sh.k1 = k1; sh.t = t; sh.ncons = 0; sh.nlam = 0; sh.cbeta = 0;
        sh.cwv = cwv; sh.cbetav = 0;
        for (int j = 0; j < TPP_MAXC; j++) sh.cw[j] = 0;
        for (W r = Cpend; r; r &= r - 1) {
            int b = tlowest(r);
            cwv[sh.ncons] = Cwrows[b];
            sh.cbetav |= (W)((Cbeta >> b) & 1) << sh.ncons;
            if (sh.ncons < TPP_MAXC) { sh.cw[sh.ncons] = Cwrows[b]; sh.cbeta |= (uint32_t)((Cbeta >> b) & 1) << sh.ncons; }
            sh.ncons++;
        }
        const bool many = sh.ncons > TPP_MAXC;      // the device routes these to the MANYC instantiation
        for (int q = 0; q < t; q++) amb[q] = Jrows[q];
        const int lam_max = std::max(0, std::min(g_emu_lam_max, (int)(8 * sizeof(W)) - t));
        if (!many && sh.ncons <= lam_max) {         // as k_pairs_tpp: the checks become Lagrange variables t .. t+nlam-1
            for (int q = 0; q < t; q++)
                for (int j = 0; j < sh.ncons; j++) amb[q] |= (W)((sh.cw[j] >> q) & 1) << (t + j);
            for (int j = 0; j < sh.ncons; j++) amb[t + j] = sh.cw[j];
            sh.D2 |= (W)sh.cbeta << t;
            sh.nlam = sh.ncons; sh.ncons = 0;
        }
        W work[128];
        Rows<W> rows; rows.base = work; rows.stride = 1; rows.sbase = 0; rows.sstride = 0;
        const auto chi_t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < nterms; i++) {
            int e, p, m;
            if (many) {
                if (exact) t_term_H<W, true>(rows, sh, (W)terms[i], e, p, m); else t_term_L<W, true>(rows, sh, (W)terms[i], e, p, m);
            } else {
                if (exact) t_term_H<W, false>(rows, sh, (W)terms[i], e, p, m); else t_term_L<W, false>(rows, sh, (W)terms[i], e, p, m);
            }
            g_bg_work.pairs++;
            BG_TRACE(-1, i);                      // end-of-pair marker
            zw_add(z, e, p, m, t / 2 + 1);
            if (epm) { epm[3 * i] = e; epm[3 * i + 1] = p; epm[3 * i + 2] = m; }
        }
        g_emu_chi_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - chi_t0).count();
    }
    if (zw_out) for (int j = 0; j < 4; j++) zw_out[j] = z.a[j];
    return alive;
}

// Brute-force self-test of the thread-level exponential sums (bg_tpp.cuh): random forms on n <= 14 variables
// scattered over the word, sum_{x} e^{i pi q(x)/4} by enumeration against t_expsum_odd (free slots: none / a few /
// all, which exercises the borrowed-variable steps and their hand-over to the rounds) and t_expsum (fold + dimers).
// Returns the number of mismatches.
template <typename W> static int expsum_selftest(uint64_t seed, int trials) {
    const int bits = 8 * (int)sizeof(W);
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
    int bad = 0;
    for (int trial = 0; trial < trials; trial++) {
        const int n = 1 + (int)(rnd() % 13);
        W A = 0;
        while (tpopc(A) < n) A |= (W)1 << (rnd() % bits);
        const unsigned podd = (unsigned)(rnd() % 5), pj = 1 + (unsigned)(rnd() % 4);
        W J[64], D1 = 0, D2 = (W)rnd() & A;
        for (int a = 0; a < bits; a++) { J[a] = 0; if (((A >> a) & 1) && rnd() % 4 < podd) D1 |= (W)1 << a; }
        for (int a = 0; a < bits; a++) for (int b = a + 1; b < bits; b++)
            if (((A >> a) & 1) && ((A >> b) & 1) && rnd() % 5 < pj) { J[a] |= (W)1 << b; J[b] |= (W)1 << a; }
        for (int a = 0; a < bits; a++) if ((D1 >> a) & 1) J[a] |= (W)1 << a;
        const uint32_t Q = (uint32_t)(rnd() % 8);
        int vars[64], nv = 0;
        for (int a = 0; a < bits; a++) if ((A >> a) & 1) vars[nv++] = a;
        long long z[4] = {0, 0, 0, 0};                       // coefficients of 1, w, w^2, w^3
        for (uint32_t mk = 0; mk < (1u << nv); mk++) {
            W x = 0;
            for (int i = 0; i < nv; i++) if ((mk >> i) & 1) x |= (W)1 << vars[i];
            int q = (int)Q;
            for (int i = 0; i < nv; i++) {
                const int a = vars[i];
                if (!((x >> a) & 1)) continue;
                q += 2 * (int)((D1 >> a) & 1) + 4 * (int)((D2 >> a) & 1);
                q += 4 * tpopc((W)(J[a] & x & ~((((W)2) << a) - 1)));
            }
            q &= 7;
            if (q < 4) z[q]++; else z[q - 4]--;
        }
        for (int variant = 0; variant < 4; variant++) {
            W work[64];
            for (int a = 0; a < bits; a++) work[a] = J[a] & A;
            Rows<W> rows; rows.base = work; rows.stride = 1; rows.sbase = 0; rows.sstride = 0;
            TF<W> f; f.A = A; f.D1 = D1; f.D2 = D2; f.Q = Q;
            int e, p, m;
            if (variant == 3) t_expsum<W>(rows, f, e, p, m);
            else {
                uint32_t fr = ~(uint32_t)A;
                if (variant == 0) fr = 0;
                if (variant == 1) { uint32_t k = 0; for (int i = 0; i < 2 && (fr & ~k); i++) k |= 1u << __builtin_ctz(fr & ~k); fr = k; }
                t_expsum_odd(rows, f, e, p, m, fr);
            }
            // eps 2^(p/2) w^m as integer coefficients: p even -> 2^(p/2) w^m; p odd -> 2^((p-1)/2) (w^(m+1) + w^(m-1))
            long long g[4] = {0, 0, 0, 0};
            if (e) {
                auto addw = [&](int mm, long long c) { mm &= 7; if (mm < 4) g[mm] += c; else g[mm - 4] -= c; };
                if (p % 2 == 0) addw(m, 1ll << (p / 2));
                else { addw(m + 1, 1ll << ((p - 1) / 2)); addw(m - 1, 1ll << ((p - 1) / 2)); }
            }
            if (g[0] != z[0] || g[1] != z[1] || g[2] != z[2] || g[3] != z[3]) bad++;
        }
    }
    return bad;
}

// One sample of the device RNG, projected by P, through BOTH formulations of k_prepare: the warp-per-sample code
// (native_random + make_ambient + project_ambient under the warp emulator) and the thread-per-sample code
// (bg_prep.cuh).  rec = [alive, k1, npf, Q, D1, D2, Cpend, Cbeta, J[64], Cw[64]] as uint64 (the fields of SampleRec).
template <int NS>
static void prepare_both(int n, uint64_t seed, uint32_t bin, uint64_t sample, const double* cdf, const bg_projector* P,
                         uint64_t* rec_warp, uint64_t* rec_thread) {
    typedef typename WordOf<NS>::T W;
    memset(rec_warp, 0, 136 * sizeof(uint64_t));
    memset(rec_thread, 0, 136 * sizeof(uint64_t));
    emu::run([&]() {
        Native<NS> st; Ambient<NS> am;
        native_random<NS>(st, n, seed, bin, sample, cdf);
        make_ambient<NS>(st, am);
        int npf = 0;
        const bool ok = project_ambient<NS>(am, P, npf);
        const int lane = bg_lane();
        if (lane == 0) rec_warp[0] = ok;
        if (!ok) return;
        if (lane == 0) {
            rec_warp[1] = (uint64_t)am.k1; rec_warp[2] = (uint64_t)npf; rec_warp[3] = am.f.Q;
            rec_warp[4] = am.f.D1; rec_warp[5] = am.f.D2; rec_warp[6] = am.Cpend; rec_warp[7] = am.Cbeta;
        }
        for (int s = 0; s < NS; s++) { rec_warp[8 + lane + 32 * s] = am.f.J[s]; rec_warp[72 + lane + 32 * s] = am.Cw[s]; }
    });
    W jr[64], cr[64];
    for (int i = 0; i < 64; i++) { jr[i] = (W)0xdeadbeefdeadbeefull; cr[i] = (W)0xfeedfacefeedfaceull; }   // stale shared memory
    Rows<W> J, C;
    J.base = jr; J.stride = 1; J.sbase = 0; J.sstride = 0;
    C.base = cr; C.stride = 1; C.sbase = 0; C.sstride = 0;
    TSample<W> s;
    t_random_ambient<W>(J, C, n, seed, bin, sample, cdf, s);
    t_project<W>(J, C, n, s, P);
    rec_thread[0] = s.alive;
    if (!s.alive) return;
    rec_thread[1] = (uint64_t)(n - tpopc(s.Cpend)); rec_thread[2] = (uint64_t)s.npf; rec_thread[3] = s.Q;
    rec_thread[4] = s.D1; rec_thread[5] = s.D2; rec_thread[6] = s.Cpend; rec_thread[7] = s.Cbeta;
    for (int v = 0; v < n; v++) {
        rec_thread[8 + v] = jr[v];
        rec_thread[72 + v] = ((s.Cpend >> v) & 1) ? cr[v] : 0;
    }
}
extern "C" {

// <b|a> through the generic path
void emu_inner_product(const bg_state* a, const bg_state* b, int32_t* epm) {
    const int n = a->n;
    emu::run([&]() {
        int e, p, m;
        if (n <= 32) warp_inner_product<1>(a, b, e, p, m); else warp_inner_product<2>(a, b, e, p, m);
        if (bg_lane() == 0) { epm[0] = e; epm[1] = p; epm[2] = m; }
    });
}

// theta (optionally projected by P) against every term of the decomposition
// epm: chi x 3; returns alive flag; npf = number of 2^-1/2 factors; kout = dim after projection
int emu_terms(const bg_state* theta, const bg_projector* P, int project, int exact, int t, int nterms,
              const uint64_t* terms, int32_t* epm, int* npf_out, int* k_out, long long* zw_out) {
    int alive = 1;
    emu::run([&]() {
        if (t <= 32) warp_sample_terms<1>(theta, P, project, exact, t, nterms, terms, epm, &alive, npf_out, k_out, zw_out);
        else warp_sample_terms<2>(theta, P, project, exact, t, nterms, terms, epm, &alive, npf_out, k_out, zw_out);
    });
    return alive;
}

int emu_terms_tpp(const bg_state* theta, const bg_projector* P, int project, int exact, int t, int nterms,
                  const uint64_t* terms, int32_t* epm, int* npf_out, int* k_out, long long* zw_out) {
    if (t <= 32) return terms_tpp<1>(theta, P, project, exact, t, nterms, terms, epm, npf_out, k_out, zw_out);
    return terms_tpp<2>(theta, P, project, exact, t, nterms, terms, epm, npf_out, k_out, zw_out);
}

// The chi loop through the shared high-block reduction (bg_shb.cuh), as k_pairs_shb runs it: plan on the host,
// relabelling + shb_reduce under the warp emulator, then one lane per term.  Returns -1 when the decomposition
// has no plan, -2 when theta has more checks than the kernel takes (the device routes those to the generic kernel).
int emu_terms_shb(const bg_state* theta, const bg_projector* P, int project, int t, int k, const uint64_t* Lrows,
                  int32_t* epm, int* npf_out, int* k_out, long long* zw_out, int* nleft_hist) {
    std::vector<uint64_t> L(Lrows, Lrows + k), terms((size_t)1 << k);
    for (size_t i = 0; i < terms.size(); i++) {
        uint64_t x = 0;
        for (int j = 0; j < k; j++) if ((i >> (k - 1 - j)) & 1) x ^= L[j];
        terms[i] = x & (t >= 64 ? ~0ull : ((1ull << t) - 1));
    }
    ShbPlan pl = shb_make_plan(t, k, L, terms);
    if (!pl.ok) return -1;
    int alive = 1, npf = 0, k1 = 0, toomany = 0;
    Zw ztot; ztot.a[0] = ztot.a[1] = ztot.a[2] = ztot.a[3] = 0;
    static uint32_t red[4 * 48];
    static uint32_t work[32][40];
    static uint64_t Jrows[64], Cwrows[64];
    emu::run([&]() {
        Native<2> st; Ambient<2> am;
        native_load<2>(st, theta);
        int n = 0; bool ok = true;
        make_ambient<2>(st, am);
        if (project) ok = project_ambient<2>(am, P, n);
        const int lane = bg_lane();
        if (lane == 0) { alive = ok; npf = n; k1 = am.k1; }
        if (!ok) return;
        if (popcw(am.Cpend) > SHB_MAXLAM) { if (lane == 0) toomany = 1; return; }
        for (int s = 0; s < 2; s++) { Jrows[lane + 32 * s] = am.f.J[s]; Cwrows[lane + 32 * s] = am.Cw[s]; }
        __syncwarp();
        ShbPerm pm; pm.nh = pl.nh; pm.nsw = pl.nsw;
        for (int i = 0; i < SHB_MAXH; i++) { pm.swp[i] = pl.swp[i]; pm.swq[i] = pl.swq[i]; }
        for (int i = 0; i < 64; i++) pm.iperm[i] = pl.iperm[i];
        ShbForm f;
        const int nlam = shb_load(Jrows, Cwrows, am.Cpend, am.Cbeta, am.f.D1, am.f.D2, am.f.Q, t, pm, f);
        const uint32_t lam_bits = ((1u << nlam) - 1u) << pm.nh;
        Zw z; z.a[0] = z.a[1] = z.a[2] = z.a[3] = 0;
        const int LPG = 32 / EMU_SHB_G, hh = lane / LPG;      // as k_pairs_shb: EMU_SHB_G classes at a time, 32 / G lanes each
        for (size_t g = 0; g < terms.size(); g += 32 * EMU_SHB_G) {
            ShbBatch sb;
            sb.red = red + hh * 48; sb.left = red + hh * 48 + 32; sb.k1 = am.k1; sb.nlam = nlam;
            sb.D1 = sb.D2 = sb.Q = 0; sb.p = 0; sb.nleft = 0; sb.left_d2 = 0;
            int nlmax = 0;
            __syncwarp();
            for (int c = 0; c < EMU_SHB_G; c++) {
                const uint32_t pattern = (uint32_t)(pl.terms[g + 32 * c] >> 32);
                ShbOut o; uint32_t Lr, Rr;
                shb_reduce(f, pattern | lam_bits, Lr, Rr, o);
                red[c * 48 + lane] = Lr;
                int nleft = 0; uint32_t ld2 = 0;
                for (uint32_t rem = o.left; rem;) {
                    const int u = shb_top(rem);
                    rem ^= 1u << u;
                    const uint32_t w = __shfl_sync(BG_FULL, Rr, u);
                    if (lane == 0) red[c * 48 + 32 + nleft] = w;
                    ld2 |= ((o.left_d2 >> u) & 1u) << nleft;
                    nleft++;
                }
                if (lane == 0 && nleft_hist) nleft_hist[nleft < 15 ? nleft : 15]++;
                nlmax = std::max(nlmax, nleft);
                if (hh == c) { sb.D1 = o.D1; sb.D2 = o.D2; sb.Q = o.Q; sb.p = (int)o.p; sb.nleft = nleft; sb.left_d2 = ld2; }
            }
            __syncwarp();
            for (int rd = 0; rd < EMU_SHB_G; rd++) {
                const size_t ti = g + 32 * hh + LPG * rd + (lane & (LPG - 1));
                const uint64_t term = pl.terms[ti];
                Rows<uint32_t> rows; rows.base = work[lane]; rows.stride = 1; rows.sbase = 0; rows.sstride = 0;
                int e, p, m;
                t_term_shb(rows, sb, term, nlmax, e, p, m);
                __syncwarp();
                zw_add(z, e, p, m, t / 2 + 1);
                if (epm) { const int nat = pl.nat[ti]; epm[3 * nat] = e; epm[3 * nat + 1] = p; epm[3 * nat + 2] = m; }
            }
        }
        for (int j = 0; j < 4; j++) {                      // lanes take turns (the emulator runs them round-robin)
            for (int l = 0; l < 32; l++) { if (lane == l) ztot.a[j] += z.a[j]; __syncwarp(); }
        }
    });
    *npf_out = npf; *k_out = k1;
    if (zw_out) for (int j = 0; j < 4; j++) zw_out[j] = ztot.a[j];
    if (toomany) return -2;
    return alive;
}

// The shared high-block plan of bg_shb_plan.h for one L: returns 0 (no plan), 1 (plan, every invariant holds) or a
// negative code naming the invariant that fails.  digest receives a hash of the plan (the search is deterministic).
int emu_shb_plan_check(int t, int k, const uint64_t* Lrows, unsigned long long* digest) {
    std::vector<uint64_t> L(Lrows, Lrows + k), terms((size_t)1 << k);
    const uint64_t maskt = t >= 64 ? ~0ull : ((1ull << t) - 1);
    for (size_t i = 0; i < terms.size(); i++) {
        uint64_t x = 0;
        for (int j = 0; j < k; j++) if ((i >> (k - 1 - j)) & 1) x ^= L[j];
        terms[i] = x & maskt;
    }
    ShbPlan pl = shb_make_plan(t, k, L, terms);
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { h = (h ^ v) * 1099511628211ull; };
    mix((unsigned long long)pl.ok);
    if (!pl.ok) { if (digest) *digest = h; return 0; }
    const size_t chi = terms.size();
    if (pl.nh != t - 32 || pl.terms.size() != chi || pl.nat.size() != chi) return -1;
    std::vector<char> seen(chi, 0);
    for (size_t i = 0; i < chi; i++) {
        const int32_t n = pl.nat[i];
        if (n < 0 || (size_t)n >= chi || seen[n]) return -2;                      // nat is a permutation
        seen[n] = 1;
        if (pl.terms[i] != shb_permute_bits(terms[n], pl)) return -3;              // terms are the relabelled natural terms
        if (pl.terms[i] >> t) return -4;
        if (32 - __builtin_popcount((uint32_t)pl.terms[i]) < SHB_RELOC) return -5; // free low slots for the leftovers
        mix(pl.terms[i]); mix((unsigned long long)n);
    }
    for (size_t i = 0; i < chi; i += 32) {
        for (size_t j = 1; j < 32; j++)
            if ((pl.terms[i + j] >> 32) != (pl.terms[i] >> 32)) return -6;          // one high pattern per class of 32
    }
    for (int i = 0; i < pl.nsw; i++) {
        if (!(pl.swp[i] < 32 && pl.swq[i] >= 32 && pl.swq[i] < t)) return -8;       // every swap crosses the word boundary
        for (int j = 0; j < i; j++) if (pl.swp[i] == pl.swp[j] || pl.swq[i] == pl.swq[j]) return -9;
    }
    // the high columns of L (after relabelling) span at most k - 5 dimensions: classes of at least 32 terms
    {
        ShbBasis B; int rk = 0;
        for (int c = 32; c < t; c++) {
            uint32_t col = 0;
            for (int j = 0; j < k; j++) col |= (uint32_t)((shb_permute_bits(L[j] & maskt, pl) >> c) & 1ull) << j;
            rk += B.add(col) ? 1 : 0;
        }
        if (rk > k - 5) return -10;
    }
    if (digest) *digest = h;
    return 1;
}

// zw32_add (the 32-bit per-thread accumulation of k_pairs_shb) against zw_add for every (eps, p, m) a pair of
// normalised states can give at 33 <= t <= 44, alone and summed over 128 worst-case terms.  Returns the mismatches.
int emu_zw32_selftest(void) {
    int bad = 0;
    for (int t = 33; t <= ZW32_MAX_T; t++) {
        const int sh = t / 2 + 1;
        for (int eps = 0; eps <= 1; eps++)
            for (int p = -t; p <= 0; p++)
                for (int m = 0; m < 16; m++) {
                    Zw a; Zw32 b;
                    for (int j = 0; j < 4; j++) { a.a[j] = 0; b.a[j] = 0; }
                    for (int rep = 0; rep < ZW32_MAX_TERMS; rep++) { zw_add(a, eps, p, m, sh); zw32_add(b, eps, p, m, sh); }
                    for (int j = 0; j < 4; j++) if (a.a[j] != (long long)b.a[j]) bad++;
                }
    }
    return bad;
}

int emu_expsum_selftest(int wordbits, unsigned long long seed, int trials) {
    return wordbits == 32 ? expsum_selftest<uint32_t>(seed, trials) : expsum_selftest<uint64_t>(seed, trials);
}

void emu_set_lam_max(int n) { g_emu_lam_max = n; }
double emu_chi_seconds(int reset) { const double v = g_emu_chi_seconds; if (reset) g_emu_chi_seconds = 0; return v; }
void emu_trace(int* buf, int cap) { g_bg_trace = buf; g_bg_trace_cap = cap; g_bg_trace_n = 0; }
int emu_trace_len(void) { return g_bg_trace_n; }

void emu_work_counters(unsigned long long* out, int reset) {
    out[0] = g_bg_work.xors; out[1] = g_bg_work.rows; out[2] = g_bg_work.dimers; out[3] = g_bg_work.monomers;
    out[4] = g_bg_work.basis_changes; out[5] = g_bg_work.pairs;
    if (reset) g_bg_work = BgWork{0, 0, 0, 0, 0, 0};
}

int emu_measure_pauli(bg_state* st, uint64_t* A, int m, uint64_t zeta, uint64_t xi) {
    int res = 0;
    const int n = st->n;
    emu::run([&]() {
        int r;
        if (n <= 32) r = warp_measure_pauli<1>(st, A, m, zeta, xi); else r = warp_measure_pauli<2>(st, A, m, zeta, xi);
        if (bg_lane() == 0) res = r;
    });
    return res;
}

void emu_random_state(int n, uint64_t seed, uint32_t bin, uint64_t sample, const double* cdf, bg_state* out, uint64_t* A) {
    emu::run([&]() {
        if (n <= 32) warp_random_state<1>(n, seed, bin, sample, cdf, out, A);
        else warp_random_state<2>(n, seed, bin, sample, cdf, out, A);
    });
}

void emu_prepare_both(int n, uint64_t seed, uint32_t bin, uint64_t sample, const double* cdf, const bg_projector* P,
                      uint64_t* rec_warp, uint64_t* rec_thread) {
    if (n <= 32) prepare_both<1>(n, seed, bin, sample, cdf, P, rec_warp, rec_thread);
    else prepare_both<2>(n, seed, bin, sample, cdf, P, rec_warp, rec_thread);
}

}  // extern "C"
