// emu_lib.cpp — TEST INFRASTRUCTURE: runs the product's warp-level device code
// (circuitsimulator_b200/csrc/bg_device.cuh, bg_philox.cuh) on the CPU warp emulator.
#define BG_COUNT_WORK 1
#include "cpu_warp.h"
#include "bg_device.cuh"
#include "bg_warp_ops.cuh"
#include "bg_tpp.cuh"
#include <string.h>
#include <algorithm>

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

BgWork g_bg_work = {0, 0, 0, 0, 0, 0};
int* g_bg_trace = nullptr; int g_bg_trace_n = 0, g_bg_trace_cap = 0;
using namespace bg;

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
    }
    if (zw_out) for (int j = 0; j < 4; j++) zw_out[j] = z.a[j];
    return alive;
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

void emu_set_lam_max(int n) { g_emu_lam_max = n; }
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

}  // extern "C"
