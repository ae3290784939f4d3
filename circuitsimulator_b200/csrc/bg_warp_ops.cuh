// bg_warp_ops.cuh — whole operations for one warp, composed from bg_device.cuh.  Shared by the
// CUDA kernels (bgnorm.cu) and, for the CPU-side check of the same source, tests/emu.
#pragma once
#include "bg_device.cuh"
#include "bg_philox.cuh"

namespace bg {

// Everything the pair loop needs about one (projected) theta.
template <int NS> struct Ambient {
    typedef typename WordOf<NS>::T W;
    QForm<NS> f;
    W Cw[NS];
    W Cpend, Cbeta;
    int k1;
};

template <int NS> BG_DEV void make_ambient(const Native<NS>& st, Ambient<NS>& am) {
    ambient<NS>(st, am.f, am.Cw, am.Cpend, am.Cbeta);
    am.k1 = popcw(st.f.A);
}

// Apply the projector's generators in order (innerprod.c:100-116).  Returns false when
// annihilated; npf counts the 2^-1/2 factors.
template <int NS> BG_DEV bool project_native(Native<NS>& st, const bg_projector* P, int& npf) {
    typedef typename WordOf<NS>::T W;
    npf = 0;
    const int ns = P->nstabs;
    for (int i = 0; i < ns; i++) {
        const int r = native_measure<NS>(st, (uint32_t)P->phase[i], (W)P->zs[i], (W)P->xs[i]);
        if (r == 0) return false;
        if (r == 2) npf++;
    }
    return true;
}

// The same on a state in ambient form (bg_device.cuh: ambient_measure) — what the hot path uses: theta is
// converted once, right after it is drawn, and every generator is a handful of mask operations.
template <int NS> BG_DEV bool project_ambient(Ambient<NS>& am, const bg_projector* P, int& npf) {
    typedef typename WordOf<NS>::T W;
    npf = 0;
    checks_echelon<NS>(am.Cw, am.Cpend, am.Cbeta);
    const int ns = P->nstabs;
    for (int i = 0; i < ns; i++) {
        const int r = ambient_measure<NS>(am.f, am.Cw, am.Cpend, am.Cbeta, (uint32_t)P->phase[i], (W)P->zs[i], (W)P->xs[i]);
        if (r == 0) return false;
        if (r == 2) npf++;
    }
    am.k1 = popcw(am.f.A) - popcw(am.Cpend);
    return true;
}

// innerProductExact(state1 = a, state2 = b) (stabilizer.c:589-659) for arbitrary states:
// both become ambient forms, q = q_a - q_b on K_a ∩ K_b.
template <int NS> BG_DEV void warp_inner_product(const bg_state* a, const bg_state* b, int& eps, int& p, int& m) {
    typedef typename WordOf<NS>::T W;
    Ambient<NS> fa, fb;
    {
        Native<NS> st;
        native_load<NS>(st, a);
        make_ambient<NS>(st, fa);
        native_load<NS>(st, b);
        make_ambient<NS>(st, fb);
    }
    QForm<NS> f;
    f.A = fa.f.A;
    f.Q = (fa.f.Q + 8u - fb.f.Q) & 7u;
    f.D1 = fa.f.D1 ^ fb.f.D1;
    f.D2 = fa.f.D2 ^ fb.f.D2 ^ (~fa.f.D1 & fb.f.D1);
#pragma unroll
    for (int s = 0; s < NS; s++) f.J[s] = fa.f.J[s] ^ fb.f.J[s];
    // membership of K_a, then of K_b; the second bank is reduced while the first is pivoted
    W pend = fa.Cpend;
    W Cbeta = fa.Cbeta, Cbeta2 = fb.Cbeta;
    bool ok = true;
    while (pend && ok) {
        const int j = lowestw(pend);
        pend ^= bitw<W>(j);
        const W w = rowb<NS>(fa.Cw, j) & f.A;
        const uint32_t beta = getw(Cbeta, j);
        if (w == 0) { if (beta) ok = false; continue; }
        const W bi = bitw<W>(lowestw(w));
        bool hit[NS], hit2[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            hit[s] = (fa.Cw[s] & bi) != 0; if (hit[s]) fa.Cw[s] ^= w;
            hit2[s] = (fb.Cw[s] & bi) != 0; if (hit2[s]) fb.Cw[s] ^= w;
        }
        const W hb = ballotw<NS>(hit), hb2 = ballotw<NS>(hit2);
        if (beta) { Cbeta ^= hb; Cbeta2 ^= hb2; }
        pivot<NS>(f, w, beta);
    }
    if (ok) ok = apply_constraints<NS>(f, fb.Cw, fb.Cpend, Cbeta2);
    if (!ok) { eps = 0; p = 0; m = 0; return; }
    expsum<NS>(f, eps, p, m);
    if (eps) p -= fa.k1 + fb.k1; else { p = 0; m = 0; }
}

// One theta against nterms decomposition terms (the body of singleProjectorSample,
// innerprod.c:88-144).  terms[i] = x~_i (|L>, exact == 0) or the pair mask e1_i (|H^t>, exact != 0).
// Outputs (lane 0 writes): epm (optional), alive, npf, k after projection, zw[4] (optional).
template <int NS>
BG_DEV void warp_sample_terms(const bg_state* theta, const bg_projector* P, int project, int exact, int t,
                              int nterms, const uint64_t* terms, int32_t* epm, int* alive_out, int* npf_out,
                              int* k_out, long long* zw_out) {
    typedef typename WordOf<NS>::T W;
    Ambient<NS> am;
    int npf = 0;
    bool alive = true;
    {
        Native<NS> st;
        native_load<NS>(st, theta);
        make_ambient<NS>(st, am);
        if (project) alive = project_ambient<NS>(am, P, npf);
        if (!alive) am.k1 = 0;
    }
    Zw z; z.a[0] = z.a[1] = z.a[2] = z.a[3] = 0;
    const int sh = t / 2 + 1;
    if (alive) {
        for (int i = 0; i < nterms; i++) {
            int e, p, m;
            if (exact) term_H<NS>(am.f, am.Cw, am.Cpend, am.Cbeta, am.k1, t, (W)terms[i], e, p, m);
            else term_L<NS>(am.f, am.Cw, am.Cpend, am.Cbeta, am.k1, (W)terms[i], e, p, m);
            zw_add(z, e, p, m, sh);
            if (epm && bg_lane() == 0) { epm[3 * i] = e; epm[3 * i + 1] = p; epm[3 * i + 2] = m; }
        }
    }
    if (bg_lane() == 0) {
        *alive_out = alive ? 1 : 0;
        if (npf_out) *npf_out = npf;
        if (k_out) *k_out = am.k1;
        if (zw_out) for (int j = 0; j < 4; j++) zw_out[j] = z.a[j];
    }
}

// measurePauli on one state, result stored in active-mask layout
template <int NS>
BG_DEV int warp_measure_pauli(bg_state* g, uint64_t* Aout, int m, uint64_t zeta, uint64_t xi) {
    typedef typename WordOf<NS>::T W;
    Native<NS> st;
    native_load<NS>(st, g);
    const int r = native_measure<NS>(st, (uint32_t)m, (W)zeta, (W)xi);
    __syncwarp();
    native_store_raw<NS>(st, g, Aout);
    return r;
}

template <int NS>
BG_DEV void warp_random_state(int n, uint64_t seed, uint32_t bin, uint64_t sample, const double* cdf,
                              bg_state* out, uint64_t* Aout) {
    Native<NS> st;
    native_random<NS>(st, n, seed, bin, sample, cdf);
    native_store_raw<NS>(st, out, Aout);
}

}  // namespace bg
