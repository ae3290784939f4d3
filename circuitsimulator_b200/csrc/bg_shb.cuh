// bg_shb.cuh — shared high-block reduction for 32 < t <= 44 (|L> terms): the L x chi loop in 32-bit words.
//
// After the relabelling of bg_shb_plan.h the 32 terms a warp takes together have the SAME active set on the
// variables >= 32 ("high": t - 32 term variables + theta's parity checks as Lagrange variables), and differ
// only on the 32 low ones.  Summing out a high variable does not depend on which low variables a term keeps
// (an update that touches a variable the term fixes to 0 is simply never read), so the warp does it ONCE for
// the whole batch, cooperatively (shb_reduce: rows in lanes, __shfl_sync / __ballot_sync):
//   * a high variable with D in {2,6}: the rank-one step of bg_tpp.cuh (t_oddblock32), on all rows at once;
//   * two coupled high variables with D in {0,4}: the dimer step of exponentialSumExact (stabilizer.c:372-418);
//   * what is left are high variables with D in {0,4} that are not coupled to each other — each is a parity
//     check on the low variables, with its own Lagrange variable.  They are handed to the threads, which give
//     them a FREE low slot (a position the term does not use) while copying the reduced form in.
// Each thread then runs the odd-first elimination on at most 32 variables in 32-bit words (t_expsum_odd).
// Same value as innerProductExact (stabilizer.c:589-659) for every pair: the order in which variables are
// summed out does not change the sum.
#pragma once
#include "bg_tpp.cuh"
#include "bg_shb_plan.h"

namespace bg {

// one sample, relabelled: lane c holds the low row c; lane j < nht holds high row j (low part R, high part S)
struct ShbForm {
    uint32_t L, R, S;                       // lane-local
    uint32_t D1lo, D2lo, D1hi, D2hi, Q;     // warp-uniform
};
struct ShbOut {
    uint32_t D1, D2, Q, p;                  // the reduced form on the low variables (rows: see shb_reduce) and 2^(p/2)
    uint32_t left;                          // high variables that are left (mask), all with D in {0,4}, mutually uncoupled
    uint32_t left_d2;                       // their D2 bits (same positions)
};

BG_DEV int shb_top(uint32_t x) { return 31 - __clz((int)x); }

// Sum out the high variables in Eh (warp-cooperative).  Returns this lane's low row of the reduced form and
// its high row's low part (read by the caller for the variables in o.left).
BG_DEV void shb_reduce(ShbForm f, uint32_t Eh, uint32_t& Lout, uint32_t& Rout, ShbOut& o) {
    const int lane = bg_lane();
    uint32_t p = 0;
    while (true) {
        const uint32_t odd = f.D1hi & Eh;
        if (odd) {                                               // rank-one step on the top odd variable
            const int a = shb_top(odd);
            Eh ^= 1u << a;
            const uint32_t vlo = __shfl_sync(BG_FULL, f.R, a), vhi = __shfl_sync(BG_FULL, f.S, a) & Eh;
            const bool neg = (f.D2hi >> a) & 1u;
            f.Q += neg ? 7u : 1u; p += 1u;
            const uint32_t dm = neg ? 0u : ~0u;
            f.D2lo ^= vlo & (f.D1lo ^ dm); f.D2hi ^= vhi & (f.D1hi ^ dm);
            f.D1lo ^= vlo; f.D1hi ^= vhi;
            if ((vlo >> lane) & 1u) f.L ^= vlo;
            if ((vhi >> lane) & 1u) { f.R ^= vlo; f.S ^= vhi; }
            continue;
        }
        // every high variable left has D in {0,4}: a coupled pair?
        const uint32_t adj = __ballot_sync(BG_FULL, ((Eh >> lane) & 1u) && (f.S & Eh & ~(1u << lane)) != 0u);
        if (!adj) break;
        const int a = shb_top(adj);
        const uint32_t Sa = __shfl_sync(BG_FULL, f.S, a);
        const int b = shb_top(Sa & Eh & ~(1u << a));
        Eh &= ~((1u << a) | (1u << b));
        const uint32_t valo = __shfl_sync(BG_FULL, f.R, a), vahi = Sa & Eh;
        const uint32_t vblo = __shfl_sync(BG_FULL, f.R, b), vbhi = __shfl_sync(BG_FULL, f.S, b) & Eh;
        const uint32_t d2a = (f.D2hi >> a) & 1u, d2b = (f.D2hi >> b) & 1u;
        p += 2u; f.Q += 4u * (d2a & d2b);
        f.D2lo ^= (d2a ? vblo : 0u) ^ (d2b ? valo : 0u) ^ (valo & vblo);
        f.D2hi ^= (d2a ? vbhi : 0u) ^ (d2b ? vahi : 0u) ^ (vahi & vbhi);
        if ((valo >> lane) & 1u) f.L ^= vblo;
        if ((vblo >> lane) & 1u) f.L ^= valo;
        if ((vahi >> lane) & 1u) { f.R ^= vblo; f.S ^= vbhi; }
        if ((vbhi >> lane) & 1u) { f.R ^= valo; f.S ^= vahi; }
    }
    Lout = f.L; Rout = f.R;
    o.D1 = f.D1lo; o.D2 = f.D2lo; o.Q = f.Q & 7u; o.p = p; o.left = Eh; o.left_d2 = f.D2hi & Eh;
}


// ---- relabelling a sample (natural variable order, as k_prepare stores it) into a ShbForm
struct ShbPerm { int nh, nsw; uint8_t swp[SHB_MAXH], swq[SHB_MAXH]; uint8_t iperm[64]; };

// every swap exchanges a position p < 32 with a position q >= 32: bit p of the low word <-> bit q - 32 of the high word
BG_DEV uint64_t shb_perm_word(uint64_t w, const ShbPerm& pm) {
    uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
    for (int i = 0; i < pm.nsw; i++) {
        const uint32_t p = pm.swp[i], q = pm.swq[i] - 32u;
        const uint32_t x = ((lo >> p) ^ (hi >> q)) & 1u;
        lo ^= x << p; hi ^= x << q;
    }
    return ((uint64_t)hi << 32) | lo;
}
BG_DEV int shb_perm_index(int c, const ShbPerm& pm) { return pm.iperm[c]; }      // the variable that sits at position c

// J: t ambient rows, Cw / Cpend / Cbeta: the parity checks (row slots in Cpend), natural labels.  Warp-cooperative.
// Returns the number of checks (<= SHB_MAXLAM, the caller's routing guarantees it); they become the high
// variables nh .. nh+nlam-1 with D = 4 beta.
// Every lane relabels exactly two words: its low row, and one more — lane j < nh the high row j, lane nh + j the
// check j, lanes 30 / 31 the vectors D1 / D2 (nh + SHB_MAXLAM <= 30) — so the bit swaps are paid twice per item,
// not once per word.
BG_DEV int shb_load(const uint64_t* J, const uint64_t* Cw, uint64_t Cpend, uint64_t Cbeta, uint64_t D1, uint64_t D2,
                    uint32_t Q, int t, const ShbPerm& pm, ShbForm& f) {
    const int lane = bg_lane();
    const int nh = pm.nh;
    const uint64_t maskt = t >= 64 ? ~0ull : ((1ull << t) - 1ull);
    const uint32_t maskh = (1u << nh) - 1u;
    f.L = (uint32_t)shb_perm_word(J[shb_perm_index(lane, pm)] & maskt, pm);
    int nlam = __popcll(Cpend);
    if (nlam > SHB_MAXLAM) nlam = SHB_MAXLAM;
    // this lane's second word
    uint64_t src = 0;
    uint32_t my_beta = 0;
    if (lane < nh) src = J[shb_perm_index(32 + lane, pm)];
    else if (lane < nh + nlam) {
        uint64_t pend = Cpend;
        for (int j = nh; j < lane; j++) pend &= pend - 1ull;           // drop the checks of the lanes before this one
        const int b = __ffsll((long long)pend) - 1;
        src = Cw[b];
        my_beta = (uint32_t)((Cbeta >> b) & 1ull);
    } else if (lane == 30) src = D1;
    else if (lane == 31) src = D2;
    const uint64_t hw = shb_perm_word(src & maskt, pm);
    const uint32_t hlo = (uint32_t)hw, hhi = (uint32_t)(hw >> 32);
    const bool high = lane < nh + nlam;
    f.R = high ? hlo : 0u; f.S = high ? (hhi & maskh) : 0u;
    for (int j = 0; j < nlam; j++) {                             // column nh + j of the high rows = check j on the high variables
        const uint32_t sl = __shfl_sync(BG_FULL, f.S, nh + j);
        if (lane < nh) f.S |= ((sl >> lane) & 1u) << (nh + j);
    }
    const uint32_t beta_bits = __ballot_sync(BG_FULL, my_beta != 0u);   // bit nh + j = right-hand side of check j
    f.D1lo = __shfl_sync(BG_FULL, hlo, 30); f.D1hi = __shfl_sync(BG_FULL, hhi, 30) & maskh;
    f.D2lo = __shfl_sync(BG_FULL, hlo, 31); f.D2hi = (__shfl_sync(BG_FULL, hhi, 31) & maskh) | beta_bits;
    f.Q = Q;
    return nlam;
}

// What a warp shares about the current batch (shared memory, written by the warp after shb_reduce)
struct ShbBatch {
    const uint32_t* red;        // 32 reduced low rows (16-byte aligned)
    const uint32_t* left;       // low parts of the rows of the variables that are left, in descending order
    uint32_t D1, D2, Q;
    int p;                      // factor 2^(p/2) collected by the warp
    int nleft;
    uint32_t left_d2;           // bit i = D2 of leftover i
    int k1, nlam;
};

// copy the reduced form in, restricted to the term's low variables, with NL leftovers relocated to the free
// slots rel[i] (column rel[i] of row c = bit c of left[i])
template <int NL>
BG_HD void t_shb_copy_in(const Rows<uint32_t>& J, const ShbBatch& sb, uint32_t keep, const uint32_t (&rel)[4]) {
    uint32_t lw[NL > 0 ? NL : 1];
#pragma unroll
    for (int i = 0; i < NL; i++) lw[i] = sb.left[i];
#if defined(__CUDA_ARCH__)
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(sb.red);
#pragma unroll
    for (int q = 0; q < 32; q += 4) {
        uint32_t x[4];
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]) : "r"(src + 4u * (uint32_t)q));
#pragma unroll
        for (int e = 0; e < 4; e++) {
            uint32_t r = x[e] & keep;
#pragma unroll
            for (int i = 0; i < NL; i++) t_pxor32(r, lw[i] & (1u << (q + e)), rel[i]);
            J.put(q + e, r);
        }
    }
#else
    for (int q = 0; q < 32; q++) {
        uint32_t r = sb.red[q] & keep;
        for (int i = 0; i < NL; i++) if ((lw[i] >> q) & 1u) r ^= rel[i];
        J.put(q, r);
    }
#endif
}

// <phi|theta> for one |L> term of the batch.  A = the term's variables (relabelled): low word = the active low
// variables, high word = the batch's high pattern.
// nlmax: the largest number of leftovers among the batches the lanes of this warp work on (warp-uniform; selects the
// copy-in variant), sb.nleft: this thread's own.
BG_HD void t_term_shb(const Rows<uint32_t>& J, const ShbBatch& sb, uint64_t A, int nlmax, int& eps, int& p, int& m) {
    const uint32_t Alo = (uint32_t)A;
    const int k2 = tpopc(A);
    TF<uint32_t> f;
    f.D1 = sb.D1; f.D2 = sb.D2; f.Q = sb.Q; f.A = Alo;
    uint32_t rel[4] = {0u, 0u, 0u, 0u};
    uint32_t fr = ~Alo;
    const int nrel = sb.nleft < SHB_RELOC ? sb.nleft : SHB_RELOC;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i < nrel) {
            const uint32_t fb = 1u << thighest(fr);              // the host checked that every term has SHB_RELOC free slots
            fr ^= fb; rel[i] = fb;
            f.A |= fb; f.D1 &= ~fb;
            f.D2 = (f.D2 & ~fb) | (((sb.left_d2 >> i) & 1u) ? fb : 0u);
        }
    }
    switch (nlmax < SHB_RELOC ? nlmax : SHB_RELOC) {                  // rel[i] = 0 beyond this thread's own leftovers
        case 0: t_shb_copy_in<0>(J, sb, Alo, rel); break;
        case 1: t_shb_copy_in<1>(J, sb, Alo, rel); break;
        case 2: t_shb_copy_in<2>(J, sb, Alo, rel); break;
        case 3: t_shb_copy_in<3>(J, sb, Alo, rel); break;
        default: t_shb_copy_in<4>(J, sb, Alo, rel); break;
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (i < nrel) J.put(thighest(rel[i]), sb.left[i] & Alo);   // the leftover's own row (uncoupled to the other leftovers)
    int extra = 0;
    bool dead = false;
    if (sb.nleft > SHB_RELOC) {                                     // (rare) more leftovers than slots: pivot their checks
        uint32_t hs[SHB_MAXHT];                                     // earlier pivoted checks (substituted into later ones)
        uint32_t hb = 0;
        for (int i = SHB_RELOC; i < sb.nleft; i++) {
            uint32_t w = sb.left[i] & Alo;
            uint32_t beta = (sb.left_d2 >> i) & 1u;
            for (int q = SHB_RELOC; q < i; q++)
                if (hs[q] && tget(w, thighest(hs[q]))) { w ^= hs[q]; beta ^= (hb >> q) & 1u; }
            w &= f.A;
            hs[i] = w; hb |= beta << i;
            extra += 2;                                             // the sum over the leftover itself: 2 [w.x = beta]
            if (w == 0u) { if (beta) dead = true; continue; }
            t_pivot<uint32_t>(J, f, w, beta);
        }
    }
    if (dead) { eps = 0; p = 0; m = 0; return; }
    t_expsum_odd(J, f, eps, p, m);
    if (eps) p += sb.p + extra - sb.k1 - k2 - 2 * sb.nlam; else { p = 0; m = 0; }
}

}  // namespace bg
