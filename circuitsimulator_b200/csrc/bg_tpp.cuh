// bg_tpp.cuh — the L x chi hot loop, one THREAD per inner product.
//
// A warp takes one projected theta (its ambient quadratic form, see bg_device.cuh) and 32
// decomposition terms at a time: lane j evaluates <phi_{i0+j} | theta>.  Every quantity that is
// warp-uniform in the warp-per-pair formulation (D, the active set, the pivot rows, ...) is here a
// per-thread register, so no lane-op is spent on replicated work; the only per-thread array is the
// working copy of J (t rows of one word), kept in shared memory in a [row][thread] layout — bank =
// thread, so the data-dependent row index never causes a bank conflict.
//
// Why: ncu on the warp-per-pair kernel (profiles/r1_v1_warp_per_pair_ncu_summary.json) shows the
// integer ALU pipe 95 % busy at 1266 warp-instructions per inner product — a t = 40 term has only
// ~20 live rows for 64 row slots and every uniform mask update is executed by all 32 lanes.  The
// same algebra per thread needs ~2000 thread-instructions per pair, i.e. ~60 warp-instructions.
//
// Same mathematics as bg_device.cuh (pivot / basis_change / expsum), same reference anchors:
// shrink (stabilizer.c:500-585), updateDJ/updateQD (:129-177), exponentialSumExact (:300-481),
// innerProductExact (:589-659), prepH/prepL (stateprep.c:36-120).
#pragma once
#include "bg_device.cuh"

// Work accounting (CPU build of this header only, -DBG_COUNT_WORK): the algorithm's own operation
// counts, from which DESIGN.md's "algorithmic lane-ops per inner product" is computed.
#if defined(BG_COUNT_WORK) && !defined(__CUDACC__)
struct BgWork { unsigned long long xors, rows, dimers, monomers, basis_changes, pairs; };
extern BgWork g_bg_work;
#define BG_WORK(field, n) (g_bg_work.field += (n))
extern int* g_bg_trace; extern int g_bg_trace_n, g_bg_trace_cap;     // per t_xor2 call: |M1|M2| (lo), (hi)
#define BG_TRACE(lo, hi) do { if (g_bg_trace && g_bg_trace_n + 2 <= g_bg_trace_cap) { g_bg_trace[g_bg_trace_n++] = (lo); g_bg_trace[g_bg_trace_n++] = (hi); } } while (0)
#else
#define BG_WORK(field, n) ((void)0)
#define BG_TRACE(lo, hi) ((void)0)
#endif

namespace bg {

// per-thread view of its working rows: row r lives at base[r * stride]
template <typename W> struct Rows {
    W* base;
    int stride;
    uint32_t sbase, sstride;     // device: shared-memory byte address of row 0, byte stride between rows
    BG_HDM W get(int r) const { return base[(size_t)r * stride]; }
    BG_HDM void put(int r, W v) const { base[(size_t)r * stride] = v; }
    BG_HDM void xr(int r, W v) const { base[(size_t)r * stride] ^= v; }
};

template <typename W> struct TF {      // per-thread quadratic form scalars (J is in Rows)
    W D1, D2, A;
    uint32_t Q;
};

BG_HD int tlowest(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
BG_HD int tlowest(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
BG_HD int tpopc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
BG_HD int tpopc(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
template <typename W> BG_HD W tbit(int i) { return (W)1 << i; }
template <typename W> BG_HD W tfill(uint32_t b) { return (W)0 - (W)(b & 1u); }
template <typename W> BG_HD uint32_t tget(W x, int i) { return (uint32_t)(x >> i) & 1u; }
template <typename W> BG_HD W tlowmask(int n) { return n >= (int)(8 * sizeof(W)) ? ~(W)0 : (((W)1 << n) - 1); }

// position of the highest set bit of a non-zero word: one FLO on the device
BG_HD int thighest(uint32_t x) {
#if defined(__CUDA_ARCH__)
    int r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
#else
    return 31 - __builtin_clz(x);
#endif
}
BG_HD int thighest(uint64_t x) {
    const uint32_t hi = (uint32_t)(x >> 32);
    return hi ? 32 + thighest(hi) : thighest((uint32_t)x);
}

// row_c ^= [c in M1] V1 ^ [c in M2] V2  for every c in M1 | M2 — the one primitive every step of the
// elimination reduces to.  A single loop (not one per mask) keeps the trip counts of the 32 lanes
// of a warp close together; words are walked in 32-bit halves from the top bit down (FLO + 2 ops).
BG_HD void t_xor2(const Rows<uint32_t>& J, uint32_t M1, uint32_t V1, uint32_t M2, uint32_t V2) {
    uint32_t U = M1 | M2;
    BG_TRACE(tpopc(U), 0);
#if defined(__CUDA_ARCH__)
    while (U) {
        const uint32_t c = (uint32_t)thighest(U);
        const uint32_t b = 1u << c;
        U ^= b;
        const uint32_t addr = c * J.sstride + J.sbase;
        uint32_t r;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
        asm("{\n\t.reg .pred p, q;\n\t"
            "setp.ne.u32 p, %1, 0;\n\tsetp.ne.u32 q, %2, 0;\n\t"
            "@p xor.b32 %0, %0, %3;\n\t@q xor.b32 %0, %0, %4;\n\t}"
            : "+r"(r) : "r"(M1 & b), "r"(M2 & b), "r"(V1), "r"(V2));
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(r) : "memory");
    }
#else
    while (U) {
        const int c = thighest(U);
        const uint32_t b = 1u << c;
        U ^= b;
        uint32_t r = J.get(c);
        if (M1 & b) r ^= V1;
        if (M2 & b) r ^= V2;
        J.put(c, r);
        BG_WORK(rows, 1); BG_WORK(xors, ((M1 & b) ? 1 : 0) + ((M2 & b) ? 1 : 0));
    }
#endif
}
BG_HD void t_xor2(const Rows<uint64_t>& J, uint64_t M1, uint64_t V1, uint64_t M2, uint64_t V2) {
    BG_TRACE(tpopc((uint32_t)(M1 | M2)), tpopc((uint32_t)((M1 | M2) >> 32)));
#if defined(__CUDA_ARCH__)
    // Hand-scheduled inner loop (31 % of the kernel's issue slots): per row one FLO, one shift, one
    // IMAD for the shared-memory address, LDS.64, two mask tests, four PREDICATED xors, STS.64.
    const uint32_t v1l = (uint32_t)V1, v1h = (uint32_t)(V1 >> 32), v2l = (uint32_t)V2, v2h = (uint32_t)(V2 >> 32);
#pragma unroll
    for (int h = 1; h >= 0; h--) {
        const uint32_t m1 = (uint32_t)(M1 >> (32 * h)), m2 = (uint32_t)(M2 >> (32 * h));
        const uint32_t hbase = J.sbase + (uint32_t)(32 * h) * J.sstride;
        uint32_t U = m1 | m2;
        while (U) {
            const uint32_t c = (uint32_t)thighest(U);
            const uint32_t b = 1u << c;
            U ^= b;
            const uint32_t addr = c * J.sstride + hbase;
            uint32_t lo, hi;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr));
            asm("{\n\t.reg .pred p, q;\n\t"
                "setp.ne.u32 p, %2, 0;\n\tsetp.ne.u32 q, %3, 0;\n\t"
                "@p xor.b32 %0, %0, %4;\n\t@p xor.b32 %1, %1, %5;\n\t"
                "@q xor.b32 %0, %0, %6;\n\t@q xor.b32 %1, %1, %7;\n\t}"
                : "+r"(lo), "+r"(hi) : "r"(m1 & b), "r"(m2 & b), "r"(v1l), "r"(v1h), "r"(v2l), "r"(v2h));
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(addr), "r"(lo), "r"(hi) : "memory");
        }
    }
#else
#pragma unroll
    for (int h = 1; h >= 0; h--) {
        const uint32_t m1 = (uint32_t)(M1 >> (32 * h)), m2 = (uint32_t)(M2 >> (32 * h));
        uint32_t U = m1 | m2;
        while (U) {
            const int c = thighest(U);
            const uint32_t b = 1u << c;
            U ^= b;
            uint64_t r = J.get(c + 32 * h);
            if (m1 & b) r ^= V1;
            if (m2 & b) r ^= V2;
            J.put(c + 32 * h, r);
            BG_WORK(rows, 1); BG_WORK(xors, ((m1 & b) ? 1 : 0) + ((m2 & b) ? 1 : 0));
        }
    }
#endif
}

// x_i = x'_i + sum_{a in Sp} x'_a.   (bg_device.cuh: basis_change)   Returns the old row i.
template <typename W> BG_HD W t_basis_change(const Rows<W>& J, TF<W>& f, int i, W Sp) {
    const W bi = tbit<W>(i);
    const W Ji = J.get(i);
    // row a += row i (a in Sp), then column a += column i: rows r whose (updated) entry (r,i) is
    // set get ^= Sp.  By symmetry that column is row i, plus J_ii on the rows of Sp.
    const W col = (Ji ^ ((Ji & bi) ? Sp : (W)0)) & f.A;
    t_xor2(J, Sp, Ji, col, Sp);
    BG_WORK(basis_changes, 1);
    const W d1i = tfill<W>(tget(f.D1, i)), d2i = tfill<W>(tget(f.D2, i));
    f.D2 ^= Sp & (d2i ^ (d1i & f.D1) ^ Ji);
    f.D1 ^= Sp & d1i;
    return Ji;
}

// impose sum_{a in S} x_a = beta, eliminating x_i, i = lowest(S)    (bg_device.cuh: pivot)
template <typename W> BG_HD void t_pivot(const Rows<W>& J, TF<W>& f, W S, uint32_t beta) {
    const int i = thighest(S);
    const W bi = tbit<W>(i), Sp = S ^ bi;
    const uint32_t d1 = tget(f.D1, i), d2 = tget(f.D2, i);
    const W Ji = t_basis_change<W>(J, f, i, Sp);
    if (beta) {
        f.Q = (f.Q + 2u * d1 + 4u * d2) & 7u;
        f.D2 ^= Ji ^ (Sp & tfill<W>(d1));
    }
    f.A &= ~bi;
}

// The monomer / dimer rounds of the exponential sum on the variables in E (all with D in {0,4}).
// HI_ONLY (64-bit words): stop as soon as no variable >= 32 is left — the caller continues with
// 32-bit words on the low halves of the same rows, at half the cost per round.
template <typename W, bool HI_ONLY>
BG_HD void t_rounds(const Rows<W>& J, W& E, W& D2, W& Js, uint32_t& cnt, uint32_t& neg0, uint32_t& neg1,
                    uint32_t& z0, uint32_t& z1, bool has_s) {
    while (HI_ONLY ? ((uint64_t)E >> 32) != 0 : E != 0) {
        const int a = thighest(E);
        const W ba = tbit<W>(a);
        const W Ja = J.get(a) & E & ~ba;
        const uint32_t d2a = tget(D2, a), sa = tget(Js, a);
        if (Ja == 0) {                               // monomer {a}
            z0 |= d2a; z1 |= d2a ^ sa; cnt++;
            E ^= ba;
            BG_WORK(monomers, 1);
            if (z0 && (z1 || !has_s)) { E = 0; break; }      // the whole sum is zero
            continue;
        }
        const int b = thighest(Ja);                  // dimer {a,b}
        const W bb = tbit<W>(b);
        const W Jb = J.get(b) & E & ~bb;
        const W rest = E & ~(ba | bb);
        const uint32_t d2b = tget(D2, b), sb = tget(Js, b);
        neg0 ^= d2a & d2b; neg1 ^= (d2a ^ sa) & (d2b ^ sb); cnt++;
        const W Jar = Ja & rest, Jbr = Jb & rest;
        BG_WORK(dimers, 1);
        t_xor2(J, Jar, Jbr, Jbr, Jar);                                  // J_c ^= [J_ca] J_b ^ [J_cb] J_a
        D2 ^= (Jar & tfill<W>(d2b)) ^ (Jbr & tfill<W>(d2a)) ^ (Jar & Jbr);
        Js ^= (Jar & tfill<W>(sb)) ^ (Jbr & tfill<W>(sa));
        E = rest;
    }
}

// sum over F_2^A of e^{i pi q/4}    (bg_device.cuh: expsum)
template <typename W> BG_HD void t_expsum(const Rows<W>& J, TF<W>& f, int& eps, int& p, int& m) {
    const W A = f.A;
    const W S = f.D1 & A;
    const bool has_s = S != 0;
    W E = A, Js = 0;
    uint32_t Ds = 0;
    if (has_s) {
        const int s = thighest(S);
        const W bs = tbit<W>(s), Sp = S ^ bs;
        Ds = 2u + 4u * tget(f.D2, s);
        if (Sp) t_basis_change<W>(J, f, s, Sp);
        E = A & ~bs;
        Js = J.get(s) & E;
    }
    W D2 = f.D2;
    uint32_t cnt = 0, neg0 = 0, neg1 = 0, z0 = 0, z1 = 0;
    // (Measured and rejected: finishing the variables >= 32 first and then continuing with 32-bit words
    // on the low halves — fewer instructions per lane, but the extra reconvergence point costs more
    // than it saves: 4.83 vs 4.67 ms at t = 40.)
    t_rounds<W, false>(J, E, D2, Js, cnt, neg0, neg1, z0, z1, has_s);
    p = 2 * (int)cnt;
    const uint32_t m0 = (f.Q + 4u * neg0) & 7u;
    if (!has_s) { eps = z0 ? 0 : 1; m = (int)m0; return; }
    const uint32_t m1 = (f.Q + Ds + 4u * neg1) & 7u;
    if (z0 && z1) { eps = 0; m = 0; p = 0; return; }
    eps = 1;
    if (z0) { m = (int)m1; return; }
    if (z1) { m = (int)m0; return; }
    const uint32_t diff = (m1 - m0) & 7u;
    p += 1;
    m = (int)((m0 + (diff == 2u ? 1u : 7u)) & 7u);
}

// What a warp shares about its theta: the ambient form and at most TPP_MAXC parity checks.
#define TPP_MAXC 6
template <typename W> struct TShared {
    const W* J;          // t ambient rows (shared memory, read-only)
    W D1, D2;
    uint32_t Q;
    int k1, t;
    int ncons;           // parity checks (t - k1)
    W cw[TPP_MAXC];      // the checks when ncons <= TPP_MAXC (register copy)
    uint32_t cbeta;      // bit j = right-hand side of check j (ncons <= TPP_MAXC)
    const W* cwv;        // all checks (shared memory) and their right-hand sides, any ncons <= t
    W cbetav;
};

// membership checks of K_theta, restricted to this term's active set.  Earlier pivots are
// substituted lazily (check j is rewritten with the pivots of checks < j).  `mrg` != 0: variables
// 2j in mrg were merged into 2j+1 beforehand (prepH), so bit 2j of a check moves to bit 2j+1.
template <typename W>
BG_HD bool t_constraints(const Rows<W>& J, TF<W>& f, const TShared<W>& sh, W mrg) {
    W hs[TPP_MAXC];
    uint32_t hb = 0;
#pragma unroll
    for (int j = 0; j < TPP_MAXC; j++) {
        if (j >= sh.ncons) break;
        W w = sh.cw[j];
        w ^= (w & mrg) << 1;
        uint32_t beta = (sh.cbeta >> j) & 1u;
#pragma unroll
        for (int q = 0; q < TPP_MAXC; q++) {            // substitute the pivots of checks q < j, in order
            if (q >= j) break;
            if (hs[q] && tget(w, thighest(hs[q]))) { w ^= hs[q]; beta ^= (hb >> q) & 1u; }
        }
        w &= f.A;
        hs[j] = w;                                       // 0 when the check is already implied
        hb |= beta << j;
        if (w == 0) { if (beta) return false; continue; }
        t_pivot<W>(J, f, w, beta);
    }
    return true;
}

// Same with any number of checks: the pivot history lives in the thread's own rows t .. t+ncons-1
// (shared memory) instead of registers.  A check only needs rewriting when it contains a variable
// eliminated by an earlier check (mask `gone`).
template <typename W>
BG_HD bool t_constraints_many(const Rows<W>& J, TF<W>& f, const TShared<W>& sh, W mrg) {
    const int t = sh.t;
    W hb = 0, gone = 0;
    for (int j = 0; j < sh.ncons; j++) {
        W w = sh.cwv[j];
        w ^= (w & mrg) << 1;
        uint32_t beta = tget(sh.cbetav, j);
        if (w & gone) {
            for (int q = 0; q < j; q++) {
                const W hq = J.get(t + q);
                if (hq && tget(w, thighest(hq))) { w ^= hq; beta ^= tget(hb, q); }
            }
        }
        w &= f.A;
        J.put(t + j, w);
        hb |= (W)beta << j;
        if (w == 0) { if (beta) return false; continue; }
        gone |= tbit<W>(thighest(w));
        t_pivot<W>(J, f, w, beta);
    }
    return true;
}

// <phi|theta> for a |L> term (prepL): |+> on supp(xt), |0> elsewhere.
template <typename W, bool MANYC = false>
BG_HD void t_term_L(const Rows<W>& J, const TShared<W>& sh, W xt, int& eps, int& p, int& m) {
    const int t = sh.t;
#pragma unroll 8
    for (int q = 0; q < t; q++) J.put(q, sh.J[q]);
    TF<W> f;
    f.D1 = sh.D1; f.D2 = sh.D2; f.Q = sh.Q;
    f.A = xt & tlowmask<W>(t);
    const int k2 = tpopc(f.A);
    if (!(MANYC ? t_constraints_many<W>(J, f, sh, (W)0) : t_constraints<W>(J, f, sh, (W)0))) { eps = 0; p = 0; m = 0; return; }
    t_expsum<W>(J, f, eps, p, m);
    if (eps) p -= sh.k1 + k2; else { p = 0; m = 0; }
}

// <phi|theta> for a |H^t> term (prepH); e1 as in bg_device.cuh: term_H.
template <typename W, bool MANYC = false>
BG_HD void t_term_H(const Rows<W>& J, const TShared<W>& sh, W e1, int& eps, int& p, int& m) {
    const int t = sh.t;
    const W maskt = tlowmask<W>(t);
    const W pairs = (W)0x5555555555555555ull & (maskt >> 1);
    const W mrg = e1 & pairs, cz = ~e1 & pairs;
    const W last = (t & 1) ? (e1 & tbit<W>(t - 1)) : (W)0;
    const W cz2 = cz | (cz << 1);
    for (int q = 0; q < t; q++) J.put(q, sh.J[q] ^ (((cz2 >> q) & 1) ? tbit<W>(q ^ 1) : (W)0));     // q1 - q2
    TF<W> f;
    f.D1 = sh.D1; f.D2 = sh.D2; f.Q = sh.Q;
    f.A = maskt & ~last;
    for (W r = mrg; r; r &= r - 1) {                    // x_{2j} = x_{2j+1}: eliminate x_{2j}
        const int i = tlowest(r);
        t_basis_change<W>(J, f, i, tbit<W>(i + 1));
        f.A &= ~tbit<W>(i);
    }
    const int k2 = t - tpopc(mrg) - tpopc(last);
    if (!(MANYC ? t_constraints_many<W>(J, f, sh, mrg) : t_constraints<W>(J, f, sh, mrg))) { eps = 0; p = 0; m = 0; return; }
    t_expsum<W>(J, f, eps, p, m);
    if (eps) p -= sh.k1 + k2; else { p = 0; m = 0; }
}

// ---------------------------------------------------------------- left-looking ("lazy") variant
// The eager routines above update every remaining row of J after each elimination step; the trip
// counts of those row loops differ from lane to lane (ncu: 20.8 of 32 lanes active), and each touched
// row costs an LDS + STS + loop overhead.  Every update has the form
//        row_c ^= [c in M1] V1 ^ [c in M2] V2        with masks that do not depend on the row,
// so instead of applying it to all rows we only REMEMBER (M1,V1,M2,V2) and materialise a row when it
// becomes a pivot:  row_c = ambient_c ^ (all remembered updates whose mask contains c).  Per round two
// rows are materialised, each with a loop over the history whose length is the round number — the
// SAME for all lanes of a warp.  No per-thread copy of J exists at all; the thread's shared-memory
// rows hold the dimer history (two words per dimer, at most t/2 dimers).
// Used for |L> terms with at most LZ_MAXB - 1 parity checks; everything else takes the eager path.
// Measured (B200, round 1): 5.32 vs 5.78 ms at t = 60 but 5.92 vs 4.67 ms at t = 40, where the unrolled
// basis-change entries and the history loads cost more than the divergence they remove — so it is an
// option (BG_LAZY=1), not the default.
#define LZ_MAXB 5                      // basis changes kept in registers: up to 4 check pivots + the fold
template <typename W> struct LzBC { W Sp, Ji, col; };

template <typename W>
BG_HD W lz_row(const TShared<W>& sh, int c, const LzBC<W> (&bc)[LZ_MAXB], int nb, const Rows<W>& H, int r) {
    W row = sh.J[c];
    const W bc_ = tbit<W>(c);
#pragma unroll
    for (int j = 0; j < LZ_MAXB; j++) {
        if (j < nb) {
            if (bc[j].Sp & bc_) row ^= bc[j].Ji;
            if (bc[j].col & bc_) row ^= bc[j].Sp;
        }
    }
    for (int q = 0; q < r; q++) {
        const W X = H.get(2 * q), Y = H.get(2 * q + 1);
        if (X & bc_) row ^= Y;
        if (Y & bc_) row ^= X;
    }
    BG_WORK(rows, 1); BG_WORK(xors, r);
    return row;
}

// x_i = x'_i + sum_{a in Sp} x'_a, remembered instead of applied (cf. t_basis_change)
template <typename W>
BG_HD W lz_basis_change(const TShared<W>& sh, TF<W>& f, int i, W Sp, LzBC<W> (&bc)[LZ_MAXB], int& nb, const Rows<W>& H) {
    const W bi = tbit<W>(i);
    const W Ji = lz_row<W>(sh, i, bc, nb, H, 0);
    const W col = (Ji ^ ((Ji & bi) ? Sp : (W)0)) & f.A;
#pragma unroll
    for (int j = 0; j < LZ_MAXB; j++) if (j == nb) { bc[j].Sp = Sp; bc[j].Ji = Ji; bc[j].col = col; }
    nb++;
    const W d1i = tfill<W>(tget(f.D1, i)), d2i = tfill<W>(tget(f.D2, i));
    f.D2 ^= Sp & (d2i ^ (d1i & f.D1) ^ Ji);
    f.D1 ^= Sp & d1i;
    BG_WORK(basis_changes, 1);
    return Ji;
}

template <typename W>
BG_HD void lz_pivot(const TShared<W>& sh, TF<W>& f, W S, uint32_t beta, LzBC<W> (&bc)[LZ_MAXB], int& nb, const Rows<W>& H) {
    const int i = thighest(S);
    const W bi = tbit<W>(i), Sp = S ^ bi;
    const uint32_t d1 = tget(f.D1, i), d2 = tget(f.D2, i);
    const W Ji = lz_basis_change<W>(sh, f, i, Sp, bc, nb, H);
    if (beta) {
        f.Q = (f.Q + 2u * d1 + 4u * d2) & 7u;
        f.D2 ^= Ji ^ (Sp & tfill<W>(d1));
    }
    f.A &= ~bi;
}

// <phi|theta> for a |L> term, left-looking.  H: the thread's scratch rows (>= t words).
// Requires sh.ncons <= LZ_MAXB - 1.
template <typename W>
BG_HD void t_term_L_lazy(const Rows<W>& H, const TShared<W>& sh, W xt, int& eps, int& p, int& m) {
    const int t = sh.t;
    TF<W> f;
    f.D1 = sh.D1; f.D2 = sh.D2; f.Q = sh.Q;
    f.A = xt & tlowmask<W>(t);
    const int k2 = tpopc(f.A);
    LzBC<W> bc[LZ_MAXB];
    int nb = 0;
    // parity checks (cf. t_constraints): earlier pivots are substituted lazily
    {
        W hs[LZ_MAXB - 1];
        uint32_t hb = 0;
#pragma unroll
        for (int j = 0; j < LZ_MAXB - 1; j++) {
            if (j >= sh.ncons) break;
            W w = sh.cw[j];
            uint32_t beta = (sh.cbeta >> j) & 1u;
#pragma unroll
            for (int q = 0; q < LZ_MAXB - 1; q++) {
                if (q >= j) break;
                if (hs[q] && tget(w, thighest(hs[q]))) { w ^= hs[q]; beta ^= (hb >> q) & 1u; }
            }
            w &= f.A;
            hs[j] = w;
            hb |= beta << j;
            if (w == 0) { if (beta) { eps = 0; p = 0; m = 0; return; } continue; }
            lz_pivot<W>(sh, f, w, beta, bc, nb, H);
        }
    }
    // exponential sum (cf. t_expsum / t_rounds)
    const W A = f.A;
    const W S = f.D1 & A;
    const bool has_s = S != 0;
    W E = A, Js = 0;
    uint32_t Ds = 0;
    if (has_s) {
        const int s = thighest(S);
        const W bs = tbit<W>(s), Sp = S ^ bs;
        Ds = 2u + 4u * tget(f.D2, s);
        if (Sp) lz_basis_change<W>(sh, f, s, Sp, bc, nb, H);
        E = A & ~bs;
        Js = lz_row<W>(sh, s, bc, nb, H, 0) & E;
    }
    W D2 = f.D2;
    uint32_t cnt = 0, neg0 = 0, neg1 = 0, z0 = 0, z1 = 0;
    int r = 0;
    while (E) {
        const int a = thighest(E);
        const W ba = tbit<W>(a);
        const W Ja = lz_row<W>(sh, a, bc, nb, H, r) & E & ~ba;
        const uint32_t d2a = tget(D2, a), sa = tget(Js, a);
        if (Ja == 0) {
            z0 |= d2a; z1 |= d2a ^ sa; cnt++;
            E ^= ba;
            BG_WORK(monomers, 1);
            if (z0 && (z1 || !has_s)) break;
            continue;
        }
        const int b = thighest(Ja);
        const W bb = tbit<W>(b);
        const W Jb = lz_row<W>(sh, b, bc, nb, H, r) & E & ~bb;
        const W rest = E & ~(ba | bb);
        const uint32_t d2b = tget(D2, b), sb = tget(Js, b);
        neg0 ^= d2a & d2b; neg1 ^= (d2a ^ sa) & (d2b ^ sb); cnt++;
        const W Jar = Ja & rest, Jbr = Jb & rest;
        BG_WORK(dimers, 1);
        H.put(2 * r, Jar); H.put(2 * r + 1, Jbr);                      // remember: row_c ^= [Jar_c] Jbr ^ [Jbr_c] Jar
        r++;
        D2 ^= (Jar & tfill<W>(d2b)) ^ (Jbr & tfill<W>(d2a)) ^ (Jar & Jbr);
        Js ^= (Jar & tfill<W>(sb)) ^ (Jbr & tfill<W>(sa));
        E = rest;
    }
    p = 2 * (int)cnt;
    const uint32_t m0 = (f.Q + 4u * neg0) & 7u;
    if (!has_s) { eps = z0 ? 0 : 1; m = (int)m0; }
    else {
        const uint32_t m1 = (f.Q + Ds + 4u * neg1) & 7u;
        if (z0 && z1) { eps = 0; }
        else {
            eps = 1;
            if (z0) m = (int)m1;
            else if (z1) m = (int)m0;
            else { p += 1; m = (int)((m0 + ((((m1 - m0) & 7u) == 2u) ? 1u : 7u)) & 7u); }
        }
    }
    if (eps) p -= sh.k1 + k2; else { p = 0; m = 0; }
}

}  // namespace bg
