// bg_device.cuh — warp-cooperative stabilizer algebra for sm_100a.
//
// One warp owns one stabilizer state / one quadratic form.  Bit matrices live
// one row per lane (row v = lane + 32*s, s < NS), rows are 32-bit words when the
// state width t <= 32 (NS = 1) and 64-bit words when t <= 64 (NS = 2).  All
// vectors (h, D1, D2, active masks) are warp-uniform registers.  Cross-lane
// traffic is __ballot_sync / __shfl_sync / __reduce_xor_sync only; all
// arithmetic is LOP3 / SHF / POPC / FLO on registers.  No shared or global
// memory is touched by anything in this header except the explicit load/store
// helpers at the bottom.
//
// What it computes is the reference's stabilizer algebra
// (libcirc/stabilizer/stabilizer.c of patrickrall/CircuitSimulator), re-derived
// for this layout; each routine cites the reference routine whose result it
// must reproduce.  The algorithms are NOT the reference's:
//   * no row swaps: the affine space is span{G_a : a in A} for an active mask A;
//   * a state is turned ONCE into an "ambient" quadratic form q^(x) on the
//     computational-basis bits x in F_2^t plus (t-k) parity checks, so that the
//     inner product with a decomposition term (a product of |0>/|+>, or of
//     2-qubit blocks) needs no elimination for the term's own constraints —
//     restricting to supp(x~) is a mask — only the few parity checks are
//     pivoted, followed by the Z8 exponential sum (eq. 63-68 of arXiv:1601.07601)
//     done as a symplectic (dimer) elimination of J with all rows in parallel.
//
// The header also compiles for the host against tests/emu/cpu_warp.h (a
// 32-fibre lock-step warp emulator) so that the SAME source is checked against
// the CPU oracle in the `-m "not gpu"` test-suite.  That build is test
// infrastructure; the product library is CUDA only.
#pragma once
#include <stdint.h>
#include "bgnorm.h"

#if defined(__CUDACC__)
#define BG_DEV __device__ __forceinline__
#define BG_HD __host__ __device__ __forceinline__
#define BG_HDM __host__ __device__ __forceinline__
BG_DEV int bg_lane() { return (int)(threadIdx.x & 31u); }
#else
#include "cpu_warp.h"   // tests/emu: __shfl_sync, __ballot_sync, __reduce_xor_sync, __popc, __ffs, ... bg_lane()
#define BG_DEV static inline
#define BG_HD static inline
#define BG_HDM inline
#endif

#define BG_FULL 0xffffffffu

namespace bg {

template <int NS> struct WordOf { typedef uint32_t T; };
template <> struct WordOf<2> { typedef uint64_t T; };

// ---------------------------------------------------------------- word helpers
BG_DEV int popcw(uint32_t x) { return __popc(x); }
BG_DEV int popcw(uint64_t x) { return __popcll(x); }
BG_DEV uint32_t parw(uint32_t x) { return (uint32_t)__popc(x) & 1u; }
BG_DEV uint32_t parw(uint64_t x) { return (uint32_t)__popc((uint32_t)x ^ (uint32_t)(x >> 32)) & 1u; }
BG_DEV int lowestw(uint32_t x) { return __ffs((int)x) - 1; }
BG_DEV int lowestw(uint64_t x) { return __ffsll((long long)x) - 1; }
template <typename W> BG_DEV W bitw(int i) { return (W)1 << i; }
template <typename W> BG_DEV W lowmaskw(int n) { return n >= (int)(8 * sizeof(W)) ? ~(W)0 : (((W)1 << n) - 1); }
template <typename W> BG_DEV W fillw(uint32_t b) { return (W)0 - (W)(b & 1u); }          // 0 or all-ones
template <typename W> BG_DEV uint32_t getw(W x, int i) { return (uint32_t)(x >> i) & 1u; }

BG_DEV uint32_t shflw(uint32_t v, int src) { return __shfl_sync(BG_FULL, v, src); }
BG_DEV uint64_t shflw(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(BG_FULL, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(BG_FULL, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
BG_DEV uint32_t shflxw(uint32_t v, int m) { return __shfl_xor_sync(BG_FULL, v, m); }
BG_DEV uint64_t shflxw(uint64_t v, int m) {
    uint32_t lo = __shfl_xor_sync(BG_FULL, (uint32_t)v, m);
    uint32_t hi = __shfl_xor_sync(BG_FULL, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}
BG_DEV uint32_t xorredw(uint32_t v) { return __reduce_xor_sync(BG_FULL, v); }
BG_DEV uint64_t xorredw(uint64_t v) {
    uint32_t lo = __reduce_xor_sync(BG_FULL, (uint32_t)v);
    uint32_t hi = __reduce_xor_sync(BG_FULL, (uint32_t)(v >> 32));
    return ((uint64_t)hi << 32) | lo;
}

// one predicate per row slot -> mask over rows
template <int NS> BG_DEV typename WordOf<NS>::T ballotw(const bool (&p)[NS]) {
    typedef typename WordOf<NS>::T W;
    W r = (W)__ballot_sync(BG_FULL, p[0]);
    if (NS == 2) r |= (W)((uint64_t)__ballot_sync(BG_FULL, p[NS - 1]) << 32 * (NS - 1));
    return r;
}
// broadcast row v (uniform) of a rows-in-lanes matrix
template <int NS> BG_DEV typename WordOf<NS>::T rowb(const typename WordOf<NS>::T (&r)[NS], int v) {
    typename WordOf<NS>::T x = r[0];
    if (NS == 2 && v >= 32) x = r[NS - 1];
    return shflw(x, v & 31);
}

// Transpose a rows-in-lanes bit matrix in place: out row v, bit u = in row u, bit v.
// 32 x 32 blocks are transposed with the log-step butterfly (5 x (SHFL.BFLY + 4 LOP3/SHF)); the
// 64 x 64 case is four such blocks plus a register swap of the off-diagonal ones.
BG_DEV uint32_t transpose32(uint32_t x) {
    const int lane = bg_lane();
    uint32_t m = 0x0000FFFFu;
#pragma unroll
    for (int j = 16; j; j >>= 1, m ^= m << j) {
        const uint32_t y = __shfl_xor_sync(BG_FULL, x, j);
        x = (lane & j) ? ((x & ~m) | ((y & ~m) >> j)) : ((x & m) | ((y & m) << j));
    }
    return x;
}
BG_DEV void transposew(uint32_t (&M)[1]) { M[0] = transpose32(M[0]); }
BG_DEV void transposew(uint64_t (&M)[2]) {
    const uint32_t a = transpose32((uint32_t)M[0]);            // block (rows 0-31, cols 0-31)
    const uint32_t b = transpose32((uint32_t)(M[0] >> 32));    // block (rows 0-31, cols 32-63) -> rows 32-63, cols 0-31
    const uint32_t c = transpose32((uint32_t)M[1]);            // block (rows 32-63, cols 0-31) -> rows 0-31, cols 32-63
    const uint32_t d = transpose32((uint32_t)(M[1] >> 32));
    M[0] = ((uint64_t)c << 32) | a;
    M[1] = ((uint64_t)d << 32) | b;
}

// ---------------------------------------------------------------- quadratic form
// q(x) = Q + sum_v D_v x_v + 4 sum_{u<v} J_uv x_u x_v  (mod 8) on the variables v in A.
// D_v = 2*D1_v + 4*D2_v, J symmetric with J_vv = D1_v (the reference's convention,
// stabilizer.h:12-15).  Bits of J/D outside A are stale and always masked at use.
template <int NS> struct QForm {
    typedef typename WordOf<NS>::T W;
    W J[NS];        // lane-local rows
    W D1, D2, A;    // warp-uniform
    uint32_t Q;     // warp-uniform, mod 8
};

// Change of variables x_i = x'_i + sum_{a in Sp} x'_a  (i not in Sp, Sp subset of A).
// Reproduces updateDJ (stabilizer.c:129-160) for R = I + sum_{a in Sp} E_{a,i}:
//   D'_a = D_a + D_i + 4 J_ai,   J' = R J R^T.           Returns the OLD row i of J.
template <int NS> BG_DEV typename WordOf<NS>::T basis_change(QForm<NS>& f, int i, typename WordOf<NS>::T Sp) {
    typedef typename WordOf<NS>::T W;
    const int lane = bg_lane();
    const W Ji = rowb<NS>(f.J, i);
    const W bi = bitw<W>(i);
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        if ((Sp >> v) & 1) f.J[s] ^= Ji;            // row a += row i
    }
#pragma unroll
    for (int s = 0; s < NS; s++)
        if (f.J[s] & bi) f.J[s] ^= Sp;              // column a += column i
    const W d1i = fillw<W>(getw(f.D1, i)), d2i = fillw<W>(getw(f.D2, i));
    f.D2 ^= Sp & (d2i ^ (d1i & f.D1) ^ Ji);
    f.D1 ^= Sp & d1i;
    return Ji;
}

// Impose the linear constraint  sum_{a in S} x_a = beta  (S != 0, S subset of A): eliminate
// x_i, i = lowest(S).  Same effect on (Q,D,J) as the non-lazy branch of shrink
// (stabilizer.c:522-583): basis change, then updateQD with y = beta e_i (:163-177).
template <int NS> BG_DEV int pivot(QForm<NS>& f, typename WordOf<NS>::T S, uint32_t beta) {
    typedef typename WordOf<NS>::T W;
    const int i = lowestw(S);
    const W bi = bitw<W>(i);
    const W Sp = S ^ bi;
    const uint32_t d1 = getw(f.D1, i), d2 = getw(f.D2, i);
    const W Ji = basis_change<NS>(f, i, Sp);
    if (beta) {
        f.Q = (f.Q + 2u * d1 + 4u * d2) & 7u;
        f.D2 ^= Ji ^ (Sp & fillw<W>(d1));           // column i of J' = J_i + [in Sp] J_ii
    }
    f.A &= ~bi;
    return i;
}

// sum_{x in F_2^A} e^{i pi q(x)/4} = eps 2^{p/2} e^{i pi m/4}
// Result of exponentialSumExact (stabilizer.c:300-481), m reduced mod 8.
// Destroys f.
template <int NS> BG_DEV void expsum(QForm<NS>& f, int& eps, int& p, int& m) {
    typedef typename WordOf<NS>::T W;
    const int lane = bg_lane();
    const W A = f.A;
    const W S = f.D1 & A;                            // D_a in {2,6}
    const bool has_s = S != 0;
    W E = A, Js = 0;
    uint32_t Ds = 0;
    if (has_s) {
        const int s = lowestw(S);
        const W bs = bitw<W>(s), Sp = S ^ bs;
        Ds = 2u + 4u * getw(f.D2, s);
        if (Sp) basis_change<NS>(f, s, Sp);          // fold S onto s (comment on p.12 of the paper)
        E = A & ~bs;
        Js = rowb<NS>(f.J, s) & E;                   // J_as, a in E
    }
    W D2 = f.D2;
    uint32_t cnt = 0, neg0 = 0, neg1 = 0, z0 = 0, z1 = 0;
    while (E) {
        const int a = lowestw(E);
        const W ba = bitw<W>(a);
        const W Ja = rowb<NS>(f.J, a) & E & ~ba;
        const uint32_t d2a = getw(D2, a), sa = getw(Js, a);
        if (Ja == 0) {                               // monomer {a}: 1 + e^{i pi D_a/4}, D_a in {0,4}
            z0 |= d2a; z1 |= d2a ^ sa; cnt++;
            E ^= ba;
            if (z0 && (z1 || !has_s)) break;         // the whole sum is zero
            continue;
        }
        const int b = lowestw(Ja);                   // dimer {a,b}
        const W bb = bitw<W>(b);
        const W Jb = rowb<NS>(f.J, b) & E & ~bb;
        const W rest = E & ~(ba | bb);
        const uint32_t d2b = getw(D2, b), sb = getw(Js, b);
        neg0 ^= d2a & d2b; neg1 ^= (d2a ^ sa) & (d2b ^ sb); cnt++;
        const W Jar = Ja & rest, Jbr = Jb & rest;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int v = lane + 32 * s;
            f.J[s] ^= (fillw<W>(getw(Jar, v)) & Jbr) ^ (fillw<W>(getw(Jbr, v)) & Jar);
        }
        D2 ^= (Jar & fillw<W>(d2b)) ^ (Jbr & fillw<W>(d2a)) ^ (Jar & Jbr);
        Js ^= (Jar & fillw<W>(sb)) ^ (Jbr & fillw<W>(sa));
        E = rest;
    }
    p = 2 * (int)cnt;
    const uint32_t m0 = (f.Q + 4u * neg0) & 7u;
    if (!has_s) { eps = z0 ? 0 : 1; m = (int)m0; return; }
    const uint32_t m1 = (f.Q + Ds + 4u * neg1) & 7u;
    if (z0 && z1) { eps = 0; m = 0; p = 0; return; }
    eps = 1;
    if (z0) { m = (int)m1; return; }
    if (z1) { m = (int)m0; return; }
    const uint32_t diff = (m1 - m0) & 7u;            // 2 or 6: 1 + e^{i pi diff/4} = sqrt2 e^{+- i pi/4}
    p += 1;
    m = (int)((m0 + (diff == 2u ? 1u : 7u)) & 7u);
}

// Pending linear constraints, one per lane slot: sum_{v in Cw_j} x_v = Cbeta_j for j in pend.
// Pivot them one by one; returns false when they are inconsistent (the sum is empty).
template <int NS>
BG_DEV bool apply_constraints(QForm<NS>& f, typename WordOf<NS>::T (&Cw)[NS], typename WordOf<NS>::T pend,
                              typename WordOf<NS>::T Cbeta) {
    typedef typename WordOf<NS>::T W;
    while (pend) {
        const int j = lowestw(pend);
        pend ^= bitw<W>(j);
        const W w = rowb<NS>(Cw, j) & f.A;
        const uint32_t beta = getw(Cbeta, j);
        if (w == 0) { if (beta) return false; continue; }
        const W bi = bitw<W>(lowestw(w));
        bool hit[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) { hit[s] = (Cw[s] & bi) != 0; if (hit[s]) Cw[s] ^= w; }
        if (beta) Cbeta ^= ballotw<NS>(hit);
        pivot<NS>(f, w, beta);
    }
    return true;
}

// ---------------------------------------------------------------- native stabilizer state
// |K,q> with K = h + span{G_a : a in f.A}; (Q,D,J) in f are in G-coordinates.  Gb = (G^-1)^T.
template <int NS> struct Native {
    typedef typename WordOf<NS>::T W;
    W G[NS], Gb[NS];
    QForm<NS> f;
    W h;
    int n;
};

template <int NS> BG_DEV void native_identity(Native<NS>& st, int n) {
    typedef typename WordOf<NS>::T W;
    const int lane = bg_lane();
    st.n = n; st.h = 0;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        st.G[s] = st.Gb[s] = v < n ? bitw<W>(v) : 0;
        st.f.J[s] = 0;
    }
    st.f.D1 = st.f.D2 = 0; st.f.Q = 0; st.f.A = lowmaskw<W>(n);
}

// rows a in A with odd overlap with xi
template <int NS> BG_DEV typename WordOf<NS>::T overlap_rows(const typename WordOf<NS>::T (&M)[NS], typename WordOf<NS>::T xi) {
    bool p[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) p[s] = parw(M[s] & xi) != 0;
    return ballotw<NS>(p);
}
// xor of the rows selected by mask
template <int NS> BG_DEV typename WordOf<NS>::T xor_rows(const typename WordOf<NS>::T (&M)[NS], typename WordOf<NS>::T sel) {
    typedef typename WordOf<NS>::T W;
    const int lane = bg_lane();
    W acc = 0;
#pragma unroll
    for (int s = 0; s < NS; s++) if ((sel >> (lane + 32 * s)) & 1) acc ^= M[s];
    return xorredw(acc);
}

// shrink (stabilizer.c:500-585) in active-mask form: K <- K ∩ {x : xi.x = alpha}.
// Returns 0 EMPTY, 1 SAME, 2 SUCCESS.  lazy: leave (Q,D,J) untouched.
template <int NS> BG_DEV int native_shrink(Native<NS>& st, typename WordOf<NS>::T xi, uint32_t alpha, bool lazy) {
    typedef typename WordOf<NS>::T W;
    const int lane = bg_lane();
    const W S = overlap_rows<NS>(st.G, xi) & st.f.A;
    const uint32_t beta = (alpha ^ parw(xi & st.h)) & 1u;
    if (S == 0) return beta ? 0 : 1;
    const int i = lowestw(S);
    const W bi = bitw<W>(i), Sp = S ^ bi;
    const W Gi = rowb<NS>(st.G, i);
    const W gb = xor_rows<NS>(st.Gb, Sp);            // gbar^i += sum_{a in Sp} gbar^a
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        if ((Sp >> v) & 1) st.G[s] ^= Gi;            // g^a += g^i
        if (v == i) st.Gb[s] ^= gb;
    }
    if (lazy) st.f.A &= ~bi;
    else pivot<NS>(st.f, S, beta);
    if (beta) st.h ^= Gi;
    return 2;
}

// extend (stabilizer.c:759-825): K <- K + span(xi) when xi is not in span(K).  Returns the
// re-activated row index, or -1 when nothing changes.
template <int NS> BG_DEV int native_extend(Native<NS>& st, typename WordOf<NS>::T xi) {
    typedef typename WordOf<NS>::T W;
    const int lane = bg_lane();
    const W S = overlap_rows<NS>(st.Gb, xi) & lowmaskw<W>(st.n);
    const W T = S & ~st.f.A;
    if (T == 0) return -1;
    const int i = lowestw(T);
    const W bi = bitw<W>(i), Sp = S ^ bi;
    const W Gbi = rowb<NS>(st.Gb, i);
    const W g = xor_rows<NS>(st.G, Sp);
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        if ((Sp >> v) & 1) st.Gb[s] ^= Gbi;
        if (v == i) st.G[s] ^= g;
    }
    st.f.A |= bi;
    return i;
}

// measurePauli (stabilizer.c:827-959): project onto the +1 eigenspace of i^m Z(zeta) X(xi).
// Returns 0 (annihilated), 1 (unchanged, factor 1) or 2 (factor 2^-1/2).
template <int NS> BG_DEV int native_measure(Native<NS>& st, uint32_t m, typename WordOf<NS>::T zeta, typename WordOf<NS>::T xi) {
    typedef typename WordOf<NS>::T W;
    const int lane = bg_lane();
    QForm<NS>& f = st.f;
    const W A = f.A;
    const W vecXi = overlap_rows<NS>(st.Gb, xi) & A;
    const W vecZeta = overlap_rows<NS>(st.G, zeta) & A;
    const W xiPrime = xor_rows<NS>(st.G, vecXi);
    // eq. 88: w = 2m + 4 zeta.h + sum_b D_b xi_b + 4 sum_{a<b} J_ab xi_a xi_b
    bool tp[NS], ep[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        const W row = f.J[s] & vecXi;
        tp[s] = ((vecXi >> v) & 1) && parw(row & lowmaskw<W>(v));
        ep[s] = parw(row) != 0;                                         // (J vecXi)_v, eq. 94
    }
    const uint32_t tri = parw(ballotw<NS>(tp));
    const W eta = ((ballotw<NS>(ep) ^ vecZeta) & A);
    const uint32_t w = (2u * m + 4u * parw(zeta & st.h) + 2u * (uint32_t)popcw(f.D1 & vecXi)
                        + 4u * (uint32_t)popcw(f.D2 & vecXi) + 4u * tri) & 7u;
    if (xi == xiPrime) {
        if (w == 0u || w == 4u) {
            const W gamma = xor_rows<NS>(st.Gb, eta);
            const uint32_t alpha = ((w >> 2) ^ parw(gamma & st.h)) & 1u;
            const int r = native_shrink<NS>(st, gamma, alpha, false);
            return r;                                                   // 0, 1, or 2
        }
        // w in {2,6}: eq. 100-101
        if (w == 2u) {           // sigma = +1: Q += 1, D_a -= 2 eta_a
            f.Q = (f.Q + 1u) & 7u;
            f.D2 ^= eta & ~f.D1;
        } else {                 // sigma = -1: Q -= 1, D_a += 2 eta_a
            f.Q = (f.Q + 7u) & 7u;
            f.D2 ^= eta & f.D1;
        }
        f.D1 ^= eta;
#pragma unroll
        for (int s = 0; s < NS; s++) if ((eta >> (lane + 32 * s)) & 1) f.J[s] ^= eta;
        return 2;
    }
    // xi not in span(K): extend, then set the new row of (D,J)  (stabilizer.c:941-951)
    const int i = native_extend<NS>(st, xi);
    const W bi = bitw<W>(i);
    const uint32_t newD = (2u * m + 4u * parw(zeta & xi) + 4u * parw(zeta & st.h)) & 7u;
    f.D1 = (f.D1 & ~bi) | (fillw<W>((newD >> 1) & 1u) & bi);
    f.D2 = (f.D2 & ~bi) | (fillw<W>((newD >> 2) & 1u) & bi);
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        if (v == i) f.J[s] = vecZeta | (fillw<W>(m & 1u) & bi);
        else f.J[s] = (f.J[s] & ~bi) | (fillw<W>(getw(vecZeta, v)) & bi);
    }
    return 2;
}

// ---------------------------------------------------------------- ambient form
// Re-express the state on the computational-basis bits x in F_2^n:
//   for x in K:  q(x) = Q^ + sum_q D^_q x_q + 4 sum_{q<r} J^_qr x_q x_r,
//   K = { x : Cw_b . x = Cbeta_b for the rows b outside A }.
// (updateDJ with R = Gbar[A]^T, stabilizer.c:129-160, then the shift x -> x + h as in
//  updateQD, :163-177.)  Output: form `o` on A = all n bits, constraints in lane slots.
template <int NS>
BG_DEV void ambient(const Native<NS>& st, QForm<NS>& o, typename WordOf<NS>::T (&Cw)[NS],
                    typename WordOf<NS>::T& Cpend, typename WordOf<NS>::T& Cbeta) {
    typedef typename WordOf<NS>::T W;
    const int lane = bg_lane();
    const int n = st.n;
    const W A = st.f.A, maskn = lowmaskw<W>(n);
    // Are the active rows of Gbar unit vectors (Gbar_a = e_a for a in A)?  Always so for a state drawn by
    // native_random — a lazy shrink only touches the Gbar row of the pivot it removes from A — and then the
    // G-coordinates of x are its own bits: R is the identity on A and (D~, J~) = (D, J) on A.
    W Dt1, Dt2, Jt[NS];
    bool un[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        un[s] = ((A >> v) & 1) && st.Gb[s] != bitw<W>(v);
    }
    if (ballotw<NS>(un) == 0) {
        Dt1 = st.f.D1 & A; Dt2 = st.f.D2 & A;
#pragma unroll
        for (int s = 0; s < NS; s++) Jt[s] = ((A >> (lane + 32 * s)) & 1) ? (st.f.J[s] & A) : (W)0;
    } else {
        // R_q = column q of Gbar restricted to the rows in A
        W R[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) R[s] = st.Gb[s];
        transposew(R);
#pragma unroll
        for (int s = 0; s < NS; s++) R[s] &= A;
        // M_q = xor_{a in R_q} J_a ;  t_q = sum_{b<a in R_q} J_ab  (strictly lower rows broadcast separately,
        // so no per-iteration mask arithmetic; rows outside A are all-zero in R and are skipped)
        W M[NS], Jlow[NS]; uint32_t tq[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) { M[s] = 0; tq[s] = 0; Jlow[s] = st.f.J[s] & A & lowmaskw<W>(lane + 32 * s); }
        for (int a = 0; a < n; a++) {
            if (!((A >> a) & 1)) continue;
            const W Ja = rowb<NS>(st.f.J, a) & A;
            const W Jl = rowb<NS>(Jlow, a);
#pragma unroll
            for (int s = 0; s < NS; s++)
                if ((R[s] >> a) & 1) { M[s] ^= Ja; tq[s] ^= parw(Jl & R[s]); }
        }
        bool p1[NS], p2[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const uint32_t c1 = (uint32_t)popcw(R[s] & st.f.D1);
            p1[s] = (c1 & 1u) != 0;
            p2[s] = (((c1 >> 1) ^ (uint32_t)popcw(R[s] & st.f.D2) ^ tq[s]) & 1u) != 0;
        }
        Dt1 = ballotw<NS>(p1) & maskn;
        Dt2 = ballotw<NS>(p2) & maskn;
        // J~ = M R^T and R^T = Gbar[A]:  J~_q = xor_{a in M_q} Gbar_a  (M_q only has bits in A)
#pragma unroll
        for (int s = 0; s < NS; s++) Jt[s] = 0;
        for (int a = 0; a < n; a++) {
            if (!((A >> a) & 1)) continue;
            const W Ga = rowb<NS>(st.Gb, a);
#pragma unroll
            for (int s = 0; s < NS; s++) if ((M[s] >> a) & 1) Jt[s] ^= Ga;
        }
#pragma unroll
        for (int s = 0; s < NS; s++) Jt[s] &= maskn;
    }
    // shift u = x + h
    const W h = st.h;
    bool pq[NS], pd[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        const W row = Jt[s] & h;
        pq[s] = ((h >> v) & 1) && parw(row & lowmaskw<W>(v));
        pd[s] = parw(row & ~bitw<W>(v)) != 0;
    }
    const uint32_t tri = parw(ballotw<NS>(pq));
    o.Q = (st.f.Q + 2u * (uint32_t)popcw(Dt1 & h) + 4u * (uint32_t)popcw(Dt2 & h) + 4u * tri) & 7u;
    Dt2 ^= (h & Dt1) ^ (ballotw<NS>(pd) & maskn);
    o.D1 = Dt1; o.D2 = Dt2; o.A = maskn;
#pragma unroll
    for (int s = 0; s < NS; s++) o.J[s] = Jt[s];
    // parity checks: rows of Gbar outside A
    Cpend = maskn & ~A;
    bool pb[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        const bool mine = ((Cpend >> v) & 1) != 0;
        Cw[s] = mine ? st.Gb[s] : 0;
        pb[s] = mine && parw(st.Gb[s] & h);
    }
    Cbeta = ballotw<NS>(pb);
}

// ---------------------------------------------------------------- projection in ambient form
// measurePauli (stabilizer.c:827-959) applied to a state that is ALREADY in ambient form:
//   theta(x) ~ w^{q(x)} for x in K = { x : Cw_b . x = Cbeta_b, b in Cpend },  w = e^{i pi/4},
// q = (Q, D, J) on all n bits.  With G = I every quantity of eq. 88-101 is a mask operation on the
// computational-basis bits: no G / Gbar, no shrink / extend, no conversion afterwards.
//
// For P = i^m Z(zeta) X(xi):  <y|P|theta> = i^m (-1)^{zeta.y} theta(y + xi), and as polynomials
//   q(y + xi) - q(y) = l0 + 4 (J xi).y,   l0 = D.xi + 4 sum_{q<r} J_qr xi_q xi_r   (J_qq = D1_q),
// so with  w0 = 2m + l0 (mod 8, always even)  and  eta = zeta + J xi:
//   xi outside the directions of K (some check has c.xi = 1): the support doubles, K' = K u (K + xi); one
//     such check c0 is dropped (the others get += c0) and, with u(y) = c0.y + beta0 its indicator,
//     q' = q + u(y) (w0 + 4 eta.y);                                   factor 2^-1/2   (the "extend" case)
//   xi inside, w0 in {0,4}: theta' = theta restricted to eta.y = w0/4: a new check, or — when eta is in
//     the span of the checks — nothing (factor 1) or annihilation;      factor 2^-1/2  (the "shrink" case)
//   xi inside, w0 in {2,6}: theta' = 2^-1/2 theta w^{+-1 -+ 2 (eta.y mod 2)}: Q +- 1, D_q -+ 2 and
//     J_qr ^= 1 on eta.                                                 factor 2^-1/2
// (n (eta.y mod 2) = n sum_q - 2n e2 + ...: for n = 2, 6 that is  n sum_{q in eta} y_q + 4 sum_{q<r in eta} y_q y_r.)
//
// Invariant of the checks: reduced echelon form by slot — the check stored in lane slot v contains bit v and
// no other check does (checks_echelon establishes it once).  Returns 0 (annihilated), 1 or 2 (2^-1/2).
template <int NS>
BG_DEV void checks_echelon(typename WordOf<NS>::T (&Cw)[NS], typename WordOf<NS>::T& Cpend, typename WordOf<NS>::T& Cbeta) {
    typedef typename WordOf<NS>::T W;
    const int lane = bg_lane();
    W Cn[NS], pn = 0, bn = 0;
#pragma unroll
    for (int s = 0; s < NS; s++) Cn[s] = 0;
    for (W rem = Cpend; rem;) {
        const int j = lowestw(rem);
        rem ^= bitw<W>(j);
        W w = rowb<NS>(Cw, j);
        uint32_t beta = getw(Cbeta, j);
        const W flags = w & pn;                                   // pivots of the checks already placed
        if (flags) { w ^= xor_rows<NS>(Cn, flags); beta ^= parw(flags & bn); }
        if (w == 0) continue;                                     // (dependent: cannot happen for rows of Gbar)
        const int p = lowestw(w);
        bool h[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int v = lane + 32 * s;
            h[s] = ((pn >> v) & 1) && ((Cn[s] >> p) & 1);
            if (h[s]) Cn[s] ^= w;
            if (v == p) Cn[s] = w;
        }
        if (beta) bn ^= ballotw<NS>(h);
        pn |= bitw<W>(p);
        bn = (bn & ~bitw<W>(p)) | ((W)beta << p);
    }
#pragma unroll
    for (int s = 0; s < NS; s++) Cw[s] = Cn[s];
    Cpend = pn; Cbeta = bn;
}

template <int NS>
BG_DEV int ambient_measure(QForm<NS>& f, typename WordOf<NS>::T (&Cw)[NS], typename WordOf<NS>::T& Cpend,
                           typename WordOf<NS>::T& Cbeta, uint32_t m, typename WordOf<NS>::T zeta, typename WordOf<NS>::T xi) {
    typedef typename WordOf<NS>::T W;
    const int lane = bg_lane();
    bool ep[NS], tp[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        const W row = f.J[s] & xi;
        ep[s] = parw(row) != 0;                                         // (J xi)_v, diagonal J_vv = D1_v included
        tp[s] = ((xi >> v) & 1) && parw(row & lowmaskw<W>(v));          // sum_{q<r} J_qr xi_q xi_r
    }
    const W eta = (zeta ^ ballotw<NS>(ep)) & f.A;
    const uint32_t tri = parw(ballotw<NS>(tp));
    W hit = 0;                                                          // checks with c.xi = 1
    if (Cpend) {
        bool cp[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) cp[s] = ((Cpend >> (lane + 32 * s)) & 1) && parw(Cw[s] & xi);
        hit = ballotw<NS>(cp);
    }
    const uint32_t w0 = (2u * m + 2u * (uint32_t)popcw(f.D1 & xi) + 4u * (uint32_t)popcw(f.D2 & xi) + 4u * tri) & 7u;
    if (hit) {
        // ---- xi leaves K: K' = K u (K + xi)
        const int p0 = lowestw(hit);
        const W bp0 = bitw<W>(p0);
        const W c0 = rowb<NS>(Cw, p0);
        const uint32_t b0 = getw(Cbeta, p0);
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int v = lane + 32 * s;
            if (((hit >> v) & 1) && v != p0) Cw[s] ^= c0;               // the other checks become xi-invariant
            if (v == p0) Cw[s] = 0;
            if ((c0 >> v) & 1) f.J[s] ^= eta;                           // 4 u(y) (eta.y): J_qr ^= c0_q eta_r + c0_r eta_q
            if ((eta >> v) & 1) f.J[s] ^= c0;
        }
        if (b0) Cbeta ^= hit;
        Cbeta &= ~bp0; Cpend &= ~bp0;
        f.D2 ^= (c0 & eta) ^ (b0 ? eta : (W)0);
        // w0 u(y), u = beta0 + (1 - 2 beta0) (c0.y mod 2)
        if (b0) f.Q = (f.Q + w0) & 7u;
        const uint32_t wp = b0 ? ((8u - w0) & 7u) : w0;
        if (wp == 4u) f.D2 ^= c0;
        else if (wp == 2u || wp == 6u) {
            f.D2 ^= (wp == 2u ? f.D1 : ~f.D1) & c0;                     // D_q += 2 / D_q -= 2 on c0
            f.D1 ^= c0;
#pragma unroll
            for (int s = 0; s < NS; s++) if ((c0 >> (lane + 32 * s)) & 1) f.J[s] ^= c0;
        }
        return 2;
    }
    if (w0 == 0u || w0 == 4u) {
        // ---- projector onto eta.y = w0/4 inside K
        const W flags = eta & Cpend;
        W er = eta;
        uint32_t br = w0 >> 2;
        if (flags) { er ^= xor_rows<NS>(Cw, flags); br ^= parw(flags & Cbeta); }
        if (er == 0) return br ? 0 : 1;
        const int p = lowestw(er);
        bool h[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int v = lane + 32 * s;
            h[s] = ((Cpend >> v) & 1) && ((Cw[s] >> p) & 1);
            if (h[s]) Cw[s] ^= er;
            if (v == p) Cw[s] = er;
        }
        if (br) Cbeta ^= ballotw<NS>(h);
        Cpend |= bitw<W>(p);
        Cbeta = (Cbeta & ~bitw<W>(p)) | ((W)br << p);
        return 2;
    }
    // ---- w0 in {2,6}: 1 + w^{w(y)} = sqrt2 w^{+-1}
    if (w0 == 2u) { f.Q = (f.Q + 1u) & 7u; f.D2 ^= eta & ~f.D1; }
    else { f.Q = (f.Q + 7u) & 7u; f.D2 ^= eta & f.D1; }
    f.D1 ^= eta;
#pragma unroll
    for (int s = 0; s < NS; s++) if ((eta >> (lane + 32 * s)) & 1) f.J[s] ^= eta;
    return 2;
}

// ---------------------------------------------------------------- decomposition terms
// <phi_i|theta> for phi_i = prepL(i) (stateprep.c:85-120): a product of |+> on supp(xt) and |0>
// elsewhere.  `base` is theta's ambient form (kept intact), k1 = dim K_theta.
// (eps, p, m) as innerProductExact(theta, phi_i) returns them (stabilizer.c:589-659).
template <int NS>
BG_DEV void term_L(const QForm<NS>& base, const typename WordOf<NS>::T (&Cw0)[NS], typename WordOf<NS>::T Cpend,
                   typename WordOf<NS>::T Cbeta, int k1, typename WordOf<NS>::T xt, int& eps, int& p, int& m) {
    typedef typename WordOf<NS>::T W;
    QForm<NS> f = base;
    f.A = base.A & xt;
    W Cw[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) Cw[s] = Cw0[s];
    if (!apply_constraints<NS>(f, Cw, Cpend, Cbeta)) { eps = 0; p = 0; m = 0; return; }
    expsum<NS>(f, eps, p, m);
    if (eps) p -= k1 + popcw(base.A & xt);          // k2 = |x~|
    else { p = 0; m = 0; }
}

// <phi_i|theta> for phi_i = prepH(i) (stateprep.c:36-81).  e1 has bit 2j set when pair j is
// |00>+|11> (and bit t-1 set when t is odd and the last qubit is |0>); unset pairs are
// |00>+|01>+|10>-|11>.
template <int NS>
BG_DEV void term_H(const QForm<NS>& base, const typename WordOf<NS>::T (&Cw0)[NS], typename WordOf<NS>::T Cpend,
                   typename WordOf<NS>::T Cbeta, int k1, int t, typename WordOf<NS>::T e1, int& eps, int& p, int& m) {
    typedef typename WordOf<NS>::T W;
    const int lane = bg_lane();
    const W maskt = lowmaskw<W>(t);
    const W even = (W)0x5555555555555555ull;
    const W pairs = even & (maskt >> 1);                 // 2j with 2j+1 < t
    const W mrg = e1 & pairs;                            // x_{2j} = x_{2j+1}
    const W cz = ~e1 & pairs;                            // phase (-1)^{x_2j x_2j+1}
    const W last = (t & 1) ? (e1 & bitw<W>(t - 1)) : 0;  // x_{t-1} = 0
    QForm<NS> f = base;
    W Cw[NS];
    const W cz2 = cz | (cz << 1);
    bool pj[NS];
    W up[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        if ((cz2 >> v) & 1) f.J[s] ^= bitw<W>(v ^ 1);    // q1 - q2
        pj[s] = ((mrg >> v) & 1) && ((f.J[s] >> (v + 1)) & 1);          // J_{2j,2j+1} of merged pairs
        up[s] = shflxw(f.J[s], 1);
    }
    const W Jpair = ballotw<NS>(pj);
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        if ((v & 1) && ((mrg >> (v - 1)) & 1)) f.J[s] ^= up[s];         // row 2j+1 += row 2j
        f.J[s] ^= (f.J[s] & mrg) << 1;                                   // col 2j+1 += col 2j
        Cw[s] = Cw0[s] ^ ((Cw0[s] & mrg) << 1);
    }
    const W c1 = (f.D1 & mrg) << 1;
    f.D2 ^= ((f.D2 & mrg) << 1) ^ (c1 & f.D1) ^ (Jpair << 1);
    f.D1 ^= c1;
    f.A = base.A & maskt & ~mrg & ~last;
    if (!apply_constraints<NS>(f, Cw, Cpend, Cbeta)) { eps = 0; p = 0; m = 0; return; }
    expsum<NS>(f, eps, p, m);
    if (eps) p -= k1 + (t - popcw(mrg) - popcw(last));
    else { p = 0; m = 0; }
}

// ---------------------------------------------------------------- exact accumulation
// sum of eps 2^{p/2} w^m (w = e^{i pi/4}) as four integers (coefficients of 1, w, w^2, w^3)
// in units of 2^-sh, sh = t/2 + 1:  p >= -t for a non-zero overlap of normalised states.
struct Zw { long long a[4]; };
// z += sgn * mag * w^e  without dynamic indexing (keeps the accumulator in registers)
BG_DEV void zw_add_pow(Zw& z, int e, long long mag) {
    const long long v = (e & 4) ? -mag : mag;
    const int j = e & 3;
    z.a[0] += (j == 0) ? v : 0;
    z.a[1] += (j == 1) ? v : 0;
    z.a[2] += (j == 2) ? v : 0;
    z.a[3] += (j == 3) ? v : 0;
}
BG_DEV void zw_add(Zw& z, int eps, int p, int m, int sh) {
    // no branches: the 32 lanes of a warp accumulate together.  eps = 0 adds 0; an odd p adds
    // sqrt2 w^m = w^{m+1} + w^{m-1}, an even p adds w^m (second term 0).
    const int f = p >> 1;                                // floor(p/2), also for p < 0
    const long long mag = eps ? (1ll << ((sh + f) & 63)) : 0ll;
    const int odd = p & 1, mm = m & 7;
    zw_add_pow(z, (mm + odd) & 7, mag);
    zw_add_pow(z, (mm + 7) & 7, odd ? mag : 0ll);
}

// The same in 32-bit integers, for a thread that adds a bounded number of terms before it hands its sums to the 64-bit
// accumulator: |<phi|theta>| <= 1 for normalised states, so p <= 0 and a term adds at most 2^sh = 2^(t/2+1) to a
// component; with t <= 44 that is 2^23, and 128 terms stay below 2^31.  (A third of the instructions of zw_add.)
struct Zw32 { int a[4]; };
BG_DEV void zw32_add_pow(Zw32& z, int e, int mag) {
    const int v = (e & 4) ? -mag : mag;
    const int j = e & 3;
    z.a[0] += (j == 0) ? v : 0;
    z.a[1] += (j == 1) ? v : 0;
    z.a[2] += (j == 2) ? v : 0;
    z.a[3] += (j == 3) ? v : 0;
}
BG_DEV void zw32_add(Zw32& z, int eps, int p, int m, int sh) {
    const int ex = sh + (p >> 1);
    const int mag = eps ? (1 << (ex < 0 ? 0 : (ex > 30 ? 30 : ex))) : 0;
    const int odd = p & 1, mm = m & 7;
    zw32_add_pow(z, (mm + odd) & 7, mag);
    zw32_add_pow(z, (mm + 7) & 7, odd ? mag : 0);
}
#define ZW32_MAX_T 44            // 2^(t/2+1) <= 2^23
#define ZW32_MAX_TERMS 128       // terms a thread may add between two flushes

}  // namespace bg

namespace bg {
// ---------------------------------------------------------------- global-memory helpers
// Load a bg_state (contiguous layout: active rows are 0..k-1).  Junk that the reference
// leaves outside the k x k block / beyond n (stabilizer.c:743-745, matrix.c:85-92) is masked.
template <int NS> BG_DEV void native_load(Native<NS>& st, const bg_state* g) {
    typedef typename WordOf<NS>::T W;
    const int lane = bg_lane();
    const int n = g->n, k = g->k;
    const W maskn = lowmaskw<W>(n), maskk = lowmaskw<W>(k);
    st.n = n;
    st.h = (W)g->h & maskn;
    st.f.A = maskk; st.f.Q = (uint32_t)g->Q & 7u;
    st.f.D1 = (W)g->D1 & maskk; st.f.D2 = (W)g->D2 & maskk;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        st.G[s] = v < n ? ((W)g->G[v] & maskn) : 0;
        st.Gb[s] = v < n ? ((W)g->Gbar[v] & maskn) : 0;
        st.f.J[s] = v < k ? ((W)g->J[v] & maskk) : 0;
    }
}
// Store in active-mask layout (rows stay where they are); *Aout receives the active mask.
// The host compacts to the contiguous layout (bg_compact_state in bgnorm.cu).
template <int NS> BG_DEV void native_store_raw(const Native<NS>& st, bg_state* g, uint64_t* Aout) {
    typedef typename WordOf<NS>::T W;
    const int lane = bg_lane();
    const W A = st.f.A;
    if (lane == 0) {
        g->n = st.n; g->k = popcw(A); g->Q = (int32_t)(st.f.Q & 7u); g->reserved = 0;
        g->h = (uint64_t)st.h; g->D1 = (uint64_t)(st.f.D1 & A); g->D2 = (uint64_t)(st.f.D2 & A);
        *Aout = (uint64_t)A;
    }
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        g->G[v] = (uint64_t)st.G[s];
        g->Gbar[v] = (uint64_t)st.Gb[s];
        g->J[v] = ((A >> v) & 1) ? (uint64_t)(st.f.J[s] & A) : 0;
    }
    if (NS == 1) { g->G[lane + 32] = 0; g->Gbar[lane + 32] = 0; g->J[lane + 32] = 0; }
}

}  // namespace bg
