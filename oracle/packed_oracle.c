/* packed_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement, on packed uint64 rows, of the reference's algorithm for
 * the stabilizer-rank norm-estimation path (Bravyi-Gosset arXiv:1601.07601 as
 * implemented by patrickrall/CircuitSimulator).  It follows the reference's
 * pivot choices, row swaps and update order step for step so that it reproduces
 * the reference's full intermediate states (not only the final amplitudes) and
 * can be pinned against the reference's own known-answer files
 * (tests/units/tests-c/*.txt) and against the compiled reference
 * (oracle/_ref/libcircref.so).  Parity status: PINNED — see tests/test_oracle_*.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this library.  The product (circuitsimulator_b200/csrc) never does.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * the reference root).  Layout: include/bgnorm.h (bit q of a word = index q).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include "bgnorm.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef bg_state S;

static inline int par64(uint64_t x) { return __builtin_parityll(x); }
static inline int pop64(uint64_t x) { return __builtin_popcountll(x); }
static inline uint64_t lowmask(int n) { return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }
static inline int bit(uint64_t w, int i) { return (int)((w >> i) & 1ull); }

/* word-op counter (instrumentation for DESIGN.md's algorithmic-work figure) */
static uint64_t g_wordops = 0;
uint64_t orc_wordops(void) { return g_wordops; }
void orc_wordops_reset(void) { g_wordops = 0; }

/* getD / setD — stabilizer.c:55-62.  D_a = 2*D1_a + 4*D2_a. */
static inline int getD(const S* s, int a) { return 2 * bit(s->D1, a) + 4 * bit(s->D2, a); }
static inline void setD(S* s, int a, int val) {
    /* the reference stores bits 1 and 2 of val (val >= 0) */
    uint64_t m = 1ull << a;
    s->D1 = (s->D1 & ~m) | ((uint64_t)((val / 2) % 2) << a);
    s->D2 = (s->D2 & ~m) | ((uint64_t)(((val / 2 - (val / 2) % 2) / 2) % 2) << a);
}

/* allocStabilizerState — stabilizer.c:9-25: K = F_2^n slice of dimension k, G = Gbar = I, q = 0 */
void orc_identity_state(S* s, int n, int k) {
    memset(s, 0, sizeof(*s));
    s->n = n; s->k = k;
    for (int i = 0; i < n; i++) { s->G[i] = 1ull << i; s->Gbar[i] = 1ull << i; }
}

/* updateDJ — stabilizer.c:129-160.  R is n x n (row i = R[i]).
 *   D'_i = sum_{j<k} D_j R_ij (mod 8)                           (eq. 49)
 *   D2_a ^= parity( sum_{c<b} J_bc R_ab R_ac )  with the OLD J, all b,c < n
 *   J  <- R J R^T over the full n x n block                     (eq. 50) */
static void update_dj(S* s, const uint64_t* R) {
    const int n = s->n, k = s->k;
    const uint64_t mk = lowmask(k);
    uint64_t nD1 = 0, nD2 = 0;
    for (int i = 0; i < n; i++) {
        int v = 2 * pop64(R[i] & s->D1 & mk) + 4 * pop64(R[i] & s->D2 & mk);
        nD1 |= (uint64_t)((v >> 1) & 1) << i;
        nD2 |= (uint64_t)((v >> 2) & 1) << i;
    }
    s->D1 = nD1; s->D2 = nD2;
    g_wordops += 6ull * n;

    uint64_t M[BG_MAX_T];
    for (int a = 0; a < n; a++) {
        uint64_t r = R[a], acc = 0;
        int val = 0;
        while (r) {
            int b = __builtin_ctzll(r); r &= r - 1;
            val += pop64(s->J[b] & R[a] & lowmask(b));
            acc ^= s->J[b];
            g_wordops += 4;
        }
        if (val & 1) s->D2 ^= 1ull << a;
        M[a] = acc;                      /* (R J)_a */
    }
    /* (M R^T)_{a,c} = parity(M_a & R_c); columns c whose R row is e_c keep M_a[c] */
    uint64_t idcols = 0;
    for (int c = 0; c < n; c++) if (R[c] == (1ull << c)) idcols |= 1ull << c;
    for (int a = 0; a < n; a++) {
        uint64_t row = M[a] & idcols;
        uint64_t rest = lowmask(n) & ~idcols;
        while (rest) {
            int c = __builtin_ctzll(rest); rest &= rest - 1;
            row |= (uint64_t)par64(M[a] & R[c]) << c;
            g_wordops += 3;
        }
        s->J[a] = row;
        g_wordops += 1;
    }
}

/* updateQD — stabilizer.c:163-177.
 *   Q += sum_{a<k} D_a y_a + 4 sum_{a<b<k} J_ab y_a y_b  (mod 8)   (eq. 52)
 *   D2 ^= J y over all n rows                                      (eq. 53) */
static void update_qd(S* s, uint64_t y) {
    const int n = s->n, k = s->k;
    const uint64_t mk = lowmask(k);
    int Q = s->Q;
    uint64_t yy = y & mk;
    Q += 2 * pop64(s->D1 & yy) + 4 * pop64(s->D2 & yy);
    uint64_t r = yy;
    while (r) {
        int a = __builtin_ctzll(r); r &= r - 1;
        /* b in (a, k) with y_b */
        Q += 4 * pop64(s->J[a] & yy & ~lowmask(a + 1));
        g_wordops += 3;
    }
    s->Q = Q % 8;
    uint64_t jy = 0;
    for (int a = 0; a < n; a++) jy |= (uint64_t)par64(s->J[a] & y) << a;
    s->D2 ^= jy;
    g_wordops += 2ull * n + 6;
}

/* shrink — stabilizer.c:500-585.  Returns 0 EMPTY, 1 SAME, 2 SUCCESS. */
int orc_shrink(S* s, uint64_t xi, int alpha, int lazy) {
    const int n = s->n;
    int Sidx[BG_MAX_T], Slen = 0;
    for (int a = 0; a < s->k; a++) if (par64(s->G[a] & xi)) Sidx[Slen++] = a;
    g_wordops += 2ull * s->k;
    int beta = (alpha + pop64(xi & s->h)) % 2;
    if (Slen == 0) return 1 - beta;

    int i = Sidx[--Slen];
    uint64_t R[BG_MAX_T];
    for (int t = 0; t < Slen; t++) {
        int a = Sidx[t];
        s->G[a] ^= s->G[i];
        if (lazy != 1) {
            for (int r = 0; r < n; r++) R[r] = 1ull << r;
            R[a] |= 1ull << i;                 /* BitMatrixSet(R, a, i, 1) */
            update_dj(s, R);
        }
        s->Gbar[i] ^= s->Gbar[a];
        g_wordops += 2;
    }
    /* swap rows i <-> k-1 of G and Gbar */
    int km1 = s->k - 1;
    uint64_t tmp;
    tmp = s->G[i]; s->G[i] = s->G[km1]; s->G[km1] = tmp;
    tmp = s->Gbar[i]; s->Gbar[i] = s->Gbar[km1]; s->Gbar[km1] = tmp;
    if (lazy != 1) {
        for (int r = 0; r < n; r++) R[r] = 1ull << r;
        tmp = R[i]; R[i] = R[km1]; R[km1] = tmp;
        update_dj(s, R);
    }
    if (beta == 1) s->h ^= s->G[km1];
    if (lazy != 1) update_qd(s, (uint64_t)beta << km1);
    s->k--;
    return 2;
}

/* Gamma — stabilizer.c:183-225: 1 + i^{A/2} + i^{B/2} - i^{(A+B)/2} for even A,B */
static void Gamma(int* eps, int* p, int* m, int A, int B) {
    static const int cre[4] = {1, 0, -1, 0}, cim[4] = {0, 1, 0, -1};
    int re = 1 + cre[(A % 8) / 2] + cre[(B % 8) / 2] - cre[((A + B) % 8) / 2];
    int im = 0 + cim[(A % 8) / 2] + cim[(B % 8) / 2] - cim[((A + B) % 8) / 2];
    if (re == 0 && im == 0) { *eps = 0; *p = 0; *m = 0; return; }
    *eps = 1; *p = 2;
    if (re == 0) *m = (im > 0) ? 2 : 6;
    else if (re / 2 == -1) *m = 4;
    else {
        if (im / 2 == 1) *m = 1;
        if (im / 2 == 0) *m = 0;
        if (im / 2 == -1) *m = 7;
    }
}

/* partialGamma — stabilizer.c:228-255: 1 + e^{i pi A/4}, A even */
static void partialGamma(int* eps, int* p, int* m, int A) {
    while (A < 0) A += 8;
    switch (A % 8) {
        case 0: *eps = 1; *p = 2; *m = 0; break;
        case 2: *eps = 1; *p = 1; *m = 1; break;
        case 4: *eps = 0; *p = 0; *m = 0; break;
        case 6: *eps = 1; *p = 1; *m = 7; break;
        default: fprintf(stderr, "oracle: partialGamma odd argument\n"); abort();
    }
}

/* Wsigma — stabilizer.c:258-297 */
static void Wsigma(const S* s, int* eps, int* p, int* m, int sigma, int sidx,
                   const int* M, int Mlen, const int* Dimers, int Dlen) {
    if (s->k == 0) { *eps = 1; *p = 0; *m = s->Q; return; }
    int tP = 0, tM = s->Q + sigma * getD(s, sidx);
    for (int i = 0; i < Mlen; i++) {
        partialGamma(eps, p, m, getD(s, M[i]) + sigma * 4 * bit(s->J[M[i]], sidx));
        if (*eps == 0) { *p = 0; *m = 0; return; }
        tP += *p; tM = (tM + *m) % 8;
    }
    for (int i = 0; i < Dlen; i++) {
        int a = Dimers[2 * i], b = Dimers[2 * i + 1];
        Gamma(eps, p, m, getD(s, a) + sigma * 4 * bit(s->J[a], sidx),
                         getD(s, b) + sigma * 4 * bit(s->J[b], sidx));
        if (*eps == 0) { *p = 0; *m = 0; return; }
        tP += *p; tM = (tM + *m) % 8;
    }
    *eps = 1; *p = tP; *m = tM;
}

/* exponentialSumExact — stabilizer.c:300-481.  Works IN PLACE on s (as the
 * reference does).  (eps,p,m): sum_{x in F_2^k} e^{i pi q(x)/4} = eps 2^{p/2} e^{i pi m/4}.
 * m is returned as the reference leaves it (possibly unreduced, see :477). */
void orc_exponential_sum_inplace(S* s, int* eps, int* p, int* m) {
    const int n = s->n, k = s->k;
    int Sidx[BG_MAX_T], Slen = 0;
    for (int a = 0; a < k; a++) if (bit(s->D1, a)) Sidx[Slen++] = a;   /* D_a in {2,6} */

    uint64_t R[BG_MAX_T];
    if (Slen > 0) {
        int a = Sidx[0];
        for (int r = 0; r < n; r++) R[r] = 1ull << r;
        for (int i = 1; i < Slen; i++) R[Sidx[i]] ^= 1ull << a;
        update_dj(s, R);
        /* identity with columns a and k-1 swapped (:329-332) */
        for (int r = 0; r < n; r++) R[r] = 1ull << r;
        if (a != k - 1) { R[a] = 1ull << (k - 1); R[k - 1] = 1ull << a; }
        update_dj(s, R);
        Sidx[0] = k - 1; Slen = 1;
    }

    int E[BG_MAX_T], Elen = 0;
    for (int c = 0; c < k; c++) if (Slen == 0 || c != Sidx[0]) E[Elen++] = c;

    int M[BG_MAX_T], Mlen = 0, Dimers[2 * BG_MAX_T], Dlen = 0;
    while (Elen > 0) {
        int a = E[0], b = -1;
        for (int i = 1; i < Elen; i++) if (bit(s->J[a], E[i])) { b = E[i]; break; }
        if (b < 0) {
            M[Mlen++] = a;
            for (int i = 0; i + 1 < Elen; i++) E[i] = E[i + 1];
            Elen--;
        } else {
            for (int r = 0; r < n; r++) R[r] = 1ull << r;
            for (int i = 0; i < Elen; i++) {
                int c = E[i];
                if (c != a && c != b) {
                    if (bit(s->J[a], c)) R[c] ^= 1ull << b;
                    if (bit(s->J[b], c)) R[c] ^= 1ull << a;
                }
            }
            update_dj(s, R);
            Dimers[2 * Dlen] = a; Dimers[2 * Dlen + 1] = b; Dlen++;
            int w = 0;
            for (int i = 0; i < Elen; i++) if (E[i] != a && E[i] != b) E[w++] = E[i];
            Elen = w;
        }
    }

    if (Slen == 0) { Wsigma(s, eps, p, m, 0, 0, M, Mlen, Dimers, Dlen); return; }
    int e0, p0, m0, e1, p1, m1;
    Wsigma(s, &e0, &p0, &m0, 0, Sidx[0], M, Mlen, Dimers, Dlen);
    Wsigma(s, &e1, &p1, &m1, 1, Sidx[0], M, Mlen, Dimers, Dlen);
    if (e0 == 0) { *eps = e1; *p = p1; *m = m1; return; }
    if (e1 == 0) { *eps = e0; *p = p0; *m = m0; return; }
    if (p0 != p1) { fprintf(stderr, "oracle: ExponentialSum p0 != p1\n"); abort(); }
    if ((m1 - m0) % 2 != 0) { fprintf(stderr, "oracle: ExponentialSum m1-m0 odd\n"); abort(); }
    partialGamma(eps, p, m, m1 - m0);
    if (*eps == 0) { *p = 0; *m = 0; }
    else { *p += p0; *m = *m + m0 % 8; }       /* sic: m + (m0 % 8), stabilizer.c:477 */
}

void orc_exponential_sum(const S* s, int* eps, int* p, int* m) {
    S tmp = *s;
    orc_exponential_sum_inplace(&tmp, eps, p, m);
}

/* innerProductExact — stabilizer.c:589-659 */
void orc_inner_product(const S* s1, const S* s2, int* eps, int* p, int* m) {
    const int n = s1->n;
    S st = *s1;
    for (int b = s2->k; b < n; b++) {
        uint64_t xi = s2->Gbar[b];
        int alpha = pop64(s2->h & xi) % 2;
        *eps = orc_shrink(&st, xi, alpha, 0);
        if (*eps == 0) { *p = 0; *m = 0; return; }
    }
    uint64_t R[BG_MAX_T]; memset(R, 0, sizeof(R));
    uint64_t hh = st.h ^ s2->h, y = 0;
    for (int a = 0; a < n; a++) {
        y |= (uint64_t)par64(hh & s2->Gbar[a]) << a;
        for (int b = 0; b < n; b++) R[b] |= (uint64_t)par64(st.G[b] & s2->Gbar[a]) << a;
    }
    g_wordops += 2ull * n * n + 2ull * n;
    S t2 = *s2;
    update_qd(&t2, y);
    update_dj(&t2, R);

    st.Q -= t2.Q; if (st.Q < 0) st.Q += 8;
    for (int i = 0; i < n; i++) {
        int val = getD(&st, i) - getD(&t2, i);
        if (val < 0) val += 8;
        setD(&st, i, val);
    }
    for (int a = 0; a < n; a++) st.J[a] ^= t2.J[a];
    g_wordops += n + 8;

    orc_exponential_sum_inplace(&st, eps, p, m);
    *p -= s1->k + s2->k;
}

/* evalW — stabilizer.c:484-488 (ComplexPolar matrix.c:15-18, ComplexMulReal :36-39) */
void orc_evalW(int eps, int p, int m, double* re, double* im) {
    double theta = M_PI * (double)m / 4.;
    double zr = 1 * cos(theta), zi = 1 * sin(theta);
    double r = eps * pow(2., (double)p / 2.);
    *re = zr * r; *im = zi * r;
}

/* extend — stabilizer.c:759-825 */
void orc_extend(S* s, uint64_t xi) {
    const int n = s->n;
    uint64_t Sm = 0;
    for (int a = 0; a < n; a++) Sm |= (uint64_t)par64(xi & s->Gbar[a]) << a;
    uint64_t T = (s->k < n) ? (Sm & ~lowmask(s->k)) : 0;
    if (!T) return;
    int i = __builtin_ctzll(T);
    uint64_t rest = Sm & ~(1ull << i);
    while (rest) {
        int a = __builtin_ctzll(rest); rest &= rest - 1;
        s->Gbar[a] ^= s->Gbar[i];
        s->G[i] ^= s->G[a];
    }
    uint64_t tmp;
    tmp = s->G[i]; s->G[i] = s->G[s->k]; s->G[s->k] = tmp;
    tmp = s->Gbar[i]; s->Gbar[i] = s->Gbar[s->k]; s->Gbar[s->k] = tmp;
    s->k++;
}

/* measurePauli — stabilizer.c:827-959: project onto the +1 eigenspace of
 * i^m Z(zeta) X(xi); returns 0, 1 or 2^-1/2. */
double orc_measure_pauli(S* s, int m, uint64_t zeta, uint64_t xi) {
    const int n = s->n, k = s->k;
    uint64_t vecXi = 0, vecZeta = 0, xiPrime = 0;
    for (int a = 0; a < k; a++) {
        vecXi   |= (uint64_t)par64(s->Gbar[a] & xi) << a;
        vecZeta |= (uint64_t)par64(s->G[a] & zeta) << a;
    }
    for (int a = 0; a < k; a++) if (bit(vecXi, a)) xiPrime ^= s->G[a];

    int w = 2 * m + 4 * (pop64(zeta & s->h) % 2);                           /* eq. 88 */
    for (int b = 0; b < k; b++) w += getD(s, b) * bit(vecXi, b);
    for (int a = 0; a < k; a++)                                             /* a < b < k, J[a][b] */
        w += 4 * pop64(s->J[a] & vecXi & lowmask(k) & ~lowmask(a + 1)) * bit(vecXi, a);
    w = w % 8;

    uint64_t eta = 0;                                                       /* eq. 94 */
    for (int a = 0; a < n; a++) eta |= (uint64_t)par64(s->J[a] & vecXi) << a;
    eta ^= vecZeta;

    if (xi == xiPrime) {
        if (w == 0 || w == 4) {
            uint64_t gamma = 0;
            for (int a = 0; a < k; a++) if (bit(eta, a)) gamma ^= s->Gbar[a];
            int alpha = (w / 4 + pop64(gamma & s->h)) % 2;
            int eps = orc_shrink(s, gamma, alpha, 0);
            if (eps == 0) return 0;
            if (eps == 1) return 1;
            return pow(2, -0.5);
        } else {                                                            /* w in {2,6} */
            int sigma = 2 - w / 2;
            s->Q = (s->Q + sigma) % 8;
            while (s->Q < 0) s->Q += 8;
            for (int a = 0; a < k; a++) {
                int val = getD(s, a) - 2 * sigma * bit(eta, a);
                while (val < 0) val += 8;
                setD(s, a, val);
            }
            for (int i = 0; i < n; i++) if (bit(eta, i)) s->J[i] ^= eta;     /* J_ij ^= eta_i eta_j */
            return pow(2, -0.5);
        }
    }
    orc_extend(s, xi);
    int newD = 2 * m + 4 * (pop64(zeta & xi) % 2) + 4 * (pop64(zeta & s->h) % 2);
    int kk = s->k - 1;
    setD(s, kk, newD);
    /* row and column k-1 of J <- vecZeta, then the diagonal <- m */
    s->J[kk] = vecZeta;
    for (int r = 0; r < n; r++)
        s->J[r] = (s->J[r] & ~(1ull << kk)) | ((uint64_t)bit(vecZeta, r) << kk);
    s->J[kk] = (s->J[kk] & ~(1ull << kk)) | ((uint64_t)(m % 2) << kk);
    return pow(2, -0.5);
}

/* binrep — stateprep.c:5-32: MSB-first; bit j of the string = bit (sz-1-j) of val */
static inline int binbit(unsigned val, int sz, int j) { return (int)((val >> (sz - 1 - j)) & 1u); }

/* prepH — stateprep.c:36-81 */
void orc_prepH(int i, int t, S* phi) {
    int size = (t + 1) / 2;
    orc_identity_state(phi, t, t);
    for (int j = 0; j < size; j++)
        if (binbit(i, size, j) == 0 && !(t % 2 && j == size - 1)) {
            phi->J[2 * j + 1] |= 1ull << (2 * j);
            phi->J[2 * j] |= 1ull << (2 * j + 1);
        }
    for (int j = 0; j < size; j++) {
        if (t % 2 && j == size - 1) {
            if (binbit(i, size, j) == 1) orc_shrink(phi, 1ull << (t - 1), 0, 0);   /* |0> */
            continue;
        }
        if (binbit(i, size, j) == 1)
            orc_shrink(phi, (1ull << (2 * j + 1)) | (1ull << (2 * j)), 0, 0);        /* |00>+|11> */
    }
}

/* x~ of term i: xor of the rows j of L whose MSB-first bit j of i is set — stateprep.c:87-103 */
uint64_t orc_Lbits(int i, int k, const uint64_t* Lrows) {
    uint64_t x = 0;
    for (int j = 0; j < k; j++) if (binbit(i, k, j)) x ^= Lrows[j];
    return x;
}

/* prepL — stateprep.c:85-120 */
void orc_prepL(int i, int t, int k, const uint64_t* Lrows, S* phi) {
    uint64_t x = orc_Lbits(i, k, Lrows);
    orc_identity_state(phi, t, t);
    for (int q = 0; q < t; q++)
        if (!bit(x, q)) orc_shrink(phi, 1ull << q, 0, 0);
}

/* ------------------------------------------------------------------------
 * Random stabilizer states.
 * randomStabilizerState — stabilizer.c:689-756 (logeta :677-687, randDouble :669-674).
 * Two sources of randomness:
 *   libc mode   : consumes rand() in exactly the reference's order, so after the
 *                 same srand() it reproduces the reference's state bit for bit;
 *   philox mode : the counter-based stream the CUDA kernel uses, so the device
 *                 generator can be checked bit for bit on the CPU.
 * ------------------------------------------------------------------------ */
static double logeta(int d, int n) {
    if (d == 0) return 0.;
    double product = 0;
    for (int a = 1; a <= d; a++) {
        product += log2(1 - pow(2, d - n - a));
        product -= log2(1 - pow(2, -a));
    }
    return (-d * (d + 1) / 2) + product;
}

/* cumulative[d], d = 0..n, exactly as stabilizer.c:693-716 computes it */
void orc_dimension_cdf(int n, double* cumulative) {
    double dist[BG_MAX_T + 1], sum = 0;
    for (int d = 0; d <= n; d++) { dist[d] = pow(2, logeta(d, n)); sum += dist[d]; }
    for (int d = 0; d <= n; d++) dist[d] /= sum;
    for (int i = 0; i <= n; i++) {
        cumulative[i] = 0;
        for (int d = 0; d <= i; d++) cumulative[i] += dist[d];
    }
    for (int i = 0; i <= n; i++) cumulative[i] /= cumulative[n];
}

/* BitVectorSetRandom — matrix.c:85-92: one rand()%256 per byte, MSB-first */
static uint64_t libc_random_vector(int n) {
    uint64_t v = 0;
    int bytes = (n + 7) / 8;
    for (int b = 0; b < bytes; b++) {
        unsigned byte = (unsigned)(rand() % 256);
        for (int j = 0; j < 8; j++) {
            int q = 8 * b + j;
            if (q < n && ((byte >> (7 - j)) & 1u)) v |= 1ull << q;
        }
    }
    return v;
}

void orc_random_state_libc(int n, S* s) {
    double cdf[BG_MAX_T + 1];
    orc_dimension_cdf(n, cdf);
    double div = RAND_MAX / 1.0;                       /* randDouble(0,1) */
    double sample = 0. + (rand() / div);
    while (sample == 0.) sample = 0. + (rand() / div);
    int d;
    for (d = 0; d <= n; d++) if (sample <= cdf[d]) break;
    int k = n - d;
    orc_identity_state(s, n, n);
    while (s->k > k) orc_shrink(s, libc_random_vector(n), 0, 1);
    s->h = libc_random_vector(n);
    s->D1 = libc_random_vector(n);
    s->D2 = libc_random_vector(n);
    for (int i = 0; i < k; i++) {
        s->J[i] = (s->J[i] & ~(1ull << i)) | ((uint64_t)bit(s->D1, i) << i);
        for (int j = 0; j < i; j++) {
            unsigned val = (unsigned)rand();
            uint64_t b = val % 2;
            s->J[i] = (s->J[i] & ~(1ull << j)) | (b << j);
            s->J[j] = (s->J[j] & ~(1ull << i)) | (b << i);
        }
    }
}

/* Philox4x32-10 (Salmon et al., SC'11), the counter-based generator of the
 * CUDA path.  key = seed, counter = (sample lo, sample hi, bin, block). */
static void philox4x32_10(uint32_t ctr[4], const uint32_t key_in[2]) {
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * ctr[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * ctr[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ ctr[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ ctr[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
void orc_philox_block(uint64_t seed, uint64_t sample, uint32_t bin, uint32_t block, uint32_t out[4]) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    out[0] = (uint32_t)sample; out[1] = (uint32_t)(sample >> 32); out[2] = bin; out[3] = block;
    philox4x32_10(out, key);
}
static inline uint64_t u64_of(const uint32_t w[4], int half) {
    return ((uint64_t)w[2 * half + 1] << 32) | w[2 * half];
}

/* The device sampler (circuitsimulator_b200/csrc/bg_philox.cuh:native_random) restated on the
 * CPU.  Same law as randomStabilizerState (stabilizer.c:689-756): d from the eq. 79 cdf, K cut
 * from F_2^n by d random hyperplanes (lazy shrink), h, D uniform, J uniform symmetric with
 * J_aa = D1_a, Q = 0.  It differs from the libc version only in bookkeeping: the device keeps
 * an active-row mask instead of swapping the pivot row to position k-1, pivots on the LOWEST
 * row of S, and draws (D, J) per row index before the rows are compacted.
 * Stream layout (key = seed, counter = (sample lo, sample hi, bin, block)):
 *   block 0            : words 0,1 -> u = ((w1:w0 >> 11) + 1) 2^-53 in (0,1]  -> d
 *   block 1 + j/2      : half j%2 -> xi_j, the j-th random hyperplane
 *   block 0x1000       : half 0 -> h, half 1 -> D1
 *   block 0x1001       : half 0 -> D2
 *   block 0x2000 + v   : half 0 -> r_v ; J_uv = bit v of r_u for v < u (both active), J_vv = D1_v */
void orc_random_state_philox(int n, uint64_t seed, uint32_t bin, uint64_t sample, S* out) {
    double cdf[BG_MAX_T + 1];
    orc_dimension_cdf(n, cdf);
    uint32_t w[4];
    orc_philox_block(seed, sample, bin, 0, w);
    double u = (double)((u64_of(w, 0) >> 11) + 1ull) * 0x1.0p-53;
    int d = 0;
    while (d < n && !(u <= cdf[d])) d++;
    int k = n - d;
    const uint64_t mn = lowmask(n);
    uint64_t G[BG_MAX_T], Gb[BG_MAX_T], J[BG_MAX_T], A = mn;
    for (int i = 0; i < n; i++) { G[i] = Gb[i] = 1ull << i; J[i] = 0; }
    for (uint32_t j = 0; pop64(A) > k && j < 100000u; j++) {
        orc_philox_block(seed, sample, bin, 1 + j / 2, w);
        uint64_t xi = u64_of(w, j % 2) & mn, Sm = 0;
        for (int a = 0; a < n; a++) if (bit(A, a) && par64(G[a] & xi)) Sm |= 1ull << a;
        if (!Sm) continue;                                   /* xi orthogonal to K: SAME */
        int i = __builtin_ctzll(Sm);
        uint64_t Sp = Sm & ~(1ull << i);
        for (int a = 0; a < n; a++) if (bit(Sp, a)) { G[a] ^= G[i]; Gb[i] ^= Gb[a]; }
        A &= ~(1ull << i);
    }
    orc_philox_block(seed, sample, bin, 0x1000, w);
    uint64_t h = u64_of(w, 0) & mn, D1 = u64_of(w, 1) & A;
    orc_philox_block(seed, sample, bin, 0x1001, w);
    uint64_t D2 = u64_of(w, 0) & A;
    for (int v = 0; v < n; v++) {
        if (!bit(A, v)) continue;
        orc_philox_block(seed, sample, bin, 0x2000 + v, w);
        uint64_t r = u64_of(w, 0) & lowmask(v) & A;
        J[v] |= r | ((uint64_t)bit(D1, v) << v);
        for (int c = 0; c < v; c++) if (bit(r, c)) J[c] |= 1ull << v;
    }
    /* compact: active rows first, order preserved */
    int perm[BG_MAX_T], r = 0;
    for (int v = 0; v < n; v++) if (bit(A, v)) perm[r++] = v;
    for (int v = 0; v < n; v++) if (!bit(A, v)) perm[r++] = v;
    memset(out, 0, sizeof(*out));
    out->n = n; out->k = k; out->Q = 0; out->h = h;
    for (int i = 0; i < n; i++) {
        out->G[i] = G[perm[i]]; out->Gbar[i] = Gb[perm[i]];
        if (i < k) {
            out->D1 |= (uint64_t)bit(D1, perm[i]) << i;
            out->D2 |= (uint64_t)bit(D2, perm[i]) << i;
            for (int c = 0; c < k; c++) out->J[i] |= (uint64_t)bit(J[perm[i]], perm[c]) << c;
        }
    }
}

/* ------------------------------------------------------------------------
 * The L x chi loop
 * ------------------------------------------------------------------------ */

/* body of singleProjectorSample after theta is drawn — innerprod.c:100-142.
 * theta is updated in place (projected).  epm (3*chi ints) optional. */
double orc_sample_from_theta(S* theta, const bg_projector* P, int exact, int k, const uint64_t* Lrows,
                             int32_t* epm, int* alive, double* total_re, double* total_im,
                             double* projfactor_out) {
    const int t = P->nqubits;
    double projfactor = 1;
    *alive = 1;
    for (int i = 0; i < P->nstabs; i++) {
        double res = orc_measure_pauli(theta, P->phase[i], P->zs[i], P->xs[i]);
        projfactor *= res;
        if (res == 0) { *alive = 0; break; }
    }
    if (projfactor_out) *projfactor_out = projfactor;
    if (!*alive) { if (total_re) { *total_re = 0; *total_im = 0; } return 0; }
    int chi = exact ? (1 << ((t + 1) / 2)) : (1 << k);
    double tre = 0, tim = 0;
    for (int i = 0; i < chi; i++) {
        S phi;
        if (exact) orc_prepH(i, t, &phi); else orc_prepL(i, t, k, Lrows, &phi);
        int eps, p, m;
        orc_inner_product(theta, &phi, &eps, &p, &m);
        if (epm) { epm[3 * i] = eps; epm[3 * i + 1] = p; epm[3 * i + 2] = m; }
        double re, im;
        orc_evalW(eps, p, m, &re, &im);
        tre += re; tim += im;
    }
    if (total_re) { *total_re = tre; *total_im = tim; }
    double zr = tre * projfactor, zi = tim * projfactor;
    return pow(2, t) * (zr * zr + zi * zi);
}

/* sampledProjector — innerprod.c:45-84, single rank, theta from the Philox stream
 * (sample indices first, first+stride, ...; `count` of them). Returns the SUM. */
double orc_sampled_sum_philox(const bg_projector* P, int exact, int k, const uint64_t* Lrows,
                              uint64_t seed, uint32_t bin, uint64_t first, uint64_t stride, uint64_t count,
                              double* per_sample) {
    double total = 0;
    for (uint64_t c = 0; c < count; c++) {
        uint64_t l = first + c * stride;
        S theta; int alive;
        orc_random_state_philox(P->nqubits, seed, bin, l, &theta);
        double v = orc_sample_from_theta(&theta, P, exact, k, Lrows, NULL, &alive, NULL, NULL, NULL);
        if (per_sample) per_sample[c] = v;
        total += v;
    }
    return total;
}

/* sampledProjector — innerprod.c:45-84 with libc rand() thetas (srand first) */
double orc_sampled_projector_libc(const bg_projector* P, int exact, int k, const uint64_t* Lrows,
                                  double norm, int samples) {
    if (P->nstabs == 0) return pow(norm, 2);
    double total = 0;
    for (int i = 0; i < samples; i++) {
        S theta; int alive;
        orc_random_state_libc(P->nqubits, &theta);
        total += orc_sample_from_theta(&theta, P, exact, k, Lrows, NULL, &alive, NULL, NULL, NULL);
    }
    return total / samples;
}

/* exactProjectorWork — innerprod.c:203-261 */
void orc_exact_projector_work(int l, const bg_projector* P, int exact, int k, const uint64_t* Lrows,
                              double* re, double* im) {
    const int t = P->nqubits;
    int chi = exact ? (1 << ((t + 1) / 2)) : (1 << k);
    int i = 0;
    while (l >= chi - i) { l -= chi - i; i += 1; }
    int j = l + i;
    S theta, phi;
    if (exact) orc_prepH(i, t, &theta); else orc_prepL(i, t, k, Lrows, &theta);
    double projfactor = 1;
    for (int r = 0; r < P->nstabs; r++) {
        double res = orc_measure_pauli(&theta, P->phase[r], P->zs[r], P->xs[r]);
        projfactor *= res;
        if (res == 0) { *re = 0; *im = 0; return; }
    }
    if (exact) orc_prepH(j, t, &phi); else orc_prepL(j, t, k, Lrows, &phi);
    int eps, p, m;
    orc_inner_product(&theta, &phi, &eps, &p, &m);
    double wr, wi;
    orc_evalW(eps, p, m, &wr, &wi);
    if (i == j) { *re = wr * projfactor; *im = wi * projfactor; }
    else { *re = wr * (2 * projfactor); *im = 0; }
}

/* exactProjector — innerprod.c:148-199, single rank */
double orc_exact_projector(const bg_projector* P, int exact, int k, const uint64_t* Lrows, double norm) {
    if (P->nstabs == 0) return pow(norm, 2);
    const int t = P->nqubits;
    if (t == 0) {
        double sum = 1;
        for (int i = 0; i < P->nstabs; i++) {
            if (P->phase[i] == 0) sum += 1;
            if (P->phase[i] == 2) sum -= 1;
        }
        return sum / (1 + (double)P->nstabs);
    }
    int size = exact ? (t + 1) / 2 : k;
    long kRange = (long)(pow(2, size - 1) * (pow(2, size) + 1));
    double tr = 0, ti = 0;
    for (long l = 0; l < kRange; l++) {
        double re, im;
        orc_exact_projector_work((int)l, P, exact, k, Lrows, &re, &im);
        tr += re; ti += im;
    }
    return sqrt(tr * tr + ti * ti);
}

/* decompose()'s fidelity loop — probability.c:373-391.  Z(L) = sum over the 2^k combinations of
 * the rows of L of pow(2, -hamming/2): `hamming/2` is an INTEGER division in the reference (both
 * operands are ints), kept as is.  Also returns the weight histogram the CUDA path produces.
 * hist may be NULL. */
double orc_decompose_ZL(int t, int k, const uint64_t* Lrows, uint64_t* hist) {
    double Z_L = 0;
    if (hist) memset(hist, 0, 65 * sizeof(uint64_t));
    for (uint64_t i = 0; i < (1ull << k); i++) {
        uint64_t x = 0;
        for (int j = 0; j < k; j++) if ((i >> (k - 1 - j)) & 1ull) x ^= Lrows[j];     /* binrep is MSB-first */
        int hamming = 0;
        for (int q = 0; q < t; q++) hamming += bit(x, q);
        if (hist) hist[hamming]++;
        Z_L += pow(2, -hamming / 2);
    }
    return Z_L;
}

size_t orc_sizeof_state(void) { return sizeof(bg_state); }
size_t orc_sizeof_projector(void) { return sizeof(bg_projector); }
