/* ref_driver.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin adapter between the packed bg_state layout (include/bgnorm.h) and the
 * UNMODIFIED reference C sources, which are compiled where they lie under
 * /root/reference by oracle/Makefile into oracle/_ref/libcircref.so.  Every
 * function here only converts layouts and then calls a reference function; no
 * algorithm is implemented in this file.
 *
 * Reference entry points used (paths relative to the reference root):
 *   libcirc/stabilizer/stabilizer.c : allocStabilizerState :9, exponentialSumExact :300,
 *       shrink :500, innerProductExact :589, randomStabilizerState :689, extend :759,
 *       measurePauli :827, getD/setD :55-62
 *   libcirc/stateprep.c             : prepH :36, prepL :85
 *   libcirc/innerprod.c             : singleProjectorSample :88, exactProjectorWork :203,
 *       sampledProjector :45, exactProjector :148
 */
#include <stdint.h>
#include <string.h>
#include <math.h>
#include "stabilizer/stabilizer.h"   /* reference header, found via -I<reference>/libcirc */
#include "bgnorm.h"

struct StabilizerState* prepH(int i, int t);
struct StabilizerState* prepL(int i, int t, struct BitMatrix* L);
double singleProjectorSample(struct Projector* P, struct BitMatrix* L, int exact);
Complex exactProjectorWork(int i, struct Projector* P, struct BitMatrix* L, int exact);
double sampledProjector(struct Projector* P, struct BitMatrix* L, int exact, double norm, int samples);
double exactProjector(struct Projector* P, struct BitMatrix* L, int exact, double norm);

/* ---------------- layout conversion ---------------- */

static struct BitVector* vec_from_u64(uint64_t v, int n) {
    struct BitVector* out = newBitVector(n);
    for (int i = 0; i < n; i++) BitVectorSet(out, i, (unsigned)((v >> i) & 1));
    return out;
}
static uint64_t vec_to_u64(struct BitVector* v) {
    uint64_t out = 0;
    for (unsigned i = 0; i < v->size; i++) out |= (uint64_t)BitVectorGet(v, i) << i;
    return out;
}
static void mat_from_rows(struct BitMatrix* M, const uint64_t* rows) {
    BitMatrixSetZero(M);
    for (unsigned r = 0; r < M->rows; r++)
        for (unsigned c = 0; c < M->cols; c++)
            BitMatrixSet(M, r, c, (unsigned)((rows[r] >> c) & 1));
}
static void mat_to_rows(struct BitMatrix* M, uint64_t* rows) {
    for (unsigned r = 0; r < M->rows; r++) {
        uint64_t w = 0;
        for (unsigned c = 0; c < M->cols; c++) w |= (uint64_t)BitMatrixGet(M, r, c) << c;
        rows[r] = w;
    }
}

static struct StabilizerState* state_from_packed(const bg_state* s) {
    struct StabilizerState* st = allocStabilizerState(s->n, s->k);
    st->Q = s->Q;
    for (int i = 0; i < s->n; i++) {
        BitVectorSet(st->h, i, (unsigned)((s->h >> i) & 1));
        BitVectorSet(st->D1, i, (unsigned)((s->D1 >> i) & 1));
        BitVectorSet(st->D2, i, (unsigned)((s->D2 >> i) & 1));
    }
    mat_from_rows(st->G, s->G);
    mat_from_rows(st->Gbar, s->Gbar);
    mat_from_rows(st->J, s->J);
    return st;
}
static void state_to_packed(struct StabilizerState* st, bg_state* s) {
    memset(s, 0, sizeof(*s));
    s->n = st->n; s->k = st->k; s->Q = st->Q;
    /* h, D1, D2 may carry junk in the pad bits of their last byte
     * (BitVectorSetRandom, matrix.c:85-92); reading bit by bit drops it. */
    for (int i = 0; i < st->n; i++) {
        s->h  |= (uint64_t)BitVectorGet(st->h, i)  << i;
        s->D1 |= (uint64_t)BitVectorGet(st->D1, i) << i;
        s->D2 |= (uint64_t)BitVectorGet(st->D2, i) << i;
    }
    mat_to_rows(st->G, s->G);
    mat_to_rows(st->Gbar, s->Gbar);
    mat_to_rows(st->J, s->J);
}

static struct Projector* projector_from_packed(const bg_projector* P) {
    struct Projector* out = (struct Projector*)malloc(sizeof(struct Projector));
    out->Nstabs = P->nstabs; out->Nqubits = P->nqubits;
    if (P->nstabs == 0) return out;
    out->phaseSign = newBitVector(P->nstabs);
    out->phaseComplex = newBitVector(P->nstabs);
    out->xs = newBitMatrixZero(P->nstabs, P->nqubits);
    out->zs = newBitMatrixZero(P->nstabs, P->nqubits);
    for (int i = 0; i < P->nstabs; i++) {
        BitVectorSet(out->phaseComplex, i, P->phase[i] % 2);
        BitVectorSet(out->phaseSign, i, (P->phase[i] / 2) % 2);
        for (int j = 0; j < P->nqubits; j++) {
            BitMatrixSet(out->xs, i, j, (unsigned)((P->xs[i] >> j) & 1));
            BitMatrixSet(out->zs, i, j, (unsigned)((P->zs[i] >> j) & 1));
        }
    }
    return out;
}
static void projector_free(struct Projector* P) {
    if (P->Nstabs > 0) {
        BitVectorFree(P->phaseSign); BitVectorFree(P->phaseComplex);
        BitMatrixFree(P->xs); BitMatrixFree(P->zs);
    }
    free(P);
}
static struct BitMatrix* L_from_rows(int k, int t, const uint64_t* rows) {
    struct BitMatrix* L = newBitMatrixZero(k, t);
    mat_from_rows(L, rows);
    return L;
}

/* ---------------- exported wrappers ---------------- */

void ref_srand(unsigned seed) { srand(seed); }

void ref_exponential_sum(const bg_state* s, int* eps, int* p, int* m) {
    struct StabilizerState* st = state_from_packed(s);
    exponentialSumExact(st, eps, p, m);
    freeStabilizerState(st);
}

int ref_shrink(bg_state* s, uint64_t xi, int alpha, int lazy) {
    struct StabilizerState* st = state_from_packed(s);
    struct BitVector* v = vec_from_u64(xi, s->n);
    int status = shrink(st, v, alpha, lazy);
    state_to_packed(st, s);
    BitVectorFree(v);
    freeStabilizerState(st);
    return status;
}

void ref_extend(bg_state* s, uint64_t xi) {
    struct StabilizerState* st = state_from_packed(s);
    struct BitVector* v = vec_from_u64(xi, s->n);
    extend(st, v);
    state_to_packed(st, s);
    BitVectorFree(v);
    freeStabilizerState(st);
}

double ref_measure_pauli(bg_state* s, int m, uint64_t zeta, uint64_t xi) {
    struct StabilizerState* st = state_from_packed(s);
    struct BitVector* vz = vec_from_u64(zeta, s->n);
    struct BitVector* vx = vec_from_u64(xi, s->n);
    double r = measurePauli(st, m, vz, vx);
    state_to_packed(st, s);
    BitVectorFree(vz); BitVectorFree(vx);
    freeStabilizerState(st);
    return r;
}

void ref_inner_product(const bg_state* a, const bg_state* b, int* eps, int* p, int* m) {
    struct StabilizerState* s1 = state_from_packed(a);
    struct StabilizerState* s2 = state_from_packed(b);
    innerProductExact(s1, s2, eps, p, m);
    freeStabilizerState(s1);
    freeStabilizerState(s2);
}

/* the reference's evalW (stabilizer.c:484-488) so tests can reproduce its fp64 */
void ref_evalW(int eps, int p, int m, double* re, double* im) {
    Complex evalW(int eps, int p, int m);
    Complex z = evalW(eps, p, m);
    *re = z.re; *im = z.im;
}

/* randomStabilizerState(n) using libc rand(); call ref_srand first */
void ref_random_state(int n, bg_state* out) {
    struct StabilizerState* st = randomStabilizerState(n);
    state_to_packed(st, out);
    freeStabilizerState(st);
}

void ref_prepH(int i, int t, bg_state* out) {
    struct StabilizerState* st = prepH(i, t);
    state_to_packed(st, out);
    freeStabilizerState(st);
}

void ref_prepL(int i, int t, int k, const uint64_t* Lrows, bg_state* out) {
    struct BitMatrix* L = L_from_rows(k, t, Lrows);
    struct StabilizerState* st = prepL(i, t, L);
    state_to_packed(st, out);
    freeStabilizerState(st);
    BitMatrixFree(L);
}

/* singleProjectorSample (innerprod.c:88-144) exactly as the reference runs it:
 * theta drawn from libc rand().  Call ref_srand first. */
double ref_single_projector_sample(const bg_projector* P, int exact, int k, const uint64_t* Lrows) {
    struct Projector* rp = projector_from_packed(P);
    struct BitMatrix* L = exact ? NULL : L_from_rows(k, P->nqubits, Lrows);
    double v = singleProjectorSample(rp, L, exact);
    if (L) BitMatrixFree(L);
    projector_free(rp);
    return v;
}

/* sampledProjector (innerprod.c:45-84), world_size 1 */
double ref_sampled_projector(const bg_projector* P, int exact, int k, const uint64_t* Lrows,
                             double norm, int samples) {
    struct Projector* rp = projector_from_packed(P);
    struct BitMatrix* L = exact ? NULL : L_from_rows(k, P->nqubits, Lrows);
    double v = sampledProjector(rp, L, exact, norm, samples);
    if (L) BitMatrixFree(L);
    projector_free(rp);
    return v;
}

/* exactProjector (innerprod.c:148-199), world_size 1 */
double ref_exact_projector(const bg_projector* P, int exact, int k, const uint64_t* Lrows, double norm) {
    struct Projector* rp = projector_from_packed(P);
    struct BitMatrix* L = exact ? NULL : L_from_rows(k, P->nqubits, Lrows);
    double v = exactProjector(rp, L, exact, norm);
    if (L) BitMatrixFree(L);
    projector_free(rp);
    return v;
}

void ref_exact_projector_work(int l, const bg_projector* P, int exact, int k, const uint64_t* Lrows,
                              double* re, double* im) {
    struct Projector* rp = projector_from_packed(P);
    struct BitMatrix* L = exact ? NULL : L_from_rows(k, P->nqubits, Lrows);
    Complex z = exactProjectorWork(l, rp, L, exact);
    *re = z.re; *im = z.im;
    if (L) BitMatrixFree(L);
    projector_free(rp);
}

/* The body of singleProjectorSample (innerprod.c:100-142) replayed on a
 * caller-supplied theta, calling the same reference functions in the same
 * order, so that the projected theta and the per-pair (eps,p,m) can be dumped.
 * Outputs: theta is replaced by the projected state; epm gets 3*chi ints;
 * returns the sample value exactly as innerprod.c:142 computes it.
 * If the projector annihilates theta, returns 0 and *alive = 0. */
double ref_sample_from_theta(bg_state* theta, const bg_projector* P, int exact, int k,
                             const uint64_t* Lrows, int32_t* epm, int* alive,
                             double* total_re, double* total_im, double* projfactor_out) {
    int t = P->nqubits;
    struct StabilizerState* th = state_from_packed(theta);
    double projfactor = 1;
    *alive = 1;
    for (int i = 0; i < P->nstabs; i++) {
        struct BitVector* zeta = vec_from_u64(P->zs[i], t);
        struct BitVector* xi = vec_from_u64(P->xs[i], t);
        double res = measurePauli(th, P->phase[i], zeta, xi);
        projfactor *= res;
        BitVectorFree(zeta); BitVectorFree(xi);
        if (res == 0) { *alive = 0; break; }
    }
    state_to_packed(th, theta);
    if (projfactor_out) *projfactor_out = projfactor;
    if (!*alive) { freeStabilizerState(th); if (total_re) { *total_re = 0; *total_im = 0; } return 0; }

    struct BitMatrix* L = exact ? NULL : L_from_rows(k, t, Lrows);
    int chi = exact ? (1 << ((t + 1) / 2)) : (1 << k);
    Complex total = {0, 0};
    for (int i = 0; i < chi; i++) {
        struct StabilizerState* phi = exact ? prepH(i, t) : prepL(i, t, L);
        int eps, p, m;
        innerProductExact(th, phi, &eps, &p, &m);
        if (epm) { epm[3*i] = eps; epm[3*i+1] = p; epm[3*i+2] = m; }
        Complex evalW(int eps, int p, int m);
        total = ComplexAdd(total, evalW(eps, p, m));
        freeStabilizerState(phi);
    }
    if (L) BitMatrixFree(L);
    freeStabilizerState(th);
    if (total_re) { *total_re = total.re; *total_im = total.im; }
    return pow(2, t) * ComplexMagSquare(ComplexMulReal(total, projfactor));
}

/* The reference's own containers as the level-2 binding sees them: a struct Projector / BitMatrix built by the
 * reference's constructors and setters (matrix.c), handed over as raw .data byte arrays.  Copies them out so that
 * the test can push them through bg_projector_from_bitmatrix / bg_set_decomposition_bitmatrix (matrix.c:124-131,
 * 330-339: MSB-first, no row padding). */
int ref_projector_bytes(const bg_projector* P, unsigned char* phase_sign, unsigned char* phase_complex,
                        unsigned char* xs, unsigned char* zs, int cap) {
    struct Projector* R = projector_from_packed(P);
    const int nb_v = (P->nstabs + 7) / 8, nb_m = (P->nstabs * P->nqubits + 7) / 8;
    if (P->nstabs == 0 || nb_m > cap) { projector_free(R); return -1; }
    memcpy(phase_sign, R->phaseSign->data, nb_v); memcpy(phase_complex, R->phaseComplex->data, nb_v);
    memcpy(xs, R->xs->data, nb_m); memcpy(zs, R->zs->data, nb_m);
    projector_free(R);
    return nb_m;
}
int ref_L_bytes(int k, int t, const uint64_t* rows, unsigned char* out, int cap) {
    struct BitMatrix* L = L_from_rows(k, t, rows);
    const int nb = (k * t + 7) / 8;
    if (nb > cap) { BitMatrixFree(L); return -1; }
    memcpy(out, L->data, nb);
    BitMatrixFree(L);
    return nb;
}

size_t ref_sizeof_state(void) { return sizeof(bg_state); }
size_t ref_sizeof_projector(void) { return sizeof(bg_projector); }
