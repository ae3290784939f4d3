/* oracle/level2_shim.c — TEST INFRASTRUCTURE: the level-2 binding of INTEGRATION.md, built for real.
 *
 * The reference's back-end host (libcirc/probability.c: main / master / decompose, UNMODIFIED, compiled where it
 * lies) is linked against THIS file instead of libcirc/innerprod.c: master() calls multiSampledProjector /
 * exactProjector (probability.c:197-201) and gets the C ABI of include/bgnorm.h.  The reference's containers are
 * passed as they are — struct Projector with BitVector / BitMatrix .data byte arrays (comms.h:4-11, matrix.c:124-131,
 * 330-339) — through the *_bitmatrix adapters.  oracle/Makefile: _ref/mpibackend_bg.
 *
 * Nothing here is part of the product; it proves that the adapters a maintainer would call work on the
 * reference's own memory layout.  Seeds: BG_SEED (default: pid, like srand(getpid()), probability.c:182), derived
 * exactly as bgbackend does, so that both executables print the same numbers for the same stream. */
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <math.h>

#include "utils/comms.h"        /* struct Projector, struct BitMatrix (reference headers, -I$(REF)/libcirc) */
#include "bgnorm.h"

static bg_ctx* g_ctx = NULL;
static int g_calls = 0;

static unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static unsigned long long next_seed(void) {
    const char* es = getenv("BG_SEED");
    const unsigned long long seed0 = es ? strtoull(es, NULL, 0) : (unsigned long long)getpid();
    return splitmix64(seed0 + (unsigned long long)(g_calls++));      /* G' then H': seed0, seed0 + 1 */
}
static void die(const char* what) {
    printf("Error: %s: %s\n", what, bg_last_error(g_ctx));
    exit(0);                                                          /* the front end looks at stdout, not at the status */
}
static void setup(struct Projector* P, struct BitMatrix* L, int exact, bg_projector* bp) {
    if (!g_ctx && bg_init(&g_ctx, 0)) die("bg_init");
    if (bg_set_decomposition_bitmatrix(g_ctx, P->Nqubits, exact, exact ? 0 : L->rows, exact ? NULL : L->data))
        die("bg_set_decomposition_bitmatrix");
    if (bg_projector_from_bitmatrix(bp, P->Nstabs, P->Nqubits, P->phaseSign->data, P->phaseComplex->data,
                                    P->xs->data, P->zs->data)) die("bg_projector_from_bitmatrix");
}

double multiSampledProjector(struct Projector* P, struct BitMatrix* L, int exact, double norm, int samples, int bins) {
    bg_projector bp;
    double out = 0;
    if (P->Nstabs == 0) { g_calls++; return pow(norm, 2); }           /* innerprod.c:47 */
    if (P->Nqubits == 0) {                                            /* Clifford circuit: closed form inside the library */
        bp.nstabs = P->Nstabs; bp.nqubits = 0;
        for (int i = 0; i < P->Nstabs; i++)
            bp.phase[i] = (unsigned char)(2 * ((P->phaseSign->data[i / 8] >> (7 - i % 8)) & 1) + ((P->phaseComplex->data[i / 8] >> (7 - i % 8)) & 1));
        if (!g_ctx && bg_init(&g_ctx, 0)) die("bg_init");
        if (bg_sampled_norm(g_ctx, &bp, (unsigned long long)samples, bins, 0, norm, &out)) die("bg_sampled_norm");
        g_calls++;
        return out;
    }
    setup(P, L, exact, &bp);
    if (bg_sampled_norm(g_ctx, &bp, (unsigned long long)samples, bins, next_seed(), norm, &out)) die("bg_sampled_norm");
    return out;
}

double exactProjector(struct Projector* P, struct BitMatrix* L, int exact, double norm) {
    bg_projector bp;
    double out = 0;
    if (P->Nstabs == 0) return pow(norm, 2);                          /* innerprod.c:150 */
    if (P->Nqubits == 0) {
        bp.nstabs = P->Nstabs; bp.nqubits = 0;
        for (int i = 0; i < P->Nstabs; i++)
            bp.phase[i] = (unsigned char)(2 * ((P->phaseSign->data[i / 8] >> (7 - i % 8)) & 1) + ((P->phaseComplex->data[i / 8] >> (7 - i % 8)) & 1));
        if (!g_ctx && bg_init(&g_ctx, 0)) die("bg_init");
        if (bg_exact_norm(g_ctx, &bp, norm, &out)) die("bg_exact_norm");
        return out;
    }
    setup(P, L, exact, &bp);
    if (bg_exact_norm(g_ctx, &bp, norm, &out)) die("bg_exact_norm");
    return out;
}

/* slave() is never entered with one rank (probability.c:40-44), but it references these */
double singleProjectorSample(struct Projector* P, struct BitMatrix* L, int exact) {
    (void)P; (void)L; (void)exact;
    printf("Error: singleProjectorSample called in the level-2 build\n");
    exit(0);
}
Complex exactProjectorWork(int i, struct Projector* P, struct BitMatrix* L, int exact) {
    (void)i; (void)P; (void)L; (void)exact;
    printf("Error: exactProjectorWork called in the level-2 build\n");
    exit(0);
}
