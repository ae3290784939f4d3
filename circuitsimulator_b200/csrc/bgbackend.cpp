// bgbackend.cpp — drop-in replacement for the reference's back-end executable
// (`libcirc/mpibackend`, built from libcirc/probability.c; the help text still calls it
// `libcirc/sample`).  Same argv / stdin token protocol, same stdout contract, so the
// UNMODIFIED Python front end (main.py, libcirc/probability.py:237-312, libcirc/sample.py)
// drives it with   cpath=<this binary> mpirun="/usr/bin/env"   (see INTEGRATION.md).
//
//   argv[1]      file name or "stdin"; absent -> first stdin token is the file name   (probability.c:52-68)
//   tokens       quiet verbose noapprox samples bins t k exact fidbound fidelity rank forceL forceSample,
//                then projectors G and H                                  (probability.c:74-127, comms.c:9-36)
//   stdout       chatter lines, then numerator and denominator as the LAST TWO lines, %.17e
//                                                                         (probability.c:207-216)
//
// The host keeps what the reference's master() keeps — parsing, decompose(), the
// sampled/exact decision, printing — and calls the C ABI (include/bgnorm.h) where the
// reference calls multiSampledProjector / exactProjector.  The MPI master/worker fan-out
// (probability.c:184-205, 221-299) is replaced by BG_GPUS host threads, one context per GPU,
// samples strided across them, and one NCCL all-reduce of the partial sums.
//
// Environment: BG_SEED (default: pid, as probability.c:182), BG_GPUS (default 1),
//              BG_DEVICE (first device, default 0), BG_QUIETER=1 (drop the chatter).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <string>
#include <thread>
#include <vector>

#include "bgnorm.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

struct Config {
    int quiet = 0, verbose = 0, noapprox = 0, samples = 0, bins = 1, t = 0, k = 0, exact = 1;
    double fidbound = 1e-5;
    int fidelity = 0, rank = 0, forceL = 0, forceSample = 0;
};

static bool read_int(FILE* f, int* v) { return fscanf(f, "%d", v) == 1; }

// readProjector (libcirc/utils/comms.c:9-36): Nstabs Nqubits, then per generator the phase
// (0..3) followed by an (x, z) pair per qubit.
static bool read_projector(FILE* f, bg_projector* P, std::string* err) {
    memset(P, 0, sizeof *P);
    int ns = 0, nq = 0;
    if (!read_int(f, &ns) || !read_int(f, &nq)) { *err = "truncated projector header"; return false; }
    if (ns < 0 || ns > BG_MAX_STABS) { *err = "too many generators"; return false; }
    if (nq < 0 || nq > BG_MAX_T) { *err = "more than 64 magic-state qubits (t) are not supported"; return false; }
    P->nstabs = ns; P->nqubits = nq;
    for (int i = 0; i < ns; i++) {
        int v;
        if (!read_int(f, &v)) { *err = "truncated projector"; return false; }
        P->phase[i] = (uint8_t)(((v / 2) % 2) * 2 + (v % 2));
        for (int q = 0; q < nq; q++) {
            int x, z;
            if (!read_int(f, &x) || !read_int(f, &z)) { *err = "truncated projector"; return false; }
            if (x % 2 == 1) P->xs[i] |= 1ull << q;
            if (z % 2 == 1) P->zs[i] |= 1ull << q;
        }
    }
    return true;
}

// rank over F_2 (the value BitMatrixRank returns, libcirc/utils/matrix.c:662-724)
static int f2_rank(std::vector<uint64_t> rows) {
    int rank = 0;
    for (int c = 0; c < 64; c++) {
        int piv = -1;
        for (size_t r = rank; r < rows.size(); r++) if ((rows[r] >> c) & 1) { piv = (int)r; break; }
        if (piv < 0) continue;
        std::swap(rows[rank], rows[piv]);
        for (size_t r = 0; r < rows.size(); r++) if ((int)r != rank && ((rows[r] >> c) & 1)) rows[r] ^= rows[rank];
        rank++;
    }
    return rank;
}

// BitMatrixSetRandom (libcirc/utils/matrix.c:301-306): one rand()%256 per byte of the k*t-bit
// row-major, MSB-first array.  Drawn in the same order so the same libc state gives the same L.
static void random_L(int k, int t, std::vector<uint64_t>& rows) {
    rows.assign(k, 0);
    const unsigned bytes = (unsigned)(k * t + 7) / 8;
    for (unsigned b = 0; b < bytes; b++) {
        const unsigned byte = (unsigned)(rand() % 256);
        for (int j = 0; j < 8; j++) {
            const unsigned loc = 8 * b + j;
            if (loc >= (unsigned)(k * t)) break;
            if ((byte >> (7 - j)) & 1u) rows[loc / t] |= 1ull << (loc % t);
        }
    }
}

// decompose (libcirc/probability.c:307-415): choose |H^t> (exact) or |L> with a random k x t L.
static void decompose(Config& c, std::vector<uint64_t>& L, double* norm) {
    const int t = c.t;
    if (t == 0) { c.exact = 0; *norm = 1; return; }
    const double v = cos(M_PI / 8);
    *norm = pow(2, floor((float)t / 2) / 2);
    if (t % 2) *norm *= 2 * v;
    if (c.exact) return;
    const bool forceK = c.k > 0;
    if (!forceK) {
        c.k = (int)ceil(1 - 2 * t * log2(v) - log2(c.fidbound));
        if (c.verbose) printf("Autopicking k = %d to achieve delta = %f.\n", c.k, c.fidbound);
    }
    if (c.k > t / 2 && !forceK && !c.forceL) {
        if (c.verbose) printf("k > t/2. Reverting to exact decomposition.\n");
        c.exact = 1;
        return;
    }
    if (c.k > t) {
        if (forceK && !c.quiet) printf("Can't have k > t. Setting k to %d.\n", t);
        c.k = t;
    }
    double overlap = 0, Z_L = 0;
    while (overlap < 1 - c.fidbound || forceK) {
        random_L(c.k, t, L);
        if (c.rank && f2_rank(L) < c.k) {
            if (!c.quiet) printf("L has insufficient rank. Sampling again...\n");
            continue;
        }
        if (!c.fidelity) break;
        // Z(L) = sum_x 2^{-|x|/2} over the 2^k combinations of rows; the reference evaluates
        // pow(2, -hamming/2) with INTEGER division (probability.c:389), kept as is.
        Z_L = 0;
        for (uint64_t i = 0; i < (1ull << c.k); i++) {
            uint64_t x = 0;
            for (int j = 0; j < c.k; j++) if ((i >> (c.k - 1 - j)) & 1) x ^= L[j];
            const int hamming = __builtin_popcountll(x);
            Z_L += pow(2, -hamming / 2);
        }
        overlap = pow(2, c.k) * pow(v, 2 * t) / Z_L;
        if (forceK) { printf("delta = 1 - <H^t|L>: %lf\n", 1 - overlap); break; }
        if (overlap < 1 - c.fidbound) { if (!c.quiet) printf("delta = 1 - <H^t|L>: %lf - Not good enough!\n", 1 - overlap); }
        else if (!c.quiet) printf("delta = 1 - <H^t|L>: %lf\n", 1 - overlap);
    }
    if (c.fidelity) *norm = sqrt(pow(2, c.k) * Z_L);
}

static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct Job {
    Config c; std::vector<uint64_t> L; double norm; bg_projector G, H; uint64_t seed;
    int gpus, device0;
    uint8_t nccl_id[128];
    double numerator = 0, denominator = 0;
    std::string error;
};

static void worker(Job* job, int rank) {
    bg_ctx* ctx = nullptr;
    auto bail = [&](const char* what) {
        if (rank == 0 || job->error.empty()) job->error = std::string(what) + ": " + bg_last_error(ctx);
        if (ctx) bg_shutdown(ctx);
    };
    if (bg_init(&ctx, job->device0 + rank)) return bail("bg_init");
    if (bg_set_shard(ctx, rank, job->gpus)) return bail("bg_set_shard");
    if (job->gpus > 1 && bg_nccl_join(ctx, job->nccl_id)) return bail("bg_nccl_join");
    const Config& c = job->c;
    double num = 0, den = 0;
    if (c.t > 0 && bg_set_decomposition(ctx, c.t, c.exact, c.exact ? 0 : c.k, job->L.data())) return bail("bg_set_decomposition");
    int rc;
    if (c.noapprox == 0) {            // multiSampledProjector x2 (probability.c:197-198)
        rc = bg_sampled_norm(ctx, &job->G, (uint64_t)c.samples, c.bins, splitmix64(job->seed), job->norm, &num);
        if (!rc) rc = bg_sampled_norm(ctx, &job->H, (uint64_t)c.samples, c.bins, splitmix64(job->seed + 1), job->norm, &den);
    } else {                          // exactProjector x2 (probability.c:200-201)
        rc = bg_exact_norm(ctx, &job->G, job->norm, &num);
        if (!rc) rc = bg_exact_norm(ctx, &job->H, job->norm, &den);
    }
    if (rc) return bail("norm evaluation");
    if (rank == 0) { job->numerator = num; job->denominator = den; }
    bg_shutdown(ctx);
}

int main(int argc, char* argv[]) {
    const bool chatter = getenv("BG_QUIETER") == nullptr;
    if (chatter) printf("B200 backend (libbgnorm) print mode is on.\n");

    char file[256] = "";
    if (argc == 1) { if (scanf("%255s", file) != 1) file[0] = 0; }
    else { strncpy(file, argv[1], 255); file[255] = 0; }
    FILE* stream;
    if (strlen(file) == 0 || strcmp(file, "stdin") == 0) {
        if (chatter) printf("Reading arguments from stdin\n");
        stream = stdin;
    } else {
        if (chatter) printf("Reading arguments from file: %s\n", file);
        stream = fopen(file, "r");
        if (!stream) { printf("Error reading file.\n"); return 0; }
    }

    Job job;
    Config& c = job.c;
    bool ok = read_int(stream, &c.quiet) && read_int(stream, &c.verbose) && read_int(stream, &c.noapprox) &&
              read_int(stream, &c.samples) && read_int(stream, &c.bins) && read_int(stream, &c.t) &&
              read_int(stream, &c.k) && read_int(stream, &c.exact) && fscanf(stream, "%lf", &c.fidbound) == 1 &&
              read_int(stream, &c.fidelity) && read_int(stream, &c.rank) && read_int(stream, &c.forceL) &&
              read_int(stream, &c.forceSample);
    if (!ok) { printf("Error: truncated argument list.\n"); return 0; }
    if (chatter) printf("samples: %d bins: %d t: %d k: %d exact: %d noapprox: %d\n", c.samples, c.bins, c.t, c.k, c.exact, c.noapprox);
    std::string perr;
    if (!read_projector(stream, &job.G, &perr) || !read_projector(stream, &job.H, &perr)) {
        printf("Error: %s.\n", perr.c_str());
        return 0;
    }
    if (c.t > BG_MAX_T) { printf("Error: t = %d exceeds the 64-qubit limit of the packed layout.\n", c.t); return 0; }

    decompose(c, job.L, &job.norm);
    if (c.verbose) {
        if (c.exact) printf("Using exact decomposition of |H^t>: 2^%d\n", (c.t + 1) / 2);
        else printf("Stabilizer rank of |L>: 2^%d\n", c.k);
    }
    // more samples than terms: fall back to the exact norm (probability.c:162-174)
    if (c.noapprox == 0 && c.forceSample == 0) {
        const double terms = c.exact ? pow(2, (c.t + 1) / 2) : pow(2, c.k);
        if ((double)c.samples * c.bins * 2 > terms - 1) {
            c.noapprox = 1;
            if (c.verbose) printf("More samples than terms in exact calculation. Disabling sampling.\n");
        }
    }
    if (!c.exact && chatter) {
        printf("L:\n");
        for (int r = 0; r < c.k; r++) {
            printf(r == 0 ? "[[" : " [");
            for (int q = 0; q < c.t; q++) printf("%d", (int)((job.L[r] >> q) & 1));
            printf(r + 1 == c.k ? "]]\n" : "]\n");
        }
    }

    const char* es = getenv("BG_SEED");
    job.seed = es ? strtoull(es, nullptr, 0) : (uint64_t)getpid();
    const char* eg = getenv("BG_GPUS");
    job.gpus = eg ? atoi(eg) : 1;
    if (job.gpus < 1) job.gpus = 1;
    const char* ed = getenv("BG_DEVICE");
    job.device0 = ed ? atoi(ed) : 0;

    double numerator = 0, denominator = 0;
    if (c.t == 0) {
        // Clifford circuit, no magic states: closed form (innerprod.c:52-62, 157-167)
        for (int which = 0; which < 2; which++) {
            const bg_projector& P = which ? job.H : job.G;
            double val;
            if (P.nstabs == 0) val = pow(job.norm, 2);
            else {
                double sum = 1;
                for (int i = 0; i < P.nstabs; i++) { if (P.phase[i] == 0) sum += 1; if (P.phase[i] == 2) sum -= 1; }
                val = sum / (1 + (double)P.nstabs);
                if (c.noapprox == 0) val *= pow(job.norm, 2);
            }
            (which ? denominator : numerator) = val;
        }
    } else {
        if (job.gpus > 1 && bg_nccl_unique_id(job.nccl_id)) { printf("Error: %s\n", bg_last_error(nullptr)); return 0; }
        if (chatter) printf("World Size: %d\n", job.gpus);
        std::vector<std::thread> th;
        for (int r = 1; r < job.gpus; r++) th.emplace_back(worker, &job, r);
        worker(&job, 0);
        for (auto& x : th) x.join();
        if (!job.error.empty()) { printf("Error: %s\n", job.error.c_str()); return 0; }
        numerator = job.numerator; denominator = job.denominator;
    }

    const int sigfigs = 17;
    if (chatter) {
        printf("|| Gprime |H^t> ||^2 ~= %.*e\n", sigfigs, numerator);
        printf("|| Hprime |H^t> ||^2 ~= %.*e\n", sigfigs, denominator);
        if (denominator > 0) printf("Output: %.*e\n", sigfigs, numerator / denominator);
    }
    printf("%.*e\n", sigfigs, numerator);
    printf("%.*e\n", sigfigs, denominator);
    return 0;
}
