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
//              BG_DEVICE (first device, default 0), BG_QUIETER=1 (drop the chatter),
//              BG_SERVER=<socket> (forward the stream to a running `bgbackend --serve <socket>`).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <sys/socket.h>
#include <sys/un.h>

#include <condition_variable>
#include <mutex>
#include <functional>
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

// Z(L)'s weight histogram on the device (bg_decomposition_weights); returns "" on success
typedef std::function<std::string(int t, const std::vector<uint64_t>& L, uint64_t hist[65])> WeightFn;

// decompose (libcirc/probability.c:307-415): choose |H^t> (exact) or |L> with a random k x t L.
// Returns "" or an error string (only the device histogram of the fidelity loop can fail).
static std::string decompose(Config& c, std::vector<uint64_t>& L, double* norm, FILE* out, const WeightFn& weights) {
    const int t = c.t;
    if (t == 0) { c.exact = 0; *norm = 1; return ""; }
    const double v = cos(M_PI / 8);
    *norm = pow(2, floor((float)t / 2) / 2);
    if (t % 2) *norm *= 2 * v;
    if (c.exact) return "";
    const bool forceK = c.k > 0;
    if (!forceK) {
        c.k = (int)ceil(1 - 2 * t * log2(v) - log2(c.fidbound));
        if (c.verbose) fprintf(out, "Autopicking k = %d to achieve delta = %f.\n", c.k, c.fidbound);
    }
    if (c.k > t / 2 && !forceK && !c.forceL) {
        if (c.verbose) fprintf(out, "k > t/2. Reverting to exact decomposition.\n");
        c.exact = 1;
        return "";
    }
    if (c.k > t) {
        if (forceK && !c.quiet) fprintf(out, "Can't have k > t. Setting k to %d.\n", t);
        c.k = t;
    }
    double overlap = 0, Z_L = 0;
    while (overlap < 1 - c.fidbound || forceK) {
        random_L(c.k, t, L);
        if (c.rank && f2_rank(L) < c.k) {
            if (!c.quiet) fprintf(out, "L has insufficient rank. Sampling again...\n");
            continue;
        }
        if (!c.fidelity) break;
        // Z(L) = sum_x 2^{-|x|/2} over the 2^k combinations of rows (probability.c:373-391).  The 2^k
        // Hamming weights are counted on the device; the reference evaluates pow(2, -hamming/2) with
        // INTEGER division (:389), kept as is.  Every term is a power of two >= 2^-(t/2) and the sum stays
        // below 2^k, so whenever k + t/2 <= 53 every partial sum is exact in fp64 and the order of
        // summation is immaterial: the same bits as the reference's loop over i.
        uint64_t hist[65];
        const std::string werr = weights(t, L, hist);
        if (!werr.empty()) return werr;
        Z_L = 0;
        for (int w = 64; w >= 0; w--) Z_L += (double)hist[w] * pow(2, -w / 2);
        overlap = pow(2, c.k) * pow(v, 2 * t) / Z_L;
        if (forceK) { fprintf(out, "delta = 1 - <H^t|L>: %lf\n", 1 - overlap); break; }
        if (overlap < 1 - c.fidbound) { if (!c.quiet) fprintf(out, "delta = 1 - <H^t|L>: %lf - Not good enough!\n", 1 - overlap); }
        else if (!c.quiet) fprintf(out, "delta = 1 - <H^t|L>: %lf\n", 1 - overlap);
    }
    if (c.fidelity) *norm = sqrt(pow(2, c.k) * Z_L);
    return "";
}

static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct Job {
    enum Kind { EVALUATE = 0, WEIGHTS = 1 };
    Kind kind = EVALUATE;
    uint64_t hist[65];              // WEIGHTS: out
    Config c; std::vector<uint64_t> L; double norm; bg_projector G, H; uint64_t seed;
    double numerator = 0, denominator = 0;
    std::string error;
};

// One context per GPU, each owned by a persistent host thread (the reference's MPI workers,
// probability.c:221-299, without the processes): created once, reused by every job — in server mode
// by every probability() call of the front end.
class Engine {
public:
    // use_nccl: combine the ranks' partial sums with the in-library NCCL all-reduce (the persistent server: the
    // communicator is created once) — or, one-shot, on the host: a probability() evaluation has 16 bytes per projector
    // to combine and creating a communicator for 8 ranks costs ~10 s (measured), a host-side add costs nothing.
    Engine(int gpus, int device0, bool use_nccl) : gpus_(gpus), device0_(device0), use_nccl_(use_nccl && gpus > 1), parts_(gpus) {}
    ~Engine() { shutdown(); }

    // returns "" on success
    std::string start() {
        if (started_) return "";
        int ndev = 0;
        if (bg_device_count(&ndev)) return std::string("bg_device_count: ") + bg_last_error(nullptr);
        if (device0_ < 0 || device0_ + gpus_ > ndev) {
            char buf[160];
            snprintf(buf, sizeof buf, "BG_GPUS=%d from device %d, but %d CUDA device(s) are visible", gpus_, device0_, ndev);
            return buf;
        }
        if (use_nccl_ && bg_nccl_unique_id(nccl_id_)) return std::string("bg_nccl_unique_id: ") + bg_last_error(nullptr);
        ready_ = 0;
        for (int r = 0; r < gpus_; r++) th_.emplace_back(&Engine::worker, this, r);
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return ready_ == gpus_; });
        started_ = true;
        return init_error_;
    }

    void run(Job* job) {
        std::unique_lock<std::mutex> lk(mu_);
        job_ = job; done_ = 0; generation_++;
        cv_job_.notify_all();
        cv_done_.wait(lk, [&] { return done_ == gpus_; });
        job_ = nullptr;
        if (!use_nccl_ && gpus_ > 1 && job->kind == Job::EVALUATE && job->error.empty()) {      // host-side reduction of the ranks' parts
            double a[4] = {0, 0, 0, 0};
            for (int r = 0; r < gpus_; r++) for (int j = 0; j < 4; j++) a[j] += parts_[r].v[j];
            if (job->c.noapprox == 0) { job->numerator = a[0]; job->denominator = a[2]; }
            else { job->numerator = sqrt(a[0] * a[0] + a[1] * a[1]); job->denominator = sqrt(a[2] * a[2] + a[3] * a[3]); }
        }
    }

    void shutdown() {
        if (th_.empty()) return;
        { std::unique_lock<std::mutex> lk(mu_); stop_ = true; generation_++; cv_job_.notify_all(); }
        for (auto& t : th_) t.join();
        th_.clear();
    }
    int gpus() const { return gpus_; }

private:
    // All ranks report whether they are fine and learn whether EVERY rank is: no rank enters a collective
    // (ncclCommInitRank, the all-reduce of a job) that a failed peer would never join.
    bool agree(bool ok) {
        std::unique_lock<std::mutex> lk(mu_);
        const unsigned long long round = agree_round_;
        if (!ok) agree_bad_ = true;
        if (++agree_count_ == gpus_) {
            agree_result_ = !agree_bad_;
            agree_count_ = 0; agree_bad_ = false; agree_round_++;
            cv_agree_.notify_all();
        } else {
            cv_agree_.wait(lk, [&] { return agree_round_ != round; });
        }
        return agree_result_;
    }

    void worker(int rank) {
        bg_ctx* ctx = nullptr;
        std::string err;
        // phase 1: the context; phase 2 (only if every rank has one): the NCCL communicator
        if (bg_init(&ctx, device0_ + rank)) err = std::string("bg_init: ") + bg_last_error(nullptr);
        else if (bg_set_shard(ctx, rank, gpus_)) err = std::string("bg_set_shard: ") + bg_last_error(ctx);
        const bool all_up = agree(err.empty());
        if (all_up && use_nccl_ && bg_nccl_join(ctx, nccl_id_)) err = std::string("bg_nccl_join: ") + bg_last_error(ctx);
        if (all_up && !use_nccl_ && gpus_ > 1 && err.empty() && bg_set_allreduce(ctx, 0)) err = std::string("bg_set_allreduce: ") + bg_last_error(ctx);
        if (!all_up && err.empty()) err = "another GPU of the job failed to initialise";
        unsigned long long seen = 0;
        {
            std::unique_lock<std::mutex> lk(mu_);
            if (!err.empty() && (init_error_.empty() || init_error_[0] == 'a')) init_error_ = err;
            ready_++;
            cv_done_.notify_all();
        }
        while (true) {
            Job* job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_job_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_) break;
                job = job_;
            }
            std::string jerr = err;
            double num = 0, den = 0;
            if (job->kind == Job::WEIGHTS) {       // decompose()'s fidelity loop: one GPU is plenty, no collective
                if (jerr.empty() && rank == 0 && bg_decomposition_weights(ctx, job->c.t, (int)job->L.size(), job->L.data(), job->hist))
                    jerr = bg_last_error(ctx);
            } else {
                const Config& c = job->c;
                // everything that can fail on one rank alone (validation, allocation, uploads) comes first ...
                int rc = jerr.empty() ? bg_set_decomposition(ctx, c.t, c.exact, c.exact ? 0 : c.k, job->L.data()) : 1;
                // the split-phase job (prepare, then run + finish): both projectors in one launch sequence.  Without a
                // communicator only bins == 1 is additive over the ranks (a median of partial bin sums is not).
                const bool split = c.noapprox == 0 && c.bins >= 1 && c.bins <= 4 && job->G.nstabs > 0 && job->H.nstabs > 0 &&
                                   job->G.nqubits > 0 && job->H.nqubits > 0 && (use_nccl_ || gpus_ == 1 || c.bins == 1);
                if (!rc && split)
                    rc = bg_sampled_prepare2(ctx, &job->G, &job->H, (uint64_t)c.samples, c.bins, splitmix64(job->seed),
                                             splitmix64(job->seed + 1));
                if (rc && jerr.empty()) jerr = bg_last_error(ctx);
                // ... then the ranks agree, and only then run what contains the all-reduce
                if (agree(jerr.empty())) {
                    double part[4] = {0, 0, 0, 0};                   // this rank's (numerator re, im, denominator re, im)
                    if (c.noapprox == 0) {            // multiSampledProjector x2 (probability.c:197-198)
                        double out[2] = {0, 0};
                        if (split) { rc = bg_sampled_run(ctx); if (!rc) rc = bg_sampled_finish2(ctx, job->norm, out); }
                        else if (use_nccl_ || gpus_ == 1)
                            rc = bg_sampled_norm2(ctx, &job->G, &job->H, (uint64_t)c.samples, c.bins, splitmix64(job->seed),
                                                  splitmix64(job->seed + 1), job->norm, out);
                        else if (rank == 0) {         // closed forms / many bins without a communicator: rank 0 alone, unsharded
                            rc = bg_set_shard(ctx, 0, 1);
                            if (!rc) rc = bg_sampled_norm2(ctx, &job->G, &job->H, (uint64_t)c.samples, c.bins, splitmix64(job->seed),
                                                           splitmix64(job->seed + 1), job->norm, out);
                            if (bg_set_shard(ctx, 0, gpus_) && !rc) rc = 1;
                            if (!rc) rc = bg_set_allreduce(ctx, 0);
                        }
                        num = out[0]; den = out[1];
                        part[0] = out[0]; part[2] = out[1];           // without the all-reduce: this rank's share of the mean
                    } else if (use_nccl_ || gpus_ == 1) {             // exactProjector x2 (probability.c:200-201)
                        rc = bg_exact_norm(ctx, &job->G, job->norm, &num);
                        if (!rc) rc = bg_exact_norm(ctx, &job->H, job->norm, &den);
                    } else {
                        rc = bg_exact_norm_parts(ctx, &job->G, job->norm, &part[0]);
                        if (!rc) rc = bg_exact_norm_parts(ctx, &job->H, job->norm, &part[2]);
                    }
                    if (rc) jerr = bg_last_error(ctx);
                    for (int j = 0; j < 4; j++) parts_[rank].v[j] = part[j];
                } else if (jerr.empty()) jerr = "another GPU of the job reported an error";
            }
            {
                std::unique_lock<std::mutex> lk(mu_);
                if (!jerr.empty() && (job->error.empty() || job->error[0] == 'a')) job->error = jerr;
                if (rank == 0) { job->numerator = num; job->denominator = den; }
                done_++;
                cv_done_.notify_all();
            }
        }
        if (ctx) bg_shutdown(ctx);
    }

    struct Part { double v[4] = {0, 0, 0, 0}; };
    int gpus_, device0_;
    bool use_nccl_;
    std::vector<Part> parts_;
    uint8_t nccl_id_[128];
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_job_, cv_done_, cv_agree_;
    int agree_count_ = 0; bool agree_bad_ = false, agree_result_ = true; unsigned long long agree_round_ = 0;
    Job* job_ = nullptr;
    unsigned long long generation_ = 0;
    int done_ = 0, ready_ = 0;
    bool stop_ = false, started_ = false;
    std::string init_error_;
};

// The 13 scalars and two projectors of one instruction stream (probability.c:74-127).  Reads token by
// token and never waits for end-of-file: the front end keeps the pipe open (probability.py:283-284).
static bool parse_job(FILE* stream, Job& job, std::string* err) {
    Config& c = job.c;
    bool ok = read_int(stream, &c.quiet) && read_int(stream, &c.verbose) && read_int(stream, &c.noapprox) &&
              read_int(stream, &c.samples) && read_int(stream, &c.bins) && read_int(stream, &c.t) &&
              read_int(stream, &c.k) && read_int(stream, &c.exact) && fscanf(stream, "%lf", &c.fidbound) == 1 &&
              read_int(stream, &c.fidelity) && read_int(stream, &c.rank) && read_int(stream, &c.forceL) &&
              read_int(stream, &c.forceSample);
    if (!ok) { *err = "truncated argument list"; return false; }
    if (!read_projector(stream, &job.G, err) || !read_projector(stream, &job.H, err)) return false;
    return true;
}

// the same stream, re-serialised (what a client forwards to the persistent server)
static std::string serialize_job(const Job& job) {
    const Config& c = job.c;
    char buf[256];
    snprintf(buf, sizeof buf, "%d %d %d %d %d %d %d %d %.17g %d %d %d %d\n", c.quiet, c.verbose, c.noapprox, c.samples,
             c.bins, c.t, c.k, c.exact, c.fidbound, c.fidelity, c.rank, c.forceL, c.forceSample);
    std::string out = buf;
    for (const bg_projector* P : {&job.G, &job.H}) {
        out += std::to_string(P->nstabs) + " " + std::to_string(P->nstabs ? P->nqubits : 0) + "\n";
        for (int i = 0; i < P->nstabs; i++) {
            out += std::to_string((int)P->phase[i]);
            for (int q = 0; q < P->nqubits; q++) {
                out += ((P->xs[i] >> q) & 1) ? " 1" : " 0";
                out += ((P->zs[i] >> q) & 1) ? " 1" : " 0";
            }
            out += "\n";
        }
    }
    return out;
}

// master() (probability.c:42-216) on one parsed instruction stream: decompose, evaluate, print.
static void process(Job& job, FILE* out, Engine& engine, bool chatter, uint64_t seed) {
    srand(1);        // decompose() draws L from libc rand() in a fresh process (default seed), probability.c:153,182
    Config& c = job.c;
    if (chatter) fprintf(out, "samples: %d bins: %d t: %d k: %d exact: %d noapprox: %d\n", c.samples, c.bins, c.t, c.k, c.exact, c.noapprox);
    if (c.t > BG_MAX_T) { fprintf(out, "Error: t = %d exceeds the 64-qubit limit of the packed layout.\n", c.t); return; }

    const WeightFn weights = [&](int t, const std::vector<uint64_t>& L, uint64_t hist[65]) -> std::string {
        std::string err = engine.start();
        if (!err.empty()) return err;
        Job wj;
        wj.kind = Job::WEIGHTS; wj.c.t = t; wj.L = L;
        engine.run(&wj);
        if (!wj.error.empty()) return wj.error;
        memcpy(hist, wj.hist, sizeof wj.hist);
        return "";
    };
    const std::string derr = decompose(c, job.L, &job.norm, out, weights);
    if (!derr.empty()) { fprintf(out, "Error: %s\n", derr.c_str()); return; }
    if (c.verbose) {
        if (c.exact) fprintf(out, "Using exact decomposition of |H^t>: 2^%d\n", (c.t + 1) / 2);
        else fprintf(out, "Stabilizer rank of |L>: 2^%d\n", c.k);
    }
    // more samples than terms: fall back to the exact norm (probability.c:162-174)
    if (c.noapprox == 0 && c.forceSample == 0) {
        const double terms = c.exact ? pow(2, (c.t + 1) / 2) : pow(2, c.k);
        if ((double)c.samples * c.bins * 2 > terms - 1) {
            c.noapprox = 1;
            if (c.verbose) fprintf(out, "More samples than terms in exact calculation. Disabling sampling.\n");
        }
    }
    if (!c.exact && chatter) {
        fprintf(out, "L:\n");
        for (int r = 0; r < c.k; r++) {
            fprintf(out, r == 0 ? "[[" : " [");
            for (int q = 0; q < c.t; q++) fprintf(out, "%d", (int)((job.L[r] >> q) & 1));
            fprintf(out, r + 1 == c.k ? "]]\n" : "]\n");
        }
    }
    job.seed = seed;

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
        std::string err = engine.start();
        if (!err.empty()) { fprintf(out, "Error: %s\n", err.c_str()); return; }
        if (chatter) fprintf(out, "World Size: %d\n", engine.gpus());
        engine.run(&job);
        if (!job.error.empty()) { fprintf(out, "Error: %s\n", job.error.c_str()); return; }
        numerator = job.numerator; denominator = job.denominator;
    }

    const int sigfigs = 17;
    if (chatter) {
        fprintf(out, "|| Gprime |H^t> ||^2 ~= %.*e\n", sigfigs, numerator);
        fprintf(out, "|| Hprime |H^t> ||^2 ~= %.*e\n", sigfigs, denominator);
        if (denominator > 0) fprintf(out, "Output: %.*e\n", sigfigs, numerator / denominator);
    }
    fprintf(out, "%.*e\n", sigfigs, numerator);
    fprintf(out, "%.*e\n", sigfigs, denominator);
}

static bool read_all(int fd, std::string* out) {
    char buf[65536];
    while (true) {
        ssize_t n = read(fd, buf, sizeof buf);
        if (n < 0) { if (errno == EINTR) continue; return false; }
        if (n == 0) return true;
        out->append(buf, (size_t)n);
    }
}
static bool write_all(int fd, const char* p, size_t n) {
    while (n) {
        ssize_t w = write(fd, p, n);
        if (w < 0) { if (errno == EINTR) continue; return false; }
        p += w; n -= (size_t)w;
    }
    return true;
}

// Server mode: the front end starts a fresh back-end process for every
// probability() call (probability.py:244); CUDA + NCCL start-up would then dominate.  `bgbackend --serve
// <socket>` keeps the contexts alive; a `bgbackend` started with BG_SERVER=<socket> only forwards its
// instruction stream and relays the answer, so the drop-in protocol is unchanged.
static int serve(const char* path, int gpus, int device0, bool chatter) {
    const char* red = getenv("BG_REDUCE");          // the server keeps an NCCL communicator unless BG_REDUCE=host
    Engine engine(gpus, device0, !(red && strcmp(red, "host") == 0));
    std::string err = engine.start();
    if (!err.empty()) { fprintf(stderr, "bgbackend --serve: %s\n", err.c_str()); return 1; }
    signal(SIGPIPE, SIG_IGN);                      // a client that goes away before reading its answer must not kill the server
    int srv = socket(AF_UNIX, SOCK_STREAM, 0);
    if (srv < 0) { perror("socket"); return 1; }
    sockaddr_un addr; memset(&addr, 0, sizeof addr);
    addr.sun_family = AF_UNIX;
    if (strlen(path) >= sizeof(addr.sun_path)) { fprintf(stderr, "bgbackend --serve: socket path too long (%zu bytes max)\n", sizeof(addr.sun_path) - 1); return 1; }
    strcpy(addr.sun_path, path);
    unlink(path);
    const mode_t old_umask = umask(0077);          // the socket is for this user only
    const bool bound = bind(srv, (sockaddr*)&addr, sizeof addr) == 0 && listen(srv, 16) == 0;
    umask(old_umask);
    if (!bound) { perror("bind/listen"); return 1; }
    printf("bgbackend serving on %s with %d GPU(s)\n", path, gpus);
    fflush(stdout);
    uint64_t calls = 0;
    const char* es = getenv("BG_SEED");
    const uint64_t seed0 = es ? strtoull(es, nullptr, 0) : (uint64_t)getpid();
    while (true) {
        int fd = accept(srv, nullptr, nullptr);
        if (fd < 0) { if (errno == EINTR) continue; break; }
        timeval tv; tv.tv_sec = 30; tv.tv_usec = 0;    // a stalled client cannot block the calls behind it for ever
        setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
        setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof tv);
        std::string in;
        if (read_all(fd, &in)) {
            if (in == "shutdown\n") { close(fd); break; }
            FILE* fin = fmemopen((void*)in.data(), in.size(), "r");
            char* obuf = nullptr; size_t olen = 0;
            FILE* fout = open_memstream(&obuf, &olen);
            if (chatter) fprintf(fout, "B200 backend (libbgnorm, persistent server) print mode is on.\n");
            Job job; std::string perr;
            if (!parse_job(fin, job, &perr)) fprintf(fout, "Error: %s.\n", perr.c_str());
            else process(job, fout, engine, chatter, seed0 + 2 * calls++);
            fclose(fin); fclose(fout);
            write_all(fd, obuf, olen);
            free(obuf);
        }
        close(fd);
    }
    close(srv);
    unlink(path);
    return 0;
}

static bool try_client(const char* path, const std::string& text) {
    int fd = socket(AF_UNIX, SOCK_STREAM, 0);
    if (fd < 0) return false;
    sockaddr_un addr; memset(&addr, 0, sizeof addr);
    addr.sun_family = AF_UNIX;
    strncpy(addr.sun_path, path, sizeof(addr.sun_path) - 1);
    if (connect(fd, (sockaddr*)&addr, sizeof addr) < 0) { close(fd); return false; }
    bool ok = write_all(fd, text.data(), text.size());
    shutdown(fd, SHUT_WR);
    std::string out;
    ok = ok && read_all(fd, &out);
    close(fd);
    if (!ok || out.empty()) return false;
    fwrite(out.data(), 1, out.size(), stdout);
    return true;
}

int main(int argc, char* argv[]) {
    const bool chatter = getenv("BG_QUIETER") == nullptr;
    const char* eg = getenv("BG_GPUS");
    int gpus = eg ? atoi(eg) : 1;
    if (gpus < 1) gpus = 1;
    const char* ed = getenv("BG_DEVICE");
    const int device0 = ed ? atoi(ed) : 0;
    if (argc >= 3 && strcmp(argv[1], "--serve") == 0) return serve(argv[2], gpus, device0, chatter);

    // argv[1] = file name or "stdin"; without it the first stdin token is the file name (probability.c:52-68)
    char file[256] = "";
    if (argc == 1) { if (scanf("%255s", file) != 1) file[0] = 0; }
    else { strncpy(file, argv[1], 255); file[255] = 0; }
    const bool from_stdin = strlen(file) == 0 || strcmp(file, "stdin") == 0;
    FILE* stream = stdin;
    if (!from_stdin) {
        stream = fopen(file, "r");
        if (!stream) { if (chatter) printf("Reading arguments from file: %s\n", file); printf("Error reading file.\n"); return 0; }
    }
    Job job; std::string perr;
    if (!parse_job(stream, job, &perr)) { printf("Error: %s.\n", perr.c_str()); return 0; }
    if (const char* srv = getenv("BG_SERVER")) {
        if (try_client(srv, serialize_job(job))) return 0;          // answered by the persistent server
    }
    if (chatter) {
        printf("B200 backend (libbgnorm) print mode is on.\n");
        if (from_stdin) printf("Reading arguments from stdin\n");
        else printf("Reading arguments from file: %s\n", file);
    }
    const char* es = getenv("BG_SEED");
    const uint64_t seed = es ? strtoull(es, nullptr, 0) : (uint64_t)getpid();
    const char* er = getenv("BG_REDUCE");                  // one-shot: the ranks' parts are added on the host unless BG_REDUCE=nccl
    Engine engine(gpus, device0, er && strcmp(er, "nccl") == 0);
    process(job, stdout, engine, chatter, seed);
    fflush(stdout);
    engine.shutdown();
    return 0;
}
