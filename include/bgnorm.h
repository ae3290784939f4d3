/* bgnorm.h — C ABI of the B200-native stabilizer-rank norm estimator.
 *
 * This is the drop-in boundary for ONE path of patrickrall/CircuitSimulator:
 * the L x chi double loop that sums stabilizer inner products (Bravyi-Gosset,
 * arXiv:1601.07601).  Every entry point names the reference interface it
 * replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, caller-owned buffers, no exceptions.
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from bg_last_error().
 *   - one host thread per context.  A context owns one CUDA device, its stream,
 *     its device buffers and (when world > 1) one NCCL communicator.
 *   - there is NO CPU fallback: if no CUDA device is usable bg_init fails.
 *
 * Packed layout
 *   Bit q of a 64-bit word is qubit / coordinate q (LSB = index 0).  The
 *   reference's BitVector/BitMatrix are byte arrays, MSB-first, with no row
 *   padding (libcirc/utils/matrix.c:124-131, 330-339); the *_bitmatrix entry
 *   points take those byte arrays directly so the reference host can pass
 *   `mat->data` unchanged.
 */
#ifndef BGNORM_H
#define BGNORM_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BG_MAX_T 64

/* Replaces struct StabilizerState (libcirc/stabilizer/stabilizer.h:5-16).
 * |K,q> = 2^{-k/2} sum_{x in K} e^{i pi q(x)/4} |x>,  K = h + span(G[0..k)),
 * q(x) = Q + sum_a D_a x_a + sum_{a<b} J_ab x_a x_b in G-coordinates,
 * D_a = 2*D1_a + 4*D2_a  (getD/setD, stabilizer.c:55-62),
 * J_ab in {0,4} stored as one bit, symmetric, J_aa = D1_a.
 * Rows >= k of G/Gbar complete G to an invertible matrix, Gbar = (G^-1)^T. */
typedef struct bg_state {
    int32_t  n;          /* state width = t of the circuit, 1..64          */
    int32_t  k;          /* dimension of the affine space, 0..n            */
    int32_t  Q;          /* Z8                                             */
    int32_t  reserved;
    uint64_t h;          /* shift vector                                   */
    uint64_t D1;         /* bit a = D1_a                                   */
    uint64_t D2;         /* bit a = D2_a                                   */
    uint64_t G[BG_MAX_T];     /* row a, bit q = G[a][q]                    */
    uint64_t Gbar[BG_MAX_T];
    uint64_t J[BG_MAX_T];     /* row a, bit b = J[a][b]/4                  */
} bg_state;

/* Replaces struct Projector (libcirc/utils/comms.h:4-11): nstabs generators
 * i^phase * Z(zeta) X(xi) on nqubits (= t) qubits, phase = 2*sign + complex. */
typedef struct bg_projector {
    int32_t  nstabs;
    int32_t  nqubits;
    uint8_t  phase[BG_MAX_T * 2];   /* up to 128 generators                */
    uint64_t xs[BG_MAX_T * 2];      /* xi   of generator i                 */
    uint64_t zs[BG_MAX_T * 2];      /* zeta of generator i                 */
} bg_projector;
#define BG_MAX_STABS (BG_MAX_T * 2)

typedef struct bg_ctx bg_ctx;

/* ---- lifetime ---------------------------------------------------------- */

/* Create a context on CUDA device `device`.  Replaces MPI_Init + the worker
 * "init" command (libcirc/probability.c:31-37, 184-191, 248-257). */
int  bg_init(bg_ctx** out, int device);
int  bg_device_count(int* out);                 /* visible sm_100 CUDA devices (checked before a multi-GPU start) */
void bg_shutdown(bg_ctx* ctx);
const char* bg_last_error(const bg_ctx* ctx);   /* ctx may be NULL: last global error */

/* Sample sharding.  Rank r of `world` evaluates samples l with l % world == r
 * — the reference's rank stride (libcirc/probability.c:268,281;
 * innerprod.c:75,187).  Default: rank 0 of 1. */
int  bg_set_shard(bg_ctx* ctx, int rank, int world);

/* NCCL plumbing for world > 1 (replaces recvDouble/recvComplex gathers,
 * libcirc/innerprod.c:79-81,192-195).  Rank 0 fills a 128-byte unique id, the
 * caller distributes it by any means, every rank then joins.  After joining,
 * bg_sampled_norm / bg_exact_norm all-reduce their partial sums in-library. */
/* With all-reduce disabled (enabled = 0) and world > 1, bg_sampled_norm returns this rank's
 * PARTIAL result, (sum over its samples)/samples, so that the caller can reduce by other means
 * (bins must be 1); bg_exact_norm refuses.  Default: enabled. */
int  bg_set_allreduce(bg_ctx* ctx, int enabled);
int  bg_nccl_unique_id(uint8_t id[128]);
int  bg_nccl_join(bg_ctx* ctx, const uint8_t id[128]);   /* uses bg_set_shard's rank/world */

/* ---- inputs ------------------------------------------------------------ */

/* The magic-state decomposition chosen by decompose()
 * (libcirc/probability.c:307-415): exact != 0 -> |H^t> as 2^ceil(t/2) terms
 * (prepH, libcirc/stateprep.c:36-81); exact == 0 -> |L> with the k x t matrix L
 * (prepL, stateprep.c:85-120), one packed row per uint64. */
int  bg_set_decomposition(bg_ctx* ctx, int t, int exact, int k, const uint64_t* L_rows);
/* Same, with L in the reference's BitMatrix byte layout (k*t bits, MSB-first). */
int  bg_set_decomposition_bitmatrix(bg_ctx* ctx, int t, int exact, int k, const uint8_t* L_bits);

/* Build a bg_projector from the reference's Projector members
 * (phaseSign->data, phaseComplex->data, xs->data, zs->data). */
int  bg_projector_from_bitmatrix(bg_projector* out, int nstabs, int nqubits,
                                 const uint8_t* phase_sign, const uint8_t* phase_complex,
                                 const uint8_t* xs, const uint8_t* zs);

/* decompose()'s fidelity loop (libcirc/probability.c:373-391), which the
 * reference runs on the host over all 2^k combinations of the rows of L:
 * hist[w] = #{ i < 2^k : |x~_i| = w }, w = 0..64, x~_i as in prepL.  From it
 * Z(L) = sum_w hist[w] * 2^-(w/2) (INTEGER w/2, as the reference's
 * pow(2, -hamming/2) evaluates) and <H^t|L> = 2^k cos(pi/8)^2t / Z(L).  Does
 * not touch the context's current decomposition. */
int  bg_decomposition_weights(bg_ctx* ctx, int t, int k, const uint64_t* L_rows, uint64_t hist[65]);

/* ---- the hot path ------------------------------------------------------ */

/* Replaces multiSampledProjector/sampledProjector/singleProjectorSample
 * (libcirc/innerprod.c:23-144): median over `bins` of the mean over `samples`
 * of 2^t |<theta| P |decomposition>|^2, theta ~ randomStabilizerState(t)
 * (stabilizer.c:689-756) drawn on the device from Philox4x32-10 keyed by
 * (seed, bin, sample index).  Empty projector -> norm^2 and t == 0 closed form
 * as in innerprod.c:47-62. */
int  bg_sampled_norm(bg_ctx* ctx, const bg_projector* P, uint64_t samples, int bins,
                     uint64_t seed, double norm, double* out);

/* Replaces exactProjector/exactProjectorWork (libcirc/innerprod.c:148-261):
 * | sum_{i<=j} c_ij <P phi_i | phi_j> |, c = 1 on the diagonal, 2 Re off it. */
int  bg_exact_norm(bg_ctx* ctx, const bg_projector* P, double norm, double* out);
/* The complex sum before the magnitude is taken (the value a rank sends with sendComplex, innerprod.c:192-195):
 * out[0] = Re, out[1] = Im of THIS rank's part when the in-library all-reduce is off (bg_set_allreduce(ctx, 0)) — a
 * host that reduces the ranks itself adds the parts and takes sqrt(re^2 + im^2) — and of the whole sum otherwise.
 * Closed forms (empty projector, t == 0) are returned in out[0] by rank 0 and as 0 by the other ranks. */
int  bg_exact_norm_parts(bg_ctx* ctx, const bg_projector* P, double norm, double out[2]);

/* ---- parity / debug entry points --------------------------------------- */

/* innerProductExact (libcirc/stabilizer/stabilizer.c:589-659) on n_pairs
 * independent pairs: epm[3*i..] = (eps, p, m mod 8) of <b_i|a_i> as the
 * reference returns it for innerProductExact(state1 = a_i, state2 = b_i). */
int  bg_inner_products(bg_ctx* ctx, size_t n_pairs, const bg_state* a, const bg_state* b,
                       int32_t* epm);

/* The L x chi loop on host-supplied theta states (libcirc/innerprod.c:100-142):
 * if `project` != 0 each theta is first projected by P with measurePauli on
 * the device.  Outputs (any may be NULL): epm[(l*chi + i)*3..] per pair,
 * per_sample[l] = 2^t |projfactor * sum_i <theta_l|phi_i>|^2, *mean = their
 * mean over n_states in index order. */
int  bg_sampled_norm_from_states(bg_ctx* ctx, const bg_projector* P, int project,
                                 size_t n_states, const bg_state* thetas,
                                 int32_t* epm, double* per_sample, double* mean);

/* measurePauli (libcirc/stabilizer/stabilizer.c:827-959) applied in place to
 * n_states states, one generator i^m Z(zeta) X(xi) each.  result[i] is the
 * reference's return value (0, 1 or 2^-1/2). */
int  bg_measure_pauli(bg_ctx* ctx, size_t n_states, bg_state* states, const int32_t* m,
                      const uint64_t* zeta, const uint64_t* xi, double* result);

/* The device RNG's randomStabilizerState(t): states for sample indices
 * [first, first+count) of bin `bin` under `seed` (exactly those that
 * bg_sampled_norm draws). */
int  bg_random_states(bg_ctx* ctx, int t, uint64_t seed, int bin, uint64_t first, size_t count,
                      bg_state* out);

/* The decomposition terms phi_i, i in [first, first+count), as full states
 * (prepH / prepL, libcirc/stateprep.c:36-120). */
int  bg_decomposition_terms(bg_ctx* ctx, uint64_t first, size_t count, bg_state* out);

/* ---- measurement ------------------------------------------------------- */

typedef struct bg_stats {
    double   kernel_ms;        /* CUDA-event time of the last hot-path kernel(s), on ctx's stream */
    uint64_t pairs;            /* inner products evaluated by this rank in the last call          */
    uint64_t pair_launches;    /* launches of the pair kernel proper in the last call (1 for a fused
                                  two-projector job, else one per projector and bin)                   */
    uint64_t launches;         /* kernels launched by the last call                               */
    uint64_t h2d_bytes;        /* host->device bytes moved by the last call                       */
    uint64_t d2h_bytes;        /* device->host bytes moved by the last call                       */
    double   prepare_ms;       /* ... of which k_prepare (theta draw + projection + ambient form)  */
    double   pairs_ms;         /* ... of which the pair kernels (the L x chi loop proper)          */
    uint64_t overlapped;       /* 1: the job ran in overlap mode (few samples per GPU, see bg_sampled_run):
                                  consecutive jobs run on two streams, so prepare_ms / pairs_ms of one job
                                  include the time its kernels shared the GPU with the other job's      */
} bg_stats;
int  bg_get_stats(const bg_ctx* ctx, bg_stats* out);

/* Split-phase variant of bg_sampled_norm for benchmarking with inputs already
 * resident: upload (projector, decomposition) once, then run the kernel only.
 * bg_sampled_norm == bg_sampled_prepare + bg_sampled_run + bg_sampled_finish. */
int  bg_sampled_prepare(bg_ctx* ctx, const bg_projector* P, uint64_t samples, int bins, uint64_t seed);
int  bg_sampled_run(bg_ctx* ctx);                       /* async: kernels on ctx's stream (captured once into a CUDA
                                                           graph, then replayed), all-reduce + read-back on a side
                                                           stream.  Up to TWO runs may be in flight.  A fused
                                                           two-projector job with few samples per GPU (its draw +
                                                           projection kernel cannot fill the machine) runs in OVERLAP
                                                           mode: every second run goes to an internal stream with
                                                           buffers of its own, so that the draw + projection of run
                                                           i+1 shares the GPU with the pair kernel of run i
                                                           (BG_OVERLAP=0 / 1 forces it off / on).                 */
int  bg_sampled_finish(bg_ctx* ctx, double norm, double* out);   /* wait for the OLDEST run in flight     */

/* Numerator and denominator of one probability() evaluation together — both projectors against the
 * same decomposition (libcirc/probability.c:197-198) — with ONE all-reduce and one host sync.
 * out[0] = G', out[1] = H'.  bg_sampled_norm2 = bg_sampled_prepare2 + run + bg_sampled_finish2. */
int  bg_sampled_norm2(bg_ctx* ctx, const bg_projector* G, const bg_projector* H, uint64_t samples, int bins,
                      uint64_t seed_g, uint64_t seed_h, double norm, double out[2]);
/* Pipelining independent probability() evaluations (bins, circuits, the clients of a served back end): with a job in
 * flight, bg_sampled_prepare2 may be called again with NEW projectors for a job of the same shape (samples, bins,
 * seeds) — they are staged through alternating pinned buffers and uploaded behind the running job on its stream — and
 * bg_set_decomposition with a new L of the same (t, k) keeps the prepared job and its captured CUDA graphs.  Then
 * bg_sampled_run, and bg_sampled_finish2 for the OLDEST job.  A job of another shape drops what is in flight. */
int  bg_sampled_prepare2(bg_ctx* ctx, const bg_projector* G, const bg_projector* H, uint64_t samples, int bins,
                         uint64_t seed_g, uint64_t seed_h);
int  bg_sampled_finish2(bg_ctx* ctx, double norm, double out[2]);

/* Per-sample values of the most recent FINISHED device-RNG job (bg_sampled_norm, bg_sampled_norm2, or prepare /
 * run / finish): out[i] = 2^t |projfactor * sum_j <theta_l|phi_j>|^2 (libcirc/innerprod.c:142) for this rank's
 * local sample index first + i, i < count — global sample l = rank + (first + i) * world (bg_set_shard).
 * projector: 0, or 1 for H' of a two-projector job.  With bins > 1 only the last bin is kept.  This is the
 * read-back the parity tests use to spot-check a full-size run against the oracle on the same Philox theta
 * (BG_DUMP). */
int  bg_sampled_per_sample(bg_ctx* ctx, int projector, uint64_t first, size_t count, double* out);

/* Run on the caller's CUDA stream (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream)
 * instead of the context's own, so that the caller's events bracket the kernels. */
int  bg_set_stream(bg_ctx* ctx, void* cuda_stream);

/* Roofline denominator: measured integer-pipe throughput of this GPU, in 32-bit lane-ops per
 * second, for LOP3 (the ALU pipe the elimination runs on) and for POPC (+IADD). */
int  bg_measure_int_peak(bg_ctx* ctx, double* lop3_lane_ops_per_s, double* popc_lane_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* BGNORM_H */
