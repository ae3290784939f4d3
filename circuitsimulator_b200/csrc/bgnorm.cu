// bgnorm.cu — CUDA kernels (sm_100a) and the C ABI of include/bgnorm.h.
//
// Hot path = the L x chi loop of libcirc/innerprod.c:88-144 (reference repo
// patrickrall/CircuitSimulator).  Per probability() evaluation (both projectors in one job) one launch of:
//
//   k_prepare     one warp per sample: draw theta on the device (Philox) or load it, turn it into its
//                 ambient quadratic form + parity checks, apply the projector's generators in those
//                 coordinates (measurePauli as mask operations), store a 1 KB record in HBM.
//   k_pairs_tpp   persistent CTAs, one THREAD per inner product, a warp = one theta x 32 terms: the chi
//                 decomposition terms are staged ONCE per CTA into shared memory with a TMA bulk copy
//                 (cp.async.bulk + mbarrier); each thread eliminates its own copy of J (shared memory,
//                 [row][thread]) two steps per row pass (bg_tpp.cuh); sums accumulated exactly in
//                 Z[e^{i pi/4}] as int64.   (k_pairs: the same with one warp per inner product,
//                 ballot / shfl / LOP3 / POPC in registers — the first version, kept under BG_KERNEL=warp.)
//   k_finalize_*  2^t |projfactor * sum|^2 per sample in fp64 and a fixed-order tree sum.
//
// There is no CPU fallback anywhere in this file: every entry point either runs the kernels
// or fails with an error string.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "bgnorm.h"
#include "bg_device.cuh"
#include "bg_philox.cuh"
#include "bg_warp_ops.cuh"
#include "bg_tpp.cuh"
#include "bg_shb.cuh"
#include "bg_prep.cuh"

using namespace bg;

// CTA size of k_pairs_tpp, a compile-time constant: the byte stride between a thread's working rows
// (threads x word size) then folds into immediate offsets.  3 warps: 6 CTAs per SM at t = 40.
#ifndef BG_TPP_WARPS
#define BG_TPP_WARPS 3
#endif
#define BG_TPP_THREADS (32 * BG_TPP_WARPS)

// ------------------------------------------------------------------------------------------
// device-side records
// ------------------------------------------------------------------------------------------
// alive: 0 = annihilated by the projector; ROUTE_TPP / ROUTE_TPP_MANY = evaluated by k_pairs_tpp (one
// thread per inner product; _MANY: more than TPP_MAXC parity checks, pivot history in shared memory);
// ROUTE_WARP = evaluated by k_pairs (one warp per inner product; only under BG_KERNEL=warp).
// ROUTE_SHB = evaluated by k_pairs_shb (shared high-block reduction, bg_shb.cuh; at most SHB_MAXLAM parity checks).
enum { ROUTE_DEAD = 0, ROUTE_TPP = 1, ROUTE_WARP = 2, ROUTE_TPP_MANY = 3, ROUTE_SHB = 4 };
struct SampleRec {          // one projected theta in ambient form (see bg_device.cuh: ambient())
    int32_t alive, k1, npf, Q;
    uint64_t D1, D2, Cpend, Cbeta;
    uint64_t J[BG_MAX_T];
    uint64_t Cw[BG_MAX_T];
};

enum { SRC_RNG = 0, SRC_STATES = 1, SRC_TERMS = 2 };

struct PrepArgs {
    SampleRec* recs;
    int n_samples;          // records to produce on this rank
    int t;
    int project;
    const bg_projector* P;
    // SRC_RNG
    uint64_t seed; uint32_t bin; uint64_t first, stride;
    // fused two-projector job (SRC_RNG): records [n_first, n_samples) belong to projector P2 / seed2, sample
    // index restarting at `first`; n_first = n_samples otherwise
    int n_first; const bg_projector* P2; uint64_t seed2;
    const double* cdf;
    // SRC_STATES
    const bg_state* states;
    // SRC_TERMS
    const uint64_t* terms; int exact;
    // optional dump of the native state before projection (active-mask layout)
    bg_state* raw_out; uint64_t* raw_A;
    long long* zw; long long* zw2;      // per-sample accumulators, zeroed here (zw2 may be null)
    int force_warp;                     // route every sample to the warp-per-pair kernel
    int shb;                            // the decomposition has a shared high-block plan: samples with few checks -> ROUTE_SHB
    unsigned long long* n_warp_routed;  // device counter: samples NOT taken by the plain k_pairs_tpp
};

struct PairArgs {
    const SampleRec* recs;
    int n_samples;
    const uint64_t* terms;   // the chi terms, SORTED by popcount so that the 32 terms a warp takes together
                             // have active sets of (nearly) equal size; sums are exact integers, order-free
    const int32_t* term_nat; // natural index i of sorted position (epm output, pair ordering of the exact norm)
    int nterms;
    int t;
    // work items: first n_samples * chunks_per_sample big items (terms [c*chunk, (c+1)*chunk) below tail_start),
    // then n_samples * tail_chunks small ones of 32 terms each covering [tail_start, nterms) — big items
    // first, small ones last, so that the last wave is short (the counter hands them out in this order)
    int chunk, chunks_per_sample, tail_start, tail_chunks, tail_size;
    unsigned long long* counter;
    long long* zw;          // [n_samples][4]
    long long* zw2;         // [n_samples][4]   (exact-norm mode: off-diagonal part)
    int32_t* epm;           // optional [n_samples][nterms][3]
    int smem_terms;         // terms staged in shared memory (0: read from global / L2)
    int tri;                // exact-norm mode: sample i meets terms j >= first_index(i)
    int natural;            // the terms are in natural order (term_nat is the identity): tri mode, so that a sample's
                            // terms j >= i are a contiguous range and no lane is masked off (innerprod.c:203-215 unranks
                            // the pair index the same way)
    uint64_t first, stride; // global index of sample idx = first + idx*stride   (tri mode)
    unsigned long long* pair_count;   // total pairs evaluated (for the throughput metric)
    const unsigned long long* n_warp_routed;   // samples routed to the warp-per-pair kernel
    const ShbPerm* shb;                        // k_pairs_shb: the relabelling of the variables (bg_shb_plan.h), in device
                                               // memory — a new L of the same shape changes no kernel argument, so the
                                               // captured graph of the job survives bg_set_decomposition
    // k_pairs_shb's work items: samples [0, split_from) are one item each (relabelling a sample costs half a batch),
    // the rest — about one wave of resident warps, handed out last — go out in `pieces` items of `piece` terms,
    // so that the last wave is short
    int split_from, piece, pieces;
    int lam_max;                               // samples with at most this many parity checks carry them as Lagrange
                                               // variables t .. t+lam_max-1 of the ambient form (t + lam_max <= word size)
};

// work item -> (sample index, term range)
__device__ __forceinline__ void item_range(const PairArgs& a, unsigned long long item, int& idx, int& i0, int& i1) {
    const unsigned long long nbig = (unsigned long long)a.n_samples * (unsigned)a.chunks_per_sample;
    if (item < nbig) {
        idx = (int)(item / (unsigned)a.chunks_per_sample);
        const int c = (int)(item % (unsigned)a.chunks_per_sample);
        i0 = c * a.chunk; i1 = min(a.tail_start, i0 + a.chunk);
    } else {
        const unsigned long long j = item - nbig;
        idx = (int)(j / (unsigned)a.tail_chunks);
        const int c = (int)(j % (unsigned)a.tail_chunks);
        i0 = a.tail_start + a.tail_size * c; i1 = min(a.nterms, i0 + a.tail_size);
    }
}

// ------------------------------------------------------------------------------------------
// k_prepare
// ------------------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void term_native(Native<NS>& st, int t, int exact, uint64_t term) {
    typedef typename WordOf<NS>::T W;
    const int lane = bg_lane();
    native_identity<NS>(st, t);
    if (!exact) {                         // prepL (stateprep.c:85-120): |+> on supp(x~), |0> elsewhere
        st.f.A = (W)term & lowmaskw<W>(t);
        return;
    }
    // prepH (stateprep.c:36-81)
    const W e1 = (W)term;
    const W maskt = lowmaskw<W>(t);
    const W pairs = (W)0x5555555555555555ull & (maskt >> 1);
    const W cz = ~e1 & pairs, cz2 = cz | (cz << 1);
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        if ((cz2 >> v) & 1) st.f.J[s] = bitw<W>(v ^ 1);
    }
    for (W rem = e1 & pairs; rem;) {
        const int q = lowestw(rem); rem &= rem - 1;
        native_shrink<NS>(st, bitw<W>(q) | bitw<W>(q + 1), 0u, false);
    }
    if ((t & 1) && ((e1 >> (t - 1)) & 1)) native_shrink<NS>(st, bitw<W>(t - 1), 0u, false);
}

template <int NS, int SRC>
__global__ void __launch_bounds__(128) k_prepare(PrepArgs a) {
    typedef typename WordOf<NS>::T W;
    const int lane = bg_lane();
    const int warps_per_block = blockDim.x >> 5;
    const int gw = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int nw = gridDim.x * warps_per_block;
    for (int idx = gw; idx < a.n_samples; idx += nw) {
        Native<NS> st;
        const bool second = SRC == SRC_RNG && idx >= a.n_first;
        const bg_projector* P = second ? a.P2 : a.P;
        if (SRC == SRC_RNG) native_random<NS>(st, a.t, second ? a.seed2 : a.seed, a.bin,
                                              a.first + (uint64_t)(second ? idx - a.n_first : idx) * a.stride, a.cdf);
        else if (SRC == SRC_STATES) native_load<NS>(st, &a.states[idx]);
        else term_native<NS>(st, a.t, a.exact, a.terms[a.first + (uint64_t)idx * a.stride]);
        if (a.raw_out) native_store_raw<NS>(st, &a.raw_out[idx], &a.raw_A[idx]);
        // theta in ambient form first (once), then the projector's generators as mask operations on it
        // (bg_device.cuh: ambient_measure) — no G / Gbar updates, no conversion afterwards
        Ambient<NS> am;
        make_ambient<NS>(st, am);
        int npf = 0;
        bool alive = true;
        if (a.project) alive = project_ambient<NS>(am, P, npf);
        SampleRec* r = &a.recs[idx];
        if (lane < 4) {
            if (a.zw) a.zw[(size_t)idx * 4 + lane] = 0;
            if (a.zw2) a.zw2[(size_t)idx * 4 + lane] = 0;
        }
        if (!alive) {
            if (lane == 0) { r->alive = 0; r->k1 = 0; r->npf = 0; r->Q = 0; }
            continue;
        }
        if (lane == 0) {
            const int nchk = popcw(am.Cpend);
            const int route = a.force_warp ? ROUTE_WARP
                              : a.shb ? (nchk > SHB_MAXLAM ? ROUTE_TPP_MANY : ROUTE_SHB)
                                      : (nchk > TPP_MAXC ? ROUTE_TPP_MANY : ROUTE_TPP);
            if (route != ROUTE_TPP && route != ROUTE_SHB) atomicAdd(a.n_warp_routed, 1ull);
            r->alive = route; r->k1 = am.k1; r->npf = npf; r->Q = (int32_t)am.f.Q;
            r->D1 = (uint64_t)am.f.D1; r->D2 = (uint64_t)am.f.D2;
            r->Cpend = (uint64_t)am.Cpend; r->Cbeta = (uint64_t)am.Cbeta;
        }
#pragma unroll
        for (int s = 0; s < NS; s++) {
            r->J[lane + 32 * s] = (uint64_t)am.f.J[s];
            r->Cw[lane + 32 * s] = (uint64_t)am.Cw[s];
        }
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------
// k_prepare_tps — the hot path's k_prepare: one THREAD per sample (bg_prep.cuh)
// ------------------------------------------------------------------------------------------
// Device-drawn theta, projected: what a sampled probability() job needs.  (k_prepare stays for host-supplied states,
// decomposition terms, unprojected samples, the raw-state dump, and under BG_PREP=warp.)
// Shared memory: 2 * nr rows (J, then the parity checks; nr = t rounded up to 8) of THREADS + 1 words: a
// thread walks its column, consecutive lanes in consecutive banks; the odd row length makes the transposed read
// of the write-out (lanes across rows, one sample at a time: coalesced 256-byte stores) at most two-way conflicted.
// CTAs of 64 threads by default; of 32 in overlap mode (bg_ctx::overlap), where one such CTA — 2048 registers, 21 KB of
// shared memory at t = 40 — fits an SM beside the ten resident CTAs of k_pairs_shb.
template <typename W, int THREADS> static size_t prep_tps_smem(int t) {
    return (size_t)2 * ((t + 7) & ~7) * (THREADS + 1) * sizeof(W);
}
template <typename W, int THREADS>
__global__ void __launch_bounds__(THREADS) k_prepare_tps(PrepArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NS = (int)(sizeof(W) / 4);
    constexpr uint32_t RS = (THREADS + 1) * (uint32_t)sizeof(W);
    const int lane = bg_lane(), wib = (int)(threadIdx.x >> 5);
    const int n = a.t, nr = (n + 7) & ~7;
    Rows<W> J, C;
    J.base = nullptr; J.stride = 0; C.base = nullptr; C.stride = 0;
    J.sbase = smem_u32(smem_raw) + threadIdx.x * (uint32_t)sizeof(W); J.sstride = RS;
    C.sbase = J.sbase + (uint32_t)nr * RS; C.sstride = RS;
    const uint32_t wbase = smem_u32(smem_raw) + (uint32_t)(wib * 32) * (uint32_t)sizeof(W);
    const int warps = (int)gridDim.x * (THREADS / 32), gw = (int)blockIdx.x * (THREADS / 32) + wib;
    const int ngroups = (a.n_samples + 31) >> 5;
    for (int g = gw; g < ngroups; g += warps) {
        const int idx = g * 32 + lane;
        TSample<W> s;
        s.Cpend = 0;
        int route = ROUTE_DEAD;
        if (idx < a.n_samples) {
            const bool second = idx >= a.n_first;
            t_random_ambient<W>(J, C, n, second ? a.seed2 : a.seed, a.bin,
                                a.first + (uint64_t)(second ? idx - a.n_first : idx) * a.stride, a.cdf, s);
            t_project<W>(J, C, n, s, second ? a.P2 : a.P);
            if (a.zw) { longlong2* z = reinterpret_cast<longlong2*>(a.zw + (size_t)idx * 4); z[0] = z[1] = make_longlong2(0, 0); }
            if (a.zw2) { longlong2* z = reinterpret_cast<longlong2*>(a.zw2 + (size_t)idx * 4); z[0] = z[1] = make_longlong2(0, 0); }
            SampleRec* r = &a.recs[idx];
            if (s.alive) {
                const int nchk = tpopc(s.Cpend);
                route = a.force_warp ? ROUTE_WARP
                        : a.shb ? (nchk > SHB_MAXLAM ? ROUTE_TPP_MANY : ROUTE_SHB)
                                : (nchk > TPP_MAXC ? ROUTE_TPP_MANY : ROUTE_TPP);
                if (route != ROUTE_TPP && route != ROUTE_SHB) atomicAdd(a.n_warp_routed, 1ull);
                *reinterpret_cast<int4*>(&r->alive) = make_int4(route, n - nchk, s.npf, (int)s.Q);
                ulonglong2* q = reinterpret_cast<ulonglong2*>(&r->D1);
                q[0] = make_ulonglong2((uint64_t)s.D1, (uint64_t)s.D2);
                q[1] = make_ulonglong2((uint64_t)s.Cpend, (uint64_t)s.Cbeta);
            } else {
                *reinterpret_cast<int4*>(&r->alive) = make_int4(0, 0, 0, 0);
            }
        }
        __syncwarp();
        for (int j = 0; j < 32; j++) {
            const int sidx = g * 32 + j;
            if (sidx >= a.n_samples) break;
            const int rt = __shfl_sync(BG_FULL, route, j);
            const W cp = shflw(s.Cpend, j);
            if (rt == ROUTE_DEAD) continue;
            SampleRec* r = &a.recs[sidx];
#pragma unroll
            for (int hh = 0; hh < NS; hh++) {
                const int v = lane + 32 * hh;
                W jv = 0, cv = 0;
                if (v < n) {
                    const uint32_t ad = wbase + (uint32_t)j * (uint32_t)sizeof(W) + (uint32_t)v * RS;
                    t_lds(ad, jv);
                    if ((cp >> v) & 1) t_lds(ad + (uint32_t)nr * RS, cv);
                }
                r->J[v] = (uint64_t)jv;
                r->Cw[v] = (uint64_t)cv;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// k_pairs
// ------------------------------------------------------------------------------------------

// Stage `bytes` (multiple of 16, 16-byte aligned on both sides) from global to shared memory
// with the TMA bulk-copy engine; completion is signalled on an mbarrier.  SASS: UBLKCP.
__device__ __forceinline__ void tma_stage(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
        uint32_t off = 0;
        while (off < bytes) {
            const uint32_t n = min(bytes - off, 32768u);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32((const char*)smem_dst + off)), "l"((const char*)gsrc + off), "r"(n),
                           "r"(smem_u32(mbar)) : "memory");
            off += n;
        }
    }
    // every thread waits for phase 0 of the barrier
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(mbar)) : "memory");
    }
}

template <int NS, bool EXACT>
__global__ void __launch_bounds__(128) k_pairs(PairArgs a) {
    typedef typename WordOf<NS>::T W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s_terms = reinterpret_cast<uint64_t*>(smem_raw);
    __shared__ __align__(8) uint64_t s_mbar;
    const int lane = bg_lane();
    if (*a.n_warp_routed == 0ull) return;       // every sample went to k_pairs_tpp
    if (a.smem_terms > 0) tma_stage(s_terms, a.terms, (uint32_t)a.smem_terms * 8u, &s_mbar);
    const uint64_t* terms = a.smem_terms > 0 ? s_terms : a.terms;

    const unsigned long long n_items = (unsigned long long)a.n_samples * (unsigned long long)(a.chunks_per_sample + a.tail_chunks);
    const int sh = a.t / 2 + 1;
    unsigned long long my_pairs = 0;
    while (true) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(a.counter, 1ull);
        item = __shfl_sync(BG_FULL, item, 0);
        if (item >= n_items) break;
        int idx, i0, i1;
        item_range(a, item, idx, i0, i1);
        const SampleRec* r = &a.recs[idx];
        if (r->alive != ROUTE_WARP) continue;
        // exactProjectorWork: sample i meets the terms j >= i (natural indices)
        const int diag_index = a.tri ? (int)(a.first + (uint64_t)idx * a.stride) : -1;
        if (i0 >= i1) continue;
        Ambient<NS> am;
        am.f.Q = (uint32_t)r->Q; am.f.D1 = (W)r->D1; am.f.D2 = (W)r->D2;
        am.f.A = lowmaskw<W>(a.t);
        am.Cpend = (W)r->Cpend; am.Cbeta = (W)r->Cbeta; am.k1 = r->k1;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            am.f.J[s] = (W)r->J[lane + 32 * s];
            am.Cw[s] = (W)r->Cw[lane + 32 * s];
        }
        Zw z, z2;
        z.a[0] = z.a[1] = z.a[2] = z.a[3] = 0;
        z2.a[0] = z2.a[1] = z2.a[2] = z2.a[3] = 0;
        for (int i = (a.tri && a.natural) ? max(i0, diag_index) : i0; i < i1; i++) {
            const W term = (W)terms[i];
            const int nat = (!a.natural && (a.tri || a.epm)) ? a.term_nat[i] : i;
            if (a.tri && nat < diag_index) continue;
            int e, p, m;
            if (EXACT) term_H<NS>(am.f, am.Cw, am.Cpend, am.Cbeta, am.k1, a.t, term, e, p, m);
            else term_L<NS>(am.f, am.Cw, am.Cpend, am.Cbeta, am.k1, term, e, p, m);
            if (a.tri && nat != diag_index) zw_add(z2, e, p, m, sh);
            else zw_add(z, e, p, m, sh);
            if (a.epm && lane == 0) {
                int32_t* o = a.epm + ((size_t)idx * a.nterms + nat) * 3;
                o[0] = e; o[1] = p; o[2] = m & 7;
            }
            my_pairs++;
        }
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (z.a[j]) atomicAdd((unsigned long long*)&a.zw[(size_t)idx * 4 + j], (unsigned long long)z.a[j]);
                if (a.tri && z2.a[j]) atomicAdd((unsigned long long*)&a.zw2[(size_t)idx * 4 + j], (unsigned long long)z2.a[j]);
            }
        }
    }
    if (lane == 0 && my_pairs) atomicAdd(a.pair_count, my_pairs);
}

// ------------------------------------------------------------------------------------------
// k_pairs_tpp: one thread per inner product (bg_tpp.cuh).  A warp = one theta x 32 terms.
// dynamic smem: [staged terms][ambient J: warps x t words][working rows: t x blockDim words]
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ long long shfl_down_ll(long long v, int d) {
    uint32_t lo = __shfl_down_sync(BG_FULL, (uint32_t)(unsigned long long)v, d);
    uint32_t hi = __shfl_down_sync(BG_FULL, (uint32_t)((unsigned long long)v >> 32), d);
    return (long long)(((unsigned long long)hi << 32) | lo);
}

// ambient rows per warp (+ the Lagrange rows of the checks, or all check rows when MANYC), padded to a
// multiple of 4 words: 16-byte aligned
__host__ __device__ __forceinline__ int tpp_amb_rows(int t, int lam_max, bool manyc) { return ((manyc ? 2 * t : t + lam_max) + 3) & ~3; }
// working rows per thread
__host__ __device__ __forceinline__ int tpp_work_rows(int t, int lam_max, bool manyc) { return manyc ? 2 * t : t + lam_max; }

template <typename W, bool EXACT, bool TRI, bool MANYC>
__global__ void __launch_bounds__(BG_TPP_THREADS) k_pairs_tpp(PairArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar;
    const int lane = bg_lane(), warp = threadIdx.x >> 5;
    constexpr int nwarps = BG_TPP_WARPS;
    const int t = a.t;
    if (MANYC && *a.n_warp_routed == 0ull) return;      // nothing has more than TPP_MAXC parity checks
    // per warp: t ambient rows (+ t check rows when MANYC); per thread: t working rows (+ t history rows)
    const int amb_rows = tpp_amb_rows(t, a.lam_max, MANYC);
    uint64_t* s_terms = reinterpret_cast<uint64_t*>(smem_raw);
    W* s_amb = reinterpret_cast<W*>(smem_raw + (size_t)a.smem_terms * 8) + warp * amb_rows;
    W* s_rows = reinterpret_cast<W*>(smem_raw + (size_t)a.smem_terms * 8) + nwarps * amb_rows + threadIdx.x;
    if (a.smem_terms > 0) tma_stage(s_terms, a.terms, (uint32_t)a.smem_terms * 8u, &s_mbar);
    const uint64_t* terms = a.smem_terms > 0 ? s_terms : a.terms;
    Rows<W> rows; rows.base = s_rows; rows.stride = BG_TPP_THREADS;
    rows.sbase = smem_u32(s_rows); rows.sstride = BG_TPP_THREADS * (uint32_t)sizeof(W);

    const unsigned long long n_items = (unsigned long long)a.n_samples * (unsigned long long)(a.chunks_per_sample + a.tail_chunks);
    const int sh_ = t / 2 + 1;
    unsigned long long my_pairs = 0;
    while (true) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(a.counter, 1ull);
        item = __shfl_sync(BG_FULL, item, 0);
        if (item >= n_items) break;
        int idx, i0, i1;
        item_range(a, item, idx, i0, i1);
        const SampleRec* r = &a.recs[idx];
        if (r->alive != (MANYC ? ROUTE_TPP_MANY : ROUTE_TPP)) continue;
        const int diag_index = TRI ? (int)(a.first + (uint64_t)idx * a.stride) : -1;
        if (i0 >= i1) continue;
        __syncwarp();
        if (MANYC) for (int q = lane; q < t; q += 32) s_amb[q] = (W)r->J[q];
        TShared<W> sh;
        sh.J = s_amb; sh.D1 = (W)r->D1; sh.D2 = (W)r->D2; sh.Q = (uint32_t)r->Q; sh.k1 = r->k1; sh.t = t;
        sh.ncons = 0; sh.nlam = 0; sh.cbeta = 0; sh.cwv = s_amb + t; sh.cbetav = 0;
        if (MANYC) {                      // compact the check rows into shared memory, in row order
            const uint64_t pend = r->Cpend, cbeta = r->Cbeta;
            for (int q = lane; q < t; q += 32)
                if ((pend >> q) & 1ull) s_amb[t + __popcll(pend & ((1ull << q) - 1ull))] = (W)r->Cw[q];
            sh.ncons = __popcll(pend);
            uint64_t cb = 0;
            for (uint64_t rem = pend; rem; rem &= rem - 1) {
                const int b = __ffsll((long long)rem) - 1;
                cb |= ((cbeta >> b) & 1ull) << __popcll(pend & ((1ull << b) - 1ull));
            }
            sh.cbetav = (W)cb;
        } else {
            const uint64_t cbeta = r->Cbeta;
#pragma unroll
            for (int j = 0; j < TPP_MAXC; j++) sh.cw[j] = 0;
            uint64_t pend = r->Cpend;
#pragma unroll
            for (int j = 0; j < TPP_MAXC; j++) {
                if (pend) {
                    const int b = __ffsll((long long)pend) - 1;
                    pend &= pend - 1;
                    sh.cw[j] = (W)r->Cw[b];
                    sh.cbeta |= (uint32_t)((cbeta >> b) & 1ull) << j;
                    sh.ncons = j + 1;
                }
            }
            // few checks: append them to the ambient form as Lagrange variables t .. t+nlam-1 (bg_tpp.cuh)
            const bool lam = sh.ncons <= a.lam_max;
            for (int q = lane; q < t; q += 32) {
                W row = (W)r->J[q];
                if (lam) {
#pragma unroll
                    for (int j = 0; j < TPP_MAXC; j++) row |= (W)((sh.cw[j] >> q) & 1) << ((t + j) & (8 * (int)sizeof(W) - 1));
                }
                s_amb[q] = row;
            }
            if (lam) {
#pragma unroll
                for (int j = 0; j < TPP_MAXC; j++) if (j < sh.ncons && lane == j) s_amb[t + j] = sh.cw[j];
                sh.D2 |= (W)sh.cbeta << t;
                sh.nlam = sh.ncons; sh.ncons = 0;
            }
        }
        __syncwarp();
        Zw z, z2;
        z.a[0] = z.a[1] = z.a[2] = z.a[3] = 0;
        z2.a[0] = z2.a[1] = z2.a[2] = z2.a[3] = 0;
        for (int g = (TRI && a.natural) ? max(i0, diag_index & ~31) : i0; g < i1; g += 32) {
            const int i = g + lane;
            const int nat = (i < i1 && !a.natural && (TRI || a.epm)) ? a.term_nat[i] : i;
            if (i < i1 && !(TRI && nat < diag_index)) {
                const unsigned group = __activemask();
                const W term = (W)terms[i];
                int e, p, m;
                if (EXACT) t_term_H<W, MANYC>(rows, sh, term, e, p, m);
                else t_term_L<W, MANYC>(rows, sh, term, e, p, m);
                __syncwarp(group);          // the lanes leave the elimination at different times: accumulate together
                if (TRI && nat != diag_index) zw_add(z2, e, p, m, sh_);
                else zw_add(z, e, p, m, sh_);
                if (a.epm) {
                    int32_t* o = a.epm + ((size_t)idx * a.nterms + nat) * 3;
                    o[0] = e; o[1] = p; o[2] = m & 7;
                }
                my_pairs++;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                z.a[j] += shfl_down_ll(z.a[j], d);
                if (TRI) z2.a[j] += shfl_down_ll(z2.a[j], d);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (z.a[j]) atomicAdd((unsigned long long*)&a.zw[(size_t)idx * 4 + j], (unsigned long long)z.a[j]);
                if (TRI && z2.a[j]) atomicAdd((unsigned long long*)&a.zw2[(size_t)idx * 4 + j], (unsigned long long)z2.a[j]);
            }
        }
    }
    for (int d = 16; d > 0; d >>= 1) my_pairs += __shfl_down_sync(BG_FULL, my_pairs, d);
    if (lane == 0 && my_pairs) atomicAdd(a.pair_count, my_pairs);
}


// ------------------------------------------------------------------------------------------
// k_pairs_shb: the L x chi loop for 32 < t <= 44 with a shared high-block plan (bg_shb.cuh).  One thread per
// inner product, a warp = one theta x 32 terms that agree on the variables >= 32: the warp sums those out once
// per batch (shb_reduce), every thread then eliminates a form on <= 32 variables in 32-bit words.
// dynamic smem: [staged terms][per warp: 32 reduced rows + SHB_MAXHT leftover rows][working rows: 32 x blockDim words]
// ------------------------------------------------------------------------------------------
// A warp works on SHB_G classes (patterns) of one sample at a time: 32 / SHB_G lanes per class, and the terms of a
// class are sorted by popcount, so the lanes that run together have active sets of nearly equal size (the warp
// waits for its largest term): round r takes the r-th slice of each of the SHB_G classes.
#ifndef SHB_G
#define SHB_G 2
#endif
#define SHB_CLASS_WORDS ((32 + SHB_MAXHT + 3) & ~3)          // 16-byte aligned: the reduced rows are read with LDS.128
#define SHB_WARP_WORDS (SHB_G * SHB_CLASS_WORDS)
// the warp's 32-bit partial sums -> the sample's 64-bit accumulators (warp-uniform call)
__device__ __forceinline__ void shb_flush(Zw32& zs, long long* zw, int lane) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        long long v = (long long)zs.a[j];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += shfl_down_ll(v, d);
        if (lane == 0 && v) atomicAdd((unsigned long long*)&zw[j], (unsigned long long)v);
        zs.a[j] = 0;
    }
}
__global__ void __launch_bounds__(BG_TPP_THREADS) k_pairs_shb(PairArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar;
    const int lane = bg_lane(), warp = threadIdx.x >> 5;
    constexpr int nwarps = BG_TPP_WARPS;
    constexpr int LPG = 32 / SHB_G;                             // lanes per class
    const int t = a.t;
    uint64_t* s_terms = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* s_warp = reinterpret_cast<uint32_t*>(smem_raw + (size_t)a.smem_terms * 8) + warp * SHB_WARP_WORDS;
    uint32_t* s_rows = reinterpret_cast<uint32_t*>(smem_raw + (size_t)a.smem_terms * 8) + nwarps * SHB_WARP_WORDS + threadIdx.x;
    __shared__ ShbPerm s_pm;
    static_assert(sizeof(ShbPerm) % 4 == 0 && sizeof(ShbPerm) / 4 <= BG_TPP_THREADS, "ShbPerm is copied one word per thread");
    if (threadIdx.x < sizeof(ShbPerm) / 4) reinterpret_cast<uint32_t*>(&s_pm)[threadIdx.x] = reinterpret_cast<const uint32_t*>(a.shb)[threadIdx.x];
    if (a.smem_terms > 0) tma_stage(s_terms, a.terms, (uint32_t)a.smem_terms * 8u, &s_mbar);
    else __syncthreads();
    const uint64_t* terms = a.smem_terms > 0 ? s_terms : a.terms;
    Rows<uint32_t> rows; rows.base = s_rows; rows.stride = BG_TPP_THREADS;
    rows.sbase = smem_u32(s_rows); rows.sstride = BG_TPP_THREADS * 4u;
    const int h = lane / LPG;                                   // the class of the group this lane works on
    uint32_t* my_red = s_warp + h * SHB_CLASS_WORDS;
    uint32_t* my_left = my_red + 32;

    const unsigned long long n_items = (unsigned long long)a.split_from + (unsigned long long)(a.n_samples - a.split_from) * (unsigned)a.pieces;
    const int sh_ = t / 2 + 1;
    unsigned long long my_pairs = 0;
    while (true) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(a.counter, 1ull);
        item = __shfl_sync(BG_FULL, item, 0);
        if (item >= n_items) break;
        int idx, i0, i1;
        if (item < (unsigned long long)a.split_from) { idx = (int)item; i0 = 0; i1 = a.nterms; }
        else {
            const uint32_t j = (uint32_t)(item - (unsigned long long)a.split_from);      // < 2^30 x pieces (host-checked)
            const uint32_t s_ = j / (uint32_t)a.pieces;
            idx = a.split_from + (int)s_;
            i0 = (int)(j - s_ * (uint32_t)a.pieces) * a.piece; i1 = min(a.nterms, i0 + a.piece);
        }
        const SampleRec* r = &a.recs[idx];
        if (r->alive != ROUTE_SHB) continue;
        if (i0 >= i1) continue;
        ShbForm f;
        const int nlam = shb_load(r->J, r->Cw, r->Cpend, r->Cbeta, r->D1, r->D2, (uint32_t)r->Q, t, s_pm, f);
        const int k1 = r->k1;
        const uint32_t lam_bits = ((1u << nlam) - 1u) << s_pm.nh;
        Zw32 zs;                                                // this thread's terms since the last flush (t <= 44: 32 bits do)
        zs.a[0] = zs.a[1] = zs.a[2] = zs.a[3] = 0;
        int nacc = 0;
        static_assert(32 + SHB_MAXH <= ZW32_MAX_T, "Zw32 holds 2^(t/2+1) x ZW32_MAX_TERMS");
        for (int g = i0; g < i1; g += 32 * SHB_G) {             // SHB_G classes of 32 terms; every class has one high pattern
            ShbBatch sb;
            sb.red = my_red; sb.left = my_left; sb.k1 = k1; sb.nlam = nlam;
            sb.D1 = sb.D2 = sb.Q = 0; sb.p = 0; sb.nleft = 0; sb.left_d2 = 0;
            int nlmax = 0;
            __syncwarp();
#pragma unroll
            for (int c = 0; c < SHB_G; c++) {
                const uint32_t pattern = (uint32_t)(terms[g + 32 * c] >> 32);
                ShbOut o; uint32_t Lr, Rr;
                shb_reduce(f, pattern | lam_bits, Lr, Rr, o);
                uint32_t* red = s_warp + c * SHB_CLASS_WORDS;
                red[lane] = Lr;
                int nleft = 0; uint32_t ld2 = 0;
                for (uint32_t rem = o.left; rem;) {
                    const int u = shb_top(rem);
                    rem ^= 1u << u;
                    const uint32_t w = __shfl_sync(BG_FULL, Rr, u);
                    if (lane == 0) red[32 + nleft] = w;
                    ld2 |= ((o.left_d2 >> u) & 1u) << nleft;
                    nleft++;
                }
                nlmax = max(nlmax, nleft);
                if (h == c) { sb.D1 = o.D1; sb.D2 = o.D2; sb.Q = o.Q; sb.p = (int)o.p; sb.nleft = nleft; sb.left_d2 = ld2; }
            }
            __syncwarp();
#pragma unroll 1
            for (int rd = 0; rd < SHB_G; rd++) {
                const int ti = g + 32 * h + LPG * rd + (lane & (LPG - 1));
                const uint64_t term = terms[ti];
                int e, p, m;
                t_term_shb(rows, sb, term, nlmax, e, p, m);
                __syncwarp();                                   // the lanes leave the elimination at different times
                zw32_add(zs, e, p, m, sh_);
                if (a.epm) {
                    int32_t* o3 = a.epm + ((size_t)idx * a.nterms + a.term_nat[ti]) * 3;
                    o3[0] = e; o3[1] = p; o3[2] = m & 7;
                }
                my_pairs++;
            }
            nacc += SHB_G;
            if (nacc > ZW32_MAX_TERMS - SHB_G && g + 32 * SHB_G < i1) {      // (an item of more than 128 terms per thread: k >= 13)
                shb_flush(zs, a.zw + (size_t)idx * 4, lane);
                nacc = 0;
            }
        }
        shb_flush(zs, a.zw + (size_t)idx * 4, lane);
    }
    for (int d = 16; d > 0; d >>= 1) my_pairs += __shfl_down_sync(BG_FULL, my_pairs, d);
    if (lane == 0 && my_pairs) atomicAdd(a.pair_count, my_pairs);
}

// generic pairs: one warp per pair
template <int NS>
__global__ void __launch_bounds__(128) k_inner_products(const bg_state* a, const bg_state* b, size_t n_pairs, int32_t* epm) {
    const int warps_per_block = blockDim.x >> 5;
    const size_t gw = (size_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const size_t nw = (size_t)gridDim.x * warps_per_block;
    for (size_t i = gw; i < n_pairs; i += nw) {
        int e, p, m;
        warp_inner_product<NS>(&a[i], &b[i], e, p, m);
        if (bg_lane() == 0) { epm[3 * i] = e; epm[3 * i + 1] = p; epm[3 * i + 2] = m & 7; }
    }
}

template <int NS>
__global__ void __launch_bounds__(128) k_measure_pauli(bg_state* st, uint64_t* A, size_t n, const int32_t* m,
                                                       const uint64_t* zeta, const uint64_t* xi, int32_t* code) {
    const int warps_per_block = blockDim.x >> 5;
    const size_t gw = (size_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const size_t nw = (size_t)gridDim.x * warps_per_block;
    for (size_t i = gw; i < n; i += nw) {
        const int r = warp_measure_pauli<NS>(&st[i], &A[i], m[i], zeta[i], xi[i]);
        if (bg_lane() == 0) code[i] = r;
    }
}

// ------------------------------------------------------------------------------------------
// k_finalize: per-sample value and fixed-order reduction
// ------------------------------------------------------------------------------------------
// (a0 + a1 w + a2 w^2 + a3 w^3), w = e^{i pi/4}:  re = a0 + (a1-a3)/sqrt2, im = a2 + (a1+a3)/sqrt2
__device__ __forceinline__ void zw_to_complex(const long long* a, double& re, double& im) {
    const double c_hi = 0.70710678118654757, c_lo = -4.8336466567264567e-17;   // 1/sqrt2 = c_hi + c_lo
    const double d1 = (double)(a[1] - a[3]), d2 = (double)(a[1] + a[3]);
    const double p1 = d1 * c_hi, e1 = fma(d1, c_hi, -p1) + d1 * c_lo;
    const double p2 = d2 * c_hi, e2 = fma(d2, c_hi, -p2) + d2 * c_lo;
    const double a0 = (double)a[0], a2 = (double)a[2];
    double s = a0 + p1, bb = s - a0; double err = (a0 - (s - bb)) + (p1 - bb);
    re = s + (err + e1);
    s = a2 + p2; bb = s - a2; err = (a2 - (s - bb)) + (p2 - bb);
    im = s + (err + e2);
}

// sampled mode: value_l = 2^t |projfactor * total|^2   (innerprod.c:142)
__global__ void k_finalize_sampled(const SampleRec* recs, const long long* zw, int n, int t, double* per_sample) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = 0.0;
    if (recs[i].alive) {
        double re, im;
        zw_to_complex(&zw[(size_t)i * 4], re, im);
        const int sh = t / 2 + 1;
        v = ldexp(re * re + im * im, t - recs[i].npf - 2 * sh);
    }
    per_sample[i] = v;
}

// finalize + fixed-order sum in ONE launch of FIN_BLOCKS CTAs: thread (b, t) adds its samples
// i = b*1024 + t, + FIN_BLOCKS*1024, ... in order; each CTA tree-reduces to a partial; the CTA that
// finishes last (atomic ticket) adds the partials in CTA order.  Same bits on every run.
#define FIN_BLOCKS 32
__global__ void __launch_bounds__(1024) k_finalize_sum_sampled(const SampleRec* recs, const long long* zw, int n, int t,
                                                               double* per_sample, double* out, double* partials,
                                                               unsigned int* ticket, int out_stride) {
    __shared__ double sh[1024];
    __shared__ bool last;
    // blockIdx.y = segment: samples [seg * n, (seg + 1) * n) are one projector's (fused two-projector job)
    const int seg = blockIdx.y;
    recs += (size_t)seg * n; zw += (size_t)seg * n * 4; per_sample += (size_t)seg * n;
    out += seg * out_stride; partials += seg * FIN_BLOCKS; ticket += seg;
    const int shf = t / 2 + 1;
    double acc = 0.0;
    for (int i = blockIdx.x * 1024 + threadIdx.x; i < n; i += FIN_BLOCKS * 1024) {
        double v = 0.0;
        if (recs[i].alive) {
            double re, im;
            zw_to_complex(&zw[(size_t)i * 4], re, im);
            v = ldexp(re * re + im * im, t - recs[i].npf - 2 * shf);
        }
        per_sample[i] = v;
        acc += v;
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = sh[0];
        __threadfence();
        last = atomicAdd(ticket, 1u) == FIN_BLOCKS - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double tot = 0.0;
        for (int b = 0; b < FIN_BLOCKS; b++) tot += ((volatile double*)partials)[b];
        *out = tot;
        *ticket = 0u;                    // ready for the next launch
    }
}

// exact mode: part_i = pf_i * (diag_i) if i == j, plus (2 pf_i Re(offdiag_i), 0)   (innerprod.c:254-260)
__global__ void k_finalize_exact(const SampleRec* recs, const long long* zw, const long long* zw2, int n, int t,
                                 double* per_re, double* per_im) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double vr = 0.0, vi = 0.0;
    if (recs[i].alive) {
        double re, im, re2, im2;
        zw_to_complex(&zw[(size_t)i * 4], re, im);
        zw_to_complex(&zw2[(size_t)i * 4], re2, im2);
        const int sh = t / 2 + 1;
        const double pf = pow(2.0, -0.5 * recs[i].npf);
        vr = ldexp((re + 2.0 * re2) * pf, -sh);
        vi = ldexp(im * pf, -sh);
    }
    per_re[i] = vr; per_im[i] = vi;
}

// out[0] = sum(x[0..n)) in a fixed order (thread-strided partials, then a binary tree)
__global__ void __launch_bounds__(1024) k_sum(const double* x, int n, double* out) {
    __shared__ double sh[1024];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) acc += x[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /*ncclUniqueId by value*/ ...) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
struct NcclUniqueId { char internal[128]; };
typedef int (*nccl_init_rank_fn)(void**, int, NcclUniqueId, int);
static NcclApi g_nccl;

struct bg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    // Two job slots so that the all-reduce + read-back of one prepared job overlap the kernels of the
    // next (bg_sampled_run may be called twice before bg_sampled_finish).  Per slot: timing events,
    // 8 reduction outputs, 8 counters, one captured graph, one pinned result buffer.
    int slot = 0;
    cudaEvent_t ev0s[2] = {nullptr, nullptr}, ev1s[2] = {nullptr, nullptr};
    cudaEvent_t evps[2][2][3] = {};          // [slot][projector]: before prepare, after prepare, after pairs
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    cudaStream_t cstream = nullptr;          // all-reduce + device->host copies
    unsigned run_seq = 0, fin_seq = 0;
    int sm_count = 0;
    int rank = 0, world = 1;
    bool allreduce = true;
    void* nccl_comm = nullptr;
    // decomposition
    int t = 0, exact = 1, k = 0;
    std::vector<uint64_t> L;
    std::vector<uint64_t> terms_host;
    uint64_t* d_terms = nullptr; size_t d_terms_cap = 0;             // natural order
    uint64_t* d_terms_sorted = nullptr; size_t d_terms_sorted_cap = 0;  // by popcount (pair kernels)
    int32_t* d_term_nat = nullptr; size_t d_term_nat_cap = 0;
    // shared high-block plan (bg_shb_plan.h): relabelled, pattern-sorted terms for k_pairs_shb
    ShbPlan shb_plan; int use_shb = 1;          // BG_SHB=0 disables
    uint64_t* d_terms_shb = nullptr; size_t d_terms_shb_cap = 0;
    int32_t* d_term_nat_shb = nullptr; size_t d_term_nat_shb_cap = 0;
    ShbPerm* d_shb_perm = nullptr;              // the plan's relabelling (read by k_pairs_shb through PairArgs::shb)
    // what the captured graphs depend on besides buffer addresses: a decomposition with the same signature keeps them
    int sig_t = -1, sig_exact = -1, sig_k = -1, sig_plan_ok = -1, sig_nh = -1; size_t sig_chi = 0;
    double* d_cdf = nullptr; int cdf_t = -1;
    // buffers
    SampleRec* d_recs = nullptr; size_t recs_cap = 0;
    long long* d_zw = nullptr; size_t zw_cap = 0;
    long long* d_zw2 = nullptr; size_t zw2_cap = 0;
    double* d_per = nullptr; size_t per_cap = 0;
    double* d_per2 = nullptr; size_t per2_cap = 0;
    int ctas_per_sm = 8, items_factor = 8; bool items_factor_set = false;
#if defined(BG_ELIM_FOLD)
    int lam_max = 0;
#else
    int lam_max = 4;
#endif                          // BG_LAM_MAX: parity checks carried as Lagrange variables (0: pivot every check per term)
    const int tpp_warps = BG_TPP_WARPS;   // warps per CTA of k_pairs_tpp
    int force_warp = 0;             // BG_KERNEL=warp: evaluate everything with the warp-per-pair kernel
    int pieces_override = 0;        // BG_PIECES: pieces of the last wave's samples in k_pairs_shb (0: chosen by launch_pairs)
    int prep_warp = 0;              // BG_PREP=warp: draw + project the samples with the warp-per-sample k_prepare
    int prep_ctas_per_sm[2][2][BG_MAX_T + 1] = {};   // k_prepare_tps: resident CTAs per SM by (word size, CTA size, t), 0 = not asked yet
    int fuse2 = 1;                  // BG_FUSE2=0: one launch sequence per projector instead of one for both
    bg_projector* d_P = nullptr;
    unsigned long long* d_counters = nullptr;   // [0] work counter, [1] pair count
    double* d_red = nullptr;                    // [2 slots][8] reduction outputs
    double* d_partials = nullptr;               // [FIN_BLOCKS] per-CTA partial sums of k_finalize_sum_sampled
    unsigned int* d_ticket = nullptr;
    // prepared sampled run
    int nproj = 1;                  // projectors of the prepared job (2: numerator and denominator together)
    uint64_t samples = 0; int bins = 1; uint64_t seeds[2] = {0, 0};
    bool prepared = false;
    bool per_valid = false;         // d_per holds the per-sample values of a finished job
    // pinned staging copies of the projectors + their upload events: two sets, used in turn, so that the projectors of
    // the next job can be staged while the previous job (and its upload) is still in flight
    bg_projector* h_stage = nullptr; cudaEvent_t ev_stage[2] = {nullptr, nullptr}; unsigned stage_seq = 0;
    bg_projector h_P[2];            // what d_P holds (a repeated prepare with the same content keeps the captured graph)
    bool P_valid = false;
    bg_projector cur_P[2];          // the prepared job's projectors (host copy)
    // Overlap mode (few samples per GPU: the draw + projection kernel cannot fill the machine and is latency-bound): the
    // jobs in the odd slot run on a stream and buffers of their own (`alt`, swapped in around their calls), so that
    // k_prepare_tps of job i+1 runs beside the pair kernel of job i instead of behind it.
    struct JobSet {
        cudaStream_t stream = nullptr;
        SampleRec* d_recs = nullptr; size_t recs_cap = 0;
        long long* d_zw = nullptr; size_t zw_cap = 0;
        long long* d_zw2 = nullptr; size_t zw2_cap = 0;
        double* d_per = nullptr; size_t per_cap = 0;
        double* d_per2 = nullptr; size_t per2_cap = 0;
        bg_projector* d_P = nullptr; double* d_partials = nullptr; unsigned int* d_ticket = nullptr;
        bg_projector h_P[2]; bool P_valid = false;
    } alt;
    bool alt_in = false;            // alt is swapped into the fields above
    int overlap_mode = -1;          // BG_OVERLAP: 1 always, 0 never, -1 when a job has fewer than OVERLAP_WARPS_PER_SM warps of prepare work per SM
    bool overlap = false;           // the prepared job runs in overlap mode
    int per_set = 0;                // which set holds the per-sample values of the last finished job
    bool phase_events = false;
    int cur = 0;                    // projector being launched (selects counters / events / d_P slot)
    // the prepared job replayed as one CUDA graph (BG_GRAPH=0 disables)
    bool use_graph = true, capturing = false;
    cudaGraphExec_t gexecs[2] = {nullptr, nullptr};
    double* h_out = nullptr;        // pinned: per slot 8 sums + 8 counters
    std::vector<double> bin_sums;
    bg_stats stats;
    std::string err;
};

#define EV0(ctx) ((ctx)->ev0s[(ctx)->slot])
#define EV1(ctx) ((ctx)->ev1s[(ctx)->slot])
#define EVP(ctx, pj, i) ((ctx)->evps[(ctx)->slot][pj][i])
#define RED(ctx) ((ctx)->d_red + 8 * (ctx)->slot)
#define CNT(ctx) ((ctx)->d_counters + 8 * (ctx)->slot)
#define HOUT(ctx) ((ctx)->h_out + 16 * (ctx)->slot)

static int fail(bg_ctx* ctx, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->err = buf;
    return 1;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// scratch device buffer of one call: freed on every return path
template <typename T> struct DevBuf {
    T* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc((void**)&p, n * sizeof(T)); }
    operator T*() const { return p; }
};

static cudaError_t rec_event(bg_ctx* ctx, cudaEvent_t ev) {
    return ctx->capturing ? cudaEventRecordWithFlags(ev, ctx->stream, cudaEventRecordExternal) : cudaEventRecord(ev, ctx->stream);
}
static void drop_graph(bg_ctx* ctx) {
    for (int sl = 0; sl < 2; sl++) if (ctx->gexecs[sl]) { cudaGraphExecDestroy(ctx->gexecs[sl]); ctx->gexecs[sl] = nullptr; }
}

// exchange the job-owned resources of ctx with the alternate set
static void swap_set(bg_ctx* ctx) {
    bg_ctx::JobSet& a = ctx->alt;
    std::swap(ctx->stream, a.stream);
    std::swap(ctx->d_recs, a.d_recs); std::swap(ctx->recs_cap, a.recs_cap);
    std::swap(ctx->d_zw, a.d_zw); std::swap(ctx->zw_cap, a.zw_cap);
    std::swap(ctx->d_zw2, a.d_zw2); std::swap(ctx->zw2_cap, a.zw2_cap);
    std::swap(ctx->d_per, a.d_per); std::swap(ctx->per_cap, a.per_cap);
    std::swap(ctx->d_per2, a.d_per2); std::swap(ctx->per2_cap, a.per2_cap);
    std::swap(ctx->d_P, a.d_P); std::swap(ctx->d_partials, a.d_partials); std::swap(ctx->d_ticket, a.d_ticket);
    std::swap(ctx->h_P[0], a.h_P[0]); std::swap(ctx->h_P[1], a.h_P[1]); std::swap(ctx->P_valid, a.P_valid);
    ctx->alt_in = !ctx->alt_in;
}
struct SetGuard {               // the alternate set for the duration of a scope
    bg_ctx* ctx; bool on;
    SetGuard(bg_ctx* c, bool use_alt) : ctx(c), on(use_alt) { if (on) swap_set(ctx); }
    ~SetGuard() { if (on) swap_set(ctx); }
};
static const int OVERLAP_WARPS_PER_SM = 16;

template <typename T> static int ensure(bg_ctx* ctx, T** p, size_t* cap, size_t need) {
    if (!(*cap >= need && *p)) drop_graph(ctx);        // a captured graph holds the old pointers
    if (*cap >= need && *p) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    size_t want = need + need / 4 + 16;
    cudaError_t e = cudaMalloc((void**)p, want * sizeof(T));
    if (e != cudaSuccess) return fail(ctx, "cudaMalloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e));
    *cap = want;
    return 0;
}

extern "C" const char* bg_last_error(const bg_ctx* ctx) {
    if (ctx && !ctx->err.empty()) return ctx->err.c_str();
    return g_last_error.c_str();
}

extern "C" int bg_device_count(int* out) {
    if (!out) return fail(nullptr, "bg_device_count: null out");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        *out = 0;
        return fail(nullptr, "no CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    *out = ndev;
    return 0;
}

extern "C" int bg_init(bg_ctx** out, int device) {
    if (!out) return fail(nullptr, "bg_init: null out pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, "bg_init: no CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return fail(nullptr, "bg_init: device %d out of range (0..%d)", device, ndev - 1);
    bg_ctx* ctx = new bg_ctx();
    ctx->device = device;
    memset(&ctx->stats, 0, sizeof ctx->stats);
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return fail(nullptr, "cudaSetDevice(%d) failed", device); }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    ctx->sm_count = prop.multiProcessorCount;
    if (prop.major != 10 || prop.minor != 0) {      // the library carries sm_100a SASS only (no PTX)
        int r = fail(nullptr, "bg_init: device %d is sm_%d%d; this build targets sm_100a (B200) only", device, prop.major, prop.minor);
        delete ctx; return r;
    }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return fail(nullptr, "cudaStreamCreate failed"); }
    for (int sl = 0; sl < 2; sl++) {
        cudaEventCreate(&ctx->ev0s[sl]); cudaEventCreate(&ctx->ev1s[sl]);
        cudaEventCreateWithFlags(&ctx->ev_done[sl], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_out[sl], cudaEventDisableTiming);
        for (int a = 0; a < 2; a++) for (int b = 0; b < 3; b++) cudaEventCreate(&ctx->evps[sl][a][b]);
    }
    cudaStreamCreateWithFlags(&ctx->cstream, cudaStreamNonBlocking);
    if (cudaStreamCreateWithFlags(&ctx->alt.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void**)&ctx->alt.d_P, 2 * sizeof(bg_projector)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->alt.d_partials, 64 * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->alt.d_ticket, 2 * sizeof(unsigned int)) != cudaSuccess) {
        int r = fail(nullptr, "bg_init: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete ctx; return r;
    }
    cudaMemset(ctx->alt.d_ticket, 0, 2 * sizeof(unsigned int));
    if (cudaMalloc((void**)&ctx->d_P, 2 * sizeof(bg_projector)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_counters, 16 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaHostAlloc((void**)&ctx->h_out, 32 * sizeof(double), cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void**)&ctx->h_stage, 4 * sizeof(bg_projector), cudaHostAllocDefault) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_stage[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_stage[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_red, 16 * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_partials, 64 * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_ticket, 2 * sizeof(unsigned int)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_shb_perm, sizeof(ShbPerm)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_cdf, (BG_MAX_T + 1) * sizeof(double)) != cudaSuccess) {
        int r = fail(nullptr, "bg_init: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete ctx; return r;
    }
    if (const char* e1 = getenv("BG_CTAS_PER_SM")) { int v = atoi(e1); if (v >= 1 && v <= 16) ctx->ctas_per_sm = v; }
    cudaMemset(ctx->d_red, 0, 16 * sizeof(double));
    cudaMemset(ctx->d_ticket, 0, 2 * sizeof(unsigned int));
    cudaMemset(ctx->d_counters, 0, 16 * sizeof(unsigned long long));
    if (const char* e5 = getenv("BG_GRAPH")) ctx->use_graph = atoi(e5) != 0;
    if (const char* e6 = getenv("BG_LAM_MAX")) ctx->lam_max = std::max(0, std::min(TPP_MAXC, atoi(e6)));
    if (const char* e7 = getenv("BG_FUSE2")) ctx->fuse2 = atoi(e7) != 0;
    if (const char* e8 = getenv("BG_SHB")) ctx->use_shb = atoi(e8) != 0;
    if (const char* e3 = getenv("BG_KERNEL")) ctx->force_warp = strcmp(e3, "warp") == 0;
    if (const char* e9 = getenv("BG_PREP")) ctx->prep_warp = strcmp(e9, "warp") == 0;
    if (const char* e10 = getenv("BG_OVERLAP")) ctx->overlap_mode = atoi(e10) != 0;
    if (const char* e11 = getenv("BG_PIECES")) ctx->pieces_override = std::max(0, atoi(e11));
    if (const char* e2 = getenv("BG_ITEMS_FACTOR")) { int v = atoi(e2); if (v >= 1 && v <= 1024) { ctx->items_factor = v; ctx->items_factor_set = true; } }
    *out = ctx;
    return 0;
}

extern "C" void bg_shutdown(bg_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->nccl_comm);
    cudaFree(ctx->d_shb_perm);
    cudaFree(ctx->d_terms); cudaFree(ctx->d_terms_shb); cudaFree(ctx->d_term_nat_shb); cudaFree(ctx->d_terms_sorted); cudaFree(ctx->d_term_nat); cudaFree(ctx->d_cdf); cudaFree(ctx->d_recs); cudaFree(ctx->d_zw); cudaFree(ctx->d_zw2);
    cudaFree(ctx->d_per); cudaFree(ctx->d_per2); cudaFree(ctx->d_P); cudaFree(ctx->d_counters); cudaFree(ctx->d_red); cudaFree(ctx->d_partials); cudaFree(ctx->d_ticket);
    for (int sl = 0; sl < 2; sl++) {
        if (ctx->ev0s[sl]) cudaEventDestroy(ctx->ev0s[sl]);
        if (ctx->ev1s[sl]) cudaEventDestroy(ctx->ev1s[sl]);
        if (ctx->ev_done[sl]) cudaEventDestroy(ctx->ev_done[sl]);
        if (ctx->ev_out[sl]) cudaEventDestroy(ctx->ev_out[sl]);
        for (int a = 0; a < 2; a++) for (int b = 0; b < 3; b++) if (ctx->evps[sl][a][b]) cudaEventDestroy(ctx->evps[sl][a][b]);
        if (ctx->gexecs[sl]) cudaGraphExecDestroy(ctx->gexecs[sl]);
    }
    if (ctx->alt_in) swap_set(ctx);
    cudaFree(ctx->alt.d_recs); cudaFree(ctx->alt.d_zw); cudaFree(ctx->alt.d_zw2); cudaFree(ctx->alt.d_per); cudaFree(ctx->alt.d_per2);
    cudaFree(ctx->alt.d_P); cudaFree(ctx->alt.d_partials); cudaFree(ctx->alt.d_ticket);
    if (ctx->alt.stream) cudaStreamDestroy(ctx->alt.stream);
    if (ctx->cstream) cudaStreamDestroy(ctx->cstream);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    for (int j = 0; j < 2; j++) if (ctx->ev_stage[j]) cudaEventDestroy(ctx->ev_stage[j]);
    if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int bg_set_shard(bg_ctx* ctx, int rank, int world) {
    if (!ctx) return fail(nullptr, "bg_set_shard: null ctx");
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, "bg_set_shard: bad rank %d of %d", rank, world);
    ctx->rank = rank; ctx->world = world;
    ctx->prepared = false;
    drop_graph(ctx);
    return 0;
}

extern "C" int bg_set_stream(bg_ctx* ctx, void* stream) {
    if (!ctx) return fail(nullptr, "bg_set_stream: null ctx");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) { cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    ctx->stream = (cudaStream_t)stream;
    drop_graph(ctx);
    return 0;
}

extern "C" int bg_set_allreduce(bg_ctx* ctx, int enabled) {
    if (!ctx) return fail(nullptr, "bg_set_allreduce: null ctx");
    ctx->allreduce = enabled != 0;
    return 0;
}

// ---- NCCL (loaded lazily so that single-GPU use never needs it) ---------------------------
static int nccl_load(bg_ctx* ctx) {
    if (g_nccl.handle) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return fail(ctx, "NCCL not found (dlopen libnccl.so.2): %s", dlerror());
    g_nccl.handle = h;
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, ...))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return fail(ctx, "NCCL symbols missing in libnccl");
    return 0;
}

extern "C" int bg_nccl_unique_id(uint8_t id[128]) {
    if (nccl_load(nullptr)) return 1;
    NcclUniqueId u; memset(&u, 0, sizeof u);
    int r = g_nccl.GetUniqueId(&u);
    if (r != 0) return fail(nullptr, "ncclGetUniqueId failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    memcpy(id, &u, 128);
    return 0;
}

extern "C" int bg_nccl_join(bg_ctx* ctx, const uint8_t id[128]) {
    if (!ctx) return fail(nullptr, "bg_nccl_join: null ctx");
    if (ctx->world <= 1) return 0;
    if (nccl_load(ctx)) return 1;
    CK(cudaSetDevice(ctx->device));
    NcclUniqueId u; memcpy(&u, id, 128);
    int r = ((nccl_init_rank_fn)g_nccl.CommInitRank)(&ctx->nccl_comm, ctx->world, u, ctx->rank);
    if (r != 0) return fail(ctx, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return 0;
}

// all-reduce (sum) n doubles of the current slot's outputs; no-op for world == 1
static int allreduce_red(bg_ctx* ctx, int n, cudaStream_t stream = nullptr) {
    if (!stream) stream = ctx->stream;
    if (ctx->world <= 1 || !ctx->allreduce) return 0;
    if (!ctx->nccl_comm) return fail(ctx, "world = %d but bg_nccl_join was not called", ctx->world);
    int r = g_nccl.AllReduce(RED(ctx), RED(ctx), (size_t)n, /*ncclDouble*/ 8, /*ncclSum*/ 0, ctx->nccl_comm, stream);
    if (r != 0) return fail(ctx, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return 0;
}

// ---- decomposition ------------------------------------------------------------------------
// binrep (stateprep.c:5-32) is MSB-first: character j of the string is bit (size-1-j) of i
static inline int binbit(uint64_t val, int sz, int j) { return (int)((val >> (sz - 1 - j)) & 1ull); }

// cumulative distribution of d = n - k, stabilizer.c:677-716
static void dimension_cdf(int n, double* cumulative) {
    double dist[BG_MAX_T + 1], sum = 0;
    for (int d = 0; d <= n; d++) {
        double le = 0.;
        if (d > 0) {
            double product = 0;
            for (int a = 1; a <= d; a++) { product += log2(1 - pow(2, d - n - a)); product -= log2(1 - pow(2, -a)); }
            le = (-d * (d + 1) / 2) + product;
        }
        dist[d] = pow(2, le); sum += dist[d];
    }
    for (int d = 0; d <= n; d++) dist[d] /= sum;
    for (int i = 0; i <= n; i++) { cumulative[i] = 0; for (int d = 0; d <= i; d++) cumulative[i] += dist[d]; }
    for (int i = 0; i <= n; i++) cumulative[i] /= cumulative[n];
}

extern "C" int bg_set_decomposition(bg_ctx* ctx, int t, int exact, int k, const uint64_t* L_rows) {
    if (!ctx) return fail(nullptr, "bg_set_decomposition: null ctx");
    if (t < 1 || t > BG_MAX_T) return fail(ctx, "bg_set_decomposition: t = %d outside 1..%d", t, BG_MAX_T);
    CK(cudaSetDevice(ctx->device));
    if (ctx->t == t && ctx->exact == (exact ? 1 : 0) && !ctx->terms_host.empty()) {
        // the same decomposition as the one in place (the back end sets it once per probability() call): keep the
        // term tables, the plan, the prepared job and its captured graph
        const uint64_t maskt = t >= 64 ? ~0ull : ((1ull << t) - 1);
        bool same = exact ? true : (k == ctx->k && (k == 0 || L_rows != nullptr));
        for (int j = 0; same && !exact && j < k; j++) same = (L_rows[j] & maskt) == ctx->L[j];
        if (same) return 0;
    }
    CK(cudaStreamSynchronize(ctx->alt.stream));     // a job on the alternate stream may still be reading the old tables
    size_t chi;
    if (exact) {
        const int size = (t + 1) / 2;
        if (size > 26) return fail(ctx, "bg_set_decomposition: exact decomposition with 2^%d terms is too large", size);
        chi = (size_t)1 << size;
        ctx->terms_host.resize(chi);
        for (size_t i = 0; i < chi; i++) {
            uint64_t e1 = 0;
            for (int j = 0; j < size; j++) if (binbit(i, size, j)) e1 |= 1ull << (2 * j);
            ctx->terms_host[i] = e1;
        }
        ctx->L.clear(); ctx->k = 0;
    } else {
        if (k < 0 || k > 26 || k > t) return fail(ctx, "bg_set_decomposition: k = %d outside 0..min(t,26)", k);
        if (k > 0 && !L_rows) return fail(ctx, "bg_set_decomposition: L_rows is null");
        const uint64_t maskt = t >= 64 ? ~0ull : ((1ull << t) - 1);
        ctx->L.assign(L_rows, L_rows + k);
        for (auto& r : ctx->L) r &= maskt;
        chi = (size_t)1 << k;
        ctx->terms_host.resize(chi);
        for (size_t i = 0; i < chi; i++) {         // x~_i = xor of the rows of L picked by binrep(i) (stateprep.c:87-103)
            uint64_t x = 0;
            for (int j = 0; j < k; j++) if (binbit(i, k, j)) x ^= ctx->L[j];
            ctx->terms_host[i] = x;
        }
        ctx->k = k;
    }
    ctx->t = t; ctx->exact = exact ? 1 : 0;
    size_t padded = (chi + 1) & ~(size_t)1;        // 16-byte multiple for the bulk copy
    if (ensure(ctx, &ctx->d_terms, &ctx->d_terms_cap, padded)) return 1;
    CK(cudaMemsetAsync(ctx->d_terms, 0, padded * 8, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_terms, ctx->terms_host.data(), chi * 8, cudaMemcpyHostToDevice, ctx->stream));
    {   // popcount-sorted copy for the pair kernels (stable: ties keep the natural order)
        std::vector<int32_t> nat(chi);
        const std::vector<uint64_t>& th = ctx->terms_host;
        // key: popcount, then popcount of the low 32-bit half (the row loops walk the two halves separately), largest
        // first; one sort of composite keys (inverted key | index; chi <= 2^26)
        {
            std::vector<uint64_t> keys(chi);
            for (size_t i = 0; i < chi; i++) {
                const uint64_t kk = ((uint64_t)__builtin_popcountll(th[i]) << 8) | (uint64_t)__builtin_popcountll(th[i] & 0xffffffffull);
                keys[i] = ((0xffffull - kk) << 32) | (uint64_t)i;
            }
            std::sort(keys.begin(), keys.end());
            for (size_t i = 0; i < chi; i++) nat[i] = (int32_t)(keys[i] & 0xffffffffull);
        }
        std::vector<uint64_t> sorted(padded, 0);
        for (size_t i = 0; i < chi; i++) sorted[i] = th[nat[i]];
        if (ensure(ctx, &ctx->d_terms_sorted, &ctx->d_terms_sorted_cap, padded)) return 1;
        if (ensure(ctx, &ctx->d_term_nat, &ctx->d_term_nat_cap, chi)) return 1;
        CK(cudaMemcpyAsync(ctx->d_terms_sorted, sorted.data(), padded * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_term_nat, nat.data(), chi * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));       // the host vectors go out of scope
    }
    ctx->shb_plan = ShbPlan();
    if (!exact && ctx->use_shb && !ctx->force_warp && t > 32 && t <= 32 + SHB_MAXH && chi >= 32 * SHB_G && chi % (32 * SHB_G) == 0 && chi <= ((size_t)1 << 20)) {
        ctx->shb_plan = shb_make_plan(t, k, ctx->L, ctx->terms_host);
        if (ctx->shb_plan.ok) {
            std::vector<uint64_t> tp(padded, 0);
            std::copy(ctx->shb_plan.terms.begin(), ctx->shb_plan.terms.end(), tp.begin());
            if (ensure(ctx, &ctx->d_terms_shb, &ctx->d_terms_shb_cap, padded)) return 1;
            if (ensure(ctx, &ctx->d_term_nat_shb, &ctx->d_term_nat_shb_cap, chi)) return 1;
            CK(cudaMemcpyAsync(ctx->d_terms_shb, tp.data(), padded * 8, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->d_term_nat_shb, ctx->shb_plan.nat.data(), chi * 4, cudaMemcpyHostToDevice, ctx->stream));
            ShbPerm pm; memset(&pm, 0, sizeof pm);
            pm.nh = ctx->shb_plan.nh; pm.nsw = ctx->shb_plan.nsw;
            for (int i = 0; i < SHB_MAXH; i++) { pm.swp[i] = ctx->shb_plan.swp[i]; pm.swq[i] = ctx->shb_plan.swq[i]; }
            for (int i = 0; i < 64; i++) pm.iperm[i] = ctx->shb_plan.iperm[i];
            CK(cudaMemcpyAsync(ctx->d_shb_perm, &pm, sizeof pm, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    if (ctx->cdf_t != t) {
        double cdf[BG_MAX_T + 1];
        dimension_cdf(t, cdf);
        CK(cudaMemcpyAsync(ctx->d_cdf, cdf, (t + 1) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ctx->cdf_t = t;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    // The captured graphs hold buffer addresses (ensure() drops them when a buffer moves) and launch shapes that depend
    // on (t, exact, k, chi) and on whether there is a shared high-block plan — not on the content of the tables.  A new L
    // of the same shape (what sampleQubits does per probability() call) keeps the prepared job and its graphs.
    const bool same_sig = ctx->sig_t == t && ctx->sig_exact == (exact ? 1 : 0) && ctx->sig_k == ctx->k && ctx->sig_chi == chi &&
                          ctx->sig_plan_ok == ctx->shb_plan.ok && ctx->sig_nh == ctx->shb_plan.nh;
    ctx->sig_t = t; ctx->sig_exact = exact ? 1 : 0; ctx->sig_k = ctx->k; ctx->sig_chi = chi;
    ctx->sig_plan_ok = ctx->shb_plan.ok; ctx->sig_nh = ctx->shb_plan.nh;
    if (!same_sig) { ctx->prepared = false; drop_graph(ctx); }
    return 0;
}

// reference bit layouts: bit `loc` of a BitVector/BitMatrix is (data[loc/8] >> (7 - loc%8)) & 1
static inline int refbit(const uint8_t* data, size_t loc) { return (data[loc / 8] >> (7 - loc % 8)) & 1; }

extern "C" int bg_set_decomposition_bitmatrix(bg_ctx* ctx, int t, int exact, int k, const uint8_t* L_bits) {
    std::vector<uint64_t> rows;
    if (!exact) {
        if (k > 0 && !L_bits) return fail(ctx, "bg_set_decomposition_bitmatrix: L_bits is null");
        if (t < 1 || t > BG_MAX_T) return fail(ctx, "bg_set_decomposition_bitmatrix: t = %d outside 1..%d", t, BG_MAX_T);
        rows.assign(k > 0 ? k : 0, 0);
        for (int r = 0; r < k; r++)
            for (int c = 0; c < t; c++) if (refbit(L_bits, (size_t)r * t + c)) rows[r] |= 1ull << c;
    }
    return bg_set_decomposition(ctx, t, exact, k, rows.data());
}

extern "C" int bg_projector_from_bitmatrix(bg_projector* out, int nstabs, int nqubits, const uint8_t* phase_sign,
                                           const uint8_t* phase_complex, const uint8_t* xs, const uint8_t* zs) {
    if (!out) return fail(nullptr, "bg_projector_from_bitmatrix: null out");
    if (nstabs < 0 || nstabs > BG_MAX_STABS) return fail(nullptr, "projector with %d generators (max %d)", nstabs, BG_MAX_STABS);
    if (nqubits < 0 || nqubits > BG_MAX_T) return fail(nullptr, "projector on %d qubits (max %d)", nqubits, BG_MAX_T);
    memset(out, 0, sizeof *out);
    out->nstabs = nstabs; out->nqubits = nqubits;
    for (int i = 0; i < nstabs; i++) {
        out->phase[i] = (uint8_t)(2 * refbit(phase_sign, i) + refbit(phase_complex, i));
        for (int q = 0; q < nqubits; q++) {
            if (refbit(xs, (size_t)i * nqubits + q)) out->xs[i] |= 1ull << q;
            if (refbit(zs, (size_t)i * nqubits + q)) out->zs[i] |= 1ull << q;
        }
    }
    return 0;
}

// ---- launch helpers -------------------------------------------------------------------------
static const int WARPS_PER_BLOCK = 4;
static const size_t SMEM_TERMS_MAX = 8192;     // 64 KB of staged terms

template <int NS> static int launch_prepare_ns(bg_ctx* ctx, int src, const PrepArgs& a) {
    const int blocks = std::max(1, std::min((a.n_samples + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, ctx->sm_count * 16));
    if (src == SRC_RNG) k_prepare<NS, SRC_RNG><<<blocks, 32 * WARPS_PER_BLOCK, 0, ctx->stream>>>(a);
    else if (src == SRC_STATES) k_prepare<NS, SRC_STATES><<<blocks, 32 * WARPS_PER_BLOCK, 0, ctx->stream>>>(a);
    else k_prepare<NS, SRC_TERMS><<<blocks, 32 * WARPS_PER_BLOCK, 0, ctx->stream>>>(a);
    CK(cudaGetLastError());
    ctx->stats.launches++;
    return 0;
}
// k_prepare_tps: as many CTAs as are resident at once (occupancy x SMs), at most one warp per group of 32 samples
template <typename W, int THREADS> static int launch_prepare_tps(bg_ctx* ctx, const PrepArgs& a) {
    const size_t smem = prep_tps_smem<W, THREADS>(a.t);
    int& per_sm = ctx->prep_ctas_per_sm[sizeof(W) == 8][THREADS == 32][a.t];
    if (per_sm == 0) {
        // the attribute is per kernel, not per launch: ask once for what the widest state of this word size needs (a
        // smaller t seen later must not lower it under a larger t whose occupancy is already cached)
        const size_t smem_max = prep_tps_smem<W, THREADS>(sizeof(W) == 8 ? BG_MAX_T : 32);
        CK(cudaFuncSetAttribute(k_prepare_tps<W, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_prepare_tps<W, THREADS>, THREADS, smem));
        if (per_sm < 1) return fail(ctx, "k_prepare_tps does not fit an SM at t = %d", a.t);
    }
    const int groups = (a.n_samples + 31) / 32, wpb = THREADS / 32;
    const int blocks = std::max(1, std::min((groups + wpb - 1) / wpb, ctx->sm_count * per_sm));
    k_prepare_tps<W, THREADS><<<blocks, THREADS, smem, ctx->stream>>>(a);
    CK(cudaGetLastError());
    ctx->stats.launches++;
    return 0;
}
static int launch_prepare(bg_ctx* ctx, int src, PrepArgs a) {
    if (a.n_samples <= 0) return 0;
    CK(cudaMemsetAsync(CNT(ctx) + 4 * ctx->cur, 0, 4 * sizeof(unsigned long long), ctx->stream));
    a.force_warp = ctx->force_warp;
    a.shb = (ctx->shb_plan.ok && src != SRC_TERMS) ? 1 : 0;     // SRC_TERMS = the exact-norm path (tri mode): generic kernels
    if (!a.P2) a.n_first = a.n_samples;
    a.n_warp_routed = CNT(ctx) + 4 * ctx->cur + 2;
    if (src == SRC_RNG && a.project && !a.raw_out && !ctx->prep_warp) {
        if (ctx->overlap) return a.t <= 32 ? launch_prepare_tps<uint32_t, 32>(ctx, a) : launch_prepare_tps<uint64_t, 32>(ctx, a);
        return a.t <= 32 ? launch_prepare_tps<uint32_t, 64>(ctx, a) : launch_prepare_tps<uint64_t, 64>(ctx, a);
    }
    return a.t <= 32 ? launch_prepare_ns<1>(ctx, src, a) : launch_prepare_ns<2>(ctx, src, a);
}

template <int NS, bool EXACT> static int launch_pairs_ns(bg_ctx* ctx, const PairArgs& a, int blocks, size_t smem) {
    if (smem > 48 * 1024)
        CK(cudaFuncSetAttribute(k_pairs<NS, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM_TERMS_MAX * 8)));
    k_pairs<NS, EXACT><<<blocks, 32 * WARPS_PER_BLOCK, smem, ctx->stream>>>(a);
    CK(cudaGetLastError());
    ctx->stats.launches++;
    return 0;
}

template <typename W, bool EXACT, bool TRI, bool MANYC>
static int launch_tpp_inst(bg_ctx* ctx, const PairArgs& a, int blocks, size_t smem) {
    if (smem > 48 * 1024)
        CK(cudaFuncSetAttribute(k_pairs_tpp<W, EXACT, TRI, MANYC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // persistent grid: exactly the CTAs that are resident at once (a multiple of the SM count), fewer if
    // there are not that many work items
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pairs_tpp<W, EXACT, TRI, MANYC>, 32 * ctx->tpp_warps, smem));
    if (per_sm < 1) return fail(ctx, "k_pairs_tpp: a CTA of %d warps with %zu bytes of shared memory does not fit an SM", ctx->tpp_warps, smem);
    blocks = std::min(blocks, ctx->sm_count * per_sm);
    k_pairs_tpp<W, EXACT, TRI, MANYC><<<blocks, 32 * ctx->tpp_warps, smem, ctx->stream>>>(a);
    CK(cudaGetLastError());
    ctx->stats.launches++;
    return 0;
}
template <typename W, bool MANYC> static int launch_tpp_w(bg_ctx* ctx, const PairArgs& a, int blocks, size_t smem) {
    if (ctx->exact) return a.tri ? launch_tpp_inst<W, true, true, MANYC>(ctx, a, blocks, smem) : launch_tpp_inst<W, true, false, MANYC>(ctx, a, blocks, smem);
    return a.tri ? launch_tpp_inst<W, false, true, MANYC>(ctx, a, blocks, smem) : launch_tpp_inst<W, false, false, MANYC>(ctx, a, blocks, smem);
}

// Fill in chunking / staging and launch the pair kernels: k_pairs_tpp for the samples routed to it,
// then k_pairs (returns at once when nothing was routed to it).  zw (and zw2) must be zeroed and
// launch_prepare must have run on the same stream (it resets the counters).
static int launch_pairs(bg_ctx* ctx, PairArgs a) {
    if (a.n_samples <= 0) return 0;
    const int resident_warps = ctx->sm_count * ctx->ctas_per_sm * WARPS_PER_BLOCK;
    // the last quarter of the (popcount-sorted, i.e. cheapest) terms goes out in 32-term items
    // (only when samples are scarce: with plenty of samples the big items balance by themselves)
    const bool shb = ctx->shb_plan.ok && !a.tri && !ctx->force_warp;       // terms beyond 2048 are read from global memory (L2)
    const int gran = shb ? 32 * SHB_G : 32;                    // items are whole groups of terms
    // k_pairs_shb relabels the sample once per item (about half a batch of work): one item per resident warp is
    // enough there (measured at 2 x 8192 samples: 0.80 ms with 1, 0.92 ms with 8 items per warp)
    const int want_items = resident_warps * (shb && !ctx->items_factor_set ? 1 : ctx->items_factor);
    const int tail_terms = (a.nterms >= 128 && a.n_samples < want_items) ? (((a.nterms / 4) + gran - 1) / gran * gran) : 0;
    a.tail_start = a.nterms - tail_terms;
    a.tail_size = gran;
    a.tail_chunks = (tail_terms + gran - 1) / gran;
    int cps = 1;
    if (a.n_samples < want_items) cps = (want_items + a.n_samples - 1) / a.n_samples;
    int chunk = (a.tail_start + cps - 1) / cps;
    chunk = (chunk + gran - 1) / gran * gran;
    if (chunk < gran) chunk = gran;
    cps = (a.tail_start + chunk - 1) / chunk;
    a.chunk = chunk; a.chunks_per_sample = cps;
    const size_t padded = ((size_t)a.nterms + 1) & ~(size_t)1;
    const unsigned long long items = (unsigned long long)a.n_samples * (cps + a.tail_chunks);
    long long blocks = (long long)ctx->sm_count * ctx->ctas_per_sm;   // persistent CTAs, a multiple of the SM count
    const long long need = (long long)((items + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    unsigned long long* cnt = CNT(ctx) + 4 * ctx->cur;
    a.pair_count = cnt + 1;
    a.n_warp_routed = cnt + 2;
    a.terms = a.tri ? ctx->d_terms : ctx->d_terms_sorted;
    a.term_nat = ctx->d_term_nat;
    a.natural = a.tri ? 1 : 0;
    if (!ctx->force_warp) {
        a.smem_terms = padded <= 2048 ? (int)padded : 0;
        const size_t wb = a.t <= 32 ? 4 : 8;
        const int tw = ctx->tpp_warps;
        long long tb = (long long)ctx->sm_count * 64;            // capped to the resident CTAs in launch_tpp_inst
        const long long tneed = (long long)((items + tw - 1) / tw);
        if (tb > tneed) tb = tneed;
        if (tb < 1) tb = 1;
        a.counter = cnt;
        if (shb) {
            // shared high-block reduction: relabelled pattern-sorted terms, 32-bit rows (samples with <= SHB_MAXLAM checks)
            PairArgs b = a;
            b.terms = ctx->d_terms_shb; b.term_nat = ctx->d_term_nat_shb;
            b.shb = ctx->d_shb_perm;
            {   // whole samples first; the last wave's worth of samples in pieces
                const int groups = b.nterms / gran;
                const int rw = ctx->sm_count * 30;                               // resident warps of k_pairs_shb (10 CTAs x 3 warps)
                int pieces = 4;
                while ((long long)std::min(b.n_samples, rw) * pieces < 2LL * rw && pieces < groups) pieces *= 2;
                if (ctx->pieces_override > 0) pieces = ctx->pieces_override;
                pieces = std::max(1, std::min(pieces, groups));
                b.piece = (groups + pieces - 1) / pieces * gran;
                b.pieces = (b.nterms + b.piece - 1) / b.piece;
                b.split_from = std::max(0, b.n_samples - rw);
                if ((long long)(b.n_samples - b.split_from) * b.pieces >= (1LL << 31)) return fail(ctx, "k_pairs_shb: too many work items");
                if (ctx->items_factor_set) { b.split_from = 0; }                 // BG_ITEMS_FACTOR: every sample in pieces
            }
            const size_t smem = (size_t)b.smem_terms * 8 + (size_t)tw * SHB_WARP_WORDS * 4 + (size_t)32 * 32 * tw * 4;
            if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_pairs_shb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pairs_shb, 32 * tw, smem));
            if (per_sm < 1) return fail(ctx, "k_pairs_shb does not fit an SM (%zu bytes of shared memory)", smem);
            const int blocks_shb = (int)std::min<long long>(tb, (long long)ctx->sm_count * per_sm);
            k_pairs_shb<<<blocks_shb, 32 * tw, smem, ctx->stream>>>(b);
            CK(cudaGetLastError());
            ctx->stats.launches++;
        } else {
            // samples with <= TPP_MAXC parity checks
            a.lam_max = std::max(0, std::min(ctx->lam_max, (int)(8 * wb) - a.t));
            size_t smem = (size_t)a.smem_terms * 8 + (size_t)tw * tpp_amb_rows(a.t, a.lam_max, false) * wb
                          + (size_t)tpp_work_rows(a.t, a.lam_max, false) * 32 * tw * wb;
            if (a.t <= 32) { if (launch_tpp_w<uint32_t, false>(ctx, a, (int)tb, smem)) return 1; }
            else { if (launch_tpp_w<uint64_t, false>(ctx, a, (int)tb, smem)) return 1; }
        }
        // the rest (returns at once if there is none)
        a.counter = cnt + 3;
        a.lam_max = 0;
        const size_t smem = (size_t)a.smem_terms * 8 + (size_t)tw * tpp_amb_rows(a.t, 0, true) * wb + (size_t)tpp_work_rows(a.t, 0, true) * 32 * tw * wb;
        if (a.t <= 32) return launch_tpp_w<uint32_t, true>(ctx, a, (int)tb, smem);
        return launch_tpp_w<uint64_t, true>(ctx, a, (int)tb, smem);
    }
    a.counter = cnt + 3;
    a.smem_terms = padded <= SMEM_TERMS_MAX ? (int)padded : 0;
    const size_t smem = (size_t)a.smem_terms * 8;
    if (a.t <= 32) return ctx->exact ? launch_pairs_ns<1, true>(ctx, a, (int)blocks, smem) : launch_pairs_ns<1, false>(ctx, a, (int)blocks, smem);
    return ctx->exact ? launch_pairs_ns<2, true>(ctx, a, (int)blocks, smem) : launch_pairs_ns<2, false>(ctx, a, (int)blocks, smem);
}

static int ensure_sample_buffers(bg_ctx* ctx, size_t n) {
    if (ensure(ctx, &ctx->d_recs, &ctx->recs_cap, n)) return 1;
    if (ensure(ctx, &ctx->d_zw, &ctx->zw_cap, n * 4)) return 1;
    if (ensure(ctx, &ctx->d_zw2, &ctx->zw2_cap, n * 4)) return 1;
    if (ensure(ctx, &ctx->d_per, &ctx->per_cap, n)) return 1;
    if (ensure(ctx, &ctx->d_per2, &ctx->per2_cap, n)) return 1;
    return 0;
}

// number of sample indices l in [0, total) with l % world == rank
static inline uint64_t shard_count(uint64_t total, int rank, int world) {
    return total / world + ((uint64_t)rank < total % world ? 1 : 0);
}

// ---- sampled norm -------------------------------------------------------------------------
static int check_projector(bg_ctx* ctx, const bg_projector* P) {
    if (!P) return fail(ctx, "null projector");
    if (P->nstabs < 0 || P->nstabs > BG_MAX_STABS) return fail(ctx, "projector with %d generators (max %d)", P->nstabs, BG_MAX_STABS);
    if (P->nstabs > 0 && P->nqubits != ctx->t)
        return fail(ctx, "projector acts on %d qubits but the decomposition has t = %d", P->nqubits, ctx->t);
    return 0;
}

// Clifford circuit (t == 0): closed form of innerprod.c:52-62 / 157-167
static double clifford_closed_form(const bg_projector* P) {
    double sum = 1;
    for (int i = 0; i < P->nstabs; i++) { if (P->phase[i] == 0) sum += 1; if (P->phase[i] == 2) sum -= 1; }
    return sum / (1 + (double)P->nstabs);
}

// d_P of the set that is swapped in := the prepared job's projectors, through a pinned staging copy, asynchronously on
// that set's stream (ordered behind the job that may still be reading d_P there); the staging copies alternate, so the
// projectors of the next job can be staged while the previous upload is still in flight.
static int sync_projectors(bg_ctx* ctx) {
    const int nproj = ctx->nproj;
    bool same = ctx->P_valid;
    for (int j = 0; same && j < nproj; j++) same = memcmp(&ctx->h_P[j], &ctx->cur_P[j], sizeof(bg_projector)) == 0;
    if (same) return 0;
    const unsigned ss = ctx->stage_seq++ & 1u;
    CK(cudaEventSynchronize(ctx->ev_stage[ss]));     // the upload before last has left this staging set
    bg_projector* hs = ctx->h_stage + 2 * ss;
    for (int j = 0; j < nproj; j++) hs[j] = ctx->cur_P[j];
    CK(cudaMemcpyAsync(ctx->d_P, hs, (size_t)nproj * sizeof(bg_projector), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->ev_stage[ss], ctx->stream));
    for (int j = 0; j < nproj; j++) ctx->h_P[j] = ctx->cur_P[j];
    ctx->P_valid = true;
    ctx->stats.h2d_bytes += (uint64_t)nproj * sizeof(bg_projector);
    return 0;
}

static int sampled_prepare_n(bg_ctx* ctx, int nproj, const bg_projector* const* Ps, uint64_t samples, int bins,
                             const uint64_t* seeds) {
    if (!ctx) return fail(nullptr, "bg_sampled_prepare: null ctx");
    if (ctx->t <= 0) return fail(ctx, "bg_sampled_prepare: call bg_set_decomposition first");
    for (int j = 0; j < nproj; j++) {
        if (check_projector(ctx, Ps[j])) return 1;
        if (Ps[j]->nstabs == 0) return fail(ctx, "bg_sampled_prepare: empty projector (use bg_sampled_norm for the closed form)");
    }
    if (bins < 1) return fail(ctx, "bg_sampled_prepare: bins = %d", bins);
    if (samples < 1) return fail(ctx, "bg_sampled_prepare: samples = 0");
    CK(cudaSetDevice(ctx->device));
    const uint64_t mine = shard_count(samples, ctx->rank, ctx->world);
    if (mine > (1ull << 30)) return fail(ctx, "bg_sampled_prepare: %llu samples per rank is too many", (unsigned long long)mine);
    // a job of the same shape and seeds keeps its captured graph (the kernels read d_P when they run)
    bool same_shape = ctx->prepared && ctx->nproj == nproj && ctx->samples == samples && ctx->bins == bins;
    for (int j = 0; same_shape && j < nproj; j++) same_shape = ctx->seeds[j] == seeds[j];
    // overlap mode: a fused two-projector job whose draw + projection is less than OVERLAP_WARPS_PER_SM warps per SM
    const bool fusable = ctx->fuse2 && nproj == 2 && bins == 1 && ctx->use_graph && !ctx->prep_warp && !ctx->force_warp;
    const bool want_overlap = fusable && (ctx->overlap_mode == 1 ||
        (ctx->overlap_mode < 0 && 2 * mine <= (uint64_t)32 * OVERLAP_WARPS_PER_SM * (uint64_t)ctx->sm_count));
    if ((!same_shape || want_overlap != ctx->overlap) && ctx->run_seq != ctx->fin_seq) {
        // an unfinished job of another shape: its buffers are about to change
        CK(cudaStreamSynchronize(ctx->stream)); CK(cudaStreamSynchronize(ctx->alt.stream)); CK(cudaStreamSynchronize(ctx->cstream));
        ctx->fin_seq = ctx->run_seq;
    }
    if (want_overlap != ctx->overlap) { drop_graph(ctx); ctx->overlap = want_overlap; }
    const size_t nrec = (size_t)std::max<uint64_t>(mine, 1) * (size_t)nproj;
    if (!same_shape && ensure_sample_buffers(ctx, nrec)) return 1;
    if (ctx->overlap) { SetGuard g(ctx, true); if (ensure_sample_buffers(ctx, nrec)) return 1; }
    ctx->nproj = nproj; ctx->samples = samples; ctx->bins = bins;
    for (int j = 0; j < nproj; j++) { ctx->seeds[j] = seeds[j]; ctx->cur_P[j] = *Ps[j]; }
    // the projectors go to the set the next bg_sampled_run uses (bg_sampled_run checks again: a job that is run many
    // times alternates between the sets)
    ctx->stats.h2d_bytes = 0;
    {
        SetGuard g(ctx, ctx->overlap && (ctx->run_seq & 1u));
        if (sync_projectors(ctx)) return 1;
    }
    ctx->prepared = true; ctx->per_valid = false;
    if (!same_shape) drop_graph(ctx);
    return 0;
}

extern "C" int bg_sampled_prepare(bg_ctx* ctx, const bg_projector* P, uint64_t samples, int bins, uint64_t seed) {
    return sampled_prepare_n(ctx, 1, &P, samples, bins, &seed);
}

extern "C" int bg_sampled_prepare2(bg_ctx* ctx, const bg_projector* G, const bg_projector* H, uint64_t samples, int bins,
                                   uint64_t seed_g, uint64_t seed_h) {
    const bg_projector* Ps[2] = {G, H};
    const uint64_t seeds[2] = {seed_g, seed_h};
    return sampled_prepare_n(ctx, 2, Ps, samples, bins, seeds);
}

// one bin of projector ctx->cur: prepare + pairs + finalize/sum into d_red[red_slot]
static int run_bin(bg_ctx* ctx, int bin, int red_slot) {
    const uint64_t mine = shard_count(ctx->samples, ctx->rank, ctx->world);
    const int n = (int)mine;
    const int pj = ctx->cur;
    if (n == 0) { CK(cudaMemsetAsync(RED(ctx) + red_slot, 0, sizeof(double), ctx->stream)); return 0; }
    PrepArgs pa; memset(&pa, 0, sizeof pa);
    pa.recs = ctx->d_recs; pa.n_samples = n; pa.t = ctx->t; pa.project = 1; pa.P = ctx->d_P + pj;
    pa.seed = ctx->seeds[pj]; pa.bin = (uint32_t)bin; pa.first = (uint64_t)ctx->rank; pa.stride = (uint64_t)ctx->world;
    pa.cdf = ctx->d_cdf;
    pa.zw = ctx->d_zw;
    CK(rec_event(ctx, EVP(ctx, pj, 0)));
    if (launch_prepare(ctx, SRC_RNG, pa)) return 1;
    CK(rec_event(ctx, EVP(ctx, pj, 1)));
    PairArgs qa; memset(&qa, 0, sizeof qa);
    qa.recs = ctx->d_recs; qa.n_samples = n; qa.terms = ctx->d_terms; qa.nterms = (int)ctx->terms_host.size();
    qa.t = ctx->t; qa.zw = ctx->d_zw; qa.zw2 = ctx->d_zw2;
    if (launch_pairs(ctx, qa)) return 1;
    CK(rec_event(ctx, EVP(ctx, pj, 2)));
    ctx->phase_events = true;
    k_finalize_sum_sampled<<<FIN_BLOCKS, 1024, 0, ctx->stream>>>(ctx->d_recs, ctx->d_zw, n, ctx->t, ctx->d_per, RED(ctx) + red_slot,
                                                                  ctx->d_partials, ctx->d_ticket, 0);
    CK(cudaGetLastError());
    ctx->stats.launches += 1;
    return 0;
}

static inline bool fused2(const bg_ctx* ctx) {
    return ctx->fuse2 && ctx->nproj == 2 && ctx->bins == 1 && shard_count(ctx->samples, ctx->rank, ctx->world) <= (1ull << 29);
}

// Both projectors of one probability() evaluation (bins = 1) in ONE launch per kernel: the records of
// G' and H' sit back to back, the terms are the same for both, so the pair kernel sees 2n samples — half
// the launches and one tail instead of two (what matters when 8 GPUs share the samples).
static int run_fused2(bg_ctx* ctx) {
    const uint64_t mine = shard_count(ctx->samples, ctx->rank, ctx->world);
    const int n = (int)mine;
    if (n == 0) { CK(cudaMemsetAsync(RED(ctx), 0, 8 * sizeof(double), ctx->stream)); return 0; }
    ctx->cur = 0;
    CK(cudaMemsetAsync(CNT(ctx) + 4, 0, 4 * sizeof(unsigned long long), ctx->stream));
    PrepArgs pa; memset(&pa, 0, sizeof pa);
    pa.recs = ctx->d_recs; pa.n_samples = 2 * n; pa.n_first = n; pa.t = ctx->t; pa.project = 1;
    pa.P = ctx->d_P; pa.P2 = ctx->d_P + 1; pa.seed = ctx->seeds[0]; pa.seed2 = ctx->seeds[1];
    pa.bin = 0; pa.first = (uint64_t)ctx->rank; pa.stride = (uint64_t)ctx->world;
    pa.cdf = ctx->d_cdf;
    pa.zw = ctx->d_zw;
    CK(rec_event(ctx, EVP(ctx, 0, 0)));
    if (launch_prepare(ctx, SRC_RNG, pa)) return 1;
    CK(rec_event(ctx, EVP(ctx, 0, 1)));
    PairArgs qa; memset(&qa, 0, sizeof qa);
    qa.recs = ctx->d_recs; qa.n_samples = 2 * n; qa.terms = ctx->d_terms; qa.nterms = (int)ctx->terms_host.size();
    qa.t = ctx->t; qa.zw = ctx->d_zw; qa.zw2 = ctx->d_zw2;
    if (launch_pairs(ctx, qa)) return 1;
    CK(rec_event(ctx, EVP(ctx, 0, 2)));
    for (int j = 0; j < 3; j++) CK(rec_event(ctx, EVP(ctx, 1, j)));        // second projector: inside the same launches
    ctx->phase_events = true;
    k_finalize_sum_sampled<<<dim3(FIN_BLOCKS, 2), 1024, 0, ctx->stream>>>(ctx->d_recs, ctx->d_zw, n, ctx->t, ctx->d_per, RED(ctx),
                                                                          ctx->d_partials, ctx->d_ticket, 4);
    CK(cudaGetLastError());
    ctx->stats.launches += 1;
    return 0;
}

// enqueue the whole prepared job (all projectors, all bins) on ctx's stream
static int enqueue_job(bg_ctx* ctx) {
    CK(rec_event(ctx, EV0(ctx)));
    if (fused2(ctx)) {
        if (run_fused2(ctx)) return 1;
        CK(rec_event(ctx, EV1(ctx)));
        return 0;
    }
    for (int pj = 0; pj < ctx->nproj; pj++) {
        ctx->cur = pj;
        for (int b = 0; b < ctx->bins; b++) if (run_bin(ctx, b, 4 * pj + b)) { ctx->cur = 0; return 1; }
    }
    ctx->cur = 0;
    CK(rec_event(ctx, EV1(ctx)));
    return 0;
}

// after the job's kernels: all-reduce of the slot's 8 outputs and the read-back, on the side stream
static int enqueue_collect(bg_ctx* ctx) {
    const int sl = ctx->slot;
    CK(cudaEventRecord(ctx->ev_done[sl], ctx->stream));
    CK(cudaStreamWaitEvent(ctx->cstream, ctx->ev_done[sl], 0));
    if (allreduce_red(ctx, 8, ctx->cstream)) return 1;          // one all-reduce for every projector and bin
    CK(cudaMemcpyAsync(HOUT(ctx), RED(ctx), 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->cstream));
    CK(cudaMemcpyAsync(HOUT(ctx) + 8, CNT(ctx), 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->cstream));
    CK(cudaEventRecord(ctx->ev_out[sl], ctx->cstream));
    return 0;
}

extern "C" int bg_sampled_run(bg_ctx* ctx) {
    if (!ctx) return fail(nullptr, "bg_sampled_run: null ctx");
    if (!ctx->prepared) return fail(ctx, "bg_sampled_run: bg_sampled_prepare not called");
    CK(cudaSetDevice(ctx->device));
    if (ctx->bins > 4) return fail(ctx, "bg_sampled_run: split-phase API supports at most 4 bins (use bg_sampled_norm)");
    if (ctx->run_seq - ctx->fin_seq >= 2) return fail(ctx, "bg_sampled_run: two jobs already in flight (call bg_sampled_finish)");
    ctx->slot = (int)(ctx->run_seq & 1u);
    SetGuard set_guard(ctx, ctx->overlap && ctx->slot == 1);      // overlap mode: the odd slot's own stream and buffers
    if (sync_projectors(ctx)) { ctx->slot = 0; return 1; }        // (no-op unless this set last held other projectors)
    int rc = 0;
    if (!ctx->use_graph) {
        ctx->stats.launches = 0;
        rc = enqueue_job(ctx);
    } else {
        cudaGraphExec_t& gexec = ctx->gexecs[ctx->slot];
        if (!gexec) {
            // first run of this job in this slot: capture the launch sequence once, replay it afterwards
            ctx->stats.launches = 0;
            cudaGraph_t g = nullptr;
            CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            ctx->capturing = true;
            rc = enqueue_job(ctx);
            ctx->capturing = false;
            cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
            if (rc) { if (g) cudaGraphDestroy(g); ctx->slot = 0; return 1; }
            if (e != cudaSuccess) { ctx->slot = 0; return fail(ctx, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e)); }
            e = cudaGraphInstantiate(&gexec, g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) { gexec = nullptr; ctx->slot = 0; return fail(ctx, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); }
        }
        CK(cudaGraphLaunch(gexec, ctx->stream));
    }
    if (!rc) rc = enqueue_collect(ctx);
    ctx->slot = 0;
    if (rc) return 1;
    ctx->run_seq++;
    return 0;
}

// Median of the bin means (multiSampledProjector, innerprod.c:23-41: even count -> mean of the middle two).
// The reference sorts with a comparator that truncates the double difference to int (innerprod.c:17-19): bin
// means closer than 1.0 — always, for these norms — compare equal, so its qsort leaves them in an order that is
// libc's business and "the median" is whatever sits in the middle.  The default here is a consistent ordering
// (the median the estimator is meant to take); BG_REF_MEDIAN=1 reproduces the reference's call for comparison.
static int ref_cmpfunc(const void* a, const void* b) { return (int)(*(const double*)a - *(const double*)b); }
static int median_cmpfunc(const void* a, const void* b) {
    const double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

static double median_of_bins(std::vector<double>& v) {
    const int bins = (int)v.size();
    if (bins == 1) return v[0];
    static const bool ref_order = getenv("BG_REF_MEDIAN") && atoi(getenv("BG_REF_MEDIAN")) != 0;
    qsort(v.data(), bins, sizeof(double), ref_order ? ref_cmpfunc : median_cmpfunc);
    if (bins % 2 == 1) return v[(bins - 1) / 2];
    return (v[bins / 2] + v[bins / 2 - 1]) / 2;
}

// after the stream has been synchronised: event times and the pair counters (h_out[8..16))
static int collect_stats(bg_ctx* ctx, int nproj, bool counters_in_h_out) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, EV0(ctx), EV1(ctx)));
    ctx->stats.kernel_ms = ms;
    ctx->stats.prepare_ms = ctx->stats.pairs_ms = 0;
    ctx->stats.overlapped = 0;
    if (ctx->phase_events) {
        for (int pj = 0; pj < nproj; pj++) {
            float a = 0, b = 0;
            if (cudaEventElapsedTime(&a, EVP(ctx, pj, 0), EVP(ctx, pj, 1)) == cudaSuccess) ctx->stats.prepare_ms += a;
            if (cudaEventElapsedTime(&b, EVP(ctx, pj, 1), EVP(ctx, pj, 2)) == cudaSuccess) ctx->stats.pairs_ms += b;
        }
        ctx->phase_events = false;
    }
    unsigned long long c[8];
    if (counters_in_h_out) memcpy(c, HOUT(ctx) + 8, sizeof c);
    else CK(cudaMemcpy(c, CNT(ctx), sizeof c, cudaMemcpyDeviceToHost));
    ctx->stats.pairs = 0;
    for (int pj = 0; pj < nproj; pj++) ctx->stats.pairs += c[4 * pj + 1];
    return 0;
}

static int sampled_finish_n(bg_ctx* ctx, double* out) {
    if (!ctx) return fail(nullptr, "bg_sampled_finish: null ctx");
    if (!ctx->prepared || ctx->run_seq == ctx->fin_seq) return fail(ctx, "bg_sampled_finish: nothing was run");
    if (!out) return fail(ctx, "bg_sampled_finish: null out");
    CK(cudaSetDevice(ctx->device));
    ctx->slot = (int)(ctx->fin_seq & 1u);                     // the oldest job in flight
    ctx->fin_seq++;
    cudaError_t e = cudaEventSynchronize(ctx->ev_out[ctx->slot]);
    if (e != cudaSuccess) { ctx->slot = 0; return fail(ctx, "cudaEventSynchronize failed: %s", cudaGetErrorString(e)); }
    ctx->stats.d2h_bytes = (uint64_t)ctx->nproj * ctx->bins * sizeof(double);
    if (ctx->use_graph)           // prepare, 2 x pairs, finalize: per projector and bin, or once for a fused two-projector job
        ctx->stats.launches = fused2(ctx) ? 4 : (uint64_t)ctx->nproj * ctx->bins * 4;
    ctx->stats.pair_launches = fused2(ctx) ? 1 : (uint64_t)ctx->nproj * ctx->bins;
    ctx->phase_events = true;
    ctx->per_valid = true;
    ctx->per_set = (ctx->overlap && ctx->slot == 1) ? 1 : 0;
    const int rc = collect_stats(ctx, ctx->nproj, true);
    ctx->stats.overlapped = ctx->overlap ? 1 : 0;
    for (int pj = 0; pj < ctx->nproj && !rc; pj++) {
        std::vector<double> v(ctx->bins);
        for (int b = 0; b < ctx->bins; b++) v[b] = HOUT(ctx)[4 * pj + b] / (double)ctx->samples;   // total/samples (innerprod.c:83)
        out[pj] = median_of_bins(v);
    }
    ctx->slot = 0;
    return rc;
}

extern "C" int bg_sampled_finish(bg_ctx* ctx, double norm, double* out) {
    (void)norm;
    if (ctx && ctx->nproj != 1) return fail(ctx, "bg_sampled_finish: the prepared job has 2 projectors (use bg_sampled_finish2)");
    return sampled_finish_n(ctx, out);
}

extern "C" int bg_sampled_finish2(bg_ctx* ctx, double norm, double out[2]) {
    (void)norm;
    if (ctx && ctx->nproj != 2) return fail(ctx, "bg_sampled_finish2: the prepared job has 1 projector");
    return sampled_finish_n(ctx, out);
}

extern "C" int bg_sampled_norm(bg_ctx* ctx, const bg_projector* P, uint64_t samples, int bins, uint64_t seed,
                               double norm, double* out) {
    if (!ctx) return fail(nullptr, "bg_sampled_norm: null ctx");
    if (!out) return fail(ctx, "bg_sampled_norm: null out");
    if (P && P->nstabs > 0 && P->nqubits == 0) { *out = pow(norm, 2) * clifford_closed_form(P); return 0; }
    if (check_projector(ctx, P)) return 1;
    if (P->nstabs == 0) { *out = pow(norm, 2); return 0; }                 // innerprod.c:47
    if (bins < 1) return fail(ctx, "bg_sampled_norm: bins = %d", bins);
    if (bg_sampled_prepare(ctx, P, samples, 1, seed)) return 1;
    std::vector<double> v(bins);
    double total_ms = 0, prep_ms = 0, pair_ms = 0; uint64_t total_pairs = 0, launches = 0;
    ctx->cur = 0;
    for (int b = 0; b < bins; b++) {
        ctx->stats.launches = 0;
        CK(cudaEventRecord(EV0(ctx), ctx->stream));
        if (run_bin(ctx, b, 0)) return 1;
        CK(cudaEventRecord(EV1(ctx), ctx->stream));
        if (allreduce_red(ctx, 1)) return 1;
        CK(cudaMemcpyAsync(HOUT(ctx), RED(ctx), sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(HOUT(ctx) + 8, CNT(ctx), 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (collect_stats(ctx, 1, true)) return 1;
        total_ms += ctx->stats.kernel_ms; total_pairs += ctx->stats.pairs; launches += ctx->stats.launches;
        prep_ms += ctx->stats.prepare_ms; pair_ms += ctx->stats.pairs_ms;
        v[b] = HOUT(ctx)[0] / (double)samples;
    }
    ctx->stats.kernel_ms = total_ms; ctx->stats.pairs = total_pairs; ctx->stats.launches = launches;
    ctx->stats.prepare_ms = prep_ms; ctx->stats.pairs_ms = pair_ms;
    ctx->stats.d2h_bytes = bins * sizeof(double);
    ctx->per_valid = true; ctx->per_set = 0;
    *out = median_of_bins(v);
    return 0;
}

// Numerator and denominator of one probability() evaluation (probability.c:197-198): both
// projectors against the same decomposition, ONE all-reduce and one host synchronisation.
extern "C" int bg_sampled_norm2(bg_ctx* ctx, const bg_projector* G, const bg_projector* H, uint64_t samples, int bins,
                                uint64_t seed_g, uint64_t seed_h, double norm, double out[2]) {
    if (!ctx) return fail(nullptr, "bg_sampled_norm2: null ctx");
    if (!out || !G || !H) return fail(ctx, "bg_sampled_norm2: null argument");
    const bool plain = G->nstabs > 0 && H->nstabs > 0 && G->nqubits > 0 && H->nqubits > 0 && bins >= 1 && bins <= 4;
    if (!plain) {            // closed forms / many bins: one projector at a time
        if (bg_sampled_norm(ctx, G, samples, bins, seed_g, norm, &out[0])) return 1;
        return bg_sampled_norm(ctx, H, samples, bins, seed_h, norm, &out[1]);
    }
    if (bg_sampled_prepare2(ctx, G, H, samples, bins, seed_g, seed_h)) return 1;      // no-op when this job is already prepared
    if (bg_sampled_run(ctx)) return 1;                                               // replays the captured graph
    return sampled_finish_n(ctx, out);
}

extern "C" int bg_sampled_per_sample(bg_ctx* ctx, int projector, uint64_t first, size_t count, double* out) {
    if (!ctx) return fail(nullptr, "bg_sampled_per_sample: null ctx");
    if (!ctx->prepared || !ctx->per_valid || ctx->run_seq != ctx->fin_seq)
        return fail(ctx, "bg_sampled_per_sample: no finished sampled job");
    if (projector < 0 || projector >= ctx->nproj) return fail(ctx, "bg_sampled_per_sample: projector %d of %d", projector, ctx->nproj);
    if (count == 0) return 0;
    if (!out) return fail(ctx, "bg_sampled_per_sample: null out");
    const uint64_t mine = shard_count(ctx->samples, ctx->rank, ctx->world);
    if (first + count > mine) return fail(ctx, "bg_sampled_per_sample: samples %llu..%llu of %llu on this rank",
                                          (unsigned long long)first, (unsigned long long)(first + count), (unsigned long long)mine);
    CK(cudaSetDevice(ctx->device));
    // a fused two-projector job keeps both segments; otherwise d_per holds the projector that ran last
    if (!fused2(ctx) && projector != ctx->nproj - 1)
        return fail(ctx, "bg_sampled_per_sample: only the last projector's values are kept for this job (bins > 1 or BG_FUSE2=0)");
    const size_t off = fused2(ctx) ? (size_t)projector * (size_t)mine : 0;
    SetGuard set_guard(ctx, ctx->per_set == 1);
    CK(cudaMemcpy(out, ctx->d_per + off + first, count * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

// ---- exact norm ---------------------------------------------------------------------------
static int exact_norm_parts(bg_ctx* ctx, const bg_projector* P, double s[2]);

extern "C" int bg_exact_norm(bg_ctx* ctx, const bg_projector* P, double norm, double* out) {
    if (!ctx) return fail(nullptr, "bg_exact_norm: null ctx");
    if (!out) return fail(ctx, "bg_exact_norm: null out");
    if (P && P->nstabs > 0 && P->nqubits == 0) { *out = clifford_closed_form(P); return 0; }
    if (P && P->nstabs == 0) { *out = pow(norm, 2); return 0; }
    if (ctx->world > 1 && !ctx->allreduce) return fail(ctx, "bg_exact_norm: world > 1 needs the in-library all-reduce (partial |sum| values do not add; bg_exact_norm_parts returns the parts)");
    double s[2];
    if (exact_norm_parts(ctx, P, s)) return 1;
    *out = sqrt(s[0] * s[0] + s[1] * s[1]);                                 // ComplexMag (innerprod.c:198)
    return 0;
}

extern "C" int bg_exact_norm_parts(bg_ctx* ctx, const bg_projector* P, double norm, double out[2]) {
    if (!ctx) return fail(nullptr, "bg_exact_norm_parts: null ctx");
    if (!out) return fail(ctx, "bg_exact_norm_parts: null out");
    out[0] = out[1] = 0;
    const bool mine = ctx->rank == 0 || ctx->allreduce;                     // closed forms: once, not once per rank
    if (P && P->nstabs > 0 && P->nqubits == 0) { if (mine) out[0] = clifford_closed_form(P); return 0; }
    if (P && P->nstabs == 0) { if (mine) out[0] = pow(norm, 2); return 0; }
    return exact_norm_parts(ctx, P, out);
}

static int exact_norm_parts(bg_ctx* ctx, const bg_projector* P, double s[2]) {
    if (ctx->t <= 0) return fail(ctx, "bg_exact_norm: call bg_set_decomposition first");
    if (check_projector(ctx, P)) return 1;
    CK(cudaSetDevice(ctx->device));
    const uint64_t chi = ctx->terms_host.size();
    const uint64_t mine = shard_count(chi, ctx->rank, ctx->world);
    const int n = (int)mine;
    if (ensure_sample_buffers(ctx, (size_t)std::max<uint64_t>(mine, 1))) return 1;
    ctx->prepared = false; ctx->P_valid = false; // d_P no longer holds the prepared job's projectors
    CK(cudaMemcpyAsync(ctx->d_P, P, sizeof(bg_projector), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes = sizeof(bg_projector);
    ctx->stats.launches = 0;
    CK(cudaEventRecord(EV0(ctx), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_red, 0, 2 * sizeof(double), ctx->stream));
    if (n > 0) {
        PrepArgs pa; memset(&pa, 0, sizeof pa);
        pa.recs = ctx->d_recs; pa.n_samples = n; pa.t = ctx->t; pa.project = 1; pa.P = ctx->d_P;
        pa.first = (uint64_t)ctx->rank; pa.stride = (uint64_t)ctx->world;
        pa.terms = ctx->d_terms; pa.exact = ctx->exact;
        pa.zw = ctx->d_zw; pa.zw2 = ctx->d_zw2;
        if (launch_prepare(ctx, SRC_TERMS, pa)) return 1;
        PairArgs qa; memset(&qa, 0, sizeof qa);
        qa.recs = ctx->d_recs; qa.n_samples = n; qa.terms = ctx->d_terms; qa.nterms = (int)chi;
        qa.t = ctx->t; qa.zw = ctx->d_zw; qa.zw2 = ctx->d_zw2; qa.tri = 1;
        qa.first = (uint64_t)ctx->rank; qa.stride = (uint64_t)ctx->world;
        if (launch_pairs(ctx, qa)) return 1;
        k_finalize_exact<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_recs, ctx->d_zw, ctx->d_zw2, n, ctx->t, ctx->d_per, ctx->d_per2);
        k_sum<<<1, 1024, 0, ctx->stream>>>(ctx->d_per, n, ctx->d_red);
        k_sum<<<1, 1024, 0, ctx->stream>>>(ctx->d_per2, n, ctx->d_red + 1);
        CK(cudaGetLastError());
        ctx->stats.launches += 3;
    }
    CK(cudaEventRecord(EV1(ctx), ctx->stream));
    if (allreduce_red(ctx, 2)) return 1;
    s[0] = s[1] = 0;
    CK(cudaMemcpyAsync(s, ctx->d_red, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes = 2 * sizeof(double);
    if (collect_stats(ctx, 1, false)) return 1;
    return 0;
}

// ---- parity / debug entry points -----------------------------------------------------------
static int check_states(bg_ctx* ctx, const bg_state* s, size_t n, int* t_out) {
    int t = -1;
    for (size_t i = 0; i < n; i++) {
        if (s[i].n < 1 || s[i].n > BG_MAX_T || s[i].k < 0 || s[i].k > s[i].n)
            return fail(ctx, "state %zu has n = %d, k = %d", i, s[i].n, s[i].k);
        if (t < 0) t = s[i].n;
        if ((s[i].n <= 32) != (t <= 32)) t = 64;      // mixed widths: use the wide kernel
    }
    *t_out = t;
    return 0;
}

extern "C" int bg_inner_products(bg_ctx* ctx, size_t n_pairs, const bg_state* a, const bg_state* b, int32_t* epm) {
    if (!ctx) return fail(nullptr, "bg_inner_products: null ctx");
    if (n_pairs == 0) return 0;
    if (!a || !b || !epm) return fail(ctx, "bg_inner_products: null buffer");
    CK(cudaSetDevice(ctx->device));
    int ta, tb;
    if (check_states(ctx, a, n_pairs, &ta) || check_states(ctx, b, n_pairs, &tb)) return 1;
    for (size_t i = 0; i < n_pairs; i++)
        if (a[i].n != b[i].n) return fail(ctx, "pair %zu: states of different size (%d vs %d)", i, a[i].n, b[i].n);
    DevBuf<bg_state> da, db; DevBuf<int32_t> de;
    CK(da.alloc(n_pairs)); CK(db.alloc(n_pairs)); CK(de.alloc(n_pairs * 3));
    CK(cudaMemcpyAsync(da, a, n_pairs * sizeof(bg_state), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(db, b, n_pairs * sizeof(bg_state), cudaMemcpyHostToDevice, ctx->stream));
    const int blocks = (int)std::min<size_t>((n_pairs + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, (size_t)ctx->sm_count * 16);
    CK(cudaEventRecord(EV0(ctx), ctx->stream));
    if (std::max(ta, tb) <= 32) k_inner_products<1><<<blocks, 32 * WARPS_PER_BLOCK, 0, ctx->stream>>>(da, db, n_pairs, de);
    else k_inner_products<2><<<blocks, 32 * WARPS_PER_BLOCK, 0, ctx->stream>>>(da, db, n_pairs, de);
    CK(cudaGetLastError());
    CK(cudaEventRecord(EV1(ctx), ctx->stream));
    CK(cudaMemcpyAsync(epm, de, n_pairs * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0; cudaEventElapsedTime(&ms, EV0(ctx), EV1(ctx));
    ctx->stats.kernel_ms = ms; ctx->stats.pairs = n_pairs; ctx->stats.launches = 1;
    ctx->stats.h2d_bytes = 2 * n_pairs * sizeof(bg_state); ctx->stats.d2h_bytes = n_pairs * 3 * sizeof(int32_t);
    return 0;
}

extern "C" int bg_sampled_norm_from_states(bg_ctx* ctx, const bg_projector* P, int project, size_t n_states,
                                           const bg_state* thetas, int32_t* epm, double* per_sample, double* mean) {
    if (!ctx) return fail(nullptr, "bg_sampled_norm_from_states: null ctx");
    if (ctx->t <= 0) return fail(ctx, "bg_sampled_norm_from_states: call bg_set_decomposition first");
    if (n_states == 0) { if (mean) *mean = 0; return 0; }
    if (!thetas) return fail(ctx, "bg_sampled_norm_from_states: null thetas");
    if (project && check_projector(ctx, P)) return 1;
    int tt;
    if (check_states(ctx, thetas, n_states, &tt)) return 1;
    for (size_t i = 0; i < n_states; i++)
        if (thetas[i].n != ctx->t) return fail(ctx, "theta %zu has n = %d but t = %d", i, thetas[i].n, ctx->t);
    CK(cudaSetDevice(ctx->device));
    const int n = (int)n_states;
    const size_t chi = ctx->terms_host.size();
    if (ensure_sample_buffers(ctx, n_states)) return 1;
    DevBuf<bg_state> dth; DevBuf<int32_t> depm;
    CK(dth.alloc(n_states));
    CK(cudaMemcpyAsync(dth, thetas, n_states * sizeof(bg_state), cudaMemcpyHostToDevice, ctx->stream));
    ctx->prepared = false; ctx->P_valid = false; // d_P / the sample buffers no longer belong to the prepared job
    if (project) CK(cudaMemcpyAsync(ctx->d_P, P, sizeof(bg_projector), cudaMemcpyHostToDevice, ctx->stream));
    if (epm) {
        CK(depm.alloc(n_states * chi * 3));
        CK(cudaMemsetAsync(depm, 0, n_states * chi * 3 * sizeof(int32_t), ctx->stream));
    }
    ctx->stats.launches = 0;
    CK(cudaEventRecord(EV0(ctx), ctx->stream));
    PrepArgs pa; memset(&pa, 0, sizeof pa);
    pa.recs = ctx->d_recs; pa.n_samples = n; pa.t = ctx->t; pa.project = project ? 1 : 0; pa.P = ctx->d_P;
    pa.states = dth;
    pa.zw = ctx->d_zw;
    if (launch_prepare(ctx, SRC_STATES, pa)) return 1;
    PairArgs qa; memset(&qa, 0, sizeof qa);
    qa.recs = ctx->d_recs; qa.n_samples = n; qa.terms = ctx->d_terms; qa.nterms = (int)chi;
    qa.t = ctx->t; qa.zw = ctx->d_zw; qa.zw2 = ctx->d_zw2; qa.epm = depm;
    if (launch_pairs(ctx, qa)) return 1;
    k_finalize_sampled<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_recs, ctx->d_zw, n, ctx->t, ctx->d_per);
    k_sum<<<1, 1024, 0, ctx->stream>>>(ctx->d_per, n, ctx->d_red);
    CK(cudaGetLastError());
    ctx->stats.launches += 2;
    CK(cudaEventRecord(EV1(ctx), ctx->stream));
    double s = 0;
    CK(cudaMemcpyAsync(&s, ctx->d_red, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (per_sample) CK(cudaMemcpyAsync(per_sample, ctx->d_per, n_states * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (epm) CK(cudaMemcpyAsync(epm, depm, n_states * chi * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (collect_stats(ctx, 1, false)) return 1;
    ctx->stats.h2d_bytes = n_states * sizeof(bg_state);
    ctx->stats.d2h_bytes = sizeof(double) + (per_sample ? n_states * sizeof(double) : 0) + (epm ? n_states * chi * 12 : 0);
    if (mean) *mean = s / (double)n_states;
    return 0;
}

// Active-mask layout -> contiguous layout (active rows first, order preserved).
static void compact_state(bg_state* s, uint64_t A) {
    const int n = s->n;
    int perm[BG_MAX_T], k = 0, r = 0;
    for (int v = 0; v < n; v++) if ((A >> v) & 1) perm[r++] = v;
    k = r;
    for (int v = 0; v < n; v++) if (!((A >> v) & 1)) perm[r++] = v;
    bg_state o; memset(&o, 0, sizeof o);
    o.n = n; o.k = k; o.Q = s->Q & 7; o.h = s->h;
    for (int i = 0; i < n; i++) {
        o.G[i] = s->G[perm[i]]; o.Gbar[i] = s->Gbar[perm[i]];
        if (i < k) {
            o.D1 |= ((s->D1 >> perm[i]) & 1ull) << i;
            o.D2 |= ((s->D2 >> perm[i]) & 1ull) << i;
            uint64_t row = 0;
            for (int c = 0; c < k; c++) row |= ((s->J[perm[i]] >> perm[c]) & 1ull) << c;
            o.J[i] = row;
        }
    }
    *s = o;
}

extern "C" int bg_measure_pauli(bg_ctx* ctx, size_t n_states, bg_state* states, const int32_t* m,
                                const uint64_t* zeta, const uint64_t* xi, double* result) {
    if (!ctx) return fail(nullptr, "bg_measure_pauli: null ctx");
    if (n_states == 0) return 0;
    if (!states || !m || !zeta || !xi) return fail(ctx, "bg_measure_pauli: null buffer");
    int tt;
    if (check_states(ctx, states, n_states, &tt)) return 1;
    CK(cudaSetDevice(ctx->device));
    DevBuf<bg_state> ds; DevBuf<uint64_t> dA, dz, dx; DevBuf<int32_t> dm, dc;
    CK(ds.alloc(n_states)); CK(dA.alloc(n_states)); CK(dz.alloc(n_states)); CK(dx.alloc(n_states));
    CK(dm.alloc(n_states)); CK(dc.alloc(n_states));
    CK(cudaMemcpyAsync(ds, states, n_states * sizeof(bg_state), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dz, zeta, n_states * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dx, xi, n_states * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dm, m, n_states * 4, cudaMemcpyHostToDevice, ctx->stream));
    const int blocks = (int)std::min<size_t>((n_states + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, (size_t)ctx->sm_count * 16);
    if (tt <= 32) k_measure_pauli<1><<<blocks, 32 * WARPS_PER_BLOCK, 0, ctx->stream>>>(ds, dA, n_states, dm, dz, dx, dc);
    else k_measure_pauli<2><<<blocks, 32 * WARPS_PER_BLOCK, 0, ctx->stream>>>(ds, dA, n_states, dm, dz, dx, dc);
    CK(cudaGetLastError());
    std::vector<uint64_t> A(n_states); std::vector<int32_t> code(n_states);
    CK(cudaMemcpyAsync(states, ds, n_states * sizeof(bg_state), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(A.data(), dA, n_states * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(code.data(), dc, n_states * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < n_states; i++) {
        compact_state(&states[i], A[i]);
        if (result) result[i] = code[i] == 0 ? 0.0 : (code[i] == 1 ? 1.0 : pow(2, -0.5));
    }
    return 0;
}

// src: SRC_RNG (t, seed, bin, first) or SRC_TERMS (ctx's decomposition, first)
static int dump_states(bg_ctx* ctx, int src, int t, uint64_t seed, int bin, uint64_t first, size_t count, bg_state* out) {
    if (count == 0) return 0;
    if (!out) return fail(ctx, "null output buffer");
    CK(cudaSetDevice(ctx->device));
    if (ensure_sample_buffers(ctx, count)) return 1;
    DevBuf<bg_state> ds; DevBuf<uint64_t> dA;
    CK(ds.alloc(count)); CK(dA.alloc(count));
    if (ctx->cdf_t != t && src == SRC_RNG) {
        double cdf[BG_MAX_T + 1];
        dimension_cdf(t, cdf);
        CK(cudaMemcpyAsync(ctx->d_cdf, cdf, (t + 1) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->cdf_t = t;
    }
    PrepArgs pa; memset(&pa, 0, sizeof pa);
    pa.recs = ctx->d_recs; pa.n_samples = (int)count; pa.t = t; pa.project = 0; pa.P = ctx->d_P;
    pa.seed = seed; pa.bin = (uint32_t)bin; pa.first = first; pa.stride = 1; pa.cdf = ctx->d_cdf;
    pa.terms = ctx->d_terms; pa.exact = ctx->exact;
    pa.raw_out = ds; pa.raw_A = dA;
    if (launch_prepare(ctx, src, pa)) return 1;
    std::vector<uint64_t> A(count);
    CK(cudaMemcpyAsync(out, ds, count * sizeof(bg_state), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(A.data(), dA, count * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < count; i++) compact_state(&out[i], A[i]);
    return 0;
}

extern "C" int bg_random_states(bg_ctx* ctx, int t, uint64_t seed, int bin, uint64_t first, size_t count, bg_state* out) {
    if (!ctx) return fail(nullptr, "bg_random_states: null ctx");
    if (t < 1 || t > BG_MAX_T) return fail(ctx, "bg_random_states: t = %d outside 1..%d", t, BG_MAX_T);
    return dump_states(ctx, SRC_RNG, t, seed, bin, first, count, out);
}

extern "C" int bg_decomposition_terms(bg_ctx* ctx, uint64_t first, size_t count, bg_state* out) {
    if (!ctx) return fail(nullptr, "bg_decomposition_terms: null ctx");
    if (ctx->t <= 0) return fail(ctx, "bg_decomposition_terms: call bg_set_decomposition first");
    if (first + count > ctx->terms_host.size()) return fail(ctx, "bg_decomposition_terms: range beyond chi = %zu", ctx->terms_host.size());
    return dump_states(ctx, SRC_TERMS, ctx->t, 0, 0, first, count, out);
}

// ---- integer-pipe peak (the roofline denominator) ---------------------------------------------
// Every thread runs `iters` x 4 rounds over 8 independent chains of LOP3 (kind 0) or POPC
// (kind 1), one instruction per chain per round (asm volatile keeps the count exact); sink prevents dead-code elimination.  lane-ops = threads * iters * 8 * ops_per_round.
template <int KIND>
__global__ void __launch_bounds__(256) k_int_peak(uint32_t* sink, int iters, uint32_t seed) {
    uint32_t x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = seed * (threadIdx.x + 1u) + 0x9E3779B9u * (j + 1) + blockIdx.x;
    const uint32_t c = seed;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (KIND == 0)      // exactly one 3-input LOP3:  x_j ^= x_{j+1} & c
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x[j]) : "r"(x[(j + 1) & 7]), "r"(c));
                else                // exactly one POPC (the result feeds itself)
                    asm volatile("popc.b32 %0, %0;" : "+r"(x[j]));
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) acc ^= x[j];
    if (acc == 0x12345u) sink[0] = acc;
}

extern "C" int bg_measure_int_peak(bg_ctx* ctx, double* lop3_lane_ops_per_s, double* popc_lane_ops_per_s) {
    if (!ctx) return fail(nullptr, "bg_measure_int_peak: null ctx");
    CK(cudaSetDevice(ctx->device));
    uint32_t* sink = nullptr;
    CK(cudaMalloc((void**)&sink, 64));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
    double res[2] = {0, 0};
    for (int kind = 0; kind < 2; kind++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(EV0(ctx), ctx->stream));
            if (kind == 0) k_int_peak<0><<<blocks, threads, 0, ctx->stream>>>(sink, iters, 12345u + rep);
            else k_int_peak<1><<<blocks, threads, 0, ctx->stream>>>(sink, iters, 12345u + rep);
            CK(cudaGetLastError());
            CK(cudaEventRecord(EV1(ctx), ctx->stream));
            CK(cudaEventSynchronize(EV1(ctx)));
            float ms = 0; CK(cudaEventElapsedTime(&ms, EV0(ctx), EV1(ctx)));
            if (rep > 0 && ms < best) best = ms;
        }
        const double ops = (double)blocks * threads * (double)iters * 32.0;     // 4 x 8 ops per iteration
        res[kind] = ops / (best * 1e-3);
    }
    cudaFree(sink);
    if (lop3_lane_ops_per_s) *lop3_lane_ops_per_s = res[0];
    if (popc_lane_ops_per_s) *popc_lane_ops_per_s = res[1];
    return 0;
}

// ------------------------------------------------------------------------------------------
// decompose()'s fidelity loop (libcirc/probability.c:373-391): the Hamming weights of all 2^k
// combinations of the rows of L.  Thread g owns the 2^low combinations whose high bits are g and walks
// them in Gray-code order (one XOR + one POPC per combination); counts go to lane-private columns of a
// per-warp histogram in shared memory (bank = lane, no atomics), flushed with one atomic per bin.
// ------------------------------------------------------------------------------------------
#define WH_WARPS 4
__global__ void __launch_bounds__(32 * WH_WARPS) k_weight_hist(const uint64_t* __restrict__ L, int k, int low,
                                                               unsigned long long* hist) {
    __shared__ uint32_t s_h[WH_WARPS][65][32];
    __shared__ uint64_t s_L[64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = threadIdx.x; j < 64; j += blockDim.x) s_L[j] = j < k ? L[j] : 0ull;
    for (int j = lane; j < 65 * 32; j += 32) (&s_h[warp][0][0])[j] = 0u;
    __syncthreads();
    const unsigned long long groups = 1ull << (k - low);
    const unsigned long long per = 1ull << low;
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups;
         g += (unsigned long long)gridDim.x * blockDim.x) {
        uint64_t x = 0;
        for (int j = low; j < k; j++) if ((g >> (j - low)) & 1ull) x ^= s_L[j];
        s_h[warp][__popcll(x)][lane]++;
        for (unsigned long long c = 1; c < per; c++) {
            x ^= s_L[__ffsll((long long)c) - 1];          // Gray code: step c flips bit ctz(c)
            s_h[warp][__popcll(x)][lane]++;
        }
    }
    __syncwarp();
    for (int w = lane; w < 65; w += 32) {
        unsigned long long sum = 0;
        for (int l = 0; l < 32; l++) sum += s_h[warp][w][l];
        if (sum) atomicAdd(&hist[w], sum);
    }
}

extern "C" int bg_decomposition_weights(bg_ctx* ctx, int t, int k, const uint64_t* L_rows, uint64_t hist[65]) {
    if (!ctx) return fail(nullptr, "bg_decomposition_weights: null ctx");
    if (!hist) return fail(ctx, "bg_decomposition_weights: null hist");
    if (t < 1 || t > BG_MAX_T) return fail(ctx, "bg_decomposition_weights: t = %d outside 1..%d", t, BG_MAX_T);
    if (k < 0 || k > t || k > 44) return fail(ctx, "bg_decomposition_weights: k = %d outside 0..min(t,44)", k);
    if (k > 0 && !L_rows) return fail(ctx, "bg_decomposition_weights: L_rows is null");
    CK(cudaSetDevice(ctx->device));
    const uint64_t maskt = t >= 64 ? ~0ull : ((1ull << t) - 1);
    uint64_t rows[64];
    for (int j = 0; j < 64; j++) rows[j] = j < k ? (L_rows[j] & maskt) : 0ull;
    unsigned long long* d = nullptr;                 // [0..64] histogram, [65..128] the rows
    CK(cudaMalloc((void**)&d, (65 + 64) * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d, 0, 65 * sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + 65, rows, sizeof rows, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        // every thread walks 2^low combinations per group, low <= 16, with 2^18 groups (more than the
        // resident threads) as soon as k allows; a lane-private counter stays far below 2^32 for k <= 44
        const int low = std::min(16, std::max(0, k - 18));
        const unsigned long long groups = 1ull << (k - low);
        const int threads = 32 * WH_WARPS;
        const int blocks = (int)std::min<unsigned long long>((groups + threads - 1) / threads, (unsigned long long)ctx->sm_count * 8);
        k_weight_hist<<<blocks, threads, 0, ctx->stream>>>((const uint64_t*)(d + 65), k, low, d);
        e = cudaGetLastError();
    }
    unsigned long long h[65];
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, "bg_decomposition_weights: %s", cudaGetErrorString(e));
    for (int w = 0; w < 65; w++) hist[w] = h[w];
    ctx->stats.launches = 1; ctx->stats.h2d_bytes = sizeof rows; ctx->stats.d2h_bytes = sizeof h;
    return 0;
}

extern "C" int bg_get_stats(const bg_ctx* ctx, bg_stats* out) {
    if (!ctx || !out) return fail(nullptr, "bg_get_stats: null argument");
    *out = ctx->stats;
    return 0;
}
