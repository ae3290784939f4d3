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
// same algebra per thread needs ~2000 thread-instructions per pair (measured: 2080, 79 warp-instructions
// at 26.3 active lanes, profiles/r1s3_k_pairs_tpp_ncu_summary.json).
//
// Structure of one inner product (t_term_L / t_term_H): working copy of the ambient rows (t_copy_in);
// the parity checks of K_theta pivoted one by one (t_constraints -> t_pivot -> t_xor2 row pass); then the
// exponential sum (t_expsum): the fold of the odd-D variables is computed but its row update stays
// pending, and the monomer / dimer rounds (t_rounds -> t_block64 / t_block32) eliminate two steps per
// pass over the rows (t_xor4; the first pass also carries the fold: t_xor6), without a branch.
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
extern int* g_bg_trace; extern int g_bg_trace_n, g_bg_trace_cap;     // per row pass: 1000 x masks (2, 4, 6) + rows touched in the low half, rows in the high half
#define BG_TRACE(lo, hi) do { if (g_bg_trace && g_bg_trace_n + 2 <= g_bg_trace_cap) { g_bg_trace[g_bg_trace_n++] = (lo); g_bg_trace[g_bg_trace_n++] = (hi); } } while (0)
#else
#define BG_WORK(field, n) ((void)0)
#define BG_TRACE(lo, hi) ((void)0)
#endif

namespace bg {

// per-thread view of its working rows: row r lives at base[r * stride]
#if defined(__CUDA_ARCH__)
// one IMAD for the address, LDS/STS on the 32-bit shared-memory window (no generic-pointer arithmetic)
BG_HD void t_lds(uint32_t addr, uint32_t& v) { asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); }
BG_HD void t_lds(uint32_t addr, uint64_t& v) {
    uint32_t lo, hi;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr));
    v = ((uint64_t)hi << 32) | lo;
}
BG_HD void t_sts(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
BG_HD void t_sts(uint32_t addr, uint64_t v) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(addr), "r"((uint32_t)v), "r"((uint32_t)(v >> 32)) : "memory");
}
#endif
template <typename W> struct Rows {
    W* base;
    int stride;
    uint32_t sbase, sstride;     // device: shared-memory byte address of row 0, byte stride between rows
#if defined(__CUDA_ARCH__)
    BG_HDM W get(int r) const { W v; t_lds((uint32_t)r * sstride + sbase, v); return v; }
    BG_HDM void put(int r, W v) const { t_sts((uint32_t)r * sstride + sbase, v); }
#else
    BG_HDM W get(int r) const { return base[(size_t)r * stride]; }
    BG_HDM void put(int r, W v) const { base[(size_t)r * stride] = v; }
#endif
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
    return thighest(hi ? hi : (uint32_t)x) + (hi ? 32 : 0);          // selects, no branch: one FLO
}

// row_c ^= [c in M1] V1 ^ [c in M2] V2  for every c in M1 | M2 — the one primitive every step of the
// elimination reduces to.  A single loop (not one per mask) keeps the trip counts of the 32 lanes
// of a warp close together; words are walked in 32-bit halves from the top bit down (FLO + 2 ops).
BG_HD void t_xor2(const Rows<uint32_t>& J, uint32_t M1, uint32_t V1, uint32_t M2, uint32_t V2) {
    uint32_t U = M1 | M2;
    BG_TRACE(2000 + tpopc(U), 0);
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
    BG_TRACE(2000 + tpopc((uint32_t)(M1 | M2)), tpopc((uint32_t)((M1 | M2) >> 32)));
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

// Two elimination steps at once:  row_c ^= [c in M1] V1 ^ [c in M2] V2 ^ [c in M3] V3 ^ [c in M4] V4.
// A row is touched when ANY of its four bits is set (15/16 of the remaining rows instead of 3/4 twice), so
// the trip counts of the 32 lanes of a warp are nearly equal and each row is loaded and stored once.
BG_HD void t_xor4(const Rows<uint32_t>& J, uint32_t M1, uint32_t V1, uint32_t M2, uint32_t V2,
                  uint32_t M3, uint32_t V3, uint32_t M4, uint32_t V4) {
    uint32_t U = M1 | M2 | M3 | M4;
    BG_TRACE(4000 + tpopc(U), 0);
#if defined(__CUDA_ARCH__)
    while (U) {
        const uint32_t c = (uint32_t)thighest(U);
        const uint32_t b = 1u << c;
        U ^= b;
        const uint32_t addr = c * J.sstride + J.sbase;
        uint32_t r;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
        asm("{\n\t.reg .pred p, q, r, s;\n\t"
            "setp.ne.u32 p, %1, 0;\n\tsetp.ne.u32 q, %2, 0;\n\tsetp.ne.u32 r, %3, 0;\n\tsetp.ne.u32 s, %4, 0;\n\t"
            "@p xor.b32 %0, %0, %5;\n\t@q xor.b32 %0, %0, %6;\n\t@r xor.b32 %0, %0, %7;\n\t@s xor.b32 %0, %0, %8;\n\t}"
            : "+r"(r) : "r"(M1 & b), "r"(M2 & b), "r"(M3 & b), "r"(M4 & b), "r"(V1), "r"(V2), "r"(V3), "r"(V4));
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
        if (M3 & b) r ^= V3;
        if (M4 & b) r ^= V4;
        J.put(c, r);
        BG_WORK(rows, 1); BG_WORK(xors, ((M1 & b) ? 1 : 0) + ((M2 & b) ? 1 : 0) + ((M3 & b) ? 1 : 0) + ((M4 & b) ? 1 : 0));
    }
#endif
}
BG_HD void t_xor4(const Rows<uint64_t>& J, uint64_t M1, uint64_t V1, uint64_t M2, uint64_t V2,
                  uint64_t M3, uint64_t V3, uint64_t M4, uint64_t V4) {
    BG_TRACE(4000 + tpopc((uint32_t)(M1 | M2 | M3 | M4)), tpopc((uint32_t)((M1 | M2 | M3 | M4) >> 32)));
#if defined(__CUDA_ARCH__)
    const uint32_t v1l = (uint32_t)V1, v1h = (uint32_t)(V1 >> 32), v2l = (uint32_t)V2, v2h = (uint32_t)(V2 >> 32);
    const uint32_t v3l = (uint32_t)V3, v3h = (uint32_t)(V3 >> 32), v4l = (uint32_t)V4, v4h = (uint32_t)(V4 >> 32);
#pragma unroll
    for (int h = 1; h >= 0; h--) {
        const uint32_t m1 = (uint32_t)(M1 >> (32 * h)), m2 = (uint32_t)(M2 >> (32 * h));
        const uint32_t m3 = (uint32_t)(M3 >> (32 * h)), m4 = (uint32_t)(M4 >> (32 * h));
        const uint32_t hbase = J.sbase + (uint32_t)(32 * h) * J.sstride;
        uint32_t U = m1 | m2 | m3 | m4;
        while (U) {
            const uint32_t c = (uint32_t)thighest(U);
            const uint32_t b = 1u << c;
            U ^= b;
            const uint32_t addr = c * J.sstride + hbase;
            uint32_t lo, hi;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr));
            asm("{\n\t.reg .pred p, q, r, s;\n\t"
                "setp.ne.u32 p, %2, 0;\n\tsetp.ne.u32 q, %3, 0;\n\tsetp.ne.u32 r, %4, 0;\n\tsetp.ne.u32 s, %5, 0;\n\t"
                "@p xor.b32 %0, %0, %6;\n\t@p xor.b32 %1, %1, %7;\n\t"
                "@q xor.b32 %0, %0, %8;\n\t@q xor.b32 %1, %1, %9;\n\t"
                "@r xor.b32 %0, %0, %10;\n\t@r xor.b32 %1, %1, %11;\n\t"
                "@s xor.b32 %0, %0, %12;\n\t@s xor.b32 %1, %1, %13;\n\t}"
                : "+r"(lo), "+r"(hi) : "r"(m1 & b), "r"(m2 & b), "r"(m3 & b), "r"(m4 & b),
                  "r"(v1l), "r"(v1h), "r"(v2l), "r"(v2h), "r"(v3l), "r"(v3h), "r"(v4l), "r"(v4h));
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(addr), "r"(lo), "r"(hi) : "memory");
        }
    }
#else
    uint64_t U = M1 | M2 | M3 | M4;
    while (U) {
        const int c = thighest(U);
        const uint64_t b = 1ull << c;
        U ^= b;
        uint64_t r = J.get(c);
        if (M1 & b) r ^= V1;
        if (M2 & b) r ^= V2;
        if (M3 & b) r ^= V3;
        if (M4 & b) r ^= V4;
        J.put(c, r);
        BG_WORK(rows, 1); BG_WORK(xors, ((M1 & b) ? 1 : 0) + ((M2 & b) ? 1 : 0) + ((M3 & b) ? 1 : 0) + ((M4 & b) ? 1 : 0));
    }
#endif
}

// ... and three updates at once (the first pass of the rounds carries the pending fold along)
BG_HD void t_xor6(const Rows<uint32_t>& J, uint32_t M1, uint32_t V1, uint32_t M2, uint32_t V2, uint32_t M3, uint32_t V3,
                  uint32_t M4, uint32_t V4, uint32_t M5, uint32_t V5, uint32_t M6, uint32_t V6) {
    uint32_t U = M1 | M2 | M3 | M4 | M5 | M6;
    BG_TRACE(6000 + tpopc(U), 0);
#if defined(__CUDA_ARCH__)
    while (U) {
        const uint32_t c = (uint32_t)thighest(U);
        const uint32_t b = 1u << c;
        U ^= b;
        const uint32_t addr = c * J.sstride + J.sbase;
        uint32_t r;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
        asm("{\n\t.reg .pred p, q, r, s, t, u;\n\t"
            "setp.ne.u32 p, %1, 0;\n\tsetp.ne.u32 q, %2, 0;\n\tsetp.ne.u32 r, %3, 0;\n\tsetp.ne.u32 s, %4, 0;\n\t"
            "setp.ne.u32 t, %5, 0;\n\tsetp.ne.u32 u, %6, 0;\n\t"
            "@p xor.b32 %0, %0, %7;\n\t@q xor.b32 %0, %0, %8;\n\t@r xor.b32 %0, %0, %9;\n\t@s xor.b32 %0, %0, %10;\n\t"
            "@t xor.b32 %0, %0, %11;\n\t@u xor.b32 %0, %0, %12;\n\t}"
            : "+r"(r) : "r"(M1 & b), "r"(M2 & b), "r"(M3 & b), "r"(M4 & b), "r"(M5 & b), "r"(M6 & b),
              "r"(V1), "r"(V2), "r"(V3), "r"(V4), "r"(V5), "r"(V6));
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
        if (M3 & b) r ^= V3;
        if (M4 & b) r ^= V4;
        if (M5 & b) r ^= V5;
        if (M6 & b) r ^= V6;
        J.put(c, r);
        BG_WORK(rows, 1);
        BG_WORK(xors, ((M1 & b) ? 1 : 0) + ((M2 & b) ? 1 : 0) + ((M3 & b) ? 1 : 0) + ((M4 & b) ? 1 : 0) + ((M5 & b) ? 1 : 0) + ((M6 & b) ? 1 : 0));
    }
#endif
}
BG_HD void t_xor6(const Rows<uint64_t>& J, uint64_t M1, uint64_t V1, uint64_t M2, uint64_t V2, uint64_t M3, uint64_t V3,
                  uint64_t M4, uint64_t V4, uint64_t M5, uint64_t V5, uint64_t M6, uint64_t V6) {
    BG_TRACE(6000 + tpopc((uint32_t)(M1 | M2 | M3 | M4 | M5 | M6)), tpopc((uint32_t)((M1 | M2 | M3 | M4 | M5 | M6) >> 32)));
#if defined(__CUDA_ARCH__)
    const uint32_t v1l = (uint32_t)V1, v1h = (uint32_t)(V1 >> 32), v2l = (uint32_t)V2, v2h = (uint32_t)(V2 >> 32);
    const uint32_t v3l = (uint32_t)V3, v3h = (uint32_t)(V3 >> 32), v4l = (uint32_t)V4, v4h = (uint32_t)(V4 >> 32);
    const uint32_t v5l = (uint32_t)V5, v5h = (uint32_t)(V5 >> 32), v6l = (uint32_t)V6, v6h = (uint32_t)(V6 >> 32);
#pragma unroll
    for (int h = 1; h >= 0; h--) {
        const uint32_t m1 = (uint32_t)(M1 >> (32 * h)), m2 = (uint32_t)(M2 >> (32 * h)), m3 = (uint32_t)(M3 >> (32 * h));
        const uint32_t m4 = (uint32_t)(M4 >> (32 * h)), m5 = (uint32_t)(M5 >> (32 * h)), m6 = (uint32_t)(M6 >> (32 * h));
        const uint32_t hbase = J.sbase + (uint32_t)(32 * h) * J.sstride;
        uint32_t U = m1 | m2 | m3 | m4 | m5 | m6;
        while (U) {
            const uint32_t c = (uint32_t)thighest(U);
            const uint32_t b = 1u << c;
            U ^= b;
            const uint32_t addr = c * J.sstride + hbase;
            uint32_t lo, hi;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr));
            asm("{\n\t.reg .pred p, q, r, s, t, u;\n\t"
                "setp.ne.u32 p, %2, 0;\n\tsetp.ne.u32 q, %3, 0;\n\tsetp.ne.u32 r, %4, 0;\n\tsetp.ne.u32 s, %5, 0;\n\t"
                "setp.ne.u32 t, %6, 0;\n\tsetp.ne.u32 u, %7, 0;\n\t"
                "@p xor.b32 %0, %0, %8;\n\t@p xor.b32 %1, %1, %9;\n\t"
                "@q xor.b32 %0, %0, %10;\n\t@q xor.b32 %1, %1, %11;\n\t"
                "@r xor.b32 %0, %0, %12;\n\t@r xor.b32 %1, %1, %13;\n\t"
                "@s xor.b32 %0, %0, %14;\n\t@s xor.b32 %1, %1, %15;\n\t"
                "@t xor.b32 %0, %0, %16;\n\t@t xor.b32 %1, %1, %17;\n\t"
                "@u xor.b32 %0, %0, %18;\n\t@u xor.b32 %1, %1, %19;\n\t}"
                : "+r"(lo), "+r"(hi) : "r"(m1 & b), "r"(m2 & b), "r"(m3 & b), "r"(m4 & b), "r"(m5 & b), "r"(m6 & b),
                  "r"(v1l), "r"(v1h), "r"(v2l), "r"(v2h), "r"(v3l), "r"(v3h), "r"(v4l), "r"(v4h),
                  "r"(v5l), "r"(v5h), "r"(v6l), "r"(v6h));
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(addr), "r"(lo), "r"(hi) : "memory");
        }
    }
#else
    uint64_t U = M1 | M2 | M3 | M4 | M5 | M6;
    while (U) {
        const int c = thighest(U);
        const uint64_t b = 1ull << c;
        U ^= b;
        uint64_t r = J.get(c);
        if (M1 & b) r ^= V1;
        if (M2 & b) r ^= V2;
        if (M3 & b) r ^= V3;
        if (M4 & b) r ^= V4;
        if (M5 & b) r ^= V5;
        if (M6 & b) r ^= V6;
        J.put(c, r);
        BG_WORK(rows, 1);
        BG_WORK(xors, ((M1 & b) ? 1 : 0) + ((M2 & b) ? 1 : 0) + ((M3 & b) ? 1 : 0) + ((M4 & b) ? 1 : 0) + ((M5 & b) ? 1 : 0) + ((M6 & b) ? 1 : 0));
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

// =====================================================================================================
// The monomer / dimer steps of the exponential sum on the variables in E (all with D in {0,4}).
//
// * TWO steps per pass over the rows (t_xor4): step 2 picks its variables from the rows a2, b2 brought up
//   to date with step 1 on the fly; every other remaining row gets both updates in one touch.  A row is
//   touched when any of its four bits is set (15/16 of the rows instead of 3/4, twice), so the trip
//   counts of the 32 lanes of a warp are nearly equal.
// * No branch in the body: a monomer {a} is a dimer whose update masks are empty, and a missing second
//   step is a step with all masks empty — the lanes of a warp stay together.
// * The fold that precedes the rounds (t_expsum) is not applied to the rows either: it is handed over as a
//   PENDING update (TPend) and rides along with the first pass (t_xor6), fixed up on the fly in the rows
//   that pass reads.
// * Bits of D2 / Js are taken with one (funnel) shift and kept as "bit 0 of a word" (upper bits are junk
//   until the end); conditional xors are predicated.
// =====================================================================================================
template <typename W> struct TPend { W M1, V1, M2, V2; };     // row_c ^= [c in M1] V1 ^ [c in M2] V2, not yet applied

BG_HD void t_cxor32(uint32_t& x, uint32_t c, uint32_t v) {            // x ^= v if c & 1
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %1, 1;\n\tsetp.ne.u32 p, t, 0;\n\t@p xor.b32 %0, %0, %2;\n\t}"
        : "+r"(x) : "r"(c), "r"(v));
#else
    if (c & 1u) x ^= v;
#endif
}
BG_HD void t_step_scalars32(uint32_t& D2, uint32_t& Js, uint32_t ia, uint32_t ib, uint32_t on, uint32_t onm, uint32_t dm,
                            uint32_t M1, uint32_t M2, uint32_t& cnt, uint32_t& neg0, uint32_t& neg1, uint32_t& z0, uint32_t& z1) {
    const uint32_t d2a = D2 >> ia, sa = Js >> ia, d2b = D2 >> ib, sb = Js >> ib;      // bit 0; upper bits junk
    const uint32_t ta = d2a ^ sa;
    cnt += on;
    neg0 ^= d2a & d2b & dm;
    neg1 ^= ta & (d2b ^ sb) & dm;
    z0 |= d2a & ~dm & onm;
    z1 |= ta & ~dm & onm;
    t_cxor32(D2, d2b, M1); t_cxor32(D2, d2a, M2); D2 ^= M1 & M2;
    t_cxor32(Js, sb, M1); t_cxor32(Js, sa, M2);
}
// row `i` as the pending update leaves it
template <bool PEND> BG_HD uint32_t t_row32(const Rows<uint32_t>& J, uint32_t i, const TPend<uint32_t>& pd) {
    uint32_t r = J.get((int)i);
    if (PEND) { t_cxor32(r, pd.M1 >> i, pd.V1); t_cxor32(r, pd.M2 >> i, pd.V2); }
    return r;
}
// two steps + one pass (32-bit words)
template <bool PEND>
BG_HD void t_block32(const Rows<uint32_t>& J, uint32_t& E, uint32_t& D2, uint32_t& Js, uint32_t& cnt, uint32_t& neg0,
                     uint32_t& neg1, uint32_t& z0, uint32_t& z1, uint32_t ns, const TPend<uint32_t>& pd) {
    // ---- step 1: a = highest variable left, b = its highest neighbour (a itself: monomer)
    const uint32_t a = (uint32_t)thighest(E), ba = 1u << a;
    const uint32_t Ja = t_row32<PEND>(J, a, pd) & E & ~ba;
    const bool dim1 = Ja != 0u;
    const uint32_t dm1 = dim1 ? ~0u : 0u;
    const uint32_t b = (uint32_t)thighest(dim1 ? Ja : ba), bb = 1u << b;
    const uint32_t r1 = E & ~(ba | bb);
    const uint32_t M1 = Ja & r1;
    const uint32_t M2 = t_row32<PEND>(J, b, pd) & r1 & dm1;
    BG_WORK(dimers, dim1 ? 1 : 0); BG_WORK(monomers, dim1 ? 0 : 1);
    t_step_scalars32(D2, Js, a, b, 1u, ~0u, dm1, M1, M2, cnt, neg0, neg1, z0, z1);
    // ---- step 2 on what is left (nothing, if the sum is already known to vanish)
    const bool go2 = ((z0 & (z1 | ns) & 1u) == 0u) & (r1 != 0u);
    const uint32_t left = go2 ? r1 : 0u;
    const uint32_t a2 = (uint32_t)thighest(go2 ? r1 : ba), ba2 = 1u << a2;
    uint32_t q = t_row32<PEND>(J, a2, pd);
    t_cxor32(q, M1 >> a2, M2); t_cxor32(q, M2 >> a2, M1);               // row a2 brought up to date with step 1
    const uint32_t Ka = q & left & ~ba2;
    const bool dim2 = Ka != 0u;
    const uint32_t dm2 = dim2 ? ~0u : 0u;
    const uint32_t b2 = (uint32_t)thighest(dim2 ? Ka : ba2), bb2 = 1u << b2;
    uint32_t r = t_row32<PEND>(J, b2, pd);
    t_cxor32(r, M1 >> b2, M2); t_cxor32(r, M2 >> b2, M1);
    const uint32_t r2 = left & ~(ba2 | bb2);
    const uint32_t M3 = Ka & r2;
    const uint32_t M4 = r & r2 & dm2;
    BG_WORK(dimers, dim2 ? 1 : 0); BG_WORK(monomers, (go2 && !dim2) ? 1 : 0);
    t_step_scalars32(D2, Js, a2, b2, go2 ? 1u : 0u, go2 ? ~0u : 0u, dm2, M3, M4, cnt, neg0, neg1, z0, z1);
    // ---- the updates on the rows that stay:  J_c ^= [J_ca] J_b ^ [J_cb] J_a, twice (+ the pending one)
    if (PEND) t_xor6(J, M1 & r2, M2, M2 & r2, M1, M3, M4, M4, M3, pd.M1 & r2, pd.V1, pd.M2 & r2, pd.V2);
    else t_xor4(J, M1 & r2, M2, M2 & r2, M1, M3, M4, M4, M3);
    E = (z0 & (z1 | ns) & 1u) ? 0u : r2;
}
BG_HD void t_rounds(const Rows<uint32_t>& J, uint32_t& E, uint32_t& D2, uint32_t& Js, uint32_t& cnt, uint32_t& neg0,
                    uint32_t& neg1, uint32_t& z0, uint32_t& z1, bool has_s, const TPend<uint32_t>& pd, bool pending = true) {
    const uint32_t ns = has_s ? 0u : 1u;
    if (pending && E != 0u) t_block32<true>(J, E, D2, Js, cnt, neg0, neg1, z0, z1, ns, pd);
    while (E != 0u) t_block32<false>(J, E, D2, Js, cnt, neg0, neg1, z0, z1, ns, pd);
    neg0 &= 1u; neg1 &= 1u; z0 &= 1u; z1 &= 1u;
}

// ---- the same for 64-bit words (t > 32), written on 32-bit halves: every 64-bit test / select / shift
// would cost the compiler two or three instructions; here a variable is (bit as two halves, index, row
// address).
struct TIdx {
    uint32_t bl, bh;       // the variable's bit, low and high half
    uint32_t idx;          // 0..63
    uint32_t addr;         // device: shared-memory address of its row
};
BG_HD TIdx t_top64(uint32_t xl, uint32_t xh, const Rows<uint64_t>& J) {      // highest variable of (xh:xl) != 0
    TIdx r;
    const bool ph = xh != 0u;
    const uint32_t c = (uint32_t)thighest(ph ? xh : xl);
    const uint32_t bm = 1u << c;
    r.bl = ph ? 0u : bm; r.bh = ph ? bm : 0u;
    r.idx = c + (ph ? 32u : 0u);
    r.addr = c * J.sstride + (ph ? J.sbase + 32u * J.sstride : J.sbase);
    return r;
}
BG_HD uint32_t t_bit64(uint32_t l, uint32_t h, uint32_t idx) {             // bit idx of (h:l) -> bit 0 (rest: junk)
    return (uint32_t)((((uint64_t)h << 32) | l) >> idx);
}
BG_HD void t_cxor64(uint32_t& l, uint32_t& h, uint32_t c, uint32_t vl, uint32_t vh) {    // (h:l) ^= (vh:vl) if c & 1
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %2, 1;\n\tsetp.ne.u32 p, t, 0;\n\t"
        "@p xor.b32 %0, %0, %3;\n\t@p xor.b32 %1, %1, %4;\n\t}" : "+r"(l), "+r"(h) : "r"(c), "r"(vl), "r"(vh));
#else
    if (c & 1u) { l ^= vl; h ^= vh; }
#endif
}
struct TPend64 { uint32_t M1l, M1h, V1l, V1h, M2l, M2h, V2l, V2h; };
template <bool PEND> BG_HD void t_ld64(const Rows<uint64_t>& J, const TIdx& i, const TPend64& pd, uint32_t& l, uint32_t& h) {
#if defined(__CUDA_ARCH__)
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(l), "=r"(h) : "r"(i.addr));
#else
    const uint64_t v = J.get((int)i.idx); l = (uint32_t)v; h = (uint32_t)(v >> 32);
#endif
    if (PEND) {
        t_cxor64(l, h, t_bit64(pd.M1l, pd.M1h, i.idx), pd.V1l, pd.V1h);
        t_cxor64(l, h, t_bit64(pd.M2l, pd.M2h, i.idx), pd.V2l, pd.V2h);
    }
}
// bookkeeping of one step: on = 0/1, onm / dm = all-ones masks for "step exists" / "dimer"
BG_HD void t_step_scalars64(uint32_t& D2l, uint32_t& D2h, uint32_t& Jsl, uint32_t& Jsh, uint32_t ia, uint32_t ib,
                            uint32_t on, uint32_t onm, uint32_t dm, uint32_t M1l, uint32_t M1h, uint32_t M2l, uint32_t M2h,
                            uint32_t& cnt, uint32_t& neg0, uint32_t& neg1, uint32_t& z0, uint32_t& z1) {
    const uint32_t d2a = t_bit64(D2l, D2h, ia), sa = t_bit64(Jsl, Jsh, ia);
    const uint32_t d2b = t_bit64(D2l, D2h, ib), sb = t_bit64(Jsl, Jsh, ib);
    const uint32_t ta = d2a ^ sa;
    cnt += on;
    neg0 ^= d2a & d2b & dm;
    neg1 ^= ta & (d2b ^ sb) & dm;
    z0 |= d2a & ~dm & onm;
    z1 |= ta & ~dm & onm;
    t_cxor64(D2l, D2h, d2b, M1l, M1h);
    t_cxor64(D2l, D2h, d2a, M2l, M2h);
    D2l ^= M1l & M2l; D2h ^= M1h & M2h;
    t_cxor64(Jsl, Jsh, sb, M1l, M1h);
    t_cxor64(Jsl, Jsh, sa, M2l, M2h);
}
BG_HD uint64_t t_mk64(uint32_t l, uint32_t h) { return ((uint64_t)h << 32) | l; }

template <bool PEND>
BG_HD void t_block64(const Rows<uint64_t>& J, uint32_t& El, uint32_t& Eh, uint32_t& D2l, uint32_t& D2h, uint32_t& Jsl,
                     uint32_t& Jsh, uint32_t& cnt, uint32_t& neg0, uint32_t& neg1, uint32_t& z0, uint32_t& z1, uint32_t ns,
                     const TPend64& pd) {
    // ---- step 1
    const TIdx a = t_top64(El, Eh, J);
    uint32_t ral, rah;
    t_ld64<PEND>(J, a, pd, ral, rah);
    const uint32_t Jal = ral & El & ~a.bl, Jah = rah & Eh & ~a.bh;
    const bool dim1 = (Jal | Jah) != 0u;
    const uint32_t dm1 = dim1 ? ~0u : 0u;
    const TIdx b = t_top64(dim1 ? Jal : a.bl, dim1 ? Jah : a.bh, J);
    uint32_t rbl, rbh;
    t_ld64<PEND>(J, b, pd, rbl, rbh);
    const uint32_t r1l = El & ~(a.bl | b.bl), r1h = Eh & ~(a.bh | b.bh);
    const uint32_t M1l = Jal & r1l, M1h = Jah & r1h;
    const uint32_t M2l = rbl & r1l & dm1, M2h = rbh & r1h & dm1;
    BG_WORK(dimers, dim1 ? 1 : 0); BG_WORK(monomers, dim1 ? 0 : 1);
    t_step_scalars64(D2l, D2h, Jsl, Jsh, a.idx, b.idx, 1u, ~0u, dm1, M1l, M1h, M2l, M2h, cnt, neg0, neg1, z0, z1);
    // ---- step 2
    const bool go2 = ((z0 & (z1 | ns) & 1u) == 0u) & ((r1l | r1h) != 0u);
    const uint32_t ll = go2 ? r1l : 0u, lh = go2 ? r1h : 0u;
    const TIdx a2 = t_top64(go2 ? r1l : a.bl, go2 ? r1h : a.bh, J);
    uint32_t ql, qh;
    t_ld64<PEND>(J, a2, pd, ql, qh);
    t_cxor64(ql, qh, t_bit64(M1l, M1h, a2.idx), M2l, M2h);              // row a2 brought up to date with step 1
    t_cxor64(ql, qh, t_bit64(M2l, M2h, a2.idx), M1l, M1h);
    const uint32_t Kal = ql & ll & ~a2.bl, Kah = qh & lh & ~a2.bh;
    const bool dim2 = (Kal | Kah) != 0u;
    const uint32_t dm2 = dim2 ? ~0u : 0u;
    const TIdx b2 = t_top64(dim2 ? Kal : a2.bl, dim2 ? Kah : a2.bh, J);
    uint32_t sl, sh;
    t_ld64<PEND>(J, b2, pd, sl, sh);
    t_cxor64(sl, sh, t_bit64(M1l, M1h, b2.idx), M2l, M2h);
    t_cxor64(sl, sh, t_bit64(M2l, M2h, b2.idx), M1l, M1h);
    const uint32_t r2l = ll & ~(a2.bl | b2.bl), r2h = lh & ~(a2.bh | b2.bh);
    const uint32_t M3l = Kal & r2l, M3h = Kah & r2h;
    const uint32_t M4l = sl & r2l & dm2, M4h = sh & r2h & dm2;
    BG_WORK(dimers, dim2 ? 1 : 0); BG_WORK(monomers, (go2 && !dim2) ? 1 : 0);
    t_step_scalars64(D2l, D2h, Jsl, Jsh, a2.idx, b2.idx, go2 ? 1u : 0u, go2 ? ~0u : 0u, dm2, M3l, M3h, M4l, M4h,
                     cnt, neg0, neg1, z0, z1);
    // ---- the updates on the rows that stay
    if (PEND)
        t_xor6(J, t_mk64(M1l & r2l, M1h & r2h), t_mk64(M2l, M2h), t_mk64(M2l & r2l, M2h & r2h), t_mk64(M1l, M1h),
               t_mk64(M3l, M3h), t_mk64(M4l, M4h), t_mk64(M4l, M4h), t_mk64(M3l, M3h),
               t_mk64(pd.M1l & r2l, pd.M1h & r2h), t_mk64(pd.V1l, pd.V1h), t_mk64(pd.M2l & r2l, pd.M2h & r2h), t_mk64(pd.V2l, pd.V2h));
    else
        t_xor4(J, t_mk64(M1l & r2l, M1h & r2h), t_mk64(M2l, M2h), t_mk64(M2l & r2l, M2h & r2h), t_mk64(M1l, M1h),
               t_mk64(M3l, M3h), t_mk64(M4l, M4h), t_mk64(M4l, M4h), t_mk64(M3l, M3h));
    const bool stop = (z0 & (z1 | ns) & 1u) != 0u;
    El = stop ? 0u : r2l; Eh = stop ? 0u : r2h;
}

BG_HD void t_rounds(const Rows<uint64_t>& J, uint64_t& E, uint64_t& D2, uint64_t& Js, uint32_t& cnt, uint32_t& neg0,
                    uint32_t& neg1, uint32_t& z0, uint32_t& z1, bool has_s, const TPend<uint64_t>& pd64, bool pending = true) {
    uint32_t El = (uint32_t)E, Eh = (uint32_t)(E >> 32);
    uint32_t D2l = (uint32_t)D2, D2h = (uint32_t)(D2 >> 32), Jsl = (uint32_t)Js, Jsh = (uint32_t)(Js >> 32);
    const uint32_t ns = has_s ? 0u : 1u;
    TPend64 pd;
    pd.M1l = (uint32_t)pd64.M1; pd.M1h = (uint32_t)(pd64.M1 >> 32); pd.V1l = (uint32_t)pd64.V1; pd.V1h = (uint32_t)(pd64.V1 >> 32);
    pd.M2l = (uint32_t)pd64.M2; pd.M2h = (uint32_t)(pd64.M2 >> 32); pd.V2l = (uint32_t)pd64.V2; pd.V2h = (uint32_t)(pd64.V2 >> 32);
    // Variables are eliminated from the top, so the high halves die first: once no lane of the warp has a
    // variable >= 32 left, the rounds continue on the low halves of the same rows with 32-bit code.
#if defined(__CUDA_ARCH__)
#define T_ALL_LOW() __all_sync(__activemask(), Eh == 0u)
#else
#define T_ALL_LOW() (Eh == 0u)
#endif
    if (pending && (El | Eh) != 0u && !T_ALL_LOW()) {
        t_block64<true>(J, El, Eh, D2l, D2h, Jsl, Jsh, cnt, neg0, neg1, z0, z1, ns, pd);
        pending = false;
    }
    while ((El | Eh) != 0u) {
        if (T_ALL_LOW()) break;
        t_block64<false>(J, El, Eh, D2l, D2h, Jsl, Jsh, cnt, neg0, neg1, z0, z1, ns, pd);
    }
#undef T_ALL_LOW
    neg0 &= 1u; neg1 &= 1u; z0 &= 1u; z1 &= 1u;
    if (El != 0u) {
        Rows<uint32_t> Jl;                       // the low words of the same rows
        Jl.base = reinterpret_cast<uint32_t*>(J.base); Jl.stride = 2 * J.stride;
        Jl.sbase = J.sbase; Jl.sstride = J.sstride;
        TPend<uint32_t> pl;
        pl.M1 = pd.M1l; pl.V1 = pd.V1l; pl.M2 = pd.M2l; pl.V2 = pd.V2l;
        t_rounds(Jl, El, D2l, Jsl, cnt, neg0, neg1, z0, z1, has_s, pl, pending);
    }
    E = 0; D2 = t_mk64(D2l, D2h); Js = t_mk64(Jsl, Jsh);
}

// sum over F_2^A of e^{i pi q/4}    (bg_device.cuh: expsum)
template <typename W> BG_HD void t_expsum(const Rows<W>& J, TF<W>& f, int& eps, int& p, int& m) {
    const W A = f.A;
    const W S = f.D1 & A;
    const bool has_s = S != 0;
    W E = A, Js = 0;
    uint32_t Ds = 0;
    TPend<W> pd;
    pd.M1 = pd.V1 = pd.M2 = pd.V2 = 0;
    if (has_s) {
        const int s = thighest(S);
        const W bs = tbit<W>(s), Sp = S ^ bs;
        Ds = 2u + 4u * tget(f.D2, s);
        // the fold x_s = x'_s + sum_{a in Sp} x'_a (t_basis_change) — its row update stays pending
        const W Ji = J.get(s);
        const W col = (Ji ^ ((Ji & bs) ? Sp : (W)0)) & f.A;
        pd.M1 = Sp; pd.V1 = Ji; pd.M2 = Sp ? col : (W)0; pd.V2 = Sp;
        BG_WORK(basis_changes, Sp ? 1 : 0);
        const W d1s = tfill<W>(tget(f.D1, s)), d2s = tfill<W>(tget(f.D2, s));
        f.D2 ^= Sp & (d2s ^ (d1s & f.D1) ^ Ji);
        f.D1 ^= Sp & d1s;
        E = A & ~bs;
        Js = (Ji ^ ((col & bs) ? Sp : (W)0)) & E;        // row s after the fold (s is not in Sp)
    }
    W D2 = f.D2;
    uint32_t cnt = 0, neg0 = 0, neg1 = 0, z0 = 0, z1 = 0;
    t_rounds(J, E, D2, Js, cnt, neg0, neg1, z0, z1, has_s, pd);
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


// =====================================================================================================
// Odd-first elimination (round 2): the exponential sum without the fold and without the sigma = 0 / 1
// double bookkeeping of exponentialSumExact (stabilizer.c:300-481); same value, hence the same (eps, p, m mod 8).
//
// * A variable a with D_a in {2,6} (D1_a = 1) is summed out on its own:
//       sum_{x_a} w^{D_a x_a + 4 x_a l(x)} = 1 + i^d (-1)^{l(x)} = sqrt2 w^d i^{-d (l(x) mod 2)},   d = +-1, w = e^{i pi/4}
//   and l mod 2 = l^2 mod 4, so with v = J_a restricted to the variables left:  Q += d, p += 1, D_c -= 2d for c in v,
//   J ^= v v^T — a RANK-ONE update whose mask and value are the same word.  With diag(J) = D1 the row update
//   row_c ^= v (c in v) also flips D1; D2_c ^= (d = +1 ? ~D1_c : D1_c) takes the carry.
// * K such steps share ONE pass over the rows (t_xork): the pivot row of step j is brought up to date with the
//   j - 1 earlier steps of its block in registers.  No branch in a block: a step with no odd variable left has
//   an empty mask.
// * Only when no variable with D in {2,6} is left does the monomer / dimer elimination (t_rounds, has_s = false)
//   finish the sum; for a random theta that is the last variable or two.
// * Parity checks of K_theta need no pivoting here: check j is a Lagrange variable lambda_j with D = 4 beta_j and
//   row = the check (1/2 sum_lambda (-1)^{lambda (c.x + beta)} = [c.x = beta]); the warp appends these rows and
//   columns to the ambient form once per sample (k_pairs_tpp) and the term only adds the lambda bits to its
//   active set; p -= 2 per check.
// =====================================================================================================
BG_HD void t_pxor32(uint32_t& r, uint32_t c, uint32_t v) {                    // r ^= v if c != 0
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p xor.b32 %0, %0, %2;\n\t}" : "+r"(r) : "r"(c), "r"(v));
#else
    if (c) r ^= v;
#endif
}
BG_HD void t_pxor64(uint32_t& l, uint32_t& h, uint32_t c, uint32_t vl, uint32_t vh) {      // (h:l) ^= (vh:vl) if c != 0
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p xor.b32 %0, %0, %3;\n\t@p xor.b32 %1, %1, %4;\n\t}"
        : "+r"(l), "+r"(h) : "r"(c), "r"(vl), "r"(vh));
#else
    if (c) { l ^= vl; h ^= vh; }
#endif
}

// row_c ^= xor of the v_j that contain c, for every row c in U  (32-bit words)
template <int K> BG_HD void t_xork(const Rows<uint32_t>& J, uint32_t U, const uint32_t (&v)[K]) {
    BG_TRACE(1000 * K + tpopc(U), 0);
    while (U) {
        const uint32_t c = (uint32_t)thighest(U);
        const uint32_t b = 1u << c;
        U ^= b;
#if defined(__CUDA_ARCH__)
        const uint32_t addr = c * J.sstride + J.sbase;
        uint32_t r;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
#pragma unroll
        for (int j = 0; j < K; j++) t_pxor32(r, v[j] & b, v[j]);
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(r) : "memory");
#else
        uint32_t r = J.get((int)c);
        for (int j = 0; j < K; j++) { if (v[j] & b) { r ^= v[j]; BG_WORK(xors, 1); } }
        J.put((int)c, r);
        BG_WORK(rows, 1);
#endif
    }
}
// ... and 64-bit words, as two halves
template <int K> BG_HD void t_xork(const Rows<uint64_t>& J, uint32_t Ul, uint32_t Uh, const uint32_t (&vl)[K], const uint32_t (&vh)[K]) {
    BG_TRACE(1000 * K + tpopc(Ul), tpopc(Uh));
#pragma unroll
    for (int h = 1; h >= 0; h--) {
        uint32_t U = h ? Uh : Ul;
#if defined(__CUDA_ARCH__)
        const uint32_t hbase = J.sbase + (uint32_t)(32 * h) * J.sstride;
#endif
        while (U) {
            const uint32_t c = (uint32_t)thighest(U);
            const uint32_t b = 1u << c;
            U ^= b;
#if defined(__CUDA_ARCH__)
            const uint32_t addr = c * J.sstride + hbase;
            uint32_t lo, hi;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr));
#pragma unroll
            for (int j = 0; j < K; j++) t_pxor64(lo, hi, (h ? vh[j] : vl[j]) & b, vl[j], vh[j]);
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(addr), "r"(lo), "r"(hi) : "memory");
#else
            uint64_t r = J.get((int)c + 32 * h);
            for (int j = 0; j < K; j++) { if ((h ? vh[j] : vl[j]) & b) { r ^= t_mk64(vl[j], vh[j]); BG_WORK(xors, 1); } }
            J.put((int)c + 32 * h, r);
            BG_WORK(rows, 1);
#endif
        }
    }
}

// K rank-one steps + one pass (32-bit words)
template <int K>
BG_HD void t_oddblock32(const Rows<uint32_t>& J, uint32_t& E, uint32_t& D1, uint32_t& D2, uint32_t& Q, uint32_t& cnt) {
    uint32_t v[K];
    uint32_t U = 0u;
#pragma unroll
    for (int j = 0; j < K; j++) {
        const uint32_t O = D1 & E;
        const bool on = O != 0u;
        const uint32_t a = (uint32_t)thighest(on ? O : 1u);
        const uint32_t ba = on ? (1u << a) : 0u;
        uint32_t r = J.get((int)a);
#pragma unroll
        for (int i = 0; i < j; i++) t_pxor32(r, v[i] & ba, v[i]);              // row a as steps < j of this block leave it
        E &= ~ba;
        v[j] = on ? (r & E) : 0u;
        const bool neg = (D2 & ba) != 0u;                                    // D_a = 6: d = -1
        Q += (on ? 1u : 0u) + (neg ? 6u : 0u);
        cnt += on ? 1u : 0u;
        D2 ^= v[j] & (D1 ^ (neg ? 0u : ~0u));
        D1 ^= v[j];
        U |= v[j];
        BG_WORK(monomers, on ? 1 : 0);
    }
    t_xork<K>(J, U & E, v);
}

// K rank-one steps + one pass (64-bit words as halves).  Returns true when the pass was done on the low words
// only because no lane of the warp has a variable >= 32 left (the caller continues with 32-bit code).
#if defined(__CUDA_ARCH__)
#define T_WARP_ALL(x) __all_sync(__activemask(), (x))
#define T_WARP_ANY(x) __any_sync(__activemask(), (x))
#elif defined(BG_EMU_FORCE_ANY)       // CPU build that takes every "some other lane needs it" branch
#define T_WARP_ALL(x) (false)
#define T_WARP_ANY(x) (true)
#else
#define T_WARP_ALL(x) (x)
#define T_WARP_ANY(x) (x)
#endif
// While a lane has variables >= 32 left the pivot is the TOP variable, odd or not, so that the high halves die
// within one block: an even top variable a borrows its oddness from a fresh variable mu (D_mu = 2, no
// couplings, sum_{x_mu} = sqrt2 w — divided out: cnt -= 1, Q -= 1) placed at a FREE low slot f of this term
// (x_mu = x'_mu + x_a makes a odd and adds bit f to its row; the rank-one step then leaves x'_mu as an even
// variable with a's couplings).  `fr` = free low slots whose columns are zero in every row (t_copy_in_masked);
// fr = 0 switches the borrowing off (then: highest odd variable, as in the 32-bit code).
template <int K>
BG_HD bool t_oddblock64(const Rows<uint64_t>& J, uint32_t& El, uint32_t& Eh, uint32_t& D1l, uint32_t& D1h, uint32_t& D2l,
                        uint32_t& D2h, uint32_t& Q, uint32_t& cnt, uint32_t& fr) {
    uint32_t vl[K], vh[K];
    uint32_t Ul = 0u, Uh = 0u;
#pragma unroll
    for (int j = 0; j < K; j++) {
        const uint32_t Ol = D1l & El, Oh = D1h & Eh;
        const bool top = (Eh != 0u) & (fr != 0u);                            // pivot = top variable (it is >= 32)
        const bool on = top | ((Ol | Oh) != 0u);
        const TIdx a = t_top64(on ? Ol : 1u, top ? Eh : Oh, J);
        const uint32_t bl = on ? a.bl : 0u, bh = a.bh;                       // off: the high word is 0, so a.bh = 0
        uint32_t rl, rh;
#if defined(__CUDA_ARCH__)
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(rl), "=r"(rh) : "r"(a.addr));
#else
        { const uint64_t r = J.get((int)a.idx); rl = (uint32_t)r; rh = (uint32_t)(r >> 32); }
#endif
#pragma unroll
        for (int i = 0; i < j; i++) t_pxor64(rl, rh, (vl[i] & bl) | (vh[i] & bh), vl[i], vh[i]);
        const bool ev = top & ((D1h & bh) == 0u);                            // even top variable: borrow from mu at slot f
        const uint32_t f = (uint32_t)thighest(fr | 1u);
        const uint32_t fb = ev ? (1u << f) : 0u;
        if (ev) {
#if defined(__CUDA_ARCH__)
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(f * J.sstride + J.sbase), "r"(fb), "r"(0u) : "memory");
#else
            J.put((int)f, (uint64_t)fb);
#endif
        }
        fr &= ~fb;
        El = (El & ~bl) | fb; Eh &= ~bh;
        D1l |= fb; D2l &= ~fb;                                               // D_mu = 2
        vl[j] = on ? ((rl & El) | fb) : 0u; vh[j] = on ? (rh & Eh) : 0u;
        const bool neg = ((D2l & bl) | (D2h & bh)) != 0u;
        Q += (on ? 1u : 0u) + (neg ? 6u : 0u) + (ev ? 7u : 0u);
        cnt += (on ? 1u : 0u) - (ev ? 1u : 0u);
        const uint32_t dm = neg ? 0u : ~0u;
        D2l ^= vl[j] & (D1l ^ dm); D2h ^= vh[j] & (D1h ^ dm);
        D1l ^= vl[j]; D1h ^= vh[j];
        Ul |= vl[j]; Uh |= vh[j];
        BG_WORK(monomers, on ? 1 : 0);
    }
    if (T_WARP_ALL(Eh == 0u)) {
        Rows<uint32_t> Jl;                       // the low words of the same rows
        Jl.base = reinterpret_cast<uint32_t*>(J.base); Jl.stride = 2 * J.stride;
        Jl.sbase = J.sbase; Jl.sstride = J.sstride;
        t_xork<K>(Jl, Ul & El, vl);
        return true;
    }
    t_xork<K>(J, Ul & El, Uh & Eh, vl, vh);
    return false;
}

// What the odd phase leaves has D in {0,4}: for a random theta one or two variables (sizes 0 / 1 / 2 / 3 / 4 with
// probability 0.29 / 0.29 / 0.19 / 0.11 / 0.06).  The blocked rounds run only while more than two are left; one or
// two variables are finished in closed form (a monomer: factor 2 or 0; a coupled pair: a dimer; an uncoupled pair: two
// monomers) — one row load, no pass.
BG_HD void t_eventail32(const Rows<uint32_t>& J, uint32_t& E, uint32_t D1, uint32_t& D2, uint32_t& Q, uint32_t& cnt, uint32_t& cnt2,
                         uint32_t& neg0, uint32_t& z0) {
    if (tpopc(E) > 2) {                                  // only even variables are left here (the odd phase ran out of odd ones)
        uint32_t Js = 0, neg1 = 0, z1 = 0;
        while (tpopc(E) > 2) t_block32<false>(J, E, D2, Js, cnt2, neg0, neg1, z0, z1, 1u, TPend<uint32_t>());
        neg0 &= 1u; z0 &= 1u;
    }
    if (E != 0u) {
        // one or two variables of ANY kind (the odd phase stops at two): sum out an odd one first (rank-one step on a
        // single neighbour), then the other alone
        const uint32_t a = (uint32_t)thighest(E);
        const uint32_t ba = 1u << a, bb = E ^ ba;                       // bb: the other variable, if any
        const bool coupled = (J.get((int)a) & bb) != 0u;
        const bool oa = (D1 & ba) != 0u, ob = (D1 & bb) != 0u;
        const bool d2a = (D2 & ba) != 0u, d2b = (D2 & bb) != 0u;
        if (!oa && !ob) {
            if (coupled) { neg0 ^= (d2a && d2b) ? 1u : 0u; cnt2 += 1u; }
            else { z0 |= (d2a || d2b) ? 1u : 0u; cnt2 += bb ? 2u : 1u; }
            BG_WORK(dimers, coupled ? 1 : 0); BG_WORK(monomers, coupled ? 0 : (bb ? 2 : 1));
        } else {
            const bool first_a = oa;                                    // the odd variable that goes first
            const bool negp = first_a ? d2a : d2b;                      // D = 6: d = -1
            Q += negp ? 7u : 1u; cnt += 1u;
            if (bb) {
                bool oq = first_a ? ob : oa, d2q = first_a ? d2b : d2a;
                if (coupled) { d2q = d2q != (oq != !negp); oq = !oq; }  // D_q -= 2d: D2_q ^= (d = +1 ? ~D1_q : D1_q), D1_q ^= 1
                if (oq) { Q += d2q ? 7u : 1u; cnt += 1u; }
                else { z0 |= d2q ? 1u : 0u; cnt2 += 1u; }
            }
            BG_WORK(monomers, bb ? 2 : 1);
        }
        E = 0u;
    }
}

// the steps of the odd phase: blocks of 8 while any lane of the warp has more than 4 variables left
BG_HD void t_oddphase(const Rows<uint32_t>& J, uint32_t& E, uint32_t& D1, uint32_t& D2, uint32_t& Q, uint32_t& cnt) {
    while ((D1 & E) != 0u && tpopc(E) > 2) {            // one or two variables left: closed form (t_eventail32)
        if (T_WARP_ANY(tpopc(E) > 4)) t_oddblock32<8>(J, E, D1, D2, Q, cnt);
        else t_oddblock32<4>(J, E, D1, D2, Q, cnt);
    }
}

// sum over F_2^A of e^{i pi q/4}: odd-first.  Same result as t_expsum.
BG_HD void t_expsum_odd(const Rows<uint32_t>& J, TF<uint32_t>& f, int& eps, int& p, int& m, uint32_t fr = 0u) {
    (void)fr;
    uint32_t E = f.A, D1 = f.D1, D2 = f.D2, Q = f.Q, cnt = 0;
    t_oddphase(J, E, D1, D2, Q, cnt);
    uint32_t cnt2 = 0, neg0 = 0, neg1 = 0, z0 = 0, z1 = 0;
    t_eventail32(J, E, D1, D2, Q, cnt, cnt2, neg0, z0);  // even variables, or at most two of any kind
    eps = z0 ? 0 : 1;
    p = (int)cnt + 2 * (int)cnt2;
    m = (int)((Q + 4u * neg0) & 7u);
}
BG_HD void t_expsum_odd(const Rows<uint64_t>& J, TF<uint64_t>& f, int& eps, int& p, int& m, uint32_t fr = 0u) {
    uint32_t El = (uint32_t)f.A, Eh = (uint32_t)(f.A >> 32), D1l = (uint32_t)f.D1, D1h = (uint32_t)(f.D1 >> 32);
    uint32_t D2l = (uint32_t)f.D2, D2h = (uint32_t)(f.D2 >> 32), Q = f.Q, cnt = 0;
    if (tpopc(fr) < tpopc(Eh)) fr = 0u;                  // not enough free low slots: no borrowing for this term
    bool low = T_WARP_ALL(Eh == 0u);
    while (!low && ((((D1l & El) | (D1h & Eh)) != 0u) | ((Eh != 0u) & (fr != 0u)))) {
        if (T_WARP_ANY(tpopc(El) + tpopc(Eh) > 4)) low = t_oddblock64<8>(J, El, Eh, D1l, D1h, D2l, D2h, Q, cnt, fr);
        else low = t_oddblock64<4>(J, El, Eh, D1l, D1h, D2l, D2h, Q, cnt, fr);
    }
    uint32_t cnt2 = 0, neg0 = 0, neg1 = 0, z0 = 0, z1 = 0;
    if (Eh == 0u) {                                      // continue on the low words with 32-bit code
        Rows<uint32_t> Jl;
        Jl.base = reinterpret_cast<uint32_t*>(J.base); Jl.stride = 2 * J.stride;
        Jl.sbase = J.sbase; Jl.sstride = J.sstride;
        t_oddphase(Jl, El, D1l, D2l, Q, cnt);
        t_eventail32(Jl, El, D1l, D2l, Q, cnt, cnt2, neg0, z0);
    } else {                                             // even variables >= 32 are left and nothing odd: 64-bit rounds
        uint64_t E = t_mk64(El, Eh), D2 = t_mk64(D2l, D2h), Js = 0;
        TPend<uint64_t> pd; pd.M1 = pd.V1 = pd.M2 = pd.V2 = 0;
        t_rounds(J, E, D2, Js, cnt2, neg0, neg1, z0, z1, false, pd, false);
    }
    eps = z0 ? 0 : 1;
    p = (int)cnt + 2 * (int)cnt2;
    m = (int)((Q + 4u * neg0) & 7u);
}

// What a warp shares about its theta: the ambient form and at most TPP_MAXC parity checks.
#define TPP_MAXC 6
template <typename W> struct TShared {
    const W* J;          // t ambient rows (shared memory, read-only)
    W D1, D2;
    uint32_t Q;
    int k1, t;
    int ncons;           // parity checks (t - k1) that the term has to pivot
    int nlam;            // parity checks carried as Lagrange variables t .. t+nlam-1 of the ambient form (then ncons = 0)
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

// the thread's working copy of the ambient J.  Device: the warp's ambient rows are 16-byte aligned
// (k_pairs_tpp pads them), so they are read with 128-bit broadcast loads.  `keep`: mask applied to the low
// 32 columns (64-bit words only): columns of inactive variables are cleared so that they can serve as free
// slots (t_oddblock64).
template <typename W>
BG_HD void t_copy_in(const Rows<W>& J, const TShared<W>& sh, uint32_t keep = ~0u) {
    const int t = sh.t + sh.nlam;
#if defined(__CUDA_ARCH__)
    const uint32_t amb = (uint32_t)__cvta_generic_to_shared(sh.J);
    constexpr int PER = 16 / (int)sizeof(W);
    int q = 0;
#pragma unroll 4
    for (; q + PER <= t; q += PER) {
        uint32_t x0, x1, x2, x3;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(amb + (uint32_t)q * (uint32_t)sizeof(W)));
        if (sizeof(W) == 8) {
            J.put(q, (W)(((uint64_t)x1 << 32) | (x0 & keep))); J.put(q + 1, (W)(((uint64_t)x3 << 32) | (x2 & keep)));
        } else {
            J.put(q, (W)x0); J.put(q + 1, (W)x1); J.put(q + 2, (W)x2); J.put(q + 3, (W)x3);
        }
    }
    for (; q < t; q++) J.put(q, sizeof(W) == 8 ? (W)(sh.J[q] & ((W)~(W)0xffffffffu | keep)) : sh.J[q]);
#else
    for (int q = 0; q < t; q++) J.put(q, sizeof(W) == 8 ? (W)(sh.J[q] & ((W)~(W)0xffffffffu | keep)) : sh.J[q]);
#endif
}

// <phi|theta> for a |L> term (prepL): |+> on supp(xt), |0> elsewhere.
template <typename W, bool MANYC = false>
BG_HD void t_term_L(const Rows<W>& J, const TShared<W>& sh, W xt, int& eps, int& p, int& m) {
    const int t = sh.t;
    TF<W> f;
    f.D1 = sh.D1; f.D2 = sh.D2; f.Q = sh.Q;
    f.A = xt & tlowmask<W>(t);
    const int k2 = tpopc(f.A);
    // low columns of the variables that are not in the term are cleared: free slots for t_oddblock64
    const uint32_t fr = (sizeof(W) == 8 && !MANYC && sh.ncons == 0) ? ~(uint32_t)f.A : 0u;
    t_copy_in<W>(J, sh, ~fr);
    if (!(MANYC ? t_constraints_many<W>(J, f, sh, (W)0) : t_constraints<W>(J, f, sh, (W)0))) { eps = 0; p = 0; m = 0; return; }
    f.A |= tlowmask<W>(sh.nlam) << (sh.nlam ? t : 0);            // the Lagrange variables of the parity checks
#if defined(BG_ELIM_FOLD)
    t_expsum<W>(J, f, eps, p, m);
#else
    t_expsum_odd(J, f, eps, p, m, fr);
#endif
    if (eps) p -= sh.k1 + k2 + 2 * sh.nlam; else { p = 0; m = 0; }
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
    f.A = (maskt & ~last) | (tlowmask<W>(sh.nlam) << (sh.nlam ? t : 0));     // + the Lagrange variables of the checks
    for (int q = t; q < t + sh.nlam; q++) J.put(q, sh.J[q]);
    for (W r = mrg; r; r &= r - 1) {                    // x_{2j} = x_{2j+1}: eliminate x_{2j}
        const int i = tlowest(r);
        t_basis_change<W>(J, f, i, tbit<W>(i + 1));
        f.A &= ~tbit<W>(i);
    }
    const int k2 = t - tpopc(mrg) - tpopc(last);
    if (!(MANYC ? t_constraints_many<W>(J, f, sh, mrg) : t_constraints<W>(J, f, sh, mrg))) { eps = 0; p = 0; m = 0; return; }
#if defined(BG_ELIM_FOLD)
    t_expsum<W>(J, f, eps, p, m);
#else
    t_expsum_odd(J, f, eps, p, m);
#endif
    if (eps) p -= sh.k1 + k2 + 2 * sh.nlam; else { p = 0; m = 0; }
}

}  // namespace bg
