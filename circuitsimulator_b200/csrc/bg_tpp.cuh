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
                    uint32_t& neg1, uint32_t& z0, uint32_t& z1, bool has_s, const TPend<uint64_t>& pd64) {
    uint32_t El = (uint32_t)E, Eh = (uint32_t)(E >> 32);
    uint32_t D2l = (uint32_t)D2, D2h = (uint32_t)(D2 >> 32), Jsl = (uint32_t)Js, Jsh = (uint32_t)(Js >> 32);
    const uint32_t ns = has_s ? 0u : 1u;
    TPend64 pd;
    pd.M1l = (uint32_t)pd64.M1; pd.M1h = (uint32_t)(pd64.M1 >> 32); pd.V1l = (uint32_t)pd64.V1; pd.V1h = (uint32_t)(pd64.V1 >> 32);
    pd.M2l = (uint32_t)pd64.M2; pd.M2h = (uint32_t)(pd64.M2 >> 32); pd.V2l = (uint32_t)pd64.V2; pd.V2h = (uint32_t)(pd64.V2 >> 32);
    // Variables are eliminated from the top, so the high halves die first: once no lane of the warp has a
    // variable >= 32 left, the rounds continue on the low halves of the same rows with 32-bit code.
    bool pending = true;
#if defined(__CUDA_ARCH__)
#define T_ALL_LOW() __all_sync(__activemask(), Eh == 0u)
#else
#define T_ALL_LOW() (Eh == 0u)
#endif
    if ((El | Eh) != 0u && !T_ALL_LOW()) {
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

// the thread's working copy of the ambient J.  Device: the warp's ambient rows are 16-byte aligned
// (k_pairs_tpp pads them), so they are read with 128-bit broadcast loads.
template <typename W>
BG_HD void t_copy_in(const Rows<W>& J, const TShared<W>& sh) {
    const int t = sh.t;
#if defined(__CUDA_ARCH__)
    const uint32_t amb = (uint32_t)__cvta_generic_to_shared(sh.J);
    constexpr int PER = 16 / (int)sizeof(W);
    int q = 0;
#pragma unroll 4
    for (; q + PER <= t; q += PER) {
        uint32_t x0, x1, x2, x3;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(amb + (uint32_t)q * (uint32_t)sizeof(W)));
        if (sizeof(W) == 8) {
            J.put(q, (W)(((uint64_t)x1 << 32) | x0)); J.put(q + 1, (W)(((uint64_t)x3 << 32) | x2));
        } else {
            J.put(q, (W)x0); J.put(q + 1, (W)x1); J.put(q + 2, (W)x2); J.put(q + 3, (W)x3);
        }
    }
    for (; q < t; q++) J.put(q, sh.J[q]);
#else
    for (int q = 0; q < t; q++) J.put(q, sh.J[q]);
#endif
}

// <phi|theta> for a |L> term (prepL): |+> on supp(xt), |0> elsewhere.
template <typename W, bool MANYC = false>
BG_HD void t_term_L(const Rows<W>& J, const TShared<W>& sh, W xt, int& eps, int& p, int& m) {
    const int t = sh.t;
    t_copy_in<W>(J, sh);
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
// Measured (B200, round 1): against the eager kernel as first committed 5.32 vs 5.78 ms at t = 60 but
// 5.92 vs 4.67 ms at t = 40; against the blocked eager rounds above it loses everywhere (t = 60, k = 12:
// 149 vs 109 ms) — the blocked passes remove most of the divergence it was meant to avoid.  Kept as an
// option (BG_LAZY=1) and as a cross-check of the eager path in the test-suite, not the default.
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
