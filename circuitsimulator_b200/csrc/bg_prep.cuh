// bg_prep.cuh — drawing and projecting a sample theta, one THREAD per sample.
//
// k_prepare (bgnorm.cu) gives a warp to every sample: rows in lanes, ballots and shuffles, every
// warp-uniform scalar replicated 32 times — 5170 warp-instructions per sample at t = 40, ~135 per generator
// of the projector (profiles/r2_k_prepare_ncu_summary.json).  That is the mapping bg_tpp.cuh moved the pair
// loop away from, and the same argument holds here: in ambient coordinates a generator is a handful of mask
// operations plus ONE symmetric rank-two update of J, J_v ^= [v in X] Y ^ [v in Z] U.  A thread that owns
// its sample — J and the parity checks as rows of shared memory in the [row][thread] layout of bg_tpp.cuh,
// everything else in registers — spends ~600 instructions per generator, so a warp does 32 samples with the
// instruction count the warp-per-sample kernel needs for four.
//
// What must be reproduced (and is, bit for bit: tests/test_device_code_on_emulator.py compares the records
// of both formulations): randomStabilizerState (stabilizer.c:689-756) on the Philox streams of
// bg_philox.cuh, the change to ambient coordinates (bg_device.cuh: ambient(), identity case), and
// measurePauli (stabilizer.c:827-959) for every generator of the projector (innerprod.c:100-116) in the
// ambient form of bg_device.cuh: ambient_measure.
//
// No warp collectives anywhere: lanes may diverge freely (the second projector's samples, dead samples).
#pragma once
#include "bg_tpp.cuh"
#include "bg_philox.cuh"

namespace bg {

template <typename W> struct TSample {     // what a thread keeps in registers; J and the checks are Rows
    W D1, D2, Cpend, Cbeta;
    uint32_t Q;
    int npf;
    bool alive;
};

template <typename W> BG_HD uint32_t tpar(W x) { return (uint32_t)tpopc(x) & 1u; }

// r ^= V where m != 0: one predicate, predicated xors (no select, no mask word)
BG_HD void t_cxor(uint32_t& r, uint32_t m, uint32_t V) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p xor.b32 %0, %0, %2;\n\t}" : "+r"(r) : "r"(m), "r"(V));
#else
    if (m) r ^= V;
#endif
}
BG_HD void t_cxor(uint64_t& r, uint32_t m, uint64_t V) {
#if defined(__CUDA_ARCH__)
    uint32_t lo = (uint32_t)r, hi = (uint32_t)(r >> 32);
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p xor.b32 %0, %0, %3;\n\t@p xor.b32 %1, %1, %4;\n\t}"
        : "+r"(lo), "+r"(hi) : "r"(m), "r"((uint32_t)V), "r"((uint32_t)(V >> 32)));
    r = ((uint64_t)hi << 32) | lo;
#else
    if (m) r ^= V;
#endif
}

// byte `byte` of row r = val (the rest of the row untouched): one STS.U8
template <typename W> BG_HD void t_row_put_byte(const Rows<W>& R, int r, int byte, uint32_t val) {
#if defined(__CUDA_ARCH__)
    asm volatile("st.shared.u8 [%0], %1;" :: "r"((uint32_t)r * R.sstride + R.sbase + (uint32_t)byte), "r"(val) : "memory");
#else
    const W x = R.get(r);
    R.put(r, (x & ~((W)0xff << (8 * byte))) | ((W)(val & 0xffu) << (8 * byte)));
#endif
}

// 8 x 8 bit transpose: byte i, bit j  <->  byte j, bit i  (three masked swaps)
BG_HD uint64_t t_transpose8(uint64_t x) {
    uint64_t y;
    y = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull; x ^= y ^ (y << 7);
    y = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull; x ^= y ^ (y << 14);
    y = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull; x ^= y ^ (y << 28);
    return x;
}

// J_v ^= [v in X] Y ^ [v in Z] U for every row v < nr (nr a multiple of 8; X and Z have no bits >= n, the
// padding rows exist and stay zero).  One pass, the same trip count in every lane.
template <typename W> BG_HD void t_rank2(const Rows<W>& J, int nr, W X, W Y, W Z, W U) {
    for (int v0 = 0; v0 < nr; v0 += 8) {
        const uint32_t xb = (uint32_t)(X >> v0), zb = (uint32_t)(Z >> v0);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            W r = J.get(v0 + k);
            t_cxor(r, xb & (1u << k), Y);
            t_cxor(r, zb & (1u << k), U);
            J.put(v0 + k, r);
        }
    }
}

// ---- theta ~ randomStabilizerState(n), directly in ambient form.
// In the warp formulation (native_random + ambient()) the state is (G, Gbar, A): a lazy shrink by a random
// hyperplane xi removes one pivot i from A.  Two facts make G and Gbar unnecessary here:
//   * an active row G_a is the unique element of K whose active coordinates are e_a (rows only ever receive
//     rows of removed pivots), and the removed coordinates of x in K are fixed by the checks; with the checks
//     kept REDUCED (check i = e_i + active coordinates)  G_a . xi = xi_a + sum_{i removed, a in C_i} xi_i;
//   * the new check is S itself (Gbar_i + sum_{a in S, a != i} Gbar_a with Gbar_a = e_a for active a).
// The checks then go to the reduced echelon form by slot that project_ambient starts from (checks_echelon:
// pivot = lowest bit; that form is unique for the space, so the basis it is computed from is immaterial).
// Scratch: the rows of J serve as temporary rows of the echelon form before J is drawn.
template <typename W>
BG_HD void t_random_ambient(const Rows<W>& J, const Rows<W>& C, int n, uint64_t seed, uint32_t bin, uint64_t sample,
                            const double* cdf, TSample<W>& o) {
    const W maskn = tlowmask<W>(n);
    const Philox4 b0 = philox4x32_10(seed, sample, bin, 0);
    const double u = (double)((philox_half(b0, 0) >> 11) + 1ull) * (1.0 / 9007199254740992.0);
    int d = 0;
    while (d < n && !(u <= cdf[d])) d++;
    const int k = n - d;
    W A = maskn;
    for (uint32_t j = 0; tpopc(A) > k && j < 100000u; j++) {
        const Philox4 b = philox4x32_10(seed, sample, bin, 1u + j / 2u);
        const W xi = (W)((j & 1u) ? philox_half(b, 1) : philox_half(b, 0)) & maskn;
        W S = xi;
        for (W rem = xi & ~A; rem; rem &= rem - 1) S ^= C.get(tlowest(rem));
        S &= A;
        if (S == 0) continue;                                   // xi vanishes on K: SAME
        const int i = tlowest(S);
        for (W rem = maskn & ~A; rem; rem &= rem - 1) {         // keep the earlier checks free of coordinate i
            const int r = tlowest(rem);
            const W c = C.get(r);
            if ((c >> i) & 1) C.put(r, c ^ S);
        }
        C.put(i, S);
        A &= ~tbit<W>(i);
    }
    const Philox4 b1 = philox4x32_10(seed, sample, bin, 0x1000u);
    const Philox4 b2 = philox4x32_10(seed, sample, bin, 0x1001u);
    const W h = (W)philox_half(b1, 0) & maskn;
    W D1 = (W)philox_half(b1, 1) & A;
    W D2 = (W)philox_half(b2, 0) & A;

    // checks -> reduced echelon form by slot (pivot = lowest bit), right-hand sides beta = c . h
    W pn = 0, bn = 0;
    for (W rem = maskn & ~A; rem; rem &= rem - 1) {
        W w = C.get(tlowest(rem));
        uint32_t beta = tpar<W>(w & h);
        for (W f = w & pn; f; f &= f - 1) {
            const int q = tlowest(f);
            w ^= J.get(q);
            beta ^= tget<W>(bn, q);
        }
        if (w == 0) continue;
        const int p = tlowest(w);
        for (W f = pn; f; f &= f - 1) {
            const int q = tlowest(f);
            const W c = J.get(q);
            if ((c >> p) & 1) { J.put(q, c ^ w); bn ^= (W)beta << q; }
        }
        J.put(p, w);
        pn |= tbit<W>(p);
        bn = (bn & ~tbit<W>(p)) | ((W)beta << p);
    }
    for (W f = pn; f; f &= f - 1) { const int q = tlowest(f); C.put(q, J.get(q)); }

    // J: row v = the strictly lower part drawn from block 0x2000 + v, mirrored into the rows above, diagonal D1.
    // Eight rows at a time (eight independent Philox chains in flight); their mirror image — bits v0 .. v0+7 of
    // the rows above — is one byte per row: 8 x 8 bit transposes and byte stores, the same work in every lane.
    const int nr = (n + 7) & ~7;
    for (int v0 = 0; v0 < nr; v0 += 8) {
        W r[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int v = v0 + k;
            const Philox4 br = philox4x32_10(seed, sample, bin, 0x2000u + (uint32_t)v);
            r[k] = (v < n && ((A >> v) & 1)) ? (((W)philox_half(br, 0) & tlowmask<W>(v) & A) | (D1 & tbit<W>(v))) : (W)0;
        }
        for (int c0 = 0; c0 < v0; c0 += 8) {                     // columns c0 .. c0+7 of the eight rows
            uint64_t x = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) x |= (uint64_t)((uint32_t)(r[k] >> c0) & 0xffu) << (8 * k);
            x = t_transpose8(x);                                  // byte j: bits v0 .. v0+7 of row c0 + j
#pragma unroll
            for (int jj = 0; jj < 8; jj++) t_row_put_byte<W>(J, c0 + jj, v0 >> 3, (uint32_t)(x >> (8 * jj)) & 0xffu);
        }
        uint64_t x = 0;                                           // the diagonal block: lower triangle + diagonal
#pragma unroll
        for (int k = 0; k < 8; k++) x |= (uint64_t)((uint32_t)(r[k] >> v0) & 0xffu) << (8 * k);
        x |= t_transpose8(x);
#pragma unroll
        for (int k = 0; k < 8; k++)
            J.put(v0 + k, (r[k] & tlowmask<W>(v0)) | ((W)((uint32_t)(x >> (8 * k)) & 0xffu) << v0));
    }
    // shift x -> x + h (updateQD, stabilizer.c:163-177, as ambient() does it)
    uint32_t tri = 0;
    W pd = 0;
    for (int v = 0; v < n; v++) {
        const W row = J.get(v) & h;
        tri ^= tget<W>(h, v) & tpar<W>(row & tlowmask<W>(v));
        pd |= (W)tpar<W>(row & ~tbit<W>(v)) << v;
    }
    o.Q = (2u * (uint32_t)tpopc(D1 & h) + 4u * (uint32_t)tpopc(D2 & h) + 4u * tri) & 7u;
    D2 ^= (h & D1) ^ (pd & maskn);
    o.D1 = D1; o.D2 = D2; o.Cpend = pn; o.Cbeta = bn;
    o.npf = 0; o.alive = true;
}

// f(q, bit q as a word) for every set bit of m, from the top: 64-bit masks are walked as two 32-bit words (one FLO
// and one xor per bit; the 64-bit "lowest bit, clear it" idiom costs a dozen instructions)
template <typename F> BG_HD void t_each_bit(uint32_t m, F f) {
    while (m) { const int c = thighest(m); const uint32_t b = 1u << c; m ^= b; f(c, b); }
}
template <typename F> BG_HD void t_each_bit(uint64_t m, F f) {
    uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
    while (hi) { const int c = thighest(hi); const uint32_t b = 1u << c; hi ^= b; f(c + 32, (uint64_t)b << 32); }
    while (lo) { const int c = thighest(lo); const uint32_t b = 1u << c; lo ^= b; f(c, (uint64_t)b); }
}

// ---- one generator i^m Z(zeta) X(xi) of the projector on a sample in ambient form: ambient_measure
// (bg_device.cuh) per thread, same case analysis, same choices (the dropped check is the one in the lowest
// hit slot, a new check goes to the slot of its lowest bit).  Returns 0 (annihilated), 1, or 2 (factor 2^-1/2).
template <typename W>
BG_HD int t_ambient_measure(const Rows<W>& J, const Rows<W>& C, int n, int nr, TSample<W>& s, uint32_t m, W zeta, W xi) {
    const W maskn = tlowmask<W>(n);
    // eta = zeta + J xi (J symmetric: the xor of the rows in xi); tri = sum_{q<r in xi} J_qr = the parity of
    // sum_q |J_q & (the bits of xi visited before q)|, every unordered pair once
    W eta = zeta, acc = 0, seen = 0;
    t_each_bit(xi, [&](int q, W b) {
        const W row = J.get(q);
        eta ^= row;
        acc ^= row & seen;
        seen |= b;
    });
    const uint32_t tri = tpar<W>(acc);
    eta &= maskn;
    W hit = 0;
    t_each_bit(s.Cpend, [&](int j, W b) { hit |= b & tfill<W>(tpar<W>(C.get(j) & xi)); });
    const uint32_t w0 = (2u * m + 2u * (uint32_t)tpopc(s.D1 & xi) + 4u * (uint32_t)tpopc(s.D2 & xi) + 4u * tri) & 7u;
    W X = 0, Y = 0, Z = 0, U = 0;
    int ret = 2;
    if (hit) {
        // xi leaves K: K' = K u (K + xi); the check c0 in the lowest hit slot is dropped
        const int p0 = tlowest(hit);
        const W bp0 = tbit<W>(p0);
        const W c0 = C.get(p0);
        const bool b0 = (s.Cbeta & bp0) != 0;
        t_each_bit(hit ^ bp0, [&](int v, W) { C.put(v, C.get(v) ^ c0); });
        if (b0) s.Cbeta ^= hit;
        s.Cbeta &= ~bp0; s.Cpend &= ~bp0;
        s.D2 ^= (c0 & eta) ^ (b0 ? eta : (W)0);
        if (b0) s.Q = (s.Q + w0) & 7u;
        const uint32_t wp = b0 ? ((8u - w0) & 7u) : w0;
        X = c0; Y = eta; Z = eta; U = c0;
        if (wp == 4u) s.D2 ^= c0;
        else if (wp == 2u || wp == 6u) {
            s.D2 ^= (wp == 2u ? s.D1 : ~s.D1) & c0;
            s.D1 ^= c0;
            Y ^= c0;
        }
    } else if (w0 == 0u || w0 == 4u) {
        // projector onto eta.y = w0/4 inside K
        W er = eta, brm = 0;
        t_each_bit(eta & s.Cpend, [&](int q, W b) { er ^= C.get(q); brm ^= b; });
        const uint32_t br = (w0 >> 2) ^ tpar<W>(brm & s.Cbeta);
        if (er == 0) ret = br ? 0 : 1;
        else {
            const int p = tlowest(er);
            const W bp = tbit<W>(p), brw = tfill<W>(br);
            t_each_bit(s.Cpend, [&](int q, W b) {
                const W c = C.get(q);
                if (c & bp) { C.put(q, c ^ er); s.Cbeta ^= b & brw; }
            });
            C.put(p, er);
            s.Cpend |= bp;
            s.Cbeta = (s.Cbeta & ~bp) | (bp & brw);
        }
    } else {
        // w0 in {2,6}: 1 + w^{w(y)} = sqrt2 w^{+-1}
        if (w0 == 2u) { s.Q = (s.Q + 1u) & 7u; s.D2 ^= eta & ~s.D1; }
        else { s.Q = (s.Q + 7u) & 7u; s.D2 ^= eta & s.D1; }
        s.D1 ^= eta;
        X = eta; Y = eta;
    }
    // No branch above returns: the three cases reconverge HERE, so that the lanes whose case updates J (the first
    // and the third) walk the rows together.  (With an early return in the middle case the reconvergence point of
    // the case analysis is the end of the function and each case runs the row pass on its own: 7 of 32 lanes.)
    if ((X | Z) != 0) t_rank2<W>(J, nr, X, Y, Z, U);
    return ret;
}

// The generators in order (innerprod.c:100-116); s.npf counts the 2^-1/2 factors.
template <typename W>
BG_HD void t_project(const Rows<W>& J, const Rows<W>& C, int n, TSample<W>& s, const bg_projector* P) {
    const int nr = (n + 7) & ~7;
    const int ns = P->nstabs;
    for (int i = 0; i < ns; i++) {
        const int r = t_ambient_measure<W>(J, C, n, nr, s, (uint32_t)P->phase[i], (W)P->zs[i], (W)P->xs[i]);
        if (r == 0) { s.alive = false; return; }
        if (r == 2) s.npf++;
    }
}

}  // namespace bg
