// bg_philox.cuh — Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw, SC'11)
// and the random stabilizer state sampler built on it.
//
// Replaces randomStabilizerState (libcirc/stabilizer/stabilizer.c:689-756), whose randomness
// is libc rand(): here sample l of bin b under seed s is a pure function of (s, b, l), so the
// set of states is identical for any number of GPUs / any launch geometry.
//
// Stream layout (key = seed, counter = (l lo, l hi, bin, block)); mirrored on the CPU by
// the test oracle (restated there for the bit-exact RNG check).
//   block 0          : words 0,1 -> u = ((w1:w0 >> 11) + 1) 2^-53 in (0,1] -> d (eq. 79 cdf), k = n - d
//   block 1 + j/2    : half j%2  -> xi_j, the j-th random hyperplane of the lazy shrink
//   block 0x1000     : half 0 -> h,  half 1 -> D1
//   block 0x1001     : half 0 -> D2
//   block 0x2000 + v : half 0 -> r_v ; J_uv = bit v of r_u for v < u (both active), J_vv = D1_v
#pragma once
#include "bg_device.cuh"

namespace bg {

struct Philox4 { uint32_t w[4]; };

BG_HD uint32_t bg_mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

BG_HD Philox4 philox4x32_10(uint64_t seed, uint64_t sample, uint32_t bin, uint32_t block) {
    uint32_t c0 = (uint32_t)sample, c1 = (uint32_t)(sample >> 32), c2 = bin, c3 = block;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t h0 = bg_mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = bg_mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o; o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}
BG_HD uint64_t philox_half(const Philox4& p, int half) { return ((uint64_t)p.w[2 * half + 1] << 32) | p.w[2 * half]; }

// theta ~ the law of randomStabilizerState(n).  cdf[0..n] is the cumulative distribution of
// d = n - k exactly as stabilizer.c:693-716 computes it (host-side table).
template <int NS>
BG_DEV void native_random(Native<NS>& st, int n, uint64_t seed, uint32_t bin, uint64_t sample, const double* cdf) {
    typedef typename WordOf<NS>::T W;
    const int lane = bg_lane();
    const W maskn = lowmaskw<W>(n);
    Philox4 b0 = philox4x32_10(seed, sample, bin, 0);
    const double u = (double)((philox_half(b0, 0) >> 11) + 1ull) * (1.0 / 9007199254740992.0);
    int d = 0;
    while (d < n && !(u <= cdf[d])) d++;
    const int k = n - d;
    native_identity<NS>(st, n);
    for (uint32_t j = 0; popcw(st.f.A) > k && j < 100000u; j++) {
        Philox4 b = philox4x32_10(seed, sample, bin, 1u + j / 2u);
        native_shrink<NS>(st, (W)philox_half(b, (int)(j & 1u)) & maskn, 0u, true);
    }
    const W A = st.f.A;
    Philox4 b1 = philox4x32_10(seed, sample, bin, 0x1000u);
    Philox4 b2 = philox4x32_10(seed, sample, bin, 0x1001u);
    st.h = (W)philox_half(b1, 0) & maskn;
    st.f.D1 = (W)philox_half(b1, 1) & A;
    st.f.D2 = (W)philox_half(b2, 0) & A;
    st.f.Q = 0;
    W r[NS], up[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        Philox4 br = philox4x32_10(seed, sample, bin, 0x2000u + (uint32_t)v);
        r[s] = ((A >> v) & 1) ? ((W)philox_half(br, 0) & lowmaskw<W>(v) & A) : 0;     // strictly lower part
        up[s] = 0;
    }
#pragma unroll
    for (int s = 0; s < NS; s++) up[s] = r[s];
    transposew(up);                               // the lower triangle, mirrored into the upper one
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int v = lane + 32 * s;
        st.f.J[s] = ((A >> v) & 1) ? (r[s] | up[s] | (st.f.D1 & bitw<W>(v))) : 0;
    }
}

}  // namespace bg
