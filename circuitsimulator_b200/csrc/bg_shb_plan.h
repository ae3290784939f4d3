// bg_shb_plan.h — host-side plan for the shared high-block reduction of the L x chi loop (bg_shb.cuh).
//
// The chi = 2^k terms of a |L> decomposition are the code words x~ = z L, z in F_2^k (prepL,
// libcirc/stateprep.c:85-103).  When 32 < t <= 32 + SHB_MAXH the plan looks for NH = t - 32 columns of L that
// lie in a subspace W of F_2^k of dimension k - 5: restricted to those columns the code words show only
// 2^(k-5) patterns, so the terms split into classes of >= 32 terms with the SAME pattern on them.  The
// variables are relabelled so that these columns are the variables 32 .. t-1 ("high"), the terms are sorted
// by pattern, and k_pairs_shb eliminates the high variables ONCE per (sample, 32-term batch) with the warp
// cooperating; each thread is left with a form on the 32 low variables in 32-bit words.
//
// Plain C++ (no CUDA): included by bgnorm.cu (product) and by tests/emu/emu_lib.cpp (the CPU emulator build).
#pragma once
#include <stdint.h>
#include <algorithm>
#include <vector>

#define SHB_MAXH 12        // high variables taken from the term (t - 32)
#define SHB_MAXLAM 6       // parity checks of theta carried along as Lagrange variables (more: generic kernel)
#define SHB_MAXHT (SHB_MAXH + SHB_MAXLAM)
#ifndef SHB_RELOC
#define SHB_RELOC 4        // leftover high variables a thread can relocate to free low slots (host-checked)
#endif

struct ShbPlan {
    int ok = 0;
    int t = 0, nh = 0;
    uint8_t perm[64];          // new position of variable v
    uint8_t iperm[64];         // variable at new position
    int nsw = 0;
    uint8_t swp[SHB_MAXH], swq[SHB_MAXH];   // the same permutation as bit swaps (p < 32 <= q)
    std::vector<uint64_t> terms;   // relabelled terms, sorted by high pattern (classes are multiples of 32)
    std::vector<int32_t> nat;      // natural index of sorted position
};

static inline uint64_t shb_permute_bits(uint64_t w, const ShbPlan& pl) {
    for (int i = 0; i < pl.nsw; i++) {
        const uint64_t x = ((w >> pl.swp[i]) ^ (w >> pl.swq[i])) & 1ull;
        w ^= (x << pl.swp[i]) | (x << pl.swq[i]);
    }
    return w;
}

// rank of a set of k-bit vectors / membership in their span, by an echelon basis indexed by leading bit
struct ShbBasis {
    uint32_t b[32];
    ShbBasis() { for (int i = 0; i < 32; i++) b[i] = 0; }
    uint32_t reduce(uint32_t v) const {
        while (v) { const int h = 31 - __builtin_clz(v); if (!b[h]) return v; v ^= b[h]; }
        return 0;
    }
    bool add(uint32_t v) { v = reduce(v); if (!v) return false; b[31 - __builtin_clz(v)] = v; return true; }
    // the representative of v + span with a zero at every pivot position (the same for all members of a coset)
    uint32_t residue(uint32_t v) const {
        for (int h = 31; h >= 0; h--) if (b[h] && ((v >> h) & 1u)) v ^= b[h];
        return v;
    }
};

// terms_nat[i] = x~_i in natural order (bit q = variable q), L = the k rows.  Returns plan.ok = 1 on success.
static inline ShbPlan shb_make_plan(int t, int k, const std::vector<uint64_t>& L, const std::vector<uint64_t>& terms_nat) {
    ShbPlan pl;
    pl.t = t;
    const int nh = t - 32, r = k - 5;
    if (nh < 1 || nh > SHB_MAXH || k < 6 || k > 26 || (int)L.size() != k) return pl;
    // column c of L as a k-bit vector
    std::vector<uint32_t> col(t, 0);
    for (int c = 0; c < t; c++) for (int j = 0; j < k; j++) col[c] |= (uint32_t)((L[j] >> c) & 1ull) << j;
    // search: a subspace W of dimension <= r containing >= nh columns.  Candidates are spans of r columns
    // (combinations in lexicographic order, bounded effort); zero columns are in every W.
    std::vector<int> best;
    {
        std::vector<int> idx(std::max(r, 1));
        long long budget = 1000;                         // lexicographic fallback after the randomised search (~2 ms in all when no plan exists)
        auto count_inside = [&](const ShbBasis& B, std::vector<int>& inside) {
            inside.clear();
            for (int c = 0; c < t; c++) if (B.reduce(col[c]) == 0) inside.push_back(c);
        };
        if (r <= 0) {
            ShbBasis B; std::vector<int> in; count_inside(B, in);
            if ((int)in.size() >= nh) best = in;
        } else if (r >= nh) {
            for (int c = t - nh; c < t; c++) best.push_back(c);       // any nh columns span at most nh <= r dimensions
        } else {
            // First a randomised search (deterministic: seeded by L).  r - 1 random columns span U; every other column
            // lies in U or in one of the cosets U + c, and a coset that holds m columns gives the subspace U + <c> with
            // |inside U| + m columns — all choices of the r-th generator are judged at once.  For a random 9 x 40 L a
            // hit takes ~40 tries of ~1 us where the lexicographic enumeration below needs thousands of dependent ones.
            uint64_t seed = 0x9E3779B97F4A7C15ull;
            for (int j = 0; j < k; j++) seed = (seed ^ L[j]) * 0xD1342543DE82EF95ull + 1ull;
            auto rnd = [&]() { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; return seed; };
            std::vector<uint32_t> res(t);
            for (int tries = 0; tries < 1500 && best.empty(); tries++) {
                ShbBasis B;
                for (int i = 0; i < r - 1; i++) B.add(col[(int)(rnd() % (uint64_t)t)]);
                int inside_u = 0, nres = 0;
                for (int c = 0; c < t; c++) { const uint32_t v = B.residue(col[c]); if (v) res[nres++] = v; else inside_u++; }
                std::sort(res.begin(), res.begin() + nres);
                int run = 0, best_run = 0; uint32_t best_v = 0;
                for (int i = 0; i < nres; i++) {
                    run = (i > 0 && res[i] == res[i - 1]) ? run + 1 : 1;
                    if (run > best_run) { best_run = run; best_v = res[i]; }
                }
                if (inside_u + best_run < nh) continue;
                if (best_v) B.add(best_v);
                std::vector<int> in;
                count_inside(B, in);
                if ((int)in.size() >= nh) best = in;
            }
            for (int i = 0; i < r; i++) idx[i] = i;
            std::vector<int> inside;
            while ((int)best.size() < nh && budget-- > 0) {
                ShbBasis B;
                for (int i = 0; i < r; i++) B.add(col[idx[i]]);
                count_inside(B, inside);
                if (inside.size() > best.size()) { best = inside; if ((int)best.size() >= nh) break; }     // the first hit will do
                int i = r - 1;
                while (i >= 0 && idx[i] == t - r + i) i--;
                if (i < 0) break;
                idx[i]++;
                for (int j = i + 1; j < r; j++) idx[j] = idx[j - 1] + 1;
            }
        }
    }
    if ((int)best.size() < nh) return pl;
    // prefer columns that already sit at positions >= 32 (fewer bit swaps), then any
    std::vector<int> H;
    for (int c : best) if (c >= 32 && (int)H.size() < nh) H.push_back(c);
    for (int c : best) if (c < 32 && (int)H.size() < nh) H.push_back(c);
    // class size must be a multiple of 32: rank of the chosen columns <= k - 5
    { ShbBasis B; int rk = 0; for (int c : H) rk += B.add(col[c]) ? 1 : 0; if (rk > r) return pl; }
    for (int v = 0; v < 64; v++) pl.perm[v] = pl.iperm[v] = (uint8_t)v;
    std::vector<int> inH(64, 0);
    for (int c : H) inH[c] = 1;
    std::vector<int> lowH, highFree;
    for (int c = 0; c < 32; c++) if (inH[c]) lowH.push_back(c);
    for (int c = 32; c < t; c++) if (!inH[c]) highFree.push_back(c);
    if (lowH.size() != highFree.size()) return pl;
    pl.nsw = (int)lowH.size();
    for (int i = 0; i < pl.nsw; i++) {
        pl.swp[i] = (uint8_t)lowH[i]; pl.swq[i] = (uint8_t)highFree[i];
        pl.perm[lowH[i]] = (uint8_t)highFree[i]; pl.perm[highFree[i]] = (uint8_t)lowH[i];
        pl.iperm[lowH[i]] = (uint8_t)highFree[i]; pl.iperm[highFree[i]] = (uint8_t)lowH[i];
    }
    pl.nh = nh;
    const size_t chi = terms_nat.size();
    std::vector<uint64_t> tp(chi);
    for (size_t i = 0; i < chi; i++) {
        tp[i] = shb_permute_bits(terms_nat[i], pl);
        if (32 - __builtin_popcount((uint32_t)tp[i]) < SHB_RELOC) return pl;       // every thread needs SHB_RELOC free low slots
    }
    pl.nat.resize(chi);
    // by pattern; inside a class by popcount of the low word, largest first (equal work for neighbouring lanes); ties
    // keep the natural order.  One sort of composite keys (pattern | 32 - popcount | index; chi <= 2^26).
    {
        std::vector<uint64_t> keys(chi);
        for (size_t i = 0; i < chi; i++)
            keys[i] = ((uint64_t)(uint32_t)(tp[i] >> 32) << 32) | ((uint64_t)(32 - __builtin_popcount((uint32_t)tp[i])) << 26) | (uint64_t)i;
        std::sort(keys.begin(), keys.end());
        for (size_t i = 0; i < chi; i++) pl.nat[i] = (int32_t)(keys[i] & 0x3ffffffull);
    }
    pl.terms.resize(chi);
    for (size_t i = 0; i < chi; i++) pl.terms[i] = tp[pl.nat[i]];
    for (size_t i = 0; i + 31 < chi; i += 32)
        if ((pl.terms[i] >> 32) != (pl.terms[i + 31] >> 32)) return pl;             // a class that is not a multiple of 32
    if (chi % 32) return pl;
    pl.ok = 1;
    return pl;
}
