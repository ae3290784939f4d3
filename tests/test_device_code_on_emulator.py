"""The product's warp-level device source (circuitsimulator_b200/csrc/bg_device.cuh, bg_philox.cuh,
bg_warp_ops.cuh) compiled for the 32-lane CPU warp emulator (tests/emu) and checked against the
reference's vectors and the oracle.  This is a check of the SOURCE where no GPU exists; the GPU
parity tests proper are in test_gpu_parity.py."""
import os

import numpy as np
import pytest

from util import load, states, epm_equal, parse_stream, GOLDEN
from oracle.oracle import states_to_numpy


@pytest.fixture(scope="module")
def emu():
    from emu.emu import Emu
    return Emu()


def test_inner_products_vs_compiled_reference_fixture(emu):
    d = load("ref_pairs.npz")
    a, b = states(d["a"]), states(d["b"])
    bad = [i for i, (s1, s2, w) in enumerate(zip(a, b, d["epm"])) if not epm_equal(emu.inner_product(s1, s2), tuple(w))]
    assert not bad, bad[:10]


def test_reference_inner_product_kat(emu):
    d = load("kat_inner_product.npz")
    for s1, s2, w in zip(states(d["a"]), states(d["b"]), d["epm"]):
        assert epm_equal(emu.inner_product(s1, s2), tuple(w))


def test_exponential_sum_kats(emu, oracle):
    d = load("kat_exponential_sum.npz")
    for s, w in list(zip(states(d["states"]), d["epm"]))[::3]:
        plus = oracle.identity_state(s.n, s.n)
        assert epm_equal(emu.inner_product(s, plus), (int(w[0]), int(w[1]) - 2 * s.n, int(w[2])))


def test_measure_pauli_kats(emu, oracle):
    d = load("kat_measure_pauli.npz")
    for s, want, m, z, x, res in zip(states(d["states_in"]), states(d["states_out"]), d["m"], d["zeta"], d["xi"], d["result"]):
        code, got = emu.measure_pauli(s, int(m), int(z), int(x))
        val = {0: 0.0, 1: 1.0, 2: 2 ** -0.5}[code]
        assert abs(val - res) < 1e-4
        assert got.k == want.k
        assert epm_equal(oracle.inner_product(got, want), (1, 0, 0))
        for j in range(3):
            pr = oracle.random_state_philox(got.n, 5, 0, j)
            assert epm_equal(oracle.inner_product(got, pr), oracle.inner_product(want, pr))


def test_rng_matches_oracle_restatement(emu, oracle):
    for n in (1, 3, 16, 32, 33, 40, 64):
        cdf = oracle.dimension_cdf(n)
        for s in range(8):
            a = emu.random_state(n, 2024, 3, s, cdf)
            b = oracle.random_state_philox(n, 2024, 3, s)
            assert a.key(full=False) == b.key(full=False)


def _terms(cfg, L, oracle):
    t = cfg["t"]
    if cfg["exact"]:
        size = (t + 1) // 2
        return [sum(((i >> (size - 1 - j)) & 1) << (2 * j) for j in range(size)) for i in range(1 << size)]
    return [oracle.Lbits(i, L) for i in range(1 << len(L))]


@pytest.mark.parametrize("tpp", [False, True, "pivot"], ids=["warp_per_pair", "thread_per_pair", "thread_per_pair_pivoted_checks"])
@pytest.mark.parametrize("stream,k,ns", [("htstack_t4.txt", 0, 24), ("hs_t16_bit6.txt", 0, 3),
                                         ("hs_t40_k9_bit0.txt", 5, 3), ("phase_estimation_q0.txt", 4, 3),
                                         ("toffoli_q0.txt", 0, 2)])
def test_L_chi_loop_vs_oracle(emu, oracle, stream, k, ns, tpp):
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", stream))
    t, exact = cfg["t"], cfg["exact"]
    rs = np.random.RandomState(3)
    L = [] if exact else [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) for _ in range(k)]
    terms = _terms(cfg, L, oracle)
    for P in (G, H):
        for s in range(ns):
            th = oracle.random_state_philox(t, 7, 0, s)
            want = oracle.sample_from_theta(th, P, exact, L)
            got = emu.terms(th, P, 1, exact, t, terms, tpp=tpp)
            assert got["alive"] == want["alive"]
            if not want["alive"]:
                continue
            assert all(epm_equal(tuple(got["epm"][i]), tuple(want["epm"][i])) for i in range(len(terms)))
            sh, a, r2 = t // 2 + 1, got["zw"], 2 ** -0.5
            tot = complex(a[0] + (a[1] - a[3]) * r2, a[2] + (a[1] + a[3]) * r2) / 2 ** sh
            assert abs(tot - want["total"]) <= 1e-12 * max(1.0, abs(want["total"]))
            assert abs(2 ** (-got["npf"] / 2) - want["projfactor"]) < 1e-15


@pytest.mark.parametrize("name", sorted(f for f in os.listdir(GOLDEN) if f.startswith("ref_samples_")))
def test_L_chi_loop_vs_compiled_reference_fixture(emu, oracle, name):
    d = load(name)
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", str(d["stream"])))
    t, exact = cfg["t"], cfg["exact"]
    L = [int(x) for x in d["L"]]
    terms = _terms(dict(cfg, exact=exact), L, oracle)
    th = states(d["theta"])
    n = min(len(th), 4 if t > 20 else 40)
    for l in range(n):
        P = (G, H)[int(d["which"][l])]
        got = emu.terms(th[l], P, 1, exact, t, terms)
        assert got["alive"] == d["alive"][l]
        if d["alive"][l]:
            assert all(epm_equal(tuple(got["epm"][i]), tuple(d["epm"][l][i])) for i in range(len(terms)))


def test_thread_per_pair_with_many_parity_checks(emu, oracle):
    """Low-dimensional thetas (many parity checks): the pivot history moves from registers to the
    thread's shared-memory rows (t_constraints_many)."""
    rs = np.random.RandomState(3)
    from oracle.oracle import Projector
    for t in (12, 40):
        L = [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) for _ in range(4)]
        terms = [oracle.Lbits(i, L) for i in range(16)]
        empty = Projector.make(t, [], [], [])
        seen = set()
        for j in range(30):
            s = oracle.random_state_philox(t, 9, 0, j)
            for _ in range((j * 3) % (t + 1)):
                oracle.measure_pauli(s, 0, int(rs.randint(1, 2 ** 62)) & ((1 << t) - 1) or 1, 0)
            seen.add(t - s.k)
            got = emu.terms(s, empty, 0, False, t, terms, tpp=True)
            assert got["alive"] == 1
            for i in range(16):
                assert epm_equal(tuple(got["epm"][i]), oracle.inner_product(s, oracle.prepL(i, t, L)))
        assert max(seen) > 6


def test_projection_in_ambient_coordinates(emu, oracle):
    """ambient_measure (measurePauli on a state that is already in ambient form) against the oracle's
    measurePauli (stabilizer.c:827-959): the reference's own measurePauli KAT states and generators, and
    random states of every dimension (many parity checks: the echelon bookkeeping, implied / contradicted
    checks, annihilation) under random Hermitian Paulis — Z-type, X-type and mixed.  Compared through what
    the hot path consumes: alive, the number of 2^-1/2 factors, and every <phi_i|P theta> of a |L> table."""
    from oracle.oracle import Projector
    rs = np.random.RandomState(17)
    kat = load("kat_measure_pauli.npz")
    cases = []
    for st, m, zeta, xi in zip(states(kat["states_in"]), kat["m"], kat["zeta"], kat["xi"]):
        cases.append((st, [(int(m), int(zeta), int(xi))]))
    for t in (5, 12, 33, 40):
        mask = (1 << t) - 1
        for j in range(24):
            s = oracle.random_state_philox(t, 21, 0, j)
            for _ in range((j * 5) % (t + 1)):                       # lower the dimension: many parity checks
                oracle.measure_pauli(s, 0, int(rs.randint(1, 2 ** 62)) & mask or 1, 0)
            gens = []
            for g in range(1 + j % 4):
                kind = (j + g) % 3
                x = 0 if kind == 0 else (int(rs.randint(1, 2 ** 62)) & mask or 1)
                z = 0 if kind == 1 else (int(rs.randint(1, 2 ** 62)) & mask or 1)
                gens.append(((bin(x & z).count("1") + 2 * int(rs.randint(0, 2))) % 4, z, x))
            cases.append((s, gens))
    dead = factors = 0
    for st, gens in cases:
        t = st.n
        L = [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) for _ in range(3)]
        terms = [oracle.Lbits(i, L) for i in range(8)]
        P = Projector.make(t, [g[0] for g in gens], [g[2] for g in gens], [g[1] for g in gens])
        want = oracle.sample_from_theta(st.copy(), P, False, L)
        for tpp in (False, True):
            got = emu.terms(st, P, 1, False, t, terms, tpp=tpp)
            assert got["alive"] == want["alive"]
            if not want["alive"]:
                continue
            assert abs(2 ** (-got["npf"] / 2) - want["projfactor"]) < 1e-15
            assert all(epm_equal(tuple(got["epm"][i]), tuple(want["epm"][i])) for i in range(8))
        dead += not want["alive"]
        factors += want["alive"] and want["projfactor"] < 1
    assert dead > 3 and factors > 20                                 # all the branches are exercised


@pytest.mark.parametrize("stream,t_expected,k,ns", [("hs_t40_k9_bit0.txt", 40, 9, 3), ("hs_t40_k9_bit2.txt", 40, 10, 2),
                                                    ("phase_estimation_q0.txt", 33, 9, 3)])
@pytest.mark.parametrize("variant", ["libbgemu.so", "libbgemu_reloc1.so"], ids=["reloc4", "reloc1_pivoted_leftovers"])
def test_shared_high_block_vs_oracle(oracle, stream, t_expected, k, ns, variant):
    """k_pairs_shb's code (bg_shb.cuh: relabelling, warp-level reduction of the variables >= 32, 32-bit threads)
    against the oracle, pair by pair, on theta drawn by the device RNG's CPU restatement and projected.
    The reloc1 build gives a thread ONE free slot for leftover high variables, so that the (otherwise rare)
    path that pivots the remaining ones as parity checks runs in most batches."""
    from emu.emu import Emu
    emu = Emu(variant)
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", stream))
    t = cfg["t"]
    assert t == t_expected
    rs = np.random.RandomState(11)
    L = [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) for _ in range(k)]
    terms = [oracle.Lbits(i, L) for i in range(1 << k)]
    seen = 0
    hist = np.zeros(16, dtype=int)
    for P in (G, H):
        for s in range(ns):
            th = oracle.random_state_philox(t, 13, 0, s)
            want = oracle.sample_from_theta(th, P, False, L)
            got = emu.terms_shb(th, P, 1, t, L)
            assert got["alive"] != -1, "no shared high-block plan for this L"
            if got["alive"] == -2:
                continue
            assert got["alive"] == want["alive"]
            if not want["alive"]:
                continue
            seen += 1
            hist += np.array(got["nleft_hist"])
            assert all(epm_equal(tuple(got["epm"][i]), tuple(want["epm"][i])) for i in range(len(terms)))
            sh, a, r2 = t // 2 + 1, got["zw"], 2 ** -0.5
            tot = complex(a[0] + (a[1] - a[3]) * r2, a[2] + (a[1] + a[3]) * r2) / 2 ** sh
            assert abs(tot - want["total"]) <= 1e-12 * max(1.0, abs(want["total"]))
    assert seen >= ns


@pytest.mark.parametrize("variant", ["libbgemu.so", "libbgemu_any.so"])
@pytest.mark.parametrize("wordbits", [32, 64])
def test_exponential_sums_by_enumeration(variant, wordbits):
    """The thread-level exponential sums of k_pairs_tpp / k_pairs_shb (odd-first with no / few / all free slots,
    and the fold + dimer rounds) against sum_x e^{i pi q(x)/4} enumerated over all x, exact integer compare, on
    random forms with <= 13 variables scattered over the word (zeros, all-even forms and out-of-slots hand-overs
    included).  libbgemu_any.so takes every warp-vote branch the way a lane does when OTHER lanes ask for it."""
    from emu.emu import Emu
    emu = Emu(variant)
    assert emu.lib.emu_expsum_selftest(wordbits, 7, 20000) == 0


def _random_projector(rs, t, ns, kind):
    """kind 0: arbitrary Paulis (light and dense X parts); kind 1: commuting Z-type generators with real phases, some
    repeated or negated — these add parity checks (measurePauli's shrink branch), leave the state alone, or kill it."""
    from oracle.oracle import Projector
    mask = (1 << t) - 1
    ph, xs, zs = [], [], []
    for _ in range(ns):
        if kind == 1:
            if zs and rs.randint(0, 6) == 0:
                j = int(rs.randint(0, len(zs)))
                xs.append(0); zs.append(zs[j]); ph.append(ph[j] if rs.randint(0, 8) else ph[j] ^ 2)
            else:
                z = int(rs.randint(0, 2 ** 62)) & mask
                if rs.randint(0, 2):
                    z &= int(rs.randint(0, 2 ** 62)) & int(rs.randint(0, 2 ** 62))
                xs.append(0); zs.append(z); ph.append(2 * int(rs.randint(0, 2)))
            continue
        ph.append(int(rs.randint(0, 4)))
        if rs.randint(0, 4):
            x = 0
            for q in rs.choice(t, size=min(t, int(rs.randint(0, 4))), replace=False):
                x |= 1 << int(q)
        else:
            x = int(rs.randint(0, 2 ** 62)) & mask
        xs.append(x)
        zs.append(int(rs.randint(0, 2 ** 62)) & mask)
    return Projector.make(t, ph, xs, zs)


@pytest.mark.parametrize("stream,ns", [("htstack_t4.txt", 64), ("hs_t16_bit6.txt", 48), ("hs_t40_k9_bit0.txt", 48),
                                       ("phase_estimation_q0.txt", 48), ("toffoli_q0.txt", 32)])
def test_thread_per_sample_prepare_equals_warp_per_sample(emu, oracle, stream, ns):
    """bg_prep.cuh (k_prepare_tps: one thread per sample) against the warp-per-sample code it replaces on the
    hot path: every field of the sample record, for the projectors of the golden streams."""
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", stream))
    t = cfg["t"]
    cdf = oracle.dimension_cdf(t)
    dead = 0
    for pi, P in enumerate((G, H)):
        for s in range(ns):
            a, b = emu.prepare_both(t, 77 + pi, 1, s, cdf, P)
            assert a == b, (stream, pi, s, [i for i in range(136) if a[i] != b[i]][:8])
            dead += a[0] == 0
    assert dead < 2 * ns


@pytest.mark.parametrize("t", [1, 2, 5, 8, 17, 31, 32, 33, 40, 47, 63, 64])
def test_thread_per_sample_prepare_random_projectors(emu, oracle, t):
    """Random generators at every word-size boundary: arbitrary Paulis (the extend and the phase branch of
    measurePauli) and commuting Z-type sets with repeats (new parity checks up to dimension 0, unchanged states,
    annihilated samples)."""
    rs = np.random.RandomState(100 + t)
    cdf = oracle.dimension_cdf(t)
    seen_dead = seen_checks = 0
    for trial in range(16):
        kind = trial & 1
        P = _random_projector(rs, t, int(rs.randint(0, (t if kind else 2 * t) + 1)), kind)
        for s in range(5):
            a, b = emu.prepare_both(t, 5 + trial, 2, 1000 * trial + s, cdf, P)
            assert a == b, (t, trial, s, [i for i in range(136) if a[i] != b[i]][:8])
            seen_dead += a[0] == 0
            seen_checks += bin(a[6]).count("1") > 3
    if t >= 8:
        assert seen_checks > 0 and seen_dead > 0


@pytest.mark.parametrize("t,k,min_rate", [(33, 6, 0.99), (36, 8, 0.95), (40, 9, 0.95), (40, 12, 0.99), (44, 9, 0.0), (44, 13, 0.9)])
def test_shared_high_block_plan_search(emu, t, k, min_rate):
    """bg_shb_plan.h (host side of k_pairs_shb): the randomised coset search finds a plan for (nearly) every random L where
    one can exist, every plan it returns satisfies the invariants the kernel relies on (nat a permutation, terms relabelled
    by word-crossing bit swaps, one high pattern per class of 32, enough free low slots), and the search is deterministic
    (the same L gives the same plan: the ranks of a multi-GPU job must agree)."""
    rs = np.random.RandomState(1000 * t + k)
    found = 0
    n = 60
    for _ in range(n):
        L = [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) for _ in range(k)]
        code, dg = emu.shb_plan_check(t, L)
        assert code >= 0, (t, k, code, L)
        assert (code, dg) == emu.shb_plan_check(t, L)
        found += code
    assert found >= min_rate * n, (t, k, found)
    # degenerate L: repeated and zero rows (columns of rank < k)
    L = [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) for _ in range(3)]
    code, _ = emu.shb_plan_check(t, (L * k)[:k - 1] + [0])
    assert code >= 0


def test_32_bit_accumulation_equals_64_bit(emu):
    """zw32_add — what a thread of k_pairs_shb adds its terms with — against zw_add over every (eps, p, m) of a pair of
    normalised states at 33 <= t <= 44, 128 terms deep (the flush interval): same four integers, no overflow."""
    assert emu.lib.emu_zw32_selftest() == 0
