"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI
(include/bgnorm.h, via circuitsimulator_b200's ctypes face) and is compared with
  * the reference's own golden vectors (tests/golden/kat_*.npz),
  * outputs of the compiled reference recorded in tests/golden/ref_*.npz,
  * the CPU oracle (oracle/) on the same seeded inputs.
Predicate for amplitudes: eps equal; if eps != 0, p equal and m equal mod 8 — the reference's own
convention (tests/units/stabtests.c:371-376).  fp64 results: 1e-12 relative."""
import os

import numpy as np
import pytest

from util import load, states, epm_equal, parse_stream, exact_sample_value, GOLDEN
from oracle.oracle import states_to_numpy, state_from_numpy, Projector as OProj

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.fixture(scope="module")
def be():
    import circuitsimulator_b200 as bg
    b = bg.Backend(0)
    yield b
    b.close()


def to_bg(P):
    import circuitsimulator_b200 as bg
    return bg.Projector.make(P.nqubits, list(P.phase[:P.nstabs]), list(P.xs[:P.nstabs]), list(P.zs[:P.nstabs]))


def exact_terms(t):
    size = (t + 1) // 2
    return 1 << size


def test_library_is_cuda_only():
    import circuitsimulator_b200 as bg
    lib = bg.load_library()
    assert os.path.basename(lib._name) == "libbgnorm.so"


def test_inner_product_reference_kat(be):
    d = load("kat_inner_product.npz")
    got = be.inner_products(d["a"], d["b"])
    for g, w in zip(got, d["epm"]):
        assert epm_equal(tuple(g), tuple(w))


def test_inner_products_vs_compiled_reference(be):
    d = load("ref_pairs.npz")
    got = be.inner_products(d["a"], d["b"])
    bad = [i for i, (g, w) in enumerate(zip(got, d["epm"])) if not epm_equal(tuple(g), tuple(w))]
    assert not bad, bad[:10]
    assert sum(1 for w in d["epm"] if w[0] == 0) > 100       # zero detection is exercised


def test_exponential_sum_kats_through_inner_product(be, oracle):
    """<+^n | K,q> = 2^-n sum_x e^{i pi q(x)/4}: the reference's 274 exponential-sum vectors
    (k = n, G = I) checked through the inner-product entry point."""
    d = load("kat_exponential_sum.npz")
    st = d["states"]
    plus = states_to_numpy([oracle.identity_state(int(s["n"]), int(s["n"])) for s in st])
    got = be.inner_products(st, plus)
    for g, w, s in zip(got, d["epm"], st):
        n = int(s["n"])
        want = (int(w[0]), int(w[1]) - 2 * n, int(w[2]))
        assert epm_equal(tuple(g), want)


def test_measure_pauli_kats(be, oracle):
    """Device measurePauli on the reference's 30 vectors.  The device keeps a different (equivalent)
    basis, so states are compared as states: same k, same affine space, and identical overlaps with
    probe states."""
    d = load("kat_measure_pauli.npz")
    out, res = be.measure_pauli(d["states_in"], d["m"], d["zeta"], d["xi"])
    want_states = states(d["states_out"])
    for i in range(len(out)):
        assert abs(res[i] - d["result"][i]) < 1e-4
        got = state_from_numpy(out[i])
        want = want_states[i]
        assert got.k == want.k
        probes = [oracle.random_state_philox(got.n, 77, 0, j) for j in range(6)] + [want]
        for pr in probes:
            assert epm_equal(oracle.inner_product(got, pr), oracle.inner_product(want, pr))
        assert epm_equal(oracle.inner_product(got, want), (1, 0, 0))      # <want|got> = 1 exactly


def test_device_rng_matches_oracle_restatement(be, oracle):
    for t in (1, 2, 5, 16, 32, 33, 40, 64):
        got = be.random_states(t, 2024, 3, 10, 24)
        for j in range(24):
            a = state_from_numpy(got[j])
            b = oracle.random_state_philox(t, 2024, 3, 10 + j)
            assert a.key(full=False) == b.key(full=False), (t, j)


def test_decomposition_terms_match_prepH_prepL(be, oracle):
    be.set_decomposition(7, True)
    got = be.decomposition_terms(0, 16)
    for i in range(16):
        a, b = state_from_numpy(got[i]), oracle.prepH(i, 7)
        assert a.k == b.k and epm_equal(oracle.inner_product(a, b), (1, 0, 0))
    L = [0b1011001110, 0b0110110101, 0b1110001011]
    be.set_decomposition(10, False, L)
    got = be.decomposition_terms(0, 8)
    for i in range(8):
        a, b = state_from_numpy(got[i]), oracle.prepL(i, 10, L)
        assert a.k == b.k and epm_equal(oracle.inner_product(a, b), (1, 0, 0))


def _fixture_names():
    return sorted(f for f in os.listdir(GOLDEN) if f.startswith("ref_samples_") and f.endswith(".npz"))


@pytest.mark.parametrize("name", _fixture_names())
def test_L_chi_loop_vs_compiled_reference(be, name):
    """theta drawn by the reference (libc rand), projected ON THE DEVICE, chi terms: per-pair
    (eps,p,m) against the compiled reference's, per-sample value against its fp64."""
    d = load(name)
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", str(d["stream"])))
    t, exact = cfg["t"], cfg["exact"]
    L = [int(x) for x in d["L"]]
    be.set_decomposition(t, exact, L)
    chi = d["epm"].shape[1]
    for which, P in enumerate((G, H)):
        sel = np.where(d["which"] == which)[0]
        out = be.sampled_norm_from_states(to_bg(P), d["theta"][sel], project=True, want_epm=True, chi=chi)
        for j, l in enumerate(sel):
            if not d["alive"][l]:
                assert out["per_sample"][j] == 0.0
                continue
            bad = [i for i in range(chi) if not epm_equal(tuple(out["epm"][j, i]), tuple(d["epm"][l, i]))]
            assert not bad, (name, l, bad[:5])
            # the device value against the EXACT value of the reference's own amplitudes (1e-12
            # relative), and against the reference's fp64, whose sequential cos/sin/pow summation
            # (innerprod.c:127-142) carries an absolute error ~1e-16 * (sum |terms|)^2
            exact, scale = exact_sample_value(d["epm"][l], t, d["projfactor"][l])
            got = out["per_sample"][j]
            assert abs(got - exact) <= RTOL * abs(exact), (got, exact)
            assert abs(got - d["value"][l]) <= RTOL * abs(exact) + 1e-13 * scale, (got, d["value"][l])
        # same thetas already projected by the reference, no device projection: same amplitudes
        alive = [l for l in sel if d["alive"][l]]
        if alive:
            out2 = be.sampled_norm_from_states(to_bg(P), d["projected"][alive], project=False, want_epm=True, chi=chi)
            for j, l in enumerate(alive):
                assert all(epm_equal(tuple(out2["epm"][j, i]), tuple(d["epm"][l, i])) for i in range(chi))


@pytest.mark.parametrize("stream,k,samples", [("htstack_t4.txt", 0, 256), ("hs_t16_bit6.txt", 0, 48),
                                              ("hs_t40_k9_bit0.txt", 6, 24), ("phase_estimation_q0.txt", 5, 24)])
def test_sampled_norm_vs_oracle_same_seed(be, oracle, stream, k, samples):
    """bg_sampled_norm (device RNG + projection + chi loop + reduction) against the oracle running
    the same Philox thetas through its restatement of singleProjectorSample."""
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", stream))
    t, exact = cfg["t"], cfg["exact"]
    rs = np.random.RandomState(1)
    L = [] if exact else [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) for _ in range(k)]
    be.set_decomposition(t, exact, L)
    for P, seed in ((G, 11), (H, 12)):
        got = be.sampled_norm(to_bg(P), samples, 1, seed, 1.0)
        tot, per = oracle.sampled_sum_philox(P, exact, L, seed, 0, 0, 1, samples)
        want = tot / samples
        assert abs(got - want) <= RTOL * max(abs(want), 1e-300), (got, want)


def test_shards_add_up(be):
    """Strided sample shards (the reference's rank stride) sum to the single-rank result."""
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", "hs_t16_bit6.txt"))
    be.set_decomposition(cfg["t"], True)
    P = to_bg(G)
    whole = be.sampled_norm(P, 1000, 1, 5, 1.0)
    be.set_allreduce(False)
    try:
        parts = []
        for r in range(3):
            be.set_shard(r, 3)
            parts.append(be.sampled_norm(P, 1000, 1, 5, 1.0))
    finally:
        be.set_shard(0, 1)
        be.set_allreduce(True)
    assert abs(sum(parts) - whole) <= 1e-13 * abs(whole)


@pytest.mark.parametrize("stream", ["htstack_t4.txt", "toffoli_q0.txt"])
def test_exact_norm_vs_oracle(be, oracle, stream):
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", stream))
    t = cfg["t"]
    be.set_decomposition(t, True)
    for P in (G, H):
        got = be.exact_norm(to_bg(P), 1.0)
        want = oracle.exact_projector(P, True, [], 1.0)
        assert abs(got - want) <= 1e-11 * max(abs(want), 1e-300), (got, want)


def test_exact_norm_L_decomposition_vs_oracle(be, oracle):
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", "phase_estimation_q0.txt"))
    t = cfg["t"]
    rs = np.random.RandomState(9)
    L = [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) for _ in range(5)]
    be.set_decomposition(t, False, L)
    got = be.exact_norm(to_bg(G), 1.0)
    want = oracle.exact_projector(G, False, L, 1.0)
    assert abs(got - want) <= 1e-11 * max(abs(want), 1e-300)


def test_backend_executable_end_to_end():
    """The unmodified front end's instruction streams through the drop-in executable
    (stdin protocol of libcirc/probability.c:74-216).  Exact-norm path: deterministic."""
    import json
    import circuitsimulator_b200 as bg
    meta = json.load(open(os.path.join(GOLDEN, "streams", "meta.json")))
    # HTstack: without -forceSample the back end switches to the exact norm (probability.c:162-167)
    txt = open(os.path.join(GOLDEN, "streams", "htstack_t4.txt")).read().split()
    txt[12] = "0"                                           # forceSample off
    num, den, _ = bg.run_backend("\n".join(txt) + "\n", env={"BG_SEED": 1})
    assert abs(num / den - meta["htstack_t4"]["expect_probability"]) < 1e-7
    num, den, _ = bg.run_backend(open(os.path.join(GOLDEN, "streams", "toffoli_111.txt")).read(), env={"BG_SEED": 1})
    assert abs(num / den - 1.0) < 1e-9                      # circuits/toffoli.circ:14-19 -> 111 with certainty
    # sampled path on the same circuit: estimator within its statistical error
    num, den, _ = bg.run_backend(open(os.path.join(GOLDEN, "streams", "htstack_t4.txt")).read(), env={"BG_SEED": 3})
    assert abs(num / den - 0.97855339) < 0.15


def test_full_size_config4_properties(be):
    """BASELINE config 4 at full size (t=40, chi=512, L=2^16): determinism under the same seed,
    independence from the work partition, and additivity of sample shards."""
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", "hs_t40_k9_bit0.txt"))
    t = cfg["t"]
    rs = np.random.RandomState(4)
    L = [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) for _ in range(9)]
    be.set_decomposition(t, False, L)
    P = to_bg(G)
    a = be.sampled_norm(P, 65536, 1, 42, 1.0)
    st = be.stats()
    assert 0.9 * 65536 * 512 <= st["pairs"] <= 65536 * 512        # annihilated samples evaluate no pairs
    b = be.sampled_norm(P, 65536, 1, 42, 1.0)
    assert a == b
    be.set_allreduce(False)
    try:
        parts = []
        for r in range(2):
            be.set_shard(r, 2)
            parts.append(be.sampled_norm(P, 65536, 1, 42, 1.0))
    finally:
        be.set_shard(0, 1)
        be.set_allreduce(True)
    assert abs(sum(parts) - a) <= 1e-13 * abs(a)
    assert a > 0


# ----------------------------------------------------------------------------- edge cases
def _random_projector(rs, t, n):
    import circuitsimulator_b200 as bg
    ph, xs, zs = [], [], []
    for _ in range(n):
        x = int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) if rs.randint(0, 2) else 0
        z = int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1)
        if t > 62:
            x |= int(rs.randint(0, 4)) << 62 if x else 0
            z |= int(rs.randint(0, 4)) << 62
        ph.append((bin(x & z).count("1") + 2 * int(rs.randint(0, 2))) % 4)      # Hermitian
        xs.append(x)
        zs.append(z)
    return bg.Projector.make(t, ph, xs, zs), OProj.make(t, ph, xs, zs)


@pytest.mark.parametrize("t,k,nst,samples", [(1, 1, 1, 64), (2, 2, 2, 64), (7, 3, 5, 48), (31, 5, 20, 24),
                                             (32, 5, 20, 24), (33, 5, 20, 16), (60, 6, 40, 10), (64, 6, 50, 10),
                                             (64, 5, 128, 8)])
def test_widths_and_generator_counts_vs_oracle(be, oracle, t, k, nst, samples):
    """State widths 1..64 (both word sizes, the 32/33 boundary, the full 64-bit row), up to the
    maximum of 128 generators; |L> decomposition with random L."""
    rs = np.random.RandomState(100 + t + nst)
    P, OP = _random_projector(rs, t, nst)
    L = [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) | ((int(rs.randint(0, 4)) << 62) if t > 62 else 0) for _ in range(k)]
    be.set_decomposition(t, False, L)
    got = be.sampled_norm(P, samples, 1, 77, 1.0)
    tot, per = oracle.sampled_sum_philox(OP, False, L, 77, 0, 0, 1, samples)
    want = tot / samples
    assert abs(got - want) <= 1e-11 * abs(want) + 1e-300, (got, want)


@pytest.mark.parametrize("t", [3, 8, 11, 16])
def test_exact_decomposition_odd_and_even_t(be, oracle, t):
    rs = np.random.RandomState(t)
    P, OP = _random_projector(rs, t, t // 2 + 1)
    be.set_decomposition(t, True)
    got = be.sampled_norm(P, 40, 1, 5, 1.0)
    tot, _ = oracle.sampled_sum_philox(OP, True, [], 5, 0, 0, 1, 40)
    assert abs(got - tot / 40) <= 1e-11 * abs(tot / 40) + 1e-300
    assert abs(be.exact_norm(P, 1.0) - oracle.exact_projector(OP, True, [], 1.0)) <= 1e-11


def test_low_dimensional_thetas_take_the_warp_kernel(be, oracle):
    """Thetas with many parity checks (k << t) are routed to the warp-per-pair kernel; host-supplied
    thetas of every dimension, including k = 0, must agree with the oracle."""
    t = 12
    rs = np.random.RandomState(3)
    L = [int(rs.randint(0, 1 << t)) for _ in range(4)]
    be.set_decomposition(t, False, L)
    thetas = []
    for j in range(40):
        s = oracle.random_state_philox(t, 9, 0, j)
        for _ in range(j % (t + 1)):                       # cut the dimension down by measuring Z-type Paulis
            oracle.measure_pauli(s, 0, int(rs.randint(1, 1 << t)), 0)
        thetas.append(s)
    import circuitsimulator_b200 as bg
    empty = bg.Projector.make(t, [], [], [])
    out = be.sampled_norm_from_states(empty, states_to_numpy(thetas), project=False, want_epm=True, chi=16)
    ks = set()
    for j, s in enumerate(thetas):
        ks.add(s.k)
        for i in range(16):
            want = oracle.inner_product(s, oracle.prepL(i, t, L))
            assert epm_equal(tuple(out["epm"][j, i]), want), (j, i, s.k)
    assert min(ks) <= 2 and max(ks) >= 10


def test_many_terms_not_staged_in_shared_memory(be, oracle):
    """chi = 4096 > the staged-table limit: terms are read from global memory."""
    t, k = 20, 12
    rs = np.random.RandomState(8)
    P, OP = _random_projector(rs, t, 9)
    L = [int(rs.randint(0, 1 << t)) for _ in range(k)]
    be.set_decomposition(t, False, L)
    got = be.sampled_norm(P, 6, 1, 1, 1.0)
    tot, _ = oracle.sampled_sum_philox(OP, False, L, 1, 0, 0, 1, 6)
    assert abs(got - tot / 6) <= 1e-11 * abs(tot / 6) + 1e-300


def test_closed_forms_and_errors(be):
    import circuitsimulator_b200 as bg
    be.set_decomposition(4, True)
    assert be.sampled_norm(bg.Projector.make(4, [], [], []), 10, 1, 0, 1.5) == 1.5 ** 2      # innerprod.c:47
    assert be.exact_norm(bg.Projector.make(4, [], [], []), 1.5) == 1.5 ** 2                  # innerprod.c:150
    with pytest.raises(bg.BGError):
        be.sampled_norm(bg.Projector.make(5, [0], [1], [0]), 10, 1, 0, 1.0)                  # width mismatch
    with pytest.raises(bg.BGError):
        be.set_decomposition(65, True)
    with pytest.raises(bg.BGError):
        be.set_decomposition(10, False, [1] * 30)                                            # k > 26
    # more shards than samples: some ranks own nothing
    be.set_allreduce(False)
    try:
        parts = []
        for r in range(4):
            be.set_shard(r, 4)
            parts.append(be.sampled_norm(bg.Projector.make(4, [0], [3], [5]), 3, 1, 2, 1.0))
    finally:
        be.set_shard(0, 1)
        be.set_allreduce(True)
    whole = be.sampled_norm(bg.Projector.make(4, [0], [3], [5]), 3, 1, 2, 1.0)
    assert abs(sum(parts) - whole) < 1e-15 and parts[3] == 0.0


def test_bins_median_of_means(be, oracle):
    """multiSampledProjector (innerprod.c:23-41): the median of the bin means (even count: mean of the middle two).
    The reference's comparator truncates differences to int (innerprod.c:17-19), so its qsort does not order values
    that are within 1 of each other; the library takes the true median (BG_REF_MEDIAN=1 gives the reference's call)."""
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", "htstack_t4.txt"))
    be.set_decomposition(4, True)
    for bins in (3, 4, 5):
        got = be.sampled_norm(to_bg(G), 50, bins, 21, 1.0)
        means = sorted(oracle.sampled_sum_philox(G, True, [], 21, b, 0, 1, 50)[0] / 50 for b in range(bins))
        want = means[bins // 2] if bins % 2 else (means[bins // 2] + means[bins // 2 - 1]) / 2
        assert abs(got - want) < 1e-12


def test_two_projector_job_and_graph_replay(be):
    """bg_sampled_norm2 / prepare2+run+finish2 (one CUDA-graph replay, one all-reduce) give exactly the
    numbers of two separate bg_sampled_norm calls, on every replay."""
    import circuitsimulator_b200 as bg
    import torch
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", "hs_t16_bit6.txt"))
    be.set_decomposition(cfg["t"], True)
    g, h = to_bg(G), to_bg(H)
    a = be.sampled_norm(g, 3000, 1, 31, 1.0)
    b = be.sampled_norm(h, 3000, 1, 32, 1.0)
    assert be.sampled_norm2(g, h, 3000, 1, 31, 32, 1.0) == (a, b)
    be.sampled_prepare2(g, h, 3000, 2, 31, 32)
    first = None
    for _ in range(3):
        be.sampled_run()
        out = be.sampled_finish2(1.0)
        first = first or out
        assert out == first
    st = be.stats()
    assert st["launches"] > 0 and st["pairs"] > 0
    # bins = 1 job equals the plain calls
    be.sampled_prepare2(g, h, 3000, 1, 31, 32)
    be.sampled_run()
    assert be.sampled_finish2(1.0) == (a, b)
    # ... and is one fused launch sequence for both projectors (prepare, pairs, pairs[many checks], finalize)
    st = be.stats()
    assert st["pair_launches"] == 1 and st["launches"] == 4
    assert 0 < st["pairs"] <= 2 * 3000 * exact_terms(cfg["t"])          # annihilated samples meet no term


def test_pipelined_runs_two_in_flight(be):
    import circuitsimulator_b200 as bg
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", "hs_t16_bit6.txt"))
    be.set_decomposition(cfg["t"], True)
    g, h = to_bg(G), to_bg(H)
    want = be.sampled_norm2(g, h, 2000, 1, 7, 8, 1.0)
    be.sampled_prepare2(g, h, 2000, 1, 7, 8)
    be.sampled_run()
    be.sampled_run()
    with pytest.raises(bg.BGError):
        be.sampled_run()                      # a third job would overwrite an unread result
    assert be.sampled_finish2(1.0) == want
    be.sampled_run()
    assert be.sampled_finish2(1.0) == want
    assert be.sampled_finish2(1.0) == want
    with pytest.raises(bg.BGError):
        be.sampled_finish2(1.0)               # nothing in flight


@pytest.mark.parametrize("overlap", ["0", "1"])
def test_new_projectors_staged_while_a_job_is_in_flight(overlap):
    """The end-to-end pipeline bench.py times: bg_sampled_prepare2 with NEW projector bytes while the previous job of
    the same shape has not been collected (staging sets alternate, the upload is ordered behind the running job on
    its stream).  Each job must return what the same projectors give one call at a time — with consecutive jobs on
    one stream (BG_OVERLAP=0) and in overlap mode (every second job on a stream and buffers of its own, which is what
    a job with few samples per GPU gets by default)."""
    import circuitsimulator_b200 as bg
    os.environ["BG_OVERLAP"] = overlap
    try:
        be = bg.Backend(0)
    finally:
        del os.environ["BG_OVERLAP"]
    try:
        _staged_in_flight(be, overlap == "1")
    finally:
        be.close()


def _staged_in_flight(be, overlapped):
    import circuitsimulator_b200 as bg
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", "hs_t16_bit6.txt"))
    be.set_decomposition(cfg["t"], True)
    g = to_bg(G)                 # (the stream's own H' projects onto the same space as G': the hidden shift is deterministic)
    h, _ = _random_projector(np.random.RandomState(4), cfg["t"], 5)
    pairs = [(g, h), (h, g), (g, g), (h, h), (g, h)]
    want = [be.sampled_norm2(a, b, 5000, 1, 7, 8, 1.0) for a, b in pairs]
    assert len(set(want)) == len(want) - 1
    got = []
    be.sampled_prepare2(*pairs[0], 5000, 1, 7, 8)
    be.sampled_run()
    for a, b in pairs[1:]:
        be.sampled_prepare2(a, b, 5000, 1, 7, 8)        # previous job still in flight
        be.sampled_run()
        got.append(be.sampled_finish2(1.0))
    got.append(be.sampled_finish2(1.0))
    assert got == want
    assert be.stats()["overlapped"] == (1 if overlapped else 0)
    # the per-sample values of the last finished job come from the set it ran in
    per = be.sampled_per_sample(0, 0, 5000)
    assert abs(per.sum() / 5000 - want[-1][0]) <= 1e-12 * abs(want[-1][0])
    # a prepared job replayed many times alternates between the sets
    be.sampled_run()
    be.sampled_run()
    assert be.sampled_finish2(1.0) == want[-1] and be.sampled_finish2(1.0) == want[-1]
    # a job of another shape drops what is in flight (its buffers change) instead of returning stale numbers
    be.sampled_run()
    be.sampled_prepare2(g, h, 1234, 1, 7, 8)
    with pytest.raises(bg.BGError):
        be.sampled_finish2(1.0)
    be.sampled_run()
    assert be.sampled_finish2(1.0) == be.sampled_norm2(g, h, 1234, 1, 7, 8, 1.0)


def test_new_decomposition_of_the_same_shape_keeps_the_prepared_job(be):
    """sampleQubits draws a new L for every probability() call.  bg_set_decomposition with another L of the same
    (t, k) keeps the prepared job and its captured CUDA graphs (the tables change, no kernel argument does — the
    relabelling of the shared high-block plan sits in device memory); the results must be those of a context that
    has seen only that L.  Changing k or t in between must not leave anything stale either."""
    import circuitsimulator_b200 as bg
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", "hs_t40_k9_bit0.txt"))
    t = cfg["t"]
    g, h = to_bg(G), to_bg(H)
    rs = np.random.RandomState(11)
    Ls = [_bench_L(9, t)] + [[int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) for _ in range(9)] for _ in range(5)]
    fresh = []
    for L in Ls:
        other = bg.Backend(0)
        try:
            other.set_decomposition(t, False, L)
            fresh.append(other.sampled_norm2(g, h, 600, 1, 21, 22, 1.0))
        finally:
            other.close()
    assert len(set(fresh)) == len(fresh)
    for rounds in range(2):
        for L, want in zip(Ls, fresh):
            be.set_decomposition(t, False, L)
            assert be.sampled_norm2(g, h, 600, 1, 21, 22, 1.0) == want
    # split-phase, two in flight, L changing between the submissions
    be.set_decomposition(t, False, Ls[0])
    be.sampled_prepare2(g, h, 600, 1, 21, 22)
    be.sampled_run()
    be.set_decomposition(t, False, Ls[1])
    be.sampled_prepare2(g, h, 600, 1, 21, 22)
    be.sampled_run()
    assert be.sampled_finish2(1.0) == fresh[0]
    assert be.sampled_finish2(1.0) == fresh[1]
    # another shape in between
    be.set_decomposition(t, False, Ls[2][:8])
    a = be.sampled_norm2(g, h, 600, 1, 21, 22, 1.0)
    be.set_decomposition(t, False, Ls[2])
    assert be.sampled_norm2(g, h, 600, 1, 21, 22, 1.0) == fresh[2]
    be.set_decomposition(t, False, Ls[2][:8])
    assert be.sampled_norm2(g, h, 600, 1, 21, 22, 1.0) == a


def test_state_widths_in_any_order_on_one_context(be):
    """One context, sampled jobs at t = 60, 40, 64, 33, 60 in this order: the draw + projection kernel's shared-memory
    budget is a per-kernel attribute — a narrower state seen later must not shrink it under a wider one seen before
    (each width's occupancy is cached).  Every job must run and repeat its own value."""
    rs = np.random.RandomState(8)
    seen = {}
    for t in (60, 40, 64, 33, 60, 40):
        P, _ = _random_projector(rs, t, 12) if t not in seen else (seen[t][0], None)
        L = [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) for _ in range(6)] if t not in seen else seen[t][1]
        be.set_decomposition(t, False, L)
        val = be.sampled_norm2(P, P, 300, 1, 3, 4, 1.0)
        if t in seen:
            assert val == seen[t][2]
        seen[t] = (P, L, val)


def test_persistent_server_mode(tmp_path):
    """`bgbackend --serve <socket>` keeps the CUDA contexts alive across probability() calls; a client
    started with BG_SERVER=<socket> relays the same protocol."""
    import subprocess
    import time
    import circuitsimulator_b200 as bg
    sock = str(tmp_path / "bg.sock")
    srv = subprocess.Popen([bg.BACKEND_PATH, "--serve", sock], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    try:
        assert b"serving" in srv.stdout.readline()
        txt = open(os.path.join(GOLDEN, "streams", "toffoli_111.txt")).read()
        direct = bg.run_backend(txt, env={"BG_SEED": 1})
        t0 = time.perf_counter()
        served = [bg.run_backend(txt, env={"BG_SERVER": sock}) for _ in range(3)]
        dt = (time.perf_counter() - t0) / 3
        for num, den, _ in served:                       # exact-norm path: deterministic
            assert num == direct[0] and den == direct[1]
        # sampled path through the server
        num, den, _ = bg.run_backend(open(os.path.join(GOLDEN, "streams", "htstack_t4.txt")).read(), env={"BG_SERVER": sock})
        assert abs(num / den - 0.97855339) < 0.15
        print("served probability() back-end call: %.1f ms" % (1e3 * dt))
    finally:
        import socket as pysock
        try:
            c = pysock.socket(pysock.AF_UNIX, pysock.SOCK_STREAM)
            c.connect(sock)
            c.sendall(b"shutdown\n")
            c.close()
            srv.wait(timeout=20)
        except Exception:
            srv.kill()


def test_bulk_random_pairs_vs_oracle(be, oracle):
    """20k generic pairs (device-drawn states, several widths; half of them cut to low dimension) through
    bg_inner_products against the oracle."""
    rs = np.random.RandomState(12)
    total = zeros = 0
    for t, n in ((5, 4000), (16, 6000), (40, 6000), (64, 4000)):
        a = be.random_states(t, 31, 0, 0, n)
        b = be.random_states(t, 32, 0, 0, n)
        # cut every other b down with random Z-type Paulis (device measurePauli)
        for _round in range(3):
            idx = np.arange(0, n, 2)
            zeta = np.array([int(rs.randint(1, 2 ** 62)) & ((1 << t) - 1) or 1 for _ in idx], dtype=np.uint64)
            cut, res = be.measure_pauli(b[idx], np.zeros(len(idx), dtype=np.int32), zeta, np.zeros(len(idx), dtype=np.uint64))
            keep = res > 0
            b[idx[keep]] = cut[keep]
        got = be.inner_products(a, b)
        for i in range(0, n, 1):
            want = oracle.inner_product(state_from_numpy(a[i]), state_from_numpy(b[i]))
            assert epm_equal(tuple(got[i]), want), (t, i, tuple(got[i]), want)
            zeros += want[0] == 0
        total += n
    assert total == 20000 and zeros > 500


def test_device_rng_law(be, oracle):
    """The device sampler follows randomStabilizerState's law: the dimension deficit d = t - k has the
    eq. 79 distribution (chi-square against stabilizer.c:693-716's cdf), h and D bits are fair."""
    t, n = 16, 60000
    st = be.random_states(t, 555, 0, 0, n)
    d = t - st["k"].astype(int)
    cdf = oracle.dimension_cdf(t)
    pmf = np.diff(np.concatenate([[0.0], cdf]))
    chi2 = 0.0
    for v in range(0, 6):
        exp = n * pmf[v]
        obs = int(np.sum(d == v))
        if exp > 5:
            chi2 += (obs - exp) ** 2 / exp
    assert chi2 < 30.0, chi2                      # 5 dof: P(chi2 > 30) ~ 1e-5
    hbits = np.unpackbits(np.ascontiguousarray(st["h"]).view(np.uint8)).sum() / (n * t)
    assert abs(hbits - 0.5) < 0.01
    full = st[st["k"] == t]
    d1 = np.unpackbits(np.ascontiguousarray(full["D1"]).view(np.uint8)).sum() / (len(full) * t)
    d2 = np.unpackbits(np.ascontiguousarray(full["D2"]).view(np.uint8)).sum() / (len(full) * t)
    assert abs(d1 - 0.5) < 0.01 and abs(d2 - 0.5) < 0.01
    # J symmetric with J_aa = D1_a, off-diagonal bits fair
    J = np.ascontiguousarray(full["J"][:, :t])
    D1full = np.ascontiguousarray(full["D1"])
    offdiag = 0
    for a in range(t):
        for b in range(a):
            bit_ab = (J[:, a] >> np.uint64(b)) & np.uint64(1)
            bit_ba = (J[:, b] >> np.uint64(a)) & np.uint64(1)
            assert np.array_equal(bit_ab, bit_ba)
            offdiag += bit_ab.sum()
        assert np.array_equal((J[:, a] >> np.uint64(a)) & np.uint64(1), (D1full >> np.uint64(a)) & np.uint64(1))
    assert abs(offdiag / (len(full) * t * (t - 1) / 2) - 0.5) < 0.01


def _fidelity_cases():
    import json
    return json.load(open(os.path.join(GOLDEN, "ref_fidelity.json")))["cases"]


def test_decomposition_weights_vs_oracle(be, oracle):
    """decompose()'s fidelity loop (libcirc/probability.c:373-391) on the device: the weight histogram of the
    2^k combinations of L's rows, bit for bit against the oracle's restatement of that loop — on the L matrices
    the compiled reference drew (tests/golden/ref_fidelity.json) and on edge cases (k = 0, 1, t = 64, a k large
    enough for every thread to walk a Gray-code run)."""
    cases = [(c["t"], [int(r) for r in c["L_rows"]]) for c in _fidelity_cases()]
    rs = np.random.RandomState(11)
    cases += [(5, []), (1, [1]), (64, [int(rs.randint(0, 2 ** 62)) | (1 << 63)]),
              (64, [int(rs.randint(0, 2 ** 62)) * 4 + int(rs.randint(0, 4)) for _ in range(16)]),
              (47, [int(rs.randint(0, 2 ** 47)) for _ in range(21)]),
              (9, [0, 0, 5])]                                   # rank-deficient L: combinations repeat
    for t, rows in cases:
        want_z, want_hist = oracle.decompose_ZL(t, [r & ((1 << t) - 1) for r in rows])
        got = be.decomposition_weights(t, rows)
        assert got == want_hist, (t, len(rows))
        assert sum(got) == 1 << len(rows)
        z = sum(float(h) * 2.0 ** (-(w // 2)) for w, h in enumerate(got))
        assert z == want_z, (t, len(rows), z, want_z)


def test_backend_fidelity_option_matches_compiled_reference():
    """The drop-in executable with fidelity = 1: same L from libc rand(), Z(L) from the device histogram; for
    empty projectors the two printed lines are norm^2 = 2^k Z(L) (innerprod.c:47) — the same 17 digits as the
    compiled reference printed — and the chatter carries the same delta."""
    import re
    import circuitsimulator_b200 as bg
    for c in _fidelity_cases():
        tok = [0, 0, 0, 1, 1, c["t"], c["k"], 0, 1e-05, 1, 0, 1, 1, 0, 0, 0, 0]
        num, den, lines = bg.run_backend("\n".join(str(v) for v in tok) + "\n", env={"BG_SEED": 1})
        assert "%.17e" % num == c["numerator"] and "%.17e" % den == c["denominator"], (c["t"], c["k"], lines[-2:])
        delta = [float(re.search(r"delta = 1 - <H\^t\|L>: ([-0-9.eE]+)", ln).group(1)) for ln in lines if "delta" in ln]
        assert delta and delta[-1] == c["delta_printed"]


# ----------------------------------------------------------------------------- parity at the headline sizes
def _bench_L(k, t):
    rs = np.random.RandomState(20240)                       # bench.py: fixed_L
    return [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) for _ in range(k)]


def test_full_size_run_spot_checked_against_oracle(be, oracle):
    """BASELINE config 4 as bench.py runs it (t=40, chi=512, L=2^16, both projectors in one fused job, device
    RNG): the per-sample values of the FULL run are read back (bg_sampled_per_sample) and 64 random sample
    indices per projector are re-computed by the oracle from the same Philox theta — value-level parity at
    the headline size, through the shared high-block kernel."""
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", "hs_t40_k9_bit0.txt"))
    t = cfg["t"]
    L = _bench_L(9, t)
    be.set_decomposition(t, False, L)
    samples = 65536
    num, den = be.sampled_norm2(to_bg(G), to_bg(H), samples, 1, 1234, 5678, 1.0)
    rs = np.random.RandomState(5)
    for pj, (P, seed) in enumerate(((G, 1234), (H, 5678))):
        per = be.sampled_per_sample(pj, 0, samples)
        assert abs(per.sum() / samples - (num, den)[pj]) <= 1e-12 * abs((num, den)[pj])
        for l in rs.randint(0, samples, 64):
            th = oracle.random_state_philox(t, seed, 0, int(l))
            want = oracle.sample_from_theta(th, P, False, L, want_epm=False)["value"]
            assert abs(per[l] - want) <= RTOL * max(abs(want), 1e-300), (pj, int(l), per[l], want)


@pytest.mark.parametrize("t,k,nsamples,stream", [(40, 9, 200, "hs_t40_k9_bit0.txt"), (60, 8, 400, None)])
def test_hundred_thousand_pairs_vs_oracle(be, oracle, t, k, nsamples, stream):
    """>= 10^5 inner products per configuration, amplitude by amplitude (eps, p, m mod 8), against the oracle on
    the same device-drawn, projected thetas: t=40 / chi=512 (shared high-block kernel) and t=60 / chi=256
    (generic 64-bit kernel)."""
    if stream:
        cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", stream))
        P, OP = to_bg(G), G
    else:
        P, OP = _random_projector(np.random.RandomState(60), t, 30)
    L = _bench_L(k, t)
    be.set_decomposition(t, False, L)
    chi = 1 << k
    thetas = be.random_states(t, 77, 0, 0, nsamples)
    got = be.sampled_norm_from_states(P, thetas, project=True, want_epm=True, chi=chi)
    pairs = bad = 0
    for l in range(nsamples):
        want = oracle.sample_from_theta(state_from_numpy(thetas[l]), OP, False, L)
        if want["alive"]:
            pairs += chi
            bad += sum(not epm_equal(tuple(int(v) for v in got["epm"][l, i]), tuple(int(v) for v in want["epm"][i]))
                       for i in range(chi))
        assert abs(got["per_sample"][l] - want["value"]) <= RTOL * max(abs(want["value"]), 1e-300)
    assert bad == 0
    assert pairs >= 100000


@pytest.mark.parametrize("stream", ["hs_t40_k9_bit0.txt", "hs_t16_bit6.txt"])
def test_projection_of_two_thousand_thetas_vs_oracle(be, oracle, stream):
    """The projection step (measurePauli per generator: alive, number of 2^-1/2 factors, dimension k1) for 2000
    device-drawn thetas per real projector: the per-sample value against a small decomposition carries all three
    (annihilated -> 0; npf and k1 enter the power of two)."""
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", stream))
    t = cfg["t"]
    L = _bench_L(3, t)
    be.set_decomposition(t, False, L)
    thetas = be.random_states(t, 99, 0, 0, 2000)
    for P in (G, H):
        got = be.sampled_norm_from_states(to_bg(P), thetas, project=True)
        wants = [oracle.sample_from_theta(state_from_numpy(th), P, False, L, want_epm=False) for th in thetas]
        # the device sums exactly (integers); the oracle follows the reference's sequential cos / sin sum, which
        # leaves ~1e-16 of the largest term where the exact sum is 0 — hence a floor on the tolerance
        floor = RTOL * float(np.mean([abs(w["value"]) for w in wants]))
        for l, want in enumerate(wants):
            if not want["alive"]:
                assert got["per_sample"][l] == 0.0
            assert abs(got["per_sample"][l] - want["value"]) <= max(RTOL * abs(want["value"]), floor)


def test_statistical_gate_htstack(be):
    """End-to-end estimator check: circuits/HTstack.circ with 4 T gates, P(0) = 0.97855339
    (HTstack.circ:10).  Mean over 64 seeds of the sampled ratio numerator / denominator, L = 1024 each, must sit
    within 3 standard errors of the exact value — a sign or normalisation error in the estimator fails this."""
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", "htstack_t4.txt"))
    be.set_decomposition(cfg["t"], True)
    Gb, Hb = to_bg(G), to_bg(H)
    ratios = []
    for seed in range(64):
        num, den = be.sampled_norm2(Gb, Hb, 1024, 1, 1000 + seed, 5000 + seed, 1.0)
        ratios.append(num / den)
    ratios = np.array(ratios)
    se = ratios.std(ddof=1) / np.sqrt(len(ratios))
    assert se < 0.01
    assert abs(ratios.mean() - 0.97855339) <= 3 * se + 1e-3, (ratios.mean(), se)   # 1e-3: bias of a ratio estimator ~ var/L


def test_statistical_gate_sampled_vs_exact_norm_t16(be, oracle):
    """Config 3 (hidden shift, t=16, exact decomposition chi=256): the sampled estimate of ||Pi|H^t>||^2 over
    2^14 samples x 8 seeds against the exact-norm path on the same projector, within 3 standard errors."""
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", "hs_t16_bit6.txt"))
    be.set_decomposition(cfg["t"], True)
    for P in (G, H):
        Pb = to_bg(P)
        exact = be.exact_norm(Pb, 1.0)
        est = np.array([be.sampled_norm(Pb, 16384, 1, 300 + s, 1.0) for s in range(8)])
        se = est.std(ddof=1) / np.sqrt(len(est))
        assert abs(est.mean() - exact) <= 3 * se, (est.mean(), exact, se)
        assert se < 0.05 * exact


# ----------------------------------------------------------------------------- build / launch variants, multi-GPU back end
def _backend_stream(name, samples=None, k=None):
    txt = open(os.path.join(GOLDEN, "streams", name)).read().split()
    if samples is not None:
        txt[3] = str(samples)
    if k is not None:
        txt[6] = str(k)
    return "\n".join(txt) + "\n"


@pytest.mark.parametrize("env", [{"BG_KERNEL": "warp"}, {"BG_FUSE2": "0"}, {"BG_SHB": "0"}, {"BG_GRAPH": "0"},
                                 {"BG_ITEMS_FACTOR": "1"}, {"BG_CTAS_PER_SM": "2"}, {"BG_SHB": "0", "BG_LAM_MAX": "4"},
                                 {"BG_SHB": "0", "BG_LAM_MAX": "0"}, {"BG_PREP": "warp"}, {"BG_OVERLAP": "1"},
                                 {"BG_OVERLAP": "0"}, {"BG_PIECES": "2"}],
                         ids=lambda e: "_".join("%s=%s" % kv for kv in e.items()))
def test_launch_variants_give_identical_sums(env):
    """Every default-off variant on the hardware: the warp-per-pair kernel (the north-star mapping, BG_KERNEL=warp),
    one launch sequence per projector (BG_FUSE2=0), the generic 64-bit kernel instead of the shared high-block one
    (BG_SHB=0) with theta's parity checks carried as Lagrange variables (BG_LAM_MAX=4) or pivoted per term (0), no CUDA
    graph, other work partitions, the warp-per-sample draw + projection kernel (BG_PREP=warp), overlap mode forced on
    and off, another split of the last wave (BG_PIECES).  Per-sample sums are exact integers, so numerator and
    denominator must agree to the last bits (only the final fp64 sum over samples depends on the order)."""
    import circuitsimulator_b200 as bg
    text = _backend_stream("hs_t40_k9_bit0.txt", samples=2048)
    base = dict(BG_SEED=5)
    num0, den0, _ = bg.run_backend(text, env=base)
    num1, den1, _ = bg.run_backend(text, env=dict(base, **env))
    assert abs(num1 - num0) <= 4e-16 * abs(num0) and abs(den1 - den0) <= 4e-16 * abs(den0), (num0, num1, den0, den1)


def test_multi_gpu_backend_matches_single_gpu():
    """bgbackend with BG_GPUS=2 (one host thread + context per GPU, in-process NCCL all-reduce): sampled path,
    exact-norm path and the persistent server against the single-GPU answers.  Needs two visible GPUs."""
    import torch
    import circuitsimulator_b200 as bg
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    text = _backend_stream("hs_t40_k9_bit0.txt", samples=4096)
    n1, d1, _ = bg.run_backend(text, env=dict(BG_SEED=5, BG_GPUS=1))
    tof = open(os.path.join(GOLDEN, "streams", "toffoli_111.txt")).read()          # exact-norm path (2 doubles per projector)
    a1, b1, _ = bg.run_backend(tof, env=dict(BG_SEED=5, BG_GPUS=1))
    bins3 = text.split()
    bins3[4] = "3"                                                                   # median of three bins: not additive over ranks
    bins3 = "\n".join(bins3) + "\n"
    m1, e1, _ = bg.run_backend(bins3, env=dict(BG_SEED=5, BG_GPUS=1))
    # one-shot default: the ranks' parts are added on the host (no communicator per call); BG_REDUCE=nccl: in-library all-reduce
    for red in ("host", "nccl"):
        env = dict(BG_SEED=5, BG_GPUS=2, BG_REDUCE=red)
        n2, d2, _ = bg.run_backend(text, env=env)
        assert abs(n2 - n1) <= 1e-13 * abs(n1) and abs(d2 - d1) <= 1e-13 * abs(d1), red
        a2, b2, _ = bg.run_backend(tof, env=env)
        assert abs(a2 - a1) <= 1e-12 * abs(a1) and abs(b2 - b1) <= 1e-12 * abs(b1), red
        assert abs(a2 / b2 - 1.0) < 1e-9
        m2, e2, _ = bg.run_backend(bins3, env=env)
        assert abs(m2 - m1) <= 1e-13 * abs(m1) and abs(e2 - e1) <= 1e-13 * abs(e1), red
    # more GPUs than the box has: an Error line, not a hang
    with pytest.raises(bg.BGError):
        bg.run_backend(text, env=dict(BG_GPUS=64), timeout=120)


def test_level2_binding_reference_host_on_the_c_abi(be, reference):
    """INTEGRATION.md level 2, built for real (oracle/Makefile: _ref/mpibackend_bg): the reference's UNMODIFIED
    probability.c (main / master / decompose) linked against the C ABI instead of libcirc/innerprod.c, passing its own
    struct Projector / BitMatrix byte arrays through the *_bitmatrix adapters.  It must print the numbers the drop-in
    executable prints for the same stream and seed (sampled path, |L> path with decompose()'s L, exact-norm path),
    and bg_set_decomposition_bitmatrix must build the same terms as the packed-row entry."""
    import subprocess
    import circuitsimulator_b200 as bg
    exe = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "mpibackend_bg")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/mpibackend_bg not built (needs the reference tree at build time)")
    for name, kw in [("htstack_t4.txt", {}), ("hs_t16_bit6.txt", dict(samples=512)), ("hs_t40_k9_bit0.txt", dict(samples=256)),
                     ("toffoli_111.txt", {})]:
        text = _backend_stream(name, **kw)
        env = dict(os.environ, BG_SEED="11")
        p = subprocess.run([exe, "stdin"], input=text.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=300)
        lines = p.stdout.decode().splitlines()
        num2, den2 = float(lines[-2]), float(lines[-1])
        num1, den1, _ = bg.run_backend(text, env={"BG_SEED": 11})
        assert abs(num2 - num1) <= 1e-12 * max(abs(num1), 1e-300), (name, num1, num2)
        assert abs(den2 - den1) <= 1e-12 * max(abs(den1), 1e-300), (name, den1, den2)
    # the decomposition adapter on the reference's BitMatrix.data
    t = 40
    L = _bench_L(9, t)
    import ctypes as C
    be.set_decomposition(t, False, L)
    a = be.decomposition_terms(0, 512)
    be.set_decomposition(16, True)                          # so that the next call cannot be recognised as unchanged
    buf = reference.L_bytes(t, L)
    lib = bg.load_library()
    assert lib.bg_set_decomposition_bitmatrix(be.ctx, t, 0, 9, C.cast(buf, C.POINTER(C.c_uint8))) == 0
    b = be.decomposition_terms(0, 512)
    assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("t,k", [(33, 9), (36, 9), (40, 9), (40, 10), (40, 12), (40, 14), (44, 9), (44, 11), (44, 13)])
def test_shared_high_block_kernel_equals_generic_kernel(t, k):
    """k_pairs_shb against the generic 64-bit kernel (BG_SHB=0) on the same device-drawn thetas, per-sample values
    bit for bit (both accumulate exact integers), for several random decompositions per (t, k): every width the plan
    covers (1 to 12 shared variables), class sizes 32 and more, and whatever the plan search makes of each L —
    including "no plan", where both contexts run the generic kernel."""
    import circuitsimulator_b200 as bg
    a = bg.Backend(0)
    os.environ["BG_SHB"] = "0"
    try:
        b = bg.Backend(0)
    finally:
        del os.environ["BG_SHB"]
    try:
        rs = np.random.RandomState(100 * t + k)
        P, _ = _random_projector(rs, t, 24)
        for trial in range(3):
            L = [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) for _ in range(k)]
            a.set_decomposition(t, False, L)
            b.set_decomposition(t, False, L)
            th = a.random_states(t, 7 + trial, 0, 0, 192 if k < 12 else 48)
            ra = a.sampled_norm_from_states(P, th, project=True)
            rb = b.sampled_norm_from_states(P, th, project=True)
            assert np.array_equal(ra["per_sample"], rb["per_sample"]), (t, k, trial)
            assert ra["per_sample"].max() > 0
    finally:
        a.close()
        b.close()


@pytest.mark.parametrize("stream,k,exact", [("hs_t40_k9_bit0.txt", 6, False), ("hs_t16_bit6.txt", 0, True),
                                            ("htstack_t4.txt", 0, True), ("phase_estimation_q0.txt", 5, False),
                                            ("random_t33", 5, False), ("random_t64", 4, False), ("random_t32", 4, False),
                                            ("zchecks_t40", 6, False)])
def test_thread_per_sample_prepare_equals_warp_per_sample(stream, k, exact):
    """k_prepare_tps (one thread per sample, bg_prep.cuh — the hot path's draw + projection) against the
    warp-per-sample k_prepare (BG_PREP=warp) on the device: the same seeds give the same per-sample values bit for bit
    (exact integer sums), for a sample count that is not a multiple of 32, both projectors of a fused job, 32- and
    64-bit words, projectors that annihilate samples and projectors that leave many parity checks."""
    import circuitsimulator_b200 as bg
    a = bg.Backend(0)
    os.environ["BG_PREP"] = "warp"
    try:
        b = bg.Backend(0)
    finally:
        del os.environ["BG_PREP"]
    try:
        if stream.startswith("random_t") or stream.startswith("zchecks_t"):
            t = int(stream.split("_t")[1])
            rs = np.random.RandomState(t)
            if stream.startswith("zchecks"):       # commuting Z-type generators: a new parity check per generator, some kill
                zs = [int(rs.randint(0, 2 ** 62)) & int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) for _ in range(9)]
                G = bg.Projector.make(t, [0, 2, 0, 2, 0, 0, 2, 0, 0], [0] * 9, zs)
                H = bg.Projector.make(t, [0, 2, 0, 2, 0, 0, 2, 0, 2, 0], [0] * 10, zs + [zs[0] ^ zs[3]])
            else:
                G, _ = _random_projector(rs, t, 2 * t)
                H, _ = _random_projector(rs, t, 7)
        else:
            cfg, Gp, Hp = parse_stream(os.path.join(GOLDEN, "streams", stream))
            t = cfg["t"]
            G, H = to_bg(Gp), to_bg(Hp)
        L = [] if exact else _bench_L(k, t)
        samples = 1000 + 13
        dead = 0
        for be in (a, b):
            be.set_decomposition(t, exact, L)
        ra = a.sampled_norm2(G, H, samples, 1, 11, 12, 1.0)
        rb = b.sampled_norm2(G, H, samples, 1, 11, 12, 1.0)
        assert ra == rb
        for pj in (0, 1):
            pa, pb = a.sampled_per_sample(pj, 0, samples), b.sampled_per_sample(pj, 0, samples)
            assert np.array_equal(pa, pb), (stream, pj, int(np.argmax(pa != pb)))
            dead += int((pa == 0).sum())
        assert a.sampled_norm(G, 77, 1, 5, 1.0) == b.sampled_norm(G, 77, 1, 5, 1.0)          # one projector, < 3 warps
        assert dead < 2 * samples
    finally:
        a.close()
        b.close()


def test_phase_estimation_chain_approaches_the_exact_distribution():
    """BASELINE config 5 end to end: the probability() calls sampleQubits makes on circuits/phaseEstimation.circ
    (one per sampled qubit, conditioned on the outcomes before it: libcirc/sample.py:32-83), written by the unmodified
    front end, through the drop-in back end.  A state-vector simulation of the 4-qubit circuit gives
    P(q0=0) = 0.92678, P(q0=0,q1=0) = 0.21339, P(q0=0,q1=1,q2=0) = 0.5, P(q0=1,q1=0) = 0.03661, and 0 for 000 / 100 /
    110; with the |L> approximation at k = 14 (t = 33) the sampled estimates must sit within 0.03 of them — an
    end-to-end check of decomposition, projection, inner products and normalisation that does not involve the oracle."""
    import json
    import circuitsimulator_b200 as bg
    S = os.path.join(GOLDEN, "streams")
    meta = json.load(open(os.path.join(S, "meta.json")))["phase_estimation_chain"]["streams"]
    exact = {"0": 0.9267766952966369, "00": 0.21338834764831827, "010": 0.5, "000": 0.0,
             "10": 0.03661165235168153, "100": 0.0, "110": 0.0}
    for key, ent in sorted(meta.items()):
        txt = open(os.path.join(S, ent["stream"])).read().split()
        txt[6] = "14"
        num, den, _ = bg.run_backend("\n".join(txt) + "\n", env={"BG_SEED": 3})
        p = 0.0 if num == 0 else 2.0 ** ent["v_minus_u"] * num / den
        assert abs(p - exact[key]) < 0.03, (key, p, exact[key])
