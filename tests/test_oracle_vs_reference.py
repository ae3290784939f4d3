"""Pin the oracle against the COMPILED reference (oracle/_ref/libcircref.so: the unmodified C sources).
Skipped where oracle/_ref was not prebuilt."""
import os

import numpy as np

from util import load, states, epm_equal, parse_stream, GOLDEN


def test_random_states_and_inner_products_identical(oracle, reference):
    for n in (1, 2, 3, 5, 8, 16, 17, 33, 40, 64):
        for seed in range(4):
            reference.srand(seed)
            a, b = reference.random_state(n), reference.random_state(n)
            oracle.srand(seed)
            a2, b2 = oracle.random_state_libc(n), oracle.random_state_libc(n)
            assert a.key() == a2.key() and b.key() == b2.key()          # same libc rand() stream, same state
            assert reference.inner_product(a, b) == oracle.inner_product(a, b)   # raw ints, quirks included


def test_measure_pauli_shrink_sequences_identical(oracle, reference):
    rs = np.random.RandomState(0)
    for n in (4, 9, 16, 40):
        reference.srand(n)
        s = reference.random_state(n)
        s2 = s.copy()
        for _ in range(2 * n):
            x = int(rs.randint(0, 2 ** 62)) & ((1 << n) - 1) if rs.randint(0, 3) else 0
            z = int(rs.randint(0, 2 ** 62)) & ((1 << n) - 1)
            m = (bin(x & z).count("1") + 2 * int(rs.randint(0, 2))) % 4
            r1, r2 = reference.measure_pauli(s, m, z, x), oracle.measure_pauli(s2, m, z, x)
            assert r1 == r2
            assert s.key() == s2.key()
            if r1 == 0:
                break


def test_prep_states_identical(oracle, reference):
    for t in (1, 2, 5, 8, 11):
        for i in range(1 << ((t + 1) // 2)):
            assert reference.prepH(i, t).key() == oracle.prepH(i, t).key()
    L = [0b1011001110, 0b0110110101, 0b1110001011]
    for i in range(8):
        assert reference.prepL(i, 10, L).key() == oracle.prepL(i, 10, L).key()


def test_reference_fixture_pairs(oracle):
    """ref_pairs.npz was produced by the compiled reference; the oracle must reproduce it exactly."""
    d = load("ref_pairs.npz")
    for s1, s2, w in zip(states(d["a"]), states(d["b"]), d["epm"]):
        assert oracle.inner_product(s1, s2) == tuple(int(v) for v in w)


def test_single_projector_sample_bitwise(oracle, reference):
    """singleProjectorSample (innerprod.c:88-144): same libc seed -> identical double."""
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", "htstack_t4.txt"))
    for seed in range(20):
        reference.srand(seed)
        want = reference.single_projector_sample(G, True, [])
        oracle.srand(seed)
        th = oracle.random_state_libc(cfg["t"])
        got = oracle.sample_from_theta(th, G, True, [])["value"]
        assert got == want


def test_exact_projector_matches(oracle, reference):
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", "htstack_t4.txt"))
    for P in (G, H):
        assert oracle.exact_projector(P, True, [], 1.0) == reference.exact_projector(P, True, [], 1.0)
    # HTstack.circ:10 — P(0) = 0.9786 for 4 T gates
    num = oracle.exact_projector(G, True, [], 1.0)
    den = oracle.exact_projector(H, True, [], 1.0)
    assert abs(num / den - 0.97855339) < 1e-7
