"""Pin the CPU oracle (oracle/packed_oracle.c) against the reference's own known-answer
vectors (tests/units/tests-c/*.txt and tests-python/Unpack*.txt of the reference, converted
to packed form by tests/golden/make_fixtures.py).  Comparison conventions are the
reference's: tests/units/stabtests.c:84-86 (exponential sum), :214-233 (shrink: full state),
:371-376 (inner product), :509-545 (measurePauli: J on the k x k block only)."""
import numpy as np

from util import load, states, unpack, epm_equal, epm_value


def test_exponential_sum_kats(oracle):
    d = load("kat_exponential_sum.npz")
    st = states(d["states"])
    assert len(st) == 274
    zeros = 0
    for s, want in zip(st, d["epm"]):
        got = oracle.exponential_sum(s)
        # stabtests.c:84-86 compares eps, p, and m mod 8 unconditionally
        assert got[0] == want[0] and got[1] == want[1] and (got[2] - want[2]) % 8 == 0
        zeros += got[0] == 0
    assert zeros == 159          # 159 of the 274 cases are exact zeros


def test_shrink_kats(oracle):
    d = load("kat_shrink.npz")
    sin, sout = states(d["states_in"]), states(d["states_out"])
    assert len(sin) == 100
    for s, want, xi, alpha, status in zip(sin, sout, d["xi"], d["alpha"], d["status"]):
        got = oracle.shrink(s, int(xi), int(alpha), 0)
        assert got == status
        assert s.key(full=True) == want.key(full=True)


def test_measure_pauli_kats(oracle):
    d = load("kat_measure_pauli.npz")
    sin, sout = states(d["states_in"]), states(d["states_out"])
    assert len(sin) == 30
    for s, want, m, z, x, res in zip(sin, sout, d["m"], d["zeta"], d["xi"], d["result"]):
        got = oracle.measure_pauli(s, int(m), int(z), int(x))
        assert abs(got - res) < 1e-4
        assert s.key(full=False) == want.key(full=False)


def test_extend_kats(oracle):
    d = load("kat_extend.npz")
    sin, sout = states(d["states_in"]), states(d["states_out"])
    assert len(sin) == 10
    for s, want, x in zip(sin, sout, d["xi"]):
        oracle.extend(s, int(x))
        assert s.k == want.k
        assert tuple(s.G[:s.n]) == tuple(want.G[:s.n])
        assert tuple(s.Gbar[:s.n]) == tuple(want.Gbar[:s.n])


def test_inner_product_kat(oracle):
    d = load("kat_inner_product.npz")
    a, b = states(d["a"]), states(d["b"])
    for s1, s2, want in zip(a, b, d["epm"]):
        assert tuple(want) == (1, -40, 4)
        assert epm_equal(oracle.inner_product(s1, s2), tuple(want))


def test_unpack_kats_and_bruteforce_inner_products(oracle):
    """The authors' MATLAB state vectors pin the meaning of (n,k,h,G,Q,D,J); with that pinned,
    every pair of those states gives a brute-force inner product to check innerProductExact."""
    d = load("kat_unpack.npz")
    st = states(d["states"])
    vecs = []
    for s, amp in zip(st, d["amplitudes"]):
        v = unpack(s)
        # the MATLAB vectors index basis states with qubit 0 as the least significant bit
        assert np.allclose(v, amp[: 1 << s.n], atol=2e-6)
        vecs.append(v)
    checked = 0
    for i, a in enumerate(st):
        for j, b in enumerate(st):
            if a.n != b.n:
                continue
            got = epm_value(oracle.inner_product(a, b))          # <b|a>
            want = np.vdot(vecs[j], vecs[i])
            assert abs(got - want) < 1e-9
            checked += 1
    assert checked > 100
