"""Shared helpers for the test-suite (fixtures, brute-force state vectors)."""
import os

import numpy as np

from oracle.oracle import State, Projector, state_from_numpy, states_to_numpy, epm_equal  # noqa: F401

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def states(arr):
    return [state_from_numpy(arr[i]) for i in range(len(arr))]


def unpack(s):
    """Brute-force state vector of a packed state, index = sum_q x_q 2^q.
    |K,q> = 2^{-k/2} sum_{z in F_2^k} w^{Q + D.z + 4 sum_{a<b} J_ab z_a z_b} |h + z G>
    (Bravyi-Gosset eq. 43-46; reference twin: libcirc/stabilizer/stabilizer.py:54-75)."""
    n, k = s.n, s.k
    assert n <= 14
    psi = np.zeros(1 << n, dtype=complex)
    w = np.exp(1j * np.pi / 4)
    D = [2 * ((s.D1 >> a) & 1) + 4 * ((s.D2 >> a) & 1) for a in range(k)]
    for z in range(1 << k):
        x = s.h
        ph = s.Q
        for a in range(k):
            if (z >> a) & 1:
                x ^= s.G[a]
                ph += D[a]
                ph += 4 * bin(s.J[a] & z & ((1 << a) - 1)).count("1")
        psi[x] = w ** (ph % 8)
    return psi / 2 ** (k / 2)


def reverse_index(vec, n):
    """Reference/MATLAB ordering puts qubit 0 at the MSB of the index."""
    out = np.zeros_like(vec)
    for i in range(1 << n):
        j = int(format(i, "0%db" % n)[::-1], 2) if n > 0 else 0
        out[j] = vec[i]
    return out


def epm_value(epm):
    eps, p, m = (int(v) for v in epm)
    return eps * 2 ** (p / 2) * np.exp(1j * np.pi * (m % 8) / 4)


def parse_stream(path):
    """Back-end instruction stream: 13 scalars + 2 projectors
    (libcirc/probability.c:74-127, libcirc/utils/comms.c:9-36)."""
    tok = open(path).read().split()
    it = iter(tok)
    names = ["quiet", "verbose", "noapprox", "samples", "bins", "t", "k", "exact", "fidbound",
             "fidelity", "rank", "forceL", "forceSample"]
    cfg = {}
    for nme in names:
        v = next(it)
        cfg[nme] = float(v) if nme == "fidbound" else int(float(v))
    projs = []
    for _ in range(2):
        ns, nq = int(next(it)), int(next(it))
        ph, xs, zs = [], [], []
        for _i in range(ns):
            ph.append(int(next(it)) % 4)
            x = z = 0
            for q in range(nq):
                if int(next(it)):
                    x |= 1 << q
                if int(next(it)):
                    z |= 1 << q
            xs.append(x)
            zs.append(z)
        projs.append(Projector.make(nq, ph, xs, zs))
    return cfg, projs[0], projs[1]


def exact_sample_value(epm, t, projfactor):
    """2^t |projfactor * sum_i eps_i 2^{p_i/2} w^{m_i}|^2 evaluated EXACTLY in Z[sqrt2] with Python
    integers (then rounded once), from per-pair (eps, p, m) triples.  Used to judge fp64 results
    whose own summation error is amplified by cancellation."""
    from decimal import Decimal, getcontext
    getcontext().prec = 80
    sh = t // 2 + 1
    a = [0, 0, 0, 0]

    def add(e, mag):
        e %= 8
        a[e & 3] += -mag if e & 4 else mag

    for eps, p, m in epm:
        eps, p, m = int(eps), int(p), int(m)
        if not eps:
            continue
        f = p // 2                      # floor
        mag = 1 << (sh + f)
        if p % 2 == 0:
            add(m, mag)
        else:
            add(m + 1, mag)
            add(m - 1, mag)
    X = sum(v * v for v in a)
    Y = a[0] * a[1] + a[1] * a[2] + a[2] * a[3] - a[0] * a[3]
    npf = int(round(-2 * np.log2(projfactor))) if projfactor > 0 else 0
    val = (Decimal(X) + Decimal(2).sqrt() * Decimal(Y)) * (Decimal(2) ** (t - npf - 2 * sh))
    mags = sum(2.0 ** (int(p_) / 2) for e_, p_, m_ in epm if int(e_))      # sum_i |term_i| (before cancellation)
    scale = mags * mags * 2.0 ** (t - npf)
    return float(val), scale
