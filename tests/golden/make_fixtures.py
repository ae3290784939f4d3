#!/usr/bin/env python
"""Generate tests/golden/ fixtures.  Runs ONLY in the build container, where the
reference tree (/root/reference) and the compiled reference (oracle/_ref) exist.
The GPU box has neither: tests read the committed outputs of this script.

    python tests/golden/make_fixtures.py kats        # reference KAT files -> packed .npz
    python tests/golden/make_fixtures.py streams     # back-end instruction streams (main.py file=)
    python tests/golden/make_fixtures.py refpairs    # pairs + (eps,p,m) from the compiled reference
    python tests/golden/make_fixtures.py samples     # theta / projected theta / per-pair epm / values
    python tests/golden/make_fixtures.py fidelity    # decompose()'s -fidelity loop through the compiled reference
    python tests/golden/make_fixtures.py all

What is read from the reference (data, never source code):
  tests/units/tests-c/{exponentialSum,shrink,measurePauli,extend}Tests.txt,
  tests/units/tests-c/innerProductTests-backup.txt   (formats: tests/units/stabtests.c:38-580)
  tests/units/tests-python/{Unpack,UnpackRevised}.txt (format: tests/units/unittests.py:240-269)
  circuits/*.circ via the unmodified front end (main.py / libcirc.probability, `file=` mode,
  libcirc/probability.py:275-280)
"""
import io
import json
import os
import sys
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("BG_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)

from oracle.oracle import (State, Projector, Reference, Oracle, states_to_numpy, STATE_DTYPE)  # noqa: E402


# ----------------------------------------------------------------------------- KATs
def _lines(path):
    with open(path) as f:
        return [ln.strip() for ln in f if ln.strip() != ""]


class _Reader:
    def __init__(self, path):
        self.l = _lines(path)
        self.i = 0

    def more(self):
        return self.i < len(self.l)

    def int(self):
        v = int(float(self.l[self.i]))
        self.i += 1
        return v

    def float(self):
        v = float(self.l[self.i])
        self.i += 1
        return v

    def arr(self):
        v = [int(float(x)) for x in self.l[self.i].split(",") if x != ""]
        self.i += 1
        return v


def _pack_vec(bits):
    v = 0
    for i, b in enumerate(bits):
        if b:
            v |= 1 << i
    return v


def _pack_mat(flat, n):
    return [_pack_vec(flat[r * n:(r + 1) * n]) for r in range(n)]


def _state(n, k, Q=0, h=None, D=None, G=None, Gbar=None, J=None):
    s = State()
    s.n, s.k, s.Q = n, k, Q
    ident = [1 << i for i in range(n)]
    if h is not None:
        s.h = _pack_vec(h)
    if D is not None:
        # setD semantics (stabilizer.c:59-62)
        s.D1 = _pack_vec([(d // 2) % 2 for d in D])
        s.D2 = _pack_vec([(d // 4) % 2 for d in D])
    for i, r in enumerate(_pack_mat(G, n) if G is not None else ident):
        s.G[i] = r
    for i, r in enumerate(_pack_mat(Gbar, n) if Gbar is not None else ident):
        s.Gbar[i] = r
    if J is not None:
        for i, r in enumerate(_pack_mat([1 if x > 0 else 0 for x in J], n)):
            s.J[i] = r
    return s


def kats():
    tc = os.path.join(REF, "tests/units/tests-c")
    # exponentialSum: n k Q D J | eps p m        (stabtests.c:52-79)
    r = _Reader(os.path.join(tc, "exponentialSumTests.txt"))
    st, out = [], []
    while r.more():
        n, k, Q = r.int(), r.int(), r.int()
        D, J = r.arr(), r.arr()
        st.append(_state(n, k, Q=Q, D=D, J=J))
        out.append([r.int(), r.int(), r.int()])
    np.savez_compressed(os.path.join(HERE, "kat_exponential_sum.npz"),
                        states=states_to_numpy(st), epm=np.array(out, dtype=np.int32))
    print("exponentialSum", len(st))

    # shrink: n k Q alpha h D xi G Gbar J | status k Q h D G Gbar J    (stabtests.c:119-212)
    r = _Reader(os.path.join(tc, "shrinkTests.txt"))
    sin, sout, xi, alpha, status = [], [], [], [], []
    while r.more():
        n, k, Q, al = r.int(), r.int(), r.int(), r.int()
        h, D, x, G, Gb, J = r.arr(), r.arr(), r.arr(), r.arr(), r.arr(), r.arr()
        sin.append(_state(n, k, Q, h, D, G, Gb, J))
        xi.append(_pack_vec(x)); alpha.append(al)
        status.append(r.int())
        ok, oQ = r.int(), r.int()
        oh, oD, oG, oGb, oJ = r.arr(), r.arr(), r.arr(), r.arr(), r.arr()
        sout.append(_state(n, ok, oQ, oh, oD + [0] * (n - len(oD)), oG, oGb, oJ))
    np.savez_compressed(os.path.join(HERE, "kat_shrink.npz"), states_in=states_to_numpy(sin),
                        states_out=states_to_numpy(sout), xi=np.array(xi, dtype=np.uint64),
                        alpha=np.array(alpha, dtype=np.int32), status=np.array(status, dtype=np.int32))
    print("shrink", len(sin))

    # measurePauli: n k Q m h D zeta xi G Gbar J | result k Q h D G Gbar J   (stabtests.c:415-507)
    r = _Reader(os.path.join(tc, "measurePauliTests.txt"))
    sin, sout, ms, zs, xs, res = [], [], [], [], [], []
    while r.more():
        n, k, Q, m = r.int(), r.int(), r.int(), r.int()
        h, D, ze, x, G, Gb, J = r.arr(), r.arr(), r.arr(), r.arr(), r.arr(), r.arr(), r.arr()
        sin.append(_state(n, k, Q, h, D + [0] * (n - len(D)), G, Gb, J))
        ms.append(m); zs.append(_pack_vec(ze)); xs.append(_pack_vec(x))
        res.append(r.float())
        ok, oQ = r.int(), r.int()
        oh, oD, oG, oGb, oJ = r.arr(), r.arr(), r.arr(), r.arr(), r.arr()
        sout.append(_state(n, ok, oQ, oh, oD + [0] * (n - len(oD)), oG, oGb, oJ))
    np.savez_compressed(os.path.join(HERE, "kat_measure_pauli.npz"), states_in=states_to_numpy(sin),
                        states_out=states_to_numpy(sout), m=np.array(ms, dtype=np.int32),
                        zeta=np.array(zs, dtype=np.uint64), xi=np.array(xs, dtype=np.uint64),
                        result=np.array(res))
    print("measurePauli", len(sin))

    # extend: n k xi G Gbar | k G Gbar
    r = _Reader(os.path.join(tc, "extendTests.txt"))
    sin, sout, xs = [], [], []
    while r.more():
        n, k = r.int(), r.int()
        x, G, Gb = r.arr(), r.arr(), r.arr()
        sin.append(_state(n, k, G=G, Gbar=Gb)); xs.append(_pack_vec(x))
        ok = r.int()
        oG, oGb = r.arr(), r.arr()
        sout.append(_state(n, ok, G=oG, Gbar=oGb))
    np.savez_compressed(os.path.join(HERE, "kat_extend.npz"), states_in=states_to_numpy(sin),
                        states_out=states_to_numpy(sout), xi=np.array(xs, dtype=np.uint64))
    print("extend", len(sin))

    # innerProduct (the one surviving case): state1, state2 | eps p m     (stabtests.c:272-364)
    r = _Reader(os.path.join(tc, "innerProductTests-backup.txt"))
    a, b, out = [], [], []
    while r.more():
        two = []
        for _ in range(2):
            n, k, Q = r.int(), r.int(), r.int()
            h, D, G, Gb, J = r.arr(), r.arr(), r.arr(), r.arr(), r.arr()
            two.append(_state(n, k, Q, h, D + [0] * (n - len(D)), G, Gb, J))
        a.append(two[0]); b.append(two[1])
        out.append([r.int(), r.int(), r.int()])
    np.savez_compressed(os.path.join(HERE, "kat_inner_product.npz"), a=states_to_numpy(a),
                        b=states_to_numpy(b), epm=np.array(out, dtype=np.int32))
    print("innerProduct", len(a))

    # Unpack: full state vectors at n = 10 (authors' MATLAB output)    (unittests.py:240-269)
    st, amps = [], []
    for name in ("Unpack.txt", "UnpackRevised.txt"):
        for tcase in json.load(open(os.path.join(REF, "tests/units/tests-python", name))):
            psi = tcase["output"]["psi_out"]
            n, k = psi["n"], psi["k"]
            Din = np.atleast_1d(np.array(psi["D"])).astype(int).tolist()
            D = Din + [0] * (n - len(Din))
            J = np.zeros((n, n), dtype=int)
            if k > 0:
                J[:k, :k] = np.atleast_2d(np.array(psi["J"])).astype(int)[:k, :k]
            s = _state(n, k, psi["Q"], np.atleast_1d(np.array(psi["h"])).astype(int).tolist(), D,
                       np.array(psi["G"]).flatten().tolist(),
                       np.array(psi["Gbar"]).flatten().tolist(), J.flatten().tolist())
            st.append(s)
            amp = np.zeros(1024, dtype=complex)       # n <= 10; entries beyond 2^n stay 0
            v = (np.atleast_1d(np.array(tcase["unpackReals"]["root"], dtype=float))
                 + 1j * np.atleast_1d(np.array(tcase["unpackImaginaries"]["root"], dtype=float)))
            amp[:len(v)] = v
            amps.append(amp)
    np.savez_compressed(os.path.join(HERE, "kat_unpack.npz"), states=states_to_numpy(st),
                        amplitudes=np.array(amps).astype(np.complex64))
    print("unpack", len(st))


# ----------------------------------------------------------------------------- instruction streams
def _front_end():
    os.chdir(REF)                      # the .circ importer resolves paths from the reference root
    sys.path.insert(0, REF)
    import warnings
    warnings.simplefilter("ignore")
    from libcirc.probability import probability
    from libcirc.compile.compilecirc import compileCircuit
    return probability, compileCircuit


def hidden_shift_circuit(n=40, toff=1, randcliff=200):
    """The random bent-function hidden-shift circuit of circuits/hiddenshift.py:46-144
    (Bravyi-Gosset appendix F), restated for python 3 with the same numpy call order so
    that seed 0 gives the reference's circuit."""
    np.random.seed(0)
    half = int(np.ceil(n / 2))
    n = 2 * half
    s = np.random.randint(0, 2, size=n)
    Og = []

    def distinct(*taken):
        while True:
            loc = np.random.randint(0, half)
            if loc not in taken:
                return loc

    def cliffords():
        for _ in range(randcliff):
            line = ["_"] * half
            if np.random.randint(0, 2) == 1:
                line[np.random.randint(0, half)] = "Z"
            else:
                a = np.random.randint(0, half)
                b = distinct(a)
                line[a], line[b] = "C", "Z"
            Og.append("".join(line))

    cliffords()
    for _ in range(toff):
        a = np.random.randint(0, half)
        b = distinct(a)
        c = distinct(a, b)
        line = ["_"] * half
        line[a], line[b], line[c] = "C", "C", "Z"
        Og.append("".join(line))
        cliffords()

    def single(g, i):
        return "_" * i + g + "_" * (n - 1 - i) + "\n"

    had = "".join(single("H", i) for i in range(n))
    xs = "".join(single("X", i) for i in range(n) if s[i] == 1)
    czs = "".join(i * "_" + "C" + "_" * (half - 1) + "Z" + "_" * (half - i - 1) + "\n" for i in range(half))
    circ = "import circuits/reference.circ\nmain:\n"
    circ += had + xs + "".join(half * "_" + g + "\n" for g in Og) + czs + xs
    circ += had + "".join(g + half * "_" + "\n" for g in Og) + czs + had
    return circ, s


def _stream(probability, compiled, measure, config, path):
    cfg = dict(config)
    cfg["file"] = path
    cfg["quiet"] = True
    if os.path.exists(path):
        os.remove(path)
    with contextlib.redirect_stdout(io.StringIO()):
        probability(compiled, dict(measure), config=cfg)
    return os.path.exists(path)


def streams():
    probability, compileCircuit = _front_end()
    out = os.path.join(HERE, "streams")
    os.makedirs(out, exist_ok=True)
    meta = {}

    # config 2: HTstack.circ output 0, samples=1024 -forceSample  (t=4, chi=4)
    np.random.seed(1)
    comp = compileCircuit(fname="circuits/HTstack.circ")
    _stream(probability, comp, {0: 0}, {"samples": 1024, "forceSample": True}, os.path.join(out, "htstack_t4.txt"))
    meta["htstack_t4"] = {"expect_probability": 0.97855339, "source": "circuits/HTstack.circ:10"}

    # config 1: toffoli.circ, first step of sampling MMM: P(qubit0 = 0)   (t=16, exact-norm path)
    np.random.seed(2)
    comp = compileCircuit(fname="circuits/toffoli.circ")
    _stream(probability, comp, {0: 0}, {"samples": 2000}, os.path.join(out, "toffoli_q0.txt"))
    np.random.seed(3)
    _stream(probability, comp, {0: 1, 1: 1, 2: 1}, {"samples": 2000}, os.path.join(out, "toffoli_111.txt"))
    meta["toffoli_111"] = {"expect_probability": 1.0, "source": "circuits/toffoli.circ:14-19"}

    # controlledT: 0.8536 (circuits/controlledT.circ:7)
    try:
        np.random.seed(4)
        comp = compileCircuit(fname="circuits/controlledT.circ")
        first = comp.splitlines()[0]
        nq = len(first)
        _stream(probability, comp, {nq - 1: 0} if nq > 1 else {0: 0}, {"samples": 2000},
                os.path.join(out, "controlledT.txt"))
    except Exception as e:     # noqa
        print("controlledT skipped:", e)

    # configs 3 and 4: hidden shift n=40, t=16 (-t 2, exact) and t=40 (-t 5 -k 9)
    for tag, toff, cfg in (("hs_t16", 2, {"samples": 16384, "exact": True, "forceSample": True}),
                           ("hs_t40_k9", 5, {"samples": 65536, "k": 9, "exact": False, "forceSample": True})):
        circ, s = hidden_shift_circuit(40, toff)
        comp = compileCircuit(raw=circ)
        written = []
        for bit in range(40):
            path = os.path.join(out, "%s_bit%d.txt" % (tag, bit))
            if _stream(probability, comp, {bit: 1}, cfg, path):
                written.append(bit)
            if len(written) >= 2:
                break
        meta[tag] = {"shift": "".join(map(str, s.tolist())), "bits_reaching_backend": written,
                     "circuit_qubits": len(comp.splitlines()[0])}
        print(tag, meta[tag])

    # config 5: phaseEstimation.circ, first sampled qubit
    try:
        np.random.seed(5)
        comp = compileCircuit(fname="circuits/phaseEstimation.circ")
        _stream(probability, comp, {0: 0}, {"samples": 4096, "k": 8, "exact": False, "forceSample": True},
                os.path.join(out, "phase_estimation_q0.txt"))
    except Exception as e:     # noqa
        print("phaseEstimation skipped:", e)

    # config 5, whole: sampleQubits(phaseEstimation.circ, "MMM_") makes one probability() call per sampled qubit, each
    # conditioned on the outcomes before it (libcirc/sample.py:32-83).  sampleQubits disables file= mode, so the calls it
    # would make are written here one by one for EVERY outcome prefix (1 + 2 + 4 streams); bench.py replays the chain
    # against the back end exactly as recursiveSample does (P0 = P / Psofar, then a draw).  When the measurement dict is
    # not empty, sampleQubits first asks for the probability of the constraints with exact=True (sample.py:37-42): those
    # are the *_given streams.
    try:
        comp = compileCircuit(fname="circuits/phaseEstimation.circ")
        cfg5 = {"samples": 16384, "k": 8, "exact": False, "forceSample": True}
        chain = {}
        for nq in range(3):
            for prefix in range(1 << nq):
                bits = [(prefix >> (nq - 1 - j)) & 1 for j in range(nq)]
                measure = {j: b for j, b in enumerate(bits)}
                measure[nq] = 0
                name = "phase_estimation_chain_%s0.txt" % "".join(map(str, bits))
                np.random.seed(50 + 8 * nq + prefix)
                if _stream(probability, comp, measure, cfg5, os.path.join(out, name)):
                    # probability() returns 2^(v-u) * numerator / denominator (probability.py:320-326): u, v are how many
                    # generators truncate() removed from G and H (probability.py:93-94) — host-side, not in the stream
                    from libcirc.compile import projectors as _pj
                    Gp, Hp, n_, t_ = _pj.projectors(comp, dict(measure), verbose=False, x=None, y=None)
                    u_ = _pj.truncate(n_, Gp)[1]
                    v_ = _pj.truncate(n_, Hp)[1]
                    chain["".join(map(str, bits)) + "0"] = {"stream": name, "v_minus_u": int(v_ - u_)}
        meta["phase_estimation_chain"] = {"streams": chain, "qubits": [0, 1, 2],
                                          "note": "key = outcomes so far + the 0 being asked for; a missing key means the "
                                                  "front end answered that call itself (probability.py:111-138)"}
    except Exception as e:     # noqa
        print("phaseEstimation chain skipped:", e)

    json.dump(meta, open(os.path.join(out, "meta.json"), "w"), indent=1, sort_keys=True)
    os.chdir(ROOT)


def parse_stream(path):
    """13 scalars + 2 projectors (libcirc/probability.c:74-127, utils/comms.c:9-36)."""
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
            xs.append(x); zs.append(z)
        projs.append(Projector.make(nq, ph, xs, zs))
    return cfg, projs[0], projs[1]


# ----------------------------------------------------------------------------- reference-run fixtures
def refpairs():
    """Random pairs through the COMPILED REFERENCE's innerProductExact."""
    ref = Reference()
    a, b, out = [], [], []
    seed = 1000

    def push(s1, s2):
        a.append(s1.copy()); b.append(s2.copy()); out.append(ref.inner_product(s1, s2))

    rng = np.random.RandomState(7)
    for n in (1, 2, 3, 4, 5, 7, 8, 12, 16, 24, 31, 32, 33, 40, 48, 63, 64):
        for rep in range(6):
            seed += 1
            ref.srand(seed)
            s1, s2 = ref.random_state(n), ref.random_state(n)
            push(s1, s2)
            # lower-dimensional partners: project s2 by random Paulis, and decomposition terms
            s3 = s2.copy()
            for _ in range(int(rng.randint(1, n + 1))):
                m = int(rng.randint(0, 4))
                z = int(rng.randint(0, 2 ** 62)) & ((1 << n) - 1)
                x = int(rng.randint(0, 2 ** 62)) & ((1 << n) - 1) if rng.randint(0, 2) else 0
                if (bin(x & z).count("1") + m) % 2:       # keep the Pauli Hermitian: i^m with m = |x&z| mod 2
                    m = (m + 1) % 4
                if ref.measure_pauli(s3, m, z, x) == 0:
                    break
            push(s1, s3)
            push(s3, s1)
            if n >= 2:
                push(s1, ref.prepH(int(rng.randint(0, 2 ** ((n + 1) // 2))), n))
                L = [int(rng.randint(0, 2 ** 62)) & ((1 << n) - 1) for _ in range(min(n, 5))]
                push(s3, ref.prepL(int(rng.randint(0, 2 ** len(L))), n, L))
    np.savez_compressed(os.path.join(HERE, "ref_pairs.npz"), a=states_to_numpy(a), b=states_to_numpy(b),
                        epm=np.array(out, dtype=np.int32))
    print("refpairs", len(a), "zeros:", sum(1 for e in out if e[0] == 0))


def samples():
    """theta (libc rand, reference generator) -> reference measurePauli projection -> reference
    per-pair (eps,p,m) and the reference's fp64 sample value, for the BASELINE configs."""
    ref = Reference()
    sdir = os.path.join(HERE, "streams")
    jobs = [("htstack_t4", 64), ("hs_t16_bit", 6), ("hs_t40_k9_bit", 3), ("phase_estimation_q0", 2)]
    for prefix, count in jobs:
        files = sorted(f for f in os.listdir(sdir) if f.startswith(prefix) and f.endswith(".txt"))
        if not files:
            print("no stream for", prefix); continue
        path = os.path.join(sdir, files[0])
        cfg, G, H = parse_stream(path)
        t, exact, k = cfg["t"], cfg["exact"], cfg["k"]
        rs = np.random.RandomState(11)
        L = [] if exact else [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) for _ in range(k)]
        rec = dict(theta=[], projected=[], epm=[], value=[], alive=[], total=[], projfactor=[], which=[])
        for which, P in enumerate((G, H)):
            for c in range(count):
                ref.srand(5000 + 17 * c + which)
                th = ref.random_state(t)
                r = ref.sample_from_theta(th, P, exact, L)
                rec["theta"].append(th); rec["projected"].append(r["theta"]); rec["epm"].append(r["epm"])
                rec["value"].append(r["value"]); rec["alive"].append(r["alive"])
                rec["total"].append(r["total"]); rec["projfactor"].append(r["projfactor"]); rec["which"].append(which)
        name = os.path.splitext(files[0])[0]
        np.savez_compressed(os.path.join(HERE, "ref_samples_%s.npz" % name),
                            stream=np.array(files[0]), L=np.array(L, dtype=np.uint64),
                            theta=states_to_numpy(rec["theta"]), projected=states_to_numpy(rec["projected"]),
                            epm=np.array(rec["epm"], dtype=np.int32), value=np.array(rec["value"]),
                            alive=np.array(rec["alive"], dtype=np.int32), total=np.array(rec["total"]),
                            projfactor=np.array(rec["projfactor"]), which=np.array(rec["which"], dtype=np.int32))
        print(name, "samples", len(rec["value"]), "alive", int(np.sum(rec["alive"])))


def fidelity():
    """decompose() with fidelity = 1 (libcirc/probability.c:361-410) through the COMPILED reference back end
    (oracle/_ref/mpibackend_ref): for an empty projector H the printed denominator is norm^2 = 2^k Z(L)
    (innerprod.c:47; G is empty too, so nothing is sampled), and the chatter line carries delta = 1 - <H^t|L>.  L comes from libc rand() in its
    default state (decompose runs before srand, probability.c:153,182), so it is reproducible: the fixture
    stores L as BitMatrixSetRandom (utils/matrix.c:301-306) lays it out."""
    import ctypes
    import re
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "mpibackend_ref")
    libc = ctypes.CDLL(None)
    cases = []
    for (t, k) in [(4, 2), (12, 5), (16, 8), (33, 7), (40, 9), (40, 12), (60, 10), (64, 11)]:
        tok = [0, 0, 0, 1, 1, t, k, 0, 1e-05, 1, 0, 1, 1]            # fidelity = 1, forceL, forceSample
        tok += [0, 0, 0, 0]                                           # G = H = empty -> norm^2, no sampling at all
        out = subprocess.run([exe, "stdin"], input="\n".join(str(v) for v in tok) + "\n", capture_output=True,
                             text=True, check=True).stdout.split("\n")
        lines = [ln for ln in out if ln.strip()]
        delta = [float(re.search(r"delta = 1 - <H\^t\|L>: ([-0-9.eE]+)", ln).group(1)) for ln in lines if "delta" in ln]
        libc.srand(1)
        rows = [0] * k
        for b in range((k * t + 7) // 8):
            byte = libc.rand() % 256
            for j in range(8):
                loc = 8 * b + j
                if loc < k * t and (byte >> (7 - j)) & 1:
                    rows[loc // t] |= 1 << (loc % t)
        cases.append({"t": t, "k": k, "L_rows": [str(r) for r in rows], "delta_printed": delta[-1],
                      "numerator": lines[-2], "denominator": lines[-1]})
        print("fidelity t=%d k=%d norm^2=%s delta=%g" % (t, k, lines[-1], delta[-1]))
    with open(os.path.join(HERE, "ref_fidelity.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_fixtures.py fidelity", "cases": cases}, f, indent=1)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("kats", "all"):
        kats()
    if what in ("streams", "all"):
        streams()
    if what in ("refpairs", "all"):
        refpairs()
    if what in ("samples", "all"):
        samples()
    if what in ("fidelity", "all"):
        fidelity()
