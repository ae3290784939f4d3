"""ctypes bindings for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

`liboracle.so`          packed-row restatement of the reference algorithm (oracle/packed_oracle.c)
`_ref/libcircref.so`    the unmodified reference C sources + layout adapters (oracle/ref_driver.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_T = 64
MAX_STABS = 128


class State(C.Structure):
    """bg_state (include/bgnorm.h)."""
    _fields_ = [("n", C.c_int32), ("k", C.c_int32), ("Q", C.c_int32), ("reserved", C.c_int32),
                ("h", C.c_uint64), ("D1", C.c_uint64), ("D2", C.c_uint64),
                ("G", C.c_uint64 * MAX_T), ("Gbar", C.c_uint64 * MAX_T), ("J", C.c_uint64 * MAX_T)]

    def copy(self):
        out = State()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(State))
        return out

    def key(self, full=True):
        """Comparable tuple; `full` compares whole n x n matrices (incl. junk outside k x k)."""
        n, k = self.n, self.k
        mk = (1 << k) - 1
        if full:
            return (n, k, self.Q % 8, self.h, self.D1 & mk, self.D2 & mk,
                    tuple(self.G[:n]), tuple(self.Gbar[:n]), tuple(self.J[:n]))
        return (n, k, self.Q % 8, self.h, self.D1 & mk, self.D2 & mk,
                tuple(self.G[:n]), tuple(self.Gbar[:n]), tuple(j & mk for j in self.J[:k]))


class Projector(C.Structure):
    """bg_projector (include/bgnorm.h)."""
    _fields_ = [("nstabs", C.c_int32), ("nqubits", C.c_int32),
                ("phase", C.c_uint8 * MAX_STABS),
                ("xs", C.c_uint64 * MAX_STABS), ("zs", C.c_uint64 * MAX_STABS)]

    @staticmethod
    def make(nqubits, phases, xs, zs):
        p = Projector()
        p.nstabs = len(phases)
        p.nqubits = nqubits
        for i, (ph, x, z) in enumerate(zip(phases, xs, zs)):
            p.phase[i] = int(ph)
            p.xs[i] = int(x)
            p.zs[i] = int(z)
        return p


STATE_DTYPE = np.dtype([("n", "<i4"), ("k", "<i4"), ("Q", "<i4"), ("reserved", "<i4"),
                        ("h", "<u8"), ("D1", "<u8"), ("D2", "<u8"),
                        ("G", "<u8", (MAX_T,)), ("Gbar", "<u8", (MAX_T,)), ("J", "<u8", (MAX_T,))])
assert STATE_DTYPE.itemsize == C.sizeof(State)


def states_to_numpy(states):
    arr = np.zeros(len(states), dtype=STATE_DTYPE)
    for i, s in enumerate(states):
        C.memmove(arr[i:i + 1].ctypes.data, C.byref(s), C.sizeof(State))
    return arr


def state_from_numpy(rec):
    s = State()
    buf = np.ascontiguousarray(rec).reshape(1)
    C.memmove(C.byref(s), buf.ctypes.data, C.sizeof(State))
    return s


def u64_array(vals):
    arr = (C.c_uint64 * max(1, len(vals)))()
    for i, v in enumerate(vals):
        arr[i] = int(v)
    return arr


def build(force=False):
    """Compile liboracle.so (always possible) and, when the reference tree is present,
    oracle/_ref (the reference compiled where it lies)."""
    if force or not os.path.exists(os.path.join(HERE, "liboracle.so")):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    ref_root = os.environ.get("BG_REFERENCE_ROOT", "/root/reference")
    if os.path.isdir(os.path.join(ref_root, "libcirc")):
        if force or not os.path.exists(os.path.join(HERE, "_ref", "libcircref.so")):
            subprocess.check_call(["make", "-C", HERE, "ref", "REF=" + ref_root], stdout=subprocess.DEVNULL)


_P = C.POINTER
_i32p = _P(C.c_int32)
_dblp = _P(C.c_double)


def _common_sigs(lib, pre):
    f = getattr(lib, pre + "inner_product")
    f.argtypes = [_P(State), _P(State), _i32p, _i32p, _i32p]
    f.restype = None
    f = getattr(lib, pre + "exponential_sum")
    f.argtypes = [_P(State), _i32p, _i32p, _i32p]
    f.restype = None
    f = getattr(lib, pre + "shrink")
    f.argtypes = [_P(State), C.c_uint64, C.c_int, C.c_int]
    f.restype = C.c_int
    f = getattr(lib, pre + "extend")
    f.argtypes = [_P(State), C.c_uint64]
    f.restype = None
    f = getattr(lib, pre + "measure_pauli")
    f.argtypes = [_P(State), C.c_int, C.c_uint64, C.c_uint64]
    f.restype = C.c_double
    f = getattr(lib, pre + "prepH")
    f.argtypes = [C.c_int, C.c_int, _P(State)]
    f.restype = None
    f = getattr(lib, pre + "prepL")
    f.argtypes = [C.c_int, C.c_int, C.c_int, _P(C.c_uint64), _P(State)]
    f.restype = None
    f = getattr(lib, pre + "evalW")
    f.argtypes = [C.c_int, C.c_int, C.c_int, _dblp, _dblp]
    f.restype = None
    f = getattr(lib, pre + "sample_from_theta")
    f.argtypes = [_P(State), _P(Projector), C.c_int, C.c_int, _P(C.c_uint64), _i32p,
                  _P(C.c_int), _dblp, _dblp, _dblp]
    f.restype = C.c_double
    f = getattr(lib, pre + "exact_projector_work")
    f.argtypes = [C.c_int, _P(Projector), C.c_int, C.c_int, _P(C.c_uint64), _dblp, _dblp]
    f.restype = None
    f = getattr(lib, pre + "exact_projector")
    f.argtypes = [_P(Projector), C.c_int, C.c_int, _P(C.c_uint64), C.c_double]
    f.restype = C.c_double
    f = getattr(lib, pre + "sizeof_state")
    f.restype = C.c_size_t
    assert f() == C.sizeof(State)
    f = getattr(lib, pre + "sizeof_projector")
    f.restype = C.c_size_t
    assert f() == C.sizeof(Projector)


class _Api:
    """Uniform python face over either library (prefix 'orc_' or 'ref_')."""

    def __init__(self, lib, pre):
        self.lib, self.pre = lib, pre
        _common_sigs(lib, pre)

    def _f(self, name):
        return getattr(self.lib, self.pre + name)

    def inner_product(self, a, b):
        e, p, m = C.c_int32(), C.c_int32(), C.c_int32()
        self._f("inner_product")(C.byref(a), C.byref(b), C.byref(e), C.byref(p), C.byref(m))
        return e.value, p.value, m.value

    def exponential_sum(self, s):
        e, p, m = C.c_int32(), C.c_int32(), C.c_int32()
        self._f("exponential_sum")(C.byref(s), C.byref(e), C.byref(p), C.byref(m))
        return e.value, p.value, m.value

    def shrink(self, s, xi, alpha, lazy=0):
        return self._f("shrink")(C.byref(s), xi, alpha, lazy)

    def extend(self, s, xi):
        self._f("extend")(C.byref(s), xi)

    def measure_pauli(self, s, m, zeta, xi):
        return self._f("measure_pauli")(C.byref(s), m, zeta, xi)

    def prepH(self, i, t):
        s = State()
        self._f("prepH")(i, t, C.byref(s))
        return s

    def prepL(self, i, t, Lrows):
        s = State()
        self._f("prepL")(i, t, len(Lrows), u64_array(Lrows), C.byref(s))
        return s

    def evalW(self, eps, p, m):
        re, im = C.c_double(), C.c_double()
        self._f("evalW")(eps, p, m, C.byref(re), C.byref(im))
        return complex(re.value, im.value)

    def sample_from_theta(self, theta, P, exact, Lrows, want_epm=True):
        """Returns dict(value, alive, total, projfactor, epm[chi,3], theta=projected state)."""
        t = P.nqubits
        k = 0 if exact else len(Lrows)
        chi = (1 << ((t + 1) // 2)) if exact else (1 << k)
        epm = np.zeros((chi, 3), dtype=np.int32)
        alive = C.c_int()
        tre, tim, pf = C.c_double(), C.c_double(), C.c_double()
        th = theta.copy()
        v = self._f("sample_from_theta")(C.byref(th), C.byref(P), int(bool(exact)), k, u64_array(Lrows or []),
                                         epm.ctypes.data_as(_i32p) if want_epm else None,
                                         C.byref(alive), C.byref(tre), C.byref(tim), C.byref(pf))
        return dict(value=v, alive=alive.value, total=complex(tre.value, tim.value),
                    projfactor=pf.value, epm=epm, theta=th)

    def exact_projector_work(self, l, P, exact, Lrows):
        re, im = C.c_double(), C.c_double()
        self._f("exact_projector_work")(l, C.byref(P), int(bool(exact)), 0 if exact else len(Lrows),
                                        u64_array(Lrows or []), C.byref(re), C.byref(im))
        return complex(re.value, im.value)

    def exact_projector(self, P, exact, Lrows, norm=1.0):
        return self._f("exact_projector")(C.byref(P), int(bool(exact)), 0 if exact else len(Lrows),
                                          u64_array(Lrows or []), norm)


class Oracle(_Api):
    """liboracle.so — the packed-row restatement."""

    def __init__(self):
        build()
        lib = C.CDLL(os.path.join(HERE, "liboracle.so"))
        super().__init__(lib, "orc_")
        lib.orc_random_state_libc.argtypes = [C.c_int, _P(State)]
        lib.orc_random_state_philox.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_uint64, _P(State)]
        lib.orc_dimension_cdf.argtypes = [C.c_int, _dblp]
        lib.orc_philox_block.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, _P(C.c_uint32)]
        lib.orc_sampled_sum_philox.argtypes = [_P(Projector), C.c_int, C.c_int, _P(C.c_uint64), C.c_uint64,
                                               C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, _dblp]
        lib.orc_sampled_sum_philox.restype = C.c_double
        lib.orc_sampled_projector_libc.argtypes = [_P(Projector), C.c_int, C.c_int, _P(C.c_uint64),
                                                   C.c_double, C.c_int]
        lib.orc_sampled_projector_libc.restype = C.c_double
        lib.orc_Lbits.argtypes = [C.c_int, C.c_int, _P(C.c_uint64)]
        lib.orc_Lbits.restype = C.c_uint64
        lib.orc_wordops.restype = C.c_uint64
        lib.orc_identity_state.argtypes = [_P(State), C.c_int, C.c_int]
        lib.orc_decompose_ZL.argtypes = [C.c_int, C.c_int, _P(C.c_uint64), _P(C.c_uint64)]
        lib.orc_decompose_ZL.restype = C.c_double
        self.libc = C.CDLL(None)

    def srand(self, seed):
        self.libc.srand(C.c_uint(seed))

    def identity_state(self, n, k):
        s = State()
        self.lib.orc_identity_state(C.byref(s), n, k)
        return s

    def random_state_libc(self, n):
        s = State()
        self.lib.orc_random_state_libc(n, C.byref(s))
        return s

    def random_state_philox(self, n, seed, bin_, sample):
        s = State()
        self.lib.orc_random_state_philox(n, seed, bin_, sample, C.byref(s))
        return s

    def dimension_cdf(self, n):
        out = (C.c_double * (n + 1))()
        self.lib.orc_dimension_cdf(n, out)
        return list(out)

    def philox_block(self, seed, sample, bin_, block):
        out = (C.c_uint32 * 4)()
        self.lib.orc_philox_block(seed, sample, bin_, block, out)
        return list(out)

    def Lbits(self, i, Lrows):
        return self.lib.orc_Lbits(i, len(Lrows), u64_array(Lrows))

    def sampled_sum_philox(self, P, exact, Lrows, seed, bin_, first, stride, count):
        per = np.zeros(count, dtype=np.float64)
        tot = self.lib.orc_sampled_sum_philox(C.byref(P), int(bool(exact)), 0 if exact else len(Lrows),
                                              u64_array(Lrows or []), seed, bin_, first, stride, count,
                                              per.ctypes.data_as(_dblp))
        return tot, per

    def sampled_projector_libc(self, P, exact, Lrows, norm, samples):
        return self.lib.orc_sampled_projector_libc(C.byref(P), int(bool(exact)), 0 if exact else len(Lrows),
                                                   u64_array(Lrows or []), norm, samples)

    def decompose_ZL(self, t, Lrows):
        """(Z(L), weight histogram) of decompose()'s fidelity loop."""
        hist = (C.c_uint64 * 65)()
        z = self.lib.orc_decompose_ZL(t, len(Lrows), u64_array(Lrows), hist)
        return z, [int(v) for v in hist]

    def wordops(self, reset=False):
        v = self.lib.orc_wordops()
        if reset:
            self.lib.orc_wordops_reset()
        return v


class Reference(_Api):
    """oracle/_ref/libcircref.so — the unmodified reference C code behind layout adapters."""

    def __init__(self):
        build()
        path = os.path.join(HERE, "_ref", "libcircref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (reference not compiled; needs /root/reference at build time)")
        lib = C.CDLL(path)
        super().__init__(lib, "ref_")
        lib.ref_srand.argtypes = [C.c_uint]
        lib.ref_random_state.argtypes = [C.c_int, _P(State)]
        lib.ref_single_projector_sample.argtypes = [_P(Projector), C.c_int, C.c_int, _P(C.c_uint64)]
        lib.ref_single_projector_sample.restype = C.c_double
        lib.ref_sampled_projector.argtypes = [_P(Projector), C.c_int, C.c_int, _P(C.c_uint64), C.c_double, C.c_int]
        lib.ref_sampled_projector.restype = C.c_double

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "libcircref.so"))

    def projector_bytes(self, P):
        """(phaseSign.data, phaseComplex.data, xs.data, zs.data) of the reference's own struct Projector for P"""
        cap = (P.nstabs * P.nqubits + 7) // 8 + 8
        bufs = [(C.c_ubyte * cap)() for _ in range(4)]
        n = self.lib.ref_projector_bytes(C.byref(P), bufs[0], bufs[1], bufs[2], bufs[3], cap)
        assert n >= 0
        return bufs

    def L_bytes(self, t, Lrows):
        """BitMatrix.data of the reference's k x t matrix L"""
        cap = (len(Lrows) * t + 7) // 8 + 8
        buf = (C.c_ubyte * cap)()
        assert self.lib.ref_L_bytes(len(Lrows), t, u64_array(Lrows), buf, cap) >= 0
        return buf

    def srand(self, seed):
        self.lib.ref_srand(seed)

    def random_state(self, n):
        s = State()
        self.lib.ref_random_state(n, C.byref(s))
        return s

    def single_projector_sample(self, P, exact, Lrows):
        return self.lib.ref_single_projector_sample(C.byref(P), int(bool(exact)), 0 if exact else len(Lrows),
                                                    u64_array(Lrows or []))

    def sampled_projector(self, P, exact, Lrows, norm, samples):
        return self.lib.ref_sampled_projector(C.byref(P), int(bool(exact)), 0 if exact else len(Lrows),
                                              u64_array(Lrows or []), norm, samples)


def epm_equal(a, b):
    """The reference's own comparison convention (tests/units/stabtests.c:84-86, 371-376):
    eps equal; if eps != 0 then p equal and m equal mod 8; if eps == 0, p and m are ignored."""
    if a[0] != b[0]:
        return False
    if a[0] == 0:
        return True
    return a[1] == b[1] and (a[2] - b[2]) % 8 == 0
