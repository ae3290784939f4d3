"""ctypes face of tests/emu/libbgemu.so — the product's warp-level device code run on the CPU
warp emulator.  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle.oracle import State, Projector, u64_array

HERE = os.path.dirname(os.path.abspath(__file__))
_P = C.POINTER


def compact_state(s, A):
    """Active-mask layout -> contiguous layout (active rows first, order preserved)."""
    n = s.n
    act = [v for v in range(n) if (A >> v) & 1]
    perm = act + [v for v in range(n) if not (A >> v) & 1]
    k = len(act)
    o = State()
    o.n, o.k, o.Q, o.h = n, k, s.Q, s.h
    d1 = d2 = 0
    for r, v in enumerate(perm):
        o.G[r] = s.G[v]
        o.Gbar[r] = s.Gbar[v]
        if r < k:
            d1 |= ((s.D1 >> v) & 1) << r
            d2 |= ((s.D2 >> v) & 1) << r
            row = 0
            for c in range(k):
                row |= ((s.J[v] >> perm[c]) & 1) << c
            o.J[r] = row
    o.D1, o.D2 = d1, d2
    return o


class Emu:
    def __init__(self, libname="libbgemu.so"):
        subprocess.check_call(["make", "-C", HERE], stdout=subprocess.DEVNULL)
        lib = C.CDLL(os.path.join(HERE, libname))
        lib.emu_inner_product.argtypes = [_P(State), _P(State), _P(C.c_int32)]
        lib.emu_terms.argtypes = [_P(State), _P(Projector), C.c_int, C.c_int, C.c_int, C.c_int, _P(C.c_uint64),
                                  _P(C.c_int32), _P(C.c_int), _P(C.c_int), _P(C.c_longlong)]
        lib.emu_terms.restype = C.c_int
        lib.emu_terms_tpp.argtypes = lib.emu_terms.argtypes
        lib.emu_terms_tpp.restype = C.c_int
        lib.emu_measure_pauli.argtypes = [_P(State), _P(C.c_uint64), C.c_int, C.c_uint64, C.c_uint64]
        lib.emu_measure_pauli.restype = C.c_int
        lib.emu_random_state.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_uint64, _P(C.c_double), _P(State),
                                         _P(C.c_uint64)]
        lib.emu_terms_shb.argtypes = [_P(State), _P(Projector), C.c_int, C.c_int, C.c_int, _P(C.c_uint64), _P(C.c_int32),
                                      _P(C.c_int), _P(C.c_int), _P(C.c_longlong), _P(C.c_int)]
        lib.emu_terms_shb.restype = C.c_int
        lib.emu_expsum_selftest.argtypes = [C.c_int, C.c_ulonglong, C.c_int]
        lib.emu_expsum_selftest.restype = C.c_int
        lib.emu_chi_seconds.argtypes = [C.c_int]
        lib.emu_chi_seconds.restype = C.c_double
        lib.emu_work_counters.argtypes = [_P(C.c_ulonglong), C.c_int]
        lib.emu_prepare_both.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_uint64, _P(C.c_double), _P(Projector),
                                         _P(C.c_uint64), _P(C.c_uint64)]
        lib.emu_shb_plan_check.argtypes = [C.c_int, C.c_int, _P(C.c_uint64), _P(C.c_ulonglong)]
        lib.emu_shb_plan_check.restype = C.c_int
        lib.emu_zw32_selftest.restype = C.c_int
        self.lib = lib

    def shb_plan_check(self, t, L):
        """(code, digest): 0 no plan, 1 a plan whose invariants hold, < 0 the invariant that fails (emu_lib.cpp)"""
        rows = (C.c_uint64 * max(1, len(L)))(*[int(x) for x in L])
        dg = C.c_ulonglong()
        return self.lib.emu_shb_plan_check(t, len(L), rows, C.byref(dg)), dg.value

    def prepare_both(self, n, seed, bin_, sample, cdf, P):
        """One device-RNG sample projected by P: the SampleRec fields (136 uint64) as the warp-per-sample code and
        as the thread-per-sample code (bg_prep.cuh) produce them."""
        a = (C.c_uint64 * 136)()
        b = (C.c_uint64 * 136)()
        arr = (C.c_double * len(cdf))(*cdf)
        self.lib.emu_prepare_both(n, seed, bin_, sample, arr, C.byref(P), a, b)
        return list(a), list(b)

    def inner_product(self, a, b):
        out = (C.c_int32 * 3)()
        self.lib.emu_inner_product(C.byref(a), C.byref(b), out)
        return tuple(out)

    def terms(self, theta, P, project, exact, t, terms, tpp=False):
        """tpp=True: the chi loop through the thread-per-pair code (parity checks as Lagrange variables, as
        k_pairs_tpp stages them); tpp="pivot": the same with every check pivoted per term (BG_LAM_MAX=0)."""
        self.lib.emu_set_lam_max(0 if tpp == "pivot" else 4)
        n = len(terms)
        epm = np.zeros((n, 3), dtype=np.int32)
        npf, k = C.c_int(), C.c_int()
        zw = (C.c_longlong * 4)()
        fn = self.lib.emu_terms_tpp if tpp else self.lib.emu_terms
        alive = fn(C.byref(theta), C.byref(P), int(project), int(bool(exact)), t, n,
                                   u64_array(terms), epm.ctypes.data_as(_P(C.c_int32)), C.byref(npf),
                                   C.byref(k), zw)
        return dict(alive=alive, epm=epm, npf=npf.value, k=k.value, zw=list(zw))

    def terms_shb(self, theta, P, project, t, L):
        """the chi loop through the shared high-block reduction (k_pairs_shb's code); epm in natural term order.
        alive = -1: the decomposition has no plan; -2: theta has more parity checks than that kernel takes."""
        n = 1 << len(L)
        epm = np.zeros((n, 3), dtype=np.int32)
        npf, k = C.c_int(), C.c_int()
        zw = (C.c_longlong * 4)()
        hist = (C.c_int * 16)()
        alive = self.lib.emu_terms_shb(C.byref(theta), C.byref(P), int(project), t, len(L), u64_array(L),
                                       epm.ctypes.data_as(_P(C.c_int32)), C.byref(npf), C.byref(k), zw, hist)
        return dict(alive=alive, epm=epm, npf=npf.value, k=k.value, zw=list(zw), nleft_hist=list(hist))

    def work_counters(self, reset=True):
        out = (C.c_ulonglong * 6)()
        self.lib.emu_work_counters(out, int(reset))
        return dict(zip(["xors", "rows", "dimers", "monomers", "basis_changes", "pairs"], [int(v) for v in out]))

    def measure_pauli(self, s, m, zeta, xi):
        """returns (code, state in contiguous layout); code 0 annihilated, 1 unchanged, 2 factor 2^-1/2"""
        w = s.copy()
        A = C.c_uint64()
        code = self.lib.emu_measure_pauli(C.byref(w), C.byref(A), m, zeta, xi)
        return code, compact_state(w, A.value)

    def random_state(self, n, seed, bin_, sample, cdf):
        out = State()
        A = C.c_uint64()
        arr = (C.c_double * len(cdf))(*cdf)
        self.lib.emu_random_state(n, seed, bin_, sample, arr, C.byref(out), C.byref(A))
        return compact_state(out, A.value)
