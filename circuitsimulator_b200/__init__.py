"""circuitsimulator_b200 — B200-native stabilizer-rank norm estimation.

Python face of the C ABI in include/bgnorm.h (libbgnorm.so: hand-written sm_100a CUDA).
It replaces ONE path of patrickrall/CircuitSimulator — the L x chi inner-product loop of
libcirc/innerprod.c — behind the reference's own back-end boundary.  The names below
follow the reference's C functions:

    Backend.sampled_norm   <-> multiSampledProjector   (libcirc/innerprod.c:23-84)
    Backend.exact_norm     <-> exactProjector          (libcirc/innerprod.c:148-199)
    Backend.inner_products <-> innerProductExact       (libcirc/stabilizer/stabilizer.c:589-659)
    Backend.measure_pauli  <-> measurePauli            (libcirc/stabilizer/stabilizer.c:827-959)
    run_backend            <-> the `mpibackend` executable's stdin/stdout protocol
                               (libcirc/probability.c:30-216), via the drop-in `bgbackend`

There is no CPU fallback: importing works anywhere, but every computation needs the CUDA
library and a B200; a missing library or device raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbgnorm.so")
BACKEND_PATH = os.path.join(HERE, "bgbackend")
MAX_T = 64
MAX_STABS = 128

STATE_DTYPE = np.dtype([("n", "<i4"), ("k", "<i4"), ("Q", "<i4"), ("reserved", "<i4"),
                        ("h", "<u8"), ("D1", "<u8"), ("D2", "<u8"),
                        ("G", "<u8", (MAX_T,)), ("Gbar", "<u8", (MAX_T,)), ("J", "<u8", (MAX_T,))])


class BGError(RuntimeError):
    pass


class Projector(C.Structure):
    """bg_projector: generators i^phase Z(zs) X(xs) (reference: struct Projector, libcirc/utils/comms.h:4-11)."""
    _fields_ = [("nstabs", C.c_int32), ("nqubits", C.c_int32),
                ("phase", C.c_uint8 * MAX_STABS),
                ("xs", C.c_uint64 * MAX_STABS), ("zs", C.c_uint64 * MAX_STABS)]

    @staticmethod
    def make(nqubits, phases, xs, zs):
        if len(phases) > MAX_STABS:
            raise BGError("projector with %d generators (max %d)" % (len(phases), MAX_STABS))
        p = Projector()
        p.nstabs, p.nqubits = len(phases), nqubits
        for i, (ph, x, z) in enumerate(zip(phases, xs, zs)):
            p.phase[i], p.xs[i], p.zs[i] = int(ph) % 4, int(x), int(z)
        return p


class Stats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("pairs", C.c_uint64), ("pair_launches", C.c_uint64),
                ("launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("prepare_ms", C.c_double), ("pairs_ms", C.c_double), ("overlapped", C.c_uint64)]


_lib = None
_P = C.POINTER


def load_library():
    """dlopen libbgnorm.so (built in-tree by `make -C circuitsimulator_b200/csrc`). Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BGError("CUDA extension missing: %s (run __graft_entry__.build()); there is no CPU fallback"
                      % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, u64, i32, dbl = C.c_void_p, C.c_uint64, C.c_int, C.c_double
    sig = {
        "bg_init": [_P(vp), i32],
        "bg_device_count": [_P(i32)],
        "bg_set_shard": [vp, i32, i32],
        "bg_set_allreduce": [vp, i32],
        "bg_nccl_unique_id": [_P(C.c_uint8)],
        "bg_nccl_join": [vp, _P(C.c_uint8)],
        "bg_set_decomposition": [vp, i32, i32, i32, _P(u64)],
        "bg_set_decomposition_bitmatrix": [vp, i32, i32, i32, _P(C.c_uint8)],
        "bg_projector_from_bitmatrix": [_P(Projector), i32, i32, _P(C.c_uint8), _P(C.c_uint8), _P(C.c_uint8), _P(C.c_uint8)],
        "bg_sampled_norm": [vp, _P(Projector), u64, i32, u64, dbl, _P(dbl)],
        "bg_exact_norm": [vp, _P(Projector), dbl, _P(dbl)],
        "bg_exact_norm_parts": [vp, _P(Projector), dbl, _P(dbl)],
        "bg_inner_products": [vp, C.c_size_t, vp, vp, _P(C.c_int32)],
        "bg_sampled_norm_from_states": [vp, _P(Projector), i32, C.c_size_t, vp, _P(C.c_int32), _P(dbl), _P(dbl)],
        "bg_measure_pauli": [vp, C.c_size_t, vp, _P(C.c_int32), _P(u64), _P(u64), _P(dbl)],
        "bg_random_states": [vp, i32, u64, i32, u64, C.c_size_t, vp],
        "bg_decomposition_terms": [vp, u64, C.c_size_t, vp],
        "bg_get_stats": [vp, _P(Stats)],
        "bg_sampled_prepare": [vp, _P(Projector), u64, i32, u64],
        "bg_sampled_run": [vp],
        "bg_sampled_finish": [vp, dbl, _P(dbl)],
        "bg_sampled_norm2": [vp, _P(Projector), _P(Projector), u64, i32, u64, u64, dbl, _P(dbl)],
        "bg_sampled_prepare2": [vp, _P(Projector), _P(Projector), u64, i32, u64, u64],
        "bg_sampled_finish2": [vp, dbl, _P(dbl)],
        "bg_sampled_per_sample": [vp, i32, u64, C.c_size_t, _P(dbl)],
        "bg_set_stream": [vp, vp],
        "bg_measure_int_peak": [vp, _P(dbl), _P(dbl)],
        "bg_decomposition_weights": [vp, i32, i32, _P(u64), _P(u64)],
    }
    for name, args in sig.items():
        f = getattr(lib, name)
        f.argtypes = args
        f.restype = C.c_int
    lib.bg_shutdown.argtypes = [vp]
    lib.bg_shutdown.restype = None
    lib.bg_last_error.argtypes = [vp]
    lib.bg_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def exported_symbols():
    """Every entry point include/bgnorm.h declares (used by the CPU-side ABI test)."""
    return ["bg_init", "bg_device_count", "bg_shutdown", "bg_last_error", "bg_set_shard", "bg_set_allreduce", "bg_nccl_unique_id", "bg_nccl_join",
            "bg_set_decomposition", "bg_set_decomposition_bitmatrix", "bg_projector_from_bitmatrix",
            "bg_sampled_norm", "bg_exact_norm", "bg_exact_norm_parts", "bg_inner_products", "bg_sampled_norm_from_states",
            "bg_measure_pauli", "bg_random_states", "bg_decomposition_terms", "bg_get_stats",
            "bg_sampled_prepare", "bg_sampled_run", "bg_sampled_finish", "bg_sampled_norm2", "bg_sampled_prepare2",
            "bg_sampled_finish2", "bg_sampled_per_sample", "bg_set_stream", "bg_measure_int_peak", "bg_decomposition_weights"]


def _states_arg(arr):
    arr = np.ascontiguousarray(arr, dtype=STATE_DTYPE)
    return arr, arr.ctypes.data


class Backend:
    """One context = one CUDA device, one stream, (optionally) one NCCL rank."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        if self.lib.bg_init(C.byref(self.ctx), device) != 0:
            raise BGError(self.lib.bg_last_error(None).decode())

    def close(self):
        if getattr(self, "ctx", None) and self.ctx.value:
            self.lib.bg_shutdown(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise BGError(self.lib.bg_last_error(self.ctx).decode())

    # -- sharding / NCCL
    def set_shard(self, rank, world):
        self._ck(self.lib.bg_set_shard(self.ctx, rank, world))

    def set_allreduce(self, enabled):
        self._ck(self.lib.bg_set_allreduce(self.ctx, int(bool(enabled))))

    def nccl_unique_id(self):
        buf = (C.c_uint8 * 128)()
        if self.lib.bg_nccl_unique_id(buf) != 0:
            raise BGError(self.lib.bg_last_error(None).decode())
        return bytes(buf)

    def nccl_join(self, uid):
        buf = (C.c_uint8 * 128)(*uid)
        self._ck(self.lib.bg_nccl_join(self.ctx, buf))

    # -- inputs
    def set_decomposition(self, t, exact, L_rows=()):
        rows = (C.c_uint64 * max(1, len(L_rows)))(*[int(r) for r in L_rows])
        self._ck(self.lib.bg_set_decomposition(self.ctx, t, int(bool(exact)), len(L_rows), rows))

    # -- hot path
    def sampled_norm(self, P, samples, bins=1, seed=0, norm=1.0):
        out = C.c_double()
        self._ck(self.lib.bg_sampled_norm(self.ctx, C.byref(P), samples, bins, seed, norm, C.byref(out)))
        return out.value

    def exact_norm(self, P, norm=1.0):
        out = C.c_double()
        self._ck(self.lib.bg_exact_norm(self.ctx, C.byref(P), norm, C.byref(out)))
        return out.value

    def sampled_prepare(self, P, samples, bins=1, seed=0):
        self._ck(self.lib.bg_sampled_prepare(self.ctx, C.byref(P), samples, bins, seed))

    def sampled_run(self):
        self._ck(self.lib.bg_sampled_run(self.ctx))

    def sampled_finish(self, norm=1.0):
        out = C.c_double()
        self._ck(self.lib.bg_sampled_finish(self.ctx, norm, C.byref(out)))
        return out.value

    def sampled_norm2(self, G, H, samples, bins=1, seed_g=0, seed_h=1, norm=1.0):
        out = (C.c_double * 2)()
        self._ck(self.lib.bg_sampled_norm2(self.ctx, C.byref(G), C.byref(H), samples, bins, seed_g, seed_h, norm, out))
        return out[0], out[1]

    def sampled_prepare2(self, G, H, samples, bins=1, seed_g=0, seed_h=1):
        self._ck(self.lib.bg_sampled_prepare2(self.ctx, C.byref(G), C.byref(H), samples, bins, seed_g, seed_h))

    def sampled_finish2(self, norm=1.0):
        out = (C.c_double * 2)()
        self._ck(self.lib.bg_sampled_finish2(self.ctx, norm, out))
        return out[0], out[1]

    # -- parity / debug
    def sampled_per_sample(self, projector, first, count):
        """per-sample values of the last finished device-RNG job (this rank's local sample indices first .. first+count)"""
        out = np.zeros(count, dtype=np.float64)
        self._ck(self.lib.bg_sampled_per_sample(self.ctx, int(projector), first, count, out.ctypes.data_as(_P(C.c_double))))
        return out

    def inner_products(self, a, b):
        a, pa = _states_arg(a)
        b, pb = _states_arg(b)
        assert len(a) == len(b)
        epm = np.zeros((len(a), 3), dtype=np.int32)
        self._ck(self.lib.bg_inner_products(self.ctx, len(a), pa, pb, epm.ctypes.data_as(_P(C.c_int32))))
        return epm

    def sampled_norm_from_states(self, P, thetas, project=True, want_epm=False, chi=None):
        th, pth = _states_arg(thetas)
        n = len(th)
        per = np.zeros(n, dtype=np.float64)
        mean = C.c_double()
        epm = None
        if want_epm:
            assert chi is not None
            epm = np.zeros((n, chi, 3), dtype=np.int32)
        self._ck(self.lib.bg_sampled_norm_from_states(
            self.ctx, C.byref(P), int(bool(project)), n, pth,
            epm.ctypes.data_as(_P(C.c_int32)) if want_epm else None,
            per.ctypes.data_as(_P(C.c_double)), C.byref(mean)))
        return dict(mean=mean.value, per_sample=per, epm=epm)

    def measure_pauli(self, states, m, zeta, xi):
        st = np.array(states, dtype=STATE_DTYPE, copy=True)
        n = len(st)
        m = np.ascontiguousarray(m, dtype=np.int32)
        zeta = np.ascontiguousarray(zeta, dtype=np.uint64)
        xi = np.ascontiguousarray(xi, dtype=np.uint64)
        res = np.zeros(n, dtype=np.float64)
        self._ck(self.lib.bg_measure_pauli(self.ctx, n, st.ctypes.data, m.ctypes.data_as(_P(C.c_int32)),
                                           zeta.ctypes.data_as(_P(C.c_uint64)), xi.ctypes.data_as(_P(C.c_uint64)),
                                           res.ctypes.data_as(_P(C.c_double))))
        return st, res

    def random_states(self, t, seed, bin_, first, count):
        out = np.zeros(count, dtype=STATE_DTYPE)
        self._ck(self.lib.bg_random_states(self.ctx, t, seed, bin_, first, count, out.ctypes.data))
        return out

    def decomposition_terms(self, first, count):
        out = np.zeros(count, dtype=STATE_DTYPE)
        self._ck(self.lib.bg_decomposition_terms(self.ctx, first, count, out.ctypes.data))
        return out

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.bg_set_stream(self.ctx, C.c_void_p(cuda_stream_ptr)))

    def measure_int_peak(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.lib.bg_measure_int_peak(self.ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def decomposition_weights(self, t, L_rows):
        """decompose()'s fidelity loop (libcirc/probability.c:373-391) on the device: hist[w] = number of
        the 2^k combinations of the rows of L with Hamming weight w."""
        k = len(L_rows)
        rows = (C.c_uint64 * max(k, 1))(*[int(r) for r in L_rows])
        hist = (C.c_uint64 * 65)()
        self._ck(self.lib.bg_decomposition_weights(self.ctx, t, k, rows, hist))
        return [int(v) for v in hist]

    def stats(self):
        s = Stats()
        self._ck(self.lib.bg_get_stats(self.ctx, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}


def run_backend(stream_text, env=None, timeout=None):
    """Feed one instruction stream (the 13 scalars + 2 projectors that libcirc/probability.py:247-272
    writes) to the drop-in back end on stdin; return (numerator, denominator, all stdout lines) the way
    libcirc/probability.py:283-305 parses them."""
    if not os.path.exists(BACKEND_PATH):
        raise BGError("back-end executable missing: %s (run __graft_entry__.build())" % BACKEND_PATH)
    e = dict(os.environ)
    if env:
        e.update({k: str(v) for k, v in env.items()})
    p = subprocess.run([BACKEND_PATH, "stdin"], input=stream_text.encode(), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, env=e, timeout=timeout)
    lines = p.stdout.decode().splitlines()
    try:
        return float(lines[-2]), float(lines[-1]), lines
    except Exception:
        raise BGError("back end gave no result: rc=%d stdout=%r stderr=%r" % (p.returncode, lines[-5:], p.stderr.decode()[-500:]))
