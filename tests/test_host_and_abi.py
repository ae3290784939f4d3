"""Host-side logic and the C-ABI surface, no GPU needed: the shared library loads and exports every
symbol include/bgnorm.h declares; without a device the product fails loudly (no CPU fallback);
the drop-in executable parses the reference's token protocol; the sample sharding used for N GPUs
is exercised with a world_size-2 gloo group."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from util import parse_stream, GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    import circuitsimulator_b200 as bg
    if not os.path.exists(bg.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return bg


def test_header_symbols_exported():
    bg = _built()
    lib = bg.load_library()
    header = open(os.path.join(ROOT, "include", "bgnorm.h")).read()
    declared = sorted(set(re.findall(r"\b(bg_[a-z_0-9]+)\s*\(", header)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(bg.exported_symbols()) == declared


def test_struct_layouts_match():
    bg = _built()
    from oracle.oracle import State, Projector
    assert bg.STATE_DTYPE.itemsize == ctypes.sizeof(State) == 16 + 24 + 3 * 64 * 8
    assert ctypes.sizeof(bg.Projector) == ctypes.sizeof(Projector)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    bg = _built()
    with pytest.raises(bg.BGError) as e:
        bg.Backend(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_touch_oracle():
    """The product package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "circuitsimulator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "packed_oracle" not in txt and "libcircref" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f


def test_backend_protocol_errors_are_unparsable_lines():
    """Errors are printed strings, never floats, so libcirc/probability.py:302-312 raises."""
    bg = _built()
    p = subprocess.run([bg.BACKEND_PATH, "stdin"], input=b"0 0 0 10 1 4 0 1 1e-5 0 0 0 1 2 4 0\n",
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    last = p.stdout.decode().splitlines()[-1]
    with pytest.raises(ValueError):
        float(last)
    p = subprocess.run([bg.BACKEND_PATH, "/nonexistent/file"], stdout=subprocess.PIPE)
    assert "Error reading file" in p.stdout.decode()


def test_backend_clifford_closed_form_needs_no_gpu():
    """t = 0 (Clifford circuit): the closed form of innerprod.c:52-62 is evaluated on the host."""
    bg = _built()
    # two projectors on 0 qubits: G = {+I, -I}, H = {+I}
    txt = "0 0 0 100 1 0 0 1 1e-5 0 0 0 0\n2 0 0 2\n1 0 0\n"
    num, den, _ = bg.run_backend(txt)
    assert abs(num - (1 + 1 - 1) / 3.0) < 1e-15 and abs(den - 1.0) < 1e-15


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle.oracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = Oracle()
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", "htstack_t4.txt"))
    samples = 37
    count = samples // world + (1 if rank < samples % world else 0)        # shard_count() of bgnorm.cu
    part, _ = o.sampled_sum_philox(G, True, [], 9, 0, rank, world, count)   # l = rank + world*j
    t = torch.tensor([part, float(count)], dtype=torch.float64)
    dist.all_reduce(t)
    if rank == 0:
        q.put((float(t[0]), float(t[1])))
    dist.destroy_process_group()


def test_strided_sharding_world2_gloo(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, count = q.get(timeout=120)
    for p in procs:
        p.join(60)
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", "htstack_t4.txt"))
    whole, _ = oracle.sampled_sum_philox(G, True, [], 9, 0, 0, 1, 37)
    assert count == 37
    assert abs(total - whole) <= 1e-13 * abs(whole)


def test_backend_answers_without_eof_on_stdin():
    """The front end writes the stream and keeps the pipe open (libcirc/probability.py:283-291): the back
    end must answer after the last token, not at end-of-file."""
    import select
    import time
    bg = _built()
    p = subprocess.Popen([bg.BACKEND_PATH, "stdin"], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    try:
        p.stdin.write(b"0 0 0 100 1 0 0 1 1e-5 0 0 0 0\n2 0 0 2\n1 0 0\n")
        p.stdin.flush()                                   # NOT closed
        deadline, buf = time.time() + 20, b""
        while time.time() < deadline and p.poll() is None:
            r, _, _ = select.select([p.stdout], [], [], 0.5)
            if r:
                chunk = os.read(p.stdout.fileno(), 65536)
                if not chunk:
                    break
                buf += chunk
        if p.poll() is None:
            p.wait(timeout=5)
        buf += p.stdout.read()
        lines = buf.decode().splitlines()
        assert abs(float(lines[-2]) - 1 / 3) < 1e-15 and float(lines[-1]) == 1.0
    finally:
        if p.poll() is None:
            p.kill()


def test_unmodified_front_end_drives_the_backend():
    """libcirc.probability() of the reference (when its tree is present) spawns bgbackend through
    cpath/mpirun, feeds it the token stream and parses the answer.  Without a GPU the back end answers
    with an Error line and the front end raises RuntimeError (probability.py:302-312) — no hang, no
    silent fallback; with a GPU it returns the probability."""
    ref = os.environ.get("BG_REFERENCE_ROOT", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "libcirc")):
        pytest.skip("reference tree not present")
    bg = _built()
    code = (
        "import sys, warnings; warnings.simplefilter('ignore'); sys.path.insert(0, %r)\n"
        "from libcirc.probability import probability\n"
        "from libcirc.compile.compilecirc import compileCircuit\n"
        "c = compileCircuit(fname='circuits/HTstack.circ')\n"
        "try:\n"
        "    p = probability(c, {0: 0}, config={'samples': 512, 'forceSample': True, 'quiet': True,\n"
        "                    'cpath': %r, 'mpirun': '/usr/bin/env'})\n"
        "    print('PROB', p)\n"
        "except RuntimeError:\n"
        "    print('RUNTIMEERROR')\n" % (ref, bg.BACKEND_PATH))
    out = subprocess.run([sys.executable, "-c", code], cwd=ref, capture_output=True, text=True, timeout=120).stdout
    import torch
    if torch.cuda.is_available():
        prob = float([ln for ln in out.splitlines() if ln.startswith("PROB")][0].split()[1])
        assert abs(prob - 0.97855339) < 0.2
    else:
        assert "RUNTIMEERROR" in out and "no CPU fallback" in out


def test_bench_host_helpers(monkeypatch):
    """bench.py's host-side pieces that no GPU run exercises in isolation: the instruction-stream writer is
    the inverse of the parser (token protocol of libcirc/probability.c:74-127), the synthetic projector only
    has Hermitian generators, and the clock sampler summarises NVML readings (in-process, fake library here)
    or falls back to the nvidia-smi poller."""
    import types
    import time as _time
    import bench
    cfg, G, H, samples, k, _ = bench.load_config("hidden_shift_n40_t16_L16384")
    text = bench.stream_text(cfg["t"], samples, k, cfg["exact"], G, H)
    path = os.path.join(ROOT, "tests", "golden", "streams", "hs_t16_bit6.txt")
    tok_a, tok_b = text.split(), open(path).read().split()
    assert tok_a[13:] == tok_b[13:]                        # both projectors, token for token
    assert int(tok_a[5]) == cfg["t"] == 16 and int(tok_a[3]) == samples
    _, Gs, Hs = bench.synthetic_stream()
    for (nq, ph, xs, zs) in (Gs, Hs):
        assert nq == 60 and all((bin(x & z).count("1") - p) % 2 == 0 for p, x, z in zip(ph, xs, zs))
    assert len(bench.fixed_L(9, 40)) == 9 and bench.fixed_L(9, 40) == bench.fixed_L(9, 40)

    fake = types.ModuleType("pynvml")
    fake.NVML_CLOCK_SM = 1
    fake.nvmlInit = lambda: None
    fake.nvmlDeviceGetHandleByIndex = lambda i: i
    fake.nvmlDeviceGetMaxClockInfo = lambda h, c: 1965
    fake.nvmlDeviceGetClockInfo = lambda h, c: 1950
    fake.nvmlDeviceGetCurrentClocksEventReasons = lambda h: 4 | 64
    fake.nvmlClocksEventReasonHwSlowdown, fake.nvmlClocksEventReasonHwThermalSlowdown = 8, 64
    fake.nvmlClocksEventReasonSwThermalSlowdown, fake.nvmlClocksEventReasonSwPowerCap = 32, 4
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    s = bench.ClockSampler(0)
    s.start()
    _time.sleep(0.2)
    s.stop_flag = True
    s.join(2)
    out = s.summary()
    assert out["source"] == "nvml" and out["sm_mhz"] == 1950 and out["sm_max_mhz"] == 1965 and out["samples"] >= 2
    assert out["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]
    broken = types.ModuleType("pynvml")                  # no NVML: the poller of the nvidia-smi binary is used
    monkeypatch.setitem(sys.modules, "pynvml", broken)
    assert bench.ClockSampler(0).nvml is None


def test_bitmatrix_adapter_on_the_reference_layout(reference):
    """bg_projector_from_bitmatrix (the adapter of the level-2 binding) fed with the .data byte
    arrays of a struct Projector that the REFERENCE's own constructors built (BitVector / BitMatrix, MSB-first,
    no row padding: matrix.c:124-131, 330-339; comms.h:4-11) must give back the packed projector."""
    import ctypes as C
    import numpy as np
    bg = _built()
    lib = bg.load_library()
    from oracle.oracle import Projector as OProj
    rs = np.random.RandomState(2)
    for t, n in [(1, 1), (4, 3), (7, 5), (16, 11), (33, 20), (40, 29), (63, 7), (64, 128)]:
        ph = [int(rs.randint(0, 4)) for _ in range(n)]
        xs = [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) | ((int(rs.randint(0, 4)) << 62) & ((1 << t) - 1)) for _ in range(n)]
        zs = [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) | ((int(rs.randint(0, 4)) << 62) & ((1 << t) - 1)) for _ in range(n)]
        P = OProj.make(t, ph, xs, zs)
        sign, cplx, bx, bz = reference.projector_bytes(P)
        out = bg.Projector()
        rc = lib.bg_projector_from_bitmatrix(C.byref(out), n, t, C.cast(sign, C.POINTER(C.c_uint8)), C.cast(cplx, C.POINTER(C.c_uint8)),
                                             C.cast(bx, C.POINTER(C.c_uint8)), C.cast(bz, C.POINTER(C.c_uint8)))
        assert rc == 0
        assert (out.nstabs, out.nqubits) == (n, t)
        assert [out.phase[i] for i in range(n)] == ph
        assert [out.xs[i] for i in range(n)] == xs and [out.zs[i] for i in range(n)] == zs
