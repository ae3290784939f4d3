#!/usr/bin/env python
"""bench.py — stabilizer inner products/sec on BASELINE.json's headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config NAME]
    (N > 1: launched by the driver under torchrun, one rank per GPU)

Workload (config.workload): BASELINE configs[3] — random hidden-shift circuit, n=40 qubits,
t=40 T gates, |L> decomposition with k=9 (chi=512), L=2^16 random stabilizer states per projector.
Input = the instruction stream the reference's unmodified front end wrote for that circuit
(tests/golden/streams/hs_t40_k9_bit0.txt; generator: tests/golden/make_fixtures.py).
One STEP = one probability() back-end evaluation = both projectors (G', H'):
2 x 2^16 x 512 = 67,108,864 stabilizer inner products (+ the 2 x 2^16 theta draws and projections).
Samples are sharded by stride across ranks (total work is fixed: "scaling": "strong"); the partial
sums are all-reduced over NCCL inside the library.

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

STREAMS = os.path.join(ROOT, "tests", "golden", "streams")
CONFIGS = {
    # name: (stream file, samples per projector, forced k for |L>, description)
    "hidden_shift_n40_t40_k9_L65536": ("hs_t40_k9_bit0.txt", 65536, 9,
                                       "random hidden-shift n=40, t=40 T gates, |L> k=9 (chi=512), L=2^16"),
    "hidden_shift_n40_t40_k9_L8192": ("hs_t40_k9_bit0.txt", 8192, 9,
                                      "config 4 at 1/8 of the samples (what one of 8 GPUs sees)"),
    "hidden_shift_n40_t16_L16384": ("hs_t16_bit6.txt", 16384, 0,
                                    "random hidden-shift n=40, t=16 T gates, exact |H^t> (chi=256), L=2^14"),
    "htstack_t4_L1024": ("htstack_t4.txt", 1024, 0, "HTstack.circ output 0, t=4 (chi=4), L=1024"),
    # BASELINE configs[4], second half: synthetic t=60 sweep (60-bit states in 64-bit rows), theta from the
    # device RNG, phi from prepL, random Hermitian Pauli projectors with 40 / 39 generators
    "synthetic_t60_k8_L65536": (None, 65536, 8, "synthetic t=60, |L> k=8 (chi=256), L=2^16"),
    "synthetic_t60_k12_L65536": (None, 65536, 12, "synthetic t=60, |L> k=12 (chi=4096), L=2^16"),
    "synthetic_t60_k16_L16384": (None, 16384, 16, "synthetic t=60, |L> k=16 (chi=65536), L=2^14"),
}
DEFAULT_CONFIG = "hidden_shift_n40_t40_k9_L65536"


def parse_stream(path):
    """13 scalars + 2 projectors (libcirc/probability.c:74-127, libcirc/utils/comms.c:9-36)."""
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
        projs.append((nq, ph, xs, zs))
    return cfg, projs[0], projs[1]


def synthetic_stream(t=60, nstabs=(40, 39), seed=60):
    """A synthetic back-end input: random Hermitian Pauli generators i^m Z(z) X(x), m = |x & z| mod 2 (+2),
    every one with a non-trivial X part — like the gadgetized-T projectors of real circuits (the hidden-shift
    and phase-estimation streams keep dim K(theta) within ~3 of t after projection); Z-only generators would
    each cut the dimension by one instead."""
    import numpy as np
    rs = np.random.RandomState(seed)
    cfg = {"t": t, "exact": 0, "k": 0}
    projs = []
    for ns in nstabs:
        ph, xs, zs = [], [], []
        for _ in range(ns):
            x = (int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1)) | 1
            z = int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1)
            ph.append((bin(x & z).count("1") + 2 * int(rs.randint(0, 2))) % 4)
            xs.append(x)
            zs.append(z)
        projs.append((t, ph, xs, zs))
    return cfg, projs[0], projs[1]


def load_config(cfgname):
    stream, samples, k, desc = CONFIGS[cfgname]
    if stream is None:
        cfg, G, H = synthetic_stream()
    else:
        cfg, G, H = parse_stream(os.path.join(STREAMS, stream))
    return cfg, G, H, samples, k, desc


def stream_text(t, samples, k, exact, G, H):
    """The token stream of libcirc/probability.py:247-272 for this input (forceSample on)."""
    tok = [0, 0, 0, samples, 1, t, k, int(bool(exact)), 1e-05, 0, 0, 0, 1]
    for (nq, ph, xs, zs) in (G, H):
        tok += [len(ph), nq if ph else 0]
        for p_, x, z in zip(ph, xs, zs):
            tok.append(p_)
            for q in range(nq):
                tok += [(x >> q) & 1, (z >> q) & 1]
    return "\n".join(str(v) for v in tok) + "\n"


def fixed_L(k, t):
    """The k x t matrix L: uniform random bits (BitMatrixSetRandom, libcirc/utils/matrix.c:301-306),
    fixed by seed so that every arm / rank / run sees the same decomposition."""
    import numpy as np
    rs = np.random.RandomState(20240)
    return [int(rs.randint(0, 2 ** 62)) & ((1 << t) - 1) for _ in range(k)]


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """SM clock + clock-event (throttle) reasons during the timed region (B200_PROFILING.md recipe).  Reads
    NVML in-process (the library nvidia-smi itself uses): a query takes microseconds and does not fork, so a
    timed region of a few tens of ms is not perturbed; falls back to polling the nvidia-smi binary."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.rows, self.stop_flag = device, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[device]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else device
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)   # static: asked once (1.5 ms a call)
            self.nvml = pynvml
            self._nvml_row()         # the first call of each query can take tens of ms: pay that here, not in the timed region
        except Exception:
            self.nvml = None

    def _nvml_row(self):
        n, h = self.nvml, self.handle
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)              # ~10 us each (measured under load)
        mx = self.max_sm
        pw = 0.0
        r = n.nvmlDeviceGetCurrentClocksEventReasons(h)

        def act(name):
            bit = getattr(n, name, 0)
            return "Active" if (r & bit) else "Not Active"
        return [str(self.device), str(sm), str(mx), "%.1f" % pw, hex(r),
                act("nvmlClocksEventReasonHwSlowdown"), act("nvmlClocksEventReasonHwThermalSlowdown"),
                act("nvmlClocksEventReasonSwThermalSlowdown"), act("nvmlClocksEventReasonSwPowerCap")]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.rows.append(self._nvml_row())
                else:
                    out = subprocess.run(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.device)], capture_output=True, text=True, timeout=5).stdout
                    for ln in out.strip().splitlines():
                        self.rows.append([c.strip() for c in ln.split(",")])
            except Exception:
                pass
            time.sleep(0.05 if self.nvml is not None else 0.3)

    def summary(self):
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for j, nme in enumerate(names):
                if len(r) > 5 + j and r[5 + j].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------ reference arm
def reference_sample(t, exact, k, chi, seconds):
    """(samples per core, k for the reference, terms per sample) so that one step takes ~`seconds`."""
    rate = 14.0 if t >= 33 else 700.0          # pairs/s/core of the reference, build container
    budget = max(1.0, rate * seconds / 2)      # pairs per projector per core
    if exact or chi <= budget:
        return max(1, int(budget // chi)), k, chi
    k_ref = max(1, int(budget).bit_length() - 1)
    return 1, k_ref, 1 << k_ref


def run_reference(args, cfgname):
    """--impl reference: the reference's own C implementation (oracle/_ref/mpibackend_ref, the
    unmodified sources behind a single-rank MPI shim, -O2) on the host cores: one process per core,
    each on a disjoint shard of samples — what the reference's MPI rank stride does
    (libcirc/probability.c:268).  Each step is a BOUNDED sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, G, H, samples, k, desc = load_config(cfgname)
    t = cfg["t"]
    exact = cfg["exact"] if k == 0 else 0
    chi = (1 << ((t + 1) // 2)) if exact else (1 << k)
    exe = os.path.join(ROOT, "oracle", "_ref", "mpibackend_ref")
    kind = "reference"
    cores = os.cpu_count() or 1
    if not os.path.exists(exe):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/mpibackend_ref not built"}))
        return
    # bounded sample: every core evaluates `per_core` samples of each projector against `chi_ref`
    # terms.  For |L> workloads the reference is given a smaller k (its own random L): the cost of
    # one inner product does not depend on k, and a full 2 x 512-term sample takes it > 70 s per core.
    # the whole --steps K run is sized to ~2.5 minutes
    per_core, k_ref, chi_ref = reference_sample(t, exact, k, chi, seconds=max(2.0, min(16.0, 150.0 / (args.steps + 1))))
    text = stream_text(t, per_core, k_ref, exact, G, H)

    def step():
        procs = [subprocess.Popen([exe, "stdin"], stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                  stderr=subprocess.DEVNULL) for _ in range(cores)]
        for p in procs:
            p.stdin.write(text.encode())
            p.stdin.close()
        for p in procs:
            p.stdout.read()
            p.wait()

    for _ in range(min(args.warmup, 1)):      # CPU code: one untimed pass is all the warm-up there is
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    pairs_per_step = cores * per_core * 2 * chi_ref
    value = pairs_per_step * args.steps / dt
    line = {"metric": "stabilizer inner products/sec", "value": value, "unit": "inner products/s",
            "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64 bit rows + Z8 integer phases", "data": "synthetic",
            "config": config_dict(cfgname, desc, t, chi, samples,
                                  {"chi": chi_ref, "k": k_ref, "samples_per_core": per_core, "cores": cores,
                                   "why": "the reference takes > 70 s per core for one full sample of 2 x %d terms; the cost of "
                                          "an inner product does not depend on k" % chi}),
            "cpu_baseline": {"value": value, "unit": "inner products/s", "cores": cores, "kind": kind,
                             "sample": "%d processes x %d samples x 2 projectors x %d terms per step "
                                       "(reference C sources unmodified, gcc -O2, single-rank MPI shim; the per-sample "
                                       "theta draw + projection is amortised over %d instead of %d terms)"
                                       % (cores, per_core, chi_ref, chi_ref, chi)},
            "e2e": {"value": value, "unit": "inner products/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def cpu_baseline_leg(cfgname, seconds=12.0):
    """Reported baseline next to the GPU number (rank 0, N=1): the compiled reference on all host
    cores for a bounded sample."""
    class A:
        pass
    cfg, G, H, samples, k, desc = load_config(cfgname)
    t = cfg["t"]
    exact = cfg["exact"] if k == 0 else 0
    chi = (1 << ((t + 1) // 2)) if exact else (1 << k)
    exe = os.path.join(ROOT, "oracle", "_ref", "mpibackend_ref")
    cores = os.cpu_count() or 1
    if not os.path.exists(exe):
        return None
    per_core, k_ref, chi_ref = reference_sample(t, exact, k, chi, seconds)
    text = stream_text(t, per_core, k_ref, exact, G, H).encode()
    t0 = time.perf_counter()
    procs = [subprocess.Popen([exe, "stdin"], stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                              stderr=subprocess.DEVNULL) for _ in range(cores)]
    for p in procs:
        p.stdin.write(text)
        p.stdin.close()
    for p in procs:
        p.stdout.read()
        p.wait()
    dt = time.perf_counter() - t0
    pairs = cores * per_core * 2 * chi_ref
    return {"value": pairs / dt, "unit": "inner products/s", "cores": cores, "kind": "reference",
            "sample": "%d processes x %d samples x 2 projectors x %d terms in %.1f s (reference C sources "
                      "unmodified, gcc -O2, single-rank MPI shim)" % (cores, per_core, chi_ref, dt)}



# ------------------------------------------------------------------------------------------ shared helpers
def config_dict(cfgname, desc, t, chi, samples, reference_sample_=None):
    """The `config` object of the JSON line: the same keys in both arms (our arm: reference_sample = null)."""
    return {"workload": cfgname, "description": desc, "t": t, "chi": chi,
            "samples_per_projector": samples, "projectors": 2,
            "step": "one probability() back-end evaluation: both projectors, theta draw + projection + "
                    "L x chi inner products + reduction (+ NCCL all-reduce for n_gpus > 1)",
            "l2": "inputs are regenerated every step from the counter-based RNG; the per-sample "
                  "records written and re-read each step (2 x %.0f MB) exceed the 126 MB L2" % (samples * 1072 / 1e6),
            "reference_sample": reference_sample_}


def pair_kernel_name(t, exact, k):
    """the kernel that evaluates the pairs of this configuration (bgnorm.cu: launch_pairs)"""
    if not exact and 32 < t <= 44 and k >= 6:
        return "k_pairs_shb"
    return "k_pairs_tpp"


def other_configs(ctx, bg, lop3_peak, skip):
    """The remaining BASELINE.json configurations, a few steps each on the same context: value, ms per step
    and — where profiles/work_model.json has the algorithmic work of that configuration — the roofline fraction."""
    import numpy as np
    wpath = os.path.join(ROOT, "profiles", "work_model.json")
    wms = json.load(open(wpath)) if os.path.exists(wpath) else {}
    out = []
    for name in ["htstack_t4_L1024", "hidden_shift_n40_t16_L16384", "synthetic_t60_k8_L65536",
                 "synthetic_t60_k12_L65536", "synthetic_t60_k16_L16384"]:
        if name == skip:
            continue
        cfg, Gd, Hd, samples, k, desc = load_config(name)
        t = cfg["t"]
        exact = cfg["exact"] if k == 0 else 0
        L = [] if exact else fixed_L(k, t)
        G, H = bg.Projector.make(*Gd), bg.Projector.make(*Hd)
        ctx.set_decomposition(t, exact, L)
        ctx.sampled_prepare2(G, H, samples, 1, 1001, 1002)
        nsteps = 3
        for _ in range(2):
            ctx.sampled_run(); ctx.sampled_finish2(1.0)
        t0 = time.perf_counter()
        pairs = kms = 0.0
        for _ in range(nsteps):
            ctx.sampled_run(); ctx.sampled_finish2(1.0)
            st = ctx.stats()
            pairs += st["pairs"]; kms += st["pairs_ms"]
        dt = time.perf_counter() - t0
        lane_ops = wms.get(name, {}).get("alu_lane_ops_per_pair")
        out.append({"workload": name, "description": desc, "t": t, "chi": (1 << ((t + 1) // 2)) if exact else (1 << k),
                    "samples_per_projector": samples, "kernel": pair_kernel_name(t, exact, k),
                    "value_this_rank": pairs / dt, "ms_per_step": 1e3 * dt / nsteps, "pair_kernel_ms": kms / nsteps,
                    "frac": (pairs / nsteps * lane_ops / (kms / nsteps * 1e-3) / lop3_peak) if lane_ops and kms > 0 else None})
    # config 1's path: the exact norm (chi (chi + 1) / 2 pair terms per projector), toffoli t=16
    cfg, Gd, Hd = parse_stream(os.path.join(STREAMS, "toffoli_q0.txt"))
    t = cfg["t"]
    ctx.set_decomposition(t, True, [])
    G = bg.Projector.make(*Gd)
    chi = 1 << ((t + 1) // 2)
    ctx.exact_norm(G, 1.0)
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        ctx.exact_norm(G, 1.0)
    dt = time.perf_counter() - t0
    st = ctx.stats()
    out.append({"workload": "toffoli_exact_norm_t16", "description": "exactProjector (innerprod.c:148-199): chi (chi + 1) / 2 pair terms, "
                "toffoli.circ qubit 0, t=16", "t": t, "chi": chi, "pair_terms_per_call": chi * (chi + 1) // 2,
                "kernel": "k_pairs_tpp<TRI>", "value_this_rank": chi * (chi + 1) // 2 * n / dt, "unit": "pair terms/s",
                "ms_per_call": 1e3 * dt / n, "kernel_ms": st["kernel_ms"],
                "reference_cpu": "1.8e3 pair terms/s/core (SURVEY section 6: 3 x 2 x 32896 terms in 107 s, -O2)"})
    return out


def stop_server(srv, sock):
    """the server's own shutdown message (bgbackend.cpp: serve), over its Unix socket"""
    import socket
    try:
        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        c.settimeout(10)
        c.connect(sock)
        c.sendall(b"shutdown\n")
        c.shutdown(socket.SHUT_WR)
        c.close()
    except Exception:
        pass
    try:
        srv.wait(timeout=15)
    except Exception:
        srv.kill()


def probability_seconds(t, samples, k, exact, Gd, Hd, gpus):
    """BASELINE metric, second half: wall seconds of one probability() back-end evaluation as the unmodified front
    end performs it (libcirc/probability.py:237-312: spawn the back end, write the token stream, read the two
    result lines), with the drop-in executable on `gpus` GPUs (BG_GPUS: one host thread + context per GPU, in-process
    NCCL) — one-shot (process start + CUDA/NCCL init every call) and against a persistent `bgbackend --serve`."""
    import tempfile
    import circuitsimulator_b200 as bg
    text = stream_text(t, samples, k, exact, Gd, Hd)
    env = {"BG_GPUS": gpus, "BG_SEED": 7}
    out = {"gpus": gpus, "stream": "config of this run through bgbackend (stdin token protocol)"}
    try:
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            num, den, _lines = bg.run_backend(text, env=env, timeout=300)
            ts.append(time.perf_counter() - t0)
        out["oneshot"] = min(ts)
        out["oneshot_all"] = ts
        out["result"] = [num, den]
        sock = os.path.join(tempfile.mkdtemp(prefix="bgsrv"), "s")
        e = dict(os.environ); e.update({k_: str(v) for k_, v in env.items()})
        srv = subprocess.Popen([bg.BACKEND_PATH, "--serve", sock], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        try:
            srv.stdout.readline()                       # "serving on ..." once the contexts are up
            cenv = dict(env); cenv["BG_SERVER"] = sock
            ts = []
            for _ in range(6):
                t0 = time.perf_counter()
                num2, den2, _lines = bg.run_backend(text, env=cenv, timeout=300)
                ts.append(time.perf_counter() - t0)
            out["served"] = min(ts[1:])
            out["served_all"] = ts
        finally:
            stop_server(srv, sock)
    except Exception as ex:                              # never fail the bench line because of this leg
        out["error"] = repr(ex)[:300]
    return out


def sample_qubits_seconds(gpus):
    """BASELINE config 5, first half: sampleQubits(phaseEstimation.circ, "MMM_") — the chain-rule weak simulation of
    libcirc/sample.py:32-83: one probability() call per sampled qubit (t=33, |L> k=8, 16384 samples), each conditioned on
    the outcomes so far, P0 = P / Psofar, then a draw.  The front end's calls were written with file= for every outcome
    prefix (tests/golden/make_fixtures.py); here the chain is replayed against `bgbackend --serve` the way
    recursiveSample walks it.  Returns wall seconds for the whole sample (3 probability() calls) and the outcome."""
    import random
    import tempfile
    import circuitsimulator_b200 as bg
    meta = json.load(open(os.path.join(STREAMS, "meta.json"))).get("phase_estimation_chain")
    if not meta:
        return None
    out = {"gpus": gpus, "circuit": "circuits/phaseEstimation.circ MMM_ (t=33, k=8, samples=16384 per projector)"}
    env = {"BG_GPUS": gpus, "BG_SEED": 11}
    sock = os.path.join(tempfile.mkdtemp(prefix="bgsq"), "s")
    e = dict(os.environ); e.update({k_: str(v) for k_, v in env.items()})
    try:
        srv = subprocess.Popen([bg.BACKEND_PATH, "--serve", sock], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    except Exception as ex:
        return {"error": repr(ex)[:300]}
    try:
        srv.stdout.readline()
        cenv = dict(env); cenv["BG_SERVER"] = sock
        runs = []
        for rep in range(4):
            rnd = random.Random(5)
            t0 = time.perf_counter()
            prefix, psofar, calls = "", 1.0, 0
            for _q in meta["qubits"]:
                ent = meta["streams"].get(prefix + "0")
                if ent is None:                       # the front end would have answered this call itself
                    break
                text = open(os.path.join(STREAMS, ent["stream"])).read()
                num, den, _lines = bg.run_backend(text, env=cenv, timeout=300)
                calls += 1
                prob = 0.0 if num == 0 else 2.0 ** ent["v_minus_u"] * num / den       # probability.py:320-326
                p0 = prob / psofar
                bit = 0 if rnd.random() < p0 else 1
                psofar = p0 if bit == 0 else 1.0 - p0                                  # sample.py:57, 83
                prefix += str(bit)
            runs.append((time.perf_counter() - t0, prefix, calls))
        out["seconds"] = min(r[0] for r in runs[1:])
        out["seconds_all"] = [r[0] for r in runs]
        out["sample"] = runs[-1][1]
        out["probability_calls"] = runs[-1][2]
    except Exception as ex:
        out["error"] = repr(ex)[:300]
    finally:
        stop_server(srv, sock)
    return out


def _packed_worker(args):
    """one host process: the shipped thread-per-pair algorithm compiled for the CPU (tests/emu, the product's own
    device source) on `n` samples of the workload"""
    cfgname, seed, n = args
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    from emu import Emu
    from oracle.oracle import Oracle
    emu, o = Emu("libbgemu_fast.so"), Oracle()
    cfg, Gd, Hd, samples, k, desc = load_config(cfgname)
    t = cfg["t"]
    exact = cfg["exact"] if k == 0 else 0
    L = [] if exact else fixed_L(k, t)
    from oracle.oracle import Projector as OP
    G = OP.make(*Gd)
    terms = [o.Lbits(i, L) for i in range(1 << len(L))] if not exact else None
    pairs = 0
    emu.lib.emu_chi_seconds(1)
    t0 = time.perf_counter()
    for s in range(n):
        th = o.random_state_philox(t, seed, 0, s)
        if exact:
            size = (t + 1) // 2
            tt = [sum(((i >> (size - 1 - j)) & 1) << (2 * j) for j in range(size)) for i in range(1 << size)]
            got = emu.terms(th, G, 1, 1, t, tt, tpp=True)
            pairs += len(tt) if got["alive"] else 0
        else:
            got = emu.terms(th, G, 1, 0, t, terms, tpp=True)
            pairs += len(terms) if got["alive"] else 0
    return pairs, time.perf_counter() - t0, emu.lib.emu_chi_seconds(1)


def cpu_baseline_packed(cfgname, per_core=300):
    """The honest "optimised CPU" row: the SAME packed-row algorithm the GPU runs (the product's thread-per-pair
    device source, compiled for the host by tests/emu — test infrastructure, timed here as a reported baseline
    only), one process per host core, bounded sample."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    try:
        with mp.get_context("spawn").Pool(cores) as pool:
            t0 = time.perf_counter()
            res = pool.map(_packed_worker, [(cfgname, 100 + c, per_core) for c in range(cores)])
            dt = time.perf_counter() - t0
        pairs = sum(r[0] for r in res)
        busy = max(r[1] for r in res)
        chi_busy = max(r[2] for r in res)
        return {"value": pairs / chi_busy, "unit": "inner products/s", "cores": cores, "kind": "port",
                "value_with_theta_draw_and_projection_under_the_warp_emulator": pairs / busy,
                "sample": "%d processes x %d samples x 1 projector x all terms; %.2f s in the chi loop per process (k_pairs_tpp's "
                          "thread-per-pair algorithm, bg_tpp.cuh compiled for the host with -O3 -march=x86-64-v3, scalar; the per-sample "
                          "theta draw + projection runs under the 32-fibre warp emulator and is timed separately: %.2f s in all)"
                          % (cores, per_core, chi_busy, busy)}
    except Exception as ex:
        return {"error": repr(ex)[:300]}


# ------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default=DEFAULT_CONFIG)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-probability", action="store_true")
    args = ap.parse_args()
    cfgname = args.config
    if args.impl == "reference":
        return run_reference(args, cfgname)

    import numpy as np
    import torch
    import torch.distributed as dist
    import circuitsimulator_b200 as bg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU path")
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")          # host-side waits (see the probability() leg)

    cfg, Gd, Hd, samples, k, desc = load_config(cfgname)
    t = cfg["t"]
    exact = cfg["exact"] if k == 0 else 0
    L = [] if exact else fixed_L(k, t)
    chi = (1 << ((t + 1) // 2)) if exact else (1 << k)
    G = bg.Projector.make(*Gd)
    H = bg.Projector.make(*Hd)

    # one context per GPU; both projectors of a probability() evaluation form one prepared job.
    # Work runs on a non-default torch stream (a CUDA graph cannot be captured on the legacy stream).
    ctx = bg.Backend(local)
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    if world > 1:
        uid = [ctx.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
    ctx.set_stream(tstream.cuda_stream)
    ctx.set_shard(rank, world)
    if world > 1:
        ctx.nccl_join(uid[0])
    ctx.set_decomposition(t, exact, L)
    lop3_peak, popc_peak = ctx.measure_int_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allsum(x):
        v = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
        return float(v.item())

    def allmax(x):
        v = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    # ---- device-resident arm: projectors + decomposition uploaded once; the step replays one CUDA graph
    # (theta draw + projection, L x chi loop, reduction for G' then H'), then one NCCL all-reduce of the
    # partial sums and a 16-byte read-back.
    ctx.sampled_prepare2(G, H, samples, 1, 1001, 1002)
    results = []

    def step_resident():
        ctx.sampled_run()
        results.append(ctx.sampled_finish2(1.0))

    sampler = ClockSampler(local) if rank == 0 else None      # one nvidia-smi poller per job, not per rank
    if sampler:
        sampler.start()              # samples through warm-up + timed region (all under the same load)
    for _ in range(max(3, args.warmup)):
        step_resident()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = {"kernel_ms": 0.0, "prepare_ms": 0.0, "pairs_ms": 0.0, "launches": 0, "pair_launches": 0, "pairs": 0}

    def account():
        st = ctx.stats()
        for key in acc:
            acc[key] += st[key]

    def timed_block(nsteps, do_account):
        """nsteps steps, two in flight: the all-reduce + 16-byte read-back of step i overlap the kernels of
        step i+1 (every step's result is delivered inside the timed region).  Returns max-over-ranks ms."""
        barrier()
        e0.record(tstream)
        ctx.sampled_run()
        for _ in range(nsteps - 1):
            ctx.sampled_run()
            results.append(ctx.sampled_finish2(1.0))
            if do_account:
                account()
        results.append(ctx.sampled_finish2(1.0))
        if do_account:
            account()
        e1.record(tstream)
        barrier()
        return allmax(e0.elapsed_time(e1))

    timed_block(max(3, args.warmup), False)                # warm-up in the timed pattern too (two jobs in flight; in overlap
                                                           # mode the second slot's stream, buffers and graph see their first use)
    # THE timed region — exactly K steps between barrier + synchronize, max over ranks — five times; `value` is the
    # median block (one block of K steps at N = 8 is 40 ms: a single stall of one rank moves it by several per cent).  The
    # first block also reads the library's per-kernel CUDA-event times after every step (the roofline's kernel_ms).
    first_ms = timed_block(args.steps, True)
    # the pairs the device actually evaluated (annihilated samples evaluate none), all ranks
    pairs_total = allsum(acc["pairs"])
    pairs_per_step = pairs_total / args.steps
    block_ms = [first_ms] + [timed_block(args.steps, False) for _ in range(4)]
    ms = sorted(block_ms)[len(block_ms) // 2]
    value = pairs_total / (ms * 1e-3)
    # a timed region of a few ms is shorter than one nvidia-smi poll: keep the same load running
    # (untimed; the same number of steps on every rank) so that the sampler sees clocks under load
    if sum(block_ms) < 1500.0:
        for _ in range(int(1500.0 / max(ms / args.steps, 0.05)) + 1):
            step_resident()
    if sampler:
        sampler.stop_flag = True

    # ---- end-to-end arm: the C-ABI calls a host makes per probability(): decomposition + both projectors from
    # HOST memory (the projectors are copied host -> pinned staging -> device every step; the decomposition is
    # recognised as unchanged and kept), the kernels, the all-reduce, the result back in host memory — every step.
    # Every step's projectors really are new bytes to upload: a host copy whose unused last generator slot carries a
    # step counter (the kernels read nstabs generators; the library compares the bytes with what the device holds, and
    # in overlap mode there are two device copies used in turn — two alternating variants would never change either).
    Gv = bg.Projector.from_buffer_copy(G)
    e2e_serial = [0]

    def fresh_G():
        e2e_serial[0] += 1
        Gv.xs[bg.MAX_STABS - 1] = e2e_serial[0]
        return Gv

    def step_e2e(i):
        ctx.set_decomposition(t, exact, L)
        return ctx.sampled_norm2(fresh_G(), H, samples, 1, 1001, 1002, 1.0)

    def submit_e2e(i):
        # the split-phase form of the same call (bg_sampled_prepare2 + bg_sampled_run): this step's projectors are
        # staged and uploaded, its kernels, all-reduce and read-back enqueued; bg_sampled_finish2 delivers the result
        ctx.set_decomposition(t, exact, L)
        ctx.sampled_prepare2(fresh_G(), H, samples, 1, 1001, 1002)
        ctx.sampled_run()

    for i in range(3):
        r_e2e = step_e2e(i)
    barrier()
    w0 = time.perf_counter()
    for i in range(args.steps):
        r_e2e = step_e2e(i)
    torch.cuda.synchronize()
    wall_sync = allmax(time.perf_counter() - w0)
    # two calls in flight, as a host with independent probability() evaluations to make (bins, circuits, the requests
    # of several clients of the served back end) issues them: the upload, all-reduce and read-back of step i overlap
    # the kernels of step i+1; every step's inputs go up and every step's result comes back inside the timed region
    submit_e2e(0)
    submit_e2e(1)
    ctx.sampled_finish2(1.0)
    ctx.sampled_finish2(1.0)
    barrier()
    w0 = time.perf_counter()
    submit_e2e(0)
    for i in range(1, args.steps):
        submit_e2e(i)
        r_pipe = ctx.sampled_finish2(1.0)
    r_pipe = ctx.sampled_finish2(1.0)
    torch.cuda.synchronize()
    wall = allmax(time.perf_counter() - w0)
    if tuple(r_pipe) != tuple(r_e2e):
        raise SystemExit("pipelined and synchronous end-to-end calls disagree: %r vs %r" % (r_pipe, r_e2e))
    st = ctx.stats()
    e2e_value = pairs_per_step * args.steps / wall
    h2d = int(st["h2d_bytes"])
    if h2d <= 0:
        raise SystemExit("the end-to-end arm did not upload its step's projectors (h2d_bytes = 0): not an end-to-end number")
    d2h = int(st["d2h_bytes"])

    # ---- end-to-end with a NEW decomposition every step (what sampleQubits does: a fresh random L per probability()
    # call): term tables rebuilt, sorted, planned and uploaded, graph re-captured
    Ls = [L, [x ^ 1 for x in L]] if L else [L, L]

    def step_fresh(i):
        ctx.set_decomposition(t, exact, Ls[i & 1])
        return ctx.sampled_norm2(G, H, samples, 1, 1001, 1002, 1.0)

    nfresh = max(2, min(args.steps, 10))
    for i in range(2):
        step_fresh(i)
    barrier()
    w0 = time.perf_counter()
    for i in range(nfresh):
        step_fresh(i)
    torch.cuda.synchronize()
    fresh_wall = allmax(time.perf_counter() - w0)
    ctx.set_decomposition(t, exact, L)

    # ---- the other configurations of BASELINE.json (a few steps each; rank 0's GPU only would do, but every rank
    # takes its shard so that the all-reduce path is the same)
    others = other_configs(ctx, bg, lop3_peak, cfgname) if not args.no_other_configs else []
    ctx.set_decomposition(t, exact, L)

    # ---- metric half 2: probability() seconds through the drop-in back end on `world` GPUs (rank 0 drives it)
    prob_s = None
    barrier()
    if rank == 0 and not args.no_probability:
        prob_s = probability_seconds(t, samples, k, exact, Gd, Hd, world)
        if prob_s is not None:
            prob_s["sample_qubits_phase_estimation"] = sample_qubits_seconds(world)
    if cpu_group is not None:
        # the other ranks wait on the HOST: an NCCL barrier would keep a spinning kernel on their GPUs, which are the
        # GPUs the back end under test is running on (measured: 7.9 instead of 5.0 ms per served call at N = 2)
        dist.barrier(group=cpu_group)
    barrier()

    if rank == 0:
        clocks = sampler.summary() if sampler else None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        wpath = os.path.join(ROOT, "profiles", "work_model.json")
        wm = json.load(open(wpath)).get(cfgname, {}) if os.path.exists(wpath) else {}
        lane_ops = wm.get("alu_lane_ops_per_pair")           # DESIGN.md "Roofline": algorithmic lane-ops / pair
        # one pair-kernel launch = this rank's samples of BOTH projectors (fused job), CUDA events around it
        launches_per_step = max(1, acc["pair_launches"]) / args.steps
        per_launch_pairs = pairs_per_step / world / launches_per_step
        k_ms = acc["pairs_ms"] / max(1, acc["pair_launches"])
        roof = {"bound": "int_alu", "unit": "Tlaneop/s",
                "kernel": pair_kernel_name(t, exact, k), "kernel_ms": k_ms, "launches_per_step": launches_per_step,
                "pairs_per_launch": per_launch_pairs,
                "kernel_share_of_step": launches_per_step * k_ms / (ms / args.steps),
                "prepare_ms": acc["prepare_ms"] / max(1, acc["pair_launches"]),
                "achieved": (per_launch_pairs * lane_ops / (k_ms * 1e-3) / 1e12) if lane_ops and k_ms > 0 else None,
                "peak": lop3_peak / 1e12,
                "peak_source": "LOP3 lane-ops/s measured in this run by bg_measure_int_peak (64 lanes/clk/SM x 148 SMs)",
                "popc_peak": popc_peak / 1e12,
                "lane_ops_per_pair": lane_ops,
                "lane_ops_source": "profiles/work_model.json: the algorithm's own operation count per inner product, frozen in round 1",
                "traffic": None,
                "hbm_gbs_peak": peaks.get("hbm_gbs"),
                # NOT measured in this run: constants read from the committed ncu capture of the same kernel at N=1
                "from_profiles": wm.get("from_profiles"),
                # overlap mode (few samples per GPU: the draw + projection kernel of step i+1 runs beside the pair kernel
                # of step i on a second stream): kernel_ms / prepare_ms are then event intervals of kernels that share the GPU
                "overlap_mode": bool(ctx.stats().get("overlapped", 0))}
        fp = wm.get("from_profiles") or {}
        if fp.get("dram_bytes_per_launch") and fp.get("pairs_per_launch_ncu"):
            roof["traffic"] = fp["dram_bytes_per_launch"] * per_launch_pairs / fp["pairs_per_launch_ncu"]
            roof["hbm_gbs_achieved"] = roof["traffic"] / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
        roof["frac"] = (roof["achieved"] / roof["peak"]) if roof["achieved"] else None
        block_ms.sort()
        line = {"metric": "stabilizer inner products/sec", "value": value, "unit": "inner products/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u64 bit rows + Z8 integer phases (int64 exact accumulation, fp64 final)",
                "data": "synthetic",
                "config": config_dict(cfgname, desc, t, chi, samples),
                "pairs_per_step": pairs_per_step,
                "blocks": {"n": len(block_ms), "steps_each": args.steps,
                           "ms_per_step": [b / args.steps for b in block_ms],
                           "value_is": "the median block (every block is exactly K timed steps)",
                           "first_block_ms_per_step": first_ms / args.steps,
                           "median_value": pairs_per_step * args.steps / (block_ms[len(block_ms) // 2] * 1e-3)},
                "e2e": {"value": e2e_value, "unit": "inner products/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h,
                        "how": "host buffers through the C ABI every step (bg_set_decomposition, bg_sampled_prepare2: "
                               "projectors host -> pinned -> device, bg_sampled_run, bg_sampled_finish2: result in host "
                               "memory), two calls in flight",
                        "one_call_at_a_time_value": pairs_per_step * args.steps / wall_sync,
                        "fresh_decomposition_value": pairs_per_step * nfresh / fresh_wall,
                        "fresh_decomposition_ms_per_step": 1e3 * fresh_wall / nfresh},
                "gpu_launches": acc["launches"],
                "roofline": roof,
                "clocks": clocks,
                "probability_s": prob_s,
                "other_configs": others,
                "result": {"numerator": results[-1][0], "denominator": results[-1][1],
                           "e2e_numerator": r_e2e[0], "e2e_denominator": r_e2e[1]}}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(cfgname)
            line["cpu_baseline_packed"] = cpu_baseline_packed(cfgname)
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
