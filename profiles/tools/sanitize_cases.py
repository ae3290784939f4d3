#!/usr/bin/env python
"""Small invocations of every hot-path kernel for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python profiles/tools/sanitize_cases.py
Cases: smoke() (t=16 exact decomposition, k_prepare + k_pairs_tpp<u32>), config 4 at 192 samples (k_pairs_shb, TMA
staging, warp-shared reduced forms), the same with BG_SHB=0 (k_pairs_tpp<u64>), low-dimensional thetas (MANYC
instantiation: pivot history in shared memory), the exact-norm path (tri mode), the warp-per-pair kernel; the draw +
projection one thread per sample (k_prepare_tps: 64-thread CTAs, 32- and 64-bit words; the default of every sampled job
above), one warp per sample (BG_PREP=warp), and overlap mode (BG_OVERLAP=1: 32-thread CTAs, two jobs in flight on two
streams with new projectors staged while a job runs)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import circuitsimulator_b200 as bg  # noqa: E402
import __graft_entry__ as ge  # noqa: E402
from oracle.oracle import Oracle, states_to_numpy  # noqa: E402

ge.smoke()
cfg, Gd, Hd, samples, k, desc = bench.load_config(bench.DEFAULT_CONFIG)
t = cfg["t"]
L = bench.fixed_L(k, t)
G, H = bg.Projector.make(*Gd), bg.Projector.make(*Hd)
for env in ({}, {"BG_SHB": "0"}, {"BG_KERNEL": "warp"}, {"BG_PREP": "warp"}):
    os.environ.update(env)
    be = bg.Backend(0)
    for key in env:
        del os.environ[key]
    be.set_decomposition(t, False, L)
    n = 32 if env.get("BG_KERNEL") else 192
    print(env, "sampled_norm2:", be.sampled_norm2(G, H, n, 1, 3, 4, 1.0), be.stats()["pairs"], "pairs", flush=True)
    be.close()
os.environ["BG_OVERLAP"] = "1"
be = bg.Backend(0)
del os.environ["BG_OVERLAP"]
be.set_decomposition(t, False, L)
got = []
be.sampled_prepare2(G, H, 100, 1, 3, 4)
be.sampled_run()
for a, b in ((H, G), (G, G), (G, H)):
    be.sampled_prepare2(a, b, 100, 1, 3, 4)
    be.sampled_run()
    got.append(be.sampled_finish2(1.0))
got.append(be.sampled_finish2(1.0))
print("overlap mode, two jobs in flight:", got, be.stats()["overlapped"], flush=True)
be.close()
be = bg.Backend(0)
o = Oracle()
tt = 12
rs = np.random.RandomState(3)
L2 = [int(rs.randint(0, 1 << tt)) for _ in range(4)]
be.set_decomposition(tt, False, L2)
thetas = []
for j in range(40):
    s = o.random_state_philox(tt, 9, 0, j)
    for _ in range(j % (tt + 1)):
        o.measure_pauli(s, 0, int(rs.randint(1, 1 << tt)), 0)
    thetas.append(s)
out = be.sampled_norm_from_states(bg.Projector.make(tt, [], [], []), states_to_numpy(thetas), project=False)
print("low-dimensional thetas (MANYC):", out["mean"], flush=True)
cfg2, G2, H2 = bench.parse_stream(os.path.join(bench.STREAMS, "toffoli_q0.txt"))
be.set_decomposition(cfg2["t"], True, [])
print("exact norm t=16:", be.exact_norm(bg.Projector.make(*G2), 1.0), flush=True)
be.close()
print("sanitize cases done")
