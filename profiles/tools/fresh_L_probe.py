#!/usr/bin/env python
"""Host cost of a NEW decomposition per probability() call (what sampleQubits does): bg_set_decomposition with another
L of the same (t, k), then the job.      python profiles/tools/fresh_L_probe.py [samples per projector]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import circuitsimulator_b200 as bg  # noqa: E402

cfg, Gd, Hd, samples, k, desc = bench.load_config(bench.DEFAULT_CONFIG)
t = cfg["t"]
L = bench.fixed_L(k, t)
import random
rnd = random.Random(7)
Ls = [L, [x ^ 1 for x in L]] + [[rnd.getrandbits(t) for _ in range(k)] for _ in range(62)]      # mostly random L, as decompose() draws them
G, H = bg.Projector.make(*Gd), bg.Projector.make(*Hd)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
be = bg.Backend(0)
for i in range(4):
    be.set_decomposition(t, False, Ls[i % len(Ls)])
    be.sampled_norm2(G, H, n, 1, 1, 2, 1.0)
reps = 200
ts = tr = 0.0
for i in range(reps):
    t0 = time.perf_counter()
    be.set_decomposition(t, False, Ls[i % len(Ls)])
    t1 = time.perf_counter()
    out = be.sampled_norm2(G, H, n, 1, 1, 2, 1.0)
    t2 = time.perf_counter()
    ts += t1 - t0
    tr += t2 - t1
same = 0.0
for i in range(reps):
    t1 = time.perf_counter()
    be.set_decomposition(t, False, Ls[1])
    out = be.sampled_norm2(G, H, n, 1, 1, 2, 1.0)
    same += time.perf_counter() - t1
print("samples/projector %d: new L: set_decomposition %.1f us + job %.1f us;  same L: call %.1f us"
      % (n, 1e6 * ts / reps, 1e6 * tr / reps, 1e6 * same / reps))
