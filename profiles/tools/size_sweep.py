#!/usr/bin/env python
"""Pair-kernel time of config 4 against the number of samples (fixed cost of a launch, wave quantisation):
    python profiles/tools/size_sweep.py [samples ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import circuitsimulator_b200 as bg  # noqa: E402

cfg, Gd, Hd, samples, k, desc = bench.load_config(bench.DEFAULT_CONFIG)
t = cfg["t"]
L = bench.fixed_L(k, t)
G, H = bg.Projector.make(*Gd), bg.Projector.make(*Hd)
be = bg.Backend(0)
be.set_decomposition(t, False, L)
sizes = [int(x) for x in sys.argv[1:]] or [1110, 2220, 3330, 4440, 6660, 8192, 8880, 16384, 65536]
for n in sizes:
    be.sampled_prepare2(G, H, n, 1, 1, 2)
    for _ in range(3):
        be.sampled_run(); be.sampled_finish2(1.0)
    tp = tq = 0.0
    for _ in range(10):
        be.sampled_run(); be.sampled_finish2(1.0)
        st = be.stats()
        tp += st["pairs_ms"]; tq += st["prepare_ms"]
    print("samples/projector %6d  pairs %.4f ms  (%.2f ns per sample)  prepare %.4f ms" % (n, tp / 10, 1e6 * tp / 10 / (2 * n), tq / 10), flush=True)
