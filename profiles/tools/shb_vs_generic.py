#!/usr/bin/env python
"""GPU: per-sample values of the shared high-block kernel (k_pairs_shb) against the generic 64-bit kernel
(BG_SHB=0) on the same device-drawn thetas, config 4 (t=40, k=9, bench L).  Prints the samples that differ.
    python profiles/tools/shb_vs_generic.py [samples]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import circuitsimulator_b200 as bg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cfg, Gd, Hd, samples, k, desc = bench.load_config(bench.DEFAULT_CONFIG)
t = cfg["t"]
L = bench.fixed_L(k, t)
G, H = bg.Projector.make(*Gd), bg.Projector.make(*Hd)
a = bg.Backend(0)
os.environ["BG_SHB"] = "0"
b = bg.Backend(0)
del os.environ["BG_SHB"]
a.set_decomposition(t, False, L)
b.set_decomposition(t, False, L)
for which, (P, seed) in enumerate(((G, 1), (H, 2))):
    for first in range(0, n, 8192):
        cnt = min(8192, n - first)
        th = a.random_states(t, seed, 0, first, cnt)
        ra = a.sampled_norm_from_states(P, th, project=True)
        rb = b.sampled_norm_from_states(P, th, project=True)
        d = np.nonzero(ra["per_sample"] != rb["per_sample"])[0]
        for i in d[:20]:
            print("projector", which, "sample", first + i, "shb", ra["per_sample"][i], "generic", rb["per_sample"][i],
                  "k", int(th[i]["k"]) if "k" in th.dtype.names else "?")
        print("projector", which, "samples", first, "..", first + cnt, "differ:", len(d), flush=True)
