#!/usr/bin/env python
"""Step time of config 4 with two jobs in flight, with and without overlap mode (BG_OVERLAP: the odd slot's jobs on a
stream and buffers of their own, so that k_prepare_tps of job i+1 runs beside the pair kernel of job i):
    python profiles/tools/overlap_probe.py [samples per projector ...]"""
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
G, H = bg.Projector.make(*Gd), bg.Projector.make(*Hd)
sizes = [int(x) for x in sys.argv[1:]] or [4096, 8192, 16384, 65536]
for n in sizes:
    ref = None
    for mode in ("0", "1"):
        os.environ["BG_OVERLAP"] = mode
        be = bg.Backend(0)
        be.set_decomposition(t, False, L)
        be.sampled_prepare2(G, H, n, 1, 1, 2)
        for _ in range(4):
            be.sampled_run()
            out = be.sampled_finish2(1.0)
        ref = ref or out
        assert out == ref, (out, ref)
        steps = 200
        t0 = time.perf_counter()
        be.sampled_run()
        for _ in range(steps - 1):
            be.sampled_run()
            out = be.sampled_finish2(1.0)
            assert out == ref
        out = be.sampled_finish2(1.0)
        dt = (time.perf_counter() - t0) / steps
        st = be.stats()
        print("samples/projector %6d  BG_OVERLAP=%s  %.4f ms per step  (pairs %.4f ms, prepare %.4f ms by events)"
              % (n, mode, 1e3 * dt, st["pairs_ms"], st["prepare_ms"]), flush=True)
        be.close()
