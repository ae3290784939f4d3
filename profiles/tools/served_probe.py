#!/usr/bin/env python
"""Wall time of a served probability() call (config 4 stream through `bgbackend --serve`) against the number of GPUs,
the reduction (NCCL in the server / host-side add) and the sample count:  python profiles/tools/served_probe.py"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import circuitsimulator_b200 as bg  # noqa: E402
import torch  # noqa: E402

cfg, Gd, Hd, samples, k, desc = bench.load_config(bench.DEFAULT_CONFIG)
t = cfg["t"]
ngpu = torch.cuda.device_count()
for gpus in [g for g in (1, 2, 4, 8) if g <= ngpu]:
    for red in ("nccl", "host"):
        if gpus == 1 and red == "host":
            continue
        for n in (samples, samples // 8):
            text = bench.stream_text(t, n, k, 0, Gd, Hd)
            env = {"BG_GPUS": gpus, "BG_SEED": 7, "BG_REDUCE": red}
            sock = os.path.join(tempfile.mkdtemp(prefix="bgsp"), "s")
            e = dict(os.environ); e.update({k_: str(v) for k_, v in env.items()})
            srv = subprocess.Popen([bg.BACKEND_PATH, "--serve", sock], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            try:
                srv.stdout.readline()
                cenv = dict(env); cenv["BG_SERVER"] = sock
                ts = []
                for _ in range(8):
                    t0 = time.perf_counter()
                    num, den, _l = bg.run_backend(text, env=cenv, timeout=300)
                    ts.append(time.perf_counter() - t0)
                print("gpus %d  reduce %s  samples %6d: served call min %.2f ms  median %.2f ms  (%.3e / %.3e)"
                      % (gpus, red, n, 1e3 * min(ts[1:]), 1e3 * sorted(ts[1:])[len(ts) // 2 - 1], num, den), flush=True)
            finally:
                bench.stop_server(srv, sock)
