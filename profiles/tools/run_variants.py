#!/usr/bin/env python
"""Scratch harness (not product): time the pair kernel of config 4 for every library variant in profiles/tools/libs/ (listed with their BG_* environment in profiles/tools/variants.json).
usage: python profiles/tools/run_variants.py [config] [steps]      env passes through (BG_FUSE2, BG_ITEMS_FACTOR, BG_LAZY, ...)"""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import sys, os, json
sys.path.insert(0, %(root)r)
import circuitsimulator_b200 as bg
bg.LIB_PATH = %(lib)r
import bench
cfgname = %(cfg)r
cfg, Gd, Hd, samples, k, desc = bench.load_config(cfgname)
t = cfg["t"]; exact = cfg["exact"] if k == 0 else 0
L = [] if exact else bench.fixed_L(k, t)
chi = (1 << ((t + 1) // 2)) if exact else (1 << k)
G = bg.Projector.make(*Gd); H = bg.Projector.make(*Hd)
ctx = bg.Backend(0)
ctx.set_decomposition(t, exact, L)
ctx.sampled_prepare2(G, H, samples, 1, 1001, 1002)
res = None
for _ in range(3):
    ctx.sampled_run(); res = ctx.sampled_finish2(1.0)
pm = []; pr = []; km = []
for _ in range(%(steps)d):
    ctx.sampled_run(); res = ctx.sampled_finish2(1.0)
    st = ctx.stats(); pm.append(st["pairs_ms"] / 2); pr.append(st["prepare_ms"] / 2); km.append(st["kernel_ms"])
pm.sort(); pr.sort(); km.sort()
print(json.dumps({"lib": os.path.basename(%(lib)r), "pairs_ms_med": pm[len(pm)//2], "pairs_ms_min": pm[0], "prepare_ms_med": pr[len(pr)//2],
                  "step_kernel_ms_med": km[len(km)//2], "num": res[0], "den": res[1], "env": {k_: v for k_, v in os.environ.items() if k_.startswith("BG_")}}))
ctx.close()
'''
cfg = sys.argv[1] if len(sys.argv) > 1 else "hidden_shift_n40_t40_k9_L65536"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
spec = os.path.join(ROOT, "profiles", "tools", "variants.json")
if os.path.exists(spec):
    variants = [(os.path.join(ROOT, "profiles", "tools", "libs", v["lib"]), v.get("env", {})) for v in json.load(open(spec))]
else:
    variants = [(l, {}) for l in sorted(glob.glob(os.path.join(ROOT, "profiles", "tools", "libs", "*.so")))]
for lib, env in variants:
    e = dict(os.environ); e.update({k: str(v) for k, v in env.items()})
    p = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "lib": lib, "cfg": cfg, "steps": steps}], capture_output=True, text=True, env=e)
    print(p.stdout.strip() or ("FAILED %s: %s" % (lib, p.stderr[-800:])), flush=True)
