#!/usr/bin/env python
"""Row passes per inner product of the thread-per-pair algorithm, from the CPU build of the kernel source
(tests/emu, BG_TRACE): how many passes over the thread's rows a pair needs, of which kind (2 masks: a
parity-check pivot; 6 masks: the first pass of the rounds with the fold; 4 masks: two elimination steps)
and how many rows each touches.  CPU only — the numbers behind DESIGN.md section 9 item 1.
    python profiles/tools/pass_model.py [config]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from emu.emu import Emu  # noqa: E402
from util import parse_stream, GOLDEN  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "hidden_shift_n40_t40_k9_L65536"
stream, samples, k, desc = bench.CONFIGS[name]
cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", stream))
t = cfg["t"]
exact = cfg["exact"] if k == 0 else 0
L = [] if exact else bench.fixed_L(k, t)
o, e = Oracle(), Emu()
if exact:
    size = (t + 1) // 2
    terms = [sum(((i >> (size - 1 - j)) & 1) << (2 * j) for j in range(size)) for i in range(1 << size)]
else:
    terms = [o.Lbits(i, L) for i in range(1 << len(L))]
# the order the pair kernel uses (bg_set_decomposition): by popcount, then by the popcount of the low half
terms.sort(key=lambda x: -((bin(x).count("1") << 8) | bin(x & 0xffffffff).count("1")))
cap = 1 << 22
buf = (C.c_int * cap)()
e.lib.emu_trace.argtypes = [C.POINTER(C.c_int), C.c_int]
pairs = []
batches = []          # per (sample, 32 consecutive terms): block-rounds of each lane
for P, seed in ((G, 1001), (H, 1002)):
    for l in range(4 if t > 20 else 16):
        th = o.random_state_philox(t, seed, 0, l)
        e.lib.emu_trace(buf, cap)
        got = e.terms(th, P, 1, exact, t, terms, tpp=True)
        n = e.lib.emu_trace_len()
        e.lib.emu_trace(None, 0)
        if not got["alive"]:
            continue
        cur, mine = [], []
        for j in range(0, n, 2):
            lo, hi = buf[j], buf[j + 1]
            if lo == -1:
                pairs.append(cur)
                mine.append(cur)
                cur = []
            else:
                cur.append((lo // 1000, lo % 1000 + hi))
        for b in range(0, len(mine), 32):
            batches.append([sum(1 for k_, r in p if k_ in (4, 6)) for p in mine[b:b + 32]])
kinds = {2: "parity-check pivot (2 masks)", 6: "first pass of the rounds, with the fold (6 masks)", 4: "two steps (4 masks)"}
out = {"config": name, "pairs": len(pairs), "per_pair": {}}
for kind, label in kinds.items():
    cnt = [sum(1 for k_, r in p if k_ == kind) for p in pairs]
    nonempty = [sum(1 for k_, r in p if k_ == kind and r > 0) for p in pairs]
    rows = [sum(r for k_, r in p if k_ == kind) for p in pairs]
    out["per_pair"][label] = {"passes": float(np.mean(cnt)), "passes_touching_rows": float(np.mean(nonempty)),
                              "rows_touched": float(np.mean(rows))}
tot = [sum(1 for k_, r in p if r > 0) for p in pairs]
out["per_pair"]["all"] = {"passes_touching_rows": float(np.mean(tot)), "rows_touched": float(np.mean([sum(r for _, r in p) for p in pairs]))}
# lanes of a warp run their block-rounds in lock step: the warp needs max(lane rounds), a lane is busy for its own
busy = sum(sum(b) for b in batches)
slots = sum(len(b) * max(b) for b in batches if b)
out["block_rounds"] = {"per_pair": busy / max(1, len(pairs)), "per_warp_batch_max": float(np.mean([max(b) for b in batches if b])),
                       "lanes_busy_fraction": busy / max(1, slots)}
print(json.dumps(out, indent=1))
