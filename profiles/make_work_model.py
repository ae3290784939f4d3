#!/usr/bin/env python
"""Algorithmic work per inner product of the thread-per-pair algorithm (bg_tpp.cuh), counted by the
CPU build of the same source (-DBG_COUNT_WORK) on the BASELINE workloads, device-RNG thetas.

lane-op model (32-bit integer ops the ALGORITHM needs, independent of how the kernel is scheduled):
    W = words * ( xors                      one XOR per (row, mask) application
                + 20 * dimers               mask algebra of a dimer round (J_a, J_b, rest, D2, Js, signs)
                +  6 * monomers
                + 14 * basis_changes        fold / parity-check pivot: column mask, D1/D2/Q updates
                +  t + 8 )                  per pair: working copy of the t ambient rows, term masks
words = 1 for t <= 32, 2 for t <= 64.   Writes profiles/work_model.json (read by bench.py)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle.oracle import Oracle
from emu.emu import Emu
from util import parse_stream, GOLDEN
import bench

o, e = Oracle(), Emu()
out = {}
for name, (stream, samples, k, desc) in bench.CONFIGS.items():
    cfg, G, H = parse_stream(os.path.join(GOLDEN, "streams", stream))
    t = cfg["t"]
    exact = cfg["exact"] if k == 0 else 0
    L = [] if exact else bench.fixed_L(k, t)
    if exact:
        size = (t + 1) // 2
        terms = [sum(((i >> (size - 1 - j)) & 1) << (2 * j) for j in range(size)) for i in range(1 << size)]
    else:
        terms = [o.Lbits(i, L) for i in range(1 << len(L))]
    e.work_counters(reset=True)
    ns = 6 if t > 20 else 24
    for P, seed in ((G, 1001), (H, 1002)):
        for l in range(ns):
            th = o.random_state_philox(t, seed, 0, l)
            e.terms(th, P, 1, exact, t, terms, tpp=True)
    w = e.work_counters(reset=True)
    words = 1 if t <= 32 else 2
    n = max(1, w["pairs"])
    per = {k_: v / n for k_, v in w.items() if k_ != "pairs"}
    W = words * (per["xors"] + 20 * per["dimers"] + 6 * per["monomers"] + 14 * per["basis_changes"] + t + 8)
    out[name] = {"t": t, "chi": len(terms), "words": words, "pairs_counted": w["pairs"],
                 "per_pair": per, "alu_lane_ops_per_pair": W}
    print(name, json.dumps(out[name]))
path = os.path.join(ROOT, "profiles", "work_model.json")
old = json.load(open(path)) if os.path.exists(path) else {}
for k_, v in out.items():
    old.setdefault(k_, {}).update(v)
json.dump(old, open(path, "w"), indent=1, sort_keys=True)
