#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the JSON kept under profiles/: per kernel the launch
duration, instruction counts, pipe utilisation, occupancy limits, DRAM bytes, the warp-stall sampling
breakdown and the SASS opcode mix.      python profiles/summarize_ncu.py <report.ncu-rep> [out.json] [note]"""
import csv
import io
import json
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.sum.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = {"what": sys.argv[3] if len(sys.argv) > 3 else "", "report": rep.split("/")[-1], "kernels": []}
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        k = {"Kernel Name": r[hdr.index("Kernel Name")]}
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                k[key] = "%s %s" % (r[i], units[i])
        out["kernels"].append(k)
    # source page of the (first) kernel: stall sampling + opcode mix
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    h = None
    data = []
    for r in src:
        if len(r) > 5 and r[0] == "Address":
            h = r
            continue
        if h and len(r) == len(h):
            data.append(r)
    if h and data:
        ix = {k: i for i, k in enumerate(h)}
        tot_s = sum(int(r[ix["# Samples"]]) for r in data) or 1
        tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data) or 1
        stalls = {k: sum(int(r[ix[k]]) for r in data) for k in h if k.startswith("stall_") and "Not Issued" not in k}
        out["warp_stall_sampling_pct"] = {k: round(100.0 * v / tot_s, 2) for k, v in sorted(stalls.items(), key=lambda x: -x[1]) if v * 200 > tot_s}
        ops, lanes = Counter(), Counter()
        for r in data:
            s = r[ix["Source"]].split()
            op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
            ops[op] += int(r[ix["Instructions Executed"]])
            lanes[op] += int(r[ix["Thread Instructions Executed"]])
        out["sass_opcode_mix_pct_of_warp_instructions"] = {op: [round(100.0 * v / tot_i, 2), round(lanes[op] / max(v, 1), 1)] for op, v in ops.most_common(16)}
        out["sass_opcode_mix_note"] = "[share of executed warp instructions, average active lanes]"
    txt = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()
