#!/usr/bin/env python
"""Basic-block view of an ncu source page: executed warp-instructions per SASS basic block (split at
BRA / BSYNC / EXIT), with the average number of active lanes and the opcode sequence — how the loops and
the straight-line parts of k_pairs_tpp share the issue slots.
    python profiles/tools/hot_blocks.py <report.ncu-rep> [n_blocks]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h = next(r for r in rows if len(r) > 5 and r[0] == "Address")
    data = [r for r in rows if len(r) == len(h) and r[0] != "Address"]
    ix = {k: i for i, k in enumerate(h)}
    tot = sum(int(r[ix["Instructions Executed"]]) for r in data) or 1
    blocks, cur = [], []
    for n, r in enumerate(data):
        src = r[ix["Source"]].strip()
        cur.append((n, src, int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]])))
        if " BRA" in " " + src or "BSYNC" in src or "EXIT" in src:
            blocks.append(cur)
            cur = []
    if cur:
        blocks.append(cur)
    print("kernel: %s" % next((r[1] for r in rows if r and r[0] == "Kernel Name"), "?"))
    print("executed warp-instructions: %d in %d SASS instructions, %d basic blocks" % (tot, len(data), len(blocks)))
    print("%7s %6s %5s %6s  %s" % ("share", "cum", "instr", "lanes", "first SASS index | opcodes"))
    acc = 0
    for b in sorted(blocks, key=lambda b: -sum(x[2] for x in b))[:top]:
        e = sum(x[2] for x in b)
        t = sum(x[3] for x in b)
        acc += e
        ops = " ".join((x[1].split()[1] if x[1].startswith("@") else x[1].split()[0]).split(".")[0] for x in b[:30])
        print("%6.2f%% %5.1f%% %5d %6.1f  %d | %s%s" % (100.0 * e / tot, 100.0 * acc / tot, len(b), t / max(e, 1), b[0][0], ops,
                                                      " ..." if len(b) > 30 else ""))


if __name__ == "__main__":
    main()
