#!/usr/bin/env python
"""Hot-spot digest of an .ncu-rep's SASS page (run here, no GPU needed):
     python tools/ncu_hot.py rep.ncu-rep [--top 25] [--kernel regex]
   prints executed warp-instructions by opcode, and the instructions with the most stall samples."""
import argparse
import collections
import csv
import re
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--ctx", type=int, default=0)
    args = ap.parse_args()
    out = subprocess.run(["ncu", "-i", args.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    insts = []
    for r in rows:
        if r and r[0] == "Kernel Name":
            print("kernel:", r[1][:160])
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) >= len(hdr) - 2:
            d = dict(zip(hdr, r))
            try:
                insts.append((d["Source"].strip(), int(d["Instructions Executed"]), int(d["Warp Stall Sampling (All Samples)"]),
                              int(d.get("L1 Wavefronts Shared", 0) or 0), int(d.get("L1 Wavefronts Shared Ideal", 0) or 0)))
            except (ValueError, KeyError):
                pass
    tot = sum(i[1] for i in insts)
    samp = sum(i[2] for i in insts)
    print(f"SASS instructions: {len(insts)}   executed warp-instructions: {tot}   stall samples: {samp}")
    by_op = collections.Counter()
    by_op_s = collections.Counter()
    for s, n, st, _, _ in insts:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", s)
        op = m.group(2) if m else s.split()[0]
        by_op[op] += n
        by_op_s[op] += st
    print("executed by opcode (share of executed / share of stall samples):")
    for op, n in by_op.most_common(22):
        print(f"  {op:12s} {100.0 * n / max(tot, 1):5.1f} %   {100.0 * by_op_s[op] / max(samp, 1):5.1f} %")
    wf = sum(i[3] for i in insts)
    wfi = sum(i[4] for i in insts)
    print(f"shared-memory wavefronts: {wf} (ideal {wfi})")
    print(f"top {args.top} instructions by stall samples:")
    order = sorted(range(len(insts)), key=lambda i: -insts[i][2])[: args.top]
    for i in sorted(order):
        for j in range(max(0, i - args.ctx), min(len(insts), i + args.ctx + 1)):
            s, n, st, w, wi = insts[j]
            mark = "*" if j == i else " "
            print(f" {mark} #{j:5d} samples {st:6d} ({100.0 * st / max(samp, 1):4.1f} %)  exec {n:9d}  {s[:110]}")


if __name__ == "__main__":
    main()
