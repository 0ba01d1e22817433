#!/usr/bin/env python
"""Warp instructions per env of a profiled kernel, bucketed by source file and line range.
usage: tools/ncu_phases.py rep.ncu-rep n_envs file:lo-hi=name [...]   (unmatched lines are listed per file)"""
import collections
import csv
import subprocess
import sys


def main():
    rep, n_envs = sys.argv[1], float(sys.argv[2])
    ranges = []
    for spec in sys.argv[3:]:
        loc, name = spec.split("=")
        f, r = loc.split(":")
        lo, hi = r.split("-")
        ranges.append((f, int(lo), int(hi), name))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    path, hdr, cur, seen = None, None, None, set()
    per = collections.Counter()
    for r in rows:
        if len(r) == 2 and r[0] in ("File Path", "File Name"):
            path = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
            ie = hdr.index("Instructions Executed")
        elif hdr and len(r) == len(hdr):
            if r[0] != "":
                cur = (path, int(r[0]))
            else:
                if r[2] in seen:
                    continue
                seen.add(r[2])
                try:
                    v = int(r[ie])
                except ValueError:
                    continue
                name = None
                for f, lo, hi, nm in ranges:
                    if cur[0] == f and lo <= cur[1] <= hi:
                        name = nm
                        break
                per[name or cur[0]] += v
    tot = sum(per.values())
    print("total %.0f warp instructions per env" % (tot / n_envs))
    for k, v in per.most_common():
        print("  %-34s %6.2f%% %8.0f /env" % (k, 100.0 * v / tot, v / n_envs))


if __name__ == "__main__":
    main()
