#!/usr/bin/env python
"""Source lines of a profiled kernel ranked by one stall reason (default stall_long_sb), with their global-load sector counts.
usage: tools/ncu_stall_lines.py rep.ncu-rep [column] [top_n]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    col = sys.argv[2] if len(sys.argv) > 2 else "stall_long_sb"
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    path, hdr, cur = None, None, None
    acc = collections.defaultdict(lambda: [0, 0, 0, ""])
    for r in rows:
        if len(r) == 2 and r[0] in ("File Path", "File Name"):
            path = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
            ic, isamp, il2 = hdr.index(col), hdr.index("# Samples"), hdr.index("L2 Theoretical Sectors Global")
        elif hdr and len(r) == len(hdr):
            if r[0] != "":
                cur = (path, r[0])
                acc[cur][3] = r[1].strip()[:100]

            def num(x):
                try:
                    return int(x)
                except ValueError:
                    return 0
            if cur is not None and r[0] == "":      # SASS rows carry the per-instruction numbers
                acc[cur][0] += num(r[ic]); acc[cur][1] += num(r[isamp]); acc[cur][2] += num(r[il2])
    tot = sum(v[0] for v in acc.values()) or 1
    tots = sum(v[1] for v in acc.values()) or 1
    print("%s: %d samples of %d" % (col, tot, tots))
    for (p, ln), v in sorted(acc.items(), key=lambda t: -t[1][0])[:top]:
        print("%6.2f%%  L2 sectors %9d  %s:%s  %s" % (100.0 * v[0] / tot, v[2], p, ln, v[3]))


if __name__ == "__main__":
    main()
