#!/usr/bin/env python
"""Instruction counts of the profiled kernel per source file / per line range (SASS addresses de-duplicated).
usage: tools/ncu_buckets.py rep.ncu-rep n_envs [top_lines]"""
import collections
import csv
import subprocess
import sys


def main():
    rep, n_envs = sys.argv[1], float(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    path, hdr, cur, seen = None, None, None, set()
    per, samp, text = collections.Counter(), collections.Counter(), {}
    for r in rows:
        if len(r) == 2 and r[0] in ("File Path", "File Name"):
            path = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
            ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        elif hdr and len(r) == len(hdr):
            if r[0] != "":
                cur = (path, int(r[0]))
                text[cur] = r[1].strip()[:100]
            else:
                if r[2] in seen:
                    continue
                seen.add(r[2])
                try:
                    per[cur] += int(r[ie])
                    samp[cur] += int(r[isamp])
                except ValueError:
                    pass
    tot, tots = sum(per.values()), sum(samp.values()) or 1
    print("total warp instructions %d = %.0f per env; stall samples %d" % (tot, tot / n_envs, tots))
    byfile = collections.Counter()
    for (f, l), v in per.items():
        byfile[f] += v
    for f, v in byfile.most_common():
        print("  %-26s %6.2f%% %8.0f /env" % (f, 100.0 * v / tot, v / n_envs))
    print("top lines (inst%%, samples%%, inst/env):")
    for k, v in per.most_common(top):
        print("  %5.2f%% %5.2f%% %7.0f  %s:%d  %s" % (100.0 * v / tot, 100.0 * samp[k] / tots, v / n_envs, k[0], k[1], text.get(k, "")))


if __name__ == "__main__":
    main()
