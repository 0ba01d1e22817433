#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples from an .ncu-rep captured with --import-source on.
usage: tools/ncu_lines.py rep.ncu-rep [top_n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    path, hdr, lines = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] in ("File Path", "File Name"):
            path = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0] != "":
            try:
                inst = int(r[hdr.index("Instructions Executed")])
                samp = int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            lines.append((inst, samp, path, r[0], r[1].strip()[:110]))
    tot_i = sum(l[0] for l in lines) or 1
    tot_s = sum(l[1] for l in lines) or 1
    print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
    print("%7s %7s  %s" % ("inst%", "samp%", "line"))
    for inst, samp, path, ln, src in sorted(lines, reverse=True)[:top]:
        print("%6.2f%% %6.2f%%  %s:%s  %s" % (100.0 * inst / tot_i, 100.0 * samp / tot_s, path, ln, src))


if __name__ == "__main__":
    main()
