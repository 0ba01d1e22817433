#!/usr/bin/env python
"""Copy the artefacts of a GPU round (gpurun_out/<tag>_*) into profiles/ (tracked): bench logs become one-line .json files
(the JSON line of the run), ncu summaries / per-line tables / launch lists / test logs are copied as they are.
usage: tools/collect_profiles.py [tag]        (default tag: r2)"""
import glob
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    n = 0
    for path in sorted(glob.glob(os.path.join(src, tag + "_*"))):
        name = os.path.basename(path)
        if name.endswith(".ncu-rep") or "_ncu_" in name and name.endswith(".log"):
            continue
        if name.startswith(tag + "_bench_") and name.endswith(".log"):
            lines = [l for l in open(path) if l.startswith("{")]
            if not lines:
                print("no JSON line in", name)
                continue
            with open(os.path.join(dst, name[:-4] + ".json"), "w") as fh:
                fh.write(lines[-1])
        else:
            shutil.copyfile(path, os.path.join(dst, name))
        n += 1
    print("copied", n, "files into profiles/")


if __name__ == "__main__":
    main()
