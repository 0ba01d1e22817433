#!/usr/bin/env python
"""Record the DRAM traffic of one profiled launch (an `ncu --set full` .ncu-rep) in profiles/r1_traffic.json, the file
bench.py reads `roofline.traffic` from.  usage: tools/ncu_traffic.py rep.ncu-rep game/obs/n_envs [out.json]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, key = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(out.splitlines()) if r]
    hdr, units, row = rows[0], rows[1], rows[2]

    def val(name):
        i = hdr.index(name)
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[i]]
        return float(row[i]) * scale

    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "r1_traffic.json")
    tr = json.load(open(path)) if os.path.exists(path) else {}
    tr[key] = {"kernel": row[hdr.index("Kernel Name")], "dram_bytes": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
               "source": os.path.basename(rep)}
    json.dump(tr, open(path, "w"), indent=1, sort_keys=True)
    print(key, tr[key])


if __name__ == "__main__":
    main()
