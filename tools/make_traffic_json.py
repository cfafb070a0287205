"""Regenerates profiles/traffic.json -- DRAM bytes per launch of the solve kernels -- from an ncu report.

    python tools/make_traffic_json.py gpurun_out/r02_head_mix.ncu-rep --batches 1024,1024,32768,32768 --source "<what was captured>"

The report must hold the `solve_kernel` launches of tools/prof_target2.py in order (`ncu --set full -k regex:solve_kernel`): for
every batch size in --batches one group of class kernels (as many as the batch has stance classes; groups are delimited by the
class sequence repeating).  The LAST group of each batch size is kept (warm).  bench.py reads the file for `roofline.traffic`.
"""
import argparse
import csv
import json
import os
import re
import subprocess
import sys

ap = argparse.ArgumentParser()
ap.add_argument("rep", nargs="+", help="one or more .ncu-rep files, in launch order")
ap.add_argument("--batches", required=True, help="batch size of every successive group of class launches, e.g. 1024,1024,32768,32768")
ap.add_argument("--source", default="")
ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"))
args = ap.parse_args()
reports = []
for rep in args.rep:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rws = list(csv.reader(raw.splitlines()))
    reports.append((rws[0], rws[1], rws[2:]))


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


batches = [int(b) for b in args.batches.split(",")]
groups = []
for hdr, units, rows in reports:
    ix = {h: i for i, h in enumerate(hdr)}
    seen = None
    for r in rows:
        name = r[ix["Kernel Name"]]
        m = re.search(r"solve_kernel(?:_warm)?<(\d+), (\d+), (\d+), (\d+)(?:, (\d+))?>", name)
        if not m:
            continue
        key = (int(m.group(1)), int(m.group(2)))
        if seen is None or key in seen:
            groups.append([])
            seen = set()
        seen.add(key)
        groups[-1].append((key, name, r, ix, units))
if len(groups) != len(batches):
    sys.exit("found %d groups of class launches, --batches names %d" % (len(groups), len(batches)))
out = {"source": args.source or ", ".join(os.path.basename(r) for r in args.rep), "tool": "tools/make_traffic_json.py", "kernels": {}}
for B, grp in zip(batches, groups):
    for (ns, n), name, r, ix, units in grp:
        rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
        wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        dur = float(r[ix["gpu__time_duration.sum"]].replace(",", ""))
        out["kernels"]["solve_kernel<NS=%d,N=%d>@%d" % (ns, n, B)] = {
            "ncu_kernel_name": re.sub(r"^void |\(.*$", "", name), "dram_bytes_read": rd, "dram_bytes_write": wr,
            "gpu_time_duration": dur, "gpu_time_unit": units[ix["gpu__time_duration.sum"]], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]}
json.dump(out, open(args.out, "w"), indent=1)
print("wrote", args.out, "with", len(out["kernels"]), "entries")
