"""DRAM traffic of ONE BUILD from an ncu launch list (csv of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --clock-control none --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline`): the sum of
dram__bytes_read + dram__bytes_write over the build_level_kernel launches of the LAST proposal BUILD in the capture (phase 0 / 1
launches: one per level and width bucket).  Writes profiles/<tag>_build_dram.json, stamped with the SHA-1 of the kernel sources;
bench.py prints `roofline.traffic` only while that stamp matches the sources it runs.
Usage: ncu_build_dram.py launches.csv [workload] [tag]"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL_SOURCES = ["spamtree_b200/csrc/st_build.cu", "spamtree_b200/csrc/st_device.cuh", "spamtree_b200/csrc/st_build_plan.cuh"]


def sources_sha1():
    h = hashlib.sha1()
    for f in KERNEL_SOURCES:
        h.update(open(os.path.join(ROOT, f), "rb").read())
    return h.hexdigest()


def main():
    path = sys.argv[1]
    workload = sys.argv[2] if len(sys.argv) > 2 else "C4"
    tag = sys.argv[3] if len(sys.argv) > 3 else "r2"
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[h], rows[h + 1:]
    ki, vi, gi, mi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size"), H.index("Metric Name")
    per = {}
    for r in data:
        if len(r) <= vi:
            continue
        d = per.setdefault(int(r[0]), {"name": r[ki], "grid": r[gi]})
        d[r[mi]] = float(r[vi].replace(",", ""))
    ids = [i for i in sorted(per) if "build_level_kernel" in per[i]["name"]]
    # a BUILD starts at the root level: one work group (grid of 1); its launches follow in id order (kernels of the other
    # stream may sit in between), the no-op launches of the deferred half (rejected proposal) add nothing
    builds = []
    for i in ids:
        if per[i]["grid"].replace(" ", "") == "(1,1,1)" or not builds:
            builds.append([])
        builds[-1].append(i)
    nmax = max(len(b) for b in builds)
    full = [b for b in builds if len(b) >= nmax - 2]
    # the BUILD of a REJECTED proposal (its deferred-half launches are no-ops): the complete group with the least traffic
    last = min(full, key=lambda b: sum(per[i].get("dram__bytes_read.sum", 0.0) + per[i].get("dram__bytes_write.sum", 0.0) for i in b))
    rd = sum(per[i].get("dram__bytes_read.sum", 0.0) for i in last)
    wr = sum(per[i].get("dram__bytes_write.sum", 0.0) for i in last)
    us = sum(per[i].get("gpu__time_duration.sum", 0.0) for i in last) / 1e3
    out = {"workload": workload, "build_launches": len(last), "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
           "kernel_time_us_under_ncu": us, "sources_sha1": sources_sha1(), "sources": KERNEL_SOURCES, "launch_list": os.path.basename(path),
           "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv "
                  "python bench.py --steps 2 --warmup 3 --no-cpu-baseline; sum over the build_level_kernel launches of one proposal BUILD"}
    dst = os.path.join(ROOT, "profiles", f"{tag}_build_dram.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
