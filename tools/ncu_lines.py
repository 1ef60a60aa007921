"""Attributes the warp-stall samples of an ncu report (--set full --import-source on) to SOURCE LINES of a kernel.
ncu's CSV source page is SASS-only, so the instruction offsets are joined with `nvdisasm -g` line info of the cubin.
Usage: ncu_lines.py <report.ncu-rep> <object-or-.so with the kernel> <kernel substring, e.g. ILi1E> [section index] [top N]"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def line_table(binary, ksub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(binary)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        infn, cur = False, None
        for ln in txt.splitlines():
            if ln.startswith("//-----") and ".text." in ln:
                infn = ksub in ln
                continue
            if not infn:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return out


def main():
    rep, binary, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
    sec = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    lt = line_table(binary, ksub)
    csv.field_size_limit(10 ** 9)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            secs.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif r and cur is not None:
            cur["rows"].append(r)
    m = re.search(r"ILi(\d+)E", ksub)
    disp = f"<(int){m.group(1)}>" if m else ksub
    secs = [s for s in secs if disp in s["name"]] or secs
    s = secs[sec]
    h = s["hdr"]
    ia, isamp, iex = h.index("Address"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_")]
    base = min(int(r[ia], 16) for r in s["rows"])
    by_line, tot = {}, 0
    for r in s["rows"]:
        off = int(r[ia], 16) - base
        key = lt.get(off, (None, ""))[0]
        d = by_line.setdefault(key, {"samples": 0, "inst": 0, "stalls": {}})
        n = int(r[isamp] or 0)
        d["samples"] += n
        d["inst"] += int(r[iex] or 0)
        tot += n
        for i in stall_cols:
            v = int(r[i] or 0)
            if v:
                d["stalls"][h[i]] = d["stalls"].get(h[i], 0) + v
    print(f"kernel: {s['name'][:80]}  total samples {tot}")
    for key, d in sorted(by_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(d["stalls"].items(), key=lambda kv: -kv[1])[:4]
        print(f"{str(key):28s} {100 * d['samples'] / max(tot, 1):6.2f}%  inst {d['inst']:>11d}  " + " ".join(f"{k[6:]}={v}" for k, v in st))


if __name__ == "__main__":
    main()
