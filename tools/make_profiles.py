#!/usr/bin/env python
"""Turns the ncu captures of a round into the tracked artefacts under profiles/:

    python tools/make_profiles.py <dir with rNN_<workload>.ncu-rep> <round tag, e.g. r02> [launches.csv]

  profiles/<tag>_ncu_full_<workload>_metrics.csv   selected raw metrics per captured kernel (tools/ncu_summary.py)
  profiles/<tag>_traffic.json                      dram__bytes_read.sum + dram__bytes_write.sum of the trace kernel per
                                                   workload, stamped with the commit the library was built from
                                                   (bench.py quotes it as roofline.traffic)
  profiles/<tag>_by_line_<workload>.txt            executed instructions / stall samples per source line (sass_by_line.py)
  profiles/<tag>_sass_bulk_copy.txt                the UBLKCP / SYNCS (cp.async.bulk + mbarrier) instructions of the library
  profiles/<tag>_launches_default.csv              copy of the launch list
"""
import csv
import glob
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, tag = sys.argv[1], sys.argv[2]
launches = sys.argv[3] if len(sys.argv) > 3 else None
prof = os.path.join(ROOT, "profiles")
head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True).stdout.strip()
lib = os.path.join(ROOT, "attosecondraytracing_b200", "libart_b200.so")

traffic = {"head": head, "how": "ncu --set full --clock-control none, one launch of the trace kernel inside "
                               "`bench.py --workload <w> --sub none --steps 2 --warmup 1 --no-graph`"}
work = "/tmp/make_profiles"
shutil.rmtree(work, ignore_errors=True)
os.makedirs(work)
subprocess.run("cuobjdump -xelf all %s > /dev/null && nvdisasm -gi -c *.cubin > dis.txt" % lib, shell=True, cwd=work, check=True)
for rep in sorted(glob.glob(os.path.join(src, "*_cfg*.ncu-rep"))):
    w = re.search(r"_(cfg[0-9a-z]+)\.ncu-rep$", rep).group(1)
    out_csv = os.path.join(prof, f"{tag}_ncu_full_{w}_metrics.csv")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, out_csv], capture_output=True, check=True)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if "trace_kernel" in d["Kernel Name"] and w not in traffic:
            units_row = dict(zip(hdr, rows[1]))
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(d["dram__bytes_read.sum"]) * scale[units_row["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"]) * scale[units_row["dram__bytes_write.sum"]]
            traffic[w] = {"traffic": rd + wr, "dram_read": rd, "dram_write": wr,
                          "kernel": d["Kernel Name"], "ncu_duration_us": float(d["gpu__time_duration.sum"]) *
                          {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(dict(zip(hdr, rows[1]))["gpu__time_duration.sum"], 1.0)}
    # by-line attribution of every distinct kernel in the capture
    for pat, kname in (("trace_kernel", None), ("detector_bulk", None)):
        sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                               "regex:" + pat], capture_output=True, text=True, cwd=work).stdout
        if "Address" not in sass:
            continue
        open(os.path.join(work, "sass.csv"), "w").write(sass)
        first = sass.splitlines()[0]
        m = re.search(r'"Kernel Name","([^"]+)"', first)
        demangled = m.group(1) if m else pat
        # mangled name: look it up in the disassembly by the demangled template arguments
        cands = [l[len(".text."):].rstrip(":\n") for l in open(os.path.join(work, "dis.txt"))
                 if l.startswith(".text._ZN3art") and pat.split("_")[0] in l]
        args = re.findall(r"\((?:bool|int)\)(\d+)", demangled)
        mangled = None
        for c in cands:
            if pat == "detector_bulk" and "detector_bulk" in c:
                mangled = c
            if pat == "trace_kernel" and "trace_kernel" in c:
                got = re.findall(r"L[bi](\d+)E", c)
                if got == args:
                    mangled = c
        if not mangled:
            continue
        units = {"cfg3": 12500000, "cfg2": 10000000, "cfg4": 50000000, "cfg4def": 50000000, "cfg5": 32000000}.get(w, 1)
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_by_line.py"), os.path.join(work, "sass.csv"),
                              os.path.join(work, "dis.txt"), mangled, str(units)], capture_output=True, text=True, cwd=ROOT)
        with open(os.path.join(prof, f"{tag}_by_line_{w}_{pat}.txt"), "w") as f:
            f.write(f"# {demangled}\n# per unit = per source ray of the launch ({units} rays); capture {os.path.basename(rep)}, "
                    f"library at {head}\n" + res.stdout[:12000])
json.dump(traffic, open(os.path.join(prof, f"{tag}_traffic.json"), "w"), indent=1)
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keep, fn = [], ""
for line in sass.splitlines():
    if "Function :" in line:
        fn = line.strip()
        if "detector_bulk" in fn:
            keep.append(fn)
    if "detector_bulk" in fn and re.search(r"UBLKCP|SYNCS", line):
        keep.append("    " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line.strip()))
# the bounded try_wait spin is unrolled by the compiler: collapse runs of the same mnemonic
out, last, run = [], None, 0
for line in keep:
    m = re.search(r"\*/\s+(\S+)", line)
    op = m.group(1) if m else line
    if op == last:
        run += 1
        continue
    if run:
        out.append(f"        ... {run} more {last}")
    out.append(line)
    last, run = op, 0
if run:
    out.append(f"        ... {run} more {last}")
keep = out
open(os.path.join(prof, f"{tag}_sass_bulk_copy.txt"), "w").write(
    f"# cuobjdump -sass libart_b200.so (built from {head}): the bulk asynchronous copies (cp.async.bulk -> UBLKCP) and\n"
    "# mbarrier operations (SYNCS) of detector_bulk_kernel\n" + "\n".join(keep) + "\n")
if launches:
    shutil.copyfile(launches, os.path.join(prof, f"{tag}_launches_default.csv"))
print(json.dumps(traffic, indent=1))
