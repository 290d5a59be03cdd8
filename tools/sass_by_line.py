#!/usr/bin/env python
"""Attribute the executed instructions / stall samples of an ncu SASS source page to CUDA source lines.

    ncu -i rep.ncu-rep --page source --csv --print-source sass --kernel-name regex:trace_kernel > sass.csv
    cuobjdump -xelf all libart_b200.so; nvdisasm -gi -c art_b200.sm_100a.cubin > dis.txt
    python tools/sass_by_line.py sass.csv dis.txt <mangled kernel name> [rays_per_launch]

The two listings are joined by instruction index (both list the kernel's SASS in address order).
An instruction is charged to the innermost frame of its inline chain that is not a lane-pack operator or
an FP64 primitive (art_optics.cuh < line 130, art_device.cuh, CUDA headers)."""
import collections
import csv
import re
import sys

sass_csv, dis, kern = sys.argv[1:4]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ie, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
inst = [(r[isrc].strip(), int(r[ie] or 0), int(r[ismp] or 0)) for r in rows[2:] if len(r) > ie and r[0].startswith("0x")]
# a page that matched several launches repeats the listing: keep the first copy
first = inst[0][0]
reps = [i for i, x in enumerate(inst) if x[0] == first and i and inst[i + 1][0] == inst[1][0]]
if reps:
    inst = inst[:reps[0]]

lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kern + ":"))
def helper(f, ln):
    return (f == "art_optics.cuh" and ln < 130) or f == "art_device.cuh" or not f.startswith("art_")


chain = []
fresh = True
locs = []
for l in lines[start + 1:]:
    if l.startswith("//--------------------- "):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            chain = []
            fresh = False
        fr = (m.group(1).split("/")[-1], int(m.group(2)))
        if not chain or chain[-1] != fr:
            chain.append(fr)
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l):
        fresh = True
        pick = next((fr for fr in chain if not helper(*fr)), chain[-1] if chain else ("?", 0))
        locs.append(pick)
assert abs(len(locs) - len(inst)) <= 2, (len(locs), len(inst))
by = collections.Counter()
smp = collections.Counter()
fp = collections.Counter()
for (s, n, k), loc in zip(inst, locs):
    key = loc[:2]
    by[key] += n
    smp[key] += k
    op = re.sub(r"^@!?U?P\d+\s+", "", s).split(".")[0].split()[0]
    if op in ("DFMA", "DMUL", "DADD", "DSETP"):
        fp[key] += n
show = [a.split("=")[1] for a in sys.argv if a.startswith("--show=")]
if show:
    f, ln = show[0].split(":")
    for (s_, n, k), loc in zip(inst, locs):
        if loc == (f, int(ln)):
            print("%8.2f  smp %5d  %s" % (n * 32 / units, k, s_))
    sys.exit(0)
tot, tots = sum(by.values()), sum(smp.values())
print("total warp instructions %d (%.1f per unit), samples %d" % (tot, tot * 32 / units, tots))
src = {}
for (f, ln), n in by.most_common(60):
    if f not in src:
        try:
            src[f] = open("attosecondraytracing_b200/csrc/" + f).read().split("\n")
        except OSError:
            src[f] = []
    text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
    print("%-18s %4d  %6.1f (%5.1f%%)  fp64 %5.1f  smp %4.1f%%  %s" % (f, ln, n * 32 / units, 100.0 * n / tot, fp[(f, ln)] * 32 / units,
                                                              100.0 * smp[(f, ln)] / max(tots, 1), text))
