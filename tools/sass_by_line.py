#!/usr/bin/env python
"""Join an ncu capture (per-SASS-instruction executions and stall samples) with the line table of the library that produced
it (nvdisasm -g on the cubin inside libdivans_b200.so): instructions and stall samples per CUDA source line / function.
Usage: python tools/sass_by_line.py gpurun_out/x.ncu-rep <kernel name substring> [top N]"""
import collections, csv, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
so = os.path.join(ROOT, "divans_b200", "lib", "libdivans_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
line_of = {}
for cub in os.listdir(tmp):
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    infn, cur = False, None
    for ln in dis.splitlines():
        if ".section" in ln and ".text." in ln:
            infn = kname in ln
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", ln)
        if infn and m:
            line_of[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
while rows and "Source" not in rows[0]:
    rows.pop(0)
H = rows[0]
ai, ci, ii = H.index("Address"), H.index("Warp Stall Sampling (All Samples)"), H.index("Instructions Executed")
base = None
inst, stall = collections.Counter(), collections.Counter()
for r in rows[1:]:
    try:
        a = int(r[ai], 16) if r[ai].startswith("0x") else int(r[ai])
    except Exception:
        continue
    if base is None:
        base = a
    key = line_of.get(a - base, ("?", 0))
    inst[key] += float(r[ii] or 0); stall[key] += float(r[ci] or 0)
ti, ts = sum(inst.values()) or 1, sum(stall.values()) or 1
print("%-34s %8s %8s" % ("source line", "instr %", "stall %"))
for k, v in sorted(inst.items(), key=lambda x: -x[1])[:top]:
    print("%-34s %8.2f %8.2f" % ("%s:%d" % k, 100 * v / ti, 100 * stall[k] / ts))
byf = collections.Counter()
for k, v in inst.items():
    byf[k[0]] += v
print({k: round(100 * v / ti, 1) for k, v in byf.most_common()})
