#!/usr/bin/env python
"""Turn the ncu artefacts of one tools/gpu_runs/gpu_round_r1.sh visit (gpurun_out/<tag>_*) into the committed summaries under profiles/.
Usage: python tools/summarize_profiles.py <tag>"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
KEEP = ["gpu__time", "inst_executed", "issue_active", "warps_active", "hit_rate", "dram__bytes", "dram__throughput", "launch__",
        "stalled", "sm__throughput", "lts__throughput", "l1tex__throughput", "lts__t_sectors_srcunit_tex_op"]


def launches():
    src = os.path.join(G, tag + "_launches.csv")
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    kn, mv, mu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = {"ns": v / 1e6, "us": v / 1e3, "ms": v, "s": v * 1e3, "second": v * 1e3}.get(r[mu], v)
        a = agg.setdefault(r[kn].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, tag + "_launch_shares.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "avg_ms", "share_pct"])
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            w.writerow([k, c, "%.3f" % t, "%.4f" % (t / c), "%.2f" % (100 * t / tot)])
    with open(os.path.join(P, tag + "_launches.csv"), "w") as f:
        f.write(open(src).read())


def full(name, out):
    rep = os.path.join(G, "%s_%s.ncu-rep" % (tag, name))
    if not os.path.exists(rep):
        return None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U, V = rows[0], rows[1], rows[2]
    d = {}
    with open(os.path.join(P, "%s_%s_ncu_full.csv" % (tag, out)), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "value", "unit"])
        for i, h in enumerate(H):
            d[h] = (V[i], U[i])
            if any(x in h for x in KEEP):
                w.writerow([h, V[i], U[i]])
    # hottest source lines by stall samples
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    try:
        rows = list(csv.reader(src.splitlines()))
        while rows and "Source" not in rows[0]:
            rows.pop(0)
        H = rows[0]
        si = H.index("Source")
        ci = H.index("Warp Stall Sampling (All Samples)") if "Warp Stall Sampling (All Samples)" in H else None
        ii = H.index("Instructions Executed") if "Instructions Executed" in H else None
        if ci is not None:
            rr = []
            for r in rows[1:]:
                try:
                    rr.append((float(r[ci] or 0), r[si].strip(), r[ii] if ii is not None else ""))
                except Exception:
                    pass
            rr.sort(reverse=True)
            tot = sum(x[0] for x in rr) or 1
            with open(os.path.join(P, "%s_%s_hot_sass.csv" % (tag, out)), "w") as f:
                w = csv.writer(f)
                w.writerow(["stall_samples", "share_pct", "instructions_executed", "sass"])
                for x in rr[:60]:
                    w.writerow([int(x[0]), "%.2f" % (100 * x[0] / tot), x[2], x[1]])
            # stall reasons summed over the kernel
            reasons = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
            sums = {h: 0.0 for h in reasons}
            for r in rows[1:]:
                for h in reasons:
                    try:
                        sums[h] += float(r[H.index(h)] or 0)
                    except Exception:
                        pass
            with open(os.path.join(P, "%s_%s_stall_reasons.csv" % (tag, out)), "w") as f:
                w = csv.writer(f)
                w.writerow(["reason", "samples", "share_pct"])
                t2 = sum(sums.values()) or 1
                for h, v in sorted(sums.items(), key=lambda x: -x[1]):
                    w.writerow([h, int(v), "%.2f" % (100 * v / t2)])
    except Exception as e:
        print("source page:", e)
    return d


launches()
d = full("decode", "decode_kernel")
full("encode", "encode_model_kernel")
d4 = full("decode4096", "decode_kernel_4096streams")
if d4:
    def b(x):
        v, u = x
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    tr = b(d4["dram__bytes_read.sum"]) + b(d4["dram__bytes_write.sum"])
    json.dump({"decode_kernel_dram_bytes_per_launch": tr, "source": "profiles/%s_decode_kernel_4096streams_ncu_full.csv (ncu --set full, 4096 x 64 KiB streams, one launch)" % tag,
               "dram_read_bytes": b(d4["dram__bytes_read.sum"]), "dram_write_bytes": b(d4["dram__bytes_write.sum"])},
              open(os.path.join(P, "traffic.json"), "w"), indent=1)
b = os.path.join(G, tag + "_bench.json")
if os.path.exists(b) and os.path.getsize(b):
    open(os.path.join(P, tag + "_bench.json"), "w").write(open(b).read())
print("ok")
