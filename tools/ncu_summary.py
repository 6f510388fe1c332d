#!/usr/bin/env python
"""Summarise one `ncu --set full --import-source on` capture of a decode kernel into the committed files under profiles/:
  <prefix>_metrics.csv   key counters (time, instructions, issue activity, hit rates, DRAM bytes, occupancy)
  <prefix>_stalls.csv    warp stall reasons, cycles per issued instruction and share
  <prefix>_hot_sass.csv  the SASS lines with the most stall samples (share, executions)
and, with --traffic, profiles/traffic.json (DRAM bytes per launch, stamped with the kernel version, lanes and stream count).
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r2_xyz [--traffic --lanes 16 --streams 4096]"""
import csv, json, os, subprocess, sys

rep, prefix = sys.argv[1], sys.argv[2]
def page(name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows
rows = page("raw")
H, U, V = rows[0], rows[1], rows[2]
val = {h: (V[i], U[i]) for i, h in enumerate(H)}
def num(h):
    try:
        return float(val[h][0].replace(",", ""))
    except Exception:
        return None
KEYS = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
with open(prefix + "_metrics.csv", "w") as f:
    w = csv.writer(f); w.writerow(["metric", "value", "unit"])
    for k in KEYS:
        if k in val:
            w.writerow([k, val[k][0], val[k][1]])
st = []
for h in H:
    if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
        v = num(h)
        if v is not None:
            st.append((v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
tot = sum(v for v, _ in st) or 1.0
with open(prefix + "_stalls.csv", "w") as f:
    w = csv.writer(f); w.writerow(["stall_reason", "warp_cycles_per_issued_instruction", "share_pct"])
    for v, h in sorted(st, reverse=True):
        w.writerow([h, "%.3f" % v, "%.1f" % (100 * v / tot)])
src = page("source")
while src and "Source" not in src[0]:
    src.pop(0)
if src:
    Hs = src[0]
    si, ci, ii = Hs.index("Source"), Hs.index("Warp Stall Sampling (All Samples)"), Hs.index("Instructions Executed")
    rr = []
    for r in src[1:]:
        try:
            rr.append((float(r[ci] or 0), r[si].strip(), r[ii]))
        except Exception:
            pass
    tots = sum(x[0] for x in rr) or 1.0
    with open(prefix + "_hot_sass.csv", "w") as f:
        w = csv.writer(f); w.writerow(["stall_samples", "share_pct", "instructions_executed", "sass"])
        for s_, t_, n_ in sorted(rr, reverse=True)[:60]:
            w.writerow([int(s_), "%.2f" % (100 * s_ / tots), n_, t_])
def to_bytes(h):
    v, u = num(h), val[h][1].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
print("%s: %.2f ms, dram read %.3f GB write %.3f GB, L1 hit %.1f%%, L2 hit %.1f%%, issue active %.1f%%" % (
    val.get("Kernel Name", ("?",))[0][:60], num("gpu__time_duration.sum") or 0, rd / 1e9, wr / 1e9, num("l1tex__t_sector_hit_rate.pct") or 0,
    num("lts__t_sector_hit_rate.pct") or 0, num("smsp__issue_active.avg.pct_of_peak_sustained_active") or 0))
if "--traffic" in sys.argv:
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    import divans_b200
    lanes = int(sys.argv[sys.argv.index("--lanes") + 1]); streams = int(sys.argv[sys.argv.index("--streams") + 1])
    entry = {"decode_kernel_dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
             "kernel_version": divans_b200.kernel_version(), "lanes_per_stream": lanes, "streams": streams,
             "source": os.path.basename(prefix) + "_metrics.csv (ncu --set full, %d x 64 KiB streams, one launch of the decode kernel)" % streams}
    # one entry per (lane layout, stream count): the N = 1 and the N > 1 workloads of bench.py run different instantiations
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    entries = []
    if os.path.exists(tp) and "--replace" not in sys.argv:
        old = json.load(open(tp))
        entries = old.get("entries", [old] if "kernel_version" in old else [])
    entries = [e for e in entries if (e.get("lanes_per_stream"), e.get("streams")) != (lanes, streams)] + [entry]
    json.dump({"entries": entries}, open(tp, "w"), indent=1)
