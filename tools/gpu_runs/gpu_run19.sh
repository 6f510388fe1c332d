#!/bin/bash
# A/B on one box: new = loop head without the rebuilt slot pointer / spilled flag (kernel r2.8), alt = r2.7
cd /root/repo; mkdir -p gpurun_out; T=r2_v16
cp divans_b200/lib/libdivans_b200.so /tmp/new.so; cp divans_b200/lib/alt/libdivans_b200.so /tmp/alt.so
for round in 1 2; do for V in new alt; do
  cp /tmp/$V.so divans_b200/lib/libdivans_b200.so
  echo "== $V (round $round)" | tee -a gpurun_out/${T}_ab.txt
  timeout 300 python tools/zprobe.py 4096 2>&1 | tail -1 | tee -a gpurun_out/${T}_ab.txt
  timeout 300 python tools/perf_probe.py --l-only --decode-once 4096 2>&1 | cut -c1-150 | tee -a gpurun_out/${T}_ab.txt
  DIVANS_B200_LPS=0 timeout 300 python tools/perf_probe.py --l-only --decode-once 8192 2>&1 | cut -c1-150 | tee -a gpurun_out/${T}_ab.txt
done; done
cp /tmp/new.so divans_b200/lib/libdivans_b200.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
