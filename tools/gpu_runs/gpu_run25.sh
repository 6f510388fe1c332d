#!/bin/bash
# final library r2.11 (= r2.10 with the 8-lane literal loop unrolled by 2 again): tests, capture of the 8-lane kernel at 8192 streams
cd /root/repo; mkdir -p gpurun_out; T=r2_fin3
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest.log
DIVANS_B200_LPS=0 timeout 300 python tools/perf_probe.py --l-only --decode-once 8192 2>&1 | cut -c1-120
DIVANS_B200_LPS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel_v2 -c 1 -o gpurun_out/${T}_dec8_8192 python tools/perf_probe.py --l-only --decode-once 8192 > gpurun_out/${T}_ncu8.log 2>&1; echo "ncu8 rc=$?"
