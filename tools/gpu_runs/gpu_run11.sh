#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; T=r2_v10
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python tools/perf_probe.py --l-only --decode-once 4096 2>&1 | cut -c1-260 | tee gpurun_out/${T}_probe_4096.txt
echo "== Z lanes 16"; timeout 300 python tools/zprobe.py 4096 2>&1 | tee gpurun_out/${T}_z16.txt
echo "== Z lanes 8"; DIVANS_B200_LPS=8 timeout 300 python tools/zprobe.py 4096 2>&1 | tee gpurun_out/${T}_z8.txt
echo "== Z lanes 8 n=8192"; DIVANS_B200_LPS=8 timeout 300 python tools/zprobe.py 8192 2>&1 | tee gpurun_out/${T}_z8_8192.txt
