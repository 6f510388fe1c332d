#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; T=r2_v11
timeout 600 python -m pytest tests/test_gpu_blend.py -m gpu -x -q > gpurun_out/${T}_pytest_blend.log 2>&1; echo "pytest blend rc=$?"; tail -15 gpurun_out/${T}_pytest_blend.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${T}_pytest.log
timeout 300 python tools/perf_probe.py --l-only --decode-once 4096 2>&1 | cut -c1-260 | tee gpurun_out/${T}_probe_4096.txt
echo "== Z lanes 16"; timeout 300 python tools/zprobe.py 4096 2>&1 | tee gpurun_out/${T}_z16.txt
