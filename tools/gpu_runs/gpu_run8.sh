#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; T=r2_v7
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python tools/perf_probe.py --l-only --decode-once 4096 > gpurun_out/${T}_probe_4096.txt 2>&1; cat gpurun_out/${T}_probe_4096.txt | cut -c1-260
timeout 300 python tools/perf_probe.py --l-only --decode-once 8192 > gpurun_out/${T}_probe_8192.txt 2>&1; cat gpurun_out/${T}_probe_8192.txt | cut -c1-260
timeout 300 python tools/zprobe.py 4096 > gpurun_out/${T}_z4096.txt 2>&1; cat gpurun_out/${T}_z4096.txt
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/${T}_bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel_v2 -c 1 -o gpurun_out/${T}_dec16_4096 python tools/perf_probe.py --l-only --decode-once 4096 > gpurun_out/${T}_ncu16.log 2>&1; echo "ncu rc=$?"
