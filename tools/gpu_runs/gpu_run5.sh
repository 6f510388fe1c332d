#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; T=r2_v4
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${T}_pytest.log
DIVANS_B200_LPS=0 timeout 300 python tools/perf_probe.py 4096 --l-only > gpurun_out/${T}_probe_auto_4096.txt 2>&1; echo "== auto n=4096"; head -2 gpurun_out/${T}_probe_auto_4096.txt
DIVANS_B200_LPS=0 timeout 300 python tools/perf_probe.py 8192 --l-only > gpurun_out/${T}_probe_auto_8192.txt 2>&1; echo "== auto n=8192"; head -2 gpurun_out/${T}_probe_auto_8192.txt
DIVANS_B200_LPS=16 timeout 300 python tools/perf_probe.py 8192 --l-only > gpurun_out/${T}_probe_l16_8192.txt 2>&1; echo "== lanes 16 n=8192"; head -2 gpurun_out/${T}_probe_l16_8192.txt
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
DIVANS_B200_LPS=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel_v2 -c 1 -o gpurun_out/${T}_dec16_4096 python tools/perf_probe.py 4096 --l-only --decode-once > gpurun_out/${T}_ncu16.log 2>&1; echo "ncu16 rc=$?"
