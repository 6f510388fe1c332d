#!/bin/bash
# final library of the round (kernel r2.10): tests, Z / L probes, traffic captures for both bench workloads, bench line, reference arm
cd /root/repo; mkdir -p gpurun_out; T=r2_fin2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python tools/zprobe.py 4096 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel_v2 -c 1 -o gpurun_out/${T}_dec16_4096 python tools/perf_probe.py --l-only --decode-once 4096 > gpurun_out/${T}_ncu16.log 2>&1; echo "ncu16 rc=$?"
DIVANS_B200_LPS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel_v2 -c 1 -o gpurun_out/${T}_dec8_8192 python tools/perf_probe.py --l-only --decode-once 8192 > gpurun_out/${T}_ncu8.log 2>&1; echo "ncu8 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-populations > gpurun_out/${T}_launches_bench.log 2>&1; echo "launch list rc=$?"
