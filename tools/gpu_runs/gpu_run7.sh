#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; T=r2_v6
timeout 300 python tools/zprobe.py 4096 > gpurun_out/${T}_z4096.txt 2>&1; cat gpurun_out/${T}_z4096.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel_v2 -c 1 -o gpurun_out/${T}_z1024 python tools/zprobe.py 1024 > gpurun_out/${T}_ncuz.log 2>&1; echo "ncu z rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel_v2 -c 1 -o gpurun_out/${T}_z4096 python tools/zprobe.py 4096 > gpurun_out/${T}_ncuz4096.log 2>&1; echo "ncu z4096 rc=$?"
timeout 600 compute-sanitizer --tool racecheck --print-limit 3000 python tools/sanitize_set.py > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/${T}_racecheck.log
