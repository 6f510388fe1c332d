#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; T=r2_v15
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel_v2 -c 1 -o gpurun_out/${T}_z4096 python tools/zprobe.py 4096 > gpurun_out/${T}_ncuz4096.log 2>&1; echo "ncu z4096 rc=$?"
