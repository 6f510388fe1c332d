#!/bin/bash
# ncu capture of the 8-lane kernel at the N > 1 workload (8192 streams per GPU), then the final N = 1 bench line (with roofline.traffic)
cd /root/repo; mkdir -p gpurun_out; T=r2_final
DIVANS_B200_LPS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel_v2 -c 1 -o gpurun_out/${T}_dec8_8192 python tools/perf_probe.py --l-only --decode-once 8192 > gpurun_out/${T}_ncu8.log 2>&1; echo "ncu8 rc=$?"
timeout 1200 python bench.py > gpurun_out/${T}_bench2.json 2> gpurun_out/${T}_bench2.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${T}_bench2.json; grep -o '"traffic": [^,]*' gpurun_out/${T}_bench2.json
timeout 600 python bench.py --impl reference > gpurun_out/${T}_reference_arm.json 2> gpurun_out/${T}_reference_arm.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/${T}_reference_arm.json
