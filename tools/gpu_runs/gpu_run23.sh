#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; T=r2_fin2
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/${T}_bench.json; grep -o '"traffic": [^,]*' gpurun_out/${T}_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/${T}_reference_arm.json 2> gpurun_out/${T}_reference_arm.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/${T}_reference_arm.json
