#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --skip-populations > gpurun_out/r2_fin4_bench_n2.json 2> gpurun_out/r2_fin4_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-200 gpurun_out/r2_fin4_bench_n2.json; grep -o '"traffic": [^,]*' gpurun_out/r2_fin4_bench_n2.json; grep -o '"scattered_e2e": {"value": [^,]*' gpurun_out/r2_fin4_bench_n2.json
