#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; T=r2_v13
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
timeout 1200 $TR bench.py --gpus 4 --steps 5 --warmup 3 --skip-populations > gpurun_out/${T}_bench_n4.json 2> gpurun_out/${T}_bench_n4.err; echo "bench n4 rc=$?"; cut -c1-300 gpurun_out/${T}_bench_n4.json; grep -o '"scattered_e2e": {[^}]*}' gpurun_out/${T}_bench_n4.json | cut -c1-400
tail -3 gpurun_out/${T}_bench_n4.err | cut -c1-300
