#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; T=r2_v9
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 1200 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-1500 gpurun_out/${T}_bench_n2.json
timeout 900 $TR bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_ref_n2.json 2> gpurun_out/${T}_ref_n2.err; echo "ref n2 rc=$?"; cut -c1-800 gpurun_out/${T}_ref_n2.json
timeout 900 $TR bench.py --gpus 2 --workload entropy --steps 6 > gpurun_out/${T}_entropy_n2.json 2> gpurun_out/${T}_entropy_n2.err; echo "entropy n2 rc=$?"; cut -c1-1200 gpurun_out/${T}_entropy_n2.json
tail -5 gpurun_out/${T}_bench_n2.err | cut -c1-300
