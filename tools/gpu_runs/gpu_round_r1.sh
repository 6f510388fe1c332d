#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, one full ncu capture of the decode kernel.  Usage: tools/gpu_round.sh <tag>
cd /root/repo
TAG=${1:-dev}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" ; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -c 1 -o gpurun_out/${TAG}_decode python bench.py --streams 1024 --steps 1 --warmup 3 --skip-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:encode_model_kernel -c 1 -o gpurun_out/${TAG}_encode python bench.py --streams 1024 --steps 1 --warmup 3 --skip-cpu > gpurun_out/${TAG}_ncu_full_enc.log 2>&1; echo "ncu full enc rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -c 1 -o gpurun_out/${TAG}_decode4096 python bench.py --streams 4096 --steps 1 --warmup 3 --skip-cpu > gpurun_out/${TAG}_ncu_full4096.log 2>&1; echo "ncu full 4096 rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
ls -la gpurun_out | tail -14
