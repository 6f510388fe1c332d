#!/bin/bash
# round-2 visit 2: first contact of the 8-lane engine
cd /root/repo; mkdir -p gpurun_out; T=r2_v1
./tools/ubench/prefetch_lat > gpurun_out/${T}_prefetch_lat.txt 2>&1; cat gpurun_out/${T}_prefetch_lat.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${T}_pytest.log
timeout 600 python tools/perf_probe.py 4096 --lz-all > gpurun_out/${T}_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/${T}_probe.txt
timeout 300 python tools/perf_probe.py 8192 > gpurun_out/${T}_probe8192.txt 2>&1; echo "probe rc=$?"; head -3 gpurun_out/${T}_probe8192.txt
