#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; T=r2_v5
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${T}_pytest.log
timeout 600 python tools/perf_probe.py 4096 --lz-all > gpurun_out/${T}_probe_4096.txt 2>&1; cat gpurun_out/${T}_probe_4096.txt
timeout 300 python tools/blocking_probe.py > gpurun_out/${T}_blocking.txt 2>&1; cat gpurun_out/${T}_blocking.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_set.py > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/${T}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 3000 python tools/sanitize_set.py > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/${T}_racecheck.log; grep -c "Error: Race" gpurun_out/${T}_racecheck.log
