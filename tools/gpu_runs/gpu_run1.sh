#!/bin/bash
# round-2 visit 1: parity tests (incl. new ones), sanitizers on the reduced set, bench + probes of the round-1 kernel
cd /root/repo; mkdir -p gpurun_out; T=r2_v0
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${T}_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_set.py > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${T}_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_set.py > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/${T}_racecheck.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cat gpurun_out/${T}_bench.json | cut -c1-1500
timeout 600 python tools/perf_probe.py 4096 --lz-all > gpurun_out/${T}_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/${T}_probe.txt
