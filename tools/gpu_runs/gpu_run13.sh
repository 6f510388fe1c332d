#!/bin/bash
# final single-GPU pass of the round: smoke, bench line, launch list, full ncu capture of the decode kernel, sanitizers
cd /root/repo; mkdir -p gpurun_out; T=r2_final
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-populations > gpurun_out/${T}_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel_v2 -c 1 -o gpurun_out/${T}_dec16_4096 python tools/perf_probe.py --l-only --decode-once 4096 > gpurun_out/${T}_ncu16.log 2>&1; echo "ncu rc=$?"
timeout 900 compute-sanitizer --tool memcheck --print-limit 50 python tools/sanitize_set.py > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/${T}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 3000 python tools/sanitize_set.py > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/${T}_racecheck.log
