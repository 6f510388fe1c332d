#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; T=r2_v8
for CO in -1 auto 28; do
  for N in 3552 4096 4736 8192; do
    if [ "$CO" = auto ]; then unset DIVANS_B200_CARVEOUT; else export DIVANS_B200_CARVEOUT=$CO; fi
    echo "== carveout $CO n=$N" | tee -a gpurun_out/${T}_sweep.txt
    timeout 300 python tools/perf_probe.py --l-only --decode-once $N 2>&1 | cut -c1-260 | tee -a gpurun_out/${T}_sweep.txt
  done
done
for CO in -1 auto; do
  if [ "$CO" = auto ]; then unset DIVANS_B200_CARVEOUT; else export DIVANS_B200_CARVEOUT=$CO; fi
  echo "== Z carveout $CO" | tee -a gpurun_out/${T}_z.txt; timeout 300 python tools/zprobe.py 4096 2>&1 | tee -a gpurun_out/${T}_z.txt
done
unset DIVANS_B200_CARVEOUT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel_v2 -c 1 -o gpurun_out/${T}_dec16_4096 python tools/perf_probe.py --l-only --decode-once 4096 > gpurun_out/${T}_ncu16.log 2>&1; echo "ncu rc=$?"
