#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; T=r2_v3
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${T}_pytest.log
for L in 16 8; do
  DIVANS_B200_PREFETCH=1 DIVANS_B200_LPS=$L timeout 300 python tools/perf_probe.py 4096 --l-only > gpurun_out/${T}_probe_pf_l${L}_4096.txt 2>&1; echo "== PREFETCH lanes $L n=4096"; head -2 gpurun_out/${T}_probe_pf_l${L}_4096.txt
  DIVANS_B200_PREFETCH=1 DIVANS_B200_LPS=$L timeout 300 python tools/perf_probe.py 8192 --l-only > gpurun_out/${T}_probe_pf_l${L}_8192.txt 2>&1; echo "== PREFETCH lanes $L n=8192"; head -2 gpurun_out/${T}_probe_pf_l${L}_8192.txt
  DIVANS_B200_LPS=$L timeout 300 python tools/perf_probe.py 4096 --l-only > gpurun_out/${T}_probe_l${L}_4096.txt 2>&1; echo "== lanes $L n=4096"; head -2 gpurun_out/${T}_probe_l${L}_4096.txt
  DIVANS_B200_LPS=$L timeout 300 python tools/perf_probe.py 8192 --l-only > gpurun_out/${T}_probe_l${L}_8192.txt 2>&1; echo "== lanes $L n=8192"; head -2 gpurun_out/${T}_probe_l${L}_8192.txt
done
DIVANS_B200_LPS=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel_v2 -c 1 -o gpurun_out/${T}_dec16_4096 python tools/perf_probe.py 4096 --l-only --decode-once > gpurun_out/${T}_ncu16.log 2>&1; echo "ncu16 rc=$?"
DIVANS_B200_PREFETCH=1 DIVANS_B200_LPS=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel_v2 -c 1 -o gpurun_out/${T}_dec16pf_4096 python tools/perf_probe.py 4096 --l-only --decode-once > gpurun_out/${T}_ncu16pf.log 2>&1; echo "ncu16pf rc=$?"
