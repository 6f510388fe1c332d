#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_fin4_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_fin4_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
