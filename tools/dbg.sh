#!/bin/bash
# debugging battery, every step under its own timeout
export DIVANS_B200_DEBUG=1
echo "== frame only"; DIVANS_B200_SKIP_DECODE=1 timeout 120 python tools/dev_min.py 32 100 2>&1 | tail -5
echo "== lps32 len 1"; timeout 120 python tools/dev_min.py 32 1 2>&1 | tail -5
echo "== lps16 len 1"; timeout 120 python tools/dev_min.py 16 1 2>&1 | tail -5
echo "== sanitizer lps32 len 1"; timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python tools/dev_min.py 32 1 2>&1 | tail -40
