"""Dev probe: decode of LZ77 command streams only (for ncu captures of the command path)."""
import sys, numpy as np
sys.path.insert(0, "/root/repo")
import divans_b200
from divans_b200 import synth
from oracle import oracle_py as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
import os
eng = divans_b200.Engine(0, 0, int(os.environ.get("DIVANS_B200_LPS", "16")))
blob, off, ln = synth.text_streams(n, 65536, seed=3)
raws = [blob[int(o):int(o + l)].tobytes() for o, l in zip(off, ln)]
streams = [O.Commands.lz77(r, 16, 2, 4).encode(O.options(window_size=16)) for r in raws]
for _ in range(2):
    res = eng.decode(streams, [len(r) + 64 for r in raws])
    print("decode ms", eng.last_main_kernel_ms(), all(st == 0 and out == r for (st, out), r in zip(res, raws)))
