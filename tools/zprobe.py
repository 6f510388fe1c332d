"""Dev probe: decode of LZ77 command streams only (for ncu captures of the command path).  python tools/zprobe.py [n_streams]"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import divans_b200
from divans_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng = divans_b200.Engine(0, 0, int(os.environ.get("DIVANS_B200_LPS", "0")))
blob, off, ln = synth.text_streams(n, 65536, seed=3)
raws = [blob[int(o):int(o + l)].tobytes() for o, l in zip(off, ln)]
cb, co, cl = divans_b200.lz77_cmds_batch(blob, off, ln, 16, 2, 4)
streams = eng.encode([cb[int(o):int(o + l)].tobytes() for o, l in zip(co, cl)], divans_b200.encode_options(window_size=16), cmds=True)
for _ in range(2):
    res = eng.decode(streams, [len(r) + 64 for r in raws])
    print("decode ms", eng.last_main_kernel_ms(), all(st == 0 and out == r for (st, out), r in zip(res, raws)))
