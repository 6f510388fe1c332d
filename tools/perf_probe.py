"""Dev probe (not a test, not the bench): device time of decode/encode for several stream populations.
  python tools/perf_probe.py [n_streams] [--l-only] [--lz-all] [--decode-once]     (DIVANS_B200_LPS selects the lane layout)"""
import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import divans_b200
from divans_b200 import synth

once = "--decode-once" in sys.argv

def run(name, raws, opts, eng, cmds=None):
    streams = eng.encode(cmds if cmds is not None else raws, opts, cmds=cmds is not None)
    enc_ms = eng.last_kernel_ms(); enc_model = eng.last_main_kernel_ms()
    caps = [len(r) + 64 for r in raws]
    for _ in range(1 if once else 2):
        res = eng.decode(streams, caps)
    dec_ms = eng.last_kernel_ms(); dec_main = eng.last_main_kernel_ms()
    ok = all(st == 0 and out == r for (st, out), r in zip(res, raws))
    tot = sum(len(r) for r in raws); comp = sum(len(s) for s in streams)
    print("%-34s n=%5d raw %7.1f MB ratio %.3f | decode %8.2f ms (main %8.2f) %7.0f MB/s | encode %8.2f ms (model %8.2f) %7.0f MB/s | ok=%s"
          % (name, len(raws), tot / 1e6, comp / tot, dec_ms, dec_main, tot / dec_ms / 1e3, enc_ms, enc_model, tot / enc_ms / 1e3, ok), flush=True)

eng = divans_b200.Engine(0, 0, int(os.environ.get("DIVANS_B200_LPS", "16")))
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if args else 4096
blob, off, ln = synth.text_streams(n, 65536, seed=3)
raws = [blob[int(o):int(o + l)].tobytes() for o, l in zip(off, ln)]
run("text 64KiB default", raws, divans_b200.encode_options(), eng)
if "--l-only" in sys.argv:
    sys.exit(0)
run("text 64KiB dcm=1", raws, divans_b200.encode_options(dynamic_context_mixing=1), eng)
run("text 64KiB dcm=2", raws, divans_b200.encode_options(dynamic_context_mixing=2), eng)
run("text 64KiB utf8 mix=1 ", raws, divans_b200.encode_options(literal_pred_mode=2, literal_mixing_value=1), eng)
if "--lz" in sys.argv or "--lz-all" in sys.argv:
    m = n if "--lz-all" in sys.argv else min(n, 512)
    cb, co, cl = divans_b200.lz77_cmds_batch(blob, off[:m], ln[:m], 16, 2, 4)
    cmds = [cb[int(o):int(o + l)].tobytes() for o, l in zip(co, cl)]
    run("text 64KiB lz77 cmds (w16)", raws[:m], divans_b200.encode_options(window_size=16), eng, cmds=cmds)
for p in (0.5, 0.9, 0.99):
    b2, o2, l2 = synth.bernoulli_streams(max(8, n // 16), 1 << 20, p, seed=5)
    r2 = [b2[int(o):int(o + l)].tobytes() for o, l in zip(o2, l2)]
    run("bernoulli p=%.2f 1MiB" % p, r2, divans_b200.encode_options(), eng)
