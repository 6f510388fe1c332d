#!/usr/bin/env python3
"""Developer scratch: oracle-encode a few stream families, decode on the GPU, compare, print rough timings."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_py as O
import divans_b200 as D
from divans_b200 import synth

lps = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n_big = int(sys.argv[2]) if len(sys.argv) > 2 else 512
eng = D.Engine(0, 0, lps)

def check(name, streams, raws):
    t = time.time()
    res = eng.decode(streams, [len(r) + 64 for r in raws])
    dt = time.time() - t
    bad = [(i, st, len(out), len(r)) for i, ((st, out), r) in enumerate(zip(res, raws)) if st != 0 or out != r]
    print("%-28s n=%4d  %s  host-call %.1f ms kernel %.2f ms" % (name, len(streams), "OK" if not bad else "MISMATCH %s" % bad[:4], dt * 1e3, eng.last_kernel_ms()))
    if bad:
        i = bad[0][0]
        st, out = res[i]
        r = raws[i]
        k = next((j for j in range(min(len(out), len(r))) if out[j] != r[j]), min(len(out), len(r)))
        print("   first diff at byte", k, "of", len(r), "status", st)
    return not bad

ok = True
text = synth.text_corpus(1 << 20)
# 1. tiny + literal-only
raws = [text[:1], text[:7], text[:8], text[:9], text[:100], text[:4097], text[:65536], b"", bytes(range(256)) * 3]
ok &= check("literal-only small", [O.encode_raw(r) for r in raws], raws)
# 2. pred modes / mixing values
for pm in range(4):
    for mv in [0, 1, 2, 3, 4, 5, 6, 7, 8]:
        raws = [text[1000 * k: 1000 * k + 3000] for k in range(2)]
        blob = np.frombuffer(b"".join(raws), np.uint8)
        out, off, ln = O.encode_batch(blob, [0, 3000], [3000, 3000], O.options(), 1, False, pm, mv)
        streams = [out[int(o):int(o + l)].tobytes() for o, l in zip(off, ln)]
        ok &= check("pm=%d mix=%d" % (pm, mv), streams, raws)
# 3. mixing math 2
for mixv in [0, 1, 2, 4]:
    raws = [text[5000:9000], text[100:10100]]
    streams = []
    for r in raws:
        blob = np.frombuffer(r, np.uint8)
        out, off, ln = O.encode_batch(blob, [0], [len(r)], O.options(dynamic_context_mixing=2), 1, False, 2, mixv)
        streams.append(out[: int(ln[0])].tobytes())
    ok &= check("mixing=2 mv=%d" % mixv, streams, raws)
# 4. lz77
for win in [10, 16, 22]:
    raws = [text[:20000], text[3000:70000], (text[:300] * 50)]
    streams = [O.Commands.lz77(r, window=win).encode(O.options(window_size=win)) for r in raws]
    ok &= check("lz77 window=%d" % win, streams, raws)
raws = [text[:30000]]
streams = [O.Commands.lz77(r, window=16).encode(O.options(window_size=16, dynamic_context_mixing=2))for r in raws]
ok &= check("lz77 mixing=2", streams, raws)
# 5. reference IR fixtures if present (dev container only)
ref = "/root/reference/testdata/"
if os.path.exists(ref):
    for name in ["alice29", "alice29-priors", "alice29-q11", "asyoulik", "random_then_unicode", "ends_with_truncated_dictionary"]:
        raw = open(ref + (name.split("-")[0] if name.startswith("alice") else name), "rb").read()
        c = O.Commands.from_ir(open(ref + name + ".ir", "rb").read())
        for mix in [0, 2]:
            for cm in [1, 0]:
                ok &= check("%s mix%d cm%d" % (name[:18], mix, cm), [c.encode(O.options(dynamic_context_mixing=mix, use_context_map=cm))], [raw])
# 6. bulk throughput
blob, off, ln = synth.text_streams(n_big, 65536)
t = time.time(); enc, eoff, elen = O.encode_batch(blob, off, ln, O.options(), os.cpu_count()); te = time.time() - t
print("oracle encode %d streams: %.2fs, ratio %.3f" % (n_big, te, elen.sum() / blob.size))
streams = [enc[int(o):int(o + l)].tobytes() for o, l in zip(eoff, elen)]
raws = [blob[int(o):int(o + l)].tobytes() for o, l in zip(off, ln)]
ok &= check("bulk text 64KiB", streams, raws)
ok &= check("bulk text 64KiB (2nd)", streams, raws)
ms = eng.last_kernel_ms()
print("bulk kernel: %.2f ms -> %.1f MB/s decompressed" % (ms, blob.size / ms / 1e3))
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
