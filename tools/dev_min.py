import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_py as O
import divans_b200 as D
from divans_b200 import synth
lps = int(sys.argv[1]) if len(sys.argv) > 1 else 32
eng = D.Engine(0, 0, lps)
text = synth.text_corpus(1 << 16)
raws = [text[:int(a)] for a in sys.argv[2:]] or [text[:100]]
res = eng.decode([O.encode_raw(r) for r in raws], [len(r) + 64 for r in raws])
for (st, out), r in zip(res, raws):
    print(st, len(out), len(r), out == r)
