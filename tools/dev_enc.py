import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import divans_b200
from divans_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
SB = 65536
eng = divans_b200.Engine(0, 0, 16)
blob, off, ln = synth.text_streams(n, SB, seed=1)
dev = torch.device("cuda", 0)
d_raw = torch.from_numpy(blob).to(dev)
d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
d_len = torch.from_numpy(ln.astype(np.int64)).to(dev)
ecap = SB + SB // 2 + 70144
d_eout = torch.zeros(n * ecap, dtype=torch.uint8, device=dev)
d_eoff = torch.arange(n, dtype=torch.int64, device=dev) * ecap
d_ecap = torch.full((n,), ecap, dtype=torch.int64, device=dev)
d_elen = torch.zeros(n, dtype=torch.int64, device=dev)
d_est = torch.zeros(n, dtype=torch.int32, device=dev)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
for it in range(3):
    eng.encode_batch_device(n, d_raw.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), SB, d_eout.data_ptr(), d_eoff.data_ptr(),
                            d_ecap.data_ptr(), d_elen.data_ptr(), d_est.data_ptr(), divans_b200.encode_options(), stream.cuda_stream)
    torch.cuda.synchronize()
    print("iter", it, "ok", int(d_est.abs().sum()), int(d_elen.sum()))
