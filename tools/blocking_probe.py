"""Dev probe: where does the blocking host call spend its time?  (4096 x 64 KiB, pinned buffers)"""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import divans_b200
from divans_b200 import synth
eng = divans_b200.Engine(0, 0, 0)
n = 4096
blob, off, ln = synth.text_streams(n, 65536, seed=3)
cap = np.full(n, 65536 + 32768 + 70144, np.uint64); eoff = np.arange(n, dtype=np.uint64) * cap[0]
out = np.zeros(int(cap.sum()), np.uint8)
out_len, status = eng.encode_batch_host(blob, off, ln, out, eoff, cap, divans_b200.encode_options())
pad = (out_len + np.uint64(15)) & ~np.uint64(15); coff = np.zeros(n, np.uint64); coff[1:] = np.cumsum(pad)[:-1]
comp = np.zeros(int(pad.sum()) + 64, np.uint8)
for i in range(n): comp[int(coff[i]):int(coff[i] + out_len[i])] = out[int(eoff[i]):int(eoff[i] + out_len[i])]
h_in = torch.from_numpy(comp).pin_memory().numpy(); h_out = torch.zeros(int(ln.sum()) + 256, dtype=torch.uint8).pin_memory().numpy()
clen = out_len.astype(np.uint64)
for k in range(5):
    t0 = time.perf_counter(); ol, st = eng.decode_batch_host(h_in, coff, clen, h_out, off, ln); t1 = time.perf_counter()
    print("blocking call %.2f ms (kernels %.2f ms) ok=%s" % ((t1 - t0) * 1e3, eng.last_kernel_ms(), bool((st == 0).all() and (h_out[:blob.size] == blob).all())))
capx = ln + np.uint64(64)
offx = np.concatenate([[0], np.cumsum(capx)[:-1]]).astype(np.uint64)
h_out2 = torch.zeros(int(capx.sum()) + 256, dtype=torch.uint8).pin_memory().numpy()
for k in range(3):
    t0 = time.perf_counter(); ol, st = eng.decode_batch_host(h_in, coff, clen, h_out2, offx, capx); t1 = time.perf_counter()
    print("blocking call, regions with 64 spare bytes each: %.2f ms" % ((t1 - t0) * 1e3))
