"""Multi-GPU sharding of a batch of independent streams: contiguous ranges of the stream index
balanced by compressed bytes; no data-path collective -- ranks only agree on the job totals."""
import numpy as np


def partition_by_bytes(lengths, n_parts):
    """[(begin, end)] * n_parts: contiguous index ranges whose byte sums are as even as a prefix-sum split allows."""
    lengths = np.asarray(lengths, dtype=np.uint64)
    n = lengths.size
    if n == 0:
        return [(0, 0)] * n_parts
    csum = np.concatenate([[0], np.cumsum(lengths.astype(np.float64))])
    total = csum[-1]
    cuts = [0]
    for k in range(1, n_parts):
        target = total * k / n_parts
        i = int(np.searchsorted(csum, target, side="left"))
        # pick the closer of the two neighbouring cut points
        if i > 0 and abs(csum[i - 1] - target) <= abs(csum[min(i, n)] - target):
            i -= 1
        cuts.append(max(cuts[-1], min(i, n)))
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(n_parts)]


def reduce_job(units_done, seconds):
    """All-reduce the per-rank (units, time) into (sum of units, max time) -- the only collective of a sharded job."""
    import torch.distributed as dist
    u = units_done.clone()
    t = seconds.clone()
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(u.item()), float(t.item())


class ShardedDecoder:
    """Decode ONE host batch of independent streams on all ranks of a torch.distributed job (the reference decodes its
    objects independently, src/divans_decompressor.rs:258-276, so a batch shards by stream index).

    The root rank owns the batch (host arrays); every rank calls ``decode`` collectively:
      1. root: contiguous stream ranges balanced by compressed bytes (``partition_by_bytes``), descriptors broadcast;
      2. scatter: root copies the batch to its GPU once, peers receive their byte range (NCCL send/recv over NVLink; no
         collective touches the decode itself);
      3. every rank decodes its shard with ``decode_fn`` (default: the engine's device API, inputs and outputs in HBM);
      4. gather: peers send their output range, out_len and status back; root assembles the batch result.
    With world size 1 (or no initialised process group) this is a plain device decode.  ``decode_fn(in, in_off, in_len, out,
    out_off, out_cap) -> (out_len, status)`` works on torch tensors of this rank's device and exists so that the plumbing can
    be tested on CPU (gloo) with a stand-in decoder."""

    def __init__(self, engine=None, device=None, root=0, group=None, decode_fn=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.engine, self.root, self.group = engine, root, group
        self.on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.on else 0
        self.world = dist.get_world_size(group) if self.on else 1
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.decode_fn = decode_fn or self._engine_decode
        self.last = {}
        self._host_out = None      # root: pinned staging buffer for the gathered output (kept between calls)

    def _engine_decode(self, d_in, d_in_off, d_in_len, d_out, d_out_off, d_out_cap):
        torch = self.torch
        n = int(d_in_off.numel())
        d_out_len = torch.zeros(n, dtype=torch.int64, device=self.device)
        d_status = torch.full((n,), 3, dtype=torch.int32, device=self.device)
        if n:
            st = torch.cuda.current_stream(self.device)
            self.engine.decode_batch_device(d_in.data_ptr(), d_in_off.data_ptr(), d_in_len.data_ptr(), d_out.data_ptr(), d_out_off.data_ptr(),
                                            d_out_cap.data_ptr(), d_out_len.data_ptr(), d_status.data_ptr(), n, int(d_in.numel()), 0, st.cuda_stream)
        return d_out_len, d_status

    def _p2p(self, ops):
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()

    def decode(self, blob=None, in_off=None, in_len=None, out_cap=None):
        """root passes numpy/torch host arrays (blob uint8, in_off / in_len / out_cap uint64-like); the other ranks pass nothing.
        Returns on root: (out uint8 tensor on host, out_off, out_len, status) -- stream i at out[out_off[i] : +out_len[i]].  On a
        GPU job `out` is a view of a pinned staging buffer that the next call overwrites."""
        torch, dist = self.torch, self.dist
        dev, root, R, W = self.device, self.root, self.rank, self.world
        i64 = lambda a: torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.int64)))
        # ---- descriptors ----
        if R == root:
            in_off_t, in_len_t, cap_t = i64(in_off), i64(in_len), i64(out_cap)
            n = int(in_len_t.numel())
            parts = partition_by_bytes(in_len_t.numpy().astype(np.uint64), W)
            hdr = torch.tensor([n] + [x for p in parts for x in p], dtype=torch.int64)
        else:
            hdr = torch.zeros(1 + 2 * W, dtype=torch.int64)
        hdr = hdr.to(dev)
        if self.on:
            dist.broadcast(hdr, src=root, group=self.group)
        hdr_h = hdr.cpu()
        n = int(hdr_h[0])
        parts = [(int(hdr_h[1 + 2 * r]), int(hdr_h[2 + 2 * r])) for r in range(W)]
        meta = torch.stack([in_off_t, in_len_t, cap_t]).to(dev) if R == root else torch.zeros((3, n), dtype=torch.int64, device=dev)
        if self.on:
            dist.broadcast(meta, src=root, group=self.group)
        a, b = parts[R]
        g_in_off, g_in_len, g_cap = meta[0], meta[1], meta[2]
        out_off_all = torch.cumsum(g_cap, 0) - g_cap                      # outputs packed back to back in stream order
        byte_lo = [int(g_in_off[p[0]]) if p[1] > p[0] else 0 for p in parts]
        byte_hi = [int(g_in_off[p[1] - 1] + g_in_len[p[1] - 1]) if p[1] > p[0] else 0 for p in parts]
        out_lo = [int(out_off_all[p[0]]) if p[1] > p[0] else 0 for p in parts]
        out_hi = [int(out_off_all[p[1] - 1] + g_cap[p[1] - 1]) if p[1] > p[0] else 0 for p in parts]
        # ---- scatter the compressed bytes ----
        if R == root:
            d_all = torch.as_tensor(blob).to(dev, non_blocking=True) if not torch.is_tensor(blob) else blob.to(dev, non_blocking=True)
            ops = [dist.P2POp(dist.isend, d_all[byte_lo[r]:byte_hi[r]], r, self.group) for r in range(W) if r != root and byte_hi[r] > byte_lo[r]]
            self._p2p(ops)
            d_in = d_all[byte_lo[R]:byte_hi[R]]
        else:
            d_in = torch.empty(byte_hi[R] - byte_lo[R], dtype=torch.uint8, device=dev)
            if d_in.numel():
                self._p2p([dist.P2POp(dist.irecv, d_in, root, self.group)])
        # ---- decode the shard where it is ----
        my_in_off = (g_in_off[a:b] - byte_lo[R]).contiguous()
        my_in_len = g_in_len[a:b].contiguous()
        my_cap = g_cap[a:b].contiguous()
        my_out_off = (out_off_all[a:b] - out_lo[R]).contiguous()
        d_out = torch.zeros(max(1, out_hi[R] - out_lo[R]), dtype=torch.uint8, device=dev)
        my_len, my_status = self.decode_fn(d_in, my_in_off, my_in_len, d_out, my_out_off, my_cap)
        if dev.type == "cuda":
            torch.cuda.current_stream(dev).synchronize()
        # ---- gather ----
        if R == root:
            out_all = torch.zeros(max(1, out_hi[-1] if W else 0, max(out_hi)), dtype=torch.uint8, device=dev)
            len_all = torch.zeros(n, dtype=torch.int64, device=dev)
            st_all = torch.full((n,), 3, dtype=torch.int32, device=dev)
            out_all[out_lo[R]:out_hi[R]] = d_out[: out_hi[R] - out_lo[R]]
            len_all[a:b] = my_len
            st_all[a:b] = my_status
            ops = []
            for r in range(W):
                if r == root or parts[r][1] <= parts[r][0]:
                    continue
                ops += [dist.P2POp(dist.irecv, out_all[out_lo[r]:out_hi[r]], r, self.group),
                        dist.P2POp(dist.irecv, len_all[parts[r][0]:parts[r][1]], r, self.group),
                        dist.P2POp(dist.irecv, st_all[parts[r][0]:parts[r][1]], r, self.group)]
            self._p2p(ops)
            self.last = dict(parts=parts, shard_bytes=[byte_hi[r] - byte_lo[r] for r in range(W)])
            if dev.type != "cuda":
                return out_all, out_off_all, len_all, st_all
            # one D2H into pinned memory (a pageable destination would cost more than the decode)
            if self._host_out is None or self._host_out.numel() < out_all.numel():
                self._host_out = torch.empty(out_all.numel(), dtype=torch.uint8, pin_memory=True)
            host = self._host_out[: out_all.numel()]
            host.copy_(out_all, non_blocking=True)
            res = (host, out_off_all.cpu(), len_all.cpu(), st_all.cpu())
            torch.cuda.current_stream(dev).synchronize()
            return res
        if b > a:
            self._p2p([dist.P2POp(dist.isend, d_out[: out_hi[R] - out_lo[R]], root, self.group),
                       dist.P2POp(dist.isend, my_len, root, self.group), dist.P2POp(dist.isend, my_status, root, self.group)])
        return None
