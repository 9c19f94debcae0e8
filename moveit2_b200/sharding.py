"""Multi-GPU plumbing for the batched validity check: one process per GPU (torch.distributed), states sharded in
contiguous blocks, fields replicated, per-rank validity bitmaps exchanged with ONE all-gather per batch.

The path has no other data exchange: every state's FK + checks depend only on replicated constants (robot model,
masks, the two fields).  The bitmap all-gather is N/8 bytes in total (1.25 MB for 10 M states): launch latency, not
NVLink bandwidth, is what matters, so exactly one collective is issued per batch, on the stream the kernel ran on.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_total, world, rank):
    """Contiguous block [lo, hi) of rank `rank`.  Block size is a multiple of 32 states so every rank's bitmap is
    whole 32-bit words and the gathered bitmap is simply the concatenation."""
    per = ((n_total + world - 1) // world + 31) // 32 * 32
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total), per


class ShardedStateValidity:
    """check_local(q_local, bitmap_local): fills `bitmap_local` (uint8 tensor, 4 * ceil(n_local / 32) bytes, bit i%8 of
    byte i//8 = state i valid) for the [n_local, V] float64 tensor `q_local`, asynchronously on the current stream.
    On a GPU that is StateValidityEngine.check_batch_device; the CPU tests plug in the oracle.

    overlap=True (CUDA only): the all-gather of batch i runs on a side stream while the kernels of batch i + 1 run on the
    compute stream -- bitmap slices and gathered buffers are double-buffered, the only cross-stream traffic is one event
    each way per batch.  The collective's launch latency (15-25 us for 125 KB slices) then never sits on the critical path
    of a stream of batches; `wait()` joins the last one."""

    def __init__(self, check_local, n_total, device, group=None, overlap=False):
        self.check_local = check_local
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = n_total
        self.lo, self.hi, self.per = shard_bounds(n_total, self.world, self.rank)
        self.bytes_per_rank = self.per // 8
        self.overlap = bool(overlap) and self.world > 1 and torch.device(device).type == "cuda"
        nbuf = 2 if self.overlap else 1
        self._local = [torch.zeros(self.bytes_per_rank, dtype=torch.uint8, device=device) for _ in range(nbuf)]
        self._gathered = [torch.zeros(self.world * self.bytes_per_rank, dtype=torch.uint8, device=device)
                          for _ in range(nbuf)] if self.world > 1 else self._local
        self._i = 0
        self.local, self.gathered = self._local[0], self._gathered[0]
        if self.overlap:
            self.comm_stream = torch.cuda.Stream(device=device)
            self._computed = [torch.cuda.Event() for _ in range(2)]
            self._gathered_ev = [torch.cuda.Event() for _ in range(2)]
            self._busy = [False, False]

    def check(self, q_local):
        """Returns the gathered bitmap tensor (every rank holds all of it; with overlap=True it is complete after wait(),
        or once the NEXT-but-one check() has been issued)."""
        n_local = self.hi - self.lo
        assert q_local.shape[0] == n_local
        k = self._i & (len(self._local) - 1)
        self._i += 1
        self.local, self.gathered = self._local[k], self._gathered[k]
        if self.overlap:
            cur = torch.cuda.current_stream()
            if self._busy[k]:
                cur.wait_event(self._gathered_ev[k])      # this buffer pair's previous gather must have read `local`
            if n_local:
                self.check_local(q_local, self.local)
            self._computed[k].record(cur)
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(self._computed[k])
                dist.all_gather_into_tensor(self.gathered, self.local, group=self.group)
                self._gathered_ev[k].record(self.comm_stream)
            self._busy[k] = True
            return self.gathered
        if n_local:
            self.check_local(q_local, self.local)
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.local, group=self.group)
        return self.gathered

    def wait(self):
        """Order the current stream behind every gather issued so far (overlap=True; a no-op otherwise)."""
        if self.overlap:
            cur = torch.cuda.current_stream()
            for k in range(2):
                if self._busy[k]:
                    cur.wait_event(self._gathered_ev[k])

    def valid_mask(self):
        """bool [n_total] on the host, from the last gathered bitmap."""
        self.wait()
        bits = np.unpackbits(self.gathered.cpu().numpy(), bitorder="little")
        out = np.zeros(self.n_total, bool)
        for r in range(self.world):
            lo, hi, per = shard_bounds(self.n_total, self.world, r)
            out[lo:hi] = bits[r * per: r * per + (hi - lo)]
        return out


def broadcast_field(engine, which, device, src=0, group=None):
    """The field `which` of rank `src`'s engine -> every rank's engine: export to one device buffer, ONE NCCL broadcast over
    NVLink, import (the EDT runs on one GPU, the compact field is replicated)."""
    buf = torch.empty(engine.field_export_size(), dtype=torch.uint8, device=device)
    s = torch.cuda.current_stream().cuda_stream
    rank = dist.get_rank(group)
    if rank == src:
        engine.field_export_device(buf.data_ptr(), which, s)
    dist.broadcast(buf, src=src, group=group)
    if rank != src:
        engine.field_import_device(buf.data_ptr(), which, s)
    torch.cuda.current_stream().synchronize()
    return buf.numel()
