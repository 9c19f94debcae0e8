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
    On a GPU that is StateValidityEngine.check_batch_device; the CPU tests plug in the oracle."""

    def __init__(self, check_local, n_total, device, group=None):
        self.check_local = check_local
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = n_total
        self.lo, self.hi, self.per = shard_bounds(n_total, self.world, self.rank)
        self.bytes_per_rank = self.per // 8
        self.local = torch.zeros(self.bytes_per_rank, dtype=torch.uint8, device=device)
        self.gathered = torch.zeros(self.world * self.bytes_per_rank, dtype=torch.uint8, device=device) \
            if self.world > 1 else self.local

    def check(self, q_local):
        """Returns the gathered bitmap tensor (every rank holds all of it)."""
        n_local = self.hi - self.lo
        assert q_local.shape[0] == n_local
        if n_local:
            self.check_local(q_local, self.local)
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.local, group=self.group)
        return self.gathered

    def valid_mask(self):
        """bool [n_total] on the host, from the last gathered bitmap."""
        bits = np.unpackbits(self.gathered.cpu().numpy(), bitorder="little")
        out = np.zeros(self.n_total, bool)
        for r in range(self.world):
            lo, hi, per = shard_bounds(self.n_total, self.world, r)
            out[lo:hi] = bits[r * per: r * per + (hi - lo)]
        return out
