"""Batch sharding over the GPUs of one node and the result gather.

Images are independent (the reference loops over them, nms.py:91), so each rank post-processes a contiguous slice of
the batch with no exchange during compute - the same partition the reference's DDP validation uses
(data/build.py:171-188 ``ContiguousDistributedSampler``).  The only collective is the gather of the per-image counts
and the fixed-stride detection rows, mirroring ``dist.gather_object(stats)`` in models/yolo/detect/val.py:226-240 but
as ONE packed all_gather (<= B*(max_det*(6+extra)+1)*4 bytes) instead of pickled Python objects.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of the batch for `rank`; the remainder goes to the first ranks."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_results(rows: torch.Tensor, count: torch.Tensor, pad_batch: int | None = None) -> torch.Tensor:
    """(b, max_det, cols) fp32 rows + (b,) int32 counts -> one (pad_batch, 1 + max_det*cols) fp32 buffer.

    The count travels bit-cast in column 0, so one collective moves everything."""
    b, max_det, cols = rows.shape
    pb = b if pad_batch is None else pad_batch
    packed = torch.zeros((pb, 1 + max_det * cols), dtype=torch.float32, device=rows.device)
    packed[:b, 0] = count.view(torch.float32)
    packed[:b, 1:] = rows.reshape(b, -1)
    return packed


def unpack_results(packed: torch.Tensor, max_det: int, cols: int):
    count = packed[:, 0].contiguous().view(torch.int32)
    rows = packed[:, 1:].reshape(packed.shape[0], max_det, cols)
    return rows, count


def gather_results(rows: torch.Tensor, count: torch.Tensor, batch: int, group=None, out: torch.Tensor | None = None,
                   async_op: bool = False):
    """all_gather the per-rank shards into the full-batch (B, max_det, cols) rows and (B,) counts on every rank.

    rows/count are this rank's shard (shard_range order).  Returns (rows, count) or, with async_op, (work, finish)
    where finish() -> (rows, count).
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    _, max_det, cols = rows.shape
    per = -(-batch // world)  # padded shard size
    packed = pack_results(rows, count, per)
    full = out if out is not None else torch.empty((world * per, packed.shape[1]), dtype=torch.float32, device=rows.device)
    work = dist.all_gather_into_tensor(full, packed, group=group, async_op=async_op)

    def finish():
        pieces = []
        for r in range(world):
            lo, hi = shard_range(batch, r, world)
            pieces.append(full[r * per: r * per + (hi - lo)])
        return unpack_results(torch.cat(pieces, 0), max_det, cols)

    if async_op:
        return work, finish
    return finish()


def gather_packed(packed: torch.Tensor, out: torch.Tensor, group=None, async_op: bool = False):
    """all_gather of a plan's own result buffer (`NmsPlan.packed`: rows and counts of one rank in ONE allocation), so the
    gather needs no packing kernels at all.  `out` is (world, packed.numel()); equal shard sizes on every rank."""
    return dist.all_gather_into_tensor(out, packed, group=group, async_op=async_op)


def split_packed(out: torch.Tensor, batch_per_rank: int, max_det: int, cols: int):
    """(world, n) gathered buffer -> rows (world*B, max_det, cols) view pieces and counts (world*B,) int32."""
    world = out.shape[0]
    nrow = batch_per_rank * max_det * cols
    rows = out[:, :nrow].reshape(world * batch_per_rank, max_det, cols)
    count = out[:, nrow:].contiguous().view(torch.int32).reshape(world * batch_per_rank)
    return rows, count


class PeerGather:
    """One-sided gather of the result buffers over NVLink peer memory (``ypb_nms_out.peer_*``).

    Every rank owns a symmetric buffer ``[world x packed | world arrival flags]`` that all ranks of the node map
    (``torch.distributed._symmetric_memory``: CUDA VMM handles exchanged once at construction).  Rank s's suppression kernel
    stores its kept rows + counts directly into slot s of EVERY rank's buffer and then raises flag s there; nothing on the
    step's critical path waits for another rank, unlike an all_gather whose kernels rendezvous.  ``wait()`` enqueues the
    consumer-side spin (``ypb_peer_wait``) after which ``gathered()`` holds the results of the matching launch of all ranks.
    Construction is collective (same order on every rank).  One instance serves one stream / lane.
    Consumer contract: slot contents are overwritten by the owner's next launch on this lane; consume (or copy out) on the
    lane's stream before enqueueing the launch after next.
    """

    def __init__(self, packed_numel: int, nrow: int, device, group=None):
        import torch.distributed._symmetric_memory as symm

        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        from . import _cabi

        if self.world > _cabi.MAX_PEERS:
            raise ValueError(f"peer gather spans one node: world {self.world} > {_cabi.MAX_PEERS}")
        self.numel, self.nrow = int(packed_numel), int(nrow)
        self.slot = (self.numel + 3) // 4 * 4  # 16-byte aligned slots
        total = self.world * self.slot + 16
        self.buf = symm.empty(total, dtype=torch.float32, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        try:
            self.handle = symm.rendezvous(self.buf, group)
        except Exception:
            symm.enable_symm_mem_for_group(group.group_name)
            self.handle = symm.rendezvous(self.buf, group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.state = torch.zeros(4, dtype=torch.int32, device=device)
        self.flags = self.buf[self.world * self.slot: self.world * self.slot + 16].view(torch.int32)
        self.my_packed = self.buf[self.rank * self.slot: self.rank * self.slot + self.numel]
        self.peer_rows = [p + self.rank * self.slot * 4 for p in ptrs]
        self.peer_count = [p + self.nrow * 4 for p in self.peer_rows]
        self.peer_flag = [p + self.world * self.slot * 4 for p in ptrs]
        dist.barrier(group)  # every rank has zeroed its buffer before anyone stores into it

    def bind(self, out) -> None:
        """Fill the peer fields of a ``_cabi.NmsOut``."""
        out.num_peers, out.my_rank = self.world, self.rank
        for i in range(self.world):
            out.peer_rows[i], out.peer_count[i], out.peer_flag[i] = self.peer_rows[i], self.peer_count[i], self.peer_flag[i]
        out.peer_state = self.state.data_ptr()

    def wait(self, lag: int = 0) -> None:
        """lag=0: the latest launch of every rank has landed; lag=k: the launch k launches back (pipelined gather)."""
        from . import _cabi

        rc = _cabi.load().ypb_peer_wait(self.flags.data_ptr(), self.world, self.state.data_ptr(), int(lag),
                                        _cabi.stream_ptr(self.buf.device))
        _cabi.check(rc, "ypb_peer_wait")

    def gathered(self, batch_per_rank: int, max_det: int, cols: int):
        """(world*B, max_det, cols) rows and (world*B,) int32 counts of the last waited-for launch (views / small copy)."""
        g = self.buf[: self.world * self.slot].view(self.world, self.slot)[:, : self.numel]
        rows = g[:, : self.nrow].reshape(self.world * batch_per_rank, max_det, cols)
        count = g[:, self.nrow:].contiguous().view(torch.int32).reshape(self.world * batch_per_rank)
        return rows, count
