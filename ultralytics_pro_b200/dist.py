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

    Every rank owns a symmetric buffer ``[depth x world x packed | world arrival flags | world acknowledgements]`` that all
    ranks of the node map (``torch.distributed._symmetric_memory``: CUDA VMM handles exchanged once at construction).  Rank
    s's suppression kernel stores its kept rows + counts of launch number q directly into slot s of ring entry ``q % depth`` of
    EVERY rank's buffer and then raises flag s there; nothing on the step's critical path waits for another rank's kernel,
    unlike an all_gather whose kernels rendezvous.  ``wait(lag)`` enqueues the consumer side (``ypb_peer_wait``): it releases
    the entry the previous ``wait`` handed out (an acknowledgement written into every producer's buffer - a producer never
    overwrites an entry a peer has not released), then spins until launch ``latest - lag`` of every rank has landed, and
    leaves that entry's index in ``slot_index`` (device int64) for the consumer's kernels: ``entry()`` / ``gathered()``.
    The entry stays valid until the next ``wait`` on this instance executes.  ``depth = 3`` serves ``lag`` 0 and 1.
    Construction is collective (same order on every rank).  One instance serves one stream / lane.
    """

    DEPTH = 3

    def __init__(self, packed_numel: int, nrow: int, device, group=None):
        import torch.distributed._symmetric_memory as symm

        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        from . import _cabi

        if self.world > _cabi.MAX_PEERS:
            raise ValueError(f"peer gather spans one node: world {self.world} > {_cabi.MAX_PEERS}")
        self.numel, self.nrow, self.depth = int(packed_numel), int(nrow), self.DEPTH
        self.slot = (self.numel + 3) // 4 * 4  # 16-byte aligned slots
        self.entry = self.world * self.slot     # floats per ring entry
        ring = self.depth * self.entry
        total = ring + 32                       # + 16 arrival flags + 16 acknowledgements (int32)
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
        self.slot_index = torch.zeros(1, dtype=torch.int64, device=device)
        self.flags = self.buf[ring: ring + 16].view(torch.int32)
        self.acks = self.buf[ring + 16: ring + 32].view(torch.int32)
        self.my_packed = torch.empty(self.numel, dtype=torch.float32, device=device)  # the plan's own result buffer
        self.peer_rows = [p + self.rank * self.slot * 4 for p in ptrs]                 # my slot in entry 0 of peer p's ring
        self.peer_count = [p + self.nrow * 4 for p in self.peer_rows]
        self.peer_flag = [p + ring * 4 for p in ptrs]
        self.peer_ack = [p + (ring + 16) * 4 for p in ptrs]
        import ctypes as C

        self._ack_array = (C.c_void_p * self.world)(*self.peer_ack)
        dist.barrier(group)  # every rank has zeroed its buffer before anyone stores into it

    def bind(self, out) -> None:
        """Fill the peer fields of a ``_cabi.NmsOut``."""
        out.num_peers, out.my_rank, out.peer_depth = self.world, self.rank, self.depth
        for i in range(self.world):
            out.peer_rows[i], out.peer_count[i], out.peer_flag[i] = self.peer_rows[i], self.peer_count[i], self.peer_flag[i]
        out.peer_state = self.state.data_ptr()
        out.peer_ack = self.acks.data_ptr()
        out.peer_entry_stride = self.entry

    def wait(self, lag: int = 0) -> None:
        """Release the previously returned entry, then wait: lag=0 for the latest launch of every rank, lag=k for the launch k
        launches back (pipelined gather; k <= depth - 2)."""
        from . import _cabi

        rc = _cabi.load().ypb_peer_wait(self.flags.data_ptr(), self.world, self.state.data_ptr(), int(lag), self.depth,
                                        self._ack_array, self.rank, self.slot_index.data_ptr(), _cabi.stream_ptr(self.buf.device))
        _cabi.check(rc, "ypb_peer_wait")

    def wait_copy(self, out: torch.Tensor, lag: int = 0) -> None:
        """``wait(lag)`` and the copy of the returned entry into ``out`` (>= world*slot float32, 16-byte aligned) in ONE kernel."""
        from . import _cabi

        if out.numel() < self.entry or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous float32 tensor of >= {self.entry} elements")
        rc = _cabi.load().ypb_peer_wait_copy(self.flags.data_ptr(), self.world, self.state.data_ptr(), int(lag), self.depth,
                                             self._ack_array, self.rank, self.slot_index.data_ptr(), self.buf.data_ptr(),
                                             self.entry, out.data_ptr(), _cabi.stream_ptr(self.buf.device))
        _cabi.check(rc, "ypb_peer_wait_copy")

    def overrun(self) -> int:
        """0, or evidence of a broken protocol (reads the device state: synchronises): > 0 = a launch of this rank gave up
        waiting for a consumer's acknowledgement and overwrote a ring entry (some rank never calls ``wait``); < 0 = a ``wait``
        gave up waiting for a peer's launch."""
        return int(self.state[3].item())

    def ring(self) -> torch.Tensor:
        """(depth, world, slot) view of the local ring."""
        return self.buf[: self.depth * self.entry].view(self.depth, self.world, self.slot)

    def entry_tensor(self) -> torch.Tensor:
        """(world, packed) copy-free-on-host selection of the entry the last executed ``wait`` returned: one device-side
        ``index_select`` with ``slot_index`` (capturable in a CUDA graph; no host synchronisation)."""
        return self.ring().index_select(0, self.slot_index)[0][:, : self.numel]

    def copy_entry(self, out: torch.Tensor) -> torch.Tensor:
        """Copy the entry the last executed ``wait`` returned into ``out`` ((1, world, slot) float32) with ONE device-side gather
        (no temporary, no host synchronisation; capturable): the cheapest complete consumer."""
        return torch.index_select(self.ring(), 0, self.slot_index, out=out)

    def gathered(self, batch_per_rank: int, max_det: int, cols: int):
        """(world*B, max_det, cols) rows and (world*B,) int32 counts of the entry the last executed ``wait`` returned."""
        g = self.entry_tensor()
        rows = g[:, : self.nrow].reshape(self.world * batch_per_rank, max_det, cols)
        count = g[:, self.nrow:].contiguous().view(torch.int32).reshape(self.world * batch_per_rank)
        return rows, count
