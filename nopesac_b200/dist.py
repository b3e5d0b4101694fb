"""Multi-GPU plumbing of the path: image pairs are independent, so they are block-sharded over the ranks
(rank r owns pairs [r*B/R, (r+1)*B/R), the reference's own data parallelism: test_NopeSAC.py:209-216 +
detectron2 InferenceSampler) and the ONLY exchange is one all-gather of the [B_local,16] fp32 per-pair result
rows (pose 7 + avg pose 7 + matched_num + pad = 64 B/pair), replacing the pickled-object `comm.gather` of
evaluation/mp3d_evaluation.py:317-318.  One process per GPU, `torch.distributed` (NCCL on GPUs, gloo in the
CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist

RESULT_WIDTH = 16


def shard_range(num_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block split; the first `num_pairs % world` ranks take one extra pair."""
    base, rem = divmod(num_pairs, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(num_pairs: int, world: int):
    return [shard_range(num_pairs, r, world)[1] - shard_range(num_pairs, r, world)[0] for r in range(world)]


def gather_results(local_rows: torch.Tensor, num_pairs: int, group=None) -> torch.Tensor:
    """All-gather of the per-pair result rows -> [num_pairs, 16] on every rank, in global pair order."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_rows
    world = dist.get_world_size(group)
    sizes = shard_sizes(num_pairs, world)
    assert local_rows.shape == (sizes[dist.get_rank(group)], RESULT_WIDTH), (local_rows.shape, sizes)
    if len(set(sizes)) == 1:
        out = torch.empty(num_pairs, RESULT_WIDTH, dtype=local_rows.dtype, device=local_rows.device)
        dist.all_gather_into_tensor(out, local_rows.contiguous(), group=group)
        return out
    # ragged tail: pad every shard to the largest one, gather once, strip the padding
    mx = max(sizes)
    padded = torch.zeros(mx, RESULT_WIDTH, dtype=local_rows.dtype, device=local_rows.device)
    padded[: local_rows.shape[0]] = local_rows
    out = torch.empty(world * mx, RESULT_WIDTH, dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * mx: r * mx + sizes[r]] for r in range(world)], 0)


class FusedResultExchange:
    """Result exchange fused into the last kernel of the path: every rank owns a [2, world*B_local, 16] buffer in
    symmetric (NVLink peer-mapped) memory; the selection kernel that finishes a pair stores its 64-byte row into the
    corresponding row of EVERY rank's buffer (plain st.global on peer mappings), so no all-gather collective runs —
    only one cross-rank barrier (`finish()`) before the rows are read.

    Double-buffered by step parity: step i writes slot i & 1.  A fast rank's stores of step i+1 therefore go to the OTHER
    slot while a slow rank still reads step i; slot i & 1 is only written again in step i+2, which a rank can start only
    after the barrier of step i+1 — and every rank enqueues that barrier after its own read-out of step i (same stream).
    So the rows returned by `finish()` stay valid until the `finish()` after next, by protocol, not by timing.

    Needs one process per GPU with an initialised NCCL process group and NVLink / P2P access between the GPUs."""

    def __init__(self, pairs_per_rank: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.pairs_per_rank = pairs_per_rank
        self.slot_rows = self.world * pairs_per_rank
        self.step = 0
        self.buf = symm_mem.empty((2 * self.slot_rows, RESULT_WIDTH), dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.peer_ptrs_dev = int(self.handle.buffer_ptrs_dev)     # device array of `world` float* (one per rank)
        torch.cuda.synchronize(device)
        dist.barrier(group)

    @property
    def row_offset(self) -> int:
        """First row (in every rank's buffer) of this rank's shard in the slot of the current step."""
        return (self.step & 1) * self.slot_rows + self.rank * self.pairs_per_rank

    @property
    def rows(self) -> torch.Tensor:
        """The current step's slot: this rank's complete [world*B_local, 16] copy once `finish()` has run."""
        s = (self.step & 1) * self.slot_rows
        return self.buf[s:s + self.slot_rows]

    def finish(self) -> torch.Tensor:
        """Cross-rank barrier on the current stream: afterwards the returned slot holds the rows of all ranks (valid until
        the `finish()` after next).  Advances to the other slot."""
        self.handle.barrier(channel=0)
        out = self.rows
        self.step += 1
        return out
