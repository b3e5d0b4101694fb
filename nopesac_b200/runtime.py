"""Host-side runtime of the path: feeds the camera head from HOST (pinned) buffers with the H2D copy of batch i+1
overlapped with the kernels of batch i (two device slots, one copy stream, CUDA events — no host synchronisation
until a result is read).  The head itself is called exactly as a user would call it.

    pipe = PairPipeline(head, matching_head, device, hyp_pairs=...)
    for pose_rows in pipe.run(host_batches):      # pose_rows: pinned [B,16] fp32 (t, q, t_avg, q_avg, m, 0)
        ...
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, Optional

import torch

FEATURE_KEYS = ("res2", "res3", "res4", "res5")


def pin_batch(batch: Dict) -> Dict:
    """Page-lock every tensor of a host batch {planes1, planes2, app1, app2, feats1{...}, feats2{...}}."""
    out = {}
    for k, v in batch.items():
        out[k] = {kk: vv.contiguous().pin_memory() for kk, vv in v.items()} if isinstance(v, dict) else v.contiguous().pin_memory()
    return out


def batch_bytes(batch: Dict) -> int:
    n = 0
    for v in batch.values():
        n += sum(t.numel() * t.element_size() for t in v.values()) if isinstance(v, dict) else v.numel() * v.element_size()
    return n


class PairPipeline:
    """`compute` (optional): callable(device_batch) -> result rows [B,16]; default = the camera head on backbone feature maps
    (batch keys planes1, planes2, app1, app2, feats1, feats2).  The full model from RGB passes e.g.
    `lambda d: model.inference_from_images(d["images"], None, d["planes1"], ...)[5]["pose"]` with a uint8 `images` entry."""

    def __init__(self, head, matching_head, device, hyp_pairs: Optional[torch.Tensor] = None, post=None, result_exchange=None,
                 compute=None):
        self.head, self.match, self.device, self.hyp_pairs, self.post = head, matching_head, device, hyp_pairs, post
        self.result_exchange = result_exchange
        self.compute = compute
        self.copy_stream = torch.cuda.Stream(device)
        self.slots = [None, None]
        self.copied = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.host_out = [None, None]
        self.done = [torch.cuda.Event(), torch.cuda.Event()]

    def _alloc_like(self, batch):
        mk = lambda t: torch.empty(t.shape, dtype=t.dtype, device=self.device)
        return {k: ({kk: mk(vv) for kk, vv in v.items()} if isinstance(v, dict) else mk(v)) for k, v in batch.items()}

    def _enqueue_copy(self, slot: int, batch: Dict):
        if self.slots[slot] is None:
            self.slots[slot] = self._alloc_like(batch)
        dst = self.slots[slot]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])          # kernels that read this slot have finished
            for k, v in batch.items():
                if isinstance(v, dict):
                    for kk, vv in v.items():
                        dst[k][kk].copy_(vv, non_blocking=True)
                else:
                    dst[k].copy_(v, non_blocking=True)
            self.copied[slot].record(self.copy_stream)

    def _enqueue_compute(self, slot: int):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.copied[slot])
        d = self.slots[slot]
        if self.compute is not None:
            rows = self.compute(d)
        else:
            out = self.head(d["feats1"], d["feats2"], d["planes1"], d["planes2"], d["app1"], d["app2"],
                            matching_net=self.match, hyp_pairs=self.hyp_pairs, result_exchange=self.result_exchange)
            rows = out[5]["pose"]
        if self.post is not None:
            rows = self.post(rows)                                     # e.g. the multi-GPU result all-gather
        self.consumed[slot].record(cur)
        if self.host_out[slot] is None or self.host_out[slot].shape != rows.shape:
            self.host_out[slot] = torch.empty(rows.shape, dtype=rows.dtype).pin_memory()
        self.host_out[slot].copy_(rows, non_blocking=True)
        self.done[slot].record(cur)

    def run(self, host_batches: Iterable[Dict]) -> Iterator[torch.Tensor]:
        """Yields the pinned [B,16] result rows of every batch, in order.

        The yielded tensor is one of TWO reused pinned buffers: it is overwritten two steps later.  Consume it inside the
        loop body, or `.clone()` it — `list(pipe.run(...))` keeps references to buffers that are rewritten."""
        it = iter(host_batches)
        try:
            nxt = next(it)
        except StopIteration:
            return
        self.consumed[0].record(torch.cuda.current_stream(self.device))
        self.consumed[1].record(torch.cuda.current_stream(self.device))
        self._enqueue_copy(0, nxt)
        i = 0
        pending = None
        while nxt is not None:
            slot = i & 1
            try:
                following = next(it)
            except StopIteration:
                following = None
            self._enqueue_compute(slot)
            if following is not None:
                self._enqueue_copy(slot ^ 1, following)               # overlaps the kernels just enqueued
            if pending is not None:
                self.done[pending].synchronize()
                yield self.host_out[pending]
            pending, nxt, i = slot, following, i + 1
        if pending is not None:
            self.done[pending].synchronize()
            yield self.host_out[pending]


class GraphedCameraHead:
    """The camera-head forward captured once in a CUDA graph and replayed (CUDA graphs instead of a tracing compiler).

    The forward is a fixed sequence of ~380 short launches for a given batch shape (no host synchronisation, no
    data-dependent shapes), so the whole step can be replayed with one graph launch: the matcher's tiny kernels are
    host-enqueue-bound when launched eagerly (profiles/r1h_step_stages.json: 15.8 ms eager vs 15.3 ms replayed).

        runner = GraphedCameraHead(head, matching_head, batch, hyp_pairs=hp)   # batch: device tensors (static buffers)
        rows = runner()                   # replays; result rows [B,16] live in a static buffer
        runner.load(new_batch)            # device-to-device copy into the static input buffers, then runner()

    The inputs are static buffers owned by the runner (the tensors given at construction are used as those buffers)."""

    def __init__(self, head, matching_head, batch: Dict, hyp_pairs: Optional[torch.Tensor] = None, warmup: int = 2):
        self.batch = batch
        self._call = lambda: head(batch["feats1"], batch["feats2"], batch["planes1"], batch["planes2"], batch["app1"], batch["app2"],
                                  matching_net=matching_head, hyp_pairs=hyp_pairs)[5]["pose"]
        dev = batch["planes1"].device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):          # warm-up off the capture: weight packs, kernel attributes, allocator pools
            for _ in range(max(1, warmup)):
                self._call()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.rows = self._call()

    def load(self, batch: Dict):
        """Copies a new batch (same shapes) into the static input buffers on the current stream."""
        for k, v in batch.items():
            if isinstance(v, dict):
                for kk, vv in v.items():
                    self.batch[k][kk].copy_(vv, non_blocking=True)
            else:
                self.batch[k].copy_(v, non_blocking=True)

    def __call__(self) -> torch.Tensor:
        self.graph.replay()
        return self.rows
