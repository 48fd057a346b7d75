"""Single-box data parallelism for the Mixer-CLIP step, one process per GPU over torch.distributed
(NCCL on the GPUs, gloo in the CPU tests).  Replaces what ``accelerate`` does for the reference
(third-party, un-vendored; call sites training/training.py:64,93-95,158-159,170):

  * ``gather``: rank-ordered all_gather of the DETACHED normalised features along dim 0
    (training.py:158-159), both towers in one [n, 2E] message;
  * gradient averaging: DDP semantics (mean over ranks) on contiguous buckets of the flat gradient
    buffer, each launched from the backward schedule as soon as its block is complete
    (engine.TowerRT.backward -> after_block), on a side stream so it overlaps the remaining backward;
  * labels: ``arange(n) + rank * n`` (training.py:165-167) - passed to the head kernel as ``rank``.

The path shards by samples only (weights replicated): there is no other collective.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def bucket_index(tags: List[Tuple[str, object]]):
    """(kind, tag) -> bucket number, in backward completion order (see CLIP._flat_order)."""
    return {t: i for i, t in enumerate(tags)}


class GradBucketReducer:
    """Averages slices of one flat gradient tensor across ranks, bucket by bucket.

    Works on any device/backend (the CPU tests drive it with gloo); on CUDA the all-reduces run on
    ``comm_stream`` behind an event recorded on the compute stream at the call point.
    """

    def __init__(self, flat_g: torch.Tensor, ranges: List[Tuple[int, int]], group=None, min_bucket_elems: int = 0,
                 close_after=(), tail_bucket_elems: int = None):
        self.flat_g, self.group = flat_g, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cuda = flat_g.is_cuda
        self.comm_stream = torch.cuda.Stream() if self.cuda else None
        # merge consecutive ranges until each bucket holds at least min_bucket_elems (launch-latency bound).  Towards the
        # end of a segment (a segment ends at a close_after index: one tower's backward) the threshold drops to
        # tail_bucket_elems once no more than 1.5 full buckets remain: what is still unreduced when the backward pass ends
        # is exposed, so the last blocks travel in smaller messages that start earlier.
        if tail_bucket_elems is None or tail_bucket_elems > min_bucket_elems:
            tail_bucket_elems = min_bucket_elems
        remaining, acc = [0] * len(ranges), 0                 # elements from range i to the end of its segment
        for i in range(len(ranges) - 1, -1, -1):
            if i in close_after:
                acc = 0
            acc += ranges[i][1] - ranges[i][0]
            remaining[i] = acc
        merged, cur_b, cur_e, members, cur_m = [], None, None, [], []
        for i, (b, e) in enumerate(ranges):
            if cur_b is None:
                cur_b, cur_e, cur_m = b, e, [i]
                thr = tail_bucket_elems if 2 * remaining[i] <= 3 * min_bucket_elems else min_bucket_elems
            else:
                assert b == cur_e, "bucket ranges must be contiguous and in backward order"
                cur_e = e
                cur_m.append(i)
            if cur_e - cur_b >= thr or i in close_after:
                merged.append((cur_b, cur_e))
                members.append(cur_m)
                cur_b = None
        if cur_b is not None:
            merged.append((cur_b, cur_e))
            members.append(cur_m)
        self.buckets = merged
        self.last_member = {m[-1]: k for k, m in enumerate(members)}   # original index that completes bucket k
        self.pending = []
        self.launched = 0
        # MC_DP_TRACE=1: CUDA events around every bucket's all-reduce (comm stream) and at the points the compute stream
        # makes a bucket ready / joins the comm stream - bench.py turns them into the measured timeline of one step
        # (how much of the all-reduce is hidden behind the backward pass, how long the exposed tail is).
        import os
        self.trace = [] if (self.cuda and os.environ.get("MC_DP_TRACE", "0") == "1") else None
        # MC_DP_BF16=1 (opt-in): buckets travel as bf16 (half the NVLink bytes); the sum is formed in bf16 by NCCL, which
        # costs one rounding of each gradient element (2^-9 relative) - measured against the oracle by tools/dp_check.py.
        self.bf16_wire = self.cuda and os.environ.get("MC_DP_BF16", "0") == "1"
        self._wire = None

    def ready(self, original_index: int):
        """Original range ``original_index`` is complete; launch its (merged) bucket if it closes one."""
        k = self.last_member.get(original_index)
        if k is None or self.world == 1:
            return
        b, e = self.buckets[k]
        view = self.flat_g[b:e]
        if self.cuda:
            ev = torch.cuda.Event(enable_timing=self.trace is not None)
            ev.record()
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                t0 = t1 = None
                if self.trace is not None:
                    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    t0.record()
                if self.bf16_wire:
                    if self._wire is None:
                        self._wire = torch.empty(max(e_ - b_ for b_, e_ in self.buckets), device=view.device, dtype=torch.bfloat16)
                    w = self._wire[:e - b]
                    w.copy_(view)
                    dist.all_reduce(w, op=dist.ReduceOp.AVG, group=self.group)
                    view.copy_(w)
                else:
                    dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group)
                if self.trace is not None:
                    t1.record()
                    self.trace.append((k, (e - b) * 4, ev, t0, t1))
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
            view.div_(self.world)
        self.launched += 1

    def finish(self):
        """Make the compute stream wait for every launched all-reduce."""
        if self.cuda and self.world > 1:
            if self.trace is not None:
                self.join_begin = torch.cuda.Event(enable_timing=True)
                self.join_begin.record()
            torch.cuda.current_stream().wait_stream(self.comm_stream)
            if self.trace is not None:
                self.join_end = torch.cuda.Event(enable_timing=True)
                self.join_end.record()
        self.launched = 0

    def timeline(self, origin):
        """After a traced step has been synchronised: per-bucket (ready, start, end) in ms relative to the CUDA event
        `origin`, the time the compute stream reached the join, and how long it then waited (the exposed tail)."""
        if not self.trace:
            return None
        rows = [{"bucket": k, "MB": round(nbytes / 1e6, 1), "ready_ms": round(origin.elapsed_time(ev), 3),
                 "start_ms": round(origin.elapsed_time(t0), 3), "end_ms": round(origin.elapsed_time(t1), 3)}
                for k, nbytes, ev, t0, t1 in self.trace]
        out = {"buckets": rows, "join_reached_ms": round(origin.elapsed_time(self.join_begin), 3),
               "join_done_ms": round(origin.elapsed_time(self.join_end), 3),
               "exposed_tail_ms": round(self.join_begin.elapsed_time(self.join_end), 3),
               "allreduce_busy_ms": round(sum(r["end_ms"] - r["start_ms"] for r in rows), 3)}
        self.trace.clear()
        return out


def gather_features(ui: torch.Tensor, ut: torch.Tensor, group=None):
    """all_gather of [n, E] image and text features in ONE [n, 2E] message, rank order along dim 0,
    no autograd through it (training.py:158-159).  Returns (ui_all, ut_all) of shape [N, E]."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ui.detach(), ut.detach()
    world = dist.get_world_size(group)
    n, E = ui.shape
    packed = torch.cat([ui.detach(), ut.detach()], dim=1).contiguous()
    out = torch.empty(world * n, 2 * E, device=ui.device, dtype=ui.dtype)
    dist.all_gather_into_tensor(out, packed, group=group)
    return out[:, :E].contiguous(), out[:, E:].contiguous()


class DataParallel:
    """Installs the bucketed gradient averaging on a CLIP model (``model._dp``) and exposes the
    feature gather.  ``isinstance(x, DataParallel)`` plays the role of the reference's
    ``isinstance(self.model, DistributedDataParallel)`` branch (training.py:174)."""

    def __init__(self, model, group=None, min_bucket_mb: float = None):
        import os
        if min_bucket_mb is None:
            # smallest all-reduce message (blocks are merged up to it).  Measured on 8 B200s (profiles/r2d_bench_8gpu_bucket*):
            # 8 MB = 27 latency-bound messages, 4.2 ms of all-reduce kernels competing with the persistent GEMMs, 16.31 ms per
            # step; 64-256 MB = 3-7 messages, 1.5-2.2 ms, 15.75-15.84 ms; 64 MB keeps the exposed tail smallest (0.68 ms)
            min_bucket_mb = float(os.environ.get("MC_DP_BUCKET_MB", "64"))
        # MC_DP_TAIL_MB: bucket size for the last 1.5 buckets' worth of each tower's backward (GradBucketReducer)
        tail_mb = float(os.environ.get("MC_DP_TAIL_MB", str(min_bucket_mb)))
        self.module, self.group = model, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        store = model._require_store()
        store.ensure_grads()
        self.index = bucket_index(model._bucket_tags)
        # never merge across a tower boundary: autograd may run the two tower backwards in either order
        close = {self.index[("text", "bottom")], self.index[("image", "bottom")]}
        self.reducer = GradBucketReducer(store.flat_g, store.bucket_ranges, group,
                                         min_bucket_elems=int(min_bucket_mb * (1 << 20) / 4), close_after=close,
                                         tail_bucket_elems=int(tail_mb * (1 << 20) / 4))
        model._dp = self

    def after_block_hook(self, kind: str) -> Optional[Callable]:
        if self.world == 1:
            return None
        return lambda tag: self.reducer.ready(self.index[(kind, tag)])

    def gather(self, ui, ut):
        return gather_features(ui, ut, self.group)

    def finish(self):
        """Call once per step after backward: reduces the tail bucket (logit_scale) and joins the streams."""
        if self.world > 1:
            self.reducer.ready(self.index[("head", "final")])
        self.reducer.finish()

    # nn.Module-like conveniences so the wrapper can stand where the DDP-wrapped model stood
    def __call__(self, *a, **k):
        return self.module(*a, **k)

    def parameters(self):
        return self.module.parameters()

    def named_parameters(self):
        return self.module.named_parameters()

    def train(self, mode=True):
        self.module.train(mode)
        return self

    def eval(self):
        self.module.eval()
        return self
