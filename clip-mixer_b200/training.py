"""Training step and loop of the reference's ``training/training.py`` on the CUDA path.

``FusedTrainStep`` is lines :144-186 of the reference loop (zero_grad -> forward -> gather(detach) ->
logits / cross-entropy -> backward -> logit-scale clamp -> clip_grad_norm_ -> AdamW -> scheduler) as one
explicit schedule over libmixerclip kernels, with no autograd graph, no host synchronisation (the
reference's ``total_loss.item()`` at :190 is a per-step sync) and, optionally, captured once into a
CUDA graph and replayed.

``Trainer`` keeps the shape of the reference class (``Trainer(model, preprocess, epochs, args)``,
``train()``, ``validate()``, ``save_model()``, ``load_model()``) on synthetic data: the Azure / LAION
/ accelerate / TensorBoard scaffolding around the loop is out of scope (SURVEY 2.1 rows 8-10,12).
"""
from __future__ import annotations

import json
import math
import os
from typing import Optional

import torch
import torch.distributed as dist

from . import checkpoint, ops
from ._lib import MixerClipError
from .dp import DataParallel
from .optim import FusedAdamW, cosine_warmup_lr


def parse_sm_split(value: str):
    """MC_SM_SPLIT -> (fixed shares or None, tune at capture?): 'auto' | 'off' | 'image_sms,text_sms'."""
    v = value.strip().lower()
    if v == "auto":
        return None, True
    if v in ("", "off", "0", "none"):
        return None, False
    try:
        shares = tuple(int(x) for x in v.split(","))
    except ValueError:
        shares = ()
    if len(shares) != 2 or min(shares) < 2:
        raise MixerClipError(f"MC_SM_SPLIT={value!r}: expected 'image_sms,text_sms', 'auto' or 'off'")
    return shares, False


class FusedTrainStep:
    """One optimisation step on a per-rank batch.  ``step(images, texts)`` returns the (device) loss of
    this rank; nothing is read back to the host."""

    def __init__(self, model, optimizer: Optional[FusedAdamW] = None, dp: Optional[DataParallel] = None,
                 total_steps: int = 10 ** 9, max_lr: float = 5e-4, warmup_steps: int = 2, use_cuda_graph: bool = False,
                 overlap_towers: bool = True, micro_batch: Optional[int] = None):
        self.model = model
        self.dp = dp
        self.world = dp.world if dp is not None else 1
        self.rank = dp.rank if dp is not None else 0
        store = model._require_store()
        store.ensure_grads()
        model._attach_grads()
        self.store = store
        self.opt = optimizer if optimizer is not None else FusedAdamW(model, lr=max_lr)
        self.total_steps, self.max_lr, self.min_lr, self.warmup = total_steps, max_lr, max_lr / 100, warmup_steps
        self._sched_step = 0
        self.opt.set_schedule(total_steps, max_lr, max_lr / 100, warmup_steps)
        dev = store.device
        self.loss = torch.zeros(1, device=dev)
        self.use_graph = use_cuda_graph
        self.graph = None
        self.static_images = self.static_texts = None
        self._bufs = {}
        self.clamp_ddp_branch = self.world > 1     # training.py:174-178 has two different clamps
        self._side = None
        self.overlap_towers = overlap_towers
        self.micro_batch = micro_batch
        # SM split: while the two towers run on two streams, each tower's persistent kernels are sized for its own
        # share of the SMs, so a tensor-bound GEMM of one tower runs beside an HBM-bound row kernel of the other
        # instead of each taking turns on the whole GPU (every kernel of this step holds a full SM per CTA).
        # MC_SM_SPLIT = "image_sms,text_sms" fixes the shares, "off" sizes every kernel for all SMs, "auto" (default,
        # CUDA-graph mode only) times a few shares around the towers' FLOP ratio at capture and keeps the fastest.
        self.sm_split, self.sm_split_auto = parse_sm_split(os.environ.get("MC_SM_SPLIT", "auto"))
        self.sm_split_trials = []

    def _tower_sms(self, which: int):
        if self.sm_split is not None and self.overlap_towers:
            ops.set_sm_limit(self.sm_split[which])

    def _side_stream(self):
        if not self.overlap_towers:
            return torch.cuda.current_stream()      # single-stream schedule (per-kernel timing passes)
        if self._side is None:
            self._side = torch.cuda.Stream()
        return self._side

    # ---- schedule -----------------------------------------------------------------------------------
    @property
    def sched_step(self):
        return self._sched_step

    @sched_step.setter
    def sched_step(self, value):
        """Host view of the scheduler step; assigning it (checkpoint resume) also positions the device counter."""
        self._sched_step = int(value)
        self.opt.sync_device_counters(sched_step=self._sched_step)

    def current_lr(self):
        return cosine_warmup_lr(self.sched_step, self.total_steps, self.max_lr, self.min_lr, self.warmup)

    # ---- the device-side schedule (graph-capturable) ----------------------------------------------
    def _device_step(self, images, texts):
        model, store = self.model, self.store
        prec = model._precision
        img_t, txt_t = model._towers["image"], model._towers["text"]
        n = images.shape[0]
        E = model._cfg["embed_dim"]
        store.flat_g.zero_()                                                        # optimizer.zero_grad()  :144
        model._prepare_weights()
        if self.micro_batch is not None and n > self.micro_batch:
            self._micro_batched_backward(images, texts)
            self._finish_step()
            return
        # The two towers are independent until the loss: they run on two streams so that one tower's kernels fill
        # the SMs the other leaves idle (partial last waves of the persistent GEMMs, HBM-bound row kernels).
        main = torch.cuda.current_stream()
        side = self._side_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        self._tower_sms(1)
        with torch.cuda.stream(side):
            ws_t = txt_t.forward(texts, prec, True)                                 # model(images, texts)   :156
        self._tower_sms(0)
        ws_i = img_t.forward(images, prec, True)
        ops.set_sm_limit(0)
        main.wait_stream(side)
        if self.dp is not None and self.world > 1:
            ui_all, ut_all = self.dp.gather(ws_i.u_feat, ws_t.u_feat)               # accelerator.gather     :158-159
        else:
            ui_all, ut_all = ws_i.u_feat, ws_t.u_feat
        N = ui_all.shape[0]
        key = (n, N, E)
        if key not in self._bufs:
            dev = store.device
            self._bufs = {key: dict(dui=torch.empty(n, E, device=dev), dut=torch.empty(n, E, device=dev),
                                    ws=torch.empty(ops.head_workspace_bytes(n, N, E) // 4, device=dev))}
        b = self._bufs[key]
        self.loss.zero_()
        ops.head_fwd_bwd(ws_i.u_feat, ws_t.u_feat, ui_all, ut_all, model.logit_scale, n, N, E, self.rank, 1.0,
                         self.loss, b["dui"], b["dut"], store.grad_view("logit_scale"), b["ws"])   # :162-170
        hook_t = self.dp.after_block_hook("text") if self.dp is not None else None
        hook_i = self.dp.after_block_hook("image") if self.dp is not None else None
        fork2 = torch.cuda.Event()
        fork2.record(main)
        side.wait_event(fork2)
        # The two backward passes are ENQUEUED block by block in alternation (text on the side stream, image on the main
        # stream): the kernels still run concurrently, and the gradient buckets reach the single in-order communication
        # stream in the order in which they become ready (see TowerRT.backward_iter).
        gen_t = txt_t.backward_iter(ws_t, b["dut"], prec, after_block=hook_t)     # accelerator.backward   :170
        gen_i = img_t.backward_iter(ws_i, b["dui"], prec, after_block=hook_i)
        live_t = live_i = True
        while live_t or live_i:
            if live_t:
                self._tower_sms(1)
                with torch.cuda.stream(side):
                    live_t = next(gen_t, None) is not None
            if live_i:
                self._tower_sms(0)
                live_i = next(gen_i, None) is not None
        ops.set_sm_limit(0)
        main.wait_stream(side)
        self._finish_step()

    def _finish_step(self):
        model = self.model
        if self.dp is not None:
            self.dp.finish()
        with torch.no_grad():                                                       # clamp                  :173-178
            if self.clamp_ddp_branch:
                model.logit_scale.data.clamp_(0, math.log(100))
            else:
                model.logit_scale.data.clamp_(max=100)
        self.opt.launch_scalars()                                                   # lr(s), bias corrections: on the device
        self.opt.launch(grad_mul=1.0)                                               # clip + step            :181,185

    def _micro_batched_backward(self, images, texts):
        """Global-batch contrastive step when n samples do not fit at once (config 3: 32768 over 2-4 GPUs,
        SURVEY 7.3-7).  Because the gathered features are detached (training.py:158-159) this is EXACT, not an
        approximation: pass 1 computes every local feature without saving activations, the features are gathered,
        pass 2 re-runs each micro-batch with activations and back-propagates its rows against the cached global
        features (labels offset by rank*n + k*m)."""
        model, store = self.model, self.store
        prec = model._precision
        img_t, txt_t = model._towers["image"], model._towers["text"]
        n, m = images.shape[0], self.micro_batch
        if n % m:
            raise MixerClipError(f"per-rank batch {n} must be a multiple of micro_batch {m}")
        K, E, dev = n // m, model._cfg["embed_dim"], store.device
        key = ("mb", n, m)
        if key not in self._bufs:
            self._bufs[key] = dict(ui=torch.empty(n, E, device=dev), ut=torch.empty(n, E, device=dev),
                                   dui=torch.empty(m, E, device=dev), dut=torch.empty(m, E, device=dev),
                                   loss=torch.zeros(1, device=dev))
        b = self._bufs[key]
        for k in range(K):                                                          # pass 1: features only
            sl = slice(k * m, (k + 1) * m)
            b["ui"][sl].copy_(img_t.forward(images[sl], prec, False).u_feat)
            b["ut"][sl].copy_(txt_t.forward(texts[sl], prec, False).u_feat)
        if self.dp is not None and self.world > 1:
            ui_all, ut_all = self.dp.gather(b["ui"], b["ut"])
        else:
            ui_all, ut_all = b["ui"], b["ut"]
        N = ui_all.shape[0]
        if "ws" not in b or b["ws_key"] != (m, N):
            b["ws"] = torch.empty(ops.head_workspace_bytes(m, N, E) // 4, device=dev)
            b["ws_key"] = (m, N)
        self.loss.zero_()
        for k in range(K):                                                          # pass 2: activations + backward
            sl = slice(k * m, (k + 1) * m)
            last = k == K - 1
            ws_i = img_t.forward(images[sl], prec, True)
            ws_t = txt_t.forward(texts[sl], prec, True)
            b["loss"].zero_()
            ops.head_fwd_bwd(ws_i.u_feat, ws_t.u_feat, ui_all, ut_all, model.logit_scale, m, N, E, self.rank * K + k,
                             m / n, b["loss"], b["dui"], b["dut"], store.grad_view("logit_scale"), b["ws"])
            self.loss.add_(b["loss"], alpha=m / n)
            hook_t = self.dp.after_block_hook("text") if (self.dp is not None and last) else None
            hook_i = self.dp.after_block_hook("image") if (self.dp is not None and last) else None
            txt_t.backward(ws_t, b["dut"], prec, after_block=hook_t)
            img_t.backward(ws_i, b["dui"], prec, after_block=hook_i)

    def step(self, images: torch.Tensor, texts: torch.Tensor) -> torch.Tensor:
        if not (images.is_cuda and texts.is_cuda):
            raise MixerClipError("FusedTrainStep needs CUDA inputs (no CPU fallback)")
        if not self.use_graph:
            self._device_step(images, texts)
        else:
            if self.graph is None:
                self._capture(images, texts)
            self.static_images.copy_(images, non_blocking=True)
            self.static_texts.copy_(texts, non_blocking=True)
            self.graph.replay()
        self.model._trusted_mirror = True     # the optimizer kernel keeps the bf16 mirror current
        self.opt.t += 1                       # host mirrors of the device counters (advanced by mc_sched_step in the step)
        self._sched_step += 1                                                       # scheduler.step()       :186
        return self.loss

    def _capture(self, images, texts):
        self.static_images = images.clone()
        self.static_texts = texts.to(torch.int64).clone()
        # warm-up on a side stream (allocations, tensor maps, NCCL channels), restoring the weights after
        snap_p = self.store.flat_p.clone()
        snap_m, snap_v = self.opt.m.clone(), self.opt.v.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._device_step(self.static_images, self.static_texts)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.store.flat_p.copy_(snap_p)
        self.opt.m.copy_(snap_m)
        self.opt.v.copy_(snap_v)
        self.opt.sync_device_counters(sched_step=self._sched_step)   # the warm-up steps advanced the device counters
        if self.model._precision.act == torch.bfloat16:
            self.store.refresh_mirror(force=True)
        self.model._trusted_mirror = True     # do not bake a redundant cast pass into the graph
        candidates = self._split_candidates(images.shape[0])
        graphs = self._capture_candidates(candidates)
        pick = 0
        if len(candidates) > 1:
            # Interleaved timing (blocks of replays alternate over the candidates, so clock / power drift hits all of
            # them alike: back-to-back blocks of one candidate differ by more than the candidates do).  The replays are
            # real optimizer steps, so the state is put back afterwards.
            times = self._time_interleaved(graphs)
            pick = min(range(len(candidates)), key=lambda i: times[i])
            self.sm_split_trials = list(zip(candidates, times))
            if self.world > 1 and candidates[pick] is not None:
                # Data parallel: NCCL's all-reduce kernels need SMs too, and a persistent kernel whose CTAs do not all
                # fit at once runs in two waves.  Second round: the winning shares with 4 / 8 SMs left unassigned
                # (NCCL keeps its own CTA count - capping it exposes the all-reduce tail, profiles/r1s3_gemm_experiments.txt).
                a, b = candidates[pick]
                extra = [(a - r // 2, b - r // 2) for r in (4, 8) if min(a, b) - r // 2 >= 8]
                g2 = [graphs[pick]] + self._capture_candidates(extra)
                t2 = self._time_interleaved(g2)
                self.sm_split_trials += list(zip(extra, t2[1:]))
                best2 = min(range(len(g2)), key=lambda i: t2[i])
                if best2 > 0:
                    candidates, graphs, pick = candidates + extra, graphs + g2[1:], len(candidates) + best2 - 1
            self.store.flat_p.copy_(snap_p)
            self.opt.m.copy_(snap_m)
            self.opt.v.copy_(snap_v)
            self.opt.sync_device_counters(sched_step=self._sched_step)
            if self.model._precision.act == torch.bfloat16:
                self.store.refresh_mirror(force=True)
            self.model._trusted_mirror = True
        self.sm_split, self.graph = candidates[pick], graphs[pick]

    def _capture_candidates(self, candidates):
        graphs = []
        for cand in candidates:
            self.sm_split = cand
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._device_step(self.static_images, self.static_texts)
            graphs.append(graph)
        return graphs

    def _time_interleaved(self, graphs, rounds: int = 3):
        total = [0.0] * len(graphs)
        for _ in range(rounds):
            for i, graph in enumerate(graphs):
                total[i] += self._time_graph(graph)
        return [t / rounds for t in total]

    def _split_candidates(self, n):
        """SM shares (image, text) to try at capture; [None] = no split."""
        if not (self.sm_split_auto and self.overlap_towers) or (self.micro_batch is not None and n > self.micro_batch):
            return [self.sm_split]
        sms = ops.device_info()[0]
        work = []
        for t in (self.model._towers["image"], self.model._towers["text"]):
            work.append(t.L * (16.0 * t.P * t.D * t.D + 16.0 * t.P * t.P * t.D))   # mixer GEMM FLOPs per sample
        centre = int(round(sms * work[0] / (work[0] + work[1]) / 2.0)) * 2
        cands = [None]
        for d in (-6, -4, -2, 0):       # measured optimum sits a little below the FLOP share (B/32: 82-84 of 148 vs 86)
            a = centre + d
            if 8 <= a <= sms - 8:
                cands.append((a, sms - a))
        return cands

    def _time_graph(self, graph, reps: int = 6) -> float:
        """Milliseconds per replay (max over ranks, so every rank keeps the same candidate)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if self.world > 1:
            t = torch.tensor([ms], device=self.store.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms


# -------------------------------------------------------------------------------------------------
# synthetic data and the reference-shaped Trainer
# -------------------------------------------------------------------------------------------------
def synthetic_batch(cfg: dict, batch: int, seed: int, device, uint8_images: bool = True):
    """Throughput inputs of SURVEY 8-d: uint8 images (the loop's /255 + Normalize is fused into the patch
    embedding), token rows SOT .. EOT 0 0 0 with a unique arg-max."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    R = cfg["image_resolution"]
    if uint8_images:
        images = torch.randint(0, 256, (batch, 3, R, R), dtype=torch.uint8, generator=g)
    else:
        images = torch.randn(batch, 3, R, R, generator=g)
    C, V = cfg["context_length"], cfg["vocab_size"]
    text = torch.randint(1, max(2, V - 2), (batch, C), generator=g)
    text[:, 0] = V - 2
    pos = torch.randint(1, C, (batch,), generator=g)
    ar = torch.arange(C)[None, :]
    text = torch.where(ar == pos[:, None], torch.full_like(text, V - 1), text)
    text = torch.where(ar > pos[:, None], torch.zeros_like(text), text)
    return images.to(device), text.to(device)


class InputPipeline:
    """Input stage of the loop: the reference's ``DataLoader(..., pin_memory=True)`` +
    ``images.to(device)`` / ``clip.tokenize(texts).to(device)`` (training/training.py:60-62,149,154) as a
    double-buffered prefetch: batch k+1 travels host -> pinned slot -> device slot on a COPY stream while step k
    computes; the compute stream only waits for the event of the batch it is about to consume.  Images stay uint8
    on the wire (38.5 MB per 256 samples instead of 154 MB fp32); /255 + Normalize + the bf16 cast happen inside the
    patch-embedding operand producer (mc_im2col)."""

    def __init__(self, device, depth: int = 2):
        self.device, self.depth = device, depth
        self.copy_stream = torch.cuda.Stream(device=device)
        self.slots = [None] * depth          # (pinned images, pinned texts, device images, device texts)
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [None] * depth           # recorded on the compute stream when the slot's batch has been consumed
        self.head = self.tail = 0            # next slot to fill / next slot to hand out
        self.h2d_bytes = 0

    def _slot(self, i, images, texts):
        s = self.slots[i]
        if s is None or s[0].shape != images.shape or s[0].dtype != images.dtype or s[1].shape != texts.shape:
            s = (torch.empty(images.shape, dtype=images.dtype).pin_memory(),
                 torch.empty(texts.shape, dtype=texts.dtype).pin_memory(),
                 torch.empty(images.shape, dtype=images.dtype, device=self.device),
                 torch.empty(texts.shape, dtype=texts.dtype, device=self.device))
            self.slots[i] = s
        return s

    def prefetch(self, images: torch.Tensor, texts: torch.Tensor):
        """Queue one HOST batch.  Pinned tensors are copied from directly; pageable ones go through the slot's pinned
        staging buffer (a host memcpy, like the DataLoader's pin-memory thread)."""
        if self.head - self.tail >= self.depth:
            raise MixerClipError("InputPipeline: all slots are in flight (call next() first)")
        i = self.head % self.depth
        pi, pt, di, dt = self._slot(i, images, texts)
        if self.free[i] is not None:
            if not (images.is_pinned() and texts.is_pinned()):
                self.ready[i].synchronize()          # the pinned staging buffer is about to be overwritten by the host
            self.copy_stream.wait_event(self.free[i])
        src_i, src_t = images, texts
        if not images.is_pinned():
            pi.copy_(images)
            src_i = pi
        if not texts.is_pinned():
            pt.copy_(texts)
            src_t = pt
        with torch.cuda.stream(self.copy_stream):
            di.copy_(src_i, non_blocking=True)
            dt.copy_(src_t, non_blocking=True)
            self.ready[i].record(self.copy_stream)
        self.h2d_bytes = images.numel() * images.element_size() + texts.numel() * texts.element_size()
        self.head += 1

    def next(self):
        """Device tensors of the oldest queued batch; the current stream waits for its copy.  Call release() once the
        work that reads them has been enqueued."""
        if self.tail >= self.head:
            raise MixerClipError("InputPipeline: nothing queued")
        i = self.tail % self.depth
        torch.cuda.current_stream().wait_event(self.ready[i])
        return self.slots[i][2], self.slots[i][3]

    def release(self):
        i = self.tail % self.depth
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.free[i] = ev
        self.tail += 1


class SyntheticPairs:
    """Stands where ``LaionCoco`` + DataLoader stood (training.py:60-62): yields (images uint8, token ids)."""

    def __init__(self, cfg, batch, steps, device, seed=1000):
        # device="cpu" yields HOST batches (what a DataLoader hands the loop); Trainer feeds those through InputPipeline
        self.cfg, self.batch, self.steps, self.device, self.seed = cfg, batch, steps, device, seed

    def __len__(self):
        return self.steps

    def __iter__(self):
        for i in range(self.steps):
            yield synthetic_batch(self.cfg, self.batch, self.seed + i, self.device)


class Trainer:
    """Shape of the reference's Trainer (training.py:30-250) over FusedTrainStep and synthetic data."""

    def __init__(self, model, preprocess=None, epochs: int = 1, args=None, batch_size: int = 256,
                 steps_per_epoch: int = 10, use_cuda_graph: bool = False, validators=None):
        self.epochs = epochs
        # the reference builds ImageNet / STS / MNIST / SST-2 validators here (training.py:98-104); their datasets are out of
        # scope, so the caller passes objects with the same ``validate(step, verbose)`` method (zeroshot.ZeroShotValidator)
        self.validators = list(validators) if validators is not None else []
        self.validation_results = []
        self.model = model
        self.preprocess = preprocess
        self.runName = getattr(args, "run_name", "run") if args is not None else "run"
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        if batch_size % self.world:
            raise MixerClipError("global batch must divide by the number of ranks (split_batches, training.py:64)")
        dev = model.logit_scale.device
        self.trainLoader = SyntheticPairs(model._cfg, batch_size // self.world, steps_per_epoch, "cpu",
                                          seed=1000 + self.rank * 100003)                      # host batches, training.py:60-62
        self.inputs = InputPipeline(dev)
        self.numBatches = len(self.trainLoader)
        self.dp = DataParallel(model) if self.world > 1 else None
        self.optimizer = FusedAdamW(model, lr=5e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2, max_grad_norm=20.0)
        self.stepper = FusedTrainStep(model, self.optimizer, self.dp, total_steps=self.epochs * self.numBatches,
                                      max_lr=5e-4, warmup_steps=2, use_cuda_graph=use_cuda_graph)
        self.startEpoch, self.currentStep = self.load_model()
        self.losses = []

    def train(self):
        global_step = 0
        for epoch in range(self.startEpoch, self.epochs):
            self.model.train()
            it = iter(self.trainLoader)
            idx = -1

            def queue_next():
                nonlocal idx
                for images, texts in it:
                    idx += 1
                    if idx < self.currentStep:
                        continue                              # resume: skip the batches already consumed (training.py:146-147)
                    self.inputs.prefetch(images, texts)
                    return idx
                return None

            cur = queue_next()
            while cur is not None:
                images, texts = self.inputs.next()
                nxt = queue_next()                            # batch k+1 crosses PCIe while step k computes
                global_step = epoch * self.numBatches + cur
                loss = self.stepper.step(images, texts)
                self.inputs.release()
                self.losses.append(loss.clone())          # device tensors: read after the loop, no per-step sync
                self.currentStep = cur + 1
                cur = nxt
                if global_step % 400 == 399:
                    self.save_model(epoch, self.currentStep)
                    self.validate(global_step)
            self.currentStep = 0
        self.validate(global_step)
        return [float(l) for l in self.losses]

    def validate(self, step):
        """training.py:211-216: every validator runs on the local main process."""
        out = None
        if self.rank == 0:
            out = [v.validate(step, True) for v in self.validators]
            self.validation_results.append((step, out))
        return out

    def save_model(self, currentEpoch: int, currentStep: int = 0, savePath: Optional[str] = None):
        """training.py:218-229: ``accelerator.save_state(path)`` + ``epoch.json`` + barrier (the Azure mirror is out of
        scope).  The directory has accelerate's file layout (checkpoint.py): ``optimizer.bin`` loads into the reference's
        torch.optim.AdamW and the model file into the reference's CLIP (both pinned by tests); the scheduler and RNG files
        carry the key sets accelerate / CosineAnnealingWarmupRestarts write, restated without the packages (unpinned)."""
        path = savePath if savePath else "outputs/checkpoints"
        checkpoint.save_state(path, self.model, self.optimizer, self.stepper.sched_step,
                              total_steps=self.epochs * self.numBatches, rank=self.rank)
        if self.rank == 0:
            checkpoint.write_epoch_json(path, currentEpoch, currentStep)                # training.py:223
        if dist.is_initialized():
            dist.barrier()

    def load_model(self, path: str = "outputs/checkpoints"):
        try:
            sched_step = checkpoint.load_state(path, self.model, self.optimizer)       # training.py:243
            meta = checkpoint.read_epoch_json(path)                                     # training.py:244
        except Exception as e:                                                          # training.py:245-248
            if self.rank == 0 and os.path.isdir(path):
                print(f"Could not load model, starting from scratch because {e}")
            return 0, 0
        self.stepper.sched_step = sched_step
        self.model.mark_weights_dirty()
        store = self.model._require_store()
        if store.flat_w16 is not None:
            store.refresh_mirror(force=True)
        return meta["epoch"], meta["step"]
