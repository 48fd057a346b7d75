"""Zero-shot scoring shape of the reference's ImageNet validator (training/clip/validation.py:119-139,142-179;
BASELINE.json configs[4]): per class, encode the prompt templates, L2-normalise, average, re-normalise
(:125-131) -> classifier W [E, classes]; per image batch, encode, normalise, ``100 * f @ W`` (:157-162), top-k
(:136-139).  The encoders are the training-path kernels in inference mode (no activation saving); the small
scoring GEMM runs on the fp32 FFMA engine."""
from __future__ import annotations

from typing import Sequence

import torch

from . import ops
from .ops import MAJOR_K, MAJOR_MN


@torch.no_grad()
def zeroshot_classifier(model, class_token_batches: Sequence[torch.Tensor]) -> torch.Tensor:
    """class_token_batches[c]: int tensor [templates, context_length] of the prompts of class c
    (validation.py:123-124 tokenises 80 templates per class).  Returns W [embed_dim, num_classes] fp32."""
    cols = []
    for tokens in class_token_batches:
        e = model.encode_text(tokens)                                  # :125-128
        e = e / e.norm(dim=-1, keepdim=True)                           # :129
        e = e.mean(dim=0)                                              # :130
        cols.append(e / e.norm())                                      # :131
    return torch.stack(cols, dim=1).contiguous()                       # :133


@torch.no_grad()
def zeroshot_logits(model, images: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    """100 * normalise(encode_image(images)) @ W   (validation.py:157-162) -> [batch, classes] fp32."""
    f = model.encode_image(images)
    f = (f / f.norm(dim=-1, keepdim=True)).contiguous()
    B, E = f.shape
    C = W.shape[1]
    out = torch.empty(B, C, device=f.device, dtype=torch.float32)
    # logits[b, c] = sum_e f[b, e] W[e, c]: A = f (K-major), B[n=c][k=e] = W[e, c] (MN-major)
    ops.gemm("simt", B, C, E, 1, f, MAJOR_K, E, 0, W, MAJOR_MN, C, 0, out, C, 0)
    return out.mul_(100.0)


class ZeroShotScorer:
    """Config 5 at full fidelity (validation.py:119-134,142-179): ``classes x templates`` prompts (ImageNet: 1000 x 80)
    -> classifier -> ``100 * f @ W`` -> top-k, with both encoders replayed from CUDA graphs.

    The text encoder runs over chunks of ``classes_per_chunk`` whole classes ([chunk * T, context] token rows); the
    per-class mean of the L2-normalised prompt embeddings and its re-normalisation (:129-131) are part of the same
    captured graph, so a 1000-class classifier is ``ceil(1000 / chunk)`` graph replays and no eager kernel.  The image
    side is one graph per image-batch shape: encode, normalise, scoring GEMM, x100."""

    def __init__(self, model, templates: int, classes_per_chunk: int = 50, use_cuda_graph: bool = True):
        self.model, self.T, self.chunk, self.use_graph = model, int(templates), int(classes_per_chunk), use_cuda_graph
        self.E = model._cfg["embed_dim"]
        self._tg = None          # (graph, static tokens, static out [chunk, E])
        self._ig = {}            # image batch shape -> (graph, static images, static logits)
        self.W = None

    def _text_chunk(self, tokens, out):
        m = self.model
        m._require_store()
        m._prepare_weights()
        ws = m._towers["text"].forward(tokens, m._precision, False)
        e = ws.u_feat.view(self.chunk, self.T, self.E).mean(dim=1)            # ws.u_feat is already L2-normalised (:129-130)
        torch.div(e, e.norm(dim=-1, keepdim=True), out=out)                   # :131

    @torch.no_grad()
    def build_classifier(self, tokens: torch.Tensor) -> torch.Tensor:
        """tokens: int [classes, templates, context] on the device.  Returns (and keeps) W [E, classes] fp32."""
        C, T, ctx = tokens.shape
        if T != self.T:
            raise ValueError(f"expected {self.T} templates per class, got {T}")
        dev = tokens.device
        Wt = torch.empty(C, self.E, device=dev)
        if self._tg is None:
            st = torch.zeros(self.chunk * T, ctx, device=dev, dtype=torch.int64)
            so = torch.empty(self.chunk, self.E, device=dev)
            st[:, 0] = 1                                                      # any valid token row for the warm-up
            graph = None
            if self.use_graph:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._text_chunk(st, so)
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._text_chunk(st, so)
            self._tg = (graph, st, so)
        graph, st, so = self._tg
        flat = tokens.reshape(C * T, ctx).to(torch.int64)
        for c0 in range(0, C, self.chunk):
            c1 = min(C, c0 + self.chunk)
            rows = (c1 - c0) * T
            st[:rows].copy_(flat[c0 * T:c1 * T])
            if rows < st.shape[0]:
                st[rows:].copy_(flat[:1].expand(st.shape[0] - rows, ctx))      # ragged last chunk: padding rows, discarded
            if graph is not None:
                graph.replay()
            else:
                self._text_chunk(st, so)
            Wt[c0:c1].copy_(so[:c1 - c0])
        if self.W is not None and tuple(self.W.shape) == (self.E, C):
            self.W.copy_(Wt.t())                                              # same buffer: the image-side graph stays valid
        else:
            self.W = Wt.t().contiguous()                                      # :133  [E, classes]
        return self.W

    def _image_batch(self, images, logits):
        m = self.model
        m._require_store()
        m._prepare_weights()
        ws = m._towers["image"].forward(images, m._precision, False)
        B, C = images.shape[0], self.W.shape[1]
        ops.gemm("simt", B, C, self.E, 1, ws.u_feat, MAJOR_K, self.E, 0, self.W, MAJOR_MN, C, 0, logits, C, 0)
        logits.mul_(100.0)                                                    # :162

    @torch.no_grad()
    def logits(self, images: torch.Tensor) -> torch.Tensor:
        """100 * normalise(encode_image(images)) @ W -> [batch, classes]; the returned tensor is reused by the next call."""
        if self.W is None:
            raise ValueError("build_classifier first")
        key = (tuple(images.shape), images.dtype, self.W.data_ptr())
        if key not in self._ig:
            si = images.clone()
            so = torch.empty(images.shape[0], self.W.shape[1], device=images.device)
            graph = None
            if self.use_graph:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._image_batch(si, so)
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._image_batch(si, so)
            self._ig = {key: (graph, si, so)}
        graph, si, so = self._ig[key]
        si.copy_(images, non_blocking=True)
        if graph is not None:
            graph.replay()
        else:
            self._image_batch(si, so)
        return so


class ZeroShotValidator:
    """Shape of the reference's ``ImageNetValidator`` (validation.py:114-181) over ZeroShotScorer: built with the class
    prompts (token ids ``[classes, templates, context]``; the reference tokenises 1000 names x 80 templates, :119-124) and
    a loader of ``(images, target)`` batches, ``validate(step, verbose)`` switches the model to eval, rebuilds the
    classifier from the current weights (:149), scores every batch (``100 * f @ W``, :157-162) and returns / logs top-1 and
    top-5 accuracy in percent (:164-179).  The datasets themselves (ImageNetV2, ...) are out of scope: the caller
    brings the loader."""

    def __init__(self, trainer, class_tokens: torch.Tensor, loader, writer=None, classes_per_chunk: int = 50,
                 use_cuda_graph: bool = True):
        self.trainer, self.class_tokens, self.loader, self.writer = trainer, class_tokens, loader, writer
        self.scorer = ZeroShotScorer(trainer.model, class_tokens.shape[1], classes_per_chunk, use_cuda_graph)

    @torch.no_grad()
    def validate(self, step, verbose=False):
        model = self.trainer.model
        was_training = model.training
        model.eval()                                                          # :143-146
        dev = model.logit_scale.device
        self.scorer.build_classifier(self.class_tokens.to(dev))               # :149
        top1 = top5 = n = 0.0
        for images, target in self.loader:                                    # :153-168
            logits = self.scorer.logits(images.to(dev))
            a1, a5 = accuracy(logits, target.to(dev), topk=(1, 5))
            top1, top5, n = top1 + a1, top5 + a5, n + images.shape[0]
        top1, top5 = top1 / max(n, 1) * 100, top5 / max(n, 1) * 100           # :170-171
        if verbose:
            print(f"Top-1 accuracy: {top1:.2f}%")
            print(f"Top-5 accuracy: {top5:.2f}%")
        if self.writer is not None:
            self.writer.add_scalar("Top-1 accuracy", top1, step)
            self.writer.add_scalar("Top-5 accuracy", top5, step)
        model.train(was_training)
        return {"top1": top1, "top5": top5, "n": int(n)}


def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1, 5)):
    """validation.py:136-139."""
    pred = output.topk(max(topk), 1, True, True)[1].t()
    correct = pred.eq(target.view(1, -1).expand_as(pred))
    return [float(correct[:k].reshape(-1).float().sum().item()) for k in topk]
