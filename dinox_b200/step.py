"""One loss-head micro-step: the step glue of the reference training loop
(scripts/phase5_big_run.py:1738-1802) around the drop-in modules, on pre-extracted features.

  loss = L_dino (+ w_ibot * L_ibot) + w_gram * L_gram ; loss / accum ; backward ;
  every `accum`-th micro-step: [optimizer step is the host's] EMA of all parameters, grads reset.
"""
from __future__ import annotations

import contextlib
import os
from typing import Dict, List, Optional

import torch

from . import losshead, ops, synth


class LossHeadStep:
    def __init__(self, shapes: synth.LossHeadShapes, device, seed_cfg: int = 2, rank: int = 0,
                 accum: int = 4, gram_weight: float = 1.0, ibot_weight: float = 1.0, ema: float = 0.996,
                 center_momentum: float = 0.9, student_temp: float = 0.1, teacher_temp: float = 0.04,
                 teacher_mode: str = "center", process_group=None, with_backbone_params: bool = True,
                 koleo_weight: float = 0.0, patch_teacher_mode: Optional[str] = None):
        self.shapes, self.device, self.accum = shapes, device, accum
        self.gram_weight, self.ibot_weight, self.ema = gram_weight, ibot_weight, ema
        # KoLeo (scripts/phase5_big_run.py:1764-1766) is off by default: the benchmark metric is the
        # north_star loss head; with a weight the global-view CLS logits are materialised for it
        self.koleo_weight = koleo_weight
        self.koleo = losshead.KoLeoLoss() if koleo_weight > 0.0 else None
        self.student_temp, self.teacher_temp = student_temp, teacher_temp
        g = synth.seeded_generator(seed_cfg, 0)  # identical replicas on every rank
        D, K = shapes.dim, shapes.out_dim
        self.student_head = losshead.ProjectionHead(D, K)
        self.teacher_head = losshead.ProjectionHead(D, K)
        self.student_head.load_state_dict(synth.head_weights(D, K, g))
        self.teacher_head.load_state_dict(synth.head_weights(D, K, g))
        self.student_head.to(device)
        self.teacher_head.to(device)
        for p in self.teacher_head.parameters():
            p.requires_grad_(False)
        self.dino_loss = losshead.DINOLoss(K, center_momentum, n_global=shapes.n_global, n_local=shapes.n_local,
                                           teacher_mode=teacher_mode, process_group=process_group,
                                           patch_teacher_mode=patch_teacher_mode).to(device)
        self.center_patch = torch.zeros(1, K, device=device)
        self._unit = torch.ones((), dtype=torch.float32, device=device)
        self._seeds: Dict[float, torch.Tensor] = {}
        # every other student/teacher parameter (backbone, scale-embed): random stand-ins with the
        # reference's shapes so that the EMA walks the real 161 / 305 tensor list
        self.student_params: List[torch.Tensor] = []
        self.teacher_params: List[torch.Tensor] = []
        if with_backbone_params:
            depth = synth.BACKBONES.get(D, dict(depth=12))["depth"]
            for shp in synth.student_param_shapes(D, depth, K)[:-4]:
                self.student_params.append(torch.randn(shp, device=device) * 0.02)
                self.teacher_params.append(torch.randn(shp, device=device) * 0.02)
        self.student_params += list(self.student_head.parameters())
        self.teacher_params += list(self.teacher_head.parameters())
        self.n_params = sum(p.numel() for p in self.student_params)
        # this loop updates parameters only through dinox entry points (ema_update bumps the weight epoch), so the
        # bf16 operand copies of the head weights are re-cast only when a weight changed
        losshead.set_weight_cache("tracked")
        # data parallel: an optim.PeerGradShards of student_head[2].weight turns the dW2 GEMM into GEMM + reduce-scatter
        self.w2_grad_shards = None
        self.micro = 0
        self._graphs: Dict[int, dict] = {}
        self.launches_per_graph = 0

    def micro_step(self, f: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """f: student_cls, teacher_cls (+ student_tok, teacher_tok) (+ student_patch, teacher_patch,
        masks_weight); student tensors may require grad.  Returns the component losses."""
        out, loss = self._losses(f)      # already divided by accum
        loss.backward()
        self.micro += 1
        if self.micro % self.accum == 0:
            # the optimizer step belongs to the host loop (scripts/phase5_big_run.py:1794-1796); then EMA
            losshead.ema_update(self.teacher_params, self.student_params, self.ema, plan_key=id(self))
            for p in self.student_head.parameters():
                p.grad = None
        return out

    # ---------------------------------------------------------------------------------------------
    # CUDA-graph stepping: the ~60 launches of a micro-step (kernels of this library + the handful of
    # scalar autograd ops of the step glue) are captured once per input slot and replayed, so the step
    # is never bound by Python / launch latency (it is at the smaller per-GPU batches of C3 / C5).
    # ---------------------------------------------------------------------------------------------
    def _head_weights(self):
        return [self.student_head[0].weight, self.student_head[2].weight,
                self.teacher_head[0].weight, self.teacher_head[2].weight]

    def static_inputs(self, like: Dict[str, torch.Tensor], slots: int = 1,
                      buffers: Optional[List[Dict[str, torch.Tensor]]] = None) -> List[Dict[str, torch.Tensor]]:
        """Device input buffers (one dict per slot) with the shapes / dtypes of `like`; the caller
        fills them (H2D copies land here directly) and calls micro_step_graph(slot).  `buffers` lets the
        caller supply the storage (e.g. typed views into one blob per slot: a single H2D copy per step)."""
        out = []
        for i in range(slots):
            d = {}
            for k, v in like.items():
                t = buffers[i][k] if buffers is not None else torch.empty(v.shape, dtype=v.dtype, device=self.device)
                assert t.shape == v.shape and t.dtype == v.dtype and t.is_cuda
                if k.startswith("student"):
                    t.requires_grad_(True)
                d[k] = t
            out.append(d)
        self._static = out
        self._graphs = {}
        return out

    def _ensure_grads(self):
        for p in self.student_head.parameters():
            if p.grad is None:
                p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)

    def capture(self, slot: int = 0, warmup: int = 2) -> None:
        """Capture forward + backward of one micro-step on the static inputs of `slot`."""
        f = self._static[slot]
        self._ensure_grads()
        was_timing = ops.TIMER.enabled
        ops.TIMER.enabled = False
        # the warm-up passes must leave no trace: centres and accumulated head gradients are restored
        saved = [self.dino_loss.center.clone(), self.center_patch.clone()] + [p.grad.clone() for p in self.student_head.parameters()]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):           # warm-up on a side stream (PyTorch capture recipe)
            for _ in range(warmup):
                for k, v in f.items():
                    if v.requires_grad:
                        v.grad = None
                self._fwd_bwd(f)
        torch.cuda.current_stream().wait_stream(side)
        with torch.no_grad():
            self.dino_loss.center.copy_(saved[0])
            self.center_patch.copy_(saved[1])
            for p, g0 in zip(self.student_head.parameters(), saved[2:]):
                p.grad.copy_(g0)
        for k, v in f.items():
            if v.requires_grad:
                v.grad = None                    # backward allocates the input grads inside the capture
        for w in self._head_weights():
            losshead.bf16_weight(w)              # casts happen OUTSIDE the graph (refreshed in place later)
        g = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        # the main chain of the step is captured on a high-priority stream (kernel nodes keep the priority of the stream
        # they were captured on): its blocks are dispatched ahead of the Gram / centre side work when SMs free up
        cap_stream = torch.cuda.Stream(device=self.device, priority=-1) if losshead.stream_priorities() else None
        with torch.cuda.graph(g, stream=cap_stream):
            out = self._fwd_bwd(f)
        self.launches_per_graph = ops.launch_count() - n0
        self._graphs[slot] = dict(graph=g, out=out)
        ops.TIMER.enabled = was_timing

    def _fwd_bwd(self, f):
        """Forward + backward of one micro-step.  The loss terms are NOT summed inside the autograd graph: the gradient of
        the total with respect to term i is the constant w_i / accum, so every term's backward is seeded with that
        constant (torch.autograd.backward on the list of terms) and the head's backward - the critical chain - starts
        right behind pass 2 instead of behind the side branches (Gram anchoring's forward used to gate it through the
        sum).  The total, for reporting, is formed on the side stream once its last term is there."""
        with losshead.contraction_precision("bf16"):
            out, terms, weights, side = self._loss_terms(f)
        main = torch.cuda.current_stream()
        if side is not None and os.environ.get("DINOX_DECOUPLED", "1") == "0":   # A/B knob: the sum inside the graph
            main.wait_stream(side)
            side = None
        if side is None:
            scaled, total = losshead.combine_losses(terms, weights, 1.0 / self.accum)
            out["loss_total"] = total
            scaled.backward(self._unit)      # a resident 1.0: no ones_like fill per step
            return out
        side.wait_stream(main)               # the head's loss values (pass 2) are on the main stream
        with torch.cuda.stream(side):
            total = torch.empty((), dtype=torch.float32, device=self.device)
            ops.scalar_combine([t.detach().reshape(()) for t in terms], weights, 1.0 / self.accum, out_unscaled=total)
            out["loss_total"] = total
        seeds = [self._seed(w) for w in weights]
        torch.autograd.backward(terms, seeds)
        main.wait_stream(side)               # Gram forward / backward and the total are part of this step
        return out

    def _seed(self, w: float) -> torch.Tensor:
        """d(total / accum) / d(term) = w / accum as a resident device scalar, rounded like the fan-out kernel of
        combine_losses rounds it (fp32: (1 * 1/accum) * w), so both routes give the same gradients bit for bit."""
        t = self._seeds.get(w)
        if t is None:
            import numpy as np
            v = np.float32(np.float32(1.0) * np.float32(1.0 / self.accum)) * np.float32(w)
            t = self._seeds[w] = torch.full((), float(v), dtype=torch.float32, device=self.device)
        return t

    def _losses(self, f):
        # this step is the bf16 tensor-core configuration of the benchmark (BASELINE.json: bf16 operands, fp32
        # accumulation) whatever the global contraction precision of the drop-in modules is
        with losshead.contraction_precision("bf16"):
            out, terms, weights, side = self._loss_terms(f)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        # loss = (L_dino + L_ibot + w_g L_gram + w_k L_koleo) / accum  (scripts/phase5_big_run.py:1749-1769), one launch
        scaled, total = losshead.combine_losses(terms, weights, 1.0 / self.accum)
        out["loss_total"] = total
        return out, scaled

    def _loss_terms(self, f):
        """Forward of all loss terms -> (outputs, terms, weights, side stream of the Gram term or None).  Gram anchoring
        is independent of the head: it is issued first, on a side stream, and its small kernels (and, by autograd's
        stream rule, their backward) run beside the head's GEMMs instead of between them."""
        gram, side, stok, rows = None, None, None, None
        if "student_tok" in f:
            stok = f["student_tok"]
            if "patch_index" in f and stok.requires_grad:
                stok, rows = losshead.token_fork(stok, f["patch_index"])
            if losshead.concurrency() >= 1 and not ops.TIMER.enabled:
                side = losshead._side_stream(self.device, 1)
                side.wait_stream(torch.cuda.current_stream())
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                gram = losshead.compute_gram_anchoring_loss(stok, f["teacher_tok"])
        # iBOT rows: materialised ("student_patch"/"teacher_patch") or named by "patch_index" inside the token tensors
        by_index = "patch_index" in f
        sp = f["student_tok"] if by_index else f.get("student_patch")
        tp = f["teacher_tok"] if by_index else f.get("teacher_patch")
        idx, t_idx = f.get("patch_index"), None
        if by_index and rows is not None:
            # the token tensor feeds Gram anchoring AND (its masked rows) the iBOT term: the fork hands the rows to the
            # head as a matrix and adds their gradient into the Gram gradient in place (see losshead.token_fork)
            sp, idx, t_idx = rows, None, f["patch_index"]
        out = losshead.fused_head_dino_loss(
            f["student_cls"], f["teacher_cls"], self.student_head, self.teacher_head, self.dino_loss,
            self.student_temp, self.teacher_temp, student_patch=sp, teacher_patch=tp,
            masks_weight=f.get("masks_weight"), center_patch=self.center_patch if sp is not None else None,
            ibot_weight=self.ibot_weight, patch_index=idx, teacher_patch_index=t_idx, grads_in_place=True,
            w2_grad_shards=self.w2_grad_shards)
        terms, weights = [out["loss"]], [1.0]
        if gram is not None:
            out["loss_gram"] = gram
            terms.append(gram)
            weights.append(self.gram_weight)
        if self.koleo is not None:
            n_glob = self.shapes.batch * self.shapes.n_global       # the reference feeds its 2B global-view rows
            z = self.student_head(f["student_cls"][:n_glob])
            out["loss_koleo"] = self.koleo(z)
            terms.append(out["loss_koleo"])
            weights.append(self.koleo_weight)
        return out, terms, weights, side

    def micro_step_graph(self, slot: int = 0) -> Dict[str, torch.Tensor]:
        """Replay the captured micro-step on the current contents of the slot's static inputs.  Returns
        the static output tensors (losses); input gradients are in `static_inputs[slot][k].grad`, head
        gradients accumulate in the parameters' .grad (zeroed at the start of every accumulation window)."""
        if slot not in self._graphs:
            self.capture(slot)
        if self.micro % self.accum == 0:
            for p in self.student_head.parameters():
                ops.fill_(p.grad.view(-1), 0.0)
        for w in self._head_weights():
            losshead.bf16_weight(w)              # no-op unless a weight changed since the last cast
        self._graphs[slot]["graph"].replay()
        self.micro += 1
        if self.micro % self.accum == 0:
            losshead.ema_update(self.teacher_params, self.student_params, self.ema, plan_key=id(self))
        return self._graphs[slot]["out"]
