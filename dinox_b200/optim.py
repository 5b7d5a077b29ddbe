"""Fused optimizer step for the student: AdamW for every parameter and the global
gradient norm in one kernel launch.

Drop-in for `torch.optim.AdamW(student.parameters(), lr=args.lr, weight_decay=args.weight_decay)`
(scripts/phase5_big_run.py:1621) together with the gradient-norm loop at :1783-1789, which does one
`.item()` host synchronisation per parameter tensor.  State layout (`step`, `exp_avg`, `exp_avg_sq`) and
`state_dict()` are torch.optim.AdamW's own, so checkpoints (`save_checkpoint`, :1104-1125) load either way.
There is no CPU path: parameters must be fp32 CUDA tensors.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _ext, losshead


class FusedAdamW(torch.optim.AdamW):
    """torch.optim.AdamW with `step()` running as ONE multi-tensor CUDA kernel (+ a one-block reduction).

    After `step()`, `last_grad_norm` is a device scalar holding ||g||_2 over all parameters that had a
    gradient (what the reference logs as `total_norm`), available without a host sync.
    `grad_scale` multiplies the gradients first (pass 1/scaler.get_scale() under fp16 loss scaling)."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self._plans = {}
        self.last_grad_norm: Optional[torch.Tensor] = None

    def _plan(self, gi, ps):
        """chunk table over (p, grad, exp_avg, exp_avg_sq); rebuilt only when a gradient tensor moved
        (zero_grad(set_to_none=True) re-allocates them every accumulation window)"""
        gkey = tuple(p.grad.data_ptr() for p in ps)
        hit = self._plans.get(gi)
        if hit is not None and hit[0] == gkey and hit[2] == len(ps):
            return hit[1]
        if hit is not None:
            _ext.lib().dinox_adamw_plan_destroy(hit[1])
        n = len(ps)
        arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])
        ne = (ctypes.c_int64 * n)(*[p.numel() for p in ps])
        h = ctypes.c_void_p()
        _ext.call("dinox_adamw_plan_create", arr(ps), arr([p.grad for p in ps]), arr([self.state[p]["exp_avg"] for p in ps]),
                  arr([self.state[p]["exp_avg_sq"] for p in ps]), ne, n, ctypes.byref(h))
        self._plans[gi] = (gkey, h, n)
        return h

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        norms = []
        for gi, group in enumerate(self.param_groups):
            if group.get("amsgrad") or group.get("maximize"):
                raise NotImplementedError("FusedAdamW: amsgrad / maximize are not supported")
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            fresh = [p for p in ps if len(self.state[p]) == 0]
            for p in fresh:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise _ext.DinoxError("FusedAdamW needs contiguous fp32 CUDA parameters (no CPU fallback)")
                st = self.state[p]
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if fresh or gi not in self._plans:
                for p in ps:
                    if not (p.grad.is_cuda and p.grad.dtype == torch.float32 and p.grad.is_contiguous()):
                        raise _ext.DinoxError("FusedAdamW needs contiguous fp32 CUDA gradients")
                steps = {float(self.state[p]["step"]) for p in ps}
                if len(steps) != 1:
                    raise _ext.DinoxError("FusedAdamW: parameters of one group must share the step count")
            t = float(self.state[ps[0]]["step"]) + 1.0
            beta1, beta2 = group["betas"]
            h = self._plan(gi, ps)
            norm = torch.empty((), dtype=torch.float32, device=ps[0].device)
            _ext.call("dinox_adamw_step", h, float(group["lr"]), float(beta1), float(beta2), float(group["eps"]),
                      float(group["weight_decay"]), 1.0 - beta1 ** t, 1.0 - beta2 ** t, float(grad_scale),
                      ctypes.c_void_p(norm.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            torch._foreach_add_([self.state[p]["step"] for p in ps], 1.0)
            norms.append(norm)
        # the kernel wrote the parameters through raw pointers: tensor version counters did not move, so cached
        # bf16 operand copies of the head weights (losshead.bf16_weight, "tracked" mode) must be told
        losshead.invalidate_weight_cache()
        if norms:
            self.last_grad_norm = norms[0] if len(norms) == 1 else torch.stack(norms).square().sum().sqrt()
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for _, h, _n in self._plans.values():   # moment tensors were replaced: rebuild the chunk tables
            _ext.lib().dinox_adamw_plan_destroy(h)
        self._plans = {}

    def __del__(self):
        try:
            for _, h, _n in self._plans.values():
                _ext.lib().dinox_adamw_plan_destroy(h)
        except Exception:
            pass


def shard_layout(shapes, world: int, min_numel: int = 1 << 20):
    """Which parameters of a data-parallel replica set get a sharded optimizer state: those with at least
    `min_numel` elements whose first dimension divides by `world` (the head's W2: 25-100 M parameters).
    Returns [(sharded?, rows_per_rank or None)] per shape.  Pure host logic (tested on CPU)."""
    out = []
    for shp in shapes:
        n = 1
        for s in shp:
            n *= int(s)
        ok = world > 1 and len(shp) >= 1 and n >= min_numel and int(shp[0]) % world == 0
        out.append((ok, int(shp[0]) // world if ok else None))
    return out


class PeerGradShards:
    """Row shards of one large gradient (the head's dW2) in SYMMETRIC memory: rank r owns rows
    [r*rows, (r+1)*rows) as an fp32 accumulator that every rank has mapped into its address space
    (torch.distributed._symmetric_memory: CUDA VMM handles exchanged once, NVLink / NVSwitch peer access).
    The backward GEMM of every rank reduce-adds each of its output tiles straight into the owner's accumulator
    (`dinox_gemm_bf16_reduce_scatter`): the gradient reduce-scatter of a data-parallel step happens inside the
    GEMM epilogue, tile by tile, and the full (K, D) gradient is never materialised on any rank.

    Protocol (all ranks): accumulate over the micro-steps of a window -> any collective (ShardedFusedAdamW issues
    one) orders every rank's last kernel before the owner reads -> owner applies its optimizer slice and zeroes the
    accumulator -> the parameter all-gather orders the zeroing before the next window's adds."""

    def __init__(self, param: torch.Tensor, process_group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.pg = None if process_group in (None, True) else process_group
        grp = self.pg if self.pg is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(grp), dist.get_rank(grp)
        K, D = param.shape
        if K % (128 * self.world) or self.world > 8:
            raise ValueError(f"PeerGradShards: {K} rows do not split into {self.world} shards of whole 128-row tiles (<= 8 ranks)")
        self.rows, self.cols = K // self.world, D
        self.acc = symm.empty(self.rows, D, dtype=torch.float32, device=param.device)
        self.acc.zero_()
        self._hdl = symm.rendezvous(self.acc, grp.group_name)
        self.ptrs = [int(p) for p in self._hdl.buffer_ptrs]
        # the backward pass ships the gradient only when `flush` is set (the caller sets it for the LAST micro-step
        # of an accumulation window; earlier micro-steps accumulate into .grad locally, and their sum rides along)
        self.flush = True
        torch.cuda.synchronize()
        dist.barrier(group=self.pg)

    def zero_(self) -> None:
        from . import ops
        ops.fill_(self.acc.view(-1), 0.0)


class ShardedFusedAdamW:
    """Data-parallel optimizer step with the big tensors' AdamW state sharded over the ranks (SURVEY 8f next
    #2, second half; the reference is single-device - scripts/phase5_big_run.py:1781-1796 - so this is the
    data-parallel form of the same update):

      large parameters (see `shard_layout`):  reduce-scatter(mean) of .grad -> each rank owns rows
          [r*rows, (r+1)*rows) -> fused AdamW on that slice with ITS slice of exp_avg / exp_avg_sq ->
          in-place all-gather of the updated rows.  Per GPU: moments and AdamW traffic / world, and the same
          NVLink bytes as the all-reduce a DDP wrapper would issue (reduce-scatter + all-gather).
      small parameters: all-reduce(mean) of .grad, replicated fused AdamW.

    One launch per class (sharded / replicated) through `dinox_adamw_step`; `last_grad_norm` is the global
    ||mean-gradient||_2 as a device scalar (the shard norms are all-reduced, no host sync).  Replicas stay
    bit-identical: every rank applies the same arithmetic to the same reduced gradients.
    `consolidated_state_dict()` returns torch.optim.AdamW's layout for checkpoints."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 process_group=None, shard_min_numel: int = 1 << 20, grad_shards=None):
        import torch.distributed as dist
        self.params = [p for p in params]
        # torch.optim-style single parameter group: the reference loop writes `param_groups[i]['lr']` every step
        # (scripts/phase5_big_run.py:1699); lr / betas / eps / weight_decay are read from it at step time
        self.param_groups = [dict(params=self.params, lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)]
        self.pg = None if process_group in (None, True) else process_group
        self.world = dist.get_world_size(self.pg) if dist.is_initialized() else 1
        self.rank = dist.get_rank(self.pg) if dist.is_initialized() else 0
        for p in self.params:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise _ext.DinoxError("ShardedFusedAdamW needs contiguous fp32 CUDA parameters (no CPU fallback)")
        self.layout = shard_layout([tuple(p.shape) for p in self.params], self.world, shard_min_numel)
        self.step_count = 0
        self.state = {}
        for p, (sharded, rows) in zip(self.params, self.layout):
            own = p.data[self.rank * rows:(self.rank + 1) * rows] if sharded else p.data
            self.state[p] = dict(own=own, exp_avg=torch.zeros_like(own), exp_avg_sq=torch.zeros_like(own),
                                 grad=torch.empty_like(own) if sharded else None)
        self._plans = {}
        self.last_grad_norm: Optional[torch.Tensor] = None
        # {parameter: PeerGradShards}: gradients that arrive already reduce-scattered (mean over ranks) in the
        # owner's peer-mapped accumulator, written by the fused GEMM + reduce-scatter of the backward pass
        self.grad_shards = dict(grad_shards or {})
        for p, sh in self.grad_shards.items():
            lay = self.layout[[id(q) for q in self.params].index(id(p))]
            if not lay[0] or lay[1] != sh.rows:
                raise ValueError("grad_shards: the parameter is not sharded by this optimizer with the same row split")
        self._token = torch.zeros(1, device=self.params[0].device) if self.grad_shards else None

    def _plan(self, key, items):
        """items: [(param slice, grad, exp_avg, exp_avg_sq)]; rebuilt when a gradient tensor moved"""
        gkey = tuple(g.data_ptr() for _, g, _, _ in items)
        hit = self._plans.get(key)
        if hit is not None and hit[0] == gkey:
            return hit[1]
        if hit is not None:
            _ext.lib().dinox_adamw_plan_destroy(hit[1])
        n = len(items)
        col = lambda i: (ctypes.c_void_p * n)(*[it[i].data_ptr() for it in items])
        ne = (ctypes.c_int64 * n)(*[it[0].numel() for it in items])
        h = ctypes.c_void_p()
        _ext.call("dinox_adamw_plan_create", col(0), col(1), col(2), col(3), ne, n, ctypes.byref(h))
        self._plans[key] = (gkey, h)
        return h

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0):
        import torch.distributed as dist
        fused = [p for p in self.params if p in self.grad_shards]
        live = [(p, lay) for p, lay in zip(self.params, self.layout) if p.grad is not None or p in self.grad_shards]
        if not live:
            return
        sharded = [p for p, (s, _) in live if s]
        repl = [p for p, (s, _) in live if not s]
        works = []
        if self.world > 1:
            if fused:   # every rank's reduce-scatter GEMMs are complete once this (4-byte) collective is
                dist.all_reduce(self._token, group=self.pg)
            for p in sharded:   # mean over replicas, each rank receives its rows
                if p in self.grad_shards:
                    continue
                works.append(dist.reduce_scatter_tensor(self.state[p]["grad"].view(-1), p.grad.view(-1),
                                                        op=dist.ReduceOp.AVG, group=self.pg, async_op=True))
            for p in repl:
                works.append(dist.all_reduce(p.grad, op=dist.ReduceOp.AVG, group=self.pg, async_op=True))
            for w in works:
                w.wait()
        self.step_count += 1
        t = float(self.step_count)
        grp = self.param_groups[0]
        lr, eps, wd = grp["lr"], grp["eps"], grp["weight_decay"]
        b1, b2 = grp["betas"]
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        sq = torch.zeros(2, dtype=torch.float32, device=self.params[0].device)
        for slot, (key, ps) in enumerate((("sharded", sharded), ("replicated", repl))):
            if not ps:
                continue
            use_shard_grad = key == "sharded" and self.world > 1
            items = [(self.state[p]["own"],
                      self.grad_shards[p].acc if p in self.grad_shards else (self.state[p]["grad"] if use_shard_grad else p.grad),
                      self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"]) for p in ps]
            h = self._plan(key, items)
            _ext.call("dinox_adamw_step", h, float(lr), float(b1), float(b2), float(eps), float(wd),
                      1.0 - b1 ** t, 1.0 - b2 ** t, float(grad_scale), ctypes.c_void_p(sq[slot:].data_ptr()), stream)
        for p in fused:   # consumed: ready for the next accumulation window (the all-gather below orders this
            self.grad_shards[p].zero_()   # before any rank's next adds)
        sq.square_()
        if self.world > 1:
            if sharded:
                dist.all_reduce(sq[0:1], group=self.pg)       # every rank holds 1/world of the rows
            for p in sharded:                                 # in place: rank r's rows are already where they belong
                dist.all_gather_into_tensor(p.data.view(-1), self.state[p]["own"].view(-1), group=self.pg)
        self.last_grad_norm = sq.sum().sqrt()
        losshead.invalidate_weight_cache()   # parameters were written through raw pointers / p.data

    # lr, betas, eps, weight_decay: views of the parameter group, so `opt.lr = x` and `param_groups[0]['lr'] = x` agree
    lr = property(lambda self: self.param_groups[0]["lr"], lambda self, v: self.param_groups[0].__setitem__("lr", v))
    betas = property(lambda self: self.param_groups[0]["betas"])
    eps = property(lambda self: self.param_groups[0]["eps"])
    weight_decay = property(lambda self: self.param_groups[0]["weight_decay"],
                            lambda self, v: self.param_groups[0].__setitem__("weight_decay", v))

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    def consolidated_state_dict(self):
        """torch.optim.AdamW-shaped state (full-size moments on every rank) for `save_checkpoint` (:1104-1125)."""
        import torch.distributed as dist
        state = {}
        for i, (p, (s, _)) in enumerate(zip(self.params, self.layout)):
            st = self.state[p]
            if s and self.world > 1:
                full = [torch.empty_like(p.data), torch.empty_like(p.data)]
                for dst, src in zip(full, (st["exp_avg"], st["exp_avg_sq"])):
                    dist.all_gather_into_tensor(dst.view(-1), src.contiguous().view(-1), group=self.pg)
            else:
                full = [st["exp_avg"].clone(), st["exp_avg_sq"].clone()]
            state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": full[0], "exp_avg_sq": full[1]}
        group = dict(lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=self.weight_decay, amsgrad=False,
                     maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                     decoupled_weight_decay=True, params=list(range(len(self.params))))
        return {"state": state, "param_groups": [group]}

    def state_dict(self):
        """torch.optim.AdamW's layout (`opt.state_dict()` of the reference checkpoint, :1119); collective: every
        rank must call it (the sharded moments are all-gathered)."""
        return self.consolidated_state_dict()

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        """Inverse of `state_dict()` (resume, scripts/phase5_big_run.py:1168): accepts torch.optim.AdamW's layout
        with full-size moments - written by this class, by FusedAdamW or by torch.optim.AdamW - and keeps only the
        rows this rank owns.  Hyper-parameters of the stored group replace the current ones."""
        groups = state_dict["param_groups"]
        order = [i for g in groups for i in g["params"]]
        if len(order) != len(self.params):
            raise ValueError(f"optimizer state holds {len(order)} parameters, this optimizer {len(self.params)}")
        steps = set()
        for pos, idx in enumerate(order):
            p, (sharded, rows) = self.params[pos], self.layout[pos]
            src = state_dict["state"].get(idx)
            st = self.state[p]
            if src is None:      # a parameter that never received a gradient
                st["exp_avg"].zero_(); st["exp_avg_sq"].zero_()
                continue
            for k in ("exp_avg", "exp_avg_sq"):
                full = src[k].to(device=p.device, dtype=torch.float32)
                if tuple(full.shape) != tuple(p.shape):
                    raise ValueError(f"{k} of parameter {pos}: shape {tuple(full.shape)} != {tuple(p.shape)}")
                st[k].copy_(full[self.rank * rows:(self.rank + 1) * rows] if sharded else full)
            steps.add(float(src["step"]))
        if len(steps) > 1:
            raise ValueError(f"ShardedFusedAdamW keeps one step count; the state holds {sorted(steps)}")
        self.step_count = int(steps.pop()) if steps else 0
        g0 = groups[0]
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in g0:
                self.param_groups[0][k] = tuple(g0[k]) if k == "betas" else g0[k]

    def __del__(self):
        try:
            for _, h in self._plans.values():
                _ext.lib().dinox_adamw_plan_destroy(h)
        except Exception:
            pass
