"""Optimiser step next to the render path: gradient clipping by global norm and Adam
over flat parameter runs, one kernel each (csrc/train_step.cu), instead of the ~15 multi-tensor launches per
parameter group that `torch.optim.Adam` + `clip_grad_norm_` issue.

Reference: `configure_optimizers` builds torch.optim.Adam(betas=(0.9, 0.999)) with one group per learning rate
(static nets, dynamic nets) plus a second Adam for the poses, stepped together by `HybridOptim`
(train_online__.py:333-372, optimizer/hybrid_optimizer.py:75-84); the Trainer clips the global gradient norm to 1.0
(train_online__.py:1170).  `FusedAdam` keeps torch.optim.Adam's constructor, param_groups and state layout
(`step`, `exp_avg`, `exp_avg_sq`), so LR schedulers and state_dict round trips work unchanged.  No CPU path."""
import ctypes as C
import math

import torch

from . import _capi
from ._capi import StarAdamSeg, check, ptr, stream
from .functional import _count


def _ws(device):
    return torch.empty((_capi.lib().star_train_ws_bytes(),), device=device, dtype=torch.uint8)


def _runs(items):
    """Merges (addresses..., n, tag) records whose every address continues the previous record's into runs.
    items: list of (tuple_of_addresses, n, tag) sorted by the first address."""
    out = []
    for addrs, n, tag in items:
        if out:
            pa, pn, ptag = out[-1]
            if ptag == tag and all(a == b + 4 * pn for a, b in zip(addrs, pa)):
                out[-1] = (pa, pn + n, ptag)
                continue
        out.append((addrs, n, tag))
    return out


def _seg_array(runs):
    arr = (StarAdamSeg * max(1, len(runs)))()
    for i, (addrs, n, tag) in enumerate(runs):
        s = arr[i]
        s.param, s.grad, s.exp_avg, s.exp_avg_sq = addrs
        s.n = n
        s.step_size, s.bc2_sqrt = tag if tag is not None else (0.0, 1.0)
    return arr


def _check_grad(p, g):
    if not (g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() and not g.is_sparse):
        raise _capi.StarError("fused optimiser step: gradients must be dense contiguous fp32 CUDA tensors")
    if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
        raise _capi.StarError("fused optimiser step: parameters must be contiguous fp32 CUDA tensors")


def clip_grad_norm_(parameters, max_norm):
    """torch.nn.utils.clip_grad_norm_(parameters, max_norm) (L2 norm) in two launches: returns the total norm (0-dim
    tensor) and scales every gradient in place by min(1, max_norm / (total_norm + 1e-6))."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    ps = [p for p in parameters if p.grad is not None]
    if not ps:
        return torch.tensor(0.0)
    for p in ps:
        _check_grad(p, p.grad)
    dev = ps[0].device
    # only the gradient address has to continue for two tensors to share a run
    merged = []
    for g in sorted((p.grad for p in ps), key=lambda t: t.data_ptr()):
        if merged and g.data_ptr() == merged[-1][0][1] + 4 * merged[-1][1]:
            merged[-1] = (merged[-1][0], merged[-1][1] + g.numel(), None)
        else:
            merged.append(((0, g.data_ptr(), 0, 0), g.numel(), None))
    segs = _seg_array(merged)
    L = _capi.lib()
    ws = _ws(dev)
    check(L.star_grad_sqnorm(segs, len(merged), ptr(ws), stream()), "star_grad_sqnorm")
    check(L.star_grad_scale(segs, len(merged), L.star_grad_sqnorm_result(ptr(ws)), float(max_norm), stream()),
          "star_grad_scale")
    _count(2 * ((len(merged) + 15) // 16))
    return ws[8:16].view(torch.float64).sqrt().to(torch.float32).reshape(())


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas, eps) (weight_decay = 0, amsgrad = False) with the whole step -- optional
    clipping of the global gradient norm over all of this optimiser's parameters, moment updates, parameter update --
    in one pass over memory.  Parameters that sit back to back (see `flatten_parameters`) are updated as one run.

    max_grad_norm: clip_grad_norm_ semantics folded into the step (set it instead of Trainer(gradient_clip_val=...)).
    write_back_clipped_grads: also store the clipped gradients, as the reference's in-place clip leaves them."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False,
                 max_grad_norm=None, write_back_clipped_grads=False):
        if weight_decay != 0 or amsgrad:
            raise ValueError("FusedAdam covers the reference's configuration: weight_decay=0, amsgrad=False")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False))
        self.max_grad_norm = max_grad_norm
        self.write_back_clipped_grads = write_back_clipped_grads
        self.last_total_norm = None
        self._plan_key = None
        self._plan_cache = None

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._plan_key = None          # the loaded moments live in new tensors

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        self._plan_key = None

    def _init_state(self, params):
        """Moment buffers are allocated per run of adjacent parameters, so that they are adjacent too.  `step` is kept
        as a Python int (torch.optim.Adam accepts a number when such a state_dict is loaded into it)."""
        todo = sorted((p for p in params if len(self.state[p]) == 0), key=lambda p: p.data_ptr())
        runs = _runs([((p.data_ptr(),), p.numel(), None) for p in todo])
        i = 0
        for _, n, _t in runs:
            m = torch.zeros((n,), device=todo[i].device, dtype=torch.float32)
            v = torch.zeros_like(m)
            off = 0
            while off < n:
                p = todo[i]
                k = p.numel()
                st = self.state[p]
                st["step"] = 0
                st["exp_avg"] = m[off:off + k].view(p.shape)
                st["exp_avg_sq"] = v[off:off + k].view(p.shape)
                off += k
                i += 1

    def _plan(self, active):
        """Groups the active parameters into runs per (betas, eps); a run = tensors whose parameter, gradient and both
        moments all continue the previous tensor's, in the same group and at the same step count."""
        for _, p in active:
            _check_grad(p, p.grad)
        self._init_state([p for _, p in active])
        by_hyper = {}
        for gi, p in active:
            st = self.state[p]
            if torch.is_tensor(st["step"]):          # a state_dict saved by torch.optim.Adam
                st["step"] = int(st["step"])
            group = self.param_groups[gi]
            rec = ((p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()), p.numel(),
                   (gi, st["step"], p))
            by_hyper.setdefault((group["betas"][0], group["betas"][1], group["eps"]), []).append(rec)
        calls, all_runs = [], []
        for hyper, recs in by_hyper.items():
            recs.sort(key=lambda r: r[0][0])
            runs = []
            for addrs, n, (gi, t, p) in recs:
                if runs:
                    pa, pn, (pgi, pt, _p0) = runs[-1]
                    if pgi == gi and pt == t and all(a == b + 4 * pn for a, b in zip(addrs, pa)):
                        runs[-1] = (pa, pn + n, runs[-1][2])
                        continue
                runs.append((addrs, n, (gi, t, p)))
            calls.append((hyper, _seg_array([(a, n, None) for a, n, _ in runs]), len(runs),
                          [(gi, p) for _a, _n, (gi, _t, p) in runs]))
            all_runs += runs
        all_segs = _seg_array([(a, n, None) for a, n, _ in all_runs])
        return calls, all_segs, len(all_runs)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _capi.lib()
        active = [(gi, p) for gi, g in enumerate(self.param_groups) for p in g["params"] if p.grad is not None]
        if not active:
            return loss
        # the plan (runs and their ctypes array) is reused while the same tensors arrive at the same addresses
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for _, p in active)
        if key != self._plan_key:
            self._plan_cache = self._plan(active)
            self._plan_key = key
        calls, all_segs, n_all = self._plan_cache
        state = self.state
        for _, p in active:
            state[p]["step"] += 1
        dev = active[0][1].device
        sq = None
        if self.max_grad_norm is not None:
            ws = _ws(dev)
            check(L.star_grad_sqnorm(all_segs, n_all, ptr(ws), stream()), "star_grad_sqnorm")
            _count((n_all + 15) // 16)
            sq = L.star_grad_sqnorm_result(ptr(ws))
            self.last_total_norm = ws[8:16].view(torch.float64)     # squared norm; .sqrt() on demand
        for (b1, b2, eps), segs, n_runs, meta in calls:
            for i, (gi, p0) in enumerate(meta):
                t = state[p0]["step"]
                segs[i].step_size = self.param_groups[gi]["lr"] / (1.0 - b1 ** t)
                segs[i].bc2_sqrt = math.sqrt(1.0 - b2 ** t)
            check(L.star_adam_step(segs, n_runs, b1, b2, eps, sq, float(self.max_grad_norm or 0.0),
                                   1 if self.write_back_clipped_grads else 0, stream()), "star_adam_step")
            _count((n_runs + 15) // 16)
        # the kernels wrote through raw pointers: tell autograd / the packed-weight caches (they key on _version)
        torch.autograd.graph.increment_version([p for _, p in active])
        return loss

    def total_grad_norm(self):
        """Global gradient norm seen by the last step (before clipping); None without max_grad_norm."""
        return None if self.last_total_norm is None else self.last_total_norm.sqrt().to(torch.float32).reshape(())


def flatten_parameters(module):
    """Re-homes every parameter of `module` into one flat fp32 buffer (each NeRF net in the master order of
    include/star_b200.h, nets in module order, other parameters last), keeping shapes and values.  Afterwards each net's master vector
    is a zero-copy view (functional.flat_master), gradients of a net arrive as one run, and FusedAdam / the gradient
    all-reduce (parallel.py) touch one contiguous range per learning-rate group.  Call after .cuda() and before
    building the optimiser.  Returns the flat buffer."""
    ps = []
    for m in module.modules():          # nets first, each in the master order its kernels use
        rt = getattr(m, "_rt", None)
        if rt is not None and hasattr(rt, "ordered_params"):
            ps += list(rt.ordered_params())
    ps += list(module.parameters())
    seen, uniq = set(), []
    for p in ps:
        if id(p) not in seen:
            seen.add(id(p))
            uniq.append(p)
    if not uniq:
        return None
    dev = uniq[0].device
    total = sum(p.numel() for p in uniq)
    flat = torch.empty((total,), device=dev, dtype=torch.float32)
    off = 0
    with torch.no_grad():
        for p in uniq:
            if p.dtype != torch.float32:
                raise _capi.StarError("flatten_parameters: fp32 parameters only")
            n = p.numel()
            flat[off:off + n].copy_(p.detach().reshape(-1))
            p.data = flat[off:off + n].view(p.shape)
            off += n
    module._star_flat = (flat, uniq)       # parallel.GradSync lays the gradients out the same way
    return flat
