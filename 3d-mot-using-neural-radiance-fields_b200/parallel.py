"""Ray-sharded data parallelism for the render path (one process per GPU, `torch.distributed`).

The reference is single-GPU (`Trainer(devices=1)`, train_online__.py:1165-1166).  Rays are independent
(callbacks/check_batch_grad.py:8-50 asserts it), so the path shards with no data-path
collective: every rank renders a contiguous slice of the ray list with replicated weights.  Training adds one
all-reduce of the flat gradient (MLP weights + pose parameters) per optimiser step; rendering all-gathers the
per-ray outputs.  Backend: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank, world_size):
    """Contiguous, balanced slice [start, end) of n items for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_rays(*tensors, rank=None, world_size=None):
    """Slices every [R, ...] tensor to this rank's rays."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    a, b = shard_bounds(tensors[0].shape[0], rank, world_size)
    out = tuple(t[a:b] for t in tensors)
    return out if len(out) > 1 else out[0]


def _grad_runs(params, max_runs=32):
    """Gradients that already sit back to back in memory (a net's flat gradient, parameters re-homed by
    optim.flatten_parameters) as flat zero-copy views, one per run; None when a gradient is missing, not dense
    contiguous fp32, or the gradients are scattered over more than `max_runs` pieces."""
    gs = []
    for p in params:
        g = p.grad
        if g is None or g.dtype != torch.float32 or g.is_sparse or not g.is_contiguous():
            return None
        gs.append(g)
    # Runs are formed in LIST order and only inside one storage, so that every rank derives the same run structure
    # (address order and accidental adjacency of separate allocations differ between processes).
    runs = []          # [first tensor, element count, end address]
    for g in gs:
        if g.numel() == 0:
            continue
        if runs and g.data_ptr() == runs[-1][2] and \
                g.untyped_storage().data_ptr() == runs[-1][0].untyped_storage().data_ptr():
            runs[-1][1] += g.numel()
            runs[-1][2] += 4 * g.numel()
            continue
        runs.append([g, g.numel(), g.data_ptr() + 4 * g.numel()])
        if len(runs) > max_runs:
            return None
    return [torch.as_strided(first, (n,), (1,)) for first, n, _ in runs]


def allreduce_gradients(params, average=True, group=None):
    """All-reduce(sum) of every .grad across ranks, in place.  Gradients that form a few contiguous runs are reduced
    where they lie (one collective per run, no staging copy); otherwise one all-reduce over a concatenated fp32
    buffer (parameters without a gradient contribute zeros, so all ranks agree on the layout), scattered back.
    average=True divides by the world size: with equal ray shards the mean of per-rank mean losses is the global
    mean loss.  Every rank must hold the same parameter list with the same gradient layout."""
    params = [p for p in params if p.requires_grad]
    _, w = world()
    if not params or w == 1:
        return None                # single process: nothing to exchange, and no .grad is materialised
    runs = _grad_runs(params)
    if runs is not None:
        if w > 1:
            for r in runs:
                dist.all_reduce(r, op=dist.ReduceOp.SUM, group=group)
                if average:
                    r.div_(w)
        return runs
    # Parameters without a gradient contribute zeros so that all ranks agree on the layout; a presence count travels
    # with the buffer, and a parameter that NO rank has a gradient for keeps .grad = None -- the single-process
    # reference skips such parameters in Adam (no step count, no momentum decay), and so must every rank here.
    have = torch.tensor([0.0 if p.grad is None else 1.0 for p in params], device=params[0].device)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params]
                     + [have])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    present = flat[-len(params):] > 0
    present = present.tolist()
    if average:
        flat.div_(w)
    off = 0
    for p, any_rank in zip(params, present):
        n = p.numel()
        if any_rank:
            g = flat[off:off + n].view_as(p).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
        off += n
    return flat[:-len(params)]


class GradSync:
    """Gradient exchange of a ray-sharded training step, overlapped with the backward pass (one
    all-reduce(sum) over the flat fp32 gradient per optimiser step, before gradient clipping, train_online__.py:1170).

        sync = GradSync(net, extra_params=[pose])        # after .cuda(), BEFORE the optimiser is built
        loss = sync.scale_loss(loss); loss.backward(); sync.finish(); optimiser.step()

    * parameters are re-homed into one flat buffer (optim.flatten_parameters) and every .grad becomes a view into ONE
      flat gradient buffer of the same layout, so a net's gradient is one contiguous range and the whole model one run
      for FusedAdam / clip_grad_norm_;
    * the MLP backward kernels accumulate straight into that buffer (functional.NerfRaw: the net's `grad_sink`), and as
      soon as a net's backward has been launched its range is all-reduced on a side stream: the fine net's collective
      runs under the coarse net's backward, only the last one is exposed;
    * averaging costs nothing: scale_loss() multiplies the loss by 1 / world_size, so the summed gradients are the mean;
    * finish() reduces what is left (nets whose backward did not run, other parameters, `extra_params` such as the
      pose vector) in one more collective and makes the current stream wait for all of them.  No host synchronisation.
    The buffer is zeroed by the first forward after finish(), i.e. gradient accumulation over several backward passes
    per optimiser step works; zero_grad(set_to_none=True) by a trainer is tolerated (views are re-attached in finish)."""

    def __init__(self, module, extra_params=(), group=None, overlap=True):
        from . import optim
        self.overlap = overlap           # False: reduce everything in finish() (several backward passes per step)
        if getattr(module, "_star_flat", None) is None:
            optim.flatten_parameters(module)
        self.flat, self.order = module._star_flat
        self.group = group
        self.extra = [p for p in extra_params if p.requires_grad]
        self.flat_grad = torch.zeros_like(self.flat)
        self.views, off = [], 0
        index = {}
        for p in self.order:
            n = p.numel()
            v = self.flat_grad[off:off + n].view(p.shape)
            self.views.append(v)
            index[id(p)] = (off, off + n)
            if p.requires_grad:
                p.grad = v
            off += n
        self.n_flat = off
        # nets: contiguous ranges in the kernels' master order -> gradient sinks
        self.sinks = []
        for m in module.modules():
            rt = getattr(m, "_rt", None)
            if rt is None or not hasattr(rt, "ordered_params"):
                continue
            ps = rt.ordered_params()
            a, b = index[id(ps[0])][0], index[id(ps[-1])][1]
            if b - a != sum(q.numel() for q in ps) or not all(q.requires_grad for q in ps):
                continue
            rt.grad_sink, rt.sink_owner = self.flat_grad[a:b], self
            self.sinks.append((rt, a, b))
        self.stream = torch.cuda.Stream(self.flat.device) if self.flat.is_cuda else None
        self._works, self._done, self._dirty = [], set(), False

    # ---- hooks called by functional.NerfRaw
    def on_forward(self, rt):
        if self._dirty:                      # first forward of a new optimiser step
            self.flat_grad.zero_()
            for p in self.extra:
                p.grad = None
            self._dirty = False
            self._done = set()
        elif rt is not None and id(rt) in self._done:
            raise RuntimeError("GradSync: a net runs forward again while its gradient of this step is already being "
                               "all-reduced; with several backward passes per optimiser step use GradSync(overlap=False)")

    def on_backward_launched(self, rt):
        """All kernels that write rt.grad_sink are queued on the current stream: reduce it behind them, on the side."""
        _, w = world()
        if w == 1 or not self.overlap or id(rt) in self._done:
            return
        self._done.add(id(rt))
        self._reduce_async(rt.grad_sink)

    def _reduce_async(self, t):
        if self.stream is None:
            self._works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            return
        ev = torch.cuda.Event()
        ev.record()
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            self._works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    # ---- user side
    def scale_loss(self, loss):
        _, w = world()
        return loss if w == 1 else loss * (1.0 / w)

    def zero_grad(self):
        self.flat_grad.zero_()
        for p in self.extra:
            p.grad = None
        self._dirty, self._done = False, set()
        for p, v in zip(self.order, self.views):
            if p.requires_grad:
                p.grad = v

    def finish(self):
        _, w = world()
        # a trainer may have dropped the views (zero_grad(set_to_none=True)): fold what autograd allocated back in
        for p, v in zip(self.order, self.views):
            if not p.requires_grad:
                continue
            g = p.grad
            if g is None or g.data_ptr() != v.data_ptr():
                if g is not None:
                    v.add_(g)
                p.grad = v
        if w > 1:
            # everything outside the ranges already in flight, as contiguous pieces of the flat buffer
            pieces, pos = [], 0
            for rt, a, b in sorted(self.sinks, key=lambda r: r[1]):
                if id(rt) in self._done:
                    if a > pos:
                        pieces.append((pos, a))
                    pos = max(pos, b)
            if pos < self.n_flat:
                pieces.append((pos, self.n_flat))
            for a, b in pieces:
                self._reduce_async(self.flat_grad[a:b])
            ex = [p for p in self.extra]
            if ex:
                buf = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in ex])
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
                off = 0
                for p in ex:
                    n = p.numel()
                    g = buf[off:off + n].view_as(p)
                    if p.grad is None:
                        p.grad = g.clone()
                    else:
                        p.grad.copy_(g)
                    off += n
            for wk in self._works:
                wk.wait()                    # current stream waits for the collective; the host does not
        self._works = []
        self._dirty = True
        return self.flat_grad


def allreduce_scalars(values, average=True, group=None):
    """Sum (or mean) of a list of 0-dim tensors across ranks in one collective (losses, regulariser sums)."""
    _, w = world()
    t = torch.stack([v.detach().float().reshape(()) for v in values])
    if w > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        if average:
            t.div_(w)
    return t


def gather_rays(t, n_total, group=None):
    """All-gathers this rank's [r_local, ...] slice into the full [n_total, ...] tensor (shards as in shard_bounds)."""
    rank, w = world()
    if w == 1:
        return t
    sizes = [shard_bounds(n_total, r, w) for r in range(w)]
    mx = max(b - a for a, b in sizes)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:b - a] for p, (a, b) in zip(parts, sizes)], 0)


def render_sharded(render_fn, rays_o, rays_d, keys=("rgb", "depth", "acc"), group=None):
    """Full-view rendering across ranks: `render_fn(rays_o_slice, rays_d_slice) -> dict` runs on this rank's rays
    (e.g. sample_pts + render_star_online); the per-ray outputs named in `keys` are all-gathered."""
    n = rays_o.shape[0]
    ro, rd = shard_rays(rays_o, rays_d)
    out = render_fn(ro, rd)
    return {k: gather_rays(out[k], n, group=group) for k in keys}
