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
    if not params:
        return None
    _, w = world()
    runs = _grad_runs(params)
    if runs is not None:
        if w > 1:
            for r in runs:
                dist.all_reduce(r, op=dist.ReduceOp.SUM, group=group)
                if average:
                    r.div_(w)
        return runs
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params])
    if w > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(w)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return flat


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
