"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the B200 box, gloo in CPU tests).

Self-play shards trivially: every game owns its search graph, item list and ranked-reward threshold, so ranks play
disjoint seed ranges with NO collective on the data path.  Collectives are used only
  (1) after self-play: all-gather of the episode scores (every rank must update the same rewards_list,
      CoachBPP.py:134-139) and of the compact training examples, and
  (2) in the learner: one flat all-reduce of the 170,583 fp32 gradients per step.
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, world_size):
    """contiguous block [lo, hi) of n units for `rank`; blocks differ in size by at most one"""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_variable(t):
    """all-gather tensors whose first dimension differs between ranks; returns the concatenation in rank order"""
    rank, ws = world()
    if ws == 1:
        return t
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def gather_examples(compact, device):
    """compact: dict of numpy arrays (or device tensors, on_device=True) as returned by
    CoachBPP.executeEpisodesBatched(expand=False) with the game dimension second for roots/counts/actions.  Returns the same dict with the games of all ranks concatenated."""
    rank, ws = world()
    if ws == 1:
        return compact
    out = {}
    for k, v in compact.items():
        if isinstance(v, torch.Tensor):  # device-resident examples: no host hop on either side of the collective
            game_axis = 1 if k in ("roots", "counts", "actions") else 0
            out[k] = all_gather_variable(v.movedim(game_axis, 0).contiguous()).movedim(0, game_axis)
            continue
        a = np.asarray(v)
        game_axis = 1 if k in ("roots", "counts", "actions") else 0
        t = torch.from_numpy(np.ascontiguousarray(np.moveaxis(a, game_axis, 0)).view(
            np.int32 if a.dtype == np.uint32 else a.dtype)).to(device)
        g = all_gather_variable(t).cpu().numpy()
        if a.dtype == np.uint32:
            g = g.view(np.uint32)
        out[k] = np.moveaxis(g, 0, game_axis)
    return out


def allreduce_gradients(module):
    """average the gradients of `module` over all ranks with ONE flat all-reduce (0.68 MB for the reference net)"""
    rank, ws = world()
    if ws == 1:
        return
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(ws)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def broadcast_parameters(module, src=0):
    rank, ws = world()
    if ws == 1:
        return
    for p in module.state_dict().values():
        dist.broadcast(p, src)
