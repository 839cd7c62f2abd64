"""Data parallelism of the attack: stereo pairs are independent, so
pair i goes to rank i mod world with NO data-path collective; the only exchanges
are a final all_gather of fixed-size per-pair statistics and, for the universal
patch (config 4), an all_reduce(sum) of the 71 KB clipped patch step.  The
reference has no working multi-GPU path (nn.DataParallel no-op,
attack/DSGN/pgd_attack.py:138; batch-0-only loop :196-207)."""
import os

import torch
import torch.distributed as dist

STAT_FIELDS = ("pair", "loss_0", "loss_K", "linf", "l2", "frac_changed", "n_det_clean", "n_det_adv")


def init(backend=None, device=None):
    """Initialise torch.distributed from the torchrun environment (no-op for 1 process)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend=backend, **kw)
    return rank(), world_size()


def rank():
    return dist.get_rank() if dist.is_initialized() else 0


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def shard_pairs(num_pairs, rank_=None, world=None):
    """Pair indices owned by this rank: i with i mod world == rank."""
    rank_ = rank() if rank_ is None else rank_
    world = world_size() if world is None else world
    return list(range(rank_, num_pairs, world))


def pair_stats(pair, losses, adv01, clean01, n_det_clean=float('nan'), n_det_adv=float('nan')):
    """Fixed-size statistics row for one attacked pair (device tensor [len(STAT_FIELDS)])."""
    delta = (adv01 - clean01).flatten()
    row = torch.stack([
        torch.tensor(float(pair), device=delta.device), losses[0].float(), losses[-1].float(),
        delta.abs().max(), delta.norm(), (delta != 0).float().mean(),
        torch.tensor(float(n_det_clean), device=delta.device),
        torch.tensor(float(n_det_adv), device=delta.device)])
    return row


def gather_stats(rows, num_pairs):
    """all_gather per-pair statistic rows from every rank -> [num_pairs, F] sorted by pair
    index (identical on every rank).  Ranks may own different numbers of pairs: rows are
    padded to ceil(num_pairs / world) with pair = -1."""
    world = world_size()
    f = len(STAT_FIELDS)
    per = (num_pairs + world - 1) // world
    device = rows[0].device if rows else torch.device('cuda' if torch.cuda.is_available() else 'cpu')
    local = torch.full((per, f), -1.0, device=device)
    if rows:
        local[:len(rows)] = torch.stack(rows)
    if world > 1:
        out = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(out, local)
        allrows = torch.cat(out, 0)
    else:
        allrows = local
    allrows = allrows[allrows[:, 0] >= 0]
    return allrows[torch.argsort(allrows[:, 0])]


def allreduce_patch_delta(delta):
    """Config 4: sum of the clipped patch steps of all ranks (synchronous mini-batch of
    ``world`` images; world == 1 reproduces the reference's sequential update exactly)."""
    if world_size() > 1:
        dist.all_reduce(delta, op=dist.ReduceOp.SUM)
    return delta


def barrier():
    if dist.is_initialized():
        dist.barrier()
