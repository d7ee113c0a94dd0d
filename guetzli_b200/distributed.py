"""torch.distributed plumbing for the sharded (one image per rank) mode.

The hot path has no data-path collective: images are independent (BASELINE
configs[4]; reference idiom: one process per image, tests/golden_test.sh:25).
The process group is only used for the barrier around the timed region and for
the max-over-ranks time / gather of per-rank results."""
import os


def setup(backend="nccl"):
    """-> (rank, world, local_rank, dist module or None)"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world <= 1:
        return 0, 1, 0, None
    import torch
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend=backend)
    return rank, world, local, dist


def image_seed(base_seed, rank):
    """Weak scaling: rank r of every step encodes image number base_seed + r."""
    return base_seed + rank


def barrier(dist, cuda=True):
    if dist is not None:
        dist.barrier()
    if cuda:
        import torch
        torch.cuda.synchronize()


def max_over_ranks(dist, seconds, device=None):
    if dist is None:
        return float(seconds)
    import torch
    t = torch.tensor([seconds], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_strings(dist, s, world):
    """All ranks' strings on every rank (used for per-rank output hashes)."""
    if dist is None:
        return [s]
    out = [None] * world
    dist.all_gather_object(out, s)
    return out


def throughput_mpix(world, steps, pixels_per_image, seconds):
    """Whole-job MPix/s: all ranks' images over the slowest rank's time."""
    return world * steps * pixels_per_image / seconds / 1e6
