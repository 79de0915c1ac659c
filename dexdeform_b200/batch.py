"""Environment-batch data parallelism: independent environments are split contiguously over ranks, one
process per GPU; the only exchange per optimisation iteration is one all-reduce of a packed fp32 buffer
``[sum of losses | action (or pose) gradients]``.  A single scene is never spatially decomposed.

Works with any ``torch.distributed`` backend: ``nccl`` over NVLink on the GPU box, ``gloo`` in the CPU tests."""
import torch
import torch.distributed as dist


def partition_envs(n_envs, world_size, rank):
    """Contiguous split: returns (first_env, n_local).  The first ``n_envs % world_size`` ranks get one extra."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(int(n_envs), int(world_size))
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def pack(loss, grads):
    """[loss | g0.flatten() | g1.flatten() ...] as one contiguous fp32 tensor on the gradients' device."""
    grads = [g for g in grads]
    dev = grads[0].device if grads else (loss.device if torch.is_tensor(loss) else torch.device("cpu"))
    loss_t = (loss.detach() if torch.is_tensor(loss) else torch.tensor(float(loss))).to(dev, torch.float32).reshape(1)
    return torch.cat([loss_t] + [g.detach().to(torch.float32).reshape(-1) for g in grads])


def unpack(buf, shapes):
    out, o = [], 1
    for shp in shapes:
        n = 1
        for s in shp:
            n *= int(s)
        out.append(buf[o:o + n].reshape(shp))
        o += n
    return buf[0], out


def allreduce_loss_and_grads(loss, grads, group=None, average=False):
    """Sum (or average) the loss and the gradients over all ranks with ONE collective.  Returns (loss, [grads])."""
    shapes = [tuple(g.shape) for g in grads]
    buf = pack(loss, grads)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        if average:
            buf /= dist.get_world_size(group)
    return unpack(buf, shapes)


def gather_scores(local_scores, n_envs, group=None):
    """Demonstration scoring (config C): every rank ends up with the (n_envs,) vector of per-environment scores."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_scores
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    full = torch.zeros(n_envs, dtype=local_scores.dtype, device=local_scores.device)
    start, count = partition_envs(n_envs, world, rank)
    full[start:start + count] = local_scores
    dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return full
