"""Multi-GPU host logic: the problem axis (clips x inits) shards embarrassingly, whole clips per rank (the
temporal smoothness term couples the frames of a clip, /root/reference/homan/lossutils.py:31-32, so a clip is
never split). There is no collective inside the iteration; the only exchange is the final gather of every
clip's best initialisation. Works over NCCL (CUDA tensors) and gloo (CPU tensors, tests)."""
import torch
import torch.distributed as dist


def shard_clips(num_clips, rank, world_size):
    """Contiguous block of clip ids owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(num_clips, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def local_best(total, clips_local, inits):
    """total [clips_local * inits] final losses (clip-major) -> (best_init [C], best_loss [C])."""
    t = total.view(clips_local, inits)
    best_loss, best_init = t.min(dim=1)
    return best_init.to(torch.int64), best_loss


def gather_best(clip_ids, best_init, best_loss, num_clips, payload=None):
    """All ranks end up with the global tables best_init [num_clips], best_loss [num_clips] (and, when given,
    payload [num_clips, D], e.g. the winner's fitted parameters). One all_gather per table; ranks may own
    different numbers of clips (padded to the maximum)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    dev = best_loss.device
    if world == 1:
        out_i = torch.zeros(num_clips, dtype=torch.int64, device=dev)
        out_l = torch.full((num_clips,), float("inf"), device=dev)
        ids = torch.as_tensor(clip_ids, dtype=torch.int64, device=dev)
        out_i[ids], out_l[ids] = best_init, best_loss
        out_p = None
        if payload is not None:
            out_p = torch.zeros(num_clips, payload.shape[1], device=dev)
            out_p[ids] = payload
        return out_i, out_l, out_p
    cap = (num_clips + world - 1) // world
    D = 0 if payload is None else payload.shape[1]
    rec = torch.full((cap, 3 + D), -1.0, device=dev)
    n = len(clip_ids)
    rec[:n, 0] = torch.as_tensor(clip_ids, dtype=torch.float32, device=dev)
    rec[:n, 1] = best_init.float()
    rec[:n, 2] = best_loss
    if D:
        rec[:n, 3:] = payload
    parts = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(parts, rec)
    allrec = torch.cat(parts)
    allrec = allrec[allrec[:, 0] >= 0]
    ids = allrec[:, 0].long()
    out_i = torch.zeros(num_clips, dtype=torch.int64, device=dev)
    out_l = torch.full((num_clips,), float("inf"), device=dev)
    out_i[ids], out_l[ids] = allrec[:, 1].long(), allrec[:, 2]
    out_p = None
    if D:
        out_p = torch.zeros(num_clips, D, device=dev)
        out_p[ids] = allrec[:, 3:]
    return out_i, out_l, out_p
