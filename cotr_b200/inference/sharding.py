"""Multi-GPU data parallelism over independent image pairs.

One process per GPU (torchrun).  Pairs / contexts are independent, so the only exchange is the gather of the
(B,Q,2) fp32 predictions - 8 KB per 1024 queries.  Works with the `nccl` backend on CUDA tensors and with `gloo` on
CPU tensors (used by the CPU tests).
"""
import torch
import torch.distributed as dist


def pair_range(n_pairs, rank, world):
    """Contiguous block of pairs owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_pairs, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_predictions(local_pred, n_pairs, group=None):
    """All-gather per-rank (b_r,Q,2) predictions into the full (n_pairs,Q,2) tensor on every rank, in pair order."""
    if not (dist.is_available() and dist.is_initialized()):
        return local_pred
    world = dist.get_world_size(group)
    q = local_pred.shape[1]
    counts = [pair_range(n_pairs, r, world) for r in range(world)]
    widest = max(e - s for s, e in counts)
    padded = torch.zeros((widest, q, 2), dtype=local_pred.dtype, device=local_pred.device)
    padded[: local_pred.shape[0]] = local_pred
    out = torch.empty((world * widest, q, 2), dtype=local_pred.dtype, device=local_pred.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = [out[r * widest: r * widest + (e - s)] for r, (s, e) in enumerate(counts)]
    return torch.cat(parts, dim=0)


def forward_sharded(model, img, queries, group=None):
    """BASELINE.json configs[3]: every rank holds the full (B,3,256,512) / (B,Q,2) batch description, runs its own
    block of pairs through `model` and receives everybody's predictions."""
    if not (dist.is_available() and dist.is_initialized()):
        return model(img, queries)['pred_corrs']
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    s, e = pair_range(img.shape[0], rank, world)
    if e > s:
        local = model(img[s:e], queries[s:e])['pred_corrs']
    else:
        local = queries.new_zeros((0, queries.shape[1], 2))
    return gather_predictions(local, img.shape[0], group)
