"""Multi-GPU sharding of the embedding path: utterances are independent, so the batch is split into
contiguous index ranges, one per rank (one process per GPU), weights are replicated, and the only collective is one
all-gather of the per-rank ``[B/R, embd]`` fp32 outputs so that every rank ends up with all embeddings in the original
order.  Ragged batches keep the reference's single-batch semantics (predict.py:244-258: pad to the longest item of the
WHOLE batch, T and the CMN mean follow that padding) by padding every shard to the global ``Lmax`` on the host.

One front-end is NOT shard-invariant: torchaudio's MFCC clamps to (max over the whole call) - top_db, so a sharded call
would clamp against the shard's maximum instead of the batch's.  Matching the single-process result needs one extra
all-reduce(MAX) of that scalar between the mel stage and the DCT; until that exists, shard MFCC configurations only when
the per-shard clamp is acceptable (Fbank / MelSpectrogram / Spectrogram have no cross-utterance term).

torch.distributed (NCCL on GPUs, gloo in the CPU tests) is plumbing only; the compute is ``embed_fn``."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced split: the first ``n % world`` ranks get one extra item."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def pad_to_global_max(waves):
    """list of 1-D float32 arrays -> ([B, Lmax] zero padded, ratio[B] = len / Lmax) (predict.py:244-255)."""
    lmax = max(w.shape[0] for w in waves)
    x = np.zeros((len(waves), lmax), dtype=np.float32)
    ratio = np.empty(len(waves), dtype=np.float32)
    for i, w in enumerate(waves):
        x[i, :w.shape[0]] = w
        ratio[i] = w.shape[0] / lmax
    return x, ratio


def embed_sharded(embed_fn, waves, embd_dim, device, group=None):
    """Every rank passes the same ``waves`` list; rank r embeds its shard with ``embed_fn(x[B_r, Lmax], ratio[B_r]) ->
    tensor [B_r, embd_dim] on `device```; returns the full ``[B, embd_dim]`` tensor on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    x, ratio = pad_to_global_max(waves)
    n = x.shape[0]
    lo, hi = shard_range(n, rank, world)
    per = -(-n // world)                                     # equal-sized slots so one all_gather_into_tensor suffices
    local = torch.zeros(per, embd_dim, dtype=torch.float32, device=device)
    if hi > lo:
        local[:hi - lo] = embed_fn(x[lo:hi], ratio[lo:hi])
    if world == 1:
        return local[:n]
    gathered = torch.empty(world * per, embd_dim, dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(gathered, local, group=group)
    parts = []
    for r in range(world):
        a, b = shard_range(n, r, world)
        parts.append(gathered[r * per: r * per + (b - a)])
    return torch.cat(parts, dim=0)
