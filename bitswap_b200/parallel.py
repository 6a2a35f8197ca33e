"""Multi-GPU plumbing: one process per GPU (torch.distributed), streams sharded by rank.

The reference compresses on a single GPU and uses no communication at all in the codec
(SURVEY.md 2b/8e); its only collectives are Horovod calls in training.  ANS streams are
independent chains (each reference "experiment" / each cropped image is its own chain with its
own initial state, cifar_compress.py:157, imagenetcrop_compress.py:122), so the data path needs
no collective: every rank codes its own shard.  NCCL (over NVLink/NVSwitch) is used only at the
end, to gather the final bitstreams and to reduce the bit count for the aggregate rate.

All functions work on any torch.distributed backend: `nccl` with CUDA tensors on the GPU box,
`gloo` with CPU tensors in the tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block partition of range(n_items): returns (first, count) for `rank`."""
    base, extra = divmod(n_items, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def shard_by_cost(costs, world: int):
    """Greedy longest-first bin packing (chains of different lengths, e.g. images tiled into a varying
    number of 32x32 blocks, imagenetcrop_compress.py:127-210).  Returns a list of index lists per rank."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (load[j], j))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(o) for o in out]


def gather_bitstreams(words, offsets, heads, device=None, group=None):
    """All-gathers packed bitstreams.  Each rank passes its packed export (words uint32 [n_r],
    offsets int64 [s_r+1], heads uint64 [s_r]); every rank receives a list (one entry per rank) of
    (words, offsets, heads) numpy triples.  Three collectives: lengths, padded words, heads."""
    world = dist.get_world_size(group)
    device = device or torch.device("cpu")
    words = np.ascontiguousarray(words, dtype=np.uint32)
    lens = np.diff(np.asarray(offsets, dtype=np.int64)).astype(np.int64)
    meta = torch.tensor([words.size, lens.size], dtype=torch.int64, device=device)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    max_w = max(int(m[0]) for m in metas)
    max_s = max(int(m[1]) for m in metas)

    def padded(a, n, dtype):
        t = torch.zeros(max(n, 1), dtype=dtype, device=device)
        if a.size:
            t[:a.size] = torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a.view(np.int64)).to(device)
        return t

    wt = padded(words, max_w, torch.int32)
    lt = padded(lens, max_s, torch.int64)
    ht = padded(np.ascontiguousarray(heads, dtype=np.uint64), max_s, torch.int64)
    W = [torch.empty_like(wt) for _ in range(world)]
    Ls = [torch.empty_like(lt) for _ in range(world)]
    H = [torch.empty_like(ht) for _ in range(world)]
    dist.all_gather(W, wt, group=group)
    dist.all_gather(Ls, lt, group=group)
    dist.all_gather(H, ht, group=group)
    out = []
    for r in range(world):
        nw, ns = int(metas[r][0]), int(metas[r][1])
        ln = Ls[r][:ns].cpu().numpy()
        offs = np.zeros(ns + 1, dtype=np.int64)
        np.cumsum(ln, out=offs[1:])
        out.append((W[r][:nw].cpu().numpy().view(np.uint32), offs, H[r][:ns].cpu().numpy().view(np.uint64)))
    return out


def reduce_sum(value: float, device=None, group=None) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device or torch.device("cpu"))
    dist.all_reduce(t, group=group)
    return float(t.item())


def reduce_max(value: float, device=None, group=None) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device or torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
