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


def gather_packed(words_t, lens_t, heads_t, base_t=None, group=None):
    """Device-resident gather of packed bitstreams: every rank contributes the tensors StreamSet.pack_device() returned
    (words int32 [>= sum(lens)], lens int64 [s_r], heads int64 [s_r], base int32 [s_r] or None) and receives
    (W [world, maxtot] int32, meta [world, 3, smax] int64 = lens / heads / base per rank, counts [world, 2] int64 =
    (total words, streams) per rank) -- tensors on the input's device, nothing staged through the host.
    Two collectives (+ one tiny one for the sizes): NCCL all_gather_into_tensor over NVLink on the GPU box, gloo in the
    CPU tests.  With trimmed packing the payload is the PRODUCED words only (about 1-5 KB per image), which is what
    SURVEY.md 8e budgets; the never-borrowed initial words stay home."""
    world = dist.get_world_size(group)
    dev = words_t.device
    s_r = int(lens_t.numel())
    total = lens_t.sum().reshape(1)
    mine = torch.cat([total, torch.tensor([s_r], dtype=torch.int64, device=dev)])
    counts = torch.empty((world, 2), dtype=torch.int64, device=dev)
    _all_gather_into(counts, mine, group)
    cmax = counts.max(dim=0).values.tolist()                  # the only host synchronisation: two integers
    maxtot, smax = max(int(cmax[0]), 1), max(int(cmax[1]), 1)
    send = torch.zeros(maxtot, dtype=torch.int32, device=dev)
    n_mine = int(counts[dist.get_rank(group), 0])
    send[:n_mine] = words_t[:n_mine]
    W = torch.empty((world, maxtot), dtype=torch.int32, device=dev)
    _all_gather_into(W, send, group)
    meta = torch.zeros((3, smax), dtype=torch.int64, device=dev)
    meta[0, :s_r] = lens_t
    meta[1, :s_r] = heads_t
    if base_t is not None:
        meta[2, :s_r] = base_t.to(torch.int64)
    M = torch.empty((world, 3, smax), dtype=torch.int64, device=dev)
    _all_gather_into(M, meta, group)
    return W, M, counts


def _all_gather_into(out, inp, group=None):
    """all_gather_into_tensor where the backend has it (NCCL), list all_gather otherwise (older gloo)."""
    try:
        dist.all_gather_into_tensor(out.view(-1), inp.contiguous().view(-1), group=group)
    except (RuntimeError, NotImplementedError):
        world = dist.get_world_size(group)
        parts = [torch.empty_like(inp) for _ in range(world)]
        dist.all_gather(parts, inp.contiguous(), group=group)
        out.copy_(torch.stack(parts).view_as(out))


def unpack_gathered(W, M, counts, rank):
    """Rank `rank`'s slice of a gather_packed result as (words int32 [total], offsets int64 [s+1], heads int64 [s],
    base int32 [s]) tensors, ready for StreamSet.unpack_device."""
    total, s = int(counts[rank, 0]), int(counts[rank, 1])
    lens = M[rank, 0, :s]
    offs = torch.zeros(s + 1, dtype=torch.int64, device=W.device)
    offs[1:] = torch.cumsum(lens, 0)
    return W[rank, :total].contiguous(), offs, M[rank, 1, :s].contiguous(), M[rank, 2, :s].to(torch.int32).contiguous()


def reduce_sum(value: float, device=None, group=None) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device or torch.device("cpu"))
    dist.all_reduce(t, group=group)
    return float(t.item())


def reduce_max(value: float, device=None, group=None) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device or torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
