"""world_size-2 gloo tests of the N>1 host path (sharding + final bitstream gather + reductions)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bitswap_b200 import parallel, synthetic
from bitswap_b200.config import preset


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            parts = [parallel.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_shard_by_cost_balances_chains():
    rs = np.random.RandomState(0)
    costs = list(rs.randint(49, 257, size=100))            # 100 images of 49..256 blocks (BASELINE config 5)
    shards = parallel.shard_by_cost(costs, 8)
    assert sorted(i for s in shards for i in s) == list(range(100))
    loads = [sum(costs[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(costs)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O                     # checker only: stands in for the GPU codec on CPU
        cfg = preset("tiny")
        n_streams = 5
        first, count = parallel.shard_range(n_streams, rank, world)
        sd = synthetic.synthetic_state_dict(cfg, seed=50, varied=True)
        zend, zcen = synthetic.synthetic_bins(cfg, seed=0)
        bs = O.BitSwapOracle(cfg, O.ModelOracle(cfg, sd), zend, zcen, coder="c", pmf="c")
        imgs = synthetic.synthetic_images(cfg, n_streams, seed=3)
        states, bits = [], 0.0
        for b in range(first, first + count):              # this rank's shard: stream b codes image b
            w, head = synthetic.initial_words(700 + b, seed=100 + b)
            st = bs.encode_image(O.CState(w, head), imgs[b])
            bits += 32.0 * (st.n - w.size)
            states.append(st)
        words = np.concatenate([s.words[:s.n] for s in states]) if states else np.zeros(0, np.uint32)
        offs = np.zeros(len(states) + 1, dtype=np.int64)
        np.cumsum([s.n for s in states], out=offs[1:])
        heads = np.array([s.head for s in states], dtype=np.uint64)
        gathered = parallel.gather_bitstreams(words, offs, heads)
        # the device-resident form (tensors in, tensors out; CPU tensors here): trimmed payload = words above `base`
        base = np.array([min(s.n, 600 + 10 * i) for i, s in enumerate(states)], dtype=np.int32)
        tw = np.concatenate([s.words[b0:s.n] for s, b0 in zip(states, base)]) if states else np.zeros(0, np.uint32)
        tl = np.array([s.n - b0 for s, b0 in zip(states, base)], dtype=np.int64)
        W, M, counts = parallel.gather_packed(torch.from_numpy(tw.view(np.int32)), torch.from_numpy(tl),
                                              torch.from_numpy(heads.view(np.int64)), torch.from_numpy(base))
        for r in range(world):
            w_r, o_r, h_r = gathered[r]
            pw, po, ph, pb = parallel.unpack_gathered(W, M, counts, r)
            assert np.array_equal(ph.numpy().view(np.uint64), h_r)
            for i in range(len(h_r)):
                full = w_r[o_r[i]:o_r[i + 1]]
                assert np.array_equal(pw[po[i]:po[i + 1]].numpy().view(np.uint32), full[int(pb[i]):])
        total_bits = parallel.reduce_sum(bits)
        slowest = parallel.reduce_max(float(rank + 1))
        ok = True
        if rank == 0:                                       # rank 0 holds every stream: decode them all
            b = 0
            for r in range(world):
                w_r, o_r, h_r = gathered[r]
                for i in range(len(h_r)):
                    st = O.CState(w_r[o_r[i]:o_r[i + 1]], int(h_r[i]))
                    st, x = bs.decode_image(st)
                    w0, head0 = synthetic.initial_words(700 + b, seed=100 + b)
                    ok &= np.array_equal(x, imgs[b].reshape(-1)) and st.n == w0.size and st.head == head0
                    b += 1
            ok &= b == n_streams
        q.put((rank, ok, total_bits, slowest, [len(g[2]) for g in gathered]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_gather_decode():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(r[1] for r in res)
    assert res[0][2] == res[1][2] > 0                      # all-reduced bit count agrees on both ranks
    assert res[0][3] == res[1][3] == 2.0                   # max over ranks
    assert res[0][4] == res[1][4] == [3, 2]                # 5 streams -> 3 + 2
