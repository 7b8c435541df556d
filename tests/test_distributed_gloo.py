"""CPU: the multi-rank host logic on the gloo backend, world_size 2 (the N>1 path of bench.py / the learner)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from resource_packing_self_play_b200 import distributed as D
    res = {}
    # (1) sharding covers every unit exactly once
    lo, hi = D.shard_range(4099, rank, ws)
    mine = torch.arange(lo, hi)
    allu = D.all_gather_variable(mine)
    res["units_ok"] = bool(torch.equal(allu, torch.arange(4099)))
    # (2) example gather: per-rank game counts differ
    G = 3 + rank
    compact = {"roots": np.full((10, G, 32), rank, dtype=np.uint32), "counts": np.full((10, G, 150), rank, np.int32),
               "actions": np.full((10, G), rank, np.int32), "moves": np.full(G, 5 + rank, np.int32),
               "r": np.full(G, 1 - 2 * rank, np.int32), "items": np.full((G, 10, 2), rank, np.int32)}
    g = D.gather_examples(compact, torch.device("cpu"))
    res["gather_ok"] = (g["roots"].shape == (10, 7, 32) and g["roots"].dtype == np.uint32 and
                        g["moves"].tolist() == [5, 5, 5, 6, 6, 6, 6] and g["counts"][:, 3:].min() == 1 and
                        g["counts"][:, :3].max() == 0 and g["r"].tolist() == [1, 1, 1, -1, -1, -1, -1])
    # (2b) the same with torch tensors (the device-resident path of learn_batched: no numpy hop around the collective)
    tcompact = {k: torch.from_numpy(v.view(np.int32) if v.dtype == np.uint32 else v) for k, v in compact.items()}
    tg = D.gather_examples(tcompact, torch.device("cpu"))
    res["gather_tensor_ok"] = all(isinstance(v, torch.Tensor) for v in tg.values()) and all(
        np.array_equal(tg[k].numpy(), g[k].view(np.int32) if g[k].dtype == np.uint32 else g[k]) for k in g)
    # (3) gradient all-reduce == mean of the per-rank gradients; parameters stay in sync
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    D.broadcast_parameters(net)
    x = torch.full((5, 8), float(rank + 1))
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    D.allreduce_gradients(net)
    gathered = [D.all_gather_variable(l.reshape(1, -1)) for l in local]
    res["grad_ok"] = all(torch.allclose(p.grad.reshape(-1), gl.mean(dim=0)) for p, gl in zip(net.parameters(), gathered))
    # (4) the bench's instance partition: ranks get disjoint consecutive seed blocks per step
    import bench
    seeds = [bench.workload(((k * ws) + rank) * 16, 16)[0] for k in range(3)]
    flat = torch.from_numpy(np.concatenate(seeds))
    every = D.all_gather_variable(flat)
    res["seeds_ok"] = len(set(every.tolist())) == 3 * 16 * ws
    out[rank] = res
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ws = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(ws, _free_port(), out), nprocs=ws, join=True)
    assert len(out) == ws
    for rank in range(ws):
        assert all(out[rank].values()), (rank, dict(out[rank]))


def test_shard_range_edges():
    from resource_packing_self_play_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 4096, 65536):
        for ws in (1, 2, 3, 8):
            blocks = [shard_range(n, r, ws) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
