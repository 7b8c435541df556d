"""Arena evaluation sweep (BASELINE.json configs[4]): every synthetic item sequence is played greedily (greedy_a = 0)
with the previous and with the new net, in lockstep, sharded over the ranks; accept iff the new net's mean raw score is
at least the previous net's (CoachBPP.arena_playing semantics, CoachBPP.py:233-291).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/arena_sweep.py \
        --seeds 65536
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resource_packing_self_play_b200 import distributed as D  # noqa: E402
from resource_packing_self_play_b200.coach import CoachBPP  # noqa: E402
from resource_packing_self_play_b200.game import BinPackingGame, ItemsGenerator  # noqa: E402
from resource_packing_self_play_b200.nnet import NNetWrapper  # noqa: E402
from resource_packing_self_play_b200.utils import dotdict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seeds", type=int, default=65536, help="item sequences over ALL ranks")
ap.add_argument("--sims", type=int, default=200)
ap.add_argument("--chunk", type=int, default=8192, help="games in lockstep per rank and pass")
a = ap.parse_args()
rank, ws, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if ws > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, N = 15, 15, 10
args = dotdict(numMCTSSims=a.sims, cpuct=1, alpha=0.75, seed=100, num_items=N, num_bins=1, numItems=N, arenaCompare=10,
               checkpoint="/tmp/_bpp_arena", epochs=1, batch_size=64)
g = BinPackingGame(W, H, N, 1)
gen = ItemsGenerator(W, H, N)
torch.manual_seed(1)
pnet = NNetWrapper(g, args, max_batch=a.chunk)   # "previous" net
torch.manual_seed(2)
nnet = NNetWrapper(g, args, max_batch=a.chunk)   # "new" net
coach = CoachBPP(g, nnet, gen.items_generator(args.seed), W * H, gen, args, saved_rewards_list=[1.0] * 100)
lo, hi = D.shard_range(a.seeds, rank, ws)
seeds = 1000 + np.arange(lo, hi)
heights = np.array([np.random.RandomState(77000 + int(b)).randint(2, 16) for b in (np.arange(lo, hi) // 20)])
torch.cuda.synchronize()
if ws > 1:
    dist.barrier()
t0 = time.perf_counter()
p_all, n_all = [], []
for c0 in range(0, hi - lo, a.chunk):
    sl = slice(c0, min(c0 + a.chunk, hi - lo))
    coach.gen.bin_height = H
    p, n_, _ = coach.arena_sweep(pnet, nnet, seeds[sl], bin_heights=heights[sl], seed=5)
    p_all.append(p)
    n_all.append(n_)
torch.cuda.synchronize()
dev = torch.device("cuda", local)
ps = D.all_gather_variable(torch.from_numpy(np.concatenate(p_all)).to(dev)).cpu().numpy()
ns = D.all_gather_variable(torch.from_numpy(np.concatenate(n_all)).to(dev)).cpu().numpy()
dt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"n_gpus": ws, "sequences": int(len(ps)), "games_played": int(2 * len(ps)), "seconds": dt,
                      "episodes_per_sec": 2 * len(ps) / dt, "mean_score_prev": float(ps.mean()),
                      "mean_score_new": float(ns.mean()), "accept_new": int(ns.mean() >= ps.mean())}))
if ws > 1:
    dist.destroy_process_group()
