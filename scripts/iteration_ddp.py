"""Full CoachBPP iteration on N GPUs (BASELINE.json configs[3]): batched self-play sharded over the ranks, NCCL
all-gather of scores/examples, data-parallel learner with one flat gradient all-reduce per step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/iteration_ddp.py \
        --games 8192 --iters 2 --sims 200 --batch 512 --epochs 2
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resource_packing_self_play_b200.coach import CoachBPP  # noqa: E402
from resource_packing_self_play_b200.game import BinPackingGame, ItemsGenerator  # noqa: E402
from resource_packing_self_play_b200.nnet import NNetWrapper  # noqa: E402
from resource_packing_self_play_b200.utils import dotdict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=4096, help="games per iteration over ALL ranks")
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--sims", type=int, default=200)
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--epochs", type=int, default=2)
a = ap.parse_args()
rank, ws, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if ws > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, N = 15, 15, 10
args = dotdict(numIters=a.iters, numEps=20, iterStepThreshold=50, maxlenOfQueue=200000, numMCTSSims=a.sims, arenaCompare=10,
               cpuct=1, alpha=0.75, seed=100, numScoresForRank=100, numItems=N, numBins=1, binH_min=2, binH=15,
               epochs=a.epochs, batch_size=a.batch, cuda=True, num_items=N, num_bins=1, checkpoint="/tmp/_bpp_iter",
               numItersForTrainExamplesHistory=50)
g = BinPackingGame(W, H, N, 1, device=local)
gen = ItemsGenerator(W, H, N)
torch.manual_seed(0)
net = NNetWrapper(g, args, max_batch=max(1, a.games // ws + 1), device=local)
coach = CoachBPP(g, net, gen.items_generator(args.seed), W * H, gen, args)
log = coach.learn_batched(a.games, num_iters=a.iters, checkpoint=(rank == 0))
# all ranks must end with identical weights and rewards buffers
flat = torch.cat([p.detach().reshape(-1) for p in net.nnet.parameters()])
chk = torch.tensor([float(flat.double().sum()), float(np.sum(coach.rewards_list))], dtype=torch.float64, device=flat.device)
same = True
if ws > 1:
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo, hi))
if rank == 0:
    last = log[-1]
    print(json.dumps({"n_gpus": ws, "games_per_iter": a.games, "sims": a.sims, "ranks_in_sync": same,
                      "episodes_per_sec_selfplay": a.games / last["t_selfplay"], "iteration_s": sum(
                          last[k] for k in ("t_setup", "t_selfplay", "t_gather", "t_train")), "log": log}))
if ws > 1:
    dist.destroy_process_group()
