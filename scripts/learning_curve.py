#!/usr/bin/env python
"""Learning-curve sanity of the batched training loop (GPU only).

Runs CoachBPP.learn_batched at the reference's DEFAULT configuration (main_bpp.py:20-47: 15x15 virtual bin, generator
15 x h with h ~ randint(2, 16) per iteration, 10 items, numEps = 20 episodes per iteration, numMCTSSims = 200, cpuct = 1,
alpha = 0.75, 10 epochs of batch 64, Adam lr 1e-3, history of 50 iterations, ranked-reward buffer of 100) for the first
30 iterations and prints the per-iteration mean reward next to the first 30 rows of the reference's archived training
runs (tests/golden/wandb_first30.json, from xw_mcts/wandb/run-*/wandb-history.jsonl).  learn_batched differs from the
reference loop in one documented way: the 20 episodes of an iteration see the rewards list of the iteration's start.

    python scripts/learning_curve.py [runs] [iters] > profiles/r02_learning_curve.json
"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from resource_packing_self_play_b200.coach import CoachBPP  # noqa: E402
from resource_packing_self_play_b200.game import BinPackingGame, ItemsGenerator  # noqa: E402
from resource_packing_self_play_b200.nnet import NNetWrapper  # noqa: E402
from resource_packing_self_play_b200.utils import dotdict  # noqa: E402


def one_run(seed, iters):
    args = dotdict(numIters=iters, numEps=20, numMCTSSims=200, cpuct=1, alpha=0.75, iterStepThreshold=50,
                   numScoresForRank=100, numItersForTrainExamplesHistory=50, maxlenOfQueue=200000, epochs=10,
                   batch_size=64, seed=100, binH_min=2, binH=15, numItems=10, num_items=10, num_bins=1, cuda=True,
                   checkpoint=tempfile.mkdtemp(prefix="bpp_curve_"))
    torch.manual_seed(seed)
    np.random.seed(seed)
    game = BinPackingGame(15, 15, 10, 1)
    net = NNetWrapper(game, args, max_batch=64)
    gen = ItemsGenerator(15, 15, 10)
    coach = CoachBPP(game, net, gen.items_generator(0), 225, gen, args)
    modes, t0 = [], time.perf_counter()
    rows = []
    for i in range(iters):
        out = coach.learn_batched(20, num_iters=1, checkpoint=False)[0]
        rows.append(out)
        modes.append(net.dnet.precision)
    return {"seed": seed, "iter_mean_reward": [r["mean_score"] for r in rows], "loss_pi": [r["loss_pi"] for r in rows],
            "loss_v": [r["loss_v"] for r in rows], "precision_mode": modes, "seconds": time.perf_counter() - t0,
            "rewards_list_tail": coach.rewards_list[-5:]}


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "wandb_first30.json")))
    ours = [one_run(s, iters) for s in range(runs)]

    def summary(x):
        x = np.asarray(x)
        return {"mean_first_%d" % len(x): float(x.mean()), "mean_last_10": float(x[-10:].mean()),
                "mean_first_10": float(x[:10].mean())}
    out = {"config": "reference default (binH 15, binH_min 2, numEps 20, numMCTSSims 200, epochs 10, batch 64)",
           "ours": [{**r, **summary(r["iter_mean_reward"])} for r in ours],
           "reference_runs": {k: {**summary(v["iter_mean_reward"][:iters]), "iter_mean_reward": v["iter_mean_reward"][:iters]}
                              for k, v in ref.items()}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
