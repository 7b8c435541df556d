"""Learner step time: the hand-written CUDA step (bpp_learner_grad + bpp_learner_adam) against the torch-autograd step
(cuDNN, CUDA-graph captured), same minibatch size.  python scripts/learner_perf.py [--batch 64 512] [--size 15]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resource_packing_self_play_b200.engine import EnvOps  # noqa: E402
from resource_packing_self_play_b200.game import BinPackingGame, ItemsGenerator  # noqa: E402
from resource_packing_self_play_b200.nnet import NNetWrapper  # noqa: E402
from resource_packing_self_play_b200.utils import dotdict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, nargs="+", default=[64, 512])
ap.add_argument("--size", type=int, default=15)
ap.add_argument("--steps", type=int, default=200)
a = ap.parse_args()
W = H = a.size
N = 10
dev = torch.device("cuda")
M = 16384
gen = ItemsGenerator(W, H, N)
items = torch.from_numpy(gen.items_batch(range(M))).to(dev)
ops = EnvOps(W, H, N)
recs = np.zeros((M, 32), dtype=np.uint32)
recs[:, 28] = (1 << N) - 1
recs = torch.from_numpy(recs.view(np.int32)).to(dev)
for k in range(4):
    valid = ops.valid_moves(recs, items).float()
    act = torch.multinomial(valid + 1e-6, 1)[:, 0].int()
    recs = torch.where((torch.rand(M, device=dev) < 0.7)[:, None], ops.next_state(recs, items, act), recs)
pis = torch.softmax(torch.randn(M, W * N, device=dev), 1)
vs = (torch.randint(0, 2, (M,), device=dev) * 2 - 1).float()
flop = {15: 4394592, 20: 7249920}.get(a.size, 0) * 3
g = BinPackingGame(W, H, N, 1)
for bs in a.batch:
    row = {"cfg": [W, H, N], "batch": bs}
    for learner in ("cuda", "torch"):
        args = dotdict(num_items=N, num_bins=1, epochs=1, batch_size=bs, cuda=True)
        torch.manual_seed(0)
        net = NNetWrapper(g, args, max_batch=64)
        net.train_compact(recs, items, pis, vs, ops, steps_per_epoch=20, seed=1, learner=learner)  # warm-up / JIT
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        l = net.train_compact(recs, items, pis, vs, ops, steps_per_epoch=a.steps, seed=1, learner=learner)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        row[learner] = {"ms_per_step": 1e3 * dt / a.steps, "samples_per_s": bs * a.steps / dt,
                        "tflops": flop * bs * a.steps / dt / 1e12, "loss": l}
    row["speedup"] = row["torch"]["ms_per_step"] / row["cuda"]["ms_per_step"]
    print(json.dumps(row))
