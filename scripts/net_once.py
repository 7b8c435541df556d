"""A few forwards of the leaf evaluator on synthetic states (the ncu target of scripts/capture_profiles.sh).
usage: net_once.py W H B mode [iters]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resource_packing_self_play_b200.game import ItemsGenerator  # noqa: E402
from resource_packing_self_play_b200.nnet import NNetWrapper  # noqa: E402
from resource_packing_self_play_b200.utils import dotdict  # noqa: E402

W, H, B, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 4
N = 10


class G:
    def getBoardSize(self):
        return (H, W)

    def getActionSize(self):
        return W * N


torch.manual_seed(0)
net = NNetWrapper(G(), dotdict(num_items=N, num_bins=1), max_batch=B, precision=mode)
rng = np.random.RandomState(0)
recs = np.zeros((B, 32), dtype=np.uint32)
recs[:, :H] = rng.randint(0, 1 << W, size=(B, H)) & rng.randint(0, 1 << W, size=(B, H))
recs[:, 28] = rng.randint(1, 1 << N, size=B)
items = ItemsGenerator(W, H, N).items_batch(np.arange(B) % 97, None)
recs_t = torch.from_numpy(recs.view(np.int32)).cuda()
items_t = torch.from_numpy(items).cuda()
for _ in range(iters):
    pol, val = net.dnet.forward(recs_t, items_t)
torch.cuda.synchronize()
print("ok", float(pol.sum()))
