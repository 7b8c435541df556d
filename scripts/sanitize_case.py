"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): fused stub search, lockstep search with the
tensor-core evaluator, env ops.  Run: compute-sanitizer --tool memcheck python scripts/sanitize_case.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resource_packing_self_play_b200 import _lib  # noqa: E402
from resource_packing_self_play_b200.engine import EnvOps, SearchEngine  # noqa: E402
from resource_packing_self_play_b200.game import BinPackingGame, ItemsGenerator  # noqa: E402
from resource_packing_self_play_b200.mcts import BatchedMCTS  # noqa: E402
from resource_packing_self_play_b200.nnet import NNetWrapper  # noqa: E402
from resource_packing_self_play_b200.utils import dotdict  # noqa: E402

W, H, N = 15, 15, 10
G = 24
items = ItemsGenerator(W, H, N).items_batch(np.arange(G) + 5, np.arange(G) % 14 + 2)
area = (W * (np.arange(G) % 14 + 2)).astype(np.int32)
eng = SearchEngine(W, H, N, G, 24, 1.0, edge_cap=60000)
eng.reset(items, area, np.full(G, 0.7001))
counts, actions = eng.play_stub("H", _lib.CHOOSE_SAMPLE, seed=3)
eng.check()
ops = EnvOps(W, H, N)
assert not bool(ops.valid_moves(eng.roots(), items).any())
g = BinPackingGame(W, H, N, 1)
args = dotdict(numMCTSSims=8, cpuct=1, alpha=0.75, num_items=N, num_bins=1)
net = NNetWrapper(g, args, max_batch=G)
bm = BatchedMCTS(g, net, args, G)
bm.reset(items, area, [])
for mv in range(3):
    c = bm.search(chunk=2)
    bm.eng.advance(bm.eng.choose(0))
bm.eng.check()
torch.cuda.synchronize()
print("sanitize case ok", int(counts.sum()), int(c.sum()))
