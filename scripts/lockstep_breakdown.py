"""Un-profiled device time of the lockstep step's kernels (CUDA events around every call, eager launches): evaluator
(trunk + heads) vs expand+select, summed over whole episodes of the real15 / real20 workload.  GPU only."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resource_packing_self_play_b200 import _lib  # noqa: E402
from resource_packing_self_play_b200.game import ItemsGenerator  # noqa: E402
from resource_packing_self_play_b200.mcts import BatchedMCTS  # noqa: E402
from resource_packing_self_play_b200.nnet import NNetWrapper  # noqa: E402
from resource_packing_self_play_b200.utils import dotdict  # noqa: E402

Wb = Hb = int(sys.argv[1]) if len(sys.argv) > 1 else 15
G = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
N = 10


class Gm:
    bin_width, bin_height, num_items = Wb, Hb, N

    def getBoardSize(self):
        return (Hb, Wb)

    def getActionSize(self):
        return Wb * N


margs = dotdict(numMCTSSims=200, cpuct=1.0, alpha=0.75, num_items=N, num_bins=1, cuda=True)
torch.manual_seed(0)
net = NNetWrapper(Gm(), margs, max_batch=G, precision="bf16")
bm = BatchedMCTS(Gm(), net, margs, G)
bm.use_graphs = False
gen = ItemsGenerator(Wb, Hb, N)
ev = {"forward": [], "expand_select": [], "select": []}


def wrap(obj, name, key):
    orig = getattr(obj, name)

    def f(*a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(*a, **kw)
        e1.record()
        ev[key].append((e0, e1))
        return r
    setattr(obj, name, f)


wrap(net.dnet, "forward", "forward")
wrap(bm.eng, "expand_select", "expand_select")
wrap(bm.eng, "select", "select")
for k in range(3):
    idx = np.arange(k * G, (k + 1) * G)
    hts = np.array([np.random.RandomState(77000 + int(b)).randint(2, Hb + 1) for b in idx // 20], dtype=np.int32)
    bm.reset(gen.items_batch(1000 + idx, hts), (Wb * hts).astype(np.int32), [])
    if k == 2:
        for v in ev.values():
            v.clear()
        bm.eng.stats(reset=True)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
    for m in range(N):
        bm.search()
        bm.eng.advance(bm.eng.choose(_lib.CHOOSE_SAMPLE, seed=7 + k))
t1.record()
torch.cuda.synchronize()
st = bm.eng.stats()
tot = t0.elapsed_time(t1)
print(f"{Wb}x{Hb}, {G} games, one episode batch, eager launches: {tot:.1f} ms wall, {st['sims'] / tot / 1e3:.2f} M sims/s")
for k, v in ev.items():
    ms = sum(a.elapsed_time(b) for a, b in v)
    print(f"  {k:14s} {len(v):6d} calls {ms:8.1f} ms  {ms / max(1, len(v)) * 1e3:7.1f} us/call  {100 * ms / tot:5.1f} % of wall")
