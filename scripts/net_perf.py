"""Throughput of the leaf evaluator kernels (evals/s, TFLOP/s) on synthetic states.  GPU only."""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resource_packing_self_play_b200.engine import pack_states  # noqa: E402
from resource_packing_self_play_b200.game import ItemsGenerator  # noqa: E402
from resource_packing_self_play_b200.nnet import NNetWrapper  # noqa: E402
from resource_packing_self_play_b200.utils import dotdict  # noqa: E402

FLOPS = {(15, 15, 10): 4394592, (20, 20, 10): 7249920}


class G:
    def __init__(self, W, H, N):
        self.W, self.H, self.N = W, H, N

    def getBoardSize(self):
        return (self.H, self.W)

    def getActionSize(self):
        return self.W * self.N


def run(W, H, N, B, modes=("bf16", "bf16_simt", "fp32"), iters=20):
    torch.manual_seed(0)
    net = NNetWrapper(G(W, H, N), dotdict(num_items=N, num_bins=1), max_batch=B, precision="bf16")
    rng = np.random.RandomState(0)
    recs = np.zeros((B, 32), dtype=np.uint32)
    recs[:, :H] = rng.randint(0, 1 << W, size=(B, H)) & rng.randint(0, 1 << W, size=(B, H))
    recs[:, 28] = rng.randint(1, 1 << N, size=B)
    items = ItemsGenerator(W, H, N).items_batch(np.arange(B) % 97, None)
    dev = net.device
    recs_t = torch.from_numpy(recs.view(np.int32)).to(dev)
    items_t = torch.from_numpy(items).to(dev)
    pol = torch.empty((B, W * N), dtype=torch.float32, device=dev)
    val = torch.empty(B, dtype=torch.float32, device=dev)
    out = {}
    for m in modes:
        net.dnet.set_precision(m)
        for _ in range(3):
            net.dnet.forward(recs_t, items_t, policy_out=pol, value_out=val)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            net.dnet.forward(recs_t, items_t, policy_out=pol, value_out=val)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out[m] = {"ms": ms, "evals_per_s": B / ms * 1e3, "tflops": B * FLOPS[(W, H, N)] / ms / 1e9}
        if m in ("bf16", "bf16x3"):
            out[m]["cta0_cycles"] = net.dnet.profile()
            if os.environ.get("BPP_NO_ROLES") is None:
                out[m]["cta0_cycles_per_role"] = net.dnet.profile_roles()
    return out


if __name__ == "__main__":
    modes = tuple(sys.argv[1].split(",")) if len(sys.argv) > 1 else ("bf16", "bf16x3", "bf16_simt", "fp32")
    for (W, H, N) in [(15, 15, 10), (20, 20, 10)]:
        for B in (2048, 2800, 8192):
            print(json.dumps({"cfg": [W, H, N], "B": B, **run(W, H, N, B, modes)}))
