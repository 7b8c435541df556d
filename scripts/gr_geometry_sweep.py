"""Grid-row kernels over odd board geometries (GPU): every (W, H, N) is checked against the CUDA-core kernels (bf16 vs
bf16_simt, split-bf16 vs fp32) on ragged batches.   usage: gr_geometry_sweep.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resource_packing_self_play_b200.nnet import NNetWrapper  # noqa: E402
from resource_packing_self_play_b200.utils import dotdict  # noqa: E402


class G:
    def __init__(self, W, H, N):
        self.W, self.H, self.N = W, H, N

    def getBoardSize(self):
        return (self.H, self.W)

    def getActionSize(self):
        return self.W * self.N


def main():
    rng = np.random.RandomState(0)
    geos = [(3, 2, 2), (2, 28, 3), (32, 3, 4), (16, 16, 8), (17, 17, 6), (31, 1, 2), (5, 5, 16), (1, 9, 3), (24, 24, 10),
            (15, 15, 10), (12, 28, 7), (7, 13, 1), (25, 10, 10), (8, 8, 15)]
    bad = 0
    for (W, H, N) in geos:
        for B in (1, 37, 600):
            recs = np.zeros((B, 32), dtype=np.uint32)
            recs[:, :H] = rng.randint(0, 1 << min(W, 30), size=(B, H))
            recs[:, 28] = rng.randint(1, 1 << N, size=B)
            items = np.stack([rng.randint(1, W + 1, size=(B, N)), rng.randint(1, H + 1, size=(B, N))], axis=2).astype(np.int32)
            r_t, i_t = torch.from_numpy(recs.view(np.int32)).cuda(), torch.from_numpy(items).cuda()
            torch.manual_seed(1)
            net = NNetWrapper(G(W, H, N), dotdict(num_items=N, num_bins=1, cuda=True, epochs=1, batch_size=8), max_batch=B,
                              precision="bf16")
            with torch.no_grad():
                net.nnet.logits_fc.weight.mul_(10.0)
            net.sync_weights()
            out = {}
            for mode in ("bf16", "bf16_simt", "bf16x3", "fp32"):
                net.dnet.set_precision(mode)
                gr = net.dnet.grid_row()
                p, v = net.dnet.forward(r_t, i_t)
                out[mode] = (p.clone(), v.clone(), gr)
            torch.cuda.synchronize()
            d1 = float((out["bf16"][0] - out["bf16_simt"][0]).abs().max())
            d3 = float((out["bf16x3"][0] - out["fp32"][0]).abs().max())
            ok = d1 < 2e-3 and d3 < 2e-4 and bool(torch.isfinite(out["bf16"][0]).all())
            bad += 0 if ok else 1
            if B == 600 or not ok:
                print(f"{'ok  ' if ok else 'FAIL'} {W}x{H} N={N} B={B}: grid-row bf16 {out['bf16'][2]} x3 {out['bf16x3'][2]}  "
                      f"|bf16 - simt| {d1:.1e}  |x3 - fp32| {d3:.1e}", flush=True)
    print("SWEEP", "PASS" if bad == 0 else f"FAIL ({bad})")


if __name__ == "__main__":
    main()
