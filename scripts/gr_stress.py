"""Stress / determinism check of the grid-row stage kernels (GPU): many batch sizes, repeated launches on the same inputs must
reproduce their outputs bit for bit (a race in the per-tile mbarrier protocol would show up as a rare difference), and every
result must stay within tolerance of the previous tensor-core path.   usage: gr_stress.py [rounds]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gr_check import data, make  # noqa: E402


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = np.random.RandomState(0)
    bad = 0
    for (W, H, N) in [(15, 15, 10), (20, 20, 10)]:
        Bmax = 9000
        recs, items = data(W, H, N, Bmax, seed=3)
        for mode in ("bf16", "bf16x3"):
            new = make(W, H, N, Bmax, True, precision=mode, scale=25.0)
            old = make(W, H, N, Bmax, False, precision=mode, scale=25.0)
            assert new.dnet.grid_row() and not old.dnet.grid_row()
            for it in range(rounds):
                B = int(rng.choice([1, 2, 7, 8, 9, 63, 148, 149, 295, 296, 297, 1200, 2800, 3139, 4096, 8192, 9000])) if it % 2 else int(rng.randint(1, Bmax + 1))
                r, i = recs[:B].contiguous(), items[:B].contiguous()
                p0, v0 = new.dnet.forward(r, i)
                p0, v0 = p0.clone(), v0.clone()
                po, vo = old.dnet.forward(r, i)
                d_old = float((p0 - po).abs().max())
                same = True
                for rep in range(4):
                    p1, v1 = new.dnet.forward(r, i)
                    same = same and bool(torch.equal(p1, p0)) and bool(torch.equal(v1, v0))
                torch.cuda.synchronize()
                ok = same and bool(torch.isfinite(p0).all()) and d_old < 1e-3
                if not ok:
                    bad += 1
                    print(f"FAIL {W}x{H} {mode} B={B}: repeatable {same} |new-old| {d_old:.2e}", flush=True)
            print(f"{W}x{H} {mode}: {rounds} batch sizes x 5 launches done", flush=True)
    print("STRESS", "PASS" if bad == 0 else f"FAIL ({bad})")


if __name__ == "__main__":
    main()
