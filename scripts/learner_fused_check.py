"""Development check of the fused stage kernels of the learner (k_lr_stage_fwd / k_lr_stage_bwd, one launch per
ConvSequence and direction) against the per-layer kernels (BPP_LEARNER_FUSED=0): gradients, losses, outputs; then the time
of bpp_learner_grad alone (CUDA events) for both.   python scripts/learner_fused_check.py [perf]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_learner import _examples, _module  # noqa: E402

from resource_packing_self_play_b200.nnet import DeviceLearner  # noqa: E402


def learner(W, H, N, B, fused, m):
    os.environ["BPP_LEARNER_FUSED"] = "1" if fused else "0"
    L = DeviceLearner(W, H, N, max_batch=B)
    L.load_state_dict(m.state_dict())
    return L


def check(W, H, N, B):
    ops, recs, items, pis, vs = _examples(B, W, H, N, seed=7)
    m = _module(W, H, N, scale=1.5)
    out = []
    for fused in (True, False):
        L = learner(W, H, N, B, fused, m)
        logp = torch.empty((B, W * N), device="cuda")
        v = torch.empty(B, device="cuda")
        losses = L.grad(recs, items, pis, vs, logp_out=logp, v_out=v).clone()
        out.append((losses, L.grads.clone(), logp, v))
        L.close()
    (l1, g1, p1, v1), (l0, g0, p0, v0) = out
    dg = float((g1 - g0).abs().max()) / max(1e-30, float(g0.abs().max()))
    print(f"{W}x{H} N={N} B={B}: |dloss| {float((l1 - l0).abs().max()):.2e}  |dgrad|/max {dg:.2e}  |dlogp| "
          f"{float((p1 - p0).abs().max()):.2e}  |dv| {float((v1 - v0).abs().max()):.2e}  finite {bool(torch.isfinite(g1).all())}",
          flush=True)
    return dg < 1e-5


def perf(W, H, N, B, iters=50):
    ops, recs, items, pis, vs = _examples(B, W, H, N, seed=3)
    m = _module(W, H, N)
    for fused in (True, False):
        L = learner(W, H, N, B, fused, m)
        for _ in range(5):
            L.grad(recs, items, pis, vs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            L.grad(recs, items, pis, vs)
        e1.record()
        torch.cuda.synchronize()
        print(f"perf {W}x{H} B={B} {'fused' if fused else 'per-layer'}: {e0.elapsed_time(e1) / iters * 1e3:.1f} us per grad",
              flush=True)
        L.close()


if __name__ == "__main__":
    good = True
    for c in [(15, 15, 10, 64), (15, 15, 10, 37), (20, 20, 10, 48), (9, 12, 5, 130), (15, 15, 10, 512), (15, 15, 10, 1),
              (15, 15, 10, 2400), (32, 28, 16, 20), (3, 2, 2, 9)]:
        good = check(*c) and good
    print("ALL CHECKS", "PASS" if good else "FAIL", flush=True)
    for B in (64, 128, 256, 512):
        perf(15, 15, 10, B)
    perf(20, 20, 10, 64)
