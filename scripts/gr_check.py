"""Development check of the grid-row trunk (bpp_net_gr.cuh) on the GPU: outputs against the CUDA-core kernel with the same
bf16 roundings and against the previous tensor-core path (BPP_NO_GR=1), ragged batches and the device-side batch size;
then throughput.   usage: gr_check.py [quick]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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


def make(W, H, N, B, gr, precision="bf16", scale=1.0):
    for k in ("BPP_NO_GR", "BPP_NO_GR3"):
        if gr:
            os.environ.pop(k, None)
        else:
            os.environ[k] = "1"
    torch.manual_seed(4)
    net = NNetWrapper(G(W, H, N), dotdict(num_items=N, num_bins=1, cuda=True, epochs=1, batch_size=8), max_batch=B,
                      precision=precision)
    if scale != 1.0:
        with torch.no_grad():
            net.nnet.logits_fc.weight.mul_(scale)
        net.sync_weights()
    return net


def data(W, H, N, B, seed=1):
    rng = np.random.RandomState(seed)
    recs = np.zeros((B, 32), dtype=np.uint32)
    recs[:, :H] = rng.randint(0, 1 << W, size=(B, H)) & rng.randint(0, 1 << W, size=(B, H))
    recs[:, 28] = rng.randint(1, 1 << N, size=B)
    items = ItemsGenerator(W, H, N).items_batch(np.arange(B) % 53 + 9, None)
    return torch.from_numpy(recs.view(np.int32)).cuda(), torch.from_numpy(items).cuda()


def check(W, H, N, B):
    recs, items = data(W, H, N, B)
    new = make(W, H, N, B, True, scale=25.0)
    old = make(W, H, N, B, False, scale=25.0)
    pn, vn = new.dnet.forward(recs, items)
    po, vo = old.dnet.forward(recs, items)
    old.dnet.set_precision("bf16_simt")
    ps, vs = old.dnet.forward(recs, items)
    torch.cuda.synchronize()
    ok = bool(torch.isfinite(pn).all())
    d_old = float((pn - po).abs().max())
    d_simt = float((pn - ps).abs().max())
    d_ref = float((po - ps).abs().max())
    dv = float((vn - vs).abs().max())
    # device-side batch size
    k = max(1, B // 3)
    cnt = torch.tensor([k], dtype=torch.int32, device="cuda")
    p2, v2 = torch.zeros_like(pn), torch.zeros_like(vn)
    new.dnet.forward(recs, items, count_dev=cnt, policy_out=p2, value_out=v2)
    torch.cuda.synchronize()
    same = bool(torch.equal(p2[:k], pn[:k])) and not bool(p2[k:].any())
    print(f"check {W}x{H} N={N} B={B}: finite {ok}  |new-old| {d_old:.2e}  |new-simt| {d_simt:.2e}  (|old-simt| {d_ref:.2e})  "
          f"|dv| {dv:.2e}  count_dev ok {same}", flush=True)
    return ok and d_simt < max(5e-3, 4 * d_ref) and same


def check_x3(W, H, N, B):
    """split-bf16 grid-row kernels against the previous split-bf16 path and the fp32 CUDA-core kernel"""
    recs, items = data(W, H, N, B)
    new = make(W, H, N, B, True, precision="bf16x3", scale=25.0)
    old = make(W, H, N, B, False, precision="bf16x3", scale=25.0)
    pn, vn = new.dnet.forward(recs, items)
    po, vo = old.dnet.forward(recs, items)
    old.dnet.set_precision("fp32")
    pf, vf = old.dnet.forward(recs, items)
    torch.cuda.synchronize()
    k = max(1, B // 3)
    cnt = torch.tensor([k], dtype=torch.int32, device="cuda")
    p2, v2 = torch.zeros_like(pn), torch.zeros_like(vn)
    new.dnet.forward(recs, items, count_dev=cnt, policy_out=p2, value_out=v2)
    torch.cuda.synchronize()
    same = bool(torch.equal(p2[:k], pn[:k])) and not bool(p2[k:].any())
    d_new, d_old = float((pn - pf).abs().max()), float((po - pf).abs().max())
    print(f"check x3 {W}x{H} N={N} B={B}: grid_row {new.dnet.grid_row()} finite {bool(torch.isfinite(pn).all())}  |new-fp32| {d_new:.2e}  "
          f"|old-fp32| {d_old:.2e}  |new-old| {float((pn - po).abs().max()):.2e}  |dv| {float((vn - vf).abs().max()):.2e}  "
          f"count_dev ok {same}", flush=True)
    return d_new < max(1e-3, 4 * d_old) and same


def perf(W, H, N, B, gr, iters=20, precision="bf16"):
    recs, items = data(W, H, N, B)
    net = make(W, H, N, B, gr, precision=precision)
    pol = torch.empty((B, W * N), dtype=torch.float32, device="cuda")
    val = torch.empty(B, dtype=torch.float32, device="cuda")
    for _ in range(3):
        net.dnet.forward(recs, items, policy_out=pol, value_out=val)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        net.dnet.forward(recs, items, policy_out=pol, value_out=val)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    extra = ""
    if gr:
        try:
            extra = str(net.dnet.profile_roles())
        except Exception as err:  # noqa: BLE001
            extra = repr(err)
    print(f"perf {precision} {W}x{H} B={B} {'grid-row' if gr else 'previous'}: {ms * 1e3:.1f} us  {B / ms / 1e3:.2f} M evals/s  "
          f"{B * FLOPS[(W, H, N)] / ms / 1e9:.1f} TFLOP/s  {extra}", flush=True)


if __name__ == "__main__":
    os.environ.setdefault("BPP_TC_VERBOSE", "1")
    t0 = time.time()
    good = True
    cases = [(15, 15, 10, 8), (15, 15, 10, 777), (20, 20, 10, 301), (9, 12, 5, 64), (15, 15, 10, 3)]
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        cases = cases[:2]
    os.environ.pop("BPP_TC_VERBOSE", None)
    if len(sys.argv) > 1 and sys.argv[1] == "serial":
        for mode in ("0", "1", "2"):
            os.environ["BPP_GR_SERIAL"] = mode
            print("== BPP_GR_SERIAL =", mode, flush=True)
            for c in cases[:3]:
                check(*c)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "x3":
        os.environ["BPP_TC_VERBOSE"] = "1"
        for c in cases:
            good = check_x3(*c) and good
        os.environ.pop("BPP_TC_VERBOSE", None)
        print("ALL X3 CHECKS", "PASS" if good else "FAIL", flush=True)
        for (W, H) in [(15, 15), (20, 20)]:
            for B in (2800, 8192):
                perf(W, H, 10, B, True, precision="bf16x3")
                perf(W, H, 10, B, False, precision="bf16x3")
        sys.exit(0)
    for c in cases:
        good = check(*c) and good
    print("ALL CHECKS", "PASS" if good else "FAIL", flush=True)
    for (W, H) in [(15, 15), (20, 20)]:
        for B in (2800, 8192):
            perf(W, H, 10, B, True)
            perf(W, H, 10, B, False)
    print("wall", time.time() - t0)
