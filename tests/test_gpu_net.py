"""-m gpu: the CUDA policy/value forward (bf16 weights + activations, fp32 accumulate) against
(a) outputs of the reference's own BinPackingNNet (fp32, CPU) stored in tests/golden/net.npz with the weights used, and
(b) a plain fp32 torch forward of the same architecture on the device.

Tolerance (stated by SURVEY.md §8(d), config 3): |d pi|_inf <= 2e-2 and |d v| <= 2e-2; the bf16 mode is held to
it on the seeded random-init 20x20 network, the split-bf16 tensor-core mode ("bf16x3") on BOTH networks.  The shipped TRAINED checkpoint has logits in [-3.2e3, -44]: rounding its
weights to bf16 alone moves the policy by up to 0.33 and the value by 0.09 (tests/golden/make_golden.py outputs vs a CPU
emulation of bf16 rounding), so that checkpoint is held to the tolerance in the fp32 mode of the kernel, and the bf16
mode is only checked to be exactly what bf16 rounding of a fp32 torch forward gives (same arg-max on >= 90 % of the
states)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN

pytestmark = pytest.mark.gpu
TOL_PI, TOL_V = 2e-2, 2e-2


class _Game:
    def __init__(self, W, H, N):
        self.W, self.H, self.N = W, H, N

    def getBoardSize(self):
        return (self.H, self.W)

    def getActionSize(self):
        return self.W * self.N


def _wrapper(W, H, N, weights, precision="bf16"):
    from resource_packing_self_play_b200.nnet import NNetWrapper
    from resource_packing_self_play_b200.utils import dotdict
    net = NNetWrapper(_Game(W, H, N), dotdict(num_items=N, num_bins=1, cuda=True, epochs=1, batch_size=8),
                      precision=precision)
    sd = {k: torch.from_numpy(v) for k, v in weights.items()}
    net.nnet.load_state_dict(sd)
    net.sync_weights()
    return net


@pytest.mark.parametrize("tag,W,H,N,precision", [("ck", 15, 15, 10, "fp32"), ("r20", 20, 20, 10, "bf16"),
                                                 ("r20", 20, 20, 10, "fp32"), ("ck", 15, 15, 10, "bf16x3"),
                                                 ("r20", 20, 20, 10, "bf16x3")])
def test_forward_matches_reference_outputs(tag, W, H, N, precision):
    from resource_packing_self_play_b200.engine import pack_states
    d = np.load(os.path.join(GOLDEN, "net.npz"))
    weights = {k[len(tag) + 3:]: d[k] for k in d.files if k.startswith(tag + "_w.")}
    net = _wrapper(W, H, N, weights, precision)
    states = d[tag + "_states"].astype(np.int64)
    recs, items = pack_states(states, W, H, N)
    dev = net.device
    pi, v = net.predict_batch(torch.from_numpy(recs.view(np.int32)).to(dev), torch.from_numpy(items).to(dev))
    pi, v = pi.cpu().numpy(), v.cpu().numpy()
    assert np.abs(pi.sum(axis=1) - 1).max() < 1e-4
    assert np.abs(pi - d[tag + "_pi"]).max() <= TOL_PI
    assert np.abs(v - d[tag + "_v"]).max() <= TOL_V
    # the same states through torch fp32 on the device (plain PyTorch reference of the op)
    with torch.no_grad():
        lp, tv = net.nnet(torch.from_numpy(states.astype(np.float32)).to(dev))
    assert np.abs(pi - lp.exp().cpu().numpy()).max() <= TOL_PI
    assert np.abs(v - tv.view(-1).cpu().numpy()).max() <= TOL_V
    # single-state drop-in call: NNetWrapper.predict(board) -> (pi (A,), v (1,))
    p1, v1 = net.predict(states[3])
    assert p1.shape == (W * N,) and v1.shape == (1,) and p1.dtype == np.float32
    assert np.array_equal(p1, pi[3]) and v1[0] == v[3]


def test_forward_with_game_index_and_device_count():
    from resource_packing_self_play_b200.engine import pack_states
    d = np.load(os.path.join(GOLDEN, "net.npz"))
    weights = {k[5:]: d[k] for k in d.files if k.startswith("ck_w.")}
    net = _wrapper(15, 15, 10, weights)
    states = d["ck_states"].astype(np.int64)
    recs, items = pack_states(states, 15, 15, 10)
    dev = net.device
    recs_t, items_t = torch.from_numpy(recs.view(np.int32)).to(dev), torch.from_numpy(items).to(dev)
    pi_ref, v_ref = net.predict_batch(recs_t, items_t)
    perm = torch.randperm(len(states), device=dev).to(torch.int32)
    count = torch.tensor([20], dtype=torch.int32, device=dev)
    pi = torch.zeros_like(pi_ref)
    v = torch.zeros_like(v_ref)
    net.dnet.forward(recs_t[perm.long()].contiguous(), items_t, game=perm, count_dev=count, policy_out=pi, value_out=v)
    assert torch.equal(pi[:20], pi_ref[perm.long()][:20]) and torch.equal(v[:20], v_ref[perm.long()][:20])
    assert float(pi[20:].abs().sum()) == 0.0  # rows beyond the device-side count are not touched


def test_checkpoint_round_trip(tmp_path):
    d = np.load(os.path.join(GOLDEN, "net.npz"))
    weights = {k[5:]: d[k] for k in d.files if k.startswith("ck_w.")}
    net = _wrapper(15, 15, 10, weights)
    net.save_checkpoint(str(tmp_path), "x.pth.tar")
    ck = torch.load(os.path.join(str(tmp_path), "x.pth.tar"), map_location="cpu")
    assert set(ck.keys()) == {"state_dict"}
    assert set(ck["state_dict"].keys()) == set(weights.keys())
    net2 = _wrapper(15, 15, 10, {k: np.zeros_like(v) for k, v in weights.items()})
    net2.load_checkpoint(str(tmp_path), "x.pth.tar")
    st = d["ck_states"][:4].astype(np.int64)
    a, b = net.predict(st[0]), net2.predict(st[0])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_bf16_mode_on_trained_checkpoint_is_plain_bf16_rounding():
    """documents the precision finding: the bf16 kernel equals a bf16-rounded torch forward, and keeps the arg-max"""
    import torch.nn.functional as F
    from resource_packing_self_play_b200.engine import pack_states
    d = np.load(os.path.join(GOLDEN, "net.npz"))
    weights = {k[5:]: d[k] for k in d.files if k.startswith("ck_w.")}
    net = _wrapper(15, 15, 10, weights, "bf16")
    states = d["ck_states"].astype(np.int64)
    recs, items = pack_states(states, 15, 15, 10)
    dev = net.device
    pi, v = net.predict_batch(torch.from_numpy(recs.view(np.int32)).to(dev), torch.from_numpy(items).to(dev))
    pi = pi.cpu().numpy()

    def bf(t):
        return t.to(torch.bfloat16).to(torch.float32)
    Wt = {k: torch.from_numpy(x) for k, x in weights.items()}
    x = torch.from_numpy(states.astype(np.float32))
    for s in range(3):
        p = f"conv_seqs.{s}."
        x = F.max_pool2d(bf(F.conv2d(x, bf(Wt[p + "conv.weight"]), Wt[p + "conv.bias"], padding=1)), 3, 2, 1)
        for b in range(2):
            q = p + f"res_block{b}."
            y = bf(F.conv2d(F.relu(x), bf(Wt[q + "conv0.weight"]), Wt[q + "conv0.bias"], padding=1))
            x = bf(F.conv2d(F.relu(y), bf(Wt[q + "conv1.weight"]), Wt[q + "conv1.bias"], padding=1) + x)
    h = bf(F.relu(F.linear(F.relu(torch.flatten(x, 1)), bf(Wt["hidden_fc.weight"]), Wt["hidden_fc.bias"])))
    emu = F.softmax(F.linear(h, bf(Wt["logits_fc.weight"]), Wt["logits_fc.bias"]), 1).numpy()
    assert np.abs(pi - emu).max() < 5e-2          # accumulation-order noise on logits of magnitude 1e2..3e3
    agree = (pi.argmax(1) == d["ck_pi"].argmax(1)).mean()
    assert agree >= 0.9, agree


@pytest.mark.parametrize("tag,W,H,N,tol", [("r20", 20, 20, 10, 2e-4), ("ck", 15, 15, 10, 6e-2)])
def test_tensor_core_kernel_matches_the_simt_kernel_with_the_same_roundings(tag, W, H, N, tol):
    """tcgen05 path (mode bf16) vs the CUDA-core kernel applying the same bf16 roundings (mode bf16_simt): only the
    fp32 accumulation order differs.  Odd batch sizes exercise partial groups and tiles."""
    from resource_packing_self_play_b200.engine import pack_states
    d = np.load(os.path.join(GOLDEN, "net.npz"))
    weights = {k[len(tag) + 3:]: d[k] for k in d.files if k.startswith(tag + "_w.")}
    net = _wrapper(W, H, N, weights, "bf16")
    states = d[tag + "_states"].astype(np.int64)
    recs, items = pack_states(states, W, H, N)
    dev = net.device
    recs_t, items_t = torch.from_numpy(recs.view(np.int32)).to(dev), torch.from_numpy(items).to(dev)
    for B in (len(states), 1, 7, 13):
        net.dnet.set_precision("bf16")
        pi_tc, v_tc = net.predict_batch(recs_t[:B].contiguous(), items_t[:B].contiguous())
        net.dnet.set_precision("bf16_simt")
        pi_s, v_s = net.predict_batch(recs_t[:B].contiguous(), items_t[:B].contiguous())
        torch.cuda.synchronize()
        assert float((pi_tc.sum(dim=1) - 1).abs().max()) < 1e-4
        assert float((pi_tc - pi_s).abs().max()) <= tol, (B, float((pi_tc - pi_s).abs().max()))
        assert float((v_tc - v_s).abs().max()) <= tol


@pytest.mark.parametrize("W,H,N,B", [(15, 15, 10, 777), (20, 20, 10, 301), (9, 12, 5, 64)])
def test_both_shared_memory_plans_of_the_tensor_core_kernel_agree_bit_for_bit(W, H, N, B, monkeypatch):
    """The classic and the compact shared-memory arena (the latter aliases the level-0 conv output with its own input
    planes and relocates triples and weights) run the same arithmetic in the same order: identical outputs, for ragged
    batch sizes (partial groups, partial tiles)."""
    from resource_packing_self_play_b200.engine import pack_states  # noqa: F401
    from resource_packing_self_play_b200.game import ItemsGenerator
    from resource_packing_self_play_b200.nnet import NNetWrapper
    from resource_packing_self_play_b200.utils import dotdict
    rng = np.random.RandomState(1)
    recs = np.zeros((B, 32), dtype=np.uint32)
    recs[:, :H] = rng.randint(0, 1 << W, size=(B, H)) & rng.randint(0, 1 << W, size=(B, H))
    recs[:, 28] = rng.randint(1, 1 << N, size=B)
    items = ItemsGenerator(W, H, N).items_batch(np.arange(B) % 53 + 9, None)
    outs = []
    monkeypatch.setenv("BPP_NO_GR", "1")   # the one-kernel trunk (fallback of the grid-row stage kernels)
    for compact in ("0", "1"):
        monkeypatch.setenv("BPP_TC_COMPACT", compact)
        torch.manual_seed(4)
        net = NNetWrapper(_Game(W, H, N), dotdict(num_items=N, num_bins=1, cuda=True, epochs=1, batch_size=8),
                          max_batch=B, precision="bf16")
        dev = net.device
        pol, val = net.dnet.forward(torch.from_numpy(recs.view(np.int32)).to(dev), torch.from_numpy(items).to(dev))
        torch.cuda.synchronize()
        outs.append((pol.cpu().numpy(), val.cpu().numpy()))
    assert np.isfinite(outs[0][0]).all() and abs(outs[0][0].sum(axis=1) - 1).max() < 1e-4
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("W,H,N,B", [(15, 15, 10, 777), (20, 20, 10, 301), (9, 12, 5, 64), (15, 15, 10, 3)])
def test_role_kernels_equal_the_one_kernel_trunk_bit_for_bit(W, H, N, B, monkeypatch):
    """The trunk split at the ConvSequence boundaries (k_net_role<0..2>, each role with its own group size, residual
    stream handed over through HBM as bf16) runs the same arithmetic in the same order as the one-kernel trunk
    (BPP_NO_ROLES=1): identical outputs for ragged batch sizes."""
    from resource_packing_self_play_b200.game import ItemsGenerator
    from resource_packing_self_play_b200.nnet import NNetWrapper
    from resource_packing_self_play_b200.utils import dotdict
    rng = np.random.RandomState(2)
    recs = np.zeros((B, 32), dtype=np.uint32)
    recs[:, :H] = rng.randint(0, 1 << W, size=(B, H)) & rng.randint(0, 1 << W, size=(B, H))
    recs[:, 28] = rng.randint(1, 1 << N, size=B)
    items = ItemsGenerator(W, H, N).items_batch(np.arange(B) % 53 + 9, None)
    outs = []
    monkeypatch.setenv("BPP_NO_GR", "1")   # role kernels / one-kernel trunk (the split-bf16 mode's kernels, bf16 fallback)
    for no_roles in ("", "1"):
        monkeypatch.setenv("BPP_ROLES_MIN_BATCH", "1")   # role kernels at any batch size (default: large batches only)
        if no_roles:
            monkeypatch.setenv("BPP_NO_ROLES", "1")
        else:
            monkeypatch.delenv("BPP_NO_ROLES", raising=False)
        torch.manual_seed(4)
        net = NNetWrapper(_Game(W, H, N), dotdict(num_items=N, num_bins=1, cuda=True, epochs=1, batch_size=8),
                          max_batch=B, precision="bf16")
        with torch.no_grad():
            net.nnet.logits_fc.weight.mul_(25.0)
        net.sync_weights()
        dev = net.device
        pol, val = net.dnet.forward(torch.from_numpy(recs.view(np.int32)).to(dev), torch.from_numpy(items).to(dev))
        count = torch.tensor([max(1, B // 3)], dtype=torch.int32, device=dev)   # device-side batch size
        pol2 = torch.zeros_like(pol)
        val2 = torch.zeros_like(val)
        net.dnet.forward(torch.from_numpy(recs.view(np.int32)).to(dev), torch.from_numpy(items).to(dev), count_dev=count,
                         policy_out=pol2, value_out=val2)
        torch.cuda.synchronize()
        outs.append((pol.cpu().numpy(), val.cpu().numpy(), pol2.cpu().numpy(), val2.cpu().numpy()))
    assert np.isfinite(outs[0][0]).all() and abs(outs[0][0].sum(axis=1) - 1).max() < 1e-4
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
    k = max(1, B // 3)
    assert np.array_equal(outs[0][2][:k], outs[0][0][:k]) and not outs[0][2][k:].any()


@pytest.mark.parametrize("W,H,N,B", [(15, 15, 10, 777), (20, 20, 10, 301), (9, 12, 5, 64), (15, 15, 10, 3), (12, 9, 16, 130),
                                     (25, 28, 10, 50)])
def test_grid_row_stage_kernels_match_the_previous_tensor_core_path_and_the_simt_kernel(W, H, N, B, monkeypatch):
    """The grid-row stage kernels (k_net_gr<0..3>, bpp_net_gr.cuh: vertical taps stacked along N, in-place layers, bias
    pre-loaded into the accumulators) apply the same bf16 roundings as the previous tensor-core path (BPP_NO_GR=1) and as the
    CUDA-core kernel bf16_simt; only the fp32 accumulation order differs.  Ragged batches (partial groups), the device-side
    batch size, 17 input channels (two K chunks in the first layer) and a 28-row board (28 tiles per layer)."""
    from resource_packing_self_play_b200.game import ItemsGenerator
    from resource_packing_self_play_b200.nnet import NNetWrapper
    from resource_packing_self_play_b200.utils import dotdict
    rng = np.random.RandomState(3)
    recs = np.zeros((B, 32), dtype=np.uint32)
    recs[:, :H] = rng.randint(0, 1 << min(W, 30), size=(B, H)) & rng.randint(0, 1 << min(W, 30), size=(B, H))
    recs[:, 28] = rng.randint(1, 1 << N, size=B)
    items = ItemsGenerator(W, H, N).items_batch(np.arange(B) % 53 + 9, None)
    outs = {}
    for tag in ("gr", "prev"):
        if tag == "gr":
            monkeypatch.delenv("BPP_NO_GR", raising=False)
        else:
            monkeypatch.setenv("BPP_NO_GR", "1")
        torch.manual_seed(4)
        net = NNetWrapper(_Game(W, H, N), dotdict(num_items=N, num_bins=1, cuda=True, epochs=1, batch_size=8),
                          max_batch=B, precision="bf16")
        with torch.no_grad():
            net.nnet.logits_fc.weight.mul_(25.0)
        net.sync_weights()
        assert net.dnet.grid_row() == (tag == "gr")
        dev = net.device
        r_t, i_t = torch.from_numpy(recs.view(np.int32)).to(dev), torch.from_numpy(items).to(dev)
        pol, val = net.dnet.forward(r_t, i_t)
        k = max(1, B // 3)
        count = torch.tensor([k], dtype=torch.int32, device=dev)   # device-side batch size
        pol2, val2 = torch.zeros_like(pol), torch.zeros_like(val)
        net.dnet.forward(r_t, i_t, count_dev=count, policy_out=pol2, value_out=val2)
        if tag == "prev":
            net.dnet.set_precision("bf16_simt")
            ps, vs = net.dnet.forward(r_t, i_t)
            outs["simt"] = (ps.cpu().numpy(), vs.cpu().numpy())
        torch.cuda.synchronize()
        outs[tag] = (pol.cpu().numpy(), val.cpu().numpy(), pol2.cpu().numpy(), val2.cpu().numpy())
    g, p, sm = outs["gr"], outs["prev"], outs["simt"]
    assert np.isfinite(g[0]).all() and abs(g[0].sum(axis=1) - 1).max() < 1e-4
    ref = max(np.abs(p[0] - sm[0]).max(), 1e-6)   # what another accumulation order of the same roundings moves
    assert np.abs(g[0] - sm[0]).max() <= max(5e-4, 4 * ref) and np.abs(g[1] - sm[1]).max() <= 2e-3
    assert np.abs(g[0] - p[0]).max() <= max(5e-4, 4 * ref)
    assert np.array_equal(g[2][:k], g[0][:k]) and not g[2][k:].any()   # results do not depend on the grouping


@pytest.mark.parametrize("W,H,N,B", [(15, 15, 10, 333), (20, 20, 10, 150), (9, 12, 5, 64), (9, 28, 5, 40)])
def test_split_bf16_grid_row_kernels_match_the_fp32_kernel_and_the_previous_split_path(W, H, N, B, monkeypatch):
    """The split-bf16 mode of the grid-row stage kernels (hi + lo arena and weights, weights streamed through two slots
    where they do not fit) against the CUDA-core fp32 kernel and the previous split-bf16 kernels (BPP_NO_GR3=1), with the
    logits scaled up as in the reference's trained checkpoints; ragged batches and the device-side batch size."""
    from resource_packing_self_play_b200.game import ItemsGenerator
    from resource_packing_self_play_b200.nnet import NNetWrapper
    from resource_packing_self_play_b200.utils import dotdict
    rng = np.random.RandomState(5)
    recs = np.zeros((B, 32), dtype=np.uint32)
    recs[:, :H] = rng.randint(0, 1 << W, size=(B, H)) & rng.randint(0, 1 << W, size=(B, H))
    recs[:, 28] = rng.randint(1, 1 << N, size=B)
    items = ItemsGenerator(W, H, N).items_batch(np.arange(B) % 53 + 9, None)
    outs = {}
    for tag in ("gr", "prev"):
        if tag == "gr":
            monkeypatch.delenv("BPP_NO_GR3", raising=False)
        else:
            monkeypatch.setenv("BPP_NO_GR3", "1")
        torch.manual_seed(4)
        net = NNetWrapper(_Game(W, H, N), dotdict(num_items=N, num_bins=1, cuda=True, epochs=1, batch_size=8),
                          max_batch=B, precision="bf16x3")
        with torch.no_grad():
            net.nnet.logits_fc.weight.mul_(25.0)
        net.sync_weights()
        assert net.dnet.grid_row() == (tag == "gr")
        dev = net.device
        r_t, i_t = torch.from_numpy(recs.view(np.int32)).to(dev), torch.from_numpy(items).to(dev)
        pol, val = net.dnet.forward(r_t, i_t)
        k = max(1, B // 3)
        count = torch.tensor([k], dtype=torch.int32, device=dev)
        pol2, val2 = torch.zeros_like(pol), torch.zeros_like(val)
        net.dnet.forward(r_t, i_t, count_dev=count, policy_out=pol2, value_out=val2)
        if tag == "prev":
            net.dnet.set_precision("fp32")
            pf, vf = net.dnet.forward(r_t, i_t)
            outs["fp32"] = (pf.cpu().numpy(), vf.cpu().numpy())
        torch.cuda.synchronize()
        outs[tag] = (pol.cpu().numpy(), val.cpu().numpy(), pol2.cpu().numpy())
    g, p, f32 = outs["gr"], outs["prev"], outs["fp32"]
    assert np.isfinite(g[0]).all() and abs(g[0].sum(axis=1) - 1).max() < 1e-4
    assert np.abs(g[0] - f32[0]).max() <= 2e-4 and np.abs(g[1] - f32[1]).max() <= 2e-4   # fp32-level accuracy
    assert np.abs(g[0] - p[0]).max() <= 2e-4
    assert np.array_equal(g[2][:k], g[0][:k]) and not g[2][k:].any()


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_grid_row_kernels_reproduce_their_outputs_bit_for_bit(precision):
    """Repeated launches on the same inputs give identical bits for many batch sizes (one group per SM, partial groups, more
    groups than SMs): a race in the per-tile mbarrier hand-over between the MMA issuer and the epilogue warps, or between the
    two groups of a CTA, would show up as a rare difference."""
    from resource_packing_self_play_b200.game import ItemsGenerator
    from resource_packing_self_play_b200.nnet import NNetWrapper
    from resource_packing_self_play_b200.utils import dotdict
    W, H, N, Bmax = 15, 15, 10, 5000
    rng = np.random.RandomState(11)
    recs = np.zeros((Bmax, 32), dtype=np.uint32)
    recs[:, :H] = rng.randint(0, 1 << W, size=(Bmax, H)) & rng.randint(0, 1 << W, size=(Bmax, H))
    recs[:, 28] = rng.randint(1, 1 << N, size=Bmax)
    items = ItemsGenerator(W, H, N).items_batch(np.arange(Bmax) % 53 + 9, None)
    torch.manual_seed(4)
    net = NNetWrapper(_Game(W, H, N), dotdict(num_items=N, num_bins=1, cuda=True, epochs=1, batch_size=8), max_batch=Bmax,
                      precision=precision)
    assert net.dnet.grid_row()
    dev = net.device
    r_t, i_t = torch.from_numpy(recs.view(np.int32)).to(dev), torch.from_numpy(items).to(dev)
    for B in (1, 9, 148, 297, 1211, 2800, 4097, 5000):
        p0, v0 = net.dnet.forward(r_t[:B].contiguous(), i_t[:B].contiguous())
        p0, v0 = p0.clone(), v0.clone()
        for _ in range(3):
            p1, v1 = net.dnet.forward(r_t[:B].contiguous(), i_t[:B].contiguous())
            assert torch.equal(p1, p0) and torch.equal(v1, v0), B
    # a leaf's result does not depend on the batch it is evaluated in
    pa, _ = net.dnet.forward(r_t[:2800].contiguous(), i_t[:2800].contiguous())
    pb, _ = net.dnet.forward(r_t[:300].contiguous(), i_t[:300].contiguous())
    assert torch.equal(pa[:300], pb)


@pytest.mark.parametrize("W,H,N", [(3, 2, 2), (2, 28, 3), (32, 3, 4), (31, 1, 2), (5, 5, 16), (1, 9, 3), (24, 24, 10), (7, 13, 1)])
def test_grid_row_kernels_on_odd_board_geometries(W, H, N):
    """One-row / one-column boards, 28 grid rows (more than the 16 accumulator slots of the first layer), 17 input channels,
    a single item: the tensor-core modes against the CUDA-core kernels with the same arithmetic (bf16 vs bf16_simt,
    split-bf16 vs fp32) on batches of 1, 37 and 600 leaves."""
    from resource_packing_self_play_b200.nnet import NNetWrapper
    from resource_packing_self_play_b200.utils import dotdict
    rng = np.random.RandomState(0)
    for B in (1, 37, 600):
        recs = np.zeros((B, 32), dtype=np.uint32)
        recs[:, :H] = rng.randint(0, 1 << min(W, 30), size=(B, H))
        recs[:, 28] = rng.randint(1, 1 << N, size=B)
        items = np.stack([rng.randint(1, W + 1, size=(B, N)), rng.randint(1, H + 1, size=(B, N))], axis=2).astype(np.int32)
        torch.manual_seed(1)
        net = NNetWrapper(_Game(W, H, N), dotdict(num_items=N, num_bins=1, cuda=True, epochs=1, batch_size=8), max_batch=B,
                          precision="bf16")
        with torch.no_grad():
            net.nnet.logits_fc.weight.mul_(10.0)
        net.sync_weights()
        dev = net.device
        r_t, i_t = torch.from_numpy(recs.view(np.int32)).to(dev), torch.from_numpy(items).to(dev)
        out = {}
        for mode in ("bf16", "bf16_simt", "bf16x3", "fp32"):
            net.dnet.set_precision(mode)
            p, v = net.dnet.forward(r_t, i_t)
            out[mode] = (p.clone(), v.clone())
        torch.cuda.synchronize()
        assert bool(torch.isfinite(out["bf16"][0]).all())
        assert float((out["bf16"][0] - out["bf16_simt"][0]).abs().max()) < 2e-3
        assert float((out["bf16x3"][0] - out["fp32"][0]).abs().max()) < 2e-4
        assert float((out["bf16x3"][1] - out["fp32"][1]).abs().max()) < 2e-4
