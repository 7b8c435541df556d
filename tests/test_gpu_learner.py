"""-m gpu: the hand-written fp32 learner step (csrc/bpp_learner.cu) against torch autograd on the
reference architecture (BinpackingNNet.py:15-81) with the reference's losses (NNet.py:87-91) and optimiser (NNet.py:31).
Tolerances: fp32 summation order differs, so gradients are compared relative to each tensor's largest entry."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False        # the comparison target is true fp32 (cuDNN convolutions default to TF32)
torch.backends.cuda.matmul.allow_tf32 = False

GRAD_RTOL = 2e-4   # max |g - g_torch| <= GRAD_RTOL * max |g_torch| per parameter tensor
LOSS_RTOL = 1e-5


def _examples(M, W, H, N, seed):
    """M compact states reached by random legal play (so that the binary input planes have the tied max-pool windows
    real states have), random target policies over the legal moves, targets z in {-1, +1}."""
    from resource_packing_self_play_b200.engine import EnvOps
    from resource_packing_self_play_b200.game import ItemsGenerator
    dev = torch.device("cuda")
    rng = np.random.default_rng(seed)
    gen = ItemsGenerator(W, H, N)
    items = torch.from_numpy(gen.items_batch(range(seed, seed + M))).to(dev)
    ops = EnvOps(W, H, N)
    recs = np.zeros((M, 32), dtype=np.uint32)
    recs[:, 28] = (1 << N) - 1
    recs = torch.from_numpy(recs.view(np.int32)).to(dev)
    depth = torch.from_numpy(rng.integers(0, N, M)).to(dev)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    for k in range(N - 1):
        valid = ops.valid_moves(recs, items).float()
        has = valid.sum(1) > 0
        act = torch.multinomial(valid + (~has)[:, None].float(), 1, generator=g)[:, 0].int()
        nxt = ops.next_state(recs, items, act)
        move = (has & (depth > k))[:, None]
        recs = torch.where(move, nxt, recs)
    valid = ops.valid_moves(recs, items).float()
    pis = torch.rand((M, W * N), device=dev, generator=g) * valid
    s = pis.sum(1, keepdim=True)
    pis = torch.where(s > 0, pis / s.clamp_min(1e-30), torch.full_like(pis, 1.0 / (W * N)))
    vs = (torch.randint(0, 2, (M,), device=dev, generator=g) * 2 - 1).float()
    return ops, recs.contiguous(), items.contiguous(), pis.contiguous(), vs.contiguous()


def _module(W, H, N, seed=0, scale=1.0):
    from resource_packing_self_play_b200.game import BinPackingGame
    from resource_packing_self_play_b200.nnet import BinPackingNNet
    from resource_packing_self_play_b200.utils import dotdict
    torch.manual_seed(seed)
    m = BinPackingNNet(BinPackingGame(W, H, N, 1), dotdict(num_items=N, num_bins=1)).cuda()
    if scale != 1.0:
        with torch.no_grad():
            for p in m.parameters():
                p.mul_(scale)
    return m


def _torch_step(m, ops, recs, items, pis, vs):
    m.zero_grad(set_to_none=True)
    logp, v = m(ops.planes(recs, items))
    l_pi = -torch.sum(pis * logp) / pis.shape[0]
    l_v = torch.sum((vs - v.view(-1)) ** 2) / vs.shape[0]
    (l_pi + l_v).backward()
    return float(l_pi.detach()), float(l_v.detach()), logp.detach(), v.detach().view(-1)


@pytest.mark.parametrize("W,H,N,B", [(15, 15, 10, 64), (15, 15, 10, 37), (20, 20, 10, 48), (9, 12, 5, 130), (15, 15, 10, 512),
                                     (15, 15, 10, 1), (15, 15, 10, 2400), (32, 28, 16, 20), (3, 2, 2, 9)])
def test_gradients_match_autograd(W, H, N, B):
    from resource_packing_self_play_b200.nnet import DeviceLearner
    ops, recs, items, pis, vs = _examples(B, W, H, N, seed=7)
    m = _module(W, H, N, scale=1.5)
    lpi, lv, logp_t, v_t = _torch_step(m, ops, recs, items, pis, vs)
    L = DeviceLearner(W, H, N, max_batch=B)
    L.load_state_dict(m.state_dict())
    logp = torch.empty((B, W * N), device="cuda")
    v = torch.empty(B, device="cuda")
    losses = L.grad(recs, items, pis, vs, logp_out=logp, v_out=v).cpu().numpy()
    assert abs(losses[0] - lpi) <= LOSS_RTOL * abs(lpi) + 1e-6 and abs(losses[1] - lv) <= LOSS_RTOL * abs(lv) + 1e-6
    assert float((logp - logp_t).abs().max()) < 1e-4 and float((v - v_t).abs().max()) < 1e-5
    gd = L.grad_dict()
    for name, p in m.named_parameters():
        ref = p.grad.reshape(-1)
        err = float((gd[name] - ref).abs().max())
        assert err <= GRAD_RTOL * float(ref.abs().max()) + 1e-9, (name, err, float(ref.abs().max()))
    L.close()


@pytest.mark.parametrize("W,H,N,B", [(15, 15, 10, 64), (15, 15, 10, 300), (20, 20, 10, 48), (9, 12, 5, 130), (15, 15, 10, 1201),
                                     (32, 28, 16, 20), (3, 2, 2, 9)])
def test_fused_stage_kernels_match_the_per_layer_kernels(W, H, N, B, monkeypatch):
    """one launch per ConvSequence and direction (k_lr_stage_fwd / k_lr_stage_bwd, the default up to 128 samples) against
    one launch per layer: the same work decomposition per layer, so the results agree to fp32 summation noise (bit for
    bit where both paths pick the same number of samples per CTA)"""
    from resource_packing_self_play_b200.nnet import DeviceLearner
    ops, recs, items, pis, vs = _examples(B, W, H, N, seed=9)
    m = _module(W, H, N, scale=1.5)
    out = []
    for fused in ("1", "0"):
        monkeypatch.setenv("BPP_LEARNER_FUSED", fused)
        L = DeviceLearner(W, H, N, max_batch=B)
        L.load_state_dict(m.state_dict())
        logp = torch.empty((B, W * N), device="cuda")
        v = torch.empty(B, device="cuda")
        losses = L.grad(recs, items, pis, vs, logp_out=logp, v_out=v).clone()
        out.append((losses, L.grads.clone(), logp, v))
        L.close()
    (l1, g1, p1, v1), (l0, g0, p0, v0) = out
    assert bool(torch.isfinite(g1).all())
    assert float((l1 - l0).abs().max()) < 1e-5
    assert float((g1 - g0).abs().max()) <= 2e-5 * float(g0.abs().max())
    assert float((p1 - p0).abs().max()) < 1e-5 and float((v1 - v0).abs().max()) < 1e-5


def test_gather_index_and_determinism():
    from resource_packing_self_play_b200.nnet import DeviceLearner
    W, H, N, M, B = 15, 15, 10, 200, 64
    ops, recs, items, pis, vs = _examples(M, W, H, N, seed=3)
    m = _module(W, H, N)
    ids = torch.randint(0, M, (B,), device="cuda")
    L = DeviceLearner(W, H, N, max_batch=B)
    L.load_state_dict(m.state_dict())
    l1 = L.grad(recs, items, pis, vs, ids=ids).clone()
    g1 = L.grads.clone()
    l2 = L.grad(recs[ids].contiguous(), items[ids].contiguous(), pis[ids].contiguous(), vs[ids].contiguous()).clone()
    assert torch.equal(l1, l2) and torch.equal(g1, L.grads)          # gather == pre-gathered, bit for bit
    L.grad(recs, items, pis, vs, ids=ids)
    assert torch.equal(g1, L.grads)                                   # run-to-run deterministic
    lpi, lv, _, _ = _torch_step(m, ops, recs[ids], items[ids], pis[ids], vs[ids])
    assert abs(float(l1[0]) - lpi) < 1e-4 and abs(float(l1[1]) - lv) < 1e-4
    L.close()


def test_adam_matches_torch_optimizer():
    """given identical gradients the update is torch.optim.Adam's (default hyper-parameters, NNet.py:31)"""
    from resource_packing_self_play_b200.nnet import DeviceLearner
    W, H, N, B = 15, 15, 10, 32
    ops, recs, items, pis, vs = _examples(B, W, H, N, seed=11)
    m = _module(W, H, N)
    L = DeviceLearner(W, H, N, max_batch=B)
    L.load_state_dict(m.state_dict())
    opt = torch.optim.Adam(m.parameters())
    for _ in range(5):
        _torch_step(m, ops, recs, items, pis, vs)
        with torch.no_grad():  # feed the torch gradients to both optimisers
            for name, p in m.named_parameters():
                off, numel = L.slices[name]
                L.grads[off:off + numel].copy_(p.grad.reshape(-1))
        opt.step()
        L.adam()
    for name, p in m.named_parameters():
        off, numel = L.slices[name]
        assert float((L.params[off:off + numel] - p.detach().reshape(-1)).abs().max()) < 2e-6, name
    L.close()


def test_training_trajectory_follows_torch():
    """30 full steps (own gradients, own Adam) on a fixed batch: the loss curve tracks torch's"""
    from resource_packing_self_play_b200.nnet import DeviceLearner
    W, H, N, B = 15, 15, 10, 64
    ops, recs, items, pis, vs = _examples(B, W, H, N, seed=5)
    m = _module(W, H, N)
    L = DeviceLearner(W, H, N, max_batch=B)
    L.load_state_dict(m.state_dict())
    opt = torch.optim.Adam(m.parameters())
    ours, theirs = [], []
    for _ in range(30):
        lpi, lv, _, _ = _torch_step(m, ops, recs, items, pis, vs)
        opt.step()
        theirs.append(lpi + lv)
        ours.append(float(L.grad(recs, items, pis, vs).sum()))
        L.adam()
    assert theirs[-1] < theirs[0] and ours[-1] < ours[0]
    assert np.abs(np.array(ours) - np.array(theirs)).max() < 2e-2 * theirs[0]
    # written back into a torch module, the trained parameters reproduce the learner's own forward
    m2 = _module(W, H, N)
    L.state_dict_into(m2)
    with torch.no_grad():
        logp_t, v_t = m2(ops.planes(recs, items))
    logp = torch.empty((B, W * N), device="cuda")
    v = torch.empty(B, device="cuda")
    L.grad(recs, items, pis, vs, train=False, logp_out=logp, v_out=v)
    assert float((logp - logp_t).abs().max()) < 1e-4 and float((v - v_t.view(-1)).abs().max()) < 1e-5
    L.close()


def test_wrapper_train_paths_agree():
    """NNetWrapper.train / train_compact with the CUDA learner against the torch-autograd cross-check: same minibatch
    stream (np.random / seeded generator), so the trained weights agree to fp32 noise amplified by Adam."""
    from resource_packing_self_play_b200.game import BinPackingGame
    from resource_packing_self_play_b200.nnet import NNetWrapper
    from resource_packing_self_play_b200.utils import dotdict
    W, H, N, M = 15, 15, 10, 256
    ops, recs, items, pis, vs = _examples(M, W, H, N, seed=21)
    args = dotdict(num_items=N, num_bins=1, epochs=2, batch_size=32, cuda=True)
    g = BinPackingGame(W, H, N, 1)
    out = {}
    for learner in ("cuda", "torch"):
        torch.manual_seed(0)
        net = NNetWrapper(g, args, max_batch=64)
        out[learner] = net.train_compact(recs, items, pis, vs, ops, seed=5, use_graph=(learner == "cuda"), learner=learner)
        out[learner + "_w"] = torch.cat([p.detach().reshape(-1) for p in net.nnet.parameters()])
    assert abs(out["cuda"][0] - out["torch"][0]) < 2e-2 * abs(out["torch"][0])
    assert abs(out["cuda"][1] - out["torch"][1]) < 5e-2 * abs(out["torch"][1]) + 1e-3
    # Adam moves a weight whose gradient is ~0 by up to lr per step in a direction set by rounding noise: bound the
    # largest difference by lr * steps and ask the typical one to be far smaller
    d = (out["cuda_w"] - out["torch_w"]).abs()
    assert float(d.max()) <= 2 * 1e-3 * 16 and float(d.mean()) < 1e-3, (float(d.max()), float(d.mean()))
    # reference-surface train(): dense boards in, same np.random minibatch stream on both paths
    planes = ops.planes(recs, items).cpu().numpy().astype(np.int64)
    examples = [(planes[i], pis[i].cpu().numpy(), float(vs[i])) for i in range(M)]
    ws = {}
    for learner in ("cuda", "torch"):
        torch.manual_seed(0)
        net = NNetWrapper(g, args, max_batch=64)
        np.random.seed(9)
        net.train(examples, learner=learner)
        ws[learner] = torch.cat([p.detach().reshape(-1) for p in net.nnet.parameters()])
        pi, v = net.predict(planes[0])
        assert abs(pi.sum() - 1) < 1e-3
    d = (ws["cuda"] - ws["torch"]).abs()
    assert float(d.max()) <= 2 * 1e-3 * 16 and float(d.mean()) < 1e-3, (float(d.max()), float(d.mean()))
