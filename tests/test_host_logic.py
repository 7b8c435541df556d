"""CPU: host-side logic of the package (no CUDA calls): instance generator, state packing, the C-ABI surface."""
import ctypes
import os
import re

import numpy as np

from helpers import load_items_golden
from oracle import bpp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from resource_packing_self_play_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "bpp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bpp_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 35
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/bpp_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.bpp_version() >= 100
    # error plumbing works without a GPU: bad arguments are rejected before any CUDA call
    assert lib.bpp_engine_create(None, None) == -1
    assert b"null" in lib.bpp_last_error()


def test_missing_library_fails_loudly(monkeypatch):
    from resource_packing_self_play_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libbpp_b200.so")
    try:
        _lib.load()
        assert False, "expected BppError"
    except _lib.BppError as e:
        assert "no CPU fallback" in str(e).lower() or "There is no CPU fallback" in str(e)


def test_items_generator_matches_reference_fixtures_and_keeps_global_rng_semantics():
    from resource_packing_self_play_b200.game import ItemsGenerator
    for rec in load_items_golden():
        gen = ItemsGenerator(rec["W"], rec["Hgen"], rec["N"])
        got = gen.items_generator(rec["seed"])
        assert [[int(v) for v in it] for it in got] == rec["items"]
    # like the reference, items_generator re-seeds numpy's global generator (BinPackingGame.py:258) ...
    gen = ItemsGenerator(15, 9, 10)
    gen.items_generator(123)
    a = np.random.randint(1 << 30)
    O.OracleItemsGenerator(15, 9, 10).items_generator(123)
    assert a == np.random.randint(1 << 30)
    # ... while the batched variant leaves it alone and agrees with it
    np.random.seed(7)
    x = np.random.randint(1 << 30)
    np.random.seed(7)
    batch = gen.items_batch([5, 6, 7], [9, 4, 15])
    assert x == np.random.randint(1 << 30)
    for s, h, got in zip([5, 6, 7], [9, 4, 15], batch):
        want = ItemsGenerator(15, h, 10).items_generator(s)
        assert [list(map(int, it[:2])) for it in want] == got.tolist()
    assert gen.bin_height == 9


def test_pack_unpack_round_trip_matches_oracle_layout():
    from resource_packing_self_play_b200.engine import pack_states, ranked_threshold, unpack_states
    rng = np.random.RandomState(0)
    for (W, H, N) in [(15, 15, 10), (20, 20, 10), (12, 9, 6), (32, 28, 16)]:
        g = O.OracleGame(W, H, N, 1)
        items = [[int(rng.randint(1, W + 1)), int(rng.randint(1, H + 1)), 0, 0] for _ in range(N)]
        planes = g.getInitItems(items)
        board = (rng.rand(H, W) < 0.3).astype(np.int64)
        for i in rng.choice(N, N // 2, replace=False):
            planes[i] = planes[i] * 0
        st = g.getBinItem(board, planes)
        recs, wh = pack_states(st, W, H, N)
        occ, rem = O.pack_state(st)
        assert recs[0, :H].tolist() == occ and int(recs[0, 28]) == rem
        back = unpack_states(recs, wh, W, H, N)[0]
        assert np.array_equal(back, st)
    assert np.isnan(ranked_threshold([], 0.75)) and ranked_threshold([0.5, 0.6, 0.7, 0.8001, 0.9], 0.75) == 0.7


def test_algorithmic_bytes_formula_matches_survey_figures():
    from resource_packing_self_play_b200.engine import algorithmic_bytes_per_sim
    b = algorithmic_bytes_per_sim(15, 15, 10, 4.0, 0.5)  # ~14.4 KB/sim at d=4.0, e=0.5
    assert 14000 < b < 14800
    assert 18000 < algorithmic_bytes_per_sim(20, 20, 10, 4.0, 0.5) < 20000


def test_dotdict_and_average_meter():
    from resource_packing_self_play_b200 import AverageMeter, dotdict
    d = dotdict(a=1)
    d.b = 2
    assert d.a == 1 and d["b"] == 2
    m = AverageMeter()
    m.update(2.0, 2)
    m.update(4.0, 2)
    assert m.avg == 3.0


def test_torch_module_is_the_reference_architecture():
    """BinPackingNNet (the parameter container / learner graph) reproduces the reference net's outputs on CPU fp32 and
    has the reference's parameter names and count (170,583 at the default size)."""
    import torch
    from resource_packing_self_play_b200.nnet import BinPackingNNet
    from resource_packing_self_play_b200.utils import dotdict
    from helpers import GOLDEN

    class G:
        def __init__(self, W, H, N):
            self.W, self.H, self.N = W, H, N

        def getBoardSize(self):
            return (self.H, self.W)

        def getActionSize(self):
            return self.W * self.N
    d = np.load(os.path.join(GOLDEN, "net.npz"))
    for tag, (W, H, N) in (("ck", (15, 15, 10)), ("r20", (20, 20, 10))):
        net = BinPackingNNet(G(W, H, N), dotdict(num_items=N, num_bins=1))
        sd = {k[len(tag) + 3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith(tag + "_w.")}
        net.load_state_dict(sd)  # strict: names and shapes must match the reference checkpoint
        net.eval()
        with torch.no_grad():
            lp, v = net(torch.from_numpy(d[tag + "_states"].astype(np.float32)))
        assert np.abs(lp.exp().numpy() - d[tag + "_pi"]).max() < 1e-4
        assert np.abs(v.view(-1).numpy() - d[tag + "_v"]).max() < 1e-4
        if tag == "ck":
            assert sum(p.numel() for p in net.parameters()) == 170583


def test_multiply_shift_division_used_by_the_forward_kernel_is_exact():
    """bpptc::fdiv (csrc/bpp_net_tc.cuh): n // d == (n * ceil(2^32 / d)) >> 32 for every divisor and dividend the kernel
    can form (row indices below 2^15 + tile padding, divisors = rows per leaf, row pitch, h*w, w, batch slices);
    d = 1 is encoded as magic 0 and bypasses the multiply."""
    n = np.arange(0, 40000, dtype=np.uint64)
    for d in list(range(2, 1200)) + [4095, 4096, 4097, 16383, 16384, 32767, 65535]:
        m = np.uint64(0xFFFFFFFF // d + 1)
        assert m < (1 << 32)
        q = (n * m) >> np.uint64(32)
        assert np.array_equal(q, n // np.uint64(d)), d
    # the bound the kernel relies on: exact whenever n * d < 2^32
    rng = np.random.RandomState(0)
    d = rng.randint(2, 65536, size=200000).astype(np.uint64)
    nn = (rng.randint(0, 1 << 31, size=200000).astype(np.uint64) * np.uint64(2)) % ((np.uint64(1) << np.uint64(32)) // d)
    m = np.uint64(0xFFFFFFFF) // d + np.uint64(1)
    assert np.array_equal((nn * m) >> np.uint64(32), nn // d)


def test_trim_rewards_equals_the_reference_loop():
    """learn_batched appends a whole batch of scores and trims once; CoachBPP.py:134-139 appends one score per episode and
    trims after the iteration with `while len > cap: pop(argmin)`.  Same multiset, same order of the survivors — also at
    G = numEps = 20 with ties (the shipped buffer is 100 x 1.0)."""
    from resource_packing_self_play_b200.coach import trim_rewards
    rng = np.random.RandomState(0)
    for trial in range(200):
        cap = int(rng.choice([5, 20, 100]))
        old = list(rng.choice([1.0, 0.5, 0.75, 0.9], size=rng.randint(0, cap + 1))) if trial % 2 else [1.0] * cap
        new = list(np.round(rng.random_sample(rng.choice([1, 20, 37])), 2))
        ref = list(old) + list(new)
        while len(ref) > cap:  # the reference's loop, CoachBPP.py:136-139
            ref.pop(int(np.argmin(ref)))
        assert trim_rewards(old, new, cap) == [float(x) for x in ref]
