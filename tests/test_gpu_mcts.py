"""-m gpu: per-move MCTS visit counts of the CUDA engine, bit-exact against fixtures produced by the unmodified
reference (MCTS_bpp.py) under the deterministic stub evaluators."""
import numpy as np
import pytest
import torch

from helpers import bl_of, load_mcts_golden

pytestmark = pytest.mark.gpu

CASES = load_mcts_golden()


def _items_wh(case):
    return np.array([[it[0], it[1]] for it in case["items"]], dtype=np.int32)


def _torch_stub(kind, A):
    """the stub evaluators of tests/golden/make_golden.py on a batch of dense planes (float64 on the device)"""
    def ev(planes):
        pop = planes[:, 0].sum(dim=(1, 2)).to(torch.int64)
        nrem = (planes[:, 1:].sum(dim=(2, 3)) > 0).sum(dim=1).to(torch.int64)
        B = planes.shape[0]
        a = torch.arange(A, device=planes.device, dtype=torch.float64)[None, :]
        v = ((7 * pop + 3 * nrem) % 16).to(torch.float64) / 16 - 0.5
        if kind == "U":
            return torch.full((B, A), 1 / A, dtype=torch.float64, device=planes.device), torch.zeros_like(v)
        if kind == "V":
            return torch.full((B, A), 1 / A, dtype=torch.float64, device=planes.device), v
        if kind == "H":
            return 1.0 / (a + 3 + (pop % 5)[:, None].to(torch.float64)), v
        ai = torch.arange(A, device=planes.device, dtype=torch.int64)[None, :]
        return ((37 * ai + 11 + pop[:, None]) % 64 + 1).to(torch.float64) / 4096.0, v
    return ev


def _play(case, mode, G=3):
    from resource_packing_self_play_b200.engine import SearchEngine
    W, H, N, A = case["W"], case["H"], case["N"], case["W"] * case["N"]
    eng = SearchEngine(W, H, N, G, case["sims"], case["cpuct"])
    items = np.repeat(_items_wh(case)[None], G, axis=0)
    area = np.full(G, case["genW"] * case["genH"], dtype=np.int32)
    bl = np.full(G, bl_of(case["rewards"], case["alpha"]))
    eng.reset(items, area, bl)
    counts_all = []
    for mv, a_ref in enumerate(case["actions"]):
        eng.begin_move()
        if mode == "fused":
            eng.search_stub(case["stub"])
        else:
            eng.search_with(_torch_stub(case["stub"], A))
        counts = eng.root_counts().cpu().numpy()
        counts_all.append(counts)
        if case["policy"] == "argmax":
            chosen = eng.choose(0).cpu().numpy()
            assert (chosen == a_ref).all(), (mv, chosen, a_ref)
        eng.advance(np.full(G, a_ref, dtype=np.int32))
    eng.check()
    st = eng.status()
    return np.stack(counts_all, axis=1), {k: v.cpu().numpy() for k, v in st.items()}, eng


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_fused_stub_visit_counts_bit_exact(ci):
    case = CASES[ci]
    counts, st, eng = _play(case, "fused")
    want = np.array(case["counts"], dtype=np.int64)
    for g in range(counts.shape[0]):
        assert np.array_equal(counts[g], want), f"case {ci} game {g}: visit counts differ"
    assert (st["done"] == 1).all()
    assert (st["r"] == case["r"]).all()
    assert (st["score"] == case["score"]).all()
    assert (st["moves"] == len(case["actions"])).all()
    nodes, _ = eng.graph_sizes()
    # every dict entry of the reference (Es) is a node here; real moves may add the played-to state as well
    n_ref = case["n_expanded"] + case["n_terminal"]
    assert ((nodes.cpu().numpy() >= n_ref) & (nodes.cpu().numpy() <= n_ref + len(case["actions"]))).all()


@pytest.mark.parametrize("ci", [0, 3, 6, 7, 8, 9, 11, 12, 15, 16, 17, 18])
def test_lockstep_external_evaluator_visit_counts_bit_exact(ci):
    case = CASES[ci]
    counts, st, _ = _play(case, "lockstep", G=2)
    want = np.array(case["counts"], dtype=np.int64)
    for g in range(counts.shape[0]):
        assert np.array_equal(counts[g], want), f"case {ci} game {g}: visit counts differ"
    assert (st["r"] == case["r"]).all() and (st["score"] == case["score"]).all()


def test_many_games_mixed_instances_match_single_game_runs():
    """4 different instances interleaved in one engine give the same counts as their fixtures."""
    from resource_packing_self_play_b200.engine import SearchEngine
    sel = [c for c in CASES if (c["W"], c["H"], c["N"], c["sims"], c["cpuct"], c["stub"], c["policy"]) ==
           (15, 15, 10, 200, 1, "V", "argmax") and not c["rewards"]]
    assert len(sel) >= 4
    G = len(sel) * 5
    order = [i % len(sel) for i in range(G)]
    eng = SearchEngine(15, 15, 10, G, 200, 1.0)
    items = np.stack([np.array([[it[0], it[1]] for it in sel[k]["items"]], dtype=np.int32) for k in order])
    area = np.array([sel[k]["genW"] * sel[k]["genH"] for k in order], dtype=np.int32)
    eng.reset(items, area, np.full(G, np.nan))
    counts, actions = eng.play_stub("V", 0)
    eng.check()
    counts, actions = counts.cpu().numpy(), actions.cpu().numpy()
    for g, k in enumerate(order):
        want = np.array(sel[k]["counts"])
        nm = want.shape[0]
        assert np.array_equal(counts[:nm, g], want)
        assert list(actions[:nm, g]) == sel[k]["actions"]
        assert (actions[nm:, g] == -1).all()
    st = eng.status()
    assert (st["done"].cpu().numpy() == 1).all()
    s = eng.stats()
    assert s["sims"] == sum(200 * len(sel[k]["actions"]) for k in order)


def test_whole_episode_kernel_equals_the_per_move_kernels():
    """bpp_engine_play_stub (one launch: search -> counts -> choose -> play for all moves) must reproduce, bit for bit,
    the per-move sequence search_stub / root_counts / choose / advance with the same sampling seed."""
    from resource_packing_self_play_b200 import _lib
    from resource_packing_self_play_b200.engine import SearchEngine
    from resource_packing_self_play_b200.game import ItemsGenerator
    W, H, N, G, SIMS = 15, 15, 10, 256, 60
    heights = (np.arange(G) % 14 + 2).astype(np.int32)
    items = ItemsGenerator(W, H, N).items_batch(np.arange(G) + 77, heights)
    area = (W * heights).astype(np.int32)
    bl = np.full(G, 0.6501)
    for mode in (_lib.CHOOSE_SAMPLE, _lib.CHOOSE_GREEDY, _lib.CHOOSE_ARGMAX_FIRST):
        a = SearchEngine(W, H, N, G, SIMS, 1.0)
        a.reset(items, area, bl)
        counts_a, actions_a = a.play_stub("D", mode, seed=11)
        a.check()
        b = SearchEngine(W, H, N, G, SIMS, 1.0)
        b.reset(items, area, bl)
        counts_b, actions_b = [], []
        for mv in range(N):
            b.begin_move()
            b.search_stub("D")
            counts_b.append(b.root_counts())
            act = b.choose(mode, seed=11)
            actions_b.append(act)
            b.advance(act)
        b.check()
        assert torch.equal(actions_a, torch.stack(actions_b))
        played = (actions_a >= 0)[:, :, None]
        assert torch.equal(counts_a * played, torch.stack(counts_b) * played)
        sa, sb = a.status(), b.status()
        for k in sa:
            assert torch.equal(sa[k], sb[k]), k
        assert a.stats()["sims"] == b.stats()["sims"]


@pytest.mark.parametrize("kind", ["U", "V"])
def test_host_buffer_episode_call_direct_and_staged_paths_agree(kind):
    """bpp_engine_play_stub_host writes the visit counts straight into PINNED result buffers from the episode kernel
    (mapped host memory) and through a device buffer + copy into pageable ones: both must equal the device-buffer call
    bpp_engine_play_stub, including the zero rows / -1 actions of moves a game does not play, and the C oracle."""
    import torch
    from oracle import c_oracle as CO
    from resource_packing_self_play_b200 import _lib
    from resource_packing_self_play_b200.engine import SearchEngine
    from resource_packing_self_play_b200.game import ItemsGenerator

    W, H, N, G, SIMS = 15, 15, 10, 37, 48
    rng = np.random.RandomState(11)
    heights = rng.randint(2, 16, size=G).astype(np.int32)
    items = ItemsGenerator(W, H, N).items_batch(np.arange(G) + 7000, heights)
    area = (W * heights).astype(np.int32)
    bl = np.full(G, np.nan)
    eng = SearchEngine(W, H, N, G, SIMS, 1.0, device=0)
    eng.reset(items, area, bl)
    counts_d, actions_d = eng.play_stub(kind, _lib.CHOOSE_ARGMAX_FIRST)
    eng.check()
    counts_d, actions_d = counts_d.cpu().numpy(), actions_d.cpu().numpy()
    moves_d = eng.status()["moves"].cpu().numpy() if "moves" in eng.status() else None

    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype).pin_memory().numpy()

    out_pin = {"counts": pinned((N, G, W * N), torch.int32), "actions": pinned((N, G), torch.int32),
               "r": pinned(G, torch.int32), "score": pinned(G, torch.float64), "moves": pinned(G, torch.int32)}
    out_pin["counts"][:] = 12345   # stale contents must be overwritten everywhere
    out_pin["actions"][:] = 777
    eng.play_stub_host(kind, items, area, bl, choose_mode=_lib.CHOOSE_ARGMAX_FIRST, out=out_pin)
    out_pag = eng.play_stub_host(kind, items, area, bl, choose_mode=_lib.CHOOSE_ARGMAX_FIRST)
    for out in (out_pin, out_pag):
        assert np.array_equal(out["counts"], counts_d)
        assert np.array_equal(out["actions"], actions_d)
        if moves_d is not None:
            assert np.array_equal(out["moves"], moves_d)
    assert (out_pin["actions"] == -1).any(), "the sample should contain episodes shorter than N moves"
    for g in range(0, G, 6):
        ref = CO.play_episode(W, H, N, items[g], int(area[g]), float("nan"), kind, SIMS, 1.0, policy=0)
        m = ref["moves"]
        assert int(out_pin["moves"][g]) == m
        assert np.array_equal(out_pin["counts"][:m, g], ref["counts"])
        assert not out_pin["counts"][m:, g].any() and (out_pin["actions"][m:, g] == -1).all()
        assert (int(out_pin["r"][g]), float(out_pin["score"][g])) == (ref["r"], ref["score"])


def test_batched_mcts_fused_graph_replay_equals_plain_lockstep():
    """BatchedMCTS.search with the fused expand+select launch replayed from CUDA graphs must give exactly the visit counts
    of the plain select -> forward -> expand_backup loop launched eagerly (games are independent and the evaluator is
    deterministic per leaf, so neither the launch structure nor the leaf order may change a count)."""
    import torch
    from resource_packing_self_play_b200 import _lib
    from resource_packing_self_play_b200.game import ItemsGenerator
    from resource_packing_self_play_b200.mcts import BatchedMCTS
    from resource_packing_self_play_b200.nnet import NNetWrapper
    from resource_packing_self_play_b200.utils import dotdict

    W, H, N, G, SIMS = 15, 15, 10, 96, 40

    class Gm:
        bin_width, bin_height, num_items = W, H, N

        def getBoardSize(self):
            return (H, W)

        def getActionSize(self):
            return W * N
    args = dotdict(numMCTSSims=SIMS, cpuct=1.0, alpha=0.75, num_items=N, num_bins=1, cuda=True)
    torch.manual_seed(3)
    net = NNetWrapper(Gm(), args, max_batch=G, precision="bf16")
    rng = np.random.RandomState(5)
    heights = rng.randint(4, 16, size=G).astype(np.int32)
    items = ItemsGenerator(W, H, N).items_batch(np.arange(G) + 31000, heights)
    area = (W * heights).astype(np.int32)
    tie = np.ones(G, dtype=np.int8)
    def play(bm, moves):
        bm.reset(items, area, [0.5, 0.7, 0.9001], tie=tie)
        per_move = []
        for m in range(moves):
            counts = bm.search(chunk=4)
            act = bm.eng.choose(_lib.CHOOSE_ARGMAX_FIRST)
            bm.eng.advance(act)
            per_move.append((counts.cpu().numpy().copy(), act.cpu().numpy().copy()))
        bm.eng.check()
        return per_move
    runs, bms = [], []
    for fused, graphs in ((True, True), (False, False), (True, False)):
        bm = BatchedMCTS(Gm(), net, args, G)
        bm.fused, bm.use_graphs = fused, graphs
        runs.append(play(bm, 6))
        bms.append(bm)
    assert bms[0].graph_launches > 0 and bms[1].graph_launches == 0, "the first run must have replayed graphs"
    for other in runs[1:]:
        for (c0, a0), (c1, a1) in zip(runs[0], other):
            assert np.array_equal(c0, c1) and np.array_equal(a0, a1)
    assert runs[0][0][0].sum() > 0
    # new weights (as after a learner step) must reach the ALREADY CAPTURED graphs: every parameter lives in device
    # buffers that are updated in place, nothing that changes is a by-value kernel argument
    with torch.no_grad():
        for prm in net.nnet.parameters():
            prm.add_(0.05 * torch.randn_like(prm))
        net.nnet.value_fc.bias.add_(0.4)
    net.sync_weights()
    before = bms[0].graph_launches
    after_graph, after_eager = play(bms[0], 3), play(bms[1], 3)
    assert bms[0].graph_launches > before
    for (c0, a0), (c1, a1) in zip(after_graph, after_eager):
        assert np.array_equal(c0, c1) and np.array_equal(a0, a1)
    assert any(not np.array_equal(x[0], y[0]) for x, y in zip(after_graph, runs[0])), "the new weights changed nothing?"
