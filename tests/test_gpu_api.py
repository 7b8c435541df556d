"""-m gpu: the reference-facing Python API (Game / MCTS / NNetWrapper / CoachBPP mirrors) used the way the reference's
own callers use it (CoachBPP.executeEpisode, MCTS.getActionProb, ...), checked against the reference fixtures and
the oracle."""
import numpy as np
import pytest
import torch

from helpers import bl_of, load_env_golden, load_mcts_golden
from oracle import bpp_oracle as O

pytestmark = pytest.mark.gpu
CASES = load_mcts_golden()


def _args(**kw):
    from resource_packing_self_play_b200.utils import dotdict
    d = dict(numMCTSSims=200, cpuct=1, alpha=0.75, num_items=10, num_bins=1, cuda=True, epochs=1, batch_size=8,
             seed=100, numItems=10)
    d.update(kw)
    return dotdict(d)


def test_game_single_state_methods_match_oracle():
    from resource_packing_self_play_b200.game import BinPackingGame, ItemsGenerator
    W, H, N = 15, 15, 10
    g, og = BinPackingGame(W, H, N, 1), O.OracleGame(W, H, N, 1)
    assert g.getBoardSize() == (H, W) and g.getActionSize() == W * N
    items = ItemsGenerator(W, 9, N).items_generator(4242)
    assert [list(map(int, it)) for it in items] == [list(map(int, it)) for it in
                                                    O.OracleItemsGenerator(W, 9, N).items_generator(4242)]
    board, planes = g.getInitBoard(), g.getInitItems(items)
    oboard, oplanes = og.getInitBoard(), og.getInitItems(items)
    assert g.max_h == og.max_h and g.sum_h == og.sum_h
    rng = np.random.RandomState(3)
    rl = [0.5, 0.6, 0.7001, 0.8001, 0.9]
    for _ in range(N):
        st, ost = g.getBinItem(board, planes), og.getBinItem(oboard, oplanes)
        assert np.array_equal(st, ost) and g.stringRepresentation(st) == og.stringRepresentation(ost)
        assert g.getGameEnded(st, W * 9, rl, 0.75) == (0, []) if og.has_valid_moves(ost) else True
        if not og.has_valid_moves(ost):
            break
        v = g.getValidMoves(st)
        assert v.dtype == np.int64 and np.array_equal(v, og.getValidMoves(ost)) and g.has_valid_moves(st)
        a = int(rng.choice(np.flatnonzero(v)))
        board, planes = g.getNextState(board, a, planes)
        oboard, oplanes = og.getNextState(oboard, a, oplanes)
        assert np.array_equal(board, oboard) and np.array_equal(np.array(planes), np.array(oplanes))
    st, ost = g.getBinItem(board, planes), og.getBinItem(oboard, oplanes)
    e, s = g.getGameEnded(st, W * 9, rl, 0.75)
    oe, os_ = og.getGameEnded(ost, W * 9, rl, 0.75)
    assert (e, float(s)) == (oe, float(os_)) and e != 0
    assert g.get_minimal_bin_height(board) == og.get_minimal_bin_height(oboard)
    with pytest.raises(AssertionError):  # no legal move: the reference asserts in getValidMoves (BinPackingGame.py:89)
        g.getValidMoves(st)


@pytest.mark.parametrize("ci", [3, 7, 16, 17])
def test_mcts_drop_in_with_python_stub_evaluator(ci):
    """MCTS(game, nnet, args).getActionProb on reference state tensors with a duck-typed Python nnet."""
    from resource_packing_self_play_b200.game import BinPackingGame
    from resource_packing_self_play_b200.mcts import MCTS
    c = CASES[ci]
    W, H, N = c["W"], c["H"], c["N"]
    g = BinPackingGame(W, H, N, 1)
    m = MCTS(g, O.StubNet(c["stub"], W * N), _args(numMCTSSims=c["sims"], cpuct=c["cpuct"], alpha=c["alpha"]))
    board, planes = g.getInitBoard(), g.getInitItems(c["items"])
    area = c["genW"] * c["genH"]
    for mv, a in enumerate(c["actions"]):
        state = g.getBinItem(board, planes)
        pi = m.getActionProb(state, area, c["rewards"])
        want = np.array(c["counts"][mv], dtype=np.float64)
        assert m.root_counts() == c["counts"][mv]
        assert pi == list(want / want.sum())
        board, planes = g.getNextState(board, a, planes)
    r, score = g.getGameEnded(g.getBinItem(board, planes), area, c["rewards"], c["alpha"])
    assert (r, float(score)) == (c["r"], c["score"])


def test_mcts_search_returns_the_reference_values():
    """MCTS.search() is one simulation and returns its backed-up value (MCTS_bpp.py:83,104,139)."""
    from resource_packing_self_play_b200.game import BinPackingGame
    from resource_packing_self_play_b200.mcts import MCTS
    c = CASES[4]
    W, H, N = c["W"], c["H"], c["N"]
    g, og = BinPackingGame(W, H, N, 1), O.OracleGame(W, H, N, 1)
    args = _args(numMCTSSims=c["sims"], cpuct=c["cpuct"])
    m, om = MCTS(g, O.StubNet("V", W * N), args), O.OracleMCTS(og, O.StubNet("V", W * N), args)
    state = g.getBinItem(g.getInitBoard(), g.getInitItems(c["items"]))
    og.getInitItems(c["items"])
    area = c["genW"] * c["genH"]
    got = [m.search(state, area, []) for _ in range(60)]
    want = [float(om.search(state, area, [])) for _ in range(60)]
    assert got == want
    assert m.root_counts() == om.root_counts(state)


def test_greedy_action_prob_is_one_hot_on_a_maximum():
    from resource_packing_self_play_b200.game import BinPackingGame
    from resource_packing_self_play_b200.mcts import MCTS
    c = CASES[0]
    g = BinPackingGame(15, 15, 10, 1)
    m = MCTS(g, O.StubNet("U", 150), _args(numMCTSSims=200))
    state = g.getBinItem(g.getInitBoard(), g.getInitItems(c["items"]))
    pi = m.getActionProb(state, 225, [], greedy_a=0)
    counts = np.array(c["counts"][0])
    assert sum(pi) == 1 and counts[int(np.argmax(pi))] == counts.max()


def test_coach_execute_episode_drop_in(monkeypatch):
    """CoachBPP.executeEpisode with the sampling made deterministic (arg-max) reproduces the fixture episode."""
    from resource_packing_self_play_b200.coach import CoachBPP
    from resource_packing_self_play_b200.game import BinPackingGame, ItemsGenerator
    c = CASES[6]  # V stub, ranked rewards
    W, H, N = c["W"], c["H"], c["N"]
    g = BinPackingGame(W, H, N, 1)

    class Net(O.StubNet):
        def __init__(self, game=None, args=None):
            super().__init__("V", W * N)
    args = _args(numMCTSSims=c["sims"], checkpoint="/tmp/_bpp_ck")
    gen = ItemsGenerator(c["genW"], c["genH"], N)
    coach = CoachBPP(g, Net(), gen.items_generator(c["seed"]), c["genW"] * c["genH"], gen, args,
                     saved_rewards_list=c["rewards"])
    monkeypatch.setattr(np.random, "choice", lambda n, p=None: int(np.argmax(p)))
    ex = coach.executeEpisode()
    assert len(ex) == len(c["actions"])
    for k, (state, pi, r) in enumerate(ex):
        want = np.array(c["counts"][k], dtype=np.float64)
        assert state.shape == (N + 1, H, W) and pi == list(want / want.sum()) and r == c["r"]
    assert float(coach.ep_score) == c["score"]


def test_real_net_lockstep_counts_match_oracle_fed_with_the_same_evaluations():
    """Batched search with the CUDA net as evaluator == the oracle's dict MCTS when the oracle is given the very same
    leaf evaluations (pi float32, v as a Python float so that all tree arithmetic is float64 on both sides)."""
    from resource_packing_self_play_b200.game import BinPackingGame, ItemsGenerator
    from resource_packing_self_play_b200.mcts import BatchedMCTS
    from resource_packing_self_play_b200.nnet import NNetWrapper
    W, H, N, G, SIMS = 15, 15, 10, 6, 40
    g = BinPackingGame(W, H, N, 1)
    args = _args(numMCTSSims=SIMS)
    torch.manual_seed(1)
    net = NNetWrapper(g, args)
    with torch.no_grad():  # make the random-init policy informative (default init gives an almost uniform prior)
        net.nnet.logits_fc.weight.mul_(40.0)
        net.nnet.value_fc.weight.mul_(20.0)
    net.sync_weights()
    heights = np.array([15, 7, 3, 11, 9, 5])
    items = ItemsGenerator(W, H, N).items_batch(np.arange(G) + 31, heights)
    areas = (W * heights).astype(np.int32)
    bm = BatchedMCTS(g, net, args, G)
    rl = [0.5, 0.6, 0.7001, 0.8001, 0.9]
    bm.reset(items, areas, rl)
    counts, actions = [], []
    for mv in range(3):
        c = bm.search().cpu().numpy()
        a = bm.eng.choose(0).cpu().numpy()
        counts.append(c)
        actions.append(a)
        bm.eng.advance(a)
    bm.eng.check()

    class Feed:
        def predict(self, board):
            pi, v = net.predict(board)
            return pi, float(v[0])
    for gi in range(G):
        og = O.OracleGame(W, H, N, 1)
        om = O.OracleMCTS(og, Feed(), args)
        its = [[int(w), int(h), 0, 0] for w, h in items[gi]]
        board, planes = og.getInitBoard(), og.getInitItems(its)
        for mv in range(3):
            st = og.getBinItem(board, planes)
            om.getActionProb(st, int(areas[gi]), rl)
            assert om.root_counts(st) == list(counts[mv][gi]), (gi, mv)
            board, planes = og.getNextState(board, int(actions[mv][gi]), planes)


def test_execute_episodes_batched_produces_reference_shaped_examples():
    from resource_packing_self_play_b200.coach import CoachBPP
    from resource_packing_self_play_b200.game import BinPackingGame, ItemsGenerator
    from resource_packing_self_play_b200.nnet import NNetWrapper
    W, H, N, G = 15, 15, 10, 32
    g = BinPackingGame(W, H, N, 1)
    args = _args(numMCTSSims=25, checkpoint="/tmp/_bpp_ck", arenaCompare=4)
    gen = ItemsGenerator(W, 8, N)
    net = NNetWrapper(g, args)
    coach = CoachBPP(g, net, gen.items_generator(1), W * 8, gen, args)
    items = gen.items_batch(np.arange(G) + 1000)
    ex, score, r = coach.executeEpisodesBatched(items, np.full(G, W * 8), seed=5)
    assert len(score) == G and set(np.unique(r)) <= {-1, 1}
    assert len(ex) >= 4 * G
    og = O.OracleGame(W, H, N, 1)
    for state, pi, rr in ex[:40]:
        assert state.shape == (N + 1, H, W) and state.dtype == np.int64
        assert abs(sum(pi) - 1) < 1e-9 and len(pi) == W * N and rr in (-1, 1)
        valid = og.getValidMoves(state)
        assert all(valid[a] for a in np.flatnonzero(pi))  # visit counts only on legal actions
    # batched arena: the same net against itself must be accepted (mean scores are equal up to tie-breaks in choose)
    p, n_, acc = coach.arena_sweep(net, net, np.arange(8) + 7, seed=3)
    assert len(p) == len(n_) == 8 and np.array_equal(p, n_) and acc == 1


@pytest.mark.parametrize("ci", [4, 7, 8, 15])
def test_whole_search_graph_equals_the_reference_dicts(ci):
    """Not only the root visit counts: after an episode the device graph, dumped as the reference's six dicts, equals
    the oracle's dicts entry by entry — Qsa and Ps bit for bit (float64), Nsa, Ns, Es, Vs exactly."""
    from resource_packing_self_play_b200.game import BinPackingGame
    from resource_packing_self_play_b200.mcts import MCTS
    c = CASES[ci]
    W, H, N = c["W"], c["H"], c["N"]
    g, og = BinPackingGame(W, H, N, 1), O.OracleGame(W, H, N, 1)
    args = _args(numMCTSSims=c["sims"], cpuct=c["cpuct"], alpha=c["alpha"])
    m, om = MCTS(g, O.StubNet(c["stub"], W * N), args), O.OracleMCTS(og, O.StubNet(c["stub"], W * N), args)
    board, planes = g.getInitBoard(), g.getInitItems(c["items"])
    og.getInitItems(c["items"])
    area = c["genW"] * c["genH"]
    for a in c["actions"][:4]:
        state = g.getBinItem(board, planes)
        m.getActionProb(state, area, c["rewards"])
        om.getActionProb(state, area, c["rewards"])
        board, planes = g.getNextState(board, a, planes)
    Qsa, Nsa, Ns, Ps, Es, Vs = m.Qsa, m.Nsa, m.Ns, m.Ps, m.Es, m.Vs
    assert Nsa == om.Nsa and Ns == om.Ns
    assert set(Es) == set(om.Es) and all(Es[k] == om.Es[k] for k in Es)
    assert set(Qsa) == set(om.Qsa) and all(float(Qsa[k]) == float(om.Qsa[k]) for k in Qsa)   # bit-exact float64
    assert set(Ps) == set(om.Ps)
    for k in Ps:
        assert np.array_equal(Ps[k], om.Ps[k]) and np.array_equal(Vs[k], om.Vs[k])
    assert len(Qsa) > 20
