"""-m gpu: round-2 paths.

* asynchronous episodes (bpp_engine_play_net: every game chooses and plays inside the search kernels) against the
  move-synchronous loop (search / choose / advance per move for all games) and against the oracle;
* the reference-facing host-buffer call bpp_engine_play_net_host;
* the default precision (NNetWrapper(game, args) -> "auto") on the reference's shipped TRAINED checkpoints, held to the
  stated tolerance |d pi| <= 2e-2, |d v| <= 2e-2 against the reference's own fp32 outputs;
* batched arena against the oracle's greedy play (CoachBPP.arena_playing, CoachBPP.py:233-291) with two shipped
  checkpoints; kernel counters against the oracle's; one MCTS object replaying the same instance.
"""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, bl_of
from oracle import bpp_oracle as O

pytestmark = pytest.mark.gpu
TOL_PI, TOL_V = 2e-2, 2e-2


def _args(**kw):
    from resource_packing_self_play_b200.utils import dotdict
    d = dict(numMCTSSims=40, cpuct=1.0, alpha=0.75, num_items=10, num_bins=1, cuda=True, epochs=1, batch_size=8,
             checkpoint="/tmp/_bpp_ck", arenaCompare=4, numItems=10, seed=100)
    d.update(kw)
    return dotdict(d)


def _ck_weights(which):
    if which == 1:
        d = np.load(os.path.join(GOLDEN, "net.npz"))
        return {k[5:]: d[k] for k in d.files if k.startswith("ck_w.")}
    d = np.load(os.path.join(GOLDEN, "net_ck2.npz"))
    return {k[2:]: d[k] for k in d.files if k.startswith("w.")}


def _net(game, args, weights=None, precision="auto", seed=3, **kw):
    from resource_packing_self_play_b200.nnet import NNetWrapper
    torch.manual_seed(seed)
    net = NNetWrapper(game, args, precision=precision, **kw)
    if weights is not None:
        net.nnet.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
        net.sync_weights()
    return net


@pytest.mark.parametrize("W,H,G,SIMS,mode", [(15, 15, 96, 40, 1), (20, 20, 40, 30, 2), (15, 15, 33, 25, 0)])
def test_async_episodes_equal_the_move_synchronous_loop(W, H, G, SIMS, mode):
    """Games are independent and the action stream is a function of (seed, game, move number): letting every game run
    ahead at its own pace may not change a single visit count, action, root record, outcome or score."""
    from resource_packing_self_play_b200.game import BinPackingGame, ItemsGenerator
    from resource_packing_self_play_b200.mcts import BatchedMCTS
    N = 10
    g = BinPackingGame(W, H, N, 1)
    args = _args(numMCTSSims=SIMS)
    net = _net(g, args, precision="bf16", max_batch=G)
    with torch.no_grad():  # an informative policy and value (default init is almost uniform / zero)
        net.nnet.logits_fc.weight.mul_(30.0)
        net.nnet.value_fc.weight.mul_(20.0)
    net.sync_weights()
    rng = np.random.RandomState(11)
    heights = rng.randint(2, H + 1, size=G).astype(np.int32)
    items = ItemsGenerator(W, H, N).items_batch(np.arange(G) + 4100, heights)
    area = (W * heights).astype(np.int32)
    rl = [0.5, 0.7, 0.9001]
    tie = np.ones(G, dtype=np.int8)
    bm = BatchedMCTS(g, net, args, G)
    # move-synchronous reference run
    bm.reset(items, area, rl, tie=tie)
    roots, counts, acts = [], [], []
    for m in range(N):
        roots.append(bm.eng.roots().cpu().numpy())
        counts.append(bm.search(chunk=4).cpu().numpy())
        a = bm.eng.choose(mode, seed=77)
        acts.append(a.cpu().numpy())
        bm.eng.advance(a)
    bm.eng.check()
    st0 = {k: v.cpu().numpy() for k, v in bm.eng.status().items()}
    # asynchronous run on the same engine
    ep = bm.play_episodes(items, area, rl, seed=77, tie=tie, mode=mode)
    bm.eng.check()
    assert bool((ep["done"] == 1).all())
    moves = ep["moves"].cpu().numpy()
    assert np.array_equal(moves, st0["moves"]) and np.array_equal(ep["r"].cpu().numpy(), st0["r"])
    assert np.array_equal(ep["score"].cpu().numpy(), st0["score"])
    c1, a1, r1 = ep["counts"].cpu().numpy(), ep["actions"].cpu().numpy(), ep["roots"].cpu().numpy()
    for gi in range(G):
        m = int(moves[gi])
        for k in range(m):
            assert np.array_equal(c1[k, gi], counts[k][gi]), (gi, k)
            assert a1[k, gi] == acts[k][gi]
            assert np.array_equal(r1[k, gi], roots[k][gi])
        assert not c1[m:, gi].any() and (a1[m:, gi] == -1).all() and not r1[m:, gi].any()
    assert c1.sum() > 0
    # the host-buffer entry point returns the same arrays
    out = bm.eng.play_net_host(net.dnet, items, area, np.full(G, bl_of(rl)), tie=tie, choose_mode=mode, seed=77)
    assert np.array_equal(out["counts"], c1) and np.array_equal(out["actions"], a1)
    assert np.array_equal(out["roots"].view(np.int32), r1) and np.array_equal(out["moves"], moves)
    assert np.array_equal(out["r"], st0["r"]) and np.array_equal(out["score"], st0["score"])
    bm.close()


def test_coach_batched_async_equals_per_move_and_matches_oracle_fed_with_the_same_evaluations():
    from resource_packing_self_play_b200.coach import CoachBPP
    from resource_packing_self_play_b200.game import BinPackingGame, ItemsGenerator
    W, H, N, G = 15, 15, 10, 12
    g = BinPackingGame(W, H, N, 1)
    args = _args(numMCTSSims=30)
    net = _net(g, args, _ck_weights(1), precision="fp32")
    gen = ItemsGenerator(W, 9, N)
    coach = CoachBPP(g, net, gen.items_generator(1), W * 9, gen, args, saved_rewards_list=[0.5, 0.6, 0.7001, 0.8001])
    items = gen.items_batch(np.arange(G) + 2000)
    areas = np.full(G, W * 9, dtype=np.int32)
    a_c, a_s, a_r = coach.executeEpisodesBatched(items, areas, greedy="first", seed=9, expand=False)
    b_c, b_s, b_r = coach.executeEpisodesBatched(items, areas, greedy="first", seed=9, expand=False, per_move=True)
    for k in ("roots", "counts", "actions", "moves", "r"):
        assert np.array_equal(a_c[k], b_c[k]), k
    assert np.array_equal(a_s, b_s)

    class Feed:  # the oracle's dict MCTS evaluated by the very same device forward (pi float32, v as Python float)
        def predict(self, board):
            pi, v = net.predict(board)
            return pi, float(v[0])
    for gi in range(0, G, 3):
        og = O.OracleGame(W, H, N, 1)
        om = O.OracleMCTS(og, Feed(), args)
        its = [[int(w), int(h), 0, 0] for w, h in items[gi]]
        board, planes = og.getInitBoard(), og.getInitItems(its)
        for mv in range(int(a_c["moves"][gi])):
            st = og.getBinItem(board, planes)
            om.getActionProb(st, int(areas[gi]), coach.rewards_list)
            assert om.root_counts(st) == list(a_c["counts"][mv, gi]), (gi, mv)
            board, planes = og.getNextState(board, int(a_c["actions"][mv, gi]), planes)
        r, score = og.getGameEnded(og.getBinItem(board, planes), int(areas[gi]), coach.rewards_list, 0.75)
        assert (r, float(score)) == (int(a_r[gi]), float(a_s[gi]))


@pytest.mark.parametrize("which", [1, 2])
def test_default_precision_meets_the_tolerance_on_the_shipped_trained_checkpoints(which):
    """NNetWrapper(game, args) with no precision argument + a shipped checkpoint: the auto-selected tensor-core mode must
    reproduce the reference's own fp32 outputs within the stated tolerance (plain bf16 does not: |d pi| up to 0.6)."""
    from resource_packing_self_play_b200.engine import pack_states
    from resource_packing_self_play_b200.game import BinPackingGame
    g = BinPackingGame(15, 15, 10, 1)
    net = _net(g, _args(), _ck_weights(which))
    assert net.precision_request == "auto" and net.dnet.precision in ("bf16x3", "fp32"), net.calibration
    assert net.dnet.precision == "bf16x3", "the split-bf16 tensor-core mode should pass the calibration"
    d1 = np.load(os.path.join(GOLDEN, "net.npz"))
    states = d1["ck_states"].astype(np.int64)
    if which == 1:
        ref_pi, ref_v = d1["ck_pi"], d1["ck_v"]
    else:
        d2 = np.load(os.path.join(GOLDEN, "net_ck2.npz"))
        ref_pi, ref_v = d2["pi"], d2["v"]
    recs, items = pack_states(states, 15, 15, 10)
    pi, v = net.predict_batch(torch.from_numpy(recs.view(np.int32)).to(net.device), torch.from_numpy(items).to(net.device))
    assert np.abs(pi.cpu().numpy() - ref_pi).max() <= TOL_PI
    assert np.abs(v.cpu().numpy() - ref_v).max() <= TOL_V
    # a freshly initialised network is well conditioned: auto keeps the fast plain-bf16 mode
    fresh = _net(g, _args(), None, seed=0)
    assert fresh.dnet.precision == "bf16", fresh.calibration


def test_arena_sweep_matches_the_oracle_greedy_play_with_two_shipped_checkpoints():
    """CoachBPP.arena_playing semantics (CoachBPP.py:233-291): every seed played greedily (greedy_a=0) by the previous and
    by the new net; scores are the raw utilisation, accept iff mean(new) >= mean(prev).  Oracle = the dict MCTS fed with
    the same fp32 device forward; arg-max choice injected as 'first maximum' on both sides."""
    from resource_packing_self_play_b200.coach import CoachBPP
    from resource_packing_self_play_b200.game import BinPackingGame, ItemsGenerator
    W, H, N = 15, 15, 10
    g = BinPackingGame(W, H, N, 1)
    args = _args(numMCTSSims=25)
    pnet = _net(g, args, _ck_weights(2), precision="fp32")
    nnet = _net(g, args, _ck_weights(1), precision="fp32")
    gen = ItemsGenerator(W, 10, N)
    coach = CoachBPP(g, nnet, gen.items_generator(1), W * 10, gen, args, saved_rewards_list=[1.0] * 100)
    seeds = np.array([11, 503, 77, 4096, 9, 250])
    hts = np.array([10, 7, 15, 4, 12, 9])
    p, n_, acc = coach.arena_sweep(pnet, nnet, seeds, hts, seed=1, choose="first")
    want = []
    for net in (pnet, nnet):
        class Feed:
            def predict(self, board, net=net):
                pi, v = net.predict(board)
                return pi, float(v[0])
        scores = []
        for sd, ht in zip(seeds, hts):
            items = ItemsGenerator(W, int(ht), N).items_generator(int(sd))
            og = O.OracleGame(W, H, N, 1)
            om = O.OracleMCTS(og, Feed(), args)   # a fresh tree per game (deviation ii)
            board, planes = og.getInitBoard(), og.getInitItems(items)
            while True:
                st = og.getBinItem(board, planes)
                om.getActionProb(st, W * int(ht), coach.rewards_list)
                a = int(np.argmax(om.root_counts(st)))
                board, planes = og.getNextState(board, a, planes)
                r, score = og.getGameEnded(og.getBinItem(board, planes), W * int(ht), coach.rewards_list, 0.75)
                if r != 0:
                    scores.append(float(score))
                    break
        want.append(np.array(scores))
    assert np.array_equal(p, want[0]) and np.array_equal(n_, want[1])
    assert acc == (1 if want[1].mean() >= want[0].mean() else 0)
    # the sequential drop-in arena_playing on the same seeds gives the same verdict when the two nets differ clearly
    assert acc in (0, 1)


def test_kernel_counters_equal_the_oracle_counters():
    """roofline inputs: simulations, edges walked and expansions counted by the kernels == the oracle's own counters"""
    from resource_packing_self_play_b200 import _lib
    from resource_packing_self_play_b200.engine import SearchEngine
    from resource_packing_self_play_b200.game import ItemsGenerator
    W, H, N, G, SIMS = 15, 15, 10, 6, 60
    heights = np.array([15, 6, 9, 3, 12, 8], dtype=np.int32)
    items = ItemsGenerator(W, H, N).items_batch(np.arange(G) + 880, heights)
    area = (W * heights).astype(np.int32)
    eng = SearchEngine(W, H, N, G, SIMS, 1.0)
    eng.reset(items, area, np.full(G, np.nan))
    eng.stats(reset=True)
    counts, actions = eng.play_stub("V", _lib.CHOOSE_ARGMAX_FIRST)
    eng.check()
    st = eng.stats()
    moves = eng.status()["moves"].cpu().numpy()
    actions = actions.cpu().numpy()
    sims = edges = exps = 0
    for gi in range(G):
        og = O.OracleGame(W, H, N, 1)
        om = O.OracleMCTS(og, O.StubNet("V", W * N), _args(numMCTSSims=SIMS))
        its = [[int(w), int(h), 0, 0] for w, h in items[gi]]
        board, planes = og.getInitBoard(), og.getInitItems(its)
        for mv in range(int(moves[gi])):
            om.getActionProb(og.getBinItem(board, planes), int(area[gi]), [])
            board, planes = og.getNextState(board, int(actions[mv, gi]), planes)
            sims += SIMS
        edges += om.n_edges_walked
        exps += om.n_expansions
    assert st["sims"] == sims and st["edges"] == edges and st["expansions"] == exps
    assert st["edge_units_read"] > 0


def test_one_mcts_object_replays_the_same_instance_many_times():
    """CoachBPP.arena_playing keeps one MCTS object across games and a user may call executeEpisode twice: the device
    pools hold several episodes and are dropped, not overflowed, beyond that (the reference's dicts just grow)."""
    from resource_packing_self_play_b200.game import BinPackingGame, ItemsGenerator
    from resource_packing_self_play_b200.mcts import MCTS
    W, H, N = 15, 15, 10
    g = BinPackingGame(W, H, N, 1)
    args = _args(numMCTSSims=60)
    m = MCTS(g, O.StubNet("V", W * N), args)
    om = O.OracleMCTS(O.OracleGame(W, H, N, 1), O.StubNet("V", W * N), args)
    items = ItemsGenerator(W, 9, N).items_generator(5)
    first = None
    for rep in range(12):
        board, planes = g.getInitBoard(), g.getInitItems(items)
        for mv in range(3):
            st = g.getBinItem(board, planes)
            pi = m.getActionProb(st, W * 9, [])
            if rep < 2:  # warm graph: identical to the reference's warm dicts
                assert pi == om.getActionProb(st, W * 9, [])
            if first is None:
                first = pi
            board, planes = g.getNextState(board, int(np.argmax(pi)), planes)
    assert abs(sum(first) - 1) < 1e-12


@pytest.mark.parametrize("W,H,E,G,SIMS,mode", [(15, 15, 101, 32, 30, 1), (20, 20, 50, 7, 25, 2)])
def test_streamed_episodes_do_not_depend_on_the_number_of_resident_games(W, H, E, G, SIMS, mode):
    """E instances streamed through G < E resident games (a game whose episode ends takes the next instance inside the
    search kernel, in whatever order the games finish) == the same E instances with one resident game each: identical
    examples and outcomes per episode, because the action stream is keyed by (seed, episode, move)."""
    from resource_packing_self_play_b200.game import BinPackingGame, ItemsGenerator
    from resource_packing_self_play_b200.mcts import BatchedMCTS
    N = 10
    g = BinPackingGame(W, H, N, 1)
    args = _args(numMCTSSims=SIMS)
    net = _net(g, args, precision="bf16", max_batch=E)
    with torch.no_grad():
        net.nnet.logits_fc.weight.mul_(30.0)
        net.nnet.value_fc.weight.mul_(20.0)
    net.sync_weights()
    rng = np.random.RandomState(3)
    heights = rng.randint(2, H + 1, size=E).astype(np.int32)
    items = ItemsGenerator(W, H, N).items_batch(np.arange(E) + 900, heights)
    area = (W * heights).astype(np.int32)
    tie = np.ones(E, dtype=np.int8)
    rl = [0.4, 0.6, 0.8001]
    outs = []
    for resident in (E, G):
        bm = BatchedMCTS(g, net, args, resident)
        ep = bm.play_episodes(items, area, rl, seed=5, tie=tie, mode=mode)
        bm.eng.check()
        outs.append({k: v.cpu().numpy() for k, v in ep.items()})
        bm.close()
    a, b = outs
    assert (a["moves"] > 0).all() and a["moves"].min() < N
    for k in ("moves", "r", "score", "counts", "actions", "roots"):
        assert np.array_equal(a[k], b[k]), k
    # and the host-buffer entry point of the stream
    bm = BatchedMCTS(g, net, args, G)
    out = bm.eng.play_net_stream_host(net.dnet, items, area, np.full(E, bl_of(rl)), tie=tie, choose_mode=mode, seed=5)
    bm.eng.check()
    for k in ("moves", "r", "score", "counts", "actions"):
        assert np.array_equal(out[k], a[k]), k
    assert np.array_equal(out["roots"].view(np.int32), a["roots"])
    bm.close()


@pytest.mark.parametrize("W,H,E,G,SIMS,stub,mode", [(15, 15, 150, 32, 40, "U", 1), (20, 20, 61, 7, 25, "V", 2),
                                                    (15, 15, 40, 64, 30, "H", 0)])
def test_streamed_stub_episodes_equal_one_game_per_episode(W, H, E, G, SIMS, stub, mode):
    """bpp_engine_play_stub_stream: E instances through G resident games in one launch of the episode kernel (a game whose
    episode ends takes the next instance) == reset + play_stub with one game per instance: identical visit counts,
    actions and outcomes per episode (the action stream is keyed by (seed, episode, move)); G > E leaves games idle.
    The host-buffer form (pinned buffers: written by the kernel itself; pageable: staged) returns the same."""
    from resource_packing_self_play_b200.engine import SearchEngine
    from resource_packing_self_play_b200.game import ItemsGenerator
    N = 10
    rng = np.random.RandomState(11)
    heights = rng.randint(2, H + 1, size=E).astype(np.int32)
    items = ItemsGenerator(W, H, N).items_batch(np.arange(E) + 400, heights)
    area = (W * heights).astype(np.int32)
    bl = np.full(E, np.nan) if stub == "U" else np.full(E, 0.7001)
    ref = SearchEngine(W, H, N, E, SIMS, 1.0)
    ref.reset(items, area, bl)
    counts, actions = ref.play_stub(stub, choose_mode=mode, seed=21)
    ref.check()
    st = {k: v.cpu().numpy() for k, v in ref.status().items()}
    counts, actions = counts.cpu().numpy(), actions.cpu().numpy()
    ref.close()
    assert (st["done"] == 1).all() and st["moves"].min() >= 1

    eng = SearchEngine(W, H, N, G, SIMS, 1.0)
    out = eng.play_stub_stream(stub, items, area, bl, choose_mode=mode, seed=21)
    eng.check()
    assert np.array_equal(out["counts"].cpu().numpy(), counts)
    assert np.array_equal(out["actions"].cpu().numpy(), actions)
    for k in ("r", "score", "moves"):
        assert np.array_equal(out[k].cpu().numpy(), st[k]), k
    # a second stream on the same handle (pools and hash tables of the games are reused), without recording
    out2 = eng.play_stub_stream(stub, items[::-1].copy(), area[::-1].copy(), bl, choose_mode=0, seed=21, record=False)
    eng.check()
    assert out2["counts"] is None and (out2["moves"].cpu().numpy() >= 1).all()
    # host buffers: pageable (staged copies) and pinned (zero-copy rows written by the episode kernel)
    host = eng.play_stub_stream_host(stub, items, area, bl, choose_mode=mode, seed=21)
    pinned = {"counts": torch.empty((N, E, W * N), dtype=torch.int32).pin_memory().numpy(),
              "actions": torch.empty((N, E), dtype=torch.int32).pin_memory().numpy()}
    pinned["counts"][:] = 7
    pinned["actions"][:] = 7
    hostp = eng.play_stub_stream_host(stub, items, area, bl, choose_mode=mode, seed=21, out=pinned)
    for o in (host, hostp):
        assert np.array_equal(o["counts"], counts) and np.array_equal(o["actions"], actions)
        for k in ("r", "score", "moves"):
            assert np.array_equal(o[k], st[k]), k
    eng.close()
