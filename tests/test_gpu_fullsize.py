"""-m gpu: BASELINE.json configs[1] at FULL size (4,096 lockstep games, 15x15 bin, 10 items, 200 simulations per move):
(a) a sample of 96 games is replayed move by move through the C oracle with the sampled actions forced, and the
    per-move visit-count matrices must be bit-identical;
(b) size-independent properties are checked on all 4,096 games."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as CO

pytestmark = pytest.mark.gpu
W, H, N, G, SIMS = 15, 15, 10, 4096, 200


@pytest.mark.parametrize("stub", ["U", "V"])
def test_4096_games_sampled_self_play(stub):
    from resource_packing_self_play_b200 import _lib
    from resource_packing_self_play_b200.engine import EnvOps, SearchEngine
    from resource_packing_self_play_b200.game import ItemsGenerator
    idx = np.arange(G)
    seeds = 1000 + idx
    heights = np.array([np.random.RandomState(77000 + int(b)).randint(2, 16) for b in idx // 20], dtype=np.int32)
    items = ItemsGenerator(W, H, N).items_batch(seeds, heights)
    area = (W * heights).astype(np.int32)
    bl = np.full(G, np.nan) if stub == "U" else np.full(G, 0.7001)
    eng = SearchEngine(W, H, N, G, SIMS, 1.0)
    eng.reset(items, area, bl)
    ops = EnvOps(W, H, N)
    roots, counts, actions = [], [], []
    for mv in range(N):
        roots.append(eng.roots())
        eng.begin_move()
        eng.search_stub(stub)
        counts.append(eng.root_counts())
        a = eng.choose(_lib.CHOOSE_SAMPLE, seed=99)
        actions.append(a)
        eng.advance(a)
    eng.check()
    st = {k: v.cpu().numpy() for k, v in eng.status().items()}
    counts_t = torch.stack(counts)                      # (N, G, A)
    actions_np = torch.stack(actions).cpu().numpy()     # (N, G)
    counts_np = counts_t.cpu().numpy()
    moves = st["moves"]
    assert (st["done"] == 1).all() and moves.min() >= 1 and moves.max() <= N
    # ---- (b) properties on every game
    played = np.arange(N)[:, None] < moves[None, :]
    sums = counts_np.sum(axis=2)
    assert (sums[0] == SIMS - 1).all()                  # a fresh tree spends its first simulation expanding the root
    # later moves search a reused tree: the root's edge counts also hold the visits of earlier moves' searches
    assert ((sums[1:] >= SIMS) | ~played[1:]).all() and (sums[~played] == 0).all()
    assert (actions_np[~played] == -1).all() and (actions_np[played] >= 0).all()
    for mv in range(N):                                 # visits only on legal actions; the sampled action was visited
        valid = ops.valid_moves(roots[mv], items).cpu().numpy().astype(bool)
        assert not (counts_np[mv][~valid] > 0).any()
        g_idx = np.flatnonzero(played[mv])
        assert (counts_np[mv][g_idx, actions_np[mv][g_idx]] > 0).all()
    final = eng.roots()
    assert not bool(ops.valid_moves(final, items).any())  # an episode ends exactly when no legal move is left
    occ = final.cpu().numpy().view(np.uint32)[:, :H]
    pop = np.array([[bin(int(v)).count("1") for v in row] for row in occ]).sum(axis=1)
    assert (pop <= area).all()
    assert ((st["score"] == 0) == (pop != area)).all() and (st["score"] <= 1.0).all()
    if stub == "U":
        assert (st["r"] == 1).all()                     # empty rewards list: every terminal state is a win
    else:
        assert ((st["r"] == 1) == ((st["score"] > 0.7001) | (st["score"] == 1.0))).all()
    s = eng.stats()
    assert s["sims"] == SIMS * int(moves.sum())
    # ---- (a) bit-exact replay of a sample through the oracle
    sample = np.linspace(0, G - 1, 96).astype(int)
    for g in sample:
        m = int(moves[g])
        ref = CO.play_episode(W, H, N, items[g], int(area[g]), float(bl[g]), stub, SIMS, 1.0, policy=2,
                              forced=actions_np[:, g].clip(min=0))
        assert ref["moves"] == m
        assert np.array_equal(ref["counts"], counts_np[:m, g]), f"game {g}: visit counts differ from the oracle"
        assert (ref["r"], ref["score"]) == (int(st["r"][g]), float(st["score"][g]))
