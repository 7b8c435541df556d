"""-m gpu: edge geometries against the C oracle — the maximum supported size (W=32, H=28, N=16: A = 512 actions,
16 mask words, a 4-leaf numpy summation plan, full 32-bit row masks), tiny bins, single-item episodes, ragged batches
(games of one engine ending after different numbers of moves)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle.np_sum import sum_plan

pytestmark = pytest.mark.gpu


def _random_instances(rng, W, H, N, G):
    """guillotine-free random items (any w<=W, h<=H): harsher than the generator's exact tilings"""
    items = np.stack([np.stack([rng.randint(1, W + 1, size=N), rng.randint(1, H + 1, size=N)], axis=1)
                      for _ in range(G)]).astype(np.int32)
    area = np.minimum(items[:, :, 0] * items[:, :, 1], 10 ** 6).sum(axis=1).astype(np.int32)
    return items, area


@pytest.mark.parametrize("W,H,N,sims,stub", [(32, 28, 16, 40, "H"), (32, 28, 16, 25, "D"), (3, 2, 2, 30, "V"),
                                             (1, 1, 1, 5, "U"), (17, 9, 13, 50, "H"), (20, 20, 10, 60, "D")])
def test_search_matches_oracle_on_edge_geometries(W, H, N, sims, stub):
    from resource_packing_self_play_b200.engine import SearchEngine
    rng = np.random.RandomState(W * 1000 + H * 10 + N)
    G = 24
    items, area = _random_instances(rng, W, H, N, G)
    if W * N > 128:
        assert len(sum_plan(W * N)[0]) >= 2  # exercises the multi-leaf pairwise-sum plan
    bl = np.where(np.arange(G) % 3 == 0, np.nan, 0.4501)
    eng = SearchEngine(W, H, N, G, sims, 1.25)
    eng.reset(items, area, bl)
    counts, actions = eng.play_stub(stub, 0)
    eng.check()
    st = {k: v.cpu().numpy() for k, v in eng.status().items()}
    counts, actions = counts.cpu().numpy(), actions.cpu().numpy()
    assert len(set(st["moves"].tolist())) >= (1 if N == 1 else 2) or N <= 2  # ragged: different episode lengths
    for g in range(G):
        ref = CO.play_episode(W, H, N, items[g], int(area[g]), float(bl[g]), stub, sims, 1.25, policy=0)
        m = ref["moves"]
        assert m == st["moves"][g]
        assert np.array_equal(counts[:m, g], ref["counts"]), f"game {g}"
        assert list(actions[:m, g]) == list(ref["actions"])
        assert (int(st["r"][g]), float(st["score"][g])) == (ref["r"], ref["score"])


@pytest.mark.parametrize("W,H,N", [(32, 28, 16), (1, 1, 1), (2, 28, 3), (32, 1, 16)])
def test_env_ops_match_oracle_on_edge_geometries(W, H, N):
    from resource_packing_self_play_b200.engine import EnvOps
    rng = np.random.RandomState(W + 100 * H + 10000 * N)
    n = 300
    items, _ = _random_instances(rng, W, H, N, n)
    recs = np.zeros((n, 32), dtype=np.uint32)
    dens = rng.rand(n, 1)
    bits = (rng.rand(n, H, W) < dens[:, :, None]).astype(np.uint64)
    recs[:, :H] = (bits << np.arange(W, dtype=np.uint64)).sum(axis=2).astype(np.uint32)
    recs[:, 28] = rng.randint(0, 1 << N, size=n)
    recs[:5, :H] = 0                     # empty bins
    recs[5:8, :H] = (1 << W) - 1 if W < 32 else 0xFFFFFFFF   # full bins
    ops = EnvOps(W, H, N)
    valid = ops.valid_moves(recs, items).cpu().numpy()
    assert np.array_equal(valid, CO.valid_moves(W, H, N, recs, items))
    rem_items = [np.flatnonzero((recs[i, 28] >> np.arange(N)) & 1) for i in range(n)]
    sel = np.array([i for i in range(n) if len(rem_items[i])])
    acts = np.array([rng.choice(rem_items[i]) * W + rng.randint(W) for i in sel], dtype=np.int32)
    nxt = ops.next_state(recs[sel], items[sel], acts).cpu().numpy().view(np.uint32)
    assert np.array_equal(nxt, CO.next_state(W, H, N, recs[sel], items[sel], acts))
    area = rng.randint(0, W * H + 1, size=n).astype(np.int32)
    pop = np.array([[bin(int(v)).count("1") for v in row[:H]] for row in recs]).sum(axis=1)
    area[::2] = pop[::2]                 # half of the states satisfy popcount == total_area
    max_h = rng.randint(1, H + 1, size=n).astype(np.int32)
    bl = np.where(rng.rand(n) < 0.3, np.nan, rng.rand(n))
    tie = np.where(rng.rand(n) < 0.5, 1, -1).astype(np.int8)
    ended, score = ops.game_ended(recs, items, area, max_h, bl, tie)
    e2, s2 = CO.game_ended(W, H, N, recs, items, area, max_h, bl, tie)
    assert np.array_equal(ended.cpu().numpy(), e2)
    term = e2 != 0
    assert np.array_equal(score.cpu().numpy()[term], s2[term])


def test_capacity_overflow_is_reported_not_hidden():
    from resource_packing_self_play_b200 import _lib
    from resource_packing_self_play_b200.engine import SearchEngine
    from resource_packing_self_play_b200.game import ItemsGenerator
    items = ItemsGenerator(15, 15, 10).items_batch([1, 2, 3, 4], [15, 15, 15, 15])
    eng = SearchEngine(15, 15, 10, 4, 200, 1.0, node_cap=40)
    eng.reset(items, np.full(4, 225, dtype=np.int32), np.full(4, np.nan))
    eng.play_stub("U", 0)
    with pytest.raises(_lib.BppError) as ei:
        eng.check()
    assert ei.value.code == -4 and "overflowed" in str(ei.value)
