"""CPU: pin the C oracle (oracle/bpp_oracle.c) against the reference fixtures and against numpy."""
import time

import numpy as np

from helpers import bl_of, load_env_golden, load_mcts_golden, recs_from_occ
from oracle import c_oracle as CO


def _groups(g):
    keys = sorted({(int(w), int(h), int(n)) for w, h, n in zip(g["W"], g["H"], g["N"])})
    for (W, H, N) in keys:
        yield W, H, N, np.flatnonzero((g["W"] == W) & (g["H"] == H) & (g["N"] == N))


def test_c_env_ops_match_reference_fixtures():
    g = load_env_golden()
    bls = np.array([bl_of(r) for r in g["reward_lists"]])
    for W, H, N, idx in _groups(g):
        recs = recs_from_occ(g["occ"][idx], g["rem"][idx], H)
        items = g["items_wh"][idx][:, :N, :].astype(np.int32)
        assert np.array_equal(CO.valid_moves(W, H, N, recs, items), g["valid"][idx][:, :W * N])
        ended, score = CO.game_ended(W, H, N, recs, items, g["total_area"][idx], g["max_h"][idx], bls[g["bl_case"][idx]])
        assert np.array_equal(ended, g["ended"][idx])
        assert np.array_equal(score[ended != 0], g["score"][idx][ended != 0])
        sel = g["action"][idx] >= 0
        nxt = CO.next_state(W, H, N, recs[sel], items[sel], g["action"][idx][sel])
        assert np.array_equal(nxt, recs_from_occ(g["next_occ"][idx][sel], g["next_rem"][idx][sel], H))


def test_c_mcts_matches_every_reference_fixture():
    t0 = time.time()
    for ci, c in enumerate(load_mcts_golden()):
        items = np.array([[it[0], it[1]] for it in c["items"]], dtype=np.int32)
        out = CO.play_episode(c["W"], c["H"], c["N"], items, c["genW"] * c["genH"], bl_of(c["rewards"], c["alpha"]),
                              c["stub"], c["sims"], c["cpuct"], policy=0 if c["policy"] == "argmax" else 1)
        assert list(out["actions"]) == c["actions"], ci
        assert np.array_equal(out["counts"], np.array(c["counts"])), ci
        assert (out["r"], out["score"]) == (c["r"], c["score"]), ci
        assert out["stats"]["expansions"] == c["n_expanded"] and out["stats"]["nsa_entries"] == c["n_edges"], ci
    assert time.time() - t0 < 60


def test_c_pairwise_sum_matches_numpy():
    rng = np.random.RandomState(1)
    for n in range(0, 600):
        a = rng.rand(n) * (10.0 ** rng.randint(-3, 4, size=n))
        assert CO.pairwise_sum(a) == float(np.sum(a)), n
