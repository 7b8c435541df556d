"""CPU: pin the oracle (oracle/bpp_oracle.py, oracle/np_sum.py) against fixtures produced by the unmodified
reference (tests/golden/make_golden.py).  The oracle is only trusted because these pass."""
import os
import sys

import numpy as np
import pytest

from helpers import bl_of, load_env_golden, load_items_golden, load_mcts_golden

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import bpp_oracle as O  # noqa: E402
from oracle.np_sum import pairwise_sum, sum_plan  # noqa: E402


def test_items_generator_matches_reference():
    for rec in load_items_golden():
        got = O.OracleItemsGenerator(rec["W"], rec["Hgen"], rec["N"]).items_generator(rec["seed"])
        assert [[int(v) for v in it] for it in got] == rec["items"]


def test_survey_known_answer_items_seed_100():
    items = O.OracleItemsGenerator(15, 15, 10).items_generator(100)
    assert [(it[0], it[1]) for it in items] == [(9, 11), (9, 3), (9, 1), (1, 8), (5, 7), (1, 7), (2, 3), (2, 5), (3, 6),
                                                (3, 2)]


def test_env_ops_match_reference():
    g = load_env_golden()
    n = len(g["W"])
    step = 3  # every third state keeps the CPU suite short; the GPU suite checks all of them
    for i in range(0, n, step):
        W, H, N = int(g["W"][i]), int(g["H"][i]), int(g["N"][i])
        game = O.OracleGame(W, H, N, 1)
        wh = [tuple(int(v) for v in x) for x in g["items_wh"][i][:N]]
        game.max_h = int(g["max_h"][i])
        st = O.unpack_state([int(v) for v in g["occ"][i][:H]], int(g["rem"][i]), wh, W, H)
        has = game.has_valid_moves(st) if g["rem"][i] else False
        assert int(has) == int(g["has_moves"][i])
        if has:
            assert np.array_equal(game.getValidMoves(st), g["valid"][i][:W * N])
        rl = g["reward_lists"][int(g["bl_case"][i])]
        ended, score = game.getGameEnded(st, int(g["total_area"][i]), rl, 0.75)
        assert ended == g["ended"][i]
        if ended != 0:
            assert float(score) == g["score"][i]
        a = int(g["action"][i])
        if a >= 0:
            nb, npl = game.getNextState(st[0], a, st[1:])
            occ, rem = O.pack_state(game.getBinItem(nb, npl))
            assert occ == [int(v) for v in g["next_occ"][i][:H]] and rem == int(g["next_rem"][i])


MCTS_CASES = load_mcts_golden()
# the seven SURVEY.md §8(c) known answers + one of each other kind (the rest run in the GPU suite)
CPU_SUBSET = [0, 1, 2, 3, 4, 5, 6, 7, 8, 12, 14, 15, 16, 17]


@pytest.mark.parametrize("ci", CPU_SUBSET)
def test_mcts_visit_counts_match_reference(ci):
    c = MCTS_CASES[ci]
    items = c["items"]
    mout = []
    counts, actions, r, score = O.play_episode(c["W"], c["H"], c["N"], items, c["genW"] * c["genH"], c["stub"],
                                               c["sims"], c["cpuct"], c["alpha"], c["rewards"], c["policy"], mout)
    assert actions == c["actions"]
    assert np.array_equal(np.array(counts, dtype=np.int64), np.array(c["counts"], dtype=np.int64))
    assert (r, score) == (c["r"], c["score"])
    m = mout[0]
    assert len(m.Ps) == c["n_expanded"] and len(m.Nsa) == c["n_edges"]


def test_survey_known_answers_sha256_prefixes():
    want = {0: "7a32df7ffe47bd89", 1: "d22dd2c676f76397", 2: "95fbb9aa4762dca1", 3: "1329dbae877b646b",
            4: "52db4a5ec8dd97ba", 5: "baac753732718832", 6: "4f168b3f22365958"}
    for ci, pre in want.items():
        assert MCTS_CASES[ci]["sha256"].startswith(pre)


def test_pairwise_sum_replica_matches_numpy():
    rng = np.random.RandomState(0)
    for n in list(range(0, 700)):
        a = rng.rand(n) * (10.0 ** rng.randint(-3, 4, size=n))
        assert pairwise_sum(a) == np.sum(a), n
    # masked priors as the search produces them (zeros for invalid actions)
    for A in (150, 200, 72, 512):
        for _ in range(20):
            p = rng.rand(A).astype(np.float32).astype(np.float64) * (rng.rand(A) < 0.2)
            assert pairwise_sum(p) == np.sum(p)


def test_sum_plan_shapes():
    assert sum_plan(150) == ([(0, 72), (72, 78)], [0, 1, -1])
    assert sum_plan(200) == ([(0, 96), (96, 104)], [0, 1, -1])
    leaves, prog = sum_plan(512)
    assert len(leaves) <= 16 and len(prog) <= 32 and sum(n for _, n in leaves) == 512


def test_ranked_threshold_helper():
    assert np.isnan(bl_of([]))
    assert bl_of([0.5, 0.6, 0.7, 0.8001, 0.9]) == 0.7
    assert bl_of([0.3]) == 0.3          # index -1 wraps
    assert bl_of([0.2, 0.95]) == 0.2
    assert bl_of([1.0] * 100) == 1.0


def test_legacy_mt19937_restatement_matches_numpy_and_reference_fixtures():
    from oracle.mt_items import LegacyMT19937, items_generator
    for seed in (0, 1, 100, 4242, 99999, 2 ** 31 + 5, 2 ** 32 - 1):
        np.random.seed(seed)
        r = LegacyMT19937(seed)
        for hi in (2, 3, 7, 10, 15, 150, 1000, 2 ** 20 + 3):
            assert int(np.random.randint(hi)) == r.randint(hi)
        assert int(np.random.randint(5, 6)) == r.randint(5, 6)   # single value: no draw on either side
        assert int(np.random.randint(9)) == r.randint(9)
    for rec in load_items_golden():
        assert items_generator(rec["W"], rec["Hgen"], rec["N"], rec["seed"]) == rec["items"]
