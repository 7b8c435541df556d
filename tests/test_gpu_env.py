"""-m gpu: the CUDA env ops (through the C ABI) against fixtures produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from helpers import bl_of, load_env_golden, recs_from_occ

pytestmark = pytest.mark.gpu


def _groups(g):
    keys = sorted({(int(w), int(h), int(n)) for w, h, n in zip(g["W"], g["H"], g["N"])})
    for (W, H, N) in keys:
        idx = np.flatnonzero((g["W"] == W) & (g["H"] == H) & (g["N"] == N))
        yield W, H, N, idx


def test_valid_moves_bit_exact():
    from resource_packing_self_play_b200.engine import EnvOps
    g = load_env_golden()
    total = 0
    for W, H, N, idx in _groups(g):
        ops = EnvOps(W, H, N)
        recs = recs_from_occ(g["occ"][idx], g["rem"][idx], H)
        items = g["items_wh"][idx][:, :N, :].astype(np.int32)
        got = ops.valid_moves(recs, items).cpu().numpy()
        want = g["valid"][idx][:, :W * N]
        assert got.shape == want.shape
        assert np.array_equal(got, want), f"valid mask mismatch for {(W, H, N)}"
        assert np.array_equal(got.any(axis=1).astype(int), g["has_moves"][idx])
        total += len(idx)
    assert total == len(g["W"])


def test_next_state_bit_exact_including_truncated_placements():
    from resource_packing_self_play_b200.engine import EnvOps
    g = load_env_golden()
    for W, H, N, idx in _groups(g):
        idx = idx[g["action"][idx] >= 0]
        ops = EnvOps(W, H, N)
        recs = recs_from_occ(g["occ"][idx], g["rem"][idx], H)
        items = g["items_wh"][idx][:, :N, :].astype(np.int32)
        got = ops.next_state(recs, items, g["action"][idx].astype(np.int32)).cpu().numpy().view(np.uint32)
        want = recs_from_occ(g["next_occ"][idx], g["next_rem"][idx], H)
        assert np.array_equal(got, want), f"next state mismatch for {(W, H, N)}"


def test_game_ended_and_ranked_reward_bit_exact():
    from resource_packing_self_play_b200.engine import EnvOps
    g = load_env_golden()
    bls = np.array([bl_of(r) for r in g["reward_lists"]])
    n_term = 0
    for W, H, N, idx in _groups(g):
        ops = EnvOps(W, H, N)
        recs = recs_from_occ(g["occ"][idx], g["rem"][idx], H)
        items = g["items_wh"][idx][:, :N, :].astype(np.int32)
        ended, score = ops.game_ended(recs, items, g["total_area"][idx].astype(np.int32),
                                      g["max_h"][idx].astype(np.int32), bls[g["bl_case"][idx]])
        ended, score = ended.cpu().numpy(), score.cpu().numpy()
        assert np.array_equal(ended, g["ended"][idx])
        term = ended != 0
        assert np.array_equal(score[term], g["score"][idx][term])  # float64, bit-exact
        n_term += int(term.sum())
    assert n_term >= 200


def test_empty_batch_and_bad_geometry():
    from resource_packing_self_play_b200 import _lib
    from resource_packing_self_play_b200.engine import EnvOps
    ops = EnvOps(15, 15, 10)
    out = ops.valid_moves(np.zeros((0, 32), dtype=np.uint32), np.zeros((0, 10, 2), dtype=np.int32))
    assert out.shape == (0, 150)
    bad = EnvOps(40, 15, 10)
    with pytest.raises(_lib.BppError):
        bad.valid_moves(np.zeros((1, 32), dtype=np.uint32), np.zeros((1, 10, 2), dtype=np.int32))


def test_device_items_generator_matches_reference_fixtures_and_numpy():
    from helpers import load_items_golden
    from oracle.mt_items import items_generator
    from resource_packing_self_play_b200.game import ItemsGenerator
    gold = load_items_golden()
    for (W, N) in sorted({(r["W"], r["N"]) for r in gold}):
        recs = [r for r in gold if (r["W"], r["N"]) == (W, N)]
        gen = ItemsGenerator(W, 1, N)
        wh, rects = gen.items_batch_device([r["seed"] for r in recs], [r["Hgen"] for r in recs], rects=True)
        assert rects.cpu().numpy().tolist() == [r["items"] for r in recs]
        assert wh.cpu().numpy().tolist() == [[it[:2] for it in r["items"]] for r in recs]
    # 4,096 bench seeds against numpy's own generator (host path) and the restatement
    W, N, n = 15, 10, 4096
    seeds = 1000 + np.arange(n)
    heights = np.array([np.random.RandomState(77000 + int(b)).randint(2, 16) for b in np.arange(n) // 20])
    gen = ItemsGenerator(W, 15, N)
    dev = gen.items_batch_device(seeds, heights).cpu().numpy()
    assert np.array_equal(dev, gen.items_batch(seeds, heights))
    big = ItemsGenerator(32, 28, 16).items_batch_device([2 ** 32 - 1, 0, 7], [28, 28, 1]).cpu().numpy()
    for s, h, got in zip([2 ** 32 - 1, 0, 7], [28, 28, 1], big):
        assert got.tolist() == [it[:2] for it in items_generator(32, h, 16, s)]
