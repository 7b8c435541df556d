"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_env_golden():
    d = np.load(os.path.join(GOLDEN, "env.npz"))
    out = {k: d[k] for k in d.files}
    out["reward_lists"] = json.loads(str(out["reward_lists"]))
    return out


def load_mcts_golden():
    return json.load(open(os.path.join(GOLDEN, "mcts.json")))


def load_items_golden():
    return json.load(open(os.path.join(GOLDEN, "items.json")))


def bl_of(rewards, alpha=0.75):
    """ranked-reward threshold of BinPackingGame.py:203-206; NaN = empty list"""
    if len(rewards) == 0:
        return float("nan")
    s = np.sort(np.asarray(rewards, dtype=np.float64))
    return float(s[int(np.floor(len(s) * alpha)) - 1])


def recs_from_occ(occ, rem, H):
    """golden occ rows (n, 20) + rem (n,) -> uint32 records (n, 32)"""
    n = occ.shape[0]
    recs = np.zeros((n, 32), dtype=np.uint32)
    recs[:, :H] = occ[:, :H].astype(np.uint32)
    recs[:, 28] = rem.astype(np.uint32)
    return recs
