"""ctypes loader of the C oracle (oracle/bpp_oracle.c).  TEST INFRASTRUCTURE ONLY — see the header of that file."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboracle.so")
STUB = {"U": 1, "V": 2, "H": 3, "D": 4}
_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return SO


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(HERE, "bpp_oracle.c")):
            build()
        _lib = C.CDLL(SO)
        _lib.oracle_pairwise_sum.restype = C.c_double
        _lib.oracle_pairwise_sum.argtypes = [C.c_void_p, C.c_int]
        _lib.oracle_play_episode.restype = C.c_int
        _lib.oracle_play_episode.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int,
                                             C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else C.c_void_p(0)


def valid_moves(W, H, N, recs, items_wh):
    recs = np.ascontiguousarray(recs, dtype=np.uint32).reshape(-1, 32)
    items = np.ascontiguousarray(items_wh, dtype=np.int32)
    out = np.zeros((recs.shape[0], W * N), dtype=np.uint8)
    load().oracle_valid_moves(W, H, N, recs.shape[0], _p(recs), _p(items), _p(out))
    return out


def next_state(W, H, N, recs, items_wh, actions):
    recs = np.ascontiguousarray(recs, dtype=np.uint32).reshape(-1, 32)
    items = np.ascontiguousarray(items_wh, dtype=np.int32)
    act = np.ascontiguousarray(actions, dtype=np.int32)
    out = np.zeros_like(recs)
    load().oracle_next_state(W, H, N, recs.shape[0], _p(recs), _p(items), _p(act), _p(out))
    return out


def game_ended(W, H, N, recs, items_wh, total_area, max_h, bl, tie=None):
    recs = np.ascontiguousarray(recs, dtype=np.uint32).reshape(-1, 32)
    items = np.ascontiguousarray(items_wh, dtype=np.int32)
    n = recs.shape[0]
    area = np.ascontiguousarray(total_area, dtype=np.int32)
    mh = np.ascontiguousarray(max_h, dtype=np.int32)
    blv = np.ascontiguousarray(bl, dtype=np.float64)
    tiev = np.ascontiguousarray(tie, dtype=np.int8) if tie is not None else None
    ended = np.zeros(n, dtype=np.int32)
    score = np.zeros(n, dtype=np.float64)
    load().oracle_game_ended(W, H, N, n, _p(recs), _p(items), _p(area), _p(mh), _p(blv), _p(tiev), _p(ended), _p(score))
    return ended, score


def pairwise_sum(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return float(load().oracle_pairwise_sum(_p(a), len(a)))


def play_episode(W, H, N, items_wh, total_area, bl, stub, sims, cpuct, policy=0, forced=None, tie=1):
    """returns dict(counts (moves, A) int32, actions (moves,), r, score, stats)"""
    items = np.ascontiguousarray(items_wh, dtype=np.int32).reshape(N, 2)
    A = W * N
    counts = np.zeros((N, A), dtype=np.int32)
    actions = np.full(N, -1, dtype=np.int32)
    r = C.c_int32(0)
    score = C.c_double(0.0)
    stats = np.zeros(6, dtype=np.int64)
    f = np.ascontiguousarray(forced, dtype=np.int32) if forced is not None else None
    kind = STUB[stub] if isinstance(stub, str) else int(stub)
    moves = load().oracle_play_episode(W, H, N, _p(items), int(total_area), float(bl), int(tie), kind, int(sims),
                                       float(cpuct), int(policy), _p(f), _p(counts), _p(actions), C.byref(r),
                                       C.byref(score), _p(stats))
    return {"counts": counts[:moves], "actions": actions[:moves], "r": int(r.value), "score": float(score.value),
            "moves": moves,
            "stats": dict(zip(["sims", "edges", "expansions", "terminals", "nodes", "nsa_entries"], stats.tolist()))}
