#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference (`/root/reference/xw_mcts`) is imported as is; the only shim is the one SURVEY.md §8(c)
describes: `ndarray.tostring()` was removed in numpy 2.3, so `stringRepresentation`
(`binpacking/BinPackingGame.py:214-218`) is overridden in a subclass with the byte-identical
`b''.join(plane.tobytes())`.

Outputs (all small, committed):
  items.json        ItemsGenerator.items_generator outputs            (BinPackingGame.py:257-285)
  env.npz           random playouts: valid masks / next states / terminal rewards
                    (BinPackingGame.py:58-116,188-212 ; BinPackingLogic.py:47-109)
  mcts.json         per-move visit-count matrices of MCTS.getActionProb under deterministic stub
                    evaluators (MCTS_bpp.py:28-139)
  net.npz           BinPackingNNet forward (binpacking/pytorch/BinpackingNNet.py:50-81) outputs for a shipped
                    checkpoint and for a seeded random init, with the weights used

Stub evaluators (shared definition with oracle/ and the CUDA in-kernel evaluators; pop = popcount(bin),
nrem = number of remaining items, A = action size):
  U: p[a] = 1/A                         v = 0.0
  V: p[a] = 1/A                         v = ((7*pop + 3*nrem) % 16)/16 - 0.5
  H: p[a] = 1/(a + 3 + pop % 5)         v = as V        (non-dyadic: exercises np.sum's pairwise order)
  D: p[a] = ((37*a + 11 + pop) % 64 + 1)/4096           v = as V
"""
import hashlib
import json
import os
import sys

import numpy as np

REF = "/root/reference/xw_mcts"
HERE = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("WANDB_MODE", "disabled")
sys.path.insert(0, REF)

from binpacking.BinPackingGame import BinPackingGame as RefGame, ItemsGenerator as RefGen  # noqa: E402
from MCTS_bpp import MCTS as RefMCTS  # noqa: E402


def _no_choice(*a, **k):
    # BinPackingGame.py:212 draws a random +-1 when r == bl, and MCTS_bpp.py:46 a random arg-max: neither may
    # happen while fixtures are generated (they would make the fixture non-deterministic).
    raise RuntimeError("reference reached a random branch (ranked-reward tie); pick other fixture parameters")


np.random.choice = _no_choice


class Game(RefGame):
    def stringRepresentation(self, board):  # numpy>=2.3 shim, byte-identical key
        return b"".join(p.tobytes() for p in board)


class dotdict(dict):
    def __getattr__(self, name):
        return self[name]


def stub_value(board):
    pop = int(board[0].sum())
    nrem = int(sum(1 for p in board[1:] if p.sum() > 0))
    return ((7 * pop + 3 * nrem) % 16) / 16 - 0.5


class Stub:
    def __init__(self, kind, A):
        self.kind, self.A = kind, A

    def predict(self, board):
        A = self.A
        pop = int(board[0].sum())
        if self.kind == "U":
            return np.full(A, 1 / A, dtype=np.float64), 0.0
        if self.kind == "V":
            return np.full(A, 1 / A, dtype=np.float64), stub_value(board)
        if self.kind == "H":
            return 1.0 / (np.arange(A, dtype=np.float64) + 3 + pop % 5), stub_value(board)
        if self.kind == "D":
            a = np.arange(A, dtype=np.int64)
            return ((37 * a + 11 + pop) % 64 + 1).astype(np.float64) / 4096.0, stub_value(board)
        raise ValueError(self.kind)


def pack_state(state):
    """(N+1,H,W) int64 -> (occ rows as ints with bit x = column x, rem mask)."""
    occ = [int(sum(int(v) << x for x, v in enumerate(row))) for row in state[0]]
    rem = 0
    for i, p in enumerate(state[1:]):
        if p.sum() > 0:
            rem |= 1 << i
    return occ, rem


# --------------------------------------------------------------------------- items
def gen_items():
    out = []
    for (W, Hg, N) in [(15, 15, 10), (15, 8, 10), (15, 2, 10), (15, 5, 10), (20, 20, 10), (20, 9, 10), (12, 7, 6),
                       (15, 11, 10), (15, 3, 10)]:
        gen = RefGen(W, Hg, N)
        for seed in [0, 1, 5, 7, 100, 1000, 1001, 4242, 99999]:
            items = gen.items_generator(seed)
            out.append({"W": W, "Hgen": Hg, "N": N, "seed": seed, "items": [[int(v) for v in it] for it in items]})
    json.dump(out, open(os.path.join(HERE, "items.json"), "w"))
    print("items.json", len(out))


# --------------------------------------------------------------------------- env
def gen_env():
    rng = np.random.RandomState(12345)
    rec = {k: [] for k in ["W", "H", "N", "items_wh", "total_area", "max_h", "occ", "rem", "valid", "action",
                           "next_occ", "next_rem", "ended", "score", "ranked", "bl_case", "has_moves"]}
    reward_lists = [[], [0.5, 0.6, 0.7001, 0.8001, 0.9], [1.0] * 100, [0.3], [0.2, 0.95], [0.61, 0.62, 0.63, 0.64]]
    n_trunc = 0
    for ep in range(260):
        W, H = [(15, 15), (20, 20), (15, 15), (12, 9)][ep % 4]
        N = [10, 10, 10, 6][ep % 4]
        Hg = int(rng.randint(2, H + 1))
        seed = int(rng.randint(100000))
        g = Game(W, H, N, 1)
        items = RefGen(W, Hg, N).items_generator(seed)
        area = W * Hg
        board = g.getInitBoard()
        planes = g.getInitItems(items)
        wh = [[int(it[0]), int(it[1])] for it in items]
        wild = (ep % 5 == 4)  # every 5th episode also takes arbitrary (possibly invalid) placements
        for step in range(N + 2):
            state = g.getBinItem(board, planes)
            occ, rem = pack_state(state)
            if rem == 0:
                has = False
                valid = np.zeros(g.getActionSize(), dtype=np.int64)
            else:
                has = g.has_valid_moves(state)
                valid = g.getValidMoves(state) if has else np.zeros(g.getActionSize(), dtype=np.int64)
            blc = int(rng.randint(len(reward_lists)))
            rl = reward_lists[blc]
            # avoid the random tie branch (BinPackingGame.py:212): fixtures never have r == bl unless r == 1
            ended, score = g.getGameEnded(state, area, rl, 0.75)
            if ended == 0:
                score = -1.0
            if wild and rem:
                rem_items = [i for i in range(N) if rem >> i & 1]
                it = int(rem_items[rng.randint(len(rem_items))])
                a = it * W + int(rng.randint(W))
            elif has:
                va = np.flatnonzero(valid)
                a = int(va[rng.randint(len(va))])
            else:
                a = -1
            if a >= 0:
                nb, np_ = g.getNextState(board, a, planes)
                nocc, nrem = pack_state(g.getBinItem(nb, np_))
                it = a // W
                if int(nb.sum() - board.sum()) != wh[it][0] * wh[it][1]:
                    n_trunc += 1
            else:
                nocc, nrem = occ, rem
            for k, v in [("W", W), ("H", H), ("N", N), ("items_wh", wh + [[0, 0]] * (10 - N)), ("total_area", area),
                         ("max_h", int(g.max_h)), ("occ", occ + [0] * (20 - H)), ("rem", rem),
                         ("valid", list(valid) + [0] * (200 - len(valid))), ("action", a),
                         ("next_occ", nocc + [0] * (20 - H)), ("next_rem", nrem), ("ended", int(ended)),
                         ("score", float(score)), ("bl_case", blc), ("has_moves", int(has))]:
                rec[k].append(v)
            rec["ranked"].append(0)
            if a < 0:
                break
            board, planes = nb, np_
    arrs = {k: np.array(v) for k, v in rec.items()}
    arrs["valid"] = arrs["valid"].astype(np.uint8)
    arrs["reward_lists"] = np.array(json.dumps(reward_lists))
    np.savez_compressed(os.path.join(HERE, "env.npz"), **arrs)
    print("env.npz states:", len(rec["W"]), "truncated placements:", n_trunc,
          "terminal:", int((arrs["ended"] != 0).sum()))


# --------------------------------------------------------------------------- mcts
def run_mcts_case(W, H, N, genW, genH, seed, stub, sims, cpuct, alpha, rewards, policy="argmax"):
    g = Game(W, H, N, 1)
    items = RefGen(genW, genH, N).items_generator(seed)
    area = genW * genH
    args = dotdict(numMCTSSims=sims, cpuct=cpuct, alpha=alpha)
    mcts = RefMCTS(g, Stub(stub, g.getActionSize()), args)
    board = g.getInitBoard()
    planes = g.getInitItems(items)
    counts_all, actions, ns_root = [], [], []
    r, score = 0, None
    while True:
        state = g.getBinItem(board, planes)
        pi = mcts.getActionProb(state, area, rewards)
        s = g.stringRepresentation(state)
        counts = [mcts.Nsa.get((s, a), 0) for a in range(g.getActionSize())]
        assert np.allclose(np.array(counts) / sum(counts), pi)
        counts_all.append(counts)
        ns_root.append(int(mcts.Ns[s]))
        if policy == "argmax":
            a = int(np.argmax(counts))
        else:  # "second": a deterministic non-greedy walk - the last action with a non-zero count
            a = int(np.flatnonzero(counts)[-1])
        actions.append(a)
        board, planes = g.getNextState(board, a, planes)
        r, score = g.getGameEnded(g.getBinItem(board, planes), area, rewards, alpha)
        if r != 0:
            break
    cm = np.array(counts_all, dtype=np.int64)
    n_term = sum(1 for v in mcts.Es.values() if v != 0)
    return {
        "W": W, "H": H, "N": N, "genW": genW, "genH": genH, "seed": seed, "stub": stub, "sims": sims,
        "cpuct": cpuct, "alpha": alpha, "rewards": rewards, "policy": policy,
        "items": [[int(v) for v in it] for it in items],
        "counts": cm.tolist(), "actions": actions, "ns_root": ns_root, "r": int(r), "score": float(score),
        "n_expanded": len(mcts.Ps), "n_terminal": n_term, "n_edges": len(mcts.Nsa),
        "sha256": hashlib.sha256(cm.tobytes()).hexdigest(),
    }


def gen_mcts():
    cases = []
    RL = [0.5, 0.6, 0.7, 0.8001, 0.9]
    spec = [
        # the SURVEY.md §8(c) known answers
        (15, 15, 10, 15, 15, 100, "U", 200, 1, 0.75, []),
        (15, 15, 10, 15, 8, 7, "U", 200, 1, 0.75, []),
        (20, 20, 10, 20, 20, 5, "U", 200, 1, 0.75, []),
        (15, 15, 10, 15, 15, 100, "V", 200, 1, 0.75, []),
        (15, 15, 10, 15, 8, 7, "V", 200, 1, 0.75, []),
        (20, 20, 10, 20, 20, 5, "V", 200, 1, 0.75, []),
        (15, 15, 10, 15, 8, 7, "V", 200, 1, 0.75, RL),
        # wider coverage: other stubs, cpuct, sims, heights
        (15, 15, 10, 15, 8, 7, "H", 200, 1, 0.75, RL),
        (15, 15, 10, 15, 8, 7, "D", 200, 1, 0.75, RL),
        (15, 15, 10, 15, 5, 1001, "H", 200, 1.5, 0.75, [1.0] * 100),
        (15, 15, 10, 15, 12, 1002, "D", 200, 0.5, 0.75, [0.4, 0.7]),
        (20, 20, 10, 20, 9, 1003, "H", 200, 1, 0.75, RL),
        (20, 20, 10, 20, 14, 1004, "D", 100, 2.0, 0.75, []),
        (15, 15, 10, 15, 2, 1005, "V", 200, 1, 0.75, RL),
        (15, 15, 10, 15, 3, 1006, "H", 50, 1, 0.75, RL),
        (12, 9, 6, 12, 7, 1007, "D", 64, 1, 0.75, RL),
        (15, 15, 10, 15, 10, 1008, "V", 25, 1, 0.75, []),
    ]
    for i, (W, H, N, gw, gh, seed, stub, sims, cpuct, alpha, rl) in enumerate(spec):
        c = run_mcts_case(W, H, N, gw, gh, seed, stub, sims, cpuct, alpha, rl)
        cases.append(c)
        print("mcts", i, stub, (W, H, N, gw, gh, seed), "actions", c["actions"], "score", c["score"], "r", c["r"],
              c["n_expanded"] + c["n_terminal"], c["n_edges"], c["sha256"][:16])
    # a non-greedy walk through the tree
    for (W, H, N, gw, gh, seed, stub, sims) in [(15, 15, 10, 15, 9, 2001, "H", 100), (15, 15, 10, 15, 13, 2002, "D", 100)]:
        c = run_mcts_case(W, H, N, gw, gh, seed, stub, sims, 1, 0.75, RL, policy="second")
        cases.append(c)
        print("mcts second", stub, seed, c["actions"], c["score"], c["sha256"][:16])
    # the bench-distribution sample: seeds 1000.., heights by the BASELINE.md protocol, stub U and V, 200 sims
    hr = np.random.RandomState(2024)
    for k in range(8):
        gh = int(hr.randint(2, 16))
        for stub in ("U", "V"):
            c = run_mcts_case(15, 15, 10, 15, gh, 1000 + k, stub, 200, 1, 0.75, [])
            cases.append(c)
            print("mcts bench", stub, 1000 + k, gh, c["actions"], c["score"], c["sha256"][:16])
    json.dump(cases, open(os.path.join(HERE, "mcts.json"), "w"))
    print("mcts.json", len(cases))


# --------------------------------------------------------------------------- net
def gen_net():
    import torch
    from binpacking.pytorch.NNet import NNetWrapper as RefNet

    out = {}
    rng = np.random.RandomState(777)

    def sample_states(W, H, N, n):
        g = Game(W, H, N, 1)
        states = []
        while len(states) < n:
            Hg = int(rng.randint(2, H + 1))
            items = RefGen(W, Hg, N).items_generator(int(rng.randint(100000)))
            board, planes = g.getInitBoard(), g.getInitItems(items)
            for _ in range(N):
                st = g.getBinItem(board, planes)
                if not g.has_valid_moves(st):
                    break
                states.append(st)
                va = np.flatnonzero(g.getValidMoves(st))
                board, planes = g.getNextState(board, int(va[rng.randint(len(va))]), planes)
        return g, np.array(states[:n])

    # (1) shipped checkpoint, default architecture (15x15, N=10)
    g, states = sample_states(15, 15, 10, 48)
    args = dotdict(num_items=10, num_bins=1, cuda=False)
    net = RefNet(g, args)
    ck = os.path.join(REF, "wandb/run-20201113_144231-15y0rcng/temp")
    net.load_checkpoint(ck, "temp.pth.tar")
    pis, vs = zip(*[net.predict(s) for s in states])
    out["ck_states"] = states.astype(np.uint8)
    out["ck_pi"] = np.array(pis, dtype=np.float32)
    out["ck_v"] = np.array(vs, dtype=np.float32).reshape(-1)
    for k, v in net.nnet.state_dict().items():
        out["ck_w." + k] = v.numpy().astype(np.float32)

    # (2) seeded random init at 20x20 (no shipped checkpoint fits)
    g, states = sample_states(20, 20, 10, 24)
    torch.manual_seed(0)
    net = RefNet(g, args)
    pis, vs = zip(*[net.predict(s) for s in states])
    out["r20_states"] = states.astype(np.uint8)
    out["r20_pi"] = np.array(pis, dtype=np.float32)
    out["r20_v"] = np.array(vs, dtype=np.float32).reshape(-1)
    for k, v in net.nnet.state_dict().items():
        out["r20_w." + k] = v.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "net.npz"), **out)
    print("net.npz", {k: v.shape for k, v in out.items() if not k.startswith(("ck_w", "r20_w"))})


if __name__ == "__main__":
    which = sys.argv[1:] or ["items", "env", "mcts", "net"]
    if "items" in which:
        gen_items()
    if "env" in which:
        gen_env()
    if "mcts" in which:
        gen_mcts()
    if "net" in which:
        gen_net()
