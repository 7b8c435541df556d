"""CPU oracle (Python/numpy) for the self-play hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, on the CPU, the algorithm of the reference's hot path so that the CUDA path can be
checked against it where `/root/reference` does not exist (the GPU box).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import it; nothing under
`resource_packing_self_play_b200/` does.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function below against fixtures produced by the
unmodified reference (`tests/golden/make_golden.py` -> items.json / env.npz / mcts.json), including the
SURVEY.md §8(c) known answers.

It deliberately keeps the reference's data representation and iteration structure — an `(N+1, H, W)` int64
tensor per state, a bytes key of 8·(N+1)·H·W bytes, six dicts, Python loops with the builtin `sum` over numpy
rows — so that timing it (`bench.py`'s `cpu_baseline`, kind "port") measures the same kind of work the
reference does on a CPU.  Each function cites the reference lines it follows
(paths relative to /root/reference/xw_mcts).
"""
import math

import numpy as np

EPS = 1e-8  # MCTS_bpp.py:6


# ----------------------------------------------------------------------------------------------------------------------
# Bin primitives                                                           binpacking/BinPackingLogic.py
# ----------------------------------------------------------------------------------------------------------------------
def left_adjacent(grid, H, x, w):
    """BinPackingLogic.py:47-78 (only the 'left' test is live).  True iff the item would touch something on its
    left: x == 0, or the cell left of the strip is occupied in the first strip row that is completely empty
    (the last row if no strip row is empty)."""
    if x == 0:
        return True
    t = 0
    for t in range(H):
        if sum(grid[t, x:x + w]) == 0:
            break
    return grid[t, x - 1] > 0


def item_dims(plane):
    """BinPackingLogic.py:84-85 / BinPackingGame.py:71-72: width and height are read back from the plane."""
    return sum(plane[0, :]), sum(plane[:, 0])


def columns_for_item(grid, W, H, planes, idx):
    """BinPackingLogic.py:80-93.  Legal columns of one remaining item: (A) the strip holds at most w*(H-h)
    occupied cells (a cell count, not a count of empty rows) and (B) left adjacency."""
    plane = planes[idx]
    assert sum(sum(plane)) > 0
    w, h = item_dims(plane)
    out = []
    for x in range(W - w + 1):
        if sum(sum(grid[0:, x:x + w])) <= (w * H - w * h):
            if left_adjacent(grid, H, x, w):
                out.append((idx, x))
    return out


def place(grid, H, x, w, h):
    """BinPackingLogic.py:95-109.  Fill the strip in the first h strip rows that are completely empty, scanning from
    row 0; rows need not be contiguous and fewer than h may exist (silent truncation)."""
    g = grid.copy()
    filled = 0
    for r in range(H):
        if sum(g[r, x:x + w]) == 0:
            g[r, x:x + w] = 1
            filled += 1
            if filled == h:
                break
    return g.copy()


# ----------------------------------------------------------------------------------------------------------------------
# Game                                                                     binpacking/BinPackingGame.py
# ----------------------------------------------------------------------------------------------------------------------
class OracleGame:
    """Restates BinPackingGame (BinPackingGame.py:8-218) — same constructor and method names."""

    def __init__(self, bin_width, bin_height, num_items, n):  # :15-22
        self.bin_width = bin_width
        self.bin_height = bin_height
        self.num_items = num_items
        self.n = n
        self.cur_item = 0
        self.sum_h = 0
        self.max_h = 0

    def getInitBoard(self):  # :24-27
        return np.array([[0] * self.bin_width for _ in range(self.bin_height)])

    def getBoardSize(self):  # :29-31
        return (self.bin_height, self.bin_width)

    def getActionSize(self):  # :33-35
        return self.bin_width * self.num_items

    def getInitItems(self, items_list):  # :37-51 (hidden state: sum_h, max_h)
        planes = []
        tot, top = 0, 0
        for i in range(self.num_items):
            w, h = items_list[i][0], items_list[i][1]
            p = self.getInitBoard()
            p[0:h, 0:w] = 1
            planes += [p]
            tot += h
            top = max(top, h)
        self.sum_h = tot
        self.max_h = top
        return planes

    def getItemsUpdated(self, planes, cur_item):  # :53-56
        planes[cur_item] -= planes[cur_item]
        return planes

    def getNextState(self, board, action, planes):  # :58-76
        planes = np.copy(planes)
        grid = np.copy(board)
        idx, x = int(action / self.bin_width), int(action % self.bin_width)
        plane = planes[idx]
        assert sum(sum(plane)) > 0
        w, h = item_dims(plane)
        grid = place(grid, self.bin_height, x, w, h)
        planes = self.getItemsUpdated(planes, idx)
        return (grid, planes)

    def getValidMoves(self, state):  # :78-92
        valids = [0] * self.getActionSize()
        grid = np.copy(state[0])
        legal = []
        for i in range(self.num_items):
            if sum(sum(state[i + 1])) == 0:
                continue
            legal += columns_for_item(grid, self.bin_width, self.bin_height, state[1:], i)
        assert len(legal) > 0
        for i, x in legal:
            valids[i * self.bin_width + x] = 1
        return np.array(valids)

    def has_valid_moves(self, state):  # :94-107 (early exit at the first item with a legal column)
        grid = np.copy(state[0])
        moves = []
        for i in range(self.num_items):
            if sum(sum(state[i + 1])) == 0:
                continue
            moves = columns_for_item(grid, self.bin_width, self.bin_height, state[1:], i)
            if len(moves) > 0:
                break
        return len(moves) > 0

    def getGameEnded(self, state, items_total_area, rewards_list, alpha):  # :109-116
        assert len(state) == self.num_items + self.n
        if not self.has_valid_moves(state):
            return self.getRankedReward(state, items_total_area, rewards_list, alpha)
        return 0, []

    def getBinItem(self, board, planes):  # :118-120
        return np.array([board] + list(planes))

    def get_minimal_bin_height(self, board):  # :181-186
        i = 0
        for i in reversed(range(self.bin_height)):
            if sum(board[i, :]) > 0:
                break
        return i + 1

    def getRankedReward(self, state, items_total_area, rewards_list, alpha, tie=None):  # :188-212
        rewards_list = list(rewards_list)
        if sum(sum(state[0, :])) != items_total_area:
            r = 0
        else:
            r = max(np.ceil(items_total_area / self.bin_width), self.max_h) / self.get_minimal_bin_height(state[0, :])
        if len(rewards_list) == 0:
            return 1, r
        srt = np.sort(rewards_list)
        bl = srt[int(np.floor(len(srt) * alpha)) - 1]
        if r > bl or r == 1:
            return 1, r
        if r < bl:
            return -1, r
        # the reference draws +-1 uniformly here (np.random.choice, :212); the oracle takes an injected bit
        if tie is None:
            raise RuntimeError("ranked-reward tie (r == bl): inject `tie`")
        return tie, r

    def stringRepresentation(self, state):  # :214-218 (tostring() == tobytes())
        return b"".join(p.tobytes() for p in state)


class OracleItemsGenerator:
    """BinPackingGame.py:250-285: guillotine splits driven by numpy's legacy global RNG."""

    def __init__(self, bin_width, bin_height, items):
        self.bin_width = bin_width
        self.bin_height = bin_height
        self.n = items

    def items_generator(self, seed):
        np.random.seed(seed)
        rects = [[self.bin_width, self.bin_height, 0, 0]]
        while len(rects) < self.n:
            axis = np.random.randint(2)
            k = np.random.randint(len(rects))
            w, h, a, b = rects[k]
            if axis == 0:
                if w == 1:
                    continue
                cut = np.random.randint(a + 1, a + w)
                rects.append([cut - a, h, a, b])
                rects.append([w - (cut - a), h, cut, b])
                rects.pop(k)
            else:
                if h == 1:
                    continue
                cut = np.random.randint(b + 1, b + h)
                rects.append([w, cut - b, a, b])
                rects.append([w, h - (cut - b), a, cut])
                rects.pop(k)
        return rects


# ----------------------------------------------------------------------------------------------------------------------
# MCTS                                                                     MCTS_bpp.py
# ----------------------------------------------------------------------------------------------------------------------
class OracleMCTS:
    """Restates MCTS (MCTS_bpp.py:11-139): six dicts keyed by the state bytes; single player, no sign flip."""

    def __init__(self, game, nnet, args):  # :16-26
        self.game, self.nnet, self.args = game, nnet, args
        self.Qsa, self.Nsa, self.Ns, self.Ps, self.Es, self.Vs = {}, {}, {}, {}, {}, {}
        # instrumentation for bench.py's algorithmic-bytes model (not in the reference)
        self.n_edges_walked = 0
        self.n_expansions = 0
        self.n_sims = 0

    def root_counts(self, state):  # :40-41
        s = self.game.stringRepresentation(state)
        return [self.Nsa[(s, a)] if (s, a) in self.Nsa else 0 for a in range(self.game.getActionSize())]

    def getActionProb(self, state, totalArea, rewardsList, greedy_a=1, rng=None):  # :28-54
        for _ in range(self.args.numMCTSSims):
            self.n_sims += 1
            self.search(state, totalArea, rewardsList)
        counts = self.root_counts(state)
        if greedy_a == 0:
            best = np.array(np.argwhere(counts == np.max(counts))).flatten()
            pick = (rng or np.random).choice(best)
            probs = [0] * len(counts)
            probs[pick] = 1
            return probs
        counts = [x ** (1. / greedy_a) for x in counts]
        tot = float(sum(counts))
        return [x / tot for x in counts]

    def search(self, state, totalArea, rewardsList):  # :56-139
        g = self.game
        s = g.stringRepresentation(state)
        if s not in self.Es:
            self.Es[s], _ = g.getGameEnded(state, totalArea, rewardsList, self.args.alpha)
        if self.Es[s] != 0:
            return self.Es[s]
        if s not in self.Ps:
            self.n_expansions += 1
            self.Ps[s], v = self.nnet.predict(state)
            valids = g.getValidMoves(state)
            self.Ps[s] = self.Ps[s] * valids
            tot = np.sum(self.Ps[s])
            if tot > 0:
                self.Ps[s] /= tot
            else:
                self.Ps[s] = self.Ps[s] + valids
                self.Ps[s] /= np.sum(self.Ps[s])
            self.Vs[s] = valids
            self.Ns[s] = 0
            return v
        valids = self.Vs[s]
        best_u, best_a = -float('inf'), -1
        for a in range(g.getActionSize()):
            if valids[a]:
                if (s, a) in self.Qsa:
                    u = self.Qsa[(s, a)] + self.args.cpuct * self.Ps[s][a] * math.sqrt(self.Ns[s]) / (
                        1 + self.Nsa[(s, a)])
                else:
                    u = self.args.cpuct * self.Ps[s][a] * math.sqrt(self.Ns[s] + EPS)
                if u > best_u:
                    best_u, best_a = u, a
        a = best_a
        self.n_edges_walked += 1
        board, planes = g.getNextState(state[0], a, state[1:])
        v = self.search(g.getBinItem(board, planes), totalArea, rewardsList)
        if (s, a) in self.Qsa:
            self.Qsa[(s, a)] = (self.Nsa[(s, a)] * self.Qsa[(s, a)] + v) / (self.Nsa[(s, a)] + 1)
            self.Nsa[(s, a)] += 1
        else:
            self.Qsa[(s, a)] = v
            self.Nsa[(s, a)] = 1
        self.Ns[s] += 1
        return v


# ----------------------------------------------------------------------------------------------------------------------
# Deterministic stub evaluators (shared definition: tests/golden/make_golden.py, csrc stub evaluators)
# ----------------------------------------------------------------------------------------------------------------------
def stub_value(state):
    pop = int(state[0].sum())
    nrem = int(sum(1 for p in state[1:] if p.sum() > 0))
    return ((7 * pop + 3 * nrem) % 16) / 16 - 0.5


class StubNet:
    """kind in 'U','V','H','D' — see tests/golden/make_golden.py for the definitions."""

    def __init__(self, kind, A):
        self.kind, self.A = kind, A

    def predict(self, state):
        A = self.A
        pop = int(state[0].sum())
        if self.kind == "U":
            return np.full(A, 1 / A, dtype=np.float64), 0.0
        if self.kind == "V":
            return np.full(A, 1 / A, dtype=np.float64), stub_value(state)
        if self.kind == "H":
            return 1.0 / (np.arange(A, dtype=np.float64) + 3 + pop % 5), stub_value(state)
        if self.kind == "D":
            a = np.arange(A, dtype=np.int64)
            return ((37 * a + 11 + pop) % 64 + 1).astype(np.float64) / 4096.0, stub_value(state)
        raise ValueError(self.kind)


class dotdict(dict):
    def __getattr__(self, name):
        return self[name]


def play_episode(W, H, N, items, total_area, stub, sims, cpuct, alpha, rewards, policy="argmax", mcts_out=None):
    """Drive one episode the way tests/golden/make_golden.py does (CoachBPP.py:50-99 with an injected,
    deterministic action choice).  Returns (counts[moves][A], actions, r, score)."""
    g = OracleGame(W, H, N, 1)
    net = stub if hasattr(stub, "predict") else StubNet(stub, g.getActionSize())
    m = OracleMCTS(g, net, dotdict(numMCTSSims=sims, cpuct=cpuct, alpha=alpha))
    board, planes = g.getInitBoard(), g.getInitItems(items)
    counts_all, actions = [], []
    while True:
        state = g.getBinItem(board, planes)
        m.getActionProb(state, total_area, rewards)
        counts = m.root_counts(state)
        counts_all.append(counts)
        a = int(np.argmax(counts)) if policy == "argmax" else int(np.flatnonzero(counts)[-1])
        actions.append(a)
        board, planes = g.getNextState(board, a, planes)
        r, score = g.getGameEnded(g.getBinItem(board, planes), total_area, rewards, alpha)
        if r != 0:
            break
    if mcts_out is not None:
        mcts_out.append(m)
    return counts_all, actions, int(r), float(score)


# ----------------------------------------------------------------------------------------------------------------------
# compact-state helpers (the layout the CUDA path uses; SURVEY.md §8 "compact equivalent of a state")
# ----------------------------------------------------------------------------------------------------------------------
def pack_state(state):
    occ = [int(sum(int(v) << x for x, v in enumerate(row))) for row in state[0]]
    rem = 0
    for i, p in enumerate(state[1:]):
        if p.sum() > 0:
            rem |= 1 << i
    return occ, rem


def unpack_state(occ, rem, items_wh, W, H):
    N = len(items_wh)
    st = np.zeros((N + 1, H, W), dtype=np.int64)
    for r in range(H):
        for x in range(W):
            st[0, r, x] = (occ[r] >> x) & 1
    for i, (w, h) in enumerate(items_wh):
        if rem >> i & 1:
            st[i + 1, 0:h, 0:w] = 1
    return st
