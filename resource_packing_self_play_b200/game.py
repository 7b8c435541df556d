"""Host-side mirror of the reference's game interface for the bin-packing path.

`BinPackingGame` keeps the constructor, method names, argument meaning, return types and assertion behaviour of
xw_mcts/binpacking/BinPackingGame.py:8-218 so it is a drop-in for CoachBPP / MCTS, but every rule evaluation runs in
the CUDA env kernels (csrc/bpp_engine.cu via include/bpp_b200.h).  The single-state methods are batch-of-one calls of
the batched ops; the `*_batch` methods are what the fast paths use.  There is no CPU implementation of the rules in
this package.
"""
import numpy as np
import torch

from .engine import EnvOps, pack_states, ranked_threshold, unpack_states


_TIE_RNG = np.random.default_rng()  # never touches numpy's global legacy generator


class BinPackingGame:
    """Rules facade, same surface as BinPackingGame.py:8-218 (dead Othello leftovers and the disabled, buggy
    getSymmetries are not provided; see DESIGN.md "out of scope")."""

    def __init__(self, bin_width, bin_height, num_items, n, device=None):  # BinPackingGame.py:15-22
        self.bin_width = bin_width
        self.bin_height = bin_height
        self.num_items = num_items
        self.n = n
        self.cur_item = 0
        self.sum_h = 0
        self.max_h = 0
        self._ops = EnvOps(bin_width, bin_height, num_items, device)

    # ---- shapes / initial state ------------------------------------------------------------------------------------
    def getInitBoard(self):  # :24-27
        return np.zeros((self.bin_height, self.bin_width), dtype=np.int64)

    def getBoardSize(self):  # :29-31
        return (self.bin_height, self.bin_width)

    def getActionSize(self):  # :33-35
        return self.bin_width * self.num_items

    def getInitItems(self, items_list):  # :37-51 — also sets the hidden state sum_h / max_h
        planes = []
        sum_h = max_h = 0
        for i in range(self.num_items):
            w, h = int(items_list[i][0]), int(items_list[i][1])
            p = self.getInitBoard()
            p[0:h, 0:w] = 1
            planes.append(p)
            sum_h += h
            max_h = max(max_h, h)
        self.sum_h, self.max_h = sum_h, max_h
        return planes

    def getItemsUpdated(self, items_list_board, cur_item):  # :53-56
        items_list_board[cur_item] -= items_list_board[cur_item]
        return items_list_board

    def getBinItem(self, board, items_list_board):  # :118-120
        return np.array([board] + list(items_list_board))

    def stringRepresentation(self, board):  # :214-218 (tostring() == tobytes())
        return b"".join(np.ascontiguousarray(p).tobytes() for p in board)

    # ---- rules (CUDA) ----------------------------------------------------------------------------------------------
    def _pack(self, state):
        return pack_states(np.asarray(state), self.bin_width, self.bin_height, self.num_items)

    def getNextState(self, board, action, items_list_board):  # :58-76
        state = self.getBinItem(np.asarray(board), list(items_list_board))
        recs, items = self._pack(state)
        cur_item = int(action / self.bin_width)
        assert (recs[0, 28] >> cur_item) & 1, "must choose a valid item"  # :69
        out = self._ops.next_state(recs, items, np.array([int(action)], dtype=np.int32)).cpu().numpy().view(np.uint32)
        nxt = unpack_states(out, items, self.bin_width, self.bin_height, self.num_items, dtype=state.dtype)[0]
        return nxt[0], nxt[1:]

    def getValidMoves(self, board):  # :78-92
        recs, items = self._pack(board)
        valids = self._ops.valid_moves(recs, items).cpu().numpy()[0].astype(np.int64)
        assert valids.sum() > 0  # :89
        return valids

    def has_valid_moves(self, board):  # :94-107
        recs, items = self._pack(board)
        return bool(self._ops.valid_moves(recs, items).any().item())

    def getGameEnded(self, total_board, items_total_area, rewards_list, alpha, tie=None):  # :109-116
        assert len(total_board) == self.num_items + self.n
        recs, items = self._pack(total_board)
        if tie is None:
            tie = 1 if _TIE_RNG.random() < 0.5 else -1  # the reference draws +-1 on r == bl (:212); private generator
        ended, score = self._ops.game_ended(recs, items, [int(items_total_area)], [int(self.max_h)],
                                            [ranked_threshold(rewards_list, alpha)], [tie])
        e = int(ended.item())
        if e == 0:
            return 0, []
        return e, np.float64(score.item())

    def getRankedReward(self, total_board, items_total_area, rewards_list, alpha, tie=None):  # :188-212
        """Ranked reward of a state *assumed* terminal (the reference calls it only from getGameEnded)."""
        recs, items = self._pack(total_board)
        recs = recs.copy()
        recs[:, 28] = 0  # no remaining items -> no legal move -> the kernel evaluates the reward branch
        if tie is None:
            tie = 1 if _TIE_RNG.random() < 0.5 else -1
        ended, score = self._ops.game_ended(recs, items, [int(items_total_area)], [int(self.max_h)],
                                            [ranked_threshold(rewards_list, alpha)], [tie])
        return int(ended.item()), np.float64(score.item())

    def get_minimal_bin_height(self, board):  # :181-186
        b = np.asarray(board)
        rows = np.flatnonzero(b.sum(axis=1) > 0)
        return int(rows[-1]) + 1 if len(rows) else 1

    # ---- batched variants (device tensors in / out) ---------------------------------------------------------------------
    def valid_moves_batch(self, recs, items_wh):
        return self._ops.valid_moves(recs, items_wh)

    def next_state_batch(self, recs, items_wh, actions):
        return self._ops.next_state(recs, items_wh, actions)

    def game_ended_batch(self, recs, items_wh, total_area, max_h, bl, tie=None):
        return self._ops.game_ended(recs, items_wh, total_area, max_h, bl, tie)


class ItemsGenerator:
    """Guillotine-split instance generator, same surface and same random stream as BinPackingGame.py:250-285
    (numpy's legacy global MT19937: `np.random.seed(seed)` then `randint` draws, including the side effect of
    re-seeding the global generator).  `bin_height` is mutable (CoachBPP.py:118)."""

    def __init__(self, bin_width, bin_height, items):
        self.bin_width = bin_width
        self.bin_height = bin_height
        self.n = items

    def _generate(self, rs):
        rects = [[self.bin_width, self.bin_height, 0, 0]]
        while len(rects) < self.n:
            axis = rs.randint(2)
            k = rs.randint(len(rects))
            w, h, a, b = rects[k]
            if axis == 0:
                if w == 1:
                    continue
                cut = rs.randint(a + 1, a + w)
                rects += [[cut - a, h, a, b], [w - (cut - a), h, cut, b]]
            else:
                if h == 1:
                    continue
                cut = rs.randint(b + 1, b + h)
                rects += [[w, cut - b, a, b], [w, h - (cut - b), a, cut]]
            del rects[k]
        return rects

    def items_generator(self, seed):
        np.random.seed(seed)
        return self._generate(np.random)

    def items_batch_device(self, seeds, bin_heights=None, device=None, rects=False):
        """items_batch on the GPU (bpp_items_generate): int32 (len(seeds), n, 2) device tensor, identical to the host
        generator for the same seeds.  With rects=True also returns the reference's [w, h, a, b] rows."""
        import ctypes as C
        from ._lib import call
        from .engine import _dev, _devidx, _ptr, _stream
        dev = torch.device("cuda", _devidx(device))
        n = len(seeds)
        seeds_t = _dev(np.asarray(seeds, dtype=np.int64), torch.int64, dev)
        hts = np.full(n, self.bin_height, dtype=np.int32) if bin_heights is None else np.asarray(bin_heights)
        hts_t = _dev(hts.astype(np.int32), torch.int32, dev)
        out = torch.empty((n, self.n, 2), dtype=torch.int32, device=dev)
        rc = torch.empty((n, self.n, 4), dtype=torch.int32, device=dev) if rects else None
        call("bpp_items_generate", self.bin_width, self.n, n, _ptr(seeds_t), _ptr(hts_t), _ptr(out),
             _ptr(rc) if rects else C.c_void_p(0), _stream())
        return (out, rc) if rects else out

    def items_batch(self, seeds, bin_heights=None):
        """(len(seeds), n, 2) int32 (w, h) array for the batched engine; leaves the global RNG untouched.
        bin_heights: optional per-seed generator height (the per-iteration draw of CoachBPP.py:117-119)."""
        out = np.empty((len(seeds), self.n, 2), dtype=np.int32)
        keep = self.bin_height
        try:
            for i, s in enumerate(seeds):
                if bin_heights is not None:
                    self.bin_height = int(bin_heights[i])
                out[i] = np.asarray(self._generate(np.random.RandomState(int(s))), dtype=np.int32)[:, :2]
        finally:
            self.bin_height = keep
        return out
