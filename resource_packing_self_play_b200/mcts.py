"""MCTS with the reference's interface (xw_mcts/MCTS_bpp.py:11-139) on the device-resident search engine.

`MCTS(game, nnet, args)` is a drop-in for CoachBPP.executeEpisode / arena_playing: `getActionProb(state, totalArea,
rewardsList, greedy_a=1)` and `search(state, totalArea, rewardsList)` take the reference's (N+1, H, W) state tensor.
The six dicts (Qsa, Nsa, Ns, Ps, Es, Vs) live on the GPU as one search graph per game (csrc/bpp_engine.cu); any
duck-typed `nnet.predict(board) -> (pi, v)` works as the leaf evaluator (one leaf per step, like the reference),
while an NNetWrapper of this package is evaluated on the device without leaving it.

`BatchedMCTS` runs G games in lockstep (the fast path).
"""
import os

import numpy as np
import torch

from . import _lib
from .engine import REC_REM, SearchEngine, pack_states, ranked_threshold


_TIE_RNG = np.random.default_rng()


class MCTS:
    EPISODES_WARM = 8  # episodes one object can search on a warm graph before it is dropped (see _prepare)

    def __init__(self, game, nnet, args, device=None):  # MCTS_bpp.py:16-26
        self.game = game
        self.nnet = nnet
        self.args = args
        self.device = device if device is not None else getattr(getattr(nnet, "device", None), "index", None)
        self._eng = None
        self._items = None    # (N, 2) dims known for this episode (0,0 = never seen)
        self._consts = None   # (total_area, bl, max_h)
        self._on_device = hasattr(nnet, "dnet") and hasattr(nnet, "predict_batch")

    # ---- engine / episode bookkeeping ----------------------------------------------------------------------------
    def _engine(self):
        if self._eng is None:
            g = self.game
            # One game: memory is no concern, so the pools hold EPISODES_WARM episodes' worth of nodes (the reference's
            # dicts keep growing when one MCTS object plays several games, CoachBPP.arena_playing :233-283, or when
            # search() is called many times); edge_cap = worst case, which can never overflow before the nodes do.
            sims, N, A = int(self.args.numMCTSSims), g.num_items, g.bin_width * g.num_items
            self._node_cap = self.EPISODES_WARM * (sims * N + N + 2)
            nvp = (A + 3) // 4 * 4
            self._eng = SearchEngine(g.bin_width, g.bin_height, N, 1, sims, float(self.args.cpuct), device=self.device,
                                     node_cap=self._node_cap, edge_cap=self._node_cap * (3 * nvp + nvp // 4))
        return self._eng

    def reset(self):
        """forget the search graph (the reference creates a new MCTS object per episode, CoachBPP.py:124)"""
        self._items = None
        self._consts = None

    def _prepare(self, state, totalArea, rewardsList):
        g = self.game
        recs, items = pack_states(np.asarray(state), g.bin_width, g.bin_height, g.num_items)
        rem = int(recs[0, REC_REM])
        items = items[0]
        bl = ranked_threshold(rewardsList, self.args.alpha)
        max_h = int(getattr(g, "max_h", 0)) or int(items[:, 1].max())
        consts = (int(totalArea), None if np.isnan(bl) else bl, max_h)  # NaN (empty rewards list) compares as None
        # same episode <=> same constants and every REMAINING item has the dims this episode started with
        same = self._items is not None and consts == self._consts
        if same:
            for i in range(g.num_items):
                if rem >> i & 1 and tuple(items[i]) != tuple(self._items[i]):
                    same = False  # a remaining item has other dims than this episode's: a different instance
                    break
        eng = self._engine()
        if same:
            # The reference's dicts grow without bound; the device pools do not.  If this call could outgrow them (more
            # than EPISODES_WARM episodes' worth of nodes on one object, e.g. the same instance replayed many times), the
            # graph is dropped: the search then starts cold, which changes visit counts only relative to a reference MCTS
            # object that was itself kept warm that long.
            nodes, _ = eng.graph_sizes()
            if int(nodes[0]) + int(self.args.numMCTSSims) + 2 > self._node_cap:
                same = False
        if not same:
            # new episode: a state key of the reference is the full tensor, so states of another instance never match
            # the old dict entries; dropping the graph is equivalent (see DESIGN.md for the one exception)
            self._items = items.copy()
            self._consts = consts
            # value of a ranked-reward tie r == bl: the reference draws +-1 from numpy's GLOBAL generator when a tie
            # happens (BinPackingGame.py:212); drawn here from a private generator so that the global stream is untouched
            tie = np.array([1 if _TIE_RNG.random() < 0.5 else -1], dtype=np.int8)
            eng.reset(self._items[None], np.array([consts[0]], dtype=np.int32), np.array([bl]), tie)
            eng.set_max_h(np.array([max_h], dtype=np.int32))
        eng.set_roots(recs)
        return eng

    def _run(self, eng, num_sims):
        eng.set_num_sims(num_sims)
        eng.begin_move()
        while True:
            eng.select()
            n = eng.leaf_count()
            if n == 0:
                return
            if self._on_device:
                _, game_ptr, recs_ptr = eng.leaf_buffers()
                pol, val = self.nnet.dnet.forward(recs_ptr, eng.items_wh, game=game_ptr, batch=n)
            else:
                board = eng.leaf_planes(1)[0].cpu().numpy().astype(np.int64)
                pi, v = self.nnet.predict(board)  # MCTS_bpp.py:87
                pi = np.asarray(pi)
                dt = torch.float64 if pi.dtype == np.float64 else torch.float32
                pol = torch.as_tensor(np.ascontiguousarray(pi), dtype=dt).reshape(1, -1).to(eng.device)
                val = torch.tensor([float(np.asarray(v).reshape(-1)[0])], dtype=torch.float64, device=eng.device)
            eng.expand_backup(pol, val)

    # ---- reference API ------------------------------------------------------------------------------------------------
    def getActionProb(self, canonicalBoard, totalArea, rewardsList, greedy_a=1):  # MCTS_bpp.py:28-54
        eng = self._prepare(canonicalBoard, totalArea, rewardsList)
        self._run(eng, int(self.args.numMCTSSims))
        eng.check()
        counts = [int(c) for c in eng.root_counts_host()[0]]
        if greedy_a == 0:
            bestAs = np.array(np.argwhere(counts == np.max(counts))).flatten()
            bestA = np.random.choice(bestAs)
            probs = [0] * len(counts)
            probs[bestA] = 1
            return probs
        counts = [x ** (1. / greedy_a) for x in counts]
        counts_sum = float(sum(counts))
        return [x / counts_sum for x in counts]

    def search(self, canonicalBoard, totalArea, rewardsList):  # MCTS_bpp.py:56-139 — one simulation
        eng = self._prepare(canonicalBoard, totalArea, rewardsList)
        self._run(eng, 1)
        return float(eng.last_values()[0].item())

    def root_counts(self):
        return [int(c) for c in self._engine().root_counts_host()[0]]

    # the reference's public dict attributes (MCTS_bpp.py:20-26), materialised on demand from the device graph.
    # Keys are the reference's state bytes; dims of items placed before the first search call are unknown (0).
    def _dicts(self):
        if self._eng is None or self._items is None:
            return ({},) * 6
        return self._eng.export_dicts(0, self._items)

    Qsa = property(lambda self: self._dicts()[0])
    Nsa = property(lambda self: self._dicts()[1])
    Ns = property(lambda self: self._dicts()[2])
    Ps = property(lambda self: self._dicts()[3])
    Es = property(lambda self: self._dicts()[4])
    Vs = property(lambda self: self._dicts()[5])


class BatchedMCTS:
    """G lockstep games: the batched counterpart of MCTS.getActionProb + CoachBPP.executeEpisode's move loop."""

    def __init__(self, game, nnet, args, G, device=None, edge_cap=0):
        self.game, self.nnet, self.args, self.G = game, nnet, args, G
        if device is None:
            device = getattr(getattr(nnet, "device", None), "index", None)
        self.eng = SearchEngine(game.bin_width, game.bin_height, game.num_items, G, int(args.numMCTSSims),
                                float(args.cpuct), device=device, edge_cap=edge_cap)
        self.steps = 0
        self.use_graphs = os.environ.get("BPP_NO_GRAPHS") is None
        self.fused = os.environ.get("BPP_NO_FUSED_STEP") is None
        self.pipeline = os.environ.get("BPP_NO_PIPELINE") is None
        self._graphs, self._eager_chunks, self.graph_launches = {}, 0, 0

    def close(self):
        """free the engine's device pools and drop the captured step graphs (they point into those pools)"""
        self._graphs = {}
        self.eng.close()

    def play_episodes(self, items_wh, total_area, rewards_list, greedy=False, seed=0, tie=None, record=True, mode=None):
        """Whole self-play episodes (CoachBPP.executeEpisode, CoachBPP.py:50-99, batched) for E >= 1 instances streamed
        through the G resident games: every game runs numMCTSSims simulations per move, chooses (sample ~ counts, or a
        random arg-max when greedy) and plays on at its own pace inside the search kernels, and takes the next instance
        when its episode ends (bpp_engine_play_net_stream); the host only polls for the end of the stream.
        Returns the compact examples as device tensors indexed by episode: roots (N,E,32), counts (N,E,A), actions
        (N,E), moves, r, score, done."""
        E = int(items_wh.shape[0])
        bl = np.full(E, ranked_threshold(rewards_list, self.args.alpha))
        if tie is None:
            tie = np.where(_TIE_RNG.random(E) < 0.5, 1, -1).astype(np.int8)
        if mode is None:
            mode = _lib.CHOOSE_GREEDY if greedy else _lib.CHOOSE_SAMPLE
        out = self.eng.play_net_stream(self.nnet.dnet, items_wh, total_area, bl, tie, mode, seed, record)
        self.steps += out.pop("steps")
        out["done"] = (out["moves"] > 0).to(torch.int32)
        return out

    def reset(self, items_wh, total_area, rewards_list, tie=None):
        bl = np.full(self.G, ranked_threshold(rewards_list, self.args.alpha))
        if tie is None:
            tie = np.where(_TIE_RNG.random(self.G) < 0.5, 1, -1).astype(np.int8)
        self.eng.reset(items_wh, total_area, bl, tie)

    def _run_chunk(self, chunk, cap, count_ptr, game_ptr, recs_ptr):
        """`chunk` lockstep steps (select -> forward -> expand + backup).  Every kernel of a step has a fixed grid and
        reads the leaf count from device memory, so the chunk is captured ONCE per (chunk, select cap) into a CUDA graph
        and replayed: one graph launch instead of 4 * chunk kernel launches through Python/ctypes, which left the GPU
        idle for a quarter of a step.  The first chunks run eagerly (lazy allocations must not happen in a capture)."""
        eng, net = self.eng, self.nnet.dnet

        def body():
            for _ in range(chunk):
                if not self.fused:
                    eng.select()
                net.forward(recs_ptr, self._items_dev, game=game_ptr, count_dev=count_ptr, policy_out=self._pol,
                            value_out=self._val, batch=self.G)
                if self.fused:   # expansion + backup of this step and the descents of the next one in ONE launch
                    eng.expand_select(self._pol, self._val)
                else:
                    eng.expand_backup(self._pol, self._val)
        # the evaluator is identified by a per-DeviceNet serial number, not by its handle's address (a freed handle's
        # address can be reused by the next DeviceNet, whose buffers a replayed graph would then miss)
        key = (chunk, cap, eng.num_sims, self.fused, torch.cuda.current_stream().cuda_stream, net.uid,
               getattr(net, "precision", None))
        graph = self._graphs.get(key) if self.use_graphs else None
        if graph is None and self.use_graphs and self._eager_chunks >= 2:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                body()
            self._graphs[key] = graph
        if graph is not None:
            graph.replay()
            self.graph_launches += (3 if self.fused else 4) * chunk   # [k_search,] trunk, heads, k_expand_[backup|search]
        else:
            body()
            self._eager_chunks += 1
        self.steps += chunk

    def search(self, chunk=8, select_cap=None):
        """numMCTSSims simulations for every running game.  One lockstep step = select (descents + terminal backups
        until every game parks one unexpanded leaf) -> ONE batched forward over the parked leaves, read straight
        from the engine's leaf buffers -> expand + backup.  The leaf count stays on the device (the evaluator and the
        expansion kernel read it there); the host only polls it every `chunk` steps to learn that the move is done."""
        eng, net = self.eng, self.nnet.dnet
        count_ptr, game_ptr, recs_ptr = eng.leaf_buffers()
        # the evaluator reads the item dimensions from a buffer that keeps its address (captured graphs point at it)
        if getattr(self, "_items_dev", None) is None:
            self._items_dev, self._items_src = torch.empty_like(eng.items_wh), None
        if self._items_src is not eng.items_wh:
            self._items_dev.copy_(eng.items_wh)
            self._items_src = eng.items_wh
        if getattr(self, "_pol", None) is None:
            self._pol = torch.empty((self.G, eng.A), dtype=torch.float32, device=eng.device)
            self._val = torch.empty(self.G, dtype=torch.float32, device=eng.device)
        eng.begin_move()
        # a game whose simulations keep ending on terminal states parks no leaf: cap its work per launch so that it
        # does not delay the leaf batch of the others; once few leaves are left the cap is lifted
        if select_cap is None:
            select_cap = int(os.environ.get("BPP_SELECT_CAP", "2"))
        lift = int(os.environ.get("BPP_SELECT_LIFT", "16"))
        eng.set_select_cap(select_cap)
        cap = select_cap
        if self.fused:
            eng.select()   # the first leaves of the move; every later select rides on the expansion launch
        if self.fused and self.use_graphs and self.pipeline:
            # chunk k+1 is queued BEFORE the host learns how chunk k ended (async read-back of the counters into pinned
            # memory + an event): no idle GPU at the chunk boundaries.  Steps queued after the move is complete are
            # no-ops (nothing parked, nobody owes simulations), so the one speculative chunk at the end changes nothing.
            if getattr(self, "_poll", None) is None:
                self._poll = [(torch.zeros(2, dtype=torch.int32).pin_memory(), torch.cuda.Event()) for _ in range(2)]
            prev, k = None, 0
            while True:
                self._run_chunk(chunk, cap, count_ptr, game_ptr, recs_ptr)
                buf, ev = self._poll[k & 1]
                eng.leaf_count_async(buf)
                ev.record()
                if prev is not None:
                    pbuf, pev = prev
                    pev.synchronize()
                    n, unfinished = int(pbuf[0]), int(pbuf[1])
                    if n == 0 and unfinished == 0:
                        break
                    if n < self.G // lift and cap != 0:
                        eng.set_select_cap(0)
                        cap = 0
                prev, k = (buf, ev), k + 1
            n = eng.leaf_count()  # synchronises and settles the engine's host-side pairing state
            assert n == 0 and eng.unfinished() == 0
        else:
            while True:
                self._run_chunk(chunk, cap, count_ptr, game_ptr, recs_ptr)
                n = eng.leaf_count()
                if n == 0 and eng.unfinished() == 0:  # nothing parked, nobody capped: the move is complete
                    break
                if n < self.G // lift and cap != 0:
                    eng.set_select_cap(0)
                    cap = 0
        eng.set_select_cap(0)
        return eng.root_counts()
