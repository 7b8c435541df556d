"""Self-play driver with the reference's Coach interface (xw_mcts/CoachBPP.py:22-291).

`CoachBPP(game, nnet, items_list, total_area, gen, args, saved_rewards_list=[])` keeps the reference's constructor,
attributes and methods (`executeEpisode`, `learn`, `arena_playing`, `save_rewards_list`, `saveTrainExamples`,
`loadTrainExamples`).  The single-game methods drive the device-resident search through the drop-in `MCTS` class;
`executeEpisodesBatched` / `arena_sweep` are the lockstep fast paths (thousands of games per call, games sharded over
ranks with torch.distributed when it is initialised).

W&B logging of the reference (`wandb.log`, CoachBPP.py:143-147) goes through `self.log_fn` (a no-op unless wandb is
importable and enabled) so that importing this module never needs a W&B login.
"""
import logging
import os
import pickle
import random
import sys
from collections import deque
from pickle import Pickler, Unpickler
from random import shuffle

import numpy as np
import torch

from . import _lib
from .engine import unpack_states
from .mcts import MCTS, BatchedMCTS

log = logging.getLogger(__name__)


def _default_log_fn():
    if os.environ.get("WANDB_MODE", "") == "disabled" or "wandb" not in sys.modules:
        return lambda metrics, step=None: None
    import wandb
    return lambda metrics, step=None: wandb.log(metrics, step=step) if wandb.run is not None else None


def trim_rewards(rewards_list, new_scores, cap):
    """CoachBPP.py:134-139 for a whole batch of episodes: append every score, then `while len > cap: pop(argmin)`.
    np.argmin returns the FIRST minimum, so the loop removes the (len - cap) smallest entries, earliest first among
    equals, and the survivors keep their order: one stable sort instead of a quadratic loop."""
    rl = np.asarray(list(rewards_list) + [float(x) for x in new_scores], dtype=np.float64)
    extra = len(rl) - int(cap)
    if extra > 0:
        drop = np.lexsort((np.arange(len(rl)), rl))[:extra]
        rl = np.delete(rl, drop)
    return [float(x) for x in rl]


class CoachBPP:
    def __init__(self, game, nnet, items_list, total_area, gen, args, saved_rewards_list=[]):  # CoachBPP.py:28-48
        self.game = game
        self.nnet = nnet
        self.args = args
        self.pnet = self.nnet.__class__(self.game, self.args)  # the competitor network
        self.items_list = items_list
        self.items_total_area = total_area
        self.rewards_list = list(saved_rewards_list)
        self.ep_score = 0
        self.mcts = MCTS(self.game, self.nnet, self.args)
        self.trainExamplesHistory = []
        self.skipFirstSelfPlay = False
        self.gen = gen
        random.seed(args.seed)
        self.seeds = random.sample(range(0, 10000), 500)
        self.log_fn = _default_log_fn()
        self.curPlayer = 1

    # ------------------------------------------------------------------------------------------------------------------
    def executeEpisode(self, greedy=False):
        """One self-play episode, CoachBPP.py:50-99: returns [(state, pi, r)] with r the final ranked outcome."""
        trainExamples = []
        board = self.game.getInitBoard()
        items_list_board = self.game.getInitItems(self.items_list)
        self.curPlayer = 1
        episodeStep = 0
        while True:
            episodeStep += 1
            bin_items_state = self.game.getBinItem(board, items_list_board)
            if greedy:
                pi = self.mcts.getActionProb(bin_items_state, self.items_total_area, self.rewards_list, greedy_a=0)
            else:
                pi = self.mcts.getActionProb(bin_items_state, self.items_total_area, self.rewards_list)
            trainExamples.append([bin_items_state, pi, None])
            np.random.seed()  # the reference re-seeds from OS entropy before every move (CoachBPP.py:86)
            action = np.random.choice(len(pi), p=pi)
            board, items_list_board = self.game.getNextState(board, action, items_list_board)
            next_bin_items_state = self.game.getBinItem(board, items_list_board)
            r, score = self.game.getGameEnded(next_bin_items_state, self.items_total_area, self.rewards_list,
                                              self.args.alpha)
            if r != 0:
                self.ep_score = score
                return [(x[0], x[1], r) for x in trainExamples]

    # ------------------------------------------------------------------------------------------------------------------
    RESIDENT_GAMES = 4096  # default number of games resident on the device; larger batches stream through them

    def executeEpisodesBatched(self, items_batch, total_areas, greedy=False, seed=None, expand=True, on_device=False,
                               per_move=False, record=True, resident=None):
        """G episodes at once on the device with the batched leaf evaluator.

        items_batch: (G, N, 2) int (w, h); total_areas: (G,) int.  All games share this call's `self.rewards_list`
        (the reference appends to it after every episode, so inside one call the ranked-reward threshold is the one at
        the start of the batch).  Every game runs at its own pace (BatchedMCTS.play_episodes: visit counts -> choose ->
        play inside the search kernels); per_move=True runs the older move-synchronous loop (search / choose / advance
        per move for all games) that gives identical results for the same seed.  Returns (examples, scores, outcomes):
        examples is the reference's list of (state (N+1,H,W) int64, pi list, r) when expand=True, else a dict of compact
        arrays (numpy, or device tensors that never visit the host when on_device=True; scores and outcomes are then
        device tensors too)."""
        items_dev = None
        if isinstance(items_batch, torch.Tensor):  # e.g. from ItemsGenerator.items_batch_device
            items_dev = items_batch.to(torch.int32)
            items_batch = items_dev.cpu().numpy()
        items_batch = np.asarray(items_batch, dtype=np.int32)
        E = items_batch.shape[0]
        # games resident on the device; a larger batch is STREAMED through them (a game whose episode ends takes the next
        # instance inside the search kernel): memory is bounded by `resident`, results do not depend on it
        G = E if per_move else min(E, int(resident or getattr(self.args, "residentGames", 0) or self.RESIDENT_GAMES))
        g = self.game
        N, A = g.num_items, g.getActionSize()
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little") >> 1
        areas = np.asarray(total_areas, dtype=np.int32)
        # greedy: False = sample ~ counts (CoachBPP.py:86-87); True = random arg-max (MCTS_bpp.py:43-49); "first" = first
        # arg-max (deterministic)
        mode = _lib.CHOOSE_ARGMAX_FIRST if greedy == "first" else (_lib.CHOOSE_GREEDY if greedy else _lib.CHOOSE_SAMPLE)
        edge_cap = 0
        while True:
            bm = getattr(self, "_bm", None)
            if bm is None or bm.G != G or edge_cap:
                if bm is not None:  # free the old pools (and the graphs captured on them) BEFORE sizing the new ones
                    bm.close()
                    self._bm = bm = None
                bm = self._bm = BatchedMCTS(g, self.nnet, self.args, G, edge_cap=edge_cap)
            bm.nnet = self.nnet  # the search engine is independent of the evaluator: swapping nets keeps the pools
            eng = bm.eng
            try:
                if per_move:
                    bm.reset(items_dev if items_dev is not None else items_batch, areas, self.rewards_list)
                    roots, counts, acts = [], [], []
                    for move in range(N):
                        roots.append(eng.roots())
                        counts.append(bm.search())
                        act = eng.choose(mode, seed)
                        acts.append(act)
                        eng.advance(act)
                    roots, counts, acts = torch.stack(roots), torch.stack(counts), torch.stack(acts)
                    live = torch.arange(N, device=eng.device)[:, None] < eng.status()["moves"][None, :]
                    roots = roots * live[:, :, None]      # rows of moves a game did not play: like play_episodes
                else:
                    ep = bm.play_episodes(items_dev if items_dev is not None else items_batch, areas, self.rewards_list,
                                          mode=mode, seed=seed, record=record)
                    roots, counts, acts = ep["roots"], ep["counts"], ep["actions"]
                eng.check()
                st = eng.status() if per_move else ep
                break
            except _lib.BppError as err:
                # a game outgrew its edge pool (sized from measured episodes, not from the worst case): re-create the
                # engine with a larger pool and replay the batch; results do not depend on the pool size
                if err.code != -4 or eng.edge_cap >= eng.edge_cap_worst:
                    raise
                edge_cap = min(eng.edge_cap_worst, 2 * eng.edge_cap)
        if on_device and not expand:
            return {"roots": roots, "counts": counts, "actions": acts, "moves": st["moves"], "r": st["r"],
                    "items": items_dev if items_dev is not None else torch.from_numpy(items_batch).to(eng.device)}, \
                st["score"], st["r"]
        st = {k: v.cpu().numpy() for k, v in st.items() if k in ("moves", "r", "score")}
        roots = roots.cpu().numpy().view(np.uint32)      # (N, G, 32)
        counts = counts.cpu().numpy()                      # (N, G, A)
        acts = acts.cpu().numpy()                          # (N, G)
        moves, r, score = st["moves"], st["r"], st["score"]
        assert (moves > 0).all()
        if not expand:
            return {"roots": roots, "counts": counts, "actions": acts, "moves": moves, "r": r,
                    "items": items_batch}, score, r
        # the reference's example list, game by game and move by move (CoachBPP.py:80,99); all array work is vectorised
        gi, mi = np.nonzero(np.arange(N)[None, :] < moves[:, None])         # game-major, then move
        states = unpack_states(roots[mi, gi], items_batch[gi], g.bin_width, g.bin_height, N)
        if greedy:   # one-hot of the arg-max the device drew (MCTS_bpp.py:43-49)
            pis = np.zeros((len(gi), A), dtype=np.int64)
            pis[np.arange(len(gi)), acts[mi, gi]] = 1
        else:
            c = counts[mi, gi].astype(np.float64)
            pis = c / c.sum(axis=1, keepdims=True)
        pis = pis.tolist()
        rr = r[gi].tolist()
        examples = [(states[k], pis[k], rr[k]) for k in range(len(gi))]
        return examples, score, r

    # ------------------------------------------------------------------------------------------------------------------
    def learn(self):
        """CoachBPP.py:101-196: numIters x (numEps self-play episodes -> train -> save)."""
        for i in range(1, self.args.numIters + 1):
            log.info(f'Starting Iter #{i} ...')
            ep_scores = []
            seeds_iter = []
            np.random.seed()
            self.gen.bin_height = np.random.randint(self.args.binH_min, self.args.binH + 1)
            self.items_total_area = self.gen.bin_height * self.gen.bin_width
            if not self.skipFirstSelfPlay or i > 1:
                iterationTrainExamples = deque([], maxlen=self.args.maxlenOfQueue)
                for _ in range(self.args.numEps):
                    self.mcts = MCTS(self.game, self.nnet, self.args)  # reset search tree
                    generator_seed = np.random.randint(int(1e5))
                    items_list = self.gen.items_generator(generator_seed)
                    seeds_iter.append(generator_seed)
                    self.items_list = np.copy(items_list)
                    iterationTrainExamples += self.executeEpisode(i > self.args.iterStepThreshold)
                    ep_scores.append(self.ep_score)
                    self.rewards_list.append(self.ep_score)
                while len(self.rewards_list) > self.args.numScoresForRank:
                    idx = np.argmin(self.rewards_list)  # drop the smallest score (CoachBPP.py:136-139)
                    self.rewards_list.pop(idx)
                self.log_fn({"iter mean reward": np.mean(ep_scores)}, step=i)
                percentage_optim = sum([item == 1.0 for item in ep_scores]) / len(ep_scores)
                self.log_fn({"optimality percentage": percentage_optim, "min reward": np.min(ep_scores),
                             "max reward": np.max(ep_scores)}, step=i)
                self.trainExamplesHistory.append(iterationTrainExamples)
            if len(self.trainExamplesHistory) > self.args.numItersForTrainExamplesHistory:
                log.warning("Removing the oldest entry in trainExamples. len(trainExamplesHistory) = "
                            f"{len(self.trainExamplesHistory)}")
                self.trainExamplesHistory.pop(0)
            trainExamples = []
            for e in self.trainExamplesHistory:
                trainExamples.extend(e)
            shuffle(trainExamples)
            self.nnet.save_checkpoint(folder=self.args.checkpoint, filename='temp.pth.tar')
            self.nnet.train(trainExamples)
            self.seeds_iter = seeds_iter
            self.save_rewards_list()

    # ------------------------------------------------------------------------------------------------------------------
    def learn_batched(self, games_per_iter, num_iters=None, checkpoint=True):
        """Lockstep / multi-GPU counterpart of learn() (BASELINE.json configs[3]).  Per iteration: rank 0 draws the
        generator height and `games_per_iter` seeds (CoachBPP.py:117-127) and broadcasts them; every rank plays its
        contiguous shard of the games in lockstep (no collective); scores and compact examples are all-gathered so that
        every rank holds the same rewards_list and replay history; the learner runs data-parallel with one flat
        gradient all-reduce per step.  Differences from learn(), by construction: all games of an iteration see the
        rewards_list of the iteration's start, and the scores are appended in global game order afterwards.
        Returns a list of per-iteration dicts (timings in seconds, mean score, losses)."""
        import time
        from . import distributed as D
        from .engine import EnvOps
        rank, ws = D.world()
        dev = self.nnet.device
        g = self.game
        if ws > 1:
            D.broadcast_parameters(self.nnet.nnet)
            self.nnet.sync_weights()
        ops = EnvOps(g.bin_width, g.bin_height, g.num_items, dev.index or 0)
        history = getattr(self, "_compact_history", [])
        out = []
        for i in range(1, (num_iters or self.args.numIters) + 1):
            t0 = time.perf_counter()
            hdr = torch.zeros(games_per_iter + 1, dtype=torch.int64, device=dev)
            if rank == 0:
                np.random.seed()
                h = np.random.randint(self.args.binH_min, self.args.binH + 1)
                hdr[0] = int(h)
                hdr[1:] = torch.from_numpy(np.random.randint(int(1e5), size=games_per_iter)).to(dev)
            if ws > 1:
                torch.distributed.broadcast(hdr, 0)
            h = int(hdr[0])
            seeds = hdr[1:].cpu().numpy()
            self.gen.bin_height = h
            self.items_total_area = h * self.gen.bin_width
            lo, hi = D.shard_range(games_per_iter, rank, ws)
            items = self.gen.items_batch_device(seeds[lo:hi], device=dev.index)
            areas = np.full(hi - lo, self.items_total_area, dtype=np.int32)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            # the examples stay on the device from the search kernels to the learner: compact arrays -> all-gather over
            # the game axis (NCCL) -> one row per played move -> replay history, all as device tensors
            compact, score, r = self.executeEpisodesBatched(items, areas, greedy=i > self.args.iterStepThreshold,
                                                            expand=False, on_device=True)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            compact = D.gather_examples(compact, dev)
            scores = D.all_gather_variable(score).cpu().numpy()
            # rewards buffer (CoachBPP.py:134-139): append every score, then drop minima until numScoresForRank remain
            # (repeated `pop(argmin)` == remove the smallest, earliest-first; survivors keep their order)
            self.rewards_list = trim_rewards(self.rewards_list, scores, self.args.numScoresForRank)
            self.log_fn({"iter mean reward": float(np.mean(scores)),
                         "optimality percentage": float(np.mean(scores == 1.0)),
                         "min reward": float(np.min(scores)), "max reward": float(np.max(scores))}, step=i)
            # flatten the iteration's examples: one row per played move (move-major, then game: np.nonzero order)
            moves = compact["moves"]
            N, G = compact["roots"].shape[0], compact["roots"].shape[1]
            played = torch.arange(N, device=dev)[:, None] < moves[None, :]
            mi, gi = played.nonzero(as_tuple=True)
            counts = compact["counts"][mi, gi].to(torch.float32)
            if i > self.args.iterStepThreshold:
                pis = torch.zeros_like(counts)
                pis[torch.arange(len(mi), device=dev), compact["actions"][mi, gi].long()] = 1.0
            else:
                pis = counts / counts.sum(dim=1, keepdim=True)
            history.append({"recs": compact["roots"][mi, gi], "items": compact["items"][gi].to(torch.int32), "pis": pis,
                            "vs": compact["r"][gi].to(torch.float32)})
            if len(history) > self.args.numItersForTrainExamplesHistory:
                history.pop(0)
            self._compact_history = history

            def cat(key, dt):
                return torch.cat([torch.as_tensor(e[key]).to(dev).view(dt) if not isinstance(e[key], torch.Tensor)
                                  else e[key] for e in history])
            recs_t, items_t = cat("recs", torch.int32), cat("items", torch.int32)
            pis_t, vs_t = cat("pis", torch.float32), cat("vs", torch.float32)
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            if checkpoint and rank == 0:
                self.nnet.save_checkpoint(folder=self.args.checkpoint, filename='temp.pth.tar')
            l_pi, l_v = self.nnet.train_compact(recs_t, items_t, pis_t, vs_t, ops)
            torch.cuda.synchronize()
            t4 = time.perf_counter()
            if checkpoint and rank == 0:
                self.save_rewards_list()
            out.append({"iter": i, "games": int(G), "examples": int(len(mi)), "history_examples": int(recs_t.shape[0]),
                        "mean_score": float(np.mean(scores)), "t_setup": t1 - t0, "t_selfplay": t2 - t1,
                        "t_gather": t3 - t2, "t_train": t4 - t3, "loss_pi": l_pi, "loss_v": l_v})
        return out

    def save_rewards_list(self):  # CoachBPP.py:198-202
        file_n = 'rewards_list_' + str(self.args.numItems) + '_items.pkl'
        if not os.path.exists(self.args.checkpoint):
            os.makedirs(self.args.checkpoint)
        with open(os.path.join(self.args.checkpoint, file_n), 'wb') as f:
            pickle.dump(self.rewards_list, f)

    def getCheckpointFile(self, iteration):  # :204-205
        return 'checkpoint_' + '.pth.tar'

    def saveTrainExamples(self, iteration):  # :207-214
        folder = self.args.checkpoint
        if not os.path.exists(folder):
            os.makedirs(folder)
        filename = os.path.join(folder, self.getCheckpointFile(iteration) + ".examples")
        with open(filename, "wb+") as f:
            Pickler(f).dump(self.trainExamplesHistory)

    def loadTrainExamples(self):  # :216-231 (without the interactive prompt: a missing file is an error)
        modelFile = os.path.join(self.args.load_folder_file[0], self.args.load_folder_file[1])
        examplesFile = modelFile + ".examples"
        if not os.path.isfile(examplesFile):
            raise FileNotFoundError(f'File "{examplesFile}" with trainExamples not found!')
        with open(examplesFile, "rb") as f:
            self.trainExamplesHistory = Unpickler(f).load()
        self.skipFirstSelfPlay = True

    # ------------------------------------------------------------------------------------------------------------------
    def _play_greedy(self, mcts, items_list):
        board = self.game.getInitBoard()
        items_list_board = self.game.getInitItems(items_list)
        state = self.game.getBinItem(board, items_list_board)
        game_ended, score = 0, None
        while game_ended == 0:
            pi = mcts.getActionProb(state, self.items_total_area, self.rewards_list, greedy_a=0)
            action = np.random.choice(len(pi), p=pi)
            board, items_list_board = self.game.getNextState(board, action, items_list_board)
            state = self.game.getBinItem(board, items_list_board)
            game_ended, score = self.game.getGameEnded(state, self.items_total_area, self.rewards_list, self.args.alpha)
        return score

    def arena_playing(self, pmcts, nmcts, seeds_iter):
        """CoachBPP.py:233-291: play arenaCompare seeds greedily with the previous and the new net; 1 iff the new
        net's mean raw score is at least the previous net's."""
        p_scores, n_scores = [], []
        random.seed()
        arena_seeds = random.sample(seeds_iter, self.args.arenaCompare)
        for t in range(self.args.arenaCompare):
            items_list = self.gen.items_generator(arena_seeds[t])
            p_scores.append(self._play_greedy(pmcts, np.copy(items_list)))
            n_scores.append(self._play_greedy(nmcts, np.copy(items_list)))
        return 1 if np.mean(n_scores) >= np.mean(p_scores) else 0

    def arena_sweep(self, pnet, nnet, seeds, bin_heights=None, seed=0, choose="greedy"):
        """Batched arena (BASELINE.json configs[4]): every seed is played greedily with both nets at once.
        Returns (p_scores, n_scores, accept) with accept as in arena_playing.  choose="greedy" draws a uniformly random
        arg-max like MCTS_bpp.py:43-49; choose="first" takes the first arg-max (deterministic; parity tests)."""
        seeds = np.asarray(seeds)
        items = self.gen.items_batch_device(seeds, bin_heights, device=self.nnet.device.index)
        hts = np.full(len(seeds), self.gen.bin_height) if bin_heights is None else np.asarray(bin_heights)
        areas = (hts * self.gen.bin_width).astype(np.int32)
        out = []
        keep = self.nnet
        try:
            for net in (pnet, nnet):
                self.nnet = net
                _, score, _ = self.executeEpisodesBatched(items, areas, greedy=True if choose == "greedy" else "first",
                                                          seed=seed, expand=False, on_device=True,
                                                          record=False)  # only the G scores leave the device
                out.append(score.cpu().numpy())
        finally:
            self.nnet = keep
        return out[0], out[1], 1 if np.mean(out[1]) >= np.mean(out[0]) else 0
