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
    def executeEpisodesBatched(self, items_batch, total_areas, greedy=False, seed=None, expand=True):
        """G episodes in lockstep on the device with the batched leaf evaluator.

        items_batch: (G, N, 2) int (w, h); total_areas: (G,) int.  All games share this call's `self.rewards_list`
        (the reference appends to it after every episode, so inside one call the ranked-reward threshold is the one at
        the start of the batch).  Returns (examples, scores, outcomes): examples is the reference's list of
        (state (N+1,H,W) int64, pi list, r) when expand=True, else a dict of compact arrays."""
        items_batch = np.asarray(items_batch, dtype=np.int32)
        G = items_batch.shape[0]
        g = self.game
        N, A = g.num_items, g.getActionSize()
        bm = getattr(self, "_bm", None)
        if bm is None or bm.G != G or bm.nnet is not self.nnet:
            bm = self._bm = BatchedMCTS(g, self.nnet, self.args, G)
        bm.reset(items_batch, np.asarray(total_areas, dtype=np.int32), self.rewards_list)
        eng = bm.eng
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        roots, counts, acts = [], [], []
        for move in range(N):
            roots.append(eng.roots())
            counts.append(bm.search())
            act = eng.choose(_lib.CHOOSE_GREEDY if greedy else _lib.CHOOSE_SAMPLE, seed + move)
            acts.append(act)
            eng.advance(act)
        eng.check()
        st = {k: v.cpu().numpy() for k, v in eng.status().items()}
        roots = torch.stack(roots).cpu().numpy().view(np.uint32)      # (N, G, 32)
        counts = torch.stack(counts).cpu().numpy()                      # (N, G, A)
        acts = torch.stack(acts).cpu().numpy()                          # (N, G)
        moves, r, score = st["moves"], st["r"], st["score"]
        assert (st["done"] == 1).all()
        if not expand:
            return {"roots": roots, "counts": counts, "actions": acts, "moves": moves, "r": r,
                    "items": items_batch}, score, r
        examples = []
        for gi in range(G):
            m = int(moves[gi])
            states = unpack_states(roots[:m, gi], np.repeat(items_batch[gi][None], m, axis=0), g.bin_width,
                                   g.bin_height, N)
            for k in range(m):
                c = counts[k, gi].astype(np.float64)
                if greedy:
                    pi = [0] * A
                    pi[int(acts[k, gi])] = 1  # one-hot of the arg-max the device drew (MCTS_bpp.py:43-49)
                else:
                    pi = list(c / c.sum())
                examples.append((states[k], pi, int(r[gi])))
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

    def arena_sweep(self, pnet, nnet, seeds, bin_heights=None, seed=0):
        """Batched arena (BASELINE.json configs[4]): every seed is played greedily with both nets in lockstep.
        Returns (p_scores, n_scores, accept) with accept as in arena_playing."""
        seeds = np.asarray(seeds)
        items = self.gen.items_batch(seeds, bin_heights)
        hts = np.full(len(seeds), self.gen.bin_height) if bin_heights is None else np.asarray(bin_heights)
        areas = (hts * self.gen.bin_width).astype(np.int32)
        out = []
        keep = self.nnet
        try:
            for net in (pnet, nnet):
                self.nnet = net
                _, score, _ = self.executeEpisodesBatched(items, areas, greedy=True, seed=seed, expand=False)
                out.append(score)
        finally:
            self.nnet = keep
        return out[0], out[1], 1 if np.mean(out[1]) >= np.mean(out[0]) else 0
