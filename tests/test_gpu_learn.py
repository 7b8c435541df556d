"""-m gpu: the outer loop: CoachBPP.learn with the reference's bookkeeping, the learner step and the
checkpoint / rewards-list files."""
import os
import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(tmp_path, **kw):
    from resource_packing_self_play_b200.coach import CoachBPP
    from resource_packing_self_play_b200.game import BinPackingGame, ItemsGenerator
    from resource_packing_self_play_b200.nnet import NNetWrapper
    from resource_packing_self_play_b200.utils import dotdict
    W, H, N = 15, 15, 10
    d = dict(numIters=2, numEps=3, iterStepThreshold=1, updateThreshold=0.6, maxlenOfQueue=200000, numMCTSSims=12,
             arenaCompare=2, cpuct=1, alpha=0.75, seed=100, numScoresForRank=4, numItems=N, numBins=1, binH_min=2,
             binH=15, lr=0.001, dropout=0.1, epochs=2, batch_size=8, cuda=True, num_channels=256, num_items=N,
             num_bins=1, checkpoint=str(tmp_path), load_model=False, numItersForTrainExamplesHistory=2)
    d.update(kw)
    args = dotdict(d)
    g = BinPackingGame(W, H, N, 1)
    gen = ItemsGenerator(W, H, N)
    torch.manual_seed(0)
    net = NNetWrapper(g, args)
    coach = CoachBPP(g, net, gen.items_generator(args.seed), W * H, gen, args)
    return coach, net, args


def test_learn_runs_the_reference_loop(tmp_path):
    coach, net, args = _setup(tmp_path)
    before = {k: v.clone() for k, v in net.nnet.state_dict().items()}
    coach.learn()
    # rewards buffer: one score per episode, trimmed to numScoresForRank by dropping minima (CoachBPP.py:134-139)
    assert len(coach.rewards_list) == args.numScoresForRank
    assert len(coach.trainExamplesHistory) == 2 and all(len(h) >= 3 for h in coach.trainExamplesHistory)
    state, pi, r = coach.trainExamplesHistory[-1][0]
    assert state.shape == (11, 15, 15) and len(pi) == 150 and r in (-1, 1)
    # iteration 2 > iterStepThreshold: greedy episodes -> one-hot policies (MCTS_bpp.py:43-49)
    assert sorted(set(pi)) in ([0, 1], [0.0, 1.0])
    # files: temp checkpoint in the reference format + the rewards pickle (CoachBPP.py:172,196-202)
    ck = torch.load(os.path.join(str(tmp_path), "temp.pth.tar"), map_location="cpu")
    assert list(ck.keys()) == ["state_dict"]
    rl = pickle.load(open(os.path.join(str(tmp_path), "rewards_list_10_items.pkl"), "rb"))
    assert rl == coach.rewards_list
    # the learner moved the weights and the CUDA forward follows the torch module
    assert any(not torch.equal(before[k], v) for k, v in net.nnet.state_dict().items())
    st = coach.trainExamplesHistory[-1][0][0]
    pi_k, v_k = net.predict(st)
    with torch.no_grad():
        lp, tv = net.nnet(torch.from_numpy(st.astype(np.float32))[None].to(net.device))
    assert np.abs(pi_k - lp.exp()[0].cpu().numpy()).max() < 2e-2 and abs(float(v_k[0]) - float(tv)) < 2e-2


def test_losses_match_reference_definitions(tmp_path):
    coach, net, args = _setup(tmp_path)
    t = torch.tensor([[0.25, 0.75], [1.0, 0.0]])
    o = torch.log(torch.tensor([[0.5, 0.5], [0.9, 0.1]]))
    assert torch.isclose(net.loss_pi(t, o), -(t * o).sum() / 2)          # NNet.py:87-88
    assert torch.isclose(net.loss_v(torch.tensor([1.0, -1.0]), torch.tensor([[0.5], [0.0]])), torch.tensor(0.625))


def test_examples_save_and_load_round_trip(tmp_path):
    coach, net, args = _setup(tmp_path, numIters=1, numEps=1)
    coach.learn()
    coach.saveTrainExamples(0)
    args.load_folder_file = (str(tmp_path), coach.getCheckpointFile(0))
    n = len(coach.trainExamplesHistory[0])
    coach.trainExamplesHistory = []
    coach.loadTrainExamples()
    assert coach.skipFirstSelfPlay and len(coach.trainExamplesHistory[0]) == n


def test_learn_batched_single_rank(tmp_path):
    coach, net, args = _setup(tmp_path, numMCTSSims=10, epochs=1, batch_size=32, numScoresForRank=20,
                              iterStepThreshold=1)
    before = {k: v.clone() for k, v in net.nnet.state_dict().items()}
    log = coach.learn_batched(games_per_iter=48, num_iters=2)
    assert len(log) == 2 and log[0]["games"] == 48 and log[0]["examples"] >= 48 * 4
    assert log[1]["history_examples"] == log[0]["examples"] + log[1]["examples"]
    assert len(coach.rewards_list) == 20 and min(coach.rewards_list) >= 0.0
    # the buffer keeps the LARGEST scores (minima are dropped, CoachBPP.py:136-139)
    assert np.isfinite(log[1]["loss_pi"]) and np.isfinite(log[1]["loss_v"])
    assert any(not torch.equal(before[k], v) for k, v in net.nnet.state_dict().items())
    assert os.path.exists(os.path.join(str(tmp_path), "temp.pth.tar"))


def test_env_planes_match_host_unpack():
    from resource_packing_self_play_b200.engine import EnvOps, pack_states, unpack_states
    from resource_packing_self_play_b200.game import ItemsGenerator
    rng = np.random.RandomState(5)
    W, H, N, n = 20, 20, 10, 40
    items = ItemsGenerator(W, H, N).items_batch(np.arange(n), rng.randint(2, 21, size=n))
    recs = np.zeros((n, 32), dtype=np.uint32)
    recs[:, :H] = rng.randint(0, 1 << W, size=(n, H))
    recs[:, 28] = rng.randint(0, 1 << N, size=n)
    planes = EnvOps(W, H, N).planes(recs, items).cpu().numpy()
    want = unpack_states(recs, items, W, H, N)
    assert planes.shape == (n, N + 1, H, W) and np.array_equal(planes, want.astype(np.float32))
    r2, i2 = pack_states(want, W, H, N)
    assert np.array_equal(r2[:, :H], recs[:, :H]) and np.array_equal(r2[:, 28], recs[:, 28])
