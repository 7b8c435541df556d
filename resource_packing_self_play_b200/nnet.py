"""Leaf evaluator with the reference's NeuralNet interface (xw_mcts/NeuralNet.py:14-50 as implemented by
binpacking/pytorch/NNet.py:17-111).

`predict` / `predict_batch` run the hand-written CUDA forward (csrc/bpp_net.cu) in bf16 with fp32 accumulation.  The
torch module below has the reference's architecture and parameter names (BinpackingNNet.py:15-81) so that reference
checkpoints (`{'state_dict': ...}`) load unchanged; it is the parameter container (and the learner's autograd graph,
SURVEY.md §8(f) rank 1), not the inference path.
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._lib import call
from .distributed import allreduce_gradients
from .engine import _devidx, _ptr, _stream, pack_states
from .utils import AverageMeter


class ResidualBlock(nn.Module):  # BinpackingNNet.py:15-27 (pre-activation residual block)
    def __init__(self, channels):
        super().__init__()
        self.conv0 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.conv1 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)

    def forward(self, x):
        y = self.conv0(F.relu(x))
        y = self.conv1(F.relu(y))
        return y + x


class ConvSequence(nn.Module):  # BinpackingNNet.py:29-48
    def __init__(self, input_shape, out_channels):
        super().__init__()
        self._input_shape = input_shape
        self._out_channels = out_channels
        self.conv = nn.Conv2d(input_shape[0], out_channels, kernel_size=3, padding=1)
        self.res_block0 = ResidualBlock(out_channels)
        self.res_block1 = ResidualBlock(out_channels)

    def forward(self, x):
        x = F.max_pool2d(self.conv(x), kernel_size=3, stride=2, padding=1)
        return self.res_block1(self.res_block0(x))

    def get_output_shape(self):
        _c, h, w = self._input_shape
        return (self._out_channels, (h + 1) // 2, (w + 1) // 2)


class BinPackingNNet(nn.Module):  # BinpackingNNet.py:50-81
    def __init__(self, game, args):
        super().__init__()
        self.board_h, self.board_w = game.getBoardSize()
        self.action_size = game.getActionSize()
        self.args = args
        self.in_channels = args.num_items + args.num_bins
        shape = (self.in_channels, self.board_h, self.board_w)
        seqs = []
        for out_channels in (16, 32, 32):
            seq = ConvSequence(shape, out_channels)
            shape = seq.get_output_shape()
            seqs.append(seq)
        self.conv_seqs = nn.ModuleList(seqs)
        self.hidden_fc = nn.Linear(shape[0] * shape[1] * shape[2], 256)
        self.logits_fc = nn.Linear(256, self.action_size)
        self.value_fc = nn.Linear(256, 1)

    def forward(self, x):
        for seq in self.conv_seqs:
            x = seq(x)
        x = F.relu(torch.flatten(x, start_dim=1))
        x = F.relu(self.hidden_fc(x))
        return F.log_softmax(self.logits_fc(x), dim=1), torch.tanh(self.value_fc(x))


class DeviceNet:
    """Handle of the CUDA forward (bpp_net_*)."""
    _serial = 0

    def __init__(self, W, H, N, max_batch, device=None):
        _lib.load()
        DeviceNet._serial += 1
        self.uid = DeviceNet._serial
        device = _devidx(device)
        self.W, self.H, self.N, self.A = W, H, N, W * N
        self.max_batch = max_batch
        self.device = torch.device("cuda", device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            call("bpp_net_create", W, H, N, max_batch, device, C.byref(h))
        self._h = h
        self.precision = "bf16"

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().bpp_net_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_precision(self, mode):
        """'bf16' (tcgen05 kernel: bf16 weights/activations, fp32 accumulate), 'fp32', or 'bf16_simt'"""
        call("bpp_net_set_precision", self._h, {"bf16": 0, "fp32": 1, "bf16_simt": 2, "bf16x3": 3}[mode])
        self.precision = mode

    def profile(self):
        arr = (C.c_int64 * 8)()
        call("bpp_net_profile", self._h, arr)
        keys = ["input", "weights", "mma_issue", "mma_wait", "epilogue", "pool", "heads", "total"]
        return {k: int(v) for k, v in zip(keys, arr)}

    def profile_roles(self):
        arr = (C.c_int64 * 32)()
        call("bpp_net_profile_roles", self._h, arr)
        if self.grid_row():   # k_net_gr stages 0..3 (thread 0 of CTA 0): phases of its groups, prologue, whole stage
            keys = ["input", "layers", "output", "prologue", "stage", "unused5", "unused6", "total"]
        else:
            keys = ["input", "weights", "mma_issue", "mma_wait", "epilogue", "pool", "output", "total"]
        return [{k: int(arr[8 * r + i]) for i, k in enumerate(keys)} for r in range(4 if self.grid_row() else 3)]

    def grid_row(self):
        """True when the bf16 mode runs the grid-row stage kernels (bpp_net_gr.cuh)"""
        return bool(_lib.load().bpp_net_grid_row(self._h))

    def load_state_dict(self, state_dict):
        for name, t in state_dict.items():
            a = np.ascontiguousarray(t.detach().to("cpu", torch.float32).numpy())
            call("bpp_net_set_param", self._h, name.encode(), a.ctypes.data_as(C.c_void_p), a.size)
        with torch.cuda.device(self.device):
            call("bpp_net_commit", self._h, _stream())

    def forward(self, recs, items_wh, game=None, count_dev=None, policy_out=None, value_out=None, batch=None):
        """recs int32/uint32 (B, 32) device, items_wh int32 (*, N, 2) device, game int32 (B,) device or None.
        recs / game / count_dev may also be raw device pointers (ctypes.c_void_p) with `batch` given."""
        B = recs.shape[0] if batch is None else int(batch)
        if policy_out is None:
            policy_out = torch.empty((B, self.A), dtype=torch.float32, device=self.device)
        if value_out is None:
            value_out = torch.empty(B, dtype=torch.float32, device=self.device)
        call("bpp_net_forward", self._h, B, _ptr(count_dev) if count_dev is not None else C.c_void_p(0), _ptr(recs),
             _ptr(game) if game is not None else C.c_void_p(0), _ptr(items_wh), _ptr(policy_out), _ptr(value_out),
             _stream())
        return policy_out, value_out


PARAM_ORDER = tuple(
    ["conv_seqs.%d.%s.%s" % (s, sub, wb) for s in range(3)
     for sub in ("conv", "res_block0.conv0", "res_block0.conv1", "res_block1.conv0", "res_block1.conv1")
     for wb in ("weight", "bias")] +
    ["hidden_fc.weight", "hidden_fc.bias", "logits_fc.weight", "logits_fc.bias", "value_fc.weight", "value_fc.bias"])


class DeviceLearner:
    """fp32 learner step in hand-written CUDA (csrc/bpp_learner.cu, bpp_learner_* in include/bpp_b200.h): forward with
    stashed activations, the two losses of NNet.py:87-91, the full backward and Adam (NNet.py:31) on flat device
    buffers in state_dict order.  Deterministic (fixed-order gradient reduction, no atomics)."""

    def __init__(self, W, H, N, max_batch, device=None):
        _lib.load()
        device = _devidx(device)
        self.device = torch.device("cuda", device)
        self.W, self.H, self.N, self.A, self.max_batch = W, H, N, W * N, int(max_batch)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            call("bpp_learner_create", W, H, N, int(max_batch), int(device), C.byref(h))
        self._h = h
        n = C.c_int64()
        call("bpp_learner_num_params", self._h, C.byref(n))
        self.num_params = int(n.value)
        self.slices = {}
        for name in PARAM_ORDER:
            off, numel = C.c_int64(), C.c_int64()
            call("bpp_learner_param_offset", self._h, name.encode(), C.byref(off), C.byref(numel))
            self.slices[name] = (int(off.value), int(numel.value))
        self.params = torch.zeros(self.num_params, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros_like(self.params)
        self.losses = torch.zeros(2, dtype=torch.float32, device=self.device)
        self.reset_optimizer()

    def close(self):
        if getattr(self, "_h", None):
            call("bpp_learner_destroy", self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset_optimizer(self):
        """a fresh torch.optim.Adam (the reference re-creates the optimiser in every train() call, NNet.py:31)"""
        self.exp_avg = torch.zeros(self.num_params, dtype=torch.float32, device=self.device)
        self.exp_avg_sq = torch.zeros_like(self.exp_avg)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)

    def load_state_dict(self, state_dict):
        for name in PARAM_ORDER:
            off, numel = self.slices[name]
            t = state_dict[name].detach().to(self.device, torch.float32).reshape(-1)
            if t.numel() != numel:
                raise ValueError("parameter %s has %d elements, expected %d" % (name, t.numel(), numel))
            self.params[off:off + numel].copy_(t)

    def state_dict_into(self, module):
        """write the flat parameters back into a torch module with the reference's parameter names"""
        sd = module.state_dict()
        with torch.no_grad():
            for name in PARAM_ORDER:
                off, numel = self.slices[name]
                sd[name].copy_(self.params[off:off + numel].view_as(sd[name]))

    def grad_dict(self):
        return {name: self.grads[off:off + numel] for name, (off, numel) in self.slices.items()}

    def grad(self, recs, items_wh, pis, vs, ids=None, batch=None, train=True, logp_out=None, v_out=None):
        """losses (and gradients into self.grads when train) of one minibatch; example b = row ids[b] of the tables"""
        B = int(batch if batch is not None else (ids.shape[0] if ids is not None else recs.shape[0]))
        call("bpp_learner_grad", self._h, B, _ptr(self.params), _ptr(recs), _ptr(items_wh), _ptr(ids), _ptr(pis),
             _ptr(vs), _ptr(self.grads) if train else C.c_void_p(0), _ptr(self.losses), _ptr(logp_out), _ptr(v_out),
             _stream())
        return self.losses

    def adam(self, grad_scale=1.0, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        """one Adam update from self.grads; the step counter lives on the device (graph-capturable)"""
        self.step_dev += 1
        call("bpp_learner_adam", self.num_params, _ptr(self.params), _ptr(self.grads), _ptr(self.exp_avg),
             _ptr(self.exp_avg_sq), 0, _ptr(self.step_dev), float(grad_scale), float(lr), float(betas[0]),
             float(betas[1]), float(eps), _stream())


class NNetWrapper:
    """Same surface as NNet.py:17-111: predict / train / save_checkpoint / load_checkpoint."""

    # calibration bound of precision="auto": half of the stated tolerance |d pi| <= 2e-2, |d v| <= 2e-2
    AUTO_TOL = 1e-2

    def __init__(self, game, args, max_batch=8192, device=None, precision="auto"):
        """precision: "auto" (default) picks, at every sync_weights(), the fastest tensor-core mode that reproduces the
        fp32 forward of the CURRENT weights on a set of calibration states: plain "bf16" for well-conditioned networks
        (fresh initialisations), split-bf16 "bf16x3" otherwise (the reference's trained checkpoints have logits down to
        -4.8e3: one rounding to bf16 or fp16 anywhere in the network moves pi by up to 0.6, profiles/
        r02_precision_study.txt), "fp32" (CUDA cores) if even that fails.  An explicit mode is taken as given."""
        device = _devidx(device)
        self.args = args
        self.game = game
        self.nnet = BinPackingNNet(game, args)
        self.board_h, self.board_w = game.getBoardSize()
        self.action_size = game.getActionSize()
        self.num_items = args.num_items
        self.device = torch.device("cuda", device)
        self.nnet.to(self.device)  # the reference moves the module only if args.cuda; this path is CUDA-only
        self.dnet = DeviceNet(self.board_w, self.board_h, self.num_items, max_batch, device)
        self.precision_request = precision
        self.calibration = None
        if precision != "auto":
            self.dnet.set_precision(precision)
        self.sync_weights()

    def sync_weights(self):
        """push the torch parameters into the CUDA forward (after init / load_checkpoint / train)"""
        self.dnet.load_state_dict(self.nnet.state_dict())
        if self.precision_request == "auto":
            self._calibrate()

    def _calibration_states(self, n=64):
        """n reachable states (random legal play from generated instances), as compact records + item lists on the device"""
        cal = getattr(self, "_cal_states", None)
        if cal is None:
            from .engine import EnvOps
            from .game import ItemsGenerator
            W, H, N = self.board_w, self.board_h, self.num_items
            rs = np.random.RandomState(20201113)
            hts = rs.randint(max(1, (N + W - 1) // W), H + 1, size=n).astype(np.int32)
            items = torch.from_numpy(ItemsGenerator(W, H, N).items_batch(np.arange(n) + 31337, hts)).to(self.device)
            ops = EnvOps(W, H, N, self.device.index)
            recs = torch.zeros((n, 32), dtype=torch.int32, device=self.device)
            recs[:, 28] = (1 << N) - 1
            depth = torch.from_numpy(rs.randint(0, N, size=n)).to(self.device)
            u = torch.from_numpy(rs.random_sample((N, n))).to(self.device)
            for k in range(N - 1):
                valid = ops.valid_moves(recs, items).bool()
                cnt = valid.sum(dim=1)
                pick = (u[k] * cnt.clamp(min=1)).long().clamp(max=(cnt - 1).clamp(min=0))
                act = (valid.cumsum(dim=1) == (pick + 1)[:, None]).int().argmax(dim=1).to(torch.int32)
                nxt = ops.next_state(recs, items, act)
                go = (depth > k) & (cnt > 0)
                still = ops.valid_moves(nxt, items).any(dim=1)   # never calibrate on a terminal state
                recs = torch.where((go & still.bool())[:, None], nxt, recs)
            cal = self._cal_states = (recs.contiguous(), items.contiguous())
        return cal

    def _calibrate(self):
        recs, items = self._calibration_states()
        self.dnet.set_precision("fp32")
        p_ref, v_ref = self.dnet.forward(recs, items)
        chosen, report = "fp32", {}
        for mode in ("bf16", "bf16x3"):
            self.dnet.set_precision(mode)
            p, v = self.dnet.forward(recs, items)
            dp, dv = float((p - p_ref).abs().max()), float((v - v_ref).abs().max())
            report[mode] = (dp, dv)
            if dp <= self.AUTO_TOL and dv <= self.AUTO_TOL and np.isfinite(dp) and np.isfinite(dv):
                chosen = mode
                break
        self.dnet.set_precision(chosen)
        self.calibration = {"mode": chosen, "errors": report}
        return chosen

    # ---- inference ---------------------------------------------------------------------------------------------------
    def predict(self, board):
        """board: (N+1, H, W) array -> (pi (A,) float32, v (1,) float32), NNet.py:69-85."""
        recs, items = pack_states(np.asarray(board), self.board_w, self.board_h, self.num_items)
        recs_t = torch.from_numpy(recs.view(np.int32)).to(self.device)
        items_t = torch.from_numpy(items).to(self.device)
        pi, v = self.dnet.forward(recs_t, items_t)
        return pi[0].cpu().numpy(), v[:1].cpu().numpy()

    def predict_batch(self, recs, items_wh, game=None, count_dev=None):
        return self.dnet.forward(recs, items_wh, game, count_dev)

    # ---- learner (same semantics as NNet.py:27-67,87-91) --------------------------------------
    def _device_learner(self, batch):
        L = getattr(self, "_learner", None)
        if L is None or L.max_batch < batch:
            if L is not None:
                L.close()
            L = self._learner = DeviceLearner(self.board_w, self.board_h, self.num_items, batch, self.device.index)
        L.load_state_dict(self.nnet.state_dict())
        L.reset_optimizer()  # the reference builds a new Adam in every train() call (NNet.py:31)
        return L

    def train(self, examples, learner=None):
        """examples: list of (board (N+1, H, W), pi (A,), v) as CoachBPP builds them (NNet.py:27-67).  Minibatches are
        drawn with np.random.randint exactly like the reference; the step itself (forward, losses, backward, Adam) is the
        hand-written CUDA learner (learner="cuda", csrc/bpp_learner.cu) or torch autograd (learner="torch", the
        cross-check)."""
        if (learner or os.environ.get("BPP_LEARNER", "cuda")) == "torch":
            return self._train_torch(examples)
        from .distributed import world
        rank, ws = world()
        bs = int(self.args.batch_size)
        boards, pis, vs = list(zip(*examples))
        recs, items = pack_states(np.asarray(boards), self.board_w, self.board_h, self.num_items)
        recs_t = torch.from_numpy(recs.view(np.int32)).to(self.device)
        items_t = torch.from_numpy(items).to(self.device)
        pis_t = torch.as_tensor(np.asarray(pis, dtype=np.float32), device=self.device)
        vs_t = torch.as_tensor(np.asarray(vs).astype(np.float64).astype(np.float32).reshape(-1), device=self.device)
        batch_count = int(len(examples) / bs)
        if ws > 1:
            # every rank must issue the same number of gradient all-reduces: agree on the smallest per-rank step count
            # (ranks holding different numbers of examples would otherwise hang in NCCL)
            bc = torch.tensor([batch_count], dtype=torch.int64, device=self.device)
            torch.distributed.all_reduce(bc, op=torch.distributed.ReduceOp.MIN)
            batch_count = int(bc.item())
        with torch.cuda.device(self.device):  # the learner's entry points launch on the current device's stream
            L = self._device_learner(bs)
            for epoch in range(self.args.epochs):
                for _ in range(batch_count):
                    ids = torch.from_numpy(np.random.randint(len(examples), size=bs).astype(np.int64)).to(self.device)
                    L.grad(recs_t, items_t, pis_t, vs_t, ids=ids)
                    if ws > 1:
                        torch.distributed.all_reduce(L.grads)
                    L.adam(grad_scale=1.0 / ws)
            L.state_dict_into(self.nnet)
        self.nnet.eval()
        self.sync_weights()

    def _train_torch(self, examples):
        optimizer = torch.optim.Adam(self.nnet.parameters())  # re-created per call, default lr (args.lr ignored)
        for epoch in range(self.args.epochs):
            self.nnet.train()
            pi_losses, v_losses = AverageMeter(), AverageMeter()
            batch_count = int(len(examples) / self.args.batch_size)
            for _ in range(batch_count):
                ids = np.random.randint(len(examples), size=self.args.batch_size)
                boards, pis, vs = list(zip(*[examples[i] for i in ids]))
                boards = torch.as_tensor(np.array(boards), dtype=torch.float32, device=self.device)
                target_pis = torch.as_tensor(np.array(pis), dtype=torch.float32, device=self.device)
                target_vs = torch.as_tensor(np.array(vs).astype(np.float64), dtype=torch.float32, device=self.device)
                out_pi, out_v = self.nnet(boards)
                l_pi = self.loss_pi(target_pis, out_pi)
                l_v = self.loss_v(target_vs, out_v)
                total = l_pi + l_v
                pi_losses.update(l_pi.item(), boards.size(0))
                v_losses.update(l_v.item(), boards.size(0))
                optimizer.zero_grad()
                total.backward()
                allreduce_gradients(self.nnet)  # no-op unless torch.distributed is initialised with > 1 rank
                optimizer.step()
        self.nnet.eval()
        self.sync_weights()

    def train_compact(self, recs, items_wh, pis, vs, ops=None, steps_per_epoch=None, seed=None, use_graph=True,
                      learner=None):
        """Learner on compact examples that already live on the device (the batched / multi-GPU path).
        recs int32 (M, 32), items_wh int32 (M, N, 2), pis float32 (M, A), vs float32 (M,).  Same optimiser, losses and
        sampling-with-replacement as `train` (NNet.py:27-67); with torch.distributed initialised every rank draws its
        own minibatches and the gradients are averaged with one flat all-reduce per step.  learner="cuda": the
        hand-written step (bpp_learner_grad gathers the minibatch rows itself -> flat gradient -> NCCL all-reduce ->
        bpp_learner_adam), captured once in a CUDA graph and replayed.  learner="torch": the same loop on torch
        autograd (`ops` = EnvOps builds the dense input planes), kept as the cross-check.  Returns (mean pi loss, mean v
        loss) of the last epoch."""
        if (learner or os.environ.get("BPP_LEARNER", "cuda")) == "torch":
            return self._train_compact_torch(recs, items_wh, pis, vs, ops, steps_per_epoch, seed, use_graph)
        from .distributed import world
        rank, ws = world()
        M = recs.shape[0]
        dev = self.device
        bs = int(self.args.batch_size)
        if steps_per_epoch is None:
            steps_per_epoch = max(1, int(M / (bs * ws)))
        gen = torch.Generator(device=dev)
        gen.manual_seed((seed if seed is not None else int(np.random.randint(1 << 30))) * 977 + rank)
        recs = recs.contiguous()
        items_wh = items_wh.contiguous()
        pis = pis.contiguous().float()
        vs = vs.contiguous().float()
        L = self._device_learner(bs)
        ids = torch.zeros(bs, dtype=torch.int64, device=dev)
        acc = torch.zeros(2, dtype=torch.float64, device=dev)

        def step():
            L.grad(recs, items_wh, pis, vs, ids=ids)
            if ws > 1:
                torch.distributed.all_reduce(L.grads)
            L.adam(grad_scale=1.0 / ws)
            acc.add_(L.losses)

        total_steps = steps_per_epoch * int(self.args.epochs)
        last_epoch_start = total_steps - steps_per_epoch
        state = {"done": 0, "n_acc": 0}

        def next_batch():
            if state["done"] == last_epoch_start:  # the returned losses are the last epoch's means
                acc.zero_()
                state["n_acc"] = 0
            ids.copy_(torch.randint(0, M, (bs,), device=dev, generator=gen))
            state["done"] += 1
            state["n_acc"] += 1

        graph = None
        if use_graph and total_steps > 8:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up on a side stream (real training steps)
                    next_batch()
                    step()
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            next_batch()
            with torch.cuda.graph(graph):
                step()
            graph.replay()  # capture records the step without running it
        while state["done"] < total_steps:
            next_batch()
            if graph is not None:
                graph.replay()
            else:
                step()
        n_acc = state["n_acc"]
        L.state_dict_into(self.nnet)
        self.nnet.eval()
        self.sync_weights()
        mean = (acc / max(1, n_acc)).cpu().numpy()
        return float(mean[0]), float(mean[1])

    def _train_compact_torch(self, recs, items_wh, pis, vs, ops, steps_per_epoch=None, seed=None, use_graph=True):
        from .distributed import world
        rank, ws = world()
        M = recs.shape[0]
        dev = self.device
        bs = int(self.args.batch_size)
        if steps_per_epoch is None:
            steps_per_epoch = max(1, int(M / (bs * ws)))
        gen = torch.Generator(device=dev)
        gen.manual_seed((seed if seed is not None else int(np.random.randint(1 << 30))) * 977 + rank)
        optimizer = torch.optim.Adam(self.nnet.parameters(), capturable=use_graph)
        self.nnet.train()
        # static minibatch buffers (graph inputs) and loss outputs
        b_recs = torch.zeros((bs, 32), dtype=torch.int32, device=dev)
        b_items = torch.zeros((bs,) + tuple(items_wh.shape[1:]), dtype=torch.int32, device=dev)
        b_pis = torch.zeros((bs, pis.shape[1]), dtype=torch.float32, device=dev)
        b_vs = torch.zeros(bs, dtype=torch.float32, device=dev)
        losses = torch.zeros(2, dtype=torch.float32, device=dev)

        def load_batch():
            ids = torch.randint(0, M, (bs,), device=dev, generator=gen)
            b_recs.copy_(recs[ids])
            b_items.copy_(items_wh[ids])
            b_pis.copy_(pis[ids])
            b_vs.copy_(vs[ids])

        def step():
            boards = ops.planes(b_recs, b_items)
            out_pi, out_v = self.nnet(boards)
            l_pi = self.loss_pi(b_pis, out_pi)
            l_v = self.loss_v(b_vs, out_v)
            (l_pi + l_v).backward()
            allreduce_gradients(self.nnet)
            optimizer.step()
            losses[0] = l_pi.detach()
            losses[1] = l_v.detach()

        total_steps = steps_per_epoch * int(self.args.epochs)
        graph = None
        done = 0
        if use_graph and total_steps > 8:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):  # warm-up steps on a side stream (they are real training steps)
                    load_batch()
                    optimizer.zero_grad(set_to_none=True)
                    step()
                    done += 1
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            load_batch()
            optimizer.zero_grad(set_to_none=True)
            with torch.cuda.graph(graph):
                step()
            graph.replay()  # capture records the step without running it
            done += 1
        acc = torch.zeros(2, dtype=torch.float64, device=dev)
        n_acc = 0
        last_epoch_start = total_steps - steps_per_epoch
        while done < total_steps:
            load_batch()
            if graph is not None:
                graph.replay()
            else:
                optimizer.zero_grad(set_to_none=True)
                step()
            if done >= last_epoch_start:
                acc += losses.double()
                n_acc += 1
            done += 1
        self.nnet.eval()
        self.sync_weights()
        mean = (acc / max(1, n_acc)).cpu().numpy()
        return float(mean[0]), float(mean[1])

    def loss_pi(self, targets, outputs):  # NNet.py:87-88
        return -torch.sum(targets * outputs) / targets.size()[0]

    def loss_v(self, targets, outputs):  # NNet.py:90-91
        return torch.sum((targets - outputs.view(-1)) ** 2) / targets.size()[0]

    # ---- checkpoints (NNet.py:93-111; same {'state_dict': ...} format) ---------------------------------------------------
    def save_checkpoint(self, folder='checkpoint', filename='checkpoint.pth.tar'):
        filepath = os.path.join(folder, filename)
        if not os.path.exists(folder):
            os.mkdir(folder)
        torch.save({'state_dict': self.nnet.state_dict()}, filepath)

    def load_checkpoint(self, folder='checkpoint', filename='checkpoint.pth.tar'):
        filepath = os.path.join(folder, filename)
        if not os.path.exists(filepath):
            raise FileNotFoundError("No model in path {}".format(filepath))
        checkpoint = torch.load(filepath, map_location=self.device)
        self.nnet.load_state_dict(checkpoint['state_dict'])
        self.sync_weights()
