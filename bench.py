#!/usr/bin/env python
"""bench.py — MCTS simulations/s of the self-play hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--games G] [--sims S] [--workload ...]

HEADLINE (BASELINE.json configs[1]): the main_bpp.py default instance — 15x15 virtual bin,
10 items from ItemsGenerator(15, h, 10) with h ~ randint(2, 16) per 20 consecutive seeds, total_area = 15*h,
numMCTSSims = 200, cpuct = 1, alpha = 0.75, empty rewards list — as G = 4096 games per GPU with the stub
uniform-prior evaluator (p = 1/A, v = 0).  One STEP = one batch of G complete self-play episodes (reset, then per
move: 200 simulations per game, visit counts out, action ~ counts (CoachBPP.py:86-87), getNextState + getGameEnded).
Seeds: generator_seed = 1000 + global episode index.  Scaling is weak: every rank plays its own G games.

value   = simulations/s with the instances already resident in HBM (whole job, all ranks, max-over-ranks time)
e2e     = the same through the host-buffer C-ABI call bpp_engine_play_stub_host: instances uploaded from pinned host
          memory and visit counts / actions / rewards downloaded inside the timed region
roofline= dominant kernel k_episode<U> (all simulations of a step in one launch): algorithmic bytes (SURVEY.md §8(d)
          formula with the kernel-counted edges and expansions per simulation) / CUDA-event time of that kernel, against
          MEASURED_PEAKS.json hbm_gbs; next to it what the layout really moves (layout_bytes_per_sim, from kernel
          counters) and, when profiles/r02_traffic.json holds an ncu capture of this configuration, the measured DRAM
          traffic per launch and the DRAM fraction it amounts to
cpu_baseline = oracle/bpp_oracle.py (a port that keeps the reference's data structures) on ONE host core, bounded
          sample of the same workload.  `--impl reference` runs the reference itself (when /root/reference/xw_mcts or
          baseline/_ref is present) or that port on all host cores instead of the GPU.

SECONDARY records in the same JSON line (`secondary.*`; each with value, e2e, roofline, clocks and, at N = 1,
cpu_baseline):
  real15     the default instance with the REAL policy/value net as leaf evaluator (NNet.predict, NNet.py:69-85), weights
             = the reference's shipped trained checkpoint (tests/golden/net.npz ck_w.*), in the precision mode that meets
             the stated tolerance (precision="auto") and, for comparison, in plain bf16
  real20     BASELINE.json configs[2]: 20x20 bin, numMCTSSims = 200, net randomly initialised (seed 0)
  iteration  configs[3]: one full CoachBPP iteration (batched self-play, example all-gather, data-parallel learner with
             one flat gradient all-reduce per step)
  arena      configs[4] share: greedy evaluation sweep, two nets per seed
  shard_checksum  sha256 of the visit-count matrices of a fixed 256-seed block sharded over the ranks (equal at any N)
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, N = 15, 15, 10
SIMS = 200
CPUCT = 1.0
ALPHA = 0.75
METRIC = "mcts_simulations_per_sec"
UNIT = "sims/s"
NET_FLOPS = {(15, 15, 10): 4394592, (20, 20, 10): 7249920}


def workload(first_episode, count, Wb=W, Hb=H):
    """(seeds, generator heights, total areas) of `count` consecutive episodes starting at `first_episode`."""
    idx = np.arange(first_episode, first_episode + count)
    seeds = 1000 + idx
    # height per batch of 20 episodes (CoachBPP.py:117-119 draws it once per iteration of numEps = 20 episodes)
    batch = idx // 20
    ub = np.unique(batch)
    hb = {int(b): np.random.RandomState(77000 + int(b)).randint(2, Hb + 1) for b in ub}
    heights = np.array([hb[int(b)] for b in batch], dtype=np.int32)
    return seeds, heights, (Wb * heights).astype(np.int32)


def config_dict(args, extra=None, mult=1):
    c = {"workload": "configs[1]: main_bpp default instance (15x15 bin, 10 items, numMCTSSims=200, cpuct=1), "
                     f"{args.games} lockstep games per GPU, stub uniform-prior net, whole self-play episodes" +
                     (f"; a step = {mult} x {args.games} episodes streamed through the resident games (a game whose "
                      "episode ends takes the next instance inside the kernel)" if mult > 1 else ""),
         "games_per_gpu": args.games, "num_mcts_sims": args.sims, "bin": [W, H], "items": N,
         "action_choice": "sample ~ visit counts (greedy=False)",
         "cache": "working set (search graphs of all games, several GB per step) is far larger than the 126 MB L2; "
                  "no explicit flush"}
    if extra:
        c.update(extra)
    return c


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def load_traffic():
    """ncu --set full captures of this round (scripts/capture_traffic.py writes the file): DRAM bytes per launch"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arms (the only place bench.py executes oracle/ or the reference)
def _reference_dir():
    for d in ("/root/reference/xw_mcts", os.path.join(ROOT, "baseline", "_ref", "xw_mcts"),
              os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isfile(os.path.join(d, "MCTS_bpp.py")) and os.path.isdir(os.path.join(d, "binpacking")):
            return d
    return None


_IMPL = {}


def _cpu_impl():
    """(kind, Game, Generator, MCTS, dotdict): the unmodified reference when it is present on this machine (one shim:
    ndarray.tostring() was removed in numpy 2.3), else the pinned port oracle/bpp_oracle.py"""
    if _IMPL:
        return _IMPL["v"]
    from oracle import bpp_oracle as O
    ref = None if os.environ.get("BPP_BENCH_FORCE_PORT") else _reference_dir()
    if ref:
        try:
            os.environ.setdefault("WANDB_MODE", "disabled")
            sys.path.insert(0, ref)
            from binpacking.BinPackingGame import BinPackingGame as RefGame, ItemsGenerator as RefGen
            from MCTS_bpp import MCTS as RefMCTS

            class Game(RefGame):
                def stringRepresentation(self, board):  # numpy >= 2.3 shim, byte-identical key
                    return b"".join(p.tobytes() for p in board)
            _IMPL["v"] = ("reference", Game, RefGen, RefMCTS, O.dotdict, ref)
            return _IMPL["v"]
        except Exception:
            pass
    _IMPL["v"] = ("port", O.OracleGame, O.OracleItemsGenerator, O.OracleMCTS, O.dotdict, "oracle/bpp_oracle.py")
    return _IMPL["v"]


def _cpu_episode(ep_index, sims, Wb=W, Hb=H, net=None):
    """one whole self-play episode of the workload on the CPU; returns the simulations it ran"""
    from oracle import bpp_oracle as O
    kind, Game, Gen, MCTS, dd, _ = _cpu_impl()
    seeds, heights, areas = workload(ep_index, 1, Wb, Hb)
    items = Gen(Wb, int(heights[0]), N).items_generator(int(seeds[0]))
    g = Game(Wb, Hb, N, 1)
    m = MCTS(g, net if net is not None else O.StubNet("U", g.getActionSize()),
             dd(numMCTSSims=sims, cpuct=CPUCT, alpha=ALPHA))
    rng = np.random.RandomState(ep_index)
    board, planes = g.getInitBoard(), g.getInitItems(items)
    moves = 0
    while True:
        state = g.getBinItem(board, planes)
        pi = m.getActionProb(state, int(areas[0]), [])
        a = int(rng.choice(len(pi), p=pi))
        board, planes = g.getNextState(board, a, planes)
        moves += 1
        r, _ = g.getGameEnded(g.getBinItem(board, planes), int(areas[0]), [], ALPHA)
        if r != 0:
            break
    return moves * sims


def _cpu_worker(q_in, q_out, sims):
    os.environ["OMP_NUM_THREADS"] = "1"
    while True:
        ep = q_in.get()
        if ep is None:
            return
        q_out.put(_cpu_episode(ep, sims))


def cpu_single_core(budget_s, sims, Wb=W, Hb=H, net=None):
    """time the CPU implementation on this process for about budget_s seconds of whole episodes"""
    _cpu_episode(0, sims, Wb, Hb, net)  # warm-up (imports, allocator)
    t0 = time.perf_counter()
    n_sims = eps = 0
    while time.perf_counter() - t0 < budget_s:
        n_sims += _cpu_episode(1 + eps, sims, Wb, Hb, net)
        eps += 1
    dt = time.perf_counter() - t0
    return n_sims / dt, eps, dt


def run_reference_arm(args):
    """The reference's CPU implementation of the path on all host cores: one episode stream per core, summed.  Uses the
    reference's own code when it is present (kind "reference"), else the port (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    kind, _, _, _, _, src = _cpu_impl()
    ctx = mp.get_context("fork")
    q_in, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(q_in, q_out, args.sims), daemon=True) for _ in range(cores)]
    for p in procs:
        p.start()
    eps_per_step = cores  # one episode per core per step

    def step(k):
        for i in range(eps_per_step):
            q_in.put(k * eps_per_step + i)
        return sum(q_out.get() for _ in range(eps_per_step))

    for k in range(args.warmup):
        step(k)
    t0 = time.perf_counter()
    total = 0
    for k in range(args.steps):
        total += step(args.warmup + k)
    dt = time.perf_counter() - t0
    for _ in procs:
        q_in.put(None)
    val = total / dt
    sample = f"{eps_per_step} episodes per step (one per core), {args.steps} steps, {kind}: {src}"
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # the same workload description as the B200 arm's line (each CPU step is a bounded sample of it: cpu_baseline.sample)
            "config": config_dict(args, {"parallelism": f"games sharded over {args.gpus} GPU(s), no data-path collective",
                                         "episodes_per_step": max(1, args.stub_stream_mult) * args.games},
                                  max(1, args.stub_stream_mult)),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "episodes_per_sec": args.steps * eps_per_step / dt, "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md clocks line)."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            import datetime
            for line in self.proc.stdout:
                r = [c.strip() for c in line.split(",")]
                try:  # nvidia-smi's own sample time (its stdout may reach us in bursts)
                    t = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except Exception:
                    t = time.time()
                self.rows.append((t, r))
        except Exception:
            pass

    def window(self, t0, t1):
        """median SM clock and throttle reasons of the samples taken in [t0, t1] (time.time() values)"""
        time.sleep(0.05)  # let the last samples of the window arrive
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def collect(lo, hi):
            sm, mx, reasons = [], 0, set()
            for t, r in list(self.rows):
                try:
                    mx = max(mx, float(r[2]))
                    if t < lo or t > hi:
                        continue
                    sm.append(float(r[1]))
                    for nm, v in zip(names, r[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
                except Exception:
                    continue
            return sm, mx, reasons
        sm, mx, reasons = collect(t0, t1)
        if not sm:  # region shorter than the sampling period: take the neighbouring samples
            sm, mx, reasons = collect(t0 - 0.1, t1 + 0.1)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)


class Ctx:
    """per-process benchmark context: device, ranks, reductions, the clock sampler"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            import datetime
            # a rank that diverges must fail the run in minutes, not after NCCL's default 10-minute watchdog
            dist.init_process_group("nccl", device_id=self.dev, timeout=datetime.timedelta(seconds=240))
        self.sampler = ClockSampler(self.local)
        self.sampler.start()
        self.peaks = load_peaks()
        self.traffic = load_traffic()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, maxes, sums):
        """MAX over ranks of `maxes`, SUM over ranks of `sums` (lists of floats)"""
        t = self.torch
        a = t.tensor(list(maxes), dtype=t.float64, device=self.dev)
        b = t.tensor(list(sums), dtype=t.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(a, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(b, op=self.dist.ReduceOp.SUM)
        return a.tolist(), b.tolist()

    def timed(self, fn, steps):
        """barrier + sync, `steps` calls of fn(k) between two CUDA events, barrier + sync; returns (device ms, wall
        ms, clocks during the region, list of fn results)"""
        t = self.torch
        self.barrier()
        w0, c0 = time.perf_counter(), time.time()
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        e0.record()
        res = [fn(k) for k in range(steps)]
        e1.record()
        self.barrier()
        w1, c1 = time.perf_counter(), time.time()
        return e0.elapsed_time(e1), 1e3 * (w1 - w0), self.sampler.window(c0, c1), res

    def close(self):
        self.sampler.stop()
        if self.world > 1:
            self.dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
def run_stub_arm(cx, mult=None, steps=None, with_e2e=True):
    """headline: whole episodes with the in-kernel stub evaluator (one k_episode launch per step).  mult > 1: a step is
    mult x G episodes streamed through the G resident games (bpp_engine_play_stub_stream), mult == 1: one episode per
    game (bpp_engine_reset + bpp_engine_play_stub)."""
    torch, args = cx.torch, cx.args
    mult = mult or max(1, args.stub_stream_mult)
    if steps is not None:
        args = argparse.Namespace(**{**vars(args), "steps": steps})
    from resource_packing_self_play_b200 import _lib
    from resource_packing_self_play_b200.engine import SearchEngine, algorithmic_bytes_per_sim
    from resource_packing_self_play_b200.game import ItemsGenerator
    dev, rank, world, local = cx.dev, cx.rank, cx.world, cx.local
    G = args.games
    E = mult * G   # episodes per step
    n_steps_total = args.warmup + args.steps
    gen = ItemsGenerator(W, H, N)
    # synthetic instances for every step of this rank (host generation is setup, not part of the timed path)
    inst = []
    for k in range(2 * n_steps_total):  # first half: device-resident arm, second half: e2e arm
        first = ((k * world) + rank) * E
        seeds, heights, areas = workload(first, E)
        # the package's device-side generator (bit-identical to numpy's legacy RNG path, tests/test_gpu_env.py)
        inst.append((gen.items_batch_device(seeds, heights, device=local).cpu().numpy(), areas))
    edge_cap = 0
    worst_units_per_node = 3 * ((W * N + 3) // 4 * 4) + ((W * N + 3) // 4 * 4) // 4
    if args.edge_frac < 1.0:  # profiling runs: a smaller edge pool keeps ncu's save/restore cheap
        edge_cap = int((args.sims * N + N + 2) * worst_units_per_node * args.edge_frac)
    eng = SearchEngine(W, H, N, G, args.sims, CPUCT, device=local, edge_cap=edge_cap)
    nan_bl = np.full(E, np.nan)
    bl_dev = torch.from_numpy(nan_bl).to(dev)
    counts_buf = torch.zeros((N, E, W * N), dtype=torch.int32, device=dev)
    actions_buf = torch.empty((N, E), dtype=torch.int32, device=dev)
    r_buf = torch.empty(E, dtype=torch.int32, device=dev)
    score_buf = torch.empty(E, dtype=torch.float64, device=dev)
    moves_buf = torch.zeros(E, dtype=torch.int32, device=dev)
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

    # ---- arm 1: inputs resident in HBM ------------------------------------------------------------------------------
    dev_inst = [(torch.from_numpy(i).to(dev), torch.from_numpy(a).to(dev)) for i, a in inst[:n_steps_total]]
    search_events = []

    def eng_stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step_resident(k, timed):
        """reset + ONE launch of the whole-episode kernel (search -> counts -> choose -> play, N moves)"""
        items, area = dev_inst[k]
        if mult == 1:
            eng.reset(items, area, bl_dev)
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if mult == 1:
            _lib.call("bpp_engine_play_stub", eng._h, _lib.STUB["U"], _lib.CHOOSE_SAMPLE, C.c_uint64(1234 + k), 0,
                      vp(counts_buf), vp(actions_buf), None, eng_stream())
        else:   # the stream's own set-up (queue counter, k_stream_begin, output clears) is inside the events
            _lib.call("bpp_engine_play_stub_stream", eng._h, _lib.STUB["U"], _lib.CHOOSE_SAMPLE, C.c_uint64(1234 + k), E,
                      vp(items), vp(area), vp(bl_dev), None, vp(counts_buf), vp(actions_buf), vp(r_buf), vp(score_buf),
                      vp(moves_buf), eng_stream())
        if timed:
            e1.record()
            search_events.append((e0, e1))

    for k in range(args.warmup):
        step_resident(k, False)
    eng.check()
    eng.stats(reset=True)
    ms, _, clocks, _ = cx.timed(lambda k: step_resident(args.warmup + k, True), args.steps)
    eng.check()
    st = eng.stats(reset=True)
    done = eng.status()["done"]
    assert bool((done == 1).all()), "some games did not finish their episode"
    if mult > 1:
        assert int(moves_buf.min()) >= 1, "some episodes of the stream were not played"
    g_nodes, g_units = eng.graph_sizes()
    graph = {"max_nodes_per_game": int(g_nodes.max()), "max_edge_units_per_game": int(g_units.max()),
             "mean_edge_units_per_game": float(g_units.double().mean()), "engine_device_bytes": eng.device_bytes,
             "edge_units_capacity_per_game": eng.edge_cap, "edge_units_worst_case_per_game": eng.edge_cap_worst}
    search_ms = sum(a.elapsed_time(b) for a, b in search_events)
    n_search = len(search_events)

    # ---- arm 2: end to end through the host-buffer C ABI --------------------------------------------------------------
    pinned_items = [torch.from_numpy(i).pin_memory() for i, _ in inst[n_steps_total:]]
    pinned_area = [torch.from_numpy(a).pin_memory() for _, a in inst[n_steps_total:]]
    out = {"counts": torch.empty((N, E, W * N), dtype=torch.int32).pin_memory().numpy(),
           "actions": torch.empty((N, E), dtype=torch.int32).pin_memory().numpy(),
           "r": torch.empty(E, dtype=torch.int32).pin_memory().numpy(),
           "score": torch.empty(E, dtype=torch.float64).pin_memory().numpy(),
           "moves": torch.empty(E, dtype=torch.int32).pin_memory().numpy()}
    pinned_bl = torch.from_numpy(nan_bl).pin_memory().numpy()

    def step_e2e(k):
        if mult == 1:
            eng.play_stub_host("U", pinned_items[k].numpy(), pinned_area[k].numpy(), pinned_bl,
                               choose_mode=_lib.CHOOSE_SAMPLE, seed=99 + k, out=out)
        else:
            eng.play_stub_stream_host("U", pinned_items[k].numpy(), pinned_area[k].numpy(), pinned_bl,
                                      choose_mode=_lib.CHOOSE_SAMPLE, seed=99 + k, out=out)
        return int(out["moves"].sum())

    e2e_ms, st2 = 0.0, {"sims": 0, "launches": 0}
    if with_e2e:
        for k in range(args.warmup):
            step_e2e(k)
        eng.stats(reset=True)
        e2e_dev_ms, e2e_wall_ms, _, _ = cx.timed(lambda k: step_e2e(args.warmup + k), args.steps)
        e2e_ms = max(e2e_dev_ms, e2e_wall_ms)  # conservative: device events vs wall clock
        st2 = eng.stats(reset=True)
    h2d = inst[0][0].nbytes + inst[0][1].nbytes + nan_bl.nbytes
    d2h = sum(v.nbytes for v in out.values())
    eng.close()

    (ms, e2e_ms, search_ms_max), (sims, sims2, launches, episodes) = cx.reduce(
        [ms, e2e_ms, search_ms], [st["sims"], st2["sims"], st["launches"] + st2["launches"], args.steps * E])
    if rank != 0:
        return None
    peak = float(cx.peaks.get("hbm_gbs", 6650.0))
    d_bar, e_bar = st["edges"] / st["sims"], st["expansions"] / st["sims"]
    bytes_per_sim = algorithmic_bytes_per_sim(W, H, N, d_bar, e_bar)
    # what THIS layout moves per simulation (kernel counters): per selection one 128-byte node record and the node's edge
    # block (valid actions only); per hash probe a table word and a candidate record; per backup level Q, {Nsa, child}
    # and Ns read + written; per expansion the new block (mean size) and the record update
    mean_block = 8.0 * float(g_units.double().sum()) / max(1.0, float(g_nodes.double().sum()))
    layout = (d_bar * 128 + 8.0 * st["edge_units_read"] / st["sims"] + (st["probes"] / st["sims"]) * 132 + d_bar * 40 +
              e_bar * (mean_block + 128))
    sims_per_launch = st["sims"] / n_search
    kernel_s = search_ms / n_search * 1e-3
    achieved = bytes_per_sim * sims_per_launch / kernel_s / 1e9
    cap = cx.traffic.get("k_episode", {})
    traffic = cap.get("dram_bytes_per_launch") if (cap.get("games") == G and cap.get("sims") == args.sims) else None
    if traffic and mult > 1:
        traffic *= mult   # the capture is of one episode per game; a streamed launch plays mult of them per game
    api = "bpp_engine_play_stub_host" if mult == 1 else "bpp_engine_play_stub_stream_host"
    line = {
        "metric": METRIC, "value": sims / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, {"parallelism": f"games sharded over {world} GPU(s), no data-path collective",
                                     "episodes_per_step": E}, mult),
        "episodes_per_sec": episodes / (ms * 1e-3),
        "e2e": {"value": sims2 / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "episodes_per_sec": episodes / (e2e_ms * 1e-3),
                "api": api + " (pinned host buffers)"} if with_e2e else None,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_episode<STUB_U,15> (whole episodes, one launch per step)",
                     "traffic_source": "ncu capture of one episode per game (profiles/r02_traffic.json)" +
                                       (f" x {mult} episodes per game" if mult > 1 else ""),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": "MEASURED_PEAKS.json (measured)" if cx.peaks else "fallback 6650 GB/s",
                     # DRAM read+write bytes of one k_episode launch: ncu --set full capture of THIS round and THIS
                     # configuration (scripts/capture_traffic.py -> profiles/r02_traffic.json), else null
                     "traffic": traffic,
                     "dram_frac": (traffic / kernel_s / 1e9 / peak) if traffic else None,
                     "algorithmic_bytes_per_sim": bytes_per_sim,
                     "layout_bytes_per_sim": layout,
                     "layout_achieved_gbs": layout * sims_per_launch / kernel_s / 1e9,
                     "edges_per_sim": d_bar, "expansions_per_sim": e_bar, "sims_per_launch": sims_per_launch,
                     "kernel_ms_per_launch": search_ms / n_search,
                     "kernel_share_of_step": search_ms / (ms if world == 1 else search_ms_max or ms),
                     "limiter": "instruction issue + dependent-load latency (ncu: profiles/), not DRAM bandwidth: "
                                "`achieved` charges SURVEY §8(d)'s dense A-wide rows, the layout stores valid actions only",
                     },
        "clocks": clocks, "graph": graph,
    }
    if world == 1 and not args.no_cpu and with_e2e:
        v, eps, dt = cpu_single_core(args.cpu_seconds, args.sims)
        kind, _, _, _, _, src = _cpu_impl()
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                                "sample": f"{eps} whole episodes of the same workload in {dt:.1f} s, {src}"}
    return line


# ---------------------------------------------------------------------------------------------------------------------
class _Gm:
    def __init__(self, Wb, Hb):
        self.bin_width, self.bin_height, self.num_items = Wb, Hb, N

    def getBoardSize(self):
        return (self.bin_height, self.bin_width)

    def getActionSize(self):
        return self.bin_width * N


def _margs(sims, **kw):
    from resource_packing_self_play_b200.utils import dotdict
    d = dict(numMCTSSims=sims, cpuct=CPUCT, alpha=ALPHA, num_items=N, num_bins=1, cuda=True)
    d.update(kw)
    return dotdict(d)


def _make_net(cx, Wb, Hb, G, precision):
    """real15: the reference's shipped trained checkpoint; real20: random init under torch.manual_seed(0)"""
    torch = cx.torch
    from resource_packing_self_play_b200.nnet import NNetWrapper
    torch.manual_seed(0)
    net = NNetWrapper(_Gm(Wb, Hb), _margs(cx.args.sims), max_batch=G, device=cx.local, precision=precision)
    weights = "random init (torch.manual_seed(0))"
    if (Wb, Hb) == (15, 15):
        d = np.load(os.path.join(ROOT, "tests", "golden", "net.npz"))
        net.nnet.load_state_dict({k[5:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("ck_w.")})
        net.sync_weights()
        weights = "reference's shipped trained checkpoint (tests/golden/net.npz ck_w.*)"
    return net, weights


def run_real_arm(cx, Wb, Hb, precision="auto", with_e2e=True, with_cpu=True, steps=None, warmup=None, games=None):
    """whole self-play episodes with the real net as leaf evaluator (bpp_engine_play_net: asynchronous per game)"""
    torch, args = cx.torch, cx.args
    from resource_packing_self_play_b200 import _lib
    from resource_packing_self_play_b200.game import ItemsGenerator
    from resource_packing_self_play_b200.mcts import BatchedMCTS
    from resource_packing_self_play_b200.nnet import BinPackingNNet
    G = games or args.games
    steps = steps or max(3, args.steps // 5)
    warmup = warmup if warmup is not None else max(3, args.warmup)
    net, weights = _make_net(cx, Wb, Hb, G, precision)
    mode = net.dnet.precision
    net_kernels = 2   # evaluator kernels per lockstep step: trunk (k_net_gr: the four levels in one launch) + heads
    bm = BatchedMCTS(_Gm(Wb, Hb), net, _margs(args.sims), G, device=cx.local)
    eng = bm.eng
    gen = ItemsGenerator(Wb, Hb, N)
    n_tot = warmup + steps

    # episodes per step, streamed through the G resident games.  A game that needs the evaluator on every simulation (flat
    # prior) takes numMCTSSims x moves sequential lockstep steps whatever the batch does, so the stream must be long against
    # that tail: 16 x G with the sharp trained policy (whose bulk of episodes is over in ~300 steps), 8 x G with the
    # random-init net (all episodes alike; 4 x G: fill 0.76, 8 x G: 0.81, 16 x G: 0.83)
    mult = args.stream_mult or (16 if (Wb, Hb) == (15, 15) else 8)
    E = mult * G

    def instances(k):
        seeds, hts, areas = workload(((k * cx.world) + cx.rank) * E, E, Wb, Hb)
        return gen.items_batch_device(seeds, hts, device=cx.local).cpu().numpy(), areas
    inst = [instances(k) for k in range(2 * n_tot)]
    dev_inst = [(torch.from_numpy(i).to(cx.dev), torch.from_numpy(a).to(cx.dev)) for i, a in inst[:n_tot]]
    roots = torch.empty((N, E, 32), dtype=torch.int32, device=cx.dev)
    counts = torch.empty((N, E, Wb * N), dtype=torch.int32, device=cx.dev)
    actions = torch.empty((N, E), dtype=torch.int32, device=cx.dev)
    r_out = torch.empty(E, dtype=torch.int32, device=cx.dev)
    score_out = torch.empty(E, dtype=torch.float64, device=cx.dev)
    moves_out = torch.zeros(E, dtype=torch.int32, device=cx.dev)
    nan_bl = torch.full((E,), float("nan"), dtype=torch.float64, device=cx.dev)
    nsteps = C.c_int32(0)
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

    def step(k):
        items, area = dev_inst[k]
        moves_out.zero_()
        _lib.call("bpp_engine_play_net_stream", eng._h, net.dnet._h, _lib.CHOOSE_SAMPLE, C.c_uint64(7 + k), E, vp(items),
                  vp(area), vp(nan_bl), C.c_void_p(0), vp(counts), vp(actions), vp(roots), vp(r_out), vp(score_out),
                  vp(moves_out), C.byref(nsteps), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        return int(nsteps.value)

    for k in range(warmup):
        step(k)
    eng.check()
    # accounting pass: the same steps with CUDA events around every evaluator call / expand+select call
    eng.set_profile(True)
    eng.stats(reset=True)
    for k in range(steps):
        step(warmup + k)
    prof = eng.profile()
    prof_st = eng.stats(reset=True)
    eng.set_profile(False)
    ms, _, clocks, lock_steps = cx.timed(lambda k: step(warmup + k), steps)
    eng.check()
    st = eng.stats(reset=True)
    assert bool((moves_out > 0).all()), "some episodes did not finish"
    e2e = None
    if with_e2e:
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
        h_inst = [(pin(i), pin(a)) for i, a in inst[n_tot:]]
        out = {"roots": pin(np.empty((N, E, 32), dtype=np.uint32)), "counts": pin(np.empty((N, E, Wb * N), dtype=np.int32)),
               "actions": pin(np.empty((N, E), dtype=np.int32)), "r": pin(np.empty(E, dtype=np.int32)),
               "score": pin(np.empty(E, dtype=np.float64)), "moves": pin(np.empty(E, dtype=np.int32))}
        h_bl = pin(np.full(E, np.nan))

        def step_e2e(k):
            eng.play_net_stream_host(net.dnet, h_inst[k][0], h_inst[k][1], h_bl, choose_mode=_lib.CHOOSE_SAMPLE,
                                     seed=99 + k, out=out)
            return int(out["moves"].sum())
        for k in range(min(warmup, 2)):
            step_e2e(k)
        eng.stats(reset=True)
        d_ms, w_ms, _, _ = cx.timed(lambda k: step_e2e(warmup + k), steps)
        st2 = eng.stats(reset=True)
        e2e = {"ms": max(d_ms, w_ms), "sims": st2["sims"],
               "h2d": int(inst[0][0].nbytes + inst[0][1].nbytes + 8 * E),
               "d2h": int(sum(v.nbytes for k2, v in out.items() if k2 != "steps"))}
    (ms, e2e_ms), (sims, sims2, exps, launches, episodes) = cx.reduce(
        [ms, e2e["ms"] if e2e else 0.0],
        [st["sims"], e2e["sims"] if e2e else 0, st["expansions"], st["launches"] + net_kernels * sum(lock_steps), steps * E])
    rec = None
    if cx.rank == 0:
        peak = float(cx.peaks.get("bf16_tflops_sustained", 1400.0))
        flops = NET_FLOPS[(Wb, Hb, N)]
        achieved = prof_st["expansions"] * flops / (prof["evaluator_ms"] * 1e-3) / 1e12 if prof["evaluator_ms"] else None
        cap = cx.traffic.get("net_%dx%d_%s" % (Wb, Hb, mode), {})
        rec = {"value": sims / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
               "episodes_per_sec": episodes / (ms * 1e-3), "leaf_evals_per_sec": exps / (ms * 1e-3),
               "dtype": f"{mode} (net) / f64 (tree)", "precision_mode": mode,
               "precision_requested": precision, "calibration": net.calibration,
               "config": {"workload": f"{Wb}x{Hb} bin, 10 items, numMCTSSims={args.sims}, real policy/value net as leaf "
                                      f"evaluator, {G} resident games per GPU, {E} whole self-play episodes per step "
                                      "streamed through them (a game whose episode ends takes the next instance)",
                          "weights": weights, "games_per_gpu": G, "episodes_per_step": E},
               "lockstep_steps_per_batch": float(np.mean(lock_steps)),
               "leaf_batch_fill": exps / max(1.0, cx.world * sum(lock_steps) * G),
               "gpu_launches": int(launches),
               "roofline": {"bound": "tensor", "kernel": ("k_net_gr + k_net_heads_tc (grid-row trunk, four levels in one launch, + FC heads)" if net.dnet.grid_row() else "k_net_forward_tc / k_net_role + k_net_heads_tc (trunk + FC heads)"),
                            "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                            "frac": achieved / peak if achieved else None,
                            "traffic": cap.get("dram_bytes_per_launch"),
                            "flop_per_eval": flops,
                            "evals_per_launch": prof_st["expansions"] / max(1, prof["steps"]),
                            "evaluator_share_of_step": prof["evaluator_ms"] /
                            max(1e-9, prof["evaluator_ms"] + prof["expand_select_ms"]),
                            "evaluator_us_per_launch": 1e3 * prof["evaluator_ms"] / max(1, prof["steps"]),
                            "expand_select_us_per_launch": 1e3 * prof["expand_select_ms"] / max(1, prof["steps"]),
                            "measured_in": "accounting pass of the same steps (CUDA events around every call)",
                            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (measured)" if cx.peaks else "fallback"},
               "clocks": clocks}
        if e2e:
            rec["e2e"] = {"value": sims2 / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"],
                          "d2h_bytes_per_step": e2e["d2h"], "api": "bpp_engine_play_net_host (pinned host buffers: items "
                          "in; root records, visit counts, actions, outcomes out)"}
        if with_cpu and cx.world == 1 and not args.no_cpu:
            cpu_net = BinPackingNNet(_Gm(Wb, Hb), _margs(args.sims))
            cpu_net.load_state_dict({k: v.cpu() for k, v in net.nnet.state_dict().items()})
            cpu_net.eval()

            class CpuNet:
                def predict(self, board):
                    with torch.no_grad():
                        lp, v = cpu_net(torch.from_numpy(board.astype(np.float32))[None])
                    return torch.exp(lp)[0].numpy(), v[0].numpy()
            v, eps, dt = cpu_single_core(min(args.cpu_seconds, 10.0), args.sims, Wb, Hb, CpuNet())
            kind, _, _, _, _, src = _cpu_impl()
            rec["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                                   "sample": f"{eps} episodes in {dt:.1f} s; {src} + fp32 torch net on CPU"}
    bm.close()
    net.dnet.close()
    return rec if cx.rank == 0 else {"precision_mode": mode, "rank_stub": True}


# ---------------------------------------------------------------------------------------------------------------------
def run_iteration_arm(cx, games_per_gpu, iters=2, batch=512, epochs=2):
    """configs[3]: full CoachBPP iteration = batched self-play + example all-gather + data-parallel learner"""
    torch, args = cx.torch, cx.args
    import tempfile
    from resource_packing_self_play_b200.coach import CoachBPP
    from resource_packing_self_play_b200.game import ItemsGenerator
    from resource_packing_self_play_b200.nnet import NNetWrapper
    margs = _margs(args.sims, numIters=iters, numEps=20, iterStepThreshold=10 ** 9, numScoresForRank=100,
                   numItersForTrainExamplesHistory=50, maxlenOfQueue=200000, epochs=epochs, batch_size=batch, seed=100,
                   binH_min=2, binH=15, numItems=N, checkpoint=tempfile.mkdtemp(prefix="bpp_bench_"))
    torch.manual_seed(0)
    gm = _Gm(W, H)
    net = NNetWrapper(gm, margs, max_batch=games_per_gpu, device=cx.local)
    gen = ItemsGenerator(W, H, N)
    coach = CoachBPP(gm, net, None, 0, gen, margs)
    cx.barrier()
    out = coach.learn_batched(games_per_gpu * cx.world, num_iters=iters, checkpoint=False)
    last = out[-1]
    steps = epochs * max(1, int(last["history_examples"] / (batch * cx.world)))
    (t_sp, t_ga, t_tr), _ = cx.reduce([last["t_selfplay"], last["t_gather"], last["t_train"]], [0.0])
    bmx = getattr(coach, "_bm", None)
    if bmx is not None:
        bmx.close()
    net.dnet.close()
    if cx.rank != 0:
        return None
    return {"config": {"workload": f"configs[3]: one CoachBPP iteration, {games_per_gpu} games per GPU x {cx.world} GPU(s), "
                                   f"numMCTSSims={args.sims}, real net (random init), learner batch {batch} per rank, "
                                   f"{epochs} epochs over the replay history ({iters} iterations run, last reported)"},
            "selfplay_s": t_sp, "gather_s": t_ga, "train_s": t_tr, "learner_steps": steps,
            "learner_ms_per_step_incl_allreduce": 1e3 * t_tr / max(1, steps),
            "examples_this_iter": last["examples"], "history_examples": last["history_examples"],
            "games": last["games"], "episodes_per_sec_selfplay": last["games"] / max(1e-9, t_sp),
            "iteration_s": t_sp + t_ga + t_tr, "precision_mode": net.dnet.precision,
            "collectives": "NCCL all-gather of compact examples + scores; one flat 0.68 MB gradient all-reduce per "
                           "learner step" if cx.world > 1 else "none (1 GPU)",
            "limiter": "learner step count grows with the replay history; the per-step all-reduce of 0.68 MB is "
                       "latency-bound" if cx.world > 1 else "learner step count grows with the replay history",
            "mean_score": last["mean_score"], "loss_pi": last["loss_pi"], "loss_v": last["loss_v"]}


def run_arena_arm(cx, seeds_per_gpu):
    """configs[4] share: greedy evaluation sweep, two nets per seed (CoachBPP.arena_playing semantics, batched)"""
    torch, args = cx.torch, cx.args
    from resource_packing_self_play_b200.coach import CoachBPP
    from resource_packing_self_play_b200.game import ItemsGenerator
    from resource_packing_self_play_b200.nnet import NNetWrapper
    margs = _margs(args.sims, seed=100, arenaCompare=seeds_per_gpu, numItems=N, checkpoint="/tmp")
    gm = _Gm(W, H)
    torch.manual_seed(1)
    pnet = NNetWrapper(gm, margs, max_batch=seeds_per_gpu, device=cx.local)
    torch.manual_seed(2)
    nnet = NNetWrapper(gm, margs, max_batch=seeds_per_gpu, device=cx.local)
    gen = ItemsGenerator(W, H, N)
    coach = CoachBPP(gm, nnet, None, 0, gen, margs)
    seeds, hts, _ = workload(900000 + cx.rank * seeds_per_gpu, seeds_per_gpu)
    coach.arena_sweep(pnet, nnet, seeds[:seeds_per_gpu], hts[:seeds_per_gpu], seed=3)   # warm-up (pools, plans)
    ms, wall, clocks, res = cx.timed(lambda k: coach.arena_sweep(pnet, nnet, seeds, hts, seed=5 + k), 1)
    p, n, acc = res[0]
    (t,), (ps, ns, cnt) = cx.reduce([max(ms, wall)], [float(p.sum()), float(n.sum()), float(len(p))])
    bmx = getattr(coach, "_bm", None)
    if bmx is not None:
        bmx.close()
    pnet.dnet.close()
    nnet.dnet.close()
    if cx.rank != 0:
        return None
    return {"config": {"workload": f"configs[4] share: {seeds_per_gpu} item sequences per GPU x {cx.world} GPU(s), each played "
                                   f"greedily (greedy_a=0) with the previous and the new net, numMCTSSims={args.sims}"},
            "sequences": int(cnt), "episodes": int(2 * cnt), "seconds": t * 1e-3,
            "episodes_per_sec": 2 * cnt / (t * 1e-3), "mean_score_prev": ps / cnt, "mean_score_new": ns / cnt,
            "accept": int(ns >= ps), "clocks": clocks}


def run_shard_checksum(cx, block=256):
    """SURVEY §4 test-plan item 5 on hardware: a fixed block of `block` seeds, sharded over the ranks, stub evaluator,
    deterministic action choice; sha256 over the per-move visit-count matrices in seed order — equal at any N"""
    torch = cx.torch
    from resource_packing_self_play_b200 import _lib
    from resource_packing_self_play_b200.distributed import all_gather_variable, shard_range
    from resource_packing_self_play_b200.engine import SearchEngine
    from resource_packing_self_play_b200.game import ItemsGenerator
    lo, hi = shard_range(block, cx.rank, cx.world)
    seeds, hts, areas = workload(500000, block)
    gen = ItemsGenerator(W, H, N)
    G = hi - lo
    eng = SearchEngine(W, H, N, G, cx.args.sims, CPUCT, device=cx.local)
    items = gen.items_batch_device(seeds[lo:hi], hts[lo:hi], device=cx.local)
    eng.reset(items, areas[lo:hi], np.full(G, 0.7001))
    counts, actions = eng.play_stub("V", _lib.CHOOSE_ARGMAX_FIRST)
    eng.check()
    allc = all_gather_variable(counts.movedim(1, 0).contiguous())      # (block, N, A) in seed order
    alla = all_gather_variable(actions.movedim(1, 0).contiguous())
    eng.close()
    if cx.rank != 0:
        return None
    h = hashlib.sha256()
    h.update(allc.cpu().numpy().tobytes())
    h.update(alla.cpu().numpy().tobytes())
    return {"sha256": h.hexdigest(), "seeds": block, "evaluator": "stub V, first arg-max", "ranks": cx.world}


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--games", type=int, default=4096)
    ap.add_argument("--sims", type=int, default=SIMS)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--edge-frac", type=float, default=1.0, help="edge pool as a fraction of the worst case")
    ap.add_argument("--workload", default="all",
                    help="comma list of: stub (headline), real15, real20, iteration, arena, checksum; all = every one")
    ap.add_argument("--precision", default="auto", help="precision mode of a single real15/real20 run")
    ap.add_argument("--stub-stream-mult", type=int, default=8,
                    help="headline workload: episodes per step = mult x games, streamed through the resident games "
                         "(1 = one episode per game and step, bpp_engine_play_stub)")
    ap.add_argument("--stream-mult", type=int, default=0,
                    help="real-net workloads: episodes per step = stream-mult x games, streamed through the resident games")
    ap.add_argument("--real20-large-games", type=int, default=16384,
                    help="resident games per GPU of the second real20 record (0 = skip): the 20x20 evaluator runs at 190 "
                         "TFLOP/s on 2,800 leaves, 285 on 8,192 and more, and this arm is evaluator-bound")
    ap.add_argument("--iteration-games", type=int, default=8192)
    ap.add_argument("--arena-seeds", type=int, default=8192)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    todo = set(args.workload.split(","))
    if "all" in todo:
        todo = {"stub", "real15", "real20", "iteration", "arena", "checksum"}
    cx = Ctx(args)
    t_start = time.perf_counter()
    line, sec = None, {}

    def release():
        """hand the finished workload's buffers back to the driver (device blocks and pinned host blocks that torch's caching
        allocators would otherwise keep for the rest of the process): every workload starts from the same memory state"""
        import gc
        gc.collect()
        cx.torch.cuda.empty_cache()
        try:
            cx.torch._C._host_emptyCache()
        except Exception:  # noqa: BLE001
            pass

    if "stub" in todo:
        line = run_stub_arm(cx)
        release()
        if args.stub_stream_mult > 1:
            # the same games with ONE episode per game and step (the launch ends with an idle tail while the longest
            # episodes finish): the figure of round 1 and of the ncu captures, for comparison
            one = run_stub_arm(cx, mult=1, steps=max(5, args.steps // 2), with_e2e=False)
            if line is not None and one is not None:
                line["one_episode_per_game"] = {k: one[k] for k in ("value", "ms_per_step", "steps", "episodes_per_sec")}
                line["one_episode_per_game"]["kernel_ms_per_launch"] = one["roofline"]["kernel_ms_per_launch"]

    def guarded(name, fn):
        """a failing secondary workload must not take the headline line with it: its record carries the error"""
        try:
            t0 = time.perf_counter()
            r = fn()
            if r is not None:
                r["wall_s"] = time.perf_counter() - t0
            release()
            return r
        except Exception as err:  # noqa: BLE001
            if cx.world > 1:
                raise  # ranks must fail together
            return {"error": f"{type(err).__name__}: {err}"}
    if "real15" in todo:
        r = guarded("real15", lambda: run_real_arm(cx, 15, 15, args.precision))
        # every rank must take the same branch (the arms contain collectives): the mode the calibration picked is a function
        # of the weights and identical on all ranks; only rank 0 holds the full record
        if r is not None and "error" not in r and args.precision == "auto" and r["precision_mode"] != "bf16":
            b = guarded("real15_bf16", lambda: run_real_arm(cx, 15, 15, "bf16", with_e2e=False, with_cpu=False))
            if b is not None and "error" not in b and not b.get("rank_stub"):
                r["bf16"] = {"value": b["value"], "leaf_evals_per_sec": b["leaf_evals_per_sec"],
                             "roofline": b["roofline"], "lockstep_steps_per_batch": b["lockstep_steps_per_batch"],
                             "note": "plain bf16 on this trained checkpoint is OUT of the stated tolerance (|d pi| up to "
                                     "0.6, profiles/r02_precision_study.txt): throughput shown for comparison only"}
            elif not r.get("rank_stub"):
                r["bf16"] = b
        sec["real15"] = r
    if "real20" in todo:
        sec["real20"] = guarded("real20", lambda: run_real_arm(cx, 20, 20, args.precision))
        if args.real20_large_games and args.real20_large_games != args.games:
            sec[f"real20_g{args.real20_large_games}"] = guarded("real20_large", lambda: run_real_arm(
                cx, 20, 20, args.precision, with_cpu=False, steps=3, games=args.real20_large_games))
    if "iteration" in todo:
        sec["iteration"] = guarded("iteration", lambda: run_iteration_arm(cx, args.iteration_games))
    if "arena" in todo:
        sec["arena"] = guarded("arena", lambda: run_arena_arm(cx, args.arena_seeds))
    if "checksum" in todo:
        sec["shard_checksum"] = guarded("checksum", lambda: run_shard_checksum(cx))
    if cx.rank == 0:
        if line is None:  # a single secondary workload was asked for: its record is the line
            name = next(iter(k for k, v in sec.items() if v is not None), None)
            r = sec.get(name) or {}
            line = {"metric": METRIC, "value": r.get("value"), "unit": UNIT, "n_gpus": cx.world,
                    "steps": r.get("steps"), "warmup": r.get("warmup"), "ms_per_step": r.get("ms_per_step"),
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": r.get("dtype"),
                    "data": "synthetic", "config": r.get("config"), "e2e": r.get("e2e"), "roofline": r.get("roofline"),
                    "clocks": r.get("clocks"), "gpu_launches": r.get("gpu_launches")}
        line["secondary"] = {k: v for k, v in sec.items() if v is not None and not (isinstance(v, dict) and v.get("rank_stub"))}
        line["bench_wall_s"] = time.perf_counter() - t_start
        print(json.dumps(line))
    cx.close()


if __name__ == "__main__":
    main()
