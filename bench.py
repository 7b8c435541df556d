#!/usr/bin/env python
"""bench.py — MCTS simulations/s of the self-play hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--games G] [--sims S]

Workload (BASELINE.json configs[1]): the main_bpp.py default instance — 15x15 virtual bin,
10 items from ItemsGenerator(15, h, 10) with h ~ randint(2, 16) per 20 consecutive seeds, total_area = 15*h,
numMCTSSims = 200, cpuct = 1, alpha = 0.75, empty rewards list — as G = 4096 lockstep games per GPU with the stub
uniform-prior evaluator (p = 1/A, v = 0).  One STEP = one batch of G complete self-play episodes (reset, then per
move: 200 simulations per game, visit counts out, action ~ counts (CoachBPP.py:86-87), getNextState + getGameEnded).
Seeds: generator_seed = 1000 + global episode index.  Scaling is weak: every rank plays its own G games.

value   = simulations/s with the instances already resident in HBM (whole job, all ranks, max-over-ranks time)
e2e     = the same through the host-buffer C-ABI call bpp_engine_play_stub_host: instances uploaded from pinned host
          memory and visit counts / actions / rewards downloaded inside the timed region
roofline= dominant kernel k_episode<U> (all simulations of a step in one launch): algorithmic bytes (SURVEY.md §8(d) formula with the kernel-counted edges and
          expansions per simulation) / CUDA-event time of that kernel, against MEASURED_PEAKS.json hbm_gbs
cpu_baseline = oracle/bpp_oracle.py (a port that keeps the reference's data structures) on ONE host core, bounded
          sample of the same workload.  `--impl reference` runs that port on all host cores instead of the GPU.
"""
import argparse
import ctypes as C
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


def workload(first_episode, count):
    """(seeds, generator heights, total areas) of `count` consecutive episodes starting at `first_episode`."""
    idx = np.arange(first_episode, first_episode + count)
    seeds = 1000 + idx
    # height per batch of 20 episodes (CoachBPP.py:117-119 draws it once per iteration of numEps = 20 episodes)
    batch = idx // 20
    heights = np.array([np.random.RandomState(77000 + int(b)).randint(2, 16) for b in batch], dtype=np.int32)
    return seeds, heights, (W * heights).astype(np.int32)


def config_dict(args, extra=None):
    c = {"workload": "configs[1]: main_bpp default instance (15x15 bin, 10 items, numMCTSSims=200, cpuct=1), "
                     f"{args.games} lockstep games per GPU, stub uniform-prior net, whole self-play episodes",
         "games_per_gpu": args.games, "num_mcts_sims": args.sims, "bin": [W, H], "items": N,
         "action_choice": "sample ~ visit counts (greedy=False)",
         "cache": "working set (search graphs of all games, several GB per step) is far larger than the 126 MB L2; "
                  "no explicit flush"}
    if extra:
        c.update(extra)
    return c


# ---------------------------------------------------------------------------------------------------------------------
# CPU arms (the only place bench.py executes oracle/)
def _cpu_episode(ep_index, sims):
    from oracle import bpp_oracle as O
    seeds, heights, areas = workload(ep_index, 1)
    items = O.OracleItemsGenerator(W, int(heights[0]), N).items_generator(int(seeds[0]))
    g = O.OracleGame(W, H, N, 1)
    m = O.OracleMCTS(g, O.StubNet("U", g.getActionSize()), O.dotdict(numMCTSSims=sims, cpuct=CPUCT, alpha=ALPHA))
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
    return moves * sims, m.n_edges_walked, m.n_expansions


def _cpu_worker(q_in, q_out, sims):
    os.environ["OMP_NUM_THREADS"] = "1"
    while True:
        ep = q_in.get()
        if ep is None:
            return
        q_out.put(_cpu_episode(ep, sims))


def cpu_single_core(budget_s, sims):
    """time the port on one core for about budget_s seconds of whole episodes"""
    _cpu_episode(0, sims)  # warm-up (imports, allocator)
    t0 = time.perf_counter()
    n_sims = eps = 0
    while time.perf_counter() - t0 < budget_s:
        s, _, _ = _cpu_episode(1 + eps, sims)
        n_sims += s
        eps += 1
    dt = time.perf_counter() - t0
    return n_sims / dt, eps, dt


def run_reference_arm(args):
    """The reference's CPU implementation of the path (the Python port; the reference itself is pure Python and is not
    present on the GPU box) on all host cores: one episode stream per core, summed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    ctx = mp.get_context("fork")
    q_in, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(q_in, q_out, args.sims), daemon=True) for _ in range(cores)]
    for p in procs:
        p.start()
    eps_per_step = cores  # one episode per core per step

    def step(k):
        for i in range(eps_per_step):
            q_in.put(k * eps_per_step + i)
        return sum(q_out.get()[0] for _ in range(eps_per_step))

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
    sample = f"{eps_per_step} episodes per step (one per core), {args.steps} steps, python port oracle/bpp_oracle.py"
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, {"note": "CPU arm: whole episodes of the same instance distribution"}),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "episodes_per_sec": args.steps * eps_per_step / dt, "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
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
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from resource_packing_self_play_b200 import _lib
    from resource_packing_self_play_b200.engine import SearchEngine, algorithmic_bytes_per_sim
    from resource_packing_self_play_b200.game import ItemsGenerator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    G = args.games
    n_steps_total = args.warmup + args.steps
    gen = ItemsGenerator(W, H, N)
    # synthetic instances for every step of this rank (host generation is setup, not part of the timed path)
    inst = []
    for k in range(2 * n_steps_total):  # first half: device-resident arm, second half: e2e arm
        first = ((k * world) + rank) * G
        seeds, heights, areas = workload(first, G)
        # the package's device-side generator (bit-identical to numpy's legacy RNG path, tests/test_gpu_env.py)
        inst.append((gen.items_batch_device(seeds, heights, device=local).cpu().numpy(), areas))
    edge_cap = 0
    if args.edge_frac < 1.0:  # profiling runs: a smaller edge pool keeps ncu's save/restore cheap
        worst_units_per_node = 3 * ((W * N + 3) // 4 * 4) + ((W * N + 3) // 4 * 4) // 4
        edge_cap = int((args.sims * N + N + 2) * worst_units_per_node * args.edge_frac)
    eng = SearchEngine(W, H, N, G, args.sims, CPUCT, device=local, edge_cap=edge_cap)
    nan_bl = np.full(G, np.nan)
    bl_dev = torch.from_numpy(nan_bl).to(dev)
    counts_buf = torch.zeros((N, G, W * N), dtype=torch.int32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- arm 1: inputs resident in HBM ------------------------------------------------------------------------------
    dev_inst = [(torch.from_numpy(i).to(dev), torch.from_numpy(a).to(dev)) for i, a in inst[:n_steps_total]]
    search_events = []

    actions_buf = torch.empty((N, G), dtype=torch.int32, device=dev)

    def step_resident(k, timed):
        """reset + ONE launch of the whole-episode kernel (search -> counts -> choose -> play, N moves)"""
        items, area = dev_inst[k]
        eng.reset(items, area, bl_dev)
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _lib.call("bpp_engine_play_stub", eng._h, _lib.STUB["U"], _lib.CHOOSE_SAMPLE, C.c_uint64(1234 + k), 0,
                  C.c_void_p(counts_buf.data_ptr()), C.c_void_p(actions_buf.data_ptr()), None, eng_stream())
        if timed:
            e1.record()
            search_events.append((e0, e1))

    def eng_stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    for k in range(args.warmup):
        step_resident(k, False)
    eng.check()
    eng.stats(reset=True)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(args.steps):
        step_resident(args.warmup + k, True)
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    clocks = sampler.stop()
    eng.check()
    st = eng.stats(reset=True)
    done = eng.status()["done"]
    assert bool((done == 1).all()), "some games did not finish their episode"
    g_nodes, g_units = eng.graph_sizes()
    graph = {"max_nodes_per_game": int(g_nodes.max()), "max_edge_units_per_game": int(g_units.max()),
             "mean_edge_units_per_game": float(g_units.double().mean()), "engine_device_bytes": eng.device_bytes}
    search_ms = sum(a.elapsed_time(b) for a, b in search_events)
    n_search = len(search_events)

    # ---- arm 2: end to end through the host-buffer C ABI --------------------------------------------------------------
    pinned_items = [torch.from_numpy(i).pin_memory() for i, _ in inst[n_steps_total:]]
    pinned_area = [torch.from_numpy(a).pin_memory() for _, a in inst[n_steps_total:]]
    out = {"counts": torch.empty((N, G, W * N), dtype=torch.int32).pin_memory().numpy(),
           "actions": torch.empty((N, G), dtype=torch.int32).pin_memory().numpy(),
           "r": torch.empty(G, dtype=torch.int32).pin_memory().numpy(),
           "score": torch.empty(G, dtype=torch.float64).pin_memory().numpy(),
           "moves": torch.empty(G, dtype=torch.int32).pin_memory().numpy()}
    pinned_bl = torch.from_numpy(nan_bl).pin_memory().numpy()

    def step_e2e(k):
        eng.play_stub_host("U", pinned_items[k].numpy(), pinned_area[k].numpy(), pinned_bl,
                           choose_mode=_lib.CHOOSE_SAMPLE, seed=99 + k, out=out)
        return int(out["moves"].sum())

    for k in range(args.warmup):
        step_e2e(k)
    eng.stats(reset=True)
    barrier()
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_moves = 0
    for k in range(args.steps):
        e2e_moves += step_e2e(args.warmup + k)
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - w0))  # conservative: device events vs wall clock
    st2 = eng.stats(reset=True)
    h2d = inst[0][0].nbytes + inst[0][1].nbytes + nan_bl.nbytes
    d2h = sum(v.nbytes for v in out.values())

    # ---- reduce over ranks ------------------------------------------------------------------------------------------
    vals = torch.tensor([ms, e2e_ms, search_ms], dtype=torch.float64, device=dev)
    sums = torch.tensor([st["sims"], st2["sims"], st["edges"], st["expansions"], st["launches"] + st2["launches"],
                         args.steps * G], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms, e2e_ms, search_ms_max = vals.tolist()
    sims, sims2, edges, exps, launches, episodes = sums.tolist()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        d_bar, e_bar = st["edges"] / st["sims"], st["expansions"] / st["sims"]
        bytes_per_sim = algorithmic_bytes_per_sim(W, H, N, d_bar, e_bar)
        # rank-0 kernel figures (per launch): algorithmic bytes of the simulations one launch processes / its duration
        sims_per_launch = st["sims"] / n_search
        achieved = bytes_per_sim * sims_per_launch / (search_ms / n_search * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": sims / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, {"parallelism": f"games sharded over {world} GPU(s), no data-path collective"}),
            "episodes_per_sec": episodes / (ms * 1e-3),
            "e2e": {"value": sims2 / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "episodes_per_sec": episodes / (e2e_ms * 1e-3),
                    "api": "bpp_engine_play_stub_host (pinned host buffers)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_episode<STUB_U,15> (whole episodes, one launch per step)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6650 GB/s",
                         # DRAM read+write bytes of one k_episode launch from the committed ncu --set full capture
                         # (profiles/r01_k_episode_ncu_summary.txt); only valid for the default configuration
                         "traffic": 5.225e9 if (G == 4096 and args.sims == 200) else None,
                         "algorithmic_bytes_per_sim": bytes_per_sim,
                         "edges_per_sim": d_bar, "expansions_per_sim": e_bar, "sims_per_launch": sims_per_launch,
                         "kernel_ms_per_launch": search_ms / n_search,
                         "kernel_share_of_step": search_ms / (ms if world == 1 else search_ms_max or ms),
                         "note": "latency/issue-bound pointer chasing: one warp per game, strictly sequential "
                                 "simulations; see DESIGN.md"},
            "clocks": clocks, "graph": graph,
        }
        if world == 1 and not args.no_cpu:
            v, eps, dt = cpu_single_core(args.cpu_seconds, args.sims)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{eps} whole episodes of the same workload in {dt:.1f} s, "
                                              "oracle/bpp_oracle.py (python port, reference data structures)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# secondary workloads (not the headline): lockstep self-play with the REAL policy/value net as leaf evaluator
# (BASELINE.json configs[2]: 20x20 bin, numMCTSSims = 200, bf16 tensor-core forward; "real15" = the default instance)
NET_FLOPS = {(15, 15, 10): 4394592, (20, 20, 10): 7249920}


def _cpu_real_episode(ep_index, sims, Wb, Hb, net):
    from oracle import bpp_oracle as O
    rs = np.random.RandomState(77000 + ep_index // 20)
    gh = int(rs.randint(2, Hb + 1))
    items = O.OracleItemsGenerator(Wb, gh, N).items_generator(1000 + ep_index)
    g = O.OracleGame(Wb, Hb, N, 1)
    m = O.OracleMCTS(g, net, O.dotdict(numMCTSSims=sims, cpuct=CPUCT, alpha=ALPHA))
    rng = np.random.RandomState(ep_index)
    board, planes = g.getInitBoard(), g.getInitItems(items)
    moves = 0
    while True:
        state = g.getBinItem(board, planes)
        pi = m.getActionProb(state, Wb * gh, [])
        a = int(rng.choice(len(pi), p=pi))
        board, planes = g.getNextState(board, a, planes)
        moves += 1
        if g.getGameEnded(g.getBinItem(board, planes), Wb * gh, [], ALPHA)[0] != 0:
            return moves * sims


NET_TRAFFIC_NCU = {(15, 15, 10): 706.3e3 + 776.4e3}


def run_real_arm(args):
    import torch
    from resource_packing_self_play_b200 import _lib
    from resource_packing_self_play_b200.game import ItemsGenerator
    from resource_packing_self_play_b200.mcts import BatchedMCTS
    from resource_packing_self_play_b200.nnet import BinPackingNNet, NNetWrapper
    from resource_packing_self_play_b200.utils import dotdict

    Wb, Hb = (20, 20) if args.workload == "real20" else (15, 15)
    G = args.games
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)

    class Gm:
        bin_width, bin_height, num_items = Wb, Hb, N

        def getBoardSize(self):
            return (Hb, Wb)

        def getActionSize(self):
            return Wb * N
    margs = dotdict(numMCTSSims=args.sims, cpuct=CPUCT, alpha=ALPHA, num_items=N, num_bins=1, cuda=True)
    torch.manual_seed(0)
    net = NNetWrapper(Gm(), margs, max_batch=G, precision="bf16")
    bm = BatchedMCTS(Gm(), net, margs, G)
    gen = ItemsGenerator(Wb, Hb, N)

    def instances(k):
        idx = np.arange(k * G, (k + 1) * G)
        hts = np.array([np.random.RandomState(77000 + int(b)).randint(2, Hb + 1) for b in idx // 20], dtype=np.int32)
        return gen.items_batch(1000 + idx, hts), (Wb * hts).astype(np.int32)
    inst = [instances(k) for k in range(args.warmup + args.steps)]
    fwd_events = []
    orig_forward = net.dnet.forward

    def timed_forward(*a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_forward(*a, **kw)
        e1.record()
        fwd_events.append((e0, e1))
        return r

    def step(k, timed):
        items, area = inst[k]
        bm.reset(items, area, [])
        net.dnet.forward = timed_forward if timed else orig_forward
        out = None
        for m in range(N):
            counts = bm.search()
            act = bm.eng.choose(_lib.CHOOSE_SAMPLE, seed=7 + k)
            bm.eng.advance(act)
            out = counts
        return out
    for k in range(args.warmup):
        step(k, False)
    # kernel-level accounting pass: the same steps with the lockstep chunks launched eagerly and CUDA events around every
    # forward (events cannot be read inside a captured graph); the timed pass below replays the captured chunks
    bm.use_graphs = False
    bm.eng.stats(reset=True)
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for k in range(args.steps):
        step(args.warmup + k, True)
    a1.record()
    torch.cuda.synchronize()
    eager_ms = a0.elapsed_time(a1)
    eager_st = bm.eng.stats(reset=True)
    fwd_ms = sum(a.elapsed_time(b) for a, b in fwd_events)
    n_fwd = len(fwd_events)
    bm.use_graphs = os.environ.get("BPP_NO_GRAPHS") is None
    for k in range(args.warmup):   # captures the chunk graphs
        step(k, False)
    bm.eng.check()
    bm.eng.stats(reset=True)
    steps_before, graph_before = bm.steps, bm.graph_launches
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(args.steps):
        step(args.warmup + k, False)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    clocks = sampler.stop()
    bm.eng.check()
    st = bm.eng.stats(reset=True)
    graph_launches = bm.graph_launches - graph_before
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    flops = NET_FLOPS[(Wb, Hb, N)]
    achieved = eager_st["expansions"] * flops / (fwd_ms * 1e-3) / 1e12
    line = {"metric": METRIC, "value": st["sims"] / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 (net) / f64 (tree)", "data": "synthetic",
            "config": {"workload": f"configs[2]-style: {Wb}x{Hb} bin, 10 items, numMCTSSims={args.sims}, real "
                                   f"policy/value net (random init, seed 0) in bf16 on tcgen05, {G} lockstep games",
                       "games_per_gpu": G},
            "episodes_per_sec": args.steps * G / (ms * 1e-3), "leaf_evals_per_sec": st["expansions"] / (ms * 1e-3),
            "lockstep_steps": bm.steps - steps_before,
            # kernels inside replayed graphs (k_search, trunk, heads, k_expand_backup per lockstep step) + eager launches
            "gpu_launches": graph_launches + st["launches"] + (0 if bm.use_graphs else 2 * (bm.steps - steps_before)),
            "cuda_graphs": bool(bm.use_graphs), "ms_per_step_eager_launches": eager_ms / args.steps,
            "roofline": {"bound": "tensor", "kernel": "k_net_forward_tc + k_net_heads_tc", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak,
                         # dram__bytes_read + write per launch of trunk + heads, ncu on this very workload
                         # (profiles/r01_real15_lockstep_kernels_ncu.csv): weights and leaf records only, everything
                         # else stays in shared / tensor memory; not captured for the other geometries
                         "traffic": NET_TRAFFIC_NCU.get((Wb, Hb, N)),
                         "flop_per_eval": flops, "evals_per_launch": eager_st["expansions"] / max(1, n_fwd),
                         "kernel_share_of_step": fwd_ms / eager_ms,
                         "measured_in": "eager-launch pass of the same steps (CUDA events around every forward)",
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (measured)" if peaks else "fallback"},
            "clocks": clocks}
    if not args.no_cpu:
        cpu_net = BinPackingNNet(Gm(), margs)
        cpu_net.load_state_dict({k: v.cpu() for k, v in net.nnet.state_dict().items()})
        cpu_net.eval()

        class CpuNet:
            def predict(self, board):
                with torch.no_grad():
                    lp, v = cpu_net(torch.from_numpy(board.astype(np.float32))[None])
                return torch.exp(lp)[0].numpy(), v[0].numpy()
        t = time.perf_counter()
        n = 0
        ep = 0
        while time.perf_counter() - t < args.cpu_seconds:
            n += _cpu_real_episode(ep, args.sims, Wb, Hb, CpuNet())
            ep += 1
        dt = time.perf_counter() - t
        line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{ep} episodes in {dt:.1f} s; oracle port + fp32 torch net on CPU"}
    print(json.dumps(line))


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
    ap.add_argument("--workload", default="stub", choices=["stub", "real15", "real20"],
                    help="stub = headline (configs[1]); real15/real20 = real net as leaf evaluator (secondary)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload != "stub":
        run_real_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
