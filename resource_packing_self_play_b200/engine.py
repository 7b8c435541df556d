"""Host-side handle of the device-resident search engine (csrc/bpp_engine.cu) and of the stateless env ops.

PyTorch is used only for device memory, streams and (elsewhere) torch.distributed; all arithmetic of the hot path
runs in the CUDA kernels behind the C ABI (include/bpp_b200.h).
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import Config, REC_REM, REC_WORDS, STUB, call

_TORCH_DT = {torch.float32: _lib.DTYPE_F32, torch.float64: _lib.DTYPE_F64}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(x, dtype, device):
    """host list / numpy / torch (any device) -> contiguous CUDA tensor of `dtype` on `device`."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    return t.to(device=device, dtype=dtype, non_blocking=True).contiguous()


def _devidx(device):
    """None -> the calling process' current CUDA device (one process per GPU: torch.cuda.set_device(LOCAL_RANK))"""
    return torch.cuda.current_device() if device is None else int(device)


def _ptr(t):
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, C.c_void_p):  # raw device pointer owned by a handle (e.g. the engine's leaf buffers)
        return t
    return C.c_void_p(t.data_ptr())


def ranked_threshold(rewards_list, alpha):
    """bl of BinPackingGame.getRankedReward (BinPackingGame.py:203-206): sorted(rewards)[floor(len*alpha) - 1]
    (index -1 wraps for short lists); NaN encodes the empty list (every terminal state is a win)."""
    if len(rewards_list) == 0:
        return float("nan")
    srt = np.sort(np.asarray(rewards_list, dtype=np.float64))
    return float(srt[int(np.floor(len(srt) * alpha)) - 1])


# ----------------------------------------------------------------------------------------------------------------------
# compact records <-> the reference's (N+1, H, W) state tensor (host-side format conversion)
def pack_states(states, W, H, N):
    """states: (n, N+1, H, W) integer array (the reference state tensor) ->
    (recs uint32 (n, 32), items_wh int32 (n, N, 2)).  Item dims are read back from the planes exactly like
    BinPackingLogic.py:84-85 (w = sum of row 0, h = sum of column 0); placed items (all-zero planes) get (0, 0)."""
    st = np.asarray(states)
    if st.ndim == 3:
        st = st[None]
    n = st.shape[0]
    assert st.shape[1:] == (N + 1, H, W), (st.shape, (N + 1, H, W))
    recs = np.zeros((n, REC_WORDS), dtype=np.uint32)
    weights = (1 << np.arange(W, dtype=np.uint64))
    recs[:, :H] = (st[:, 0].astype(np.uint64) * weights).sum(axis=2).astype(np.uint32)
    planes = st[:, 1:]
    remaining = planes.reshape(n, N, -1).sum(axis=2) > 0
    recs[:, REC_REM] = (remaining.astype(np.uint32) << np.arange(N, dtype=np.uint32)).sum(axis=1).astype(np.uint32)
    items = np.zeros((n, N, 2), dtype=np.int32)
    items[:, :, 0] = planes[:, :, 0, :].sum(axis=2)
    items[:, :, 1] = planes[:, :, :, 0].sum(axis=2)
    return recs, items


def unpack_states(recs, items_wh, W, H, N, dtype=np.int64):
    """inverse of pack_states (items_wh supplies the dims of the remaining items)."""
    recs = np.asarray(recs, dtype=np.uint32).reshape(-1, REC_WORDS)
    items_wh = np.asarray(items_wh).reshape(recs.shape[0], N, 2)
    n = recs.shape[0]
    st = np.zeros((n, N + 1, H, W), dtype=dtype)
    st[:, 0] = (recs[:, :H, None] >> np.arange(W, dtype=np.uint32)) & 1
    rem = (recs[:, REC_REM, None] >> np.arange(N, dtype=np.uint32)) & 1
    ys = np.arange(H)[None, None, :, None]
    xs = np.arange(W)[None, None, None, :]
    st[:, 1:] = ((ys < items_wh[:, :, 1, None, None]) & (xs < items_wh[:, :, 0, None, None]) &
                 (rem[:, :, None, None] > 0))
    return st


# ----------------------------------------------------------------------------------------------------------------------
class EnvOps:
    """Batched stateless environment ops on device tensors (bpp_env_* entry points)."""

    def __init__(self, W, H, N, device=None):
        _lib.load()
        self.W, self.H, self.N, self.A = W, H, N, W * N
        self.device = torch.device("cuda", _devidx(device))

    def valid_moves(self, recs, items_wh):
        recs = self._recs(recs)
        items = _dev(items_wh, torch.int32, self.device)
        n = recs.shape[0]
        out = torch.empty((n, self.A), dtype=torch.uint8, device=self.device)
        call("bpp_env_valid_moves", self.W, self.H, self.N, n, _ptr(recs), _ptr(items), _ptr(out), _stream())
        return out

    def planes(self, recs, items_wh):
        """dense float32 (n, N+1, H, W) input planes of compact states (items_wh per state)"""
        recs = self._recs(recs)
        items = _dev(items_wh, torch.int32, self.device)
        n = recs.shape[0]
        out = torch.empty((n, self.N + 1, self.H, self.W), dtype=torch.float32, device=self.device)
        call("bpp_env_planes", self.W, self.H, self.N, n, _ptr(recs), _ptr(items), _ptr(out), _stream())
        return out

    def next_state(self, recs, items_wh, actions):
        recs = self._recs(recs)
        items = _dev(items_wh, torch.int32, self.device)
        act = _dev(actions, torch.int32, self.device)
        n = recs.shape[0]
        out = torch.empty_like(recs)
        call("bpp_env_next_state", self.W, self.H, self.N, n, _ptr(recs), _ptr(items), _ptr(act), _ptr(out), _stream())
        return out

    def game_ended(self, recs, items_wh, total_area, max_h, bl, tie=None):
        recs = self._recs(recs)
        items = _dev(items_wh, torch.int32, self.device)
        area = _dev(total_area, torch.int32, self.device)
        mh = _dev(max_h, torch.int32, self.device)
        blt = _dev(bl, torch.float64, self.device)
        tiet = _dev(tie, torch.int8, self.device) if tie is not None else None
        n = recs.shape[0]
        ended = torch.empty(n, dtype=torch.int32, device=self.device)
        score = torch.empty(n, dtype=torch.float64, device=self.device)
        call("bpp_env_game_ended", self.W, self.H, self.N, n, _ptr(recs), _ptr(items), _ptr(area), _ptr(mh), _ptr(blt),
             _ptr(tiet), _ptr(ended), _ptr(score), _stream())
        return ended, score

    def _recs(self, recs):
        if isinstance(recs, torch.Tensor):
            t = recs
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(recs, dtype=np.uint32)).view(np.int32))
        t = t.to(self.device).contiguous()
        if t.dtype != torch.int32:
            t = t.view(torch.int32) if t.element_size() == 4 else t.to(torch.int32)
        return t.reshape(-1, REC_WORDS)


# ----------------------------------------------------------------------------------------------------------------------
class SearchEngine:
    """G lockstep games with device-resident search graphs (one handle per device)."""

    def __init__(self, W, H, N, G, num_sims, cpuct=1.0, device=None, node_cap=0, edge_cap=0):
        _lib.load()
        self.W, self.H, self.N, self.G, self.A = W, H, N, G, W * N
        self.num_sims, self.cpuct = int(num_sims), float(cpuct)
        device = _devidx(device)
        self.device = torch.device("cuda", device)
        cfg = Config(W, H, N, G, int(num_sims), float(cpuct), int(node_cap), int(edge_cap), int(device))
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            call("bpp_engine_create", C.byref(cfg), C.byref(h))
        self._h = h
        self._keep = []
        nvp = (self.A + 3) // 4 * 4
        node_cap = int(node_cap) if node_cap else self.num_sims * N + N + 2
        self.edge_cap_worst = node_cap * (3 * nvp + nvp // 4)   # every node expanded with all A actions valid
        cap = C.c_int64(0)
        call("bpp_engine_edge_cap", self._h, C.byref(cap))
        self.edge_cap = int(cap.value)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.load().bpp_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_bytes(self):
        return int(_lib.load().bpp_engine_device_bytes(self._h))

    # -- episode / move control -----------------------------------------------------------------------------------
    def reset(self, items_wh, total_area, bl=None, tie=None):
        """items_wh (G, N, 2), total_area (G,), bl (G,) float64 (NaN = empty rewards list), tie (G,) int8."""
        items = _dev(items_wh, torch.int32, self.device).reshape(self.G, self.N, 2)
        area = _dev(total_area, torch.int32, self.device).reshape(self.G)
        if bl is None:
            bl = np.full(self.G, np.nan)
        blt = _dev(bl, torch.float64, self.device).reshape(self.G)
        tiet = _dev(tie, torch.int8, self.device).reshape(self.G) if tie is not None else None
        self._keep = [items, area, blt, tiet]
        self.items_wh = items
        call("bpp_engine_reset", self._h, _ptr(items), _ptr(area), _ptr(blt), _ptr(tiet), _stream())

    def reset_host(self, items_wh, total_area, bl, tie=None):
        """same, from HOST numpy buffers through the C ABI's own host entry point (copies inside the call)."""
        items = np.ascontiguousarray(items_wh, dtype=np.int32)
        area = np.ascontiguousarray(total_area, dtype=np.int32)
        blh = np.ascontiguousarray(bl, dtype=np.float64)
        tieh = np.ascontiguousarray(tie, dtype=np.int8) if tie is not None else None
        self._keep = [items, area, blh, tieh]
        call("bpp_engine_reset_host", self._h, items.ctypes.data_as(C.c_void_p), area.ctypes.data_as(C.c_void_p),
             blh.ctypes.data_as(C.c_void_p), tieh.ctypes.data_as(C.c_void_p) if tieh is not None else C.c_void_p(0),
             _stream())

    def set_roots(self, recs):
        t = _dev(np.asarray(recs, dtype=np.uint32).view(np.int32) if not isinstance(recs, torch.Tensor) else recs,
                 torch.int32, self.device).reshape(self.G, REC_WORDS)
        call("bpp_engine_set_roots", self._h, _ptr(t), _stream())

    def set_max_h(self, max_h):
        t = _dev(max_h, torch.int32, self.device).reshape(self.G)
        self._keep_mh = t
        call("bpp_engine_set_max_h", self._h, _ptr(t), _stream())

    def set_num_sims(self, n):
        call("bpp_engine_set_num_sims", self._h, int(n))
        self.num_sims = int(n)

    def set_select_cap(self, k):
        call("bpp_engine_set_select_cap", self._h, int(k))

    def unfinished(self):
        n = C.c_int32(0)
        call("bpp_engine_unfinished", self._h, C.byref(n))
        return int(n.value)

    def last_values(self):
        out = torch.empty(self.G, dtype=torch.float64, device=self.device)
        call("bpp_engine_last_values", self._h, _ptr(out), _stream())
        return out

    def begin_move(self):
        call("bpp_engine_begin_move", self._h, _stream())

    # -- lockstep search with an external evaluator ------------------------------------------------------------------
    def select(self):
        call("bpp_engine_select", self._h, _stream())

    def leaf_count(self):
        n = C.c_int32(0)
        call("bpp_engine_leaf_count", self._h, C.byref(n), _stream())
        return int(n.value)

    def leaf_count_async(self, pinned2):
        """queue a copy of [parked leaves, capped games] into a pinned int32[2] tensor (no synchronisation)"""
        call("bpp_engine_leaf_count_async", self._h, C.c_void_p(pinned2.data_ptr()), _stream())

    def leaf_planes(self, count):
        out = torch.empty((count, self.N + 1, self.H, self.W), dtype=torch.float32, device=self.device)
        if count:
            call("bpp_engine_leaf_planes", self._h, _ptr(out), _stream())
        return out

    def leaf_buffers(self):
        c, g, r = C.c_void_p(), C.c_void_p(), C.c_void_p()
        call("bpp_engine_leaf_buffers", self._h, C.byref(c), C.byref(g), C.byref(r))
        return c, g, r

    def expand_backup(self, policy, value):
        policy = policy.contiguous()
        value = value.contiguous()
        self._keep_eval = (policy, value)
        call("bpp_engine_expand_backup", self._h, _ptr(policy), _TORCH_DT[policy.dtype], _ptr(value),
             _TORCH_DT[value.dtype], _stream())

    def expand_select(self, policy, value):
        """expand_backup + select in one launch (bpp_engine_expand_select)"""
        policy = policy.contiguous()
        value = value.contiguous()
        self._keep_eval = (policy, value)
        call("bpp_engine_expand_select", self._h, _ptr(policy), _TORCH_DT[policy.dtype], _ptr(value),
             _TORCH_DT[value.dtype], _stream())

    def search_with(self, evaluator):
        """Run the numMCTSSims simulations of this move for every game with a batched evaluator:
        evaluator(planes float32 (B, N+1, H, W) on device) -> (policy (B, A) f32|f64, value (B,) f32|f64)."""
        steps = 0
        while True:
            self.select()
            n = self.leaf_count()
            if n == 0:  # nothing parked: all simulations of this move are done (or every game has ended)
                break
            pol, val = evaluator(self.leaf_planes(n))
            self.expand_backup(pol, val.reshape(-1))
            steps += 1
        return steps

    # -- fused search with an in-kernel stub evaluator ------------------------------------------------------------------
    def search_stub(self, kind):
        call("bpp_engine_search_stub", self._h, STUB[kind] if isinstance(kind, str) else int(kind), _stream())

    def play_stub(self, kind, choose_mode=_lib.CHOOSE_ARGMAX_FIRST, seed=0, record=True):
        """whole episodes for all games; returns (counts (N, G, A) int32, actions (N, G) int32) when record."""
        counts = actions = None
        if record:
            counts = torch.zeros((self.N, self.G, self.A), dtype=torch.int32, device=self.device)
            actions = torch.full((self.N, self.G), -1, dtype=torch.int32, device=self.device)
        moves = C.c_int32(0)
        call("bpp_engine_play_stub", self._h, STUB[kind] if isinstance(kind, str) else int(kind), int(choose_mode),
             C.c_uint64(seed), 0, _ptr(counts), _ptr(actions), C.byref(moves), _stream())
        return counts, actions

    def play_stub_host(self, kind, items_wh, total_area, bl, choose_mode=_lib.CHOOSE_ARGMAX_FIRST, seed=0, out=None):
        """HOST-buffer episode batch through the C ABI (uploads, plays, downloads, synchronises).
        `out` may hold preallocated (ideally pinned) numpy arrays: counts (N,G,A) i32, actions (N,G) i32,
        r (G,) i32, score (G,) f64, moves (G,) i32."""
        items = np.ascontiguousarray(items_wh, dtype=np.int32)
        area = np.ascontiguousarray(total_area, dtype=np.int32)
        blh = np.ascontiguousarray(bl, dtype=np.float64)
        if out is None:
            out = {}
        out.setdefault("counts", np.empty((self.N, self.G, self.A), dtype=np.int32))
        out.setdefault("actions", np.empty((self.N, self.G), dtype=np.int32))
        out.setdefault("r", np.empty(self.G, dtype=np.int32))
        out.setdefault("score", np.empty(self.G, dtype=np.float64))
        out.setdefault("moves", np.empty(self.G, dtype=np.int32))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        call("bpp_engine_play_stub_host", self._h, STUB[kind] if isinstance(kind, str) else int(kind), int(choose_mode),
             C.c_uint64(seed), vp(items), vp(area), vp(blh), C.c_void_p(0), vp(out["counts"]), vp(out["actions"]),
             vp(out["r"]), vp(out["score"]), vp(out["moves"]), _stream())
        return out

    def play_stub_stream(self, kind, items_wh, total_area, bl, tie=None, choose_mode=_lib.CHOOSE_ARGMAX_FIRST, seed=0,
                         record=True):
        """E >= 1 episodes streamed through the G resident games with a stub evaluator, one launch
        (bpp_engine_play_stub_stream): a game whose episode ends takes the next instance of the queue inside the episode
        kernel.  items_wh (E, N, 2), total_area (E,), bl (E,) float64 (NaN = empty rewards list), tie (E,) int8 or None.
        Returns device tensors indexed by EPISODE: counts (N, E, A), actions (N, E) (None unless record), r, score, moves."""
        items = _dev(items_wh, torch.int32, self.device)
        E = items.shape[0]
        items = items.reshape(E, self.N, 2)
        area = _dev(total_area, torch.int32, self.device).reshape(E)
        blt = _dev(bl, torch.float64, self.device).reshape(E)
        tiet = _dev(tie, torch.int8, self.device).reshape(E) if tie is not None else None
        counts = actions = None
        if record:
            counts = torch.empty((self.N, E, self.A), dtype=torch.int32, device=self.device)
            actions = torch.empty((self.N, E), dtype=torch.int32, device=self.device)
        r = torch.zeros(E, dtype=torch.int32, device=self.device)
        score = torch.zeros(E, dtype=torch.float64, device=self.device)
        moves = torch.zeros(E, dtype=torch.int32, device=self.device)
        self.items_wh = items[:self.G] if E >= self.G else items
        call("bpp_engine_play_stub_stream", self._h, STUB[kind] if isinstance(kind, str) else int(kind), int(choose_mode),
             C.c_uint64(seed), int(E), _ptr(items), _ptr(area), _ptr(blt), _ptr(tiet), _ptr(counts), _ptr(actions), _ptr(r),
             _ptr(score), _ptr(moves), _stream())
        return {"counts": counts, "actions": actions, "r": r, "score": score, "moves": moves}

    def play_stub_stream_host(self, kind, items_wh, total_area, bl, tie=None, choose_mode=_lib.CHOOSE_ARGMAX_FIRST, seed=0,
                              out=None):
        """the same from / to HOST buffers through the C ABI (uploads the queue, plays, downloads, synchronises).
        `out` may hold preallocated (ideally pinned) numpy arrays: counts (N,E,A) i32, actions (N,E) i32, r (E,) i32,
        score (E,) f64, moves (E,) i32."""
        items = np.ascontiguousarray(items_wh, dtype=np.int32)
        E = items.shape[0]
        area = np.ascontiguousarray(total_area, dtype=np.int32)
        blh = np.ascontiguousarray(bl, dtype=np.float64)
        tieh = np.ascontiguousarray(tie, dtype=np.int8) if tie is not None else None
        if out is None:
            out = {}
        out.setdefault("counts", np.empty((self.N, E, self.A), dtype=np.int32))
        out.setdefault("actions", np.empty((self.N, E), dtype=np.int32))
        out.setdefault("r", np.empty(E, dtype=np.int32))
        out.setdefault("score", np.empty(E, dtype=np.float64))
        out.setdefault("moves", np.empty(E, dtype=np.int32))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        call("bpp_engine_play_stub_stream_host", self._h, STUB[kind] if isinstance(kind, str) else int(kind),
             int(choose_mode), C.c_uint64(seed), int(E), vp(items), vp(area), vp(blh),
             vp(tieh) if tieh is not None else C.c_void_p(0), vp(out["counts"]), vp(out["actions"]), vp(out["r"]),
             vp(out["score"]), vp(out["moves"]), _stream())
        return out

    # -- asynchronous whole episodes with the batched device evaluator ---------------------------------------------------
    def set_auto_play(self, mode, seed=0, counts=None, actions=None, roots=None):
        """arm (mode = CHOOSE_*) / disarm (mode = -1) per-game move completion inside expand_select"""
        self._keep_auto = (counts, actions, roots)
        call("bpp_engine_set_auto_play", self._h, int(mode), C.c_uint64(seed), _ptr(counts), _ptr(actions), _ptr(roots))

    def progress_async(self, pinned4):
        """queue a copy of [parked leaves, capped games, games still playing, 0] into a pinned int32[4] tensor"""
        call("bpp_engine_progress_async", self._h, C.c_void_p(pinned4.data_ptr()), _stream())

    def set_profile(self, on):
        call("bpp_engine_set_profile", self._h, int(bool(on)))

    def profile(self):
        arr = (C.c_double * 4)()
        call("bpp_engine_profile", self._h, arr)
        return {"evaluator_ms": arr[0], "expand_select_ms": arr[1], "steps": int(arr[2])}

    def play_net(self, dnet, choose_mode=_lib.CHOOSE_SAMPLE, seed=0, record=True):
        """Whole episodes for all games after reset(), leaves evaluated by `dnet` (a DeviceNet): one C call, no per-move
        host round trip.  Returns a dict of device tensors: roots (N, G, 32) int32, counts (N, G, A) int32, actions
        (N, G) int32 (rows of moves a game did not play: 0 / 0 / -1) and `steps` (lockstep steps queued)."""
        roots = counts = actions = None
        if record:
            roots = torch.empty((self.N, self.G, REC_WORDS), dtype=torch.int32, device=self.device)
            counts = torch.empty((self.N, self.G, self.A), dtype=torch.int32, device=self.device)
            actions = torch.empty((self.N, self.G), dtype=torch.int32, device=self.device)
        steps = C.c_int32(0)
        call("bpp_engine_play_net", self._h, dnet._h, int(choose_mode), C.c_uint64(seed), _ptr(counts), _ptr(actions),
             _ptr(roots), C.byref(steps), _stream())
        return {"roots": roots, "counts": counts, "actions": actions, "steps": int(steps.value)}

    def play_net_stream(self, dnet, items_wh, total_area, bl, tie=None, choose_mode=_lib.CHOOSE_SAMPLE, seed=0,
                        record=True):
        """E >= 1 episodes streamed through the G resident games (bpp_engine_play_net_stream): a game whose episode ends
        takes the next instance of the queue inside the search kernel, so no batch waits for its slowest episode.
        items_wh (E, N, 2), total_area (E,), bl (E,) float64 (NaN = empty rewards list), tie (E,) int8 or None.
        Returns device tensors indexed by EPISODE: roots (N, E, 32), counts (N, E, A), actions (N, E) (None unless
        record), r (E,), score (E,), moves (E,), and `steps`."""
        items = _dev(items_wh, torch.int32, self.device)
        E = items.shape[0]
        items = items.reshape(E, self.N, 2)
        area = _dev(total_area, torch.int32, self.device).reshape(E)
        blt = _dev(bl, torch.float64, self.device).reshape(E)
        tiet = _dev(tie, torch.int8, self.device).reshape(E) if tie is not None else None
        roots = counts = actions = None
        if record:
            roots = torch.empty((self.N, E, REC_WORDS), dtype=torch.int32, device=self.device)
            counts = torch.empty((self.N, E, self.A), dtype=torch.int32, device=self.device)
            actions = torch.empty((self.N, E), dtype=torch.int32, device=self.device)
        r = torch.zeros(E, dtype=torch.int32, device=self.device)
        score = torch.zeros(E, dtype=torch.float64, device=self.device)
        moves = torch.zeros(E, dtype=torch.int32, device=self.device)
        steps = C.c_int32(0)
        self.items_wh = items[:self.G] if E >= self.G else items
        call("bpp_engine_play_net_stream", self._h, dnet._h, int(choose_mode), C.c_uint64(seed), int(E), _ptr(items),
             _ptr(area), _ptr(blt), _ptr(tiet), _ptr(counts), _ptr(actions), _ptr(roots), _ptr(r), _ptr(score),
             _ptr(moves), C.byref(steps), _stream())
        return {"roots": roots, "counts": counts, "actions": actions, "r": r, "score": score, "moves": moves,
                "steps": int(steps.value)}

    def play_net_stream_host(self, dnet, items_wh, total_area, bl, tie=None, choose_mode=_lib.CHOOSE_SAMPLE, seed=0,
                             out=None):
        """the same from / to HOST buffers through the C ABI (uploads the queue, plays, downloads, synchronises)"""
        items = np.ascontiguousarray(items_wh, dtype=np.int32)
        E = items.shape[0]
        area = np.ascontiguousarray(total_area, dtype=np.int32)
        blh = np.ascontiguousarray(bl, dtype=np.float64)
        tieh = np.ascontiguousarray(tie, dtype=np.int8) if tie is not None else None
        if out is None:
            out = {}
        out.setdefault("roots", np.empty((self.N, E, REC_WORDS), dtype=np.uint32))
        out.setdefault("counts", np.empty((self.N, E, self.A), dtype=np.int32))
        out.setdefault("actions", np.empty((self.N, E), dtype=np.int32))
        out.setdefault("r", np.empty(E, dtype=np.int32))
        out.setdefault("score", np.empty(E, dtype=np.float64))
        out.setdefault("moves", np.empty(E, dtype=np.int32))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        steps = C.c_int32(0)
        call("bpp_engine_play_net_stream_host", self._h, dnet._h, int(choose_mode), C.c_uint64(seed), int(E), vp(items),
             vp(area), vp(blh), vp(tieh) if tieh is not None else C.c_void_p(0), vp(out["roots"]), vp(out["counts"]),
             vp(out["actions"]), vp(out["r"]), vp(out["score"]), vp(out["moves"]), C.byref(steps), _stream())
        out["steps"] = int(steps.value)
        return out

    def play_net_host(self, dnet, items_wh, total_area, bl, tie=None, choose_mode=_lib.CHOOSE_SAMPLE, seed=0, out=None):
        """HOST-buffer episode batch with the real net through the C ABI (uploads, plays, downloads, synchronises).
        `out` may hold preallocated (ideally pinned) numpy arrays: roots (N,G,32) u32, counts (N,G,A) i32, actions
        (N,G) i32, r (G,) i32, score (G,) f64, moves (G,) i32."""
        items = np.ascontiguousarray(items_wh, dtype=np.int32)
        area = np.ascontiguousarray(total_area, dtype=np.int32)
        blh = np.ascontiguousarray(bl, dtype=np.float64)
        tieh = np.ascontiguousarray(tie, dtype=np.int8) if tie is not None else None
        if out is None:
            out = {}
        out.setdefault("roots", np.empty((self.N, self.G, REC_WORDS), dtype=np.uint32))
        out.setdefault("counts", np.empty((self.N, self.G, self.A), dtype=np.int32))
        out.setdefault("actions", np.empty((self.N, self.G), dtype=np.int32))
        out.setdefault("r", np.empty(self.G, dtype=np.int32))
        out.setdefault("score", np.empty(self.G, dtype=np.float64))
        out.setdefault("moves", np.empty(self.G, dtype=np.int32))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        steps = C.c_int32(0)
        call("bpp_engine_play_net_host", self._h, dnet._h, int(choose_mode), C.c_uint64(seed), vp(items), vp(area),
             vp(blh), vp(tieh) if tieh is not None else C.c_void_p(0), vp(out["roots"]), vp(out["counts"]),
             vp(out["actions"]), vp(out["r"]), vp(out["score"]), vp(out["moves"]), C.byref(steps), _stream())
        out["steps"] = int(steps.value)
        return out

    # -- results ------------------------------------------------------------------------------------------------------
    def root_counts(self):
        out = torch.empty((self.G, self.A), dtype=torch.int32, device=self.device)
        call("bpp_engine_root_counts", self._h, _ptr(out), _stream())
        return out

    def root_counts_host(self, out=None):
        if out is None:
            out = np.empty((self.G, self.A), dtype=np.int32)
        call("bpp_engine_root_counts_host", self._h, out.ctypes.data_as(C.c_void_p), _stream())
        return out

    def choose(self, mode=_lib.CHOOSE_ARGMAX_FIRST, seed=0):
        out = torch.empty(self.G, dtype=torch.int32, device=self.device)
        call("bpp_engine_choose", self._h, int(mode), C.c_uint64(seed), _ptr(out), _stream())
        return out

    def advance(self, actions):
        act = _dev(actions, torch.int32, self.device).reshape(self.G)
        self._keep_act = act
        call("bpp_engine_advance", self._h, _ptr(act), _stream())

    def status(self):
        done = torch.empty(self.G, dtype=torch.int32, device=self.device)
        r = torch.empty(self.G, dtype=torch.int32, device=self.device)
        score = torch.empty(self.G, dtype=torch.float64, device=self.device)
        moves = torch.empty(self.G, dtype=torch.int32, device=self.device)
        call("bpp_engine_status", self._h, _ptr(done), _ptr(r), _ptr(score), _ptr(moves), _stream())
        return {"done": done, "r": r, "score": score, "moves": moves}

    def roots(self):
        out = torch.empty((self.G, REC_WORDS), dtype=torch.int32, device=self.device)
        call("bpp_engine_roots", self._h, _ptr(out), _stream())
        return out

    def graph_sizes(self):
        nodes = torch.empty(self.G, dtype=torch.int32, device=self.device)
        units = torch.empty(self.G, dtype=torch.int32, device=self.device)
        call("bpp_engine_graph_sizes", self._h, _ptr(nodes), _ptr(units), _stream())
        return nodes, units

    def export_game(self, g=0):
        """search graph of game g as host arrays: (nodes uint32 (n, 32), edges uint64 (units,))"""
        nn, nu = C.c_int32(0), C.c_int64(0)
        call("bpp_engine_export_game", self._h, int(g), C.c_void_p(0), 0, C.c_void_p(0), 0, C.byref(nn), C.byref(nu),
             _stream())
        nodes = np.zeros((max(1, nn.value), REC_WORDS), dtype=np.uint32)
        edges = np.zeros(max(1, nu.value), dtype=np.uint64)
        call("bpp_engine_export_game", self._h, int(g), nodes.ctypes.data_as(C.c_void_p), nodes.shape[0],
             edges.ctypes.data_as(C.c_void_p), edges.shape[0], C.byref(nn), C.byref(nu), _stream())
        return nodes[:nn.value], edges[:nu.value]

    def export_dicts(self, g, items_wh, dtype=np.int64):
        """The reference's six dicts (MCTS_bpp.py:16-26) for game g, keyed by the reference's state bytes
        (stringRepresentation of the (N+1, H, W) tensor): Qsa, Nsa, Ns, Ps, Es, Vs."""
        nodes, edges = self.export_game(g)
        states = unpack_states(nodes, np.repeat(np.asarray(items_wh)[None], len(nodes), axis=0), self.W, self.H, self.N,
                               dtype=dtype) if len(nodes) else []
        Qsa, Nsa, Ns, Ps, Es, Vs = {}, {}, {}, {}, {}, {}
        for rec, st in zip(nodes, states):
            key = st.tobytes()
            kind = (int(rec[31]) >> 16) & 0xFF
            if kind == 0:
                continue  # key known only (created by a real move, not yet visited by search)
            Es[key] = {1: 0, 2: 1, 3: -1}[kind]
            if kind != 1:
                continue
            nv = int(rec[31]) & 0xFFFF
            nvp = (nv + 3) & ~3
            off = int(rec[30])
            blk = edges[off:off + 3 * nvp + nvp // 4]
            Q = blk[:nvp].view(np.float64)
            Pv = blk[nvp:2 * nvp].view(np.float64)
            NC = blk[2 * nvp:3 * nvp].view(np.int32).reshape(nvp, 2)
            act = blk[3 * nvp:].view(np.uint16)[:nv]
            p = np.zeros(self.A, dtype=np.float64)
            v = np.zeros(self.A, dtype=np.int64)
            p[act] = Pv[:nv]
            v[act] = 1
            Ps[key], Vs[key], Ns[key] = p, v, int(rec[29])
            for e in range(nv):
                if NC[e, 0] > 0:
                    Qsa[(key, int(act[e]))] = float(Q[e])
                    Nsa[(key, int(act[e]))] = int(NC[e, 0])
        return Qsa, Nsa, Ns, Ps, Es, Vs

    def stats(self, reset=False):
        arr = (C.c_uint64 * 8)()
        call("bpp_engine_stats", self._h, arr, int(reset), _stream())
        keys = ["sims", "edges", "expansions", "terminals", "nodes_created", "probes", "launches", "edge_units_read"]
        return {k: int(v) for k, v in zip(keys, arr)}

    def check(self):
        call("bpp_engine_check", self._h, _stream())


def algorithmic_bytes_per_sim(W, H, N, edges_per_sim, expansions_per_sim):
    """algorithmic HBM bytes of one simulation with dense A-wide rows.
    K = compact key bytes (row masks + remaining mask), A = W*N."""
    A = W * N
    row_bytes = 2 if W <= 16 else 4
    K = (H * row_bytes + 2 + 3) // 4 * 4
    select = A * (4 + 8 + 8) + math.ceil(A / 8) + K + 4
    lookup = K + 12
    backup = 32
    expand = A * (8 + 12) + math.ceil(A / 8) + K + 12 + 4 * A + 4
    return edges_per_sim * (select + lookup + backup) + expansions_per_sim * expand
