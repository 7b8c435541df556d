/*
 * bpp_b200.h — C ABI of the B200-native self-play hot path (bin-packing MCTS).
 *
 * The reference (Wang-Xiaoyang/resource_packing_self_play) is pure Python and has no FFI; its plug-in boundary is
 * the duck-typed Game / MCTS / NeuralNet API (xw_mcts/Game.py:14-113, xw_mcts/NeuralNet.py:14-50) as specialised by
 *   xw_mcts/binpacking/BinPackingGame.py:8-285, xw_mcts/binpacking/BinPackingLogic.py:19-109,
 *   xw_mcts/MCTS_bpp.py:11-139, xw_mcts/binpacking/pytorch/NNet.py:69-85.
 * Each entry point below names the reference method(s) it replaces.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success and a negative BPP_E_* code on failure; bpp_last_error() returns a
 *     thread-local, NUL-terminated description of the last failure;
 *   - no C++ types, exceptions or torch types cross this boundary: plain pointers, sizes and scalars;
 *   - pointer arguments named *_dev are CALLER-OWNED DEVICE pointers (e.g. torch.Tensor.data_ptr()); pointer arguments
 *     named *_host are host pointers (copied with cudaMemcpyAsync on `stream`; pass pinned memory for overlap);
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream); 0 = legacy default stream;
 *   - a handle is bound to one CUDA device and is NOT thread-safe; use one handle per device/process.
 *
 * Compact state layout ("record"): 32 little-endian uint32 words per state,
 *   word r (0 <= r < H)  occupancy of bin row r, bit x = column x        (plane 0 of the reference state tensor)
 *   word 28              remaining-items mask, bit i = item i not yet placed (planes 1..N non-zero)
 *   words 29..31         tree bookkeeping inside the engine; ignored on input, zero on output of the env ops
 * which, together with the per-episode item list (w_i, h_i), is bijective with the reference's (N+1, H, W) int64
 * state tensor (BinPackingGame.py:118-120).  Limits: W <= 32, H <= 28, N <= 16.
 */
#ifndef BPP_B200_H
#define BPP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPP_REC_WORDS 32
#define BPP_REC_REM 28
#define BPP_MAX_ITEMS 16

#define BPP_OK 0
#define BPP_E_INVALID (-1)  /* bad argument / configuration */
#define BPP_E_CUDA (-2)     /* CUDA runtime error */
#define BPP_E_NOMEM (-3)    /* device allocation failed */
#define BPP_E_CAPACITY (-4) /* a game's node/edge pool overflowed (sticky, see bpp_engine_check) */
#define BPP_E_STATE (-5)    /* call sequence error */

/* stub leaf evaluators computed inside the kernels (definitions: tests/golden/make_golden.py) */
#define BPP_STUB_U 1
#define BPP_STUB_V 2
#define BPP_STUB_H 3
#define BPP_STUB_D 4

#define BPP_DTYPE_F32 0
#define BPP_DTYPE_F64 1

/* action choice after a search (bpp_engine_choose) */
#define BPP_CHOOSE_ARGMAX_FIRST 0 /* first maximum of the visit counts (deterministic; parity tests) */
#define BPP_CHOOSE_SAMPLE 1       /* a ~ counts / sum(counts)   (CoachBPP.py:86-87, greedy=False) */
#define BPP_CHOOSE_GREEDY 2       /* uniformly random arg-max     (MCTS_bpp.py:43-49, greedy_a=0) */

const char *bpp_last_error(void);
/* library/ABI version: major*10000 + minor*100 + patch */
int bpp_version(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Stateless, batched environment ops (one warp per state).
 * n states; recs_dev: uint32 [n][32]; items_wh_dev: int32 [n][N][2] = (w, h) of every item of that state's episode.
 * ---------------------------------------------------------------------------------------------------------------- */

/* BinPackingGame.getValidMoves (BinPackingGame.py:78-92) + Bin.get_moves_for_square / get_adjacency
 * (BinPackingLogic.py:47-93).  valid_out_dev: uint8 [n][W*N], 1 = legal.  Unlike the reference it does not assert
 * on "no legal move": the row is all-zero (that is has_valid_moves == False, BinPackingGame.py:94-107). */
int bpp_env_valid_moves(int W, int H, int N, int n, const uint32_t *recs_dev, const int32_t *items_wh_dev,
                        uint8_t *valid_out_dev, void *stream);

/* ItemsGenerator.items_generator (BinPackingGame.py:257-285) for n seeds at once, on the device: replays numpy's legacy
 * global RNG (np.random.seed(seed); np.random.randint draws) so the instances equal the reference's for the same
 * seeds.  seeds_dev int64 [n] (0 <= seed < 2^32), heights_dev int32 [n] generator bin height (the generator's width is
 * W; a rectangle W x height must be splittable into N items, i.e. W*height >= N); items_wh_out_dev int32 [n][N][2] =
 * (w, h); rects_out_dev (may be NULL) int32 [n][N][4] = the reference's [w, h, a, b]. */
int bpp_items_generate(int W, int N, int n, const int64_t *seeds_dev, const int32_t *heights_dev,
                       int32_t *items_wh_out_dev, int32_t *rects_out_dev, void *stream);

/* BinPackingGame.getBinItem (BinPackingGame.py:118-120) for compact states: the dense evaluator / learner input
 * planes_out_dev float32 [n][N+1][H][W] (plane 0 = bin occupancy, plane i+1 = item i's [0:h, 0:w] block while it is
 * still to be placed). */
int bpp_env_planes(int W, int H, int N, int n, const uint32_t *recs_dev, const int32_t *items_wh_dev,
                   float *planes_out_dev, void *stream);

/* BinPackingGame.getNextState (BinPackingGame.py:58-76) + Bin.execute_move (BinPackingLogic.py:95-109), including
 * the silent truncation when fewer than h strip rows are empty.  actions_dev: int32 [n] (item*W + x; the item must
 * be remaining).  recs_out_dev: uint32 [n][32]. */
int bpp_env_next_state(int W, int H, int N, int n, const uint32_t *recs_dev, const int32_t *items_wh_dev,
                       const int32_t *actions_dev, uint32_t *recs_out_dev, void *stream);

/* BinPackingGame.getGameEnded + getRankedReward + get_minimal_bin_height (BinPackingGame.py:109-116,181-212).
 * total_area_dev, max_h_dev: int32 [n]; bl_dev: float64 [n] ranked-reward threshold
 * sorted(rewards)[floor(len*alpha)-1], NaN = empty rewards list (always +1); tie_dev: int8 [n] value returned on
 * r == bl (the reference draws it from numpy's global RNG, BinPackingGame.py:212), may be NULL (=+1).
 * ended_out_dev: int32 [n] in {0,+1,-1}; score_out_dev: float64 [n] raw reward r (undefined where ended == 0). */
int bpp_env_game_ended(int W, int H, int N, int n, const uint32_t *recs_dev, const int32_t *items_wh_dev,
                       const int32_t *total_area_dev, const int32_t *max_h_dev, const double *bl_dev,
                       const int8_t *tie_dev, int32_t *ended_out_dev, double *score_out_dev, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Search engine: G lockstep games, each with its own device-resident search graph (replaces the six dicts of
 * MCTS.__init__, MCTS_bpp.py:16-26).
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct bpp_engine bpp_engine;

typedef struct bpp_config {
    int32_t W, H, N;      /* virtual bin width/height, items per episode */
    int32_t G;            /* games resident on this device */
    int32_t num_sims;     /* args.numMCTSSims */
    double cpuct;         /* args.cpuct */
    int32_t node_cap;     /* nodes per game; 0 = num_sims*N + 2 (the per-episode worst case) */
    int64_t edge_cap;     /* 8-byte edge units per game; 0 = auto: a quarter of the worst case (3x the measured maximum),
                             clipped to half of the free device memory; overflow = sticky BPP_E_CAPACITY, re-create larger */
    int32_t device;       /* CUDA device ordinal */
} bpp_config;

int bpp_engine_create(const bpp_config *cfg, bpp_engine **out);
int bpp_engine_destroy(bpp_engine *e);
/* bytes of device memory held by the handle */
int64_t bpp_engine_device_bytes(const bpp_engine *e);
/* edge-pool capacity per game actually allocated (8-byte units), see bpp_config.edge_cap */
int bpp_engine_edge_cap(const bpp_engine *e, int64_t *units_out);

/* New episode for every game (a fresh MCTS object, CoachBPP.py:124, plus getInitBoard/getInitItems,
 * BinPackingGame.py:24-51).  items_wh: int32 [G][N][2]; total_area: int32 [G]; bl: float64 [G] (NaN = empty
 * rewards list); tie: int8 [G] or NULL.  max_h (hidden state of the reference Game, BinPackingGame.py:48-50) is
 * derived from the items. */
int bpp_engine_reset(bpp_engine *e, const int32_t *items_wh_dev, const int32_t *total_area_dev, const double *bl_dev,
                     const int8_t *tie_dev, void *stream);
int bpp_engine_reset_host(bpp_engine *e, const int32_t *items_wh_host, const int32_t *total_area_host,
                          const double *bl_host, const int8_t *tie_host, void *stream);

/* Overwrite the root state of every game (drop-in single-state use: MCTS.getActionProb(state, ...) on an arbitrary
 * state).  roots_dev: uint32 [G][32] records.  Keeps the search graph. */
int bpp_engine_set_roots(bpp_engine *e, const uint32_t *roots_dev, void *stream);

/* Override max_h, the hidden state BinPackingGame.getInitItems leaves on the Game object (BinPackingGame.py:48-50) and
 * getRankedReward reads (:198); needed when a search starts from a mid-episode state whose placed items' dims are no
 * longer in the state tensor.  max_h_dev: int32 [G].  Call after bpp_engine_reset. */
int bpp_engine_set_max_h(bpp_engine *e, const int32_t *max_h_dev, void *stream);
/* Change the per-move simulation budget (args.numMCTSSims); MCTS.search() is one simulation: budget 1. */
int bpp_engine_set_num_sims(bpp_engine *e, int num_sims);
/* Lockstep mode: cap the simulations one game runs inside one bpp_engine_select launch (0 = no cap, the default).  A game
 * whose simulations keep ending on terminal states parks no leaf; with a cap it continues in the next step instead of
 * delaying the leaf batch of every other game.  The order of a game's simulations, hence every result, is unchanged. */
int bpp_engine_set_select_cap(bpp_engine *e, int max_sims_per_launch);
/* games that stopped at the cap in the last select (valid after bpp_engine_leaf_count); the move is finished when both
 * the leaf count and this count are 0 */
int bpp_engine_unfinished(bpp_engine *e, int32_t *count_host);
/* v returned by the most recent simulation of each game (the return value of MCTS.search, MCTS_bpp.py:83,104,139).
 * values_out_dev: float64 [G]. */
int bpp_engine_last_values(bpp_engine *e, double *values_out_dev, void *stream);

/* Begin a getActionProb call: zero the per-move simulation counters (MCTS_bpp.py:37). */
int bpp_engine_begin_move(bpp_engine *e, void *stream);

/* One lockstep step of MCTS.search (MCTS_bpp.py:56-139) for every game that still owes simulations this move:
 * runs simulations (PUCT descent :107-125, terminal hits :78-83 with backup :130-139) until the game reaches an
 * unexpanded non-terminal state, then parks that leaf for evaluation.  After it returns, the leaf batch is
 * (leaf_count, leaf records); feed an evaluator and call bpp_engine_expand_backup. */
int bpp_engine_select(bpp_engine *e, void *stream);
/* number of parked leaves (synchronises `stream`); a count of 0 completes the select/expand pairing, i.e. no
 * bpp_engine_expand_backup call is needed (every game has finished its simulations for this move) */
int bpp_engine_leaf_count(bpp_engine *e, int32_t *count_host, void *stream);
/* the same without synchronising: copies {parked leaves, games stopped by the select cap} into counts_host2 (int32[2],
 * pinned host memory) in stream order; the caller waits on its own event.  Lets a driver queue the next chunk of
 * lockstep steps before it learns whether the previous one finished the move (steps after the end are no-ops). */
int bpp_engine_leaf_count_async(bpp_engine *e, int32_t *counts_host2, void *stream);
/* device-side leaf batch for a device evaluator: count int32[1], game index int32 [<=G], records uint32 [<=G][32]
 * (row b = leaf b).  Pointers stay valid for the life of the handle. */
int bpp_engine_leaf_buffers(bpp_engine *e, const int32_t **count_dev, const int32_t **game_dev,
                            const uint32_t **recs_dev);
/* Dense evaluator input, the reference's NNet.predict argument (getBinItem, BinPackingGame.py:118-120):
 * planes_out_dev float32 [leaf_count][N+1][H][W]. */
int bpp_engine_leaf_planes(bpp_engine *e, float *planes_out_dev, void *stream);
/* Leaf expansion (MCTS_bpp.py:85-104: mask by valid moves, renormalise with numpy's pairwise summation order,
 * uniform fallback) and path backup (:130-139) for every parked leaf.
 * policy_dev: [leaf_count][A] of policy_dtype; value_dev: [leaf_count] of value_dtype. */
int bpp_engine_expand_backup(bpp_engine *e, const void *policy_dev, int policy_dtype, const void *value_dev,
                             int value_dtype, void *stream);

/* bpp_engine_expand_backup followed by bpp_engine_select in ONE launch (one warp per game: the expansion and backup of
 * the game's parked leaf, MCTS_bpp.py:85-104,130-139, then the next descents :107-125 until it parks its next leaf).
 * Results are identical to the two separate calls; games without a parked leaf only select.  The new leaf batch
 * replaces the old one (same buffers, bpp_engine_leaf_buffers). */
int bpp_engine_expand_select(bpp_engine *e, const void *policy_dev, int policy_dtype, const void *value_dev,
                             int value_dtype, void *stream);

/* Whole getActionProb simulation loop (MCTS_bpp.py:37-38) in ONE launch with an in-kernel stub evaluator:
 * every game runs simulations until its per-move counter reaches num_sims. */
int bpp_engine_search_stub(bpp_engine *e, int stub_kind, void *stream);

/* counts[a] = Nsa[(root, a)] (MCTS_bpp.py:40-41).  counts_out_dev: int32 [G][A]. */
int bpp_engine_root_counts(bpp_engine *e, int32_t *counts_out_dev, void *stream);
int bpp_engine_root_counts_host(bpp_engine *e, int32_t *counts_out_host, void *stream);
/* choose an action per game from the root visit counts; actions_out_dev int32 [G] (-1 for finished games) */
int bpp_engine_choose(bpp_engine *e, int mode, uint64_t seed, int32_t *actions_out_dev, void *stream);
/* Play actions_dev[g] in every unfinished game (CoachBPP.py:88-98): root <- getNextState(root, a); evaluates
 * getGameEnded on the new root and latches (r, score) when the episode ends.  Starts the next move. */
int bpp_engine_advance(bpp_engine *e, const int32_t *actions_dev, void *stream);
/* Per-game episode status: done int32 [G] (0 running, 1 ended), r int32 [G] (+1/-1), score float64 [G],
 * moves int32 [G].  Any pointer may be NULL. */
int bpp_engine_status(bpp_engine *e, int32_t *done_out_dev, int32_t *r_out_dev, double *score_out_dev,
                      int32_t *moves_out_dev, void *stream);
/* current root records uint32 [G][32] (word 28 = remaining mask) */
int bpp_engine_roots(bpp_engine *e, uint32_t *roots_out_dev, void *stream);

/* Whole self-play episodes in one call with an in-kernel stub evaluator: reset must have been called; loops
 * search_stub -> choose(mode, seed) -> advance until every game has ended.  counts_out_dev (may be NULL):
 * int32 [max_moves][G][A] per-move visit counts; actions_out_dev (may be NULL): int32 [max_moves][G].
 * moves_run_host receives the number of move rounds executed. */
int bpp_engine_play_stub(bpp_engine *e, int stub_kind, int choose_mode, uint64_t seed, int max_moves,
                         int32_t *counts_out_dev, int32_t *actions_out_dev, int32_t *moves_run_host, void *stream);

/* The reference-facing batched "executeEpisode" with HOST buffers (CoachBPP.py:50-99 for G games at once): uploads the
 * instances, resets, plays whole episodes with the in-kernel stub evaluator, downloads the results and synchronises.
 * counts_out_host int32 [N][G][A] (rows of moves a game did not play stay 0), actions_out_host int32 [N][G] (-1 when
 * the game had already ended), r_out_host int32 [G], score_out_host float64 [G], moves_out_host int32 [G]; any output
 * may be NULL.  Fails with BPP_E_CAPACITY if a game overflowed its pools. */
int bpp_engine_play_stub_host(bpp_engine *e, int stub_kind, int choose_mode, uint64_t seed,
                              const int32_t *items_wh_host, const int32_t *total_area_host, const double *bl_host,
                              const int8_t *tie_host, int32_t *counts_out_host, int32_t *actions_out_host,
                              int32_t *r_out_host, double *score_out_host, int32_t *moves_out_host, void *stream);

/* Episode STREAM with a stub evaluator (CoachBPP.executeEpisode x num_episodes, CoachBPP.py:50-99, in ONE launch):
 * num_episodes >= 1 instances played through the G resident games.  Games 0..G-1 start with episodes 0..G-1; a game whose
 * episode ends latches the outcome and takes the next instance of the queue inside the episode kernel, so the launch has
 * no idle tail but the end of the whole stream.  Inputs are indexed by episode: items_wh int32 [E][N][2], total_area int32
 * [E], bl float64 [E], tie int8 [E] or NULL; outputs too: counts int32 [N][E][A], actions int32 [N][E], r int32 [E], score
 * float64 [E], moves int32 [E]; any output may be NULL.  The action stream is keyed by (seed, episode, move): results do
 * not depend on G, and with num_episodes == G they equal bpp_engine_reset + bpp_engine_play_stub.  No bpp_engine_reset is
 * needed before the call.  The _host form takes / fills HOST buffers (pinned result buffers are written by the kernel
 * itself while the episodes run), synchronises and fails with BPP_E_CAPACITY if a game overflowed its pools. */
int bpp_engine_play_stub_stream(bpp_engine *e, int stub_kind, int choose_mode, uint64_t seed, int num_episodes,
                                const int32_t *items_wh_dev, const int32_t *total_area_dev, const double *bl_dev,
                                const int8_t *tie_dev, int32_t *counts_out_dev, int32_t *actions_out_dev,
                                int32_t *r_out_dev, double *score_out_dev, int32_t *moves_out_dev, void *stream);
int bpp_engine_play_stub_stream_host(bpp_engine *e, int stub_kind, int choose_mode, uint64_t seed, int num_episodes,
                                     const int32_t *items_wh_host, const int32_t *total_area_host, const double *bl_host,
                                     const int8_t *tie_host, int32_t *counts_out_host, int32_t *actions_out_host,
                                     int32_t *r_out_host, double *score_out_host, int32_t *moves_out_host, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Asynchronous self-play episodes with the batched device evaluator (CoachBPP.executeEpisode, CoachBPP.py:50-99, for
 * all G games at once; the real-net counterpart of bpp_engine_play_stub).
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct bpp_net bpp_net;

/* Arm (choose_mode = BPP_CHOOSE_*) or disarm (choose_mode = -1) per-game move completion inside
 * bpp_engine_expand_select: a game that has run its num_sims simulations writes its visit-count row
 * counts_out_dev[m][g][:] (MCTS_bpp.py:40-41), the root record it searched from roots_out_dev[m][g][:] (getBinItem of the
 * example, CoachBPP.py:74-80), chooses with the stream (seed, game, move number) exactly like bpp_engine_choose, stores
 * actions_out_dev[m][g], plays the move like bpp_engine_advance (CoachBPP.py:88-98) and continues with the next move in
 * the same launch.  No game waits for the slowest game of a move; results per game are identical to the per-move
 * sequence select/expand ... root_counts, choose, advance.  counts int32 [N][G][A], actions int32 [N][G], roots uint32
 * [N][G][32]; any may be NULL; caller-zeroed. */
int bpp_engine_set_auto_play(bpp_engine *e, int choose_mode, uint64_t seed, int32_t *counts_out_dev,
                             int32_t *actions_out_dev, uint32_t *roots_out_dev);
/* progress counters of the last select / expand_select launch, copied in stream order into counts_host4 (int32[4],
 * pinned): [0] parked leaves, [1] games stopped by the select cap, [2] games still playing their episode, [3] 0 */
int bpp_engine_progress_async(bpp_engine *e, int32_t *counts_host4, void *stream);
/* Whole episodes for all G games after bpp_engine_reset: loops {bpp_net_forward on the parked leaves ->
 * bpp_engine_expand_select} with auto-play armed until every game has ended; the lockstep steps are queued in chunks and
 * the host only reads the progress counters of the previous chunk.  Output rows of moves a game did not play: counts 0,
 * action -1, root 0.  Synchronises the stream.  steps_run_host (may be NULL) = lockstep steps queued. */
int bpp_engine_play_net(bpp_engine *e, bpp_net *net, int choose_mode, uint64_t seed, int32_t *counts_out_dev,
                        int32_t *actions_out_dev, uint32_t *roots_out_dev, int32_t *steps_run_host, void *stream);
/* Episode STREAM: num_episodes >= 1 instances played through the G resident games.  Games 0..G-1 start with episodes
 * 0..G-1; a game whose episode ends latches the outcome and takes the next instance of the queue inside the same search
 * launch, so the device always holds G running games until the queue is empty (no batch waits for its slowest episode).
 * Inputs are indexed by episode: items_wh int32 [E][N][2], total_area int32 [E], bl float64 [E], tie int8 [E] or NULL.
 * Outputs are indexed by episode too: counts int32 [N][E][A], actions int32 [N][E], roots uint32 [N][E][32], r int32 [E],
 * score float64 [E], moves int32 [E]; any may be NULL.  The action stream is keyed by (seed, episode, move), so results do
 * not depend on which game slot an episode lands in; with num_episodes == G they equal bpp_engine_reset +
 * bpp_engine_play_net.  No bpp_engine_reset is needed before the call. */
int bpp_engine_play_net_stream(bpp_engine *e, bpp_net *net, int choose_mode, uint64_t seed, int num_episodes,
                               const int32_t *items_wh_dev, const int32_t *total_area_dev, const double *bl_dev,
                               const int8_t *tie_dev, int32_t *counts_out_dev, int32_t *actions_out_dev,
                               uint32_t *roots_out_dev, int32_t *r_out_dev, double *score_out_dev, int32_t *moves_out_dev,
                               int32_t *steps_run_host, void *stream);
int bpp_engine_play_net_stream_host(bpp_engine *e, bpp_net *net, int choose_mode, uint64_t seed, int num_episodes,
                                    const int32_t *items_wh_host, const int32_t *total_area_host, const double *bl_host,
                                    const int8_t *tie_host, uint32_t *roots_out_host, int32_t *counts_out_host,
                                    int32_t *actions_out_host, int32_t *r_out_host, double *score_out_host,
                                    int32_t *moves_out_host, int32_t *steps_run_host, void *stream);
/* Accounting pass for bench.py: with on != 0 bpp_engine_play_net records CUDA events around every evaluator call and every
 * expand_select call (slower: use it for kernel shares, not for throughput) and accumulates their durations;
 * bpp_engine_profile returns {evaluator ms, expand+select ms, lockstep steps timed, 0} since the last set_profile. */
int bpp_engine_set_profile(bpp_engine *e, int on);
int bpp_engine_profile(bpp_engine *e, double ms_out4[4]);
/* The same with HOST buffers, the reference-facing batched executeEpisode with the real net: uploads the instances,
 * resets, plays, downloads the compact examples (root records + visit counts + actions per move) and the outcomes.
 * roots_out_host uint32 [N][G][32], counts_out_host int32 [N][G][A], actions_out_host int32 [N][G], r_out_host int32
 * [G], score_out_host float64 [G], moves_out_host int32 [G]; any output may be NULL.  BPP_E_CAPACITY if a game
 * overflowed its pools. */
int bpp_engine_play_net_host(bpp_engine *e, bpp_net *net, int choose_mode, uint64_t seed, const int32_t *items_wh_host,
                             const int32_t *total_area_host, const double *bl_host, const int8_t *tie_host,
                             uint32_t *roots_out_host, int32_t *counts_out_host, int32_t *actions_out_host,
                             int32_t *r_out_host, double *score_out_host, int32_t *moves_out_host,
                             int32_t *steps_run_host, void *stream);

/* Counters since creation (or the last bpp_engine_stats with reset != 0), copied to the host (synchronises):
 * [0] simulations, [1] edges traversed, [2] expansions, [3] terminal hits, [4] nodes created, [5] hash probes,
 * [6] kernels launched by this handle, [7] 8-byte edge-block units read by the PUCT selections (layout traffic). */
int bpp_engine_stats(bpp_engine *e, uint64_t stats_host[8], int reset, void *stream);
/* Synchronises and returns BPP_E_CAPACITY if any game overflowed its pools (the search of that game stopped). */
int bpp_engine_check(bpp_engine *e, void *stream);
/* Dump the search graph of one game to the host (the content of the reference's dicts Qsa/Nsa/Ns/Ps/Es/Vs,
 * MCTS_bpp.py:16-26).  Call with NULL buffers to obtain the sizes.  nodes_out_host: uint32 [n_nodes][32] records (words
 * 0..H-1 rows, 28 remaining mask, 29 Ns, 30 edge-block offset in 8-byte units, 31 = nvalid | kind<<16 with kind 0 = key
 * only, 1 = expanded, 2 = terminal +1, 3 = terminal -1).  edges_out_host: uint64 [n_units]; the block of an expanded
 * node with nv valid actions (nvp = nv rounded up to 4) is Q float64[nvp] | P float64[nvp] | {int32 Nsa, int32 child}[nvp]
 * | uint16 action[nvp].  Synchronises. */
int bpp_engine_export_game(bpp_engine *e, int game, uint32_t *nodes_out_host, int32_t nodes_cap,
                           uint64_t *edges_out_host, int64_t units_cap, int32_t *n_nodes_host, int64_t *n_units_host,
                           void *stream);
/* per-game graph sizes: nodes int32 [G], edge units int32 [G] (either may be NULL) */
int bpp_engine_graph_sizes(bpp_engine *e, int32_t *nodes_out_dev, int32_t *units_out_dev, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Batched policy/value network forward (NNetWrapper.predict, NNet.py:69-85, over BinPackingNNet.forward,
 * BinpackingNNet.py:72-81) in bf16 with fp32 accumulation.  Declared in bpp_net section of the library.
 * ---------------------------------------------------------------------------------------------------------------- */
/* Create for board H x W, N items (in_channels = N + 1), action size A = W*N. */
int bpp_net_create(int W, int H, int N, int max_batch, int device, bpp_net **out);
int bpp_net_destroy(bpp_net *n);
/* Load one parameter tensor by its reference state_dict name (e.g. "conv_seqs.0.res_block1.conv0.weight"),
 * fp32, host pointer, PyTorch layout (OIHW for conv, [out][in] for linear). */
int bpp_net_set_param(bpp_net *n, const char *name, const float *data_host, int64_t numel);
/* commit parameters (converts/re-lays-out to the device formats the kernels use) */
int bpp_net_commit(bpp_net *n, void *stream);
/* Arithmetic of the forward.  BPP_NET_BF16 (default): bf16 weights and inter-layer activations, fp32 accumulation - within
 * 1e-3 of the fp32 reference for freshly initialised networks.  BPP_NET_FP32: fp32 weights and activations - needed for
 * the reference's TRAINED checkpoints, whose logits span ~3e3 so that bf16 rounding of the weights alone moves the
 * policy by up to 0.3 (measured, DESIGN.md "leaf evaluation"); BPP_NET_BF16X3 keeps fp32-level accuracy on the tensor cores
 * and is what the Python wrapper's precision="auto" selects for such weights. */
#define BPP_NET_BF16 0      /* tcgen05 tensor-core kernels (grid-row implicit GEMM, TMEM accumulators; bpp_net_gr.cuh) */
#define BPP_NET_FP32 1      /* CUDA-core kernel, fp32 weights and activations */
#define BPP_NET_BF16_SIMT 2 /* CUDA-core kernel with the bf16 roundings of mode 0 (cross-check of the tensor-core path) */
#define BPP_NET_BF16X3 3    /* tcgen05 kernel in split-bf16: activations and weights as hi + lo bf16 halves, three MMAs per
                               product, fp32 heads (~16 mantissa bits): the tensor-core mode for TRAINED checkpoints */
int bpp_net_set_precision(bpp_net *n, int mode);
/* Phase timers (SM clock cycles, CTA 0 of the last tensor-core forward; synchronises the device):
 * [0] input planes, [1] weight staging, [2] MMA issue, [3] MMA wait, [4] epilogue, [5] pooling, [6] heads, [7] total. */
int bpp_net_profile(bpp_net *n, int64_t cycles_host[8]);
/* Timers of the four level stages of the grid-row trunk (k_net_gr, CTA 0 / thread 0 of the last forward):
 * cycles_host[8 * stage + slot], slots [0] input, [1] layers, [2] output (pooling / features), [3] stage prologue, [4] whole
 * stage, [7] groups only.  (Fallback kernels: the phase timers above per role kernel k_net_role<0..2>.) */
int bpp_net_profile_roles(bpp_net *n, int64_t cycles_host[32]);
/* 1 when the handle's current precision mode (bf16 or split-bf16) runs the grid-row stage kernels (k_net_gr), else 0 */
int bpp_net_grid_row(bpp_net *n);
/* Forward for B compact states.  recs_dev uint32 [B][32], game_dev int32 [B] (index into items_wh_dev rows; may be
 * NULL for identity), items_wh_dev int32 [*][N][2]; if count_dev != NULL the batch size is read from device memory
 * (*count_dev <= B).  policy_out_dev float32 [B][A] = exp(log_softmax(logits)); value_out_dev float32 [B]. */
int bpp_net_forward(bpp_net *n, int B, const int32_t *count_dev, const uint32_t *recs_dev, const int32_t *game_dev,
                    const int32_t *items_wh_dev, float *policy_out_dev, float *value_out_dev, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Learner step (NNetWrapper.train, NNet.py:27-67; losses NNet.py:87-91) in fp32 on the device: forward with stashed
 * activations, loss_pi + loss_v, backward through the whole network, gradients reduced in a fixed order.  Parameters,
 * gradients and Adam moments are flat caller-owned fp32 device buffers in the reference's state_dict order
 * (conv_seqs.{0,1,2}.{conv,res_block0.conv0,res_block0.conv1,res_block1.conv0,res_block1.conv1}.{weight,bias},
 * hidden_fc, logits_fc, value_fc; PyTorch layouts), see bpp_learner_param_offset.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct bpp_learner bpp_learner;
int bpp_learner_create(int W, int H, int N, int max_batch, int device, bpp_learner **out);
int bpp_learner_destroy(bpp_learner *l);
int bpp_learner_num_params(bpp_learner *l, int64_t *out);
int bpp_learner_param_offset(bpp_learner *l, const char *name, int64_t *offset, int64_t *numel);
/* One minibatch of B examples: example b is row ids_dev[b] (int64; NULL = identity) of recs_dev uint32 [*][32],
 * items_wh_dev int32 [*][N][2], pis_dev float32 [*][A], vs_dev float32 [*].  losses_out_dev float32 [2] =
 * {loss_pi, loss_v} (NNet.py:87-91, means over the batch).  grads_out_dev float32 [num_params] = d(loss_pi + loss_v)/dp;
 * NULL = forward and losses only.  Optional outputs: logp_out_dev float32 [B][A] (log_softmax), v_out_dev float32 [B]. */
int bpp_learner_grad(bpp_learner *l, int B, const float *params_dev, const uint32_t *recs_dev,
                     const int32_t *items_wh_dev, const int64_t *ids_dev, const float *pis_dev, const float *vs_dev,
                     float *grads_out_dev, float *losses_out_dev, float *logp_out_dev, float *v_out_dev, void *stream);
/* torch.optim.Adam update (optimizer of NNet.py:31; no weight decay) on flat buffers: g = grads * grad_scale (1/world
 * after a sum all-reduce).  The step count t >= 1 comes from step_dev (int32 in device memory, for CUDA-graph replay)
 * when it is not NULL, else from `step`. */
int bpp_learner_adam(int64_t n, float *params_dev, const float *grads_dev, float *exp_avg_dev, float *exp_avg_sq_dev,
                     int step, const int32_t *step_dev, float grad_scale, float lr, float beta1, float beta2, float eps,
                     void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BPP_B200_H */
