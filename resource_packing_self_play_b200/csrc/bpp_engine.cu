// bpp_engine.cu — device-resident MCTS for G lockstep bin-packing games + the stateless env ops, behind the C ABI of
// include/bpp_b200.h.  sm_100a only.
//
// Layout in HBM (per game g, struct-of-arrays across games):
//   nodes  [G][node_cap][32] u32   one 128-byte record per known state: rows 0..H-1, rem, Ns, edge offset, meta
//   table  [G][table_cap]    u32   open-addressing hash table over the FULL compact state (tag<<20 | node+1)
//   edges  [G][edge_cap]     8 B   bump-allocated edge blocks (Q f64 | P f64 | {Nsa, child} | action u16), only the
//                                  VALID actions of a node are stored, in ascending action order
// This replaces the reference's six Python dicts keyed by 19.8 KB byte strings (MCTS_bpp.py:16-26).  The graph is a
// DAG (transpositions share Ns/Ps, edges keep their own Nsa/Qsa), exactly like the dict version.
//
// Parallelisation: the reference's simulations of one game are strictly sequential (each reads the statistics the
// previous one wrote), so ONE WARP owns one game and all parallelism is across games; inside the warp the 32 lanes
// split the actions (PUCT, valid sweep), the bin rows (placement, hashing, key compare) and the path (backup).
// No atomics are needed on the graph; atomics are only used for the leaf-batch cursor and the statistics.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/bpp_b200.h"
#include "bpp_device.cuh"

using namespace bpp;

// ---------------------------------------------------------------------------------------------------------------------
// error plumbing
static thread_local char g_err[512] = "";
static int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
extern "C" const char* bpp_last_error(void) { return g_err; }
int bpp_set_error_message(int code, const char* msg) { return set_err(code, "%s", msg); }  // used by bpp_net.cu
extern "C" int bpp_version(void) { return 100; }

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return set_err(BPP_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                           __LINE__);                                                                    \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------------
struct Params {
    Geom geom;
    int G, num_sims, node_cap;
    unsigned table_mask;
    long long edge_cap;
    double cpuct;
    SumPlan plan;
    // episode constants
    uint8_t* item_w;   // [G][16]
    uint8_t* item_h;   // [G][16]
    int* total_area;   // [G]
    int* numer;        // [G]
    double* bl;        // [G]
    int8_t* tie;       // [G]
    // per-game state
    uint32_t* root_rec;  // [G][32]
    int* root_node;      // [G]
    int* n_nodes;        // [G]
    int* n_units;        // [G]
    int* sims_done;      // [G]
    int* moves_done;     // [G]
    int* status;         // [G] 0 running, 1 episode ended, <0 error
    int* ep_r;           // [G]
    double* ep_score;    // [G]
    double* last_v;      // [G] value backed up by the most recent simulation (return value of MCTS.search)
    // parked leaves (lockstep)
    int* pend_depth;         // [G] -1 = none
    int* pend_leaf;          // [G]
    int* pend_slot;          // [G] slot of the parked leaf in the leaf batch (row of the evaluator's outputs)
    int4* pend_path;         // [G][32]
    uint32_t* pend_valid;    // [G][16]
    int* leaf_count;         // [1]
    int* leaf_game;          // [G]
    uint32_t* leaf_rec;      // [G][32]
    // graph
    uint32_t* nodes;
    uint32_t* table;
    unsigned long long* edges;
    unsigned long long* stats;  // [8]
    const double* sqrt_tab;     // [3][sqrt_n]: sqrt(n), sqrt(n + 1e-8), 1/n (correctly rounded, host-computed)
    int sqrt_n;
    int select_cap;             // lockstep: simulations per game and select launch (0 = until a leaf is parked)
    int edge_budget;            // lockstep: a game starts no new simulation in a launch once it has walked this many edges
                                // in it (0 = no limit); bounds the launch time in work units instead of simulations
    // asynchronous episodes (bpp_engine_set_auto_play): a game that completes its num_sims simulations inside
    // k_expand_search writes its visit-count row, chooses, plays the move and goes on with the next move in the same launch
    int auto_mode;              // -1 = off, else BPP_CHOOSE_*
    unsigned long long auto_seed;
    int32_t* ep_counts;         // [N][G][A] or nullptr
    int32_t* ep_actions;        // [N][G] or nullptr
    uint32_t* ep_roots;         // [N][G][32] root record each move was searched from, or nullptr
    // episode stream (bpp_engine_play_net_stream): q_total >= 1 episodes through the G resident games; a game whose
    // episode ends takes the next instance of the queue in the same launch.  The per-move outputs above are then indexed
    // [move][episode] (row length q_total) and the outcomes go to the epq_* arrays.
    int q_total;                // 0 = no stream (game g plays episode g only)
    int* q_next;                // [1] next episode to hand out
    const int32_t* q_items;     // [E][N][2]
    const int32_t* q_area;      // [E]
    const double* q_bl;         // [E]
    const int8_t* q_tie;        // [E] or nullptr
    int* slot_ep;               // [G] episode the game is playing
    int32_t* items_i32;         // [G][N][2] item list of the resident games, read by the evaluator
    int32_t* epq_r;             // [E] outcomes per episode (any may be nullptr)
    double* epq_score;
    int32_t* epq_moves;
};

struct __align__(16) WarpSmem {
    uint32_t occ[32];
    uint32_t col[VW_SCRATCH];   // valid_words: range-OR table over the columns + prefix sums of the column counts
    uint32_t vw[MAX_AW];
    uint16_t items[BPP_MAX_ITEMS];
    uint8_t tab[16];
    double scratch[MAX_LEAVES];
};

struct Stats {
    unsigned sims, edges, expansions, terminals, created, probes, units;
};

struct GameCtx {
    int g;
    uint32_t* nodes;
    uint32_t* table;
    unsigned long long* edges;
    int n_nodes, n_units, root_node, err;
    RewardCtx rc;
};

__device__ __forceinline__ void load_ctx(const Params& P, int g, int lane, GameCtx& gm, WarpSmem& sm) {
    gm.g = g;
    gm.nodes = P.nodes + (size_t)g * P.node_cap * REC_WORDS;
    gm.table = P.table + (size_t)g * (P.table_mask + 1u);
    gm.edges = P.edges + (size_t)g * (size_t)P.edge_cap;
    gm.n_nodes = P.n_nodes[g];
    gm.n_units = P.n_units[g];
    gm.root_node = P.root_node[g];
    gm.err = 0;
    gm.rc.total_area = P.total_area[g];
    gm.rc.numer = P.numer[g];
    gm.rc.bl = P.bl[g];
    gm.rc.tie = P.tie[g];
    if (lane < BPP_MAX_ITEMS)
        sm.items[lane] = (uint16_t)(P.item_w[g * BPP_MAX_ITEMS + lane] | (P.item_h[g * BPP_MAX_ITEMS + lane] << 8));
    __syncwarp();
}
__device__ __forceinline__ void store_ctx(const Params& P, const GameCtx& gm, int lane) {
    if (lane == 0) {
        P.n_nodes[gm.g] = gm.n_nodes;
        P.n_units[gm.g] = gm.n_units;
        P.root_node[gm.g] = gm.root_node;
        if (gm.err) P.status[gm.g] = -gm.err;
    }
}
__device__ __forceinline__ void flush_stats(const Params& P, const Stats& st, int lane) {
    if (lane == 0) {
        if (st.sims) atomicAdd(&P.stats[0], (unsigned long long)st.sims);
        if (st.edges) atomicAdd(&P.stats[1], (unsigned long long)st.edges);
        if (st.expansions) atomicAdd(&P.stats[2], (unsigned long long)st.expansions);
        if (st.terminals) atomicAdd(&P.stats[3], (unsigned long long)st.terminals);
        if (st.created) atomicAdd(&P.stats[4], (unsigned long long)st.created);
        if (st.probes) atomicAdd(&P.stats[5], (unsigned long long)st.probes);
        if (st.units) atomicAdd(&P.stats[7], (unsigned long long)st.units);
    }
}

// Find the node of a state (full-key equality, like the dict lookup of MCTS_bpp.py:76-85) or create it.
// (scalars instead of `const Params&`: a reference into kernel-parameter space from a non-inlined function would force
// a 480-byte local copy of Params in every kernel)
// The context fields travel by value and the counters come back packed (bits 0..31 node index or -1, bits 32..47 probes,
// bit 48 created): a reference into the caller's GameCtx / Stats would pin both structures in local memory for the whole
// search loop (LDL / STL on the critical path of every level).
__device__ __noinline__ long long lookup_or_insert_v(int H, int node_cap, uint32_t table_mask, uint32_t* table, uint32_t* nodes,
                                                     int n_nodes, uint32_t rec, int lane) {
    const uint32_t h = hash_state(rec, lane, H);
    const uint32_t tag = h >> 20;
    uint32_t slot = h & table_mask;
    long long probes = 0;
    for (;;) {
        probes++;
        const uint32_t ent = table[slot];
        if (ent == 0u) {
            if (n_nodes >= node_cap) return (probes << 32) | 0xffffffffll;
            const int idx = n_nodes;
            nodes[(size_t)idx * REC_WORDS + lane] = state_lane(lane, H) ? rec : 0u;
            if (lane == 0) table[slot] = (tag << 20) | (uint32_t)(idx + 1);
            __syncwarp();
            return (1ll << 48) | (probes << 32) | (long long)idx;
        }
        if ((ent >> 20) == tag) {
            const int cand = (int)(ent & 0xfffffu) - 1;
            const uint32_t o = nodes[(size_t)cand * REC_WORDS + lane];
            const bool same = !state_lane(lane, H) || o == rec;
            if (__all_sync(FULL, same)) return (probes << 32) | (long long)cand;
        }
        slot = (slot + 1u) & table_mask;
    }
}
__device__ __forceinline__ int lookup_or_insert(int H, int node_cap, uint32_t table_mask, GameCtx& gm, uint32_t rec, int lane,
                                                Stats& st) {
    const long long r = lookup_or_insert_v(H, node_cap, table_mask, gm.table, gm.nodes, gm.n_nodes, rec, lane);
    const int idx = (int)(uint32_t)r;
    const unsigned created = (unsigned)(r >> 48) & 1u;
    st.probes += (unsigned)(r >> 32) & 0xffffu;
    st.created += created;
    gm.n_nodes += (int)created;
    if (idx < 0) gm.err = 4;
    return idx;
}

// Leaf expansion, MCTS_bpp.py:85-104.  sm.vw holds the valid mask.  prior(a) is the evaluator's p[a] as float64.
template <typename F>
__device__ __forceinline__ bool expand_node(const Params& P, GameCtx& gm, WarpSmem& sm, int cur, int lane, F prior) {
    const uint32_t myw = lane < MAX_AW ? sm.vw[lane] : 0u;
    const int nv = (int)__reduce_add_sync(FULL, (unsigned)__popc(myw));
    const int nvp = (nv + 3) & ~3;
    const int units = edge_units(nv);
    if ((long long)gm.n_units + units > P.edge_cap) {
        gm.err = 4;
        return false;
    }
    const int off = gm.n_units;
    gm.n_units += units;
    EdgeBlock eb(gm.edges + off, nvp);
    const uint32_t* vw = sm.vw;
    // sum of the masked prior; if it is not positive ("all valid moves were masked", MCTS_bpp.py:93-100) the prior
    // becomes Ps + valids, i.e. every valid entry is bumped by 1.0, and is renormalised by its own sum.
    // (p + 0.0 == p, so one code path serves both.)
    double bump = 0.0, tot;
#pragma unroll 1
    for (;;) {
        auto term = [&](int a) -> double { return __dadd_rn(prior(a), bump); };
        tot = np_masked_sum(P.plan, vw, term, lane, sm.scratch);
        if (tot > 0.0 || bump != 0.0) break;
        bump = 1.0;
    }
    int base = 0;
    for (int k = 0; k < P.geom.AW; ++k) {
        const uint32_t wk = vw[k];
        if ((wk >> lane) & 1u) {
            const int a = k * 32 + lane;
            const int e = base + __popc(wk & ((1u << lane) - 1u));
            const double p = prior(a);
            eb.P[e] = __ddiv_rn(__dadd_rn(p, bump), tot);
            eb.Q[e] = 0.0;
            eb.NC[e] = make_int2(0, -1);
            eb.ACT[e] = (uint16_t)a;
        }
        base += __popc(wk);
    }
    uint32_t* rec = gm.nodes + (size_t)cur * REC_WORDS;
    if (lane == REC_NS) rec[REC_NS] = 0u;
    if (lane == REC_OFF) rec[REC_OFF] = (uint32_t)off;
    if (lane == REC_META) rec[REC_META] = (uint32_t)nv | ((uint32_t)KIND_EXP << 16);
    __syncwarp();
    return true;
}

// One simulation (MCTS.search from the root, MCTS_bpp.py:56-139).
// Returns 0 = finished and backed up, 1 = parked an unexpanded leaf (STUB == 0 only), -1 = pool overflow.
template <int STUB, int HC>
__device__ __forceinline__ int simulate(const Params& P, GameCtx& gm, WarpSmem& sm, int lane, Stats& st) {
    const Geom& ge = P.geom;
    int cur = gm.root_node;
    if (cur < 0) {
        const uint32_t rrec = P.root_rec[(size_t)gm.g * REC_WORDS + lane];
        cur = lookup_or_insert(ge.H, P.node_cap, P.table_mask, gm, rrec, lane, st);
        if (cur < 0) return -1;
        gm.root_node = cur;
    }
    PathEntry pe = {0, 0, 0, 0};
    int depth = 0;
    double v = 0.0;
    for (;;) {
        const uint32_t rec = gm.nodes[(size_t)cur * REC_WORDS + lane];
        const uint32_t meta = __shfl_sync(FULL, rec, REC_META);
        const int kind = (int)((meta >> 16) & 0xffu);
        if (kind == KIND_TPOS || kind == KIND_TNEG) {  // Es[s] != 0, :81-83
            v = kind == KIND_TPOS ? 1.0 : -1.0;
            st.terminals++;
            break;
        }
        const uint32_t rem = __shfl_sync(FULL, rec, REC_REM);
        if (kind == KIND_NEW) {
            // first visit of this state: Es (getGameEnded) then, if not terminal, expansion
            sm.occ[lane] = rec;
            __syncwarp();
            const uint32_t mine = valid_words<HC>(ge, sm.occ, sm.items, rem, lane, sm.vw, sm.tab, sm.col);
            if (!__any_sync(FULL, mine != 0u)) {
                double score;
                const int r = terminal_value(ge, gm.rc, rec, lane, &score);
                if (lane == REC_META)
                    gm.nodes[(size_t)cur * REC_WORDS + REC_META] = (uint32_t)(r > 0 ? KIND_TPOS : KIND_TNEG) << 16;
                __syncwarp();
                v = (double)r;
                st.terminals++;
                break;
            }
            if (STUB == 0) {
                // park the leaf for the batched evaluator
                int b = 0;
                if (lane == 0) b = atomicAdd(P.leaf_count, 1);
                b = __shfl_sync(FULL, b, 0);
                if (lane == 0) {
                    P.leaf_game[b] = gm.g;
                    P.pend_depth[gm.g] = depth;
                    P.pend_leaf[gm.g] = cur;
                    P.pend_slot[gm.g] = b;
                }
                P.leaf_rec[(size_t)b * REC_WORDS + lane] = state_lane(lane, ge.H) ? rec : 0u;
                P.pend_path[(size_t)gm.g * 32 + lane] = make_int4(pe.node, pe.off, pe.nvp, pe.e);
                if (lane < MAX_AW) P.pend_valid[(size_t)gm.g * MAX_AW + lane] = sm.vw[lane];
                return 1;
            } else {
                const uint32_t rowv = lane < ge.H ? rec : 0u;
                const int pop = (int)__reduce_add_sync(FULL, (unsigned)__popc(rowv));
                const int A = ge.A;
                auto prior = [=](int a) -> double { return stub_prior<STUB>(a, A, pop); };
                if (!expand_node(P, gm, sm, cur, lane, prior)) return -1;
                v = STUB == 1 ? 0.0 : stub_value_of(pop, __popc(rem));
                st.expansions++;
                break;
            }
        }
        // expanded node: PUCT
        const int Ns = (int)__shfl_sync(FULL, rec, REC_NS);
        const int off = (int)__shfl_sync(FULL, rec, REC_OFF);
        const int nv = (int)(meta & 0xffffu);
        const int nvp = (nv + 3) & ~3;
        EdgeBlock eb(gm.edges + off, nvp);
        const int e = puct_select(eb, nv, Ns, P.cpuct, lane, P.sqrt_tab, P.sqrt_n);
        const int act = eb.ACT[e];
        int child = eb.NC[e].y;
        if (lane == depth) {
            pe.node = cur; pe.off = off; pe.nvp = nvp; pe.e = e;
        }
        depth++;
        st.edges++;
        st.units += (unsigned)(3 * nvp + (nvp >> 2));  // 8-byte units of the edge block this selection read
        if (child < 0) {  // first traversal of this edge: getNextState + key lookup (:125-128)
            const int item = div_w(ge, act);
            const int x = act - item * ge.W;
            const int w = sm.items[item] & 0xff, h = sm.items[item] >> 8;
            const uint32_t nrec = apply_move(ge, rec, lane, item, w, h, x);
            child = lookup_or_insert(ge.H, P.node_cap, P.table_mask, gm, nrec, lane, st);
            if (child < 0) return -1;
            if (lane == 0) eb.NC[e].y = child;
            __syncwarp();
        }
        cur = child;
    }
    backup_path(gm.nodes, gm.edges, pe, depth, v, lane, P.sqrt_tab, P.sqrt_n);
    if (lane == 0) P.last_v[gm.g] = v;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// kernels: one warp per game, 4 warps per CTA
constexpr int WARPS_PER_CTA = 4;

// __launch_bounds__(128, 7): 4,096 games = 1,024 CTAs must all be resident at once on 148 SMs (7 CTAs per SM), which caps
// the kernel at 72 registers per thread; at 80 registers only 6 fit and a second, almost empty wave appears.
template <int STUB, int HC>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 7) k_search(Params P) {
    __shared__ WarpSmem smem[WARPS_PER_CTA];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = blockIdx.x * WARPS_PER_CTA + wid;
    if (g >= P.G) return;
    if (P.status[g] != 0) return;
    if (STUB == 0 && P.pend_depth[g] >= 0) return;
    WarpSmem& sm = smem[wid];
    GameCtx gm;
    load_ctx(P, g, lane, gm, sm);
    Stats st = {0, 0, 0, 0, 0, 0, 0};
    int done = P.sims_done[g];
    // lockstep mode: a game whose simulations keep ending on terminal states needs no evaluator; capping its work per
    // launch keeps it from holding back the leaf batch of all other games (it simply continues in the next step)
    const int stop = (STUB == 0 && P.select_cap > 0) ? min(P.num_sims, done + P.select_cap) : P.num_sims;
    while (done < stop) {
        const int rc = simulate<STUB, HC>(P, gm, sm, lane, st);
        if (rc != 0) break;  // parked leaf (counted when it is expanded) or overflow
        done++;
        st.sims++;
    }
    if (lane == 0) P.sims_done[g] = done;
    if (STUB == 0 && lane == 0 && done < P.num_sims && P.pend_depth[g] < 0) atomicAdd(P.leaf_count + 1, 1);
    store_ctx(P, gm, lane);
    flush_stats(P, st, lane);
}

// expansion + backup of the parked leaves; one warp per leaf
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_expand_backup(Params P, const void* policy, int policy_f64, const void* value, int value_f64) {
    __shared__ WarpSmem smem[WARPS_PER_CTA];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.x * WARPS_PER_CTA + wid;
    if (b >= *P.leaf_count) return;
    const int g = P.leaf_game[b];
    WarpSmem& sm = smem[wid];
    GameCtx gm;
    load_ctx(P, g, lane, gm, sm);
    Stats st = {0, 0, 0, 0, 0, 0, 0};
    const int cur = P.pend_leaf[g];
    const int depth = P.pend_depth[g];
    if (lane < MAX_AW) sm.vw[lane] = P.pend_valid[(size_t)g * MAX_AW + lane];
    __syncwarp();
    const int4 pp = P.pend_path[(size_t)g * 32 + lane];
    PathEntry pe = {pp.x, pp.y, pp.z, pp.w};
    const int A = P.geom.A;
    bool ok;
    if (policy_f64) {
        const double* pol = reinterpret_cast<const double*>(policy) + (size_t)b * A;
        ok = expand_node(P, gm, sm, cur, lane, [=](int a) -> double { return pol[a]; });
    } else {
        const float* pol = reinterpret_cast<const float*>(policy) + (size_t)b * A;
        ok = expand_node(P, gm, sm, cur, lane, [=](int a) -> double { return (double)pol[a]; });
    }
    if (ok) {
        const double v = value_f64 ? reinterpret_cast<const double*>(value)[b]
                                   : (double)reinterpret_cast<const float*>(value)[b];
        backup_path(gm.nodes, gm.edges, pe, depth, v, lane, P.sqrt_tab, P.sqrt_n);
        if (lane == 0) P.last_v[g] = v;
        st.expansions++;
        st.sims++;
        if (lane == 0) P.sims_done[g] += 1;
    }
    if (lane == 0) P.pend_depth[g] = -1;
    store_ctx(P, gm, lane);
    flush_stats(P, st, lane);
}

// Put episode `ep` of the queue into game slot g (warp-collective): a fresh MCTS object (CoachBPP.py:124) plus
// getInitBoard / getInitItems (BinPackingGame.py:24-51) - the per-game part of k_reset - and an empty hash table.
__device__ __forceinline__ void load_episode(const Params& P, int g, int ep, int lane) {
    const int N = P.geom.N;
    int w = 0, h = 0;
    if (lane < N) {
        w = P.q_items[((size_t)ep * N + lane) * 2 + 0];
        h = P.q_items[((size_t)ep * N + lane) * 2 + 1];
        P.items_i32[((size_t)g * N + lane) * 2 + 0] = w;
        P.items_i32[((size_t)g * N + lane) * 2 + 1] = h;
    }
    if (lane < BPP_MAX_ITEMS) {
        P.item_w[g * BPP_MAX_ITEMS + lane] = (uint8_t)w;
        P.item_h[g * BPP_MAX_ITEMS + lane] = (uint8_t)h;
    }
    const int max_h = (int)__reduce_max_sync(FULL, (unsigned)h);
    uint32_t* table = P.table + (size_t)g * (P.table_mask + 1u);
    for (unsigned i = lane; i <= P.table_mask; i += 32) table[i] = 0u;
    P.root_rec[(size_t)g * REC_WORDS + lane] =
        lane == REC_REM ? ((N >= 32) ? 0xffffffffu : ((1u << N) - 1u)) : 0u;
    if (lane == 0) {
        const int area = P.q_area[ep];
        const int cdiv = (area + P.geom.W - 1) / P.geom.W;
        P.total_area[g] = area;
        P.numer[g] = cdiv > max_h ? cdiv : max_h;
        P.bl[g] = P.q_bl[ep];
        P.tie[g] = P.q_tie ? P.q_tie[ep] : (int8_t)1;
        P.root_node[g] = -1;
        P.n_nodes[g] = 0;
        P.n_units[g] = 0;
        P.sims_done[g] = 0;
        P.moves_done[g] = 0;
        P.status[g] = 0;
        P.ep_r[g] = 0;
        P.ep_score[g] = 0.0;
        P.pend_depth[g] = -1;
        P.slot_ep[g] = ep;
    }
    __syncwarp();
}

// first G episodes of the stream into the G games; games beyond the queue length stay idle
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_stream_begin(Params P) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = blockIdx.x * WARPS_PER_CTA + wid;
    if (g >= P.G) return;
    if (g < P.q_total) {
        load_episode(P, g, g, lane);
    } else if (lane == 0) {
        P.status[g] = 1;
        P.pend_depth[g] = -1;
        P.slot_ep[g] = -1;
    }
}

__global__ void k_reset(Params P, const int32_t* items_wh, const int32_t* total_area, const double* bl,
                        const int8_t* tie) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P.G) return;
    const int N = P.geom.N;
    int max_h = 0;
    for (int i = 0; i < BPP_MAX_ITEMS; ++i) {
        int w = 0, h = 0;
        if (i < N) {
            w = items_wh[((size_t)g * N + i) * 2 + 0];
            h = items_wh[((size_t)g * N + i) * 2 + 1];
            max_h = h > max_h ? h : max_h;
        }
        P.item_w[g * BPP_MAX_ITEMS + i] = (uint8_t)w;
        P.item_h[g * BPP_MAX_ITEMS + i] = (uint8_t)h;
    }
    const int area = total_area[g];
    const int cdiv = (area + P.geom.W - 1) / P.geom.W;  // np.ceil(items_total_area / bin_width), BinPackingGame.py:198
    P.total_area[g] = area;
    P.numer[g] = cdiv > max_h ? cdiv : max_h;
    P.bl[g] = bl[g];
    P.tie[g] = tie ? tie[g] : (int8_t)1;
    for (int r = 0; r < REC_WORDS; ++r) P.root_rec[(size_t)g * REC_WORDS + r] = 0u;
    P.root_rec[(size_t)g * REC_WORDS + REC_REM] = (N >= 32) ? 0xffffffffu : ((1u << N) - 1u);
    P.root_node[g] = -1;
    P.n_nodes[g] = 0;
    P.n_units[g] = 0;
    P.sims_done[g] = 0;
    P.moves_done[g] = 0;
    P.status[g] = 0;
    P.ep_r[g] = 0;
    P.ep_score[g] = 0.0;
    P.pend_depth[g] = -1;
    P.slot_ep[g] = g;
}

__global__ void k_set_roots(Params P, const uint32_t* roots) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.G * REC_WORDS) return;
    const int lane = i & 31, g = i >> 5;
    P.root_rec[i] = state_lane(lane, P.geom.H) ? roots[i] : 0u;
    if (lane == 0) {
        P.root_node[g] = -1;
        P.sims_done[g] = 0;
        P.pend_depth[g] = -1;
        if (P.status[g] == 1) P.status[g] = 0;
    }
}

// counts[a] = Nsa[(root, a)] for one game (warp-collective)
__device__ __forceinline__ void root_counts_warp(const Params& P, int g, int root, int lane, int32_t* row) {
    const int A = P.geom.A;
    for (int a = lane; a < A; a += 32) row[a] = 0;
    __syncwarp();
    if (root < 0) return;
    const uint32_t* nodes = P.nodes + (size_t)g * P.node_cap * REC_WORDS;
    const uint32_t meta = nodes[(size_t)root * REC_WORDS + REC_META];
    if (((meta >> 16) & 0xffu) != KIND_EXP) return;
    const int nv = (int)(meta & 0xffffu), nvp = (nv + 3) & ~3;
    EdgeBlock eb(P.edges + (size_t)g * (size_t)P.edge_cap + nodes[(size_t)root * REC_WORDS + REC_OFF], nvp);
    for (int e = lane; e < nv; e += 32) row[eb.ACT[e]] = eb.NC[e].x;
}

// Mirror of one visit-count row into mapped host memory as full 128-byte stores (k_episode writes the results of
// bpp_engine_play_stub_host straight into the caller's pinned buffer while the episodes run).  noinline + scalar
// arguments: called once per move, kept out of the search loop's register allocation.
__device__ __noinline__ void mirror_row(const int32_t* dev_row, int32_t* host_row, int A, int lane) {
    __syncwarp();
    for (int a = lane; a < A; a += 32) host_row[a] = dev_row[a];
}
// rows of moves a game did not play: counts 0, action -1
__device__ __noinline__ void fill_unplayed(int32_t* counts_host, int32_t* actions_out, int g, int G, int A, int m0, int m1,
                                           int lane) {
    for (int m = m0; m < m1; ++m) {
        if (counts_host)
            for (int a = lane; a < A; a += 32) counts_host[((size_t)m * G + g) * A + a] = 0;
        if (actions_out && lane == 0) actions_out[(size_t)m * G + g] = -1;
    }
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_root_counts(Params P, int32_t* out) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = blockIdx.x * WARPS_PER_CTA + wid;
    if (g >= P.G) return;
    root_counts_warp(P, g, P.root_node[g], lane, out + (size_t)g * P.geom.A);
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// action choice from the root visit counts of one game (executed by ONE thread; <= A edges).  The random stream is a
// pure function of (seed, game, move number), so the per-move kernels and the whole-episode kernel draw the same.
__device__ __forceinline__ int choose_action(const Params& P, int g, int root, int move_no, int mode,
                                             unsigned long long seed, int key = -1) {
    if (key < 0) key = g;  // stream key: the episode (== the game index unless episodes are streamed through the games)
    int act = -1;
    if (root < 0) return act;
    const uint32_t* nodes = P.nodes + (size_t)g * P.node_cap * REC_WORDS;
    const uint32_t meta = nodes[(size_t)root * REC_WORDS + REC_META];
    if (((meta >> 16) & 0xffu) != KIND_EXP) return act;
    const int nv = (int)(meta & 0xffffu), nvp = (nv + 3) & ~3;
    EdgeBlock eb(P.edges + (size_t)g * (size_t)P.edge_cap + nodes[(size_t)root * REC_WORDS + REC_OFF], nvp);
    const unsigned long long rnd =
        splitmix64(seed ^ splitmix64(((unsigned long long)key << 20) ^ (unsigned long long)move_no));
    if (mode == BPP_CHOOSE_SAMPLE) {  // a ~ counts / sum(counts), CoachBPP.py:86-87
        long long tot = 0;
        for (int e = 0; e < nv; ++e) tot += eb.NC[e].x;
        if (tot > 0) {
            long long t = (long long)(rnd % (unsigned long long)tot);
            for (int e = 0; e < nv; ++e) {
                t -= eb.NC[e].x;
                if (t < 0) { act = eb.ACT[e]; break; }
            }
        }
    } else {
        int best = -1, nbest = 0;
        for (int e = 0; e < nv; ++e) {
            const int n = eb.NC[e].x;
            if (n > best) { best = n; nbest = 1; act = eb.ACT[e]; }
            else if (n == best) nbest++;
        }
        if (mode == BPP_CHOOSE_GREEDY && nbest > 1) {  // uniformly random arg-max, MCTS_bpp.py:43-49
            int k = (int)(rnd % (unsigned long long)nbest);
            for (int e = 0; e < nv; ++e)
                if (eb.NC[e].x == best && k-- == 0) { act = eb.ACT[e]; break; }
        }
    }
    return act;
}

__global__ void k_choose(Params P, int mode, unsigned long long seed, int32_t* actions) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P.G) return;
    actions[g] = P.status[g] == 0 ? choose_action(P, g, P.root_node[g], P.moves_done[g], mode, seed) : -1;
}

// Play one real move in one game (CoachBPP.py:88-98): root <- getNextState(root, a), then getGameEnded on the new root.
// Warp-collective.  `rec` = the lane's word of the root record (in/out).  Returns the new status (0 running, 1 ended,
// -1 illegal action) and latches (r, score) when the episode ends.
__device__ __forceinline__ int advance_game(const Params& P, GameCtx& gm, WarpSmem& sm, int lane, int a, uint32_t& rec,
                                            Stats& st) {
    const Geom& ge = P.geom;
    const int g = gm.g;
    const uint32_t rem = __shfl_sync(FULL, rec, REC_REM);
    const int item = a >= 0 ? div_w(ge, a) : -1;
    if (a < 0 || a >= ge.A || !((rem >> item) & 1u)) {  // reference: assert sum(sum(item)) > 0, BinPackingGame.py:69
        if (lane == 0) P.status[g] = -1;
        return -1;
    }
    const int x = a - item * ge.W;
    rec = apply_move(ge, rec, lane, item, sm.items[item] & 0xff, sm.items[item] >> 8, x);
    rec = state_lane(lane, ge.H) ? rec : 0u;
    P.root_rec[(size_t)g * REC_WORDS + lane] = rec;
    gm.root_node = lookup_or_insert(ge.H, P.node_cap, P.table_mask, gm, rec, lane, st);
    // getGameEnded on the new root (CoachBPP.py:91)
    sm.occ[lane] = rec;
    __syncwarp();
    const uint32_t nrem = __shfl_sync(FULL, rec, REC_REM);
    const uint32_t mine = valid_words<0>(ge, sm.occ, sm.items, nrem, lane, sm.vw, sm.tab, sm.col);
    int status = 0;
    if (!__any_sync(FULL, mine != 0u)) {
        double score;
        const int r = terminal_value(ge, gm.rc, rec, lane, &score);
        status = 1;
        if (lane == 0) {
            P.status[g] = 1;
            P.ep_r[g] = r;
            P.ep_score[g] = score;
        }
    }
    if (lane == 0) {
        P.moves_done[g] += 1;
        P.sims_done[g] = 0;
    }
    if (gm.err) {
        if (lane == 0) P.status[g] = -gm.err;
        status = -gm.err;
        gm.err = 0;
    }
    return status;
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_advance(Params P, const int32_t* actions) {
    __shared__ WarpSmem smem[WARPS_PER_CTA];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = blockIdx.x * WARPS_PER_CTA + wid;
    if (g >= P.G) return;
    if (P.status[g] != 0) return;
    WarpSmem& sm = smem[wid];
    GameCtx gm;
    load_ctx(P, g, lane, gm, sm);
    Stats st = {0, 0, 0, 0, 0, 0, 0};
    uint32_t rec = P.root_rec[(size_t)g * REC_WORDS + lane];
    if (advance_game(P, gm, sm, lane, actions[g], rec, st) == -1) return;
    store_ctx(P, gm, lane);
    flush_stats(P, st, lane);
}

// Lockstep step in ONE launch per game warp: expansion + backup of the game's parked leaf (evaluator outputs of the
// previous step, row pend_slot[g]) followed by the next descents until the game parks its next leaf (k_expand_backup +
// k_search<0>): the game context is loaded once, and the second launch with its tail disappears from every step.  New
// leaves go to slots counted from zero (leaf_count is cleared before the launch); nothing in this kernel reads the
// old leaf records, and the evaluator outputs are only read.
template <int HC>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 7)
k_expand_search(Params P, const void* policy, int policy_f64, const void* value, int value_f64) {
    __shared__ WarpSmem smem[WARPS_PER_CTA];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = blockIdx.x * WARPS_PER_CTA + wid;
    // the evaluator's trunk kernel, launched programmatically behind this one, may set itself up (barriers, tensor memory,
    // first weights) on the SMs this kernel's short warps have left while its longest descents still run
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (g >= P.G) return;
    if (P.status[g] != 0) return;
    WarpSmem& sm = smem[wid];
    GameCtx gm;
    load_ctx(P, g, lane, gm, sm);
    Stats st = {0, 0, 0, 0, 0, 0, 0};
    int done = P.sims_done[g];
    const int depth = P.pend_depth[g];
    bool ok = true;
    if (depth >= 0) {
        const int b = P.pend_slot[g];
        const int cur = P.pend_leaf[g];
        if (lane < MAX_AW) sm.vw[lane] = P.pend_valid[(size_t)g * MAX_AW + lane];
        __syncwarp();
        const int4 pp = P.pend_path[(size_t)g * 32 + lane];
        PathEntry pe = {pp.x, pp.y, pp.z, pp.w};
        const size_t row = (size_t)b * P.geom.A;
        const double* pol64 = reinterpret_cast<const double*>(policy) + row;
        const float* pol32 = reinterpret_cast<const float*>(policy) + row;
        ok = expand_node(P, gm, sm, cur, lane,
                         [=](int a) -> double { return policy_f64 ? pol64[a] : (double)pol32[a]; });
        if (ok) {
            const double v = value_f64 ? reinterpret_cast<const double*>(value)[b]
                                       : (double)reinterpret_cast<const float*>(value)[b];
            backup_path(gm.nodes, gm.edges, pe, depth, v, lane, P.sqrt_tab, P.sqrt_n);
            if (lane == 0) P.last_v[g] = v;
            st.expansions++;
            st.sims++;
            done++;
        }
        if (lane == 0) P.pend_depth[g] = -1;
        __syncwarp();
    }
    bool running = ok;
    if (ok) {
        int budget = P.select_cap > 0 ? P.select_cap : 0x7fffffff;
        for (;;) {
            bool parked = false;
            while (done < P.num_sims && budget > 0 && (P.edge_budget == 0 || st.edges < (unsigned)P.edge_budget)) {
                const int rc = simulate<0, HC>(P, gm, sm, lane, st);
                if (rc != 0) {  // parked leaf (counted when it is expanded) or overflow
                    parked = true;
                    running = rc > 0;
                    break;
                }
                done++;
                st.sims++;
                budget--;
            }
            if (parked || done < P.num_sims || P.auto_mode < 0) break;
            // asynchronous episodes: this game's move is searched (MCTS_bpp.py:37-41) -> counts out, choose
            // (CoachBPP.py:86-87 / MCTS_bpp.py:43-49), play (CoachBPP.py:88-98) and straight on to the next move; no
            // game waits for the slowest game of a move and the host never synchronises per move
            const int m = P.moves_done[g];
            uint32_t rec = P.root_rec[(size_t)g * REC_WORDS + lane];
            const int ep = P.q_total ? P.slot_ep[g] : g;
            const size_t mg = (size_t)m * (P.q_total ? P.q_total : P.G) + ep;
            if (P.ep_roots) P.ep_roots[mg * REC_WORDS + lane] = rec;
            if (P.ep_counts) root_counts_warp(P, g, gm.root_node, lane, P.ep_counts + mg * P.geom.A);
            int a = lane == 0 ? choose_action(P, g, gm.root_node, m, P.auto_mode, P.auto_seed, ep) : 0;
            a = __shfl_sync(FULL, a, 0);
            if (P.ep_actions && lane == 0) P.ep_actions[mg] = a;
            const int status = advance_game(P, gm, sm, lane, a, rec, st);
            done = 0;
            if (status == 0) continue;
            running = false;
            if (status != 1 || !P.q_total) break;
            // the episode has ended: latch its outcome and take the next instance of the queue into this game
            int nxt = 0;
            if (lane == 0) {
                if (P.epq_r) P.epq_r[ep] = P.ep_r[g];
                if (P.epq_score) P.epq_score[ep] = P.ep_score[g];
                if (P.epq_moves) P.epq_moves[ep] = P.moves_done[g];
                nxt = atomicAdd(P.q_next, 1);
            }
            nxt = __shfl_sync(FULL, nxt, 0);
            if (nxt >= P.q_total) break;
            load_episode(P, g, nxt, lane);
            load_ctx(P, g, lane, gm, sm);
            running = true;
        }
    }
    if (lane == 0) {
        P.sims_done[g] = done;
        if (running && done < P.num_sims && P.pend_depth[g] < 0) atomicAdd(P.leaf_count + 1, 1);
        if (running) atomicAdd(P.leaf_count + 2, 1);  // games still playing their episode
    }
    store_ctx(P, gm, lane);
    flush_stats(P, st, lane);
}

// Whole self-play episodes in ONE launch (stub evaluators): per game, loop {numMCTSSims simulations -> visit counts out
// -> choose -> play the move} until the episode ends.  Removes the per-move launch boundaries: a game that needs longer
// for one move no longer holds back the others (the only tail left is the end of the batch).
template <int STUB, int HC>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 7)
k_episode(Params P, int mode, unsigned long long seed, int max_moves, int32_t* counts_out, int32_t* actions_out,
          int32_t* counts_host) {
    // counts_host != nullptr: results go to mapped host memory (counts_out is then the device scratch the rows are built
    // in, actions_out the host alias) and this kernel also writes the rows of unplayed moves
    __shared__ WarpSmem smem[WARPS_PER_CTA];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = blockIdx.x * WARPS_PER_CTA + wid;
    if (g >= P.G) return;
    if (P.status[g] != 0) {
        if (counts_host && !P.q_total) fill_unplayed(counts_host, actions_out, g, P.G, P.geom.A, 0, max_moves, lane);
        return;
    }
    WarpSmem& sm = smem[wid];
    GameCtx gm;
    load_ctx(P, g, lane, gm, sm);
    Stats st = {0, 0, 0, 0, 0, 0, 0};
    uint32_t rec = P.root_rec[(size_t)g * REC_WORDS + lane];
    int move_no = P.moves_done[g];
    int done = P.sims_done[g];
    // rows of the outputs: [move][game] or, with an episode stream (q_total > 0, bpp_engine_play_stub_stream),
    // [move][episode]: a game whose episode ends takes the next instance of the queue, so the launch has no idle tail
    // but the end of the whole stream
    const int rows = P.q_total ? P.q_total : P.G;
    int ep = P.q_total ? P.slot_ep[g] : g;
    const size_t RA = (size_t)rows * P.geom.A;
    int m = 0;
    bool open = true;   // an episode is in progress in this game (its unplayed rows are still to be written)
    while (m < max_moves) {
        while (done < P.num_sims) {
            if (simulate<STUB, HC>(P, gm, sm, lane, st) != 0) break;
            done++;
            st.sims++;
        }
        if (gm.err) break;
        if (counts_out) {
            int32_t* row = counts_out + m * RA + (size_t)ep * P.geom.A;
            root_counts_warp(P, g, gm.root_node, lane, row);
            if (counts_host) mirror_row(row, counts_host + m * RA + (size_t)ep * P.geom.A, P.geom.A, lane);
        }
        int a = lane == 0 ? choose_action(P, g, gm.root_node, move_no, mode, seed, ep) : 0;
        a = __shfl_sync(FULL, a, 0);
        if (actions_out && lane == 0) actions_out[(size_t)m * rows + ep] = a;
        const int status = advance_game(P, gm, sm, lane, a, rec, st);
        move_no++;
        ++m;
        done = 0;
        if (status == 0) continue;
        if (!P.q_total || status != 1) break;
        // the episode has ended: latch its outcome and take the next instance of the queue into this game
        if (counts_host) fill_unplayed(counts_host, actions_out, ep, rows, P.geom.A, m, max_moves, lane);
        open = false;
        int nxt = 0;
        if (lane == 0) {
            if (P.epq_r) P.epq_r[ep] = P.ep_r[g];
            if (P.epq_score) P.epq_score[ep] = P.ep_score[g];
            if (P.epq_moves) P.epq_moves[ep] = P.moves_done[g];
            nxt = atomicAdd(P.q_next, 1);
        }
        nxt = __shfl_sync(FULL, nxt, 0);
        if (nxt >= P.q_total) break;
        load_episode(P, g, nxt, lane);
        load_ctx(P, g, lane, gm, sm);
        rec = P.root_rec[(size_t)g * REC_WORDS + lane];
        ep = nxt;
        move_no = 0;
        m = 0;
        open = true;
    }
    store_ctx(P, gm, lane);
    flush_stats(P, st, lane);
    if (counts_host && open) fill_unplayed(counts_host, actions_out, ep, rows, P.geom.A, m, max_moves, lane);
}

// dense evaluator input: planes [B][N+1][H][W] float32 (getBinItem, BinPackingGame.py:118-120)
__global__ void k_leaf_planes(Params P, float* out) {
    const int B = *P.leaf_count;
    const Geom& ge = P.geom;
    const int per = (ge.N + 1) * ge.H * ge.W;
    const long long total = (long long)B * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / per);
        int r = (int)(i - (long long)b * per);
        const int c = r / (ge.H * ge.W);
        r -= c * ge.H * ge.W;
        const int y = r / ge.W, x = r - y * ge.W;
        const uint32_t* rec = P.leaf_rec + (size_t)b * REC_WORDS;
        float v;
        if (c == 0) {
            v = (float)((rec[y] >> x) & 1u);
        } else {
            const int g = P.leaf_game[b];
            const int it = c - 1;
            const bool remaining = (rec[REC_REM] >> it) & 1u;
            v = (remaining && y < P.item_h[g * BPP_MAX_ITEMS + it] && x < P.item_w[g * BPP_MAX_ITEMS + it]) ? 1.f : 0.f;
        }
        out[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stateless env kernels (one warp per state)
struct EnvArgs {
    Geom geom;
    int n;
    const uint32_t* recs;
    const int32_t* items_wh;
};
__device__ __forceinline__ void env_load(const EnvArgs& E, int s, int lane, WarpSmem& sm, uint32_t& rec) {
    rec = E.recs[(size_t)s * REC_WORDS + lane];
    rec = state_lane(lane, E.geom.H) ? rec : 0u;
    sm.occ[lane] = rec;
    if (lane < BPP_MAX_ITEMS) {
        int w = 0, h = 0;
        if (lane < E.geom.N) {
            w = E.items_wh[((size_t)s * E.geom.N + lane) * 2 + 0];
            h = E.items_wh[((size_t)s * E.geom.N + lane) * 2 + 1];
        }
        sm.items[lane] = (uint16_t)((w & 0xff) | ((h & 0xff) << 8));
    }
    __syncwarp();
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_env_valid(EnvArgs E, uint8_t* out) {
    __shared__ WarpSmem smem[WARPS_PER_CTA];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int s = blockIdx.x * WARPS_PER_CTA + wid;
    if (s >= E.n) return;
    WarpSmem& sm = smem[wid];
    uint32_t rec;
    env_load(E, s, lane, sm, rec);
    const uint32_t rem = __shfl_sync(FULL, rec, REC_REM);
    valid_words<0>(E.geom, sm.occ, sm.items, rem, lane, sm.vw, sm.tab, sm.col);
    for (int a = lane; a < E.geom.A; a += 32) out[(size_t)s * E.geom.A + a] = (uint8_t)((sm.vw[a >> 5] >> (a & 31)) & 1u);
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_env_next(EnvArgs E, const int32_t* actions, uint32_t* out) {
    __shared__ WarpSmem smem[WARPS_PER_CTA];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int s = blockIdx.x * WARPS_PER_CTA + wid;
    if (s >= E.n) return;
    WarpSmem& sm = smem[wid];
    uint32_t rec;
    env_load(E, s, lane, sm, rec);
    const int a = actions[s];
    if (a >= 0 && a < E.geom.A) {
        const int item = a / E.geom.W, x = a - item * E.geom.W;
        rec = apply_move(E.geom, rec, lane, item, sm.items[item] & 0xff, sm.items[item] >> 8, x);
    }
    out[(size_t)s * REC_WORDS + lane] = state_lane(lane, E.geom.H) ? rec : 0u;
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_env_ended(EnvArgs E, const int32_t* total_area, const int32_t* max_h, const double* bl, const int8_t* tie,
            int32_t* ended, double* score_out) {
    __shared__ WarpSmem smem[WARPS_PER_CTA];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int s = blockIdx.x * WARPS_PER_CTA + wid;
    if (s >= E.n) return;
    WarpSmem& sm = smem[wid];
    uint32_t rec;
    env_load(E, s, lane, sm, rec);
    const uint32_t rem = __shfl_sync(FULL, rec, REC_REM);
    const uint32_t mine = valid_words<0>(E.geom, sm.occ, sm.items, rem, lane, sm.vw, sm.tab, sm.col);
    int res = 0;
    double score = 0.0;
    if (!__any_sync(FULL, mine != 0u)) {
        RewardCtx rc;
        rc.total_area = total_area[s];
        const int cdiv = (rc.total_area + E.geom.W - 1) / E.geom.W;
        rc.numer = cdiv > max_h[s] ? cdiv : max_h[s];
        rc.bl = bl[s];
        rc.tie = tie ? tie[s] : 1;
        res = terminal_value(E.geom, rc, rec, lane, &score);
    }
    if (lane == 0) {
        ended[s] = res;
        score_out[s] = score;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Device-side ItemsGenerator.items_generator (BinPackingGame.py:257-285): one thread per seed replays numpy's legacy
// global RNG — MT19937 seeded by init_genrand(seed) (np.random.seed(int)), randint = masked rejection over 32-bit
// draws with NO draw when the range holds a single value — and the guillotine-split loop, so the instances are
// identical to the reference's for the same seeds.  The 2.5 KB generator state lives in local memory.
struct DevMT {
    uint32_t mt[624];
    int pos;
    __device__ void seed(uint32_t s) {
        for (int i = 0; i < 624; ++i) {
            mt[i] = s;
            s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i + 1u;
        }
        pos = 624;
    }
    __device__ void twist() {
        for (int i = 0; i < 624; ++i) {
            const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
            mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        pos = 0;
    }
    __device__ uint32_t u32() {
        if (pos >= 624) twist();
        uint32_t y = mt[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    __device__ int randint(int low, int high) {  // np.random.randint(low, high), high exclusive
        const uint32_t rng = (uint32_t)(high - 1 - low);
        if (rng == 0u) return low;
        uint32_t mask = rng;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        do { v = u32() & mask; } while (v > rng);
        return low + (int)v;
    }
};

__global__ void k_items_generate(int W, int N, int n, const int64_t* __restrict__ seeds,
                                 const int32_t* __restrict__ heights, int32_t* __restrict__ items_wh,
                                 int32_t* __restrict__ rects_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DevMT r;
    r.seed((uint32_t)seeds[i]);
    short rc[2 * BPP_MAX_ITEMS + 2][4];  // [w, h, a, b]
    int len = 1;
    rc[0][0] = (short)W; rc[0][1] = (short)heights[i]; rc[0][2] = 0; rc[0][3] = 0;
    while (len < N) {
        const int axis = r.randint(0, 2);
        const int k = r.randint(0, len);
        const int w = rc[k][0], h = rc[k][1], a = rc[k][2], b = rc[k][3];
        if (axis == 0) {
            if (w == 1) continue;
            const int cut = r.randint(a + 1, a + w);
            rc[len][0] = (short)(cut - a); rc[len][1] = (short)h; rc[len][2] = (short)a; rc[len][3] = (short)b;
            rc[len + 1][0] = (short)(w - (cut - a)); rc[len + 1][1] = (short)h; rc[len + 1][2] = (short)cut; rc[len + 1][3] = (short)b;
        } else {
            if (h == 1) continue;
            const int cut = r.randint(b + 1, b + h);
            rc[len][0] = (short)w; rc[len][1] = (short)(cut - b); rc[len][2] = (short)a; rc[len][3] = (short)b;
            rc[len + 1][0] = (short)w; rc[len + 1][1] = (short)(h - (cut - b)); rc[len + 1][2] = (short)a; rc[len + 1][3] = (short)cut;
        }
        len += 2;
        for (int j = k; j < len - 1; ++j) {  // item_list.pop(idx_item)
            rc[j][0] = rc[j + 1][0]; rc[j][1] = rc[j + 1][1]; rc[j][2] = rc[j + 1][2]; rc[j][3] = rc[j + 1][3];
        }
        len -= 1;
    }
    for (int j = 0; j < N; ++j) {
        items_wh[((size_t)i * N + j) * 2 + 0] = rc[j][0];
        items_wh[((size_t)i * N + j) * 2 + 1] = rc[j][1];
        if (rects_out) {
            for (int c = 0; c < 4; ++c) rects_out[((size_t)i * N + j) * 4 + c] = rc[j][c];
        }
    }
}

// dense planes of arbitrary compact states (learner input batches): out float32 [n][N+1][H][W]
__global__ void k_env_planes(EnvArgs E, float* out) {
    const Geom& ge = E.geom;
    const int per = (ge.N + 1) * ge.H * ge.W;
    const long long total = (long long)E.n * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / per);
        int r = (int)(i - (long long)b * per);
        const int c = r / (ge.H * ge.W);
        r -= c * ge.H * ge.W;
        const int y = r / ge.W, x = r - y * ge.W;
        const uint32_t* rec = E.recs + (size_t)b * REC_WORDS;
        float v;
        if (c == 0) {
            v = (float)((rec[y] >> x) & 1u);
        } else {
            const int it = c - 1;
            const int32_t* wh = E.items_wh + ((size_t)b * ge.N + it) * 2;
            v = (((rec[REC_REM] >> it) & 1u) && y < wh[1] && x < wh[0]) ? 1.f : 0.f;
        }
        out[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// returns the index of the leaf slot that will hold the sum of [base, base+n)
static int build_sum_plan_rec(SumPlan& p, int base, int n) {
    if (n <= 128) {  // numpy PW_BLOCKSIZE
        p.leaf_base[p.n_leaves] = (short)base;
        p.leaf_n[p.n_leaves] = (short)n;
        return p.n_leaves++;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    const int l = build_sum_plan_rec(p, base, n2);
    const int r = build_sum_plan_rec(p, base + n2, n - n2);
    p.op_dst[p.n_ops] = (signed char)l;
    p.op_src[p.n_ops] = (signed char)r;
    p.n_ops++;
    return l;
}

static int make_geom(int W, int H, int N, Geom* g) {
    if (W < 1 || W > 32 || H < 1 || H > 28 || N < 1 || N > BPP_MAX_ITEMS)
        return set_err(BPP_E_INVALID, "unsupported geometry W=%d H=%d N=%d (limits: W<=32, H<=28, N<=16)", W, H, N);
    g->W = W; g->H = H; g->N = N; g->A = W * N; g->AW = (W * N + 31) / 32;
    g->wmask = W >= 32 ? 0xffffffffu : ((1u << W) - 1u);
    g->invW = (65536u + (unsigned)W - 1u) / (unsigned)W;
    return BPP_OK;
}

struct bpp_engine {
    bpp_config cfg;
    Params P;
    std::vector<void*> allocs;
    int64_t bytes = 0;
    unsigned long long launches = 0;
    bool leaf_parked = false;
    int unfinished = 0;  // games that hit the select cap in the last select (still owe simulations)
    int32_t* d_actions = nullptr;  // scratch for play_stub
    int32_t* d_items = nullptr;    // staging for the *_host entry points
    int32_t* d_area = nullptr;
    double* d_bl = nullptr;
    int8_t* d_tie = nullptr;
    int32_t* d_counts = nullptr;
    int32_t* d_counts_all = nullptr;   // [N][G][A], lazily allocated by play_stub_host
    int32_t* d_actions_all = nullptr;  // [N][G]
    int* h_status = nullptr;
    // bpp_engine_play_net: evaluator outputs of one lockstep step, progress counters read back asynchronously
    float* d_pol = nullptr;            // [G][A]
    float* d_val = nullptr;            // [G]
    uint32_t* d_roots_all = nullptr;   // [N][G][32], lazily allocated by play_net_host
    int32_t* h_prog = nullptr;         // pinned [4][8]: leaf_count[4], simulations (u64), -
    cudaEvent_t ev_prog[4] = {nullptr, nullptr, nullptr, nullptr};
    const int32_t* items_ref = nullptr;  // int32 [G][N][2] item list of the current episodes (evaluator input)
    int num_sms = 148;
    // staging of bpp_engine_play_net_stream_host (device; element counts)
    int32_t* q_items = nullptr; size_t q_items_n = 0;
    int32_t* q_area = nullptr; size_t q_area_n = 0;
    double* q_bl = nullptr; size_t q_bl_n = 0;
    int8_t* q_tie = nullptr; size_t q_tie_n = 0;
    int32_t* q_r = nullptr; size_t q_r_n = 0;
    double* q_score = nullptr; size_t q_score_n = 0;
    int32_t* q_moves = nullptr; size_t q_moves_n = 0;
    int32_t* q_counts = nullptr; size_t q_counts_n = 0;
    int32_t* q_actions = nullptr; size_t q_actions_n = 0;
    uint32_t* q_roots = nullptr; size_t q_roots_n = 0;
    bool prof_on = false;              // bpp_engine_set_profile: CUDA events around every call of play_net's steps
    std::vector<cudaEvent_t> prof_ev;
    double prof_ms[4] = {0, 0, 0, 0};  // evaluator ms, expand+select ms, steps timed, -
};

template <typename T>
static int dev_alloc(bpp_engine* e, T** p, size_t count) {
    void* q = nullptr;
    const size_t bytes = count * sizeof(T);
    cudaError_t err = cudaMalloc(&q, bytes ? bytes : 1);
    if (err != cudaSuccess) {
        cudaGetLastError();
        return set_err(BPP_E_NOMEM, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(err));
    }
    e->allocs.push_back(q);
    e->bytes += (int64_t)bytes;
    *p = reinterpret_cast<T*>(q);
    return BPP_OK;
}

extern "C" int bpp_engine_destroy(bpp_engine* e) {
    if (!e) return BPP_OK;
    for (void* p : e->allocs) cudaFree(p);
    if (e->h_status) cudaFreeHost(e->h_status);
    if (e->h_prog) cudaFreeHost(e->h_prog);
    for (int i = 0; i < 4; ++i)
        if (e->ev_prog[i]) cudaEventDestroy(e->ev_prog[i]);
    for (cudaEvent_t ev : e->prof_ev) cudaEventDestroy(ev);
    delete e;
    return BPP_OK;
}

// inside bpp_engine_create, after the handle exists: a failing CUDA call frees everything the handle owns
#define CREATE_TRY(expr)                                                                                 \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            bpp_engine_destroy(e);                                                                       \
            return set_err(BPP_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                           __LINE__);                                                                    \
        }                                                                                                \
    } while (0)

extern "C" int bpp_engine_create(const bpp_config* cfg, bpp_engine** out) {
    if (!cfg || !out) return set_err(BPP_E_INVALID, "null argument");
    *out = nullptr;
    Geom ge;
    int rc = make_geom(cfg->W, cfg->H, cfg->N, &ge);
    if (rc) return rc;
    if (cfg->G < 1 || cfg->num_sims < 0) return set_err(BPP_E_INVALID, "G=%d num_sims=%d", cfg->G, cfg->num_sims);
    CUDA_TRY(cudaSetDevice(cfg->device));
    bpp_engine* e = new bpp_engine();
    e->cfg = *cfg;
    if (cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, cfg->device) != cudaSuccess || e->num_sms < 1)
        e->num_sms = 148;
    Params& P = e->P;
    memset(&P, 0, sizeof(P));
    P.geom = ge;
    P.G = cfg->G;
    P.num_sims = cfg->num_sims;
    P.cpuct = cfg->cpuct;
    // every simulation creates at most one node, every real move at most one more (+ the first root)
    P.node_cap = cfg->node_cap > 0 ? cfg->node_cap : cfg->num_sims * cfg->N + cfg->N + 2;
    if (P.node_cap >= (1 << 20) - 1) {
        bpp_engine_destroy(e);
        return set_err(BPP_E_INVALID, "node_cap %d too large (max 2^20-2)", P.node_cap);
    }
    unsigned tc = 64;
    while (tc < 2u * (unsigned)P.node_cap) tc <<= 1;
    P.table_mask = tc - 1u;
    memset(&P.plan, 0, sizeof(P.plan));
    build_sum_plan_rec(P.plan, 0, ge.A);
    const size_t G = (size_t)cfg->G;
#define ALLOC(ptr, count)                          \
    if ((rc = dev_alloc(e, &(ptr), (count))) != 0) { \
        bpp_engine_destroy(e);                     \
        return rc;                                 \
    }
    ALLOC(P.item_w, G * BPP_MAX_ITEMS);
    ALLOC(P.item_h, G * BPP_MAX_ITEMS);
    ALLOC(P.total_area, G);
    ALLOC(P.numer, G);
    ALLOC(P.bl, G);
    ALLOC(P.tie, G);
    ALLOC(P.root_rec, G * REC_WORDS);
    ALLOC(P.root_node, G);
    ALLOC(P.n_nodes, G);
    ALLOC(P.n_units, G);
    ALLOC(P.sims_done, G);
    ALLOC(P.moves_done, G);
    ALLOC(P.status, G);
    ALLOC(P.ep_r, G);
    ALLOC(P.ep_score, G);
    ALLOC(P.last_v, G);
    ALLOC(P.pend_depth, G);
    ALLOC(P.pend_leaf, G);
    ALLOC(P.pend_slot, G);
    ALLOC(P.pend_path, G * 32);
    ALLOC(P.pend_valid, G * MAX_AW);
    ALLOC(P.leaf_count, 4);  // [0] parked leaves, [1] games that stopped at the per-launch cap, [2] games still playing
    ALLOC(P.leaf_game, G);
    ALLOC(P.leaf_rec, G * REC_WORDS);
    ALLOC(P.stats, 8);
    ALLOC(P.slot_ep, G);
    ALLOC(P.q_next, 1);
    ALLOC(e->d_actions, G);
    ALLOC(e->d_items, G * (size_t)ge.N * 2);
    P.items_i32 = e->d_items;
    ALLOC(e->d_area, G);
    ALLOC(e->d_bl, G);
    ALLOC(e->d_tie, G);
    ALLOC(e->d_counts, G * (size_t)ge.A);
    ALLOC(P.nodes, G * (size_t)P.node_cap * REC_WORDS);
    ALLOC(P.table, G * (size_t)tc);
    // Edge pool.  Worst case = every node expanded with all A actions valid (7.9 MB per game at the default config);
    // measured need over whole 10-move episodes: mean 47 K units, maximum 87 K units per game = 9 % of that worst case
    // (15x15, 200 simulations).  Default = a quarter of the worst case (3x the measured maximum), clipped to half of the
    // free memory so that a second engine, the evaluator and the learner still fit in the same process.  A game that
    // outgrows its pool sets the sticky BPP_E_CAPACITY error (bpp_engine_check): the caller re-creates the engine
    // with a larger cfg.edge_cap (cfg.edge_cap = worst case never overflows) and replays the batch.
    const long long worst = (long long)P.node_cap * edge_units(ge.A);
    long long cap = cfg->edge_cap > 0 ? cfg->edge_cap : worst / 4;
    if (cfg->edge_cap <= 0) {
        size_t free_b = 0, total_b = 0;
        CREATE_TRY(cudaMemGetInfo(&free_b, &total_b));
        const long long fit = (long long)((double)free_b * 0.50 / (double)G / 8.0);
        if (cap > fit) cap = fit;
        const long long floor_units = (long long)edge_units(ge.A) * 4;
        if (cap < floor_units) cap = floor_units;
    }
    if (cap > 0x7fffffffll) cap = 0x7fffffffll;
    P.edge_cap = cap;
    ALLOC(P.edges, G * (size_t)cap);
#undef ALLOC
#define ALLOC2(ptr, count)                         \
    if ((rc = dev_alloc(e, &(ptr), (count))) != 0) { \
        bpp_engine_destroy(e);                     \
        return rc;                                 \
    }
    {   // Ns[s] never exceeds the simulations of an episode (num_sims per move, at most N moves)
        const int tn = cfg->num_sims * cfg->N + 2;
        std::vector<double> tab(3 * (size_t)tn);
        for (int i = 0; i < tn; ++i) {
            tab[i] = sqrt((double)i);
            tab[tn + i] = sqrt((double)i + 1e-8);
            tab[2 * tn + i] = i ? 1.0 / (double)i : 0.0;
        }
        double* d_tab = nullptr;
        ALLOC2(d_tab, 3 * (size_t)tn);
        CREATE_TRY(cudaMemcpy(d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
        P.sqrt_tab = d_tab;
        P.sqrt_n = tn;
    }
    CREATE_TRY(cudaMemset(P.stats, 0, 8 * sizeof(unsigned long long)));
    CREATE_TRY(cudaMemset(P.status, 0xff, G * sizeof(int)));  // not reset yet
    CREATE_TRY(cudaMemset(P.pend_depth, 0xff, G * sizeof(int)));
    CREATE_TRY(cudaMemset(P.leaf_count, 0, 4 * sizeof(int)));
    CREATE_TRY(cudaMallocHost(&e->h_status, G * sizeof(int)));
    CREATE_TRY(cudaMallocHost(&e->h_prog, 32 * sizeof(int32_t)));
    for (int i = 0; i < 4; ++i) CREATE_TRY(cudaEventCreate(&e->ev_prog[i]));
    P.auto_mode = -1;
    *out = e;
    return BPP_OK;
}

extern "C" int64_t bpp_engine_device_bytes(const bpp_engine* e) { return e ? e->bytes : 0; }
extern "C" int bpp_engine_edge_cap(const bpp_engine* e, int64_t* units_out) {
    if (!e || !units_out) return set_err(BPP_E_INVALID, "null argument");
    *units_out = e->P.edge_cap;
    return BPP_OK;
}

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int grid_warps(int n) { return (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA; }
#define LAUNCH_CHECK(e)                                                                          \
    do {                                                                                         \
        (e)->launches++;                                                                         \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess) return set_err(BPP_E_CUDA, "kernel launch failed: %s (%s:%d)",    \
                                              cudaGetErrorString(_e), __FILE__, __LINE__);       \
    } while (0)

extern "C" int bpp_engine_reset(bpp_engine* e, const int32_t* items_wh_dev, const int32_t* total_area_dev,
                                const double* bl_dev, const int8_t* tie_dev, void* stream) {
    if (!e || !items_wh_dev || !total_area_dev || !bl_dev) return set_err(BPP_E_INVALID, "null argument");
    Params& P = e->P;
    CUDA_TRY(cudaMemsetAsync(P.table, 0, (size_t)P.G * (P.table_mask + 1u) * sizeof(uint32_t), S(stream)));
    CUDA_TRY(cudaMemsetAsync(P.leaf_count, 0, sizeof(int), S(stream)));
    k_reset<<<(P.G + 127) / 128, 128, 0, S(stream)>>>(P, items_wh_dev, total_area_dev, bl_dev, tie_dev);
    LAUNCH_CHECK(e);
    e->leaf_parked = false;
    // the evaluator reads the item dimensions as int32 [G][N][2]: keep an engine-owned copy (bpp_engine_play_net)
    if (items_wh_dev != e->d_items)
        CUDA_TRY(cudaMemcpyAsync(e->d_items, items_wh_dev, (size_t)P.G * P.geom.N * 2 * sizeof(int32_t),
                                 cudaMemcpyDeviceToDevice, S(stream)));
    e->items_ref = e->d_items;
    return BPP_OK;
}

extern "C" int bpp_engine_reset_host(bpp_engine* e, const int32_t* items_wh_host, const int32_t* total_area_host,
                                     const double* bl_host, const int8_t* tie_host, void* stream) {
    if (!e || !items_wh_host || !total_area_host || !bl_host) return set_err(BPP_E_INVALID, "null argument");
    const size_t G = (size_t)e->P.G, N = (size_t)e->P.geom.N;
    int32_t* d_items = e->d_items;
    int32_t* d_area = e->d_area;
    double* d_bl = e->d_bl;
    int8_t* d_tie = e->d_tie;
    CUDA_TRY(cudaMemcpyAsync(d_items, items_wh_host, G * N * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, S(stream)));
    CUDA_TRY(cudaMemcpyAsync(d_area, total_area_host, G * sizeof(int32_t), cudaMemcpyHostToDevice, S(stream)));
    CUDA_TRY(cudaMemcpyAsync(d_bl, bl_host, G * sizeof(double), cudaMemcpyHostToDevice, S(stream)));
    if (tie_host) CUDA_TRY(cudaMemcpyAsync(d_tie, tie_host, G * sizeof(int8_t), cudaMemcpyHostToDevice, S(stream)));
    return bpp_engine_reset(e, d_items, d_area, d_bl, tie_host ? d_tie : nullptr, stream);
}

extern "C" int bpp_engine_set_roots(bpp_engine* e, const uint32_t* roots_dev, void* stream) {
    if (!e || !roots_dev) return set_err(BPP_E_INVALID, "null argument");
    const int n = e->P.G * REC_WORDS;
    k_set_roots<<<(n + 255) / 256, 256, 0, S(stream)>>>(e->P, roots_dev);
    LAUNCH_CHECK(e);
    e->leaf_parked = false;
    return BPP_OK;
}

__global__ void k_set_max_h(Params P, const int32_t* max_h) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P.G) return;
    const int cdiv = (P.total_area[g] + P.geom.W - 1) / P.geom.W;
    P.numer[g] = cdiv > max_h[g] ? cdiv : max_h[g];
}

extern "C" int bpp_engine_set_max_h(bpp_engine* e, const int32_t* max_h_dev, void* stream) {
    if (!e || !max_h_dev) return set_err(BPP_E_INVALID, "null argument");
    k_set_max_h<<<(e->P.G + 127) / 128, 128, 0, S(stream)>>>(e->P, max_h_dev);
    LAUNCH_CHECK(e);
    return BPP_OK;
}

extern "C" int bpp_engine_set_num_sims(bpp_engine* e, int num_sims) {
    if (!e || num_sims < 0) return set_err(BPP_E_INVALID, "bad argument");
    e->P.num_sims = num_sims;
    return BPP_OK;
}

extern "C" int bpp_engine_set_select_cap(bpp_engine* e, int max_sims_per_launch) {
    if (!e || max_sims_per_launch < 0) return set_err(BPP_E_INVALID, "bad argument");
    e->P.select_cap = max_sims_per_launch;
    return BPP_OK;
}

extern "C" int bpp_engine_unfinished(bpp_engine* e, int32_t* count_host) {
    if (!e || !count_host) return set_err(BPP_E_INVALID, "null argument");
    *count_host = e->unfinished;
    return BPP_OK;
}

extern "C" int bpp_engine_last_values(bpp_engine* e, double* values_out_dev, void* stream) {
    if (!e || !values_out_dev) return set_err(BPP_E_INVALID, "null argument");
    CUDA_TRY(cudaMemcpyAsync(values_out_dev, e->P.last_v, (size_t)e->P.G * sizeof(double), cudaMemcpyDeviceToDevice,
                             S(stream)));
    return BPP_OK;
}

extern "C" int bpp_engine_begin_move(bpp_engine* e, void* stream) {
    if (!e) return set_err(BPP_E_INVALID, "null argument");
    CUDA_TRY(cudaMemsetAsync(e->P.sims_done, 0, (size_t)e->P.G * sizeof(int), S(stream)));
    return BPP_OK;
}

// bin heights with a specialised (fully unrolled, rows-in-registers) valid-move sweep; anything else is generic
template <int STUB>
static void launch_search(bpp_engine* e, void* stream) {
    const int grid = grid_warps(e->P.G), block = WARPS_PER_CTA * 32;
    switch (e->P.geom.H) {
        case 15: k_search<STUB, 15><<<grid, block, 0, S(stream)>>>(e->P); break;
        case 20: k_search<STUB, 20><<<grid, block, 0, S(stream)>>>(e->P); break;
        default: k_search<STUB, 0><<<grid, block, 0, S(stream)>>>(e->P); break;
    }
}

extern "C" int bpp_engine_select(bpp_engine* e, void* stream) {
    if (!e) return set_err(BPP_E_INVALID, "null argument");
    if (e->leaf_parked) return set_err(BPP_E_STATE, "bpp_engine_select called with leaves still parked");
    CUDA_TRY(cudaMemsetAsync(e->P.leaf_count, 0, 4 * sizeof(int), S(stream)));
    launch_search<0>(e, stream);
    LAUNCH_CHECK(e);
    e->leaf_parked = true;
    return BPP_OK;
}

extern "C" int bpp_engine_leaf_count(bpp_engine* e, int32_t* count_host, void* stream) {
    if (!e || !count_host) return set_err(BPP_E_INVALID, "null argument");
    int both[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(both, e->P.leaf_count, 2 * sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
    CUDA_TRY(cudaStreamSynchronize(S(stream)));
    *count_host = both[0];
    e->unfinished = both[1];
    if (both[0] == 0) e->leaf_parked = false;  // nothing to expand: the select/expand pairing is complete
    return BPP_OK;
}

extern "C" int bpp_engine_leaf_count_async(bpp_engine* e, int32_t* counts_host2, void* stream) {
    if (!e || !counts_host2) return set_err(BPP_E_INVALID, "null argument");
    CUDA_TRY(cudaMemcpyAsync(counts_host2, e->P.leaf_count, 2 * sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
    return BPP_OK;
}

extern "C" int bpp_engine_leaf_buffers(bpp_engine* e, const int32_t** count_dev, const int32_t** game_dev,
                                       const uint32_t** recs_dev) {
    if (!e) return set_err(BPP_E_INVALID, "null argument");
    if (count_dev) *count_dev = e->P.leaf_count;
    if (game_dev) *game_dev = e->P.leaf_game;
    if (recs_dev) *recs_dev = e->P.leaf_rec;
    return BPP_OK;
}

extern "C" int bpp_engine_leaf_planes(bpp_engine* e, float* planes_out_dev, void* stream) {
    if (!e || !planes_out_dev) return set_err(BPP_E_INVALID, "null argument");
    k_leaf_planes<<<2 * e->num_sms, 256, 0, S(stream)>>>(e->P, planes_out_dev);
    LAUNCH_CHECK(e);
    return BPP_OK;
}

extern "C" int bpp_engine_expand_backup(bpp_engine* e, const void* policy_dev, int policy_dtype,
                                        const void* value_dev, int value_dtype, void* stream) {
    if (!e || !policy_dev || !value_dev) return set_err(BPP_E_INVALID, "null argument");
    if (!e->leaf_parked) return set_err(BPP_E_STATE, "bpp_engine_expand_backup without a preceding bpp_engine_select");
    k_expand_backup<<<grid_warps(e->P.G), WARPS_PER_CTA * 32, 0, S(stream)>>>(
        e->P, policy_dev, policy_dtype == BPP_DTYPE_F64, value_dev, value_dtype == BPP_DTYPE_F64);
    LAUNCH_CHECK(e);
    e->leaf_parked = false;
    return BPP_OK;
}

extern "C" int bpp_engine_expand_select(bpp_engine* e, const void* policy_dev, int policy_dtype, const void* value_dev,
                                        int value_dtype, void* stream) {
    if (!e || !policy_dev || !value_dev) return set_err(BPP_E_INVALID, "null argument");
    CUDA_TRY(cudaMemsetAsync(e->P.leaf_count, 0, 4 * sizeof(int), S(stream)));
    const int grid = grid_warps(e->P.G), block = WARPS_PER_CTA * 32;
    const int pf = policy_dtype == BPP_DTYPE_F64, vf = value_dtype == BPP_DTYPE_F64;
    switch (e->P.geom.H) {
        case 15: k_expand_search<15><<<grid, block, 0, S(stream)>>>(e->P, policy_dev, pf, value_dev, vf); break;
        case 20: k_expand_search<20><<<grid, block, 0, S(stream)>>>(e->P, policy_dev, pf, value_dev, vf); break;
        default: k_expand_search<0><<<grid, block, 0, S(stream)>>>(e->P, policy_dev, pf, value_dev, vf); break;
    }
    LAUNCH_CHECK(e);
    e->leaf_parked = true;
    return BPP_OK;
}

extern "C" int bpp_engine_search_stub(bpp_engine* e, int stub_kind, void* stream) {
    if (!e) return set_err(BPP_E_INVALID, "null argument");
    if (e->leaf_parked) return set_err(BPP_E_STATE, "leaves are parked; call bpp_engine_expand_backup first");
    switch (stub_kind) {
        case BPP_STUB_U: launch_search<1>(e, stream); break;
        case BPP_STUB_V: launch_search<2>(e, stream); break;
        case BPP_STUB_H: launch_search<3>(e, stream); break;
        case BPP_STUB_D: launch_search<4>(e, stream); break;
        default: return set_err(BPP_E_INVALID, "unknown stub kind %d", stub_kind);
    }
    LAUNCH_CHECK(e);
    return BPP_OK;
}

extern "C" int bpp_engine_root_counts(bpp_engine* e, int32_t* counts_out_dev, void* stream) {
    if (!e || !counts_out_dev) return set_err(BPP_E_INVALID, "null argument");
    k_root_counts<<<grid_warps(e->P.G), WARPS_PER_CTA * 32, 0, S(stream)>>>(e->P, counts_out_dev);
    LAUNCH_CHECK(e);
    return BPP_OK;
}

extern "C" int bpp_engine_root_counts_host(bpp_engine* e, int32_t* counts_out_host, void* stream) {
    if (!e || !counts_out_host) return set_err(BPP_E_INVALID, "null argument");
    const size_t n = (size_t)e->P.G * e->P.geom.A;
    int32_t* d_counts = e->d_counts;
    int rc = bpp_engine_root_counts(e, d_counts, stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(counts_out_host, d_counts, n * sizeof(int32_t), cudaMemcpyDeviceToHost, S(stream)));
    CUDA_TRY(cudaStreamSynchronize(S(stream)));
    return BPP_OK;
}

extern "C" int bpp_engine_choose(bpp_engine* e, int mode, uint64_t seed, int32_t* actions_out_dev, void* stream) {
    if (!e || !actions_out_dev) return set_err(BPP_E_INVALID, "null argument");
    if (mode < 0 || mode > 2) return set_err(BPP_E_INVALID, "unknown choose mode %d", mode);
    k_choose<<<(e->P.G + 127) / 128, 128, 0, S(stream)>>>(e->P, mode, (unsigned long long)seed, actions_out_dev);
    LAUNCH_CHECK(e);
    return BPP_OK;
}

extern "C" int bpp_engine_advance(bpp_engine* e, const int32_t* actions_dev, void* stream) {
    if (!e || !actions_dev) return set_err(BPP_E_INVALID, "null argument");
    if (e->leaf_parked) return set_err(BPP_E_STATE, "leaves are parked; call bpp_engine_expand_backup first");
    k_advance<<<grid_warps(e->P.G), WARPS_PER_CTA * 32, 0, S(stream)>>>(e->P, actions_dev);
    LAUNCH_CHECK(e);
    return BPP_OK;
}

extern "C" int bpp_engine_status(bpp_engine* e, int32_t* done_out_dev, int32_t* r_out_dev, double* score_out_dev,
                                 int32_t* moves_out_dev, void* stream) {
    if (!e) return set_err(BPP_E_INVALID, "null argument");
    const size_t G = (size_t)e->P.G;
    if (done_out_dev)
        CUDA_TRY(cudaMemcpyAsync(done_out_dev, e->P.status, G * sizeof(int), cudaMemcpyDeviceToDevice, S(stream)));
    if (r_out_dev) CUDA_TRY(cudaMemcpyAsync(r_out_dev, e->P.ep_r, G * sizeof(int), cudaMemcpyDeviceToDevice, S(stream)));
    if (score_out_dev)
        CUDA_TRY(cudaMemcpyAsync(score_out_dev, e->P.ep_score, G * sizeof(double), cudaMemcpyDeviceToDevice, S(stream)));
    if (moves_out_dev)
        CUDA_TRY(cudaMemcpyAsync(moves_out_dev, e->P.moves_done, G * sizeof(int), cudaMemcpyDeviceToDevice, S(stream)));
    return BPP_OK;
}

extern "C" int bpp_engine_roots(bpp_engine* e, uint32_t* roots_out_dev, void* stream) {
    if (!e || !roots_out_dev) return set_err(BPP_E_INVALID, "null argument");
    CUDA_TRY(cudaMemcpyAsync(roots_out_dev, e->P.root_rec, (size_t)e->P.G * REC_WORDS * sizeof(uint32_t),
                             cudaMemcpyDeviceToDevice, S(stream)));
    return BPP_OK;
}

template <int STUB>
static void launch_episode(bpp_engine* e, int mode, uint64_t seed, int moves, int32_t* counts, int32_t* actions,
                           int32_t* counts_host, void* stream) {
    const int grid = grid_warps(e->P.G), block = WARPS_PER_CTA * 32;
    const unsigned long long sd = (unsigned long long)seed;
    switch (e->P.geom.H) {
        case 15: k_episode<STUB, 15><<<grid, block, 0, S(stream)>>>(e->P, mode, sd, moves, counts, actions, counts_host); break;
        case 20: k_episode<STUB, 20><<<grid, block, 0, S(stream)>>>(e->P, mode, sd, moves, counts, actions, counts_host); break;
        default: k_episode<STUB, 0><<<grid, block, 0, S(stream)>>>(e->P, mode, sd, moves, counts, actions, counts_host); break;
    }
}

// counts_host_map != nullptr: results go to mapped host memory (counts_out_dev is the device scratch, actions_out_dev
// the host alias) and the kernel itself writes the rows of unplayed moves
static int play_stub_impl(bpp_engine* e, int stub_kind, int choose_mode, uint64_t seed, int max_moves,
                          int32_t* counts_out_dev, int32_t* actions_out_dev, int32_t* moves_run_host,
                          int32_t* counts_host_map, void* stream) {
    if (!e) return set_err(BPP_E_INVALID, "null argument");
    if (e->leaf_parked) return set_err(BPP_E_STATE, "leaves are parked; call bpp_engine_expand_backup first");
    if (choose_mode < 0 || choose_mode > 2) return set_err(BPP_E_INVALID, "unknown choose mode %d", choose_mode);
    // every move places exactly one item, so an episode has at most N moves
    const int moves = max_moves > 0 && max_moves < e->P.geom.N ? max_moves : e->P.geom.N;
    const size_t G = e->P.q_total ? (size_t)e->P.q_total : (size_t)e->P.G, GA = G * e->P.geom.A;  // rows per move
    // rows of moves a game does not play: counts 0, action -1
    if (!counts_host_map) {
        if (counts_out_dev) CUDA_TRY(cudaMemsetAsync(counts_out_dev, 0, (size_t)moves * GA * sizeof(int32_t), S(stream)));
        if (actions_out_dev)
            CUDA_TRY(cudaMemsetAsync(actions_out_dev, 0xff, (size_t)moves * G * sizeof(int32_t), S(stream)));
    }
    int32_t* f = counts_host_map;
    switch (stub_kind) {
        case BPP_STUB_U: launch_episode<1>(e, choose_mode, seed, moves, counts_out_dev, actions_out_dev, f, stream); break;
        case BPP_STUB_V: launch_episode<2>(e, choose_mode, seed, moves, counts_out_dev, actions_out_dev, f, stream); break;
        case BPP_STUB_H: launch_episode<3>(e, choose_mode, seed, moves, counts_out_dev, actions_out_dev, f, stream); break;
        case BPP_STUB_D: launch_episode<4>(e, choose_mode, seed, moves, counts_out_dev, actions_out_dev, f, stream); break;
        default: return set_err(BPP_E_INVALID, "unknown stub kind %d", stub_kind);
    }
    LAUNCH_CHECK(e);
    if (moves_run_host) *moves_run_host = moves;
    return BPP_OK;
}

extern "C" int bpp_engine_play_stub(bpp_engine* e, int stub_kind, int choose_mode, uint64_t seed, int max_moves,
                                    int32_t* counts_out_dev, int32_t* actions_out_dev, int32_t* moves_run_host,
                                    void* stream) {
    return play_stub_impl(e, stub_kind, choose_mode, seed, max_moves, counts_out_dev, actions_out_dev, moves_run_host,
                          nullptr, stream);
}

// device alias of a host buffer the kernels can write directly (pinned + mapped: cudaHostAlloc / cudaHostRegister under
// unified addressing), or nullptr for pageable memory
static void* mapped_alias(void* host_ptr) {
    if (!host_ptr) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host_ptr) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

extern "C" int bpp_engine_play_stub_host(bpp_engine* e, int stub_kind, int choose_mode, uint64_t seed,
                                         const int32_t* items_wh_host, const int32_t* total_area_host,
                                         const double* bl_host, const int8_t* tie_host, int32_t* counts_out_host,
                                         int32_t* actions_out_host, int32_t* r_out_host, double* score_out_host,
                                         int32_t* moves_out_host, void* stream) {
    if (!e) return set_err(BPP_E_INVALID, "null argument");
    const size_t G = (size_t)e->P.G, A = (size_t)e->P.geom.A, N = (size_t)e->P.geom.N;
    int rc;
    // Pinned (mapped) result buffers are written by the episode kernel itself, row by row while the games run: the
    // 25 MB of visit counts then cross PCIe under the kernel instead of in a copy after it.  Pageable buffers take the
    // staged path (device buffer + copy).
    int32_t* counts_map = static_cast<int32_t*>(mapped_alias(counts_out_host));
    int32_t* actions_map = static_cast<int32_t*>(mapped_alias(actions_out_host));
    const bool direct = counts_map && (!actions_out_host || actions_map) && getenv("BPP_NO_ZERO_COPY") == nullptr;
    if (counts_out_host && !e->d_counts_all) {
        if ((rc = dev_alloc(e, &e->d_counts_all, N * G * A))) return rc;
    }
    if (!direct && !e->d_actions_all && (rc = dev_alloc(e, &e->d_actions_all, N * G))) return rc;
    if ((rc = bpp_engine_reset_host(e, items_wh_host, total_area_host, bl_host, tie_host, stream))) return rc;
    if (direct) {
        if ((rc = play_stub_impl(e, stub_kind, choose_mode, seed, 0, e->d_counts_all, actions_map, nullptr, counts_map,
                                 stream)))
            return rc;
    } else {
        if ((rc = bpp_engine_play_stub(e, stub_kind, choose_mode, seed, 0, counts_out_host ? e->d_counts_all : nullptr,
                                       e->d_actions_all, nullptr, stream)))
            return rc;
        if (counts_out_host)
            CUDA_TRY(cudaMemcpyAsync(counts_out_host, e->d_counts_all, N * G * A * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                     S(stream)));
        if (actions_out_host)
            CUDA_TRY(cudaMemcpyAsync(actions_out_host, e->d_actions_all, N * G * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                     S(stream)));
    }
    if (r_out_host)
        CUDA_TRY(cudaMemcpyAsync(r_out_host, e->P.ep_r, G * sizeof(int32_t), cudaMemcpyDeviceToHost, S(stream)));
    if (score_out_host)
        CUDA_TRY(cudaMemcpyAsync(score_out_host, e->P.ep_score, G * sizeof(double), cudaMemcpyDeviceToHost, S(stream)));
    if (moves_out_host)
        CUDA_TRY(cudaMemcpyAsync(moves_out_host, e->P.moves_done, G * sizeof(int32_t), cudaMemcpyDeviceToHost, S(stream)));
    CUDA_TRY(cudaStreamSynchronize(S(stream)));
    return bpp_engine_check(e, stream);
}

// queue / result staging of the *_stream_host entry points, grown on demand
template <typename T>
static int ensure(bpp_engine* e, T** p, size_t* have, size_t need) {
    if (*p && *have >= need) return BPP_OK;
    *have = need;
    return dev_alloc(e, p, need);  // the old block stays owned by the handle until destroy (rare: sizes repeat)
}

// E >= 1 episodes streamed through the G resident games with a stub evaluator, ONE launch: a game whose episode ends
// latches its outcome and takes the next instance of the queue inside k_episode (CoachBPP.executeEpisode x E; the action
// stream is keyed by (seed, episode, move), so the results do not depend on G).  Outputs have E rows per move.
static int play_stub_stream_impl(bpp_engine* e, int stub_kind, int choose_mode, uint64_t seed, int num_episodes,
                                 const int32_t* items_wh_dev, const int32_t* total_area_dev, const double* bl_dev,
                                 const int8_t* tie_dev, int32_t* counts_out_dev, int32_t* actions_out_dev,
                                 int32_t* r_out_dev, double* score_out_dev, int32_t* moves_out_dev,
                                 int32_t* counts_host_map, void* stream) {
    if (!e || !items_wh_dev || !total_area_dev || !bl_dev) return set_err(BPP_E_INVALID, "null argument");
    if (num_episodes < 1) return set_err(BPP_E_INVALID, "num_episodes = %d", num_episodes);
    Params& P = e->P;
    P.q_total = num_episodes;
    P.q_items = items_wh_dev;
    P.q_area = total_area_dev;
    P.q_bl = bl_dev;
    P.q_tie = tie_dev;
    P.epq_r = r_out_dev;
    P.epq_score = score_out_dev;
    P.epq_moves = moves_out_dev;
    const int first = P.G;  // episodes 0..G-1 start in the games, the queue hands out the rest
    cudaError_t ce = cudaMemcpyAsync(P.q_next, &first, sizeof(int), cudaMemcpyHostToDevice, S(stream));
    int rc = BPP_OK;
    if (ce != cudaSuccess) {
        rc = set_err(BPP_E_CUDA, "bpp_engine_play_stub_stream: %s", cudaGetErrorString(ce));
    } else {
        k_stream_begin<<<grid_warps(P.G), WARPS_PER_CTA * 32, 0, S(stream)>>>(P);
        e->launches++;
        e->leaf_parked = false;
        e->items_ref = e->d_items;
        rc = play_stub_impl(e, stub_kind, choose_mode, seed, 0, counts_out_dev, actions_out_dev, nullptr, counts_host_map,
                            stream);
    }
    P.q_total = 0;
    P.q_items = nullptr; P.q_area = nullptr; P.q_bl = nullptr; P.q_tie = nullptr;
    P.epq_r = nullptr; P.epq_score = nullptr; P.epq_moves = nullptr;
    return rc;
}

extern "C" int bpp_engine_play_stub_stream(bpp_engine* e, int stub_kind, int choose_mode, uint64_t seed, int num_episodes,
                                           const int32_t* items_wh_dev, const int32_t* total_area_dev, const double* bl_dev,
                                           const int8_t* tie_dev, int32_t* counts_out_dev, int32_t* actions_out_dev,
                                           int32_t* r_out_dev, double* score_out_dev, int32_t* moves_out_dev, void* stream) {
    return play_stub_stream_impl(e, stub_kind, choose_mode, seed, num_episodes, items_wh_dev, total_area_dev, bl_dev, tie_dev,
                                 counts_out_dev, actions_out_dev, r_out_dev, score_out_dev, moves_out_dev, nullptr, stream);
}

extern "C" int bpp_engine_play_stub_stream_host(bpp_engine* e, int stub_kind, int choose_mode, uint64_t seed,
                                                int num_episodes, const int32_t* items_wh_host,
                                                const int32_t* total_area_host, const double* bl_host, const int8_t* tie_host,
                                                int32_t* counts_out_host, int32_t* actions_out_host, int32_t* r_out_host,
                                                double* score_out_host, int32_t* moves_out_host, void* stream) {
    if (!e || !items_wh_host || !total_area_host || !bl_host) return set_err(BPP_E_INVALID, "null argument");
    if (num_episodes < 1) return set_err(BPP_E_INVALID, "num_episodes = %d", num_episodes);
    const size_t E = (size_t)num_episodes, G = (size_t)e->P.G, A = (size_t)e->P.geom.A, N = (size_t)e->P.geom.N;
    int rc;
    if ((rc = ensure(e, &e->q_items, &e->q_items_n, E * N * 2)) || (rc = ensure(e, &e->q_area, &e->q_area_n, E)) ||
        (rc = ensure(e, &e->q_bl, &e->q_bl_n, E)) || (rc = ensure(e, &e->q_tie, &e->q_tie_n, E)) ||
        (rc = ensure(e, &e->q_r, &e->q_r_n, E)) || (rc = ensure(e, &e->q_score, &e->q_score_n, E)) ||
        (rc = ensure(e, &e->q_moves, &e->q_moves_n, E)))
        return rc;
    // Pinned (mapped) result buffers are written by the episode kernel itself, row by row while the games run (the rows
    // are built in a per-game device scratch of N x G x A); pageable buffers take the staged path.
    int32_t* counts_map = static_cast<int32_t*>(mapped_alias(counts_out_host));
    int32_t* actions_map = static_cast<int32_t*>(mapped_alias(actions_out_host));
    const bool direct = counts_map && (!actions_out_host || actions_map) && getenv("BPP_NO_ZERO_COPY") == nullptr;
    if (counts_out_host && (rc = ensure(e, &e->q_counts, &e->q_counts_n, N * E * A))) return rc;
    if (!direct && actions_out_host && (rc = ensure(e, &e->q_actions, &e->q_actions_n, N * E))) return rc;
    (void)G;
    CUDA_TRY(cudaMemcpyAsync(e->q_items, items_wh_host, E * N * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, S(stream)));
    CUDA_TRY(cudaMemcpyAsync(e->q_area, total_area_host, E * sizeof(int32_t), cudaMemcpyHostToDevice, S(stream)));
    CUDA_TRY(cudaMemcpyAsync(e->q_bl, bl_host, E * sizeof(double), cudaMemcpyHostToDevice, S(stream)));
    if (tie_host) CUDA_TRY(cudaMemcpyAsync(e->q_tie, tie_host, E * sizeof(int8_t), cudaMemcpyHostToDevice, S(stream)));
    CUDA_TRY(cudaMemsetAsync(e->q_moves, 0, E * sizeof(int32_t), S(stream)));
    if (direct) {
        rc = play_stub_stream_impl(e, stub_kind, choose_mode, seed, num_episodes, e->q_items, e->q_area, e->q_bl,
                                   tie_host ? e->q_tie : nullptr, e->q_counts, actions_map, e->q_r, e->q_score, e->q_moves,
                                   counts_map, stream);
        if (rc) return rc;
    } else {
        rc = play_stub_stream_impl(e, stub_kind, choose_mode, seed, num_episodes, e->q_items, e->q_area, e->q_bl,
                                   tie_host ? e->q_tie : nullptr, counts_out_host ? e->q_counts : nullptr,
                                   actions_out_host ? e->q_actions : nullptr, e->q_r, e->q_score, e->q_moves, nullptr,
                                   stream);
        if (rc) return rc;
        if (counts_out_host)
            CUDA_TRY(cudaMemcpyAsync(counts_out_host, e->q_counts, N * E * A * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                     S(stream)));
        if (actions_out_host)
            CUDA_TRY(cudaMemcpyAsync(actions_out_host, e->q_actions, N * E * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                     S(stream)));
    }
    if (r_out_host) CUDA_TRY(cudaMemcpyAsync(r_out_host, e->q_r, E * sizeof(int32_t), cudaMemcpyDeviceToHost, S(stream)));
    if (score_out_host)
        CUDA_TRY(cudaMemcpyAsync(score_out_host, e->q_score, E * sizeof(double), cudaMemcpyDeviceToHost, S(stream)));
    if (moves_out_host)
        CUDA_TRY(cudaMemcpyAsync(moves_out_host, e->q_moves, E * sizeof(int32_t), cudaMemcpyDeviceToHost, S(stream)));
    CUDA_TRY(cudaStreamSynchronize(S(stream)));
    return bpp_engine_check(e, stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// Asynchronous episodes with an external (batched, device) evaluator
extern "C" int bpp_engine_set_auto_play(bpp_engine* e, int choose_mode, uint64_t seed, int32_t* counts_out_dev,
                                        int32_t* actions_out_dev, uint32_t* roots_out_dev) {
    if (!e) return set_err(BPP_E_INVALID, "null argument");
    if (choose_mode < -1 || choose_mode > 2) return set_err(BPP_E_INVALID, "unknown choose mode %d", choose_mode);
    e->P.auto_mode = choose_mode;
    e->P.auto_seed = (unsigned long long)seed;
    e->P.ep_counts = choose_mode < 0 ? nullptr : counts_out_dev;
    e->P.ep_actions = choose_mode < 0 ? nullptr : actions_out_dev;
    e->P.ep_roots = choose_mode < 0 ? nullptr : roots_out_dev;
    return BPP_OK;
}

extern "C" int bpp_engine_progress_async(bpp_engine* e, int32_t* counts_host4, void* stream) {
    if (!e || !counts_host4) return set_err(BPP_E_INVALID, "null argument");
    CUDA_TRY(cudaMemcpyAsync(counts_host4, e->P.leaf_count, 4 * sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
    return BPP_OK;
}

extern "C" int bpp_net_forward(bpp_net* n, int B, const int32_t* count_dev, const uint32_t* recs_dev,
                               const int32_t* game_dev, const int32_t* items_wh_dev, float* policy_out_dev,
                               float* value_out_dev, void* stream);

// The lockstep loop of bpp_engine_play_net / bpp_engine_play_net_stream: games reset (or the stream begun), outputs with
// `rows` entries per move (G games, or E episodes of a stream).
static int run_lockstep(bpp_engine* e, bpp_net* net, int choose_mode, uint64_t seed, int32_t* counts_out_dev,
                        int32_t* actions_out_dev, uint32_t* roots_out_dev, size_t rows, int32_t* steps_run_host,
                        void* stream) {
    Params& P = e->P;
    const size_t G = (size_t)P.G, A = (size_t)P.geom.A, N = (size_t)P.geom.N;
    int rc;
    if (!e->d_pol && ((rc = dev_alloc(e, &e->d_pol, G * A)) || (rc = dev_alloc(e, &e->d_val, G)))) return rc;
    // rows of moves a game does not play: counts 0, action -1, root record 0
    if (counts_out_dev) CUDA_TRY(cudaMemsetAsync(counts_out_dev, 0, N * rows * A * sizeof(int32_t), S(stream)));
    if (actions_out_dev) CUDA_TRY(cudaMemsetAsync(actions_out_dev, 0xff, N * rows * sizeof(int32_t), S(stream)));
    if (roots_out_dev) CUDA_TRY(cudaMemsetAsync(roots_out_dev, 0, N * rows * REC_WORDS * sizeof(uint32_t), S(stream)));
    // Work per game and launch.  A launch lasts as long as its slowest warp, and a game whose simulations end on terminal
    // states parks no leaf, so its work per launch is bounded - in EDGES WALKED (a late-game simulation is 1-2 edges deep,
    // an early one 5-10), not in simulations.  The bound adapts to the leaf supply: with a sharp trained policy only ~2 %
    // of the simulations need the evaluator; with a flat prior most do.  The bound is therefore tuned while the stream
    // runs, by hill climbing on the measured simulations per device-second of the last chunk.
    const int keep_cap = P.select_cap;
    const char* bud_env = getenv("BPP_EDGE_BUDGET");
    const bool adapt = bud_env == nullptr;
    int budget = bud_env ? atoi(bud_env) : 12;
    if (budget < 0) budget = 0;
    P.select_cap = 0;
    P.edge_budget = budget;
    if ((rc = bpp_engine_set_auto_play(e, choose_mode, seed, counts_out_dev, actions_out_dev, roots_out_dev))) return rc;
    auto restore = [&]() {
        P.select_cap = keep_cap;
        P.edge_budget = 0;
        bpp_engine_set_auto_play(e, -1, 0, nullptr, nullptr, nullptr);
    };
    if ((rc = bpp_engine_begin_move(e, stream)) || (rc = bpp_engine_select(e, stream))) {
        restore();
        return rc;
    }
    // Lockstep steps are queued in chunks; the progress counters of chunk k are read back asynchronously and looked at
    // after chunk k+1 has been queued, so the GPU never waits for the host.  Steps queued after the last game has ended
    // are no-ops (no parked leaf, every game's status != 0).
    const int chunk = 8;
    int steps = 0;
    size_t nev = 0;
    int used[4] = {budget, budget, budget, budget};   // edge budget the chunks of the progress ring were queued with
    int dir = 1, acc_n = 0;
    double last_rate = 0.0, acc_sims = 0.0, acc_ms = 0.0;
    const bool trace = getenv("BPP_PLAY_TRACE") != nullptr;
    auto mark = [&]() {  // profiling pass only: one event per call boundary
        if (!e->prof_on) return;
        if (nev == e->prof_ev.size()) {
            cudaEvent_t ev;
            if (cudaEventCreate(&ev) != cudaSuccess) return;
            e->prof_ev.push_back(ev);
        }
        cudaEventRecord(e->prof_ev[nev++], S(stream));
    };
    for (int k = 0;; ++k) {
        for (int i = 0; i < chunk; ++i) {
            mark();
            rc = bpp_net_forward(net, P.G, P.leaf_count, P.leaf_rec, P.leaf_game, e->items_ref, e->d_pol, e->d_val, stream);
            mark();
            if (!rc) rc = bpp_engine_expand_select(e, e->d_pol, BPP_DTYPE_F32, e->d_val, BPP_DTYPE_F32, stream);
            mark();
            if (rc) {
                restore();
                return rc;
            }
        }
        steps += chunk;
        int32_t* hp = e->h_prog + 8 * (k & 3);
        cudaError_t ce = cudaMemcpyAsync(hp, P.leaf_count, 4 * sizeof(int), cudaMemcpyDeviceToHost, S(stream));
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(hp + 4, P.stats, sizeof(unsigned long long), cudaMemcpyDeviceToHost, S(stream));
        if (ce == cudaSuccess) ce = cudaEventRecord(e->ev_prog[k & 3], S(stream));
        used[k & 3] = budget;
        if (ce == cudaSuccess && k > 0) {
            ce = cudaEventSynchronize(e->ev_prog[(k - 1) & 3]);
            const int32_t* pp = e->h_prog + 8 * ((k - 1) & 3);
            if (trace) fprintf(stderr, "play_net chunk %d: leaves %d capped %d running %d budget %d rate %.1f M sims/s\n", k - 1,
                               pp[0], pp[1], pp[2], used[(k - 1) & 3], last_rate * 1e-3);
            if (ce == cudaSuccess && pp[0] == 0 && pp[2] == 0) break;  // nothing parked, nobody playing
            if (adapt && k > 1 && pp[2] > 0 && ce == cudaSuccess) {
                // Edge budget control.  (1) While fewer than 15 % of the running games stop at the budget there is nothing
                // to gain from a larger one (nearly every game parks a leaf first): shrink towards the floor.  (2) Otherwise
                // hill-climb on the measured throughput: simulations completed (device counter) per device time (events)
                // over the last TWO chunks that ran entirely with the current setting; a new setting takes effect one chunk
                // after it is chosen, because the next chunk is already queued.
                float dt = 0.f;
                unsigned long long s1, s0;
                memcpy(&s1, pp + 4, 8);
                memcpy(&s0, e->h_prog + 8 * ((k - 2) & 3) + 4, 8);
                const bool same = used[(k - 1) & 3] == budget;
                if (same && cudaEventElapsedTime(&dt, e->ev_prog[(k - 2) & 3], e->ev_prog[(k - 1) & 3]) == cudaSuccess && dt > 0.f) {
                    acc_sims += (double)(s1 - s0);
                    acc_ms += dt;
                    acc_n++;
                } else if (!same) {
                    acc_sims = acc_ms = 0.0;
                    acc_n = 0;
                }
                // (steps of 1.5x: with steps of 2x the climber spent half its time one step off an optimum that is flat over
                // about a factor of two - 20x20, flat prior: 35.7 / 36.4 / 36.0 M sims/s at budgets 8 / 12 / 24)
                if (pp[1] * 100 < pp[2] * 15) {
                    if (same && budget > 8) {
                        budget = budget * 2 / 3;
                        if (budget < 8) budget = 8;
                        dir = -1;
                        last_rate = 0.0;
                    }
                } else if (acc_n >= 2) {
                    const double rate = acc_sims / acc_ms;  // simulations per ms
                    if (last_rate > 0.0 && rate < 0.98 * last_rate) dir = -dir;
                    last_rate = rate;
                    int nb = dir > 0 ? budget * 3 / 2 : budget * 2 / 3;
                    if (nb < 8) { nb = 8; dir = 1; }
                    if (nb > 8192) { nb = 8192; dir = -1; }
                    budget = nb;
                }
                if (budget != P.edge_budget) {
                    P.edge_budget = budget;
                    acc_sims = acc_ms = 0.0;
                    acc_n = 0;
                }
            }
        }
        if (ce != cudaSuccess) {
            restore();
            return set_err(BPP_E_CUDA, "bpp_engine_play_net: %s", cudaGetErrorString(ce));
        }
        if ((double)steps > (400.0 * P.num_sims * (double)N + 4096.0) * (double)((rows + G - 1) / G)) {  // cannot happen
            restore();
            return set_err(BPP_E_STATE, "bpp_engine_play_net made no progress");
        }
    }
    restore();
    CUDA_TRY(cudaStreamSynchronize(S(stream)));
    e->leaf_parked = false;
    if (steps_run_host) *steps_run_host = steps;
    if (e->prof_on) {
        for (size_t i = 0; i + 2 < nev; i += 3) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, e->prof_ev[i], e->prof_ev[i + 1]);
            cudaEventElapsedTime(&b, e->prof_ev[i + 1], e->prof_ev[i + 2]);
            e->prof_ms[0] += a;
            e->prof_ms[1] += b;
            e->prof_ms[2] += 1.0;
        }
    }
    return BPP_OK;
}

extern "C" int bpp_engine_play_net(bpp_engine* e, bpp_net* net, int choose_mode, uint64_t seed, int32_t* counts_out_dev,
                                   int32_t* actions_out_dev, uint32_t* roots_out_dev, int32_t* steps_run_host,
                                   void* stream) {
    if (!e || !net) return set_err(BPP_E_INVALID, "null argument");
    if (e->leaf_parked) return set_err(BPP_E_STATE, "leaves are parked; call bpp_engine_expand_backup first");
    if (choose_mode < 0 || choose_mode > 2) return set_err(BPP_E_INVALID, "unknown choose mode %d", choose_mode);
    if (!e->items_ref) return set_err(BPP_E_STATE, "bpp_engine_play_net before bpp_engine_reset");
    e->P.q_total = 0;
    return run_lockstep(e, net, choose_mode, seed, counts_out_dev, actions_out_dev, roots_out_dev, (size_t)e->P.G,
                        steps_run_host, stream);
}

extern "C" int bpp_engine_play_net_stream(bpp_engine* e, bpp_net* net, int choose_mode, uint64_t seed, int num_episodes,
                                          const int32_t* items_wh_dev, const int32_t* total_area_dev, const double* bl_dev,
                                          const int8_t* tie_dev, int32_t* counts_out_dev, int32_t* actions_out_dev,
                                          uint32_t* roots_out_dev, int32_t* r_out_dev, double* score_out_dev,
                                          int32_t* moves_out_dev, int32_t* steps_run_host, void* stream) {
    if (!e || !net || !items_wh_dev || !total_area_dev || !bl_dev) return set_err(BPP_E_INVALID, "null argument");
    if (num_episodes < 1) return set_err(BPP_E_INVALID, "num_episodes = %d", num_episodes);
    if (choose_mode < 0 || choose_mode > 2) return set_err(BPP_E_INVALID, "unknown choose mode %d", choose_mode);
    Params& P = e->P;
    P.q_total = num_episodes;
    P.q_items = items_wh_dev;
    P.q_area = total_area_dev;
    P.q_bl = bl_dev;
    P.q_tie = tie_dev;
    P.epq_r = r_out_dev;
    P.epq_score = score_out_dev;
    P.epq_moves = moves_out_dev;
    const int first = P.G;  // episodes 0..G-1 start in the games, the queue hands out the rest
    cudaError_t ce = cudaMemcpyAsync(P.q_next, &first, sizeof(int), cudaMemcpyHostToDevice, S(stream));
    if (ce == cudaSuccess) ce = cudaMemsetAsync(P.leaf_count, 0, 4 * sizeof(int), S(stream));
    if (ce != cudaSuccess) {
        P.q_total = 0;
        return set_err(BPP_E_CUDA, "bpp_engine_play_net_stream: %s", cudaGetErrorString(ce));
    }
    k_stream_begin<<<grid_warps(P.G), WARPS_PER_CTA * 32, 0, S(stream)>>>(P);
    e->launches++;
    e->leaf_parked = false;
    e->items_ref = e->d_items;
    int rc = run_lockstep(e, net, choose_mode, seed, counts_out_dev, actions_out_dev, roots_out_dev, (size_t)num_episodes,
                          steps_run_host, stream);
    P.q_total = 0;
    P.q_items = nullptr; P.q_area = nullptr; P.q_bl = nullptr; P.q_tie = nullptr;
    P.epq_r = nullptr; P.epq_score = nullptr; P.epq_moves = nullptr;
    return rc;
}

extern "C" int bpp_engine_set_profile(bpp_engine* e, int on) {
    if (!e) return set_err(BPP_E_INVALID, "null argument");
    e->prof_on = on != 0;
    for (double& v : e->prof_ms) v = 0.0;
    return BPP_OK;
}

extern "C" int bpp_engine_profile(bpp_engine* e, double ms_out4[4]) {
    if (!e || !ms_out4) return set_err(BPP_E_INVALID, "null argument");
    for (int i = 0; i < 4; ++i) ms_out4[i] = e->prof_ms[i];
    return BPP_OK;
}

extern "C" int bpp_engine_play_net_host(bpp_engine* e, bpp_net* net, int choose_mode, uint64_t seed,
                                        const int32_t* items_wh_host, const int32_t* total_area_host,
                                        const double* bl_host, const int8_t* tie_host, uint32_t* roots_out_host,
                                        int32_t* counts_out_host, int32_t* actions_out_host, int32_t* r_out_host,
                                        double* score_out_host, int32_t* moves_out_host, int32_t* steps_run_host,
                                        void* stream) {
    if (!e || !net) return set_err(BPP_E_INVALID, "null argument");
    const size_t G = (size_t)e->P.G, A = (size_t)e->P.geom.A, N = (size_t)e->P.geom.N;
    int rc;
    if (counts_out_host && !e->d_counts_all && (rc = dev_alloc(e, &e->d_counts_all, N * G * A))) return rc;
    if (actions_out_host && !e->d_actions_all && (rc = dev_alloc(e, &e->d_actions_all, N * G))) return rc;
    if (roots_out_host && !e->d_roots_all && (rc = dev_alloc(e, &e->d_roots_all, N * G * REC_WORDS))) return rc;
    if ((rc = bpp_engine_reset_host(e, items_wh_host, total_area_host, bl_host, tie_host, stream))) return rc;
    if ((rc = bpp_engine_play_net(e, net, choose_mode, seed, counts_out_host ? e->d_counts_all : nullptr,
                                  actions_out_host ? e->d_actions_all : nullptr,
                                  roots_out_host ? e->d_roots_all : nullptr, steps_run_host, stream)))
        return rc;
    if (roots_out_host)
        CUDA_TRY(cudaMemcpyAsync(roots_out_host, e->d_roots_all, N * G * REC_WORDS * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, S(stream)));
    if (counts_out_host)
        CUDA_TRY(cudaMemcpyAsync(counts_out_host, e->d_counts_all, N * G * A * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                 S(stream)));
    if (actions_out_host)
        CUDA_TRY(cudaMemcpyAsync(actions_out_host, e->d_actions_all, N * G * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                 S(stream)));
    if (r_out_host)
        CUDA_TRY(cudaMemcpyAsync(r_out_host, e->P.ep_r, G * sizeof(int32_t), cudaMemcpyDeviceToHost, S(stream)));
    if (score_out_host)
        CUDA_TRY(cudaMemcpyAsync(score_out_host, e->P.ep_score, G * sizeof(double), cudaMemcpyDeviceToHost, S(stream)));
    if (moves_out_host)
        CUDA_TRY(cudaMemcpyAsync(moves_out_host, e->P.moves_done, G * sizeof(int32_t), cudaMemcpyDeviceToHost, S(stream)));
    CUDA_TRY(cudaStreamSynchronize(S(stream)));
    return bpp_engine_check(e, stream);
}

extern "C" int bpp_engine_play_net_stream_host(bpp_engine* e, bpp_net* net, int choose_mode, uint64_t seed, int num_episodes,
                                               const int32_t* items_wh_host, const int32_t* total_area_host,
                                               const double* bl_host, const int8_t* tie_host, uint32_t* roots_out_host,
                                               int32_t* counts_out_host, int32_t* actions_out_host, int32_t* r_out_host,
                                               double* score_out_host, int32_t* moves_out_host, int32_t* steps_run_host,
                                               void* stream) {
    if (!e || !net || !items_wh_host || !total_area_host || !bl_host) return set_err(BPP_E_INVALID, "null argument");
    if (num_episodes < 1) return set_err(BPP_E_INVALID, "num_episodes = %d", num_episodes);
    const size_t E = (size_t)num_episodes, A = (size_t)e->P.geom.A, N = (size_t)e->P.geom.N;
    int rc;
    if ((rc = ensure(e, &e->q_items, &e->q_items_n, E * N * 2)) || (rc = ensure(e, &e->q_area, &e->q_area_n, E)) ||
        (rc = ensure(e, &e->q_bl, &e->q_bl_n, E)) || (rc = ensure(e, &e->q_tie, &e->q_tie_n, E)) ||
        (rc = ensure(e, &e->q_r, &e->q_r_n, E)) || (rc = ensure(e, &e->q_score, &e->q_score_n, E)) ||
        (rc = ensure(e, &e->q_moves, &e->q_moves_n, E)))
        return rc;
    if (counts_out_host && (rc = ensure(e, &e->q_counts, &e->q_counts_n, N * E * A))) return rc;
    if (actions_out_host && (rc = ensure(e, &e->q_actions, &e->q_actions_n, N * E))) return rc;
    if (roots_out_host && (rc = ensure(e, &e->q_roots, &e->q_roots_n, N * E * REC_WORDS))) return rc;
    CUDA_TRY(cudaMemcpyAsync(e->q_items, items_wh_host, E * N * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, S(stream)));
    CUDA_TRY(cudaMemcpyAsync(e->q_area, total_area_host, E * sizeof(int32_t), cudaMemcpyHostToDevice, S(stream)));
    CUDA_TRY(cudaMemcpyAsync(e->q_bl, bl_host, E * sizeof(double), cudaMemcpyHostToDevice, S(stream)));
    if (tie_host) CUDA_TRY(cudaMemcpyAsync(e->q_tie, tie_host, E * sizeof(int8_t), cudaMemcpyHostToDevice, S(stream)));
    CUDA_TRY(cudaMemsetAsync(e->q_moves, 0, E * sizeof(int32_t), S(stream)));
    if ((rc = bpp_engine_play_net_stream(e, net, choose_mode, seed, num_episodes, e->q_items, e->q_area, e->q_bl,
                                         tie_host ? e->q_tie : nullptr, counts_out_host ? e->q_counts : nullptr,
                                         actions_out_host ? e->q_actions : nullptr, roots_out_host ? e->q_roots : nullptr,
                                         e->q_r, e->q_score, e->q_moves, steps_run_host, stream)))
        return rc;
    if (roots_out_host)
        CUDA_TRY(cudaMemcpyAsync(roots_out_host, e->q_roots, N * E * REC_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                 S(stream)));
    if (counts_out_host)
        CUDA_TRY(cudaMemcpyAsync(counts_out_host, e->q_counts, N * E * A * sizeof(int32_t), cudaMemcpyDeviceToHost, S(stream)));
    if (actions_out_host)
        CUDA_TRY(cudaMemcpyAsync(actions_out_host, e->q_actions, N * E * sizeof(int32_t), cudaMemcpyDeviceToHost, S(stream)));
    if (r_out_host) CUDA_TRY(cudaMemcpyAsync(r_out_host, e->q_r, E * sizeof(int32_t), cudaMemcpyDeviceToHost, S(stream)));
    if (score_out_host)
        CUDA_TRY(cudaMemcpyAsync(score_out_host, e->q_score, E * sizeof(double), cudaMemcpyDeviceToHost, S(stream)));
    if (moves_out_host)
        CUDA_TRY(cudaMemcpyAsync(moves_out_host, e->q_moves, E * sizeof(int32_t), cudaMemcpyDeviceToHost, S(stream)));
    CUDA_TRY(cudaStreamSynchronize(S(stream)));
    return bpp_engine_check(e, stream);
}

extern "C" int bpp_engine_stats(bpp_engine* e, uint64_t stats_host[8], int reset, void* stream) {
    if (!e || !stats_host) return set_err(BPP_E_INVALID, "null argument");
    CUDA_TRY(cudaMemcpyAsync(stats_host, e->P.stats, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost, S(stream)));
    CUDA_TRY(cudaStreamSynchronize(S(stream)));
    stats_host[6] = e->launches;
    if (reset) {
        CUDA_TRY(cudaMemsetAsync(e->P.stats, 0, 8 * sizeof(uint64_t), S(stream)));
        e->launches = 0;
    }
    return BPP_OK;
}

extern "C" int bpp_engine_check(bpp_engine* e, void* stream) {
    if (!e) return set_err(BPP_E_INVALID, "null argument");
    const size_t G = (size_t)e->P.G;
    CUDA_TRY(cudaMemcpyAsync(e->h_status, e->P.status, G * sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
    CUDA_TRY(cudaStreamSynchronize(S(stream)));
    for (size_t g = 0; g < G; ++g) {
        if (e->h_status[g] == -4)
            return set_err(BPP_E_CAPACITY, "game %zu overflowed its node/edge pool (node_cap=%d, edge_cap=%lld units)", g,
                           e->P.node_cap, e->P.edge_cap);
        if (e->h_status[g] < 0)
            return set_err(BPP_E_STATE, "game %zu is in error state %d (illegal action or engine not reset)", g,
                           e->h_status[g]);
    }
    return BPP_OK;
}

extern "C" int bpp_engine_export_game(bpp_engine* e, int game, uint32_t* nodes_out_host, int32_t nodes_cap,
                                      uint64_t* edges_out_host, int64_t units_cap, int32_t* n_nodes_host,
                                      int64_t* n_units_host, void* stream) {
    if (!e || !n_nodes_host || !n_units_host) return set_err(BPP_E_INVALID, "null argument");
    if (game < 0 || game >= e->P.G) return set_err(BPP_E_INVALID, "game index %d out of range", game);
    int nn = 0, nu = 0;
    CUDA_TRY(cudaMemcpyAsync(&nn, e->P.n_nodes + game, sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
    CUDA_TRY(cudaMemcpyAsync(&nu, e->P.n_units + game, sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
    CUDA_TRY(cudaStreamSynchronize(S(stream)));
    *n_nodes_host = nn;
    *n_units_host = nu;
    if (nodes_out_host) {
        if (nodes_cap < nn) return set_err(BPP_E_INVALID, "nodes buffer too small (%d < %d)", nodes_cap, nn);
        CUDA_TRY(cudaMemcpyAsync(nodes_out_host, e->P.nodes + (size_t)game * e->P.node_cap * REC_WORDS,
                                 (size_t)nn * REC_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, S(stream)));
    }
    if (edges_out_host) {
        if (units_cap < nu) return set_err(BPP_E_INVALID, "edges buffer too small");
        CUDA_TRY(cudaMemcpyAsync(edges_out_host, e->P.edges + (size_t)game * (size_t)e->P.edge_cap,
                                 (size_t)nu * sizeof(uint64_t), cudaMemcpyDeviceToHost, S(stream)));
    }
    CUDA_TRY(cudaStreamSynchronize(S(stream)));
    return BPP_OK;
}

extern "C" int bpp_engine_graph_sizes(bpp_engine* e, int32_t* nodes_out_dev, int32_t* units_out_dev, void* stream) {
    if (!e) return set_err(BPP_E_INVALID, "null argument");
    const size_t G = (size_t)e->P.G;
    if (nodes_out_dev)
        CUDA_TRY(cudaMemcpyAsync(nodes_out_dev, e->P.n_nodes, G * sizeof(int), cudaMemcpyDeviceToDevice, S(stream)));
    if (units_out_dev)
        CUDA_TRY(cudaMemcpyAsync(units_out_dev, e->P.n_units, G * sizeof(int), cudaMemcpyDeviceToDevice, S(stream)));
    return BPP_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// env C ABI
static int env_args(int W, int H, int N, int n, const uint32_t* recs, const int32_t* items, EnvArgs* E) {
    int rc = make_geom(W, H, N, &E->geom);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!recs || !items))) return set_err(BPP_E_INVALID, "bad env arguments");
    E->n = n;
    E->recs = recs;
    E->items_wh = items;
    return BPP_OK;
}
#define ENV_LAUNCH_CHECK()                                                                                     \
    do {                                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                                   \
        if (_e != cudaSuccess) return set_err(BPP_E_CUDA, "kernel launch failed: %s", cudaGetErrorString(_e)); \
    } while (0)

extern "C" int bpp_env_valid_moves(int W, int H, int N, int n, const uint32_t* recs_dev, const int32_t* items_wh_dev,
                                   uint8_t* valid_out_dev, void* stream) {
    EnvArgs E;
    int rc = env_args(W, H, N, n, recs_dev, items_wh_dev, &E);
    if (rc) return rc;
    if (n == 0) return BPP_OK;
    if (!valid_out_dev) return set_err(BPP_E_INVALID, "null output");
    k_env_valid<<<grid_warps(n), WARPS_PER_CTA * 32, 0, S(stream)>>>(E, valid_out_dev);
    ENV_LAUNCH_CHECK();
    return BPP_OK;
}

extern "C" int bpp_items_generate(int W, int N, int n, const int64_t* seeds_dev, const int32_t* heights_dev,
                                  int32_t* items_wh_out_dev, int32_t* rects_out_dev, void* stream) {
    if (W < 1 || W > 32 || N < 1 || N > BPP_MAX_ITEMS || n < 0)
        return set_err(BPP_E_INVALID, "bad generator arguments W=%d N=%d n=%d", W, N, n);
    if (n == 0) return BPP_OK;
    if (!seeds_dev || !heights_dev || !items_wh_out_dev) return set_err(BPP_E_INVALID, "null argument");
    k_items_generate<<<(n + 63) / 64, 64, 0, S(stream)>>>(W, N, n, seeds_dev, heights_dev, items_wh_out_dev, rects_out_dev);
    ENV_LAUNCH_CHECK();
    return BPP_OK;
}

extern "C" int bpp_env_planes(int W, int H, int N, int n, const uint32_t* recs_dev, const int32_t* items_wh_dev,
                              float* planes_out_dev, void* stream) {
    EnvArgs E;
    int rc = env_args(W, H, N, n, recs_dev, items_wh_dev, &E);
    if (rc) return rc;
    if (n == 0) return BPP_OK;
    if (!planes_out_dev) return set_err(BPP_E_INVALID, "null output");
    k_env_planes<<<592, 256, 0, S(stream)>>>(E, planes_out_dev);
    ENV_LAUNCH_CHECK();
    return BPP_OK;
}

extern "C" int bpp_env_next_state(int W, int H, int N, int n, const uint32_t* recs_dev, const int32_t* items_wh_dev,
                                  const int32_t* actions_dev, uint32_t* recs_out_dev, void* stream) {
    EnvArgs E;
    int rc = env_args(W, H, N, n, recs_dev, items_wh_dev, &E);
    if (rc) return rc;
    if (n == 0) return BPP_OK;
    if (!actions_dev || !recs_out_dev) return set_err(BPP_E_INVALID, "null argument");
    k_env_next<<<grid_warps(n), WARPS_PER_CTA * 32, 0, S(stream)>>>(E, actions_dev, recs_out_dev);
    ENV_LAUNCH_CHECK();
    return BPP_OK;
}

extern "C" int bpp_env_game_ended(int W, int H, int N, int n, const uint32_t* recs_dev, const int32_t* items_wh_dev,
                                  const int32_t* total_area_dev, const int32_t* max_h_dev, const double* bl_dev,
                                  const int8_t* tie_dev, int32_t* ended_out_dev, double* score_out_dev, void* stream) {
    EnvArgs E;
    int rc = env_args(W, H, N, n, recs_dev, items_wh_dev, &E);
    if (rc) return rc;
    if (!total_area_dev || !max_h_dev || !bl_dev || !ended_out_dev || !score_out_dev)
        return n == 0 ? BPP_OK : set_err(BPP_E_INVALID, "null argument");
    if (n == 0) return BPP_OK;
    k_env_ended<<<grid_warps(n), WARPS_PER_CTA * 32, 0, S(stream)>>>(E, total_area_dev, max_h_dev, bl_dev, tie_dev,
                                                                    ended_out_dev, score_out_dev);
    ENV_LAUNCH_CHECK();
    return BPP_OK;
}
