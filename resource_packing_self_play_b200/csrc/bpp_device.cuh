// bpp_device.cuh — warp-cooperative device primitives of the bin-packing self-play hot path (sm_100a).
//
// One warp owns one game / one state.  The occupancy grid lives as one 32-bit row mask per lane (lane r = bin row r),
// mirrored into a 32-word shared-memory strip per warp for the broadcast sweeps of the valid-move computation.
// Every function here is warp-collective: all 32 lanes must call it convergently.
//
// Reference semantics implemented (paths relative to /root/reference/xw_mcts):
//   valid_words      BinPackingGame.getValidMoves :78-92, Bin.get_moves_for_square / get_adjacency
//                    (binpacking/BinPackingLogic.py:47-93)
//   apply_move       BinPackingGame.getNextState :58-76, Bin.execute_move (BinPackingLogic.py:95-109)
//   terminal_value   BinPackingGame.getRankedReward :188-212, get_minimal_bin_height :181-186
//   np_pairwise_sum  numpy's pairwise float64 add-reduction (np.sum at MCTS_bpp.py:90,100)
//   puct_select      MCTS.search PUCT loop (MCTS_bpp.py:107-121), first-max tie-break
//   backup_path      MCTS.search backup (MCTS_bpp.py:130-139)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bpp {

constexpr unsigned FULL = 0xffffffffu;
constexpr int REC_WORDS = 32;
constexpr int REC_REM = 28;   // remaining-items mask
constexpr int REC_NS = 29;    // Ns[s]
constexpr int REC_OFF = 30;   // edge block offset (8-byte units) in the game's edge pool
constexpr int REC_META = 31;  // nvalid | kind << 16
constexpr int KIND_NEW = 0;   // key known, never visited by search (Es/Ps not computed yet)
constexpr int KIND_EXP = 1;   // expanded (Ps, Vs, Ns exist)
constexpr int KIND_TPOS = 2;  // terminal, Es = +1
constexpr int KIND_TNEG = 3;  // terminal, Es = -1
constexpr int MAX_AW = 16;    // valid-mask words (A <= 512)
constexpr int MAX_LEAVES = 16;

// numpy pairwise-sum plan for a length-A float64 reduction (host-built, see build_sum_plan_rec in bpp_engine.cu):
// leaf blocks of <= 128 elements, then the recursion's additions as in-place pair operations s[dst] += s[src]
// executed in order; the total ends in s[0].
struct SumPlan {
    int n_leaves;
    int n_ops;
    short leaf_base[MAX_LEAVES];
    short leaf_n[MAX_LEAVES];
    signed char op_dst[MAX_LEAVES];
    signed char op_src[MAX_LEAVES];
};

struct Geom {
    int W, H, N, A, AW;
    unsigned wmask;  // low W bits
    unsigned invW;   // ceil(2^16 / W): a / W == (a * invW) >> 16 for 0 <= a < 512 (W <= 32)
};
__device__ __forceinline__ int div_w(const Geom& g, int a) { return (int)(((unsigned)a * g.invW) >> 16); }

__device__ __forceinline__ bool state_lane(int lane, int H) { return lane < H || lane == REC_REM; }

// ---------------------------------------------------------------------------------------------------------------------
// key hash: murmur-style per-lane mix, XOR-reduced with REDUX
__device__ __forceinline__ uint32_t hash_state(uint32_t rec, int lane, int H) {
    uint32_t v = state_lane(lane, H) ? rec : 0u;
    uint32_t h = v + 0x9E3779B9u * (uint32_t)(lane + 1);
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    h = __reduce_xor_sync(FULL, h);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}

__device__ __forceinline__ uint32_t strip_mask(int w, int x, unsigned wmask) {
    uint32_t m = (w >= 32) ? 0xffffffffu : ((1u << w) - 1u);
    return (m << x) & wmask;
}

// ---------------------------------------------------------------------------------------------------------------------
// Valid-move mask.  s_occ[0..H) = occupancy rows (shared, this warp's 16-byte aligned strip), s_items[i] = w | h << 8,
// s_tab = 16-entry scratch, s_col = VW_SCRATCH words of scratch.  The (remaining item, column x) pairs are enumerated
// densely: pair p = k*W + x, k = rank of the item among the remaining ones, so late in an episode (few items left) only
// ceil(nrem*W/32) rounds run.
//   (A) cell-count test (BinPackingLogic.py:89): occupied cells of the strip [:, x:x+w] <= w*(H-h);
//   (B) left adjacency (BinPackingLogic.py:63-70): x == 0, or the cell left of the strip is occupied in the first strip
//       row that is completely empty (row H-1 if none: the reference's loop variable keeps its last value).
// COLUMN view (round 2; the row sweep per pair - H x {and, popc, add, compare, or} - was 40 % of k_episode's
// instructions): lane c builds col[c] = the H-bit occupancy of column c once per call; then per pair
//   occupied rows of the strip = OR of col[x .. x+w)  -> two look-ups in a sparse table of range ORs (level k covers 2^k
//                                                       columns, built with one shuffle per level),
//   cell count                 = pre[x+w] - pre[x]      (prefix sums of popc(col), one warp scan),
//   left neighbour             = bit t of col[x-1].
// HC > 0 = compile-time bin height (the column build is unrolled over vector loads of the rows); HC == 0 = generic.
// Returns word k of the A-bit mask in lane k (0 elsewhere); also leaves the words in s_vw[0..MAX_AW).
constexpr int VW_LEVELS = 6;                       // range ORs over 1, 2, 4, 8, 16, 32 columns
constexpr int VW_PRE = VW_LEVELS * 32;             // prefix sums pre[0..32] behind the table
constexpr int VW_SCRATCH = VW_PRE + 36;            // 32-bit words
template <int HC>
__device__ __forceinline__ uint32_t valid_words(const Geom& g, const uint32_t* s_occ, const uint16_t* s_items,
                                                uint32_t rem, int lane, uint32_t* s_vw, uint8_t* s_tab, uint32_t* s_col) {
    const int H = HC ? HC : g.H;
    uint32_t col = 0;
    if (HC) {
        constexpr int NR = (HC + 3) & ~3;
#pragma unroll
        for (int r4 = 0; r4 < NR; r4 += 4) {
            const uint4 v = *reinterpret_cast<const uint4*>(s_occ + r4);
            col |= ((v.x >> lane) & 1u) << r4;
            if (r4 + 1 < HC) col |= ((v.y >> lane) & 1u) << (r4 + 1);
            if (r4 + 2 < HC) col |= ((v.z >> lane) & 1u) << (r4 + 2);
            if (r4 + 3 < HC) col |= ((v.w >> lane) & 1u) << (r4 + 3);
        }
    } else {
        for (int r = 0; r < H; ++r) col |= ((s_occ[r] >> lane) & 1u) << r;
    }
    if (lane >= g.W) col = 0u;
    rem &= (g.N >= 32) ? 0xffffffffu : ((1u << g.N) - 1u);
    if (lane < MAX_AW) s_vw[lane] = 0u;
    if ((rem >> lane) & 1u) s_tab[__popc(rem & ((1u << lane) - 1u))] = (uint8_t)lane;  // rank -> item index
    {   // sparse table of range ORs: level k, entry c = OR of col[c .. c + 2^k) (clipped at the last lane)
        uint32_t t = col;
        s_col[lane] = t;
        t |= __shfl_down_sync(FULL, t, 1);
        s_col[32 + lane] = t;
        t |= __shfl_down_sync(FULL, t, 2);
        s_col[64 + lane] = t;
        t |= __shfl_down_sync(FULL, t, 4);
        s_col[96 + lane] = t;
        if (g.W > 15) {   // items at least 16 (32) wide exist only in bins that wide
            t |= __shfl_down_sync(FULL, t, 8);
            s_col[128 + lane] = t;
            if (g.W > 31) {
                t |= __shfl_down_sync(FULL, t, 16);
                s_col[160 + lane] = t;
            }
        }
        int inc = __popc(col);   // inclusive scan of the column counts
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += o;
        }
        s_col[VW_PRE + 1 + lane] = (uint32_t)inc;
        if (lane == 0) s_col[VW_PRE] = 0u;
    }
    __syncwarp();
    const uint32_t hmask = (1u << H) - 1u;   // H <= 28
    const int npairs = __popc(rem) * g.W;
    for (int p0 = 0; p0 < npairs; p0 += 32) {
        const int p = p0 + lane;
        if (p < npairs) {
            const int k = div_w(g, p);
            const int x = p - k * g.W;
            const int i = s_tab[k];
            const int w = s_items[i] & 0xff, h = s_items[i] >> 8;
            if (x + w <= g.W) {
                const int lv = 31 - __clz(w | 1);
                const uint32_t* tk = s_col + lv * 32;
                const uint32_t occ_rows = w > 0 ? (tk[x] | tk[x + w - (1 << lv)]) : 0u;  // rows with a cell in the strip
                const uint32_t emp = ~occ_rows & hmask;                                  // strip rows completely empty
                const int cnt = (int)s_col[VW_PRE + x + w] - (int)s_col[VW_PRE + x];
                const int t = emp ? __ffs(emp) - 1 : H - 1;
                if ((cnt <= w * (H - h)) && (x == 0 || ((s_col[x - 1] >> t) & 1u))) {
                    const int a = i * g.W + x;
                    atomicOr(&s_vw[a >> 5], 1u << (a & 31));
                }
            }
        }
    }
    __syncwarp();
    return lane < MAX_AW ? s_vw[lane] : 0u;
}

// ---------------------------------------------------------------------------------------------------------------------
// Placement: fill the strip in the first h strip rows that are completely empty (rows need not be contiguous; fewer
// than h -> silent truncation).  rec is the lane's word of the state record; returns the new word.
__device__ __forceinline__ uint32_t apply_move(const Geom& g, uint32_t rec, int lane, int item, int w, int h, int x) {
    const uint32_t m = strip_mask(w, x, g.wmask);
    const bool e = (lane < g.H) && ((rec & m) == 0);
    const uint32_t E = __ballot_sync(FULL, e);
    const int before = __popc(E & ((1u << lane) - 1u));
    if (e && before < h) rec |= m;
    if (lane == REC_REM) rec &= ~(1u << item);
    return rec;
}

// ---------------------------------------------------------------------------------------------------------------------
struct RewardCtx {
    int total_area;  // items_total_area
    int numer;       // max(ceil(total_area / W), max_h)
    double bl;       // ranked-reward threshold, NaN = empty rewards list
    int tie;         // value on r == bl
};

// returns +1 / -1 and the raw reward r (score)
__device__ __forceinline__ int terminal_value(const Geom& g, const RewardCtx& rc, uint32_t rec, int lane, double* score) {
    const uint32_t rowv = lane < g.H ? rec : 0u;
    const int pop = __reduce_add_sync(FULL, (unsigned)__popc(rowv));
    const uint32_t ne = __ballot_sync(FULL, rowv != 0u);
    double r;
    if (pop != rc.total_area) {
        r = 0.0;  // some item discarded or truncated, BinPackingGame.py:193-195
    } else {
        const int top = ne ? 32 - __clz(ne) : 1;  // get_minimal_bin_height, :181-186
        r = __ddiv_rn((double)rc.numer, (double)top);
    }
    *score = r;
    if (rc.bl != rc.bl) return 1;  // empty rewards list, :203-204
    if (r > rc.bl || r == 1.0) return 1;
    if (r < rc.bl) return -1;
    return rc.tie;
}

// ---------------------------------------------------------------------------------------------------------------------
// numpy float64 pairwise add-reduction (np.sum) of the dense length-A vector whose entry a is f(a) where bit a of the
// valid mask vw is set and 0.0 elsewhere, bit-exact: blocks of <= 128 elements use 8 interleaved accumulators
// (accumulator j takes positions j, j+8, ...) combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then a sequential tail;
// longer inputs split recursively at n/2 rounded down to a multiple of 8 (the SumPlan).  Adding +0.0 never changes a
// partial sum, so only the VALID positions are visited (in the same ascending order).  Each 8-lane group of the warp
// evaluates one block.  scratch: >= MAX_LEAVES doubles of this warp's shared memory.  Result is warp-uniform.
template <typename F>
__device__ __forceinline__ double np_masked_sum(const SumPlan& plan, const uint32_t* vw, F f, int lane,
                                                double* scratch) {
    const int j = lane & 7;
#pragma unroll 1
    for (int pass = 0; pass * 4 < plan.n_leaves; ++pass) {
        const int li = pass * 4 + (lane >> 3);
        const bool act = li < plan.n_leaves;
        const int base = act ? plan.leaf_base[li] : 0;  // multiple of 8
        const int n = act ? plan.leaf_n[li] : 0;
        const int lim = n >= 8 ? n - (n & 7) : 0;
        double r = 0.0;
        if (lim > 0) {
            const int pend = base + lim;
            const uint32_t sel = 0x01010101u << j;  // positions == j (mod 8)
#pragma unroll 1
            for (int k = base >> 5; k <= (pend - 1) >> 5; ++k) {
                const int lo = k << 5;
                uint32_t m = vw[k] & sel;
                if (lo < base) m &= 0xffffffffu << (base - lo);
                if (pend - lo < 32) m &= (1u << (pend - lo)) - 1u;
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1u;
                    r = __dadd_rn(r, f(lo + bit));
                }
            }
        }
        double t = __dadd_rn(r, __shfl_xor_sync(FULL, r, 1));
        t = __dadd_rn(t, __shfl_xor_sync(FULL, t, 2));
        t = __dadd_rn(t, __shfl_xor_sync(FULL, t, 4));
        double res = lim > 0 ? t : 0.0;
#pragma unroll 1
        for (int a = base + lim; a < base + n; ++a)  // sequential tail (< 8 positions)
            if ((vw[a >> 5] >> (a & 31)) & 1u) res = __dadd_rn(res, f(a));
        if (act && j == 0) scratch[li] = res;
    }
    __syncwarp();
    if (plan.n_ops > 0) {
        if (lane == 0)
#pragma unroll 1
            for (int t = 0; t < plan.n_ops; ++t)
                scratch[plan.op_dst[t]] = __dadd_rn(scratch[plan.op_dst[t]], scratch[plan.op_src[t]]);
        __syncwarp();
    }
    const double total = scratch[0];
    __syncwarp();
    return total;
}

// ---------------------------------------------------------------------------------------------------------------------
// Edge block of an expanded node with nv valid actions, nvp = nv rounded up to 4; 8-byte units:
//   [0, nvp)        Q   float64     Qsa
//   [nvp, 2nvp)     P   float64     Ps[s][a] (masked + renormalised prior)
//   [2nvp, 3nvp)    NC  {int32 Nsa, int32 child node (-1 = not resolved yet)}
//   [3nvp, 3nvp + nvp/4)  ACT uint16 action index, ascending
struct EdgeBlock {
    double* Q;
    double* P;
    int2* NC;
    uint16_t* ACT;
    __device__ __forceinline__ EdgeBlock(unsigned long long* base, int nvp)
        : Q(reinterpret_cast<double*>(base)),
          P(reinterpret_cast<double*>(base) + nvp),
          NC(reinterpret_cast<int2*>(base + 2 * (size_t)nvp)),
          ACT(reinterpret_cast<uint16_t*>(base + 3 * (size_t)nvp)) {}
};
__host__ __device__ __forceinline__ int edge_units(int nv) {
    const int nvp = (nv + 3) & ~3;
    return 3 * nvp + (nvp >> 2);
}

// PUCT arg-max with the reference's exact operation order and first-max tie-break (ascending action == ascending
// edge index).  u = Q + ((cpuct*P)*sqrt(Ns)) / (1+Nsa) for visited edges, (cpuct*P)*sqrt(Ns+1e-8) otherwise.
// Correctly rounded a / d for a small positive integer d without the ~35-instruction division sequence: with
// y = RN(1/d) from a host-computed table, q0 = RN(a*y) is within one ulp, the residual r = a - d*q0 is exact in one
// FMA, and q1 = RN(q0 + r*y) is the correctly rounded quotient (Markstein's theorem; d's significand is never all
// ones here).  Verified against IEEE division on 3e5 random operands and by the bit-exact Q / visit-count
// parity tests.  tab3 = [sqrt(n) | sqrt(n + 1e-8) | 1/n], each tab_n entries.
__device__ __forceinline__ double div_small(double a, int d, const double* __restrict__ tab3, int tab_n) {
    if (d < tab_n) {
        const double y = __ldg(tab3 + 2 * tab_n + d);
        const double q0 = __dmul_rn(a, y);
        const double r = __fma_rn(-(double)d, q0, a);
        return __fma_rn(r, y, q0);
    }
    return __ddiv_rn(a, (double)d);
}

// sqrt_tab: optional host-computed tables sqrt(n) [0..tab_n) and sqrt(n + 1e-8) [tab_n..2*tab_n) — the same correctly
// rounded IEEE values __dsqrt_rn returns, as one broadcast load instead of a ~50-instruction software sequence.
__device__ __forceinline__ int puct_select(const EdgeBlock& eb, int nv, int Ns, double cpuct, int lane,
                                           const double* __restrict__ sqrt_tab, int tab_n) {
    if (nv == 1) return 0;  // a single legal action: the arg-max needs no arithmetic
    double bu = __longlong_as_double((long long)0xfff0000000000000ull);  // -inf
    int be = 0x7fffffff;
    // first round from registers: issue all loads, then decide which square root is needed (the root has every edge
    // visited, a fresh node none), then the arithmetic
    const bool in0 = lane < nv;
    const double q0 = in0 ? eb.Q[lane] : 0.0;
    const double p0 = in0 ? eb.P[lane] : 0.0;
    const int n0 = in0 ? eb.NC[lane].x : 0;
    bool need_sq = in0 && n0 > 0, need_sqe = in0 && n0 <= 0;
#pragma unroll 1
    for (int e = lane + 32; e < nv; e += 32) {
        const int n = eb.NC[e].x;
        need_sq |= n > 0;
        need_sqe |= n <= 0;
    }
    double sq, sqe;
    if (Ns < tab_n) {
        sq = __ldg(sqrt_tab + Ns);
        sqe = __ldg(sqrt_tab + tab_n + Ns);
    } else {
        sq = __any_sync(FULL, need_sq) ? __dsqrt_rn((double)Ns) : 0.0;
        sqe = __any_sync(FULL, need_sqe) ? __dsqrt_rn(__dadd_rn((double)Ns, 1e-8)) : 0.0;
    }
    if (in0) {
        const double cp = __dmul_rn(cpuct, p0);
        double u = n0 > 0 ? __dadd_rn(q0, div_small(__dmul_rn(cp, sq), 1 + n0, sqrt_tab, tab_n)) : __dmul_rn(cp, sqe);
        bu = __dadd_rn(u, 0.0);  // canonicalise -0.0 (Python's `>` treats it as equal to +0.0)
        be = lane;
    }
#pragma unroll 1
    for (int e = lane + 32; e < nv; e += 32) {
        const double q = eb.Q[e];
        const double p = eb.P[e];
        const int n = eb.NC[e].x;
        const double cp = __dmul_rn(cpuct, p);
        double u = n > 0 ? __dadd_rn(q, div_small(__dmul_rn(cp, sq), 1 + n, sqrt_tab, tab_n)) : __dmul_rn(cp, sqe);
        u = __dadd_rn(u, 0.0);
        if (u > bu) { bu = u; be = e; }
    }
    // order-preserving map double -> uint64, then two 32-bit REDUX max + one REDUX min on the edge index
    unsigned long long bits = (unsigned long long)__double_as_longlong(bu);
    bits ^= (bits >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull;
    const uint32_t hi = (uint32_t)(bits >> 32), lo = (uint32_t)bits;
    const uint32_t mh = __reduce_max_sync(FULL, hi);
    const uint32_t ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
    const bool match = (hi == mh) && (lo == ml) && (be != 0x7fffffff);
    return (int)__reduce_min_sync(FULL, match ? (unsigned)be : 0x7fffffffu);
}

// One path entry per lane (lane d = depth d).  Paths never revisit a node (every move removes an item), so the
// lanes update disjoint memory.
struct PathEntry {
    int node;
    int off;   // edge block offset (units)
    int nvp;
    int e;
};

__device__ __forceinline__ void backup_path(uint32_t* nodes, unsigned long long* edges, const PathEntry& pe, int depth,
                                            double v, int lane, const double* __restrict__ tab3, int tab_n) {
    if (lane < depth) {
        EdgeBlock eb(edges + pe.off, pe.nvp);
        const int n = eb.NC[pe.e].x;
        const double q = eb.Q[pe.e];
        // MCTS_bpp.py:130-136: Q <- (N*Q + v)/(N+1), first visit Q <- v
        const double qn = n > 0 ? div_small(__dadd_rn(__dmul_rn((double)n, q), v), n + 1, tab3, tab_n) : v;
        eb.Q[pe.e] = qn;
        eb.NC[pe.e].x = n + 1;
        nodes[(size_t)pe.node * REC_WORDS + REC_NS] += 1u;  // Ns[s] += 1, :138
    }
    __syncwarp();
}

// stub evaluators (tests/golden/make_golden.py)
__device__ __forceinline__ double stub_value_of(int pop, int nrem) {
    return (double)((7 * pop + 3 * nrem) & 15) / 16.0 - 0.5;
}
template <int STUB>
__device__ __forceinline__ double stub_prior(int a, int A, int pop) {
    if (STUB == 1 || STUB == 2) return __ddiv_rn(1.0, (double)A);
    if (STUB == 3) return __ddiv_rn(1.0, (double)(a + 3 + pop % 5));
    return __ddiv_rn((double)((37 * a + 11 + pop) % 64 + 1), 4096.0);
}

}  // namespace bpp
