/*
 * bpp_oracle.c — scalar C restatement of the reference's self-play hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may load this library; nothing under
 * resource_packing_self_play_b200/ does.  It exists because the Python oracle (oracle/bpp_oracle.py, which mirrors
 * the reference's data structures) is too slow to check thousands of games: this one runs ~10^5 simulations/s on one
 * core and lets the -m gpu tests compare whole 4096-game batches bit for bit.
 *
 * Parity status: PINNED — tests/test_oracle_c.py checks it against the fixtures produced by the unmodified
 * reference (tests/golden/) and against oracle/bpp_oracle.py on random instances.
 *
 * It is written independently of the CUDA kernels: the bin is a plain H x W cell array swept with the reference's own
 * loops (no bit tricks), the search graph is a dense per-node table behind a byte-key hash map (the dicts of
 * MCTS_bpp.py:16-26).  Reference lines are cited per function (paths relative to /root/reference/xw_mcts).
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: no FMA contraction, float64 arithmetic must match CPython's).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXW 32
#define MAXH 32
#define MAXN 16
#define MAXA (MAXW * MAXN)

typedef struct {
    int W, H, N, A;
    int iw[MAXN], ih[MAXN];
    int total_area, max_h;
    double bl;
    int has_bl, tie;
} Inst;

typedef struct {
    uint8_t cell[MAXH][MAXW];
    uint32_t rem; /* bit i = item i still to place (its plane is non-zero) */
} State;

/* ---------------------------------------------------------------- BinPackingLogic.py */
/* Bin.get_adjacency, BinPackingLogic.py:47-78 (left test only) */
static int left_adjacent(const Inst *in, const State *s, int x, int w) {
    if (x == 0) return 1;
    int t = 0;
    for (t = 0; t < in->H; ++t) {
        int sum = 0;
        for (int c = x; c < x + w && c < in->W; ++c) sum += s->cell[t][c];
        if (sum == 0) break;
    }
    if (t == in->H) t = in->H - 1; /* Python keeps the last loop value */
    return s->cell[t][x - 1] > 0;
}

/* Bin.get_moves_for_square, BinPackingLogic.py:80-93 */
static int columns_for_item(const Inst *in, const State *s, int item, uint8_t *valid) {
    const int w = in->iw[item], h = in->ih[item];
    int found = 0;
    for (int x = 0; x < in->W - w + 1; ++x) {
        int occ = 0;
        for (int r = 0; r < in->H; ++r)
            for (int c = x; c < x + w; ++c) occ += s->cell[r][c];
        if (occ <= w * in->H - w * h && left_adjacent(in, s, x, w)) {
            if (valid) valid[item * in->W + x] = 1;
            found++;
        }
    }
    return found;
}

/* Bin.execute_move, BinPackingLogic.py:95-109 + getNextState, BinPackingGame.py:58-76 */
static void next_state(const Inst *in, const State *s, int action, State *out) {
    *out = *s;
    const int item = action / in->W, x = action % in->W;
    const int w = in->iw[item], h = in->ih[item];
    int filled = 0;
    for (int r = 0; r < in->H && filled < h; ++r) {
        int sum = 0;
        for (int c = x; c < x + w && c < in->W; ++c) sum += out->cell[r][c];
        if (sum == 0) {
            for (int c = x; c < x + w && c < in->W; ++c) out->cell[r][c] = 1;
            filled++;
        }
    }
    out->rem &= ~(1u << item);
}

/* getValidMoves, BinPackingGame.py:78-92 (returns the count; 0 means has_valid_moves() is False, :94-107) */
static int valid_moves(const Inst *in, const State *s, uint8_t *valid) {
    int n = 0;
    if (valid) memset(valid, 0, (size_t)in->A);
    for (int i = 0; i < in->N; ++i)
        if (s->rem >> i & 1u) n += columns_for_item(in, s, i, valid);
    return n;
}

/* getRankedReward + get_minimal_bin_height, BinPackingGame.py:181-212 */
static int ranked_reward(const Inst *in, const State *s, double *score) {
    int pop = 0;
    for (int r = 0; r < in->H; ++r)
        for (int c = 0; c < in->W; ++c) pop += s->cell[r][c];
    double r;
    if (pop != in->total_area) {
        r = 0.0;
    } else {
        int top = 0;
        for (int i = in->H - 1; i >= 0; --i) {
            int sum = 0;
            for (int c = 0; c < in->W; ++c) sum += s->cell[i][c];
            top = i;
            if (sum > 0) break;
        }
        double lower = ceil((double)in->total_area / (double)in->W);
        if ((double)in->max_h > lower) lower = (double)in->max_h;
        r = lower / (double)(top + 1);
    }
    *score = r;
    if (!in->has_bl) return 1;
    if (r > in->bl || r == 1.0) return 1;
    if (r < in->bl) return -1;
    return in->tie;
}

/* getGameEnded, BinPackingGame.py:109-116 */
static int game_ended(const Inst *in, const State *s, double *score) {
    if (valid_moves(in, s, NULL) > 0) {
        *score = 0.0;
        return 0;
    }
    return ranked_reward(in, s, score);
}

/* ---------------------------------------------------------------- numpy pairwise sum (np.sum, MCTS_bpp.py:90) */
static double pairwise_sum(const double *a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
}

/* ---------------------------------------------------------------- stub evaluators (tests/golden/make_golden.py) */
static void stub_eval(int kind, const Inst *in, const State *s, double *p, double *v) {
    int pop = 0, nrem = 0;
    for (int r = 0; r < in->H; ++r)
        for (int c = 0; c < in->W; ++c) pop += s->cell[r][c];
    for (int i = 0; i < in->N; ++i) nrem += (s->rem >> i) & 1u;
    const double val = (double)((7 * pop + 3 * nrem) % 16) / 16.0 - 0.5;
    for (int a = 0; a < in->A; ++a) {
        if (kind == 1 || kind == 2) p[a] = 1.0 / (double)in->A;
        else if (kind == 3) p[a] = 1.0 / (double)(a + 3 + pop % 5);
        else p[a] = (double)((37 * a + 11 + pop) % 64 + 1) / 4096.0;
    }
    *v = kind == 1 ? 0.0 : val;
}

/* ---------------------------------------------------------------- MCTS, MCTS_bpp.py:11-139 */
typedef struct {
    State key;
    int es_known, es; /* Es[s] */
    int expanded;     /* s in Ps */
    int ns;           /* Ns[s] */
    uint8_t *valid;   /* Vs[s] */
    double *P;        /* Ps[s] */
    int *nsa;         /* Nsa[(s,a)], 0 = (s,a) not in Qsa */
    double *qsa;      /* Qsa[(s,a)] */
} Node;

typedef struct {
    Inst in;
    int num_sims, stub;
    double cpuct;
    Node *nodes;
    int n_nodes, cap_nodes;
    int *table;
    int table_cap;
    long long edges_walked, expansions, sims, terminals;
    /* external evaluator (lockstep cross-checks): if set, called instead of the stub */
    void (*eval)(void *ctx, const uint8_t *cells, uint32_t rem, double *p, double *v);
    void *eval_ctx;
} Mcts;

static uint64_t key_hash(const Inst *in, const State *s) {
    uint64_t h = 1469598103934665603ull;
    for (int r = 0; r < in->H; ++r)
        for (int c = 0; c < in->W; ++c) { h ^= s->cell[r][c]; h *= 1099511628211ull; }
    h ^= s->rem; h *= 1099511628211ull;
    return h ^ (h >> 29);
}
static int key_equal(const Inst *in, const State *a, const State *b) {
    if (a->rem != b->rem) return 0;
    for (int r = 0; r < in->H; ++r)
        if (memcmp(a->cell[r], b->cell[r], (size_t)in->W)) return 0;
    return 1;
}
static void table_grow(Mcts *m) {
    int ncap = m->table_cap ? m->table_cap * 2 : 4096;
    int *t = (int *)malloc(sizeof(int) * (size_t)ncap);
    for (int i = 0; i < ncap; ++i) t[i] = -1;
    for (int k = 0; k < m->n_nodes; ++k) {
        uint64_t h = key_hash(&m->in, &m->nodes[k].key) & (uint64_t)(ncap - 1);
        while (t[h] >= 0) h = (h + 1) & (uint64_t)(ncap - 1);
        t[h] = k;
    }
    free(m->table);
    m->table = t;
    m->table_cap = ncap;
}
static Node *find_or_add(Mcts *m, const State *s) {
    if (m->n_nodes * 2 >= m->table_cap) table_grow(m);
    uint64_t h = key_hash(&m->in, s) & (uint64_t)(m->table_cap - 1);
    while (m->table[h] >= 0) {
        Node *n = &m->nodes[m->table[h]];
        if (key_equal(&m->in, &n->key, s)) return n;
        h = (h + 1) & (uint64_t)(m->table_cap - 1);
    }
    if (m->n_nodes == m->cap_nodes) {
        m->cap_nodes = m->cap_nodes ? m->cap_nodes * 2 : 1024;
        m->nodes = (Node *)realloc(m->nodes, sizeof(Node) * (size_t)m->cap_nodes);
    }
    Node *n = &m->nodes[m->n_nodes];
    memset(n, 0, sizeof(*n));
    n->key = *s;
    m->table[h] = m->n_nodes++;
    return n;
}

static double search(Mcts *m, const State *s) {
    const Inst *in = &m->in;
    Node *n = find_or_add(m, s);
    const int ni = (int)(n - m->nodes);
    if (!n->es_known) { /* :78-79 */
        double sc;
        n->es = game_ended(in, s, &sc);
        n->es_known = 1;
    }
    if (n->es != 0) { /* :81-83 */
        m->terminals++;
        return (double)n->es;
    }
    if (!n->expanded) { /* :85-104 */
        const int A = in->A;
        n->valid = (uint8_t *)malloc((size_t)A);
        n->P = (double *)malloc(sizeof(double) * (size_t)A);
        n->nsa = (int *)calloc((size_t)A, sizeof(int));
        n->qsa = (double *)calloc((size_t)A, sizeof(double));
        double v;
        if (m->eval) m->eval(m->eval_ctx, &s->cell[0][0], s->rem, n->P, &v);
        else stub_eval(m->stub, in, s, n->P, &v);
        valid_moves(in, s, n->valid);
        for (int a = 0; a < A; ++a) n->P[a] = n->P[a] * (double)n->valid[a];
        double tot = pairwise_sum(n->P, A);
        if (tot > 0) {
            for (int a = 0; a < A; ++a) n->P[a] /= tot;
        } else {
            for (int a = 0; a < A; ++a) n->P[a] = n->P[a] + (double)n->valid[a];
            tot = pairwise_sum(n->P, A);
            for (int a = 0; a < A; ++a) n->P[a] /= tot;
        }
        n->ns = 0;
        n->expanded = 1;
        m->expansions++;
        return v;
    }
    double best = -INFINITY;
    int best_a = -1;
    for (int a = 0; a < in->A; ++a) { /* :111-121 */
        if (!n->valid[a]) continue;
        double u;
        if (n->nsa[a] > 0) u = n->qsa[a] + ((m->cpuct * n->P[a]) * sqrt((double)n->ns)) / (double)(1 + n->nsa[a]);
        else u = (m->cpuct * n->P[a]) * sqrt((double)n->ns + 1e-8);
        if (u > best) { best = u; best_a = a; }
    }
    State child;
    next_state(in, s, best_a, &child);
    m->edges_walked++;
    const double v = search(m, &child);
    n = &m->nodes[ni]; /* the node array may have moved */
    if (n->nsa[best_a] > 0) { /* :130-136 */
        n->qsa[best_a] = ((double)n->nsa[best_a] * n->qsa[best_a] + v) / (double)(n->nsa[best_a] + 1);
        n->nsa[best_a] += 1;
    } else {
        n->qsa[best_a] = v;
        n->nsa[best_a] = 1;
    }
    n->ns += 1;
    return v;
}

/* ---------------------------------------------------------------- exported API (ctypes) */
static void inst_init(Inst *in, int W, int H, int N, const int32_t *items_wh, int total_area, double bl, int tie) {
    memset(in, 0, sizeof(*in));
    in->W = W; in->H = H; in->N = N; in->A = W * N;
    in->max_h = 0;
    for (int i = 0; i < N; ++i) {
        in->iw[i] = items_wh[2 * i];
        in->ih[i] = items_wh[2 * i + 1];
        if (in->ih[i] > in->max_h) in->max_h = in->ih[i];
    }
    in->total_area = total_area;
    in->has_bl = !(bl != bl);
    in->bl = bl;
    in->tie = tie;
}
static void state_from_rec(const Inst *in, const uint32_t *rec, State *s) {
    memset(s, 0, sizeof(*s));
    for (int r = 0; r < in->H; ++r)
        for (int c = 0; c < in->W; ++c) s->cell[r][c] = (uint8_t)((rec[r] >> c) & 1u);
    s->rem = rec[28];
}
static void state_to_rec(const Inst *in, const State *s, uint32_t *rec) {
    memset(rec, 0, 32 * sizeof(uint32_t));
    for (int r = 0; r < in->H; ++r)
        for (int c = 0; c < in->W; ++c) rec[r] |= (uint32_t)s->cell[r][c] << c;
    rec[28] = s->rem;
}

/* batched env ops on compact records (same layout as include/bpp_b200.h) */
void oracle_valid_moves(int W, int H, int N, int n, const uint32_t *recs, const int32_t *items_wh, uint8_t *valid_out) {
    for (int i = 0; i < n; ++i) {
        Inst in; State s;
        inst_init(&in, W, H, N, items_wh + (size_t)i * N * 2, 0, NAN, 1);
        state_from_rec(&in, recs + (size_t)i * 32, &s);
        valid_moves(&in, &s, valid_out + (size_t)i * in.A);
    }
}
void oracle_next_state(int W, int H, int N, int n, const uint32_t *recs, const int32_t *items_wh, const int32_t *actions,
                       uint32_t *recs_out) {
    for (int i = 0; i < n; ++i) {
        Inst in; State s, o;
        inst_init(&in, W, H, N, items_wh + (size_t)i * N * 2, 0, NAN, 1);
        state_from_rec(&in, recs + (size_t)i * 32, &s);
        next_state(&in, &s, actions[i], &o);
        state_to_rec(&in, &o, recs_out + (size_t)i * 32);
    }
}
void oracle_game_ended(int W, int H, int N, int n, const uint32_t *recs, const int32_t *items_wh,
                       const int32_t *total_area, const int32_t *max_h, const double *bl, const int8_t *tie,
                       int32_t *ended, double *score) {
    for (int i = 0; i < n; ++i) {
        Inst in; State s;
        inst_init(&in, W, H, N, items_wh + (size_t)i * N * 2, total_area[i], bl[i], tie ? tie[i] : 1);
        in.max_h = max_h[i];
        state_from_rec(&in, recs + (size_t)i * 32, &s);
        ended[i] = game_ended(&in, &s, &score[i]);
    }
}
double oracle_pairwise_sum(const double *a, int n) { return pairwise_sum(a, n); }

static void mcts_free(Mcts *m) {
    for (int k = 0; k < m->n_nodes; ++k) {
        free(m->nodes[k].valid); free(m->nodes[k].P); free(m->nodes[k].nsa); free(m->nodes[k].qsa);
    }
    free(m->nodes); free(m->table);
}

/* One self-play episode driven like tests/golden/make_golden.py (CoachBPP.py:50-99 with a deterministic action
 * choice): policy 0 = first arg-max of the visit counts, 1 = last action with a non-zero count, 2 = take
 * forced_actions[move].  counts_out: int32 [N][A] (rows beyond the last move untouched), actions_out: int32 [N].
 * stats_out (may be NULL): [sims, edges walked, expansions, terminal hits, nodes, edges(Nsa entries)].
 * Returns the number of moves played. */
int oracle_play_episode(int W, int H, int N, const int32_t *items_wh, int total_area, double bl, int tie, int stub,
                        int num_sims, double cpuct, int policy, const int32_t *forced_actions, int32_t *counts_out,
                        int32_t *actions_out, int32_t *r_out, double *score_out, long long *stats_out) {
    Mcts m;
    memset(&m, 0, sizeof(m));
    inst_init(&m.in, W, H, N, items_wh, total_area, bl, tie);
    m.num_sims = num_sims; m.stub = stub; m.cpuct = cpuct;
    State root;
    memset(&root, 0, sizeof(root));
    root.rem = N >= 32 ? 0xffffffffu : ((1u << N) - 1u);
    const int A = m.in.A;
    int moves = 0, r = 0;
    double score = 0.0;
    while (moves < N) {
        for (int i = 0; i < num_sims; ++i) { m.sims++; search(&m, &root); } /* getActionProb, :37-38 */
        Node *n = find_or_add(&m, &root);
        int best = -1, best_a = -1, last = -1;
        for (int a = 0; a < A; ++a) {
            const int c = (n->expanded) ? n->nsa[a] : 0;
            if (counts_out) counts_out[(size_t)moves * A + a] = c;
            if (c > best) { best = c; best_a = a; }
            if (c > 0) last = a;
        }
        int a = policy == 0 ? best_a : (policy == 1 ? last : forced_actions[moves]);
        if (actions_out) actions_out[moves] = a;
        moves++;
        State nxt;
        next_state(&m.in, &root, a, &nxt);
        root = nxt;
        r = game_ended(&m.in, &root, &score);
        if (r != 0) break;
    }
    if (r_out) *r_out = r;
    if (score_out) *score_out = score;
    if (stats_out) {
        long long edges = 0;
        for (int k = 0; k < m.n_nodes; ++k)
            if (m.nodes[k].expanded)
                for (int a = 0; a < A; ++a) edges += m.nodes[k].nsa[a] > 0;
        stats_out[0] = m.sims; stats_out[1] = m.edges_walked; stats_out[2] = m.expansions;
        stats_out[3] = m.terminals; stats_out[4] = m.n_nodes; stats_out[5] = edges;
    }
    mcts_free(&m);
    return moves;
}
