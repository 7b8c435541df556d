// bpp_net_gr.cuh — "grid-row" tcgen05 trunk of the policy/value network (sm_100a).  Included by bpp_net.cu inside its
// anonymous namespace, after NetParams.
//
// Why a second formulation.  With both operands in shared memory a tcgen05.mma of M = 128, K = 16 costs the fetch of its
// 4 KB A operand (~40 cycles) however few output channels it feeds (profiles/r01_umma_probe.txt): with N = Cout = 16 / 32
// the tensor pipe does math 20 % / 40 % of the time, and a 3x3 convolution as nine shifted N = Cout MMAs per 16 input
// channels is bound by 89 x 40 cycles per leaf.  The only way to feed more columns per A fetch in this network is to
// stack kernel taps along N - and a stacked tap has to be un-shifted again when the accumulators are read.  A shift along
// the pixel ROW axis is a shift across TMEM lanes (impossible without going through shared memory); this layout makes the
// vertical shift a shift across accumulator COLUMN blocks instead, which costs nothing:
//
//   * row axis of a level = [grid row y][leaf j][padded column xp], xp = 0 being the zero halo column shared by
//     neighbouring leaves; one 128-row MMA tile = grid row y of the J leaves of a group (J*(w+1) <= 128: 8 leaves at 15x15,
//     14 at 8x8, 25 at 4x4, 42 at 2x2 - hence one STAGE per level, each with its own group size and shared-memory plan);
//   * for INPUT tile t the three vertical taps are stacked along N: B = [W(dy=0,dx) | W(dy=1,dx) | W(dy=2,dx)], N = 3*Cout,
//     accumulated over dx (A shifted by dx-1 rows, i.e. only the descriptor's start address moves) and the input-channel
//     chunks; block dy belongs to OUTPUT tile t+1-dy.  Output tile Y owns slot Y mod R of a ring of R accumulator slots of
//     Cout TMEM columns (256 columns per group), laid out so that the three blocks of one MMA land in three consecutive
//     column blocks = three consecutive output tiles, on the SAME lanes (two MMAs where the ring wraps);
//   * 3*cin16 MMAs of N = 48 / 96 per tile instead of 9*cin16 of N = 16 / 32: ~2.3x fewer tensor cycles.  A stacked MMA
//     cannot overwrite one block and accumulate into the other two, so the epilogue warps pre-load every slot they drain
//     with the bias of its next user (tcgen05.st) and every MMA accumulates;
//   * an input tile is read by ITS OWN MMAs only, so every layer runs IN PLACE: the epilogue of output tile Y (which waits
//     for the MMAs of input tile Y+1) overwrites input tile Y.  A residual block needs two buffers (raw stream, relu'd
//     activations) instead of three plus a pooling buffer; halo rows are never written and stay zero for the whole stage;
//   * all weights of a level stay in shared memory (one bulk async copy per layer at the start of the stage; split mode:
//     streamed through two slots where they do not fit), and the layers of a group are chained by per-tile mbarriers
//     instead of CTA-wide barriers: the issuer starts layer l+1 on tile t as soon as the epilogue warps have stored tiles
//     <= t+1 of layer l, so the tensor pipe does not drain between layers;
//   * a CTA runs up to two independent groups ("subs": 8 epilogue warps + 1 issuer warp each) that share the weights and
//     split the 512 TMEM columns - one sub's epilogue / input / pooling phases overlap the other's MMAs;
//   * the four stages run in ONE launch (k_net_gr): every CTA takes the same contiguous share of the batch through all
//     levels, so a stage only reads hand-over data written by its own CTA and no grid-wide barrier is needed.
//
// Stages: 0 = input planes from the compact records -> conv -> max-pool -> x1;  1 = x1 -> two residual blocks -> conv ->
// max-pool -> x2;  2 = the same one level down -> x3;  3 = x3 -> two residual blocks -> relu(flatten) -> feat (the FC heads
// run in k_net_heads_tc).  Hand-over layout: [leaf][hi | lo (split mode)][plane of 8 channels][pixel] 16-byte units,
// interior pixels only (BinpackingNNet.py:29-48,72-81).
#pragma once

#ifndef BPP_GR_TB23
#define BPP_GR_TB23 1   // tiles per epilogue batch in stages 2, 3 (16 columns per warp)
#endif

namespace bppgr {
using namespace bpptc;

constexpr int SUB_THREADS = 288;   // 8 epilogue warps + 1 issuer warp
constexpr int MAX_TILES = 28;      // grid rows of a level (H <= 28)
constexpr int G0 = 8;              // zero guard rows in front of tile 0 (the dx = 0 window starts one row early)
constexpr int MAX_LAY = 5;

struct GrStage {
    int h, w, wp, J, TS, NT, RT;   // level geometry, leaves per group, tile stride (rows), tiles (= h), rows per plane
    int arena_planes;              // 8-channel planes per sub arena
    int cp;                        // planes of the residual stream (= input channels / 8)
    int h2, w2;                    // pooled geometry (stages 0..2)
    int nsub, tmem_cols, col_sub;  // groups in flight per CTA, TMEM columns allocated, column offset of sub 1
    int nlay;                      // conv layers of the stage; the LAST one of stages 0..2 is the sequence's first conv
    int cin16[MAX_LAY], cout[MAX_LAY];
    int w_soff[MAX_LAY];           // byte offset of each layer's weights in the shared weight block
    int w_len[MAX_LAY];            // bytes
    long long w_goff[MAX_LAY];     // element offset in the grid-row weight buffer
    int b_goff[MAX_LAY];           // offset of the layer's bias in NetParams::bias
    int w_bytes;                   // shared weight block
    int arena_off, arena_bytes, smem_bytes;
    uint32_t m_w, m_w2, m_hw2, m_php2, m_flat, m_pw2;   // fdiv magics: w, w2, h2*w2, planes_out*h2*w2, flat, planes_out*w2
    int planes_out;                // planes of the pooled hand-over (stages 0..2)
    // split-bf16 mode (hi + lo halves, three MMAs per product): the lo planes of the arena sit lo_off bytes behind the hi
    // planes, a layer's lo weights w_len bytes behind its hi weights; `stream`: the weights do not fit next to the arena and
    // pass through two slots of slot_bytes, one layer ahead
    int lo_off, stream, slot_bytes;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// one bulk asynchronous copy global -> shared (TMA engine, no tensor map), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void sub_sync(int sub) {   // named barrier of one sub's 288 threads
    asm volatile("bar.sync %0, %1;" ::"r"(sub + 1), "r"(SUB_THREADS) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
template <int NC>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float* v) {
    if (NC == 8) tmem_ld8(taddr, v);
    else tmem_ld16(taddr, v);
}

__host__ __device__ constexpr uint32_t gr_idesc(int n) {   // kind::f16: D = f32, A = B = bf16, K-major, M = 128, N = n
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// All MMAs of one layer, issued by one thread: straight-line code per tile (the issuing thread's own instruction stream is
// what bounds the MMA rate of these small tiles: branches and address arithmetic per MMA cost more than the MMA).
//   alo0 = descriptor low word of tile 0's dx = 0 window (it starts one row before the tile); wlo = descriptor low word of the
//   layer's weights [dx][kc][k-half][3*COUT rows]; both advance in 16-byte units.
// Accumulators: a RING of R slots of COUT columns; output tile Y owns slot Y mod R at columns cbase + (R-1 - Y mod R)*COUT,
// so that the three blocks of input tile t (output tiles t+1, t, t-1) are consecutive ascending column blocks = ONE MMA of
// N = 3*COUT, except where the ring wraps (two MMAs).  The epilogue warps pre-load every slot with the bias of the tile that
// uses it next (tcgen05.st), so every MMA accumulates.  Waits: `ready` of the PREVIOUS layer (parity rpar) hands over tile t's
// data and the drained slots when the layers are chained tile by tile (chain; needs NT <= R), and `ready` of THIS layer
// (parity cpar) frees the slot of output tile t+1 when the image has more grid rows than the ring has slots.
// XM: 0 = plain bf16; 1 = split-bf16, every product is a_hi*w_hi + a_hi*w_lo + a_lo*w_hi (a_lo16 / w_lo16 = distance of the lo
// operands in 16-byte units); 2 = split-bf16 with an exact bf16 A operand (the 0/1 input planes): no a_lo term
template <int CIN16, int COUT, int R, bool TALL, bool PRE, int XM>
__device__ __forceinline__ void issue_layer(uint32_t alo0, uint32_t ahi, uint32_t a_kc, uint32_t TS, uint32_t wlo, uint32_t bhi,
                                            int NT, uint32_t cbase, uint32_t full, uint32_t ready, uint32_t rpar, uint32_t cpar,
                                            bool chain, uint32_t a_lo16, uint32_t w_lo16) {
    constexpr uint32_t id1 = gr_idesc(COUT), id2 = gr_idesc(2 * COUT), id3 = gr_idesc(3 * COUT);
    constexpr uint32_t blk = 6u * COUT;   // 16-byte units of one (dx, kc) block: 2 K halves x 3*COUT rows
    constexpr uint32_t top = (uint32_t)(R - 1) * COUT;
#define GR_MMA1(DCOL, AT, BT, ID, ACC)                                                                                         \
    do {                                                                                                                      \
        umma_bf16_lh((DCOL), (AT), ahi, (BT), bhi, (ID), (ACC));                                                              \
        if (XM) umma_bf16_lh((DCOL), (AT), ahi, (BT) + w_lo16, bhi, (ID), 1u);                                                \
        if (XM == 1) umma_bf16_lh((DCOL), (AT) + a_lo16, ahi, (BT), bhi, (ID), 1u);                                           \
    } while (0)
#define GR_MMAS(DCOL, BOFF, ID)                                                                                               \
    _Pragma("unroll") for (int i = 0; i < 3 * CIN16; ++i)                                                                     \
        GR_MMA1((DCOL), alo + (uint32_t)(i / CIN16) + (uint32_t)(i % CIN16) * a_kc, wlo + (uint32_t)i * blk + (BOFF), (ID), 1u)
    for (int t = 0; t < NT; ++t) {
        if (chain && t + 1 < NT) {
            mbar_wait(ready + 8u * (uint32_t)(t + 1), rpar);   // data of tile t (and t-1, t), slot of t+1 drained
            tc_fence_after();
        }
        const uint32_t alo = alo0 + (uint32_t)t * TS;
        if (!TALL) {   // every grid row has its own slot (NT <= R): no wrap, no slot is reused inside a layer
            const uint32_t c_a = cbase + top - (uint32_t)(t + 1) * COUT;   // columns of output tile t+1
            if (PRE) {
                if (NT == 1) { GR_MMAS(cbase + top, COUT, id1); }
                else if (t == 0) { GR_MMAS(c_a, 0u, id2); }                    // output tiles 1 and 0
                else if (t == NT - 1) { GR_MMAS(c_a + COUT, COUT, id2); }      // no tile below: dy = 1, 2 only
                else { GR_MMAS(c_a, 0u, id3); }
            } else {
                // accumulators not pre-loaded: the first MMA into a slot overwrites.  Tile 0 opens output tiles 1 and 0; a
                // middle tile's dy = 0 block opens output tile t+1 while its other two blocks accumulate: its first MMA is split
#define GR_MMAS0(DCOL, BOFF, ID, A0)                                                                                           \
    _Pragma("unroll") for (int i = 0; i < 3 * CIN16; ++i)                                                                     \
        GR_MMA1((DCOL), alo + (uint32_t)(i / CIN16) + (uint32_t)(i % CIN16) * a_kc, wlo + (uint32_t)i * blk + (BOFF), (ID), (i || (A0)) ? 1u : 0u)
                if (NT == 1) { GR_MMAS0(cbase + top, COUT, id1, 0); }
                else if (t == 0) { GR_MMAS0(c_a, 0u, id2, 0); }
                else if (t == NT - 1) { GR_MMAS0(c_a + COUT, COUT, id2, 1); }
                else {
                    umma_bf16_lh(c_a, alo, ahi, wlo, bhi, id1, 0u);
                    umma_bf16_lh(c_a + COUT, alo, ahi, wlo + COUT, bhi, id2, 1u);
                    if (XM) umma_bf16_lh(c_a, alo, ahi, wlo + w_lo16, bhi, id3, 1u);
                    if (XM == 1) umma_bf16_lh(c_a, alo + a_lo16, ahi, wlo, bhi, id3, 1u);
                    _Pragma("unroll") for (int i = 1; i < 3 * CIN16; ++i)
                        GR_MMA1(c_a, alo + (uint32_t)(i / CIN16) + (uint32_t)(i % CIN16) * a_kc, wlo + (uint32_t)i * blk, id3, 1u);
                }
#undef GR_MMAS0
            }
        } else {
            if (t + 1 >= R && t + 1 < NT) {
                mbar_wait(ready + 8u * (uint32_t)(t + 1 - R), cpar);   // the slot of output tile t+1 was tile t+1-R's
                tc_fence_after();
            }
            const uint32_t sa = (uint32_t)(t + 1) & (uint32_t)(R - 1), sb = (uint32_t)t & (uint32_t)(R - 1);
            const uint32_t c_a = cbase + top - sa * COUT;   // columns of output tile t+1
            if (NT == 1) {
                GR_MMAS(cbase + top, COUT, id1);
            } else if (t == 0) {              // output tiles 1 and 0
                GR_MMAS(c_a, 0u, id2);
            } else if (t == NT - 1) {         // no tile below: dy = 1, 2 only
                if (sb != 0) { GR_MMAS(cbase + top - sb * COUT, COUT, id2); }
                else { GR_MMAS(cbase + top, COUT, id1); GR_MMAS(cbase, 2u * COUT, id1); }
            } else if (sa == 0) {             // tile t+1 wrapped to the top slot; t and t-1 sit in the two lowest column blocks
                GR_MMAS(cbase + top, 0u, id1);
                GR_MMAS(cbase, COUT, id2);
            } else if (sb == 0) {             // t+1 and t are consecutive, t-1 wrapped to the lowest block
                GR_MMAS(c_a, 0u, id2);
                GR_MMAS(cbase, 2u * COUT, id1);
            } else {
                GR_MMAS(c_a, 0u, id3);
            }
        }
        umma_commit(full + 8u * (uint32_t)t);
    }
#undef GR_MMAS
#undef GR_MMA1
}

enum { GR_CONV = 0, GR_RES0 = 1, GR_RES1 = 2 };

// TMEM loads without the wait: the registers may be read only behind tmem_wait_dep on the same registers
__device__ __forceinline__ void tmem_ld8_nw(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nw(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// tcgen05.wait::ld with the loaded registers as in/out operands, so that no use of them is scheduled in front of the wait
__device__ __forceinline__ void tmem_wait_dep8(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])::"memory");
}
template <int NC>
__device__ __forceinline__ void tmem_ld_nw(uint32_t taddr, uint32_t* r) {
    if (NC == 8) tmem_ld8_nw(taddr, r);
    else tmem_ld16_nw(taddr, r);
}
template <int NC>
__device__ __forceinline__ void tmem_wait_dep(uint32_t* r) {
    tmem_wait_dep8(r);
    if (NC == 16) tmem_wait_dep8(r + 8);
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
template <int NC>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t* r) {
    tmem_st8(taddr, r);
    if (NC == 16) tmem_st8(taddr + 8u, r + 8);
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Epilogue of one (tile, column part) unit of a thread: NC accumulator columns = NC/8 planes of its pixel row.  `out`, `raw`
// = shared-space byte addresses of the row's entry in the first of those planes.  BIAS: the bias is added here (layers whose
// accumulators were not pre-loaded with it).  Branch-free up to the stores, so that the units of a batch interleave.
//   GR_CONV: out <- bf16(acc + bias)
//   GR_RES0: out <- relu(bf16(acc + bias))                                  (in place over the layer's input)
//   GR_RES1: v = acc + bias + raw; raw <- bf16(v); out <- relu(bf16(v))
template <int NC, int KIND, bool BIAS>
__device__ __forceinline__ void epi_unit(const uint32_t* acc, const float* bias_s, bool interior, uint32_t out, uint32_t raw,
                                         uint32_t PS) {
    float v[NC];
    if (BIAS) {
        const float4* b4 = reinterpret_cast<const float4*>(bias_s);
#pragma unroll
        for (int i = 0; i < NC / 4; ++i) {
            const float4 bb = b4[i];
            v[4 * i] = __uint_as_float(acc[4 * i]) + bb.x; v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + bb.y;
            v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + bb.z; v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + bb.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NC; ++i) v[i] = __uint_as_float(acc[i]);
    }
#pragma unroll
    for (int p = 0; p < NC / 8; ++p) {
        float* u = v + 8 * p;
        uint4 o;
        if (KIND == GR_RES1) {
            const uint4 rv = lds128(raw + (uint32_t)p * PS);
            u[0] += bf16_lo(rv.x); u[1] += bf16_hi(rv.x); u[2] += bf16_lo(rv.y); u[3] += bf16_hi(rv.y);
            u[4] += bf16_lo(rv.z); u[5] += bf16_hi(rv.z); u[6] += bf16_lo(rv.w); u[7] += bf16_hi(rv.w);
            o = make_uint4(pack_bf16(u[0], u[1]), pack_bf16(u[2], u[3]), pack_bf16(u[4], u[5]), pack_bf16(u[6], u[7]));
            if (interior) sts128(raw + (uint32_t)p * PS, o);
        } else {
            o = make_uint4(pack_bf16(u[0], u[1]), pack_bf16(u[2], u[3]), pack_bf16(u[4], u[5]), pack_bf16(u[6], u[7]));
        }
        if (KIND != GR_CONV) o = make_uint4(relu_bf16x2(o.x), relu_bf16x2(o.y), relu_bf16x2(o.z), relu_bf16x2(o.w));
        if (interior) sts128(out + (uint32_t)p * PS, o);
    }
}

// split-bf16 variant: every stored value is hi + lo (the lo entry LO bytes behind the hi one)
template <int NC, int KIND, bool BIAS>
__device__ __forceinline__ void epi_unit_x3(const uint32_t* acc, const float* bias_s, bool interior, uint32_t out, uint32_t raw,
                                            uint32_t PS, uint32_t LO) {
#pragma unroll
    for (int p = 0; p < NC / 8; ++p) {
        float u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = __uint_as_float(acc[8 * p + i]) + (BIAS ? bias_s[8 * p + i] : 0.f);
        uint4 h4, l4;
        if (KIND == GR_RES1) {
            const uint4 rh = lds128(raw + (uint32_t)p * PS), rl = lds128(raw + (uint32_t)p * PS + LO);
            u[0] += bf16_lo(rh.x) + bf16_lo(rl.x); u[1] += bf16_hi(rh.x) + bf16_hi(rl.x);
            u[2] += bf16_lo(rh.y) + bf16_lo(rl.y); u[3] += bf16_hi(rh.y) + bf16_hi(rl.y);
            u[4] += bf16_lo(rh.z) + bf16_lo(rl.z); u[5] += bf16_hi(rh.z) + bf16_hi(rl.z);
            u[6] += bf16_lo(rh.w) + bf16_lo(rl.w); u[7] += bf16_hi(rh.w) + bf16_hi(rl.w);
            split8(u, h4, l4);
            if (interior) {
                sts128(raw + (uint32_t)p * PS, h4);
                sts128(raw + (uint32_t)p * PS + LO, l4);
            }
        }
        if (KIND != GR_CONV) {
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = fmaxf(u[i], 0.f);
        }
        split8(u, h4, l4);
        if (interior) {
            sts128(out + (uint32_t)p * PS, h4);
            sts128(out + (uint32_t)p * PS + LO, l4);
        }
    }
}

// this warp's share of the first min(NT, R) accumulator slots <- the bias of the layer that is about to run
template <int NC, int R>
__device__ __forceinline__ void preload_bias(uint32_t tbase, int NT, uint32_t cout, const float* bias_s) {
    uint32_t bv[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) bv[i] = __float_as_uint(bias_s[i]);
    const int ns = NT < R ? NT : R;
    for (int sl = 0; sl < ns; ++sl) tmem_st<NC>(tbase + (uint32_t)(R - 1 - sl) * cout, bv);
    tmem_wait_st();
    tc_fence_before();
}

#ifdef BPP_GR_PROF
#define GR_T(slot) do { if (tprof) { const long long _t = clock64(); tprof[slot] += _t - tq; tq = _t; } } while (0)
#else
#define GR_T(slot) do { } while (0)
#endif
// All epilogue units of one layer for this warp, in batches of TB tiles (R % TB == 0: a batch never straddles the ring's
// wrap): one wait for the batch's last tile, the TMEM loads of the whole batch in flight together, one cross-proxy fence per
// batch, then one arrival per tile on the tiles' "ready" barriers.  `publish` = a later layer reads the stores through the
// tensor core's proxy.  A drained slot is re-loaded with the bias of its next user: this layer's (bias_s) while the image
// has tiles left for it (T + R < NT), else the next layer's (bias_next, nullptr = none).
//   tbase = TMEM address of this warp's lane quarter and column part in slot R-1 (column block 0); out0 / raw0 = shared
//   addresses of this thread's row in tile 0, first plane of its column part
template <int NC, int KIND, int TB, int R, bool BIAS, bool TALL, bool X3 = false>
__device__ __forceinline__ void epi_layer(int NT, uint32_t tile_bytes, uint32_t PS, uint32_t tbase, uint32_t full, uint32_t ready,
                                          uint32_t par, uint32_t cout, const float* bias_s, const float* bias_next, bool interior,
                                          uint32_t out0, uint32_t raw0, int lane, bool publish, long long* tprof, uint32_t LO = 0) {
    uint32_t bn[NC];
    if (bias_next) {
#pragma unroll
        for (int i = 0; i < NC; ++i) bn[i] = __float_as_uint(bias_next[i]);
    }
#ifdef BPP_GR_PROF
    long long tq = tprof ? clock64() : 0;
#endif
    uint32_t orow = out0, rrow = raw0;
    for (int T0 = 0; T0 < NT; T0 += TB) {
        const int nb = min(TB, NT - T0);
        const int fb = min(T0 + nb, NT - 1);   // the MMAs of input tile T+1 complete output tile T
        mbar_wait(full + 8u * (uint32_t)fb, par);
        tc_fence_after();
        GR_T(0);
        const uint32_t tcol = tbase + (uint32_t)(R - 1 - (T0 & (R - 1))) * cout;   // slot of tile T0; T0 + b sits b blocks lower
        uint32_t acc[TB][NC];
#pragma unroll
        for (int b = 0; b < TB; ++b)
            if (b < nb) tmem_ld_nw<NC>(tcol - (uint32_t)b * cout, acc[b]);   // warp-collective
#pragma unroll
        for (int b = 0; b < TB; ++b)
            if (b < nb) {
                tmem_wait_dep<NC>(acc[b]);
                if (b == 0) GR_T(1);
                if (TALL && T0 + b + R < NT) {   // the image is taller than the ring: this layer uses the slot again
                    uint32_t bt[NC];
#pragma unroll
                    for (int i = 0; i < NC; ++i) bt[i] = __float_as_uint(bias_s[i]);
                    tmem_st<NC>(tcol - (uint32_t)b * cout, bt);
                } else if (bias_next) tmem_st<NC>(tcol - (uint32_t)b * cout, bn);
                if (X3) epi_unit_x3<NC, KIND, BIAS>(acc[b], bias_s, interior, orow + (uint32_t)b * tile_bytes, rrow + (uint32_t)b * tile_bytes, PS, LO);
                else epi_unit<NC, KIND, BIAS>(acc[b], bias_s, interior, orow + (uint32_t)b * tile_bytes, rrow + (uint32_t)b * tile_bytes, PS);
            }
        GR_T(2);
        if (publish) fence_proxy_async();   // this thread's stores -> visible to the tensor core's reads of the next layer
        if (bias_next || TALL) tmem_wait_st();
        GR_T(3);
        tc_fence_before();                  // its TMEM accesses are ordered before the MMAs that will use the columns
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int b = 0; b < TB; ++b)
                if (b < nb) mbar_arrive(ready + 8u * (uint32_t)(T0 + b));
        }
        GR_T(4);
        orow += (uint32_t)TB * tile_bytes;
        rrow += (uint32_t)TB * tile_bytes;
    }
}

// barriers and small tables of a CTA (static shared memory, shared by the four stages of the fused kernel)
struct GrShared {
    uint64_t full[2][MAX_TILES];    // [sub][tile]: the MMAs of input tile t are complete
    uint64_t ready[2][MAX_TILES];   // [sub][tile]: output tile t stored, its accumulator slot drained (8 warp arrivals)
    uint64_t wbar;                  // resident weights have landed
    uint64_t cbar[2];               // conv bias pre-loaded (8 warps), per sub
    uint64_t wslot[2];              // streamed weights: slot filled
    uint32_t tmem;
    long long tprof[8];
    uint16_t rowmap[128];           // row r of a tile -> leaf j | xp << 8 (0xffff: not a pixel row)
    float bias[15][32];             // biases of all conv layers (fetched once per launch), global layer index
};
// stage 0's per-group tables live in dynamic shared memory behind its arenas (12.3 KB)
struct GrStage0Tables {
    uint32_t rec[2][8][32];                  // [sub][leaf]: compact record
    uint32_t msk[2][8][2][32];               // [sub][leaf][rows | columns][index]: items reaching grid row y / column x
    int it[2][8][2 * BPP_MAX_ITEMS];         // [sub][leaf][item: w, h]
    uint4 lut[256];                          // 8 channel bits -> 8 x bf16 {0, 1}
};

// One level of the trunk for this CTA's share of the batch.  X3 = split-bf16 mode: activations and weights as hi + lo bf16
// halves, three MMAs per product, fp32-level accuracy for the reference's trained checkpoints; one group per
// CTA (the doubled arena leaves no room for two).  Called by every thread of the CTA; `first` = the CTA has not run a stage
// yet (barriers are fresh, nothing to wait for).
template <int STAGE, bool X3>
__device__ __forceinline__ void gr_stage(const NetParams& P, const GrStage& S, int B, const uint32_t* __restrict__ recs,
                                         const int32_t* __restrict__ game, const int32_t* __restrict__ items_wh,
                                         const uint4* xin, uint4* xout, __nv_bfloat16* __restrict__ feat_out,
                                         long long feat_lo_off, const __nv_bfloat16* __restrict__ wts_gr,
                                         const __nv_bfloat16* __restrict__ wts_gr_lo, long long* prof, unsigned char* smem,
                                         GrShared& sh, bool first) {
    GrStage0Tables& t0 = *reinterpret_cast<GrStage0Tables*>(smem + S.arena_off + (size_t)S.nsub * S.arena_bytes);
    const long long t_enter = clock64();
    const int tid = threadIdx.x;
    const int sub = tid >= SUB_THREADS ? 1 : 0;
    const int st = tid - sub * SUB_THREADS;          // thread inside the sub
    const int warp_s = st >> 5, lane = tid & 31;
    const int nthreads = (int)blockDim.x;   // (a stage that runs one group per CTA leaves the second sub idle)
    float* s_bias = &sh.bias[STAGE == 0 ? 0 : 5 * (STAGE - 1) + 1][0];   // [nlay][32], this stage's layers
    unsigned char* arena = smem + S.arena_off + (size_t)sub * S.arena_bytes;
    const uint32_t PS = (uint32_t)S.RT * 16u;
    const uint32_t LO = (uint32_t)S.lo_off;   // split mode: lo planes behind the hi planes

    // every stage starts with fresh barriers (its own phase counting)
    if (!first) {
        __syncthreads();
        if (tid < 2 * MAX_TILES) {
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&sh.full[0][0]) + 8u * (uint32_t)tid) : "memory");
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&sh.ready[0][0]) + 8u * (uint32_t)tid) : "memory");
        }
        if (tid == 0) {
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&sh.wbar)) : "memory");
            for (int k = 0; k < 2; ++k) {
                asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&sh.cbar[k])) : "memory");
                asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&sh.wslot[k])) : "memory");
            }
        }
        __syncthreads();
    }
    if (tid < 2 * MAX_TILES) {
        mbar_init(smem_u32(&sh.full[0][0]) + 8u * (uint32_t)tid, 1);
        mbar_init(smem_u32(&sh.ready[0][0]) + 8u * (uint32_t)tid, 8);
    }
    if (tid == 0) {
        mbar_init(smem_u32(&sh.wbar), 1); mbar_init(smem_u32(&sh.cbar[0]), 8); mbar_init(smem_u32(&sh.cbar[1]), 8);
        mbar_init(smem_u32(&sh.wslot[0]), 1); mbar_init(smem_u32(&sh.wslot[1]), 1);
    }
    if (tid < 128) {
        const int j = tid / S.wp, xp = tid - j * S.wp;
        sh.rowmap[tid] = (tid < S.J * S.wp && xp != 0) ? (uint16_t)(j | (xp << 8)) : (uint16_t)0xffff;
    }
    if (STAGE == 0 && tid < 256) {
        const uint32_t b8 = (uint32_t)tid;
        uint4 v;
        v.x = ((b8 & 1u) | ((b8 & 2u) << 15)) * 0x3f80u;          // bf16 1.0 pairs
        v.y = (((b8 >> 2) & 1u) | ((b8 & 8u) << 13)) * 0x3f80u;
        v.z = (((b8 >> 4) & 1u) | ((b8 & 32u) << 11)) * 0x3f80u;
        v.w = (((b8 >> 6) & 1u) | ((b8 & 128u) << 9)) * 0x3f80u;
        t0.lut[tid] = v;
    }
    // halo rows, guards and padding rows are zero from here on: no epilogue, input or hand-over pass ever writes them
    {
        uint4* q = reinterpret_cast<uint4*>(smem + S.arena_off);
        const int n16 = S.nsub * S.arena_bytes / 16;
        for (int i = tid; i < n16; i += nthreads) q[i] = make_uint4(0, 0, 0, 0);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // weights: resident for the whole kernel (one bulk async copy per layer now), or - split mode, where they do not fit next
    // to the doubled arena - streamed through two slots, one layer ahead (the issuing thread refills, see below)
    auto fetch_layer = [&](int l, uint32_t dst, uint32_t bar) {   // one thread
        bulk_g2s(dst, wts_gr + S.w_goff[l], (uint32_t)S.w_len[l], bar);
        if (X3) bulk_g2s(dst + (uint32_t)S.w_len[l], wts_gr_lo + S.w_goff[l], (uint32_t)S.w_len[l], bar);
    };
    if (tid == 0) {
        fence_proxy_async();   // the previous stage used this memory through the generic proxy
        if (!S.stream) {
            const uint32_t wb = smem_u32(&sh.wbar);
            mbar_expect_tx(wb, (uint32_t)S.w_bytes);
            for (int l = 0; l < S.nlay; ++l) fetch_layer(l, smem_u32(smem + S.w_soff[l]), wb);
        } else {
            for (int g = 0; g < 2 && g < S.nlay; ++g) {
                mbar_expect_tx(smem_u32(&sh.wslot[g]), (uint32_t)S.w_len[g] * (X3 ? 2u : 1u));
                fetch_layer(g, smem_u32(smem + g * S.slot_bytes), smem_u32(&sh.wslot[g]));
            }
        }
    }
    const long long t_start = clock64();   // (only the issuing threads wait for the weights, before their first MMA)
    long long t_in = 0, t_cv = 0, t_out = 0;

    // this CTA's share of the batch is the same contiguous range of leaves in every stage (its subs split it), so a stage
    // only ever reads hand-over data written by its own CTA
    const int nwork = gridDim.x * S.nsub, wid = blockIdx.x * S.nsub + sub;
    const int slice_lo = (int)(((long long)wid * B) / nwork);
    const int slice_hi = (int)(((long long)(wid + 1) * B) / nwork);
    const int n_slice = slice_hi - slice_lo;
    const int n_groups = (n_slice + S.J - 1) / S.J;
    const int gsz = n_groups > 0 ? (n_slice + n_groups - 1) / n_groups : S.J;

    const uint32_t tmem_sub = sh.tmem + (uint32_t)(sub * 256);
    const uint32_t full = smem_u32(&sh.full[sub][0]), ready = smem_u32(&sh.ready[sub][0]);
    const int NT = S.NT;
    // epilogue warps: this thread's pixel row inside every tile
    // a warp reads the TMEM lane quarter of its CTA-wide warp index (sub 1 starts at warp 9): the eight epilogue warps of a
    // sub cover every (quarter, half) pair once either way
    const int quarter = (tid >> 5) & 3, half = (warp_s >> 2) & 1;
    const int r = quarter * 32 + lane;
    const uint32_t rm = sh.rowmap[r];
    const int my_j = (int)(rm & 0xff);
    uint32_t lc = 0;   // layers this sub has run (parity of its barriers)
    uint32_t cbar_par = 0;
    const bool profiling = prof != nullptr && blockIdx.x == 0 && tid == 0;
    long long* tprof = profiling ? sh.tprof : nullptr;
    if (profiling) for (int i = 0; i < 8; ++i) sh.tprof[i] = 0;

    // stage 0: the compact record words and item-list entries this thread fetches ahead for its sub's next group
    uint32_t pf_rec = 0;
    int pf_it = 0, pf_j = 0, pf_q = 0;
    const int n2 = 2 * P.N;
    if (STAGE == 0 && sub < S.nsub) {
        pf_j = st / n2;
        pf_q = st - pf_j * n2;
        const int nv = min(gsz, slice_hi - slice_lo);
        if (st < nv * 32) pf_rec = recs[(size_t)(slice_lo + (st >> 5)) * 32 + (st & 31)];
        if (st < nv * n2) {
            const int g = game ? game[slice_lo + pf_j] : slice_lo + pf_j;
            pf_it = items_wh[(size_t)g * n2 + pf_q];
        }
    }
    if (sub < S.nsub)
    for (int b0 = slice_lo; b0 < slice_hi; b0 += gsz) {
        const int nvalid = min(gsz, slice_hi - b0);
        long long tq = profiling ? clock64() : 0;
        // ---------------------------------------------------------------------------------------------------- input
        if (STAGE == 0) {
            // input planes (getBinItem, BinPackingGame.py:118-120): channel 0 = bin occupancy, channel i+1 = item i's
            // [0:h, 0:w] block while it is still to be placed.  Per leaf two small tables - items still to place that reach
            // grid row y, items that reach column x - make a pixel's channel bits one AND; an 8-channel plane entry (bf16 0/1)
            // comes from a 256-entry table.
            // (the records and item lists were fetched into registers while the previous group was being computed)
            if (st < nvalid * 32) t0.rec[sub][st >> 5][st & 31] = pf_rec;
            if (st < nvalid * n2) t0.it[sub][pf_j][pf_q] = pf_it;
            sub_sync(sub);
            for (int i = st; i < nvalid * 64; i += SUB_THREADS) {
                const int j = i >> 6, k = i & 31, isx = (i >> 5) & 1;   // k = grid row y (isx = 0) or column x (isx = 1)
                const uint32_t rem = t0.rec[sub][j][BPP_REC_REM];
                uint32_t m = 0;
                for (int q = 0; q < P.N; ++q)
                    if (((rem >> q) & 1u) && k < t0.it[sub][j][2 * q + (isx ? 0 : 1)]) m |= 2u << q;
                t0.msk[sub][j][isx][k] = m;
            }
            {   // next group's inputs: in flight during this group's layers
                const int nb0 = b0 + gsz, nv = min(gsz, slice_hi - nb0);
                if (st < nv * 32) pf_rec = recs[(size_t)(nb0 + (st >> 5)) * 32 + (st & 31)];
                if (st < nv * n2) {
                    const int g = game ? game[nb0 + pf_j] : nb0 + pf_j;
                    pf_it = items_wh[(size_t)g * n2 + pf_q];
                }
            }
            sub_sync(sub);
            for (int idx = st; idx < NT * 128; idx += SUB_THREADS) {   // one pixel row entry of one tile per step
                const int t = idx >> 7;
                const uint32_t m = sh.rowmap[idx & 127];
                const int j = (int)(m & 0xff);
                if (m == 0xffffu || j >= nvalid) continue;
                const int x = (int)(m >> 8) - 1;
                const uint32_t bits = ((t0.rec[sub][j][t] >> x) & 1u) | (t0.msk[sub][j][0][t] & t0.msk[sub][j][1][x]);
                const uint32_t dst = smem_u32(arena) + (uint32_t)(G0 + t * S.TS + (idx & 127)) * 16u;
                sts128(dst, t0.lut[bits & 0xffu]);
                sts128(dst + PS, t0.lut[(bits >> 8) & 0xffu]);
                if (S.cp > 2) {
                    sts128(dst + 2u * PS, t0.lut[(bits >> 16) & 0xffu]);
                    sts128(dst + 3u * PS, t0.lut[0]);
                }
            }
        } else {
            // the previous stage's residual stream [leaf][hi | lo][plane][pixel] -> raw planes, relu(raw) -> activation planes;
            // the loads of LR rows are in flight together
            const int hw = S.h * S.w;
            constexpr int CPX = STAGE == 1 ? 2 : 4, LR = (STAGE == 1 && !X3) ? 2 : 1, FX = X3 ? 2 : 1;
            for (int i0 = st; i0 < NT * 128; i0 += LR * SUB_THREADS) {
                uint4 v[LR][FX * CPX];
                uint32_t dst[LR];
#pragma unroll
                for (int k = 0; k < LR; ++k) {
                    const int idx = i0 + k * SUB_THREADS;
                    dst[k] = 0;
                    if (idx < NT * 128) {
                        const int t = idx >> 7;
                        const uint32_t m = sh.rowmap[idx & 127];
                        const int j = (int)(m & 0xff);
                        if (m != 0xffffu && j < nvalid) {
                            const uint4* src = xin + (size_t)(b0 + j) * (FX * CPX) * hw + t * S.w + ((int)(m >> 8) - 1);
                            dst[k] = smem_u32(arena) + (uint32_t)(G0 + t * S.TS + (idx & 127)) * 16u;
#pragma unroll
                            for (int p = 0; p < FX * CPX; ++p) v[k][p] = __ldcg(src + (size_t)p * hw);   // written by this kernel: L2, not the read-only path
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < LR; ++k)
                    if (dst[k]) {
#pragma unroll
                        for (int p = 0; p < CPX; ++p) {
                            const uint32_t d = dst[k] + (uint32_t)p * PS;
                            sts128(d, v[k][p]);
                            if (!X3) {
                                sts128(d + (uint32_t)CPX * PS, make_uint4(relu_bf16x2(v[k][p].x), relu_bf16x2(v[k][p].y),
                                                                          relu_bf16x2(v[k][p].z), relu_bf16x2(v[k][p].w)));
                            } else {
                                const uint4 lo = v[k][CPX + p];
                                sts128(d + LO, lo);
                                float a[8], c[8];
                                unpack8(v[k][p], a);
                                unpack8(lo, c);
#pragma unroll
                                for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i] + c[i], 0.f);
                                uint4 h4, l4;
                                split8(a, h4, l4);
                                sts128(d + (uint32_t)CPX * PS, h4);
                                sts128(d + (uint32_t)CPX * PS + LO, l4);
                            }
                        }
                    }
            }
        }
        constexpr int NCR = STAGE <= 1 ? 8 : 16;          // columns per warp in the residual layers (and stage 0's conv)
        constexpr int RR = STAGE <= 1 ? 16 : 8;           // accumulator ring: 256 columns per sub
        constexpr uint32_t CO = STAGE <= 1 ? 16u : 32u;
        const uint32_t tq_ = tmem_sub + ((uint32_t)(quarter * 32) << 16);
        if (warp_s < 8)   // accumulator slots of the group's first layer <- its bias
            preload_bias<NCR, RR>(tq_ + half * NCR, NT, CO, s_bias + half * NCR);
        fence_proxy_async();
        sub_sync(sub);
        if (profiling) { const long long t_ = clock64(); t_in += t_ - tq; tq = t_; }
        // --------------------------------------------------------------------------------------------------- layers
        const int nres = STAGE == 0 ? 0 : 4;
        const bool has_conv = STAGE != 3;
        if (warp_s == 8) {
            if (elect_one()) {
                tc_fence_after();
                const uint32_t ahi = (uint32_t)(umma_desc(0, (uint32_t)S.RT, 8u) >> 32);
                const uint32_t a_lbo = ((uint32_t)S.RT & 0x3fffu) << 16;
                const uint32_t arena_a = smem_u32(arena);
                uint32_t lcl = lc;
                const uint32_t a_lo16 = LO >> 4;
                if (!S.stream && lc == 0) mbar_wait(smem_u32(&sh.wbar), 0);   // resident weights have landed
                const uint32_t nl_total = (uint32_t)n_groups * (uint32_t)S.nlay;   // layers this CTA runs in all (streaming)
                for (int l = 0; l < nres + (has_conv ? 1 : 0); ++l, ++lcl) {
                    const bool conv = l == nres;
                    // residual layers read the activation planes, the sequence's first conv the raw stream (stage 0: input)
                    const uint32_t abase = arena_a + ((conv || STAGE == 0) ? 0u : (uint32_t)S.cp * PS);
                    const uint32_t alo0 = (((abase + (uint32_t)(G0 - 1) * 16u) >> 4) & 0x3fffu) | a_lbo;
                    const uint32_t w_sm = smem_u32(smem) + (S.stream ? (lcl & 1u) * (uint32_t)S.slot_bytes : (uint32_t)S.w_soff[l]);
                    const uint64_t bd = umma_desc(w_sm, 3u * (uint32_t)S.cout[l], 8u);
                    const uint32_t wlo = (uint32_t)bd, bhi = (uint32_t)(bd >> 32);
                    const uint32_t w_lo16 = (uint32_t)S.w_len[l] >> 4;
                    const uint32_t a_kc = 2u * (uint32_t)S.RT;
                    const uint32_t rpar = (lcl - 1u) & 1u, cpar = lcl & 1u;
                    // layers are chained tile by tile while every grid row has its own accumulator slot; the sequence's first
                    // conv has another column map and starts when every slot of the last residual layer is drained: with up to
                    // 8 grid rows its first MMAs overwrite the slots and the bias is added in the epilogue, a taller image
                    // needs the ring (wraps) and gets the bias pre-loaded behind sh.cbar.  Streamed weights: no chaining, the
                    // previous layer is complete when this one starts, so its slot is refilled with the layer after this one.
                    const bool chain = l > 0 && !conv && NT <= RR && !S.stream;
                    if (l > 0) {
                        if (conv && NT > 8) {   // (a conv of at most 8 grid rows opens its slots itself, see below)
                            mbar_wait(smem_u32(&sh.cbar[sub]), cbar_par);
                            cbar_par ^= 1u;
                        } else if (chain) mbar_wait(ready, rpar);
                        else
                            for (int t = 0; t < NT; ++t) mbar_wait(ready + 8u * (uint32_t)t, rpar);
                        tc_fence_after();
                    }
                    if (S.stream) {
                        if (lcl >= 1u && lcl + 1u < nl_total) {
                            const uint32_t sl = (lcl + 1u) & 1u, nl = (lcl + 1u) % (uint32_t)S.nlay;
                            mbar_expect_tx(smem_u32(&sh.wslot[sl]), (uint32_t)S.w_len[nl] * (X3 ? 2u : 1u));
                            fetch_layer((int)nl, smem_u32(smem) + sl * (uint32_t)S.slot_bytes, smem_u32(&sh.wslot[sl]));
                        }
                        mbar_wait(smem_u32(&sh.wslot[lcl & 1u]), (lcl >> 1) & 1u);
                    }
#define GR_ISSUE(C16, CO_, R_, PRE_, XM_)                                                                                     \
    do {                                                                                                                      \
        if (NT <= (R_)) issue_layer<C16, CO_, R_, false, PRE_, XM_>(alo0, ahi, a_kc, (uint32_t)S.TS, wlo, bhi, NT, tmem_sub, full, ready, rpar, cpar, chain, a_lo16, w_lo16); \
        else issue_layer<C16, CO_, R_, true, true, XM_>(alo0, ahi, a_kc, (uint32_t)S.TS, wlo, bhi, NT, tmem_sub, full, ready, rpar, cpar, chain, a_lo16, w_lo16);           \
    } while (0)
#define GR_ISSUE_FLAT(C16, CO_, R_, PRE_, XM_) /* layers whose grid rows always fit the ring (guaranteed by the plan) */       \
    issue_layer<C16, CO_, R_, false, PRE_, XM_>(alo0, ahi, a_kc, (uint32_t)S.TS, wlo, bhi, NT, tmem_sub, full, ready, rpar, cpar, chain, a_lo16, w_lo16)
                    constexpr int XMR = X3 ? 1 : 0;
                    if (STAGE == 0) {   // the 0/1 input planes are exact in bf16: no a_lo term in the network's first layer
                        if (S.cin16[0] == 1) GR_ISSUE(1, 16, 16, true, (X3 ? 2 : 0));
                        else GR_ISSUE(2, 16, 16, true, (X3 ? 2 : 0));
                    } else if (STAGE == 1) {
                        if (!conv) GR_ISSUE_FLAT(1, 16, 16, true, XMR);
                        else GR_ISSUE(1, 32, 8, false, XMR);
                    } else if (!conv) GR_ISSUE_FLAT(2, 32, 8, true, XMR);
                    else GR_ISSUE_FLAT(2, 32, 8, false, XMR);
#undef GR_ISSUE_FLAT
#undef GR_ISSUE
                }
            }
            __syncwarp();
        } else {
            const bool interior = rm != 0xffffu && my_j < nvalid;
            uint32_t lcl = lc;
            // every layer's accumulator slots are pre-loaded with its bias; a drained slot is re-loaded for its next user
            const uint32_t tile_bytes = (uint32_t)S.TS * 16u;
            const uint32_t row_a = smem_u32(arena) + (uint32_t)(G0 + r) * 16u;
            const uint32_t raw_r = row_a + (uint32_t)(half * (NCR / 8)) * PS;            // raw planes of this warp's column part
            const uint32_t act_r = raw_r + (uint32_t)S.cp * PS;
            const float* bs = s_bias + half * NCR;
            if (STAGE == 0) {
                if (NT <= RR) epi_layer<8, GR_CONV, 2, RR, false, false, X3>(NT, tile_bytes, PS, tq_ + half * 8, full, ready, lcl & 1u, 16u, bs, nullptr, interior, raw_r, raw_r, lane, false, tprof, LO);
                else epi_layer<8, GR_CONV, 2, RR, false, true, X3>(NT, tile_bytes, PS, tq_ + half * 8, full, ready, lcl & 1u, 16u, bs, nullptr, interior, raw_r, raw_r, lane, false, tprof, LO);
            } else {
                constexpr int TBR = STAGE == 1 ? 2 : BPP_GR_TB23;
                for (int blk = 0; blk < 2; ++blk) {
                    epi_layer<NCR, GR_RES0, TBR, RR, false, false, X3>(NT, tile_bytes, PS, tq_ + half * NCR, full, ready, lcl & 1u, CO, bs + 32 * (2 * blk), bs + 32 * (2 * blk + 1), interior, act_r, raw_r, lane, true, tprof, LO);
                    ++lcl;
                    epi_layer<NCR, GR_RES1, TBR, RR, false, false, X3>(NT, tile_bytes, PS, tq_ + half * NCR, full, ready, lcl & 1u, CO, bs + 32 * (2 * blk + 1), blk == 0 ? bs + 32 * 2 : nullptr, interior, act_r, raw_r, lane, STAGE != 3 || blk == 0, tprof, LO);
                    ++lcl;
                }
                if (STAGE != 3) {
                    // the sequence's first conv: 32 output channels, 16 columns = 2 planes per warp, its own column map (ring
                    // of 8 slots)
                    const uint32_t t_r = row_a + (uint32_t)(half * 2) * PS;
                    const float* bc = s_bias + 32 * 4 + half * 16;
                    if (NT > 8) {
                        // all eight warps must have drained the residual layers' slots before any of them pre-loads the
                        // conv's bias; the issuer starts the conv behind sh.cbar
                        asm volatile("bar.sync %0, 256;" ::"r"(3 + sub) : "memory");
                        preload_bias<16, 8>(tq_ + half * 16, NT, 32u, bc);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&sh.cbar[sub]));
                        epi_layer<16, GR_CONV, 1, 8, false, true, X3>(NT, tile_bytes, PS, tq_ + half * 16, full, ready, lcl & 1u, 32u, bc, nullptr, interior, t_r, t_r, lane, false, tprof, LO);
                    } else
                        epi_layer<16, GR_CONV, 1, 8, true, false, X3>(NT, tile_bytes, PS, tq_ + half * 16, full, ready, lcl & 1u, 32u, bc, nullptr, interior, t_r, t_r, lane, false, tprof, LO);
                }
            }
        }
        lc += (uint32_t)(nres + (has_conv ? 1 : 0));
        tc_fence_before();
        sub_sync(sub);
        tc_fence_after();
        if (profiling) { const long long t_ = clock64(); t_cv += t_ - tq; tq = t_; }
        // --------------------------------------------------------------------------------------------------- output
        if (STAGE < 3) {
            // max_pool2d(3, 2, 1) of the conv output T (planes 0.. of the arena, interior rows) -> x[leaf][plane][pixel].  One
            // thread = one output column of one plane of one leaf: the horizontal 3-max of every input row is taken once and
            // shared by the two output rows that use it.  A tap outside the image is replaced by its clamped neighbour, which
            // lies inside the same window (the maximum is unchanged, no branches).
            const int hw2 = S.h2 * S.w2, per = S.planes_out * hw2, pw2 = S.planes_out * S.w2;
            const uint32_t abase = smem_u32(arena) + (uint32_t)(G0 + 1) * 16u;   // entry of leaf 0, column 0 in tile 0
            const uint32_t trow = (uint32_t)S.TS * 16u;
            for (int it = st; it < nvalid * pw2; it += SUB_THREADS) {
                const int j = fdiv(it, S.m_pw2);
                int q = it - j * pw2;
                const int p = fdiv(q, S.m_w2), ox = q - p * S.w2;
                const uint32_t a0 = abase + (uint32_t)p * PS + (uint32_t)(j * S.wp) * 16u;
                const uint32_t c0 = a0 + (uint32_t)max(2 * ox - 1, 0) * 16u, c1 = a0 + (uint32_t)(2 * ox) * 16u,
                               c2 = a0 + (uint32_t)min(2 * ox + 1, S.w - 1) * 16u;
                if (!X3) {
                    auto hmax = [&](int y) {
                        const uint32_t ro = (uint32_t)y * trow;
                        const uint4 a = lds128(c0 + ro), b = lds128(c1 + ro), c = lds128(c2 + ro);
                        return make_uint4(max_bf16x2(max_bf16x2(a.x, b.x), c.x), max_bf16x2(max_bf16x2(a.y, b.y), c.y),
                                          max_bf16x2(max_bf16x2(a.z, b.z), c.z), max_bf16x2(max_bf16x2(a.w, b.w), c.w));
                    };
                    uint4* dst = xout + (size_t)(b0 + j) * per + p * hw2 + ox;
                    uint4 up = hmax(0);
                    for (int oy = 0; oy < S.h2; ++oy) {
                        const uint4 mid = oy ? hmax(2 * oy) : up;
                        const uint4 lo = hmax(min(2 * oy + 1, S.h - 1));
                        dst[oy * S.w2] = make_uint4(max_bf16x2(max_bf16x2(up.x, mid.x), lo.x), max_bf16x2(max_bf16x2(up.y, mid.y), lo.y),
                                                    max_bf16x2(max_bf16x2(up.z, mid.z), lo.z), max_bf16x2(max_bf16x2(up.w, mid.w), lo.w));
                        up = lo;
                    }
                } else {
                    // split mode: values are hi + lo; hand-over layout [leaf][hi | lo][plane][pixel]
                    struct F8 { float f[8]; };
                    auto ld8 = [&](uint32_t a) {
                        F8 r, q2;
                        unpack8(lds128(a), r.f);
                        unpack8(lds128(a + LO), q2.f);
#pragma unroll
                        for (int i = 0; i < 8; ++i) r.f[i] += q2.f[i];
                        return r;
                    };
                    auto hmax = [&](int y) {
                        const uint32_t ro = (uint32_t)y * trow;
                        F8 a = ld8(c0 + ro);
                        const F8 b = ld8(c1 + ro), c = ld8(c2 + ro);
#pragma unroll
                        for (int i = 0; i < 8; ++i) a.f[i] = fmaxf(fmaxf(a.f[i], b.f[i]), c.f[i]);
                        return a;
                    };
                    uint4* dst = xout + (size_t)(b0 + j) * (2 * per) + p * hw2 + ox;
                    F8 up = hmax(0);
                    for (int oy = 0; oy < S.h2; ++oy) {
                        const F8 mid = oy ? hmax(2 * oy) : up;
                        const F8 lo = hmax(min(2 * oy + 1, S.h - 1));
                        float m[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) m[i] = fmaxf(fmaxf(up.f[i], mid.f[i]), lo.f[i]);
                        uint4 h4, l4;
                        split8(m, h4, l4);
                        dst[oy * S.w2] = h4;
                        dst[per + oy * S.w2] = l4;
                        up = lo;
                    }
                }
            }
        } else {
            // relu(flatten(x)) as bf16 [B][flat] for the FC heads (flat index = channel * h*w + y*w + x)
            const int hw = S.h * S.w;
            const uint32_t m_hw = fdiv_magic((uint32_t)hw);
            for (int idx = st; idx < nvalid * P.flat; idx += SUB_THREADS) {
                const int j = fdiv(idx, S.m_flat), f = idx - j * P.flat;
                const int c = fdiv(f, m_hw), q = f - c * hw;
                const int y = fdiv(q, S.m_w), x = q - y * S.w;
                const size_t row = (size_t)G0 + (size_t)y * S.TS + (size_t)j * S.wp + (x + 1);
                const size_t eoff = ((size_t)(c >> 3) * S.RT + row) * 16;
                const uint16_t e = *(reinterpret_cast<const uint16_t*>(arena + eoff) + (c & 7));
                if (!X3) {
                    reinterpret_cast<uint16_t*>(feat_out)[(size_t)(b0 + j) * P.flat + f] = (e & 0x8000u) ? (uint16_t)0 : e;
                } else {   // hi, and lo feat_lo_off elements behind it
                    const uint16_t el = *(reinterpret_cast<const uint16_t*>(arena + LO + eoff) + (c & 7));
                    const float v = fmaxf(__uint_as_float((uint32_t)e << 16) + __uint_as_float((uint32_t)el << 16), 0.f);
                    const uint32_t hb = pack_bf16(v, 0.f) & 0xffffu;
                    const uint32_t lb = pack_bf16(v - __uint_as_float(hb << 16), 0.f) & 0xffffu;
                    reinterpret_cast<uint16_t*>(feat_out)[(size_t)(b0 + j) * P.flat + f] = (uint16_t)hb;
                    reinterpret_cast<uint16_t*>(feat_out)[(size_t)feat_lo_off + (size_t)(b0 + j) * P.flat + f] = (uint16_t)lb;
                }
            }
        }
        sub_sync(sub);
        if (profiling) { const long long t_ = clock64(); t_out += t_ - tq; tq = t_; }
    }
    if (profiling) {
        prof[8 * STAGE + 0] = t_in;
        prof[8 * STAGE + 1] = t_cv;
        prof[8 * STAGE + 2] = t_out;
#ifdef BPP_GR_PROF
        for (int i = 0; i < 4; ++i) prof[8 * STAGE + 3 + i] = sh.tprof[i] + (i == 3 ? sh.tprof[4] : 0);
#endif
        prof[8 * STAGE + 7] = clock64() - t_start;
        prof[8 * STAGE + 3] = t_start - t_enter;   // the stage's prologue (barriers, tables, arena clear, weight fetch issue)
    }
    __threadfence();       // the hand-over written by this stage is read by the CTA's next stage
    fence_proxy_async();   // this thread's shared-memory writes are ordered before the next stage's bulk weight copies
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (profiling) prof[8 * STAGE + 4] = clock64() - t_enter;   // whole stage incl. prologue and the closing fence + barrier
}

// The whole trunk in ONE launch: every CTA takes the same contiguous share of the batch through the four levels (no
// dependency between CTAs, hence no grid-wide barrier and no launch boundary between the levels: four launches cost ~4 us
// each in prologue, tail imbalance and launch latency - 15 % of a forward at the lockstep batch sizes).  Each level re-plans
// the shared memory (its weights, its arenas) and re-initialises the barriers; TMEM is allocated once.
template <bool X3>
__global__ void __launch_bounds__(X3 ? SUB_THREADS : 2 * SUB_THREADS, 1)
k_net_gr(NetParams P, GrStage S0, GrStage S1, GrStage S2, GrStage S3, int Bmax, const int32_t* __restrict__ count_dev,
         const uint32_t* __restrict__ recs, const int32_t* __restrict__ game, const int32_t* __restrict__ items_wh, uint4* x1,
         uint4* x2, uint4* x3, __nv_bfloat16* __restrict__ feat_out, long long feat_lo_off,
         const __nv_bfloat16* __restrict__ wts_gr, const __nv_bfloat16* __restrict__ wts_gr_lo, long long* prof) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) GrShared sh;
    const int tid = threadIdx.x;
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem)),
                     "r"((uint32_t)S0.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    for (int i = tid; i < 15 * 32; i += (int)blockDim.x) {
        const int l = i >> 5, c = i & 31;
        sh.bias[l][c] = c < P.conv[l].co ? P.bias[P.conv[l].b_off + c] : 0.f;
    }
    // nothing above depends on the kernel that produced the leaf batch; the heads kernel behind us may start launching
    pdl_launch_dependents();
    pdl_wait();
    const int B = count_dev ? min(*count_dev, Bmax) : Bmax;
    gr_stage<0, X3>(P, S0, B, recs, game, items_wh, nullptr, x1, feat_out, feat_lo_off, wts_gr, wts_gr_lo, prof, smem, sh, true);
    gr_stage<1, X3>(P, S1, B, recs, game, items_wh, x1, x2, feat_out, feat_lo_off, wts_gr, wts_gr_lo, prof, smem, sh, false);
    gr_stage<2, X3>(P, S2, B, recs, game, items_wh, x2, x3, feat_out, feat_lo_off, wts_gr, wts_gr_lo, prof, smem, sh, false);
    gr_stage<3, X3>(P, S3, B, recs, game, items_wh, x3, nullptr, feat_out, feat_lo_off, wts_gr, wts_gr_lo, prof, smem, sh, false);
    if (tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sh.tmem), "r"((uint32_t)S0.tmem_cols));
}

}  // namespace bppgr
