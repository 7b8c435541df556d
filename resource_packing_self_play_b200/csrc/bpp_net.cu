// bpp_net.cu — batched policy/value network forward for the leaf evaluation of the search
// (NNetWrapper.predict, binpacking/pytorch/NNet.py:69-85, over BinPackingNNet.forward,
// binpacking/pytorch/BinpackingNNet.py:72-81: three ConvSequences [conv3x3 -> maxpool(3,2,1) -> 2 residual blocks],
// flatten, relu, fc256, relu, {fc A -> log_softmax, fc 1 -> tanh}; predict returns exp(log_pi) and v).
//
// Numerics: weights and inter-layer activations are bf16, every accumulation is fp32 (the tolerance against the fp32
// torch reference is stated in tests/test_gpu_net.py).  The input planes are never materialised in HBM: they are
// synthesised in shared memory from the 128-byte compact state record (row masks + remaining mask) and the episode's
// item list.
//
// Kernel v1 (this file): one CTA per leaf, whole network fused in one launch, activations resident in shared memory,
// weights (341 KB bf16) served from L1/L2.  CUDA-core FMA; the tcgen05 implicit-GEMM version replaces the conv stages
// ("leaf evaluation").
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/bpp_b200.h"
#include "bpp_net_tc.cuh"

namespace {

constexpr int NCONV = 15;   // 3 sequences x (1 + 2*2) convolutions
constexpr int HIDDEN = 256;
constexpr int NET_THREADS = 128;
constexpr int CO_T = 4;     // output channels per thread

struct ConvDesc {
    int ci, co, h, w;       // input channels, output channels, spatial size
    long long w_off;        // offset (elements) into the bf16 weight buffer, layout [ci][tap][co]
    int b_off;              // offset into the fp32 bias buffer
};

struct NetParams {
    int W, H, N, Cin, A;
    int hs[4], ws[4];       // spatial size at the input of sequence s (s = 3: after the last pool)
    int flat;               // 32 * hs[3] * ws[3]
    int buf_elems;          // floats per activation buffer
    ConvDesc conv[NCONV];
    long long fc_hidden_off, fc_logits_off, fc_value_off;  // bf16, transposed [in][out]
    int b_hidden_off, b_logits_off, b_value_off;
    const __nv_bfloat16* wts;
    const float* wts32;     // same layout in fp32 (precision mode BPP_NET_FP32)
    const float* bias;
};

template <bool F32>
__device__ __forceinline__ float act_round(float x) {
    return F32 ? x : __bfloat162float(__float2bfloat16_rn(x));
}
template <bool F32>
__device__ __forceinline__ float wt(const NetParams& P, long long idx) {
    return F32 ? P.wts32[idx] : __bfloat162float(P.wts[idx]);
}

// out[co][y][x] = bias[co] + sum_{ci,dy,dx} act(in[ci][y+dy-1][x+dx-1]) * w[ci][dy*3+dx][co] (+ residual)
template <bool RELU_IN, bool F32>
__device__ void conv3x3(const NetParams& P, const ConvDesc& d, const float* __restrict__ in, float* __restrict__ out,
                        const float* residual) {
    const int hw = d.h * d.w;
    const int groups = d.co / CO_T;
    const __nv_bfloat16* wbase = P.wts + d.w_off;
    const float* wbase32 = P.wts32 + d.w_off;
    for (int idx = threadIdx.x; idx < groups * hw; idx += blockDim.x) {
        const int cg = idx / hw, p = idx - cg * hw;
        const int y = p / d.w, x = p - y * d.w;
        const int co0 = cg * CO_T;
        float acc[CO_T];
#pragma unroll
        for (int j = 0; j < CO_T; ++j) acc[j] = P.bias[d.b_off + co0 + j];
        for (int ci = 0; ci < d.ci; ++ci) {
            const float* ip = in + ci * hw;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int yy = y + dy - 1;
                if (yy < 0 || yy >= d.h) continue;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int xx = x + dx - 1;
                    if (xx < 0 || xx >= d.w) continue;
                    float v = ip[yy * d.w + xx];
                    if (RELU_IN) v = fmaxf(v, 0.f);
                    const size_t wi = (size_t)(ci * 9 + dy * 3 + dx) * d.co + co0;
                    if (F32) {
                        const float4 wv = *reinterpret_cast<const float4*>(wbase32 + wi);
                        acc[0] = fmaf(v, wv.x, acc[0]);
                        acc[1] = fmaf(v, wv.y, acc[1]);
                        acc[2] = fmaf(v, wv.z, acc[2]);
                        acc[3] = fmaf(v, wv.w, acc[3]);
                    } else {
                        const uint2 wv = *reinterpret_cast<const uint2*>(wbase + wi);
                        acc[0] = fmaf(v, __uint_as_float(wv.x << 16), acc[0]);
                        acc[1] = fmaf(v, __uint_as_float(wv.x & 0xffff0000u), acc[1]);
                        acc[2] = fmaf(v, __uint_as_float(wv.y << 16), acc[2]);
                        acc[3] = fmaf(v, __uint_as_float(wv.y & 0xffff0000u), acc[3]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < CO_T; ++j) {
            float r = acc[j];
            if (residual) r += residual[(co0 + j) * hw + p];
            out[(co0 + j) * hw + p] = act_round<F32>(r);
        }
    }
    __syncthreads();
}

// max_pool2d(kernel 3, stride 2, padding 1): (h, w) -> ((h+1)/2, (w+1)/2)
__device__ void maxpool3s2(const float* __restrict__ in, float* __restrict__ out, int c, int h, int w) {
    const int ho = (h + 1) / 2, wo = (w + 1) / 2;
    for (int idx = threadIdx.x; idx < c * ho * wo; idx += blockDim.x) {
        const int ch = idx / (ho * wo), p = idx - ch * ho * wo;
        const int oy = p / wo, ox = p - oy * wo;
        float m = -INFINITY;
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = 2 * oy + dy;
            if (yy < 0 || yy >= h) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = 2 * ox + dx;
                if (xx < 0 || xx >= w) continue;
                m = fmaxf(m, in[(ch * h + yy) * w + xx]);
            }
        }
        out[idx] = m;
    }
    __syncthreads();
}

template <bool F32>
__global__ void __launch_bounds__(NET_THREADS)
k_net_forward(NetParams P, int Bmax, const int32_t* __restrict__ count_dev, const uint32_t* __restrict__ recs,
              const int32_t* __restrict__ game, const int32_t* __restrict__ items_wh, float* __restrict__ policy,
              float* __restrict__ value) {
    extern __shared__ float smem[];
    float* bufA = smem;
    float* bufB = smem + P.buf_elems;
    float* bufC = smem + 2 * P.buf_elems;
    __shared__ float s_red[NET_THREADS / 32 + 1];
    const int B = count_dev ? min(*count_dev, Bmax) : Bmax;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        // ---- input planes from the compact record (getBinItem, BinPackingGame.py:118-120)
        const uint32_t* rec = recs + (size_t)b * 32;
        const int g = game ? game[b] : b;
        const int32_t* it = items_wh + (size_t)g * P.N * 2;
        const uint32_t rem = rec[BPP_REC_REM];
        const int hw0 = P.H * P.W;
        for (int idx = threadIdx.x; idx < P.Cin * hw0; idx += blockDim.x) {
            const int c = idx / hw0, p = idx - c * hw0;
            const int y = p / P.W, x = p - y * P.W;
            float v;
            if (c == 0) v = (float)((rec[y] >> x) & 1u);
            else v = (((rem >> (c - 1)) & 1u) && y < it[(c - 1) * 2 + 1] && x < it[(c - 1) * 2]) ? 1.f : 0.f;
            bufA[idx] = v;
        }
        __syncthreads();
        // ---- three ConvSequences (BinpackingNNet.py:29-48)
        float* x = bufA;
        float* t1 = bufB;
        float* t2 = bufC;
        int li = 0;
        for (int s = 0; s < 3; ++s) {
            const ConvDesc& c0 = P.conv[li++];
            conv3x3<false, F32>(P, c0, x, t1, nullptr);
            maxpool3s2(t1, x, c0.co, c0.h, c0.w);
            for (int blk = 0; blk < 2; ++blk) {  // ResidualBlock, BinpackingNNet.py:15-27
                const ConvDesc& ca = P.conv[li++];
                const ConvDesc& cb = P.conv[li++];
                conv3x3<true, F32>(P, ca, x, t1, nullptr);
                conv3x3<true, F32>(P, cb, t1, t2, x);
                float* sw = x; x = t2; t2 = sw;
            }
        }
        // ---- heads: flatten -> relu -> fc256 -> relu -> {logits, value} (BinpackingNNet.py:74-81)
        for (int o = threadIdx.x; o < HIDDEN; o += blockDim.x) {
            float acc = P.bias[P.b_hidden_off + o];
            for (int i = 0; i < P.flat; ++i)
                acc = fmaf(fmaxf(x[i], 0.f), wt<F32>(P, P.fc_hidden_off + (long long)i * HIDDEN + o), acc);
            t1[o] = act_round<F32>(fmaxf(acc, 0.f));
        }
        __syncthreads();
        float lmax = -INFINITY;
        for (int o = threadIdx.x; o < P.A; o += blockDim.x) {
            float acc = P.bias[P.b_logits_off + o];
            for (int i = 0; i < HIDDEN; ++i) acc = fmaf(t1[i], wt<F32>(P, P.fc_logits_off + (long long)i * P.A + o), acc);
            t2[o] = acc;
            lmax = fmaxf(lmax, acc);
        }
        if (threadIdx.x < 32) {  // value head: one warp
            float acc = 0.f;
            for (int i = threadIdx.x; i < HIDDEN; i += 32) acc = fmaf(t1[i], wt<F32>(P, P.fc_value_off + i), acc);
            for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (threadIdx.x == 0) value[b] = tanhf(acc + P.bias[P.b_value_off]);
        }
        // softmax = exp(log_softmax(logits)) (BinpackingNNet.py:81, NNet.py:85)
        for (int o = 16; o; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = lmax;
        __syncthreads();
        lmax = s_red[0];
        for (int i = 1; i < NET_THREADS / 32; ++i) lmax = fmaxf(lmax, s_red[i]);
        __syncthreads();
        float lsum = 0.f;
        for (int o = threadIdx.x; o < P.A; o += blockDim.x) lsum += expf(t2[o] - lmax);
        for (int o = 16; o; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = lsum;
        __syncthreads();
        lsum = 0.f;
        for (int i = 0; i < NET_THREADS / 32; ++i) lsum += s_red[i];
        const float lse = lmax + logf(lsum);
        for (int o = threadIdx.x; o < P.A; o += blockDim.x) policy[(size_t)b * P.A + o] = expf(t2[o] - lse);
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// tcgen05 forward: one CTA = S leaves through the whole network (see bpp_net_tc.cuh)
// NS = leaves per CTA the head accumulators are unrolled for (4 or 8); X3 = split-bf16 mode (hi + lo halves, 3 MMAs per
// product, fp32 heads) for networks whose dynamic range exceeds plain bf16
// TRUNK = the kernel ends at relu(flatten) (feat_out) and the in-kernel CUDA-core heads are compiled out; MINB = CTAs per
// SM the register allocation is bounded for
template <int NS, bool X3, bool TRUNK = false, int MINB = ((NS <= 4 && !X3) ? 2 : 1)>
__global__ void __launch_bounds__(bpptc::TC_THREADS, MINB)
k_net_forward_tc(NetParams P, bpptc::TcParams T, int Bmax, const int32_t* __restrict__ count_dev,
                 const uint32_t* __restrict__ recs, const int32_t* __restrict__ game,
                 const int32_t* __restrict__ items_wh, float* __restrict__ policy, float* __restrict__ value,
                 long long* prof, __nv_bfloat16* __restrict__ feat_out) {
    using namespace bpptc;
    extern __shared__ __align__(1024) unsigned char arena[];
    __shared__ __align__(8) uint64_t s_bar[MAX_BARS];
    __shared__ uint32_t s_rec[NS][32];
    __shared__ int s_it[NS][BPP_MAX_ITEMS][2];
    __shared__ uint32_t s_tmem;
    __shared__ long long s_prof[8];
    unsigned char* regA = arena;
    unsigned char* regB = arena + T.regA_bytes;
    unsigned char* wbuf = regB + T.regB_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < MAX_BARS) mbar_init(smem_u32(&s_bar[tid]), 1);
    if (tid < 8) s_prof[tid] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"((uint32_t)T.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    Ctx cx;
    cx.tmem = s_tmem;
    cx.bar = smem_u32(&s_bar[0]);
    cx.phase = 0;
    cx.wbuf = wbuf;
    cx.prof = s_prof;
    const long long t_start = clock64();
    const int S = T.S;
    const int cin16_0 = (P.Cin + 15) / 16;
    WPre pre;
    auto lay_w = [&](int l) { return T.wts_umma + T.w_off[l]; };
    auto lay_b = [&](int l) { return P.bias + P.conv[l].b_off; };
    auto lay_wl = [&](int l) { return T.wts_umma_lo + T.w_off[l]; };
    if (!X3) wpre_load(pre, lay_w(0), T.lay_n16[0], lay_b(0), P.conv[0].co);
    // everything above (barriers, TMEM, first weights) is independent of the kernel that produced the leaf batch: with a
    // programmatic launch it overlaps that kernel's tail; the heads kernel behind us may start launching right away
    pdl_launch_dependents();
    pdl_wait();
    const int B = count_dev ? min(*count_dev, Bmax) : Bmax;
    // every CTA takes one contiguous, equally sized slice of the batch and walks it in groups of <= S leaves: all CTAs
    // finish together (with "group g -> CTA g mod grid" a batch of 2,800 leaves left 36 % of the CTAs idle in the last wave)
    const int slice_lo = (int)(((long long)blockIdx.x * B) / gridDim.x);
    const int slice_hi = (int)(((long long)(blockIdx.x + 1) * B) / gridDim.x);
    // ... in equally sized groups (a 9-leaf slice at S = 5 is 5 + 4, a 6-leaf slice 3 + 3): a group's time is dominated
    // by the latency chain of its 15 layers, not by its size
    const int n_slice = slice_hi - slice_lo;
    const int n_groups = (n_slice + S - 1) / S;
    const int gsz = n_groups > 0 ? (n_slice + n_groups - 1) / n_groups : S;
    for (int b0 = slice_lo; b0 < slice_hi; b0 += gsz) {
        const int nvalid = min(gsz, slice_hi - b0);
        long long tq = clock64();
        // ---- level-0 operand planes from the compact records (getBinItem, BinPackingGame.py:118-120)
        const Level& L0 = T.lv[0];
        zero_bytes(regA, (X3 ? 2 : 1) * 2 * cin16_0 * L0.RT * 16);  // split mode: the lo planes of the 0/1 input stay zero
        for (int i = tid; i < nvalid * 32; i += TC_THREADS) s_rec[i >> 5][i & 31] = recs[(size_t)(b0 + (i >> 5)) * 32 + (i & 31)];
        for (int i = tid; i < nvalid * P.N * 2; i += TC_THREADS) {
            const int j = i / (P.N * 2), r = i - j * P.N * 2;
            const int b = b0 + j;
            const int g = game ? game[b] : b;
            s_it[j][r >> 1][r & 1] = items_wh[(size_t)g * P.N * 2 + r];
        }
        __syncthreads();
        {
            // one thread builds half a grid row of one 8-channel plane: the 8 channels' column masks of that row once, then
            // one 16-byte pixel per step.  The lanes of a warp walk their rows from different starting columns, so that
            // their simultaneous 16-byte stores fall into different bank groups (rows are wp * 16 bytes apart).
            const int nplanes = 2 * cin16_0, rows = nvalid * P.H;
            const int xh = (P.W + 1) >> 1;
            const uint32_t mrows = fdiv_magic((uint32_t)rows);
            for (int idx = tid; idx < 2 * nplanes * rows; idx += TC_THREADS) {
                const int hp = fdiv(idx, mrows);
                int r = idx - hp * rows;
                const int j = fdiv(r, T.mH), y = r - j * P.H;
                const int half = hp >= nplanes ? 1 : 0, p = hp - half * nplanes;
                const uint32_t rem = s_rec[j][BPP_REC_REM];
                uint32_t m[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int c = p * 8 + k;
                    m[k] = 0;
                    if (c == 0) m[k] = s_rec[j][y];
                    else if (c <= P.N && ((rem >> (c - 1)) & 1u) && y < s_it[j][c - 1][1]) {
                        const int iw = s_it[j][c - 1][0];
                        m[k] = iw >= 32 ? 0xffffffffu : (1u << iw) - 1u;
                    }
                }
                const int x0 = half ? xh : 0, n = half ? P.W - xh : xh;
                uint4* dst = reinterpret_cast<uint4*>(regA) + (size_t)p * L0.RT + L0.guard + j * L0.P + (y + 1) * L0.wp + 1 + x0;
                int i = n > 0 ? lane % n : 0;
                for (int step = 0; step < n; ++step) {
                    const int x = x0 + i;
                    uint32_t w[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        w[q] = (((m[2 * q] >> x) & 1u) | (((m[2 * q + 1] >> x) & 1u) << 16)) * 0x3f80u;  // bf16 1.0 pairs
                    dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
                    if (++i == n) i = 0;
                }
            }
        }
        __syncthreads();
        TC_PROF(0, tq);
        // ---- three ConvSequences
        // Arena use.  Classic: region A = level-0 input, then the {raw, actA, actB} triples; region B = conv outputs T
        // awaiting pooling; weights behind both.  Compact (bf16 trunk, two CTAs per SM): the triples live in region B;
        // T_0 ALIASES the input planes, shifted down by `guard` rows - tile t's epilogue writes physical rows
        // [128t, 128t+128) while the MMAs still queued (tiles > t, issued and completed in order) read input rows
        // >= 128(t+1) - wp - 1, i.e. physical rows >= 128(t+1) because guard >= wp + 1; T_1, T_2 and every layer's
        // weights sit in whatever region is idle at that layer.  This halves the level-0 footprint: larger groups.
        const bool compact = !X3 && T.compact;
        unsigned char* const tri = compact ? regB : regA;
        const unsigned char* in = regA;
        int cin16 = cin16_0, li = 0;
        unsigned char* raw = tri;
        for (int s = 0; s < 3; ++s) {
            const Level& La = T.lv[s];
            const Level& Lb = T.lv[s + 1];
            const int cout = P.conv[li].co;
            const uint32_t in_lo = (uint32_t)(2 * cin16) * (uint32_t)La.RT * 16u;   // bytes of the input's hi planes
            const uint32_t t_lo = (uint32_t)(cout / 8) * (uint32_t)La.RT * 16u;     // bytes of T's hi planes
            unsigned char* Tbuf = regB;
            if (compact) {
                Tbuf = s == 0 ? regA - (size_t)La.guard * 16 : regA;
                cx.wbuf = s == 0 ? regB : regA + ((t_lo + 127u) & ~127u);
            }
            conv_layer<X3>(T, cx, La, nvalid, cin16, cout, lay_w(li), lay_b(li), in, EPI_CONV, Tbuf, nullptr, pre,
                           lay_w(li + 1), T.lay_n16[li + 1], lay_b(li + 1), P.conv[li + 1].co, lay_wl(li), in_lo, t_lo, 0u);
            ++li;
            tq = clock64();
            const int planes = cout / 8;
            const size_t pb = (size_t)planes * Lb.RT * 16;   // one logical buffer (hi planes)
            const size_t bs = X3 ? 2 * pb : pb;               // stride between logical buffers (hi [+ lo])
            raw = tri;
            unsigned char* actA = tri + bs;
            unsigned char* actB = tri + 2 * bs;
            zero_bytes(tri, (int)(3 * bs));
            if (!X3) pool_pad_tail(La, nvalid, planes, Tbuf);
            __syncthreads();
            if (X3) pool_level_x3(La, Lb, nvalid, planes, Tbuf, t_lo, raw, (uint32_t)pb, actA, (uint32_t)pb);
            else pool_level(La, Lb, nvalid, planes, Tbuf, raw, actA);
            __syncthreads();
            TC_PROF(5, tq);
            if (compact) cx.wbuf = regA;  // T is dead: the residual layers' weights go to region A
            for (int blk = 0; blk < 2; ++blk) {
                conv_layer<X3>(T, cx, Lb, nvalid, cout / 16, cout, lay_w(li), lay_b(li), actA, EPI_RES0, actB, nullptr, pre,
                               lay_w(li + 1), T.lay_n16[li + 1], lay_b(li + 1), P.conv[li + 1].co, lay_wl(li),
                               (uint32_t)pb, (uint32_t)pb, (uint32_t)pb);
                ++li;
                {
                    const int nx = (li + 1) % NCONV;  // after the last layer: layer 0 of this CTA's next group
                    conv_layer<X3>(T, cx, Lb, nvalid, cout / 16, cout, lay_w(li), lay_b(li), actB, EPI_RES1, actA, raw, pre,
                                   lay_w(nx), T.lay_n16[nx], lay_b(nx), P.conv[nx].co, lay_wl(li), (uint32_t)pb,
                                   (uint32_t)pb, (uint32_t)pb);
                }
                ++li;
            }
            in = raw;
            cin16 = cout / 16;
        }
        tq = clock64();
        // ---- heads on the CUDA cores (1.6 % of the FLOPs): flatten -> relu -> fc256 -> relu -> {logits, value}
        const Level& L3 = T.lv[3];
        float* feat = reinterpret_cast<float*>(regB);            // [NS][flat]
        float* hid = feat + NS * P.flat;                          // [NS][256]
        float* lg = hid + NS * HIDDEN;                            // [NS][A]
        float* part = lg + NS * P.A;                              // [2][NS][256] / [2][NS][A] partial sums
        const int hw3 = L3.h * L3.w;
        const size_t raw_lo3 = (size_t)(P.conv[NCONV - 1].co / 8) * L3.RT * 16;
        if (TRUNK || (!X3 && feat_out)) {
            // trunk only: relu(flatten(x)) goes to HBM as bf16 [B][flat]; the two FC heads run as real GEMMs over the
            // whole batch in k_net_heads_tc (128-leaf M tiles) instead of per group of S leaves here
            const uint32_t mflat = fdiv_magic((uint32_t)P.flat);
            for (int idx = tid; idx < nvalid * P.flat; idx += TC_THREADS) {
                const int j = fdiv(idx, mflat), f = idx - j * P.flat;
                const int c = fdiv(f, L3.mhw), q = f - c * hw3;
                const int y = fdiv(q, L3.mw), x = q - y * L3.w;
                const size_t row = (size_t)L3.guard + (size_t)j * L3.P + (size_t)(y + 1) * L3.wp + (x + 1);
                const uint16_t e = *(reinterpret_cast<const uint16_t*>(raw + ((size_t)(c >> 3) * L3.RT + row) * 16) + (c & 7));
                reinterpret_cast<uint16_t*>(feat_out)[(size_t)(b0 + j) * P.flat + f] = (e & 0x8000u) ? (uint16_t)0 : e;
            }
            __syncthreads();
            TC_PROF(6, tq);
            continue;
        }
        if (TRUNK) continue;
        for (int idx = tid; idx < NS * P.flat; idx += TC_THREADS) {
            const int j = idx / P.flat, f = idx - j * P.flat;
            float v = 0.f;
            if (j < nvalid) {
                const int c = f / hw3, q = f - c * hw3;
                const int y = q / L3.w, x = q - y * L3.w;
                const size_t row = (size_t)L3.guard + (size_t)j * L3.P + (size_t)(y + 1) * L3.wp + (x + 1);
                const size_t eoff = ((size_t)(c >> 3) * L3.RT + row) * 16;
                const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(raw + eoff) + (c & 7);
                v = __bfloat162float(*e);
                if (X3) v += __bfloat162float(*(reinterpret_cast<const __nv_bfloat16*>(raw + raw_lo3 + eoff) + (c & 7)));
                v = fmaxf(v, 0.f);
            }
            feat[idx] = v;
        }
        __syncthreads();
        const int hh = tid >> 7, ot = tid & 127;  // input half, output pair
        if (tid < TC_WORKERS) {   // hidden layer: thread -> outputs 2*ot, 2*ot+1 over one half of the inputs (coalesced bf16x2 loads)
            const int o = 2 * ot;
            float a0[NS], a1[NS];
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                a0[j] = hh ? 0.f : P.bias[P.b_hidden_off + o];
                a1[j] = hh ? 0.f : P.bias[P.b_hidden_off + o + 1];
            }
            const uint32_t* wp32 = reinterpret_cast<const uint32_t*>(P.wts + P.fc_hidden_off) + ot;
            const int i0 = hh * (P.flat / 2), i1 = hh ? P.flat : P.flat / 2;
#pragma unroll 16
            for (int i = i0; i < i1; ++i) {
                float w0, w1;
                if (X3) {
                    const float2 wf = __ldg(reinterpret_cast<const float2*>(P.wts32 + P.fc_hidden_off + (size_t)i * HIDDEN) + ot);
                    w0 = wf.x; w1 = wf.y;
                } else {
                    const uint32_t wv = __ldg(wp32 + (size_t)i * (HIDDEN / 2));
                    w0 = bf16_lo(wv); w1 = bf16_hi(wv);
                }
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    const float f = feat[j * P.flat + i];
                    a0[j] = fmaf(f, w0, a0[j]);
                    a1[j] = fmaf(f, w1, a1[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                part[(hh * NS + j) * HIDDEN + o] = a0[j];
                part[(hh * NS + j) * HIDDEN + o + 1] = a1[j];
            }
        }
        __syncthreads();
        for (int idx = tid; idx < NS * HIDDEN; idx += TC_THREADS)
            hid[idx] = act_round<X3>(fmaxf(part[idx] + part[NS * HIDDEN + idx], 0.f));
        __syncthreads();
        if (tid < TC_WORKERS && 2 * ot < T.A_pad) {  // logits: padded [256][A_pad] copy of the weights, outputs 2*ot, 2*ot+1
            const int o = 2 * ot;
            float a0[NS], a1[NS];
            const float b0 = hh ? 0.f : P.bias[P.b_logits_off + o];
            const float b1 = (hh || o + 1 >= P.A) ? 0.f : P.bias[P.b_logits_off + o + 1];
#pragma unroll
            for (int j = 0; j < NS; ++j) { a0[j] = b0; a1[j] = b1; }
            const uint32_t* wp32 = reinterpret_cast<const uint32_t*>(T.wts_logits_pad) + ot;
            const int i0 = hh * (HIDDEN / 2), i1 = i0 + HIDDEN / 2;
#pragma unroll 16
            for (int i = i0; i < i1; ++i) {
                float w0, w1;
                if (X3) {
                    const float* wr = P.wts32 + P.fc_logits_off + (size_t)i * P.A + o;
                    w0 = __ldg(wr);
                    w1 = o + 1 < P.A ? __ldg(wr + 1) : 0.f;
                } else {
                    const uint32_t wv = __ldg(wp32 + (size_t)i * (T.A_pad / 2));
                    w0 = bf16_lo(wv); w1 = bf16_hi(wv);
                }
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    const float f = hid[j * HIDDEN + i];
                    a0[j] = fmaf(f, w0, a0[j]);
                    a1[j] = fmaf(f, w1, a1[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                part[(hh * NS + j) * T.A_pad + o] = a0[j];
                part[(hh * NS + j) * T.A_pad + o + 1] = a1[j];
            }
        }
        __syncthreads();
        for (int idx = tid; idx < NS * P.A; idx += TC_THREADS) {
            const int j = idx / P.A, o = idx - j * P.A;
            lg[idx] = part[j * T.A_pad + o] + part[(NS + j) * T.A_pad + o];
        }
        __syncthreads();
        for (int j = warp; j < nvalid; j += TC_THREADS / 32) {  // one warp per leaf: value head + softmax
            const int b = b0 + j;
            float acc = 0.f;
            for (int i = lane; i < HIDDEN; i += 32)
                acc = fmaf(hid[j * HIDDEN + i], wt<X3>(P, P.fc_value_off + i), acc);
            float mx = -INFINITY;
            for (int o = lane; o < P.A; o += 32) mx = fmaxf(mx, lg[j * P.A + o]);
            for (int o = 16; o; o >>= 1) {
                acc += __shfl_xor_sync(0xffffffffu, acc, o);
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
            float sum = 0.f;
            for (int o = lane; o < P.A; o += 32) sum += expf(lg[j * P.A + o] - mx);
            for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float lse = mx + logf(sum);
            for (int o = lane; o < P.A; o += 32) policy[(size_t)b * P.A + o] = expf(lg[j * P.A + o] - lse);
            if (lane == 0) value[b] = tanhf(acc + P.bias[P.b_value_off]);
        }
        __syncthreads();
        TC_PROF(6, tq);
    }
    if (prof && blockIdx.x == 0 && tid == 0) {
        cx.prof[7] = clock64() - t_start;
        for (int i = 0; i < 8; ++i) prof[i] = cx.prof[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(cx.tmem), "r"((uint32_t)T.tmem_cols));
}


// ---------------------------------------------------------------------------------------------------------------------
// Role kernels: the trunk split at the ConvSequence boundaries.  k_net_role<SEQ> runs sequence SEQ (conv3x3 -> max-pool ->
// two residual blocks, BinpackingNNet.py:29-48) for the WHOLE batch and hands its residual stream to the next role
// through HBM (8x8x16 / 4x4x32 bf16 per leaf at 15x15: 2 KB / 1 KB), so that every role picks its own group size: the
// deep levels are tiny (25 and 9 rows per leaf), and with the one-kernel trunk's 5 leaves per group nine of its fifteen
// layers ran a single, mostly empty 128-row tile behind a full per-layer latency chain (weights -> barrier -> MMAs ->
// epilogue -> barrier).  Here role 1 takes ~8 and role 2 ~25 leaves per group: full tiles, the chain amortised 2-5x.
// Hand-over layout: [leaf][hi|lo (X3 only)][8-channel plane][pixel] 16-byte units, interior pixels only.
template <int SEQ, bool X3, int MINB>
__global__ void __launch_bounds__(bpptc::TC_THREADS, MINB)
k_net_role(NetParams P, bpptc::TcParams T, int Bmax, const int32_t* __restrict__ count_dev,
           const uint32_t* __restrict__ recs, const int32_t* __restrict__ game, const int32_t* __restrict__ items_wh,
           const uint4* __restrict__ xin, uint4* __restrict__ xout, __nv_bfloat16* __restrict__ feat_out,
           long long feat_lo_off, long long* prof) {
    using namespace bpptc;
    constexpr int NSR = SEQ == 0 ? 8 : 1;
    extern __shared__ __align__(1024) unsigned char arena[];
    __shared__ __align__(8) uint64_t s_bar[MAX_BARS];
    __shared__ uint32_t s_rec[NSR][32];
    __shared__ int s_it[NSR][BPP_MAX_ITEMS][2];
    __shared__ uint32_t s_tmem;
    __shared__ long long s_prof[8];
    unsigned char* regA = arena;
    unsigned char* regB = arena + T.regA_bytes;
    unsigned char* wbuf = regB + T.regB_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < MAX_BARS) mbar_init(smem_u32(&s_bar[tid]), 1);
    if (tid < 8) s_prof[tid] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"((uint32_t)T.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    Ctx cx;
    cx.tmem = s_tmem;
    cx.bar = smem_u32(&s_bar[0]);
    cx.phase = 0;
    cx.wbuf = wbuf;
    cx.prof = s_prof;
    const long long t_start = clock64();
    const int S = T.S;
    constexpr int l0 = 5 * SEQ;  // this role's first conv layer
    const int cin16 = (P.conv[l0].ci + 15) / 16;
    const int cout = P.conv[l0].co;
    const int fx = X3 ? 2 : 1;
    WPre pre;
    auto lay_w = [&](int l) { return T.wts_umma + T.w_off[l]; };
    auto lay_b = [&](int l) { return P.bias + P.conv[l].b_off; };
    auto lay_wl = [&](int l) { return T.wts_umma_lo + T.w_off[l]; };
    if (!X3) wpre_load(pre, lay_w(l0), T.lay_n16[l0], lay_b(l0), cout);
    pdl_launch_dependents();
    pdl_wait();
    const int B = count_dev ? min(*count_dev, Bmax) : Bmax;
    const int slice_lo = (int)(((long long)blockIdx.x * B) / gridDim.x);
    const int slice_hi = (int)(((long long)(blockIdx.x + 1) * B) / gridDim.x);
    const int n_slice = slice_hi - slice_lo;
    const int n_groups = (n_slice + S - 1) / S;
    const int gsz = n_groups > 0 ? (n_slice + n_groups - 1) / n_groups : S;
    const Level& La = T.lv[SEQ];
    const Level& Lb = T.lv[SEQ + 1];
    const int in_planes_n = 2 * cin16;
    const int hwa = La.h * La.w, hwb = Lb.h * Lb.w;
    const int planes = cout / 8;
    for (int b0 = slice_lo; b0 < slice_hi; b0 += gsz) {
        const int nvalid = min(gsz, slice_hi - b0);
        long long tq = clock64();
        // ---- this role's input in the level-SEQ operand planes (region A).  One thread per grid entry (16 bytes = the 8
        // channels of one padded pixel), consecutive threads on consecutive entries: halo entries are written as zeros in
        // the same pass (no separate clearing of the region), up to one grid row behind the last present sample - the
        // furthest entry a stored output row reads.  Entries beyond that keep stale data: they only feed accumulator rows
        // that are never stored.
        if (SEQ == 0) {
            for (int i = tid; i < nvalid * 32; i += TC_THREADS) s_rec[i >> 5][i & 31] = recs[(size_t)(b0 + (i >> 5)) * 32 + (i & 31)];
            for (int i = tid; i < nvalid * P.N * 2; i += TC_THREADS) {
                const int j = i / (P.N * 2), r = i - j * P.N * 2;
                const int b = b0 + j;
                const int g = game ? game[b] : b;
                s_it[j][r >> 1][r & 1] = items_wh[(size_t)g * P.N * 2 + r];
            }
            __syncthreads();
        }
        {
            const int nrows = nvalid * La.P + La.wp + 1;
            const int nplanes = (SEQ == 0 ? 1 : fx) * in_planes_n;   // the 0/1 input has no low halves
            const uint32_t mrows = fdiv_magic((uint32_t)nrows);
            uint4* base = reinterpret_cast<uint4*>(regA);
            for (int idx = tid; idx < nplanes * La.guard; idx += TC_THREADS) {   // front guard rows
                const int p = idx / La.guard;
                base[(size_t)p * La.RT + (idx - p * La.guard)] = make_uint4(0, 0, 0, 0);
            }
            const int per = nplanes * hwa;   // hand-over entries per leaf
            for (int idx = tid; idx < nplanes * nrows; idx += TC_THREADS) {
                const int p = fdiv(idx, mrows);
                const int r = idx - p * nrows;
                const int j = fdiv(r, La.mP), q = r - j * La.P;
                const int yp = fdiv(q, La.mwp), xp = q - yp * La.wp;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (j < nvalid && yp >= 1 && xp >= 1) {
                    const int y = yp - 1, x = xp - 1;
                    if (SEQ == 0) {
                        // input planes from the compact record (getBinItem, BinPackingGame.py:118-120): channel 0 = bin
                        // occupancy, channel i+1 = item i's [0:h, 0:w] block while it is still to be placed
                        const uint32_t rem = s_rec[j][BPP_REC_REM];
                        uint32_t bits = 0;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int c = p * 8 + k;
                            uint32_t on;
                            if (c == 0) on = (s_rec[j][y] >> x) & 1u;
                            else on = (c <= P.N && ((rem >> (c - 1)) & 1u) && y < s_it[j][c - 1][1] && x < s_it[j][c - 1][0]) ? 1u : 0u;
                            bits |= on << k;
                        }
                        v.x = ((bits & 1u) | ((bits & 2u) << 15)) * 0x3f80u;          // bf16 1.0 pairs
                        v.y = (((bits >> 2) & 1u) | ((bits & 8u) << 13)) * 0x3f80u;
                        v.z = (((bits >> 4) & 1u) | ((bits & 32u) << 11)) * 0x3f80u;
                        v.w = (((bits >> 6) & 1u) | ((bits & 128u) << 9)) * 0x3f80u;
                    } else {
                        // the previous role's residual stream: [leaf][hi|lo][plane][pixel]
                        v = __ldg(xin + (size_t)(b0 + j) * per + (size_t)p * hwa + y * La.w + x);
                    }
                }
                base[(size_t)p * La.RT + La.guard + r] = v;
            }
        }
        __syncthreads();
        TC_PROF(0, tq);
        int li = l0;
        const uint32_t in_lo = (uint32_t)in_planes_n * (uint32_t)La.RT * 16u;   // bytes of the input's hi planes
        const uint32_t t_lo = (uint32_t)planes * (uint32_t)La.RT * 16u;         // bytes of T's hi planes
        unsigned char* Tbuf = regB;
        conv_layer<X3>(T, cx, La, nvalid, cin16, cout, lay_w(li), lay_b(li), regA, EPI_CONV, Tbuf, nullptr, pre,
                       lay_w(li + 1), T.lay_n16[li + 1], lay_b(li + 1), P.conv[li + 1].co, lay_wl(li), in_lo, t_lo, 0u,
                       SEQ == 0);   // the 0/1 input planes are exact in bf16: no a_lo term in the network's first layer
        ++li;
        tq = clock64();
        const size_t pb = (size_t)planes * Lb.RT * 16;   // one logical buffer (hi planes)
        const size_t bs = X3 ? 2 * pb : pb;               // stride between logical buffers (hi [+ lo])
        unsigned char* raw = regA;
        unsigned char* actA = regA + bs;
        unsigned char* actB = regA + 2 * bs;
        zero_bytes(regA, (int)(3 * bs));
        if (!X3) pool_pad_tail(La, nvalid, planes, Tbuf);
        __syncthreads();
        if (X3) pool_level_x3(La, Lb, nvalid, planes, Tbuf, t_lo, raw, (uint32_t)pb, actA, (uint32_t)pb);
        else pool_level(La, Lb, nvalid, planes, Tbuf, raw, actA);
        __syncthreads();
        TC_PROF(5, tq);
        for (int blk = 0; blk < 2; ++blk) {
            conv_layer<X3>(T, cx, Lb, nvalid, cout / 16, cout, lay_w(li), lay_b(li), actA, EPI_RES0, actB, nullptr, pre,
                           lay_w(li + 1), T.lay_n16[li + 1], lay_b(li + 1), P.conv[li + 1].co, lay_wl(li), (uint32_t)pb,
                           (uint32_t)pb, (uint32_t)pb);
            ++li;
            const int nx = li + 1 == l0 + 5 ? l0 : li + 1;   // after the role's last layer: its first layer (next group)
            conv_layer<X3>(T, cx, Lb, nvalid, cout / 16, cout, lay_w(li), lay_b(li), actB, EPI_RES1, actA, raw, pre,
                           lay_w(nx), T.lay_n16[nx], lay_b(nx), P.conv[nx].co, lay_wl(li), (uint32_t)pb, (uint32_t)pb,
                           (uint32_t)pb);
            ++li;
        }
        tq = clock64();
        if (SEQ < 2) {
            // hand the residual stream (interior pixels) to the next role
            const int per = fx * planes * hwb;
            const uint32_t mper = fdiv_magic((uint32_t)per);
            const uint32_t mhwb = fdiv_magic((uint32_t)hwb);
            uint4* dst = xout + (size_t)b0 * per;
            for (int idx = tid; idx < nvalid * per; idx += TC_THREADS) {
                const int j = fdiv(idx, mper);
                int r = idx - j * per;
                const int hp = fdiv(r, mhwb), q = r - hp * hwb;
                const int y = fdiv(q, Lb.mw), x = q - y * Lb.w;
                const size_t row = (size_t)Lb.guard + (size_t)j * Lb.P + (size_t)(y + 1) * Lb.wp + (x + 1);
                dst[idx] = reinterpret_cast<const uint4*>(raw)[(size_t)hp * Lb.RT + row];
            }
        } else {
            // relu(flatten(x)) as bf16 [B][flat] (hi, and lo `flat_lo_off` elements behind it in X3 mode) for the FC heads
            const uint32_t mflat = fdiv_magic((uint32_t)P.flat);
            for (int idx = tid; idx < nvalid * P.flat; idx += TC_THREADS) {
                const int j = fdiv(idx, mflat), f = idx - j * P.flat;
                const int c = fdiv(f, Lb.mhw), q = f - c * hwb;
                const int y = fdiv(q, Lb.mw), x = q - y * Lb.w;
                const size_t row = (size_t)Lb.guard + (size_t)j * Lb.P + (size_t)(y + 1) * Lb.wp + (x + 1);
                const size_t eoff = ((size_t)(c >> 3) * Lb.RT + row) * 16;
                const uint16_t e = *(reinterpret_cast<const uint16_t*>(raw + eoff) + (c & 7));
                if (!X3) {
                    reinterpret_cast<uint16_t*>(feat_out)[(size_t)(b0 + j) * P.flat + f] = (e & 0x8000u) ? (uint16_t)0 : e;
                } else {
                    const uint16_t el = *(reinterpret_cast<const uint16_t*>(raw + pb + eoff) + (c & 7));
                    const float v = fmaxf(__uint_as_float((uint32_t)e << 16) + __uint_as_float((uint32_t)el << 16), 0.f);
                    const uint32_t hb = pack_bf16(v, 0.f) & 0xffffu;
                    const uint32_t lb = pack_bf16(v - __uint_as_float(hb << 16), 0.f) & 0xffffu;
                    reinterpret_cast<uint16_t*>(feat_out)[(size_t)(b0 + j) * P.flat + f] = (uint16_t)hb;
                    reinterpret_cast<uint16_t*>(feat_out)[(size_t)feat_lo_off + (size_t)(b0 + j) * P.flat + f] = (uint16_t)lb;
                }
            }
        }
        __syncthreads();
        TC_PROF(6, tq);
    }
    if (prof && blockIdx.x == 0 && tid == 0) {
        cx.prof[7] = clock64() - t_start;
        for (int i = 0; i < 8; ++i) prof[8 * SEQ + i] = cx.prof[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(cx.tmem), "r"((uint32_t)T.tmem_cols));
}

#include "bpp_net_gr.cuh"

// ---------------------------------------------------------------------------------------------------------------------
// FC heads as batched GEMMs on the tensor core: one CTA = 128 leaves (one M tile).
//   hidden = relu(feat[128 x flat] * W1^T + b1)   -> TMEM columns [0, 256)
//   logits = hidden[128 x 256] * W2^T + b2        -> TMEM columns [256, 256 + N2)
//   value  = tanh(hidden . wv + bv);  policy = softmax(logits)   (thread t = leaf t = TMEM lane t)
// The A operand (features, then the bf16 hidden activations) sits in shared memory in the 8-channel-plane layout; the
// weights stream through two 32 KB stages in the UMMA B layout prepared on the host.
struct HeadParams {
    int flat, A, N2;                      // N2 = A rounded up to 16 (UMMA N granularity for M = 128)
    const __nv_bfloat16* w1u;             // [flat/16][2][256][8]
    const __nv_bfloat16* w2u;             // [16][2][N2][8]
    const float* b1;                      // [256]
    const float* b2;                      // [A]
    const float* wv;                      // [256] (bf16-rounded values, as fp32) + [256] = the value head's bias
    // split-bf16 (X3) mode: low halves of the FC weights in the same layouts, the exact fp32 value weights (+ bias), and
    // the offset (elements) of the low halves of the features behind the high ones
    const __nv_bfloat16* w1u_lo;
    const __nv_bfloat16* w2u_lo;
    const float* wv_f32;
    long long feat_lo_off;
};
#ifdef BPP_HEADS_PROF  // phase timers of the heads kernel (debug builds: nvcc -DBPP_HEADS_PROF), printed by CTA 0
#define HP_T(i) t_[i] = clock64()
#else
#define HP_T(i)
#endif
constexpr int HEAD_THREADS = 512;           // warps w, w + 4, w + 8, w + 12 share a TMEM lane quarter (rows) and split the columns
constexpr int HEAD_NQ = HEAD_THREADS / 128;  // column parts per row
constexpr int HEAD_STAGES = 4;               // ring of weight stages filled by cp.async, drained by the MMAs
constexpr int HEAD_STAGE_BYTES = 16 * 1024;

__device__ __forceinline__ void head_cp16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(bpptc::smem_u32(dst_smem)), "l"(src));
}

// X3 = split-bf16: features, hidden activations and both weight matrices are hi + lo bf16 pairs and every product is the
// three MMAs hi*hi + lo*hi + hi*lo (the weight chunks stream twice: high halves, then low halves)
template <bool X3>
__global__ void __launch_bounds__(HEAD_THREADS, 1)
k_net_heads_tc(HeadParams Hp, int Bmax, const int32_t* __restrict__ count_dev, const __nv_bfloat16* __restrict__ feat,
               float* __restrict__ policy, float* __restrict__ value) {
    using namespace bpptc;
    extern __shared__ __align__(1024) unsigned char hsm[];
    __shared__ __align__(8) uint64_t s_bar[HEAD_STAGES + 1];  // [s]: MMAs that read stage s are done; [last]: GEMM done
    __shared__ uint32_t s_tmem;
    __shared__ float s_b1[HIDDEN], s_wv[HIDDEN], s_b2[256], s_x0[HEAD_NQ][128], s_x1[HEAD_NQ][128];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row_l = tid & 127, hv = tid >> 7;  // row of the tile, column part (0 .. HEAD_NQ-1)
#ifdef BPP_HEADS_PROF
    long long t_[8];
#endif
    HP_T(0);
    const int row0 = blockIdx.x * 128;
    const int kplanes = (Hp.flat > HIDDEN ? Hp.flat : HIDDEN) / 8;
    unsigned char* areg = hsm;                                   // planes x 128 rows x 16 B (X3: the lo planes behind)
    const uint32_t alo_off = (uint32_t)kplanes * 2048u;
    unsigned char* stage0 = hsm + (size_t)(X3 ? 2 : 1) * kplanes * 2048;
    // the weight chunks of both GEMMs as one list: chunk c covers k-blocks [kb, kb + nk) of GEMM g
    const int blk1 = 2 * HIDDEN * 16, blk2 = 2 * Hp.N2 * 16;     // bytes per 16-deep k-block of W1 / W2 (UMMA B layout)
    const int kpc1 = HEAD_STAGE_BYTES / blk1, kpc2 = HEAD_STAGE_BYTES / blk2;
    const int nkb1 = Hp.flat / 16, nkb2 = HIDDEN / 16;
    const int nch1 = (nkb1 + kpc1 - 1) / kpc1, nch2 = (nkb2 + kpc2 - 1) / kpc2;
    const int n1 = (X3 ? 2 : 1) * nch1, n2 = (X3 ? 2 : 1) * nch2;
    const int nch = n1 + n2;
    // chunk c -> (GEMM 2?, low-half pass?, index inside the pass)
    auto chunk_info = [&](int c, bool& g2, bool& lo, int& cc) {
        g2 = c >= n1;
        const int c2 = g2 ? c - n1 : c, np = g2 ? nch2 : nch1;
        lo = c2 >= np;
        cc = lo ? c2 - np : c2;
    };
    auto issue_chunk = [&](int c) {  // cp.async this thread's share of chunk c into stage c % HEAD_STAGES
        if (c < nch) {
            bool g2, lo;
            int cc;
            chunk_info(c, g2, lo, cc);
            const int kpc = g2 ? kpc2 : kpc1, blk = g2 ? blk2 : blk1, nkb = g2 ? nkb2 : nkb1;
            const int kb = cc * kpc, nk = min(kpc, nkb - kb);
            const __nv_bfloat16* wsrc = g2 ? (lo ? Hp.w2u_lo : Hp.w2u) : (lo ? Hp.w1u_lo : Hp.w1u);
            const unsigned char* src = reinterpret_cast<const unsigned char*>(wsrc) + (size_t)kb * blk;
            unsigned char* dst = stage0 + (size_t)(c % HEAD_STAGES) * HEAD_STAGE_BYTES;
            for (int i = tid * 16; i < nk * blk; i += HEAD_THREADS * 16) head_cp16(dst + i, src + i);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");  // one group per chunk slot, possibly empty
    };
    // Prologue that does not depend on the trunk kernel (with a programmatic launch it runs under the trunk's tail): the
    // first weight chunks in flight, barriers, TMEM, biases.
#pragma unroll
    for (int c = 0; c < HEAD_STAGES; ++c) issue_chunk(c);
    if (tid <= HEAD_STAGES) mbar_init(smem_u32(&s_bar[tid]), 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    for (int i = tid; i < HIDDEN; i += HEAD_THREADS) {
        s_b1[i] = Hp.b1[i];
        s_wv[i] = X3 ? Hp.wv_f32[i] : Hp.wv[i];
        s_b2[i] = i < Hp.A ? Hp.b2[i] : 0.f;
    }
    pdl_wait();  // the trunk's features and the leaf count are complete and visible from here on
    const int B = count_dev ? min(*count_dev, Bmax) : Bmax;
    if (row0 >= B) {  // no leaves for this tile: drain the copies, give the tensor memory back
        asm volatile("cp.async.wait_all;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512u));
        return;
    }
    // A operand: features of rows row0..row0+127 (zero beyond the batch): thread = (row, plane parity), 16 bytes per
    // plane by cp.async, all copies in flight at once
    {
        const int r = row0 + row_l;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(feat + (size_t)r * Hp.flat);
        const unsigned char* src_lo = reinterpret_cast<const unsigned char*>(feat + Hp.feat_lo_off + (size_t)r * Hp.flat);
        for (int p = hv; p < Hp.flat / 8; p += HEAD_NQ) {
            unsigned char* dst = areg + (size_t)p * 2048 + row_l * 16;
            if (r < B) {
                head_cp16(dst, src + p * 16);
                if (X3) head_cp16(dst + alo_off, src_lo + p * 16);
            } else {
                *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
                if (X3) *reinterpret_cast<uint4*>(dst + alo_off) = make_uint4(0, 0, 0, 0);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");  // the A operand (newest group) and the first weight chunks have landed
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    HP_T(1);
    const uint32_t tmem = s_tmem;
    const uint32_t bar0 = smem_u32(&s_bar[0]);
    const uint32_t bar_done = bar0 + 8u * HEAD_STAGES;
    uint32_t ph_stage = 0;  // bit s = parity to wait for on stage s
    uint32_t ph_done = 0;

    // consume chunks [c0, c1): wait for the chunk, issue its MMAs, refill the stage with chunk c + HEAD_STAGES
    auto run_chunks = [&](int c0, int c1, bool g2) {
        const int N = g2 ? Hp.N2 : HIDDEN, blk = g2 ? blk2 : blk1, kpc = g2 ? kpc2 : kpc1, nkb = g2 ? nkb2 : nkb1;
        const uint32_t idesc = umma_idesc(N), dcol = g2 ? 256u : 0u;
        for (int c = c0; c < c1; ++c) {
            const int st = c % HEAD_STAGES;
            asm volatile("cp.async.wait_group %0;" ::"n"(HEAD_STAGES - 2) : "memory");  // this thread's part of chunk c
            fence_proxy_async();
            __syncthreads();
            bool g2c, lo;
            int cc;
            chunk_info(c, g2c, lo, cc);
            const int kb = cc * kpc, nk = min(kpc, nkb - kb);
            if (warp == 0 && elect_one()) {
                tc_fence_after();
                const uint64_t a0 = umma_desc(smem_u32(areg), 128u, 8u);
                const uint64_t a0l = umma_desc(smem_u32(areg + alo_off), 128u, 8u);
                const uint64_t b0 = umma_desc(smem_u32(stage0 + (size_t)st * HEAD_STAGE_BYTES), (uint32_t)N, 8u);
                for (int k = 0; k < nk; ++k) {
                    const uint64_t ak = (uint64_t)((kb + k) * 2 * 2048 >> 4), bk = (uint64_t)(k * blk >> 4);
                    umma_bf16(tmem + dcol, a0 + ak, b0 + bk, idesc, (lo || (kb + k) > 0) ? 1u : 0u);  // a_hi * w_(hi|lo)
                    if (X3 && !lo) umma_bf16(tmem + dcol, a0l + ak, b0 + bk, idesc, 1u);                // a_lo * w_hi
                }
                umma_commit(bar0 + 8u * st);
                if (c == c1 - 1) umma_commit(bar_done);
            }
            __syncwarp();
            // refill the stage of the PREVIOUS chunk (its MMAs were issued one iteration ago and are normally complete
            // by now, so this wait does not stall) with chunk c - 1 + HEAD_STAGES
            const int pc = c - 1;
            if (pc >= 0 && pc + HEAD_STAGES < nch) {
                const int ps = pc % HEAD_STAGES;
                mbar_wait(bar0 + 8u * ps, (ph_stage >> ps) & 1u);
                ph_stage ^= 1u << ps;
                issue_chunk(pc + HEAD_STAGES);
            } else {
                asm volatile("cp.async.commit_group;" ::: "memory");  // keep one group per iteration
            }
        }
        mbar_wait(bar_done, ph_done);
        ph_done ^= 1u;
        tc_fence_after();
    };

    run_chunks(0, n1, false);
    HP_T(2);
    // epilogue 1: hidden = relu(acc + b1) -> bf16 planes (A operand of the logits GEMM) + value head dot product
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float vacc = 0.f;
    for (int c0 = hv * (HIDDEN / HEAD_NQ); c0 < (hv + 1) * (HIDDEN / HEAD_NQ); c0 += 16) {
        float v[16];
        tmem_ld16(lane_base + (uint32_t)c0, v);
        uint32_t pk[8], pl[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float h0 = fmaxf(v[2 * i] + s_b1[c0 + 2 * i], 0.f), h1 = fmaxf(v[2 * i + 1] + s_b1[c0 + 2 * i + 1], 0.f);
            pk[i] = pack_bf16(h0, h1);
            if (X3) {  // hi + lo halves of the hidden activations; the value head in full fp32
                pl[i] = pack_bf16(h0 - bf16_lo(pk[i]), h1 - bf16_hi(pk[i]));
                vacc = fmaf(h0, s_wv[c0 + 2 * i], vacc);
                vacc = fmaf(h1, s_wv[c0 + 2 * i + 1], vacc);
            } else {
                vacc = fmaf(bf16_lo(pk[i]), s_wv[c0 + 2 * i], vacc);
                vacc = fmaf(bf16_hi(pk[i]), s_wv[c0 + 2 * i + 1], vacc);
            }
        }
        *reinterpret_cast<uint4*>(areg + (size_t)(c0 / 8) * 2048 + row_l * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(areg + (size_t)(c0 / 8 + 1) * 2048 + row_l * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        if (X3) {
            *reinterpret_cast<uint4*>(areg + alo_off + (size_t)(c0 / 8) * 2048 + row_l * 16) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            *reinterpret_cast<uint4*>(areg + alo_off + (size_t)(c0 / 8 + 1) * 2048 + row_l * 16) = make_uint4(pl[4], pl[5], pl[6], pl[7]);
        }
    }
    s_x0[hv][row_l] = vacc;  // the value head's dot product: one part per thread
    tc_fence_before();
    __syncthreads();  // all hidden planes written (run_chunks fences them towards the async proxy before its first MMA)
    HP_T(3);
    run_chunks(n1, nch, true);
    HP_T(4);
    // epilogue 2 (the operand planes and the weight stages are free now and become the [128][A] policy tile):
    //   A. thread = (row, column half): logits = acc + b2 from TMEM into the tile, maximum of the half
    //   B. thread = (row, column half): e = exp(logit - row max) back into the tile, sum of the half
    //   C. whole CTA: policy = e / row sum, one row per warp and pass, coalesced stores
    const int r = row0 + row_l;
    if (hv == 0 && r < B) {
        float dot = 0.f;
#pragma unroll
        for (int q = 0; q < HEAD_NQ; ++q) dot += s_x0[q][row_l];
        value[r] = tanhf(dot + __ldg((X3 ? Hp.wv_f32 : Hp.wv) + HIDDEN));
    }
    float* s_pol = reinterpret_cast<float*>(hsm);
    const int ldp = Hp.A | 1;  // odd row stride: the 32 rows of a warp fall into distinct banks
    float* prow = s_pol + row_l * ldp;
    const int cw = (((Hp.A + HEAD_NQ - 1) / HEAD_NQ + 15) >> 4) << 4;  // columns per part, a multiple of the 16-column TMEM load
    const int cbeg = min(Hp.A, hv * cw), cend = min(Hp.A, (hv + 1) * cw);
    float mx = -INFINITY;
    for (int c0 = cbeg; c0 < cend; c0 += 16) {
        float v[16];
        tmem_ld16(lane_base + 256u + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (c0 + i < cend) {
                const float l = v[i] + s_b2[c0 + i];
                mx = fmaxf(mx, l);
                prow[c0 + i] = l;
            }
        }
    }
    s_x1[hv][row_l] = mx;
    tc_fence_before();
    __syncthreads();
    HP_T(5);
    mx = s_x1[0][row_l];
#pragma unroll
    for (int q = 1; q < HEAD_NQ; ++q) mx = fmaxf(mx, s_x1[q][row_l]);
    float sum = 0.f;
#pragma unroll 4
    for (int c = cbeg; c < cend; ++c) {
        const float e = __expf(prow[c] - mx);
        sum += e;
        prow[c] = e;
    }
    s_x0[hv][row_l] = sum;  // (the value head has consumed s_x0 before the barrier above)
    __syncthreads();
    HP_T(6);
    {
        const int nrows = min(128, B - row0), lane = tid & 31;
        for (int rr = warp; rr < nrows; rr += HEAD_THREADS / 32) {  // one row per warp and pass, 8 independent loads
            float den = 0.f;
#pragma unroll
            for (int q = 0; q < HEAD_NQ; ++q) den += s_x0[q][rr];
            const float inv = __fdividef(1.f, den);
            const float* src = s_pol + rr * ldp;
            float* dst = policy + (size_t)(row0 + rr) * Hp.A;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (lane + 32 * k < Hp.A) ? src[lane + 32 * k] : 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (lane + 32 * k < Hp.A) dst[lane + 32 * k] = v[k] * inv;
        }
    }
#ifdef BPP_HEADS_PROF
    HP_T(7);
    if (blockIdx.x == 0 && tid == 0)
        printf("heads cyc: prologue %lld gemm1 %lld epi1 %lld gemm2 %lld logits %lld exp %lld store %lld\n", t_[1] - t_[0],
               t_[2] - t_[1], t_[3] - t_[2], t_[4] - t_[3], t_[5] - t_[4], t_[6] - t_[5], t_[7] - t_[6]);
#endif
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

uint16_t f32_to_bf16_rne(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);  // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

}  // namespace

int bpp_set_error_message(int code, const char* msg);  // defined in bpp_engine.cu

struct bpp_net {
    NetParams P;
    int max_batch, device;
    std::map<std::string, std::vector<float>> host;
    std::map<std::string, long long> expect;  // name -> numel
    __nv_bfloat16* d_wts = nullptr;
    float* d_wts32 = nullptr;
    float* d_bias = nullptr;
    int precision = 0;  // BPP_NET_BF16
    bpptc::TcParams T;    // plain bf16 mode
    bpptc::TcParams T3;   // split-bf16 (x3) mode: doubled activation and weight buffers
    bool tc3_ok = false;
    __nv_bfloat16* d_wts_umma_lo = nullptr;
    __nv_bfloat16* d_wts_umma = nullptr;
    __nv_bfloat16* d_wts_logits_pad = nullptr;
    long long umma_elems = 0;
    long long* d_prof = nullptr;  // phase timers of CTA 0 (bpp_net_profile)
    __nv_bfloat16* d_feat = nullptr;     // [max_batch][flat] trunk output (bf16 mode)
    __nv_bfloat16* d_w1u = nullptr;      // FC weights in the UMMA B layout
    __nv_bfloat16* d_w2u = nullptr;
    float* d_wv32 = nullptr;
    __nv_bfloat16* d_w1u_lo = nullptr;   // split mode: low halves of the FC weights, exact fp32 value weights (+ bias)
    __nv_bfloat16* d_w2u_lo = nullptr;
    float* d_wvf = nullptr;
    HeadParams Hp;
    int heads_smem = 0, heads_smem3 = 0;
    bool heads_ok = false, heads3_ok = false;
    bool tc_ok = false;
    int ctas_per_sm = 1;
    bool committed = false;
    int smem_bytes = 0;
    int num_sms = 148;
    // role kernels (k_net_role): one plan per ConvSequence, bf16 [0] and split-bf16 [1]
    bpptc::TcParams Tr[2][3];
    int role_ctas[2][3] = {{2, 1, 2}, {1, 1, 1}};   // measured: profiles/r02_role_sweep.txt
    int roles_min_batch = 6144;  // bf16: below this batch the one-kernel trunk is faster (three launches, emptier groups)
    bool roles_ok[2] = {false, false};
    uint4* d_x1 = nullptr;   // hand-over buffers between the roles (sized for the split mode: hi + lo)
    uint4* d_x2 = nullptr;
    // grid-row trunk (k_net_gr, bpp_net_gr.cuh): one plan per level
    bppgr::GrStage Gr[4], Gr3[4];   // bf16, split-bf16
    bool gr_ok = false, gr3_ok = false;
    int gr_smem[2] = {0, 0}, gr_threads[2] = {0, 0};   // fused launch: dynamic shared memory, threads per CTA
    __nv_bfloat16* d_wts_gr = nullptr;
    __nv_bfloat16* d_wts_gr_lo = nullptr;
    long long gr_elems = 0;
    uint4* d_g[3] = {nullptr, nullptr, nullptr};   // hand-over buffers x1, x2, x3
};

static const char* kSeqConvNames[5] = {"conv", "res_block0.conv0", "res_block0.conv1", "res_block1.conv0",
                                       "res_block1.conv1"};

static int nerr(int code, const std::string& msg) { return bpp_set_error_message(code, msg.c_str()); }

extern "C" int bpp_net_create(int W, int H, int N, int max_batch, int device, bpp_net** out) {
    if (!out) return nerr(BPP_E_INVALID, "null argument");
    *out = nullptr;
    if (W < 1 || W > 32 || H < 1 || H > 28 || N < 1 || N > BPP_MAX_ITEMS || max_batch < 1)
        return nerr(BPP_E_INVALID, "unsupported network geometry");
    bpp_net* n = new bpp_net();
    n->max_batch = max_batch;
    n->device = device;
    NetParams& P = n->P;
    memset(&P, 0, sizeof(P));
    P.W = W; P.H = H; P.N = N; P.Cin = N + 1; P.A = W * N;
    P.hs[0] = H; P.ws[0] = W;
    for (int s = 0; s < 3; ++s) {
        P.hs[s + 1] = (P.hs[s] + 1) / 2;
        P.ws[s + 1] = (P.ws[s] + 1) / 2;
    }
    const int chans[3] = {16, 32, 32};
    P.flat = 32 * P.hs[3] * P.ws[3];
    long long woff = 0;
    int boff = 0, li = 0, cin = P.Cin;
    int buf = P.Cin * H * W;
    for (int s = 0; s < 3; ++s) {
        for (int k = 0; k < 5; ++k) {
            ConvDesc& d = P.conv[li++];
            d.ci = k == 0 ? cin : chans[s];
            d.co = chans[s];
            d.h = k == 0 ? P.hs[s] : P.hs[s + 1];
            d.w = k == 0 ? P.ws[s] : P.ws[s + 1];
            d.w_off = woff;
            d.b_off = boff;
            woff += (long long)d.ci * 9 * d.co;
            boff += d.co;
            if (d.co * d.h * d.w > buf) buf = d.co * d.h * d.w;
            const std::string base = "conv_seqs." + std::to_string(s) + "." + kSeqConvNames[k];
            n->expect[base + ".weight"] = (long long)d.co * d.ci * 9;
            n->expect[base + ".bias"] = d.co;
        }
        cin = chans[s];
    }
    if (buf < HIDDEN) buf = HIDDEN;
    if (buf < P.A) buf = P.A;
    P.buf_elems = (buf + 3) & ~3;
    woff = (woff + 7) & ~7ll;
    P.fc_hidden_off = woff; woff += (long long)P.flat * HIDDEN;
    P.fc_logits_off = woff; woff += (long long)HIDDEN * P.A;
    woff = (woff + 7) & ~7ll;
    P.fc_value_off = woff; woff += HIDDEN;
    P.b_hidden_off = boff; boff += HIDDEN;
    P.b_logits_off = boff; boff += P.A;
    P.b_value_off = boff; boff += 1;
    n->expect["hidden_fc.weight"] = (long long)HIDDEN * P.flat;
    n->expect["hidden_fc.bias"] = HIDDEN;
    n->expect["logits_fc.weight"] = (long long)P.A * HIDDEN;
    n->expect["logits_fc.bias"] = P.A;
    n->expect["value_fc.weight"] = HIDDEN;
    n->expect["value_fc.bias"] = 1;
    if (cudaSetDevice(device) != cudaSuccess) { delete n; return nerr(BPP_E_CUDA, "cudaSetDevice failed"); }
    if (cudaDeviceGetAttribute(&n->num_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n->num_sms < 1)
        n->num_sms = 148;
    {   // UMMA-layout conv weights (size known before the Tc setup below: recompute here)
        long long u = 0;
        for (int l = 0; l < NCONV; ++l) u += 9LL * ((P.conv[l].ci + 15) / 16) * 2 * P.conv[l].co * 8;
        if (cudaMalloc(&n->d_wts_umma, (size_t)u * sizeof(__nv_bfloat16)) != cudaSuccess ||
            cudaMalloc(&n->d_wts_umma_lo, (size_t)u * sizeof(__nv_bfloat16)) != cudaSuccess) {
            cudaGetLastError();
            delete n;
            return nerr(BPP_E_NOMEM, "cudaMalloc of the network parameters failed");
        }
    }
    if (cudaMalloc(&n->d_wts, (size_t)woff * sizeof(__nv_bfloat16)) != cudaSuccess ||
        cudaMalloc(&n->d_wts32, (size_t)woff * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&n->d_bias, (size_t)boff * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        delete n;
        return nerr(BPP_E_NOMEM, "cudaMalloc of the network parameters failed");
    }
    P.wts = n->d_wts;
    P.wts32 = n->d_wts32;
    P.bias = n->d_bias;
    // tensor-core path: pick the largest group size S <= 8 whose buffers fit in shared memory
    {
        bpptc::TcParams& T = n->T;
        memset(&T, 0, sizeof(T));
        long long uoff = 0;
        int wmax = 0;
        for (int l = 0; l < NCONV; ++l) {
            const int c16 = (P.conv[l].ci + 15) / 16;
            T.w_off[l] = uoff;
            T.lay_n16[l] = 9 * c16 * 2 * P.conv[l].co;
            const int bytes = 9 * c16 * 2 * P.conv[l].co * 16;
            uoff += bytes / 2;
            if (bytes > wmax) wmax = bytes;
        }
        n->umma_elems = uoff;
        const int cin16_0 = (P.Cin + 15) / 16;
        // ctas_want = CTAs per SM the plan is made for (bf16 mode: 1 or 2; they share the SM's 227 KB of shared memory,
        // its registers and its 512 TMEM columns); the largest group size S whose buffers fit is taken
        auto plan = [&](bpptc::TcParams& T, bool x3, int cap_bytes, int ctas_want, int& ctas_out, bool allow_compact) -> bool {
            const int f = x3 ? 2 : 1;
            const char* smax_env = getenv("BPP_TC_SMAX");  // experiments: cap the group size
            const int smax = smax_env ? std::max(1, std::min(8, atoi(smax_env))) : 8;
            for (int S = smax; S >= 1; --S) {
                for (int l = 0; l < 4; ++l) {
                    bpptc::Level& L = T.lv[l];
                    // shared halos: one zero column between consecutive grid rows (the right halo of row y IS the left
                    // halo of row y+1) and one zero row between consecutive samples, so a sample is (h+1) x (w+1) rows
                    // instead of (h+2) x (w+2); the guards in front of the first and behind the last sample stay zero
                    L.h = P.hs[l]; L.w = P.ws[l]; L.hp = L.h + 1; L.wp = L.w + 1; L.P = L.hp * L.wp;
                    L.mP = bpptc::fdiv_magic((uint32_t)L.P); L.mwp = bpptc::fdiv_magic((uint32_t)L.wp);
                    L.mhw = bpptc::fdiv_magic((uint32_t)(L.h * L.w)); L.mw = bpptc::fdiv_magic((uint32_t)L.w);
                    L.guard = (L.wp + 1 + 7) & ~7;
                    L.RT = L.guard + S * L.P + L.guard;
                    L.ntiles = (S * L.P + 127) / 128;
                }
                long long a = (long long)f * 2 * cin16_0 * T.lv[0].RT * 16, b = 0;
                for (int s2 = 0; s2 < 3; ++s2) {
                    const int planes = chans[s2] / 8;
                    a = std::max(a, (long long)f * 3 * planes * T.lv[s2 + 1].RT * 16);
                    b = std::max(b, (long long)f * planes * T.lv[s2].RT * 16);
                }
                const int NSs = S <= 4 ? 4 : 8;
                if (x3 || ctas_want < 2)  // in-kernel CUDA-core heads need their scratch
                    b = std::max(b, (long long)NSs * (P.flat + 4 * HIDDEN + 3 * (P.A + 1)) * 4);
                // the last tile's shifted windows over-read up to 128 + wp + 1 rows behind region A: region B must cover that
                const long long over = (long long)(128 + T.lv[0].wp + 8) * 16;
                b = std::max(b, over);
                T.S = S;
                T.compact = (allow_compact && !x3 && ctas_want >= 2) ? 1 : 0;
                if (T.compact) {
                    // compact arena (see k_net_forward_tc): A = max(input [= aliased T_0], T_s + conv weights of
                    // sequences 1 and 2, residual weights); B = max(level triples, level-0 conv weights) + over-read slack
                    auto wbytes = [&](int l) { return (long long)T.lay_n16[l] * 16 + 256; };
                    long long ca = (long long)2 * cin16_0 * T.lv[0].RT * 16, cb = wbytes(0);
                    for (int s2 = 0; s2 < 3; ++s2) {
                        const long long planes = chans[s2] / 8;
                        if (s2 > 0) ca = std::max(ca, ((planes * T.lv[s2].RT * 16 + 127) & ~127LL) + wbytes(5 * s2));
                        cb = std::max(cb, 3 * planes * T.lv[s2 + 1].RT * 16);
                        for (int l = 5 * s2 + 1; l < 5 * s2 + 5; ++l) ca = std::max(ca, wbytes(l));
                    }
                    if (chans[0] / 8 > 2 * cin16_0) T.compact = 0;  // T_0 must fit into the input's planes
                    else { a = ca; b = cb + over; }
                }
                T.regA_bytes = (int)((a + 127) & ~127LL);
                T.regB_bytes = (int)((b + 127) & ~127LL);
                T.wbuf_bytes = T.compact ? 0 : ((f * wmax + 256 + 127) & ~127);     // + the layer's bias behind the weights
                T.smem_bytes = T.regA_bytes + T.regB_bytes + T.wbuf_bytes;
                if (T.smem_bytes <= cap_bytes && T.lv[0].RT < 16384) {
                    T.tmem_cols = ctas_want == 2 ? 256 : 512;
                    ctas_out = ctas_want;
                    return true;
                }
            }
            return false;
        };
        {
            // 227 KB per SM, 1 KB reserved per CTA, 4.3 KB static shared memory per CTA.  Two CTAs per SM: while one waits
            // for its MMAs the other runs its CUDA-core phases (BPP_TC_CTAS=1: one CTA with larger groups, experiments)
            const char* e = getenv("BPP_TC_CTAS");
            int want = e ? atoi(e) : 2;
            if (want < 1 || want > 2) want = 2;
            const int hk = (P.flat > HIDDEN ? P.flat : HIDDEN) / 8;
            const int hsm = std::max(hk * 2048 + HEAD_STAGES * HEAD_STAGE_BYTES, 128 * ((P.A | 1) + 1) * 4);
            const bool heads_possible = (P.flat % 16 == 0) && ((P.A + 15) & ~15) <= 256 && hsm <= 220 * 1024 &&
                                        getenv("BPP_NO_TC_HEADS") == nullptr;
            if (!heads_possible) want = 1;  // the multi-CTA instantiations are trunk-only
            // The compact arena holds larger groups but costs ~3 % at equal group size (measured: 15x15, S = 5, both
            // layouts), so it is taken only where it buys at least a third more leaves per group (20x20: 4 instead of 3,
            // +8..15 %); BPP_TC_COMPACT=0/1 forces the choice.
            const int cap = (227 * 1024) / want - 1024 - 4352;
            bpptc::TcParams Tc = T;
            int ctas_c = 1;
            n->tc_ok = plan(T, false, cap, want, n->ctas_per_sm, false);
            const bool okc = plan(Tc, false, cap, want, ctas_c, true) && Tc.compact;
            const char* fc = getenv("BPP_TC_COMPACT");
            const bool take = okc && (fc ? atoi(fc) != 0 : (!n->tc_ok || 3 * Tc.S >= 4 * T.S));
            if (take) {
                T = Tc;
                n->ctas_per_sm = ctas_c;
                n->tc_ok = true;
            }
        }
        if (getenv("BPP_TC_VERBOSE"))
            fprintf(stderr, "bpp_net: tcgen05 plan %dx%d: S = %d, %d CTA(s)/SM, %s arena, %d B shared memory (A %d, B %d)\n",
                    P.W, P.H, T.S, n->ctas_per_sm, T.compact ? "compact" : "classic", T.smem_bytes, T.regA_bytes,
                    T.regB_bytes);
        n->T3 = T;
        int c3 = 1;
        n->tc3_ok = plan(n->T3, true, 220 * 1024, 1, c3, false);
        // ---- role kernels: one plan per ConvSequence.  Region A = the role's input planes (level s), later the
        // {raw, actA, actB} triple of level s+1; region B = the conv output T_s awaiting pooling; one layer of weights.
        {
            auto plan_role = [&](bpptc::TcParams& R, int seq, bool x3, int ctas) -> bool {
                const int f = x3 ? 2 : 1;
                const int cap_bytes = (227 * 1024) / ctas - 1024 - 4352;
                const int cin16 = seq == 0 ? cin16_0 : chans[seq - 1] / 16;
                const int planes_out = chans[seq] / 8;
                int wrole = 0;
                for (int l = 5 * seq; l < 5 * seq + 5; ++l) wrole = std::max(wrole, T.lay_n16[l] * 16);
                const char* smax_env = getenv(seq == 0 ? "BPP_ROLE_S0" : seq == 1 ? "BPP_ROLE_S1" : "BPP_ROLE_S2");
                const int smax = smax_env ? std::max(1, atoi(smax_env)) : (seq == 0 ? 8 : 48);
                for (int S = std::min(smax, seq == 0 ? 8 : 48); S >= 1; --S) {
                    R = T;  // weight tables, pointers
                    for (int l = 0; l < 4; ++l) {
                        bpptc::Level& L = R.lv[l];
                        L.h = P.hs[l]; L.w = P.ws[l]; L.hp = L.h + 1; L.wp = L.w + 1; L.P = L.hp * L.wp;
                        L.mP = bpptc::fdiv_magic((uint32_t)L.P); L.mwp = bpptc::fdiv_magic((uint32_t)L.wp);
                        L.mhw = bpptc::fdiv_magic((uint32_t)(L.h * L.w)); L.mw = bpptc::fdiv_magic((uint32_t)L.w);
                        L.guard = (L.wp + 1 + 7) & ~7;
                        L.RT = L.guard + S * L.P + L.guard;
                        L.ntiles = (S * L.P + 127) / 128;
                    }
                    const bpptc::Level& La = R.lv[seq];
                    const bpptc::Level& Lb = R.lv[seq + 1];
                    const long long a = std::max((long long)f * 2 * cin16 * La.RT * 16, (long long)f * 3 * planes_out * Lb.RT * 16);
                    const long long over = (long long)(128 + La.wp + 8) * 16;
                    const long long b = std::max((long long)f * planes_out * La.RT * 16, over);
                    R.S = S;
                    R.compact = 0;
                    R.regA_bytes = (int)((a + 127) & ~127LL);
                    R.regB_bytes = (int)((b + 127) & ~127LL);
                    R.wbuf_bytes = (f * wrole + 256 + 127) & ~127;
                    R.smem_bytes = R.regA_bytes + R.regB_bytes + R.wbuf_bytes;
                    R.tmem_cols = ctas == 1 ? 512 : ctas == 2 ? 256 : 128;
                    if (R.smem_bytes <= cap_bytes && La.RT < 16384 && Lb.RT < 16384) return true;
                }
                return false;
            };
            for (int x = 0; x < 2; ++x) {
                const char* ce = getenv(x ? "BPP_ROLE_CTAS_X3" : "BPP_ROLE_CTAS");   // e.g. "2,2,2"
                if (ce) sscanf(ce, "%d,%d,%d", &n->role_ctas[x][0], &n->role_ctas[x][1], &n->role_ctas[x][2]);
                if (const char* mb = getenv("BPP_ROLES_MIN_BATCH")) n->roles_min_batch = atoi(mb);
                bool ok = getenv("BPP_NO_ROLES") == nullptr && (P.flat % 16 == 0);
                for (int sq = 0; sq < 3 && ok; ++sq) {
                    int& c = n->role_ctas[x][sq];
                    if (c < 1 || c > 2) c = x ? 1 : 2;
                    ok = plan_role(n->Tr[x][sq], sq, x != 0, c);
                }
                n->roles_ok[x] = ok;
                if (ok && getenv("BPP_TC_VERBOSE"))
                    for (int sq = 0; sq < 3; ++sq)
                        fprintf(stderr, "bpp_net: role %d (%s): S = %d, %d CTA(s)/SM, %d B shared memory (A %d, B %d, W %d)\n", sq,
                                x ? "bf16x3" : "bf16", n->Tr[x][sq].S, n->role_ctas[x][sq], n->Tr[x][sq].smem_bytes,
                                n->Tr[x][sq].regA_bytes, n->Tr[x][sq].regB_bytes, n->Tr[x][sq].wbuf_bytes);
            }
        }
    }
    n->smem_bytes = 3 * P.buf_elems * (int)sizeof(float);
    if (cudaFuncSetAttribute(k_net_forward<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, n->smem_bytes) !=
            cudaSuccess ||
        cudaFuncSetAttribute(k_net_forward<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n->smem_bytes) !=
            cudaSuccess ||
        (n->tc_ok && (cudaFuncSetAttribute(k_net_forward_tc<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           n->T.smem_bytes) != cudaSuccess ||
                      cudaFuncSetAttribute(k_net_forward_tc<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           n->T.smem_bytes) != cudaSuccess ||
                      cudaFuncSetAttribute(k_net_forward_tc<8, false, true, 2>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, n->T.smem_bytes) != cudaSuccess)) ||
        (n->tc3_ok && (cudaFuncSetAttribute(k_net_forward_tc<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            n->T3.smem_bytes) != cudaSuccess ||
                       cudaFuncSetAttribute(k_net_forward_tc<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            n->T3.smem_bytes) != cudaSuccess))) {
        cudaGetLastError();
        delete n;
        return nerr(BPP_E_CUDA, "cannot reserve shared memory for the forward kernel");
    }
    n->T.wts_umma = n->d_wts_umma;
    n->T.wts_umma_lo = n->d_wts_umma_lo;
    n->T.A_pad = (P.A + 1) & ~1;
    n->T.mH = bpptc::fdiv_magic((uint32_t)P.H);
    if (cudaMalloc(&n->d_wts_logits_pad, (size_t)HIDDEN * n->T.A_pad * 2) != cudaSuccess) {
        cudaGetLastError();
        delete n;
        return nerr(BPP_E_NOMEM, "cudaMalloc of the network parameters failed");
    }
    n->T.wts_logits_pad = n->d_wts_logits_pad;
    {   // tensor-core heads
        HeadParams& Hp = n->Hp;
        Hp.flat = P.flat; Hp.A = P.A; Hp.N2 = (P.A + 15) & ~15;
        const int kplanes = (P.flat > HIDDEN ? P.flat : HIDDEN) / 8;
        n->heads_smem = std::max(kplanes * 2048 + HEAD_STAGES * HEAD_STAGE_BYTES, 128 * ((P.A | 1) + 1) * 4);
        n->heads_ok = (P.flat % 16 == 0) && Hp.N2 <= 256 && n->heads_smem <= 220 * 1024 && getenv("BPP_NO_TC_HEADS") == nullptr;
        n->heads_smem3 = std::max(2 * kplanes * 2048 + HEAD_STAGES * HEAD_STAGE_BYTES, 128 * ((P.A | 1) + 1) * 4);
        n->heads3_ok = n->heads_ok && n->heads_smem3 <= 220 * 1024;
        if (n->heads_ok) {
            if (cudaMalloc(&n->d_feat, (size_t)max_batch * P.flat * 2 * 2) != cudaSuccess ||   // hi [+ lo]
                cudaMalloc(&n->d_w1u, (size_t)P.flat * HIDDEN * 2) != cudaSuccess ||
                cudaMalloc(&n->d_w2u, (size_t)HIDDEN * Hp.N2 * 2) != cudaSuccess ||
                cudaMalloc(&n->d_w1u_lo, (size_t)P.flat * HIDDEN * 2) != cudaSuccess ||
                cudaMalloc(&n->d_w2u_lo, (size_t)HIDDEN * Hp.N2 * 2) != cudaSuccess ||
                cudaMalloc(&n->d_wv32, (HIDDEN + 1) * sizeof(float)) != cudaSuccess ||
                cudaMalloc(&n->d_wvf, (HIDDEN + 1) * sizeof(float)) != cudaSuccess ||
                cudaFuncSetAttribute(k_net_heads_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, n->heads_smem) != cudaSuccess ||
                (n->heads3_ok && cudaFuncSetAttribute(k_net_heads_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      n->heads_smem3) != cudaSuccess)) {
                cudaGetLastError();
                delete n;
                return nerr(BPP_E_NOMEM, "cudaMalloc of the head buffers failed");
            }
            Hp.w1u = n->d_w1u; Hp.w2u = n->d_w2u; Hp.wv = n->d_wv32;
            Hp.w1u_lo = n->d_w1u_lo; Hp.w2u_lo = n->d_w2u_lo; Hp.wv_f32 = n->d_wvf;
            Hp.feat_lo_off = (long long)max_batch * P.flat;
            Hp.b1 = n->d_bias + P.b_hidden_off; Hp.b2 = n->d_bias + P.b_logits_off;
        }
    }
    {   // the split-mode plan shares the weight pointers and layer tables
        bpptc::TcParams keep = n->T3;
        n->T3 = n->T;
        n->T3.S = keep.S; n->T3.regA_bytes = keep.regA_bytes; n->T3.regB_bytes = keep.regB_bytes;
        n->T3.wbuf_bytes = keep.wbuf_bytes; n->T3.smem_bytes = keep.smem_bytes; n->T3.tmem_cols = keep.tmem_cols;
        for (int l = 0; l < 4; ++l) n->T3.lv[l] = keep.lv[l];
    }
    for (int x = 0; x < 2; ++x)
        for (int sq = 0; sq < 3; ++sq) {  // the role plans share the weight pointers and tables too
            bpptc::TcParams& R = n->Tr[x][sq];
            R.wts_umma = n->T.wts_umma; R.wts_umma_lo = n->T.wts_umma_lo; R.A_pad = n->T.A_pad; R.mH = n->T.mH;
            R.wts_logits_pad = n->T.wts_logits_pad;
            for (int l = 0; l < NCONV; ++l) { R.w_off[l] = n->T.w_off[l]; R.lay_n16[l] = n->T.lay_n16[l]; }
        }
    if (n->roles_ok[0] || n->roles_ok[1]) {
        const size_t x1 = (size_t)max_batch * 2 * 2 * P.hs[1] * P.ws[1] * 16, x2 = (size_t)max_batch * 2 * 4 * P.hs[2] * P.ws[2] * 16;
        bool ok = cudaMalloc(&n->d_x1, x1) == cudaSuccess && cudaMalloc(&n->d_x2, x2) == cudaSuccess;
        auto attr = [&](const void* fn, int bytes) {
            return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
        };
        if (ok && n->roles_ok[0])
            ok = attr((const void*)k_net_role<0, false, 2>, n->Tr[0][0].smem_bytes) &&
                 attr((const void*)k_net_role<1, false, 2>, n->Tr[0][1].smem_bytes) &&
                 attr((const void*)k_net_role<2, false, 2>, n->Tr[0][2].smem_bytes) &&
                 attr((const void*)k_net_role<0, false, 1>, n->Tr[0][0].smem_bytes) &&
                 attr((const void*)k_net_role<1, false, 1>, n->Tr[0][1].smem_bytes) &&
                 attr((const void*)k_net_role<2, false, 1>, n->Tr[0][2].smem_bytes);
        if (ok && n->roles_ok[1])
            ok = attr((const void*)k_net_role<0, true, 2>, n->Tr[1][0].smem_bytes) &&
                 attr((const void*)k_net_role<1, true, 2>, n->Tr[1][1].smem_bytes) &&
                 attr((const void*)k_net_role<2, true, 2>, n->Tr[1][2].smem_bytes) &&
                 attr((const void*)k_net_role<0, true, 1>, n->Tr[1][0].smem_bytes) &&
                 attr((const void*)k_net_role<1, true, 1>, n->Tr[1][1].smem_bytes) &&
                 attr((const void*)k_net_role<2, true, 1>, n->Tr[1][2].smem_bytes);
        if (!ok) {
            cudaGetLastError();
            n->roles_ok[0] = n->roles_ok[1] = false;
        }
    }

    // ---- grid-row trunk (bpp_net_gr.cuh): one plan per level and arithmetic mode.  J = leaves per group (J * (w+1) <= 128
    // rows = one MMA tile per grid row).  bf16: two groups in flight per CTA where shared memory allows it.  Split-bf16 (x3):
    // arena and weights are doubled (hi + lo), one group per CTA, and where the level's weights do not fit next to the arena
    // they stream through two slots.
    {
        const int cin16_0 = (P.Cin + 15) / 16;
        const int chans_in[4] = {16 * cin16_0, 16, 32, 32};
        const int cap = 232448 - 4608;   // 227 KB per block minus the kernel's static shared memory (barriers, row map, biases)
        n->gr_elems = 0;   // all 15 layers in the grid-row layout, whichever plans succeed
        for (int l = 0; l < NCONV; ++l) n->gr_elems += 9LL * ((P.conv[l].ci + 15) / 16) * 2 * P.conv[l].co * 8;
        for (int x3 = 0; x3 < 2; ++x3) {
            long long goff = 0;
            bool ok = getenv(x3 ? "BPP_NO_GR3" : "BPP_NO_GR") == nullptr && P.flat % 16 == 0 && (x3 ? n->heads3_ok : n->heads_ok);
            const int f = x3 ? 2 : 1;
            for (int s = 0; s < 4 && ok; ++s) {
                bppgr::GrStage& G = (x3 ? n->Gr3 : n->Gr)[s];
                memset(&G, 0, sizeof(G));
                G.h = P.hs[s]; G.w = P.ws[s]; G.wp = G.w + 1; G.NT = G.h;
                G.h2 = s < 3 ? P.hs[s + 1] : 0; G.w2 = s < 3 ? P.ws[s + 1] : 0;
                G.cp = chans_in[s] / 8;
                G.arena_planes = s == 0 ? std::max(G.cp, 2) : 2 * G.cp;
                G.planes_out = s < 3 ? chans[s] / 8 : 0;
                // layers: stage 0 = conv 0; stage s = the four residual convs at this level, then the next sequence's conv
                int first = s == 0 ? 0 : 5 * (s - 1) + 1;
                G.nlay = s == 0 ? 1 : s == 3 ? 4 : 5;
                int off = 0, wmax = 0;
                for (int l = 0; l < G.nlay; ++l) {
                    const ConvDesc& d = P.conv[first + l];
                    G.cin16[l] = (d.ci + 15) / 16;
                    G.cout[l] = d.co;
                    G.w_len[l] = 9 * G.cin16[l] * 2 * d.co * 16;
                    G.w_soff[l] = off;
                    G.w_goff[l] = goff;
                    G.b_goff[l] = d.b_off;
                    off += f * G.w_len[l];
                    goff += G.w_len[l] / 2;
                    wmax = std::max(wmax, f * G.w_len[l]);
                }
                // the residual layers (and stage 2's conv) keep one accumulator slot per grid row: 16 slots of 16 or 8 of 32 columns
                if (G.NT > bppgr::MAX_TILES || G.wp > 128 || (s >= 1 && G.NT > (s == 1 ? 16 : 8))) { ok = false; break; }
                const char* je = getenv(s == 0 ? "BPP_GR_J0" : s == 1 ? "BPP_GR_J1" : s == 2 ? "BPP_GR_J2" : "BPP_GR_J3");
                int jmax = std::min(128 / G.wp, s == 0 ? 8 : 255);
                if (je) jmax = std::max(1, std::min(jmax, atoi(je)));
                const char* se = getenv("BPP_GR_NSUB");
                const int nsub_max = x3 ? 1 : (se ? std::max(1, std::min(2, atoi(se))) : 2);
                bool fit = false;
                // every (J, groups per CTA, weights resident | streamed) that fits; the one with the most leaves in flight per
                // SM wins (streamed weights count 10 % less: their layers are not chained tile by tile), ties go to larger groups
                bppgr::GrStage best = G;
                double best_score = -1.0;
                for (int J = jmax; J >= 1; --J)
                    for (int ns = nsub_max; ns >= 1; --ns)
                        for (int stream = 0; stream < (x3 && G.nlay > 2 ? 2 : 1); ++stream) {
                            G.J = J;
                            G.TS = (J * G.wp + 7) & ~7;
                            G.RT = bppgr::G0 + (G.NT - 1) * G.TS + 128 + 8;
                            G.lo_off = G.arena_planes * G.RT * 16;
                            G.arena_bytes = f * G.lo_off;
                            G.nsub = ns;
                            G.stream = stream;
                            G.slot_bytes = (wmax + 127) & ~127;
                            G.w_bytes = stream ? 2 * G.slot_bytes : off;
                            G.arena_off = (G.w_bytes + G.nlay * 32 * 4 + 127) & ~127;
                            G.smem_bytes = G.arena_off + ns * G.arena_bytes + (s == 0 ? (int)sizeof(bppgr::GrStage0Tables) : 0);
                            if (G.smem_bytes > cap || G.RT >= 16384) continue;
                            const double score = (double)J * ns * (stream ? 0.9 : 1.0) + 1e-3 * J;
                            if (score > best_score) { best_score = score; best = G; fit = true; }
                        }
                G = best;
                if (!fit) { ok = false; break; }
                // accumulators: a ring of 256 TMEM columns per group in flight (16 slots of 16 or 8 slots of 32 columns)
                G.tmem_cols = G.nsub == 2 ? 512 : 256;
                G.col_sub = 256;
                G.m_w = bpptc::fdiv_magic((uint32_t)G.w);
                G.m_w2 = bpptc::fdiv_magic((uint32_t)std::max(1, G.w2));
                G.m_hw2 = bpptc::fdiv_magic((uint32_t)std::max(1, G.h2 * G.w2));
                G.m_php2 = bpptc::fdiv_magic((uint32_t)std::max(1, G.planes_out * G.h2 * G.w2));
                G.m_flat = bpptc::fdiv_magic((uint32_t)P.flat);
                G.m_pw2 = bpptc::fdiv_magic((uint32_t)std::max(1, G.planes_out * G.w2));
                if (getenv("BPP_TC_VERBOSE"))
                    fprintf(stderr, "bpp_net: grid-row stage %d (%s): %dx%d, J = %d, tile stride %d, %d group(s) per CTA, weights %s, "
                            "%d B shared memory (weights %d)\n", s, x3 ? "bf16x3" : "bf16", G.h, G.w, G.J, G.TS, G.nsub,
                            G.stream ? "streamed" : "resident", G.smem_bytes, G.w_bytes);
            }
            (x3 ? n->gr3_ok : n->gr_ok) = ok;
        }
        if (n->gr_ok || n->gr3_ok) {   // weights in the grid-row layout (hi, lo) and the hand-over buffers x1..x3 (hi + lo)
            const size_t x1 = (size_t)max_batch * 2 * 2 * P.hs[1] * P.ws[1] * 16, x2 = (size_t)max_batch * 2 * 4 * P.hs[2] * P.ws[2] * 16,
                         x3b = (size_t)max_batch * 2 * 4 * P.hs[3] * P.ws[3] * 16;
            bool ok = cudaMalloc(&n->d_wts_gr, (size_t)n->gr_elems * 2) == cudaSuccess &&
                      cudaMalloc(&n->d_wts_gr_lo, (size_t)n->gr_elems * 2) == cudaSuccess && cudaMalloc(&n->d_g[0], x1) == cudaSuccess &&
                      cudaMalloc(&n->d_g[1], x2) == cudaSuccess && cudaMalloc(&n->d_g[2], x3b) == cudaSuccess;
            auto attr = [&](const void* fn, int bytes) {
                return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
            };
            for (int x3 = 0; x3 < 2; ++x3) {   // one launch runs the four stages: common TMEM allocation, largest footprint
                bppgr::GrStage* G = x3 ? n->Gr3 : n->Gr;
                int cols = 0, sm = 0, ns = 1;
                for (int s2 = 0; s2 < 4; ++s2) { cols = std::max(cols, G[s2].tmem_cols); sm = std::max(sm, G[s2].smem_bytes); ns = std::max(ns, G[s2].nsub); }
                for (int s2 = 0; s2 < 4; ++s2) G[s2].tmem_cols = cols;
                n->gr_smem[x3] = sm;
                n->gr_threads[x3] = ns * bppgr::SUB_THREADS;
            }
            if (ok && n->gr_ok) ok = attr((const void*)bppgr::k_net_gr<false>, n->gr_smem[0]);
            if (ok && n->gr3_ok) ok = attr((const void*)bppgr::k_net_gr<true>, n->gr_smem[1]);
            if (!ok) {
                cudaGetLastError();
                n->gr_ok = n->gr3_ok = false;
            }
        }
    }
    if (cudaMalloc(&n->d_prof, 32 * sizeof(long long)) == cudaSuccess) cudaMemset(n->d_prof, 0, 32 * sizeof(long long));
    *out = n;
    return BPP_OK;
}

extern "C" int bpp_net_destroy(bpp_net* n) {
    if (!n) return BPP_OK;
    cudaFree(n->d_wts);
    cudaFree(n->d_wts32);
    cudaFree(n->d_wts_umma);
    cudaFree(n->d_wts_umma_lo);
    cudaFree(n->d_wts_logits_pad);
    cudaFree(n->d_prof);
    cudaFree(n->d_feat);
    cudaFree(n->d_w1u);
    cudaFree(n->d_w2u);
    cudaFree(n->d_wv32);
    cudaFree(n->d_w1u_lo);
    cudaFree(n->d_w2u_lo);
    cudaFree(n->d_wvf);
    cudaFree(n->d_bias);
    cudaFree(n->d_x1);
    cudaFree(n->d_x2);
    cudaFree(n->d_wts_gr);
    cudaFree(n->d_wts_gr_lo);
    for (int i = 0; i < 3; ++i) cudaFree(n->d_g[i]);
    delete n;
    return BPP_OK;
}

extern "C" int bpp_net_set_param(bpp_net* n, const char* name, const float* data_host, int64_t numel) {
    if (!n || !name || !data_host) return nerr(BPP_E_INVALID, "null argument");
    auto it = n->expect.find(name);
    if (it == n->expect.end()) return nerr(BPP_E_INVALID, std::string("unknown parameter name: ") + name);
    if (it->second != numel)
        return nerr(BPP_E_INVALID, std::string("parameter ") + name + " has " + std::to_string(numel) +
                                       " elements, expected " + std::to_string(it->second));
    n->host[name].assign(data_host, data_host + numel);
    n->committed = false;
    return BPP_OK;
}

extern "C" int bpp_net_commit(bpp_net* n, void* stream) {
    if (!n) return nerr(BPP_E_INVALID, "null argument");
    for (auto& kv : n->expect)
        if (!n->host.count(kv.first)) return nerr(BPP_E_STATE, "parameter not set: " + kv.first);
    const NetParams& P = n->P;
    const long long nw = P.fc_value_off + HIDDEN;
    const int nb = P.b_value_off + 1;
    std::vector<uint16_t> w((size_t)nw, 0);
    std::vector<float> w32((size_t)nw, 0.f);
    std::vector<float> b((size_t)nb, 0.f);
    auto put = [&](size_t idx, float v) { w[idx] = f32_to_bf16_rne(v); w32[idx] = v; };
    int li = 0;
    for (int s = 0; s < 3; ++s)
        for (int k = 0; k < 5; ++k) {
            const ConvDesc& d = P.conv[li++];
            const std::string base = "conv_seqs." + std::to_string(s) + "." + kSeqConvNames[k];
            const std::vector<float>& src = n->host[base + ".weight"];  // OIHW
            for (int co = 0; co < d.co; ++co)
                for (int ci = 0; ci < d.ci; ++ci)
                    for (int t = 0; t < 9; ++t)
                        put((size_t)d.w_off + ((size_t)ci * 9 + t) * d.co + co, src[((size_t)co * d.ci + ci) * 9 + t]);
            const std::vector<float>& bs = n->host[base + ".bias"];
            for (int co = 0; co < d.co; ++co) b[d.b_off + co] = bs[co];
        }
    {
        const std::vector<float>& src = n->host["hidden_fc.weight"];  // [256][flat]
        for (int o = 0; o < HIDDEN; ++o)
            for (int i = 0; i < P.flat; ++i)
                put((size_t)P.fc_hidden_off + (size_t)i * HIDDEN + o, src[(size_t)o * P.flat + i]);
        const std::vector<float>& src2 = n->host["logits_fc.weight"];  // [A][256]
        for (int o = 0; o < P.A; ++o)
            for (int i = 0; i < HIDDEN; ++i)
                put((size_t)P.fc_logits_off + (size_t)i * P.A + o, src2[(size_t)o * HIDDEN + i]);
        const std::vector<float>& src3 = n->host["value_fc.weight"];
        for (int i = 0; i < HIDDEN; ++i) put((size_t)P.fc_value_off + i, src3[i]);
        memcpy(&b[P.b_hidden_off], n->host["hidden_fc.bias"].data(), HIDDEN * sizeof(float));
        memcpy(&b[P.b_logits_off], n->host["logits_fc.bias"].data(), (size_t)P.A * sizeof(float));
        b[P.b_value_off] = n->host["value_fc.bias"][0];
    }
    // conv weights in the UMMA K-major B layout: [tap][kc][k-half][cout][8 cin]
    std::vector<uint16_t> wu((size_t)n->umma_elems, 0), wul((size_t)n->umma_elems, 0);
    li = 0;
    for (int s = 0; s < 3; ++s)
        for (int k = 0; k < 5; ++k) {
            const ConvDesc& d = P.conv[li];
            const std::string base = "conv_seqs." + std::to_string(s) + "." + kSeqConvNames[k];
            const std::vector<float>& src = n->host[base + ".weight"];  // OIHW
            const int c16 = (d.ci + 15) / 16;
            for (int t = 0; t < 9; ++t)
                for (int kc = 0; kc < c16; ++kc)
                    for (int kh = 0; kh < 2; ++kh)
                        for (int co = 0; co < d.co; ++co)
                            for (int j = 0; j < 8; ++j) {
                                const int ci = kc * 16 + kh * 8 + j;
                                const float v = ci < d.ci ? src[((size_t)co * d.ci + ci) * 9 + t] : 0.f;
                                const size_t ui = (size_t)n->T.w_off[li] + ((((size_t)t * c16 + kc) * 2 + kh) * d.co + co) * 8 + j;
                                wu[ui] = f32_to_bf16_rne(v);
                                uint32_t hb = (uint32_t)wu[ui] << 16;
                                float hf;
                                memcpy(&hf, &hb, 4);
                                wul[ui] = f32_to_bf16_rne(v - hf);  // low half for the split-bf16 mode
                            }
            ++li;
        }
    // conv weights for the grid-row kernels: per layer [dx][kc][k-half][dy * cout + co][8 cin] - the three vertical taps of
    // one horizontal tap are the N axis of one UMMA B operand (bpp_net_gr.cuh)
    std::vector<uint16_t> wg, wgl;
    if (n->gr_ok || n->gr3_ok) {
        wg.assign((size_t)n->gr_elems, 0);
        wgl.assign((size_t)n->gr_elems, 0);
        for (int s = 0; s < 4; ++s) {
            const bppgr::GrStage& G = n->gr_ok ? n->Gr[s] : n->Gr3[s];   // (layer tables are the same in both plans)
            const int first = s == 0 ? 0 : 5 * (s - 1) + 1;
            for (int l = 0; l < G.nlay; ++l) {
                const ConvDesc& d = P.conv[first + l];
                const int sq = (first + l) / 5, k = (first + l) % 5;
                const std::vector<float>& src = n->host["conv_seqs." + std::to_string(sq) + "." + kSeqConvNames[k] + ".weight"];  // OIHW
                const int c16 = G.cin16[l];
                for (int dx = 0; dx < 3; ++dx)
                    for (int kc = 0; kc < c16; ++kc)
                        for (int kh = 0; kh < 2; ++kh)
                            for (int dy = 0; dy < 3; ++dy)
                                for (int co = 0; co < d.co; ++co)
                                    for (int j = 0; j < 8; ++j) {
                                        const int ci = kc * 16 + kh * 8 + j;
                                        const float v = ci < d.ci ? src[((size_t)co * d.ci + ci) * 9 + dy * 3 + dx] : 0.f;
                                        const size_t ui = (size_t)G.w_goff[l] +
                                                          ((((size_t)dx * c16 + kc) * 2 + kh) * (3 * d.co) + dy * d.co + co) * 8 + j;
                                        wg[ui] = f32_to_bf16_rne(v);
                                        uint32_t hb = (uint32_t)wg[ui] << 16;
                                        float hf;
                                        memcpy(&hf, &hb, 4);
                                        wgl[ui] = f32_to_bf16_rne(v - hf);   // low half for the split-bf16 mode
                                    }
            }
        }
    }
    // every upload below is queued on the caller's stream (a forward or a captured graph still in flight on that stream is
    // ordered before it) and the staging vectors live until the cudaStreamSynchronize at the end
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    std::vector<uint16_t> w1u, w2u, w1l, w2l;
    std::vector<float> wv32, wvf;
    auto split_bf16 = [](float v, uint16_t& hi, uint16_t& lo) {
        hi = f32_to_bf16_rne(v);
        uint32_t hb = (uint32_t)hi << 16;
        float hf;
        memcpy(&hf, &hb, 4);
        lo = f32_to_bf16_rne(v - hf);
    };
    if (n->heads_ok) {  // FC weights in the UMMA K-major B layout [k-block][k-half][n][8]
        const HeadParams& Hp = n->Hp;
        w1u.assign((size_t)P.flat * HIDDEN, 0);
        w2u.assign((size_t)HIDDEN * Hp.N2, 0);
        w1l.assign((size_t)P.flat * HIDDEN, 0);
        w2l.assign((size_t)HIDDEN * Hp.N2, 0);
        wv32.assign(HIDDEN + 1, 0.f);  // + the value head's bias
        wvf.assign(HIDDEN + 1, 0.f);
        const std::vector<float>& h1 = n->host["hidden_fc.weight"];  // [256][flat]
        for (int kc = 0; kc < P.flat / 16; ++kc)
            for (int kh = 0; kh < 2; ++kh)
                for (int o = 0; o < HIDDEN; ++o)
                    for (int j = 0; j < 8; ++j) {
                        const size_t ui = ((((size_t)kc * 2 + kh) * HIDDEN + o) * 8) + j;
                        split_bf16(h1[(size_t)o * P.flat + kc * 16 + kh * 8 + j], w1u[ui], w1l[ui]);
                    }
        const std::vector<float>& l2 = n->host["logits_fc.weight"];  // [A][256]
        for (int kc = 0; kc < HIDDEN / 16; ++kc)
            for (int kh = 0; kh < 2; ++kh)
                for (int o = 0; o < P.A; ++o)
                    for (int j = 0; j < 8; ++j) {
                        const size_t ui = ((((size_t)kc * 2 + kh) * Hp.N2 + o) * 8) + j;
                        split_bf16(l2[(size_t)o * HIDDEN + kc * 16 + kh * 8 + j], w2u[ui], w2l[ui]);
                    }
        const std::vector<float>& v3 = n->host["value_fc.weight"];
        for (int i = 0; i < HIDDEN; ++i) {
            uint32_t hb = (uint32_t)f32_to_bf16_rne(v3[i]) << 16;
            memcpy(&wv32[i], &hb, 4);
            wvf[i] = v3[i];
        }
        wvf[HIDDEN] = n->host["value_fc.bias"][0];
        // in device memory, not in the by-value kernel parameter: captured CUDA graphs replay the parameters of their
        // capture, and the weights change between iterations
        wv32[HIDDEN] = n->host["value_fc.bias"][0];
        if (cudaMemcpyAsync(n->d_w1u, w1u.data(), w1u.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(n->d_w2u, w2u.data(), w2u.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(n->d_w1u_lo, w1l.data(), w1l.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(n->d_w2u_lo, w2l.data(), w2l.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(n->d_wvf, wvf.data(), wvf.size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(n->d_wv32, wv32.data(), wv32.size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess)
            return nerr(BPP_E_CUDA, "head parameter upload failed");
    }
    std::vector<uint16_t> wlp((size_t)HIDDEN * n->T.A_pad, 0);
    {
        const std::vector<float>& src2 = n->host["logits_fc.weight"];  // [A][256]
        for (int o = 0; o < P.A; ++o)
            for (int i = 0; i < HIDDEN; ++i) wlp[(size_t)i * n->T.A_pad + o] = f32_to_bf16_rne(src2[(size_t)o * HIDDEN + i]);
    }
    if ((n->gr_ok || n->gr3_ok) &&
        (cudaMemcpyAsync(n->d_wts_gr, wg.data(), wg.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess ||
         cudaMemcpyAsync(n->d_wts_gr_lo, wgl.data(), wgl.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess))
        return nerr(BPP_E_CUDA, "grid-row weight upload failed");
    if (cudaMemcpyAsync(n->d_wts_logits_pad, wlp.data(), wlp.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(n->d_wts_umma, wu.data(), wu.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(n->d_wts_umma_lo, wul.data(), wul.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(n->d_wts, w.data(), w.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(n->d_wts32, w32.data(), w32.size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(n->d_bias, b.data(), b.size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return nerr(BPP_E_CUDA, std::string("parameter upload failed: ") + cudaGetErrorString(cudaGetLastError()));
    n->committed = true;
    return BPP_OK;
}

extern "C" int bpp_net_profile(bpp_net* n, int64_t cycles_host[8]) {
    if (!n || !cycles_host || !n->d_prof) return nerr(BPP_E_INVALID, "null argument");
    long long all[32];
    if (cudaMemcpy(all, n->d_prof, sizeof(all), cudaMemcpyDeviceToHost) != cudaSuccess)
        return nerr(BPP_E_CUDA, "profile copy failed");
    // role kernels write one block of 8 timers each (rows 0..2): report their sum; the one-kernel trunk writes row 0
    const bool roles = (n->precision == BPP_NET_BF16 && n->roles_ok[0] && n->heads_ok && n->max_batch >= n->roles_min_batch) ||
                       (n->precision == BPP_NET_BF16X3 && n->roles_ok[1] && n->heads3_ok);
    for (int i = 0; i < 8; ++i) cycles_host[i] = roles ? all[i] + all[8 + i] + all[16 + i] : all[i];
    return BPP_OK;
}
extern "C" int bpp_net_grid_row(bpp_net* n) {
    return n && ((n->precision == BPP_NET_BF16 && n->gr_ok) || (n->precision == BPP_NET_BF16X3 && n->gr3_ok)) ? 1 : 0;
}
extern "C" int bpp_net_profile_roles(bpp_net* n, int64_t cycles_host[32]) {
    if (!n || !cycles_host || !n->d_prof) return nerr(BPP_E_INVALID, "null argument");
    if (cudaMemcpy(cycles_host, n->d_prof, 32 * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess)
        return nerr(BPP_E_CUDA, "profile copy failed");
    return BPP_OK;
}

extern "C" int bpp_net_set_precision(bpp_net* n, int mode) {
    if (!n || (mode != BPP_NET_BF16 && mode != BPP_NET_FP32 && mode != BPP_NET_BF16_SIMT && mode != BPP_NET_BF16X3))
        return nerr(BPP_E_INVALID, "unknown precision mode");
    n->precision = mode;
    return BPP_OK;
}

// launch with the programmatic-stream-serialization attribute (see pdl_wait in bpp_net_tc.cuh); BPP_NO_PDL=1 falls back
// to plain stream order
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args&&... args) {
    static const bool use_pdl = getenv("BPP_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// three role kernels (one per ConvSequence, each with its own group size) chained by programmatic launches
static int launch_roles(bpp_net* n, int x3, int B, const int32_t* count_dev, const uint32_t* recs_dev, const int32_t* game_dev,
                        const int32_t* items_wh_dev, cudaStream_t st) {
    const uint4* xin[3] = {nullptr, n->d_x1, n->d_x2};
    uint4* xout[3] = {n->d_x1, n->d_x2, nullptr};
    const long long flo = (long long)n->max_batch * n->P.flat;
    cudaError_t ce = cudaSuccess;
    for (int sq = 0; sq < 3 && ce == cudaSuccess; ++sq) {
        const bpptc::TcParams& R = n->Tr[x3][sq];
        const int c = n->role_ctas[x3][sq];
        const int gr = std::max(1, std::min((B + R.S - 1) / R.S, n->num_sms * c));
#define ROLE_LAUNCH(SQ, X, MB)                                                                                              \
    ce = launch_pdl(k_net_role<SQ, X, MB>, gr, bpptc::TC_THREADS, (size_t)R.smem_bytes, st, n->P, R, B, count_dev, recs_dev, \
                    game_dev, items_wh_dev, xin[SQ], xout[SQ], n->d_feat, flo, n->d_prof)
#define ROLE_PICK(SQ)                                                     \
    do {                                                                  \
        if (x3) { if (c == 2) ROLE_LAUNCH(SQ, true, 2); else ROLE_LAUNCH(SQ, true, 1); }    \
        else { if (c == 2) ROLE_LAUNCH(SQ, false, 2); else ROLE_LAUNCH(SQ, false, 1); }     \
    } while (0)
        if (sq == 0) ROLE_PICK(0);
        else if (sq == 1) ROLE_PICK(1);
        else ROLE_PICK(2);
#undef ROLE_PICK
#undef ROLE_LAUNCH
    }
    if (ce != cudaSuccess) return nerr(BPP_E_CUDA, std::string("role kernel launch failed: ") + cudaGetErrorString(ce));
    return BPP_OK;
}


// grid-row trunk: the four levels in one launch (bpp_net_gr.cuh); every CTA takes its share of the batch through all of them
static int launch_gr(bpp_net* n, int x3, int B, const int32_t* count_dev, const uint32_t* recs_dev, const int32_t* game_dev,
                     const int32_t* items_wh_dev, cudaStream_t st) {
    const bppgr::GrStage* G = x3 ? n->Gr3 : n->Gr;
    const long long flo = (long long)n->max_batch * n->P.flat;
    const int grid = std::max(1, std::min((B + 7) / 8, n->num_sms));
    cudaError_t ce;
    if (x3)
        ce = launch_pdl(bppgr::k_net_gr<true>, grid, n->gr_threads[1], (size_t)n->gr_smem[1], st, n->P, G[0], G[1], G[2], G[3], B,
                        count_dev, recs_dev, game_dev, items_wh_dev, n->d_g[0], n->d_g[1], n->d_g[2], n->d_feat, flo,
                        (const __nv_bfloat16*)n->d_wts_gr, (const __nv_bfloat16*)n->d_wts_gr_lo, n->d_prof);
    else
        ce = launch_pdl(bppgr::k_net_gr<false>, grid, n->gr_threads[0], (size_t)n->gr_smem[0], st, n->P, G[0], G[1], G[2], G[3], B,
                        count_dev, recs_dev, game_dev, items_wh_dev, n->d_g[0], n->d_g[1], n->d_g[2], n->d_feat, flo,
                        (const __nv_bfloat16*)n->d_wts_gr, (const __nv_bfloat16*)n->d_wts_gr_lo, n->d_prof);
    if (ce != cudaSuccess) return nerr(BPP_E_CUDA, std::string("grid-row kernel launch failed: ") + cudaGetErrorString(ce));
    return BPP_OK;
}

extern "C" int bpp_net_forward(bpp_net* n, int B, const int32_t* count_dev, const uint32_t* recs_dev,
                               const int32_t* game_dev, const int32_t* items_wh_dev, float* policy_out_dev,
                               float* value_out_dev, void* stream) {
    if (!n || !recs_dev || !items_wh_dev || !policy_out_dev || !value_out_dev)
        return B == 0 ? BPP_OK : nerr(BPP_E_INVALID, "null argument");
    if (!n->committed) return nerr(BPP_E_STATE, "bpp_net_forward before bpp_net_commit");
    if (B < 0 || B > n->max_batch) return nerr(BPP_E_INVALID, "batch larger than max_batch");
    if (B == 0) return BPP_OK;
    int grid = B < n->num_sms * 8 ? B : n->num_sms * 8;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (n->precision == BPP_NET_BF16 && n->tc_ok) {
        const int groups = (B + n->T.S - 1) / n->T.S;
        const int cap = n->num_sms * n->ctas_per_sm;
        const int g2 = groups < cap ? groups : cap;
        __nv_bfloat16* fo = n->heads_ok ? n->d_feat : nullptr;
        if (fo && n->gr_ok) {
            int rc = launch_gr(n, 0, B, count_dev, recs_dev, game_dev, items_wh_dev, st);
            if (rc) return rc;
        } else if (fo && n->roles_ok[0] && B >= n->roles_min_batch) {
            int rc = launch_roles(n, 0, B, count_dev, recs_dev, game_dev, items_wh_dev, st);
            if (rc) return rc;
        } else if (fo && n->ctas_per_sm == 2)
            launch_pdl(k_net_forward_tc<8, false, true, 2>, g2, bpptc::TC_THREADS, (size_t)n->T.smem_bytes, st, n->P, n->T, B,
                       count_dev, recs_dev, game_dev, items_wh_dev, policy_out_dev, value_out_dev, n->d_prof, fo);
        else if (n->T.S <= 4)
            k_net_forward_tc<4, false><<<g2, bpptc::TC_THREADS, n->T.smem_bytes, st>>>(
                n->P, n->T, B, count_dev, recs_dev, game_dev, items_wh_dev, policy_out_dev, value_out_dev, n->d_prof, fo);
        else
            k_net_forward_tc<8, false><<<g2, bpptc::TC_THREADS, n->T.smem_bytes, st>>>(
                n->P, n->T, B, count_dev, recs_dev, game_dev, items_wh_dev, policy_out_dev, value_out_dev, n->d_prof, fo);
        if (fo)
            launch_pdl(k_net_heads_tc<false>, (B + 127) / 128, HEAD_THREADS, (size_t)n->heads_smem, st, n->Hp, B, count_dev,
                       (const __nv_bfloat16*)n->d_feat, policy_out_dev, value_out_dev);
    } else if (n->precision == BPP_NET_BF16X3 && n->gr3_ok && n->heads3_ok) {
        int rc = launch_gr(n, 1, B, count_dev, recs_dev, game_dev, items_wh_dev, st);
        if (rc) return rc;
        launch_pdl(k_net_heads_tc<true>, (B + 127) / 128, HEAD_THREADS, (size_t)n->heads_smem3, st, n->Hp, B, count_dev,
                   (const __nv_bfloat16*)n->d_feat, policy_out_dev, value_out_dev);
    } else if (n->precision == BPP_NET_BF16X3 && n->roles_ok[1] && n->heads3_ok) {
        int rc = launch_roles(n, 1, B, count_dev, recs_dev, game_dev, items_wh_dev, st);
        if (rc) return rc;
        launch_pdl(k_net_heads_tc<true>, (B + 127) / 128, HEAD_THREADS, (size_t)n->heads_smem3, st, n->Hp, B, count_dev,
                   (const __nv_bfloat16*)n->d_feat, policy_out_dev, value_out_dev);
    } else if (n->precision == BPP_NET_BF16X3 && n->tc3_ok) {
        const int groups = (B + n->T3.S - 1) / n->T3.S;
        const int g2 = groups < n->num_sms ? groups : n->num_sms;
        if (n->T3.S <= 4)
            k_net_forward_tc<4, true><<<g2, bpptc::TC_THREADS, n->T3.smem_bytes, st>>>(
                n->P, n->T3, B, count_dev, recs_dev, game_dev, items_wh_dev, policy_out_dev, value_out_dev, n->d_prof,
                nullptr);
        else
            k_net_forward_tc<8, true><<<g2, bpptc::TC_THREADS, n->T3.smem_bytes, st>>>(
                n->P, n->T3, B, count_dev, recs_dev, game_dev, items_wh_dev, policy_out_dev, value_out_dev, n->d_prof,
                nullptr);
    } else if (n->precision == BPP_NET_FP32)
        k_net_forward<true><<<grid, NET_THREADS, n->smem_bytes, st>>>(n->P, B, count_dev, recs_dev, game_dev, items_wh_dev,
                                                                      policy_out_dev, value_out_dev);
    else
        k_net_forward<false><<<grid, NET_THREADS, n->smem_bytes, st>>>(n->P, B, count_dev, recs_dev, game_dev,
                                                                       items_wh_dev, policy_out_dev, value_out_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return nerr(BPP_E_CUDA, std::string("forward launch failed: ") + cudaGetErrorString(e));
    return BPP_OK;
}
