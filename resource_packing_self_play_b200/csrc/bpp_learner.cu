// Learner step of the policy/value network in fp32 on the CUDA cores: forward with stashed activations, losses, full
// backward, deterministic gradient reduction, Adam.  Follows NNetWrapper.train (NNet.py:27-67: Adam with default
// hyper-parameters, loss = loss_pi + loss_v, NNet.py:87-91) over BinPackingNNet.forward (BinpackingNNet.py:72-81),
// ConvSequence (:29-48: conv -> max_pool2d(3, 2, 1) -> two residual blocks) and ResidualBlock (:15-27, pre-activation).
//
// Why fp32 SIMT and not tcgen05: the gradients must reproduce the reference's fp32 training (the trained reference
// checkpoints reach logits of -3e3, which bf16 operands cannot represent, DESIGN.md "leaf evaluation"), and the net is
// 4.4 MFLOP per sample, so a minibatch step is ~7 GFLOP: launch- and latency-bound, not FLOP-bound.
//
// Launches per step: up to 128 samples one kernel per ConvSequence and direction (k_lr_stage_fwd / k_lr_stage_bwd: the
// running activation stays in shared memory across the stage's five layers; 13 launches), beyond that one kernel per layer
// (k_lr_conv / k_lr_pool_*; 41 launches) - the fused form measured slower from 256 samples on (profiles/r02_notes.md).
//
// Layout: activations fp32 [B][C][h*w] per tensor in one device buffer.  Parameters, gradients and Adam moments are
// flat caller-owned fp32 device buffers in state_dict order (bpp_learner_param_offset).  Per-chunk partial gradients
// [chunk][param] are summed in a fixed order (no atomics): the step is run-to-run deterministic.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bpp_b200.h"

int bpp_set_error_message(int code, const char* msg);  // defined in bpp_engine.cu

namespace {

constexpr int REC_WORDS = 32, REC_REM = 28;
constexpr int HIDDEN = 256;
constexpr int NSTAGE = 3, NCONV = 15;
constexpr int SMAX = 16;      // samples per gradient chunk (register arrays in the heads kernel)
constexpr int THREADS = 256;

int lerr(int code, const std::string& m) { return bpp_set_error_message(code, m.c_str()); }

// asynchronous global -> shared copies (LDGSTS): the staging loops issue them back to back instead of paying one
// L2 round trip per loop iteration
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src));
}
__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct ConvL {
    int cin, cout, h, w;
    long long w_off, b_off;     // offsets into the flat parameter vector
};

// ---------------------------------------------------------------------------------------------------------------------
// input planes (BinPackingGame.getBinItem layout: plane 0 = bin, plane 1+i = item i while it is still to be placed)
__global__ void k_lr_planes(int B, int W, int H, int N, const uint32_t* __restrict__ recs, const int32_t* __restrict__ items_wh,
                            const int64_t* __restrict__ ids, float* __restrict__ out) {
    const int hw = H * W, per = (N + 1) * hw;
    const long long total = (long long)B * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / per);
        int r = (int)(i - (long long)b * per);
        const int c = r / hw;
        r -= c * hw;
        const int y = r / W, x = r - y * W;
        const long long e = ids ? ids[b] : b;
        const uint32_t* rec = recs + (size_t)e * REC_WORDS;
        float v;
        if (c == 0) {
            v = (float)((rec[y] >> x) & 1u);
        } else {
            const int32_t* wh = items_wh + ((size_t)e * N + (c - 1)) * 2;
            v = (((rec[REC_REM] >> (c - 1)) & 1u) && y < wh[1] && x < wh[0]) ? 1.f : 0.f;
        }
        out[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// conv weights re-laid out once per step for coalesced staging: fwd[k = ci*9 + tap][co] = W[co][ci][tap] and, for the
// data gradient, bwd[k = co*9 + tap][ci] = W[co][ci][8 - tap] (transposed, rotated by 180 degrees); both at w_off.
struct RelayoutArgs {
    long long w_off[NCONV];
    int cin[NCONV], cout[NCONV];
};
__global__ void k_lr_relayout(RelayoutArgs a, const float* __restrict__ params, float* __restrict__ fwd,
                              float* __restrict__ bwd) {
    const int l = blockIdx.y, cin = a.cin[l], cout = a.cout[l], n = cin * cout * 9;
    const float* w = params + a.w_off[l];
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        {   // forward layout: idx = (ci*9 + tap)*cout + co
            const int k = idx / cout, co = idx - k * cout, ci = k / 9, tap = k - ci * 9;
            fwd[a.w_off[l] + idx] = w[((size_t)co * cin + ci) * 9 + tap];
        }
        {   // dgrad layout: idx = (co*9 + tap)*cin + ci
            const int k = idx / cin, ci = idx - k * cin, co = k / 9, tap = k - co * 9;
            bwd[a.w_off[l] + idx] = w[((size_t)co * cin + ci) * 9 + (8 - tap)];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// 3x3 / pad 1 convolution over [B][cin][h*w] -> [B][cout][h*w] (cout a multiple of 8), used for the forward and, with
// the transposed + rotated weights, for the data gradient.
//   out = (mask ? (mask > 0) : 1) * (bias + conv(relu_in ? relu(in) : in)) + (add ? add : 0)
// One CTA = G samples staged with zero halos in shared memory; one work item = a row segment of TP output positions x 8
// output channels: per (input channel, kernel row) TP + 2 inputs and 3 x 8 weights feed 24 * TP FMAs, which keeps the
// shared-memory pipe (128 B/clk) below the FMA pipe.
struct ConvArgs {
    const float* __restrict__ in;
    const float* __restrict__ wt;    // [cin*9][cout] (k_lr_relayout)
    const float* __restrict__ bias;  // nullable
    const float* __restrict__ mask;  // nullable, [B][cout][hw]
    const float* __restrict__ add;   // nullable, [B][cout][hw]
    float* __restrict__ out;
    int B, cin, cout, h, w, G, relu_in;
    int wp;             // padded row stride in shared memory (odd: rows of one column fall into distinct banks)
    int KS;             // input-channel split: KS threads share one work item (small layers: latency, not FMAs, bounds them)
};

// acc[j][k] += sum over ci in [ci0, ci1), 3x3 taps: in(position k + tap) * w[ci][tap][channel j]
template <int TP>
__device__ __forceinline__ void conv_tile(float (&acc)[8][TP], const float* sp, const float* wp0, int ci0, int ci1,
                                          int cstride, int wp, int cout, float lo) {
    for (int ci = ci0; ci < ci1; ++ci) {
#pragma unroll
        for (int ty = 0; ty < 3; ++ty) {
            float in[TP + 2];
#pragma unroll
            for (int k = 0; k < TP + 2; ++k) in[k] = fmaxf(sp[ci * cstride + ty * wp + k], lo);  // lo = 0: relu(in)
#pragma unroll
            for (int tx = 0; tx < 3; ++tx) {
                const float4* wv = reinterpret_cast<const float4*>(wp0 + (size_t)(ci * 9 + ty * 3 + tx) * cout);
                const float4 w0 = wv[0], w1 = wv[1];
                const float wj[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int k = 0; k < TP; ++k) acc[j][k] = fmaf(in[k + tx], wj[j], acc[j][k]);
            }
        }
    }
}

template <int TP>
__global__ void __launch_bounds__(THREADS, 2) k_lr_conv(ConvArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int hw = a.h * a.w, wp = a.wp, PP = (a.h + 2) * wp;
    float* s_w = sm;                                            // [cin*9][cout]
    float* s_in = sm + a.cin * 9 * a.cout;                      // [cin][G][PP] + 8 floats of slack (segment over-read)
    float* s_part = s_in + ((a.cin * a.G * PP + 8 + 3) & ~3);   // [8*TP][KS*items] (KS > 1 only)
    const int tid = threadIdx.x, b0 = blockIdx.x * a.G;
    {
        const int n4 = a.cin * 9 * a.cout / 4;
        for (int idx = tid; idx < n4; idx += THREADS) cp_async16(s_w + 4 * idx, a.wt + 4 * idx);
    }
    for (int idx = tid; idx < a.cin * a.G * PP + 8; idx += THREADS) s_in[idx] = 0.f;
    __syncthreads();
    const int per = a.cin * hw;
    {   // one (sample, channel) plane per warp and pass; y = pos / w by multiply-shift (exact for pos < 1024, w <= 32)
        const int lane = tid & 31, warp = tid >> 5;
        const unsigned magic = (65536u + a.w - 1) / a.w;
        const int gmax = min(a.G, a.B - b0);
        for (int pl = warp; pl < gmax * a.cin; pl += THREADS / 32) {
            const int g = pl / a.cin, ci = pl - g * a.cin;
            const float* src = a.in + (size_t)(b0 + g) * per + ci * hw;
            float* dst = s_in + (ci * a.G + g) * PP + wp + 1;
            for (int pos = lane; pos < hw; pos += 32) {
                const int y = (int)(((unsigned)pos * magic) >> 16);
                cp_async4(dst + pos + y * (wp - a.w), src + pos);
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();
    const float lo = a.relu_in ? 0.f : -INFINITY;
    const int nseg = (a.w + TP - 1) / TP;
    const int rows = a.G * a.h * nseg, items = rows * (a.cout >> 3);
    const int cstride = a.G * PP;
    if (a.KS == 1) {
        for (int item = tid; item < items; item += THREADS) {
            const int cb = item / rows;
            int r = item - cb * rows;
            const int g = r / (a.h * nseg);
            r -= g * a.h * nseg;
            const int y = r / nseg, x0 = (r - y * nseg) * TP;
            const int b = b0 + g;
            if (b >= a.B) continue;
            float acc[8][TP];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float bj = a.bias ? __ldg(a.bias + cb * 8 + j) : 0.f;
#pragma unroll
                for (int k = 0; k < TP; ++k) acc[j][k] = bj;
            }
            conv_tile<TP>(acc, s_in + g * PP + y * wp + x0, s_w + cb * 8, 0, a.cin, cstride, wp, a.cout, lo);
#pragma unroll
            for (int j0 = 0; j0 < 8; j0 += 2) {  // per channel pair: all loads of the epilogue first, then the stores
                float mk[2][TP], ad[2][TP];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const size_t o0 = ((size_t)b * a.cout + cb * 8 + j0 + j) * hw + y * a.w + x0;
#pragma unroll
                    for (int k = 0; k < TP; ++k) {
                        const bool in_row = x0 + k < a.w;
                        mk[j][k] = (a.mask && in_row) ? a.mask[o0 + k] : 1.f;
                        ad[j][k] = (a.add && in_row) ? a.add[o0 + k] : 0.f;
                    }
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const size_t o0 = ((size_t)b * a.cout + cb * 8 + j0 + j) * hw + y * a.w + x0;
#pragma unroll
                    for (int k = 0; k < TP; ++k)
                        if (x0 + k < a.w) a.out[o0 + k] = (mk[j][k] > 0.f ? acc[j0 + j][k] : 0.f) + ad[j][k];
                }
            }
        }
        return;
    }
    // split over the input channels: thread = (work item, channel slice); partial sums meet in shared memory and every
    // thread then reduces and writes a slice of its own item's 8 x TP outputs
    const int slots = items * a.KS;
    const bool live = tid < slots;
    const int item = live ? tid % items : 0, ks = live ? tid / items : 0;
    const int cb = item / rows;
    int r = item - cb * rows;
    const int g = r / (a.h * nseg);
    r -= g * a.h * nseg;
    const int y = r / nseg, x0 = (r - y * nseg) * TP;
    const int b = b0 + g;
    if (live) {
        const int cpk = (a.cin + a.KS - 1) / a.KS;
        float acc[8][TP];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int k = 0; k < TP; ++k) acc[j][k] = 0.f;
        conv_tile<TP>(acc, s_in + g * PP + y * wp + x0, s_w + cb * 8, ks * cpk, min(a.cin, (ks + 1) * cpk), cstride, wp,
                      a.cout, lo);
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int k = 0; k < TP; ++k) s_part[(j * TP + k) * slots + tid] = acc[j][k];
    }
    __syncthreads();
    if (!live || b >= a.B) return;
    const size_t obase = ((size_t)b * a.cout + cb * 8) * hw + y * a.w + x0;
    for (int jk = ks; jk < 8 * TP; jk += a.KS) {
        const int j = jk / TP, k = jk % TP;  // TP is a compile-time power of two
        if (x0 + k >= a.w) continue;
        const size_t o = obase + (size_t)j * hw + k;
        const float mk = a.mask ? a.mask[o] : 1.f;
        const float ad = a.add ? a.add[o] : 0.f;
        float v = a.bias ? __ldg(a.bias + cb * 8 + j) : 0.f;
        const float* sp = s_part + jk * slots + item;
        for (int q = 0; q < a.KS; ++q) v += sp[q * items];
        a.out[o] = (mk > 0.f ? v : 0.f) + ad;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// One launch per ConvSequence and direction: a CTA takes G samples through the stage's five convolutions and its pooling
// with the activations of the running layer in shared memory (zero halos), so that a layer costs its FMAs and one weight
// fetch instead of a launch, a staging pass over global memory and a tail.  Every tensor the backward pass or the weight
// gradient needs is still written to global memory (c, p, arg-max codes, a0, q, a1, o / da1, dq, da0, dp, dc, dout).
// The per-layer arithmetic is conv_tile's, with the same work decomposition as k_lr_conv (row segments of TP positions x 8
// output channels, input channels split over KS threads when a layer has fewer work items than the CTA has threads).
struct StageLayer {
    const float* wt;    // [cin*9][cout] (forward) or the transposed + rotated layout (data gradient)
    const float* bias;  // nullable
    const float* mask;  // nullable [B][cout][hw]: stashed activation whose sign gates the gradient
    const float* add;   // nullable [B][cout][hw]: residual / gradient of the skip connection
    float* out;         // [B][cout][hw]
    int cin, cout, h, w, wp, TP, KS, relu_in;
};
struct StageArgs {
    StageLayer L[5];    // forward: conv, res0.conv0, res0.conv1, res1.conv0, res1.conv1; backward: the reverse order
    const float* in;    // forward: the stage's input [B][cin][h0*w0]; backward: d(stage output) [B][ch][h1*w1]
    float* p;           // forward: pooled activation out           backward: unused
    uint8_t* amax;      // arg-max codes of the pooling (written forward, read backward)
    float* dc;          // backward: gradient w.r.t. the conv output (input resolution)
    int B, G, ch, h0, w0, h1, w1, wp0, wp1;
    int l0_dgrad;       // backward: the stage has a predecessor, its output gradient is L[4] (dgrad of the stage's conv)
    int off_a, off_x, off_y, off_part, n_act;   // float offsets into dynamic shared memory (weights at 0)
};

__device__ __forceinline__ void stage_weights(float* s_w, const StageLayer& L, int tid) {
    const int n4 = L.cin * 9 * L.cout / 4;
    for (int idx = tid; idx < n4; idx += THREADS) cp_async16(s_w + 4 * idx, L.wt + 4 * idx);
}

// global [B][C][h*w] planes of samples b0 .. b0 + gmax - 1 -> shared [C][G][PP] with halo (buffer zeroed before)
__device__ __forceinline__ void stage_planes(float* s_dst, const float* src_base, int C, int h, int w, int wp, int G, int gmax,
                                             int b0, int tid) {
    const int hw = h * w, PP = (h + 2) * wp, per = C * hw;
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned magic = (65536u + w - 1) / w;
    for (int pl = warp; pl < gmax * C; pl += THREADS / 32) {
        const int g = pl / C, ci = pl - g * C;
        const float* src = src_base + (size_t)(b0 + g) * per + ci * hw;
        float* dst = s_dst + (ci * G + g) * PP + wp + 1;
        for (int pos = lane; pos < hw; pos += 32) {
            const int y = (int)(((unsigned)pos * magic) >> 16);
            cp_async4(dst + pos + y * (wp - w), src + pos);
        }
    }
}

// one convolution of the stage: input in shared memory (s_in, [cin][G][PP]), weights in s_w; result to global memory and,
// when s_out != nullptr, into the interior of the next layer's shared input ([cout][G][PP], same resolution).
// Ends with a CTA barrier.
template <int TP>
__device__ __noinline__ void stage_conv(const StageLayer& a, const float* s_w, const float* s_in, float* s_out,
                                        float* s_part, int B, int G, int b0, int tid) {
    const int hw = a.h * a.w, wp = a.wp, PP = (a.h + 2) * wp;
    const float lo = a.relu_in ? 0.f : -INFINITY;
    const int nseg = (a.w + TP - 1) / TP;
    const int rows = G * a.h * nseg, items = rows * (a.cout >> 3);
    const int cstride = G * PP;
    if (a.KS == 1) {
        for (int item = tid; item < items; item += THREADS) {
            const int cb = item / rows;
            int r = item - cb * rows;
            const int g = r / (a.h * nseg);
            r -= g * a.h * nseg;
            const int y = r / nseg, x0 = (r - y * nseg) * TP;
            const int b = b0 + g;
            if (b >= B) continue;
            float acc[8][TP];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float bj = a.bias ? __ldg(a.bias + cb * 8 + j) : 0.f;
#pragma unroll
                for (int k = 0; k < TP; ++k) acc[j][k] = bj;
            }
            conv_tile<TP>(acc, s_in + g * PP + y * wp + x0, s_w + cb * 8, 0, a.cin, cstride, wp, a.cout, lo);
#pragma unroll
            for (int j0 = 0; j0 < 8; j0 += 2) {
                float mk[2][TP], ad[2][TP];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const size_t o0 = ((size_t)b * a.cout + cb * 8 + j0 + j) * hw + y * a.w + x0;
#pragma unroll
                    for (int k = 0; k < TP; ++k) {
                        const bool in_row = x0 + k < a.w;
                        mk[j][k] = (a.mask && in_row) ? a.mask[o0 + k] : 1.f;
                        ad[j][k] = (a.add && in_row) ? a.add[o0 + k] : 0.f;
                    }
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const size_t o0 = ((size_t)b * a.cout + cb * 8 + j0 + j) * hw + y * a.w + x0;
                    float* so = s_out ? s_out + ((cb * 8 + j0 + j) * G + g) * PP + (y + 1) * wp + x0 + 1 : nullptr;
#pragma unroll
                    for (int k = 0; k < TP; ++k)
                        if (x0 + k < a.w) {
                            const float v = (mk[j][k] > 0.f ? acc[j0 + j][k] : 0.f) + ad[j][k];
                            a.out[o0 + k] = v;
                            if (so) so[k] = v;
                        }
                }
            }
        }
        __syncthreads();
        return;
    }
    {
        const int slots = items * a.KS;
        const bool live = tid < slots;
        const int item = live ? tid % items : 0, ks = live ? tid / items : 0;
        const int cb = item / rows;
        int r = item - cb * rows;
        const int g = r / (a.h * nseg);
        r -= g * a.h * nseg;
        const int y = r / nseg, x0 = (r - y * nseg) * TP;
        const int b = b0 + g;
        if (live) {
            const int cpk = (a.cin + a.KS - 1) / a.KS;
            float acc[8][TP];
    #pragma unroll
            for (int j = 0; j < 8; ++j)
    #pragma unroll
                for (int k = 0; k < TP; ++k) acc[j][k] = 0.f;
            conv_tile<TP>(acc, s_in + g * PP + y * wp + x0, s_w + cb * 8, ks * cpk, min(a.cin, (ks + 1) * cpk), cstride, wp,
                          a.cout, lo);
    #pragma unroll
            for (int j = 0; j < 8; ++j)
    #pragma unroll
                for (int k = 0; k < TP; ++k) s_part[(j * TP + k) * slots + tid] = acc[j][k];
        }
        __syncthreads();
        if (live && b < B) {
            const size_t obase = ((size_t)b * a.cout + cb * 8) * hw + y * a.w + x0;
            for (int jk = ks; jk < 8 * TP; jk += a.KS) {
                const int j = jk / TP, k = jk % TP;  // TP is a compile-time power of two
                if (x0 + k >= a.w) continue;
                const size_t o = obase + (size_t)j * hw + k;
                const float mk = a.mask ? a.mask[o] : 1.f;
                const float ad = a.add ? a.add[o] : 0.f;
                float v = a.bias ? __ldg(a.bias + cb * 8 + j) : 0.f;
                const float* sp = s_part + jk * slots + item;
                for (int q = 0; q < a.KS; ++q) v += sp[q * items];
                v = (mk > 0.f ? v : 0.f) + ad;
                a.out[o] = v;
                if (s_out) s_out[((cb * 8 + j) * G + g) * PP + (y + 1) * wp + x0 + k + 1] = v;
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void stage_conv_any(const StageLayer& a, const float* s_w, const float* s_in, float* s_out,
                                               float* s_part, int B, int G, int b0, int tid) {
    if (a.TP == 8) stage_conv<8>(a, s_w, s_in, s_out, s_part, B, G, b0, tid);
    else if (a.TP == 4) stage_conv<4>(a, s_w, s_in, s_out, s_part, B, G, b0, tid);
    else stage_conv<2>(a, s_w, s_in, s_out, s_part, B, G, b0, tid);
}

__global__ void __launch_bounds__(THREADS, 2) k_lr_stage_fwd(StageArgs a) {
    extern __shared__ __align__(16) float sm[];
    float *s_w = sm, *s_a = sm + a.off_a, *s_x = sm + a.off_x, *s_y = sm + a.off_y, *s_part = sm + a.off_part;
    const int tid = threadIdx.x, b0 = blockIdx.x * a.G, gmax = min(a.G, a.B - b0);
    stage_weights(s_w, a.L[0], tid);
    for (int idx = tid; idx < a.n_act; idx += THREADS) s_a[idx] = 0.f;  // all activation buffers: the halos stay zero
    __syncthreads();
    stage_planes(s_a, a.in, a.L[0].cin, a.h0, a.w0, a.wp0, a.G, gmax, b0, tid);
    cp_async_wait_all();
    __syncthreads();
    stage_conv_any(a.L[0], s_w, s_a, nullptr, s_part, a.B, a.G, b0, tid);   // c (global), barrier
    stage_weights(s_w, a.L[1], tid);   // arrive behind the pooling
    {   // max_pool2d(3, 2, 1) with torch's arg-max rule (k_lr_pool_fwd) -> p, codes, and the first residual conv's input
        const int h = a.h0, w = a.w0, oh = a.h1, ow = a.w1, PP1 = (oh + 2) * a.wp1;
        const int per = a.ch * oh * ow;
        const float* cbuf = a.L[0].out;
        for (int i = tid; i < gmax * per; i += THREADS) {
            const int g = i / per;
            int r = i - g * per;
            const int c = r / (oh * ow);
            r -= c * oh * ow;
            const int oy = r / ow, ox = r - oy * ow;
            const float* pl = cbuf + ((size_t)(b0 + g) * a.ch + c) * h * w;
            float best = -INFINITY;
            int code = 0;
            bool first = true;
            for (int ky = 0; ky < 3; ++ky) {
                const int iy = 2 * oy - 1 + ky;
                if (iy < 0 || iy >= h) continue;
                for (int kx = 0; kx < 3; ++kx) {
                    const int ix = 2 * ox - 1 + kx;
                    if (ix < 0 || ix >= w) continue;
                    const float v = pl[iy * w + ix];
                    if (first || v > best || v != v) {
                        best = v;
                        code = ky * 3 + kx;
                        first = false;
                    }
                }
            }
            const size_t o = ((size_t)(b0 + g) * a.ch + c) * oh * ow + r;
            a.p[o] = best;
            a.amax[o] = (uint8_t)code;
            s_x[(c * a.G + g) * PP1 + (oy + 1) * a.wp1 + ox + 1] = best;
        }
    }
    cp_async_wait_all();
    __syncthreads();
    float *cur = s_x, *nxt = s_y;
    for (int k = 1; k < 5; ++k) {
        stage_conv_any(a.L[k], s_w, cur, k < 4 ? nxt : nullptr, s_part, a.B, a.G, b0, tid);
        if (k < 4) {
            stage_weights(s_w, a.L[k + 1], tid);
            cp_async_wait_all();
            __syncthreads();
        }
        float* t = cur; cur = nxt; nxt = t;
    }
}

__global__ void __launch_bounds__(THREADS, 2) k_lr_stage_bwd(StageArgs a) {
    extern __shared__ __align__(16) float sm[];
    float *s_w = sm, *s_a = sm + a.off_a, *s_x = sm + a.off_x, *s_y = sm + a.off_y, *s_part = sm + a.off_part;
    const int tid = threadIdx.x, b0 = blockIdx.x * a.G, gmax = min(a.G, a.B - b0);
    stage_weights(s_w, a.L[0], tid);
    for (int idx = tid; idx < a.n_act; idx += THREADS) s_a[idx] = 0.f;
    __syncthreads();
    stage_planes(s_x, a.in, a.ch, a.h1, a.w1, a.wp1, a.G, gmax, b0, tid);
    cp_async_wait_all();
    __syncthreads();
    float *cur = s_x, *nxt = s_y;
    for (int k = 0; k < 4; ++k) {   // da1, dq, da0, dp
        stage_conv_any(a.L[k], s_w, cur, k < 3 ? nxt : nullptr, s_part, a.B, a.G, b0, tid);
        if (k < 3 || a.l0_dgrad) {
            stage_weights(s_w, a.L[k + 1], tid);   // the conv's own data-gradient weights arrive behind the pooling
            if (k < 3) {
                cp_async_wait_all();
                __syncthreads();
            }
        }
        float* t = cur; cur = nxt; nxt = t;
    }
    {   // pooling backward (k_lr_pool_bwd): dp, codes -> dc (global) and the conv's data-gradient input
        const int h = a.h0, w = a.w0, oh = a.h1, ow = a.w1, PP0 = (h + 2) * a.wp0;
        const int per = a.ch * h * w;
        const float* dpb = a.L[3].out;
        for (int i = tid; i < gmax * per; i += THREADS) {
            const int g = i / per;
            int r = i - g * per;
            const int c = r / (h * w);
            r -= c * h * w;
            const int iy = r / w, ix = r - iy * w;
            const size_t pb = ((size_t)(b0 + g) * a.ch + c) * oh * ow;
            const float* dp = dpb + pb;
            const uint8_t* am = a.amax + pb;
            const int oy0 = iy >> 1, oy1 = (iy + 1) >> 1, ox0 = ix >> 1, ox1 = (ix + 1) >> 1;
            const bool vy1 = oy1 != oy0 && oy1 < oh, vx1 = ox1 != ox0 && ox1 < ow;
            const int cy1 = vy1 ? oy1 : oy0, cx1 = vx1 ? ox1 : ox0;
            const int i00 = oy0 * ow + ox0, i01 = oy0 * ow + cx1, i10 = cy1 * ow + ox0, i11 = cy1 * ow + cx1;
            const int a00 = am[i00], a01 = am[i01], a10 = am[i10], a11 = am[i11];
            const float d00 = dp[i00], d01 = dp[i01], d10 = dp[i10], d11 = dp[i11];
            const int ky0 = iy - (2 * oy0 - 1), ky1 = iy - (2 * cy1 - 1), kx0 = ix - (2 * ox0 - 1), kx1 = ix - (2 * cx1 - 1);
            float gsum = (a00 == ky0 * 3 + kx0) ? d00 : 0.f;
            if (vx1 && a01 == ky0 * 3 + kx1) gsum += d01;
            if (vy1 && a10 == ky1 * 3 + kx0) gsum += d10;
            if (vy1 && vx1 && a11 == ky1 * 3 + kx1) gsum += d11;
            a.dc[((size_t)(b0 + g) * a.ch + c) * h * w + r] = gsum;
            s_a[(c * a.G + g) * PP0 + (iy + 1) * a.wp0 + ix + 1] = gsum;
        }
    }
    if (!a.l0_dgrad) return;
    cp_async_wait_all();
    __syncthreads();
    stage_conv_any(a.L[4], s_w, s_a, nullptr, s_part, a.B, a.G, b0, tid);   // gradient of the previous stage's output
}

// ---------------------------------------------------------------------------------------------------------------------
// max_pool2d(kernel 3, stride 2, padding 1) with the arg-max rule of torch (first maximum in row-major window order)
__global__ void k_lr_pool_fwd(int BC, int h, int w, int oh, int ow, const float* __restrict__ in, float* __restrict__ out,
                              uint8_t* __restrict__ amax) {
    const long long total = (long long)BC * oh * ow;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int bc = (int)(i / (oh * ow));
        const int r = (int)(i - (long long)bc * oh * ow);
        const int oy = r / ow, ox = r - oy * ow;
        const float* p = in + (size_t)bc * h * w;
        float best = -INFINITY;
        int code = 0;
        bool first = true;
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = 2 * oy - 1 + ky;
            if (iy < 0 || iy >= h) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = 2 * ox - 1 + kx;
                if (ix < 0 || ix >= w) continue;
                const float v = p[iy * w + ix];
                if (first || v > best || v != v) {
                    best = v;
                    code = ky * 3 + kx;
                    first = false;
                }
            }
        }
        out[i] = best;
        amax[i] = (uint8_t)code;
    }
}

__global__ void k_lr_pool_bwd(int BC, int h, int w, int oh, int ow, const float* __restrict__ dout,
                              const uint8_t* __restrict__ amax, float* __restrict__ din) {
    const long long total = (long long)BC * h * w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int bc = (int)(i / (h * w));
        const int r = (int)(i - (long long)bc * h * w);
        const int iy = r / w, ix = r - iy * w;
        const float* dp = dout + (size_t)bc * oh * ow;
        const uint8_t* am = amax + (size_t)bc * oh * ow;
        // at most two candidate windows per axis; all loads first (clamped indices), then the selects
        const int oy0 = iy >> 1, oy1 = (iy + 1) >> 1, ox0 = ix >> 1, ox1 = (ix + 1) >> 1;
        const bool vy1 = oy1 != oy0 && oy1 < oh, vx1 = ox1 != ox0 && ox1 < ow;
        const int cy1 = vy1 ? oy1 : oy0, cx1 = vx1 ? ox1 : ox0;
        const int i00 = oy0 * ow + ox0, i01 = oy0 * ow + cx1, i10 = cy1 * ow + ox0, i11 = cy1 * ow + cx1;
        const int a00 = am[i00], a01 = am[i01], a10 = am[i10], a11 = am[i11];
        const float d00 = dp[i00], d01 = dp[i01], d10 = dp[i10], d11 = dp[i11];
        const int ky0 = iy - (2 * oy0 - 1), ky1 = iy - (2 * cy1 - 1), kx0 = ix - (2 * ox0 - 1), kx1 = ix - (2 * cx1 - 1);
        float g = (a00 == ky0 * 3 + kx0) ? d00 : 0.f;
        if (vx1 && a01 == ky0 * 3 + kx1) g += d01;
        if (vy1 && a10 == ky1 * 3 + kx0) g += d10;
        if (vy1 && vx1 && a11 == ky1 * 3 + kx1) g += d11;
        din[i] = g;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// heads: relu(flatten) -> hidden_fc -> relu -> {logits_fc -> log_softmax, value_fc -> tanh}, the two losses
// (NNet.py:87-91), their backward, and this chunk's partial gradients of the three linear layers.
// One CTA = one chunk of S <= SMAX samples.
struct HeadArgs {
    int B, S, flat, A;
    const float* o2;       // [B][flat] last stage output (pre-relu)
    const float* params;   // flat parameter vector
    long long w1, b1, w2, b2, wv, bv;
    const float* pis;      // [*][A] target policies
    const float* vs;       // [*]    target values
    const int64_t* ids;    // nullable gather index into pis / vs
    float* dfeat;          // [B][flat] gradient w.r.t. o2
    float* partial;        // [chunk][nparams rounded up to 4]
    long long nparams;     // row stride of `partial`
    float* loss_partial;   // [chunk][2]
    float* logp_out;       // nullable [B][A]: log-softmax output (evaluation)
    float* v_out;          // nullable [B]
};

template <int SM>  // SM = samples per chunk rounded up to a power of two (register arrays)
__global__ void __launch_bounds__(THREADS) k_lr_heads(HeadArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* s_f = sm;                        // [S][flat]   relu(features)
    float* s_h = s_f + SM * a.flat;       // [S][256]    hidden activations, later their gradients
    float* s_l = s_h + SM * HIDDEN;       // [S][A + 1]  logits -> dlogits; column A = dv (pre-tanh gradient)
    __shared__ float s_loss[2][THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = blockIdx.x * a.S;
    const int S = min(a.S, a.B - b0);
    const int A1 = a.A + 1;
    const float invB = 1.f / (float)a.B;
    const float* W1 = a.params + a.w1;
    const float* W2 = a.params + a.w2;
    const float* WV = a.params + a.wv;
    float* part = a.partial + (size_t)blockIdx.x * a.nparams;

    for (int idx = tid; idx < SM * a.flat; idx += THREADS) {
        const int s = idx / a.flat, i = idx - s * a.flat;
        s_f[idx] = s < S ? fmaxf(a.o2[(size_t)(b0 + s) * a.flat + i], 0.f) : 0.f;
    }
    __syncthreads();
    // hidden = relu(W1 f + b1): thread = hidden unit
    float hreg[SM];
    {
        const int o = tid;
        float acc[SM];
        const float bias = a.params[a.b1 + o];
#pragma unroll
        for (int s = 0; s < SM; ++s) acc[s] = bias;
        const float4* wrow = reinterpret_cast<const float4*>(W1 + (size_t)o * a.flat);
#pragma unroll 8
        for (int i4 = 0; i4 < a.flat / 4; ++i4) {
            const float4 w4 = __ldg(wrow + i4);
#pragma unroll
            for (int s = 0; s < SM; ++s) {
                const float4 f4 = *reinterpret_cast<const float4*>(s_f + s * a.flat + 4 * i4);
                acc[s] = fmaf(w4.x, f4.x, acc[s]);
                acc[s] = fmaf(w4.y, f4.y, acc[s]);
                acc[s] = fmaf(w4.z, f4.z, acc[s]);
                acc[s] = fmaf(w4.w, f4.w, acc[s]);
            }
        }
#pragma unroll
        for (int s = 0; s < SM; ++s) {
            hreg[s] = fmaxf(acc[s], 0.f);
            s_h[s * HIDDEN + o] = hreg[s];
        }
    }
    __syncthreads();
    // logits (threads j < A) and the value pre-activation (j == A)
    for (int j = tid; j <= a.A; j += THREADS) {
        float acc[SM];
        const float bias = a.params[(j < a.A ? a.b2 : a.bv) + (j < a.A ? j : 0)];
#pragma unroll
        for (int s = 0; s < SM; ++s) acc[s] = bias;
        const float* wrow = j < a.A ? W2 + (size_t)j * HIDDEN : WV;  // WV may not be 16-byte aligned (A odd multiples)
#pragma unroll 8
        for (int o4 = 0; o4 < HIDDEN / 4; ++o4) {
            const float4 w4 = j < a.A ? __ldg(reinterpret_cast<const float4*>(wrow) + o4)
                                      : make_float4(__ldg(wrow + 4 * o4), __ldg(wrow + 4 * o4 + 1), __ldg(wrow + 4 * o4 + 2),
                                                    __ldg(wrow + 4 * o4 + 3));
#pragma unroll
            for (int s = 0; s < SM; ++s) {
                const float4 h4 = *reinterpret_cast<const float4*>(s_h + s * HIDDEN + 4 * o4);
                acc[s] = fmaf(w4.x, h4.x, acc[s]);
                acc[s] = fmaf(w4.y, h4.y, acc[s]);
                acc[s] = fmaf(w4.z, h4.z, acc[s]);
                acc[s] = fmaf(w4.w, h4.w, acc[s]);
            }
        }
#pragma unroll
        for (int s = 0; s < SM; ++s) s_l[s * A1 + j] = acc[s];
    }
    __syncthreads();
    // per sample: log-softmax, losses, gradient of the logits and of the value pre-activation (one warp per sample)
    float lpi = 0.f, lv = 0.f;
    for (int s = warp; s < S; s += THREADS / 32) {
        const long long e = a.ids ? a.ids[b0 + s] : (b0 + s);
        const float* pi = a.pis + (size_t)e * a.A;
        float* row = s_l + s * A1;
        float mx = -INFINITY;
        for (int j = lane; j < a.A; j += 32) mx = fmaxf(mx, row[j]);
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f, psum = 0.f;
        for (int j = lane; j < a.A; j += 32) {
            sum += expf(row[j] - mx);
            psum += pi[j];
        }
        for (int o = 16; o; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            psum += __shfl_xor_sync(0xffffffffu, psum, o);
        }
        const float lse = mx + logf(sum);
        float dot = 0.f;
        for (int j = lane; j < a.A; j += 32) {
            const float lp = row[j] - lse, t = pi[j];
            dot += t * lp;
            if (a.logp_out) a.logp_out[(size_t)(b0 + s) * a.A + j] = lp;
            row[j] = (expf(lp) * psum - t) * invB;  // d(-sum(t * logp) / B) / dlogit
        }
        for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (lane == 0) {
            const float v = tanhf(row[a.A]), z = a.vs[e];
            if (a.v_out) a.v_out[b0 + s] = v;
            lpi -= dot;
            lv += (z - v) * (z - v);
            row[a.A] = 2.f * (v - z) * invB * (1.f - v * v);
        }
    }
    if (lane == 0) {
        s_loss[0][warp] = lpi;
        s_loss[1][warp] = lv;
    }
    for (int idx = tid; idx < (SM - S) * A1; idx += THREADS) s_l[S * A1 + idx] = 0.f;  // rows of absent samples
    __syncthreads();
    if (tid < 2) {
        float t = 0.f;
        for (int i = 0; i < THREADS / 32; ++i) t += s_loss[tid][i];
        a.loss_partial[blockIdx.x * 2 + tid] = t * invB;
    }
    if (!a.partial) return;  // evaluation only
    // partial gradients of logits_fc / value_fc: thread = hidden unit o (coalesced over o)
    {
        const int o = tid;
#pragma unroll 8
        for (int j = 0; j <= a.A; ++j) {
            float g = 0.f;
#pragma unroll
            for (int s = 0; s < SM; ++s) g = fmaf(s_l[s * A1 + j], hreg[s], g);
            part[(j < a.A ? a.w2 + (long long)j * HIDDEN : a.wv) + o] = g;
        }
    }
    for (int j = tid; j <= a.A; j += THREADS) {
        float g = 0.f;
        for (int s = 0; s < SM; ++s) g += s_l[s * A1 + j];
        part[j < a.A ? a.b2 + j : a.bv] = g;
    }
    // gradient of the hidden activations: dh[s][o] = (sum_j dl[s][j] W2[j][o] + dv[s] wv[o]) * (h > 0)
    float dh[SM];
    {
        const int o = tid;
#pragma unroll
        for (int s = 0; s < SM; ++s) dh[s] = 0.f;
#pragma unroll 8
        for (int j = 0; j <= a.A; ++j) {
            const float w = j < a.A ? __ldg(W2 + (size_t)j * HIDDEN + o) : __ldg(WV + o);
#pragma unroll
            for (int s = 0; s < SM; ++s) dh[s] = fmaf(s_l[s * A1 + j], w, dh[s]);
        }
        float gb = 0.f;
        __syncthreads();  // every thread is done reading s_h through hreg's producers; s_h is reused for dh
#pragma unroll
        for (int s = 0; s < SM; ++s) {
            dh[s] = hreg[s] > 0.f ? dh[s] : 0.f;
            s_h[s * HIDDEN + o] = dh[s];
            gb += dh[s];
        }
        part[a.b1 + o] = gb;
    }
    __syncthreads();
    // partial gradient of hidden_fc.weight [256][flat] (coalesced over the flattened index)
    for (int idx = tid; idx < HIDDEN * a.flat; idx += THREADS) {
        const int o = idx / a.flat, i = idx - o * a.flat;
        float g = 0.f;
#pragma unroll
        for (int s = 0; s < SM; ++s) g = fmaf(s_h[s * HIDDEN + o], s_f[s * a.flat + i], g);
        part[a.w1 + idx] = g;
    }
    // gradient w.r.t. the last stage output: (W1^T dh) * (o2 > 0)
    for (int i = tid; i < a.flat; i += THREADS) {
        float acc[SM];
#pragma unroll
        for (int s = 0; s < SM; ++s) acc[s] = 0.f;
#pragma unroll 8
        for (int o = 0; o < HIDDEN; ++o) {
            const float w = __ldg(W1 + (size_t)o * a.flat + i);
#pragma unroll
            for (int s = 0; s < SM; ++s) acc[s] = fmaf(s_h[s * HIDDEN + o], w, acc[s]);
        }
        for (int s = 0; s < S; ++s) a.dfeat[(size_t)(b0 + s) * a.flat + i] = s_f[s * a.flat + i] > 0.f ? acc[s] : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// weight / bias gradients of all 15 convolutions in one launch: grid = (chunks, layers).
//   dW[co][ci][t] = sum_{b, y, x} dY[b][co][y][x] * X[b][ci][y + ty - 1][x + tx - 1]     (X = relu(in) where the layer has it)
// thread = (ci, 4 output channels, sample subgroup): 36 accumulators; subgroups split the chunk's samples.
struct WgradLayer {
    const float* in;
    const float* dy;
    int cin, cout, h, w, relu_in;
    int pg;      // sample subgroups working in parallel (threads = pairs * pg)
    int round;   // samples staged in shared memory at once
    long long w_off, b_off;
};
struct WgradArgs {
    WgradLayer L[NCONV];
    int B, S;
    float* partial;
    long long nparams;
};

#define WG_FMA(L0, L1, L2, M0, M1, M2, R0, R1, R2, D)                        \
    {                                                                        \
        const float win[9] = {L0, M0, R0, L1, M1, R1, L2, M2, R2};           \
        const float4 d4 = (D);                                               \
        _Pragma("unroll") for (int t = 0; t < 9; ++t) {                      \
            acc[0][t] = fmaf(d4.x, win[t], acc[0][t]);                       \
            acc[1][t] = fmaf(d4.y, win[t], acc[1][t]);                       \
            acc[2][t] = fmaf(d4.z, win[t], acc[2][t]);                       \
            acc[3][t] = fmaf(d4.w, win[t], acc[3][t]);                       \
        }                                                                    \
        accb[0] += d4.x;                                                     \
        accb[1] += d4.y;                                                     \
        accb[2] += d4.z;                                                     \
        accb[3] += d4.w;                                                     \
    }

__global__ void __launch_bounds__(THREADS) k_lr_wgrad(WgradArgs a) {
    extern __shared__ __align__(16) float sm[];
    const WgradLayer& L = a.L[blockIdx.y];
    const int cin = L.cin, cout = L.cout, h = L.h, w = L.w, hw = h * w, wp = w + 2;
    const int PPs = ((h + 2) * wp) | 1;  // odd plane stride: lanes (consecutive ci) hit distinct banks
    const int cq4 = cout >> 2;
    const int pairs = cin * cq4, PG = L.pg, R = L.round;
    float* s_x = sm;                                 // [R][cin][PPs]   inputs with zero halo
    float* s_dy = sm + ((R * cin * PPs + 3) & ~3);   // [R][cout/4][hw][4]  output gradients, channel quads innermost
    const int tid = threadIdx.x;
    const int pair = tid % pairs, sg = tid / pairs;
    const int ci = pair % cin, cq = pair / cin;
    const bool active = sg < PG;
    const int b0 = blockIdx.x * a.S;
    const int S = min(a.S, a.B - b0);
    const float lo = L.relu_in ? 0.f : -INFINITY;
    float acc[4][9];
    float accb[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        accb[c] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[c][t] = 0.f;
    }
    for (int idx = tid; idx < R * cin * PPs; idx += THREADS) s_x[idx] = 0.f;  // halos stay zero
    for (int j0 = 0; j0 < S; j0 += R) {
        __syncthreads();
        const int ns = min(R, S - j0);
        const int perx = cin * hw, pery = cout * hw;
        const int lane = tid & 31, warp = tid >> 5;
        const unsigned magic = (65536u + w - 1) / w;
        const unsigned mcin = (65536u + cin - 1) / cin, mcq = (65536u + cq4 - 1) / cq4;  // exact for operands < 2048
        for (int pl = warp; pl < ns * cin; pl += THREADS / 32) {  // one input plane per warp and pass
            const int g = (int)(((unsigned)pl * mcin) >> 16), c = pl - g * cin;
            const float* src = L.in + (size_t)(b0 + j0 + g) * perx + c * hw;
            float* dst = s_x + (g * cin + c) * PPs + wp + 1;
            for (int pos = lane; pos < hw; pos += 32) {
                const int y = (int)(((unsigned)pos * magic) >> 16);
                cp_async4(dst + pos + 2 * y, src + pos);
            }
        }
        for (int gq = warp; gq < ns * cq4; gq += THREADS / 32) {  // one (sample, channel quad) per warp and pass
            const int g = (int)(((unsigned)gq * mcq) >> 16), q = gq - g * cq4;
            const float* src = L.dy + (size_t)(b0 + j0 + g) * pery + (size_t)q * 4 * hw;
            float* dst = s_dy + (size_t)gq * hw * 4;
            for (int r = lane; r < ((hw + 7) >> 3) * 32; r += 32) {  // r = (8-position group, channel in quad, position)
                const int grp = r >> 5, pos = (grp << 3) + (r & 7), ce = (r >> 3) & 3;  // 4 runs of 8 floats per warp load
                if (pos < hw) cp_async4(dst + pos * 4 + ce, src + ce * hw + pos);
            }
        }
        cp_async_wait_all();
        __syncthreads();
        if (active) {
            const unsigned mh = (65536u + h - 1) / h;
            for (int u = sg; u < ns * h; u += PG) {  // unit = one row of one sample; the 3x3 window slides along it
                const int g = (int)(((unsigned)u * mh) >> 16), y = u - g * h;
                const float* r0 = s_x + (g * cin + ci) * PPs + y * wp;
                const float *r1 = r0 + wp, *r2 = r1 + wp;
                const float4* dr = reinterpret_cast<const float4*>(s_dy) + (size_t)(g * cq4 + cq) * hw + y * w;
#define WG_LD(p, i) fmaxf((p)[i], lo)  /* lo = 0: the layer's input is relu(in) */
                float a0 = WG_LD(r0, 0), a1 = WG_LD(r1, 0), a2 = WG_LD(r2, 0);
                float b0v = WG_LD(r0, 1), b1v = WG_LD(r1, 1), b2v = WG_LD(r2, 1), c0, c1, c2;
                for (int x = 0; x < w; x += 3) {
                    c0 = WG_LD(r0, x + 2); c1 = WG_LD(r1, x + 2); c2 = WG_LD(r2, x + 2);
                    WG_FMA(a0, a1, a2, b0v, b1v, b2v, c0, c1, c2, dr[x]);
                    if (x + 1 < w) {
                        a0 = WG_LD(r0, x + 3); a1 = WG_LD(r1, x + 3); a2 = WG_LD(r2, x + 3);
                        WG_FMA(b0v, b1v, b2v, c0, c1, c2, a0, a1, a2, dr[x + 1]);
                    }
                    if (x + 2 < w) {
                        b0v = WG_LD(r0, x + 4); b1v = WG_LD(r1, x + 4); b2v = WG_LD(r2, x + 4);
                        WG_FMA(c0, c1, c2, a0, a1, a2, b0v, b1v, b2v, dr[x + 2]);
                    }
                }
            }
        }
    }
    __syncthreads();
    // reduce the subgroups through shared memory (fixed order: deterministic), lay the sums out in parameter order and
    // write this chunk's partial gradient with coalesced vector stores
    float* s_red = sm;                                     // [PG][pairs][41] (odd stride)
    float* s_out = sm + ((PG * pairs * 41 + 3) & ~3);      // [cout*cin*9] weights | [cout] bias
    if (active) {
        float* r = s_red + ((size_t)sg * pairs + pair) * 41;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int t = 0; t < 9; ++t) r[c * 9 + t] = acc[c][t];
            r[36 + c] = accb[c];
        }
    }
    __syncthreads();
    const int nw = cout * cin * 9;
    if (tid < pairs) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float* o = s_out + ((size_t)(cq * 4 + c) * cin + ci) * 9;
            const float* rp = s_red + (size_t)pair * 41 + c * 9;
            const int sstride = pairs * 41;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                float g = 0.f;
                for (int s2 = 0; s2 < PG; ++s2) g += rp[s2 * sstride + t];
                o[t] = g;
            }
            if (ci == 0) {
                float g = 0.f;
                for (int s2 = 0; s2 < PG; ++s2) g += s_red[(size_t)s2 * sstride + pair * 41 + 36 + c];
                s_out[nw + cq * 4 + c] = g;
            }
        }
    }
    __syncthreads();
    float* part = a.partial + (size_t)blockIdx.x * a.nparams;
    for (int i4 = tid; i4 < nw / 4; i4 += THREADS)
        reinterpret_cast<float4*>(part + L.w_off)[i4] = reinterpret_cast<const float4*>(s_out)[i4];
    for (int co = tid; co < cout; co += THREADS) part[L.b_off + co] = s_out[nw + co];
}

// grads[p] = sum over chunks (fixed order); losses[0..1] likewise
__global__ void k_lr_reduce(long long nparams, long long pstride, int nchunk, const float* __restrict__ partial,
                            float* __restrict__ grads,
                            const float* __restrict__ loss_partial, float* __restrict__ losses) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < nparams) {
        float g = 0.f;
        for (int k = 0; k < nchunk; ++k) g += partial[(size_t)k * pstride + i];
        grads[i] = g;
    }
    if (i < 2 && losses) {
        float t = 0.f;
        for (int k = 0; k < nchunk; ++k) t += loss_partial[k * 2 + i];
        losses[i] = t;
    }
}

// torch.optim.Adam (no weight decay, no amsgrad): the update of torch/optim/adam.py::_single_tensor_adam
__global__ void k_lr_adam(long long n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                          float* __restrict__ v, float grad_scale, float lr, float beta1, float beta2, float eps,
                          float bc1, float bc2_sqrt, const int32_t* __restrict__ step_dev) {
    __shared__ float s_bc[2];
    if (step_dev) {  // step count in device memory (CUDA-graph replays): bias corrections computed here
        if (threadIdx.x == 0) {
            const double t = (double)*step_dev;
            s_bc[0] = (float)(1.0 - pow((double)beta1, t));
            s_bc[1] = (float)sqrt(1.0 - pow((double)beta2, t));
        }
        __syncthreads();
        bc1 = s_bc[0];
        bc2_sqrt = s_bc[1];
    }
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i] * grad_scale;
    const float mi = m[i] + (gi - m[i]) * (1.f - beta1);            // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;        // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
struct bpp_learner {
    int W, H, N, A, Cin, max_batch, device;
    int hs[4], ws[4], chans[3];
    int flat;
    ConvL conv[NCONV];
    long long w1, b1, w2, b2, wv, bv, nparams;
    std::vector<std::pair<std::string, std::pair<long long, long long>>> names;  // state_dict order
    // activation tensors (device); per stage: c (input resolution), p, a0, q, a1, o (pooled resolution)
    float* planes = nullptr;
    float *c[3] = {}, *p[3] = {}, *a0[3] = {}, *q[3] = {}, *a1[3] = {}, *o[3] = {};
    float *dc[3] = {}, *dp[3] = {}, *da0[3] = {}, *dq[3] = {}, *da1[3] = {}, *dout[3] = {};
    uint8_t* amax[3] = {};
    float* wfwd = nullptr;   // conv weights as [cin*9][cout] (forward) and transposed + rotated (data gradient)
    float* wbwd = nullptr;
    float* partial = nullptr;
    float* loss_partial = nullptr;
    int max_chunks = 0;
    int sms = 148;
    int smem_cap = 0;
    int fused_max_batch = 128;   // one kernel per stage and direction up to this batch size, one per layer beyond
    std::vector<void*> allocs;
};

namespace {

bool dalloc(bpp_learner* l, void** p, size_t bytes) {
    if (cudaMalloc(p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    l->allocs.push_back(*p);
    return true;
}

// samples per gradient chunk: about one chunk per SM, at most SMAX samples each
void chunking(int B, int sms, int& S, int& nchunk) {
    S = std::min(SMAX, (B + sms - 1) / sms);
    nchunk = (B + S - 1) / S;
}

int launch_conv(bpp_learner* l, cudaStream_t st, int B, const float* in, const float* wt, const float* bias,
                const float* mask, const float* add, float* out, int cin, int cout, int h, int w, int relu_in) {
    ConvArgs a;
    a.in = in; a.wt = wt; a.bias = bias; a.mask = mask; a.add = add; a.out = out;
    a.B = B; a.cin = cin; a.cout = cout; a.h = h; a.w = w; a.relu_in = relu_in;
    const int TP = w >= 7 ? 8 : (w >= 3 ? 4 : 2);  // output positions per work item (row segment)
    a.wp = (w + 2) | 1;
    const int PP = (h + 2) * a.wp;
    const size_t wbytes = (size_t)cin * 9 * cout * 4;
    const size_t part_bytes = (size_t)THREADS * 8 * TP * 4;
    auto bytes = [&](int G) { return wbytes + ((size_t)cin * G * PP + 12) * 4 + part_bytes; };
    int G = std::max(1, (B + 2 * l->sms - 1) / (2 * l->sms));  // two CTAs per SM when the batch allows
    while (G > 1 && bytes(G) > (size_t)l->smem_cap) --G;
    if (bytes(G) > (size_t)l->smem_cap) return lerr(BPP_E_INVALID, "convolution does not fit in shared memory");
    a.G = G;
    const int items = G * h * ((w + TP - 1) / TP) * (cout / 8);
    a.KS = items >= THREADS ? 1 : std::max(1, std::min(cin, THREADS / items));
    const int grid = (B + G - 1) / G;
    const size_t smem = bytes(G);
    if (TP == 8) k_lr_conv<8><<<grid, THREADS, smem, st>>>(a);
    else if (TP == 4) k_lr_conv<4><<<grid, THREADS, smem, st>>>(a);
    else k_lr_conv<2><<<grid, THREADS, smem, st>>>(a);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return lerr(BPP_E_CUDA, std::string("convolution launch failed: ") + cudaGetErrorString(e) + " (smem " +
                                    std::to_string(smem) + ", G " + std::to_string(G) + ", TP " + std::to_string(TP) + ")");
    return BPP_OK;
}

// plan and launch one fused stage kernel (k_lr_stage_fwd / k_lr_stage_bwd); L[] = the five layers in execution order with
// their global pointers filled in, geometry taken from the stage
int launch_stage(bpp_learner* l, cudaStream_t st, int B, int s, bool fwd, StageArgs& a) {
    a.B = B;
    a.ch = l->chans[s];
    a.h0 = l->hs[s]; a.w0 = l->ws[s]; a.h1 = l->hs[s + 1]; a.w1 = l->ws[s + 1];
    a.wp0 = (a.w0 + 2) | 1; a.wp1 = (a.w1 + 2) | 1;
    const int PP0 = (a.h0 + 2) * a.wp0, PP1 = (a.h1 + 2) * a.wp1;
    const int nl = fwd ? 5 : (a.l0_dgrad ? 5 : 4);
    size_t wfl = 0;
    for (int k = 0; k < nl; ++k) {
        a.L[k].wp = (a.L[k].w + 2) | 1;
        wfl = std::max(wfl, (size_t)a.L[k].cin * 9 * a.L[k].cout);
    }
    const int cin_a = fwd ? a.L[0].cin : a.ch;   // channels of the input-resolution buffer
    auto r4 = [](size_t n) { return (n + 3) & ~(size_t)3; };
    auto plan = [&](int G, bool commit) {
        const size_t na = r4((size_t)cin_a * G * PP0 + 8), nx = r4((size_t)a.ch * G * PP1 + 8);
        size_t npart = 0;
        for (int k = 0; k < nl; ++k) {
            StageLayer& L = a.L[k];
            auto items = [&](int TP) { return G * L.h * ((L.w + TP - 1) / TP) * (L.cout / 8); };
            const int TP = L.w >= 7 ? 8 : (L.w >= 3 ? 4 : 2);   // output positions per work item (row segment), as k_lr_conv
            const int KS = items(TP) >= THREADS ? 1 : std::max(1, std::min(L.cin, THREADS / items(TP)));
            if (commit) {
                L.TP = TP;
                L.KS = KS;
            }
            if (KS > 1) npart = std::max(npart, (size_t)THREADS * 8 * TP);
        }
        if (commit) {
            a.G = G;
            a.off_a = (int)r4(wfl);
            a.off_x = a.off_a + (int)na;
            a.off_y = a.off_x + (int)nx;
            a.off_part = a.off_y + (int)nx;
            a.n_act = (int)(na + 2 * nx);
        }
        return (r4(wfl) + na + 2 * nx + npart) * 4;
    };
    int G = std::max(1, std::min(4, (B + 2 * l->sms - 1) / (2 * l->sms)));
    while (G > 1 && plan(G, false) > (size_t)l->smem_cap) --G;
    const size_t smem = plan(G, true);
    if (smem > (size_t)l->smem_cap) return lerr(BPP_E_INVALID, "stage does not fit in shared memory");
    const int grid = (B + G - 1) / G;
    if (fwd) k_lr_stage_fwd<<<grid, THREADS, smem, st>>>(a);
    else k_lr_stage_bwd<<<grid, THREADS, smem, st>>>(a);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return lerr(BPP_E_CUDA, std::string("stage kernel launch failed: ") + cudaGetErrorString(e) + " (smem " +
                                    std::to_string(smem) + ", G " + std::to_string(G) + ")");
    return BPP_OK;
}

void stage_layer(StageLayer& L, const float* wt, const float* bias, const float* mask, const float* add, float* out, int cin,
                 int cout, int h, int w, int relu_in) {
    L.wt = wt; L.bias = bias; L.mask = mask; L.add = add; L.out = out;
    L.cin = cin; L.cout = cout; L.h = h; L.w = w; L.relu_in = relu_in;
    L.wp = 0; L.TP = 0; L.KS = 1;
}

}  // namespace

extern "C" int bpp_learner_create(int W, int H, int N, int max_batch, int device, bpp_learner** out) {
    if (!out) return lerr(BPP_E_INVALID, "null argument");
    if (W < 1 || W > 32 || H < 1 || H > 28 || N < 1 || N > 16 || max_batch < 1)
        return lerr(BPP_E_INVALID, "unsupported network geometry");
    if (cudaSetDevice(device) != cudaSuccess) return lerr(BPP_E_CUDA, "cudaSetDevice failed");
    bpp_learner* l = new bpp_learner();
    l->W = W; l->H = H; l->N = N; l->A = W * N; l->Cin = N + 1; l->max_batch = max_batch; l->device = device;
    l->chans[0] = 16; l->chans[1] = 32; l->chans[2] = 32;
    l->hs[0] = H; l->ws[0] = W;
    for (int s = 0; s < 3; ++s) {
        l->hs[s + 1] = (l->hs[s] + 1) / 2;
        l->ws[s + 1] = (l->ws[s] + 1) / 2;
    }
    l->flat = 32 * l->hs[3] * l->ws[3];
    if (l->flat % 4) { delete l; return lerr(BPP_E_INVALID, "unsupported network geometry"); }
    long long off = 0;
    auto reg = [&](const std::string& name, long long numel) {
        l->names.push_back({name, {off, numel}});
        const long long o = off;
        off += numel;
        return o;
    };
    int cin = l->Cin;
    for (int s = 0; s < NSTAGE; ++s) {
        const int ch = l->chans[s];
        const std::string pre = "conv_seqs." + std::to_string(s) + ".";
        const char* sub[5] = {"conv", "res_block0.conv0", "res_block0.conv1", "res_block1.conv0", "res_block1.conv1"};
        for (int k = 0; k < 5; ++k) {
            ConvL& L = l->conv[s * 5 + k];
            L.cin = k == 0 ? cin : ch;
            L.cout = ch;
            L.h = k == 0 ? l->hs[s] : l->hs[s + 1];
            L.w = k == 0 ? l->ws[s] : l->ws[s + 1];
            L.w_off = reg(pre + sub[k] + ".weight", (long long)L.cout * L.cin * 9);
            L.b_off = reg(pre + sub[k] + ".bias", L.cout);
        }
        cin = ch;
    }
    l->w1 = reg("hidden_fc.weight", (long long)HIDDEN * l->flat);
    l->b1 = reg("hidden_fc.bias", HIDDEN);
    l->w2 = reg("logits_fc.weight", (long long)l->A * HIDDEN);
    l->b2 = reg("logits_fc.bias", l->A);
    l->wv = reg("value_fc.weight", HIDDEN);
    l->bv = reg("value_fc.bias", 1);
    l->nparams = off;
    // every tensor offset must keep float4 alignment for the vector loads of the FC rows
    if ((l->w1 % 4) || (l->w2 % 4)) { delete l; return lerr(BPP_E_INVALID, "unaligned parameter layout"); }

    int smem_optin = 0;
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    l->smem_cap = std::min(smem_optin, 200 * 1024);
    cudaDeviceGetAttribute(&l->sms, cudaDevAttrMultiProcessorCount, device);
    if (l->sms < 1) l->sms = 148;
    {   // fused stage kernels: measured faster than one launch per layer up to ~128 samples (one sample group per CTA),
        // slower beyond (profiles/r02_notes.md); BPP_LEARNER_FUSED=0 / 1 forces the choice
        const char* e = std::getenv("BPP_LEARNER_FUSED");
        l->fused_max_batch = e ? (e[0] == '1' ? 0x7fffffff : 0) : 128;
        if (std::getenv("BPP_LEARNER_UNFUSED")) l->fused_max_batch = 0;
    }
    const cudaFuncAttribute dyn = cudaFuncAttributeMaxDynamicSharedMemorySize;
    if (cudaFuncSetAttribute(k_lr_conv<8>, dyn, l->smem_cap) != cudaSuccess ||
        cudaFuncSetAttribute(k_lr_conv<4>, dyn, l->smem_cap) != cudaSuccess ||
        cudaFuncSetAttribute(k_lr_conv<2>, dyn, l->smem_cap) != cudaSuccess ||
        cudaFuncSetAttribute(k_lr_stage_fwd, dyn, l->smem_cap) != cudaSuccess ||
        cudaFuncSetAttribute(k_lr_stage_bwd, dyn, l->smem_cap) != cudaSuccess ||
        cudaFuncSetAttribute(k_lr_wgrad, dyn, l->smem_cap) != cudaSuccess ||
        cudaFuncSetAttribute(k_lr_heads<1>, dyn, l->smem_cap) != cudaSuccess ||
        cudaFuncSetAttribute(k_lr_heads<2>, dyn, l->smem_cap) != cudaSuccess ||
        cudaFuncSetAttribute(k_lr_heads<4>, dyn, l->smem_cap) != cudaSuccess ||
        cudaFuncSetAttribute(k_lr_heads<8>, dyn, l->smem_cap) != cudaSuccess ||
        cudaFuncSetAttribute(k_lr_heads<16>, dyn, l->smem_cap) != cudaSuccess) {
        cudaGetLastError();
        delete l;
        return lerr(BPP_E_CUDA, "cannot reserve shared memory for the learner kernels");
    }
    const size_t B = (size_t)max_batch;
    bool ok = dalloc(l, (void**)&l->planes, B * l->Cin * H * W * 4);
    for (int s = 0; s < 3 && ok; ++s) {
        const size_t ch = l->chans[s], hw0 = (size_t)l->hs[s] * l->ws[s], hw1 = (size_t)l->hs[s + 1] * l->ws[s + 1];
        ok = ok && dalloc(l, (void**)&l->c[s], B * ch * hw0 * 4) && dalloc(l, (void**)&l->dc[s], B * ch * hw0 * 4) &&
             dalloc(l, (void**)&l->amax[s], B * ch * hw1);
        float** t[10] = {&l->p[s], &l->a0[s], &l->q[s], &l->a1[s], &l->o[s], &l->dp[s], &l->da0[s], &l->dq[s], &l->da1[s],
                         &l->dout[s]};
        for (int k = 0; k < 10 && ok; ++k) ok = dalloc(l, (void**)t[k], B * ch * hw1 * 4);
    }
    l->max_chunks = std::max(l->sms, (max_batch + SMAX - 1) / SMAX);
    ok = ok && dalloc(l, (void**)&l->wfwd, (size_t)l->w1 * 4) && dalloc(l, (void**)&l->wbwd, (size_t)l->w1 * 4);
    ok = ok && dalloc(l, (void**)&l->partial, (size_t)l->max_chunks * ((l->nparams + 3) & ~3LL) * 4) &&
         dalloc(l, (void**)&l->loss_partial, (size_t)l->max_chunks * 2 * 4);
    if (!ok) {
        for (void* p : l->allocs) cudaFree(p);
        delete l;
        return lerr(BPP_E_NOMEM, "cudaMalloc of the learner buffers failed");
    }
    *out = l;
    return BPP_OK;
}

extern "C" int bpp_learner_destroy(bpp_learner* l) {
    if (!l) return BPP_OK;
    cudaSetDevice(l->device);
    for (void* p : l->allocs) cudaFree(p);
    delete l;
    return BPP_OK;
}

extern "C" int bpp_learner_num_params(bpp_learner* l, int64_t* out) {
    if (!l || !out) return lerr(BPP_E_INVALID, "null argument");
    *out = l->nparams;
    return BPP_OK;
}

extern "C" int bpp_learner_param_offset(bpp_learner* l, const char* name, int64_t* offset, int64_t* numel) {
    if (!l || !name || !offset || !numel) return lerr(BPP_E_INVALID, "null argument");
    for (auto& e : l->names)
        if (e.first == name) {
            *offset = e.second.first;
            *numel = e.second.second;
            return BPP_OK;
        }
    return lerr(BPP_E_INVALID, std::string("unknown parameter ") + name);
}

// forward + losses (+ backward and flat gradient when grads_out_dev != NULL)
extern "C" int bpp_learner_grad(bpp_learner* l, int B, const float* params_dev, const uint32_t* recs_dev,
                                const int32_t* items_wh_dev, const int64_t* ids_dev, const float* pis_dev,
                                const float* vs_dev, float* grads_out_dev, float* losses_out_dev, float* logp_out_dev,
                                float* v_out_dev, void* stream) {
    if (!l || !params_dev || !recs_dev || !items_wh_dev || !pis_dev || !vs_dev)
        return lerr(BPP_E_INVALID, "null argument");
    if (B < 1 || B > l->max_batch) return lerr(BPP_E_INVALID, "batch size out of range");
    cudaStream_t st = (cudaStream_t)stream;
    const bool train = grads_out_dev != nullptr;
    const bool fused = B <= l->fused_max_batch;
    const float* P = params_dev;
    int rc;
    {
        RelayoutArgs ra;
        for (int i = 0; i < NCONV; ++i) {
            ra.w_off[i] = l->conv[i].w_off;
            ra.cin[i] = l->conv[i].cin;
            ra.cout[i] = l->conv[i].cout;
        }
        k_lr_relayout<<<dim3(8, NCONV), 256, 0, st>>>(ra, P, l->wfwd, l->wbwd);
        const long long total = (long long)B * l->Cin * l->H * l->W;
        k_lr_planes<<<(int)std::min<long long>((total + 255) / 256, 4096), 256, 0, st>>>(B, l->W, l->H, l->N, recs_dev,
                                                                                       items_wh_dev, ids_dev, l->planes);
    }
    // ---- forward ----
    const float* u = l->planes;
    for (int s = 0; s < NSTAGE; ++s) {
        const ConvL* L = &l->conv[s * 5];
        const int ch = l->chans[s], h1 = l->hs[s + 1], w1 = l->ws[s + 1];
        if (fused) {   // one launch per stage: conv -> pool -> two residual blocks
            StageArgs a;
            a.in = u; a.p = l->p[s]; a.amax = l->amax[s]; a.dc = nullptr; a.l0_dgrad = 0;
            stage_layer(a.L[0], l->wfwd + L[0].w_off, P + L[0].b_off, nullptr, nullptr, l->c[s], L[0].cin, ch, L[0].h, L[0].w, 0);
            stage_layer(a.L[1], l->wfwd + L[1].w_off, P + L[1].b_off, nullptr, nullptr, l->a0[s], ch, ch, h1, w1, 1);
            stage_layer(a.L[2], l->wfwd + L[2].w_off, P + L[2].b_off, nullptr, l->p[s], l->q[s], ch, ch, h1, w1, 1);
            stage_layer(a.L[3], l->wfwd + L[3].w_off, P + L[3].b_off, nullptr, nullptr, l->a1[s], ch, ch, h1, w1, 1);
            stage_layer(a.L[4], l->wfwd + L[4].w_off, P + L[4].b_off, nullptr, l->q[s], l->o[s], ch, ch, h1, w1, 1);
            if ((rc = launch_stage(l, st, B, s, true, a))) return rc;
            u = l->o[s];
            continue;
        }
        if ((rc = launch_conv(l, st, B, u, l->wfwd + L[0].w_off, P + L[0].b_off, nullptr, nullptr, l->c[s], L[0].cin, ch,
                              L[0].h, L[0].w, 0)))
            return rc;
        {
            const long long total = (long long)B * ch * h1 * w1;
            k_lr_pool_fwd<<<(int)std::min<long long>((total + 255) / 256, 4096), 256, 0, st>>>(
                B * ch, L[0].h, L[0].w, h1, w1, l->c[s], l->p[s], l->amax[s]);
        }
        if ((rc = launch_conv(l, st, B, l->p[s], l->wfwd + L[1].w_off, P + L[1].b_off, nullptr, nullptr, l->a0[s], ch, ch,
                              h1, w1, 1)) ||
            (rc = launch_conv(l, st, B, l->a0[s], l->wfwd + L[2].w_off, P + L[2].b_off, nullptr, l->p[s], l->q[s], ch, ch,
                              h1, w1, 1)) ||
            (rc = launch_conv(l, st, B, l->q[s], l->wfwd + L[3].w_off, P + L[3].b_off, nullptr, nullptr, l->a1[s], ch, ch,
                              h1, w1, 1)) ||
            (rc = launch_conv(l, st, B, l->a1[s], l->wfwd + L[4].w_off, P + L[4].b_off, nullptr, l->q[s], l->o[s], ch, ch,
                              h1, w1, 1)))
            return rc;
        u = l->o[s];
    }
    // ---- heads: forward, losses, backward of the linear layers ----
    int S, nchunk;
    chunking(B, l->sms, S, nchunk);
    {
        HeadArgs a;
        a.B = B; a.S = S; a.flat = l->flat; a.A = l->A;
        a.o2 = l->o[2]; a.params = P;
        a.w1 = l->w1; a.b1 = l->b1; a.w2 = l->w2; a.b2 = l->b2; a.wv = l->wv; a.bv = l->bv;
        a.pis = pis_dev; a.vs = vs_dev; a.ids = ids_dev;
        a.dfeat = l->dout[2];
        a.partial = train ? l->partial : nullptr;
        a.nparams = (l->nparams + 3) & ~3LL;
        a.loss_partial = l->loss_partial;
        a.logp_out = logp_out_dev; a.v_out = v_out_dev;
        int SM = 1;
        while (SM < S) SM <<= 1;
        const size_t smem = ((size_t)SM * l->flat + (size_t)SM * HIDDEN + (size_t)SM * (l->A + 1)) * 4;
        if (smem > (size_t)l->smem_cap) return lerr(BPP_E_INVALID, "heads do not fit in shared memory");
        switch (SM) {
            case 1: k_lr_heads<1><<<nchunk, THREADS, smem, st>>>(a); break;
            case 2: k_lr_heads<2><<<nchunk, THREADS, smem, st>>>(a); break;
            case 4: k_lr_heads<4><<<nchunk, THREADS, smem, st>>>(a); break;
            case 8: k_lr_heads<8><<<nchunk, THREADS, smem, st>>>(a); break;
            default: k_lr_heads<16><<<nchunk, THREADS, smem, st>>>(a); break;
        }
    }
    if (!train) {
        k_lr_reduce<<<1, 32, 0, st>>>(0, 0, nchunk, l->partial, nullptr, l->loss_partial, losses_out_dev);
        const cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? BPP_OK
                                : lerr(BPP_E_CUDA, std::string("learner forward launch failed: ") + cudaGetErrorString(e));
    }
    // ---- backward: data gradients, last stage first ----
    for (int s = NSTAGE - 1; s >= 0; --s) {
        const ConvL* L = &l->conv[s * 5];
        const int ch = l->chans[s], h1 = l->hs[s + 1], w1 = l->ws[s + 1];
        if (fused) {   // one launch per stage: four data gradients, pooling backward, the conv's data gradient
            StageArgs a;
            a.in = l->dout[s]; a.p = nullptr; a.amax = l->amax[s]; a.dc = l->dc[s]; a.l0_dgrad = s > 0;
            stage_layer(a.L[0], l->wbwd + L[4].w_off, nullptr, l->a1[s], nullptr, l->da1[s], ch, ch, h1, w1, 0);
            stage_layer(a.L[1], l->wbwd + L[3].w_off, nullptr, l->q[s], l->dout[s], l->dq[s], ch, ch, h1, w1, 0);
            stage_layer(a.L[2], l->wbwd + L[2].w_off, nullptr, l->a0[s], nullptr, l->da0[s], ch, ch, h1, w1, 0);
            stage_layer(a.L[3], l->wbwd + L[1].w_off, nullptr, l->p[s], l->dq[s], l->dp[s], ch, ch, h1, w1, 0);
            stage_layer(a.L[4], l->wbwd + L[0].w_off, nullptr, nullptr, nullptr, s > 0 ? l->dout[s - 1] : nullptr, ch,
                        s > 0 ? L[0].cin : 8, L[0].h, L[0].w, 0);
            if ((rc = launch_stage(l, st, B, s, false, a))) return rc;
            continue;
        }
        // da1 = dgrad(res1.conv1, do) * (a1 > 0)
        if ((rc = launch_conv(l, st, B, l->dout[s], l->wbwd + L[4].w_off, nullptr, l->a1[s], nullptr, l->da1[s], ch, ch, h1,
                              w1, 0)) ||
            // dq = do + dgrad(res1.conv0, da1) * (q > 0)
            (rc = launch_conv(l, st, B, l->da1[s], l->wbwd + L[3].w_off, nullptr, l->q[s], l->dout[s], l->dq[s], ch, ch, h1,
                              w1, 0)) ||
            // da0 = dgrad(res0.conv1, dq) * (a0 > 0)
            (rc = launch_conv(l, st, B, l->dq[s], l->wbwd + L[2].w_off, nullptr, l->a0[s], nullptr, l->da0[s], ch, ch, h1,
                              w1, 0)) ||
            // dp = dq + dgrad(res0.conv0, da0) * (p > 0)
            (rc = launch_conv(l, st, B, l->da0[s], l->wbwd + L[1].w_off, nullptr, l->p[s], l->dq[s], l->dp[s], ch, ch, h1,
                              w1, 0)))
            return rc;
        {
            const long long total = (long long)B * ch * L[0].h * L[0].w;
            k_lr_pool_bwd<<<(int)std::min<long long>((total + 255) / 256, 4096), 256, 0, st>>>(
                B * ch, L[0].h, L[0].w, h1, w1, l->dp[s], l->amax[s], l->dc[s]);
        }
        if (s > 0) {  // gradient w.r.t. the previous stage's output (no activation in between)
            if ((rc = launch_conv(l, st, B, l->dc[s], l->wbwd + L[0].w_off, nullptr, nullptr, nullptr, l->dout[s - 1], ch,
                                  L[0].cin, L[0].h, L[0].w, 0)))
                return rc;
        }
    }
    // ---- weight gradients of all convolutions, then the ordered reduction over chunks ----
    {
        WgradArgs a;
        a.B = B; a.S = S; a.partial = l->partial; a.nparams = (l->nparams + 3) & ~3LL;
        size_t smem_max = 0;
        for (int s = 0; s < NSTAGE; ++s) {
            const float* ins[5] = {s == 0 ? l->planes : l->o[s - 1], l->p[s], l->a0[s], l->q[s], l->a1[s]};
            const float* dys[5] = {l->dc[s], l->da0[s], l->dq[s], l->da1[s], l->dout[s]};
            for (int k = 0; k < 5; ++k) {
                const ConvL& C = l->conv[s * 5 + k];
                WgradLayer& Lw = a.L[s * 5 + k];
                Lw.in = ins[k]; Lw.dy = dys[k];
                Lw.cin = C.cin; Lw.cout = C.cout; Lw.h = C.h; Lw.w = C.w; Lw.relu_in = k != 0;
                Lw.w_off = C.w_off; Lw.b_off = C.b_off;
                const int pairs = C.cin * (C.cout / 4);
                if (pairs > THREADS) return lerr(BPP_E_INVALID, "unsupported channel counts");
                const int PPs = ((C.h + 2) * (C.w + 2)) | 1;
                const size_t per = ((size_t)C.cin * PPs + (size_t)C.h * C.w * C.cout) * 4;
                int round = S;  // samples staged at once: keep a CTA near 56 KB so that four of them share an SM
                while (round > 1 && per * round + 16 > (size_t)std::min(l->smem_cap, 56 * 1024)) --round;
                const int pg = std::max(1, std::min(THREADS / pairs, round * C.h));
                Lw.pg = pg;
                Lw.round = round;
                const size_t need = std::max(per * round + 16, ((size_t)pg * pairs * 41 + 4 + (size_t)C.cout * C.cin * 9 + C.cout) * 4);
                if (need > (size_t)l->smem_cap) return lerr(BPP_E_INVALID, "weight gradient does not fit in shared memory");
                smem_max = std::max(smem_max, need);
            }
        }
        k_lr_wgrad<<<dim3(nchunk, NCONV), THREADS, smem_max, st>>>(a);
    }
    k_lr_reduce<<<(int)((l->nparams + 255) / 256), 256, 0, st>>>(l->nparams, (l->nparams + 3) & ~3LL, nchunk, l->partial,
                                                               grads_out_dev,
                                                               l->loss_partial, losses_out_dev);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? BPP_OK : lerr(BPP_E_CUDA, std::string("learner step launch failed: ") + cudaGetErrorString(e));
}

extern "C" int bpp_learner_adam(int64_t n, float* params_dev, const float* grads_dev, float* exp_avg_dev,
                                float* exp_avg_sq_dev, int step, const int32_t* step_dev, float grad_scale, float lr,
                                float beta1, float beta2, float eps, void* stream) {
    if (!params_dev || !grads_dev || !exp_avg_dev || !exp_avg_sq_dev || n < 0 || (step < 1 && !step_dev))
        return lerr(BPP_E_INVALID, "invalid argument");
    if (n == 0) return BPP_OK;
    const float bc1 = (float)(1.0 - std::pow((double)beta1, std::max(step, 1)));
    const float bc2s = (float)std::sqrt(1.0 - std::pow((double)beta2, std::max(step, 1)));
    k_lr_adam<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, params_dev, grads_dev, exp_avg_dev,
                                                                       exp_avg_sq_dev, grad_scale, lr, beta1, beta2, eps,
                                                                       bc1, bc2s, step_dev);
    return cudaGetLastError() == cudaSuccess ? BPP_OK : lerr(BPP_E_CUDA, "adam launch failed");
}
