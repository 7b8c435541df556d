// bpp_net_tc.cuh — tcgen05 / TMEM implicit-GEMM forward of the policy/value network (sm_100a).
//
// One CTA evaluates a group of S leaves through the WHOLE network without touching HBM in between: activations live
// in shared memory in the UMMA canonical K-major (no-swizzle) layout, accumulators in tensor memory.
//
// Convolution as implicit GEMM without im2col.  At spatial level l every sample is a zero-haloed grid and the S grids are
// concatenated into one long ROW axis (row = one padded pixel).  Halos are SHARED: a grid row is w+1 rows long (its zero
// column is the right halo of the grid row above and the left halo of its own), and a sample is h+1 grid rows (its zero
// row is the bottom halo of the previous sample), so a sample costs (h+1)(w+1) rows instead of (h+2)(w+2) - 29 % fewer
// MMAs and epilogue rows at 15x15 (256 rows per sample at level 0: a 128-row tile is exactly half a sample).  Activations are stored channels-last
// in 8-channel planes: plane p holds, for every row, the 16 bytes of channels 8p..8p+7 — exactly the 8-row x 16-byte
// "core matrix" tiling the tensor core reads for a K-major operand (SBO = 128 B between 8-row groups, LBO = plane
// stride between the two 8-channel halves of a K=16 slice).  For the 3x3 tap (dy, dx) the A operand of output rows
// [r0, r0+128) is the SAME buffer shifted by (dy-1)*(w+2) + (dx-1) rows, i.e. only the descriptor's start address
// changes: 9 (x Cin/16) tcgen05.mma instructions of shape 128 x Cout x 16 accumulate one output tile in TMEM.
// Rows that are halo or tile padding produce garbage accumulators that the epilogue never stores (halo rows of the
// activation buffers stay zero, which is the convolution's zero padding).
//
// Epilogue (all 128 threads, thread i = TMEM lane i = output row r0+i): tcgen05.ld -> + bias (+ residual) -> bf16 ->
// 16-byte stores into the next layer's operand planes.  Pooling re-grids level l into level l+1 on the CUDA cores.
// Weights are pre-arranged on the host in the UMMA B layout and staged per layer into shared memory.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace bpptc {

constexpr int TC_WORKERS = 256;   // 8 epilogue warps: warps w and w+4 own the same TMEM lane quarter and share the tiles
constexpr int TC_THREADS = 288;   // + warp 8, which issues the MMAs of a layer tile by tile, in order
constexpr int MAX_BARS = 16;

struct Level {
    int h, w, hp, wp, P, guard, RT, ntiles;  // RT = rows of one plane = guard + S*P + guard
    uint32_t mP, mwp, mhw, mw;               // division magics (fdiv) of P, wp, h*w, w
};

// n / d for n, d < 65536 with the host-computed magic m = ceil(2^32 / d) (0 encodes d = 1): one IMAD.HI instead of the
// ~35-instruction software division; index math was a quarter of the kernel's instructions before
__host__ __device__ __forceinline__ uint32_t fdiv_magic(uint32_t d) { return d <= 1 ? 0u : 0xffffffffu / d + 1u; }
__device__ __forceinline__ int fdiv(int n, uint32_t m) { return m ? (int)__umulhi((uint32_t)n, m) : n; }

struct TcParams {
    int S;                 // leaves per CTA
    Level lv[4];
    int regA_bytes;        // region A: input planes of level 0, then the {raw, actA, actB} triple of levels 1..3
    int regB_bytes;        // region B: conv outputs awaiting pooling (T_0..T_2), then the head scratch
    int wbuf_bytes;        // one layer of weights
    int compact;           // 1: compact arena (see the plan in bpp_net_create): T_0 aliases the input planes, the level
                           // triples live in region B, the weights wherever the current layer leaves room
    int smem_bytes;
    int tmem_cols;         // TMEM columns allocated per CTA: 512 / (CTAs per SM), a power of two
    long long w_off[15];   // element offsets of each conv layer in wts_umma
    int lay_n16[15];       // 16-byte chunks of each layer's staged weights (9 * cin16 * 2 * cout)
    const __nv_bfloat16* wts_umma;
    const __nv_bfloat16* wts_umma_lo;      // bf16(w - bf16(w)): low halves for the split-bf16 (x3) mode
    uint32_t mH;                           // fdiv magic of the bin height
    int A_pad;                             // action size rounded up to even
    const __nv_bfloat16* wts_logits_pad;   // logits weights [256][A_pad] (bf16x2 loads)
};

// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may become
// resident while its predecessor in the stream still runs; pdl_wait() blocks until the predecessor grid has completed and
// its writes are visible (a no-op for a normal launch), pdl_launch_dependents() lets the successor start launching.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading byte
// offset (between the two 8-element K halves) and stride byte offset (between 8-row groups) in 16-byte units,
// descriptor version 1 for sm_100.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(lbo16 & 0x3fffu) << 16) |
           ((uint64_t)(sbo16 & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16: D = f32, A = B = bf16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same, descriptors given as (lo, hi) halves: only the 14-bit start-address field in `lo` changes between MMAs of a
// layer, so the issuing thread needs one 32-bit add per operand instead of 64-bit arithmetic
__device__ __forceinline__ void umma_bf16_lh(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
        "}" ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&b);
}
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t x) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&x);
    v = __hmax2(v, __floats2bfloat162_rn(0.f, 0.f));
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
    __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a), y = *reinterpret_cast<__nv_bfloat162*>(&b);
    x = __hmax2(x, y);
    return *reinterpret_cast<uint32_t*>(&x);
}
__device__ __forceinline__ float bf16_lo(uint32_t x) { return __uint_as_float(x << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t x) { return __uint_as_float(x & 0xffff0000u); }

enum { EPI_CONV = 0, EPI_RES0 = 1, EPI_RES1 = 2 };

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi), for 8 values packed as two uint4 of bf16 pairs
__device__ __forceinline__ void split8(const float* u, uint4& h4, uint4& l4) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = pack_bf16(u[2 * i], u[2 * i + 1]);
        l[i] = pack_bf16(u[2 * i] - bf16_lo(h[i]), u[2 * i + 1] - bf16_hi(h[i]));
    }
    h4 = make_uint4(h[0], h[1], h[2], h[3]);
    l4 = make_uint4(l[0], l[1], l[2], l[3]);
}

// Register prefetch of the NEXT layer's weights and bias: the global loads are issued at the start of a layer and are
// consumed (stored to shared memory) at the start of the next one, so their L2 latency hides behind the layer's MMAs
// and epilogues without a second shared-memory weight buffer.
constexpr int WPRE = 4;  // 18,432 B / 16 B / 288 threads
struct WPre {
    uint4 w[WPRE];
    float b;
};
__device__ __forceinline__ void wpre_load(WPre& p, const __nv_bfloat16* wsrc, int n16, const float* bias, int cout) {
    const uint4* src = reinterpret_cast<const uint4*>(wsrc);
#pragma unroll
    for (int k = 0; k < WPRE; ++k) {
        const int i = threadIdx.x + k * TC_THREADS;
        if (i < n16) p.w[k] = __ldg(src + i);
    }
    if ((int)threadIdx.x < cout) p.b = __ldg(bias + threadIdx.x);
}

struct Ctx {
    uint32_t tmem;      // TMEM base (lane 0, column 0)
    uint32_t bar;       // shared address of mbarrier[0] (MAX_BARS barriers, 8 bytes apart)
    uint32_t phase;     // one parity bit per barrier
    unsigned char* wbuf;
    long long* prof;    // phase timers of thread 0, in shared memory (registers are scarce here): 0 input, 1 weights,
                        // 2 mma issue, 3 mma wait, 4 epilogue, 5 pool+zero, 6 heads / hand-over, 7 total
};
#define TC_PROF(slot, t0) do { if (threadIdx.x == 0) { const long long _t = clock64(); cx.prof[slot] += _t - (t0); (t0) = _t; } } while (0)

// One 3x3 convolution at level L: in_planes (cin16*2 planes of stride RT*16 bytes) -> epilogue `kind`.
//   EPI_CONV: out0 <- bf16(acc + bias)                       (raw conv output, to be pooled)
//   EPI_RES0: out0 <- relu(bf16(acc + bias))
//   EPI_RES1: v = acc + bias + raw; raw <- bf16(v); out0 <- relu(bf16(v))          (raw is read and updated in place)
// X3 = split-bf16 mode: every activation and weight is carried as hi + lo bf16 halves (lo planes / lo weights sit
// `*_lo_off` bytes behind the hi ones) and each product is the three MMAs hi*hi + hi*lo + lo*hi into the same fp32
// accumulator (the lo*lo term is below fp32 resolution): ~16 mantissa bits, enough for the reference's trained
// checkpoints whose logits reach 3e3.
template <bool X3>
__device__ __forceinline__ void conv_layer(const TcParams& T, Ctx& cx, const Level& L, int nvalid, int cin16, int cout,
                                           const __nv_bfloat16* __restrict__ wsrc, const float* __restrict__ bias,
                                           const unsigned char* in_planes, int kind, unsigned char* out0,
                                           unsigned char* raw, WPre& pre, const __nv_bfloat16* next_wsrc, int next_n16,
                                           const float* next_bias, int next_cout, const __nv_bfloat16* wsrc_lo = nullptr,
                                           uint32_t in_lo_off = 0, uint32_t out_lo_off = 0, uint32_t raw_lo_off = 0,
                                           bool a_lo_zero = false) {
    const int tid = threadIdx.x;
    long long tp = clock64();
    const int wbytes = 9 * cin16 * 2 * cout * 16;
    float* s_bias = reinterpret_cast<float*>(cx.wbuf + (X3 ? 2 : 1) * wbytes);  // Cout floats behind the staged weights
    if (X3) {
        // split mode: stage hi and lo weights straight from global memory (this mode runs one CTA per SM)
        const uint4* sh = reinterpret_cast<const uint4*>(wsrc);
        const uint4* sl = reinterpret_cast<const uint4*>(wsrc_lo);
        uint4* dst = reinterpret_cast<uint4*>(cx.wbuf);
        const int n16 = wbytes / 16;
        for (int i = tid; i < n16; i += TC_THREADS) {
            dst[i] = __ldg(sh + i);
            dst[n16 + i] = __ldg(sl + i);
        }
        if (tid < cout) s_bias[tid] = __ldg(bias + tid);
    } else {
        // this layer's weights (already in the UMMA B layout) were prefetched into registers: park them in shared memory
        uint4* dst = reinterpret_cast<uint4*>(cx.wbuf);
#pragma unroll
        for (int k = 0; k < WPRE; ++k) {
            const int i = tid + k * TC_THREADS;
            if (i < wbytes / 16) dst[i] = pre.w[k];
        }
        if (tid < cout) s_bias[tid] = pre.b;
        wpre_load(pre, next_wsrc, next_n16, next_bias, next_cout);  // in flight during this layer
    }
    fence_proxy_async();  // generic-proxy writes (weights, previous epilogue) -> visible to the tensor core's async proxy
    __syncthreads();
    TC_PROF(1, tp);
    const uint32_t idesc = umma_idesc(cout);
    const uint32_t a_base = smem_u32(in_planes);
    const uint32_t w_base = smem_u32(cx.wbuf);
    const uint32_t plane_b = (uint32_t)L.RT * 16u;
    const int rows_valid = nvalid * L.P;
    const int nt = min(L.ntiles, (rows_valid + 127) >> 7);  // tiles that hold at least one present sample
    // tiles in flight per pass (cout is 16 or 32 in this network: no integer division on the issuer's critical path)
    int tpp = cout == 16 ? T.tmem_cols >> 4 : cout == 32 ? T.tmem_cols >> 5 : T.tmem_cols / cout;
    if (tpp > MAX_BARS) tpp = MAX_BARS;
    for (int t0 = 0; t0 < nt; t0 += tpp) {
      const int nb = min(tpp, nt - t0);
      // ---- issue: one elected lane of the issuer warp queues the MMAs of all nb tiles IN TILE ORDER (each tile into its
      // own TMEM columns, followed by a commit to its own mbarrier): tile 0 completes after its own 9 * cin16 MMAs and
      // its epilogue runs while the tensor core works on the later tiles.  (Issuing tile b from warp b, as an earlier
      // version did, interleaves the tiles in the queue: they all complete together at the end and nothing overlaps.)
      if (tid >= TC_WORKERS) {
       if (elect_one()) {
        tc_fence_after();
        const uint64_t b0 = umma_desc(w_base, (uint32_t)cout, 8u);
        const uint64_t a00 = umma_desc(a_base + (uint32_t)L.guard * 16u, (uint32_t)L.RT, 8u);
        const uint32_t ahi = (uint32_t)(a00 >> 32), bhi = (uint32_t)(b0 >> 32), blo0 = (uint32_t)b0;
        const uint32_t a_kc = 2u * plane_b >> 4, b_blk = (uint32_t)(2 * cout), b_tap = b_blk * (uint32_t)cin16;
        const uint32_t a_lo16 = in_lo_off >> 4, w_lo16 = (uint32_t)wbytes >> 4;
        for (int b = 0; b < nb; ++b) {
            const uint32_t alo0 = (uint32_t)a00 + (uint32_t)(t0 + b) * 128u;  // +128 rows (16-byte units) per tile
            const uint32_t d = cx.tmem + (uint32_t)(b * cout);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                uint32_t at = alo0 + (uint32_t)((tap / 3 - 1) * L.wp + (tap % 3 - 1));
                const uint32_t bt = blo0 + (uint32_t)tap * b_tap;
                umma_bf16_lh(d, at, ahi, bt, bhi, idesc, tap > 0 ? 1u : 0u);
                if (X3) {
                    umma_bf16_lh(d, at, ahi, bt + w_lo16, bhi, idesc, 1u);                          // a_hi * w_lo
                    if (!a_lo_zero) umma_bf16_lh(d, at + a_lo16, ahi, bt, bhi, idesc, 1u);          // a_lo * w_hi
                }
                if (cin16 == 2) {
                    umma_bf16_lh(d, at + a_kc, ahi, bt + b_blk, bhi, idesc, 1u);
                    if (X3) {
                        umma_bf16_lh(d, at + a_kc, ahi, bt + b_blk + w_lo16, bhi, idesc, 1u);
                        if (!a_lo_zero) umma_bf16_lh(d, at + a_kc + a_lo16, ahi, bt + b_blk, bhi, idesc, 1u);
                    }
                }
            }
            umma_commit(cx.bar + 8u * (uint32_t)b);
        }
       }
       __syncwarp();
      }
      TC_PROF(2, tp);
      const int half = tid >> 7, lt = tid & 127;  // lt = TMEM lane = row inside the tile
      const int nch = cout >> 4;                    // 16-column chunks per tile
      // work unit = (tile, chunk): the two halves (warps w and w + 4 read the same TMEM lane quarter) alternate units, so
      // that a single-tile layer with 32 output channels still occupies all eight warps
      for (int un = half; un < nb * nch && half < 2; un += 2) {
        const int b = nch == 2 ? un >> 1 : un;
        const int c0 = nch == 2 ? (un & 1) << 4 : 0;
        const int t = t0 + b;
        mbar_wait(cx.bar + 8u * (uint32_t)b, (cx.phase >> b) & 1u);
        tc_fence_after();
        TC_PROF(3, tp);
        // ---- epilogue: thread tid owns output row t*128 + tid
        const int rl = t * 128 + lt;
        const int j = fdiv(rl, L.mP), q = rl - j * L.P;
        const int yp = fdiv(q, L.mwp), xp = q - yp * L.wp;
        const bool interior = rl < rows_valid && yp >= 1 && yp <= L.h && xp >= 1 && xp <= L.w;
        const size_t rowb = (size_t)(L.guard + rl) * 16;
        const uint32_t taddr = cx.tmem + ((uint32_t)(((lt >> 5) & 3) * 32) << 16) + (uint32_t)(b * cout);
        // EPI_CONV (bf16 mode): the halo rows of T are set to -inf, the max-pool's padding value, so that the pooling
        // pass reads its 3x3 windows without bounds checks (rows up to one grid row behind the last sample)
        const bool pad_row = !X3 && kind == EPI_CONV && !interior && rl < rows_valid + L.wp + 1;
        {
            float v[16];
            tmem_ld16(taddr + (uint32_t)c0, v);  // warp-collective: executed by every lane
            if (pad_row) {
                const uint4 ninf = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);
                *reinterpret_cast<uint4*>(out0 + (size_t)(c0 >> 3) * plane_b + rowb) = ninf;
                *reinterpret_cast<uint4*>(out0 + (size_t)((c0 >> 3) + 1) * plane_b + rowb) = ninf;
            }
            if (interior) {
                const float4* b4 = reinterpret_cast<const float4*>(s_bias + c0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 bb = b4[i];
                    v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
                }
#pragma unroll
                for (int hp8 = 0; hp8 < 2; ++hp8) {
                    const int plane = (c0 >> 3) + hp8;
                    float* u = v + hp8 * 8;
                    const size_t poff = (size_t)plane * plane_b + rowb;
                    if (!X3) {
                        uint4 o;
                        if (kind == EPI_RES1) {
                            uint4* rp = reinterpret_cast<uint4*>(raw + poff);
                            const uint4 rv = *rp;
                            u[0] += bf16_lo(rv.x); u[1] += bf16_hi(rv.x); u[2] += bf16_lo(rv.y); u[3] += bf16_hi(rv.y);
                            u[4] += bf16_lo(rv.z); u[5] += bf16_hi(rv.z); u[6] += bf16_lo(rv.w); u[7] += bf16_hi(rv.w);
                            o = make_uint4(pack_bf16(u[0], u[1]), pack_bf16(u[2], u[3]), pack_bf16(u[4], u[5]),
                                           pack_bf16(u[6], u[7]));
                            *rp = o;
                        } else {
                            o = make_uint4(pack_bf16(u[0], u[1]), pack_bf16(u[2], u[3]), pack_bf16(u[4], u[5]),
                                           pack_bf16(u[6], u[7]));
                        }
                        if (kind != EPI_CONV)
                            o = make_uint4(relu_bf16x2(o.x), relu_bf16x2(o.y), relu_bf16x2(o.z), relu_bf16x2(o.w));
                        *reinterpret_cast<uint4*>(out0 + poff) = o;
                    } else {
                        if (kind == EPI_RES1) {
                            const uint4 rh = *reinterpret_cast<const uint4*>(raw + poff);
                            const uint4 rl2 = *reinterpret_cast<const uint4*>(raw + raw_lo_off + poff);
                            u[0] += bf16_lo(rh.x) + bf16_lo(rl2.x); u[1] += bf16_hi(rh.x) + bf16_hi(rl2.x);
                            u[2] += bf16_lo(rh.y) + bf16_lo(rl2.y); u[3] += bf16_hi(rh.y) + bf16_hi(rl2.y);
                            u[4] += bf16_lo(rh.z) + bf16_lo(rl2.z); u[5] += bf16_hi(rh.z) + bf16_hi(rl2.z);
                            u[6] += bf16_lo(rh.w) + bf16_lo(rl2.w); u[7] += bf16_hi(rh.w) + bf16_hi(rl2.w);
                            uint4 h4, l4;
                            split8(u, h4, l4);
                            *reinterpret_cast<uint4*>(raw + poff) = h4;
                            *reinterpret_cast<uint4*>(raw + raw_lo_off + poff) = l4;
                        }
                        if (kind != EPI_CONV) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) u[i] = fmaxf(u[i], 0.f);
                        }
                        uint4 h4, l4;
                        split8(u, h4, l4);
                        *reinterpret_cast<uint4*>(out0 + poff) = h4;
                        *reinterpret_cast<uint4*>(out0 + out_lo_off + poff) = l4;
                    }
                }
            }
        }
        TC_PROF(4, tp);
      }
      cx.phase ^= (1u << nb) - 1u;  // every barrier of this pass completed one phase
      tc_fence_before();
      __syncthreads();  // TMEM columns are free again; epilogue stores are ordered before the next layer's proxy fence
    }
}

// max_pool2d(kernel 3, stride 2, padding 1) from the conv output T (level La) into raw/actA of level Lb
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}

// split-bf16 variant: values are hi + lo (lo buffers `*_lo` bytes behind)
__device__ __forceinline__ void pool_level_x3(const Level& La, const Level& Lb, int nvalid, int planes,
                                              const unsigned char* Tbuf, uint32_t t_lo, unsigned char* raw, uint32_t raw_lo,
                                              unsigned char* actA, uint32_t act_lo) {
    const int per = Lb.h * Lb.w;
    const int total = nvalid * per * planes;
    const uint32_t mnp = fdiv_magic((uint32_t)(nvalid * per));
    for (int idx = threadIdx.x; idx < total; idx += TC_THREADS) {
        const int p = fdiv(idx, mnp);
        int r = idx - p * nvalid * per;
        const int j = fdiv(r, Lb.mhw);
        r -= j * per;
        const int oy = fdiv(r, Lb.mw), ox = r - oy * Lb.w;
        float m[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = 2 * oy + dy;
            if (yy < 0 || yy >= La.h) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = 2 * ox + dx;
                if (xx < 0 || xx >= La.w) continue;
                const size_t row = (size_t)La.guard + (size_t)j * La.P + (size_t)(yy + 1) * La.wp + (xx + 1);
                const size_t off = ((size_t)p * La.RT + row) * 16;
                float a[8], b[8];
                unpack8(*reinterpret_cast<const uint4*>(Tbuf + off), a);
                unpack8(*reinterpret_cast<const uint4*>(Tbuf + t_lo + off), b);
#pragma unroll
                for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], a[i] + b[i]);
            }
        }
        const size_t orow = (size_t)Lb.guard + (size_t)j * Lb.P + (size_t)(oy + 1) * Lb.wp + (ox + 1);
        const size_t ooff = ((size_t)p * Lb.RT + orow) * 16;
        uint4 h4, l4;
        split8(m, h4, l4);
        *reinterpret_cast<uint4*>(raw + ooff) = h4;
        *reinterpret_cast<uint4*>(raw + raw_lo + ooff) = l4;
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], 0.f);
        split8(m, h4, l4);
        *reinterpret_cast<uint4*>(actA + ooff) = h4;
        *reinterpret_cast<uint4*>(actA + act_lo + ooff) = l4;
    }
}

// bf16 mode: the halo rows of T hold -inf (written by the EPI_CONV epilogue and pool_pad_tail), so every 3x3 window is
// read without bounds checks: window row yy = 2*oy + dy maps to grid row yy + 1, i.e. 2*oy .. 2*oy + 2
__device__ __forceinline__ void pool_level(const Level& La, const Level& Lb, int nvalid, int planes,
                                           const unsigned char* Tbuf, unsigned char* raw, unsigned char* actA) {
    const int per = Lb.h * Lb.w;
    const int total = nvalid * per * planes;
    const uint32_t mnp = fdiv_magic((uint32_t)(nvalid * per));
    for (int idx = threadIdx.x; idx < total; idx += TC_THREADS) {
        const int p = fdiv(idx, mnp);
        int r = idx - p * nvalid * per;
        const int j = fdiv(r, Lb.mhw);
        r -= j * per;
        const int oy = fdiv(r, Lb.mw), ox = r - oy * Lb.w;
        const uint4* src = reinterpret_cast<const uint4*>(Tbuf) + (size_t)p * La.RT + La.guard + j * La.P +
                           (2 * oy) * La.wp + 2 * ox;
        uint4 v[9];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) v[dy * 3 + dx] = src[dy * La.wp + dx];
        uint4 m = v[0];
#pragma unroll
        for (int k = 1; k < 9; ++k)
            m = make_uint4(max_bf16x2(m.x, v[k].x), max_bf16x2(m.y, v[k].y), max_bf16x2(m.z, v[k].z), max_bf16x2(m.w, v[k].w));
        const size_t orow = (size_t)Lb.guard + (size_t)j * Lb.P + (size_t)(oy + 1) * Lb.wp + (ox + 1);
        *reinterpret_cast<uint4*>(raw + ((size_t)p * Lb.RT + orow) * 16) = m;
        *reinterpret_cast<uint4*>(actA + ((size_t)p * Lb.RT + orow) * 16) =
            make_uint4(relu_bf16x2(m.x), relu_bf16x2(m.y), relu_bf16x2(m.z), relu_bf16x2(m.w));
    }
}

// the grid row behind the last present sample (its bottom halo) belongs to no processed tile when the samples end on a
// tile boundary: set it to -inf here (runs between the convolution's final barrier and the barrier before pool_level)
__device__ __forceinline__ void pool_pad_tail(const Level& La, int nvalid, int planes, unsigned char* Tbuf) {
    const uint4 ninf = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);
    const int n = La.wp + 1;
    for (int idx = threadIdx.x; idx < planes * n; idx += TC_THREADS) {
        const int p = idx / n, k = idx - p * n;
        reinterpret_cast<uint4*>(Tbuf)[(size_t)p * La.RT + La.guard + nvalid * La.P + k] = ninf;
    }
}

__device__ __forceinline__ void zero_bytes(unsigned char* p, int bytes) {
    uint4* q = reinterpret_cast<uint4*>(p);
    for (int i = threadIdx.x; i < bytes / 16; i += TC_THREADS) q[i] = make_uint4(0, 0, 0, 0);
}

}  // namespace bpptc
