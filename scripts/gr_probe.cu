// gr_probe.cu — micro-measurements behind the grid-row trunk (bpp_net_gr.cuh): tcgen05.mma cost for the stacked-tap shapes
// (N = 48 / 96), A operand from shared memory vs tensor memory, issue -> commit -> mbarrier latency of short MMA batches, and
// the per-tile hand-over costs of the epilogue warps (cross-proxy fence, tcgen05 fences, mbarrier arrive / wake-up).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/gr_probe scripts/gr_probe.cu && scripts/bin/gr_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(lbo16 & 0x3fffu) << 16) | ((uint64_t)(sbo16 & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                 "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tLW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra LD;\n\tbra LW;\n\tLD:\n\t}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}

__global__ void __launch_bounds__(256, 1) k_probe(long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar, bar2;
    __shared__ uint32_t s_tmem;
    __shared__ volatile long long s_t;
    const int tid = threadIdx.x;
    for (int i = tid; i < 160 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 7);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    uint32_t par = 0, par2 = 0;
    int o = 0;
    const uint32_t a0 = smem_u32(smem) + 4096, b0 = smem_u32(smem) + 140 * 1024;
    const int NS[5] = {16, 32, 48, 64, 96};
    // 1. throughput, SS and TS, long chains
    for (int mode = 0; mode < 2; ++mode)
        for (int k = 0; k < 5; ++k) {
            const int N = NS[k];
            if (tid == 0) {
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
                const uint64_t bd = mk_desc(b0, (uint32_t)N, 8u);
                const long long t0 = clock64();
                for (int i = 0; i < 900; ++i) {
                    if (mode == 0) mma_ss(tmem, mk_desc(a0 + 16u * (uint32_t)(i % 3), 1072, 8), bd, idesc, i ? 1u : 0u);
                    else mma_ts(tmem, tmem + 256u + 8u * (uint32_t)(i % 3), bd, idesc, i ? 1u : 0u);
                }
                commit(smem_u32(&bar));
                wait(smem_u32(&bar), par);
                out[o] = clock64() - t0;
            }
            par ^= 1u;
            ++o;
            __syncthreads();
        }
    // 2. latency of short batches (N = 48, SS): k MMAs + commit -> barrier seen by the issuing thread
    const int KS[5] = {1, 2, 4, 7, 14};
    for (int k = 0; k < 5; ++k) {
        if (tid == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(48 >> 3) << 17) | ((128u >> 4) << 24);
            const uint64_t bd = mk_desc(b0, 48u, 8u);
            long long tot = 0;
            for (int rep = 0; rep < 16; ++rep) {
                const long long t0 = clock64();
                for (int i = 0; i < KS[k]; ++i) mma_ss(tmem, mk_desc(a0 + 16u * (uint32_t)(i % 3), 1072, 8), bd, idesc, i ? 1u : 0u);
                commit(smem_u32(&bar));
                wait(smem_u32(&bar), par);
                tot += clock64() - t0;
                par ^= 1u;
            }
            out[o] = tot / 16;
        }
        if (tid != 0) par ^= 0u;
        ++o;
        __syncthreads();
    }
    // keep the other threads' parity in step: thread 0 flipped 16 times per case (even) -> unchanged
    // 3. one warp: costs of the hand-over instructions (average of 32 repetitions)
    if (tid < 32) {
        uint4* buf = reinterpret_cast<uint4*>(smem + 64 * 1024) + tid;
        long long t0, acc[6] = {0, 0, 0, 0, 0, 0};
        for (int rep = 0; rep < 32; ++rep) {
            t0 = clock64();
            acc[0] += clock64() - t0;                                            // clock overhead
            t0 = clock64();
            buf[0] = make_uint4(rep, 1, 2, 3); buf[64] = make_uint4(rep, 1, 2, 3);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            acc[1] += clock64() - t0;                                            // 2 x st.shared.v4 + cross-proxy fence
            t0 = clock64();
            buf[128] = make_uint4(rep, 1, 2, 3); buf[192] = make_uint4(rep, 1, 2, 3);
            acc[2] += clock64() - t0;                                            // 2 x st.shared.v4 alone
            uint32_t r[8];
            t0 = clock64();
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                         : "r"(tmem + 64u));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc[3] += clock64() - t0;                                            // tcgen05.ld x8 + wait
            if (r[0] == 0x12345u) out[63] = r[1];
            t0 = clock64();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            acc[4] += clock64() - t0;
            t0 = clock64();
            __syncwarp();
            if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar2)) : "memory");
            acc[5] += clock64() - t0;                                            // syncwarp + arrive
        }
        if (tid == 0) for (int i = 0; i < 6; ++i) out[o + i] = acc[i] / 32;
    }
    o += 6;
    __syncthreads();
    // 4. wake-up latency: warp 1 waits on bar (fresh phase), warp 0 arrives and records when
    {
        long long tot = 0;
        for (int rep = 0; rep < 16; ++rep) {
            __syncthreads();
            if (tid == 32) {
                wait(smem_u32(&bar), par);
                tot += clock64() - s_t;
            }
            if (tid == 0) {
                for (int spin = 0; spin < 2000; ++spin) asm volatile("nanosleep.u32 20;");
                s_t = clock64();
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            par ^= 1u;
            __syncthreads();
        }
        if (tid == 32) out[o] = tot / 16;
        ++o;
    }
    (void)par2;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 64 * sizeof(long long));
    cudaMemset(d_out, 0, 64 * sizeof(long long));
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int rep = 0; rep < 2; ++rep) {
        k_probe<<<1, 256, 200 * 1024>>>(d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    long long out[64];
    cudaMemcpy(out, d_out, sizeof(out), cudaMemcpyDeviceToHost);
    const int NS[5] = {16, 32, 48, 64, 96};
    int o = 0;
    for (int mode = 0; mode < 2; ++mode)
        for (int k = 0; k < 5; ++k, ++o)
            printf("throughput M=128 K=16 N=%3d A from %s: %6.1f cycles/MMA\n", NS[k], mode ? "TMEM  " : "shared", out[o] / 900.0);
    const int KS[5] = {1, 2, 4, 7, 14};
    for (int k = 0; k < 5; ++k, ++o) printf("latency: %2d MMAs (N=48, SS) + commit -> barrier seen: %lld cycles\n", KS[k], out[o]);
    const char* nm[6] = {"clock64 pair", "2 x st.shared.v4 + fence.proxy.async", "2 x st.shared.v4", "tcgen05.ld x8 + wait::ld",
                         "tcgen05.fence::before_thread_sync", "syncwarp + mbarrier.arrive"};
    for (int i = 0; i < 6; ++i, ++o) printf("warp cost: %-40s %lld cycles\n", nm[i], out[o]);
    printf("mbarrier arrive -> try_wait returns in another warp: %lld cycles\n", out[o]);
    return 0;
}
