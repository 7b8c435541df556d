// umma_probe.cu — how many SM cycles does one tcgen05.mma (M = 128, K = 16, bf16, operands in shared memory) cost as a
// function of N, the shared-memory layout of the A operand, the alignment of its start address and the accumulator
// pattern?  Measurement tool for the leaf-evaluator design; results in profiles/.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/umma_probe scripts/umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Case {
    int n;          // MMA N
    int layout;     // descriptor layout_type: 0 none, 6 SW32, 4 SW64, 2 SW128
    int lbo16, sbo16;
    int a_step16;   // start-address increment (16-byte units) between consecutive MMAs, cycling over 9 positions
    int a_base16;   // start-address offset of the first MMA
    int nd;         // accumulators written round-robin (1 = one dependent chain)
    int chain;      // MMAs per accumulator before moving to the next
    int m64;        // 1: M = 64
    int bg;         // background traffic of the other warps: 0 none, 1 st.shared.v4, 2 ld.shared.v4, 3 tcgen05.ld x16
};

__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(lbo16 & 0x3fffu) << 16) | ((uint64_t)(sbo16 & 0x3fffu) << 32) |
           (1ull << 46) | ((uint64_t)layout << 61);
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}

__global__ void __launch_bounds__(256, 1) k_probe(const Case* cases, int ncases, int nmma, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x;
    __shared__ volatile int s_stop;
    for (int i = tid; i < 160 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 7);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
    uint32_t parity = 0;
    for (int c = 0; c < ncases; ++c) {
        const Case cs = cases[c];
        long long dt = 0;
        if (tid == 0) s_stop = 0;
        __syncthreads();
        if (tid >= 32 && cs.bg) {
            // background traffic until the issuing thread has seen its MMAs complete
            uint4* buf = reinterpret_cast<uint4*>(smem + 64 * 1024) + (tid - 32);
            uint4 acc4 = make_uint4(tid, 0, 0, 0);
            uint32_t sink = 0;
            while (!s_stop) {
                if (cs.bg == 1) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) buf[k * 224] = acc4;
                } else if (cs.bg == 2) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        uint4 v;
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                                     : "r"(smem_u32(buf + k * 224)));
                        sink += v.x;
                    }
                } else {
                    uint32_t r[16];
                    const uint32_t taddr = tmem + ((uint32_t)(((tid >> 5) & 3) * 32) << 16) + 256u;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    sink += r[0] + r[15];
                }
            }
            if (sink == 0x12345678u) out[63] = sink;
        }
        if (tid < 32 && elect_one()) {
            const uint32_t m = cs.m64 ? 64u : 128u;
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(cs.n >> 3) << 17) | ((m >> 4) << 24);
            const uint32_t a0 = smem_u32(smem) + 4096;
            const uint32_t b0 = smem_u32(smem) + 140 * 1024;
            const uint64_t bdesc = mk_desc(b0, (uint32_t)cs.n, 8u, 0u);  // [2 K halves][N rows][16 B]
            const long long t0 = clock64();
            const uint64_t ad0 = mk_desc(a0 + 16u * (uint32_t)cs.a_base16, (uint32_t)cs.lbo16, (uint32_t)cs.sbo16,
                                         (uint32_t)cs.layout);
            const uint32_t ahi = (uint32_t)(ad0 >> 32), alo0 = (uint32_t)ad0;
            const uint32_t bhi = (uint32_t)(bdesc >> 32), blo = (uint32_t)bdesc;
            const uint32_t step = (uint32_t)cs.a_step16, dn = (uint32_t)cs.n;
            const uint32_t first_acc = cs.chain == 1 ? 0u : 1u;
            uint32_t d = 0, dmax = (uint32_t)(cs.nd * cs.n);
            const uint32_t dpos = cs.chain == 1 ? dn : 0u;   // chain 1: consecutive MMAs go to different accumulators
            const uint32_t dmask = (uint32_t)(cs.nd - 1);      // nd is a power of two
            for (int i = 0; i < nmma; i += 9) {
#pragma unroll
                for (int pos = 0; pos < 9; ++pos) {
                    const uint32_t acc = pos ? first_acc : 0u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
                        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem + d + dpos * ((uint32_t)pos & dmask)),
                        "r"(alo0 + step * (uint32_t)pos), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc)
                        : "memory");
                }
                if (cs.chain != 1) { d += dn; if (d >= dmax) d = 0; }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                         : "memory");
            // bounded wait
            uint32_t done = 0;
            for (long long spin = 0; spin < (1ll << 26) && !done; ++spin) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done)
                    : "r"(smem_u32(&bar)), "r"(parity)
                    : "memory");
            }
            dt = done ? clock64() - t0 : -1;
            if (blockIdx.x == 0) out[c] = dt;
            s_stop = 1;
        }
        parity ^= 1u;
        __syncthreads();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main(int argc, char** argv) {
    const int nmma = 1152;
    Case cs[64];
    const char* names[64];
    int n = 0;
    auto add = [&](const char* nm, Case c) { names[n] = nm; cs[n++] = c; };
    const int PL = 1072;  // plane stride of the production layout at 15x15, S = 4 (rows of 16 B)
    for (int N : {16, 32, 64, 128, 256}) {
        add("planar    aligned  1 acc", Case{N, 0, PL, 8, 0, 0, 1, 9, 0});
    }
    for (int N : {16, 32}) {
        add("planar    aligned  8 acc chain 9", Case{N, 0, PL, 8, 0, 0, 8, 9, 0});
        add("planar    aligned  8 acc chain 1", Case{N, 0, PL, 8, 0, 0, 8, 1, 0});
        add("planar    +1 row   1 acc", Case{N, 0, PL, 8, 0, 1, 1, 9, 0});
        add("planar    taps (step 1 row)", Case{N, 0, PL, 8, 1, 0, 1, 9, 0});
        add("planar    taps (step 17 rows)", Case{N, 0, PL, 8, 17, 0, 1, 9, 0});
        add("planar    step 8 rows (aligned)", Case{N, 0, PL, 8, 8, 0, 1, 9, 0});
        add("interleav aligned  (LBO 128 B, SBO 256 B)", Case{N, 0, 8, 16, 0, 0, 1, 9, 0});
        add("interleav step 1 row-pair (32 B)", Case{N, 0, 8, 16, 2, 0, 1, 9, 0});
        add("SW32      aligned  (SBO 256 B)", Case{N, 6, 1, 16, 0, 0, 1, 9, 0});
        add("SW32      step 1 row (32 B)", Case{N, 6, 1, 16, 2, 0, 1, 9, 0});
        add("SW64      aligned  (SBO 512 B)", Case{N, 4, 1, 32, 0, 0, 1, 9, 0});
        add("SW64      step 1 row (64 B)", Case{N, 4, 1, 32, 4, 0, 1, 9, 0});
        add("SW128     aligned  (SBO 1024 B)", Case{N, 2, 1, 64, 0, 0, 1, 9, 0});
        add("SW128     step 1 row (128 B)", Case{N, 2, 1, 64, 8, 0, 1, 9, 0});
        add("M=64 planar aligned", Case{N, 0, PL, 8, 0, 0, 1, 9, 1});
        add("planar taps + 7 warps st.shared.v4", Case{N, 0, PL, 8, 17, 0, 1, 9, 0, 1});
        add("planar taps + 7 warps ld.shared.v4", Case{N, 0, PL, 8, 17, 0, 1, 9, 0, 2});
        add("planar taps + 7 warps tcgen05.ld", Case{N, 0, PL, 8, 17, 0, 1, 9, 0, 3});
    }
    Case* d_cs;
    long long* d_out;
    cudaMalloc(&d_cs, sizeof(cs));
    cudaMalloc(&d_out, 64 * sizeof(long long));
    cudaMemcpy(d_cs, cs, sizeof(cs), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int blocks = argc > 1 ? atoi(argv[1]) : 1;
    for (int rep = 0; rep < 2; ++rep) {
        k_probe<<<blocks, 256, 200 * 1024>>>(d_cs, n, nmma, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    long long out[64];
    cudaMemcpy(out, d_out, sizeof(out), cudaMemcpyDeviceToHost);
    printf("tcgen05.mma M=128 K=16 bf16 SS, %d MMAs per case, one issuing thread, %d CTA(s)\n", nmma, blocks);
    for (int i = 0; i < n; ++i)
        printf("N=%3d  %-44s %8.1f cycles/MMA\n", cs[i].n, names[i], out[i] < 0 ? -1.0 : (double)out[i] / nmma);
    return 0;
}
