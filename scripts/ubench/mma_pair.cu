// Micro-benchmark: cycles per tcgen05.mma (kind::f16, K = 16, operands in shared memory, no swizzle) issued back to back,
// cta_group::1 (M = 128) against cta_group::2 (M = 256 over two SMs, B split between the CTAs), for N = 64 / 128 / 256.
// Answers: is the N = 64 MMA bound by shared-memory operand traffic (A 4 KB + B 2 KB per 32-cycle MMA), and does a CTA pair
// relieve it (A 4 KB + B 1 KB per SM)?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_pair mma_pair.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../alphazero-implementation_b200/csrc/tcgen05.cuh"
using namespace tc05;

__device__ __forceinline__ void umma2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tWC_LOOP:\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra WC_DONE;\n\tbra WC_LOOP;\n\tWC_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

constexpr int ROWS = 160;                      // 128 + guard, K-group-major: LBO = ROWS * 16
constexpr uint32_t LBO_A = ROWS * 16, SBO_A = 128;
constexpr int KG = 8;                          // 64 channels
constexpr uint32_t A_BYTES = KG * LBO_A;       // 20480
constexpr uint32_t B_BYTES = 256 * 64 * 2;     // up to N = 256, K = 64

template <bool PAIR>
__global__ void __launch_bounds__(128, 1) k_bench(int N, int iters, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + A_BYTES + B_BYTES);
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + A_BYTES + B_BYTES + 16);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = PAIR ? ctarank() : 0u;
    for (uint32_t i = tid; i < (A_BYTES + B_BYTES) / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3C003C00u;  // fp16 1.0
    if (warp == 0) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
        } else {
            tmem_alloc(smem_u32(slot), 256u);
        }
    }
    if (tid == 32) {
        mbar_init(smem_u32(bar), 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    if (PAIR) cluster_sync();
    fence_after();
    const uint32_t tmem = *slot;
    const uint32_t idesc = instr_desc(PAIR ? 256 : 128, N, true);
    const uint32_t a0 = smem_u32(smem) + 16 * 16, b0 = smem_u32(smem + A_BYTES);
    long long t0 = 0, t1 = 0;
    if (warp == 1 && rank == 0) {
        const uint64_t ad = smem_desc(a0, LBO_A, SBO_A);
        // B tile [rows][16] per K step: a pair holds N / 2 rows per CTA
        const uint32_t kstep_bytes = (uint32_t)(PAIR ? N / 2 : N) * 16 * 2;
        t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t bd = smem_desc(b0 + ks * kstep_bytes, 128, 256);
                    if (PAIR) umma2(tmem, ad + (uint64_t)(ks * (2 * LBO_A >> 4)) + (uint64_t)(i & 7), bd, idesc, 1u);
                    else umma(tmem, ad + (uint64_t)(ks * (2 * LBO_A >> 4)) + (uint64_t)(i & 7), bd, idesc, 1u);
                }
            }
            if (PAIR) umma_commit2(smem_u32(bar));
            else umma_commit(smem_u32(bar));
        }
        __syncwarp();
    }
    if (PAIR) mbar_wait_cluster(smem_u32(bar), 0u);
    else mbar_wait(smem_u32(bar), 0u);
    t1 = clock64();
    if (warp == 1 && rank == 0 && (tid & 31) == 0) out[0] = t1 - t0;
    fence_before();
    __syncthreads();
    if (PAIR) cluster_sync();
    if (warp == 0) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
        else tmem_dealloc(tmem, 256u);
    }
}

int main() {
    long long *d, h;
    cudaMalloc(&d, 8);
    const int smem = A_BYTES + B_BYTES + 64, iters = 2000;
    cudaFuncSetAttribute(k_bench<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_bench<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int pair = 0; pair < 2; ++pair)
        for (int N : {64, 128, 256}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (pair) {
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    cfg.attrs = at; cfg.numAttrs = 1;
                    cudaLaunchKernelEx(&cfg, k_bench<true>, N, iters, d);
                } else {
                    k_bench<false><<<1, 128, smem>>>(N, iters, d);
                }
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            const double per = (double)h / (4.0 * iters);
            const double bytes = 128 * 16 * 2 + (pair ? N / 2 : N) * 16 * 2;
            printf("cta_group::%d  M=%d N=%3d : %6.1f cycles per MMA (math floor %3d; shared-memory operand bytes per SM %5.0f -> %4.1f cycles at 128 B/clk)\n",
                   pair + 1, pair ? 256 : 128, N, per, 128 * N / 256, bytes, bytes / 128.0);
        }
    return 0;
}
