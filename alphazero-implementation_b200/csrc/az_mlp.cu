// Fused policy/value MLP of the reference's BasicNN (models/games/connect4/basic.py:8-39) on the 5th-generation
// tensor cores: 42 -> 512 (ReLU) -> 512 (ReLU) -> {7 policy logits, 2 values (tanh)} for a batch of leaf
// positions, ONE kernel, activations never leave the SM.
//
// One CTA (128 threads) owns a tile of 128 positions.  All three layers are tcgen05.mma (kind::f16, bf16 inputs,
// fp32 accumulation) issued by one thread, with the accumulator in tensor memory (all 512 TMEM columns: a
// 128 x 512 fp32 tile).  Between layers the four warps read their 32 accumulator lanes back with tcgen05.ld,
// apply bias + ReLU in fp32, round to bf16 and write the next layer's A operand straight into shared memory in
// the canonical K-major no-swizzle core-matrix layout the MMA descriptors address (8 rows x 16 bytes per core
// matrix; LBO = 128 B between K-adjacent core matrices, SBO between 8-row groups).  Weights are packed once per
// weight update (az_mlp_set_weights) into the same canonical layout, bf16, in 64-wide K chunks, so a chunk
// is a flat 64 KB copy into shared memory.  Only the conv/GEMM work runs on tensor cores; bias, ReLU and tanh
// are fp32 epilogues (north_star: "uses tensor cores only for its conv/GEMM layers").
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/az_engine.h"

namespace {

constexpr int TILE_M = 128;  // positions per CTA = UMMA M
constexpr int IN = 42;       // 6 x 7 grid
constexpr int K1 = 64;       // IN padded to one K chunk
constexpr int HID = 512;
constexpr int NH = 16;       // 7 logits + 2 values, padded to the smallest UMMA N for M = 128
constexpr int KC = 64;       // K chunk staged in shared memory
constexpr uint32_t LBO = 128;                       // bytes between K-adjacent core matrices
constexpr uint32_t SBO_ACT = (HID / 8) * 128;       // 8192: bytes between 8-row groups of the [128][512] activation tile
constexpr uint32_t SBO_CHUNK = (KC / 8) * 128;      // 1024: same for any [rows][64] tile (input tile, weight chunks)
constexpr uint32_t ACT_BYTES = TILE_M * HID * 2;    // 131072
constexpr uint32_t WBUF_BYTES = HID * KC * 2;       // 65536
constexpr uint32_t SMEM_BYTES = ACT_BYTES + WBUF_BYTES + 64;
constexpr uint32_t W1_ELEMS = HID * K1, W2_ELEMS = HID * HID, WH_ELEMS = NH * HID;

__host__ __device__ inline uint32_t canon(uint32_t r, uint32_t k, uint32_t sbo) {  // byte offset of element (r, k)
    return (r >> 3) * sbo + (k >> 3) * LBO + (r & 7) * 16 + (k & 7) * 2;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride byte offsets in
// 16-byte units, version 1 (Blackwell), no swizzle
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = bf16, both K-major, shape M x N
__host__ __device__ constexpr uint32_t instr_desc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(phase)
        : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&p);
}

// flat copy of a packed weight chunk into the shared-memory staging buffer
__device__ __forceinline__ void load_chunk(uint8_t *wbuf, const uint8_t *__restrict__ src, uint32_t bytes) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(wbuf);
    for (uint32_t i = threadIdx.x; i < bytes / 16; i += TILE_M) d[i] = __ldg(s + i);
}

// accumulator (128 x 512 fp32 in TMEM) -> bias + ReLU -> bf16 -> canonical [128][512] A operand in shared memory
__device__ __forceinline__ void epilogue_hidden(uint32_t tmem_base, uint8_t *act, const float *__restrict__ bias) {
    const uint32_t row = threadIdx.x;                                  // lane of TMEM = row of the tile
    const uint32_t taddr = tmem_base + ((row & ~31u) << 16);          // this warp's 32 lanes
    uint8_t *rowp = act + (row >> 3) * SBO_ACT + (row & 7) * 16;
#pragma unroll 1
    for (int cb = 0; cb < HID / 32; ++cb) {
        uint32_t v[32];
        tmem_ld32(taddr + cb * 32, v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaxf(__uint_as_float(v[q * 8 + j]) + __ldg(bias + cb * 32 + q * 8 + j), 0.0f);
            *reinterpret_cast<uint4 *>(rowp + (cb * 4 + q) * LBO) =
                make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
        }
    }
}

__global__ void __launch_bounds__(TILE_M, 1)
k_mlp_fused(const float *__restrict__ grid, long long n, const uint8_t *__restrict__ w1p, const float *__restrict__ b1,
            const uint8_t *__restrict__ w2p, const float *__restrict__ b2, const uint8_t *__restrict__ whp,
            const float *__restrict__ bh, float *__restrict__ logits, float *__restrict__ values) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *act = smem;
    uint8_t *wbuf = smem + ACT_BYTES;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + ACT_BYTES + WBUF_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + ACT_BYTES + WBUF_BYTES + 16);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const long long row0 = (long long)blockIdx.x * TILE_M;
    const uint32_t bar_addr = smem_u32(bar);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // input tile: raw grid values (-1 / 0 / 1, exact in bf16) as a [128][64] K-major tile, zero padded
    for (uint32_t i = tid; i < TILE_M * K1 * 2 / 16; i += TILE_M) reinterpret_cast<uint4 *>(act)[i] = make_uint4(0, 0, 0, 0);
    load_chunk(wbuf, w1p, W1_ELEMS * 2);
    __syncthreads();
    for (uint32_t e = tid; e < TILE_M * IN; e += TILE_M) {
        const uint32_t r = e / IN, k = e - r * IN;
        if (row0 + r < n) *reinterpret_cast<__nv_bfloat16 *>(act + canon(r, k, SBO_CHUNK)) = __float2bfloat16_rn(__ldg(grid + (row0 + r) * IN + k));
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t act_addr = smem_u32(act), wbuf_addr = smem_u32(wbuf);
    uint32_t phase = 0;

    // ---- layer 1: [128 x 64] . [512 x 64]^T
    if (tid == 0) {
#pragma unroll
        for (int ks = 0; ks < K1 / 16; ++ks)
#pragma unroll
            for (int half = 0; half < 2; ++half)
                umma(tmem_base + half * 256, smem_desc(act_addr + ks * 2 * LBO, SBO_CHUNK),
                     smem_desc(wbuf_addr + half * (256 / 8) * SBO_CHUNK + ks * 2 * LBO, SBO_CHUNK), instr_desc(128, 256), ks > 0);
        umma_commit(bar_addr);
    }
    mbar_wait(bar_addr, phase);
    phase ^= 1;
    tc_fence_after();
    load_chunk(wbuf, w2p, WBUF_BYTES);  // first K chunk of layer 2 (layer-1 MMAs are done with wbuf)
    epilogue_hidden(tmem_base, act, b1);

    // ---- layer 2: [128 x 512] . [512 x 512]^T, K in 8 chunks of 64
    for (int kc = 0; kc < HID / KC; ++kc) {
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks)
#pragma unroll
                for (int half = 0; half < 2; ++half)
                    umma(tmem_base + half * 256, smem_desc(act_addr + (kc * 8 + ks * 2) * LBO, SBO_ACT),
                         smem_desc(wbuf_addr + half * (256 / 8) * SBO_CHUNK + ks * 2 * LBO, SBO_CHUNK), instr_desc(128, 256),
                         (kc | ks) > 0);
            umma_commit(bar_addr);
        }
        mbar_wait(bar_addr, phase);
        phase ^= 1;
        tc_fence_after();
        if (kc + 1 < HID / KC) load_chunk(wbuf, w2p + (size_t)(kc + 1) * WBUF_BYTES, WBUF_BYTES);
        else load_chunk(wbuf, whp, WH_ELEMS * 2);  // head weights: 8 chunks of [16][64], 16 KB in all
    }
    epilogue_hidden(tmem_base, act, b2);

    // ---- heads: [128 x 512] . [16 x 512]^T
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        for (int kc = 0; kc < HID / KC; ++kc)
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks)
                umma(tmem_base, smem_desc(act_addr + (kc * 8 + ks * 2) * LBO, SBO_ACT),
                     smem_desc(wbuf_addr + kc * (NH / 8) * SBO_CHUNK + ks * 2 * LBO, SBO_CHUNK), instr_desc(128, NH), (kc | ks) > 0);
        umma_commit(bar_addr);
    }
    mbar_wait(bar_addr, phase);
    tc_fence_after();
    {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((tid & ~31u) << 16), v);
        const long long row = row0 + tid;
        if (row < n) {
#pragma unroll
            for (int j = 0; j < 7; ++j) logits[row * 7 + j] = __uint_as_float(v[j]) + __ldg(bh + j);
            values[row * 2 + 0] = tanhf(__uint_as_float(v[7]) + __ldg(bh + 7));
            values[row * 2 + 1] = tanhf(__uint_as_float(v[8]) + __ldg(bh + 8));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

// fp32 [N][K] row-major (nn.Linear.weight) -> bf16, canonical K-major chunks of 64 along K, rows padded to n_pad
__global__ void __launch_bounds__(256)
k_pack_weight(const float *__restrict__ w, int N, int K, int n_pad, int k_pad, uint8_t *__restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad * k_pad) return;
    const int nn = i / k_pad, k = i - nn * k_pad;
    const float v = (nn < N && k < K) ? w[(size_t)nn * K + k] : 0.0f;
    const size_t off = (size_t)(k / KC) * ((size_t)n_pad * KC * 2) + canon((uint32_t)nn, (uint32_t)(k % KC), SBO_CHUNK);
    *reinterpret_cast<__nv_bfloat16 *>(dst + off) = __float2bfloat16_rn(v);
}

__global__ void k_pack_head_rows(const float *__restrict__ wp, const float *__restrict__ wv, float *__restrict__ wh) {
    // rows 0..6 = policy_head.weight [7][512], rows 7..8 = value_head[0].weight [2][512]
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9 * HID) return;
    const int r = i / HID, k = i - r * HID;
    wh[i] = r < 7 ? wp[r * HID + k] : wv[(r - 7) * HID + k];
}

}  // namespace

struct az_mlp {
    int device;
    uint8_t *w1p, *w2p, *whp;
    float *b1, *b2, *bh, *wh_tmp;
    long long launches;
    char err[256];
};

extern "C" {

int32_t az_mlp_create(int32_t device, az_mlp **out) {
    if (!out) return AZ_E_INVALID;
    *out = nullptr;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return AZ_E_CUDA;
    az_mlp *m = (az_mlp *)calloc(1, sizeof(az_mlp));
    if (!m) return AZ_E_NOMEM;
    m->device = device;
    cudaSetDevice(device);
    cudaError_t e = cudaSuccess;
    void *p;
#define AL(field, bytes) if (e == cudaSuccess) { e = cudaMalloc(&p, (bytes)); m->field = (decltype(m->field))p; }
    AL(w1p, W1_ELEMS * 2); AL(w2p, W2_ELEMS * 2); AL(whp, WH_ELEMS * 2);
    AL(b1, HID * 4); AL(b2, HID * 4); AL(bh, NH * 4); AL(wh_tmp, 9 * HID * 4);
#undef AL
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) {
        cudaGetLastError();
        free(m);
        return AZ_E_CUDA;
    }
    *out = m;
    return AZ_OK;
}

int32_t az_mlp_destroy(az_mlp *m) {
    if (!m) return AZ_OK;
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    cudaFree(m->w1p); cudaFree(m->w2p); cudaFree(m->whp); cudaFree(m->b1); cudaFree(m->b2); cudaFree(m->bh); cudaFree(m->wh_tmp);
    free(m);
    return AZ_OK;
}

const char *az_mlp_last_error(const az_mlp *m) { return m ? m->err : "az_mlp: null handle"; }

/* fp32 device pointers in nn.Linear layout: w1 [512][42], w2 [512][512], wp [7][512], wv [2][512] and their biases */
int32_t az_mlp_set_weights(az_mlp *m, const float *w1, const float *b1, const float *w2, const float *b2, const float *wp,
                           const float *bp, const float *wv, const float *bv, void *stream) {
    if (!m || !w1 || !b1 || !w2 || !b2 || !wp || !bp || !wv || !bv) return AZ_E_INVALID;
    cudaSetDevice(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    k_pack_weight<<<(HID * K1 + 255) / 256, 256, 0, st>>>(w1, HID, IN, HID, K1, m->w1p);
    k_pack_weight<<<(HID * HID + 255) / 256, 256, 0, st>>>(w2, HID, HID, HID, HID, m->w2p);
    k_pack_head_rows<<<(9 * HID + 255) / 256, 256, 0, st>>>(wp, wv, m->wh_tmp);
    k_pack_weight<<<(NH * HID + 255) / 256, 256, 0, st>>>(m->wh_tmp, 9, HID, NH, HID, m->whp);
    m->launches += 4;
    cudaMemcpyAsync(m->b1, b1, HID * 4, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(m->b2, b2, HID * 4, cudaMemcpyDeviceToDevice, st);
    cudaMemsetAsync(m->bh, 0, NH * 4, st);
    cudaMemcpyAsync(m->bh, bp, 7 * 4, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(m->bh + 7, bv, 2 * 4, cudaMemcpyDeviceToDevice, st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(m->err, sizeof m->err, "az_mlp_set_weights: %s", cudaGetErrorString(e));
        return AZ_E_CUDA;
    }
    return AZ_OK;
}

/* BasicNN.forward (basic.py:29-39) on n positions: grid [n][42] f32 (AZ_LAYOUT_GRID_F32) -> logits [n][7], values [n][2] (tanh) */
int32_t az_mlp_forward(az_mlp *m, const float *grid, int64_t n, float *logits, float *values, void *stream) {
    if (!m || !grid || !logits || !values || n < 0) return AZ_E_INVALID;
    if (n == 0) return AZ_OK;
    cudaSetDevice(m->device);
    const int blocks = (int)((n + TILE_M - 1) / TILE_M);
    k_mlp_fused<<<blocks, TILE_M, SMEM_BYTES, (cudaStream_t)stream>>>(grid, n, m->w1p, m->b1, m->w2p, m->b2, m->whp, m->bh, logits, values);
    m->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(m->err, sizeof m->err, "k_mlp_fused launch: %s", cudaGetErrorString(e));
        return AZ_E_CUDA;
    }
    return AZ_OK;
}

int64_t az_mlp_launch_count(const az_mlp *m) { return m ? m->launches : 0; }

}  // extern "C"
