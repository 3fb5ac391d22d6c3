// Fused policy/value MLP of the reference's BasicNN (models/games/connect4/basic.py:8-39) on the 5th-generation
// tensor cores: 42 -> 512 (ReLU) -> 512 (ReLU) -> {7 policy logits, 2 values (tanh)} for a batch of leaf
// positions, ONE kernel, activations never leave the SM.
//
// One CTA owns a tile of 128 positions.  All three layers are tcgen05.mma (kind::f16, bf16 inputs, fp32 accumulation)
// with the accumulator in tensor memory (all 512 TMEM columns: a 128 x 512 fp32 tile).  Roles: warp 8 only streams weights
// - packed once per weight update (az_mlp_set_weights) into the canonical K-major no-swizzle core-matrix layout the MMA
// descriptors address, bf16, in 16-wide K chunks, so a chunk is a flat 16 KB region - through a five-stage shared-memory
// ring with bulk async copies (cp.async.bulk + mbarrier complete_tx); warp 0 waits for a stage and issues its MMAs (one
// elected lane), tcgen05.commit on the stage's mbarrier hands it back to the producer; warps 0..7 are the epilogue: between
// layers they read their accumulator lanes back with tcgen05.ld (two warps per 32 lanes, half the columns each, the next
// 32-column block in flight while the current one is processed), apply bias + ReLU in fp32, round to bf16 and write the
// next layer's A operand straight into shared memory (8 rows x 16 bytes per core matrix; LBO = 128 B between K-adjacent
// core matrices, SBO between 8-row groups).  Only the GEMM work runs on tensor cores; bias, ReLU and tanh are fp32
// epilogues (north_star: "uses tensor cores only for its conv/GEMM layers").
// Measured (scripts/mlp_clocks.py, one CTA, 16384 positions): 25.0 k cycles = set-up 4.6 k, layer 1 1.7 k + epilogue 3.1 k,
// layer 2 9.8 k (floor: 64 MMAs x 128 cycles = 8.2 k = the time the SM's ~64 B/clk L2 port needs for the layer's 512 KB of
// weights) + epilogue 3.0 k, heads 1.9 k.  The first version (issuing warp also refilling a two-stage ring, 4 epilogue
// warps with blocking tensor-memory loads) took 33.6 k; kernel 24.2 -> 18.7 us.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/az_engine.h"
#include "c4_bitboard.cuh"

namespace {

constexpr int TILE_M = 128;  // positions per CTA = UMMA M
constexpr int THREADS = 288;  // warps 0..7: epilogue (w and w + 4 share the accumulator lanes 32 * (w % 4).., half the columns each),
                              // warp 0 also issues the MMAs; warp 8 only streams weights
constexpr int ETHREADS = 256;
constexpr int IN = 42;       // 6 x 7 grid
constexpr int K1 = 64;       // IN padded to a multiple of the K chunk
constexpr int HID = 512;
constexpr int NH = 16;       // 7 logits + 2 values, padded to the smallest UMMA N for M = 128
constexpr int KC = 16;       // K chunk staged in shared memory (one MMA K step): 16 KB per chunk of a 512-row layer
constexpr int NS = 5;        // stages of the weight ring: 64 KB in flight keep the SM's ~64 B/clk L2 read port busy
constexpr uint32_t LBO = 128;                       // bytes between K-adjacent core matrices
constexpr uint32_t SBO_ACT = (HID / 8) * 128;       // 8192: bytes between 8-row groups of the [128][512] activation tile
constexpr uint32_t SBO_X = (K1 / 8) * 128;          // 1024: same for the [128][64] input tile
constexpr uint32_t SBO_CHUNK = (KC / 8) * 128;      // same for a [rows][KC] weight chunk
constexpr uint32_t ACT_BYTES = TILE_M * HID * 2;    // 131072
constexpr uint32_t STAGE_BYTES = HID * KC * 2;      // 32768
constexpr uint32_t BIAS_BYTES = 2 * HID * 4;          // both hidden layers' biases, fp32
constexpr uint32_t SMEM_BYTES = ACT_BYTES + NS * STAGE_BYTES + BIAS_BYTES + (2 * NS + 1) * 8 + 16;
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
// operand format: bf16 (1) or fp16 (0) - same rate and storage; fp16's 11 significand bits bring the outputs within 1e-3 of fp32
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, bool f16) {
    return (1u << 4) | (f16 ? 0u : (1u << 7) | (1u << 10)) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(phase), "r"(0x989680u)  /* suspend-time hint: the warp sleeps in the barrier unit instead of spinning on the issue slots the MMA / producer warps share */
        : "memory");
}
// one lane of a converged warp (the caller keeps every operand warp-uniform, so the MMA operands stay in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
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
// the same load without the wait: the caller overlaps it with the arithmetic on the previous block (tmem_ld_wait before use)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <bool F16>
__device__ __forceinline__ uint32_t pack16(float a, float b) {
    if (F16) {
        __half2 h = __floats2half2_rn(fminf(a, 65504.f), fminf(b, 65504.f));  // saturate instead of overflowing to inf
        return *reinterpret_cast<uint32_t *>(&h);
    }
    __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&p);
}

#ifdef AZ_TRUNK_CLOCKS
__device__ long long g_mlp_clk[16];
#define MCLK(i) do { if (blockIdx.x == 60 && threadIdx.x == 0) g_mlp_clk[i] = clock64(); } while (0)
#else
#define MCLK(i) do { } while (0)
#endif

// ---- weight pipeline.  Warp 8 streams every chunk of every layer, in order, through an NS-stage ring with bulk async copies
// (TMA, 1-D): an mbarrier per stage reports the bytes landed ("full"); tcgen05.commit reports when the MMAs that read a stage are
// done ("empty").  Warp 0 only waits for "full" and issues - it never waits for a copy it would have to start itself (the MMA
// queue is 2-3 instructions deep: a stalled issuer is a stalled tensor core), and 4 chunks stay in flight: a 16 KB chunk takes
// ~560 cycles to arrive (L2 latency + 64 B/clk), the two MMAs that consume it take 256.
struct Pipe {
    uint32_t full0, empty0, done, stage0;  // shared-memory addresses: full[NS], empty[NS], done barrier, first stage
    uint32_t g;                            // chunks consumed (issuer) / produced (producer) so far (stage = g % NS, use = g / NS)
};

__device__ __forceinline__ void produce_chunks(Pipe &p, const uint8_t *src, uint32_t nchunks, uint32_t chunk_bytes) {
#pragma unroll 1
    for (uint32_t i = 0; i < nchunks; ++i, ++p.g) {
        const uint32_t st = p.g % NS;
        if (p.g >= NS) mbar_wait(p.empty0 + st * 8, ((p.g / NS) - 1u) & 1u);  // the MMAs of the previous use have finished reading
        if (elect_one()) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(p.full0 + st * 8), "r"(chunk_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             p.stage0 + st * STAGE_BYTES),
                         "l"(src + (size_t)i * chunk_bytes), "r"(chunk_bytes), "r"(p.full0 + st * 8)
                         : "memory");
        }
        __syncwarp();
    }
}

// One layer's MMAs (warp 0, converged; one elected lane issues): D[128 x N] (+)= A[128 x K] . W[N x K]^T, W arriving in `nchunks`
// chunks; `ksteps` MMAs of K = 16 per chunk and N half.
__device__ __forceinline__ void issue_layer(Pipe &p, uint32_t nchunks, uint32_t ksteps, uint32_t a_addr, uint32_t a_sbo, uint32_t n_halves,
                                            uint32_t idesc, uint32_t tmem_base) {
#pragma unroll 1
    for (uint32_t i = 0; i < nchunks; ++i, ++p.g) {
        const uint32_t st = p.g % NS, stage = p.stage0 + st * STAGE_BYTES;
        mbar_wait(p.full0 + st * 8, (p.g / NS) & 1u);
        tc_fence_after();
        if (elect_one()) {
            for (uint32_t ks = 0; ks < ksteps; ++ks)
                for (uint32_t half = 0; half < n_halves; ++half)
                    umma(tmem_base + half * 256, smem_desc(a_addr + (i * ksteps + ks) * 2 * LBO, a_sbo),
                         // inside a stage: [rows][KC] sub-chunks back to back; K step ks -> sub-chunk ks / (KC/16), part ks % (KC/16)
                         smem_desc(stage + half * (256 / 8) * SBO_CHUNK + (ks / (KC / 16)) * (NH * KC * 2) + (ks % (KC / 16)) * 2 * LBO, SBO_CHUNK),
                         idesc, (i | ks) > 0);
            umma_commit(p.empty0 + st * 8);
        }
        __syncwarp();
    }
    if (elect_one()) umma_commit(p.done);
    __syncwarp();
}

// accumulator (128 x 512 fp32 in TMEM) -> bias + ReLU -> bf16 -> canonical [128][512] A operand in shared memory.
// 8 warps: thread = (row, column half); the next 32-column block is in flight while the current one is processed.
// packed fp32 pair add (FADD2) and ReLU + round-to-16-bit in one instruction (F2FP.RELU; fp16 saturates to the largest finite value):
// the epilogue is issue-bound - 1.3 instead of ~4 instructions per value
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 a, b, c;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tadd.rn.f32x2 c, a, b;\n\tmov.b64 {%0, %1}, c;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
template <bool F16>
__device__ __forceinline__ uint32_t pack16_relu(float2 v) {
    uint32_t r;
    if (F16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v.y), "f"(v.x));
    else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v.y), "f"(v.x));
    return r;
}
template <bool F16>
__device__ __forceinline__ void epilogue_block(const uint32_t (&v)[32], uint8_t *rowp, const float *bias, int cb) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b0 = *reinterpret_cast<const float4 *>(bias + cb * 32 + q * 8), b1 = *reinterpret_cast<const float4 *>(bias + cb * 32 + q * 8 + 4);
        const float2 f0 = fadd2(make_float2(__uint_as_float(v[q * 8]), __uint_as_float(v[q * 8 + 1])), make_float2(b0.x, b0.y));
        const float2 f1 = fadd2(make_float2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3])), make_float2(b0.z, b0.w));
        const float2 f2 = fadd2(make_float2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5])), make_float2(b1.x, b1.y));
        const float2 f3 = fadd2(make_float2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7])), make_float2(b1.z, b1.w));
        *reinterpret_cast<uint4 *>(rowp + (cb * 4 + q) * LBO) = make_uint4(pack16_relu<F16>(f0), pack16_relu<F16>(f1), pack16_relu<F16>(f2), pack16_relu<F16>(f3));
    }
}
template <bool F16>
__device__ __forceinline__ void epilogue_hidden(uint32_t tmem_base, uint8_t *act, const float *bias /* shared memory */) {
    const uint32_t row = threadIdx.x & 127u;                           // lane of TMEM = row of the tile
    const int cb0 = (int)(threadIdx.x >> 7) * (HID / 64);              // this warp group's 8 column blocks of 32
    const uint32_t taddr = tmem_base + ((row & ~31u) << 16);          // this warp's 32 lanes
    uint8_t *rowp = act + (row >> 3) * SBO_ACT + (row & 7) * 16;
    uint32_t va[32], vb[32];
    tmem_ld32_issue(taddr + cb0 * 32, va);
#pragma unroll 1
    for (int i = 0; i < HID / 64; i += 2) {
        tmem_ld_wait();
        tmem_ld32_issue(taddr + (cb0 + i + 1) * 32, vb);
        epilogue_block<F16>(va, rowp, bias, cb0 + i);
        tmem_ld_wait();
        if (i + 2 < HID / 64) tmem_ld32_issue(taddr + (cb0 + i + 2) * 32, va);
        epilogue_block<F16>(vb, rowp, bias, cb0 + i + 1);
    }
}

// column-major bitboard (bit = 7*col + row) -> row-major 42-bit mask (bit = 7*row + col): the element order of state.grid
__device__ __forceinline__ uint64_t row_major42(uint64_t bb) {
    uint64_t m = 0;
#pragma unroll
    for (int r = 0; r < c4::H; ++r) m |= (((((bb >> r) & 0x40810204081ull) * c4::LEGAL_MAGIC) >> 36) & 0x7Full) << (7 * r);
    return m;
}

// FROM_LEAVES: the input rows are built in the kernel from the engine's leaf bitboards (the leaf gather fused in);
// otherwise they are read from an AZ_LAYOUT_GRID_F32 batch.
template <bool FROM_LEAVES, bool F16>
__global__ void __launch_bounds__(THREADS, 1)
k_mlp_fused(const float *__restrict__ grid, const uint64_t *__restrict__ leaf_bb0, const uint64_t *__restrict__ leaf_bb1,
            const uint8_t *__restrict__ leaf_status, const int32_t *__restrict__ eval_list, const int32_t *__restrict__ eval_count, long long n_rows, const uint8_t *__restrict__ w1p, const float *__restrict__ b1,
            const uint8_t *__restrict__ w2p, const float *__restrict__ b2, const uint8_t *__restrict__ whp,
            const float *__restrict__ bh, float *__restrict__ logits, float *__restrict__ values) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *act = smem;
    float *s_bias = reinterpret_cast<float *>(smem + ACT_BYTES + NS * STAGE_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + ACT_BYTES + NS * STAGE_BYTES + BIAS_BYTES);  // full[NS] empty[NS] done
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + ACT_BYTES + NS * STAGE_BYTES + BIAS_BYTES + (2 * NS + 1) * 8);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const long long row0 = (long long)blockIdx.x * TILE_M;
    // FROM_LEAVES: row j is the leaf of slot eval_list[j], j < *eval_count (only the leaves that wait for an evaluation; the
    // outputs go to the slots' rows).  A block beyond the list has nothing to do.
    const long long n = (FROM_LEAVES && eval_list) ? (long long)__ldg(eval_count) : n_rows;
    if (row0 >= n) return;
    MCLK(0);
    // the leaf record of this thread's row is requested before the CTA set-up and consumed after it
    uint64_t in_b0 = 0, in_b1 = 0;
    bool in_live = false;
    long long out_row = -1;  // the row of logits / values this thread's row goes to
    if (FROM_LEAVES && tid < ETHREADS) {
        const long long row = row0 + (tid & 127u);
        if (row < n) {
            out_row = eval_list ? (long long)__ldg(eval_list + row) : row;
            in_live = leaf_status[out_row] == AZ_LEAF_EVAL;
            in_b0 = leaf_bb0[out_row];
            in_b1 = leaf_bb1[out_row];
        }
    } else if (!FROM_LEAVES && tid < ETHREADS) {
        const long long row = row0 + (tid & 127u);
        if (row < n) out_row = row;
    }
    Pipe p;
    p.full0 = smem_u32(bars);
    p.empty0 = smem_u32(bars + NS);
    p.done = smem_u32(bars + 2 * NS);
    p.stage0 = smem_u32(smem + ACT_BYTES);
    p.g = 0;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 32) {
#pragma unroll
        for (int i = 0; i < 2 * NS + 1; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();  // barriers initialised, tensor memory allocated
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t act_addr = smem_u32(act);

    if (warp == 8) {
        // ===== weight producer: layer 1 (K1 / KC chunks), layer 2 (HID / KC chunks), heads (one 16 KB chunk of [16][KC] sub-chunks) =====
        produce_chunks(p, w1p, K1 / KC, STAGE_BYTES);
        produce_chunks(p, w2p, HID / KC, STAGE_BYTES);
        produce_chunks(p, whp, 1, WH_ELEMS * 2);
    } else {
        // input tile: raw grid values (-1 / 0 / 1, exact in bf16) as a [128][64] K-major tile, zero padded
        for (uint32_t i = tid; i < 2 * HID; i += ETHREADS) s_bias[i] = i < HID ? __ldg(b1 + i) : __ldg(b2 + i - HID);
        if (FROM_LEAVES) {
            // thread = (row, half of the 8 K groups): 42 grid values from the leaf's bitboards, written as 16-byte core-matrix rows
            const uint32_t r = tid & 127u, kg0 = (tid >> 7) * (K1 / 16);
            const bool live = in_live;
            const uint64_t m0 = row_major42(in_b0), m1 = row_major42(in_b1);
            uint8_t *rowp = act + (r >> 3) * SBO_X + (r & 7) * 16;
#pragma unroll
            for (int kk = 0; kk < K1 / 16; ++kk) {
                const int kg = (int)kg0 + kk;
                uint32_t w[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    uint32_t pair = 0;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int e = kg * 8 + h * 2 + q;
                        uint32_t v = 0;
                        if (e < IN) v = ((m0 >> e) & 1ull) ? 0x0000u : (((m1 >> e) & 1ull) ? (F16 ? 0x3C00u : 0x3F80u) : (F16 ? 0xBC00u : 0xBF80u));  // 0, +1, -1
                        pair |= v << (16 * q);
                    }
                    w[h] = live ? pair : 0u;
                }
                *reinterpret_cast<uint4 *>(rowp + kg * LBO) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else {
            for (uint32_t i = tid; i < TILE_M * K1 * 2 / 16; i += ETHREADS) reinterpret_cast<uint4 *>(act)[i] = make_uint4(0, 0, 0, 0);
            asm volatile("bar.sync 1, %0;" ::"n"(ETHREADS) : "memory");
#pragma unroll 6
            for (uint32_t e = tid; e < TILE_M * IN; e += ETHREADS) {
                const uint32_t r = e / IN, k = e - r * IN;
                if (row0 + r < n) {
                    const float xv = __ldg(grid + (row0 + r) * IN + k);
                    if (F16) *reinterpret_cast<__half *>(act + canon(r, k, SBO_X)) = __float2half_rn(xv);
                    else *reinterpret_cast<__nv_bfloat16 *>(act + canon(r, k, SBO_X)) = __float2bfloat16_rn(xv);
                }
            }
        }
        fence_async_smem();
        asm volatile("bar.sync 1, %0;" ::"n"(ETHREADS) : "memory");

        // ---- layer 1: [128 x 64] . [512 x 64]^T
        MCLK(1);
        if (warp == 0) issue_layer(p, K1 / KC, KC / 16, act_addr, SBO_X, 2, instr_desc(128, 256, F16), tmem_base);
        MCLK(2);
        mbar_wait(p.done, 0);
        tc_fence_after();
        MCLK(3);
        epilogue_hidden<F16>(tmem_base, act, s_bias);
        MCLK(4);
        fence_async_smem();
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(ETHREADS) : "memory");
        tc_fence_after();

        // ---- layer 2: [128 x 512] . [512 x 512]^T, K in chunks of KC
        MCLK(5);
        if (warp == 0) issue_layer(p, HID / KC, KC / 16, act_addr, SBO_ACT, 2, instr_desc(128, 256, F16), tmem_base);
        MCLK(6);
        mbar_wait(p.done, 1);
        tc_fence_after();
        MCLK(7);
        epilogue_hidden<F16>(tmem_base, act, s_bias + HID);
        MCLK(8);
        fence_async_smem();
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(ETHREADS) : "memory");
        tc_fence_after();

        // ---- heads: [128 x 512] . [16 x 512]^T
        MCLK(9);
        if (warp == 0) issue_layer(p, 1, HID / 16, act_addr, SBO_ACT, 1, instr_desc(128, NH, F16), tmem_base);
        mbar_wait(p.done, 0);
        tc_fence_after();
        MCLK(10);
        if (tid < TILE_M) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((tid & ~31u) << 16), v);
            const long long row = out_row;
            if (row >= 0) {
#pragma unroll
                for (int j = 0; j < 7; ++j) logits[row * 7 + j] = __uint_as_float(v[j]) + __ldg(bh + j);
                values[row * 2 + 0] = tanhf(__uint_as_float(v[7]) + __ldg(bh + 7));
                values[row * 2 + 1] = tanhf(__uint_as_float(v[8]) + __ldg(bh + 8));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    MCLK(11);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

// fp32 [N][K] row-major (nn.Linear.weight) -> bf16, canonical K-major chunks of 64 along K, rows padded to n_pad
__global__ void __launch_bounds__(256)
k_pack_weight(const float *__restrict__ w, int N, int K, int n_pad, int k_pad, uint8_t *__restrict__ dst, int f16) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad * k_pad) return;
    const int nn = i / k_pad, k = i - nn * k_pad;
    const float v = (nn < N && k < K) ? w[(size_t)nn * K + k] : 0.0f;
    const size_t off = (size_t)(k / KC) * ((size_t)n_pad * KC * 2) + canon((uint32_t)nn, (uint32_t)(k % KC), SBO_CHUNK);
    if (f16) *reinterpret_cast<__half *>(dst + off) = __float2half_rn(v);
    else *reinterpret_cast<__nv_bfloat16 *>(dst + off) = __float2bfloat16_rn(v);
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
    int fmt;  // AZ_FMT_*: operand format of the packed weights and of the activations between layers
    uint8_t *w1p, *w2p, *whp;
    float *b1, *b2, *bh, *wh_tmp;
    long long launches;
    char err[256];
};

extern "C" {

#ifdef AZ_TRUNK_CLOCKS
int32_t az_debug_mlp_clocks(long long *out) { return cudaMemcpyFromSymbol(out, g_mlp_clk, sizeof(g_mlp_clk)) == cudaSuccess ? 0 : 1; }
#endif

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
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_fused<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_fused<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_fused<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_fused<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
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

/* operand format of the evaluator (AZ_FMT_BF16 default, AZ_FMT_F16); call before az_mlp_set_weights, which packs in that format */
int32_t az_mlp_set_operand_format(az_mlp *m, int32_t fmt) {
    if (!m || (fmt != AZ_FMT_BF16 && fmt != AZ_FMT_F16)) return AZ_E_INVALID;
    m->fmt = fmt;
    return AZ_OK;
}

const char *az_mlp_last_error(const az_mlp *m) { return m ? m->err : "az_mlp: null handle"; }

/* fp32 device pointers in nn.Linear layout: w1 [512][42], w2 [512][512], wp [7][512], wv [2][512] and their biases */
int32_t az_mlp_set_weights(az_mlp *m, const float *w1, const float *b1, const float *w2, const float *b2, const float *wp,
                           const float *bp, const float *wv, const float *bv, void *stream) {
    if (!m || !w1 || !b1 || !w2 || !b2 || !wp || !bp || !wv || !bv) return AZ_E_INVALID;
    cudaSetDevice(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int f16 = m->fmt == AZ_FMT_F16;
    k_pack_weight<<<(HID * K1 + 255) / 256, 256, 0, st>>>(w1, HID, IN, HID, K1, m->w1p, f16);
    k_pack_weight<<<(HID * HID + 255) / 256, 256, 0, st>>>(w2, HID, HID, HID, HID, m->w2p, f16);
    k_pack_head_rows<<<(9 * HID + 255) / 256, 256, 0, st>>>(wp, wv, m->wh_tmp);
    k_pack_weight<<<(NH * HID + 255) / 256, 256, 0, st>>>(m->wh_tmp, 9, HID, NH, HID, m->whp, f16);
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
    auto kern = m->fmt == AZ_FMT_F16 ? k_mlp_fused<false, true> : k_mlp_fused<false, false>;
    kern<<<blocks, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(grid, nullptr, nullptr, nullptr, nullptr, nullptr, n, m->w1p, m->b1, m->w2p, m->b2, m->whp, m->bh, logits, values);
    m->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(m->err, sizeof m->err, "k_mlp_fused launch: %s", cudaGetErrorString(e));
        return AZ_E_CUDA;
    }
    return AZ_OK;
}

/* same, with the leaf gather fused in: rows are the leaves chosen by the engine's last az_select_leaves
 * (row i = slot i; rows of terminal / idle slots are zero), so no az_gather_leaves launch is needed */
int32_t az_mlp_forward_leaves(az_mlp *m, az_engine *engine, float *logits, float *values, void *stream) {
    if (!m || !engine || !logits || !values) return AZ_E_INVALID;
    const uint64_t *bb0 = nullptr, *bb1 = nullptr;
    const uint8_t *status = nullptr;
    int32_t n = 0;
    if (az_leaf_arrays(engine, &bb0, &bb1, &status, &n) != AZ_OK || n <= 0) return AZ_E_INVALID;
    // walks the engine's compacted leaf list when the last selection produced one (az_set_leaf_compaction); at <= 148 tiles of 128
    // rows the kernel is one wave either way, so the BasicNN path leaves compaction off and saves the extra launch
    const int32_t *elist = nullptr, *ecount = nullptr;
    if (az_leaf_compact(engine, &elist, &ecount) != AZ_OK) return AZ_E_INVALID;
    cudaSetDevice(m->device);
    const int blocks = (n + TILE_M - 1) / TILE_M;
    auto kern = m->fmt == AZ_FMT_F16 ? k_mlp_fused<true, true> : k_mlp_fused<true, false>;
    kern<<<blocks, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(nullptr, bb0, bb1, status, elist, ecount, n, m->w1p, m->b1, m->w2p, m->b2, m->whp, m->bh, logits, values);
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
