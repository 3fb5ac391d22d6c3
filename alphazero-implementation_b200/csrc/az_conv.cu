// Fused ResNet-style policy/value network on the 5th-generation tensor cores (tcgen05 + TMEM): stem conv3x3 (3 -> 64),
// any number of residual blocks (2 x conv3x3 64 -> 64, skip, ReLU) and the policy / value heads of the reference's
// ResNet (src/alphazero_simple/resnet.py:13-103, BatchNorm folded) for a batch of leaf positions in ONE kernel; the
// activations of a position never leave the SM between layers.  The leaf gather is fused in: the stem's input
// planes (empty / side to move / opponent, cnn.py:93-95) are built from the engine's leaf bitboards.
//
// Convolution as implicit GEMM without im2col.  Pixel = one GEMM row, channels = K.  The activation buffer is
// stored K-group-major: for every group of 8 channels, all rows back to back at 16 bytes per row.  In the MMA's
// K-major no-swizzle shared-memory descriptor this is "stride between 8-row groups = 128 B, stride between
// K-adjacent core matrices = rows * 16 B", so the A operand of filter tap (dy, dx) is THE SAME buffer with the
// start address moved by (8*dy + dx) rows: nine taps = nine descriptor offsets, no data movement.
//  * Compact padding.  A position is 7 x 8 pixel rows (6 x 7 cells + ONE zero column on the right + ONE zero row on
//    top): with a pixel-row stride of 8 the cell left of x = 0 is the zero column of the row below, and the row under
//    y = 0 is the zero row of the previous position (a leading zero row opens each group).  Outputs computed for
//    padding rows are written back as zeros, which keeps the padding zero for the next layer.  4 positions + the
//    leading row = 232 <= 256 rows = 2 accumulator tiles of 128 rows x 64 fp32 columns in tensor memory; a CTA takes
//    two such groups: 8 positions in 4 tiles.
//  * Two groups, ping-pong.  Warp 8 only issues MMAs: for every layer group A's 72 (9 taps x 2 tiles x 4 K-steps),
//    then group B's 72.  The eight epilogue warps (two per 32 accumulator lanes, half the channels each) trail it:
//    tcgen05.ld, bias (+ skip) + ReLU in fp32, round to bf16, write the next layer's A operand - while they work on
//    A's accumulators the tensor core works on B, and vice versa.  Hand-offs are mbarriers: tcgen05.commit ->
//    mma_done[g]; one arrival per epilogue warp -> epi_done[g].
//  * Warp 9 only streams weights ([64 out][64 in] bf16 per tap, 8 KB, packed once per weight update) with bulk async
//    copies.  The ring has 9 stages = the 9 taps of a layer (stage = tap, parity = layer): a tap is loaded once per
//    layer, used by A and then by B, whose commit releases the stage for the next layer's tap.  The issuer never
//    waits for a copy it has to start itself: the MMA queue is only 2-3 instructions deep (measured with the
//    AZ_TRUNK_CLOCKS build, scripts/trunk_clocks.py), so every stall of the issuing warp is a stall of the tensor core.
//  * Persistent CTAs.  One CTA per SM walks over batches of 8 positions: tensor memory, barriers (their phases just continue),
//    biases and the zero padding are set up once, the weight producer runs ahead across the batch boundary, and the next
//    batch's leaf records are requested while the FC tail of the current one runs.
//  * Heads.  After the last block the policy conv1x1 and the value conv3x3 run as ONE 48-channel conv layer (the 1x1
//    weights sit in the centre tap); the two small FC layers run on CUDA cores in fp32, each weight read once per CTA.
// Measured (16384 positions, 4 blocks): 0.70 ms; a group's 72 MMAs take ~3750 cycles = 52 cycles per 128x64x16 MMA
// against the 32-cycle floor: each fetches 6 KB of operands from shared memory (~115 B/clk), i.e. with 64 output
// channels the conv MMAs are shared-memory-operand-bound.  A CTA-pair variant (template PAIR: cta_group::2 MMAs with M = 256, each
// CTA staging half of every tap so that B is fetched once per SM pair) is built in and bit-identical, but measured 8 % slower
// (az_trunk_set_cta_pair).  History: a serial version (one 5-tile group of 8 x 9 padded
// positions, the issuing warp also refilling the ring) took 1.33 ms - every tap waited for its own stage to drain.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/az_engine.h"
#include "c4_bitboard.cuh"

namespace {

constexpr int C = 64;            // trunk channels
constexpr int GUARD = 16;        // zero rows before / after (a tap moves the window by up to 9 rows)
constexpr uint32_t ROWB = 16;                        // bytes per row per K group
constexpr uint32_t SBO_A = 128;                      // between 8-row groups
constexpr uint32_t TAP_BYTES = C * C * 2;            // 8192: one tap's [64][64] weights
constexpr uint32_t SBO_W = (C / 8) * 128;            // 1024 (canonical K-major [64][64])
constexpr uint32_t LBO_W = 128;
constexpr uint32_t STEM_TAP_BYTES = C * 16 * 2;      // 2048: stem tap [64][16]
constexpr uint32_t SBO_WS = (16 / 8) * 128;          // 256
constexpr int NHC = 48;                              // head conv channels: 32 policy (1x1, centre tap) + 3 value (3x3) + padding
constexpr uint32_t HEAD_TAP_BYTES = NHC * C * 2;     // 6144
constexpr int NHU = 35;                              // head channels actually used
constexpr int MAX_LAYERS = 24;                       // biases of every layer are staged in shared memory once

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// D = f32; A and B K-major, both bf16 (format 1) or both fp16 (format 0): same rate, same storage, 8 vs 11 significand bits
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
// ---- CTA-pair (cta_group::2) variants: one MMA spans the two SMs of a cluster, 128 rows of A and half of B from each
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of all MMAs issued so far -> the mbarrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a barrier that threads / the tensor core of the OTHER CTA arrive on
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t phase) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAITC_LOOP:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra WAITC_DONE;\n\t"
        "bra WAITC_LOOP;\n\t"
        "WAITC_DONE:\n\t"
        "}" ::"r"(bar), "r"(phase), "r"(0x989680u)  /* suspend-time hint: the warp sleeps in the barrier unit instead of spinning on the issue slots the MMA / producer warps share */
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
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&p);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
// the 16-bit operand format of a kernel instance: bf16 (north_star's default) or fp16 (same tensor-core rate; 8x smaller rounding
// error per operand, which is what brings priors / values within 1e-3 of the fp32 reference `predict`, tests/test_gpu_trunk.py)
template <bool F16>
__device__ __forceinline__ uint32_t pack16(float a, float b) {
    if (F16) {
        __half2 p = __floats2half2_rn(fminf(a, 65504.f), fminf(b, 65504.f));  // saturate instead of overflowing to inf
        return *reinterpret_cast<uint32_t *>(&p);
    }
    return pack_bf16(a, b);
}
template <bool F16>
__device__ __forceinline__ float2 unpack16(uint32_t w) {
    if (F16) return __half22float2(*reinterpret_cast<const __half2 *>(&w));
    return make_float2(bf16_lo(w), bf16_hi(w));
}

#ifdef AZ_TRUNK_CLOCKS
// debug build only: per-layer timestamps of one CTA (issue start / issue end / accumulators ready / epilogue end)
__device__ long long g_clk[4 * 2 * MAX_LAYERS];
#define CLK(kind, l, g) do { if (blockIdx.x == 100 && (threadIdx.x & 31) == 0) g_clk[((kind) * MAX_LAYERS + (l)) * 2 + (g)] = clock64(); } while (0)
#else
#define CLK(kind, l, g) do { } while (0)
#endif

constexpr int THREADS = 320;  // warps 0..7 epilogue, warp 8 MMA issuer, warp 9 weight producer
constexpr int PW = 8, PIX = 56, LEAD = 8;  // pixel-row stride, rows per position, leading zero row of a group
constexpr int GPOS = 4, P = 2 * GPOS;        // positions per group / per CTA
constexpr int GTILES = 2, TILES = 2 * GTILES, GROWS = GTILES * 128;
constexpr int RTOT = TILES * 128 + 2 * GUARD;       // 544 rows per buffer
constexpr uint32_t LBO_A = RTOT * ROWB;             // 8704
constexpr uint32_t BUF_BYTES = (C / 8) * LBO_A;     // 69632
constexpr int NS = 9;
constexpr uint32_t OFF_RING = 2 * BUF_BYTES;
constexpr uint32_t OFF_BIAS = OFF_RING + NS * TAP_BYTES;
constexpr uint32_t OFF_BARS = OFF_BIAS + MAX_LAYERS * C * 4;
constexpr uint32_t SMEM_BYTES = OFF_BARS + (3 * NS + 4) * 8 + 16;  // full[9] empty[9] mma_done[2] epi_done[2] pfull[9], TMEM slot
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(LEAD + GPOS * PIX <= GROWS, "a group must fit its accumulator tiles");

__device__ __forceinline__ bool decode_row(int r, int &pos, int &y, int &x) {
    const int g = r >= GROWS;
    const int rr = r - g * GROWS - LEAD;
    const int p = rr / PIX;
    const int q = rr - p * PIX;
    y = q >> 3;
    x = q & 7;
    pos = g * GPOS + p;
    return rr >= 0 && p < GPOS && y < c4::H && x < c4::W;
}

// epilogue of tiles [t0, t1) of one conv layer (8 warps; warp group `half` takes 32 of the 64 channels)
template <bool F16>
__device__ __forceinline__ void conv_epilogue(uint32_t tmem_base, uint8_t *dst, const uint8_t *skip, const float *bias, int t0, int t1) {
    const uint32_t lane_row = threadIdx.x & 127u;
    const int half = (threadIdx.x >> 7) & 1;
    const uint32_t taddr = tmem_base + ((lane_row & ~31u) << 16) + half * 32;
#pragma unroll 1
    for (int t = t0; t < t1; ++t) {
        const int r = t * 128 + (int)lane_row;
        int pos, y, x;
        const bool valid = decode_row(r, pos, y, x);
        uint8_t *drow = dst + (GUARD + r) * ROWB;
        const uint8_t *srow = skip ? skip + (GUARD + r) * ROWB : nullptr;
        uint32_t v[32];
        tmem_ld32(taddr + t * C, v);
#pragma unroll
        for (int kg = 0; kg < 4; ++kg) {
            const int grp = half * 4 + kg;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[kg * 8 + j]) + bias[grp * 8 + j];
            if (srow) {
                const uint4 s = *reinterpret_cast<const uint4 *>(srow + grp * LBO_A);
                const float2 s0 = unpack16<F16>(s.x), s1 = unpack16<F16>(s.y), s2 = unpack16<F16>(s.z), s3 = unpack16<F16>(s.w);
                f[0] += s0.x; f[1] += s0.y; f[2] += s1.x; f[3] += s1.y;
                f[4] += s2.x; f[5] += s2.y; f[6] += s3.x; f[7] += s3.y;
            }
            uint4 o = make_uint4(0, 0, 0, 0);
            if (valid)
                o = make_uint4(pack16<F16>(fmaxf(f[0], 0.f), fmaxf(f[1], 0.f)), pack16<F16>(fmaxf(f[2], 0.f), fmaxf(f[3], 0.f)),
                               pack16<F16>(fmaxf(f[4], 0.f), fmaxf(f[5], 0.f)), pack16<F16>(fmaxf(f[6], 0.f), fmaxf(f[7], 0.f)));
            *reinterpret_cast<uint4 *>(drow + grp * LBO_A) = o;
        }
    }
}

// MMAs of one (layer, group), fully unrolled: 9 taps x 2 tiles x KSTEPS.  Every operand is warp-uniform; one elected
// lane issues.  Group A waits for each tap's weights; group B finds them there and releases the stage afterwards.
// Before its last tap the warp already waits for what the NEXT (layer, group) needs (`pre_bar`: that group's epi_done;
// `pre_full`: tap 0 of the next layer), while the queued MMAs keep the tensor core busy - the hand-over costs no bubble.
// PAIR: the MMAs are cta_group::2 (M = 256: this CTA's tile and the peer's), a tap's weights are complete when this CTA's
// half (full) and the peer's half (pfull, forwarded by the peer) have landed, and commits arrive in both CTAs.
template <int KSTEPS, bool GROUP_B, bool PAIR>
__device__ __forceinline__ void issue_group(uint32_t full0, uint32_t pfull0, uint32_t empty0, uint32_t parity, bool wait_tap0, uint64_t a_desc,
                                            uint64_t b_desc, uint32_t idesc, uint32_t tmem_tile0, uint32_t pre_bar, uint32_t pre_parity,
                                            bool pre_full) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int shift = (tap / 3 - 1) * PW + (tap % 3 - 1);  // rows; one row = 16 B = one descriptor address unit
        if (!GROUP_B && (tap > 0 || wait_tap0)) {
            mbar_wait(full0 + tap * 8, parity);
            if (PAIR) mbar_wait_cluster(pfull0 + tap * 8, parity);
        }
        if (tap == 8) {
            if (pre_bar) {
                if (PAIR) mbar_wait_cluster(pre_bar, pre_parity);
                else mbar_wait(pre_bar, pre_parity);
            }
            if (pre_full) {
                mbar_wait(full0, parity ^ 1u);
                if (PAIR) mbar_wait_cluster(pfull0, parity ^ 1u);
            }
        }
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
            for (int t = 0; t < GTILES; ++t)
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks) {
                    const uint64_t ad = a_desc + (uint64_t)(int64_t)(shift + t * 128 + ks * (int)(2 * LBO_A >> 4));
                    const uint64_t bd = b_desc + (uint64_t)(tap * (int)(TAP_BYTES >> 4) + ks * (int)(2 * LBO_W >> 4));
                    if (PAIR) umma2(tmem_tile0 + t * C, ad, bd, idesc, (tap | ks) > 0);
                    else umma(tmem_tile0 + t * C, ad, bd, idesc, (tap | ks) > 0);
                }
            if (GROUP_B) {
                if (PAIR) umma_commit2(empty0 + tap * 8);
                else umma_commit(empty0 + tap * 8);
            }
        }
        __syncwarp();
    }
}

template <bool PAIR, bool F16>
__global__ void __launch_bounds__(THREADS, 1)
k_resnet_trunk(const uint64_t *__restrict__ leaf_bb0, const uint64_t *__restrict__ leaf_bb1, const uint8_t *__restrict__ leaf_player,
                   const uint8_t *__restrict__ leaf_status, const int32_t *__restrict__ eval_list, const int32_t *__restrict__ eval_count,
                   long long n_slots, const uint8_t *__restrict__ weights,
                   const float *__restrict__ biases, int num_blocks, __nv_bfloat16 *__restrict__ out,
                   const uint8_t *__restrict__ head_w, const float *__restrict__ head_b, const float *__restrict__ fc_policy_w,
                   const float *__restrict__ fc_policy_b, const float *__restrict__ fc_value_w, const float *__restrict__ fc_value_b,
                   float *__restrict__ logits, float *__restrict__ values) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *buf[2] = {smem, smem + BUF_BYTES};
    float *s_bias = reinterpret_cast<float *>(smem + OFF_BIAS);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BARS);  // full[9] empty[9] mma_done[2] epi_done[2] pfull[9]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BARS + (3 * NS + 4) * 8);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) CLK(0, MAX_LAYERS - 1, 0);
    // eval_list: position j of the batch is slot eval_list[j], j < *eval_count (only the leaves that wait for an evaluation are
    // processed, outputs go to the slots' rows); without it position j is slot j and rows of other slots are computed as zeros
    const long long n = eval_list ? (long long)__ldg(eval_count) : n_slots;
    const int n_conv = 1 + 2 * num_blocks;            // stem + block convs
    const int n_layers = n_conv + (head_w ? 1 : 0);   // + head conv
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NS), ring0 = smem_u32(smem + OFF_RING);
    const uint32_t mma_done0 = smem_u32(bars + 2 * NS), epi_done0 = smem_u32(bars + 2 * NS + 2), pfull0 = smem_u32(bars + 2 * NS + 4);
    // PAIR: two CTAs (one cluster, the two SMs of a TPC) run every MMA together; rank 0 issues them
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    if (warp == 0) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
    }
    if (tid == 32) {
        for (int i = 0; i < 3 * NS + 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + i)));
        // epi_done[g]: one arrival per epilogue warp of the CTA - and, in a pair, of the peer - on the LEADER's barrier
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(epi_done0), "r"(PAIR ? 16u : 8u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(epi_done0 + 8), "r"(PAIR ? 16u : 8u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = tid; i < 2 * BUF_BYTES / 16; i += THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (uint32_t i = tid; i < (uint32_t)n_conv * C; i += THREADS) s_bias[i] = __ldg(biases + i);
    if (head_w)
        for (uint32_t i = tid; i < NHC; i += THREADS) s_bias[n_conv * C + i] = __ldg(head_b + i);
    const uint32_t a0 = smem_u32(buf[0]) + GUARD * ROWB, a1 = smem_u32(buf[1]) + GUARD * ROWB;
    const uint32_t epi_done_lead = PAIR ? mapa(epi_done0, 0u) : epi_done0;  // shared::cluster address in a pair
    uint32_t tmem_base = 0;
    constexpr int WTHREADS = THREADS - 32;  // everyone but the weight producer: it free-runs across batches (single-CTA mode)
    auto batch_sync = [&]() {
        if (PAIR) __syncthreads();
        else asm volatile("bar.sync 2, %0;" ::"n"(WTHREADS) : "memory");
    };
    __syncthreads();  // barriers initialised, tensor memory allocated, buffers zeroed, biases staged

    // Persistent: the CTA keeps its tensor memory, barriers, biases and zero padding and walks over batches of P positions.  Barrier
    // phases simply continue: gl counts every layer the CTA has run (full / empty / mma_done / pfull complete once per layer), ge the
    // conv layers (epi_done is not used by the head layer).  The weight producer runs ahead across the batch boundary.
    const long long n_batches = (n + P - 1) / P + (PAIR ? ((n + P - 1) / P & 1) : 0);
    uint32_t it = 0;
    for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x, ++it) {
    const long long pos0 = batch * P;
    const uint32_t gl0 = it * (uint32_t)n_layers, ge0 = it * (uint32_t)n_conv;
    // The leaf records of this thread's stem rows are requested first, in one round of independent loads, and consumed after the
    // set-up below (first batch: tensor-memory allocation, barrier init, zero fill; later: the previous batch's FC tail): their HBM / L2 latency hides behind it.
    constexpr int ROWS_PER_THREAD = (TILES * 128 + WTHREADS - 1) / WTHREADS;
    uint64_t in_b0[ROWS_PER_THREAD], in_b1[ROWS_PER_THREAD];
    uint32_t in_meta[ROWS_PER_THREAD];  // bit 0: cell of an evaluated leaf, bit 1: side to move, bit 2: row is a board cell, bits 8..: bit index of the cell
#pragma unroll
    for (int k = 0; k < ROWS_PER_THREAD; ++k) {
        const int r = (int)tid + k * WTHREADS;
        int pos, y, x;
        const bool cell = tid < WTHREADS && r < TILES * 128 && decode_row(r, pos, y, x);
        const long long gp = pos0 + pos;
        const bool in = cell && gp < n;
        const long long g = in ? (eval_list ? (long long)__ldg(eval_list + gp) : gp) : 0;
        const uint8_t st = leaf_status[g];
        in_b0[k] = leaf_bb0[g];
        in_b1[k] = leaf_bb1[g];
        in_meta[k] = ((in && st == AZ_LEAF_EVAL) ? 1u : 0u) | (cell ? 4u : 0u) | ((uint32_t)(leaf_player[g] & 1) << 1) | ((uint32_t)(x * c4::STRIDE + y) << 8);
    }
    if (PAIR || warp != 9) {
    batch_sync();  // the previous batch is finished: buffers and tensor memory are this batch's
    // the FC tail of the previous batch used K groups 2..7 of buf[1] as scratch, guard rows included: those must be zero again
    if (it > 0)
        for (uint32_t i = tid; i < 6 * 2 * GUARD; i += WTHREADS) {
            const uint32_t kg = 2 + i / (2 * GUARD), j = i % (2 * GUARD);
            const uint32_t row = j < GUARD ? j : (uint32_t)(TILES * 128 + j);  // rows before / after the tiles
            *reinterpret_cast<uint4 *>(buf[1] + kg * LBO_A + row * ROWB) = make_uint4(0, 0, 0, 0);
        }
    // stem input in buf[1]: channels 0..2 = empty / side to move / opponent (cnn.py:93-95), K group 0.  Cells of slots without
    // an evaluation are written as zeros (the buffer holds the previous batch's activations); K group 1 keeps stale finite
    // values, which the stem's zero weights for channels 8..15 cancel.
#pragma unroll
    for (int k = 0; k < ROWS_PER_THREAD; ++k) {
        const int r = (int)tid + k * WTHREADS;
        if (!(in_meta[k] & 4u)) continue;
        const int pl = (in_meta[k] >> 1) & 1, bit = (int)(in_meta[k] >> 8);
        const uint32_t live = in_meta[k] & 1u;
        const uint32_t s0 = (uint32_t)((in_b0[k] >> bit) & 1ull), s1 = (uint32_t)((in_b1[k] >> bit) & 1ull);
        const uint32_t mine = live * (pl ? s1 : s0), theirs = live * (pl ? s0 : s1), emp = live * (1u - (s0 | s1));
        const uint32_t one = F16 ? 0x3C00u : 0x3F80u;  // 1.0
        *reinterpret_cast<uint4 *>(buf[1] + (GUARD + r) * ROWB) = make_uint4(emp * one | (mine * one) << 16, theirs * one, 0u, 0u);
    }
    fence_async_smem();
    tc_fence_before();
    batch_sync();
    }
    if (PAIR) cluster_sync_all();  // the peer's barriers are initialised and its stem input is written before anything remote happens
    tc_fence_after();
    tmem_base = *tmem_slot;

    if (warp == 9) {
        // ===== weight producer: tap `tap` of layer l into stage `tap`, once group B of layer l-1 has released it =====
        for (int l = 0; l < n_layers; ++l) {
            const uint8_t *w = l == 0 ? weights : (l < n_conv ? weights + 9 * STEM_TAP_BYTES + (size_t)(l - 1) * 9 * TAP_BYTES : head_w);
            const uint32_t tap_bytes = l == 0 ? STEM_TAP_BYTES : (l < n_conv ? TAP_BYTES : HEAD_TAP_BYTES);
            // a pair splits B by output channel: this CTA stages rows [rank * N/2, (rank + 1) * N/2) of every tap - the first /
            // second half of the packed tap, whose 8-row groups are contiguous
            const uint32_t bytes = PAIR ? tap_bytes / 2 : tap_bytes;
#pragma unroll 1
            for (uint32_t tap = 0; tap < 9; ++tap) {
                if (gl0 + l > 0) mbar_wait_cluster(empty0 + tap * 8, (gl0 + (uint32_t)l - 1u) & 1u);
                if (elect_one()) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + tap * 8), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     ring0 + tap * TAP_BYTES),
                                 "l"(w + (size_t)tap * tap_bytes + (size_t)rank * bytes), "r"(bytes), "r"(full0 + tap * 8)
                                 : "memory");
                }
                __syncwarp();
            }
        }
    } else if (warp == 8 && !leader) {
        // ===== peer CTA of a pair: tell the leader when this CTA's half of a tap has landed =====
        const uint32_t pfull_lead = mapa(pfull0, 0u);
        for (int l = 0; l < n_layers; ++l)
#pragma unroll 1
            for (uint32_t tap = 0; tap < 9; ++tap) {
                mbar_wait(full0 + tap * 8, (gl0 + (uint32_t)l) & 1u);
                if (elect_one()) mbar_arrive_cluster(pfull_lead + tap * 8);
                __syncwarp();
            }
    } else if (warp == 8) {
        // ===== MMA issuer (converged; one elected lane issues) =====
        const uint32_t idesc_c = instr_desc(PAIR ? 256 : 128, C, F16), idesc_h = instr_desc(PAIR ? 256 : 128, NHC, F16);
        for (int l = 0; l < n_layers; ++l) {
            const bool is_head = l >= n_conv;
            const uint32_t src = l == 0 ? a1 : (is_head ? a0 : ((l & 1) ? a0 : a1));  // conv1 (odd l) reads x = buf[0]; conv2 reads t = buf[1]
            const uint32_t idesc = is_head ? idesc_h : idesc_c;
            const uint64_t b_desc = smem_desc(ring0, LBO_W, l == 0 ? SBO_WS : SBO_W);
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {
                // this group's input is written and its TMEM tiles are free once epi_done[g] of the previous layer completes: the
                // previous (layer, group) waited for that before its last tap (pre_bar below)
                const uint32_t tile0 = tmem_base + g * GTILES * C;
                const uint64_t a_desc = smem_desc(src + g * GROWS * ROWB, LBO_A, SBO_A);
                const bool last = l + 1 == n_layers;
                const uint32_t pre_bar = g == 0 ? (l > 0 ? epi_done0 + 8 : 0u) : (last ? 0u : epi_done0);
                const uint32_t pre_parity = g == 0 ? (ge0 + (uint32_t)l - 1u) & 1u : (ge0 + (uint32_t)l) & 1u;
                const uint32_t par = (gl0 + (uint32_t)l) & 1u;
                CLK(0, l, g);
                if (l == 0) {
                    if (g == 0) issue_group<1, false, PAIR>(full0, pfull0, empty0, par, true, a_desc, b_desc, idesc, tile0, pre_bar, pre_parity, false);
                    else issue_group<1, true, PAIR>(full0, pfull0, empty0, par, false, a_desc, b_desc, idesc, tile0, pre_bar, pre_parity, !last);
                } else {
                    if (g == 0)
                        issue_group<C / 16, false, PAIR>(full0, pfull0, empty0, par, false, a_desc, b_desc, idesc, tile0, pre_bar, pre_parity, false);
                    else
                        issue_group<C / 16, true, PAIR>(full0, pfull0, empty0, par, false, a_desc, b_desc, idesc, tile0, pre_bar, pre_parity, !last);
                }
                if (elect_one()) {
                    if (PAIR) umma_commit2(mma_done0 + g * 8);
                    else umma_commit(mma_done0 + g * 8);
                }
                __syncwarp();
                CLK(1, l, g);
            }
        }
    } else {
        // ===== epilogue warps (256 threads) =====
        for (int l = 0; l < n_conv; ++l) {
            // stem (l = 0) and conv2 of a block (even l) write x into buf[0]; conv1 (odd l) writes t into buf[1]
            uint8_t *dst = (l & 1) ? buf[1] : buf[0];
            const uint8_t *skip = (l > 0 && !(l & 1)) ? buf[0] : nullptr;
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {
                if (PAIR) mbar_wait_cluster(mma_done0 + g * 8, (gl0 + (uint32_t)l) & 1u);
                else mbar_wait(mma_done0 + g * 8, (gl0 + (uint32_t)l) & 1u);
                tc_fence_after();
                if (warp == 0) CLK(2, l, g);
                conv_epilogue<F16>(tmem_base, dst, skip, s_bias + l * C, g * GTILES, (g + 1) * GTILES);
                // every thread orders its own rows for the tensor core's reads, the warp syncs, one lane arrives for the warp
                tc_fence_before();
                fence_async_smem();
                __syncwarp();
                if ((tid & 31u) == 0) {
                    if (PAIR && !leader) mbar_arrive_cluster(epi_done_lead + g * 8);
                    else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(epi_done0 + g * 8) : "memory");
                }
                if (warp == 0) CLK(3, l, g);
            }
        }
        if (out) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int r = tid; r < TILES * 128; r += 256) {
                int pos, y, x;
                if (!decode_row(r, pos, y, x)) continue;
                const long long gp = pos0 + pos;
                if (gp >= n) continue;
                const long long slot = eval_list ? (long long)__ldg(eval_list + gp) : gp;
                uint4 *o = reinterpret_cast<uint4 *>(out + ((slot * c4::H + y) * c4::W + x) * C);
                const uint8_t *row = buf[0] + (GUARD + r) * ROWB;
#pragma unroll
                for (int grp = 0; grp < C / 8; ++grp) o[grp] = *reinterpret_cast<const uint4 *>(row + grp * LBO_A);
            }
        }
        if (head_w) {
            // [P][35][42] fp32, the Flatten() order of NCHW - behind K groups 0 and 1 of buf[1], whose zero padding the next batch's stem needs
            float *hact = reinterpret_cast<float *>(buf[1] + 2 * LBO_A);
            const float *hb = s_bias + n_conv * C;
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {
                if (PAIR) mbar_wait_cluster(mma_done0 + g * 8, (gl0 + (uint32_t)n_conv) & 1u);
                else mbar_wait(mma_done0 + g * 8, (gl0 + (uint32_t)n_conv) & 1u);
                tc_fence_after();
                if (warp == 0) CLK(2, n_conv, g);
                // warps 0..3: policy channels 0..31; warps 4..7: value channels 32..34
                const int half = (tid >> 7) & 1;
                const uint32_t taddr = tmem_base + (((tid & 127u) & ~31u) << 16) + half * 32;
                for (int t = g * GTILES; t < (g + 1) * GTILES; ++t) {
                    const int r = t * 128 + (int)(tid & 127u);
                    int pos, y, x;
                    const bool valid = decode_row(r, pos, y, x);
                    uint32_t v[32];
                    tmem_ld32(taddr + t * C, v);
                    if (valid) {
                        if (half == 0) {
#pragma unroll
                            for (int c = 0; c < 32; ++c) hact[(pos * NHU + c) * 42 + y * c4::W + x] = fmaxf(__uint_as_float(v[c]) + hb[c], 0.f);
                        } else {
#pragma unroll
                            for (int c = 32; c < NHU; ++c) hact[(pos * NHU + c) * 42 + y * c4::W + x] = fmaxf(__uint_as_float(v[c - 32]) + hb[c], 0.f);
                        }
                    }
                }
                if (warp == 0) CLK(3, n_conv, g);
            }
            tc_fence_before();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // Both FC layers for the CTA's 8 positions at once, in fp32 on CUDA cores: thread t takes inputs k = t, t + 256, ...
            // so every weight is read once per CTA (coalesced) and used for 8 positions; acc[p][j], j = 7 is the value head.
            float acc[P * 8];
#pragma unroll
            for (int i = 0; i < P * 8; ++i) acc[i] = 0.f;
#pragma unroll
            for (int it = 0; it < (32 * 42 + 255) / 256; ++it) {
                const int k = (int)tid + it * 256;
                const bool in = k < 32 * 42;
                const int kk = in ? k : 0;
                float w[7];
#pragma unroll
                for (int j = 0; j < 7; ++j) w[j] = in ? __ldg(fc_policy_w + j * (32 * 42) + kk) : 0.f;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const float xv = hact[p * NHU * 42 + kk];
#pragma unroll
                    for (int j = 0; j < 7; ++j) acc[p * 8 + j] = fmaf(w[j], xv, acc[p * 8 + j]);
                }
            }
            {
                const bool in = tid < 3 * 42;
                const int kk = in ? (int)tid : 0;
                const float w = in ? __ldg(fc_value_w + kk) : 0.f;
#pragma unroll
                for (int p = 0; p < P; ++p) acc[p * 8 + 7] = fmaf(w, hact[p * NHU * 42 + 32 * 42 + kk], acc[p * 8 + 7]);
            }
            // warp reduction by recursive halving: each step exchanges half of the values, so 62 shuffles reduce all 64
            // sums; lane L ends with the totals of indices 2L and 2L + 1
            const uint32_t lane = tid & 31u;
#pragma unroll
            for (int h = 32; h >= 2; h >>= 1) {
                const bool up = (lane & (uint32_t)(h >> 1)) != 0;
#pragma unroll
                for (int i = 0; i < h; ++i) {
                    const float send = up ? acc[i] : acc[i + h];
                    const float keep = up ? acc[i + h] : acc[i];
                    acc[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, h >> 1);
                }
            }
            float *red = reinterpret_cast<float *>(buf[1] + 2 * LBO_A + 48 * 1024);  // [8 warps][64], behind hact (47040 B)
            *reinterpret_cast<float2 *>(red + warp * 64 + 2 * lane) = make_float2(acc[0], acc[1]);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid < P * 8) {
                float sum = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) sum += red[w8 * 64 + tid];
                const int p = (int)tid >> 3, j = (int)tid & 7;
                const long long gp = pos0 + p;
                if (gp < n) {
                    const long long slot = eval_list ? (long long)__ldg(eval_list + gp) : gp;
                    if (j < 7) {
                        logits[slot * 7 + j] = sum + __ldg(fc_policy_b + j);
                    } else {
                        const float v = tanhf(sum + __ldg(fc_value_b));
                        values[slot * 2] = v;
                        values[slot * 2 + 1] = -v;
                    }
                }
            }
        }
    }
    }  // batches
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while the pair's MMAs may still touch it
    if (warp == 0) CLK(1, MAX_LAYERS - 1, 0);
    if (warp == 0) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
    }
}


}  // namespace

extern "C" {

int32_t az_resnet_pipe_launch(az_engine *engine, const az_resnet_desc *d, float *logits, float *values, void *stream);  // csrc/az_resnet_pipe.cu
int32_t az_resnet_wide_launch(az_engine *engine, const az_resnet_desc *d, float *logits, float *values, void *stream);  // csrc/az_resnet_wide.cu

#ifdef AZ_TRUNK_CLOCKS
int32_t az_debug_trunk_clocks(long long *out) {
    return cudaMemcpyFromSymbol(out, g_clk, sizeof(g_clk)) == cudaSuccess ? 0 : 1;
}
#endif

/* bytes of packed weights the trunk kernel expects for `num_blocks` residual blocks */
int64_t az_trunk_weight_bytes(int32_t num_blocks) { return 9ll * STEM_TAP_BYTES + (int64_t)num_blocks * 2 * 9 * TAP_BYTES; }

// 0 (default): one CTA per 8 positions.  1: CTA pairs (cta_group::2 MMAs, M = 256, each CTA stages half of B) - bit-identical,
// measured 8 % SLOWER on B200 for this shape (0.75 vs 0.70 ms at 16384 positions: the per-MMA time does not drop below the single-CTA
// kernel's ~52 cycles and the cross-CTA hand-offs lengthen the epilogues), kept selectable for measurements and tests.
static int g_trunk_cta_pair = 0;

int32_t az_trunk_set_cta_pair(int32_t on) {
    const int prev = g_trunk_cta_pair;
    g_trunk_cta_pair = on ? 1 : 0;
    return prev;
}

static int32_t launch_trunk(az_engine *engine, const void *weights, const float *biases, int32_t num_blocks, void *out,
                            const void *head_w, const float *head_b, const float *fcp_w, const float *fcp_b, const float *fcv_w,
                            const float *fcv_b, float *logits, float *values, void *stream, int32_t fmt = AZ_FMT_BF16) {
    if (fmt != AZ_FMT_BF16 && fmt != AZ_FMT_F16) return AZ_E_INVALID;
    if (!engine || !weights || !biases || num_blocks < 0 || 1 + 2 * num_blocks + 1 > MAX_LAYERS) return AZ_E_INVALID;
    const uint64_t *bb0 = nullptr, *bb1 = nullptr;
    const uint8_t *status = nullptr, *player = nullptr;
    int32_t n = 0;
    if (az_leaf_arrays(engine, &bb0, &bb1, &status, &n) != AZ_OK || az_leaf_players(engine, &player) != AZ_OK || n <= 0) return AZ_E_INVALID;
    // the full net walks the engine's compacted list of leaves that wait for an evaluation; the trunk-only entry point (out != null)
    // produces activations for every slot
    const int32_t *elist = nullptr, *ecount = nullptr;
    if (!out && az_leaf_compact(engine, &elist, &ecount) != AZ_OK) return AZ_E_INVALID;
    static bool attr_set[64] = {false};  // the opt-in to > 48 KB of dynamic shared memory is per device
    const int dev = az_device(engine);
    if (dev < 0 || dev >= 64 || cudaSetDevice(dev) != cudaSuccess) return AZ_E_CUDA;
    if (!attr_set[dev]) {
        if (cudaFuncSetAttribute(k_resnet_trunk<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess) return AZ_E_CUDA;
        if (cudaFuncSetAttribute(k_resnet_trunk<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess) return AZ_E_CUDA;
        if (cudaFuncSetAttribute(k_resnet_trunk<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess) return AZ_E_CUDA;
        if (cudaFuncSetAttribute(k_resnet_trunk<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess) return AZ_E_CUDA;
        attr_set[dev] = true;
    }
    const int pair = g_trunk_cta_pair;
    const long long nn = n;
    const uint8_t *w8 = (const uint8_t *)weights, *hw8 = (const uint8_t *)head_w;
    __nv_bfloat16 *o16 = (__nv_bfloat16 *)out;
    if (pair) {
        // clusters of two CTAs = the two SMs of a TPC; an odd tail CTA gets an empty partner
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(((n + P - 1) / P + 1) / 2 * 2));
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = SMEM_BYTES;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        auto kern = fmt == AZ_FMT_F16 ? k_resnet_trunk<true, true> : k_resnet_trunk<true, false>;
        if (cudaLaunchKernelEx(&cfg, kern, bb0, bb1, player, status, elist, ecount, nn, w8, biases, (int)num_blocks, o16, hw8, head_b, fcp_w, fcp_b,
                               fcv_w, fcv_b, logits, values) != cudaSuccess)
            return AZ_E_CUDA;
    } else {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return AZ_E_CUDA;
        const int batches = (n + P - 1) / P;
        auto kern = fmt == AZ_FMT_F16 ? k_resnet_trunk<false, true> : k_resnet_trunk<false, false>;
        kern<<<batches < sms ? batches : sms, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(bb0, bb1, player, status, elist, ecount, nn, w8, biases, num_blocks, o16, hw8, head_b,
                                                                                            fcp_w, fcp_b, fcv_w, fcv_b, logits, values);
    }
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_E_CUDA;
}

/* ResNet trunk (stem + num_blocks residual blocks, 64 channels, BatchNorm folded) on the leaves chosen by the last
 * az_select_leaves of `engine`: out[slot][6][7][64] bf16 (NHWC, row 0 = bottom).  `weights`: bf16 MMA operands packed as
 * alphazero-implementation_b200/models.py:pack_trunk_weights does; `biases`: fp32 [1 + 2*num_blocks][64]. */
int32_t az_trunk_forward_leaves(az_engine *engine, const void *weights, const float *biases, int32_t num_blocks, void *out,
                                void *stream) {
    if (!out) return AZ_E_INVALID;
    return launch_trunk(engine, weights, biases, num_blocks, out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
}

/* The whole net (trunk + policy / value heads, resnet.py:30-103 with tanh and the [v, -v] value convention of cnn.py:73) in the
 * same kernel: logits [E][7] and values [E][2] fp32, the inputs of az_expand_backup.  head_conv_w: 9 taps x [48][64] bf16
 * (rows 0..31 policy conv1x1 in the centre tap, rows 32..34 value conv3x3), head_conv_b [48]; fc_*: fp32 nn.Linear layout. */
int32_t az_resnet_forward_leaves(az_engine *engine, const void *weights, const float *biases, int32_t num_blocks,
                                 const void *head_conv_w, const float *head_conv_b, const float *fc_policy_w, const float *fc_policy_b,
                                 const float *fc_value_w, const float *fc_value_b, float *logits, float *values, void *stream) {
    if (!head_conv_w || !head_conv_b || !fc_policy_w || !fc_policy_b || !fc_value_w || !fc_value_b || !logits || !values) return AZ_E_INVALID;
    return launch_trunk(engine, weights, biases, num_blocks, nullptr, head_conv_w, head_conv_b, fc_policy_w, fc_policy_b, fc_value_w,
                        fc_value_b, logits, values, stream);
}

/* The same with an explicit descriptor: operand format (bf16 / fp16) and channel count.  Replaces `Connect4Model.predict`'s
 * forward (models/games/connect4/model.py:19-43) for the ResNet-style net on the leaves of the last az_select_leaves. */
int32_t az_resnet_forward_leaves_v2(az_engine *engine, const az_resnet_desc *d, float *logits, float *values, void *stream) {
    if (!d || !logits || !values) return AZ_E_INVALID;
    if (!d->head_conv_w || !d->head_conv_b || !d->fc_policy_w || !d->fc_policy_b || !d->fc_value_w || !d->fc_value_b) return AZ_E_INVALID;
    // default: the layer-pipelined kernel (csrc/az_resnet_pipe.cu) for 64 and 128 channels; variant 1 = this file's ping-pong kernel
    if (d->num_channels == 64 && d->variant == 4) return az_resnet_wide_launch(engine, d, logits, values, stream);  // filter rows fused (N = 192)
    if (d->num_channels == 128 || (d->num_channels == 64 && d->variant != 1)) return az_resnet_pipe_launch(engine, d, logits, values, stream);  // variants 0, 2, 3
    if (d->num_channels != 64 || d->variant != 1) return AZ_E_INVALID;
    return launch_trunk(engine, d->trunk_w, d->trunk_b, d->num_blocks, nullptr, d->head_conv_w, d->head_conv_b, d->fc_policy_w, d->fc_policy_b,
                        d->fc_value_w, d->fc_value_b, logits, values, stream, d->operand_format);
}

}  // extern "C"
