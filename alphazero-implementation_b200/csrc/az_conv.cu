// Fused ResNet-style trunk on the 5th-generation tensor cores (tcgen05 + TMEM): stem conv3x3 (3 -> 64) and any
// number of residual blocks (2 x conv3x3 64 -> 64, skip, ReLU) of the reference's ResNet
// (src/alphazero_simple/resnet.py:13-53, BatchNorm folded) for a batch of leaf positions in ONE kernel; the
// activations of a position never leave the SM between layers.
//
// Convolution as implicit GEMM without im2col.  Every position is laid out as a zero-padded 8 x 9 grid of
// "pixels" (6 x 7 cells plus a border), pixel = one GEMM row, channels = K.  The activation buffer is stored
// K-group-major: for every group of 8 channels, all rows back to back at 16 bytes per row.  In the MMA's
// K-major no-swizzle shared-memory descriptor this is "stride between 8-row groups = 128 B, stride between
// K-adjacent core matrices = rows * 16 B", so the A operand of filter tap (dy, dx) is THE SAME buffer with the
// start address moved by (9*dy + dx) rows: nine taps = nine descriptor offsets, no data movement.  The zero
// border supplies the padding; outputs computed for border pixels are discarded (written back as zeros, which
// keeps the border zero for the next layer).  8 positions = 576 pixels = 5 accumulator tiles of 128 rows x 64
// fp32 columns in tensor memory.  Per layer one thread issues 9 taps x 5 tiles x 4 K-steps tcgen05.mma; the
// tap weights ([64 out][64 in] bf16, 8 KB, packed once per weight update) stream through a 4-stage ring filled
// by bulk async copies.  Epilogue (8 warps: two per 32 accumulator lanes, half the channels each): tcgen05.ld, bias
// (+ skip) + ReLU in fp32, round to bf16, write the next layer's A operand.  After the last block the policy conv1x1 and the
// value conv3x3 run as one 48-channel conv layer (the 1x1 weights in the centre tap) and the two small FC layers on CUDA cores.
// Measured (ncu, 16384 positions, 4 blocks): 1.35 ms trunk, tensor pipe 24-28 % active; a 128x64x16 MMA takes ~65 cycles, not the
// 32-cycle floor, because it fetches 6 KB of operands from shared memory (~96 B/clk): with 64 output channels the conv MMAs are
// shared-memory-operand-bound.  A software-pipelined variant (issuer warp + two position groups so that one group's epilogue
// overlaps the other's MMAs) was built and is bit-identical but ran SLOWER (2.4-2.6 ms): the MMAs slow down to ~97 ns each when
// the epilogue warps' shared-memory traffic competes with the operand fetch; it was removed.  The leaf gather is fused in: the stem's input planes
// (empty / side to move / opponent, cnn.py:93-95) are built from the engine's leaf bitboards.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/az_engine.h"
#include "c4_bitboard.cuh"

namespace {

constexpr int C = 64;            // trunk channels
constexpr int P = 8;             // positions per CTA
constexpr int PW = 9, PH = 8;    // padded grid
constexpr int PIX = PW * PH;     // 72 rows per position
constexpr int ROWS = P * PIX;    // 576 real rows
constexpr int TILES = 5;         // accumulator tiles of 128 rows (640 >= 576)
constexpr int GUARD = 16;        // rows before / after (a tap moves the window by up to 10 rows)
constexpr int RTOT = TILES * 128 + 2 * GUARD;       // 672 rows per buffer
constexpr uint32_t ROWB = 16;                        // bytes per row per K group
constexpr uint32_t LBO_A = RTOT * ROWB;              // 10752: between K groups (8 channels)
constexpr uint32_t SBO_A = 128;                      // between 8-row groups
constexpr uint32_t BUF_BYTES = (C / 8) * LBO_A;      // 86016
constexpr uint32_t TAP_BYTES = C * C * 2;            // 8192: one tap's [64][64] weights
constexpr uint32_t SBO_W = (C / 8) * 128;            // 1024 (canonical K-major [64][64])
constexpr uint32_t LBO_W = 128;
constexpr uint32_t STEM_TAP_BYTES = C * 16 * 2;      // 2048: stem tap [64][16]
constexpr uint32_t SBO_WS = (16 / 8) * 128;          // 256
constexpr int NS = 4;                                // weight ring stages (6 measured no faster: the MMAs, not the copies, pace a layer)
constexpr int NHC = 48;                              // head conv channels: 32 policy (1x1, centre tap) + 3 value (3x3) + padding
constexpr uint32_t HEAD_TAP_BYTES = NHC * C * 2;     // 6144
constexpr int NHU = 35;                              // head channels actually used
constexpr int MAX_LAYERS = 24;                       // biases of every layer are staged in shared memory once
constexpr uint32_t OFF_RING = 2 * BUF_BYTES;
constexpr uint32_t OFF_BIAS = OFF_RING + NS * TAP_BYTES;
constexpr uint32_t OFF_BARS = OFF_BIAS + MAX_LAYERS * C * 4;
constexpr uint32_t SMEM_BYTES = OFF_BARS + (2 * NS + 1) * 8 + 16;
constexpr int THREADS = 256;  // 8 warps: warps w and w+4 share the 32 accumulator lanes 32*(w%4).., each takes half the channels

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
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
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&p);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

struct Pipe {
    uint32_t full[NS], empty[NS], done, stage[NS];
    uint32_t g;  // taps consumed so far (stage = g % NS, use = g / NS)
};
__device__ __forceinline__ void pipe_load(const Pipe &p, uint32_t gj, const uint8_t *src, uint32_t bytes) {
    const uint32_t st = gj % NS;
    if (gj >= NS) mbar_wait(p.empty[st], ((gj / NS) - 1u) & 1u);
    if (elect_one()) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(p.full[st]), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(p.stage[st]),
                     "l"(src), "r"(bytes), "r"(p.full[st])
                     : "memory");
    }
    __syncwarp();
}

// row index inside a CTA -> is it a real board cell, and which (position, y, x)
__device__ __forceinline__ bool decode_row(int r, int &pos, int &y, int &x) {
    pos = r / PIX;
    const int q = r - pos * PIX;
    const int py = q / PW, px = q - py * PW;
    y = py - 1;
    x = px - 1;
    return r < ROWS && py >= 1 && py <= c4::H && px >= 1 && px <= c4::W;
}

// One conv layer's MMAs (warp 0, converged; one elected lane issues): 9 taps x TILES x ksteps.  A = `a_addr` (row 0 of the buffer, after the guard).
__device__ __forceinline__ void conv_mmas(Pipe &p, const uint8_t *w, uint32_t tap_bytes, uint32_t ksteps, uint32_t sbo_w,
                                          uint32_t a_addr, uint32_t tmem_base, bool prefetched, int n_out = C) {
    const uint32_t idesc = instr_desc(128, n_out);
    if (!prefetched)
        for (uint32_t i = 0; i < NS; ++i) pipe_load(p, p.g + i, w + (size_t)i * tap_bytes, tap_bytes);
    for (uint32_t tap = 0; tap < 9; ++tap) {
        const uint32_t gi = p.g + tap, st = gi % NS;
        const int dy = (int)(tap / 3) - 1, dx = (int)(tap % 3) - 1;
        const uint32_t a_tap = a_addr + (uint32_t)((dy * PW + dx) * (int)ROWB);
        mbar_wait(p.full[st], (gi / NS) & 1u);
        tc_fence_after();
        // descriptors differ only in the start-address field (low 14 bits, 16-byte units): build once, add offsets
        const uint64_t a_base = smem_desc(a_tap, LBO_A, SBO_A), b_base = smem_desc(p.stage[st], LBO_W, sbo_w);
        const uint32_t acc0 = tap > 0;
        if (elect_one()) {
            if (ksteps == 1) {
#pragma unroll
                for (uint32_t t = 0; t < TILES; ++t) umma(tmem_base + t * C, a_base + t * (128 * ROWB >> 4), b_base, idesc, acc0);
            } else {
#pragma unroll
                for (uint32_t t = 0; t < TILES; ++t)
#pragma unroll
                    for (uint32_t ks = 0; ks < C / 16; ++ks)
                        umma(tmem_base + t * C, a_base + t * (128 * ROWB >> 4) + ks * (2 * LBO_A >> 4), b_base + ks * (2 * LBO_W >> 4), idesc,
                             acc0 | (ks > 0));
            }
            umma_commit(p.empty[st]);
        }
        __syncwarp();
        if (tap + NS < 9) pipe_load(p, gi + NS, w + (size_t)(tap + NS) * tap_bytes, tap_bytes);
    }
    if (elect_one()) umma_commit(p.done);
    __syncwarp();
    p.g += 9;
}

// Epilogue of one layer: accumulators -> (+ bias, + skip) -> ReLU -> bf16 -> destination buffer (K-group-major);
// border / padding rows are written as zeros.  `skip` (may be null) is the residual input buffer.
__device__ __forceinline__ void conv_epilogue(uint32_t tmem_base, uint8_t *dst, const uint8_t *skip, const float *bias) {
    const uint32_t lane_row = threadIdx.x & 127u;
    const int half = threadIdx.x >> 7;  // which 32 of the 64 channels this warp group handles
    const uint32_t taddr = tmem_base + ((lane_row & ~31u) << 16) + half * 32;
#pragma unroll 1
    for (int t = 0; t < TILES; ++t) {
        const int r = t * 128 + (int)lane_row;
        int pos, y, x;
        const bool valid = decode_row(r, pos, y, x);
        uint8_t *drow = dst + (GUARD + r) * ROWB;
        const uint8_t *srow = skip ? skip + (GUARD + r) * ROWB : nullptr;
        uint32_t v[32];
        tmem_ld32(taddr + t * C, v);
#pragma unroll
        for (int kg = 0; kg < 4; ++kg) {
            const int grp = half * 4 + kg;  // channel group of 8
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[kg * 8 + j]) + bias[grp * 8 + j];
            if (srow) {
                const uint4 s = *reinterpret_cast<const uint4 *>(srow + grp * LBO_A);
                f[0] += bf16_lo(s.x); f[1] += bf16_hi(s.x); f[2] += bf16_lo(s.y); f[3] += bf16_hi(s.y);
                f[4] += bf16_lo(s.z); f[5] += bf16_hi(s.z); f[6] += bf16_lo(s.w); f[7] += bf16_hi(s.w);
            }
            uint4 o = make_uint4(0, 0, 0, 0);
            if (valid)
                o = make_uint4(pack_bf16(fmaxf(f[0], 0.f), fmaxf(f[1], 0.f)), pack_bf16(fmaxf(f[2], 0.f), fmaxf(f[3], 0.f)),
                               pack_bf16(fmaxf(f[4], 0.f), fmaxf(f[5], 0.f)), pack_bf16(fmaxf(f[6], 0.f), fmaxf(f[7], 0.f)));
            *reinterpret_cast<uint4 *>(drow + grp * LBO_A) = o;
        }
    }
}

// weights: [stem: 9 taps x [64][16]] [layer 1: 9 taps x [64][64]] ... all bf16 canonical; biases: [L][64] fp32
__global__ void __launch_bounds__(THREADS, 1)
k_resnet_trunk(const uint64_t *__restrict__ leaf_bb0, const uint64_t *__restrict__ leaf_bb1, const uint8_t *__restrict__ leaf_player,
               const uint8_t *__restrict__ leaf_status, long long n, const uint8_t *__restrict__ weights,
               const float *__restrict__ biases, int num_blocks, __nv_bfloat16 *__restrict__ out /*[n][6][7][64] or null*/,
               const uint8_t *__restrict__ head_w /*9 taps x [48][64] or null*/, const float *__restrict__ head_b /*[48]*/,
               const float *__restrict__ fc_policy_w /*[7][1344]*/, const float *__restrict__ fc_policy_b,
               const float *__restrict__ fc_value_w /*[126]*/, const float *__restrict__ fc_value_b,
               float *__restrict__ logits /*[n][7]*/, float *__restrict__ values /*[n][2]*/) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *buf[2] = {smem, smem + BUF_BYTES};
    float *s_bias = reinterpret_cast<float *>(smem + OFF_BIAS);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BARS);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BARS + (2 * NS + 1) * 8);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const long long pos0 = (long long)blockIdx.x * P;
    Pipe p;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        p.full[i] = smem_u32(bars + i);
        p.empty[i] = smem_u32(bars + NS + i);
        p.stage[i] = smem_u32(smem + OFF_RING) + i * TAP_BYTES;
    }
    p.done = smem_u32(bars + 2 * NS);
    p.g = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        for (int i = 0; i < 2 * NS + 1; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (warp == 0)
        for (uint32_t i = 0; i < NS; ++i) pipe_load(p, i, weights + (size_t)i * STEM_TAP_BYTES, STEM_TAP_BYTES);
    // zero both activation buffers (borders, guards and the unused K groups of the stem input must be zero)
    for (uint32_t i = tid; i < 2 * BUF_BYTES / 16; i += THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    const int n_layers = 1 + 2 * num_blocks;
    for (uint32_t i = tid; i < (uint32_t)n_layers * C; i += THREADS) s_bias[i] = __ldg(biases + i);
    if (head_w)
        for (uint32_t i = tid; i < NHC; i += THREADS) s_bias[n_layers * C + i] = __ldg(head_b + i);
    __syncthreads();
    // stem input in buf[1]: channels 0..2 = empty / side to move / opponent (cnn.py:93-95), K group 0
    for (int r = tid; r < ROWS; r += THREADS) {
        int pos, y, x;
        if (!decode_row(r, pos, y, x)) continue;
        const long long gp = pos0 + pos;
        if (gp >= n || leaf_status[gp] != AZ_LEAF_EVAL) continue;
        const uint64_t b0 = leaf_bb0[gp], b1 = leaf_bb1[gp];
        const int pl = leaf_player[gp] & 1;
        const int bit = x * c4::STRIDE + y;
        const uint32_t s0 = (uint32_t)((b0 >> bit) & 1ull), s1 = (uint32_t)((b1 >> bit) & 1ull);
        const uint32_t mine = pl ? s1 : s0, theirs = pl ? s0 : s1, empty = 1u - (s0 | s1);
        const uint32_t one = 0x3F80u;
        *reinterpret_cast<uint4 *>(buf[1] + (GUARD + r) * ROWB) = make_uint4(empty * one | (mine * one) << 16, theirs * one, 0u, 0u);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a0 = smem_u32(buf[0]) + GUARD * ROWB, a1 = smem_u32(buf[1]) + GUARD * ROWB;
    uint32_t done_phase = 0;
    const uint8_t *w_layer = weights + 9 * STEM_TAP_BYTES;  // first trunk layer's weights

    // ---- stem: buf[1] (16 input channels, 3 used) -> buf[0]
    if (warp == 0) {
        conv_mmas(p, weights, STEM_TAP_BYTES, 1, SBO_WS, a1, tmem_base, true);
        if (num_blocks > 0)
            for (uint32_t i = 0; i < NS; ++i) pipe_load(p, p.g + i, w_layer + (size_t)i * TAP_BYTES, TAP_BYTES);
        else if (head_w)
            for (uint32_t i = 0; i < NS; ++i) pipe_load(p, p.g + i, head_w + (size_t)i * HEAD_TAP_BYTES, HEAD_TAP_BYTES);
    }
    mbar_wait(p.done, done_phase);
    done_phase ^= 1;
    tc_fence_after();
    conv_epilogue(tmem_base, buf[0], nullptr, s_bias);

    // ---- residual blocks: x in buf[0]; t = relu(conv1(x)) -> buf[1]; x = relu(conv2(t) + x) -> buf[0]
    for (int blk = 0; blk < num_blocks; ++blk) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int layer = 1 + 2 * blk + half;
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
            if (warp == 0) {
                conv_mmas(p, w_layer, TAP_BYTES, C / 16, SBO_W, half == 0 ? a0 : a1, tmem_base, true);
                const bool last = (blk == num_blocks - 1 && half == 1);
                if (!last)
                    for (uint32_t i = 0; i < NS; ++i) pipe_load(p, p.g + i, w_layer + 9 * TAP_BYTES + (size_t)i * TAP_BYTES, TAP_BYTES);
                else if (head_w)
                    for (uint32_t i = 0; i < NS; ++i) pipe_load(p, p.g + i, head_w + (size_t)i * HEAD_TAP_BYTES, HEAD_TAP_BYTES);
            }
            w_layer += 9 * TAP_BYTES;
            mbar_wait(p.done, done_phase);
            done_phase ^= 1;
            tc_fence_after();
            if (half == 0) conv_epilogue(tmem_base, buf[1], nullptr, s_bias + layer * C);
            else conv_epilogue(tmem_base, buf[0], buf[0], s_bias + layer * C);
        }
    }
    __syncthreads();
    // ---- trunk output, NHWC bf16 (optional)
    if (out) {
        for (int r = tid; r < ROWS; r += THREADS) {
            int pos, y, x;
            if (!decode_row(r, pos, y, x)) continue;
            const long long gp = pos0 + pos;
            if (gp >= n) continue;
            uint4 *o = reinterpret_cast<uint4 *>(out + ((gp * c4::H + y) * c4::W + x) * C);
            const uint8_t *row = buf[0] + (GUARD + r) * ROWB;
#pragma unroll
            for (int grp = 0; grp < C / 8; ++grp) o[grp] = *reinterpret_cast<const uint4 *>(row + grp * LBO_A);
        }
    }
    // ---- heads (resnet.py:56-72): policy conv1x1 -> 32 and value conv3x3 -> 3 as ONE 48-channel conv layer (the 1x1
    // weights sit in the centre tap), ReLU, then the two small fully connected layers on CUDA cores in fp32
    if (head_w) {
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (warp == 0) conv_mmas(p, head_w, HEAD_TAP_BYTES, C / 16, SBO_W, a0, tmem_base, true, NHC);
        mbar_wait(p.done, done_phase);
        done_phase ^= 1;
        tc_fence_after();
        float *hact = reinterpret_cast<float *>(buf[1]);  // [P][35][42] fp32, the Flatten() order of NCHW
        const float *hb = s_bias + n_layers * C;
        if (tid < 128) {
            const uint32_t taddr = tmem_base + ((tid & ~31u) << 16);
#pragma unroll 1
            for (int t = 0; t < TILES; ++t) {
                const int r = t * 128 + (int)tid;
                int pos, y, x;
                const bool valid = decode_row(r, pos, y, x);
                uint32_t v[32];
                tmem_ld32(taddr + t * C, v);  // policy channels 0..31
                if (valid) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) hact[(pos * NHU + c) * 42 + y * c4::W + x] = fmaxf(__uint_as_float(v[c]) + hb[c], 0.f);
                }
                tmem_ld32(taddr + t * C + 32, v);  // value channels 32..34 (+ padding)
                if (valid) {
#pragma unroll
                    for (int c = 32; c < NHU; ++c) hact[(pos * NHU + c) * 42 + y * c4::W + x] = fmaxf(__uint_as_float(v[c - 32]) + hb[c], 0.f);
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        // warp = position; lanes stride over the 1344 (policy) / 126 (value) inputs with coalesced weight reads and
        // seven independent accumulators, then a shuffle reduction
        {
            const int pos = warp, lane = tid & 31;
            const long long gp = pos0 + pos;
            const float *a = hact + pos * NHU * 42;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
            for (int i = lane; i < 32 * 42; i += 32) {
                const float x = a[i];
#pragma unroll
                for (int j = 0; j < 7; ++j) acc[j] = fmaf(__ldg(fc_policy_w + j * (32 * 42) + i), x, acc[j]);
            }
            for (int i = lane; i < 3 * 42; i += 32) acc[7] = fmaf(__ldg(fc_value_w + i), a[32 * 42 + i], acc[7]);
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) acc[j] += __shfl_xor_sync(0xFFFFFFFFu, acc[j], off);
            if (gp < n) {
                if (lane < 7) {
                    float v = acc[0];
#pragma unroll
                    for (int j = 1; j < 7; ++j) v = lane == j ? acc[j] : v;
                    logits[gp * 7 + lane] = v + __ldg(fc_policy_b + lane);
                } else if (lane == 7) {
                    const float v = tanhf(acc[7] + __ldg(fc_value_b));
                    values[gp * 2] = v;
                    values[gp * 2 + 1] = -v;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

}  // namespace

extern "C" {

/* bytes of packed weights the trunk kernel expects for `num_blocks` residual blocks */
int64_t az_trunk_weight_bytes(int32_t num_blocks) { return 9ll * STEM_TAP_BYTES + (int64_t)num_blocks * 2 * 9 * TAP_BYTES; }

static int32_t launch_trunk(az_engine *engine, const void *weights, const float *biases, int32_t num_blocks, void *out,
                            const void *head_w, const float *head_b, const float *fcp_w, const float *fcp_b, const float *fcv_w,
                            const float *fcv_b, float *logits, float *values, void *stream) {
    if (!engine || !weights || !biases || num_blocks < 0 || 1 + 2 * num_blocks + 1 > MAX_LAYERS) return AZ_E_INVALID;
    const uint64_t *bb0 = nullptr, *bb1 = nullptr;
    const uint8_t *status = nullptr, *player = nullptr;
    int32_t n = 0;
    if (az_leaf_arrays(engine, &bb0, &bb1, &status, &n) != AZ_OK || az_leaf_players(engine, &player) != AZ_OK || n <= 0) return AZ_E_INVALID;
    static bool attr_set[64] = {false};  // the opt-in to > 48 KB of dynamic shared memory is per device
    const int dev = az_device(engine);
    if (dev < 0 || dev >= 64 || cudaSetDevice(dev) != cudaSuccess) return AZ_E_CUDA;
    if (!attr_set[dev]) {
        if (cudaFuncSetAttribute(k_resnet_trunk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess) return AZ_E_CUDA;
        attr_set[dev] = true;
    }
    const int blocks = (n + P - 1) / P;
    k_resnet_trunk<<<blocks, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(bb0, bb1, player, status, n, (const uint8_t *)weights, biases,
                                                                          num_blocks, (__nv_bfloat16 *)out, (const uint8_t *)head_w, head_b,
                                                                          fcp_w, fcp_b, fcv_w, fcv_b, logits, values);
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

}  // extern "C"
