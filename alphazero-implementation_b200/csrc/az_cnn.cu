// The reference's CNNModel (models/games/connect4/cnn.py:8-75: 3 x (conv3x3 + BatchNorm + ReLU) 3 -> 64 -> 128 -> 256,
// Flatten, Linear 10752 -> 512 + ReLU (+ Dropout, identity in eval), policy Linear 512 -> 7, value Linear 512 -> 1 + tanh,
// returned as [v, -v]) for the leaves of the search, BatchNorm folded, as two hand-written tcgen05 kernels:
//
//  k_cnn_conv  the three convolutions.  Same implicit GEMM as csrc/az_resnet_pipe.cu (pixel = GEMM row, channels = K, activation
//      buffers K-group-major so that a filter tap is the same buffer with the MMA descriptor moved by 8 dy + dx rows, compact
//      7 x 8 padding, 4 positions = 2 accumulator tiles per CTA) and the same layer pipeline through double-buffered tensor memory
//      with per-K-chunk hand-over barriers, over a fixed layer table:
//        L0  3(16) -> 64    1 piece  [9][ 64][16] 18 KB   accumulators: set 0
//        L1  64    -> 128   4 pieces [9][128][16] 36 KB                 set 1   (K chunk ks waits for L0's epilogue chunk ks)
//        L2a 128   -> 128   8 pieces (output channels   0..127)         set 0   (waits for L1's epilogue chunks)
//        L2b 128   -> 128   8 pieces (output channels 128..255)         set 1   (same input, already complete)
//      The last layer's epilogue writes straight to HBM in the layout the second kernel's A operand wants: tiles of
//      [128 positions][32 k] in the MMA's canonical K-major core-matrix order, k = pixel * 256 + channel.
//  k_cnn_fc    the 10752 -> 512 layer as a GEMM with M = 128 positions per CTA, N = 512 (all 512 tensor-memory columns), K streamed
//      in chunks of 32 (A tile 8 KB + weight tile 32 KB per stage, 4 stages); epilogue: bias + ReLU in fp32 and both heads on
//      CUDA cores straight from the accumulators (8 outputs per row), tanh, [v, -v].
// Only the leaves that wait for an evaluation are processed (the engine's compacted list); outputs go to the slots' rows.
#include <stdint.h>
#include <stdio.h>

#include "../../include/az_engine.h"
#include "c4_bitboard.cuh"
#include "tcgen05.cuh"

namespace {

using namespace tc05;

// ---- geometry shared with csrc/az_resnet_pipe.cu
constexpr int GUARD = 16;
constexpr uint32_t ROWB = 16;
constexpr int PW = 8, PIX = 56, LEAD = 8, TPOS = 2, TILES = 2, POS = TPOS * TILES, ROWS = TILES * 128, RTOT = ROWS + 2 * GUARD;
constexpr uint32_t LBO_A = RTOT * ROWB, SBO_A = 128, LBO_W = 128, SBO_W = 256;
constexpr int THREADS = 320, WTHREADS = 288;

// ---- conv kernel
constexpr int NL = 4;                                     // L0, L1, L2a, L2b
constexpr uint32_t BUF_IN = 2 * LBO_A, BUF_1 = 8 * LBO_A, BUF_2 = 16 * LBO_A;  // 3(16), 64 and 128 channels
constexpr uint32_t PIECE0 = 9 * 64 * 16 * 2, PIECE = 9 * 128 * 16 * 2;          // 18432, 36864
constexpr int NS = 2;
constexpr uint32_t OFF_B1 = BUF_IN, OFF_B2 = OFF_B1 + BUF_1, OFF_RING = OFF_B2 + BUF_2;
constexpr uint32_t OFF_BIAS = OFF_RING + NS * PIECE;
constexpr int NBIAS = 64 + 128 + 256;
constexpr uint32_t OFF_BARS = OFF_BIAS + NBIAS * 4;
constexpr int NBARS = 2 * NS + NL + 8 + 1;                // full[NS] empty[NS] mma_done[NL] chunk[8] drain0
constexpr uint32_t CONV_SMEM = OFF_BARS + NBARS * 8 + 16;
static_assert(CONV_SMEM <= 232448, "shared memory budget");
constexpr int64_t CONV_W_BYTES = (int64_t)PIECE0 + 4 * PIECE + 16 * PIECE;

// ---- FC kernel
constexpr int FC_K = 42 * 256, FC_N = 512, KC = 32, NCHUNK = FC_K / KC;  // 336 chunks of 32
constexpr uint32_t A_TILE = 128 * KC * 2, B_TILE = FC_N * KC * 2;        // 8192, 32768
constexpr uint32_t SBO_FC = (KC / 8) * 128;                              // 512: next 8-row group of a [rows][32] tile
constexpr int FC_NS = 4;
constexpr uint32_t FC_STAGE = A_TILE + B_TILE;
constexpr uint32_t FC_OFF_WH = FC_NS * FC_STAGE;                        // head weights, transposed to [512][8] fp32
constexpr uint32_t FC_OFF_B1 = FC_OFF_WH + 8 * FC_N * 4;                // fc bias [512]
constexpr uint32_t FC_OFF_RED = FC_OFF_B1 + FC_N * 4;                   // [128][8] partial sums of the upper column half
constexpr uint32_t FC_OFF_BARS = FC_OFF_RED + 128 * 8 * 4;
constexpr uint32_t FC_SMEM = FC_OFF_BARS + (2 * FC_NS + 1) * 8 + 16;
static_assert(FC_SMEM <= 232448, "shared memory budget");

__device__ __forceinline__ bool decode_row(int r, int &pos, int &y, int &x) {
    const int tile = r >> 7;
    const int rr = (r & 127) - LEAD;
    const int p = rr >= 0 ? rr / PIX : 0;
    const int q = rr - p * PIX;
    y = q >> 3;
    x = q & 7;
    pos = tile * TPOS + p;
    return rr >= 0 && p < TPOS && y < c4::H && x < c4::W;
}

// byte offset of 8 consecutive k (16 bytes) of row `j` (compacted position index) in the FC kernel's A operand: tiles
// [j / 128][k / 32] of 8 KB, canonical K-major inside
__device__ __forceinline__ size_t fc_a_offset(long long j, int k) {
    const long long blk = j >> 7;
    const int row = (int)(j & 127);
    return ((size_t)blk * NCHUNK + (size_t)(k >> 5)) * A_TILE + (size_t)(row >> 3) * SBO_FC + (size_t)((k & 31) >> 3) * 128 + (size_t)(row & 7) * 16;
}

template <bool F16>
__global__ void __launch_bounds__(THREADS, 1)
k_cnn_conv(const uint64_t *__restrict__ leaf_bb0, const uint64_t *__restrict__ leaf_bb1, const uint8_t *__restrict__ leaf_player,
           const uint8_t *__restrict__ leaf_status, const int32_t *__restrict__ eval_list, const int32_t *__restrict__ eval_count, long long n_slots,
           const uint8_t *__restrict__ weights, const float *__restrict__ biases, uint8_t *__restrict__ fc_a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *bufIn = smem, *buf1 = smem + OFF_B1, *buf2 = smem + OFF_B2;
    float *s_bias = reinterpret_cast<float *>(smem + OFF_BIAS);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BARS);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BARS + NBARS * 8);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const long long n = eval_list ? (long long)__ldg(eval_count) : n_slots;
    // mma_done: one barrier PER LAYER (each completes once per batch).  With a single one the issuer, which needs nothing from the
    // epilogue warps to run L0 of the next batch, could complete two layers (L2b, next L0) before a late epilogue warp looks - and a
    // parity wait that misses two phase flips waits forever.
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NS), mma_done = smem_u32(bars + 2 * NS), chunk0 = smem_u32(bars + 2 * NS + NL);
    const uint32_t drain0 = chunk0 + 8 * 8;  // accumulator set 0 is free again: L2a's epilogue is complete (one arrival per epilogue warp)
    const uint32_t ring0 = smem_u32(smem + OFF_RING);
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512u);
    if (tid == 32) {
        for (int i = 0; i < 2 * NS + NL; ++i) mbar_init(smem_u32(bars + i), 1u);
        for (int i = 0; i < 9; ++i) mbar_init(chunk0 + i * 8, 8u);  // chunk[8] and drain0: one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = tid; i < OFF_RING / 16; i += THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (uint32_t i = tid; i < NBIAS; i += THREADS) s_bias[i] = __ldg(biases + i);
    const uint32_t aIn = smem_u32(bufIn) + GUARD * ROWB, a1 = smem_u32(buf1) + GUARD * ROWB, a2 = smem_u32(buf2) + GUARD * ROWB;
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // The batches of a CTA form ONE pipeline without CTA-wide barriers: the issuer warp stages the next batch's input planes itself
    // (its own buffer) while it waits, the first layers of the next batch run while the epilogue warps still write the previous
    // batch's conv3 output to HBM; the only extra hand-over is `drain0` (accumulator set 0 free again).
    auto stage_input = [&](long long first_pos) {  // issuer warp: 8 rows per lane
        uint64_t b0[POS], b1[POS];
        uint32_t meta[POS];  // bit 0: leaf waits for an evaluation, bit 1: side to move
#pragma unroll
        for (int p = 0; p < POS; ++p) {
            const long long gp = first_pos + p;
            const bool in = gp < n;
            const long long slot = in ? (eval_list ? (long long)__ldg(eval_list + gp) : gp) : 0;
            b0[p] = leaf_bb0[slot];
            b1[p] = leaf_bb1[slot];
            meta[p] = ((in && leaf_status[slot] == AZ_LEAF_EVAL) ? 1u : 0u) | ((uint32_t)(leaf_player[slot] & 1) << 1);
        }
        const uint32_t one = F16 ? 0x3C00u : 0x3F80u;
#pragma unroll
        for (int k = 0; k < ROWS / 32; ++k) {
            const int r = (int)lane + 32 * k;
            int pos, y, x;
            if (!decode_row(r, pos, y, x)) continue;
            uint64_t c0 = b0[0], c1 = b1[0];
            uint32_t m = meta[0];
#pragma unroll
            for (int p = 1; p < POS; ++p)
                if (pos == p) { c0 = b0[p]; c1 = b1[p]; m = meta[p]; }
            const int bit = x * c4::STRIDE + y, pl = (int)(m >> 1) & 1;
            const uint32_t live = m & 1u;
            const uint32_t s0 = (uint32_t)((c0 >> bit) & 1ull), s1 = (uint32_t)((c1 >> bit) & 1ull);
            const uint32_t mine = live * (pl ? s1 : s0), theirs = live * (pl ? s0 : s1), emp = live * (1u - (s0 | s1));
            // channels 0..2 = empty / side to move / opponent (cnn.py:93-95); the buffer's second K group stays zero
            *reinterpret_cast<uint4 *>(bufIn + (GUARD + r) * ROWB) = make_uint4(emp * one | (mine * one) << 16, theirs * one, 0u, 0u);
        }
        fence_async_smem();
        __syncwarp();
    };
    if (warp == 8 && (long long)blockIdx.x < (n + POS - 1) / POS) stage_input((long long)blockIdx.x * POS);

    const long long n_batches = (n + POS - 1) / POS;
    uint32_t it = 0, g = 0;  // g: weight pieces produced / consumed so far
    for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x, ++it) {
        const long long pos0 = batch * POS;
        if (warp == 9) {
            // ===== weight producer =====
            const uint8_t *src = weights;
#pragma unroll 1
            for (int i = 0; i < 1 + 4 + 16; ++i, ++g) {
                const uint32_t bytes = i == 0 ? PIECE0 : PIECE;
                const uint32_t st = g % NS;
                if (g >= NS) mbar_wait(empty0 + st * 8, ((g / NS) - 1u) & 1u);
                if (elect_one()) bulk_load(ring0 + st * PIECE, src, bytes, full0 + st * 8);
                __syncwarp();
                src += bytes;
            }
        } else if (warp == 8) {
            // ===== MMA issuer =====
            if (it > 0) {  // set 0 (L0's accumulators) was L2a's in the previous batch: its epilogue must have drained it
                mbar_wait(drain0, (it - 1u) & 1u);
                fence_after();
            }
#pragma unroll 1
            for (int l = 0; l < NL; ++l) {
                const uint32_t src = l == 0 ? aIn : (l == 1 ? a1 : a2);
                const int N = l == 0 ? 64 : 128, ksteps = l == 0 ? 1 : (l == 1 ? 4 : 8);
                const uint32_t idesc = instr_desc(128, N, F16);
                const uint32_t acc = tmem_base + (uint32_t)(l & 1) * 256u;
                const uint64_t a_desc = smem_desc(src, LBO_A, SBO_A);
                const uint32_t tap_units = (uint32_t)(N * 16 * 2) >> 4;
#pragma unroll 1
                for (int ks = 0; ks < ksteps; ++ks, ++g) {
                    // chunk[c] completes twice per batch for c < 4 (after L0's and after L1's epilogue) and once for c >= 4 (L1's)
                    if (l == 1) mbar_wait(chunk0 + ks * 8, 0u);
                    if (l == 2) mbar_wait(chunk0 + ks * 8, ks < 4 ? 1u : (it & 1u));
                    const uint32_t st = g % NS;
                    mbar_wait(full0 + st * 8, (g / NS) & 1u);
                    fence_after();
                    if (elect_one()) {
                        const uint64_t bd = smem_desc(ring0 + st * PIECE, LBO_W, SBO_W);
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const int shift = (tap / 3 - 1) * PW + (tap % 3 - 1);
#pragma unroll
                            for (int t = 0; t < TILES; ++t)
                                umma(acc + t * N, a_desc + (uint64_t)(int64_t)(shift + t * 128 + ks * (int)(2 * LBO_A >> 4)), bd + (uint64_t)(tap * tap_units), idesc,
                                     (ks | tap) > 0);
                        }
                        umma_commit(empty0 + st * 8);
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(mma_done + l * 8);
                __syncwarp();
            }
            // the next batch's input: L0 of this batch (the only reader of the input buffer) completed long ago
            if (batch + gridDim.x < n_batches) stage_input((batch + gridDim.x) * POS);
        } else {
            // ===== epilogue warps: thread = (tile, row) =====
            const int tile = (int)(warp >> 2);
            const int r = tile * 128 + (int)((warp & 3u) * 32u + lane);
            int pos, y, x;
            const bool valid = decode_row(r, pos, y, x);
            const uint32_t lane_addr = tmem_base + (((warp & 3u) * 32u) << 16);
            const uint32_t row_off = (GUARD + r) * ROWB;
            const long long gp = pos0 + pos;
            const int pix = y * c4::W + x;
#pragma unroll 1
            for (int l = 0; l < NL; ++l) {
                const int N = l == 0 ? 64 : 128, chunks = N / 16;
                const float *bias = s_bias + (l == 0 ? 0 : (l == 1 ? 64 : (l == 2 ? 192 : 320)));
                const uint32_t acc = lane_addr + (uint32_t)(l & 1) * 256u + (uint32_t)(tile * N);
                uint8_t *dst = l == 0 ? buf1 : buf2;
                mbar_wait(mma_done + l * 8, it & 1u);
                fence_after();
                uint32_t va[16], vb[16];
                auto chunk = [&](const uint32_t (&v)[16], int c) {
                    uint4 o[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        // packed pair adds (FADD2), ReLU + rounding in one F2FP.RELU
                        const float4 b0 = *reinterpret_cast<const float4 *>(bias + c * 16 + h * 8), b1 = *reinterpret_cast<const float4 *>(bias + c * 16 + h * 8 + 4);
                        o[h] = make_uint4(
                            pack16_relu<F16>(fadd2(make_float2(__uint_as_float(v[h * 8]), __uint_as_float(v[h * 8 + 1])), make_float2(b0.x, b0.y))),
                            pack16_relu<F16>(fadd2(make_float2(__uint_as_float(v[h * 8 + 2]), __uint_as_float(v[h * 8 + 3])), make_float2(b0.z, b0.w))),
                            pack16_relu<F16>(fadd2(make_float2(__uint_as_float(v[h * 8 + 4]), __uint_as_float(v[h * 8 + 5])), make_float2(b1.x, b1.y))),
                            pack16_relu<F16>(fadd2(make_float2(__uint_as_float(v[h * 8 + 6]), __uint_as_float(v[h * 8 + 7])), make_float2(b1.z, b1.w))));
                    }
                    if (l < 2) {
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            *reinterpret_cast<uint4 *>(dst + (2 * c + h) * LBO_A + row_off) = valid ? o[h] : make_uint4(0, 0, 0, 0);
                        fence_before();
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(chunk0 + c * 8);
                    } else if (valid && gp < n) {
                        // conv3 output -> the FC kernel's A operand: k = pixel * 256 + channel, channel = (l - 2) * 128 + 16 c + 8 h
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            *reinterpret_cast<uint4 *>(fc_a + fc_a_offset(gp, pix * 256 + (l - 2) * 128 + 16 * c + 8 * h)) = o[h];
                    }
                };
                tmem_ld16_issue(acc, va);
#pragma unroll 1
                for (int c = 0; c < chunks; c += 2) {
                    tmem_ld_wait();
                    tmem_ld16_issue(acc + (c + 1) * 16, vb);
                    chunk(va, c);
                    tmem_ld_wait();
                    if (c + 2 < chunks) tmem_ld16_issue(acc + (c + 2) * 16, va);
                    chunk(vb, c + 1);
                }
                if (l == 2) {  // set 0 is drained: the next batch's L0 may overwrite it
                    fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(drain0);
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512u);
}

// 10752 -> 512 (+ ReLU) -> {7 logits, tanh value}: one CTA per 128 compacted positions
template <bool F16>
__global__ void __launch_bounds__(THREADS, 1)
k_cnn_fc(const uint8_t *__restrict__ fc_a, const uint8_t *__restrict__ fc_w, const float *__restrict__ fc_b, const float *__restrict__ head_w,
         const float *__restrict__ head_b, const int32_t *__restrict__ eval_list, const int32_t *__restrict__ eval_count, long long n_slots,
         float *__restrict__ logits, float *__restrict__ values) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float *s_wh = reinterpret_cast<float *>(smem + FC_OFF_WH), *s_b1 = reinterpret_cast<float *>(smem + FC_OFF_B1);
    float *s_red = reinterpret_cast<float *>(smem + FC_OFF_RED);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + FC_OFF_BARS);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + FC_OFF_BARS + (2 * FC_NS + 1) * 8);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const long long n = eval_list ? (long long)__ldg(eval_count) : n_slots;
    const long long row0 = (long long)blockIdx.x * 128;
    if (row0 >= n) return;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + FC_NS), done = smem_u32(bars + 2 * FC_NS);
    const uint32_t ring0 = smem_u32(smem);
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512u);
    if (tid == 32) {
        for (int i = 0; i < 2 * FC_NS + 1; ++i) mbar_init(smem_u32(bars + i), 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = tid; i < 8 * FC_N; i += THREADS) s_wh[(i % FC_N) * 8 + i / FC_N] = __ldg(head_w + i);  // [8][512] -> [512][8]: two 16-byte loads per column
    for (uint32_t i = tid; i < FC_N; i += THREADS) s_b1[i] = __ldg(fc_b + i);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 9) {
        // ===== producer: per K chunk the A tile of this CTA's 128 positions and the weight tile [512][32] =====
        const uint8_t *a_src = fc_a + (size_t)blockIdx.x * NCHUNK * A_TILE;
#pragma unroll 1
        for (uint32_t g = 0; g < (uint32_t)NCHUNK; ++g) {
            const uint32_t st = g % FC_NS;
            if (g >= FC_NS) mbar_wait(empty0 + st * 8, ((g / FC_NS) - 1u) & 1u);
            if (elect_one()) {
                const uint32_t bar = full0 + st * 8, dst = ring0 + st * FC_STAGE;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(FC_STAGE) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                             "l"(a_src + (size_t)g * A_TILE), "r"(A_TILE), "r"(bar)
                             : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + A_TILE),
                             "l"(fc_w + (size_t)g * B_TILE), "r"(B_TILE), "r"(bar)
                             : "memory");
            }
            __syncwarp();
        }
    } else if (warp == 8) {
        // ===== MMA issuer: D[128 x 512] += A[128 x 32] . W[512 x 32]^T per chunk, as 2 K steps x 2 halves of N = 256 =====
        const uint32_t idesc = instr_desc(128, 256, F16);
#pragma unroll 1
        for (uint32_t g = 0; g < (uint32_t)NCHUNK; ++g) {
            const uint32_t st = g % FC_NS;
            mbar_wait(full0 + st * 8, (g / FC_NS) & 1u);
            fence_after();
            if (elect_one()) {
                const uint32_t sa = ring0 + st * FC_STAGE, sb = sa + A_TILE;
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks)
#pragma unroll
                    for (int half = 0; half < 2; ++half)
                        umma(tmem_base + half * 256, smem_desc(sa + ks * 256, 128, SBO_FC), smem_desc(sb + half * (256 / 8) * SBO_FC + ks * 256, 128, SBO_FC), idesc,
                             (g | (uint32_t)ks) > 0);
                umma_commit(empty0 + st * 8);
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(done);
        __syncwarp();
    } else {
        // ===== epilogue: thread = (row, column half); h = ReLU(acc + b) in fp32, heads as 8 dot products per row =====
        const uint32_t row = tid & 127u, half = tid >> 7;
        const uint32_t taddr = tmem_base + ((row & ~31u) << 16) + half * 256;
        mbar_wait(done, 0u);
        fence_after();
        float out[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] = 0.f;
        uint32_t va[32], vb[32];
        auto block = [&](const uint32_t (&v)[32], int cb) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int col = (int)half * 256 + cb * 32 + i;
                const float h = fmaxf(__uint_as_float(v[i]) + s_b1[col], 0.f);
                const float4 w0 = *reinterpret_cast<const float4 *>(s_wh + col * 8), w1 = *reinterpret_cast<const float4 *>(s_wh + col * 8 + 4);
                out[0] = fmaf(h, w0.x, out[0]); out[1] = fmaf(h, w0.y, out[1]); out[2] = fmaf(h, w0.z, out[2]); out[3] = fmaf(h, w0.w, out[3]);
                out[4] = fmaf(h, w1.x, out[4]); out[5] = fmaf(h, w1.y, out[5]); out[6] = fmaf(h, w1.z, out[6]); out[7] = fmaf(h, w1.w, out[7]);
            }
        };
        tmem_ld32_issue(taddr, va);
#pragma unroll 1
        for (int cb = 0; cb < 8; cb += 2) {
            tmem_ld_wait();
            tmem_ld32_issue(taddr + (cb + 1) * 32, vb);
            block(va, cb);
            tmem_ld_wait();
            if (cb + 2 < 8) tmem_ld32_issue(taddr + (cb + 2) * 32, va);
            block(vb, cb + 1);
        }
        if (half == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s_red[row * 8 + j] = out[j];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (half == 0) {
            const long long gp = row0 + row;
            if (gp < n) {
                const long long slot = eval_list ? (long long)__ldg(eval_list + gp) : gp;
#pragma unroll
                for (int j = 0; j < 7; ++j) logits[slot * 7 + j] = (out[j] + s_red[row * 8 + j]) + __ldg(head_b + j);
                const float v = tanhf((out[7] + s_red[row * 8 + 7]) + __ldg(head_b + 7));
                values[slot * 2] = v;  // cnn.py:73: cat([value, -value])
                values[slot * 2 + 1] = -v;
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512u);
}

}  // namespace

extern "C" {

/* bytes of the packed convolution weights (models.py:pack_cnn_weights) and of the workspace between the two kernels for n slots */
int64_t az_cnn_conv_weight_bytes(void) { return CONV_W_BYTES; }
int64_t az_cnn_fc_weight_bytes(void) { return (int64_t)NCHUNK * B_TILE; }
int64_t az_cnn_workspace_bytes(int64_t n) { return ((n + 127) / 128) * (int64_t)NCHUNK * A_TILE; }

/* CNNModel.forward (cnn.py:52-75) on the leaves of the last selection: logits [E][7], values [E][2] = [v, -v] */
int32_t az_cnn_forward_leaves(az_engine *engine, const az_cnn_desc *d, float *logits, float *values, void *stream) {
    if (!engine || !d || !logits || !values || !d->conv_w || !d->conv_b || !d->fc_w || !d->fc_b || !d->head_w || !d->head_b || !d->workspace) return AZ_E_INVALID;
    if (d->operand_format != AZ_FMT_BF16 && d->operand_format != AZ_FMT_F16) return AZ_E_INVALID;
    const uint64_t *bb0 = nullptr, *bb1 = nullptr;
    const uint8_t *status = nullptr, *player = nullptr;
    const int32_t *elist = nullptr, *ecount = nullptr;
    int32_t n = 0;
    if (az_leaf_arrays(engine, &bb0, &bb1, &status, &n) != AZ_OK || az_leaf_players(engine, &player) != AZ_OK ||
        az_leaf_compact(engine, &elist, &ecount) != AZ_OK || n <= 0)
        return AZ_E_INVALID;
    if (d->workspace_bytes < az_cnn_workspace_bytes(n)) return AZ_E_INVALID;
    static bool attr_set[64] = {false};
    const int dev = az_device(engine);
    if (dev < 0 || dev >= 64 || cudaSetDevice(dev) != cudaSuccess) return AZ_E_CUDA;
    if (!attr_set[dev]) {
        if (cudaFuncSetAttribute(k_cnn_conv<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CONV_SMEM) != cudaSuccess) return AZ_E_CUDA;
        if (cudaFuncSetAttribute(k_cnn_conv<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CONV_SMEM) != cudaSuccess) return AZ_E_CUDA;
        if (cudaFuncSetAttribute(k_cnn_fc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FC_SMEM) != cudaSuccess) return AZ_E_CUDA;
        if (cudaFuncSetAttribute(k_cnn_fc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FC_SMEM) != cudaSuccess) return AZ_E_CUDA;
        attr_set[dev] = true;
    }
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return AZ_E_CUDA;
    const bool f16 = d->operand_format == AZ_FMT_F16;
    const int batches = (n + POS - 1) / POS;
    auto conv = f16 ? k_cnn_conv<true> : k_cnn_conv<false>;
    conv<<<batches < sms ? batches : sms, THREADS, CONV_SMEM, (cudaStream_t)stream>>>(bb0, bb1, player, status, elist, ecount, (long long)n,
                                                                                      (const uint8_t *)d->conv_w, d->conv_b, (uint8_t *)d->workspace);
    auto fc = f16 ? k_cnn_fc<true> : k_cnn_fc<false>;
    fc<<<(n + 127) / 128, THREADS, FC_SMEM, (cudaStream_t)stream>>>((const uint8_t *)d->workspace, (const uint8_t *)d->fc_w, d->fc_b, d->head_w, d->head_b,
                                                                    elist, ecount, (long long)n, logits, values);
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_E_CUDA;
}

}  // extern "C"
