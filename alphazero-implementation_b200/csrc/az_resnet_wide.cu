// ResNet(board, 7, num_res_blocks, 64) (src/alphazero_simple/resnet.py:30-103) behind the main package's Model API, BatchNorm
// folded, as ONE tcgen05 kernel for the leaves of the search - the 64-channel net with the three taps of a filter row fused
// into one MMA (N = 3 x 64 = 192).
//
// Why: in the implicit-GEMM formulation of csrc/az_conv.cu / az_resnet_pipe.cu (pixel = GEMM row, channels = K, a filter tap =
// the activation buffer with the MMA descriptor moved by 8*dy + dx rows) a 64-channel layer issues 128x64x16 MMAs: 4 KB of A +
// 2 KB of B from shared memory for 32 cycles of math.  The operand crossbar moves 128 B/clk, so each takes 48 cycles
// (scripts/ubench/mma_pair.cu) and the whole kernel runs at the speed of shared memory.  The A tile of a tap is
// used for 64 output columns only.  Here the three taps (dy, -1), (dy, 0), (dy, +1) of a filter row share ONE A tile - the
// activations moved by 8*dy rows only:
//     D[p][(dx, co)] = sum_dy sum_ci X[p + 8 dy][ci] * W[dy][dx][ci][co]                      (128 x 192 x 16 MMAs)
//     out[q][co]     = D[q - 1][(-1, co)] + D[q][(0, co)] + D[q + 1][(+1, co)]                (epilogue)
// A 128x192x16 MMA is 96 cycles of math; measured 112-115, because the tensor core fetches the A tile once per 128 output columns
// (14 KB of operands per MMA at 128 B/clk) - still 12 MMAs per tile-layer instead of 36 of 48 cycles.  The row shift of the epilogue is
// a warp shuffle: a 128-row tile is 2 positions x 7 x 8 pixel rows, pixel column = row & 7, and column 7 is the shared zero column -
// its accumulator rows are exact zeros - so lane 0 of a warp (column 0) takes its left neighbour from lane 31 of the same warp
// (column 7 of another row: zero, as the true neighbour), and lane 31 itself is never a board cell.  The neighbours' partial sums
// travel as fp16 pairs (one shuffle for two channels; no measurable change of the outputs, AZ_WIDE_SHFL16).
//
// Schedule: one CTA per SM, 2 tiles (4 positions); an accumulator tile is 192 fp32 columns, two of them = 384 of the 512
// columns, so the tiles ping-pong: while the eight epilogue warps drain tile 0 of layer l (tcgen05.ld, shuffles, bias (+ skip),
// ReLU, round to 16 bits, write the next layer's A operand), the tensor core runs tile 1 of layer l, then tile 0 of layer
// l + 1, ...  Both tiles use every weight piece, so a layer's weights are streamed once per CTA: a piece = one K chunk =
// [3 dy][192 = (dx, co)][16 ci] = 18 KB - byte for byte the [9 taps][64][16] piece of the layer-pipelined kernel
// (models.py:pack_trunk_weights_pipe), so both kernels take the same packed trunk weights - through a 6-stage ring (1.5 layers ahead).
// The heads (policy conv1x1 -> 32, value conv3x3 -> 3) are one more layer of the same form with 64 MMA columns (policy 32 + 3 filter
// columns x 8; models.py:pack_head_weights(wide=True)) accumulating in the remaining 2 x 64 tensor-memory columns, so at a batch
// boundary the next batch's stem MMAs follow the heads' without waiting for the head epilogue, and the epilogue warps take that stem
// BEFORE the heads.  The next batch's stem input is written by a stager warp when tcgen05.commit reports the tile's last trunk
// layer complete (the t buffer is free by then); both fully connected layers of the heads run on an FC warp.
// Roles: warps 0..7 epilogue (thread = one pixel row of a tile; warps 0..3 take output channels 0..31, warps 4..7 32..63),
// warp 8 issues the MMAs (one elected lane), warp 9 streams the weights, warp 10 stages the stem inputs, warp 11 runs the FC
// layers - the last two keep batch boundaries off the epilogue warps, which (with the operand crossbar) bound this kernel.
// Measurements, the variants that lost, and the timeline of a CTA: DESIGN.md 3b, profiles/r02_k_resnet_wide_*.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/az_engine.h"
#include "c4_bitboard.cuh"
#include "tcgen05.cuh"

namespace {

using namespace tc05;

constexpr int C = 64, KS = C / 16, KG = C / 8;
constexpr int GUARD = 16;              // zero rows before / after the tiles (a filter row moves the window by 8 rows)
constexpr uint32_t ROWB = 16;
constexpr int PW = 8, PIX = 56, LEAD = 8;  // pixel-row stride, rows per position, leading zero rows of a tile
constexpr int TPOS = 2, TILES = 2, POS = TPOS * TILES, ROWS = TILES * 128;
constexpr int RTOT = ROWS + 2 * GUARD;
constexpr uint32_t SBO_A = 128, LBO_A = RTOT * ROWB;   // next 8-row group, next K group
constexpr uint32_t BUF_BYTES = KG * LBO_A;
constexpr uint32_t LBO_W = 128, SBO_W = 256;           // canonical K-major [N][16]
constexpr int NHC = 48, NHU = 35;      // head biases: 32 policy (1x1, centre tap) + 3 value (3x3) + padding
constexpr int NH = 64;                 // head MMA columns: 32 policy + 3 filter columns x 8 (3 value channels + padding) + 8 padding
constexpr uint32_t DY_BYTES = 3 * C * 16 * 2, PIECE_BYTES = 3 * DY_BYTES;              // [192][16], three filter rows
constexpr uint32_t HEAD_DY_BYTES = NH * 16 * 2, HEAD_PIECE_BYTES = 3 * HEAD_DY_BYTES;  // [64][16], three filter rows
constexpr int NS = 6;                  // ring stages
#ifndef WIDE_EW
#define WIDE_EW 8
#endif
constexpr int EW = WIDE_EW, THREADS = (EW + 4) * 32;   // + issuer, weight producer, stager, FC warp
constexpr int CPW = C / (EW / 4), NCC = CPW / 16;   // output channels per epilogue warp (4 warps cover the 128 rows), 16-channel chunks of them
static_assert(EW == 8 || EW == 16, "epilogue warps: 2 or 4 per 32 accumulator lanes");
constexpr int MAX_CONV = 23;           // 11 blocks
constexpr uint32_t TILE_COLS = 192;    // tensor-memory columns of an accumulator tile: columns 0..383 = the two tiles,
constexpr uint32_t HEAD_COLS = 384;    // 384..511 = the two tiles' head accumulators (64 columns each)
constexpr uint32_t OFF_RING = 2 * BUF_BYTES;
constexpr uint32_t OFF_HACT = OFF_RING + NS * PIECE_BYTES;
constexpr uint32_t HACT_BYTES = POS * NHU * 42 * 4;
constexpr uint32_t OFF_RED = OFF_HACT + HACT_BYTES;
constexpr uint32_t OFF_BIAS = OFF_RED;
constexpr uint32_t OFF_BARS = OFF_BIAS + (MAX_CONV * C + NHC) * 4;
constexpr int NBARS = 2 * NS + 14;     // full[NS] empty[NS] mma_done[2] epi_done[2] stage_go[2] stem_ready[2] hact_ready hact_free head_mma[2] head_epi[2]
constexpr uint32_t SMEM_BYTES = OFF_BARS + NBARS * 8 + 16;
#ifdef WIDE_TRACE
constexpr uint32_t SMEM_LAUNCH = SMEM_BYTES + 3 * 340 * 8;
#else
constexpr uint32_t SMEM_LAUNCH = SMEM_BYTES;
#endif
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(LEAD + TPOS * PIX <= 128, "a tile's positions must fit its 128 rows");
static_assert(OFF_HACT % 16 == 0 && OFF_BIAS % 16 == 0 && OFF_BARS % 8 == 0, "alignment");

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

// the leaf record behind one pixel row of a batch (requested early, consumed when the stem input is staged)
struct RowRec {
    uint64_t b0, b1;
    uint32_t meta;  // bit 0: row of a position that exists and waits for an evaluation, bit 1: side to move, bit 2: board cell, bits 8..: bit index
};

#ifdef WIDE_TRACE
// timeline of CTA 0 (timing study only): each traced warp keeps (code, clock) pairs in shared memory (a store and a clock read
// per event - nothing that waits), copied out when the kernel ends.  code = layer * 100 + tile * 10 + phase
constexpr int TRACE_WARPS = 3, TRACE_CAP = 340;
__device__ unsigned int g_trace[TRACE_WARPS * TRACE_CAP * 2];
#define TRACE(slot, l, t, phase)                                                                 \
    do {                                                                                          \
        if (blockIdx.x == 0 && lane == 0 && tr_n < TRACE_CAP) {                                   \
            s_trace[((slot) * TRACE_CAP + tr_n) * 2] = (unsigned int)((l) * 100 + (t) * 10 + (phase)); \
            s_trace[((slot) * TRACE_CAP + tr_n) * 2 + 1] = (unsigned int)clock64();               \
            ++tr_n;                                                                               \
        }                                                                                         \
    } while (0)
#else
#define TRACE(slot, l, t, phase)
#endif

template <bool F16, bool S16>
__global__ void __launch_bounds__(THREADS, 1)
k_resnet_wide(const uint64_t *__restrict__ leaf_bb0, const uint64_t *__restrict__ leaf_bb1, const uint8_t *__restrict__ leaf_player,
              const uint8_t *__restrict__ leaf_status, const int32_t *__restrict__ eval_list, const int32_t *__restrict__ eval_count,
              long long n_slots, const uint8_t *__restrict__ weights, const float *__restrict__ biases, int num_blocks,
              const uint8_t *__restrict__ head_w, const float *__restrict__ head_b, const float *__restrict__ fc_policy_w,
              const float *__restrict__ fc_policy_b, const float *__restrict__ fc_value_w, const float *__restrict__ fc_value_b,
              float *__restrict__ logits, float *__restrict__ values, unsigned long long *__restrict__ timing, int timing_cap) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *bufX = smem, *bufT = smem + BUF_BYTES;
    float *hact = reinterpret_cast<float *>(smem + OFF_HACT);
    float *s_bias = reinterpret_cast<float *>(smem + OFF_BIAS);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BARS);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BARS + NBARS * 8);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
#ifdef WIDE_TRACE
    unsigned int *s_trace = reinterpret_cast<unsigned int *>(smem + SMEM_BYTES);
    int tr_n = 0;
#endif
    // position j of a batch is slot eval_list[j], j < *eval_count: only the leaves that wait for an evaluation are processed
    const long long n = eval_list ? (long long)__ldg(eval_count) : n_slots;
    const int n_conv = 1 + 2 * num_blocks;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NS), mma_done0 = smem_u32(bars + 2 * NS), epi_done0 = smem_u32(bars + 2 * NS + 2);
    const uint32_t stage_go0 = epi_done0 + 16, stem_ready0 = stage_go0 + 16, hact_ready = stem_ready0 + 16, hact_free = hact_ready + 8;
    const uint32_t head_mma0 = hact_free + 8, head_epi0 = head_mma0 + 16;
    const uint32_t ring0 = smem_u32(smem + OFF_RING);
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
    if (tid == 32) {
        for (int i = 0; i < 2 * NS + 2; ++i) mbar_init(smem_u32(bars + i), 1u);
        for (int i = 0; i < 2; ++i) mbar_init(epi_done0 + i * 8, (uint32_t)EW);  // one arrival per epilogue warp
        for (int i = 0; i < 4; ++i) mbar_init(stage_go0 + i * 8, 1u);            // stage_go[2] (tensor core), stem_ready[2] (stager warp)
        mbar_init(hact_ready, (uint32_t)EW);
        mbar_init(hact_free, 1u);
        for (int i = 0; i < 2; ++i) mbar_init(head_mma0 + i * 8, 1u);
        for (int i = 0; i < 2; ++i) mbar_init(head_epi0 + i * 8, (uint32_t)EW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = tid; i < 2 * BUF_BYTES / 16; i += THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (uint32_t i = tid; i < (uint32_t)n_conv * C; i += THREADS) s_bias[i] = __ldg(biases + i);
    for (uint32_t i = tid; i < NHC; i += THREADS) s_bias[n_conv * C + i] = __ldg(head_b + i);
    const uint32_t aX = smem_u32(bufX) + GUARD * ROWB, aT = smem_u32(bufT) + GUARD * ROWB;
    const long long n_batches = (n + POS - 1) / POS;
    // measurement hook (az_resnet_wide_set_timing): first start / last end of every launch on the device's global timer
    unsigned long long t_slot = 0;
    if (timing && tid == 0) {
        t_slot = 4 + 2 * (*reinterpret_cast<volatile unsigned long long *>(timing) % (unsigned long long)timing_cap);
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        atomicMin(timing + t_slot, now);
    }

    // the leaf record of pixel row r (0..255) of batch `batch`
    auto load_rec = [&](long long batch, int r) {
        RowRec rec;
        int pos, y, x;
        const bool cell = decode_row(r, pos, y, x);
        const long long gp = batch * POS + pos;
        const bool in = cell && gp < n;
        const long long slot = in ? (eval_list ? (long long)__ldg(eval_list + gp) : gp) : 0;
        const bool live = in && leaf_status[slot] == AZ_LEAF_EVAL;
        rec.b0 = leaf_bb0[slot];
        rec.b1 = leaf_bb1[slot];
        rec.meta = (live ? 1u : 0u) | ((uint32_t)(leaf_player[slot] & 1) << 1) | (cell ? 4u : 0u) | ((uint32_t)(x * c4::STRIDE + y) << 8);
        return rec;
    };
    // stem input in t, K group 0: channels 0..2 = empty / side to move / opponent (cnn.py:93-95).  K group 1 keeps stale finite
    // activations, which the stem's zero weights for channels 8..15 cancel.
    auto stage_row = [&](const RowRec &rec, int r) {
        if (!(rec.meta & 4u)) return;
        const int pl = (rec.meta >> 1) & 1, bit = (int)(rec.meta >> 8);
        const uint32_t live = rec.meta & 1u;
        const uint32_t s0 = (uint32_t)((rec.b0 >> bit) & 1ull), s1 = (uint32_t)((rec.b1 >> bit) & 1ull);
        const uint32_t mine = live * (pl ? s1 : s0), theirs = live * (pl ? s0 : s1), emp = live * (1u - (s0 | s1));
        const uint32_t one = F16 ? 0x3C00u : 0x3F80u;
        *reinterpret_cast<uint4 *>(bufT + (GUARD + r) * ROWB) = make_uint4(emp * one | (mine * one) << 16, theirs * one, 0u, 0u);
    };
    fence_async_smem();
    fence_before();
    __syncthreads();  // barriers initialised, tensor memory allocated, buffers zeroed, biases staged
    fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Order of the work at a batch boundary (per tile t): ... last trunk layer of batch k, heads of batch k, stem of batch k + 1,
    // conv1 of batch k + 1 ...  The heads accumulate in their own tensor-memory columns, so the next stem's MMAs do not wait for the
    // head epilogue, and the epilogue warps take the next stem BEFORE the heads: the tensor core gets conv1 of the next batch as
    // early as possible.  Trunk tile-layers (stem included) of a tile are counted by idx = batch * n_conv + l.
    if (warp == EW + 1) {
        // ===== weight producer: every piece in the order the issuer consumes them: stem (first batch), then per batch the trunk
        // pieces, the 4 head pieces and the next batch's stem piece =====
        uint32_t g = 0;
        auto piece = [&](const uint8_t *src, uint32_t bytes) {
            const uint32_t st = g % NS;
            if (g >= NS) mbar_wait(empty0 + st * 8, ((g / NS) - 1u) & 1u);  // both tiles' MMAs of the previous use have read the stage
            if (elect_one()) bulk_load(ring0 + st * PIECE_BYTES, src, bytes, full0 + st * 8);
            __syncwarp();
            ++g;
        };
        if ((long long)blockIdx.x < n_batches) piece(weights, PIECE_BYTES);
        for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
            for (int l = 1; l < n_conv; ++l)
#pragma unroll 1
                for (int i = 0; i < KS; ++i) piece(weights + PIECE_BYTES + ((size_t)(l - 1) * KS + i) * PIECE_BYTES, PIECE_BYTES);
#pragma unroll 1
            for (int i = 0; i < KS; ++i) piece(head_w + (size_t)i * HEAD_PIECE_BYTES, HEAD_PIECE_BYTES);
            if (batch + gridDim.x < n_batches) piece(weights, PIECE_BYTES);
        }
    } else if (warp == EW) {
        // ===== MMA issuer (converged; one elected lane issues) =====
        uint32_t g = 0, it = 0;  // pieces consumed; batches
        // the MMAs of one tile over `ksteps` ring stages starting at piece g: three filter rows per K chunk
        auto mmas = [&](int t, uint32_t src, uint32_t acc, uint32_t idesc, int ksteps, uint32_t dy_units) {
            const uint64_t a_desc = smem_desc(src, LBO_A, SBO_A);
#pragma unroll 1
            for (int ks = 0; ks < ksteps; ++ks) {
                const uint32_t gg = g + (uint32_t)ks, st = gg % NS;
                if (t == 0) {
                    mbar_wait(full0 + st * 8, (gg / NS) & 1u);  // the K chunk's three filter rows have landed
                    fence_after();
                }
                if (elect_one()) {
                    const uint64_t bd = smem_desc(ring0 + st * PIECE_BYTES, LBO_W, SBO_W);
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint64_t ad = a_desc + (uint64_t)(int64_t)((dy - 1) * PW + t * 128 + ks * (int)(2 * LBO_A >> 4));
                        umma(acc, ad, bd + (uint64_t)(dy * dy_units), idesc, (ks | dy) > 0);
                    }
                    if (t == TILES - 1) umma_commit(empty0 + st * 8);  // both tiles have used the stage
                }
                __syncwarp();
            }
        };
        // one stem / trunk layer of one tile; idx = its number in the tile's sequence
        auto trunk = [&](int l, int t, uint32_t idx, uint32_t batch_no) {
            TRACE(0, l, t, 0);  // issuer: about to wait for the tile's previous epilogue
            if (l == 0) mbar_wait(stem_ready0 + t * 8, batch_no & 1u);  // the stager warp has written this batch's stem input
            // the tile's previous epilogue is through: its accumulator columns are read, this layer's input rows are written
            if (idx > 0) mbar_wait(epi_done0 + t * 8, (idx - 1u) & 1u);
            fence_after();
            TRACE(0, l, t, 1);  // issuer: epilogue seen
            mmas(t, (l == 0 || !(l & 1)) ? aT : aX, tmem_base + (uint32_t)t * TILE_COLS, instr_desc(128, 3 * C, F16), l == 0 ? 1 : KS, DY_BYTES >> 4);
            if (elect_one()) {
                umma_commit(mma_done0 + t * 8);
                // a batch's last trunk layer: when it is complete nothing reads the tile's rows of t any more - the stager warp may
                // write the next batch's stem input there
                if (l == n_conv - 1) umma_commit(stage_go0 + t * 8);
            }
            __syncwarp();
            TRACE(0, l, t, 2);  // issuer: tile-layer issued and committed
        };
        if ((long long)blockIdx.x < n_batches) {
            for (int t = 0; t < TILES; ++t) trunk(0, t, 0u, 0u);
            g += 1;
        }
        for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x, ++it) {
            const bool has_next = batch + gridDim.x < n_batches;
            for (int l = 1; l < n_conv; ++l) {
                for (int t = 0; t < TILES; ++t) trunk(l, t, it * (uint32_t)n_conv + (uint32_t)l, it);
                g += KS;
            }
            for (int t = 0; t < TILES; ++t) {
                // heads of this batch (read x: the last trunk epilogue of the tile; their own accumulator columns: the previous
                // batch's head epilogue), then the next batch's stem - the head pieces are ring stages g .. g + 3, the stem piece g + 4
                mbar_wait(epi_done0 + t * 8, (it * (uint32_t)n_conv + (uint32_t)n_conv - 1u) & 1u);
                if (it > 0) mbar_wait(head_epi0 + t * 8, (it - 1u) & 1u);
                fence_after();
                mmas(t, aX, tmem_base + HEAD_COLS + (uint32_t)t * NH, instr_desc(128, NH, F16), KS, HEAD_DY_BYTES >> 4);
                if (elect_one()) umma_commit(head_mma0 + t * 8);
                __syncwarp();
                if (has_next) {
                    g += KS;
                    trunk(0, t, (it + 1u) * (uint32_t)n_conv, it + 1u);
                    g -= KS;
                }
            }
            g += KS + (has_next ? 1u : 0u);
        }
    } else if (warp < EW) {
        // ===== epilogue warps: thread = one pixel row of a tile, CPW of the 64 output channels =====
        const int half = (int)(warp >> 2);  // which CPW channels
        const int row_in_tile = (int)((warp & 3u) * 32u + lane);
        int pos[TILES], yy[TILES], xx[TILES];
        bool valid[TILES];
        uint32_t row_off[TILES];
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
            const int r = t * 128 + row_in_tile;
            valid[t] = decode_row(r, pos[t], yy[t], xx[t]);
            row_off[t] = (GUARD + r) * ROWB;
        }
        const uint32_t lane_addr = tmem_base + (((warp & 3u) * 32u) << 16);
        const int lm = (int)((lane + 31u) & 31u), lp = (int)((lane + 1u) & 31u);  // left / right neighbour row (see the header)
        // stem / trunk layer l of both tiles; idx = its number in a tile's sequence
        auto trunk = [&](int l, uint32_t idx) {
            uint8_t *dst = (l & 1) ? bufT : bufX;          // stem and conv2 write x, conv1 writes t
            const bool skip = l > 0 && !(l & 1);            // conv2: + x, in place
            float2 bias2[CPW / 2];  // this warp's output channels
#pragma unroll
            for (int i = 0; i < CPW / 4; ++i) {
                const float4 b4 = *reinterpret_cast<const float4 *>(s_bias + l * C + half * CPW + 4 * i);
                bias2[2 * i] = make_float2(b4.x, b4.y);
                bias2[2 * i + 1] = make_float2(b4.z, b4.w);
            }
#pragma unroll
            for (int t = 0; t < TILES; ++t) {
                const uint32_t acc = lane_addr + (uint32_t)t * TILE_COLS + (uint32_t)(half * CPW);
                if (warp == 0 || warp == EW - 1) TRACE(warp == 0 ? 1 : 2, l, t, 0);  // epilogue: about to wait for the MMAs
                mbar_wait(mma_done0 + t * 8, idx & 1u);
                fence_after();
                if (warp == 0 || warp == EW - 1) TRACE(warp == 0 ? 1 : 2, l, t, 1);  // MMAs complete
                uint32_t vm[NCC][16], v0[NCC][16], vp[NCC][16];
#pragma unroll
                for (int cc = 0; cc < NCC; ++cc) {
                    tmem_ld16_issue(acc + 16 * cc, vm[cc]);
                    tmem_ld16_issue(acc + C + 16 * cc, v0[cc]);
                    tmem_ld16_issue(acc + 2 * C + 16 * cc, vp[cc]);
                }
                tmem_ld_wait();
                if (warp == 0 || warp == EW - 1) TRACE(warp == 0 ? 1 : 2, l, t, 2);  // accumulators in registers
#pragma unroll
                for (int cc = 0; cc < NCC; ++cc) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint8_t *p = dst + (uint32_t)(half * (CPW / 8) + cc * 2 + h) * LBO_A + row_off[t];
                        float2 f[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int j = h * 8 + 2 * i;
                            float2 m, q;
                            if (S16) {
                                // the neighbours' partial sums travel as fp16 pairs: one shuffle per direction for two channels;
                                // 11 significand bits against the 8 / 11 the sum is rounded to anyway
                                m = unpack16<true>(__shfl_sync(0xFFFFFFFFu, pack16_sat(__uint_as_float(vm[cc][j]), __uint_as_float(vm[cc][j + 1])), lm));
                                q = unpack16<true>(__shfl_sync(0xFFFFFFFFu, pack16_sat(__uint_as_float(vp[cc][j]), __uint_as_float(vp[cc][j + 1])), lp));
                            } else {
                                m = make_float2(__shfl_sync(0xFFFFFFFFu, __uint_as_float(vm[cc][j]), lm),
                                                __shfl_sync(0xFFFFFFFFu, __uint_as_float(vm[cc][j + 1]), lm));
                                q = make_float2(__shfl_sync(0xFFFFFFFFu, __uint_as_float(vp[cc][j]), lp),
                                                __shfl_sync(0xFFFFFFFFu, __uint_as_float(vp[cc][j + 1]), lp));
                            }
                            const float2 c0 = make_float2(__uint_as_float(v0[cc][j]), __uint_as_float(v0[cc][j + 1]));
                            f[i] = fadd2(fadd2(fadd2(c0, m), q), bias2[cc * 8 + h * 4 + i]);
                        }
                        if (skip) {
                            const uint4 s = *reinterpret_cast<const uint4 *>(p);
                            f[0] = fadd2(f[0], unpack16<F16>(s.x));
                            f[1] = fadd2(f[1], unpack16<F16>(s.y));
                            f[2] = fadd2(f[2], unpack16<F16>(s.z));
                            f[3] = fadd2(f[3], unpack16<F16>(s.w));
                        }
                        uint4 o = make_uint4(0, 0, 0, 0);
                        if (valid[t]) o = make_uint4(pack16_relu<F16>(f[0]), pack16_relu<F16>(f[1]), pack16_relu<F16>(f[2]), pack16_relu<F16>(f[3]));
                        *reinterpret_cast<uint4 *>(p) = o;
                    }
                }
                // this warp's part of the tile is in place for the tensor core and its accumulator reads are complete
                if (warp == 0 || warp == EW - 1) TRACE(warp == 0 ? 1 : 2, l, t, 3);  // outputs computed and stored
                fence_before();
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(epi_done0 + t * 8);
                if (warp == 0 || warp == EW - 1) TRACE(warp == 0 ? 1 : 2, l, t, 4);  // arrived
            }
        };
        uint32_t it = 0;
        if ((long long)blockIdx.x < n_batches) trunk(0, 0u);
        for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x, ++it) {
            for (int l = 1; l < n_conv; ++l) trunk(l, it * (uint32_t)n_conv + (uint32_t)l);
            // the next batch's stem first (its MMAs were issued right behind this batch's heads): conv1 of the next batch can start
            if (batch + gridDim.x < n_batches) trunk(0, (it + 1u) * (uint32_t)n_conv);
            // ---- heads: [POS][35][42] fp32 in the Flatten() order of NCHW.  Accumulator columns of a tile: 0..31 policy channels
            // (1x1 conv: centre tap only, no neighbour terms), 32 + 8 dx + v = value channel v, filter column dx (models.py:
            // pack_head_weights(wide=True)).  Warps 0..3: policy 0..15 and the 3 value channels; warps 4..7: policy 16..31.
            const float *hb = s_bias + n_conv * C;
            if (warp == 0 || warp == EW - 1) TRACE(warp == 0 ? 1 : 2, 50, 0, 0);
            if (it > 0) mbar_wait(hact_free, (it - 1u) & 1u);  // the FC warp is through with the previous batch's head activations
            if (warp == 0 || warp == EW - 1) TRACE(warp == 0 ? 1 : 2, 50, 0, 1);
#pragma unroll
            for (int t = 0; t < TILES; ++t) {
                const uint32_t acc = lane_addr + HEAD_COLS + (uint32_t)t * NH;
                float hbr[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 b4 = *reinterpret_cast<const float4 *>(hb + (half & 1) * 16 + 4 * i);
                    hbr[4 * i] = b4.x; hbr[4 * i + 1] = b4.y; hbr[4 * i + 2] = b4.z; hbr[4 * i + 3] = b4.w;
                }
                const float4 hbv = *reinterpret_cast<const float4 *>(hb + 32);
                mbar_wait(head_mma0 + t * 8, it & 1u);
                fence_after();
                if (warp == 0 || warp == EW - 1) TRACE(warp == 0 ? 1 : 2, 50, t, 2);
                uint32_t v[16], wa[16], wb[16];
                if (half < 2) tmem_ld16_issue(acc + (uint32_t)half * 16u, v);
                if (half == 0) {
                    tmem_ld16_issue(acc + 32, wa);  // value channels: filter columns -1 (0..7) and 0 (8..15)
                    tmem_ld16_issue(acc + 48, wb);  // filter column +1 (0..7)
                }
                tmem_ld_wait();
                float val[3] = {0.f, 0.f, 0.f};
                if (half == 0) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float m = __shfl_sync(0xFFFFFFFFu, __uint_as_float(wa[c]), lm);
                        const float q = __shfl_sync(0xFFFFFFFFu, __uint_as_float(wb[c]), lp);
                        val[c] = ((__uint_as_float(wa[8 + c]) + m) + q) + (c == 0 ? hbv.x : (c == 1 ? hbv.y : hbv.z));
                    }
                }
                if (valid[t] && half < 2) {
                    float *o = hact + pos[t] * NHU * 42 + yy[t] * c4::W + xx[t];
#pragma unroll
                    for (int c = 0; c < 16; ++c) o[(half * 16 + c) * 42] = fmaxf(__uint_as_float(v[c]) + hbr[c], 0.f);
                    if (half == 0) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) o[(32 + c) * 42] = fmaxf(val[c], 0.f);
                    }
                }
                fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(head_epi0 + t * 8);
                if (warp == 0 || warp == EW - 1) TRACE(warp == 0 ? 1 : 2, 50, t, 3);
            }
            // both tiles' head activations are in place for the FC warp
            if (lane == 0) mbar_arrive(hact_ready);
        }
    } else if (warp == EW + 2) {
        // ===== stager warp: the stem input of every batch (lane = 4 pixel rows of a tile).  The first batch's goes in at once, the
        // next batch's when the tensor core reports the tile's last trunk layer complete (stage_go) =====
        uint32_t it = 0;
        for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x, ++it) {
            RowRec rec[TILES][4];
#pragma unroll
            for (int t = 0; t < TILES; ++t)
#pragma unroll
                for (int j = 0; j < 4; ++j) rec[t][j] = load_rec(batch, t * 128 + j * 32 + (int)lane);
#pragma unroll
            for (int t = 0; t < TILES; ++t) {
                if (it > 0) mbar_wait(stage_go0 + t * 8, (it - 1u) & 1u);
#pragma unroll
                for (int j = 0; j < 4; ++j) stage_row(rec[t][j], t * 128 + j * 32 + (int)lane);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(stem_ready0 + t * 8);
            }
        }
    } else if (warp == EW + 3) {
        // ===== FC warp: both fully connected layers on CUDA cores, off the epilogue warps' critical path.  Lane L takes inputs
        // k = L, L + 32, ...: every FC weight is read once per CTA (coalesced) and used for POS positions; acc[p][j], j = 7 = value head
        uint32_t it = 0;
        for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x, ++it) {
            mbar_wait(hact_ready, it & 1u);
            float acc[POS * 8];
#pragma unroll
            for (int i = 0; i < POS * 8; ++i) acc[i] = 0.f;
#pragma unroll 6
            for (int k = (int)lane; k < 32 * 42; k += 32) {
                float wj[7];
#pragma unroll
                for (int j = 0; j < 7; ++j) wj[j] = __ldg(fc_policy_w + j * (32 * 42) + k);
#pragma unroll
                for (int p = 0; p < POS; ++p) {
                    const float xv = hact[p * NHU * 42 + k];
#pragma unroll
                    for (int j = 0; j < 7; ++j) acc[p * 8 + j] = fmaf(wj[j], xv, acc[p * 8 + j]);
                }
            }
#pragma unroll 1
            for (int k = (int)lane; k < 3 * 42; k += 32) {
                const float wv = __ldg(fc_value_w + k);
#pragma unroll
                for (int p = 0; p < POS; ++p) acc[p * 8 + 7] = fmaf(wv, hact[p * NHU * 42 + 32 * 42 + k], acc[p * 8 + 7]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(hact_free);  // the head activations are consumed
            // warp reduction by recursive halving: each step exchanges half of the values; lane L ends with the total of index L
            constexpr int V = POS * 8;
            static_assert(V == 32, "one value per lane after the halving");
#pragma unroll
            for (int h = V / 2; h >= 1; h >>= 1) {
                const bool up = (lane & (uint32_t)h) != 0;
#pragma unroll
                for (int i = 0; i < h; ++i) {
                    const float send = up ? acc[i] : acc[i + h];
                    const float keep = up ? acc[i + h] : acc[i];
                    acc[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, h);
                }
            }
            const float sum = acc[0];
            const int p = (int)lane >> 3, j = (int)lane & 7;
            const long long gp = batch * POS + p;
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
    fence_before();
    __syncthreads();
#ifdef WIDE_TRACE
    if (blockIdx.x == 0) {
        if (lane == 0 && (warp == EW || warp == 0 || warp == EW - 1)) {
            const int slot = warp == EW ? 0 : (warp == 0 ? 1 : 2);
            for (int i = tr_n; i < TRACE_CAP; ++i) s_trace[(slot * TRACE_CAP + i) * 2] = 0xFFFFFFFFu;
        }
        __syncthreads();
        for (uint32_t i = tid; i < TRACE_WARPS * TRACE_CAP * 2; i += THREADS) g_trace[i] = s_trace[i];
    }
#endif
    if (warp == 0) tmem_dealloc(tmem_base, 512);
    if (timing && tid == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        atomicMax(timing + t_slot + 1, now);
        __threadfence();
        if (atomicAdd(timing + 1, 1ull) == (unsigned long long)gridDim.x - 1ull) {  // the launch's last CTA: next launch, next slot
            timing[1] = 0;
            __threadfence();
            atomicAdd(timing, 1ull);
        }
    }
}

}  // namespace

static unsigned long long *g_timing = nullptr;
static int g_timing_cap = 0;

extern "C" {

/* Measurement hook: `buf` = device array of 4 + 2 * cap uint64, [0] = launches so far, [1] = scratch, then per launch (slot =
 * launch % cap) {first CTA start, last CTA end} on %globaltimer (ns); the caller initialises the pairs to {~0, 0}.  The pointer is
 * baked into launches (and captured graphs) made after this call; NULL switches the hook off. */
int32_t az_resnet_wide_set_timing(void *buf, int32_t cap) {
    g_timing = (unsigned long long *)buf;
    g_timing_cap = buf ? cap : 0;
    return AZ_OK;
}

#ifdef WIDE_TRACE
/* timing study: copy out the timeline of CTA 0 of the last launch: [3 warps][TRACE_CAP][code, clock] */
int32_t az_resnet_wide_trace(unsigned int *out, int32_t cap) {
    cudaDeviceSynchronize();
    const int32_t n = TRACE_WARPS * TRACE_CAP * 2;
    if (cap < n) return -1;
    cudaMemcpyFromSymbol(out, g_trace, (size_t)n * 4);
    return n;
}
#endif

/* internal: called by az_resnet_forward_leaves_v2 (csrc/az_conv.cu), variant 4.  Takes the packed weights of the layer-pipelined
 * kernel (az_resnet_pipe_weight_bytes(num_blocks, 64) bytes, models.py:pack_trunk_weights_pipe). */
int32_t az_resnet_wide_launch(az_engine *engine, const az_resnet_desc *d, float *logits, float *values, void *stream) {
    if (!engine || !d || !d->trunk_w || !d->trunk_b || d->num_channels != 64) return AZ_E_INVALID;
    if (d->operand_format != AZ_FMT_BF16 && d->operand_format != AZ_FMT_F16) return AZ_E_INVALID;
    if (d->num_blocks < 0 || 1 + 2 * d->num_blocks > MAX_CONV) return AZ_E_INVALID;
    const uint64_t *bb0 = nullptr, *bb1 = nullptr;
    const uint8_t *status = nullptr, *player = nullptr;
    const int32_t *elist = nullptr, *ecount = nullptr;
    int32_t n = 0;
    if (az_leaf_arrays(engine, &bb0, &bb1, &status, &n) != AZ_OK || az_leaf_players(engine, &player) != AZ_OK ||
        az_leaf_compact(engine, &elist, &ecount) != AZ_OK || n <= 0)
        return AZ_E_INVALID;
    static bool attr_set[64] = {false};
    const int dev = az_device(engine);
    if (dev < 0 || dev >= 64 || cudaSetDevice(dev) != cudaSuccess) return AZ_E_CUDA;
    if (!attr_set[dev]) {
        if (cudaFuncSetAttribute(k_resnet_wide<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LAUNCH) != cudaSuccess) return AZ_E_CUDA;
        if (cudaFuncSetAttribute(k_resnet_wide<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LAUNCH) != cudaSuccess) return AZ_E_CUDA;
        if (cudaFuncSetAttribute(k_resnet_wide<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LAUNCH) != cudaSuccess) return AZ_E_CUDA;
        if (cudaFuncSetAttribute(k_resnet_wide<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LAUNCH) != cudaSuccess) return AZ_E_CUDA;
        attr_set[dev] = true;
    }
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return AZ_E_CUDA;
    const int batches = (n + POS - 1) / POS;
    // neighbour shuffles as fp16 pairs (default; AZ_WIDE_SHFL16=0 = fp32 shuffles): 2-3 % of the kernel, and no measurable change of the
    // outputs - max deviation from the fp32 `predict` over 16384 positions 3.8e-5 / 1.6e-4 (priors / values) with fp16 operands either
    // way, 3.0e-4 / 1.19e-3 against 3.0e-4 / 1.13e-3 with bf16 operands (profiles/r02_evaluator_deviation.json)
    static int s16_env = -2;
    if (s16_env == -2) {
        const char *e = getenv("AZ_WIDE_SHFL16");
        s16_env = e ? atoi(e) : 1;
    }
    const bool f16 = d->operand_format == AZ_FMT_F16;
    const bool s16 = s16_env != 0;
    auto kern = f16 ? (s16 ? k_resnet_wide<true, true> : k_resnet_wide<true, false>) : (s16 ? k_resnet_wide<false, true> : k_resnet_wide<false, false>);
    kern<<<batches < sms ? batches : sms, THREADS, SMEM_LAUNCH, (cudaStream_t)stream>>>(
        bb0, bb1, player, status, elist, ecount, (long long)n, (const uint8_t *)d->trunk_w, d->trunk_b, d->num_blocks, (const uint8_t *)d->head_conv_w,
        d->head_conv_b, d->fc_policy_w, d->fc_policy_b, d->fc_value_w, d->fc_value_b, logits, values, g_timing, g_timing_cap);
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_E_CUDA;
}

}  // extern "C"
