// ResNet(board, 7, num_res_blocks, C) (src/alphazero_simple/resnet.py:30-103; the reference's own instance is 9 x 128,
// src/alphazero_less_simple/main.py:13) behind the main package's Model API, BatchNorm folded, as ONE tcgen05 kernel for the
// leaves of the search: stem conv3x3 (3 -> C), residual blocks (2 x conv3x3 C -> C, skip, ReLU), the policy / value head
// convolutions and the two small FC layers; activations never leave the SM.  Two instances: C = 128 with 2 accumulator tiles
// (4 positions) per CTA - 223.9 MFLOP per position at 9 blocks - and C = 64 with 4 tiles (8 positions), the headline net 4 x 64.
// The comments below give the numbers of the 128-channel instance.
//
// Same implicit-GEMM formulation as csrc/az_conv.cu (pixel = GEMM row, channels = K, activation buffer K-group-major so that a
// filter tap is the same buffer with the MMA descriptor moved by 8*dy + dx rows; compact padding: a position is 7 x 8 pixel rows),
// but a different schedule, because with 128 channels nothing that kernel relies on fits: the 9 taps of a layer are 288 KB.
//  * A CTA works on 4 positions = 2 accumulator tiles of 128 rows x 128 fp32 columns.  Both tiles use every weight piece, so a
//    layer's weights are streamed ONCE per CTA, K-chunk-major: a piece is the 9 taps of one K chunk, [9][128 out][16 in] = 36 KB,
//    through a 2-stage ring of bulk async copies - 36 KB per 18 MMAs of 64 cycles = 32 B/clk, half of the SM's L2 port.  The
//    pieces must be this large: a bulk copy costs ~300 cycles whatever its size (measured: 4 KB pieces arrived one per
//    ~350 cycles = 11.7 B/clk and the tensor pipe sat at 34 %; profiles/r02_k_resnet128_4k_pieces_ncu_raw.json).
//  * Layers are pipelined through tensor memory instead of ping-ponging two groups: the accumulators are double-buffered
//    (2 sets x 2 tiles x 128 columns = all 512 columns).  While the eight epilogue warps drain layer l (tcgen05.ld, bias (+ skip),
//    ReLU, round to 16 bits, write the next layer's A operand), the tensor core already runs layer l + 1 into the other set:
//    its MMAs are ordered by K chunk, and chunk ks needs only input channels 16 ks .. 16 ks + 15, so it waits for an mbarrier
//    that the epilogue warps arrive on after writing exactly those channels (`chunk[ks]`).  The epilogue of a layer takes
//    ~1.5 k cycles against 9.2 k cycles of MMAs (144 x 64), so after the first chunk the tensor core never waits; the only
//    bubble is the hand-over at the start of a layer (last MMA done -> first chunk written).
//  * The residual sum is formed in place: conv2's epilogue reads the skip value of its row and overwrites it, which no MMA reads
//    any more (conv1 of the same block is complete) - two activation buffers (x, t) of 72 KB are enough.
//  * Shared-memory operand traffic: A 4 KB + B 4 KB per 64-cycle MMA = 128 B/clk, the crossbar's rate; N = 128 is therefore
//    balanced where the 64-channel kernel (6 KB per 32 cycles) is operand-bound.
//    With 64 channels the same MMAs are operand-bound as in csrc/az_conv.cu (6 KB per 32 cycles -> ~48 cycles each); what this
//    schedule saves there is everything around them: no per-group hand-overs, weights once per layer for all four tiles.
// Roles: warps 0..7 epilogue (thread = one pixel row of each pair of tiles), warp 8 issues MMAs (one elected lane), warp 9 streams weights.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/az_engine.h"
#include "c4_bitboard.cuh"
#include "tcgen05.cuh"

namespace {

using namespace tc05;

constexpr int GUARD = 16;              // zero rows before / after the tiles (a tap moves the window by up to 9 rows)
constexpr uint32_t ROWB = 16;
constexpr int PW = 8, PIX = 56, LEAD = 8;  // pixel-row stride, rows per position, leading zero rows of a tile
constexpr int TPOS = 2;                // positions per 128-row tile
constexpr uint32_t SBO_A = 128;        // next 8-row group
constexpr uint32_t LBO_W = 128, SBO_W = 256;     // canonical K-major [N][16]
constexpr int NHC = 48, NHU = 35;      // head conv channels: 32 policy (1x1, centre tap) + 3 value (3x3) + padding
constexpr uint32_t HEAD_TAP_BYTES = NHC * 16 * 2;    // 1536
constexpr uint32_t HEAD_PIECE_BYTES = 9 * HEAD_TAP_BYTES;
static_assert(LEAD + TPOS * PIX <= 128, "a tile's positions must fit its 128 rows");

// C channels, TILES accumulator tiles per CTA (TILES * C tensor-memory columns per accumulator set), NS ring stages, EW epilogue
// warps (warp EW issues the MMAs, warp EW + 1 streams the weights), CTAS resident CTAs per SM
template <int C_, int TILES_, int NS_, int MAX_CONV_, int EW_ = 8, int CTAS_ = 1, bool PAIR_ = false>
struct Cfg {
    static constexpr int C = C_, TILES = TILES_, NS = NS_, MAX_CONV = MAX_CONV_, EW = EW_, CTAS = CTAS_;
    static constexpr bool PAIR = PAIR_;           // clusters of two CTAs on two SMs share every MMA (cta_group::2, M = 256)
    static constexpr int THREADS = (EW + 2) * 32, WTHREADS = (EW + 1) * 32, ETHREADS = EW * 32;
    static constexpr int TSTEP = EW / 4;          // an epilogue thread owns one row of the tiles w / 4, w / 4 + TSTEP, ...
    static constexpr int SETCOLS = TILES * C;     // tensor-memory columns of one accumulator set
    static constexpr int KG = C / 8;              // K groups of 8 channels (16 bytes per row)
    static constexpr int KS = C / 16;             // K steps per tap = pieces per layer = chunks of an epilogue
    static constexpr int POS = TPOS * TILES;      // positions per CTA
    static constexpr int ROWS = TILES * 128;
    static constexpr int RTOT = ROWS + 2 * GUARD; // rows per K group
    static constexpr int RPT = TILES / TSTEP;     // rows per epilogue thread
    static constexpr uint32_t LBO_A = RTOT * ROWB;          // next K group
    static constexpr uint32_t BUF_BYTES = KG * LBO_A;
    static constexpr uint32_t TAP_BYTES = C * 16 * 2;       // one tap of a K chunk, [C out][16 in]
    static constexpr uint32_t PIECE_BYTES = 9 * TAP_BYTES;  // the 9 taps of a K chunk
    static constexpr uint32_t OFF_RING = 2 * BUF_BYTES;
    static constexpr uint32_t OFF_BIAS = OFF_RING + NS * PIECE_BYTES;
    static constexpr uint32_t OFF_BARS = OFF_BIAS + (MAX_CONV * C + NHC) * 4;
    static constexpr int NBARS = 2 * NS + 1 + KS + NS + 1;  // full[NS] empty[NS] mma_done chunk[KS] | pair: pfull[NS] batch_ready
    static constexpr uint32_t SMEM_BYTES = OFF_BARS + NBARS * 8 + 16;
    // FC-tail scratch inside t's K groups 2..7 (K groups 0 / 1 take the next batch's stem input): head activations, per-warp sums
    static constexpr uint32_t OFF_HACT = 2 * LBO_A;
    static constexpr uint32_t HACT_BYTES = POS * NHU * 42 * 4;
    static constexpr uint32_t OFF_RED = OFF_HACT + ((HACT_BYTES + 1023) / 1024) * 1024;
    static_assert(SETCOLS == 128 || SETCOLS == 256, "two accumulator sets: 256 or 512 tensor-memory columns");
    static_assert((EW == 4 || EW == 8) && TILES % TSTEP == 0, "epilogue warps come in groups of four (one per 32 accumulator lanes)");
    static_assert(CTAS * (SMEM_BYTES + 1024) <= 233472 && SMEM_BYTES <= 232448, "shared memory budget");
    static_assert(HEAD_PIECE_BYTES <= PIECE_BYTES, "head pieces use the same ring");
    static_assert(OFF_RED + EW * POS * 8 * 4 <= 8 * LBO_A, "FC scratch must stay inside K groups 2..7");
};
using Cfg128 = Cfg<128, 2, 2, 19>;       // 9 blocks; 36 KB pieces
using Cfg64 = Cfg<64, 4, 4, 23>;         // 11 blocks; 18 KB pieces; 8 positions per CTA
// 64 channels, TWO CTAs per SM of 4 positions each (4 epilogue warps, 6 warps in all): while one CTA is in an epilogue, a layer
// hand-over, its FC tail or its prologue, the other one's MMAs keep the tensor core busy - the overlap the ping-pong kernel builds
// by hand inside one CTA, here between two independent instruction streams.  5 blocks at most (bias table).
using Cfg64x2 = Cfg<64, 2, 2, 11, 4, 2>;
// The same with CTA pairs: cta_group::2 MMAs (M = 256: this CTA's tile and the peer's), each CTA stages HALF of every weight piece
// and the tensor cores fetch B once per pair: 5 KB instead of 6 KB of shared-memory operands per MMA per SM - measured 43 against 48
// cycles per 128x64x16 MMA (scripts/ubench/mma_pair.cu) - and half the weight traffic into each SM.
using Cfg64x2p = Cfg<64, 2, 2, 11, 4, 2, true>;
using Cfg128p = Cfg<128, 2, 2, 19, 8, 1, true>;  // 128 channels with CTA pairs: B 2 KB instead of 4 KB per MMA per SM, 18 KB half-pieces

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

template <typename K, bool F16>
__global__ void __launch_bounds__(K::THREADS, K::CTAS)
k_resnet_pipe(const uint64_t *__restrict__ leaf_bb0, const uint64_t *__restrict__ leaf_bb1, const uint8_t *__restrict__ leaf_player,
            const uint8_t *__restrict__ leaf_status, const int32_t *__restrict__ eval_list, const int32_t *__restrict__ eval_count,
            long long n_slots, const uint8_t *__restrict__ weights, const float *__restrict__ biases, int num_blocks,
            const uint8_t *__restrict__ head_w, const float *__restrict__ head_b, const float *__restrict__ fc_policy_w,
            const float *__restrict__ fc_policy_b, const float *__restrict__ fc_value_w, const float *__restrict__ fc_value_b,
            float *__restrict__ logits, float *__restrict__ values, uint32_t stagger_ns) {
    constexpr int C = K::C, TILES = K::TILES, NS = K::NS, KS = K::KS, POS = K::POS, ROWS = K::ROWS, RPT = K::RPT, NBARS = K::NBARS;
    constexpr int EW = K::EW, THREADS = K::THREADS, WTHREADS = K::WTHREADS, ETHREADS = K::ETHREADS, TSTEP = K::TSTEP;
    constexpr uint32_t SETCOLS = K::SETCOLS;
    constexpr bool PAIR = K::PAIR;
    constexpr uint32_t LBO_A = K::LBO_A, BUF_BYTES = K::BUF_BYTES, TAP_BYTES = K::TAP_BYTES, PIECE_BYTES = K::PIECE_BYTES;
    constexpr uint32_t OFF_RING = K::OFF_RING, OFF_BIAS = K::OFF_BIAS, OFF_BARS = K::OFF_BARS, OFF_HACT = K::OFF_HACT, OFF_RED = K::OFF_RED;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *bufX = smem, *bufT = smem + BUF_BYTES;
    float *s_bias = reinterpret_cast<float *>(smem + OFF_BIAS);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BARS);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BARS + NBARS * 8);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    // position j of a batch is slot eval_list[j], j < *eval_count: only the leaves that wait for an evaluation are processed
    const long long n = eval_list ? (long long)__ldg(eval_count) : n_slots;
    const int n_conv = 1 + 2 * num_blocks, n_layers = n_conv + 1;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NS), mma_done = smem_u32(bars + 2 * NS), chunk0 = smem_u32(bars + 2 * NS + 1);
    const uint32_t pfull0 = chunk0 + KS * 8, batch_ready = pfull0 + NS * 8;
    const uint32_t ring0 = smem_u32(smem + OFF_RING);
    // PAIR: the two CTAs of a cluster (two SMs) run every MMA together; rank 0 issues them.  Barriers that gate the MMAs live in
    // the leader: chunk[] and batch_ready collect arrivals from both CTAs, pfull[] is the peer's "my half of the weights has landed".
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    if (warp == 0) {
        if (PAIR) tmem_alloc2(smem_u32(tmem_slot), 2 * SETCOLS);
        else tmem_alloc(smem_u32(tmem_slot), 2 * SETCOLS);
    }
    if (tid == 32) {
        for (int i = 0; i < 2 * NS + 1; ++i) mbar_init(smem_u32(bars + i), 1u);
        for (int i = 0; i < KS; ++i) mbar_init(chunk0 + i * 8, (uint32_t)(PAIR ? 2 * EW : EW));  // one arrival per epilogue warp (of both CTAs)
        for (int i = 0; i < NS; ++i) mbar_init(pfull0 + i * 8, 1u);
        mbar_init(batch_ready, 2u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const uint32_t chunk_lead = PAIR ? mapa(chunk0, 0u) : chunk0;  // shared::cluster addresses of the leader's barriers
    const uint32_t pfull_lead = PAIR ? mapa(pfull0, 0u) : pfull0, ready_lead = PAIR ? mapa(batch_ready, 0u) : batch_ready;
    for (uint32_t i = tid; i < 2 * BUF_BYTES / 16; i += THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (uint32_t i = tid; i < (uint32_t)n_conv * C; i += THREADS) s_bias[i] = __ldg(biases + i);
    for (uint32_t i = tid; i < NHC; i += THREADS) s_bias[n_conv * C + i] = __ldg(head_b + i);
    const uint32_t aX = smem_u32(bufX) + GUARD * ROWB, aT = smem_u32(bufT) + GUARD * ROWB;
    fence_before();
    __syncthreads();  // barriers initialised, tensor memory allocated, buffers zeroed, biases staged
    if (PAIR) cluster_sync_all();  // ... in the peer as well, before anything remote happens
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    auto batch_sync = [&]() { asm volatile("bar.sync 2, %0;" ::"n"(WTHREADS) : "memory"); };  // everyone but the weight producer

    const long long n_batches = (n + POS - 1) / POS;
    // PAIR: cluster c walks over pairs of batches (2 pb, 2 pb + 1), one per CTA; both CTAs run the same number of iterations (a
    // batch beyond the last one is empty: zero rows, no outputs)
    const long long first = PAIR ? (long long)cluster_id_x() : (long long)blockIdx.x, stride = PAIR ? (long long)cluster_count_x() : (long long)gridDim.x;
    const long long n_units = PAIR ? (n_batches + 1) / 2 : n_batches;
    uint32_t it = 0;
    uint32_t g = 0;  // pieces produced (warp 9) / consumed (warp 8) so far: stage = g % NS, use = g / NS
    for (long long unit = first; unit < n_units; unit += stride, ++it) {
        const long long batch = PAIR ? 2 * unit + rank : unit;
        const long long pos0 = batch * POS;
        const uint32_t gl0 = it * (uint32_t)n_layers, ge0 = it * (uint32_t)n_conv;
        // this thread's stem rows: the leaf records are requested now and consumed after the batch barrier
        constexpr int SROWS = (ROWS + WTHREADS - 1) / WTHREADS;
        uint64_t in_b0[SROWS], in_b1[SROWS];
        uint32_t in_meta[SROWS];  // bit 0: row of a position that exists, bit 1: side to move, bit 2: row is a board cell, bits 8..: bit index
#pragma unroll
        for (int k = 0; k < SROWS; ++k) {
            const int r = (int)tid + k * WTHREADS;
            int pos, y, x;
            const bool cell = tid < WTHREADS && r < ROWS && decode_row(r, pos, y, x);
            const long long gp = pos0 + pos;
            const bool in = cell && gp < n;
            const long long slot = in ? (eval_list ? (long long)__ldg(eval_list + gp) : gp) : 0;
            const bool live = in && leaf_status[slot] == AZ_LEAF_EVAL;
            in_b0[k] = leaf_bb0[slot];
            in_b1[k] = leaf_bb1[slot];
            in_meta[k] = (live ? 1u : 0u) | ((uint32_t)(leaf_player[slot] & 1) << 1) | (cell ? 4u : 0u) | ((uint32_t)(x * c4::STRIDE + y) << 8);
        }
        if (K::CTAS == 2 && stagger_ns && it == 0 && warp == EW && blockIdx.x >= gridDim.x / 2) {
            // Experiment kept as a switch (default off): two co-resident CTAs run identical batches, and a start offset for the
            // second wave of CTAs (blockIdx >= gridDim / 2 fills the second slot of each SM) was meant to keep them out of lock
            // step.  They drift apart on their own (contention for the tensor core); any offset only costs its own length.
            for (int d = 0; d < n_layers; ++d) __nanosleep(stagger_ns);
        }
        if (warp != EW + 1) {
            batch_sync();  // the previous batch is finished: buffers and tensor memory are this batch's
            // the FC tail of the previous batch used K groups 2..7 of t as scratch, guard rows included: those must be zero again
            if (it > 0)
                for (uint32_t i = tid; i < 6 * 2 * GUARD; i += WTHREADS) {
                    const uint32_t kg = 2 + i / (2 * GUARD), j = i % (2 * GUARD);
                    const uint32_t row = j < GUARD ? j : (uint32_t)(ROWS + j);
                    *reinterpret_cast<uint4 *>(bufT + kg * LBO_A + row * ROWB) = make_uint4(0, 0, 0, 0);
                }
            // stem input in t, K group 0: channels 0..2 = empty / side to move / opponent (cnn.py:93-95).  K group 1 keeps stale
            // finite activations, which the stem's zero weights for channels 8..15 cancel.
#pragma unroll
            for (int k = 0; k < SROWS; ++k) {
                if (!(in_meta[k] & 4u)) continue;
                const int r = (int)tid + k * WTHREADS;
                const int pl = (in_meta[k] >> 1) & 1, bit = (int)(in_meta[k] >> 8);
                const uint32_t live = in_meta[k] & 1u;
                const uint32_t s0 = (uint32_t)((in_b0[k] >> bit) & 1ull), s1 = (uint32_t)((in_b1[k] >> bit) & 1ull);
                const uint32_t mine = live * (pl ? s1 : s0), theirs = live * (pl ? s0 : s1), emp = live * (1u - (s0 | s1));
                const uint32_t one = F16 ? 0x3C00u : 0x3F80u;
                *reinterpret_cast<uint4 *>(bufT + (GUARD + r) * ROWB) = make_uint4(emp * one | (mine * one) << 16, theirs * one, 0u, 0u);
            }
            fence_async_smem();
            fence_before();
            batch_sync();
            fence_after();
            if (PAIR && tid == 0) mbar_arrive_cluster(ready_lead);  // this CTA's stem input is in place (the leader's issuer waits for both)
        }

        if (warp == EW + 1) {
            // ===== weight producer: every piece of every layer, in the order the issuer consumes them =====
            for (int l = 0; l < n_layers; ++l) {
                const bool head = l >= n_conv;
                const uint8_t *src = l == 0 ? weights : (head ? head_w : weights + PIECE_BYTES + (size_t)(l - 1) * KS * PIECE_BYTES);
                const uint32_t bytes = head ? HEAD_PIECE_BYTES : PIECE_BYTES;
                // a pair splits B by output channel: this CTA stages rows [rank * N / 2, (rank + 1) * N / 2) of every tap - packed as the
                // first / second half of the piece (models.py:pack_trunk_weights_pipe(pair=True))
                const uint32_t mine = PAIR ? bytes / 2 : bytes;
                const int pieces = l == 0 ? 1 : KS;
#pragma unroll 1
                for (int i = 0; i < pieces; ++i, ++g) {
                    const uint32_t st = g % NS;
                    if (g >= NS) {  // the MMAs of the previous use have read the stage
                        if (PAIR) mbar_wait_cluster(empty0 + st * 8, ((g / NS) - 1u) & 1u);
                        else mbar_wait(empty0 + st * 8, ((g / NS) - 1u) & 1u);
                    }
                    if (elect_one()) bulk_load(ring0 + st * PIECE_BYTES, src + (size_t)i * bytes + (size_t)rank * mine, mine, full0 + st * 8);
                    __syncwarp();
                }
            }
        } else if (warp == EW && !leader) {
            // ===== peer CTA of a pair: tell the leader when this CTA's half of a weight piece has landed =====
            for (int l = 0; l < n_layers; ++l) {
                const int pieces = l == 0 ? 1 : KS;
#pragma unroll 1
                for (int i = 0; i < pieces; ++i, ++g) {
                    const uint32_t st = g % NS;
                    mbar_wait(full0 + st * 8, (g / NS) & 1u);
                    if (elect_one()) mbar_arrive_cluster(pfull_lead + st * 8);
                    __syncwarp();
                }
            }
        } else if (warp == EW) {
            // ===== MMA issuer (converged; one elected lane issues) =====
            if (PAIR) mbar_wait_cluster(batch_ready, it & 1u);  // the peer's stem input is in place as well
            for (int l = 0; l < n_layers; ++l) {
                const bool head = l >= n_conv;
                const uint32_t src = l == 0 ? aT : ((head || (l & 1)) ? aX : aT);  // conv1 (odd l) and the heads read x; conv2 reads t
                const uint32_t idesc = head ? instr_desc(PAIR ? 256 : 128, NHC, F16) : instr_desc(PAIR ? 256 : 128, C, F16);
                const uint32_t acc = tmem_base + (uint32_t)(l & 1) * SETCOLS;
                const uint64_t a_desc = smem_desc(src, LBO_A, SBO_A);
                const int ksteps = l == 0 ? 1 : KS;
                // bytes of one tap inside a staged piece: a pair holds half of the output channels per CTA
                const uint32_t tap_units = ((head ? HEAD_TAP_BYTES : TAP_BYTES) >> 4) / (PAIR ? 2u : 1u);
#pragma unroll 1
                for (int ks = 0; ks < ksteps; ++ks, ++g) {
                    // input channels 16 ks .. 16 ks + 15 of every row are written once all epilogue warps (of both CTAs) have passed them
                    if (l > 0) {
                        if (PAIR) mbar_wait_cluster(chunk0 + ks * 8, (ge0 + (uint32_t)l - 1u) & 1u);
                        else mbar_wait(chunk0 + ks * 8, (ge0 + (uint32_t)l - 1u) & 1u);
                    }
                    const uint32_t st = g % NS;
                    mbar_wait(full0 + st * 8, (g / NS) & 1u);  // the K chunk's 9 taps have landed
                    if (PAIR) mbar_wait_cluster(pfull0 + st * 8, (g / NS) & 1u);  // ... in the peer too
                    fence_after();
                    if (elect_one()) {
                        const uint64_t bd = smem_desc(ring0 + st * PIECE_BYTES, LBO_W, SBO_W);
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const int shift = (tap / 3 - 1) * PW + (tap % 3 - 1);  // rows; one row = 16 B = one descriptor address unit
#pragma unroll
                            for (int t = 0; t < TILES; ++t) {
                                const uint64_t ad = a_desc + (uint64_t)(int64_t)(shift + t * 128 + ks * (int)(2 * LBO_A >> 4));
                                if (PAIR) umma2(acc + t * C, ad, bd + (uint64_t)(tap * tap_units), idesc, (ks | tap) > 0);
                                else umma(acc + t * C, ad, bd + (uint64_t)(tap * tap_units), idesc, (ks | tap) > 0);
                            }
                        }
                        if (PAIR) umma_commit2(empty0 + st * 8);
                        else umma_commit(empty0 + st * 8);
                    }
                    __syncwarp();
                }
                if (elect_one()) {
                    if (PAIR) umma_commit2(mma_done);
                    else umma_commit(mma_done);
                }
                __syncwarp();
            }
        } else {
            // ===== epilogue warps: thread = one row of the tiles w/4, w/4 + TSTEP, ... (RPT rows) =====
            const int tile0 = (int)(warp >> 2);
            const int row_in_tile = (int)((warp & 3u) * 32u + lane);
            int pos[RPT], yy[RPT], xx[RPT];
            bool valid[RPT];
            uint32_t row_off[RPT];
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const int r = (tile0 + TSTEP * j) * 128 + row_in_tile;
                valid[j] = decode_row(r, pos[j], yy[j], xx[j]);
                row_off[j] = (GUARD + r) * ROWB;
            }
            const uint32_t lane_addr = tmem_base + (((warp & 3u) * 32u) << 16);
            constexpr int NQ = KS * RPT;  // (chunk, row) pairs of a layer, chunk-major
            static_assert(NQ % 2 == 0, "the double-buffered loop handles pairs");
            for (int l = 0; l < n_conv; ++l) {
                uint8_t *dst = (l & 1) ? bufT : bufX;          // stem and conv2 write x, conv1 writes t
                const bool skip = l > 0 && !(l & 1);            // conv2: + x, in place
                const float *bias = s_bias + l * C;
                const uint32_t acc = lane_addr + (uint32_t)(l & 1) * SETCOLS;
                if (PAIR) mbar_wait_cluster(mma_done, (gl0 + (uint32_t)l) & 1u);
                else mbar_wait(mma_done, (gl0 + (uint32_t)l) & 1u);
                fence_after();
                uint32_t va[16], vb[16];
                auto load = [&](int q, uint32_t (&v)[16]) { tmem_ld16_issue(acc + (uint32_t)((tile0 + TSTEP * (q % RPT)) * C + (q / RPT) * 16), v); };
                auto chunk = [&](const uint32_t (&v)[16], int q) {
                    const int c = q / RPT, j = q % RPT;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint8_t *p = dst + (2 * c + h) * LBO_A + row_off[j];
                        // packed pair adds (FADD2), ReLU + rounding in one F2FP.RELU: a third of the scalar version's instructions
                        const float4 b0 = *reinterpret_cast<const float4 *>(bias + c * 16 + h * 8), b1 = *reinterpret_cast<const float4 *>(bias + c * 16 + h * 8 + 4);
                        float2 f[4];
                        f[0] = fadd2(make_float2(__uint_as_float(v[h * 8]), __uint_as_float(v[h * 8 + 1])), make_float2(b0.x, b0.y));
                        f[1] = fadd2(make_float2(__uint_as_float(v[h * 8 + 2]), __uint_as_float(v[h * 8 + 3])), make_float2(b0.z, b0.w));
                        f[2] = fadd2(make_float2(__uint_as_float(v[h * 8 + 4]), __uint_as_float(v[h * 8 + 5])), make_float2(b1.x, b1.y));
                        f[3] = fadd2(make_float2(__uint_as_float(v[h * 8 + 6]), __uint_as_float(v[h * 8 + 7])), make_float2(b1.z, b1.w));
                        if (skip) {
                            const uint4 s = *reinterpret_cast<const uint4 *>(p);
                            f[0] = fadd2(f[0], unpack16<F16>(s.x));
                            f[1] = fadd2(f[1], unpack16<F16>(s.y));
                            f[2] = fadd2(f[2], unpack16<F16>(s.z));
                            f[3] = fadd2(f[3], unpack16<F16>(s.w));
                        }
                        uint4 o = make_uint4(0, 0, 0, 0);
                        if (valid[j]) o = make_uint4(pack16_relu<F16>(f[0]), pack16_relu<F16>(f[1]), pack16_relu<F16>(f[2]), pack16_relu<F16>(f[3]));
                        *reinterpret_cast<uint4 *>(p) = o;
                    }
                    if (j == RPT - 1) {
                        // this warp's rows of channels 16 c .. 16 c + 15 are in place for the tensor core (and its reads of those
                        // accumulator columns are complete): one arrival per warp
                        fence_before();
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (PAIR) mbar_arrive_cluster(chunk_lead + c * 8);
                            else mbar_arrive(chunk0 + c * 8);
                        }
                    }
                };
                load(0, va);
#pragma unroll
                for (int q = 0; q < NQ; q += 2) {
                    tmem_ld_wait();
                    load(q + 1, vb);
                    chunk(va, q);
                    tmem_ld_wait();
                    if (q + 2 < NQ) load(q + 2, va);
                    chunk(vb, q + 1);
                }
            }
            // ---- heads: [POS][35][42] fp32 in the Flatten() order of NCHW, then both FC layers on CUDA cores
            float *hact = reinterpret_cast<float *>(bufT + OFF_HACT);
            const float *hb = s_bias + n_conv * C;
            {
                if (PAIR) mbar_wait_cluster(mma_done, (gl0 + (uint32_t)n_conv) & 1u);
                else mbar_wait(mma_done, (gl0 + (uint32_t)n_conv) & 1u);
                fence_after();
#pragma unroll
                for (int j = 0; j < RPT; ++j) {
                    const uint32_t acc = lane_addr + (uint32_t)(n_conv & 1) * SETCOLS + (uint32_t)((tile0 + TSTEP * j) * C);
                    uint32_t v[32], w[16];
                    tmem_ld32_issue(acc, v);
                    tmem_ld16_issue(acc + 32, w);
                    tmem_ld_wait();
                    if (valid[j]) {
                        float *o = hact + pos[j] * NHU * 42 + yy[j] * c4::W + xx[j];
#pragma unroll
                        for (int c = 0; c < 32; ++c) o[c * 42] = fmaxf(__uint_as_float(v[c]) + hb[c], 0.f);
#pragma unroll
                        for (int c = 32; c < NHU; ++c) o[c * 42] = fmaxf(__uint_as_float(w[c - 32]) + hb[c], 0.f);
                    }
                }
            }
            fence_before();
            asm volatile("bar.sync 1, %0;" ::"n"(ETHREADS) : "memory");
            // thread t takes inputs k = t, t + ETHREADS, ...: every FC weight is read once per CTA (coalesced) and used for POS positions;
            // acc[p][j], j = 7 is the value head
            float acc[POS * 8];
#pragma unroll
            for (int i = 0; i < POS * 8; ++i) acc[i] = 0.f;
#pragma unroll
            for (int k0 = 0; k0 < 32 * 42; k0 += ETHREADS) {
                const int k = k0 + (int)tid;
                const bool in = k < 32 * 42;
                const int kk = in ? k : 0;
                float wj[7];
#pragma unroll
                for (int j = 0; j < 7; ++j) wj[j] = in ? __ldg(fc_policy_w + j * (32 * 42) + kk) : 0.f;
#pragma unroll
                for (int p = 0; p < POS; ++p) {
                    const float xv = hact[p * NHU * 42 + kk];
#pragma unroll
                    for (int j = 0; j < 7; ++j) acc[p * 8 + j] = fmaf(wj[j], xv, acc[p * 8 + j]);
                }
            }
            {
                const bool in = tid < 3 * 42;
                const int kk = in ? (int)tid : 0;
                const float wv = in ? __ldg(fc_value_w + kk) : 0.f;
#pragma unroll
                for (int p = 0; p < POS; ++p) acc[p * 8 + 7] = fmaf(wv, hact[p * NHU * 42 + 32 * 42 + kk], acc[p * 8 + 7]);
            }
            // warp reduction by recursive halving: each step exchanges half of the values; with V = POS * 8 values lane L ends
            // with the totals of the V / 32 indices L * V / 32 ...
            constexpr int V = POS * 8, VL = V / 32;
#pragma unroll
            for (int h = V / 2; h >= VL; h >>= 1) {
                const bool up = (lane & (uint32_t)(h / VL)) != 0;
#pragma unroll
                for (int i = 0; i < h; ++i) {
                    const float send = up ? acc[i] : acc[i + h];
                    const float keep = up ? acc[i + h] : acc[i];
                    acc[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, h / VL);
                }
            }
            float *red = reinterpret_cast<float *>(bufT + OFF_RED);  // [EW warps][V]
#pragma unroll
            for (int i = 0; i < VL; ++i) red[warp * V + VL * lane + i] = acc[i];
            asm volatile("bar.sync 1, %0;" ::"n"(ETHREADS) : "memory");
            if (tid < V) {
                float sum = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < EW; ++w8) sum += red[w8 * V + tid];
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
    }  // batches
    fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while the pair's MMAs may still touch it
    if (warp == 0) {
        if (PAIR) tmem_dealloc2(tmem_base, 2 * SETCOLS);
        else tmem_dealloc(tmem_base, 2 * SETCOLS);
    }
}

}  // namespace

// start offset of the second wave of co-resident CTAs, per layer of the net (AZ_PIPE_STAGGER_NS overrides; 0 = none)
static uint32_t stagger_ns() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("AZ_PIPE_STAGGER_NS");
        v = e ? atoi(e) : 0;  // measured (scripts/gpu/run9.sh): 0 / 600 / 1100 / 1600 / 2200 ns -> 0.536 / 0.543 / 0.550 / 0.548 / 0.563 ms per launch
    }
    return (uint32_t)v;
}

template <typename K>
static int32_t launch_pipe(az_engine *engine, const az_resnet_desc *d, float *logits, float *values, void *stream) {
    if (d->num_blocks < 0 || 1 + 2 * d->num_blocks > K::MAX_CONV) return AZ_E_INVALID;
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
        if (cudaFuncSetAttribute(k_resnet_pipe<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES) != cudaSuccess) return AZ_E_CUDA;
        if (cudaFuncSetAttribute(k_resnet_pipe<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES) != cudaSuccess) return AZ_E_CUDA;
        attr_set[dev] = true;
    }
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return AZ_E_CUDA;
    const int batches = (n + K::POS - 1) / K::POS;
    auto kern = d->operand_format == AZ_FMT_F16 ? k_resnet_pipe<K, true> : k_resnet_pipe<K, false>;
    const int resident = sms * K::CTAS;
    if (K::PAIR) {
        // clusters of two CTAs = the two SMs of a TPC; an odd tail gets an empty partner
        cudaLaunchConfig_t cfg = {};
        const int want = (batches + 1) / 2 * 2;
        cfg.gridDim = dim3((unsigned)(want < resident ? want : resident / 2 * 2));
        cfg.blockDim = dim3(K::THREADS);
        cfg.dynamicSmemBytes = K::SMEM_BYTES;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, kern, bb0, bb1, player, status, elist, ecount, (long long)n, (const uint8_t *)d->trunk_w, d->trunk_b, (int)d->num_blocks,
                               (const uint8_t *)d->head_conv_w, d->head_conv_b, d->fc_policy_w, d->fc_policy_b, d->fc_value_w, d->fc_value_b, logits, values,
                               stagger_ns()) != cudaSuccess)
            return AZ_E_CUDA;
        return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_E_CUDA;
    }
    kern<<<batches < resident ? batches : resident, K::THREADS, K::SMEM_BYTES, (cudaStream_t)stream>>>(
        bb0, bb1, player, status, elist, ecount, (long long)n, (const uint8_t *)d->trunk_w, d->trunk_b, d->num_blocks, (const uint8_t *)d->head_conv_w,
        d->head_conv_b, d->fc_policy_w, d->fc_policy_b, d->fc_value_w, d->fc_value_b, logits, values, stagger_ns());
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_E_CUDA;
}

extern "C" {

/* bytes of packed trunk weights for `num_blocks` residual blocks of 64 / 128 channels (models.py:pack_trunk_weights_pipe):
 * one piece [9 taps][C out][16 in] for the stem, C / 16 pieces per block convolution */
int64_t az_resnet_pipe_weight_bytes(int32_t num_blocks, int32_t num_channels) {
    if (num_channels == 128) return (int64_t)Cfg128::PIECE_BYTES * (1 + (int64_t)num_blocks * 2 * Cfg128::KS);
    if (num_channels == 64) return (int64_t)Cfg64::PIECE_BYTES * (1 + (int64_t)num_blocks * 2 * Cfg64::KS);
    return -1;
}

/* internal: called by az_resnet_forward_leaves_v2 (csrc/az_conv.cu) */
int32_t az_resnet_pipe_launch(az_engine *engine, const az_resnet_desc *d, float *logits, float *values, void *stream) {
    if (!engine || !d || !d->trunk_w || !d->trunk_b) return AZ_E_INVALID;
    if (d->operand_format != AZ_FMT_BF16 && d->operand_format != AZ_FMT_F16) return AZ_E_INVALID;
    if (d->num_channels == 128 && d->variant == 3) return launch_pipe<Cfg128p>(engine, d, logits, values, stream);
    if (d->num_channels == 128) return launch_pipe<Cfg128>(engine, d, logits, values, stream);
    if (d->num_channels == 64 && d->variant == 2) return launch_pipe<Cfg64x2>(engine, d, logits, values, stream);
    if (d->num_channels == 64 && d->variant == 3) return launch_pipe<Cfg64x2p>(engine, d, logits, values, stream);
    if (d->num_channels == 64) return launch_pipe<Cfg64>(engine, d, logits, values, stream);
    return AZ_E_INVALID;
}

}  // extern "C"
