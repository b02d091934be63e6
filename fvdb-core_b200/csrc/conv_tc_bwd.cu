// conv_tc_bwd.cu -- fused backward for narrow channels: dgrad AND wgrad from ONE gather of grad_output.
//
//   dX[i]        = sum_k dY[m_k(i)] . W[k]^T                         (GatherScatterDefault.cu:803-804)
//   dW[k][ci][co] = sum_i X[i][ci] * dY[m_k(i)][co]                   (:806-807, written input-stationary)
//
// with m_k(i) = the output row that input row i reaches through tap k (the reversed / input-stationary map).  Both sums
// consume the SAME gathered operand A_k = dY[m_k(rows of a tile)] -- a 128-row x 128-byte block in shared memory:
//   MMA1 (dgrad)  D1[tile][128 rows x CX]       += A_k (K-major, 64 reduction elements = G taps x CG)  . W^T chunk
//   MMA2 (wgrad)  D2[unit][(tap, co) 128 x CX]  += A_k^T (the same block, MN-major; reduction = 128 rows) . X tile
// The forward / dgrad / wgrad kernels of this engine are bound by the gather (LSU instruction issue + shared-memory fill:
// profiles/r02_*), not by the tensor pipe, so running dgrad and wgrad off one gather nearly halves the backward pass.  The
// price is tensor memory: every tap's dW accumulator must stay resident next to the dX tiles, 2 taps x CG lanes per 64-wide
// block.  That fits for the narrow shapes -- CG, CX in {16, 32}: 5^3 16 -> 16 needs 16 units x 16 columns + 2 x 16, 3^3
// 32 -> 32 needs 7 x 32 + 2 x 32 of the 512 columns -- which are exactly the shapes where the gather dominates most
// (BASELINE.json configs[4], and the 32-channel level of configs[2]).  Wider layers keep the two separate kernels.
//
// One persistent CTA per SM walks a contiguous chunk of 128-row input tiles:
//   warps 0-3   epilogue: drain D1 of a finished tile (tcgen05.ld -> bf16 -> dX rows) while the next tile is gathered --
//               D1 is double-buffered -- and, once per CTA, D2 -> the CTA's fp32 partial slice [K^3][CX][CG]
//   warps 4-11  gather producers: warp w gathers whole 16 KB blocks (32 x 16-byte zero-filling cp.async per lane), two blocks
//               = one stage; its 32 map entries per lane come from eight 16-byte loads issued one block ahead
//   warp 12     TMEM allocation + the single-thread tcgen05.mma issuer
//   warp 13     X-tile loader (128 contiguous rows, SWIZZLE_32B / 64B MN-major, double-buffered)
// The W^T image of the whole kernel (<= 64 KB) stays resident in shared memory.  Partial slices are summed by
// wgrad_reduce_partials in a fixed order: no atomics, run-to-run deterministic.
#include "conv_internal.cuh"
#include "tc_ptx.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace fvc {

using namespace tc;

constexpr int BF_TILE = 128;
constexpr int BF_BLOCK_BYTES = BF_TILE * 128;
constexpr int BF_PW = 8;                      // producer warps
constexpr int BF_WARP_PROD0 = 4, BF_WARP_MMA = BF_WARP_PROD0 + BF_PW, BF_WARP_X = BF_WARP_MMA + 1;
constexpr int BF_THREADS = (BF_WARP_X + 1) * 32;
constexpr int BF_MAX_UNITS = 16;

template <int CG, int CX> struct TcBwdCfg {
    static constexpr int G = 64 / CG;                  // taps per gathered block
    static constexpr int CPT = CG / 8;                 // 16-byte chunks one tap contributes to a 128-byte row
    static constexpr int XROW = CX * 2;                // bytes of an X-tile row (32 or 64): SWIZZLE_32B / SWIZZLE_64B MN-major
    static constexpr int XTILE = BF_TILE * XROW;
    static constexpr int WCHUNK = CX * 128;            // W^T image bytes per block: CX rows x 64 reduction elements
    // gather stages of one unit (two blocks) each.  STAGES >= BF_PW / 2 (the units the producer warps hold in flight) is a
    // correctness condition, not a tuning knob: a warp passes its wait for unit u - PW/2 once unit u - PW/2 - STAGES is consumed,
    // and its next wait (unit u) is only unambiguous if unit u - 2 STAGES is consumed by then -- a parity wait cannot tell a
    // barrier two phases behind from one that is ready, and the warp would overwrite a unit that was never consumed
    static constexpr int STAGES = 4;
    static constexpr int A_STAGE = 2 * BF_BLOCK_BYTES;
    static constexpr int NUM_BARS = 2 * STAGES + 10;   // full / empty per stage, 2 x (xfull, xempty, dxfull, dxempty), accum, weights
    static_assert((CG == 16 || CG == 32) && (CX == 16 || CX == 32), "fused backward serves 16 / 32 channels");
    static_assert(2 * STAGES >= BF_PW, "producer warps may run at most one stage phase ahead");
    static size_t smem_bytes(int blocks) { return 1024 + size_t(STAGES) * A_STAGE + align_up(size_t(blocks) * WCHUNK, 1024) + 2 * size_t(XTILE) + 8 * NUM_BARS + 64; }
};

// shared-memory matrix descriptor with an explicit layout type (2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B)
__device__ __forceinline__ uint64_t make_smem_desc_any(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);
    d |= uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(layout_type) << 61;
    return d;
}

// mbarrier wait that names the waiting site when it times out (debug runs: FVC_DEBUG_WAITS=1 hands the kernel a host-mapped
// report slot and synchronises after the launch; otherwise `report` is null and a protocol bug just traps)
__device__ __forceinline__ void bf_wait(uint32_t bar, uint32_t parity, int site, int a, int b, int *report) {
    uint32_t done = 0;
    long long start = 0;
    while (true) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done)
            break;
        const long long now = clock64();
        if (start == 0)
            start = now;
        else if (now - start > 2000000000ll) {
            if (report && (threadIdx.x & 31) == 0 && atomicCAS(report, 0, site) == 0) {
                report[1] = int(blockIdx.x), report[2] = int(threadIdx.x >> 5), report[3] = int(parity), report[4] = a, report[5] = b;
                __threadfence_system();
            }
            __trap();
        }
    }
}

template <int CG, int CX>
__global__ void __launch_bounds__(BF_THREADS, 1)
conv_tc_bwd_fused_kernel(const uint16_t *__restrict__ dy, const uint16_t *__restrict__ x, const uint8_t *__restrict__ w_img,
                         const int32_t *__restrict__ map, int64_t pitch, const unsigned long long *__restrict__ tile_mask, int64_t n_in,
                         int k3, int tiles_per_chunk, int flip_taps, int is_bf16, uint32_t idesc1, uint32_t idesc2, void *__restrict__ dx_,
                         float *__restrict__ partial, int *report) {
    using Cfg = TcBwdCfg<CG, CX>;
    constexpr int G = Cfg::G, CPT = Cfg::CPT, STAGES = Cfg::STAGES, XROW = Cfg::XROW;
    extern __shared__ uint8_t smem_raw[];
    const int total_blocks = (k3 + G - 1) / G;
    const int nunits = (total_blocks + 1) / 2;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_w = smem_a + STAGES * Cfg::A_STAGE;
    const uint32_t smem_x = smem_w + ((uint32_t(total_blocks) * Cfg::WCHUNK + 1023u) & ~1023u);
    const uint32_t bars = smem_x + 2 * Cfg::XTILE;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * STAGES;
    const uint32_t bar_xfull = bar_empty + 8 * STAGES, bar_xempty = bar_xfull + 16;
    const uint32_t bar_dxfull = bar_xempty + 16, bar_dxempty = bar_dxfull + 16;
    const uint32_t bar_accum = bar_dxempty + 16, bar_w = bar_accum + 8;
    const uint32_t tmem_slot = bar_w + 8;
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (smem_base - smem_u32(smem_raw)) + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t total_tiles = (n_in + BF_TILE - 1) / BF_TILE;
    const int64_t tile_begin = int64_t(blockIdx.x) * tiles_per_chunk;
    const int64_t tile_end = tile_begin + tiles_per_chunk < total_tiles ? tile_begin + tiles_per_chunk : total_tiles;

    // live units of a tile from its tap bitmask (K^3 <= 128; larger kernels do not skip)
    const int words = (k3 + 63) >> 6;
    const bool use_mask = tile_mask != nullptr && words <= 2;
    __shared__ unsigned long long s_unit_taps[BF_MAX_UNITS][2], s_block_taps[2 * BF_MAX_UNITS][2];
    if (threadIdx.x < 2 * BF_MAX_UNITS) {
        const int blk = threadIdx.x;
        unsigned long long lo = 0ull, hi = 0ull;
        for (int tap = blk * G; tap < (blk + 1) * G && tap < k3 && tap < 128; ++tap)
            (tap < 64 ? lo : hi) |= 1ull << (tap & 63);
        s_block_taps[blk][0] = lo;
        s_block_taps[blk][1] = hi;
    }
    __syncthreads();
    if (threadIdx.x < BF_MAX_UNITS) {
        s_unit_taps[threadIdx.x][0] = s_block_taps[2 * threadIdx.x][0] | s_block_taps[2 * threadIdx.x + 1][0];
        s_unit_taps[threadIdx.x][1] = s_block_taps[2 * threadIdx.x][1] | s_block_taps[2 * threadIdx.x + 1][1];
    }
    auto tile_words = [&](int64_t tile, unsigned long long &m0, unsigned long long &m1) {
        m0 = use_mask ? __ldg(tile_mask + tile * words) : ~0ull;
        m1 = use_mask ? (words > 1 ? __ldg(tile_mask + tile * words + 1) : 0ull) : ~0ull;
    };
    auto live_units = [&](unsigned long long m0, unsigned long long m1) -> uint32_t {
        uint32_t live = 0;
        for (int u = 0; u < nunits; ++u)
            live |= uint32_t(((m0 & s_unit_taps[u][0]) | (m1 & s_unit_taps[u][1])) != 0ull) << u;
        return live;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 64); // two gathering warps x 32 completion-triggered arrivals
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_xfull + 8 * b, 32);
            mbar_init(bar_xempty + 8 * b, 1);
            mbar_init(bar_dxfull + 8 * b, 1);
            mbar_init(bar_dxempty + 8 * b, 128); // the four epilogue warps
        }
        mbar_init(bar_accum, 1);
        mbar_init(bar_w, 1);
        fence_mbar_init();
    }
    if (warp == BF_WARP_MMA)
        tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t d1_col0 = uint32_t(nunits * CX); // D2 units first, then the two D1 tile buffers

    if (warp < 4) {
        // ================= epilogue: dX tiles as they finish, then the dW accumulators =================
        const bool bf16 = is_bf16 != 0;
        int tb = 0;
        for (int64_t tile = tile_begin; tile < tile_end; ++tile) {
            unsigned long long m0, m1;
            tile_words(tile, m0, m1);
            const bool live = live_units(m0, m1) != 0u;
            const int64_t row = tile * BF_TILE + warp * 32 + lane;
            uint32_t acc[32];
            if (live) {
                const int b = tb & 1;
                bf_wait(bar_dxfull + 8 * b, (tb >> 1) & 1, 1, tb, int(tile - tile_begin), report);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (uint32_t(warp * 32) << 16) + d1_col0 + uint32_t(b * CX);
                if (CX == 32)
                    tmem_ld_32x32b_x32(taddr, acc);
                else
                    tmem_ld_32x32b_x16(taddr, acc);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(bar_dxempty + 8 * b); // the accumulator may be overwritten by the tile after next
                ++tb;
            } else { // no tap reaches this tile: its gradient rows are zero (GatherScatterDefault.cu:771-777)
#pragma unroll
                for (int z = 0; z < 32; ++z)
                    acc[z] = 0u;
            }
            if (row < n_in) {
                uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(dx_) + row * CX);
#pragma unroll
                for (int v4 = 0; v4 < CX / 8; ++v4) {
                    uint32_t p[4];
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const float a = __uint_as_float(acc[v4 * 8 + 2 * h]), c = __uint_as_float(acc[v4 * 8 + 2 * h + 1]);
                        if (bf16) {
                            __nv_bfloat162 hv = __floats2bfloat162_rn(a, c);
                            p[h] = *reinterpret_cast<uint32_t *>(&hv);
                        } else {
                            __half2 hv = __floats2half2_rn(a, c);
                            p[h] = *reinterpret_cast<uint32_t *>(&hv);
                        }
                    }
                    dst[v4] = make_uint4(p[0], p[1], p[2], p[3]);
                }
            }
        }
        // ---- dW: TMEM lanes 32 w .. 32 w + 31 of unit u = reduction-side element (tap, co) of its block `warp >> 1` ----
        bf_wait(bar_accum, 0, 2, tb, 0, report);
        tc_fence_after();
        const uint32_t started = *reinterpret_cast<volatile uint32_t *>(smem_raw + (smem_base - smem_u32(smem_raw)) + (tmem_slot + 4 - smem_base));
        float *slice = partial + int64_t(blockIdx.x) * k3 * CX * CG;
        const int half = warp >> 1;            // which of the unit's two blocks
        const int kk = (warp & 1) * 32 + lane; // element inside the block: (sub tap, co)
        for (int u = 0; u < nunits; ++u) {
            const int blk = 2 * u + half;
            const int tap = blk * G + kk / CG, co = kk % CG;
            const bool live_row = blk < total_blocks && tap < k3;
            uint32_t acc[32];
            if ((started >> u) & 1u) {
                const uint32_t taddr = tmem_base + (uint32_t(warp * 32) << 16) + uint32_t(u * CX);
                if (CX == 32)
                    tmem_ld_32x32b_x32(taddr, acc);
                else
                    tmem_ld_32x32b_x16(taddr, acc);
                tmem_ld_wait();
            } else { // no tile of this CTA reached the unit's taps: its accumulator was never written
#pragma unroll
                for (int z = 0; z < 32; ++z)
                    acc[z] = 0u;
            }
            if (live_row) {
                const int tap_out = flip_taps ? k3 - 1 - tap : tap; // a mirrored (symmetric) map walks the taps backwards
                float *dst = slice + int64_t(tap_out) * CX * CG + co;
#pragma unroll
                for (int ci = 0; ci < CX; ++ci)
                    dst[ci * CG] = __uint_as_float(acc[ci]); // consecutive lanes = consecutive co: coalesced per ci
            }
        }
        tc_fence_before();
    } else if (warp < BF_WARP_MMA) {
        // ================= gather producers: warp pw gathers whole blocks, item it -> warp it % BF_PW =================
        // lane = (lg, q): 8 lanes q cover one 128-byte row; lane group lg owns the 32 consecutive rows 32 lg .. 32 lg + 31 and
        // copies 16-byte chunk q of each: chunk q belongs to tap `sub` of the block and to channel chunk q % CPT of that
        // tap's grad_output row.  The 32 map entries (rows 32 lg .. of tap `sub`) are eight 16-byte loads, one block ahead.
        const int pw = warp - BF_WARP_PROD0;
        const int q = lane & 7, lg = lane >> 3, sub = q / CPT;
        const uint32_t lane_off = (uint32_t(lg) << 12) | (uint32_t(q) << 4);
        const uint16_t *dyq = dy + (q % CPT) * 8;
        int it = 0; // block items of this CTA so far (every producer warp counts them all)
        int4 idx[8], nxt[8];
        auto load_idx = [&](int64_t tile, int blk, bool blk_live, int4 (&out)[8]) {
            const int tap = blk * G + sub;
            if (blk_live && blk < total_blocks && tap < k3) {
                const int4 *src = reinterpret_cast<const int4 *>(map + int64_t(tap) * pitch + tile * BF_TILE + lg * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    out[j] = __ldg(src + j);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    out[j] = make_int4(-1, -1, -1, -1);
            }
        };
        // walk (tile, live unit, block) in the order every role uses; keep one of this warp's items prefetched
        struct Item {
            int64_t tile;
            int blk, stage, use;
            bool live;
        };
        auto next_item = [&](int64_t &tile, uint32_t &rest, unsigned long long &m0, unsigned long long &m1, int &half, int &unit_no, Item &out) -> bool {
            // advances the shared enumeration by one block; returns false at the end of the chunk
            while (true) {
                if (tile >= tile_end)
                    return false;
                if (rest == 0u && half == 0) {
                    ++tile;
                    if (tile >= tile_end)
                        return false;
                    tile_words(tile, m0, m1);
                    rest = live_units(m0, m1);
                    continue;
                }
                const int u = __ffs(rest) - 1;
                out.tile = tile;
                out.blk = 2 * u + half;
                out.live = ((m0 & s_block_taps[out.blk][0]) | (m1 & s_block_taps[out.blk][1])) != 0ull;
                out.stage = unit_no % STAGES;
                out.use = unit_no / STAGES;
                if (half == 1) {
                    rest &= rest - 1u;
                    ++unit_no;
                }
                half ^= 1;
                return true;
            }
        };
        int64_t e_tile = tile_begin - 1;
        uint32_t e_rest = 0u;
        unsigned long long e_m0 = 0ull, e_m1 = 0ull;
        int e_half = 0, e_unit = 0;
        Item cur, fut;
        bool have = false;
        // find this warp's first item
        while (next_item(e_tile, e_rest, e_m0, e_m1, e_half, e_unit, cur)) {
            if (it++ % BF_PW == pw) {
                have = true;
                break;
            }
        }
        if (have)
            load_idx(cur.tile, cur.blk, cur.live, idx);
        while (have) {
            bool have_next = false;
            while (next_item(e_tile, e_rest, e_m0, e_m1, e_half, e_unit, fut)) {
                if (it++ % BF_PW == pw) {
                    have_next = true;
                    break;
                }
            }
            if (have_next)
                load_idx(fut.tile, fut.blk, fut.live, nxt);
            const int64_t rows_left = n_in - cur.tile * BF_TILE - lg * 32; // row 32 lg + i exists iff i < rows_left
            bf_wait(bar_empty + 8 * cur.stage, (cur.use & 1) ^ 1u, 3, cur.use * STAGES + cur.stage, cur.blk, report);
            const uint32_t dst = smem_a + cur.stage * Cfg::A_STAGE + (cur.blk & 1) * BF_BLOCK_BYTES;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int v[4] = {idx[j].x, idx[j].y, idx[j].z, idx[j].w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int i = 4 * j + c;
                    const bool ok = v[c] >= 0 && i < rows_left;
                    cp_async16(dst + (lane_off ^ uint32_t(i * 128 + ((i & 7) << 4))), ok ? dyq + int64_t(v[c]) * CG : dy, ok ? 16u : 0u);
                }
            }
            cp_async_arrive_noinc(bar_full + 8 * cur.stage);
            cur = fut;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                idx[j] = nxt[j];
            have = have_next;
        }
        cp_async_wait_all();
    } else if (warp == BF_WARP_MMA) {
        // ================= MMA issuer (one thread): resident W^T image, then per live unit MMA1 x 8 + MMA2 x 8 =================
        if (elect_one()) {
            const uint32_t w_bytes = uint32_t(total_blocks) * Cfg::WCHUNK;
            mbar_expect_tx(bar_w, w_bytes);
            for (uint32_t off = 0; off < w_bytes; off += 32768u) // bulk copies of at most 32 KB
                bulk_g2s(smem_w + off, w_img + off, (w_bytes - off < 32768u ? w_bytes - off : 32768u), bar_w);
            bf_wait(bar_w, 0, 4, 0, 0, report);
            const uint64_t hi_k = make_smem_desc_any(0, 16, 1024, 2) & 0xFFFFFFFF00000000ull;            // K-major SWIZZLE_128B (A of MMA1, W^T)
            const uint64_t hi_mn = make_smem_desc_any(0, BF_BLOCK_BYTES, 1024, 2) & 0xFFFFFFFF00000000ull; // MN-major SWIZZLE_128B (A of MMA2)
            const uint32_t lbo_k = 1u << 16, lbo_mn = uint32_t(BF_BLOCK_BYTES >> 4) << 16;
            // X tile: MN-major, CX elements (XROW bytes) per row, 8-row atoms of 8 * XROW bytes
            const uint64_t hi_x = make_smem_desc_any(0, 16, 8 * XROW, XROW == 32 ? 6 : 4) & 0xFFFFFFFF00000000ull;
            const uint32_t x_step = uint32_t(16 * XROW) >> 4; // 16 rows per MMA
            int unit_no = 0, tb = 0, tx = 0;
            uint32_t started = 0; // dW accumulators written so far
            for (int64_t tile = tile_begin; tile < tile_end; ++tile) {
                unsigned long long m0, m1;
                tile_words(tile, m0, m1);
                const uint32_t live = live_units(m0, m1);
                if (live == 0u)
                    continue;
                const int xb = tx & 1, db = tb & 1;
                bf_wait(bar_xfull + 8 * xb, (tx >> 1) & 1, 5, tx, int(tile - tile_begin), report);
                bf_wait(bar_dxempty + 8 * db, ((tb >> 1) & 1) ^ 1u, 6, tb, int(tile - tile_begin), report);
                tc_fence_after();
                const uint32_t d1 = tmem_base + d1_col0 + uint32_t(db * CX);
                const uint32_t x_lo = (((smem_x + xb * Cfg::XTILE) & 0x3FFFFu) >> 4) | (1u << 16);
                bool first = true;
                for (uint32_t rest = live; rest; rest &= rest - 1u, ++unit_no) {
                    const int u = __ffs(rest) - 1;
                    const int s = unit_no % STAGES;
                    bf_wait(bar_full + 8 * s, (unit_no / STAGES) & 1, 7, unit_no, u, report);
                    fence_proxy_async(); // cp.async (generic proxy) filled the stage; tcgen05.mma reads through the async proxy
                    tc_fence_after();
                    const uint32_t a_addr = smem_a + s * Cfg::A_STAGE;
                    // MMA1: dX tile += block . W^T chunk, for the unit's two blocks (a block past the kernel volume is all zero)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int blk = 2 * u + h;
                        if (blk >= total_blocks)
                            break;
                        const uint32_t a_lo = (((a_addr + h * BF_BLOCK_BYTES) & 0x3FFFFu) >> 4) | lbo_k;
                        const uint32_t b_lo = (((smem_w + blk * Cfg::WCHUNK) & 0x3FFFFu) >> 4) | lbo_k;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_f16(d1, hi_k | (a_lo + 2 * kk), hi_k | (b_lo + 2 * kk), idesc1, uint32_t(!(first && kk == 0)));
                        first = false;
                    }
                    // MMA2: dW unit += (both blocks)^T . X tile, reduction over the tile's 128 rows in 8 steps of 16
                    const uint32_t a_mn = ((a_addr & 0x3FFFFu) >> 4) | lbo_mn;
                    const uint32_t acc0 = (started >> u) & 1u;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        umma_f16(tmem_base + uint32_t(u * CX), hi_mn | (a_mn + 128 * kk), hi_x | (x_lo + x_step * kk), idesc2, acc0 | uint32_t(kk != 0));
                    started |= 1u << u;
                    umma_commit(bar_empty + 8 * s);
                }
                umma_commit(bar_dxfull + 8 * db);
                umma_commit(bar_xempty + 8 * xb);
                ++tb;
                ++tx;
            }
            // accumulators no tile of this chunk reached were never written: the drain takes `started` as the validity mask
            *reinterpret_cast<volatile uint32_t *>(smem_raw + (smem_base - smem_u32(smem_raw)) + (tmem_slot + 4 - smem_base)) = started;
            __threadfence_block();
            umma_commit(bar_accum);
        }
        __syncwarp();
    } else {
        // ================= X-tile loader: 128 contiguous rows, XROW bytes each, swizzled for the MN-major B operand of MMA2 =================
        int tx = 0;
        for (int64_t tile = tile_begin; tile < tile_end; ++tile) {
            unsigned long long m0, m1;
            tile_words(tile, m0, m1);
            if (live_units(m0, m1) == 0u)
                continue;
            const int xb = tx & 1;
            bf_wait(bar_xempty + 8 * xb, ((tx >> 1) & 1) ^ 1u, 8, tx, int(tile - tile_begin), report);
            const uint32_t dst = smem_x + xb * Cfg::XTILE;
            constexpr int CHUNKS = XROW / 16; // 16-byte chunks per row
            for (int e = lane; e < BF_TILE * CHUNKS; e += 32) {
                const int r = e / CHUNKS, c = e % CHUNKS;
                const int64_t row = tile * BF_TILE + r;
                // SWIZZLE_32B: chunk ^= (address >> 7) & 1; SWIZZLE_64B: chunk ^= (address >> 7) & 3  (atoms are 8 rows, 1024-byte aligned base)
                const int sw = CHUNKS == 2 ? ((r >> 2) & 1) : ((r >> 1) & 3);
                const bool ok = row < n_in;
                cp_async16(dst + r * XROW + ((c ^ sw) << 4), ok ? x + row * CX + c * 8 : x, ok ? 16u : 0u);
            }
            cp_async_arrive_noinc(bar_xfull + 8 * xb);
            ++tx;
        }
        cp_async_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == BF_WARP_MMA)
        tmem_dealloc(tmem_base, 512);
}

// ---- host side ------------------------------------------------------------------------------------
bool tc_bwd_fused_supported(int32_t cin, int32_t cout, int64_t k3, int32_t dtype) {
    if (dtype != FVC_F16 && dtype != FVC_BF16)
        return false;
    if (!((cin == 16 || cin == 32) && (cout == 16 || cout == 32)) || k3 < 1 || k3 > 128)
        return false;
    const int g = 64 / cout, blocks = int(ceil_div(k3, g)), units = (blocks + 1) / 2;
    if (units > BF_MAX_UNITS || units * cin + 2 * cin > 512)
        return false;
    const size_t w_bytes = align_up(size_t(blocks) * size_t(cin) * 128, 1024);
    return 1024 + 4 * 2 * BF_BLOCK_BYTES + w_bytes + 2 * size_t(BF_TILE) * cin * 2 + 1024 <= 232448 - 1024;
}

static inline int bwd_chunks(int64_t n_in) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = ceil_div(n_in > 0 ? n_in : 1, BF_TILE);
    return int(tiles < sms ? tiles : sms);
}

size_t tc_bwd_fused_scratch_bytes(int64_t n_in, int32_t cin, int32_t cout, int64_t k3) {
    return align_up(size_t(bwd_chunks(n_in)) * size_t(k3) * size_t(cin) * size_t(cout) * 4, 256) + 256;
}

template <int CG, int CX> static int launch_bwd_fused(const BwdFusedArgs &a) {
    using Cfg = TcBwdCfg<CG, CX>;
    auto kernel = conv_tc_bwd_fused_kernel<CG, CX>;
    const int blocks = int(ceil_div(a.k3, Cfg::G));
    const size_t smem = Cfg::smem_bytes(blocks);
    static std::atomic<unsigned long long> configured{0};
    // the resident weight image makes the shared-memory size depend on the kernel volume: opt in to the maximum once per device
    const int rc = ensure_dynamic_smem(kernel, 232448 - 1024, configured);
    if (rc)
        return rc;
    const int chunks = bwd_chunks(a.n_in);
    const int64_t tiles = ceil_div(a.n_in, BF_TILE);
    const int tiles_per_chunk = int(ceil_div(tiles, chunks));
    const int grid = int(ceil_div(tiles, tiles_per_chunk));
    const bool bf16 = a.dtype == FVC_BF16;
    const uint32_t idesc1 = make_idesc_f16(128, CX, bf16, false, false), idesc2 = make_idesc_f16(128, CX, bf16, true, true);
    float *partial = reinterpret_cast<float *>(a.scratch);
    static int *report = nullptr; // debug runs only: a host-mapped slot the kernel names a timed-out wait in
    static const bool debug = getenv("FVC_DEBUG_WAITS") != nullptr; // read once per process
    if (debug && !report && cudaHostAlloc(reinterpret_cast<void **>(&report), 64, cudaHostAllocMapped) != cudaSuccess)
        report = nullptr;
    if (debug && report)
        memset(report, 0, 64);
    kernel<<<grid, BF_THREADS, smem, a.stream>>>(reinterpret_cast<const uint16_t *>(a.dy), reinterpret_cast<const uint16_t *>(a.x),
                                                 reinterpret_cast<const uint8_t *>(a.w_img), a.map, a.pitch,
                                                 reinterpret_cast<const unsigned long long *>(a.tile_mask), a.n_in, a.k3, tiles_per_chunk,
                                                 a.flip_taps, bf16 ? 1 : 0, idesc1, idesc2, a.dx, partial, debug ? report : nullptr);
    FVC_LAUNCH_CHECK();
    if (debug) {
        const cudaError_t err = cudaStreamSynchronize(a.stream);
        fprintf(stderr, "[fvc debug] fused backward <%d,%d> grid %d tiles/chunk %d k3 %d smem %zu: %s; wait report site %d block %d warp %d parity %d a %d b %d\n", CG, CX,
                grid, tiles_per_chunk, a.k3, smem, cudaGetErrorString(err), report ? report[0] : -1, report ? report[1] : -1, report ? report[2] : -1,
                report ? report[3] : -1, report ? report[4] : -1, report ? report[5] : -1);
    }
    return wgrad_reduce_partials(partial, grid, CX, CG, a.k3, a.dtype, a.grad_w, a.stream);
}

int tc_bwd_fused(const BwdFusedArgs &a) {
    FVC_REQUIRE(tc_bwd_fused_supported(a.cin, a.cout, a.k3, a.dtype), FVC_ERR_UNSUPPORTED, "fused backward does not admit dtype code %d with channels %d -> %d, %d taps",
                a.dtype, a.cin, a.cout, a.k3);
    FVC_REQUIRE(a.scratch && a.scratch_bytes >= tc_bwd_fused_scratch_bytes(a.n_in, a.cin, a.cout, a.k3), FVC_ERR_RUNTIME, "fused backward scratch too small");
    FVC_REQUIRE((reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.dx) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a.map) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.w_img) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.scratch) & 255) == 0,
                FVC_ERR_RUNTIME, "fused backward needs 16-byte aligned operands and 256-byte aligned scratch");
    FVC_REQUIRE(a.pitch % 4 == 0 && a.pitch >= ceil_div(a.n_in, BF_TILE) * BF_TILE, FVC_ERR_RUNTIME,
                "fused backward needs the map pitch (%lld) to be a multiple of 4 covering whole 128-row tiles", (long long)a.pitch);
    // gathered channels CG = the public Cout (rows of grad_output), CX = the public Cin (rows of the features)
    if (a.cout == 16 && a.cin == 16)
        return launch_bwd_fused<16, 16>(a);
    if (a.cout == 16 && a.cin == 32)
        return launch_bwd_fused<16, 32>(a);
    if (a.cout == 32 && a.cin == 16)
        return launch_bwd_fused<32, 16>(a);
    return launch_bwd_fused<32, 32>(a);
}

} // namespace fvc
