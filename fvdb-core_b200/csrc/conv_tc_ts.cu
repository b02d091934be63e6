// conv_tc_ts.cu -- forward / dgrad for NARROW layers (Cin, Cout in {16, 32}, f16 / bf16): the gathered operand goes through
// tensor memory, not through a zero-filled shared-memory tile.
//
// Why a second executor.  conv_tc.cu writes every (row, tap) slot of a 128-row x 64-element operand block into shared
// memory -- 16-byte zero-filling cp.async for the missing neighbours -- and tcgen05.mma reads the whole block back.  On the
// bandwidth-bound small-channel workloads most slots are misses (BASELINE.json configs[4]: 5^3 at 20 % occupancy, 80 % of the
// slots), so the kernel is bound by shared-memory bandwidth spent on zeros (profiles/r01_ncu_full_c5_v2_summary.txt: l1tex
// 72 %, tensor pipe 8 %, DRAM far from peak).  Here
//   * one gather THREAD owns one output row of the tile (= one TMEM lane).  For every reduction block (G = 64 / Cin taps x Cin
//     channels) it issues cp.async ONLY for the taps that hit (32 or 64 bytes each, through L1: rows shared by neighbouring
//     taps / tiles of the same CTA can hit) into its private slice of a staging buffer, DEPTH units ahead of their use;
//   * when a unit's copies have landed (cp.async.wait_group -- the thread reads back only what it wrote itself: no
//     cross-thread hand-off), it loads the hits into registers, zeros for the misses, and writes its 128-byte operand row
//     with ONE tcgen05.st into an A-operand stage in tensor memory (lane = row, column j = reduction elements 2j, 2j + 1);
//   * per unit 4 x tcgen05.mma kind::f16 (M = 128, N = Cout, K = 16) with A FROM TENSOR MEMORY and B = the resident weight
//     image (the whole kernel's W, <= 64 KB, loaded once per CTA by cp.async.bulk);
//   * the CTA runs NG = 4 independent pipelines: gather group g (four warps, one per TMEM lane quarter) serves the units
//     u = g (mod NG) of the CTA's unit sequence, its own MMA-issuing warp accumulates them into its own partial accumulator
//     (measured: ONE issuing thread needs ~850 cycles of dependent instructions per unit -- uniform-register moves, mbarrier
//     round trip -- and was the bottleneck of the single-pipeline version, profiles/r02_ncu_ts_v2_hot.txt);
//   * accumulators are double-buffered in TMEM: four epilogue warps drain tile t -- tcgen05.ld of the NG partials, summed in a
//     fixed order -> bias / scale / shift / residual / ReLU / statistics -> 128-bit stores -- while tile t + 1 accumulates.
// Shared memory carries only the hits (one write + one read of real bytes); misses cost a predicate and a register move.
// One persistent CTA per SM walks a contiguous chunk of tiles (so gathered rows re-used by the next tile are still in L1).
// Output-stationary, no atomics: run-to-run deterministic like conv_tc.cu.  Reference semantics: GatherScatterDefault.cu:706-721
// (forward) and :803-804 (dgrad, on the input-stationary map with W^T).
#include "conv_internal.cuh"
#include "tc_math.cuh"
#include "tc_ptx.cuh"

namespace fvc {

using namespace tc;

constexpr int TS_TILE = 128;
constexpr int TS_NG = 4;    // pipelines: gather group (four warps, one per TMEM lane quarter) + MMA warp + partial accumulator
constexpr int TS_DEPTH = 2; // units a gather thread keeps in flight (cp.async groups)
constexpr int TS_SAG = 2;   // A-operand stages in tensor memory per pipeline, 32 columns each
constexpr int TS_WARP_G0 = 4, TS_WARP_MMA0 = TS_WARP_G0 + 4 * TS_NG, TS_THREADS = (TS_WARP_MMA0 + TS_NG) * 32;
constexpr int TS_UNIT_BYTES = TS_TILE * 128; // staging of one unit: G taps x 128 rows x Cin x 2 B
constexpr int TS_A_COL0 = 256;               // TMEM columns [0, 256): NG x 2 accumulator buffers of <= 32 columns; then the A stages
constexpr int TS_MAX_BLOCKS = 32;            // reduction blocks of one kernel (live-block mask of a tile = one 32-bit word)
constexpr int TS_NUM_BARS = 2 * TS_NG * TS_SAG + 4 + 1;
constexpr int TS_MAP_RING = 4;               // units of map entries a gather warp keeps in its shared-memory ring
static_assert((TS_NG & (TS_NG - 1)) == 0 && TS_A_COL0 + 32 * TS_NG * TS_SAG <= 512 && TS_NG * 2 * 32 <= TS_A_COL0, "tensor memory");

template <int CIN, int COUT> struct TsCfg {
    static constexpr int G = 64 / CIN;       // taps per reduction block
    static constexpr int ROWB = CIN * 2;     // bytes one tap contributes to an operand row (32 / 64)
    static constexpr int CH = ROWB / 16;     // 16-byte chunks of it
    static constexpr int WCHUNK = COUT * 128; // weight image bytes per block: COUT rows x 64 reduction elements
    static_assert((CIN == 16 || CIN == 32) && (COUT == 16 || COUT == 32), "the tensor-memory executor serves 16 / 32 channels");
    static size_t smem_bytes(int blocks) { return 1024 + size_t(TS_NG) * TS_DEPTH * TS_UNIT_BYTES + align_up(size_t(blocks) * WCHUNK, 1024) + 8 * TS_NUM_BARS + 128 + 8 * COUT * 4 + size_t(4 * TS_NG) * TS_MAP_RING * G * 128; }
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(TS_THREADS, 1)
conv_tc_ts_kernel(const uint16_t *__restrict__ x, const uint8_t *__restrict__ w_img, const Epilogue epi, void *__restrict__ y_,
                  const int32_t *__restrict__ nbr, int64_t pitch, const unsigned long long *__restrict__ tile_mask, int64_t n_out, int k3,
                  int tiles_per_chunk, uint32_t idesc, int is_bf16) {
    using Cfg = TsCfg<CIN, COUT>;
    constexpr int G = Cfg::G, ROWB = Cfg::ROWB, CH = Cfg::CH;
    extern __shared__ uint8_t smem_raw[];
    const int total_blocks = (k3 + G - 1) / G;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_stage = smem_base;
    const uint32_t smem_w = smem_stage + TS_NG * TS_DEPTH * TS_UNIT_BYTES;
    const uint32_t bars = smem_w + ((uint32_t(total_blocks) * Cfg::WCHUNK + 1023u) & ~1023u);
    const uint32_t bar_afull = bars, bar_aempty = bars + 8 * TS_NG * TS_SAG; // [pipeline][stage]
    const uint32_t bar_dfull = bar_aempty + 8 * TS_NG * TS_SAG, bar_dempty = bar_dfull + 16;
    const uint32_t bar_w = bar_dempty + 16;
    const uint32_t tmem_slot = bar_w + 8;
    const uint32_t smem_zero = tmem_slot + 56;   // 64 bytes of zeros (16-byte aligned: the barrier block is a multiple of 8 bytes, + 8 + 56)
    const uint32_t smem_stats = smem_zero + 64;  // [4 quarters][2][COUT] floats
    const uint32_t smem_map = smem_stats + 8 * COUT * 4; // [gather warp][TS_MAP_RING] map-entry slots (16-byte aligned)
    uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = int((n_out + TS_TILE - 1) / TS_TILE);
    const int tile_begin = int(blockIdx.x) * tiles_per_chunk;
    const int tile_end = tile_begin + tiles_per_chunk < total_tiles ? tile_begin + tiles_per_chunk : total_tiles;
    const int full_tiles = int(n_out / TS_TILE), tail_rows = int(n_out % TS_TILE); // row r of tile t exists iff t < full_tiles || r < tail_rows

    // live reduction blocks of a tile from its tap bitmask (bit b: some tap of block b reaches some row of the tile)
    const int words = (k3 + 63) >> 6;
    const bool use_mask = tile_mask != nullptr && words <= 2;
    const uint32_t all_blocks = total_blocks >= 32 ? 0xffffffffu : ((1u << total_blocks) - 1u);
    // bit b of compress(word): any of the G tap bits of block b (64 % G == 0: a block never straddles the two mask words)
    auto compress = [](unsigned long long m) -> uint32_t {
        m |= m >> 1;
        if (G == 4) {
            m |= m >> 2;
            m &= 0x1111111111111111ull;
            m = (m | (m >> 3)) & 0x0303030303030303ull;
            m = (m | (m >> 6)) & 0x000F000F000F000Full;
            m = (m | (m >> 12)) & 0x000000FF000000FFull;
            m = (m | (m >> 24)) & 0xFFFFull;
        } else {
            m &= 0x5555555555555555ull;
            m = (m | (m >> 1)) & 0x3333333333333333ull;
            m = (m | (m >> 2)) & 0x0F0F0F0F0F0F0F0Full;
            m = (m | (m >> 4)) & 0x00FF00FF00FF00FFull;
            m = (m | (m >> 8)) & 0x0000FFFF0000FFFFull;
            m = (m | (m >> 16)) & 0xFFFFFFFFull;
        }
        return uint32_t(m);
    };
    auto tile_live = [&](int tile) -> uint32_t {
        if (!use_mask)
            return all_blocks;
        uint32_t live = compress(__ldg(tile_mask + int64_t(tile) * words));
        if (G == 4 && words > 1)
            live |= compress(__ldg(tile_mask + int64_t(tile) * words + 1)) << 16;
        return live & all_blocks;
    };
    // The CTA's units = (tile ascending, live block ascending), numbered u = 0, 1, ...; pipeline g owns u = g (mod NG).  Every role
    // walks a tile with `take`: `rest` = live blocks not yet passed, `skip` = units of other pipelines before this one's next unit
    // (at the start of a tile whose first unit has number `base`: skip = (g - base) mod NG).
    auto take = [](uint32_t &rest, int &skip, int &blk) -> bool {
        if (__popc(rest) <= skip) {
            skip -= __popc(rest);
            rest = 0u;
            return false;
        }
        if (skip > 0)
            rest &= rest - 1u;
        if (skip > 1)
            rest &= rest - 1u;
        if (skip > 2)
            rest &= rest - 1u;
        blk = __ffs(rest) - 1;
        rest &= rest - 1u;
        skip = TS_NG - 1;
        return true;
    };
    static_assert(TS_NG == 4, "`take` skips at most three units");

    if (threadIdx.x == 0) {
        for (int s = 0; s < TS_NG * TS_SAG; ++s) {
            mbar_init(bar_afull + 8 * s, 4);  // one arrival per warp of the gather group (after its warp-collective tcgen05.wait::st)
            mbar_init(bar_aempty + 8 * s, 1); // tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_dfull + 8 * b, TS_NG); // one tcgen05.commit per pipeline
            mbar_init(bar_dempty + 8 * b, 128);  // the four epilogue warps
        }
        mbar_init(bar_w, 1);
        fence_mbar_init();
    }
    if (threadIdx.x < 16) // 64 bytes of zeros: where the gather threads "load" a missing neighbour from
        reinterpret_cast<uint32_t *>(smem_gen + (smem_zero - smem_base))[threadIdx.x] = 0u;
    if (warp == TS_WARP_MMA0)
        tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp < 4) {
        // ================= epilogue: drain the partial accumulators of a finished tile while the next one accumulates =================
        const bool bf16 = is_bf16 != 0;
        const int quarter = warp;
        const bool fancy = epi.scale != nullptr || epi.shift != nullptr || epi.residual != nullptr || epi.relu != 0 || epi.stats != nullptr;
        float st_sum = 0.f, st_sq = 0.f; // lane l: column l & (COUT - 1) of this warp's rows (Epilogue::stats)
        int tb = 0, base = 0;
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            const uint32_t live = tile_live(tile);
            const int cnt = __popc(live);
            const int64_t row = int64_t(tile) * TS_TILE + quarter * 32 + lane;
            const bool row_ok = row < n_out;
            float v[COUT];
#pragma unroll
            for (int z = 0; z < COUT; ++z)
                v[z] = 0.f; // a tile no tap reaches was never accumulated: its rows are zero (+ bias)
            if (cnt) {
                const int b = tb & 1;
                mbar_wait_idle(bar_dfull + 8 * b, (tb >> 1) & 1, 256);
                tc_fence_after();
#pragma unroll
                for (int g = 0; g < TS_NG; ++g) { // fixed order: deterministic
                    if (((g - base) & (TS_NG - 1)) < cnt) { // pipeline g owned a unit of this tile (else its buffer holds stale data)
                        uint32_t acc[32];
                        const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t((g * 2 + b) * COUT);
                        if (COUT == 32)
                            tmem_ld_32x32b_x32(taddr, acc);
                        else
                            tmem_ld_32x32b_x16(taddr, acc);
                        tmem_ld_wait();
#pragma unroll
                        for (int z = 0; z < COUT; ++z)
                            v[z] += __uint_as_float(acc[z]);
                    }
                }
                tc_fence_before();
                mbar_arrive(bar_dempty + 8 * b); // the buffers may be overwritten by the tile after next
                ++tb;
                base = (base + cnt) & (TS_NG - 1);
            }
            if (epi.bias) {
#pragma unroll
                for (int z = 0; z < COUT; ++z)
                    v[z] += half_to_float(__ldg(reinterpret_cast<const uint16_t *>(epi.bias) + z), bf16);
            }
            if (fancy) { // stored = act2(act1((acc + bias) * scale + shift) + residual), statistics of the stored values
#pragma unroll
                for (int z = 0; z < COUT; ++z) {
                    if (epi.scale)
                        v[z] *= __ldg(epi.scale + z);
                    if (epi.shift)
                        v[z] += __ldg(epi.shift + z);
                    if (epi.relu & 1)
                        v[z] = fmaxf(v[z], 0.f);
                }
                if (epi.residual && row_ok) {
                    const uint4 *res = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint16_t *>(epi.residual) + row * COUT);
#pragma unroll
                    for (int v8 = 0; v8 < COUT / 8; ++v8) {
                        const uint4 r = __ldg(res + v8);
                        const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            float a, b;
                            unpack_half2(rw[h], bf16, a, b);
                            v[8 * v8 + 2 * h] += a, v[8 * v8 + 2 * h + 1] += b;
                        }
                    }
                }
                if (epi.relu & 2) {
#pragma unroll
                    for (int z = 0; z < COUT; ++z)
                        v[z] = fmaxf(v[z], 0.f);
                }
            }
            uint32_t p[COUT / 2];
#pragma unroll
            for (int h = 0; h < COUT / 2; ++h) {
                p[h] = pack_half2(v[2 * h], v[2 * h + 1], bf16);
                unpack_half2(p[h], bf16, v[2 * h], v[2 * h + 1]); // statistics see what a later pass over y would read
            }
            if (row_ok) {
                uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(y_) + row * COUT);
#pragma unroll
                for (int v4 = 0; v4 < COUT / 8; ++v4)
                    dst[v4] = make_uint4(p[4 * v4], p[4 * v4 + 1], p[4 * v4 + 2], p[4 * v4 + 3]);
            }
            if (epi.stats) { // per tile: the epilogue warps have time to spare, registers they have not
                float sq[COUT];
#pragma unroll
                for (int z = 0; z < COUT; ++z) {
                    v[z] = row_ok ? v[z] : 0.f;
                    sq[z] = v[z] * v[z];
                }
                st_sum += warp_column_sum<COUT>(v, lane);
                st_sq += warp_column_sum<COUT>(sq, lane);
            }
        }
        if (epi.stats) {
            float *s_stats = reinterpret_cast<float *>(smem_gen + (smem_stats - smem_base));
            if (lane < COUT) {
                s_stats[(quarter * 2 + 0) * COUT + lane] = st_sum;
                s_stats[(quarter * 2 + 1) * COUT + lane] = st_sq;
            }
        }
    } else if (warp < TS_WARP_MMA0) {
        // ================= gather: thread = one output row of the tile = one TMEM lane =================
        const int gw = warp - TS_WARP_G0, g = gw >> 2, q = gw & 3; // (warp & 3) == q: the lane quarter this warp may write
        const int r = q * 32 + lane;
        // chunk c of a tap's row sits at position c ^ swz: consecutive threads' 16-byte accesses then fall on distinct banks
        const uint32_t swz = CH == 2 ? uint32_t((r >> 2) & 1) : uint32_t((r >> 1) & 3);
        uint32_t coff[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c)
            coff[c] = (uint32_t(c) ^ swz) << 4;
        const uint32_t my_stage = smem_stage + g * (TS_DEPTH * TS_UNIT_BYTES) + r * ROWB;
        const uint32_t my_taddr = tmem_base + (uint32_t(q * 32) << 16) + TS_A_COL0 + uint32_t(g * TS_SAG * 32);
        const uint32_t my_afull = bar_afull + 8 * g * TS_SAG, my_aempty = bar_aempty + 8 * g * TS_SAG;
        const int32_t *nbr_r = nbr + r;
        // this pipeline's units, in the order its MMA warp consumes them
        int e_tile = tile_begin - 1, e_skip = g;
        uint32_t e_rest = 0u;
        auto next_own = [&](int &tile, int &blk) -> bool {
            while (true) {
                if (take(e_rest, e_skip, blk)) {
                    tile = e_tile;
                    return true;
                }
                if (e_tile + 1 >= tile_end)
                    return false;
                ++e_tile;
                e_rest = tile_live(e_tile);
                // the map is a pure HBM stream (4 K^3 bytes per row, read once): pull the NEXT tile's entries of this warp's 32 rows
                // into L2 now, one 128-byte line per tap -- taps g, g + NG, ... of this group, one per lane
                if (e_tile + 1 < tile_end && g + TS_NG * lane < k3)
                    prefetch_l2(nbr_r - lane + int64_t(g + TS_NG * lane) * pitch + int64_t(e_tile + 1) * TS_TILE);
            }
        };
        // Map entries travel through a per-warp shared-memory ring, two units ahead of the copies that need them: ONE 16-byte
        // cp.async per lane fetches the entries of the warp's 32 rows x G taps of a unit (lane = (chunk of four rows, tap));
        // a thread then reads its own G entries back (consecutive lanes, consecutive words).  Loading them into registers instead
        // left every step waiting on the scoreboard of the loads issued a few instructions earlier (profiles/r02_ncu_ts_v3_hot.txt).
        constexpr int MAP_SLOT = G * 128; // bytes per unit: G taps x 32 rows x 4 B
        const uint32_t my_ring = smem_map + gw * (TS_MAP_RING * MAP_SLOT);
        const int m_tap = lane & (G - 1), m_chunk = lane / G; // this lane's share of a unit's map copy (lanes < 8 G take part)
        const int32_t *nbr_w = nbr + q * 32 + m_chunk * 4;
        auto copy_map = [&](int tile, int blk, int j) {
            const int tap = blk * G + m_tap;
            if (lane < 8 * G && tap < k3)
                cp_async16(my_ring + (j & (TS_MAP_RING - 1)) * MAP_SLOT + m_tap * 128 + m_chunk * 16, nbr_w + int64_t(tap) * pitch + int64_t(tile) * TS_TILE, 16u);
        };
        auto read_idx = [&](int tile, int blk, int j, int (&idx)[G]) {
            const bool row_ok = tile < full_tiles || r < tail_rows;
            const uint32_t src = my_ring + (j & (TS_MAP_RING - 1)) * MAP_SLOT + lane * 4;
#pragma unroll
            for (int t = 0; t < G; ++t) {
                int v;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(src + t * 128) : "memory");
                idx[t] = (row_ok && blk * G + t < k3) ? v : -1;
            }
        };
        // a landed unit: hits from the staging slice, zeros for the misses (read from a zeroed 64-byte region: one select per
        // tap instead of eight register moves), one tcgen05.st of the 128-byte operand row
        auto complete = [&](int j, uint32_t valid, int slot) {
            const int sa = j % TS_SAG;
            mbar_wait(my_aempty + 8 * sa, ((j / TS_SAG) & 1) ^ 1u); // the MMAs that read this stage last have retired
            tc_fence_after();
            uint32_t v[32];
            const uint32_t src = my_stage + slot * TS_UNIT_BYTES;
#pragma unroll
            for (int t = 0; t < G; ++t) {
                const uint32_t base = ((valid >> t) & 1u) ? src + t * (TS_TILE * ROWB) : smem_zero;
#pragma unroll
                for (int c = 0; c < CH; ++c)
                    lds128(base + coff[c], v[(t * CH + c) * 4 + 0], v[(t * CH + c) * 4 + 1], v[(t * CH + c) * 4 + 2], v[(t * CH + c) * 4 + 3]);
            }
            tmem_st_32x32b_x32(my_taddr + uint32_t(sa * 32), v);
            tmem_st_wait(); // warp-collective: every lane's row is in tensor memory
            tc_fence_before();
            if (lane == 0)
                mbar_arrive(my_afull + 8 * sa);
        };
        // Software pipeline per thread: step k copies the map entries of unit k + 2 into the ring, issues the feature copies of
        // unit k (its map entries landed with group k - 2) and completes unit k - 1 (copies issued one step ago); one cp.async
        // group per step.
        struct Unit {
            int tile, blk;
            bool ok;
        };
        int own = 0; // index of the unit `cur` (units of this pipeline issued so far)
        auto fetch = [&](Unit &un, int j) {
            un.ok = next_own(un.tile, un.blk);
            if (un.ok)
                copy_map(un.tile, un.blk, j);
        };
        Unit cur, nx1, nx2;
        fetch(cur, 0);
        fetch(nx1, 1);
        cp_async_commit();
        cp_async_wait_all();
        __syncwarp();
        uint32_t prev_valid = 0u;
        bool has_prev = false;
        while (cur.ok || has_prev) {
            uint32_t valid = 0u;
            nx2.ok = false;
            if (cur.ok) {
                fetch(nx2, own + 2);
                int idx[G];
                read_idx(cur.tile, cur.blk, own, idx);
                const uint32_t dst = my_stage + (own & (TS_DEPTH - 1)) * TS_UNIT_BYTES;
#pragma unroll
                for (int t = 0; t < G; ++t) {
                    if (idx[t] >= 0) {
                        const uint16_t *src = x + int64_t(idx[t]) * CIN;
#pragma unroll
                        for (int c = 0; c < CH; ++c)
                            cp_async16_ca(dst + t * (TS_TILE * ROWB) + coff[c], src + c * 8);
                        valid |= 1u << t;
                    }
                }
            }
            cp_async_commit(); // (an empty group when nothing is left to issue: keeps the group count uniform)
            if (has_prev) {
                cp_async_wait<TS_DEPTH - 1>(); // every group but the newest has landed: the previous unit's copies are in ...
                __syncwarp();                  // ... and so are the map entries every lane of the warp copied a step earlier
                complete(own - 1, prev_valid, (own - 1) & (TS_DEPTH - 1));
            }
            prev_valid = valid, has_prev = cur.ok;
            if (cur.ok)
                ++own;
            cur = nx1, nx1 = nx2;
        }
        cp_async_wait_all();
    } else {
        // ================= MMA issuer of pipeline g (one thread): resident weight image, A from tensor memory =================
        const int g = warp - TS_WARP_MMA0;
        if (elect_one()) {
            if (g == 0) {
                const uint32_t w_bytes = uint32_t(total_blocks) * Cfg::WCHUNK;
                mbar_expect_tx(bar_w, w_bytes);
                for (uint32_t off = 0; off < w_bytes; off += 32768u)
                    bulk_g2s(smem_w + off, w_img + off, (w_bytes - off < 32768u ? w_bytes - off : 32768u), bar_w);
            }
            mbar_wait(bar_w, 0);
            const uint64_t desc_hi = make_smem_desc_sw128(0, 16, 1024) & 0xFFFFFFFF00000000ull;
            const uint32_t my_afull = bar_afull + 8 * g * TS_SAG, my_aempty = bar_aempty + 8 * g * TS_SAG;
            const uint32_t a0 = tmem_base + TS_A_COL0 + uint32_t(g * TS_SAG * 32);
            const uint32_t w_lo = ((smem_w & 0x3FFFFu) >> 4) | (1u << 16);
            int j = 0, tb = 0, skip = g;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                uint32_t rest = tile_live(tile);
                if (rest == 0u)
                    continue;
                const int db = tb & 1;
                // every pipeline takes part in every live tile's hand-shake, with or without a unit of its own in it: the phase
                // of bar_dfull counts NG commits per tile
                mbar_wait_idle(bar_dempty + 8 * db, ((tb >> 1) & 1) ^ 1u, 64);
                tc_fence_after();
                const uint32_t d = tmem_base + uint32_t((g * 2 + db) * COUT);
                uint32_t accumulate = 0u;
                int blk;
                while (take(rest, skip, blk)) {
                    const int sa = j % TS_SAG;
                    mbar_wait_idle(my_afull + 8 * sa, (j / TS_SAG) & 1, 32);
                    tc_fence_after();
                    const uint32_t a = a0 + uint32_t(sa * 32);
                    const uint32_t b_lo = w_lo + uint32_t(blk) * (Cfg::WCHUNK >> 4);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) { // K = 16 per MMA: 8 columns of A, 32 bytes of the swizzled weight rows
                        umma_f16_ts(d, a + 8 * kk, desc_hi | (b_lo + 2 * kk), idesc, accumulate);
                        accumulate = 1u;
                    }
                    umma_commit(my_aempty + 8 * sa);
                    ++j;
                }
                umma_commit(bar_dfull + 8 * db);
                ++tb;
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (epi.stats) { // fixed-order sum of the four lane quarters -> this CTA's partial [2][COUT]
        const float *s_stats = reinterpret_cast<const float *>(smem_gen + (smem_stats - smem_base));
        for (int e = threadIdx.x; e < 2 * COUT; e += TS_THREADS)
            epi.stats[int64_t(blockIdx.x) * 2 * COUT + e] = (s_stats[e] + s_stats[2 * COUT + e]) + (s_stats[4 * COUT + e] + s_stats[6 * COUT + e]);
    }
    if (warp == TS_WARP_MMA0)
        tmem_dealloc(tmem_base, 512);
}

// ---- host side ------------------------------------------------------------------------------------
bool tc_ts_supported(int32_t cin, int32_t cout, int64_t k3, int32_t dtype) {
    if (dtype != FVC_F16 && dtype != FVC_BF16)
        return false;
    if (!((cin == 16 || cin == 32) && (cout == 16 || cout == 32)) || k3 < 1 || k3 > 128)
        return false;
    const int64_t blocks = ceil_div(k3, 64 / cin);
    return blocks <= TS_MAX_BLOCKS && blocks * cout * 128 <= 65536;
}

static inline void ts_chunking(int64_t n_out, int *grid, int *tiles_per_chunk) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = ceil_div(n_out > 0 ? n_out : 1, TS_TILE);
    const int64_t chunks = tiles < sms ? tiles : sms;
    *tiles_per_chunk = int(ceil_div(tiles, chunks));
    *grid = int(ceil_div(tiles, *tiles_per_chunk));
}

int64_t tc_ts_stats_blocks(int64_t n_out, int32_t *rows_per_block) {
    int grid = 0, tpc = 0;
    ts_chunking(n_out, &grid, &tpc);
    if (rows_per_block)
        *rows_per_block = tpc * TS_TILE;
    return n_out > 0 ? grid : 0;
}

template <int CIN, int COUT> static int launch_ts(const ConvArgs &a, const void *x, const uint8_t *img) {
    using Cfg = TsCfg<CIN, COUT>;
    auto kernel = conv_tc_ts_kernel<CIN, COUT>;
    const int blocks = int(ceil_div(a.k3, Cfg::G));
    const size_t smem = Cfg::smem_bytes(blocks);
    static std::atomic<unsigned long long> configured{0};
    const int rc = ensure_dynamic_smem(kernel, 232448 - 1024, configured); // the resident weight image makes the size depend on K^3
    if (rc)
        return rc;
    int grid = 0, tpc = 0;
    ts_chunking(a.n_out, &grid, &tpc);
    const bool bf16 = a.dtype == FVC_BF16;
    kernel<<<grid, TS_THREADS, smem, a.stream>>>(reinterpret_cast<const uint16_t *>(x), img, a.epi, a.y, a.nbr, a.pitch,
                                                 reinterpret_cast<const unsigned long long *>(a.tile_mask), a.n_out, a.k3, tpc,
                                                 make_idesc_f16(128, COUT, bf16, false, false), bf16 ? 1 : 0);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int tc_ts_forward(const ConvArgs &a, const void *x, const uint8_t *img) {
    if (a.cin == 16 && a.cout == 16)
        return launch_ts<16, 16>(a, x, img);
    if (a.cin == 16 && a.cout == 32)
        return launch_ts<16, 32>(a, x, img);
    if (a.cin == 32 && a.cout == 16)
        return launch_ts<32, 16>(a, x, img);
    if (a.cin == 32 && a.cout == 32)
        return launch_ts<32, 32>(a, x, img);
    return set_error(FVC_ERR_UNSUPPORTED, "no tensor-memory executor for channels %d -> %d", a.cin, a.cout);
}

} // namespace fvc
