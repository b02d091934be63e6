// conv_tc_wgrad.cu -- tcgen05 / TMEM weight gradient for f16 / bf16 and (three-way bf16 split) fp32: fp32 accumulate,
// fixed-order reduction.
//
//   dW[k][ci][co] = sum over output rows o of  X[nbr[k][o]][ci] * dY[o][co]        (GatherScatterDefault.cu:806-807)
//
// Output-stationary over the dense tap-major map, like the forward kernel, but with the roles of the GEMM
// dimensions rotated: the REDUCTION runs over output rows (K = 128 rows per tile, 8 x tcgen05.mma of K=16),
// M stacks (tap, input channel) and N is the output channel.  Both operands are therefore MN-major:
//   A = gathered X rows   [128 rows][64 ch]  -- exactly the 128B-swizzled block the forward producers write,
//   B = the dY tile       [128 rows][64 ch]  -- loaded once per tile and shared by every tap.
// An "M-unit" is two such A blocks (M = 128 = two taps x 64 channels when Cin = 64, or two channel blocks of
// one tap when Cin = 128 ...), its accumulator is 128 TMEM lanes x Cout columns and stays resident while
// the CTA streams over its share of the row tiles.  TMEM (512 columns) holds 512 / Cout units, so the
// taps are split into groups (grid.y) and the rows into chunks (grid.x); every CTA writes one fp32 partial
// [K^3][Cin][Cout] slice that wgrad_reduce_partials sums in a fixed order (deterministic, no atomics).
//
// fp32 (SPLIT = true): X and dY arrive as bf16 split rows [N][3][C] (conv_tc.cu: tc_split_rows); per (unit, X split i) the
// issuer runs the MMAs against dY splits 0 .. 2 - i, x0.d0 into the unit's main accumulator, the five small terms into a
// second one.  tcgen05 truncates on every accumulate, so the full-magnitude chain is kept short: every `seg_tiles` row
// tiles the CTA drains its accumulators and adds them (rounded fp32 adds, same CTA, fixed order) into its partial slice.
#include "conv_internal.cuh"
#include "tc_ptx.cuh"

namespace fvc {

using namespace tc;

int g_wgrad_variant = 0;

constexpr int WG_PW = 4;                  // gather-producer warps (also the epilogue warps)
constexpr int WG_WARP_MMA = WG_PW;        // warp WG_PW + 1 streams the kernel map
constexpr int WG_THREADS = (WG_PW + 2) * 32;
constexpr int WG_THREADS_PIPE = (WG_PW + 3) * 32; // MODE 1: producers | two MMA warps | dY loader
constexpr int wg_threads(int mode) { return mode == 1 ? WG_THREADS_PIPE : WG_THREADS; }
constexpr int WG_TILE = 128;
constexpr int WG_BLOCK_BYTES = WG_TILE * 128; // 128 rows x 64 reduction-side elements x 2 B

// Small channel counts are packed like in the forward kernel: an A block holds G = 64 / CIN taps x CIN channels;
// a dY row narrower than 64 channels is zero-padded to one 128-byte row (N = 64 for the MMA, extra columns unused).
// CTAS: co-resident CTAs per SM (TMEM columns are split between them).  Two half-size CTAs hide each other's per-unit
// hand-off latency on the narrow-output shapes, where a dY tile is cheap to load twice.
template <int CIN, int COUT, int STAGES, bool SPLIT = false, int CTAS = 1, int BSTG = 0> struct TcWgradCfg {
    static constexpr int G = CIN >= 64 ? 1 : 64 / CIN;      // taps per A block
    static constexpr int CB = CIN >= 64 ? CIN / 64 : 1;     // A channel blocks per tap (group)
    static constexpr int CPT = CIN >= 64 ? 8 : CIN / 8;     // 16-byte chunks one tap contributes to a row
    static constexpr int NB = COUT >= 64 ? COUT / 64 : 1;   // B blocks
    static constexpr int NPAD = COUT >= 64 ? COUT : 64;     // MMA N = TMEM columns per unit
    static constexpr int BQ = COUT >= 64 ? 8 : COUT / 8;    // valid 16-byte chunks of a dY row per B block
    static constexpr int NS = SPLIT ? 3 : 1;                // bf16 splits per fp32 operand
    static constexpr int XS = NS * CIN, YS = NS * COUT;     // row strides (elements) of the (split) feature / grad rows
    static constexpr int ACC_COLS = (SPLIT ? 2 : 1) * NPAD; // TMEM columns per unit (SPLIT: main | small-term accumulator)
    static constexpr int TMEM_COLS = 512 / CTAS;            // this CTA's share of the SM's tensor memory
    static constexpr int MAX_UNITS = TMEM_COLS / ACC_COLS;  // accumulators that fit
    // dY tile buffers.  Two let the next tile's dY load overlap this tile's MMAs (one buffer stalls every tile on the load:
    // 64 -> 64 on C2 0.615 -> 0.549 ms, profiles/r02_variants_c2_wgrad.json); the packed 16-channel shape has no room for it
    static constexpr int BSTAGES = BSTG ? BSTG : (((SPLIT && COUT >= 128) || (CTAS > 1 && CIN < 32)) ? 1 : 2);
    static constexpr int RING = G >= 2 ? 4 : 8;             // kernel-map ring depth
    static constexpr int SUB_STRIDE = G > 1 ? 528 : 512;    // 128 int32 per tap (+ a 16-byte pad so packed taps sit on different banks)
    static constexpr int RING_BYTES = 2 * G * SUB_STRIDE;   // two blocks x G taps
    static constexpr int A_STAGE = 2 * WG_BLOCK_BYTES;
    static constexpr int B_STAGE = NS * NB * WG_BLOCK_BYTES; // [split][64-channel block] of one dY tile
    static constexpr int NUM_BARS = 2 * STAGES + 5 + 2 * RING;
    static constexpr size_t SMEM = 1024 + size_t(STAGES) * A_STAGE + size_t(BSTAGES) * B_STAGE + size_t(RING) * RING_BYTES + 8 * NUM_BARS + 16;
    static_assert(CIN == 16 || CIN == 32 || CIN == 64 || CIN == 128 || CIN == 256, "unsupported Cin");
    static_assert(COUT == 16 || COUT == 32 || COUT == 64 || COUT == 128 || COUT == 256, "unsupported Cout");
    static_assert(!SPLIT || COUT <= 128, "fp32 split: Cout <= 128");
    static_assert((SMEM + 1024 + 768) * CTAS <= 233472, "shared memory budget exceeded");
    static_assert(CTAS == 1 || CTAS == 2, "one or two CTAs per SM");
};

// One warp's share (rows [32w, 32w+32)) of a 128-row x 128-byte swizzled block: 8 lanes q cover one row (one full
// 128-byte line); lane group g = lane >> 3 owns the 8 consecutive rows row0 = 32w + 8g .. row0 + 7 (row0 is a multiple
// of 8, so row i's swizzle term is i); idx[i] < 0 zero-fills row row0 + i.
__device__ __forceinline__ void gather_rows(uint32_t dst /* block + row0 * 128 */, int q,
                                            const uint16_t *__restrict__ base /* + column offset */, int64_t row_stride,
                                            const int (&idx)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
        cp_async16(dst + i * 128 + (uint32_t(q ^ i) << 4), idx[i] >= 0 ? base + int64_t(idx[i]) * row_stride : base,
                   idx[i] >= 0 ? 16u : 0u);
}
__device__ __forceinline__ void lds_v4x2(uint32_t addr, int (&v)[8]) {
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr) : "memory");
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr + 16) : "memory");
}

// MODE 0: the four producer warps share every unit (a quarter of its rows each), a streamer warp feeds them map entries through
//         a shared-memory ring (round 1; still serves fp32, whose per-segment drain keeps the producers in step with the issuer).
// MODE 1: two pipelines per CTA (half precision).  Pipeline p owns the units (accumulators) ul = p (mod 2), STAGES / 2 gather
//         stages, two producer warps -- warp (p, h) gathers WHOLE 16 KB blocks, block h of each of the pipeline's units: 32 x
//         16-byte cp.async per lane, map entries by eight 16-byte loads one block ahead, no ring -- and its OWN MMA-issuing warp;
//         a seventh warp loads the dY tiles both pipelines read.  One issuing thread per SM needs ~850-1200 dependent cycles per
//         unit (uniform-register moves, mbarrier round trips: profiles/r02_ts_executor.md) and was what the wide shapes, which
//         run one CTA per SM, waited for.
template <int CIN, int COUT, int STAGES, bool SPLIT, int CTAS, int BSTG, int MODE>
__global__ void __launch_bounds__(wg_threads(MODE), CTAS)
conv_tc_wgrad_kernel(const uint16_t *__restrict__ x, const uint16_t *__restrict__ dy, const int32_t *__restrict__ nbr,
                     int64_t pitch, const unsigned long long *__restrict__ tile_mask, int64_t n_out, int k3, int units_per_group,
                     int tiles_per_chunk, int mini, int seg_tiles, uint32_t idesc, float *__restrict__ partial) {
    using Cfg = TcWgradCfg<CIN, COUT, STAGES, SPLIT, CTAS, BSTG>;
    constexpr int G = Cfg::G, CB = Cfg::CB, CPT = Cfg::CPT, NB = Cfg::NB, NPAD = Cfg::NPAD, RING = Cfg::RING;
    constexpr int NS = Cfg::NS, XS = Cfg::XS, YS = Cfg::YS, ACC = Cfg::ACC_COLS, BSTAGES = Cfg::BSTAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_a + STAGES * Cfg::A_STAGE;
    const uint32_t smem_idx = smem_b + BSTAGES * Cfg::B_STAGE;
    const uint32_t bars = smem_idx + RING * Cfg::RING_BYTES;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * STAGES;
    const uint32_t bar_bfull = bars + 16 * STAGES, bar_bempty = bar_bfull + 16;
    const uint32_t bar_accum = bar_bempty + 16;
    const uint32_t bar_ifull = bar_accum + 8, bar_iempty = bar_ifull + 8 * RING;
    const uint32_t tmem_slot = bar_iempty + 8 * RING;
    __shared__ uint32_t s_started[2]; // per segment parity: units whose accumulator was written (MMA thread -> drain)
    volatile uint32_t *tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t *>(smem_raw + (smem_base - smem_u32(smem_raw)) + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // A block b: Cin >= 64 -> (tap b / CB, channels 64 * (b % CB) ..); Cin < 64 -> taps [G b, G b + G) x all channels
    const int total_blocks = CIN >= 64 ? k3 * CB : (k3 + G - 1) / G;
    const int total_units = (total_blocks + 1) / 2;
    const int unit0 = blockIdx.y * units_per_group;
    const int nunits = total_units - unit0 < units_per_group ? total_units - unit0 : units_per_group;
    const int64_t total_tiles = (n_out + WG_TILE - 1) / WG_TILE;
    // This CTA's tiles, q = 0 .. n_my - 1.  mini > 0: mini-chunks of `mini` consecutive tiles dealt round-robin over the gridDim.x
    // row chunks, so that ALL resident CTAs -- every row chunk and every tap group -- sweep the tile space together and what
    // they touch at one time (dY tiles re-read by each tap group, feature rows re-gathered by neighbouring taps) stays in L2;
    // with one contiguous range per CTA (mini == 0, round 1) the tap groups of a chunk drift apart and the ranges of the
    // chunks are far apart to begin with: the 128 -> 128 weight gradient read 3.6 GB from DRAM for 0.9 GB of operands.
    const int64_t tile_begin = int64_t(blockIdx.x) * tiles_per_chunk;
    const int64_t tile_end = tile_begin + tiles_per_chunk < total_tiles ? tile_begin + tiles_per_chunk : total_tiles;
    int64_t n_my = tile_end - tile_begin;
    if (mini > 0) {
        const int64_t mcs = (total_tiles + mini - 1) / mini, x = blockIdx.x, X = gridDim.x;
        const int64_t mine = mcs > x ? (mcs - 1 - x) / X + 1 : 0;
        n_my = mine * mini - ((mine > 0 && (mcs - 1) % X == x) ? mcs * mini - total_tiles : 0); // (the last mini-chunk may be short)
    }
    auto tile_of = [&](int64_t q) -> int64_t { return mini > 0 ? ((q / mini) * gridDim.x + blockIdx.x) * mini + q % mini : tile_begin + q; };
    auto first_tap = [&](int blk) -> int { return CIN >= 64 ? blk / CB : blk * G; };

    // (tile, unit) skipping: a unit is live for a tile iff one of its taps reaches a row of the tile.  Every role derives
    // the same live-unit bitmask from the tile's tap bitmask (K^3 <= 128; larger kernels do not skip).
    const int words = (k3 + 63) >> 6;
    const bool use_mask = tile_mask != nullptr && words <= 2;
    __shared__ unsigned long long s_unit_taps[8][2]; // taps (bits of the two mask words) that belong to each unit
    if (threadIdx.x < 8) {
        unsigned long long lo = 0ull, hi = 0ull;
        const int ul = threadIdx.x;
        if (ul < nunits) {
            const int blk = 2 * (unit0 + ul);
            const int t_lo = first_tap(blk);
            int t_hi = first_tap(blk + 1 < total_blocks ? blk + 1 : blk) + G; // exclusive
            t_hi = t_hi < k3 ? t_hi : k3;
            for (int tap = t_lo; tap < t_hi && tap < 128; ++tap)
                (tap < 64 ? lo : hi) |= 1ull << (tap & 63);
        }
        s_unit_taps[ul][0] = lo;
        s_unit_taps[ul][1] = hi;
    }
    auto live_units = [&](int64_t tile) -> uint32_t {
        if (!use_mask)
            return (1u << nunits) - 1u;
        const unsigned long long m0 = __ldg(tile_mask + tile * words), m1 = words > 1 ? __ldg(tile_mask + tile * words + 1) : 0ull;
        uint32_t live = 0;
#pragma unroll
        for (int ul = 0; ul < 8; ++ul)
            live |= uint32_t(((m0 & s_unit_taps[ul][0]) | (m1 & s_unit_taps[ul][1])) != 0ull) << ul;
        return live;
    };

    if (threadIdx.x == 0) {
        s_started[0] = s_started[1] = 0u;
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, MODE == 1 ? 64 : WG_PW * 32); // MODE 1: the two warps that gather a unit's two blocks
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_bfull + 8 * b, MODE == 1 ? 32 : WG_PW * 32); // MODE 1: the dY loader warp
            mbar_init(bar_bempty + 8 * b, MODE == 1 ? 2 : 1); // MODE 1: both pipelines release a dY tile
        }
        mbar_init(bar_accum, MODE == 1 ? 2 : 1);
        for (int e = 0; e < RING; ++e) {
            mbar_init(bar_ifull + 8 * e, 32);
            mbar_init(bar_iempty + 8 * e, WG_PW);
        }
        fence_mbar_init();
    }
    if (warp == WG_WARP_MMA)
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp < WG_PW) {
        // ================= producers: dY tile, then the gathered X blocks of every live unit; drain per segment =================
        const int q = lane & 7, sub = q / CPT;
        const int row0 = warp * 32 + (lane >> 3) * 8; // rows row0 + i, i < 8
        const uint32_t dst0 = uint32_t(row0) * 128u;
        const uint16_t *xq = x + (CIN >= 64 ? q * 8 : (q % CPT) * 8), *dyq = dy + q * 8;
        // drain role of this warp: TMEM lanes 32w .. 32w+31 = reduction-side element kk of A block `half` of every unit
        const int half = warp >> 1;
        const int kk = (warp & 1) * 32 + lane;
        float *slice = partial + int64_t(blockIdx.x) * k3 * CIN * COUT;
        // ---- drain: accumulators -> fp32 partial slice [k][ci][co] (first segment stores, later ones add).  The issuer
        //      cannot touch TMEM again before every producer thread has arrived on the next stage, i.e. after this.
        auto drain = [&](int seg) {
            mbar_wait(bar_accum, uint32_t(seg & 1));
            tc_fence_after();
            const uint32_t started = *reinterpret_cast<volatile uint32_t *>(&s_started[seg & 1]);
            constexpr int EC = COUT >= 32 ? 32 : 16; // columns drained per tcgen05.ld
            for (int ul = 0; ul < nunits; ++ul) {
                const int blk = 2 * (unit0 + ul) + half;
                const int tap = CIN >= 64 ? blk / CB : blk * G + kk / CIN;
                const int ci = CIN >= 64 ? (blk % CB) * 64 + kk : kk % CIN;
                const bool live_row = blk < total_blocks && tap < k3;
                const bool touched = (started >> ul) & 1u;
                if (!touched && seg > 0)
                    continue; // nothing to add
#pragma unroll
                for (int c0 = 0; c0 < COUT; c0 += EC) {
                    uint32_t acc[32];
                    if (touched) {
                        const uint32_t taddr = tmem_base + (uint32_t(warp * 32) << 16) + uint32_t(ul * ACC + c0);
                        if (EC == 32)
                            tmem_ld_32x32b_x32(taddr, acc);
                        else
                            tmem_ld_32x32b_x16(taddr, acc);
                        if (SPLIT) {
                            uint32_t small[32];
                            if (EC == 32)
                                tmem_ld_32x32b_x32(taddr + NPAD, small);
                            else
                                tmem_ld_32x32b_x16(taddr + NPAD, small);
                            tmem_ld_wait();
#pragma unroll
                            for (int z = 0; z < EC; ++z)
                                acc[z] = __float_as_uint(__uint_as_float(acc[z]) + __uint_as_float(small[z]));
                        } else {
                            tmem_ld_wait();
                        }
                    } else { // no row of this CTA's tiles ever reached these taps
#pragma unroll
                        for (int z = 0; z < 32; ++z)
                            acc[z] = 0u;
                    }
                    if (live_row) {
                        uint4 *dst = reinterpret_cast<uint4 *>(slice + (int64_t(tap) * CIN + ci) * COUT + c0);
#pragma unroll
                        for (int v = 0; v < EC / 4; ++v) {
                            uint4 o = make_uint4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
                            if (seg > 0) {
                                const uint4 old = dst[v];
                                o.x = __float_as_uint(__uint_as_float(old.x) + __uint_as_float(o.x));
                                o.y = __float_as_uint(__uint_as_float(old.y) + __uint_as_float(o.y));
                                o.z = __float_as_uint(__uint_as_float(old.z) + __uint_as_float(o.z));
                                o.w = __float_as_uint(__uint_as_float(old.w) + __uint_as_float(o.w));
                            }
                            dst[v] = o;
                        }
                    }
                }
            }
        };
        if constexpr (MODE == 1) {
            // ================= MODE 1 producers: warp (p, h) gathers block h of every unit of pipeline p =================
            // lane = (lg, q): 8 lanes q cover one 128-byte row; lane group lg owns the 32 consecutive rows 32 lg .. 32 lg + 31 and
            // copies 16-byte chunk q of each.  Its 32 map entries (tap `sub` of the block) are eight 16-byte loads, one block ahead.
            static_assert(!SPLIT && STAGES % 2 == 0, "pipelines: half precision, an even number of gather stages");
            constexpr int SPW = STAGES / 2; // stages per pipeline
            const int p = warp >> 1, h = warp & 1;
            const uint32_t pmask = p ? 0xAAAAAAAAu : 0x55555555u; // units ul = p (mod 2)
            const int lg = lane >> 3;
            const uint32_t lane_off = (uint32_t(lg) << 12) | (uint32_t(q) << 4);
            struct Item {
                int64_t tile;
                int blk, stage, use;
            };
            int64_t e_q = -1, e_tile = 0;
            uint32_t e_rest = 0u;
            int e_j = 0; // units of this pipeline so far
            auto next_item = [&](Item &out) -> bool { // (tile ascending, own live unit ascending): the order the pipeline's MMA warp walks
                while (true) {
                    if (e_rest == 0u) {
                        if (e_q + 1 >= n_my)
                            return false;
                        e_tile = tile_of(++e_q);
                        e_rest = live_units(e_tile) & pmask;
                        continue;
                    }
                    out.tile = e_tile;
                    out.blk = 2 * (unit0 + __ffs(e_rest) - 1) + h;
                    out.stage = p * SPW + e_j % SPW;
                    out.use = e_j / SPW;
                    e_rest &= e_rest - 1u;
                    ++e_j;
                    return true;
                }
            };
            auto load_idx = [&](const Item &item, int4 (&out)[8]) {
                const int tap = first_tap(item.blk) + sub;
                if (item.blk < total_blocks && tap < k3) { // (an odd block count leaves a zero dummy block in the last unit)
                    const int4 *src = reinterpret_cast<const int4 *>(nbr + int64_t(tap) * pitch + item.tile * WG_TILE + lg * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        out[j] = __ldg(src + j);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        out[j] = make_int4(-1, -1, -1, -1);
                }
            };
            Item cur, fut;
            int4 idx[8], nxt[8];
            bool have = next_item(cur);
            if (have)
                load_idx(cur, idx);
            while (have) {
                const bool have_next = next_item(fut);
                if (have_next)
                    load_idx(fut, nxt);
                const int64_t rows_left = n_out - cur.tile * WG_TILE - lg * 32; // row 32 lg + i exists iff i < rows_left
                const uint16_t *src0 = xq + (CIN >= 64 ? (cur.blk % CB) * 64 : 0);
                mbar_wait(bar_empty + 8 * cur.stage, (cur.use & 1) ^ 1u);
                const uint32_t dst = smem_a + cur.stage * Cfg::A_STAGE + (cur.blk & 1) * WG_BLOCK_BYTES;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int v[4] = {idx[j].x, idx[j].y, idx[j].z, idx[j].w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int i = 4 * j + c;
                        const bool ok = v[c] >= 0 && i < rows_left;
                        cp_async16(dst + (lane_off ^ uint32_t(i * 128 + ((i & 7) << 4))), ok ? src0 + int64_t(v[c]) * XS : x, ok ? 16u : 0u);
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
            drain(0); // half precision: one segment, drained once every tile of the chunk has been accumulated
            tc_fence_before();
        } else {
        int s = 0, e = 0, tb = 0, seg = 0, seg_t = 0;
        uint32_t ph = 0, eph = 0;
        for (int64_t qi = 0; qi < n_my; ++qi, ++tb) {
            const int64_t tile = tile_of(qi);
            const int64_t rows_left = n_out - tile * WG_TILE - row0; // row row0 + i exists iff i < rows_left
            const uint32_t live = live_units(tile);
            { // B: plain rows of dY (identity "map"), every split; lanes beyond a narrow row's chunks zero-fill
                const int bs = tb % BSTAGES;
                mbar_wait(bar_bempty + 8 * bs, ((tb / BSTAGES) & 1) ^ 1);
                int self[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    self[i] = (i < rows_left && q < Cfg::BQ) ? int(tile * WG_TILE + row0 + i) : -1;
#pragma unroll
                for (int j = 0; j < NS; ++j)
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb)
                        gather_rows(smem_b + bs * Cfg::B_STAGE + (j * NB + nb) * WG_BLOCK_BYTES + dst0, q, dyq + j * COUT + nb * 64, YS, self);
                cp_async_arrive_noinc(bar_bfull + 8 * bs);
            }
            for (uint32_t rest = live; rest; rest &= rest - 1u) {
                const int blk = 2 * (unit0 + __ffs(rest) - 1);
                const bool ok0 = first_tap(blk) + sub < k3;
                const bool ok1 = blk + 1 < total_blocks && first_tap(blk + 1) + sub < k3; // odd block count: zero dummy
                for (int i = 0; i < NS; ++i) {
                    mbar_wait(bar_ifull + 8 * e, eph);
                    int idx0[8], idx1[8];
                    const uint32_t entry = smem_idx + e * Cfg::RING_BYTES + sub * Cfg::SUB_STRIDE + row0 * 4;
                    lds_v4x2(entry, idx0);
                    lds_v4x2(entry + G * Cfg::SUB_STRIDE, idx1);
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        if (r >= rows_left || !ok0)
                            idx0[r] = -1;
                        if (r >= rows_left || !ok1)
                            idx1[r] = -1;
                    }
                    mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                    const uint32_t stage = smem_a + s * Cfg::A_STAGE + dst0;
                    gather_rows(stage, q, xq + i * CIN + (CIN >= 64 ? (blk % CB) * 64 : 0), XS, idx0);
                    gather_rows(stage + WG_BLOCK_BYTES, q, xq + i * CIN + (CIN >= 64 ? ((blk + 1) % CB) * 64 : 0), XS, idx1);
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(bar_iempty + 8 * e);
                    cp_async_arrive_noinc(bar_full + 8 * s);
                    if (++s == STAGES) {
                        s = 0;
                        ph ^= 1u;
                    }
                    if (++e == RING) {
                        e = 0;
                        eph ^= 1u;
                    }
                }
            }
            if (++seg_t < seg_tiles && qi + 1 < n_my)
                continue;
            drain(seg);
            tc_fence_before();
            ++seg;
            seg_t = 0;
        }
        cp_async_wait_all();
        } // MODE 0
    } else if (MODE == 1 && (warp == WG_WARP_MMA || warp == WG_WARP_MMA + 1)) {
        // ================= MODE 1: MMA issuer of pipeline p (its own units, its own stages; the dY tile is shared) =================
        if (elect_one()) {
            constexpr int SPW = STAGES / 2;
            const int p = warp - WG_WARP_MMA;
            const uint32_t pmask = p ? 0xAAAAAAAAu : 0x55555555u;
            const uint64_t desc_hi = make_smem_desc_sw128(0, WG_BLOCK_BYTES, 1024) & 0xFFFFFFFF00000000ull;
            const uint32_t lbo = uint32_t(WG_BLOCK_BYTES >> 4) << 16;
            const uint32_t a_lo0 = ((smem_a & 0x3FFFFu) >> 4) | lbo, b_lo0 = ((smem_b & 0x3FFFFu) >> 4) | lbo;
            int j = 0, tb = 0;
            uint32_t started = 0;
            for (int64_t qi = 0; qi < n_my; ++qi, ++tb) {
            const int64_t tile = tile_of(qi);
                const int bs = tb % BSTAGES;
                const uint32_t live = live_units(tile) & pmask;
                mbar_wait(bar_bfull + 8 * bs, (tb / BSTAGES) & 1); // both pipelines take part in every tile's dY hand-shake
                fence_proxy_async();
                const uint32_t b_lo = b_lo0 + uint32_t(bs) * (Cfg::B_STAGE >> 4);
                for (uint32_t rest = live; rest; rest &= rest - 1u, ++j) {
                    const int ul = __ffs(rest) - 1;
                    const int s = p * SPW + j % SPW;
                    mbar_wait(bar_full + 8 * s, (j / SPW) & 1);
                    fence_proxy_async();
                    tc_fence_after();
                    const uint32_t a_lo = a_lo0 + uint32_t(s) * (Cfg::A_STAGE >> 4);
                    const uint32_t acc0 = (started >> ul) & 1u;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        umma_f16(tmem_base + uint32_t(ul * ACC), desc_hi | (a_lo + 128 * kk), desc_hi | (b_lo + 128 * kk), idesc, acc0 | uint32_t(kk != 0));
                    started |= 1u << ul;
                    umma_commit(bar_empty + 8 * s);
                }
                umma_commit(bar_bempty + 8 * bs);
            }
            atomicOr(&s_started[0], started); // the two pipelines' accumulators are disjoint: the drain reads the union
            __threadfence_block();
            umma_commit(bar_accum);
        }
        __syncwarp();
    } else if (MODE != 1 && warp == WG_WARP_MMA) {
        // ================= MMA issuer =================
        if (elect_one()) {
            // MN-major SWIZZLE_128B descriptors: LBO = distance between the two 64-wide blocks, SBO = 8-row group
            const uint64_t desc_hi = make_smem_desc_sw128(0, WG_BLOCK_BYTES, 1024) & 0xFFFFFFFF00000000ull;
            const uint32_t lbo = uint32_t(WG_BLOCK_BYTES >> 4) << 16;
            const uint32_t a_lo0 = ((smem_a & 0x3FFFFu) >> 4) | lbo, b_lo0 = ((smem_b & 0x3FFFFu) >> 4) | lbo;
            int s = 0, tb = 0, seg = 0, seg_t = 0;
            uint32_t ph = 0, started = 0, started_small = 0;
            for (int64_t qi = 0; qi < n_my; ++qi, ++tb) {
            const int64_t tile = tile_of(qi);
                const int bs = tb % BSTAGES;
                const uint32_t live = live_units(tile);
                mbar_wait(bar_bfull + 8 * bs, (tb / BSTAGES) & 1);
                fence_proxy_async(); // cp.async (generic proxy) wrote the dY tile; tcgen05.mma reads it through the async proxy
                const uint32_t b_lo = b_lo0 + uint32_t(bs) * (Cfg::B_STAGE >> 4);
                for (uint32_t rest = live; rest; rest &= rest - 1u) {
                    const int ul = __ffs(rest) - 1;
                    for (int i = 0; i < NS; ++i) {
                        mbar_wait(bar_full + 8 * s, ph);
                        fence_proxy_async();
                        tc_fence_after();
                        const uint32_t a_lo = a_lo0 + uint32_t(s) * (Cfg::A_STAGE >> 4);
                        for (int jj = 0; jj + i <= (SPLIT ? 2 : 0); ++jj) { // X split i meets dY splits 0 .. 2 - i
                            const bool main_term = (i | jj) == 0;
                            const uint32_t acc0 = ((main_term ? started : started_small) >> ul) & 1u;
                            const uint32_t d = tmem_base + uint32_t(ul * ACC + (main_term ? 0 : NPAD));
                            const uint32_t bj = b_lo + uint32_t(jj) * ((NB * WG_BLOCK_BYTES) >> 4);
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk) // 16 rows (K) per MMA = two 8-row swizzle groups = 2048 B = 128 units
                                umma_f16(d, desc_hi | (a_lo + 128 * kk), desc_hi | (bj + 128 * kk), idesc, acc0 | uint32_t(kk != 0));
                            if (main_term)
                                started |= 1u << ul;
                            else
                                started_small |= 1u << ul;
                        }
                        umma_commit(bar_empty + 8 * s);
                        if (++s == STAGES) {
                            s = 0;
                            ph ^= 1u;
                        }
                    }
                }
                umma_commit(bar_bempty + 8 * bs);
                if (++seg_t == seg_tiles || qi + 1 == n_my) { // segment done: hand the accumulators to the drain
                    *reinterpret_cast<volatile uint32_t *>(&s_started[seg & 1]) = started;
                    __threadfence_block();
                    umma_commit(bar_accum);
                    started = started_small = 0;
                    ++seg;
                    seg_t = 0;
                }
            }
        }
        __syncwarp();
    } else if constexpr (MODE == 1) {
        // ================= dY tile loader (whole warp, the seventh): 128 plain rows per tile, every 64-channel block =================
        const int q = lane & 7, lg = lane >> 3;
        const uint32_t lane_off = (uint32_t(lg) << 12) | (uint32_t(q) << 4);
        const uint16_t *dyq = dy + q * 8;
        int tb = 0;
        for (int64_t qi = 0; qi < n_my; ++qi, ++tb) {
            const int64_t tile = tile_of(qi);
            const int bs = tb % BSTAGES;
            const int64_t row0 = tile * WG_TILE + lg * 32;
            mbar_wait(bar_bempty + 8 * bs, ((tb / BSTAGES) & 1) ^ 1);
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const uint32_t dst = smem_b + bs * Cfg::B_STAGE + nb * WG_BLOCK_BYTES;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const bool ok = row0 + i < n_out && q < Cfg::BQ; // lanes beyond a narrow row's chunks zero-fill
                    cp_async16(dst + (lane_off ^ uint32_t(i * 128 + ((i & 7) << 4))), ok ? dyq + (row0 + i) * YS + nb * 64 : dy, ok ? 16u : 0u);
                }
            }
            cp_async_arrive_noinc(bar_bfull + 8 * bs);
        }
        cp_async_wait_all();
    } else {
        // ================= kernel-map streamer (whole warp): every tap of both blocks of each live unit =================
        int e = 0;
        uint32_t eph = 0;
        const int32_t *lane_nbr = nbr + lane * 4;
        for (int64_t qi = 0; qi < n_my; ++qi) {
            const int64_t tile = tile_of(qi);
            const uint32_t live = live_units(tile);
            for (uint32_t rest = live; rest; rest &= rest - 1u) {
                const int blk = 2 * (unit0 + __ffs(rest) - 1);
                for (int i = 0; i < NS; ++i) { // one ring entry per gathered stage (the splits re-read the same entries)
                    mbar_wait(bar_iempty + 8 * e, eph ^ 1u);
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int sb = 0; sb < G; ++sb) {
                            const int tap = first_tap(blk + h) + sb;
                            if (blk + h < total_blocks && tap < k3)
                                cp_async16(smem_idx + e * Cfg::RING_BYTES + (h * G + sb) * Cfg::SUB_STRIDE + lane * 16,
                                           lane_nbr + int64_t(tap) * pitch + tile * WG_TILE, 16u);
                        }
                    cp_async_arrive_noinc(bar_ifull + 8 * e);
                    if (++e == RING) {
                        e = 0;
                        eph ^= 1u;
                    }
                }
            }
        }
        cp_async_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WG_WARP_MMA)
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---- host side ------------------------------------------------------------------------------------
struct WgradPlan {
    int groups, units_per_group, chunks, tiles_per_chunk, mini, seg_tiles;
};

constexpr int WG_SPLIT_SEG_TILES = 32; // fp32: drain the accumulators every 32 row tiles (<= 256 full-magnitude MMA steps)

// outputs up to 64 channels (bf16 / f16; 64 TMEM columns per accumulator) run two half-TMEM CTAs per SM
static inline int wgrad_ctas(int cin, int cout, bool split) {
    if (g_wgrad_variant == 2 && cin == 64 && cout == 64 && !split)
        return 1; // experiment: one full-TMEM CTA per SM
    return (!split && cout <= 64) ? 2 : 1;
}

static WgradPlan plan_wgrad(int64_t n_out, int cin, int cout, int k3, bool split) {
    WgradPlan p;
    const int ctas = wgrad_ctas(cin, cout, split);
    const int total_blocks = cin >= 64 ? k3 * (cin / 64) : int(ceil_div(k3, 64 / cin));
    const int total_units = (total_blocks + 1) / 2;
    const int max_units = (512 / ctas) / ((split ? 2 : 1) * (cout >= 64 ? cout : 64));
    p.groups = int(ceil_div(total_units, max_units));
    p.units_per_group = int(ceil_div(total_units, p.groups)); // balanced groups
    const int64_t tiles = ceil_div(n_out, WG_TILE);
    int64_t chunks = 148 * ctas / p.groups; // one wave of resident CTAs (TMEM: 512 columns per SM)
    if (chunks < 1)
        chunks = 1;
    if (chunks > tiles)
        chunks = tiles;
    // tiles per mini-chunk (see the kernel): 8 by default; knob: 5 = contiguous ranges (round 1), 6 / 7 / 8 = 16 / 32 / 64
    p.mini = g_wgrad_variant == 5 ? 0 : g_wgrad_variant == 6 ? 16 : g_wgrad_variant == 7 ? 32 : g_wgrad_variant == 8 ? 64 : 8;
    p.tiles_per_chunk = int(ceil_div(tiles, chunks));
    if (p.mini > 0 && p.tiles_per_chunk < 4 * p.mini) // small batches: keep at least ~4 mini-chunks per CTA, or the deal is uneven
        p.mini = p.tiles_per_chunk >= 4 ? p.tiles_per_chunk / 4 : 1;
    p.chunks = int(ceil_div(tiles, p.tiles_per_chunk));
    if (p.mini > 0) { // every CTA gets at least one mini-chunk
        const int64_t mcs = ceil_div(tiles, p.mini);
        p.chunks = int(chunks < mcs ? chunks : mcs);
    }
    p.seg_tiles = split ? WG_SPLIT_SEG_TILES : (1 << 30); // half precision: one segment, drained after the CTA's last tile
    return p;
}

// tc_split_rows (conv_tc.cu) writes the bf16 split rows of an fp32 operand

template <int CIN, int COUT, int STAGES, bool SPLIT = false, int CTAS = 1, int BSTG = 0, int MODE = 0>
static int launch_tc_wgrad(const WgradArgs &a, const void *x, const void *dy, float *partial) {
    using Cfg = TcWgradCfg<CIN, COUT, STAGES, SPLIT, CTAS, BSTG>;
    auto kernel = conv_tc_wgrad_kernel<CIN, COUT, STAGES, SPLIT, CTAS, BSTG, MODE>;
    const int ctas_planned = wgrad_ctas(CIN, COUT, SPLIT);
    FVC_REQUIRE(CTAS == ctas_planned, FVC_ERR_RUNTIME, "weight-gradient plan / kernel shape mismatch");
    static std::atomic<unsigned long long> configured{0}; // per instantiation, one bit per device
    const int rc_attr = ensure_dynamic_smem(kernel, Cfg::SMEM, configured);
    if (rc_attr)
        return rc_attr;
    const WgradPlan p = plan_wgrad(a.n_out, CIN, COUT, a.k3, SPLIT);
    const uint32_t idesc = make_idesc_f16(128, Cfg::NPAD, SPLIT || a.dtype == FVC_BF16, true, true);
    dim3 grid((unsigned)p.chunks, (unsigned)p.groups);
    kernel<<<grid, wg_threads(MODE), Cfg::SMEM, a.stream>>>(reinterpret_cast<const uint16_t *>(x), reinterpret_cast<const uint16_t *>(dy), a.nbr,
                                                      a.pitch, reinterpret_cast<const unsigned long long *>(a.tile_mask), a.n_out, a.k3,
                                                      p.units_per_group, p.tiles_per_chunk, p.mini, p.seg_tiles, idesc, partial);
    FVC_LAUNCH_CHECK();
    return wgrad_reduce_partials(partial, p.chunks, a.cin, a.cout, a.k3, a.dtype, a.grad_w, a.stream);
}

bool tc_wgrad_supported(int32_t cin, int32_t cout, int64_t k3, int32_t dtype) {
    const bool split = dtype == FVC_F32;
    if (dtype != FVC_F16 && dtype != FVC_BF16 && !split)
        return false;
    if (k3 < 1 || k3 > 4096)
        return false;
    auto pow2 = [](int c) { return c == 16 || c == 32 || c == 64 || c == 128 || c == 256; };
    return pow2(cin) && pow2(cout) && (!split || cout <= 128);
}

static inline size_t wg_partial_bytes(int64_t n_out, int32_t cin, int32_t cout, int64_t k3, bool split) {
    const WgradPlan p = plan_wgrad(n_out > 0 ? n_out : 1, cin, cout, int(k3), split);
    return align_up(size_t(p.chunks) * size_t(k3) * size_t(cin) * size_t(cout) * 4, 256);
}

// scratch = [fp32 partial slices | split X rows | split dY rows (fp32 only)]
size_t tc_wgrad_scratch_bytes(int64_t n_in, int64_t n_out, int32_t cin, int32_t cout, int64_t k3, int32_t dtype) {
    const bool split = dtype == FVC_F32;
    size_t bytes = wg_partial_bytes(n_out, cin, cout, k3, split) + 256;
    if (split)
        bytes += align_up(size_t(n_in > 0 ? n_in : 0) * 3 * size_t(cin) * 2, 256) + align_up(size_t(n_out > 0 ? n_out : 0) * 3 * size_t(cout) * 2, 256);
    return bytes;
}

int tc_wgrad(const WgradArgs &a) {
    const bool split = a.dtype == FVC_F32;
    const size_t x_rows = (split && !a.x_split) ? align_up(size_t(a.n_in) * 3 * size_t(a.cin) * 2, 256) : 0;
    const size_t dy_rows = (split && !a.dy_split) ? align_up(size_t(a.n_out) * 3 * size_t(a.cout) * 2, 256) : 0;
    const size_t need = wg_partial_bytes(a.n_out, a.cin, a.cout, a.k3, split) + x_rows + dy_rows;
    FVC_REQUIRE(a.scratch && a.scratch_bytes >= need, FVC_ERR_RUNTIME, "tensor-core wgrad scratch too small: %zu < %zu",
                a.scratch_bytes, need);
    FVC_REQUIRE((reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.dy) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a.scratch) & 255) == 0 && (reinterpret_cast<uintptr_t>(a.nbr) & 15) == 0,
                FVC_ERR_RUNTIME, "tensor-core wgrad needs 16-byte aligned feature / grad / map pointers and 256-byte aligned scratch");
    FVC_REQUIRE(a.pitch % 4 == 0 && a.pitch >= ceil_div(a.n_out, WG_TILE) * WG_TILE, FVC_ERR_RUNTIME,
                "tensor-core wgrad needs the map pitch (%lld) to be a multiple of 4 covering whole 128-row tiles", (long long)a.pitch);
    float *partial = reinterpret_cast<float *>(a.scratch);
    if (split) {
        uint8_t *base = reinterpret_cast<uint8_t *>(a.scratch) + wg_partial_bytes(a.n_out, a.cin, a.cout, a.k3, true);
        const void *xs = a.x, *dys = a.dy; // already split by the caller (the rows a layer's forward / dgrad call produced), or split here
        if (!a.x_split) {
            const int rc = tc_split_rows(reinterpret_cast<const float *>(a.x), a.n_in, a.cin, reinterpret_cast<uint16_t *>(base), a.stream);
            if (rc)
                return rc;
            xs = base;
        }
        if (!a.dy_split) {
            const int rc = tc_split_rows(reinterpret_cast<const float *>(a.dy), a.n_out, a.cout, reinterpret_cast<uint16_t *>(base + x_rows), a.stream);
            if (rc)
                return rc;
            dys = base + x_rows;
        }
#define FVC_WGS_CASE(CI, CO, S)      \
    if (a.cin == CI && a.cout == CO) \
        return launch_tc_wgrad<CI, CO, S, true>(a, xs, dys, partial);
#define FVC_WGS_CIN(CI)      \
    FVC_WGS_CASE(CI, 16, 3)  \
    FVC_WGS_CASE(CI, 32, 3)  \
    FVC_WGS_CASE(CI, 64, 3)  \
    FVC_WGS_CASE(CI, 128, 2)
        FVC_WGS_CIN(16)
        FVC_WGS_CIN(32)
        FVC_WGS_CIN(64)
        FVC_WGS_CIN(128)
        FVC_WGS_CIN(256)
#undef FVC_WGS_CIN
#undef FVC_WGS_CASE
        return set_error(FVC_ERR_UNSUPPORTED, "no fp32 tensor-core wgrad kernel for channels %d -> %d", a.cin, a.cout);
    }
    // experiment knob (fvc_set_tuning(1, v)): the 64 -> 64 shape with two dY tile buffers per CTA (v = 1) / three gather stages (v = 2)
    if (g_wgrad_variant == 1 && a.cin == 64 && a.cout == 64)
        return launch_tc_wgrad<64, 64, 2, false, 2, 1>(a, a.x, a.dy, partial); // round-1 shape: one dY buffer
    if (g_wgrad_variant == 2 && a.cin == 64 && a.cout == 64)
        return launch_tc_wgrad<64, 64, 4, false, 1, 2>(a, a.x, a.dy, partial);
    // two pipelines per CTA (MODE 1) are the default of the 128- / 256-wide outputs (one CTA per SM: 128 -> 128 1.348 -> 1.265 ms,
    // 256 -> 256 4.60 -> 4.51 ms on the C2 batch); on 64 -> 64, which runs two CTAs per SM, they lose (0.562 -> 0.591 ms): knob only
    if (g_wgrad_variant == 3 && a.cin == 64 && a.cout == 64)
        return launch_tc_wgrad<64, 64, 2, false, 2, 0, 1>(a, a.x, a.dy, partial);
    if (g_wgrad_variant == 3 && a.cin == 128 && a.cout == 128)
        return launch_tc_wgrad<128, 128, 4>(a, a.x, a.dy, partial); // the single-pipeline shape (A/B baseline)
    if (g_wgrad_variant == 3 && a.cin == 256 && a.cout == 256)
        return launch_tc_wgrad<256, 256, 2>(a, a.x, a.dy, partial);
#define FVC_WG_CASE(CI, CO, S)       \
    if (a.cin == CI && a.cout == CO) \
        return launch_tc_wgrad<CI, CO, S>(a, a.x, a.dy, partial);
#define FVC_WG_CASE2(CI, CO, S)      \
    if (a.cin == CI && a.cout == CO) \
        return launch_tc_wgrad<CI, CO, S, false, 2>(a, a.x, a.dy, partial);
#define FVC_WG_CASEP(CI, CO, S)      \
    if (a.cin == CI && a.cout == CO) \
        return launch_tc_wgrad<CI, CO, S, false, 1, 0, 1>(a, a.x, a.dy, partial);
#define FVC_WG_CIN(CI)       \
    FVC_WG_CASE2(CI, 16, 2)  \
    FVC_WG_CASE2(CI, 32, 2)  \
    FVC_WG_CASE2(CI, 64, 2)  \
    FVC_WG_CASE(CI, 128, 4)  \
    FVC_WG_CASE(CI, 256, 2)
#define FVC_WG_CIN_WIDE(CI)  \
    FVC_WG_CASE2(CI, 16, 2)  \
    FVC_WG_CASE2(CI, 32, 2)  \
    FVC_WG_CASE2(CI, 64, 2)  \
    FVC_WG_CASEP(CI, 128, 4) \
    FVC_WG_CASEP(CI, 256, 2)
    FVC_WG_CIN(16)
    FVC_WG_CIN(32)
    FVC_WG_CIN_WIDE(64)
    FVC_WG_CIN_WIDE(128)
    FVC_WG_CIN_WIDE(256)
#undef FVC_WG_CIN_WIDE
#undef FVC_WG_CASEP
#undef FVC_WG_CIN
#undef FVC_WG_CASE
#undef FVC_WG_CASE2
    return set_error(FVC_ERR_UNSUPPORTED, "no tensor-core wgrad kernel for channels %d -> %d", a.cin, a.cout);
}

} // namespace fvc
