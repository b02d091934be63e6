// conv_tc.cu -- tcgen05 / TMEM implicit-GEMM sparse convolution for f16 / bf16 (fp32 accumulate) and for fp32
// (three-way bf16 split, see "fp32 on the tensor pipe" below).
//
// Replaces PredGatherIGemm.cu (SM80 mma.sync TF32, forward only, one CTA per leaf) and the cuBLAS
// gather -> mm -> atomic-scatter pipeline of GatherScatterDefault.cu:706-721,786-808 for the half types.
//
// Forward / dgrad kernel (output-stationary, no atomics):
//   a CTA owns TILES consecutive 128-row output tiles; their fp32 accumulators (128 lanes x COUT columns
//   each) stay in TMEM for the whole tap loop.  For every (tap, 64-channel block) the weight chunk
//   B = W[tap][:, block] (COUT x 64, K-major, 128B-swizzled image) arrives by ONE cp.async.bulk; then for each
//   tile the gather producers copy the 128 neighbour rows (128 B each) with 16-byte zero-filling cp.async
//   straight into the canonical K-major SWIZZLE_128B layout, and one elected thread issues
//   4 x tcgen05.mma (M=128, N=COUT, K=16) accumulating into that tile's TMEM slice.  Stages are recycled by
//   tcgen05.commit -> mbarrier.  Epilogue: tcgen05.ld -> (+bias, *scale+shift, +residual, ReLU, column
//   statistics) -> bf16/f16 -> one plain 128-bit store stream per output row.
//
// Two producer organisations (template MODE):
//   MODE 1 ("warp per unit", Cin >= 64; the default there): producer warp w gathers WHOLE units u = w, w + PW, ... --
//     32 x 16-byte cp.async per lane behind ONE stage hand-shake, so PW independent wait -> gather -> arrive chains
//     run concurrently and the per-unit fixed latencies (two mbarrier round trips, ~200 cycles) are paid once per
//     16 KB instead of once per 4 KB by every warp.  A lane reads the unit's 128 map entries as one 16-byte load two
//     units ahead and the warp redistributes them by shuffle: no map ring, no streamer warp, two barriers less per unit.
//     (Round 1's isolation runs had shown the hand-off skeleton, not HBM / L2 / the tensor pipe, to bound the kernel.)
//   MODE 0 (packed taps, Cin = 16 / 32): 4 producer warps each copy a quarter of every unit; one warp streams the
//     kernel-map entries of upcoming units into a shared-memory ring (profiles/r01_experiments.md).
//
// The list of live (tap, channel block, tile) units is built once per CTA in shared memory; every role walks that list.
//
// fp32 on the tensor pipe (SPLIT = true).  An fp32 value is the exact sum of three bf16 values x = x0 + x1 + x2
// (8 + 8 + 8 mantissa bits).  A pre-pass writes the split features [N][3][Cin] and the split weight image once per
// call (N*Cin elements, not P*Cin); the executor gathers the three split rows like three channel blocks and issues
//   x0.(w0 + w1 + w2)  +  x1.(w0 + w1)  +  x2.w0          (six bf16 products, exact in fp32; dropped terms <= 2^-24)
// The x0.w0 term accumulates in its own TMEM columns and the five small terms (<= 2^-8 of it) in a second set, added
// once in the epilogue: tcgen05 truncates on every accumulate (measured: -0.5 ulp per MMA on same-sign data,
// scripts/exp_accum_precision.py), so the full-magnitude chain must stay as short as the bf16 one.
#include "conv_internal.cuh"
#include "tc_ptx.cuh"
#include "tc_math.cuh"

#include <cstring>

namespace fvc {

using namespace tc;

int g_tc_variant = 0;

constexpr int TC_IDX_RING = 8;              // map-entry ring depth (units of 128 int32), power of two (MODE 0)
constexpr int TC_TILE_M = 128;
constexpr int TC_A_BYTES = TC_TILE_M * 128; // one stage: 128 rows x 64 channels x 2 B
constexpr int TC_MASK_WORDS = 8;            // tile tap-mask words (K^3 <= 512)
constexpr int TC_MAX_UNITS = 4096;          // capacity of the per-CTA unit list (uint16 entries)
constexpr int TC_MAX_UNITS_SPLIT = 2048;    // ... of the small / fp32 shapes (more shared memory goes elsewhere)

constexpr int tmem_cols_for(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

// Small channel counts are packed: a 64-wide reduction block holds G = 64 / CIN consecutive taps x CIN channels
// ("tap group"), so Cin = 16 / 32 feed the same K = 64 pipeline (the bandwidth-bound small-channel path).
template <int CIN, int COUT, int TILES, int STAGES, int BST, int PW, bool SPLIT = false, int MODE = 0> struct TcFwdCfg {
    static constexpr int G = CIN >= 64 ? 1 : 64 / CIN;      // taps per reduction block
    static constexpr int KB = CIN >= 64 ? CIN / 64 : 1;     // 64-wide reduction blocks per tap group
    static constexpr int CPT = CIN >= 64 ? 8 : CIN / 8;     // 16-byte chunks one tap contributes to a 128-byte row
    static constexpr int NS = SPLIT ? 3 : 1;                 // bf16 splits of an fp32 operand (A blocks and B chunks per (group, block))
    static constexpr int XS = NS * CIN;                      // feature row stride in elements
    static constexpr int CHUNK_BYTES = COUT * 128;           // one weight chunk: COUT rows x 64 channels x 2 B
    static constexpr int B_BYTES = NS * CHUNK_BYTES;         // one weight stage: the chunk(s) of one (tap group, channel block)
    static constexpr int ACC_COLS = (SPLIT ? 2 : 1) * COUT;  // TMEM columns per tile (SPLIT: main | small-term accumulator)
    static constexpr int TMEM_COLS = tmem_cols_for(TILES * ACC_COLS);
    static constexpr bool RING = MODE == 0;                  // kernel-map ring + streamer warp
    // MODE 3: NP pipelines per CTA -- each = two producer warps, its own MMA-issuing warp, its own gather stages and its own
    // TILES / NP tiles -- sharing ONE stream of weight chunks: a fat CTA loads every live chunk once for all of its tiles
    // (the L2 -> SM traffic of a wide layer is mostly re-streamed weights) without falling back on one
    // issuing thread for the whole SM (~850 dependent cycles per unit, profiles/r02_ts_executor.md)
    static constexpr int NP = MODE == 3 ? PW / 2 : 1;
    static constexpr int SP = STAGES / NP;                   // gather stages per pipeline
    static constexpr int TP = TILES / NP;                    // tiles per pipeline
    // unit-list capacity: packed taps (Cin < 64) never exceed 128 groups x 8 tiles; the small and fp32 shapes trade list
    // space for a co-resident CTA / weight chunks
    static constexpr int MAX_UNITS = (SPLIT || TILES <= 2 || CIN < 64) ? TC_MAX_UNITS_SPLIT : TC_MAX_UNITS;
    static constexpr int WARPS = RING ? PW + 3 : PW + NP + 1; // producers | MMA issuer(s) | weight loader | (map streamer)
    static constexpr int THREADS = WARPS * 32;
    static constexpr int NUM_BARS = 2 * STAGES + 2 * BST + 1 + (RING ? 2 * TC_IDX_RING : 0);
    // one ring entry: 128 map entries per tap of the group; with several taps per entry each tap's 512 bytes are followed
    // by a 16-byte pad, so the four taps a quarter-warp reads together sit on different banks
    static constexpr int SUB_STRIDE = G > 1 ? 528 : 512;
    static constexpr int RING_BYTES = G * SUB_STRIDE;
    static constexpr size_t SMEM = 1024 + size_t(STAGES) * TC_A_BYTES + size_t(BST) * B_BYTES + size_t(RING ? TC_IDX_RING : 0) * RING_BYTES +
                                   size_t(MAX_UNITS) * 2 * (MODE == 3 ? 2 : 1) + 8 * NUM_BARS + 16;
    // co-resident CTAs: limited by TMEM columns (512 per SM) and shared memory (228 KB per SM, 1 KB reserved per CTA).
    // Three small CTAs per SM beat two larger ones by ~10 % on the 32- and 64-channel shapes: more independent
    // producer -> MMA -> commit chains hide the per-unit hand-off latency (profiles/r01_experiments.md)
    static constexpr int BY_TMEM = 512 / TMEM_COLS, BY_SMEM = int(233472 / (SMEM + 1024 + 768));
    static constexpr int CTA_CAP = MODE == 1 ? 4 : 3;
    static constexpr int CTAS_PER_SM = BY_TMEM < BY_SMEM ? (BY_TMEM > CTA_CAP ? CTA_CAP : BY_TMEM) : (BY_SMEM > CTA_CAP ? CTA_CAP : BY_SMEM);
    static_assert(CTAS_PER_SM >= 1, "configuration does not fit one SM");
    static_assert((CIN % 64 == 0 || CIN == 32 || CIN == 16) && COUT % 16 == 0 && COUT >= 16 && COUT <= 256, "unsupported channel counts");
    static_assert(TILES * ACC_COLS <= 512 && TILES <= 8 && KB <= 4, "accumulators exceed TMEM / unit encoding");
    static_assert(MODE == 0 ? PW == 4 : MODE == 3 ? (PW % 2 == 0 && PW >= 4 && PW <= 8 && G == 1 && !SPLIT) : (PW >= 2 && PW <= 4 && G == 1), "producer warps / mode");
    static_assert(MODE != 3 || (STAGES % NP == 0 && SP >= 2 && TILES % NP == 0), "pipelines: two producer warps need two stages each");
    static_assert(WARPS >= 4, "the epilogue needs one warp per TMEM lane quarter");
    // warp-per-unit producers hold PW units in flight: with fewer stages a warp could meet a stage barrier two phases behind,
    // which a parity wait cannot tell from a ready one
    static_assert(MODE != 1 || STAGES >= PW, "warp-per-unit producers need at least PW stages");
};

__device__ __forceinline__ float load_any_f(const void *p, int64_t i, int dtype) {
    switch (dtype) {
    case FVC_F16: return __half2float(reinterpret_cast<const __half *>(p)[i]);
    case FVC_BF16: return __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(p)[i]);
    case FVC_F32: return reinterpret_cast<const float *>(p)[i];
    default: return float(reinterpret_cast<const double *>(p)[i]);
    }
}

// x = s0 + s1 + s2 with every s_i a bf16 (round-to-nearest each time; the remainders are exact in fp32)
__device__ __forceinline__ void split3(float v, uint16_t (&out)[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        out[i] = *reinterpret_cast<const uint16_t *>(&h);
        v -= __bfloat162float(h);
    }
}

struct WeightSource { // the public [Cout, Cin, k0, k1, k2] tensor (any strides), or the packed [K][Cin][Cout] array
    const void *w;
    int64_t s_n, s_k, s_t0, s_t1, s_t2; // element strides of (output channel n, reduction channel kk, tap axes)
    int k1, k2, k3, flip, dtype_in;
};

__device__ __forceinline__ float weight_at(const WeightSource &ws, int tap, int kk, int n) {
    const int t = ws.flip ? ws.k3 - 1 - tap : tap;
    const int t0 = t / (ws.k1 * ws.k2), t1 = (t / ws.k2) % ws.k1, t2 = t % ws.k2;
    return load_any_f(ws.w, n * ws.s_n + kk * ws.s_k + t0 * ws.s_t0 + t1 * ws.s_t1 + t2 * ws.s_t2, ws.dtype_in);
}

// Weight image: chunk c = (group * KB + j) * NS + i holds COUT rows of 128 B; row n = the 64 reduction elements of that
// chunk for output channel n (Cin >= 64: channels [64j, 64j+64) of tap `group`; Cin < 64: taps [G*group, G*group+G) x Cin
// channels, zero beyond the last tap), 16-byte chunk q stored at position q ^ (n & 7) (SWIZZLE_128B K-major atom);
// i = bf16 split of an fp32 weight (NS = 3) or 0.  One launch from the public layout: replaces
// `weights.permute(2,3,4,1,0).reshape(K,Cin,Cout).contiguous()` (GatherScatterDefault.cu:691-694) + the operand staging.
__global__ void tc_weight_image_kernel(WeightSource ws, int k3, int cin, int cout, int ns, int bf16_out, uint4 *__restrict__ img) {
    const int g = cin >= 64 ? 1 : 64 / cin, kb = cin >= 64 ? cin / 64 : 1;
    const int groups = (k3 + g - 1) / g;
    const int64_t total = int64_t(groups) * kb * ns * cout * 8; // 16-byte chunks
    for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
        const int pos = int(e & 7);
        const int n = int((e >> 3) % cout);
        const int64_t chunk_s = (e >> 3) / cout;
        const int i = int(chunk_s % ns);
        const int64_t chunk = chunk_s / ns;
        const int j = int(chunk % kb), group = int(chunk / kb);
        const int q = pos ^ (n & 7);
        uint32_t v[4] = {0u, 0u, 0u, 0u};
        const int kk0 = q * 8; // first of the 8 reduction elements of this 16-byte chunk
        const int tap = cin >= 64 ? group : group * g + kk0 / cin;
        const int ci0 = cin >= 64 ? j * 64 + kk0 : kk0 % cin;
        if (tap < k3) {
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const float a = weight_at(ws, tap, ci0 + 2 * h, n), b = weight_at(ws, tap, ci0 + 2 * h + 1, n);
                if (ns == 3) {
                    uint16_t lo[3], hi[3];
                    split3(a, lo);
                    split3(b, hi);
                    v[h] = uint32_t(lo[i]) | (uint32_t(hi[i]) << 16);
                } else if (bf16_out) {
                    const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
                    v[h] = *reinterpret_cast<const uint32_t *>(&p);
                } else {
                    const __half2 p = __floats2half2_rn(a, b);
                    v[h] = *reinterpret_cast<const uint32_t *>(&p);
                }
            }
        }
        img[e] = make_uint4(v[0], v[1], v[2], v[3]);
    }
}

// fp32 rows [n][c] -> bf16 split rows [n][3][c]; one thread per 8 consecutive channels (32 B in, 3 x 16 B out)
__global__ void tc_split_rows_kernel(const float *__restrict__ x, int64_t n, int c, uint16_t *__restrict__ xs) {
    const int c8 = c >> 3;
    const int64_t total = n * c8;
    for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
        const int64_t row = e / c8;
        const int ch = int(e - row * c8) * 8;
        const float4 a = __ldg(reinterpret_cast<const float4 *>(x + row * c + ch)), b = __ldg(reinterpret_cast<const float4 *>(x + row * c + ch) + 1);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t packed[3][4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            uint16_t lo[3], hi[3];
            split3(v[2 * h], lo);
            split3(v[2 * h + 1], hi);
#pragma unroll
            for (int i = 0; i < 3; ++i)
                packed[i][h] = uint32_t(lo[i]) | (uint32_t(hi[i]) << 16);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
            *reinterpret_cast<uint4 *>(xs + (row * 3 + i) * c + ch) = make_uint4(packed[i][0], packed[i][1], packed[i][2], packed[i][3]);
    }
}

// x: feature rows (SPLIT: the bf16 split rows [N][3][CIN]); bias / residual / y: in the output dtype (SPLIT: fp32)
template <int CIN, int COUT, int TILES, int STAGES, int BST, int PW, bool SPLIT, int MODE>
__global__ void __launch_bounds__((TcFwdCfg<CIN, COUT, TILES, STAGES, BST, PW, SPLIT, MODE>::THREADS), (TcFwdCfg<CIN, COUT, TILES, STAGES, BST, PW, SPLIT, MODE>::CTAS_PER_SM))
conv_tc_fwd_kernel(const uint16_t *__restrict__ x, const uint8_t *__restrict__ w_img, const Epilogue epi, void *__restrict__ y_,
                   const int32_t *__restrict__ nbr, int64_t pitch, const unsigned long long *__restrict__ tile_mask, int64_t n_in, int64_t n_out,
                   int k3, uint32_t idesc, int is_bf16) {
    using Cfg = TcFwdCfg<CIN, COUT, TILES, STAGES, BST, PW, SPLIT, MODE>;
    constexpr int KB = Cfg::KB, G = Cfg::G, CPT = Cfg::CPT, THREADS = Cfg::THREADS, NS = Cfg::NS, XS = Cfg::XS, ACC = Cfg::ACC_COLS;
    constexpr bool RING = Cfg::RING;
    constexpr int NP = Cfg::NP, SP = Cfg::SP, TP = Cfg::TP, PCAP = Cfg::MAX_UNITS / NP;
    constexpr int WARP_MMA = PW, WARP_B = PW + NP; // MMA issuers: warps PW .. PW + NP - 1
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u; // SWIZZLE_128B atoms need 1024-byte alignment
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_a + STAGES * TC_A_BYTES;
    const uint32_t smem_idx = smem_b + BST * Cfg::B_BYTES;
    const uint32_t smem_units = smem_idx + (RING ? TC_IDX_RING : 0) * Cfg::RING_BYTES;
    const uint32_t bars = smem_units + Cfg::MAX_UNITS * 2 * (MODE == 3 ? 2 : 1);
    const uint32_t bar_full = bars, bar_empty = bars + 8 * STAGES;
    const uint32_t bar_bfull = bars + 16 * STAGES, bar_bempty = bar_bfull + 8 * BST;
    const uint32_t bar_accum = bar_bempty + 8 * BST;
    const uint32_t bar_ifull = bar_accum + 8, bar_iempty = bar_ifull + 8 * TC_IDX_RING; // (RING only)
    const uint32_t tmem_slot = bar_accum + 8 + (RING ? 16 * TC_IDX_RING : 0);
    uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));
    uint16_t *units = reinterpret_cast<uint16_t *>(smem_gen + (smem_units - smem_base));
    uint16_t *units_p = units + Cfg::MAX_UNITS; // MODE 3: pipeline p's own units, in order, at units_p + p * PCAP
    __shared__ int s_np[4];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t total_tiles = (n_out + TC_TILE_M - 1) / TC_TILE_M;
    const int64_t tile0 = int64_t(blockIdx.x) * TILES;
    const int ntiles = int(total_tiles - tile0 < TILES ? total_tiles - tile0 : TILES);

    // ---- prologue: which tiles does each tap group reach (bit t of s_tiles[g]), then the ordered list of live units ----
    __shared__ uint8_t s_tiles[64 * TC_MASK_WORDS];
    __shared__ int s_nunits, s_live;
    const int ngroups = (k3 + G - 1) / G;
    {
        const int words = (k3 + 63) >> 6;
        for (int g = threadIdx.x; g < ngroups; g += THREADS) {
            uint32_t bits = 0;
            for (int sub = 0; sub < G; ++sub) {
                const int k = g * G + sub;
                if (k >= k3)
                    break;
                for (int t = 0; t < ntiles; ++t) {
                    const unsigned long long m = tile_mask ? __ldg(tile_mask + (tile0 + t) * words + (k >> 6)) : ~0ull;
                    bits |= uint32_t((m >> (k & 63)) & 1ull) << t;
                }
            }
            s_tiles[g] = uint8_t(bits);
        }
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, RING ? PW * 32 : 32); // one completion-triggered arrival per gathering thread
            mbar_init(bar_empty + 8 * s, 1);                  // tcgen05.commit
        }
        for (int b = 0; b < BST; ++b) {
            mbar_init(bar_bfull + 8 * b, 1); // expect_tx arrival + bytes
            mbar_init(bar_bempty + 8 * b, NP); // every pipeline releases every chunk
        }
        mbar_init(bar_accum, NP);
        if (RING) {
            for (int e = 0; e < TC_IDX_RING; ++e) {
                mbar_init(bar_ifull + 8 * e, 32);  // one completion-triggered arrival per streamer lane
                mbar_init(bar_iempty + 8 * e, PW); // one arrival per producer warp
            }
        }
        fence_mbar_init();
    }
    if (warp == WARP_MMA)
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    __syncthreads();
    if (warp == 0) { // unit list in [tap group][channel block][split][tile] order: entry = group << 7 | block << 5 | split << 3 | tile
        int base = 0;
        uint32_t live = 0;
        for (int k0 = 0; k0 < ngroups; k0 += 32) {
            const int k = k0 + lane;
            const uint32_t bits = k < ngroups ? s_tiles[k] : 0u;
            const int cnt = KB * NS * __popc(bits);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d)
                    incl += v;
            }
            int off = base + incl - cnt;
            for (int j = 0; j < KB; ++j)
                for (int i = 0; i < NS; ++i)
                    for (uint32_t rest = bits; rest; rest &= rest - 1u)
                        units[off++] = uint16_t((k << 7) | (j << 5) | (i << 3) | (__ffs(rest) - 1));
            base += __shfl_sync(0xffffffffu, incl, 31);
            live |= bits;
        }
#pragma unroll
        for (int d = 16; d; d >>= 1)
            live |= __shfl_xor_sync(0xffffffffu, live, d);
        if (lane == 0) {
            s_nunits = base;
            s_live = int(live);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const int nunits = s_nunits;
    if constexpr (MODE == 3) { // warp p < NP compacts pipeline p's units (tiles p TP .. p TP + TP - 1) out of the CTA's list, in order
        if (warp < NP) {
            int base = 0;
            for (int u0 = 0; u0 < nunits; u0 += 32) {
                const int u = u0 + lane;
                const uint32_t unit = u < nunits ? units[u] : 0u;
                const bool mine = u < nunits && int(unit & 7u) / TP == warp;
                const uint32_t ballot = __ballot_sync(0xffffffffu, mine);
                if (mine)
                    units_p[warp * PCAP + base + __popc(ballot & ((1u << lane) - 1u))] = uint16_t(unit);
                base += __popc(ballot);
            }
            if (lane == 0)
                s_np[warp] = base;
        }
        __syncthreads();
    }

    if (warp < PW) {
        if constexpr (!RING) {
            // ================= MODE 1 gather producers: warp w gathers whole units u = w, w + PW, ... =================
            // lane = (lg, q): 8 lanes q cover one 128-byte row (one full line); instruction i of a unit copies rows
            // i + 32 lg, lg = 0..3.  The unit's 128 map entries arrive as ONE 16-byte load per lane (lane l holds rows
            // 4l .. 4l+3), two of this warp's units ahead; row i + 32 lg then comes from lane (i >> 2) + 8 lg, component i & 3.
            const int q = lane & 7, lg = lane >> 3;
            const uint32_t lane_off = (uint32_t(lg) << 12) | (uint32_t(q) << 4); // row 32 lg, chunk q; XOR with the per-i constant below
            const int32_t *lane_nbr = nbr + tile0 * TC_TILE_M + lane * 4;
            const int src_lane0 = 8 * lg;
            // whose units: the CTA's (MODE 1: warp w takes u = w, w + PW, ...) or this warp's pipeline's (MODE 3: the two
            // producer warps of pipeline p alternate over its list and fill its SP stages)
            const uint16_t *ulist = MODE == 3 ? units_p + (warp >> 1) * PCAP : units;
            const int n_own = MODE == 3 ? s_np[warp >> 1] : nunits;
            const int first = MODE == 3 ? (warp & 1) : warp, stage0 = MODE == 3 ? (warp >> 1) * SP : 0;
            constexpr int STRIDE = MODE == 3 ? 2 : PW;
            auto load_idx = [&](int uu) -> int4 {
                if (uu >= n_own)
                    return make_int4(-1, -1, -1, -1);
                const uint32_t un = ulist[uu];
                return __ldg(reinterpret_cast<const int4 *>(lane_nbr + int64_t(un >> 7) * pitch + int(un & 7) * TC_TILE_M));
            };
            int4 pf0 = load_idx(first), pf1 = load_idx(first + STRIDE);
            for (int u = first; u < n_own; u += STRIDE) {
                const uint32_t unit = ulist[u];
                const int j = (unit >> 5) & 3, t = unit & 7;
                const int4 cur = pf0;
                pf0 = pf1;
                pf1 = load_idx(u + 2 * STRIDE);
                const int s = stage0 + u % SP;
                const int64_t rows_left = n_out - (tile0 + t) * TC_TILE_M - 32 * lg; // row i + 32 lg exists iff i < rows_left
                const uint16_t *xj = x + q * 8 + j * 64 + (SPLIT ? int((unit >> 3) & 3u) * CIN : 0);
                mbar_wait(bar_empty + 8 * s, ((u / SP) & 1) ^ 1u);
                const uint32_t dst = smem_a + s * TC_A_BYTES;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int comp = (i & 3) == 0 ? cur.x : (i & 3) == 1 ? cur.y : (i & 3) == 2 ? cur.z : cur.w;
                    const int idx = __shfl_sync(0xffffffffu, comp, src_lane0 + (i >> 2));
                    const bool ok = idx >= 0 && i < rows_left;
                    // row r = i + 32 lg: offset r * 128 + ((q ^ (r & 7)) << 4) == lane_off ^ (i * 128 + ((i & 7) << 4))
                    cp_async16(dst + (lane_off ^ uint32_t(i * 128 + ((i & 7) << 4))), ok ? xj + int64_t(idx) * XS : x, ok ? 16u : 0u);
                }
                // completion-triggered arrival (the CUTLASS sm100 cp.async -> UMMA idiom): the producer never blocks on
                // its own loads, so every stage this warp owns can be in flight
                cp_async_arrive_noinc(bar_full + 8 * s);
            }
            cp_async_wait_all();
        } else {
            // ================= MODE 0 gather producers: warp w copies rows [RPW*w, RPW*(w+1)) of every unit =================
            // lane = (lg, q): 8 lanes q cover one 128-byte row (one full line).  With one tap per row (Cin >= 64) lane group
            // lg takes rows lg, lg + 4, ... (an instruction copies 4 consecutive output rows, whose neighbours tend to be
            // consecutive feature rows).  With packed taps (Cin < 64) it takes NI CONSECUTIVE rows instead, so that its map
            // entries are contiguous (128-bit shared loads, no bank conflicts between the taps a quarter-warp reads).
            constexpr int RPW = TC_TILE_M / PW, NI = RPW / 4;
            constexpr bool SEQ = G > 1;
            const int q = lane & 7;
            const int row0 = warp * RPW + (lane >> 3) * (SEQ ? NI : 1); // rows row0 + i * (SEQ ? 1 : 4), i < NI
            // lane q copies 16-byte chunk q of each of its rows; with packed taps that chunk belongs to tap `sub` of the
            // group and to channel chunk q % CPT of that tap's feature row
            const int sub = q / CPT;
            const uint16_t *xq = x + (CIN >= 64 ? q * 8 : (q % CPT) * 8);
            uint32_t dst_off[NI]; // row offset + swizzled chunk position inside a stage
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const int r = row0 + i * (SEQ ? 1 : 4);
                dst_off[i] = uint32_t(r) * 128u + (uint32_t(q ^ (r & 7)) << 4);
            }
            const uint32_t idx_src = smem_idx + sub * Cfg::SUB_STRIDE + row0 * 4;
            int s = 0;
            uint32_t ph = 0;
            for (int u = 0; u < nunits; ++u) {
                const uint32_t unit = units[u];
                const int g = int(unit >> 7), j = (unit >> 5) & 3, t = unit & 7;
                const int e = u & (TC_IDX_RING - 1);
                mbar_wait(bar_ifull + 8 * e, (u / TC_IDX_RING) & 1);
                int idx[NI];
                if (SEQ) {
#pragma unroll
                    for (int i = 0; i < NI; i += 4)
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(idx[i]), "=r"(idx[i + 1]), "=r"(idx[i + 2]), "=r"(idx[i + 3])
                                     : "r"(idx_src + e * Cfg::RING_BYTES + i * 4)
                                     : "memory");
                } else {
#pragma unroll
                    for (int i = 0; i < NI; ++i)
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(idx[i]) : "r"(idx_src + e * Cfg::RING_BYTES + i * 16) : "memory");
                }
                // row row0 + i * step exists iff i * step < rows_left; a tap beyond the kernel volume (last, partial group) is empty
                const int64_t rows_left = (G == 1 || g * G + sub < k3) ? n_out - (tile0 + t) * TC_TILE_M - row0 : 0;
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                const uint32_t dst = smem_a + s * TC_A_BYTES;
                const uint16_t *xj = xq + j * 64 + (SPLIT ? int((unit >> 3) & 3u) * CIN : 0);
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const bool ok = idx[i] >= 0 && i * (SEQ ? 1 : 4) < rows_left;
                    cp_async16(dst + dst_off[i], ok ? xj + int64_t(idx[i]) * XS : x, ok ? 16u : 0u);
                }
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(bar_iempty + 8 * e); // ring entry consumed by this warp (values are in registers)
                cp_async_arrive_noinc(bar_full + 8 * s);
                if (++s == STAGES) {
                    s = 0;
                    ph ^= 1u;
                }
            }
            cp_async_wait_all();
        }
    } else if (warp >= WARP_MMA && warp < WARP_MMA + NP) {
        // ================= MMA issuer of pipeline p (one thread): walks the CTA's unit list, issues its own tiles' units,
        //                   and takes part in every weight chunk's hand-shake (a chunk is free once ALL pipelines are past it) =====
        if (elect_one()) { // (not `lane == 0`: see tc_ptx.cuh)
            const int p = NP > 1 ? warp - WARP_MMA : 0; // (a compile-time 0 for the single-pipeline shapes: nothing extra on their issuing thread)
            // descriptor = {hi: SBO 1024 | version 1 | SWIZZLE_128B, lo: (addr >> 4) | LBO 16 B}; only lo changes
            const uint64_t desc_hi = make_smem_desc_sw128(0, 16, 1024) & 0xFFFFFFFF00000000ull;
            const uint32_t a_lo0 = ((smem_a & 0x3FFFFu) >> 4) | (1u << 16), b_lo0 = ((smem_b & 0x3FFFFu) >> 4) | (1u << 16);
            int sl = 0, c = -1, prev_kj = -1; // sl: this pipeline's next stage (of its SP), ph: its phase
            uint32_t ph = 0, started = 0, started_small = 0, b_lo = 0;
            uint32_t next_unit = nunits > 0 ? units[0] : 0u;
            for (int u = 0; u < nunits; ++u) {
                const uint32_t unit = next_unit;
                next_unit = units[u + 1 < nunits ? u + 1 : u]; // (one entry ahead: its shared-memory latency is off this thread's per-unit chain)
                const int kj = int(unit >> 5), t = unit & 7;
                if (kj != prev_kj) { // next weight stage
                    if (c >= 0)
                        umma_commit(bar_bempty + 8 * (c % BST));
                    ++c;
                    mbar_wait(bar_bfull + 8 * (c % BST), (c / BST) & 1);
                    b_lo = b_lo0 + uint32_t(c % BST) * (Cfg::B_BYTES >> 4);
                    prev_kj = kj;
                }
                if (NP > 1 && t / TP != p)
                    continue; // another pipeline's tile
                const int s = p * SP + sl;
                mbar_wait(bar_full + 8 * s, ph);
                // the gathered rows were written through the generic proxy (cp.async); tcgen05.mma reads shared memory
                // through the async proxy: order the two before the first MMA of the stage
                fence_proxy_async();
                tc_fence_after();
                const uint32_t a_lo = a_lo0 + uint32_t(s) * (TC_A_BYTES >> 4);
                if (!SPLIT) {
                    const uint32_t acc0 = (started >> t) & 1u;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) // 4 x K=16 inside the 128-byte swizzle span (+32 B = +2 descriptor units)
                        umma_f16(tmem_base + uint32_t(t * ACC), desc_hi | (a_lo + 2 * kk), desc_hi | (b_lo + 2 * kk), idesc,
                                 acc0 | uint32_t(kk != 0));
                } else {
                    // split i of the features meets weight splits 0 .. 2 - i; only x0.w0 goes to the main accumulator
                    const int i = int((unit >> 3) & 3u);
                    for (int jj = 0; jj + i <= 2; ++jj) {
                        const bool main_term = (i | jj) == 0;
                        const uint32_t acc0 = ((main_term ? started : started_small) >> t) & 1u;
                        const uint32_t d = tmem_base + uint32_t(t * ACC + (main_term ? 0 : COUT));
                        const uint32_t bj = b_lo + uint32_t(jj) * (Cfg::CHUNK_BYTES >> 4);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_f16(d, desc_hi | (a_lo + 2 * kk), desc_hi | (bj + 2 * kk), idesc, acc0 | uint32_t(kk != 0));
                        if (main_term)
                            started |= 1u << t;
                        else
                            started_small |= 1u << t;
                    }
                }
                umma_commit(bar_empty + 8 * s); // stage reusable once these MMAs retire
                started |= 1u << t;
                if (++sl == SP) { // (counters, not own % SP: a division on the issuing thread is time every unit waits for)
                    sl = 0;
                    ph ^= 1u;
                }
            }
            umma_commit(bar_accum);
        }
        __syncwarp();
    } else if (warp == WARP_B) {
        // ================= weight-chunk loader (one thread) =================
        if (lane == 0) {
            int c = 0, prev_kj = -1;
            for (int u = 0; u < nunits; ++u) {
                const uint32_t unit = units[u];
                const int kj = int(unit >> 5);
                if (kj == prev_kj)
                    continue;
                prev_kj = kj;
                const int b = c % BST;
                mbar_wait(bar_bempty + 8 * b, ((c / BST) & 1) ^ 1);
                mbar_expect_tx(bar_bfull + 8 * b, Cfg::B_BYTES);
                bulk_g2s(smem_b + b * Cfg::B_BYTES, w_img + int64_t((kj >> 2) * KB + (kj & 3)) * Cfg::B_BYTES, Cfg::B_BYTES,
                         bar_bfull + 8 * b);
                ++c;
            }
        }
        __syncwarp();
    } else if constexpr (RING) {
        // ================= kernel-map streamer (whole warp): 128 map entries = 32 lanes x 16 B per unit =================
        const int32_t *lane_nbr = nbr + tile0 * TC_TILE_M + lane * 4;
        for (int u = 0; u < nunits; ++u) {
            const uint32_t unit = units[u];
            const int g = int(unit >> 7), t = unit & 7;
            const int e = u & (TC_IDX_RING - 1);
            mbar_wait(bar_iempty + 8 * e, ((u / TC_IDX_RING) & 1) ^ 1);
#pragma unroll
            for (int sub = 0; sub < G; ++sub)
                if (g * G + sub < k3)
                    cp_async16(smem_idx + e * Cfg::RING_BYTES + sub * Cfg::SUB_STRIDE + lane * 16, lane_nbr + int64_t(g * G + sub) * pitch + t * TC_TILE_M, 16u);
            cp_async_arrive_noinc(bar_ifull + 8 * e);
        }
        cp_async_wait_all();
    }

    // ================= epilogue: warp w < 4 drains TMEM lanes 32 w .. 32 w + 31 of every tile =================
    if (warp < 4) {
        mbar_wait(bar_accum, 0);
        tc_fence_after();
        const bool bf16 = is_bf16 != 0;
        const int quarter = warp;
        const uint32_t live = uint32_t(s_live);
        constexpr int EC = COUT >= 32 ? 32 : 16; // columns drained per tcgen05.ld
        constexpr int NCH = COUT / EC;
        const bool fancy = epi.scale != nullptr || epi.shift != nullptr || epi.residual != nullptr || epi.relu != 0 || epi.stats != nullptr;
        float st_sum[NCH], st_sq[NCH]; // lane l: column c0 + (l & (EC-1)) of chunk c0 (Epilogue::stats)
#pragma unroll
        for (int c = 0; c < NCH; ++c)
            st_sum[c] = st_sq[c] = 0.f;
        // narrow outputs keep per-thread column sums over all of the CTA's tiles and cross the warp ONCE at the end (the
        // shuffle reduction per tile cost ~13 % of the 32-channel kernel); wide outputs reduce per tile (registers)
        constexpr bool DEFER = COUT <= 32;
        float df_sum[DEFER ? COUT : 1], df_sq[DEFER ? COUT : 1];
#pragma unroll
        for (int z = 0; z < (DEFER ? COUT : 1); ++z)
            df_sum[z] = df_sq[z] = 0.f;
        for (int tt = 0; tt < ntiles; ++tt) {
            const int64_t row = (tile0 + tt) * TC_TILE_M + quarter * 32 + lane;
            const bool row_ok = row < n_out;
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int c0 = ch * EC;
                uint32_t acc[32];
                if ((live >> tt) & 1u) {
                    const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(tt * ACC + c0);
                    if (EC == 32)
                        tmem_ld_32x32b_x32(taddr, acc);
                    else
                        tmem_ld_32x32b_x16(taddr, acc);
                    if (SPLIT) { // + the small-term accumulator (one rounded fp32 add per output)
                        uint32_t small[32];
                        if (EC == 32)
                            tmem_ld_32x32b_x32(taddr + COUT, small);
                        else
                            tmem_ld_32x32b_x16(taddr + COUT, small);
                        tmem_ld_wait();
#pragma unroll
                        for (int z = 0; z < EC; ++z)
                            acc[z] = __float_as_uint(__uint_as_float(acc[z]) + __uint_as_float(small[z]));
                    } else {
                        tmem_ld_wait();
                    }
                } else { // a tile no tap reaches was never accumulated: its rows are zero (+ bias)
#pragma unroll
                    for (int z = 0; z < 32; ++z)
                        acc[z] = 0u;
                }
                if (!fancy) { // bias only (the plain ConvolutionPlan / SparseConv3d call)
                    if (row_ok) {
                        if (SPLIT) {
                            const float *bias = reinterpret_cast<const float *>(epi.bias);
                            uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<float *>(y_) + row * COUT + c0);
#pragma unroll
                            for (int v4 = 0; v4 < EC / 4; ++v4) {
                                uint32_t p[4];
#pragma unroll
                                for (int h = 0; h < 4; ++h)
                                    p[h] = bias ? __float_as_uint(__uint_as_float(acc[v4 * 4 + h]) + __ldg(bias + c0 + v4 * 4 + h)) : acc[v4 * 4 + h];
                                dst[v4] = make_uint4(p[0], p[1], p[2], p[3]);
                            }
                        } else {
                            const uint16_t *bias = reinterpret_cast<const uint16_t *>(epi.bias);
                            uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(y_) + row * COUT + c0);
#pragma unroll
                            for (int v4 = 0; v4 < EC / 8; ++v4) {
                                uint32_t p[4];
#pragma unroll
                                for (int h = 0; h < 4; ++h) {
                                    float a = __uint_as_float(acc[v4 * 8 + 2 * h]), b = __uint_as_float(acc[v4 * 8 + 2 * h + 1]);
                                    if (bias) {
                                        a += half_to_float(bias[c0 + v4 * 8 + 2 * h], bf16);
                                        b += half_to_float(bias[c0 + v4 * 8 + 2 * h + 1], bf16);
                                    }
                                    p[h] = pack_half2(a, b, bf16);
                                }
                                dst[v4] = make_uint4(p[0], p[1], p[2], p[3]);
                            }
                        }
                    }
                } else {
                    // fused block epilogue: stored = act(((acc + bias) * scale + shift) + residual), statistics of the stored values
                    float v[EC];
#pragma unroll
                    for (int z = 0; z < EC; ++z) {
                        float a = __uint_as_float(acc[z]);
                        if (epi.bias)
                            a += SPLIT ? __ldg(reinterpret_cast<const float *>(epi.bias) + c0 + z)
                                       : half_to_float(__ldg(reinterpret_cast<const uint16_t *>(epi.bias) + c0 + z), bf16);
                        if (epi.scale)
                            a *= __ldg(epi.scale + c0 + z);
                        if (epi.shift)
                            a += __ldg(epi.shift + c0 + z);
                        v[z] = a;
                    }
                    if (epi.relu & 1) { // the block's activation (before a skip connection joins)
#pragma unroll
                        for (int z = 0; z < EC; ++z)
                            v[z] = fmaxf(v[z], 0.f);
                    }
                    if (epi.residual && row_ok) {
                        if (SPLIT) {
                            const float4 *res = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(epi.residual) + row * COUT + c0);
#pragma unroll
                            for (int v4 = 0; v4 < EC / 4; ++v4) {
                                const float4 r = __ldg(res + v4);
                                v[4 * v4] += r.x, v[4 * v4 + 1] += r.y, v[4 * v4 + 2] += r.z, v[4 * v4 + 3] += r.w;
                            }
                        } else {
                            const uint4 *res = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint16_t *>(epi.residual) + row * COUT + c0);
#pragma unroll
                            for (int v8 = 0; v8 < EC / 8; ++v8) {
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
                    }
                    if (epi.relu & 2) { // activation after the skip connection (fvdb/nn/simple_unet.py:187-188)
#pragma unroll
                        for (int z = 0; z < EC; ++z)
                            v[z] = fmaxf(v[z], 0.f);
                    }
                    if (SPLIT) {
                        if (row_ok) {
                            uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<float *>(y_) + row * COUT + c0);
#pragma unroll
                            for (int v4 = 0; v4 < EC / 4; ++v4)
                                dst[v4] = make_uint4(__float_as_uint(v[4 * v4]), __float_as_uint(v[4 * v4 + 1]), __float_as_uint(v[4 * v4 + 2]),
                                                     __float_as_uint(v[4 * v4 + 3]));
                        }
                    } else {
                        uint32_t p[EC / 2];
#pragma unroll
                        for (int h = 0; h < EC / 2; ++h) {
                            p[h] = pack_half2(v[2 * h], v[2 * h + 1], bf16);
                            unpack_half2(p[h], bf16, v[2 * h], v[2 * h + 1]); // statistics see what a later pass over y would read
                        }
                        if (row_ok) {
                            uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(y_) + row * COUT + c0);
#pragma unroll
                            for (int v4 = 0; v4 < EC / 8; ++v4)
                                dst[v4] = make_uint4(p[4 * v4], p[4 * v4 + 1], p[4 * v4 + 2], p[4 * v4 + 3]);
                        }
                    }
                    if (epi.stats) {
                        if constexpr (DEFER) {
                            if (row_ok) {
#pragma unroll
                                for (int z = 0; z < EC; ++z) {
                                    df_sum[c0 + z] += v[z];
                                    df_sq[c0 + z] = fmaf(v[z], v[z], df_sq[c0 + z]);
                                }
                            }
                        } else {
                            float sq[EC];
#pragma unroll
                            for (int z = 0; z < EC; ++z) {
                                v[z] = row_ok ? v[z] : 0.f;
                                sq[z] = v[z] * v[z];
                            }
                            st_sum[ch] += warp_column_sum<EC>(v, lane);
                            st_sq[ch] += warp_column_sum<EC>(sq, lane);
                        }
                    }
                }
            }
        }
        if (epi.stats) { // every MMA has retired (bar_accum), so the gather stages are free: [quarter][2][COUT] floats
            if constexpr (DEFER) {
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    float a[EC], b[EC];
#pragma unroll
                    for (int z = 0; z < EC; ++z)
                        a[z] = df_sum[ch * EC + z], b[z] = df_sq[ch * EC + z];
                    st_sum[ch] = warp_column_sum<EC>(a, lane);
                    st_sq[ch] = warp_column_sum<EC>(b, lane);
                }
            }
            float *s_stats = reinterpret_cast<float *>(smem_gen);
            if (lane < EC) {
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    s_stats[(quarter * 2 + 0) * COUT + ch * EC + lane] = st_sum[ch];
                    s_stats[(quarter * 2 + 1) * COUT + ch * EC + lane] = st_sq[ch];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (epi.stats) { // fixed-order sum of the four lane quarters -> this CTA's partial [2][COUT]
        const float *s_stats = reinterpret_cast<const float *>(smem_gen);
        for (int e = threadIdx.x; e < 2 * COUT; e += THREADS)
            epi.stats[int64_t(blockIdx.x) * 2 * COUT + e] = (s_stats[e] + s_stats[2 * COUT + e]) + (s_stats[4 * COUT + e] + s_stats[6 * COUT + e]);
    }
    if (warp == WARP_MMA)
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---- host side ------------------------------------------------------------------------------------
static inline int tc_groups(int64_t k3, int32_t cin) { return cin >= 64 ? int(k3) : int(ceil_div(k3, 64 / cin)); }
static inline size_t tc_image_bytes(int32_t cin, int32_t cout, int64_t k3, int splits) {
    const int64_t kb = cin >= 64 ? cin / 64 : 1;
    return align_up(size_t(tc_groups(k3, cin)) * size_t(kb) * size_t(splits) * size_t(cout) * 128, 256);
}

size_t tc_weight_image_bytes(int32_t cin, int32_t cout, int64_t k3, int32_t dtype) { return tc_image_bytes(cin, cout, k3, dtype == FVC_F32 ? 3 : 1); }

static int launch_weight_image(const WeightSource &ws, int k3, int cin, int cout, int32_t dtype, void *image, cudaStream_t stream) {
    const int ns = dtype == FVC_F32 ? 3 : 1;
    const int64_t chunks16 = int64_t(tc_groups(k3, cin)) * (cin >= 64 ? cin / 64 : 1) * ns * cout * 8;
    if (chunks16 == 0)
        return FVC_OK;
    tc_weight_image_kernel<<<int(ceil_div(chunks16, 256) > 1184 ? 1184 : ceil_div(chunks16, 256)), 256, 0, stream>>>(
        ws, k3, cin, cout, ns, (dtype == FVC_BF16 || dtype == FVC_F32) ? 1 : 0, reinterpret_cast<uint4 *>(image));
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

// cin / cout are the PUBLIC weight dimensions [Cout, Cin, k0, k1, k2]; transpose != 0 prepares W[k]^T (dgrad: the executor's
// reduction runs over the public Cout and its outputs are the public Cin)
int tc_prepare_weights(const void *weights, const int64_t strides[5], int32_t dtype_in, int32_t cout, int32_t cin, int32_t k0, int32_t k1,
                       int32_t k2, int32_t transpose, int32_t flip_taps, int32_t dtype, void *image, cudaStream_t stream) {
    WeightSource ws;
    ws.w = weights;
    ws.s_n = transpose ? strides[1] : strides[0];
    ws.s_k = transpose ? strides[0] : strides[1];
    ws.s_t0 = strides[2], ws.s_t1 = strides[3], ws.s_t2 = strides[4];
    ws.k1 = k1, ws.k2 = k2, ws.k3 = k0 * k1 * k2, ws.flip = flip_taps ? 1 : 0, ws.dtype_in = dtype_in;
    return launch_weight_image(ws, ws.k3, transpose ? cout : cin, transpose ? cin : cout, dtype, image, stream);
}

// image from the packed [K][Cin][Cout] array in the working dtype (the fvc_conv_forward entry point without prepared weights)
static int image_from_packed(const ConvArgs &a, void *image) {
    WeightSource ws;
    ws.w = a.w;
    ws.s_n = 1, ws.s_k = a.cout, ws.s_t0 = int64_t(a.cin) * a.cout, ws.s_t1 = 0, ws.s_t2 = 0;
    ws.k1 = 1, ws.k2 = 1, ws.k3 = a.k3, ws.flip = 0, ws.dtype_in = a.dtype;
    return launch_weight_image(ws, a.k3, a.cin, a.cout, a.dtype, image, a.stream);
}

template <int CIN, int COUT, int TILES, int STAGES, int BST, int PW, bool SPLIT = false, int MODE = 0>
static int launch_tc_fwd(const ConvArgs &a, const void *x, const uint8_t *w_img) {
    using Cfg = TcFwdCfg<CIN, COUT, TILES, STAGES, BST, PW, SPLIT, MODE>;
    auto kernel = conv_tc_fwd_kernel<CIN, COUT, TILES, STAGES, BST, PW, SPLIT, MODE>;
    static std::atomic<unsigned long long> configured{0}; // per instantiation, one bit per device
    const int rc = ensure_dynamic_smem(kernel, Cfg::SMEM, configured);
    if (rc)
        return rc;
    FVC_REQUIRE(ceil_div(a.k3, Cfg::G) * Cfg::KB * Cfg::NS * TILES <= Cfg::MAX_UNITS, FVC_ERR_UNSUPPORTED,
                "kernel volume %d too large for the tensor-core unit list", a.k3);
    const int64_t tiles = ceil_div(a.n_out, TC_TILE_M);
    const unsigned grid = unsigned(ceil_div(tiles, TILES));
    const bool bf16 = SPLIT || a.dtype == FVC_BF16;
    const uint32_t idesc = make_idesc_f16(TC_TILE_M, COUT, bf16, false, false);
    kernel<<<grid, Cfg::THREADS, Cfg::SMEM, a.stream>>>(reinterpret_cast<const uint16_t *>(x), w_img, a.epi, a.y, a.nbr, a.pitch,
                                                        reinterpret_cast<const unsigned long long *>(a.tile_mask), a.n_in, a.n_out, a.k3,
                                                        idesc, bf16 ? 1 : 0);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

static inline bool tc_channels_ok(int32_t c, int32_t max_c) { return (c == 16 || c == 32 || c == 64 || c == 128 || c == 256) && c <= max_c; }

// ---- kernel shapes: ONE table (tiles per CTA / stages / weight slots / producer warps / mode) per (Cin class, Cout, dtype) --------
// Cin >= 64 runs MODE 1 (warp-per-unit producers); packed taps (Cin = 16 / 32) run MODE 0.
struct TcShape {
    int tiles, stages, bst, pw, mode;
};
// Small fp32 batches (BASELINE.json configs[0]: 100 k voxels = 782 tiles) do not fill the machine with the multi-tile CTAs of
// the table -- 196 four-tile CTAs walk 168 units each, one after the other -- so they run 1-tile CTAs, three per SM.
static inline bool tc_small_batch(int64_t n_out, int32_t cin, int32_t cout, bool split) {
    return split && cin < 64 && cout <= 64 && ceil_div(n_out > 0 ? n_out : 1, TC_TILE_M) <= 4096;
}
static inline TcShape tc_shape(int32_t cin, int32_t cout, bool split, bool small = false) {
    const bool wide = cin >= 64;
    if (small)
        return TcShape{1, 2, 2, 4, 0};
    if (split) {
        switch (cout) {
        case 16: return wide ? TcShape{8, 4, 3, 4, 1} : TcShape{8, 4, 3, 4, 0};
        case 32: return wide ? TcShape{4, 4, 3, 4, 1} : TcShape{4, 4, 3, 4, 0};
        case 64: return wide ? TcShape{2, 3, 2, 3, 1} : TcShape{2, 3, 2, 4, 0};
        default: return wide ? TcShape{1, 4, 2, 4, 1} : TcShape{1, 3, 2, 4, 0};
        }
    }
    switch (cout) {
    case 16: return wide ? TcShape{8, 4, 2, 4, 1} : TcShape{8, 3, 2, 4, 0};
    case 32: return wide ? TcShape{4, 4, 2, 4, 1} : TcShape{4, 3, 2, 4, 0};
    case 64: return wide ? TcShape{2, 2, 2, 2, 1} : TcShape{2, 3, 2, 4, 0};
    // 128- and 256-wide outputs: two pipelines (a tile, two producer warps and an MMA warp each) share the weight chunks, which
    // are 48 % / 64 % of these shapes' L2 -> SM traffic (128 -> 128: 0.817 -> 0.757 ms, 256 -> 256: 2.67 -> 2.40 ms on the C2 batch)
    case 128: return wide ? TcShape{2, 4, 2, 4, 3} : TcShape{1, 2, 2, 4, 0};
    default: return wide ? TcShape{2, 4, 2, 4, 3} : TcShape{2, 6, 3, 4, 0};
    }
}

bool tc_forward_supported(int32_t cin, int32_t cout, int64_t k3, int32_t dtype) {
    const bool split = dtype == FVC_F32;
    if (dtype != FVC_F16 && dtype != FVC_BF16 && !split)
        return false;
    if (!tc_channels_ok(cin, 256) || !tc_channels_ok(cout, split ? 128 : 256))
        return false;
    const int64_t kb = cin >= 64 ? cin / 64 : 1;
    const int64_t units = tc_groups(k3, cin) * kb * (split ? 3 : 1) * 8;
    return k3 >= 1 && k3 <= 64 * TC_MASK_WORDS && units <= (split ? TC_MAX_UNITS_SPLIT : TC_MAX_UNITS);
}

// Variant 12 (fvc_set_tuning(0, 12)) serves narrow half-precision layers with the tensor-memory executor (conv_tc_ts.cu).  Measured
// (profiles/r02_ts_executor.md): it cuts the L2 -> SM traffic to the compulsory bytes but, at ~230 dependent instructions per
// (warp, unit) with 16 gather warps per SM, only ties the shared-memory kernels (faster on 16 -> 32, slower on 32 -> 32 and on
// BASELINE.json configs[4]), so those stay the default.
static inline bool takes_ts(int32_t cin, int32_t cout, int64_t k3, int32_t dtype) { return g_tc_variant == 12 && tc_ts_supported(cin, cout, k3, dtype); }

int64_t tc_stats_blocks(int64_t n_out, int32_t cin, int32_t cout, int64_t k3, int32_t dtype, int32_t *rows_per_block) {
    if (takes_ts(cin, cout, k3, dtype))
        return tc_ts_stats_blocks(n_out, rows_per_block);
    const int tiles = tc_shape(cin, cout, dtype == FVC_F32, tc_small_batch(n_out, cin, cout, dtype == FVC_F32)).tiles;
    if (rows_per_block)
        *rows_per_block = tiles * TC_TILE_M;
    return ceil_div(ceil_div(n_out > 0 ? n_out : 0, TC_TILE_M), tiles);
}

// scratch = [weight image (unless prepared) | split feature rows (fp32, unless the caller passes split rows)]
size_t tc_forward_scratch_bytes(int64_t n_in, int64_t, int32_t cin, int32_t cout, int64_t k3, int32_t dtype) {
    if (dtype == FVC_F32)
        return tc_image_bytes(cin, cout, k3, 3) + align_up(size_t(n_in > 0 ? n_in : 0) * 3 * size_t(cin) * 2, 256) + 256;
    return tc_image_bytes(cin, cout, k3, 1) + 256;
}

int tc_split_rows(const float *x, int64_t n, int c, uint16_t *xs, cudaStream_t stream) {
    const int64_t work = n * (c / 8);
    if (work <= 0)
        return FVC_OK;
    tc_split_rows_kernel<<<int(ceil_div(work, 256) > 148 * 16 ? 148 * 16 : ceil_div(work, 256)), 256, 0, stream>>>(x, n, c, xs);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

// One instantiation per row of tc_shape(); the launch checks that table and macro agree.
#define FVC_TC_LAUNCH(CI, CO, T, S, B, P, SPLIT_, M)                                                                                     \
    if (a.cin == CI && a.cout == CO) {                                                                                                   \
        const TcShape sh = tc_shape(CI, CO, SPLIT_);                                                                                     \
        FVC_REQUIRE(sh.tiles == T && sh.stages == S && sh.bst == B && sh.pw == P && sh.mode == M, FVC_ERR_RUNTIME,                      \
                    "tensor-core shape table / instantiation mismatch for %d -> %d", CI, CO);                                            \
        return launch_tc_fwd<CI, CO, T, S, B, P, SPLIT_, M>(a, x, img);                                                                  \
    }

static int tc_forward_split(const ConvArgs &a, const void *x, const uint8_t *img) {
    if (tc_small_batch(a.n_out, a.cin, a.cout, true)) {
#define FVC_TCS_SMALL(CI, CO)          \
    if (a.cin == CI && a.cout == CO) \
        return launch_tc_fwd<CI, CO, 1, 2, 2, 4, true, 0>(a, x, img);
        FVC_TCS_SMALL(16, 16)
        FVC_TCS_SMALL(16, 32)
        FVC_TCS_SMALL(16, 64)
        FVC_TCS_SMALL(32, 16)
        FVC_TCS_SMALL(32, 32)
        FVC_TCS_SMALL(32, 64)
#undef FVC_TCS_SMALL
    }
#define FVC_TCS_NARROW(CI)                    \
    FVC_TC_LAUNCH(CI, 16, 8, 4, 3, 4, true, 0) \
    FVC_TC_LAUNCH(CI, 32, 4, 4, 3, 4, true, 0) \
    FVC_TC_LAUNCH(CI, 64, 2, 3, 2, 4, true, 0) \
    FVC_TC_LAUNCH(CI, 128, 1, 3, 2, 4, true, 0)
#define FVC_TCS_WIDE(CI)                      \
    FVC_TC_LAUNCH(CI, 16, 8, 4, 3, 4, true, 1) \
    FVC_TC_LAUNCH(CI, 32, 4, 4, 3, 4, true, 1) \
    FVC_TC_LAUNCH(CI, 64, 2, 3, 2, 3, true, 1) \
    FVC_TC_LAUNCH(CI, 128, 1, 4, 2, 4, true, 1)
    FVC_TCS_NARROW(16)
    FVC_TCS_NARROW(32)
    FVC_TCS_WIDE(64)
    FVC_TCS_WIDE(128)
    FVC_TCS_WIDE(256)
#undef FVC_TCS_NARROW
#undef FVC_TCS_WIDE
    return set_error(FVC_ERR_UNSUPPORTED, "no fp32 tensor-core kernel for channels %d -> %d", a.cin, a.cout);
}

static int tc_forward_half(const ConvArgs &a, const void *x, const uint8_t *img) {
    if (takes_ts(a.cin, a.cout, a.k3, a.dtype))
        return tc_ts_forward(a, x, img);
    if (g_tc_variant == 15 && a.cin == 256 && a.cout == 256 && !a.epi.stats)
        return launch_tc_fwd<256, 256, 2, 6, 3, 4, false, 0>(a, x, img); // the round-1 shape of the 256-wide outputs (A/B baseline)
    // experiment knob (fvc_set_tuning(0, v), scripts/bench_variants.py): alternative pipeline shapes of the two headline shapes
    if (g_tc_variant != 0 && a.cin == a.cout && (a.cin == 64 || a.cin == 128) && !a.epi.stats) { // (statistics blocks follow the default shape)
        const bool c64 = a.cin == 64;
        switch (g_tc_variant) {
        case 2: return c64 ? launch_tc_fwd<64, 64, 4, 4, 2, 4, false, 1>(a, x, img) : launch_tc_fwd<128, 128, 2, 4, 2, 4, false, 1>(a, x, img); // 2 CTAs / SM
        case 3: return c64 ? launch_tc_fwd<64, 64, 8, 8, 3, 4, false, 1>(a, x, img) : launch_tc_fwd<128, 128, 4, 8, 2, 4, false, 1>(a, x, img); // 1 CTA / SM
        case 4: return c64 ? launch_tc_fwd<64, 64, 2, 4, 2, 4, false, 1>(a, x, img) : launch_tc_fwd<128, 128, 1, 4, 2, 4, false, 1>(a, x, img);
        case 5: return c64 ? launch_tc_fwd<64, 64, 2, 4, 2, 2, false, 1>(a, x, img) : launch_tc_fwd<128, 128, 1, 4, 2, 2, false, 1>(a, x, img);
        case 6: return c64 ? launch_tc_fwd<64, 64, 4, 5, 2, 3, false, 1>(a, x, img) : launch_tc_fwd<128, 128, 2, 4, 2, 3, false, 1>(a, x, img);
        case 7: return c64 ? launch_tc_fwd<64, 64, 2, 3, 2, 3, false, 1>(a, x, img) : launch_tc_fwd<128, 128, 1, 2, 1, 2, false, 1>(a, x, img); // 3 CTAs / SM (64), 4 (128)
        case 8: return c64 ? launch_tc_fwd<64, 64, 1, 2, 2, 2, false, 1>(a, x, img) : launch_tc_fwd<128, 128, 2, 3, 2, 2, false, 1>(a, x, img);
        case 9: return c64 ? launch_tc_fwd<64, 64, 2, 3, 2, 2, false, 1>(a, x, img) : launch_tc_fwd<128, 128, 4, 6, 2, 3, false, 1>(a, x, img);
        case 13: return c64 ? launch_tc_fwd<64, 64, 4, 4, 3, 4, false, 3>(a, x, img) : launch_tc_fwd<128, 128, 2, 4, 2, 3, false, 1>(a, x, img); // 64: 2 pipelines, 2 CTAs / SM; 128: the single-pipeline shape
        case 14: return c64 ? launch_tc_fwd<64, 64, 8, 8, 3, 8, false, 3>(a, x, img) : launch_tc_fwd<128, 128, 4, 8, 3, 8, false, 3>(a, x, img); // 4 pipelines, 1 CTA / SM
        case 10: return c64 ? launch_tc_fwd<64, 64, 4, 3, 2, 3, false, 1>(a, x, img) : launch_tc_fwd<128, 128, 2, 3, 2, 3, false, 1>(a, x, img);
        default: break;
        }
    }
#define FVC_TC_LEGACY(CI, CO, T, S, B) \
    if (a.cin == CI && a.cout == CO)  \
        return launch_tc_fwd<CI, CO, T, S, B, 4, false, 0>(a, x, img);
#define FVC_TC_LEGACY_CIN(CI)     \
    FVC_TC_LEGACY(CI, 16, 8, 3, 2)  \
    FVC_TC_LEGACY(CI, 32, 4, 3, 2)  \
    FVC_TC_LEGACY(CI, 64, 2, 3, 2)  \
    FVC_TC_LEGACY(CI, 128, 1, 2, 2) \
    FVC_TC_LEGACY(CI, 256, 2, 6, 3)
    if (g_tc_variant == 1 && !a.epi.stats) { // the round-1 kernel (four producer warps share every unit, map ring) for every wide shape: A/B baseline
        FVC_TC_LEGACY_CIN(64)
        FVC_TC_LEGACY_CIN(128)
        FVC_TC_LEGACY_CIN(256)
    }
#undef FVC_TC_LEGACY_CIN
#undef FVC_TC_LEGACY
#define FVC_TC_NARROW(CI)                      \
    FVC_TC_LAUNCH(CI, 16, 8, 3, 2, 4, false, 0) \
    FVC_TC_LAUNCH(CI, 32, 4, 3, 2, 4, false, 0) \
    FVC_TC_LAUNCH(CI, 64, 2, 3, 2, 4, false, 0) \
    FVC_TC_LAUNCH(CI, 128, 1, 2, 2, 4, false, 0) \
    FVC_TC_LAUNCH(CI, 256, 2, 6, 3, 4, false, 0)
#define FVC_TC_WIDE(CI)                        \
    FVC_TC_LAUNCH(CI, 16, 8, 4, 2, 4, false, 1) \
    FVC_TC_LAUNCH(CI, 32, 4, 4, 2, 4, false, 1) \
    FVC_TC_LAUNCH(CI, 64, 2, 2, 2, 2, false, 1) \
    FVC_TC_LAUNCH(CI, 128, 2, 4, 2, 4, false, 3) \
    FVC_TC_LAUNCH(CI, 256, 2, 4, 2, 4, false, 3)
    FVC_TC_NARROW(16)
    FVC_TC_NARROW(32)
    FVC_TC_WIDE(64)
    FVC_TC_WIDE(128)
    FVC_TC_WIDE(256)
#undef FVC_TC_NARROW
#undef FVC_TC_WIDE
    return set_error(FVC_ERR_UNSUPPORTED, "no tensor-core kernel for channels %d -> %d", a.cin, a.cout);
}

int tc_forward(const ConvArgs &a) {
    const bool split = a.dtype == FVC_F32;
    const size_t img_bytes = a.w_prepared ? 0 : tc_image_bytes(a.cin, a.cout, a.k3, split ? 3 : 1);
    const size_t rows_bytes = (split && !a.x_split) ? align_up(size_t(a.n_in > 0 ? a.n_in : 0) * 3 * size_t(a.cin) * 2, 256) : 0;
    const size_t need = img_bytes + rows_bytes;
    FVC_REQUIRE(need == 0 || (a.scratch && a.scratch_bytes >= need), FVC_ERR_RUNTIME, "tensor-core conv scratch too small: %zu < %zu",
                a.scratch_bytes, need);
    FVC_REQUIRE((reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a.scratch) & 255) == 0 && (reinterpret_cast<uintptr_t>(a.nbr) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a.w) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.epi.residual) & 15) == 0,
                FVC_ERR_RUNTIME, "tensor-core conv needs 16-byte aligned feature / output / residual / map pointers and 256-byte aligned scratch");
    FVC_REQUIRE(a.pitch % 4 == 0 && a.pitch >= ceil_div(a.n_out, TC_TILE_M) * TC_TILE_M, FVC_ERR_RUNTIME,
                "tensor-core conv needs the map pitch (%lld) to be a multiple of 4 covering whole 128-row tiles", (long long)a.pitch);
    const uint8_t *img = reinterpret_cast<const uint8_t *>(a.w);
    if (!a.w_prepared) {
        img = reinterpret_cast<const uint8_t *>(a.scratch);
        const int rc = image_from_packed(a, a.scratch);
        if (rc)
            return rc;
    }
    const void *x = a.x;
    if (split && !a.x_split) {
        uint16_t *xs = reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(a.scratch) + img_bytes);
        const int rc = tc_split_rows(reinterpret_cast<const float *>(a.x), a.n_in, a.cin, xs, a.stream);
        if (rc)
            return rc;
        x = xs;
    }
    return split ? tc_forward_split(a, x, img) : tc_forward_half(a, x, img);
}

// tensor-core weight gradient: conv_tc_wgrad.cu

} // namespace fvc
