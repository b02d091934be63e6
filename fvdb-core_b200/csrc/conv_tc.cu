// conv_tc.cu -- tcgen05 / TMEM implicit-GEMM sparse convolution for f16 / bf16 (fp32 accumulate).
//
// Replaces PredGatherIGemm.cu (SM80 mma.sync TF32, forward only, one CTA per leaf) and the cuBLAS
// gather -> mm -> atomic-scatter pipeline of GatherScatterDefault.cu:706-721,786-808 for the half types.
//
// Forward / dgrad kernel (output-stationary, no atomics):
//   a CTA owns TILES consecutive 128-row output tiles; their fp32 accumulators (128 lanes x COUT columns
//   each) stay in TMEM for the whole tap loop.  For every (tap, 64-channel block) the weight chunk
//   B = W[tap][:, block] (COUT x 64, K-major, 128B-swizzled image prepared by tc_pack_b_kernel) arrives by
//   ONE cp.async.bulk; then for each tile the 4 producer warps gather the 128 neighbour rows (128 B each)
//   with 16-byte zero-filling cp.async straight into the canonical K-major SWIZZLE_128B layout, and one
//   elected thread issues 4 x tcgen05.mma (M=128, N=COUT, K=16) accumulating into that tile's TMEM slice.
//   Stages are recycled by tcgen05.commit -> mbarrier.  Epilogue: tcgen05.ld -> (+bias) -> bf16/f16 -> one
//   plain 128-bit store stream per output row.
//
// Warp roles (352 threads): warps 0-7 gather producers, then epilogue (warp w owns TMEM lanes 32*(w&3)..+31 of the
// tiles with parity w>>2); warp 8 TMEM allocator + MMA issuer; warp 9 weight-chunk loader; warp 10 streams the
// kernel-map entries of upcoming units into a 16-deep shared-memory ring (one 512-byte cp.async per unit), so the
// producers never wait on an index load (ncu showed that wait as the top stall of the first version; the second
// showed the producers' own instruction stream, hence 8 producer warps and a bit-scan unit iterator).
#include "conv_internal.cuh"
#include "tc_ptx.cuh"

#include <cstdlib>

namespace fvc {

using namespace tc;

constexpr int TC_PRODUCER_WARPS = 8;
constexpr int TC_PRODUCERS = TC_PRODUCER_WARPS * 32;
constexpr int TC_WARP_MMA = TC_PRODUCER_WARPS, TC_WARP_B = TC_PRODUCER_WARPS + 1;
constexpr int TC_THREADS = (TC_PRODUCER_WARPS + 3) * 32;
constexpr int TC_IDX_RING = 16;             // map-entry ring depth (units of 128 int32)
constexpr int TC_TILE_M = 128;
constexpr int TC_A_BYTES = TC_TILE_M * 128; // one stage: 128 rows x 64 channels x 2 B
constexpr int TC_MASK_WORDS = 8;            // tile tap-mask words kept in shared memory (K^3 <= 512)

constexpr int tmem_cols_for(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

template <int CIN, int COUT, int TILES, int STAGES, int BST> struct TcFwdCfg {
    static constexpr int KB = CIN / 64;                      // 64-channel reduction blocks per tap
    static constexpr int B_BYTES = COUT * 128;               // one weight chunk: COUT rows x 64 channels x 2 B
    static constexpr int TMEM_COLS = tmem_cols_for(TILES * COUT);
    static constexpr int NUM_BARS = 2 * STAGES + 2 * BST + 1 + 2 * TC_IDX_RING;
    static constexpr size_t SMEM = 1024 + size_t(STAGES) * TC_A_BYTES + size_t(BST) * B_BYTES + size_t(TC_IDX_RING) * 512 + 8 * NUM_BARS + 16;
    // co-resident CTAs: limited by TMEM columns (512 per SM) and shared memory (227 KB per SM, 1 KB reserved per CTA)
    static constexpr int CTAS_PER_SM = (512 / TMEM_COLS) < int(232448 / (SMEM + 1024)) ? (512 / TMEM_COLS) : int(232448 / (SMEM + 1024));
    static_assert(CTAS_PER_SM >= 1, "configuration does not fit one SM");
    static_assert(CIN % 64 == 0 && COUT % 16 == 0 && COUT >= 16 && COUT <= 256, "unsupported channel counts");
    static_assert(TILES * COUT <= 512, "accumulators exceed TMEM");
};

// Weight image: for chunk c = tap * KB + j, COUT rows of 128 B; row n holds channels [64j, 64j+64) of
// W[tap][:, n] with 16-byte chunk q stored at position q ^ (n & 7)  (the SWIZZLE_128B K-major atom).
__global__ void tc_pack_b_kernel(const uint16_t *__restrict__ w /*[k3][cin][cout]*/, int k3, int cin, int cout,
                                 uint4 *__restrict__ img) {
    const int kb = cin / 64;
    const int64_t total = int64_t(k3) * kb * cout * 8; // 16-byte chunks
    for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
        const int pos = int(e & 7);
        const int n = int((e >> 3) % cout);
        const int64_t chunk = (e >> 3) / cout;
        const int j = int(chunk % kb), tap = int(chunk / kb);
        const int q = pos ^ (n & 7);
        const uint16_t *src = w + (int64_t(tap) * cin + j * 64 + q * 8) * cout + n;
        uint32_t v[4];
#pragma unroll
        for (int h = 0; h < 4; ++h)
            v[h] = uint32_t(src[int64_t(2 * h) * cout]) | (uint32_t(src[int64_t(2 * h + 1) * cout]) << 16);
        img[e] = make_uint4(v[0], v[1], v[2], v[3]);
    }
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b, bool bf16) {
    if (bf16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    }
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ float half_to_float(uint16_t v, bool bf16) {
    return bf16 ? __bfloat162float(*reinterpret_cast<__nv_bfloat16 *>(&v)) : __half2float(*reinterpret_cast<__half *>(&v));
}

template <int CIN, int COUT, int TILES, int STAGES, int BST>
__global__ void __launch_bounds__(TC_THREADS, (TcFwdCfg<CIN, COUT, TILES, STAGES, BST>::CTAS_PER_SM > 2 ? 2 : TcFwdCfg<CIN, COUT, TILES, STAGES, BST>::CTAS_PER_SM))
conv_tc_fwd_kernel(const uint16_t *__restrict__ x, const uint8_t *__restrict__ w_img, const uint16_t *__restrict__ bias,
                   uint16_t *__restrict__ y, const int32_t *__restrict__ nbr, int64_t pitch,
                   const unsigned long long *__restrict__ tile_mask, int64_t n_out, int k3, uint32_t idesc, int is_bf16) {
    using Cfg = TcFwdCfg<CIN, COUT, TILES, STAGES, BST>;
    constexpr int KB = Cfg::KB;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u; // SWIZZLE_128B atoms need 1024-byte alignment
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_a + STAGES * TC_A_BYTES;
    const uint32_t smem_idx = smem_b + BST * Cfg::B_BYTES;
    const uint32_t bars = smem_idx + TC_IDX_RING * 512;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * STAGES;
    const uint32_t bar_bfull = bars + 16 * STAGES, bar_bempty = bar_bfull + 8 * BST;
    const uint32_t bar_accum = bar_bempty + 8 * BST;
    const uint32_t bar_ifull = bar_accum + 8, bar_iempty = bar_ifull + 8 * TC_IDX_RING;
    const uint32_t tmem_slot = bar_iempty + 8 * TC_IDX_RING;
    uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t total_tiles = (n_out + TC_TILE_M - 1) / TC_TILE_M;
    const int64_t tile0 = int64_t(blockIdx.x) * TILES;
    const int ntiles = int(total_tiles - tile0 < TILES ? total_tiles - tile0 : TILES);

    // s_tiles[k]: bit t set iff tile t of this CTA has a row with a neighbour through tap k.  Units whose bit is
    // clear are skipped by every role (about half of all units on planar scenes).
    __shared__ uint8_t s_tiles[64 * TC_MASK_WORDS];
    {
        const int words = (k3 + 63) >> 6;
        for (int k = threadIdx.x; k < k3; k += TC_THREADS) {
            uint32_t bits = 0;
            for (int t = 0; t < ntiles; ++t) {
                const unsigned long long m = tile_mask ? __ldg(tile_mask + (tile0 + t) * words + (k >> 6)) : ~0ull;
                bits |= uint32_t((m >> (k & 63)) & 1ull) << t;
            }
            s_tiles[k] = uint8_t(bits);
        }
    }
    // next active (tap, channel block, tile) unit in [tap][block][tile] order; every role walks the same sequence.
    // `bits` caches s_tiles[k]; start with k = -1.
    auto advance = [&](int &k, int &j, int &t, uint32_t &bits) {
        const uint32_t rest = k >= 0 ? bits & ~((2u << t) - 1u) : 0u;
        if (rest) {
            t = __ffs(rest) - 1;
            return;
        }
        if (k >= 0 && ++j < KB) {
            t = __ffs(bits) - 1;
            return;
        }
        j = 0;
        do {
            ++k;
        } while (k < k3 && (bits = s_tiles[k]) == 0u);
        t = k < k3 ? __ffs(bits) - 1 : 0;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, TC_PRODUCERS); // one completion-triggered arrival per producer thread
            mbar_init(bar_empty + 8 * s, 1);            // tcgen05.commit
        }
        for (int b = 0; b < BST; ++b) {
            mbar_init(bar_bfull + 8 * b, 1); // expect_tx arrival + bytes
            mbar_init(bar_bempty + 8 * b, 1);
        }
        mbar_init(bar_accum, 1);
        for (int e = 0; e < TC_IDX_RING; ++e) {
            mbar_init(bar_ifull + 8 * e, 32);            // one completion-triggered arrival per streamer lane
            mbar_init(bar_iempty + 8 * e, TC_PRODUCERS); // every producer thread has read its entries
        }
        fence_mbar_init();
    }
    if (warp == TC_WARP_MMA)
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp < TC_PRODUCER_WARPS) {
        // ================= gather producers: warp w copies rows [16w, 16w+16) of the tile, 4 rows per instruction ====
        const int q = lane & 7, rsub = lane >> 3;
        int k = -1, j = 0, t = 0;
        uint32_t bits = 0;
        advance(k, j, t, bits);
        for (int u = 0; k < k3; ++u) {
            const int e = u % TC_IDX_RING;
            mbar_wait(bar_ifull + 8 * e, (u / TC_IDX_RING) & 1);
            int idx[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(idx[i]) : "r"(smem_idx + e * 512 + (warp * 16 + 4 * i + rsub) * 4) : "memory");
            const int64_t row0 = (tile0 + t) * TC_TILE_M + warp * 16 + rsub;
            const int col = j * 64 + q * 8;
            advance(k, j, t, bits);
            const int s = u % STAGES;
            mbar_wait(bar_empty + 8 * s, ((u / STAGES) & 1) ^ 1);
            const uint32_t stage = smem_a + s * TC_A_BYTES;
#pragma unroll
            for (int i = 0; i < 4; ++i) { // 8 lanes cover one 128-byte row (one full line)
                const int row = warp * 16 + 4 * i + rsub;
                const bool ok = idx[i] >= 0 && row0 + 4 * i < n_out;
                const uint16_t *src = x + (ok ? int64_t(idx[i]) * CIN + col : 0);
                cp_async16(stage + row * 128 + ((q ^ (row & 7)) << 4), src, ok ? 16u : 0u);
            }
            mbar_arrive(bar_iempty + 8 * e); // ring entry consumed (its values are in registers)
            // completion-triggered arrival (the CUTLASS sm100 cp.async -> UMMA idiom): the producer never blocks on
            // its own loads, so up to STAGES gathers per CTA stay in flight
            cp_async_arrive_noinc(bar_full + 8 * s);
        }
        cp_async_wait_all();

        // ================= epilogue: warp w drains TMEM lanes 32*(w&3).., tiles of parity w>>2 =================
        mbar_wait(bar_accum, 0);
        tc_fence_after();
        const bool bf16 = is_bf16 != 0;
        const int quarter = warp & 3;
        for (int tt = warp >> 2; tt < ntiles; tt += TC_PRODUCER_WARPS / 4) {
            const int64_t row = (tile0 + tt) * TC_TILE_M + quarter * 32 + lane;
            bool live = false; // a tile no tap reaches was never accumulated: its rows are zero
            for (int kk = 0; kk < k3; ++kk)
                live = live || ((s_tiles[kk] >> tt) & 1u);
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += 32) {
                uint32_t acc[32];
                if (live) {
                    tmem_ld_32x32b_x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(tt * COUT + c0), acc);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int z = 0; z < 32; ++z)
                        acc[z] = 0u;
                }
                if (row < n_out) {
                    uint4 *dst = reinterpret_cast<uint4 *>(y + row * COUT + c0);
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4) {
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
        }
    } else if (warp == TC_WARP_MMA) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            int u = 0, c = 0;
            uint32_t started = 0; // tiles whose accumulator has been written at least once
            for (int k = 0; k < k3; ++k) {
                const uint32_t tiles = s_tiles[k];
                if (!tiles)
                    continue;
                for (int j = 0; j < KB; ++j, ++c) {
                    const int b = c % BST;
                    mbar_wait(bar_bfull + 8 * b, (c / BST) & 1);
                    const uint32_t b_base = smem_b + b * Cfg::B_BYTES;
                    for (uint32_t rest = tiles; rest; rest &= rest - 1u, ++u) {
                        const int t = __ffs(rest) - 1;
                        const int s = u % STAGES;
                        mbar_wait(bar_full + 8 * s, (u / STAGES) & 1);
                        tc_fence_after();
                        const uint32_t a_base = smem_a + s * TC_A_BYTES;
                        const uint32_t acc0 = (started >> t) & 1u;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) // 4 x K=16 inside the 128-byte swizzle span
                            umma_f16(tmem_base + uint32_t(t * COUT), make_smem_desc_sw128(a_base + kk * 32, 16, 1024),
                                     make_smem_desc_sw128(b_base + kk * 32, 16, 1024), idesc, (acc0 | uint32_t(kk != 0)));
                        umma_commit(bar_empty + 8 * s); // stage reusable once these MMAs retire
                        started |= 1u << t;
                    }
                    umma_commit(bar_bempty + 8 * b);
                }
            }
            umma_commit(bar_accum);
        }
        __syncwarp();
    } else if (warp == TC_WARP_B) {
        // ================= weight-chunk loader (one thread) =================
        if (lane == 0) {
            int c = 0;
            for (int k = 0; k < k3; ++k) {
                if (!s_tiles[k])
                    continue;
                for (int j = 0; j < KB; ++j, ++c) {
                    const int b = c % BST;
                    mbar_wait(bar_bempty + 8 * b, ((c / BST) & 1) ^ 1);
                    mbar_expect_tx(bar_bfull + 8 * b, Cfg::B_BYTES);
                    bulk_g2s(smem_b + b * Cfg::B_BYTES, w_img + int64_t(k * KB + j) * Cfg::B_BYTES, Cfg::B_BYTES, bar_bfull + 8 * b);
                }
            }
        }
        __syncwarp();
    } else {
        // ================= kernel-map streamer (whole warp): 128 map entries = 32 lanes x 16 B per unit =================
        int k = -1, j = 0, t = 0;
        uint32_t bits = 0;
        advance(k, j, t, bits);
        for (int u = 0; k < k3; ++u) {
            const int e = u % TC_IDX_RING;
            mbar_wait(bar_iempty + 8 * e, ((u / TC_IDX_RING) & 1) ^ 1);
            cp_async16(smem_idx + e * 512 + lane * 16, nbr + int64_t(k) * pitch + (tile0 + t) * TC_TILE_M + lane * 4, 16u);
            cp_async_arrive_noinc(bar_ifull + 8 * e);
            advance(k, j, t, bits);
        }
        cp_async_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_WARP_MMA)
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---- host side ------------------------------------------------------------------------------------
template <int CIN, int COUT, int TILES, int STAGES, int BST> static int launch_tc_fwd(const ConvArgs &a, const uint8_t *w_img) {
    using Cfg = TcFwdCfg<CIN, COUT, TILES, STAGES, BST>;
    auto kernel = conv_tc_fwd_kernel<CIN, COUT, TILES, STAGES, BST>;
    static bool configured = false; // per instantiation
    if (!configured) {
        FVC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SMEM)));
        configured = true;
    }
    const int64_t tiles = ceil_div(a.n_out, TC_TILE_M);
    const unsigned grid = unsigned(ceil_div(tiles, TILES));
    const bool bf16 = a.dtype == FVC_BF16;
    const uint32_t idesc = make_idesc_f16(TC_TILE_M, COUT, bf16, false, false);
    kernel<<<grid, TC_THREADS, Cfg::SMEM, a.stream>>>(reinterpret_cast<const uint16_t *>(a.x), w_img,
                                                      reinterpret_cast<const uint16_t *>(a.bias), reinterpret_cast<uint16_t *>(a.y),
                                                      a.nbr, a.pitch, reinterpret_cast<const unsigned long long *>(a.tile_mask), a.n_out,
                                                      a.k3, idesc, bf16 ? 1 : 0);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

bool tc_forward_supported(int32_t cin, int32_t cout, int64_t k3, int32_t dtype) {
    if (dtype != FVC_F16 && dtype != FVC_BF16)
        return false;
    if (k3 < 1 || k3 > 64 * TC_MASK_WORDS)
        return false;
    const bool cin_ok = cin == 64 || cin == 128 || cin == 256;
    const bool cout_ok = cout == 32 || cout == 64 || cout == 128 || cout == 256;
    return cin_ok && cout_ok;
}

size_t tc_forward_scratch_bytes(int64_t, int32_t cin, int32_t cout, int64_t k3, int32_t) {
    return size_t(k3) * size_t(cin) * size_t(cout) * 2 + 256;
}

int tc_forward(const ConvArgs &a) {
    const size_t need = tc_forward_scratch_bytes(a.n_out, a.cin, a.cout, a.k3, a.dtype);
    FVC_REQUIRE(a.scratch && a.scratch_bytes >= need, FVC_ERR_RUNTIME, "tensor-core conv scratch too small: %zu < %zu",
                a.scratch_bytes, need);
    FVC_REQUIRE((reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a.scratch) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.nbr) & 15) == 0,
                FVC_ERR_RUNTIME, "tensor-core conv needs 16-byte aligned feature / output / scratch / map pointers");
    FVC_REQUIRE(a.pitch % 4 == 0 && a.pitch >= ceil_div(a.n_out, TC_TILE_M) * TC_TILE_M, FVC_ERR_RUNTIME,
                "tensor-core conv needs the map pitch (%lld) to be a multiple of 4 covering whole 128-row tiles", (long long)a.pitch);
    uint8_t *img = reinterpret_cast<uint8_t *>(a.scratch);
    const int64_t chunks16 = int64_t(a.k3) * a.cin * a.cout / 8;
    tc_pack_b_kernel<<<int(ceil_div(chunks16, 256) > 1184 ? 1184 : ceil_div(chunks16, 256)), 256, 0, a.stream>>>(
        reinterpret_cast<const uint16_t *>(a.w), a.k3, a.cin, a.cout, reinterpret_cast<uint4 *>(img));
    FVC_LAUNCH_CHECK();
    // experiment knob (scripts/bench_variants.py): alternative pipeline shapes for the 64 -> 64 kernel
    if (a.cin == 64 && a.cout == 64) {
        const char *v = getenv("FVC_TC_VARIANT");
        const int variant = v ? atoi(v) : 0;
        switch (variant) {
        case 1: return launch_tc_fwd<64, 64, 4, 4, 4>(a, img);
        case 2: return launch_tc_fwd<64, 64, 8, 9, 4>(a, img);
        case 3: return launch_tc_fwd<64, 64, 4, 3, 2>(a, img);
        case 4: return launch_tc_fwd<64, 64, 2, 4, 4>(a, img);
        case 5: return launch_tc_fwd<64, 64, 4, 5, 2>(a, img);
        case 6: return launch_tc_fwd<64, 64, 8, 6, 6>(a, img);
        default: break;
        }
    }
#define FVC_TC_CASE(CI, CO, T, S, B) \
    if (a.cin == CI && a.cout == CO) \
        return launch_tc_fwd<CI, CO, T, S, B>(a, img);
    FVC_TC_CASE(64, 32, 8, 4, 4)
    FVC_TC_CASE(64, 64, 4, 4, 4)
    FVC_TC_CASE(64, 128, 4, 8, 3)
    FVC_TC_CASE(64, 256, 2, 6, 3)
    FVC_TC_CASE(128, 32, 8, 4, 4)
    FVC_TC_CASE(128, 64, 4, 4, 4)
    FVC_TC_CASE(128, 128, 4, 8, 3)
    FVC_TC_CASE(128, 256, 2, 6, 3)
    FVC_TC_CASE(256, 32, 8, 4, 4)
    FVC_TC_CASE(256, 64, 4, 4, 4)
    FVC_TC_CASE(256, 128, 4, 8, 3)
    FVC_TC_CASE(256, 256, 2, 6, 3)
#undef FVC_TC_CASE
    return set_error(FVC_ERR_UNSUPPORTED, "no tensor-core kernel for channels %d -> %d", a.cin, a.cout);
}

// tensor-core weight gradient: conv_tc_wgrad.cu

} // namespace fvc
