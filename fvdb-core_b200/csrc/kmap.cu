// kmap.cu -- kernel-map construction, CSR view, lookups and target-topology candidate emission.
//
// Replaces GatherScatterDefault.cu:92-294 (two full sweeps with one un-cached root-to-leaf walk per
// (voxel, tap) probe and K^3 contended global counters).  Here one warp owns one *output leaf*: its
// lanes walk the tree once per source leaf of the probe neighbourhood (<= 3x3x3 leaves for every
// K <= 9 stride-1 kernel), stage those leaves' mask / prefix / base lines in shared memory with
// 16-byte loads, and answer all voxel x tap probes from shared memory.  The map is written tap-major
// so that consecutive rows of a leaf store coalesced.
#include "fvc_common.cuh"

#include <cub/cub.cuh>

namespace fvc {

constexpr int KM_WARPS = 4;       // output leaves per CTA: one warp owns one leaf
constexpr int KM_THREADS = KM_WARPS * 32;
constexpr int KM_MAX_NB = 32;     // cached source leaves per output leaf (3x3x3 = 27 covers every K <= 9 stride-1 kernel)
constexpr int KM_MAX_TAPS = 512;  // taps with shared-memory tap table / counters

struct LeafBox {
    int lmin[3];
    int n[3];
    int total;
};

// Source-leaf box reached from the output leaf at `origin` (ConvolutionGeometry.h:99-124 applied to
// the two corners of the leaf).
__device__ __forceinline__ LeafBox probe_box(const Geometry &g, const int origin[3], int transposed) {
    LeafBox box;
    box.total = 1;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        int lo, hi;
        if (!transposed) {
            lo = g.s[d] * origin[d] - g.pad[d];
            hi = g.s[d] * (origin[d] + 7) + g.k[d] - 1 - g.pad[d];
        } else {
            lo = floor_div(origin[d] - (g.k[d] - 1 - g.pad[d]), g.s[d]);
            hi = floor_div(origin[d] + 7 + g.pad[d], g.s[d]);
        }
        box.lmin[d] = lo >> 3;
        box.n[d] = (hi >> 3) - (lo >> 3) + 1;
        // saturate so that the product cannot overflow
        box.total = (box.total > KM_MAX_NB || box.n[d] > KM_MAX_NB) ? KM_MAX_NB + 1 : box.total * box.n[d];
    }
    return box;
}

// One WARP per output leaf (four leaves per CTA keep four independent latency chains in flight per CTA: ncu showed
// the one-CTA-per-leaf version waiting at block barriers behind 27 serial tree walks).  The warp's lanes walk the
// tree once per source leaf of the probe box, stage those leaves' mask / prefix / base lines in shared memory with
// 16-byte loads, then every lane keeps one output voxel in registers and answers its K^3 probes from shared memory.
__global__ void __launch_bounds__(KM_THREADS)
kmap_build_kernel(FvcGridBatch feat, FvcGridBatch out, Geometry g, int transposed, int32_t *__restrict__ nbr,
                  int64_t pitch, unsigned long long *__restrict__ tap_counts, unsigned long long *__restrict__ tile_mask, int mask_words) {
    __shared__ __align__(16) uint64_t s_mask[KM_WARPS][KM_MAX_NB][8];
    __shared__ __align__(16) uint16_t s_prefix[KM_WARPS][KM_MAX_NB][8];
    __shared__ int s_base[KM_WARPS][KM_MAX_NB];
    __shared__ int s_leaf[KM_WARPS][KM_MAX_NB];
    __shared__ uint16_t s_vox[KM_WARPS][512];
    __shared__ uint32_t s_tap[KM_MAX_TAPS];
    __shared__ uint32_t s_cnt[KM_MAX_TAPS];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k3 = int(g.volume);
    const bool small_taps = k3 <= KM_MAX_TAPS;
    const int k12 = g.k[1] * g.k[2];
    if (small_taps) {
        for (int k = tid; k < k3; k += KM_THREADS) {
            s_tap[k] = (uint32_t(k / k12) << 20) | (uint32_t((k / g.k[2]) % g.k[1]) << 10) | uint32_t(k % g.k[2]);
            s_cnt[k] = 0;
        }
    }
    __syncthreads();

    const int leaf_id = blockIdx.x * KM_WARPS + warp;
    if (leaf_id < out.num_leaves) {
        const FvcLeaf *L = out.leaves + leaf_id;
        const int b = __ldg(&L->batch);
        const int origin[3] = {__ldg(&L->origin[0]), __ldg(&L->origin[1]), __ldg(&L->origin[2])};
        const int base = __ldg(&L->base);
        const int cnt = __ldg(&L->count);
        const LeafBox box = probe_box(g, origin, transposed);
        const bool cached = box.total <= KM_MAX_NB;

        // (1) lanes walk the tree in parallel, one source leaf each
        for (int t = lane; cached && t < box.total; t += 32) {
            const int c = t % box.n[2], bb = (t / box.n[2]) % box.n[1], a = t / (box.n[2] * box.n[1]);
            const int leaf = find_leaf(feat, b, (box.lmin[0] + a) << 3, (box.lmin[1] + bb) << 3, (box.lmin[2] + c) << 3);
            s_leaf[warp][t] = leaf;
            s_base[warp][t] = leaf >= 0 ? __ldg(&feat.leaves[leaf].base) : 0;
        }
        // (2) compact the output leaf's active voxels: rank inside the leaf == row - base.  Lane l scans 16 bits.
        {
            const int w = lane >> 2, shift = (lane & 3) * 16;
            const uint64_t m = __ldg(L->mask + w);
            const int before = int(__ldg(L->prefix + w)) + __popcll(m & ((1ull << shift) - 1ull));
            uint32_t bits = uint32_t(m >> shift) & 0xFFFFu;
            for (int r = before; bits; bits &= bits - 1u, ++r)
                s_vox[warp][r] = uint16_t((w << 6) + shift + __ffs(bits) - 1);
        }
        __syncwarp();
        // (3) stage mask (4 x 16 B) + prefix (1 x 16 B) of every source leaf
        if (cached) {
            for (int e = lane; e < box.total * 5; e += 32) {
                const int li = e / 5, part = e % 5, leaf = s_leaf[warp][li];
                uint4 v = make_uint4(0, 0, 0, 0);
                if (leaf >= 0)
                    v = __ldg(reinterpret_cast<const uint4 *>(feat.leaves + leaf) + part);
                if (part < 4)
                    reinterpret_cast<uint4 *>(&s_mask[warp][li][0])[part] = v;
                else
                    *reinterpret_cast<uint4 *>(&s_prefix[warp][li][0]) = v;
            }
        }
        __syncwarp();

        // (4) probes: a lane keeps one voxel's coordinates in registers and walks all taps; stores to nbr[k][base + j]
        //     coalesce across the lanes; per-tap pair counts come from one ballot per tap.
        const bool unit_stride = g.s[0] == 1 && g.s[1] == 1 && g.s[2] == 1;
        for (int j0 = 0; j0 < cnt; j0 += 32) {
            const int j = j0 + lane;
            const bool has = j < cnt;
            const int n = has ? s_vox[warp][j] : 0;
            // forward: S * c - pad (then + tap);  transposed: c + pad (then - tap, then / S)
            int c[3] = {origin[0] + (n >> 6), origin[1] + ((n >> 3) & 7), origin[2] + (n & 7)};
#pragma unroll
            for (int d = 0; d < 3; ++d)
                c[d] = transposed ? c[d] + g.pad[d] : g.s[d] * c[d] - g.pad[d];
            // tile tap-mask, fused (it used to be a second pass over the whole map): the 32 rows of this chunk lie in at most two
            // 128-row tiles; the taps that hit are collected per tile in registers and OR-ed into the mask once per 64 taps
            const int64_t tile_lo = (int64_t(base) + j0) >> 7;
            const unsigned in_lo = __ballot_sync(0xffffffffu, ((int64_t(base) + j) >> 7) == tile_lo);
            unsigned long long taps_lo = 0ull, taps_hi = 0ull;
            for (int k = 0; k < k3; ++k) {
                int t[3];
                if (small_taps) {
                    const uint32_t packed = s_tap[k];
                    t[0] = int(packed >> 20), t[1] = int((packed >> 10) & 1023), t[2] = int(packed & 1023);
                } else {
                    t[0] = k / k12, t[1] = (k / g.k[2]) % g.k[1], t[2] = k % g.k[2];
                }
                int p[3];
                bool ok = has;
                if (!transposed) {
#pragma unroll
                    for (int d = 0; d < 3; ++d)
                        p[d] = c[d] + t[d]; // fineFromCoarse
                } else if (unit_stride) {
#pragma unroll
                    for (int d = 0; d < 3; ++d)
                        p[d] = c[d] - t[d];
                } else {
#pragma unroll
                    for (int d = 0; d < 3; ++d) { // coarseFromFine with the divisibility test
                        const int numer = c[d] - t[d];
                        ok = ok && floor_mod(numer, g.s[d]) == 0;
                        p[d] = floor_div(numer, g.s[d]);
                    }
                }
                int val = -1;
                if (ok) {
                    if (cached) {
                        const int li = (((p[0] >> 3) - box.lmin[0]) * box.n[1] + ((p[1] >> 3) - box.lmin[1])) * box.n[2] +
                                       ((p[2] >> 3) - box.lmin[2]);
                        if (s_leaf[warp][li] >= 0) {
                            const int w = p[0] & 7;
                            val = leaf_value(s_mask[warp][li][w], s_prefix[warp][li][w], s_base[warp][li], p[1], p[2]);
                        }
                    } else {
                        val = lookup_row(feat, b, p[0], p[1], p[2]);
                    }
                }
                if (has)
                    nbr[int64_t(k) * pitch + base + j] = val;
                const unsigned hits = __ballot_sync(0xffffffffu, val >= 0);
                if (lane == 0 && hits) {
                    if (small_taps)
                        atomicAdd(&s_cnt[k], uint32_t(__popc(hits)));
                    else
                        atomicAdd(tap_counts + k, (unsigned long long)__popc(hits));
                }
                if (tile_mask) {
                    taps_lo |= (unsigned long long)((hits & in_lo) != 0u) << (k & 63);
                    taps_hi |= (unsigned long long)((hits & ~in_lo) != 0u) << (k & 63);
                    if ((k & 63) == 63 || k == k3 - 1) {
                        if (lane == 0 && taps_lo)
                            atomicOr(tile_mask + tile_lo * mask_words + (k >> 6), taps_lo);
                        if (lane == 0 && taps_hi)
                            atomicOr(tile_mask + (tile_lo + 1) * mask_words + (k >> 6), taps_hi);
                        taps_lo = taps_hi = 0ull;
                    }
                }
            }
        }
    }
    if (small_taps) {
        __syncthreads();
        for (int k = tid; k < k3; k += KM_THREADS)
            if (s_cnt[k])
                atomicAdd(tap_counts + k, (unsigned long long)s_cnt[k]);
    }
}

// Fast path of the same build: every forward map (any stride) and the unit-stride transposed maps whose probe box fits the
// cache.  ncu on the general kernel showed it issue-bound (~130 warp instructions per (32 voxels, tap): 64-bit mask
// arithmetic, per-probe leaf-index and address arithmetic, one shared-memory atomic per tap).  Here
//   * a cached source leaf is re-encoded once per output leaf as [x word][32-bit half] entries {mask half, row of the half's
//     first voxel} -- one 8-byte shared load answers a probe with 32-bit arithmetic, a missing leaf is an all-zero mask;
//   * taps run as nested (t0, t1, t2) loops, so the x / y parts of the probe, of the leaf index and of the bit position are
//     hoisted, and the store pointer advances by the pitch;
//   * kernels of <= 32 taps keep the per-tap pair counts in registers (lane k owns tap k) instead of shared-memory atomics.
template <bool SMALL32>
__global__ void __launch_bounds__(KM_THREADS)
kmap_build_fast_kernel(FvcGridBatch feat, FvcGridBatch out, Geometry g, int transposed, int32_t *__restrict__ nbr,
                       int64_t pitch, unsigned long long *__restrict__ tap_counts, unsigned long long *__restrict__ tile_mask, int mask_words) {
    __shared__ __align__(16) uint2 s_entry[KM_WARPS][KM_MAX_NB][8][2]; // {mask half, batch-cumulative row before the half}
    __shared__ int s_leaf[KM_WARPS][KM_MAX_NB];
    __shared__ uint16_t s_vox[KM_WARPS][512];
    __shared__ uint32_t s_cnt[SMALL32 ? 1 : KM_MAX_TAPS];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k3 = int(g.volume);
    if (!SMALL32) {
        for (int k = tid; k < k3; k += KM_THREADS)
            s_cnt[k] = 0;
        __syncthreads();
    }
    uint32_t my_count = 0; // SMALL32: pairs of tap `lane` seen by this warp
    const int leaf_id = blockIdx.x * KM_WARPS + warp;
    if (leaf_id < out.num_leaves) {
        const FvcLeaf *L = out.leaves + leaf_id;
        const int b = __ldg(&L->batch);
        const int origin[3] = {__ldg(&L->origin[0]), __ldg(&L->origin[1]), __ldg(&L->origin[2])};
        const int base = __ldg(&L->base);
        const int cnt = __ldg(&L->count);
        const LeafBox box = probe_box(g, origin, transposed);
        // (1) lanes walk the tree in parallel, one source leaf each
        for (int t = lane; t < box.total; t += 32) {
            const int c = t % box.n[2], bb = (t / box.n[2]) % box.n[1], a = t / (box.n[2] * box.n[1]);
            s_leaf[warp][t] = find_leaf(feat, b, (box.lmin[0] + a) << 3, (box.lmin[1] + bb) << 3, (box.lmin[2] + c) << 3);
        }
        // (2) compact the output leaf's active voxels: rank inside the leaf == row - base.  Lane l scans 16 bits.
        {
            const int w = lane >> 2, shift = (lane & 3) * 16;
            const uint64_t m = __ldg(L->mask + w);
            const int before = int(__ldg(L->prefix + w)) + __popcll(m & ((1ull << shift) - 1ull));
            uint32_t bits = uint32_t(m >> shift) & 0xFFFFu;
            for (int r = before; bits; bits &= bits - 1u, ++r)
                s_vox[warp][r] = uint16_t((w << 6) + shift + __ffs(bits) - 1);
        }
        __syncwarp();
        // (3) re-encode every source leaf: entry (leaf, x word w, half h) <- one 8-byte mask word + its prefix count
        for (int e = lane; e < box.total * 8; e += 32) {
            const int li = e >> 3, w = e & 7, leaf = s_leaf[warp][li];
            uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
            if (leaf >= 0) {
                const FvcLeaf *S = feat.leaves + leaf;
                const uint64_t m = __ldg(S->mask + w);
                const uint32_t first = uint32_t(__ldg(&S->base)) + uint32_t(__ldg(S->prefix + w));
                lo = make_uint2(uint32_t(m), first);
                hi = make_uint2(uint32_t(m >> 32), first + uint32_t(__popc(uint32_t(m))));
            }
            s_entry[warp][li][w][0] = lo;
            s_entry[warp][li][w][1] = hi;
        }
        __syncwarp();
        // (4) probes
        const uint2 *entries = &s_entry[warp][0][0][0];
        const int dir = transposed ? -1 : 1; // forward: S * c - pad + t;  unit-stride transposed: c + pad - t
        for (int j0 = 0; j0 < cnt; j0 += 32) {
            const int j = j0 + lane;
            const bool has = j < cnt;
            const int n = has ? s_vox[warp][j] : 0;
            int c0[3] = {origin[0] + (n >> 6), origin[1] + ((n >> 3) & 7), origin[2] + (n & 7)};
#pragma unroll
            for (int d = 0; d < 3; ++d)
                c0[d] = transposed ? c0[d] + g.pad[d] : g.s[d] * c0[d] - g.pad[d];
            const int64_t tile_lo = (int64_t(base) + j0) >> 7;
            const unsigned in_lo = __ballot_sync(0xffffffffu, ((int64_t(base) + j) >> 7) == tile_lo);
            unsigned long long taps_lo = 0ull, taps_hi = 0ull;
            int32_t *dst = nbr + base + j; // + k * pitch
            int k = 0;
            for (int t0 = 0; t0 < g.k[0]; ++t0) {
                const int px = c0[0] + dir * t0;
                const int ex = (((px >> 3) - box.lmin[0]) * box.n[1]) , wx = px & 7;
                for (int t1 = 0; t1 < g.k[1]; ++t1) {
                    const int py = c0[1] + dir * t1;
                    const int exy = (ex + ((py >> 3) - box.lmin[1])) * box.n[2] - box.lmin[2];
                    const int half = (py >> 2) & 1, ybit = (py & 3) << 3;
                    for (int t2 = 0; t2 < g.k[2]; ++t2, ++k, dst += pitch) {
                        const int pz = c0[2] + dir * t2;
                        const int li = exy + (pz >> 3);
                        const uint2 en = entries[((li << 3) + wx) * 2 + half];
                        const int bit = ybit | (pz & 7);
                        const bool hit = has && ((en.x >> bit) & 1u);
                        const int val = hit ? int(en.y + uint32_t(__popc(en.x & ((1u << bit) - 1u)))) : -1;
                        if (has)
                            *dst = val;
                        const unsigned hits = __ballot_sync(0xffffffffu, hit);
                        if (SMALL32) {
                            if (lane == k)
                                my_count += uint32_t(__popc(hits));
                        } else if (lane == 0 && hits) {
                            atomicAdd(&s_cnt[k], uint32_t(__popc(hits)));
                        }
                        if (tile_mask) {
                            taps_lo |= (unsigned long long)((hits & in_lo) != 0u) << (k & 63);
                            taps_hi |= (unsigned long long)((hits & ~in_lo) != 0u) << (k & 63);
                            if ((k & 63) == 63 || k == k3 - 1) {
                                if (lane == 0 && taps_lo)
                                    atomicOr(tile_mask + tile_lo * mask_words + (k >> 6), taps_lo);
                                if (lane == 0 && taps_hi)
                                    atomicOr(tile_mask + (tile_lo + 1) * mask_words + (k >> 6), taps_hi);
                                taps_lo = taps_hi = 0ull;
                            }
                        }
                    }
                }
            }
        }
    }
    if (SMALL32) {
        if (lane < k3 && my_count)
            atomicAdd(tap_counts + lane, (unsigned long long)my_count);
    } else {
        __syncthreads();
        for (int k = tid; k < k3; k += KM_THREADS)
            if (s_cnt[k])
                atomicAdd(tap_counts + k, (unsigned long long)s_cnt[k]);
    }
}

// ---- CSR-by-tap view ------------------------------------------------------------------------------
constexpr int CSR_THREADS = 256;
constexpr int CSR_ROWS_PER_THREAD = 8;
constexpr int CSR_ROWS = CSR_THREADS * CSR_ROWS_PER_THREAD;

__global__ void __launch_bounds__(CSR_THREADS)
csr_count_kernel(const int32_t *__restrict__ nbr, int64_t pitch, int64_t n_out, int64_t nblk,
                 int64_t *__restrict__ counts) {
    const int64_t k = blockIdx.y, blk = blockIdx.x;
    const int64_t row0 = blk * CSR_ROWS + int64_t(threadIdx.x) * CSR_ROWS_PER_THREAD;
    int local = 0;
#pragma unroll
    for (int r = 0; r < CSR_ROWS_PER_THREAD; ++r)
        if (row0 + r < n_out)
            local += nbr[k * pitch + row0 + r] >= 0;
    using BlockReduce = cub::BlockReduce<int, CSR_THREADS>;
    __shared__ typename BlockReduce::TempStorage temp;
    const int total = BlockReduce(temp).Sum(local);
    if (threadIdx.x == 0)
        counts[k * nblk + blk] = total;
}

__global__ void __launch_bounds__(CSR_THREADS)
csr_fill_kernel(const int32_t *__restrict__ nbr, int64_t pitch, int64_t n_out, int64_t nblk,
                const int64_t *__restrict__ block_offsets, int32_t *__restrict__ gather, int32_t *__restrict__ scatter) {
    const int64_t k = blockIdx.y, blk = blockIdx.x;
    const int64_t row0 = blk * CSR_ROWS + int64_t(threadIdx.x) * CSR_ROWS_PER_THREAD;
    int vals[CSR_ROWS_PER_THREAD];
    int local = 0;
#pragma unroll
    for (int r = 0; r < CSR_ROWS_PER_THREAD; ++r) {
        vals[r] = row0 + r < n_out ? nbr[k * pitch + row0 + r] : -1;
        local += vals[r] >= 0;
    }
    using BlockScan = cub::BlockScan<int, CSR_THREADS>;
    __shared__ typename BlockScan::TempStorage temp;
    int rank;
    BlockScan(temp).ExclusiveSum(local, rank);
    int64_t pos = block_offsets[k * nblk + blk] + rank;
#pragma unroll
    for (int r = 0; r < CSR_ROWS_PER_THREAD; ++r)
        if (vals[r] >= 0) {
            gather[pos] = vals[r];
            scatter[pos] = int32_t(row0 + r);
            ++pos;
        }
}

__global__ void csr_offsets_kernel(const int64_t *__restrict__ block_offsets, const int64_t *__restrict__ last_count,
                                   int64_t nblk, int64_t k3, int64_t *__restrict__ offsets) {
    const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (k < k3)
        offsets[k] = block_offsets[k * nblk];
    else if (k == k3)
        offsets[k3] = k3 * nblk > 0 ? block_offsets[k3 * nblk - 1] + *last_count : 0;
}

__global__ void reverse_dense_kernel(const int32_t *__restrict__ gather, const int32_t *__restrict__ scatter,
                                     const int64_t *__restrict__ offsets, int k3, int64_t total,
                                     int32_t *__restrict__ nbr_rev, int64_t pitch_rev) {
    for (int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; p < total; p += int64_t(gridDim.x) * blockDim.x) {
        int lo = 0, hi = k3; // largest k with offsets[k] <= p
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(offsets + mid) <= p)
                lo = mid;
            else
                hi = mid;
        }
        nbr_rev[int64_t(lo) * pitch_rev + gather[p]] = scatter[p];
    }
}

// nbr_rev[k][i] = o for every (k, o) with nbr[k][o] = i >= 0: the input-stationary map straight from the output-stationary
// one (no CSR detour); a feature row is reached through tap k by at most one output row, so there is no write conflict
__global__ void reverse_from_dense_kernel(const int32_t *__restrict__ nbr, int64_t pitch, int64_t n_out, int k3,
                                          int32_t *__restrict__ nbr_rev, int64_t pitch_rev) {
    const int64_t total = int64_t(k3) * n_out;
    for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
        const int64_t k = e / n_out, o = e - k * n_out;
        const int32_t i = __ldg(nbr + k * pitch + o);
        if (i >= 0)
            nbr_rev[k * pitch_rev + i] = int32_t(o);
    }
}

__global__ void degree_kernel(const int32_t *__restrict__ nbr, int64_t pitch, int64_t n_out, int k3,
                              int32_t *__restrict__ degree) {
    for (int64_t o = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; o < n_out; o += int64_t(gridDim.x) * blockDim.x) {
        int d = 0;
        for (int k = 0; k < k3; ++k)
            d += nbr[int64_t(k) * pitch + o] >= 0;
        degree[o] = d;
    }
}

// one CTA per 128-row tile; bit k of the tile's mask = "some row of the tile has a neighbour through tap k"
__global__ void __launch_bounds__(128)
tile_mask_kernel(const int32_t *__restrict__ nbr, int64_t pitch, int64_t n_out, int k3, int words,
                 unsigned long long *__restrict__ mask) {
    extern __shared__ unsigned long long s_words[];
    for (int w = threadIdx.x; w < words; w += blockDim.x)
        s_words[w] = 0ull;
    __syncthreads();
    const int64_t row = int64_t(blockIdx.x) * 128 + threadIdx.x;
    for (int k = 0; k < k3; ++k) {
        const bool valid = row < n_out && nbr[int64_t(k) * pitch + row] >= 0;
        if (__ballot_sync(0xffffffffu, valid) != 0u && (threadIdx.x & 31) == 0)
            atomicOr(&s_words[k >> 6], 1ull << (k & 63));
    }
    __syncthreads();
    for (int w = threadIdx.x; w < words; w += blockDim.x)
        mask[int64_t(blockIdx.x) * words + w] = s_words[w];
}

// ---- lookups --------------------------------------------------------------------------------------
__global__ void neighbor_indexes_kernel(FvcGridBatch grid, const int32_t *__restrict__ q_ijk,
                                        const int32_t *__restrict__ q_bidx, int64_t nq, int extent, int shift,
                                        int64_t *__restrict__ out) {
    const int w = 2 * extent + 1, w3 = w * w * w;
    const int64_t total = nq * w3;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t q = i / w3;
        const int o = int(i - q * w3);
        const int b = q_bidx ? q_bidx[q] : 0;
        const int x = (q_ijk[3 * q] << shift) + o / (w * w) - extent;
        const int y = (q_ijk[3 * q + 1] << shift) + (o / w) % w - extent;
        const int z = (q_ijk[3 * q + 2] << shift) + o % w - extent;
        const int row = lookup_row(grid, b, x, y, z);
        out[i] = row >= 0 ? int64_t(row) - __ldg(grid.voxel_offsets + b) : -1; // per-grid-local (NeighborIndexes.cu:40)
    }
}

__global__ void ijk_to_index_kernel(FvcGridBatch grid, const int32_t *__restrict__ q_ijk,
                                    const int32_t *__restrict__ q_bidx, int64_t nq, int cumulative,
                                    int64_t *__restrict__ out) {
    for (int64_t q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; q < nq; q += int64_t(gridDim.x) * blockDim.x) {
        const int b = q_bidx ? q_bidx[q] : 0;
        const int row = lookup_row(grid, b, q_ijk[3 * q], q_ijk[3 * q + 1], q_ijk[3 * q + 2]);
        out[q] = row < 0 ? -1 : (cumulative ? int64_t(row) : int64_t(row) - __ldg(grid.voxel_offsets + b));
    }
}

// ---- generated target topologies ------------------------------------------------------------------
// number of taps t in [0, k) with (c - t + pad) divisible by s  == taps congruent to (c + pad) mod s
__device__ __forceinline__ int divisible_taps(int c, int k, int s, int pad) {
    const int r = floor_mod(c + pad, s);
    return r < k ? (k - 1 - r) / s + 1 : 0;
}

__global__ void conv_grid_count_kernel(const int32_t *__restrict__ ijk, int64_t n, Geometry g,
                                       unsigned long long *__restrict__ counter) {
    unsigned long long local = 0;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        local += (unsigned long long)divisible_taps(ijk[3 * i], g.k[0], g.s[0], g.pad[0]) *
                 divisible_taps(ijk[3 * i + 1], g.k[1], g.s[1], g.pad[1]) *
                 divisible_taps(ijk[3 * i + 2], g.k[2], g.s[2], g.pad[2]);
    }
    using BlockReduce = cub::BlockReduce<unsigned long long, 256>;
    __shared__ typename BlockReduce::TempStorage temp;
    const unsigned long long total = BlockReduce(temp).Sum(local);
    if (threadIdx.x == 0 && total)
        atomicAdd(counter, total);
}

// forward: each fine voxel emits floorDiv(fine - tap + pad, S) for every divisible tap
// (BuildGridForConv.cu:485-510); candidates are claimed with one aggregated atomic per warp (order is irrelevant,
// the grid builder sorts and de-duplicates).
__global__ void conv_grid_emit_fwd_kernel(const int32_t *__restrict__ ijk, const int32_t *__restrict__ bidx, int64_t n,
                                          Geometry g, int64_t capacity, int32_t *__restrict__ cand_ijk,
                                          int32_t *__restrict__ cand_bidx, unsigned long long *__restrict__ counter) {
    const int lane = threadIdx.x & 31;
    // warp-uniform loop; one aggregated atomic per warp claims the candidates of its 32 voxels
    for (int64_t i0 = blockIdx.x * int64_t(blockDim.x) + threadIdx.x - lane; i0 < n; i0 += int64_t(gridDim.x) * blockDim.x) {
        const int64_t i = i0 + lane;
        const bool valid = i < n;
        int c[3] = {0, 0, 0}, r[3] = {0, 0, 0}, m[3] = {0, 0, 0};
        if (valid) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                c[d] = ijk[3 * i + d];
                r[d] = floor_mod(c[d] + g.pad[d], g.s[d]); // smallest admissible tap
                m[d] = divisible_taps(c[d], g.k[d], g.s[d], g.pad[d]);
            }
        }
        const int total = m[0] * m[1] * m[2];
        int incl = total;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d)
                incl += v;
        }
        const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned long long warp_base = 0ull;
        if (lane == 0 && warp_total > 0)
            warp_base = atomicAdd(counter, (unsigned long long)warp_total);
        warp_base = __shfl_sync(0xffffffffu, warp_base, 0);
        if (total == 0)
            continue;
        int64_t pos = int64_t(warp_base) + incl - total;
        const int b = bidx ? bidx[i] : 0;
        for (int a = 0; a < m[0]; ++a)
            for (int bb = 0; bb < m[1]; ++bb)
                for (int cc = 0; cc < m[2]; ++cc, ++pos) {
                    if (pos >= capacity)
                        break;
                    cand_ijk[3 * pos] = floor_div(c[0] - (r[0] + a * g.s[0]) + g.pad[0], g.s[0]);
                    cand_ijk[3 * pos + 1] = floor_div(c[1] - (r[1] + bb * g.s[1]) + g.pad[1], g.s[1]);
                    cand_ijk[3 * pos + 2] = floor_div(c[2] - (r[2] + cc * g.s[2]) + g.pad[2], g.s[2]);
                    cand_bidx[pos] = b;
                }
    }
}

// transposed: every coarse voxel spreads through every tap (BuildGridForConvTranspose.cu:326-337)
__global__ void conv_grid_emit_tr_kernel(const int32_t *__restrict__ ijk, const int32_t *__restrict__ bidx, int64_t n,
                                         Geometry g, int32_t *__restrict__ cand_ijk, int32_t *__restrict__ cand_bidx) {
    const int64_t k3 = g.volume, total = n * k3;
    const int k12 = g.k[1] * g.k[2];
    for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
        const int64_t i = e / k3;
        const int k = int(e - i * k3);
        cand_ijk[3 * e] = g.s[0] * ijk[3 * i] + k / k12 - g.pad[0];
        cand_ijk[3 * e + 1] = g.s[1] * ijk[3 * i + 1] + (k / g.k[2]) % g.k[1] - g.pad[1];
        cand_ijk[3 * e + 2] = g.s[2] * ijk[3 * i + 2] + k % g.k[2] - g.pad[2];
        cand_bidx[e] = bidx ? bidx[i] : 0;
    }
}

static inline int grid_for(int64_t n, int block) {
    int64_t blocks = ceil_div(n, block);
    return int(blocks < 1 ? 1 : (blocks > 148 * 16 ? 148 * 16 : blocks));
}

static int check_geometry_args(const int32_t kernel_size[3], const int32_t stride[3]) {
    for (int d = 0; d < 3; ++d) {
        FVC_REQUIRE(kernel_size[d] > 0, FVC_ERR_RUNTIME, "kernel_size[%d] must be positive, got %d", d, kernel_size[d]);
        FVC_REQUIRE(stride[d] > 0, FVC_ERR_RUNTIME, "stride[%d] must be positive, got %d", d, stride[d]);
    }
    return FVC_OK;
}

} // namespace fvc

using namespace fvc;

extern "C" {

int fvc_kmap_build(const FvcGridBatch *feature_grid, const FvcGridBatch *output_grid, const int32_t kernel_size[3],
                   const int32_t stride[3], int32_t transposed, int32_t *nbr, int64_t pitch, int64_t *tap_counts, uint64_t *tile_mask,
                   fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    FVC_REQUIRE(feature_grid && output_grid, FVC_ERR_RUNTIME, "feature_grid and output_grid must be provided");
    int rc = check_geometry_args(kernel_size, stride); // GatherScatterDefault.cu:63-67
    if (rc)
        return rc;
    FVC_REQUIRE(feature_grid->num_grids == output_grid->num_grids, FVC_ERR_RUNTIME,
                "feature_grid and output_grid must have the same batch size, got %d and %d", feature_grid->num_grids,
                output_grid->num_grids);
    FVC_REQUIRE(feature_grid->total_voxels <= INT32_MAX, FVC_ERR_RUNTIME,
                "feature_grid has %lld voxels, exceeding the int32 index limit (%d)",
                (long long)feature_grid->total_voxels, INT32_MAX); // :68-74
    FVC_REQUIRE(output_grid->total_voxels <= INT32_MAX, FVC_ERR_RUNTIME,
                "output_grid has %lld voxels, exceeding the int32 index limit (%d)",
                (long long)output_grid->total_voxels, INT32_MAX); // :75-80
    const Geometry g = make_geometry(kernel_size, stride);
    FVC_REQUIRE(g.volume <= INT32_MAX, FVC_ERR_RUNTIME, "kernel volume %lld exceeds INT32_MAX", (long long)g.volume);
    FVC_REQUIRE(pitch >= output_grid->total_voxels, FVC_ERR_RUNTIME, "map pitch %lld < output voxels %lld",
                (long long)pitch, (long long)output_grid->total_voxels);
    FVC_CUDA(cudaMemsetAsync(tap_counts, 0, size_t(g.volume) * 8, stream));
    const int mask_words = int((g.volume + 63) / 64);
    if (tile_mask) {
        FVC_REQUIRE(g.volume <= 4096, FVC_ERR_UNSUPPORTED, "kernel volume %lld exceeds the tile-mask limit 4096", (long long)g.volume);
        FVC_CUDA(cudaMemsetAsync(tile_mask, 0, size_t(ceil_div(output_grid->total_voxels, 128)) * mask_words * 8, stream));
    }
    if (output_grid->num_leaves == 0)
        return FVC_OK;
    // fast path: forward maps of any stride and unit-stride transposed maps whose probe box (the same for every leaf: origins
    // are multiples of 8) fits the per-warp leaf cache, with at most KM_MAX_TAPS taps
    bool fast = g.volume <= KM_MAX_TAPS && (!transposed || (g.s[0] == 1 && g.s[1] == 1 && g.s[2] == 1));
    if (fast) {
        int total = 1;
        for (int d = 0; d < 3; ++d) {
            // widest case over the origin's residue: span of probes of one leaf, in leaves (+1 for an unaligned start)
            const int span = transposed ? (7 + g.k[d] - 1) : (g.s[d] * 7 + g.k[d] - 1);
            total *= span / 8 + 2;
        }
        fast = total <= KM_MAX_NB;
    }
    const unsigned grid = unsigned(ceil_div(output_grid->num_leaves, KM_WARPS));
    unsigned long long *counts = reinterpret_cast<unsigned long long *>(tap_counts), *tmask = reinterpret_cast<unsigned long long *>(tile_mask);
    if (fast && g.volume <= 32)
        kmap_build_fast_kernel<true><<<grid, KM_THREADS, 0, stream>>>(*feature_grid, *output_grid, g, transposed ? 1 : 0, nbr, pitch, counts, tmask, mask_words);
    else if (fast)
        kmap_build_fast_kernel<false><<<grid, KM_THREADS, 0, stream>>>(*feature_grid, *output_grid, g, transposed ? 1 : 0, nbr, pitch, counts, tmask, mask_words);
    else
        kmap_build_kernel<<<grid, KM_THREADS, 0, stream>>>(*feature_grid, *output_grid, g, transposed ? 1 : 0, nbr, pitch, counts, tmask, mask_words);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

size_t fvc_kmap_csr_scratch_bytes(int64_t n_out, int64_t kernel_volume) {
    const int64_t nblk = ceil_div(n_out > 0 ? n_out : 1, CSR_ROWS);
    const int64_t m = nblk * (kernel_volume > 0 ? kernel_volume : 1);
    size_t temp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, temp, (int64_t *)nullptr, (int64_t *)nullptr, m);
    return align_up(size_t(m) * 8, 256) + 256 + align_up(temp, 256);
}

int fvc_kmap_to_csr(const int32_t *nbr, int64_t pitch, int64_t n_out, int64_t kernel_volume,
                    const int64_t *tap_counts, int64_t *offsets_dev, int32_t *gather, int32_t *scatter, void *scratch,
                    size_t scratch_bytes, fvc_stream_t stream_) {
    (void)tap_counts;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    FVC_REQUIRE(kernel_volume >= 0 && kernel_volume <= 65535, FVC_ERR_UNSUPPORTED,
                "kernel volume %lld exceeds the CSR builder limit 65535", (long long)kernel_volume);
    if (n_out == 0 || kernel_volume == 0) {
        FVC_CUDA(cudaMemsetAsync(offsets_dev, 0, size_t(kernel_volume + 1) * 8, stream));
        return FVC_OK;
    }
    const int64_t nblk = ceil_div(n_out, CSR_ROWS), m = nblk * kernel_volume;
    FVC_REQUIRE(scratch && scratch_bytes >= fvc_kmap_csr_scratch_bytes(n_out, kernel_volume), FVC_ERR_RUNTIME,
                "CSR scratch too small");
    int64_t *counts = reinterpret_cast<int64_t *>(scratch);
    int64_t *last = reinterpret_cast<int64_t *>(reinterpret_cast<char *>(scratch) + align_up(size_t(m) * 8, 256));
    void *temp = reinterpret_cast<char *>(last) + 256;
    size_t temp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, counts, counts, m);
    dim3 grid((unsigned)nblk, (unsigned)kernel_volume);
    csr_count_kernel<<<grid, CSR_THREADS, 0, stream>>>(nbr, pitch, n_out, nblk, counts);
    FVC_LAUNCH_CHECK();
    FVC_CUDA(cudaMemcpyAsync(last, counts + (m - 1), 8, cudaMemcpyDeviceToDevice, stream));
    FVC_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, counts, counts, m, stream));
    g_launch_count.fetch_add(1);
    csr_offsets_kernel<<<int(ceil_div(kernel_volume + 1, 256)), 256, 0, stream>>>(counts, last, nblk, kernel_volume,
                                                                                   offsets_dev);
    FVC_LAUNCH_CHECK();
    csr_fill_kernel<<<grid, CSR_THREADS, 0, stream>>>(nbr, pitch, n_out, nblk, counts, gather, scatter);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int fvc_kmap_reverse_dense(const int32_t *gather, const int32_t *scatter, const int64_t *offsets_dev,
                           int64_t kernel_volume, int64_t total_pairs, int64_t n_feature, int32_t *nbr_rev,
                           int64_t pitch_rev, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    FVC_REQUIRE(pitch_rev >= n_feature, FVC_ERR_RUNTIME, "reverse map pitch %lld < feature voxels %lld",
                (long long)pitch_rev, (long long)n_feature);
    if (kernel_volume == 0 || pitch_rev == 0)
        return FVC_OK;
    FVC_CUDA(cudaMemsetAsync(nbr_rev, 0xFF, size_t(kernel_volume) * size_t(pitch_rev) * 4, stream));
    if (total_pairs == 0)
        return FVC_OK;
    reverse_dense_kernel<<<grid_for(total_pairs, 256), 256, 0, stream>>>(gather, scatter, offsets_dev, int(kernel_volume),
                                                                         total_pairs, nbr_rev, pitch_rev);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int fvc_kmap_reverse_from_dense(const int32_t *nbr, int64_t pitch, int64_t n_out, int64_t kernel_volume, int64_t n_feature, int32_t *nbr_rev,
                                int64_t pitch_rev, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    FVC_REQUIRE(pitch_rev >= n_feature && pitch >= n_out, FVC_ERR_RUNTIME, "map pitch smaller than the row count");
    if (kernel_volume == 0 || pitch_rev == 0)
        return FVC_OK;
    FVC_CUDA(cudaMemsetAsync(nbr_rev, 0xFF, size_t(kernel_volume) * size_t(pitch_rev) * 4, stream));
    if (n_out == 0)
        return FVC_OK;
    reverse_from_dense_kernel<<<grid_for(kernel_volume * n_out, 256), 256, 0, stream>>>(nbr, pitch, n_out, int(kernel_volume), nbr_rev, pitch_rev);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int fvc_kmap_tile_mask(const int32_t *nbr, int64_t pitch, int64_t n_out, int64_t kernel_volume, uint64_t *mask,
                       fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n_out == 0 || kernel_volume == 0)
        return FVC_OK;
    FVC_REQUIRE(kernel_volume <= 4096, FVC_ERR_UNSUPPORTED, "kernel volume %lld exceeds the tile-mask limit 4096", (long long)kernel_volume);
    const int words = int((kernel_volume + 63) / 64);
    tile_mask_kernel<<<unsigned(ceil_div(n_out, 128)), 128, words * 8, stream>>>(nbr, pitch, n_out, int(kernel_volume), words,
                                                                                reinterpret_cast<unsigned long long *>(mask));
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int fvc_kmap_degree(const int32_t *nbr, int64_t pitch, int64_t n_out, int64_t kernel_volume, int32_t *degree,
                    fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n_out == 0)
        return FVC_OK;
    degree_kernel<<<grid_for(n_out, 256), 256, 0, stream>>>(nbr, pitch, n_out, int(kernel_volume), degree);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int fvc_neighbor_indexes(const FvcGridBatch *grid, const int32_t *query_ijk, const int32_t *query_bidx, int64_t nq,
                         int32_t extent, int32_t shift, int64_t *out, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    FVC_REQUIRE(grid, FVC_ERR_RUNTIME, "grid must be provided");
    FVC_REQUIRE(extent >= 0, FVC_ERR_VALUE, "extent must be >= 0");
    FVC_REQUIRE(shift >= 0 && shift < 31, FVC_ERR_VALUE, "bitshift must be in [0, 31)");
    if (nq == 0)
        return FVC_OK;
    const int64_t w = 2 * extent + 1;
    neighbor_indexes_kernel<<<grid_for(nq * w * w * w, 256), 256, 0, stream>>>(*grid, query_ijk, query_bidx, nq, extent,
                                                                              shift, out);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int fvc_ijk_to_index(const FvcGridBatch *grid, const int32_t *query_ijk, const int32_t *query_bidx, int64_t nq,
                     int32_t cumulative, int64_t *out, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    FVC_REQUIRE(grid, FVC_ERR_RUNTIME, "grid must be provided");
    if (nq == 0)
        return FVC_OK;
    ijk_to_index_kernel<<<grid_for(nq, 256), 256, 0, stream>>>(*grid, query_ijk, query_bidx, nq, cumulative, out);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int fvc_conv_grid_count(const int32_t *src_ijk, int64_t n, const int32_t kernel_size[3], const int32_t stride[3],
                        int32_t transposed, void *scratch8, int64_t *count_host, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_geometry_args(kernel_size, stride);
    if (rc)
        return rc;
    const Geometry g = make_geometry(kernel_size, stride);
    FVC_REQUIRE(g.volume <= INT32_MAX, FVC_ERR_RUNTIME, "kernel volume %lld exceeds INT32_MAX (BuildGridForConv.cu:151-152)",
                (long long)g.volume);
    *count_host = 0;
    if (n == 0)
        return FVC_OK;
    if (transposed) {
        *count_host = n * g.volume;
        return FVC_OK;
    }
    FVC_CUDA(cudaMemsetAsync(scratch8, 0, 8, stream));
    conv_grid_count_kernel<<<grid_for(n, 256), 256, 0, stream>>>(src_ijk, n, g,
                                                                 reinterpret_cast<unsigned long long *>(scratch8));
    FVC_LAUNCH_CHECK();
    unsigned long long total = 0;
    FVC_CUDA(cudaMemcpyAsync(&total, scratch8, 8, cudaMemcpyDeviceToHost, stream));
    FVC_CUDA(cudaStreamSynchronize(stream));
    *count_host = int64_t(total);
    return FVC_OK;
}

int fvc_conv_grid_emit(const int32_t *src_ijk, const int32_t *src_bidx, int64_t n, const int32_t kernel_size[3],
                       const int32_t stride[3], int32_t transposed, int64_t count, int32_t *cand_ijk,
                       int32_t *cand_bidx, void *counter8, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_geometry_args(kernel_size, stride);
    if (rc)
        return rc;
    const Geometry g = make_geometry(kernel_size, stride);
    if (n == 0 || count == 0)
        return FVC_OK;
    if (transposed) {
        FVC_REQUIRE(count == n * g.volume, FVC_ERR_RUNTIME, "transposed candidate count mismatch");
        conv_grid_emit_tr_kernel<<<grid_for(count, 256), 256, 0, stream>>>(src_ijk, src_bidx, n, g, cand_ijk, cand_bidx);
        FVC_LAUNCH_CHECK();
        return FVC_OK;
    }
    FVC_CUDA(cudaMemsetAsync(counter8, 0, 8, stream));
    conv_grid_emit_fwd_kernel<<<grid_for(n, 256), 256, 0, stream>>>(src_ijk, src_bidx, n, g, count, cand_ijk, cand_bidx,
                                                                   reinterpret_cast<unsigned long long *>(counter8));
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

} // extern "C"
