// grid_build.cu -- batched (grid, ijk) -> index grid on device.
//
// Replaces ops/BuildGridFromIjk.cu:52-111 (NanoVDB voxelsToGrid + mergeGridHandles, looped over the
// batch on the host there; one batched pass here).  Row order = lexicographic in
// (grid, root tile, upper offset, lower offset, leaf offset), each x-major: a 106-bit key sorted with
// two stable LSD radix passes (low word, then high word).
#include "fvc_common.cuh"

#include <cub/cub.cuh>

namespace fvc {

struct Add4 {
    __host__ __device__ __forceinline__ int4 operator()(const int4 &a, const int4 &b) const {
        return make_int4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
};

struct TileRange { // smallest / largest root-tile coordinate of the batch per axis (see "single-pass key" below)
    int lo[3], hi[3];
};

struct BuildScratch {
    uint64_t *key_a, *key_b;
    uint32_t *idx_a, *idx_b;
    int4 *scan;
    TileRange *range;
    void *cub_temp;
    size_t cub_bytes;
    size_t total;
};

static size_t cub_temp_bytes(int64_t n) {
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DoubleBuffer<uint64_t> keys(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> vals(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys, vals, n, 0, 64);
    cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, (int4 *)nullptr, (int4 *)nullptr, Add4(), n);
    return sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
}

static BuildScratch carve(void *scratch, int64_t n) {
    BuildScratch s;
    const size_t m = size_t(n > 0 ? n : 1);
    char *p = reinterpret_cast<char *>(scratch);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char *q = p ? p + off : nullptr;
        off += align_up(bytes, 256);
        return q;
    };
    s.key_a = reinterpret_cast<uint64_t *>(take(m * 8));
    s.key_b = reinterpret_cast<uint64_t *>(take(m * 8));
    s.idx_a = reinterpret_cast<uint32_t *>(take(m * 4));
    s.idx_b = reinterpret_cast<uint32_t *>(take(m * 4));
    s.scan = reinterpret_cast<int4 *>(take(m * 16 + 16));
    s.range = reinterpret_cast<TileRange *>(take(sizeof(TileRange)));
    s.cub_bytes = cub_temp_bytes(n > 0 ? n : 1);
    s.cub_temp = take(s.cub_bytes);
    s.total = off;
    return s;
}

// low word: [ty & 255 : 8][tz : 20][upper offset : 15][lower offset : 12][leaf offset : 9]
__device__ __forceinline__ uint64_t key_lo(int x, int y, int z) {
    const uint64_t ty = uint64_t((y >> 12) + (1 << 19)), tz = uint64_t((z >> 12) + (1 << 19));
    const uint64_t up = uint64_t(((((x >> 7) & 31) << 5) | ((y >> 7) & 31)) << 5 | ((z >> 7) & 31));
    const uint64_t lo = uint64_t(((((x >> 3) & 15) << 4) | ((y >> 3) & 15)) << 4 | ((z >> 3) & 15));
    const uint64_t vx = uint64_t(((x & 7) << 6) | ((y & 7) << 3) | (z & 7));
    return ((ty & 255ull) << 56) | (tz << 36) | (up << 21) | (lo << 9) | vx;
}
// high word (42 bits): [grid : 10][tx : 20][ty >> 8 : 12]
__device__ __forceinline__ uint64_t key_hi(int b, int x, int y) {
    const uint64_t tx = uint64_t((x >> 12) + (1 << 19)), ty = uint64_t((y >> 12) + (1 << 19));
    return (uint64_t(b) << 32) | (tx << 12) | (ty >> 8);
}

// ---- single-pass key: when every grid of the batch spans at most 64 root tiles per axis (262 144 voxels) the whole order
// fits ONE 64-bit key  [grid : 10][tx - tx_min : 6][ty - ty_min : 6][tz - tz_min : 6][upper : 15][lower : 12][leaf : 9]
// (subtracting the batch-wide minimum tile keeps the lexicographic order), i.e. one 8-digit radix sort instead of 8 + 6 digits
// and one encode pass instead of two.  The tile range is found on the device; whether it fits is read back with the counts
// the build synchronises on anyway, and the rare wide batch re-runs the two-pass path.
__global__ void tile_range_init_kernel(TileRange *r) {
    if (threadIdx.x < 3) {
        r->lo[threadIdx.x] = INT32_MAX;
        r->hi[threadIdx.x] = INT32_MIN;
    }
}

__global__ void tile_range_kernel(const int32_t *__restrict__ ijk, int64_t n, TileRange *__restrict__ range) {
    int lo[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, hi[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int t = ijk[3 * i + d] >> 12;
            lo[d] = t < lo[d] ? t : lo[d];
            hi[d] = t > hi[d] ? t : hi[d];
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const int a = __shfl_xor_sync(0xffffffffu, lo[d], o), b = __shfl_xor_sync(0xffffffffu, hi[d], o);
            lo[d] = a < lo[d] ? a : lo[d];
            hi[d] = b > hi[d] ? b : hi[d];
        }
        if ((threadIdx.x & 31) == 0 && lo[d] <= hi[d]) {
            atomicMin(&range->lo[d], lo[d]);
            atomicMax(&range->hi[d], hi[d]);
        }
    }
}

__global__ void encode64_kernel(const int32_t *__restrict__ ijk, const int32_t *__restrict__ bidx, int64_t n, const TileRange *__restrict__ range,
                                uint64_t *__restrict__ keys, uint32_t *__restrict__ idx) {
    const int t0 = range->lo[0], t1 = range->lo[1], t2 = range->lo[2];
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const int x = ijk[3 * i], y = ijk[3 * i + 1], z = ijk[3 * i + 2];
        const uint64_t tx = uint64_t((x >> 12) - t0) & 63ull, ty = uint64_t((y >> 12) - t1) & 63ull, tz = uint64_t((z >> 12) - t2) & 63ull;
        const uint64_t up = uint64_t(((((x >> 7) & 31) << 5) | ((y >> 7) & 31)) << 5 | ((z >> 7) & 31));
        const uint64_t lo = uint64_t(((((x >> 3) & 15) << 4) | ((y >> 3) & 15)) << 4 | ((z >> 3) & 15));
        const uint64_t vx = uint64_t(((x & 7) << 6) | ((y & 7) << 3) | (z & 7));
        keys[i] = (uint64_t(bidx ? bidx[i] : 0) << 54) | (tx << 48) | (ty << 42) | (tz << 36) | (up << 21) | (lo << 9) | vx;
        idx[i] = uint32_t(i);
    }
}

__global__ void encode_lo_kernel(const int32_t *__restrict__ ijk, int64_t n, uint64_t *__restrict__ keys,
                                 uint32_t *__restrict__ idx) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        keys[i] = key_lo(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]);
        idx[i] = uint32_t(i);
    }
}

__global__ void encode_hi_kernel(const int32_t *__restrict__ ijk, const int32_t *__restrict__ bidx, int64_t n,
                                 const uint32_t *__restrict__ idx, uint64_t *__restrict__ keys) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t s = idx[i];
        keys[i] = key_hi(bidx ? bidx[s] : 0, ijk[3 * s], ijk[3 * s + 1]);
    }
}

// head flags of the sorted sequence: x = new voxel, y = new leaf, z = new lower node, w = new root tile
__global__ void head_flags_kernel(const int32_t *__restrict__ ijk, const int32_t *__restrict__ bidx, int64_t n,
                                  const uint32_t *__restrict__ perm, int4 *__restrict__ flags) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        int4 f = make_int4(1, 1, 1, 1);
        if (i > 0) {
            const int64_t a = perm[i], p = perm[i - 1];
            const int xa = ijk[3 * a], ya = ijk[3 * a + 1], za = ijk[3 * a + 2];
            const int xp = ijk[3 * p], yp = ijk[3 * p + 1], zp = ijk[3 * p + 2];
            const bool same_b = !bidx || bidx[a] == bidx[p];
            const int dx = xa ^ xp, dy = ya ^ yp, dz = za ^ zp;
            const int d = dx | dy | dz;
            f.x = !(same_b && d == 0);
            f.y = !(same_b && (d >> 3) == 0);
            f.z = !(same_b && (d >> 7) == 0);
            f.w = !(same_b && (d >> 12) == 0);
        }
        flags[i] = f;
    }
}

__global__ void fill_nodes_kernel(const int32_t *__restrict__ ijk, const int32_t *__restrict__ bidx, int64_t n,
                                  int32_t num_grids, const uint32_t *__restrict__ perm, const int4 *__restrict__ scan,
                                  FvcLeaf *__restrict__ leaves, int32_t *__restrict__ lower, int32_t *__restrict__ upper,
                                  int32_t *__restrict__ root_keys, int32_t *__restrict__ root_offsets,
                                  int64_t *__restrict__ voxel_offsets, int32_t *__restrict__ leaf_offsets,
                                  int32_t *__restrict__ out_ijk, int32_t *__restrict__ out_bidx) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const int4 cur = scan[i];
        const int4 prev = i > 0 ? scan[i - 1] : make_int4(0, 0, 0, 0);
        const int64_t src = perm[i];
        const int x = ijk[3 * src], y = ijk[3 * src + 1], z = ijk[3 * src + 2];
        const int b = bidx ? bidx[src] : 0;
        const int row = cur.x - 1, leaf = cur.y - 1, low = cur.z - 1, up = cur.w - 1;
        if (cur.x != prev.x) { // first occurrence of this voxel
            out_ijk[3 * int64_t(row)] = x;
            out_ijk[3 * int64_t(row) + 1] = y;
            out_ijk[3 * int64_t(row) + 2] = z;
            out_bidx[row] = b;
            atomicOr(reinterpret_cast<unsigned long long *>(&leaves[leaf].mask[x & 7]),
                     1ull << (((y & 7) << 3) | (z & 7)));
        }
        if (cur.y != prev.y) { // first voxel of a leaf
            FvcLeaf *L = leaves + leaf;
            L->base = row;
            L->batch = b;
            L->origin[0] = x & ~7;
            L->origin[1] = y & ~7;
            L->origin[2] = z & ~7;
            lower[(int64_t(low) << 12) + (((((x >> 3) & 15) << 4) | ((y >> 3) & 15)) << 4 | ((z >> 3) & 15))] = leaf;
        }
        if (cur.z != prev.z)
            upper[(int64_t(up) << 15) + (((((x >> 7) & 31) << 5) | ((y >> 7) & 31)) << 5 | ((z >> 7) & 31))] = low;
        if (cur.w != prev.w) {
            root_keys[4 * up] = b;
            root_keys[4 * up + 1] = x >> 12;
            root_keys[4 * up + 2] = y >> 12;
            root_keys[4 * up + 3] = z >> 12;
        }
        // per-grid offsets: this element starts grid b if the previous one belongs to an earlier grid
        const int bprev = i > 0 ? (bidx ? bidx[perm[i - 1]] : 0) : -1;
        for (int g = bprev + 1; g <= b; ++g) {
            voxel_offsets[g] = row;
            leaf_offsets[g] = leaf;
            root_offsets[g] = up;
        }
        if (i == n - 1) {
            for (int g = b + 1; g <= num_grids; ++g) {
                voxel_offsets[g] = cur.x;
                leaf_offsets[g] = cur.y;
                root_offsets[g] = cur.w;
            }
        }
    }
}

__global__ void finalize_leaves_kernel(FvcLeaf *__restrict__ leaves, int32_t num_leaves) {
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= num_leaves)
        return;
    FvcLeaf *L = leaves + leaf;
    int running = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        L->prefix[w] = uint16_t(running);
        running += __popcll(L->mask[w]);
    }
    L->count = running;
}

__global__ void zero_offsets_kernel(int32_t num_grids, int32_t *root_offsets, int64_t *voxel_offsets,
                                    int32_t *leaf_offsets) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g <= num_grids) {
        root_offsets[g] = 0;
        voxel_offsets[g] = 0;
        leaf_offsets[g] = 0;
    }
}

static inline int grid_for(int64_t n, int block) {
    int64_t blocks = ceil_div(n, block);
    return int(blocks < 1 ? 1 : (blocks > 148 * 16 ? 148 * 16 : blocks));
}

} // namespace fvc

using namespace fvc;

extern "C" {

size_t fvc_grid_build_scratch_bytes(int64_t n) { return carve(nullptr, n).total; }

int fvc_grid_build_count(const int32_t *ijk, const int32_t *bidx, int64_t n, int32_t num_grids, void *scratch,
                         size_t scratch_bytes, int64_t counts_host[4], fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    FVC_REQUIRE(n >= 0 && n <= INT32_MAX, FVC_ERR_RUNTIME, "ijk has %lld rows, exceeding the int32 index limit",
                (long long)n);
    FVC_REQUIRE(num_grids >= 0 && num_grids <= 1024, FVC_ERR_RUNTIME,
                "batch size %d exceeds the 1024-grid limit (GridBatchData.h:26)", num_grids);
    for (int d = 0; d < 4; ++d)
        counts_host[d] = 0;
    if (n == 0)
        return FVC_OK;
    BuildScratch s = carve(scratch, n);
    FVC_REQUIRE(scratch && scratch_bytes >= s.total, FVC_ERR_RUNTIME, "grid build scratch too small: %zu < %zu",
                scratch_bytes, s.total);
    const int block = 256, grid = grid_for(n, block);
    tile_range_init_kernel<<<1, 32, 0, stream>>>(s.range);
    FVC_LAUNCH_CHECK();
    tile_range_kernel<<<grid, block, 0, stream>>>(ijk, n, s.range);
    FVC_LAUNCH_CHECK();
    auto two_pass = [&]() -> int {
        encode_lo_kernel<<<grid, block, 0, stream>>>(ijk, n, s.key_a, s.idx_a);
        FVC_LAUNCH_CHECK();
        cub::DoubleBuffer<uint64_t> keys(s.key_a, s.key_b);
        cub::DoubleBuffer<uint32_t> vals(s.idx_a, s.idx_b);
        size_t temp = s.cub_bytes;
        FVC_CUDA(cub::DeviceRadixSort::SortPairs(s.cub_temp, temp, keys, vals, n, 0, 64, stream));
        g_launch_count.fetch_add(1);
        uint32_t *idx_sorted = vals.Current();
        uint32_t *idx_other = vals.Alternate();
        // second (most significant) pass; keys are regenerated from the permuted inputs
        encode_hi_kernel<<<grid, block, 0, stream>>>(ijk, bidx, n, idx_sorted, s.key_a);
        FVC_LAUNCH_CHECK();
        cub::DoubleBuffer<uint64_t> keys2(s.key_a, s.key_b);
        cub::DoubleBuffer<uint32_t> vals2(idx_sorted, idx_other);
        temp = s.cub_bytes;
        FVC_CUDA(cub::DeviceRadixSort::SortPairs(s.cub_temp, temp, keys2, vals2, n, 0, 42, stream));
        g_launch_count.fetch_add(1);
        // keep the final permutation in idx_a so that stage 2 finds it
        if (vals2.Current() != s.idx_a)
            FVC_CUDA(cudaMemcpyAsync(s.idx_a, vals2.Current(), size_t(n) * 4, cudaMemcpyDeviceToDevice, stream));
        return FVC_OK;
    };
    auto one_pass = [&]() -> int {
        encode64_kernel<<<grid, block, 0, stream>>>(ijk, bidx, n, s.range, s.key_a, s.idx_a);
        FVC_LAUNCH_CHECK();
        cub::DoubleBuffer<uint64_t> keys(s.key_a, s.key_b);
        cub::DoubleBuffer<uint32_t> vals(s.idx_a, s.idx_b);
        size_t temp = s.cub_bytes;
        FVC_CUDA(cub::DeviceRadixSort::SortPairs(s.cub_temp, temp, keys, vals, n, 0, 64, stream));
        g_launch_count.fetch_add(1);
        if (vals.Current() != s.idx_a)
            FVC_CUDA(cudaMemcpyAsync(s.idx_a, vals.Current(), size_t(n) * 4, cudaMemcpyDeviceToDevice, stream));
        return FVC_OK;
    };
    auto flags_and_counts = [&](int4 *totals, TileRange *range_host) -> int {
        head_flags_kernel<<<grid, block, 0, stream>>>(ijk, bidx, n, s.idx_a, s.scan);
        FVC_LAUNCH_CHECK();
        size_t temp = s.cub_bytes;
        FVC_CUDA(cub::DeviceScan::InclusiveScan(s.cub_temp, temp, s.scan, s.scan, Add4(), n, stream));
        g_launch_count.fetch_add(1);
        FVC_CUDA(cudaMemcpyAsync(totals, s.scan + (n - 1), sizeof(int4), cudaMemcpyDeviceToHost, stream));
        if (range_host)
            FVC_CUDA(cudaMemcpyAsync(range_host, s.range, sizeof(TileRange), cudaMemcpyDeviceToHost, stream));
        FVC_CUDA(cudaStreamSynchronize(stream));
        return FVC_OK;
    };
    int4 totals;
    TileRange range_host;
    int rc = one_pass();
    if (rc)
        return rc;
    rc = flags_and_counts(&totals, &range_host);
    if (rc)
        return rc;
    bool narrow = true;
    for (int d = 0; d < 3; ++d)
        narrow = narrow && (int64_t(range_host.hi[d]) - range_host.lo[d] < 64);
    if (!narrow) { // a batch wider than 64 root tiles on some axis: the 64-bit key wrapped -- redo with the 106-bit key
        rc = two_pass();
        if (rc)
            return rc;
        rc = flags_and_counts(&totals, nullptr);
        if (rc)
            return rc;
    }
    counts_host[0] = totals.x;
    counts_host[1] = totals.y;
    counts_host[2] = totals.z;
    counts_host[3] = totals.w;
    return FVC_OK;
}

int fvc_grid_build_fill(const int32_t *ijk, const int32_t *bidx, int64_t n, int32_t num_grids, void *scratch,
                        size_t scratch_bytes, const int64_t counts_host[4], FvcLeaf *leaves, int32_t *lower,
                        int32_t *upper, int32_t *root_keys, int32_t *root_offsets, int64_t *voxel_offsets,
                        int32_t *leaf_offsets, int32_t *out_ijk, int32_t *out_bidx, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    FVC_REQUIRE(root_offsets && voxel_offsets && leaf_offsets, FVC_ERR_RUNTIME, "offset arrays must be provided");
    if (n == 0 || counts_host[0] == 0) {
        zero_offsets_kernel<<<ceil_div(num_grids + 1, 256), 256, 0, stream>>>(num_grids, root_offsets, voxel_offsets,
                                                                             leaf_offsets);
        FVC_LAUNCH_CHECK();
        return FVC_OK;
    }
    BuildScratch s = carve(scratch, n);
    FVC_REQUIRE(scratch && scratch_bytes >= s.total, FVC_ERR_RUNTIME, "grid build scratch too small: %zu < %zu",
                scratch_bytes, s.total);
    FVC_REQUIRE((reinterpret_cast<uintptr_t>(leaves) & 127) == 0, FVC_ERR_RUNTIME, "leaves must be 128-byte aligned");
    FVC_CUDA(cudaMemsetAsync(leaves, 0, size_t(counts_host[1]) * sizeof(FvcLeaf), stream));
    FVC_CUDA(cudaMemsetAsync(lower, 0xFF, size_t(counts_host[2]) * 4096 * 4, stream));
    FVC_CUDA(cudaMemsetAsync(upper, 0xFF, size_t(counts_host[3]) * 32768 * 4, stream));
    const int block = 256, grid = grid_for(n, block);
    fill_nodes_kernel<<<grid, block, 0, stream>>>(ijk, bidx, n, num_grids, s.idx_a, s.scan, leaves, lower, upper,
                                                  root_keys, root_offsets, voxel_offsets, leaf_offsets, out_ijk,
                                                  out_bidx);
    FVC_LAUNCH_CHECK();
    finalize_leaves_kernel<<<int(ceil_div(counts_host[1], 256)), 256, 0, stream>>>(leaves, int32_t(counts_host[1]));
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

} // extern "C"
