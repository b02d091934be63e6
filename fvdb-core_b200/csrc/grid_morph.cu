// grid_morph.cu -- stride-1 generated topologies by leaf-mask morphology (SURVEY.md 8(f) rank 1 "fast paths").
//
// conv_grid / conv_transpose_grid at stride 1 are the dilation of the source by a box of tap offsets
// [lo, hi]^3 (BuildGridForConv.cu:465-525: every tap of every voxel emits a voxel).  The general path stages K^3
// candidate coordinates per voxel and radix-sorts them (44 M keys for 1.6 M voxels at 3^3).  Here the work is per LEAF:
//   1. host: the <= 27 neighbour leaf origins of every source leaf go through the ordinary grid builder (27 keys per
//      leaf, ~50x fewer than per voxel), giving the output tree with placeholder leaves;
//   2. dilate_leaves_kernel: one warp per output leaf gathers the 3x3x3 source leaf masks into a 24^3-bit volume in
//      shared memory and applies the box dilation separably (z: shifts inside 24-bit lines, y and x: ORs of lines);
//   3. a scan of the leaf counts gives the row bases; expand_leaves_kernel writes base / prefix and the voxel list.
// Bit layout (NanoVDB leaf): word = x & 7, bit = (y & 7) * 8 + (z & 7); rows are ordered by (word, bit).
#include "fvc_common.cuh"

namespace fvc {

constexpr int GM_WARPS = 4;

__global__ void __launch_bounds__(GM_WARPS * 32)
dilate_leaves_kernel(FvcGridBatch src, FvcLeaf *__restrict__ dst, int n_dst, int lox, int hix, int loy, int hiy, int loz, int hiz,
                     int32_t *__restrict__ leaf_counts) {
    __shared__ uint32_t s_line[GM_WARPS][24][24]; // [x][y] -> 24 z-bits, volume origin = output leaf origin - 8
    __shared__ uint32_t s_tmp[GM_WARPS][24][8];   // after the y pass: [x][y - 8]
    __shared__ int s_leaf[GM_WARPS][27];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int leaf_id = blockIdx.x * GM_WARPS + warp;
    if (leaf_id >= n_dst)
        return;
    FvcLeaf *L = dst + leaf_id;
    const int b = L->batch;
    const int ox = L->origin[0], oy = L->origin[1], oz = L->origin[2];
    if (lane < 27) {
        const int dx = lane / 9 - 1, dy = (lane / 3) % 3 - 1, dz = lane % 3 - 1;
        s_leaf[warp][lane] = find_leaf(src, b, ox + 8 * dx, oy + 8 * dy, oz + 8 * dz);
    }
    for (int e = lane; e < 24 * 24; e += 32)
        (&s_line[warp][0][0])[e] = 0u;
    __syncwarp();
    // scatter the source masks: leaf (lx, ly, lz), word x, byte y -> 8 z-bits at bit 8 * lz of line [8 lx + x][8 ly + y]
    for (int e = lane; e < 27 * 8; e += 32) {
        const int li = e >> 3, x = e & 7;
        const int leaf = s_leaf[warp][li];
        if (leaf < 0)
            continue;
        const int lx = li / 9, ly = (li / 3) % 3, lz = li % 3;
        const uint64_t word = __ldg(src.leaves[leaf].mask + x);
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const uint32_t byte = uint32_t(word >> (8 * y)) & 0xFFu;
            if (byte)
                atomicOr(&s_line[warp][8 * lx + x][8 * ly + y], byte << (8 * lz)); // three z-leaves share a line
        }
    }
    __syncwarp();
    // z: out(z) = OR over o in [loz, hiz] of in(z - o)
    for (int e = lane; e < 24 * 24; e += 32) {
        const uint32_t v = (&s_line[warp][0][0])[e];
        uint32_t acc = 0u;
        for (int o = loz; o <= hiz; ++o)
            acc |= o >= 0 ? (v << o) : (v >> (-o));
        (&s_line[warp][0][0])[e] = acc & 0xFFFFFFu;
    }
    __syncwarp();
    // y: only the centre rows y in [8, 16) are needed from here on
    for (int e = lane; e < 24 * 8; e += 32) {
        const int x = e >> 3, y = 8 + (e & 7);
        uint32_t acc = 0u;
        for (int o = loy; o <= hiy; ++o)
            acc |= s_line[warp][x][y - o];
        s_tmp[warp][x][y - 8] = acc;
    }
    __syncwarp();
    // x, then pack: lane handles (x, y-pair); word x = OR over y of byte(x, y) << 8y
    uint64_t my_word = 0ull;
    {
        const int x = lane >> 2, y0 = (lane & 3) * 2; // 8 words x 4 lanes, two y rows per lane
        for (int yy = 0; yy < 2; ++yy) {
            uint32_t acc = 0u;
            for (int o = lox; o <= hix; ++o)
                acc |= s_tmp[warp][8 + x - o][y0 + yy];
            my_word |= uint64_t((acc >> 8) & 0xFFu) << (8 * (y0 + yy));
        }
    }
    my_word |= __shfl_xor_sync(0xffffffffu, my_word, 1);
    my_word |= __shfl_xor_sync(0xffffffffu, my_word, 2);
    // lanes 4x .. 4x+3 now hold word x; prefix = popcount of the words before it
    const int pc = __popcll(my_word);
    int incl = (lane & 3) == 0 ? pc : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
            incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if ((lane & 3) == 0) {
        L->mask[lane >> 2] = my_word;
        L->prefix[lane >> 2] = uint16_t(incl - pc);
    }
    if (lane == 0) {
        L->count = total;
        leaf_counts[leaf_id] = total;
    }
}

// base[leaf] -> leaf record, and the voxel list: row base + rank(word, bit) = origin + (x, y, z)
__global__ void __launch_bounds__(GM_WARPS * 32)
expand_leaves_kernel(FvcLeaf *__restrict__ leaves, int n_leaves, const int32_t *__restrict__ leaf_base, int32_t *__restrict__ out_ijk,
                     int32_t *__restrict__ out_bidx) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int leaf_id = blockIdx.x * GM_WARPS + warp;
    if (leaf_id >= n_leaves)
        return;
    FvcLeaf *L = leaves + leaf_id;
    const int base = leaf_base[leaf_id];
    if (lane == 0)
        L->base = base;
    const int b = L->batch, ox = L->origin[0], oy = L->origin[1], oz = L->origin[2];
    // lane scans 16 bits: word lane >> 2, bits 16 * (lane & 3) ..
    const int w = lane >> 2, shift = (lane & 3) * 16;
    const uint64_t m = L->mask[w];
    int r = base + int(L->prefix[w]) + __popcll(m & ((1ull << shift) - 1ull));
    for (uint32_t bits = uint32_t(m >> shift) & 0xFFFFu; bits; bits &= bits - 1u, ++r) {
        const int bit = shift + __ffs(bits) - 1;
        out_ijk[3 * int64_t(r)] = ox + w;
        out_ijk[3 * int64_t(r) + 1] = oy + (bit >> 3);
        out_ijk[3 * int64_t(r) + 2] = oz + (bit & 7);
        out_bidx[r] = b;
    }
}

} // namespace fvc

using namespace fvc;

extern "C" {

int fvc_grid_dilate_leaves(const FvcGridBatch *src, FvcLeaf *dst_leaves, int32_t n_dst_leaves, const int32_t lo[3], const int32_t hi[3],
                           int32_t *leaf_counts, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    FVC_REQUIRE(src && lo && hi, FVC_ERR_RUNTIME, "fvc_grid_dilate_leaves: null argument");
    for (int d = 0; d < 3; ++d)
        FVC_REQUIRE(lo[d] <= 0 && hi[d] >= 0 && lo[d] > -8 && hi[d] < 8, FVC_ERR_VALUE,
                    "fvc_grid_dilate_leaves: offsets [%d, %d] on axis %d must contain 0 and stay inside one leaf (|o| < 8)", lo[d], hi[d], d);
    if (n_dst_leaves == 0)
        return FVC_OK;
    FVC_REQUIRE(dst_leaves && leaf_counts, FVC_ERR_RUNTIME, "fvc_grid_dilate_leaves: null output");
    dilate_leaves_kernel<<<int(ceil_div(n_dst_leaves, GM_WARPS)), GM_WARPS * 32, 0, stream>>>(*src, dst_leaves, n_dst_leaves, lo[0], hi[0], lo[1], hi[1],
                                                                                            lo[2], hi[2], leaf_counts);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int fvc_grid_expand_leaves(FvcLeaf *leaves, int32_t n_leaves, const int32_t *leaf_base, int32_t *out_ijk, int32_t *out_bidx, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n_leaves == 0)
        return FVC_OK;
    FVC_REQUIRE(leaves && leaf_base, FVC_ERR_RUNTIME, "fvc_grid_expand_leaves: null argument");
    expand_leaves_kernel<<<int(ceil_div(n_leaves, GM_WARPS)), GM_WARPS * 32, 0, stream>>>(leaves, n_leaves, leaf_base, out_ijk, out_bidx);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

} // extern "C"
