/*
 * fvdbconv.h -- C ABI of the B200-native sparse-convolution engine (libfvdbconv.so).
 *
 * This library is the drop-in for the convolution section of fVDB's pybind11 module
 * `fvdb._fvdb_cpp` (reference: src/python/Bindings.cpp:491-674 and src/python/GridBatchOps.cpp:764-792)
 * and for the C++ ops those bindings call (src/fvdb/detail/ops/convolution/GatherScatterDefault.{h,cu},
 * PredGatherIGemm.{h,cu}, ops/BuildGridForConv*.cu, ops/BuildGridFromIjk.cu, ops/NeighborIndexes.cu).
 * Each entry point names the reference interface it replaces.
 *
 * Conventions
 *   - plain C, no torch types: raw device pointers (tensor.data_ptr()), sizes, an explicit stream;
 *   - the library never allocates or frees user-visible memory: outputs and scratch are allocated by
 *     the caller (PyTorch's caching allocator owns everything);
 *   - every call is asynchronous on `stream` unless the comment says it synchronises (only the calls
 *     that must return a count to size the caller's next allocation do);
 *   - return value 0 = success, non-zero = failure; fvc_last_error() gives the thread-local message.
 *     The Python shim maps FVC_ERR_VALUE -> ValueError, FVC_ERR_INDEX -> IndexError, others ->
 *     RuntimeError, as TORCH_CHECK_VALUE / TORCH_CHECK_INDEX / TORCH_CHECK do in the reference;
 *   - there is no CPU fallback: every compute entry point needs a CUDA device of compute capability 10.x.
 */
#ifndef FVDBCONV_H
#define FVDBCONV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FVC_ABI_VERSION 3

#if defined(__GNUC__)
#define FVC_API __attribute__((visibility("default")))
#else
#define FVC_API
#endif

/* status codes */
#define FVC_OK 0
#define FVC_ERR_VALUE 1       /* invalid argument value   (reference: TORCH_CHECK_VALUE -> ValueError)   */
#define FVC_ERR_RUNTIME 2     /* precondition violated    (reference: TORCH_CHECK       -> RuntimeError) */
#define FVC_ERR_INDEX 3       /* index out of range       (reference: TORCH_CHECK_INDEX -> IndexError)   */
#define FVC_ERR_CUDA 4        /* CUDA runtime failure                                                    */
#define FVC_ERR_UNSUPPORTED 5 /* no kernel for this dtype / shape (reference: dispatch_lookup_error)     */

/* dtype codes (torch.float16 / bfloat16 / float32 / float64) */
#define FVC_F16 0
#define FVC_BF16 1
#define FVC_F32 2
#define FVC_F64 3

typedef void *fvc_stream_t; /* cudaStream_t */

/* -------------------------------------------------------------------------------------------------
 * Index grid (replaces the NanoVDB OnIndexGrid batch held by fvdb::GridBatchData,
 * reference: src/fvdb/GridBatchData.h:29-55,120-226).
 *
 * Tree shape follows NanoVDB: root (sorted table of 4096^3 tiles per grid) -> upper node 32^3 ->
 * lower node 16^3 -> leaf 8^3.  One leaf is one 128-byte record = one L2 line:
 *   mask[w] bit (y&7)*8+(z&7) for w = x&7;  prefix[w] = number of active voxels in words < w;
 *   value(x,y,z) = base + prefix[w] + popcount(mask[w] below the bit)  (0-based, batch-cumulative row).
 * Voxel rows are ordered by (grid, root tile, upper offset, lower offset, leaf offset), each x-major.
 * ------------------------------------------------------------------------------------------------- */
typedef struct FvcLeaf {
    uint64_t mask[8];
    uint16_t prefix[8];
    int32_t base;      /* batch-cumulative row of the leaf's first active voxel */
    int32_t batch;     /* grid index inside the batch */
    int32_t origin[3]; /* ijk of the leaf's (0,0,0) corner, multiples of 8 */
    int32_t count;     /* active voxels in this leaf */
    int32_t reserved[6];
} FvcLeaf; /* sizeof == 128 */

typedef struct FvcGridBatch {
    int32_t num_grids;
    int32_t num_leaves;
    int32_t num_lower;
    int32_t num_upper; /* == number of root tiles */
    int64_t total_voxels;
    const FvcLeaf *leaves;        /* [num_leaves], 128-byte aligned                                  */
    const int32_t *lower;         /* [num_lower][4096]   leaf index or -1                            */
    const int32_t *upper;         /* [num_upper][32768]  lower-node index or -1                      */
    const int32_t *root_keys;     /* [num_upper][4]      (grid, x>>12, y>>12, z>>12), sorted         */
    const int32_t *root_offsets;  /* [num_grids+1]       range of root tiles owned by each grid      */
    const int64_t *voxel_offsets; /* [num_grids+1]       cumulative voxel count (JaggedTensor joffsets) */
    const int32_t *leaf_offsets;  /* [num_grids+1]       cumulative leaf count                       */
} FvcGridBatch;

/* -------- library / device ------------------------------------------------------------------- */
FVC_API int fvc_abi_version(void);
FVC_API const char *fvc_last_error(void);
/* Fills sm_count / cc_major / cc_minor of the current device; fails (FVC_ERR_CUDA) without a GPU. */
FVC_API int fvc_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* Number of kernels this library has launched in the calling process (for bench.py's gpu_launches). */
FVC_API int64_t fvc_launch_count(void);

/* Which kernel family serves (dtype, channels, kernel volume) under `path` (0 auto / 1 CUDA-core / 2 tensor-core):
 * returns 1 = CUDA-core kernels, 2 = tcgen05 kernels.  pass 0 = forward / dgrad, 1 = weight gradient (given a dense map),
 * 2 = the fused backward (fvc_conv_backward_fused; cin / cout = the public weight dimensions). */
FVC_API int32_t fvc_conv_kernel_family(int32_t cin, int32_t cout, int64_t kernel_volume, int32_t dtype, int32_t path, int32_t pass);
/* Experiment knob for the benchmark scripts: key 0 = pipeline-shape variant of the tensor-core forward kernel, key 1 = of the
 * weight-gradient kernel (0 = the shape table's default).  Not part of the reference interface. */
FVC_API int fvc_set_tuning(int32_t key, int32_t value);

/* -------- geometry (host-only; replaces ConvolutionGeometry.h:30-207) ----------------------------- */
/* Validates kernel_size / stride (> 0, volume fits int64) and returns padding_before / padding_after /
 * kernel_volume.  FVC_ERR_VALUE mirrors TORCH_CHECK_VALUE at ConvolutionGeometry.h:150-172,186-198. */
FVC_API int fvc_geometry(const int32_t kernel_size[3], const int32_t stride[3], int32_t padding_before[3],
                 int32_t padding_after[3], int64_t *kernel_volume);
/* tapCoord (ConvolutionGeometry.h:85-90), fineFromCoarse (:99-104), coarseFromFine (:107-124; returns
 * *divisible = 0 and leaves coarse untouched when an axis does not divide). */
FVC_API int fvc_geometry_tap_coord(const int32_t kernel_size[3], int64_t tap_index, int32_t tap[3]);
FVC_API int fvc_geometry_fine_from_coarse(const int32_t kernel_size[3], const int32_t stride[3], const int32_t coarse[3],
                                  const int32_t tap[3], int32_t fine[3]);
FVC_API int fvc_geometry_coarse_from_fine(const int32_t kernel_size[3], const int32_t stride[3], const int32_t fine[3],
                                  const int32_t tap[3], int32_t coarse[3], int32_t *divisible);

/* -------- index-grid construction (replaces ops/BuildGridFromIjk.cu:52-111, NanoVDB voxelsToGrid) -- */
/* Scratch bytes needed by fvc_grid_build_* for n input coordinates. */
FVC_API size_t fvc_grid_build_scratch_bytes(int64_t n);
/* Stage 1: sort the (grid, ijk) keys into index-grid row order, mark duplicates, count nodes.
 * ijk: int32 [n][3]; bidx: int32 [n] grid index of each row (NULL = all grid 0); coordinates may be
 * negative and duplicated.  counts_host[0..3] = {unique voxels, leaves, lower nodes, upper nodes}.
 * SYNCHRONISES the stream (the caller sizes the node arrays from the counts). */
FVC_API int fvc_grid_build_count(const int32_t *ijk, const int32_t *bidx, int64_t n, int32_t num_grids, void *scratch,
                         size_t scratch_bytes, int64_t counts_host[4], fvc_stream_t stream);
/* Stage 2: fill caller-allocated node arrays (sized from counts_host) and the row-ordered ijk table.
 * `scratch` must be the buffer stage 1 wrote; the call initialises every output itself.  out_ijk: int32 [unique][3]; out_bidx: int32 [unique] (JaggedTensor jidx). */
FVC_API int fvc_grid_build_fill(const int32_t *ijk, const int32_t *bidx, int64_t n, int32_t num_grids, void *scratch,
                        size_t scratch_bytes, const int64_t counts_host[4], FvcLeaf *leaves, int32_t *lower,
                        int32_t *upper, int32_t *root_keys, int32_t *root_offsets, int64_t *voxel_offsets,
                        int32_t *leaf_offsets, int32_t *out_ijk, int32_t *out_bidx, fvc_stream_t stream);

/* -------- generated target topologies (replaces conv_grid / conv_transpose_grid,
 *          src/python/GridBatchOps.cpp:764-792, ops/BuildGridForConv.cu:392-544,
 *          ops/BuildGridForConvTranspose.cu:201-359) ------------------------------------------- */
/* Number of candidate target coordinates the source grid emits (forward: one per divisible tap per
 * voxel; transposed: voxels * kernel volume).  SYNCHRONISES (forward only). */
FVC_API int fvc_conv_grid_count(const int32_t *src_ijk, int64_t n, const int32_t kernel_size[3], const int32_t stride[3],
                        int32_t transposed, void *scratch8, int64_t *count_host, fvc_stream_t stream);
/* Emits the candidates (unsorted, with duplicates) into cand_ijk [count][3] / cand_bidx [count]; feed
 * them to fvc_grid_build_*.  counter8: 8 bytes of zero-initialised-by-the-call device scratch. */
FVC_API int fvc_conv_grid_emit(const int32_t *src_ijk, const int32_t *src_bidx, int64_t n, const int32_t kernel_size[3],
                       const int32_t stride[3], int32_t transposed, int64_t count, int32_t *cand_ijk,
                       int32_t *cand_bidx, void *counter8, fvc_stream_t stream);

/* -------- kernel map (replaces gs_build_topology / gs_build_transpose_topology,
 *          Bindings.cpp:570-583,612-625 -> GatherScatterDefault.cu:92-271) --------------------- */
/* Output-stationary, tap-major dense map: nbr[k * pitch + o] = feature row reached from output row o
 * through tap k, or -1.  Forward probes fineFromCoarse on the feature grid, transposed probes
 * coarseFromFine with the divisibility test (GatherScatterDefault.cu:129-141); probes never leave the
 * output voxel's own grid (:126).  tap_counts: int64 [K^3] device, number of pairs per tap.
 * FVC_ERR_RUNTIME when either grid exceeds INT32_MAX voxels or batch sizes differ (:58-80).
 * tile_mask (may be NULL): uint64 [ceil(N_out / 128)][ceil(K^3 / 64)], the per-tile tap bitmask of fvc_kmap_tile_mask written in
 * the same pass (zero-initialised by the call).  Asynchronous: no host synchronisation. */
FVC_API int fvc_kmap_build(const FvcGridBatch *feature_grid, const FvcGridBatch *output_grid, const int32_t kernel_size[3],
                   const int32_t stride[3], int32_t transposed, int32_t *nbr, int64_t pitch, int64_t *tap_counts, uint64_t *tile_mask,
                   fvc_stream_t stream);
/* CSR-by-tap view of a dense map (GatherScatterDefaultTopology, GatherScatterDefault.h:59-81): per tap
 * segment, pairs ordered by output row.  offsets_dev: int64 [K^3+1] = exclusive scan of tap_counts
 * (written by the call); gather / scatter: int32 [total pairs].  scratch: fvc_kmap_csr_scratch_bytes. */
FVC_API size_t fvc_kmap_csr_scratch_bytes(int64_t n_out, int64_t kernel_volume);
FVC_API int fvc_kmap_to_csr(const int32_t *nbr, int64_t pitch, int64_t n_out, int64_t kernel_volume,
                    const int64_t *tap_counts, int64_t *offsets_dev, int32_t *gather, int32_t *scatter, void *scratch,
                    size_t scratch_bytes, fvc_stream_t stream);
/* Dense map of the reversed rulebook from a CSR view: nbr_rev[k * pitch_rev + gather[p]] = scatter[p]
 * (the input-stationary map dgrad consumes; GatherScatterDefault.cu:794-804).  Fills -1 first. */
FVC_API int fvc_kmap_reverse_dense(const int32_t *gather, const int32_t *scatter, const int64_t *offsets_dev,
                           int64_t kernel_volume, int64_t total_pairs, int64_t n_feature, int32_t *nbr_rev,
                           int64_t pitch_rev, fvc_stream_t stream);
/* The same reversed dense map straight from the output-stationary dense map (no CSR needed): nbr_rev[k][nbr[k][o]] = o. */
FVC_API int fvc_kmap_reverse_from_dense(const int32_t *nbr, int64_t pitch, int64_t n_out, int64_t kernel_volume, int64_t n_feature,
                                int32_t *nbr_rev, int64_t pitch_rev, fvc_stream_t stream);
/* Per 128-row tile of a dense map, a bitmask over taps: bit k of mask[tile * words + (k >> 6)] (words =
 * ceil(K^3 / 64), bit index k & 63) is set iff some row of the tile has a neighbour through tap k.  The
 * tensor-core executors use it to skip whole (tile, tap) units: on planar surfaces two thirds of them are empty. */
FVC_API int fvc_kmap_tile_mask(const int32_t *nbr, int64_t pitch, int64_t n_out, int64_t kernel_volume, uint64_t *mask,
                               fvc_stream_t stream);
/* Per-row degree (number of taps that hit) of a dense map: degree int32 [n_out]. */
FVC_API int fvc_kmap_degree(const int32_t *nbr, int64_t pitch, int64_t n_out, int64_t kernel_volume, int32_t *degree,
                    fvc_stream_t stream);

/* -------- lookups (replaces ops/NeighborIndexes.cu:22-121, ops/IjkToIndex.cu:30-40) ------------- */
/* out: int64 [nq][w][w][w], w = 2*extent+1, x-major offsets in [-extent, extent]; per-grid-local index
 * or -1; query ijk is shifted left by `shift` first (NeighborIndexes.cu:35-44). */
FVC_API int fvc_neighbor_indexes(const FvcGridBatch *grid, const int32_t *query_ijk, const int32_t *query_bidx, int64_t nq,
                         int32_t extent, int32_t shift, int64_t *out, fvc_stream_t stream);
/* out: int64 [nq]; per-grid-local (cumulative = 0) or batch-cumulative (cumulative = 1) row, or -1. */
FVC_API int fvc_ijk_to_index(const FvcGridBatch *grid, const int32_t *query_ijk, const int32_t *query_bidx, int64_t nq,
                     int32_t cumulative, int64_t *out, fvc_stream_t stream);

/* -------- weights ----------------------------------------------------------------------------------
 * Replaces `weights.permute({2,3,4,1,0}).reshape({K,Cin,Cout}).contiguous()` (+ cast)
 * (GatherScatterDefault.cu:691-694,762-765).  Reads public-layout weights [Cout,Cin,k0,k1,k2] through
 * arbitrary element strides (fvdb.nn stores them as a permuted view, fvdb/nn/modules.py:282-289).
 *   layout 0: out[k][ci][co]   (forward:  Y  = X  . W[k])
 *   layout 1: out[k][co][ci]   (dgrad:    dX = dY . W[k]^T)
 * flip_taps != 0 writes tap k to slot K-1-k (dgrad of a same-grid stride-1 odd-K plan reuses the
 * forward map with flipped taps). */
FVC_API int fvc_pack_weights(const void *weights, const int64_t strides[5], int32_t dtype_in, int32_t cout, int32_t cin,
                     int32_t k0, int32_t k1, int32_t k2, int32_t layout, int32_t flip_taps, int32_t dtype_out, void *out,
                     fvc_stream_t stream);

/* -------- execution (replaces gs_conv / gs_conv_transpose / gs_conv_backward /
 *          gs_conv_transpose_backward / pred_gather_igemm_conv, Bindings.cpp:585-674 ->
 *          GatherScatterDefault.cu:673-924, PredGatherIGemm.cu:1121-1172) ----------------------- */
/* Output-stationary sparse convolution: y[o,:] = sum_k x[nbr[k*pitch+o],:] . w[k]   (w: [K][Cin][Cout]).
 * One entry point serves forward (either direction) and dgrad (x = grad_output, nbr = the reversed
 * map, w packed with layout 1): the maps are already oriented (GatherScatterDefault.cu:734-740).
 * x, y, w share `dtype`; accumulation is fp32 (fp64 for FVC_F64).  y is fully overwritten (rows with
 * no neighbour become 0; GatherScatterDefault.cu:696).  bias (may be NULL): [Cout] in `dtype`, added in
 * the epilogue (fvdb/nn/modules.py:370-371).  path: 0 = automatic (tensor-core kernel when the dtype is
 * f16/bf16/f32 and the channel counts allow it, else the CUDA-core kernel), 1 = force CUDA-core path,
 * 2 = force tensor-core path (FVC_ERR_UNSUPPORTED if not admissible).  fp32 runs on the tensor pipe as a
 * three-way bf16 split (six exact bf16 products per fp32 product, fp32 accumulate; relative error ~1e-6,
 * inside the reference's 1e-5 bar); the CUDA-core path is plain fp32 FMA.  tile_mask: fvc_kmap_tile_mask of
 * `nbr` (may be NULL: no unit skipping).  scratch: 256-byte aligned, fvc_conv_scratch_bytes() bytes (weight
 * image + for fp32 the split copy of the n_in feature rows); 0 bytes when only the CUDA-core path applies. */
FVC_API size_t fvc_conv_scratch_bytes(int64_t n_in, int64_t n_out, int32_t cin, int32_t cout, int64_t kernel_volume, int32_t dtype);
FVC_API int fvc_conv_forward(const void *x, const void *w_packed, const void *bias, void *y, const int32_t *nbr, int64_t pitch,
                     const uint64_t *tile_mask, int64_t n_in, int64_t n_out, int32_t cin, int32_t cout, int64_t kernel_volume, int32_t dtype,
                     int32_t path, void *scratch, size_t scratch_bytes, fvc_stream_t stream);
/* ---- prepared weights + fused block epilogue (bias + BatchNorm-apply + ReLU + residual in
 *      the GEMM epilogue, fvdb/nn/modules.py:484-521, fvdb/nn/simple_unet.py:233-243) ---------------------------------
 * fvc_conv_prepare_weights writes, in ONE launch from the public [Cout,Cin,k0,k1,k2] tensor (any strides), the operand the
 * executor chosen by (dtype, channels, path) consumes: the pre-swizzled shared-memory image of the tcgen05 kernels (fp32:
 * its three-way bf16 split) or the packed [K][Cin][Cout] array of the CUDA-core kernels.  transpose != 0 prepares W[k]^T
 * (dgrad); flip_taps as in fvc_pack_weights.  The blob can be cached for as long as the weights do not change.
 * fvc_conv_weights_bytes takes the EXECUTOR's channel counts (dgrad: cin = public Cout, cout = public Cin). */
FVC_API size_t fvc_conv_weights_bytes(int32_t cin, int32_t cout, int64_t kernel_volume, int32_t dtype, int32_t path);
FVC_API int fvc_conv_prepare_weights(const void *weights, const int64_t strides[5], int32_t dtype_in, int32_t cout, int32_t cin, int32_t k0,
                                     int32_t k1, int32_t k2, int32_t transpose, int32_t flip_taps, int32_t dtype, int32_t path, void *prepared,
                                     size_t prepared_bytes, fvc_stream_t stream);
/* stored[o,c] = act2(act1((acc[o,c] + bias[c]) * scale[c] + shift[c]) + residual[o,c]); every member may be NULL / 0.
 * relu bit 0: act1 = ReLU (the block's activation, before a skip connection joins); bit 1: act2 = ReLU (after it:
 * fvdb/nn/simple_unet.py:187-188).  bias / residual in `dtype`, scale / shift fp32 (an eval-mode BatchNorm folds into them).  stats: fp32
 * [fvc_conv_stats_blocks()][2][Cout], per block of rows_per_block consecutive output rows the column sums and sums of squares
 * of the STORED values (what a BatchNorm statistics pass over y would read) -- written, not accumulated; deterministic.
 * scale / shift / residual / relu / stats need the tensor-core path (FVC_ERR_UNSUPPORTED otherwise). */
typedef struct FvcConvEpilogue {
    const void *bias;
    const float *scale;
    const float *shift;
    const void *residual;
    int32_t relu;
    float *stats;
} FvcConvEpilogue;
FVC_API int64_t fvc_conv_stats_blocks(int64_t n_out, int32_t cin, int32_t cout, int64_t kernel_volume, int32_t dtype, int32_t path,
                                      int32_t *rows_per_block);
/* fvc_conv_forward over prepared weights.  x_is_split != 0 (fp32 only): x holds the bf16 split rows fvc_split_rows wrote
 * ([n_in][3][Cin] bf16) -- a layer splits its input once and reuses the rows for forward and wgrad.  scratch: only for fp32
 * with x_is_split == 0 (fvc_conv_scratch_bytes is always enough). */
FVC_API int fvc_conv_forward_ex(const void *x, int32_t x_is_split, const void *w_prepared, const FvcConvEpilogue *epilogue, void *y,
                                const int32_t *nbr, int64_t pitch, const uint64_t *tile_mask, int64_t n_in, int64_t n_out, int32_t cin,
                                int32_t cout, int64_t kernel_volume, int32_t dtype, int32_t path, void *scratch, size_t scratch_bytes,
                                fvc_stream_t stream);
/* fp32 rows [n][channels] -> bf16 split rows [n][3][channels] (x = s0 + s1 + s2 exactly to 24 bits); channels % 8 == 0 */
FVC_API int fvc_split_rows(const float *x, int64_t n, int32_t channels, void *split_rows, fvc_stream_t stream);

/* Weight gradient: grad_w[co][ci][k0][k1][k2] = sum over pairs p of tap k of x[gather[p]][ci] * dy[scatter[p]][co]
 * (GatherScatterDefault.cu:806-813), written contiguous in the public layout in `dtype`;
 * fixed-order (deterministic) fp32/fp64 reduction.  offsets_host / offsets_dev: the same int64 [K^3+1]
 * CSR offsets on the host and on the device.  nbr/pitch: the output-stationary dense map of the same
 * rulebook (used by the tensor-core path; may be NULL, then the CSR path runs).  When the tensor-core path runs (nbr given,
 * dtype / channels admitted, path != 1) the CSR arguments (gather, scatter, offsets_host, offsets_dev) may all be NULL:
 * the kernel reads the dense map only, so a plan never has to materialise the CSR view for training. */
/* Scratch for the kernel family fvc_conv_wgrad will take: max_pairs_per_tap = the largest CSR tap segment (sizes the
 * CUDA-core partials), path as in the call, has_dense_map = whether nbr will be passed. */
FVC_API size_t fvc_conv_wgrad_scratch_bytes(int64_t n_in, int64_t n_out, int64_t max_pairs_per_tap, int32_t cin, int32_t cout,
                                    int64_t kernel_volume, int32_t dtype, int32_t path, int32_t has_dense_map);
FVC_API int fvc_conv_wgrad(const void *x, const void *dy, const int32_t *gather, const int32_t *scatter,
                   const int64_t *offsets_host, const int64_t *offsets_dev, const int32_t *nbr, int64_t pitch,
                   const uint64_t *tile_mask, int64_t n_in, int64_t n_out,
                   int32_t cin, int32_t cout, int64_t kernel_volume, int32_t dtype, int32_t path, void *grad_w,
                   void *scratch, size_t scratch_bytes, fvc_stream_t stream);
/* Same, with fp32 operands optionally passed as the bf16 split rows of fvc_split_rows (x: [n_in][3][Cin], dy: [n_out][3][Cout]). */
FVC_API int fvc_conv_wgrad_ex(const void *x, int32_t x_is_split, const void *dy, int32_t dy_is_split, const int32_t *gather,
                      const int32_t *scatter, const int64_t *offsets_host, const int64_t *offsets_dev, const int32_t *nbr, int64_t pitch,
                      const uint64_t *tile_mask, int64_t n_in, int64_t n_out, int32_t cin, int32_t cout, int64_t kernel_volume,
                      int32_t dtype, int32_t path, void *grad_w, void *scratch, size_t scratch_bytes, fvc_stream_t stream);

/* Fused backward for narrow layers (cin, cout in {16, 32}, f16 / bf16, K^3 <= 128; fvc_conv_kernel_family(..., pass = 2) == 2):
 * grad_features AND grad_w from ONE gather of grad_output (both sums of GatherScatterDefault.cu:803-807 consume the same
 * gathered rows).  in_map / in_tile_mask: the INPUT-stationary dense map (fvc_kmap_reverse_from_dense, or the forward map of a
 * same-grid stride-1 odd-kernel plan with flip_taps = 1) and its tile mask; w_prepared_transposed: fvc_conv_prepare_weights
 * (transpose = 1, the same flip_taps).  cin / cout are the public weight dimensions.  Deterministic (no atomics). */
FVC_API size_t fvc_conv_backward_fused_scratch_bytes(int64_t n_in, int32_t cin, int32_t cout, int64_t kernel_volume);
FVC_API int fvc_conv_backward_fused(const void *grad_output, const void *features, const void *w_prepared_transposed, const int32_t *in_map,
                            int64_t pitch, const uint64_t *in_tile_mask, int64_t n_in, int64_t n_out, int32_t cin, int32_t cout,
                            int64_t kernel_volume, int32_t dtype, int32_t flip_taps, void *grad_features, void *grad_w, void *scratch,
                            size_t scratch_bytes, fvc_stream_t stream);

/* -------- stride-1 generated topologies by leaf-mask morphology (fast path of conv_grid / conv_transpose_grid;
 *          replaces the NanoVDB DilateGrid route of ops/BuildGridForConv.cu:392-463) --------------------------------
 * dst_leaves: the leaves of a grid whose tree already holds every leaf the result can touch (built by the ordinary
 * grid builder from the neighbour-leaf origins of `src`); their masks / prefixes / counts are overwritten with
 * out(c) = OR over o in [lo, hi]^3 of src(c - o), |o| < 8; leaf_counts[n] receives the voxel count of every leaf. */
FVC_API int fvc_grid_dilate_leaves(const FvcGridBatch *src, FvcLeaf *dst_leaves, int32_t n_dst_leaves, const int32_t lo[3],
                           const int32_t hi[3], int32_t *leaf_counts, fvc_stream_t stream);
/* leaf_base[n] (exclusive scan of the counts, batch-cumulative) -> leaf records, voxel list out_ijk [rows][3] and grid
 * index out_bidx [rows] in row order */
FVC_API int fvc_grid_expand_leaves(FvcLeaf *leaves, int32_t n_leaves, const int32_t *leaf_base, int32_t *out_ijk, int32_t *out_bidx,
                           fvc_stream_t stream);

/* -------- normalisation around the convolution (replaces torch.nn.BatchNorm1d over jdata
 *          + the separate ReLU pass of fvdb/nn/modules.py:484-521,91-110 and the bias-gradient reduction) -------------
 * Rows are [n][channels] row-major in `dtype` (f16 / bf16 / f32), channels a multiple of 16 bytes' worth of elements.
 * Statistics, affine parameters and sums are fp32 [channels]; gamma / beta may be NULL (1 / 0).  `scratch` holds the
 * per-CTA partials (fvc_bn_scratch_bytes).  All entry points are asynchronous on `stream` and deterministic. */
FVC_API size_t fvc_bn_scratch_bytes(int32_t channels);
/* mean / biased variance over the n rows; running_mean / running_var (may be NULL) are updated in place with `momentum`
 * (running_var from the unbiased variance), as torch.nn.BatchNorm1d does in training mode. */
FVC_API int fvc_bn_stats(const void *x, int64_t n, int32_t channels, int32_t dtype, float *mean, float *var, float *running_mean,
                 float *running_var, float momentum, void *scratch, size_t scratch_bytes, fvc_stream_t stream);
/* y = act((x - mean) / sqrt(var + eps) * gamma + beta), act = ReLU when relu != 0 */
/* Same statistics from the per-block column sums the convolution epilogue wrote (FvcConvEpilogue.stats): partial
 * [blocks][2][channels], block b covering rows [b * rows_per_block, min(n, (b + 1) * rows_per_block)). */
FVC_API int fvc_bn_stats_from_partials(const float *partial, int64_t blocks, int32_t rows_per_block, int64_t n, int32_t channels, float *mean,
                               float *var, float *running_mean, float *running_var, float momentum, fvc_stream_t stream);
FVC_API int fvc_bn_apply(const void *x, int64_t n, int32_t channels, int32_t dtype, const float *mean, const float *var, const float *gamma,
                 const float *beta, float eps, int32_t relu, void *y, fvc_stream_t stream);
/* sums[0][c] = sum dz (= grad beta), sums[1][c] = sum dz * xhat (= grad gamma); dz = dy masked by the fused ReLU */
FVC_API int fvc_bn_backward_reduce(const void *dy, const void *x, int64_t n, int32_t channels, int32_t dtype, const float *mean, const float *var,
                           const float *gamma, const float *beta, float eps, int32_t relu, float *sums, void *scratch, size_t scratch_bytes,
                           fvc_stream_t stream);
/* dx from dy, x and the (possibly all-reduced) sums over `count` rows (count_dev, if not NULL, is a device float holding
 * the row count of a distributed batch and overrides `count`); training == 0: statistics are constants */
FVC_API int fvc_bn_backward_apply(const void *dy, const void *x, int64_t n, int32_t channels, int32_t dtype, const float *mean, const float *var,
                          const float *gamma, const float *beta, float eps, int32_t relu, int32_t training, const float *sums, int64_t count,
                          const float *count_dev, void *dx, fvc_stream_t stream);
/* sums[c] = sum over rows of x[:, c] (bias gradient of SparseConv3d, fvdb/nn/modules.py:370-371) */
FVC_API int fvc_column_sums(const void *x, int64_t n, int32_t channels, int32_t dtype, float *sums, void *scratch, size_t scratch_bytes,
                    fvc_stream_t stream);

/* -------- pooling / refinement between a fine and a coarse grid (replaces
 *          ops/MaxPool.cu:16-122, ops/AvgPool.cu:17-110, ops/Refine.cu:17-110 behind GridBatch.max_pool / avg_pool / refine)
 * idx[n_out][taps] (int32, row-major): for every output row the rows of its window children in x, -1 = inactive; the
 * host builds it with fvc_ijk_to_index.  Rows are [n][channels] in `dtype` (f16 / bf16 / f32), channels a multiple of a
 * 16-byte vector.  mode 0: max over the children (0 when there is none: the documented contract, fvdb/nn/modules.py:125-128;
 * the reference kernel leaves -inf there, MaxPool.cu:46); mode 1: sum * scale. */
FVC_API int fvc_pool_rows(const void *x, const int32_t *idx, int64_t n_out, int32_t taps, int32_t channels, int32_t dtype, int32_t mode,
                  float scale, void *y, fvc_stream_t stream);
/* dx (zero-initialised by the caller): mode 0 routes dy[o] to the first maximal child per channel (MaxPool.cu:100-118),
 * mode 1 writes dy[o] * scale to every child (AvgPool.cu:98-108); windows must not overlap (one writer per element). */
FVC_API int fvc_pool_rows_backward(const void *dy, const void *x, const int32_t *idx, int64_t n_out, int32_t taps, int32_t channels,
                           int32_t dtype, int32_t mode, float scale, void *dx, fvc_stream_t stream);
/* y[r] = idx[r] >= 0 ? x[idx[r]] : 0  (nearest-neighbour refinement, Refine.cu:43-56) */
FVC_API int fvc_gather_rows(const void *x, const int32_t *idx, int64_t n_out, int32_t channels, int32_t dtype, void *y, fvc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FVDBCONV_H */
