/* libaocb200 -- C ABI of the B200-native (sm_100a) AOC-Net per-frame inference kernels.
 *
 * The reference (JerryX1110/Robust-Video-Object-Segmentation) is pure Python/PyTorch and has no FFI of its own:
 * its boundary for this path is the Python duck type `AOCNet.forward_for_eval` (networks/aoc/aocnet.py:84) and the
 * op callables imported at networks/aoc/aocnet.py:6-8.  Each entry point below replaces one of those torch call
 * sites (cited per function; paths relative to AOC-Net/complete_project/AOCNet/).  The Python host
 * (aocb200/model.py) binds them with ctypes; INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (except where "host" is stated); no hidden allocation;
 *   - activations are fp32 NHWC ("pixel-major": [N][H*W][C]) with an explicit row stride `ld*` in floats, so a
 *     kernel can read/write a channel slice of a wider concat buffer; label maps are uint8 object ids
 *     (0 = background slot, 1..K objects, anything >= O, e.g. 125 = "uncertain", belongs to no object);
 *   - O = K + 1 object slots, O <= AOC_MAX_OBJECTS; the embedding width is fixed at 100
 *     (cfg.MODEL_SEMANTIC_EMBEDDING_DIM);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); calls are asynchronous and re-entrant per
 *     stream; workspaces are sized by the matching *_workspace_bytes();
 *   - return value: AOC_OK or a negative AOC_E* code; aoc_last_error_string() describes the last failure of the
 *     calling thread.  Nothing throws or aborts.  There is no CPU fallback: without a CUDA device every launch
 *     returns AOC_ELAUNCH.
 */
#ifndef AOCB200_H_
#define AOCB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define AOC_OK 0
#define AOC_EINVAL (-1)   /* bad argument (shape, alignment, null pointer, workspace too small) */
#define AOC_ELAUNCH (-2)  /* CUDA launch/runtime failure (cudaGetLastError) */
#define AOC_EARCH (-3)    /* device is not sm_100 */

#define AOC_MAX_OBJECTS 16   /* O = K+1 slots */
#define AOC_KMEANS_MAX_K 64  /* largest cluster_num (matching.py:507; the reference wires 16) */
#define AOC_PROXY_SLOTS 36   /* per object at kmax = 16: 16 centroids, 16 centroid_avg, 1 mean proxy, 3 pad (2*kmax + 4) */
#define AOC_META_INTS (2 * AOC_MAX_OBJECTS + 3)

int aoc_version(void);
const char* aoc_last_error_string(void);
/* 0 if device `dev` can run this library (compute capability 10.x), else AOC_EARCH / AOC_ELAUNCH. */
int aoc_check_device(int dev);
/* Tuning / diagnostic switches for tests and tools -- the ONE piece of process-global mutable state in the library
 * (everything that selects arithmetic on the product path is a per-call argument).  "conv_chunk" (default 8): default
 * length, in 16-channel stages, of the TMEM accumulation chains of the tensor-core convolution (see
 * aoc_conv2d_nhwc_tc).  "conv_splitk" (default 1): allow the split-K schedule for layers with few output tiles.
 * "conv_pdl" (default 1): programmatic dependent launch.  "match_f16" (default 1): split-fp16 operands in the matching
 * contraction (0 = 3xTF32).  "conv_halo" (default 1): halo variant of the convolution for 3x3 / stride-1 / pad = dilation <= 2
 * layers with at least one 16 x 8 tile per SM.  "conv_tail" (default 1) / "conv_tail_min_stages" (default 192): K slices
 * for the tiles of a partial last wave of long K loops.  "match_fast" (default 0): FAST precision mode of the global matching
 * -- one fp16 MMA per product instead of three; the only switch that changes results beyond fp32 rounding (schedules differ in
 * summation order only), off unless a caller asks for it.  "match_collector" (default 0): tcgen05 collector hints on the
 * query operand of the matching contraction (identical results; measured slower, kept as an experiment switch).
 * "glue_pdl" (default 0): programmatic dependent launch of the GroupNorm coefficient kernel itself (identical results;
 * measured slower). */
int aoc_set_option(const char* key, int value);

/* ---------------------------------------------------------------- convolutions (conv_simt.cu, umma_conv2.cu) */
/* nn.Conv2d (+ folded FrozenBatchNorm2d bias, + residual, + ReLU): resnet.py:23-42,108-123; deeplab/aspp.py:62-74;
 * deeplab/decoder.py:32-41; layers/gct.py:68-91; layers/aspp.py:57-70; decoding_module.py:162-190,228-240.
 * x [N][H][W][ldx>=Cin], w [Cout][kh][kw][Cin], y [N][Ho][Wo][ldy>=Cout]; in_scale (optional) [N][Cin] multiplies the
 * input per (sample, channel) -- a fused IA/GCT gate.  fp32 FMA, exact-path. */
int aoc_conv2d_nhwc_f32(const float* x, const float* w, const float* bias, const float* residual,
                        const float* in_scale, float* y, int N, int H, int W, int Cin, int ldx, int Cout, int ldy,
                        int ldres, int kh, int kw, int stride, int pad, int dil, int relu, cudaStream_t stream);
/* Same contract on the tcgen05 tensor cores (csrc/umma_conv2.cu).  fp32 products are rebuilt from two-term operand
 * splits (22 mantissa bits, the lo*lo term dropped), selected per call by `operand_mode`:
 *     AOC_CONV_SPLIT_F16  hi = fp16(x), lo = fp16((x - hi) * 2^11): three K = 16 kind::f16 MMAs per 16 channels.
 *                         Range: |x|, |w| saturate at 65504; `overflow_flag` (optional device word) is set to 1 when an
 *                         activation operand reaches 6e4, so the caller can repeat the frame in the other mode.
 *     AOC_CONV_TF32X3     hi = rna_tf32(x), lo = rna_tf32(x - hi): six K = 8 kind::tf32 MMAs; fp32 exponent range.
 * fp32 accumulation in TMEM in chains of `chunk_stages` x 16 input channels that are summed in fp32 registers (round to
 * nearest), so the result has fp32-FMA quality for any K.  The activation patch is fetched by TMA (zero fill =
 * padding) and the per-(sample, channel) affine that precedes the convolution in the reference is fused into the
 * operand path:
 *     x_eff = relu?(x * in_a[n,c] + in_b[n,c])      (GroupNorm apply + ReLU, GCT gate, IA gate; each optional)
 * w_packed comes from aoc_conv_pack_weights with the SAME operand_mode (size aoc_conv_packed_weight_bytes) and must be
 * COMPLETE when the convolution is launched (synchronise the stream once after packing): under programmatic dependent
 * launch (option conv_pdl, default on) the kernel's weight-copy warp starts before the preceding kernel of the stream
 * has finished -- weights are constants of the model.  Requires
 * ldx % 4 == 0, stride in {1, 2}; Cin <= 1024 when an input affine is given.  chunk_stages <= 0 selects the default (8). */
#define AOC_CONV_TF32X3 0
#define AOC_CONV_SPLIT_F16 1
size_t aoc_conv_packed_weight_bytes(int Cout, int Cin, int kh, int kw, int operand_mode);
int aoc_conv_pack_weights(const float* w, int Cout, int Cin, int kh, int kw, int operand_mode, void* w_packed,
                          cudaStream_t stream);
int aoc_conv2d_nhwc_tc(const float* x, const void* w_packed, const float* bias, const float* residual,
                       const float* in_a, const float* in_b, int in_relu, float* y, float* tile_stats, int N, int H,
                       int W, int Cin, int ldx, int Cout, int ldy, int ldres, int kh, int kw, int stride, int pad,
                       int dil, int relu, int chunk_stages, int operand_mode, int* overflow_flag, void* workspace,
                       size_t ws_bytes, cudaStream_t stream);
/* workspace (optional, aoc_conv_workspace_bytes): partial sums for split-K, used when the layer has too few output
 * tiles to fill the chip (the 31x54 maps of the backbone); without it such layers run unsplit.  Its first 4096 bytes
 * are arrival counters that must be ZERO when the caller first hands the buffer over; every launch leaves them zero
 * again (the K slice that arrives last at a tile adds the slices in index order, finishes the tile and resets it). */
size_t aoc_conv_workspace_bytes(int N, int H, int W, int Cout, int kh, int kw, int stride, int pad, int dil);
/* tile_stats (optional): [N * aoc_conv_tiles_per_image(...)][2][Cout] floats receiving, per statistics row (= a 32-pixel
 * quadrant of a 128-pixel output tile: what one epilogue warp stores; aoc_conv_tiles_per_image counts these rows), the
 * per-channel sum and sum of squares of the stored output -- the GroupNorm / GCT statistics of the next layer come
 * out of the convolution epilogue instead of a second pass over the tensor (aoc_tile_stats_reduce_f32 folds them into
 * the [N][2][C] double layout of aoc_channel_stats_f32). */
int aoc_conv_tiles_per_image(int N, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int dil,
                             int operand_mode);
/* tooling: when non-null, CTA 0 of every later aoc_conv2d_nhwc_tc launch records clock64() of its pipeline events
 * (8 events x the first 256 stages, uint64) into this device buffer; see tools/conv_trace.py.  NULL switches it off. */
int aoc_conv_trace(void* device_buffer_16x256_u64);
int aoc_tile_stats_reduce_f32(const float* tile_stats, int N, int tiles_per_image, int C, double* stats,
                              cudaStream_t stream);
/* depthwise 3x3 pad 1 + bias (aocnet.py:19 seperate_conv); w [C][3][3] */
int aoc_dwconv3x3_nhwc_f32(const float* x, const float* w, const float* bias, float* y, int N, int H, int W, int C,
                           cudaStream_t stream);
/* F.max_pool2d(x, 3, 2, 1) (resnet.py:113) */
int aoc_maxpool3x3s2_nhwc_f32(const float* x, float* y, int N, int H, int W, int C, cudaStream_t stream);

/* ---------------------------------------------------------------- statistics / affine (norm.cu) */
size_t aoc_channel_stats_workspace_bytes(int N, int HW, int C);
/* stats [N][2][C] doubles = per-(sample, channel) sum and sum of squares over HW; with (phi, thr) only pixels with
 * phi[n,p] > thr[n] are summed (masked GAP of conditioning_layer.py:38-43). */
int aoc_channel_stats_f32(const float* x, int N, int HW, int C, int ldx, const float* phi, const float* thr,
                          double* stats, void* workspace, size_t ws_bytes, cudaStream_t stream);
/* nn.GroupNorm as y = x*a[n,c] + b[n,c] */
/* the same coefficients straight from a convolution's tile_stats (aoc_conv2d_nhwc_tc), C / groups <= 32 */
int aoc_gn_coeffs_tiles_f32(const float* tile_stats, int tiles_per_image, const float* gamma, const float* beta, int N,
                            int C, int groups, int HW, float eps, float* a, float* b, cudaStream_t stream);
int aoc_gn_coeffs_f32(const double* stats, const float* gamma, const float* beta, int N, int C, int groups, int HW,
                      float eps, float* a, float* b, cudaStream_t stream);
/* GCT gate (layers/gct.py:17-36): a[n,c] = pre*(1 + tanh(emb*norm + beta)) */
int aoc_gct_coeffs_f32(const double* stats, const float* alpha, const float* gamma, const float* beta,
                       const float* pre_scale, int N, int C, float eps, float* a, cudaStream_t stream);
/* adaptive_avg_pool2d(x, 1) from the statistics */
int aoc_gap_from_stats_f32(const double* stats, int N, int C, int HW, float* out, cudaStream_t stream);
/* y = x*a[n,c] (+ b[n,c]) (+ residual*res_scale[n,c]) (ReLU) */
int aoc_affine_nc_f32(const float* x, const float* a, const float* b, const float* residual, const float* res_scale,
                      float* y, int N, int HW, int C, int ldx, int ldy, int ldres, int relu, cudaStream_t stream);

/* same, and the [N][2][C] statistics of y as a by-product of the pass (workspace as for aoc_channel_stats_f32) */
int aoc_affine_stats_nc_f32(const float* x, const float* a, const float* b, const float* residual,
                            const float* res_scale, float* y, int N, int HW, int C, int ldx, int ldy, int ldres, int relu,
                            double* stats, void* workspace, size_t ws_bytes, cudaStream_t stream);

/* ---------------------------------------------------------------- FiLM conditioning (film.cu) */
/* phi_layer: 1x1 conv C->1 (conditioning_layer.py:27) */
int aoc_cond_phi_f32(const float* x, const float* w, const float* b, float* phi, int N, int HW, int C, int ldx,
                     cudaStream_t stream);
/* torch.topk(vals, k)[..., -1] per row (conditioning_layer.py:32-36); k is 1-based */
int aoc_kth_largest_f32(const float* vals, int N, int L, int k, float* out, cudaStream_t stream);
/* nn.Linear; act 0 = identity, 1 = 1 + tanh (IA_gate attention.py:12-17, conditioning_block :82-85) */
int aoc_linear_f32(const float* x, const float* W, const float* b, float* y, int N, int M, int K, int ldx, int ldy,
                   int act, cudaStream_t stream);
/* out[n] = sum_n' v[n'] - v[n] (conditioning_layer.py:69, decoding_module.py IA9/IA10/IA11 heads) */
int aoc_delta_sum_f32(const float* v, float* out, int N, int C, int ldo, cudaStream_t stream);

/* ---------------------------------------------------------------- layout / resampling (resize.cu) */
int aoc_image_to_nhwc4_f32(const float* x_chw3, float* y_hwc4, int H, int W, cudaStream_t stream);
/* Space-to-depth form of the frame for the ResNet stem (resnet.py:108-110: 7x7 / stride 2 / pad 3): y is
 * [ceil(H/2) + 1][ceil(W/2) + 1][16], pixel (i, j) = the 2x2 source block at (2(i-1), 2(j-1)) as channels (py*2+px)*4 + c,
 * zero outside the image, in row 0 / column 0 and in every c = 3.  The stem is then a 4x4 / stride-1 / pad-1 convolution
 * over y (weights rearranged by the caller: source tap r = 2a' + py - 1), 16 K stages instead of 49. */
int aoc_image_to_s2d16_f32(const float* x_chw3, float* y_s2d, int H, int W, cudaStream_t stream);
int aoc_nchw_to_nhwc_f32(const float* x, float* y, int N, int C, int HW, int ldy, cudaStream_t stream);
int aoc_nhwc_to_nchw_f32(const float* x, float* y, int N, int C, int HW, int ldx, cudaStream_t stream);
/* F.interpolate(mode='bilinear', align_corners=True); with (ids, table) the source is table[ids[pixel]] */
int aoc_resize_bilinear_nhwc_f32(const float* x, const uint8_t* ids, const float* table, int n_table, float* y, int N,
                                 int Hi, int Wi, int Ho, int Wo, int C, int ldx, int ldy, cudaStream_t stream);
/* F.interpolate(mode='bicubic', align_corners=True) (decoding_module.py:163) */
int aoc_resize_bicubic_nhwc_f32(const float* x, float* y, int N, int Hi, int Wi, int Ho, int Wo, int C, int ldx,
                                int ldy, cudaStream_t stream);
/* torch.cat along channels: copy a channel slice between buffers with different row strides */
int aoc_copy_channels_f32(const float* x, float* y, long long rows, int C, int ldx, int ldy, cudaStream_t stream);
/* input edge of the eval loop: MultiRestrictSize's cv2.resize(INTER_CUBIC) + optional mirror + MultiToTensor's /255, -mean,
 * /std, HWC -> CHW (dataloaders/custom_transforms.py:387-463, :465-487) in one kernel.  img_hwc: uint8 [H][W][3] on the
 * device; mean3 / std3: HOST arrays of 3 floats; out_chw: float [3][Ho][Wo].  Ho == H && Wo == W: no resize (as the transform). */
int aoc_prepare_frame_u8(const uint8_t* img_hwc, int H, int W, int Ho, int Wo, int flip, const float* mean3,
                         const float* std3, float* out_chw, cudaStream_t stream);
/* F.interpolate(mode='nearest') on label maps (aocnet.py:128-135) */
int aoc_resize_nearest_u8(const uint8_t* x, uint8_t* y, int Hi, int Wi, int Ho, int Wo, cudaStream_t stream);
/* the same for an int64 label map (torch.argmax output, eval_manager_mm.py:318-320), values clamped to 0..255 */
int aoc_resize_nearest_i64(const long long* x, uint8_t* y, int Hi, int Wi, int Ho, int Wo, cudaStream_t stream);
/* plumbing that would otherwise be eager torch kernels inside the captured frame: fill n 32-bit words; out = a + b
 * (op 0) or a * b (op 1) over n floats (sums / products of per-(sample, channel) coefficient vectors) */
int aoc_fill_u32(void* p, unsigned int value, long long n_words, cudaStream_t stream);
int aoc_vec_op_f32(const float* a, const float* b, float* out, int n, int op, cudaStream_t stream);

/* ---------------------------------------------------------------- reference bank (matching.cu) */
size_t aoc_bank_workspace_bytes(int total_pixels, int O);
/* Object-sorted index of all bank pixels (matching.py:2486-2495, :533-545).  ids: uint8 [total_pixels] (frames
 * concatenated).  meta_out (device int32[AOC_META_INTS]): [o] = rows of object o, [AOC_MAX_OBJECTS+o] = first sorted
 * row of object o (segments padded to `align` rows), [2*MAX+1] = padded row count, [2*MAX+2] = valid pixels. */
int aoc_bank_index_build(const uint8_t* ids, int total_pixels, int O, int align, int* meta_out, int* row_src,
                         int cap_rows, int* nat2sorted, void* workspace, size_t ws_bytes, cudaStream_t stream);
int aoc_bank_gather_f32(const float* emb_all, const int* row_src, int rows, float* S, float* r2, cudaStream_t stream);
/* tcgen05 operand image ("tc image", see csrc/umma.cuh): 3xTF32 hi/lo split of x[rows][K] in the K-major core-matrix
 * layout, row blocks of RB rows, K zero-padded to K_img (multiple of 8).  Used for the sorted bank (RB = 256,
 * K_img = 104) and consumed by TMA bulk copies without further transformation. */
size_t aoc_tc_image_bytes(long long rows, int K_img, int RB);
int aoc_pack_tc_image_f32(const float* x, long long rows, int K, int ld, int RB, int K_img, void* out,
                          cudaStream_t stream);

/* ---------------------------------------------------------------- matching (matching.cu, umma_match.cu) */
/* global_matching_for_eval (matching.py:2384-2510): out [HW][O] */
int aoc_global_match_simt_f32(const float* q, int HW, const float* S, const float* r2, const int* meta,
                              const float* bias, int O, float* mins_ws, float* out, cudaStream_t stream);
/* Same result on the tcgen05 tensor cores (3xTF32, TMEM accumulators, TMA-fed).  S / r2 = sorted bank rows and their
 * squared norms (+inf on padding rows) from aoc_bank_gather_f32 built with align = 256; rows_padded =
 * meta[2*AOC_MAX_OBJECTS+1].  Both operands are translated by the column mean of q and packed into tc images inside
 * the workspace (|q-r|^2 is translation invariant; centred operands keep the TMEM accumulation unbiased). */
size_t aoc_global_match_tc_workspace_bytes(int HW, int rows_padded);
int aoc_global_match_tc(const float* q, int HW, const float* S, const float* r2, const int* meta_dev, int rows_padded,
                        const float* bias, int O, void* workspace, size_t ws_bytes, float* out, cudaStream_t stream);
/* ---- bank-sharded matching of ONE sequence over several GPUs (the op is matching.py:2384-2516, the bank
 * growth eval_manager_mm.py:309-312).  One process per GPU, all ranks run the same frames; rank r contracts the queries
 * with row blocks [nrb*r/world, nrb*(r+1)/world) of the object-sorted bank only, and the min over ranks -- associative,
 * so `out` is bit-identical to aoc_global_match_tc's -- is exchanged from INSIDE the matching kernel: each CTA stores
 * the minima of its 128 query rows into its slot of every peer's exchange area (peer-mapped memory: NVLink writes),
 * fences at system scope and bumps the peer's arrival counter; a finalize kernel waits for the counters and reduces.
 * No collective call and no host synchronisation per frame.
 * areas[g]: rank g's exchange area (aoc_match_shard_area_bytes(world, cap_hw) bytes from aoc_peer_alloc, exported with
 * aoc_peer_export, mapped with aoc_peer_open; areas[rank] = the local allocation) -- a HOST array of `world` pointers;
 * state: 64 zeroed device bytes owned by this rank; HW <= cap_hw. */
#define AOC_PEER_HANDLE_BYTES 64
int aoc_peer_alloc(size_t bytes, void** ptr_out);
int aoc_peer_free(void* ptr);
int aoc_peer_export(void* ptr, void* handle_out);
int aoc_peer_open(const void* handle, void** ptr_out);
int aoc_peer_close(void* ptr);
size_t aoc_match_shard_area_bytes(int world, int cap_hw);
int aoc_match_shard_range(int rows_padded, int rank, int world, int* rb_begin, int* rb_end);
int aoc_global_match_tc_sharded(const float* q, int HW, const float* S, const float* r2, const int* meta_dev,
                                int rows_padded, const float* bias, int O, int rank, int world, void* const* areas,
                                int cap_hw, void* state, void* workspace, size_t ws_bytes, float* out,
                                cudaStream_t stream);
int aoc_global_match_finalize_f32(const float* mins, const int* meta, const float* bias, int HW, int O, float* out,
                                  cudaStream_t stream);
/* cluster level (matching.py:602-637) and k=1 proxy level (matching.py:149-197): out_cluster [HW][O][2], out_proxy [HW][O];
 * P / pvalid / kmax as produced by aoc_kmeans_proxies_f32 */
int aoc_proxy_match_f32(const float* q, int HW, const float* P, const int* pvalid, const float* bias, int O, int kmax,
                        float* out_cluster, float* out_proxy, cudaStream_t stream);
size_t aoc_head_pool_workspace_bytes(int total_pixels);
/* calculate_attention_head_for_eval_p_m (attention.py:155-189): masked means written into head rows */
int aoc_head_pool_f32(const float* emb, const uint8_t* ids, int total_pixels, int O, float eps, float* head,
                      int ld_head, int off_pos, int off_neg, float* pos_out, int ld_pos, void* workspace,
                      size_t ws_bytes, cudaStream_t stream);
int aoc_row_sqnorm_f32(const float* x, int rows, float* out, cudaStream_t stream);
/* local_matching / local_matching_proxy on the half-resolution grid (matching.py:2710-2851): out [hh*ww][ld_out], o*6+ch */
int aoc_local_match_f32(const float* xq, const float* yp, const float* x2, const float* y2, const uint8_t* ids, int hh,
                        int ww, int O, const float* bias, float* out, int ld_out, cudaStream_t stream);
/* foreground2background + concat (matching.py:9-23, aocnet.py:349-358): out [O][HW][24] */
int aoc_prehead_assemble_f32(const float* g, const float* gc, const float* gp, const float* loc, const float* locp,
                             int ld_loc, const uint8_t* prev_ids, int HW, int O, float* out, cudaStream_t stream);
int aoc_broadcast_rows_f32(const float* x, float* y, int N, int HW, int C, int ldx, int ldy, cudaStream_t stream);

/* ---------------------------------------------------------------- adaptive object proxies (kmeans.cu) */
size_t aoc_kmeans_workspace_bytes(int max_rows_per_object, int O, int kmax);
/* scipy.cluster.vq.kmeans2(X_i, k, 'points', iter) per object + centroid_avg (matching.py:533-595, cluster_num :507) in
 * ONE persistent cooperative launch (grid-wide barriers between assignment and update; row tiles resident in shared
 * memory when the bank fits the chip).  kmax = 16 (the reference's cluster_num) or AOC_KMEANS_MAX_K = 64 is the
 * compile-time width the call runs at; kk[o] <= kmax clusters for object o (0 = none), init_idx [O][kmax] object-local
 * rows drawn by the host RNG.  Outputs: cent [O][kmax][100], labels (int32 per sorted bank row),
 * P [O][2*kmax+4][100] (slots 0..kmax-1 centroids, kmax..2*kmax-1 centroid_avg, 2*kmax the mean proxy written by
 * aoc_head_pool_f32) and pvalid [O][2*kmax+4]. */
int aoc_kmeans_proxies_f32(const float* S, const int* meta, const int* nat2sorted, const int* kk, const int* init_idx,
                           int O, int max_rows_per_object, int iters, int kmax, float* cent, int* labels, float* P,
                           int* pvalid, void* workspace, size_t ws_bytes, cudaStream_t stream);

/* ---------------------------------------------------------------- output head (head.cu) */
int aoc_dyn_logits_f32(const float* x, const float* wfg, const float* wbg, float* fg, float* bg, float* logits, int O,
                       int HW, int C, int ldx, cudaStream_t stream);
int aoc_upsample_softmax_f32(const float* logits, float* probs, uint8_t* label, int O, int h, int w, int H, int W,
                             cudaStream_t stream);
/* The same upsample + softmax with the eval loop's per-frame label bookkeeping fused behind it (SURVEY 8f rows 1-2;
 * eval_manager_mm.py:252-270 "delete the label that hasn't existed in the GT label", :318-320 argmax, :339-349
 * uncertainty region filter; layers/shannon_entropy.py:10-13).  exist_bits: device int32 word, bit o set <=> label o
 * was seen in a ground-truth frame so far (NULL = every slot).  probs (optional) is always the plain softmax that
 * AOCNet.forward_for_eval returns (aocnet.py:100-107) -- the filter is the CALLER's bookkeeping and only shapes the
 * derived maps: label = argmax over the seen slots (uint8, lowest index on ties; == argmax of probs * keep);
 * conf_label (optional) = 125 where the entropy over the seen slots exceeds unc_ratio, else label -- the "confident"
 * mask the memory bank stores; entropy (optional) [H][W] floats. */
int aoc_upsample_softmax_label_f32(const float* logits, float* probs, uint8_t* label, uint8_t* conf_label,
                                   float* entropy, const int* exist_bits, float unc_ratio, int O, int h, int w, int H,
                                   int W, cudaStream_t stream);

/* ---------------------------------------------------------------- tcgen05 self-test (umma_gemm.cu) */
/* C[M][N] = A[M][K] * B[N][K]^T through the same tcgen05 pipeline as aoc_global_match_tc (M%128==0, N%256==0,
 * K<=104); workspace >= tc images of A (RB 128) and B (RB 256).  variant 1 swaps LBO/SBO (diagnostic). */
int aoc_gemm_tf32x3_test(const float* A, const float* B, float* C, int M, int N, int K, int variant, void* workspace,
                         size_t ws_bytes, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AOCB200_H_ */
