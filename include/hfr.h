/* hfr.h - C ABI of libhfr.so: the B200-native replacement for the numeric back ends under the hot path of
 * av-savchenko/HSE_FaceRec_tf (TensorFlow `sess.run`, Keras `model.predict`, scikit-learn 1-NN).
 *
 * The reference has no FFI of its own (it is Python calling TF/Keras/sklearn); each entry point below names the
 * reference call site it stands in for (paths relative to the reference repository).  Plain pointers and sizes only;
 * device pointers are CUDA device memory on the handle's device, `stream` is a cudaStream_t passed as void*.
 *
 * Error convention: every function returns 0 on success or a negative hfr_status; the message is available from
 * hfr_last_error() (thread-local).  There is NO CPU fallback: compute entry points fail with HFR_ERR_CUDA when no
 * sm_100 device is usable.
 */
#ifndef HFR_H_
#define HFR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hfr_model hfr_model;
typedef struct hfr_knn hfr_knn;

enum hfr_status {
  HFR_OK = 0,
  HFR_ERR_INVALID = -1,     /* bad argument (ValueError in the Python layer)                         */
  HFR_ERR_IO = -2,          /* file missing / unreadable                                             */
  HFR_ERR_FORMAT = -3,      /* not a GraphDef / HDF5 file, or malformed                              */
  HFR_ERR_NOT_FOUND = -4,   /* tensor name not in graph (KeyError from graph.get_tensor_by_name)     */
  HFR_ERR_UNSUPPORTED = -5, /* op / pattern outside the hot path                                     */
  HFR_ERR_CUDA = -6,        /* CUDA runtime / driver failure, or no usable GPU                       */
  HFR_ERR_STATE = -7        /* call order (e.g. query before set_gallery)                            */
};

enum hfr_precision {
  HFR_FP32 = 0, /* fp32 storage, fp32 CUDA-core GEMMs (exact-parity mode)        */
  HFR_TF32 = 1, /* fp32 storage, tcgen05 kind::tf32 GEMMs, fp32 accumulate       */
  HFR_BF16 = 2  /* bf16 storage, tcgen05 kind::f16 (bf16) GEMMs, fp32 accumulate */
};

enum hfr_input_dtype {
  HFR_IN_F32 = 0, /* float32 NHWC, already pre-processed: what the reference feeds the placeholder               */
  HFR_IN_U8 = 1   /* uint8 NHWC RGB crops; pre-processing (flags below) is fused into the first kernel           */
};

/* hfr_model_forward flags.  Pre-processing of facerec_test.py:96-110 / facial_analysis.py:102-107, for HFR_IN_U8. */
#define HFR_FLAG_BGR 0x1            /* convert2BGR: reverse channel order                                */
#define HFR_FLAG_MEAN_IMAGENET 0x2  /* subtract (103.939, 116.779, 123.68)   (imageNetUtilsMean=True)     */
#define HFR_FLAG_MEAN_VGGFACE2 0x4  /* subtract (91.4953, 103.8827, 131.0912) (imageNetUtilsMean=False)   */
#define HFR_FLAG_SCALE_PM1 0x8      /* x / 127.5 - 1                          (convert2BGR=False)         */
#define HFR_FLAG_L2NORM 0x10        /* L2-normalise output 0 (sklearn.preprocessing.normalize, facerec_test.py:405) */
#define HFR_FLAG_CUDA_GRAPH 0x20    /* replay the step from a captured CUDA graph (stream must not be the legacy stream) */

const char* hfr_last_error(void);
int hfr_version(void);
/* Number of kernels this library has launched in this process (all handles); for `gpu_launches` accounting. */
int64_t hfr_launch_count(void);

/* ---- model: load + compile -------------------------------------------------------------------------------------
 * Replaces load_graph + TensorFlowInference.__init__ (facerec_test.py:41-78), FacialImageProcessing.load_graph_def /
 * load_age_gender (facial_analysis.py:83-92,319-325) and Keras load_weights('models/vgg2_mobilenet.h5')
 * (facerec_test.py:326-334).  `path` is a frozen GraphDef (.pb) or a Keras HDF5 file (.h5, MobileNet-v1 layer names).
 * `output_names_csv`: comma-separated tensor names, e.g. "age_pred/Softmax:0,gender_pred/Sigmoid:0,global_pooling/Mean:0".
 * `phase_name` (nullable): learning-phase placeholder fed with `phase_value` (learning_phase_tensor /
 * additional_input_value).  `input_hw`: 0 = the placeholder's static size; otherwise run the fully convolutional
 * body at this size.  `device` < 0 parses and compiles on the host only (no CUDA needed): the handle then supports
 * hfr_model_info / hfr_model_plan_json but not forward. */
int hfr_model_load(const char* path, const char* input_name, const char* output_names_csv, const char* phase_name,
                   float phase_value, int input_hw, int device, int precision, hfr_model** out);
int hfr_model_info(const hfr_model* m, int* in_h, int* in_w, int* in_c, int* n_outputs, int* out_dims /*[n_outputs]*/);
/* Fused layer plan as JSON (for tests / inspection): layers, outputs, and "arena" = the activation arena's layout, one
 * [value id, producer layer, byte offset per image, bytes per image, layer after which the slot is free (-1: never)]
 * per materialised value.  Returns the length needed (incl. NUL) if buf is too small. */
int64_t hfr_model_plan_json(const hfr_model* m, char* buf, int64_t buf_len);

/* Folded fp32 weights / bias of plan layer `layer` as the compiler produced them (layouts: see "w" in the plan JSON
 * kinds - stem [kh][kw][3][cout], dw [9][c], pw/conv [cout][kh*kw][cin], fc [k][n]).  Returns the weight element
 * count; copies only when the capacities suffice.  Host-side, for checking the graph compiler without a GPU. */
int64_t hfr_model_layer_weights(const hfr_model* m, int layer, float* w, int64_t w_cap, float* bias, int64_t b_cap);

/* Replaces sess.run(outputs, {input: x}) (facerec_test.py:120, facial_analysis.py:109) on a whole batch.
 * x: device pointer, NHWC [batch, H, W, 3] of `in_dtype`.  outs[i]: device pointer to float32 [batch, out_dims[i]]. */
int hfr_model_forward(hfr_model* m, const void* x, int in_dtype, int batch, int flags, void* const* outs, void* stream);
/* Same call with HOST buffers (numpy arrays in, numpy arrays out - the reference's calling convention): copies the
 * batch host->device, runs, copies the outputs device->host and synchronises the stream. */
int hfr_model_forward_host(hfr_model* m, const void* x_host, int in_dtype, int batch, int flags, void* const* outs_host,
                           void* stream);
/* The same call split in two for callers that stream many batches (the reference's loop over a dataset,
 * facerec_test.py:394): submit returns as soon as the H2D copy, the forward and the D2H copy of the batch are enqueued -
 * the copies on the handle's own copy streams, the forward on `stream` - so the next batch's upload and the previous
 * batch's download overlap the current batch's compute.  Up to HFR_HOST_SLOTS batches may be in flight, one per slot;
 * outs_host[i] are valid after hfr_model_wait_host(m, slot).  x_host / outs_host should be page-locked for the copies to
 * be asynchronous.  Submitting to a slot that has not been waited for is an error. */
#define HFR_HOST_SLOTS 4
int hfr_model_submit_host(hfr_model* m, int slot, const void* x_host, int in_dtype, int batch, int flags,
                          void* const* outs_host, void* stream);
int hfr_model_wait_host(hfr_model* m, int slot);
/* Debug: give every activation its own buffer (no arena reuse) so hfr_model_debug_layer can read any of them. */
int hfr_model_set_keep_activations(hfr_model* m, int keep);
/* Intermediate activation of layer `layer_index` (plan order) from the last forward, converted to float32 NHWC on the
 * device; for layer-by-layer parity checks.  Returns the element count per image, or a negative status. */
int64_t hfr_model_debug_layer(hfr_model* m, int layer_index, int batch, float* dst, void* stream);
/* Per-layer device timing: while enabled, forward runs eagerly with a CUDA-event pair around every layer launch (on
 * the caller's stream) and accumulates the durations; get returns the sums in ms (plan order) and the step count. */
int hfr_model_set_layer_timing(hfr_model* m, int enable);
int hfr_model_get_layer_times(const hfr_model* m, double* ms_per_layer, int* steps);
void hfr_model_free(hfr_model* m);

/* Input staging: crop `n` boxes out of RGB uint8 frames [n_frames, frame_h, frame_w, 3] and resize each crop to
 * out_h x out_w, bit-exactly like cv2.resize(face_img, (w, h)) (INTER_LINEAR, uint8) in age_gender_fun
 * (facial_analysis.py:95) applied to face_img = img[y1:y2, x1:x2] (facial_analysis.py:267).  boxes: device int32
 * [n][5] = (frame index, x1, y1, x2, y2), already clamped to the frame with x2 > x1, y2 > y1.  out: [n,out_h,out_w,3]. */
int hfr_crop_resize_u8(const uint8_t* frames, int n_frames, int frame_h, int frame_w, const int32_t* boxes, int n,
                       uint8_t* out, int out_h, int out_w, int device, void* stream);
/* Host-side launch heuristics, exposed so that they can be tested without a GPU: the tile the GEMM launcher uses for an
 * m x n x k problem on a device with `sms` SMs (ctas = 1 or 2 CTAs per tile, block_n = 64 / 128 / 256; conv_taps = kh*kw
 * of an implicit-GEMM convolution, 0 for a plain GEMM), and the gallery split of the k-NN kernel. */
int hfr_debug_gemm_tile_choice(int64_t m, int n, int k, int conv_taps, int sms, int* ctas, int* block_n);
int hfr_debug_knn_plan(int64_t nq, int64_t n, int* splits, int* n_blocks_per_unit);
/* ... and whether two dependent 1x1 convolutions  y = act([a0 | a] w1^T (+ r)) [m, n1],  z = act(y w2^T) [m, n2]  run as
 * one gemm_pair_kernel launch (k0 = columns of the concatenated operand, 0: none), and with which configuration: staging
 * buffers per epilogue warpgroup, residual prefetch distance, A buffers, 16 KB ring slots, dynamic shared memory. */
int hfr_debug_gemm_pair_config(int64_t m, int k0, int k1, int n1, int n2, int has_residual, int precision, int* eligible,
                               int* nbuf, int* pf, int* na, int* stages, int* smem_bytes);
/* Distance matrix for the clustering scripts: out[i,j] = ||x_i - y_j||_2 (float32 [n,m], device), the per-pair
 * expression of process_photos.py:46-48 and sklearn.metrics.pairwise_distances(X_norm) of facial_clustering_test.py:396;
 * y == NULL: y = x (m == n, exact zeros on the diagonal).  With year/born (device float32 [n] / [m], all four or none):
 * out = max(0, dist + age_weight * (a_i - a_j)^2 / (a_i + a_j)), a = max(year_i, year_j) - born: the album variant of
 * process_photos.py:49-56 (age_weight = 0.1 there).  Linkage itself stays on the host. */
int hfr_pairwise_dist(const float* x, int64_t n, const float* y, int64_t m, int dim, const float* year_x,
                      const float* born_x, const float* year_y, const float* born_y, float age_weight, float* out,
                      int device, void* stream);
/* Input staging, the other resize the reference uses: scipy.misc.imresize(img, (w, h), interp='bilinear')
 * (facerec_test.py:84,93) = Pillow's Image.resize(BILINEAR) on the uint8 array (antialiasing triangle filter, 22-bit
 * fixed-point coefficients, horizontal pass rounded to uint8 before the vertical one) - bit-exact.  `images`: device
 * buffer holding n RGB uint8 images of arbitrary sizes; desc_host: HOST int64 [n][4] = (byte offset of pixel (0,0),
 * height, width, row pitch in bytes) - a crop is an offset plus the parent's pitch.  out: device [n,out_h,out_w,3].
 * Reductions beyond 31x are HFR_ERR_UNSUPPORTED. */
int hfr_resize_pil_u8(const uint8_t* images, const int64_t* desc_host, int n, uint8_t* out, int out_h, int out_w,
                      int device, void* stream);
/* age = 1 + sum_{i in top2} i * p_i / sum_{top2} p   (facial_analysis.py:113-124); age_probs [batch, n] float32. */
int hfr_age_gender_post(const float* age_probs, int batch, int n, float* age_out, int device, void* stream);
/* sklearn.preprocessing.normalize(X, norm='l2') (facerec_test.py:262,265,405); in-place allowed (y == x). */
int hfr_l2_normalize(const float* x, float* y, int64_t n, int dim, int device, void* stream);

/* ---- 1-NN identification ------------------------------------------------------------------------------------------
 * Replaces KNeighborsClassifier(n_neighbors=1, p=2).fit / .kneighbors (facerec_test.py:272,284-285,422) - the euclidean
 * ArgKmin reduction; label lookup stays with the caller.  A handle owns one gallery shard.
 *
 * Exactness contract.  The result of a query equals the brute-force answer computed in double precision from the
 * float32 rows (what scikit-learn's ArgKmin computes), ties going to the lowest gallery index:
 *   1. the tensor-core distance GEMM (bf16 or tf32 operands) proposes candidates - the best 2 (n_neighbors = 1) or 4
 *      approximate scores per 16k-row gallery split and epilogue warpgroup;
 *   2. the best 4 (8) candidates per query are re-scored in fp64 from the float32 originals;
 *   3. a query is CERTIFIED when every row that was not re-scored has an approximate score above the k-th exact
 *      distance by more than the rigorous rounding bound of step 1 (operand rounding 2u|q||g|, u = 2^-9 bf16 / 2^-10
 *      tf32, plus fp32 accumulation and norm terms); otherwise the query is re-scored against the WHOLE shard in fp64
 *      (knn_exact_kernel).  hfr_knn_stats reports how many queries took that second path. */
typedef struct hfr_neighbor {
  double dist2;   /* squared euclidean distance, accumulated in fp64 from the float32 rows */
  int64_t index;  /* global gallery row (global_row_offset + local row); -1: no such neighbour */
} hfr_neighbor;
int hfr_knn_create(int device, int dim, int precision /* HFR_TF32 or HFR_BF16 */, hfr_knn** out);
/* gallery: device float32 [n_local, dim], must stay alive and unchanged while the handle uses it (it is re-read for
 * the exact re-scoring).  global_row_offset: index of the shard's first row in the full gallery. */
int hfr_knn_set_gallery(hfr_knn* k, const float* gallery, int64_t n_local, int64_t global_row_offset, void* stream);
/* queries: device float32 [nq, dim] -> out: device hfr_neighbor [nq][n_neighbors], 1 <= n_neighbors <= 4 (the
 * reference's classifier list holds n_neighbors = 1 and 3, facerec_test.py:272-275), ascending by (dist2, index) like
 * sklearn's kneighbors; (inf, -1) where the shard has fewer rows. */
int hfr_knn_query(hfr_knn* k, const float* queries, int64_t nq, int n_neighbors, hfr_neighbor* out, void* stream);
/* Host-buffer variant for the end-to-end path (copies in, runs, copies out, synchronises). */
int hfr_knn_query_host(hfr_knn* k, const float* queries_host, int64_t nq, int n_neighbors, hfr_neighbor* out_host,
                       void* stream);
/* Merge per-shard results gathered as [n_parts][nq][n_neighbors] (e.g. by ONE NCCL all-gather of the packed records)
 * into the global n_neighbors nearest per query: ascending (dist2, index), ties -> lowest index. */
int hfr_knn_merge(const hfr_neighbor* parts, int n_parts, int64_t nq, int n_neighbors, hfr_neighbor* out, int device,
                  void* stream);
/* Row-sharded gallery, exact and shard-independent (what KNeighborsClassifier(sharded=True) runs; the merge above only
 * combines finished per-shard answers).  Certifying per shard would send every query whose neighbour lives in ANOTHER
 * shard through this shard's fp64 pass; instead
 *   1. hfr_knn_query_partial: out [nq][n_neighbors + 1] = the shard's re-scored candidates (ascending, (inf,-1) padded)
 *      followed by one record {lower bound on dist2 of every row of the shard that was not re-scored, index -2};
 *   2. all-gather the records ([n_parts][nq][n_neighbors + 1], one collective), hfr_knn_merge_certify: global nearest of
 *      all candidates into out [nq][n_neighbors]; a query whose n_neighbors-th distance is not strictly below the
 *      smallest bound is appended to unc_list (device int32 [nq]); *unc_count (device int32) receives the count -
 *      identical on every rank, so every rank takes the same branch;
 *   3. only if the count is not zero: hfr_knn_query_exact fills the listed rows of a local [nq][n_neighbors] array
 *      with the shard's fp64 brute-force answer, the arrays are all-gathered and hfr_knn_merge_listed overwrites the
 *      listed rows of `out`. */
int hfr_knn_query_partial(hfr_knn* k, const float* queries, int64_t nq, int n_neighbors, hfr_neighbor* out, void* stream);
int hfr_knn_merge_certify(const hfr_neighbor* parts, int n_parts, int64_t nq, int n_neighbors, hfr_neighbor* out,
                          int32_t* unc_list, int32_t* unc_count, int device, void* stream);
int hfr_knn_query_exact(hfr_knn* k, const float* queries, int64_t nq, int n_neighbors, const int32_t* unc_list,
                        const int32_t* unc_count, hfr_neighbor* out, void* stream);
int hfr_knn_merge_listed(const hfr_neighbor* parts, int n_parts, int64_t nq, int n_neighbors, const int32_t* unc_list,
                         const int32_t* unc_count, hfr_neighbor* out, int device, void* stream);
/* Counters of the last hfr_knn_query on this handle (synchronises its stream): queries certified by the error bound /
 * queries re-scored against the whole shard in fp64. */
int hfr_knn_stats(hfr_knn* k, int64_t* certified, int64_t* rescored);
/* Debug: the approximate candidate records of the last query, [nq][records] each (testing the certification bound);
 * returns the number of records per query; score / index may be NULL to ask for the count only. */
int64_t hfr_knn_debug_candidates(hfr_knn* k, float* score_host, int32_t* index_host, int64_t nq);
void hfr_knn_free(hfr_knn* k);

/* ---- MTCNN face detector networks (scope row 8f-4) -------------------------------------------------------------------
 * Replaces the three sess.run lambdas of FacialImageProcessing.load_mtcnn (facial_analysis.py:334-352) over the
 * reference's mtcnn.pb; the cascade around them (image pyramid, NMS, box regression: facial_analysis.py:354-604) is host
 * code.  net: 0 = P-Net (any h x w), 1 = R-Net (24 x 24), 2 = O-Net (48 x 48).  x: device float32 [n, h, w, 3] exactly as
 * the reference feeds the placeholder ((pixel - 127.5) * 0.0078125, width-major as mtcnn_detect_faces transposes it).
 * Outputs (device float32, slots in the order load_mtcnn binds them):
 *   P-Net  out0 = pnet/conv4-2/BiasAdd [n, ho, wo, 4]   out1 = pnet/prob1 [n, ho, wo, 2]
 *   R-Net  out0 = rnet/conv5-2 [n, 4]                   out1 = rnet/prob1 [n, 2]
 *   O-Net  out0 = onet/conv6-2 [n, 4]   out1 = onet/conv6-3 [n, 10]   out2 = onet/prob1 [n, 2]
 * hfr_mtcnn_out_shape gives (ho, wo, channels) of an output slot for an input size. */
typedef struct hfr_mtcnn hfr_mtcnn;
int hfr_mtcnn_load(const char* mtcnn_pb_path, int device, hfr_mtcnn** out);
int hfr_mtcnn_out_shape(const hfr_mtcnn* m, int net, int h, int w, int slot, int* oh, int* ow, int* oc);
int hfr_mtcnn_run(hfr_mtcnn* m, int net, const float* x, int n, int h, int w, float* out0, float* out1, float* out2,
                  void* stream);
void hfr_mtcnn_free(hfr_mtcnn* m);

/* ---- single operators (kernel-level parity tests and profiling) ---------------------------------------------------
 * dtype: HFR_FP32/HFR_TF32 -> float32 tensors, HFR_BF16 -> bfloat16 tensors.  act: 0 none, 1 relu, 2 relu6. */
int hfr_op_dwconv3x3(const void* x, const float* w9c, const float* bias, void* y, int batch, int h, int w, int c,
                     int stride, int pad_t, int pad_l, int ho, int wo, int act, int dtype, int device, void* stream);
/* y[M,N] = act(a[M,K] * b[N,K]^T + bias (+ residual[M,N])) */
int hfr_op_gemm_bias_act(const void* a, const void* b, const float* bias, const void* residual, void* y, int64_t m,
                         int n, int k, int act, int dtype, int device, void* stream);
/* The fused forms of the ResNet bottleneck seams (testing hooks of gemm_pair.cuh / GemmParams::kb_split):
 *   y[M,N1] = act1([a0 | a] * w1^T + bias1 (+ residual)),   w1 = [N1, K0 + K1]  (a0 == NULL: no concatenated operand);
 *   z == NULL:  that single K-concatenated GEMM;   otherwise also  z[M,N2] = act2(y * w2^T + bias2),  w2 = [N2, N1],
 *   both GEMMs in ONE launch of gemm_pair_kernel (HFR_ERR_UNSUPPORTED when the shapes are not eligible). */
int hfr_op_gemm_pair(const void* a0, int k0, const void* a, int k1, const void* w1, const float* bias1, const void* residual,
                     void* y, int64_t m, int n1, int act1, const void* w2, const float* bias2, void* z, int n2, int act2,
                     int dtype, int device, void* stream);
int hfr_op_stem_conv(const void* x, int in_dtype, const float* w, const float* bias, void* y, int batch, int h, int w_,
                     int kh, int kw, int stride, int pad_t, int pad_l, int ho, int wo, int cout, int flags, int act,
                     int dtype, int device, void* stream);

/* Stride-2 stem convolution on the tensor cores (bf16 out, uint8 RGB in, pre-processing `flags` folded into the
 * weights): w_host is a HOST pointer [kh][kw][3][cout] fp32; synchronises the stream before returning. */
int hfr_op_stem_conv_tc(const void* x_u8, const float* w_host, const float* bias, void* y, int batch, int h, int w_,
                        int kh, int kw, int pad_t, int pad_l, int ho, int wo, int cout, int flags, int act, int device,
                        void* stream);
/* Stride-1 KHxKW convolution, bf16, cout 32|64, through the smem-window kernel (conv_window.cuh).  w_host: HOST pointer
 * [cout][kh*kw][cin] fp32; synchronises the stream before returning. */
int hfr_op_conv2d_window(const void* x, const float* w_host, const float* bias, void* y, int batch, int h, int w_,
                         int cin, int kh, int kw, int pad_t, int pad_l, int ho, int wo, int cout, int act, int device,
                         void* stream);
/* KxK convolution (square kernel, implicit GEMM on tensor cores; tf32/bf16 only). w: [cout][kh*kw][cin] of dtype. */
int hfr_op_conv2d(const void* x, const void* w, const float* bias, const void* residual, void* y, int batch, int h,
                  int w_, int cin, int kh, int kw, int stride, int pad_t, int pad_l, int ho, int wo, int cout, int act,
                  int dtype, int device, void* stream);
int hfr_op_maxpool(const void* x, void* y, int batch, int h, int w, int c, int k, int stride, int pad_t, int pad_l,
                   int ho, int wo, int explicit_zero, int dtype, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HFR_H_ */
