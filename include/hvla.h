/*
 * hvla.h -- C ABI of the B200-native HyperVLA inference hot path (libhvla.so).
 *
 * The reference (MasterXiong/Hyper-VLA) has no FFI or plugin interface: its boundary is
 * the Python object API of `HyperVLA` (hypervla/model.py:24-137).  These entry points are
 * what a `jax.ffi` / XLA custom-call binding for that path would bind instead of the
 * Flax `apply` calls; each one cites the reference code it replaces.  INTEGRATION.md
 * shows the reference-side stub.
 *
 * Conventions
 *   - plain pointers and sizes only; every buffer is owned by the caller;
 *   - every function returns 0 (HVLA_OK) or a negative HVLA_ERR_* code and never
 *     throws or aborts; `hvla_last_error()` returns a thread-local message;
 *   - all launches are asynchronous on `stream`; no host synchronisation inside
 *     (except the *_host entry points, which say so);
 *   - `dtype` selects the arithmetic: HVLA_F32 = exact fp32 CUDA-core path (parity
 *     <= 1e-5 against the fp32 oracle); HVLA_BF16 = tensor-core path (bf16 operands,
 *     fp32 accumulation; parity <= 2e-2); HVLA_BF16X3 = fp32-class accuracy ON the tensor cores: every DINOv2 matrix
 *     product runs with operands split into two bf16 numbers (hi + lo, three products, fp32 accumulation; csrc/dino_x3.cuh),
 *     generate and the base net on the fp32 path; all buffers as for HVLA_F32 except dino_mat (below);
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     returns HVLA_ERR_CUDA.
 *
 * Blob layouts (element offsets can be queried with hvla_layout_offset()):
 *   HN blob (fp32):   tok_proj_w[768,128] tok_proj_b[128] img_proj_w[768,128] img_proj_b[128]
 *                     task_pos[32,128] img_pos[128] layer_pos[128]
 *                     6 x { ln0_s ln0_b wqkv[128,384] bqkv[384] wo[128,128] bo ln1_s ln1_b
 *                           w0[128,512] b0[512] w1[512,128] b1[128] }  encn_s encn_b
 *   heads:            W[128, NGP] (dtype) and b[NGP] (fp32); NGP = 201504 (201500 used).
 *                     Column order = packed per-task weight row (below).
 *   generated row:    [NGP] (dtype): proj_w[768,64] proj_b[64] pos[257,64]
 *                     4 x { ln0_s ln0_b wq[64,64] bq wk bk wv bv wo[64,64] bo ln1_s ln1_b
 *                           w0[64,128] b0[128] w1[128,64] b1[64] }
 *                     encn_s encn_b wc[64,24] bc[24] wd[64,4] bd[4]
 *   DINO vec (fp32):  patch_b[768] cls[768] pos[257,768]
 *                     12 x { ln1_s ln1_b bqkv[2304] bo ls1 ln2_s ln2_b b1[3072] b2 ls2 } lnf_s lnf_b
 *   DINO mat:         patch_w, 12 x { wqkv wo w1 w2 }.  HVLA_F32: fp32, Flax [K,N] layout,
 *                     patch K padded 588->640.  HVLA_BF16: bf16, transposed [N,K] (K contiguous).
 *                     HVLA_BF16X3: bf16, every matrix [N, 3K] = [hi | hi | lo] of the transposed fp32 matrix, at 3x the offsets.
 */
#ifndef HVLA_H_
#define HVLA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* hvla_stream_t; /* cudaStream_t */

enum { HVLA_OK = 0, HVLA_ERR_ARG = -1, HVLA_ERR_CUDA = -2, HVLA_ERR_UNSUPPORTED = -3, HVLA_ERR_WORKSPACE = -4 };
enum { HVLA_F32 = 0, HVLA_BF16 = 1, HVLA_BF16X3 = 2 };

/* ---- host-only queries (usable without a GPU) -------------------------------------- */
int hvla_version(void);
const char* hvla_last_error(void);
int64_t hvla_hn_blob_elems(void);
int64_t hvla_generated_elems(void);        /* 201500 */
int64_t hvla_generated_row_stride(void);   /* NGP = 201504 */
int64_t hvla_dino_vec_elems(void);
int64_t hvla_dino_mat_elems(void);
/* element offset of a named field: "hn.<field>", "gen.<field>", "dvec.<field>", "dmat.<field>";
 * per-layer fields are "l<k>.<field>".  Returns -1 for an unknown name. */
int64_t hvla_layout_offset(const char* name);
/* bytes of scratch the compute entry points need for B environments and T tasks.  The scratch is owned by ONE call at a time: two calls
 * that may run concurrently (different streams) need different buffers -- besides the activations it holds scheduler state of the
 * persistent GEMM kernels (unit counter, per-row-block completion counters), which every call zeroes itself on its stream.  Its contents
 * need not be preserved or initialised between calls. */
size_t hvla_workspace_bytes(int B, int T, int dtype);

/* ---- generate: replaces HyperNetwork.apply inside HyperVLA.create_tasks -----------------
 * (hypervla/model.py:73-80 -> hypervla/components/hypernetwork.py:99-233).
 *   tok_emb   [T,32,768] f32  instruction_dict.language_instruction.token_embedding
 *   tok_mask  [T,32]     i32  ...attention_mask
 *   lang_pad  [T]        u8   tasks.pad_mask_dict.language_instruction (model.py:69: all ones); NULL = ones
 *   init_cls  [T,768]    f32  initial_state.patch_embeddings[:, 0]   (hypernetwork.py:126)
 *   out_weights [T,NGP] (dtype)  packed generated base-net weights, one row per task
 *   out_ctx   [T,128] f32 or NULL  context embedding (second return of HyperNetwork.__call__)
 *   hn_blob_f16: the HN blob converted element-wise to IEEE fp16 (same offsets); operands of the fused
 *   context-encoder kernel of HVLA_BF16 mode (fp16 keeps the context embedding ~8x closer to the fp32
 *   reference than bf16).  NULL (or HVLA_F32) selects the fp32 CUDA-core path. */
int hvla_generate(hvla_stream_t stream, const float* hn_blob, const void* hn_blob_f16, const void* heads_w,
                  const float* heads_b, const float* tok_emb, const int32_t* tok_mask, const uint8_t* lang_pad,
                  const float* init_cls, int T, void* out_weights, float* out_ctx, void* workspace,
                  size_t workspace_bytes, int dtype);

/* ---- task-switch scheduler: regenerate only the T tasks that switched, IN PLACE ----------------------------
 * (the reference regenerates at every episode reset: data/utils/hypervla_interface.py:141-146,
 *  data/simpler/evaluate.py:263-277).  Inputs as hvla_generate for the T switched tasks; row_index [T] i32 (device):
 *  task t is written to row row_index[t] of the persistent `weights` [T_max,NGP] (dtype) and `ctx` [T_max,128] f32
 *  (or NULL) buffers; other rows are untouched, so a CUDA graph captured over `weights` stays valid.
 *  Rows outside [0, T_max) are skipped. */
int hvla_generate_rows(hvla_stream_t stream, const float* hn_blob, const void* hn_blob_f16, const void* heads_w,
                       const float* heads_b, const float* tok_emb, const int32_t* tok_mask, const uint8_t* lang_pad,
                       const float* init_cls, int T, const int32_t* row_index, int T_max, void* weights, float* ctx,
                       void* workspace, size_t workspace_bytes, int dtype);

/* ---- DINOv2 image encoder: replaces self.image_encoder(raw_images) + normalisation ------
 * (hypervla/components/base_vit.py:111-122; FlaxDinov2Module of transformers==4.50.0).
 *   images  [B,224,224,3] u8 (NHWC)      out_emb [B,257,768] (dtype) = last_hidden_state */
int hvla_dino_forward(hvla_stream_t stream, const float* dino_vec, const void* dino_mat, const uint8_t* images,
                      int B, void* out_emb, void* workspace, size_t workspace_bytes, int dtype);

/* ---- base ViT with per-task generated weights + mix action head -------------------------
 * (hypervla/components/base_vit.py:130-226, transformer.py:127-262, action_heads.py:455-470,
 *  524-538; batching semantics of scripts/train.py:559-579).
 *   emb [B,257,768] (dtype) DINOv2 last_hidden_state (row 0 = CLS is skipped, base_vit.py:122)
 *   weights [T,NGP] (dtype); task_index [B] i32 or NULL (NULL: env b uses row b, needs T==B,
 *   or T==1 shared)   out_action [B,4,7] f32   out_logit [B,4] f32 (gripper logits) or NULL */
int hvla_base_act(hvla_stream_t stream, const void* emb, const void* weights, const int32_t* task_index,
                  int B, int T, float* out_action, float* out_logit, void* workspace, size_t workspace_bytes, int dtype);

/* ---- act = dino_forward + base_act: replaces base_net.apply(method=predict_action) -------
 * inside HyperVLA.sample_actions (hypervla/model.py:125-136). */
int hvla_act(hvla_stream_t stream, const float* dino_vec, const void* dino_mat, const uint8_t* images,
             const void* weights, const int32_t* task_index, int B, int T,
             float* out_action, float* out_logit, void* workspace, size_t workspace_bytes, int dtype);

/* ---- act with the `intermediates` the reference sows (hypervla/model.py:125-137 mutable=['intermediates']; read by
 * InferenceWrapper.save_attention_map, data/utils/hypervla_interface.py:208-217): the softmax attention weights of
 * every DINOv2 layer, dino_maps [12,B,12,257,257] f32 (FlaxDinov2 `outputs.attentions`, base_vit.py:118), and of every base
 * encoder block, base_maps [4,B,4,257,257] f32 (transformer.py:172-191); either may be NULL.  A debugging call: the base net
 * runs on the generic CUDA-core kernels and every map is an extra pass over q|k|v (38 MB + 1 MB of maps per image). */
int hvla_act_debug(hvla_stream_t stream, const float* dino_vec, const void* dino_mat, const uint8_t* images,
                   const void* weights, const int32_t* task_index, int B, int T, float* out_action, float* out_logit,
                   float* dino_maps, float* base_maps, void* workspace, size_t workspace_bytes, int dtype);

/* ---- DiscreteActionHead variant: replaces the same predict_action call when the config says
 * base_net_kwargs.action_head_type = "discrete" (hypervla/components/base_network.py:22-33, 92-99; action_heads.py:252-396;
 * BinTokenizer.decode, octo/model/components/tokenizers.py:235-275).  The base ViT carries n_action_tokens readout tokens
 * (4 = discrete_token_type "action_horizon", 28 = "action_dim_and_action_horizon") and the generated row has its own layout:
 *   proj_w[768,64] proj_b[64] pos[256+A,64] 4 x block (as above) encn_s encn_b vocab_w[64,V] vocab_b[V],  V = 1792 (A=4) | 256 (A=28)
 * hvla_discrete_generated_elems / hvla_discrete_row_stride give its size / padded stride; hvla_generate_n is hvla_generate with an
 * explicit row stride (heads_w [128,row_stride], heads_b [row_stride], out_weights [T,row_stride]).
 *   out_action [B,4,7] f32 = bin centres of the argmax tokens; out_tokens [B,4,7] i32; out_top2 [B,4,7,2] f32 (largest and second
 *   largest logit of every slot, for margin-aware parity checks) or NULL.  Runs the base net on the generic CUDA-core kernels. */
int64_t hvla_discrete_generated_elems(int n_action_tokens);
int64_t hvla_discrete_row_stride(int n_action_tokens);
int hvla_generate_n(hvla_stream_t stream, const float* hn_blob, const void* hn_blob_f16, const void* heads_w, const float* heads_b,
                    const float* tok_emb, const int32_t* tok_mask, const uint8_t* lang_pad, const float* init_cls, int T,
                    int64_t row_stride, void* out_weights, float* out_ctx, void* workspace, size_t workspace_bytes, int dtype);
int hvla_act_discrete(hvla_stream_t stream, const float* dino_vec, const void* dino_mat, const uint8_t* images, const void* weights,
                      const int32_t* task_index, int B, int T, int n_action_tokens, float* out_action, int32_t* out_tokens,
                      float* out_top2, void* workspace, size_t workspace_bytes, int dtype);

/* ---- same, HOST image/action buffers (pinned or pageable): H2D copy, act, D2H copy on
 * `stream`, then ONE cudaStreamSynchronize.  The call InferenceWrapper.step makes
 * (data/utils/hypervla_interface.py:197-207: model call followed by the host read). */
int hvla_act_host(hvla_stream_t stream, const float* dino_vec, const void* dino_mat, const uint8_t* images_host,
                  const void* weights, const int32_t* task_index, int B, int T,
                  float* out_action_host, float* out_logit_host, void* workspace, size_t workspace_bytes, int dtype);

/* ---- building blocks exported for tests / profiling -----------------------------------
 * C[M,N] (bf16) = act(A[M,K] (bf16, row-major) * Wt[N,K]^T (bf16) + bias[N] (f32));
 * act: 0 none, 2 erf-GELU.  The tcgen05/TMEM/TMA GEMM used by the DINOv2 blocks. */
int hvla_gemm_bf16(hvla_stream_t stream, const void* A, const void* Wt, const float* bias, void* C,
                   int M, int N, int K, int act);
/* DINOv2 self-attention alone: qkv [B*257, 2304] bf16 (q | k | v, q pre-divided by sqrt(64)) -> out [B*257, 768] bf16.
 * impl 0 = warp-level mma.sync kernel, 1 = tcgen05/TMEM kernel (the one the pipeline uses). */
int hvla_dino_attention(hvla_stream_t stream, const void* qkv, void* out, int B, int impl);
/* ---- batched action post-processing: replaces the host code after the model call in
 * InferenceWrapper.step (data/utils/hypervla_interface.py:219-300) and BatchActionEnsembler
 * (data/utils/action_ensemble.py:6-27) for B environments at once.
 *   raw_action [B,4,7] f32 (device)   state [B, hvla_postprocess_state_floats()] f32 (device, zero-initialised)
 *   reset [B] u8 or NULL (1 = first step of an episode)      stat_a/stat_b/mask: HOST arrays of 7
 *   norm_type 0 = NORMAL (a=std, b=mean), 1 = BOUNDS (a=p01, b=p99); policy_setup 0 google_robot, 1 widowx_bridge, 2 libero
 *   out_raw [B,7] un-normalised (ensembled) action; out_action [B,7] = world_vector | rot_axangle | gripper */
int64_t hvla_postprocess_state_floats(void);
int hvla_postprocess(hvla_stream_t stream, const float* raw_action, float* state, const uint8_t* reset, int B, int norm_type,
                     const float* stat_a, const float* stat_b, const uint8_t* mask, int ensemble, float temp,
                     int policy_setup, int sticky_repeat, float* out_raw, float* out_action);
/* ---- per-step image preprocessing: replaces InferenceWrapper._resize_image
 * (data/utils/hypervla_interface.py:89-121: tf.image.resize lanczos3 antialias -> optional centre crop_and_resize
 * -> round/clip/uint8) for B camera frames at once.
 *   images [B,H,W,3] u8 (device) -> out [B,S,S,3] u8 (device; the buffer hvla_act reads when S = 224)
 *   starts_y, starts_x, weights_y, weights_x (device): span start and normalised lanczos3 weights per output row / column,
 *     weights_y [S, span_y], weights_x [S, span_x] (computed once per input size by the caller: hvla/preprocess.py)
 *   crop 0/1; crop_params (HOST, 4 floats): y1*(S-1), x1*(S-1), height_scale, width_scale of the crop box
 *   workspace (device): hvla_resize_workspace_bytes(B,H,W,S,crop) */
size_t hvla_resize_workspace_bytes(int B, int H, int W, int S, int crop);
int hvla_resize_lanczos3(hvla_stream_t stream, const uint8_t* images, int B, int H, int W, int S, const int32_t* starts_y,
                         const float* weights_y, int span_y, const int32_t* starts_x, const float* weights_x, int span_x, int crop,
                         const float* crop_params, uint8_t* out, void* workspace, size_t workspace_bytes);
/* ---- T5-base token embedder: replaces FlaxT5EncoderModel('t5-base')(input_ids, attention_mask)
 * .last_hidden_state, the `token_embedding` the hypernetwork consumes (octo/model/components/tokenizers.py:186-211,
 * data/utils/language_tokenizer.py:9-28).  fp32.
 *   t5_blob (device, hvla_t5_blob_elems() floats, HF torch [out,in] weight layout, packed by hvla/t5.py):
 *     shared[32128,768] | 12 x { ln0[768] wq|wk|wv[2304,768] wo[768,768] ln1[768] wi[3072,768] wo2[768,3072] } | final_ln[768]
 *   pos_bias (device) [12,S,S]: relative-position bias per head, query, key (block 0's bucket table expanded by the caller)
 *   input_ids, attention_mask (device) [T,S] i32, S <= 32      out_emb (device) [T,S,768] f32 */
int64_t hvla_t5_blob_elems(void);
size_t hvla_t5_workspace_bytes(int T, int S);
int hvla_t5_encode(hvla_stream_t stream, const float* t5_blob, const float* pos_bias, const int32_t* input_ids, const int32_t* attention_mask,
                   int T, int S, float* out_emb, void* workspace, size_t workspace_bytes);
/* The same encoder with every GEMM on the tcgen05 kernel (the fp32 entry above runs them on CUDA cores: exact, but 6.7 ms for
 * one instruction and 21 ms for 64).  Operands are split into two bf16 numbers (hi + lo) and A W^T is accumulated in fp32 as
 * Ahi Whi^T + Alo Whi^T + Ahi Wlo^T ("bf16x3": 3e-5 of the fp64 result after 12 blocks) in ONE launch per matrix, the three
 * products concatenated along K.  LayerNorm weights and the embedding table still come from t5_blob;
 *   t5_mat (device, hvla_t5_mat_elems() bf16, packed by hvla/t5.py: split_matrices): 12 x { wq|wk|wv[2304,3*768] wo[768,3*768]
 *   wi[3072,3*768] wo2[768,3*3072] }, every row = [hi | hi | lo] of the HF torch [out,in] weight row. */
int64_t hvla_t5_mat_elems(void);
size_t hvla_t5_tc_workspace_bytes(int T, int S);
int hvla_t5_encode_tc(hvla_stream_t stream, const float* t5_blob, const void* t5_mat, const float* pos_bias, const int32_t* input_ids,
                      const int32_t* attention_mask, int T, int S, float* out_emb, void* workspace, size_t workspace_bytes);
/* number of kernels launched by this library since load (for bench.py's gpu_launches) */
int64_t hvla_launch_count(void);
/* per-kernel-class CUDA-event timing on the launching stream (bench.py's live roofline numbers).
 * enable(1) clears and starts recording an event pair around every launch; report() synchronises
 * and writes one "name launches total_ms" line per kernel class. */
int hvla_profile_enable(int on);
int hvla_profile_report(char* buf, size_t cap);

/* ---- legacy XLA GPU custom-call wrappers (jax<=0.4.30: xla_client.register_custom_call_target;
 * signature void(cudaStream_t, void** buffers, const char* opaque, size_t opaque_len)).
 * `opaque` = struct hvla_xla_opaque.  Buffer order = the argument order above (inputs, then
 * outputs, then workspace).  jax is absent from this image, so tests/test_gpu_parity.py drives both targets through
 * ctypes exactly as XLA would (pointer array + packed opaque) and checks them bit-for-bit against the direct calls. */
struct hvla_xla_opaque { int32_t B, T, dtype, reserved; uint64_t workspace_bytes; };
void hvla_xla_generate(void* stream, void** buffers, const char* opaque, size_t opaque_len);
void hvla_xla_act(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* Status-returning form (xla_client.register_custom_call_target(..., api_version=1)): the trailing argument is XLA's
 * `XlaCustomCallStatus*` (opaque here -- no XLA header needed).  On failure (short opaque, bad argument, CUDA error) the
 * message is reported through XlaCustomCallStatusSetFailure -- resolved as a weak symbol from the hosting process, or set
 * explicitly with hvla_xla_register_status_setter -- and the outputs are zero-filled on `stream`. */
typedef void (*hvla_xla_status_setter)(void* status, const char* message, size_t message_len);
void hvla_xla_register_status_setter(hvla_xla_status_setter fn);
void hvla_xla_generate_status(void* stream, void** buffers, const char* opaque, size_t opaque_len, void* status);
void hvla_xla_act_status(void* stream, void** buffers, const char* opaque, size_t opaque_len, void* status);

#ifdef __cplusplus
}
#endif
#endif /* HVLA_H_ */
