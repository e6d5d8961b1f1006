/*
 * aihab_clip.h — C ABI of libaihab_clip.so: the B200-native (sm_100a) implementation of the aihab-clip
 * image-encode + zero-shot scoring hot path.
 *
 * The reference (WhiteGiveFive/aihab-clip) is pure Python/PyTorch and has no FFI of its own; its boundary for
 * this path is the Python object protocol of the model returned by clip.load (clip/clip.py:89-137).  Each entry
 * point below names the reference call it replaces.  The Python mirror of that protocol lives in
 * aihab_clip_b200/clip/ and binds these symbols with ctypes (see INTEGRATION.md for the reference-side stub).
 *
 * Conventions
 *   - every function returns 0 on success and a non-zero code on failure; aihab_last_error() returns a
 *     thread-local message for the last failure on the calling thread.  No C++ exception crosses the ABI.
 *   - all data pointers are DEVICE pointers unless the parameter is documented "host or device"; the caller owns
 *     every input/output buffer; a handle owns its packed weights and workspace.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); work is enqueued, not awaited.
 *   - a handle must not be used from two threads at the same time.
 */
#ifndef AIHAB_CLIP_H_
#define AIHAB_CLIP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIHAB_ABI_VERSION 1

#if defined(__GNUC__)
#define AIHAB_API __attribute__((visibility("default")))
#else
#define AIHAB_API
#endif

/* element types */
enum { AIHAB_F32 = 0, AIHAB_F16 = 1, AIHAB_BF16 = 2 };

/* GEMM epilogues (aihab_gemm16) */
enum {
  AIHAB_EPI_BIAS_16 = 0,      /* out16 = acc + bias                                  (in_proj, clip/model.py:181) */
  AIHAB_EPI_BIAS_GELU_16 = 1, /* out16 = quickgelu(acc + bias)                       (c_fc + QuickGELU, :162,172) */
  AIHAB_EPI_BIAS_RES_32 = 2,  /* out32 += acc + bias                                 (out_proj / c_proj + residual, :184-185) */
  AIHAB_EPI_PATCH_32 = 3,     /* out32[token] = acc + positional_embedding           (conv1 + pos-emb, :217-221) */
  AIHAB_EPI_SCALE_32 = 4      /* out32 = scale * acc + bias                          (plain fp32 store) */
};

typedef struct aihab_vit aihab_vit;

/* Geometry of the image tower: VisionTransformer.__init__ (clip/model.py:199-214). heads = width / 64 (:267). */
typedef struct {
  int image_size;  /* input_resolution R            */
  int patch_size;  /* p                              */
  int width;       /* D, multiple of 128             */
  int layers;      /* number of ResidualAttentionBlock */
  int heads;       /* D / 64                         */
  int dtype;       /* AIHAB_F16 or AIHAB_BF16: tensor-core operand / activation format (fp32 accumulate, fp32 residual) */
  int max_batch;   /* images processed per internal chunk (sizes the workspace) */
} aihab_vit_config;

/* One ResidualAttentionBlock (clip/model.py:165-186); fp32, reference state_dict shapes; host or device. */
typedef struct {
  const float* ln_1_weight;     /* [D]      */
  const float* ln_1_bias;       /* [D]      */
  const float* in_proj_weight;  /* [3D, D]  rows ordered q, k, v */
  const float* in_proj_bias;    /* [3D]     */
  const float* out_proj_weight; /* [D, D]   */
  const float* out_proj_bias;   /* [D]      */
  const float* ln_2_weight;     /* [D]      */
  const float* ln_2_bias;       /* [D]      */
  const float* c_fc_weight;     /* [4D, D]  */
  const float* c_fc_bias;       /* [4D]     */
  const float* c_proj_weight;   /* [D, 4D]  */
  const float* c_proj_bias;     /* [D]      */
} aihab_vit_block_weights;

/* visual.* tensors of the reference state_dict (clip/model.py:204-214); fp32; host or device. */
typedef struct {
  const float* conv1_weight;         /* [D, 3, p, p] (no bias) */
  const float* class_embedding;      /* [D]                    */
  const float* positional_embedding; /* [L, D], L = (R/p)^2+1  */
  const float* ln_pre_weight;        /* [D] */
  const float* ln_pre_bias;          /* [D] */
  const float* ln_post_weight;       /* [D] */
  const float* ln_post_bias;         /* [D] */
  const aihab_vit_block_weights* blocks; /* [layers], host array */
} aihab_vit_weights;

/* Text tower: CLIP.encode_text (clip/model.py:338-353) on the same GEMM / LayerNorm / attention
 * kernels with the causal mask of build_attention_mask (:323-329).  heads = width / 64. */
typedef struct aihab_text aihab_text;
typedef struct {
  int context_length; /* 77; 65..224 supported (persistent tcgen05 attention) */
  int vocab_size;     /* 49408 */
  int width;          /* transformer_width, multiple of 128 */
  int layers;
  int heads;          /* width / 64 */
  int dtype;          /* AIHAB_F16 or AIHAB_BF16 */
  int max_batch;      /* prompts per internal chunk */
} aihab_text_config;
typedef struct {
  const float* token_embedding;      /* [vocab, width] */
  const float* positional_embedding; /* [context_length, width] */
  const float* ln_final_weight;      /* [width] */
  const float* ln_final_bias;        /* [width] */
  const aihab_vit_block_weights* blocks; /* transformer.resblocks.{i}.*, [layers], host array */
} aihab_text_weights;

/* Library / device ------------------------------------------------------------------------------------- */
AIHAB_API int aihab_abi_version(void);
AIHAB_API const char* aihab_last_error(void);
/* number of kernels this library has launched in this process (all handles, all streams) */
AIHAB_API uint64_t aihab_kernel_launches(void);

/* Optional per-kernel-class timing (CUDA events recorded on the launch stream around every launch of the class).
 * Classes: 0 = tcgen05 GEMM, 1 = attention, 2 = LayerNorm, 3 = im2col / preprocess, 4 = scoring.
 * aihab_profile_read sums elapsed milliseconds, launches and algorithmic work (FLOPs for 0/1/4, bytes for 2/3)
 * recorded since the last reset; it synchronises on the recorded events. */
AIHAB_API int aihab_profile_enable(int on);
AIHAB_API int aihab_profile_read(int cls, double* ms, uint64_t* launches, double* work, int reset);
/* Per launch site: groups the records of class `cls` by (algorithmic work per launch, site tag; for a GEMM 2*M*N*K and
 * N, so one group per GEMM shape of the step) and writes work per launch, tag, summed milliseconds and launch count of
 * up to `cap` groups.  Returns the number of groups, -1 on error.  Does not reset; call it before
 * aihab_profile_read(reset). */
AIHAB_API int aihab_profile_sites(int cls, double* work_per_launch, int* tag, double* ms, uint64_t* launches, int cap);

/* Image tower -------------------------------------------------------------------------------------------
 * Replaces build_model(...).visual construction + convert_weights (clip/model.py:372-433): packs the weights
 * into 16-bit K-major tensor-core layout, builds TMA descriptors, allocates the workspace on `device`. */
AIHAB_API int aihab_vit_create(const aihab_vit_config* cfg, const aihab_vit_weights* w, int device, aihab_vit** out);
AIHAB_API void aihab_vit_destroy(aihab_vit* h);
AIHAB_API size_t aihab_vit_workspace_bytes(const aihab_vit* h);
/* Images per call (<= max_batch, >= max_batch / 2) whose transformer GEMMs fill `device`'s SMs with whole waves of
 * 256 x 256 tiles: the batch size to feed aihab_vit_encode* / the extraction loop (methods/utils.py:142-173 uses a
 * fixed 16).  tokens = (image_size / patch_size)^2 + 1, width = transformer width.  No handle needed. */
AIHAB_API int aihab_preferred_batch(int tokens, int width, int max_batch, int device);

/* Replaces CLIP.encode_image / VisionTransformer.forward (clip/model.py:216-235, 335-336).
 *   images   : [n, 3, R, R] NCHW, element type in_dtype, already normalised (output of the preprocess)
 *   feats_out: [n, D] PRE-projection features ln_post(x[:,0,:]) — visual.proj is NOT applied (model.py:228-235)
 * n may exceed max_batch; the call loops over chunks. */
AIHAB_API int aihab_vit_encode(aihab_vit* h, const void* images, int in_dtype, int n, void* feats_out, int out_dtype,
                     void* stream);

/* Same, fed by raw uint8 HWC images: the eval preprocessing of data/clip_transforms.py:50-56 is fused in front
 * (bit-exact Pillow bicubic resize of the short side to R, centre crop, /255, CLIP mean/std).
 *   images_u8: [n, sh, sw, 3] */
AIHAB_API int aihab_vit_encode_u8(aihab_vit* h, const uint8_t* images_u8, int n, int sh, int sw, void* feats_out,
                        int out_dtype, void* stream);

/* Preprocessing alone.  Replaces build_clip_transforms(is_train=False) (data/clip_transforms.py:50-56) and
 * clip._transform (clip/clip.py:74-81) for uint8 arrays.  out: [n, 3, R, R] of out_dtype. */
AIHAB_API int aihab_preprocess_u8(const uint8_t* images_u8, int n, int sh, int sw, int R, void* out, int out_dtype,
                        void* stream);

/* Scoring --------------------------------------------------------------------------------------------------
 * Replaces VisProjViT.forward (methods/ProLIP.py:31-41), F.normalize (methods/utils.py:184),
 * `100. * f @ text_weights` (methods/utils.py:185) and argmax / topk (methods/utils.py:16-21,186;
 * aihab_utils/evaluation.py:261-273), all in fp32.
 *   feats   : [n, D] fp32 pre-projection features
 *   proj    : [D, E] fp32 (visual.proj) or NULL to score `feats` directly (then E = D)
 *   text_w  : [E, C] fp32 class text embeddings (columns = classes), or NULL to stop after normalisation
 *   scale   : logit temperature (100 in the reference)
 *   emb_out : [n, E] fp32 L2-normalised embeddings, nullable
 *   logits_out: [n, C] fp32, nullable
 *   topk_idx: [n, k] int64 sorted descending, lowest index first among exact ties, nullable when k == 0
 *   topk_val: [n, k] fp32, nullable */
AIHAB_API int aihab_score(const float* feats, int n, int D, const float* proj, int E, const float* text_w, int C, float scale,
                int k, float* emb_out, float* logits_out, int64_t* topk_idx, float* topk_val, void* stream);

/* Row L2 normalisation y = x / max(||x||_2, eps) for rows of fp32 / fp16 / bf16 (dtype codes as above), fp32
 * statistics, in place allowed when the dtypes match.  Replaces F.normalize(feats, dim=-1) in the cache writers
 * (aihab_utils/feature_cache.py:126-127, methods/utils.py:39, eps = 1e-12) and, with eps = 0,
 * `image_features /= image_features.norm(dim=-1, keepdim=True)` (utils.py:69). */
AIHAB_API int aihab_l2_normalize(const void* x, int in_dtype, int rows, int cols, float eps, void* y, int out_dtype,
                                 void* stream);

/* Same scoring over CACHED 16-bit features on the tensor cores (ProLIP / linear-probe scoring over a feature cache,
 * methods/ProLIP.py:288-293; the reference caches features in the model dtype, fp16 on GPU — feature_cache.py:215).
 *   feats16 : [n, D] fp16/bf16 (dtype)          proj16 : [D, E] same dtype (visual.proj after convert_weights)
 *   text_w  : [E, C] fp32, C % 4 == 0
 * feats16 @ proj16 runs as a tcgen05 GEMM with fp32 accumulation: products of 16-bit values are exact in fp32, so
 * the result differs from the fp32 reference only by summation order.  The logits GEMM multiplies fp16 hi/lo splits
 * of the embedding and of the text weights (e_hi w_hi + e_hi w_lo + e_lo w_hi, relative error ~2^-21): of the
 * NORMALISED embedding when emb_out is requested, else of the raw one with F.normalize applied to the accumulator row
 * (scale / max(||e||, 1e-12)) - the same value up to fp32 rounding.  Top-k indices equal the fp32 reference's
 * wherever its scores are untied.  Outputs as in aihab_score. */
AIHAB_API int aihab_score16(const void* feats16, int n, int D, int dtype, const void* proj16, int E, const float* text_w,
                            int C, float scale, int k, float* emb_out, float* logits_out, int64_t* topk_idx,
                            float* topk_val, void* stream);

/* Text tower: create / destroy as for the image tower.  aihab_text_encode replaces the transformer part of
 * CLIP.encode_text (clip/model.py:341-350): tokens [n, context_length] int64 on the device ->
 * feats_out [n, width] = ln_final(x)[arange(n), tokens.argmax(-1)], i.e. the reference's x_before_proj
 * (the caller applies @ text_projection, :351).  n may exceed max_batch. */
AIHAB_API int aihab_text_create(const aihab_text_config* cfg, const aihab_text_weights* w, int device, aihab_text** out);
AIHAB_API void aihab_text_destroy(aihab_text* h);
AIHAB_API int aihab_text_encode(aihab_text* h, const int64_t* tokens, int n, void* feats_out, int out_dtype, void* stream);

/* Metrics epilogue after the logits, one launch (aihab_utils/evaluation.py): aggregate_logits_to_l2 (:92-142,
 * reduce 0 = sum, 1 = mean, 2 = logsumexp, accumulated in L3-id order like the reference loop), the top-k over the
 * L2 logits that L2MetricsAccumulator.update takes (:186-221), and top3_metrics (:261-273): top-3 indices of the L3
 * logits with their softmax probabilities.  l3_to_l2: device int32 [C3] (the L3 -> L2 label map, 20 -> 11 as shipped).
 * Outputs (all nullable except topk_idx when k > 0): logits_l2_out [n,C2], topk_idx [n,k] int64 / topk_val [n,k],
 * top3_idx [n,3] int64, top3_prob [n,3].  C3 <= 1024, C2 <= 256, k <= C2. */
AIHAB_API int aihab_l2_metrics(const float* logits_l3, int n, int C3, const int32_t* l3_to_l2, int C2, int reduce, int k,
                               float* logits_l2_out, int64_t* topk_idx, float* topk_val, int64_t* top3_idx,
                               float* top3_prob, void* stream);

/* Multi-prototype / centroid scoring over cached embeddings (tools/outlier_cleaning.py:553-668; with one prototype
 * per class it is the centroid similarity of :296-337): sim = emb @ prototypes^T in fp32, then per row the best
 * similarity among the prototypes of its own class (prototype_id = index inside the class block, first maximum), the
 * best among all other classes (NaN when there is no other class) and margin = own - other.
 *   emb [n,E] fp32 L2-normalised, labels [n] int64, prototypes_t [E,P] fp32 (TRANSPOSED, class blocks contiguous),
 *   owner [P] int64 = class id of each prototype.  prototype_id / sim_to_other / margin may be null. */
AIHAB_API int aihab_prototype_scores(const float* emb, const int64_t* labels, int n, int E, const float* prototypes_t,
                                     const int64_t* owner, int P, float* sim_to_prototype, int64_t* prototype_id,
                                     float* sim_to_other, float* margin, void* stream);

/* Building blocks (exported for per-kernel parity tests; same kernels the tower uses) ---------------------- */
/* D[M,N] = A[M,K] * W[N,K]^T with a fused epilogue; A, W 16-bit (ab_dtype), K % 8 == 0. */
AIHAB_API int aihab_gemm16(const void* A, const void* W, int M, int N, int K, int ab_dtype, int epilogue, const float* bias,
                 void* out16, float* out32, int ldo, const float* pos, int g2, float scale, void* stream);
/* LayerNorm rows of D fp32 (eps 1e-5) -> out32 (nullable) and/or out16 (nullable, out16_dtype). */
AIHAB_API int aihab_layernorm(const float* x, int rows, int D, const float* gamma, const float* beta, float* out32,
                    void* out16, int out16_dtype, void* stream);
/* softmax(q k^T / 8) v per (image, head); qkv [n*L, 3*H*64] 16-bit -> out [n*L, H*64]. */
AIHAB_API int aihab_attention(const void* qkv, void* out, int n, int L, int H, int dtype, void* stream);
/* Same with the text tower's causal mask (key j visible to query i only for j <= i); 64 < L <= 224. */
AIHAB_API int aihab_attention_causal(const void* qkv, void* out, int n, int L, int H, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AIHAB_CLIP_H_ */
