// Launchers for the non-GEMM kernels of the hot path (all sm_100a, all on a caller-supplied stream).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace aihab {

// LayerNorm over rows of D fp32 (eps 1e-5, affine, fp32 statistics; clip/model.py:151-157).
//   x      : fp32, row r starts at x + r * ldx
//   cls0   : optional fp32 [D]; when non-null, rows with r % L == 0 take cls0 as their input instead of x
//            (class_embedding + positional_embedding[0], clip/model.py:220-221)
//   out32  : optional fp32 [rows, D] (may alias x when ldx == D)
//   out16  : optional fp16/bf16 [rows, D]
// D must be a multiple of 128 and <= 2048.
cudaError_t launch_layernorm(const float* x, size_t ldx, const float* cls0, int L, const float* gamma,
                             const float* beta, float* out32, void* out16, int out_bf16, int rows, int D,
                             cudaStream_t stream);

// Two chained LayerNorms in one pass (ln_pre -> ln_1 of the first block, clip/model.py:222,181): out32 = LN(x; g1, b1) (may
// alias x), out16 = LN(out32; g2, b2); bit-identical to two launch_layernorm calls.
cudaError_t launch_layernorm2(const float* x, size_t ldx, const float* cls0, int L, const float* g1, const float* b1,
                              float* out32, const float* g2, const float* b2, void* out16, int out_bf16, int rows, int D,
                              cudaStream_t stream);

// im2col for the stride-p patch embedding (clip/model.py:204,217): images [N,3,R,R] -> rows [N*g*g, Kpad]
// 16-bit with column order (c, ky, kx) matching conv1.weight.reshape(D, 3*p*p); columns >= 3*p*p are zero.
//   in_dtype: 0 = fp32, 1 = fp16, 2 = bf16
cudaError_t launch_im2col(const void* images, int in_dtype, int n, int R, int p, int Kpad, void* out, int out_bf16,
                          cudaStream_t stream);

// Fused single-pass attention for one (image, head) per CTA: softmax(q k^T / 8) v, head dim 64, no mask
// (clip/model.py:179-181; nn.MultiheadAttention with batch_first=False, need_weights=False).
//   qkv : [N*L, 3*D] 16-bit, columns [0,D) = q, [D,2D) = k, [2D,3D) = v, head h at h*64
//   out : [N*L, D] 16-bit
cudaError_t launch_attention(const void* qkv, void* out, int n_img, int L, int H, int is_bf16, cudaStream_t stream);
cudaError_t attention_init(int max_L);
// Same kernel restricted to the query rows >= q_row0 (q_row0 % 16 == 0): the tail of launch_attention_tcf.
cudaError_t launch_attention_rows(const void* qkv, void* out, int n_img, int L, int H, int is_bf16, int q_row0,
                                  cudaStream_t stream);

// Persistent, software-pipelined variant (64 < L <= 224; opt-in for 16 <= L <= 64, where several images share a tile
// behind a block-diagonal mask): one CTA per SM, double-buffered Q/K/V stages and S buffers,
// epilogue of item i-1 overlapped with the PV MMA.  tmap_q: make_tmap_2d_16bit over qkv [n*L, 3D] with a 128-row box;
// tmap_kv: same matrix with a K / V box of
// attention_tcp_key_rows(L) rows.
bool attention_tcp_supported(int L);
int attention_tcp_pack(int L);      // images per packed sequence (1 for L > 64)
int attention_tcp_key_rows(int L);  // rows of the K / V TMA box
// causal != 0: key j is visible to query i only if j <= i (the text tower's attn_mask, clip/model.py:323-329).
// tmap_out3 (optional): 3-D output map for the dual-stream kernel; without it the single-stream kernel runs.
cudaError_t launch_attention_tcp(const CUtensorMap& tmap_q, const CUtensorMap& tmap_kv, void* out, int n_img, int L,
                                 int H, int is_bf16, int num_sms, cudaStream_t stream, int reverse = 0, int causal = 0,
                                 const CUtensorMap* tmap_out3 = nullptr);

// Dual-stream variant for 128 < L <= 224 (the two 128-row query tiles of an (image, head) unit as two streams with one
// softmax thread per row, K / V shared in smem; attention_tcd.cu).  Same tensor maps as launch_attention_tcp; used by it
// for unmasked sequences unless AIHAB_ATTN_DUAL=0.
bool attention_tcd_supported(int L);
// tmap_out: make_tmap_3d_16bit_seq over out viewed as [n_img][L][H * 64] with a 32-row box.
cudaError_t launch_attention_tcd(const CUtensorMap& tmap_q, const CUtensorMap& tmap_kv, const CUtensorMap& tmap_out,
                                 int n_img, int L, int H, int is_bf16, int num_sms, cudaStream_t stream, int reverse = 0);

// Persistent flash-style variant for long sequences (128 < L <= 1024; ViT-L/14: 257, ViT-L/14@336px: 577): keys in
// blocks of <= 160 with an online softmax, O rescaled in TMEM only when the running max moves by more than 2^8.
// tmap_q as above; tmap_kv32: the same matrix with a 32-row box.  A tail of <= 16 query rows beyond the last full
// 128-row tile is computed by launch_attention_rows (needs the raw qkv pointer) after the flash kernel on the same
// stream, or by the caller on another stream when run_tail is false (rows >= L - attention_tcf_tail_rows(L)).
bool attention_tcf_supported(int L);
int attention_tcf_tail_rows(int L);
cudaError_t launch_attention_tcf(const CUtensorMap& tmap_q, const CUtensorMap& tmap_kv32, const void* qkv, void* out,
                                 int n_img, int L, int H, int is_bf16, int num_sms, cudaStream_t stream,
                                 int reverse = 0, bool run_tail = true);

// Launch helper: cudaLaunchKernelEx with an optional cluster dimension and programmatic dependent launch (the kernel
// must call ptx::griddep_wait() before its first global-memory access).  AIHAB_PDL=0 disables the latter.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
cudaError_t launch_kernel(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, int cluster,
                          bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned na = 0;
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl && pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Text tower ends (clip/model.py:341-342, 350): embedding lookup + positional embedding into the fp32 residual stream;
// copy of each sequence's EOT row (tokens.argmax(-1), first maximum) into out [n, D].
cudaError_t launch_embed_tokens(const int64_t* tokens, const float* emb, const float* pos, float* x, long rows, int L,
                                int D, int vocab, cudaStream_t stream);
cudaError_t launch_eot_gather(const int64_t* tokens, const float* x, float* out, int n, int L, int D, cudaStream_t stream);

// fp32 -> 16-bit cast of a weight matrix [rows, cols] into [rows, cols_pad] (zero padded columns).
cudaError_t launch_cast_pad(const float* src, int rows, int cols, void* dst, int cols_pad, int out_bf16,
                            cudaStream_t stream);

// ---- scoring (methods/ProLIP.py:40, methods/utils.py:183-186, aihab_utils/evaluation.py:261-273)
// C[M,N] = alpha * A[M,K] @ B[K,N], all fp32 row-major, fp32 FMA accumulation in k order.
cudaError_t launch_sgemm(const float* A, const float* B, float* C, int M, int N, int K, float alpha,
                         cudaStream_t stream);
// rows / max(||row||_2, eps)   (F.normalize, eps 1e-12); in place allowed.
cudaError_t launch_l2norm(const float* x, float* y, int rows, int cols, float eps, cudaStream_t stream);
// Same for rows of any dtype (0 = fp32, 1 = fp16, 2 = bf16) in and out, 128-bit loads/stores; eps = 0 reproduces
// `f /= f.norm(dim=-1, keepdim=True)` (utils.py:69).  The feature-cache writers normalise with it in the model dtype.
cudaError_t launch_l2norm_rows(const void* x, int in_dtype, void* y, int out_dtype, int rows, int cols, float eps,
                               cudaStream_t stream);
// top-k per row, descending, lowest index first among exact ties (torch.topk / argmax semantics). k <= 16.
cudaError_t launch_topk(const float* logits, int rows, int cols, int k, int64_t* idx, float* val,
                        cudaStream_t stream);

// final top-k over the per-row candidates of the EPI_TOPK_32 GEMM epilogue ([rows, slots, 8] values / int32 columns)
cudaError_t launch_topk_merge(const float* cand_val, const int* cand_idx, int rows, int slots, int k, int64_t* idx,
                              float* val, cudaStream_t stream);

// One-launch small-batch scoring (n <= 8192): proj -> normalise -> logits -> top-k; bit-identical to the chain above.
// mid-size batches (65..8192 rows) with projection and class head: two column-sliced launches (+ the top-k kernel) whose
// results are bit-identical to the chunked sgemm -> l2norm -> sgemm path; emb_raw [n, E] and logits [n, C] are scratch / outputs
bool score_mid_supported(int n, int D, int E, int C);
cudaError_t launch_score_mid(const float* feats, int n, int D, const float* proj, int E, const float* text_w, int C, float scale,
                             float* emb_raw, float* emb_out, float* logits, cudaStream_t stream);
bool score_fused_supported(int n, int D, int E, int C);
cudaError_t launch_score_fused(const float* feats, int n, int D, const float* proj, int E, const float* text_w, int C,
                               float scale, int k, float* emb_out, float* logits_out, int64_t* topk_idx,
                               float* topk_val, cudaStream_t stream);

// L3 -> L2 aggregation (reduce: 0 sum, 1 mean, 2 logsumexp, accumulated in L3-id order) + top-k over the L2 logits +
// top-3 / softmax probabilities over the L3 logits, one launch (aihab_utils/evaluation.py:92-142, 186-221, 261-273).
// l3_to_l2: device int32 [C3]; C3 <= 1024, C2 <= 256; every output pointer may be null (topk_idx only if k == 0).
cudaError_t launch_l2_metrics(const float* logits_l3, int n, int C3, const int* l3_to_l2, int C2, int reduce, int k,
                              float* logits_l2_out, int64_t* topk_idx, float* topk_val, int64_t* top3_idx,
                              float* top3_prob, cudaStream_t stream);

// masked row maxima of a similarity matrix sim [n, P] against prototype owners (tools/outlier_cleaning.py:553-668)
cudaError_t launch_prototype_reduce(const float* sim, const int64_t* labels, const int64_t* owner, int n, int P, int ld,
                                    float* sim_own, int64_t* proto_id, float* sim_other, float* margin,
                                    cudaStream_t stream);

// tensor-core scoring helpers (aihab_score16): 16-bit transpose, fp16 hi/lo splits of the text weights and of the
// normalised embedding (A' = e_hi | e_hi | e_lo against W' = w_hi | w_lo | w_hi reproduces the fp32 product to ~2^-21)
cudaError_t launch_transpose16(const void* src, void* dst, int R, int Cc, cudaStream_t stream);
cudaError_t launch_split_textw(const float* w, int E, int C, void* out, cudaStream_t stream);
cudaError_t launch_l2norm_split(const float* emb, float* emb_out, void* a3, int rows, int E, cudaStream_t stream,
                                int normalize = 1);

// ---- preprocessing (data/clip_transforms.py:50-56; Pillow ImagingResample fixed-point bicubic)
struct ResampleTables {
  // device pointers; *_bounds = {first input index, tap count} per output index; coeffs [out, ksize] int32 (2^22)
  const int* h_bounds;
  const int* h_coeffs;
  int h_ksize;
  const int* v_bounds;
  const int* v_coeffs;
  int v_ksize;
  int new_w, new_h;    // resized image size before the centre crop
  int crop_left, crop_top;
  int need_h, need_v;  // Pillow skips a pass whose input and output sizes are equal
};
// u8 [N,SH,SW,3] -> normalised output.  layout: 0 = NCHW [N,3,R,R]; 1 = im2col rows [N*g*g, Kpad] (16-bit only)
//   out_dtype: 0 = fp32, 1 = fp16, 2 = bf16
cudaError_t launch_preprocess(const uint8_t* in, int n, int sh, int sw, int R, const ResampleTables& t, void* out,
                              int out_dtype, int layout, int p, int Kpad, cudaStream_t stream);
cudaError_t preprocess_init();

}  // namespace aihab
