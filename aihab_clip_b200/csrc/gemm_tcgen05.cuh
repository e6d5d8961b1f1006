// Persistent warp-specialised tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
//   A : activations, row-major [M, K] 16-bit (fp16 or bf16)         -> K-major UMMA operand A
//   W : weights in the reference's nn.Linear layout [N, K] 16-bit   -> K-major UMMA operand B
//   accumulate fp32 in TMEM; epilogue variants cover every GEMM site of the reference hot path
//   (clip/model.py:171-175,181,184-185,217-221 ; methods/ProLIP.py:40 ; methods/utils.py:185).
//
// Roles (384 threads, 256 for the residual epilogue; 1 CTA / SM, persistent, static round-robin tile schedule;
// CTA pairs = clusters of 2 with tcgen05 cta_group::2 and UMMA M = 256 when wave quantisation allows):
//   warp 0 (one lane)  TMA producer : cp.async.bulk.tensor 128x64 A box + BNx64 W box (half of it per CTA of a pair)
//                                     per stage, SWIZZLE_128B, 3-5 stage mbarrier ring
//   warp 1 (one lane)  MMA issuer   : 4 x tcgen05.mma (M=128|256, N=BN, K=16) per stage, tcgen05.commit -> barriers
//   warp 2             TMEM allocator (2 accumulator stages x BN fp32 columns: epilogue of tile i overlaps MMAs of i+1)
//   warps 4..11        epilogue     : tcgen05.ld -> bias / LayerNorm fold / QuickGELU in fp32 -> swizzled smem staging
//                                     -> TMA store (16-bit outputs) or 128 B row segments (fp32 outputs);
//                      warps 4..7 only for EPI_BIAS_RES_32: fp32 residual read-modify-write through a ring of 4 KB
//                                     TMA boxes, LayerNorm statistics + gamma * x for the next GEMM
// Every kernel calls griddepcontrol.launch_dependents at entry and griddepcontrol.wait before its first global access
// (programmatic dependent launch along the stream).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace aihab {

enum GemmEpilogue : int {
  EPI_BIAS_16 = 0,       // out16[m,n] = acc + bias[n]
  EPI_BIAS_GELU_16 = 1,  // out16[m,n] = quickgelu(acc + bias[n])             clip/model.py:160-162
  EPI_BIAS_RES_32 = 2,   // out32[m,n] += acc + bias[n]   (fp32 residual stream) clip/model.py:184-185
  EPI_PATCH_32 = 3,      // out32[tok(m),n] = acc + pos[1 + m % g2, n]         clip/model.py:217-221
  EPI_SCALE_32 = 4,      // out32[m,n] = scale * acc (+ bias[n] if bias)       plain fp32 store
  // LayerNorm folded into the consumer GEMM: A holds gamma * x (16-bit, un-normalised), the epilogue applies
  //   out = r_m * acc - r_m * mu_m * s_n + b'_n,  s_n = sum_k gamma_k W_nk,  b'_n = bias_n + sum_k beta_k W_nk,
  // with the row statistics (mu, r = rsqrt(var + 1e-5)) rebuilt from the partial sums the producer GEMM stored.
  EPI_LN_BIAS_16 = 5,       // ln_1 -> in_proj                                  clip/model.py:181,184
  EPI_LN_BIAS_GELU_16 = 6,  // ln_2 -> c_fc -> QuickGELU                        clip/model.py:172,185
  // scale * acc (+ bias) is NOT stored: every epilogue warp keeps, per accumulator row, the TOPK_SLOTS best of the
  // columns it sees (value descending, lower column first among equals = torch.topk order) and writes them to
  // cand_val / cand_idx [M, 2 * n_tiles, TOPK_SLOTS]; a merge pass picks the row's final top-k.  The logits of the
  // scoring path never reach HBM (methods/utils.py:16-21,185-186).
  EPI_TOPK_32 = 7,
  // out16 = [M, 3N] fp16 rows (hi | hi | lo) of acc (+ bias): hi = fp16(v), lo = fp16(v - hi) - the A operand of the exact
  // three-term logits GEMM of aihab_score16 (W' = w_hi | w_lo | w_hi) - and stats_out[m, n / 64] = sum of v^2 over the
  // 64-column chunk.  The rows are NOT normalised: the consumer GEMM scales its accumulator rows by
  // 1 / max(sqrt(sum of the chunks), 1e-12) (GemmParams.row_ss), which is F.normalize applied after the product.
  EPI_SPLIT3_16 = 8,
};
constexpr int TOPK_SLOTS = 8;  // candidates kept per (row, epilogue warp) in EPI_TOPK_32; k <= TOPK_SLOTS

struct GemmParams {
  int M, N, K;
  int ab_format;      // 0 = fp16 operands, 1 = bf16 operands
  int epilogue;       // GemmEpilogue
  const float* bias;  // [N] fp32 or nullptr
  void* out16;        // [M, ldo] fp16/bf16
  float* out32;       // [*, ldo] fp32
  int ldo;            // leading dimension (elements) of the output
  const float* pos;   // EPI_PATCH_32: positional embedding [L, N] fp32
  int g2;             // EPI_PATCH_32: patches per image (L = g2 + 1)
  float scale;        // EPI_SCALE_32
  int reverse_m;      // walk M blocks from last to first (L2 reuse of the producer's freshest rows)
  int debug;          // 0 = normal; 77 = skip the 16-bit epilogues entirely (main-loop ceiling measurement, AIHAB_GEMM_DEBUG)
  // EPI_BIAS_RES_32 as LayerNorm PRODUCER (optional, ln_gamma != nullptr): besides updating the residual it stores
  //   a16_out[m,n] = round16(ln_gamma[n] * x_new[m,n])            (A operand of the next GEMM, ld = N)
  //   stats_out[m, n / 128, 0..1] = (sum, sum of squares) of x_new over 128-column blocks (deterministic order,
  //   independent of tile width and warp layout)
  const float* ln_gamma;
  void* a16_out;
  float* stats_out;
  // EPI_LN_* as LayerNorm CONSUMER: partial sums written by the producer ([M, ln_nsb, 2]) and s_n ([N]); `bias` = b'
  const float* ln_stats;
  int ln_nsb;
  const float* ln_s;
  // EPI_TOPK_32: candidate buffers [M, 2 * ceil(N / BN), TOPK_SLOTS]
  float* cand_val;
  int* cand_idx;
  // EPI_SCALE_32 / EPI_TOPK_32 (optional): per-row chunk sums of squares [M, row_ss_n] from an EPI_SPLIT3_16 producer; the
  // epilogue multiplies `scale` by 1 / max(sqrt(sum), 1e-12) of its row
  const float* row_ss;
  int row_ss_n;
  int topk_k;  // the k the caller needs (1..TOPK_SLOTS; 0 = TOPK_SLOTS): k <= 5 keeps 5 candidates per slot group
  // Pipelined GEMM pair through an L2-resident ring (CTA pairs only; c_fc -> c_proj): the PRODUCER GEMM (ring_mode 1,
  // 16-bit epilogue) stores its output rows modulo ring_rows, counts every finished epilogue warp of a 256-row pair-row in
  // ctr_done[pair_row] once its TMA stores are complete, and before overwriting a ring slot waits until the consumer has
  // counted need_consumed tiles of the pair-row ring_rows / 256 earlier in ctr_consumed.  The CONSUMER GEMM (ring_mode
  // 2) loads its A rows modulo ring_rows after ctr_done[pair_row] reached need_done and counts every tile whose MMAs
  // have completed in ctr_consumed[pair_row].  The two kernels run CONCURRENTLY on two streams, each on half of the SMs.
  int ring_mode;
  int ring_rows;
  unsigned* ctr_done;
  unsigned* ctr_consumed;
  unsigned need_done;
  unsigned need_consumed;
};

// Fused MLP (c_fc -> QuickGELU -> c_proj + residual as ONE persistent kernel of CTA pairs; gemm_tcgen05.cu
// mlp_fused_kernel).  M must be a multiple of 256 and D of 256.
struct MlpFusedParams {
  int M, D;               // token rows, width (hidden = 4 D)
  int ab_format;
  const uint32_t* tiles;  // device tile list from build_mlp_tiles(): bit 31 = c_proj, bits 8..30 pair-row, bits 0..7 n tile
  int num_tiles;          // entries (rounds x units; MLP_TILE_NONE pads)
  int ring_pairs;         // pair-rows (256 rows each) in the hidden ring
  unsigned* ctr_done;     // [M / 256] c_fc epilogue warps finished per pair-row (zeroed before the launch)
  unsigned* ctr_cons;     // [M / 256] c_proj tiles whose MMAs retired per pair-row (zeroed before the launch)
  unsigned need_done, need_cons;
  const float* fc_bias;   // b' of the LayerNorm fold
  const float* fc_s;      // s_n of the LayerNorm fold
  const float* ln_stats;  // [M, ln_nsb, 2] from the previous residual GEMM
  int ln_nsb;
  const float* proj_bias;
  const float* ln_gamma;  // optional LayerNorm producer for the next block (nullptr: none)
  void* a16_out;
  float* stats_out;
};
// Tile list for `units` CTA pairs: entry r * units + u is the r-th tile of pair u.  c_proj tiles follow the c_fc tiles
// of their pair-row by >= lag rounds, are spread evenly over the pairs, and one full round of them is kept for the end
// so that the tail is balanced; a c_fc tile is dealt only after the c_proj tiles of the pair-row ring_pairs earlier, so
// every wait points to an earlier round (tools/probes/mlp_tiles_sim.py replays the schedule).
void build_mlp_tiles(int pair_rows, int n_fc, int n_proj, int units, int ring_pairs, std::vector<uint32_t>* out, int lag = 3,
                     int extra = 4);
// tmap_y: A of c_fc {64,128}; tmap_wfc / tmap_wproj: W halves {64,128}; tmap_hld: ring as A of c_proj {64,128};
// tmap_x: make_tmap_2d_f32_box32 over the residual.  The store map over the ring is built here from ring_base.
cudaError_t launch_mlp_fused(const CUtensorMap& tmap_y, const CUtensorMap& tmap_wfc, const CUtensorMap& tmap_hld,
                             const CUtensorMap& tmap_wproj, const CUtensorMap& tmap_x, void* ring_base,
                             const MlpFusedParams& p, int num_sms, cudaStream_t stream);

// Encodes a 2-D tiled tensor map over a row-major [rows, cols] 16-bit matrix with a {64, box_rows} box and
// SWIZZLE_128B.  row_pitch_bytes must be a multiple of 16.  Returns cudaSuccess or an error.
cudaError_t make_tmap_2d_16bit(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                               uint64_t row_pitch_bytes, uint32_t box_rows, int ab_format);

// 3-D map over a 16-bit activation buffer viewed as [n_seq][seq_rows][cols] (row-major, sequences back to back) with a
// {64 cols, box_rows, 1} box and SWIZZLE_128B: a store of a row block that runs past seq_rows is clipped at the end
// of ITS sequence (the attention epilogue's partial last query tile).
cudaError_t make_tmap_3d_16bit_seq(CUtensorMap* map, const void* base, uint64_t n_seq, uint64_t seq_rows, uint64_t cols,
                                   uint32_t box_rows, int ab_format);

// Tensor map over the fp32 residual stream [rows, cols] with a {32 cols, 32 rows} box (4 KB) and SWIZZLE_128B:
// used by EPI_BIAS_RES_32 to load, and store back in place, the residual slice of every epilogue warp.
cudaError_t make_tmap_2d_f32_box32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                                   uint64_t row_pitch_bytes);

// Launches the GEMM.  tmap_a box = {64,128}; tmap_w box = {64,BN} with BN = gemm_block_n(N, M);
// tmap_c (EPI_BIAS_RES_32 only, N % 32 == 0) = make_tmap_2d_f32_box32 over out32.
// pair = true (block_n must be 256, tmap_w box = {64,128}): CTA pairs with tcgen05 cta_group::2 (UMMA M = 256).
cudaError_t launch_gemm(const CUtensorMap& tmap_a, const CUtensorMap& tmap_w, const CUtensorMap* tmap_c,
                        const GemmParams& p, int block_n, int num_sms, cudaStream_t stream, bool pair = false);

// Tile-N policy shared by map construction and launch.
int gemm_block_n(int M, int N, int num_sms);
// CTA-pair policy: true when 256-wide tiles are used and pairing M blocks does not cost more in wave quantisation
// than the ~5 % it gains per tile.
bool gemm_use_pair(int M, int N, int num_sms);

cudaError_t gemm_init();  // sets max dynamic smem attributes; resolves cuTensorMapEncodeTiled

}  // namespace aihab
