// C ABI of libaihab_clip.so (see include/aihab_clip.h).  Host-side orchestration only: weight packing, TMA
// descriptor construction, workspace layout, the per-layer launch sequence and Pillow's coefficient tables.
#include "../../include/aihab_clip.h"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "gemm_tcgen05.cuh"
#include "kernels.cuh"

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

// ---- optional per-kernel-class timing with CUDA events on the launch stream (bench.py roofline evidence)
enum ProfClass { PC_GEMM = 0, PC_ATTN = 1, PC_LN = 2, PC_PRE = 3, PC_SCORE = 4, PC_COUNT = 5 };
struct ProfRec {
  cudaEvent_t a, b;
  int cls;
  double work;
  int aux;  // site tag within a class (GEMM: N)
};
std::mutex g_prof_mu;
std::atomic<bool> g_prof_on{false};
std::vector<ProfRec> g_prof_recs;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_free;

struct ProfScope {
  bool on = false;
  ProfRec r{};
  cudaStream_t s;
  ProfScope(int cls, double work, cudaStream_t stream, int aux = 0) : s(stream) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_free.empty()) {
      r.a = g_prof_free.back().first;
      r.b = g_prof_free.back().second;
      g_prof_free.pop_back();
    } else if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    r.cls = cls;
    r.work = work;
    r.aux = aux;
    on = true;
    cudaEventRecord(r.a, s);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(r.b, s);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back(r);
  }
};

int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}
#define CK(expr)                                                                                          \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) {                                                                              \
      char _b[512];                                                                                       \
      snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return fail(_b);                                                                                    \
    }                                                                                                     \
  } while (0)
// a launcher that enqueues exactly one kernel
#define CKL(expr)     \
  do {                \
    CK(expr);         \
    g_launches += 1;  \
  } while (0)

int round_up(int v, int m) { return (v + m - 1) / m * m; }

cudaMemPool_t temp_pool(int dev);

// Stream-ordered temporary that is returned to the pool on EVERY exit path (the CK macros return early on errors).
template <typename T>
struct AsyncTemp {
  T* p = nullptr;
  cudaStream_t s;
  explicit AsyncTemp(cudaStream_t stream) : s(stream) {}
  AsyncTemp(const AsyncTemp&) = delete;
  AsyncTemp& operator=(const AsyncTemp&) = delete;
  cudaError_t alloc(size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaMemPool_t pool = temp_pool(dev);
    return pool ? cudaMallocFromPoolAsync(reinterpret_cast<void**>(&p), bytes, pool, s)
                : cudaMallocAsync(reinterpret_cast<void**>(&p), bytes, s);
  }
  ~AsyncTemp() {
    if (p != nullptr) cudaFreeAsync(p, s);
  }
};

// attention kernel choice: 3 = persistent flash tcgen05 (L > 224), 2 = persistent tcgen05 (64 < L <= 224; with
// AIHAB_ATTN_PACK=1 also 16 <= L <= 64, images packed block-diagonally), 0 = mma.sync (any L <= 908).
// AIHAB_ATTN=legacy|tcp caps the choice (A/B measurements); default picks the fastest supported kernel.
int attention_kind(int L) {
  int cap = 3;
  if (const char* e = getenv("AIHAB_ATTN")) {
    if (!strcmp(e, "legacy")) cap = 0;
    else if (!strcmp(e, "tcp")) cap = 2;
  }
  if (cap >= 2 && aihab::attention_tcp_supported(L)) return 2;
  if (cap >= 3 && aihab::attention_tcf_supported(L)) return 3;
  return 0;
}
int attention_key_box(int kind, int L) {
  (void)kind;
  return kind == 3 ? 32 : aihab::attention_tcp_key_rows(L);
}

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int device_of(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice) return a.device;
  cudaGetLastError();
  int d = 0;
  cudaGetDevice(&d);
  return d;
}

// Stream-ordered temporaries (aihab_score / aihab_score16 / aihab_prototype_scores) come from a PRIVATE memory pool per
// device whose freed blocks stay cached (release threshold = max) - the process-wide default pool is left alone.
cudaMemPool_t temp_pool(int dev) {
  static std::mutex mu;
  static std::map<int, cudaMemPool_t> pools;
  std::lock_guard<std::mutex> lk(mu);
  auto it = pools.find(dev);
  if (it != pools.end()) return it->second;
  cudaMemPool_t pool = nullptr;
  cudaMemPoolProps props{};
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = dev;
  if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
    uint64_t threshold = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  } else {
    cudaGetLastError();
    pool = nullptr;  // fall back to the default pool (cudaMallocAsync)
  }
  pools[dev] = pool;
  return pool;
}
void keep_pool_warm(int dev) { (void)temp_pool(dev); }

// CTA-pair (tcgen05 cta_group::2) GEMM tiles: AIHAB_GEMM_PAIR=1/0 forces them on/off, default = gemm_use_pair()
bool gemm_pair_enabled(int M, int N, int sms) {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("AIHAB_GEMM_PAIR");
    v = (e != nullptr) ? (e[0] != '0') : -1;
  }
  if (v >= 0) return v != 0 && aihab::gemm_block_n(M, N, sms) == 256;
  return aihab::gemm_use_pair(M, N, sms);
}

int sm_count(int dev) {
  static std::mutex mu;
  static std::map<int, int> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(dev);
  if (it != cache.end()) return it->second;
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  cache[dev] = n;
  return n;
}

// ------------------------------------------------------------------------------------------------------------
// Pillow ImagingResample coefficient tables (third-party: Pillow 12.2.0, src/libImaging/Resample.c —
// precompute_coeffs + normalize_coeffs_8bpc; restated from the published algorithm).  One axis.
struct AxisTable {
  std::vector<int> bounds;  // {xmin, count} per output index
  std::vector<int> coeffs;  // [out, ksize], 22-bit fixed point
  int ksize = 0;
};

double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

AxisTable build_resample_axis(int in_size, int out_size) {
  AxisTable t;
  const double scale = static_cast<double>(in_size) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  t.ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  t.bounds.assign(static_cast<size_t>(out_size) * 2, 0);
  t.coeffs.assign(static_cast<size_t>(out_size) * t.ksize, 0);
  std::vector<double> k(t.ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
    }
    for (int x = 0; x < xmax; ++x) {
      const double v = k[x] * (1 << 22);
      t.coeffs[static_cast<size_t>(xx) * t.ksize + x] = static_cast<int>(v < 0 ? -0.5 + v : 0.5 + v);
    }
    t.bounds[2 * xx] = xmin;
    t.bounds[2 * xx + 1] = xmax;
  }
  return t;
}

struct DeviceTables {
  int* h_bounds = nullptr;
  int* h_coeffs = nullptr;
  int* v_bounds = nullptr;
  int* v_coeffs = nullptr;
  aihab::ResampleTables t{};
  uint64_t last_use = 0;
};

constexpr size_t kMaxTableSets = 32;  // distinct (device, input size, R) combinations kept; least recently used evicted
uint64_t g_tab_clock = 0;
std::mutex g_tab_mu;
std::map<std::tuple<int, int, int, int>, DeviceTables> g_tables;  // (device, sh, sw, R)

int upload_ints(const std::vector<int>& v, int** out) {
  CK(cudaMalloc(out, v.size() * sizeof(int)));
  CK(cudaMemcpy(*out, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}

// torchvision v2.Resize(int) output size (short side -> R, long side int(R * long / short)) and v2.CenterCrop
// offsets int(round((size - R) / 2.0)); data/clip_transforms.py:51-52.
int get_tables(int dev, int sh, int sw, int R, aihab::ResampleTables* out) {
  std::lock_guard<std::mutex> lk(g_tab_mu);
  auto key = std::make_tuple(dev, sh, sw, R);
  auto it = g_tables.find(key);
  if (it == g_tables.end()) {
    int new_h, new_w;
    if (sw <= sh) {
      new_w = R;
      new_h = static_cast<int>(static_cast<double>(R) * sh / sw);
    } else {
      new_h = R;
      new_w = static_cast<int>(static_cast<double>(R) * sw / sh);
    }
    if (new_h < R || new_w < R) return fail("preprocess: resized image smaller than the crop");
    DeviceTables d;
    AxisTable th = build_resample_axis(sw, new_w);
    AxisTable tv = build_resample_axis(sh, new_h);
    if (upload_ints(th.bounds, &d.h_bounds) || upload_ints(th.coeffs, &d.h_coeffs) ||
        upload_ints(tv.bounds, &d.v_bounds) || upload_ints(tv.coeffs, &d.v_coeffs))
      return 1;
    d.t.h_bounds = d.h_bounds;
    d.t.h_coeffs = d.h_coeffs;
    d.t.h_ksize = th.ksize;
    d.t.v_bounds = d.v_bounds;
    d.t.v_coeffs = d.v_coeffs;
    d.t.v_ksize = tv.ksize;
    d.t.new_w = new_w;
    d.t.new_h = new_h;
    d.t.crop_top = static_cast<int>(std::nearbyint((new_h - R) / 2.0));   // Python round(): half to even
    d.t.crop_left = static_cast<int>(std::nearbyint((new_w - R) / 2.0));
    d.t.need_h = (new_w != sw);
    d.t.need_v = (new_h != sh);
    if (g_tables.size() >= kMaxTableSets) {  // variable-size image sets must not grow device memory without bound
      auto victim = g_tables.begin();
      for (auto j = g_tables.begin(); j != g_tables.end(); ++j)
        if (j->second.last_use < victim->second.last_use) victim = j;
      // cudaFree synchronises the device, so no kernel still reads the evicted tables
      cudaFree(victim->second.h_bounds);
      cudaFree(victim->second.h_coeffs);
      cudaFree(victim->second.v_bounds);
      cudaFree(victim->second.v_coeffs);
      g_tables.erase(victim);
    }
    it = g_tables.emplace(key, d).first;
  }
  it->second.last_use = ++g_tab_clock;
  *out = it->second.t;
  return 0;
}

}  // namespace

// ================================================================================================ handle
struct aihab_vit {
  aihab_vit_config cfg{};
  int device = 0;
  int num_sms = 148;
  int g = 0, g2 = 0, L = 0, D = 0, Kp = 0, Kpad = 0;
  int bf16 = 0;
  size_t cap_rows = 0;  // max_batch * L

  // packed weights (16-bit, [N, K] K-major) and fp32 vectors
  struct Block {
    void *w_in = nullptr, *w_out = nullptr, *w_fc = nullptr, *w_proj = nullptr;
    float *b_in = nullptr, *b_out = nullptr, *b_fc = nullptr, *b_proj = nullptr;
    float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
    // LayerNorm folded into the consumer GEMM: s_n = sum_k gamma_k W16_nk, b'_n = bias_n + sum_k beta_k W16_nk
    float *s_in = nullptr, *bp_in = nullptr, *s_fc = nullptr, *bp_fc = nullptr;
    CUtensorMap m_in[2], m_out[2], m_fc[2], m_proj[2];  // [0] = 128-row box, [1] = 256-row box
  };
  std::vector<Block> blocks;
  void* w_conv = nullptr;
  CUtensorMap m_conv[2];
  float *pos = nullptr, *cls0 = nullptr, *lnpre_g = nullptr, *lnpre_b = nullptr, *lnpost_g = nullptr,
        *lnpost_b = nullptr;

  // workspace
  void* patches = nullptr;  // [max_batch*g2, Kpad] 16-bit
  float* x = nullptr;       // [cap_rows, D] fp32 residual stream
  void* y = nullptr;        // [cap_rows, D] 16-bit (LN output / attention output)
  void* big = nullptr;      // [cap_rows, 4D] 16-bit (qkv [.,3D] and MLP hidden [.,4D] share it)
  void* y2 = nullptr;       // [cap_rows, D] 16-bit: gamma * x written by the residual epilogues (LayerNorm fold)
  float* ln_stats = nullptr;  // [cap_rows, kMaxStatBlocks, 2] per 128-column block (mean, sum of squared deviations) of the residual rows
  bool ln_fold = true;
  bool zigzag = true;  // consecutive kernels walk the rows in opposite directions (L2 keeps the producer's last rows)
  CUtensorMap m_y2;
  CUtensorMap m_patches, m_y, m_h, m_x;  // m_x: fp32 residual stream, {32,32} boxes (EPI_BIAS_RES_32)
  CUtensorMap m_attn_q, m_attn_kv;       // qkv view [cap_rows, 3D] of `big` for the tcgen05 attention
  CUtensorMap m_attn_out3;               // attention output `y` as [max_batch][L][D] (dual-stream kernel's TMA stores)
  bool has_attn_out3 = false;
  int attn_kind = 0;
  int causal = 0;  // text tower: key j visible to query i only for j <= i (clip/model.py:323-329)
  // text tower ends (aihab_text_*): embedding table, ln_final, EOT rows
  float *tok_emb = nullptr, *lnf_g = nullptr, *lnf_b = nullptr, *xe = nullptr;
  int vocab = 0;
  // side stream for the few query rows the flash attention leaves to the SIMT row kernel (L = 257): it runs
  // concurrently with the tensor-core kernel, fork / join through the two events
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // pipelined MLP (AIHAB_MLP_PIPE=1): c_fc and c_proj run concurrently on half of the SMs each, the hidden activations
  // pass through an L2-resident ring in the front of `big`
  bool mlp_pipe = false;
  int ring_rows = 0;
  unsigned* mlp_ctr = nullptr;   // [2][cap_rows / 256] progress counters (done | consumed)
  CUtensorMap m_hring;           // A operand of c_proj over the ring
  // AIHAB_MLP_FUSED=1: c_fc + c_proj as ONE kernel (mlp_fused_kernel) with the same ring / counters; tile lists per M
  bool mlp_fused = false;
  bool ln_chain = true;  // ln_pre and ln_1 of the first block in one pass (AIHAB_LN_CHAIN=0: two launches)
  struct TileList {
    std::vector<uint32_t> host;
    uint32_t* dev = nullptr;
  };
  std::map<int, TileList> mlp_tiles;  // key: pair-rows (M / 256)
  size_t ws_bytes = 0;
  std::vector<void*> allocs;
};

namespace {

int dev_alloc(aihab_vit* h, void** p, size_t bytes) {
  CK(cudaMalloc(p, bytes));
  // zero once: masked attention keys (padding rows of a 64-row slot, rows behind the last image of a partial batch) are
  // multiplied by P = 0, which only gives 0 if what they hold is finite
  CK(cudaMemset(*p, 0, bytes));
  h->allocs.push_back(*p);
  h->ws_bytes += bytes;
  return 0;
}

// copy an fp32 tensor (host or device) to a new device buffer
int upload_f32(aihab_vit* h, const float* src, size_t n, float** out) {
  if (src == nullptr) return fail("aihab_vit_create: null weight pointer");
  if (dev_alloc(h, reinterpret_cast<void**>(out), n * sizeof(float))) return 1;
  CK(cudaMemcpy(*out, src, n * sizeof(float), cudaMemcpyDefault));
  return 0;
}

// fp32 [rows, cols] (host or device) -> 16-bit [rows, cols_pad] device
int upload_16(aihab_vit* h, const float* src, int rows, int cols, int cols_pad, void** out) {
  if (src == nullptr) return fail("aihab_vit_create: null weight pointer");
  float* tmp = nullptr;
  const size_t n = static_cast<size_t>(rows) * cols;
  CK(cudaMalloc(&tmp, n * sizeof(float)));
  cudaError_t e = cudaMemcpy(tmp, src, n * sizeof(float), cudaMemcpyDefault);
  if (e == cudaSuccess) {
    if (dev_alloc(h, out, static_cast<size_t>(rows) * cols_pad * 2)) {
      cudaFree(tmp);
      return 1;
    }
    e = aihab::launch_cast_pad(tmp, rows, cols, *out, cols_pad, h->bf16, 0);
    g_launches += 1;
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
  }
  cudaFree(tmp);
  CK(e);
  return 0;
}

// LayerNorm fold constants for a consumer Linear W [N, K] fed by LayerNorm(gamma, beta):
//   s[n] = sum_k gamma[k] * W16[n][k],   bp[n] = bias[n] + sum_k beta[k] * W16[n][k]
// with W16 = W rounded to the tensor-core operand format actually multiplied.  Pointers may be host or device.
int fold_ln(aihab_vit* h, const float* W, const float* bias, const float* gamma, const float* beta, int N, int K,
            float** s_out, float** bp_out) {
  std::vector<float> w(static_cast<size_t>(N) * K), b(N), g(K), be(K), sv(N), bp(N);
  CK(cudaMemcpy(w.data(), W, w.size() * sizeof(float), cudaMemcpyDefault));
  CK(cudaMemcpy(b.data(), bias, N * sizeof(float), cudaMemcpyDefault));
  CK(cudaMemcpy(g.data(), gamma, K * sizeof(float), cudaMemcpyDefault));
  CK(cudaMemcpy(be.data(), beta, K * sizeof(float), cudaMemcpyDefault));
  for (int n = 0; n < N; ++n) {
    double acc_s = 0.0, acc_b = 0.0;
    const float* row = w.data() + static_cast<size_t>(n) * K;
    for (int k = 0; k < K; ++k) {
      const float w16 = h->bf16 ? __bfloat162float(__float2bfloat16_rn(row[k])) : __half2float(__float2half_rn(row[k]));
      acc_s += static_cast<double>(g[k]) * w16;
      acc_b += static_cast<double>(be[k]) * w16;
    }
    sv[n] = static_cast<float>(acc_s);
    bp[n] = static_cast<float>(static_cast<double>(b[n]) + acc_b);
  }
  if (upload_f32(h, sv.data(), N, s_out) || upload_f32(h, bp.data(), N, bp_out)) return 1;
  return 0;
}

int weight_maps(aihab_vit* h, const void* w, int N, int K, CUtensorMap (&m)[2]) {
  CK(aihab::make_tmap_2d_16bit(&m[0], w, N, K, static_cast<uint64_t>(K) * 2, 128, h->bf16));
  CK(aihab::make_tmap_2d_16bit(&m[1], w, N, K, static_cast<uint64_t>(K) * 2, 256, h->bf16));
  return 0;
}

constexpr int kMaxStatBlocks = 8;  // N / 128 for N <= 1024

struct LnOpt {
  const float* gamma = nullptr;  // producer: gamma of the next LayerNorm -> y2 / ln_stats are written
  const float* s = nullptr;      // consumer: s_n; statistics are read from ln_stats
  int nsb = 0;                   // consumer: stat blocks per row written by the producer
};

int run_gemm(aihab_vit* h, const CUtensorMap& ma, const CUtensorMap (&mw)[2], int M, int N, int K, int epi,
             const float* bias, void* out16, float* out32, int ldo, cudaStream_t s, const LnOpt& ln = LnOpt(),
             int* n_blocks_out = nullptr, int reverse = 0) {
  aihab::GemmParams p{};
  p.ln_gamma = ln.gamma;
  p.a16_out = ln.gamma ? h->y2 : nullptr;
  p.stats_out = ln.gamma ? h->ln_stats : nullptr;
  p.ln_stats = ln.s ? h->ln_stats : nullptr;
  p.ln_nsb = ln.nsb;
  p.ln_s = ln.s;
  p.M = M;
  p.N = N;
  p.K = K;
  p.ab_format = h->bf16;
  p.epilogue = epi;
  p.bias = bias;
  p.out16 = out16;
  p.out32 = out32;
  p.ldo = ldo;
  p.pos = h->pos;
  p.g2 = h->g2;
  p.scale = 1.0f;
  p.reverse_m = reverse;
  const int bn = aihab::gemm_block_n(M, N, h->num_sms);
  if (n_blocks_out) *n_blocks_out = (N + 127) / 128;  // LayerNorm statistics: one partial per row and 128 columns
  const bool pair = gemm_pair_enabled(M, N, h->num_sms);  // a pair stages the 256-wide W tile as two 128-row halves
  ProfScope ps(PC_GEMM, 2.0 * M * N * K, s, N);
  CKL(aihab::launch_gemm(ma, mw[(bn == 256 && !pair) ? 1 : 0], epi == aihab::EPI_BIAS_RES_32 ? &h->m_x : nullptr, p, bn,
                         h->num_sms, s, pair));
  return 0;
}

// the residual attention blocks (clip/model.py:165-197) on the M = n * L rows of the fp32 residual stream h->x
// ln1_done: ln_1 of the first block is already in h->y2 (fused into the ln_pre pass by run_tower)
int run_blocks(aihab_vit* h, int n, cudaStream_t s, bool ln1_done = false) {
  const int D = h->D, L = h->L, M = n * L;
  const int layers = static_cast<int>(h->blocks.size());
  int nsb = 0;  // stat blocks per row written by the last residual GEMM
  // Zig-zag traversal: each kernel walks the token rows in the direction opposite to its producer, so the ~100 MB the
  // producer wrote last are still in the 126 MB L2 when the consumer reads them first (results do not depend on order).
  int dir = 0;
  auto next_dir = [&]() {
    const int d = h->zigzag ? dir : 0;
    dir ^= 1;
    return d;
  };
  for (int l = 0; l < layers; ++l) {
    aihab_vit::Block& b = h->blocks[l];
    // x = x + out_proj(attn(in_proj(ln_1(x))))   (clip/model.py:181,184)
    if (l == 0 || !h->ln_fold) {
      if (!(l == 0 && ln1_done)) {
        ProfScope ps(PC_LN, 6.0 * M * D, s);
        CKL(aihab::launch_layernorm(h->x, D, nullptr, 0, b.ln1_g, b.ln1_b, nullptr, h->y2, h->bf16, M, D, s));
      }
      if (run_gemm(h, h->m_y2, b.m_in, M, 3 * D, D, aihab::EPI_BIAS_16, b.b_in, h->big, nullptr, 3 * D, s, LnOpt(), nullptr,
                   next_dir()))
        return 1;
    } else {  // ln_1 folded: y2 = gamma_1 * x and the row statistics came out of the previous c_proj epilogue
      LnOpt o;
      o.s = b.s_in;
      o.nsb = nsb;
      if (run_gemm(h, h->m_y2, b.m_in, M, 3 * D, D, aihab::EPI_LN_BIAS_16, b.bp_in, h->big, nullptr, 3 * D, s, o, nullptr,
                   next_dir()))
        return 1;
    }
    {
      ProfScope ps(PC_ATTN, 4.0 * n * L * L * D, s);
      if (h->attn_kind == 3) {
        const int tail = aihab::attention_tcf_tail_rows(L);
        if (tail && h->side) CK(cudaEventRecord(h->ev_fork, s));
        CKL(aihab::launch_attention_tcf(h->m_attn_q, h->m_attn_kv, h->big, h->y, n, L, h->cfg.heads, h->bf16,
                                        h->num_sms, s, next_dir(), /*run_tail=*/!(tail && h->side)));
        if (tail && h->side) {
          // fork: the leftover rows run beside the tensor-core kernel.  Launched AFTER it so that the persistent CTAs
          // (one per SM) are resident first and the small row CTAs fill the remaining thread slots.
          CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
          CKL(aihab::launch_attention_rows(h->big, h->y, n, L, h->cfg.heads, h->bf16, L - tail, h->side));
          CK(cudaEventRecord(h->ev_join, h->side));
        }
        if (tail && h->side) CK(cudaStreamWaitEvent(s, h->ev_join, 0));
      }
      else if (h->attn_kind == 2)
        CKL(aihab::launch_attention_tcp(h->m_attn_q, h->m_attn_kv, h->y, n, L, h->cfg.heads, h->bf16, h->num_sms, s,
                                        next_dir(), h->causal, h->has_attn_out3 ? &h->m_attn_out3 : nullptr));
      else
        CKL(aihab::launch_attention(h->big, h->y, n, L, h->cfg.heads, h->bf16, s));
    }
    {
      LnOpt o;
      if (h->ln_fold) o.gamma = b.ln2_g;
      if (run_gemm(h, h->m_y, b.m_out, M, D, D, aihab::EPI_BIAS_RES_32, b.b_out, nullptr, h->x, D, s, o, &nsb, next_dir()))
        return 1;
    }
    // x = x + c_proj(quickgelu(c_fc(ln_2(x))))   (clip/model.py:171-175,185)
    if (h->mlp_fused && h->ln_fold && (M % 256) == 0 && M / 256 > h->ring_rows / 256 && h->num_sms >= 2) {
      // ONE kernel: the two GEMMs' tiles interleaved on the whole machine, hidden activations through the L2 ring
      const int pairs = M / 256, units = h->num_sms / 2;
      aihab_vit::TileList& tl = h->mlp_tiles[pairs];
      if (tl.dev == nullptr) {
        const char* el = getenv("AIHAB_MLP_LAG");
        const char* ex = getenv("AIHAB_MLP_EXTRA");
        aihab::build_mlp_tiles(pairs, 4 * D / 256, D / 256, units, h->ring_rows / 256, &tl.host, el ? atoi(el) : 3,
                               ex ? atoi(ex) : 4);
        if (dev_alloc(h, reinterpret_cast<void**>(&tl.dev), tl.host.size() * sizeof(uint32_t))) return 1;
        CK(cudaMemcpyAsync(tl.dev, tl.host.data(), tl.host.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
      }
      CK(cudaMemsetAsync(h->mlp_ctr, 0, 2 * (h->cap_rows / 256 + 1) * sizeof(unsigned), s));
      aihab::MlpFusedParams fp{};
      fp.M = M;
      fp.D = D;
      fp.ab_format = h->bf16;
      fp.tiles = tl.dev;
      fp.num_tiles = static_cast<int>(tl.host.size());
      fp.ring_pairs = h->ring_rows / 256;
      fp.ctr_done = h->mlp_ctr;
      fp.ctr_cons = h->mlp_ctr + h->cap_rows / 256 + 1;
      fp.need_done = static_cast<unsigned>((4 * D / 256) * 2 * 8);  // n tiles x 2 CTAs x 8 epilogue warps
      fp.need_cons = static_cast<unsigned>((D / 256) * 2);          // n tiles x 2 CTAs
      if (getenv("AIHAB_MLP_NOWAIT") != nullptr) fp.need_done = fp.need_cons = 0;  // DEBUG: wrong results, cost of the waits
      fp.fc_bias = b.bp_fc;
      fp.fc_s = b.s_fc;
      fp.ln_stats = h->ln_stats;
      fp.ln_nsb = nsb;
      fp.proj_bias = b.b_proj;
      if (l + 1 < layers) {
        fp.ln_gamma = h->blocks[l + 1].ln1_g;
        fp.a16_out = h->y2;
        fp.stats_out = h->ln_stats;
      }
      {
        ProfScope ps(PC_GEMM, 2.0 * M * (4.0 * D) * D * 2.0, s, -4 * D);
        CKL(aihab::launch_mlp_fused(h->m_y2, b.m_fc[0], h->m_hring, b.m_proj[0], h->m_x, h->big, fp, h->num_sms, s));
      }
      nsb = (D + 127) / 128;
      dir = 1;  // the fused kernel walks forward: its consumer starts from the freshest rows
      continue;
    }
    if (h->mlp_pipe && h->ln_fold && (M % 256) == 0 && M / 256 > h->ring_rows / 256 && h->num_sms >= 4) {
      // Pipelined pair: c_fc (stream s) and c_proj (side stream) run CONCURRENTLY, each on half of the SMs; c_fc's
      // 16-bit hidden rows go through a ring of h->ring_rows rows that stays in L2 (per-pair-row progress counters in
      // global memory order the two kernels), so the [M, 4D] hidden activations never make the HBM round trip.
      const int pairs = M / 256, half = (h->num_sms / 4) * 2;  // SMs per kernel (an even number: CTA pairs)
      unsigned* ctr_done = h->mlp_ctr;
      unsigned* ctr_cons = h->mlp_ctr + h->cap_rows / 256 + 1;
      CK(cudaMemsetAsync(h->mlp_ctr, 0, 2 * (h->cap_rows / 256 + 1) * sizeof(unsigned), s));
      ProfScope ps(PC_GEMM, 2.0 * M * (4.0 * D) * D * 2.0, s, -4 * D);  // both GEMMs as one site (they overlap)
      CK(cudaEventRecord(h->ev_fork, s));
      CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
      aihab::GemmParams p{};
      p.M = M;
      p.ab_format = h->bf16;
      p.scale = 1.0f;
      p.ring_rows = h->ring_rows;
      p.ctr_done = ctr_done;
      p.ctr_consumed = ctr_cons;
      p.need_done = static_cast<unsigned>((4 * D / 256) * 2 * 8);  // n tiles x 2 CTAs x 8 epilogue warps
      p.need_consumed = static_cast<unsigned>((D / 256) * 2);       // n tiles x 2 CTAs
      {  // producer: c_fc (LN fold + bias + QuickGELU) -> ring
        aihab::GemmParams a = p;
        a.ring_mode = 1;
        a.N = 4 * D;
        a.K = D;
        a.epilogue = aihab::EPI_LN_BIAS_GELU_16;
        a.bias = b.bp_fc;
        a.out16 = h->big;
        a.ldo = 4 * D;
        a.ln_stats = h->ln_stats;
        a.ln_nsb = nsb;
        a.ln_s = b.s_fc;
        CKL(aihab::launch_gemm(h->m_y2, b.m_fc[0], nullptr, a, 256, half, s, true));
      }
      {  // consumer: c_proj (+ fp32 residual, LayerNorm producer for the next block) <- ring
        aihab::GemmParams c = p;
        c.ring_mode = 2;
        c.N = D;
        c.K = 4 * D;
        c.epilogue = aihab::EPI_BIAS_RES_32;
        c.bias = b.b_proj;
        c.out32 = h->x;
        c.ldo = D;
        if (l + 1 < layers) {
          c.ln_gamma = h->blocks[l + 1].ln1_g;
          c.a16_out = h->y2;
          c.stats_out = h->ln_stats;
        }
        CKL(aihab::launch_gemm(h->m_hring, b.m_proj[0], &h->m_x, c, 256, half, h->side, true));
      }
      nsb = (D + 127) / 128;
      CK(cudaEventRecord(h->ev_join, h->side));
      CK(cudaStreamWaitEvent(s, h->ev_join, 0));
      continue;
    }
    if (!h->ln_fold) {
      {
        ProfScope ps(PC_LN, 6.0 * M * D, s);
        CKL(aihab::launch_layernorm(h->x, D, nullptr, 0, b.ln2_g, b.ln2_b, nullptr, h->y2, h->bf16, M, D, s));
      }
      if (run_gemm(h, h->m_y2, b.m_fc, M, 4 * D, D, aihab::EPI_BIAS_GELU_16, b.b_fc, h->big, nullptr, 4 * D, s, LnOpt(),
                   nullptr, next_dir()))
        return 1;
    } else {
      LnOpt o;
      o.s = b.s_fc;
      o.nsb = nsb;
      if (run_gemm(h, h->m_y2, b.m_fc, M, 4 * D, D, aihab::EPI_LN_BIAS_GELU_16, b.bp_fc, h->big, nullptr, 4 * D, s, o, nullptr,
                   next_dir()))
        return 1;
    }
    {
      LnOpt o;
      if (h->ln_fold && l + 1 < layers) o.gamma = h->blocks[l + 1].ln1_g;  // the last block feeds ln_post (kernel)
      if (run_gemm(h, h->m_h, b.m_proj, M, D, 4 * D, aihab::EPI_BIAS_RES_32, b.b_proj, nullptr, h->x, D, s, o, &nsb, next_dir()))
        return 1;
    }
  }
  return 0;
}

// transformer stack + ln_post on a chunk of n images whose patch rows are already in h->patches
int run_tower(aihab_vit* h, int n, void* feats_out, int out_dtype, cudaStream_t s) {
  const int D = h->D, L = h->L, M = n * L;
  // conv1 as GEMM + positional embedding (clip/model.py:217-221)
  if (run_gemm(h, h->m_patches, h->m_conv, n * h->g2, D, h->Kpad, aihab::EPI_PATCH_32, nullptr, nullptr, h->x, D, s))
    return 1;
  // class token row + ln_pre, in place on the fp32 residual stream (clip/model.py:220-222)
  const bool chain = h->ln_chain && !h->blocks.empty();
  {
    ProfScope ps(PC_LN, (chain ? 10.0 : 8.0) * M * D, s);
    if (chain)  // ... and ln_1 of the first block from the same registers
      CKL(aihab::launch_layernorm2(h->x, D, h->cls0, L, h->lnpre_g, h->lnpre_b, h->x, h->blocks[0].ln1_g, h->blocks[0].ln1_b,
                                   h->y2, h->bf16, M, D, s));
    else
      CKL(aihab::launch_layernorm(h->x, D, h->cls0, L, h->lnpre_g, h->lnpre_b, h->x, nullptr, 0, M, D, s));
  }
  if (run_blocks(h, n, s, chain)) return 1;
  // ln_post on token 0 of every image (clip/model.py:228); rows are L*D apart
  float* o32 = out_dtype == AIHAB_F32 ? static_cast<float*>(feats_out) : nullptr;
  void* o16 = out_dtype == AIHAB_F32 ? nullptr : feats_out;
  CKL(aihab::launch_layernorm(h->x, static_cast<size_t>(L) * D, nullptr, 0, h->lnpost_g, h->lnpost_b, o32, o16,
                              out_dtype == AIHAB_BF16, n, D, s));
  return 0;
}

// Common part of the two towers: packs the residual attention blocks, allocates the activation workspace for
// h->cap_rows token rows (+ patch_rows im2col rows for the image tower) and builds the TMA descriptors.
// Returns non-zero after fail(); the caller destroys the handle.
int build_stack(aihab_vit* h, int layers, const aihab_vit_block_weights* blocks, size_t patch_rows) {
  const int D = h->D, L = h->L;
  {
    const char* e = getenv("AIHAB_LNFOLD");
    h->ln_fold = !(e && e[0] == '0') && D / 128 <= kMaxStatBlocks;
    const char* z = getenv("AIHAB_ZIGZAG");
    h->zigzag = !(z && z[0] == '0');
  }
  h->blocks.resize(layers);
  for (int i = 0; i < layers; ++i) {
    const aihab_vit_block_weights& s = blocks[i];
    aihab_vit::Block& b = h->blocks[i];
    if (upload_16(h, s.in_proj_weight, 3 * D, D, D, &b.w_in) || upload_16(h, s.out_proj_weight, D, D, D, &b.w_out) ||
        upload_16(h, s.c_fc_weight, 4 * D, D, D, &b.w_fc) || upload_16(h, s.c_proj_weight, D, 4 * D, 4 * D, &b.w_proj))
      return 1;
    if (upload_f32(h, s.in_proj_bias, 3 * D, &b.b_in) || upload_f32(h, s.out_proj_bias, D, &b.b_out) ||
        upload_f32(h, s.c_fc_bias, 4 * D, &b.b_fc) || upload_f32(h, s.c_proj_bias, D, &b.b_proj) ||
        upload_f32(h, s.ln_1_weight, D, &b.ln1_g) || upload_f32(h, s.ln_1_bias, D, &b.ln1_b) ||
        upload_f32(h, s.ln_2_weight, D, &b.ln2_g) || upload_f32(h, s.ln_2_bias, D, &b.ln2_b))
      return 1;
    if (weight_maps(h, b.w_in, 3 * D, D, b.m_in) || weight_maps(h, b.w_out, D, D, b.m_out) ||
        weight_maps(h, b.w_fc, 4 * D, D, b.m_fc) || weight_maps(h, b.w_proj, D, 4 * D, b.m_proj))
      return 1;
    if (h->ln_fold &&
        (fold_ln(h, s.in_proj_weight, s.in_proj_bias, s.ln_1_weight, s.ln_1_bias, 3 * D, D, &b.s_in, &b.bp_in) ||
         fold_ln(h, s.c_fc_weight, s.c_fc_bias, s.ln_2_weight, s.ln_2_bias, 4 * D, D, &b.s_fc, &b.bp_fc)))
      return 1;
  }
  // workspace
  const size_t prow = patch_rows;
  if ((prow > 0 && dev_alloc(h, &h->patches, prow * h->Kpad * 2)) ||
      dev_alloc(h, reinterpret_cast<void**>(&h->x), h->cap_rows * D * 4) || dev_alloc(h, &h->y, h->cap_rows * D * 2) ||
      dev_alloc(h, &h->big, h->cap_rows * 4 * D * 2) || dev_alloc(h, &h->y2, h->cap_rows * D * 2) ||
      dev_alloc(h, reinterpret_cast<void**>(&h->ln_stats), h->cap_rows * kMaxStatBlocks * 2 * sizeof(float)))
    return 1;
  if ((prow > 0 && cudaMemset(h->patches, 0, prow * h->Kpad * 2) != cudaSuccess) || cudaMemset(h->y, 0, h->cap_rows * D * 2) != cudaSuccess ||
      cudaMemset(h->big, 0, h->cap_rows * 4 * D * 2) != cudaSuccess ||
      cudaMemset(h->y2, 0, h->cap_rows * D * 2) != cudaSuccess) {
    fail("aihab_vit_create: cudaMemset failed");
    return 1;
  }
  if ((prow > 0 && aihab::make_tmap_2d_16bit(&h->m_patches, h->patches, prow, h->Kpad, static_cast<uint64_t>(h->Kpad) * 2, 128,
                                             h->bf16) != cudaSuccess) ||
      aihab::make_tmap_2d_16bit(&h->m_y, h->y, h->cap_rows, D, static_cast<uint64_t>(D) * 2, 128, h->bf16) != cudaSuccess ||
      aihab::make_tmap_2d_16bit(&h->m_h, h->big, h->cap_rows, 4 * D, static_cast<uint64_t>(4 * D) * 2, 128, h->bf16) != cudaSuccess ||
      aihab::make_tmap_2d_16bit(&h->m_y2, h->y2, h->cap_rows, D, static_cast<uint64_t>(D) * 2, 128, h->bf16) != cudaSuccess ||
      aihab::make_tmap_2d_f32_box32(&h->m_x, h->x, h->cap_rows, D, static_cast<uint64_t>(D) * 4) != cudaSuccess) {
    fail("aihab_vit_create: cuTensorMapEncodeTiled failed for the workspace");
    return 1;
  }
  h->attn_kind = attention_kind(L);
  if (h->attn_kind == 3 && aihab::attention_tcf_tail_rows(L) > 0 && getenv("AIHAB_ATTN_SERIAL_TAIL") == nullptr) {
    if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) {
      fail("aihab_vit_create: could not create the attention side stream");
      return 1;
    }
  }
  {
    if (const char* ec = getenv("AIHAB_LN_CHAIN")) h->ln_chain = ec[0] != '0';
    const char* e = getenv("AIHAB_MLP_PIPE");
    const char* ef = getenv("AIHAB_MLP_FUSED");
    const bool fused = ef != nullptr && ef[0] == '1';
    if (fused) e = ef;
    int ring = 8192;  // rows: 32 pair-rows x 4D x 2 B = 50 MB at D = 768
    if (const char* er = getenv("AIHAB_MLP_RING_PAIRS")) ring = 256 * std::max(4, atoi(er));
    if (e != nullptr && e[0] == '1' && h->ln_fold && (D % 256) == 0 && h->cap_rows > static_cast<size_t>(2 * ring)) {
      if (dev_alloc(h, reinterpret_cast<void**>(&h->mlp_ctr), 2 * (h->cap_rows / 256 + 1) * sizeof(unsigned))) return 1;
      if (aihab::make_tmap_2d_16bit(&h->m_hring, h->big, ring, 4 * D, static_cast<uint64_t>(4 * D) * 2, 128, h->bf16) !=
          cudaSuccess) {
        fail("aihab_vit_create: cuTensorMapEncodeTiled failed for the MLP ring view");
        return 1;
      }
      if (h->side == nullptr && cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess) {
        fail("aihab_vit_create: could not create the side stream");
        return 1;
      }
      if (h->ev_fork == nullptr && (cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                                    cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess)) {
        fail("aihab_vit_create: could not create the fork / join events");
        return 1;
      }
      h->ring_rows = ring;
      h->mlp_pipe = !fused;
      h->mlp_fused = fused;
    }
  }
  if (h->attn_kind > 0) {
    const uint64_t pitch = static_cast<uint64_t>(3 * D) * 2;
    if (aihab::make_tmap_2d_16bit(&h->m_attn_q, h->big, h->cap_rows, 3 * D, pitch, 128, h->bf16) != cudaSuccess ||
        aihab::make_tmap_2d_16bit(&h->m_attn_kv, h->big, h->cap_rows, 3 * D, pitch,
                                  attention_key_box(h->attn_kind, L), h->bf16) != cudaSuccess) {
      fail("aihab_vit_create: cuTensorMapEncodeTiled failed for the attention views");
      return 1;
    }
    if (h->attn_kind == 2 && !h->causal && aihab::attention_tcd_supported(L)) {
      if (aihab::make_tmap_3d_16bit_seq(&h->m_attn_out3, h->y, h->cap_rows / L, L, D, 32, h->bf16) != cudaSuccess) {
        fail("aihab_vit_create: cuTensorMapEncodeTiled failed for the attention output view");
        return 1;
      }
      h->has_attn_out3 = true;
    }
  }
  return 0;
}

size_t dtype_size(int dt) { return dt == AIHAB_F32 ? 4 : 2; }

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int aihab_abi_version(void) { return AIHAB_ABI_VERSION; }
const char* aihab_last_error(void) { return g_err.c_str(); }
uint64_t aihab_kernel_launches(void) { return g_launches.load(); }

int aihab_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  return 0;
}

int aihab_profile_read(int cls, double* ms_out, uint64_t* launches_out, double* work_out, int reset) {
  if (cls < 0 || cls >= PC_COUNT) return fail("aihab_profile_read: bad class");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double ms = 0.0, work = 0.0;
  uint64_t cnt = 0;
  for (const ProfRec& r : g_prof_recs) {
    if (r.cls != cls) continue;
    CK(cudaEventSynchronize(r.b));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, r.a, r.b));
    ms += t;
    work += r.work;
    ++cnt;
  }
  if (reset) {
    std::vector<ProfRec> keep;
    for (const ProfRec& r : g_prof_recs) {
      if (r.cls == cls)
        g_prof_free.emplace_back(r.a, r.b);
      else
        keep.push_back(r);
    }
    g_prof_recs.swap(keep);
  }
  if (ms_out) *ms_out = ms;
  if (launches_out) *launches_out = cnt;
  if (work_out) *work_out = work;
  return 0;
}

int aihab_profile_sites(int cls, double* work_out, int* aux_out, double* ms_out, uint64_t* launches_out, int cap) {
  if (cls < 0 || cls >= PC_COUNT || cap < 0) {
    fail("aihab_profile_sites: bad argument");
    return -1;
  }
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int n = 0;
  for (const ProfRec& r : g_prof_recs) {
    if (r.cls != cls) continue;
    float t = 0.f;
    if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) {
      fail("aihab_profile_sites: event query failed");
      return -1;
    }
    int g = 0;
    while (g < n && (work_out[g] != r.work || aux_out[g] != r.aux)) ++g;
    if (g == n) {
      if (n == cap) continue;  // more distinct sites than the caller asked for: the rest are dropped
      work_out[n] = r.work;
      aux_out[n] = r.aux;
      ms_out[n] = 0.0;
      launches_out[n] = 0;
      ++n;
    }
    ms_out[g] += t;
    launches_out[g] += 1;
  }
  return n;
}

int aihab_vit_create(const aihab_vit_config* cfg, const aihab_vit_weights* w, int device, aihab_vit** out) {
  if (cfg == nullptr || w == nullptr || out == nullptr) return fail("aihab_vit_create: null argument");
  *out = nullptr;
  if (cfg->patch_size <= 0 || cfg->image_size % cfg->patch_size != 0)
    return fail("aihab_vit_create: image_size must be a multiple of patch_size");
  if (cfg->width % 128 != 0 || cfg->width > 2048) return fail("aihab_vit_create: width must be a multiple of 128, <= 2048");
  if (cfg->heads * 64 != cfg->width) return fail("aihab_vit_create: heads must equal width / 64");
  if (cfg->dtype != AIHAB_F16 && cfg->dtype != AIHAB_BF16) return fail("aihab_vit_create: dtype must be AIHAB_F16 or AIHAB_BF16");
  if (cfg->layers <= 0 || cfg->max_batch <= 0 || w->blocks == nullptr) return fail("aihab_vit_create: bad layers/max_batch/blocks");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail("aihab_vit_create: no CUDA device (this library has no CPU fallback)");
  }
  if (device < 0 || device >= ndev) return fail("aihab_vit_create: bad device index");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail("aihab_vit_create: device is not sm_100 (Blackwell B200) — kernels are sm_100a only");
  DeviceGuard guard(device);
  if (!guard.ok) return fail("aihab_vit_create: cudaSetDevice failed");
  CK(aihab::gemm_init());

  aihab_vit* h = new aihab_vit();
  h->cfg = *cfg;
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->g = cfg->image_size / cfg->patch_size;
  h->g2 = h->g * h->g;
  h->L = h->g2 + 1;
  h->D = cfg->width;
  h->Kp = 3 * cfg->patch_size * cfg->patch_size;
  h->Kpad = round_up(h->Kp, 64);
  h->bf16 = cfg->dtype == AIHAB_BF16;
  h->cap_rows = static_cast<size_t>(cfg->max_batch) * h->L;
  const int D = h->D, L = h->L;

  auto bail = [&](int) {
    aihab_vit_destroy(h);
    return 1;
  };
  if (aihab::attention_init(L) != cudaSuccess) {
    fail("aihab_vit_create: sequence too long for the attention kernel (L <= 908)");
    return bail(1);
  }
  if (upload_16(h, w->conv1_weight, D, h->Kp, h->Kpad, &h->w_conv)) return bail(1);
  if (weight_maps(h, h->w_conv, D, h->Kpad, h->m_conv)) return bail(1);
  if (upload_f32(h, w->positional_embedding, static_cast<size_t>(L) * D, &h->pos)) return bail(1);
  if (upload_f32(h, w->ln_pre_weight, D, &h->lnpre_g) || upload_f32(h, w->ln_pre_bias, D, &h->lnpre_b) ||
      upload_f32(h, w->ln_post_weight, D, &h->lnpost_g) || upload_f32(h, w->ln_post_bias, D, &h->lnpost_b))
    return bail(1);
  {  // cls0 = class_embedding + positional_embedding[0]  (fp32 add, clip/model.py:220-221)
    std::vector<float> cls(D), p0(D);
    if (cudaMemcpy(cls.data(), w->class_embedding, D * sizeof(float), cudaMemcpyDefault) != cudaSuccess ||
        cudaMemcpy(p0.data(), w->positional_embedding, D * sizeof(float), cudaMemcpyDefault) != cudaSuccess) {
      fail("aihab_vit_create: cannot read class/positional embedding");
      return bail(1);
    }
    for (int i = 0; i < D; ++i) cls[i] += p0[i];
    if (upload_f32(h, cls.data(), D, &h->cls0)) return bail(1);
  }
  if (build_stack(h, cfg->layers, w->blocks, static_cast<size_t>(cfg->max_batch) * h->g2)) return bail(1);
  if (cudaError_t e = cudaDeviceSynchronize(); e != cudaSuccess) {
    fail(std::string("aihab_vit_create: ") + cudaGetErrorString(e));
    return bail(1);
  }
  *out = h;
  return 0;
}

void aihab_vit_destroy(aihab_vit* h) {
  if (h == nullptr) return;
  DeviceGuard guard(h->device);
  for (void* p : h->allocs) cudaFree(p);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  delete h;
}

size_t aihab_vit_workspace_bytes(const aihab_vit* h) { return h ? h->ws_bytes : 0; }

// ---- text tower: the same handle type with token-embedding ends instead of the patch-embedding ends
int aihab_text_create(const aihab_text_config* cfg, const aihab_text_weights* w, int device, aihab_text** out) {
  if (cfg == nullptr || w == nullptr || out == nullptr) return fail("aihab_text_create: null argument");
  *out = nullptr;
  if (cfg->context_length <= 64 || cfg->context_length > 224)
    return fail("aihab_text_create: context_length must be in 65..224 (causal tcgen05 attention)");
  if (cfg->width % 128 != 0 || cfg->width > 2048) return fail("aihab_text_create: width must be a multiple of 128, <= 2048");
  if (cfg->heads * 64 != cfg->width) return fail("aihab_text_create: heads must equal width / 64");
  if (cfg->dtype != AIHAB_F16 && cfg->dtype != AIHAB_BF16) return fail("aihab_text_create: dtype must be AIHAB_F16 or AIHAB_BF16");
  if (cfg->layers <= 0 || cfg->max_batch <= 0 || cfg->vocab_size <= 0 || w->blocks == nullptr)
    return fail("aihab_text_create: bad layers/max_batch/vocab_size/blocks");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail("aihab_text_create: no CUDA device (this library has no CPU fallback)");
  }
  if (device < 0 || device >= ndev) return fail("aihab_text_create: bad device index");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail("aihab_text_create: device is not sm_100 (Blackwell B200) — kernels are sm_100a only");
  DeviceGuard guard(device);
  if (!guard.ok) return fail("aihab_text_create: cudaSetDevice failed");
  CK(aihab::gemm_init());

  aihab_vit* h = new aihab_vit();
  h->cfg.width = cfg->width;
  h->cfg.layers = cfg->layers;
  h->cfg.heads = cfg->heads;
  h->cfg.dtype = cfg->dtype;
  h->cfg.max_batch = cfg->max_batch;
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->L = cfg->context_length;
  h->D = cfg->width;
  h->bf16 = cfg->dtype == AIHAB_BF16;
  h->cap_rows = static_cast<size_t>(cfg->max_batch) * h->L;
  h->causal = 1;
  h->vocab = cfg->vocab_size;
  const int D = h->D, L = h->L;
  auto bail = [&](int) {
    aihab_vit_destroy(h);
    return 1;
  };
  if (upload_f32(h, w->token_embedding, static_cast<size_t>(cfg->vocab_size) * D, &h->tok_emb) ||
      upload_f32(h, w->positional_embedding, static_cast<size_t>(L) * D, &h->pos) ||
      upload_f32(h, w->ln_final_weight, D, &h->lnf_g) || upload_f32(h, w->ln_final_bias, D, &h->lnf_b) ||
      dev_alloc(h, reinterpret_cast<void**>(&h->xe), static_cast<size_t>(cfg->max_batch) * D * sizeof(float)))
    return bail(1);
  if (build_stack(h, cfg->layers, w->blocks, 0)) return bail(1);
  if (h->attn_kind != 2) {
    fail("aihab_text_create: the causal mask needs the persistent tcgen05 attention (AIHAB_ATTN must not cap it)");
    return bail(1);
  }
  if (cudaError_t e = cudaDeviceSynchronize(); e != cudaSuccess) {
    fail(std::string("aihab_text_create: ") + cudaGetErrorString(e));
    return bail(1);
  }
  *out = reinterpret_cast<aihab_text*>(h);
  return 0;
}

void aihab_text_destroy(aihab_text* h) { aihab_vit_destroy(reinterpret_cast<aihab_vit*>(h)); }

int aihab_text_encode(aihab_text* ht, const int64_t* tokens, int n, void* feats_out, int out_dtype, void* stream) {
  aihab_vit* h = reinterpret_cast<aihab_vit*>(ht);
  if (h == nullptr || !h->causal) return fail("aihab_text_encode: not a text handle");
  if (n < 0 || out_dtype < 0 || out_dtype > 2) return fail("aihab_text_encode: bad argument");
  if (n == 0) return 0;
  if (tokens == nullptr || feats_out == nullptr) return fail("aihab_text_encode: null buffer");
  DeviceGuard guard(h->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int L = h->L, D = h->D;
  for (int i0 = 0; i0 < n; i0 += h->cfg.max_batch) {
    const int nb = std::min(h->cfg.max_batch, n - i0);
    const int64_t* tk = tokens + static_cast<size_t>(i0) * L;
    {  // x = token_embedding(text) + positional_embedding   (clip/model.py:341-342)
      ProfScope ps(PC_PRE, static_cast<double>(nb) * L * D * 8.0, s);
      CKL(aihab::launch_embed_tokens(tk, h->tok_emb, h->pos, h->x, static_cast<long>(nb) * L, L, D, h->vocab, s));
    }
    if (run_blocks(h, nb, s)) return 1;
    // ln_final(x)[arange, text.argmax(-1)]   (clip/model.py:345, 350): gather the EOT rows, then LayerNorm on those
    CKL(aihab::launch_eot_gather(tk, h->x, h->xe, nb, L, D, s));
    void* dst = static_cast<uint8_t*>(feats_out) + static_cast<size_t>(i0) * D * dtype_size(out_dtype);
    float* o32 = out_dtype == AIHAB_F32 ? static_cast<float*>(dst) : nullptr;
    void* o16 = out_dtype == AIHAB_F32 ? nullptr : dst;
    CKL(aihab::launch_layernorm(h->xe, D, nullptr, 0, h->lnf_g, h->lnf_b, o32, o16, out_dtype == AIHAB_BF16, nb, D, s));
  }
  return 0;
}

int aihab_preferred_batch(int tokens, int width, int max_batch, int device) {
  if (tokens <= 0 || width <= 0 || max_batch <= 0) return max_batch > 0 ? max_batch : 1;
  const int sms = sm_count(device);
  const long shapes[4][2] = {{3L * width, width}, {width, width}, {4L * width, width}, {width, 4L * width}};
  int best = max_batch;
  double best_eff = -1.0;
  for (int b = std::max(1, max_batch / 2); b <= max_batch; ++b) {
    const long M = static_cast<long>(b) * tokens;
    double flops = 0.0, time = 0.0;  // time in per-SM tile MACs summed over whole waves
    for (const auto& nk : shapes) {
      const int N = static_cast<int>(nk[0]);
      const long K = nk[1];
      const int bn = aihab::gemm_block_n(static_cast<int>(M), N, sms);
      const bool pair = aihab::gemm_use_pair(static_cast<int>(M), N, sms);
      const long m_blocks = (M + 127) / 128, n_blocks = (N + bn - 1) / bn;
      const long tiles = (pair ? (m_blocks + 1) / 2 : m_blocks) * n_blocks;
      const long units = pair ? sms / 2 : sms;
      const long waves = (tiles + units - 1) / units;
      // measured: a CTA pair moves half the W bytes per SM, a 128-wide tile twice the A bytes per FLOP
      const double tile_penalty = pair ? 1.0 : (bn == 256 ? 1.05 : 1.15);
      time += static_cast<double>(waves) * 128.0 * bn * K * tile_penalty;
      flops += static_cast<double>(M) * N * K;
    }
    const double eff = flops / (time * sms);
    if (eff >= best_eff) {  // ties go to the larger batch
      best_eff = eff;
      best = b;
    }
  }
  return best;
}

int aihab_vit_encode(aihab_vit* h, const void* images, int in_dtype, int n, void* feats_out, int out_dtype,
                     void* stream) {
  if (h == nullptr) return fail("aihab_vit_encode: null handle");
  if (n < 0 || in_dtype < 0 || in_dtype > 2 || out_dtype < 0 || out_dtype > 2) return fail("aihab_vit_encode: bad argument");
  if (n == 0) return 0;
  if (images == nullptr || feats_out == nullptr) return fail("aihab_vit_encode: null buffer");
  DeviceGuard guard(h->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int R = h->cfg.image_size;
  const size_t img_bytes = static_cast<size_t>(3) * R * R * dtype_size(in_dtype);
  for (int i0 = 0; i0 < n; i0 += h->cfg.max_batch) {
    const int nb = std::min(h->cfg.max_batch, n - i0);
    const uint8_t* src = static_cast<const uint8_t*>(images) + img_bytes * i0;
    {
      ProfScope ps(PC_PRE, static_cast<double>(nb) * (img_bytes + static_cast<double>(h->g2) * h->Kpad * 2), s);
      CKL(aihab::launch_im2col(src, in_dtype, nb, R, h->cfg.patch_size, h->Kpad, h->patches, h->bf16, s));
    }
    void* dst = static_cast<uint8_t*>(feats_out) + static_cast<size_t>(i0) * h->D * dtype_size(out_dtype);
    if (run_tower(h, nb, dst, out_dtype, s)) return 1;
  }
  return 0;
}

int aihab_vit_encode_u8(aihab_vit* h, const uint8_t* images_u8, int n, int sh, int sw, void* feats_out,
                        int out_dtype, void* stream) {
  if (h == nullptr) return fail("aihab_vit_encode_u8: null handle");
  if (n < 0 || sh <= 0 || sw <= 0 || out_dtype < 0 || out_dtype > 2) return fail("aihab_vit_encode_u8: bad argument");
  if (n == 0) return 0;
  if (images_u8 == nullptr || feats_out == nullptr) return fail("aihab_vit_encode_u8: null buffer");
  DeviceGuard guard(h->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int R = h->cfg.image_size;
  aihab::ResampleTables t;
  if (get_tables(h->device, sh, sw, R, &t)) return 1;
  const size_t img_bytes = static_cast<size_t>(sh) * sw * 3;
  for (int i0 = 0; i0 < n; i0 += h->cfg.max_batch) {
    const int nb = std::min(h->cfg.max_batch, n - i0);
    {
      ProfScope ps(PC_PRE, static_cast<double>(nb) * (img_bytes + static_cast<double>(h->g2) * h->Kpad * 2), s);
      CK(aihab::launch_preprocess(images_u8 + img_bytes * i0, nb, sh, sw, R, t, h->patches,
                                  h->bf16 ? AIHAB_BF16 : AIHAB_F16, 1, h->cfg.patch_size, h->Kpad, s));
      g_launches += (h->Kpad > h->Kp) ? 2 : 1;
    }
    void* dst = static_cast<uint8_t*>(feats_out) + static_cast<size_t>(i0) * h->D * dtype_size(out_dtype);
    if (run_tower(h, nb, dst, out_dtype, s)) return 1;
  }
  return 0;
}

int aihab_preprocess_u8(const uint8_t* images_u8, int n, int sh, int sw, int R, void* out, int out_dtype,
                        void* stream) {
  if (n < 0 || sh <= 0 || sw <= 0 || R <= 0 || out_dtype < 0 || out_dtype > 2) return fail("aihab_preprocess_u8: bad argument");
  if (n == 0) return 0;
  if (images_u8 == nullptr || out == nullptr) return fail("aihab_preprocess_u8: null buffer");
  const int dev = device_of(images_u8);
  DeviceGuard guard(dev);
  aihab::ResampleTables t;
  if (get_tables(dev, sh, sw, R, &t)) return 1;
  CKL(aihab::launch_preprocess(images_u8, n, sh, sw, R, t, out, out_dtype, 0, 0, 0, static_cast<cudaStream_t>(stream)));
  return 0;
}

int aihab_score(const float* feats, int n, int D, const float* proj, int E, const float* text_w, int C, float scale,
                int k, float* emb_out, float* logits_out, int64_t* topk_idx, float* topk_val, void* stream) {
  if (n < 0 || D <= 0) return fail("aihab_score: bad argument");
  if (n == 0) return 0;
  if (feats == nullptr) return fail("aihab_score: null feats");
  if (proj == nullptr) E = D;
  if (E <= 0) return fail("aihab_score: bad embed dim");
  if (text_w != nullptr && C <= 0) return fail("aihab_score: bad class count");
  if (k < 0 || k > 16 || (k > 0 && (text_w == nullptr || k > C || topk_idx == nullptr))) return fail("aihab_score: bad k (0..16, <= C)");
  const int dev = device_of(feats);
  DeviceGuard guard(dev);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(feats) | reinterpret_cast<uintptr_t>(proj) | reinterpret_cast<uintptr_t>(text_w) |
                           reinterpret_cast<uintptr_t>(emb_out) | reinterpret_cast<uintptr_t>(logits_out)) & 15) == 0;
  if (text_w != nullptr && aligned16 && aihab::score_mid_supported(n, D, E, C)) {
    // one extraction batch (the headline step's 256 rows): column-sliced projection and logits, then the top-k kernel
    keep_pool_warm(dev);
    AsyncTemp<float> raw_t(s), logit_t(s);
    if (proj != nullptr) CK(raw_t.alloc(static_cast<size_t>(n) * E * 4));
    if (logits_out == nullptr) CK(logit_t.alloc(static_cast<size_t>(n) * C * 4));
    float* lg = logits_out ? logits_out : logit_t.p;
    ProfScope ps(PC_SCORE, 2.0 * n * (static_cast<double>(D) * E + static_cast<double>(E) * C), s);
    CKL(aihab::launch_score_mid(feats, n, D, proj, E, text_w, C, scale, raw_t.p, emb_out, lg, s));
    if (proj != nullptr) g_launches += 1;  // two kernels
    if (k > 0) CKL(aihab::launch_topk(lg, n, C, k, topk_idx, topk_val, s));
    return 0;
  }
  if (aihab::score_fused_supported(n, D, E, text_w ? C : 0)) {
    ProfScope ps(PC_SCORE, 2.0 * n * (static_cast<double>(proj ? D : 0) * E + static_cast<double>(text_w ? E : 0) * C), s);
    CKL(aihab::launch_score_fused(feats, n, D, proj, E, text_w, C, scale, k, emb_out, logits_out, topk_idx, topk_val, s));
    return 0;
  }
  // rows per pass bounded so temporaries stay small (config 5: 1M rows x 1000 classes)
  keep_pool_warm(dev);
  const int chunk = 65536;
  AsyncTemp<float> emb_t(s), logit_t(s);
  const int rows_tmp = std::min(n, chunk);
  if (emb_out == nullptr) CK(emb_t.alloc(static_cast<size_t>(rows_tmp) * E * 4));
  if (text_w != nullptr && logits_out == nullptr) CK(logit_t.alloc(static_cast<size_t>(rows_tmp) * C * 4));
  float *emb_tmp = emb_t.p, *logit_tmp = logit_t.p;
  ProfScope ps(PC_SCORE, 2.0 * n * (static_cast<double>(proj ? D : 0) * E + static_cast<double>(text_w ? E : 0) * C), s);
  for (int i0 = 0; i0 < n; i0 += chunk) {
    const int nb = std::min(chunk, n - i0);
    const float* f = feats + static_cast<size_t>(i0) * D;
    float* emb = emb_out ? emb_out + static_cast<size_t>(i0) * E : emb_tmp;
    if (proj != nullptr) {
      CKL(aihab::launch_sgemm(f, proj, emb, nb, E, D, 1.0f, s));
      CKL(aihab::launch_l2norm(emb, emb, nb, E, 1e-12f, s));
    } else {
      CKL(aihab::launch_l2norm(f, emb, nb, E, 1e-12f, s));
    }
    if (text_w != nullptr) {
      float* lg = logits_out ? logits_out + static_cast<size_t>(i0) * C : logit_tmp;
      CKL(aihab::launch_sgemm(emb, text_w, lg, nb, C, E, scale, s));
      if (k > 0)
        CKL(aihab::launch_topk(lg, nb, C, k, topk_idx + static_cast<size_t>(i0) * k,
                               topk_val ? topk_val + static_cast<size_t>(i0) * k : nullptr, s));
    }
  }
  return 0;
}

int aihab_l2_normalize(const void* x, int in_dtype, int rows, int cols, float eps, void* y, int out_dtype, void* stream) {
  if (rows < 0 || cols <= 0 || in_dtype < 0 || in_dtype > 2 || out_dtype < 0 || out_dtype > 2 || !(eps >= 0.f))
    return fail("aihab_l2_normalize: bad argument");
  if (rows == 0) return 0;
  if (x == nullptr || y == nullptr) return fail("aihab_l2_normalize: null buffer");
  DeviceGuard guard(device_of(x));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const double bytes = static_cast<double>(rows) * cols * ((in_dtype == 0 ? 4 : 2) + (out_dtype == 0 ? 4 : 2));
  ProfScope ps(PC_SCORE, bytes, s);
  CKL(aihab::launch_l2norm_rows(x, in_dtype, y, out_dtype, rows, cols, eps, s));
  return 0;
}

int aihab_score16(const void* feats16, int n, int D, int dtype, const void* proj16, int E, const float* text_w, int C,
                  float scale, int k, float* emb_out, float* logits_out, int64_t* topk_idx, float* topk_val,
                  void* stream) {
  if (n < 0 || D <= 0 || E <= 0 || C <= 0) return fail("aihab_score16: bad argument");
  if (n == 0) return 0;
  if (feats16 == nullptr || proj16 == nullptr || text_w == nullptr) return fail("aihab_score16: null buffer");
  if (dtype != AIHAB_F16 && dtype != AIHAB_BF16) return fail("aihab_score16: dtype must be AIHAB_F16 or AIHAB_BF16");
  if ((D & 7) || (E & 7) || (C & 3)) return fail("aihab_score16: needs D % 8 == 0, E % 8 == 0, C % 4 == 0");
  if (k < 0 || k > 16 || k > C || (k > 0 && topk_idx == nullptr)) return fail("aihab_score16: bad k (0..16, <= C)");
  const int dev = device_of(feats16);
  DeviceGuard guard(dev);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(aihab::gemm_init());
  keep_pool_warm(dev);
  const int bf16 = dtype == AIHAB_BF16;
  const int sms = sm_count(dev);
  // rows per pass (AIHAB_SCORE16_CHUNK overrides).  Measured on B200 for 1 M x 768 -> 512 -> 1000 classes, top-5:
  // 32768 rows 9.5 ms, 16384 10.3, 8192 11.4, 4096 15.5 - L2-sized passes do NOT pay: the five launches per pass are
  // bound by their own latency / wave quantisation, not by the HBM round trip of the intermediates.
  static const int chunk_env = [] {
    const char* e = getenv("AIHAB_SCORE16_CHUNK");
    return e ? atoi(e) : 0;
  }();
  const int chunk = chunk_env > 0 ? chunk_env : 32768;
  const int rows_tmp = std::min(n, chunk);
  AsyncTemp<uint8_t> projT_t(s), w3_t(s), a3_t(s);
  AsyncTemp<float> emb_raw_t(s), logit_t(s);
  CK(projT_t.alloc(static_cast<size_t>(E) * D * 2));
  CK(w3_t.alloc(static_cast<size_t>(C) * 3 * E * 2));
  CK(a3_t.alloc(static_cast<size_t>(rows_tmp) * 3 * E * 2));
  // (allocated below only if the fp32 embeddings are materialised)
  // k <= 8 and no logits requested: the logits GEMM keeps per-row top-k candidates in its epilogue (EPI_TOPK_32) and the
  // [rows, C] logits never reach HBM; AIHAB_SCORE16_FUSED=0 restores the store + top-k kernel pair (A/B, tests)
  static const bool fused_env = [] {
    const char* e = getenv("AIHAB_SCORE16_FUSED");
    return !(e != nullptr && e[0] == '0');
  }();
  const bool fused = fused_env && logits_out == nullptr && k > 0 && k <= aihab::TOPK_SLOTS;
  const int bn_logits = aihab::gemm_block_n(rows_tmp, C, sms);
  const int cand_slots = 2 * ((C + bn_logits - 1) / bn_logits);
  AsyncTemp<float> cand_val_t(s);
  AsyncTemp<int> cand_idx_t(s);
  if (fused) {
    CK(cand_val_t.alloc(static_cast<size_t>(rows_tmp) * cand_slots * aihab::TOPK_SLOTS * 4));
    CK(cand_idx_t.alloc(static_cast<size_t>(rows_tmp) * cand_slots * aihab::TOPK_SLOTS * 4));
  } else if (logits_out == nullptr) {
    CK(logit_t.alloc(static_cast<size_t>(rows_tmp) * C * 4));
  }
  // no embeddings requested: GEMM 1 writes the hi | hi | lo split of the RAW embedding rows and their chunk sums of
  // squares itself (EPI_SPLIT3_16) and the logits GEMM normalises its accumulator rows - the fp32 embeddings and the
  // normalise + split pass over them disappear (AIHAB_SCORE16_SPLIT=0: the three-kernel sequence)
  static const bool split_env = [] {
    const char* e = getenv("AIHAB_SCORE16_SPLIT");
    return !(e != nullptr && e[0] == '0');
  }();
  const bool split = split_env && emb_out == nullptr && (E % 64) == 0;
  AsyncTemp<float> ss_t(s);
  if (split) CK(ss_t.alloc(static_cast<size_t>(rows_tmp) * (E / 64) * 4));
  else CK(emb_raw_t.alloc(static_cast<size_t>(rows_tmp) * E * 4));
  void *projT = projT_t.p, *w3 = w3_t.p, *a3 = a3_t.p;
  float *emb_raw = emb_raw_t.p, *logit_tmp = logit_t.p;
  ProfScope ps(PC_SCORE, 2.0 * n * (static_cast<double>(D) * E + static_cast<double>(E) * C), s);
  CKL(aihab::launch_transpose16(proj16, projT, D, E, s));   // [D, E] -> [E, D]: K-major operand B
  CKL(aihab::launch_split_textw(text_w, E, C, w3, s));      // [E, C] fp32 -> [C, 3E] fp16 (hi | lo | hi)
  for (int i0 = 0; i0 < n; i0 += chunk) {
    const int nb = std::min(chunk, n - i0);
    const uint8_t* f = static_cast<const uint8_t*>(feats16) + static_cast<size_t>(i0) * D * 2;
    aihab::GemmParams p{};
    CUtensorMap ma, mw;
    // emb_raw = feats16 @ proj16: products of 16-bit values are exact in fp32, so only the summation order differs
    // from the fp32 reference (methods/ProLIP.py:40)
    int bn = aihab::gemm_block_n(nb, E, sms);
    CK(aihab::make_tmap_2d_16bit(&ma, f, nb, D, static_cast<uint64_t>(D) * 2, 128, bf16));
    CK(aihab::make_tmap_2d_16bit(&mw, projT, E, D, static_cast<uint64_t>(D) * 2, bn, bf16));
    p.M = nb;
    p.N = E;
    p.K = D;
    p.ab_format = bf16;
    p.scale = 1.0f;
    p.row_ss = nullptr;
    if (split) {
      p.epilogue = aihab::EPI_SPLIT3_16;
      p.out16 = a3;
      p.out32 = nullptr;
      p.ldo = 3 * E;
      p.stats_out = ss_t.p;
      CKL(aihab::launch_gemm(ma, mw, nullptr, p, bn, sms, s));
      p.out16 = nullptr;
      p.stats_out = nullptr;
      p.row_ss = ss_t.p;
      p.row_ss_n = E / 64;
    } else {
      p.epilogue = aihab::EPI_SCALE_32;
      p.out32 = emb_raw;
      p.ldo = E;
      CKL(aihab::launch_gemm(ma, mw, nullptr, p, bn, sms, s));
      CKL(aihab::launch_l2norm_split(emb_raw, emb_out ? emb_out + static_cast<size_t>(i0) * E : nullptr, a3, nb, E, s));
    }
    // logits = scale * (e_hi w_hi + e_hi w_lo + e_lo w_hi): one K = 3E fp16 GEMM (methods/utils.py:185)
    float* lg = logits_out ? logits_out + static_cast<size_t>(i0) * C : logit_tmp;
    bn = fused ? bn_logits : aihab::gemm_block_n(nb, C, sms);
    CK(aihab::make_tmap_2d_16bit(&ma, a3, nb, 3 * E, static_cast<uint64_t>(3 * E) * 2, 128, 0));
    CK(aihab::make_tmap_2d_16bit(&mw, w3, C, 3 * E, static_cast<uint64_t>(3 * E) * 2, bn, 0));
    p.N = C;
    p.K = 3 * E;
    p.ab_format = 0;
    p.out32 = fused ? nullptr : lg;
    p.ldo = C;
    p.scale = scale;
    p.epilogue = aihab::EPI_SCALE_32;
    if (fused) {
      p.epilogue = aihab::EPI_TOPK_32;
      p.topk_k = k;
      p.cand_val = cand_val_t.p;
      p.cand_idx = cand_idx_t.p;
      CKL(aihab::launch_gemm(ma, mw, nullptr, p, bn, sms, s));
      CKL(aihab::launch_topk_merge(cand_val_t.p, cand_idx_t.p, nb, cand_slots, k, topk_idx + static_cast<size_t>(i0) * k,
                                   topk_val ? topk_val + static_cast<size_t>(i0) * k : nullptr, s));
      p.epilogue = aihab::EPI_SCALE_32;  // GEMM 1 of the next pass
      continue;
    }
    CKL(aihab::launch_gemm(ma, mw, nullptr, p, bn, sms, s));
    if (k > 0)
      CKL(aihab::launch_topk(lg, nb, C, k, topk_idx + static_cast<size_t>(i0) * k,
                             topk_val ? topk_val + static_cast<size_t>(i0) * k : nullptr, s));
  }
  return 0;
}

int aihab_prototype_scores(const float* emb, const int64_t* labels, int n, int E, const float* prototypes_t,
                           const int64_t* owner, int P, float* sim_to_prototype, int64_t* prototype_id,
                           float* sim_to_other, float* margin, void* stream) {
  if (n < 0 || E <= 0 || P <= 0) return fail("aihab_prototype_scores: bad argument");
  if (n == 0) return 0;
  if (emb == nullptr || labels == nullptr || prototypes_t == nullptr || owner == nullptr || sim_to_prototype == nullptr)
    return fail("aihab_prototype_scores: null buffer");
  const int dev = device_of(emb);
  DeviceGuard guard(dev);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  keep_pool_warm(dev);
  const int chunk = 65536;  // rows per pass: the [rows, P] similarity tile stays small
  const int rows_tmp = std::min(n, chunk);
  AsyncTemp<float> sim_t(s);
  CK(sim_t.alloc(static_cast<size_t>(rows_tmp) * P * 4));
  float* sim = sim_t.p;
  ProfScope ps(PC_SCORE, 2.0 * n * static_cast<double>(E) * P, s);
  const bool tensor = (E % 8) == 0 && (P % 4) == 0;
  AsyncTemp<uint8_t> a3_t(s), w3_t(s);
  void *a3 = nullptr, *w3 = nullptr;
  const int sms = sm_count(dev);
  if (tensor) {
    // sim = e_hi p_hi + e_hi p_lo + e_lo p_hi as ONE K = 3E fp16 tcgen05 GEMM with fp32 accumulation (relative error
    // ~2^-21, the hi/lo split of aihab_score16) instead of an fp32 CUDA-core GEMM
    CK(aihab::gemm_init());
    CK(a3_t.alloc(static_cast<size_t>(rows_tmp) * 3 * E * 2));
    CK(w3_t.alloc(static_cast<size_t>(P) * 3 * E * 2));
    a3 = a3_t.p;
    w3 = w3_t.p;
    CKL(aihab::launch_split_textw(prototypes_t, E, P, w3, s));  // [E, P] fp32 -> [P, 3E] fp16 (hi | lo | hi)
  }
  for (int i0 = 0; i0 < n; i0 += chunk) {
    const int nb = std::min(chunk, n - i0);
    const float* e = emb + static_cast<size_t>(i0) * E;
    if (tensor) {
      CKL(aihab::launch_l2norm_split(e, nullptr, a3, nb, E, s, /*normalize=*/0));
      aihab::GemmParams p{};
      CUtensorMap ma, mw;
      const int bn = aihab::gemm_block_n(nb, P, sms);
      CK(aihab::make_tmap_2d_16bit(&ma, a3, nb, 3 * E, static_cast<uint64_t>(3 * E) * 2, 128, 0));
      CK(aihab::make_tmap_2d_16bit(&mw, w3, P, 3 * E, static_cast<uint64_t>(3 * E) * 2, bn, 0));
      p.M = nb;
      p.N = P;
      p.K = 3 * E;
      p.ab_format = 0;
      p.epilogue = aihab::EPI_SCALE_32;
      p.out32 = sim;
      p.ldo = P;
      p.scale = 1.0f;
      CKL(aihab::launch_gemm(ma, mw, nullptr, p, bn, sms, s));
    } else {
      CKL(aihab::launch_sgemm(e, prototypes_t, sim, nb, P, E, 1.0f, s));
    }
    CKL(aihab::launch_prototype_reduce(sim, labels + i0, owner, nb, P, P, sim_to_prototype + i0,
                                       prototype_id ? prototype_id + i0 : nullptr, sim_to_other ? sim_to_other + i0 : nullptr,
                                       margin ? margin + i0 : nullptr, s));
  }
  return 0;
}

int aihab_l2_metrics(const float* logits_l3, int n, int C3, const int32_t* l3_to_l2, int C2, int reduce, int k,
                     float* logits_l2_out, int64_t* topk_idx, float* topk_val, int64_t* top3_idx, float* top3_prob,
                     void* stream) {
  if (logits_l3 == nullptr || l3_to_l2 == nullptr || n < 0) return fail("aihab_l2_metrics: bad argument");
  if (C3 <= 0 || C3 > 1024 || C2 <= 0 || C2 > 256) return fail("aihab_l2_metrics: needs 1 <= C3 <= 1024 and 1 <= C2 <= 256");
  if (reduce < 0 || reduce > 2) return fail("aihab_l2_metrics: reduce must be 0 (sum), 1 (mean) or 2 (logsumexp)");
  if (k < 0 || k > C2 || (k > 0 && topk_idx == nullptr)) return fail("aihab_l2_metrics: 0 <= k <= C2 and topk_idx for k > 0");
  if (n == 0) return 0;
  DeviceGuard guard(device_of(logits_l3));
  ProfScope ps(PC_SCORE, static_cast<double>(n) * C3 * 2.0, static_cast<cudaStream_t>(stream));
  CKL(aihab::launch_l2_metrics(logits_l3, n, C3, reinterpret_cast<const int*>(l3_to_l2), C2, reduce, k, logits_l2_out,
                               topk_idx, topk_val, top3_idx, top3_prob, static_cast<cudaStream_t>(stream)));
  return 0;
}

int aihab_gemm16(const void* A, const void* W, int M, int N, int K, int ab_dtype, int epilogue, const float* bias,
                 void* out16, float* out32, int ldo, const float* pos, int g2, float scale, void* stream) {
  if (A == nullptr || W == nullptr || M <= 0 || N <= 0 || K <= 0 || (K & 7)) return fail("aihab_gemm16: bad argument (K % 8 == 0)");
  if (ab_dtype != AIHAB_F16 && ab_dtype != AIHAB_BF16) return fail("aihab_gemm16: ab_dtype must be AIHAB_F16 or AIHAB_BF16");
  const int dev = device_of(A);
  DeviceGuard guard(dev);
  CK(aihab::gemm_init());
  const int bf16 = ab_dtype == AIHAB_BF16;
  const int sms = sm_count(dev);
  const int bn = aihab::gemm_block_n(M, N, sms);
  CUtensorMap ma, mw;
  const bool pair = gemm_pair_enabled(M, N, sms);
  CK(aihab::make_tmap_2d_16bit(&ma, A, M, K, static_cast<uint64_t>(K) * 2, 128, bf16));
  CK(aihab::make_tmap_2d_16bit(&mw, W, N, K, static_cast<uint64_t>(K) * 2, pair ? 128 : bn, bf16));
  aihab::GemmParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  p.ab_format = bf16;
  p.epilogue = epilogue;
  p.bias = bias;
  p.out16 = out16;
  p.out32 = out32;
  p.ldo = ldo;
  p.pos = pos;
  p.g2 = g2;
  p.scale = scale;
  if (const char* dbg = getenv("AIHAB_GEMM_DEBUG")) p.debug = atoi(dbg);  // 77: no 16-bit epilogue (main-loop ceiling)
  CUtensorMap mc;
  const bool res = epilogue == AIHAB_EPI_BIAS_RES_32;
  if (res) {
    if (out32 == nullptr || (N & 31)) return fail("aihab_gemm16: EPI_BIAS_RES_32 needs out32 and N % 32 == 0");
    CK(aihab::make_tmap_2d_f32_box32(&mc, out32, M, N, static_cast<uint64_t>(ldo) * 4));
  }
  CKL(aihab::launch_gemm(ma, mw, res ? &mc : nullptr, p, bn, sms, static_cast<cudaStream_t>(stream), pair));
  return 0;
}

int aihab_layernorm(const float* x, int rows, int D, const float* gamma, const float* beta, float* out32,
                    void* out16, int out16_dtype, void* stream) {
  if (x == nullptr || gamma == nullptr || beta == nullptr || rows < 0) return fail("aihab_layernorm: bad argument");
  DeviceGuard guard(device_of(x));
  CKL(aihab::launch_layernorm(x, D, nullptr, 0, gamma, beta, out32, out16, out16_dtype == AIHAB_BF16, rows, D,
                              static_cast<cudaStream_t>(stream)));
  return 0;
}

int aihab_attention_causal(const void* qkv, void* out, int n, int L, int H, int dtype, void* stream) {
  if (qkv == nullptr || out == nullptr || n < 0) return fail("aihab_attention_causal: bad argument");
  if (dtype != AIHAB_F16 && dtype != AIHAB_BF16) return fail("aihab_attention_causal: dtype must be AIHAB_F16 or AIHAB_BF16");
  if (!aihab::attention_tcp_supported(L) || aihab::attention_tcp_pack(L) != 1)
    return fail("aihab_attention_causal: needs 64 < L <= 224");
  DeviceGuard guard(device_of(qkv));
  if (n == 0) return 0;
  CUtensorMap mq, mkv;
  const int bf16 = dtype == AIHAB_BF16;
  const uint64_t rows = static_cast<uint64_t>(n) * L, pitch = static_cast<uint64_t>(3 * H * 64) * 2;
  CK(aihab::gemm_init());
  CK(aihab::make_tmap_2d_16bit(&mq, qkv, rows, 3 * H * 64, pitch, 128, bf16));
  CK(aihab::make_tmap_2d_16bit(&mkv, qkv, rows, 3 * H * 64, pitch, aihab::attention_tcp_key_rows(L), bf16));
  CKL(aihab::launch_attention_tcp(mq, mkv, out, n, L, H, bf16, sm_count(device_of(qkv)), static_cast<cudaStream_t>(stream),
                                  0, 1));
  return 0;
}

int aihab_attention(const void* qkv, void* out, int n, int L, int H, int dtype, void* stream) {
  if (qkv == nullptr || out == nullptr || n < 0) return fail("aihab_attention: bad argument");
  if (dtype != AIHAB_F16 && dtype != AIHAB_BF16) return fail("aihab_attention: dtype must be AIHAB_F16 or AIHAB_BF16");
  DeviceGuard guard(device_of(qkv));
  if (n == 0) return 0;
  const int kind = attention_kind(L);
  if (kind > 0) {
    CUtensorMap mq, mkv;
    const int bf16 = dtype == AIHAB_BF16;
    const uint64_t rows = static_cast<uint64_t>(n) * L, pitch = static_cast<uint64_t>(3 * H * 64) * 2;
    CK(aihab::gemm_init());
    CK(aihab::make_tmap_2d_16bit(&mq, qkv, rows, 3 * H * 64, pitch, 128, bf16));
    CK(aihab::make_tmap_2d_16bit(&mkv, qkv, rows, 3 * H * 64, pitch, attention_key_box(kind, L), bf16));
    if (kind == 3)
      CKL(aihab::launch_attention_tcf(mq, mkv, qkv, out, n, L, H, bf16, sm_count(device_of(qkv)), static_cast<cudaStream_t>(stream)));
    else {
      CUtensorMap mo;
      const bool dual = aihab::attention_tcd_supported(L);
      if (dual) CK(aihab::make_tmap_3d_16bit_seq(&mo, out, n, L, H * 64, 32, bf16));
      CKL(aihab::launch_attention_tcp(mq, mkv, out, n, L, H, bf16, sm_count(device_of(qkv)), static_cast<cudaStream_t>(stream), 0,
                                      0, dual ? &mo : nullptr));
    }
    return 0;
  }
  CKL(aihab::launch_attention(qkv, out, n, L, H, dtype == AIHAB_BF16, static_cast<cudaStream_t>(stream)));
  return 0;
}

}  // extern "C"
