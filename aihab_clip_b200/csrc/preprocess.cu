// uint8 HWC -> resize (short side -> R, bicubic) -> centre crop R x R -> /255 -> (x - mean) / std, on the GPU.
// Replaces data/clip_transforms.py:50-56 (v2.Resize(BICUBIC) + v2.CenterCrop + v2.ToTensor + v2.Normalize with
// CLIP_MEAN/CLIP_STD at :22-23; twin pipeline clip/clip.py:74-81).  The resize reproduces Pillow's
// ImagingResample for 8-bit images bit for bit: separable, horizontal pass then vertical pass, each pass in
// 22-bit fixed point with a uint8 round-and-clamp between the passes.  Coefficient tables come from the host
// (api.cu: build_resample_axis) and are indexed in resized-image coordinates.
#include "kernels.cuh"
#include "ptx.cuh"

namespace aihab {

namespace {

constexpr int TH = 16;  // output rows per CTA

__device__ __forceinline__ int clip8(int v) {
  v >>= 22;  // arithmetic shift, as Pillow's clip8 lookup index
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

template <typename OutT>
__device__ __forceinline__ OutT to_out(float v);
template <>
__device__ __forceinline__ float to_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half to_out<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename OutT>
__global__ void __launch_bounds__(256) preprocess_kernel(const uint8_t* __restrict__ in, int sh, int sw, int R,
                                                         ResampleTables t, OutT* __restrict__ out, int layout, int p,
                                                         int Kpad) {
  extern __shared__ uint8_t tile[];  // [nrows][R*3] horizontally resampled rows (HWC, crop columns only)
  const int img = blockIdx.y;
  const int y0 = blockIdx.x * TH;
  const int th = min(TH, R - y0);
  const int RC = R * 3;
  const uint8_t* src = in + static_cast<size_t>(img) * sh * sw * 3;

  int r_first, r_last;
  if (t.need_v) {
    const int ry0 = t.crop_top + y0, ry1 = t.crop_top + y0 + th - 1;
    r_first = t.v_bounds[2 * ry0];
    r_last = t.v_bounds[2 * ry1] + t.v_bounds[2 * ry1 + 1];
  } else {
    r_first = t.crop_top + y0;
    r_last = r_first + th;
  }
  const int nrows = r_last - r_first;

  // ---- pass 1: horizontal, input rows [r_first, r_last) -> uint8 tile
  for (int idx = threadIdx.x; idx < nrows * RC; idx += blockDim.x) {
    const int r = idx / RC;
    const int rem = idx - r * RC;
    const int ox = rem / 3, c = rem - ox * 3;
    const int rx = t.crop_left + ox;
    const uint8_t* row = src + static_cast<size_t>(r_first + r) * sw * 3 + c;
    int v;
    if (t.need_h) {
      const int xmin = t.h_bounds[2 * rx], cnt = t.h_bounds[2 * rx + 1];
      const int* k = t.h_coeffs + static_cast<size_t>(rx) * t.h_ksize;
      int ss = 1 << 21;
      for (int j = 0; j < cnt; ++j) ss += static_cast<int>(__ldg(row + (xmin + j) * 3)) * __ldg(k + j);
      v = clip8(ss);
    } else {
      v = __ldg(row + rx * 3);
    }
    tile[idx] = static_cast<uint8_t>(v);
  }
  __syncthreads();

  // ---- pass 2: vertical + normalise, coalesced along x for every (row, channel)
  const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
  const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
  const int g = (layout == 1) ? R / p : 0;
  for (int idx = threadIdx.x; idx < th * RC; idx += blockDim.x) {
    const int oy = idx / RC;
    const int rem = idx - oy * RC;
    const int c = rem / R, ox = rem - c * R;
    int v;
    if (t.need_v) {
      const int ry = t.crop_top + y0 + oy;
      const int ymin = t.v_bounds[2 * ry] - r_first, cnt = t.v_bounds[2 * ry + 1];
      const int* k = t.v_coeffs + static_cast<size_t>(ry) * t.v_ksize;
      int ss = 1 << 21;
      for (int j = 0; j < cnt; ++j) ss += static_cast<int>(tile[(ymin + j) * RC + ox * 3 + c]) * __ldg(k + j);
      v = clip8(ss);
    } else {
      v = tile[oy * RC + ox * 3 + c];
    }
    const float f = (static_cast<float>(v) / 255.0f - mean[c]) / stdv[c];
    const int y = y0 + oy;
    size_t o;
    if (layout == 0) {
      o = ((static_cast<size_t>(img) * 3 + c) * R + y) * R + ox;
    } else {
      const int gy = y / p, ky = y - gy * p, gx = ox / p, kx = ox - gx * p;
      o = (static_cast<size_t>(img) * g * g + gy * g + gx) * Kpad + c * p * p + ky * p + kx;
    }
    out[o] = to_out<OutT>(f);
  }
}

// Fast path when the input already has the crop size (new == in, Pillow skips both passes): one thread converts 16
// consecutive pixels of one image row (48 contiguous input bytes, 3 x 128-bit loads) into three 32-byte runs of the
// im2col row (one per channel).  Requires p % 16 == 0 (ViT-B/16, ViT-B/32) so a run never straddles a patch.
template <bool BF16>
__global__ void __launch_bounds__(256) normalize_im2col16_kernel(const uint8_t* __restrict__ in, int n, int sh, int sw,
                                                                 int R, int top, int left, uint16_t* __restrict__ out,
                                                                 int p, int Kpad) {
  const int segs = R >> 4;
  const long total = static_cast<long>(n) * R * segs;
  const int g = R / p;
  const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
  const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int xs = static_cast<int>(idx % segs);
    const long t = idx / segs;
    const int y = static_cast<int>(t % R);
    const int img = static_cast<int>(t / R);
    const uint8_t* src = in + ((static_cast<size_t>(img) * sh + top + y) * sw + left + xs * 16) * 3;
    uint32_t w[12];  // 48 bytes = 16 RGB pixels
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
        w[4 * i] = v.x;
        w[4 * i + 1] = v.y;
        w[4 * i + 2] = v.z;
        w[4 * i + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 12; ++i)
        w[i] = __ldg(src + 4 * i) | (__ldg(src + 4 * i + 1) << 8) | (__ldg(src + 4 * i + 2) << 16) |
               (static_cast<uint32_t>(__ldg(src + 4 * i + 3)) << 24);
    }
    auto px = [&](int i) { return static_cast<float>((w[i >> 2] >> ((i & 3) * 8)) & 0xffu); };
    const int x0 = xs * 16;
    const int gy = y / p, ky = y - gy * p, gx = x0 / p, kx = x0 - gx * p;
    uint16_t* dst = out + (static_cast<size_t>(img) * g * g + gy * g + gx) * Kpad + ky * p + kx;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = (px((2 * j) * 3 + c) / 255.0f - mean[c]) / stdv[c];
        const float b = (px((2 * j + 1) * 3 + c) / 255.0f - mean[c]) / stdv[c];
        pk[j] = ptx::pack2<BF16>(a, b);
      }
      uint4* d = reinterpret_cast<uint4*>(dst + c * p * p);
      d[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      d[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
  }
}

// zero the K padding columns of im2col rows (layout 1 writes only the first 3*p*p columns)
__global__ void zero_pad_kernel(uint16_t* out, long rows, int K, int Kpad) {
  const int w = Kpad - K;
  const long total = rows * w;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / w;
    out[r * Kpad + K + (i - r * w)] = 0;
  }
}

template <typename OutT>
cudaError_t launch_t(const uint8_t* in, int n, int sh, int sw, int R, const ResampleTables& t, void* out, int layout,
                     int p, int Kpad, int max_rows, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(max_rows) * R * 3;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(preprocess_kernel<OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  dim3 grid((R + TH - 1) / TH, n);
  preprocess_kernel<OutT><<<grid, 256, smem, stream>>>(in, sh, sw, R, t, static_cast<OutT*>(out), layout, p, Kpad);
  return cudaGetLastError();
}

}  // namespace

cudaError_t preprocess_init() { return cudaSuccess; }

cudaError_t launch_preprocess(const uint8_t* in, int n, int sh, int sw, int R, const ResampleTables& t, void* out,
                              int out_dtype, int layout, int p, int Kpad, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (layout == 1 && (out_dtype == 0 || p <= 0 || R % p != 0)) return cudaErrorInvalidValue;
  // rows of the (horizontally resampled) input one CTA needs: ceil(scale) * TH + ksize is a safe bound
  int max_rows = TH;
  if (t.need_v) {
    const int scale_up = (sh + t.new_h - 1) / t.new_h;  // ceil(in / out)
    max_rows = TH * scale_up + t.v_ksize + 2;
    if (max_rows > sh) max_rows = sh;
  }
  if (layout == 1 && !t.need_h && !t.need_v && (p % 16) == 0 && (R % 16) == 0 && Kpad == 3 * p * p &&
      out_dtype != 0) {
    const long total = static_cast<long>(n) * R * (R / 16);
    const int grid = static_cast<int>(std::min<long>((total + 255) / 256, 148L * 16));
    if (out_dtype == 2)
      normalize_im2col16_kernel<true><<<grid, 256, 0, stream>>>(in, n, sh, sw, R, t.crop_top, t.crop_left,
                                                                static_cast<uint16_t*>(out), p, Kpad);
    else
      normalize_im2col16_kernel<false><<<grid, 256, 0, stream>>>(in, n, sh, sw, R, t.crop_top, t.crop_left,
                                                                 static_cast<uint16_t*>(out), p, Kpad);
    return cudaGetLastError();
  }
  if (layout == 1 && Kpad > 3 * p * p) {
    const long rows = static_cast<long>(n) * (R / p) * (R / p);
    zero_pad_kernel<<<148, 256, 0, stream>>>(static_cast<uint16_t*>(out), rows, 3 * p * p, Kpad);
  }
  switch (out_dtype) {
    case 0: return launch_t<float>(in, n, sh, sw, R, t, out, layout, p, Kpad, max_rows, stream);
    case 1: return launch_t<__half>(in, n, sh, sw, R, t, out, layout, p, Kpad, max_rows, stream);
    case 2: return launch_t<__nv_bfloat16>(in, n, sh, sw, R, t, out, layout, p, Kpad, max_rows, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace aihab
