// On-device view generation (TPT/data/datautils.py:76-128, TPT/data/augmix_ops.py): the reference builds the 64 views
// of every test image on the host with PIL (RandomResizedCrop + flip, optionally AugMix chains) and ships 38.5 MB of
// fp32 views per image to the GPU.  Here the decoded uint8 image (a few hundred KB) is uploaded once and the views are
// produced next to the towers that consume them.  Everything below reproduces Pillow's arithmetic bit for bit:
//   resample_h/v   ImagingResample (libImaging/Resample.c): two-pass separable filter with 22-bit fixed-point
//                  coefficients (computed on the host in double, as Pillow does) and uint8 rounding after each pass
//   augmix_kernel  ImageOps.autocontrast / equalize / posterize / solarize (histogram + LUT, integer / double math as
//                  in the Python source) and Image.transform(AFFINE, BILINEAR) (libImaging/Geometry.c: double
//                  coordinates, bilinear in double, truncation to uint8), then ToTensor + Normalize and the AugMix
//                  blend in fp32 with the reference's operation order (no FMA contraction anywhere).
// Byte / integer work, HBM- and latency-bound: one CTA per (view, channel) keeps the 224x224 plane in shared memory.
#include "ptx.cuh"
#include "rlcf_internal.h"

namespace rlcf {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Resample.c: PRECISION_BITS

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;   // arithmetic shift = floor, as clip8_lookups[in >> PRECISION_BITS]
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// hdr[v] = {x0, y0, row_first, n_rows, flip, 0, 0, 0}: crop origin in the source, first source row (relative to y0) the
// vertical pass needs and how many.  hb [V][out][2] = (xmin, count), hk [V][out][ks] fixed-point taps.
__global__ void __launch_bounds__(256)
resample_h_kernel(const uint8_t* __restrict__ src, int W, const int* __restrict__ hdr, const int* __restrict__ hb,
                  const int* __restrict__ hk, int ks, int out_w, uint8_t* __restrict__ tmp, int tmp_rows) {
  const int v = blockIdx.y;
  const int* h = hdr + v * 8;
  const int n_rows = h[3];
  const int total = n_rows * out_w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / out_w, x = i % out_w;
    const int xmin = hb[(v * out_w + x) * 2], cnt = hb[(v * out_w + x) * 2 + 1];
    const int* k = hk + static_cast<size_t>(v * out_w + x) * ks;
    const uint8_t* line = src + (static_cast<size_t>(h[1] + h[2] + r) * W + h[0] + xmin) * 3;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int t = 0; t < cnt; ++t) {
      const int kk = k[t];
      s0 += line[t * 3 + 0] * kk;
      s1 += line[t * 3 + 1] * kk;
      s2 += line[t * 3 + 2] * kk;
    }
    uint8_t* o = tmp + ((static_cast<size_t>(v) * tmp_rows + r) * out_w + x) * 3;
    o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
  }
}

// vb [V][out][2] = (ymin relative to row_first, count), vk [V][out][ks].  out [V][out_h][out_w][3] uint8, mirrored in x
// when hdr.flip (RandomHorizontalFlip, datautils.py:91).
__global__ void __launch_bounds__(256)
resample_v_kernel(const uint8_t* __restrict__ tmp, int tmp_rows, const int* __restrict__ hdr,
                  const int* __restrict__ vb, const int* __restrict__ vk, int ks, int out_h, int out_w,
                  uint8_t* __restrict__ out) {
  const int v = blockIdx.y;
  const int flip = hdr[v * 8 + 4];
  const int total = out_h * out_w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / out_w, x = i % out_w;
    const int ymin = vb[(v * out_h + y) * 2], cnt = vb[(v * out_h + y) * 2 + 1];
    const int* k = vk + static_cast<size_t>(v * out_h + y) * ks;
    const uint8_t* col = tmp + ((static_cast<size_t>(v) * tmp_rows + ymin) * out_w + x) * 3;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int t = 0; t < cnt; ++t) {
      const int kk = k[t];
      const uint8_t* px = col + static_cast<size_t>(t) * out_w * 3;
      s0 += px[0] * kk; s1 += px[1] * kk; s2 += px[2] * kk;
    }
    const int xo = flip ? out_w - 1 - x : x;
    uint8_t* o = out + ((static_cast<size_t>(v) * out_h + y) * out_w + xo) * 3;
    o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
  }
}

// precompute_coeffs + normalize_coeffs_8bpc (libImaging/Resample.c) on the device, in double with the C source's
// operation order (explicit _rn intrinsics: no FMA contraction), so the host only draws the crop boxes.
// geom [V][8] int32 = {in_w, in_h, res_w, res_h, lo_x, lo_y, filter (0 bilinear, 1 bicubic), 0}: the source region of
// in_w x in_h pixels is resized to res_w x res_h and the window of `out` outputs starting at (lo_x, lo_y) is kept.
// grid (2 axes, V), one thread per output index.  The vertical block also fills hdr[v][2..3] = (first source row the
// window needs, number of rows) and re-bases its first-tap indices to that row (ImagingResample: ybox_first).
__device__ __forceinline__ double pil_filter(int filt, double x) {
  if (x < 0.0) x = -x;
  if (filt == 0) return x < 1.0 ? __dsub_rn(1.0, x) : 0.0;
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.5, x), 2.5), x), x), 1.0);
  if (x < 2.0) return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), -0.5);
  return 0.0;
}

__global__ void resample_taps_kernel(const int* __restrict__ geom, int out, int ks_h, int ks_v, int* __restrict__ hdr,
                                     int* __restrict__ hb, int* __restrict__ hk, int* __restrict__ vb,
                                     int* __restrict__ vk) {
  __shared__ int s_first, s_last;
  const int axis = blockIdx.x, v = blockIdx.y, t = threadIdx.x;
  const int* g = geom + v * 8;
  const int in_size = g[axis], res = g[2 + axis], lo = g[4 + axis], filt = g[6];
  const int ks = axis == 0 ? ks_h : ks_v;
  int* bounds = (axis == 0 ? hb : vb) + (static_cast<size_t>(v) * out) * 2;
  int* taps = (axis == 0 ? hk : vk) + (static_cast<size_t>(v) * out) * ks;
  int xmin = 0, cnt = 0;
  if (t < out) {
    const double scale = __ddiv_rn(static_cast<double>(static_cast<float>(in_size)), static_cast<double>(res));
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = __dmul_rn(filt == 0 ? 1.0 : 2.0, filterscale);
    const double center = __dadd_rn(0.0, __dmul_rn(__dadd_rn(static_cast<double>(lo + t), 0.5), scale));
    const double ss = __ddiv_rn(1.0, filterscale);
    xmin = static_cast<int>(__dadd_rn(__dsub_rn(center, support), 0.5));
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
    if (xmax > in_size) xmax = in_size;
    cnt = xmax - xmin;
    double ww = 0.0;
    for (int x = 0; x < cnt; ++x)
      ww = __dadd_rn(ww, pil_filter(filt, __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss)));
    int* k = taps + static_cast<size_t>(t) * ks;
    for (int x = 0; x < ks; ++x) {
      int fixed = 0;
      if (x < cnt) {
        double w = pil_filter(filt, __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss));
        if (ww != 0.0) w = __ddiv_rn(w, ww);
        fixed = w < 0 ? static_cast<int>(__dadd_rn(-0.5, __dmul_rn(w, 4194304.0)))
                      : static_cast<int>(__dadd_rn(0.5, __dmul_rn(w, 4194304.0)));
      }
      k[x] = fixed;
    }
  }
  if (axis == 1) {
    if (t == 0) s_first = xmin;
    if (t == out - 1) s_last = xmin + cnt;
    __syncthreads();
    if (t < out) xmin -= s_first;
    if (t == 0) { hdr[v * 8 + 2] = s_first; hdr[v * 8 + 3] = s_last - s_first; }
  }
  if (t < out) { bounds[t * 2] = xmin; bounds[t * 2 + 1] = cnt; }
}

int resample_taps(const int* geom, int n_views, int out, int ks_h, int ks_v, int* hdr, int* hb, int* hk, int* vb,
                  int* vk, cudaStream_t stream) {
  if (n_views <= 0 || n_views > 65535 || out <= 0 || out > 1024 || ks_h <= 0 || ks_v <= 0)
    return set_error(RLCF_ERR_ARG, "resample_taps: bad shape");
  dim3 grid(2, n_views);
  resample_taps_kernel<<<grid, (out + 31) / 32 * 32, 0, stream>>>(geom, out, ks_h, ks_v, hdr, hb, hk, vb, vk);
  RLCF_CHECK_LAUNCH("resample_taps");
  return 0;
}

int resample_u8(const uint8_t* src, int H, int W, int n_views, const int* hdr, const int* hb, const int* hk, int ks_h,
                const int* vb, const int* vk, int ks_v, int out_h, int out_w, uint8_t* tmp, int tmp_rows, uint8_t* out,
                cudaStream_t stream) {
  if (H <= 0 || W <= 0 || n_views <= 0 || n_views > 65535 || ks_h <= 0 || ks_v <= 0 || out_h <= 0 || out_w <= 0 ||
      tmp_rows <= 0)
    return set_error(RLCF_ERR_ARG, "resample_u8: bad shape");
  dim3 gh((tmp_rows * out_w + 255) / 256, n_views), gv((out_h * out_w + 255) / 256, n_views);
  resample_h_kernel<<<gh, 256, 0, stream>>>(src, W, hdr, hb, hk, ks_h, out_w, tmp, tmp_rows);
  RLCF_CHECK_LAUNCH("resample_h");
  resample_v_kernel<<<gv, 256, 0, stream>>>(tmp, tmp_rows, hdr, vb, vk, ks_v, out_h, out_w, out);
  RLCF_CHECK_LAUNCH("resample_v");
  return 0;
}

// ------------------------------------------------------------------------------------------------ bicubic resize
// nn.functional.interpolate(images, size, mode="bicubic", align_corners=True) (TPT/clip_reward.py:133-134): the views
// are resized to the reward model's own input resolution when it is not 224 (e.g. ViT-L/14@336px).  PyTorch's cubic
// convolution kernel (A = -0.75), source index = dst * (in - 1) / (out - 1), border indices clamped.  view_idx gathers
// the selected views.  One thread per output pixel and channel plane.
__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

__global__ void __launch_bounds__(256)
bicubic_resize_kernel(const float* __restrict__ in, const int32_t* __restrict__ view_idx, int C, int H, int W, int oh,
                      int ow, float sy, float sx, long long total, float* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % ow);
    const int y = static_cast<int>((i / ow) % oh);
    const long long plane = i / (static_cast<long long>(ow) * oh);      // view * C + channel
    const int c = static_cast<int>(plane % C);
    const long long v = plane / C;
    const long long sv = view_idx ? view_idx[v] : v;
    const float* src = in + (sv * C + c) * static_cast<long long>(H) * W;
    const float fy = sy * y, fx = sx * x;
    const int iy = static_cast<int>(floorf(fy)), ix = static_cast<int>(floorf(fx));
    float wy[4], wx[4];
    cubic_coeffs(fy - iy, wy);
    cubic_coeffs(fx - ix, wx);
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), H - 1);
      const float* row = src + static_cast<long long>(yy) * W;
      float r = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) r += row[min(max(ix - 1 + b, 0), W - 1)] * wx[b];
      acc += r * wy[a];
    }
    out[i] = acc;
  }
}

int bicubic_resize(const float* in, const int32_t* view_idx, int n_views, int C, int H, int W, int oh, int ow, float* out,
                   cudaStream_t stream) {
  if (n_views <= 0 || C <= 0 || H <= 0 || W <= 0 || oh <= 0 || ow <= 0)
    return set_error(RLCF_ERR_ARG, "bicubic_resize: bad shape");
  const float sy = oh > 1 ? static_cast<float>(H - 1) / static_cast<float>(oh - 1) : 0.f;
  const float sx = ow > 1 ? static_cast<float>(W - 1) / static_cast<float>(ow - 1) : 0.f;
  const long long total = static_cast<long long>(n_views) * C * oh * ow;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  bicubic_resize_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(in, view_idx, C, H, W, oh, ow, sy, sx, total, out);
  RLCF_CHECK_LAUNCH("bicubic_resize");
  return 0;
}

// ------------------------------------------------------------------------------------------------ AugMix
enum { OP_AUTOCONTRAST = 0, OP_EQUALIZE = 1, OP_POSTERIZE = 2, OP_SOLARIZE = 3, OP_AFFINE = 4 };
constexpr int kPlane = 224;
constexpr int kMaxOps = 3, kChains = 3;

__device__ __forceinline__ float preprocess_px(uint8_t u, float mean, float stdv) {
  // ToTensor (x / 255) then Normalize ((x - mean) / std), each a separate fp32 rounding as in torch
  return __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(u), 255.f), mean), stdv);
}

// Geometry.c: affine_transform + bilinear_filter32RGB for one band; returns false where the source point is outside.
__device__ __forceinline__ bool affine_bilinear(const uint8_t* __restrict__ in, const double* __restrict__ a, int x,
                                                int y, uint8_t& out) {
  const double xc = x + 0.5, yc = y + 0.5;
  double xin = __dadd_rn(__dadd_rn(__dmul_rn(a[0], xc), __dmul_rn(a[1], yc)), a[2]);
  double yin = __dadd_rn(__dadd_rn(__dmul_rn(a[3], xc), __dmul_rn(a[4], yc)), a[5]);
  if (xin < 0.0 || xin >= kPlane || yin < 0.0 || yin >= kPlane) return false;
  xin = __dsub_rn(xin, 0.5);
  yin = __dsub_rn(yin, 0.5);
  const int xi = xin < 0.0 ? static_cast<int>(floor(xin)) : static_cast<int>(xin);
  const int yi = yin < 0.0 ? static_cast<int>(floor(yin)) : static_cast<int>(yin);
  const double dx = __dsub_rn(xin, static_cast<double>(xi)), dy = __dsub_rn(yin, static_cast<double>(yi));
  auto clipc = [](int c) { return c < 0 ? 0 : (c < kPlane ? c : kPlane - 1); };
  const int x0 = clipc(xi), x1 = clipc(xi + 1);
  const uint8_t* r0 = in + clipc(yi) * kPlane;
  double v1 = __dadd_rn(static_cast<double>(r0[x0]), __dmul_rn(static_cast<double>(static_cast<int>(r0[x1]) - static_cast<int>(r0[x0])), dx));
  double v2 = v1;
  if (yi + 1 >= 0 && yi + 1 < kPlane) {
    const uint8_t* r1 = in + (yi + 1) * kPlane;
    v2 = __dadd_rn(static_cast<double>(r1[x0]), __dmul_rn(static_cast<double>(static_cast<int>(r1[x1]) - static_cast<int>(r1[x0])), dx));
  }
  v1 = __dadd_rn(v1, __dmul_rn(__dsub_rn(v2, v1), dy));
  out = static_cast<uint8_t>(static_cast<int>(v1));   // (UINT8)v1: truncation
  return true;
}

// grid (3 channels, n_views).  x_orig u8 [V][224][224][3].  vflag[v]: 0 = plain view (out = preprocess(x_orig)).
// wts [V][4] = (w0, w1, w2, m) as float32, omm[v] = float32(1 - m); n_ops [V][3]; ops [V][3][3][2] = (type, int param);
// mats [V][3][3][6] doubles (affine coefficients of op slots whose type is OP_AFFINE).  out f32 [V][3][224][224].
__global__ void __launch_bounds__(256)
augmix_kernel(const uint8_t* __restrict__ x_orig, const int* __restrict__ vflag, const float* __restrict__ wts,
              const float* __restrict__ omm, const int* __restrict__ n_ops, const int* __restrict__ ops,
              const double* __restrict__ mats, float mean0, float mean1, float mean2, float std0, float std1,
              float std2, float* __restrict__ out) {
  extern __shared__ uint8_t sm_u8[];
  uint8_t* bufA = sm_u8;
  uint8_t* bufB = sm_u8 + kPlane * kPlane;
  int* hist = reinterpret_cast<int*>(sm_u8 + 2 * kPlane * kPlane);   // [256]
  int* lut = hist + 256;                                             // [256]
  const int c = blockIdx.x, v = blockIdx.y, tid = threadIdx.x;
  const float mean = c == 0 ? mean0 : (c == 1 ? mean1 : mean2);
  const float stdv = c == 0 ? std0 : (c == 1 ? std1 : std2);
  const uint8_t* src = x_orig + static_cast<size_t>(v) * kPlane * kPlane * 3 + c;
  float* o = out + (static_cast<size_t>(v) * 3 + c) * kPlane * kPlane;
  constexpr int N = kPlane * kPlane;
  if (vflag[v] == 0) {
    for (int i = tid; i < N; i += blockDim.x) o[i] = preprocess_px(src[static_cast<size_t>(i) * 3], mean, stdv);
    return;
  }
  for (int chain = 0; chain < kChains; ++chain) {
    for (int i = tid; i < N; i += blockDim.x) bufA[i] = src[static_cast<size_t>(i) * 3];   // x_aug = x_orig.copy()
    __syncthreads();
    uint8_t* cur = bufA;
    uint8_t* nxt = bufB;
    const int nops = n_ops[v * kChains + chain];
    for (int k = 0; k < nops; ++k) {
      const int type = ops[((v * kChains + chain) * kMaxOps + k) * 2];
      const int param = ops[((v * kChains + chain) * kMaxOps + k) * 2 + 1];
      if (type == OP_AFFINE) {
        const double* a = mats + static_cast<size_t>((v * kChains + chain) * kMaxOps + k) * 6;
        for (int i = tid; i < N; i += blockDim.x) {
          uint8_t px;
          nxt[i] = affine_bilinear(cur, a, i % kPlane, i / kPlane, px) ? px : 0;   // fill = 0 outside
        }
        __syncthreads();
        uint8_t* t = cur; cur = nxt; nxt = t;
        continue;
      }
      if (type == OP_AUTOCONTRAST || type == OP_EQUALIZE) {
        for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < N; i += blockDim.x) atomicAdd(&hist[cur[i]], 1);
        __syncthreads();
        if (tid == 0) {
          if (type == OP_AUTOCONTRAST) {            // ImageOps.autocontrast, cutoff = 0
            int lo = 0, hi = 255;
            while (lo < 256 && hist[lo] == 0) ++lo;
            while (hi >= 0 && hist[hi] == 0) --hi;
            if (hi <= lo) {
              for (int i = 0; i < 256; ++i) lut[i] = i;
            } else {
              const double scale = __ddiv_rn(255.0, static_cast<double>(hi - lo));
              const double offset = __dmul_rn(-static_cast<double>(lo), scale);
              for (int i = 0; i < 256; ++i) {
                int ix = static_cast<int>(__dadd_rn(__dmul_rn(static_cast<double>(i), scale), offset));
                lut[i] = ix < 0 ? 0 : (ix > 255 ? 255 : ix);
              }
            }
          } else {                                  // ImageOps.equalize
            int nonzero = 0, total = 0, last = 0;
            for (int i = 0; i < 256; ++i)
              if (hist[i]) { ++nonzero; total += hist[i]; last = hist[i]; }
            const int step = nonzero <= 1 ? 0 : (total - last) / 255;
            if (step == 0) {
              for (int i = 0; i < 256; ++i) lut[i] = i;
            } else {
              int n = step / 2;
              for (int i = 0; i < 256; ++i) { lut[i] = min(255, n / step); n += hist[i]; }   // Image.point clips to uint8
            }
          }
        }
        __syncthreads();
        for (int i = tid; i < N; i += blockDim.x) cur[i] = static_cast<uint8_t>(lut[cur[i]]);
        __syncthreads();
      } else if (type == OP_POSTERIZE) {            // lut[i] = i & ~(2^(8-bits) - 1), param = bits
        const int mask = ~((1 << (8 - param)) - 1);
        for (int i = tid; i < N; i += blockDim.x) cur[i] = static_cast<uint8_t>(cur[i] & mask);
        __syncthreads();
      } else if (type == OP_SOLARIZE) {             // lut[i] = i if i < threshold else 255 - i, param = threshold
        for (int i = tid; i < N; i += blockDim.x) {
          const int p = cur[i];
          cur[i] = static_cast<uint8_t>(p < param ? p : 255 - p);
        }
        __syncthreads();
      }
    }
    // mix += w[chain] * preprocess(x_aug)          (datautils.py:109)
    const float w = wts[v * 4 + chain];
    for (int i = tid; i < N; i += blockDim.x) {
      const float term = __fmul_rn(w, preprocess_px(cur[i], mean, stdv));
      o[i] = chain == 0 ? __fadd_rn(0.f, term) : __fadd_rn(o[i], term);
    }
    __syncthreads();
  }
  // mix = m * x_processed + (1 - m) * mix          (datautils.py:110)
  const float m = wts[v * 4 + 3], om = omm[v];
  for (int i = tid; i < N; i += blockDim.x) {
    const float xp = preprocess_px(src[static_cast<size_t>(i) * 3], mean, stdv);
    o[i] = __fadd_rn(__fmul_rn(m, xp), __fmul_rn(om, o[i]));
  }
}

int augmix_views(const uint8_t* x_orig, int n_views, const int* vflag, const float* wts, const float* omm,
                 const int* n_ops, const int* ops, const double* mats, const float* mean, const float* stdv,
                 float* out, cudaStream_t stream) {
  if (n_views <= 0 || n_views > 65535) return set_error(RLCF_ERR_ARG, "augmix_views: bad shape");
  const size_t smem = 2 * kPlane * kPlane + 512 * sizeof(int);
  static DynSmemState st;
  if (cudaError_t e = ensure_dyn_smem(augmix_kernel, smem, st))
    return set_error(RLCF_ERR_CUDA, "augmix attr: %s", cudaGetErrorString(e));
  dim3 grid(3, n_views);
  augmix_kernel<<<grid, 256, smem, stream>>>(x_orig, vflag, wts, omm, n_ops, ops, mats, mean[0], mean[1], mean[2],
                                             stdv[0], stdv[1], stdv[2], out);
  RLCF_CHECK_LAUNCH("augmix_views");
  return 0;
}

}  // namespace rlcf
