// EigenPlaces on sm_100a (see eigenplaces.cuh).  Model: gmberton/eigenplaces ResNet18 / 512-d, restated in
// oracle/eigenplaces.py; host contract: /root/reference/src/EigenPlaces.cc:123-174 (preprocess, inference,
// final cv::normalize) and src/PlaceRecognizer.cc:10-52 (index).
#include "eigenplaces.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>

#include "umma_core.cuh"

namespace ssb {

// =================================================================================================
// kernels
// =================================================================================================

// cv::resize(INTER_LINEAR) on 8-bit data in OpenCV's fixed point (11 coefficient bits; horizontal pass in
// int, vertical pass ((b * (S >> 4)) >> 16 ... + 2) >> 2), then convertTo(CV_32F, 1/255) and the ImageNet
// normalisation of EigenPlaces::preprocess, written as fp16 RGB0 pixels.  mode 0: bilinear tables,
// mode 1: exact 2x decimation (OpenCV routes it to the INTER_AREA 2x2 mean), mode 2: same size.
__global__ void __launch_bounds__(256)
ep_preprocess_kernel(const uint8_t* __restrict__ src, int sh, int sw, int cn, const int* __restrict__ tab,
                     int in_h, int in_w, int mode, __half* __restrict__ x0) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (x >= in_w) return;
  const uint8_t* img = src + static_cast<size_t>(b) * sh * sw * cn;
  int rgb[3];
  if (mode == 0) {
    const int* xofs = tab;
    const int* xa0 = tab + in_w;
    const int* xa1 = tab + 2 * in_w;
    const int* yofs = tab + 3 * in_w;
    const int* yb0 = yofs + in_h;
    const int* yb1 = yofs + 2 * in_h;
    const int sx = xofs[x], sx1 = min(sx + 1, sw - 1);
    const int a0 = xa0[x], a1 = xa1[x];
    const int sy = yofs[y];
    const int r0 = min(max(sy, 0), sh - 1), r1 = min(max(sy + 1, 0), sh - 1);
    const int b0 = yb0[y], b1 = yb1[y];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int sc = cn == 1 ? 0 : 2 - c;   // GRAY2RGB replicates, BGR2RGB reverses
      const int p00 = img[(static_cast<size_t>(r0) * sw + sx) * cn + sc], p01 = img[(static_cast<size_t>(r0) * sw + sx1) * cn + sc];
      const int p10 = img[(static_cast<size_t>(r1) * sw + sx) * cn + sc], p11 = img[(static_cast<size_t>(r1) * sw + sx1) * cn + sc];
      const int S0 = (p00 * a0 + p01 * a1) >> 4, S1 = (p10 * a0 + p11 * a1) >> 4;
      const int v = (((b0 * S0) >> 16) + ((b1 * S1) >> 16) + 2) >> 2;
      rgb[c] = min(max(v, 0), 255);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int sc = cn == 1 ? 0 : 2 - c;
      if (mode == 1) {
        const size_t o = (static_cast<size_t>(2 * y) * sw + 2 * x) * cn + sc;
        rgb[c] = (img[o] + img[o + cn] + img[o + static_cast<size_t>(sw) * cn] + img[o + static_cast<size_t>(sw) * cn + cn] + 2) >> 2;
      } else {
        rgb[c] = img[(static_cast<size_t>(y) * sw + x) * cn + sc];
      }
    }
  }
  const float mean[3] = {0.485f, 0.456f, 0.406f}, sd[3] = {0.229f, 0.224f, 0.225f};
  __half h[4];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = __fmul_rn(static_cast<float>(rgb[c]), 1.0f / 255.0f);
    h[c] = __float2half(__fdiv_rn(__fsub_rn(v, mean[c]), sd[c]));
  }
  h[3] = __float2half(0.f);
  *reinterpret_cast<uint2*>(x0 + ((static_cast<size_t>(b) * in_h + y) * in_w + x) * 4) = *reinterpret_cast<uint2*>(h);
}

// im2col of the 7x7 / stride 2 / pad 3 stem: one warp per output pixel, lanes 0..23 write eight consecutive
// columns each; column k = (kh*7 + kw)*3 + c for k < 147, zero above.
__global__ void __launch_bounds__(256)
ep_im2col_kernel(const __half* __restrict__ x0, int in_h, int in_w, int Ho, int Wo, __half* __restrict__ col) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pix = blockIdx.x * 8 + warp;
  const int b = blockIdx.y;
  if (pix >= Ho * Wo || lane >= kEpStemK / 8) return;
  const int yo = pix / Wo, xo = pix - yo * Wo;
  const __half* img = x0 + static_cast<size_t>(b) * in_h * in_w * 4;
  __half v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = lane * 8 + i;
    __half r = __float2half(0.f);
    if (k < 147) {
      const int tap = k / 3, c = k - tap * 3;
      const int kh = tap / 7, kw = tap - kh * 7;
      const int yi = 2 * yo + kh - 3, xi = 2 * xo + kw - 3;
      if (yi >= 0 && yi < in_h && xi >= 0 && xi < in_w) r = img[(static_cast<size_t>(yi) * in_w + xi) * 4 + c];
    }
    v[i] = r;
  }
  *reinterpret_cast<uint4*>(col + (static_cast<size_t>(b) * Ho * Wo + pix) * kEpStemK + lane * 8) =
      *reinterpret_cast<uint4*>(v);
}

// max_pool2d(3, stride 2, pad 1) on NHWC fp16, eight channels per thread.
__global__ void __launch_bounds__(256)
ep_maxpool_kernel(const __half* __restrict__ in, int H, int W, int C, int Ho, int Wo, __half* __restrict__ out) {
  const int groups = C / 8;
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= static_cast<size_t>(Ho) * Wo * groups) return;
  const int g = static_cast<int>(idx % groups);
  const int pix = static_cast<int>(idx / groups);
  const int yo = pix / Wo, xo = pix - yo * Wo;
  __half2 m[4];
  bool first = true;
  for (int dy = -1; dy <= 1; ++dy) {
    const int yi = 2 * yo + dy;
    if (yi < 0 || yi >= H) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int xi = 2 * xo + dx;
      if (xi < 0 || xi >= W) continue;
      const uint4 raw = *reinterpret_cast<const uint4*>(in + ((static_cast<size_t>(b) * H + yi) * W + xi) * C + g * 8);
      const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int i = 0; i < 4; ++i) m[i] = first ? h2[i] : __hmax2(m[i], h2[i]);
      first = false;
    }
  }
  *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(b) * Ho + yo) * Wo + xo) * C + g * 8) = *reinterpret_cast<uint4*>(m);
}

// Aggregation head, one CTA (512 threads) per image:
//   F.normalize(x, dim=channels)  ->  GeM: (mean_p clamp(x, 1e-6)^p)^(1/p)  ->  Linear(512, 512)  ->  F.normalize
__global__ void __launch_bounds__(kEpDim)
ep_head_kernel(const __half* __restrict__ feat, int npix, float gem_p, const float* __restrict__ wt,
               const float* __restrict__ bias, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* inv = sm;            // [npix]
  float* g = sm + npix;       // [512]
  float* red = g + kEpDim;    // [16]
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __half* f = feat + static_cast<size_t>(b) * npix * kEpDim;
  for (int p = warp; p < npix; p += kEpDim / 32) {
    const uint4* row = reinterpret_cast<const uint4*>(f + static_cast<size_t>(p) * kEpDim);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint4 raw = row[lane + 32 * i];
      const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 v = __half22float2(h2[j]);
        ss = fmaf(v.x, v.x, ss);
        ss = fmaf(v.y, v.y, ss);
      }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);
    if (lane == 0) inv[p] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  }
  __syncthreads();
  const int c = threadIdx.x;
  float acc = 0.f;
  for (int p = 0; p < npix; ++p) {
    const float v = fmaxf(__half2float(f[static_cast<size_t>(p) * kEpDim + c]) * inv[p], 1e-6f);
    acc += powf(v, gem_p);
  }
  g[c] = powf(acc / static_cast<float>(npix), 1.0f / gem_p);
  __syncthreads();
  float o = bias[c];
  for (int k = 0; k < kEpDim; ++k) o = fmaf(wt[static_cast<size_t>(k) * kEpDim + c], g[k], o);
  float ss = o * o;
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < kEpDim / 32; ++i) tot += red[i];
  out[static_cast<size_t>(b) * kEpDim + c] = o / fmaxf(sqrtf(tot), 1e-12f);
}

// scores[i] = <db[i], q>, one warp per database row (CosineDescriptorIndex::query's cand * q^T).
__global__ void __launch_bounds__(256)
ep_index_scores_kernel(const float* __restrict__ db, const float* __restrict__ q, int rows, int dim,
                       float* __restrict__ scores) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= rows) return;
  const float* row = db + static_cast<size_t>(r) * dim;
  float acc = 0.f;
  for (int k = lane; k < dim; k += 32) acc = fmaf(row[k], q[k], acc);
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
  if (lane == 0) scores[r] = acc;
}

// ---- TMEM epilogue: folded-BatchNorm bias (+ residual) (+ ReLU) -> fp16 NHWC through a staged TMA store.
// kSplitCols: the two warps of a TMEM lane quadrant split the tile's columns (needs >= 64 columns each).
template <bool kSplitCols>
struct EpiBnAct {
  const float* bias;
  CUtensorMap tm_out;      // 4-D (C, W, H, B), box (64, 16, 2, 1)
  const __half* res;       // residual, same NHWC geometry as the output, or nullptr
  int H, W, C;
  int relu;
  static constexpr bool kSplit = kSplitCols;
  __device__ void operator()(EpiCtx& c, bool has_acc) const {
    const int px0 = __shfl_sync(0xffffffffu, c.px, 0), py0 = __shfl_sync(0xffffffffu, c.py, 0);
    const bool inside = c.px < W && c.py < H;
    const __half* rrow = res != nullptr && inside
                             ? res + ((static_cast<size_t>(c.z) * H + c.py) * W + c.px) * C + c.n0
                             : nullptr;
    for (int g0 = c.col_begin; g0 < c.col_end; g0 += 64) {
      stage_begin(c);
#pragma unroll 1
      for (int hc = 0; hc < 2; ++hc) {
        const int col = g0 + hc * 32;
        float v[32];
        tmem_ld_32x32(c.tmem_row + col, v);
        uint4 rr[4] = {};
        if (rrow != nullptr) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rr[j] = *reinterpret_cast<const uint4*>(rrow + col + 8 * j);
        }
        const float4* b4 = reinterpret_cast<const float4*>(bias + c.n0 + col);
        float bv[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = __ldg(b4 + j);
          bv[4 * j] = t.x, bv[4 * j + 1] = t.y, bv[4 * j + 2] = t.z, bv[4 * j + 3] = t.w;
        }
        tmem_ld_wait();
        const __half2* r2 = reinterpret_cast<const __half2*>(rr);
        uint32_t h[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 r = __half22float2(r2[j]);
          float x0 = (has_acc ? v[2 * j] : 0.f) + bv[2 * j] + r.x;
          float x1 = (has_acc ? v[2 * j + 1] : 0.f) + bv[2 * j + 1] + r.y;
          if (relu) {
            x0 = fmaxf(x0, 0.f);
            x1 = fmaxf(x1, 0.f);
          }
          h[j] = pack_half2(x0, x1);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          stage_put(c, c.lane, hc * 4 + j, make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]));
      }
      stage_fence(c);
      if (c.lane == 0) {
        tma_store_4d(&tm_out, c.stage_cur, c.n0 + g0, px0, py0, c.z);
        bulk_commit();
      }
    }
  }
};

// =================================================================================================
// host
// =================================================================================================
EigenPlaces::~EigenPlaces() {
  cudaSetDevice(device_);
  for (void* p : owned_)
    if (p) cudaFree(p);
  if (src_dev_) cudaFree(src_dev_);
  if (tab_dev_) cudaFree(tab_dev_);
  if (db_) cudaFree(db_);
  if (q_dev_) cudaFree(q_dev_);
  if (sc_dev_) cudaFree(sc_dev_);
  if (src_host_) cudaFreeHost(src_host_);
  if (out_host_) cudaFreeHost(out_host_);
  if (io_host_) cudaFreeHost(io_host_);
  if (stream_) cudaStreamDestroy(stream_);
}

// conv weight (cout, cin, t, t) + BatchNorm (eps 1e-5) -> fp16 [tap][cout][k_per_tap] with the BN scale
// folded in, fp32 bias = beta - mean * scale.  The 7x7 stem becomes one [cout][192] im2col matrix.
int EigenPlaces::load_conv(const WeightArchive& ar, const std::string& conv, const std::string& bn, int cin,
                           int cout, int taps, int stride, EpConv* L) {
  const HostTensor* w = ar.get(conv + ".weight", {cout, cin, taps, taps});
  const HostTensor* ga = ar.get(bn + ".weight", {cout});
  const HostTensor* be = ar.get(bn + ".bias", {cout});
  const HostTensor* mu = ar.get(bn + ".running_mean", {cout});
  const HostTensor* va = ar.get(bn + ".running_var", {cout});
  if (!w || !ga || !be || !mu || !va) return SSB_ERR_IO;
  const bool stem = taps == 7;
  const int T = stem ? 1 : taps * taps;
  const int kp = stem ? kEpStemK : cin;
  SSB_CHECK(stem || cin % 64 == 0, SSB_ERR_INVALID, "%s: input channels must be a multiple of 64", conv.c_str());
  L->cin = cin, L->cout = cout, L->taps = stem ? 1 : taps, L->stride = stem ? 1 : stride, L->k_per_tap = kp;
  std::vector<__half> hw(static_cast<size_t>(T) * cout * kp, __float2half(0.f));
  std::vector<float> hb(cout);
  for (int co = 0; co < cout; ++co) {
    const float scale = ga->data[co] / std::sqrt(va->data[co] + 1e-5f);
    hb[co] = be->data[co] - mu->data[co] * scale;
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < taps * taps; ++t) {
        const float v = w->data[(static_cast<size_t>(co) * cin + ci) * taps * taps + t] * scale;
        const size_t dst = stem ? static_cast<size_t>(co) * kp + t * 3 + ci
                                : (static_cast<size_t>(t) * cout + co) * kp + ci;
        hw[dst] = __float2half(v);
      }
  }
  SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&L->w), hw.size() * sizeof(__half)));
  owned_.push_back(L->w);
  SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&L->bias), hb.size() * sizeof(float)));
  owned_.push_back(L->bias);
  SSB_CUDA_CHECK(cudaMemcpy(L->w, hw.data(), hw.size() * sizeof(__half), cudaMemcpyHostToDevice));
  SSB_CUDA_CHECK(cudaMemcpy(L->bias, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice));
  uint64_t dims[3] = {static_cast<uint64_t>(kp), static_cast<uint64_t>(T) * cout, 1};
  uint64_t strides[2] = {static_cast<uint64_t>(kp) * 2, static_cast<uint64_t>(T) * cout * kp * 2};
  uint32_t box[3] = {64, static_cast<uint32_t>(cout > 256 ? 256 : cout), 1};
  return encode_tmap_f16(&L->tmB, L->w, 3, dims, strides, box);
}

// A-operand map over an NHWC activation: 16 x 8 output pixels per tile; stride 2 reads every second pixel.
static int ep_load_map(CUtensorMap* tm, const __half* base, int C, int W, int H, int B, int stride) {
  uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                      static_cast<uint64_t>(B)};
  uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(W) * C * 2,
                         static_cast<uint64_t>(H) * W * C * 2};
  uint32_t box[4] = {64, static_cast<uint32_t>(16 * stride), static_cast<uint32_t>(8 * stride), 1};
  uint32_t es[4] = {1, static_cast<uint32_t>(stride), static_cast<uint32_t>(stride), 1};
  return encode_tmap_f16(tm, base, 4, dims, strides, box, es);
}
static int ep_store_map(CUtensorMap* tm, const __half* base, int C, int W, int H, int B) {
  uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                      static_cast<uint64_t>(B)};
  uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(W) * C * 2,
                         static_cast<uint64_t>(H) * W * C * 2};
  uint32_t box[4] = {64, 16, 2, 1};
  return encode_tmap_f16(tm, base, 4, dims, strides, box);
}

int EigenPlaces::init(const char* weights_path, int in_w, int in_h, int max_batch, int device) {
  SSB_CHECK(weights_path != nullptr, SSB_ERR_INVALID, "null weights path");
  SSB_CHECK(in_w >= 64 && in_h >= 64 && in_w <= 2048 && in_h <= 2048 && in_w % 32 == 0 && in_h % 32 == 0,
            SSB_ERR_INVALID, "input size %dx%d: both sides must be multiples of 32 in [64, 2048]", in_w, in_h);
  SSB_CHECK(max_batch >= 1 && max_batch <= 64, SSB_ERR_INVALID, "max_batch out of range (1..64)");
  device_ = device, in_w_ = in_w, in_h_ = in_h, max_batch_ = max_batch;
  SSB_CUDA_CHECK(cudaSetDevice(device));
  SSB_CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  WeightArchive ar;
  SSB_RETURN_IF(load_archive(weights_path, &ar));
  SSB_RETURN_IF(load_conv(ar, "backbone.0", "backbone.1", 3, 64, 7, 2, &stem_));
  const int chans[4] = {64, 128, 256, 512};
  int cin = 64;
  for (int s = 0; s < 4; ++s) {
    EpStage& S = stage_[s];
    S.C = chans[s];
    const int stride = s == 0 ? 1 : 2;
    const std::string base = "backbone." + std::to_string(4 + s) + ".";
    for (int blk = 0; blk < 2; ++blk) {
      const std::string p = base + std::to_string(blk) + ".";
      SSB_RETURN_IF(load_conv(ar, p + "conv1", p + "bn1", blk == 0 ? cin : S.C, S.C, 3, blk == 0 ? stride : 1, &S.c1[blk]));
      SSB_RETURN_IF(load_conv(ar, p + "conv2", p + "bn2", S.C, S.C, 3, 1, &S.c2[blk]));
    }
    S.has_ds = ar.has(base + "0.downsample.0.weight");
    SSB_CHECK(S.has_ds == (s != 0), SSB_ERR_IO, "unexpected downsample layout in stage %d", s + 1);
    if (S.has_ds) SSB_RETURN_IF(load_conv(ar, base + "0.downsample.0", base + "0.downsample.1", cin, S.C, 1, stride, &S.ds));
    cin = S.C;
  }
  {
    const HostTensor* p = ar.get("aggregation.1.p", {1});
    const HostTensor* w = ar.get("aggregation.3.weight", {kEpDim, kEpDim});
    const HostTensor* b = ar.get("aggregation.3.bias", {kEpDim});
    if (!p || !w || !b) return SSB_ERR_IO;
    gem_p_ = p->data[0];
    std::vector<float> wt(static_cast<size_t>(kEpDim) * kEpDim);
    for (int o = 0; o < kEpDim; ++o)
      for (int k = 0; k < kEpDim; ++k) wt[static_cast<size_t>(k) * kEpDim + o] = w->data[static_cast<size_t>(o) * kEpDim + k];
    SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&fc_wt_), wt.size() * 4));
    owned_.push_back(fc_wt_);
    SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&fc_b_), kEpDim * 4));
    owned_.push_back(fc_b_);
    SSB_CUDA_CHECK(cudaMemcpy(fc_wt_, wt.data(), wt.size() * 4, cudaMemcpyHostToDevice));
    SSB_CUDA_CHECK(cudaMemcpy(fc_b_, b->data.data(), kEpDim * 4, cudaMemcpyHostToDevice));
  }
  // activations
  const size_t B = max_batch;
  auto alloc = [&](__half** p, size_t halves) -> int {
    SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(p), halves * 2 + 1024));
    SSB_CUDA_CHECK(cudaMemset(*p, 0, halves * 2 + 1024));
    owned_.push_back(*p);
    return SSB_OK;
  };
  const int H2 = in_h / 2, W2 = in_w / 2, H4 = in_h / 4, W4 = in_w / 4;
  SSB_RETURN_IF(alloc(&x0_, B * in_h * in_w * 4));
  SSB_RETURN_IF(alloc(&col_, B * H2 * W2 * kEpStemK));
  SSB_RETURN_IF(alloc(&s0_, B * H2 * W2 * 64));
  SSB_RETURN_IF(alloc(&p0_, B * H4 * W4 * 64));
  SSB_RETURN_IF(ep_load_map(&ld_col_, col_, kEpStemK, W2, H2, max_batch, 1));
  SSB_RETURN_IF(ep_store_map(&st_s0_, s0_, 64, W2, H2, max_batch));
  const __half* prev = p0_;
  int pc = 64, ph = H4, pw = W4;
  for (int s = 0; s < 4; ++s) {
    EpStage& S = stage_[s];
    const int stride = s == 0 ? 1 : 2;
    S.H = ph / stride, S.W = pw / stride;
    const size_t n = B * S.H * S.W * S.C;
    SSB_RETURN_IF(alloc(&S.t, n));
    SSB_RETURN_IF(alloc(&S.a, n));
    SSB_RETURN_IF(alloc(&S.b, n));
    if (S.has_ds) SSB_RETURN_IF(alloc(&S.d, n));
    SSB_RETURN_IF(ep_load_map(&S.ld_in, prev, pc, pw, ph, max_batch, stride));
    SSB_RETURN_IF(ep_load_map(&S.ld_t, S.t, S.C, S.W, S.H, max_batch, 1));
    SSB_RETURN_IF(ep_load_map(&S.ld_a, S.a, S.C, S.W, S.H, max_batch, 1));
    SSB_RETURN_IF(ep_store_map(&S.st_t, S.t, S.C, S.W, S.H, max_batch));
    SSB_RETURN_IF(ep_store_map(&S.st_a, S.a, S.C, S.W, S.H, max_batch));
    SSB_RETURN_IF(ep_store_map(&S.st_b, S.b, S.C, S.W, S.H, max_batch));
    if (S.has_ds) SSB_RETURN_IF(ep_store_map(&S.st_d, S.d, S.C, S.W, S.H, max_batch));
    prev = S.b, pc = S.C, ph = S.H, pw = S.W;
  }
  SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&out_dev_), B * kEpDim * 4));
  owned_.push_back(out_dev_);
  SSB_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&out_host_), B * kEpDim * 4));
  SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&tab_dev_), static_cast<size_t>(3) * (in_w + in_h) * 4));
  return SSB_OK;
}

// cv::resize's coefficient tables for this source size (modules/imgproc resize.cpp, INTER_LINEAR, 8U):
// fx = (float)((dx + 0.5) * scale - 0.5), sx = floor(fx), clamped taps get fx = 0 horizontally (rows are
// clipped instead), weights = saturate_cast<short>(w * 2048).
static void ep_axis_table(int src, int dst, bool clamp_frac, int* ofs, int* c0, int* c1) {
  const double scale = static_cast<double>(src) / dst;
  for (int d = 0; d < dst; ++d) {
    float f = static_cast<float>((d + 0.5) * scale - 0.5);
    int s = static_cast<int>(std::floor(f));
    f -= static_cast<float>(s);
    if (clamp_frac) {
      if (s < 0) f = 0.f, s = 0;
      if (s >= src - 1) f = 0.f, s = src - 1;
    }
    ofs[d] = s;
    c0[d] = static_cast<int>(std::max(-32768.0, std::min(32767.0, std::nearbyint(static_cast<double>((1.f - f) * 2048.f)))));
    c1[d] = static_cast<int>(std::max(-32768.0, std::min(32767.0, std::nearbyint(static_cast<double>(f * 2048.f)))));
  }
}

int EigenPlaces::ensure_source(int h, int w, int channels) {
  if (h == src_h_ && w == src_w_ && channels == src_c_) return SSB_OK;
  SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
  const size_t bytes = static_cast<size_t>(max_batch_) * h * w * channels;
  if (bytes > src_bytes_) {
    if (src_dev_) cudaFree(src_dev_);
    if (src_host_) cudaFreeHost(src_host_);
    src_dev_ = nullptr, src_host_ = nullptr, src_bytes_ = 0;
    SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&src_dev_), bytes));
    SSB_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&src_host_), bytes));
    src_bytes_ = bytes;
  }
  if (w == in_w_ && h == in_h_) {
    resize_mode_ = 2;
  } else if (w == 2 * in_w_ && h == 2 * in_h_) {
    resize_mode_ = 1;
  } else {
    resize_mode_ = 0;
    std::vector<int> tab(static_cast<size_t>(3) * (in_w_ + in_h_));
    ep_axis_table(w, in_w_, true, tab.data(), tab.data() + in_w_, tab.data() + 2 * in_w_);
    int* ty = tab.data() + 3 * in_w_;
    ep_axis_table(h, in_h_, false, ty, ty + in_h_, ty + 2 * in_h_);
    SSB_CUDA_CHECK(cudaMemcpy(tab_dev_, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
  }
  src_h_ = h, src_w_ = w, src_c_ = channels;
  return SSB_OK;
}

int EigenPlaces::conv(const EpConv& L, const CUtensorMap& in, const CUtensorMap& out, const __half* residual,
                      bool relu, int Ho, int Wo, int batch, const char* label) {
  CoreParams p;
  std::memset(&p, 0, sizeof(p));
  p.label = label;
  p.taps_h = p.taps_w = L.taps;
  p.pad = L.taps / 2;
  p.a_stride = L.stride;
  p.kc0 = L.k_per_tap / 64;
  p.b_tap_rows = L.cout;
  p.tile_w = 16;
  p.tile_h = 8;
  p.tiles_w = (Wo + 15) / 16;
  p.block_n = L.cout > 256 ? 256 : L.cout;
  p.a_z_mul = 1;
  const dim3 grid(p.tiles_w * ((Ho + 7) / 8), L.cout / p.block_n, batch);
  if (p.block_n < 128) {
    EpiBnAct<false> e{L.bias, out, residual, Ho, Wo, L.cout, relu ? 1 : 0};
    return launch_core(in, in, L.tmB, p, e, grid, stream_);
  }
  EpiBnAct<true> e{L.bias, out, residual, Ho, Wo, L.cout, relu ? 1 : 0};
  return launch_core(in, in, L.tmB, p, e, grid, stream_);
}

int EigenPlaces::run(int batch) {
  const int H2 = in_h_ / 2, W2 = in_w_ / 2, H4 = in_h_ / 4, W4 = in_w_ / 4;
  prof_begin(stream_);
  ep_preprocess_kernel<<<dim3((in_w_ + 255) / 256, in_h_, batch), 256, 0, stream_>>>(
      src_dev_, src_h_, src_w_, src_c_, tab_dev_, in_h_, in_w_, resize_mode_, x0_);
  SSB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  prof_mark(stream_, "ep.preprocess");
  ep_im2col_kernel<<<dim3((H2 * W2 + 7) / 8, batch), 256, 0, stream_>>>(x0_, in_h_, in_w_, H2, W2, col_);
  SSB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  prof_mark(stream_, "ep.im2col");
  SSB_RETURN_IF(conv(stem_, ld_col_, st_s0_, nullptr, true, H2, W2, batch, "ep.stem"));
  {
    const size_t n = static_cast<size_t>(H4) * W4 * 8;
    ep_maxpool_kernel<<<dim3(static_cast<unsigned>((n + 255) / 256), batch), 256, 0, stream_>>>(s0_, H2, W2, 64, H4, W4, p0_);
    SSB_CUDA_CHECK(cudaGetLastError());
    count_launch();
    prof_mark(stream_, "ep.maxpool");
  }
  const __half* stage_in = p0_;
  static const char* names[4] = {"ep.layer1", "ep.layer2", "ep.layer3", "ep.layer4"};
  for (int s = 0; s < 4; ++s) {
    EpStage& S = stage_[s];
    // block 0: t = relu(conv1(in)); [d = ds(in)]; a = relu(conv2(t) + (d | in))
    SSB_RETURN_IF(conv(S.c1[0], S.ld_in, S.st_t, nullptr, true, S.H, S.W, batch, names[s]));
    if (S.has_ds) SSB_RETURN_IF(conv(S.ds, S.ld_in, S.st_d, nullptr, false, S.H, S.W, batch, names[s]));
    SSB_RETURN_IF(conv(S.c2[0], S.ld_t, S.st_a, S.has_ds ? S.d : stage_in, true, S.H, S.W, batch, names[s]));
    // block 1: t = relu(conv1(a)); b = relu(conv2(t) + a)
    SSB_RETURN_IF(conv(S.c1[1], S.ld_a, S.st_t, nullptr, true, S.H, S.W, batch, names[s]));
    SSB_RETURN_IF(conv(S.c2[1], S.ld_t, S.st_b, S.a, true, S.H, S.W, batch, names[s]));
    stage_in = S.b;
  }
  const int npix = stage_[3].H * stage_[3].W;
  const size_t smem = (static_cast<size_t>(npix) + kEpDim + 16) * 4;
  SSB_CHECK(smem <= 48 * 1024, SSB_ERR_INVALID, "input too large for the aggregation head");
  ep_head_kernel<<<batch, kEpDim, smem, stream_>>>(stage_[3].b, npix, gem_p_, fc_wt_, fc_b_, out_dev_);
  SSB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  prof_mark(stream_, "ep.head");
  return SSB_OK;
}

int EigenPlaces::compute(const uint8_t* const* images, int count, int h, int w, int channels, int row_stride,
                         float* out) {
  SSB_CHECK(images != nullptr && out != nullptr && count >= 1, SSB_ERR_INVALID, "bad arguments");
  SSB_CHECK(channels == 1 || channels == 3, SSB_ERR_INVALID, "images must be 8-bit gray or BGR");
  SSB_CHECK(h >= 2 && w >= 2 && h <= 8192 && w <= 8192 && row_stride >= w * channels, SSB_ERR_INVALID, "bad image geometry");
  SSB_CUDA_CHECK(cudaSetDevice(device_));
  SSB_RETURN_IF(ensure_source(h, w, channels));
  const size_t img_bytes = static_cast<size_t>(h) * w * channels;
  for (int first = 0; first < count; first += max_batch_) {
    const int n = std::min(max_batch_, count - first);
    for (int i = 0; i < n; ++i) {
      SSB_CHECK(images[first + i] != nullptr, SSB_ERR_INVALID, "image %d is null", first + i);
      for (int y = 0; y < h; ++y)
        std::memcpy(src_host_ + i * img_bytes + static_cast<size_t>(y) * w * channels,
                    images[first + i] + static_cast<size_t>(y) * row_stride, static_cast<size_t>(w) * channels);
    }
    SSB_CUDA_CHECK(cudaMemcpyAsync(src_dev_, src_host_, n * img_bytes, cudaMemcpyHostToDevice, stream_));
    SSB_RETURN_IF(run(n));
    SSB_CUDA_CHECK(cudaMemcpyAsync(out_host_, out_dev_, static_cast<size_t>(n) * kEpDim * 4, cudaMemcpyDeviceToHost, stream_));
    SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
    // cv::normalize(desc, desc, 1.0, 0.0, NORM_L2) (src/EigenPlaces.cc:172): norm accumulated in double
    for (int i = 0; i < n; ++i) {
      const float* d = out_host_ + static_cast<size_t>(i) * kEpDim;
      double ss = 0.0;
      for (int k = 0; k < kEpDim; ++k) ss += static_cast<double>(d[k]) * d[k];
      const float nrm = static_cast<float>(std::sqrt(ss));
      float* o = out + static_cast<size_t>(first + i) * kEpDim;
      for (int k = 0; k < kEpDim; ++k) o[k] = nrm > 0.f ? d[k] / nrm : d[k];
    }
  }
  return SSB_OK;
}

// normalizedRow (src/PlaceRecognizer.cc:10-19): L2 norm in double, divide only when > 1e-12.
static void ep_normalized_row(const float* in, int dim, float* out) {
  double ss = 0.0;
  for (int k = 0; k < dim; ++k) ss += static_cast<double>(in[k]) * in[k];
  const double n = std::sqrt(ss);
  for (int k = 0; k < dim; ++k) out[k] = n > 1e-12 ? static_cast<float>(in[k] / n) : in[k];
}

int EigenPlaces::add(uint64_t keyframe_id, const float* desc, int dim) {
  SSB_CHECK(desc != nullptr && dim >= 1 && dim <= 65536, SSB_ERR_INVALID, "bad descriptor");
  SSB_CHECK(db_dim_ == 0 || dim == db_dim_, SSB_ERR_INVALID, "descriptor width %d differs from the index's %d", dim, db_dim_);
  SSB_CUDA_CHECK(cudaSetDevice(device_));
  if (db_dim_ == 0) {
    db_dim_ = dim;
    io_host_floats_ = static_cast<size_t>(dim);
    SSB_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&io_host_), io_host_floats_ * 4));
    SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&q_dev_), static_cast<size_t>(dim) * 4));
  }
  const int M = static_cast<int>(ids_.size());
  if (M + 1 > db_cap_) {   // grow by doubling; rows keep their insertion order (recency)
    const int cap = std::max(256, db_cap_ * 2);
    float* nd = nullptr;
    SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&nd), static_cast<size_t>(cap) * dim * 4));
    if (db_) {
      SSB_CUDA_CHECK(cudaMemcpy(nd, db_, static_cast<size_t>(M) * dim * 4, cudaMemcpyDeviceToDevice));
      cudaFree(db_);
    }
    db_ = nd, db_cap_ = cap;
  }
  std::vector<float> row(dim);
  ep_normalized_row(desc, dim, row.data());
  SSB_CUDA_CHECK(cudaMemcpy(db_ + static_cast<size_t>(M) * dim, row.data(), static_cast<size_t>(dim) * 4, cudaMemcpyHostToDevice));
  ids_.push_back(keyframe_id);
  return SSB_OK;
}

int EigenPlaces::query(const float* desc, int dim, uint64_t exclude_recent, int top_k, float min_score,
                       uint64_t* ids, float* scores, int capacity, int* n_out) {
  SSB_CHECK(n_out != nullptr, SSB_ERR_INVALID, "n_out is null");
  *n_out = 0;
  const size_t M = ids_.size();
  if (M == 0 || M <= exclude_recent) return SSB_OK;   // nothing old enough to be a loop (:31-32)
  SSB_CHECK(desc != nullptr && dim == db_dim_, SSB_ERR_INVALID, "descriptor width %d differs from the index's %d", dim, db_dim_);
  SSB_CUDA_CHECK(cudaSetDevice(device_));
  const int limit = static_cast<int>(M - exclude_recent);
  if (limit > sc_cap_) {
    if (sc_dev_) cudaFree(sc_dev_);
    sc_dev_ = nullptr;
    const int cap = std::max(1024, limit * 2);
    SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&sc_dev_), static_cast<size_t>(cap) * 4));
    sc_cap_ = cap;
  }
  std::vector<float> q(dim);
  ep_normalized_row(desc, dim, q.data());
  SSB_CUDA_CHECK(cudaMemcpyAsync(q_dev_, q.data(), static_cast<size_t>(dim) * 4, cudaMemcpyHostToDevice, stream_));
  ep_index_scores_kernel<<<(limit + 7) / 8, 256, 0, stream_>>>(db_, q_dev_, limit, dim, sc_dev_);
  SSB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  std::vector<float> sc(limit);
  SSB_CUDA_CHECK(cudaMemcpyAsync(sc.data(), sc_dev_, static_cast<size_t>(limit) * 4, cudaMemcpyDeviceToHost, stream_));
  SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
  std::vector<std::pair<float, uint64_t>> cand;
  for (int i = 0; i < limit; ++i)
    if (sc[i] >= min_score) cand.emplace_back(sc[i], ids_[i]);
  std::stable_sort(cand.begin(), cand.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
  size_t n = cand.size();
  if (top_k > 0 && n > static_cast<size_t>(top_k)) n = top_k;
  if (n > static_cast<size_t>(std::max(capacity, 0))) n = std::max(capacity, 0);
  for (size_t i = 0; i < n; ++i) {
    if (ids) ids[i] = cand[i].second;
    if (scores) scores[i] = cand[i].first;
  }
  *n_out = static_cast<int>(n);
  return SSB_OK;
}

int EigenPlaces::debug_read(const char* what, void* dst, size_t bytes) {
  SSB_CUDA_CHECK(cudaSetDevice(device_));
  SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
  const std::string k(what ? what : "");
  const size_t B = max_batch_;
  struct Ent {
    const char* name;
    const void* ptr;
    size_t bytes;
  } tab[] = {
      {"x0", x0_, B * in_h_ * in_w_ * 4 * 2},
      {"col", col_, B * (in_h_ / 2) * (in_w_ / 2) * kEpStemK * 2},
      {"stem", s0_, B * (in_h_ / 2) * (in_w_ / 2) * 64 * 2},
      {"pool", p0_, B * (in_h_ / 4) * (in_w_ / 4) * 64 * 2},
      {"layer1", stage_[0].b, B * stage_[0].H * stage_[0].W * stage_[0].C * 2},
      {"layer2", stage_[1].b, B * stage_[1].H * stage_[1].W * stage_[1].C * 2},
      {"layer2.t", stage_[1].t, B * stage_[1].H * stage_[1].W * stage_[1].C * 2},
      {"layer2.d", stage_[1].d, B * stage_[1].H * stage_[1].W * stage_[1].C * 2},
      {"layer3", stage_[2].b, B * stage_[2].H * stage_[2].W * stage_[2].C * 2},
      {"layer4", stage_[3].b, B * stage_[3].H * stage_[3].W * stage_[3].C * 2},
      {"desc", out_dev_, B * kEpDim * 4},
  };
  for (const Ent& e : tab) {
    if (k == e.name) {
      SSB_CHECK(bytes <= e.bytes, SSB_ERR_INVALID, "debug_read: '%s' holds %zu bytes, asked %zu", what, e.bytes, bytes);
      SSB_CUDA_CHECK(cudaMemcpy(dst, e.ptr, bytes, cudaMemcpyDeviceToHost));
      return SSB_OK;
    }
  }
  set_last_error("debug_read: unknown buffer '%s'", what);
  return SSB_ERR_INVALID;
}

}  // namespace ssb
