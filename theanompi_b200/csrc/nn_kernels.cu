// Fused non-GEMM layer kernels for sm_100a: LRN fwd/bwd, max/avg pooling fwd/bwd, dropout (Philox),
// softmax + NLL + top-1/top-5 error (+ dlogits), ReLU-mask + bias-gradient reduction, NHWC im2col /
// col2im-gather for the implicit-GEMM convolutions, normalise+crop+mirror for the loader.
// All activations are NHWC bf16 with C % 8 == 0 (16-byte vectors) unless noted; math is fp32.
// Reference ops: theanompi/models/layers2.py (LRN :753-809, Pool :402-428, Dropout :864-908,
// Softmax :937-997, Crop/Subtract :223-347) and data/utils.py:42-129 (crop_and_mirror).
#include "common.cuh"
#include "api.h"
#include <algorithm>

namespace tmpi {

static inline int grid_for(long long n, int block) { return (int)((n + block - 1) / block); }

// ============================================================================ LRN
template <int HALF>
__global__ void lrn_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long rows, int C,
                               float k, float alpha, float beta) {
  const int nvec = C >> 3;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;           // host guarantees rows * nvec < 2^32
  if (idx >= (unsigned)rows * (unsigned)nvec) return;
  const long long r = idx / (unsigned)nvec;
  const int cv = (int)(idx - (unsigned)r * (unsigned)nvec);
  const __nv_bfloat16* row = x + r * C;
  float xs[24];
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    const int c = (cv - 1 + v) * 8;
    if (c >= 0 && c < C) unpack8(*reinterpret_cast<const bf16x8*>(row + c), xs + 8 * v);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) xs[8 * v + i] = 0.f;
    }
  }
  float out[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float s = 0.f;
#pragma unroll
    for (int j = -HALF; j <= HALF; ++j) { float t = xs[8 + i + j]; s += t * t; }
    const float scale = k + alpha * s;
    out[i] = xs[8 + i] * exp2f(-beta * __log2f(scale));
  }
  *reinterpret_cast<bf16x8*>(y + r * C + cv * 8) = pack8(out);
}

template <int HALF>
__global__ void lrn_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                               __nv_bfloat16* __restrict__ dx, long long rows, int C, float k, float alpha, float beta) {
  const int nvec = C >> 3;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;           // host guarantees rows * nvec < 2^32
  if (idx >= (unsigned)rows * (unsigned)nvec) return;
  const long long r = idx / (unsigned)nvec;
  const int cv = (int)(idx - (unsigned)r * (unsigned)nvec);
  float xs[24], ds[24];
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    const int c = (cv - 1 + v) * 8;
    if (c >= 0 && c < C) {
      unpack8(*reinterpret_cast<const bf16x8*>(x + r * C + c), xs + 8 * v);
      unpack8(*reinterpret_cast<const bf16x8*>(dy + r * C + c), ds + 8 * v);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { xs[8 * v + i] = 0.f; ds[8 * v + i] = 0.f; }
    }
  }
  // t_i = dy_i * x_i * s_i^(-beta-1) for i in [8-HALF, 16+HALF); p_i = s_i^-beta for the centre 8
  float t[8 + 2 * HALF];
  float pc[8];
#pragma unroll
  for (int a = 0; a < 8 + 2 * HALF; ++a) {
    const int L = 8 - HALF + a;
    float s = 0.f;
#pragma unroll
    for (int j = -HALF; j <= HALF; ++j) { float q = xs[L + j]; s += q * q; }
    const float scale = k + alpha * s;
    const float p = exp2f(-beta * __log2f(scale));
    t[a] = ds[L] * xs[L] * p / scale;
    if (a >= HALF && a < HALF + 8) pc[a - HALF] = p;
  }
  float out[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j <= 2 * HALF; ++j) acc += t[i + j];
    out[i] = ds[8 + i] * pc[i] - 2.f * alpha * beta * xs[8 + i] * acc;
  }
  *reinterpret_cast<bf16x8*>(dx + r * C + cv * 8) = pack8(out);
}

void lrn_fwd(const void* x, void* y, long long rows, int C, int n, float k, float alpha, float beta, cudaStream_t st) {
  if (C % 8) throw std::runtime_error("lrn: C must be a multiple of 8");
  const int half = n / 2;
  long long total = rows * (C / 8);
  if (total >= (1LL << 32)) throw std::runtime_error("lrn: tensor too large for 32-bit indexing");
  const int B = 256;
  auto X = (const __nv_bfloat16*)x; auto Y = (__nv_bfloat16*)y;
  switch (half) {
    case 1: lrn_fwd_kernel<1><<<grid_for(total, B), B, 0, st>>>(X, Y, rows, C, k, alpha, beta); break;
    case 2: lrn_fwd_kernel<2><<<grid_for(total, B), B, 0, st>>>(X, Y, rows, C, k, alpha, beta); break;
    case 3: lrn_fwd_kernel<3><<<grid_for(total, B), B, 0, st>>>(X, Y, rows, C, k, alpha, beta); break;
    case 4: lrn_fwd_kernel<4><<<grid_for(total, B), B, 0, st>>>(X, Y, rows, C, k, alpha, beta); break;
    default: throw std::runtime_error("lrn: window n must be 3,5,7 or 9");
  }
  count_launch(); TMPI_CHECK_LAUNCH("lrn_fwd"); ::tmpi::check_capture(st, "lrn_fwd");
}

void lrn_bwd(const void* x, const void* dy, void* dx, long long rows, int C, int n, float k, float alpha, float beta, cudaStream_t st) {
  if (C % 8) throw std::runtime_error("lrn: C must be a multiple of 8");
  const int half = n / 2;
  long long total = rows * (C / 8);
  if (total >= (1LL << 32)) throw std::runtime_error("lrn: tensor too large for 32-bit indexing");
  const int B = 256;
  auto X = (const __nv_bfloat16*)x; auto DY = (const __nv_bfloat16*)dy; auto DX = (__nv_bfloat16*)dx;
  switch (half) {
    case 1: lrn_bwd_kernel<1><<<grid_for(total, B), B, 0, st>>>(X, DY, DX, rows, C, k, alpha, beta); break;
    case 2: lrn_bwd_kernel<2><<<grid_for(total, B), B, 0, st>>>(X, DY, DX, rows, C, k, alpha, beta); break;
    case 3: lrn_bwd_kernel<3><<<grid_for(total, B), B, 0, st>>>(X, DY, DX, rows, C, k, alpha, beta); break;
    case 4: lrn_bwd_kernel<4><<<grid_for(total, B), B, 0, st>>>(X, DY, DX, rows, C, k, alpha, beta); break;
    default: throw std::runtime_error("lrn: window n must be 3,5,7 or 9");
  }
  count_launch(); TMPI_CHECK_LAUNCH("lrn_bwd"); ::tmpi::check_capture(st, "lrn_bwd");
}

// ============================================================================ pooling
struct PoolGeom { int N, H, W, C, Ho, Wo, k, s, p; };

template <int MAXW, bool FUSE>      // defined with the fused conv→pool backward further down
__global__ void maxpool_relu_bias_bwd_kernel(const __nv_bfloat16* __restrict__ dyp, const uint8_t* __restrict__ arg,
                                             const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ dym,
                                             float* __restrict__ db0, float* __restrict__ db1, int c_split, PoolGeom g, int VT);

// One CTA per output row (n, ho): (n, ho) come from blockIdx, threads sweep (wo, channel vector) — no per-thread
// division chain, 32-bit offsets inside the row (the one-thread-per-output version was instruction bound at 2.5x the
// memory roofline).  K > 0: compile-time window size (unrolled), K == 0: generic.
template <int K>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                          uint8_t* __restrict__ arg, PoolGeom g) {
  const int k = K > 0 ? K : g.k;
  const int nvec = g.C >> 3;
  const int ho = blockIdx.x % g.Ho, n = blockIdx.x / g.Ho;
  const int h0 = ho * g.s - g.p;
  const __nv_bfloat16* xin = x + (long long)n * g.H * g.W * g.C;
  const long long orow = ((long long)n * g.Ho + ho) * g.Wo * g.C;
  const int items = g.Wo * nvec;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int wo = it / nvec, cv = it - wo * nvec;
    const int w0 = wo * g.s - g.p;
    // running max and its window index stay PACKED (two bf16 / two 16-bit indices per register): one __hgt2_mask and two
    // bit-selects per register and tap instead of unpack + compare + two selects per channel
    uint32_t best[4], bidx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { best[i] = 0xFF80FF80u; bidx[i] = 0u; }          // -inf, index 0
#pragma unroll
    for (int kh = 0; kh < (K > 0 ? K : 1); ++kh) {
      for (int kh2 = (K > 0 ? kh : 0); kh2 < (K > 0 ? kh + 1 : k); ++kh2) {
        const int h = h0 + kh2;
        if (h < 0 || h >= g.H) continue;
#pragma unroll
        for (int kw = 0; kw < (K > 0 ? K : 1); ++kw) {
          for (int kw2 = (K > 0 ? kw : 0); kw2 < (K > 0 ? kw + 1 : k); ++kw2) {
            const int w = w0 + kw2;
            if (w < 0 || w >= g.W) continue;
            const bf16x8 v = *reinterpret_cast<const bf16x8*>(xin + ((unsigned)(h * g.W + w) * (unsigned)g.C + cv * 8));
            const uint32_t* vw = reinterpret_cast<const uint32_t*>(&v);
            const uint32_t tt = (uint32_t)(kh2 * k + kw2) * 0x00010001u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&vw[i]), *reinterpret_cast<const __nv_bfloat162*>(&best[i]));
              best[i] = (vw[i] & m) | (best[i] & ~m);
              bidx[i] = (tt & m) | (bidx[i] & ~m);
            }
          }
        }
      }
    }
    const long long o = orow + (unsigned)(wo * g.C + cv * 8);
    *reinterpret_cast<uint4*>(y + o) = make_uint4(best[0], best[1], best[2], best[3]);
    uint2 packed;                                   // 8 window indices, one byte per channel
    packed.x = __byte_perm(bidx[0], bidx[1], 0x6420);
    packed.y = __byte_perm(bidx[2], bidx[3], 0x6420);
    *reinterpret_cast<uint2*>(arg + o) = packed;
  }
}

__global__ void maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ arg,
                                   __nv_bfloat16* __restrict__ dx, PoolGeom g) {
  const int nvec = g.C >> 3;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;             // host guarantees total < 2^32
  const unsigned total = (unsigned)g.N * g.H * g.W * nvec;
  if (idx >= total) return;
  const int cv = (int)(idx % (unsigned)nvec); unsigned t = idx / (unsigned)nvec;
  const int w = (int)(t % g.W); t /= g.W;
  const int h = (int)(t % g.H); const int n = (int)(t / g.H);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  // outputs whose window covers (h, w): ho*s - p <= h < ho*s - p + k
  int ho_lo = (h + g.p - g.k + g.s) / g.s; if (h + g.p - g.k + 1 <= 0) ho_lo = 0;
  int wo_lo = (w + g.p - g.k + g.s) / g.s; if (w + g.p - g.k + 1 <= 0) wo_lo = 0;
  const int ho_hi = min(g.Ho - 1, (h + g.p) / g.s);
  const int wo_hi = min(g.Wo - 1, (w + g.p) / g.s);
  for (int ho = ho_lo; ho <= ho_hi; ++ho) {
    const int kh = h + g.p - ho * g.s;
    if (kh < 0 || kh >= g.k) continue;
    for (int wo = wo_lo; wo <= wo_hi; ++wo) {
      const int kw = w + g.p - wo * g.s;
      if (kw < 0 || kw >= g.k) continue;
      const long long o = (((long long)n * g.Ho + ho) * g.Wo + wo) * g.C + cv * 8;
      const uint2 a = *reinterpret_cast<const uint2*>(arg + o);
      float d[8];
      unpack8(*reinterpret_cast<const bf16x8*>(dy + o), d);
      const uint32_t me = (uint32_t)(kh * g.k + kw);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t ai = ((i < 4 ? a.x : a.y) >> (8 * (i & 3))) & 0xFFu;
        if (ai == me) acc[i] += d[i];
      }
    }
  }
  *reinterpret_cast<bf16x8*>(dx + (((long long)n * g.H + h) * g.W + w) * g.C + cv * 8) = pack8(acc);
}

__device__ __forceinline__ int avg_count(const PoolGeom& g, int ho, int wo) {
  const int h0 = max(0, ho * g.s - g.p), h1 = min(g.H, ho * g.s - g.p + g.k);
  const int w0 = max(0, wo * g.s - g.p), w1 = min(g.W, wo * g.s - g.p + g.k);
  return max(1, (h1 - h0) * (w1 - w0));
}

__global__ void avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, PoolGeom g) {
  const int nvec = g.C >> 3;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;             // host guarantees total < 2^32
  const unsigned total = (unsigned)g.N * g.Ho * g.Wo * nvec;
  if (idx >= total) return;
  const int cv = (int)(idx % (unsigned)nvec); unsigned t = idx / (unsigned)nvec;
  const int wo = (int)(t % g.Wo); t /= g.Wo;
  const int ho = (int)(t % g.Ho); const int n = (int)(t / g.Ho);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int kh = 0; kh < g.k; ++kh) {
    const int h = ho * g.s - g.p + kh;
    if (h < 0 || h >= g.H) continue;
    for (int kw = 0; kw < g.k; ++kw) {
      const int w = wo * g.s - g.p + kw;
      if (w < 0 || w >= g.W) continue;
      float v[8];
      unpack8(*reinterpret_cast<const bf16x8*>(x + (((long long)n * g.H + h) * g.W + w) * g.C + cv * 8), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
  const float inv = 1.f / (float)avg_count(g, ho, wo);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] *= inv;
  *reinterpret_cast<bf16x8*>(y + (((long long)n * g.Ho + ho) * g.Wo + wo) * g.C + cv * 8) = pack8(acc);
}

__global__ void avgpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, PoolGeom g) {
  const int nvec = g.C >> 3;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;             // host guarantees total < 2^32
  const unsigned total = (unsigned)g.N * g.H * g.W * nvec;
  if (idx >= total) return;
  const int cv = (int)(idx % (unsigned)nvec); unsigned t = idx / (unsigned)nvec;
  const int w = (int)(t % g.W); t /= g.W;
  const int h = (int)(t % g.H); const int n = (int)(t / g.H);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const int ho_hi = min(g.Ho - 1, (h + g.p) / g.s);
  const int wo_hi = min(g.Wo - 1, (w + g.p) / g.s);
  for (int ho = 0; ho <= ho_hi; ++ho) {
    const int kh = h + g.p - ho * g.s;
    if (kh < 0 || kh >= g.k) continue;
    for (int wo = 0; wo <= wo_hi; ++wo) {
      const int kw = w + g.p - wo * g.s;
      if (kw < 0 || kw >= g.k) continue;
      float d[8];
      unpack8(*reinterpret_cast<const bf16x8*>(dy + (((long long)n * g.Ho + ho) * g.Wo + wo) * g.C + cv * 8), d);
      const float inv = 1.f / (float)avg_count(g, ho, wo);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += d[i] * inv;
    }
  }
  *reinterpret_cast<bf16x8*>(dx + (((long long)n * g.H + h) * g.W + w) * g.C + cv * 8) = pack8(acc);
}

void pool_fwd(const void* x, void* y, void* arg, int N, int H, int W, int C, int Ho, int Wo, int k, int s, int p, int is_max, cudaStream_t st) {
  if (C % 8) throw std::runtime_error("pool: C must be a multiple of 8");
  PoolGeom g{N, H, W, C, Ho, Wo, k, s, p};
  long long total = (long long)N * Ho * Wo * (C / 8);
  if ((long long)N * H * W * (C / 8) >= (1LL << 32)) throw std::runtime_error("pool: tensor too large for 32-bit indexing");
  if ((long long)H * W * C >= (1LL << 31)) throw std::runtime_error("pool: image too large for 32-bit in-image offsets");
  if (is_max) {
    const unsigned rows = (unsigned)N * Ho;
    auto X = (const __nv_bfloat16*)x; auto Y = (__nv_bfloat16*)y; auto A = (uint8_t*)arg;
    if (k == 3) maxpool_fwd_kernel<3><<<rows, 256, 0, st>>>(X, Y, A, g);
    else if (k == 2) maxpool_fwd_kernel<2><<<rows, 256, 0, st>>>(X, Y, A, g);
    else maxpool_fwd_kernel<0><<<rows, 256, 0, st>>>(X, Y, A, g);
  } else avgpool_fwd_kernel<<<grid_for(total, 256), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, g);
  count_launch(); TMPI_CHECK_LAUNCH("pool_fwd"); ::tmpi::check_capture(st, "pool_fwd");
}

void pool_bwd(const void* dy, const void* arg, void* dx, int N, int H, int W, int C, int Ho, int Wo, int k, int s, int p, int is_max, cudaStream_t st) {
  PoolGeom g{N, H, W, C, Ho, Wo, k, s, p};
  long long total = (long long)N * H * W * (C / 8);
  const int maxw = (k + s - 1) / s;
  if (is_max && maxw <= 3 && (long long)H * W * C < (1LL << 31)) {
    // row-per-CTA kernel shared with the fused conv→pool backward (all candidate windows in flight at once)
    const int nvec = C / 8;
    const int VT = nvec < 32 ? nvec : 32;
    dim3 grid((unsigned)N * H, (unsigned)((nvec + VT - 1) / VT));
    auto DY = (const __nv_bfloat16*)dy; auto A = (const uint8_t*)arg; auto DX = (__nv_bfloat16*)dx;
    if (maxw == 1) maxpool_relu_bias_bwd_kernel<1, false><<<grid, 256, 0, st>>>(DY, A, nullptr, DX, nullptr, nullptr, C, g, VT);
    else if (maxw == 2) maxpool_relu_bias_bwd_kernel<2, false><<<grid, 256, 0, st>>>(DY, A, nullptr, DX, nullptr, nullptr, C, g, VT);
    else maxpool_relu_bias_bwd_kernel<3, false><<<grid, 256, 0, st>>>(DY, A, nullptr, DX, nullptr, nullptr, C, g, VT);
  } else if (is_max) maxpool_bwd_kernel<<<grid_for(total, 256), 256, 0, st>>>((const __nv_bfloat16*)dy, (const uint8_t*)arg, (__nv_bfloat16*)dx, g);
  else avgpool_bwd_kernel<<<grid_for(total, 256), 256, 0, st>>>((const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, g);
  count_launch(); TMPI_CHECK_LAUNCH("pool_bwd"); ::tmpi::check_capture(st, "pool_bwd");
}

// ============================================================================ dropout (Philox4x32-10)
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// y = x * keep; keep drawn with P(keep) = 1 - p_drop from Philox keyed by (seed, layer) and counter (idx, *step)
__global__ void dropout_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ mask,
                                   long long n8, float p_drop, unsigned long long seed, uint32_t layer,
                                   const unsigned long long* __restrict__ step) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n8) return;
  const unsigned long long stp = *step;
  uint32_t r[4];
  philox4x32((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)stp, (uint32_t)(stp >> 32), (uint32_t)seed,
             (uint32_t)(seed >> 32) ^ (layer * 0x9E3779B9u), r);
  const uint32_t thr = (uint32_t)(p_drop * 65536.f);
  float v[8];
  unpack8(*reinterpret_cast<const bf16x8*>(x + idx * 8), v);
  uint32_t mlo = 0, mhi = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t u = (r[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
    const uint32_t keep = u >= thr ? 1u : 0u;
    v[i] = keep ? v[i] : 0.f;
    if (i < 4) mlo |= keep << (8 * i); else mhi |= keep << (8 * (i - 4));
  }
  *reinterpret_cast<bf16x8*>(y + idx * 8) = pack8(v);
  *reinterpret_cast<uint2*>(mask + idx * 8) = make_uint2(mlo, mhi);
}

__global__ void dropout_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ mask,
                                   __nv_bfloat16* __restrict__ dx, long long n8) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n8) return;
  float v[8];
  unpack8(*reinterpret_cast<const bf16x8*>(dy + idx * 8), v);
  const uint2 m = *reinterpret_cast<const uint2*>(mask + idx * 8);
#pragma unroll
  for (int i = 0; i < 8; ++i) { const uint32_t k = ((i < 4 ? m.x : m.y) >> (8 * (i & 3))) & 0xFFu; if (!k) v[i] = 0.f; }
  *reinterpret_cast<bf16x8*>(dx + idx * 8) = pack8(v);
}

__global__ void advance_step_kernel(unsigned long long* step) { if (threadIdx.x == 0 && blockIdx.x == 0) *step += 1ull; }

void dropout_fwd(const void* x, void* y, void* mask, long long n, float p_drop, unsigned long long seed, int layer, const void* step, cudaStream_t st) {
  if (n % 8) throw std::runtime_error("dropout: numel must be a multiple of 8");
  dropout_fwd_kernel<<<grid_for(n / 8, 256), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, (uint8_t*)mask, n / 8, p_drop, seed,
                                                           (uint32_t)layer, (const unsigned long long*)step);
  count_launch(); TMPI_CHECK_LAUNCH("dropout_fwd"); ::tmpi::check_capture(st, "dropout_fwd");
}
void dropout_bwd(const void* dy, const void* mask, void* dx, long long n, cudaStream_t st) {
  dropout_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, st>>>((const __nv_bfloat16*)dy, (const uint8_t*)mask, (__nv_bfloat16*)dx, n / 8);
  count_launch(); TMPI_CHECK_LAUNCH("dropout_bwd"); ::tmpi::check_capture(st, "dropout_bwd");
}
void advance_step(void* step, cudaStream_t st) {
  advance_step_kernel<<<1, 32, 0, st>>>((unsigned long long*)step);
  count_launch(); TMPI_CHECK_LAUNCH("advance_step"); ::tmpi::check_capture(st, "advance_step");
}

// ============================================================================ softmax + NLL + errors + dlogits
// one CTA per row; rowstat[b] = {nll, err1, err5}; dlogits = (softmax - onehot) * scale
__global__ void softmax_xent_kernel(const __nv_bfloat16* __restrict__ logits, const long long* __restrict__ labels,
                                    __nv_bfloat16* __restrict__ dlogits, float* __restrict__ rowstat, int C, float scale) {
  const int b = blockIdx.x;
  const __nv_bfloat16* row = logits + (long long)b * C;
  const int label = (int)labels[b];
  __shared__ float red[32];
  __shared__ float bcast;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, bf16_to_f(row[c]));
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (warp == 0) { float v = lane < nw ? red[lane] : -INFINITY; v = warp_max(v); if (lane == 0) bcast = v; }
  __syncthreads();
  mx = bcast;
  __syncthreads();
  const float lab = bf16_to_f(row[label]);
  float se = 0.f, gt = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = bf16_to_f(row[c]);
    se += __expf(v - mx);
    gt += (v > lab || (v == lab && c < label)) ? 1.f : 0.f;     // rank of the label's logit
  }
  se = warp_sum(se); gt = warp_sum(gt);
  if (lane == 0) { red[warp] = se; }
  __syncthreads();
  if (warp == 0) { float v = lane < nw ? red[lane] : 0.f; v = warp_sum(v); if (lane == 0) bcast = v; }
  __syncthreads();
  se = bcast;
  __syncthreads();
  if (lane == 0) { red[warp] = gt; }
  __syncthreads();
  if (warp == 0) { float v = lane < nw ? red[lane] : 0.f; v = warp_sum(v); if (lane == 0) bcast = v; }
  __syncthreads();
  gt = bcast;
  const float inv = 1.f / se;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float pr = __expf(bf16_to_f(row[c]) - mx) * inv;
    if (c == label) pr -= 1.f;
    dlogits[(long long)b * C + c] = f_to_bf16(pr * scale);
  }
  if (threadIdx.x == 0) {
    rowstat[3 * b + 0] = -(lab - mx - __logf(se));
    rowstat[3 * b + 1] = gt >= 1.f ? 1.f : 0.f;
    rowstat[3 * b + 2] = gt >= 5.f ? 1.f : 0.f;
  }
}

__global__ void rowstat_mean_kernel(const float* __restrict__ rowstat, float* __restrict__ out, int B, float weight) {
  __shared__ float red[3][32];
  float a = 0.f, b = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) { a += rowstat[3 * i]; b += rowstat[3 * i + 1]; c += rowstat[3 * i + 2]; }
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) { red[0][warp] = a; red[1][warp] = b; red[2][warp] = c; }
  __syncthreads();
  if (warp == 0) {
    a = lane < nw ? red[0][lane] : 0.f; b = lane < nw ? red[1][lane] : 0.f; c = lane < nw ? red[2][lane] : 0.f;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) { out[0] = weight * a / B; out[1] = b / B; out[2] = c / B; }
  }
}

void softmax_xent(const void* logits, const void* labels, void* dlogits, void* rowstat, void* out3, int B, int C, float weight, cudaStream_t st) {
  softmax_xent_kernel<<<B, 256, 0, st>>>((const __nv_bfloat16*)logits, (const long long*)labels, (__nv_bfloat16*)dlogits, (float*)rowstat, C, weight / (float)B);
  count_launch(); TMPI_CHECK_LAUNCH("softmax_xent"); ::tmpi::check_capture(st, "softmax_xent");
  rowstat_mean_kernel<<<1, 256, 0, st>>>((const float*)rowstat, (float*)out3, B, weight);
  count_launch(); TMPI_CHECK_LAUNCH("rowstat_mean"); ::tmpi::check_capture(st, "rowstat_mean");
}

// ============================================================================ ReLU mask + bias gradient
// dym = dy * (y > 0) (bf16, contiguous [R, C]);  db[c] += sum_r dym[r, c]   (db pre-zeroed by the launcher)
// dy / y have row pitch ld (elements) so channel slices of a wider tensor work (grouped conv).
template <bool RELU, bool WRITE>
__global__ void relu_bias_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                     __nv_bfloat16* __restrict__ dym, float* __restrict__ db, float* __restrict__ db1, int c_split,
                                     long long R, int C, long long ld, int VT, int rows_per_cta) {
  extern __shared__ float sm[];                       // [RL][VT*8]
  const int nvec = C >> 3;
  const int RL = blockDim.x / VT;
  const int tv = threadIdx.x % VT, tr = threadIdx.x / VT;
  const int cv = blockIdx.y * VT + tv;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (tr < RL && cv < nvec) {
    const long long rend = min(R, r0 + rows_per_cta);
    for (long long r = r0 + tr; r < rend; r += 4 * RL) {
      // 4 rows per trip, all loads issued before any use (memory-level parallelism: this kernel is pure streaming)
      bf16x8 dv[4], yv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long rr = r + (long long)u * RL;
        if (rr < rend) {
          dv[u] = *reinterpret_cast<const bf16x8*>(dy + rr * ld + cv * 8);
          if (RELU) yv[u] = *reinterpret_cast<const bf16x8*>(y + rr * ld + cv * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long rr = r + (long long)u * RL;
        if (rr < rend) {
          float d[8];
          unpack8(dv[u], d);
          if (RELU) {
            float v[8];
            unpack8(yv[u], v);
#pragma unroll
            for (int i = 0; i < 8; ++i) if (!(v[i] > 0.f)) d[i] = 0.f;
          }
          if (WRITE) *reinterpret_cast<bf16x8*>(dym + rr * C + cv * 8) = pack8(d);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += d[i];
        }
      }
    }
  }
  if (db == nullptr) return;
  if (tr < RL) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[(tr * VT + tv) * 8 + i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < VT * 8) {
    const int v = threadIdx.x / 8, i = threadIdx.x % 8;
    const int c = (blockIdx.y * VT + v) * 8 + i;
    if (c < C) {
      float s = 0.f;
      for (int t = 0; t < RL; ++t) s += sm[(t * VT + v) * 8 + i];
      atomicAdd(c < c_split ? db + c : db1 + (c - c_split), s);
    }
  }
}

// db: bias gradient for channels [0, c_split), db1: for channels [c_split, C) (the two parameter sets of a 2-group block);
// pass db1 = nullptr / c_split = C for a single bias vector.
void relu_bias_bwd2(const void* dy, const void* y, void* dym, void* db, void* db1, int c_split, long long R, int C, long long ld, int relu,
                    cudaStream_t st) {
  if (C % 8) throw std::runtime_error("relu_bias_bwd: C must be a multiple of 8");
  const int nvec = C / 8;
  const int VT = nvec < 32 ? nvec : 32;
  const int RL = 256 / VT;
  // deterministic mode: one CTA per channel group sums ALL rows (fixed order) instead of row slabs + atomics
  const long long rows_per_cta_ll = deterministic_mode() ? std::max<long long>(R, 1) : (long long)RL * 8;
  if (rows_per_cta_ll >= (1LL << 31)) throw std::runtime_error("relu_bias_bwd: too many rows for the deterministic mode");
  const int rows_per_cta = (int)rows_per_cta_ll;
  dim3 grid((unsigned)((R + rows_per_cta - 1) / rows_per_cta), (unsigned)((nvec + VT - 1) / VT));
  const size_t smem = (size_t)RL * VT * 8 * sizeof(float);
  if (!db1 || c_split > C) c_split = C;
  if (db) check_cuda(cudaMemsetAsync(db, 0, (size_t)c_split * 4, st), "relu_bias_bwd memset");
  if (db && c_split < C) check_cuda(cudaMemsetAsync(db1, 0, (size_t)(C - c_split) * 4, st), "relu_bias_bwd memset");
  const bool write = dym != nullptr;
  auto DY = (const __nv_bfloat16*)dy; auto Y = (const __nv_bfloat16*)y; auto DM = (__nv_bfloat16*)dym; auto DB = (float*)db; auto DB1 = (float*)db1;
  if (relu && write) relu_bias_bwd_kernel<true, true><<<grid, 256, smem, st>>>(DY, Y, DM, DB, DB1, c_split, R, C, ld, VT, rows_per_cta);
  else if (relu) relu_bias_bwd_kernel<true, false><<<grid, 256, smem, st>>>(DY, Y, DM, DB, DB1, c_split, R, C, ld, VT, rows_per_cta);
  else if (write) relu_bias_bwd_kernel<false, true><<<grid, 256, smem, st>>>(DY, Y, DM, DB, DB1, c_split, R, C, ld, VT, rows_per_cta);
  else relu_bias_bwd_kernel<false, false><<<grid, 256, smem, st>>>(DY, Y, DM, DB, DB1, c_split, R, C, ld, VT, rows_per_cta);
  count_launch(); TMPI_CHECK_LAUNCH("relu_bias_bwd"); ::tmpi::check_capture(st, "relu_bias_bwd");
}
void relu_bias_bwd(const void* dy, const void* y, void* dym, void* db, long long R, int C, long long ld, int relu, cudaStream_t st) {
  relu_bias_bwd2(dy, y, dym, db, nullptr, C, R, C, ld, relu, st);
}

// y[r, c] (bf16) = act(acc[r, c] (fp32) + bias[c]) — finishing pass of a split-K forward GEMM (small-batch FC layers: the
// parallelism has to come from splitting K, and split-K accumulates in fp32 with reductions, so bias / ReLU / cast run here)
__global__ void bias_act_cast_kernel(const float* __restrict__ acc, const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                                     int R, int C, int relu) {
  const int nvec = C >> 3;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)R * (unsigned)nvec) return;
  const unsigned r = idx / (unsigned)nvec; const int cv = (int)(idx - r * (unsigned)nvec);
  const float4 a0 = *reinterpret_cast<const float4*>(acc + (size_t)r * C + cv * 8);
  const float4 a1 = *reinterpret_cast<const float4*>(acc + (size_t)r * C + cv * 8 + 4);
  float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  if (bias) {
    const float4 b0 = *reinterpret_cast<const float4*>(bias + cv * 8), b1 = *reinterpret_cast<const float4*>(bias + cv * 8 + 4);
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  if (relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  *reinterpret_cast<bf16x8*>(y + (size_t)r * C + cv * 8) = pack8(v);
}
void bias_act_cast(const void* acc, const void* bias, void* y, int R, int C, int relu, cudaStream_t st) {
  if (C % 8) throw std::runtime_error("bias_act_cast: C must be a multiple of 8");
  const long long total = (long long)R * (C / 8);
  bias_act_cast_kernel<<<grid_for(total, 256), 256, 0, st>>>((const float*)acc, (const float*)bias, (__nv_bfloat16*)y, R, C, relu);
  count_launch(); TMPI_CHECK_LAUNCH("bias_act_cast"); ::tmpi::check_capture(st, "bias_act_cast");
}

// Fused backward of  conv(+bias+ReLU) -> max-pool : one pass over the conv output instead of three.
//   dym[n,h,w,c] = (sum over pooling windows whose argmax is (h,w) of dyp) * (y[n,h,w,c] > 0)     (bf16, contiguous)
//   db[c]       += sum_{n,h,w} dym                                                                 (pre-zeroed by the launcher)
// Channels < c_split accumulate into db0, the rest into db1 (the two parameter sets of a 2-group AlexNet block).
// Replaces maxpool_bwd (write dx) + relu_bias_bwd (read dx, read y, write dym): 2 reads + 2 writes of the big tensor -> 1 + 1.
// One CTA per input row (n, h): the candidate output rows are CTA-uniform, each thread owns one channel vector (so its bias
// partial sums stay in registers) and sweeps w.
template <int MAXW, bool FUSE>
__global__ void __launch_bounds__(256) maxpool_relu_bias_bwd_kernel(const __nv_bfloat16* __restrict__ dyp, const uint8_t* __restrict__ arg,
                                             const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ dym,
                                             float* __restrict__ db0, float* __restrict__ db1, int c_split, PoolGeom g, int VT) {
  extern __shared__ float sm[];                       // [RL][VT*8]
  const int nvec = g.C >> 3;
  const int RL = blockDim.x / VT;
  const int tv = threadIdx.x % VT, tw = threadIdx.x / VT;
  const int cv = blockIdx.y * VT + tv;
  const int h = blockIdx.x % g.H, n = blockIdx.x / g.H;
  int ho_lo = (h + g.p - g.k + g.s) / g.s; if (h + g.p - g.k + 1 <= 0) ho_lo = 0;
  const int ho_hi = min(g.Ho - 1, (h + g.p) / g.s);
  const long long irow = ((long long)n * g.H + h) * g.W * g.C;
  const long long obase = (long long)n * g.Ho * g.Wo * g.C;
  float tot[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) tot[i] = 0.f;
  if (tw < RL && cv < nvec) {
    for (int w = tw; w < g.W; w += RL) {
      int wo_lo = (w + g.p - g.k + g.s) / g.s; if (w + g.p - g.k + 1 <= 0) wo_lo = 0;
      const int wo_hi = min(g.Wo - 1, (w + g.p) / g.s);
      // phase 1: all candidate windows (at most MAXW x MAXW = ceil(k/s)^2) and the activation in flight at once
      uint2 av[MAXW * MAXW];
      bf16x8 dv[MAXW * MAXW];
      uint32_t me[MAXW * MAXW];
      bool ok[MAXW * MAXW];
#pragma unroll
      for (int a = 0; a < MAXW; ++a) {
#pragma unroll
        for (int b = 0; b < MAXW; ++b) {
          const int ho = ho_lo + a, wo = wo_lo + b;
          const int t = a * MAXW + b;
          ok[t] = (ho <= ho_hi && wo <= wo_hi);
          me[t] = (uint32_t)((h + g.p - ho * g.s) * g.k + (w + g.p - wo * g.s));
          if (ok[t]) {
            const long long o = obase + (unsigned)((ho * g.Wo + wo) * g.C + cv * 8);
            av[t] = *reinterpret_cast<const uint2*>(arg + o);
            dv[t] = *reinterpret_cast<const bf16x8*>(dyp + o);
          }
        }
      }
      const long long xi = irow + (unsigned)(w * g.C + cv * 8);
      bf16x8 yv;
      if (FUSE) yv = *reinterpret_cast<const bf16x8*>(y + xi);
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
      for (int t = 0; t < MAXW * MAXW; ++t) {
        if (ok[t]) {
          // argmax match for 4 channels per instruction (__vcmpeq4 on the packed argmax bytes), byte mask widened to the
          // bf16 lanes with PRMT, gradient kept packed until the add
          const uint32_t me4 = me[t] * 0x01010101u;
          const uint32_t eq_lo = __vcmpeq4(av[t].x, me4), eq_hi = __vcmpeq4(av[t].y, me4);
          const uint32_t* dw = reinterpret_cast<const uint32_t*>(&dv[t]);
          const uint32_t w0 = dw[0] & __byte_perm(eq_lo, 0, 0x1100), w1 = dw[1] & __byte_perm(eq_lo, 0, 0x3322);
          const uint32_t w2 = dw[2] & __byte_perm(eq_hi, 0, 0x1100), w3 = dw[3] & __byte_perm(eq_hi, 0, 0x3322);
          acc[0] += __uint_as_float(w0 << 16); acc[1] += __uint_as_float(w0 & 0xFFFF0000u);
          acc[2] += __uint_as_float(w1 << 16); acc[3] += __uint_as_float(w1 & 0xFFFF0000u);
          acc[4] += __uint_as_float(w2 << 16); acc[5] += __uint_as_float(w2 & 0xFFFF0000u);
          acc[6] += __uint_as_float(w3 << 16); acc[7] += __uint_as_float(w3 & 0xFFFF0000u);
        }
      }
      bf16x8 pk = pack8(acc);
      if (FUSE) {
        // ReLU mask on the packed result: a bf16 is > 0 exactly when its bits, read as int16, are > 0
        uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
        const uint32_t* yw = reinterpret_cast<const uint32_t*>(&yv);
#pragma unroll
        for (int i = 0; i < 4; ++i) pw[i] &= __vcmpgts2(yw[i], 0u);
      }
      *reinterpret_cast<bf16x8*>(dym + xi) = pk;
      if (FUSE) {
        unpack8(pk, acc);                               // db sums what wgrad / dgrad will actually see (bf16-rounded)
#pragma unroll
        for (int i = 0; i < 8; ++i) tot[i] += acc[i];
      }
    }
  }
  if (!FUSE) return;                                    // plain max-pool backward: no mask, no bias gradient
  if (tw < RL) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[(tw * VT + tv) * 8 + i] = tot[i];
  }
  __syncthreads();
  if (threadIdx.x < VT * 8) {
    const int v = threadIdx.x / 8, i = threadIdx.x % 8;
    const int c = (blockIdx.y * VT + v) * 8 + i;
    if (c < g.C) {
      float sacc = 0.f;
      for (int t = 0; t < RL; ++t) sacc += sm[(t * VT + v) * 8 + i];
      atomicAdd(c < c_split ? db0 + c : db1 + (c - c_split), sacc);
    }
  }
}

void maxpool_relu_bias_bwd(const void* dyp, const void* arg, const void* y, void* dym, void* db0, void* db1, int c_split, int N, int H,
                           int W, int C, int Ho, int Wo, int k, int s, int p, cudaStream_t st) {
  if (C % 8) throw std::runtime_error("maxpool_relu_bias_bwd: C must be a multiple of 8");
  if ((long long)H * W * C >= (1LL << 31)) throw std::runtime_error("maxpool_relu_bias_bwd: image too large for 32-bit in-image offsets");
  PoolGeom g{N, H, W, C, Ho, Wo, k, s, p};
  const int nvec = C / 8;
  const int VT = nvec < 32 ? nvec : 32;
  const int RL = 256 / VT;
  dim3 grid((unsigned)N * H, (unsigned)((nvec + VT - 1) / VT));
  const size_t smem = (size_t)RL * VT * 8 * sizeof(float);
  if (c_split > C) c_split = C;
  check_cuda(cudaMemsetAsync(db0, 0, (size_t)c_split * 4, st), "maxpool_relu_bias_bwd memset");
  if (c_split < C) check_cuda(cudaMemsetAsync(db1, 0, (size_t)(C - c_split) * 4, st), "maxpool_relu_bias_bwd memset");
  const int maxw = (k + s - 1) / s;
  auto DYP = (const __nv_bfloat16*)dyp; auto A = (const uint8_t*)arg; auto Y = (const __nv_bfloat16*)y; auto DM = (__nv_bfloat16*)dym;
  if (maxw == 1) maxpool_relu_bias_bwd_kernel<1, true><<<grid, 256, smem, st>>>(DYP, A, Y, DM, (float*)db0, (float*)db1, c_split, g, VT);
  else if (maxw == 2) maxpool_relu_bias_bwd_kernel<2, true><<<grid, 256, smem, st>>>(DYP, A, Y, DM, (float*)db0, (float*)db1, c_split, g, VT);
  else if (maxw == 3) maxpool_relu_bias_bwd_kernel<3, true><<<grid, 256, smem, st>>>(DYP, A, Y, DM, (float*)db0, (float*)db1, c_split, g, VT);
  else throw std::runtime_error("maxpool_relu_bias_bwd: pooling windows overlapping more than 3x3 outputs are not supported");
  count_launch(); TMPI_CHECK_LAUNCH("maxpool_relu_bias_bwd"); ::tmpi::check_capture(st, "maxpool_relu_bias_bwd");
}

// ============================================================================ im2col / col2im (NHWC, bf16)
struct ConvGeom { int N, H, W, Ctot, c_off, Cg, KH, KW, Ho, Wo, s, p; long long ldcol; int K; };

// col[m, (kh*KW+kw)*Cg + c] = x[n, ho*s-p+kh, wo*s-p+kw, c_off+c]   (zero outside the image)
__global__ void __launch_bounds__(256) im2col_vec8_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col, ConvGeom g) {
  // one CTA per IM2COL_ROWS consecutive output pixels; (n, ho, wo) is decoded once per row, the threads sweep the
  // row's KH*KW*(Cg/8) 16-byte vectors with 32-bit index math only.
  constexpr int IM2COL_ROWS = 8;
  const int cvn = g.Cg >> 3;
  const int per_m = g.KH * g.KW * cvn;
  const long long M = (long long)g.N * g.Ho * g.Wo;
  const long long m_base = (long long)blockIdx.x * IM2COL_ROWS;
  for (int rr = 0; rr < IM2COL_ROWS; ++rr) {
    const long long m = m_base + rr;
    if (m >= M) return;
    const int wo = (int)(m % g.Wo); const long long t = m / g.Wo;
    const int ho = (int)(t % g.Ho); const int n = (int)(t / g.Ho);
    const int h0 = ho * g.s - g.p, w0 = wo * g.s - g.p;
    const __nv_bfloat16* xin = x + (long long)n * g.H * g.W * g.Ctot + g.c_off;
    __nv_bfloat16* dst = col + m * g.ldcol;
    for (int v = threadIdx.x; v < per_m; v += blockDim.x) {
      const int kk = v / cvn, cv = v - kk * cvn;
      const int kh = kk / g.KW, kw = kk - kh * g.KW;
      const int h = h0 + kh, w = w0 + kw;
      bf16x8 val;
      if (h >= 0 && h < g.H && w >= 0 && w < g.W)
        val = *reinterpret_cast<const bf16x8*>(xin + ((long long)h * g.W + w) * g.Ctot + cv * 8);
      else {
#pragma unroll
        for (int i = 0; i < 4; ++i) val.v[i] = __floats2bfloat162_rn(0.f, 0.f);
      }
      *reinterpret_cast<bf16x8*>(dst + kk * g.Cg + cv * 8) = val;
    }
  }
}

// generic (any Cg): one thread per (m, kh*KW+kw); the extra index KH*KW zero-fills the K..ldcol padding
__global__ void im2col_scalar_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col, ConvGeom g) {
  const int per_m = g.KH * g.KW + 1;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long M = (long long)g.N * g.Ho * g.Wo;
  if (idx >= M * per_m) return;
  const long long m = idx / per_m; const int kk = (int)(idx % per_m);
  __nv_bfloat16* dst = col + m * g.ldcol;
  if (kk == g.KH * g.KW) { for (int c = g.K; c < g.ldcol; ++c) dst[c] = f_to_bf16(0.f); return; }
  const int kh = kk / g.KW, kw = kk % g.KW;
  const int wo = (int)(m % g.Wo); long long t = m / g.Wo;
  const int ho = (int)(t % g.Ho); const int n = (int)(t / g.Ho);
  const int h = ho * g.s - g.p + kh, w = wo * g.s - g.p + kw;
  const bool ok = (h >= 0 && h < g.H && w >= 0 && w < g.W);
  const __nv_bfloat16* src = x + (((long long)n * g.H + h) * g.W + w) * g.Ctot + g.c_off;
  dst += (long long)kk * g.Cg;
  for (int c = 0; c < g.Cg; ++c) dst[c] = ok ? src[c] : f_to_bf16(0.f);
}

// small-C path (conv1: C = 3): one CTA per (image, output row).  The KH input rows the output row needs are staged in
// shared memory with coalesced loads; the CTA then emits its Wo consecutive col rows (one contiguous Wo*ldcol span)
// as 16-byte vectors — each col row is KH runs of KW*Cg contiguous input elements.
__global__ void __launch_bounds__(256) im2col_rows_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col, ConvGeom g) {
  extern __shared__ __nv_bfloat16 rows[];                  // [KH][W*Cg]
  const int n = blockIdx.x / g.Ho, ho = blockIdx.x % g.Ho;
  const int rowlen = g.W * g.Cg;
  const bool dense = (g.Cg == g.Ctot);                     // all channels used: an input row is one contiguous span
  for (int kh = 0; kh < g.KH; ++kh) {
    const int h = ho * g.s - g.p + kh;
    __nv_bfloat16* dst = rows + kh * rowlen;
    if (h < 0 || h >= g.H) {
      for (int e = threadIdx.x; e < rowlen; e += blockDim.x) dst[e] = f_to_bf16(0.f);
    } else if (dense) {
      const __nv_bfloat16* src = x + ((long long)n * g.H + h) * g.W * g.Ctot;
      for (int e = threadIdx.x; e < rowlen; e += blockDim.x) dst[e] = src[e];
    } else {
      const __nv_bfloat16* src = x + ((long long)n * g.H + h) * g.W * g.Ctot + g.c_off;
      for (int e = threadIdx.x; e < rowlen; e += blockDim.x) { const int w = e / g.Cg; dst[e] = src[(long long)w * g.Ctot + (e - w * g.Cg)]; }
    }
  }
  __syncthreads();
  const int run = g.KW * g.Cg;                             // contiguous elements per (kh)
  const int vec_per_row = (int)(g.ldcol >> 3);
  __nv_bfloat16* out = col + ((long long)n * g.Ho + ho) * g.Wo * g.ldcol;
  for (int i = threadIdx.x; i < g.Wo * vec_per_row; i += blockDim.x) {
    const int wo = i / vec_per_row, e0 = (i - wo * vec_per_row) * 8;
    const int xbase = (wo * g.s - g.p) * g.Cg;              // may be negative with padding
    int kh = e0 / run, r = e0 - kh * run;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __nv_bfloat16 t = f_to_bf16(0.f);
      if (e0 + j < g.K) {
        const int xi = xbase + r;
        if (xi >= 0 && xi < rowlen) t = rows[kh * rowlen + xi];
      }
      v[j] = t;
      if (++r == run) { r = 0; ++kh; }
    }
    *reinterpret_cast<uint4*>(out + (long long)wo * g.ldcol + e0) = *reinterpret_cast<const uint4*>(v);
  }
}

// dx[n,h,w,c_off+c] = sum over (kh,kw) with (h+p-kh)%s==0, (w+p-kw)%s==0 of dcol[m(n,ho,wo), (kh*KW+kw)*Cg + c]
__global__ void col2im_vec8_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ dx, ConvGeom g) {
  const unsigned cvn = (unsigned)(g.Cg >> 3);
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;           // host guarantees total < 2^32
  const unsigned total = (unsigned)g.N * g.H * g.W * cvn;
  if (idx >= total) return;
  const int cv = (int)(idx % cvn); unsigned t = idx / cvn;
  const int w = (int)(t % g.W); t /= g.W;
  const int h = (int)(t % g.H); const int n = (int)(t / g.H);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int kh = 0; kh < g.KH; ++kh) {
    const int th = h + g.p - kh;
    if (th < 0 || th % g.s) continue;
    const int ho = th / g.s;
    if (ho >= g.Ho) continue;
    for (int kw = 0; kw < g.KW; ++kw) {
      const int tw = w + g.p - kw;
      if (tw < 0 || tw % g.s) continue;
      const int wo = tw / g.s;
      if (wo >= g.Wo) continue;
      const long long m = ((long long)n * g.Ho + ho) * g.Wo + wo;
      float v[8];
      unpack8(*reinterpret_cast<const bf16x8*>(dcol + m * g.ldcol + (long long)(kh * g.KW + kw) * g.Cg + cv * 8), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
  *reinterpret_cast<bf16x8*>(dx + (((long long)n * g.H + h) * g.W + w) * g.Ctot + g.c_off + cv * 8) = pack8(acc);
}

void im2col(const void* x, void* col, int N, int H, int W, int Ctot, int c_off, int Cg, int KH, int KW, int Ho, int Wo, int s, int p,
            long long ldcol, cudaStream_t st) {
  ConvGeom g{N, H, W, Ctot, c_off, Cg, KH, KW, Ho, Wo, s, p, ldcol, KH * KW * Cg};
  long long M = (long long)N * Ho * Wo;
  if (Cg % 8 == 0 && c_off % 8 == 0 && Ctot % 8 == 0 && ldcol % 8 == 0) {
    im2col_vec8_kernel<<<grid_for(M, 8), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)col, g);
  } else if (ldcol % 8 == 0 && (size_t)KH * W * Cg * 2 <= 48 * 1024) {
    im2col_rows_kernel<<<N * Ho, 256, (size_t)KH * W * Cg * 2, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)col, g);
  } else {
    long long total = M * (KH * KW + 1);
    im2col_scalar_kernel<<<grid_for(total, 256), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)col, g);
  }
  count_launch(); TMPI_CHECK_LAUNCH("im2col"); ::tmpi::check_capture(st, "im2col");
}

void col2im(const void* dcol, void* dx, int N, int H, int W, int Ctot, int c_off, int Cg, int KH, int KW, int Ho, int Wo, int s, int p,
            long long ldcol, cudaStream_t st) {
  if (Cg % 8 || c_off % 8 || Ctot % 8 || ldcol % 8) throw std::runtime_error("col2im: channel counts must be multiples of 8");
  ConvGeom g{N, H, W, Ctot, c_off, Cg, KH, KW, Ho, Wo, s, p, ldcol, KH * KW * Cg};
  long long total = (long long)N * H * W * (Cg / 8);
  if (total >= (1LL << 32)) throw std::runtime_error("col2im: tensor too large for 32-bit indexing");
  col2im_vec8_kernel<<<grid_for(total, 256), 256, 0, st>>>((const __nv_bfloat16*)dcol, (__nv_bfloat16*)dx, g);
  count_launch(); TMPI_CHECK_LAUNCH("col2im"); ::tmpi::check_capture(st, "col2im");
}

// ============================================================================ small utility kernels
// rows x cols (pitch src_ld) → rows x dst_ld with zero padding (K-padding of conv1 weights for TMA pitch rules)
__global__ void pad_rows_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long rows, int cols,
                                long long src_ld, long long dst_ld) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * dst_ld) return;
  const long long r = idx / dst_ld; const int c = (int)(idx % dst_ld);
  dst[idx] = c < cols ? src[r * src_ld + c] : f_to_bf16(0.f);
}
void pad_rows(const void* src, void* dst, long long rows, int cols, long long src_ld, long long dst_ld, cudaStream_t st) {
  pad_rows_kernel<<<grid_for(rows * dst_ld, 256), 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, rows, cols, src_ld, dst_ld);
  count_launch(); TMPI_CHECK_LAUNCH("pad_rows"); ::tmpi::check_capture(st, "pad_rows");
}

__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int R, int C) {
  __shared__ __nv_bfloat16 tile[32][33];
  int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 32 + threadIdx.y;
  for (int j = 0; j < 32; j += 8) if (c < C && r + j < R) tile[threadIdx.y + j][threadIdx.x] = src[(long long)(r + j) * C + c];
  __syncthreads();
  int oc = blockIdx.y * 32 + threadIdx.x, orow = blockIdx.x * 32 + threadIdx.y;
  for (int j = 0; j < 32; j += 8) if (oc < R && orow + j < C) dst[(long long)(orow + j) * R + oc] = tile[threadIdx.x][threadIdx.y + j];
}
void transpose_bf16(const void* src, void* dst, int R, int C, cudaStream_t st) {
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  transpose_bf16_kernel<<<grid, block, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, R, C);
  count_launch(); TMPI_CHECK_LAUNCH("transpose_bf16"); ::tmpi::check_capture(st, "transpose_bf16");
}

// dgrad of a stride-1 convolution is a forward convolution of dy with the spatially flipped, channel-transposed filter:
// wt[c][KH-1-r][KW-1-s][o] = w[o][r][s][c]     (tiny: runs once per layer per step)
__global__ void conv_weight_flip_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ wt, int O, int KH, int KW, int Cg) {
  const int total = O * KH * KW * Cg;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int o = i % O; int t = i / O;                      // iterate over the OUTPUT layout [c][r'][s'][o] (coalesced writes)
    const int s2 = t % KW; t /= KW;
    const int r2 = t % KH; const int c = t / KH;
    wt[i] = w[(((long long)o * KH + (KH - 1 - r2)) * KW + (KW - 1 - s2)) * Cg + c];
  }
}
void conv_weight_flip(const void* w, void* wt, int O, int KH, int KW, int Cg, cudaStream_t st) {
  const int total = O * KH * KW * Cg;
  conv_weight_flip_kernel<<<std::min(grid_for(total, 256), sm_count() * 8), 256, 0, st>>>((const __nv_bfloat16*)w, (__nv_bfloat16*)wt, O, KH, KW, Cg);
  count_launch(); TMPI_CHECK_LAUNCH("conv_weight_flip"); ::tmpi::check_capture(st, "conv_weight_flip");
}

// ============================================================================ space-to-depth for strided few-channel convs
// A KHxKW / stride-S convolution on C (< 8) channels is the ceil(KH/S) x ceil(KW/S) / stride-1 convolution of the
// space-to-depth image x'[n, i, j, (dy*S+dx)*C + c] = x[n, S*i+dy, S*j+dx, c] with the re-indexed (zero padded) filter.
// That turns AlexNet's conv1 (11x11/4 on RGB, which TMA cannot gather: 6-byte pixels) into a 3x3 conv on 48 channels
// that runs on the implicit-GEMM tcgen05 path.
// One CTA per output row (n, i): for a fixed dy the S*C output channels of pixel j are S*C CONTIGUOUS input elements of image
// row S*i+dy starting at j*S*C — a strided copy of short runs; threads sweep the row's output elements (coalesced stores).
template <int SCT>
__global__ void __launch_bounds__(256) space_to_depth_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H,
                                                             int W, int C, int S, int Hs, int Ws, int Cp, int P) {
  // blockDim = (Cp, 256 / Cp): threadIdx.x = output channel cp, threadIdx.y strides over the pixels j of output row (n, i).
  // SCT > 0: S*C is a compile-time constant, so cp / SC and cp % SC are a multiply-shift (the generic version was
  // issue-bound on runtime divisions: 74 instructions per 2-byte element).
  // P = zero padding of the ORIGINAL convolution, folded in here: output (i, j, (dy, dx, c)) reads x[S*i + dy - P, S*j + dx - P, c].
  const int i = blockIdx.x % Hs, n = blockIdx.x / Hs;
  const int SC = SCT > 0 ? SCT : S * C;
  const int WC = W * C;
  const int cp = threadIdx.x;
  const int dy = cp / SC, e = cp - dy * SC;
  const int h = i * S + dy - P;
  const bool chan_ok = cp < S * SC && h >= 0 && h < H;
  __nv_bfloat16* orow = y + ((long long)n * Hs + i) * Ws * Cp + cp;
  const __nv_bfloat16* irow = x + ((long long)n * H + (chan_ok ? h : 0)) * WC;
  const __nv_bfloat16 zero = f_to_bf16(0.f);
  const int shift = e - P * C;
  for (int j = threadIdx.y; j < Ws; j += blockDim.y) {
    const int col = j * SC + shift;                      // element offset inside the image row
    orow[(unsigned)(j * Cp)] = (chan_ok && col >= 0 && col < WC) ? irow[col] : zero;
  }
}
void space_to_depth(const void* x, void* y, int N, int H, int W, int C, int S, int Hs, int Ws, int Cp, int P, cudaStream_t st) {
  if (Cp > 256) throw std::runtime_error("space_to_depth: more than 256 packed channels");
  const dim3 blk((unsigned)Cp, (unsigned)std::max(1, 256 / Cp));
  auto X = (const __nv_bfloat16*)x; auto Y = (__nv_bfloat16*)y;
  if (S * C == 12) space_to_depth_kernel<12><<<(unsigned)N * Hs, blk, 0, st>>>(X, Y, N, H, W, C, S, Hs, Ws, Cp, P);
  else space_to_depth_kernel<0><<<(unsigned)N * Hs, blk, 0, st>>>(X, Y, N, H, W, C, S, Hs, Ws, Cp, P);
  count_launch(); TMPI_CHECK_LAUNCH("space_to_depth"); ::tmpi::check_capture(st, "space_to_depth");
}

// w [O][KH][KW][C]  <->  ws [O][KHs][KWs][Cp]   with  ws[o, a, b, (dy*S+dx)*C + c] = w[o, S*a+dy, S*b+dx, c]  (0 outside the filter)
// dir 0: pack bf16 filter (w -> ws);  dir 1: unpack fp32 gradient (gs -> g)
__global__ void s2d_filter_kernel(const void* __restrict__ src, void* __restrict__ dst, int O, int KH, int KW, int C, int S, int KHs, int KWs,
                                  int Cp, int dir) {
  if (dir == 0) {
    const int total = O * KHs * KWs * Cp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int cp = i % Cp; int t = i / Cp;
      const int b = t % KWs; t /= KWs;
      const int a = t % KHs; const int o = t / KHs;
      __nv_bfloat16 v = f_to_bf16(0.f);
      if (cp < S * S * C) {
        const int c = cp % C, d = cp / C, dy = d / S, dx = d % S;
        const int kh = a * S + dy, kw = b * S + dx;
        if (kh < KH && kw < KW) v = reinterpret_cast<const __nv_bfloat16*>(src)[(((long long)o * KH + kh) * KW + kw) * C + c];
      }
      reinterpret_cast<__nv_bfloat16*>(dst)[i] = v;
    }
  } else {
    const int total = O * KH * KW * C;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int c = i % C; int t = i / C;
      const int kw = t % KW; t /= KW;
      const int kh = t % KH; const int o = t / KH;
      const int a = kh / S, dy = kh % S, b = kw / S, dx = kw % S;
      reinterpret_cast<float*>(dst)[i] =
          reinterpret_cast<const float*>(src)[(((long long)o * KHs + a) * KWs + b) * Cp + (dy * S + dx) * C + c];
    }
  }
}
void s2d_filter(const void* src, void* dst, int O, int KH, int KW, int C, int S, int KHs, int KWs, int Cp, int dir, cudaStream_t st) {
  const int total = dir == 0 ? O * KHs * KWs * Cp : O * KH * KW * C;
  s2d_filter_kernel<<<std::min(grid_for(total, 256), sm_count() * 8), 256, 0, st>>>(src, dst, O, KH, KW, C, S, KHs, KWs, Cp, dir);
  count_launch(); TMPI_CHECK_LAUNCH("s2d_filter"); ::tmpi::check_capture(st, "s2d_filter");
}

// ============================================================================ loader: normalise + crop + mirror → NHWC bf16/fp32
// One CTA per output row (n, oy): crop offsets / flip flag are CTA-uniform, threads sweep the row's (ox, c) elements with 32-bit
// math (coalesced stores; loads are contiguous runs of the source row, reversed when mirrored).
// grid = (N * ch, ceil(cw / 128)): the output row (n, oy) comes from blockIdx.x, the pixel from blockIdx.y / threadIdx.x — no
// per-thread divisions (the one-thread-per-pixel version with 64-bit div/mod was issue-bound: 230 instructions per pixel).
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(128) crop_mirror_norm_kernel(const Tin* __restrict__ x, const float* __restrict__ mean, int mean_mode,
                                        float scale, const float* __restrict__ cscale, Tout* __restrict__ out, const int* __restrict__ offs,
                                        const uint8_t* __restrict__ flips, int N, int H, int W, int C, int ch, int cw, int Cout) {
  const int ox = blockIdx.y * blockDim.x + threadIdx.x;
  if (ox >= cw) return;
  const int oy = blockIdx.x % ch, n = blockIdx.x / ch;
  const int sy = offs[2 * n] + oy;
  const int sx = offs[2 * n + 1] + (flips[n] ? (cw - 1 - ox) : ox);
  const unsigned pix = (unsigned)(sy * W + sx) * (unsigned)C;          // inside one image (host checks H*W*C < 2^31)
  const Tin* src = x + (long long)n * H * W * C + pix;
  const float* mp = mean_mode == 2 ? mean + pix : mean;
  Tout* o = out + ((long long)blockIdx.x * cw + ox) * Cout;
  if (C == 3 && Cout == 3) {
    const float m0 = mean_mode == 0 ? mp[0] : mp[0], m1 = mean_mode == 0 ? mp[0] : mp[1], m2 = mean_mode == 0 ? mp[0] : mp[2];
    // per-channel 1/std on top of the scalar scale (ref proc_load_mpi.py:99: (arr - img_mean) / 255. / img_std)
    const float s0 = cscale ? scale * cscale[0] : scale, s1 = cscale ? scale * cscale[1] : scale, s2 = cscale ? scale * cscale[2] : scale;
    const float v0 = ((float)src[0] - m0) * s0, v1 = ((float)src[1] - m1) * s1, v2 = ((float)src[2] - m2) * s2;
    o[0] = (Tout)v0; o[1] = (Tout)v1; o[2] = (Tout)v2;
    return;
  }
  for (int c = 0; c < Cout; ++c) {
    float v = 0.f;
    if (c < C) {
      const float m = mean_mode == 0 ? mp[0] : mp[c];
      v = ((float)src[c] - m) * (cscale ? scale * cscale[c] : scale);
    }
    o[c] = (Tout)v;
  }
}

void crop_mirror_norm(const void* x, int in_kind /*0 u8, 1 bf16, 2 f32*/, const void* mean, int mean_mode, float scale, const void* cscale, void* out,
                      int out_bf16, const void* offs, const void* flips, int N, int H, int W, int C, int ch, int cw, int Cout, cudaStream_t st) {
  if ((long long)H * W * C >= (1LL << 31) || (long long)N * ch >= (1LL << 31)) throw std::runtime_error("crop_mirror_norm: image too large");
  const dim3 g((unsigned)(N * ch), (unsigned)((cw + 127) / 128));
  auto M = (const float*)mean; auto O = (const int*)offs; auto F = (const uint8_t*)flips;
#define CMN(TI, TO) crop_mirror_norm_kernel<TI, TO><<<g, 128, 0, st>>>((const TI*)x, M, mean_mode, scale, (const float*)cscale, (TO*)out, O, F, N, H, W, C, ch, cw, Cout)
  if (in_kind == 0) { if (out_bf16) CMN(uint8_t, __nv_bfloat16); else CMN(uint8_t, float); }
  else if (in_kind == 1) { if (out_bf16) CMN(__nv_bfloat16, __nv_bfloat16); else CMN(__nv_bfloat16, float); }
  else { if (out_bf16) CMN(float, __nv_bfloat16); else CMN(float, float); }
#undef CMN
  count_launch(); TMPI_CHECK_LAUNCH("crop_mirror_norm"); ::tmpi::check_capture(st, "crop_mirror_norm");
}

}  // namespace tmpi
