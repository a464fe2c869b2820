// fp32-storage variants of the non-GEMM layer kernels: the activation path of the tf32 precision mode (fp32 tensors end to end,
// GEMMs / convolutions on tcgen05 kind::tf32 — the reference computes in fp32, theanompi/models/layers2.py:380-388, :927-929).
// Same semantics as nn_kernels.cu (NHWC, C % 4 == 0 → 16-byte float4 vectors, argmax-exact max pooling, Philox dropout keyed by
// the device step counter); these are plain streaming kernels — the bf16 versions carry the packed-SIMD tricks.
#include "common.cuh"
#include "api.h"
#include <algorithm>

namespace tmpi {

static inline int grid_f(long long n, int block) { return (int)((n + block - 1) / block); }
static inline void need_c4(int C, const char* what) {
  if (C % 4) throw std::runtime_error(std::string(what) + ": C must be a multiple of 4 (fp32 path)");
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float& comp(float4& v, int i) { return reinterpret_cast<float*>(&v)[i]; }

// ============================================================================ LRN (across channels, window n = 2*half+1)
__global__ void lrn_fwd_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows, int C, int half, float k,
                                   float alpha, float beta) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * C) return;
  const long long r = idx / C; const int c = (int)(idx - r * C);
  const float* row = x + r * C;
  float s = 0.f;
  for (int j = max(0, c - half); j <= min(C - 1, c + half); ++j) s += row[j] * row[j];
  y[idx] = row[c] * exp2f(-beta * __log2f(k + alpha * s));
}
__global__ void lrn_bwd_f32_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, long long rows, int C,
                                   int half, float k, float alpha, float beta) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * C) return;
  const long long r = idx / C; const int c = (int)(idx - r * C);
  const float* xr = x + r * C; const float* dr = dy + r * C;
  // dx_c = dy_c * s_c^-beta - 2 alpha beta x_c * sum_{i : |i - c| <= half} dy_i x_i s_i^(-beta-1)
  float acc = 0.f, pc = 0.f;
  for (int i = max(0, c - half); i <= min(C - 1, c + half); ++i) {
    float s = 0.f;
    for (int j = max(0, i - half); j <= min(C - 1, i + half); ++j) s += xr[j] * xr[j];
    const float scale = k + alpha * s;
    const float p = exp2f(-beta * __log2f(scale));
    acc += dr[i] * xr[i] * p / scale;
    if (i == c) pc = p;
  }
  dx[idx] = dr[c] * pc - 2.f * alpha * beta * xr[c] * acc;
}
void lrn_fwd_f32(const void* x, void* y, long long rows, int C, int n, float k, float alpha, float beta, cudaStream_t st) {
  lrn_fwd_f32_kernel<<<grid_f(rows * C, 256), 256, 0, st>>>((const float*)x, (float*)y, rows, C, n / 2, k, alpha, beta);
  count_launch(); TMPI_CHECK_LAUNCH("lrn_fwd_f32"); ::tmpi::check_capture(st, "lrn_fwd_f32");
}
void lrn_bwd_f32(const void* x, const void* dy, void* dx, long long rows, int C, int n, float k, float alpha, float beta, cudaStream_t st) {
  lrn_bwd_f32_kernel<<<grid_f(rows * C, 256), 256, 0, st>>>((const float*)x, (const float*)dy, (float*)dx, rows, C, n / 2, k, alpha, beta);
  count_launch(); TMPI_CHECK_LAUNCH("lrn_bwd_f32"); ::tmpi::check_capture(st, "lrn_bwd_f32");
}

// ============================================================================ pooling
struct PoolGeomF { int N, H, W, C, Ho, Wo, k, s, p; };

__global__ void pool_fwd_f32_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ arg, PoolGeomF g, int is_max) {
  const int nvec = g.C >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)g.N * g.Ho * g.Wo * nvec;
  if (idx >= total) return;
  const int cv = (int)(idx % nvec); long long t = idx / nvec;
  const int wo = (int)(t % g.Wo); t /= g.Wo;
  const int ho = (int)(t % g.Ho); const int n = (int)(t / g.Ho);
  float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), sum = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t bi[4] = {0u, 0u, 0u, 0u};
  int cnt = 0;
  for (int kh = 0; kh < g.k; ++kh) {
    const int h = ho * g.s - g.p + kh;
    if (h < 0 || h >= g.H) continue;
    for (int kw = 0; kw < g.k; ++kw) {
      const int w = wo * g.s - g.p + kw;
      if (w < 0 || w >= g.W) continue;
      float4 v = ld4(x + (((long long)n * g.H + h) * g.W + w) * g.C + cv * 4);
      ++cnt;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float vi = comp(v, i);
        if (vi > comp(best, i)) { comp(best, i) = vi; bi[i] = (uint32_t)(kh * g.k + kw); }
        comp(sum, i) += vi;
      }
    }
  }
  const long long o = (((long long)n * g.Ho + ho) * g.Wo + wo) * g.C + cv * 4;
  if (is_max) {
    st4(y + o, best);
    *reinterpret_cast<uint32_t*>(arg + o) = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
  } else {
    const float inv = 1.f / (float)max(1, cnt);
    st4(y + o, make_float4(sum.x * inv, sum.y * inv, sum.z * inv, sum.w * inv));
  }
}
__global__ void pool_bwd_f32_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ arg, float* __restrict__ dx, PoolGeomF g,
                                    int is_max) {
  const int nvec = g.C >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)g.N * g.H * g.W * nvec;
  if (idx >= total) return;
  const int cv = (int)(idx % nvec); long long t = idx / nvec;
  const int w = (int)(t % g.W); t /= g.W;
  const int h = (int)(t % g.H); const int n = (int)(t / g.H);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int ho_hi = min(g.Ho - 1, (h + g.p) / g.s), wo_hi = min(g.Wo - 1, (w + g.p) / g.s);
  for (int ho = max(0, (h + g.p - g.k + g.s) / g.s); ho <= ho_hi; ++ho) {
    const int kh = h + g.p - ho * g.s;
    if (kh < 0 || kh >= g.k) continue;
    for (int wo = max(0, (w + g.p - g.k + g.s) / g.s); wo <= wo_hi; ++wo) {
      const int kw = w + g.p - wo * g.s;
      if (kw < 0 || kw >= g.k) continue;
      const long long o = (((long long)n * g.Ho + ho) * g.Wo + wo) * g.C + cv * 4;
      float4 d = ld4(dy + o);
      if (is_max) {
        const uint32_t a = *reinterpret_cast<const uint32_t*>(arg + o);
        const uint32_t me = (uint32_t)(kh * g.k + kw);
#pragma unroll
        for (int i = 0; i < 4; ++i) if (((a >> (8 * i)) & 0xFFu) == me) comp(acc, i) += comp(d, i);
      } else {
        const int h0 = max(0, ho * g.s - g.p), h1 = min(g.H, ho * g.s - g.p + g.k);
        const int w0 = max(0, wo * g.s - g.p), w1 = min(g.W, wo * g.s - g.p + g.k);
        const float inv = 1.f / (float)max(1, (h1 - h0) * (w1 - w0));
        acc.x += d.x * inv; acc.y += d.y * inv; acc.z += d.z * inv; acc.w += d.w * inv;
      }
    }
  }
  st4(dx + (((long long)n * g.H + h) * g.W + w) * g.C + cv * 4, acc);
}
void pool_fwd_f32(const void* x, void* y, void* arg, int N, int H, int W, int C, int Ho, int Wo, int k, int s, int p, int is_max, cudaStream_t st) {
  need_c4(C, "pool_f32");
  PoolGeomF g{N, H, W, C, Ho, Wo, k, s, p};
  pool_fwd_f32_kernel<<<grid_f((long long)N * Ho * Wo * (C / 4), 256), 256, 0, st>>>((const float*)x, (float*)y, (uint8_t*)arg, g, is_max);
  count_launch(); TMPI_CHECK_LAUNCH("pool_fwd_f32"); ::tmpi::check_capture(st, "pool_fwd_f32");
}
void pool_bwd_f32(const void* dy, const void* arg, void* dx, int N, int H, int W, int C, int Ho, int Wo, int k, int s, int p, int is_max,
                  cudaStream_t st) {
  need_c4(C, "pool_f32");
  PoolGeomF g{N, H, W, C, Ho, Wo, k, s, p};
  pool_bwd_f32_kernel<<<grid_f((long long)N * H * W * (C / 4), 256), 256, 0, st>>>((const float*)dy, (const uint8_t*)arg, (float*)dx, g, is_max);
  count_launch(); TMPI_CHECK_LAUNCH("pool_bwd_f32"); ::tmpi::check_capture(st, "pool_bwd_f32");
}

// ============================================================================ dropout (same Philox stream layout as the bf16 kernel: 8 draws per counter)
__device__ __forceinline__ void philox4x32_f(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
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
__global__ void dropout_fwd_f32_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ mask, long long n8, float p_drop,
                                       unsigned long long seed, uint32_t layer, const unsigned long long* __restrict__ step) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n8) return;
  const unsigned long long stp = *step;
  uint32_t r[4];
  philox4x32_f((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)stp, (uint32_t)(stp >> 32), (uint32_t)seed,
               (uint32_t)(seed >> 32) ^ (layer * 0x9E3779B9u), r);
  const uint32_t thr = (uint32_t)(p_drop * 65536.f);
  float4 a = ld4(x + idx * 8), b = ld4(x + idx * 8 + 4);
  uint32_t mlo = 0, mhi = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t u = (r[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
    const uint32_t keep = u >= thr ? 1u : 0u;
    float& v = i < 4 ? comp(a, i) : comp(b, i - 4);
    v = keep ? v : 0.f;
    if (i < 4) mlo |= keep << (8 * i); else mhi |= keep << (8 * (i - 4));
  }
  st4(y + idx * 8, a); st4(y + idx * 8 + 4, b);
  *reinterpret_cast<uint2*>(mask + idx * 8) = make_uint2(mlo, mhi);
}
__global__ void dropout_bwd_f32_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ mask, float* __restrict__ dx, long long n4) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n4) return;
  float4 v = ld4(dy + idx * 4);
  const uint32_t m = *reinterpret_cast<const uint32_t*>(mask + idx * 4);
#pragma unroll
  for (int i = 0; i < 4; ++i) if (!((m >> (8 * i)) & 0xFFu)) comp(v, i) = 0.f;
  st4(dx + idx * 4, v);
}
void dropout_fwd_f32(const void* x, void* y, void* mask, long long n, float p_drop, unsigned long long seed, int layer, const void* step,
                     cudaStream_t st) {
  if (n % 8) throw std::runtime_error("dropout_f32: numel must be a multiple of 8");
  dropout_fwd_f32_kernel<<<grid_f(n / 8, 256), 256, 0, st>>>((const float*)x, (float*)y, (uint8_t*)mask, n / 8, p_drop, seed, (uint32_t)layer,
                                                            (const unsigned long long*)step);
  count_launch(); TMPI_CHECK_LAUNCH("dropout_fwd_f32"); ::tmpi::check_capture(st, "dropout_fwd_f32");
}
void dropout_bwd_f32(const void* dy, const void* mask, void* dx, long long n, cudaStream_t st) {
  dropout_bwd_f32_kernel<<<grid_f(n / 4, 256), 256, 0, st>>>((const float*)dy, (const uint8_t*)mask, (float*)dx, n / 4);
  count_launch(); TMPI_CHECK_LAUNCH("dropout_bwd_f32"); ::tmpi::check_capture(st, "dropout_bwd_f32");
}

// ============================================================================ softmax + NLL + errors + dlogits (one CTA per row)
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = lane < nw ? red[lane] : (is_max ? -INFINITY : 0.f);
  r = is_max ? warp_max(r) : warp_sum(r);
  return __shfl_sync(0xffffffffu, r, 0);
}
__global__ void softmax_xent_f32_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, float* __restrict__ dlogits,
                                        float* __restrict__ rowstat, int C, float scale) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float* row = logits + (long long)b * C;
  const int label = (int)labels[b];
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, row[c]);
  mx = block_reduce(mx, red, true);
  const float lab = row[label];
  float se = 0.f, gt = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = row[c];
    se += __expf(v - mx);
    gt += (v > lab || (v == lab && c < label)) ? 1.f : 0.f;
  }
  se = block_reduce(se, red, false);
  gt = block_reduce(gt, red, false);
  const float inv = 1.f / se;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float pr = __expf(row[c] - mx) * inv;
    if (c == label) pr -= 1.f;
    dlogits[(long long)b * C + c] = pr * scale;
  }
  if (threadIdx.x == 0) {
    rowstat[3 * b + 0] = -(lab - mx - __logf(se));
    rowstat[3 * b + 1] = gt >= 1.f ? 1.f : 0.f;
    rowstat[3 * b + 2] = gt >= 5.f ? 1.f : 0.f;
  }
}
__global__ void rowstat_mean_f32_kernel(const float* __restrict__ rowstat, float* __restrict__ out, int B, float weight) {
  __shared__ float red[32];
  float a = 0.f, b = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) { a += rowstat[3 * i]; b += rowstat[3 * i + 1]; c += rowstat[3 * i + 2]; }
  a = block_reduce(a, red, false); b = block_reduce(b, red, false); c = block_reduce(c, red, false);
  if (threadIdx.x == 0) { out[0] = weight * a / B; out[1] = b / B; out[2] = c / B; }
}
void softmax_xent_f32(const void* logits, const void* labels, void* dlogits, void* rowstat, void* out3, int B, int C, float weight, cudaStream_t st) {
  softmax_xent_f32_kernel<<<B, 256, 0, st>>>((const float*)logits, (const long long*)labels, (float*)dlogits, (float*)rowstat, C, weight / (float)B);
  rowstat_mean_f32_kernel<<<1, 256, 0, st>>>((const float*)rowstat, (float*)out3, B, weight);
  count_launch(2); TMPI_CHECK_LAUNCH("softmax_xent_f32"); ::tmpi::check_capture(st, "softmax_xent_f32");
}

// ============================================================================ ReLU mask + bias gradient
// dym = dy * (y > 0) (contiguous [R, C]);  db[c] = sum_r dym[r, c].  dy / y have row pitch ld.  One CTA per 64-row slab and
// channel vector group: per-thread column sums in registers → shared memory → one atomicAdd per (CTA, channel).
__global__ void __launch_bounds__(256) relu_bias_bwd_f32_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dym,
                                                               float* __restrict__ db, float* __restrict__ db1, int c_split, long long R, int C,
                                                               long long ld, int relu, int VT, int rows_per_cta) {
  extern __shared__ float sm[];                       // [RL][VT*4]
  const int nvec = C >> 2;
  const int RL = blockDim.x / VT;
  const int tv = threadIdx.x % VT, tr = threadIdx.x / VT;
  const int cv = blockIdx.y * VT + tv;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tr < RL && cv < nvec) {
    const long long rend = min(R, r0 + rows_per_cta);
    for (long long r = r0 + tr; r < rend; r += RL) {
      float4 d = ld4(dy + r * ld + cv * 4);
      if (relu) {
        const float4 v = ld4(y + r * ld + cv * 4);
        if (!(v.x > 0.f)) d.x = 0.f; if (!(v.y > 0.f)) d.y = 0.f; if (!(v.z > 0.f)) d.z = 0.f; if (!(v.w > 0.f)) d.w = 0.f;
      }
      if (dym) st4(dym + r * C + cv * 4, d);
      acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
    }
  }
  if (db == nullptr) return;
  if (tr < RL) st4(sm + (tr * VT + tv) * 4, acc);
  __syncthreads();
  if (threadIdx.x < VT * 4) {
    const int v = threadIdx.x / 4, i = threadIdx.x % 4;
    const int c = (blockIdx.y * VT + v) * 4 + i;
    if (c < C) {
      float s = 0.f;
      for (int t = 0; t < RL; ++t) s += sm[(t * VT + v) * 4 + i];
      atomicAdd(c < c_split ? db + c : db1 + (c - c_split), s);
    }
  }
}
void relu_bias_bwd2_f32(const void* dy, const void* y, void* dym, void* db, void* db1, int c_split, long long R, int C, long long ld, int relu,
                        cudaStream_t st) {
  need_c4(C, "relu_bias_bwd_f32");
  const int nvec = C / 4;
  const int VT = nvec < 32 ? nvec : 32;
  const int RL = 256 / VT;
  const int rows_per_cta = deterministic_mode() ? (int)std::max<long long>(R, 1) : RL * 8;
  dim3 grid((unsigned)((R + rows_per_cta - 1) / rows_per_cta), (unsigned)((nvec + VT - 1) / VT));
  const size_t smem = (size_t)RL * VT * 4 * sizeof(float);
  if (!db1 || c_split > C) c_split = C;
  if (db) check_cuda(cudaMemsetAsync(db, 0, (size_t)c_split * 4, st), "relu_bias_bwd_f32 memset");
  if (db && c_split < C) check_cuda(cudaMemsetAsync(db1, 0, (size_t)(C - c_split) * 4, st), "relu_bias_bwd_f32 memset");
  relu_bias_bwd_f32_kernel<<<grid, 256, smem, st>>>((const float*)dy, (const float*)y, (float*)dym, (float*)db, (float*)db1, c_split, R, C, ld,
                                                    relu, VT, rows_per_cta);
  count_launch(); TMPI_CHECK_LAUNCH("relu_bias_bwd_f32"); ::tmpi::check_capture(st, "relu_bias_bwd_f32");
}

// y[r, c] = act(acc[r, c] + bias[c])  (finishing pass of a split-K forward GEMM; may run in place)
__global__ void bias_act_f32_kernel(const float* __restrict__ acc, const float* __restrict__ bias, float* __restrict__ y, long long R, int C,
                                    int relu) {
  const int nvec = C >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * nvec) return;
  const long long r = idx / nvec; const int cv = (int)(idx - r * nvec);
  float4 a = ld4(acc + r * C + cv * 4);
  if (bias) { const float4 b = ld4(bias + cv * 4); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
  if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
  st4(y + r * C + cv * 4, a);
}
void bias_act_f32(const void* acc, const void* bias, void* y, int R, int C, int relu, cudaStream_t st) {
  need_c4(C, "bias_act_f32");
  bias_act_f32_kernel<<<grid_f((long long)R * (C / 4), 256), 256, 0, st>>>((const float*)acc, (const float*)bias, (float*)y, R, C, relu);
  count_launch(); TMPI_CHECK_LAUNCH("bias_act_f32"); ::tmpi::check_capture(st, "bias_act_f32");
}

// ============================================================================ explicit im2col (first layers with C = 3) / col2im (strided dgrad)
struct ConvGeomF { int N, H, W, Ctot, c_off, Cg, KH, KW, Ho, Wo, s, p; long long ldcol; int K; };
__global__ void im2col_f32_kernel(const float* __restrict__ x, float* __restrict__ col, ConvGeomF g) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long M = (long long)g.N * g.Ho * g.Wo;
  if (idx >= M * g.ldcol) return;
  const long long m = idx / g.ldcol; const int e = (int)(idx - m * g.ldcol);
  float v = 0.f;
  if (e < g.K) {
    const int kk = e / g.Cg, c = e - kk * g.Cg;
    const int kh = kk / g.KW, kw = kk - kh * g.KW;
    const int wo = (int)(m % g.Wo); long long t = m / g.Wo;
    const int ho = (int)(t % g.Ho); const int n = (int)(t / g.Ho);
    const int h = ho * g.s - g.p + kh, w = wo * g.s - g.p + kw;
    if (h >= 0 && h < g.H && w >= 0 && w < g.W) v = x[(((long long)n * g.H + h) * g.W + w) * g.Ctot + g.c_off + c];
  }
  col[idx] = v;
}
__global__ void col2im_f32_kernel(const float* __restrict__ dcol, float* __restrict__ dx, ConvGeomF g) {
  const int cvn = g.Cg >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)g.N * g.H * g.W * cvn;
  if (idx >= total) return;
  const int cv = (int)(idx % cvn); long long t = idx / cvn;
  const int w = (int)(t % g.W); t /= g.W;
  const int h = (int)(t % g.H); const int n = (int)(t / g.H);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
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
      const float4 v = ld4(dcol + m * g.ldcol + (long long)(kh * g.KW + kw) * g.Cg + cv * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  st4(dx + (((long long)n * g.H + h) * g.W + w) * g.Ctot + g.c_off + cv * 4, acc);
}
void im2col_f32(const void* x, void* col, int N, int H, int W, int Ctot, int c_off, int Cg, int KH, int KW, int Ho, int Wo, int s, int p,
                long long ldcol, cudaStream_t st) {
  ConvGeomF g{N, H, W, Ctot, c_off, Cg, KH, KW, Ho, Wo, s, p, ldcol, KH * KW * Cg};
  im2col_f32_kernel<<<grid_f((long long)N * Ho * Wo * ldcol, 256), 256, 0, st>>>((const float*)x, (float*)col, g);
  count_launch(); TMPI_CHECK_LAUNCH("im2col_f32"); ::tmpi::check_capture(st, "im2col_f32");
}
void col2im_f32(const void* dcol, void* dx, int N, int H, int W, int Ctot, int c_off, int Cg, int KH, int KW, int Ho, int Wo, int s, int p,
                long long ldcol, cudaStream_t st) {
  if (Cg % 4 || c_off % 4 || Ctot % 4 || ldcol % 4) throw std::runtime_error("col2im_f32: channel counts must be multiples of 4");
  ConvGeomF g{N, H, W, Ctot, c_off, Cg, KH, KW, Ho, Wo, s, p, ldcol, KH * KW * Cg};
  col2im_f32_kernel<<<grid_f((long long)N * H * W * (Cg / 4), 256), 256, 0, st>>>((const float*)dcol, (float*)dx, g);
  count_launch(); TMPI_CHECK_LAUNCH("col2im_f32"); ::tmpi::check_capture(st, "col2im_f32");
}
__global__ void pad_rows_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long rows, int cols, long long src_ld, long long dst_ld) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * dst_ld) return;
  const long long r = idx / dst_ld; const int c = (int)(idx % dst_ld);
  dst[idx] = c < cols ? src[r * src_ld + c] : 0.f;
}
void pad_rows_f32(const void* src, void* dst, long long rows, int cols, long long src_ld, long long dst_ld, cudaStream_t st) {
  pad_rows_f32_kernel<<<grid_f(rows * dst_ld, 256), 256, 0, st>>>((const float*)src, (float*)dst, rows, cols, src_ld, dst_ld);
  count_launch(); TMPI_CHECK_LAUNCH("pad_rows_f32"); ::tmpi::check_capture(st, "pad_rows_f32");
}

// ============================================================================ space-to-depth (strided few-channel first layer) — see nn_kernels.cu
__global__ void space_to_depth_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C, int S, int Hs, int Ws,
                                          int Cp, int P) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * Hs * Ws * Cp;
  if (idx >= total) return;
  const int cp = (int)(idx % Cp); long long t = idx / Cp;
  const int j = (int)(t % Ws); t /= Ws;
  const int i = (int)(t % Hs); const int n = (int)(t / Hs);
  float v = 0.f;
  if (cp < S * S * C) {
    const int c = cp % C, d = cp / C, dy = d / S, dx = d % S;
    const int h = i * S + dy - P, w = j * S + dx - P;          // P: zero padding of the original convolution, folded in
    if (h >= 0 && h < H && w >= 0 && w < W) v = x[(((long long)n * H + h) * W + w) * C + c];
  }
  y[idx] = v;
}
void space_to_depth_f32(const void* x, void* y, int N, int H, int W, int C, int S, int Hs, int Ws, int Cp, int P, cudaStream_t st) {
  space_to_depth_f32_kernel<<<grid_f((long long)N * Hs * Ws * Cp, 256), 256, 0, st>>>((const float*)x, (float*)y, N, H, W, C, S, Hs, Ws, Cp, P);
  count_launch(); TMPI_CHECK_LAUNCH("space_to_depth_f32"); ::tmpi::check_capture(st, "space_to_depth_f32");
}
// fp32 filter pack: ws[o, a, b, (dy*S+dx)*C + c] = w[o, S*a+dy, S*b+dx, c]   (the unpack direction is type-agnostic: s2d_filter dir 1)
__global__ void s2d_filter_pack_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int O, int KH, int KW, int C, int S, int KHs,
                                           int KWs, int Cp) {
  const int total = O * KHs * KWs * Cp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cp = i % Cp; int t = i / Cp;
    const int b = t % KWs; t /= KWs;
    const int a = t % KHs; const int o = t / KHs;
    float v = 0.f;
    if (cp < S * S * C) {
      const int c = cp % C, d = cp / C, dy = d / S, dx = d % S;
      const int kh = a * S + dy, kw = b * S + dx;
      if (kh < KH && kw < KW) v = src[(((long long)o * KH + kh) * KW + kw) * C + c];
    }
    dst[i] = v;
  }
}
void s2d_filter_pack_f32(const void* src, void* dst, int O, int KH, int KW, int C, int S, int KHs, int KWs, int Cp, cudaStream_t st) {
  const int total = O * KHs * KWs * Cp;
  s2d_filter_pack_f32_kernel<<<std::min(grid_f(total, 256), sm_count() * 8), 256, 0, st>>>((const float*)src, (float*)dst, O, KH, KW, C, S, KHs, KWs, Cp);
  count_launch(); TMPI_CHECK_LAUNCH("s2d_filter_pack_f32"); ::tmpi::check_capture(st, "s2d_filter_pack_f32");
}

}  // namespace tmpi
