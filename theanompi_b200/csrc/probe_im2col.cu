// Stand-alone probe (not part of the extension): discover the exact semantics of TMA im2col loads on sm_100a by
// comparing the shared-memory tile they produce with a reference gather, for several candidate conventions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o probe probe_im2col.cu && ./probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*PFN_im2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one CTA: load a [PIX x CH] tile with one im2col TMA, copy the raw smem bytes to global
__global__ void probe_kernel(const __grid_constant__ CUtensorMap tmap, __nv_bfloat16* out, int pix, int ch, int c0, int w0, int h0, int n0,
                             int off_w, int off_h) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t b = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const uint32_t dst = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(pix * ch * 2) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(dst), "l"(&tmap), "r"(b), "r"(c0), "r"(w0), "r"(h0), "r"(n0), "h"((uint16_t)off_w), "h"((uint16_t)off_h) : "memory");
  }
  // bounded wait
  uint32_t ok = 0; long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
    if (clock64() - t0 > 2000000000LL) { if (threadIdx.x == 0) printf("TIMEOUT waiting for TMA\n"); return; }
  }
  const uint8_t* src = smem + (dst - smem_u32(smem));
  for (int i = threadIdx.x; i < pix * ch; i += blockDim.x) out[i] = reinterpret_cast<const __nv_bfloat16*>(src)[i];
}

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no cuTensorMapEncodeIm2col\n"); return 1; }
  PFN_im2col enc = (PFN_im2col)fn;
  // activation NHWC: N=2, H=6, W=7, C=64 (value encodes (n,h,w,c)): v = n*1000 + h*100 + w*10 + c/8 (fits bf16 exactly for small ints? use small ranges)
  const int N = 2, H = 6, W = 7, C = 64;
  std::vector<__nv_bfloat16> hx((size_t)N * H * W * C);
  for (int n = 0; n < N; ++n) for (int h = 0; h < H; ++h) for (int w = 0; w < W; ++w) for (int c = 0; c < C; ++c)
    hx[(((size_t)n * H + h) * W + w) * C + c] = __float2bfloat16((float)(n * 100 + h * 10 + w) + (c == 0 ? 0.f : 0.f));   // pixel id (exact in bf16 up to 256)
  __nv_bfloat16 *dx, *dout; cudaMalloc(&dx, hx.size() * 2); cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  const int PIX = 64, CH = 64;
  cudaMalloc(&dout, PIX * CH * 2);
  std::vector<__nv_bfloat16> hout(PIX * CH);
  // conv: 3x3, pad 1, stride S (1 or 2)
  for (int S = 1; S <= 2; ++S) {
    const int pad = 1, R = 3;
    const int P = (H + 2 * pad - R) / S + 1, Q = (W + 2 * pad - R) / S + 1;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    int lower[2] = {-pad, -pad};
    int upper[2] = {pad - (R - 1), pad - (R - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)S, (cuuint32_t)S, 1};
    CUtensorMap tm;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, lower, upper, CH, PIX, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("== stride %d: encode result %d, P=%d Q=%d\n", S, (int)r, P, Q);
    if (r != CUDA_SUCCESS) continue;
    // candidate start-coordinate conventions for output pixel m0 = (n=0, p=1, q=2) and tap (r=2, s=0)
    const int p0 = 1, q0 = 2, tr = 2, ts = 0;
    int cands[3][2] = {{q0 * S - pad, p0 * S - pad}, {q0 * S, p0 * S}, {q0, p0}};
    for (int ci = 0; ci < 3; ++ci) {
      cudaMemset(dout, 0xFF, PIX * CH * 2);
      probe_kernel<<<1, 128, PIX * CH * 2 + 2048>>>(tm, dout, PIX, CH, 0, cands[ci][0], cands[ci][1], 0, ts, tr);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("cand %d: CUDA error %s\n", ci, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(hout.data(), dout, PIX * CH * 2, cudaMemcpyDeviceToHost);
      // reference: pixel j of the tile = output position m0 + j in (n,p,q) order; input pixel (p*S - pad + tr, q*S - pad + ts), zero if outside
      int match = 0, first_bad = -1;
      printf("cand %d start(w=%d,h=%d): got  ", ci, cands[ci][0], cands[ci][1]);
      for (int j = 0; j < PIX; ++j) {
        int m = (0 * P + p0) * Q + q0 + j; int n = m / (P * Q), rem = m % (P * Q), p = rem / Q, qq = rem % Q;
        int ih = p * S - pad + tr, iw = qq * S - pad + ts;
        float want = (n < N && ih >= 0 && ih < H && iw >= 0 && iw < W) ? (float)(n * 100 + ih * 10 + iw) : 0.f;
        float got = __bfloat162float(hout[j * CH + 5]);
        if (j < 24) printf("%g ", got);
        if (got == want) ++match; else if (first_bad < 0) first_bad = j;
      }
      printf("\n        match %d/%d first_bad %d\n", match, PIX, first_bad);
      if (ci == 0) {
        printf("        want: ");
        for (int j = 0; j < 24; ++j) {
          int m = (0 * P + p0) * Q + q0 + j; int n = m / (P * Q), rem = m % (P * Q), p = rem / Q, qq = rem % Q;
          int ih = p * S - pad + tr, iw = qq * S - pad + ts;
          printf("%g ", (n < N && ih >= 0 && ih < H && iw >= 0 && iw < W) ? (float)(n * 100 + ih * 10 + iw) : 0.f);
        }
        printf("\n");
      }
    }
  }
  return 0;
}
