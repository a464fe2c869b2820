// Hand-written Blackwell GEMM / implicit-GEMM convolution:  C[M,N] = alpha * op(A) · op(B)  (+bias, +ReLU), bf16 in, fp32 accumulate.
//
//   * persistent: one CTA per SM walks the tile list; 320 threads = warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer,
//     warps 2..9 epilogue (warp % 4 selects the TMEM lane quarter it may read, (warp - 2) / 4 the column half);
//   * operands staged global→shared by TMA (tiled 2-D / 3-D boxes or im2col-mode 4-D boxes, 128 B swizzle) into a 4–8 stage
//     mbarrier ring; tcgen05.mma.cta_group::1.kind::f16 issued by one elected lane, accumulators in TMEM (2 stages, so the
//     epilogue of tile i overlaps the main loop of tile i+1); tcgen05.commit hands smem slots back / signals the epilogue;
//   * MT = 2: a CTA computes a 256-row tile as two MMAs per k-step against ONE B tile (bf16-output GEMMs with enough tiles);
//   * epilogue: tcgen05.ld 32x32b.x32 → bias (prefetched, shuffle-broadcast) / ReLU → per-warp smem transpose of one
//     32x32 chunk → coalesced 16 B row-segment stores, or red.global.add.v4.f32 for split-K;
//   * two same-shape problems (the two groups of a grouped convolution) can share one launch.
//
// Both operands may be K-major ([rows, K], K contiguous) or MN-major ([K, rows], rows contiguous): forward, dgrad and wgrad of
// FC and conv layers all run on this one kernel without transposed copies (reference: cuBLAS SGEMM / cuDNN through Theano,
// layers2.py:927-929, :380-388, GpuCorrMM :597-653).  What bounded the first versions and how it was found:
// profiles/gemm_probe.md.
#include "common.cuh"
#include "api.h"
#include <cuda.h>
#include <algorithm>
#include <map>
#include <vector>
#include <mutex>
#include <tuple>

namespace tmpi {
std::atomic<unsigned long long> g_launch_count{0};

namespace gemm {

constexpr int BM = 128;
constexpr int NUM_THREADS = 320;       // warp0 TMA, warp1 MMA, warps 2..9 epilogue (2 per TMEM lane quarter)
constexpr int NUM_EPI_WARPS = 8;

// Operand element type: __nv_bfloat16 (tcgen05 kind::f16) or float (kind::tf32: fp32 storage, the tensor core reads the top
// 19 bits).  A k-block is always ONE 128-byte swizzle row per operand row, so the byte geometry of the smem ring, the UMMA
// descriptors' K-major stepping (32 B per MMA) and the TMA box widths in bytes are identical for both types; what changes is
// the number of ELEMENTS per k-block / MMA / MN-major atom.
template <typename T> struct Elem {
  static constexpr int ESZ = (int)sizeof(T);
  static constexpr int BK = 128 / ESZ;            // elements per k-block: 64 (bf16) / 32 (tf32)
  static constexpr int UMMA_K = 32 / ESZ;         // K of one tcgen05.mma: 16 / 8
  static constexpr int ATOM = 128 / ESZ;          // MN-major: elements of one 128-byte atom row (TMA box width): 64 / 32
  static constexpr bool TF32 = ESZ == 4;
};

// MT = number of 128-row sub-tiles a CTA computes per k-block against ONE copy of the B tile (MT = 2: a 256 x BN tile as two
// MMAs per k-step — half the B (weight) reads per flop, half the per-k-block barrier / issue overhead per flop).
template <typename T, int BN, int MT> struct Cfg {
  // persistent kernel, one CTA per SM: operand ring + per-warp epilogue staging chunks + 2 TMEM accumulator stages
  static constexpr int A_BYTES = MT * BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // epilogue staging: each epilogue warp transposes one 32-row x 32-column chunk at a time through its own region
  // (row pitch = chunk bytes + 16 B: conflict-free 16 B accesses).  bf16 MT = 2 tiles are bf16-output only (fprop / dgrad).
  static constexpr int STAGING_ROW = ((MT == 2 && sizeof(T) == 2) ? 32 * 2 : 32 * 4) + 16;
  static constexpr int STAGING_BYTES = NUM_EPI_WARPS * 32 * STAGING_ROW;
  static constexpr int SMEM_LIMIT = 232448;                                // 227 KB per CTA
  static constexpr int RING_BUDGET = SMEM_LIMIT - STAGING_BYTES - 1024 /*align slack*/ - 256 /*barriers*/;
  static constexpr int STAGES = (RING_BUDGET / STAGE_BYTES) > 8 ? 8 : (RING_BUDGET / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + 256;
  static constexpr int ACC_STAGES = 2;
  static constexpr int ACC_COLS = MT * BN;                                 // TMEM columns of one accumulator stage
  static constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512)));
  static_assert(STAGES >= 3, "operand ring too shallow");
  static_assert(2 * ACC_COLS <= 512, "accumulators exceed TMEM");
};

struct Params {
  void* C;
  const float* bias;
  float alpha;
  int M, N, K;
  long long ldc;
  int a_mn, b_mn;       // 1 = operand is MN-major in global memory
  int out_bf16;         // 1 = bf16 output, 0 = fp32
  int bias_mode;        // 0 none, 1 per-column (N), 2 per-row (M)
  int relu;
  int kb_per_split;     // k-blocks per split-K slice
  int mt, nt, splits;   // tile grid (the kernel is persistent: tiles are walked round-robin by the CTAs)
  int num_kb;           // total k-blocks (GEMM: ceil(K/64); conv: taps*chunks or pixel blocks)
  // implicit-GEMM convolution (operands gathered by TMA im2col loads, no col matrix in memory):
  //   conv_mode 1  fprop / dgrad : A = activation im2col tile [128 pixels x 64 ch] per (tap, chunk); B = weights [O][tap][C] 3-D tiled
  //   conv_mode 2  wgrad         : A = dy (MN-major tiled);  B = activation im2col boxes [64 pixels x 64 ch]; n-tiles = (tap, channel chunk)
  int conv_mode;
  int cHo, cWo, cS, cP, cKH, cKW, cCg, c_chunks;
  int atomic_out;       // 1 = fp32 atomicAdd (split-K)
  // second problem of the same shape run by the same launch (the two groups of an AlexNet-style grouped convolution): tiles
  // [0, per_group) belong to group 0 (maps a/b, C, bias), tiles [per_group, 2*per_group) to group 1 (maps a1/b1, C1, bias1)
  void* C1;
  const float* bias1;
  int groups;           // 1 or 2
  int group_m;          // tile raster: m-tiles per band (0 = plain m-fastest order); see tile_mn()
  int dbg;              // bottleneck probe (scripts/gemm_probe.py): 1 = skip A loads, 2 = skip B loads, 4 = skip the MMAs
  // reduce-scatter fused into the epilogue (fp32 wgrad outputs living in the symmetric gradient arena): element e of the G
  // region is owned by rank ((e >> 10) - rs_blo) / rs_per; every 16-byte vector is red.add-ed into the OWNER's G over NVLink
  // (rs_g[q] = rank q's G region as mapped here) instead of being stored locally.  rs_world == 0: off.
  int rs_world = 0, rs_rank = 0;
  unsigned rs_blo = 0, rs_per = 1;
  long long rs_e0 = 0;  // element index of C[0, 0] inside the G region
  float* rs_g[kMaxRanks] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // epilogue through the bulk copy engine (cp.async.bulk / cp.reduce.async.bulk, one 64–128 byte row segment per lane):
  // bit 0 plain stores, bit 1 split-K reductions, bit 2 reduce-scatter reductions over NVLink (TMPI_GEMM_BULK, see launch())
  int bulk = 0;
};

// owner-rank address of element e of the gradient region (no dynamic indexing of the kernel-parameter array)
__device__ __forceinline__ float* rs_addr(const Params& p, long long e, bool& local) {
  unsigned owner = ((unsigned)(e >> 10) - p.rs_blo) / p.rs_per;
  if (owner >= (unsigned)p.rs_world) owner = (unsigned)p.rs_world - 1u;
  float* base = p.rs_g[0];
#pragma unroll
  for (int q = 1; q < kMaxRanks; ++q) if (owner == (unsigned)q) base = p.rs_g[q];
  local = owner == (unsigned)p.rs_rank;
  return base + e;
}
__device__ __forceinline__ void red_add_sys_f32(float* addr, float v) {
  asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU.  The spin loop lives INSIDE the asm
// statement so the compiler sees straight-line code: a C++ loop around try_wait has a per-thread exit condition, which
// makes everything after it "potentially divergent" and pushes the producer / MMA loop indices out of the uniform
// registers (R2UR before every UTMALDG / UTCHMMA operand).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1, P2;\n\t"
      ".reg .u32 cnt;\n\t"
      "mov.u32 cnt, 0;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "add.u32 cnt, cnt, 1;\n\t"
      "setp.lt.u32 P2, cnt, 0x10000000;\n\t"
      "@P2 bra LAB_WAIT;\n\t"
      "trap;\n\t"
      "DONE:\n\t"
      "}"
      ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// im2col-mode TMA on an NHWC tensor (dims C,W,H,N): coordinates are the input position of the window's top-left corner
// (w = q*stride - pad, h = p*stride - pad) of the FIRST pixel of the tile; the filter tap goes in the 16-bit offsets.
// The unit then walks pixelsPerColumn output positions (W, then H, then N) and zero-fills padding / out-of-range pixels.
// (Semantics established with csrc/probe_im2col.cu on a B200.)
__device__ __forceinline__ void tma_load_im2col(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n,
                                                int off_w, int off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"((uint16_t)off_w), "h"((uint16_t)off_h) : "memory");
}
// Bulk (TMA engine) epilogue ops: one contiguous row segment shared → global per call; the reduce form adds fp32 into global
// (or peer-mapped) memory as ONE packet per segment instead of 16-byte vector atomics.
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_reduce_add_f32(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// one lane of a fully converged warp (elect.sync): the issuing thread of the TMA / MMA warps
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <bool TF32>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B, version 1.
// layout_type: 2 = SWIZZLE_128B (16-byte swizzle atoms), 1 = SWIZZLE_128B_BASE32B (32-byte atoms: the only layout the tensor
// core accepts for MN-major 32-bit (tf32) operands — cute::UMMA::Layout_MN_SW128_32B_Atom, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2u) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);            // start address      bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;       // leading byte off   bits [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;       // stride byte off    bits [32,46)
  d |= (uint64_t)1 << 46;                                 // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;                       // layout type
  return d;
}

// Tile raster.  Default: m fastest (all CTAs of a wave share the few B tiles — right for conv / FC where one operand is
// small).  For GEMMs with many tiles in both directions the wave is folded into bands of group_m m-tiles so the tiles that
// run concurrently form a near-square block and re-use both operands out of L2 instead of streaming one from HBM.
__device__ __forceinline__ void tile_mn(const Params& p, int rem, int& mti, int& nti) {
  if (p.group_m <= 0) { nti = rem / p.mt; mti = rem - nti * p.mt; return; }
  const int band_sz = p.group_m * p.nt;
  const int band = rem / band_sz, within = rem - band * band_sz;
  const int rows = min(p.group_m, p.mt - band * p.group_m);
  nti = within / rows; mti = band * p.group_m + (within - nti * rows);
}

// ------------------------------------------------------------------ the kernel
// Persistent: grid = min(#tiles, #SMs); CTA c processes tiles c, c+grid, … .  Three pipelines run concurrently:
//   TMA warp  → smem ring (full/empty mbarriers)           → MMA thread
//   MMA thread → TMEM accumulator A/B (tmem_full/tmem_empty) → epilogue warps
// so the epilogue of tile i overlaps the main loop of tile i+1 and all per-CTA setup (TMEM alloc, barrier init,
// descriptor fetch) is paid once per SM instead of once per tile.
template <typename T, int BN, int MT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_b0,
             const __grid_constant__ CUtensorMap tmap_a1, const __grid_constant__ CUtensorMap tmap_b1, const Params p) {
  using C = Cfg<T, BN, MT>;
  using E = Elem<T>;
  constexpr int BK = E::BK, UMMA_K = E::UMMA_K, ATOM = E::ATOM;
  constexpr int SUB_BYTES = BM * 128;              // one 128-row operand sub-tile of one k-block
  constexpr int TM = MT * BM;                      // rows of a CTA tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B needs 1024B alignment
  const uint32_t staging_base = smem_base + C::STAGES * C::STAGE_BYTES;
  const uint32_t bar_base = staging_base + C::STAGING_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + C::ACC_STAGES + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 2 * C::ACC_STAGES);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb_total = p.num_kb;
  const int tiles_mn = p.mt * p.nt;
  const int per_group = tiles_mn * p.splits;
  const int total_tiles = per_group * p.groups;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a0);
    tma_prefetch_desc(&tmap_b0);
    if (p.groups > 1) { tma_prefetch_desc(&tmap_a1); tma_prefetch_desc(&tmap_b1); }
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < C::ACC_STAGES; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), NUM_EPI_WARPS); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, C::TMEM_COLS); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The WHOLE warp walks the loops (so every index stays warp-uniform and lives in uniform registers — a loop body that
    // sits inside `if (lane == 0)` is compiled with per-thread registers plus an R2UR/ELECT dance in front of every
    // UTMALDG and cost ~450 cycles per k-block, more than the MMAs it feeds); one elected lane issues.  All per-k-block
    // index math is incremental: no divisions inside the k loop.
    const bool leader = elect_one();
    const bool skip_a = (p.dbg & 1) != 0, skip_b = (p.dbg & 2) != 0;
    const bool issue_a = leader && !skip_a, issue_b = leader && !skip_b;
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int grp = tile >= per_group ? 1 : 0;
      const int t2 = tile - grp * per_group;
      const CUtensorMap* const tma_a = grp ? &tmap_a1 : &tmap_a0;
      const CUtensorMap* const tma_b = grp ? &tmap_b1 : &tmap_b0;
      const int split = t2 / tiles_mn, rem = t2 - split * tiles_mn;
      int mti, nti;
      tile_mn(p, rem, mti, nti);
      const int m0 = mti * TM, n0 = nti * BN;
      const int kb0 = split * p.kb_per_split, kb1 = min(num_kb_total, kb0 + p.kb_per_split);
      // sub-tiles that start beyond M are not loaded at all (their MMAs chew on stale smem; the epilogue drops the rows)
      const int n_sub = (MT == 2 && m0 + BM < p.M) ? 2 : 1;
      if (p.conv_mode == 1) {
        // ---- conv fprop / dgrad: A = im2col box [128 pixels x 64 ch] of filter tap (r_, s_), channel chunk cc; B = weights
        const int hw = p.cHo * p.cWo;
        int img0[MT], bw0[MT], bh0[MT];
#pragma unroll
        for (int u = 0; u < MT; ++u) {
          const int mu = m0 + u * BM;
          img0[u] = mu / hw; const int r0_ = mu - img0[u] * hw; const int p0 = r0_ / p.cWo, q0 = r0_ - p0 * p.cWo;
          bw0[u] = q0 * p.cS - p.cP; bh0[u] = p0 * p.cS - p.cP;
        }
        int tap = kb0 / p.c_chunks, cc = kb0 - tap * p.c_chunks;
        int r_ = tap / p.cKW, s_ = tap - r_ * p.cKW;
        const uint32_t tx = (skip_a ? 0u : (uint32_t)(n_sub * SUB_BYTES)) + (skip_b ? 0u : (uint32_t)C::B_BYTES);
        const bool b_t = p.b_mn != 0;
        const int ntaps = p.cKH * p.cKW;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t fb = full_bar(stage);
          if (leader) mbar_expect_tx(fb, tx);
          if (issue_a) {
#pragma unroll
            for (int u = 0; u < MT; ++u)
              if (u < n_sub) tma_load_im2col(sa + u * SUB_BYTES, tma_a, fb, cc * BK, bw0[u], bh0[u], img0[u], s_, r_);
          }
          if (issue_b) {
            if (!b_t) {
              tma_load_3d(sa + C::A_BYTES, tma_b, fb, cc * BK, tap, n0);            // weights [n][tap][k]: K-major box
            } else {
              // dgrad reads the FORWARD filter [k = out-ch][tap][n = in-ch] in place: MN-major boxes {64 n, 1 tap, 64 k} of
              // the mirrored tap (no flipped / transposed copy of the weights)
#pragma unroll
              for (int j = 0; j < (BN >= ATOM ? BN / ATOM : 1); ++j)
                tma_load_3d(sa + C::A_BYTES + j * (BK * 128), tma_b, fb, n0 + ATOM * j, ntaps - 1 - tap, cc * BK);
            }
          }
          if (++cc == p.c_chunks) { cc = 0; ++tap; if (++s_ == p.cKW) { s_ = 0; ++r_; } }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      } else if (p.conv_mode == 2) {
        // ---- conv wgrad: A = dy (MN-major) [64 m x 64 pixels] x2;  B = BN/64 im2col boxes [64 pixels x 64 ch], one per
        //      (filter tap, channel chunk); the k loop walks pixels 64 at a time (W, then H, then N)
        constexpr int NBOX = (BN >= ATOM) ? BN / ATOM : 1;
        const int total_boxes = p.cKH * p.cKW * p.c_chunks;
        const int w_box0 = nti * NBOX;
        int bc[NBOX], bs[NBOX], br[NBOX];
        int nbox = 0;
#pragma unroll
        for (int j = 0; j < NBOX; ++j) {
          const int box = w_box0 + j;
          const int tapj = box / p.c_chunks, c64 = box - tapj * p.c_chunks;
          br[j] = tapj / p.cKW; bs[j] = tapj - br[j] * p.cKW; bc[j] = c64 * ATOM;
          if (box < total_boxes) ++nbox;
        }
        if (skip_b) nbox = 0;
        const uint32_t tx = (skip_a ? 0u : (uint32_t)C::A_BYTES) + (uint32_t)(nbox * (BK * 128));
        const int hw = p.cHo * p.cWo;
        int pix = kb0 * BK;
        int img = pix / hw; const int r2 = pix - img * hw; int pp = r2 / p.cWo, qq = r2 - pp * p.cWo;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          const uint32_t fb = full_bar(stage);
          if (leader) mbar_expect_tx(fb, tx);
          if (issue_a) {
#pragma unroll
            for (int j = 0; j < BM / ATOM; ++j) tma_load_2d(sa + j * (BK * 128), tma_a, fb, m0 + ATOM * j, pix);
          }
          if (issue_b) {
            const int cw = qq * p.cS - p.cP, ch = pp * p.cS - p.cP;
#pragma unroll
            for (int j = 0; j < NBOX; ++j)
              if (j < nbox) tma_load_im2col(sb + j * (BK * 128), tma_b, fb, bc[j], cw, ch, img, bs[j], br[j]);
          }
          pix += BK; qq += BK;
          while (qq >= p.cWo) { qq -= p.cWo; if (++pp == p.cHo) { pp = 0; ++img; } }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      } else {
        // ---- plain GEMM: K-major operands are one box {64 k, rows}; MN-major operands are 64-wide column boxes {64 mn, 64 k}
        const uint32_t tx = (skip_a ? 0u : (uint32_t)(n_sub * SUB_BYTES)) + (skip_b ? 0u : (uint32_t)C::B_BYTES);
        const bool a_mn = p.a_mn != 0, b_mn = p.b_mn != 0;
        int k0 = kb0 * BK;
        for (int kb = kb0; kb < kb1; ++kb, k0 += BK) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          const uint32_t fb = full_bar(stage);
          if (leader) mbar_expect_tx(fb, tx);
          if (issue_a) {
#pragma unroll
            for (int u = 0; u < MT; ++u) {
              if (u < n_sub) {
                if (!a_mn) {
                  tma_load_2d(sa + u * SUB_BYTES, tma_a, fb, k0, m0 + u * BM);
                } else {
#pragma unroll
                  for (int j = 0; j < BM / ATOM; ++j) tma_load_2d(sa + u * SUB_BYTES + j * (BK * 128), tma_a, fb, m0 + u * BM + ATOM * j, k0);
                }
              }
            }
          }
          if (issue_b) {
            if (!b_mn) {
              tma_load_2d(sb, tma_b, fb, k0, n0);
            } else {
#pragma unroll
              for (int j = 0; j < (BN >= ATOM ? BN / ATOM : 1); ++j) tma_load_2d(sb + j * (BK * 128), tma_b, fb, n0 + ATOM * j, k0);
            }
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the loop, one elected lane issues) =====================
    const bool leader = elect_one();
    const bool skip_mma = (p.dbg & 4) != 0;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16 (format 1) or tf32 (format 2), majors, N>>3, M>>4
    constexpr uint32_t kFmt = E::TF32 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    // K-major, SW128: 8-row groups 1024 B apart (SBO); advance 32 B per UMMA_K inside the 128 B row.
    // MN-major, SW128: 64-element MN atoms BK*128 B apart (LBO), 8-k-row groups 1024 B apart (SBO);
    //                  advance UMMA_K k-rows = UMMA_K * 128 B per MMA (2048 B bf16, 1024 B tf32).
    constexpr uint32_t kMnStep = (uint32_t)(UMMA_K * 128) >> 4;
    const uint32_t a_lbo = p.a_mn ? (uint32_t)(BK * 128) : 16u, a_step = p.a_mn ? kMnStep : 2u;
    const uint32_t b_lbo = p.b_mn ? (uint32_t)(BK * 128) : 16u, b_step = p.b_mn ? kMnStep : 2u;
    // tf32 MN-major: 32-byte swizzle atoms, 4 k-rows (512 B) per atom along K; everything else: 16-byte atoms, 8 rows (1024 B)
    const uint32_t a_lt = (E::TF32 && p.a_mn) ? 1u : 2u, b_lt = (E::TF32 && p.b_mn) ? 1u : 2u;
    const uint64_t adesc_base = make_smem_desc(smem_base, a_lbo, a_lt == 1u ? 512u : 1024u, a_lt);
    const uint64_t bdesc_base = make_smem_desc(smem_base + C::A_BYTES, b_lbo, b_lt == 1u ? 512u : 1024u, b_lt);
    int stage = 0; uint32_t phase = 0;
    int t = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
      const int t2 = tile >= per_group ? tile - per_group : tile;
      const int split = t2 / tiles_mn;
      const int kb0 = split * p.kb_per_split, kb1 = min(num_kb_total, kb0 + p.kb_per_split);
      const int acc = t & 1;
      mbar_wait(tmem_empty_bar(acc), (uint32_t)(((t >> 1) & 1) ^ 1));       // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * C::ACC_COLS);
      uint32_t accumulate = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (leader) {
          const uint64_t adesc0 = adesc_base + (uint64_t)((uint32_t)(stage * C::STAGE_BYTES) >> 4);
          const uint64_t bdesc0 = bdesc_base + (uint64_t)((uint32_t)(stage * C::STAGE_BYTES) >> 4);
          if (!skip_mma) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
#pragma unroll
              for (int u = 0; u < MT; ++u)             // the MT sub-tiles share the B descriptor
                umma_ss<E::TF32>(tmem_acc + (uint32_t)(u * BN), adesc0 + (uint64_t)(u * (SUB_BYTES >> 4)) + (uint64_t)(a_step * k),
                          bdesc0 + (uint64_t)(b_step * k), idesc, accumulate);
              accumulate = 1u;
            }
          }
          umma_commit(empty_bar(stage));             // frees the smem slot once these MMAs retire
        }
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
      if (leader) umma_commit(tmem_full_bar(acc));   // accumulator complete → epilogue
      __syncwarp();
    }
  } else {
    // ===================== epilogue: TMEM → registers → (per-warp smem transpose) → global =====================
    // 8 warps: warp w reads TMEM lane quarter q = w % 4 (hardware rule) and the column half (w - 2) / 4 of the tile, one
    // 32-column chunk at a time: tcgen05.ld → bias / ReLU → the warp's staging region (thread = row) → read back with
    // thread = 16-byte vector so every global store / red covers whole 64–128 B row segments.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int COLS_PER_WARP = (BN >= 64) ? BN / 2 : BN;          // BN = 32: only the first four warps carry data
    constexpr int NCHUNK = COLS_PER_WARP / 32;
    const bool active = (BN >= 64) || half == 0;
    const int col0 = (BN >= 64) ? half * COLS_PER_WARP : 0;
    const int row = 32 * q + lane;                 // row inside a 128-row sub-tile
    const int esz = p.out_bf16 ? 2 : 4;
    const uint32_t pitch = (uint32_t)(32 * esz + 16);
    uint8_t* wstage = smem_raw + (staging_base - smem_u32(smem_raw)) + (size_t)(warp - 2) * 32 * C::STAGING_ROW;
    const int vec_per_row = (32 * esz) / 16;       // 16-byte vectors per chunk row: 8 (fp32) or 4 (bf16)
    const int rows_per_it = 32 / vec_per_row;
    const int lr = lane / vec_per_row, lv = lane % vec_per_row;
    const bool ld_ok = (((long long)p.ldc * esz) % 16) == 0;
    int t = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
      const int grp = tile >= per_group ? 1 : 0;
      const int rem = (tile - grp * per_group) % tiles_mn;
      uint8_t* const Cg_ptr = reinterpret_cast<uint8_t*>(grp ? p.C1 : p.C);
      const float* const bias_ptr = grp ? p.bias1 : p.bias;
      int mti, nti;
      tile_mn(p, rem, mti, nti);
      const int m0 = mti * TM;
      const int n0 = nti * BN, n_end = p.N;          // global column of tile column cc is n0 + cc (GEMM / conv fprop)
      // wgrad (conv_mode 2): every ATOM-column box of the tile is one (filter tap, ATOM-channel chunk) and lands at column
      // tap*Cg + chunk*ATOM of dW — resolved per 32-column chunk below (a chunk never straddles two boxes)
      // bias prefetch BEFORE waiting for the accumulator (hidden behind the main loop): per-column bias — lane j holds the
      // bias of column (chunk base + j), broadcast later with shuffles; per-row bias — one value per sub-tile row.
      float bias_m0 = 0.f, bias_m1 = 0.f;
      if (p.bias_mode == 2) {
        if (m0 + row < p.M) bias_m0 = __ldg(bias_ptr + m0 + row);
        if (MT == 2 && m0 + BM + row < p.M) bias_m1 = __ldg(bias_ptr + m0 + BM + row);
      }
      float bias_c0 = 0.f, bias_c1 = 0.f;
      if (p.bias_mode == 1 && active) {
        const int cb = n0 + col0 + lane;
        if (cb < n_end) bias_c0 = __ldg(bias_ptr + cb);
        if (NCHUNK > 1 && cb + 32 < n_end) bias_c1 = __ldg(bias_ptr + cb + 32);
      }
      const int acc = t & 1;
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * C::ACC_COLS);
      mbar_wait(tmem_full_bar(acc), (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      if (!active) {                               // nothing to read: just hand the accumulator back
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
        continue;
      }
#pragma unroll 1
      for (int u = 0; u < MT; ++u) {
        const int mbase = m0 + u * BM;               // first row of this sub-tile
        const int m = mbase + row;
        const bool m_ok = m < p.M;
#pragma unroll 1
        for (int c = 0; c < NCHUNK; ++c) {
          uint32_t r[32];
          const int cc = col0 + 32 * c;              // column offset inside the tile
          __syncwarp();                              // tcgen05.ld is .sync.aligned: reconverge first
          tmem_ld32(tmem_acc + ((uint32_t)(32 * q) << 16) + (uint32_t)(u * BN + cc), r);
          tmem_ld_wait();
          if (u == MT - 1 && c == NCHUNK - 1) {
            // all of this warp's TMEM reads for the tile are done → hand the accumulator back to the MMA thread
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
          }
          if (mbase >= p.M) continue;                // warp-uniform: whole sub-tile out of range
          int nb = n0 + cc, n_end_c = n_end;
          if (p.conv_mode == 2) {
            const int box = nti * (BN >= ATOM ? BN / ATOM : 1) + cc / ATOM;
            if (box < p.cKH * p.cKW * p.c_chunks) {
              const int tap = box / p.c_chunks, cch = box - tap * p.c_chunks;
              nb = tap * p.cCg + cch * ATOM + (cc % ATOM);
              n_end_c = tap * p.cCg + min(p.cCg, cch * ATOM + ATOM);
            } else {
              nb = 0; n_end_c = 0;
            }
          }
          const float bias_sel = (c == 0) ? bias_c0 : bias_c1;
          float v[32];
          if (p.bias_mode == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), p.alpha, __shfl_sync(0xffffffffu, bias_sel, j));
          } else {
            const float bm = (u == 0) ? bias_m0 : bias_m1;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), p.alpha, bm);
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          // Fast path: chunk fully in range and 16-byte aligned → transposed through smem, coalesced 16 B row-segment
          // stores (plain, or vector reductions red.global.add.v4.f32 for split-K).
          const bool staged = (nb + 32 <= n_end_c) && ld_ok && (((reinterpret_cast<uintptr_t>(Cg_ptr) + (long long)nb * esz) % 16) == 0);
          if (staged) {
            uint8_t* dst = wstage + (size_t)lane * pitch;
            if (p.out_bf16) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) { bf16x8 pk = pack8(v + j); *reinterpret_cast<bf16x8*>(dst + j * 2) = pk; }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            __syncwarp();
            const int bulk_kind = p.rs_world > 0 ? 4 : (p.atomic_out ? 2 : 1);
            if (p.bulk & bulk_kind) {
              // one row segment (32 columns) per lane through the bulk copy engine: the staging row this lane just wrote is
              // the source; generic-proxy writes are fenced into the async proxy first
              fence_proxy_async();
              const long long gm = (long long)mbase + row;
              if (gm < p.M) {
                const uint32_t ssrc = smem_u32(wstage + (size_t)lane * pitch);
                const uint32_t nbytes = (uint32_t)(32 * esz);
                if (p.rs_world > 0) {
                  bool local;
                  float* d = rs_addr(p, p.rs_e0 + gm * p.ldc + nb, local);
                  bulk_reduce_add_f32(d, ssrc, nbytes);
                } else if (p.atomic_out) {
                  bulk_reduce_add_f32(Cg_ptr + ((long long)gm * p.ldc + nb) * esz, ssrc, nbytes);
                } else {
                  bulk_store(Cg_ptr + ((long long)gm * p.ldc + nb) * esz, ssrc, nbytes);
                }
              }
              bulk_commit();
              bulk_wait_read();                                          // staging row may be overwritten by the next chunk
              __syncwarp();
              continue;
            }
            uint8_t* gbase = Cg_ptr + (long long)nb * esz + (long long)lv * 16;
            for (int r0 = 0; r0 < 32; r0 += rows_per_it) {
              const int rr = r0 + lr;                                  // row inside this warp's 32-row band
              const long long gm = (long long)mbase + 32 * q + rr;
              if (gm < p.M) {
                const uint4 val = *reinterpret_cast<const uint4*>(wstage + (size_t)rr * pitch + (size_t)lv * 16);
                uint8_t* gp = gbase + gm * p.ldc * esz;
                if (p.rs_world > 0) {
                  // fused reduce-scatter: this 16-byte vector of dW goes to the rank that owns it in the exchange
                  bool local;
                  float* d = rs_addr(p, p.rs_e0 + gm * p.ldc + nb + lv * 4, local);
                  if (local) {
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(__uint_as_float(val.x)),
                                 "f"(__uint_as_float(val.y)), "f"(__uint_as_float(val.z)), "f"(__uint_as_float(val.w)) : "memory");
                  } else {
                    // ONE 16-byte vector reduction per NVLink packet (the first version issued four scalar 4-byte atomics per
                    // vector and was 60 % slower than the separate exchange kernel)
                    asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(__uint_as_float(val.x)),
                                 "f"(__uint_as_float(val.y)), "f"(__uint_as_float(val.z)), "f"(__uint_as_float(val.w)) : "memory");
                  }
                } else if (p.atomic_out) {
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gp), "f"(__uint_as_float(val.x)),
                               "f"(__uint_as_float(val.y)), "f"(__uint_as_float(val.z)), "f"(__uint_as_float(val.w)) : "memory");
                } else {
                  *reinterpret_cast<uint4*>(gp) = val;
                }
              }
            }
            __syncwarp();                                              // staging region is free for the next chunk
            continue;
          }
          if (!m_ok || nb >= n_end_c) continue;
          const bool full = (nb + 32 <= n_end_c);
          if (p.rs_world > 0) {
            const long long e0 = p.rs_e0 + (long long)m * p.ldc + nb;
            for (int j = 0; j < 32; ++j) {
              if (full || nb + j < n_end_c) { bool local; red_add_sys_f32(rs_addr(p, e0 + j, local), v[j]); }
            }
          } else if (p.atomic_out) {
            float* dst = reinterpret_cast<float*>(Cg_ptr) + (long long)m * p.ldc + nb;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (full || nb + j < n_end_c) atomicAdd(dst + j, v[j]);
          } else if (p.out_bf16) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(Cg_ptr) + (long long)m * p.ldc + nb;
            if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) { bf16x8 pk = pack8(v + j); *reinterpret_cast<bf16x8*>(dst + j) = pk; }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (nb + j < n_end_c) dst[j] = f_to_bf16(v[j]);
            }
          } else {
            float* dst = reinterpret_cast<float*>(Cg_ptr) + (long long)m * p.ldc + nb;
            if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (nb + j < n_end_c) dst[j] = v[j];
            }
          }
        }
      }
    }
  }
  if (warp >= 2 && p.bulk) bulk_wait_all();                // bulk stores / reductions have landed before the kernel retires
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, C::TMEM_COLS); }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled)f;
    (void)cudaGetLastError();
  });
  if (!fn) throw std::runtime_error("tmpi_native: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  return fn;
}

// fp32 operands of the tf32 path are described to the TMA unit as TFLOAT32: the copy engine then rounds every element to the
// nearest tf32 value on its way into shared memory (the tensor core alone would truncate the low 13 mantissa bits, a biased
// error that does not average out over K).  TMPI_TF32_TMA_ROUND=0 falls back to raw fp32 bits.
static CUtensorMapDataType f32_map_type() {
  static const bool rnd = [] { const char* e = getenv("TMPI_TF32_TMA_ROUND"); return !(e && e[0] == '0'); }();
  return rnd ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
}

// 2-D tensor map (bf16 or fp32 elements): dims {inner, outer}, row pitch in bytes, box {128 bytes, box_outer}, 128B swizzle,
// zero OOB fill.
static CUtensorMap make_tmap(const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_outer, int esz = 2,
                             int mn_major = 0) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) throw std::runtime_error("tmpi_native: TMA operand base must be 16B aligned");
  if ((pitch_bytes & 15) != 0) throw std::runtime_error("tmpi_native: TMA operand row pitch must be a multiple of 16 bytes");
  using Key = std::tuple<const void*, uint64_t, uint64_t, uint64_t, uint32_t, int, int>;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  Key key{ptr, inner, outer, pitch_bytes, box_outer, esz, mn_major};
  // MN-major fp32 (tf32) operands must land in the 32-byte-atom swizzle (see make_smem_desc)
  const CUtensorMapSwizzle swz = (esz == 4 && mn_major) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = get_encode()(&m, esz == 4 ? f32_map_type() : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("tmpi_native: cuTensorMapEncodeTiled failed, code " + std::to_string((int)r));
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  return m;
}

static int g_dbg = 0;
static int g_bulk = -1;            // gemm_set_bulk(): -1 = TMPI_GEMM_BULK env (default 4: reduce-scatter reductions only)
static int bulk_default() {
  static const int v = [] { const char* e = getenv("TMPI_GEMM_BULK"); return e ? atoi(e) : 4; }();
  return v;
}

// Split-K factor for a persistent grid of `sms` CTAs walking equal-length tiles round-robin: minimise
// waves x (k-blocks per slice + per-tile overhead).  A plain ceil(sms / tiles) overshoots the machine by a few tiles
// and pays a whole second wave for them (conv2 wgrad: 13 tiles x 12 slices = 156 > 148).
static int choose_splits(int tiles, int num_kb, int sms) {
  // TMPI_DETERMINISTIC=1: never split K — split-K slices combine with fp32 red.global.add in arrival order, so weight gradients
  // differ in the last bits from run to run; without it every output element is produced by ONE CTA in a fixed k order
  if (deterministic_mode()) return 1;
  if (tiles >= sms || num_kb < 8) return 1;
  const int kTileOverheadKb = 4;                    // pipeline fill + accumulator hand-off, in k-block units
  int best = 1;
  long long best_cost = -1;
  const int max_s = std::min(num_kb / 4, 4 * sms);
  for (int s = 1; s <= max_s; ++s) {
    const int kb_per = (num_kb + s - 1) / s;
    const int s_eff = (num_kb + kb_per - 1) / kb_per;
    const long long waves = ((long long)tiles * s_eff + sms - 1) / sms;
    const long long cost = waves * (kb_per + kTileOverheadKb) * 64 + s_eff;     // tie-break: fewer slices (less atomic traffic)
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s_eff; }
  }
  return best;
}

template <typename T, int BN, int MT>
static void launch(const CUtensorMap& ta, const CUtensorMap& tb, Params& p, int splits, cudaStream_t st,
                   const CUtensorMap* ta1 = nullptr, const CUtensorMap* tb1 = nullptr) {
  using C = Cfg<T, BN, MT>;
  p.dbg = g_dbg;
  p.bulk = g_bulk < 0 ? bulk_default() : g_bulk;
  if (!ta1) { p.groups = 1; p.C1 = nullptr; p.bias1 = nullptr; }
  static bool attr_set = false;
  if (!attr_set) {
    check_cuda(cudaFuncSetAttribute(gemm_tcgen05<T, BN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES), "gemm smem attr");
    attr_set = true;
  }
  if (MT == 2 && sizeof(T) == 2 && !p.out_bf16) throw std::runtime_error("tmpi_native: 256-row bf16 GEMM tiles are bf16-output only");
  if (sizeof(T) == 4 && p.out_bf16) throw std::runtime_error("tmpi_native: the tf32 path stores fp32");
  const long long total = (long long)p.mt * p.nt * splits * p.groups;
  const int grid = (int)std::min<long long>(total, (long long)sm_count());
  gemm_tcgen05<T, BN, MT><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(ta, tb, ta1 ? *ta1 : ta, tb1 ? *tb1 : tb, p);
  count_launch();
  TMPI_CHECK_LAUNCH("gemm_tcgen05"); ::tmpi::check_capture(st, "gemm_tcgen05");
}

// 256-row tiles (MT = 2) when the output is bf16 (no split-K) and the wave arithmetic favours them: a tall tile costs ~1.7x a
// 128-row tile (measured: conv3 fprop 21.5 us vs 12.8 us per wave — the B tile and the per-k-block overhead are shared), so
// it wins unless halving the tile count wastes most of a wave (conv3 dgrad: 170 tall tiles = 2 waves vs 338 = 3 short ones).
static bool use_tall_tiles(long long M, int nt, int eligible, int sms) {
  static const bool enabled = [] { const char* e = getenv("TMPI_GEMM_TALL"); return !(e && e[0] == '0'); }();
  if (!enabled || !eligible || M < 2 * BM) return false;
  const long long t1 = ((M + BM - 1) / BM) * nt, t2 = ((M + 2 * BM - 1) / (2 * BM)) * nt;
  const long long w1 = (t1 + sms - 1) / sms, w2 = (t2 + sms - 1) / sms;
  return w2 * 17 <= w1 * 10;
}

}  // namespace gemm

void gemm_set_debug(int flags) { gemm::g_dbg = flags; }
void gemm_set_bulk(int mask) { gemm::g_bulk = mask; }

// ---- reduce-scatter epilogue registry (see api.h)
namespace gemm {
struct RsRange { const char* lo; const char* hi; long long blo, per; };
static int g_rs_world = 0, g_rs_rank = 0;
static float* g_rs_peer[kMaxRanks] = {};
static const char* g_rs_local = nullptr;
static std::vector<RsRange> g_rs_ranges;
static std::mutex g_rs_mu;
// fills the rs_* fields of p when C lies in a registered tensor; returns true if the epilogue will reduce-scatter
static bool rs_lookup(const void* C, Params& p) {
  std::lock_guard<std::mutex> lk(g_rs_mu);
  if (g_rs_world < 2) return false;
  const char* c = reinterpret_cast<const char*>(C);
  for (const RsRange& r : g_rs_ranges) {
    if (c >= r.lo && c < r.hi) {
      p.rs_world = g_rs_world; p.rs_rank = g_rs_rank;
      p.rs_blo = (unsigned)r.blo; p.rs_per = (unsigned)std::max<long long>(1, r.per);
      p.rs_e0 = (long long)((c - g_rs_local) / 4);
      for (int q = 0; q < kMaxRanks; ++q) p.rs_g[q] = q < g_rs_world ? g_rs_peer[q] : nullptr;
      return true;
    }
  }
  return false;
}
}  // namespace gemm

void gemm_rs_configure(int world, const void* const* peer_g, const void* local_g) {
  std::lock_guard<std::mutex> lk(gemm::g_rs_mu);
  if (world > kMaxRanks) throw std::runtime_error("gemm_rs_configure: too many ranks");
  gemm::g_rs_world = world;
  gemm::g_rs_local = reinterpret_cast<const char*>(local_g);
  gemm::g_rs_rank = 0;
  for (int q = 0; q < world; ++q) {
    gemm::g_rs_peer[q] = reinterpret_cast<float*>(const_cast<void*>(peer_g[q]));
    if (peer_g[q] == local_g) gemm::g_rs_rank = q;
  }
  gemm::g_rs_ranges.clear();
}
void gemm_rs_add_range(const void* c_lo, const void* c_hi, long long blo, long long per) {
  std::lock_guard<std::mutex> lk(gemm::g_rs_mu);
  gemm::g_rs_ranges.push_back({reinterpret_cast<const char*>(c_lo), reinterpret_cast<const char*>(c_hi), blo, per});
}
void gemm_rs_clear() {
  std::lock_guard<std::mutex> lk(gemm::g_rs_mu);
  gemm::g_rs_world = 0; gemm::g_rs_ranges.clear();
}

// host-side planning helpers, exported so the wave arithmetic can be unit-tested without a GPU
int gemm_plan_splits(int tiles, int num_kb, int sms) { return gemm::choose_splits(tiles, num_kb, sms); }
int gemm_plan_tall(long long M, int nt, int out_bf16, int sms) { return gemm::use_tall_tiles(M, nt, out_bf16, sms) ? 1 : 0; }

// C[M,N] (ldc) = alpha * op(A) op(B) + bias, optional ReLU.
//   a_mn == 0: A is [M, K] with row pitch lda (elements);  a_mn == 1: A is [K, M] with row pitch lda.
//   b_mn == 0: B is [N, K] with row pitch ldb;             b_mn == 1: B is [K, N] with row pitch ldb.
//   bn_hint: 0 = auto, else 32/64/128.  splitk: 0 = auto, 1 = none, >1 = forced (fp32 output only, no bias/relu).
//   T = __nv_bfloat16: bf16 operands (kind::f16), bf16 or fp32 output.  T = float: fp32 operands (kind::tf32), fp32 output.
namespace gemm {
template <typename T>
static void gemm_host(const void* A, const void* B, void* C, const float* bias, int M, int N, int K, long long lda, long long ldb,
                      long long ldc, int a_mn, int b_mn, int out_bf16, int bias_mode, int relu, float alpha, int bn_hint, int splitk,
                      cudaStream_t st) {
  using E = Elem<T>;
  constexpr int BK = E::BK, ATOM = E::ATOM, ESZ = E::ESZ;
  if (M <= 0 || N <= 0 || K <= 0) return;
  const int sms = sm_count();
  const int mt = (M + BM - 1) / BM;
  const bool fused_epi = bias_mode != 0 || relu;
  const bool can_split = (!out_bf16) && !fused_epi;
  int BN = bn_hint;
  if (BN == 0) {
    BN = 128;
    // Narrow tiles only buy parallelism when split-K cannot (fused bias/ReLU or bf16 output): with split-K available the
    // wide tile is always better — every extra n-tile re-reads the whole A operand through L2.
    if (!(can_split && splitk != 1)) {
      if (mt * ((N + 127) / 128) < sms && N >= 64) BN = 64;
      if (BN == 64 && mt * ((N + 63) / 64) < sms && !b_mn && N >= 32) BN = 32;
    } else if (N <= 64) {
      BN = 64;
    }
  }
  if (b_mn && BN < 64) BN = 64;
  const int nt = (N + BN - 1) / BN;
  const int num_kb = (K + BK - 1) / BK;
  int splits = 1;
  if (splitk > 1 && can_split) splits = splitk;
  else if (splitk == 0 && can_split) splits = choose_splits(mt * nt, num_kb, sms);
  int kb_per = (num_kb + splits - 1) / splits;
  splits = (num_kb + kb_per - 1) / kb_per;          // every slice owns >= 1 k-block

  Params p;
  p.C = C; p.bias = bias; p.alpha = alpha; p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.a_mn = a_mn; p.b_mn = b_mn;
  p.out_bf16 = out_bf16; p.bias_mode = bias ? bias_mode : 0; p.relu = relu; p.kb_per_split = kb_per; p.atomic_out = splits > 1;
  // tall tiles: bf16 path — bf16 outputs (fprop / dgrad); tf32 path — any un-split output (activations are fp32 there)
  const int tall_ok = ESZ == 2 ? out_bf16 : 1;
  const bool tall = splits == 1 && BN >= 64 && use_tall_tiles(M, nt, tall_ok, sms);
  p.mt = tall ? (M + 2 * BM - 1) / (2 * BM) : mt; p.nt = nt; p.splits = splits; p.num_kb = num_kb; p.conv_mode = 0;
  p.group_m = (p.mt > 12 && nt > 12) ? (tall ? 8 : 12) : 0;
  p.cHo = p.cWo = p.cS = p.cP = p.cKH = p.cKW = p.cCg = p.c_chunks = 0;
  // wgrad outputs registered for the fused reduce-scatter: every vector is red.add-ed into its owner's G (the exchange kernel
  // clears G after consuming it, so there is no memset here — a memset would race with the peers' adds)
  const bool rs = (!out_bf16) && bias_mode == 0 && !relu && alpha == 1.f && (ldc % 4) == 0 && rs_lookup(C, p);
  if (rs) {
    p.atomic_out = 1;
  } else if (splits > 1) {
    // split-K accumulates with fp32 atomics: clear the (possibly strided) output first
    check_cuda(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st), "gemm split-K memset");
  }
  CUtensorMap ta = a_mn ? make_tmap(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda * ESZ, (uint32_t)BK, ESZ, 1)
                        : make_tmap(A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * ESZ, (uint32_t)BM, ESZ, 0);
  CUtensorMap tb = b_mn ? make_tmap(B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * ESZ, (uint32_t)BK, ESZ, 1)
                        : make_tmap(B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * ESZ, (uint32_t)BN, ESZ, 0);
  (void)ATOM;
  if (tall) { if (BN == 128) launch<T, 128, 2>(ta, tb, p, splits, st); else launch<T, 64, 2>(ta, tb, p, splits, st); }
  else if (BN == 128) launch<T, 128, 1>(ta, tb, p, splits, st);
  else if (BN == 64) launch<T, 64, 1>(ta, tb, p, splits, st);
  else launch<T, 32, 1>(ta, tb, p, splits, st);
}
}  // namespace gemm

void gemm_bf16(const void* A, const void* B, void* C, const float* bias, int M, int N, int K, long long lda, long long ldb,
               long long ldc, int a_mn, int b_mn, int out_bf16, int bias_mode, int relu, float alpha, int bn_hint, int splitk,
               cudaStream_t st, int tf32) {
  if (tf32) gemm::gemm_host<float>(A, B, C, bias, M, N, K, lda, ldb, ldc, a_mn, b_mn, 0, bias_mode, relu, alpha, bn_hint, splitk, st);
  else gemm::gemm_host<__nv_bfloat16>(A, B, C, bias, M, N, K, lda, ldb, ldc, a_mn, b_mn, out_bf16, bias_mode, relu, alpha, bn_hint, splitk, st);
}

// ------------------------------------------------------------------ implicit-GEMM convolution (TMA im2col)
namespace gemm {
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                     const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeIm2col get_encode_im2col() {
  static PFN_encodeIm2col fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeIm2col)f;
    (void)cudaGetLastError();
  });
  if (!fn) throw std::runtime_error("tmpi_native: cuTensorMapEncodeIm2col unavailable");
  return fn;
}

// NHWC activation (channel slice [c_off, c_off+Cg) of a tensor with Ctot channels) as an im2col tensor map:
// box = pixels x 128 bytes of channels (64 bf16 / 32 fp32), 128 B swizzle, zero fill for padding / out-of-range pixels /
// channels >= Cg.
static CUtensorMap make_im2col_map(const void* x, int N, int H, int W, int Ctot, int c_off, int Cg, int KH, int KW, int S, int P, int pixels,
                                   int esz, int mn_major) {
  const char* base = reinterpret_cast<const char*>(x) + (size_t)c_off * esz;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((Ctot * esz) % 16) != 0) throw std::runtime_error("tmpi_native: im2col operand must be 16B aligned");
  using Key = std::tuple<const void*, int, int, int, int, int, int, int, int, int, int, int, int>;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  Key key{base, N, H, W, Ctot, Cg, KH, KW, S, P, pixels, esz, mn_major};
  const CUtensorMapSwizzle swz = (esz == 4 && mn_major) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)Cg, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)Ctot * esz, (cuuint64_t)W * Ctot * esz, (cuuint64_t)H * W * Ctot * esz};
  int lower[2] = {-P, -P};
  int upper[2] = {P - (KW - 1), P - (KH - 1)};
  cuuint32_t estr[4] = {1u, (cuuint32_t)S, (cuuint32_t)S, 1u};
  CUresult r = get_encode_im2col()(&m, esz == 4 ? f32_map_type() : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(base),
                                   dims, strides, lower, upper, (cuuint32_t)(128 / esz), (cuuint32_t)pixels, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("tmpi_native: cuTensorMapEncodeIm2col failed, code " + std::to_string((int)r));
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  return m;
}

// weights [O][KH*KW][Cg] (contiguous) as a 3-D tiled map, box {box_c ch, 1 tap, box_o out-channels}
static CUtensorMap make_weight_map(const void* w, int O, int taps, int Cg, int box_o, int box_c, int esz, int mn_major) {
  if ((reinterpret_cast<uintptr_t>(w) & 15) != 0 || ((Cg * esz) % 16) != 0) throw std::runtime_error("tmpi_native: conv weights must be 16B aligned, C * esz % 16 == 0");
  using Key = std::tuple<const void*, int, int, int, int, int, int, int>;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  Key key{w, O, taps, Cg, box_o, box_c, esz, mn_major};
  const CUtensorMapSwizzle swz = (esz == 4 && mn_major) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)Cg, (cuuint64_t)taps, (cuuint64_t)O};
  cuuint64_t strides[2] = {(cuuint64_t)Cg * esz, (cuuint64_t)taps * Cg * esz};
  cuuint32_t box[3] = {(cuuint32_t)box_c, 1u, (cuuint32_t)box_o};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = get_encode()(&m, esz == 4 ? f32_map_type() : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("tmpi_native: cuTensorMapEncodeTiled(3D weights) failed, code " + std::to_string((int)r));
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  return m;
}

// y[N*Ho*Wo, O] (ld = ldc) = relu(conv(x[.., c_off:c_off+Cg], w[O][KH][KW][Cg]) + bias)   — no col matrix in memory
// dgrad = 1: the same kernel computes the input gradient — x is dy, (Cg, O) are (#dy channels, #dx channels) and w is the
// FORWARD filter [Cg][KH][KW][O], read mirrored and transposed by the TMA loads (stride-1 convolutions only).
// ngroups = 2: both groups of a grouped convolution in ONE persistent launch (group g reads channel slice c_off[g] of x,
// filter w[g], writes y[g] / adds bias[g]) — their tiles fill the machine together instead of two under-filled waves.
template <typename T>
static void conv_fprop_groups(int ngroups, const void* x, const int* c_off, const void* const* w, void* const* y, const float* const* bias,
                              int N, int H, int W, int Ctot, int Cg, int KH, int KW, int Ho, int Wo, int S, int P, int O, long long ldc,
                              int relu, int out_bf16, int dgrad, cudaStream_t st) {
  using E = Elem<T>;
  constexpr int BK = E::BK, ATOM = E::ATOM, ESZ = E::ESZ;
  const long long M = (long long)N * Ho * Wo;
  if (M <= 0 || O <= 0) return;
  if (M >= (1LL << 31)) throw std::runtime_error("conv_fprop: too many output pixels");
  if (ESZ == 4) out_bf16 = 0;
  const int BN = O > 64 ? 128 : 64;
  Params p;
  p.C = y[0]; p.bias = bias[0]; p.alpha = 1.f; p.M = (int)M; p.N = O; p.K = KH * KW * Cg; p.ldc = ldc; p.a_mn = 0; p.b_mn = dgrad ? 1 : 0;
  p.groups = ngroups; p.C1 = ngroups > 1 ? y[1] : nullptr; p.bias1 = ngroups > 1 ? bias[1] : nullptr;
  p.out_bf16 = out_bf16; p.bias_mode = bias[0] ? 1 : 0; p.relu = relu; p.atomic_out = 0;
  if (ngroups > 1 && ((bias[0] == nullptr) != (bias[1] == nullptr))) throw std::runtime_error("conv_fprop: both groups need a bias or none");
  if (dgrad && (S != 1 || ((O * ESZ) % 16) != 0)) throw std::runtime_error("conv dgrad through the fprop kernel needs stride 1 and 16-byte channel rows");
  p.nt = (O + BN - 1) / BN; p.splits = 1;
  const bool tall = use_tall_tiles(M, p.nt * ngroups, ESZ == 2 ? out_bf16 : 1, sm_count());
  p.mt = tall ? (int)((M + 2 * BM - 1) / (2 * BM)) : (int)((M + BM - 1) / BM);
  p.group_m = 0;
  p.conv_mode = 1; p.cHo = Ho; p.cWo = Wo; p.cS = S; p.cP = P; p.cKH = KH; p.cKW = KW; p.cCg = Cg; p.c_chunks = (Cg + BK - 1) / BK;
  p.num_kb = KH * KW * p.c_chunks; p.kb_per_split = p.num_kb;
  CUtensorMap ta[2], tb[2];
  for (int g = 0; g < ngroups; ++g) {
    ta[g] = make_im2col_map(x, N, H, W, Ctot, c_off[g], Cg, KH, KW, S, P, BM, ESZ, 0);                     // K-major (channels = K)
    tb[g] = dgrad ? make_weight_map(w[g], Cg, KH * KW, O, BK, ATOM, ESZ, 1) : make_weight_map(w[g], O, KH * KW, Cg, BN, BK, ESZ, 0);
  }
  const CUtensorMap* a1 = ngroups > 1 ? &ta[1] : nullptr;
  const CUtensorMap* b1 = ngroups > 1 ? &tb[1] : nullptr;
  if (tall) { if (BN == 128) launch<T, 128, 2>(ta[0], tb[0], p, 1, st, a1, b1); else launch<T, 64, 2>(ta[0], tb[0], p, 1, st, a1, b1); }
  else if (BN == 128) launch<T, 128, 1>(ta[0], tb[0], p, 1, st, a1, b1);
  else launch<T, 64, 1>(ta[0], tb[0], p, 1, st, a1, b1);
}

// dw[O][KH*KW][Cg] (fp32, contiguous) = sum over pixels dy[pix, o] * im2col(x)[pix, (tap, c)]   (dy: [M, O], row pitch ldy)
// ngroups = 2: both groups in one launch (dy[g] = the group's channel slice of the output gradient, x slice c_off[g], dw[g]).
template <typename T>
static void conv_wgrad_groups(int ngroups, const void* const* dy, const void* x, void* const* dw, const int* c_off, int N, int H, int W,
                              int Ctot, int Cg, int KH, int KW, int Ho, int Wo, int S, int P, int O, long long ldy, cudaStream_t st) {
  using E = Elem<T>;
  constexpr int BK = E::BK, ATOM = E::ATOM, ESZ = E::ESZ;
  const long long M = (long long)N * Ho * Wo;
  if (M <= 0 || O <= 0) return;
  if (M >= (1LL << 31)) throw std::runtime_error("conv_wgrad: too many output pixels");
  const int sms = sm_count();
  const int BN = 128;                                  // BN / ATOM (tap, channel-chunk) boxes per n-tile: fewer re-reads of dy
  Params p;
  p.C = dw[0]; p.bias = nullptr; p.alpha = 1.f; p.M = O; p.N = KH * KW * Cg; p.K = (int)M; p.ldc = (long long)KH * KW * Cg; p.a_mn = 1; p.b_mn = 1;
  p.groups = ngroups; p.C1 = ngroups > 1 ? dw[1] : nullptr; p.bias1 = nullptr;
  p.out_bf16 = 0; p.bias_mode = 0; p.relu = 0;
  p.group_m = 0;
  p.conv_mode = 2; p.cHo = Ho; p.cWo = Wo; p.cS = S; p.cP = P; p.cKH = KH; p.cKW = KW; p.cCg = Cg; p.c_chunks = (Cg + ATOM - 1) / ATOM;
  p.mt = (O + BM - 1) / BM; p.nt = (KH * KW * p.c_chunks + BN / ATOM - 1) / (BN / ATOM);
  p.num_kb = (int)((M + BK - 1) / BK);
  int splits = choose_splits(p.mt * p.nt * ngroups, p.num_kb, sms);
  p.kb_per_split = (p.num_kb + splits - 1) / splits;
  splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.splits = splits; p.atomic_out = splits > 1;
  CUtensorMap ta[2], tb[2];
  for (int g = 0; g < ngroups; ++g) {
    if (splits > 1) check_cuda(cudaMemsetAsync(dw[g], 0, (size_t)O * KH * KW * Cg * 4, st), "conv_wgrad memset");
    ta[g] = make_tmap(dy[g], (uint64_t)O, (uint64_t)M, (uint64_t)ldy * ESZ, (uint32_t)BK, ESZ, 1);          // both operands MN-major
    tb[g] = make_im2col_map(x, N, H, W, Ctot, c_off[g], Cg, KH, KW, S, P, BK, ESZ, 1);
  }
  launch<T, 128, 1>(ta[0], tb[0], p, splits, st, ngroups > 1 ? &ta[1] : nullptr, ngroups > 1 ? &tb[1] : nullptr);
}
}  // namespace gemm

void conv_fprop_bf16(const void* x, const void* w, void* y, const float* bias, int N, int H, int W, int Ctot, int c_off, int Cg, int KH,
                     int KW, int Ho, int Wo, int S, int P, int O, long long ldc, int relu, int out_bf16, int dgrad, cudaStream_t st, int tf32) {
  const void* ws[1] = {w}; void* ys[1] = {y}; const float* bs[1] = {bias};
  if (tf32) gemm::conv_fprop_groups<float>(1, x, &c_off, ws, ys, bs, N, H, W, Ctot, Cg, KH, KW, Ho, Wo, S, P, O, ldc, relu, 0, dgrad, st);
  else gemm::conv_fprop_groups<__nv_bfloat16>(1, x, &c_off, ws, ys, bs, N, H, W, Ctot, Cg, KH, KW, Ho, Wo, S, P, O, ldc, relu, out_bf16, dgrad, st);
}

void conv_fprop2_bf16(const void* x, const void* w0, const void* w1, void* y0, void* y1, const float* bias0, const float* bias1, int N, int H,
                      int W, int Ctot, int c_off0, int c_off1, int Cg, int KH, int KW, int Ho, int Wo, int S, int P, int O, long long ldc,
                      int relu, int out_bf16, int dgrad, cudaStream_t st, int tf32) {
  const int co[2] = {c_off0, c_off1}; const void* ws[2] = {w0, w1}; void* ys[2] = {y0, y1}; const float* bs[2] = {bias0, bias1};
  if (tf32) gemm::conv_fprop_groups<float>(2, x, co, ws, ys, bs, N, H, W, Ctot, Cg, KH, KW, Ho, Wo, S, P, O, ldc, relu, 0, dgrad, st);
  else gemm::conv_fprop_groups<__nv_bfloat16>(2, x, co, ws, ys, bs, N, H, W, Ctot, Cg, KH, KW, Ho, Wo, S, P, O, ldc, relu, out_bf16, dgrad, st);
}

void conv_wgrad_bf16(const void* dy, const void* x, void* dw, int N, int H, int W, int Ctot, int c_off, int Cg, int KH, int KW, int Ho,
                     int Wo, int S, int P, int O, long long ldy, cudaStream_t st, int tf32) {
  const void* dys[1] = {dy}; void* dws[1] = {dw};
  if (tf32) gemm::conv_wgrad_groups<float>(1, dys, x, dws, &c_off, N, H, W, Ctot, Cg, KH, KW, Ho, Wo, S, P, O, ldy, st);
  else gemm::conv_wgrad_groups<__nv_bfloat16>(1, dys, x, dws, &c_off, N, H, W, Ctot, Cg, KH, KW, Ho, Wo, S, P, O, ldy, st);
}

void conv_wgrad2_bf16(const void* dy0, const void* dy1, const void* x, void* dw0, void* dw1, int N, int H, int W, int Ctot, int c_off0,
                      int c_off1, int Cg, int KH, int KW, int Ho, int Wo, int S, int P, int O, long long ldy, cudaStream_t st, int tf32) {
  const void* dys[2] = {dy0, dy1}; void* dws[2] = {dw0, dw1}; const int co[2] = {c_off0, c_off1};
  if (tf32) gemm::conv_wgrad_groups<float>(2, dys, x, dws, co, N, H, W, Ctot, Cg, KH, KW, Ho, Wo, S, P, O, ldy, st);
  else gemm::conv_wgrad_groups<__nv_bfloat16>(2, dys, x, dws, co, N, H, W, Ctot, Cg, KH, KW, Ho, Wo, S, P, O, ldy, st);
}

}  // namespace tmpi
