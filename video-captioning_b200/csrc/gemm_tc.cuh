// bf16 tensor-core GEMM for sm_100a: TMA -> 128B-swizzled shared memory -> tcgen05.mma -> TMEM ->
// tcgen05.ld -> fused epilogue -> swizzled shared-memory staging -> TMA store.  Hand-written PTX.
//
//   C[M,N] = A[M,K] . W[N,K]^T   A, W bf16 K-major; fp32 accumulation in TMEM.
//
// Warp roles (192 threads):  warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA
// issuer (one elected lane), warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4, one output row per
// thread).  kStages-deep smem ring with full/empty mbarriers; tcgen05.commit releases smem slots and
// signals the epilogue.  One 128 x BN output tile per CTA, BK = 64 (one 128-byte swizzle atom per row).
//
// Epilogue I/O never touches global memory with row-per-thread accesses (32 sectors per request; the
// round-1a profile showed that pattern and the exposed bias-load latency costing 5-10x the MMA time):
//   * bias is staged in shared memory while the main loop runs,
//   * the LSTM epilogue's inputs (all-timestep input-projection tile, previous cell state) are TMA-loaded
//     into swizzled shared memory at kernel start,
//   * outputs are written as 16-byte chunks into 128B-swizzled staging boxes (bank-conflict free for
//     row-per-thread writers) and leave through cp.async.bulk.tensor stores, which also clip M/N tails.
//
// Descriptor bit layouts follow the PTX ISA tcgen05 "shared memory descriptor" / "instruction
// descriptor" tables (cross-checked against cute/arch/mma_sm100_desc.hpp in the image).
//
// Kernels in this file:
//   gemm_tc_kernel             one 128 x BN tile per CTA (small GEMMs: query / context projections; masked encoder steps);
//                              the 128x128 store variant can take its A rows over tile row by tile row from a producer
//                              LSTM GEMM that is still running (GemmArgs::sync_wait)
//   gemm_tc_persistent_kernel  one CTA per SM loops over 128 x 256 tiles, two TMEM accumulators (epilogue of tile i overlaps
//                              the main loop of tile i+1); store / fused-LSTM / vocabulary-statistics epilogues; kind::f16 or
//                              kind::tf32 operands; MC = CTA pairs (clusters of 2, tcgen05 cta_group::2, 256 x 256 pair tiles,
//                              4-5 stage ring of 32 KB stages); stacked LSTM GEMMs hand their h rows over per tile row
//                              (EpiLstm::sync_signal / sync_wait) instead of kernel by kernel
//   gemm_tc_direct_kernel      generic functor epilogue (packed-sequence masking, unaligned outputs)
// All are launched with the programmatic-dependent-launch attribute (common.cuh): the prologue and the first weight loads
// overlap the previous kernel's tail.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <string.h>

#include <unordered_map>

#include "gemm_common.cuh"

namespace vc {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;          // bf16 elements = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;
constexpr int kBoxBytes = BM * 128;   // one 128-row x 128-byte staging box

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a broken pipeline traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  printf("vc::tc mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::tf32: fp32 operands in shared memory are read as tf32 (8 elements = 32 bytes per K step)
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- CTA pairs (thread-block clusters of 2, tcgen05 cta_group::2).  One 256x256 tile per pair: each CTA stages its own
// 128 rows of A and HALF of the W tile (128 of its 256 rows), the leader CTA issues M=256 MMAs that read both CTAs' shared
// memory and write each CTA's 128 accumulator rows into that CTA's TMEM.  A ring stage is 32 KB instead of 48 KB per 512
// MMA cycles: 5 stages fit where 3 did, which is what these GEMMs were short of (bytes in flight = bandwidth x latency).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the pair's leader CTA
// TMA load whose completion bytes are counted on the LEADER CTA's barrier (executed by both CTAs of the pair)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
// completion of the pair's MMAs -> the same barrier in both CTAs
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (8-row x 128B atoms, SBO = 1024B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);      // start address  [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell) [46,48)
  d |= (uint64_t)2 << 61;                        // layout type SWIZZLE_128B [61,64)
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=BN.
__host__ __device__ constexpr uint32_t make_idesc(int bn, int m = BM) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// kind::tf32 instruction descriptor: D=f32, A=B=tf32 (format code 2), both K-major, M=128, N=BN.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int bn, int m = BM) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// 16-byte chunk `c` (0..7) of row `r` inside a 128B-swizzled box
__device__ __forceinline__ uint32_t swz(uint32_t box, int r, int c) { return box + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void lds128(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
template <class OutT> __device__ __forceinline__ uint32_t pack16(float lo, float hi) { return pack_bf16(lo, hi); }
template <> __device__ __forceinline__ uint32_t pack16<__half>(float lo, float hi) {
  __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

enum { EPI_STORE = 0, EPI_LSTM = 1 };

struct alignas(64) TcMaps {
  CUtensorMap A[2], W[2];
  CUtensorMap Wh[2];      // CTA-pair kernels: W with a 128-row box (each CTA of a pair stages half of the 256-row tile), per z
  // EPI_STORE: io[0] = C.  EPI_LSTM: io[0] = addend (input projections), io[1] = c_prev, io[2] = c_new,
  // io[3] = h destination 0, io[4] = h destination 1.
  CUtensorMap io[5];
  CUtensorMap io1[5];     // dual-problem LSTM launch (TcArgs::dual): the io maps of problem z = 1
};
struct TcArgs {
  int M, N, K;
  int a_col0[2], a_split, a_skip;
  const float* bias[2];
  int io_col0[5][2];      // column offset of the tile origin inside each io map, per z
  int has_h1;
  // persistent EPI_LSTM kernel, encoder form: nz = 2 directions share one launch (z is the slowest tile index), and the
  // all-timestep input projections are ADDED ON THE TENSOR CORES: after the nkb operand k-blocks, four more ring passes bring
  // the [128 x 64] pieces of the addend tile (A part of the slot only) and one N = 64 MMA each multiplies them by a 64 x 64
  // identity kept in shared memory into the accumulator's column range -- no epilogue change, no extra staging boxes
  int nz, has_add;
  // tile-level hand-over between stacked LSTM GEMMs (EpiLstm::sync_*): counters per scheduled m-tile row
  unsigned int* sync_signal;
  const unsigned int* sync_wait;
  unsigned int sync_target;
  int sync_row_shift;     // gemm_tc_kernel: counter index = m0 >> sync_row_shift
  // Dual-problem form of the persistent EPI_LSTM kernel (the decoder's two stacked LSTM layers of one step in ONE launch, z =
  // layer): the layers share M and N but not K, buffers or hand-over counters.  Tiles are ordered layer 0 first, so every CTA
  // (pair) finishes its layer-0 tiles before it waits on a layer-1 tile's rows (sync_wait1: the same per-tile-row counters
  // layer 0 signals), and the ring streams across the layer boundary: the partial last tile round of layer 0 and the launch /
  // prologue / drain of a second kernel disappear into one balanced tile list.  z = 0 uses K, io[], sync_signal (and the
  // grid-level dependency wait); z = 1 uses K1, io1[], sync_wait1 / sync_signal1.
  int dual;
  int K1;
  unsigned int* sync_signal1;
  const unsigned int* sync_wait1;
  const float* a32;       // CVT kernels: the fp32 A operand [M, lda32] (converted to bf16 by the producer warps)
  int64_t lda32;
#ifdef VC_GEMM_PROBE
  long long* dbg;         // probe builds (scripts/gemm_probe.cu): clock64 stamps of CTA 0 of the persistent kernel, [tile][event]
#endif
};
#ifdef VC_GEMM_PROBE
inline long long*& probe_dbg() { static long long* p = nullptr; return p; }   // set by the probe before a launch
#define GEMM_PROBE(it, ev) do { if (g.dbg != nullptr && blockIdx.x == 0 && (it) < 16) g.dbg[(it) * 8 + (ev)] = clock64(); } while (0)
#else
#define GEMM_PROBE(it, ev) do { } while (0)
#endif

// ---------------------------------------------------------------- the kernel
// EPI_STORE: OutT = float | bf16, optional tanh.   EPI_LSTM: BN = 256 (64 hidden units per tile),
// HAS_ADD selects the encoder form (gates += input-projection tile).
template <int BN, int kStages, int kMinBlocks, int EPI, class OutT, bool TANH, bool HAS_ADD>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
gemm_tc_kernel(const __grid_constant__ TcMaps maps, const TcArgs g) {
  constexpr uint32_t kABytes = BM * BK * 2;
  constexpr uint32_t kBBytes = BN * BK * 2;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  constexpr int kAddBoxes = (EPI == EPI_LSTM && HAS_ADD) ? BN * 2 / 128 : 0;   // bf16 input-projection tile
  constexpr int kCBoxes = (EPI == EPI_LSTM) ? (BN / 4) * 4 / 128 : 0;          // fp32 cell-state tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* io_smem = smem + (size_t)kStages * kStageBytes;          // [kAddBoxes + kCBoxes] boxes
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ __align__(8) uint64_t in_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[BN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int z = blockIdx.z;
  const CUtensorMap* mapA = &maps.A[z];
  const CUtensorMap* mapW = &maps.W[z];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nkb = g.K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    mbar_init(smem_u32(&in_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(mapW) : "memory");
  }
  constexpr int kTmemCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));   // allocations are powers of two
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_slot), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  // Weights never depend on the previous kernel: the producer starts the W loads of the first ring pass before the
  // dependency wait (PDL, common.cuh); the A loads of those stages follow after it.
  const int npre = nkb < kStages ? nkb : kStages;
  if (threadIdx.x == 0) {
    for (int kb = 0; kb < npre; ++kb) {
      const uint32_t fb = smem_u32(&full_bar[kb]);
      mbar_expect_tx(fb, kStageBytes);
      tma_load_2d(smem_u32(smem + (size_t)kb * kStageBytes + kABytes), mapW, fb, kb * BK, n0);
    }
  }
  // everything above overlaps the tail of the previous kernel in the stream; with a tile-level hand-over (sync_wait) the
  // kernel does not wait for the previous grid as a whole but for the A rows of its own tile
  const bool handover = EPI == EPI_STORE && g.sync_wait != nullptr;
  if (!handover) pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      if (handover) {
        const unsigned int* flag = g.sync_wait + (m0 >> g.sync_row_shift);
        uint32_t spin = 0;
        while (ld_acquire_gpu_u32(flag) < g.sync_target) {
          if (++spin > (1u << 26)) {
            printf("vc::tc hand-over timeout (block %d,%d)\n", blockIdx.x, blockIdx.y);
            __trap();
          }
        }
        asm volatile("fence.proxy.async;" ::: "memory");     // order the acquire before the async-proxy loads
      }
      if (EPI == EPI_LSTM) {
        // epilogue inputs first: they are resident long before the accumulator is
        const uint32_t ib = smem_u32(&in_bar);
        mbar_expect_tx(ib, (uint32_t)(kAddBoxes + kCBoxes) * kBoxBytes);
        for (int i = 0; i < kAddBoxes; ++i)
          tma_load_2d(smem_u32(io_smem + (size_t)i * kBoxBytes), &maps.io[0], ib, g.io_col0[0][z] + n0 + i * 64, m0);
        for (int i = 0; i < kCBoxes; ++i)
          tma_load_2d(smem_u32(io_smem + (size_t)(kAddBoxes + i) * kBoxBytes), &maps.io[1], ib,
                      g.io_col0[1][z] + n0 / 4 + i * 32, m0);
      }
      uint32_t stage = 0, phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t fb = smem_u32(&full_bar[stage]);
        if (kb >= npre) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          mbar_expect_tx(fb, kStageBytes);
        }
        const int k = kb * BK;
        const int acol = g.a_col0[z] + k + (k >= g.a_split ? g.a_skip : 0);
        uint8_t* sa = smem + (size_t)stage * kStageBytes;
        tma_load_2d(smem_u32(sa), mapA, fb, acol, m0);
        if (kb >= npre) tma_load_2d(smem_u32(sa + kABytes), mapW, fb, k, n0);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc(BN);
      uint32_t stage = 0, phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        uint8_t* sa = smem + (size_t)stage * kStageBytes;
        const uint64_t da = make_smem_desc(smem_u32(sa));
        const uint64_t db = make_smem_desc(smem_u32(sa + kABytes));
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 16 bf16 = 32 bytes inside the swizzle atom: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty_bar[stage]));   // frees the smem slot when these MMAs retire
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(smem_u32(&tmem_full_bar));         // accumulator complete (all smem reads retired)
    }
  } else {
    // ===== epilogue: warp (warp % 4) owns TMEM lanes [32*(warp%4), +32) = output rows =====
    const int q = warp & 3;
    const int r = q * 32 + lane;                     // row inside the tile
    const int et = threadIdx.x - 64;                 // 0..127 among the epilogue threads
    // stage the bias while the main loop runs
    for (int i = et; i < BN; i += 128) {
      const int col = n0 + i;
      bias_s[i] = (g.bias[z] != nullptr && col < g.N) ? g.bias[z][col] : 0.f;
    }
    epi_bar_sync();
    if (nkb > 0) {
      mbar_wait(smem_u32(&tmem_full_bar), 0);
      tc_fence_after();
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    // After tmem_full every MMA (hence every read of the stage buffers) has retired: reuse them as staging.
    const uint32_t stg = smem_u32(smem);

    if (EPI == EPI_STORE) {
      constexpr int kColsPerBox = 128 / (int)sizeof(OutT);          // 32 (fp32) or 64 (bf16)
      constexpr int kLdPerBox = kColsPerBox / 32;
#pragma unroll 1
      for (int bx = 0; bx < BN / kColsPerBox; ++bx) {
        const uint32_t box = stg + (uint32_t)bx * kBoxBytes;
#pragma unroll
        for (int h = 0; h < kLdPerBox; ++h) {
          const int c = bx * kColsPerBox + h * 32;
          uint32_t v[32];
          if (nkb > 0) {
            tmem_ld32(taddr + (uint32_t)c, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            f[j] = __uint_as_float(v[j]) + bias_s[c + j];
            if (TANH) f[j] = tanh_<false>(f[j]);
          }
          if (sizeof(OutT) == 4) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(swz(box, r, j), __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]),
                     __float_as_uint(f[4 * j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128(swz(box, r, h * 4 + j), pack16<OutT>(f[8 * j], f[8 * j + 1]), pack16<OutT>(f[8 * j + 2], f[8 * j + 3]),
                     pack16<OutT>(f[8 * j + 4], f[8 * j + 5]), pack16<OutT>(f[8 * j + 6], f[8 * j + 7]));
          }
        }
        fence_proxy_async_smem();
        epi_bar_sync();
        if (et == 0) {
          const int col = n0 + bx * kColsPerBox;
          if (col < g.N) tma_store_2d(&maps.io[0], box, g.io_col0[0][z] + col, m0);
          tma_store_commit();
        }
      }
      if (et == 0) {
        tma_store_wait_read();
        if (g.sync_signal != nullptr) {
          // publish this tile's rows to the consumer GEMM (GemmArgs::sync_signal): stores complete, then a release increment
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          asm volatile("fence.proxy.async;" ::: "memory");
          asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(g.sync_signal + (m0 >> 7)), "r"(1u) : "memory");
        }
      }
    } else {
      // ---- fused LSTM cell: 8 hidden units (32 gate columns) per TMEM load
      constexpr int kUnits = BN / 4;                                 // 64
      const uint32_t add_s = smem_u32(io_smem);
      const uint32_t c_s = smem_u32(io_smem + (size_t)kAddBoxes * kBoxBytes);
      const uint32_t h_box = stg;                                    // 64 units x bf16 = 128B rows
      mbar_wait(smem_u32(&in_bar), 0);
#pragma unroll 1
      for (int ci = 0; ci < BN / 32; ++ci) {
        const int c = ci * 32;
        uint32_t v[32];
        if (nkb > 0) {
          tmem_ld32(taddr + (uint32_t)c, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        float gte[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) gte[j] = __uint_as_float(v[j]) + bias_s[c + j];
        if (HAS_ADD) {
          // 32 bf16 = 4 chunks of the addend box (64 columns per box)
          const uint32_t abox = add_s + (uint32_t)(c / 64) * kBoxBytes;
          const int ch0 = (c % 64) / 8;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w0, w1, w2, w3;
            lds128(swz(abox, r, ch0 + j), w0, w1, w2, w3);
            gte[8 * j + 0] += bf16_lo(w0); gte[8 * j + 1] += bf16_hi(w0);
            gte[8 * j + 2] += bf16_lo(w1); gte[8 * j + 3] += bf16_hi(w1);
            gte[8 * j + 4] += bf16_lo(w2); gte[8 * j + 5] += bf16_hi(w2);
            gte[8 * j + 6] += bf16_lo(w3); gte[8 * j + 7] += bf16_hi(w3);
          }
        }
        // previous cell state of units u0..u0+7 (fp32, 32 units per box)
        const int u0 = ci * 8;
        const uint32_t cbox = c_s + (uint32_t)(u0 / 32) * kBoxBytes;
        const int cch = (u0 % 32) / 4;
        uint32_t cw[8];
        lds128(swz(cbox, r, cch), cw[0], cw[1], cw[2], cw[3]);
        lds128(swz(cbox, r, cch + 1), cw[4], cw[5], cw[6], cw[7]);
        float hn[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float ig = sigmoid_<false>(gte[4 * u + 0]);
          const float fg = sigmoid_<false>(gte[4 * u + 1]);
          const float gg = tanh_<false>(gte[4 * u + 2]);
          const float og = sigmoid_<false>(gte[4 * u + 3]);
          const float cn = fmaf(fg, __uint_as_float(cw[u]), ig * gg);
          hn[u] = og * tanh_<false>(cn);
          cw[u] = __float_as_uint(cn);
        }
        sts128(swz(cbox, r, cch), cw[0], cw[1], cw[2], cw[3]);          // c_new in place
        sts128(swz(cbox, r, cch + 1), cw[4], cw[5], cw[6], cw[7]);
        sts128(swz(h_box, r, ci), pack_bf16(hn[0], hn[1]), pack_bf16(hn[2], hn[3]), pack_bf16(hn[4], hn[5]),
               pack_bf16(hn[6], hn[7]));
      }
      fence_proxy_async_smem();
      epi_bar_sync();
      if (et == 0) {
        const int u_tile = n0 / 4;
        tma_store_2d(&maps.io[3], h_box, g.io_col0[3][z] + u_tile, m0);
        if (g.has_h1) tma_store_2d(&maps.io[4], h_box, g.io_col0[4][z] + u_tile, m0);
        for (int i = 0; i < kCBoxes; ++i)
          tma_store_2d(&maps.io[2], c_s + (uint32_t)i * kBoxBytes, g.io_col0[2][z] + u_tile + i * 32, m0);
        tma_store_commit();
        tma_store_wait_read();
      }
      (void)kUnits;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------- persistent variant (large GEMMs)
// Same roles and epilogues, but each CTA loops over output tiles (static round-robin, n fastest so CTAs that
// run together share the A rows in L2), the smem ring keeps streaming across tile boundaries, and TWO TMEM
// accumulators (2 x 256 columns) let the epilogue of tile i overlap the tcgen05 main loop of tile i+1.
// No per-tile prologue, no wave quantisation.  BN = 256 only; EPI_LSTM without the addend (decoder form).
//   EPI_STORE: EIGHT epilogue warps = two column halves x four TMEM lane quarters (a 128x256 fp32 tile costs
//              more epilogue time than tcgen05 time with four), each half with its own pair of alternating
//              16 KB staging boxes.   smem: kStages x 48 KB ring + 4 staging boxes
//   EPI_LSTM : four epilogue warps.   smem: kStages x 48 KB ring + c tile (2 boxes, in place) + h staging box
//
// STATS (vocabulary projection, fp32 logits; video_captioning_model.py:209 log_softmax, :215 topk;
// decoder.py:269 argmax): the logits tile never goes through the staging boxes.  Per row the epilogue emits
// the maximum of every 32-column chunk and a (max, sum exp) pair per 128-column half tile, and writes a
// chunk's 32 logits to HBM ONLY IF the chunk can still belong to the row's `topk` best chunks: CTAs own
// contiguous m-major tile ranges, every epilogue thread keeps the sorted `topk` largest chunk maxima it has
// seen for its row in the current range, and a chunk is stored iff its maximum beats the topk-th of them
// (strictly: on a tie the earlier chunk wins, matching the selection's value-desc / index-asc order).  A
// chunk among the row's global top-`topk` chunks always passes (the range-local list is a subset), and
// those are the only chunks the selection kernel (decode.cuh: select_fused_kernel) reads.  Expected stores:
// ~topk*(1+ln(n/topk)) of the n chunks of a range, i.e. ~20% of the 4*V bytes per row; the plain fp32 store
// of all logits was HBM-write bound (205 MB per step at the MSVD shape, 63 us vs 38 us of tcgen05 time).
struct VocabStats {
  float* cmax;      // [M, nc] maximum logit of each 32-column chunk (-inf for chunks past N)
  float2* part;     // [M, np] (max, sum 2^((x-max)*log2e)) per 128-column half tile
  int nc, np;       // nc = 8 * tiles_n, np = 2 * tiles_n
  float* logits;    // [M, ld] fp32, sparsely written (see above)
  int64_t ld;
  int topk;         // 1..8: prune; 0: store every chunk
  int* rowthr;      // [M] or nullptr: pruning threshold shared by all CTAs (and both column halves) working on a row: the
                    // largest "topk-th best chunk maximum" any of them has seen, as an ordered-int key; a chunk whose
                    // maximum is strictly below it cannot be among the row's topk best chunks.  The caller presets it to a
                    // very negative key before every launch.
  int dbg;          // timing experiments only (VC_DEBUG_VOCAB): 1 = skip the logits stores, 2 = skip the exp sums, 4 = skip stats stores
};

constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// packed fp32 pairs (FADD2 / FFMA2 on sm_100): half the issue slots of the scalar forms
__device__ __forceinline__ uint64_t pack_f2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t add_f2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma_f2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

template <int EPI, bool CVT = false> struct PersistentCfg {
  static constexpr int kEpiWarps = (EPI == EPI_STORE) ? 8 : 4;
  static constexpr int kCvtWarps = CVT ? 16 : 0;          // fp32 -> bf16 converting A producers (CVT)
  static constexpr int kThreads = 64 + 32 * kEpiWarps + 32 * kCvtWarps;
};

// TF32: the operands in memory are fp32 (read by the tensor cores as tf32); a k-block is still 128 bytes per row,
// i.e. 32 elements, and one MMA covers 8 of them.  Used for the feature projection so that the fp32 input
// features are consumed as they are (no fp32 -> bf16 conversion pass over 1.3 MB per video).
// MC: launched as clusters of 2 CTAs (cta_group::2, see the helpers above).  The pair works on the m-tiles (2i, 2i+1) of the
// same n-tile; both producers signal the leader's full barrier, the leader's MMA completions release the ring slot and
// publish the accumulator in both CTAs, and both epilogues hand the accumulator back on the leader's barrier
// (maps.Wh[z]: W with a 128-row box).
// CVT: A is fp32 in global memory and is NOT staged by TMA: 16 converter warps load it with ld.global.nc (three k-blocks in
// flight in registers), round to bf16 and write the 128B-swizzled A tile of the ring themselves (fence.proxy.async before
// the arrive: the MMA reads shared memory through the async proxy).  The feature projection then runs at the bf16 MMA rate
// instead of the tf32 one without a separate conversion pass; a conversion THROUGH shared memory would need ~250 B/clk of
// the SM's 128 B/clk.  EPI_STORE, single CTA, no STATS.
template <int kStages, int EPI, class OutT, bool TANH, bool STATS, bool TF32 = false, bool MC = false, bool CVT = false>
__global__ void __launch_bounds__(PersistentCfg<EPI, CVT>::kThreads, 1)
gemm_tc_persistent_kernel(const __grid_constant__ TcMaps maps, const TcArgs g, const int tiles_m, const int tiles_n,
                          const VocabStats vstat) {
  static_assert(!STATS || (EPI == EPI_STORE && sizeof(OutT) == 4 && !TANH), "STATS: fp32 logits store only");
  static_assert(!TF32 || (EPI == EPI_STORE && !STATS), "TF32 operands: plain store epilogue only");
  static_assert(!CVT || (EPI == EPI_STORE && !STATS && !TF32 && !MC), "CVT: plain store epilogue, single CTA, bf16 MMA");
  constexpr int BN = 256;
  constexpr int BKE = TF32 ? 32 : BK;              // elements per k-block (128 bytes per row)
  constexpr int kEpiThreads = 32 * PersistentCfg<EPI>::kEpiWarps;
  constexpr uint32_t kABytes = BM * BK * 2;
  constexpr uint32_t kBBytes = (MC ? BN / 2 : BN) * BK * 2;  // the part of the W tile this CTA stages
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* io_smem = smem + (size_t)kStages * kStageBytes;   // STORE: 4 staging boxes; LSTM: c box0, c box1, h box
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages];
  __shared__ __align__(8) uint64_t tmem_full[2], tmem_empty[2], c_full, c_empty;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[2][BN];
  __shared__ __align__(16) float stat_bias_s[STATS ? 8 * 2 * 128 : 4];      // STATS epilogue: a private bias copy per warp and accumulator
  float* const stat_bias = stat_bias_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = g.K / BKE;
  // scheduling units: CTAs, or CTA pairs working on pairs of m-tiles (MC; tiles_m is even)
  const int crank = MC ? (int)cluster_ctarank() : 0;
  const int unit = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int units = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int kMT = MC ? 2 : 1;                  // m-tiles per scheduled tile
  const int tiles_z = (tiles_m / kMT) * tiles_n;                                     // tiles of one z (direction)
  const int num_tiles = tiles_z * ((EPI == EPI_LSTM && g.nz > 1) ? g.nz : 1);
  const bool has_add = EPI == EPI_LSTM && g.has_add != 0;
  const bool dual = EPI == EPI_LSTM && g.dual != 0;                                  // two stacked layers in one launch (TcArgs::dual)
  auto nkb_of = [&](int z) { return (dual && z != 0) ? g.K1 / BKE : nkb; };
  uint8_t* ident_s = nullptr;                                                        // identity block of the addend MMAs
  // tile schedule (m-major tile index, n fastest): round-robin, or contiguous ranges when the epilogue carries
  // per-row state from tile to tile (STATS)
  const int t_first = STATS ? (int)((int64_t)unit * num_tiles / units) : unit;
  const int t_last = STATS ? (int)((int64_t)(unit + 1) * num_tiles / units) : num_tiles;
  const int t_step = STATS ? 1 : units;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1 + PersistentCfg<EPI, CVT>::kCvtWarps);   // W producer (+ one arrival per converter warp)
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full[a]), 1);
      mbar_init(smem_u32(&tmem_empty[a]), MC ? 2 * kEpiThreads : kEpiThreads);   // MC: both CTAs' epilogues, on the leader
    }
    mbar_init(smem_u32(&c_full), 1);
    mbar_init(smem_u32(&c_empty), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.A[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(MC ? &maps.Wh[0] : &maps.W[0]) : "memory");
  }
  if (EPI == EPI_LSTM && has_add) {
    // 64 x 64 identity, K-major, 128B-swizzled like a W box (MC: this CTA's 32 rows n = 32 * crank + r of it)
    ident_s = io_smem + 3 * kBoxBytes;
    constexpr int kRows = MC ? 32 : 64;
    for (int i = threadIdx.x; i < kRows * 8; i += blockDim.x) {
      const int r = i >> 3, c = i & 7;                      // row, 16-byte chunk (8 k values)
      const int n = (MC ? crank * 32 : 0) + r;
      uint32_t w4[4] = {0u, 0u, 0u, 0u};
      if ((n >> 3) == c) w4[(n & 7) >> 1] = (n & 1) ? 0x3F800000u : 0x00003F80u;   // bf16 1.0 at k = n
      sts128(swz(smem_u32(ident_s), r, c), w4[0], w4[1], w4[2], w4[3]);
    }
    fence_proxy_async_smem();                               // generic-proxy writes -> visible to the tensor core's async proxy
  }
  if (warp == 1) {
    if (MC) tmem_alloc_2sm(smem_u32(&tmem_base_slot), 512);
    else tmem_alloc(smem_u32(&tmem_base_slot), 512);
  }
  tc_fence_before();
  if (MC) cluster_sync_all();      // the peer's barriers are initialised before anything is signalled on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  // operand tiles of a k-block -> ring slot.  MC: this CTA's half of the W tile, completion counted on the leader's barrier
  auto load_w = [&](uint8_t* sb, uint32_t fb, int k, int n0, int z) {
    if (MC) tma_load_2d_2sm(smem_u32(sb), &maps.Wh[z], fb, k, n0 + crank * (BN / 2));
    else tma_load_2d(smem_u32(sb), &maps.W[z], fb, k, n0);
  };
  auto load_a = [&](uint8_t* sa, uint32_t fb, const CUtensorMap* map, int acol, int m0) {
    if (MC) tma_load_2d_2sm(smem_u32(sa), map, fb, acol, m0);
    else tma_load_2d(smem_u32(sa), map, fb, acol, m0);
  };
  // the barrier of a slot is armed once per use, by the leader, for both CTAs' bytes
  auto arm_bytes = [&](uint32_t fb, uint32_t bytes) {
    if (CVT) mbar_expect_tx(fb, kBBytes);                  // the A half of the slot is written by the converter warps
    else if (!MC) mbar_expect_tx(fb, bytes);
    else if (crank == 0) mbar_expect_tx(fb, 2 * bytes);
  };
  auto arm = [&](uint32_t fb) { arm_bytes(fb, kStageBytes); };
  // tile index -> (z, m-tile row, n-tile)
  auto tile_z = [&](int tile) { return (EPI == EPI_LSTM && g.nz > 1) ? tile / tiles_z : 0; };
  // an epilogue thread is done with accumulator a: tell the MMA issuer (MC: the leader's barrier, from either CTA)
  auto arrive_tmem_empty = [&](uint32_t bar) {
    if (MC && crank != 0) mbar_arrive_remote(bar, 0u);
    else mbar_arrive_cta(bar);
  };
  // Weights never depend on the previous kernel: the producer starts the W loads of the first ring pass (first tile)
  // before the dependency wait (PDL, common.cuh); the A loads of those stages follow after it.
  const int npre = (t_first < t_last) ? (nkb < kStages ? nkb : kStages) : 0;
  if (threadIdx.x == 0) {
    const int n0 = (t_first % tiles_n) * BN;
    for (int kb = 0; kb < npre; ++kb) {
      const uint32_t fb = smem_u32(&full_bar[kb]);
      arm(fb);
      load_w(smem + (size_t)kb * kStageBytes + kABytes, fb, kb * BKE, n0, tile_z(t_first));
    }
  }
  // everything above overlaps the tail of the previous kernel in the stream.  A kernel that is handed its A rows tile by
  // tile (sync_wait, below) does not wait for the previous grid as a whole: everything else it reads is older.
  if (g.sync_wait == nullptr) pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int tile = t_first; tile < t_last; tile += t_step, ++it) {
        const int z = tile_z(tile), tz = tile - z * tiles_z;
        const int m0 = ((tz / tiles_n) * kMT + crank) * BM, n0 = (tz % tiles_n) * BN;
        const unsigned int* sync_wait = (dual && z != 0) ? g.sync_wait1 : g.sync_wait;
        if (sync_wait != nullptr) {
          // the producer GEMM (still running: the previous LSTM layer, or the context projection in front of the vocabulary
          // projection; dual form: layer 0's tiles of this launch) publishes its rows per m-tile row: acquire them
          const unsigned int* flag = sync_wait + tz / tiles_n;
          uint32_t spin = 0;
          while (ld_acquire_gpu_u32(flag) < g.sync_target) {
            if (++spin > (1u << 26)) {
              printf("vc::tc layer hand-over timeout (block %d tile %d)\n", blockIdx.x, tile);
              __trap();
            }
          }
          asm volatile("fence.proxy.async;" ::: "memory");   // order the acquire before the async-proxy loads
        }
        GEMM_PROBE(it, 0);
        const CUtensorMap* io = (dual && z != 0) ? maps.io1 : maps.io;
        const int nk = nkb_of(z);
        for (int kb = 0; kb < nk; ++kb) {
          const bool pre = (it == 0 && kb < npre);         // W already on its way, barrier already armed
          const uint32_t fb = smem_u32(&full_bar[stage]);
          if (!pre) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
            arm(fb);
          }
          const int k = kb * BKE;
          const int acol = g.a_col0[z] + k + (k >= g.a_split ? g.a_skip : 0);
          uint8_t* sa = smem + (size_t)stage * kStageBytes;
          if (!CVT) load_a(sa, fb, &maps.A[z], acol, m0);
          if (!pre) load_w(sa + kABytes, fb, k, n0, z);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (EPI == EPI_LSTM && has_add) {
          // the addend tile [128 rows x 256 gate columns] as four A-only ring passes (see TcArgs::has_add)
          for (int j = 0; j < BN / BK; ++j) {
            const uint32_t fb = smem_u32(&full_bar[stage]);
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
            arm_bytes(fb, kABytes);
            load_a(smem + (size_t)stage * kStageBytes, fb, &maps.io[0], g.io_col0[0][z] + n0 + j * BK, m0);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
        if (EPI == EPI_LSTM) {
          // previous cell state of this tile, needed only by epilogue(it): issued AFTER the operand loads so
          // waiting for epilogue(it-1) to release the buffer never delays main loop(it)
          mbar_wait(smem_u32(&c_empty), (uint32_t)((it & 1) ^ 1));
          const uint32_t cb = smem_u32(&c_full);
          mbar_expect_tx(cb, 2 * kBoxBytes);
          tma_load_2d(smem_u32(io_smem), &io[1], cb, g.io_col0[1][z] + n0 / 4, m0);
          tma_load_2d(smem_u32(io_smem + kBoxBytes), &io[1], cb, g.io_col0[1][z] + n0 / 4 + 32, m0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {
      // ===== MMA issuer (MC: the pair's leader, M = 256 over both CTAs) =====
      constexpr uint32_t idesc = TF32 ? make_idesc_tf32(BN, MC ? 256 : BM) : make_idesc(BN, MC ? 256 : BM);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int tile = t_first; tile < t_last; tile += t_step, ++it) {
        const int a = it & 1;
        mbar_wait(smem_u32(&tmem_empty[a]), (uint32_t)(((it >> 1) & 1) ^ 1));   // epilogue drained accumulator a
        tc_fence_after();
        GEMM_PROBE(it, 1);
        const uint32_t d = tmem_base + (uint32_t)(a * BN);
        const int nk = nkb_of(tile_z(tile));
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          if (kb == 0) GEMM_PROBE(it, 2);
          tc_fence_after();
          uint8_t* sa = smem + (size_t)stage * kStageBytes;
          const uint64_t da = make_smem_desc(smem_u32(sa));
          const uint64_t db = make_smem_desc(smem_u32(sa + kABytes));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {      // 4 MMAs of 32 bytes of K each, either operand type
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            if (MC && TF32) umma_tf32_2sm(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, acc);
            else if (MC) umma_bf16_2sm(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, acc);
            else if (TF32) umma_tf32(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, acc);
            else umma_bf16(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, acc);
          }
          if (MC) umma_commit_2sm(smem_u32(&empty_bar[stage]));     // slot free in both CTAs
          else umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (EPI == EPI_LSTM && has_add) {
          // accumulator[:, 64 j .. 64 j + 63] += addend piece j . I  (N = 64 MMAs against the identity block)
          constexpr uint32_t idesc64 = make_idesc(64, MC ? 256 : BM);
          const uint64_t di = make_smem_desc(smem_u32(ident_s));
          for (int j = 0; j < BN / BK; ++j) {
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            tc_fence_after();
            const uint64_t da = make_smem_desc(smem_u32(smem + (size_t)stage * kStageBytes));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (MC) umma_bf16_2sm(d + (uint32_t)(j * BK), da + (uint64_t)(2 * k), di + (uint64_t)(2 * k), idesc64, 1u);
              else umma_bf16(d + (uint32_t)(j * BK), da + (uint64_t)(2 * k), di + (uint64_t)(2 * k), idesc64, 1u);
            }
            if (MC) umma_commit_2sm(smem_u32(&empty_bar[stage]));
            else umma_commit(smem_u32(&empty_bar[stage]));
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
        if (MC) umma_commit_2sm(smem_u32(&tmem_full[a]));           // accumulator halves complete in both CTAs
        else umma_commit(smem_u32(&tmem_full[a]));
        GEMM_PROBE(it, 3);
      }
    }
  } else if (CVT && warp >= 2 + PersistentCfg<EPI, CVT>::kEpiWarps) {
    // ===== fp32 -> bf16 converting A producers =====
    // Lane mapping: a warp instruction reads two whole 256-byte row segments (lanes 0-15 row r, lanes 16-31 row r+1), so every
    // 32-byte sector is fetched once; a warp owns 8 rows of the tile (4 instructions), a thread 4 floats of each of 4 rows.
    const int cw = warp - (2 + PersistentCfg<EPI, CVT>::kEpiWarps);      // 0..15
    const int lrow = lane >> 4, lcol = lane & 15;                        // row parity inside an instruction, 16-byte column
    int l_tile = t_first, l_kb = 0;                        // next (tile, k-block) of the load stream
    auto issue = [&](uint4 (&b)[4]) {
      if (l_tile < t_last) {
        const int row0 = (l_tile / tiles_n) * BM + cw * 8 + lrow;
        const float* p = g.a32 + (size_t)row0 * g.lda32 + (size_t)l_kb * BK + lcol * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (row0 + 2 * i < g.M) {
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(b[i].x), "=r"(b[i].y), "=r"(b[i].z), "=r"(b[i].w) : "l"(p + (size_t)(2 * i) * g.lda32));
          } else {
            b[i] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        if (++l_kb == nkb) { l_kb = 0; l_tile += t_step; }
      }
    };
    uint32_t stage = 0, phase = 0;
    auto consume = [&](const uint4 (&b)[4]) {
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
      const uint32_t sa = smem_u32(smem + (size_t)stage * kStageBytes);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = cw * 8 + 2 * i + lrow;               // row inside the tile
        const uint32_t w0 = pack_bf16(__uint_as_float(b[i].x), __uint_as_float(b[i].y));
        const uint32_t w1 = pack_bf16(__uint_as_float(b[i].z), __uint_as_float(b[i].w));
        // 4 bf16 = 8 bytes: half of 16-byte chunk lcol/2 of the 128B-swizzled row
        const uint32_t addr = swz(sa, r, lcol >> 1) + (uint32_t)(lcol & 1) * 8u;
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(w0), "r"(w1) : "memory");
      }
      fence_proxy_async_smem();                            // generic-proxy writes -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(smem_u32(&full_bar[stage]));
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    };
    int n_mine = 0;
    for (int tile = t_first; tile < t_last; tile += t_step) ++n_mine;
    const int total = n_mine * nkb;
    uint4 b0[4], b1[4], b2[4];
    issue(b0);
    issue(b1);
    issue(b2);
    for (int j = 0; j < total; j += 3) {
      consume(b0);
      issue(b0);
      if (j + 1 < total) { consume(b1); issue(b1); }
      if (j + 2 < total) { consume(b2); issue(b2); }
    }
  } else {
    // ===== epilogue =====
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                          // row inside the tile
    int it = 0;
    if constexpr (STATS) {
      // warps 2..5 = column half 0, warps 6..9 = column half 1; thread = one row x 128 columns per tile
      const int half = (warp - 2) >> 2;
      const int eh = ((warp - 2) & 3) * 32 + lane;
      const int bar_id = 1 + half;
      constexpr int KS = 8;
      const uint32_t wslot = smem_u32(io_smem) + (uint32_t)(warp - 2) * (32u * 144u);   // this warp's 32 row slots
      const uint32_t slot = wslot + (uint32_t)lane * 144u;                             // 128 B + 16 B pad: conflict-free sts128
      float top[KS];                                      // largest chunk maxima of this row in this tile range
      float thr = -INFINITY;                              // the topk-th of them (chunks must beat it to be stored)
      int cur_mb = -1;
      for (int tile = t_first; tile < t_last; tile += t_step, ++it) {
        const int tmi = tile / tiles_n, tn = tile - tmi * tiles_n;
        const int m0 = (tmi * kMT + crank) * BM, n0 = tn * BN;
        const int a = it & 1;
        if (tmi != cur_mb) {
          cur_mb = tmi;
          thr = -INFINITY;
#pragma unroll
          for (int j = 0; j < KS; ++j) top[j] = -INFINITY;
        }
        // the 128 bias values of this half: every warp stages its own copy (4 per lane) and only syncs with itself -- staged once
        // per half behind a 128-thread barrier, the four warps of a half waited for each other at every tile (4% of the kernel's
        // stall samples, ncu source page of the r3 build)
        float* bs = stat_bias + (size_t)((warp - 2) * 2 + a) * 128;
        {
          const int col = n0 + half * 128 + lane * 4;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (g.bias[0] != nullptr) {
            if (col + 3 < g.N) b4 = __ldg(reinterpret_cast<const float4*>(g.bias[0] + col));
            else {
              if (col < g.N) b4.x = g.bias[0][col];
              if (col + 1 < g.N) b4.y = g.bias[0][col + 1];
              if (col + 2 < g.N) b4.z = g.bias[0][col + 2];
            }
          }
          *reinterpret_cast<float4*>(bs + lane * 4) = b4;
        }
        const int row = m0 + r;
        // threshold published by the other CTAs / the other column half working on this row (strictly-below test);
        // requested before the wait for the accumulator and first looked at after it, so that its latency hides under that wait
        const bool shared_thr = vstat.rowthr != nullptr && vstat.topk > 0 && row < g.M;
        int gkey = 0;
        if (shared_thr) asm volatile("ld.global.cg.b32 %0, [%1];" : "=r"(gkey) : "l"(vstat.rowthr + row));
        const float thr_in = thr;
        mbar_wait(smem_u32(&tmem_full[a]), (uint32_t)((it >> 1) & 1));
        tc_fence_after();
        if (warp == 2 && lane == 0) GEMM_PROBE(it, 4);
        __syncwarp();                                                         // this warp's bias copy is complete
        const float gthr = shared_thr ? key2f(gkey) : -INFINITY;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + half * 128);
        const bool tail_tile = n0 + BN > g.N;
        // (the epilogue paces this kernel: ~4 us per tile against 2.3 us of main loop, scripts/gemm_probe.cu vocab; of the 25 us it
        // adds to the bare GEMM, ~19 are the keep logic + kept-chunk stores, ~5-10 the exponentials, ~2 the statistics stores;
        // per-chunk (max, sum) pairs merged after the loop instead of the running pair measured the same)
        float run_m = -1e30f, run_s = 0.f;                // online (max, sum 2^((x-max) log2e)) of this half tile
        float cm[4];
        // TMEM loads are software-pipelined: the load of chunk bx+1 is in flight while chunk bx is processed
        uint32_t vbuf[2][32];
        tmem_ld32(taddr, vbuf[0]);
#pragma unroll
        for (int bx = 0; bx < 4; ++bx) {
          const int c = bx * 32;                          // column inside the half
          tmem_ld_wait();
          if (bx + 1 < 4) tmem_ld32(taddr + (uint32_t)(c + 32), vbuf[(bx + 1) & 1]);
          const uint32_t(&v)[32] = vbuf[bx & 1];
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {               // logits = accumulator + bias, two per FADD2
            const uint64_t x = add_f2(pack_f2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), pack_f2(bs[c + j], bs[c + j + 1]));
            unpack_f2(x, f[j], f[j + 1]);
          }
          const int col0 = n0 + half * 128 + c;
          if (tail_tile) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j >= g.N) f[j] = -INFINITY;
          }
          float t[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) t[j] = fmaxf(f[j], f[j + 16]);
#pragma unroll
          for (int w = 8; w > 0; w >>= 1)
#pragma unroll
            for (int j = 0; j < w; ++j) t[j] = fmaxf(t[j], t[j + w]);
          const float bm = t[0];
          cm[bx] = bm;
          if (!(vstat.dbg & 2)) {
            const float nm = fmaxf(run_m, bm);
            run_s *= ex2_approx((run_m - nm) * kLog2e);
            run_m = nm;
            const uint64_t sc2 = pack_f2(kLog2e, kLog2e), nb2 = pack_f2(-nm * kLog2e, -nm * kLog2e);
            uint64_t e2[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float y0, y1;
              unpack_f2(fma_f2(pack_f2(f[2 * j], f[2 * j + 1]), sc2, nb2), y0, y1);
              e2[j] = pack_f2(ex2_approx(y0), ex2_approx(y1));
            }
#pragma unroll
            for (int w = 8; w > 0; w >>= 1)
#pragma unroll
              for (int j = 0; j < w; ++j) e2[j] = add_f2(e2[j], e2[j + w]);
            float s0, s1;
            unpack_f2(e2[0], s0, s1);
            run_s += s0 + s1;
          }
          bool keep = false;
          if (bm > thr && !(bm < gthr) && !(vstat.dbg & 1)) {
            if (vstat.topk > 0) {
              // sorted insert (descending; an equal earlier chunk stays ahead), then refresh the threshold
#pragma unroll
              for (int j = KS - 1; j > 0; --j) top[j] = (bm > top[j - 1]) ? top[j - 1] : ((bm > top[j]) ? bm : top[j]);
              top[0] = (bm > top[0]) ? bm : top[0];
              thr = top[0];
#pragma unroll
              for (int j = 1; j < KS; ++j)
                if (j < vstat.topk) thr = top[j];
            }
            keep = row < g.M;
          }
          // Kept chunks leave through the warp's shared-memory slots so that the warp can write them as full
          // 128-byte lines (8 lanes per row, 4 rows per instruction).  Row-per-thread st.global.v4 touches 32
          // lines per instruction, and a per-lane cp.async.bulk is serialised lane by lane (uniform-register
          // operands): both cost ~36 us per step at the MSVD shape.
          const unsigned km = __ballot_sync(0xffffffffu, keep);
          if (km != 0u) {
            if (keep) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                sts128(slot + 16u * j, __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]),
                       __float_as_uint(f[4 * j + 3]));
            }
            __syncwarp();
            const int piece = lane & 7, sub = lane >> 3;
            const int gcol = col0 + piece * 4;
            float* dst0 = vstat.logits + (size_t)(m0 + q * 32) * vstat.ld + gcol;
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8) {
              const int rr = i8 * 4 + sub;
              if (((km >> rr) & 1u) && gcol < g.N) {
                uint32_t x0, x1, x2, x3;
                lds128(wslot + (uint32_t)rr * 144u + (uint32_t)piece * 16u, x0, x1, x2, x3);
                *reinterpret_cast<uint4*>(dst0 + (size_t)rr * vstat.ld) = make_uint4(x0, x1, x2, x3);
              }
            }
            __syncwarp();
          }
        }
        tc_fence_before();
        arrive_tmem_empty(smem_u32(&tmem_empty[a]));
        if (warp == 2 && lane == 0) GEMM_PROBE(it, 5);
        if (shared_thr && thr > thr_in && thr > gthr) atomicMax(vstat.rowthr + row, f2key(thr));
        if (row < g.M && !(vstat.dbg & 4)) {
          *reinterpret_cast<float4*>(vstat.cmax + (size_t)row * vstat.nc + tn * 8 + half * 4) =
              make_float4(cm[0], cm[1], cm[2], cm[3]);
          vstat.part[(size_t)row * vstat.np + tn * 2 + half] = make_float2(run_m, run_s);
        }
      }
    } else if (EPI == EPI_STORE) {
      // warps 2..5 = column half 0, warps 6..9 = column half 1 (each group covers the four lane quarters)
      const int half = (warp - 2) >> 2;
      const int eh = ((warp - 2) & 3) * 32 + lane;        // 0..127 inside the half
      const int bar_id = 1 + half;
      constexpr int kColsPerBox = 128 / (int)sizeof(OutT);
      constexpr int kLdPerBox = kColsPerBox / 32;
      constexpr int kBoxesPerHalf = (BN / 2) / kColsPerBox;   // 4 (fp32) or 2 (16-bit outputs)
      const uint32_t box_base = smem_u32(io_smem) + (uint32_t)(half * 2) * kBoxBytes;
      for (int tile = t_first; tile < t_last; tile += t_step, ++it) {
        const int m0 = ((tile / tiles_n) * kMT + crank) * BM, n0 = (tile % tiles_n) * BN;
        const int a = it & 1;
        float* bs = bias_s[a] + half * 128;
        {
          const int col = n0 + half * 128 + eh;
          bs[eh] = (g.bias[0] != nullptr && col < g.N) ? g.bias[0][col] : 0.f;
        }
        mbar_wait(smem_u32(&tmem_full[a]), (uint32_t)((it >> 1) & 1));
        tc_fence_after();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");          // bias of this half staged
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + half * 128);
#pragma unroll 1
        for (int bx = 0; bx < kBoxesPerHalf; ++bx) {
          const uint32_t box = box_base + (uint32_t)(bx & 1) * kBoxBytes;
          // the staging box written two boxes ago must have been read by its store
          if (eh == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
          for (int h = 0; h < kLdPerBox; ++h) {
            const int c = bx * kColsPerBox + h * 32;     // column inside the half
            uint32_t v[32];
            tmem_ld32(taddr + (uint32_t)c, v);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              f[j] = __uint_as_float(v[j]) + bs[c + j];
              if (TANH) f[j] = tanh_<false>(f[j]);
            }
            if (sizeof(OutT) == 4) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                sts128(swz(box, r, j), __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]),
                       __float_as_uint(f[4 * j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                sts128(swz(box, r, h * 4 + j), pack16<OutT>(f[8 * j], f[8 * j + 1]), pack16<OutT>(f[8 * j + 2], f[8 * j + 3]),
                       pack16<OutT>(f[8 * j + 4], f[8 * j + 5]), pack16<OutT>(f[8 * j + 6], f[8 * j + 7]));
            }
          }
          if (bx == kBoxesPerHalf - 1) {                  // all TMEM reads of this accumulator half are done
            tc_fence_before();
            arrive_tmem_empty(smem_u32(&tmem_empty[a]));
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (eh == 0) {
            const int col = n0 + half * 128 + bx * kColsPerBox;
            if (col < g.N) tma_store_2d(&maps.io[0], box, g.io_col0[0][0] + col, m0);
            tma_store_commit();
          }
        }
      }
      if (eh == 0) tma_store_wait_read();
    } else {
      const int et = threadIdx.x - 64;
      for (int tile = t_first; tile < t_last; tile += t_step, ++it) {
        const int z = tile_z(tile), tz = tile - z * tiles_z;
        const int m0 = ((tz / tiles_n) * kMT + crank) * BM, n0 = (tz % tiles_n) * BN;
        const int a = it & 1;
        float* bs = bias_s[a];
        for (int i = et; i < BN; i += 128) {
          const int col = n0 + i;
          bs[i] = (g.bias[z] != nullptr && col < g.N) ? g.bias[z][col] : 0.f;
        }
        mbar_wait(smem_u32(&c_full), (uint32_t)(it & 1));
        mbar_wait(smem_u32(&tmem_full[a]), (uint32_t)((it >> 1) & 1));
        tc_fence_after();
        if (et == 0) GEMM_PROBE(it, 4);
        epi_bar_sync();                                  // bias staged by all
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN);
        const uint32_t c_s = smem_u32(io_smem);
        const uint32_t h_box = smem_u32(io_smem + 2 * kBoxBytes);
#pragma unroll 1
        for (int ci = 0; ci < BN / 32; ++ci) {
          uint32_t v[32];
          tmem_ld32(taddr + (uint32_t)(ci * 32), v);
          tmem_ld_wait();
          const int u0 = ci * 8;
          const uint32_t cbox = c_s + (uint32_t)(u0 / 32) * kBoxBytes;
          const int cch = (u0 % 32) / 4;
          uint32_t cw[8];
          lds128(swz(cbox, r, cch), cw[0], cw[1], cw[2], cw[3]);
          lds128(swz(cbox, r, cch + 1), cw[4], cw[5], cw[6], cw[7]);
          float hn[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float ig = sigmoid_<false>(__uint_as_float(v[4 * u + 0]) + bs[ci * 32 + 4 * u + 0]);
            const float fg = sigmoid_<false>(__uint_as_float(v[4 * u + 1]) + bs[ci * 32 + 4 * u + 1]);
            const float gg = tanh_<false>(__uint_as_float(v[4 * u + 2]) + bs[ci * 32 + 4 * u + 2]);
            const float og = sigmoid_<false>(__uint_as_float(v[4 * u + 3]) + bs[ci * 32 + 4 * u + 3]);
            const float cn = fmaf(fg, __uint_as_float(cw[u]), ig * gg);
            hn[u] = og * tanh_<false>(cn);
            cw[u] = __float_as_uint(cn);
          }
          sts128(swz(cbox, r, cch), cw[0], cw[1], cw[2], cw[3]);
          sts128(swz(cbox, r, cch + 1), cw[4], cw[5], cw[6], cw[7]);
          sts128(swz(h_box, r, ci), pack_bf16(hn[0], hn[1]), pack_bf16(hn[2], hn[3]), pack_bf16(hn[4], hn[5]),
                 pack_bf16(hn[6], hn[7]));
        }
        tc_fence_before();
        arrive_tmem_empty(smem_u32(&tmem_empty[a]));
        fence_proxy_async_smem();
        if (et == 0) GEMM_PROBE(it, 5);
        epi_bar_sync();
        if (et == 0) {
          const int u_tile = n0 / 4;
          const CUtensorMap* io = (dual && z != 0) ? maps.io1 : maps.io;
          tma_store_2d(&io[3], h_box, g.io_col0[3][z] + u_tile, m0);
          if (g.has_h1) tma_store_2d(&io[4], h_box, g.io_col0[4][z] + u_tile, m0);
          tma_store_2d(&io[2], c_s, g.io_col0[2][z] + u_tile, m0);
          tma_store_2d(&io[2], c_s + kBoxBytes, g.io_col0[2][z] + u_tile + 32, m0);
          tma_store_commit();
          tma_store_wait_read();
          GEMM_PROBE(it, 6);
          mbar_arrive_cta(smem_u32(&c_empty));          // c / h boxes may be refilled for the next tile
          unsigned int* sync_signal = (dual && z != 0) ? g.sync_signal1 : g.sync_signal;
          if (sync_signal != nullptr) {
            // publish this tile's h rows to the next layer's GEMM: stores complete, then a release increment
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence();
            asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(sync_signal + tz / tiles_n), "r"(1u) : "memory");
          }
        }
        epi_bar_sync();                                 // nobody rewrites h_box before the stores have read it
      }
    }
  }
  tc_fence_before();
  if (MC) cluster_sync_all();      // neither CTA leaves while the peer may still signal its barriers / read its ring
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (MC) tmem_dealloc_2sm(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- legacy direct-store kernel
// Generic functor epilogue with row-per-thread global accesses: slow, kept only for the cases the staged
// epilogue does not cover (packed-sequence masking, a second fp32 output, unaligned pitches).
template <int BN, int kStages, int kMinBlocks, class Epi>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
gemm_tc_direct_kernel(const __grid_constant__ TcMaps maps, const TcArgs g, const Epi epi) {
  constexpr uint32_t kABytes = BM * BK * 2;
  constexpr uint32_t kBBytes = BN * BK * 2;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int z = blockIdx.z;
  const CUtensorMap* mapA = &maps.A[z];
  const CUtensorMap* mapW = &maps.W[z];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nkb = g.K / BK;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_slot), BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();                 // everything above overlaps the tail of the previous kernel in the stream
  pdl_launch_dependents();
  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t fb = smem_u32(&full_bar[stage]);
        mbar_expect_tx(fb, kStageBytes);
        const int k = kb * BK;
        const int acol = g.a_col0[z] + k + (k >= g.a_split ? g.a_skip : 0);
        uint8_t* sa = smem + (size_t)stage * kStageBytes;
        tma_load_2d(smem_u32(sa), mapA, fb, acol, m0);
        tma_load_2d(smem_u32(sa + kABytes), mapW, fb, k, n0);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      uint32_t stage = 0, phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        uint8_t* sa = smem + (size_t)stage * kStageBytes;
        const uint64_t da = make_smem_desc(smem_u32(sa));
        const uint64_t db = make_smem_desc(smem_u32(sa + kABytes));
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(smem_u32(&empty_bar[stage]));
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(smem_u32(&tmem_full_bar));
    }
  } else {
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    if (nkb > 0) {
      mbar_wait(smem_u32(&tmem_full_bar), 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      if (nkb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (row < g.M) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = n0 + c + j * 4;
          if (col < g.N) {
            float v[4] = {__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                          __uint_as_float(r[4 * j + 3])};
            epi(z, row, col, v);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ---------------------------------------------------------------- host side: tensor maps (cached)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct MapKey {
  const void* base; uint64_t rows, cols, ld; uint32_t box_rows, box_cols, esize;
  bool operator==(const MapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
           box_cols == o.box_cols && esize == o.esize;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.base);
    auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); };
    mix(k.rows); mix(k.cols); mix(k.ld); mix(k.box_rows); mix(k.box_cols); mix(k.esize);
    return h;
  }
};

// 2D row-major [rows, cols] (elements of esize bytes: 2 = bf16, 4 = fp32) with row pitch ld; box =
// box_rows x box_cols with box_cols*esize == 128 bytes, 128B swizzle.  Encodes are cached per thread:
// the decode loop re-uses the same few dozen maps for every step.
inline int get_map(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                   uint32_t esize) {
  static thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const uint32_t box_cols = 128 / esize;
  MapKey key{base, rows, cols, ld, box_rows, box_cols, esize};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return VC_OK;
  }
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return VC_ERR_CUDA;
  }
  VC_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * esize) % 16 == 0,
           "TMA operand must be 16B aligned with a 16B-multiple pitch (ptr=%p ld=%llu)", base, (unsigned long long)ld);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * esize};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu esize=%u", (int)r, (unsigned long long)rows,
              (unsigned long long)cols, (unsigned long long)ld, esize);
    return VC_ERR_CUDA;
  }
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return VC_OK;
}

// A buffer reference for TMA: `ptr` may point inside a row of a larger 2D buffer whose row 0 starts at
// `origin` (nullptr -> ptr is the origin) and has `cols` addressable columns from origin.
struct Ref {
  const void* ptr; const void* origin; int64_t cols; int64_t ld;
};
inline int ref_map(CUtensorMap* out, int* col0, const Ref& r, uint64_t rows, uint32_t box_rows, uint32_t esize) {
  const char* o = reinterpret_cast<const char*>(r.origin ? r.origin : r.ptr);
  *col0 = (int)((reinterpret_cast<const char*>(r.ptr) - o) / esize);
  return get_map(out, o, rows, (uint64_t)r.cols, (uint64_t)r.ld, box_rows, esize);
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
    else n = 148;
  }
  return n;
}

inline bool tma_ok(const void* p, int64_t ld, int esize) {
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * esize) % 16 == 0;
}

// CTA-pair (cta_group::2) variant of the persistent kernel: even number of m-tiles and SMs, at least one pair-tile per pair.
// VC_DISABLE_MC=1: single-CTA kernel everywhere (A/B).
inline bool use_mc(int tm, int tn) {
  const char* e = getenv("VC_DISABLE_MC");          // read per call: tests toggle it inside one process
  const bool en = !(e != nullptr && e[0] == '1');
  return en && tm % 2 == 0 && num_sms() % 2 == 0 && (int64_t)(tm / 2) * tn >= num_sms() / 2;
}
// Arrivals per scheduled m-tile row and launch of the persistent LSTM GEMM's tile-level hand-over (EpiLstm::sync_*), or 0
// when a decoder-form LSTM GEMM of this shape does not take the persistent path (then stacked layers use stream order).
inline unsigned int lstm_sync_arrivals(int M, int N) {
  const int tm = (M + BM - 1) / BM, tn = N / 256;
  if (N % 256 != 0 || tm * tn < num_sms()) return 0u;
  return use_mc(tm, tn) ? 2u * (unsigned)tn : (unsigned)tn;
}
// rows per hand-over counter of that LSTM GEMM, as a shift: 256-row pair tiles or 128-row tiles
inline int lstm_sync_row_shift(int M, int N) {
  const int tm = (M + BM - 1) / BM, tn = N / 256;
  return use_mc(tm, tn) ? 8 : 7;
}
// The consumer side is implemented in the 128x128-tile kernel: true when a store-epilogue GEMM [M, N] takes that kernel
// (see launch_gemm_tc: BN = 128 unless N >= 256 and there are at least #SMs 128x256 tiles)
inline bool ctx_handover_ok(int M, int N) {
  const int64_t tiles256 = (int64_t)((M + 127) / 128) * ((N + 255) / 256);
  return !(N >= 256 && tiles256 >= num_sms());
}
// 128 x 192 tiles for a small store GEMM (GemmArgs::force_persistent == 3): taken when they fit one CTA per SM in a single wave
// where 128 x 128 tiles do not.  The 128 x 128 kernel at two CTAs per SM is bound by its ring's round trips (K / 64 k-blocks through
// 3 slots, ~2 us each: 18.8 us for the context projection at 26% tensor-pipe activity with 160 CTAs = 0.54 waves); one CTA per SM
// affords 5 slots.
inline bool wide_tiles_ok(const GemmArgs& g) {
  const int tm = (g.M + BM - 1) / BM;
  return g.force_persistent == 3 && g.N > 256 && tm * ((g.N + 191) / 192) <= num_sms() && tm * ((g.N + 127) / 128) > num_sms();
}
// n-tiles (hand-over arrivals per 128-row tile) of the store GEMM [M, N] on the one-tile-per-CTA kernels
inline int store_tiles_n(const GemmArgs& g) { return wide_tiles_ok(g) ? (g.N + 191) / 192 : (g.N + 127) / 128; }
// maps.Wh[z] <- the W operand with a 128-row box (each CTA of a pair stages half of the 256-row tile)
inline int fill_w_half(TcMaps& mp, const GemmArgs& g, int esize) {
  for (int z = 0; z < 2; ++z)
    VC_TRY(get_map(&mp.Wh[z], g.W[z < g.nz ? z : 0], (uint64_t)g.N, (uint64_t)g.K, (uint64_t)g.ldw, 128u, esize));
  return VC_OK;
}

inline int fill_ab(TcMaps& mp, TcArgs& ta, const GemmArgs& g, int64_t a_cols, int BN, int esize = 2) {
  VC_CHECK(g.K % (128 / esize) == 0, "tensor-core GEMM needs K to be a multiple of %d (K=%d)", 128 / esize, g.K);
  VC_CHECK(g.N % 4 == 0, "bf16 tensor-core GEMM needs N %% 4 == 0 (N=%d)", g.N);
  VC_CHECK(g.a_col0 % 8 == 0 && g.a_split % BK == 0 && g.a_skip % 8 == 0, "A column offsets must be multiples of 8/64");
  memset(&ta, 0, sizeof(ta));
  ta.M = g.M; ta.N = g.N; ta.K = g.K; ta.a_split = g.a_split; ta.a_skip = g.a_skip;
  ta.sync_wait = g.sync_wait; ta.sync_target = g.sync_target; ta.sync_row_shift = g.sync_row_shift;
  ta.sync_signal = g.sync_signal;
  for (int z = 0; z < 2; ++z) {
    const int zz = z < g.nz ? z : 0;
    int c0 = 0;
    Ref ra{g.A[zz], g.a_origin, g.a_origin ? g.a_origin_cols : a_cols, g.lda};
    VC_TRY(ref_map(&mp.A[z], &c0, ra, (uint64_t)g.M, BM, esize));
    ta.a_col0[z] = c0 + g.a_col0;
    VC_TRY(get_map(&mp.W[z], g.W[zz], (uint64_t)g.N, (uint64_t)g.K, (uint64_t)g.ldw, (uint32_t)BN, esize));
  }
  return VC_OK;
}

template <class Epi>
int launch_direct(const GemmArgs& g, int64_t a_cols, const Epi& epi, cudaStream_t stream) {
  TcMaps mp;
  TcArgs ta;
  VC_TRY(fill_ab(mp, ta, g, a_cols, 128));
  constexpr int kStages = 3;
  const size_t smem = (size_t)kStages * (BM * BK * 2 + 128 * BK * 2) + 1024;
  auto kern = gemm_tc_direct_kernel<128, kStages, 2, Epi>;
  VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((g.N + 127) / 128, (g.M + BM - 1) / BM, g.nz);
  VC_CUDA(launch_pdl(kern, dim3(grid), dim3(kThreads), smem, stream, mp, ta, epi));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// ---- plain store (+bias, +tanh)
// `stats` (fp32 output only): also emit the per-row chunk maxima / log-sum-exp partials described at
// VocabStats; forces the persistent kernel.  Caller guarantees N >= 256 and the buffers' nc/np match.
template <class OutT, bool TANH>
int launch_gemm_tc(const GemmArgs& g, int64_t a_cols, const EpiStore<OutT, TANH, false>& e, cudaStream_t stream,
                   const VocabStats* stats = nullptr) {
  if (g.M == 0 || g.N == 0) return VC_OK;
  const bool direct = e.C2[0] != nullptr || g.nz != 1 || !tma_ok(e.C[0], e.ldc, sizeof(OutT));
  if (stats != nullptr) VC_CHECK(!direct && sizeof(OutT) == 4 && !TANH && g.N >= 256, "vocab statistics need a TMA-storable fp32 output");
  if (direct) return launch_direct(g, a_cols, e, stream);
  TcMaps mp;
  TcArgs ta;
  const int64_t tiles256 = (int64_t)((g.M + 127) / 128) * ((g.N + 255) / 256);
  const int BN = (stats != nullptr || (g.N >= 256 && (tiles256 >= 148 || (g.force_persistent != 0 && g.force_persistent != 3)))) ? 256 : 128;
  VC_TRY(fill_ab(mp, ta, g, a_cols, (BN == 128 && wide_tiles_ok(g)) ? 192 : BN));
  ta.bias[0] = ta.bias[1] = e.bias[0];
  VC_TRY(get_map(&mp.io[0], e.C[0], (uint64_t)g.M, (uint64_t)g.N, (uint64_t)e.ldc, BM, sizeof(OutT)));
  // consumer side of the tile-level hand-over (GemmArgs::sync_wait): the 128x128-tile kernel, and the single-CTA persistent
  // kernel with the vocabulary statistics (flags per 128-row m-tile); producer side (sync_signal): the 128x128-tile kernel
  if (BN == 256 && stats == nullptr) ta.sync_wait = nullptr;
  if (BN == 256) ta.sync_signal = nullptr;
  if (BN == 256) {
    // persistent, TMEM double-buffered: one CTA per SM loops over the tiles
    // (4 stages + staging + static smem would exceed the 227 KB limit)
    constexpr int kStages = 3;
    const size_t smem = (size_t)kStages * (BM * BK * 2 + 256 * BK * 2) + 4 * kBoxBytes + 1024;
    const int tm = (g.M + BM - 1) / BM, tn = (g.N + 255) / 256;
    // the vocabulary GEMM (stats) is paced by its epilogue, not by the operand ring: pairing only couples two epilogues
    // (measured 72 vs 68 us); VC_MC_STATS=1 pairs it anyway
    static const bool mc_stats = getenv("VC_MC_STATS") != nullptr && getenv("VC_MC_STATS")[0] == '1';
    const bool mc = (use_mc(tm, tn) || (g.force_persistent == 2 && tm % 2 == 0 && num_sms() % 2 == 0)) && (stats == nullptr || mc_stats);
    if (mc) ta.sync_wait = nullptr;          // (pair tiles span two 128-row flag rows)
    if (mc) VC_TRY(fill_w_half(mp, g, 2));
    const int ctas = mc ? num_sms() : (tm * tn < num_sms() ? tm * tn : num_sms());
    VocabStats vs;
    memset(&vs, 0, sizeof(vs));
    if constexpr (sizeof(OutT) == 4 && !TANH) {
      if (stats != nullptr) {
        vs = *stats;
        VC_CHECK(vs.nc == 8 * tn && vs.np == 2 * tn, "vocab statistics buffers do not match the tiling (nc=%d np=%d tn=%d)", vs.nc, vs.np, tn);
        vs.logits = e.C[0];
        vs.ld = e.ldc;
#ifdef VC_GEMM_PROBE
        ta.dbg = probe_dbg();
#endif
        // no staging boxes: one 144-byte slot per epilogue thread for the sparse row-chunk stores
        constexpr int kStatStages = 3;
        const size_t smem_st = (size_t)kStatStages * (BM * BK * 2 + 256 * BK * 2) + 256 * 144 + 1024;
        if (mc) {
          constexpr int kMcStages = 5;     // 32 KB per stage
          const size_t smem_mc = (size_t)kMcStages * (BM * BK * 2 + 128 * BK * 2) + 256 * 144 + 1024;
          auto kern = gemm_tc_persistent_kernel<kMcStages, EPI_STORE, OutT, TANH, true, false, true>;
          VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mc));
          VC_CUDA(launch_pdl_cluster(kern, dim3(ctas), dim3(PersistentCfg<EPI_STORE>::kThreads), smem_mc, stream, 2, mp, ta, tm, tn, vs));
        } else {
          auto kern = gemm_tc_persistent_kernel<kStatStages, EPI_STORE, OutT, TANH, true>;
          VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_st));
          VC_CUDA(launch_pdl(kern, dim3(ctas), dim3(PersistentCfg<EPI_STORE>::kThreads), smem_st, stream, mp, ta, tm, tn, vs));
        }
        VC_CUDA(cudaGetLastError());
        return VC_OK;
      }
    }
    if (mc) {
      constexpr int kMcStages = 4;       // 32 KB per stage (5 would not fit beside the 64 KB of staging boxes)
      const size_t smem_mc = (size_t)kMcStages * (BM * BK * 2 + 128 * BK * 2) + 4 * kBoxBytes + 1024;
      auto kern = gemm_tc_persistent_kernel<kMcStages, EPI_STORE, OutT, TANH, false, false, true>;
      VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mc));
      VC_CUDA(launch_pdl_cluster(kern, dim3(ctas), dim3(PersistentCfg<EPI_STORE>::kThreads), smem_mc, stream, 2, mp, ta, tm, tn, vs));
    } else {
      auto kern = gemm_tc_persistent_kernel<kStages, EPI_STORE, OutT, TANH, false>;
      VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      VC_CUDA(launch_pdl(kern, dim3(ctas), dim3(PersistentCfg<EPI_STORE>::kThreads), smem, stream, mp, ta, tm, tn, vs));
    }
  } else if (wide_tiles_ok(g)) {
    // 128 x 192 tiles, one CTA per SM, 5 stages x 40 KB (GemmArgs::force_persistent == 3: the context projection)
    constexpr int kStages = 5;
    const size_t smem = (size_t)kStages * (BM * BK * 2 + 192 * BK * 2) + 1024;
    auto kern = gemm_tc_kernel<192, kStages, 1, EPI_STORE, OutT, TANH, false>;
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((g.N + 191) / 192, (g.M + BM - 1) / BM, 1);
    VC_CUDA(launch_pdl(kern, dim3(grid), dim3(kThreads), smem, stream, mp, ta));
  } else {
    // 3 stages x 32 KB: two CTAs co-reside per SM, so one CTA's epilogue overlaps the other's main loop
    constexpr int kStages = 3;
    const size_t smem = (size_t)kStages * (BM * BK * 2 + 128 * BK * 2) + 1024;
    auto kern = gemm_tc_kernel<128, kStages, 2, EPI_STORE, OutT, TANH, false>;
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((g.N + 127) / 128, (g.M + BM - 1) / BM, 1);
    VC_CUDA(launch_pdl(kern, dim3(grid), dim3(kThreads), smem, stream, mp, ta));
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// ---- fp32 operands read as tf32 (feature projection): C[M,N] bf16 = A[M,K] fp32 . W[N,K]^T fp32 + bias
inline int launch_gemm_tc_tf32(const GemmArgs& g, int64_t a_cols, const EpiStore<bf16, false, false>& e, cudaStream_t stream) {
  if (g.M == 0 || g.N == 0) return VC_OK;
  VC_CHECK(e.C2[0] == nullptr && g.nz == 1 && tma_ok(e.C[0], e.ldc, 2) && tma_ok(g.A[0], g.lda, 4) && tma_ok(g.W[0], g.ldw, 4) &&
               g.N >= 256 && g.a_split >= g.K,
           "tf32 GEMM: unsupported operand layout");
  TcMaps mp;
  TcArgs ta;
  VC_TRY(fill_ab(mp, ta, g, a_cols, 256, 4));
  ta.bias[0] = ta.bias[1] = e.bias[0];
  VC_TRY(get_map(&mp.io[0], e.C[0], (uint64_t)g.M, (uint64_t)g.N, (uint64_t)e.ldc, BM, 2));
  constexpr int kStages = 3;
  const size_t smem = (size_t)kStages * (BM * BK * 2 + 256 * BK * 2) + 4 * kBoxBytes + 1024;
  const int tm = (g.M + BM - 1) / BM, tn = (g.N + 255) / 256;
  const bool mc = use_mc(tm, tn);
  if (mc) VC_TRY(fill_w_half(mp, g, 4));
  const int ctas = mc ? num_sms() : (tm * tn < num_sms() ? tm * tn : num_sms());
  VocabStats vs;
  memset(&vs, 0, sizeof(vs));
  if (mc) {
    constexpr int kMcStages = 4;
    const size_t smem_mc = (size_t)kMcStages * (BM * BK * 2 + 128 * BK * 2) + 4 * kBoxBytes + 1024;
    auto kern = gemm_tc_persistent_kernel<kMcStages, EPI_STORE, bf16, false, false, true, true>;
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mc));
    VC_CUDA(launch_pdl_cluster(kern, dim3(ctas), dim3(PersistentCfg<EPI_STORE>::kThreads), smem_mc, stream, 2, mp, ta, tm, tn, vs));
  } else {
    auto kern = gemm_tc_persistent_kernel<kStages, EPI_STORE, bf16, false, false, true>;
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VC_CUDA(launch_pdl(kern, dim3(ctas), dim3(PersistentCfg<EPI_STORE>::kThreads), smem, stream, mp, ta, tm, tn, vs));
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// ---- fp32 A converted to bf16 by the kernel's producer warps (feature projection): C[M,N] bf16 = bf16(A[M,K] fp32) . W[N,K]^T bf16 + bias
inline int launch_gemm_tc_cvt(const float* A32, int64_t lda32, const void* Wbf16, int64_t ldw, int M, int N, int K,
                              const EpiStore<bf16, false, false>& e, cudaStream_t stream) {
  if (M == 0 || N == 0) return VC_OK;
  VC_CHECK(e.C2[0] == nullptr && tma_ok(e.C[0], e.ldc, 2) && tma_ok(Wbf16, ldw, 2) && N >= 256 && K % BK == 0 &&
               (reinterpret_cast<uintptr_t>(A32) & 15) == 0 && lda32 % 4 == 0,
           "converting GEMM: unsupported operand layout");
  TcMaps mp;
  TcArgs ta;
  memset(&ta, 0, sizeof(ta));
  memset(&mp, 0, sizeof(mp));
  ta.M = M; ta.N = N; ta.K = K; ta.a_split = 1 << 30;
  ta.a32 = A32; ta.lda32 = lda32;
  ta.bias[0] = ta.bias[1] = e.bias[0];
  VC_TRY(get_map(&mp.W[0], Wbf16, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, 256u, 2));
  mp.W[1] = mp.W[0];
  mp.A[0] = mp.A[1] = mp.W[0];        // never loaded through (only prefetched in the prologue)
  VC_TRY(get_map(&mp.io[0], e.C[0], (uint64_t)M, (uint64_t)N, (uint64_t)e.ldc, BM, 2));
  constexpr int kStages = 3;
  const size_t smem = (size_t)kStages * (BM * BK * 2 + 256 * BK * 2) + 4 * kBoxBytes + 1024;
  const int tm = (M + BM - 1) / BM, tn = (N + 255) / 256;
  const int ctas = tm * tn < num_sms() ? tm * tn : num_sms();
  VocabStats vs;
  memset(&vs, 0, sizeof(vs));
  auto kern = gemm_tc_persistent_kernel<kStages, EPI_STORE, bf16, false, false, false, false, true>;
  VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VC_CUDA(launch_pdl(kern, dim3(ctas), dim3(PersistentCfg<EPI_STORE, true>::kThreads), smem, stream, mp, ta, tm, tn, vs));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// ---- fused LSTM cell
inline int launch_gemm_tc(const GemmArgs& g, int64_t a_cols, const EpiLstm<bf16, bf16, false>& e, cudaStream_t stream) {
  if (g.M == 0 || g.N == 0) return VC_OK;
  const int H = g.N / 4;
  const bool staged = e.lengths == nullptr && g.N % 256 == 0 && e.c_tma_cols > 0;
  if (!staged) return launch_direct(g, a_cols, e, stream);
  TcMaps mp;
  TcArgs ta;
  VC_TRY(fill_ab(mp, ta, g, a_cols, 256));
  const bool has_add = e.addend[0] != nullptr;
  for (int z = 0; z < 2; ++z) {
    const int zz = z < g.nz ? z : 0;
    ta.bias[z] = e.bias[zz];
    if (has_add) {
      Ref r{e.addend[zz], e.add_origin, e.add_origin_cols, e.add_ld};
      VC_TRY(ref_map(&mp.io[0], &ta.io_col0[0][z], r, (uint64_t)g.M, BM, 2));
    }
    Ref rc{e.c_prev[zz], e.c_origin_in, e.c_tma_cols, e.c_ld};
    VC_TRY(ref_map(&mp.io[1], &ta.io_col0[1][z], rc, (uint64_t)g.M, BM, 4));
    Ref rn{e.c_new[zz], e.c_origin_out, e.c_tma_cols, e.c_ld};
    VC_TRY(ref_map(&mp.io[2], &ta.io_col0[2][z], rn, (uint64_t)g.M, BM, 4));
    Ref rh{e.h_out0[zz], e.h0_origin, e.h0_origin ? e.h0_origin_cols : (int64_t)H, e.h0_ld};
    VC_TRY(ref_map(&mp.io[3], &ta.io_col0[3][z], rh, (uint64_t)g.M, BM, 2));
    if (e.h_out1[zz] != nullptr) {
      Ref r1{e.h_out1[zz], e.h1_origin, e.h1_origin ? e.h1_origin_cols : (int64_t)H, e.h1_ld};
      VC_TRY(ref_map(&mp.io[4], &ta.io_col0[4][z], r1, (uint64_t)g.M, BM, 2));
      ta.has_h1 = 1;
    }
  }
  dim3 grid(g.N / 256, (g.M + BM - 1) / BM, g.nz);
  // persistent kernel: decoder form (one z, no addend), or the encoder's per-timestep form -- both directions in one launch,
  // input projections added by identity MMAs (TcArgs::has_add); VC_DISABLE_PERSISTENT_ENC_STEP=1: one tile per CTA (A/B)
  const char* eso = getenv("VC_DISABLE_PERSISTENT_ENC_STEP");       // read per call: tests toggle it inside one process
  const bool enc_step_off = eso != nullptr && eso[0] == '1';
  const bool enc_form = has_add || g.nz != 1;
  if ((!enc_form || !enc_step_off) && (int)(grid.x * grid.y * grid.z) >= (enc_form ? num_sms() / 2 : num_sms())) {
    constexpr int kStages = 3;
    const size_t ident = has_add ? 64 * 128 : 0;
    const size_t smem = (size_t)kStages * (BM * BK * 2 + 256 * BK * 2) + 3 * kBoxBytes + ident + 1024;
    VocabStats vs;
    memset(&vs, 0, sizeof(vs));
    ta.sync_signal = e.sync_signal;
    ta.sync_wait = e.sync_wait;
    ta.sync_target = e.sync_target;
    ta.nz = g.nz;
    ta.has_add = has_add ? 1 : 0;
#ifdef VC_GEMM_PROBE
    ta.dbg = probe_dbg();
#endif
    if (use_mc((int)grid.y, (int)(grid.x * grid.z))) {
      VC_TRY(fill_w_half(mp, g, 2));
      constexpr int kMcStages = 5;
      const size_t smem_mc = (size_t)kMcStages * (BM * BK * 2 + 128 * BK * 2) + 3 * kBoxBytes + ident + 1024;
      auto kern = gemm_tc_persistent_kernel<kMcStages, EPI_LSTM, bf16, false, false, false, true>;
      VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mc));
      VC_CUDA(launch_pdl_cluster(kern, dim3(num_sms()), dim3(PersistentCfg<EPI_LSTM>::kThreads), smem_mc, stream, 2, mp, ta, (int)grid.y, (int)grid.x, vs));
    } else {
      auto kern = gemm_tc_persistent_kernel<kStages, EPI_LSTM, bf16, false, false>;
      VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      VC_CUDA(launch_pdl(kern, dim3(num_sms()), dim3(PersistentCfg<EPI_LSTM>::kThreads), smem, stream, mp, ta, (int)grid.y, (int)grid.x, vs));
    }
    VC_CUDA(cudaGetLastError());
    return VC_OK;
  }
  if (has_add) {
    constexpr int kStages = 2;
    const size_t smem = (size_t)kStages * (BM * BK * 2 + 256 * BK * 2) + (size_t)(4 + 2) * kBoxBytes + 1024;
    auto kern = gemm_tc_kernel<256, kStages, 1, EPI_LSTM, bf16, false, true>;
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VC_CUDA(launch_pdl(kern, dim3(grid), dim3(kThreads), smem, stream, mp, ta));
  } else {
    constexpr int kStages = 3;
    const size_t smem = (size_t)kStages * (BM * BK * 2 + 256 * BK * 2) + (size_t)2 * kBoxBytes + 1024;
    auto kern = gemm_tc_kernel<256, kStages, 1, EPI_LSTM, bf16, false, false>;
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VC_CUDA(launch_pdl(kern, dim3(grid), dim3(kThreads), smem, stream, mp, ta));
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// ---- the decoder's two stacked LSTM layers of one step in ONE launch (TcArgs::dual).  g0 / e0: layer 0, g1 / e1: layer 1
// (decoder form: one problem each, no addend, staged cell-state tiles); e0.sync_signal -> the per-tile-row counters layer 1
// waits on (e1.sync_wait), e1.sync_signal -> the next consumer's (the context projection).  Returns VC_ERR_UNSUPPORTED-like
// false through `*taken` when the shapes do not take the persistent kernel, so the caller launches the layers one by one.
inline int launch_gemm_tc_lstm_dual(const GemmArgs& g0, int64_t a_cols0, const EpiLstm<bf16, bf16, false>& e0, const GemmArgs& g1,
                                    int64_t a_cols1, const EpiLstm<bf16, bf16, false>& e1, cudaStream_t stream, bool* taken) {
  *taken = false;
  const int H = g0.N / 4;
  const bool ok = g0.M == g1.M && g0.N == g1.N && g0.M > 0 && g0.N % 256 == 0 && g0.nz == 1 && g1.nz == 1 && e0.lengths == nullptr &&
                  e1.lengths == nullptr && e0.c_tma_cols > 0 && e1.c_tma_cols > 0 && e0.addend[0] == nullptr && e1.addend[0] == nullptr &&
                  e0.sync_signal != nullptr && e1.sync_wait == e0.sync_signal && e0.sync_wait == nullptr &&
                  (e0.h_out1[0] != nullptr) == (e1.h_out1[0] != nullptr);
  const int tm = (g0.M + BM - 1) / BM, tn = g0.N / 256;
  if (!ok || tm * tn < num_sms()) return VC_OK;
  TcMaps mp, mq;
  TcArgs ta, tb;
  const GemmArgs* gs[2] = {&g0, &g1};
  const EpiLstm<bf16, bf16, false>* es[2] = {&e0, &e1};
  const int64_t acs[2] = {a_cols0, a_cols1};
  const bool mc = use_mc(tm, tn);      // (the same decision lstm_sync_arrivals() bases the counters' targets on)
  for (int l = 0; l < 2; ++l) {
    TcMaps& m_ = l == 0 ? mp : mq;
    TcArgs& t_ = l == 0 ? ta : tb;
    const GemmArgs& g = *gs[l];
    const EpiLstm<bf16, bf16, false>& e = *es[l];
    VC_CHECK(g.a_split >= g.K, "dual LSTM launch: split A operands are not supported");
    VC_TRY(fill_ab(m_, t_, g, acs[l], 256));
    if (mc) VC_TRY(fill_w_half(m_, g, 2));
    t_.bias[0] = e.bias[0];
    Ref rc{e.c_prev[0], e.c_origin_in, e.c_tma_cols, e.c_ld};
    VC_TRY(ref_map(&m_.io[1], &t_.io_col0[1][0], rc, (uint64_t)g.M, BM, 4));
    Ref rn{e.c_new[0], e.c_origin_out, e.c_tma_cols, e.c_ld};
    VC_TRY(ref_map(&m_.io[2], &t_.io_col0[2][0], rn, (uint64_t)g.M, BM, 4));
    Ref rh{e.h_out0[0], e.h0_origin, e.h0_origin ? e.h0_origin_cols : (int64_t)H, e.h0_ld};
    VC_TRY(ref_map(&m_.io[3], &t_.io_col0[3][0], rh, (uint64_t)g.M, BM, 2));
    if (e.h_out1[0] != nullptr) {
      Ref r1{e.h_out1[0], e.h1_origin, e.h1_origin ? e.h1_origin_cols : (int64_t)H, e.h1_ld};
      VC_TRY(ref_map(&m_.io[4], &t_.io_col0[4][0], r1, (uint64_t)g.M, BM, 2));
      t_.has_h1 = 1;
    }
  }
  // problem 1 -> the z = 1 slots of problem 0's launch arguments
  mp.A[1] = mq.A[0]; mp.W[1] = mq.W[0]; mp.Wh[1] = mq.Wh[0];
  for (int i = 0; i < 5; ++i) { mp.io1[i] = mq.io[i]; ta.io_col0[i][1] = tb.io_col0[i][0]; }
  mp.io1[0] = mp.io[0] = mp.io[1];     // (addend maps are never used in the decoder form; keep them valid)
  ta.a_col0[1] = tb.a_col0[0];
  ta.bias[1] = tb.bias[0];
  ta.dual = 1; ta.K1 = g1.K; ta.nz = 2; ta.has_add = 0;
  ta.sync_wait = nullptr; ta.sync_signal = e0.sync_signal;
  ta.sync_wait1 = e1.sync_wait; ta.sync_signal1 = e1.sync_signal;
  ta.sync_target = e1.sync_target;
  VocabStats vs;
  memset(&vs, 0, sizeof(vs));
#ifdef VC_GEMM_PROBE
  ta.dbg = probe_dbg();
#endif
  if (mc) {
    constexpr int kMcStages = 5;
    const size_t smem_mc = (size_t)kMcStages * (BM * BK * 2 + 128 * BK * 2) + 3 * kBoxBytes + 1024;
    auto kern = gemm_tc_persistent_kernel<kMcStages, EPI_LSTM, bf16, false, false, false, true>;
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mc));
    VC_CUDA(launch_pdl_cluster(kern, dim3(num_sms()), dim3(PersistentCfg<EPI_LSTM>::kThreads), smem_mc, stream, 2, mp, ta, tm, tn, vs));
  } else {
    constexpr int kStages = 3;
    const size_t smem = (size_t)kStages * (BM * BK * 2 + 256 * BK * 2) + 3 * kBoxBytes + 1024;
    auto kern = gemm_tc_persistent_kernel<kStages, EPI_LSTM, bf16, false, false>;
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VC_CUDA(launch_pdl(kern, dim3(num_sms()), dim3(PersistentCfg<EPI_LSTM>::kThreads), smem, stream, mp, ta, tm, tn, vs));
  }
  VC_CUDA(cudaGetLastError());
  *taken = true;
  return VC_OK;
}

}  // namespace tc
}  // namespace vc
