// bf16 tensor-core GEMM for sm_100a: TMA -> 128B-swizzled shared memory -> tcgen05.mma -> TMEM ->
// tcgen05.ld -> fused epilogue.  Hand-written PTX, no CUTLASS/cuBLAS.
//
//   C[M,N] = A[M,K] . W[N,K]^T   A, W bf16 K-major; fp32 accumulation in TMEM.
//
// Warp roles (192 threads):  warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA
// issuer (one elected lane), warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4, one output row per
// thread).  kStages-deep smem ring with full/empty mbarriers; tcgen05.commit releases smem slots and
// signals the epilogue.  One 128 x BN output tile per CTA, BK = 64 (one 128-byte swizzle atom per row).
//
// Descriptor bit layouts follow the PTX ISA tcgen05 "shared memory descriptor" / "instruction
// descriptor" tables (cross-checked against cute/arch/mma_sm100_desc.hpp in the image).
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>

#include "gemm_common.cuh"

namespace vc {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;          // bf16 elements = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a broken pipeline traps instead of hanging the GPU (gpurun strikes).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct TcArgs {
  int M, N, K;
  int a_col0, a_split, a_skip;
};

template <int BN, int kStages, int kMinBlocks, class Epi>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                    const __grid_constant__ CUtensorMap mapW0, const __grid_constant__ CUtensorMap mapW1,
                    const TcArgs g, const Epi epi) {
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
  const CUtensorMap* mapA = z ? &mapA1 : &mapA0;
  const CUtensorMap* mapW = z ? &mapW1 : &mapW0;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nkb = g.K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(mapW) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_slot), BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t stage = 0, phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t fb = smem_u32(&full_bar[stage]);
        mbar_expect_tx(fb, kStageBytes);
        const int k = kb * BK;
        const int acol = g.a_col0 + k + (k >= g.a_split ? g.a_skip : 0);
        uint8_t* sa = smem + (size_t)stage * kStageBytes;
        tma_load_2d(smem_u32(sa), mapA, fb, acol, m0);
        tma_load_2d(smem_u32(sa + kABytes), mapW, fb, k, n0);
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
      umma_commit(smem_u32(&tmem_full_bar));         // accumulator complete
    }
  } else {
    // ===== epilogue: warp (warp % 4) owns TMEM lanes [32*(warp%4), +32) = output rows =====
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

// ---------------------------------------------------------------- host side: tensor maps
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

// 2D bf16 row-major [rows, cols] with row pitch ld (elements); box = box_rows x 64, 128B swizzle.
inline int make_map_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return VC_ERR_CUDA;
  }
  VC_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0,
           "TMA operand must be 16B aligned with a 16B-multiple pitch (ptr=%p ld=%llu)", base, (unsigned long long)ld);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu", (int)r, (unsigned long long)rows,
              (unsigned long long)cols, (unsigned long long)ld);
    return VC_ERR_CUDA;
  }
  return VC_OK;
}

// A operand: `a_cols` = number of addressable columns in a row of the A buffer (>= a_col0 + K + a_skip).
template <class Epi>
int launch_gemm_tc(const GemmArgs& g, int64_t a_cols, const Epi& epi, cudaStream_t stream) {
  VC_CHECK(g.K % BK == 0, "bf16 tensor-core GEMM needs K %% 64 == 0 (K=%d)", g.K);
  VC_CHECK(g.N % 4 == 0, "bf16 tensor-core GEMM needs N %% 4 == 0 (N=%d)", g.N);
  VC_CHECK(g.a_col0 % 8 == 0 && g.a_split % BK == 0 && g.a_skip % 8 == 0, "A column offsets must be multiples of 8/64");
  if (g.M == 0 || g.N == 0) return VC_OK;
  CUtensorMap ma[2], mw[2];
  const int BN = (g.N >= 256 && ((int64_t)((g.M + 127) / 128) * ((g.N + 255) / 256) * g.nz >= 148)) ? 256 : 128;
  for (int z = 0; z < 2; ++z) {
    const int zz = z < g.nz ? z : 0;
    VC_TRY(make_map_bf16(&ma[z], g.A[zz], (uint64_t)g.M, (uint64_t)a_cols, (uint64_t)g.lda, BM));
    VC_TRY(make_map_bf16(&mw[z], g.W[zz], (uint64_t)g.N, (uint64_t)g.K, (uint64_t)g.ldw, (uint32_t)BN));
  }
  TcArgs ta{g.M, g.N, g.K, g.a_col0, g.a_split, g.a_skip};
  if (BN == 256) {
    constexpr int kStages = 4;
    const size_t smem = (size_t)kStages * (BM * BK * 2 + 256 * BK * 2) + 1024;
    auto kern = gemm_bf16_tc_kernel<256, kStages, 1, Epi>;
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((g.N + 255) / 256, (g.M + BM - 1) / BM, g.nz);
    kern<<<grid, kThreads, smem, stream>>>(ma[0], ma[1], mw[0], mw[1], ta, epi);
  } else {
    // 3 stages x 32 KB: two CTAs co-reside per SM, so one CTA's epilogue overlaps the other's main loop
    constexpr int kStages = 3;
    const size_t smem = (size_t)kStages * (BM * BK * 2 + 128 * BK * 2) + 1024;
    auto kern = gemm_bf16_tc_kernel<128, kStages, 2, Epi>;
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((g.N + 127) / 128, (g.M + BM - 1) / BM, g.nz);
    kern<<<grid, kThreads, smem, stream>>>(ma[0], ma[1], mw[0], mw[1], ta, epi);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

}  // namespace tc
}  // namespace vc
