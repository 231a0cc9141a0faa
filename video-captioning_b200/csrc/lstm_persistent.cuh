// Persistent, weights-stationary bidirectional LSTM layer for sm_100a (bf16 mode).
//
// Replaces the 2*T dependent recurrent launches of one encoder layer (nn.LSTM inside encoder.py:84) with
// ONE cooperative launch that runs all T timesteps of both directions:
//
//   grid  = (4H/128 gate-column tiles) x (ceil(B/256) row tiles) x (2 directions)   <= #SMs, 1 CTA / SM
//   smem  = W_hh tile [128 x H] bf16, RESIDENT for the whole layer (128 KB at H=512)
//           + 3-stage ring of h_{t-1} k-blocks (TMA) + h staging boxes for the TMA stores.
//           The input-projection tile does NOT go through shared memory: each epilogue thread loads its own 128
//           contiguous bytes (one row x 16 hidden units x 4 gates) into registers before it waits for the step's
//           accumulator, i.e. under the step's own load + MMA time.
//   TMEM  = two 128x128 fp32 accumulators (rows 0-127 / 128-255 of the CTA's row tile): the LSTM epilogue
//           of one half overlaps the tcgen05 main loop of the other
//   regs  = the cell state c (one row x 32 hidden units per epilogue thread) never leaves registers
//
// Per step each CTA computes gates[256 rows, 128 cols] = h_{t-1}[256, H] . W_tile^T + xproj_t + bias, applies
// the fused cell (gate-interleaved columns: 32 hidden units per tile) and TMA-stores its h_t slice into the
// layer output [B, T, 2H].  The 4H/128 CTAs that share a (direction, row-half) exchange h_t through L2:
// a global arrival counter per group, released after the bulk store has completed and acquired by the TMA
// producers before they load h_t for step t+1.  The spin is safe because the launch is cooperative
// (all CTAs co-resident) and bounded (trap instead of hang).
//
// Where a step's 7.2 us go (clock64 stamps of one CTA, scripts/plstm_probe.cu, VC_PLSTM_PROBE): flag seen -> first k-block
// in shared memory 0.9 us, -> all 8 k-blocks through the 3-slot ring and their MMAs committed 3.3 us, cell epilogue 1.0 us,
// bulk store of the h slice 0.4 us, release + flag visible to the 16 consumers 1.5 us.  The two row-halves of a CTA are two
// such chains half a period apart.  A 16-CTA cluster variant (remote mbarrier arrivals instead of the L2 flag, every k-block
// multicast to the cluster: 2 MB instead of 32 MB of L2 reads per step) measured the same chain length -- a slot can only be
// refilled when all 16 CTAs have consumed it -- and only 7 clusters of 16 fit a B200, so it was not kept.
//
// Warp roles (608 threads): w0 = h_{t-1} TMA producer, w1 = TMEM alloc + MMA issuer, w2-9 = epilogue of
// row-half 0, w10-17 = epilogue of row-half 1 (two warps per TMEM lane quarter, each taking 16 of the tile's 32 hidden
// units: the epilogue is on the step's critical path, so its latency matters more than its issue slots).
#pragma once
#include <cooperative_groups.h>

#include "gemm_tc.cuh"

namespace vc {
namespace tc {

constexpr int kPlThreads = 576;
constexpr int kPlEpiThreads = 256;  // epilogue threads per row-half
// Ring depth.  Measured with scripts/plstm_probe.cu (B = 1024, T = 80, H = 512, us per step): 3 stages 7.58, 4 stages 8.71,
// 5 stages 8.79.  All 16 CTAs of a group pull the same 16 KB boxes of h_{t-1} from L2 at the same moment; more boxes in flight
// only lengthen the latency of the first one (0.9 -> 1.5 us) and of everything else that goes through L2.
constexpr int kPlStages = 3;
constexpr int kPlBN = 128;          // gate columns per CTA = 32 hidden units
#ifndef PL_PAIR_STAGES
#define PL_PAIR_STAGES 3
#endif
constexpr int kPlPairStages = PL_PAIR_STAGES;

struct alignas(64) PLstmMaps {
  CUtensorMap out_ld;   // layer output [B, T*2H] bf16, box 64 x 128, 128B swizzle (h_{t-1} loads)
  CUtensorMap out_st;   // same buffer, box 32 x 128, 64B swizzle (h_t stores)
  CUtensorMap W[2];     // W_hh [4H, H] per direction, box 64 x 128
};
struct PLstmArgs {
  int B, T, H;
  const float* bias[2];      // nullable (the encoder's biases are folded into xp)
  const bf16* xp;            // input projections [B, T*8H] bf16 (both directions, gate-interleaved)
  unsigned int* flags;       // [2 dirs][MT][2 halves] arrival counters, zeroed before launch
  long long* dbg;            // VC_PLSTM_PROBE builds (scripts/plstm_probe.cu): clock64 stamps of CTA (0,0,0), steps 40..47
};

#ifdef VC_PLSTM_PROBE
#define PL_PROBE(ev, t) do { if (g.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (t) >= 40 && (t) < 48) g.dbg[((t) - 40) * 32 + (ev)] = clock64(); } while (0)
#else
#define PL_PROBE(ev, t) do { } while (0)
#endif
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// 16-byte chunk c (0..3) of row r in a 64-byte-row box with TMA SWIZZLE_64B (Swizzle<2,4,3>)
__device__ __forceinline__ uint32_t swz64(uint32_t box, int r, int c) {
  return box + (uint32_t)r * 64u + (uint32_t)((c ^ ((r >> 1) & 3)) << 4);
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// TMA load delivered to the same shared-memory offset of every CTA in `mask` of the cluster; each destination's barrier (same
// offset) receives the bytes
__device__ __forceinline__ void tma_load_2d_mcast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// completion of this CTA's MMAs -> the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

// CL = 1: one CTA per (gate-column tile, row tile, direction) as described above.  CL = 2 (clusters of two neighbouring
// gate-column tiles, which need the same h_{t-1} rows): the two producers issue alternate k-blocks and TMA-multicast each box
// into both CTAs' rings, so a box leaves L2 once per cluster instead of twice; a ring slot is refilled when BOTH CTAs' MMAs
// have consumed it (tcgen05.commit multicast onto both empty barriers).  MMAs, epilogue and flags are unchanged.
template <int CL>
__global__ void __launch_bounds__(kPlThreads, 1) lstm_layer_persistent_kernel(const __grid_constant__ PLstmMaps maps,
                                                                             const PLstmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int H = g.H, T = g.T;
  const int nkb = H / BK;
  uint8_t* w_s = smem;                                    // nkb boxes of [128 n x 64 k] bf16
  uint8_t* a_s = w_s + (size_t)nkb * kBoxBytes;           // kPlStages boxes of [128 rows x 64 k]
  uint8_t* h_s = a_s + (size_t)kPlStages * kBoxBytes;     // 2 x [128 rows x 32 units] bf16 (64B rows)
  __shared__ __align__(8) uint64_t w_bar, full_bar[kPlStages], empty_bar[kPlStages];
  __shared__ __align__(8) uint64_t tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[kPlBN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int z = blockIdx.z;                               // direction
  const int n0 = blockIdx.x * kPlBN;                      // gate column tile
  const int m0 = blockIdx.y * 256;                        // row tile
  const int NT = gridDim.x;
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  unsigned int* flag0 = g.flags + ((size_t)(z * gridDim.y + blockIdx.y) * 2);

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&w_bar), 1);
    for (int s = 0; s < kPlStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), CL);               // one commit per CTA of the cluster
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(smem_u32(&tmem_full[h]), 1);
      mbar_init(smem_u32(&tmem_empty[h]), kPlEpiThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < kPlBN) {
    const float* bz = g.bias[z];
    bias_s[threadIdx.x] = bz ? bz[n0 + threadIdx.x] : 0.f;
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_slot), 256);
  tc_fence_before();
  if (CL > 1) cluster_sync_all();                           // the peer's barriers are initialised before anything lands on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== W_hh tile (once) + h_{t-1} k-blocks (every step) =====
      mbar_expect_tx(smem_u32(&w_bar), (uint32_t)nkb * kBoxBytes);
      for (int kb = 0; kb < nkb; ++kb)
        tma_load_2d(smem_u32(w_s + (size_t)kb * kBoxBytes), &maps.W[z], smem_u32(&w_bar), kb * BK, n0);
      uint32_t stage = 0, phase = 0;
      for (int t = 1; t < T; ++t) {
        const int tprev = (z == 0) ? (t - 1) : (T - t);   // time index holding h_{t-1} of this direction
        const int col0 = tprev * 2 * H + z * H;
        for (int half = 0; half < 2; ++half) {
          // all NT column tiles of this (direction, row-half) must have published h_{t-1}
          const unsigned int need = (unsigned int)NT * (unsigned int)t;
          uint32_t spin = 0;
          while (ld_acquire_gpu(flag0 + half) < need) {
            if (++spin > (1u << 24)) {
              printf("vc::lstm_persistent flag timeout (block %d,%d,%d step %d half %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, t, half);
              __trap();
            }
          }
          PL_PROBE(0 + half, t);
          asm volatile("fence.proxy.async;" ::: "memory");   // order the acquire before the async-proxy loads
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            mbar_expect_tx(fb, kBoxBytes);
            if (CL == 1) tma_load_2d(smem_u32(a_s + (size_t)stage * kBoxBytes), &maps.out_ld, fb, col0 + kb * BK, m0 + half * 128);
            else if ((kb & 1) == crank)
              tma_load_2d_mcast(smem_u32(a_s + (size_t)stage * kBoxBytes), &maps.out_ld, fb, col0 + kb * BK, m0 + half * 128, (uint16_t)3);
            if (++stage == kPlStages) { stage = 0; phase ^= 1; }
          }
          PL_PROBE(2 + half, t);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc(kPlBN);
      mbar_wait(smem_u32(&w_bar), 0);
      uint32_t stage = 0, phase = 0;
      for (int t = 1; t < T; ++t) {
        for (int half = 0; half < 2; ++half) {
          mbar_wait(smem_u32(&tmem_empty[half]), (uint32_t)(((t - 1) & 1) ^ 1));   // epilogue drained this accumulator
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)(half * 128);
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            if (kb == 0) PL_PROBE(4 + half, t);
            tc_fence_after();
            const uint64_t da = make_smem_desc(smem_u32(a_s + (size_t)stage * kBoxBytes));
            const uint64_t db = make_smem_desc(smem_u32(w_s + (size_t)kb * kBoxBytes));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            if (CL == 1) umma_commit(smem_u32(&empty_bar[stage]));
            else umma_commit_mcast(smem_u32(&empty_bar[stage]), (uint16_t)3);     // slot consumed by this CTA: tell both producers
            if (++stage == kPlStages) { stage = 0; phase ^= 1; }
          }
          umma_commit(smem_u32(&tmem_full[half]));
          PL_PROBE(6 + half, t);
        }
      }
    }
  } else if (warp >= 2 && warp < 18) {
    // ===== epilogue: fused LSTM cell, one row x 16 hidden units per thread =====
    const int half = (warp - 2) >> 3;
    const int sub = ((warp - 2) >> 2) & 1;                // which 2 of the 4 32-column chunks (16 of the 32 hidden units)
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                          // row inside the half
    const int et = (warp - 2 - half * 8) * 32 + lane;     // 0..255 inside the half
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 128);
    const uint32_t hbox = smem_u32(h_s + (size_t)half * (128 * 64));
    // this thread's input-projection row: 128 contiguous bytes per step (16 hidden units x 4 gates, gate-interleaved)
    const int grow = m0 + half * 128 + r;
    const bool row_ok = grow < g.B;
    const bf16* xrow = g.xp + (size_t)(row_ok ? grow : 0) * ((size_t)T * 8 * H) + (size_t)z * 4 * H + n0 + sub * 64;
    float c[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) c[u] = 0.f;
    for (int t = 0; t < T; ++t) {
      const int tt = (z == 0) ? t : (T - 1 - t);
      // issued before the wait for the accumulator: the loads complete under the step's MMA time
      uint32_t xw[32];
      {
        const bf16* xp_t = xrow + (size_t)tt * 8 * H;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (row_ok) {
            asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(xw[8 * i + 0]), "=r"(xw[8 * i + 1]), "=r"(xw[8 * i + 2]), "=r"(xw[8 * i + 3]), "=r"(xw[8 * i + 4]),
                           "=r"(xw[8 * i + 5]), "=r"(xw[8 * i + 6]), "=r"(xw[8 * i + 7])
                         : "l"(xp_t + 16 * i));
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) xw[8 * i + j] = 0u;
          }
        }
      }
      if (t > 0) {
        mbar_wait(smem_u32(&tmem_full[half]), (uint32_t)((t - 1) & 1));
        tc_fence_after();
        if (et == 0) PL_PROBE(8 + half, t);
      }
#pragma unroll
      for (int cj = 0; cj < 2; ++cj) {
        const int ci = 2 * sub + cj;
        uint32_t v[32];
        if (t > 0) {
          tmem_ld32(taddr + (uint32_t)(ci * 32), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;        // h_{-1} = 0: gates = xproj + bias
        }
        float gte[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t w = xw[cj * 16 + i];
          gte[2 * i] = __uint_as_float(v[2 * i]) + bias_s[ci * 32 + 2 * i] + bf16_lo(w);
          gte[2 * i + 1] = __uint_as_float(v[2 * i + 1]) + bias_s[ci * 32 + 2 * i + 1] + bf16_hi(w);
        }
        float hn[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float ig = sigmoid_<false>(gte[4 * u + 0]);
          const float fg = sigmoid_<false>(gte[4 * u + 1]);
          const float gg = tanh_<false>(gte[4 * u + 2]);
          const float og = sigmoid_<false>(gte[4 * u + 3]);
          const float cn = fmaf(fg, c[cj * 8 + u], ig * gg);
          c[cj * 8 + u] = cn;
          hn[u] = og * tanh_<false>(cn);
        }
        sts128(swz64(hbox, r, ci), pack_bf16(hn[0], hn[1]), pack_bf16(hn[2], hn[3]), pack_bf16(hn[4], hn[5]),
               pack_bf16(hn[6], hn[7]));
      }
      if (t > 0) {
        tc_fence_before();
        mbar_arrive(smem_u32(&tmem_empty[half]));         // accumulator may be overwritten by step t+1
      }
      fence_proxy_async_smem();
      if (et == 0) PL_PROBE(10 + half, t);
      asm volatile("bar.sync %0, 256;" ::"r"(2 + half) : "memory");
      if (et == 0) {
        tma_store_2d(&maps.out_st, hbox, tt * 2 * H + z * H + n0 / 4, m0 + half * 128);
        tma_store_commit();
        tma_store_wait_read();                            // staging box may be rewritten
        PL_PROBE(12 + half, t);
      }
      asm volatile("bar.sync %0, 256;" ::"r"(2 + half) : "memory");
      if (et == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // h_t slice is in global memory
        PL_PROBE(14 + half, t);
        // The bulk store has completed (its writes are visible to this thread); the gpu-scope release orders them before the
        // increment.  (A fence.proxy.async + __threadfence() in front of it cost 0.35 us per step and half: 7.53 -> 7.19 us.)
        red_release_gpu_add(flag0 + half, 1u);
        PL_PROBE(16 + half, t);
      }
    }
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all();                           // neither CTA leaves while the peer may still write its ring / barriers
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---------------------------------------------------------------- CTA-pair form (clusters of 2, tcgen05 cta_group::2)
// The kernel above is bound by what an SM can take in per step: each CTA needs h_{t-1} of its whole 256-row tile, 256 KB per
// step through one SM's L2 port (3.3 us of the 7.2 us step; multicasting the boxes inside a cluster does not change what each
// SM has to receive).  Here two CTAs with neighbouring gate-column tiles form a pair and split the ROWS instead: CTA r of the
// pair stages only rows [128 r, 128 r + 128) of h_{t-1} (128 KB per step) next to its resident W_hh tile, the leader issues
// M = 256, N = 256 MMAs (A rows from both CTAs, W columns from both CTAs), and each CTA's TMEM receives its own 128 rows x
// all 256 gate columns of the pair -- so CTA r runs the cell epilogue of row-half r for 64 hidden units and publishes them.
// Same grid, same flags (one counter per (direction, row tile, row-half); the NT/2 CTAs that own that half arrive), same
// arithmetic per element (bit-identical layer outputs, scripts/plstm_probe.cu).  The two row-halves advance in lockstep (one MMA
// chain per step), which gives up the overlap of one half's epilogue with the other half's main loop that the kernel above has.
// MEASURED, NOT ADOPTED (opt-in, VC_PLSTM_PAIR=1): 8.7 us per step against 7.2.  The load phase did not shrink with the bytes
// (128 KB still take 2.2 us after the first box, with 3, 4 or 5 ring stages alike: the 16 MB per step that the 128 CTAs pull
// out of the same few L2 lines are bound on the L2 side, not at the SM port), and the cell epilogue of all 512 threads at once
// is MUFU-bound (5 transcendentals x 8192 cell updates per CTA = 1.3 us) on the critical path: flag -> first box 0.9, loads +
// MMAs 2.2, cell 1.7, stores 0.9, release -> flag 1.5 us.
// No tmem_empty barrier: the MMAs of step t+1 need both producers' boxes, each producer has acquired its half's flag, and a
// half's flag includes the CTA's own arrival, which follows its epilogue's TMEM reads.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPlThreads, 1)
    lstm_layer_pair_kernel(const __grid_constant__ PLstmMaps maps, const PLstmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int H = g.H, T = g.T;
  const int nkb = H / BK;
  uint8_t* w_s = smem;                                    // nkb boxes of [128 n x 64 k] bf16: this CTA's half of the pair's W tile
  uint8_t* a_s = w_s + (size_t)nkb * kBoxBytes;           // kPlPairStages boxes of [128 rows x 64 k]: this CTA's rows of h_{t-1}
  uint8_t* h_s = a_s + (size_t)kPlPairStages * kBoxBytes;     // 2 x [128 rows x 32 units] bf16 (64B rows): this CTA's rows of h_t
  __shared__ __align__(8) uint64_t w_bar, full_bar[kPlPairStages], empty_bar[kPlPairStages], tmem_full;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[2 * kPlBN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();               // = blockIdx.x & 1: row-half of this CTA
  const int z = blockIdx.z;                               // direction
  const int n0 = blockIdx.x * kPlBN;                      // this CTA's W tile (gate columns)
  const int np0 = (blockIdx.x & ~1) * kPlBN;              // the pair's 256 gate columns = 64 hidden units
  const int m0 = blockIdx.y * 256 + crank * 128;          // this CTA's rows
  const int NT = gridDim.x;
  unsigned int* flag = g.flags + ((size_t)(z * gridDim.y + blockIdx.y) * 2) + crank;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&w_bar), 1);
    for (int s = 0; s < kPlPairStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&tmem_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 2 * kPlBN) {
    const float* bz = g.bias[z];
    bias_s[threadIdx.x] = bz ? bz[np0 + threadIdx.x] : 0.f;
  }
  if (warp == 1) tmem_alloc_2sm(smem_u32(&tmem_base_slot), 256);
  tc_fence_before();
  cluster_sync_all();                                     // the peer's barriers are initialised before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== W_hh tile (once) + this CTA's rows of h_{t-1} (every step); completion bytes are counted on the leader's barriers =====
      if (crank == 0) mbar_expect_tx(smem_u32(&w_bar), 2u * (uint32_t)nkb * kBoxBytes);
      for (int kb = 0; kb < nkb; ++kb)
        tma_load_2d_2sm(smem_u32(w_s + (size_t)kb * kBoxBytes), &maps.W[z], smem_u32(&w_bar), kb * BK, n0);
      uint32_t stage = 0, phase = 0;
      for (int t = 1; t < T; ++t) {
        const int tprev = (z == 0) ? (t - 1) : (T - t);   // time index holding h_{t-1} of this direction
        const int col0 = tprev * 2 * H + z * H;
        // the NT/2 CTAs that own this row-half (one per pair) must have published h_{t-1}
        const unsigned int need = (unsigned int)(NT / 2) * (unsigned int)t;
        uint32_t spin = 0;
        while (ld_acquire_gpu(flag) < need) {
          if (++spin > (1u << 24)) {
            printf("vc::lstm_pair flag timeout (block %d,%d,%d step %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, t);
            __trap();
          }
        }
        PL_PROBE(0 + crank, t);
        asm volatile("fence.proxy.async;" ::: "memory");   // order the acquire before the async-proxy loads
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          if (crank == 0) mbar_expect_tx(fb, 2u * kBoxBytes);
          tma_load_2d_2sm(smem_u32(a_s + (size_t)stage * kBoxBytes), &maps.out_ld, fb, col0 + kb * BK, m0);
          if (++stage == kPlPairStages) { stage = 0; phase ^= 1; }
        }
        PL_PROBE(2 + crank, t);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {
      // ===== MMA issuer: the pair's leader, M = 256 (both CTAs' rows) x N = 256 (both CTAs' W tiles) =====
      constexpr uint32_t idesc = make_idesc(2 * kPlBN, 256);
      mbar_wait(smem_u32(&w_bar), 0);
      uint32_t stage = 0, phase = 0;
      for (int t = 1; t < T; ++t) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          if (kb == 0) PL_PROBE(4, t);
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(a_s + (size_t)stage * kBoxBytes));
          const uint64_t db = make_smem_desc(smem_u32(w_s + (size_t)kb * kBoxBytes));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16_2sm(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(smem_u32(&empty_bar[stage]));   // slot free in both CTAs
          if (++stage == kPlPairStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(smem_u32(&tmem_full));            // both CTAs' accumulator rows are complete
        PL_PROBE(6, t);
      }
    }
  } else if (warp >= 2 && warp < 18) {
    // ===== epilogue: fused LSTM cell for this CTA's 128 rows x 64 hidden units, one row x 16 hidden units per thread =====
    const int sub = (warp - 2) >> 2;                      // which 64 of the pair's 256 gate columns
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                          // row inside the half
    const int et = (warp - 2) * 32 + lane;                // 0..511
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t hbox = smem_u32(h_s);
    const int grow = m0 + r;
    const bool row_ok = grow < g.B;
    const bf16* xrow = g.xp + (size_t)(row_ok ? grow : 0) * ((size_t)T * 8 * H) + (size_t)z * 4 * H + np0 + sub * 64;
    float c[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) c[u] = 0.f;
    for (int t = 0; t < T; ++t) {
      const int tt = (z == 0) ? t : (T - 1 - t);
      uint32_t xw[32];
      {
        const bf16* xp_t = xrow + (size_t)tt * 8 * H;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (row_ok) {
            asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(xw[8 * i + 0]), "=r"(xw[8 * i + 1]), "=r"(xw[8 * i + 2]), "=r"(xw[8 * i + 3]), "=r"(xw[8 * i + 4]),
                           "=r"(xw[8 * i + 5]), "=r"(xw[8 * i + 6]), "=r"(xw[8 * i + 7])
                         : "l"(xp_t + 16 * i));
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) xw[8 * i + j] = 0u;
          }
        }
      }
      if (t > 0) {
        mbar_wait(smem_u32(&tmem_full), (uint32_t)((t - 1) & 1));
        tc_fence_after();
        if (et == 0) PL_PROBE(8 + crank, t);
      }
#pragma unroll
      for (int cj = 0; cj < 2; ++cj) {
        const int ci = 2 * sub + cj;                      // 32-column chunk (8 hidden units) of the pair's 256 columns
        uint32_t v[32];
        if (t > 0) {
          tmem_ld32(taddr + (uint32_t)(ci * 32), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;        // h_{-1} = 0: gates = xproj + bias
        }
        float gte[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t w = xw[cj * 16 + i];
          gte[2 * i] = __uint_as_float(v[2 * i]) + bias_s[ci * 32 + 2 * i] + bf16_lo(w);
          gte[2 * i + 1] = __uint_as_float(v[2 * i + 1]) + bias_s[ci * 32 + 2 * i + 1] + bf16_hi(w);
        }
        float hn[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float ig = sigmoid_<false>(gte[4 * u + 0]);
          const float fg = sigmoid_<false>(gte[4 * u + 1]);
          const float gg = tanh_<false>(gte[4 * u + 2]);
          const float og = sigmoid_<false>(gte[4 * u + 3]);
          const float cn = fmaf(fg, c[cj * 8 + u], ig * gg);
          c[cj * 8 + u] = cn;
          hn[u] = og * tanh_<false>(cn);
        }
        sts128(swz64(hbox + (uint32_t)(ci >> 2) * (128u * 64u), r, ci & 3), pack_bf16(hn[0], hn[1]), pack_bf16(hn[2], hn[3]),
               pack_bf16(hn[4], hn[5]), pack_bf16(hn[6], hn[7]));
      }
      tc_fence_before();
      fence_proxy_async_smem();
      if (et == 0) PL_PROBE(10 + crank, t);
      asm volatile("bar.sync 2, 512;" ::: "memory");
      if (et == 0) {
        const int ucol = tt * 2 * H + z * H + np0 / 4;
        tma_store_2d(&maps.out_st, hbox, ucol, m0);
        tma_store_2d(&maps.out_st, hbox + 128u * 64u, ucol + 32, m0);
        tma_store_commit();
        tma_store_wait_read();                            // staging boxes may be rewritten
        PL_PROBE(12 + crank, t);
      }
      asm volatile("bar.sync 2, 512;" ::: "memory");
      if (et == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this CTA's h_t slice is in global memory
        PL_PROBE(14 + crank, t);
        red_release_gpu_add(flag, 1u);
        PL_PROBE(16 + crank, t);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();      // neither CTA leaves while the peer may still signal its barriers / read its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 256);
  }
}

// box_cols x box_rows map with an explicit swizzle (the h store uses 32-column = 64-byte rows)
inline int get_map_sw(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                      uint32_t box_cols, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return VC_ERR_CUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (persistent LSTM) failed (%d)", (int)r);
    return VC_ERR_CUDA;
  }
  return VC_OK;
}

inline size_t plstm_smem_bytes(int H, int stages = kPlStages) {
  return (size_t)(H / BK) * kBoxBytes + (size_t)stages * kBoxBytes + 2 * (128 * 64) + 1024;
}

// Largest batch one cooperative launch can take (0 = shape not supported by the persistent kernel).
inline int plstm_max_batch(int H, int num_sms) {
  if (H % BK != 0 || (4 * H) % kPlBN != 0 || plstm_smem_bytes(H) > 227 * 1024) return 0;
  const int NT = 4 * H / kPlBN;
  const int MT = num_sms / (2 * NT);
  return MT * 256;
}

// One bidirectional layer, all T steps.  out: [B, T, 2H] bf16 (written), xp: [B, T, 8H] bf16 (both directions'
// input projections incl. biases, gate-interleaved), W[dir]: [4H, H] bf16 gate-interleaved, flags: >= 4*MT uints.
inline int launch_lstm_layer_persistent(bf16* out, const bf16* xp, const void* W0, const void* W1, int B, int T, int H,
                                        unsigned int* flags, cudaStream_t stream, long long* dbg = nullptr, int pair_mode = -1, int cluster_mode = -1) {
  PLstmMaps mp;
  VC_TRY(get_map(&mp.out_ld, out, (uint64_t)B, (uint64_t)T * 2 * H, (uint64_t)T * 2 * H, BM, 2));
  VC_TRY(get_map_sw(&mp.out_st, out, (uint64_t)B, (uint64_t)T * 2 * H, (uint64_t)T * 2 * H, 128, 32, CU_TENSOR_MAP_SWIZZLE_64B));
  VC_TRY(get_map(&mp.W[0], W0, (uint64_t)4 * H, (uint64_t)H, (uint64_t)H, 128, 2));
  VC_TRY(get_map(&mp.W[1], W1, (uint64_t)4 * H, (uint64_t)H, (uint64_t)H, 128, 2));
  PLstmArgs a;
  a.B = B; a.T = T; a.H = H;
  a.bias[0] = a.bias[1] = nullptr;
  a.xp = xp;
  a.flags = flags;
  a.dbg = dbg;
  const int NT = 4 * H / kPlBN, MT = (B + 255) / 256;
  dim3 grid(NT, MT, 2);
  VC_CUDA(cudaMemsetAsync(flags, 0, sizeof(unsigned int) * (size_t)4 * MT, stream));
  const size_t smem = plstm_smem_bytes(H);
  // VC_PLSTM_PAIR=1: the CTA-pair form (measured slower: 8.7 vs 7.2 us per step, see its header); read per call like the other switches
  const char* e = getenv("VC_PLSTM_PAIR");
  const bool pair = (pair_mode < 0 ? (e != nullptr && e[0] == '1') : pair_mode != 0) && NT % 2 == 0;
  if (pair) {
    const size_t smem = plstm_smem_bytes(H, kPlPairStages);
    VC_CUDA(cudaFuncSetAttribute(lstm_layer_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = dim3(kPlThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;            // all CTAs co-resident: the flag spins cannot starve a CTA that has not started
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    VC_CUDA(cudaLaunchKernelEx(&cfg, lstm_layer_pair_kernel, mp, a));
    return VC_OK;
  }
  // VC_PLSTM_CLUSTER=2: clusters of two gate-column tiles share every h_{t-1} box through TMA multicast (A/B testing)
  const char* ec = getenv("VC_PLSTM_CLUSTER");
  const bool cl2 = (cluster_mode < 0 ? (ec != nullptr && ec[0] == '2') : cluster_mode == 2) && NT % 2 == 0;
  if (cl2) {
    VC_CUDA(cudaFuncSetAttribute(lstm_layer_persistent_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = dim3(kPlThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = 2; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    VC_CUDA(cudaLaunchKernelEx(&cfg, lstm_layer_persistent_kernel<2>, mp, a));
    return VC_OK;
  }
  VC_CUDA(cudaFuncSetAttribute(lstm_layer_persistent_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&mp, (void*)&a};
  VC_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_layer_persistent_kernel<1>, grid, dim3(kPlThreads), args, smem, stream));
  return VC_OK;
}

}  // namespace tc
}  // namespace vc
