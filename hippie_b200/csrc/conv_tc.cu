// tcgen05 implicit-GEMM conv1d (forward and dgrad) with fp32-level accuracy via the 3xTF32 split.
//
//   C[m, n] (+)= sum_k A[row(m), k] * W[n, k]        A rows = contiguous k*Cin windows of the padded
//                                                   channels-last input, W = [N][K] K-major weights
//
// Per CTA: one 128-row x BN-column output tile, accumulators in TMEM (BN fp32 columns).
//   warp 0    TMA producer: per 32-float k-block one 3-D box of A ([samples][Lout][32 floats], 128B swizzle)
//             and one 2-D box of W into a STAGES-deep shared-memory ring
//   warp 1    TMEM allocator + MMA issuer: 3 x tcgen05.mma.kind::tf32 per 8-wide k-step
//             (lo*hi + hi*lo + hi*hi, fp32 accumulation in TMEM).  The tensor core truncates when it adds
//             into the fp32 accumulator, which biases a long K loop by ~n_steps * 2^-24 (measured: 4e-5 at
//             K = 1536); the hi*hi products therefore alternate between two accumulators per k-block and the
//             small lo terms use a third, and the epilogue adds the three with round-to-nearest.
//   warps 2-5 split every landed fp32 tile in place into hi = tf32-rounded value and a second buffer
//             lo = x - hi (exact in fp32), then run the epilogue: TMEM -> registers -> shared staging ->
//             coalesced global stores, bias / accumulate, BatchNorm (sum, centred M2) partials per column.
//
// Replaces the nn.Conv1d forward calls of hippie/backbones.py:11,24,26,31,50,55 and their input gradients.
#include <cuda.h>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace hp {

namespace {

using namespace tc;

struct TcConv {
  float* C;
  const float* bias;
  float* part;
  int B, N, K;        // samples, output channels, k*Cin
  int Lout, nb;       // logical rows per sample, samples per 128-row tile
  int out_rows, out_off, out_lstride, accumulate;
  int dbg;  // tools/tc_test timing experiments: 1 = skip the hi/lo split, 2 = hi*hi MMA only
};

// PASSES = 3: fp32-accurate 3xTF32 (stage = A_hi | A_lo | W_hi | W_lo);  PASSES = 1: single tf32 pass with
// round-to-nearest operands (stage = A | W), used for the backward GEMMs in the fast-backward mode.
template <int BN, int STAGES, int PASSES>
struct TcSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 4;  // 16 KB
  static constexpr int B_BYTES = BN * TC_BK * 4;
  static constexpr int A_LO = A_BYTES;                             // offset of A_lo (3-pass only)
  static constexpr int B_OFF = PASSES == 3 ? 2 * A_BYTES : A_BYTES;  // offset of W (hi)
  static constexpr int B_LO = B_OFF + B_BYTES;                     // offset of W_lo (3-pass only)
  static constexpr int STAGE_BYTES = PASSES == 3 ? 2 * A_BYTES + 2 * B_BYTES : A_BYTES + B_BYTES;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int EPI_STRIDE = BN + 4;
  static_assert(TC_BM * EPI_STRIDE * 4 <= RING_BYTES, "epilogue staging must fit in the ring");
  static constexpr int TOTAL = RING_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = PASSES == 3 ? (BN == 128 ? 512 : 256) : 2 * BN;  // 3 (or 2) accumulators
};

template <int BN, int STAGES, int PASSES>
__global__ void __launch_bounds__(TC_THREADS, 1)
    conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW, TcConv p) {
  using S = TcSmem<BN, STAGES, PASSES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + S::RING_BYTES);
  uint64_t* full = bars;                 // [STAGES]  TMA bytes landed
  uint64_t* conv = bars + STAGES;        // [STAGES]  hi/lo split written
  uint64_t* empty = bars + 2 * STAGES;   // [STAGES]  MMAs that read the stage retired
  uint64_t* accum = bars + 3 * STAGES;   // accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = blockIdx.x * p.nb, n0 = blockIdx.y * BN;
  const int nkb = p.K / TC_BK;
  const int rows_tile = p.nb * p.Lout;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapW);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&conv[s], 128);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(S::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)(rows_tile * TC_BK * 4 + S::B_BYTES);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        uint8_t* st = ring + s * S::STAGE_BYTES;
        mbar_expect_tx(&full[s], tx_bytes);
        tma_load_3d(st, &mapA, &full[s], kb * TC_BK, 0, b0);
        tma_load_2d(st + S::B_OFF, &mapW, &full[s], kb * TC_BK, n0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&full[s], ph);
        mbar_wait(&conv[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = smem_u32(ring + s * S::STAGE_BYTES);
        const uint64_t a_hi = umma_desc_k_sw128(st), a_lo = umma_desc_k_sw128(st + S::A_LO);
        const uint64_t b_hi = umma_desc_k_sw128(st + S::B_OFF), b_lo = umma_desc_k_sw128(st + S::B_LO);
#pragma unroll
        for (int k8 = 0; k8 < TC_BK / 8; ++k8) {
          const uint64_t adv = (uint64_t)(k8 * 32 >> 4);  // 8 tf32 = 32 bytes along the swizzled row
          if (PASSES == 3 && !(p.dbg & 2)) {
            umma_tf32(tmem_base + 2 * BN, a_lo + adv, b_hi + adv, idesc, (kb | k8) != 0 ? 1u : 0u);
            umma_tf32(tmem_base + 2 * BN, a_hi + adv, b_lo + adv, idesc, 1u);
          }
          umma_tf32(tmem_base + (kb & 1) * BN, a_hi + adv, b_hi + adv, idesc, (kb >= 2 || k8 != 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(accum);
    }
  } else {
    // ===================== split (hi / lo) warps, then epilogue =====================
    const int t = threadIdx.x - 64;  // 0..127
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
      mbar_wait(&full[s], ph);
      uint8_t* st = ring + s * S::STAGE_BYTES;
      // A: [hi | lo] at st, st + A_BYTES;  W: [hi | lo] at st + 2*A_BYTES, + B_BYTES.  Element-wise on raw bytes,
      // so the 128-byte swizzle the TMA applied is preserved.
      if (p.dbg & 1) {
        mbar_arrive(&conv[s]);
        continue;
      }
#pragma unroll
      for (int i = 0; i < S::A_BYTES / 16 / 128; ++i) {
        uint4* ph_ = reinterpret_cast<uint4*>(st) + t + i * 128;
        uint4 v = *ph_, h;
        const uint32_t rb = (p.dbg & 4) ? 0u : 0x1000u;  // dbg 4: truncate and leave the raw value as hi
        h.x = (v.x + rb) & 0xFFFFE000u, h.y = (v.y + rb) & 0xFFFFE000u;
        h.z = (v.z + rb) & 0xFFFFE000u, h.w = (v.w + rb) & 0xFFFFE000u;
        float4 l;
        l.x = __uint_as_float(v.x) - __uint_as_float(h.x), l.y = __uint_as_float(v.y) - __uint_as_float(h.y);
        l.z = __uint_as_float(v.z) - __uint_as_float(h.z), l.w = __uint_as_float(v.w) - __uint_as_float(h.w);
        if (!(p.dbg & 4)) *ph_ = h;
        if (PASSES == 3) *reinterpret_cast<float4*>(st + S::A_LO + (size_t)(t + i * 128) * 16) = l;
      }
#pragma unroll
      for (int i = 0; i < S::B_BYTES / 16 / 128; ++i) {
        uint4* ph_ = reinterpret_cast<uint4*>(st + S::B_OFF) + t + i * 128;
        uint4 v = *ph_, h;
        const uint32_t rb = (p.dbg & 4) ? 0u : 0x1000u;
        h.x = (v.x + rb) & 0xFFFFE000u, h.y = (v.y + rb) & 0xFFFFE000u;
        h.z = (v.z + rb) & 0xFFFFE000u, h.w = (v.w + rb) & 0xFFFFE000u;
        float4 l;
        l.x = __uint_as_float(v.x) - __uint_as_float(h.x), l.y = __uint_as_float(v.y) - __uint_as_float(h.y);
        l.z = __uint_as_float(v.z) - __uint_as_float(h.z), l.w = __uint_as_float(v.w) - __uint_as_float(h.w);
        if (!(p.dbg & 4)) *ph_ = h;
        if (PASSES == 3) *reinterpret_cast<float4*>(st + S::B_LO + (size_t)(t + i * 128) * 16) = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
      mbar_arrive(&conv[s]);
    }

    // ---- epilogue ----
    mbar_wait(accum, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* stage = reinterpret_cast<float*>(ring);  // [128][BN + 4]; the ring is idle once `accum` has fired
    const int q = warp & 3;                         // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32], r1[32], r2[32];
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
      tmem_ld32(ta, r);
      tmem_ld32(ta + BN, r1);
      if (PASSES == 3) {
        tmem_ld32(ta + 2 * BN, r2);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r2[i] = 0u;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i)
        r[i] = __float_as_uint((__uint_as_float(r[i]) + __uint_as_float(r1[i])) + __uint_as_float(r2[i]));
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<uint4*>(&stage[row * S::EPI_STRIDE + c * 32 + i * 4]) =
            make_uint4(r[i * 4], r[i * 4 + 1], r[i * 4 + 2], r[i * 4 + 3]);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only

    const int nvalid = min(rows_tile, (p.B - b0) * p.Lout);  // rows of this tile that are real outputs
    {  // coalesced stores: thread -> (row group, fixed column quad)
      constexpr int QUADS = BN / 4;
      const int quad = t % QUADS;
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias) bv = *reinterpret_cast<const float4*>(p.bias + n0 + quad * 4);
      for (int r = t / QUADS; r < nvalid; r += 128 / QUADS) {
        const int bs = r / p.Lout, l = r - bs * p.Lout;
        float4 v = *reinterpret_cast<const float4*>(&stage[r * S::EPI_STRIDE + quad * 4]);
        v.x += bv.x, v.y += bv.y, v.z += bv.z, v.w += bv.w;
        float4* dst = reinterpret_cast<float4*>(
            p.C + ((int64_t)(b0 + bs) * p.out_rows + p.out_off + (int64_t)l * p.out_lstride) * p.N + n0 + quad * 4);
        if (p.accumulate) {
          const float4 o = *dst;
          v.x += o.x, v.y += o.y, v.z += o.z, v.w += o.w;
        }
        *dst = v;
      }
    }
    if (p.part && t < BN) {  // BatchNorm statistics of this tile: (sum, centred sum of squares) per column
      const float bias = p.bias ? p.bias[n0 + t] : 0.f;
      // four independent partial sums keep the shared-memory loads in flight (a single dependent chain of
      // 128 adds costs ~4 us per tile)
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int r = 0;
      for (; r + 4 <= nvalid; r += 4) {
        s0 += stage[(r + 0) * S::EPI_STRIDE + t], s1 += stage[(r + 1) * S::EPI_STRIDE + t];
        s2 += stage[(r + 2) * S::EPI_STRIDE + t], s3 += stage[(r + 3) * S::EPI_STRIDE + t];
      }
      for (; r < nvalid; ++r) s0 += stage[r * S::EPI_STRIDE + t];
      const float s = (s0 + s1) + (s2 + s3);
      const float mean = s / (float)nvalid;
      float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
      for (r = 0; r + 4 <= nvalid; r += 4) {
        const float d0 = stage[(r + 0) * S::EPI_STRIDE + t] - mean, d1 = stage[(r + 1) * S::EPI_STRIDE + t] - mean;
        const float d2 = stage[(r + 2) * S::EPI_STRIDE + t] - mean, d3 = stage[(r + 3) * S::EPI_STRIDE + t] - mean;
        q0 = fmaf(d0, d0, q0), q1 = fmaf(d1, d1, q1), q2 = fmaf(d2, d2, q2), q3 = fmaf(d3, d3, q3);
      }
      for (; r < nvalid; ++r) {
        const float d = stage[r * S::EPI_STRIDE + t] - mean;
        q0 = fmaf(d, d, q0);
      }
      const float m2 = (q0 + q1) + (q2 + q3);
      *reinterpret_cast<float2*>(p.part + ((int64_t)blockIdx.x * p.N + n0 + t) * 2) =
          make_float2(s + bias * (float)nvalid, m2);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(S::TMEM_COLS) : "memory");
  }
}


// ------------------------------------------------------------------------------------------------
// Weight gradient on tcgen05:  dW[m, n] += sum_r dY[r, m] * X[(r + roff) * Cin + n]   (split over row ranges)
//
// Both operands are "MN-major" for the tensor core: the reduction index r is the slow (row) index of the
// channels-last tensors.  A tile = 128 output channels x 32 rows is fetched as 4 groups of [32 rows][32 floats]
// (one 3-D TMA box, 128-byte swizzle), the canonical UMMA Major-MN SWIZZLE_128B layout; one MMA consumes 8 rows.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t ltype) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;  // leading byte offset: next group of 32 MN elements
  d |= (uint64_t)(sbo >> 4) << 32;  // stride byte offset: next group of K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)ltype << 61;
  return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32_mn(int n) {
  return umma_idesc_tf32(n) | (1u << 15) | (1u << 16);  // A and B MN-major
}

struct TcWgrad {
  float* dW;
  int M, N;            // Cout, k*Cin
  int R;               // padded rows to reduce over
  int rows_per_split;  // multiple of TC_BK
  uint32_t lbo, sbo, ltype;  // shared-memory descriptor fields of the MN-major operands
};

template <int BN, int STAGES, int PASSES>
__global__ void __launch_bounds__(TC_THREADS, 1)
    wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapDY, const __grid_constant__ CUtensorMap mapX, TcWgrad p) {
  using S = TcSmem<BN, STAGES, PASSES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + S::RING_BYTES);
  uint64_t* full = bars;
  uint64_t* conv = bars + STAGES;
  uint64_t* empty = bars + 2 * STAGES;
  uint64_t* accum = bars + 3 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * TC_BM;
  const int r_begin = blockIdx.z * p.rows_per_split;
  const int r_end = min(p.R, r_begin + p.rows_per_split);
  const int nkb = (r_end - r_begin + TC_BK - 1) / TC_BK;  // >= 1 by construction of the grid

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapDY);
    tma_prefetch_desc(&mapX);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&conv[s], 128);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(S::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        uint8_t* st = ring + s * S::STAGE_BYTES;
        mbar_expect_tx(&full[s], (uint32_t)(S::A_BYTES + S::B_BYTES));
        // one [32 rows][32 channels] box per group of 32 channels (groups land 4 KB apart = the descriptor's LBO)
#pragma unroll
        for (int g = 0; g < TC_BM / 32; ++g)
          tma_load_2d(st + g * (TC_BK * 128), &mapDY, &full[s], m0 + g * 32, r_begin + kb * TC_BK);
#pragma unroll
        for (int g = 0; g < BN / 32; ++g)
          tma_load_2d(st + S::B_OFF + g * (TC_BK * 128), &mapX, &full[s], n0 + g * 32, r_begin + kb * TC_BK);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32_mn(BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&full[s], ph);
        mbar_wait(&conv[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = smem_u32(ring + s * S::STAGE_BYTES);
        const uint64_t a_hi = umma_desc_mn(st, p.lbo, p.sbo, p.ltype), a_lo = umma_desc_mn(st + S::A_LO, p.lbo, p.sbo, p.ltype);
        const uint64_t b_hi = umma_desc_mn(st + S::B_OFF, p.lbo, p.sbo, p.ltype);
        const uint64_t b_lo = umma_desc_mn(st + S::B_LO, p.lbo, p.sbo, p.ltype);
#pragma unroll
        for (int k8 = 0; k8 < TC_BK / 8; ++k8) {
          const uint64_t adv = (uint64_t)(k8 * 1024 >> 4);  // 8 reduction rows = 8 x 128 B
          if (PASSES == 3) {
            umma_tf32(tmem_base + 2 * BN, a_lo + adv, b_hi + adv, idesc, (kb | k8) != 0 ? 1u : 0u);
            umma_tf32(tmem_base + 2 * BN, a_hi + adv, b_lo + adv, idesc, 1u);
          }
          umma_tf32(tmem_base + (kb & 1) * BN, a_hi + adv, b_hi + adv, idesc, (kb >= 2 || k8 != 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(accum);
    }
  } else {
    const int t = threadIdx.x - 64;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
      mbar_wait(&full[s], ph);
      uint8_t* st = ring + s * S::STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < S::A_BYTES / 16 / 128; ++i) {
        uint4* ph_ = reinterpret_cast<uint4*>(st) + t + i * 128;
        uint4 v = *ph_, h;
        h.x = (v.x + 0x1000u) & 0xFFFFE000u, h.y = (v.y + 0x1000u) & 0xFFFFE000u;
        h.z = (v.z + 0x1000u) & 0xFFFFE000u, h.w = (v.w + 0x1000u) & 0xFFFFE000u;
        float4 l;
        l.x = __uint_as_float(v.x) - __uint_as_float(h.x), l.y = __uint_as_float(v.y) - __uint_as_float(h.y);
        l.z = __uint_as_float(v.z) - __uint_as_float(h.z), l.w = __uint_as_float(v.w) - __uint_as_float(h.w);
        *ph_ = h;
        if (PASSES == 3) *reinterpret_cast<float4*>(st + S::A_LO + (size_t)(t + i * 128) * 16) = l;
      }
#pragma unroll
      for (int i = 0; i < S::B_BYTES / 16 / 128; ++i) {
        uint4* ph_ = reinterpret_cast<uint4*>(st + S::B_OFF) + t + i * 128;
        uint4 v = *ph_, h;
        h.x = (v.x + 0x1000u) & 0xFFFFE000u, h.y = (v.y + 0x1000u) & 0xFFFFE000u;
        h.z = (v.z + 0x1000u) & 0xFFFFE000u, h.w = (v.w + 0x1000u) & 0xFFFFE000u;
        float4 l;
        l.x = __uint_as_float(v.x) - __uint_as_float(h.x), l.y = __uint_as_float(v.y) - __uint_as_float(h.y);
        l.z = __uint_as_float(v.z) - __uint_as_float(h.z), l.w = __uint_as_float(v.w) - __uint_as_float(h.w);
        *ph_ = h;
        if (PASSES == 3) *reinterpret_cast<float4*>(st + S::B_LO + (size_t)(t + i * 128) * 16) = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&conv[s]);
    }

    mbar_wait(accum, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* stage = reinterpret_cast<float*>(ring);
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool two = nkb >= 2;  // the second hi*hi accumulator is only written when there are two k-blocks
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32], r1[32], r2[32];
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
      tmem_ld32(ta, r);
      tmem_ld32(ta + BN, r1);
      if (PASSES == 3) {
        tmem_ld32(ta + 2 * BN, r2);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r2[i] = 0u;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i)
        r[i] = __float_as_uint((__uint_as_float(r[i]) + (two ? __uint_as_float(r1[i]) : 0.f)) + __uint_as_float(r2[i]));
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<uint4*>(&stage[row * S::EPI_STRIDE + c * 32 + i * 4]) =
            make_uint4(r[i * 4], r[i * 4 + 1], r[i * 4 + 2], r[i * 4 + 3]);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int mvalid = min(TC_BM, p.M - m0);
    const int col = t % BN;
    for (int r = t / BN; r < mvalid; r += 128 / BN)
      atomicAdd(p.dW + (int64_t)(m0 + r) * p.N + n0 + col, stage[r * S::EPI_STRIDE + col]);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(S::TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_conv_dbg = 0;
// MN-major fp32/tf32 operands: the tensor core transposes at 32-byte granularity, so the pairing is
// TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B <-> UMMA layout type SWIZZLE_128B_BASE32B (1), groups of 32 channels
// 4 KB apart (LBO), groups of 4 reduction rows 512 B apart (SBO).  Found by tools/tc_test's probe on a B200: the
// plain SWIZZLE_128B pairing (type 2) yields zeros for MN-major tf32.
uint32_t g_wg_lbo = TC_BK * 128, g_wg_sbo = 512, g_wg_ltype = 1;
CUtensorMapSwizzle g_wg_swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;

}  // namespace

bool tc_init(std::string* err) {
  if (g_encode) return true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    if (err) *err = "cuTensorMapEncodeTiled is not available from the driver";
    return false;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
#define HP_SET_SMEM(kern, BN, ST, PS) \
  cudaFuncSetAttribute(kern<BN, ST, PS>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<BN, ST, PS>::TOTAL)
  HP_SET_SMEM(conv_gemm_tc_kernel, 128, 3, 3), HP_SET_SMEM(conv_gemm_tc_kernel, 64, 4, 3);
  HP_SET_SMEM(conv_gemm_tc_kernel, 128, 6, 1), HP_SET_SMEM(conv_gemm_tc_kernel, 64, 8, 1);
  HP_SET_SMEM(wgrad_tc_kernel, 128, 3, 3), HP_SET_SMEM(wgrad_tc_kernel, 64, 4, 3);
  HP_SET_SMEM(wgrad_tc_kernel, 128, 6, 1), HP_SET_SMEM(wgrad_tc_kernel, 64, 8, 1);
#undef HP_SET_SMEM
  return cudaGetLastError() == cudaSuccess;
}

// 3-D map over a padded channels-last tensor: (k within the k*Cin window, logical output row, sample)
bool tc_make_act_map(TcMap* out, const float* base, int in_C, int K, int Lout, int in_rows, int in_stride, int in_off,
                     int max_batch) {
  const int nb = TC_BM / Lout;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)Lout, (cuuint64_t)max_batch};
  cuuint64_t strides[2] = {(cuuint64_t)in_stride * in_C * 4, (cuuint64_t)in_rows * in_C * 4};
  cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)Lout, (cuuint32_t)nb};
  cuuint32_t estr[3] = {1, 1, 1};
  void* gptr = const_cast<float*>(base + (int64_t)in_off * in_C);
  CUresult r = g_encode(reinterpret_cast<CUtensorMap*>(out->opaque), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, gptr, dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// 2-D map over K-major weights [N][K]; box = 32 floats x bn rows
bool tc_make_weight_map(TcMap* out, const float* w, int N, int K, int bn) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(reinterpret_cast<CUtensorMap*>(out->opaque), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                        const_cast<float*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// 2-D map for the wgrad operands: (channel, reduction row) with box 32 channels x 32 rows.  For the input tensor the
// "row" is the k*Cin-float window starting at that padded row, so consecutive rows overlap (row pitch = Cin floats).
bool tc_make_rows_map(TcMap* out, const float* base, int row_floats, int channels, int rows, int /*box_groups*/) {
  cuuint64_t dims[2] = {(cuuint64_t)channels, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_floats * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)TC_BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(reinterpret_cast<CUtensorMap*>(out->opaque), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                        const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        g_wg_swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

void tc_debug_conv(int dbg) { g_conv_dbg = dbg; }

void tc_debug_wgrad_knobs(uint32_t lbo, uint32_t sbo, uint32_t ltype, int tma_swizzle) {
  g_wg_lbo = lbo, g_wg_sbo = sbo, g_wg_ltype = ltype, g_wg_swz = (CUtensorMapSwizzle)tma_swizzle;
}

void launch_wgrad_tc(const WgradGemm& g, const TcMap& mapDY, const TcMap& mapX, int bn, int sm_count, int passes,
                     cudaStream_t s) {
  TcWgrad p{};
  p.dW = g.dW, p.M = g.M, p.N = g.N, p.R = g.R;
  p.lbo = g_wg_lbo, p.sbo = g_wg_sbo, p.ltype = g_wg_ltype;
  const int tiles = ((g.M + TC_BM - 1) / TC_BM) * (g.N / bn);
  // one wave: every CTA pays ~6-9 us of prologue + epilogue, so fewer, longer CTAs beat more splits
  int splits = sm_count / tiles;
  const int kblocks = (g.R + TC_BK - 1) / TC_BK;
  if (splits > (kblocks + 7) / 8) splits = (kblocks + 7) / 8;  // at least 8 k-blocks (256 rows) per CTA
  if (splits < 1) splits = 1;
  int kb_per = (kblocks + splits - 1) / splits;
  p.rows_per_split = kb_per * TC_BK;
  splits = (g.R + p.rows_per_split - 1) / p.rows_per_split;
  dim3 grid(g.N / bn, (g.M + TC_BM - 1) / TC_BM, splits);
  const CUtensorMap& a = *reinterpret_cast<const CUtensorMap*>(mapDY.opaque);
  const CUtensorMap& x = *reinterpret_cast<const CUtensorMap*>(mapX.opaque);
  if (passes == 3) {
    if (bn == 128)
      wgrad_tc_kernel<128, 3, 3><<<grid, TC_THREADS, TcSmem<128, 3, 3>::TOTAL, s>>>(a, x, p);
    else
      wgrad_tc_kernel<64, 4, 3><<<grid, TC_THREADS, TcSmem<64, 4, 3>::TOTAL, s>>>(a, x, p);
  } else {
    if (bn == 128)
      wgrad_tc_kernel<128, 6, 1><<<grid, TC_THREADS, TcSmem<128, 6, 1>::TOTAL, s>>>(a, x, p);
    else
      wgrad_tc_kernel<64, 8, 1><<<grid, TC_THREADS, TcSmem<64, 8, 1>::TOTAL, s>>>(a, x, p);
  }
}

int tc_pick_bn(int B, int N, int Lout, int sm_count) {
  if (N % 128 != 0) return 64;
  const int nb = TC_BM / Lout;
  const int mtiles = (B + nb - 1) / nb;
  return (mtiles * (N / 128) * 4 >= sm_count * 3) ? 128 : 64;  // fall back to 64-wide tiles when the grid is too small
}

// returns the number of logical rows per statistics tile
int launch_conv_gemm_tc(const ConvGemm& g, const TcMap& mapA, const TcMap& mapW, int bn, int B, int passes,
                        cudaStream_t s) {
  TcConv p{};
  p.C = g.C, p.bias = g.bias, p.part = g.part, p.B = B, p.N = g.N, p.K = g.K, p.Lout = g.Lout;
  p.nb = TC_BM / g.Lout;
  p.out_rows = g.out_rows, p.out_off = g.out_off, p.out_lstride = g.out_lstride, p.accumulate = g.accumulate;
  p.dbg = g_conv_dbg;
  dim3 grid((B + p.nb - 1) / p.nb, g.N / bn);
  const CUtensorMap& a = *reinterpret_cast<const CUtensorMap*>(mapA.opaque);
  const CUtensorMap& w = *reinterpret_cast<const CUtensorMap*>(mapW.opaque);
  if (passes == 3) {
    if (bn == 128)
      conv_gemm_tc_kernel<128, 3, 3><<<grid, TC_THREADS, TcSmem<128, 3, 3>::TOTAL, s>>>(a, w, p);
    else
      conv_gemm_tc_kernel<64, 4, 3><<<grid, TC_THREADS, TcSmem<64, 4, 3>::TOTAL, s>>>(a, w, p);
  } else {
    if (bn == 128)
      conv_gemm_tc_kernel<128, 6, 1><<<grid, TC_THREADS, TcSmem<128, 6, 1>::TOTAL, s>>>(a, w, p);
    else
      conv_gemm_tc_kernel<64, 8, 1><<<grid, TC_THREADS, TcSmem<64, 8, 1>::TOTAL, s>>>(a, w, p);
  }
  return p.nb * g.Lout;
}

}  // namespace hp
