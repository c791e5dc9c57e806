// tcgen05 implicit-GEMM conv1d (forward, dgrad) and weight gradient over 16-bit PAIR PLANES.
//
// Every GEMM operand lives in HBM as two 16-bit planes (hi, lo) with  x ~= hi + lo :
//   activations fp16 pair (scale 1), weights fp16 pair (scale 2^8), output gradients fp16 pair scaled by a power of
//   two chosen on the device from a bound of max|g| (bn_bwd_apply_kernel)                      -> ~22 mantissa bits
// (mixed fp16 x bf16 operands are an illegal instruction on sm_100a; a bf16 pair is supported by the kernels and used by
// tools/pair_test only) and one product is three tensor-core MMAs  hi*hi + hi*lo + lo*hi  (kind::f16, K = 16, fp32 accumulation in TMEM).
// tools/pair_precision.py shows the fp16 pair matches fp32 operands on the loss / embedding bounds (like 3xTF32),
// at HALF the tensor time and ~1/3 of the shared-memory traffic of an in-kernel 3xTF32 split: the kernels below are
// pure TMA -> MMA pipelines, the producers of the tensors write the planes (kernels_ew.cu).
//
//   C[m, n] (+)= scale * sum_k A[row(m), k] * B[n, k]
//     A rows  = contiguous k*Cin windows of the padded channels-last pair tensor (3-D TMA box per plane)
//     B       = K-major weight planes [N][K] (forward), or the SAME planes read MN-major as [ci][(t, co)] (dgrad:
//               no transposed weight copy exists)
//
// Per CTA (192 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-5 epilogue
// (TMEM -> registers -> swizzled shared staging -> ONE elected thread issues TMA tensor stores of the whole tile
// (add-reduce when the layer accumulates into its output) while the four warps compute the BatchNorm (sum, centred M2)
// partials from the same staging; the eval-mode fold uses per-thread stores).
// The B planes of a stage sit next to each other in shared memory, so ONE descriptor spans [B hi | B lo] as 2 * BN
// columns: a k-step is   A hi x [B hi | B lo]  (N = 2 BN: hi*hi and hi*lo land side by side in tensor memory)  and
// A lo x B hi  (N = BN, added onto the hi*lo columns).  Same tensor math as three N = BN instructions, but the A hi
// tile is read from shared memory once instead of twice: 14 KB instead of 18 KB per k-step through the 128 B/clk
// port that bounds the main loop (tools/mma_rate2: 48 / 64 cycles per N = 64 / 128 instruction).
// The tensor core truncates when it adds into the fp32 accumulator; to keep that bias below the parity bound
// consecutive k-steps alternate between two accumulator sets (summed with round-to-nearest in the epilogue).
//
// Replaces the nn.Conv1d forward calls of hippie/backbones.py:11,24,26,31,50,55 and their autograd backward.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include "kernels.cuh"
#include "pair_fmt.cuh"
#include "pdl.cuh"
#include "tc_common.cuh"

namespace hp {

namespace {

using namespace tc;

constexpr int PK = 64;  // K elements per k-block = one 128-byte swizzle row of 16-bit values
// Tile configuration: 128 x 64 output tile, 2-stage ring (96 KB), 256 TMEM columns -> TWO CTAs per SM, so that the
// prologue / epilogue of one CTA (or of the next kernel, see pdl.cuh) overlaps the main loop of the other and the two
// branches of the model (and the weight gradients) share every SM.  Tiles that own the SM (128 x 128, 3 stages) are
// 13-17 % faster on the deep layers in isolation and slower in every real pass (round 1: 4.77 vs 4.36 ms per bs512 step;
// round 2: embedding pass -1.5 % at bs4096, -9 % at bs512, profiles/r02_exp46.txt, r02_exp47.txt).
// Phase stamps of one CTA (tools/pair_test, PT_STAMPS=1): prologue 0.26 us, first fill 0.67 us, main loop 0.42-0.46 us per
// k-block alone and ~0.55 us when every SM holds two CTAs (the chip-wide L2 -> SM rate: 2 x 48 KB per k-block pair at
// ~42 B/clk per SM), epilogue 1.8 us (TMEM -> staging 0.45, store issue 0.16, BatchNorm partials 1.2).
// 32-wide k-blocks (64-byte swizzle, 4 stages: tried) cost 0.38 us per half block; deeper rings change nothing.
#ifndef HP_KBN
#define HP_KBN 64
#endif
#ifndef HP_KSTAGES
#define HP_KSTAGES 2
#endif
constexpr int kBN = HP_KBN, kStages = HP_KSTAGES;  // overridable for tools/pair_test experiments
// Grids of at most one CTA per SM run a deeper ring (kStagesAlone, one CTA per SM by its shared-memory footprint):
// the block scheduler otherwise packs part of such a grid two to an SM, and those CTAs share one SM's ~64 B/clk of
// L2 -> shared-memory ingest while other SMs idle (512->512 L=4, 128 CTAs: 21.0 -> 16.1 us, tools/pair_test).
constexpr int kStagesAlone = HP_KSTAGES > 3 ? HP_KSTAGES : 3;
constexpr int ctas_per_sm(int bn, int stages) { return (PK * 2 * (128 + bn) * 2 * stages <= 100 * 1024 && bn <= 64) ? 2 : 1; }
// Register budget: the kernels are compiled as if THREE CTAs had to fit on an SM (ptxas settles on 96 registers, 4-8 bytes of
// spill).  Two fit by shared memory; the registers they leave free decide whether a CTA of an elementwise kernel (BatchNorm
// apply / backward, on the critical chains) can start on an SM that already holds two GEMM CTAs: 154 registers 3.06 ms per
// bs512 step, 96 registers 2.98 ms, 80 registers (76-84 bytes of spill) 3.02 ms (profiles/r02_exp32.txt - r02_exp34.txt).
constexpr int reg_ctas_per_sm(int bn, int stages) { return ctas_per_sm(bn, stages) == 2 ? 3 : 1; }

struct PairConv {
  float* C;
  const float* bias;
  float* part;
  double* tot;
  int B, N, K;
  int Lout, nb;
  int out_rows, out_off, out_lstride, accumulate;
  float out_scale;
  uint32_t idesc;   // N = BN       (A lo x B hi)
  uint32_t idesc2;  // N = 2 * BN   (A hi x [B hi | B lo])
  int nmma;         // 3: all three products; 2: experiment, A lo x B hi dropped (A at 11 bits)
  int fused;        // 1: A hi x [B hi | B lo] as one N = 2 BN instruction; 0: two N = BN instructions
  int fake_reuse, fake_k;
  int b_mn;        // 1: B tiles are MN-major boxes of the forward weight planes [co][t][ci]
  int kb_per_tap;  // b_mn: k-blocks per tap (= Cout / 64)
  int taps;        // b_mn: kernel size (3 or 1); tap of k-block kb = taps - 1 - kb / kb_per_tap
  const float* dyn_scale;  // optional device scalar multiplied into out_scale
  unsigned long long* stamps;  // tools/pair_test only: %globaltimer at the phase boundaries of CTA (0, 0)
  int pdl_late;                // trigger the dependent launch when the main loop is issued instead of at kernel start
  int stamps_all;              // every CTA writes its stamps (8 slots per CTA, slot 7 = %smid)
  EvalFold fold;               // eval-mode BatchNorm (+ residual, LeakyReLU, pair planes) applied in the epilogue
  int tma_out;                 // mapC describes the fp32 output: whole tiles leave through TMA stores
};
__device__ __forceinline__ void stamp(const PairConv& p, int i) {
  if (p.stamps && (p.stamps_all || (blockIdx.x == 0 && blockIdx.y == 0))) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    const size_t base = p.stamps_all ? (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 : 0;
    p.stamps[base + i] = t;
    if (p.stamps_all && i == 0) {  // slot 7: the SM this CTA runs on
      unsigned smid;
      asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      p.stamps[base + 7] = smid;
    }
  }
}

template <int BN, int STAGES>
struct PairSmem {
  static constexpr int A_PLANE = TC_BM * 128;  // 16 KB: 128 rows x 64 halfs
  static constexpr int B_PLANE = BN * 128;
  static constexpr int A_LO = A_PLANE;
  static constexpr int B_OFF = 2 * A_PLANE;
  static constexpr int B_LO = B_OFF + B_PLANE;
  static constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  // epilogue staging: BN / 32 boxes of [128 rows][32 floats] in the 128-byte swizzle a TMA store expects
  // (16-byte chunk j of row r sits at chunk j ^ (r & 7)): row-per-thread writes and column reads are conflict-free
  static constexpr int EPI_BOX_FLOATS = TC_BM * 32;
  static_assert((BN / 32) * EPI_BOX_FLOATS * 4 <= RING_BYTES, "epilogue staging must fit in the ring");
  __device__ static __forceinline__ int epi(int r, int col) {
    return (col >> 5) * EPI_BOX_FLOATS + r * 32 + (((((col & 31) >> 2) ^ (r & 7))) << 2) + (col & 3);
  }
  static constexpr int TOTAL = RING_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 4 * BN;  // two accumulator sets of [hi*hi | cross terms], BN columns each
};

// MODE: kConvTrain = training forward (K-major weights, bias, BatchNorm partials, TMA stores), kConvEval = eval forward
// (BatchNorm folded into the epilogue, per-thread stores, no statistics), kConvDgrad = dgrad (MN-major weights, dynamic
// scale, TMA store / add-reduce).  One instantiation per mode keeps each pipeline's code lean: with all three in one
// kernel the eval epilogue ran 30 % slower (register allocation), measured on the embedding pass.
constexpr int kConvTrain = 0, kConvEval = 1, kConvDgrad = 2;
constexpr int kConvProducers = 4;   // TMA-issuing threads of conv_pair_kernel (warps 0, 2, 3, 4)
constexpr int kWgradProducers = 5;  // of wgrad_pair_kernel (warps 0, 2, 3, 4, 5)
template <int BN, int STAGES, int MODE>
__global__ void __launch_bounds__(TC_THREADS, reg_ctas_per_sm(BN, STAGES))
    conv_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                     const __grid_constant__ CUtensorMap mapC, PairConv p) {
  constexpr bool EVAL = MODE == kConvEval, B_MN = MODE == kConvDgrad;
  using S = PairSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + S::RING_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* accum = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  if (!p.pdl_late) pdl_trigger();
  if (threadIdx.x == 0) stamp(p, 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = blockIdx.x * p.nb, n0 = blockIdx.y * BN;
  const int nkb = p.fake_k ? min(p.fake_k, p.K / PK) : p.K / PK;  // fake_k: SPEED EXPERIMENT ONLY (main loops cut short)
  const int rows_tile = p.nb * p.Lout;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], kConvProducers);
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
  pdl_wait();  // everything above touched only shared / tensor memory
  if (threadIdx.x == 0) stamp(p, 1);

  // ===================== TMA producers =====================
  // One thread issues its tensor loads one after the other: a box costs ~235 cycles + 0.9 per 128-byte row whatever the
  // ring depth (tools/tma_bw: 47 B/clk from one thread, up to ~90 B/clk per SM from several), and a k-block is four
  // boxes.  So each of the four operand planes gets its own issuing thread in its own warp (warp 0 and the first three
  // epilogue warps, which have nothing to do until the accumulator is complete).
  if (warp != 1 && warp <= kConvProducers && lane == 0) {
    const int pid = warp == 0 ? 0 : warp - 1;  // 0: A hi, 1: A lo, 2: B hi, 3: B lo
    const uint32_t tx_bytes = pid < 2 ? (uint32_t)(rows_tile * 128) : (uint32_t)S::B_PLANE;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
      mbar_wait(&empty[s], ph ^ 1u);
      uint8_t* st = ring + s * S::STAGE_BYTES;
      if (p.fake_reuse && pid < 2 && kb >= STAGES && kb % 3 != 0) {  // SPEED EXPERIMENT ONLY (wrong results): A fetched for one k-block in three
        mbar_arrive(&full[s]);
        continue;
      }
      mbar_expect_tx(&full[s], tx_bytes);
      if (pid < 2) {
        tma_load_4d(st + pid * S::A_LO, &mapA, &full[s], kb * PK, 0, b0, pid);
      } else if (!B_MN) {
        tma_load_3d(st + (pid == 2 ? S::B_OFF : S::B_LO), &mapB, &full[s], kb * PK, n0, pid - 2);
      } else {
        const int u = kb / p.kb_per_tap, cob = kb - u * p.kb_per_tap;
        const int tap = p.taps - 1 - u;
#pragma unroll
        for (int g = 0; g < BN / 64; ++g)  // one [64 co][64 ci] box per group of 64 output columns
          tma_load_4d(st + (pid == 2 ? S::B_OFF : S::B_LO) + g * 8192, &mapB, &full[s], n0 + g * 64, tap, cob * 64, pid - 2);
      }
    }
  }
  __syncwarp();

  if (warp == 0) {
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop and an elected lane issues: inside an `if (lane == 0)` branch the compiler wraps
    // every UTCHMMA in an ELECT / R2UR / BRA.U.ANY loop (~90 cycles per MMA instead of the 48 of the tensor pipe at
    // N = 64, tools/mma_rate2).
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
      mbar_wait(&full[s], ph);
      if (kb == 0 && lane == 0) stamp(p, 2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t st = smem_u32(ring + s * S::STAGE_BYTES);
        const uint64_t a_hi = umma_desc(st, 16, 1024, 2), a_lo = umma_desc(st + S::A_LO, 16, 1024, 2);
        // the B lo plane follows the B hi plane: the same descriptor read with N = 2 BN covers [B hi | B lo]
        constexpr uint64_t blo = (uint64_t)(S::B_PLANE >> 4);  // descriptor offset of the B lo plane
        uint64_t b_hi, badv;
        if (!B_MN) {  // K-major: 8-row groups 1024 B apart, 16 halfs = 32 B along the swizzled row per step
          b_hi = umma_desc(st + S::B_OFF, 16, 1024, 2);
          badv = 32 >> 4;
        } else {  // MN-major: groups of 64 columns 8 KB apart (LBO), 8 k-rows 1024 B apart (SBO), 16 k-rows per step
          b_hi = umma_desc(st + S::B_OFF, 8192, 1024, 2);
          badv = 2048 >> 4;
        }
#pragma unroll
        for (int k16 = 0; k16 < PK / 16; ++k16) {
          const uint64_t aadv = (uint64_t)(k16 * 32 >> 4), bad = (uint64_t)k16 * badv;
          const int step = kb * (PK / 16) + k16;
          const uint32_t acc = tmem_base + (uint32_t)((step & 1) * 2 * BN);
          if (p.fused) {
            umma_f16(acc, a_hi + aadv, b_hi + bad, p.idesc2, step >= 2 ? 1u : 0u);  // [hi*hi | hi*lo]
          } else {
            umma_f16(acc, a_hi + aadv, b_hi + bad, p.idesc, step >= 2 ? 1u : 0u);
            umma_f16(acc + BN, a_hi + aadv, b_hi + blo + bad, p.idesc, step >= 2 ? 1u : 0u);
          }
          if (p.nmma == 3) umma_f16(acc + BN, a_lo + aadv, b_hi + bad, p.idesc, 1u);  // + lo*hi
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(accum);
    __syncwarp();
    if (p.pdl_late) pdl_trigger();  // this CTA is about to enter its epilogue
  } else {
    // ===================== epilogue =====================
    const int t = threadIdx.x - 64;  // 0..127
    mbar_wait(accum, 0);
    if (t == 0) stamp(p, 3);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* stage = reinterpret_cast<float*>(ring);  // swizzled boxes (PairSmem::epi); the ring is idle once `accum` has fired
    const int q = warp & 3;                         // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const float sc = p.dyn_scale ? p.out_scale * __ldg(p.dyn_scale) : p.out_scale;
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      // two live register arrays: the register footprint of these CTAs decides whether an elementwise kernel's CTA still
      // fits on an SM that holds two of them (122 -> 154 registers cost 4 % of the bs512 step, gpurun_out/r02_exp32.txt)
      uint32_t r[32], r1[32];
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
      tmem_ld32(ta + BN, r);       // cross terms, even k-steps
      tmem_ld32(ta + 3 * BN, r1);  // cross terms, odd k-steps
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r1[i]));
      tmem_ld32(ta, r1);           // hi*hi, even k-steps
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r1[i]));
      tmem_ld32(ta + 2 * BN, r1);  // hi*hi, odd k-steps
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = __float_as_uint((__uint_as_float(r[i]) + __uint_as_float(r1[i])) * sc);
      if (!B_MN && p.bias) {  // warp-uniform addresses: broadcast loads
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c * 32 + i * 4));
          r[i * 4 + 0] = __float_as_uint(__uint_as_float(r[i * 4 + 0]) + bv.x);
          r[i * 4 + 1] = __float_as_uint(__uint_as_float(r[i * 4 + 1]) + bv.y);
          r[i * 4 + 2] = __float_as_uint(__uint_as_float(r[i * 4 + 2]) + bv.z);
          r[i * 4 + 3] = __float_as_uint(__uint_as_float(r[i * 4 + 3]) + bv.w);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<uint4*>(&stage[S::epi(row, c * 32 + i * 4)]) =
            make_uint4(r[i * 4], r[i * 4 + 1], r[i * 4 + 2], r[i * 4 + 3]);
    }
    const int nvalid = min(rows_tile, (p.B - b0) * p.Lout);  // rows of this tile that are real outputs
    // training-mode tiles go out through the TMA unit (it skips the pad rows by construction of mapC)
    const bool tma_out = !EVAL && p.tma_out;  // mapC ends at the batch size: a ragged last tile is clipped by the unit
    if (tma_out) fence_proxy_async_smem();
    asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
    if (t == 0) stamp(p, 4);

    if (tma_out) {
      if (t == 0) {
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          if (B_MN && p.accumulate)
            tma_reduce_add_3d(&mapC, stage + c * S::EPI_BOX_FLOATS, n0 + c * 32, 0, b0);
          else
            tma_store_3d(&mapC, stage + c * S::EPI_BOX_FLOATS, n0 + c * 32, 0, b0);
        }
        tma_store_commit();
      }
    } else {  // coalesced stores: thread -> (row group, fixed column quad)
      constexpr int QUADS = BN / 4;
      const int quad = t % QUADS;
      const float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      const EvalFold& fo = p.fold;
      float4 fsc = bv, fbe = bv, fmu = bv;
      if (EVAL && fo.coef) {
        fsc = *reinterpret_cast<const float4*>(fo.coef + 0 * p.N + n0 + quad * 4);
        fbe = *reinterpret_cast<const float4*>(fo.coef + 1 * p.N + n0 + quad * 4);
        fmu = *reinterpret_cast<const float4*>(fo.coef + 2 * p.N + n0 + quad * 4);
      }
      constexpr int RSTEP = 128 / QUADS;  // rows covered by one pass of the 128 threads
      if (EVAL && fo.coef) {
        // Four rows per thread and iteration, every load (staging, residual) issued before the first use: with one row per
        // iteration the residual's L2 round trip (~0.6 us) was paid 16 times per tile in sequence and this epilogue, not the
        // main loop, bounded the shallow layers of the embedding pass (K = 192: 1.3 us of MMAs per tile).
        constexpr int U = 4;
        for (int r0 = t / QUADS; r0 < nvalid; r0 += U * RSTEP) {
          float4 v[U], rr[U];
          int64_t off[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int r = min(r0 + u * RSTEP, nvalid - 1);  // clamped rows are loaded twice and stored never
            const int bs = r / p.Lout, l = r - bs * p.Lout;
            off[u] = ((int64_t)(b0 + bs) * p.out_rows + p.out_off + (int64_t)l * p.out_lstride) * p.N + n0 + quad * 4;
            v[u] = *reinterpret_cast<const float4*>(&stage[S::epi(r, quad * 4)]);
            rr[u] = fo.res ? __ldg(reinterpret_cast<const float4*>(fo.res + off[u])) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int r = r0 + u * RSTEP;
            if (r >= nvalid) break;
            // same expressions as bn_apply_kernel: the folded result is bit-identical to conv + apply
            float4 y;
            y.x = fmaf(v[u].x - fmu.x, fsc.x, fbe.x), y.y = fmaf(v[u].y - fmu.y, fsc.y, fbe.y);
            y.z = fmaf(v[u].z - fmu.z, fsc.z, fbe.z), y.w = fmaf(v[u].w - fmu.w, fsc.w, fbe.w);
            if (fo.res) y.x += rr[u].x, y.y += rr[u].y, y.z += rr[u].z, y.w += rr[u].w;
            y.x = y.x > 0.f ? y.x : y.x * fo.slope, y.y = y.y > 0.f ? y.y : y.y * fo.slope;
            y.z = y.z > 0.f ? y.z : y.z * fo.slope, y.w = y.w > 0.f ? y.w : y.w * fo.slope;
            if (fo.write_f32) *reinterpret_cast<float4*>(p.C + off[u]) = y;
            if (fo.out_p) store_pair4(fo.out_p, fo.out_ps, off[u], y, fo.flags);
            if (fo.up_p) {
              const int bs = r / p.Lout, l = r - bs * p.Lout;
              const int64_t ou = ((int64_t)(b0 + bs) * (2 * p.Lout + 2) + 1 + 2 * l) * p.N + n0 + quad * 4;
              store_pair4(fo.up_p, fo.up_ps, ou, y);
              store_pair4(fo.up_p, fo.up_ps, ou + p.N, y);
            }
          }
        }
      } else {
        for (int r = t / QUADS; r < nvalid; r += RSTEP) {
          const int bs = r / p.Lout, l = r - bs * p.Lout;
          float4 v = *reinterpret_cast<const float4*>(&stage[S::epi(r, quad * 4)]);
          const int64_t off = ((int64_t)(b0 + bs) * p.out_rows + p.out_off + (int64_t)l * p.out_lstride) * p.N + n0 + quad * 4;
          float4* dst = reinterpret_cast<float4*>(p.C + off);
          if (B_MN && p.accumulate) {
            const float4 o = *dst;
            v.x += o.x, v.y += o.y, v.z += o.z, v.w += o.w;
          }
          *dst = v;
        }
      }
    }
    if (t == 0) stamp(p, 5);
    if (MODE == kConvTrain && (p.part || p.tot)) {  // BatchNorm statistics of this tile: (sum, centred sum of squares) per column
      // BN == 64: two threads per column (rows split in halves, combined with Chan's formula through shared memory)
      constexpr int TPC = 128 / BN;  // threads per column
      const int col = t % BN, part_id = t / BN;
      const int r_lo = (nvalid * part_id) / TPC, r_hi = (nvalid * (part_id + 1)) / TPC;
      const int cnt = r_hi - r_lo;
      // the thread's (at most 64) rows of its column sit in registers between the two passes: every shared-memory load
      // is issued before the first use
      constexpr int RMAX = TC_BM / TPC;
      const int cbase = (col >> 5) * S::EPI_BOX_FLOATS + (col & 3), c0 = (col & 31) >> 2;
      float v[RMAX];
#pragma unroll
      for (int j = 0; j < RMAX; ++j) {
        const int r = r_lo + j;
        v[j] = j < cnt ? stage[cbase + r * 32 + ((c0 ^ (r & 7)) << 2)] : 0.f;
      }
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int j = 0; j < RMAX; j += 4) s0 += v[j], s1 += v[j + 1], s2 += v[j + 2], s3 += v[j + 3];
      const float s = (s0 + s1) + (s2 + s3);
      const float mean = cnt > 0 ? s / (float)cnt : 0.f;
      float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
      for (int j = 0; j < RMAX; j += 4) {
        const float d0 = j + 0 < cnt ? v[j + 0] - mean : 0.f, d1 = j + 1 < cnt ? v[j + 1] - mean : 0.f;
        const float d2 = j + 2 < cnt ? v[j + 2] - mean : 0.f, d3 = j + 3 < cnt ? v[j + 3] - mean : 0.f;
        q0 = fmaf(d0, d0, q0), q1 = fmaf(d1, d1, q1), q2 = fmaf(d2, d2, q2), q3 = fmaf(d3, d3, q3);
      }
      float m2 = (q0 + q1) + (q2 + q3);
      float sum = s;
      if (TPC == 2) {
        asm volatile("bar.sync 1, 128;" ::: "memory");  // every thread is done reading its rows' staging columns
        float* xch = stage + (BN / 32) * S::EPI_BOX_FLOATS;  // [64][2] exchange, past the staging a TMA store may still read
        if (part_id == 1) xch[col * 2] = s, xch[col * 2 + 1] = m2;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (part_id == 0) {
          const float sb = xch[col * 2], m2b = xch[col * 2 + 1];
          const int cb = nvalid - cnt;
          if (cb > 0) {
            const float mb = sb / (float)cb;
            const float delta = mb - mean;
            m2 = m2 + m2b + delta * delta * ((float)cnt * (float)cb / (float)nvalid);
            sum = s + sb;
          }
        }
      }
      if (part_id == 0) {
        if (p.tot) {  // per-channel totals (sum x, sum x^2): exact enough in double, no per-tile partials to re-read
          const double sd = (double)sum;
          double* tot = p.tot + (int64_t)(blockIdx.x % kBnFwdTotCopies) * 2 * p.N;  // spread the collisions
          atomicAdd(tot + n0 + col, sd);
          atomicAdd(tot + p.N + n0 + col, (double)m2 + sd * sd / (double)nvalid);
        } else {
          *reinterpret_cast<float2*>(p.part + ((int64_t)blockIdx.x * p.N + n0 + col) * 2) = make_float2(sum, m2);
        }
      }
    }
    if (tma_out && t == 0) tma_store_wait_read();  // the staging must outlive the bulk stores' reads
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) stamp(p, 6);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(S::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient:  dW[m, n] += scale * sum_r dY[r, m] * X[(r + roff) * Cin + n]   (split over row ranges, atomics)
// Both operands are MN-major: the reduction index r is the row index of the channels-last pair tensors.  A k-block is
// 64 rows; a tile of 128 (BN) channels arrives as 2 (BN / 64) boxes of [64 rows][64 channels] (128-byte swizzle).
// ------------------------------------------------------------------------------------------------
struct PairWgrad {
  float* dW;
  int M, N;
  int R;
  int rows_per_split;  // multiple of PK
  float out_scale;
  const float* dyn_scale;
  uint32_t idesc, idesc2;  // N = BN, N = 2 * BN (see PairConv)
  int nmma, fused, fake_k;
  int pdl_late;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, reg_ctas_per_sm(BN, STAGES))
    wgrad_pair_kernel(const __grid_constant__ CUtensorMap mapDY, const __grid_constant__ CUtensorMap mapX,
                      const __grid_constant__ CUtensorMap mapDW, PairWgrad p) {
  using S = PairSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + S::RING_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* accum = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  if (!p.pdl_late) pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * TC_BM;
  const int r_begin = blockIdx.z * p.rows_per_split;
  const int r_end = min(p.R, r_begin + p.rows_per_split);
  const int nkb_all = (r_end - r_begin + PK - 1) / PK;  // >= 1 by construction of the grid
  const int nkb = p.fake_k ? min(p.fake_k, nkb_all) : nkb_all;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapDY);
    tma_prefetch_desc(&mapX);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], kWgradProducers);
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
  pdl_wait();  // everything above touched only shared / tensor memory

  // TMA producers: a k-block is 4 + 2 * (BN / 64) boxes of [64 rows][64 channels]; five issuing threads, one per warp
  // (see conv_pair_kernel): 0..3 = dY (group, plane), 4 = X (both planes)
  if (warp != 1 && lane == 0) {
    const int pid = warp == 0 ? 0 : warp - 1;
    const uint32_t tx_bytes = pid < 4 ? 8192u : (uint32_t)(2 * S::B_PLANE);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
      mbar_wait(&empty[s], ph ^ 1u);
      uint8_t* st = ring + s * S::STAGE_BYTES;
      mbar_expect_tx(&full[s], tx_bytes);
      const int r0 = r_begin + kb * PK;
      if (pid < 4) {
        const int g = pid >> 1, pl = pid & 1;
        tma_load_3d(st + pl * S::A_LO + g * 8192, &mapDY, &full[s], m0 + g * 64, r0, pl);
      } else {
#pragma unroll
        for (int g = 0; g < BN / 64; ++g) {
          tma_load_3d(st + S::B_OFF + g * 8192, &mapX, &full[s], n0 + g * 64, r0, 0);
          tma_load_3d(st + S::B_LO + g * 8192, &mapX, &full[s], n0 + g * 64, r0, 1);
        }
      }
    }
  }
  __syncwarp();

  if (warp == 0) {
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop, elected lane (see conv_pair_kernel)
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
      mbar_wait(&full[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t st = smem_u32(ring + s * S::STAGE_BYTES);
        const uint64_t a_hi = umma_desc(st, 8192, 1024, 2), a_lo = umma_desc(st + S::A_LO, 8192, 1024, 2);
        const uint64_t b_hi = umma_desc(st + S::B_OFF, 8192, 1024, 2);  // N = 2 BN: [X hi | X lo], groups 8 KB apart
        constexpr uint64_t blo = (uint64_t)(S::B_PLANE >> 4);
#pragma unroll
        for (int k16 = 0; k16 < PK / 16; ++k16) {
          const uint64_t adv = (uint64_t)(k16 * 2048 >> 4);  // 16 reduction rows = 16 x 128 B
          const int step = kb * (PK / 16) + k16;
          const uint32_t acc = tmem_base + (uint32_t)((step & 1) * 2 * BN);
          if (p.fused) {
            umma_f16(acc, a_hi + adv, b_hi + adv, p.idesc2, step >= 2 ? 1u : 0u);  // [hi*hi | hi*lo]
          } else {
            umma_f16(acc, a_hi + adv, b_hi + adv, p.idesc, step >= 2 ? 1u : 0u);
            umma_f16(acc + BN, a_hi + adv, b_hi + blo + adv, p.idesc, step >= 2 ? 1u : 0u);
          }
          if (p.nmma == 3) umma_f16(acc + BN, a_lo + adv, b_hi + adv, p.idesc, 1u);  // + lo*hi
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(accum);
    __syncwarp();
    if (p.pdl_late) pdl_trigger();
  } else {
    const int t = threadIdx.x - 64;
    mbar_wait(accum, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* stage = reinterpret_cast<float*>(ring);
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const float sc = p.dyn_scale ? p.out_scale * __ldg(p.dyn_scale) : p.out_scale;
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      // two live register arrays: the register footprint of these CTAs decides whether an elementwise kernel's CTA still
      // fits on an SM that holds two of them (122 -> 154 registers cost 4 % of the bs512 step, gpurun_out/r02_exp32.txt)
      uint32_t r[32], r1[32];
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
      tmem_ld32(ta + BN, r);       // cross terms, even k-steps
      tmem_ld32(ta + 3 * BN, r1);  // cross terms, odd k-steps
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r1[i]));
      tmem_ld32(ta, r1);           // hi*hi, even k-steps
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r1[i]));
      tmem_ld32(ta + 2 * BN, r1);  // hi*hi, odd k-steps
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = __float_as_uint((__uint_as_float(r[i]) + __uint_as_float(r1[i])) * sc);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<uint4*>(&stage[S::epi(row, c * 32 + i * 4)]) =
            make_uint4(r[i * 4], r[i * 4 + 1], r[i * 4 + 2], r[i * 4 + 3]);
    }
    // the split-K partial tile is added into dW by the TMA unit (fp32 add-reduce; rows past Cout are clipped by the map)
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (t == 0) {
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) tma_reduce_add_3d(&mapDW, stage + c * S::EPI_BOX_FLOATS, n0 + c * 32, m0, 0);
      tma_store_commit();
      tma_store_wait_read();
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(S::TMEM_COLS) : "memory");
  }
}

// fp32 -> pair planes (tools/pair_test and the weight refresh)
template <int FMT>
__global__ void to_pair_kernel(const float* __restrict__ src, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                               int64_t n, float scale, unsigned* __restrict__ flags) {
  bool sat = false;
  const int64_t n4 = n >> 2;  // four elements per thread and iteration: one 128-bit load, two 64-bit stores
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(src + q * 4);
    const float x[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
    uint16_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      pair_split<FMT>(x[e], h[e], l[e]);
      if (FMT == kPairF16) sat |= !(fabsf(x[e]) < kPairF16Max);
    }
    *reinterpret_cast<uint2*>(hi + q * 4) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
    *reinterpret_cast<uint2*>(lo + q * 4) = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = n4 * 4; i < n; ++i) {
      const float x = src[i] * scale;
      pair_split<FMT>(x, hi[i], lo[i]);
      if (FMT == kPairF16) sat |= !(fabsf(x) < kPairF16Max);
    }
  if (flags && sat) atomicOr(flags, kFlagWeightSaturated);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_enc = nullptr;
constexpr int kMapF32 = 100;  // encode(): fp32 output tensors (the pair formats are 0 / 1)

bool encode(TcMap* out, int fmt, int rank, const void* base, const cuuint64_t* dims, const cuuint64_t* strides,
            const cuuint32_t* box, CUtensorMapL2promotion promo) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapDataType dt = fmt == kMapF32    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : fmt == kPairBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                    : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = g_enc(reinterpret_cast<CUtensorMap*>(out->opaque), dt, (cuuint32_t)rank,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "hippie_b200: cuTensorMapEncodeTiled -> %d (rank %d, base %p, dims", (int)r, rank, base);
    for (int i = 0; i < rank; ++i) fprintf(stderr, " %llu", (unsigned long long)dims[i]);
    fprintf(stderr, ", strides");
    for (int i = 0; i + 1 < rank; ++i) fprintf(stderr, " %llu", (unsigned long long)strides[i]);
    fprintf(stderr, ", box");
    for (int i = 0; i < rank; ++i) fprintf(stderr, " %u", box[i]);
    fprintf(stderr, ")\n");
  }
  return r == CUDA_SUCCESS;
}

}  // namespace

template <int STAGES>
static void set_smem_attrs() {
  const int bytes = PairSmem<kBN, STAGES>::TOTAL;
  cudaFuncSetAttribute(conv_pair_kernel<kBN, STAGES, kConvTrain>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(conv_pair_kernel<kBN, STAGES, kConvEval>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(conv_pair_kernel<kBN, STAGES, kConvDgrad>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(wgrad_pair_kernel<kBN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
static int g_variant = 0, g_wgrad_variant = 1, g_pdl_late = 0;
static int g_mma_scheme = 7;  // HIPPIE_B200_MMA_SCHEME bits: 1 forward, 2 dgrad, 4 wgrad issue A hi x [B hi | B lo] as one instruction
static int g_fake_k = 0;
static bool g_fake_reuse = false;
static int g_bwd_nmma = 3;  // HIPPIE_B200_BWD_MMA=2: accuracy experiment, dgrad / wgrad without the (gradient lo) x (hi) product
static int g_wgrad_min_kb = 16;  // k-blocks (of 64 reduction rows) a weight-gradient CTA processes at least: the kernels are off
// the critical path, so few long-lived CTAs (less prologue / fill / epilogue time per MAC) beat many short ones -- the step is
// bound by the time CTAs occupy SM slots (bs512: 4 -> 3.134 ms, 16 -> 3.054 ms, 64 -> 3.308 ms per step)
static int g_sm_count = 148, g_alone_max = 74;

bool pair_init(std::string* err) {
  if (!g_enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      if (err) *err = "cuTensorMapEncodeTiled is not available from the driver";
      return false;
    }
    g_enc = reinterpret_cast<EncodeTiledFn>(fn);
  }
  // function attributes belong to the current device's context: set them on every bind
  set_smem_attrs<kStages>();
  set_smem_attrs<kStagesAlone>();
  if (const char* v = getenv("HIPPIE_B200_PAIR_VARIANT")) g_variant = atoi(v);  // 0 auto, 1 always shared, 2 always alone
  if (const char* v = getenv("HIPPIE_B200_PDL_LATE")) g_pdl_late = atoi(v);
  if (const char* v = getenv("HIPPIE_B200_MMA_SCHEME")) g_mma_scheme = atoi(v);
#ifdef HP_EXPERIMENTS  // result-changing speed experiments of DESIGN 4.3: only in builds made with EXTRA=-DHP_EXPERIMENTS
  if (const char* v = getenv("HIPPIE_B200_FAKE_K")) g_fake_k = atoi(v);
  if (const char* v = getenv("HIPPIE_B200_BWD_MMA")) g_bwd_nmma = atoi(v) == 2 ? 2 : 3;
  g_fake_reuse = getenv("HIPPIE_B200_FAKE_REUSE") != nullptr;
#endif
  if (const char* v = getenv("HIPPIE_B200_WGRAD_VARIANT")) g_wgrad_variant = atoi(v);  // 1 shared (two CTAs per SM), 2 alone
  if (const char* v = getenv("HIPPIE_B200_WGRAD_MIN_KB")) g_wgrad_min_kb = atoi(v) > 0 ? atoi(v) : 16;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  // "alone" variant for grids of at most half the SMs: two such kernels (the two branches of the model) still run side
  // by side; larger grids keep two CTAs per SM so that the chains and the weight gradients share every SM (bs512 step:
  // 3.11 ms vs 3.24 ms with every grid <= 148 CTAs alone; bs64 step 1.70 -> 1.52 ms, gpurun_out/r02_variants_step.txt)
  g_alone_max = g_sm_count / 2;
  if (const char* v = getenv("HIPPIE_B200_ALONE_MAX")) g_alone_max = atoi(v);
  if (cudaGetLastError() != cudaSuccess) {
    if (err) *err = "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed";
    return false;
  }
  return true;
}

// (k within the k*Cin window, logical output row, sample, plane)
bool pair_make_act_map(TcMap* out, const void* planes, int64_t plane_stride, int fmt, int in_C, int K, int Lout,
                       int in_rows, int in_stride, int in_off, int max_batch) {
  const int nb = TC_BM / Lout;
  cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)Lout, (cuuint64_t)max_batch, 2};
  cuuint64_t strides[3] = {(cuuint64_t)in_stride * in_C * 2, (cuuint64_t)in_rows * in_C * 2, (cuuint64_t)plane_stride * 2};
  cuuint32_t box[4] = {(cuuint32_t)PK, (cuuint32_t)Lout, (cuuint32_t)nb, 1};
  const uint16_t* base = static_cast<const uint16_t*>(planes) + (int64_t)in_off * in_C;
  return encode(out, fmt, 4, base, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
}

// K-major weight planes [N][K]: (k, n, plane), box 64 x bn x 1
bool pair_make_w_map(TcMap* out, const void* planes, int64_t plane_stride, int fmt, int N, int K, int bn) {
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, 2};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)plane_stride * 2};
  cuuint32_t box[3] = {(cuuint32_t)PK, (cuuint32_t)bn, 1};
  return encode(out, fmt, 3, planes, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

// the same planes [co][t][ci] read MN-major for dgrad: (ci, t, co, plane), box 64 x 1 x 64 x 1
bool pair_make_wmn_map(TcMap* out, const void* planes, int64_t plane_stride, int fmt, int Cout, int Cin, int k) {
  cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)k, (cuuint64_t)Cout, 2};
  cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)k * Cin * 2, (cuuint64_t)plane_stride * 2};
  cuuint32_t box[4] = {64, 1, 64, 1};
  return encode(out, fmt, 4, planes, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

// wgrad operands: (channel, reduction row, plane), box 64 x 64 x 1; row pitch `row_elems` (rows may overlap)
bool pair_make_rows_map(TcMap* out, const void* planes, int64_t plane_stride, int fmt, int64_t row_elems, int channels,
                        int rows) {
  cuuint64_t dims[3] = {(cuuint64_t)channels, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)row_elems * 2, (cuuint64_t)plane_stride * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)PK, 1};
  return encode(out, fmt, 3, planes, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
}

// fp32 conv output [B][out_rows][N], logical rows at out_off..: (channel, row, sample), box 32 x Lout x nb -- one
// 128-byte-swizzled staging box of the epilogue per 32 channels
bool pair_make_out_map(TcMap* out, float* C, int N, int Lout, int out_rows, int out_off, int batch) {
  const int nb = TC_BM / Lout;
  cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)Lout, (cuuint64_t)batch};  // rows past the batch are clipped
  cuuint64_t strides[2] = {(cuuint64_t)N * 4, (cuuint64_t)out_rows * N * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)Lout, (cuuint32_t)nb};
  return encode(out, kMapF32, 3, C + (int64_t)out_off * N, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_NONE);
}

// fp32 weight gradient [M][N]: (n, m, 1), box 32 x 128 x 1
bool pair_make_dw_map(TcMap* out, float* dW, int M, int N) {
  cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, 1};
  cuuint64_t strides[2] = {(cuuint64_t)N * 4, (cuuint64_t)M * N * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)TC_BM, 1};
  return encode(out, kMapF32, 3, dW, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_NONE);
}

int pair_pick_bn(int, int, int, int) { return kBN; }

template <int STAGES>
static void launch_conv_variant(const PairOpts& o, dim3 grid, cudaStream_t s, const CUtensorMap& a, const CUtensorMap& w,
                                const CUtensorMap& c, const PairConv& p) {
  constexpr int smem = PairSmem<kBN, STAGES>::TOTAL;
  if (o.b_mn)
    launch_pdl(conv_pair_kernel<kBN, STAGES, kConvDgrad>, grid, dim3(TC_THREADS), smem, s, a, w, c, p);
  else if (o.fold)
    launch_pdl(conv_pair_kernel<kBN, STAGES, kConvEval>, grid, dim3(TC_THREADS), smem, s, a, w, c, p);
  else
    launch_pdl(conv_pair_kernel<kBN, STAGES, kConvTrain>, grid, dim3(TC_THREADS), smem, s, a, w, c, p);
}

int launch_conv_pair(const ConvGemm& g, const TcMap& mapA, const TcMap& mapB, int bn, int B, const PairOpts& o,
                     cudaStream_t s) {
  PairConv p{};
  p.C = g.C, p.bias = g.bias, p.part = g.part, p.tot = g.tot, p.B = B, p.N = g.N, p.K = g.K, p.Lout = g.Lout;
  p.nb = TC_BM / g.Lout;
  p.out_rows = g.out_rows, p.out_off = g.out_off, p.out_lstride = g.out_lstride, p.accumulate = g.accumulate;
  p.out_scale = o.out_scale;
  p.idesc = umma_idesc_16(bn, o.a_fmt, o.b_fmt, 0, o.b_mn ? 1 : 0);
  p.idesc2 = umma_idesc_16(2 * bn, o.a_fmt, o.b_fmt, 0, o.b_mn ? 1 : 0);
  p.nmma = o.b_mn ? g_bwd_nmma : 3;
  p.fused = (g_mma_scheme >> (o.b_mn ? 1 : 0)) & 1;
  p.fake_reuse = (g_fake_reuse && o.taps == 3) ? 1 : 0;
  p.fake_k = g_fake_k;
  p.b_mn = o.b_mn, p.taps = o.taps, p.kb_per_tap = o.taps > 0 ? (g.K / o.taps) / PK : 1;
  p.dyn_scale = o.dyn_scale, p.stamps = o.stamps, p.stamps_all = o.stamps_all, p.pdl_late = g_pdl_late & 1;
  if (o.fold) p.fold = *o.fold;
  dim3 grid((B + p.nb - 1) / p.nb, g.N / bn);

  const CUtensorMap& a = *reinterpret_cast<const CUtensorMap*>(mapA.opaque);
  const CUtensorMap& w = *reinterpret_cast<const CUtensorMap*>(mapB.opaque);
  p.tma_out = (o.out_map && g.out_lstride == 1) ? 1 : 0;
  const CUtensorMap& c = p.tma_out ? *reinterpret_cast<const CUtensorMap*>(o.out_map->opaque) : a;
  const bool alone = g_variant == 2 || (g_variant == 0 && (int)(grid.x * grid.y) <= g_alone_max);
  if (alone)
    launch_conv_variant<kStagesAlone>(o, grid, s, a, w, c, p);
  else
    launch_conv_variant<kStages>(o, grid, s, a, w, c, p);
  return p.nb * g.Lout;
}

void launch_wgrad_pair(const WgradGemm& g, const TcMap& mapDY, const TcMap& mapX, const TcMap& mapDW, int bn, int sm_count,
                       const PairOpts& o, cudaStream_t s) {
  PairWgrad p{};
  p.dW = g.dW, p.M = g.M, p.N = g.N, p.R = g.R;
  p.out_scale = o.out_scale, p.dyn_scale = o.dyn_scale, p.pdl_late = (g_pdl_late >> 1) & 1;
  bn = kBN;
  p.idesc = umma_idesc_16(bn, o.a_fmt, o.b_fmt, 1, 1);
  p.idesc2 = umma_idesc_16(2 * bn, o.a_fmt, o.b_fmt, 1, 1);
  p.nmma = g_bwd_nmma;
  p.fused = (g_mma_scheme >> 2) & 1;
  p.fake_k = g_fake_k;
  const int tiles = ((g.M + TC_BM - 1) / TC_BM) * (g.N / bn);
  const bool alone = g_wgrad_variant == 2;
  int splits = ((alone ? 1 : 2) * sm_count) / tiles;  // one or two CTAs per SM
  const int kblocks = (g.R + PK - 1) / PK;
  if (splits > (kblocks + g_wgrad_min_kb - 1) / g_wgrad_min_kb) splits = (kblocks + g_wgrad_min_kb - 1) / g_wgrad_min_kb;
  if (splits < 1) splits = 1;
  const int kb_per = (kblocks + splits - 1) / splits;
  p.rows_per_split = kb_per * PK;
  splits = (g.R + p.rows_per_split - 1) / p.rows_per_split;
  dim3 grid(g.N / bn, (g.M + TC_BM - 1) / TC_BM, splits);
  const CUtensorMap& a = *reinterpret_cast<const CUtensorMap*>(mapDY.opaque);
  const CUtensorMap& x = *reinterpret_cast<const CUtensorMap*>(mapX.opaque);
  const CUtensorMap& dw = *reinterpret_cast<const CUtensorMap*>(mapDW.opaque);
  if (alone)
    launch_pdl(wgrad_pair_kernel<kBN, kStagesAlone>, grid, dim3(TC_THREADS), PairSmem<kBN, kStagesAlone>::TOTAL, s, a, x, dw, p);
  else
    launch_pdl(wgrad_pair_kernel<kBN, kStages>, grid, dim3(TC_THREADS), PairSmem<kBN, kStages>::TOTAL, s, a, x, dw, p);
}

void launch_to_pair(const float* src, void* planes, int64_t plane_stride, int64_t n, float scale, int fmt,
                    cudaStream_t s, unsigned* flags) {
  uint16_t* hi = static_cast<uint16_t*>(planes);
  int blocks = (int)((n / 4 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  if (fmt == kPairBF16)
    to_pair_kernel<kPairBF16><<<blocks, 256, 0, s>>>(src, hi, hi + plane_stride, n, scale, flags);
  else
    to_pair_kernel<kPairF16><<<blocks, 256, 0, s>>>(src, hi, hi + plane_stride, n, scale, flags);
}

}  // namespace hp
