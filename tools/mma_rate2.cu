// Issue rate of tcgen05.mma.kind::f16 (M = 128, K = 16, SS, 128-byte swizzle), second experiment: HOW the MMAs are issued.
//   issue 0: `if (lane == 0)` branch (the compiler wraps every UTCHMMA in an elect / BRA.U.ANY waterfall loop)
//   issue 1: warp-uniform loop, elect.sync leader (straight-line UTCHMMA)
//   issue 2: as 1, two issuing warps (even / odd k-blocks, separate accumulators)
// Varied also: N, CTAs per SM (1 / 2).   make -C tools mma_rate2 && ./tools/mma_rate2   (on a B200)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "../hippie_b200/csrc/tc_common.cuh"

using namespace hp::tc;

struct Args {
  int n, iters;
  long long* out;
};

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// 12 MMAs per k-block: 4 k-steps x {lo*hi, hi*lo -> accumulator X, hi*hi -> accumulator M}
template <int NMMA>
__device__ __forceinline__ void issue_kblock(uint32_t st, int n, uint32_t accM, uint32_t accX, uint32_t idesc, bool first) {
  const uint64_t a_hi = umma_desc(st, 16, 1024, 2), a_lo = umma_desc(st + 16384, 16, 1024, 2);
  const uint64_t b_hi = umma_desc(st + 32768, 16, 1024, 2), b_lo = umma_desc(st + 32768 + n * 128, 16, 1024, 2);
#pragma unroll
  for (int k16 = 0; k16 < 4; ++k16) {
    const uint64_t adv = (uint64_t)(k16 * 32 >> 4);
    const uint32_t acc = (first && k16 == 0) ? 0u : 1u;
    if (NMMA == 3) {
      umma_f16(accX, a_lo + adv, b_hi + adv, idesc, acc);
      umma_f16(accX, a_hi + adv, b_lo + adv, idesc, 1u);
    } else if (NMMA == 2) {
      umma_f16(accX, a_hi + adv, b_lo + adv, idesc, acc);
    }
    umma_f16(accM, a_hi + adv, b_hi + adv, idesc, acc);
  }
}

template <int ISSUE, int NMMA, int MINB>
__global__ void __launch_bounds__(128, MINB) mma_rate_kernel(Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = 2 * 16384 + 2 * a.n * 128;
  for (int i = threadIdx.x; i < 2 * stage_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ring)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr int COLS = MINB == 1 ? 512 : 256;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = umma_idesc_16(a.n, 0, 0, 0, 0);
  // accumulators: M at column 0, X at column n (two issuers: second pair at 2n, 3n when it fits, else shared)
  if (ISSUE == 0) {
    if (warp == 1 && lane == 0) {
      const long long t0 = clock64();
      for (int kb = 0; kb < a.iters; ++kb)
        issue_kblock<NMMA>(smem_u32(ring + (kb & 1) * stage_bytes), a.n, tmem, tmem + a.n, idesc, kb == 0);
      umma_commit(&bar[0]);
      mbar_wait(&bar[0], 0);
      if (blockIdx.x == 0) a.out[0] = clock64() - t0;
    }
  } else if (ISSUE == 1) {
    if (warp == 1) {
      const long long t0 = clock64();
      for (int kb = 0; kb < a.iters; ++kb) {
        if (elect_one()) issue_kblock<NMMA>(smem_u32(ring + (kb & 1) * stage_bytes), a.n, tmem, tmem + a.n, idesc, kb == 0);
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bar[0]);
      __syncwarp();
      mbar_wait(&bar[0], 0);
      if (blockIdx.x == 0 && lane == 0) a.out[0] = clock64() - t0;
    }
  } else {
    if (warp == 1 || warp == 2) {
      const int w = warp - 1;
      const uint32_t base = (4 * a.n <= COLS) ? tmem + w * 2 * a.n : tmem;
      const long long t0 = clock64();
      for (int kb = w; kb < a.iters; kb += 2) {
        if (elect_one()) issue_kblock<NMMA>(smem_u32(ring + (kb & 1) * stage_bytes), a.n, base, base + a.n, idesc, kb == w);
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bar[w]);
      __syncwarp();
      mbar_wait(&bar[w], 0);
      if (blockIdx.x == 0 && lane == 0) a.out[w] = clock64() - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(COLS) : "memory");
}

template <int ISSUE, int NMMA, int MINB>
static void run(int n, int ctas, long long* d) {
  if (2 * n > (MINB == 1 ? 512 : 256)) return;
  Args a{n, 200, d};
  const size_t smem = 2 * (2 * 16384 + 2 * n * 128) + 2048;
  cudaFuncSetAttribute(mma_rate_kernel<ISSUE, NMMA, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaMemset(d, 0, 64);
  mma_rate_kernel<ISSUE, NMMA, MINB><<<ctas, 128, smem>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("N=%d issue %d: %s\n", n, ISSUE, cudaGetErrorString(e));
    exit(1);
  }
  long long cyc[2];
  cudaMemcpy(cyc, d, 16, cudaMemcpyDeviceToHost);
  const long long c = cyc[0] > cyc[1] ? cyc[0] : cyc[1];
  const int mmas = 200 * 4 * NMMA;
  const double per_sm = (double)c / mmas / (ctas > 148 ? 2 : 1);
  printf("%4d %6d %5d %7d %6d | %10lld %12.1f %14.1f %10.1f\n", n, ISSUE, NMMA, MINB, ctas, c, (double)c / mmas, per_sm,
         128.0 * n / 256.0);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  printf("%4s %6s %5s %7s %6s | %10s %12s %14s %10s\n", "N", "issue", "nmma", "cta/sm", "ctas", "cycles", "cyc/MMA/CTA",
         "cyc/MMA/SM", "floor");
  for (int n : {64, 128, 256}) {
    run<0, 3, 1>(n, 148, d);
    run<1, 3, 1>(n, 148, d);
    run<2, 3, 1>(n, 148, d);
    run<0, 2, 1>(n, 148, d);
    run<1, 2, 1>(n, 148, d);
    run<1, 1, 1>(n, 148, d);
    run<2, 1, 1>(n, 148, d);
    run<0, 3, 2>(n, 296, d);
    run<1, 3, 2>(n, 296, d);
    run<2, 3, 2>(n, 296, d);
  }
  return 0;
}
