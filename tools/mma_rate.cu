// Issue rate of tcgen05.mma.kind::f16 (M = 128, K = 16, both operands in shared memory, 128-byte swizzle) as the pair
// GEMMs of conv_pair.cu use it: one thread issues `iters` k-blocks of 12 MMAs (4 k-steps x {lo*hi, hi*lo, hi*hi}) on
// operand tiles that are already in shared memory -- no TMA, no epilogue.  Varied: N, how the MMAs are spread over
// accumulators, whether consecutive MMAs re-use the A descriptor, CTAs per SM.
//   make -C tools mma_rate && ./tools/mma_rate        (on a B200)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "../hippie_b200/csrc/tc_common.cuh"

using namespace hp::tc;

struct Args {
  int n;         // MMA N
  int iters;     // k-blocks
  int pattern;   // 0: conv_pair pattern (3 accumulators), 1: all into one accumulator, 2: only the hi*hi MMA (4 per k-block),
                 // 3: 12 MMAs, hi*hi only operands (same descriptors re-used three times), 4: 6 accumulators round robin
  int b_mn;
  long long* out;
};

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // zero the operand tiles (2 stages x (A hi, A lo 16 KB each, B hi, B lo n * 128 B each))
  const int stage_bytes = 2 * 16384 + 2 * a.n * 128;
  for (int i = threadIdx.x; i < 2 * stage_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ring)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = umma_idesc_16(a.n, 0, 0, 0, a.b_mn);
    const long long t0 = clock64();
    int step = 0;
    for (int kb = 0; kb < a.iters; ++kb) {
      const uint32_t st = smem_u32(ring + (kb & 1) * stage_bytes);
      const uint64_t a_hi = umma_desc(st, 16, 1024, 2), a_lo = umma_desc(st + 16384, 16, 1024, 2);
      uint64_t b_hi, b_lo, badv;
      if (!a.b_mn) {
        b_hi = umma_desc(st + 32768, 16, 1024, 2), b_lo = umma_desc(st + 32768 + a.n * 128, 16, 1024, 2), badv = 32 >> 4;
      } else {
        b_hi = umma_desc(st + 32768, 8192, 1024, 2), b_lo = umma_desc(st + 32768 + a.n * 128, 8192, 1024, 2), badv = 2048 >> 4;
      }
#pragma unroll
      for (int k16 = 0; k16 < 4; ++k16, ++step) {
        const uint64_t aadv = (uint64_t)(k16 * 32 >> 4), bad = (uint64_t)k16 * badv;
        const int nn = a.n;
        if (a.pattern == 0) {
          umma_f16(tmem + 2 * nn, a_lo + aadv, b_hi + bad, idesc, step != 0);
          umma_f16(tmem + 2 * nn, a_hi + aadv, b_lo + bad, idesc, 1u);
          umma_f16(tmem + (step & 1) * nn, a_hi + aadv, b_hi + bad, idesc, step >= 2);
        } else if (a.pattern == 1) {
          umma_f16(tmem, a_lo + aadv, b_hi + bad, idesc, step != 0);
          umma_f16(tmem, a_hi + aadv, b_lo + bad, idesc, 1u);
          umma_f16(tmem, a_hi + aadv, b_hi + bad, idesc, 1u);
        } else if (a.pattern == 2) {
          umma_f16(tmem, a_hi + aadv, b_hi + bad, idesc, step != 0);
        } else if (a.pattern == 3) {
          umma_f16(tmem, a_hi + aadv, b_hi + bad, idesc, step != 0);
          umma_f16(tmem, a_hi + aadv, b_hi + bad, idesc, 1u);
          umma_f16(tmem, a_hi + aadv, b_hi + bad, idesc, 1u);
        } else {
          const int acc = (step * 3) % 6;
          if (nn * 6 <= 512) {
            umma_f16(tmem + ((acc + 0) % 6) * nn, a_lo + aadv, b_hi + bad, idesc, step >= 2);
            umma_f16(tmem + ((acc + 1) % 6) * nn, a_hi + aadv, b_lo + bad, idesc, step >= 2);
            umma_f16(tmem + ((acc + 2) % 6) * nn, a_hi + aadv, b_hi + bad, idesc, step >= 2);
          }
        }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) a.out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("%4s %8s %5s %6s | %10s %12s %14s\n", "N", "pattern", "b_mn", "ctas", "cycles", "cyc/MMA", "floor cyc/MMA");
  for (int ctas : {1, 148})
    for (int b_mn : {0, 1})
      for (int n : {64, 128, 256})
        for (int pattern : {0, 1, 2, 3, 4}) {
          if (n * 3 > 512 && pattern == 0) continue;
          if (pattern == 4 && n * 6 > 512) continue;
          if (b_mn && n > 128) continue;
          Args a{n, 200, pattern, b_mn, d};
          const size_t smem = 2 * (2 * 16384 + 2 * n * 128) + 2048;
          mma_rate_kernel<<<ctas, 128, smem>>>(a);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("N=%d pattern %d: %s\n", n, pattern, cudaGetErrorString(e));
            return 1;
          }
          long long cyc;
          cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
          const int mmas = 200 * (pattern == 2 ? 4 : 12);
          printf("%4d %8d %5d %6d | %10lld %12.1f %14.1f\n", n, pattern, b_mn, ctas, cyc, (double)cyc / mmas, 128.0 * n / 256.0);
        }
  return 0;
}
