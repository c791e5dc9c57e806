// What a grid-wide barrier costs on a B200 -- the price of one BatchNorm coupling inside a persistent layer loop (DESIGN.md
// section 8) next to the ~9 us a kernel boundary costs the bs512 step today.
//   make -C tools grid_barrier && ./tools/grid_barrier
// Variants: cooperative-groups grid.sync(); a hand-rolled barrier (one red.release.gpu per CTA on a counter, every CTA
// polls a generation word with ld.acquire.gpu); the same with a payload phase (every CTA streams a slice of a buffer
// between barriers, the shape of a BatchNorm apply chunk).
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

namespace cg = cooperative_groups;

__device__ __forceinline__ void barrier_arrive_wait(unsigned* count, unsigned* gen, unsigned nctas, unsigned& my_gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    my_gen += 1;
    unsigned prev;
    asm volatile("atom.add.acq_rel.gpu.u32 %0, [%1], 1;" : "=r"(prev) : "l"(count) : "memory");
    if (prev == nctas * my_gen - 1) {  // last arriver of this generation publishes it
      asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(gen), "r"(my_gen) : "memory");
    } else {
      unsigned g;
      do {
        asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(g) : "l"(gen) : "memory");
      } while (g < my_gen);
    }
  }
  __syncthreads();
}

__global__ void k_coop(int iters, long long* cycles) {
  cg::grid_group grid = cg::this_grid();
  grid.sync();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) grid.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = clock64() - t0;
}

__global__ void k_hand(int iters, unsigned* count, unsigned* gen, long long* cycles, float* buf, size_t per_cta) {
  unsigned my_gen = 0;
  barrier_arrive_wait(count, gen, gridDim.x, my_gen);
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (per_cta) {  // payload: read-modify-write a slice (L2-resident), like one BatchNorm apply chunk
      float4* p = reinterpret_cast<float4*>(buf + (size_t)blockIdx.x * per_cta);
      for (size_t j = threadIdx.x; j < per_cta / 4; j += blockDim.x) {
        float4 v = p[j];
        v.x += 1.f, v.y += 1.f, v.z += 1.f, v.w += 1.f;
        p[j] = v;
      }
      __threadfence();
    }
    barrier_arrive_wait(count, gen, gridDim.x, my_gen);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = clock64() - t0;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  long long* cycles;
  unsigned* sync_words;
  float* buf;
  const size_t buf_floats = (size_t)8 << 20;  // 32 MB
  cudaMalloc(&cycles, 8), cudaMalloc(&sync_words, 256), cudaMalloc(&buf, buf_floats * 4);
  cudaMemset(buf, 0, buf_floats * 4);
  const int iters = 2000;
  printf("%d SMs, %.0f MHz (nominal)\n", sms, khz / 1e3);
  for (int per_sm = 1; per_sm <= 2; ++per_sm) {
    const int nctas = sms * per_sm;
    for (int threads : {192, 256}) {
      long long c = 0;
      int it = iters;
      void* args[] = {&it, &cycles};
      cudaError_t e = cudaLaunchCooperativeKernel((void*)k_coop, dim3(nctas), dim3(threads), args, 0, 0);
      cudaDeviceSynchronize();
      cudaMemcpy(&c, cycles, 8, cudaMemcpyDeviceToHost);
      printf("grid.sync()          %3d CTAs x %3d threads: %7.0f cycles = %.2f us per barrier%s\n", nctas, threads,
             (double)c / iters, (double)c / iters / (khz / 1e3), e == cudaSuccess ? "" : "  (launch failed)");
      for (size_t per_cta : {(size_t)0, (size_t)8192, (size_t)32768}) {  // floats per CTA and phase: 0 / 32 KB / 128 KB
        if ((size_t)nctas * per_cta > buf_floats) continue;
        cudaMemset(sync_words, 0, 256);
        unsigned *count = sync_words, *gen = sync_words + 32;
        size_t pc = per_cta;
        void* a2[] = {&it, &count, &gen, &cycles, &buf, &pc};
        e = cudaLaunchCooperativeKernel((void*)k_hand, dim3(nctas), dim3(threads), a2, 0, 0);
        cudaDeviceSynchronize();
        cudaMemcpy(&c, cycles, 8, cudaMemcpyDeviceToHost);
        printf("atomic + acquire poll %3d CTAs x %3d threads, %3zu KB per CTA and phase: %7.0f cycles = %.2f us per phase%s\n",
               nctas, threads, per_cta * 4 / 1024, (double)c / iters, (double)c / iters / (khz / 1e3),
               e == cudaSuccess ? "" : "  (launch failed)");
      }
    }
  }
  return 0;
}
