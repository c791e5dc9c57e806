// How fast can one CTA / one SM ingest operand tiles?  Every CTA runs a TMA producer / consumer ring with NO tensor-core
// work: a stage is 48 KB (the pair GEMM's k-block) fetched as `boxes` boxes, the consumer only waits for the stage and
// hands it back.  Varied: box shape (rows per box, two planes per box), row pitch in global memory, number of
// producer threads (different warps), tensor-map vs plain 1-D bulk copies, CTAs per SM.  Prints the aggregate ingest
// rate -- the ceiling every pair-plane GEMM main loop runs against.
//   make -C tools tma_bw && ./tools/tma_bw          (on a B200)
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../hippie_b200/csrc/tc_common.cuh"

using namespace hp::tc;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

enum { kTensor2D = 0, kTensor3D = 1, kBulk1D = 2 };

struct Args {
  int stages, boxes, box_bytes, iters;
  int mode;       // kTensor2D: box {64, rows}; kTensor3D: box {64, rows / 2, 2 planes}; kBulk1D: contiguous chunks
  int rows;       // rows per box (kTensor3D: both planes together)
  int producers;  // producer threads, one per warp; box b of a stage is issued by producer b % producers
  int ntiles;     // distinct tiles in the working set
  int col_tiles;
  const uint8_t* base;
};

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(192, 2) ingest_kernel(const __grid_constant__ CUtensorMap map, Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = a.boxes * a.box_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + a.stages * stage_bytes);
  uint64_t* empty = full + a.stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) mbar_init(&full[s], a.producers), mbar_init(&empty[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned t = (unsigned)blockIdx.x * 7919u;
  if (warp < a.producers && lane == 0) {
    const int mine = (a.boxes - warp + a.producers - 1) / a.producers;  // boxes this producer issues per stage
    for (int i = 0; i < a.iters; ++i) {
      const int s = i % a.stages;
      mbar_wait(&empty[s], (((unsigned)(i / a.stages)) & 1u) ^ 1u);
      mbar_expect_tx(&full[s], (uint32_t)(mine * a.box_bytes));
      for (int b = warp; b < a.boxes; b += a.producers) {
        const unsigned tile = (t + (unsigned)(i * a.boxes + b)) % (unsigned)a.ntiles;
        uint8_t* dst = ring + s * stage_bytes + b * a.box_bytes;
        if (a.mode == kTensor2D)
          tma_load_2d(dst, &map, &full[s], (int)(tile % a.col_tiles) * 64, (int)(tile / a.col_tiles) * a.rows);
        else if (a.mode == kTensor3D)
          tma_load_3d(dst, &map, &full[s], (int)(tile % a.col_tiles) * 64, (int)(tile / a.col_tiles) * (a.rows / 2), 0);
        else
          bulk_load_1d(dst, a.base + (size_t)tile * a.box_bytes, (uint32_t)a.box_bytes, &full[s]);
      }
    }
  } else if (warp == 5 && lane == 0) {
    for (int i = 0; i < a.iters; ++i) {
      const int s = i % a.stages;
      mbar_wait(&full[s], ((unsigned)(i / a.stages)) & 1u);
      mbar_arrive(&empty[s]);
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  CK(cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const size_t set_bytes = (size_t)32 << 20;  // L2-resident working set
  void* buf;
  CK(cudaMalloc(&buf, set_bytes));
  CK(cudaMemset(buf, 0, set_bytes));
  printf("%8s %6s %6s %6s %5s %5s %6s | %9s %9s %9s %12s\n", "mode", "boxes", "rows", "pitch", "prod", "stg", "cta/sm", "us", "TB/s",
         "B/clk/SM", "cyc/box/CTA");
  struct Cfg {
    int mode, rows, col_tiles, producers, stages, per_sm, swz;
  };
  std::vector<Cfg> cfgs;
  for (int per_sm : {1, 2}) {
    for (int prod : {1, 2, 3})
      for (int rows : {64, 128, 256}) cfgs.push_back({kTensor2D, rows, 16, prod, 2, per_sm, 1});
    cfgs.push_back({kTensor2D, 128, 1, 1, 2, per_sm, 1});   // contiguous rows (pitch 128 B)
    cfgs.push_back({kTensor2D, 256, 1, 1, 2, per_sm, 1});
    cfgs.push_back({kTensor2D, 128, 4, 1, 2, per_sm, 1});   // pitch 512 B (256-channel tensors)
    cfgs.push_back({kTensor2D, 128, 16, 1, 2, per_sm, 0});  // no swizzle
    cfgs.push_back({kTensor3D, 256, 16, 1, 2, per_sm, 1});  // two planes of 128 rows in one box
    cfgs.push_back({kTensor3D, 128, 16, 1, 2, per_sm, 1});  // two planes of 64 rows in one box
    cfgs.push_back({kTensor3D, 256, 16, 2, 2, per_sm, 1});
    for (int prod : {1, 2, 3})
      for (int rows : {128, 256, 384}) cfgs.push_back({kBulk1D, rows, 1, prod, 2, per_sm, 1});
  }
  cfgs.push_back({kTensor2D, 128, 16, 1, 4, 1, 1});
  cfgs.push_back({kTensor2D, 128, 16, 3, 4, 1, 1});
  for (const Cfg& c : cfgs) {
    Args a{};
    a.mode = c.mode, a.rows = c.rows, a.producers = c.producers, a.stages = c.stages, a.iters = 200, a.col_tiles = c.col_tiles;
    a.box_bytes = c.rows * 128;
    a.boxes = 48 * 1024 / a.box_bytes;
    if (a.boxes < 1) continue;
    const int stage_bytes = a.boxes * a.box_bytes;
    a.base = static_cast<const uint8_t*>(buf);
    const size_t plane_bytes = set_bytes / 2;
    CUtensorMap map;
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r;
    const CUtensorMapSwizzle sw = c.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
    if (c.mode == kTensor3D) {
      const int prow = c.rows / 2;
      const int row_tiles = (int)(plane_bytes / ((size_t)c.col_tiles * 128 * prow));
      a.ntiles = row_tiles * c.col_tiles;
      cuuint64_t dims[3] = {(cuuint64_t)c.col_tiles * 64, (cuuint64_t)row_tiles * prow, 2};
      cuuint64_t strides[2] = {(cuuint64_t)c.col_tiles * 128, (cuuint64_t)plane_bytes};
      cuuint32_t box[3] = {64, (cuuint32_t)prow, 2};
      r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      const int row_tiles = (int)(set_bytes / ((size_t)c.col_tiles * 128 * c.rows));
      a.ntiles = c.mode == kBulk1D ? (int)(set_bytes / a.box_bytes) : row_tiles * c.col_tiles;
      cuuint64_t dims[2] = {(cuuint64_t)c.col_tiles * 64, (cuuint64_t)row_tiles * c.rows};
      cuuint64_t strides[1] = {(cuuint64_t)c.col_tiles * 128};
      cuuint32_t box[2] = {64, (cuuint32_t)c.rows};
      r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
      printf("encode failed %d (mode %d rows %d)\n", (int)r, c.mode, c.rows);
      continue;
    }
    const size_t smem = (size_t)c.stages * stage_bytes + 2048;
    if (smem > (size_t)(c.per_sm == 2 ? 110 : 200) * 1024) continue;
    const int grid = sms * c.per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    ingest_kernel<<<grid, 192, smem>>>(map, a);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int it = 0; it < 5; ++it) ingest_kernel<<<grid, 192, smem>>>(map, a);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    const double bytes = (double)grid * a.iters * stage_bytes;
    const char* mn = c.mode == kTensor2D ? (c.swz ? "2D" : "2D-nosw") : c.mode == kTensor3D ? "3D-2pl" : "bulk1D";
    printf("%8s %6d %6d %6d %5d %5d %6d | %9.1f %9.2f %9.1f %12.0f\n", mn, a.boxes, c.rows, c.col_tiles * 128, c.producers, c.stages,
           c.per_sm, ms * 1e3, bytes / (ms * 1e-3) / 1e12, bytes / (ms * 1e-3) / sms / 1.965e9,
           ms * 1e-3 * 1.965e9 / ((double)a.iters * a.boxes));
  }
  return 0;
}
