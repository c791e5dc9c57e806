// FP32 CUDA-core implicit-GEMM kernels: conv1d forward / dgrad (conv_gemm_simt) and wgrad
// (wgrad_simt).  This is the bit-faithful fp32 path (conv_path = 1) and the parity yardstick
// for the tcgen05 path in conv_tc.cu.
//
// Replaces the nn.Conv1d calls of the reference (hippie/backbones.py:11,24,26,31,50,55,78)
// and their autograd backward.
#include "kernels.cuh"

namespace hp {

namespace {

constexpr int BK = 16;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// C[m, n] (+)= sum_k A[row(m), k] * W[n, k] (+ bias[n]); optional BatchNorm statistics partials.
// 256 threads as 16 (ty: rows) x 16 (tx: cols); each thread owns TM x TN outputs in 4-wide groups.
template <int BM, int BN>
__global__ void __launch_bounds__(256) conv_gemm_simt_kernel(ConvGemm p) {
  constexpr int TM = BM / 16, TN = BN / 16;  // 4 or 8
  constexpr int HM = TM / 4, HN = TN / 4;    // number of 4-wide groups
  constexpr int AP = BM + 4, BP = BN + 4;
  constexpr int LA = BM * 4 / 256, LB = BN * 4 / 256;  // float4 loads per thread per k-tile
  __shared__ __align__(16) float As[2][BK][AP];
  __shared__ __align__(16) float Bs[2][BK][BP];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  const float* aptr[LA];
  const float* bptr[LB];
#pragma unroll
  for (int i = 0; i < LA; ++i) {
    int id = tid + i * 256, row = id >> 2, kq = id & 3;
    int m = min(m0 + row, p.M - 1);
    int b = m / p.Lout, l = m - b * p.Lout;
    aptr[i] = p.A + ((int64_t)b * p.in_rows + (int64_t)l * p.in_stride + p.in_off) * p.in_C + kq * 4;
  }
#pragma unroll
  for (int i = 0; i < LB; ++i) {
    int id = tid + i * 256, row = id >> 2, kq = id & 3;
    bptr[i] = p.W + (int64_t)(n0 + row) * p.K + kq * 4;
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[LA], rb[LB];
  const int nk = p.K / BK;
#pragma unroll
  for (int i = 0; i < LA; ++i) ra[i] = ldg4(aptr[i]);
#pragma unroll
  for (int i = 0; i < LB; ++i) rb[i] = ldg4(bptr[i]);

  auto stage = [&](int buf) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int id = tid + i * 256, row = id >> 2, kq = (id & 3) * 4;
      As[buf][kq + 0][row] = ra[i].x;
      As[buf][kq + 1][row] = ra[i].y;
      As[buf][kq + 2][row] = ra[i].z;
      As[buf][kq + 3][row] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int id = tid + i * 256, row = id >> 2, kq = (id & 3) * 4;
      Bs[buf][kq + 0][row] = rb[i].x;
      Bs[buf][kq + 1][row] = rb[i].y;
      Bs[buf][kq + 2][row] = rb[i].z;
      Bs[buf][kq + 3][row] = rb[i].w;
    }
  };
  stage(0);
  __syncthreads();

  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
#pragma unroll
      for (int i = 0; i < LA; ++i) ra[i] = ldg4(aptr[i] + (kt + 1) * BK);
#pragma unroll
      for (int i = 0; i < LB; ++i) rb[i] = ldg4(bptr[i] + (kt + 1) * BK);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int h = 0; h < HM; ++h) {
        float4 v = *reinterpret_cast<const float4*>(&As[cur][kk][h * (BM / 2) + ty * 4]);
        a[h * 4 + 0] = v.x, a[h * 4 + 1] = v.y, a[h * 4 + 2] = v.z, a[h * 4 + 3] = v.w;
      }
#pragma unroll
      for (int h = 0; h < HN; ++h) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[cur][kk][h * (BN / 2) + tx * 4]);
        b[h * 4 + 0] = v.x, b[h * 4 + 1] = v.y, b[h * 4 + 2] = v.z, b[h * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) stage(cur ^ 1);
    __syncthreads();
  }

  // ---- epilogue: bias, (accumulate,) store ----
  if (p.bias) {
#pragma unroll
    for (int h = 0; h < HN; ++h) {
      float4 bv = ldg4(p.bias + n0 + h * (BN / 2) + tx * 4);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        acc[i][h * 4 + 0] += bv.x, acc[i][h * 4 + 1] += bv.y, acc[i][h * 4 + 2] += bv.z, acc[i][h * 4 + 3] += bv.w;
      }
    }
  }
  bool valid[TM];
#pragma unroll
  for (int hi = 0; hi < TM; ++hi) {
    int row = (hi >> 2) * (BM / 2) + ty * 4 + (hi & 3);
    int m = m0 + row;
    valid[hi] = m < p.M;
    if (!valid[hi]) continue;
    int b = m / p.Lout, l = m - b * p.Lout;
    float* crow = p.C + ((int64_t)b * p.out_rows + p.out_off + (int64_t)l * p.out_lstride) * p.N + n0;
#pragma unroll
    for (int h = 0; h < HN; ++h) {
      float4* dst = reinterpret_cast<float4*>(crow + h * (BN / 2) + tx * 4);
      float4 v = make_float4(acc[hi][h * 4 + 0], acc[hi][h * 4 + 1], acc[hi][h * 4 + 2], acc[hi][h * 4 + 3]);
      if (p.accumulate) {
        float4 o = *dst;
        v.x += o.x, v.y += o.y, v.z += o.z, v.w += o.w;
      }
      *dst = v;
    }
  }

  // ---- BatchNorm statistics of this tile: per column (sum, centred sum of squares) ----
  if (p.part) {
    float* red = &As[0][0][0];    // [16][BN]
    float* smean = &Bs[0][0][0];  // [BN]
    const int nvalid = min(BM, p.M - m0);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < TM; ++i) s += valid[i] ? acc[i][j] : 0.f;
      red[ty * BN + (j >> 2) * (BN / 2) + tx * 4 + (j & 3)] = s;
    }
    __syncthreads();
    float colsum = 0.f;
    if (tid < BN) {
#pragma unroll
      for (int r = 0; r < 16; ++r) colsum += red[r * BN + tid];
      smean[tid] = colsum / (float)nvalid;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = (j >> 2) * (BN / 2) + tx * 4 + (j & 3);
      const float mu = smean[col];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        float d = acc[i][j] - mu;
        s += valid[i] ? d * d : 0.f;
      }
      red[ty * BN + col] = s;
    }
    __syncthreads();
    if (tid < BN) {
      float m2 = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) m2 += red[r * BN + tid];
      float2* dst = reinterpret_cast<float2*>(p.part + ((int64_t)blockIdx.x * p.N + n0 + tid) * 2);
      *dst = make_float2(colsum, m2);
    }
  }
}

// dW[m, n] += sum_{r in chunk} dY[r, m] * X[(r + roff) * Cin + n]
template <int BM, int BN>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(WgradGemm p, int rows_per_split) {
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int HM = TM / 4, HN = TN / 4;
  constexpr int LA = BM * BK / 4 / 256, LB = BN * BK / 4 / 256;  // float4 per thread per k-tile (1 or 2)
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int r_begin = blockIdx.z * rows_per_split;
  const int r_end = min(p.R, r_begin + rows_per_split);
  if (r_begin >= r_end) return;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // loader mapping: item id -> (kk = id / (BM/4), quad = id % (BM/4))
  float4 ra[LA], rb[LB];
  auto load = [&](int r0) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int id = tid + i * 256, kk = id / (BM / 4), q = id % (BM / 4);
      int r = r0 + kk;
      ra[i] = r < r_end ? ldg4(p.dY + (int64_t)r * p.M + m0 + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int id = tid + i * 256, kk = id / (BN / 4), q = id % (BN / 4);
      int r = r0 + kk;
      rb[i] = r < r_end ? ldg4(p.X + ((int64_t)r + p.roff) * p.Cin + n0 + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int id = tid + i * 256, kk = id / (BM / 4), q = id % (BM / 4);
      *reinterpret_cast<float4*>(&As[buf][kk][q * 4]) = ra[i];
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int id = tid + i * 256, kk = id / (BN / 4), q = id % (BN / 4);
      *reinterpret_cast<float4*>(&Bs[buf][kk][q * 4]) = rb[i];
    }
  };

  const int nk = (r_end - r_begin + BK - 1) / BK;
  load(r_begin);
  stage(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) load(r_begin + (kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int h = 0; h < HM; ++h) {
        float4 v = *reinterpret_cast<const float4*>(&As[cur][kk][h * (BM / 2) + ty * 4]);
        a[h * 4 + 0] = v.x, a[h * 4 + 1] = v.y, a[h * 4 + 2] = v.z, a[h * 4 + 3] = v.w;
      }
#pragma unroll
      for (int h = 0; h < HN; ++h) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[cur][kk][h * (BN / 2) + tx * 4]);
        b[h * 4 + 0] = v.x, b[h * 4 + 1] = v.y, b[h * 4 + 2] = v.z, b[h * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) stage(cur ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int hi = 0; hi < TM; ++hi) {
    int m = m0 + (hi >> 2) * (BM / 2) + ty * 4 + (hi & 3);
    float* drow = p.dW + (int64_t)m * p.N + n0;
#pragma unroll
    for (int j = 0; j < TN; ++j) atomicAdd(drow + (j >> 2) * (BN / 2) + tx * 4 + (j & 3), acc[hi][j]);
  }
}

}  // namespace

int launch_conv_gemm_simt(const ConvGemm& g, cudaStream_t s) {
  const bool n128 = (g.N % 128) == 0;
  const int tiles128 = ((g.M + 127) / 128) * (g.N / (n128 ? 128 : 64));
  const bool m128 = tiles128 >= 120;  // prefer 64-row tiles when 128-row tiles cannot fill the 148 SMs
  const int BM = m128 ? 128 : 64;
  dim3 grid((g.M + BM - 1) / BM, g.N / (n128 ? 128 : 64));
  if (m128 && n128)
    conv_gemm_simt_kernel<128, 128><<<grid, 256, 0, s>>>(g);
  else if (m128)
    conv_gemm_simt_kernel<128, 64><<<grid, 256, 0, s>>>(g);
  else if (n128)
    conv_gemm_simt_kernel<64, 128><<<grid, 256, 0, s>>>(g);
  else
    conv_gemm_simt_kernel<64, 64><<<grid, 256, 0, s>>>(g);
  return BM;
}

void launch_wgrad_simt(const WgradGemm& g, int sm_count, cudaStream_t s) {
  const bool m128 = (g.M % 128) == 0, n128 = (g.N % 128) == 0;
  const int BM = m128 ? 128 : 64, BN = n128 ? 128 : 64;
  const int tiles = (g.M / BM) * (g.N / BN);
  int splits = (3 * sm_count + tiles - 1) / tiles;
  int max_splits = (g.R + 4 * BK - 1) / (4 * BK);  // at least 64 rows per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int rows_per_split = (g.R + splits - 1) / splits;
  rows_per_split = (rows_per_split + BK - 1) / BK * BK;
  splits = (g.R + rows_per_split - 1) / rows_per_split;
  dim3 grid(g.N / BN, g.M / BM, splits);
  if (m128 && n128)
    wgrad_simt_kernel<128, 128><<<grid, 256, 0, s>>>(g, rows_per_split);
  else if (m128)
    wgrad_simt_kernel<128, 64><<<grid, 256, 0, s>>>(g, rows_per_split);
  else if (n128)
    wgrad_simt_kernel<64, 128><<<grid, 256, 0, s>>>(g, rows_per_split);
  else
    wgrad_simt_kernel<64, 64><<<grid, 256, 0, s>>>(g, rows_per_split);
}

}  // namespace hp
