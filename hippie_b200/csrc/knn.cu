// k-nearest-neighbour evaluation of embeddings on the device: the step that follows the embedding pass in every
// stage-3 run of the reference (scripts/train_model_with_multimodal.py:916-934: KNeighborsClassifier(k).fit(train).
// predict(test) for k = 5..19, balanced_accuracy_score, confusion_matrix).  Three kernels:
//
//   knn_neighbors_kernel   one warp per query row; train rows staged through shared memory in tiles shared by the
//                          CTA's 8 queries; squared Euclidean distance in double, summed in feature order with separate
//                          multiply and add (what the KD-tree / brute-force code of scikit-learn evaluates for float32
//                          input, which it widens to float64); the warp keeps its k <= 32 best (distance, index) pairs
//                          sorted ONE PER LANE and inserts a candidate with a ballot + shuffle.  Ties: lower index first.
//   knn_vote_kernel        one thread per query: majority vote among the first k neighbours for every k in
//                          [k_lo, k_hi] (ties: smallest class, as scipy's mode / argmax of the class counts), prediction
//                          table and confusion-matrix counts (integer atomics, exact).
//   knn_balacc_kernel      balanced accuracy per k from the confusion matrix: mean recall over the classes that occur in
//                          the true labels, summed in numpy's order (sequential below 8 terms, 8 interleaved partial
//                          sums above) so that the double result is the one sklearn prints.
//
// Distances are O(n_query * n_train * dim) double operations; the labelled sets this runs on have 10^2..10^5 rows.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/hippie_b200.h"

#define KNN_API extern "C" __attribute__((visibility("default")))

namespace {

constexpr int kWarps = 8;             // queries per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kTileFloats = 10240;    // 40 KB of train rows per tile

__global__ void __launch_bounds__(kThreads) knn_neighbors_kernel(const float* __restrict__ train, long long n_train,
                                                                const float* __restrict__ query, long long n_query,
                                                                int dim, int k, int tile_pts,
                                                                long long* __restrict__ out_idx,
                                                                double* __restrict__ out_d2) {
  extern __shared__ float sm[];
  const int stride = dim | 1;  // odd row stride: lanes read different banks
  float* tile = sm;
  float* qv = sm + (size_t)tile_pts * stride;  // [kWarps][dim]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * kWarps + warp;
  const bool live = q < n_query;
  if (live)
    for (int j = lane; j < dim; j += 32) qv[warp * dim + j] = query[q * dim + j];
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  double best_d = INF;       // lane i: i-th smallest squared distance so far (lanes >= k stay +inf)
  long long best_i = -1;
  for (long long base = 0; base < n_train; base += tile_pts) {
    const int npts = (int)min((long long)tile_pts, n_train - base);
    __syncthreads();
    for (int e = threadIdx.x; e < npts * dim; e += kThreads) {
      int p = e / dim, j = e - p * dim;
      tile[p * stride + j] = train[base * dim + e];
    }
    __syncthreads();
    if (!live) continue;
    const float* qq = qv + warp * dim;
    for (int p0 = 0; p0 < npts; p0 += 32) {
      const int p = p0 + lane;
      double d = INF;
      if (p < npts) {
        const float* t = tile + p * stride;
        d = 0.0;
        for (int j = 0; j < dim; ++j) {
          double diff = __dsub_rn((double)qq[j], (double)t[j]);
          d = __dadd_rn(d, __dmul_rn(diff, diff));
        }
      }
      double thr = __shfl_sync(0xffffffffu, best_d, k - 1);
      unsigned cand = __ballot_sync(0xffffffffu, d < thr);
      while (cand) {
        const int src = __ffs(cand) - 1;
        cand &= cand - 1;
        const double cd = __shfl_sync(0xffffffffu, d, src);
        if (!(cd < thr)) continue;  // the list tightened since the ballot
        const long long ci = base + p0 + src;
        // position = number of kept entries that are <= the candidate (candidates arrive in index order, so equal
        // distances keep the lower index in front)
        const int pos = __popc(__ballot_sync(0xffffffffu, best_d <= cd));
        const double up_d = __shfl_up_sync(0xffffffffu, best_d, 1);
        const long long up_i = __shfl_up_sync(0xffffffffu, best_i, 1);
        if (lane > pos) { best_d = up_d; best_i = up_i; }
        if (lane == pos) { best_d = cd; best_i = ci; }
        if (lane >= k) { best_d = INF; best_i = -1; }
        thr = __shfl_sync(0xffffffffu, best_d, k - 1);
      }
    }
  }
  if (live && lane < k) {
    out_idx[q * k + lane] = best_i;
    if (out_d2) out_d2[q * k + lane] = best_d;
  }
}

__global__ void knn_vote_kernel(const long long* __restrict__ neigh, long long n_query, int k_stride,
                                const long long* __restrict__ train_class, const long long* __restrict__ true_class,
                                int n_classes, int k_lo, int k_hi, long long* __restrict__ pred,
                                unsigned long long* __restrict__ confusion) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_query) return;
  int lab[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) lab[i] = i < k_hi ? (int)train_class[neigh[q * k_stride + i]] : -1;
  for (int k = k_lo; k <= k_hi; ++k) {
    int best_c = -1, best_n = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i >= k) break;
      int n = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) n += (j < k && lab[j] == lab[i]) ? 1 : 0;
      if (n > best_n || (n == best_n && lab[i] < best_c)) { best_n = n; best_c = lab[i]; }
    }
    if (pred) pred[(long long)(k - k_lo) * n_query + q] = best_c;
    if (confusion && true_class) {
      const long long t = true_class[q];
      if (t >= 0 && t < n_classes)
        atomicAdd(&confusion[((size_t)(k - k_lo) * n_classes + t) * n_classes + best_c], 1ULL);
    }
  }
}

// numpy's add.reduce over a contiguous double vector of n <= 128 terms (pairwise_sum in loops_utils.h): plain loop
// below 8 terms, otherwise 8 running sums combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus the tail in order.
__global__ void knn_balacc_kernel(const unsigned long long* __restrict__ confusion, int n_classes, int nk,
                                  double* __restrict__ out) {
  const int ki = blockIdx.x * blockDim.x + threadIdx.x;
  if (ki >= nk) return;
  const unsigned long long* cm = confusion + (size_t)ki * n_classes * n_classes;
  int n = 0;
  for (int c = 0; c < n_classes; ++c) {
    unsigned long long rs = 0;
    for (int j = 0; j < n_classes; ++j) rs += cm[(size_t)c * n_classes + j];
    n += rs ? 1 : 0;
  }
  if (n == 0) { out[ki] = __longlong_as_double(0x7ff8000000000000LL); return; }
  const int blocked = n < 8 ? 0 : n - (n % 8);
  double r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double res = 0.0;
  int seen = 0;
  for (int c = 0; c < n_classes; ++c) {
    unsigned long long rs = 0;
    for (int j = 0; j < n_classes; ++j) rs += cm[(size_t)c * n_classes + j];
    if (!rs) continue;
    const double recall = __ddiv_rn((double)cm[(size_t)c * n_classes + c], (double)rs);
    if (seen < blocked) {
      const int s = seen & 7;
      r[s] = seen < 8 ? recall : __dadd_rn(r[s], recall);
      if (seen == blocked - 1)
        res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                        __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    } else {
      res = __dadd_rn(res, recall);
    }
    ++seen;
  }
  out[ki] = __ddiv_rn(res, (double)n);
}

}  // namespace

KNN_API int hippie_knn_neighbors(const float* train, int64_t n_train, const float* query, int64_t n_query, int32_t dim,
                                 int32_t k, int64_t* out_index, double* out_sqdist, void* stream) {
  if (n_query == 0) return 0;
  if (!train || !query || !out_index || dim < 1 || dim > 1024 || k < 1 || k > 32 || n_train < k || n_query < 0) return -1;
  const int stride = dim | 1;
  int tile_pts = (kTileFloats / stride) & ~31;
  if (tile_pts < 32) tile_pts = 32;
  const size_t smem = ((size_t)tile_pts * stride + (size_t)kWarps * dim) * sizeof(float);
  static bool attr_set = false;  // one device per process (include/hippie_b200.h)
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(knn_neighbors_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const long long grid = (n_query + kWarps - 1) / kWarps;
  if (grid > 0x7fffffffLL) return -1;
  knn_neighbors_kernel<<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(
      train, (long long)n_train, query, (long long)n_query, dim, k, tile_pts, (long long*)out_index, out_sqdist);
  return (int)cudaGetLastError();
}

KNN_API int hippie_knn_evaluate(const int64_t* neighbors, int64_t n_query, int32_t k_stride, const int64_t* train_class,
                                const int64_t* true_class, int32_t n_classes, int32_t k_lo, int32_t k_hi,
                                int64_t* out_pred, int64_t* out_confusion, double* out_balanced_accuracy, void* stream) {
  if (!neighbors || !train_class || k_lo < 1 || k_hi < k_lo || k_hi > 32 || k_hi > k_stride || n_classes < 1 ||
      n_classes > 128 || n_query < 0)
    return -1;
  if ((out_confusion || out_balanced_accuracy) && !(true_class && out_confusion)) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  const int nk = k_hi - k_lo + 1;
  if (out_confusion) {
    cudaError_t e = cudaMemsetAsync(out_confusion, 0, sizeof(int64_t) * (size_t)nk * n_classes * n_classes, s);
    if (e != cudaSuccess) return (int)e;
  }
  if (n_query > 0) {
    const long long grid = (n_query + 127) / 128;
    if (grid > 0x7fffffffLL) return -1;
    knn_vote_kernel<<<(unsigned)grid, 128, 0, s>>>((const long long*)neighbors, (long long)n_query, k_stride,
                                                   (const long long*)train_class, (const long long*)true_class, n_classes,
                                                   k_lo, k_hi, (long long*)out_pred, (unsigned long long*)out_confusion);
  }
  if (out_balanced_accuracy)
    knn_balacc_kernel<<<1, 32, 0, s>>>((const unsigned long long*)out_confusion, n_classes, nk, out_balanced_accuracy);
  return (int)cudaGetLastError();
}
