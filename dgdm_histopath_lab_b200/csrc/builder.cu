// TissueGraphBuilder edge construction (preprocessing/tissue_graph_builder.py:290-352) on the GPU.
//
//  * morphological edges: for each query row the k+1 nearest rows by Euclidean distance, ranked on fp64 squared
//    distances from the Gram identity |x|^2 - 2 x.y + |y|^2 exactly as sklearn's EuclideanArgKmin32 does for
//    float32 input (chunks upcast to float64), then the cosine similarity of each kept pair with sklearn
//    cosine_similarity semantics (rows normalised in float32, dot product rounded to float32).
//  * spatial edges: exact k+1 nearest neighbours of 2-D float64 coordinates.
//  The caller (builder.py) drops neighbour 0 ("skip self", :303/:338) and applies the thresholds.
//
// Round-1 implementation: fp64 Gram tiles on CUDA cores (64x64 tile, 4x4 per thread) + warp-per-row top-k
// selection. The tcgen05 candidate generator + fp64 re-rank that config 3 (1M rows) needs is the next step
// (DESIGN.md "edge builder"); this version is exact and is what the parity tests pin.
#include "common.cuh"

namespace dgdm {

constexpr int KNN_MAXK = 64;

// squared row norms in fp64
__global__ void __launch_bounds__(256) row_sqnorm_kernel(const float* __restrict__ F, int n, int d, double* __restrict__ sq) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  double s = 0.0;
  for (int c = lane; c < d; c += 32) { const double v = (double)F[(size_t)row * d + c]; s += v * v; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) sq[row] = s;
}

// D2[q, j] = max(0, sq[q0+q] - 2 F[q0+q].F[j] + sq[j])  for q in [0,nq), j in [0,n): 64x64 tiles, fp64 FMA
__global__ void __launch_bounds__(256) gram_dist_kernel(const float* __restrict__ F, int n, int d, int q0, int nq,
                                                        const double* __restrict__ sq, double* __restrict__ D2) {
  __shared__ double As[16][64 + 1];
  __shared__ double Bs[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < d; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, kk = i & 15;
      As[kk][r] = (m0 + r < nq && k0 + kk < d) ? (double)F[(size_t)(q0 + m0 + r) * d + k0 + kk] : 0.0;
      Bs[kk][r] = (n0 + r < n && k0 + kk < d) ? (double)F[(size_t)(n0 + r) * d + k0 + kk] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, c = n0 + tx * 4 + j;
      if (m < nq && c < n) {
        const double v = sq[q0 + m] + (-2.0 * acc[i][j]) + sq[c];
        D2[(size_t)m * n + c] = v > 0.0 ? v : 0.0;
      }
    }
}

// 2-D squared distances in fp64 (direct differences, like sklearn's kd-tree reduced distance)
__global__ void __launch_bounds__(256) dist2d_kernel(const double* __restrict__ P, int n, int q0, int nq, double* __restrict__ D2) {
  const size_t total = (size_t)nq * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int q = (int)(i / n), j = (int)(i - (size_t)q * n);
    const double dx = P[2 * (size_t)(q0 + q)] - P[2 * (size_t)j], dy = P[2 * (size_t)(q0 + q) + 1] - P[2 * (size_t)j + 1];
    D2[i] = dx * dx + dy * dy;
  }
}

__device__ __forceinline__ bool knn_less(double da, int ia, double db, int ib) { return da < db || (da == db && ia < ib); }

// warp per row: the K smallest (value, index) pairs of D2[row, 0..n), ascending. Each lane keeps a sorted private list
// over its strided elements (in shared memory), then a 32-way merge picks the global K.
template <int K>
__global__ void __launch_bounds__(128) row_topk_kernel(const double* __restrict__ D2, int n, int nq, int k_out,
                                                       int* __restrict__ idx_out, double* __restrict__ dist_out) {
  extern __shared__ uint8_t sm[];
  const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp_in_block;
  double* lv = reinterpret_cast<double*>(sm) + (size_t)(warp_in_block * 32 + lane) * K;
  int* li = reinterpret_cast<int*>(sm + (size_t)4 * 32 * K * sizeof(double)) + (size_t)(warp_in_block * 32 + lane) * K;
  if (row >= nq) return;
  for (int i = 0; i < K; ++i) { lv[i] = INFINITY; li[i] = 0x7fffffff; }
  const double* r = D2 + (size_t)row * n;
  for (int j = lane; j < n; j += 32) {
    const double v = r[j];
    if (knn_less(v, j, lv[K - 1], li[K - 1])) {
      int p = K - 1;
      while (p > 0 && knn_less(v, j, lv[p - 1], li[p - 1])) { lv[p] = lv[p - 1]; li[p] = li[p - 1]; --p; }
      lv[p] = v; li[p] = j;
    }
  }
  __syncwarp();
  int head = 0;
  for (int o = 0; o < k_out; ++o) {
    double v = head < K ? lv[head] : INFINITY;
    int id = head < K ? li[head] : 0x7fffffff;
    double bv = v; int bi = id;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, s);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
      if (knn_less(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (bi == id && bv == v && head < K) ++head;       // indices are unique, so exactly one lane advances
    if (lane == 0) { idx_out[(size_t)row * k_out + o] = bi; dist_out[(size_t)row * k_out + o] = sqrt(bv); }
  }
}

// cosine similarity of (query row, neighbour) pairs with sklearn semantics: rows normalised in float32, dot -> float32
__global__ void __launch_bounds__(256) pair_cosine_kernel(const float* __restrict__ F, int d, int q0, int nq, int k,
                                                          const int* __restrict__ nbr, const double* __restrict__ sq,
                                                          float* __restrict__ cosv) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)nq * k) return;
  const int q = (int)(w / k);
  const int j = nbr[w];
  const int i = q0 + q;
  const float ni = (float)sqrt(sq[i]), nj = (float)sqrt(sq[j]);
  const float di = ni == 0.f ? 1.f : ni, dj = nj == 0.f ? 1.f : nj;         // sklearn normalize(): zero norms -> 1
  double acc = 0.0;
  for (int c = lane; c < d; c += 32) {
    const float a = F[(size_t)i * d + c] / di, b = F[(size_t)j * d + c] / dj;
    acc += (double)a * (double)b;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) cosv[w] = (float)acc;
}

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

static int launch_topk(const double* D2, int n, int nq, int k, int* idx, double* dist, cudaStream_t s) {
  const int blocks = (nq + 3) / 4;
  if (k <= 16) {
    constexpr int K = 16; const size_t sm = (size_t)4 * 32 * K * 12;
    row_topk_kernel<K><<<blocks, 128, sm, s>>>(D2, n, nq, k, idx, dist);
  } else if (k <= 32) {
    constexpr int K = 32; const size_t sm = (size_t)4 * 32 * K * 12;
    static bool set = false;
    if (!set) { if (cudaFuncSetAttribute(row_topk_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return DGDM_ERR_LAUNCH; set = true; }
    row_topk_kernel<K><<<blocks, 128, sm, s>>>(D2, n, nq, k, idx, dist);
  } else {
    constexpr int K = KNN_MAXK; const size_t sm = (size_t)4 * 32 * K * 12;
    static bool set = false;
    if (!set) { if (cudaFuncSetAttribute(row_topk_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return DGDM_ERR_LAUNCH; set = true; }
    row_topk_kernel<K><<<blocks, 128, sm, s>>>(D2, n, nq, k, idx, dist);
  }
  return DGDM_OK;
}

}  // namespace dgdm
using namespace dgdm;

// k here counts the neighbours returned per row INCLUDING the nearest one (the reference asks sklearn for k+1 and skips
// column 0), so callers pass morphological_k + 1 / spatial_k + 1.
extern "C" size_t dgdm_knn_cosine_work_bytes(int nq, int n, int d, int k) {
  (void)d; (void)k;
  return al256((size_t)n * 8) + al256((size_t)nq * n * 8) + 256;
}

extern "C" int dgdm_knn_cosine(const float* F, int n, int d, int q0, int nq, int k, int* nbr_idx, float* nbr_cos,
                               double* nbr_dist, void* work, size_t work_bytes, void* stream) {
  if (!F || !nbr_idx || !nbr_cos || !nbr_dist || !work || n <= 0 || d <= 0 || k <= 0 || k > KNN_MAXK || k > n) return DGDM_ERR_INVALID;
  if (q0 < 0 || nq <= 0 || q0 + nq > n) return DGDM_ERR_INVALID;
  if (work_bytes < dgdm_knn_cosine_work_bytes(nq, n, d, k)) return DGDM_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* w = (uint8_t*)(((uintptr_t)work + 255) & ~(uintptr_t)255);
  double* sq = (double*)w; w += al256((size_t)n * 8);
  double* D2 = (double*)w;
  row_sqnorm_kernel<<<(n + 7) / 8, 256, 0, s>>>(F, n, d, sq);
  dim3 grid((n + 63) / 64, (nq + 63) / 64);
  gram_dist_kernel<<<grid, 256, 0, s>>>(F, n, d, q0, nq, sq, D2);
  int rc = launch_topk(D2, n, nq, k, nbr_idx, nbr_dist, s);
  if (rc != DGDM_OK) return rc;
  const long long warps = (long long)nq * k;
  pair_cosine_kernel<<<(int)((warps + 7) / 8), 256, 0, s>>>(F, d, q0, nq, k, nbr_idx, sq, nbr_cos);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" int dgdm_knn2d(const double* coords, int n, int q0, int nq, int k, int* nbr_idx, double* nbr_dist, void* stream) {
  if (!coords || !nbr_idx || !nbr_dist || n <= 0 || k <= 0 || k > KNN_MAXK || k > n || q0 < 0 || nq <= 0 || q0 + nq > n) return DGDM_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  double* D2 = nullptr;
  // scratch for the distance rows of this query chunk (stream-ordered allocation; freed below)
  if (cudaMallocAsync((void**)&D2, (size_t)nq * n * sizeof(double), s) != cudaSuccess) return DGDM_ERR_LAUNCH;
  const size_t total = (size_t)nq * n;
  const size_t b = (total + 255) / 256;
  dist2d_kernel<<<(int)(b > 148 * 32 ? 148 * 32 : b), 256, 0, s>>>(coords, n, q0, nq, D2);
  int rc = launch_topk(D2, n, nq, k, nbr_idx, nbr_dist, s);
  cudaFreeAsync(D2, s);
  if (rc != DGDM_OK) return rc;
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}
