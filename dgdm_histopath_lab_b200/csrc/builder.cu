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


// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core candidate pass (config 3: 1M x 768 rows). The N^2 x 768 Gram work runs on tcgen05 through dgdm_gemm_f16x2
// (3-term split-fp16 products, fp32 accumulate: |error| <~ 1e-5 |x||y|), 512 key columns at a time; per query row a sorted
// list of the KC = 32 nearest candidates by APPROXIMATE distance is maintained across the column groups; the candidates are
// then re-ranked with sklearn's exact fp64 arithmetic. A row is flagged when exactness cannot be PROVED, i.e. unless
//     (approximate distance of the worst kept candidate) - eps  >  (exact distance of the k-th neighbour),
// in which case some non-candidate could still belong to the true top-k; the caller recomputes flagged rows with the exact
// fp64 path (dgdm_knn_cosine). Results are therefore always exactly the reference's.
constexpr int KNN_KC = 32;

__global__ void knn_cand_init_kernel(float* __restrict__ cd, int* __restrict__ ci, size_t total) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) { cd[i] = INFINITY; ci[i] = 0x7fffffff; }
}

// warp per query row; lane l holds the l-th smallest (distance, index) so far. G [nq, ncols] = approximate dot products.
__global__ void __launch_bounds__(256) knn_cand_merge_kernel(const float* __restrict__ G, int ldg, int nq, int ncols, int g0, int n,
                                                             int q0, const double* __restrict__ sq, float* __restrict__ cd,
                                                             int* __restrict__ ci) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= nq) return;
  float ld = cd[(size_t)row * KNN_KC + lane];
  int li = ci[(size_t)row * KNN_KC + lane];
  float thr = __shfl_sync(0xffffffffu, ld, 31);
  const double sqq = sq[q0 + row];
  const float* g = G + (size_t)row * ldg;
  for (int c = 0; c < ncols; c += 32) {
    const int j = g0 + c + lane;
    float v = INFINITY;
    if (c + lane < ncols && j < n) {
      const double d2 = sqq + sq[j] - 2.0 * (double)g[c + lane];
      v = d2 > 0.0 ? (float)d2 : 0.f;
    }
    unsigned mask = __ballot_sync(0xffffffffu, v < thr);
    while (mask) {
      const int b = __ffs(mask) - 1;
      mask &= mask - 1;
      const float vb = __shfl_sync(0xffffffffu, v, b);
      const int jb = g0 + c + b;
      if (!(vb < thr)) continue;                                  // the threshold moved since the ballot
      const int pos = __popc(__ballot_sync(0xffffffffu, ld <= vb));   // entries that stay in front of the new one
      const float up_d = __shfl_up_sync(0xffffffffu, ld, 1);
      const int up_i = __shfl_up_sync(0xffffffffu, li, 1);
      if (lane > pos) { ld = up_d; li = up_i; }
      else if (lane == pos) { ld = vb; li = jb; }
      thr = __shfl_sync(0xffffffffu, ld, 31);
    }
  }
  cd[(size_t)row * KNN_KC + lane] = ld;
  ci[(size_t)row * KNN_KC + lane] = li;
}

// warp per query row: exact fp64 squared distances of the 32 candidates (sklearn: |x|^2 - 2 x.y + |y|^2, clipped at 0), sorted by
// (distance, index); the first k are written; flag = 1 unless exactness is proved (see above).
__global__ void __launch_bounds__(128) knn_rerank_kernel(const float* __restrict__ F, int n, int d, int q0, int nq, int k,
                                                         const double* __restrict__ sq, const double* __restrict__ sq_max_ptr, const float* __restrict__ cd,
                                                         const int* __restrict__ ci, int* __restrict__ idx_out,
                                                         double* __restrict__ dist_out, int* __restrict__ flags) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= nq) return;
  const int my_i = ci[(size_t)row * KNN_KC + lane];
  const float worst_approx = __shfl_sync(0xffffffffu, cd[(size_t)row * KNN_KC + lane], 31);
  const float* xq = F + (size_t)(q0 + row) * d;
  double my_d = INFINITY;
  for (int c = 0; c < KNN_KC; ++c) {
    const int j = __shfl_sync(0xffffffffu, my_i, c);
    if (j == 0x7fffffff) continue;                                 // fewer than 32 rows exist
    const float* xj = F + (size_t)j * d;
    double acc = 0.0;
    for (int t = lane; t < d; t += 32) acc = fma((double)xq[t], (double)xj[t], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const double d2 = sq[q0 + row] + (-2.0 * acc) + sq[j];
    if (lane == c) my_d = d2 > 0.0 ? d2 : 0.0;
  }
  // rank of each candidate among the 32 by (distance, index): O(32) compares per lane
  int rank = 0;
  for (int c = 0; c < KNN_KC; ++c) {
    const double od = __shfl_sync(0xffffffffu, my_d, c);
    const int oi = __shfl_sync(0xffffffffu, my_i, c);
    if (c != lane && knn_less(od, oi, my_d, my_i)) ++rank;
  }
  if (rank < k && my_i != 0x7fffffff) { idx_out[(size_t)row * k + rank] = my_i; dist_out[(size_t)row * k + rank] = sqrt(my_d); }
  // exactness proof
  double kth = 0.0;
  {
    const unsigned who = __ballot_sync(0xffffffffu, rank == k - 1);
    const int src = who ? __ffs(who) - 1 : 0;
    kth = __shfl_sync(0xffffffffu, my_d, src);
  }
  if (lane == 0) {
    const double eps = 4e-5 * fmax(sq[q0 + row], *sq_max_ptr) + 1e-30;
    const bool all_seen = n <= KNN_KC;
    flags[row] = (all_seen || (double)worst_approx - eps > kth) ? 0 : 1;
  }
}

__global__ void max_f64_kernel(const double* __restrict__ v, int n, double* __restrict__ out) {
  __shared__ double red[32];
  double m = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, v[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) { for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmax(m, red[i]); *out = m; }
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

extern "C" int dgdm_gemm_f16x2_strided(const dgdm_gemm_params* p, long long w_plane_stride_bytes, void* stream);

extern "C" int dgdm_row_sqnorm_f64(const float* F, int n, int d, double* sq, void* stream) {
  if (!F || !sq || n <= 0 || d <= 0) return DGDM_ERR_INVALID;
  row_sqnorm_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(F, n, d, sq);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

constexpr int KNN_TC_GROUP = 512;
extern "C" size_t dgdm_knn_cosine_tc_work_bytes(int nq) {
  return al256((size_t)nq * KNN_TC_GROUP * 4) + 2 * al256((size_t)nq * KNN_KC * 4) + 512;
}

// F: [n_pad, d] fp32 with n_pad = round_up(n, 16) (pad rows zero), W_img = dgdm_split_weight_f16 images of the same matrix
// (n_pad rows), sq = fp64 squared row norms [n_pad]. d % 4 == 0. Query rows [q0, q0+nq). k <= 32 neighbours per row incl. the row
// itself. row_flags[nq]: 0 = provably exact, 1 = recompute this row with dgdm_knn_cosine.
extern "C" int dgdm_knn_cosine_tc(const float* F, const void* W_img, const double* sq, int n, int n_pad, int d, int q0, int nq, int k,
                                  int* nbr_idx, float* nbr_cos, double* nbr_dist, int* row_flags, void* work, size_t work_bytes,
                                  void* stream) {
  if (!F || !W_img || !sq || !nbr_idx || !nbr_cos || !nbr_dist || !row_flags || !work) return DGDM_ERR_INVALID;
  if (n <= 0 || n_pad < n || (n_pad % 16) != 0 || d <= 0 || (d % 4) != 0 || k <= 0 || k > KNN_KC || k > n) return DGDM_ERR_INVALID;
  if (q0 < 0 || nq <= 0 || q0 + nq > n || work_bytes < dgdm_knn_cosine_tc_work_bytes(nq)) return DGDM_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* w = (uint8_t*)(((uintptr_t)work + 255) & ~(uintptr_t)255);
  float* G = (float*)w; w += al256((size_t)nq * KNN_TC_GROUP * 4);
  float* cd = (float*)w; w += al256((size_t)nq * KNN_KC * 4);
  int* ci = (int*)w; w += al256((size_t)nq * KNN_KC * 4);
  double* sq_max = (double*)w;
  knn_cand_init_kernel<<<148 * 4, 256, 0, s>>>(cd, ci, (size_t)nq * KNN_KC);
  max_f64_kernel<<<1, 1024, 0, s>>>(sq, n, sq_max);
  const int kpad = (d + 63) / 64 * 64;
  for (int g0 = 0; g0 < n_pad; g0 += KNN_TC_GROUP) {
    const int nc = (n_pad - g0) < KNN_TC_GROUP ? (n_pad - g0) : KNN_TC_GROUP;
    dgdm_gemm_params p = {};
    p.A = F + (size_t)q0 * d; p.lda = d; p.M = nq; p.K = d;
    p.W_hi = (const uint8_t*)W_img + (size_t)g0 * 128; p.W_lo = (const uint8_t*)p.W_hi + (size_t)n_pad * 128; p.ldw = kpad;
    p.N = nc; p.C = G; p.ldc = KNN_TC_GROUP; p.a_scale = 1.0f; p.ln_eps = 1e-5f;
    const int rc = dgdm_gemm_f16x2_strided(&p, (long long)n_pad * 128, stream);
    if (rc != DGDM_OK) return rc;
    knn_cand_merge_kernel<<<(nq + 7) / 8, 256, 0, s>>>(G, KNN_TC_GROUP, nq, nc, g0, n, q0, sq, cd, ci);
  }
  knn_rerank_kernel<<<(nq + 3) / 4, 128, 0, s>>>(F, n, d, q0, nq, k, sq, sq_max, cd, ci, nbr_idx, nbr_dist, row_flags);
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
