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

// ---------------------------------------------------------------------------------------------------------------------
// Exact 2-D k nearest neighbours on a uniform grid (spatial edges, preprocessing/tissue_graph_builder.py:290-320): O(N) instead
// of the O(N^2) distance matrix. Points are binned into G x G cells over their bounding box (counting sort by cell); a query
// scans the cells ring by ring (Chebyshev distance r around its own cell) and stops after ring r once it holds k candidates whose
// k-th squared distance is strictly below (r*h + m)^2, h = cell size, m = distance from the query to the nearest edge of its
// cell — every unvisited cell lies at least that far away. Distances are fp64 direct differences and ties break towards the lower
// index, exactly like the brute-force kernel, so the result is identical to it (and to sklearn's kd-tree on tie-free data).
struct Grid2D { double x0, y0, inv_h, h; int G; };

__global__ void grid_bbox_kernel(const double* __restrict__ P, int n, double* __restrict__ bb /*minx,miny,maxx,maxy*/) {
  __shared__ double sm[4][32];
  double mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double x = P[2 * (size_t)i], y = P[2 * (size_t)i + 1];
    mnx = fmin(mnx, x); mny = fmin(mny, y); mxx = fmax(mxx, x); mxy = fmax(mxy, y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sm[0][w] = mnx; sm[1][w] = mny; sm[2][w] = mxx; sm[3][w] = mxy; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) { mnx = fmin(mnx, sm[0][i]); mny = fmin(mny, sm[1][i]); mxx = fmax(mxx, sm[2][i]); mxy = fmax(mxy, sm[3][i]); }
    bb[0] = mnx; bb[1] = mny; bb[2] = mxx; bb[3] = mxy;
  }
}
__device__ __forceinline__ Grid2D grid_from_bbox(const double* bb, int G) {
  Grid2D g;
  const double ext = fmax(fmax(bb[2] - bb[0], bb[3] - bb[1]), 1e-300);
  g.x0 = bb[0]; g.y0 = bb[1]; g.G = G; g.h = ext / (double)G; g.inv_h = (double)G / ext;
  return g;
}
__device__ __forceinline__ int grid_cell(const Grid2D& g, double v, double v0) {
  int c = (int)floor((v - v0) * g.inv_h);
  return c < 0 ? 0 : (c >= g.G ? g.G - 1 : c);
}
__global__ void grid_count_kernel(const double* __restrict__ P, int n, const double* __restrict__ bb, int G, int* __restrict__ cell_of,
                                  int* __restrict__ counts) {
  const Grid2D g = grid_from_bbox(bb, G);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = grid_cell(g, P[2 * (size_t)i + 1], g.y0) * G + grid_cell(g, P[2 * (size_t)i], g.x0);
    cell_of[i] = c;
    atomicAdd(counts + c, 1);
  }
}
// exclusive scan of counts[0..m) into start[0..m] by one block
__global__ void __launch_bounds__(1024) grid_scan_kernel(const int* __restrict__ counts, int* __restrict__ start, int m) {
  __shared__ int wt[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < m; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < m ? counts[i] : 0;
    int sc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
    if (lane == 31) wt[warp] = sc;
    __syncthreads();
    if (warp == 0) {
      int w = wt[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      wt[lane] = w;
    }
    __syncthreads();
    const int carry = carry_s;
    if (i < m) start[i] = carry + (warp ? wt[warp - 1] : 0) + sc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wt[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) start[m] = carry_s;
}
__global__ void grid_fill_kernel(const int* __restrict__ cell_of, int n, const int* __restrict__ start, int* __restrict__ cursor,
                                 int* __restrict__ sorted_idx) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = cell_of[i];
    sorted_idx[start[c] + atomicAdd(cursor + c, 1)] = i;      // order inside a cell is irrelevant: selection is by (distance, index)
  }
}
template <int K>
__global__ void __launch_bounds__(128) grid_knn_kernel(const double* __restrict__ P, int n, int q0, int nq, int k, const double* __restrict__ bb,
                                                       int G, const int* __restrict__ start, const int* __restrict__ sorted_idx,
                                                       int* __restrict__ idx_out, double* __restrict__ dist_out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const Grid2D g = grid_from_bbox(bb, G);
  const double qx = P[2 * (size_t)(q0 + q)], qy = P[2 * (size_t)(q0 + q) + 1];
  const int cx = grid_cell(g, qx, g.x0), cy = grid_cell(g, qy, g.y0);
  double bd[K]; int bi[K];
#pragma unroll
  for (int i = 0; i < K; ++i) { bd[i] = INFINITY; bi[i] = 0x7fffffff; }
  // distance from the query to the nearest edge of its own cell (0 if the point was clamped into the cell)
  const double lx = qx - (g.x0 + cx * g.h), ly = qy - (g.y0 + cy * g.h);
  double m = fmin(fmin(lx, g.h - lx), fmin(ly, g.h - ly));
  if (!(m > 0.0)) m = 0.0;
  auto visit = [&](int ccx, int ccy) {
    const int c = ccy * G + ccx;
    for (int s = start[c]; s < start[c + 1]; ++s) {
      const int j = sorted_idx[s];
      const double dx = qx - P[2 * (size_t)j], dy = qy - P[2 * (size_t)j + 1];
      const double d2 = dx * dx + dy * dy;
      if (knn_less(d2, j, bd[k - 1], bi[k - 1])) {
        int p = k - 1;
        while (p > 0 && knn_less(d2, j, bd[p - 1], bi[p - 1])) { bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p; }
        bd[p] = d2; bi[p] = j;
      }
    }
  };
  for (int r = 0; r < G; ++r) {
    const int x_lo = cx - r, x_hi = cx + r, y_lo = cy - r, y_hi = cy + r;
    if (r == 0) visit(cx, cy);
    else {
      for (int xx = max(x_lo, 0); xx <= min(x_hi, G - 1); ++xx) {
        if (y_lo >= 0) visit(xx, y_lo);
        if (y_hi < G) visit(xx, y_hi);
      }
      for (int yy = max(y_lo + 1, 0); yy <= min(y_hi - 1, G - 1); ++yy) {
        if (x_lo >= 0) visit(x_lo, yy);
        if (x_hi < G) visit(x_hi, yy);
      }
    }
    if (x_lo <= 0 && y_lo <= 0 && x_hi >= G - 1 && y_hi >= G - 1) break;        // whole grid visited
    const double bound = ((double)r * g.h + m) * (1.0 - 1e-12);
    if (bi[k - 1] != 0x7fffffff && bd[k - 1] < bound * bound) break;
  }
  for (int o = 0; o < k; ++o) { idx_out[(size_t)q * k + o] = bi[o]; dist_out[(size_t)q * k + o] = sqrt(bd[o]); }
}


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
    static PerDeviceOnce set;
    if (set.needed()) { if (cudaFuncSetAttribute(row_topk_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return DGDM_ERR_LAUNCH; set.mark(); }
    row_topk_kernel<K><<<blocks, 128, sm, s>>>(D2, n, nq, k, idx, dist);
  } else {
    constexpr int K = KNN_MAXK; const size_t sm = (size_t)4 * 32 * K * 12;
    static PerDeviceOnce set;
    if (set.needed()) { if (cudaFuncSetAttribute(row_topk_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return DGDM_ERR_LAUNCH; set.mark(); }
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

extern "C" size_t dgdm_knn2d_grid_work_bytes(int n) {
  int G = (int)ceil(sqrt((double)n / 3.0)); if (G < 1) G = 1;
  return al256(4 * 8) + al256((size_t)n * 4) * 2 + al256(((size_t)G * G + 1) * 4) * 3 + 256;
}
static int knn2d_grid(const double* coords, int n, int q0, int nq, int k, int* nbr_idx, double* nbr_dist, void* work, cudaStream_t s) {
  int G = (int)ceil(sqrt((double)n / 3.0)); if (G < 1) G = 1;
  const size_t cells = (size_t)G * G;
  uint8_t* w = (uint8_t*)(((uintptr_t)work + 255) & ~(uintptr_t)255);
  double* bb = (double*)w; w += al256(4 * 8);
  int* cell_of = (int*)w; w += al256((size_t)n * 4);
  int* sorted_idx = (int*)w; w += al256((size_t)n * 4);
  int* counts = (int*)w; w += al256((cells + 1) * 4);
  int* start = (int*)w; w += al256((cells + 1) * 4);
  int* cursor = (int*)w;
  if (cudaMemsetAsync(counts, 0, (cells + 1) * 4, s) != cudaSuccess || cudaMemsetAsync(cursor, 0, (cells + 1) * 4, s) != cudaSuccess) return DGDM_ERR_LAUNCH;
  grid_bbox_kernel<<<1, 1024, 0, s>>>(coords, n, bb);
  const int blocks = (n + 255) / 256 > 148 * 8 ? 148 * 8 : (n + 255) / 256;
  grid_count_kernel<<<blocks, 256, 0, s>>>(coords, n, bb, G, cell_of, counts);
  grid_scan_kernel<<<1, 1024, 0, s>>>(counts, start, (int)cells);
  grid_fill_kernel<<<blocks, 256, 0, s>>>(cell_of, n, start, cursor, sorted_idx);
  const int qb = (nq + 127) / 128;
  if (k <= 16) grid_knn_kernel<16><<<qb, 128, 0, s>>>(coords, n, q0, nq, k, bb, G, start, sorted_idx, nbr_idx, nbr_dist);
  else if (k <= 32) grid_knn_kernel<32><<<qb, 128, 0, s>>>(coords, n, q0, nq, k, bb, G, start, sorted_idx, nbr_idx, nbr_dist);
  else grid_knn_kernel<KNN_MAXK><<<qb, 128, 0, s>>>(coords, n, q0, nq, k, bb, G, start, sorted_idx, nbr_idx, nbr_dist);
  if (cudaGetLastError() != cudaSuccess) return DGDM_ERR_LAUNCH;
  return DGDM_OK;
}

// mode 0: brute force (O(N^2) fp64 distance rows, small n); mode 1: uniform grid (O(N)); work: dgdm_knn2d_grid_work_bytes(n)
// bytes for mode 1 (may be NULL for mode 0). Both return identical results.
extern "C" int dgdm_knn2d(const double* coords, int n, int q0, int nq, int k, int* nbr_idx, double* nbr_dist, int mode, void* work,
                          void* stream) {
  if (!coords || !nbr_idx || !nbr_dist || n <= 0 || k <= 0 || k > KNN_MAXK || k > n || q0 < 0 || nq <= 0 || q0 + nq > n) return DGDM_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == 1) {
    if (!work) return DGDM_ERR_INVALID;
    return knn2d_grid(coords, n, q0, nq, k, nbr_idx, nbr_dist, work, s);
  }
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
