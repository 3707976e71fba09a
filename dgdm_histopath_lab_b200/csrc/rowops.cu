// Row-wise kernels: fp32 SIMT GEMM (bring-up / comparator), stand-alone fused epilogue (bias, residual, LayerNorm x2,
// activations) for widths that do not fit TMEM, GroupNorm+SiLU, diffusion element-wise ops, spatial positional
// encoding, un-pool, input validation. All HBM-bound: 128-bit coalesced row accesses, warp per row.
#include "common.cuh"

namespace dgdm {

// ------------------------------------------------------------------------------------------ SIMT GEMM
// C[M,N] = A[M,K] @ W[N,K]^T, 64x64 tile, 16x16 threads, 4x4 outputs per thread. Exact fp32 FMA chain.
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W,
                                                        int ldw, float* __restrict__ C, int ldc, int M, int N, int K,
                                                        float a_scale) {
  __shared__ float As[16][64 + 1];
  __shared__ float Ws[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, kk = i & 15;
      As[kk][r] = (m0 + r < M && k0 + kk < K) ? A[(size_t)(m0 + r) * lda + k0 + kk] : 0.f;
      Ws[kk][r] = (n0 + r < N && k0 + kk < K) ? W[(size_t)(n0 + r) * ldw + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) C[(size_t)m * ldc + n] = acc[i][j];
    }
}

// ---------------------------------------------------------------------------------- row-wise epilogue
// One warp per row, row kept in registers (N <= 128*MAXV). Same semantics as the GEMM epilogue.
template <int MAXV>
__global__ void __launch_bounds__(256) rowwise_epilogue_kernel(const float* __restrict__ X, int ldx, int M, int N,
                                                               const dgdm_gemm_params p) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int m = warp;
  float4 v[MAXV];
  const float invN = 1.0f / (float)N;
  const int rbi = p.row_bias_index ? p.row_bias_index[m] : 0;
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int c = (j * 32 + lane) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < N) {
      x = ldg4(X + (size_t)m * ldx + c);
      if (p.bias) { float4 b = ldg4(p.bias + c); x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w; }
      if (p.row_bias) { float4 b = ldg4(p.row_bias + (size_t)rbi * p.ld_row_bias + c); x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w; }
      if (p.r_pre) { float4 b = ldg4(p.r_pre + (size_t)m * p.ld_r_pre + c); x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w; }
      apply_act4(x, p.act_pre);
      sum += (x.x + x.y) + (x.z + x.w);
    }
    v[j] = x;
  }
  for (int pass = 0; pass < 2; ++pass) {
    const float* g = pass == 0 ? p.ln1_g : p.ln2_g;
    const float* b = pass == 0 ? p.ln1_b : p.ln2_b;
    if (g == nullptr) continue;
    if (pass == 1 && p.ln1_g != nullptr) {
      sum = 0.f;
#pragma unroll
      for (int j = 0; j < MAXV; ++j) if ((j * 32 + lane) * 4 < N) sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mean = warp_sum(sum) * invN;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) if ((j * 32 + lane) * 4 < N) {
      float d0 = v[j].x - mean, d1 = v[j].y - mean, d2 = v[j].z - mean, d3 = v[j].w - mean;
      ss += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    const float rstd = rsqrtf(warp_sum(ss) * invN + p.ln_eps);
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c < N) {
        float4 gg = ldg4(g + c), bb = ldg4(b + c);
        v[j].x = (v[j].x - mean) * rstd * gg.x + bb.x; v[j].y = (v[j].y - mean) * rstd * gg.y + bb.y;
        v[j].z = (v[j].z - mean) * rstd * gg.z + bb.z; v[j].w = (v[j].w - mean) * rstd * gg.w + bb.w;
      }
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int c = (j * 32 + lane) * 4;
    if (c < N) {
      float4 x = v[j];
      apply_act4(x, p.act_post);
      if (p.r_post) { float4 b = ldg4(p.r_post + (size_t)m * p.ld_r_post + c); x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w; }
      if (p.dot_w) { float4 w = ldg4(p.dot_w + c); dot += x.x * w.x + x.y * w.y + x.z * w.z + x.w * w.w; }
      if (p.C) *reinterpret_cast<float4*>(p.C + (size_t)m * p.ldc + c) = x;
    }
  }
  if (p.dot_out) {
    dot = warp_sum(dot);
    if (lane == 0) p.dot_out[m] = apply_act(dot + p.dot_b, p.dot_act);
  }
}

// ------------------------------------------------------------------------------------------ GroupNorm + SiLU
// nn.GroupNorm(G, C) on 2-D [M,C] input (core/diffusion.py:96-101 under P6): per row, per group of C/G channels.
__global__ void __launch_bounds__(256) groupnorm_silu_kernel(const float* __restrict__ X, int ldx, int M, int C, int G,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             float eps, float* __restrict__ out, int ldo) {
  // one warp per (row, group)
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)M * G) return;
  const int m = (int)(w / G), g = (int)(w - (long long)m * G);
  const int cg = C / G;
  const float* x = X + (size_t)m * ldx + g * cg;
  float s = 0.f;
  for (int c = lane; c < cg; c += 32) s += x[c];
  const float mean = warp_sum(s) / (float)cg;
  float ss = 0.f;
  for (int c = lane; c < cg; c += 32) { float d = x[c] - mean; ss += d * d; }
  const float rstd = rsqrtf(warp_sum(ss) / (float)cg + eps);
  for (int c = lane; c < cg; c += 32) {
    const int ch = g * cg + c;
    out[(size_t)m * ldo + ch] = silu((x[c] - mean) * rstd * gamma[ch] + beta[ch]);
  }
}

// ------------------------------------------------------------------------------------------ diffusion
__global__ void diffusion_add_noise_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                           const float* __restrict__ sa, const float* __restrict__ sb,
                                           const int* __restrict__ seg_ptr, const int* __restrict__ seg_t, int num_seg,
                                           int N, int C, float* __restrict__ out) {
  const size_t total = (size_t)N * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / C);
    int lo = 0, hi = num_seg - 1;           // segment of this row (binary search over seg_ptr)
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (seg_ptr[mid] <= row) lo = mid; else hi = mid - 1; }
    const int t = seg_t[lo];
    out[i] = sa[t] * x0[i] + sb[t] * noise[i];
  }
}
__global__ void ddpm_update_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ noise,
                                   float s1mac, float isac, float salpha, float svar, size_t n, float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float x0 = (x[i] - s1mac * eps[i]) * isac;     // core/diffusion.py:258-260 (division -> multiply by 1/sqrt(ac))
    out[i] = noise ? salpha * x0 + svar * noise[i] : x0;
  }
}

// ------------------------------------------------------------------------------------------ spatial PE
// per-segment min/max over all coordinates of pos[seg] (core/attention.py:238-240: scalar min/max of the graph's positions)
__global__ void __launch_bounds__(1024) segment_minmax_kernel(const float* __restrict__ pos, const int* __restrict__ seg_ptr,
                                                              float* __restrict__ minmax) {
  const int g = blockIdx.x;
  const int a = seg_ptr[g] * 2, b = seg_ptr[g + 1] * 2;       // [N,2] rows are 8-byte aligned: float2 loads, 4 in flight per thread
  float mn = INFINITY, mx = -INFINITY;
  const float2* p2 = reinterpret_cast<const float2*>(pos);
  const int a2 = a >> 1, b2 = b >> 1;
  int i = a2 + threadIdx.x;
  for (; i + 3 * 1024 < b2; i += 4 * 1024) {
    const float2 v0 = __ldg(p2 + i), v1 = __ldg(p2 + i + 1024), v2 = __ldg(p2 + i + 2048), v3 = __ldg(p2 + i + 3072);
    mn = fminf(fminf(fminf(mn, fminf(v0.x, v0.y)), fminf(v1.x, v1.y)), fminf(fminf(v2.x, v2.y), fminf(v3.x, v3.y)));
    mx = fmaxf(fmaxf(fmaxf(mx, fmaxf(v0.x, v0.y)), fmaxf(v1.x, v1.y)), fmaxf(fmaxf(v2.x, v2.y), fmaxf(v3.x, v3.y)));
  }
  for (; i < b2; i += 1024) { const float2 v = __ldg(p2 + i); mn = fminf(mn, fminf(v.x, v.y)); mx = fmaxf(mx, fmaxf(v.x, v.y)); }
  __shared__ float smn[32], smx[32];
  mn = warp_min(mn); mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x < 32) {
    mn = warp_min(smn[threadIdx.x]); mx = warp_max(smx[threadIdx.x]);
    if (threadIdx.x == 0) { minmax[2 * g] = mn; minmax[2 * g + 1] = mx; }
  }
}
// out = X + PE: channel 4f+0 = sin(x*div_f), 4f+1 = cos(x*div_f), 4f+2 = sin(y*div_f), 4f+3 = cos(y*div_f),
// div_f = exp(2f * -(ln 1e4 / (D/2)))   (core/attention.py:243-257)
__global__ void __launch_bounds__(256) spatial_pe_add_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ pos,
                                                             const int* __restrict__ seg_ptr, int num_seg,
                                                             const float* __restrict__ minmax, int N, int D,
                                                             float* __restrict__ out, int ldo) {
  const int nf = D / 4;
  const size_t total = (size_t)N * nf;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / nf), f = (int)(i - (size_t)row * nf);
    int lo = 0, hi = num_seg - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (seg_ptr[mid] <= row) lo = mid; else hi = mid - 1; }
    const float mn = minmax[2 * lo], mx = minmax[2 * lo + 1];
    const float den = (mx - mn) + 1e-8f;
    const float px = (pos[2 * row] - mn) / den, py = (pos[2 * row + 1] - mn) / den;
    const float div = expf((float)(2 * f) * -(9.210340371976184f / (float)(D / 2)));
    const float ax = px * div, ay = py * div;
    float4 x = ldg4(X + (size_t)row * ldx + 4 * f);
    x.x += sinf(ax); x.y += cosf(ax); x.z += sinf(ay); x.w += cosf(ay);
    *reinterpret_cast<float4*>(out + (size_t)row * ldo + 4 * f) = x;
  }
}

// ------------------------------------------------------------------------------------------ un-pool
__global__ void __launch_bounds__(256) unpool_add_act_kernel(const float* __restrict__ skip, int lds, const float* __restrict__ x,
                                                             int ldx, const int* __restrict__ node_map, int N, int C, int act,
                                                             float* __restrict__ out, int ldo) {
  const int c4 = C / 4;
  const size_t total = (size_t)N * c4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / c4), c = (int)(i - (size_t)row * c4) * 4;
    float4 s = ldg4(skip + (size_t)row * lds + c);
    const int j = node_map[row];
    if (j >= 0) { float4 u = ldg4(x + (size_t)j * ldx + c); s.x += u.x; s.y += u.y; s.z += u.z; s.w += u.w; }
    apply_act4(s, act);
    *reinterpret_cast<float4*>(out + (size_t)row * ldo + c) = s;
  }
}

// ------------------------------------------------------------------------------------------ validation
__global__ void __launch_bounds__(256) validate_kernel(const float* __restrict__ x, size_t n_x, const long long* __restrict__ ei,
                                                       size_t n_ei, int N, int* __restrict__ flags, float* __restrict__ stats) {
  float amax = 0.f, s = 0.f, ss = 0.f;
  int fnan = 0, finf = 0, fhi = 0, flo = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto take = [&](float v) {
    if (isnan(v)) fnan = 1; else if (isinf(v)) finf = 1;
    else { amax = fmaxf(amax, fabsf(v)); s += v; ss = fmaf(v, v, ss); }
  };
  // 128-bit loads, four in flight per thread (x is 16-byte aligned: torch allocations are); scalar tail
  const size_t n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? n_x / 4 : 0;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  size_t i = t0;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    const float4 a = __ldg(x4 + i), b = __ldg(x4 + i + stride), c = __ldg(x4 + i + 2 * stride), d = __ldg(x4 + i + 3 * stride);
    take(a.x); take(a.y); take(a.z); take(a.w); take(b.x); take(b.y); take(b.z); take(b.w);
    take(c.x); take(c.y); take(c.z); take(c.w); take(d.x); take(d.y); take(d.z); take(d.w);
  }
  for (; i < n4; i += stride) { const float4 a = __ldg(x4 + i); take(a.x); take(a.y); take(a.z); take(a.w); }
  for (size_t j = 4 * n4 + t0; j < n_x; j += stride) take(x[j]);
  for (size_t i = t0; i < n_ei; i += stride) {
    const long long v = ei[i];
    if (v > (long long)N - 1) fhi = 1;
    if (v < 0) flo = 1;
  }
  amax = warp_max(amax); s = warp_sum(s); ss = warp_sum(ss);
  fnan = __any_sync(0xffffffffu, fnan); finf = __any_sync(0xffffffffu, finf);
  fhi = __any_sync(0xffffffffu, fhi); flo = __any_sync(0xffffffffu, flo);
  if ((threadIdx.x & 31) == 0) {
    if (fnan) atomicOr(flags + 0, 1);
    if (finf) atomicOr(flags + 1, 1);
    if (fhi) atomicOr(flags + 2, 1);
    if (flo) atomicOr(flags + 3, 1);
    atomicMax(reinterpret_cast<unsigned int*>(stats + 0), __float_as_uint(amax));   // amax >= 0: uint order == float order
    atomicAdd(stats + 1, s);        // sum / sum-of-squares only feed a warning threshold; order-insensitive use
    atomicAdd(stats + 2, ss);
  }
}

static inline int grid_for(size_t n, int block = 256, int cap = 148 * 16) {
  size_t b = (n + block - 1) / block;
  return (int)(b < 1 ? 1 : (b > (size_t)cap ? cap : b));
}

}  // namespace dgdm
using namespace dgdm;

// ------------------------------------------------------------------------------------------ contrastive loss (training/losses.py:123-176)
// out[i,:] = x[i,:] / max(|x[i,:]|_2, eps)   (F.normalize, :131); warp per row
__global__ void __launch_bounds__(256) row_normalize_kernel(const float* __restrict__ X, int ldx, int M, int C, float eps,
                                                            float* __restrict__ out, int ldo) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  float sq = 0.f;
  for (int c = lane; c < C; c += 32) { const float v = X[(size_t)row * ldx + c]; sq = fmaf(v, v, sq); }
  sq = warp_sum(sq);
  const float inv = 1.0f / fmaxf(sqrtf(sq), eps);
  for (int c = lane; c < C; c += 32) out[(size_t)row * ldo + c] = X[(size_t)row * ldx + c] * inv;
}
// One chunk of the N x N similarity matrix: S[i, j] = e_i . e_{k0+j} for j < nk (fp32, from the tensor-core Gram GEMM).
// sum_all[i] += sum_j exp((S_ij - 1) / T); sum_pos[i] += the same over keys of the SAME graph, j != i; cnt_pos[i] += their number.
// The shift by 1/T (cosine <= 1) keeps exp() in range for small temperatures; it cancels in the ratio and is added back in the
// logarithms by the caller. Warp per row, fixed lane -> key assignment and one read-modify-write per chunk: deterministic.
__global__ void __launch_bounds__(256) contrastive_accumulate_kernel(const float* __restrict__ S, int lds, int M, int nk, int row0, int k0,
                                                                     const long long* __restrict__ batch, float inv_temp,
                                                                     float* __restrict__ sum_all, float* __restrict__ sum_pos,
                                                                     int* __restrict__ cnt_pos) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const long long b = batch ? batch[row0 + row] : 0;
  float a = 0.f, pz = 0.f; int c = 0;
  for (int j = lane; j < nk; j += 32) {
    const float e = expf((S[(size_t)row * lds + j] - 1.0f) * inv_temp);
    a += e;
    const bool same = (batch ? batch[k0 + j] == b : true) && (k0 + j != row0 + row);
    if (same) { pz += e; ++c; }
  }
  a = warp_sum(a); pz = warp_sum(pz);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) { sum_all[row] += a; sum_pos[row] += pz; cnt_pos[row] += c; }
}

extern "C" int dgdm_row_normalize(const float* X, int ldx, int M, int C, float eps, float* out, int ldo, void* stream) {
  if (!X || !out || C <= 0) return DGDM_ERR_INVALID;
  if (M <= 0) return DGDM_OK;
  row_normalize_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(X, ldx, M, C, eps, out, ldo);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}
extern "C" int dgdm_contrastive_accumulate(const float* S, int lds, int M, int nk, int row0, int k0, const int64_t* batch,
                                           float temperature, float* sum_all, float* sum_pos, int* cnt_pos, void* stream) {
  if (!S || !sum_all || !sum_pos || !cnt_pos || !(temperature > 0.f) || nk < 0) return DGDM_ERR_INVALID;
  if (M <= 0 || nk == 0) return DGDM_OK;
  contrastive_accumulate_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(S, lds, M, nk, row0, k0, (const long long*)batch,
                                                                              1.0f / temperature, sum_all, sum_pos, cnt_pos);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" int dgdm_abi_version(void) { return 1; }

extern "C" int dgdm_device_check(int dev) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return DGDM_ERR_NODEV;
  return prop.major == 10 ? DGDM_OK : DGDM_ERR_NODEV;
}

extern "C" int dgdm_rowwise_epilogue(const float* X, int ldx, int M, int N, const dgdm_gemm_params* ep, void* stream) {
  if (!X || !ep || N <= 0 || (N % 4) != 0 || N > 2048 || (ldx % 4) != 0) return DGDM_ERR_INVALID;
  if (M <= 0) return DGDM_OK;
  const int blocks = (M + 7) / 8;
  cudaStream_t s = (cudaStream_t)stream;
  if (N <= 128) rowwise_epilogue_kernel<1><<<blocks, 256, 0, s>>>(X, ldx, M, N, *ep);
  else if (N <= 256) rowwise_epilogue_kernel<2><<<blocks, 256, 0, s>>>(X, ldx, M, N, *ep);
  else if (N <= 512) rowwise_epilogue_kernel<4><<<blocks, 256, 0, s>>>(X, ldx, M, N, *ep);
  else if (N <= 1024) rowwise_epilogue_kernel<8><<<blocks, 256, 0, s>>>(X, ldx, M, N, *ep);
  else rowwise_epilogue_kernel<16><<<blocks, 256, 0, s>>>(X, ldx, M, N, *ep);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" int dgdm_gemm_simt(const dgdm_gemm_params* p, const float* W32, void* stream) {
  if (!p || !p->A || !W32 || !p->C || p->N <= 0 || p->K <= 0) return DGDM_ERR_INVALID;
  if (p->M <= 0) return DGDM_OK;
  dim3 grid((p->N + 63) / 64, (p->M + 63) / 64);
  gemm_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p->A, p->lda, W32, p->K, p->C, p->ldc, p->M, p->N, p->K, 1.0f);
  DGDM_CHECK_LAUNCH();
  return dgdm_rowwise_epilogue(p->C, p->ldc, p->M, p->N, p, stream);
}

extern "C" int dgdm_groupnorm_silu(const float* X, int ldx, int M, int C, int groups, const float* gamma, const float* beta,
                                   float eps, float* out, int ldo, void* stream) {
  if (!X || !out || !gamma || !beta || groups <= 0 || C <= 0 || (C % groups) != 0) return DGDM_ERR_INVALID;
  if (M <= 0) return DGDM_OK;
  const long long warps = (long long)M * groups;
  const int blocks = (int)((warps + 7) / 8);
  groupnorm_silu_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(X, ldx, M, C, groups, gamma, beta, eps, out, ldo);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" int dgdm_diffusion_add_noise(const float* x0, const float* noise, const float* sqrt_ac, const float* sqrt_1mac,
                                        const int* seg_ptr, const int* seg_t, int num_seg, int N, int C, float* out, void* stream) {
  if (!x0 || !noise || !sqrt_ac || !sqrt_1mac || !seg_ptr || !seg_t || !out || num_seg <= 0) return DGDM_ERR_INVALID;
  if (N <= 0) return DGDM_OK;
  diffusion_add_noise_kernel<<<grid_for((size_t)N * C), 256, 0, (cudaStream_t)stream>>>(x0, noise, sqrt_ac, sqrt_1mac, seg_ptr,
                                                                                        seg_t, num_seg, N, C, out);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" int dgdm_ddpm_update(const float* x, const float* eps_pred, const float* noise, float sqrt_1mac, float inv_sqrt_ac,
                                float sqrt_alpha, float sqrt_var, size_t n, float* out, void* stream) {
  if (!x || !eps_pred || !out) return DGDM_ERR_INVALID;
  if (n == 0) return DGDM_OK;
  ddpm_update_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, eps_pred, noise, sqrt_1mac, inv_sqrt_ac, sqrt_alpha,
                                                                    sqrt_var, n, out);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" int dgdm_segment_minmax(const float* pos, const int* seg_ptr, int num_seg, float* minmax, void* stream) {
  if (!pos || !seg_ptr || !minmax || num_seg <= 0) return DGDM_ERR_INVALID;
  segment_minmax_kernel<<<num_seg, 1024, 0, (cudaStream_t)stream>>>(pos, seg_ptr, minmax);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" int dgdm_spatial_pe_add(const float* X, int ldx, const float* pos, const int* seg_ptr, int num_seg,
                                   const float* minmax, int N, int D, float* out, int ldo, void* stream) {
  if (!X || !pos || !seg_ptr || !minmax || !out || D <= 0 || (D % 4) != 0 || (ldx % 4) != 0 || (ldo % 4) != 0) return DGDM_ERR_INVALID;
  if (N <= 0) return DGDM_OK;
  spatial_pe_add_kernel<<<grid_for((size_t)N * (D / 4)), 256, 0, (cudaStream_t)stream>>>(X, ldx, pos, seg_ptr, num_seg, minmax, N, D, out, ldo);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" int dgdm_unpool_add_act(const float* skip, int ld_skip, const float* x, int ldx, const int* node_map, int N,
                                   int C, int act, float* out, int ldo, void* stream) {
  if (!skip || !x || !node_map || !out || C <= 0 || (C % 4) != 0) return DGDM_ERR_INVALID;
  if (N <= 0) return DGDM_OK;
  unpool_add_act_kernel<<<grid_for((size_t)N * (C / 4)), 256, 0, (cudaStream_t)stream>>>(skip, ld_skip, x, ldx, node_map, N, C, act, out, ldo);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" int dgdm_validate_inputs(const float* x, size_t n_x, const int64_t* edge_index, size_t n_ei, int N,
                                    int* flags4, float* stats4, void* stream) {
  if (!flags4 || !stats4) return DGDM_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(flags4, 0, 4 * sizeof(int), s) != cudaSuccess) return DGDM_ERR_LAUNCH;
  if (cudaMemsetAsync(stats4, 0, 4 * sizeof(float), s) != cudaSuccess) return DGDM_ERR_LAUNCH;
  const size_t n = n_x > n_ei ? n_x : n_ei;
  if (n == 0) return DGDM_OK;
  validate_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, s>>>(x, n_x, (const long long*)edge_index, n_ei, N, flags4, stats4);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}
