// Generic exact-fp32 flash attention for head dimensions the tensor-core kernels do not cover (DGDM-Large: 16 heads x 8;
// MultiHeadAttention with other widths): same contract as the head_dim-16 kernels — per head
//   out_i = softmax_j( q_i.k_j / sqrt(hd) - |p_i - p_j| / T ) v_j,   keys restricted to the query's graph segment,
// never forming the N x N scores. CUDA-core FMAs, thread = (query row, head), 64-key tiles staged in shared memory
// (every lane reads the same key -> broadcast, conflict-free), online softmax per tile. FMA/MUFU-bound; it is the coverage
// path, not the benchmarked one (core/attention.py:135-157,261-283).
#include "common.cuh"

namespace dgdm {

constexpr int SA_ROWS = 128;
constexpr int SA_KT = 64;

template <int HD>
__global__ void __launch_bounds__(SA_ROWS) attn_simt_kernel(const float* __restrict__ q_qkv, int ld_q, const float* __restrict__ q_pos,
                                                            const int* __restrict__ qseg, const float* __restrict__ kv_qkv, int ld_kv,
                                                            const float* __restrict__ kv_pos, const int* __restrict__ kseg, int num_seg,
                                                            int D, float qscale, float pscale, float* __restrict__ out, int ldo) {
  __shared__ __align__(16) float Ks[SA_KT][HD];
  __shared__ __align__(16) float Vs[SA_KT][HD];
  __shared__ float kx[SA_KT], ky[SA_KT];
  int seg = -1, row0 = 0;
  {
    int acc = 0;
    for (int s = 0; s < num_seg; ++s) {
      const int a = qseg[s], b = qseg[s + 1];
      const int nt = (b - a + SA_ROWS - 1) / SA_ROWS;
      if ((int)blockIdx.x < acc + nt) { seg = s; row0 = a + ((int)blockIdx.x - acc) * SA_ROWS; break; }
      acc += nt;
    }
  }
  if (seg < 0) return;
  const int h = blockIdx.y, tid = threadIdx.x;
  const int row = row0 + tid;
  const bool valid = row < qseg[seg + 1];
  float q[HD], acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) { q[d] = valid ? q_qkv[(size_t)row * ld_q + h * HD + d] * qscale : 0.f; acc[d] = 0.f; }
  const bool bias = q_pos != nullptr;
  const float qx = (bias && valid) ? q_pos[2 * (size_t)row] * pscale : 0.f, qy = (bias && valid) ? q_pos[2 * (size_t)row + 1] * pscale : 0.f;
  float m = -INFINITY, l = 0.f;
  const int k_beg = kseg[seg], k_end = kseg[seg + 1];
  for (int k0 = k_beg; k0 < k_end; k0 += SA_KT) {
    const int nk = min(SA_KT, k_end - k0);
    __syncthreads();
    for (int i = tid; i < SA_KT * HD; i += SA_ROWS) {
      const int r = i / HD, d = i - r * HD;
      const bool ok = r < nk;
      Ks[r][d] = ok ? kv_qkv[(size_t)(k0 + r) * ld_kv + D + h * HD + d] : 0.f;
      Vs[r][d] = ok ? kv_qkv[(size_t)(k0 + r) * ld_kv + 2 * D + h * HD + d] : 0.f;
    }
    if (tid < SA_KT) {
      const bool ok = bias && tid < nk;
      kx[tid] = ok ? kv_pos[2 * (size_t)(k0 + tid)] * pscale : 0.f;
      ky[tid] = ok ? kv_pos[2 * (size_t)(k0 + tid) + 1] * pscale : 0.f;
    }
    __syncthreads();
    float s[SA_KT];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < SA_KT; ++j) {
      float v = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) v = fmaf(q[d], Ks[j][d], v);
      if (bias) { const float dx = qx - kx[j], dy = qy - ky[j]; v -= sqrtf(dx * dx + dy * dy); }
      v = (j < nk) ? v : -INFINITY;
      s[j] = v;
      tmax = fmaxf(tmax, v);
    }
    const float mn = fmaxf(m, tmax);
    const float corr = exp2f(m - mn);            // m = -inf on the first tile -> 0
    l *= corr;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] *= corr;
#pragma unroll
    for (int j = 0; j < SA_KT; ++j) {
      const float pj = exp2f(s[j] - mn);         // masked keys: exp2(-inf) = 0
      l += pj;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] = fmaf(pj, Vs[j][d], acc[d]);
    }
    m = mn;
  }
  if (valid) {
    const float inv = l > 0.f ? 1.0f / l : 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) out[(size_t)row * ldo + h * HD + d] = acc[d] * inv;
  }
}

int attn_simt_launch(const float* q_qkv, int ld_q, const float* q_pos, const int* qseg_ptr, int Nq,
                     const float* kv_qkv, int ld_kv, const float* kv_pos, const int* kseg_ptr, int num_seg,
                     int D, int heads, float temperature, float* out, int ldo, cudaStream_t s) {
  const int hd = D / heads;
  const float LOG2E = 1.4426950408889634f;
  const float qscale = LOG2E / sqrtf((float)hd), pscale = LOG2E / temperature;
  dim3 grid((Nq + SA_ROWS - 1) / SA_ROWS + num_seg, heads);
#define SA_LAUNCH(H) attn_simt_kernel<H><<<grid, SA_ROWS, 0, s>>>(q_qkv, ld_q, q_pos, qseg_ptr, kv_qkv, ld_kv, kv_pos, kseg_ptr, num_seg, D, qscale, pscale, out, ldo)
  switch (hd) {
    case 4: SA_LAUNCH(4); break;
    case 8: SA_LAUNCH(8); break;
    case 16: SA_LAUNCH(16); break;
    case 32: SA_LAUNCH(32); break;
    case 64: SA_LAUNCH(64); break;
    default: return DGDM_ERR_INVALID;
  }
#undef SA_LAUNCH
  if (cudaGetLastError() != cudaSuccess) return DGDM_ERR_LAUNCH;
  return DGDM_OK;
}

}  // namespace dgdm
