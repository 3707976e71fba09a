// Backward of the spatial attention core (the training step of models/dgdm_model.py:405-506 needs
// d(loss)/d(q,k,v) of   O_i = sum_j softmax_j(q_i.k_j / sqrt(hd) - |p_i - p_j| / T) v_j   per head, keys of the query's graph).
// Flash-style: the N x N weights are never stored; three passes recompute the score tiles from q, k and the positions:
//   stats : lse_i = log sum_j exp(s_ij)            and   delta_i = O_i . dO_i
//   dkdv  : dV_j = sum_i P_ij dO_i ,  dK_j = scale * sum_i dS_ij q_i           (one CTA per 64-key tile and head, loops over queries)
//   dq    : dQ_i = scale * sum_j dS_ij k_j                                      (one CTA per 64-query tile and head, loops over keys)
// with P_ij = exp(s_ij - lse_i), dP_ij = dO_i . v_j, dS_ij = P_ij (dP_ij - delta_i). Every output element is produced by exactly
// one CTA in a fixed order: deterministic, no atomics. Exact fp32 on the CUDA cores (first backward; the tensor-core version of
// these three passes is the obvious next step — the forward kernel is 17x faster per score). Head width 16 (DGDM-Base).
#include "common.cuh"

namespace dgdm {

constexpr int BT = 64;            // queries / keys per tile
constexpr int BHD = 16;           // head width
constexpr int BLD = 17;           // padded smem row stride

struct AttnBwdArgs {
  const float* qkv; int ld;       // [N, 3D]  q | k | v
  const float* pos;               // [N, 2] or NULL
  const int* seg_of;              // [N] segment (graph) id per row
  const int* seg_ptr;             // [num_seg + 1]
  int N, D;
  float scale, inv_temp;
  const float* O; int ldo;
  const float* dO; int lddo;
  float* lse; float* delta;       // [N, heads]
  float* dqkv; int ldd;           // [N, 3D]
};

__device__ __forceinline__ void load_tile(float* dst, const float* src, int ld, int row0, int N, int col0, int t) {
  const int r = t >> 2, c = (t & 3) * 4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row0 + r < N) v = ldg4(src + (size_t)(row0 + r) * ld + col0 + c);
  dst[r * BLD + c] = v.x; dst[r * BLD + c + 1] = v.y; dst[r * BLD + c + 2] = v.z; dst[r * BLD + c + 3] = v.w;
}
__device__ __forceinline__ void load_meta(float* px, float* py, int* sg, const float* pos, const int* seg_of, int row0, int N, int t) {
  if (t < BT) {
    const int r = row0 + t;
    px[t] = (pos && r < N) ? pos[2 * (size_t)r] : 0.f;
    py[t] = (pos && r < N) ? pos[2 * (size_t)r + 1] : 0.f;
    sg[t] = r < N ? seg_of[r] : -1 - t;           // rows past the end match nothing
  }
}
// s[r][c] for this thread's 4 queries x 4 keys; invalid pairs (other graph / past the end) get -inf
__device__ __forceinline__ void score_tile(const float* Qs, const float* Ks, const float* qx, const float* qy, const int* qs,
                                           const float* kx, const float* ky, const int* ks, int tq, int tk, float scale,
                                           float inv_temp, bool use_pos, float s[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) s[r][c] = 0.f;
#pragma unroll
  for (int d = 0; d < BHD; ++d) {
    float q[4], k[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) q[r] = Qs[(tq * 4 + r) * BLD + d];
#pragma unroll
    for (int c = 0; c < 4; ++c) k[c] = Ks[(tk + 16 * c) * BLD + d];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[r][c] = fmaf(q[r], k[c], s[r][c]);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int qi = tq * 4 + r, ki = tk + 16 * c;
      float v = s[r][c] * scale;
      if (use_pos) { const float dx = qx[qi] - kx[ki], dy = qy[qi] - ky[ki]; v -= sqrtf(dx * dx + dy * dy) * inv_temp; }
      s[r][c] = (qs[qi] == ks[ki]) ? v : -INFINITY;
    }
}
__device__ __forceinline__ void tile_range(const AttnBwdArgs& a, int row0, int& lo, int& hi) {
  const int s0 = a.seg_of[row0], s1 = a.seg_of[min(row0 + BT - 1, a.N - 1)];
  lo = (a.seg_ptr[s0] / BT) * BT;
  hi = a.seg_ptr[s1 + 1];
}

// dP[r][c] = dO_r . v_c for this thread's 4 x 4 pairs (d outermost: 8 shared-memory reads per 16 FMAs)
__device__ __forceinline__ void dp_tile(const float* dOs, const float* Vs, int tq, int tk, float dp[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) dp[r][c] = 0.f;
#pragma unroll
  for (int d = 0; d < BHD; ++d) {
    float g[4], v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) g[r] = dOs[(tq * 4 + r) * BLD + d];
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = Vs[(tk + 16 * c) * BLD + d];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) dp[r][c] = fmaf(g[r], v[c], dp[r][c]);
  }
}

struct BwdSmem {
  float A[BT * BLD], B[BT * BLD], C[BT * BLD], Dd[BT * BLD];
  float ax[BT], ay[BT], bx[BT], by[BT], v0[BT], v1[BT];
  int as[BT], bs[BT];
};

// ---- pass 1: lse and delta. grid (query tiles, heads)
__global__ void __launch_bounds__(256) attn_bwd_stats_kernel(const AttnBwdArgs a) {
  __shared__ BwdSmem sm;
  const int t = threadIdx.x, tq = t >> 4, tk = t & 15, h = blockIdx.y, q0 = blockIdx.x * BT;
  const int heads = a.D / BHD;
  load_tile(sm.A, a.qkv, a.ld, q0, a.N, h * BHD, t);
  load_meta(sm.ax, sm.ay, sm.as, a.pos, a.seg_of, q0, a.N, t);
  float m[4], l[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { m[r] = -INFINITY; l[r] = 0.f; }
  int lo, hi;
  tile_range(a, q0, lo, hi);
  for (int k0 = lo; k0 < hi; k0 += BT) {
    __syncthreads();
    load_tile(sm.B, a.qkv, a.ld, k0, a.N, a.D + h * BHD, t);
    load_meta(sm.bx, sm.by, sm.bs, a.pos, a.seg_of, k0, a.N, t);
    __syncthreads();
    float s[4][4];
    score_tile(sm.A, sm.B, sm.ax, sm.ay, sm.as, sm.bx, sm.by, sm.bs, tq, tk, a.scale, a.inv_temp, a.pos != nullptr, s);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float mx = fmaxf(fmaxf(s[r][0], s[r][1]), fmaxf(s[r][2], s[r][3]));
      if (mx > m[r]) { l[r] *= expf(m[r] - mx); m[r] = mx; }
      if (m[r] > -INFINITY) {
#pragma unroll
        for (int c = 0; c < 4; ++c) l[r] += expf(s[r][c] - m[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {               // merge the 16 key-lanes of this query (half-warp)
      const float mo = __shfl_xor_sync(0xffffffffu, m[r], o), lo_ = __shfl_xor_sync(0xffffffffu, l[r], o);
      const float mn = fmaxf(m[r], mo);
      l[r] = (m[r] > -INFINITY ? l[r] * expf(m[r] - mn) : 0.f) + (mo > -INFINITY ? lo_ * expf(mo - mn) : 0.f);
      m[r] = mn;
    }
    const int row = q0 + tq * 4 + r;
    if (tk == 0 && row < a.N) a.lse[(size_t)row * heads + h] = m[r] + logf(l[r]);
  }
  if (t < BT && q0 + t < a.N) {
    const float* o = a.O + (size_t)(q0 + t) * a.ldo + h * BHD;
    const float* g = a.dO + (size_t)(q0 + t) * a.lddo + h * BHD;
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < BHD; ++j) d = fmaf(o[j], g[j], d);
    a.delta[(size_t)(q0 + t) * heads + h] = d;
  }
}

// ---- pass 2: dQ. grid (query tiles, heads)
__global__ void __launch_bounds__(256) attn_bwd_dq_kernel(const AttnBwdArgs a) {
  __shared__ BwdSmem sm;
  const int t = threadIdx.x, tq = t >> 4, tk = t & 15, h = blockIdx.y, q0 = blockIdx.x * BT;
  const int heads = a.D / BHD;
  load_tile(sm.A, a.qkv, a.ld, q0, a.N, h * BHD, t);                 // Q
  load_tile(sm.C, a.dO, a.lddo, q0, a.N, h * BHD, t);                // dO
  load_meta(sm.ax, sm.ay, sm.as, a.pos, a.seg_of, q0, a.N, t);
  if (t < BT) {
    const int row = q0 + t;
    sm.v0[t] = row < a.N ? a.lse[(size_t)row * heads + h] : 0.f;
    sm.v1[t] = row < a.N ? a.delta[(size_t)row * heads + h] : 0.f;
  }
  float acc[4][BHD];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int d = 0; d < BHD; ++d) acc[r][d] = 0.f;
  int lo, hi;
  tile_range(a, q0, lo, hi);
  for (int k0 = lo; k0 < hi; k0 += BT) {
    __syncthreads();
    load_tile(sm.B, a.qkv, a.ld, k0, a.N, a.D + h * BHD, t);         // K
    load_tile(sm.Dd, a.qkv, a.ld, k0, a.N, 2 * a.D + h * BHD, t);    // V
    load_meta(sm.bx, sm.by, sm.bs, a.pos, a.seg_of, k0, a.N, t);
    __syncthreads();
    float s[4][4], dp[4][4];
    score_tile(sm.A, sm.B, sm.ax, sm.ay, sm.as, sm.bx, sm.by, sm.bs, tq, tk, a.scale, a.inv_temp, a.pos != nullptr, s);
    dp_tile(sm.C, sm.Dd, tq, tk, dp);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int qi = tq * 4 + r;
#pragma unroll
      for (int c = 0; c < 4; ++c) s[r][c] = expf(s[r][c] - sm.v0[qi]) * (dp[r][c] - sm.v1[qi]);     // dS; exp(-inf) = 0 for masked pairs
    }
#pragma unroll
    for (int d = 0; d < BHD; ++d) {
      float k[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) k[c] = sm.B[(tk + 16 * c) * BLD + d];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][d] = fmaf(s[r][c], k[c], acc[r][d]);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = q0 + tq * 4 + r;
#pragma unroll
    for (int d = 0; d < BHD; ++d) {
      float v = acc[r][d];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (tk == 0 && row < a.N) a.dqkv[(size_t)row * a.ldd + h * BHD + d] = v * a.scale;
    }
  }
}

// ---- pass 3: dK, dV. grid (key tiles, heads)
__global__ void __launch_bounds__(256) attn_bwd_dkdv_kernel(const AttnBwdArgs a) {
  __shared__ union { BwdSmem sm; float red[8 * 16 * 4 * BHD]; } u;     // tiles while looping, cross-warp reduction afterwards (32 KB)
  BwdSmem& sm = u.sm;
  const int t = threadIdx.x, tq = t >> 4, tk = t & 15, h = blockIdx.y, k0 = blockIdx.x * BT;
  const int heads = a.D / BHD;
  load_tile(sm.B, a.qkv, a.ld, k0, a.N, a.D + h * BHD, t);           // K
  load_tile(sm.Dd, a.qkv, a.ld, k0, a.N, 2 * a.D + h * BHD, t);      // V
  load_meta(sm.bx, sm.by, sm.bs, a.pos, a.seg_of, k0, a.N, t);
  float dv[4][BHD], dk[4][BHD];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int d = 0; d < BHD; ++d) { dv[c][d] = 0.f; dk[c][d] = 0.f; }
  int lo, hi;
  tile_range(a, k0, lo, hi);
  for (int q0 = lo; q0 < hi; q0 += BT) {
    __syncthreads();
    load_tile(sm.A, a.qkv, a.ld, q0, a.N, h * BHD, t);               // Q
    load_tile(sm.C, a.dO, a.lddo, q0, a.N, h * BHD, t);              // dO
    load_meta(sm.ax, sm.ay, sm.as, a.pos, a.seg_of, q0, a.N, t);
    if (t < BT) {
      const int row = q0 + t;
      sm.v0[t] = row < a.N ? a.lse[(size_t)row * heads + h] : 0.f;
      sm.v1[t] = row < a.N ? a.delta[(size_t)row * heads + h] : 0.f;
    }
    __syncthreads();
    float s[4][4], dp[4][4];
    score_tile(sm.A, sm.B, sm.ax, sm.ay, sm.as, sm.bx, sm.by, sm.bs, tq, tk, a.scale, a.inv_temp, a.pos != nullptr, s);
    dp_tile(sm.C, sm.Dd, tq, tk, dp);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int qi = tq * 4 + r;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s[r][c] = expf(s[r][c] - sm.v0[qi]);                         // P
        dp[r][c] = s[r][c] * (dp[r][c] - sm.v1[qi]);                 // dS
      }
    }
#pragma unroll
    for (int d = 0; d < BHD; ++d) {
      float g[4], q[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) { g[r] = sm.C[(tq * 4 + r) * BLD + d]; q[r] = sm.A[(tq * 4 + r) * BLD + d]; }
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          dv[c][d] = fmaf(s[r][c], g[r], dv[c][d]);
          dk[c][d] = fmaf(dp[r][c], q[r], dk[c][d]);
        }
    }
  }
  // reduce over the 16 query-lanes that share this thread's keys: lane pair (xor 16) in registers, the 8 warps through smem
  const int warp = t >> 5;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int d = 0; d < BHD; ++d) {
        float v = pass == 0 ? dv[c][d] : dk[c][d];
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if ((t & 16) == 0) u.red[((warp * 16 + tk) * 4 + c) * BHD + d] = v;
      }
    __syncthreads();
    for (int o = t; o < BT * BHD; o += 256) {
      const int ki = o / BHD, d = o - ki * BHD;                       // key ki = tk + 16 c
      const int c = ki >> 4, kk = ki & 15;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += u.red[((w * 16 + kk) * 4 + c) * BHD + d];
      const int row = k0 + ki;
      if (row < a.N) a.dqkv[(size_t)row * a.ldd + (pass == 0 ? 2 * a.D : a.D) + h * BHD + d] = pass == 0 ? v : v * a.scale;
    }
  }
}

}  // namespace dgdm
using namespace dgdm;

extern "C" int dgdm_spatial_attention_backward(const float* qkv, int ld, const float* pos, const int* seg_of, const int* seg_ptr,
                                               int N, int D, int heads, float temperature, const float* O, int ldo,
                                               const float* dO, int lddo, float* lse, float* delta, float* dqkv, int ldd,
                                               void* stream) {
  if (!qkv || !seg_of || !seg_ptr || !O || !dO || !lse || !delta || !dqkv) return DGDM_ERR_INVALID;
  if (heads <= 0 || D != heads * BHD || !(temperature > 0.f) || (ld % 4) || (ldo % 4) || (lddo % 4)) return DGDM_ERR_INVALID;
  if (N <= 0) return DGDM_OK;
  AttnBwdArgs a;
  a.qkv = qkv; a.ld = ld; a.pos = pos; a.seg_of = seg_of; a.seg_ptr = seg_ptr; a.N = N; a.D = D;
  a.scale = 1.0f / sqrtf((float)BHD); a.inv_temp = 1.0f / temperature;
  a.O = O; a.ldo = ldo; a.dO = dO; a.lddo = lddo; a.lse = lse; a.delta = delta; a.dqkv = dqkv; a.ldd = ldd;
  const dim3 grid((N + BT - 1) / BT, heads);
  cudaStream_t s = (cudaStream_t)stream;
  attn_bwd_stats_kernel<<<grid, 256, 0, s>>>(a);
  attn_bwd_dkdv_kernel<<<grid, 256, 0, s>>>(a);
  attn_bwd_dq_kernel<<<grid, 256, 0, s>>>(a);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}
