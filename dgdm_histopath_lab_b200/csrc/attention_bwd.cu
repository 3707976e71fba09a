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

// have_lse: lse already holds the forward's log-sum-exp (dgdm_spatial_attention_lse) -> the statistics pass only forms delta
__global__ void __launch_bounds__(256) attn_bwd_delta_kernel(const float* __restrict__ O, int ldo, const float* __restrict__ dO, int lddo,
                                                             int N, int heads, float* __restrict__ delta) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * heads) return;
  const int row = (int)(i / heads), h = (int)(i - (size_t)row * heads);
  const float* o = O + (size_t)row * ldo + h * dgdm::BHD;
  const float* g = dO + (size_t)row * lddo + h * dgdm::BHD;
  float d = 0.f;
#pragma unroll
  for (int j = 0; j < dgdm::BHD; j += 4) {
    const float4 a = dgdm::ldg4(o + j), b = dgdm::ldg4(g + j);
    d = fmaf(a.x, b.x, d); d = fmaf(a.y, b.y, d); d = fmaf(a.z, b.z, d); d = fmaf(a.w, b.w, d);
  }
  delta[i] = d;
}

extern "C" int dgdm_spatial_attention_backward(const float* qkv, int ld, const float* pos, const int* seg_of, const int* seg_ptr,
                                               int N, int D, int heads, float temperature, const float* O, int ldo,
                                               const float* dO, int lddo, int have_lse, float* lse, float* delta, float* dqkv, int ldd,
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
  if (have_lse) attn_bwd_delta_kernel<<<(unsigned)(((size_t)N * heads + 255) / 256), 256, 0, s>>>(O, ldo, dO, lddo, N, heads, delta);
  else attn_bwd_stats_kernel<<<grid, 256, 0, s>>>(a);
  attn_bwd_dkdv_kernel<<<grid, 256, 0, s>>>(a);
  attn_bwd_dq_kernel<<<grid, 256, 0, s>>>(a);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

// =====================================================================================================================
// Tensor-core version of the three passes (default; DGDM_ATTN_BWD=simt selects the exact fp32 kernels above).
// Warp-level mma.sync m16n8k16 (fp16 operands, fp32 accumulate): a warp owns 16 rows of the CTA's 64-row tile and walks the
// 64-row tiles of the other side. The score tile is produced directly in the orientation the next product needs
// (S^T = K Q^T in the dK/dV pass, S = Q K^T in the dQ / stats passes), so the accumulator fragments of two adjacent 8-column tiles
// ARE the A fragment of the following product (P^T dO, dS^T Q, dS K) — no shuffles, no shared-memory round trip.
// Precision: single fp16 operand planes (q, k, v, dO, P, dS), fp32 accumulation, dO pre-scaled by a power of two (go_scale) so
// that gradients of ~1e-6 sit in the normal fp16 range. Gradients are within ~2e-3 of the fp32 path (stated tensor-core
// tolerance; the forward keeps its fp32-class split arithmetic because the top-k pooling downstream needs it, a gradient does not).
namespace dgdm {

constexpr int MT = 64;             // rows per CTA tile (4 warps x 16)
constexpr int TS = 72;             // half stride of the transposed tiles (144 B: conflict-free fragment loads)

__device__ __forceinline__ void mma16816(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

struct TcSmem {
  __half rowm[2][MT * 16];          // two row-major [64][16] tiles (B operands with k = head dim)
  __half tr[2][16 * TS];            // two transposed [16][64(+8)] tiles (B operands with k = rows of the other side)
  float4 meta[MT];                  // per row of the streamed tile: (x, y, lse in log2 units, delta) — one 128-bit load per row
  int sg[MT];
};
constexpr float LOG2E_F = 1.4426950408889634f;
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// fp32 [64 x 16] tile (rows row0.., columns col0..) -> half, row-major and/or transposed; `mul` scales the values
__device__ __forceinline__ void tc_load(const float* src, int ld, int row0, int N, int col0, float mul, __half* rm, __half* tr, int t) {
  for (int i = t; i < MT * 4; i += 128) {
    const int r = i >> 2, c = (i & 3) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < N) v = ldg4(src + (size_t)(row0 + r) * ld + col0 + c);
    const __half h0 = __float2half_rn(v.x * mul), h1 = __float2half_rn(v.y * mul), h2 = __float2half_rn(v.z * mul), h3 = __float2half_rn(v.w * mul);
    if (rm) { rm[r * 16 + c] = h0; rm[r * 16 + c + 1] = h1; rm[r * 16 + c + 2] = h2; rm[r * 16 + c + 3] = h3; }
    if (tr) { tr[c * TS + r] = h0; tr[(c + 1) * TS + r] = h1; tr[(c + 2) * TS + r] = h2; tr[(c + 3) * TS + r] = h3; }
  }
}
// A fragment (16 rows x 16 dims) of this warp's rows straight from global fp32
__device__ __forceinline__ void tc_afrag(const float* src, int ld, int row0, int N, int col0, float mul, int g, int tq, uint32_t a[4]) {
  auto ld2 = [&](int r, int c) {
    float2 v = make_float2(0.f, 0.f);
    if (row0 + r < N) v = *reinterpret_cast<const float2*>(src + (size_t)(row0 + r) * ld + col0 + c);
    return pack_half2(v.x * mul, v.y * mul);
  };
  a[0] = ld2(g, 2 * tq); a[1] = ld2(g + 8, 2 * tq); a[2] = ld2(g, 2 * tq + 8); a[3] = ld2(g + 8, 2 * tq + 8);
}

// ---- dK / dV: CTA = (key tile, head); warp = 16 keys; loop over query tiles
__global__ void __launch_bounds__(128) attn_bwd_dkdv_tc_kernel(const AttnBwdArgs a, const float* __restrict__ go_scale) {
  __shared__ TcSmem sm;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, g = lane >> 2, tq = lane & 3;
  const int h = blockIdx.y, k0 = blockIdx.x * MT, heads = a.D / BHD;
  const float gs = *go_scale;
  const int kr0 = k0 + warp * 16;                                   // this warp's keys: kr0 + g, kr0 + g + 8
  uint32_t kf[4], vf[4];
  tc_afrag(a.qkv, a.ld, kr0, a.N, a.D + h * BHD, 1.0f, g, tq, kf);
  tc_afrag(a.qkv, a.ld, kr0, a.N, 2 * a.D + h * BHD, 1.0f, g, tq, vf);
  float kx[2], ky[2]; int ks[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int key = kr0 + g + 8 * r;
    kx[r] = (a.pos && key < a.N) ? a.pos[2 * (size_t)key] : 0.f;
    ky[r] = (a.pos && key < a.N) ? a.pos[2 * (size_t)key + 1] : 0.f;
    ks[r] = key < a.N ? a.seg_of[key] : -1000000 - key;
  }
  float dv[2][4], dk[2][4];
#pragma unroll
  for (int n = 0; n < 2; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) { dv[n][i] = 0.f; dk[n][i] = 0.f; }
  int lo, hi;
  tile_range(a, k0, lo, hi);
  const bool use_pos = a.pos != nullptr;
  const float sc2 = a.scale * LOG2E_F, it2 = a.inv_temp * LOG2E_F;
  for (int q0 = lo; q0 < hi; q0 += MT) {
    __syncthreads();
    tc_load(a.qkv, a.ld, q0, a.N, h * BHD, 1.0f, sm.rowm[0], sm.tr[0], t);          // Q, Q^T
    tc_load(a.dO, a.lddo, q0, a.N, h * BHD, gs, sm.rowm[1], sm.tr[1], t);            // dO, dO^T (scaled)
    if (t < MT) {
      const int row = q0 + t;
      const bool ok = row < a.N;
      sm.meta[t] = make_float4((use_pos && ok) ? a.pos[2 * (size_t)row] : 0.f, (use_pos && ok) ? a.pos[2 * (size_t)row + 1] : 0.f,
                               ok ? a.lse[(size_t)row * heads + h] * LOG2E_F : 0.f, ok ? a.delta[(size_t)row * heads + h] * gs : 0.f);
      sm.sg[t] = ok ? a.seg_of[row] : -1 - t;
    }
    __syncthreads();
#pragma unroll
    for (int kstep = 0; kstep < 4; ++kstep) {                         // 16 queries per step = two 8-query tiles
      uint32_t pf[4], sf[4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int qb = kstep * 16 + half * 8;                         // B fragment: queries qb + g, dims 2 tq (+8)
        const uint32_t* qrow = reinterpret_cast<const uint32_t*>(sm.rowm[0] + (qb + g) * 16);
        const uint32_t* grow = reinterpret_cast<const uint32_t*>(sm.rowm[1] + (qb + g) * 16);
        float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816(s, kf, qrow[tq], qrow[tq + 4]);                      // S^T[key, query]
        mma16816(dp, vf, grow[tq], grow[tq + 4]);                     // dP^T[key, query] = v_key . dO_query
        float p[4], ds[4];
        const float4 m0 = sm.meta[qb + 2 * tq], m1 = sm.meta[qb + 2 * tq + 1];      // the two queries of this lane's columns
        const int sg0 = sm.sg[qb + 2 * tq], sg1 = sm.sg[qb + 2 * tq + 1];
#pragma unroll
        for (int i = 0; i < 4; ++i) {                                 // i: (key g + 8 (i>>1), query qb + 2 tq + (i&1))
          const int r = i >> 1;
          const float4 mq = (i & 1) ? m1 : m0;
          float v = fmaf(s[i], sc2, -mq.z);                           // log2 units
          if (use_pos) { const float dx = mq.x - kx[r], dy = mq.y - ky[r]; v = fmaf(sqrtf(fmaf(dx, dx, dy * dy)), -it2, v); }
          p[i] = (((i & 1) ? sg1 : sg0) == ks[r]) ? ex2f(v) : 0.f;
          ds[i] = p[i] * (dp[i] - mq.w);
        }
        pf[2 * half] = pack_half2(p[0], p[1]); pf[2 * half + 1] = pack_half2(p[2], p[3]);
        sf[2 * half] = pack_half2(ds[0], ds[1]); sf[2 * half + 1] = pack_half2(ds[2], ds[3]);
      }
      // A fragment order is (row g, k lo), (row g+8, k lo), (row g, k hi), (row g+8, k hi): exactly pf / sf
#pragma unroll
      for (int n = 0; n < 2; ++n) {                                   // dims n*8 + g
        const uint32_t* gt = reinterpret_cast<const uint32_t*>(sm.tr[1] + (n * 8 + g) * TS + kstep * 16);
        const uint32_t* qt = reinterpret_cast<const uint32_t*>(sm.tr[0] + (n * 8 + g) * TS + kstep * 16);
        mma16816(dv[n], pf, gt[tq], gt[tq + 4]);                      // dV[key, dim] += P^T[key, q] dO[q, dim]
        mma16816(dk[n], sf, qt[tq], qt[tq + 4]);                      // dK[key, dim] += dS^T[key, q] Q[q, dim]
      }
    }
  }
  const float inv = 1.0f / gs;
#pragma unroll
  for (int n = 0; n < 2; ++n)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int key = kr0 + g + 8 * r;
      if (key < a.N) {
        float* o = a.dqkv + (size_t)key * a.ldd + h * BHD + n * 8 + 2 * tq;
        *reinterpret_cast<float2*>(o + 2 * a.D) = make_float2(dv[n][2 * r] * inv, dv[n][2 * r + 1] * inv);
        *reinterpret_cast<float2*>(o + a.D) = make_float2(dk[n][2 * r] * inv * a.scale, dk[n][2 * r + 1] * inv * a.scale);
      }
    }
}

// ---- dQ (MODE 1) and the softmax statistics (MODE 0): CTA = (query tile, head); warp = 16 queries; loop over key tiles
template <int MODE>
__global__ void __launch_bounds__(128) attn_bwd_q_tc_kernel(const AttnBwdArgs a, const float* __restrict__ go_scale) {
  __shared__ TcSmem sm;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, g = lane >> 2, tq = lane & 3;
  const int h = blockIdx.y, q0 = blockIdx.x * MT, heads = a.D / BHD;
  const float gs = MODE ? *go_scale : 1.0f;
  const int qr0 = q0 + warp * 16;
  uint32_t qf[4], gf[4];
  tc_afrag(a.qkv, a.ld, qr0, a.N, h * BHD, 1.0f, g, tq, qf);
  if (MODE) tc_afrag(a.dO, a.lddo, qr0, a.N, h * BHD, gs, g, tq, gf);
  float qx[2], qy[2], lse[2], dl[2]; int qs[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = qr0 + g + 8 * r;
    qx[r] = (a.pos && row < a.N) ? a.pos[2 * (size_t)row] : 0.f;
    qy[r] = (a.pos && row < a.N) ? a.pos[2 * (size_t)row + 1] : 0.f;
    qs[r] = row < a.N ? a.seg_of[row] : -1000000 - row;
    lse[r] = (MODE && row < a.N) ? a.lse[(size_t)row * heads + h] * LOG2E_F : 0.f;
    dl[r] = (MODE && row < a.N) ? a.delta[(size_t)row * heads + h] * gs : 0.f;
  }
  float dq[2][4], m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
#pragma unroll
  for (int n = 0; n < 2; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) dq[n][i] = 0.f;
  int lo, hi;
  tile_range(a, q0, lo, hi);
  const bool use_pos = a.pos != nullptr;
  const float sc2 = a.scale * LOG2E_F, it2 = a.inv_temp * LOG2E_F;
  for (int k0 = lo; k0 < hi; k0 += MT) {
    __syncthreads();
    tc_load(a.qkv, a.ld, k0, a.N, a.D + h * BHD, 1.0f, sm.rowm[0], MODE ? sm.tr[0] : nullptr, t);       // K (, K^T)
    if (MODE) tc_load(a.qkv, a.ld, k0, a.N, 2 * a.D + h * BHD, 1.0f, sm.rowm[1], nullptr, t);           // V
    if (t < MT) {
      const int row = k0 + t;
      const bool ok = row < a.N;
      sm.meta[t] = make_float4((use_pos && ok) ? a.pos[2 * (size_t)row] : 0.f, (use_pos && ok) ? a.pos[2 * (size_t)row + 1] : 0.f, 0.f, 0.f);
      sm.sg[t] = ok ? a.seg_of[row] : -1 - t;
    }
    __syncthreads();
#pragma unroll
    for (int kstep = 0; kstep < 4; ++kstep) {
      uint32_t sf[4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int kb = kstep * 16 + half * 8;
        const uint32_t* krow = reinterpret_cast<const uint32_t*>(sm.rowm[0] + (kb + g) * 16);
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816(s, qf, krow[tq], krow[tq + 4]);                      // S[query, key]
        float sv[4];                                                  // scores in log2 units
        const float4 m0 = sm.meta[kb + 2 * tq], m1 = sm.meta[kb + 2 * tq + 1];
        const int sg0 = sm.sg[kb + 2 * tq], sg1 = sm.sg[kb + 2 * tq + 1];
#pragma unroll
        for (int i = 0; i < 4; ++i) {                                 // i: (query g + 8 (i>>1), key kb + 2 tq + (i&1))
          const int r = i >> 1;
          const float4 mk = (i & 1) ? m1 : m0;
          float v = s[i] * sc2;
          if (use_pos) { const float dx = qx[r] - mk.x, dy = qy[r] - mk.y; v = fmaf(sqrtf(fmaf(dx, dx, dy * dy)), -it2, v); }
          sv[i] = (((i & 1) ? sg1 : sg0) == qs[r]) ? v : -INFINITY;
        }
        if (MODE == 0) {
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const float mx = fmaxf(sv[2 * r], sv[2 * r + 1]);
            if (mx > m[r]) { l[r] *= ex2f(m[r] - mx); m[r] = mx; }
            if (m[r] > -INFINITY) l[r] += ex2f(sv[2 * r] - m[r]) + ex2f(sv[2 * r + 1] - m[r]);
          }
        } else {
          const uint32_t* vrow = reinterpret_cast<const uint32_t*>(sm.rowm[1] + (kb + g) * 16);
          float dp[4] = {0.f, 0.f, 0.f, 0.f};
          mma16816(dp, gf, vrow[tq], vrow[tq + 4]);                   // dP[query, key] = dO_query . v_key
          float ds[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { const int r = i >> 1; ds[i] = ex2f(sv[i] - lse[r]) * (dp[i] - dl[r]); }
          sf[2 * half] = pack_half2(ds[0], ds[1]); sf[2 * half + 1] = pack_half2(ds[2], ds[3]);
        }
      }
      if (MODE) {
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const uint32_t* kt = reinterpret_cast<const uint32_t*>(sm.tr[0] + (n * 8 + g) * TS + kstep * 16);
          mma16816(dq[n], sf, kt[tq], kt[tq + 4]);                    // dQ[query, dim] += dS[query, key] K[key, dim]
        }
      }
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {                              // merge the 4 lanes of the quad (same row, different keys)
        const float mo = __shfl_xor_sync(0xffffffffu, m[r], o), lo_ = __shfl_xor_sync(0xffffffffu, l[r], o);
        const float mn = fmaxf(m[r], mo);
        l[r] = (m[r] > -INFINITY ? l[r] * ex2f(m[r] - mn) : 0.f) + (mo > -INFINITY ? lo_ * ex2f(mo - mn) : 0.f);
        m[r] = mn;
      }
      const int row = qr0 + g + 8 * r;
      if (tq == 0 && row < a.N) a.lse[(size_t)row * heads + h] = (m[r] + log2f(l[r])) * 0.6931471805599453f;      // natural log units
    }
    if (t < MT && q0 + t < a.N) {
      const float* o = a.O + (size_t)(q0 + t) * a.ldo + h * BHD;
      const float* gg = a.dO + (size_t)(q0 + t) * a.lddo + h * BHD;
      float d = 0.f;
#pragma unroll
      for (int j = 0; j < BHD; ++j) d = fmaf(o[j], gg[j], d);
      a.delta[(size_t)(q0 + t) * heads + h] = d;
    }
  } else {
    const float inv = a.scale / gs;
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = qr0 + g + 8 * r;
        if (row < a.N)
          *reinterpret_cast<float2*>(a.dqkv + (size_t)row * a.ldd + h * BHD + n * 8 + 2 * tq) = make_float2(dq[n][2 * r] * inv, dq[n][2 * r + 1] * inv);
      }
  }
}

}  // namespace dgdm

extern "C" int dgdm_spatial_attention_backward_tc(const float* qkv, int ld, const float* pos, const int* seg_of, const int* seg_ptr,
                                                  int N, int D, int heads, float temperature, const float* O, int ldo,
                                                  const float* dO, int lddo, const float* go_scale, int have_lse, float* lse, float* delta,
                                                  float* dqkv, int ldd, void* stream) {
  if (!qkv || !seg_of || !seg_ptr || !O || !dO || !go_scale || !lse || !delta || !dqkv) return DGDM_ERR_INVALID;
  if (heads <= 0 || D != heads * BHD || !(temperature > 0.f) || (ld % 4) || (ldo % 4) || (lddo % 4) || (ldd % 2)) return DGDM_ERR_INVALID;
  if (N <= 0) return DGDM_OK;
  AttnBwdArgs a;
  a.qkv = qkv; a.ld = ld; a.pos = pos; a.seg_of = seg_of; a.seg_ptr = seg_ptr; a.N = N; a.D = D;
  a.scale = 1.0f / sqrtf((float)BHD); a.inv_temp = 1.0f / temperature;
  a.O = O; a.ldo = ldo; a.dO = dO; a.lddo = lddo; a.lse = lse; a.delta = delta; a.dqkv = dqkv; a.ldd = ldd;
  const dim3 grid((N + MT - 1) / MT, heads);
  cudaStream_t s = (cudaStream_t)stream;
  if (have_lse) attn_bwd_delta_kernel<<<(unsigned)(((size_t)N * heads + 255) / 256), 256, 0, s>>>(O, ldo, dO, lddo, N, heads, delta);
  else attn_bwd_q_tc_kernel<0><<<grid, 128, 0, s>>>(a, go_scale);
  attn_bwd_dkdv_tc_kernel<<<grid, 128, 0, s>>>(a, go_scale);
  attn_bwd_q_tc_kernel<1><<<grid, 128, 0, s>>>(a, go_scale);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}
