// General dense multi-head attention for the PUBLIC MultiHeadAttention / SpatialAttention API (core/attention.py:73-181,
// 285-327): additive / boolean attn_mask with torch broadcasting strides, key_padding_mask, cross attention (L_q != L_k),
// distinct value tensor, per-head or head-averaged weights. This is the O(L_q * L_k)-output path (the caller asked for the
// weights, or passed an arbitrary mask); the model's own forward never comes here — it uses the flash kernels.
// One warp per (batch, query row); exact fp32; two passes over the keys per head (online max/sum, then weights + PV).
#include "common.cuh"

namespace dgdm {

struct DenseMhaArgs {
  const float* q; const float* k; const float* v;
  int ldq, ldk, ldv;
  int B, Lq, Lk, D, heads;
  const float* mask; const unsigned char* bmask;        // additive / boolean (True = -inf); same 4-D strides
  long long ms_b, ms_h, ms_q, ms_k;
  const unsigned char* kpm;                              // [B, Lk] True = masked
  const float* pos_q; const float* pos_k;                // [P, Lq, 2] / [P, Lk, 2] or NULL: bias = -|p_i - p_j| * inv_temp
  int pos_by_head;                                       // 0: P indexed by the batch element, 1: by the HEAD (see dgdm_b200.h)
  float inv_temp;
  float* out; int ldo;
  float* weights; int per_head;                          // NULL | [B*H, Lq, Lk] | [B, Lq, Lk]
};

constexpr int DENSE_WARPS = 8;
constexpr int DENSE_MAX_D = 1024;

__device__ __forceinline__ float dense_score(const DenseMhaArgs& a, const float* qh, int b, int h, int i, int j, int hd, float scale,
                                             float pix, float piy, const float* pk) {
  const float* kr = a.k + ((size_t)b * a.Lk + j) * a.ldk + h * hd;
  float s = 0.f;
  for (int d = 0; d < hd; ++d) s = fmaf(qh[d], __ldg(kr + d), s);
  s *= scale;                                                     // matmul / sqrt(head_dim)  (:135)
  const long long mo = (long long)b * a.ms_b + (long long)h * a.ms_h + (long long)i * a.ms_q + (long long)j * a.ms_k;
  if (pk) { const float dx = pix - pk[2 * j], dy = piy - pk[2 * j + 1]; s += -sqrtf(dx * dx + dy * dy) * a.inv_temp; }
  if (a.mask) s += a.mask[mo];                                    // attn_scores += attn_mask  (:142)
  if (a.bmask && a.bmask[mo]) s = -INFINITY;                      // masked_fill_(attn_mask, -inf)  (:140)
  if (a.kpm && a.kpm[(size_t)b * a.Lk + j]) s = -INFINITY;        // key_padding_mask  (:145-150)
  return s;
}

__global__ void __launch_bounds__(DENSE_WARPS * 32) mha_dense_kernel(const DenseMhaArgs a) {
  extern __shared__ float q_s[];                                   // [DENSE_WARPS][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * DENSE_WARPS + warp;
  if (row >= (long long)a.B * a.Lq) return;
  const int b = (int)(row / a.Lq), i = (int)(row - (long long)b * a.Lq);
  const int hd = a.D / a.heads;
  const float scale = 1.0f / sqrtf((float)hd);
  float* qrow = q_s + warp * a.D;
  for (int d = lane; d < a.D; d += 32) qrow[d] = a.q[(size_t)row * a.ldq + d];
  __syncwarp();
  const float invh = 1.0f / (float)a.heads;
  float* wmean = (a.weights && !a.per_head) ? a.weights + (size_t)row * a.Lk : nullptr;
  if (wmean) for (int j = lane; j < a.Lk; j += 32) wmean[j] = 0.f;
  for (int h = 0; h < a.heads; ++h) {
    const float* qh = qrow + h * hd;
    const int p = a.pos_by_head ? h : b;
    const float* pq = a.pos_q ? a.pos_q + ((size_t)p * a.Lq + i) * 2 : nullptr;
    const float* pk = a.pos_k ? a.pos_k + (size_t)p * a.Lk * 2 : nullptr;
    const float pix = pq ? pq[0] : 0.f, piy = pq ? pq[1] : 0.f;
    float m = -INFINITY, l = 0.f;
    for (int j = lane; j < a.Lk; j += 32) {
      const float s = dense_score(a, qh, b, h, i, j, hd, scale, pix, piy, pk);
      if (s > m) { l = l * expf(m - s) + 1.0f; m = s; } else if (s > -INFINITY) l += expf(s - m);
    }
    const float mw = warp_max(m);
    l = (m > -INFINITY) ? l * expf(m - mw) : 0.f;
    l = warp_sum(l);
    // every key masked: softmax of a row of -inf is NaN in torch; reproduce (0 * inf below)
    const float inv = 1.0f / l;
    float* wh = (a.weights && a.per_head) ? a.weights + (((size_t)b * a.heads + h) * a.Lq + i) * a.Lk : nullptr;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int base = 0; base < a.Lk; base += 32) {
      const int j = base + lane;
      float w = 0.f;
      if (j < a.Lk) {
        const float s = dense_score(a, qh, b, h, i, j, hd, scale, pix, piy, pk);
        w = (mw > -INFINITY) ? expf(s - mw) * inv : __int_as_float(0x7fc00000);
        if (wh) wh[j] = w;
        if (wmean) wmean[j] += w * invh;
      }
      const int cnt = min(32, a.Lk - base);
      for (int jj = 0; jj < cnt; ++jj) {
        const float wj = __shfl_sync(0xffffffffu, w, jj);
        const float* vr = a.v + ((size_t)b * a.Lk + base + jj) * a.ldv + h * hd;
#pragma unroll
        for (int t = 0; t < 4; ++t) { const int d = lane + 32 * t; if (d < hd) acc[t] = fmaf(wj, __ldg(vr + d), acc[t]); }
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) { const int d = lane + 32 * t; if (d < hd) a.out[(size_t)row * a.ldo + h * hd + d] = acc[t]; }
  }
}

}  // namespace dgdm
using namespace dgdm;

extern "C" int dgdm_mha_dense(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, int B, int Lq, int Lk,
                              int D, int heads, const float* mask, const unsigned char* bool_mask, long long ms_b, long long ms_h,
                              long long ms_q, long long ms_k, const unsigned char* key_padding_mask, const float* pos_q,
                              const float* pos_k, int pos_by_head, float temperature, float* out, int ldo, float* weights,
                              int per_head, void* stream) {
  if (!q || !k || !v || !out || B <= 0 || Lq <= 0 || Lk <= 0 || heads <= 0 || D <= 0 || D % heads != 0) return DGDM_ERR_INVALID;
  if (D > DENSE_MAX_D || D / heads > 128 || !(temperature > 0.f) || ((pos_q == nullptr) != (pos_k == nullptr))) return DGDM_ERR_INVALID;
  DenseMhaArgs a;
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.B = B; a.Lq = Lq; a.Lk = Lk; a.D = D; a.heads = heads;
  a.mask = mask; a.bmask = bool_mask; a.ms_b = ms_b; a.ms_h = ms_h; a.ms_q = ms_q; a.ms_k = ms_k;
  a.kpm = key_padding_mask; a.pos_q = pos_q; a.pos_k = pos_k; a.pos_by_head = pos_by_head; a.inv_temp = 1.0f / temperature;
  a.out = out; a.ldo = ldo; a.weights = weights; a.per_head = per_head;
  const long long rows = (long long)B * Lq;
  const size_t smem = (size_t)DENSE_WARPS * D * sizeof(float);
  mha_dense_kernel<<<(unsigned)((rows + DENSE_WARPS - 1) / DENSE_WARPS), DENSE_WARPS * 32, smem, (cudaStream_t)stream>>>(a);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}
