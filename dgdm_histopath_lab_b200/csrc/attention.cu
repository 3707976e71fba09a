// Spatial attention (core/attention.py:225-327) as a flash-style kernel that never forms the N x N scores:
//   out_i = softmax_j( q_i.k_j/sqrt(hd) - |p_i - p_j|/T ) v_j     per head, keys restricted to the query's graph.
//
// Roofline: the kernel is bound by exp2 (one MUFU op per score; 8 heads x N^2 scores) — not by HBM and not by
// the tensor pipe (QK^T has K = head_dim = 16, PV has N = 16). Design:
//   * one warp owns 16 query rows and ALL heads of its head group, so the distance bias -|p_i-p_j| (a sqrt per
//     (i,j)) is computed once in registers and reused by every head as the C operand of the QK^T MMA;
//   * QK^T and PV run on mma.sync m16n8k16 fp16 with fp32 accumulate, P stays in registers (accumulator layout
//     == A-operand layout), K/V tiles staged through shared memory with cp.async double buffering + ldmatrix;
//   * softmax uses a lazily-updated row reference max: the common path is sub, exp2, add per score; the max is
//     only recomputed (with shuffles and an O/l rescale) when a lane's block sum signals fp16 overflow risk;
//   * PRECISE=true splits q,k,p,v into fp16 hi+lo planes (3 MMAs per product) for ~fp32 accuracy;
//   * keys can be split across CTAs (grid.y); partial (O, m, l) are merged by a tiny combine kernel.
#include "common.cuh"

namespace dgdm {

constexpr int ATT_WARPS = 8;                 // row groups (16 query rows each) per CTA
constexpr int ATT_THREADS = 2 * ATT_WARPS * 32; // 16 warps: two head-halves x 8 row groups
constexpr int ATT_QT = ATT_WARPS * 16;      // 128 query rows per CTA
constexpr int ATT_KS = 64;                  // keys per smem stage
constexpr int ATT_DG = 128;                 // columns (heads_in_group * head_dim) handled per CTA
constexpr int ATT_ROWB = ATT_DG * 2 + 16;   // padded smem row (bytes): 272 -> conflict-free ldmatrix
constexpr float ATT_RESCALE_THR = 8192.0f;  // lane block-sum above this => refresh the reference max

struct AttnParams {
  const __half* Qh; const __half* Ql; const float2* qpos;       // [Nq, D], [Nq]
  const __half* Kh; const __half* Kl; const __half* Vh; const __half* Vl; const float2* kpos;   // [Nkv, D], [Nkv]
  const int* qseg; const int* kseg; int num_seg; int Nq; int D; int heads; int use_bias; int nsplit;
  float* Opart; float* Mpart; float* Lpart;                      // [nsplit][Nq][D], [nsplit][Nq][heads] x2
};

__device__ __forceinline__ void mma16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1, const float* c) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}
__device__ __forceinline__ void ldsm4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp8(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

template <int HD, bool PRECISE, bool FULL>
__global__ void __launch_bounds__(ATT_THREADS, 1) spatial_attn_kernel(const AttnParams p) {
  static_assert(HD == 16, "head_dim 16 specialisation");
  constexpr int G = ATT_DG / HD / 2;                   // heads per warp (4): warps 0-7 own heads 0-3, warps 8-15 heads 4-7
  constexpr int NPL = PRECISE ? 2 : 1;                 // fp16 planes per operand
  constexpr int KB = PRECISE ? 16 : 32;                // keys per inner block (register budget: 128 regs/thread)
  constexpr int NT = KB / 8;                           // 8-key n-tiles per inner block
  constexpr int PLANE_B = ATT_KS * ATT_ROWB;           // bytes per K or V plane per stage
  constexpr int STAGE_B = 2 * NPL * PLANE_B + ATT_KS * 8;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int rw = warp & (ATT_WARPS - 1), hh = warp / ATT_WARPS;     // row group, head half
  // ---- locate this CTA's (segment, query tile)
  int seg = -1, row0 = 0;
  {
    int acc = 0;
    for (int s = 0; s < p.num_seg; ++s) {
      const int a = p.qseg[s], b = p.qseg[s + 1];
      const int nt = (b - a + ATT_QT - 1) / ATT_QT;
      if ((int)blockIdx.x < acc + nt) { seg = s; row0 = a + ((int)blockIdx.x - acc) * ATT_QT; break; }
      acc += nt;
    }
  }
  if (seg < 0) return;
  const int qend = p.qseg[seg + 1];
  const int kbeg = p.kseg[seg], kend = p.kseg[seg + 1];
  const int split = blockIdx.y;
  int per = (kend - kbeg + p.nsplit - 1) / p.nsplit;
  per = (per + ATT_KS - 1) / ATT_KS * ATT_KS;
  const int ks = kbeg + split * per;
  const int ke = min(kend, ks + per);
  const int hg = blockIdx.z;                           // head group of 8 heads (128 columns)
  const int col0 = hg * ATT_DG;                        // first column staged in smem
  const int head0 = hg * (2 * G) + hh * G;             // first head of this warp
  const int nheads = FULL ? G : max(0, min(G, p.heads - head0));

  const int r_lo = row0 + rw * 16 + g, r_hi = r_lo + 8;
  const bool warp_active = (row0 + rw * 16) < qend && nheads > 0;
  const int rl = min(r_lo, qend - 1), rh = min(r_hi, qend - 1);

  // ---- Q fragments (A operand, 16x16 per head), query positions
  uint32_t qh[G][4], ql[PRECISE ? G : 1][4];
#pragma unroll
  for (int h = 0; h < G; ++h) {
    const int c = (head0 + h) * HD + 2 * t;
    const bool hv = FULL || h < nheads;
    qh[h][0] = hv ? *reinterpret_cast<const uint32_t*>(p.Qh + (size_t)rl * p.D + c) : 0u;
    qh[h][1] = hv ? *reinterpret_cast<const uint32_t*>(p.Qh + (size_t)rh * p.D + c) : 0u;
    qh[h][2] = hv ? *reinterpret_cast<const uint32_t*>(p.Qh + (size_t)rl * p.D + c + 8) : 0u;
    qh[h][3] = hv ? *reinterpret_cast<const uint32_t*>(p.Qh + (size_t)rh * p.D + c + 8) : 0u;
    if constexpr (PRECISE) {
      ql[h][0] = hv ? *reinterpret_cast<const uint32_t*>(p.Ql + (size_t)rl * p.D + c) : 0u;
      ql[h][1] = hv ? *reinterpret_cast<const uint32_t*>(p.Ql + (size_t)rh * p.D + c) : 0u;
      ql[h][2] = hv ? *reinterpret_cast<const uint32_t*>(p.Ql + (size_t)rl * p.D + c + 8) : 0u;
      ql[h][3] = hv ? *reinterpret_cast<const uint32_t*>(p.Ql + (size_t)rh * p.D + c + 8) : 0u;
    }
  }
  float2 qp_lo = make_float2(0.f, 0.f), qp_hi = make_float2(0.f, 0.f);
  if (p.use_bias) { qp_lo = p.qpos[rl]; qp_hi = p.qpos[rh]; }

  float o[G][2][4];                                     // [head][n-tile (8 dims)][c-frag]
  float m_lo[G], m_hi[G], l_lo[G], l_hi[G];
#pragma unroll
  for (int h = 0; h < G; ++h) {
    m_lo[h] = m_hi[h] = -INFINITY; l_lo[h] = l_hi[h] = 0.f;
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) o[h][n][i] = 0.f;
  }

  // ---- K/V stage loader (all threads)
  auto load_stage = [&](int stage, int key0) {
    const uint32_t st = sbase + stage * STAGE_B;
    for (int i = tid; i < ATT_KS * 16; i += ATT_THREADS) {        // 16-byte chunks: 64 rows x 16 per plane
      const int r = i >> 4, c = i & 15;
      const int key = key0 + r;
      const bool v = key < ke && (col0 + c * 8) < p.D;      // zero-fill padded keys and columns beyond D
      const size_t goff = v ? (size_t)key * p.D + col0 + c * 8 : (size_t)kbeg * p.D;
      const uint32_t so = r * ATT_ROWB + c * 16;
      cp16(st + so, p.Kh + goff, v);
      cp16(st + NPL * PLANE_B + so, p.Vh + goff, v);
      if constexpr (PRECISE) {
        cp16(st + PLANE_B + so, p.Kl + goff, v);
        cp16(st + 3 * PLANE_B + so, p.Vl + goff, v);
      }
    }
    if (p.use_bias && tid < ATT_KS) {
      const int key = key0 + tid;
      cp8(st + 2 * NPL * PLANE_B + tid * 8, p.kpos + (key < ke ? key : kbeg), key < ke);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int nstages = (ke > ks) ? (ke - ks + ATT_KS - 1) / ATT_KS : 0;
  if (nstages > 0) load_stage(0, ks);
  for (int si = 0; si < nstages; ++si) {
    if (si + 1 < nstages) { load_stage((si + 1) & 1, ks + (si + 1) * ATT_KS); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (warp_active) {
      const uint32_t st = sbase + (si & 1) * STAGE_B;
      const uint32_t sK = st, sV = st + NPL * PLANE_B;
      const float4* kp4 = reinterpret_cast<const float4*>(smem + (si & 1) * STAGE_B + 2 * NPL * PLANE_B);
      const int key0 = ks + si * ATT_KS;
#pragma unroll 1
      for (int kb = 0; kb < ATT_KS; kb += KB) {
        if (key0 + kb >= ke) break;
        // ---- bias tile: NT n-tiles x (2 rows x 2 cols); -|p_i - p_j| (positions pre-scaled by log2e/T), -inf for padded keys
        float bias[NT][4];
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const int kk = kb + n * 8 + 2 * t;                // keys kk, kk+1 of this stage
          float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
          if (p.use_bias) {
            const float4 kp = kp4[kk >> 1];                 // (x0,y0,x1,y1)
            float dx = qp_lo.x - kp.x, dy = qp_lo.y - kp.y; b0 = -sqrt_approx(fmaf(dx, dx, dy * dy));
            dx = qp_lo.x - kp.z; dy = qp_lo.y - kp.w;       b1 = -sqrt_approx(fmaf(dx, dx, dy * dy));
            dx = qp_hi.x - kp.x; dy = qp_hi.y - kp.y;       b2 = -sqrt_approx(fmaf(dx, dx, dy * dy));
            dx = qp_hi.x - kp.z; dy = qp_hi.y - kp.w;       b3 = -sqrt_approx(fmaf(dx, dx, dy * dy));
          }
          if (key0 + kk >= ke) { b0 = -INFINITY; b2 = -INFINITY; }
          if (key0 + kk + 1 >= ke) { b1 = -INFINITY; b3 = -INFINITY; }
          bias[n][0] = b0; bias[n][1] = b1; bias[n][2] = b2; bias[n][3] = b3;
        }
#pragma unroll
        for (int h = 0; h < G; ++h) {
          if (!FULL && h >= nheads) break;
          const int hc = (hh * G + h) * (HD * 2);           // byte offset of this head inside a staged row
          // ---- S = Q K^T + bias : NT n-tiles of 8 keys
          float s[NT][4];
          const uint32_t kaddr = sK + (kb + (lane >> 4) * 8 + (lane & 7)) * ATT_ROWB + hc + ((lane >> 3) & 1) * 16;
#pragma unroll
          for (int n2 = 0; n2 < NT / 2; ++n2) {             // each ldmatrix.x4 covers 16 keys (2 n-tiles)
            uint32_t kf[4];
            ldsm4(kf, kaddr + n2 * 16 * ATT_ROWB);
            mma16816(s[2 * n2], qh[h], kf[0], kf[1], bias[2 * n2]);
            mma16816(s[2 * n2 + 1], qh[h], kf[2], kf[3], bias[2 * n2 + 1]);
            if constexpr (PRECISE) {
              uint32_t kl[4];
              ldsm4(kl, kaddr + PLANE_B + n2 * 16 * ATT_ROWB);
              mma16816(s[2 * n2], ql[h], kf[0], kf[1], s[2 * n2]);
              mma16816(s[2 * n2 + 1], ql[h], kf[2], kf[3], s[2 * n2 + 1]);
              mma16816(s[2 * n2], qh[h], kl[0], kl[1], s[2 * n2]);
              mma16816(s[2 * n2 + 1], qh[h], kl[2], kl[3], s[2 * n2 + 1]);
            }
          }
          // ---- lazy-max softmax
          float pr[NT][4];
          float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            pr[n][0] = ex2(s[n][0] - m_lo[h]); pr[n][1] = ex2(s[n][1] - m_lo[h]);
            pr[n][2] = ex2(s[n][2] - m_hi[h]); pr[n][3] = ex2(s[n][3] - m_hi[h]);
            sum_lo += pr[n][0] + pr[n][1]; sum_hi += pr[n][2] + pr[n][3];
          }
          const bool bad = !(sum_lo <= ATT_RESCALE_THR) || !(sum_hi <= ATT_RESCALE_THR);
          if (__any_sync(0xffffffffu, bad)) {
            float mx_lo = m_lo[h], mx_hi = m_hi[h];
#pragma unroll
            for (int n = 0; n < NT; ++n) {
              mx_lo = fmaxf(mx_lo, fmaxf(s[n][0], s[n][1])); mx_hi = fmaxf(mx_hi, fmaxf(s[n][2], s[n][3]));
            }
            mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1)); mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
            mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1)); mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
            const float sc_lo = (mx_lo == -INFINITY) ? 1.f : ex2(m_lo[h] - mx_lo);
            const float sc_hi = (mx_hi == -INFINITY) ? 1.f : ex2(m_hi[h] - mx_hi);
            const float ml = (mx_lo == -INFINITY) ? 0.f : mx_lo, mh = (mx_hi == -INFINITY) ? 0.f : mx_hi;
            m_lo[h] = mx_lo; m_hi[h] = mx_hi;
            l_lo[h] *= sc_lo; l_hi[h] *= sc_hi;
#pragma unroll
            for (int n = 0; n < 2; ++n) { o[h][n][0] *= sc_lo; o[h][n][1] *= sc_lo; o[h][n][2] *= sc_hi; o[h][n][3] *= sc_hi; }
            sum_lo = 0.f; sum_hi = 0.f;
#pragma unroll
            for (int n = 0; n < NT; ++n) {
              pr[n][0] = ex2(s[n][0] - ml); pr[n][1] = ex2(s[n][1] - ml);
              pr[n][2] = ex2(s[n][2] - mh); pr[n][3] = ex2(s[n][3] - mh);
              sum_lo += pr[n][0] + pr[n][1]; sum_hi += pr[n][2] + pr[n][3];
            }
          }
          l_lo[h] += sum_lo; l_hi[h] += sum_hi;
          // ---- O += P V : KB/16 k16 steps (16 keys each) x two n-tiles (8 dims each)
          const uint32_t vaddr = sV + (kb + ((lane >> 3) & 1) * 8 + (lane & 7)) * ATT_ROWB + hc + (lane >> 4) * 16;
#pragma unroll
          for (int k2 = 0; k2 < NT / 2; ++k2) {
            uint32_t pa[4], vf[4];
            pa[0] = pack_h2(pr[2 * k2][0], pr[2 * k2][1]);     pa[1] = pack_h2(pr[2 * k2][2], pr[2 * k2][3]);
            pa[2] = pack_h2(pr[2 * k2 + 1][0], pr[2 * k2 + 1][1]); pa[3] = pack_h2(pr[2 * k2 + 1][2], pr[2 * k2 + 1][3]);
            ldsm4t(vf, vaddr + k2 * 16 * ATT_ROWB);
            mma16816(o[h][0], pa, vf[0], vf[1], o[h][0]);
            mma16816(o[h][1], pa, vf[2], vf[3], o[h][1]);
            if constexpr (PRECISE) {
              uint32_t pl[4], vl[4];
              {
                __half2 a;
                a = *reinterpret_cast<__half2*>(&pa[0]); pl[0] = pack_h2(pr[2 * k2][0] - __low2float(a), pr[2 * k2][1] - __high2float(a));
                a = *reinterpret_cast<__half2*>(&pa[1]); pl[1] = pack_h2(pr[2 * k2][2] - __low2float(a), pr[2 * k2][3] - __high2float(a));
                a = *reinterpret_cast<__half2*>(&pa[2]); pl[2] = pack_h2(pr[2 * k2 + 1][0] - __low2float(a), pr[2 * k2 + 1][1] - __high2float(a));
                a = *reinterpret_cast<__half2*>(&pa[3]); pl[3] = pack_h2(pr[2 * k2 + 1][2] - __low2float(a), pr[2 * k2 + 1][3] - __high2float(a));
              }
              ldsm4t(vl, vaddr + PLANE_B + k2 * 16 * ATT_ROWB);
              mma16816(o[h][0], pl, vf[0], vf[1], o[h][0]);
              mma16816(o[h][1], pl, vf[2], vf[3], o[h][1]);
              mma16816(o[h][0], pa, vl[0], vl[1], o[h][0]);
              mma16816(o[h][1], pa, vl[2], vl[3], o[h][1]);
            }
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- write partials (unnormalised O, reference max m, row sum l)
  if (warp_active) {
#pragma unroll
    for (int h = 0; h < G; ++h) {
      if (!FULL && h >= nheads) break;
      float ll = l_lo[h], lh = l_hi[h];
      ll += __shfl_xor_sync(0xffffffffu, ll, 1); ll += __shfl_xor_sync(0xffffffffu, ll, 2);
      lh += __shfl_xor_sync(0xffffffffu, lh, 1); lh += __shfl_xor_sync(0xffffffffu, lh, 2);
      const int head = head0 + h;
      const size_t pbase = (size_t)split * p.Nq;
      if (r_lo < qend) {
        float* op = p.Opart + (pbase + r_lo) * p.D + head * HD + 2 * t;
        *reinterpret_cast<float2*>(op) = make_float2(o[h][0][0], o[h][0][1]);
        *reinterpret_cast<float2*>(op + 8) = make_float2(o[h][1][0], o[h][1][1]);
        if (t == 0) { p.Mpart[(pbase + r_lo) * p.heads + head] = m_lo[h]; p.Lpart[(pbase + r_lo) * p.heads + head] = ll; }
      }
      if (r_hi < qend) {
        float* op = p.Opart + (pbase + r_hi) * p.D + head * HD + 2 * t;
        *reinterpret_cast<float2*>(op) = make_float2(o[h][0][2], o[h][0][3]);
        *reinterpret_cast<float2*>(op + 8) = make_float2(o[h][1][2], o[h][1][3]);
        if (t == 0) { p.Mpart[(pbase + r_hi) * p.heads + head] = m_hi[h]; p.Lpart[(pbase + r_hi) * p.heads + head] = lh; }
      }
    }
  }
}

// merge key-split partials: out[r, c] = sum_s O_s 2^(m_s - m) / sum_s l_s 2^(m_s - m)
__global__ void __launch_bounds__(256) attn_combine_kernel(const float* __restrict__ Opart, const float* __restrict__ Mpart,
                                                           const float* __restrict__ Lpart, int nsplit, int Nq, int D, int heads,
                                                           float* __restrict__ out, int ldo, float* __restrict__ lse_out) {
  const int hd = D / heads;
  const size_t total = (size_t)Nq * (D / 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / (D / 4)), c = (int)(i - (size_t)r * (D / 4)) * 4;
    const int h = c / hd;
    float m = -INFINITY;
    for (int s = 0; s < nsplit; ++s) m = fmaxf(m, Mpart[((size_t)s * Nq + r) * heads + h]);
    float l = 0.f; float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < nsplit; ++s) {
      const float ms = Mpart[((size_t)s * Nq + r) * heads + h];
      if (ms == -INFINITY) continue;
      const float w = exp2f(ms - m);
      l = fmaf(Lpart[((size_t)s * Nq + r) * heads + h], w, l);
      const float4 o = *reinterpret_cast<const float4*>(Opart + ((size_t)s * Nq + r) * D + c);
      acc.x = fmaf(o.x, w, acc.x); acc.y = fmaf(o.y, w, acc.y); acc.z = fmaf(o.z, w, acc.z); acc.w = fmaf(o.w, w, acc.w);
    }
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    *reinterpret_cast<float4*>(out + (size_t)r * ldo + c) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    // log-sum-exp of the row's scores in natural units (m is a log2-domain reference, l the sum of 2^(s - m)): saved for the backward
    if (lse_out != nullptr && c == h * hd) lse_out[(size_t)r * heads + h] = (m + log2f(l)) * 0.6931471805599453f;
  }
}

// fp32 qkv -> fp16 operand planes. Q is pre-multiplied by log2(e)/sqrt(hd) so the MMA yields log2-domain scores;
// positions are pre-multiplied by log2(e)/T so that -|dp| is the log2-domain bias.
__global__ void __launch_bounds__(256) attn_prep_q_kernel(const float* __restrict__ qkv, int ld, int N, int D, float qscale,
                                                          __half* __restrict__ Qh, __half* __restrict__ Ql) {
  const size_t total = (size_t)N * D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / D), c = (int)(i - (size_t)r * D);
    const float q = qkv[(size_t)r * ld + c] * qscale;
    const __half h = __float2half_rn(q);
    Qh[i] = h;
    if (Ql) Ql[i] = __float2half_rn(q - __half2float(h));
  }
}
__global__ void __launch_bounds__(256) attn_prep_kv_kernel(const float* __restrict__ qkv, int ld, int N, int D,
                                                           __half* __restrict__ Kh, __half* __restrict__ Kl,
                                                           __half* __restrict__ Vh, __half* __restrict__ Vl) {
  const size_t total = (size_t)N * D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / D), c = (int)(i - (size_t)r * D);
    const float k = qkv[(size_t)r * ld + D + c], v = qkv[(size_t)r * ld + 2 * D + c];
    const __half kh = __float2half_rn(k), vh = __float2half_rn(v);
    Kh[i] = kh; Vh[i] = vh;
    if (Kl) { Kl[i] = __float2half_rn(k - __half2float(kh)); Vl[i] = __float2half_rn(v - __half2float(vh)); }
  }
}
__global__ void attn_prep_pos_kernel(const float* __restrict__ pos, int N, float pscale, float2* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
    out[i] = make_float2(pos[2 * i] * pscale, pos[2 * i + 1] * pscale);
}

// ---------------------------------------------------------------- dense head-averaged weights (return_attention)
// One warp per query row; two passes over the keys per head (max, then normalised weights), fp32 throughout.
__global__ void __launch_bounds__(256) attn_weights_kernel(const float* __restrict__ qkv, int ld, const float* __restrict__ pos,
                                                           int n, int D, int heads, float inv_temp, float* __restrict__ W) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  const int hd = D / heads;
  const float scale = 1.0f / sqrtf((float)hd);
  const float px = pos ? pos[2 * row] : 0.f, py = pos ? pos[2 * row + 1] : 0.f;
  for (int j = lane; j < n; j += 32) W[(size_t)row * n + j] = 0.f;
  const float invh = 1.0f / (float)heads;
  for (int h = 0; h < heads; ++h) {
    const float* q = qkv + (size_t)row * ld + h * hd;
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) {
      const float* k = qkv + (size_t)j * ld + D + h * hd;
      float s = 0.f;
      for (int d = 0; d < hd; ++d) s = fmaf(q[d], k[d], s);
      s *= scale;
      if (pos) { const float dx = px - pos[2 * j], dy = py - pos[2 * j + 1]; s -= sqrtf(dx * dx + dy * dy) * inv_temp; }
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) {
      const float* k = qkv + (size_t)j * ld + D + h * hd;
      float s = 0.f;
      for (int d = 0; d < hd; ++d) s = fmaf(q[d], k[d], s);
      s *= scale;
      if (pos) { const float dx = px - pos[2 * j], dy = py - pos[2 * j + 1]; s -= sqrtf(dx * dx + dy * dy) * inv_temp; }
      sum += expf(s - mx);
    }
    sum = warp_sum(sum);
    const float inv = invh / sum;
    for (int j = lane; j < n; j += 32) {
      const float* k = qkv + (size_t)j * ld + D + h * hd;
      float s = 0.f;
      for (int d = 0; d < hd; ++d) s = fmaf(q[d], k[d], s);
      s *= scale;
      if (pos) { const float dx = px - pos[2 * j], dy = py - pos[2 * j + 1]; s -= sqrtf(dx * dx + dy * dy) * inv_temp; }
      W[(size_t)row * n + j] += expf(s - mx) * inv;
    }
  }
}

// ---------------------------------------------------------------- GlobalAttentionPool (one query per graph)
constexpr int POOL_CHUNKS = 64;
__global__ void __launch_bounds__(256) pool_scores_kernel(const float* __restrict__ q_tok, const float* __restrict__ kv, int ld, int N,
                                                          int D, int heads, float* __restrict__ scores) {
  const int hd = D / heads;
  const float scale = 1.0f / sqrtf((float)hd);
  const size_t total = (size_t)N * heads;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i / heads), h = (int)(i - (size_t)j * heads);
    const float* k = kv + (size_t)j * ld + h * hd;
    const float* q = q_tok + h * hd;
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s = fmaf(q[d], k[d], s);
    scores[i] = s * scale;
  }
}
// block per (segment, head): max and sum(exp) over the segment's nodes
__global__ void __launch_bounds__(256) pool_stats_kernel(const float* __restrict__ scores, const int* __restrict__ seg_ptr, int heads,
                                                         float* __restrict__ stats) {
  const int seg = blockIdx.x / heads, h = blockIdx.x - seg * heads;
  const int a = seg_ptr[seg], b = seg_ptr[seg + 1];
  __shared__ float red[8];
  float mx = -INFINITY;
  for (int j = a + threadIdx.x; j < b; j += 256) mx = fmaxf(mx, scores[(size_t)j * heads + h]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int j = a + threadIdx.x; j < b; j += 256) sum += expf(scores[(size_t)j * heads + h] - mx);
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    stats[2 * blockIdx.x] = mx; stats[2 * blockIdx.x + 1] = s;
  }
}
// block per (segment, chunk): thread d accumulates sum_j w_j v[j,d] over its chunk of the segment
__global__ void pool_partial_kernel(const float* __restrict__ scores, const float* __restrict__ stats, const float* __restrict__ kv,
                                    int ld, const int* __restrict__ seg_ptr, int D, int heads, float* __restrict__ partial) {
  const int seg = blockIdx.x / POOL_CHUNKS, ch = blockIdx.x - seg * POOL_CHUNKS;
  const int a = seg_ptr[seg], b = seg_ptr[seg + 1];
  const int len = b - a, per = (len + POOL_CHUNKS - 1) / POOL_CHUNKS;
  const int j0 = a + ch * per, j1 = min(b, j0 + per);
  const int hd = D / heads;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const int h = d / hd;
    const float mx = stats[2 * (seg * heads + h)], sum = stats[2 * (seg * heads + h) + 1];
    float acc = 0.f;
    for (int j = j0; j < j1; ++j) acc = fmaf(expf(scores[(size_t)j * heads + h] - mx), kv[(size_t)j * ld + D + d], acc);
    partial[(size_t)blockIdx.x * D + d] = (sum > 0.f) ? acc / sum : 0.f;
  }
}
__global__ void pool_reduce_kernel(const float* __restrict__ partial, int D, float* __restrict__ out) {
  const int seg = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < POOL_CHUNKS; ++c) acc += partial[((size_t)seg * POOL_CHUNKS + c) * D + d];
    out[(size_t)seg * D + d] = acc;
  }
}

// Fused fast path (D <= 128, head_dim % 4 == 0, head_dim/4 a power of two): one pass over k|v with an online softmax.
// Block = (segment, chunk of the segment's rows), 8 warps; a lane owns 4 consecutive columns (LPH lanes per head), a warp owns
// rows j0+w, j0+w+8, ... (two rows in flight). The 8 warps and then the chunks are merged in a fixed order -> deterministic.
constexpr int POOL_FCHUNKS = 256;
template <int LPH>
__global__ void __launch_bounds__(256) pool_fused_kernel(const float* __restrict__ q_tok, const float* __restrict__ kv, int ld,
                                                         const int* __restrict__ seg_ptr, int D, int heads, int chunks,
                                                         float* __restrict__ pacc, float* __restrict__ pml) {
  const int seg = blockIdx.x / chunks, ch = blockIdx.x - seg * chunks;
  const int a = seg_ptr[seg], b = seg_ptr[seg + 1];
  const int per = (b - a + chunks - 1) / chunks;
  const int j0 = a + ch * per, j1 = min(b, j0 + per);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = lane * 4;
  const bool on = col < D;
  const float scale = 1.0f / sqrtf((float)(D / heads));
  const float4 q = on ? ldg4(q_tok + col) : make_float4(0.f, 0.f, 0.f, 0.f);
  float m = -INFINITY, l = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  auto head_dot = [&](float4 k) {
    float sdot = q.x * k.x + q.y * k.y + q.z * k.z + q.w * k.w;
#pragma unroll
    for (int o = 1; o < LPH; o <<= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
    return sdot * scale;
  };
  auto update = [&](float sc, float4 v) {
    const float mn = fmaxf(m, sc);
    const float corr = expf(m - mn), pr = expf(sc - mn);     // m = -inf on the first row: corr = 0
    l = l * corr + pr;
    acc.x = acc.x * corr + pr * v.x; acc.y = acc.y * corr + pr * v.y; acc.z = acc.z * corr + pr * v.z; acc.w = acc.w * corr + pr * v.w;
    m = mn;
  };
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  int j = j0 + warp;
  for (; j + 8 < j1; j += 16) {
    const float* r0 = kv + (size_t)j * ld + col; const float* r1 = kv + (size_t)(j + 8) * ld + col;
    const float4 k0 = on ? ldg4(r0) : z4, v0 = on ? ldg4(r0 + D) : z4, k1 = on ? ldg4(r1) : z4, v1 = on ? ldg4(r1 + D) : z4;
    update(head_dot(k0), v0);
    update(head_dot(k1), v1);
  }
  for (; j < j1; j += 8) {
    const float* r0 = kv + (size_t)j * ld + col;
    const float4 k0 = on ? ldg4(r0) : z4, v0 = on ? ldg4(r0 + D) : z4;
    update(head_dot(k0), v0);
  }
  __shared__ float sm_m[8][32], sm_l[8][32];
  __shared__ float4 sm_a[8][32];
  sm_m[warp][lane] = m; sm_l[warp][lane] = l; sm_a[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && on) {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) M = fmaxf(M, sm_m[w][lane]);
    float L = 0.f; float4 A = z4;
    if (M > -INFINITY) {
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const float e = expf(sm_m[w][lane] - M);
        const float4 t = sm_a[w][lane];
        L += sm_l[w][lane] * e; A.x += t.x * e; A.y += t.y * e; A.z += t.z * e; A.w += t.w * e;
      }
    }
    *reinterpret_cast<float4*>(pacc + (size_t)blockIdx.x * D + col) = A;
    if ((lane % LPH) == 0) { const int h = lane / LPH; pml[((size_t)blockIdx.x * heads + h) * 2] = M; pml[((size_t)blockIdx.x * heads + h) * 2 + 1] = L; }
  }
}
__global__ void pool_fused_merge_kernel(const float* __restrict__ pacc, const float* __restrict__ pml, int D, int heads, int chunks,
                                        float* __restrict__ out) {
  const int seg = blockIdx.x;
  const int hd = D / heads;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const int h = d / hd;
    float M = -INFINITY;
    for (int c = 0; c < chunks; ++c) M = fmaxf(M, pml[((size_t)(seg * chunks + c) * heads + h) * 2]);
    float L = 0.f, A = 0.f;
    if (M > -INFINITY)
      for (int c = 0; c < chunks; ++c) {
        const size_t blk = (size_t)seg * chunks + c;
        const float e = expf(pml[(blk * heads + h) * 2] - M);
        L += pml[(blk * heads + h) * 2 + 1] * e;
        A += pacc[blk * D + d] * e;
      }
    out[(size_t)seg * D + d] = L > 0.f ? A / L : 0.f;
  }
}

static inline int agrid(size_t n, int block = 256, int cap = 148 * 16) {
  size_t b = (n + block - 1) / block;
  return (int)(b < 1 ? 1 : (b > (size_t)cap ? cap : b));
}
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static int choose_nsplit(int Nq, int Nkv, int num_seg) {
  const int tiles = (Nq + ATT_QT - 1) / ATT_QT + num_seg;
  int s = (148 * 8 + tiles - 1) / tiles;            // aim for >= ~8 CTAs of work per SM
  const int avg_keys = Nkv / (num_seg > 0 ? num_seg : 1);
  const int max_s = avg_keys / (2 * ATT_KS) > 1 ? avg_keys / (2 * ATT_KS) : 1;
  if (s > max_s) s = max_s;
  if (s > 16) s = 16;
  if (s < 1) s = 1;
  return s;
}

void attn_launch_combine(const float* Opart, const float* Mpart, const float* Lpart, int nsplit, int Nq, int D, int heads,
                         float* out, int ldo, float* lse_out, cudaStream_t s) {
  attn_combine_kernel<<<agrid((size_t)Nq * (D / 4)), 256, 0, s>>>(Opart, Mpart, Lpart, nsplit, Nq, D, heads, out, ldo, lse_out);
}
// attention_simt.cu: exact fp32 CUDA-core path for head dimensions other than 16 (DGDM-Large: 16 heads x 8)
int attn_simt_launch(const float* q_qkv, int ld_q, const float* q_pos, const int* qseg_ptr, int Nq,
                     const float* kv_qkv, int ld_kv, const float* kv_pos, const int* kseg_ptr, int num_seg,
                     int D, int heads, float temperature, float* out, int ldo, cudaStream_t s);
// attention_tc.cu: tcgen05 / TMEM path (D = 128, 8 heads)
size_t attn_tc_work_bytes(int Nq, int Nkv, int D, int heads, int num_seg, int max_seg_keys);
int attn_tc_launch(const float* q_qkv, int ld_q, const float* q_pos, const int* qseg_ptr, int Nq,
                   const float* kv_qkv, int ld_kv, const float* kv_pos, const int* kseg_ptr, int Nkv,
                   int num_seg, int max_seg_keys, float temperature, float* out, int ldo, float* lse_out, void* work, size_t work_bytes,
                   cudaStream_t s);

}  // namespace dgdm
using namespace dgdm;

extern "C" size_t dgdm_spatial_attention_work_bytes(int Nq, int Nkv, int D, int heads, int num_seg, int max_seg_keys) {
  const int ns = choose_nsplit(Nq, Nkv, num_seg);
  size_t b = 0;
  if (D == 128 && heads == 8) b = attn_tc_work_bytes(Nq, Nkv, D, heads, num_seg, max_seg_keys);
  const size_t b_tc = b;
  b = 0;
  b += 2 * align256((size_t)Nq * D * 2);            // Qh, Ql
  b += 4 * align256((size_t)Nkv * D * 2);           // Kh, Kl, Vh, Vl
  b += align256((size_t)Nq * 8) + align256((size_t)Nkv * 8);
  b += align256((size_t)ns * Nq * D * 4);
  b += 2 * align256((size_t)ns * Nq * heads * 4);
  b += 256;
  return b > b_tc ? b : b_tc;
}

extern "C" int dgdm_spatial_attention(const float* q_qkv, int ld_q, const float* q_pos, const int* qseg_ptr, int Nq,
                                      const float* kv_qkv, int ld_kv, const float* kv_pos, const int* kseg_ptr, int Nkv,
                                      int num_seg, int max_seg_keys, int D, int heads, float temperature, int precise,
                                      float* out, int ldo, void* work, size_t work_bytes, void* stream) {
  return dgdm_spatial_attention_lse(q_qkv, ld_q, q_pos, qseg_ptr, Nq, kv_qkv, ld_kv, kv_pos, kseg_ptr, Nkv, num_seg, max_seg_keys, D, heads,
                                    temperature, precise, out, ldo, nullptr, work, work_bytes, stream);
}

extern "C" int dgdm_spatial_attention_lse(const float* q_qkv, int ld_q, const float* q_pos, const int* qseg_ptr, int Nq,
                                          const float* kv_qkv, int ld_kv, const float* kv_pos, const int* kseg_ptr, int Nkv,
                                          int num_seg, int max_seg_keys, int D, int heads, float temperature, int precise,
                                          float* out, int ldo, float* lse_out, void* work, size_t work_bytes, void* stream) {
  if (!q_qkv || !kv_qkv || !qseg_ptr || !kseg_ptr || !out || !work) return DGDM_ERR_INVALID;
  if (heads <= 0 || D % heads != 0 || (ldo % 4) != 0) return DGDM_ERR_INVALID;
  if ((q_pos == nullptr) != (kv_pos == nullptr)) return DGDM_ERR_INVALID;
  if (Nq <= 0) return DGDM_OK;
  if (Nkv <= 0 || num_seg <= 0 || !(temperature > 0.f)) return DGDM_ERR_INVALID;
  if ((D / heads != 16 || precise == 3) && lse_out != nullptr) return DGDM_ERR_INVALID;      // the CUDA-core path keeps no statistics
  if (D / heads != 16 || precise == 3)      // other head widths (DGDM-Large 16 x 8) / forced: exact fp32 CUDA-core kernel
    return attn_simt_launch(q_qkv, ld_q, q_pos, qseg_ptr, Nq, kv_qkv, ld_kv, kv_pos, kseg_ptr, num_seg, D, heads, temperature, out, ldo,
                            (cudaStream_t)stream);
  if (work_bytes < dgdm_spatial_attention_work_bytes(Nq, Nkv, D, heads, num_seg, max_seg_keys)) return DGDM_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  if (precise == 1 && D == 128 && heads == 8)       // product path for DGDM-Base: tcgen05 / TMEM kernel (attention_tc.cu)
    return attn_tc_launch(q_qkv, ld_q, q_pos, qseg_ptr, Nq, kv_qkv, ld_kv, kv_pos, kseg_ptr, Nkv, num_seg, max_seg_keys, temperature, out, ldo,
                          lse_out, work, work_bytes, s);
  const int ns = choose_nsplit(Nq, Nkv, num_seg);
  uint8_t* w = (uint8_t*)(((uintptr_t)work + 255) & ~(uintptr_t)255);
  auto take = [&](size_t bytes) { uint8_t* r = w; w += align256(bytes); return r; };
  __half* Qh = (__half*)take((size_t)Nq * D * 2); __half* Ql = (__half*)take((size_t)Nq * D * 2);
  __half* Kh = (__half*)take((size_t)Nkv * D * 2); __half* Kl = (__half*)take((size_t)Nkv * D * 2);
  __half* Vh = (__half*)take((size_t)Nkv * D * 2); __half* Vl = (__half*)take((size_t)Nkv * D * 2);
  float2* qpos = (float2*)take((size_t)Nq * 8); float2* kpos = (float2*)take((size_t)Nkv * 8);
  float* Opart = (float*)take((size_t)ns * Nq * D * 4);
  float* Mpart = (float*)take((size_t)ns * Nq * heads * 4); float* Lpart = (float*)take((size_t)ns * Nq * heads * 4);

  const float LOG2E = 1.4426950408889634f;
  attn_prep_q_kernel<<<agrid((size_t)Nq * D), 256, 0, s>>>(q_qkv, ld_q, Nq, D, LOG2E / sqrtf((float)(D / heads)), Qh, precise ? Ql : nullptr);
  attn_prep_kv_kernel<<<agrid((size_t)Nkv * D), 256, 0, s>>>(kv_qkv, ld_kv, Nkv, D, Kh, precise ? Kl : nullptr, Vh, precise ? Vl : nullptr);
  if (q_pos) {
    attn_prep_pos_kernel<<<agrid(Nq), 256, 0, s>>>(q_pos, Nq, LOG2E / temperature, qpos);
    attn_prep_pos_kernel<<<agrid(Nkv), 256, 0, s>>>(kv_pos, Nkv, LOG2E / temperature, kpos);
  }
  AttnParams p;
  p.Qh = Qh; p.Ql = Ql; p.qpos = qpos; p.Kh = Kh; p.Kl = Kl; p.Vh = Vh; p.Vl = Vl; p.kpos = kpos;
  p.qseg = qseg_ptr; p.kseg = kseg_ptr; p.num_seg = num_seg; p.Nq = Nq; p.D = D; p.heads = heads;
  p.use_bias = q_pos != nullptr; p.nsplit = ns; p.Opart = Opart; p.Mpart = Mpart; p.Lpart = Lpart;
  const int tiles = (Nq + ATT_QT - 1) / ATT_QT + num_seg;
  const int hgroups = (D + ATT_DG - 1) / ATT_DG;
  dim3 grid(tiles, ns, hgroups);
  const bool full = (heads % 8) == 0;
#define ATT_LAUNCH(PREC, FULLV)                                                                                          \
  do {                                                                                                                   \
    constexpr int smem = 2 * ((PREC ? 4 : 2) * ATT_KS * ATT_ROWB + ATT_KS * 8);                                          \
    static PerDeviceOnce set;                                                                                            \
    if (set.needed()) {                                                                                                  \
      if (cudaFuncSetAttribute(spatial_attn_kernel<16, PREC, FULLV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)  \
          != cudaSuccess) return DGDM_ERR_LAUNCH;                                                                        \
      set.mark();                                                                                                        \
    }                                                                                                                    \
    spatial_attn_kernel<16, PREC, FULLV><<<grid, ATT_THREADS, smem, s>>>(p);                                             \
  } while (0)
  if (precise) { if (full) ATT_LAUNCH(true, true); else ATT_LAUNCH(true, false); }
  else         { if (full) ATT_LAUNCH(false, true); else ATT_LAUNCH(false, false); }
#undef ATT_LAUNCH
  attn_combine_kernel<<<agrid((size_t)Nq * (D / 4)), 256, 0, s>>>(Opart, Mpart, Lpart, ns, Nq, D, heads, out, ldo, lse_out);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" int dgdm_spatial_attention_weights(const float* qkv, int ld, const float* pos, int n, int D, int heads,
                                              float temperature, float* weights, void* stream) {
  if (!qkv || !weights || heads <= 0 || D % heads != 0 || !(temperature > 0.f)) return DGDM_ERR_INVALID;
  if (n <= 0) return DGDM_OK;
  attn_weights_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(qkv, ld, pos, n, D, heads, 1.0f / temperature, weights);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" size_t dgdm_attention_pool_work_floats(int N, int D, int heads, int num_seg) {
  const size_t generic = (size_t)N * heads + (size_t)2 * num_seg * heads + (size_t)num_seg * POOL_CHUNKS * D + 64;
  const size_t fused = (size_t)num_seg * POOL_FCHUNKS * ((size_t)D + 2 * heads) + 64;
  return generic > fused ? generic : fused;
}

extern "C" int dgdm_attention_pool(const float* q_tok, const float* kv, int ld_kv, const int* seg_ptr, int num_seg, int N,
                                   int D, int heads, float* out, float* work, void* stream) {
  if (!q_tok || !kv || !seg_ptr || !out || !work || num_seg <= 0 || heads <= 0 || D % heads != 0 || D > 1024) return DGDM_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  const int hd = D / heads, lph = hd / 4;
  if (D <= 128 && (D % 4) == 0 && (hd % 4) == 0 && (lph & (lph - 1)) == 0 && (ld_kv % 4) == 0) {
    int chunks = (148 * 4 + num_seg - 1) / num_seg;
    if (chunks > POOL_FCHUNKS) chunks = POOL_FCHUNKS;
    if (chunks < 1) chunks = 1;
    float* pacc = work; float* pml = work + (size_t)num_seg * chunks * D;
    const int grid = num_seg * chunks;
    switch (lph) {
      case 1: pool_fused_kernel<1><<<grid, 256, 0, s>>>(q_tok, kv, ld_kv, seg_ptr, D, heads, chunks, pacc, pml); break;
      case 2: pool_fused_kernel<2><<<grid, 256, 0, s>>>(q_tok, kv, ld_kv, seg_ptr, D, heads, chunks, pacc, pml); break;
      case 4: pool_fused_kernel<4><<<grid, 256, 0, s>>>(q_tok, kv, ld_kv, seg_ptr, D, heads, chunks, pacc, pml); break;
      case 8: pool_fused_kernel<8><<<grid, 256, 0, s>>>(q_tok, kv, ld_kv, seg_ptr, D, heads, chunks, pacc, pml); break;
      case 16: pool_fused_kernel<16><<<grid, 256, 0, s>>>(q_tok, kv, ld_kv, seg_ptr, D, heads, chunks, pacc, pml); break;
      default: pool_fused_kernel<32><<<grid, 256, 0, s>>>(q_tok, kv, ld_kv, seg_ptr, D, heads, chunks, pacc, pml); break;
    }
    pool_fused_merge_kernel<<<num_seg, 128, 0, s>>>(pacc, pml, D, heads, chunks, out);
    DGDM_CHECK_LAUNCH();
    return DGDM_OK;
  }
  float* scores = work; float* stats = scores + (size_t)N * heads; float* partial = stats + (size_t)2 * num_seg * heads;
  if (N > 0) pool_scores_kernel<<<agrid((size_t)N * heads), 256, 0, s>>>(q_tok, kv, ld_kv, N, D, heads, scores);
  pool_stats_kernel<<<num_seg * heads, 256, 0, s>>>(scores, seg_ptr, heads, stats);
  pool_partial_kernel<<<num_seg * POOL_CHUNKS, D < 128 ? 128 : (D > 512 ? 1024 : ((D + 31) / 32 * 32)), 0, s>>>(scores, stats, kv, ld_kv, seg_ptr, D, heads, partial);
  pool_reduce_kernel<<<num_seg, 128, 0, s>>>(partial, D, out);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}
