// Spatial attention core on tcgen05 / TMEM (head_dim 16, 8 heads, D = 128 — DGDM-Base; core/attention.py:135-157,261-283):
//   out_i = softmax_j( q_i.k_j / sqrt(hd) - |p_i - p_j| / T ) v_j      per head, keys restricted to the query's graph.
//
// The N x N scores are never formed in memory. Work unit = (128 query rows) x (64-key block) x (one head):
//   S  = Q_h K_h^T        tcgen05.mma kind::f16, M=128 N=64 K=16, A = Q (fp16 hi/lo planes resident in TMEM), B = K tile (smem)
//   P  = exp2(S - |dp| - m)  by the softmax warps: tcgen05.ld S -> registers -> packed fp32x2 math + MUFU.EX2 -> fp16 hi/lo planes
//                            written back IN PLACE over S (tcgen05.st), interleaved per 16-key K-step [hi(8 cols) | lo(8 cols)]
//   O_h += P V_h          tcgen05.mma M=128 N=16 K=16 x 4 K-steps, A = P (TMEM), B = V_h^T tile (smem), accumulator in TMEM
// All products use the 3-term split (hi*hi + lo*hi + hi*lo) -> fp32-class results (same arithmetic as the mma.sync kernel).
//
// CTA = 384 threads: warps 0-3 softmax group 0 (heads 0-3), warps 4-7 softmax group 1 (heads 4-7) — thread = query row, so the
// distance bias of a (row, key block) is computed once into registers and reused by the group's 4 heads; warps 8/9 issue the MMAs
// for group 0/1 (one elected thread each); warp 10 streams pre-swizzled K / V^T / position block images global -> smem with
// cp.async.bulk (TMA engine) through a 3-stage mbarrier ring. TMEM (512 columns): 4 S/P slots x 64 | O 8 x 16 | Q hi 8 x 8 | Q lo 8 x 8.
// Softmax uses a lazily updated row reference max: a Cauchy-Schwarz bound |q_i| max_j |k_j| per key block decides whether the
// explicit row-max pass (and an O/l rescale through tcgen05.ld/st) is needed at all; the common path is add, add, exp2, add, split.
// Roofline: MUFU-bound (one exp2 per score + one sqrt per (row,key) per head group = 1.25 MUFU/score at 16 lanes/clk/SM).
#include "tc05.cuh"
#include <cstdlib>

namespace dgdm {

constexpr int TC_BN = 64;                      // keys per block
constexpr int TC_QT = 128;                     // query rows per CTA
constexpr int TC_IMG = 66560;                  // bytes per key-block image (65 KB, keeps every stage 1024-aligned)
constexpr int TC_OFF_V = 32768;                // K: [plane 2][half 2][64 keys x 128 B]; V^T: [plane 2][head 8][16 dims x 128 B]
constexpr int TC_OFF_POS = 65536;              // kx[64], ky[64] (pre-scaled by log2e/T; +inf marks padded keys)
constexpr int TC_OFF_KMAX = 65536 + 512;       // max_j |k_j,h| per head (8 floats)
constexpr int TC_IMG_USED = 65536 + 512 + 32;  // bytes actually copied per block
#ifndef TC_STAGES_N
#define TC_STAGES_N 3
#endif
constexpr int TC_STAGES = TC_STAGES_N;   // 2 stages measure the same 35.65 ms (the ring is not the limiter); 3 kept for L2-latency slack
constexpr int TC_THREADS = 384;                 // 3 warpgroups: softmax 0, softmax 1, {MMA 0, MMA 1, loader, idle}
constexpr int TC_SMEM = TC_STAGES * TC_IMG + 1024 + 256;
constexpr float TC_TAU = 12.0f;                // lazy-max slack (log2 units): p <= 2^12 is far inside the fp16 range of the hi plane

// TMEM column map
__device__ __forceinline__ constexpr uint32_t tc_slot(int g, int s) { return (uint32_t)((2 * g + s) * 64); }
__device__ __forceinline__ constexpr uint32_t tc_ocol(int h) { return (uint32_t)(256 + 16 * h); }
__device__ __forceinline__ constexpr uint32_t tc_qhi(int h) { return (uint32_t)(384 + 8 * h); }
__device__ __forceinline__ constexpr uint32_t tc_qlo(int h) { return (uint32_t)(448 + 8 * h); }

struct AttnTcParams {
  const float* q_qkv; int ld_q; const float* q_pos;
  const int* qseg; const int* kseg; int num_seg; int Nq; int nsplit;
  const uint8_t* img; const int* blk_ptr;
  float qscale, pscale;
  float* Opart; float* Mpart; float* Lpart;
  int mma_sleep_ns;            // single-thread roles (MMA issuers, loader): 0 = parked mbarrier waits, > 0 = poll + nanosleep
};
// wait of a single-thread role
#define TC_ROLE_WAIT(bar, par) do { if (p.mma_sleep_ns > 0) mbar_wait_sleep((bar), (par), (uint32_t)p.mma_sleep_ns); else mbar_wait_parked((bar), (par)); } while (0)

// ---------------------------------------------------------------------------------------------- packed fp32x2 helpers
__device__ __forceinline__ unsigned long long pk2(float a, float b) {
  return (unsigned long long)__float_as_uint(a) | ((unsigned long long)__float_as_uint(b) << 32);
}
__device__ __forceinline__ float lo32(unsigned long long v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi32(unsigned long long v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ float tc_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 2^x on the FMA / integer pipes (no MUFU): round-to-nearest split x = n + f with the 1.5*2^23 magic constant, degree-5
// minimax polynomial for 2^f on [-0.5, 0.5] (max rel. error 2.1e-7, the class of ex2.approx), exponent inserted with one
// integer multiply-add. Packed: two elements per call. Inputs are clamped at -125 (results below 2^-125 flush towards 0).
__device__ __forceinline__ unsigned long long tc_ex2_poly2(unsigned long long x) {
  const float x0 = fmaxf(lo32(x), -125.0f), x1 = fmaxf(hi32(x), -125.0f);
  const unsigned long long xc = pk2(x0, x1);
  const unsigned long long magic = pk2(12582912.0f, 12582912.0f);
  const unsigned long long t = add2(xc, magic);
  const unsigned long long f = sub2(xc, sub2(t, magic));
  unsigned long long q = fma2(pk2(0.0013276466634124517f, 0.0013276466634124517f), f, pk2(0.009675540961325169f, 0.009675540961325169f));
  q = fma2(q, f, pk2(0.05550713464617729f, 0.05550713464617729f));
  q = fma2(q, f, pk2(0.24022120237350464f, 0.24022120237350464f));
  q = fma2(q, f, pk2(0.6931469440460205f, 0.6931469440460205f));
  q = fma2(q, f, pk2(1.0000001192092896f, 1.0000001192092896f));
  const uint32_t r0 = (uint32_t)q + ((uint32_t)t << 23), r1 = (uint32_t)(q >> 32) + ((uint32_t)(t >> 32) << 23);
  return (unsigned long long)r0 | ((unsigned long long)r1 << 32);
}
#ifndef TC_POLY_EVERY
// Every TC_POLY_EVERY-th element pair takes the polynomial path (0 = never; T2_POLY is the same knob for the v2 kernel, measured
// 33.30 ms at 8 and 34.28 ms at 4 vs 33.16 ms at 0). MEASURED on B200 at 100k nodes (tools/attn_exp.sh):
// 0 -> 35.2 ms, 4 -> 38.5 ms, 3 -> 38.9 ms, 2 -> 43.9 ms: the softmax warps are issue/latency-bound (2 warps per scheduler,
// 216 registers each), so trading one MUFU op for ~6 FMA/ALU issue slots loses; the default therefore stays 0.
#define TC_POLY_EVERY 0
#endif
__device__ __forceinline__ float tc_ex2v(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#ifndef T2_DIST
#define T2_DIST 8
#endif
#ifndef T2_ROLL
#define T2_ROLL 0          // 1 / 2: head loop rolled / unrolled twice (less code) — measured 37.9 / 36.1 ms vs 33.1 ms fully unrolled
#endif
#ifndef T2_WATCH
// 1: the fourth role warp polls the S_ready barriers and publishes a unit counter (st.release / ld.acquire on shared memory), so that
// the softmax warps replace their one barrier test per unit (~180 cycles, 7 % of all samples) by a load. Correct (attention tests
// green) but MEASURED NOT FASTER: 33.52 vs 33.18 ms — the stall overlaps the scheduler's other warp; off by default.
#define T2_WATCH 0
#endif
#ifndef T2_JUNROLL
#define T2_JUNROLL 1       // 2: key-block loop unrolled twice — measured 34.08 vs 33.16 ms (instruction-cache misses, ncu no_instruction 0.5)
#endif
#ifndef T2_PIPE
// 1: chunk-level TMEM pipeline (every read issued one chunk before its wait, P_ready of unit u arrived after the first chunk of
// unit u + 1). MEASURED SLOWER: 35.8 vs 33.06 ms at 100k nodes (ptxas merges the two chunks into one schedule and sinks the TMEM
// accesses; with scheduling fences between the chunks it reaches 32.8 ms in the instrumented build) -> the unit-level order stays.
#define T2_PIPE 0
#endif
#ifndef T2_POLY
#define T2_POLY 0
#endif
#ifndef T2_ABL
#define T2_ABL 0           // timing ablations (WRONG results): 1 no exp2, 2 pass-through softmax, 3 no distance
#endif
__device__ __forceinline__ float tc_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t tc_pack_h2(float a, float b) {       // a -> low half, b -> high half
  uint32_t r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r;
}
__device__ __forceinline__ void split_pack(float a, float b, uint32_t& hi, uint32_t& lo) {   // round-to-nearest hi, residual lo
  const __half2 h = __floats2half2_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float2 hf = __half22float2(h);
  lo = tc_pack_h2(a - hf.x, b - hf.y);
}

// ---------------------------------------------------------------------------------------------- block-image builder
__global__ void attn_tc_blkptr_kernel(const int* __restrict__ kseg, int num_seg, int* __restrict__ blk_ptr) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int acc = 0;
    blk_ptr[0] = 0;
    for (int g = 0; g < num_seg; ++g) { acc += (kseg[g + 1] - kseg[g] + TC_BN - 1) / TC_BN; blk_ptr[g + 1] = acc; }
  }
}

// One CTA per 64-key block: fp32 k|v rows -> fp16 hi/lo planes in the exact shared-memory image the MMA descriptors expect
// (K-major SWIZZLE_128B tiles), V transposed per head, plus scaled key positions and the per-head max key norm.
__global__ void __launch_bounds__(256) attn_tc_prep_kernel(const float* __restrict__ kv, int ld, const float* __restrict__ pos,
                                                           const int* __restrict__ kseg, const int* __restrict__ blk_ptr, int num_seg,
                                                           float pscale, uint8_t* __restrict__ img) {
  __shared__ float vt[TC_BN][129];
  __shared__ float ksq[TC_BN][8];
  const int gb = blockIdx.x;
  if (gb >= blk_ptr[num_seg]) return;
  int lo = 0, hi = num_seg - 1;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (blk_ptr[mid] <= gb) lo = mid; else hi = mid - 1; }
  const int key0 = kseg[lo] + (gb - blk_ptr[lo]) * TC_BN;
  const int nvalid = min(TC_BN, kseg[lo + 1] - key0);
  uint8_t* out = img + (size_t)gb * TC_IMG;
  const int tid = threadIdx.x;
  constexpr int D = 128;
  // ---- K planes (and squared head norms)
  for (int idx = tid; idx < TC_BN * 16; idx += 256) {
    const int r = idx >> 4, c16 = idx & 15;
    float v[8];
    if (r < nvalid) {
      const float4 a = ldg4(kv + (size_t)(key0 + r) * ld + D + c16 * 8), b = ldg4(kv + (size_t)(key0 + r) * ld + D + c16 * 8 + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    uint32_t h[4], l[4];
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { split_pack(v[2 * i], v[2 * i + 1], h[i], l[i]); sq += v[2 * i] * v[2 * i] + v[2 * i + 1] * v[2 * i + 1]; }
    const uint32_t off = (uint32_t)((c16 >> 3) * 8192) + tc::sw128_off(r, c16 & 7);
    *reinterpret_cast<uint4*>(out + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(out + 16384 + off) = make_uint4(l[0], l[1], l[2], l[3]);
    // head = c16 / 2; the two 8-dim chunks of a head are handled by neighbouring lanes
    const float other = __shfl_xor_sync(0xffffffffu, sq, 1);
    if ((c16 & 1) == 0) ksq[r][c16 >> 1] = sq + other;
  }
  // ---- V tile -> smem (coalesced), zero rows for padded keys
  for (int idx = tid; idx < TC_BN * 32; idx += 256) {
    const int r = idx >> 5, c4 = idx & 31;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nvalid) a = ldg4(kv + (size_t)(key0 + r) * ld + 2 * D + c4 * 4);
    vt[r][c4 * 4] = a.x; vt[r][c4 * 4 + 1] = a.y; vt[r][c4 * 4 + 2] = a.z; vt[r][c4 * 4 + 3] = a.w;
  }
  __syncthreads();
  // ---- V^T planes: (head, dim) rows of 64 keys
  for (int idx = tid; idx < 8 * 16 * 8; idx += 256) {
    const int kc = idx & 7, d = (idx >> 3) & 15, h = idx >> 7;
    uint32_t hh[4], ll[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_pack(vt[8 * kc + 2 * i][16 * h + d], vt[8 * kc + 2 * i + 1][16 * h + d], hh[i], ll[i]);
    const uint32_t off = TC_OFF_V + (uint32_t)(h * 2048) + tc::sw128_off(d, kc);
    *reinterpret_cast<uint4*>(out + off) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    *reinterpret_cast<uint4*>(out + 16384 + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
  }
  // ---- key positions (scaled) and per-head max key norm
  if (tid < TC_BN) {
    float kx = INFINITY, ky = 0.f;                       // +inf => distance inf => bias -inf => p = 0 for padded keys
    if (tid < nvalid) { kx = pos ? pos[2 * (size_t)(key0 + tid)] * pscale : 0.f; ky = pos ? pos[2 * (size_t)(key0 + tid) + 1] * pscale : 0.f; }
    reinterpret_cast<float*>(out + TC_OFF_POS)[tid] = kx;
    reinterpret_cast<float*>(out + TC_OFF_POS)[TC_BN + tid] = ky;
  }
  if (tid >= 64 && tid < 72) {
    const int h = tid - 64;
    float mx = 0.f;
    for (int r = 0; r < TC_BN; ++r) mx = fmaxf(mx, ksq[r][h]);
    reinterpret_cast<float*>(out + TC_OFF_KMAX)[h] = sqrtf(mx) * 1.0001f;
  }
}

// ---------------------------------------------------------------------------------------------- main kernel
__global__ void __launch_bounds__(TC_THREADS, 1) spatial_attn_tc_kernel(const AttnTcParams p) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + TC_STAGES * TC_IMG;
  // barrier map (8 B each): full[3] @0, empty[3] @24, S_ready[g][s] @48, P_ready[g][s] @80, O_done[g] @112, Q_ready[g] @128
  const uint32_t bar_full = bars, bar_empty = bars + 24, bar_S = bars + 48, bar_P = bars + 80, bar_O = bars + 112, bar_Q = bars + 128;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bp + TC_STAGES * TC_IMG + 192);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- locate this CTA's (segment, query tile)
  int seg = -1, row0 = 0;
  {
    int acc = 0;
    for (int s = 0; s < p.num_seg; ++s) {
      const int a = p.qseg[s], b = p.qseg[s + 1];
      const int nt = (b - a + TC_QT - 1) / TC_QT;
      if ((int)blockIdx.x < acc + nt) { seg = s; row0 = a + ((int)blockIdx.x - acc) * TC_QT; break; }
      acc += nt;
    }
  }
  if (seg < 0) return;
  const int qend = p.qseg[seg + 1];
  const int split = blockIdx.y;
  const int nb_seg = p.blk_ptr[seg + 1] - p.blk_ptr[seg];
  const int per = (nb_seg + p.nsplit - 1) / p.nsplit;
  const int b_lo = min(nb_seg, split * per), b_hi = min(nb_seg, b_lo + per);
  const int nblk = b_hi - b_lo;
  const uint8_t* img0 = p.img + (size_t)(p.blk_ptr[seg] + b_lo) * TC_IMG;

  if (tid == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 2 + 8); }
    for (int i = 0; i < 4; ++i) { mbar_init(bar_S + 8 * i, 1); mbar_init(bar_P + 8 * i, 4); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_O + 8 * i, 1); mbar_init(bar_Q + 8 * i, 4); }
    mbar_fence_init();
  }
  if (warp == 10) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tm = *tmem_slot;

  if (warp < 8) {
    // =========================================================================== softmax groups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");      // registers handed over by the MMA / loader warpgroup
    const int g = warp >> 2, wq = warp & 3;
    const int row = row0 + wq * 32 + lane;
    const bool rvalid = row < qend;
    const uint32_t trow = tm + ((uint32_t)(wq * 32) << 16);
    float qn[4];
    // ---- Q: fp32 row -> scaled -> fp16 hi/lo -> TMEM (A operand of QK^T), and |q_h| for the lazy-max bound
    {
      const float* qp = p.q_qkv + (size_t)(rvalid ? row : row0) * p.ld_q + 64 * g;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t qh[8], ql[8];
        float sq = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float4 v = ldg4(qp + 16 * i + 4 * c);
          if (!rvalid) v = make_float4(0.f, 0.f, 0.f, 0.f);
          v.x *= p.qscale; v.y *= p.qscale; v.z *= p.qscale; v.w *= p.qscale;
          sq += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          split_pack(v.x, v.y, qh[2 * c], ql[2 * c]);
          split_pack(v.z, v.w, qh[2 * c + 1], ql[2 * c + 1]);
        }
        qn[i] = sqrtf(sq) * 1.0001f;
        tmem_st8_nowait(trow + tc_qhi(4 * g + i), qh);
        tmem_st8_nowait(trow + tc_qlo(4 * g + i), ql);
      }
      tmem_wait_st();
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_Q + 8 * g);
    }
    float qx = 0.f, qy = 0.f;
    if (p.q_pos != nullptr && rvalid) { qx = p.q_pos[2 * (size_t)row] * p.pscale; qy = p.q_pos[2 * (size_t)row + 1] * p.pscale; }
    const unsigned long long qx2 = pk2(qx, qx), qy2 = pk2(qy, qy);

    float m[4];
    unsigned long long lsum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; lsum[i] = 0ull; }

    // S of the unit being processed lives in registers (ra = keys 0-31, rb = keys 32-63 of the block); the loads for the NEXT
    // unit are issued while the current one is still in the exp2 / split phase, so neither the S_ready wait nor the TMEM
    // read latency sits on the critical path.
    uint32_t ra[32], rb[32];
    const int total = 4 * nblk;
    if (nblk > 0) {
      mbar_wait_parked(bar_S + 8 * (2 * g), 0);
      fence_after();
      tmem_ld32_nowait(trow + tc_slot(g, 0), ra);
      tmem_ld32_nowait(trow + tc_slot(g, 0) + 32, rb);
    }
    const unsigned long long neg1 = pk2(-1.0f, -1.0f);
    // exp2 + fp16 hi/lo split of one 32-key chunk held in r (in place), accumulating the row sum
    auto softmax_chunk = [&](uint32_t* r, const unsigned long long* nbc, unsigned long long nmu2, unsigned long long& ls) {
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int e = 16 * ks + 2 * t;                              // element pair (e, e+1) of this 32-key chunk
          const unsigned long long x = add2(add2(pk2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), nbc[8 * ks + t]), nmu2);
          float p0, p1;
          unsigned long long pp;
          if (TC_POLY_EVERY > 0 && (t % (TC_POLY_EVERY > 0 ? TC_POLY_EVERY : 1)) == (TC_POLY_EVERY > 0 ? TC_POLY_EVERY - 1 : 0)) {
            pp = tc_ex2_poly2(x); p0 = lo32(pp); p1 = hi32(pp);
          } else {
            p0 = tc_ex2(lo32(x)); p1 = tc_ex2(hi32(x)); pp = pk2(p0, p1);
          }
          ls = add2(ls, pp);
          const float h0 = __uint_as_float(__float_as_uint(p0) & 0xffffe000u), h1 = __uint_as_float(__float_as_uint(p1) & 0xffffe000u);
          const unsigned long long ll = fma2(pk2(h0, h1), neg1, pp);  // p - hi, exact
          hi[t] = tc_pack_h2(h0, h1);
          lo[t] = tc_pack_h2(lo32(ll), hi32(ll));
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) { r[16 * ks + t] = hi[t]; r[16 * ks + 8 + t] = lo[t]; }
      }
    };

    for (int j = 0; j < nblk; ++j) {
      const int st = j % TC_STAGES;
      mbar_wait_parked(bar_full + 8 * st, (uint32_t)((j / TC_STAGES) & 1));
      // ---- negated distance bias of this row against the block's 64 keys (shared by the group's 4 heads)
      unsigned long long nb[32];
      float kmax[4];
      {
        const float* kp = reinterpret_cast<const float*>(bp + st * TC_IMG + TC_OFF_POS);
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const float4 kx = *reinterpret_cast<const float4*>(kp + 4 * t);
          const float4 ky = *reinterpret_cast<const float4*>(kp + TC_BN + 4 * t);
          unsigned long long dx = sub2(qx2, pk2(kx.x, kx.y)), dy = sub2(qy2, pk2(ky.x, ky.y));
          unsigned long long d2 = fma2(dx, dx, mul2(dy, dy));
          nb[2 * t] = pk2(-tc_sqrt(lo32(d2)), -tc_sqrt(hi32(d2)));
          dx = sub2(qx2, pk2(kx.z, kx.w)); dy = sub2(qy2, pk2(ky.z, ky.w));
          d2 = fma2(dx, dx, mul2(dy, dy));
          nb[2 * t + 1] = pk2(-tc_sqrt(lo32(d2)), -tc_sqrt(hi32(d2)));
        }
        const float4 km = *reinterpret_cast<const float4*>(bp + st * TC_IMG + TC_OFF_KMAX + 16 * g);
        kmax[0] = km.x; kmax[1] = km.y; kmax[2] = km.z; kmax[3] = km.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + 8 * st);            // this warp no longer reads the stage's position data

#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int u = 4 * j + i, s = u & 1;
        const uint32_t slot = trow + tc_slot(g, s);
        const uint32_t nslot = trow + tc_slot(g, s ^ 1);
        tmem_wait_ld();                                           // ra / rb of this unit (issued one unit ago)
        // ---- lazy reference max: only when the Cauchy-Schwarz bound says this block could overflow the current reference
        const bool need = (j == 0) || (qn[i] * kmax[i] > m[i] + TC_TAU);
        if (__any_sync(0xffffffffu, need)) {
          float rmax = -INFINITY;
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            const unsigned long long xa = add2(pk2(__uint_as_float(ra[2 * t]), __uint_as_float(ra[2 * t + 1])), nb[t]);
            const unsigned long long xb = add2(pk2(__uint_as_float(rb[2 * t]), __uint_as_float(rb[2 * t + 1])), nb[16 + t]);
            rmax = fmaxf(rmax, fmaxf(fmaxf(lo32(xa), hi32(xa)), fmaxf(lo32(xb), hi32(xb))));
          }
          float sc = 1.0f;
          if (j == 0) m[i] = rmax;
          else if (rmax > m[i] + TC_TAU) { sc = tc_ex2(m[i] - rmax); m[i] = rmax; }
          if (__any_sync(0xffffffffu, sc != 1.0f)) {
            // rescale O_h and l of this row; every PV MMA on O_h issued so far has completed (in-order MMA pipe, S_ready waited)
            uint32_t o[16];
            tmem_ld16_nowait(trow + tc_ocol(4 * g + i), o);
            tmem_wait_ld();
#pragma unroll
            for (int t = 0; t < 16; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * sc);
            tmem_st16_nowait(trow + tc_ocol(4 * g + i), o);
            lsum[i] = mul2(lsum[i], pk2(sc, sc));
          }
        }
        const float mu = (m[i] == -INFINITY) ? 0.f : m[i];
        const unsigned long long nmu2 = pk2(-mu, -mu);
        softmax_chunk(ra, nb, nmu2, lsum[i]);
        tmem_st32_nowait(slot, ra);
        if (u + 1 < total) {                                      // prefetch the first half of the next unit's S into ra
          mbar_wait_parked(bar_S + 8 * (2 * g + (s ^ 1)), (uint32_t)(((u + 1) >> 1) & 1));
          fence_after();
          tmem_ld32_nowait(nslot, ra);
        }
        softmax_chunk(rb, nb + 16, nmu2, lsum[i]);
        tmem_st32_nowait(slot + 32, rb);
        tmem_wait_st();
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_P + 8 * (2 * g + s));
        if (u + 1 < total) tmem_ld32_nowait(nslot + 32, rb);      // second half of the next unit
      }
    }
    // ---- epilogue: unnormalised O, reference max m, row sum l of this key split
    if (nblk > 0) { mbar_wait_parked(bar_O + 8 * g, 0); fence_after(); }
    const size_t prow = (size_t)split * p.Nq + (size_t)row;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      if (nblk > 0) { tmem_ld32_nowait(trow + tc_ocol(4 * g) + 32 * c, r); tmem_wait_ld(); }
      else {
#pragma unroll
        for (int t = 0; t < 32; ++t) r[t] = 0u;
      }
      if (rvalid) {
        float* op = p.Opart + prow * 128 + 64 * g + 32 * c;
#pragma unroll
        for (int t = 0; t < 8; ++t)
          *reinterpret_cast<float4*>(op + 4 * t) = make_float4(__uint_as_float(r[4 * t]), __uint_as_float(r[4 * t + 1]),
                                                               __uint_as_float(r[4 * t + 2]), __uint_as_float(r[4 * t + 3]));
      }
    }
    if (rvalid) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        p.Mpart[prow * 8 + 4 * g + i] = m[i];
        p.Lpart[prow * 8 + 4 * g + i] = lo32(lsum[i]) + hi32(lsum[i]);
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp < 10) {
    // =========================================================================== MMA issuers (one per head group)
    const int g = warp - 8;
    if (nblk > 0 && elect_one()) {
      const uint32_t id64 = idesc_f16(128, 64), id16 = idesc_f16(128, 16);
      const int total = 4 * nblk;
      TC_ROLE_WAIT(bar_Q + 8 * g, 0);
      fence_after();
      auto issue_qk = [&](int u) {
        const int b = u >> 2, i = u & 3, st = b % TC_STAGES;
        if (i == 0) { TC_ROLE_WAIT(bar_full + 8 * st, (uint32_t)((b / TC_STAGES) & 1)); fence_after(); }
        const uint32_t d = tm + tc_slot(g, u & 1);
        const uint32_t kb = base + st * TC_IMG + 8192 * g + 32 * i;
        mma_ts(d, tm + tc_qhi(4 * g + i), desc_sw128(kb), id64, 0);
        mma_ts(d, tm + tc_qlo(4 * g + i), desc_sw128(kb), id64, 1);
        mma_ts(d, tm + tc_qhi(4 * g + i), desc_sw128(kb + 16384), id64, 1);
        mma_commit(bar_S + 8 * (2 * g + (u & 1)));
      };
      issue_qk(0);
      if (total > 1) issue_qk(1);
      for (int u = 0; u < total; ++u) {
        const int b = u >> 2, i = u & 3, st = b % TC_STAGES, h = 4 * g + i;
        TC_ROLE_WAIT(bar_P + 8 * (2 * g + (u & 1)), (uint32_t)((u >> 1) & 1));
        fence_after();
        const uint32_t vb = base + st * TC_IMG + TC_OFF_V + 2048 * h;
        const uint32_t ps = tm + tc_slot(g, u & 1);
        const uint32_t od = tm + tc_ocol(h);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          mma_ts(od, ps + 16 * ks, desc_sw128(vb + 32 * ks), id16, (b > 0 || ks > 0) ? 1u : 0u);
          mma_ts(od, ps + 16 * ks + 8, desc_sw128(vb + 32 * ks), id16, 1);
          mma_ts(od, ps + 16 * ks, desc_sw128(vb + 16384 + 32 * ks), id16, 1);
        }
        if (i == 3) mma_commit(bar_empty + 8 * st);                  // every MMA reading this stage has been issued
        if (u + 2 < total) issue_qk(u + 2);
      }
      mma_commit(bar_O + 8 * g);
    }
    __syncwarp();
    } else if (warp == 10) {
    // =========================================================================== loader (cp.async.bulk, 3-stage ring)
    if (elect_one()) {
      for (int j = 0; j < nblk; ++j) {
        const int st = j % TC_STAGES;
        TC_ROLE_WAIT(bar_empty + 8 * st, (uint32_t)(((j / TC_STAGES) & 1) ^ 1));
        mbar_arrive_expect_tx(bar_full + 8 * st, TC_IMG_USED);
        const uint8_t* src = img0 + (size_t)j * TC_IMG;
        const uint32_t dst = base + st * TC_IMG;
#pragma unroll
        for (int c = 0; c < 4; ++c) bulk_g2s(dst + 16384 * c, src + 16384 * c, 16384, bar_full + 8 * st);
        bulk_g2s(dst + 65536, src + 65536, TC_IMG_USED - 65536, bar_full + 8 * st);
      }
    }
    __syncwarp();
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 10) tmem_dealloc(tm, 512);
}


// ---------------------------------------------------------------------------------------------- v2: restructured around measured latencies
// Same work decomposition, data images and 3-term split arithmetic as spatial_attn_tc_kernel. In-kernel cycle accounting
// (tools/attn_prof.py) and timing ablations (T2_ABL) showed what the ncu counters could not: the exp2 / split chunks already ran at
// ~90 % of the XU bound, and the rest of a unit's ~1950 cycles went to (a) ~350 cycles parked on S_ready of the next unit — its
// QK^T MMAs could only be issued after the PV MMAs that free the in-place P slot, ~1000 cycles after P_ready — and (b) barrier tests
// that cost ~180 cycles even when the phase is already complete. v2 therefore:
//   * keeps THREE S/P slots per head group, so QK^T of unit u + 3 is issued right behind PV of unit u and S is ready two units ahead;
//     the TMEM columns come from moving Q to shared memory (SS-mode MMAs on fp16 hi/lo planes in the SWIZZLE_128B layout of the K
//     tiles; slower per MMA than TMEM operands, but off the critical path now);
//   * folds the reference maximum into the MMA:  S = Q_h K_h^T - m_h.  A fourth MMA per unit multiplies a [128 x 16] "extension" A
//     tile, whose columns 4i..4i+2 hold -m of head i as three fp16 pieces (exact to 2^-33), with a constant 0/1 selector B tile.
//     The softmax thread rewrites its pieces only when its lazy reference changes (first block, then rarely), so the common path no
//     longer subtracts m per score; the distance is kept positive and subtracted (the negation folds into the packed add);
//   * fetches S of the next unit at the END of a unit and tests the barriers one chunk before their results are needed.
// Per two scores: 2 FADD2, 2 MUFU.EX2, 2 LOP3, 1 FFMA2, 2 F2FP (9 instead of 10). TMEM (512 columns): 6 S/P slots x 64 | O 8 x 16.
// Shared memory: K/V ring 2 x 65 KB (2 stages measured equal to 3 in round 2) | Q 64 KB | extension 16 KB | selector 8 KB.
// (A variant that also computed the row sums on the tensor pipe — ones-tile MMAs into 16-column accumulators, 8 instructions per two
// scores — measured 33.7 ms with two slots; it needs the TMEM columns of the third slot and lengthens the MMA chain, see git history.)
constexpr int T2_STAGES = 2;
constexpr int T2_OFF_Q = T2_STAGES * TC_IMG;            // [plane 2][group 2][128 rows x 128 B]
constexpr int T2_OFF_EXT = T2_OFF_Q + 65536;            // [128 rows x 128 B]; group g uses the 16-column K slice g
constexpr int T2_OFF_SEL = T2_OFF_EXT + 16384;          // [64 keys x 128 B]; K slice i selects head i's extension columns
constexpr int T2_OFF_BARS = T2_OFF_SEL + 8192;
constexpr int T2_SMEM = T2_OFF_BARS + 256 + 1024;
__device__ __forceinline__ uint32_t t2_slot(int g, int k) { return (uint32_t)((3 * g + k) * 64); }     // k = unit % 3
__device__ __forceinline__ constexpr uint32_t t2_ocol(int h) { return (uint32_t)(384 + 16 * h); }

#ifndef T2_PROF
#define T2_PROF 0          // 1: cycle accounting of one softmax warp and one MMA issuer of CTA (5, 1) into g_t2_prof (dev builds only)
#endif
#if T2_PROF
__device__ long long g_t2_prof[32];
#define T2_TICK(idx) do { if (prof_on) { const long long t__ = clock64(); pc[idx] += t__ - tprev; tprev = t__; } } while (0)
#else
#define T2_TICK(idx) do { } while (0)
#endif
__global__ void __launch_bounds__(TC_THREADS, 1) spatial_attn_tc2_kernel(const AttnTcParams p) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + T2_OFF_BARS;
  // barrier map (8 B each): full[2] @0, empty[2] @16, S_ready[g][k] @32, P_ready[g][k] @80, O_done[g] @128, Q_ready[g] @144
  const uint32_t bar_full = bars, bar_empty = bars + 16, bar_S = bars + 32, bar_P = bars + 80, bar_O = bars + 128, bar_Q = bars + 144;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bp + T2_OFF_BARS + 192);
  // watcher words (T2_WATCH): [0], [1] = number of units of group 0 / 1 whose S_ready the watcher warp has observed, [2] = softmax
  // warps that have left their loop
  const uint32_t wflag = bars + 168;
  (void)wflag;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int seg = -1, row0 = 0;
  {
    int acc = 0;
    for (int s = 0; s < p.num_seg; ++s) {
      const int a = p.qseg[s], b = p.qseg[s + 1];
      const int nt = (b - a + TC_QT - 1) / TC_QT;
      if ((int)blockIdx.x < acc + nt) { seg = s; row0 = a + ((int)blockIdx.x - acc) * TC_QT; break; }
      acc += nt;
    }
  }
  if (seg < 0) return;
  const int qend = p.qseg[seg + 1];
  const int split = blockIdx.y;
  const int nb_seg = p.blk_ptr[seg + 1] - p.blk_ptr[seg];
  const int per = (nb_seg + p.nsplit - 1) / p.nsplit;
  const int b_lo = min(nb_seg, split * per), b_hi = min(nb_seg, b_lo + per);
  const int nblk = b_hi - b_lo;
  const uint8_t* img0 = p.img + (size_t)(p.blk_ptr[seg] + b_lo) * TC_IMG;

  if (tid == 0) {
    for (int i = 0; i < T2_STAGES; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 2 + 8); }
    for (int i = 0; i < 6; ++i) { mbar_init(bar_S + 8 * i, 1); mbar_init(bar_P + 8 * i, 4); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_O + 8 * i, 1); mbar_init(bar_Q + 8 * i, 4); }
    mbar_fence_init();
#if T2_WATCH
    *reinterpret_cast<volatile uint32_t*>(bp + T2_OFF_BARS + 168) = 0u;
    *reinterpret_cast<volatile uint32_t*>(bp + T2_OFF_BARS + 172) = 0u;
    *reinterpret_cast<volatile uint32_t*>(bp + T2_OFF_BARS + 176) = 0u;
#endif
  }
  // constant selector B tile: element [key, 16 i + c] = 1 iff c / 4 == i and c % 4 < 3
  for (int idx = tid; idx < 64 * 8; idx += TC_THREADS) {
    const int k = idx >> 3, ch = idx & 7;                         // 16-byte chunk ch = columns 8 ch .. 8 ch + 7
    uint32_t w[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int c0 = 8 * ch + 2 * t, c1 = c0 + 1;
      const uint32_t v0 = ((((c0 & 15) >> 2) == (c0 >> 4)) && ((c0 & 3) < 3)) ? 0x3C00u : 0u;
      const uint32_t v1 = ((((c1 & 15) >> 2) == (c1 >> 4)) && ((c1 & 3) < 3)) ? 0x3C00u : 0u;
      w[t] = v0 | (v1 << 16);
    }
    *reinterpret_cast<uint4*>(bp + T2_OFF_SEL + sw128_off(k, ch)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  fence_async_smem();
  if (warp == 10) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tm = *tmem_slot;

  if (warp < 8) {
    // =========================================================================== softmax groups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int g = warp >> 2, wq = warp & 3;
    const int rl = wq * 32 + lane;
    const int row = row0 + rl;
    const bool rvalid = row < qend;
    const uint32_t trow = tm + ((uint32_t)(wq * 32) << 16);
    uint8_t* ext_row = bp + T2_OFF_EXT;                          // this row's extension chunks: 2 g (heads 0, 1), 2 g + 1 (heads 2, 3)
    float qn[4];
    // ---- Q: fp32 row -> scaled -> fp16 hi/lo planes in shared memory (A operand of QK^T), |q_h| for the lazy-max bound
    {
      const float* qp = p.q_qkv + (size_t)(rvalid ? row : row0) * p.ld_q + 64 * g;
      uint8_t* qhi = bp + T2_OFF_Q + g * 16384;
      uint8_t* qlo = qhi + 32768;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t qh[8], ql[8];
        float sq = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float4 v = ldg4(qp + 16 * i + 4 * c);
          if (!rvalid) v = make_float4(0.f, 0.f, 0.f, 0.f);
          v.x *= p.qscale; v.y *= p.qscale; v.z *= p.qscale; v.w *= p.qscale;
          sq += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          split_pack(v.x, v.y, qh[2 * c], ql[2 * c]);
          split_pack(v.z, v.w, qh[2 * c + 1], ql[2 * c + 1]);
        }
        qn[i] = sqrtf(sq) * 1.0001f;
        *reinterpret_cast<uint4*>(qhi + sw128_off(rl, 2 * i)) = make_uint4(qh[0], qh[1], qh[2], qh[3]);
        *reinterpret_cast<uint4*>(qhi + sw128_off(rl, 2 * i + 1)) = make_uint4(qh[4], qh[5], qh[6], qh[7]);
        *reinterpret_cast<uint4*>(qlo + sw128_off(rl, 2 * i)) = make_uint4(ql[0], ql[1], ql[2], ql[3]);
        *reinterpret_cast<uint4*>(qlo + sw128_off(rl, 2 * i + 1)) = make_uint4(ql[4], ql[5], ql[6], ql[7]);
      }
      *reinterpret_cast<uint4*>(ext_row + sw128_off(rl, 2 * g)) = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(ext_row + sw128_off(rl, 2 * g + 1)) = make_uint4(0u, 0u, 0u, 0u);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_Q + 8 * g);
    }
    float qx = 0.f, qy = 0.f;
    if (p.q_pos != nullptr && rvalid) { qx = p.q_pos[2 * (size_t)row] * p.pscale; qy = p.q_pos[2 * (size_t)row + 1] * p.pscale; }
    const unsigned long long qx2 = pk2(qx, qx), qy2 = pk2(qy, qy);

    float m[4];                                                   // reference maximum folded into S from the head's NEXT unit on
    unsigned long long lsum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; lsum[i] = 0ull; }

    uint32_t ra[32], rb[32];
    const int total = 4 * nblk;
    if (nblk > 0) {
      mbar_wait_parked(bar_S + 8 * (3 * g), 0);
      fence_after();
      tmem_ld32_nowait(trow + t2_slot(g, 0), ra);
#if !T2_PIPE
      tmem_ld32_nowait(trow + t2_slot(g, 0) + 32, rb);
#endif
    }
    const unsigned long long neg1 = pk2(-1.0f, -1.0f);
    // exp2 + fp16 hi/lo split of one 32-key chunk held in r (in place); r already carries q.k - m, dc the (scaled) distances.
    // Written as a software pipeline (the MUFU.EX2 of pair t + T2_DIST before the split of pair t). ptxas re-schedules the chunk
    // anyway: it paces the MUFUs with 8-cycle issue stalls and places each dependent LOP3 two to three MUFUs behind its producer
    // (six scoreboards), so the source order only fixes which registers are live; kept because it is the order that was measured.
    auto softmax_chunk = [&](uint32_t* r, const unsigned long long* dc, unsigned long long& ls) {
      constexpr int D = T2_DIST;
      float pv[32];
      auto ex = [&](int t) {                                      // pair t = elements (2t, 2t+1) of the chunk
        const unsigned long long x = sub2(pk2(__uint_as_float(r[2 * t]), __uint_as_float(r[2 * t + 1])), dc[t]);
#if T2_ABL == 1
        pv[2 * t] = lo32(x); pv[2 * t + 1] = hi32(x);
#else
        if (T2_POLY > 0 && (t % (T2_POLY > 0 ? T2_POLY : 1)) == (T2_POLY > 0 ? T2_POLY - 1 : 0)) {
          const unsigned long long pp = tc_ex2_poly2(x);          // every T2_POLY-th pair on the FMA / integer pipes instead of the XU
          pv[2 * t] = lo32(pp); pv[2 * t + 1] = hi32(pp);
        } else {
          pv[2 * t] = tc_ex2v(lo32(x)); pv[2 * t + 1] = tc_ex2v(hi32(x));
        }
#endif
      };
#if T2_ABL == 2 || T2_ABL == 4
      if (dc != nullptr) return;
#endif
      uint32_t hi[16], lo[16];
      auto sp = [&](int t) {
        const float p0 = pv[2 * t], p1 = pv[2 * t + 1];
        const float h0 = __uint_as_float(__float_as_uint(p0) & 0xffffe000u), h1 = __uint_as_float(__float_as_uint(p1) & 0xffffe000u);
        const unsigned long long pp = pk2(p0, p1);
        ls = add2(ls, pp);
        const unsigned long long ll = fma2(pk2(h0, h1), neg1, pp);
        hi[t] = tc_pack_h2(h0, h1);
        lo[t] = tc_pack_h2(lo32(ll), hi32(ll));
      };
#pragma unroll
      for (int t = 0; t < D; ++t) ex(t);
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        if (t + D < 16) ex(t + D);
        sp(t);
      }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int t = 0; t < 8; ++t) { r[16 * ks + t] = hi[8 * ks + t]; r[16 * ks + 8 + t] = lo[8 * ks + t]; }
      }
    };

#if T2_PROF
    const bool prof_on = (blockIdx.x == 5 && blockIdx.y == 1 && warp == 0);
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
    const long long tstart = tprev;
#endif
    bool f_ok = false;                                           // early test of the next block's full barrier succeeded
#if T2_JUNROLL == 2
#pragma unroll 2
#endif
    for (int j = 0; j < nblk; ++j) {
      const int st = j % T2_STAGES;
      if (!f_ok) mbar_wait_parked(bar_full + 8 * st, (uint32_t)((j / T2_STAGES) & 1));
      f_ok = false;
      T2_TICK(0);
      // ---- scaled distance of this row to the block's 64 keys (shared by the group's 4 heads); +inf for padded keys
      unsigned long long db[32];
      float kmax[4];
      {
        const float* kp = reinterpret_cast<const float*>(bp + st * TC_IMG + TC_OFF_POS);
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const float4 kx = *reinterpret_cast<const float4*>(kp + 4 * t);
          const float4 ky = *reinterpret_cast<const float4*>(kp + TC_BN + 4 * t);
          unsigned long long dx = sub2(qx2, pk2(kx.x, kx.y)), dy = sub2(qy2, pk2(ky.x, ky.y));
          unsigned long long d2 = fma2(dx, dx, mul2(dy, dy));
#if T2_ABL == 3 || T2_ABL == 4
          db[2 * t] = d2;
#else
          db[2 * t] = pk2(tc_sqrt(lo32(d2)), tc_sqrt(hi32(d2)));
#endif
          dx = sub2(qx2, pk2(kx.z, kx.w)); dy = sub2(qy2, pk2(ky.z, ky.w));
          d2 = fma2(dx, dx, mul2(dy, dy));
#if T2_ABL == 3 || T2_ABL == 4
          db[2 * t + 1] = d2;
#else
          db[2 * t + 1] = pk2(tc_sqrt(lo32(d2)), tc_sqrt(hi32(d2)));
#endif
        }
        const float4 km = *reinterpret_cast<const float4*>(bp + st * TC_IMG + TC_OFF_KMAX + 16 * g);
        kmax[0] = km.x; kmax[1] = km.y; kmax[2] = km.z; kmax[3] = km.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + 8 * st);
      T2_TICK(1);

      float bound[4];                                            // Cauchy-Schwarz bound on q.k of this row against the block, per head
#pragma unroll
      for (int i = 0; i < 4; ++i) bound[i] = qn[i] * kmax[i];
#if T2_ROLL == 1
#pragma unroll 1
#elif T2_ROLL == 2
#pragma unroll 2
#else
#pragma unroll
#endif
      for (int i = 0; i < 4; ++i) {
        const int u = 4 * j + i, k3 = u % 3, k3n = (u + 1) % 3;     // S/P slot of this unit and of the next
        const uint32_t slot = trow + t2_slot(g, k3);
        const uint32_t nslot = trow + t2_slot(g, k3n);
        const uint32_t nbarS = bar_S + 8 * (3 * g + k3n), nparS = (uint32_t)(((u + 1) / 3) & 1);
        // head-indexed state through selects (constant register indices also when the loop is not unrolled)
        const float m_i = (i == 0) ? m[0] : (i == 1) ? m[1] : (i == 2) ? m[2] : m[3];
        const float bound_i = (i == 0) ? bound[0] : (i == 1) ? bound[1] : (i == 2) ? bound[2] : bound[3];
        tmem_wait_ld();
#if T2_PIPE
        // Chunk-level pipeline: every TMEM read is issued one chunk (half a unit) before its tcgen05.wait::ld and every store is
        // waited for one chunk after its issue, so neither latency is exposed (both waits cover ALL outstanding accesses of the
        // thread, hence the order: wait for the current half, THEN issue the read of the other half).
        tmem_ld32_nowait(slot + 32, rb);                          // second half of this unit; rb was released by the previous store
#endif
        T2_TICK(2);
        // ---- lazy reference: S already carries -m[i] (0 while m[i] = -inf). Only when the Cauchy-Schwarz bound admits an overflow
        // of the current reference is the block's row maximum looked at; a new reference shifts this unit's scores in registers,
        // rescales O_h / l_h and is published to the extension tile for the head's following units.
        const bool need = (m_i == -INFINITY) || (bound_i > m_i + TC_TAU);
        if (__any_sync(0xffffffffu, need)) {
#if T2_PROF
          if (prof_on) pc[7] += 1;                                // how often the row-maximum pass runs
#endif
#if T2_PIPE
          tmem_wait_ld();                                         // rare path: the row maximum needs both halves
#endif
          float rmax = -INFINITY;                                 // relative to the folded reference
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            const unsigned long long xa = sub2(pk2(__uint_as_float(ra[2 * t]), __uint_as_float(ra[2 * t + 1])), db[t]);
            const unsigned long long xb = sub2(pk2(__uint_as_float(rb[2 * t]), __uint_as_float(rb[2 * t + 1])), db[16 + t]);
            rmax = fmaxf(rmax, fmaxf(fmaxf(lo32(xa), hi32(xa)), fmaxf(lo32(xb), hi32(xb))));
          }
          // The new reference is what the tensor core will subtract: -m as three fp16 pieces (hi + mid + lo reproduces an fp32 value
          // to 2^-33; each piece is clamped to the finite fp16 range, because an infinite piece times the selector's zeros would put
          // NaNs into the OTHER heads' scores). m_new is read back from the pieces, so the bookkeeping is exact by construction.
          float delta = 0.f, sc = 1.0f, m_new = m_i;
          __half e0 = __float2half_rn(0.f), e1 = e0, e2 = e0;
          const bool first = (m_i == -INFINITY);
          if ((first && rmax > -INFINITY) || (!first && rmax > TC_TAU)) {
            const float m_old = first ? 0.f : m_i;                 // reference folded into this unit's S
            const float nm = -(m_old + rmax);
            e0 = __float2half_rn(fminf(fmaxf(nm, -65000.f), 65000.f));
            const float r1 = nm - __half2float(e0);
            e1 = __float2half_rn(fminf(fmaxf(r1, -65000.f), 65000.f));
            e2 = __float2half_rn(fminf(fmaxf(r1 - __half2float(e1), -65000.f), 65000.f));
            m_new = -((__half2float(e0) + __half2float(e1)) + __half2float(e2));
            delta = m_old - m_new;                                 // shift of this unit's scores (~ -rmax)
            if (!first) sc = tc_ex2(delta);                        // everything accumulated so far is 0 when first
          }
          if (i == 0) m[0] = m_new; else if (i == 1) m[1] = m_new; else if (i == 2) m[2] = m_new; else m[3] = m_new;
          if (__any_sync(0xffffffffu, sc != 1.0f)) {
            // every PV MMA on head i issued so far has completed: PV of unit u - 4 was issued before QK^T of unit u - 1, whose
            // S_ready this warp has waited for (in-order MMA pipe)
            uint32_t o[16];
            tmem_ld16_nowait(trow + t2_ocol(4 * g + i), o);
            tmem_wait_ld();
#pragma unroll
            for (int t = 0; t < 16; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * sc);
            tmem_st16_nowait(trow + t2_ocol(4 * g + i), o);
            const unsigned long long sc2 = pk2(sc, sc);
            if (i == 0) lsum[0] = mul2(lsum[0], sc2); else if (i == 1) lsum[1] = mul2(lsum[1], sc2);
            else if (i == 2) lsum[2] = mul2(lsum[2], sc2); else lsum[3] = mul2(lsum[3], sc2);
          }
          if (__any_sync(0xffffffffu, delta != 0.f)) {
            const unsigned long long d2 = pk2(delta, delta);
#pragma unroll
            for (int t = 0; t < 16; ++t) {
              const unsigned long long xa = add2(pk2(__uint_as_float(ra[2 * t]), __uint_as_float(ra[2 * t + 1])), d2);
              const unsigned long long xb = add2(pk2(__uint_as_float(rb[2 * t]), __uint_as_float(rb[2 * t + 1])), d2);
              ra[2 * t] = (uint32_t)xa; ra[2 * t + 1] = (uint32_t)(xa >> 32);
              rb[2 * t] = (uint32_t)xb; rb[2 * t + 1] = (uint32_t)(xb >> 32);
            }
            if (delta != 0.f) {
              // publish the pieces in columns 4 i .. 4 i + 2 of this row's slice of the extension tile
              const uint32_t w0 = (uint32_t)__half_as_ushort(e0) | ((uint32_t)__half_as_ushort(e1) << 16);
              const uint32_t w1 = (uint32_t)__half_as_ushort(e2);
              *reinterpret_cast<uint2*>(ext_row + sw128_off(rl, 2 * g + (i >> 1)) + 8 * (i & 1)) = make_uint2(w0, w1);
              fence_async_smem();
            }
          }
        }
        unsigned long long ls = (i == 0) ? lsum[0] : (i == 1) ? lsum[1] : (i == 2) ? lsum[2] : lsum[3];
        // A barrier test costs ~180 cycles even when the phase is complete (tools/tc_probe.cu): S_ready of the next unit and the
        // next block's full barrier are tested one chunk before the results are needed and only parked on if that test failed.
        bool s_ok = false;
#if T2_PIPE
        if (u + 1 < total) s_ok = mbar_test_wait(nbarS, nparS);
        if (i == 3 && j + 1 < nblk) f_ok = mbar_test_wait(bar_full + 8 * ((j + 1) % T2_STAGES), (uint32_t)(((j + 1) / T2_STAGES) & 1));
        softmax_chunk(ra, db, ls);
        if (u > 0) {                                              // P of the PREVIOUS unit: its last store was issued one chunk ago
          tmem_wait_st();
          fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_P + 8 * (3 * g + (k3 + 2) % 3));
        }
        tmem_st32_nowait(slot, ra);
        T2_TICK(3);
        tmem_wait_ld();                                           // second half of this unit (issued one chunk ago)
        if (u + 1 < total) {                                      // first half of the next unit; ra was released by the store above
          if (!s_ok) mbar_wait_parked(nbarS, nparS);
          fence_after();
          tmem_ld32_nowait(nslot, ra);
        }
        T2_TICK(4);
        softmax_chunk(rb, db + 16, ls);
        if (i == 0) lsum[0] = ls; else if (i == 1) lsum[1] = ls; else if (i == 2) lsum[2] = ls; else lsum[3] = ls;
        tmem_st32_nowait(slot + 32, rb);
        T2_TICK(5);
#else
        softmax_chunk(ra, db, ls);
#if T2_WATCH
        // S_ready of the next unit through the watcher's counter: a shared-memory load instead of a barrier test (the test stalls
        // the warp ~180 cycles even when the phase is complete — 7 % of all samples in profiles/attn_tc_lines_r02h.txt)
        if (u + 1 < total) s_ok = lds_acquire(wflag + 4 * g) >= (uint32_t)(u + 2);
#else
        if (u + 1 < total) s_ok = mbar_test_wait(nbarS, nparS);
#endif
        if (i == 3 && j + 1 < nblk) f_ok = mbar_test_wait(bar_full + 8 * ((j + 1) % T2_STAGES), (uint32_t)(((j + 1) / T2_STAGES) & 1));
        tmem_st32_nowait(slot, ra);
        T2_TICK(3);
        softmax_chunk(rb, db + 16, ls);
        if (i == 0) lsum[0] = ls; else if (i == 1) lsum[1] = ls; else if (i == 2) lsum[2] = ls; else lsum[3] = ls;
        tmem_st32_nowait(slot + 32, rb);
        T2_TICK(5);
        // S of the next unit is fetched at the unit end: with three slots its QK^T MMAs were issued two units ago
        tmem_wait_st();
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_P + 8 * (3 * g + k3));
        if (u + 1 < total) {
          if (!s_ok) mbar_wait_parked(nbarS, nparS);
          fence_after();
          tmem_ld32_nowait(nslot, ra);
          tmem_ld32_nowait(nslot + 32, rb);
        }
        T2_TICK(6);
#endif
      }
    }
#if T2_WATCH
    __syncwarp();
    if (lane == 0) atomicAdd(reinterpret_cast<unsigned int*>(bp + T2_OFF_BARS + 176), 1u);   // lets the watcher leave
#endif
#if T2_PIPE
    if (nblk > 0) {                                               // P of the last unit
      tmem_wait_st();
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_P + 8 * (3 * g + (total - 1) % 3));
    }
#endif
#if T2_PROF
    if (prof_on && lane == 0) {
      for (int q = 0; q < 8; ++q) g_t2_prof[q] = pc[q];
      g_t2_prof[8] = clock64() - tstart; g_t2_prof[9] = nblk;
    }
#endif
    // ---- epilogue: unnormalised O, reference max m, row sum l of this key split
    if (nblk > 0) { mbar_wait_parked(bar_O + 8 * g, 0); fence_after(); }
    const size_t prow = (size_t)split * p.Nq + (size_t)row;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      if (nblk > 0) { tmem_ld32_nowait(trow + t2_ocol(4 * g) + 32 * c, r); tmem_wait_ld(); }
      else {
#pragma unroll
        for (int t = 0; t < 32; ++t) r[t] = 0u;
      }
      if (rvalid) {
        float* op = p.Opart + prow * 128 + 64 * g + 32 * c;
#pragma unroll
        for (int t = 0; t < 8; ++t)
          *reinterpret_cast<float4*>(op + 4 * t) = make_float4(__uint_as_float(r[4 * t]), __uint_as_float(r[4 * t + 1]),
                                                               __uint_as_float(r[4 * t + 2]), __uint_as_float(r[4 * t + 3]));
      }
    }
    if (rvalid) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        p.Mpart[prow * 8 + 4 * g + i] = m[i];
        p.Lpart[prow * 8 + 4 * g + i] = lo32(lsum[i]) + hi32(lsum[i]);
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp < 10) {
    // =========================================================================== MMA issuers (one per head group)
    const int g = warp - 8;
    if (nblk > 0 && elect_one()) {
      const uint32_t id64 = idesc_f16(128, 64), id16 = idesc_f16(128, 16);
      const int total = 4 * nblk;
      const uint32_t qh_b = base + T2_OFF_Q + 16384 * g, ql_b = qh_b + 32768;
      const uint64_t ext_d = desc_sw128(base + T2_OFF_EXT + 32 * g);
      TC_ROLE_WAIT(bar_Q + 8 * g, 0);
      fence_after();
      auto issue_qk = [&](int u) {
        const int b = u >> 2, i = u & 3, st = b % T2_STAGES;
        if (i == 0) { TC_ROLE_WAIT(bar_full + 8 * st, (uint32_t)((b / T2_STAGES) & 1)); fence_after(); }
        const uint32_t d = tm + t2_slot(g, u % 3);
        const uint32_t kb = base + st * TC_IMG + 8192 * g + 32 * i;
        mma_ss(d, desc_sw128(qh_b + 32 * i), desc_sw128(kb), id64, 0);
        mma_ss(d, desc_sw128(ql_b + 32 * i), desc_sw128(kb), id64, 1);
        mma_ss(d, desc_sw128(qh_b + 32 * i), desc_sw128(kb + 16384), id64, 1);
        mma_ss(d, ext_d, desc_sw128(base + T2_OFF_SEL + 32 * i), id64, 1);          // - m_h (0 until the first reference is published)
        mma_commit(bar_S + 8 * (3 * g + u % 3));
      };
#if T2_PROF
      const bool prof_on = (blockIdx.x == 5 && blockIdx.y == 1 && g == 0);
      long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      long long tprev = clock64();
      const long long tstart = tprev;
#endif
      issue_qk(0);
      if (total > 1) issue_qk(1);
      if (total > 2) issue_qk(2);
      T2_TICK(0);
      for (int u = 0; u < total; ++u) {
        const int b = u >> 2, i = u & 3, st = b % T2_STAGES, h = 4 * g + i;
        TC_ROLE_WAIT(bar_P + 8 * (3 * g + u % 3), (uint32_t)((u / 3) & 1));
        fence_after();
        T2_TICK(1);
        const uint32_t vb = base + st * TC_IMG + TC_OFF_V + 2048 * h;
        const uint32_t ps = tm + t2_slot(g, u % 3);
        const uint32_t od = tm + t2_ocol(h);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          mma_ts(od, ps + 16 * ks, desc_sw128(vb + 32 * ks), id16, (b > 0 || ks > 0) ? 1u : 0u);
          mma_ts(od, ps + 16 * ks + 8, desc_sw128(vb + 32 * ks), id16, 1);
          mma_ts(od, ps + 16 * ks, desc_sw128(vb + 16384 + 32 * ks), id16, 1);
        }
        if (i == 3) mma_commit(bar_empty + 8 * st);
        T2_TICK(2);
        if (u + 3 < total) issue_qk(u + 3);                          // into the slot whose P this unit's PV MMAs just consumed
        T2_TICK(3);
      }
      mma_commit(bar_O + 8 * g);
#if T2_PROF
      if (prof_on) {
        for (int q = 0; q < 8; ++q) g_t2_prof[16 + q] = pc[q];
        g_t2_prof[24] = clock64() - tstart;
      }
#endif
    }
    __syncwarp();
    } else if (warp == 10) {
    // =========================================================================== loader (cp.async.bulk, 2-stage ring)
    if (elect_one()) {
      for (int j = 0; j < nblk; ++j) {
        const int st = j % T2_STAGES;
        TC_ROLE_WAIT(bar_empty + 8 * st, (uint32_t)(((j / T2_STAGES) & 1) ^ 1));
        mbar_arrive_expect_tx(bar_full + 8 * st, TC_IMG_USED);
        const uint8_t* src = img0 + (size_t)j * TC_IMG;
        const uint32_t dst = base + st * TC_IMG;
#pragma unroll
        for (int c = 0; c < 4; ++c) bulk_g2s(dst + 16384 * c, src + 16384 * c, 16384, bar_full + 8 * st);
        bulk_g2s(dst + 65536, src + 65536, TC_IMG_USED - 65536, bar_full + 8 * st);
      }
    }
    __syncwarp();
    }
#if T2_WATCH
    else {
    // =========================================================================== watcher (the fourth role warp)
    // Polls the S_ready barriers of both head groups in unit order and publishes how many units are ready, so that the softmax
    // warps learn it from a shared-memory load. Observing a barrier with acquire semantics and releasing the counter orders the
    // MMA's TMEM writes before any softmax thread that acquires the counter. If the watcher ever lags (or mis-reads a barrier
    // that has run two phases ahead) the counter merely stays behind and the softmax warps fall back to the barrier itself.
    if (nblk > 0 && elect_one()) {
      const int total = 4 * nblk;
      int c0 = 0, c1 = 0;
      while (c0 < total || c1 < total) {
        if (c0 < total && mbar_test_wait(bar_S + 8 * (c0 % 3), (uint32_t)((c0 / 3) & 1))) { ++c0; sts_release(wflag, (uint32_t)c0); }
        if (c1 < total && mbar_test_wait(bar_S + 8 * (3 + c1 % 3), (uint32_t)((c1 / 3) & 1))) { ++c1; sts_release(wflag + 4, (uint32_t)c1); }
        if (lds_acquire(wflag + 8) >= 8u) break;                   // every softmax warp has left its loop
      }
    }
    __syncwarp();
    }
#endif
  }
  fence_before();
  __syncthreads();
  if (warp == 10) tmem_dealloc(tm, 512);
}

// ---------------------------------------------------------------------------------------------- 16-softmax-warp variant
// MEASURED SLOWER than the 8-warp kernel (37.5 vs 35.9 ms at 100k nodes, profiles/attn_tc16_r02b.txt, profiles/microbench_r02.txt) and therefore NOT the default
// (DGDM_ATTN_TC_WARPS=16 selects it): the kernel is not latency-bound. tools/mufu_probe.cu shows that the softmax instruction mix
// itself (per two scores: 2 MUFU.EX2, 3 FADD2, 2 LOP3, 1 FFMA2, 2 F2FP) sustains only 24.5 / 22.0 / 21.4 cycles per pair at 2 / 4 / 8
// warps per scheduler against the 16-cycle MUFU ideal — the packed-fp32 and conversion instructions co-limit with the XU pipe — so
// doubling the warps buys at most 10 % while halving the keys per thread doubles the per-unit synchronisation overhead.
// Same arithmetic, data layout and MMA schedule as spatial_attn_tc_kernel, but 4 softmax warps per scheduler instead of 2: the
// round-1 capture showed the two 216-register warps of a scheduler latency-bound (issue 53 %, XU 64 %, "wait" the top stall), i.e.
// neither the issue slots nor the MUFU pipe could be kept busy by two in-order instruction streams. Here a softmax thread owns
// (query row, 4 heads, 32 of the block's 64 keys): 16 softmax warps = {head group g} x {key half kh} x {row quarter}, 112 registers
// each, processing 16-key quarters (one K-step of the PV MMA) so that the TMEM read of the next quarter overlaps the exp2 phase.
// The two threads that share a (row, head) keep IDENTICAL lazy reference maxima: the decision to refresh depends only on values both
// hold, and the refreshed maximum is exchanged through shared memory under a 64-thread named barrier (rare path).
constexpr int T16_THREADS = 640;                // 16 softmax warps + {MMA 0, MMA 1, loader, idle}
constexpr int T16_XCH = 8192;                   // need-exchange [2 buf][2 g][2 kh][128] floats + epilogue l exchange [2 g][128][4]
constexpr int T16_SMEM = TC_SMEM + T16_XCH;
#ifndef T16_MMA_SLEEP_NS
#define T16_MMA_SLEEP_NS 32
#endif

__global__ void __launch_bounds__(T16_THREADS, 1) spatial_attn_tc16_kernel(const AttnTcParams p) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + TC_STAGES * TC_IMG;
  const uint32_t bar_full = bars, bar_empty = bars + 24, bar_S = bars + 48, bar_P = bars + 80, bar_O = bars + 112, bar_Q = bars + 128;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bp + TC_STAGES * TC_IMG + 192);
  float* xch = reinterpret_cast<float*>(bp + TC_STAGES * TC_IMG + 1024 + 256);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int seg = -1, row0 = 0;
  {
    int acc = 0;
    for (int s = 0; s < p.num_seg; ++s) {
      const int a = p.qseg[s], b = p.qseg[s + 1];
      const int nt = (b - a + TC_QT - 1) / TC_QT;
      if ((int)blockIdx.x < acc + nt) { seg = s; row0 = a + ((int)blockIdx.x - acc) * TC_QT; break; }
      acc += nt;
    }
  }
  if (seg < 0) return;
  const int qend = p.qseg[seg + 1];
  const int split = blockIdx.y;
  const int nb_seg = p.blk_ptr[seg + 1] - p.blk_ptr[seg];
  const int per = (nb_seg + p.nsplit - 1) / p.nsplit;
  const int b_lo = min(nb_seg, split * per), b_hi = min(nb_seg, b_lo + per);
  const int nblk = b_hi - b_lo;
  const uint8_t* img0 = p.img + (size_t)(p.blk_ptr[seg] + b_lo) * TC_IMG;

  if (tid == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 2 + 16); }
    for (int i = 0; i < 4; ++i) { mbar_init(bar_S + 8 * i, 1); mbar_init(bar_P + 8 * i, 8); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_O + 8 * i, 1); mbar_init(bar_Q + 8 * i, 4); }
    mbar_fence_init();
  }
  if (warp == 18) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tm = *tmem_slot;

  if (warp < 16) {
    // =========================================================================== softmax warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int wq = warp & 3, g = warp >> 3, kh = (warp >> 2) & 1;
    const int rl = wq * 32 + lane;                                  // row inside the tile
    const int row = row0 + rl;
    const bool rvalid = row < qend;
    const uint32_t trow = tm + ((uint32_t)(wq * 32) << 16);
    const int bar_id = 1 + g * 4 + wq;                              // named barrier of the two key halves of this (g, row quarter)
    float qn[4];
    {
      const float* qp = p.q_qkv + (size_t)(rvalid ? row : row0) * p.ld_q + 64 * g;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t qh[8], ql[8];
        float sq = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float4 v = ldg4(qp + 16 * i + 4 * c);
          if (!rvalid) v = make_float4(0.f, 0.f, 0.f, 0.f);
          v.x *= p.qscale; v.y *= p.qscale; v.z *= p.qscale; v.w *= p.qscale;
          sq += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          split_pack(v.x, v.y, qh[2 * c], ql[2 * c]);
          split_pack(v.z, v.w, qh[2 * c + 1], ql[2 * c + 1]);
        }
        qn[i] = sqrtf(sq) * 1.0001f;
        if (kh == 0) {
          tmem_st8_nowait(trow + tc_qhi(4 * g + i), qh);
          tmem_st8_nowait(trow + tc_qlo(4 * g + i), ql);
        }
      }
      if (kh == 0) {
        tmem_wait_st();
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_Q + 8 * g);
      }
    }
    float qx = 0.f, qy = 0.f;
    if (p.q_pos != nullptr && rvalid) { qx = p.q_pos[2 * (size_t)row] * p.pscale; qy = p.q_pos[2 * (size_t)row + 1] * p.pscale; }

    float m[4];
    unsigned long long lsum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; lsum[i] = 0ull; }
    int nx = 0;                                                      // need-exchange counter (buffer parity)

    uint32_t ra[16], rb[16];                                         // keys [32 kh, 32 kh + 16) and [32 kh + 16, 32 kh + 32) of the unit
    const int total = 4 * nblk;
    const uint32_t kcol = (uint32_t)(32 * kh);
    if (nblk > 0) {
      mbar_wait_parked(bar_S + 8 * (2 * g), 0);
      fence_after();
      tmem_ld16_nowait(trow + tc_slot(g, 0) + kcol, ra);
      tmem_ld16_nowait(trow + tc_slot(g, 0) + kcol + 16, rb);
    }
    const unsigned long long neg1 = pk2(-1.0f, -1.0f);
    // exp2 + fp16 hi/lo split of one 16-key quarter held in r (in place: r[0..7] = hi, r[8..15] = lo), accumulating the row sum
    auto softmax_q = [&](uint32_t* r, const unsigned long long* nbc, unsigned long long nmu2, unsigned long long& ls) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const unsigned long long x = add2(add2(pk2(__uint_as_float(r[2 * t]), __uint_as_float(r[2 * t + 1])), nbc[t]), nmu2);
        const float p0 = tc_ex2(lo32(x)), p1 = tc_ex2(hi32(x));
        const unsigned long long pp = pk2(p0, p1);
        ls = add2(ls, pp);
        const float h0 = __uint_as_float(__float_as_uint(p0) & 0xffffe000u), h1 = __uint_as_float(__float_as_uint(p1) & 0xffffe000u);
        const unsigned long long ll = fma2(pk2(h0, h1), neg1, pp);
        hi[t] = tc_pack_h2(h0, h1);
        lo[t] = tc_pack_h2(lo32(ll), hi32(ll));
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) { r[t] = hi[t]; r[8 + t] = lo[t]; }
    };

    for (int j = 0; j < nblk; ++j) {
      const int st = j % TC_STAGES;
      mbar_wait_parked(bar_full + 8 * st, (uint32_t)((j / TC_STAGES) & 1));
      unsigned long long nb[16];                                      // negated distance bias of this row against my 32 keys
      float kmax[4];
      {
        const float* kp = reinterpret_cast<const float*>(bp + st * TC_IMG + TC_OFF_POS) + 32 * kh;
        const unsigned long long qx2 = pk2(qx, qx), qy2 = pk2(qy, qy);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float4 kx = *reinterpret_cast<const float4*>(kp + 4 * t);
          const float4 ky = *reinterpret_cast<const float4*>(kp + TC_BN + 4 * t);
          unsigned long long dx = sub2(qx2, pk2(kx.x, kx.y)), dy = sub2(qy2, pk2(ky.x, ky.y));
          unsigned long long d2 = fma2(dx, dx, mul2(dy, dy));
          nb[2 * t] = pk2(-tc_sqrt(lo32(d2)), -tc_sqrt(hi32(d2)));
          dx = sub2(qx2, pk2(kx.z, kx.w)); dy = sub2(qy2, pk2(ky.z, ky.w));
          d2 = fma2(dx, dx, mul2(dy, dy));
          nb[2 * t + 1] = pk2(-tc_sqrt(lo32(d2)), -tc_sqrt(hi32(d2)));
        }
        const float4 km = *reinterpret_cast<const float4*>(bp + st * TC_IMG + TC_OFF_KMAX + 16 * g);
        kmax[0] = km.x; kmax[1] = km.y; kmax[2] = km.z; kmax[3] = km.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + 8 * st);

#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int u = 4 * j + i, s = u & 1;
        const uint32_t slot = trow + tc_slot(g, s) + kcol;
        const uint32_t nslot = trow + tc_slot(g, s ^ 1) + kcol;
        tmem_wait_ld();
        const bool need = (j == 0) || (qn[i] * kmax[i] > m[i] + TC_TAU);
        if (__any_sync(0xffffffffu, need)) {
          float rmax = -INFINITY;
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const unsigned long long xa = add2(pk2(__uint_as_float(ra[2 * t]), __uint_as_float(ra[2 * t + 1])), nb[t]);
            const unsigned long long xb = add2(pk2(__uint_as_float(rb[2 * t]), __uint_as_float(rb[2 * t + 1])), nb[8 + t]);
            rmax = fmaxf(rmax, fmaxf(fmaxf(lo32(xa), hi32(xa)), fmaxf(lo32(xb), hi32(xb))));
          }
          // the other key half of this (row, head) takes the same branch (same qn, kmax, m): exchange the partial maxima
          float* xb_ = xch + (((nx & 1) * 2 + g) * 2) * 128;
          xb_[kh * 128 + rl] = rmax;
          named_bar_sync(bar_id, 64);
          rmax = fmaxf(rmax, xb_[(kh ^ 1) * 128 + rl]);
          ++nx;
          float sc = 1.0f;
          if (j == 0) m[i] = rmax;
          else if (rmax > m[i] + TC_TAU) { sc = tc_ex2(m[i] - rmax); m[i] = rmax; }
          if (__any_sync(0xffffffffu, sc != 1.0f)) {
            if (kh == 0) {                                            // O_h of this row is rescaled once, by the kh = 0 thread
              uint32_t o[16];
              tmem_ld16_nowait(trow + tc_ocol(4 * g + i), o);
              tmem_wait_ld();
#pragma unroll
              for (int t = 0; t < 16; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * sc);
              tmem_st16_nowait(trow + tc_ocol(4 * g + i), o);
            }
            lsum[i] = mul2(lsum[i], pk2(sc, sc));
          }
        }
        const float mu = (m[i] == -INFINITY) ? 0.f : m[i];
        const unsigned long long nmu2 = pk2(-mu, -mu);
        softmax_q(ra, nb, nmu2, lsum[i]);
        tmem_st16_nowait(slot, ra);
        if (u + 1 < total) {                                          // prefetch the first quarter of the next unit's S into ra
          mbar_wait_parked(bar_S + 8 * (2 * g + (s ^ 1)), (uint32_t)(((u + 1) >> 1) & 1));
          fence_after();
          tmem_ld16_nowait(nslot, ra);
        }
        softmax_q(rb, nb + 8, nmu2, lsum[i]);
        tmem_st16_nowait(slot + 16, rb);
        tmem_wait_st();
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_P + 8 * (2 * g + s));
        if (u + 1 < total) tmem_ld16_nowait(nslot + 16, rb);
      }
    }
    // ---- epilogue: unnormalised O (each key-half warp stores 32 of the group's 64 columns), reference max m, row sum l
    if (nblk > 0) { mbar_wait_parked(bar_O + 8 * g, 0); fence_after(); }
    float* lx = xch + 1024 + (g * 128 + rl) * 4;
    if (kh == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) lx[i] = lo32(lsum[i]) + hi32(lsum[i]);
    }
    named_bar_sync(bar_id, 64);
    const size_t prow = (size_t)split * p.Nq + (size_t)row;
    {
      uint32_t r[32];
      if (nblk > 0) { tmem_ld32_nowait(trow + tc_ocol(4 * g) + 32 * kh, r); tmem_wait_ld(); }
      else {
#pragma unroll
        for (int t = 0; t < 32; ++t) r[t] = 0u;
      }
      if (rvalid) {
        float* op = p.Opart + prow * 128 + 64 * g + 32 * kh;
#pragma unroll
        for (int t = 0; t < 8; ++t)
          *reinterpret_cast<float4*>(op + 4 * t) = make_float4(__uint_as_float(r[4 * t]), __uint_as_float(r[4 * t + 1]),
                                                               __uint_as_float(r[4 * t + 2]), __uint_as_float(r[4 * t + 3]));
      }
    }
    if (rvalid && kh == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        p.Mpart[prow * 8 + 4 * g + i] = m[i];
        p.Lpart[prow * 8 + 4 * g + i] = (lo32(lsum[i]) + hi32(lsum[i])) + lx[i];
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp < 18) {
    // =========================================================================== MMA issuers (one per head group)
    const int g = warp - 16;
    if (nblk > 0 && elect_one()) {
      const uint32_t id64 = idesc_f16(128, 64), id16 = idesc_f16(128, 16);
      const int total = 4 * nblk;
      TC_ROLE_WAIT(bar_Q + 8 * g, 0);
      fence_after();
      auto issue_qk = [&](int u) {
        const int b = u >> 2, i = u & 3, st = b % TC_STAGES;
        if (i == 0) { TC_ROLE_WAIT(bar_full + 8 * st, (uint32_t)((b / TC_STAGES) & 1)); fence_after(); }
        const uint32_t d = tm + tc_slot(g, u & 1);
        const uint32_t kb = base + st * TC_IMG + 8192 * g + 32 * i;
        mma_ts(d, tm + tc_qhi(4 * g + i), desc_sw128(kb), id64, 0);
        mma_ts(d, tm + tc_qlo(4 * g + i), desc_sw128(kb), id64, 1);
        mma_ts(d, tm + tc_qhi(4 * g + i), desc_sw128(kb + 16384), id64, 1);
        mma_commit(bar_S + 8 * (2 * g + (u & 1)));
      };
      issue_qk(0);
      if (total > 1) issue_qk(1);
      for (int u = 0; u < total; ++u) {
        const int b = u >> 2, i = u & 3, st = b % TC_STAGES, h = 4 * g + i;
        TC_ROLE_WAIT(bar_P + 8 * (2 * g + (u & 1)), (uint32_t)((u >> 1) & 1));
        fence_after();
        const uint32_t vb = base + st * TC_IMG + TC_OFF_V + 2048 * h;
        const uint32_t ps = tm + tc_slot(g, u & 1);
        const uint32_t od = tm + tc_ocol(h);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          mma_ts(od, ps + 16 * ks, desc_sw128(vb + 32 * ks), id16, (b > 0 || ks > 0) ? 1u : 0u);
          mma_ts(od, ps + 16 * ks + 8, desc_sw128(vb + 32 * ks), id16, 1);
          mma_ts(od, ps + 16 * ks, desc_sw128(vb + 16384 + 32 * ks), id16, 1);
        }
        if (i == 3) mma_commit(bar_empty + 8 * st);
        if (u + 2 < total) issue_qk(u + 2);
      }
      mma_commit(bar_O + 8 * g);
    }
    __syncwarp();
    } else if (warp == 18) {
    // =========================================================================== loader (cp.async.bulk, 3-stage ring)
    if (elect_one()) {
      for (int j = 0; j < nblk; ++j) {
        const int st = j % TC_STAGES;
        TC_ROLE_WAIT(bar_empty + 8 * st, (uint32_t)(((j / TC_STAGES) & 1) ^ 1));
        mbar_arrive_expect_tx(bar_full + 8 * st, TC_IMG_USED);
        const uint8_t* src = img0 + (size_t)j * TC_IMG;
        const uint32_t dst = base + st * TC_IMG;
#pragma unroll
        for (int c = 0; c < 4; ++c) bulk_g2s(dst + 16384 * c, src + 16384 * c, 16384, bar_full + 8 * st);
        bulk_g2s(dst + 65536, src + 65536, TC_IMG_USED - 65536, bar_full + 8 * st);
      }
    }
    __syncwarp();
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 18) tmem_dealloc(tm, 512);
}

static inline size_t tc_al256(size_t x) { return (x + 255) & ~(size_t)255; }

// Key splits per query tile. Two constraints: (1) occupancy — at least ~8 CTAs of work per SM; (2) ACCURACY — the tensor core
// accumulates O in fp32 with truncation, 12 accumulate steps per key block (4 K-steps x 3 split terms), which biases a long
// chain low by ~3.7e-8 per step (measured: -3.5e-4 at 781 blocks per CTA, tests/test_gpu_properties.py). The chain is capped
// at TC_MAX_CHAIN key blocks per CTA (bias <= ~2e-5; 32 costs +1.1 ms at 100k nodes, 48 costs +0.5 ms); the partial (O, m, l) of the splits are merged in fp32 round-to-nearest
// by attn_combine_kernel.
constexpr int TC_MAX_CHAIN = 48;
int attn_tc_nsplit(int Nq, int Nkv, int num_seg, int max_seg_keys) {
  const int tiles = (Nq + TC_QT - 1) / TC_QT + num_seg;
  int s = (148 * 8 + tiles - 1) / tiles;                       // aim for >= ~8 CTAs of work per SM
  const int avg_blocks = (Nkv / (num_seg > 0 ? num_seg : 1)) / TC_BN;
  const int max_s = avg_blocks / 8 > 1 ? avg_blocks / 8 : 1;   // at least ~8 key blocks per CTA
  if (s > max_s) s = max_s;
  if (s > 16) s = 16;
  const int longest = (max_seg_keys > 0 && max_seg_keys < Nkv) ? max_seg_keys : Nkv;   // keys of the largest segment (Nkv when unknown)
  const int max_blocks = (longest + TC_BN - 1) / TC_BN;
  const int s_chain = (max_blocks + TC_MAX_CHAIN - 1) / TC_MAX_CHAIN;
  if (s < s_chain) s = s_chain;
  if (s < 1) s = 1;
  return s;
}

size_t attn_tc_work_bytes(int Nq, int Nkv, int D, int heads, int num_seg, int max_seg_keys) {
  const int ns = attn_tc_nsplit(Nq, Nkv, num_seg, max_seg_keys);
  const size_t nblk_max = (size_t)(Nkv + TC_BN - 1) / TC_BN + (size_t)num_seg;
  size_t b = 1024;
  b += tc_al256(nblk_max * TC_IMG) + 1024;
  b += tc_al256((size_t)(num_seg + 1) * 4);
  b += tc_al256((size_t)ns * Nq * D * 4);
  b += 2 * tc_al256((size_t)ns * Nq * heads * 4);
  return b;
}

// defined in attention.cu
void attn_launch_combine(const float* Opart, const float* Mpart, const float* Lpart, int nsplit, int Nq, int D, int heads,
                         float* out, int ldo, float* lse_out, cudaStream_t s);

int attn_tc_launch(const float* q_qkv, int ld_q, const float* q_pos, const int* qseg_ptr, int Nq,
                   const float* kv_qkv, int ld_kv, const float* kv_pos, const int* kseg_ptr, int Nkv,
                   int num_seg, int max_seg_keys, float temperature, float* out, int ldo, float* lse_out, void* work, size_t work_bytes,
                   cudaStream_t s) {
  constexpr int D = 128, heads = 8;
  if (work_bytes < attn_tc_work_bytes(Nq, Nkv, D, heads, num_seg, max_seg_keys)) return DGDM_ERR_INVALID;
  const int ns = attn_tc_nsplit(Nq, Nkv, num_seg, max_seg_keys);
  const size_t nblk_max = (size_t)(Nkv + TC_BN - 1) / TC_BN + (size_t)num_seg;
  uint8_t* w = (uint8_t*)(((uintptr_t)work + 1023) & ~(uintptr_t)1023);
  uint8_t* img = w; w += tc_al256(nblk_max * TC_IMG);
  int* blk_ptr = (int*)w; w += tc_al256((size_t)(num_seg + 1) * 4);
  float* Opart = (float*)w; w += tc_al256((size_t)ns * Nq * D * 4);
  float* Mpart = (float*)w; w += tc_al256((size_t)ns * Nq * heads * 4);
  float* Lpart = (float*)w;
  const float LOG2E = 1.4426950408889634f;
  const float pscale = LOG2E / temperature;
  attn_tc_blkptr_kernel<<<1, 32, 0, s>>>(kseg_ptr, num_seg, blk_ptr);
  attn_tc_prep_kernel<<<(unsigned)nblk_max, 256, 0, s>>>(kv_qkv, ld_kv, kv_pos, kseg_ptr, blk_ptr, num_seg, pscale, img);
  static PerDeviceOnce attr_set;
  if (attr_set.needed()) {
    if (cudaFuncSetAttribute(spatial_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(spatial_attn_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(spatial_attn_tc16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T16_SMEM) != cudaSuccess) return DGDM_ERR_LAUNCH;
    attr_set.mark();
  }
  // 2 = spatial_attn_tc2_kernel (default), 8 = the round-1/2 kernel, 16 = its 16-softmax-warp variant
  static const int variant = [] { const char* e = getenv("DGDM_ATTN_TC_WARPS"); return e ? atoi(e) : 2; }();
  static const int sleep_ns = [] { const char* e = getenv("DGDM_ATTN_TC_SLEEP"); return e ? atoi(e) : 0; }();
  AttnTcParams p;
  p.q_qkv = q_qkv; p.ld_q = ld_q; p.q_pos = q_pos; p.qseg = qseg_ptr; p.kseg = kseg_ptr; p.num_seg = num_seg; p.Nq = Nq; p.nsplit = ns;
  p.img = img; p.blk_ptr = blk_ptr; p.qscale = LOG2E / 4.0f; p.pscale = pscale;
  p.Opart = Opart; p.Mpart = Mpart; p.Lpart = Lpart;
  p.mma_sleep_ns = sleep_ns;
  const int tiles = (Nq + TC_QT - 1) / TC_QT + num_seg;
  if (variant == 2) spatial_attn_tc2_kernel<<<dim3(tiles, ns, 1), TC_THREADS, T2_SMEM, s>>>(p);
  else if (variant == 8) spatial_attn_tc_kernel<<<dim3(tiles, ns, 1), TC_THREADS, TC_SMEM, s>>>(p);
  else spatial_attn_tc16_kernel<<<dim3(tiles, ns, 1), T16_THREADS, T16_SMEM, s>>>(p);
  attn_launch_combine(Opart, Mpart, Lpart, ns, Nq, D, heads, out, ldo, lse_out, s);
  if (cudaGetLastError() != cudaSuccess) return DGDM_ERR_LAUNCH;
  return DGDM_OK;
}

}  // namespace dgdm

#if T2_PROF
extern "C" int dgdm_debug_attn_prof(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, dgdm::g_t2_prof, sizeof(long long) * 32) == cudaSuccess ? 0 : 1;
}
#endif
