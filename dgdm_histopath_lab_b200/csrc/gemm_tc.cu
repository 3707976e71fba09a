// tcgen05 / TMEM GEMM with fused row-wise epilogue for the DGDM node-level linear layers.
//
//   C[M,N] = epilogue( A[M,K] (fp32) @ W[N,K]^T )     N <= 512, whole output rows live in TMEM (512 columns)
//
// Precision: A and W are each split into two fp16 planes (hi, lo); D += Alo*Whi + Ahi*Wlo + Ahi*Whi with fp32
// accumulation in TMEM (3 tcgen05.mma kind::f16 per K-step) -> ~2^-21 relative, i.e. fp32-class results.
//
// Persistent, warp-specialised kernel: one CTA per SM walks the 128-row tiles m = blockIdx.x, +gridDim.x, ...
//   warps 0-3   epilogue: TMEM lane == output row == thread, so LayerNorm statistics are thread-local; global I/O goes
//               through a per-warp 32x32 smem transpose so that every global access is a full 128-byte row segment.
//               With N <= 256 there are two 256-column accumulator buffers and the epilogue of tile t overlaps the
//               main loop of tile t+1; with N > 256 the single 512-column buffer serialises MMA and epilogue, but the
//               loads / conversions of the next tile still run ahead.
//   warps 4-11  A converters: fp32 global rows (128-bit coalesced, next K-block prefetched into registers before the
//               current one is converted -> ~64 KB of loads in flight per SM) -> hi/lo fp16 -> smem ring in the K-major
//               SWIZZLE_128B canonical layout.
//   warp 12     W producer: one elected lane streams the pre-swizzled weight K-block images (dgdm_split_weight_f16) with
//               cp.async.bulk (TMA engine) into the W ring; completion is counted in bytes on an mbarrier.
//   warp 13     MMA issuer: one elected lane; tcgen05.commit frees ring slots and publishes finished accumulators.
// N > 256 is covered by column chunks of <= 256 into adjacent TMEM column ranges (one W ring unit per chunk).
#include "tc05.cuh"

namespace dgdm {
using namespace tc;

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_NCHUNK = 256;
constexpr int GEMM_EPI_WARPS = 4;
constexpr int GEMM_CONV_WARPS = 8;
constexpr int GEMM_CONV_THREADS = GEMM_CONV_WARPS * 32;
constexpr int GEMM_THREADS = (GEMM_EPI_WARPS + GEMM_CONV_WARPS + 2) * 32;     // 448
constexpr int A_PLANE_BYTES = GEMM_BM * GEMM_BK * 2;                           // 16 KB
constexpr int A_SLOT_BYTES = 2 * A_PLANE_BYTES;                                // 32 KB (hi | lo)
constexpr int GEMM_RING_BYTES = 192 * 1024;                                    // A slots + W slots
constexpr int TR_STRIDE = 36;                                                  // floats; conflict-free for 128-bit accesses
constexpr int GEMM_STAGING_BYTES = GEMM_EPI_WARPS * 32 * TR_STRIDE * 4;        // 18432
constexpr int GEMM_MAX_SLOTS = 4;
constexpr int GEMM_VEC_BYTES = 6 * 512 * 4;                                    // bias, ln1 g/b, ln2 g/b, dot_w
constexpr int GEMM_SMEM = GEMM_RING_BYTES + GEMM_STAGING_BYTES + GEMM_VEC_BYTES + 256 /*barriers*/ + 1024 /*align*/;

#ifdef DGDM_GEMM_TRACE
// dev-only instrumentation (tools/gemm_trace.sh): cycles each role spends in its waits, per CTA
__device__ long long g_gemm_events[12][64];      // CTA 0 event times (cycles since the thread entered the kernel)
#define TRACE_EVENT(role, idx) do { if (blockIdx.x == 0 && (idx) < 64) g_gemm_events[role][idx] = clock64() - t_kernel0; } while (0)
#else
#define TRACE_EVENT(role, idx)
#endif

struct GemmSched {
  int na, nw;              // ring depths
  int w_slot_bytes;        // bytes per W ring slot (2 planes x chunk rows x 128 B)
  int nbuf;                // accumulator buffers in TMEM (2 when N <= 256, else 1)
  int ntiles;
  long long w_plane_stride;   // bytes between the hi and lo plane of one K-block image ( = total weight rows x 128 )
};

__device__ __forceinline__ uint32_t pack_half2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
// 4 fp32 -> packed fp16 hi (round to nearest) and lo = fp16(x - hi): cvt.rn.f16x2.f32 pairs, 3 instructions per element
__device__ __forceinline__ void split4(float4 v, uint2& hi, uint2& lo) {
  const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
  hi = make_uint2(h2_bits(h01), h2_bits(h23));
  lo = make_uint2(h2_bits(l01), h2_bits(l23));
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  tmem_ld32_nowait(taddr, reinterpret_cast<uint32_t*>(v));
  tmem_wait_ld();
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  tmem_st32_nowait(taddr, reinterpret_cast<const uint32_t*>(v));
  tmem_wait_st();
}

// ---- per-warp 32x32 transposing I/O through smem ----
__device__ __forceinline__ void multimem_st4(float* mc, float4 t) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
}
// thread-row-owned regs v[32] (row = lane) -> coalesced global store of rows [row0, row0+32), cols [c0, c0+32).
// Node-partitioned graphs: the same 128-byte row segments also go to the peers' copies of C (P2P stores over NVLink, or one
// NVLS multicast store that the switch replicates to every rank) — the halo all-gather fused into the GEMM epilogue.
__device__ __forceinline__ void warp_store_tile(float* st, const float* v, float* G, int ldg, int row0, int c0,
                                                int M, int N, int lane) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(st + lane * TR_STRIDE + j * 4) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    int r = it * 4 + (lane >> 3), c = (lane & 7) * 4;
    float4 t = *reinterpret_cast<const float4*>(st + r * TR_STRIDE + c);
    if (row0 + r < M && c0 + c < N) *reinterpret_cast<float4*>(G + (size_t)(row0 + r) * ldg + c0 + c) = t;
  }
  __syncwarp();
}
// Same tile also to the peers' copies of C (the fused halo all-gather): kept out of line of the single-GPU store above so that
// the common path stays as lean as before (an in-loop peer branch cost 30-45 % on the plain GEMMs).
__device__ __forceinline__ void warp_store_tile_peers(const float* st, const dgdm_gemm_params& p, int row0, int c0, int lane) {
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    int r = it * 4 + (lane >> 3), c = (lane & 7) * 4;
    if (row0 + r < p.M && c0 + c < p.N) {
      const float4 t = *reinterpret_cast<const float4*>(st + r * TR_STRIDE + c);
      const size_t off = (size_t)(row0 + r) * p.ldc + c0 + c;
      if (p.peers_multicast) multimem_st4(p.C_peers[0] + off, t);
      else {
        *reinterpret_cast<float4*>(p.C + off) = t;
        for (int d = 0; d < p.n_peers; ++d) *reinterpret_cast<float4*>(p.C_peers[d] + off) = t;
      }
    }
  }
  __syncwarp();
}
template <bool PEERS>
__device__ __forceinline__ void warp_store_out(float* st, const float* v, const dgdm_gemm_params& p, int row0, int c0, int lane) {
  if (!PEERS) { warp_store_tile(st, v, p.C, p.ldc, row0, c0, p.M, p.N, lane); return; }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(st + lane * TR_STRIDE + j * 4) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  warp_store_tile_peers(st, p, row0, c0, lane);
}
// coalesced global load of a 32x32 tile -> thread-row-owned regs v[32]
__device__ __forceinline__ void warp_load_tile(float* st, float* v, const float* G, int ldg, int row0, int c0,
                                               int M, int N, int lane) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    int r = it * 4 + (lane >> 3), c = (lane & 7) * 4;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < M && c0 + c < N) t = ldg4(G + (size_t)(row0 + r) * ldg + c0 + c);
    *reinterpret_cast<float4*>(st + r * TR_STRIDE + c) = t;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 t = *reinterpret_cast<const float4*>(st + lane * TR_STRIDE + j * 4);
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
  __syncwarp();
}

// Split form of warp_load_tile for software pipelining: the global loads of the NEXT 32-column chunk of a residual are issued into
// registers right after the current chunk was consumed, so their DRAM latency overlaps the epilogue arithmetic instead of sitting in
// front of the shared-memory transpose (ncu r02e: 98 % long-scoreboard stalls on that store, ~1/4 of the epilogue warps' time).
__device__ __forceinline__ void warp_load_issue(float4 r[8], const float* G, int ldg, int row0, int c0, int M, int N, int lane) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + (lane >> 3), c = (lane & 7) * 4;
    r[it] = (row0 + rr < M && c0 + c < N) ? ldg4(G + (size_t)(row0 + rr) * ldg + c0 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void warp_load_commit(float* st, float* v, const float4 r[8], int lane) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + (lane >> 3), c = (lane & 7) * 4;
    *reinterpret_cast<float4*>(st + rr * TR_STRIDE + c) = r[it];
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = *reinterpret_cast<const float4*>(st + lane * TR_STRIDE + j * 4);
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
  __syncwarp();
}

// Column vectors of the epilogue staged once per CTA in shared memory (zero / one filled when absent) so that the element
// loops below are branch-free FFMA chains fed by broadcast LDS.128 instead of per-element predicated global loads.
struct EpiVec { const float* bias; const float* g1; const float* b1; const float* g2; const float* b2; const float* dotw; };

__device__ __forceinline__ void lds_row32(const float* s, float* o) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 q = *reinterpret_cast<const float4*>(s + 4 * j);
    o[4 * j] = q.x; o[4 * j + 1] = q.y; o[4 * j + 2] = q.z; o[4 * j + 3] = q.w;
  }
}

// Row-wise epilogue of one 128-row tile by one epilogue warp (32 rows): thread = row, accumulator row in TMEM at t_row.
// Every feature test is warp-uniform and hoisted out of the 32-element loops.
template <bool PEERS>
__device__ __forceinline__ void epilogue_rows(const dgdm_gemm_params& p, const EpiVec& ev, float* st, uint32_t t_row, int row0, int lane) {
  const int row = row0 + lane;
  const float inv_scale = 1.0f / p.a_scale;
  const bool has_ln1 = p.ln1_g != nullptr, has_ln2 = p.ln2_g != nullptr;
  const bool writeback = has_ln1 || has_ln2;
  const float invN = 1.0f / (float)p.N;
  float v[32], t[32];
  float4 pre[8];                // prefetched residual chunk (r_pre in pass 1, r_post in the finishing pass)
  float sum = 0.f, dot = 0.f;

  auto finish = [&](int c0) {   // v holds the post-LN values of columns [c0, c0+32)
    apply_act_chunk<32>(v, p.act_post);
    if (p.r_post != nullptr) {
      warp_load_commit(st, t, pre, lane);
      if (c0 + 32 < p.N) warp_load_issue(pre, p.r_post, p.ld_r_post, row0, c0 + 32, p.M, p.N, lane);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += t[j];
    }
    if (p.dot_w != nullptr) {        // columns >= N hold zeros in ev.dotw
      lds_row32(ev.dotw + c0, t);
#pragma unroll
      for (int j = 0; j < 32; ++j) dot = fmaf(v[j], t[j], dot);
    }
    if (p.C != nullptr) warp_store_out<PEERS>(st, v, p, row0, c0, lane);
  };
  auto normalise = [&](int c0, const float* g, const float* b, float mean, float rstd) {
    lds_row32(g + c0, t);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (v[j] - mean) * rstd * t[j];
    lds_row32(b + c0, t);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += t[j];
    if (c0 + 32 > p.N) {
#pragma unroll
      for (int j = 16; j < 32; ++j) v[j] = 0.f;
    }
  };
  auto sq_dev = [&](int c0, float mean, float& ss) {
    if (c0 + 32 > p.N) {
#pragma unroll
      for (int j = 0; j < 16; ++j) { const float d = v[j] - mean; ss = fmaf(d, d, ss); }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) { const float d = v[j] - mean; ss = fmaf(d, d, ss); }
    }
  };

  // pass 1: v = acc/scale + bias (+row_bias) (+r_pre); act_pre
  if (p.r_pre != nullptr) warp_load_issue(pre, p.r_pre, p.ld_r_pre, row0, 0, p.M, p.N, lane);
  for (int c0 = 0; c0 < p.N; c0 += 32) {
    tmem_ld32(t_row + (uint32_t)c0, v);
    if (p.bias != nullptr) {
      lds_row32(ev.bias + c0, t);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], inv_scale, t[j]);
    } else if (p.a_scale != 1.0f) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= inv_scale;
    }
    if (p.row_bias != nullptr) {
      const int rbi = row < p.M ? p.row_bias_index[row] : 0;
      const float* rb = p.row_bias + (size_t)rbi * p.ld_row_bias + c0;
      const int nv = (c0 + 32 > p.N) ? 4 : 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < nv) {
          const float4 q = ldg4(rb + 4 * j);
          v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
        }
      }
    }
    if (p.r_pre != nullptr) {
      warp_load_commit(st, t, pre, lane);
      if (c0 + 32 < p.N) warp_load_issue(pre, p.r_pre, p.ld_r_pre, row0, c0 + 32, p.M, p.N, lane);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += t[j];
    }
    if (p.act_pre != DGDM_ACT_NONE) apply_act_chunk<32>(v, p.act_pre);
    if (c0 + 32 > p.N) {          // N % 16 == 0: only the upper half of the last chunk can lie outside the row
#pragma unroll
      for (int j = 16; j < 32; ++j) v[j] = 0.f;
    }
    if (writeback) {
#pragma unroll
      for (int j = 0; j < 32; ++j) sum += v[j];
      tmem_st32(t_row + (uint32_t)c0, v);
    } else {
      // r_pre and r_post together without LayerNorm do not occur in the model; keep it correct with an unpipelined load
      if (p.r_post != nullptr && (c0 == 0 || p.r_pre != nullptr)) warp_load_issue(pre, p.r_post, p.ld_r_post, row0, c0, p.M, p.N, lane);
      finish(c0);
      if (p.r_post != nullptr && p.r_pre != nullptr && c0 + 32 < p.N) warp_load_issue(pre, p.r_pre, p.ld_r_pre, row0, c0 + 32, p.M, p.N, lane);
    }
  }
  if (writeback) {
    const float* g = has_ln1 ? ev.g1 : ev.g2;
    const float* b = has_ln1 ? ev.b1 : ev.b2;
    float mean = sum * invN, ss = 0.f;
    for (int c0 = 0; c0 < p.N; c0 += 32) {
      tmem_ld32(t_row + (uint32_t)c0, v);
      sq_dev(c0, mean, ss);
    }
    float rstd = rsqrtf(ss * invN + p.ln_eps);
    if (has_ln1 && has_ln2) {
      float sum2 = 0.f;
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        tmem_ld32(t_row + (uint32_t)c0, v);
        normalise(c0, g, b, mean, rstd);
#pragma unroll
        for (int j = 0; j < 32; ++j) sum2 += v[j];
        tmem_st32(t_row + (uint32_t)c0, v);
      }
      mean = sum2 * invN; ss = 0.f;
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        tmem_ld32(t_row + (uint32_t)c0, v);
        sq_dev(c0, mean, ss);
      }
      rstd = rsqrtf(ss * invN + p.ln_eps);
      g = ev.g2; b = ev.b2;
    }
    if (p.r_post != nullptr) warp_load_issue(pre, p.r_post, p.ld_r_post, row0, 0, p.M, p.N, lane);
    for (int c0 = 0; c0 < p.N; c0 += 32) {
      tmem_ld32(t_row + (uint32_t)c0, v);
      normalise(c0, g, b, mean, rstd);
      finish(c0);
    }
  }
  if (p.dot_out != nullptr && row < p.M) p.dot_out[row] = apply_act(dot + p.dot_b, p.dot_act);
}

template <bool PEERS>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_f16x2_kernel(const dgdm_gemm_params p, const GemmSched sc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_ring = base;
  const uint32_t w_ring = base + (uint32_t)sc.na * A_SLOT_BYTES;
  float* staging = reinterpret_cast<float*>(base_ptr + GEMM_RING_BYTES);
  float* vecs = reinterpret_cast<float*>(base_ptr + GEMM_RING_BYTES + GEMM_STAGING_BYTES);
  const uint32_t bar_base = base + GEMM_RING_BYTES + GEMM_STAGING_BYTES + GEMM_VEC_BYTES;
  const uint32_t bar_full_a = bar_base, bar_empty_a = bar_base + 32, bar_full_w = bar_base + 64, bar_empty_w = bar_base + 96;
  const uint32_t bar_acc_full = bar_base + 128, bar_acc_empty = bar_base + 144;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + GEMM_RING_BYTES + GEMM_STAGING_BYTES + GEMM_VEC_BYTES + 192);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef DGDM_GEMM_TRACE
  const long long t_kernel0 = clock64();
#endif
  const int nkb = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int nchunk = (p.N + GEMM_NCHUNK - 1) / GEMM_NCHUNK;

  if (tid == 0) {
    for (int i = 0; i < GEMM_MAX_SLOTS; ++i) {
      mbar_init(bar_full_a + 8 * i, GEMM_CONV_THREADS);
      mbar_init(bar_empty_a + 8 * i, 1);
      mbar_init(bar_full_w + 8 * i, 1);
      mbar_init(bar_empty_w + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, GEMM_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 13) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
  {   // epilogue column vectors -> smem (absent vectors: zeros; columns >= N: zeros)
    auto stage_vec = [&](const float* src, int k) {
      for (int c = tid; c < 512; c += GEMM_THREADS) vecs[k * 512 + c] = (src != nullptr && c < p.N) ? __ldg(src + c) : 0.f;
    };
    stage_vec(p.bias, 0); stage_vec(p.ln1_g, 1); stage_vec(p.ln1_b, 2); stage_vec(p.ln2_g, 3); stage_vec(p.ln2_b, 4); stage_vec(p.dot_w, 5);
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) TRACE_EVENT(11, 0);

  if (warp < GEMM_EPI_WARPS) {
    // ------------------------------------------------------------------------------------------ epilogue warps
    float* st = staging + warp * (32 * TR_STRIDE);
    const EpiVec ev = {vecs, vecs + 512, vecs + 1024, vecs + 1536, vecs + 2048, vecs + 2560};
    int t = 0;
    for (int tile = blockIdx.x; tile < sc.ntiles; tile += gridDim.x, ++t) {
      const int buf = t % sc.nbuf, use = t / sc.nbuf;
      mbar_wait_warp(bar_acc_full + 8 * buf, (uint32_t)(use & 1), lane);
      fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * GEMM_NCHUNK);
      if (tid == 0) TRACE_EVENT(2, t);
      epilogue_rows<PEERS>(p, ev, st, t_row, tile * GEMM_BM + warp * 32, lane);
      if (tid == 0) TRACE_EVENT(3, t);
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf);
    }
  } else if (warp < GEMM_EPI_WARPS + GEMM_CONV_WARPS) {
    // ------------------------------------------------------------------------------------------ A converters
    const int ct = tid - GEMM_EPI_WARPS * 32;
    const int a_f4 = ct & 15, a_r0 = ct >> 4;       // float4 column 0..15, row offset 0..15
    const int my_tiles = (sc.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int items = my_tiles * nkb;
    // Rolling two-item register window: while item i is converted row by row, each freed register is refilled at once
    // with the same row of item i+2, so ~12-16 float4 (up to 64 KB per SM) of global loads stay in flight.
    float4 b0[8], b1[8];
    struct ItemAddr { const float* ptr; int rows_left; bool kok; };
    auto item_addr = [&](int item) {
      ItemAddr ia;
      const int tl = item / nkb, kb = item - tl * nkb;
      const int m0 = ((int)blockIdx.x + tl * (int)gridDim.x) * GEMM_BM + a_r0;
      const int kcol = kb * GEMM_BK + a_f4 * 4;
      ia.ptr = p.A + (size_t)m0 * p.lda + kcol;
      ia.rows_left = p.M - m0;                   // row i of the item is valid iff 16*i < rows_left
      ia.kok = item < items && kcol < p.K;
      return ia;
    };
    auto load_row = [&](const ItemAddr& ia, int i) {
      return (ia.kok && 16 * i < ia.rows_left) ? ldg4(ia.ptr + (size_t)(16 * i) * p.lda) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto process = [&](int item, float4* buf) {
      const ItemAddr nx = item_addr(item + 2);
      const int slot = item % sc.na, use = item / sc.na;
      mbar_wait_poll(bar_empty_a + 8 * slot, (uint32_t)((use & 1) ^ 1));
      if (ct == 0) TRACE_EVENT(4, item);
      uint8_t* dst = base_ptr + slot * A_SLOT_BYTES;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = a_r0 + 16 * i;
        float4 x = buf[i];
        if (p.a_scale != 1.0f) { x.x *= p.a_scale; x.y *= p.a_scale; x.z *= p.a_scale; x.w *= p.a_scale; }
        uint2 hi, lo;
        split4(x, hi, lo);
        const uint32_t off = sw128_off(r, a_f4 >> 1) + (uint32_t)(a_f4 & 1) * 8;
        *reinterpret_cast<uint2*>(dst + off) = hi;
        *reinterpret_cast<uint2*>(dst + A_PLANE_BYTES + off) = lo;
        buf[i] = load_row(nx, i);
      }
      if (ct == 0) TRACE_EVENT(5, item);
      fence_async_smem();
      if (ct == 0) TRACE_EVENT(6, item);
      mbar_arrive(bar_full_a + 8 * slot);      // per-thread arrive: a warp-elected arrive behind __syncwarp measured 2x slower
      if (ct == 0) TRACE_EVENT(0, item);
    };
    {
      const ItemAddr i0 = item_addr(0), i1 = item_addr(1);
#pragma unroll
      for (int i = 0; i < 8; ++i) b0[i] = load_row(i0, i);
#pragma unroll
      for (int i = 0; i < 8; ++i) b1[i] = load_row(i1, i);
    }
    for (int item = 0; item < items; item += 2) {
      process(item, b0);
      if (item + 1 < items) process(item + 1, b1);
    }
  } else if (warp == 12) {
    // ------------------------------------------------------------------------------------------ W producer
    if (elect_one()) {
      const uint8_t* Wimg = reinterpret_cast<const uint8_t*>(p.W_hi);
      int unit = 0;
      for (int tile = blockIdx.x; tile < sc.ntiles; tile += gridDim.x) {
        for (int kb = 0; kb < nkb; ++kb) {
          for (int ch = 0; ch < nchunk; ++ch, ++unit) {
            const int n0 = ch * GEMM_NCHUNK;
            const int nc = min(GEMM_NCHUNK, p.N - n0);
            const int slot = unit % sc.nw, use = unit / sc.nw;
            mbar_wait_poll(bar_empty_w + 8 * slot, (uint32_t)((use & 1) ^ 1));
            TRACE_EVENT(7, unit);
            const uint32_t bytes = (uint32_t)nc * 128u;
            const uint32_t dst = w_ring + (uint32_t)slot * (uint32_t)sc.w_slot_bytes;
            const uint8_t* src = Wimg + (size_t)kb * 2 * (size_t)sc.w_plane_stride + (size_t)n0 * 128;
            mbar_arrive_expect_tx(bar_full_w + 8 * slot, 2 * bytes);
            bulk_g2s(dst, src, bytes, bar_full_w + 8 * slot);
            bulk_g2s(dst + (uint32_t)(sc.w_slot_bytes >> 1), src + sc.w_plane_stride, bytes, bar_full_w + 8 * slot);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int item = 0, unit = 0, t = 0;
      for (int tile = blockIdx.x; tile < sc.ntiles; tile += gridDim.x, ++t) {
        const int buf = t % sc.nbuf, ause = t / sc.nbuf;
        mbar_wait_poll(bar_acc_empty + 8 * buf, (uint32_t)((ause & 1) ^ 1));
        fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++item) {
          const int sa = item % sc.na;
          mbar_wait_poll(bar_full_a + 8 * sa, (uint32_t)((item / sc.na) & 1));
          TRACE_EVENT(8, item);
          const uint32_t a_hi = a_ring + (uint32_t)sa * A_SLOT_BYTES, a_lo = a_hi + A_PLANE_BYTES;
          for (int ch = 0; ch < nchunk; ++ch, ++unit) {
            const int n0 = ch * GEMM_NCHUNK;
            const int nc = min(GEMM_NCHUNK, p.N - n0);
            const int sw = unit % sc.nw;
            mbar_wait_poll(bar_full_w + 8 * sw, (uint32_t)((unit / sc.nw) & 1));
            TRACE_EVENT(9, unit);
            fence_after();
            const uint32_t w_hi = w_ring + (uint32_t)sw * (uint32_t)sc.w_slot_bytes, w_lo = w_hi + (uint32_t)(sc.w_slot_bytes >> 1);
            const uint32_t idesc = idesc_f16(GEMM_BM, nc);
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * GEMM_NCHUNK + n0);
#pragma unroll
            for (int ks = 0; ks < GEMM_BK / 16; ++ks) {
              const uint32_t ko = ks * 32;
              mma_ss(d_tmem, desc_sw128(a_lo + ko), desc_sw128(w_hi + ko), idesc, (kb | ks) != 0);
              mma_ss(d_tmem, desc_sw128(a_hi + ko), desc_sw128(w_lo + ko), idesc, 1);
              mma_ss(d_tmem, desc_sw128(a_hi + ko), desc_sw128(w_hi + ko), idesc, 1);
            }
            mma_commit(bar_empty_w + 8 * sw);
          }
          mma_commit(bar_empty_a + 8 * sa);
          TRACE_EVENT(10, item);
        }
        mma_commit(bar_acc_full + 8 * buf);
        TRACE_EVENT(1, t);
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 13) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
}

// fp32 W[N,K] -> pre-swizzled fp16 K-block images: img[kb][plane][N rows x 128 B] (SWIZZLE_128B canonical, K-major),
// element (n, k = kb*64 + c*8 + e) of plane pl at  ((kb*2 + pl)*N)*128 + sw128_off(n, c) + e*2.  K is zero padded to Kpad.
__global__ void split_weight_kernel(const float* __restrict__ W, int N, int K, uint8_t* __restrict__ img, int Kpad) {
  const size_t total = (size_t)N * (Kpad / 8);                 // one thread per 16-byte chunk (8 elements)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / (Kpad / 8)), kc = (int)(i - (size_t)n * (Kpad / 8));
    const int kb = kc >> 3, c = kc & 7;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k0 = kc * 8 + 2 * e;
      const float w0 = (k0 < K) ? W[(size_t)n * K + k0] : 0.f;
      const float w1 = (k0 + 1 < K) ? W[(size_t)n * K + k0 + 1] : 0.f;
      const __half h0 = __float2half_rn(w0), h1 = __float2half_rn(w1);
      hi[e] = pack_half2(h0, h1);
      lo[e] = pack_half2(__float2half_rn(w0 - __half2float(h0)), __float2half_rn(w1 - __half2float(h1)));
    }
    const size_t off = ((size_t)kb * 2 * N) * 128 + sw128_off(n, c);
    *reinterpret_cast<uint4*>(img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(img + off + (size_t)N * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

static int g_num_sms = 0;

static int launch_gemm(const dgdm_gemm_params& p, long long plane_stride, cudaStream_t stream) {
  static PerDeviceOnce attr_set;
  if (attr_set.needed()) {
    if (cudaFuncSetAttribute(gemm_f16x2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(gemm_f16x2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM) != cudaSuccess)
      return DGDM_ERR_LAUNCH;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_num_sms <= 0) g_num_sms = 148;
    attr_set.mark();
  }
  GemmSched sc;
  const int ch_rows = p.N < GEMM_NCHUNK ? p.N : GEMM_NCHUNK;
  sc.w_slot_bytes = 2 * ch_rows * 128;
  if (sc.w_slot_bytes > 32768) { sc.na = 2; sc.nw = 2; }          // 64 + 128 KB
  else if (sc.w_slot_bytes > 16384) { sc.na = 3; sc.nw = 3; }     // 96 + 96 KB
  else { sc.na = 4; sc.nw = 4; }                                  // 128 + <= 64 KB
  sc.nbuf = p.N <= GEMM_NCHUNK ? 2 : 1;
  sc.ntiles = (p.M + GEMM_BM - 1) / GEMM_BM;
  sc.w_plane_stride = plane_stride;
  const int grid = sc.ntiles < g_num_sms ? sc.ntiles : g_num_sms;
  if (p.n_peers > 0) gemm_f16x2_kernel<true><<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(p, sc);
  else gemm_f16x2_kernel<false><<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(p, sc);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

}  // namespace dgdm

using namespace dgdm;

#ifdef DGDM_GEMM_TRACE
extern "C" int dgdm_debug_gemm_trace(long long* host_out, int reset) {
  if (host_out && cudaMemcpyFromSymbol(host_out, g_gemm_events, sizeof(long long) * 12 * 64) != cudaSuccess) return DGDM_ERR_LAUNCH;
  if (reset) { static long long z[12 * 64]; if (cudaMemcpyToSymbol(g_gemm_events, z, sizeof(z)) != cudaSuccess) return DGDM_ERR_LAUNCH; }
  return DGDM_OK;
}
#endif

extern "C" int dgdm_gemm_f16x2(const dgdm_gemm_params* p, void* stream) {
  if (!p || !p->A || !p->W_hi || !p->W_lo) return DGDM_ERR_INVALID;
  if (p->N < 16 || (p->N % 16) != 0) return DGDM_ERR_INVALID;
  if (p->K < 4 || (p->K % 4) != 0 || (p->lda % 4) != 0 || (p->ldw % 64) != 0 || p->ldw < p->K) return DGDM_ERR_INVALID;
  if ((const uint8_t*)p->W_lo != (const uint8_t*)p->W_hi + (size_t)p->N * 128) return DGDM_ERR_INVALID;   // image layout of dgdm_split_weight_f16
  if (p->M <= 0) return DGDM_OK;
  if (p->C && (p->ldc % 4) != 0) return DGDM_ERR_INVALID;
  if (p->r_pre && (p->ld_r_pre % 4) != 0) return DGDM_ERR_INVALID;
  if (p->r_post && (p->ld_r_post % 4) != 0) return DGDM_ERR_INVALID;
  if ((p->ln1_g == nullptr) != (p->ln1_b == nullptr) || (p->ln2_g == nullptr) != (p->ln2_b == nullptr)) return DGDM_ERR_INVALID;
  if (!(p->a_scale > 0.f)) return DGDM_ERR_INVALID;
  if (p->n_peers < 0 || p->n_peers > 8 || (p->n_peers > 0 && (!p->C || p->N > 512))) return DGDM_ERR_INVALID;
  for (int d = 0; d < p->n_peers; ++d) if (!p->C_peers[d]) return DGDM_ERR_INVALID;
  const long long plane_stride = (long long)p->N * 128;
  if (p->N <= 512) return launch_gemm(*p, plane_stride, (cudaStream_t)stream);
  // Rows wider than TMEM (DGDM-Large 768-d layers): raw products in column groups of <= 512, then the stand-alone
  // row-wise epilogue over the full width.
  if (!p->C) return DGDM_ERR_INVALID;
  for (int g0 = 0; g0 < p->N; g0 += 512) {
    dgdm_gemm_params raw = {};
    raw.A = p->A; raw.lda = p->lda; raw.M = p->M; raw.K = p->K;
    raw.W_hi = (const uint8_t*)p->W_hi + (size_t)g0 * 128;
    raw.W_lo = (const uint8_t*)p->W_lo + (size_t)g0 * 128;
    raw.ldw = p->ldw;
    raw.N = (p->N - g0) < 512 ? (p->N - g0) : 512;
    raw.C = p->C + g0; raw.ldc = p->ldc;
    raw.a_scale = p->a_scale; raw.ln_eps = p->ln_eps;
    const int rc = launch_gemm(raw, plane_stride, (cudaStream_t)stream);
    if (rc != DGDM_OK) return rc;
  }
  dgdm_gemm_params ep = *p;
  ep.a_scale = 1.0f;        // the raw pass already divided the accumulators by a_scale
  return dgdm_rowwise_epilogue(p->C, p->ldc, p->M, p->N, &ep, stream);
}

// Same kernel on a row sub-range [row0, row0 + p->N) of a larger weight image (p->N <= 512): W_hi points at row0 of the hi plane
// of K-block 0 and w_plane_stride_bytes = (rows of the whole image) * 128. Used by the Gram-matrix candidate pass of the edge builder.
extern "C" int dgdm_gemm_f16x2_strided(const dgdm_gemm_params* p, long long w_plane_stride_bytes, void* stream) {
  if (!p || !p->A || !p->W_hi || !p->C) return DGDM_ERR_INVALID;
  if (p->N < 16 || p->N > 512 || (p->N % 16) != 0 || w_plane_stride_bytes < (long long)p->N * 128) return DGDM_ERR_INVALID;
  if (p->K < 4 || (p->K % 4) != 0 || (p->lda % 4) != 0 || (p->ldw % 64) != 0 || p->ldw < p->K || (p->ldc % 4) != 0) return DGDM_ERR_INVALID;
  if (!(p->a_scale > 0.f) || p->n_peers != 0) return DGDM_ERR_INVALID;
  if (p->M <= 0) return DGDM_OK;
  return launch_gemm(*p, w_plane_stride_bytes, (cudaStream_t)stream);
}

extern "C" int dgdm_split_weight_f16(const float* W, int N, int K, void* W_hi, void* W_lo, int Kpad, void* stream) {
  if (!W || !W_hi || !W_lo || N <= 0 || (N % 8) != 0 || K <= 0 || Kpad < K || (Kpad % 64) != 0) return DGDM_ERR_INVALID;
  if ((uint8_t*)W_lo != (uint8_t*)W_hi + (size_t)N * 128) return DGDM_ERR_INVALID;
  const size_t total = (size_t)N * (Kpad / 8);
  const int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
  split_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, N, K, (uint8_t*)W_hi, Kpad);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}
