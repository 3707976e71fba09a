// tcgen05 / TMEM GEMM with fused row-wise epilogue for the DGDM node-level linear layers.
//
//   C[M,N] = epilogue( A[M,K] (fp32) @ W[N,K]^T )     N <= 512, whole output rows live in TMEM (512 columns)
//
// Precision: A and W are each split into two fp16 planes (hi, lo); D += Ahi*Whi + Alo*Whi + Ahi*Wlo with fp32
// accumulation in TMEM (3 tcgen05.mma kind::f16 per K-step) -> ~2^-21 relative, i.e. fp32-class results.
// Layout: one CTA = 128 rows x all N columns. K is walked in 64-wide blocks, 2-stage smem ring:
//   A block: fp32 global -> registers (128-bit coalesced) -> hi/lo fp16 -> smem, K-major SWIZZLE_128B canonical layout
//   W block: pre-split fp16 planes [N,Kpad] -> cp.async 16 B -> same canonical layout
// One elected thread issues the MMAs; tcgen05.commit on an mbarrier frees the stage. N > 256 is covered by
// N-passes of <=256 columns into adjacent TMEM column ranges. The epilogue keeps one output row per thread
// (TMEM lane == row), so LayerNorm statistics are thread-local; global I/O goes through a per-warp smem
// transpose so that every global access is a full 128-byte row segment.
#include "common.cuh"

namespace dgdm {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_NCHUNK = 256;
constexpr int GEMM_THREADS = 256;
constexpr int A_PLANE_BYTES = GEMM_BM * GEMM_BK * 2;        // 16 KB
constexpr int W_PLANE_BYTES = GEMM_NCHUNK * GEMM_BK * 2;    // 32 KB
constexpr int STAGE_BYTES = 2 * A_PLANE_BYTES + 2 * W_PLANE_BYTES;   // 96 KB
constexpr int GEMM_SMEM = 2 * STAGE_BYTES + 1024 /*align*/ + 64 /*barriers*/;

// ---- PTX wrappers --------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a wrong descriptor must not hang the GPU — after ~2^28 polls we trap.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 28)) { asm volatile("trap;"); }
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// K-major SWIZZLE_128B canonical layout: 8-row x 128-byte atoms, 1024 B apart (SBO); LBO unused.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                         // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset = 1024 B
  d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}
// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a [rows x 64 fp16] SW128 tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

__device__ __forceinline__ uint32_t pack_half2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
__device__ __forceinline__ void split4(float4 v, float s, uint2& hi, uint2& lo) {
  float x0 = v.x * s, x1 = v.y * s, x2 = v.z * s, x3 = v.w * s;
  __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1), h2 = __float2half_rn(x2), h3 = __float2half_rn(x3);
  __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
  __half l2 = __float2half_rn(x2 - __half2float(h2)), l3 = __float2half_rn(x3 - __half2float(h3));
  hi = make_uint2(pack_half2(h0, h1), pack_half2(h2, h3));
  lo = make_uint2(pack_half2(l0, l1), pack_half2(l2, l3));
}

// ---- per-warp 32x32 transposing I/O through smem (stride 36 floats, conflict-free for 128-bit accesses) ----
constexpr int TR_STRIDE = 36;
// thread-row-owned regs v[32] (row = lane) -> coalesced global store of rows [row0, row0+32), cols [c0, c0+32)
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

__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_f16x2_kernel(const dgdm_gemm_params p, uint32_t tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_base = base + 2 * STAGE_BYTES;     // [0],[1]: stage free; [2]: all MMAs done
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + 2 * STAGE_BYTES + 32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * GEMM_BM;
  const int nkb = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int npass = (p.N + GEMM_NCHUNK - 1) / GEMM_NCHUNK;
  const int total = nkb * npass;

  if (tid == 0) {
    mbar_init(bar_base + 0, 1); mbar_init(bar_base + 8, 1); mbar_init(bar_base + 16, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const __half* Whi = reinterpret_cast<const __half*>(p.W_hi);
  const __half* Wlo = reinterpret_cast<const __half*>(p.W_lo);
  const int a_f4 = tid & 15, a_r0 = tid >> 4;       // float4 column 0..15, row offset 0..15

  for (int it = 0; it < total; ++it) {
    const int pass = it / nkb, kb = it - pass * nkb;
    const int s = it & 1;
    const int n0 = pass * GEMM_NCHUNK;
    const int nc = min(GEMM_NCHUNK, p.N - n0);
    const uint32_t st_base = base + s * STAGE_BYTES;
    if (it >= 2) mbar_wait(bar_base + 8 * s, ((it >> 1) - 1) & 1);   // MMAs that read this stage have retired

    // W planes via cp.async (nc rows x 8 chunks x 2 planes)
    {
      const int nchunks = nc * 8;
      for (int i = tid; i < 2 * nchunks; i += GEMM_THREADS) {
        const int plane = i >= nchunks;
        const int j = plane ? i - nchunks : i;
        const int r = j >> 3, c = j & 7;
        const __half* src = (plane ? Wlo : Whi) + (size_t)(n0 + r) * p.ldw + kb * GEMM_BK + c * 8;
        cp_async16(st_base + 2 * A_PLANE_BYTES + plane * W_PLANE_BYTES + sw128_off(r, c), src);
      }
    }
    // A: fp32 -> (hi, lo) fp16
    {
      float4 v[8];
      const int kcol = kb * GEMM_BK + a_f4 * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = a_r0 + 16 * i;
        const int gm = m0 + r;
        v[i] = (gm < p.M && kcol < p.K) ? ldg4(p.A + (size_t)gm * p.lda + kcol) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = a_r0 + 16 * i;
        uint2 hi, lo;
        split4(v[i], p.a_scale, hi, lo);
        const uint32_t off = sw128_off(r, a_f4 >> 1) + (a_f4 & 1) * 8;
        *reinterpret_cast<uint2*>(base_ptr + s * STAGE_BYTES + off) = hi;
        *reinterpret_cast<uint2*>(base_ptr + s * STAGE_BYTES + A_PLANE_BYTES + off) = lo;
      }
    }
    cp_async_wait_all();
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = (1u << 4) | ((uint32_t)(nc >> 3) << 17) | ((uint32_t)(GEMM_BM >> 4) << 24);
      const uint32_t d_tmem = tmem_base + (uint32_t)n0;
      const uint32_t a_hi = st_base, a_lo = st_base + A_PLANE_BYTES;
      const uint32_t w_hi = st_base + 2 * A_PLANE_BYTES, w_lo = w_hi + W_PLANE_BYTES;
#pragma unroll
      for (int ks = 0; ks < GEMM_BK / 16; ++ks) {
        const uint32_t ko = ks * 32;
        umma_f16(d_tmem, make_desc_sw128(a_lo + ko), make_desc_sw128(w_hi + ko), idesc, (kb | ks) != 0);
        umma_f16(d_tmem, make_desc_sw128(a_hi + ko), make_desc_sw128(w_lo + ko), idesc, 1);
        umma_f16(d_tmem, make_desc_sw128(a_hi + ko), make_desc_sw128(w_hi + ko), idesc, 1);
      }
      umma_commit(bar_base + 8 * s);
      if (it == total - 1) umma_commit(bar_base + 16);
    }
  }
  mbar_wait(bar_base + 16, 0);
  tc_fence_after();
  __syncthreads();   // every MMA retired: pipeline smem is free for the epilogue staging tiles

  // ------------------------------------------------------------------ epilogue: warps 0..3, thread = row
  if (warp < 4) {
    float* st = reinterpret_cast<float*>(base_ptr) + warp * (32 * TR_STRIDE);
    const int row0 = m0 + warp * 32;
    const int row = row0 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float inv_scale = 1.0f / p.a_scale;
    const bool has_ln1 = p.ln1_g != nullptr, has_ln2 = p.ln2_g != nullptr;
    const bool writeback = has_ln1 || has_ln2;
    const float invN = 1.0f / (float)p.N;
    const int rbi = (p.row_bias_index != nullptr && row < p.M) ? p.row_bias_index[row] : 0;
    float v[32], t[32];
    float sum = 0.f, dot = 0.f;

    auto finish = [&](int c0) {   // v holds the post-LN values of columns [c0, c0+32)
      apply_act_chunk<32>(v, p.act_post);
      if (p.r_post != nullptr) {
        warp_load_tile(st, t, p.r_post, p.ld_r_post, row0, c0, p.M, p.N, lane);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += t[j];
      }
      if (p.dot_w != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (c0 + j < p.N) dot = fmaf(v[j], __ldg(p.dot_w + c0 + j), dot);
      }
      if (p.C != nullptr) warp_store_tile(st, v, p.C, p.ldc, row0, c0, p.M, p.N, lane);
    };

    // pass 1: v = acc/scale + bias (+row_bias) (+r_pre); act_pre
    for (int c0 = 0; c0 < p.N; c0 += 32) {
      tmem_ld32(t_row + (uint32_t)c0, v);
      if (p.r_pre != nullptr) warp_load_tile(st, t, p.r_pre, p.ld_r_pre, row0, c0, p.M, p.N, lane);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int c = c0 + j;
        float x = v[j] * inv_scale;
        if (c < p.N) {
          if (p.bias != nullptr) x += __ldg(p.bias + c);
          if (p.row_bias != nullptr) x += __ldg(p.row_bias + (size_t)rbi * p.ld_row_bias + c);
          if (p.r_pre != nullptr) x += t[j];
        } else x = 0.f;
        v[j] = x;
      }
      if (p.act_pre != DGDM_ACT_NONE) {
        apply_act_chunk<32>(v, p.act_pre);
#pragma unroll
        for (int j = 0; j < 32; ++j) if (c0 + j >= p.N) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) sum += v[j];
      if (writeback) tmem_st32(t_row + (uint32_t)c0, v);
      else finish(c0);
    }
    if (writeback) {
      const float* g = has_ln1 ? p.ln1_g : p.ln2_g;
      const float* b = has_ln1 ? p.ln1_b : p.ln2_b;
      float mean = sum * invN, ss = 0.f;
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        tmem_ld32(t_row + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) if (c0 + j < p.N) { float d = v[j] - mean; ss = fmaf(d, d, ss); }
      }
      float rstd = rsqrtf(ss * invN + p.ln_eps);
      if (has_ln1 && has_ln2) {
        float sum2 = 0.f;
        for (int c0 = 0; c0 < p.N; c0 += 32) {
          tmem_ld32(t_row + (uint32_t)c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = c0 + j;
            float y = (c < p.N) ? (v[j] - mean) * rstd * __ldg(g + c) + __ldg(b + c) : 0.f;
            v[j] = y; sum2 += y;
          }
          tmem_st32(t_row + (uint32_t)c0, v);
        }
        mean = sum2 * invN; ss = 0.f;
        for (int c0 = 0; c0 < p.N; c0 += 32) {
          tmem_ld32(t_row + (uint32_t)c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) if (c0 + j < p.N) { float d = v[j] - mean; ss = fmaf(d, d, ss); }
        }
        rstd = rsqrtf(ss * invN + p.ln_eps);
        g = p.ln2_g; b = p.ln2_b;
      }
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        tmem_ld32(t_row + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = c0 + j;
          v[j] = (c < p.N) ? (v[j] - mean) * rstd * __ldg(g + c) + __ldg(b + c) : 0.f;
        }
        finish(c0);
      }
    }
    if (p.dot_out != nullptr && row < p.M) p.dot_out[row] = apply_act(dot + p.dot_b, p.dot_act);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// fp32 W[N,K] -> fp16 hi/lo planes [N,Kpad]
__global__ void split_weight_kernel(const float* __restrict__ W, int N, int K, __half* __restrict__ hi,
                                    __half* __restrict__ lo, int Kpad) {
  const size_t total = (size_t)N * Kpad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / Kpad), k = (int)(i - (size_t)n * Kpad);
    float w = (k < K) ? W[(size_t)n * K + k] : 0.f;
    __half h = __float2half_rn(w);
    hi[i] = h;
    lo[i] = __float2half_rn(w - __half2float(h));
  }
}

}  // namespace dgdm

using namespace dgdm;

extern "C" int dgdm_gemm_f16x2(const dgdm_gemm_params* p, void* stream) {
  if (!p || !p->A || !p->W_hi || !p->W_lo) return DGDM_ERR_INVALID;
  if (p->N < 16 || p->N > 512 || (p->N % 16) != 0) return DGDM_ERR_INVALID;
  if (p->K < 4 || (p->K % 4) != 0 || (p->lda % 4) != 0 || (p->ldw % 64) != 0 || p->ldw < p->K) return DGDM_ERR_INVALID;
  if (p->M <= 0) return DGDM_OK;
  if (p->C && (p->ldc % 4) != 0) return DGDM_ERR_INVALID;
  if (p->r_pre && (p->ld_r_pre % 4) != 0) return DGDM_ERR_INVALID;
  if (p->r_post && (p->ld_r_post % 4) != 0) return DGDM_ERR_INVALID;
  if ((p->ln1_g == nullptr) != (p->ln1_b == nullptr) || (p->ln2_g == nullptr) != (p->ln2_b == nullptr)) return DGDM_ERR_INVALID;
  if (!(p->a_scale > 0.f)) return DGDM_ERR_INVALID;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_f16x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM) != cudaSuccess)
      return DGDM_ERR_LAUNCH;
    attr_set = true;
  }
  uint32_t cols = 32;
  while ((int)cols < p->N) cols <<= 1;
  const int grid = (p->M + GEMM_BM - 1) / GEMM_BM;
  gemm_f16x2_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, (cudaStream_t)stream>>>(*p, cols);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" int dgdm_split_weight_f16(const float* W, int N, int K, void* W_hi, void* W_lo, int Kpad, void* stream) {
  if (!W || !W_hi || !W_lo || N <= 0 || K <= 0 || Kpad < K || (Kpad % 64) != 0) return DGDM_ERR_INVALID;
  const size_t total = (size_t)N * Kpad;
  const int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
  split_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, N, K, (__half*)W_hi, (__half*)W_lo, Kpad);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}
