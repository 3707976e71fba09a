// Bring-up probe for the tcgen05 building blocks of the spatial-attention kernel (run on the GPU box):
//   1. S_h = Q_h K_h^T with A/B from SW128 K-major smem tiles addressed at +32*h bytes (head slice), N = 64, K = 16
//   2. P (fp16 hi/lo) written IN PLACE over S with tcgen05.st, then O_h += P V_h with A from TMEM, B = V^T tile, N = 16
//   3. cycle counts of tcgen05.ld / tcgen05.st / small-N MMAs / mbarrier waits (clock64 inside one CTA)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc_probe tools/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../dgdm_histopath_lab_b200/csrc/tc05.cuh"

using namespace dgdm;
using namespace dgdm::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int BN = 64;          // keys per block
constexpr int NH = 4;           // heads in the probe (dims 0..63)

struct ProbeOut { float S[NH][128][BN]; float O[NH][128][16]; long long cyc[16]; };

// smem map (bytes, all 1024-aligned): Q 16K | K 8K | Vt NH*2K | bars
constexpr int OFF_Q = 0, OFF_K = 16384, OFF_V = OFF_K + BN * 128, OFF_BAR = OFF_V + NH * 2048, SMEM = OFF_BAR + 256 + 1024;

__global__ void __launch_bounds__(192, 1) probe_kernel(const __half* Q, const __half* K, const __half* V, ProbeOut* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* bp = raw + (base - smem_u32(raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar = base + OFF_BAR;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(bp + OFF_BAR + 128);
  if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 128); mbar_init(bar + 16, 1); mbar_init(bar + 24, 1); mbar_fence_init(); }
  if (warp == 4) tmem_alloc(smem_u32((const void*)slot), 512);
  // ---- stage tiles: Q [128 x 64] and K [BN x 64] as SW128 rows; Vt_h [16 dims x 64 keys]
  for (int i = tid; i < 128 * 8; i += 192) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(bp + OFF_Q + sw128_off(r, c)) = *reinterpret_cast<const uint4*>(Q + r * 64 + c * 8);
  }
  for (int i = tid; i < BN * 8; i += 192) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(bp + OFF_K + sw128_off(r, c)) = *reinterpret_cast<const uint4*>(K + r * 64 + c * 8);
  }
  for (int i = tid; i < NH * 16 * BN; i += 192) {       // element-wise transpose: V[key][dim] -> Vt_h[d][key]
    const int h = i / (16 * BN), d = (i / BN) % 16, key = i % BN;
    const uint32_t off = OFF_V + h * 2048 + sw128_off(d, key >> 3) + (key & 7) * 2;
    *reinterpret_cast<__half*>(bp + off) = V[key * 64 + h * 16 + d];
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tm = *slot;
  const uint32_t O_COL = 256;

  if (warp == 4) {
    if (elect_one()) {
      for (int h = 0; h < NH; ++h)
        mma_ss(tm + h * BN, desc_sw128(base + OFF_Q + 32 * h), desc_sw128(base + OFF_K + 32 * h), idesc_f16(128, BN), 0);
      mma_commit(bar);
      mbar_wait(bar + 8, 0);
      fence_after();
      for (int h = 0; h < NH; ++h)
        for (int k = 0; k < BN / 16; ++k) {
          const uint64_t vd = desc_sw128(base + OFF_V + h * 2048 + 32 * k);
          mma_ts(tm + O_COL + 16 * h, tm + h * BN + 16 * k + 8, vd, idesc_f16(128, 16), k > 0);   // lo
          mma_ts(tm + O_COL + 16 * h, tm + h * BN + 16 * k, vd, idesc_f16(128, 16), 1);           // hi
        }
      mma_commit(bar + 16);
    }
    __syncwarp();
  } else {
    const int row = warp * 32 + lane;
    const uint32_t trow = tm + ((uint32_t)(warp * 32) << 16);
    mbar_wait(bar, 0);
    fence_after();
    for (int h = 0; h < NH; ++h)
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32], pk[32];
        tmem_ld32_nowait(trow + h * BN + 32 * c, r);
        tmem_wait_ld();
        for (int j = 0; j < 32; ++j) out->S[h][row][32 * c + j] = __uint_as_float(r[j]);
        // P = S/16 split into hi/lo fp16, K-step interleaved: [hi(8 cols) lo(8 cols)] per 16 keys
        for (int ks = 0; ks < 2; ++ks)
          for (int j = 0; j < 8; ++j) {
            const float p0 = __uint_as_float(r[16 * ks + 2 * j]) * 0.0625f, p1 = __uint_as_float(r[16 * ks + 2 * j + 1]) * 0.0625f;
            const __half h0 = __float2half_rn(p0), h1 = __float2half_rn(p1);
            const __half l0 = __float2half_rn(p0 - __half2float(h0)), l1 = __float2half_rn(p1 - __half2float(h1));
            pk[16 * ks + j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
            pk[16 * ks + 8 + j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
          }
        tmem_st32_nowait(trow + h * BN + 32 * c, pk);
      }
    tmem_wait_st();
    fence_before();
    mbar_arrive(bar + 8);
    mbar_wait(bar + 16, 0);
    fence_after();
    uint32_t r[32];
    for (int c = 0; c < NH * 16 / 32; ++c) {
      tmem_ld32_nowait(trow + O_COL + 32 * c, r);
      tmem_wait_ld();
      for (int j = 0; j < 32; ++j) out->O[(32 * c + j) / 16][row][(32 * c + j) % 16] = __uint_as_float(r[j]);
    }
  }
  // ------------------------------------------------------------------ timing section
  fence_before();
  __syncthreads();
  fence_after();
  if (warp < 4) {
    const uint32_t trow = tm + ((uint32_t)(warp * 32) << 16);
    uint32_t r[32];
    long long t0 = clock64();
    uint32_t acc0 = 0;
    for (int i = 0; i < 64; ++i) { tmem_ld32_nowait(trow + (i & 7) * 32 + (acc0 & 1), r); tmem_wait_ld(); acc0 += r[0] & 2; }
    long long t1 = clock64();
    if (tid == 0) out->cyc[0] = (t1 - t0) / 64;            // ld.x32 + wait, 4 warps concurrently (latency-bound)
    uint32_t acc = acc0;
    t0 = clock64();
    for (int i = 0; i < 64; i += 4) {
      uint32_t a[32], b[32], c[32], d[32];
      tmem_ld32_nowait(trow + 0, a); tmem_ld32_nowait(trow + 32, b); tmem_ld32_nowait(trow + 64, c); tmem_ld32_nowait(trow + 96, d);
      tmem_wait_ld();
      for (int j = 0; j < 32; ++j) acc += a[j] ^ b[j] ^ c[j] ^ d[j];
    }
    t1 = clock64();
    if (tid == 0) out->cyc[1] = (t1 - t0) / 64;            // ld.x32 pipelined 4 deep (throughput)
    for (int j = 0; j < 32; ++j) r[j] = acc + j;
    t0 = clock64();
    for (int i = 0; i < 64; ++i) tmem_st32_nowait(trow + 384 + (i & 3) * 32, r);
    tmem_wait_st();
    t1 = clock64();
    if (tid == 0) out->cyc[2] = (t1 - t0) / 64;            // st.x32 throughput
    if (acc == 0x12345) out->cyc[15] = acc;
  }
  fence_before();
  __syncthreads();
  fence_after();
  if (warp == 4 && elect_one()) {
    long long t0 = clock64();
    for (int i = 0; i < 64; ++i)
      mma_ts(tm + O_COL, tm + 16 * (i & 3), desc_sw128(base + OFF_V + 32 * (i & 3)), idesc_f16(128, 16), 1);
    mma_commit(bar + 24);
    long long t1 = clock64();
    mbar_wait(bar + 24, 0);
    long long t2 = clock64();
    out->cyc[3] = (t1 - t0) / 64;                          // issue cost per TS MMA (N=16)
    out->cyc[4] = (t2 - t0) / 64;                          // completion per TS MMA (N=16, K=16)
    t0 = clock64();
    for (int i = 0; i < 64; ++i)
      mma_ss(tm + 64 * (i & 3), desc_sw128(base + OFF_Q + 32 * (i & 3)), desc_sw128(base + OFF_K + 32 * (i & 3)), idesc_f16(128, 64), 0);
    mma_commit(bar + 24);
    t1 = clock64();
    mbar_wait(bar + 24, 1);
    t2 = clock64();
    out->cyc[5] = (t1 - t0) / 64;                          // issue cost per SS MMA (N=64)
    out->cyc[6] = (t2 - t0) / 64;                          // completion per SS MMA (N=64, K=16)
    t0 = clock64();
    for (int i = 0; i < 64; ++i) { mma_commit(bar + 24); mbar_wait(bar + 24, i & 1); }
    t1 = clock64();
    out->cyc[7] = (t1 - t0) / 64;                          // empty commit + wait round trip
    t0 = clock64();
    for (int i = 0; i < 64; ++i) mbar_wait(bar + 24, 1);   // already-complete phase
    t1 = clock64();
    out->cyc[8] = (t1 - t0) / 64;
  }
  // ------------------------------------------------------------------ MMA issue under MUFU load (the attention kernel's situation:
  // the issuing thread shares its scheduler and the MIO queue with exp2-heavy softmax warps)
  volatile int* stop = reinterpret_cast<volatile int*>(bp + OFF_BAR + 192);
  if (tid == 0) *stop = 0;
  fence_before();
  __syncthreads();
  fence_after();
  if (warp < 4) {
    float x = 0.001f * tid, acc = 0.f;
    long long n = 0;
    while (*stop == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x + i * 0.01f)); acc += y; }
      x = acc * 1e-9f; ++n;
    }
    if (acc == 12345.f) out->cyc[15] = n;
  } else if (warp == 4 && elect_one()) {
    for (int rep = 0; rep < 2; ++rep) {
      long long t0 = clock64();
      for (int i = 0; i < 64; ++i)
        mma_ts(tm + O_COL, tm + 16 * (i & 3), desc_sw128(base + OFF_V + 32 * (i & 3)), idesc_f16(128, 16), 1);
      mma_commit(bar + 24);
      long long t1 = clock64();
      mbar_wait(bar + 24, rep & 1);
      long long t2 = clock64();
      out->cyc[9] = (t1 - t0) / 64;                        // issue cost per TS MMA (N=16) under MUFU load
      out->cyc[10] = (t2 - t0) / 64;
    }
    *stop = 1;
  }
  fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tm, 512);
}

int main() {
  std::vector<__half> Q(128 * 64), K(BN * 64), V(BN * 64);
  srand(1);
  auto rnd = [] { return (float)((rand() % 2001) - 1000) / 1000.0f; };
  for (auto& x : Q) x = __float2half(rnd());
  for (auto& x : K) x = __float2half(rnd());
  for (auto& x : V) x = __float2half(rnd());
  __half *dQ, *dK, *dV; ProbeOut* dO;
  CK(cudaMalloc(&dQ, Q.size() * 2)); CK(cudaMalloc(&dK, K.size() * 2)); CK(cudaMalloc(&dV, V.size() * 2)); CK(cudaMalloc(&dO, sizeof(ProbeOut)));
  CK(cudaMemcpy(dQ, Q.data(), Q.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dK, K.data(), K.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dV, V.data(), V.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0, sizeof(ProbeOut)));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  probe_kernel<<<1, 192, SMEM>>>(dQ, dK, dV, dO);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  ProbeOut* o = new ProbeOut;
  CK(cudaMemcpy(o, dO, sizeof(ProbeOut), cudaMemcpyDeviceToHost));
  double eS = 0, eO = 0, mS = 0, mO = 0;
  for (int h = 0; h < NH; ++h)
    for (int r = 0; r < 128; ++r) {
      double P[BN];
      for (int j = 0; j < BN; ++j) {
        double s = 0;
        for (int d = 0; d < 16; ++d) s += (double)__half2float(Q[r * 64 + h * 16 + d]) * (double)__half2float(K[j * 64 + h * 16 + d]);
        eS = fmax(eS, fabs(s - o->S[h][r][j])); mS = fmax(mS, fabs(s));
        const float p = o->S[h][r][j] * 0.0625f;
        const __half hi = __float2half_rn(p); const __half lo = __float2half_rn(p - __half2float(hi));
        P[j] = (double)__half2float(hi) + (double)__half2float(lo);
      }
      for (int d = 0; d < 16; ++d) {
        double acc = 0;
        for (int j = 0; j < BN; ++j) acc += P[j] * (double)__half2float(V[j * 64 + h * 16 + d]);
        eO = fmax(eO, fabs(acc - o->O[h][r][d])); mO = fmax(mO, fabs(acc));
      }
    }
  printf("S  max abs err %.3e (max |S| %.3f)  -> %s\n", eS, mS, eS < 1e-5 * mS ? "OK" : "MISMATCH");
  printf("O  max abs err %.3e (max |O| %.3f)  -> %s\n", eO, mO, eO < 1e-5 * mO ? "OK" : "MISMATCH");
  const char* names[] = {"tmem ld.x32+wait (latency)", "tmem ld.x32 4-deep (per instr)", "tmem st.x32 (per instr)", "TS MMA N=16 issue", "TS MMA N=16 complete",
                         "SS MMA N=64 issue", "SS MMA N=64 complete", "commit+wait round trip", "mbar wait (complete phase)",
                         "TS MMA N=16 issue, MUFU-loaded SM", "TS MMA N=16 complete, MUFU-loaded"};
  for (int i = 0; i < 11; ++i) printf("cycles  %-34s %lld\n", names[i], o->cyc[i]);
  return (eS < 1e-5 * mS && eO < 1e-5 * mO) ? 0 : 1;
}
