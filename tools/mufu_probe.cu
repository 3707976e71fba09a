// Microbenchmark: what the B200 SM sustains for the attention softmax instruction mix.
// Per "pair" (two scores) the kernel's inner loop is: 3 FADD2, 2 MUFU.EX2, 2 LOP3, 1 FFMA2, 2 F2FP (+0.5 MUFU.SQRT amortised).
// Measures cycles per pair per warp for: pure EX2, the full mix, and the mix with one ingredient removed, at W warps per SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mufu_probe tools/mufu_probe.cu && tools/mufu_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned long long pk2(float a, float b) { return (unsigned long long)__float_as_uint(a) | ((unsigned long long)__float_as_uint(b) << 32); }
__device__ __forceinline__ float lo32(unsigned long long v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi32(unsigned long long v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }

// MODE bits: 1 = EX2, 2 = FADD2 x (2 + lsum), 4 = LOP3 + FFMA2 residual, 8 = F2FP packs
template <int MODE>
__global__ void __launch_bounds__(1024) probe(const float* __restrict__ in, uint32_t* __restrict__ out, int iters, long long* cyc) {
  constexpr int P = 8;                       // independent pairs in flight per thread (the real loop interleaves 8-16)
  unsigned long long s[P], nb[P];
  for (int i = 0; i < P; ++i) { s[i] = pk2(in[threadIdx.x + 32 * i], in[threadIdx.x + 32 * i + 7]); nb[i] = pk2(-0.01f * i, -0.02f * i); }
  const unsigned long long nmu = pk2(-0.5f, -0.5f), neg1 = pk2(-1.f, -1.f);
  unsigned long long ls = 0ull;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < P; ++i) {
      unsigned long long x = s[i];
      if (MODE & 2) x = add2(add2(x, nb[i]), nmu);
      float p0 = lo32(x), p1 = hi32(x);
      if (MODE & 1) { p0 = ex2(p0); p1 = ex2(p1); }
      const unsigned long long pp = pk2(p0, p1);
      if (MODE & 2) ls = add2(ls, pp);
      float h0 = p0, h1 = p1;
      unsigned long long ll = pp;
      if (MODE & 4) {
        h0 = __uint_as_float(__float_as_uint(p0) & 0xffffe000u); h1 = __uint_as_float(__float_as_uint(p1) & 0xffffe000u);
        ll = fma2(pk2(h0, h1), neg1, pp);
      }
      if (MODE & 8) { acc ^= pack_h2(h0, h1); acc += pack_h2(lo32(ll), hi32(ll)); }
      else { acc ^= __float_as_uint(h0) + __float_as_uint(hi32(ll)); }
      s[i] = pk2(p0 * 0.25f - 1.0f, p1 * 0.25f - 1.0f) ^ (unsigned long long)(it & 1);       // keep the chain data-dependent, values bounded
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (uint32_t)ls;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_smsp, const float* in, uint32_t* out, long long* cyc) {
  const int threads = warps_per_smsp * 4 * 32, iters = 2000;
  probe<MODE><<<148, threads>>>(in, out, 10, cyc);
  probe<MODE><<<148, threads>>>(in, out, iters, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_pair_per_warp = (double)c / (iters * 8.0);
  printf("%-44s W=%d  %.2f cycles per pair per warp  -> %.2f cycles per pair per SMSP-slot (x%d warps)\n", name, warps_per_smsp,
         per_pair_per_warp, per_pair_per_warp / warps_per_smsp, warps_per_smsp);
}

int main() {
  float* in; uint32_t* out; long long* cyc;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = -0.001f * (i % 97);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int w : {1, 2, 4, 8}) {
    if (w == 1) { run<1>("EX2 only (+1 FFMA/el chain)", 1, in, out, cyc); run<15>("full mix", 1, in, out, cyc); }
    if (w == 2) { run<1>("EX2 only", 2, in, out, cyc); run<15>("full mix", 2, in, out, cyc); run<7>("mix without F2FP", 2, in, out, cyc);
                  run<11>("mix without LOP3/FFMA2 residual", 2, in, out, cyc); run<13>("mix without FADD2", 2, in, out, cyc); run<14>("mix without EX2", 2, in, out, cyc); }
    if (w == 4) { run<1>("EX2 only", 4, in, out, cyc); run<15>("full mix", 4, in, out, cyc); run<7>("mix without F2FP", 4, in, out, cyc);
                  run<11>("mix without LOP3/FFMA2 residual", 4, in, out, cyc); run<13>("mix without FADD2", 4, in, out, cyc); run<14>("mix without EX2", 4, in, out, cyc); }
    if (w == 8) { run<1>("EX2 only", 8, in, out, cyc); run<15>("full mix", 8, in, out, cyc); }
  }
  printf("ideal XU-bound: 2 EX2 per pair x 8 cycles = 16 cycles per pair per SMSP-slot\n");
  return 0;
}
