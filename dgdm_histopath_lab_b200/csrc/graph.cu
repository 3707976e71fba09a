// Graph-structure kernels: CSR-by-destination build (edge order kept), GCN normalisation, fused
// gather / scale / fixed-order accumulate / bias / activation SpMM, edge-attribute aggregation,
// exact top-k pooling (radix select + ordered compaction), pooled-CSR construction, per-destination
// segment softmax. HBM-bound integer/float work: warp per destination row, 128-bit coalesced feature-row
// loads, no atomics on the data path (deterministic results run to run).
#include "common.cuh"
#include <cstdlib>

namespace dgdm {

// ------------------------------------------------------------------------------------------ scans
// Exclusive scan of `in[0..n)` into `out[0..n]` (out[n] = total) by ONE block; n up to a few million.
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int* __restrict__ in, int* __restrict__ out, int n) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n ? in[i] : 0;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
    if (lane == 31) warp_tot[warp] = s;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const int carry = carry_s;
    const int excl = carry + (warp ? warp_tot[warp - 1] : 0) + s - v;
    if (i < n) out[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_tot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry_s;
}

// ------------------------------------------------------------------------------------------ CSR build
// Edges whose source or destination lies outside [0, N) are left out of the CSR (both passes apply the same test), so a forward
// whose input validation is read back BEHIND it (captured graphs, loader.GraphServer) cannot index out of bounds; the validation
// kernel flags such inputs and the host raises the reference's error for them.
__device__ __forceinline__ bool edge_in_range(long long s, long long d, int N) {
  return (unsigned long long)s < (unsigned long long)N && (unsigned long long)d < (unsigned long long)N;
}
__global__ void count_dst_kernel(const long long* __restrict__ ei, int E, int N, int* __restrict__ cnt) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const long long s = ei[e], d = ei[(size_t)E + e];
    if (edge_in_range(s, d, N)) atomicAdd(cnt + (int)d, 1);
  }
}
__global__ void fill_unordered_kernel(const long long* __restrict__ ei, int E, int N, const int* __restrict__ rowptr,
                                      int* __restrict__ cursor, int* __restrict__ tmp_eid) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const long long s = ei[e], dl = ei[(size_t)E + e];
    if (!edge_in_range(s, dl, N)) continue;
    const int d = (int)dl;
    tmp_eid[rowptr[d] + atomicAdd(cursor + d, 1)] = e;
  }
}
__global__ void dinv_kernel(const int* __restrict__ rowptr, int N, float* __restrict__ dinv) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
    dinv[i] = 1.0f / sqrtf((float)(rowptr[i + 1] - rowptr[i] + 1));   // deg incl. self loop; deg.pow(-0.5)
}
// warp per row: rank-sort the (unique) edge ids of the row ascending => original edge order; emit col / norm.
__global__ void __launch_bounds__(256) order_rows_kernel(const long long* __restrict__ ei, int E, const int* __restrict__ rowptr,
                                                         const int* __restrict__ tmp_eid, const float* __restrict__ dinv, int N,
                                                         int* __restrict__ col, int* __restrict__ eid, float* __restrict__ norm) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= N) return;
  const int beg = rowptr[row], end = rowptr[row + 1];
  const float dr = dinv[row];
  for (int i = beg + lane; i < end; i += 32) {
    const int mine = tmp_eid[i];
    int rank = 0;
    for (int j = beg; j < end; ++j) rank += (tmp_eid[j] < mine);
    const int src = (int)ei[mine];
    eid[beg + rank] = mine;
    col[beg + rank] = src;
    norm[beg + rank] = dinv[src] * dr;      // deg_inv_sqrt[row] * deg_inv_sqrt[col], core/graph_layers.py:84
  }
}

// Multi-block exclusive scan (n beyond a few thousand): per-block sums -> one-block scan of the sums -> per-block rescan + offset.
constexpr int SCAN_TILE = 2048;       // elements per block (1024 threads x 2)
__device__ __forceinline__ int block_exclusive_scan_1024(int v, int* warp_tot /*[32] smem*/, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int s = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
  if (lane == 31) warp_tot[warp] = s;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    warp_tot[lane] = w;
  }
  __syncthreads();
  total = warp_tot[31];
  const int excl = (warp ? warp_tot[warp - 1] : 0) + s - v;
  __syncthreads();
  return excl;
}
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(const int* __restrict__ in, int n, int* __restrict__ sums) {
  __shared__ int wt[32];
  const int i0 = blockIdx.x * SCAN_TILE + 2 * threadIdx.x;
  int v = (i0 < n ? in[i0] : 0) + (i0 + 1 < n ? in[i0 + 1] : 0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) wt[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    int w = wt[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (threadIdx.x == 0) sums[blockIdx.x] = w;
  }
}
__global__ void __launch_bounds__(1024) scan_apply_kernel(const int* __restrict__ in, int n, const int* __restrict__ offs /*exclusive, [nb+1]*/,
                                                          int* __restrict__ out) {
  __shared__ int wt[32];
  const int i0 = blockIdx.x * SCAN_TILE + 2 * threadIdx.x;
  const int a = i0 < n ? in[i0] : 0, b = i0 + 1 < n ? in[i0 + 1] : 0;
  int total;
  const int excl = block_exclusive_scan_1024(a + b, wt, total) + offs[blockIdx.x];
  if (i0 < n) out[i0] = excl;
  if (i0 + 1 < n) out[i0 + 1] = excl + a;
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = offs[gridDim.x];
}
// out[0..n] = exclusive scan of in[0..n); scratch: >= n / SCAN_TILE + 2 ints (NULL or too small -> single-block scan)
static void launch_exclusive_scan(const int* in, int* out, int n, int* scratch, int scratch_ints, cudaStream_t s) {
  const int nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (nb <= 2 || scratch == nullptr || scratch_ints < 2 * (nb + 1)) { exclusive_scan_kernel<<<1, 1024, 0, s>>>(in, out, n); return; }
  int* sums = scratch; int* offs = scratch + (nb + 1);
  scan_block_sums_kernel<<<nb, 1024, 0, s>>>(in, n, sums);
  exclusive_scan_kernel<<<1, 1024, 0, s>>>(sums, offs, nb);
  scan_apply_kernel<<<nb, 1024, 0, s>>>(in, n, offs, out);
}

// ------------------------------------------------------------------------------------------ SpMM
template <int VPL, int U>
__global__ void __launch_bounds__(256) spmm_gcn_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                       const float* __restrict__ norm, const float* __restrict__ dinv,
                                                       int n_csr_rows, int n_rows, int h_row0, const float* __restrict__ H, int ldh, int C,
                                                       const float* __restrict__ addend, int ld_add,
                                                       const float* __restrict__ bias, int act, float* __restrict__ out, int ldo) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  float4 acc[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  float selfw = 1.0f;
  if (row < n_csr_rows) {
    const int beg = rowptr[row], end = rowptr[row + 1];
    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      int c = 0; float w = 0.f;
      if (lane < cnt) { c = __ldg(col + base + lane); w = __ldg(norm + base + lane); }
      for (int j0 = 0; j0 < cnt; j0 += U) {
        float4 h[U][VPL]; float wj[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {               // issue all loads of U neighbours first (memory-level parallelism)
          const int j = j0 + u;
          const int cj = __shfl_sync(0xffffffffu, c, j & 31);
          wj[u] = (j < cnt) ? __shfl_sync(0xffffffffu, w, j & 31) : 0.f;
          const float* hp = H + (size_t)cj * ldh;
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            const int idx = (v * 32 + lane) * 4;
            h[u][v] = (j < cnt && idx < C) ? ldg4(hp + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)                 // accumulate in edge order (fixed, deterministic)
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            acc[v].x = fmaf(wj[u], h[u][v].x, acc[v].x); acc[v].y = fmaf(wj[u], h[u][v].y, acc[v].y);
            acc[v].z = fmaf(wj[u], h[u][v].z, acc[v].z); acc[v].w = fmaf(wj[u], h[u][v].w, acc[v].w);
          }
      }
    }
    const float d = dinv[row];
    selfw = d * d;
  }
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int idx = (v * 32 + lane) * 4;
    if (idx < C) {
      const float4 hs = ldg4(H + (size_t)(h_row0 + row) * ldh + idx);  // self loop last (add_self_loops appends)
      float4 a = acc[v];
      a.x = fmaf(selfw, hs.x, a.x); a.y = fmaf(selfw, hs.y, a.y); a.z = fmaf(selfw, hs.z, a.z); a.w = fmaf(selfw, hs.w, a.w);
      if (addend && row < n_csr_rows) { const float4 t = ldg4(addend + (size_t)row * ld_add + idx); a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
      if (bias) { const float4 t = ldg4(bias + idx); a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
      apply_act4(a, act);
      *reinterpret_cast<float4*>(out + (size_t)row * ldo + idx) = a;
    }
  }
}

// agg[i,:] = sum_{e in row i} norm[e] * edge_attr[eid[e], :]   (D <= 128, D % 4 == 0; lane handles one float4)
__global__ void __launch_bounds__(256) edge_attr_aggregate_kernel(const int* __restrict__ rowptr, const int* __restrict__ eid,
                                                                  const float* __restrict__ norm, int N,
                                                                  const float* __restrict__ ea, int ld, int D, float* __restrict__ agg) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= N) return;
  const int idx = lane * 4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = rowptr[row]; e < rowptr[row + 1]; ++e) {
    const float w = __ldg(norm + e);
    if (idx < D) {
      const float4 t = ldg4(ea + (size_t)__ldg(eid + e) * ld + idx);
      a.x = fmaf(w, t.x, a.x); a.y = fmaf(w, t.y, a.y); a.z = fmaf(w, t.z, a.z); a.w = fmaf(w, t.w, a.w);
    }
  }
  if (idx < D) *reinterpret_cast<float4*>(agg + (size_t)row * D + idx) = a;
}

// ------------------------------------------------------------------------------------------ top-k select
// state[0] = prefix (ordered-uint key bits decided so far), state[1] = remaining k, state[2] = pass
__global__ void __launch_bounds__(256) radix_hist_kernel(const float* __restrict__ s, int N, const unsigned* __restrict__ state,
                                                         unsigned* __restrict__ hist) {
  __shared__ unsigned sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const unsigned pass = state[2], prefix = state[0];
  const int shift = 24 - 8 * (int)pass;
  const unsigned mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const unsigned k = float_to_ordered(s[i]);
    if ((k & mask) == (prefix & mask)) atomicAdd(&sh[(k >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
__global__ void __launch_bounds__(256) radix_pick_kernel(unsigned* __restrict__ state, unsigned* __restrict__ hist) {
  // digit d* = the largest d with suffix[d] = sum_{d' >= d} hist[d'] >= k (k-th largest lies in it); 256 threads, reverse inclusive scan
  __shared__ unsigned suf[256];
  const int d = threadIdx.x;
  suf[d] = hist[255 - d];                       // reversed: index r = 255 - d
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    const unsigned t = d >= o ? suf[d - o] : 0u;
    __syncthreads();
    suf[d] += t;
    __syncthreads();
  }
  const unsigned k = state[1];
  // suf[r] = sum of hist over digits >= 255 - r; the wanted digit is the smallest r with suf[r] >= k
  const bool here = suf[d] >= k && (d == 0 || suf[d - 1] < k);
  __syncthreads();
  if (here) {
    const int shift = 24 - 8 * (int)state[2];
    state[0] |= ((unsigned)(255 - d) << shift);
    state[1] = k - (d == 0 ? 0u : suf[d - 1]);  // how many elements with this digit (finally: equal to the threshold) to take
    state[2] += 1;
  }
  hist[threadIdx.x] = 0;
}
// Ordered compaction, multi-block. selected = key > T, or key == T for the first state[1] such nodes (lowest index first).
// (a) per-block counts of (key > T, key == T); (b) one block turns them into exclusive offsets of the ties and of the selected
// nodes; (c) every block re-derives its local prefixes and writes perm / node_map.
constexpr int TOPK_TILE = 1024;
__global__ void __launch_bounds__(1024) topk_count_kernel(const float* __restrict__ s, int N, const unsigned* __restrict__ state,
                                                          int* __restrict__ blk_gt, int* __restrict__ blk_eq) {
  __shared__ int wg[32], we[32];
  const unsigned T = state[0];
  const int i = blockIdx.x * TOPK_TILE + threadIdx.x;
  unsigned key = 0; const bool valid = i < N;
  if (valid) key = float_to_ordered(s[i]);
  int g = valid && key > T, e = valid && key == T;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { g += __shfl_xor_sync(0xffffffffu, g, o); e += __shfl_xor_sync(0xffffffffu, e, o); }
  if ((threadIdx.x & 31) == 0) { wg[threadIdx.x >> 5] = g; we[threadIdx.x >> 5] = e; }
  __syncthreads();
  if (threadIdx.x < 32) {
    g = wg[threadIdx.x]; e = we[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { g += __shfl_xor_sync(0xffffffffu, g, o); e += __shfl_xor_sync(0xffffffffu, e, o); }
    if (threadIdx.x == 0) { blk_gt[blockIdx.x] = g; blk_eq[blockIdx.x] = e; }
  }
}
__global__ void __launch_bounds__(1024) topk_offsets_kernel(const int* __restrict__ blk_gt, const int* __restrict__ blk_eq, int nb,
                                                            const unsigned* __restrict__ state, int* __restrict__ eq_off,
                                                            int* __restrict__ sel_off) {
  __shared__ int wt[32];
  __shared__ int carry_eq, carry_sel;
  if (threadIdx.x == 0) { carry_eq = 0; carry_sel = 0; }
  __syncthreads();
  const int need_eq = (int)state[1];
  for (int base = 0; base < nb; base += 1024) {
    const int b = base + threadIdx.x;
    const int e = b < nb ? blk_eq[b] : 0;
    int tot_e;
    const int eo = block_exclusive_scan_1024(e, wt, tot_e) + carry_eq;
    int take = need_eq - eo; take = take < 0 ? 0 : (take > e ? e : take);
    const int sel = b < nb ? blk_gt[b] + take : 0;
    int tot_s;
    const int so = block_exclusive_scan_1024(sel, wt, tot_s) + carry_sel;
    if (b < nb) { eq_off[b] = eo; sel_off[b] = so; }
    __syncthreads();
    if (threadIdx.x == 0) { carry_eq += tot_e; carry_sel += tot_s; }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(1024) topk_scatter_kernel(const float* __restrict__ s, int N, const unsigned* __restrict__ state,
                                                            const int* __restrict__ eq_off, const int* __restrict__ sel_off,
                                                            int* __restrict__ perm, int* __restrict__ node_map) {
  __shared__ int wt[32];
  const unsigned T = state[0];
  const int need_eq = (int)state[1];
  const int i = blockIdx.x * TOPK_TILE + threadIdx.x;
  unsigned key = 0; const bool valid = i < N;
  if (valid) key = float_to_ordered(s[i]);
  const int eq = valid && key == T;
  int tot;
  const int eq_before = eq_off[blockIdx.x] + block_exclusive_scan_1024(eq, wt, tot);
  const int sel = valid && (key > T || (eq && eq_before < need_eq));
  const int pos = sel_off[blockIdx.x] + block_exclusive_scan_1024(sel, wt, tot);
  if (valid) {
    node_map[i] = sel ? pos : -1;
    if (sel) perm[pos] = i;
  }
}
__global__ void init_select_state_kernel(unsigned* state, unsigned* hist, int k) {
  if (threadIdx.x == 0) { state[0] = 0; state[1] = (unsigned)k; state[2] = 0; }
  hist[threadIdx.x] = 0;
}

// pooled features: out[r,:] = act(X[perm[r],:] * score[perm[r]])   (multiplier = 1.0, core/graph_layers.py:316)
__global__ void __launch_bounds__(256) pool_gather_kernel(const float* __restrict__ X, int ldx, int C, const float* __restrict__ s,
                                                          const int* __restrict__ perm, int k, int act, float* __restrict__ out, int ldo) {
  const int c4 = C / 4;
  const size_t total = (size_t)k * c4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / c4), c = (int)(i - (size_t)r * c4) * 4;
    const int src = perm[r];
    const float sc = s[src];
    float4 v = ldg4(X + (size_t)src * ldx + c);
    v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    apply_act4(v, act);
    *reinterpret_cast<float4*>(out + (size_t)r * ldo + c) = v;
  }
}

// pooled CSR: new row r = old row perm[r]; keep in-edges whose source survives, in order.
__global__ void __launch_bounds__(256) csr_pool_count_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                             const int* __restrict__ perm, const int* __restrict__ node_map, int k,
                                                             int* __restrict__ cnt, float* __restrict__ p_dinv) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= k) return;
  const int old = perm[r];
  int c = 0;
  for (int e = rowptr[old] + lane; e < rowptr[old + 1]; e += 32) c += (node_map[col[e]] >= 0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) { cnt[r] = c; p_dinv[r] = 1.0f / sqrtf((float)(c + 1)); }
}
__global__ void __launch_bounds__(256) csr_pool_fill_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                            const int* __restrict__ perm, const int* __restrict__ node_map, int k,
                                                            const int* __restrict__ p_rowptr, const float* __restrict__ p_dinv,
                                                            int* __restrict__ p_col, float* __restrict__ p_norm,
                                                            const int* __restrict__ eid, int* __restrict__ p_eid) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= k) return;
  const int old = perm[r];
  const int beg = rowptr[old], end = rowptr[old + 1];
  int outp = p_rowptr[r];
  const float dr = p_dinv[r];
  for (int base = beg; base < end; base += 32) {
    const int e = base + lane;
    int nm = -1;
    if (e < end) nm = node_map[col[e]];
    const unsigned b = __ballot_sync(0xffffffffu, nm >= 0);
    if (nm >= 0) {
      const int pos = outp + __popc(b & ((1u << lane) - 1u));
      p_col[pos] = nm;
      p_norm[pos] = p_dinv[nm] * dr;
      if (p_eid) p_eid[pos] = eid[e];         // level-0 edge id: edge_attr[edge_mask] (core/graph_layers.py:327) without a copy
    }
    outp += __popc(b);
  }
}

// ------------------------------------------------------------------------------------------ segment softmax
__global__ void __launch_bounds__(256) segment_softmax_scores_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                                     const int* __restrict__ eid, int N, const float* __restrict__ qkv,
                                                                     int ld_qkv, const float* __restrict__ ek, int ld_ek, int hid,
                                                                     int heads, float* __restrict__ out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= N) return;
  const int beg = rowptr[row], end = rowptr[row + 1];
  if (beg == end) return;
  const int hd = hid / heads;
  const float scale = 1.0f / sqrtf((float)hd);
  const float* krow = qkv + (size_t)row * ld_qkv + hid;            // k[dst]
  for (int h = 0; h < heads; ++h) {
    float mx = -INFINITY;
    for (int i = beg + lane; i < end; i += 32) {
      const float* q = qkv + (size_t)col[i] * ld_qkv + h * hd;     // q[src]
      const float* kk = krow + h * hd;
      const float* e = ek + (size_t)eid[i] * ld_ek + h * hd;
      float s = 0.f;
      for (int d = 0; d < hd; ++d) s = fmaf(q[d], kk[d] + e[d], s);
      s *= scale;
      out[(size_t)eid[i] * heads + h] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int i = beg + lane; i < end; i += 32) {
      const float ex = expf(out[(size_t)eid[i] * heads + h] - mx);
      out[(size_t)eid[i] * heads + h] = ex;
      sum += ex;
    }
    sum = warp_sum(sum) + 1e-16f;
    for (int i = beg + lane; i < end; i += 32) out[(size_t)eid[i] * heads + h] /= sum;
  }
}

static inline int grid1d(size_t n, int block = 256, int cap = 148 * 16) {
  size_t b = (n + block - 1) / block;
  return (int)(b < 1 ? 1 : (b > (size_t)cap ? cap : b));
}

}  // namespace dgdm
using namespace dgdm;

extern "C" size_t dgdm_csr_build_work_ints(int N, int E) { return (size_t)2 * (N + 1) + (size_t)E + 16; }

extern "C" int dgdm_csr_build(const int64_t* edge_index, int E, int N, int* rowptr, int* col, int* eid, float* norm,
                              float* dinv, int* work, void* stream) {
  if (!edge_index && E > 0) return DGDM_ERR_INVALID;
  if (!rowptr || !dinv || !work || N <= 0 || E < 0) return DGDM_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  int* cnt = work; int* cursor = work + (N + 1); int* tmp_eid = work + 2 * (N + 1);
  if (cudaMemsetAsync(cnt, 0, sizeof(int) * 2 * (N + 1), s) != cudaSuccess) return DGDM_ERR_LAUNCH;
  const long long* ei = (const long long*)edge_index;
  if (E > 0) count_dst_kernel<<<grid1d(E), 256, 0, s>>>(ei, E, N, cnt);
  launch_exclusive_scan(cnt, rowptr, N, tmp_eid, E + 16, s);      // tmp_eid is free until fill_unordered_kernel
  dinv_kernel<<<grid1d(N), 256, 0, s>>>(rowptr, N, dinv);
  if (E > 0) {
    fill_unordered_kernel<<<grid1d(E), 256, 0, s>>>(ei, E, N, rowptr, cursor, tmp_eid);
    order_rows_kernel<<<(N + 7) / 8, 256, 0, s>>>(ei, E, rowptr, tmp_eid, dinv, N, col, eid, norm);
  }
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" int dgdm_spmm_gcn(const int* rowptr, const int* col, const float* norm, const float* dinv, int n_csr_rows,
                             int n_rows, int h_row0, const float* H, int ldh, int C, const float* addend, int ld_addend,
                             const float* bias, int act, float* out, int ldo, void* stream) {
  if (!rowptr || !dinv || !H || !out || C <= 0 || (C % 4) != 0 || C > 1024 || (ldh % 4) != 0 || (ldo % 4) != 0) return DGDM_ERR_INVALID;
  if (addend && (ld_addend % 4) != 0) return DGDM_ERR_INVALID;
  if (n_rows <= 0) return DGDM_OK;
  if (n_csr_rows > n_rows || h_row0 < 0) return DGDM_ERR_INVALID;
  const int blocks = (n_rows + 7) / 8;
  cudaStream_t s = (cudaStream_t)stream;
  const int vpl = (C + 127) / 128;
  static const int unroll = [] { const char* e = getenv("DGDM_SPMM_UNROLL"); return e ? atoi(e) : 0; }();
  // neighbours in flight per warp: measured per width (tools/spmm_bench.py); DGDM_SPMM_UNROLL overrides for A/B runs
  const int U = unroll ? unroll : 2;
#define LAUNCH_SPMM_U(V, UU) spmm_gcn_kernel<V, UU><<<blocks, 256, 0, s>>>(rowptr, col, norm, dinv, n_csr_rows, n_rows, h_row0, H, ldh, C, addend, ld_addend, bias, act, out, ldo)
#define LAUNCH_SPMM(V) do { if (U <= 1) LAUNCH_SPMM_U(V, 1); else if (U == 2) LAUNCH_SPMM_U(V, 2); else if (U == 3) LAUNCH_SPMM_U(V, 3); else LAUNCH_SPMM_U(V, 4); } while (0)
  switch (vpl) {
    case 1: LAUNCH_SPMM(1); break;
    case 2: LAUNCH_SPMM(2); break;
    case 3: LAUNCH_SPMM(3); break;
    case 4: LAUNCH_SPMM(4); break;
    case 5: case 6: LAUNCH_SPMM(6); break;
    default: LAUNCH_SPMM(8); break;
  }
#undef LAUNCH_SPMM
#undef LAUNCH_SPMM_U
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" int dgdm_edge_attr_aggregate(const int* rowptr, const int* eid, const float* norm, int N,
                                        const float* edge_attr, int ld_attr, int D, float* agg, void* stream) {
  if (!rowptr || !eid || !norm || !edge_attr || !agg || D <= 0 || D > 128 || (D % 4) != 0 || (ld_attr % 4) != 0) return DGDM_ERR_INVALID;
  if (N <= 0) return DGDM_OK;
  edge_attr_aggregate_kernel<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(rowptr, eid, norm, N, edge_attr, ld_attr, D, agg);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" size_t dgdm_topk_pool_work_ints(int N) { return (size_t)N + 512 + 4 * ((size_t)N / TOPK_TILE + 2); }

extern "C" int dgdm_topk_select(const float* scores, int N, int k, int* perm, int* node_map, int* work, void* stream) {
  if (!scores || !perm || !node_map || !work || N <= 0 || k <= 0 || k > N) return DGDM_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned* state = (unsigned*)work; unsigned* hist = state + 8;
  init_select_state_kernel<<<1, 256, 0, s>>>(state, hist, k);
  for (int pass = 0; pass < 4; ++pass) {
    radix_hist_kernel<<<grid1d(N, 256, 148 * 4), 256, 0, s>>>(scores, N, state, hist);
    radix_pick_kernel<<<1, 256, 0, s>>>(state, hist);
  }
  const int nb = (N + TOPK_TILE - 1) / TOPK_TILE;
  int* blk = work + 512;                 // [gt | eq | eq_off | sel_off], nb + 1 ints each (the state / histogram use the first 264 ints)
  int* blk_gt = blk; int* blk_eq = blk + (nb + 1); int* eq_off = blk + 2 * (nb + 1); int* sel_off = blk + 3 * (nb + 1);
  topk_count_kernel<<<nb, 1024, 0, s>>>(scores, N, state, blk_gt, blk_eq);
  topk_offsets_kernel<<<1, 1024, 0, s>>>(blk_gt, blk_eq, nb, state, eq_off, sel_off);
  topk_scatter_kernel<<<nb, 1024, 0, s>>>(scores, N, state, eq_off, sel_off, perm, node_map);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" int dgdm_pool_gather(const float* X, int ldx, int C, const float* scores, const int* perm, int k, int act,
                                float* out, int ldo, void* stream) {
  if (!X || !scores || !perm || !out || C <= 0 || (C % 4) != 0 || (ldx % 4) != 0 || (ldo % 4) != 0) return DGDM_ERR_INVALID;
  if (k <= 0) return DGDM_OK;
  pool_gather_kernel<<<grid1d((size_t)k * (C / 4)), 256, 0, (cudaStream_t)stream>>>(X, ldx, C, scores, perm, k, act, out, ldo);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" int dgdm_csr_pool(const int* rowptr, const int* col, const int* eid, int N, const int* perm, const int* node_map, int k,
                             int* p_rowptr, int* p_col, int* p_eid, float* p_norm, float* p_dinv, int* work, void* stream) {
  if (!rowptr || !perm || !node_map || !p_rowptr || !p_dinv || !work || k <= 0 || k > N) return DGDM_ERR_INVALID;
  if (p_eid && !eid) return DGDM_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  int* cnt = work;
  csr_pool_count_kernel<<<(k + 7) / 8, 256, 0, s>>>(rowptr, col, perm, node_map, k, cnt, p_dinv);
  launch_exclusive_scan(cnt, p_rowptr, k, work + k + 8, (k / SCAN_TILE + 2) * 2 + 8, s);
  csr_pool_fill_kernel<<<(k + 7) / 8, 256, 0, s>>>(rowptr, col, perm, node_map, k, p_rowptr, p_dinv, p_col, p_norm, eid, p_eid);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}

extern "C" int dgdm_segment_softmax_scores(const int* rowptr, const int* col, const int* eid, int N, const float* qkv,
                                           int ld_qkv, const float* ek, int ld_ek, int hid, int heads, float* out, void* stream) {
  if (!rowptr || !col || !eid || !qkv || !ek || !out || hid <= 0 || heads <= 0 || (hid % heads) != 0) return DGDM_ERR_INVALID;
  if (N <= 0) return DGDM_OK;
  segment_softmax_scores_kernel<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(rowptr, col, eid, N, qkv, ld_qkv, ek, ld_ek, hid, heads, out);
  DGDM_CHECK_LAUNCH();
  return DGDM_OK;
}
