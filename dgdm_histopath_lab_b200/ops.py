"""Thin Python wrappers over the C-ABI (include/dgdm_b200.h). Torch supplies device memory and the current
stream only; every computation here is a hand-written sm_100a kernel. No fallback paths."""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import GemmParams

ACT = {None: 0, "none": 0, "gelu": 1, "relu": 2, "silu": 3, "tanh": 4, "elu": 5}


class KernelError(RuntimeError):
    pass


# kernels launched per C-ABI call (for bench.py's gpu_launches claim; memsets not counted)
KERNELS_PER_CALL = {
    "dgdm_gemm_f16x2": 1, "dgdm_gemm_simt": 2, "dgdm_split_weight_f16": 1, "dgdm_rowwise_epilogue": 1, "dgdm_csr_build": 7,
    "dgdm_spmm_gcn": 1, "dgdm_edge_attr_aggregate": 1, "dgdm_topk_select": 12, "dgdm_pool_gather": 1, "dgdm_csr_pool": 5,
    "dgdm_unpool_add_act": 1, "dgdm_segment_softmax_scores": 1, "dgdm_segment_minmax": 1, "dgdm_spatial_pe_add": 1,
    "dgdm_spatial_attention": 6, "dgdm_spatial_attention_weights": 1, "dgdm_attention_pool": 2, "dgdm_diffusion_add_noise": 1,
    "dgdm_groupnorm_silu": 1, "dgdm_ddpm_update": 1, "dgdm_spatial_attention_backward": 3, "dgdm_spatial_attention_backward_tc": 3, "dgdm_mha_dense": 1,
    "dgdm_row_normalize": 1, "dgdm_contrastive_accumulate": 1, "dgdm_validate_inputs": 1, "dgdm_knn_cosine": 4, "dgdm_knn2d": 2,
    "dgdm_row_sqnorm_f64": 1, "dgdm_knn_cosine_tc": 6,
}
LAUNCHES = {"kernels": 0, "calls": 0}
# MUFU (XU pipe) operations per attention score in spatial_attn_tc2_kernel (and its predecessor): one exp2 per (row, key, head) + the distance sqrt of a
# (row, key) shared by the heads of a softmax group (bench.py's roofline denominator; see attention_tc.cu)
ATTN_MUFU_PER_SCORE = 1.25
_PROFILE = None      # when set to a list: (name, start_event, end_event) per C call (bench.py's per-kernel timing pass)


def _check(rc, name):
    if rc != 0:
        raise KernelError(f"{name} failed with code {rc}")
    LAUNCHES["kernels"] += KERNELS_PER_CALL.get(name, 1)
    LAUNCHES["calls"] += 1
    if _PROFILE is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        _PROFILE.append((name, ev))


_PROFILE_META = {}


def take_profile_meta():
    """Shapes recorded during the last profile pass: {'spmm': [(rows, edges, C)], 'gemm': [(M, K, N)]}."""
    global _PROFILE_META
    m, _PROFILE_META = _PROFILE_META, {}
    return m


def profile_begin():
    """Start recording one CUDA event after every C-ABI call on the current stream (durations = gaps between events)."""
    global _PROFILE, _PROFILE_META
    _PROFILE_META = {}
    _PROFILE = []
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    _PROFILE.append(("__start__", ev))


def profile_end():
    global _PROFILE
    rec, _PROFILE = _PROFILE, None
    torch.cuda.synchronize()
    out = {}
    for (n0, e0), (n1, e1) in zip(rec[:-1], rec[1:]):
        d = out.setdefault(n1, [0, 0.0])
        d[0] += 1
        d[1] += e0.elapsed_time(e1)
    return out


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t):
    assert t.is_cuda and t.dtype == torch.float32, (t.device, t.dtype)
    return t


def device_check(dev=0):
    _check(_lib.load().dgdm_device_check(int(dev)), "dgdm_device_check")


# ------------------------------------------------------------------------------------------------ weights
class SplitWeight:
    """fp16 hi/lo planes of an fp32 [N, K] weight as pre-swizzled K-block images (dgdm_split_weight_f16:
    img[kb][plane][N rows x 128 B], SWIZZLE_128B canonical layout) + the fp32 original (for the SIMT comparator)."""
    __slots__ = ("img", "N", "K", "Kpad", "w32")

    def __init__(self, w):
        w = _f32c(w.detach()).contiguous()
        self.N, self.K = w.shape
        self.Kpad = (self.K + 63) // 64 * 64
        self.img = torch.empty(self.N * self.Kpad * 4, dtype=torch.uint8, device=w.device)
        self.w32 = w
        _check(_lib.load().dgdm_split_weight_f16(_p(w), self.N, self.K, self.hi_ptr, self.lo_ptr, self.Kpad, _stream()),
               "dgdm_split_weight_f16")

    @property
    def hi_ptr(self):
        return C.c_void_p(self.img.data_ptr())

    @property
    def lo_ptr(self):
        return C.c_void_p(self.img.data_ptr() + self.N * 128)


def gemm_mode():
    return os.environ.get("DGDM_GEMM", "tc")


def linear(a, w, bias=None, *, r_pre=None, act_pre=None, ln1=None, ln2=None, act_post=None, r_post=None,
           dot=None, out=None, want_out=True, ln_eps=1e-5, row_bias=None, a_scale=1.0, mode=None, peers=None, multicast=False):
    """out = epilogue(a @ w^T). a [M,K] fp32 (row stride may exceed K), w SplitWeight [N,K].

    ln1/ln2: (gamma, beta) tuples. dot: (w_vec [N], b float, act) -> also returns dot_out [M].
    row_bias: (index int32 [M], table [R,N]). Returns out, or (out, dot_out) when dot is given.
    peers: addresses (ints) of the peers' copies of `out` (symmetric memory) — the epilogue stores every tile there as well
    (fused halo all-gather); multicast=True: peers[0] is an NVLS multicast address covering all ranks incl. this one.
    """
    lib = _lib.load()
    mode = mode or gemm_mode()
    assert a.dim() == 2 and a.stride(1) == 1
    M, K = a.shape
    assert K == w.K, (K, w.K)
    N = w.N
    p = GemmParams()
    p.A, p.lda, p.M, p.K = _p(_f32c(a)), a.stride(0), M, K
    p.W_hi, p.W_lo, p.ldw, p.N = w.hi_ptr, w.lo_ptr, w.Kpad, N
    if want_out or mode == "simt" or N > 512:
        if out is None:
            out = torch.empty(M, N, dtype=torch.float32, device=a.device)
        assert out.stride(1) == 1
        p.C, p.ldc = _p(out), out.stride(0)
    else:
        out = None
        p.C, p.ldc = None, 0
    p.bias = _p(bias)
    if r_pre is not None:
        assert r_pre.stride(1) == 1 and r_pre.shape == (M, N)
        p.r_pre, p.ld_r_pre = _p(r_pre), r_pre.stride(0)
    if r_post is not None:
        assert r_post.stride(1) == 1 and r_post.shape == (M, N)
        p.r_post, p.ld_r_post = _p(r_post), r_post.stride(0)
    p.act_pre, p.act_post = ACT[act_pre], ACT[act_post]
    if ln1 is not None:
        p.ln1_g, p.ln1_b = _p(ln1[0]), _p(ln1[1])
    if ln2 is not None:
        p.ln2_g, p.ln2_b = _p(ln2[0]), _p(ln2[1])
    p.ln_eps = ln_eps
    dot_out = None
    if dot is not None:
        dot_out = torch.empty(M, dtype=torch.float32, device=a.device)
        p.dot_w, p.dot_b, p.dot_act, p.dot_out = _p(dot[0]), float(dot[1]), ACT[dot[2]], _p(dot_out)
    p.a_scale = float(a_scale)
    if row_bias is not None:
        p.row_bias_index, p.row_bias, p.ld_row_bias = _p(row_bias[0]), _p(row_bias[1]), row_bias[1].stride(0)
    if peers:
        assert out is not None and mode != "simt" and N <= 512 and len(peers) <= 8
        for i, a_ in enumerate(peers):
            p.C_peers[i] = int(a_)
        p.n_peers, p.peers_multicast = len(peers), int(bool(multicast))
    if _PROFILE is not None:
        _PROFILE_META.setdefault("gemm", []).append((M, K, N))
    if M > 0:
        if mode == "simt":
            _check(lib.dgdm_gemm_simt(C.byref(p), _p(w.w32), _stream()), "dgdm_gemm_simt")
        else:
            if N % 16 != 0:
                raise KernelError(f"linear: output width {N} must be a multiple of 16")
            # N > 512 (DGDM-Large 768-d rows do not fit TMEM): the C side runs column groups + the stand-alone epilogue
            _check(lib.dgdm_gemm_f16x2(C.byref(p), _stream()), "dgdm_gemm_f16x2")
    return (out, dot_out) if dot is not None else out


# ------------------------------------------------------------------------------------------------ graph
class CSR:
    """CSR by destination, original edge order inside rows, GCN norm (core/graph_layers.py:76-84)."""
    __slots__ = ("rowptr", "col", "eid", "norm", "dinv", "N", "E", "n_rows_total")

    def __init__(self, rowptr, col, eid, norm, dinv, N, E):
        self.rowptr, self.col, self.eid, self.norm, self.dinv, self.N, self.E = rowptr, col, eid, norm, dinv, N, E


def csr_build(edge_index, N):
    lib = _lib.load()
    assert edge_index.is_cuda and edge_index.dtype == torch.int64 and edge_index.dim() == 2 and edge_index.size(0) == 2
    ei = edge_index.contiguous()
    E = ei.size(1)
    dev = ei.device
    rowptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    col = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
    eid = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
    norm = torch.empty(max(E, 1), dtype=torch.float32, device=dev)
    dinv = torch.empty(N, dtype=torch.float32, device=dev)
    work = torch.empty(int(lib.dgdm_csr_build_work_ints(N, E)), dtype=torch.int32, device=dev)
    _check(lib.dgdm_csr_build(_p(ei), E, N, _p(rowptr), _p(col), _p(eid), _p(norm), _p(dinv), _p(work), _stream()),
           "dgdm_csr_build")
    return CSR(rowptr, col, eid, norm, dinv, N, E)


def spmm_gcn(csr, h, *, n_rows=None, rows=None, addend=None, bias=None, act=None, out=None):
    """GCN propagate. Single GPU: rows [0, n_rows) of `h`. Node-partitioned: rows=(a, b) are the output rows this rank owns,
    `h` is the all-gathered feature matrix (global row indices), out/addend are local [b-a, C]."""
    lib = _lib.load()
    Cw = h.size(1)
    if rows is None:
        a, n_rows = 0, (csr.N if n_rows is None else n_rows)
    else:
        a, n_rows = int(rows[0]), int(rows[1]) - int(rows[0])
    n_csr = max(0, min(csr.N - a, n_rows))
    assert h.stride(1) == 1 and h.size(0) >= a + n_rows
    if out is None:
        out = torch.empty(n_rows, Cw, dtype=torch.float32, device=h.device)
    if _PROFILE is not None:      # profiling pass only: the edge count of pooled levels lives on the device
        e0 = int(csr.rowptr[min(csr.N, a)].item()); e1 = int(csr.rowptr[min(csr.N, a + n_rows)].item())
        _PROFILE_META.setdefault("spmm", []).append((n_rows, e1 - e0, Cw))
    a_csr = min(a, csr.N)          # rows beyond the CSR (D8) never dereference rowptr/dinv
    _check(lib.dgdm_spmm_gcn(C.c_void_p(csr.rowptr.data_ptr() + 4 * a_csr), _p(csr.col), _p(csr.norm),
                             C.c_void_p(csr.dinv.data_ptr() + 4 * min(a_csr, max(csr.N - 1, 0))), n_csr, n_rows, a,
                             _p(_f32c(h)), h.stride(0), Cw, _p(addend), addend.stride(0) if addend is not None else 0,
                             _p(bias), ACT[act], _p(out), out.stride(0), _stream()), "dgdm_spmm_gcn")
    return out


def edge_attr_aggregate(csr, edge_attr):
    lib = _lib.load()
    D = edge_attr.size(1)
    agg = torch.empty(csr.N, D, dtype=torch.float32, device=edge_attr.device)
    _check(lib.dgdm_edge_attr_aggregate(_p(csr.rowptr), _p(csr.eid), _p(csr.norm), csr.N, _p(_f32c(edge_attr)),
                                        edge_attr.stride(0), D, _p(agg), _stream()), "dgdm_edge_attr_aggregate")
    return agg


def topk_select(scores, k):
    lib = _lib.load()
    N = scores.numel()
    dev = scores.device
    perm = torch.empty(k, dtype=torch.int32, device=dev)
    node_map = torch.empty(N, dtype=torch.int32, device=dev)
    work = torch.empty(int(lib.dgdm_topk_pool_work_ints(N)), dtype=torch.int32, device=dev)
    _check(lib.dgdm_topk_select(_p(_f32c(scores)), N, k, _p(perm), _p(node_map), _p(work), _stream()), "dgdm_topk_select")
    return perm, node_map


def pool_gather(x, scores, perm, act=None):
    lib = _lib.load()
    k, Cw = perm.numel(), x.size(1)
    out = torch.empty(k, Cw, dtype=torch.float32, device=x.device)
    _check(lib.dgdm_pool_gather(_p(_f32c(x)), x.stride(0), Cw, _p(scores), _p(perm), k, ACT[act], _p(out), out.stride(0),
                                _stream()), "dgdm_pool_gather")
    return out


def csr_pool(csr, perm, node_map, with_eid=False):
    """with_eid: the pooled CSR carries the level-0 edge ids (for edge_attr aggregation on pooled levels)."""
    lib = _lib.load()
    k = perm.numel()
    dev = perm.device
    p_rowptr = torch.empty(k + 1, dtype=torch.int32, device=dev)
    p_col = torch.empty(max(csr.E, 1), dtype=torch.int32, device=dev)
    p_norm = torch.empty(max(csr.E, 1), dtype=torch.float32, device=dev)
    p_dinv = torch.empty(k, dtype=torch.float32, device=dev)
    work = torch.empty(k + 2 * (k // 2048 + 2) + 64, dtype=torch.int32, device=dev)      # counts + scan scratch
    with_eid = with_eid and csr.eid is not None
    p_eid = torch.empty(max(csr.E, 1), dtype=torch.int32, device=dev) if with_eid else None
    _check(lib.dgdm_csr_pool(_p(csr.rowptr), _p(csr.col), _p(csr.eid) if with_eid else None, csr.N, _p(perm), _p(node_map), k,
                             _p(p_rowptr), _p(p_col), _p(p_eid), _p(p_norm), _p(p_dinv), _p(work), _stream()), "dgdm_csr_pool")
    return CSR(p_rowptr, p_col, p_eid, p_norm, p_dinv, k, csr.E)   # E is an upper bound; the true count is rowptr[k]


def unpool_add_act(skip, x, node_map, act=None):
    lib = _lib.load()
    N, Cw = skip.shape
    out = torch.empty(N, Cw, dtype=torch.float32, device=skip.device)
    _check(lib.dgdm_unpool_add_act(_p(_f32c(skip)), skip.stride(0), _p(_f32c(x)), x.stride(0), _p(node_map), N, Cw,
                                   ACT[act], _p(out), out.stride(0), _stream()), "dgdm_unpool_add_act")
    return out


def segment_softmax_scores(csr, qkv, ek, hid, heads):
    lib = _lib.load()
    out = torch.zeros(csr.E, heads, dtype=torch.float32, device=qkv.device)
    _check(lib.dgdm_segment_softmax_scores(_p(csr.rowptr), _p(csr.col), _p(csr.eid), csr.N, _p(_f32c(qkv)), qkv.stride(0),
                                           _p(_f32c(ek)), ek.stride(0), hid, heads, _p(out), _stream()),
           "dgdm_segment_softmax_scores")
    return out


# ------------------------------------------------------------------------------------------------ attention
def spatial_pe_add(x, pos, seg_ptr, num_seg, minmax_from=None):
    """minmax_from = (full_pos, full_seg_ptr): per-segment min/max taken over the whole batch while x / pos / seg_ptr describe
    only the rows this rank owns (node-partitioned graphs)."""
    lib = _lib.load()
    N, D = x.shape
    minmax = torch.empty(num_seg, 2, dtype=torch.float32, device=x.device)
    mm_pos, mm_seg = (pos, seg_ptr) if minmax_from is None else minmax_from
    _check(lib.dgdm_segment_minmax(_p(_f32c(mm_pos)), _p(mm_seg), num_seg, _p(minmax), _stream()), "dgdm_segment_minmax")
    out = torch.empty(N, D, dtype=torch.float32, device=x.device)
    _check(lib.dgdm_spatial_pe_add(_p(_f32c(x)), x.stride(0), _p(pos), _p(seg_ptr), num_seg, _p(minmax), N, D, _p(out),
                                   out.stride(0), _stream()), "dgdm_spatial_pe_add")
    return out


def attention_precise_default():
    return int(os.environ.get("DGDM_ATTN_PRECISE", "1"))


def spatial_attention(qkv, pos, seg_ptr, num_seg, heads, temperature=1.0, *, kv_qkv=None, kv_pos=None, kseg_ptr=None,
                      precise=None, max_seg=0, want_lse=False):
    """Flash-style attention core. qkv [N, 3D] (q|k|v). pos may be None (no spatial bias). Returns [N, D].
    max_seg: number of keys of the largest segment when known on the host (0 = unknown)."""
    lib = _lib.load()
    Nq, D3 = qkv.shape
    D = D3 // 3
    kv_qkv = qkv if kv_qkv is None else kv_qkv
    kv_pos = pos if kv_pos is None else kv_pos
    kseg_ptr = seg_ptr if kseg_ptr is None else kseg_ptr
    Nkv = kv_qkv.size(0)
    precise = attention_precise_default() if precise is None else int(precise)
    max_seg = Nkv if num_seg == 1 else int(max_seg)
    wb = int(lib.dgdm_spatial_attention_work_bytes(Nq, Nkv, D, heads, num_seg, max_seg))
    work = torch.empty(wb, dtype=torch.uint8, device=qkv.device)
    out = torch.empty(Nq, D, dtype=torch.float32, device=qkv.device)
    if want_lse:            # training path: keep the softmax statistics for the backward ([Nq, heads], natural-log units)
        lse = torch.empty(Nq, heads, dtype=torch.float32, device=qkv.device)
        _check(lib.dgdm_spatial_attention_lse(_p(_f32c(qkv)), qkv.stride(0), _p(pos), _p(seg_ptr), Nq, _p(_f32c(kv_qkv)),
                                              kv_qkv.stride(0), _p(kv_pos), _p(kseg_ptr), Nkv, num_seg, max_seg, D, heads,
                                              float(temperature), precise, _p(out), out.stride(0), _p(lse), _p(work), wb, _stream()),
               "dgdm_spatial_attention")
        return out, lse
    _check(lib.dgdm_spatial_attention(_p(_f32c(qkv)), qkv.stride(0), _p(pos), _p(seg_ptr), Nq, _p(_f32c(kv_qkv)),
                                      kv_qkv.stride(0), _p(kv_pos), _p(kseg_ptr), Nkv, num_seg, max_seg, D, heads,
                                      float(temperature), precise, _p(out), out.stride(0), _p(work), wb, _stream()),
           "dgdm_spatial_attention")
    return out


def spatial_attention_any(qkv, pos, seg_ptr, num_seg, heads, temperature=1.0, *, kv_qkv=None, kv_pos=None, kseg_ptr=None, max_seg=0):
    """spatial_attention for any supported head width. head_dim 16 goes straight to the kernels. Narrower heads whose width divides
    16 (DGDM-Large: 16 heads x 8) run on the SAME tcgen05 kernel: every head is zero-padded to 16 dims (dot products unchanged),
    q is pre-scaled by sqrt(16/hd) so that the kernel's 1/sqrt(16) becomes 1/sqrt(hd), and the heads go through in groups of 8
    (one 128-wide launch per group) — ~5x faster than the exact CUDA-core kernel, same fp32-class split arithmetic."""
    N, D3 = qkv.shape
    D = D3 // 3
    hd = D // heads
    if hd == 16 or 16 % hd != 0 or heads % 8 != 0 or hd > 16:
        return spatial_attention(qkv, pos, seg_ptr, num_seg, heads, temperature, kv_qkv=kv_qkv, kv_pos=kv_pos, kseg_ptr=kseg_ptr, max_seg=max_seg)
    kvq = qkv if kv_qkv is None else kv_qkv
    out = torch.empty(N, D, dtype=torch.float32, device=qkv.device)
    qs = (16.0 / hd) ** 0.5

    def padded(src, col0, scale):             # [n, 8*hd] columns col0.. of src -> [n, 128] with each head padded to 16 dims
        n = src.size(0)
        t = src.new_zeros(n, 8, 16)
        t[:, :, :hd] = src[:, col0:col0 + 8 * hd].reshape(n, 8, hd)
        return (t * scale if scale != 1.0 else t).reshape(n, 128)

    for g in range(heads // 8):
        c0 = g * 8 * hd
        k_p, v_p = padded(kvq, D + c0, 1.0), padded(kvq, 2 * D + c0, 1.0)
        q_p = padded(qkv, c0, qs)
        if kv_qkv is None:
            o = spatial_attention(torch.cat([q_p, k_p, v_p], 1), pos, seg_ptr, num_seg, 8, temperature, max_seg=max_seg)
        else:
            qq = torch.cat([q_p, q_p.new_zeros(N, 256)], 1)          # only the q third of the query-side matrix is read
            kk = torch.cat([k_p.new_zeros(kvq.size(0), 128), k_p, v_p], 1)
            o = spatial_attention(qq, pos, seg_ptr, num_seg, 8, temperature, kv_qkv=kk, kv_pos=kv_pos, kseg_ptr=kseg_ptr, max_seg=max_seg)
        out[:, c0:c0 + 8 * hd] = o.reshape(N, 8, 16)[:, :, :hd].reshape(N, 8 * hd)
    return out


def attention_backward_mode():
    return os.environ.get("DGDM_ATTN_BWD", "tc")


def spatial_attention_backward(qkv, pos, seg_of, seg_ptr, heads, temperature, out, grad_out, mode=None, lse=None):
    """d(qkv) of the spatial self-attention core (training path). qkv [N, 3D], out / grad_out [N, D]; seg_of int32 [N].
    mode "tc" (default): mma.sync tensor-core passes, fp16 operands (~2e-3); "simt": exact fp32 CUDA-core passes."""
    lib = _lib.load()
    mode = mode or attention_backward_mode()
    N, D3 = qkv.shape
    D = D3 // 3
    dqkv = torch.empty(N, D3, dtype=torch.float32, device=qkv.device)
    have_lse = int(lse is not None)
    if lse is None:
        lse = torch.empty(N, heads, dtype=torch.float32, device=qkv.device)
    delta = torch.empty(N, heads, dtype=torch.float32, device=qkv.device)
    go = _f32c(grad_out.contiguous())
    if mode == "tc":
        sc = torch.exp2(torch.round(torch.log2(8.0 / go.abs().amax().clamp_min(1e-30)))).reshape(1).contiguous()
        _check(lib.dgdm_spatial_attention_backward_tc(_p(_f32c(qkv)), qkv.stride(0), _p(pos), _p(seg_of), _p(seg_ptr), N, D, heads,
                                                      float(temperature), _p(_f32c(out)), out.stride(0), _p(go), go.stride(0), _p(sc),
                                                      have_lse, _p(lse), _p(delta), _p(dqkv), dqkv.stride(0), _stream()),
               "dgdm_spatial_attention_backward_tc")
        return dqkv
    _check(lib.dgdm_spatial_attention_backward(_p(_f32c(qkv)), qkv.stride(0), _p(pos), _p(seg_of), _p(seg_ptr), N, D, heads,
                                               float(temperature), _p(_f32c(out)), out.stride(0), _p(go), go.stride(0), have_lse, _p(lse),
                                               _p(delta), _p(dqkv), dqkv.stride(0), _stream()), "dgdm_spatial_attention_backward")
    return dqkv


def spatial_attention_weights(qkv, pos, heads, temperature=1.0):
    lib = _lib.load()
    n, D3 = qkv.shape
    w = torch.empty(n, n, dtype=torch.float32, device=qkv.device)
    _check(lib.dgdm_spatial_attention_weights(_p(_f32c(qkv)), qkv.stride(0), _p(pos), n, D3 // 3, heads, float(temperature),
                                              _p(w), _stream()), "dgdm_spatial_attention_weights")
    return w


def mha_dense(q, k, v, B, Lq, Lk, heads, *, attn_mask=None, key_padding_mask=None, pos_q=None, pos_k=None, pos_by_head=False,
              temperature=1.0, need_weights=True, per_head=False):
    """Dense MultiHeadAttention core on projected rows q [B*Lq, D], k / v [B*Lk, D] (row stride may exceed D).
    attn_mask: float (additive) or bool (True = -inf), any shape broadcastable to [B, H, Lq, Lk] by torch's rules.
    Returns (out [B*Lq, D], weights [B*H, Lq, Lk] | [B, Lq, Lk] | None)."""
    lib = _lib.load()
    D = q.size(1)
    dev = q.device
    fmask = bmask = None
    strides = (0, 0, 0, 0)
    if attn_mask is not None:
        m = attn_mask
        if m.dim() > 4:
            raise KernelError(f"attn_mask with {m.dim()} dims")
        m4 = m.reshape((1,) * (4 - m.dim()) + tuple(m.shape))
        for have, want in zip(m4.shape, (B, heads, Lq, Lk)):
            if have not in (1, want):           # same failure torch's `attn_scores += attn_mask` raises
                raise RuntimeError(f"The size of tensor a ({(B, heads, Lq, Lk)}) must match the size of tensor b ({tuple(m.shape)}) "
                                   "at a non-singleton dimension")
        strides = tuple(0 if m4.size(i) == 1 else m4.stride(i) for i in range(4))
        if m.dtype == torch.bool:
            bmask = m4
        else:
            fmask = m4 = m4.float()
            strides = tuple(0 if m4.size(i) == 1 else m4.stride(i) for i in range(4))
    kpm = None
    if key_padding_mask is not None:
        kpm = key_padding_mask.to(torch.bool).contiguous()
        assert kpm.shape == (B, Lk), (kpm.shape, B, Lk)
    out = torch.empty(B * Lq, D, dtype=torch.float32, device=dev)
    w = None
    if need_weights:
        w = torch.empty((B * heads, Lq, Lk) if per_head else (B, Lq, Lk), dtype=torch.float32, device=dev)
    pq = None if pos_q is None else _f32c(pos_q.float().contiguous())
    pk = None if pos_k is None else _f32c(pos_k.float().contiguous())
    _check(lib.dgdm_mha_dense(_p(_f32c(q)), q.stride(0), _p(_f32c(k)), k.stride(0), _p(_f32c(v)), v.stride(0), B, Lq, Lk, D, heads,
                              _p(fmask), _p(bmask.view(torch.uint8)) if bmask is not None else None, strides[0], strides[1],
                              strides[2], strides[3], _p(kpm.view(torch.uint8)) if kpm is not None else None, _p(pq), _p(pk),
                              int(bool(pos_by_head)), float(temperature), _p(out), out.stride(0), _p(w), int(bool(per_head)),
                              _stream()), "dgdm_mha_dense")
    return out, w


def attention_pool(q_tok, kv, seg_ptr, num_seg, heads):
    lib = _lib.load()
    N, D2 = kv.shape
    D = D2 // 2
    out = torch.empty(num_seg, D, dtype=torch.float32, device=kv.device)
    work = torch.empty(int(lib.dgdm_attention_pool_work_floats(N, D, heads, num_seg)), dtype=torch.float32, device=kv.device)
    _check(lib.dgdm_attention_pool(_p(_f32c(q_tok)), _p(_f32c(kv)), kv.stride(0), _p(seg_ptr), num_seg, N, D, heads, _p(out),
                                   _p(work), _stream()), "dgdm_attention_pool")
    return out


# ------------------------------------------------------------------------------------------------ diffusion
def diffusion_add_noise(x0, noise, sqrt_ac, sqrt_1mac, seg_ptr, seg_t):
    lib = _lib.load()
    N, Cw = x0.shape
    out = torch.empty_like(x0)
    _check(lib.dgdm_diffusion_add_noise(_p(_f32c(x0.contiguous())), _p(_f32c(noise.contiguous())), _p(sqrt_ac), _p(sqrt_1mac),
                                        _p(seg_ptr), _p(seg_t), seg_t.numel(), N, Cw, _p(out), _stream()),
           "dgdm_diffusion_add_noise")
    return out


def groupnorm_silu(x, groups, gamma, beta, eps=1e-5):
    lib = _lib.load()
    M, Cw = x.shape
    out = torch.empty(M, Cw, dtype=torch.float32, device=x.device)
    _check(lib.dgdm_groupnorm_silu(_p(_f32c(x)), x.stride(0), M, Cw, groups, _p(gamma), _p(beta), eps, _p(out), out.stride(0),
                                   _stream()), "dgdm_groupnorm_silu")
    return out


def ddpm_update(x, eps_pred, noise, sqrt_1mac, inv_sqrt_ac, sqrt_alpha, sqrt_var):
    lib = _lib.load()
    out = torch.empty_like(x)
    _check(lib.dgdm_ddpm_update(_p(_f32c(x)), _p(_f32c(eps_pred)), _p(noise), sqrt_1mac, inv_sqrt_ac, sqrt_alpha, sqrt_var,
                                x.numel(), _p(out), _stream()), "dgdm_ddpm_update")
    return out


# ------------------------------------------------------------------------------------------------ validation
def validate_inputs_device(x, edge_index, N, buf=None):
    """The validation kernel only: returns the 32-byte device buffer (flags[4] int32 | stats[4] float32), no readback — capturable."""
    lib = _lib.load()
    if buf is None:
        buf = torch.empty(8, dtype=torch.int32, device=x.device)
    flags, stats = buf[:4], buf[4:].view(torch.float32)
    _check(lib.dgdm_validate_inputs(_p(x), x.numel(), _p(edge_index), edge_index.numel(), N, _p(flags), _p(stats), _stream()),
           "dgdm_validate_inputs")
    return buf


def validate_inputs(x, edge_index, N):
    """One kernel + one 32-byte readback. Returns (flags[4] ints, stats[4] floats) on the host."""
    host = validate_inputs_device(x, edge_index, N).cpu()
    return host[:4].tolist(), host[4:].view(torch.float32).tolist()
