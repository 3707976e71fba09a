"""ContrastiveLoss — drop-in for training/losses.py:73-176 (the within-batch InfoNCE path).

The reference forms the dense N x N matrix exp(E E^T / T) (4 N^2 bytes: 40 GB at 100k nodes). Here the similarity matrix is produced
512 key columns at a time by the split-fp16 tcgen05 GEMM (fp32-class products) and folded straight into per-row sums; nothing of
size N^2 is stored. Forward only (like the rest of the package): the value is returned without an autograd graph."""
import math

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import GemmParams
from .ops import KernelError, _check, _f32c, _p, _stream

import ctypes as C

KEY_CHUNK = 512


def _gram_chunk(a, img, n_img_rows, r0, nrows, out):
    """out[:, :nrows] = a @ W[r0:r0+nrows]^T where W is the fp16 hi/lo image of a [n_img_rows, K] matrix (dgdm_gemm_f16x2_strided)."""
    lib = _lib.load()
    p = GemmParams()
    M, K = a.shape
    p.A, p.lda, p.M, p.K = _p(a), a.stride(0), M, K
    base = img.img.data_ptr() + r0 * 128
    p.W_hi, p.W_lo, p.ldw, p.N = C.c_void_p(base), C.c_void_p(base + n_img_rows * 128), img.Kpad, nrows
    p.C, p.ldc = _p(out), out.stride(0)
    p.a_scale, p.ln_eps = 1.0, 1e-5
    _check(lib.dgdm_gemm_f16x2_strided(C.byref(p), n_img_rows * 128, _stream()), "dgdm_gemm_f16x2_strided")


class ContrastiveLoss(nn.Module):
    """Same constructor and call signature as the reference (training/losses.py:80-121)."""

    def __init__(self, temperature=0.1, similarity_function="cosine", reduction="mean"):
        super().__init__()
        self.temperature, self.similarity_function, self.reduction = temperature, similarity_function, reduction

    @torch.no_grad()
    def forward(self, embeddings, batch_indices=None, positive_pairs=None):
        if positive_pairs is not None:
            return self._pair_contrastive_loss(embeddings, positive_pairs)
        return self._batch_contrastive_loss(embeddings, batch_indices)

    def _batch_contrastive_loss(self, embeddings, batch_indices):
        if self.similarity_function not in ("cosine", "dot"):
            raise ValueError(f"Unknown similarity function: {self.similarity_function}")
        if not embeddings.is_cuda:
            raise KernelError("ContrastiveLoss runs on a B200 (CUDA) device only; there is no CPU fallback")
        lib = _lib.load()
        x = _f32c(embeddings.float().contiguous())
        n, d = x.shape
        dev = x.device
        n_pad = (n + 15) // 16 * 16
        e = torch.zeros(n_pad, d, dtype=torch.float32, device=dev)
        _check(lib.dgdm_row_normalize(_p(x), x.stride(0), n, d, 1e-12, _p(e), e.stride(0), _stream()), "dgdm_row_normalize")     # both branches of
        img = ops.SplitWeight(e)                                                     # :134-137 multiply the NORMALISED rows
        batch = None if batch_indices is None else batch_indices.to(torch.int64).contiguous()
        sum_all = torch.zeros(n, dtype=torch.float32, device=dev)
        sum_pos = torch.zeros(n, dtype=torch.float32, device=dev)
        cnt_pos = torch.zeros(n, dtype=torch.int32, device=dev)
        row_chunk = 65536
        for r0 in range(0, n, row_chunk):
            m = min(row_chunk, n - r0)
            a = e[r0:r0 + m]
            s = torch.empty(m, KEY_CHUNK, dtype=torch.float32, device=dev)
            for k0 in range(0, n, KEY_CHUNK):
                nk = min(KEY_CHUNK, n - k0)
                nk16 = (nk + 15) // 16 * 16                                          # padded key rows are zero rows of the image
                _gram_chunk(a, img, n_pad, k0, nk16, s)
                _check(lib.dgdm_contrastive_accumulate(_p(s), s.stride(0), m, nk, r0, k0, _p(batch), float(self.temperature),
                                                       C.c_void_p(sum_all.data_ptr() + 4 * r0), C.c_void_p(sum_pos.data_ptr() + 4 * r0),
                                                       C.c_void_p(cnt_pos.data_ptr() + 4 * r0), _stream()), "dgdm_contrastive_accumulate")
        # -log(clamp(sum_pos, 1e-8) / sum_all) with the 1/T shift restored inside the logarithms (:160-166); rows without positives
        # are dropped (:168-170). O(N) bookkeeping on graph-level vectors.
        shift = 1.0 / float(self.temperature)
        log_all = torch.log(sum_all) + shift
        log_pos = torch.clamp(torch.log(sum_pos) + shift, min=math.log(1e-8))
        loss = (log_all - log_pos)[cnt_pos > 0]
        if self.reduction == "mean":
            return loss.mean()
        if self.reduction == "sum":
            return loss.sum()
        return loss

    def _pair_contrastive_loss(self, embeddings, positive_pairs):
        raise NotImplementedError("explicit positive pairs (training/losses.py:178-210: a Python loop of one softmax row per pair) are not "
                                  "on the accelerated path")
