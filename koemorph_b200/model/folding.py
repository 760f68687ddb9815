"""Fold the batch-invariant products of DualStreamCrossAttention into the kernel weight pack.

Reference forward: src/model/dual_stream_attention.py:162-280.  In eval mode several of its
steps are products of parameters only, so they are multiplied out once per weight update
(in float64 on the device, then stored float32):

* the 28 mouth queries and ``W_q`` never see the input (``:221,225-230``): with
  ``Qt = (Wq q + bq) / sqrt(head_dim)`` the per-head scores are ``Qt_h (Wk_h enc + bk_h)``.  ``Qt_h . bk_h``
  is constant along the key axis, which softmax ignores, so scores = ``(Qt_h Wk_h) enc``: one
  [224 x 256] matrix ``Qk`` replaces the K projection and the QK^T product;
* ``out_proj``, ``mel_output_proj`` (``:231``) and ``blendshape_decoder.0`` (``:150-151``) are consecutive
  affine maps: ``Wa = W1 Wmo Wo`` [128 x 256];
* the emotion stream attends over ONE key (``:234-239``), so its softmax is 1 and every expression query
  yields ``out_proj(Wv e + bv)``; together with ``emotion_output_proj`` and ``decoder.0`` that is one
  [128 x 256] affine map ``We2`` after the LayerNorm; before it, ``emotion_encoder`` absorbs the
  264 -> 256 eGeMAPS compression (src/features/opensmile_extractor.py:586-602);
* stream-weight fusion (``:252-267``): ``0.5 * (softmax(mel_w / T) + softmax(emo_w / T))`` per coefficient.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from .. import _lib

MOUTH_INDICES = list(range(14, 41)) + [51]          # dual_stream_attention.py:44
EXPRESSION_INDICES = [i for i in range(52) if i not in MOUTH_INDICES]  # :45


TC_STAGE_BYTES = 16384


def _tile_kmajor(mat: torch.Tensor, rows_per_stage: int, k_per_stage: int, k_pad: int):
    """Cut a [rows, K] matrix into pipeline-stage images in the UMMA no-swizzle K-major core-matrix layout:
    inside a stage, element (r, k) sits at (r//8)*(chunks*128) + (k//8)*128 + (r%8)*16 + (k%8)*2 bytes (bf16)."""
    rows, k = mat.shape
    m = torch.zeros(rows, k_pad, dtype=mat.dtype, device=mat.device)
    m[:, :k] = mat
    out = []
    for r0 in range(0, rows, rows_per_stage):
        for k0 in range(0, k_pad, k_per_stage):
            sub = m[r0:r0 + rows_per_stage, k0:k0 + k_per_stage]
            img = sub.reshape(rows_per_stage // 8, 8, k_per_stage // 8, 8).permute(0, 2, 1, 3).contiguous()
            out.append(img.reshape(-1))
    return out


def _ceil_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class CoreWeights:
    """Device-resident, kernel-ready weights + the ctypes struct that points at them."""

    def __init__(self, tensors: Dict[str, torch.Tensor], k_mel: int, emo_in: int, b2: float, eps: float):
        self.tensors = tensors  # keeps the storage alive
        s = _lib.CoreWeightsStruct()
        s.k_mel, s.k_mel_pad = k_mel, tensors["wc_t"].shape[0]
        s.emo_in, s.emo_in_pad = emo_in, tensors["we1_t"].shape[0]
        s.b2, s.ln_eps = b2, eps
        for name in ("wc_t", "bc", "ln_g", "ln_b", "qk_t", "wv_t", "bv", "wa_t", "ba", "w2", "coef", "mouth_idx",
                     "expr_idx", "we1_t", "be1", "eln_g", "eln_b", "we2_t", "be2"):
            setattr(s, name, tensors[name].data_ptr())
        tc = tensors.get("tc_bf16")
        s.tc_bf16 = tc.data_ptr() if tc is not None else None
        s.tc_stages = 0 if tc is None else tc.numel() * 2 // TC_STAGE_BYTES
        s.tc_bv = tensors["tc_bv"].data_ptr() if "tc_bv" in tensors else None
        self.struct = s
        self.device = tensors["wc_t"].device


@torch.no_grad()
def fold(sd: Dict[str, torch.Tensor], num_heads: int, temperature: float, device,
         compression: Optional[Dict[str, torch.Tensor]] = None, eps: float = 1e-5) -> CoreWeights:
    """sd: parameters of DualStreamCrossAttention keyed as in its state_dict (no prefix)."""
    dd = torch.float64
    g = lambda k: sd[k].detach().to(device=device, dtype=dd)
    d = g("mel_norm.weight").shape[0]
    if d != 256 or num_heads != 8:
        raise NotImplementedError("the CUDA core is built for d_model=256, num_heads=8 (the reference configuration)")
    hd = d // num_heads
    # ---- mel stream
    wc, bc = g("mel_channel_encoder.weight"), g("mel_channel_encoder.bias")
    k_mel = wc.shape[1]
    inw, inb = g("mel_attention.in_proj_weight"), g("mel_attention.in_proj_bias")
    wq, wk, wv = inw[:d], inw[d:2 * d], inw[2 * d:]
    bq, bv = inb[:d], inb[2 * d:]
    qt = (g("mouth_queries") @ wq.T + bq) / math.sqrt(hd)                       # [28, 256]
    nq = qt.shape[0]
    qk = torch.zeros(256, d, dtype=dd, device=device)                           # rows h*28+q, padded to 256
    for h in range(num_heads):
        qk[h * nq:(h + 1) * nq] = qt[:, h * hd:(h + 1) * hd] @ wk[h * hd:(h + 1) * hd]
    w1, b1 = g("blendshape_decoder.0.weight"), g("blendshape_decoder.0.bias")
    wo, bo = g("mel_attention.out_proj.weight"), g("mel_attention.out_proj.bias")
    wmo, bmo = g("mel_output_proj.weight"), g("mel_output_proj.bias")
    wa = w1 @ wmo @ wo
    ba = w1 @ (wmo @ bo + bmo) + b1
    # ---- emotion stream
    we, be = g("emotion_encoder.weight"), g("emotion_encoder.bias")
    if compression is not None:
        cw = compression["weight"].detach().to(device=device, dtype=dd)
        cb = compression["bias"].detach().to(device=device, dtype=dd)
        we1, be1 = we @ cw, we @ cb + be
    else:
        we1, be1 = we, be
    emo_in = we1.shape[1]
    einw, einb = g("emotion_attention.in_proj_weight"), g("emotion_attention.in_proj_bias")
    wve, bve = einw[2 * d:], einb[2 * d:]
    woe, boe = g("emotion_attention.out_proj.weight"), g("emotion_attention.out_proj.bias")
    weo, beo = g("emotion_output_proj.weight"), g("emotion_output_proj.bias")
    we2 = w1 @ weo @ woe @ wve
    be2 = w1 @ (weo @ (woe @ bve + boe) + beo) + b1
    coef = 0.5 * (torch.softmax(g("mel_weights") / temperature, 0) + torch.softmax(g("emotion_weights") / temperature, 0))

    def pad_rows(m, rows):
        out = torch.zeros(rows, m.shape[1], dtype=dd, device=device)
        out[:m.shape[0]] = m
        return out

    f32 = lambda t: t.to(torch.float32).contiguous()
    tensors = {
        "wc_t": f32(pad_rows(wc.T, _ceil_to(k_mel, 16))), "bc": f32(bc),
        "ln_g": f32(g("mel_norm.weight")), "ln_b": f32(g("mel_norm.bias")),
        "qk_t": f32(qk.T), "wv_t": f32(wv.T), "bv": f32(bv),
        "wa_t": f32(wa.T), "ba": f32(ba), "w2": f32(g("blendshape_decoder.3.weight").reshape(-1)),
        "coef": f32(coef),
        "mouth_idx": torch.tensor(MOUTH_INDICES, dtype=torch.int32, device=device),
        "expr_idx": torch.tensor(EXPRESSION_INDICES, dtype=torch.int32, device=device),
        "we1_t": f32(pad_rows(we1.T, _ceil_to(emo_in, 8))), "be1": f32(be1),
        "eln_g": f32(g("emotion_norm.weight")), "eln_b": f32(g("emotion_norm.bias")),
        "we2_t": f32(we2.T), "be2": f32(be2),
    }
    if k_mel in (259, 515):
        # tcgen05 path: stage images in consumption order -- Wc (9 x [256 x 32] at 30 fps, 17 at 60 fps), Qk tiles 0/1,
        # Wv tiles 0/1, Wa (each 4 x [128 x 64]); Qk rows are re-indexed to 32*h + q (queries 28..31 of each head are zero rows)
        # The affine parts around the LayerNorm are folded away (in float64, like everything here): the encoder bias is
        # weight column k_mel of the first GEMM (the kernel puts a constant one in that operand column); mel_norm.weight
        # scales the input columns of the key and value projections; mel_norm.bias adds a per-(head, query) constant to
        # the scores, which the softmax ignores, and Wv . bias to the values, which joins the value bias (tc_bv).
        ln_g, ln_b = g("mel_norm.weight"), g("mel_norm.bias")
        qk_pad = torch.zeros(256, d, dtype=dd, device=device)
        for h in range(num_heads):
            qk_pad[32 * h:32 * h + nq] = qk[h * nq:(h + 1) * nq] * ln_g
        wc_b = torch.cat([wc, bc.reshape(-1, 1)], dim=1)          # [256, k_mel + 1]
        assert k_mel + 1 <= _ceil_to(k_mel, 16)
        tensors["tc_bv"] = f32(wv @ ln_b + bv)
        n_g1 = (_ceil_to(k_mel, 16) + 31) // 32
        stages = _tile_kmajor(wc_b, 256, 32, 32 * n_g1) + _tile_kmajor(qk_pad, 128, 64, 256) + \
            _tile_kmajor(wv * ln_g, 128, 64, 256) + _tile_kmajor(wa, 128, 64, 256)
        tensors["tc_bf16"] = torch.cat(stages).to(torch.bfloat16).contiguous()
        assert tensors["tc_bf16"].numel() * 2 == (n_g1 + 20) * TC_STAGE_BYTES
    b2 = float(sd["blendshape_decoder.3.bias"].detach().reshape(-1)[0])
    return CoreWeights(tensors, k_mel, emo_in, b2, eps)
