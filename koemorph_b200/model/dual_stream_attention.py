"""B200 drop-in for the reference's ``src/model/dual_stream_attention.py``.

Same constructor, parameter names / shapes (``state_dict`` compatible, SURVEY.md section 8 a-W)
and ``forward`` contract as ``DualStreamCrossAttention`` (reference ``:48-294``); the arithmetic
runs in the hand-written sm_100a kernels behind ``koe_dual_stream_features`` /
``koe_emotion_stream`` (include/koemorph_b200.h).  Inference only (eval semantics: dropout off).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import _lib
from .folding import EXPRESSION_INDICES, MOUTH_INDICES, CoreWeights, fold

# ARKit coefficient order, dual_stream_attention.py:29-41 (public constant of the reference module)
ARKIT_BLENDSHAPES = (
    "eyeBlinkLeft eyeLookDownLeft eyeLookInLeft eyeLookOutLeft eyeLookUpLeft eyeSquintLeft eyeWideLeft "
    "eyeBlinkRight eyeLookDownRight eyeLookInRight eyeLookOutRight eyeLookUpRight eyeSquintRight eyeWideRight "
    "jawForward jawLeft jawRight jawOpen mouthClose mouthFunnel mouthPucker mouthLeft mouthRight mouthSmileLeft "
    "mouthSmileRight mouthFrownLeft mouthFrownRight mouthDimpleLeft mouthDimpleRight mouthStretchLeft "
    "mouthStretchRight mouthRollLower mouthRollUpper mouthShrugLower mouthShrugUpper mouthPressLeft "
    "mouthPressRight mouthLowerDownLeft mouthLowerDownRight mouthUpperUpLeft mouthUpperUpRight browDownLeft "
    "browDownRight browInnerUp browOuterUpLeft browOuterUpRight cheekPuff cheekSquintLeft cheekSquintRight "
    "noseSneerLeft noseSneerRight tongueOut").split()
MOUTH_BLENDSHAPES = [ARKIT_BLENDSHAPES[i] for i in MOUTH_INDICES]


class DualStreamCrossAttention(nn.Module):
    """28 mouth queries over 80 mel-channel tokens + 24 expression queries over one eGeMAPS token."""

    def __init__(self, d_model: int = 256, num_heads: int = 8, num_mel_channels: int = 80,
                 mel_sequence_length: int = 256, mel_temporal_frames: int = 3, emotion_dim: int = 256,
                 emotion_sequence_length: int = 1, dropout: float = 0.1, num_blendshapes: int = 52,
                 use_learnable_weights: bool = True, temperature: float = 1.0):
        super().__init__()
        if (d_model, num_heads, num_mel_channels, num_blendshapes, mel_temporal_frames) != (256, 8, 80, 52, 3):
            raise NotImplementedError(
                "koemorph_b200 kernels are specialised for d_model=256, num_heads=8, 80 mel channels, 52 blendshapes, "
                "3 temporal frames (the configuration of SimplifiedDualStreamModel, reference :148-159)")
        self.d_model, self.num_heads = d_model, num_heads
        self.num_mel_channels, self.mel_sequence_length = num_mel_channels, mel_sequence_length
        self.mel_temporal_frames, self.emotion_dim = mel_temporal_frames, emotion_dim
        self.emotion_sequence_length, self.num_blendshapes = emotion_sequence_length, num_blendshapes
        self.temperature = temperature
        self.total_mel_dim = num_mel_channels * (mel_sequence_length + mel_temporal_frames)
        # parameter creation order follows the reference constructor (:102-160) so that a seeded
        # construction draws the same random initial weights
        self.mel_channel_encoder = nn.Linear(mel_sequence_length + mel_temporal_frames, d_model)
        self.mel_attention = nn.MultiheadAttention(d_model, num_heads, dropout=dropout, batch_first=True)
        self.emotion_encoder = nn.Linear(emotion_dim, d_model)
        self.emotion_attention = nn.MultiheadAttention(d_model, num_heads, dropout=dropout, batch_first=True)
        self.mouth_queries = nn.Parameter(torch.randn(len(MOUTH_INDICES), d_model) * 0.02)
        self.expression_queries = nn.Parameter(torch.randn(len(EXPRESSION_INDICES), d_model) * 0.02)
        if use_learnable_weights:
            self.mel_weights = nn.Parameter(torch.ones(num_blendshapes))
            self.emotion_weights = nn.Parameter(torch.ones(num_blendshapes))
            with torch.no_grad():
                self.mel_weights[MOUTH_INDICES] = 2.0
                self.mel_weights[EXPRESSION_INDICES] = 0.5
                self.emotion_weights[MOUTH_INDICES] = 0.5
                self.emotion_weights[EXPRESSION_INDICES] = 2.0
        else:
            mw, ew = torch.zeros(num_blendshapes), torch.zeros(num_blendshapes)
            mw[MOUTH_INDICES] = 1.0
            ew[EXPRESSION_INDICES] = 1.0
            self.register_buffer("mel_weights", mw)
            self.register_buffer("emotion_weights", ew)
        self.mel_output_proj = nn.Linear(d_model, d_model)
        self.emotion_output_proj = nn.Linear(d_model, d_model)
        self.blendshape_decoder = nn.Sequential(nn.Linear(d_model, d_model // 2), nn.ReLU(), nn.Dropout(dropout),
                                                nn.Linear(d_model // 2, 1), nn.Sigmoid())
        self.mel_norm = nn.LayerNorm(d_model)
        self.emotion_norm = nn.LayerNorm(d_model)
        self.precision = "fp32"  # "fp32" (CUDA-core FMA) | "bf16" (bf16 operands on tcgen05, fp32 accumulation)
        self._folded: Dict = {}
        # bumped by load_state_dict / invalidate_kernel_weights(): in-place edits through ``param.data`` do not change
        # ``_version``, so code that writes weights that way must call invalidate_kernel_weights() itself
        self._generation = 0
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_kernel_weights())

    def invalidate_kernel_weights(self) -> None:
        """Drop the folded kernel weights: call after editing parameters in place through ``.data`` (such writes do not
        bump the version counter the cache key uses)."""
        self._generation += 1
        self._folded.clear()

    # ---- folded kernel weights, rebuilt when any parameter (or the compression layer) changes --------------
    def _weight_slots(self):
        """(owner dict, name) of every parameter / buffer, in named_parameters() + named_buffers() order.  The module tree
        is fixed after construction, so the walk is done once; looking the tensors up through the owners' dicts still sees
        replaced tensors (``.to()``, a new ``nn.Parameter``)."""
        slots = getattr(self, "_slots", None)
        if slots is None:
            slots = []
            for prefix, mod in self.named_modules():
                for name in mod._parameters:
                    slots.append((mod._parameters, name, (prefix + "." if prefix else "") + name))
            for prefix, mod in self.named_modules():
                for name in mod._buffers:
                    if name not in mod._non_persistent_buffers_set:
                        slots.append((mod._buffers, name, (prefix + "." if prefix else "") + name))
            self._slots = slots
        return slots

    def kernel_weights(self, compression: Optional[Dict[str, torch.Tensor]] = None) -> CoreWeights:
        slots = self._weight_slots()
        slots = [sl for sl in slots if sl[0][sl[1]] is not None]   # (nn.MultiheadAttention registers unused None slots)
        tensors = [d[n] for d, n, _ in slots]
        comp = [] if compression is None else [compression["weight"], compression["bias"]]
        key = tuple((t.data_ptr(), t._version) for t in tensors) + \
            tuple((t.data_ptr(), t._version) for t in comp) + (self.temperature, self._generation)
        slot = "comp" if compression is not None else "plain"
        hit = self._folded.get(slot)
        if hit is None or hit[0] != key:
            device = self.mel_norm.weight.device
            if device.type != "cuda":
                raise RuntimeError(f"DualStreamCrossAttention parameters are on {device}; move the module to a CUDA "
                                   "device (koemorph_b200 has no CPU path)")
            sd = {full: t for (_, _, full), t in zip(slots, tensors)}
            hit = (key, fold(sd, self.num_heads, self.temperature, device, compression, eps=self.mel_norm.eps))
            self._folded[slot] = hit
        return hit[1]

    @torch.no_grad()
    def forward(self, mel_features: torch.Tensor, mel_temporal_features: torch.Tensor,
                emotion_features: torch.Tensor, return_attention: bool = False) -> Dict[str, torch.Tensor]:
        """mel_features (B, T, 80), mel_temporal_features (B, 3, 80), emotion_features (B, 256) -> dict
        (reference :162-280).  T is zero padded / truncated to ``mel_sequence_length`` (:192-202)."""
        mel = _lib.require_cuda(mel_features, "mel_features")
        short = _lib.require_cuda(mel_temporal_features, "mel_temporal_features")
        emo = _lib.require_cuda(emotion_features, "emotion_features")
        if mel.dim() != 3 or mel.shape[2] != 80 or short.shape != (mel.shape[0], 3, 80):
            raise ValueError(f"expected mel (B, T, 80) and temporal (B, 3, 80), got {tuple(mel.shape)}, {tuple(short.shape)}")
        if emo.dim() != 2 or emo.shape != (mel.shape[0], self.emotion_dim):
            raise ValueError(f"expected emotion_features (B, {self.emotion_dim}), got {tuple(emo.shape)}")
        B, T = mel.shape[0], mel.shape[1]
        w = self.kernel_weights()
        lib = _lib.load()
        dev = mel.device
        expr = torch.empty(B, dtype=torch.float32, device=dev)
        out = torch.empty(B, 52, dtype=torch.float32, device=dev)
        sig = torch.empty(B, 52, dtype=torch.float32, device=dev) if return_attention else None
        attn = torch.empty(B, 28, 80, dtype=torch.float32, device=dev) if return_attention else None
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            _lib.check(lib.koe_emotion_stream(C.byref(w.struct), emo.data_ptr(), B, expr.data_ptr(), st),
                       "koe_emotion_stream")
            _lib.check(lib.koe_dual_stream_features(
                C.byref(w.struct), mel.data_ptr(), T, short.data_ptr(), B, expr.data_ptr(), out.data_ptr(),
                sig.data_ptr() if sig is not None else None, attn.data_ptr() if attn is not None else None,
                _lib.PRECISIONS[self.precision], st), "koe_dual_stream_features")
        return _package(out, sig, attn, return_attention)

    def get_frequency_bands(self) -> Dict[str, List[int]]:
        """Mel-channel groups used by the attention visualiser (reference :282-294)."""
        return {"low": list(range(0, 20)), "mid_low": list(range(20, 40)),
                "mid_high": list(range(40, 60)), "high": list(range(60, 80))}


def _package(out, sig, attn, return_attention):
    """Output dict of the reference (:272-280); sig is the decoder output before stream-weight fusion."""
    res = {"blendshapes": out}
    if return_attention:
        res["mel_attention_weights"] = attn
        lead = out.shape[:-1]
        # one key per clip -> softmax weight is exactly 1 for each of the 24 expression queries (:234-239)
        res["emotion_attention_weights"] = torch.ones(*lead, len(EXPRESSION_INDICES), 1, dtype=out.dtype,
                                                      device=out.device)
        mb, eb = torch.zeros_like(sig), torch.zeros_like(sig)
        mb[..., MOUTH_INDICES] = sig[..., MOUTH_INDICES]
        eb[..., EXPRESSION_INDICES] = sig[..., EXPRESSION_INDICES]
        res["mel_blendshapes"], res["emotion_blendshapes"] = mb, eb
    return res
