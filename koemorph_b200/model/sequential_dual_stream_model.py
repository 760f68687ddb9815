"""B200 drop-in for the reference's ``src/model/sequential_dual_stream_model.py``.

The reference slides a window of ``mel_sequence_length`` hops over the clip and, for EVERY output
frame, recomputes the whole window's librosa mel (``:101-120``) before running the core and the EMA.
Here each STFT frame is computed once: frame k of window i is global frame i*stride + k, except the
frames within n_fft/2 of a window edge, which see zeros beyond the edge (note E).
Those are produced as separate "edge variants" by ``koe_logmel_power`` (lo_rel_hops / hi_rel_hops) on
the window grid; the dB reference of a window is the max over ITS frames, taken inside the core kernel.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from .. import _lib
from .dual_stream_attention import _package
from .simplified_dual_stream_model import SimplifiedDualStreamModel


class SequentialDualStreamModel(SimplifiedDualStreamModel):
    """audio (B, L) -> blendshapes (B, T_out, 52), T_out = max(1, (L // hop - W) // stride + 1)."""

    def __init__(self, d_model: int = 256, num_heads: int = 8, num_blendshapes: int = 52, sample_rate: int = 16000,
                 target_fps: int = 30, mel_sequence_length: int = 256, emotion_config: Optional[Dict] = None,
                 device: str = "cuda", real_time_mode: bool = False, stride_frames: int = 1):
        super().__init__(d_model=d_model, num_heads=num_heads, num_blendshapes=num_blendshapes,
                         sample_rate=sample_rate, target_fps=target_fps, mel_sequence_length=mel_sequence_length,
                         emotion_config=emotion_config, device=device, real_time_mode=real_time_mode)
        if stride_frames < 1:
            raise ValueError("stride_frames must be >= 1")
        self.stride_frames = stride_frames
        self.window_frames = mel_sequence_length
        self.window_samples = self.window_frames * self.hop_length
        self.stride_samples = self.stride_frames * self.hop_length

    def num_output_frames(self, audio_length: int) -> int:
        """reference :84,96."""
        return max(1, (audio_length // self.hop_length - self.window_frames) // self.stride_frames + 1)

    def _forward_frames(self, audio, eg, return_attention, out=None):
        """The stateless part of forward(): (B, T_out, 52) smoothed blendshapes (+ sigmoid, attention)."""
        B, L = audio.shape
        hop, W, stride = self.hop_length, self.window_frames, self.stride_frames
        n_out = self.num_output_frames(L)
        T_w = W + 1                                        # librosa: 1 + (W * hop) // hop frames per window
        n_frames = (n_out - 1) * stride + T_w              # global frames touched
        # a single window that already spans the whole (zero padded) clip has no interior edges
        n_edge = 0 if (n_out == 1 and L <= self.window_samples) else math.ceil((self.n_fft // 2) / hop)
        if n_edge > _lib.MAX_EDGE:
            raise NotImplementedError(f"hop {hop} needs {n_edge} edge variants per side (max {_lib.MAX_EDGE})")
        return self._forward_windows(audio, eg, n_frames, T_w, stride, n_out, n_edge, self.use_temporal_smoothing,
                                     return_attention, out=out)

    @torch.no_grad()
    def forward(self, audio: torch.Tensor, return_attention: bool = False,
                egemaps: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """reference :63-167.  ``out`` (extension): a preallocated contiguous (B, T_out, 52) float32 CUDA tensor that
        receives the blendshapes (a corpus sweep lets every batch write where the results will be gathered)."""
        audio = self._check_audio(audio)
        B, L = audio.shape
        eg = self._check_egemaps(egemaps, B, audio.device)
        out, sig, attn = self._forward_frames(audio, eg, return_attention, out=out)
        # the reference leaves the smoothing state at the last frame of the sequence (:99,136): a view, like its .detach()
        self.prev_blendshapes = out[:, -1] if self.use_temporal_smoothing else None
        res = _package(out, sig, attn, return_attention)
        res["num_frames"] = out.shape[1]
        res["fps"] = self.target_fps
        res["emotion_backend"] = "egemaps_input"
        res["emotion_processing_time"] = 0.0
        return res

    @torch.no_grad()
    def forward_single_frame(self, audio: torch.Tensor, frame_idx: int = None,
                             egemaps: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """reference :169-181."""
        return SimplifiedDualStreamModel.forward(self, audio, return_attention=False, egemaps=egemaps)
