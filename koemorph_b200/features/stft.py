"""B200 drop-in for the reference's ``src/features/stft.py`` ``MelSpectrogramExtractor``:
torchaudio ``T.MelSpectrogram(n_fft=512, hop=sr/fps, mel_scale="htk", norm=None, normalized=True, pad_mode="reflect")``
followed by ``log(mel + eps)`` and the truncate / repeat-last-frame fix-up to ``int(L / sr * fps)`` frames
(``stft.py:84-142``).  Same FFT kernel as the log-mel frontend of the dual-stream path (a 512-point frame is transformed
as the middle of a zero-extended 1024-point one), a different bank and epilogue (``koe_frontend_create_ex``)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import _lib
from .mel_frontend import LogMelFrontend


class MelSpectrogramExtractor(nn.Module):
    """waveform (B, L) or (L,) float32 CUDA -> log-mel (B, T, 80), T = int(L / sample_rate * target_fps)."""

    def __init__(self, sample_rate: int = 16000, target_fps: float = 30.0, n_fft: int = 512, n_mels: int = 80,
                 f_min: float = 80.0, f_max: Optional[float] = None, power: float = 2.0, normalized: bool = True,
                 center: bool = True, pad_mode: str = "reflect", eps: float = 1e-8):
        super().__init__()
        if n_fft not in (512, 1024) or n_mels != 80:
            raise NotImplementedError("the CUDA frontend implements n_fft 512 / 1024 and 80 mel bands")
        if power != 2.0 or not center or pad_mode not in ("reflect", "constant"):
            raise NotImplementedError("the CUDA frontend implements power=2.0, center=True, pad_mode reflect / constant")
        self.sample_rate, self.target_fps, self.n_fft, self.n_mels = sample_rate, target_fps, n_fft, n_mels
        self.f_min, self.f_max = f_min, f_max or sample_rate // 2
        self.power, self.eps, self.normalized, self.pad_mode = power, eps, normalized, pad_mode
        self.hop_length = int(sample_rate / target_fps)          # stft.py:76
        self.win_length = n_fft
        if self.hop_length <= 0:
            raise ValueError(f"Invalid hop_length {self.hop_length} for sr={sample_rate}, fps={target_fps}")
        self.register_buffer("mel_scale", torch.zeros(n_fft // 2 + 1, n_mels), persistent=True)  # filled on first use
        self._filled = False

    def _frontend(self, device) -> LogMelFrontend:
        fe = LogMelFrontend.get(device, self.sample_rate, self.n_fft, self.n_mels, self.f_min, float(self.f_max),
                                mel_scale="htk", window_normalized=self.normalized, log_mode="ln", log_eps=self.eps)
        if not self._filled:
            self.mel_scale = torch.from_numpy(fe.filterbank().T.copy()).to(self.mel_scale.device)
            self._filled = True
        return fe

    @torch.no_grad()
    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        if waveform.dim() == 1:
            waveform = waveform.unsqueeze(0)
        if waveform.dim() != 2:
            raise ValueError(f"Expected 1D or 2D input, got {waveform.dim()}D")
        x = _lib.require_cuda(waveform, "waveform")
        L = x.shape[1]
        if L <= 512:
            raise ValueError("waveform too short: reflect padding needs more than 512 samples")
        n_frames = 1 + L // self.hop_length                      # torch.stft(center=True)
        log_mel, _ = self._frontend(x.device).power(x, self.hop_length, n_frames, pad_mode=self.pad_mode)
        expected = int(L / self.sample_rate * self.target_fps)   # stft.py:128-140
        if n_frames > expected:
            log_mel = log_mel[:, :expected, :]
        elif n_frames < expected:
            log_mel = torch.cat([log_mel, log_mel[:, -1:, :].repeat(1, expected - n_frames, 1)], dim=1)
        return log_mel

    def get_output_length(self, input_length: int) -> int:
        """stft.py:144-158."""
        input_length += 2 * (self.n_fft // 2)
        return (input_length - self.n_fft) // self.hop_length + 1

    def get_time_axis(self, seq_length: int) -> torch.Tensor:
        """stft.py:160-173."""
        return torch.arange(seq_length, dtype=torch.float32) * self.hop_length / self.sample_rate
