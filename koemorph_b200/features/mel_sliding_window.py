"""B200 drop-in for the reference's ``src/features/mel_sliding_window.py``.

``MelAudioBuffer`` (reference ``:21-154``) and ``MelSlidingWindowExtractor`` (``:157-412``) keep the
reference's constructor arguments, methods and return types; the ring lives in device memory and the whole-ring
log-mel of ``process_audio_frame`` (reference ``:280-307``: librosa mel, reflect padding, ``power_to_db(ref=max)``
in [-80, 0] dB, truncate / last-frame-pad to ``int(context / update)`` frames) is one ``koe_logmel_power_ex`` +
``koe_logmel_normalise`` pair on the current CUDA stream.

Quirks of the reference that are preserved because callers can observe them:
the ring advances by ``int(sample_rate / (1 / update_interval))`` = 532 samples per frame while the STFT hop is
533; frames of +-1 sample are padded / truncated; a wall-clock throttle returns the cached features when called
again within 30 % of the update interval.  The hop-aligned incremental path (3 FFTs per hop instead of 256) is
``koemorph_b200.streaming.StreamingEngine``.
"""
from __future__ import annotations

import logging
import threading
import time
from collections import deque
from typing import Any, Dict, Optional, Union

import numpy as np
import torch

from .. import _lib
from .mel_frontend import LogMelFrontend

logger = logging.getLogger(__name__)
ArrayLike = Union[np.ndarray, torch.Tensor]


def _resolve_device(device) -> torch.device:
    dev = torch.device("cuda" if device in (None, "cpu") and torch.cuda.is_available() else device)
    if dev.type != "cuda":
        raise RuntimeError(f"koemorph_b200 mel extraction needs a CUDA device (got {device!r}); there is no CPU path")
    return torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())


class MelAudioBuffer:
    """Circular audio buffer of ``context_window`` seconds in device memory (reference ``:21-154``)."""

    def __init__(self, context_window: float = 8.5, sample_rate: int = 16000, update_interval: float = 0.0333,
                 device="cuda"):
        self.context_window, self.sample_rate, self.update_interval = context_window, sample_rate, update_interval
        self.buffer_size = int(context_window * sample_rate)
        target_fps = 1.0 / update_interval
        self.hop_length = int(sample_rate / target_fps)   # 532 for 0.0333 s (reference :49-50)
        self.device = _resolve_device(device)
        self.audio_buffer = torch.zeros(self.buffer_size, dtype=torch.float32, device=self.device)
        self.write_ptr = 0
        self.is_full = False
        self._lock = threading.Lock()
        self.total_frames_added = 0
        self.buffer_overruns = 0

    def add_audio_frame(self, audio_frame: ArrayLike) -> bool:
        """Append one hop (+-1 sample tolerated: padded with a zero / truncated), reference ``:70-116``."""
        frame = torch.as_tensor(audio_frame, dtype=torch.float32).reshape(-1)
        n = frame.numel()
        if abs(n - self.hop_length) > 1:
            logger.warning(f"Frame size mismatch: expected ~{self.hop_length}, got {n}")
            return False
        if n < self.hop_length:
            frame = torch.cat([frame, frame.new_zeros(self.hop_length - n)])
        elif n > self.hop_length:
            frame = frame[:self.hop_length]
        frame = frame.to(self.device, non_blocking=True)
        with self._lock:
            end = (self.write_ptr + self.hop_length) % self.buffer_size
            if end > self.write_ptr:
                self.audio_buffer[self.write_ptr:end] = frame
            else:
                first = self.buffer_size - self.write_ptr
                self.audio_buffer[self.write_ptr:] = frame[:first]
                self.audio_buffer[:end] = frame[first:]
            self.write_ptr = end
            self.total_frames_added += 1
            if not self.is_full and self.total_frames_added * self.hop_length >= self.buffer_size:
                self.is_full = True
        return True

    def get_current_audio(self, as_tensor: bool = False):
        """Chronological window of the whole ring, or None until it has filled once (reference ``:118-140``)."""
        with self._lock:
            if not self.is_full:
                return None
            win = self.audio_buffer if self.write_ptr == 0 else \
                torch.cat([self.audio_buffer[self.write_ptr:], self.audio_buffer[:self.write_ptr]])
        return win if as_tensor else win.cpu().numpy()

    def get_stats(self) -> Dict[str, Any]:
        with self._lock:
            return {"context_window": self.context_window, "buffer_size": self.buffer_size,
                    "hop_length": self.hop_length, "total_frames_added": self.total_frames_added,
                    "buffer_overruns": self.buffer_overruns, "is_full": self.is_full, "write_ptr": self.write_ptr,
                    "buffer_utilization": self.total_frames_added * self.hop_length / self.buffer_size
                    if self.total_frames_added > 0 else 0.0}


class MelSlidingWindowExtractor:
    """Frame-by-frame log-mel over a sliding 8.5 s context (reference ``:157-412``)."""

    def __init__(self, context_window: float = 8.5, update_interval: float = 0.0333, sample_rate: int = 16000,
                 n_mels: int = 80, n_fft: int = 512, hop_length: Optional[int] = None,
                 win_length: Optional[int] = None, f_min: float = 80.0, f_max: Optional[float] = None,
                 power: float = 2.0, center: bool = True, pad_mode: str = "reflect", device: str = "cuda"):
        if n_fft not in (512, 1024) or n_mels != 80 or power != 2.0 or not center or \
                (win_length is not None and not 2 <= win_length <= n_fft):
            raise NotImplementedError(
                "koemorph_b200's frontend kernel transforms n_fft = 1024 (what SimplifiedDualStreamModel passes, reference "
                "simplified_dual_stream_model.py:122-135) or 512 (this class's default in the reference, :169), "
                "win_length <= n_fft, 80 mels, power 2, center=True")
        if pad_mode not in ("reflect", "constant"):
            raise NotImplementedError(f"pad_mode {pad_mode!r}: only 'reflect' and 'constant' are implemented")
        self.context_window, self.update_interval, self.sample_rate = context_window, update_interval, sample_rate
        self.n_mels, self.n_fft, self.f_min = n_mels, n_fft, f_min
        self.f_max = f_max or sample_rate // 2
        self.power, self.center, self.pad_mode = power, center, pad_mode
        self.device = _resolve_device(device)
        target_fps = 1.0 / update_interval
        self.hop_length = hop_length or int(sample_rate / target_fps)
        self.win_length = win_length or n_fft
        self.audio_buffer = MelAudioBuffer(context_window, sample_rate, update_interval, self.device)
        self._fe = LogMelFrontend.get(self.device, sample_rate, n_fft, n_mels, f_min, self.f_max, win_length=self.win_length)
        self.mel_transform = self._fe.filterbank()         # (80, 1 + n_fft / 2) float32, as librosa.filters.mel (:224-230)
        self.current_features = None
        self._current_tensor = None
        self.last_update_time = 0
        self.features_ready = False
        self.extraction_times = deque(maxlen=100)
        self.total_extractions = 0
        self.failed_extractions = 0
        self.feature_shape = (int(context_window / update_interval), n_mels)

    # ---- device helpers -------------------------------------------------------------------------------------
    def _log_mel(self, audio: torch.Tensor, rescale: bool) -> torch.Tensor:
        """(L,) CUDA -> (1 + L // hop, 80) dB relative to the clip max, clamped at -80 [optionally (x+80)/80]."""
        n_frames = 1 + audio.numel() // self.hop_length
        db, fmax = self._fe.power(audio.reshape(1, -1), self.hop_length, n_frames, pad_mode=self.pad_mode)
        long_term, _ = self._fe.normalise(db, fmax, db_only=not rescale)
        return long_term[0]

    def _fit_frames(self, feats: torch.Tensor) -> torch.Tensor:
        expected = self.feature_shape[0]
        if feats.shape[0] > expected:
            return feats[:expected]
        if feats.shape[0] < expected:   # pad with the last frame (reference :304-307)
            return torch.cat([feats, feats[-1:].expand(expected - feats.shape[0], -1)])
        return feats

    # ---- reference API --------------------------------------------------------------------------------------
    def process_audio_frame(self, audio_frame: ArrayLike, as_tensor: bool = False, rescale: bool = False):
        """Append one hop; returns (T, 80) features (numpy float32 like the reference, or the CUDA tensor with
        ``as_tensor=True``; ``rescale=True`` applies the model's (dB + 80) / 80) or None until the ring is full."""
        if not self.audio_buffer.add_audio_frame(audio_frame):
            return None
        now = time.time()
        if now - self.last_update_time < self.update_interval * 0.3:     # reference :266-269
            return self._current_tensor if as_tensor else self.current_features
        window = self.audio_buffer.get_current_audio(as_tensor=True)
        if window is None:
            return None
        start = time.time()
        feats = self._fit_frames(self._log_mel(window, rescale))
        self._current_tensor = feats
        self.current_features = None if as_tensor else feats.cpu().numpy().astype(np.float32)
        self.last_update_time = now
        self.features_ready = True
        self.extraction_times.append(time.time() - start)
        self.total_extractions += 1
        return feats if as_tensor else self.current_features

    def process_audio_batch(self, audio: ArrayLike, as_tensor: bool = False):
        """Whole-clip log-mel in dB, (T, 80) (reference ``:326-365``)."""
        a = torch.as_tensor(audio, dtype=torch.float32).reshape(-1).to(self.device)
        feats = self._log_mel(a, rescale=False)
        return feats if as_tensor else feats.cpu().numpy().astype(np.float32)

    def get_current_features(self):
        if not self.features_ready:
            return None
        if self.current_features is None and self._current_tensor is not None:
            self.current_features = self._current_tensor.cpu().numpy().astype(np.float32)
        return self.current_features

    def reset(self):
        self.audio_buffer = MelAudioBuffer(self.context_window, self.sample_rate, self.update_interval, self.device)
        self.current_features = None
        self._current_tensor = None
        self.last_update_time = 0
        self.features_ready = False

    def get_stats(self) -> Dict[str, Any]:
        ex = {"total_extractions": self.total_extractions, "failed_extractions": self.failed_extractions,
              "success_rate": (self.total_extractions - self.failed_extractions) / max(1, self.total_extractions),
              "features_ready": self.features_ready}
        if self.extraction_times:
            ex.update({"avg_extraction_time": float(np.mean(self.extraction_times)),
                       "max_extraction_time": float(np.max(self.extraction_times)),
                       "min_extraction_time": float(np.min(self.extraction_times))})
        return {"context_window": self.context_window, "update_interval": self.update_interval,
                "feature_shape": self.feature_shape, "buffer_stats": self.audio_buffer.get_stats(),
                "extraction_stats": ex}

    @property
    def feature_dim(self) -> int:
        return self.n_mels


def create_mel_extractor(context_window: float = 8.5, update_interval: float = 0.0333, sample_rate: int = 16000,
                         n_mels: int = 80, **kwargs) -> MelSlidingWindowExtractor:
    """Factory of the reference (``:415-441``)."""
    return MelSlidingWindowExtractor(context_window=context_window, update_interval=update_interval,
                                     sample_rate=sample_rate, n_mels=n_mels, **kwargs)
