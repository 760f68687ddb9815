"""Device log-mel frontend: thin host wrapper over koe_frontend_* / koe_logmel_* (include/koemorph_b200.h).

Replaces the per-clip librosa loop of the reference
(src/model/simplified_dual_stream_model.py:184-229: melspectrogram(n_fft=1024, hop, n_mels=80,
fmin=80, fmax=8000) -> power_to_db(ref=np.max) -> (x + 80) / 80).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .. import _lib

N_FFT = 1024
N_MELS = 80


def pcm16_to_float(pcm: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int16 PCM on the device -> float32 audio, sample / 32768 (what sf.read(dtype="float32") returns for a 16-bit WAV,
    src/data/io.py:71); `out` may be a preallocated float32 tensor of the same shape."""
    pcm = _lib.require_cuda(pcm, "pcm", torch.int16)
    if out is None:
        out = torch.empty(pcm.shape, dtype=torch.float32, device=pcm.device)
    else:
        out = _lib.require_cuda(out, "out")
        if out.shape != pcm.shape:
            raise ValueError(f"out has shape {tuple(out.shape)}, pcm {tuple(pcm.shape)}")
    with torch.cuda.device(pcm.device):
        _lib.check(_lib.load().koe_pcm16_to_float(pcm.data_ptr(), pcm.numel(), out.data_ptr(),
                                                  _lib.stream_ptr(pcm.device)), "koe_pcm16_to_float")
    return out


class LogMelFrontend:
    """One handle (Hann window, FFT twiddles, sparse Slaney filterbank) per CUDA device."""

    _cache: Dict[Tuple, "LogMelFrontend"] = {}

    def __init__(self, device, sample_rate: int = 16000, n_fft: int = N_FFT, n_mels: int = N_MELS,
                 f_min: float = 80.0, f_max: float = 8000.0, mel_scale: str = "slaney", window_normalized: bool = False,
                 log_mode: str = "db", log_eps: float = 0.0, win_length: int = 0):
        """``mel_scale`` "slaney" (librosa: Slaney scale, slaney norm) or "htk" (torchaudio default: HTK scale, no norm);
        ``log_mode`` "db" = 10 log10(max(p, 1e-10)) or "ln" = ln(p + log_eps); ``n_fft`` 1024 or 512."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(f"LogMelFrontend needs a CUDA device, got {device}; there is no CPU path")
        if mel_scale not in ("slaney", "htk") or log_mode not in ("db", "ln"):
            raise ValueError("mel_scale must be 'slaney' or 'htk', log_mode 'db' or 'ln'")
        self.device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        self.sample_rate, self.n_fft, self.n_mels, self.f_min, self.f_max = sample_rate, n_fft, n_mels, f_min, f_max
        self._lib = _lib.load()
        cfg = _lib.FrontendConfig(self.device.index, sample_rate, n_fft, n_mels, float(f_min), float(f_max),
                                  1 if mel_scale == "htk" else 0, 0 if mel_scale == "htk" else 1,
                                  int(bool(window_normalized)), 1 if log_mode == "ln" else 0, float(log_eps),
                                  int(win_length or 0))
        h = C.c_void_p()
        _lib.check(self._lib.koe_frontend_create_ex(C.byref(cfg), C.byref(h)), "koe_frontend_create_ex")
        self._h = h

    @classmethod
    def get(cls, device, sample_rate=16000, n_fft=N_FFT, n_mels=N_MELS, f_min=80.0, f_max=8000.0, mel_scale="slaney",
            window_normalized=False, log_mode="db", log_eps=0.0, win_length=0):
        device = torch.device(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        win_length = 0 if win_length in (None, n_fft) else int(win_length)
        key = (idx, sample_rate, n_fft, n_mels, float(f_min), float(f_max), mel_scale, bool(window_normalized), log_mode,
               float(log_eps), win_length)
        if key not in cls._cache:
            cls._cache[key] = cls(torch.device("cuda", idx), sample_rate, n_fft, n_mels, f_min, f_max, mel_scale,
                                  window_normalized, log_mode, log_eps, win_length)
        return cls._cache[key]

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.koe_frontend_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def uses_unrolled_bank(self) -> bool:
        """True when this bank has the structure of the path's default (sr 16000, 80 mels, 80..8000 Hz) and the kernel
        runs the filterbank phase unrolled from it; False: generic looped kernel (same results, slower)."""
        return bool(self._lib.koe_frontend_uses_unrolled_bank(self._h))

    def filterbank(self) -> np.ndarray:
        """Dense (n_mels, 1 + n_fft // 2) float32 filterbank (the library keeps it on the bins of its 1024-point
        transform; a 512-point bank sits on the even bins)."""
        fb = np.empty((self.n_mels, 1 + N_FFT // 2), np.float32)
        _lib.check(self._lib.koe_frontend_filterbank_host(self._h, fb.ctypes.data_as(C.c_void_p)))
        return np.ascontiguousarray(fb[:, ::N_FFT // self.n_fft])

    def power(self, audio: torch.Tensor, hop: int, n_frames: int, frame_offset: int = 0, frame_step: int = 1,
              lo_rel: Optional[int] = None, hi_rel: Optional[int] = None,
              out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, sample_offset: int = 0,
              pad_mode: str = "constant", out_row: int = 0,
              out_b: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """audio (B, L) float32 CUDA -> mel power in dB, 10*log10(max(p, 1e-10)), (B, n_frames, 80) and its
        per-frame max (B, n_frames).

        With ``out=(power, fmax)`` of shapes (B, R, 80) / (B, R) the n_frames rows are written at rows
        ``out_row ...`` of every clip's block (ring buffers of the streaming paths).  With ``out_b=(power_b, fmax_b)`` of
        shapes (B, Rb, 80) / (B, Rb) the odd frames 2j + 1 go to row j there instead (koe_logmel_args.power_b)."""
        audio = _lib.require_cuda(audio, "audio")
        if audio.dim() != 2:
            raise ValueError(f"audio must be (B, L), got {tuple(audio.shape)}")
        if pad_mode not in ("constant", "reflect"):
            raise ValueError(f"pad_mode must be 'constant' or 'reflect', got {pad_mode!r}")
        B, L = audio.shape
        if L == 0 and B > 0:   # a clip without samples is all padding: nothing is read, but the ABI wants a pointer
            audio = torch.zeros(B, 1, dtype=torch.float32, device=audio.device)
        if out is None:
            power = torch.empty((B, n_frames, self.n_mels), dtype=torch.float32, device=audio.device)
            fmax = torch.empty((B, n_frames), dtype=torch.float32, device=audio.device)
        else:
            power, fmax = out
            rows = n_frames if out_b is None else 2 * ((n_frames - 1) // 2) + 1   # with out_b only the even frames land here
            if power.shape[0] != B or power.shape[2] != self.n_mels or out_row + rows > power.shape[1] or \
                    fmax.shape[:2] != power.shape[:2] or not power.is_contiguous() or not fmax.is_contiguous():
                raise ValueError("out must be contiguous (B, R, 80) / (B, R) tensors with out_row + n_frames <= R")
        a = _lib.LogmelArgs()
        a.audio, a.audio_stride = audio.data_ptr(), audio.stride(0)
        a.n_clips, a.n_samples, a.hop, a.n_frames = B, L, hop, n_frames
        a.frame_offset, a.frame_step, a.sample_offset = frame_offset, frame_step, sample_offset
        a.lo_rel_hops = _lib.KOE_NO_EDGE if lo_rel is None else lo_rel
        a.hi_rel_hops = _lib.KOE_NO_EDGE if hi_rel is None else hi_rel
        a.pad_mode = 1 if pad_mode == "reflect" else 0
        a.power = power.data_ptr() + out_row * self.n_mels * 4
        a.power_clip_stride = power.stride(0)
        a.frame_max = fmax.data_ptr() + out_row * 4
        a.frame_max_clip_stride = fmax.stride(0)
        if out_b is not None:
            pb, fb = out_b
            if pb.shape[0] != B or pb.shape[2] != self.n_mels or pb.shape[1] < n_frames // 2 or \
                    fb.shape[:2] != pb.shape[:2] or not pb.is_contiguous() or not fb.is_contiguous():
                raise ValueError("out_b must be contiguous (B, Rb, 80) / (B, Rb) tensors with n_frames // 2 <= Rb")
            a.power_b, a.power_b_clip_stride = pb.data_ptr(), pb.stride(0)
            a.frame_max_b, a.frame_max_b_clip_stride = fb.data_ptr(), fb.stride(0)
        with torch.cuda.device(audio.device):
            _lib.check(self._lib.koe_logmel_power_ex(self._h, C.byref(a), _lib.stream_ptr(audio.device)),
                       "koe_logmel_power")
        return power, fmax

    def normalise(self, power: torch.Tensor, fmax: torch.Tensor, db_only: bool = False):
        """power_to_db(ref=clip max, top_db=80) [+ (x+80)/80] -> long-term (B, T, 80), short-term (B, 3, 80)."""
        B, T, M = power.shape
        long_term = torch.empty_like(power)
        short_term = torch.empty((B, 3, M), dtype=torch.float32, device=power.device)
        with torch.cuda.device(power.device):
            _lib.check(self._lib.koe_logmel_normalise(power.data_ptr(), fmax.data_ptr(), B, T, int(db_only),
                                                      long_term.data_ptr(), short_term.data_ptr(),
                                                      _lib.stream_ptr(power.device)), "koe_logmel_normalise")
        return long_term, short_term
