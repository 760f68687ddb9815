"""B200 drop-in for the reference's ``src/model/simplified_dual_stream_model.py``.

``SimplifiedDualStreamModel`` keeps the reference constructor (``:28-40``), ``state_dict`` keys
(``smoothing_alpha`` + ``dual_stream_attention.*``) and ``forward`` contract (``:370-415``).
Two deliberate differences, both required by BASELINE.json's north_star:

* the eGeMAPS windows are an explicit input (``egemaps``: (B, 264) or (B, 3, 88)) instead of an
  openSMILE call on the CPU (reference ``:231-300`` -> ``src/features/opensmile_extractor.py``);
* nothing runs on the CPU: audio must be a CUDA tensor, and every stage is a hand-written sm_100a kernel
  reached through the C ABI (include/koemorph_b200.h).
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib
from ..features.mel_frontend import LogMelFrontend
from .dual_stream_attention import DualStreamCrossAttention, _package

COMPRESSION_SEED = 91011  # seed of the 264 -> 256 layer


class SimplifiedDualStreamModel(nn.Module):
    """audio (B, L) float32 @16 kHz + eGeMAPS windows (B, 264) -> 52 ARKit coefficients per clip."""

    def __init__(self, d_model: int = 256, num_heads: int = 8, num_blendshapes: int = 52, sample_rate: int = 16000,
                 target_fps: int = 30, mel_sequence_length: int = 256, emotion_config: Optional[Dict] = None,
                 mel_config: Optional[Dict] = None, device: str = "cuda", real_time_mode: bool = False):
        super().__init__()
        self.d_model, self.num_blendshapes = d_model, num_blendshapes
        self.sample_rate, self.target_fps = sample_rate, target_fps
        self.mel_sequence_length = mel_sequence_length
        self.device = device
        self.real_time_mode = real_time_mode
        self.n_mels = 80
        self.hop_length = int(sample_rate / target_fps)   # reference :54 (533 @30 fps, 266 @60 fps)
        self.n_fft = 1024
        self.emotion_config = dict(emotion_config or {})
        self.mel_config = dict(mel_config or {})
        self.f_min, self.f_max = 80.0, 8000.0             # hard-coded in the reference's librosa call (:194-195)
        # concatenated eGeMAPS (3 x 88 -> 256) is the production configuration (:99-103)
        self.emotion_dim = 256
        self.dual_stream_attention = DualStreamCrossAttention(
            d_model=d_model, num_heads=num_heads, num_mel_channels=self.n_mels,
            mel_sequence_length=mel_sequence_length, mel_temporal_frames=3, emotion_dim=self.emotion_dim,
            dropout=0.1, num_blendshapes=num_blendshapes, use_learnable_weights=True, temperature=1.0)
        self.use_temporal_smoothing = True
        self.smoothing_alpha = nn.Parameter(torch.tensor(0.8))
        self.prev_blendshapes: Optional[torch.Tensor] = None
        # 264 -> 256 compression: in the reference an unseeded nn.Linear created lazily on the extractor and kept
        # OUT of state_dict (src/features/opensmile_extractor.py:586-592); here a seeded plain tensor pair.
        gen = torch.Generator().manual_seed(COMPRESSION_SEED)
        bound = 1.0 / (264 ** 0.5)
        self._compression = {
            "weight": (torch.rand(256, 264, generator=gen) * 2 - 1) * bound,
            "bias": (torch.rand(256, generator=gen) * 2 - 1) * bound,
        }
        self._compression_dev: Dict = {}
        self._mel_extractor = None
        if real_time_mode:
            from ..features.mel_sliding_window import MelSlidingWindowExtractor
            mc = self.mel_config
            self._mel_extractor = MelSlidingWindowExtractor(
                context_window=mc.get("context_window", 8.5), update_interval=mc.get("update_interval", 0.0333),
                sample_rate=sample_rate, n_mels=self.n_mels, n_fft=mc.get("n_fft", 1024),
                hop_length=self.hop_length, f_min=mc.get("f_min", 80.0), f_max=mc.get("f_max", sample_rate // 2),
                device=device)
            self.mel_context_window = mc.get("context_window", 8.5)
            self.mel_update_interval = mc.get("update_interval", 0.0333)

    # ---- plumbing --------------------------------------------------------------------------------------
    @property
    def mel_extractor(self):
        return self._mel_extractor

    @property
    def precision(self) -> str:
        return self.dual_stream_attention.precision

    @precision.setter
    def precision(self, value: str):
        if value not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        self.dual_stream_attention.precision = value

    def set_compression_layer(self, weight: torch.Tensor, bias: torch.Tensor) -> None:
        """Install the (264 -> 256) eGeMAPS compression weights (nn.Linear layout: (256, 264), (256,))."""
        if tuple(weight.shape) != (256, 264) or tuple(bias.shape) != (256,):
            raise ValueError("compression layer must be weight (256, 264), bias (256,)")
        self._compression = {"weight": weight.detach().clone().float(), "bias": bias.detach().clone().float()}
        self._compression_dev = {}
        self.dual_stream_attention.invalidate_kernel_weights()   # a new tensor may reuse a freed address at version 0

    def _alpha(self) -> float:
        """sigmoid(smoothing_alpha) as a host float, read back once per parameter version (a device sync), not per call."""
        p = self.smoothing_alpha
        key = (p.data_ptr(), p._version, self.dual_stream_attention._generation)
        hit = getattr(self, "_alpha_cache", None)
        if hit is None or hit[0] != key:
            hit = (key, float(torch.sigmoid(p.detach().float())))
            self._alpha_cache = hit
        return hit[1]

    def _frontend(self, device) -> LogMelFrontend:
        return LogMelFrontend.get(device, self.sample_rate, self.n_fft, self.n_mels, self.f_min, self.f_max)

    def _check_audio(self, audio) -> torch.Tensor:
        audio = _lib.require_cuda(audio, "audio")
        if audio.dim() != 2:
            raise ValueError(f"audio must be (B, T), got {tuple(audio.shape)}")
        return audio

    def _check_egemaps(self, egemaps, B: int, device) -> torch.Tensor:
        if egemaps is None:
            raise RuntimeError(
                "egemaps is required: koemorph_b200 takes the three 88-D eGeMAPS windows as an input ((B, 264) or "
                "(B, 3, 88)); the reference's CPU openSMILE call (simplified_dual_stream_model.py:231-300) is out of "
                "scope and there is no CPU fallback")
        eg = _lib.require_cuda(egemaps, "egemaps")
        if eg.dim() == 3:
            eg = eg.reshape(eg.shape[0], -1)
        if eg.shape != (B, 264):
            raise ValueError(f"egemaps must be (B, 264) or (B, 3, 88) with B={B}, got {tuple(egemaps.shape)}")
        if eg.device != device:
            raise RuntimeError("audio and egemaps must be on the same device")
        return eg.contiguous()

    # ---- reference API ---------------------------------------------------------------------------------
    @torch.no_grad()
    def extract_mel_features(self, audio: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(B, L) -> long-term (B, 1 + L // hop, 80) and short-term (B, 3, 80), both (dB + 80) / 80
        (reference :166-229)."""
        audio = self._check_audio(audio)
        fe = self._frontend(audio.device)
        n_frames = 1 + audio.shape[1] // self.hop_length
        power, fmax = fe.power(audio, self.hop_length, n_frames)
        return fe.normalise(power, fmax)

    @torch.no_grad()
    def extract_emotion_features(self, audio: torch.Tensor = None, egemaps: torch.Tensor = None):
        """(B, 264) eGeMAPS windows -> compressed (B, 256) features + metadata (reference :231-300,
        src/features/opensmile_extractor.py:583-604)."""
        if egemaps is None:
            self._check_egemaps(None, 0, None)
        eg = _lib.require_cuda(egemaps, "egemaps")
        eg = eg.reshape(eg.shape[0], 264).contiguous()
        out = torch.empty(eg.shape[0], 256, dtype=torch.float32, device=eg.device)
        with torch.cuda.device(eg.device):
            _lib.check(_lib.load().koe_affine_rows(eg.data_ptr(), eg.shape[0], 264, self._compression_t(eg.device).data_ptr(),
                                                   self._compression_on(eg.device)["bias"].data_ptr(), 256, out.data_ptr(),
                                                   _lib.stream_ptr(eg.device)), "koe_affine_rows")
        return out, {"backend_used": "egemaps_input", "processing_time": 0.0}

    def _compression_on(self, device) -> Dict[str, torch.Tensor]:
        key = str(device)
        if key not in self._compression_dev:
            self._compression_dev = {key: {"weight": self._compression["weight"].to(device),
                                           "bias": self._compression["bias"].to(device),
                                           "weight_t": self._compression["weight"].t().contiguous().to(device)}}
        return self._compression_dev[key]

    def _compression_t(self, device) -> torch.Tensor:
        return self._compression_on(device)["weight_t"]

    def align_features(self, mel_features, emotion_features):
        """No-op for concatenated eGeMAPS (reference :317-322)."""
        return mel_features, emotion_features

    @torch.no_grad()
    def apply_temporal_smoothing(self, blendshapes: torch.Tensor) -> torch.Tensor:
        """Stateful EMA across calls: first call (or batch-size change) passes through (reference :341-368)."""
        if not self.use_temporal_smoothing:
            return blendshapes
        B = blendshapes.shape[0]
        if self.prev_blendshapes is None or self.prev_blendshapes.shape[0] != B:
            self.prev_blendshapes = blendshapes.detach().clone()
            return blendshapes
        x = _lib.require_cuda(blendshapes, "blendshapes").clone()
        # the state may be a view of an earlier result (SequentialDualStreamModel.forward leaves out[:, -1] here): the scan
        # updates its state in place, so it gets a private contiguous copy
        state = self.prev_blendshapes.clone(memory_format=torch.contiguous_format)
        alpha = self._alpha()
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().koe_ema_scan(x.data_ptr(), B, 1, alpha, state.data_ptr(), 1,
                                                _lib.stream_ptr(x.device)), "koe_ema_scan")
        self.prev_blendshapes = state
        return x

    def _core_windows(self, power_list, fmax_list, n_edge, B, n_frames, n_out, stride, frames_per_window, eg,
                      return_attention, out=None):
        """koe_emotion_stream + koe_dual_stream_windows on prepared mel-power buffers; ``out`` may be a preallocated
        contiguous (B, n_out, 52) float32 tensor (a corpus sweep writes every chunk's frames where they will be gathered)."""
        lib = _lib.load()
        dev = eg.device
        w = self.dual_stream_attention.kernel_weights(self._compression)
        expr = torch.empty(B, dtype=torch.float32, device=dev)
        if out is None:
            out = torch.empty(B, n_out, 52, dtype=torch.float32, device=dev)
        elif out.shape != (B, n_out, 52) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
            raise ValueError(f"out must be a contiguous float32 ({B}, {n_out}, 52) tensor on {dev}")
        sig = torch.empty(B, n_out, 52, dtype=torch.float32, device=dev) if return_attention else None
        attn = torch.empty(B, n_out, 28, 80, dtype=torch.float32, device=dev) if return_attention else None
        n_buf = 1 + 2 * _lib.MAX_EDGE
        pw = (C.c_void_p * n_buf)(*[t.data_ptr() for t in power_list] + [None] * (n_buf - len(power_list)))
        fm = (C.c_void_p * n_buf)(*[t.data_ptr() for t in fmax_list] + [None] * (n_buf - len(fmax_list)))
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            _lib.check(lib.koe_emotion_stream(C.byref(w.struct), eg.data_ptr(), B, expr.data_ptr(), st),
                       "koe_emotion_stream")
            _lib.check(lib.koe_dual_stream_windows(
                C.byref(w.struct), pw, fm, n_edge, B, n_frames, n_out, stride, frames_per_window, expr.data_ptr(),
                out.data_ptr(), sig.data_ptr() if sig is not None else None,
                attn.data_ptr() if attn is not None else None, _lib.PRECISIONS[self.precision], st),
                "koe_dual_stream_windows")
        return out, sig, attn

    def _forward_windows(self, audio, eg, n_frames, frames_per_window, stride, n_out, n_edge, smooth, return_attention,
                         out=None):
        """koe_forward_windows: frontend (+ edge variants), emotion stream, core and EMA scan as ONE native call.
        Returns (out (B, n_out, 52), sigmoid or None, attention or None); ``out`` may be preallocated."""
        lib = _lib.load()
        dev = audio.device
        B, L = audio.shape
        w = self.dual_stream_attention.kernel_weights(self._compression)
        fe = self._frontend(dev)
        if out is None:
            out = torch.empty(B, n_out, 52, dtype=torch.float32, device=dev)
        elif out.shape != (B, n_out, 52) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
            raise ValueError(f"out must be a contiguous float32 ({B}, {n_out}, 52) tensor on {dev}")
        if B == 0:   # nothing to launch (and no device pointers to hand over)
            empty = (lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)) if return_attention else (lambda *shape: None)
            return out, empty(0, n_out, 52), empty(0, n_out, 28, 80)
        if L == 0:   # a clip without samples is all padding: the kernels read nothing, but the ABI wants a pointer
            audio = torch.zeros(B, 1, dtype=torch.float32, device=dev)
        # one workspace allocation per call: mel power (dB) of the global frames and of the edge variants, their per-frame
        # maxima, the emotion stream's per-clip value
        n_main, n_var = B * n_frames, B * n_out
        floats = n_main * 81 + 2 * n_edge * n_var * 81 + B
        ws = torch.empty(floats + 4, dtype=torch.float32, device=dev)
        a = _lib.ForwardArgs()
        a.frontend, a.weights = fe._h, C.addressof(w.struct)
        a.audio, a.audio_stride = audio.data_ptr(), audio.stride(0)
        a.n_clips, a.n_samples, a.hop = B, L, self.hop_length
        a.n_frames, a.frames_per_window, a.stride_frames, a.n_out, a.n_edge = n_frames, frames_per_window, stride, n_out, n_edge
        a.egemaps = eg.data_ptr()
        base, off = ws.data_ptr(), 0
        a.power[0] = base
        off += n_main * 80
        for j in range(1, 1 + 2 * n_edge):
            a.power[j] = base + 4 * off
            off += n_var * 80
        a.frame_max[0] = base + 4 * off
        off += n_main
        for j in range(1, 1 + 2 * n_edge):
            a.frame_max[j] = base + 4 * off
            off += n_var
        a.expr_sigmoid = base + 4 * off
        sig = torch.empty(B, n_out, 52, dtype=torch.float32, device=dev) if return_attention else None
        attn = torch.empty(B, n_out, 28, 80, dtype=torch.float32, device=dev) if return_attention else None
        a.out = out.data_ptr()
        a.sigmoid_out = sig.data_ptr() if sig is not None else None
        a.attn_out = attn.data_ptr() if attn is not None else None
        a.alpha = self._alpha() if smooth else 0.0
        a.smooth = 1 if smooth else 0
        a.precision = _lib.PRECISIONS[self.precision]
        with torch.cuda.device(dev):
            _lib.check(lib.koe_forward_windows(C.byref(a), _lib.stream_ptr(dev)), "koe_forward_windows")
        return out, sig, attn

    def _single_frames(self, audio, eg, return_attention, out=None):
        """The stateless part of forward(): (B, 52) pre-smoothing blendshapes (+ sigmoid, attention)."""
        B, L = audio.shape
        T = 1 + L // self.hop_length
        o, sig, attn = self._forward_windows(audio, eg, T, T, 1, 1, 0, False, return_attention,
                                             out=None if out is None else out.view(B, 1, 52))
        return o[:, 0], None if sig is None else sig[:, 0], None if attn is None else attn[:, 0]

    def _forward_frames(self, audio, eg, return_attention, out=None):
        """What HostPipeline runs per chunk: the stateless part of THIS class's forward (overridden by the sequence model)."""
        return self._single_frames(audio, eg, return_attention, out=out)

    @torch.no_grad()
    def forward(self, audio: torch.Tensor, return_attention: bool = False,
                egemaps: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """One 52-D frame per clip from the whole clip's mel (reference :370-415)."""
        audio = self._check_audio(audio)
        B, L = audio.shape
        eg = self._check_egemaps(egemaps, B, audio.device)
        out, sig, attn = self._single_frames(audio, eg, return_attention)
        res = _package(out, sig, attn, return_attention)
        res["blendshapes"] = self.apply_temporal_smoothing(res["blendshapes"])
        return res

    def reset_temporal_state(self):
        """Reset temporal state for a new sequence (reference :417-419)."""
        self.prev_blendshapes = None

    def get_model_info(self) -> Dict[str, Any]:
        """Reference :421-450 (emotion extractor statistics are replaced by the input contract)."""
        info = {
            "model_type": type(self).__name__, "d_model": self.d_model,
            "num_heads": self.dual_stream_attention.num_heads, "num_blendshapes": self.num_blendshapes,
            "emotion_backend": "egemaps_input", "emotion_fallback_level": 1,
            "mel_sequence_length": self.mel_sequence_length, "n_mels": self.n_mels, "emotion_dim": self.emotion_dim,
            "total_parameters": sum(p.numel() for p in self.parameters() if p.requires_grad),
            "emotion_extraction_stats": {}, "real_time_mode": self.real_time_mode, "precision": self.precision,
        }
        if self.real_time_mode and self._mel_extractor is not None:
            info.update({"mel_context_window": self.mel_context_window,
                         "mel_update_interval": self.mel_update_interval,
                         "mel_extraction_stats": self._mel_extractor.get_stats()})
        return info

    # ---- realtime helpers (reference :452-522) ------------------------------------------------------------
    @torch.no_grad()
    def process_audio_frame_realtime(self, audio_frame, return_attention: bool = False,
                                     egemaps: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """One hop of audio in, (52,) out once the 8.5 s context is full (reference :452-500).

        The reference calls the core without ``mel_temporal_features`` here and raises TypeError
        (SURVEY.md section 3.5); this implementation feeds the last three context frames, which is what
        ``extract_mel_features`` defines as the short-term detail."""
        if not self.real_time_mode:
            raise RuntimeError("Model not in real-time mode. Use forward() for batch processing.")
        if self._mel_extractor is None:
            raise RuntimeError("Mel extractor not initialized for real-time mode.")
        feats = self._mel_extractor.process_audio_frame(audio_frame, as_tensor=True, rescale=True)
        if feats is None:
            return None
        dev = feats.device
        eg = self._check_egemaps(egemaps, 1, dev)
        emo, _ = self.extract_emotion_features(egemaps=eg)
        mel = feats.unsqueeze(0)
        out = self.dual_stream_attention(mel, mel[:, -3:].contiguous(), emo, return_attention=return_attention)
        return self.apply_temporal_smoothing(out["blendshapes"]).squeeze(0)

    def reset_realtime_state(self):
        if self.real_time_mode and self._mel_extractor is not None:
            self._mel_extractor.reset()
        self.reset_temporal_state()

    def get_realtime_stats(self) -> Dict[str, Any]:
        if not self.real_time_mode:
            return {"error": "Not in real-time mode"}
        stats: Dict[str, Any] = {}
        if self._mel_extractor is not None:
            stats["mel_stats"] = self._mel_extractor.get_stats()
        stats["emotion_stats"] = {}
        return stats
