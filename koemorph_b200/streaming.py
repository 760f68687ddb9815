"""Hop-aligned streaming inference for many concurrent streams (BASELINE.json configs[3]).

The reference's real-time loop (``scripts/rt.py:465-519`` + ``SimplifiedDualStreamModel.process_audio_frame_realtime``,
``src/model/simplified_dual_stream_model.py:452-500``) recomputes the whole 8.5 s mel for every 33 ms hop.  A window
of ``SequentialDualStreamModel`` that starts at hop i only differs from the previous one by (SURVEY.md section 8,
note E; 30 fps geometry, hop 533 >= n_fft / 2):

* one new plain frame           F[n]    centred on sample n*hop,
* one new "window starts here"  R[n]    the same frame with everything before n*hop zeroed (frame 0 of window n),
* one "window ends here" frame  L[n+1]  centred on (n+1)*hop with everything from (n+1)*hop on zeroed,

so a step costs 3 FFTs per stream, not 257.  F and R rows go into per-stream rings of ``W`` mel rows in HBM
(164 KB per stream), the only audio kept is the last ``n_fft/2 + hop`` samples, and the core kernel addresses the
rings directly (``koe_dual_stream_ring``).  After the 256th hop every step emits the frame that
``SequentialDualStreamModel.forward`` would produce for the window ending at the newest sample, EMA included.
All streams advance in lockstep (one hop per step each).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib


class StreamingEngine:
    def __init__(self, model, n_streams: int):
        if model.hop_length < model.n_fft // 2:
            raise NotImplementedError("StreamingEngine implements the 30 fps geometry (hop >= n_fft/2, one edge frame "
                                      "per window side)")
        self.model = model
        self.S = int(n_streams)
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("StreamingEngine needs the model on a CUDA device")
        self.hop = model.hop_length
        self.W = model.mel_sequence_length
        self.half = model.n_fft // 2
        self.tail_len = self.half + self.hop               # samples [ (n+1)hop - tail_len, (n+1)hop )
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.tail = torch.zeros(self.S, self.tail_len, **f32)
        self.ring_f = torch.zeros(self.S, self.W, 80, **f32)
        self.ring_r = torch.zeros(self.S, self.W, 80, **f32)
        self.fmax_f = torch.zeros(self.S, self.W, **f32)
        self.fmax_r = torch.zeros(self.S, self.W, **f32)
        self.row_l = torch.zeros(self.S, 1, 80, **f32)
        self.fmax_l = torch.zeros(self.S, 1, **f32)
        self.expr = torch.zeros(self.S, **f32)
        self.state = torch.zeros(self.S, 52, **f32)
        self.out = torch.zeros(self.S, 52, **f32)
        self.n = 0                                          # hops pushed so far
        self.emitted = 0
        self._fe = model._frontend(self.dev)
        self._alpha, self._alpha_key = 0.0, None

    def reset(self):
        for t in (self.tail, self.ring_f, self.ring_r, self.fmax_f, self.fmax_r, self.state):
            t.zero_()
        self.n = 0
        self.emitted = 0

    @torch.no_grad()
    def set_egemaps(self, egemaps: torch.Tensor) -> None:
        """(S, 264) eGeMAPS windows; the reference refreshes them every 300 ms, not every frame
        (src/features/opensmile_extractor.py:168)."""
        eg = self.model._check_egemaps(egemaps, self.S, self.dev)
        w = self.model.dual_stream_attention.kernel_weights(self.model._compression)
        with torch.cuda.device(self.dev):
            _lib.check(_lib.load().koe_emotion_stream(C.byref(w.struct), eg.data_ptr(), self.S, self.expr.data_ptr(),
                                                      _lib.stream_ptr(self.dev)), "koe_emotion_stream")

    @torch.no_grad()
    def step(self, hop_audio: torch.Tensor) -> Optional[torch.Tensor]:
        """hop_audio (S, hop) float32 CUDA: the next hop of every stream.  Returns (S, 52) smoothed blendshapes once
        the 8.5 s context is full (from the W-th hop on), else None.  The returned tensor is reused by the next step."""
        x = _lib.require_cuda(hop_audio, "hop_audio")
        if x.shape != (self.S, self.hop):
            raise ValueError(f"hop_audio must be ({self.S}, {self.hop}), got {tuple(x.shape)}")
        n, hop, W, half = self.n, self.hop, self.W, self.half
        # slide the audio tail: keep the last n_fft/2 samples, append the hop
        self.tail = torch.cat([self.tail[:, hop:], x], dim=1)
        slot = n % W
        fe = self._fe
        # F[n]: centred on local sample `half`; before the first hop the tail is zeros = librosa's zero padding
        fe.power(self.tail, hop, 1, sample_offset=half, out=(self.ring_f, self.fmax_f), out_row=slot)
        # R[n]: same frame, nothing before its centre
        fe.power(self.tail, hop, 1, sample_offset=half, lo_rel=0, out=(self.ring_r, self.fmax_r), out_row=slot)
        # L[n+1]: centred on the end of the tail, nothing from its centre on
        fe.power(self.tail, hop, 1, sample_offset=self.tail_len, hi_rel=0, out=(self.row_l, self.fmax_l))
        self.n += 1
        if self.n < W:
            return None
        base = self.n - W                                    # first frame of the window that ends now
        m = self.model
        w = m.dual_stream_attention.kernel_weights(m._compression)
        lib = _lib.load()
        with torch.cuda.device(self.dev):
            st = _lib.stream_ptr(self.dev)
            _lib.check(lib.koe_dual_stream_ring(
                C.byref(w.struct), self.ring_f.data_ptr(), self.fmax_f.data_ptr(), self.ring_r.data_ptr(),
                self.fmax_r.data_ptr(), self.row_l.data_ptr(), self.fmax_l.data_ptr(), self.S, W, base % W, W + 1,
                self.expr.data_ptr(), self.out.data_ptr(), None, None, _lib.PRECISIONS[m.precision], st),
                "koe_dual_stream_ring")
            if m.use_temporal_smoothing:
                # sigmoid(smoothing_alpha) is read back once per parameter version, not once per hop (a device sync)
                ver = (m.smoothing_alpha.data_ptr(), m.smoothing_alpha._version)
                if self._alpha_key != ver:
                    self._alpha = float(torch.sigmoid(m.smoothing_alpha.detach().float()))
                    self._alpha_key = ver
                alpha = self._alpha
                _lib.check(lib.koe_ema_scan(self.out.data_ptr(), self.S, 1, alpha, self.state.data_ptr(),
                                            1 if self.emitted > 0 else 0, st), "koe_ema_scan")
        self.emitted += 1
        return self.out
