"""Hop-aligned streaming inference for many concurrent streams (BASELINE.json configs[3]).

The reference's real-time loop (``scripts/rt.py:465-519`` + ``SimplifiedDualStreamModel.process_audio_frame_realtime``,
``src/model/simplified_dual_stream_model.py:452-500``) recomputes the whole 8.5 s mel for every 33 ms hop.  A window
of ``SequentialDualStreamModel`` that starts at hop i only differs from the previous one by (SURVEY.md section 8,
note E; 30 fps geometry, hop 533 >= n_fft / 2):

* one new plain frame           F[n]    centred on sample n*hop,
* one new "window starts here"  R[n]    the same frame with everything before n*hop zeroed (frame 0 of window n),
* one "window ends here" frame  L[n+1]  centred on (n+1)*hop with everything from (n+1)*hop on zeroed,

so a step costs 3 FFTs per stream, not 257.  At 60 fps (hop 266 < n_fft / 2) two frames per window side are edge
frames: the newest complete plain frame is F[n-1], and a step adds its two window-start variants R0[n-1] (nothing before
(n-1)*hop) and R1[n-1] (nothing before (n-2)*hop) plus the last two frames of the window that ends now, L1[n] and
L0[n+1] -- 5 FFTs per hop against 513.  F and R rows go into per-stream rings of ``W`` mel rows in HBM
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
        self.model = model
        self.S = int(n_streams)
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("StreamingEngine needs the model on a CUDA device")
        self._dev_index = self.dev.index if self.dev.index is not None else torch.cuda.current_device()
        self.hop = model.hop_length
        self.W = model.mel_sequence_length
        self.half = model.n_fft // 2
        # frames within n_fft/2 of a window edge see zeros beyond it: one per side at 30 fps (hop 533 >= 512), two at 60 fps
        # (hop 266): frames 0, 1 and W-1, W of every window are edge variants (SURVEY.md section 8 note E: 5 FFTs per hop)
        self.n_edge = -(-self.half // self.hop)
        if self.n_edge > _lib.MAX_EDGE:
            raise NotImplementedError(f"hop {self.hop} needs {self.n_edge} edge frames per window side (max {_lib.MAX_EDGE})")
        self.lag = self.n_edge - 1                         # the newest complete plain frame is global frame n - lag
        self.tail_len = self.half + self.n_edge * self.hop  # samples [ (n+1)hop - tail_len, (n+1)hop )
        f32 = dict(dtype=torch.float32, device=self.dev)
        self._tails = [torch.zeros(self.S, self.tail_len, **f32) for _ in range(2)]   # ping-pong (koe_stream_push)
        self.ring_f = torch.zeros(self.S, self.W, 80, **f32)
        self.ring_r = torch.zeros(self.S, self.W, 80, **f32)
        self.fmax_f = torch.zeros(self.S, self.W, **f32)
        self.fmax_r = torch.zeros(self.S, self.W, **f32)
        self.row_l = torch.zeros(self.S, 1, 80, **f32)
        self.fmax_l = torch.zeros(self.S, 1, **f32)
        if self.n_edge == 2:
            self.ring_r2 = torch.zeros(self.S, self.W, 80, **f32)
            self.fmax_r2 = torch.zeros(self.S, self.W, **f32)
            self.row_l2 = torch.zeros(self.S, 1, 80, **f32)
            self.fmax_l2 = torch.zeros(self.S, 1, **f32)
        else:
            self.ring_r2 = self.fmax_r2 = self.row_l2 = self.fmax_l2 = None
        self.expr = torch.zeros(self.S, **f32)
        self.state = torch.zeros(self.S, 52, **f32)
        self.out = torch.zeros(self.S, 52, **f32)
        self.n = 0                                          # hops pushed so far
        self.emitted = 0
        self._fe = model._frontend(self.dev)
        self._alpha, self._alpha_key = 0.0, None
        self._args, self._args_key, self._w_tensors = None, None, []
        self.native = True                                  # one koe_stream_push per step (False: call by call from Python)

    def reset(self):
        for t in (*self._tails, self.ring_f, self.ring_r, self.fmax_f, self.fmax_r, self.state, self.ring_r2, self.fmax_r2):
            if t is not None:
                t.zero_()
        self.n = 0
        self.emitted = 0
        self._args = None                                   # the next step walks the module's weights again

    @torch.no_grad()
    def set_egemaps(self, egemaps: torch.Tensor) -> None:
        """(S, 264) eGeMAPS windows; the reference refreshes them every 300 ms, not every frame
        (src/features/opensmile_extractor.py:168)."""
        eg = self.model._check_egemaps(egemaps, self.S, self.dev)
        self._args = None                                   # (a natural point to pick up replaced weights, too)
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
        if not self.native:
            return self._step_python(x)
        # the whole step -- tail shift, the three new frames F[n], R[n], L[n+1] into the rings, the window that ends at
        # the newest sample through the core, the smoothing -- is one native call that queues five kernels
        # (host time before the first launch is latency nobody hides: the argument block is built once, a step only
        # updates what changes)
        m = self.model
        a = self._args
        # the folded weights are re-validated against the module's parameters by their in-place version counters every
        # step, and by the full walk over the module (replaced tensors: .to(), new Parameters) every 256th
        key = (m.precision, m.use_temporal_smoothing, tuple(t._version for t in self._w_tensors))
        if a is None or self._args_key != key or (self.n & 255) == 0:
            a = self._args = self._build_args()
            self._args_key = (m.precision, m.use_temporal_smoothing, tuple(t._version for t in self._w_tensors))
        if m.use_temporal_smoothing:
            # sigmoid(smoothing_alpha) is read back once per parameter version, not once per hop (a device sync)
            ver = (m.smoothing_alpha.data_ptr(), m.smoothing_alpha._version)
            if self._alpha_key != ver:
                self._alpha = float(torch.sigmoid(m.smoothing_alpha.detach().float()))
                self._alpha_key = ver
            a.alpha = self._alpha
        a.hop_audio = x.data_ptr()
        a.has_state = 1 if self.emitted > 0 else 0
        a.step = self.n
        emitted = self._emitted_flag
        if torch.cuda.current_device() == self._dev_index:
            rc = self._push(self._args_ref, self._emitted_ref, torch.cuda.current_stream(self.dev).cuda_stream)
        else:
            with torch.cuda.device(self.dev):
                rc = self._push(self._args_ref, self._emitted_ref, torch.cuda.current_stream(self.dev).cuda_stream)
        _lib.check(rc, "koe_stream_push")
        self.n += 1
        if not emitted.value:
            return None
        self.emitted += 1
        return self.out

    def _build_args(self):
        m = self.model
        self._weights = m.dual_stream_attention.kernel_weights(m._compression)   # (kept alive: the block points into it)
        att = m.dual_stream_attention
        self._w_tensors = [t for _, t in att.named_parameters()] + [t for _, t in att.named_buffers()] + \
            ([] if m._compression is None else [m._compression["weight"], m._compression["bias"]])
        a = _lib.StreamArgs()
        a.frontend, a.weights = self._fe._h, C.cast(C.pointer(self._weights.struct), C.c_void_p)
        a.n_streams, a.hop, a.window_frames, a.half_fft = self.S, self.hop, self.W, self.half
        a.tail[0], a.tail[1] = self._tails[0].data_ptr(), self._tails[1].data_ptr()
        a.ring_f, a.fmax_f = self.ring_f.data_ptr(), self.fmax_f.data_ptr()
        a.ring_r, a.fmax_r = self.ring_r.data_ptr(), self.fmax_r.data_ptr()
        a.row_l, a.fmax_l = self.row_l.data_ptr(), self.fmax_l.data_ptr()
        if self.n_edge == 2:
            a.ring_r2, a.fmax_r2 = self.ring_r2.data_ptr(), self.fmax_r2.data_ptr()
            a.row_l2, a.fmax_l2 = self.row_l2.data_ptr(), self.fmax_l2.data_ptr()
        a.expr_sigmoid, a.out = self.expr.data_ptr(), self.out.data_ptr()
        a.ema_state, a.alpha = (self.state.data_ptr(), self._alpha) if m.use_temporal_smoothing else (None, 1.0)
        a.precision = _lib.PRECISIONS[m.precision]
        self._args_ref = C.byref(a)
        self._emitted_flag = C.c_int(0)
        self._emitted_ref = C.byref(self._emitted_flag)
        self._push = _lib.load().koe_stream_push
        return a

    def _step_python(self, x: torch.Tensor) -> Optional[torch.Tensor]:
        """The same step issued call by call from Python (``native = False``): the readable statement of what
        koe_stream_push does, kept as its cross-check (tests/test_gpu_streaming.py)."""
        n, hop, W, half = self.n, self.hop, self.W, self.half
        # slide the audio tail: drop the oldest hop, append the new one
        tail = self._tails[(n + 1) & 1]
        torch.cat([self._tails[n & 1][:, hop:], x], dim=1, out=tail)
        slot = (n - self.lag) % W                           # ring slot of the newest complete frame g = n - lag
        fe = self._fe
        # frame g and the frame one hop later are consecutive frames of the tail: one launch, g into the ring slot, the
        # pair's second frame into its own row -- at 30 fps that is L (centred on the tail's end, beyond which the clip
        # reads as zeros), at 60 fps the last-but-one frame of the window that ends now
        second = (self.row_l, self.fmax_l) if self.n_edge == 1 else (self.row_l2, self.fmax_l2)
        fe.power(tail, hop, 2, sample_offset=half, out=(self.ring_f, self.fmax_f), out_row=slot, out_b=second)
        if self.n_edge == 2:                                # the frame centred on the tail's end: the window's last frame
            fe.power(tail, hop, 1, frame_offset=2, sample_offset=half, out=(self.row_l, self.fmax_l))
        # window-start variants of frame g: nothing before its centre; at 60 fps also nothing before frame g-1's centre
        fe.power(tail, hop, 1, sample_offset=half, lo_rel=0, out=(self.ring_r, self.fmax_r), out_row=slot)
        if self.n_edge == 2:
            fe.power(tail, hop, 1, sample_offset=half, lo_rel=-1, out=(self.ring_r2, self.fmax_r2), out_row=slot)
        self.n += 1
        if self.n < W:
            return None
        base = self.n - W                                    # first frame of the window that ends now
        m = self.model
        w = m.dual_stream_attention.kernel_weights(m._compression)
        lib = _lib.load()
        with torch.cuda.device(self.dev):
            st = _lib.stream_ptr(self.dev)
            bufs = [self.ring_f, self.ring_r, self.row_l] + ([self.ring_r2, self.row_l2] if self.n_edge == 2 else [])
            fmx = [self.fmax_f, self.fmax_r, self.fmax_l] + ([self.fmax_r2, self.fmax_l2] if self.n_edge == 2 else [])
            n_buf = 1 + 2 * _lib.MAX_EDGE
            pw = (C.c_void_p * n_buf)(*[t.data_ptr() for t in bufs] + [None] * (n_buf - len(bufs)))
            fm = (C.c_void_p * n_buf)(*[t.data_ptr() for t in fmx] + [None] * (n_buf - len(fmx)))
            _lib.check(lib.koe_dual_stream_ring_edges(
                C.byref(w.struct), pw, fm, self.n_edge, self.S, W, base % W, W + 1, self.expr.data_ptr(),
                self.out.data_ptr(), None, None, _lib.PRECISIONS[m.precision], st), "koe_dual_stream_ring_edges")
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
