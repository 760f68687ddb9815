"""CPU oracle for the KoeMorph audio -> ARKit-blendshape inference path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``koemorph_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker
or the timed CPU baseline -- never as the product path.

What it restates (reference paths are relative to /root/reference):

* log-mel frontend -- ``src/model/simplified_dual_stream_model.py:166-229``.
  The arithmetic lives in **librosa** (``pyproject.toml:18`` pins only
  ``librosa>=0.10.0``; librosa is NOT vendored in the reference and NOT
  installed in this image).  The functions below restate the published
  librosa >= 0.10 algorithm (``core/spectrum.py::stft/_spectrogram/power_to_db``,
  ``feature/spectral.py::melspectrogram``, ``filters.py::mel``,
  ``core/convert.py::hz_to_mel/mel_to_hz/mel_frequencies``).
* dual-stream attention core -- ``src/model/dual_stream_attention.py:162-280``.
* 264 -> 256 eGeMAPS compression -- ``src/features/opensmile_extractor.py:583-604``.
* learnable-alpha EMA smoothing -- ``simplified_dual_stream_model.py:341-368``.
* single-frame forward -- ``simplified_dual_stream_model.py:370-415``.
* sliding-window sequence forward -- ``src/model/sequential_dual_stream_model.py:63-167``.
* streaming mel -- ``src/features/mel_sliding_window.py:252-324``.

PARITY PINNING STATUS
---------------------
* Attention core, smoothing, sequence driver: **pinned** -- the unmodified
  reference modules are executed in the build container by
  ``oracle/run_reference.py`` and their outputs are committed under
  ``tests/golden/`` (generator: ``tests/golden/make_golden.py``); this oracle is
  checked against them in ``tests/test_oracle.py``.
* librosa boundary (STFT / Slaney mel / power_to_db): **parity unpinned** --
  the reference holds no test, fixture or golden vector for this path
   and librosa itself is absent, so the restatement is
  cross-checked only against independent implementations
  (``torchaudio.functional.melscale_fbanks`` and ``torch.stft``), not against
  librosa output.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import scipy.signal
import torch

# ---------------------------------------------------------------------------
# Constants of the path (reference: simplified_dual_stream_model.py:51-55,
# dual_stream_attention.py:14-45)
# ---------------------------------------------------------------------------
SAMPLE_RATE = 16000
N_FFT = 1024
N_MELS = 80
F_MIN = 80.0
F_MAX = 8000.0
N_BLENDSHAPES = 52
# dual_stream_attention.py:44-45 evaluates to mouth = 14..40 + 51, expression = rest
MOUTH_INDICES = list(range(14, 41)) + [51]
EXPRESSION_INDICES = [i for i in range(N_BLENDSHAPES) if i not in MOUTH_INDICES]


def hop_length_for(fps: int, sample_rate: int = SAMPLE_RATE) -> int:
    """simplified_dual_stream_model.py:54 -- int(sample_rate / target_fps)."""
    return int(sample_rate / fps)


# ---------------------------------------------------------------------------
# librosa restatement (published algorithm, librosa >= 0.10)
# ---------------------------------------------------------------------------
def hz_to_mel(freq):
    """librosa.core.convert.hz_to_mel(htk=False): Slaney scale."""
    freq = np.asanyarray(freq, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = freq / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if freq.ndim:
        big = freq >= min_log_hz
        mels[big] = min_log_mel + np.log(freq[big] / min_log_hz) / logstep
    elif freq >= min_log_hz:
        mels = min_log_mel + np.log(freq / min_log_hz) / logstep
    return mels


def mel_to_hz(mels):
    """librosa.core.convert.mel_to_hz(htk=False)."""
    mels = np.asanyarray(mels, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        big = mels >= min_log_mel
        freqs[big] = min_log_hz * np.exp(logstep * (mels[big] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def mel_filterbank(sr: int = SAMPLE_RATE, n_fft: int = N_FFT, n_mels: int = N_MELS,
                   fmin: float = F_MIN, fmax: float = F_MAX) -> np.ndarray:
    """librosa.filters.mel(htk=False, norm='slaney', dtype=float32) -> (n_mels, 1+n_fft//2)."""
    n_bins = 1 + n_fft // 2
    weights = np.zeros((n_mels, n_bins), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


_FB_CACHE: Dict[Tuple, np.ndarray] = {}


def _fb(sr, n_fft, n_mels, fmin, fmax):
    key = (sr, n_fft, n_mels, float(fmin), float(fmax))
    if key not in _FB_CACHE:
        _FB_CACHE[key] = mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    return _FB_CACHE[key]


def hann_window(n_fft: int = N_FFT) -> np.ndarray:
    """librosa.filters.get_window('hann', n, fftbins=True) -> scipy periodic Hann, float64."""
    return scipy.signal.get_window("hann", n_fft, fftbins=True)


def stft(y: np.ndarray, n_fft: int = N_FFT, hop_length: int = 533, center: bool = True,
         pad_mode: str = "constant", exact: bool = False) -> np.ndarray:
    """librosa.stft restated: centre padding n_fft//2 each side, frames at k*hop,
    1 + len(y)//hop frames; window (float64) * frames -> numpy rfft in float64 ->
    stored as complex64 for float32 input (``exact=True`` keeps complex128)."""
    y = np.asarray(y)
    win = hann_window(n_fft)
    if center:
        y = np.pad(y, n_fft // 2, mode=pad_mode)
    if len(y) < n_fft:
        raise ValueError("input too short for one frame")
    n_frames = 1 + (len(y) - n_fft) // hop_length
    frames = np.lib.stride_tricks.as_strided(
        y, shape=(n_fft, n_frames), strides=(y.strides[0], y.strides[0] * hop_length), writeable=False)
    spec = np.fft.rfft(win[:, None] * frames, axis=0)
    if exact:
        return spec
    return spec.astype(np.complex64 if y.dtype == np.float32 else np.complex128)


def melspectrogram(y: np.ndarray, sr: int = SAMPLE_RATE, n_fft: int = N_FFT, hop_length: int = 533,
                   n_mels: int = N_MELS, fmin: float = F_MIN, fmax: float = F_MAX, power: float = 2.0,
                   center: bool = True, pad_mode: str = "constant", exact: bool = False) -> np.ndarray:
    """librosa.feature.melspectrogram -> (n_mels, T) mel *power* (float32 unless exact)."""
    spec = stft(y, n_fft=n_fft, hop_length=hop_length, center=center, pad_mode=pad_mode, exact=exact)
    S = np.abs(spec) ** power
    fb = _fb(sr, n_fft, n_mels, fmin, fmax)
    if exact:
        return fb.astype(np.float64) @ S
    return np.einsum("ft,mf->mt", S, fb, optimize=True)


def power_to_db(S: np.ndarray, amin: float = 1e-10, top_db: Optional[float] = 80.0) -> np.ndarray:
    """librosa.power_to_db(S, ref=np.max)."""
    magnitude = np.abs(S)
    ref_value = np.abs(np.max(magnitude))
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def extract_mel_features(audio: np.ndarray, fps: int = 30, sample_rate: int = SAMPLE_RATE,
                         exact: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """simplified_dual_stream_model.py:166-229 for equal-length clips.

    audio (B, L) float32 -> long-term (B, T, 80), short-term (B, 3, 80) in the
    ``(dB + 80) / 80`` normalisation; per-clip loop exactly like the reference."""
    hop = hop_length_for(fps, sample_rate)
    long_term, short_term = [], []
    for clip in np.asarray(audio):
        mel = melspectrogram(clip, sr=sample_rate, n_fft=N_FFT, hop_length=hop, n_mels=N_MELS,
                             fmin=80, fmax=8000, exact=exact)
        mel = (power_to_db(mel) + 80) / 80
        mel_t = mel.T
        long_term.append(mel_t)
        if mel_t.shape[0] >= 3:
            detail = mel_t[-3:]
        else:
            detail = np.zeros((3, N_MELS))
            detail[:mel_t.shape[0]] = mel_t
        short_term.append(detail)
    out_dtype = np.float64 if exact else np.float32
    return np.stack(long_term).astype(out_dtype), np.stack(short_term).astype(out_dtype)


# ---------------------------------------------------------------------------
# Deterministic synthetic weights (state_dict layout of the reference,
# SURVEY.md section 8 a-W).  numpy PCG64 so every box regenerates identical bytes.
# ---------------------------------------------------------------------------
def make_weights(seed: int = 1234, fps: int = 30, d_model: int = 256, style: str = "init") -> Dict[str, np.ndarray]:
    """Random weights keyed like ``SimplifiedDualStreamModel.state_dict()``.

    style="init":   distributions of the reference's constructors
                    (dual_stream_attention.py:99-160) but with non-zero biases and
                    non-trivial LayerNorm affine so every term is exercised.
    style="stress": larger queries / stream weights so the 28x80 softmax is peaky
                    and errors in the score path are visible in the output.
    Also returns the (out-of-state_dict) 264->256 compression layer as
    ``compression.weight`` / ``compression.bias`` (opensmile_extractor.py:586-592)."""
    rng = np.random.default_rng(seed)
    mel_seq = 256 if fps == 30 else 512
    k_mel = mel_seq + 3
    hd = d_model // 2
    qs = 0.02 if style == "init" else 8.0

    def lin(out_f, in_f):
        b = 1.0 / math.sqrt(in_f)
        return (rng.uniform(-b, b, (out_f, in_f)).astype(np.float32),
                rng.uniform(-b, b, (out_f,)).astype(np.float32))

    def mha(prefix, sd):
        b = math.sqrt(6.0 / (d_model + 3 * d_model))  # xavier_uniform on (3d, d)
        sd[prefix + ".in_proj_weight"] = rng.uniform(-b, b, (3 * d_model, d_model)).astype(np.float32)
        sd[prefix + ".in_proj_bias"] = rng.uniform(-0.05, 0.05, (3 * d_model,)).astype(np.float32)
        w, bb = lin(d_model, d_model)
        sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"] = w, bb

    sd: Dict[str, np.ndarray] = {}
    sd["smoothing_alpha"] = np.float32(0.8) * np.ones((), np.float32)
    p = "dual_stream_attention."
    sd[p + "mouth_queries"] = (rng.standard_normal((len(MOUTH_INDICES), d_model)) * qs).astype(np.float32)
    sd[p + "expression_queries"] = (rng.standard_normal((len(EXPRESSION_INDICES), d_model)) * qs).astype(np.float32)
    mw = np.full(N_BLENDSHAPES, 0.5, np.float32)
    ew = np.full(N_BLENDSHAPES, 2.0, np.float32)
    mw[MOUTH_INDICES] = 2.0
    ew[MOUTH_INDICES] = 0.5
    jitter = 0.05 if style == "init" else 1.0
    sd[p + "mel_weights"] = (mw + jitter * rng.standard_normal(N_BLENDSHAPES)).astype(np.float32)
    sd[p + "emotion_weights"] = (ew + jitter * rng.standard_normal(N_BLENDSHAPES)).astype(np.float32)
    sd[p + "mel_channel_encoder.weight"], sd[p + "mel_channel_encoder.bias"] = lin(d_model, k_mel)
    mha(p + "mel_attention", sd)
    sd[p + "emotion_encoder.weight"], sd[p + "emotion_encoder.bias"] = lin(d_model, 256)
    mha(p + "emotion_attention", sd)
    sd[p + "mel_output_proj.weight"], sd[p + "mel_output_proj.bias"] = lin(d_model, d_model)
    sd[p + "emotion_output_proj.weight"], sd[p + "emotion_output_proj.bias"] = lin(d_model, d_model)
    sd[p + "blendshape_decoder.0.weight"], sd[p + "blendshape_decoder.0.bias"] = lin(hd, d_model)
    sd[p + "blendshape_decoder.3.weight"], sd[p + "blendshape_decoder.3.bias"] = lin(1, hd)
    for n in ("mel_norm", "emotion_norm"):
        sd[p + n + ".weight"] = (1.0 + 0.1 * rng.standard_normal(d_model)).astype(np.float32)
        sd[p + n + ".bias"] = (0.1 * rng.standard_normal(d_model)).astype(np.float32)
    if style == "stress":
        # make the decoder output move: scale the last layer so sigmoid is not pinned near 0.5
        sd[p + "blendshape_decoder.3.weight"] = sd[p + "blendshape_decoder.3.weight"] * 4.0
    sd["compression.weight"], sd["compression.bias"] = lin(256, 264)
    return sd


def model_state_dict(weights: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
    """The subset of ``make_weights`` that is in the reference's state_dict, as torch tensors."""
    return {k: torch.from_numpy(np.array(v)) for k, v in weights.items() if not k.startswith("compression.")}


def make_inputs(seed: int, batch: int, n_samples: int, kind: str = "noise") -> Tuple[np.ndarray, np.ndarray]:
    """Seeded synthetic audio (B, L) float32 and eGeMAPS windows (B, 264) float32.

    kinds: noise | silence_burst | sine | level_step | speechlike | silence"""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / SAMPLE_RATE
    clips = []
    for b in range(batch):
        if kind == "noise":
            a = 0.1 * rng.standard_normal(n_samples)
        elif kind == "silence":
            a = np.zeros(n_samples)
        elif kind == "silence_burst":
            a = np.zeros(n_samples)
            s = int(rng.integers(n_samples // 4, n_samples // 2))
            w = min(4000, n_samples - s)
            a[s:s + w] = 0.5 * rng.standard_normal(w)
        elif kind == "sine":
            a = 0.3 * np.sin(2 * np.pi * (440.0 * (b + 1)) * t)
        elif kind == "level_step":
            a = rng.standard_normal(n_samples)
            a[: n_samples // 2] *= 1e-4
            a[n_samples // 2:] *= 0.5
        elif kind == "speechlike":
            f0 = 110.0 + 40.0 * b
            env = 0.5 * (1 + np.sin(2 * np.pi * 3.1 * t + b))
            a = sum((0.4 / h) * np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 6.28)) for h in range(1, 30))
            a = env * a + 0.003 * rng.standard_normal(n_samples)
        else:
            raise ValueError(kind)
        clips.append(a)
    audio = np.stack(clips).astype(np.float32)
    egemaps = rng.standard_normal((batch, 264)).astype(np.float32)
    return audio, egemaps


# ---------------------------------------------------------------------------
# Model restatement (torch CPU, explicit math; dtype-parametrised)
# ---------------------------------------------------------------------------
def _t(x, dtype):
    return torch.as_tensor(np.asarray(x)).to(dtype)


def _layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _mha(query, kv, in_w, in_b, out_w, out_b, num_heads, need_weights):
    """torch.nn.MultiheadAttention(batch_first=True) in eval mode, written out."""
    d = query.shape[-1]
    hd = d // num_heads
    q = query @ in_w[:d].T + in_b[:d]
    k = kv @ in_w[d:2 * d].T + in_b[d:2 * d]
    v = kv @ in_w[2 * d:].T + in_b[2 * d:]
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    q = q.view(B, Lq, num_heads, hd).transpose(1, 2)
    k = k.view(B, Lk, num_heads, hd).transpose(1, 2)
    v = v.view(B, Lk, num_heads, hd).transpose(1, 2)
    p = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, Lq, d)
    o = o @ out_w.T + out_b
    return o, (p.mean(dim=1) if need_weights else None)


def compress_egemaps(egemaps264, weights, dtype=torch.float32):
    """opensmile_extractor.py:583-604: concat(3x88) -> Linear(264, 256)."""
    x = _t(egemaps264, dtype).reshape(-1, 264)
    return x @ _t(weights["compression.weight"], dtype).T + _t(weights["compression.bias"], dtype)


def dual_stream_core(weights, mel_long, mel_short, emotion256, num_heads: int = 8, temperature: float = 1.0,
                     mel_sequence_length: int = 256, return_attention: bool = False, dtype=torch.float32,
                     return_presigmoid: bool = False):
    """dual_stream_attention.py:162-280 (eval mode: dropout is a no-op)."""
    W = {k[len("dual_stream_attention."):]: _t(v, dtype) for k, v in weights.items()
         if k.startswith("dual_stream_attention.")}
    mel = _t(mel_long, dtype).transpose(1, 2)
    B, C, T = mel.shape
    if T < mel_sequence_length:
        mel = torch.cat([mel, torch.zeros(B, C, mel_sequence_length - T, dtype=dtype)], dim=2)
    elif T > mel_sequence_length:
        mel = mel[:, :, :mel_sequence_length]
    x = torch.cat([mel, _t(mel_short, dtype).transpose(1, 2)], dim=2)
    enc = _layer_norm(x @ W["mel_channel_encoder.weight"].T + W["mel_channel_encoder.bias"],
                      W["mel_norm.weight"], W["mel_norm.bias"])
    emo = _t(emotion256, dtype)
    eenc = _layer_norm((emo @ W["emotion_encoder.weight"].T + W["emotion_encoder.bias"]).unsqueeze(1),
                       W["emotion_norm.weight"], W["emotion_norm.bias"])
    mq = W["mouth_queries"].unsqueeze(0).expand(B, -1, -1)
    eq = W["expression_queries"].unsqueeze(0).expand(B, -1, -1)
    mo, mw = _mha(mq, enc, W["mel_attention.in_proj_weight"], W["mel_attention.in_proj_bias"],
                  W["mel_attention.out_proj.weight"], W["mel_attention.out_proj.bias"], num_heads, return_attention)
    mo = mo @ W["mel_output_proj.weight"].T + W["mel_output_proj.bias"]
    eo, ew = _mha(eq, eenc, W["emotion_attention.in_proj_weight"], W["emotion_attention.in_proj_bias"],
                  W["emotion_attention.out_proj.weight"], W["emotion_attention.out_proj.bias"], num_heads,
                  return_attention)
    eo = eo @ W["emotion_output_proj.weight"].T + W["emotion_output_proj.bias"]
    d = mo.shape[-1]
    comb = torch.zeros(B, N_BLENDSHAPES, d, dtype=dtype)
    comb[:, MOUTH_INDICES] = mo
    comb[:, EXPRESSION_INDICES] = eo
    h = torch.relu(comb @ W["blendshape_decoder.0.weight"].T + W["blendshape_decoder.0.bias"])
    logit = (h @ W["blendshape_decoder.3.weight"].T + W["blendshape_decoder.3.bias"]).squeeze(-1)
    blend = torch.sigmoid(logit)
    nm = torch.softmax(W["mel_weights"] / temperature, dim=0)
    ne = torch.softmax(W["emotion_weights"] / temperature, dim=0)
    final = torch.clamp(nm * blend * 0.5 + ne * blend * 0.5, 0, 1)
    out = {"blendshapes": final, "sigmoid": blend}
    if return_presigmoid:
        out["logit"] = logit
    if return_attention:
        out["mel_attention_weights"] = mw
        out["emotion_attention_weights"] = ew
        mb = torch.zeros_like(blend)
        eb = torch.zeros_like(blend)
        mb[:, MOUTH_INDICES] = blend[:, MOUTH_INDICES]
        eb[:, EXPRESSION_INDICES] = blend[:, EXPRESSION_INDICES]
        out["mel_blendshapes"] = mb
        out["emotion_blendshapes"] = eb
    return out


def smoothing_alpha(weights) -> float:
    """simplified_dual_stream_model.py:362 -- sigmoid(smoothing_alpha)."""
    return 1.0 / (1.0 + math.exp(-float(np.asarray(weights["smoothing_alpha"]))))


def ema_smooth(frames, alpha: float):
    """simplified_dual_stream_model.py:341-368 applied along dim 1 of (B, T, 52):
    s_0 = b_0 ; s_t = alpha * b_t + (1 - alpha) * s_{t-1}."""
    out = torch.empty_like(frames)
    prev = None
    for t in range(frames.shape[1]):
        cur = frames[:, t]
        prev = cur if prev is None else alpha * cur + (1 - alpha) * prev
        out[:, t] = prev
    return out


def forward_single(weights, audio, egemaps264, fps: int = 30, return_attention: bool = False,
                   dtype=torch.float32):
    """SimplifiedDualStreamModel.forward (first call after reset: smoothing is a passthrough,
    simplified_dual_stream_model.py:357-359).  audio (B, L), egemaps264 (B, 264) -> dict."""
    exact = dtype == torch.float64
    long_t, short_t = extract_mel_features(np.asarray(audio, np.float32), fps=fps, exact=exact)
    emo = compress_egemaps(egemaps264, weights, dtype)
    out = dual_stream_core(weights, long_t, short_t, emo, mel_sequence_length=256 if fps == 30 else 512,
                           return_attention=return_attention, dtype=dtype)
    out["logmel"] = torch.from_numpy(long_t)
    out["logmel_short"] = torch.from_numpy(short_t)
    return out


def forward_sequence(weights, audio, egemaps264, fps: int = 30, stride_frames: int = 1,
                     return_attention: bool = False, dtype=torch.float32, max_frames: Optional[int] = None):
    """SequentialDualStreamModel.forward (sequential_dual_stream_model.py:63-167):
    per output frame a zero-padded window of W*hop samples goes through the full
    librosa mel (W+1 frames), the core and the EMA.  Returns (B, T_out, 52)."""
    exact = dtype == torch.float64
    audio = np.asarray(audio, np.float32)
    B, L = audio.shape
    hop = hop_length_for(fps)
    W = 256 if fps == 30 else 512
    win = W * hop
    n_frames = L // hop
    t_out = max(1, (n_frames - W) // stride_frames + 1)
    if max_frames is not None:
        t_out = min(t_out, max_frames)
    emo = compress_egemaps(egemaps264, weights, dtype)
    alpha = smoothing_alpha(weights)
    frames, sig, mws = [], [], []
    for i in range(t_out):
        s = i * stride_frames * hop
        e = min(s + win, L)
        w = np.zeros((B, win), np.float32)
        w[:, : e - s] = audio[:, s:e]
        long_t, short_t = extract_mel_features(w, fps=fps, exact=exact)
        o = dual_stream_core(weights, long_t, short_t, emo, mel_sequence_length=W,
                             return_attention=return_attention, dtype=dtype)
        frames.append(o["blendshapes"])
        sig.append(o["sigmoid"])
        if return_attention:
            mws.append(o["mel_attention_weights"])
    raw = torch.stack(frames, dim=1)
    out = {"blendshapes": ema_smooth(raw, alpha), "raw_blendshapes": raw, "sigmoid": torch.stack(sig, dim=1),
           "num_frames": t_out, "fps": fps}
    if return_attention:
        out["mel_attention_weights"] = torch.stack(mws, dim=1)
    return out


def streaming_mel(ring_audio: np.ndarray, hop_length: int = 533, n_fft: int = 1024, f_min: float = 80.0,
                  f_max: float = 8000.0, context_window: float = 8.5, update_interval: float = 0.0333,
                  pad_mode: str = "reflect") -> np.ndarray:
    """mel_sliding_window.py:272-310: whole-ring mel, reflect padding, dB in [-80, 0]
    (NOT rescaled), truncated / last-frame-padded to int(context/update) frames."""
    mel = melspectrogram(np.asarray(ring_audio, np.float32), n_fft=n_fft, hop_length=hop_length,
                         fmin=f_min, fmax=f_max, pad_mode=pad_mode)
    log_mel = power_to_db(mel).T
    expected = int(context_window / update_interval)
    if log_mel.shape[0] > expected:
        log_mel = log_mel[:expected]
    elif log_mel.shape[0] < expected:
        log_mel = np.vstack([log_mel, np.tile(log_mel[-1:], (expected - log_mel.shape[0], 1))])
    return log_mel.astype(np.float32)
