"""Golden vectors for the torchaudio-flavoured frontend: the UNMODIFIED reference ``src/features/stft.py``
``MelSpectrogramExtractor`` (it imports fine in the build container: torch + torchaudio only) on seeded inputs.

    python tests/golden/make_golden_stft.py        # needs /root/reference; writes tests/golden/stft_reference.npz

The inputs are regenerated from the seeds by tests/test_gpu_stft.py (``stft_inputs`` below), only the outputs and the
filterbank are stored."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (seed, B, L, kind, constructor kwargs)
    "default_noise_1s": (11, 2, 16000, "noise", {}),
    "default_ramp_8p5s": (12, 2, 136000, "ramp", {}),
    "default_tone_3s": (13, 1, 48000, "tone", {}),
    "nfft1024_noise_2s": (14, 2, 32000, "noise", {"n_fft": 1024}),
    "unnormalised_constant_pad": (15, 2, 20000, "noise", {"normalized": False, "pad_mode": "constant"}),
    "fps60_noise_2s": (16, 2, 32000, "noise", {"target_fps": 60.0}),
}


def stft_inputs(seed: int, B: int, L: int, kind: str) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    x = 0.1 * torch.randn(B, L, generator=g)
    if kind == "ramp":
        x = x * torch.linspace(1e-4, 1.0, L)
    elif kind == "tone":
        t = torch.arange(L) / 16000.0
        x = 0.5 * torch.sin(2 * torch.pi * 440.0 * t).expand(B, L).clone() + 1e-3 * x
    return x.float()


def main():
    sys.path.insert(0, "/root/reference")
    warnings.filterwarnings("ignore")
    from src.features.stft import MelSpectrogramExtractor
    out = {}
    for name, (seed, B, L, kind, kw) in CASES.items():
        m = MelSpectrogramExtractor(**kw).eval()
        with torch.no_grad():
            y = m(stft_inputs(seed, B, L, kind))
        out[f"{name}/log_mel"] = y.numpy().astype(np.float32)
        out[f"{name}/mel_scale"] = m.mel_scale.numpy().astype(np.float32)
        print(name, tuple(y.shape), float(y.min()), float(y.max()))
    np.savez_compressed(os.path.join(HERE, "stft_reference.npz"), **out)


if __name__ == "__main__":
    main()
