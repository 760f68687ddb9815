"""GPU parity of the torchaudio-flavoured frontend (koemorph_b200.features.stft.MelSpectrogramExtractor) against golden
vectors produced by the unmodified reference ``src/features/stft.py`` (tests/golden/make_golden_stft.py): PINNED parity.

Tolerance: |d| <= 2e-4 + 1e-4 * |ref| on ln(mel + 1e-8) -- 2e-4 in the log is 2e-4 relative in mel power; the reference
itself computes the STFT in float32 (torch.stft), which is where the last digits differ."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_stft import CASES, stft_inputs  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stft_reference.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("name", sorted(CASES))
def test_mel_spectrogram_extractor_vs_reference(golden, name):
    from koemorph_b200.features.stft import MelSpectrogramExtractor
    seed, B, L, kind, kw = CASES[name]
    m = MelSpectrogramExtractor(**kw).cuda().eval()
    y = m(stft_inputs(seed, B, L, kind).cuda()).cpu().double().numpy()
    ref = golden[f"{name}/log_mel"].astype(np.float64)
    assert y.shape == ref.shape
    err = np.abs(y - ref) - (2e-4 + 1e-4 * np.abs(ref))
    assert err.max() <= 0, f"max |d| = {np.abs(y - ref).max():.3e}"
    fb = golden[f"{name}/mel_scale"]
    got = m.mel_scale.cpu().numpy()
    # torchaudio builds its bank in float32 (weights in [0, 1]); ours is float64 rounded once
    assert got.shape == fb.shape and np.abs(got - fb).max() <= 2e-5 and ((got > 1e-4) == (fb > 1e-4)).mean() > 0.999


def test_extractor_api_and_errors():
    from koemorph_b200.features.stft import MelSpectrogramExtractor
    m = MelSpectrogramExtractor().cuda()
    assert m.hop_length == 533 and m.get_output_length(16000) == 31
    assert torch.allclose(m.get_time_axis(3), torch.tensor([0.0, 533 / 16000, 1066 / 16000]))
    assert m(torch.zeros(16000, device="cuda")).shape == (1, 30, 80)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 16000))                      # no CPU path
    with pytest.raises(ValueError):
        m(torch.zeros(1, 2, 16000, device="cuda"))
    with pytest.raises(NotImplementedError):
        MelSpectrogramExtractor(n_fft=400)
