"""Host-side formats either side of the path (koemorph_b200.io): window grid, JSONL / UDP frames."""
import io
import json
import socket

import pytest
import torch

from koemorph_b200 import io as kio


def test_window_grid_matches_the_reference_contract():
    # sequential_dataset.py:182-192: num_windows = (frames - window) // stride + 1, samples = frames * hop
    g = kio.window_grid(600, window_frames=256, stride_frames=3, hop_length=533)
    assert len(g) == (600 - 256) // 3 + 1
    assert g[0] == (0, 256, 0, 256 * 533)
    assert g[-1][0] == 3 * (len(g) - 1) and g[-1][1] <= 600
    assert all(e - s == 256 and se - ss == 256 * 533 for s, e, ss, se in g)
    assert kio.window_grid(255, 256, 1, 533) == []
    assert len(kio.window_grid(256, 256, 1, 533)) == 1
    with pytest.raises(ValueError):
        kio.window_grid(10, 0, 1, 533)


def test_jsonl_round_trip_and_schema():
    torch.manual_seed(0)
    x = torch.rand(7, 52)
    buf = io.StringIO()
    assert kio.write_jsonl(x, 30, buf) == 7
    lines = buf.getvalue().splitlines()
    first = json.loads(lines[0])
    assert set(first) == {"timestamp", "blendshapes"} and len(first["blendshapes"]) == 52
    assert abs(first["timestamp"] - 1 / 30) < 1e-12            # README.md:97 -- 0.0333 at 30 fps
    ts, y = kio.read_jsonl(lines + ["", "  "])
    assert torch.equal(y, x) and torch.allclose(ts, torch.arange(1, 8, dtype=torch.float64) / 30)
    with pytest.raises(ValueError):
        kio.write_jsonl(torch.rand(3, 51), 30, io.StringIO())
    with pytest.raises(ValueError):
        kio.read_jsonl(['{"timestamp": 0.1, "blendshapes": [0.0]}'])


def test_streamer_udp_and_file(tmp_path):
    rx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    rx.bind(("127.0.0.1", 0))
    rx.settimeout(5)
    s = kio.BlendshapeStreamer("udp", "127.0.0.1", rx.getsockname()[1])
    s.send(torch.full((52,), 0.25), 1.5)
    msg = json.loads(rx.recv(65536).decode())
    assert msg["timestamp"] == 1.5 and msg["blendshapes"] == [0.25] * 52
    s.close()
    rx.close()
    path = tmp_path / "out.jsonl"
    f = kio.BlendshapeStreamer("file", output_file=str(path))
    f.send([0.5] * 52, 0.1)
    f.send([0.0] * 52, 0.2)
    f.close()
    ts, y = kio.read_jsonl(open(path))
    assert ts.tolist() == [0.1, 0.2] and y.shape == (2, 52)
    with pytest.raises(ValueError):
        kio.BlendshapeStreamer("file")
    with pytest.raises(ValueError):
        kio.BlendshapeStreamer("osc")
    with pytest.raises(ValueError):
        kio.BlendshapeStreamer("udp").send([0.0] * 3, 0.0)


def test_stft_golden_fixture_is_complete():
    """tests/golden/stft_reference.npz (outputs of the unmodified reference src/features/stft.py) covers every case of
    tests/golden/make_golden_stft.py with the frame count the reference contract gives: int(L / sr * fps)."""
    import os
    import sys
    import numpy as np
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    from make_golden_stft import CASES, stft_inputs
    data = np.load(os.path.join(here, "stft_reference.npz"))
    for name, (seed, B, L, kind, kw) in CASES.items():
        fps = kw.get("target_fps", 30.0)
        assert data[f"{name}/log_mel"].shape == (B, int(L / 16000 * fps), 80)
        assert data[f"{name}/mel_scale"].shape == (kw.get("n_fft", 512) // 2 + 1, 80)
        assert np.isfinite(data[f"{name}/log_mel"]).all()
    x = stft_inputs(11, 2, 16000, "noise")
    assert x.shape == (2, 16000) and torch.equal(x, stft_inputs(11, 2, 16000, "noise"))
