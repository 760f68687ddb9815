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


# ---- window slicing of long recordings against the UNMODIFIED reference dataset class (tests/golden/make_golden_dataset.py)
def _dataset_cases():
    import os
    import numpy as np
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset_windows.npz")
    data = np.load(here)
    return json.loads(str(data["cases"])), data


def _check_recording_windows(device):
    import numpy as np
    cases, data = _dataset_cases()
    assert len(cases) >= 7
    for name, n_samples, n_labels, W, stride, fps in cases:
        rng = np.random.default_rng(int(data[f"{name}.seed"]))
        audio = torch.from_numpy(rng.standard_normal(n_samples).astype(np.float32)).to(device)
        labels = torch.from_numpy(rng.random((n_labels, 52)).astype(np.float32)).to(device)
        hop = int(16000 / fps)
        out = kio.recording_windows(audio, labels, W, stride, hop)
        want_start = data[f"{name}.start_frames"]
        assert out["start_frames"].cpu().tolist() == want_start.tolist(), name
        assert out["audio"].shape == (len(want_start), W * hop) and out["blendshapes"].shape == (len(want_start), W, 52)
        assert out["audio"].device == audio.device
        if len(want_start) == 0:
            continue
        assert out["audio"].data_ptr() == audio.data_ptr(), "windows must be views, not copies"
        np.testing.assert_allclose(out["audio"].double().sum(1).cpu().numpy(), data[f"{name}.audio_sum"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(out["blendshapes"].double().sum((1, 2)).cpu().numpy(), data[f"{name}.label_sum"], rtol=0, atol=1e-9)
        fl = torch.stack([out["audio"][:, 0], out["audio"][:, -1]], 1).cpu().numpy()
        assert (fl == data[f"{name}.audio_first_last"]).all(), name
        # the window grid of the model-side helper describes the same windows
        n_s, n_f = kio.align_recording(n_samples, n_labels, hop)
        grid = [g for g in kio.window_grid(n_f, W, stride, hop) if g[3] <= n_s]
        assert [g[0] for g in grid] == want_start.tolist()


def test_recording_windows_match_the_reference_dataset_cpu():
    _check_recording_windows("cpu")


@pytest.mark.gpu
def test_recording_windows_match_the_reference_dataset_on_device():
    _check_recording_windows("cuda")


@pytest.mark.gpu
def test_windows_of_a_recording_fed_as_clips_equal_the_sliding_forward():
    """Window i of a long recording, cut out on the device and fed to the model as a clip of its own, gives frame i of the
    model's own slide over the whole recording (smoothing off: it couples consecutive frames)."""
    import numpy as np
    import koemorph_b200 as K
    from oracle import koemorph_oracle as O
    w = O.make_weights(1235, 30, style="stress")
    m = K.SequentialDualStreamModel(stride_frames=2).cuda().eval()
    m.load_state_dict(O.model_state_dict(w), strict=True)
    m.set_compression_layer(torch.from_numpy(w["compression.weight"]), torch.from_numpy(w["compression.bias"]))
    m.use_temporal_smoothing = False
    audio, eg = O.make_inputs(4321, 1, 136000 + 11 * 533, "speechlike")
    audio, eg = torch.from_numpy(audio).cuda(), torch.from_numpy(eg).cuda()
    whole = m(audio, egemaps=eg)["blendshapes"][0]                       # (T_out, 52)
    labels = torch.zeros(audio.shape[1] // 533, 52, device="cuda")
    wins = kio.recording_windows(audio[0], labels, 256, 2, 533)
    assert wins["audio"].shape[0] == whole.shape[0] == 6
    per_window = m(wins["audio"].contiguous(), egemaps=eg.expand(6, -1).contiguous())["blendshapes"][:, 0]
    # not bitwise: a frame's rounding depends on the frame it shares its complex transform with, and the window-edge frames
    # are paired differently in the two runs
    assert (per_window - whole).abs().max().item() < 2e-7
