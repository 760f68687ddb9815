"""GPU tests (-m gpu) of the streaming paths: the MelSlidingWindowExtractor drop-in against the golden vectors of
the unmodified reference, and the hop-aligned StreamingEngine against SequentialDualStreamModel / the oracle."""
import numpy as np
import pytest
import torch

from oracle import koemorph_oracle as O

pytestmark = pytest.mark.gpu


def _close(got, ref, atol, what):
    got = got.detach().cpu().double().numpy() if torch.is_tensor(got) else np.asarray(got, np.float64)
    ref = ref.detach().cpu().double().numpy() if torch.is_tensor(ref) else np.asarray(ref, np.float64)
    assert got.shape == ref.shape, f"{what}: {got.shape} vs {ref.shape}"
    assert np.abs(got - ref).max() <= atol, f"{what}: max |d| = {np.abs(got - ref).max():.3e}"


def test_mel_sliding_window_extractor_vs_reference(golden):
    """300 hops of 532 samples through the device ring == the reference extractor's last features (dB, 255 x 80)."""
    from koemorph_b200.features.mel_sliding_window import MelSlidingWindowExtractor, create_mel_extractor
    _, data = golden
    ex = MelSlidingWindowExtractor(context_window=8.5, update_interval=0.0333, sample_rate=16000, n_mels=80,
                                   n_fft=1024, f_min=80.0, f_max=8000)
    hop = ex.audio_buffer.hop_length
    assert hop == int(data["streaming/hop"]) == 532 and ex.hop_length == 532 and ex.feature_shape == (255, 80)
    audio, _ = O.make_inputs(4242, 1, hop * 300, "speechlike")
    feats, n_none = None, 0
    for i in range(300):
        ex.last_update_time = 0   # defeat the wall-clock throttle exactly like tests/golden/make_golden.py
        f = ex.process_audio_frame(audio[0, i * hop:(i + 1) * hop])
        if f is None:
            n_none += 1
        else:
            feats = f
    assert n_none == 255          # ring (136000 samples) fills on the 256th hop of 532
    assert isinstance(feats, np.ndarray) and feats.dtype == np.float32
    # tolerance: 1e-4 of the 80 dB range = 8e-3 dB (north_star log-mel tolerance, here on the un-rescaled dB values)
    _close(feats, data["streaming/features"], 8e-3, "streaming features (dB)")
    assert feats.max() == 0.0 and feats.min() >= -80.0
    _close(ex.process_audio_batch(audio[0, :136000]), data["streaming/batch_features"], 8e-3, "batch features (dB)")
    st = ex.get_stats()
    assert st["buffer_stats"]["is_full"] and st["extraction_stats"]["total_extractions"] == 45
    assert ex.feature_dim == 80 and ex.get_current_features() is not None
    assert ex.process_audio_frame(np.zeros(100, np.float32)) is None      # wrong frame size is refused
    ex.reset()
    assert ex.get_current_features() is None and not ex.audio_buffer.is_full
    assert create_mel_extractor().n_fft == 512          # the reference's default (mel_sliding_window.py:169)
    with pytest.raises(NotImplementedError):
        MelSlidingWindowExtractor(n_fft=2048)


def test_mel_sliding_window_extractor_default_geometry_vs_reference():
    """The reference's DEFAULT constructor (n_fft = win_length = 512, f_max = sr // 2, reflect padding): golden from the
    unmodified reference class, tests/golden/make_golden_extractor.py."""
    import os
    from koemorph_b200.features.mel_sliding_window import MelSlidingWindowExtractor
    data = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "extractor_default.npz"))
    ex = MelSlidingWindowExtractor()
    assert (ex.n_fft, ex.win_length, ex.f_max, ex.pad_mode, ex.hop_length) == (512, 512, 8000, "reflect", int(data["stft_hop"]))
    hop = ex.audio_buffer.hop_length
    assert hop == int(data["hop"])
    fb = ex.mel_transform
    assert fb.shape == (80, 257)
    _close(fb, data["filterbank"], 2e-7 * float(np.abs(data["filterbank"]).max()), "512-point Slaney bank")
    audio, _ = O.make_inputs(4343, 1, hop * 300, "speechlike")
    feats = None
    for i in range(300):
        ex.last_update_time = 0
        f = ex.process_audio_frame(audio[0, i * hop:(i + 1) * hop])
        feats = f if f is not None else feats
    _close(feats, data["features"], 8e-3, "default-geometry streaming features (dB)")
    _close(ex.process_audio_batch(audio[0, :100000]), data["batch_features"], 8e-3, "default-geometry batch features (dB)")


def test_win_length_shorter_than_n_fft_matches_torch_stft():
    """win_length < n_fft: the Hann window of win_length points centred in the n_fft frame (librosa pad_center, the same
    convention as torch.stft, used here as an independent float64 cross-check)."""
    from koemorph_b200.features.mel_frontend import LogMelFrontend
    fe = LogMelFrontend.get("cuda", n_fft=1024, win_length=400)
    audio, _ = O.make_inputs(99, 2, 40000, "speechlike")
    a = torch.from_numpy(audio).cuda()
    n_frames = 1 + 40000 // 533
    db, _ = fe.power(a, 533, n_frames)
    win = torch.hann_window(400, periodic=True, dtype=torch.float64, device="cuda")
    S = torch.stft(a.double(), 1024, 533, 400, win, center=True, pad_mode="constant", return_complex=True)
    mel = torch.einsum("mf,bft->btm", torch.from_numpy(fe.filterbank()).cuda().double(), S.abs() ** 2)[:, :n_frames]
    ref = 10 * torch.log10(mel.clamp_min(1e-10))
    big = mel > 1e-6 * mel.amax(dim=(1, 2), keepdim=True)
    assert (db.double() - ref).abs()[big].max().item() < 1e-3


def test_streaming_engine_60fps_matches_sequence_model(golden):
    """hop 266 < n_fft / 2: two edge frames per window side, five FFTs per hop (SURVEY.md section 8 note E).  Fed hop by
    hop, the engine reproduces SequentialDualStreamModel.forward at 60 fps -- whose frames are pinned by the reference's
    golden vector ``seq_60fps`` -- on the native one-call step and on the call-by-call driver, fp32 and bf16."""
    import koemorph_b200 as K
    from koemorph_b200.streaming import StreamingEngine
    cases, data = golden
    spec = cases["seq_60fps"]
    w = O.make_weights(spec["wseed"], 60, style=spec["style"])
    audio, eg = O.make_inputs(spec["iseed"], spec["B"], spec["L"], spec["kind"])
    m = K.SequentialDualStreamModel(target_fps=60, mel_sequence_length=512).cuda().eval()
    m.load_state_dict(O.model_state_dict(w), strict=True)
    m.set_compression_layer(torch.from_numpy(w["compression.weight"]), torch.from_numpy(w["compression.bias"]))
    a, e = torch.from_numpy(audio).cuda(), torch.from_numpy(eg).cuda()
    want = torch.from_numpy(data["seq_60fps/blendshapes"])
    n_hops = spec["L"] // 266
    assert want.shape[1] == n_hops - 512 + 1
    for native in (True, False):
        eng = StreamingEngine(m, spec["B"])
        assert eng.n_edge == 2 and eng.tail_len == 512 + 2 * 266
        eng.native = native
        eng.set_egemaps(e)
        outs = []
        for n in range(n_hops):
            o = eng.step(a[:, n * 266:(n + 1) * 266].contiguous())
            assert (o is None) == (n < 511)
            if o is not None:
                outs.append(o.clone())
        got = torch.stack(outs, dim=1)
        _close(got, want, 2e-6, f"60 fps streaming (native={native}) vs the reference's golden frames")
        _close(got, m(a, egemaps=e)["blendshapes"], 2e-6, "60 fps streaming vs sequence forward")
        if native:
            first = got
        else:
            assert torch.equal(got, first), "native step and call-by-call driver differ at 60 fps"
    # ring wrap-around (more than one revolution of the 512-slot rings) and the tensor-core core
    m.precision = "bf16"
    eng = StreamingEngine(m, 2)
    eng.set_egemaps(e.expand(2, -1).contiguous())
    L2 = (2 * 512 + 7) * 266
    audio2, _ = O.make_inputs(1234, 2, L2, "speechlike")
    a2 = torch.from_numpy(audio2).cuda()
    last = None
    for n in range(L2 // 266):
        last = eng.step(a2[:, n * 266:(n + 1) * 266].contiguous())
    ref2 = m(a2, egemaps=e.expand(2, -1).contiguous())["blendshapes"]
    _close(last, ref2[:, -1], 1e-4, "60 fps bf16 after ring wrap")


def test_streaming_engine_matches_sequence_model():
    """Feeding a clip hop by hop gives, from hop 256 on, the frames of SequentialDualStreamModel.forward on the
    prefix heard so far -- 3 FFTs per hop instead of 257 -- and both match the CPU oracle."""
    import koemorph_b200 as K
    from koemorph_b200.streaming import StreamingEngine
    w = O.make_weights(1235, 30, style="stress")
    n_hops_extra = 9
    L = (256 + n_hops_extra) * 533
    audio, eg = O.make_inputs(77, 3, L, "speechlike")
    audio[2, : L // 2] = 0.0                                  # a stream that starts silent: exercises amin / the clamp
    m = K.SequentialDualStreamModel().cuda().eval()
    m.load_state_dict(O.model_state_dict(w), strict=True)
    m.set_compression_layer(torch.from_numpy(w["compression.weight"]), torch.from_numpy(w["compression.bias"]))
    a, e = torch.from_numpy(audio).cuda(), torch.from_numpy(eg).cuda()
    ref_seq = m(a, egemaps=e)["blendshapes"]                  # (3, 10, 52)
    assert ref_seq.shape == (3, n_hops_extra + 1, 52)
    eng = StreamingEngine(m, 3)
    eng.set_egemaps(e)
    outs = []
    for n in range(256 + n_hops_extra):
        o = eng.step(a[:, n * 533:(n + 1) * 533].contiguous())
        assert (o is None) == (n < 255)
        if o is not None:
            outs.append(o.clone())
    got = torch.stack(outs, dim=1)
    _close(got, ref_seq, 2e-6, "streaming engine vs sequence model")
    oracle = O.forward_sequence(w, audio, eg)["blendshapes"]
    _close(got, oracle, 2e-6, "streaming engine vs oracle")
    # ring wrap-around: keep going past one full ring revolution, compare with a fresh sequence forward
    eng.reset()
    eng.set_egemaps(e)
    L2 = (2 * 256 + 5) * 533
    audio2, _ = O.make_inputs(78, 3, L2, "noise")
    a2 = torch.from_numpy(audio2).cuda()
    last = None
    for n in range(2 * 256 + 5):
        last = eng.step(a2[:, n * 533:(n + 1) * 533].contiguous())
    ref2 = m(a2, egemaps=e)["blendshapes"]
    _close(last, ref2[:, -1], 2e-6, "after ring wrap")


def test_streaming_engine_bf16_and_realtime_model():
    import koemorph_b200 as K
    from koemorph_b200.streaming import StreamingEngine
    w = O.make_weights(1235, 30, style="stress")
    m = K.SequentialDualStreamModel().cuda().eval()
    m.load_state_dict(O.model_state_dict(w), strict=True)
    m.set_compression_layer(torch.from_numpy(w["compression.weight"]), torch.from_numpy(w["compression.bias"]))
    audio, eg = O.make_inputs(79, 2, 258 * 533, "speechlike")
    a, e = torch.from_numpy(audio).cuda(), torch.from_numpy(eg).cuda()
    ref = m(a, egemaps=e)["blendshapes"]
    m.precision = "bf16"
    eng = StreamingEngine(m, 2)
    eng.set_egemaps(e)
    outs = [eng.step(a[:, n * 533:(n + 1) * 533].contiguous()) for n in range(258)]
    _close(outs[-1], ref[:, -1], 1e-4, "bf16 streaming")
    # single-stream realtime API of the model (reference :452-500), driven with 532-sample hops like rt.py
    rt = K.SimplifiedDualStreamModel(real_time_mode=True).cuda().eval()
    rt.load_state_dict(O.model_state_dict(w), strict=True)
    rt.set_compression_layer(torch.from_numpy(w["compression.weight"]), torch.from_numpy(w["compression.bias"]))
    res = None
    for n in range(258):
        rt.mel_extractor.last_update_time = 0
        res = rt.process_audio_frame_realtime(audio[0, n * 532:(n + 1) * 532], egemaps=e[:1])
        assert (res is None) == (n < 255)
    assert res.shape == (52,) and bool(torch.isfinite(res).all())
    assert "mel_stats" in rt.get_realtime_stats()
    rt.reset_realtime_state()


def test_native_stream_step_equals_the_call_by_call_driver():
    """koe_stream_push (one native call per hop: tail shift, three frontend launches, the core, the smoothing) against
    the same step issued call by call from Python: every emitted frame bit for bit, and the persistent state with it."""
    import koemorph_b200 as K
    from koemorph_b200.streaming import StreamingEngine
    w = O.make_weights(77, 30, style="stress")
    m = K.SequentialDualStreamModel().cuda().eval()
    m.load_state_dict(O.model_state_dict(w), strict=True)
    m.set_compression_layer(torch.from_numpy(w["compression.weight"]), torch.from_numpy(w["compression.bias"]))
    S, steps = 3, 256 + 7
    audio, eg = O.make_inputs(5, S, steps * 533, "speechlike")
    a, e = torch.from_numpy(audio).cuda(), torch.from_numpy(eg).cuda()
    engines = []
    for native in (True, False):
        eng = StreamingEngine(m, S)
        eng.native = native
        eng.set_egemaps(e)
        engines.append(eng)
    emitted = 0
    for n in range(steps):
        hop = a[:, n * 533:(n + 1) * 533].contiguous()
        o_native, o_python = engines[0].step(hop), engines[1].step(hop)
        assert (o_native is None) == (o_python is None) == (n + 1 < 256)
        if o_native is not None:
            emitted += 1
            assert torch.equal(o_native, o_python), f"hop {n}"
    assert emitted == 8
    for name in ("ring_f", "ring_r", "fmax_f", "fmax_r", "row_l", "fmax_l", "state"):
        assert torch.equal(getattr(engines[0], name), getattr(engines[1], name)), name
    # reset() starts both from silence again
    engines[0].reset()
    assert engines[0].step(a[:, :533].contiguous()) is None and engines[0].n == 1
