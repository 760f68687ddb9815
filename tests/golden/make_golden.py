"""Generate tests/golden/reference_outputs.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every case is (weights seed/style, input seed/kind/shape) -> outputs of the
reference's own ``SimplifiedDualStreamModel`` / ``SequentialDualStreamModel`` /
``MelSlidingWindowExtractor`` executed through ``oracle/run_reference.py``
(librosa + opensmile replaced by the stand-ins described there; everything
else is reference code).  Inputs and weights are NOT stored: tests regenerate
them from the seeds with ``oracle.koemorph_oracle.make_weights/make_inputs``.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import koemorph_oracle as O  # noqa: E402
from oracle import run_reference as R  # noqa: E402

CLIP = 136000  # 8.5 s at 16 kHz

# name -> spec.  mode: single (SimplifiedDualStreamModel.forward), sequence (SequentialDualStreamModel.forward)
CASES = {
    "single_noise_init":      dict(mode="single", fps=30, wseed=1234, style="init",   iseed=5678, kind="noise",         B=2, L=CLIP),
    "single_speech_stress":   dict(mode="single", fps=30, wseed=1235, style="stress", iseed=5679, kind="speechlike",    B=2, L=CLIP),
    "single_burst_stress":    dict(mode="single", fps=30, wseed=1235, style="stress", iseed=5680, kind="silence_burst", B=2, L=CLIP),
    "single_sine_stress":     dict(mode="single", fps=30, wseed=1235, style="stress", iseed=5681, kind="sine",          B=2, L=CLIP),
    "single_step_stress":     dict(mode="single", fps=30, wseed=1235, style="stress", iseed=5682, kind="level_step",    B=2, L=CLIP),
    "single_silence_init":    dict(mode="single", fps=30, wseed=1234, style="init",   iseed=5683, kind="silence",       B=1, L=CLIP),
    "single_short_clip":      dict(mode="single", fps=30, wseed=1235, style="stress", iseed=5684, kind="speechlike",    B=2, L=40000),
    "single_long_clip":       dict(mode="single", fps=30, wseed=1235, style="stress", iseed=5685, kind="speechlike",    B=1, L=160000),
    "single_60fps":           dict(mode="single", fps=60, wseed=1236, style="stress", iseed=5686, kind="speechlike",    B=2, L=CLIP),
    "seq_clip_T1":            dict(mode="sequence", fps=30, wseed=1235, style="stress", iseed=5687, kind="level_step",  B=2, L=CLIP),
    "seq_13_frames":          dict(mode="sequence", fps=30, wseed=1235, style="stress", iseed=5688, kind="speechlike",  B=2, L=CLIP + 12 * 533 + 5),
    "seq_20s":                dict(mode="sequence", fps=30, wseed=1235, style="stress", iseed=5689, kind="speechlike",  B=1, L=320000),
    "seq_stride3":            dict(mode="sequence", fps=30, wseed=1235, style="stress", iseed=5690, kind="noise",       B=1, L=CLIP + 20 * 533, stride=3),
    "seq_60fps":              dict(mode="sequence", fps=60, wseed=1236, style="stress", iseed=5691, kind="speechlike",  B=1, L=CLIP + 9 * 266 + 100),
    "seq_burst_edge":         dict(mode="sequence", fps=30, wseed=1235, style="stress", iseed=5692, kind="silence_burst", B=1, L=CLIP + 40 * 533),
}


def run_case(spec):
    w = O.make_weights(spec["wseed"], spec["fps"], style=spec["style"])
    audio, eg = O.make_inputs(spec["iseed"], spec["B"], spec["L"], spec["kind"])
    seq = spec["mode"] == "sequence"
    model = R.build_reference_model(w, spec["fps"], sequential=seq, stride_frames=spec.get("stride", 1))
    model.set_egemaps(eg)
    out = {}
    with torch.no_grad():
        res = model(torch.from_numpy(audio), return_attention=True)
        out["blendshapes"] = res["blendshapes"].numpy()
        out["mel_attention_weights"] = res["mel_attention_weights"].numpy().astype(np.float32)
        if not seq:
            lt, st = model.extract_mel_features(torch.from_numpy(audio))
            out["logmel"] = lt.numpy()
            out["logmel_short"] = st.numpy()
            out["mel_blendshapes"] = res["mel_blendshapes"].numpy()
            out["emotion_blendshapes"] = res["emotion_blendshapes"].numpy()
        else:
            if out["mel_attention_weights"].shape[1] > 16:  # keep the fixture small
                out["mel_attention_weights"] = out["mel_attention_weights"][:, :16]
    return out


def streaming_case():
    """MelSlidingWindowExtractor driven like test_realtime_dual_stream.py: 300 hops of 532 samples."""
    _, _, _, Mel = R.import_reference()
    ex = Mel(context_window=8.5, update_interval=0.0333, sample_rate=16000, n_mels=80, n_fft=1024,
             f_min=80.0, f_max=8000)
    audio, _ = O.make_inputs(4242, 1, ex.audio_buffer.hop_length * 300, "speechlike")
    hop = ex.audio_buffer.hop_length
    feats = None
    for i in range(300):
        ex.last_update_time = 0  # defeat the wall-clock throttle (mel_sliding_window.py:266-269)
        f = ex.process_audio_frame(audio[0, i * hop:(i + 1) * hop])
        if f is not None:
            feats = f
    batch = ex.process_audio_batch(audio[0, :136000])
    return {"hop": np.int64(hop), "features": feats, "batch_features": batch}


def main():
    arrays = {}
    for name, spec in CASES.items():
        print("case", name, spec, flush=True)
        for k, v in run_case(spec).items():
            arrays[f"{name}/{k}"] = v
    for k, v in streaming_case().items():
        arrays[f"streaming/{k}"] = v
    here = os.path.dirname(os.path.abspath(__file__))
    np.savez_compressed(os.path.join(here, "reference_outputs.npz"), **arrays)
    with open(os.path.join(here, "cases.json"), "w") as f:
        json.dump(CASES, f, indent=1, sort_keys=True)
    print("wrote", len(arrays), "arrays;", os.path.getsize(os.path.join(here, "reference_outputs.npz")), "bytes")


if __name__ == "__main__":
    main()
