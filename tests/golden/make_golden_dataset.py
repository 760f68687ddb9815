"""Generate tests/golden/dataset_windows.npz by running the UNMODIFIED reference dataset class
(``SequentialKoeMorphDataset._process_file_pair``, src/data/sequential_dataset.py:157-206) on synthetic recordings.

Build container only (needs /root/reference).  ``soundfile`` and ``librosa`` are absent from this image; the dataset module
only uses them to read the WAV, so two stand-ins are installed before the import: ``soundfile.read`` returns the samples of a
``.npy`` file written next to the (empty) ``.wav``.  Everything that decides WHICH windows exist and what they contain --
the alignment rule, the window arithmetic, the size check -- is the reference's own code.

    python tests/golden/make_golden_dataset.py
"""
import json
import os
import sys
import tempfile
import types
import zlib
from pathlib import Path

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REFERENCE = os.environ.get("KOEMORPH_REFERENCE", "/root/reference")

# (name, n_samples, n_label_frames, window_frames, stride_frames, fps)
CASES = [
    ("aligned_stride1", 300 * 533 + 7, 300, 256, 1, 30),
    ("aligned_stride3", 400 * 533, 400, 256, 3, 30),
    ("labels_short_by_5", 300 * 533, 295, 256, 2, 30),       # > 1 frame off: both trimmed to 295 frames
    ("labels_long_by_4", 290 * 533 + 100, 294, 256, 1, 30),   # trimmed to 290 frames
    ("labels_long_by_1", 270 * 533, 271, 256, 1, 30),         # within tolerance: last window fails the size check
    ("too_short", 200 * 533, 200, 256, 1, 30),
    ("fps60_window64", 200 * 266 + 3, 200, 64, 5, 60),
]


def main():
    sf = types.ModuleType("soundfile")
    sf.read = lambda path, dtype="float32": (np.load(str(Path(path).with_suffix(".npy"))).astype(dtype), 16000)
    sys.modules["soundfile"] = sf
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    sys.path.insert(0, REFERENCE)
    from src.data.sequential_dataset import SequentialKoeMorphDataset

    out = {}
    for name, n_samples, n_labels, W, stride, fps in CASES:
        seed = zlib.crc32(name.encode())
        rng = np.random.default_rng(seed)
        audio = rng.standard_normal(n_samples).astype(np.float32)
        labels = rng.random((n_labels, 52)).astype(np.float32)
        with tempfile.TemporaryDirectory() as d:
            np.save(os.path.join(d, "rec.npy"), audio)
            open(os.path.join(d, "rec.wav"), "wb").close()
            with open(os.path.join(d, "rec.jsonl"), "w") as f:
                for i, row in enumerate(labels):
                    f.write(json.dumps({"timestamp": (i + 1) / fps, "blendshapes": [float(v) for v in row]}) + "\n")
            ds = SequentialKoeMorphDataset(d, window_frames=W, stride_frames=stride, target_fps=fps, shuffle_files=False,
                                           loop_dataset=False)
            wins = list(ds._process_file_pair(*ds.file_pairs[0]))
        out[f"{name}.seed"] = np.int64(seed)
        out[f"{name}.start_frames"] = np.array([int(w["start_frames"]) for w in wins], np.int64)
        out[f"{name}.audio_sum"] = np.array([float(w["audio"].double().sum()) for w in wins], np.float64)
        out[f"{name}.audio_first_last"] = np.array([[float(w["audio"][0]), float(w["audio"][-1])] for w in wins], np.float32).reshape(-1, 2)
        out[f"{name}.label_sum"] = np.array([float(w["blendshapes"].double().sum()) for w in wins], np.float64)
        print(name, len(wins), "windows")
    out["cases"] = np.array(json.dumps(CASES))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dataset_windows.npz"), **out)


if __name__ == "__main__":
    main()
