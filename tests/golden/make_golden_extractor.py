"""Generate tests/golden/extractor_default.npz: the UNMODIFIED reference ``MelSlidingWindowExtractor`` at its DEFAULT
geometry (n_fft = win_length = 512, f_max = sr // 2; src/features/mel_sliding_window.py:165-180), driven with 300 hops of
532 samples, plus its whole-clip ``process_audio_batch``.  Build container only (needs /root/reference); librosa is the
stand-in of oracle/run_reference.py (numpy restatement), everything else is reference code.

    python tests/golden/make_golden_extractor.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import koemorph_oracle as O  # noqa: E402
from oracle import run_reference as R  # noqa: E402


def main():
    _, _, _, Mel = R.import_reference()
    ex = Mel()                                         # every argument at the reference's default
    assert ex.n_fft == 512 and ex.win_length == 512 and ex.f_max == 8000 and ex.pad_mode == "reflect"
    hop = ex.audio_buffer.hop_length
    audio, _ = O.make_inputs(4343, 1, hop * 300, "speechlike")
    feats = None
    for i in range(300):
        ex.last_update_time = 0                        # defeat the wall-clock throttle (mel_sliding_window.py:266-269)
        f = ex.process_audio_frame(audio[0, i * hop:(i + 1) * hop])
        if f is not None:
            feats = f
    batch = ex.process_audio_batch(audio[0, :100000])
    out = {"hop": np.int64(hop), "stft_hop": np.int64(ex.hop_length), "features": feats.astype(np.float32),
           "batch_features": batch.astype(np.float32), "filterbank": np.asarray(ex.mel_transform, np.float32)}
    path = os.path.join(ROOT, "tests", "golden", "extractor_default.npz")
    np.savez_compressed(path, **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
