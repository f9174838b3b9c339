"""ORACLE SHIM for librosa (audio_tokenizer.py:2,211,214): to_mono = mean over channels;
resample is not exercised by any BASELINE config (all synthetic audio is 16 kHz)."""
import numpy as np


def to_mono(y):
    y = np.asarray(y)
    return y.mean(axis=0) if y.ndim > 1 else y


def resample(y, *, orig_sr, target_sr, **kw):
    if orig_sr == target_sr:
        return y
    raise NotImplementedError("oracle shim: resampling is outside the hot path under test")
