"""Host-side chunk utilities of the agent loop, same names and semantics as
/root/reference/realtime_codec_agent/utils/audio_utils.py (smooth_join :4-17, create_crossfade_ramps
:19-23, pad_or_trim :25-37, normalize_audio_rms :39-46).

With a B200Generator these run on the device, fused behind the decoder (``emit_chunk_kernel``,
csrc/post_sm100.cuh, reached through ``OutputChunkEmitter``); the numpy versions here serve callers
that hold a plain array (and the duck-typed CPU models the host tests use).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def create_crossfade_ramps(sr: int, fade_secs: float) -> Tuple[int, np.ndarray, np.ndarray]:
    """(L, fade_in, fade_out): quarter-sine ramps over L = int(sr * fade_secs) samples."""
    n = int(sr * fade_secs)
    phase = np.linspace(0, 1, n, endpoint=False, dtype=np.float32)
    rising = np.sin(0.5 * np.pi * phase)
    return n, rising, rising[::-1]


def smooth_join(chunk1: np.ndarray, chunk2: np.ndarray, L: int, fade_in: np.ndarray, fade_out: np.ndarray) -> np.ndarray:
    """chunk1 ++ chunk2 with the last L samples of chunk1 cross-faded into the first L of chunk2."""
    if chunk1.shape[-1] == 0:
        return chunk2
    if L == 0:
        return np.concatenate((chunk1, chunk2), axis=-1)
    blended = chunk1[..., -L:] * fade_out + chunk2[..., :L] * fade_in
    return np.concatenate((chunk1[..., :-L], blended, chunk2[..., L:]), axis=-1)


def pad_or_trim(chunk: np.ndarray, target_length: int, pad_side: str = "right") -> np.ndarray:
    if chunk.ndim > 1:
        raise ValueError("Input chunk must be a 1D array.")
    have = chunk.shape[-1]
    if have > target_length:
        return chunk[..., :target_length]
    if have < target_length:
        missing = target_length - have
        return np.pad(chunk, (0, missing) if pad_side == "right" else (missing, 0), mode="constant")
    return chunk


def normalize_audio_rms(audio, target_rms=0.05, silence_rms_threshold=0.003):
    level = np.sqrt(np.mean(audio ** 2))
    if level < silence_rms_threshold:
        return audio                      # silence stays silence
    return audio * (target_rms / level)
