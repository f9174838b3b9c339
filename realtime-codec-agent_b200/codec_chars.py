"""Integer code <-> unicode char map of the path.

Mirrors the interface of the third-party ``codec_bpe`` converter the reference calls at
/root/reference/realtime_codec_agent/audio_tokenizer.py:89-95 (codes_to_chars) and :119-127
(chars_to_codes): one char per code, ``chr(unicode_offset + k*codebook_size + code)``,
frames major / codebooks minor.  Vectorised through a UTF-32 round trip instead of a
per-char Python loop.
"""
from __future__ import annotations

import numpy as np

UNICODE_OFFSET = 0x4E00
#: first code point after the surrogate block; prep_lm_dataset_magicodec.sh:4 pins 0xE000
UNICODE_OFFSET_LARGE = 0xE000


def _as_numpy(codes) -> np.ndarray:
    if hasattr(codes, "detach"):            # torch.Tensor without importing torch here
        codes = codes.detach().cpu().numpy()
    return np.asarray(codes)


def codes_to_chars(codes, codebook_size: int, copy_before_conversion: bool = True,
                   unicode_offset: int = UNICODE_OFFSET) -> str:
    arr = _as_numpy(codes)
    if arr.ndim != 2:
        raise ValueError("codes must be a 2D array of shape (num_codebooks, seq_length).")
    shift = unicode_offset + codebook_size * np.arange(arr.shape[0], dtype=np.int64)[:, None]
    if copy_before_conversion:
        pts = arr.astype(np.int64) + shift
    else:                                   # upstream mutates the caller's array in this mode
        arr += shift.astype(arr.dtype)
        pts = arr
    flat = np.ascontiguousarray(pts.T).reshape(-1).astype("<u4")
    return flat.tobytes().decode("utf-32-le", errors="surrogatepass")


def chars_to_codes(chars: str, num_codebooks: int, codebook_size: int, return_tensors=None,
                   unicode_offset: int = UNICODE_OFFSET):
    pts = np.frombuffer(chars.encode("utf-32-le", errors="surrogatepass"), dtype="<u4").astype(np.int64)
    pts = pts.reshape(-1, num_codebooks).T
    shift = unicode_offset + codebook_size * np.arange(num_codebooks, dtype=np.int64)[:, None]
    codes = np.ascontiguousarray(pts - shift)
    if return_tensors == "pt":
        import torch
        return torch.from_numpy(codes)
    return codes
