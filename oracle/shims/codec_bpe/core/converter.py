"""ORACLE SHIM: codec_bpe.core.converter (call sites: audio_tokenizer.py:89-95,119-127;
lm_dataset_builder.py:10,412-417; run_stream_codes.py:47-53)."""
import numpy as np

UNICODE_OFFSET = 0x4E00
UNICODE_OFFSET_LARGE = 0xE000


def codes_to_chars(codes, codebook_size, copy_before_conversion=True, unicode_offset=UNICODE_OFFSET):
    try:
        import torch
        if isinstance(codes, torch.Tensor):
            codes = codes.cpu().numpy()
    except ImportError:  # pragma: no cover
        pass
    codes = np.asarray(codes)
    if codes.ndim != 2:
        raise ValueError("codes must be a 2D array of shape (num_codebooks, seq_length).")
    if copy_before_conversion:
        codes = codes.copy()
    for k in range(codes.shape[0]):
        codes[k] += unicode_offset + k * codebook_size
    return "".join(chr(int(c)) for c in codes.T.reshape(-1))


def chars_to_codes(chars, num_codebooks, codebook_size, return_tensors=None, unicode_offset=UNICODE_OFFSET):
    codes = np.array([ord(c) for c in chars], dtype=np.int64).reshape(-1, num_codebooks).T.copy()
    for k in range(codes.shape[0]):
        codes[k] -= unicode_offset + k * codebook_size
    if return_tensors == "pt":
        import torch
        return torch.tensor(codes)
    return codes
