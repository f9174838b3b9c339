"""ORACLE SHIM (test infrastructure) for the un-vendored pip package ``codec-bpe``
(requirements.txt:2, unpinned), so that the reference wrapper
/root/reference/realtime_codec_agent/audio_tokenizer.py:7-8 imports unmodified.

Restates the published converter (codec_bpe/core/converter.py upstream):
one unicode char per code, ``chr(unicode_offset + k*codebook_size + code)``,
time-major flatten of a [num_codebooks, T] array, and its inverse.
UNICODE_OFFSET_LARGE must equal 0xE000: the LM data is built with
``--unicode_offset=0xE000`` (prep_lm_dataset_magicodec.sh:4) while the agent
uses the UNICODE_OFFSET_LARGE default (audio_tokenizer.py:16).
"""
from .core.converter import codes_to_chars, chars_to_codes, UNICODE_OFFSET, UNICODE_OFFSET_LARGE  # noqa: F401
