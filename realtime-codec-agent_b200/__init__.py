"""realtime-codec-agent_b200 — B200-native (sm_100a) MagiCodec tokenization path.

The directory name carries a hyphen (it mirrors the reference repo's name), so it is imported
under the module name ``realtime_codec_agent_b200`` through ``rca_b200_loader`` at the repo
root (``import rca_b200_loader`` registers the package in ``sys.modules``).

Public surface (mirrors the reference's interface for this path):
    AudioTokenizer            — realtime_codec_agent/audio_tokenizer.py:10
    codes_to_chars / chars_to_codes / UNICODE_OFFSET(_LARGE)   — codec_bpe converter
    B200Generator             — the model duck type of audio_tokenizer.py:28-36,158,190-200
    MagiCodecSpec, DEFAULT_SPEC, init_random_weights
    smooth_join / create_crossfade_ramps / pad_or_trim / normalize_audio_rms   — utils/audio_utils.py:4-46
    OutputChunkEmitter        — RealtimeAgent.detokenize_output_chunk, realtime_agent_v2.py:556-579
    ExternalTTSDuplexAligner  — external_tts_duplex_aligner.py:6-27
    SessionBatcher / ThreadedSessionBatcher   — many live streams on one engine (tts_server.py:59,158)
    audio_io (read_audio / load_audio / resample / DeviceIngest), audio_to_codes (the offline CLI)
"""
from .spec import MagiCodecSpec, DEFAULT_SPEC, TINY_SPEC, MID_SPEC  # noqa: F401
from .weights import init_random_weights, param_shapes, save_checkpoint, load_checkpoint  # noqa: F401
from .codec_chars import codes_to_chars, chars_to_codes, UNICODE_OFFSET, UNICODE_OFFSET_LARGE  # noqa: F401
from .audio_tokenizer import AudioTokenizer, load_magicodec_model  # noqa: F401
from .synth import synth_audio  # noqa: F401
from .audio_utils import smooth_join, create_crossfade_ramps, pad_or_trim, normalize_audio_rms  # noqa: F401
from .output_chunks import OutputChunkEmitter  # noqa: F401
from .duplex_aligner import ExternalTTSDuplexAligner  # noqa: F401


def __getattr__(name):
    # the native binding is imported lazily so that host-only tooling works without the .so
    if name == "B200Generator":
        from .generator import B200Generator
        return B200Generator
    if name in ("SessionBatcher", "ThreadedSessionBatcher", "SessionPool"):
        from . import session_batcher
        return getattr(session_batcher, name)
    if name == "native":
        from . import _native
        return _native
    raise AttributeError(name)
