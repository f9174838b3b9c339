"""ORACLE (test infrastructure): import the reference's AudioTokenizer UNMODIFIED.

Loads /root/reference/realtime_codec_agent/audio_tokenizer.py by file path (not via the
package __init__, which drags in llama_cpp: realtime_codec_agent/__init__.py:1-5) with the
shims in oracle/shims on sys.path.  Only available where /root/reference exists (the build
container); the GPU box uses the golden vectors made with it (tests/golden/make_golden.py).
"""
import importlib.util
import os
import sys

REFERENCE_FILE = "/root/reference/realtime_codec_agent/audio_tokenizer.py"
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available() -> bool:
    return os.path.isfile(REFERENCE_FILE)


def load_reference_audio_tokenizer():
    """Returns the reference's AudioTokenizer class, byte-for-byte as shipped."""
    if not reference_available():
        raise FileNotFoundError(REFERENCE_FILE)
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    name = "_reference_audio_tokenizer"
    if name in sys.modules:
        return sys.modules[name].AudioTokenizer
    spec = importlib.util.spec_from_file_location(name, REFERENCE_FILE)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod.AudioTokenizer
