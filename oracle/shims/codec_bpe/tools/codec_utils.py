"""ORACLE SHIM: codec_bpe.tools.codec_utils.load_magicodec_model (audio_tokenizer.py:8,27).
No checkpoint exists offline; tests always pass a model OBJECT, so a string is an error."""


def load_magicodec_model(name, device):
    raise RuntimeError(
        f"oracle shim: no checkpoint for {name!r} is available offline; pass a model object as codec_model=")
