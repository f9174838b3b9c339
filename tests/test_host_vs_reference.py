"""CPU, build container only: the product AudioTokenizer and the UNMODIFIED reference wrapper,
each driving the same oracle model object, must agree call by call (strings equal, waveforms
bit-equal, bookkeeping equal) across the call patterns the reference's callers use
(realtime_agent_v2.py:504-579, run_stream_codes.py:14-68, tts_server.py:59)."""
import numpy as np
import pytest
import torch

import realtime_codec_agent_b200 as pkg
from oracle.magicodec_oracle import OracleGenerator
from oracle.reference_wrapper import load_reference_audio_tokenizer, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference is only mounted in the build container")


@pytest.fixture(scope="module")
def model():
    return OracleGenerator(pkg.TINY_SPEC, pkg.init_random_weights(pkg.TINY_SPEC, seed=0))


@pytest.mark.parametrize("channels", [1, 2])
@pytest.mark.parametrize("chunk_ms", [20, 100, 580, 1000])
def test_streaming_encode_decode_agree(model, channels, chunk_ms):
    Ref = load_reference_audio_tokenizer()
    ref = Ref(codec_model=model, num_channels=channels, device="cpu")
    ours = pkg.AudioTokenizer(codec_model=model, num_channels=channels, device="cpu")
    assert ours.framerate == ref.framerate and ours.context_frames == ref.context_frames
    n = chunk_ms * 16
    total = min(3.0, 12 * chunk_ms / 1000.0)
    wav = np.stack([pkg.synth_audio(int(total * 16000), file_id=3, channel=c).numpy() for c in range(channels)])
    wav = wav[0] if channels == 1 else wav
    pre_r = pre_o = 0
    for s0 in range(0, wav.shape[-1], n):
        chunk = wav[..., s0:s0 + n]
        a, b = ref.tokenize_audio(chunk), ours.tokenize_audio(chunk)
        assert a == b
        (sr_r, w_r), h_r, pre_r = ref.detokenize_audio(a, preroll_samples=320)
        (sr_o, w_o), h_o, pre_o = ours.detokenize_audio(b, preroll_samples=320)
        assert sr_r == sr_o and h_r == h_o and pre_r == pre_o
        assert w_r.shape == w_o.shape and np.array_equal(w_r, w_o)
    assert np.array_equal(ref.tokenize_context, ours.tokenize_context)
    assert ref.detokenize_context == ours.detokenize_context


def test_int16_tuple_mono_downmix_and_hanging(model):
    Ref = load_reference_audio_tokenizer()
    ref = Ref(codec_model=model, device="cpu")
    ours = pkg.AudioTokenizer(codec_model=model, device="cpu")
    st = np.stack([pkg.synth_audio(8000, channel=c).numpy() for c in range(2)])
    pcm = (st * 32767).astype(np.int16)
    assert ref.tokenize_audio((16000, pcm)) == ours.tokenize_audio((16000, pcm))
    ref2 = Ref(codec_model=model, num_channels=2, device="cpu")
    ours2 = pkg.AudioTokenizer(codec_model=model, num_channels=2, device="cpu")
    s = ref2.tokenize_audio(st)
    assert s == ours2.tokenize_audio(st)
    for cut in (len(s) - 1, len(s) - 3, 7):
        ref2.reset_context(); ours2.reset_context()
        (_, w_r), h_r, p_r = ref2.detokenize_audio(s[:cut], preroll_samples=100)
        (_, w_o), h_o, p_o = ours2.detokenize_audio(s[:cut], preroll_samples=100)
        assert h_r == h_o and p_r == p_o and np.array_equal(w_r, w_o)
    assert ref.get_audio_codes_str_secs(s) == ours.get_audio_codes_str_secs(s)
    assert torch.equal(ref.get_codec_embeddings(), ours.get_codec_embeddings())
    assert torch.equal(ref._encode_silence(1.0), ours._encode_silence(1.0))
    assert ref._drop_hanging_channel_codes("abc") == ours._drop_hanging_channel_codes("abc")


def test_codec_chars_match_shim():
    from oracle.shims.codec_bpe.core import converter as shim
    rng = np.random.default_rng(0)
    for nb, K in ((1, 131072), (2, 1024), (4, 2048)):
        codes = rng.integers(0, K, size=(nb, 37))
        for off in (pkg.UNICODE_OFFSET, pkg.UNICODE_OFFSET_LARGE):
            if off + nb * K > 0x110000:
                continue
            a = shim.codes_to_chars(codes, K, unicode_offset=off)
            assert a == pkg.codes_to_chars(codes, K, unicode_offset=off)
            assert np.array_equal(shim.chars_to_codes(a, nb, K, unicode_offset=off),
                                  pkg.chars_to_codes(a, nb, K, unicode_offset=off))
            assert np.array_equal(pkg.chars_to_codes(a, nb, K, unicode_offset=off), codes)
