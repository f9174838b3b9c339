"""CPU: the oracle restatement reproduces the golden vectors that the UNMODIFIED reference
wrapper produced around it (tests/golden/make_golden.py), and the product's host-side
AudioTokenizer reproduces the same strings/waveforms when it drives the same model object."""
import os

import numpy as np
import pytest
import torch

import realtime_codec_agent_b200 as pkg
from oracle.magicodec_oracle import OracleGenerator

SPECS = {"tiny": pkg.TINY_SPEC, "mid": pkg.MID_SPEC}


@pytest.fixture(scope="module", params=["tiny", "mid"])
def bundle(request, golden_dir):
    name = request.param
    spec = SPECS[name]
    g = np.load(os.path.join(golden_dir, f"golden_{name}.npz"))
    model = OracleGenerator(spec, pkg.init_random_weights(spec, seed=0))
    return name, spec, g, model


def _codes(s, tok):
    return np.array([ord(c) - tok.unicode_offset for c in s], dtype=np.int32)


def test_synth_audio_is_reproducible(bundle):
    _, spec, g, _ = bundle
    n = g["wav0"].shape[0]
    w = pkg.synth_audio(n, seed=1234, file_id=0, channel=0).numpy().astype(np.float16).astype(np.float32)
    assert np.array_equal(w, g["wav0"])


def test_oracle_taps_match_golden(bundle):
    _, spec, g, model = bundle
    x = torch.from_numpy(np.stack([g["wav0"][-32000:], g["wav1"][-32000:]]))
    with torch.no_grad():
        z_e = model.encoder(model.pad_audio(x))
        z_q, idx, margin = model.quantizer.inference(z_e, return_margin=True)
        rec = model.decoder(z_q)
    assert np.allclose(z_e.numpy(), g["tap_z_e"], atol=2e-5, rtol=1e-5)
    stable = g["tap_margin"] > 1e-3
    assert np.array_equal(idx.numpy()[stable], g["tap_idx"][stable])
    assert np.allclose(rec.numpy()[:, 0], g["tap_rec"], atol=1e-4, rtol=1e-4)


def test_host_tokenizer_reproduces_reference_strings(bundle):
    _, spec, g, model = bundle
    tok = pkg.AudioTokenizer(codec_model=model, device="cpu")
    assert tok.framerate == float(g["framerate"]) == 50.0
    assert tok.context_samples == 32000 and tok.context_frames == 100
    wav0 = g["wav0"]
    s = tok.tokenize_audio(wav0)
    assert np.array_equal(_codes(s, tok), g["mono_oneshot_codes"])
    (sr, rec), hang, pre = tok.detokenize_audio(s)
    assert sr == 16000 and hang == "" and pre == 0
    assert np.allclose(rec, g["mono_oneshot_wav"], atol=1e-4)
    tok.reset_context()
    s = tok.chunked_tokenize_audio(wav0, 0.1)
    assert np.array_equal(_codes(s, tok), g["mono_chunked_codes"])
    tok.reset_context()
    pieces = []
    for i in range(0, len(s), 5):
        (sr, rec), hang, pre_left = tok.detokenize_audio(s[i:i + 5], preroll_samples=320)
        assert rec.shape[-1] == min(1920, (i // 5 + 1) * 1600)       # preroll available after the 1st chunk
        pieces.append(rec[-1600:])
    assert np.allclose(np.concatenate(pieces), g["mono_stream_decode_wav"], atol=1e-4)
    tok.reset_context()
    assert len(tok.tokenize_audio(wav0[: int(0.58 * 16000)])) == int(g["len_058"]) == 28
    assert len(tok.tokenize_audio(wav0[:320])) == int(g["len_002"]) == 1


def test_host_tokenizer_stereo(bundle):
    _, spec, g, model = bundle
    tok = pkg.AudioTokenizer(codec_model=model, num_channels=2, device="cpu")
    assert tok.context_frames == 200
    st = np.stack([g["wav0"], g["wav1"]])
    s = tok.chunked_tokenize_audio(st, 0.1)
    assert np.array_equal(_codes(s, tok), g["stereo_chunked_codes"])
    (sr, rec), hang, pre = tok.detokenize_audio(s[:201])
    assert [ord(c) for c in hang] == list(g["stereo_hanging"])
    assert rec.shape == g["stereo_decode_wav"].shape
    assert np.allclose(rec, g["stereo_decode_wav"], atol=1e-4)
