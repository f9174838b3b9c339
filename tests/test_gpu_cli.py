"""GPU (B200): the offline CLI end to end on the native engine (row f1): `python -m ...audio_to_codes` semantics of
encode_audio_gpu_1.sh / encode_audio_stereo.sh through main(), a checkpoint file, real .wav decoding, the on-disk
layout LMDatasetBuilder parses — and codes equal to AudioTokenizer.chunked_tokenize_audio on the same engine."""
import json
import os

import numpy as np
import pytest
import torch
from scipy.io import wavfile

import realtime_codec_agent_b200 as pkg
from realtime_codec_agent_b200 import audio_to_codes

pytestmark = pytest.mark.gpu


def test_cli_main_stereo_on_the_engine(tmp_path, monkeypatch):
    spec = pkg.TINY_SPEC
    w = pkg.init_random_weights(spec, seed=0)
    ckpt = tmp_path / "tiny.b200.pt"
    pkg.save_checkpoint(str(ckpt), spec, w)
    monkeypatch.setenv("MAGICODEC_B200_CHECKPOINT", str(ckpt))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    raw = tmp_path / "raw" / "CallFriend_eng"
    raw.mkdir(parents=True)
    a = pkg.synth_audio(16000 * 5 + 800, file_id=21).numpy()
    b = pkg.synth_audio(16000 * 5 + 800, file_id=22, channel=1).numpy()
    wavfile.write(raw / "0001.wav", 16000, (np.stack([a, b], axis=1) * 32767).astype(np.int16))
    wavfile.write(raw / "0002.wav", 8000, (a[::2] * 32767).astype(np.int16))           # mono, resampled 8 k -> 16 k
    audio_to_codes.main(["--audio_path", str(tmp_path / "raw"), "--codes_path", str(tmp_path / "codes"), "--stereo",
                         "--audio_filter", "CallFriend", "--batch_size", "16"])
    out = tmp_path / "codes" / "MagiCodec-50Hz-Base" / "0.1s_2.0s" / "stereo"
    info = json.load(open(out / "codec_info.json"))
    assert info["framerate"] == 50.0 and info["codebook_size"] == spec.codebook_size
    man = json.load(open(out / "manifest.json"))
    assert len(man) == 4
    gen = pkg.B200Generator(spec, w, device="cuda")
    tok = pkg.AudioTokenizer(codec_model=gen, device="cuda")
    st = audio_to_codes.load_audio(str(raw / "0001.wav"), 16000, mono=False)
    for c in range(2):
        arr = np.load(out / "CallFriend_eng" / f"0001_c{c}.npy")
        assert arr.shape == (1, 252) and arr.dtype == np.int32                         # 5.05 s -> 250 + int(0.05*50) = 252
        tok.reset_context()
        s = tok.chunked_tokenize_audio(st[c], 0.1)
        assert np.array_equal(arr[0], [ord(ch) - tok.unicode_offset for ch in s])
    assert np.load(out / "CallFriend_eng" / "0002_c0.npy").shape[-1] == np.load(out / "CallFriend_eng" / "0002_c1.npy").shape[-1]
