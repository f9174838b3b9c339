"""CPU: the corpus CLI's on-disk layout is what the reference's consumers parse
(lm_dataset_builder.py:75-101 name regex, :396-408 array rank; prep_lm_dataset.py:47-52 codec_info)."""
import json
import os
import re

import numpy as np
import torch
from scipy.io import wavfile

import realtime_codec_agent_b200 as pkg
from oracle.magicodec_oracle import OracleGenerator
from realtime_codec_agent_b200 import audio_to_codes
from tests.fake_gen import OracleBackedGen


def test_cli_layout_and_resume(tmp_path):
    raw = tmp_path / "raw" / "CallHome_eng"
    raw.mkdir(parents=True)
    a = pkg.synth_audio(16000 * 2 + 800, file_id=1).numpy()
    b = pkg.synth_audio(16000 * 2, file_id=2, channel=1).numpy()
    wavfile.write(raw / "4156.wav", 16000, (np.stack([a[: len(b)], b], axis=1) * 32767).astype(np.int16))
    np.save(raw / "mono_clip.npy", a)
    (tmp_path / "raw" / "other").mkdir()
    np.save(tmp_path / "raw" / "other" / "skipme.npy", a)
    gen = OracleBackedGen(OracleGenerator(pkg.TINY_SPEC, pkg.init_random_weights(pkg.TINY_SPEC, seed=0)))
    man = audio_to_codes.encode_corpus(gen, str(tmp_path / "raw"), str(tmp_path / "codes"), stereo=True,
                                       audio_filter=["CallHome"], batch_size=16)
    out = tmp_path / "codes" / "MagiCodec-50Hz-Base" / "0.1s_2.0s" / "stereo"
    names = sorted(os.path.relpath(os.path.join(d, f), out) for d, _, fs in os.walk(out) for f in fs)
    assert names == ["CallHome_eng/4156_c0.npy", "CallHome_eng/4156_c1.npy", "CallHome_eng/mono_clip_c0.npy",
                     "CallHome_eng/mono_clip_c1.npy", "codec_info.json"]
    for n in names[:-1]:
        assert re.match(r"(.+)_c(\d+)[_.]", n)                       # lm_dataset_builder.py:79
        arr = np.load(out / n)
        assert arr.ndim == 2 and arr.shape[0] == 1 and arr.dtype == np.int32
    info = json.load(open(out / "codec_info.json"))
    assert info["framerate"] == 50.0 and info["num_codebooks"] == 1 and info["codebook_size"] == pkg.TINY_SPEC.codebook_size
    assert np.load(out / "CallHome_eng/4156_c0.npy").shape[-1] == 100       # 2.0 s -> 100 frames
    assert np.load(out / "CallHome_eng/mono_clip_c0.npy").shape[-1] == 100 + 2   # ragged 50 ms tail -> int(0.05*50) = 2
    assert len(man) == 4
    # same semantics as AudioTokenizer.chunked_tokenize_audio on the int16-decoded channel
    tok = pkg.AudioTokenizer(codec_model=gen.oracle, device="cpu")
    ch0 = audio_to_codes.load_audio(str(raw / "4156.wav"), 16000, mono=False)[0]
    ref = tok.chunked_tokenize_audio(ch0, 0.1)
    assert np.array_equal(np.load(out / "CallHome_eng/4156_c0.npy")[0], [ord(c) - tok.unicode_offset for c in ref])
    # resume: nothing left to do
    calls = len(gen.calls)
    assert audio_to_codes.encode_corpus(gen, str(tmp_path / "raw"), str(tmp_path / "codes"), stereo=True,
                                        audio_filter=["CallHome"]) == []
    assert len(gen.calls) == calls


def test_two_ranks_partition_the_corpus(tmp_path):
    """encode_audio_gpu_N.sh launches one process per GPU over the same corpus: with world_size 2 every file is encoded
    by exactly one rank (LPT by size), outputs equal the single-rank run, and the manifests are disjoint."""
    raw = tmp_path / "raw"
    raw.mkdir()
    for i, secs in enumerate((2.3, 0.7, 1.5, 3.1, 0.4)):
        np.save(raw / f"clip{i}.npy", pkg.synth_audio(int(secs * 16000), file_id=30 + i).numpy())
    gen = OracleBackedGen(OracleGenerator(pkg.TINY_SPEC, pkg.init_random_weights(pkg.TINY_SPEC, seed=0)))
    solo = audio_to_codes.encode_corpus(gen, str(raw), str(tmp_path / "solo"), batch_size=16)
    man = []
    for rank in range(2):
        man.append(audio_to_codes.encode_corpus(gen, str(raw), str(tmp_path / "duo"), batch_size=16, rank=rank, world_size=2))
    assert sorted(e.file_id for m in man for e in m) == sorted(e.file_id for e in solo) == list(range(5))
    assert not {e.file_id for e in man[0]} & {e.file_id for e in man[1]} and man[0] and man[1]
    assert {e.rank for e in man[1]} == {1}
    by_id = {e.file_id: e.crc32 for e in solo}
    assert all(by_id[e.file_id] == e.crc32 for m in man for e in m)
    for i in range(5):
        a = np.load(tmp_path / "solo" / "MagiCodec-50Hz-Base" / "0.1s_2.0s" / "mono" / f"clip{i}_c0.npy")
        b = np.load(tmp_path / "duo" / "MagiCodec-50Hz-Base" / "0.1s_2.0s" / "mono" / f"clip{i}_c0.npy")
        assert np.array_equal(a, b)
