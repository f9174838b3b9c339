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
from tests.fake_gen import HostIngest, OracleBackedGen


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
    man, errs = audio_to_codes.encode_corpus(gen, str(tmp_path / "raw"), str(tmp_path / "codes"), stereo=True,
                                             audio_filter=["CallHome"], batch_size=16, ingest=HostIngest(gen))
    assert errs == []
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
    # resume: nothing left to encode, and the skipped files re-enter the manifest from disk
    calls = len(gen.calls)
    man2, errs2 = audio_to_codes.encode_corpus(gen, str(tmp_path / "raw"), str(tmp_path / "codes"), stereo=True,
                                               audio_filter=["CallHome"], ingest=HostIngest(gen))
    assert len(gen.calls) == calls and errs2 == []
    assert [(e.file_id, e.channel, e.n_frames, e.crc32, e.path) for e in man2] == [(e.file_id, e.channel, e.n_frames, e.crc32, e.path) for e in man]
    assert man[0].path == "CallHome_eng/4156"
    # a truncated output (run killed mid-write before the rename existed) is re-encoded, not trusted
    victim = out / "CallHome_eng/4156_c1.npy"
    victim.write_bytes(victim.read_bytes()[:70])
    man3, _ = audio_to_codes.encode_corpus(gen, str(tmp_path / "raw"), str(tmp_path / "codes"), stereo=True,
                                           audio_filter=["CallHome"], ingest=HostIngest(gen))
    assert len(gen.calls) > calls and np.load(victim).shape == (1, 100)
    assert [(e.file_id, e.channel, e.crc32) for e in man3] == [(e.file_id, e.channel, e.crc32) for e in man]
    assert not [f for d, _, fs in os.walk(out) for f in fs if ".tmp." in f]


def test_two_ranks_partition_the_corpus(tmp_path):
    """encode_audio_gpu_N.sh launches one process per GPU over the same corpus: with world_size 2 every file is encoded
    by exactly one rank (LPT by size), outputs equal the single-rank run, and the manifests are disjoint."""
    raw = tmp_path / "raw"
    raw.mkdir()
    for i, secs in enumerate((2.3, 0.7, 1.5, 3.1, 0.4)):
        np.save(raw / f"clip{i}.npy", pkg.synth_audio(int(secs * 16000), file_id=30 + i).numpy())
    gen = OracleBackedGen(OracleGenerator(pkg.TINY_SPEC, pkg.init_random_weights(pkg.TINY_SPEC, seed=0)))
    solo = audio_to_codes.encode_corpus(gen, str(raw), str(tmp_path / "solo"), batch_size=16, ingest=HostIngest(gen))[0]
    man = []
    for rank in range(2):
        man.append(audio_to_codes.encode_corpus(gen, str(raw), str(tmp_path / "duo"), batch_size=16, rank=rank, world_size=2,
                                                ingest=HostIngest(gen))[0])
    assert sorted(e.file_id for m in man for e in m) == sorted(e.file_id for e in solo) == list(range(5))
    assert not {e.file_id for e in man[0]} & {e.file_id for e in man[1]} and man[0] and man[1]
    assert {e.rank for e in man[1]} == {1}
    by_id = {e.file_id: e.crc32 for e in solo}
    assert all(by_id[e.file_id] == e.crc32 for m in man for e in m)
    for i in range(5):
        a = np.load(tmp_path / "solo" / "MagiCodec-50Hz-Base" / "0.1s_2.0s" / "mono" / f"clip{i}_c0.npy")
        b = np.load(tmp_path / "duo" / "MagiCodec-50Hz-Base" / "0.1s_2.0s" / "mono" / f"clip{i}_c0.npy")
        assert np.array_equal(a, b)


def test_mixed_corpus_formats_and_per_file_errors(tmp_path):
    """The reference's corpora: 8 kHz mu-law SPHERE telephone speech (stereo), FLAC (libri-light), wav — and an mp3 that
    cannot be decoded offline: it is reported per file and the run goes on.  Ranks are balanced on decoded duration."""
    import struct
    from realtime_codec_agent_b200 import audio_io
    from tests import flac_writer as fw
    raw = tmp_path / "raw"
    (raw / "fisher").mkdir(parents=True)
    (raw / "libri").mkdir()
    a = pkg.synth_audio(16000 * 2, file_id=5).numpy()
    # 8 kHz 2-channel mu-law SPHERE (1.5 s): every byte value is a valid sample
    rng = np.random.default_rng(0)
    ul = rng.integers(0, 255, size=(12000, 2), dtype=np.uint8)
    hdr = ("NIST_1A\n   1024\nsample_count -i 12000\nsample_n_bytes -i 1\nchannel_count -i 2\nsample_rate -i 8000\n"
           "sample_coding -s4 ulaw\nend_head\n").encode().ljust(1024, b" ")
    (raw / "fisher" / "fe_03_00001.sph").write_bytes(hdr + ul.tobytes())
    pcm16 = np.round(a[:16000] * 20000).astype(np.int64)[None]
    plan = lambda b, c: ("lpc", {"order": 8, "porder": 4})
    (raw / "libri" / "book.flac").write_bytes(fw.encode_flac(pcm16, 16000, 16, 4096, plan))
    wavfile.write(raw / "libri" / "clip.wav", 16000, (a * 32767).astype(np.int16))
    (raw / "fisher" / "broken.mp3").write_bytes(b"\xff\xfb\x90\x00" * 4000)
    gen = OracleBackedGen(OracleGenerator(pkg.TINY_SPEC, pkg.init_random_weights(pkg.TINY_SPEC, seed=0)))
    man, errs = audio_to_codes.encode_corpus(gen, str(raw), str(tmp_path / "codes"), stereo=True, batch_size=16, ingest=HostIngest(gen))
    assert len(errs) == 1 and errs[0]["file"].endswith("broken.mp3") and "no decoder" in errs[0]["error"]
    # corrupt inputs of the formats that ARE decoded here: a wav cut inside its fmt chunk, a FLAC with a flipped byte
    # (frame CRC) and an .npy that is not an array — one errors.json line each, nothing else changes
    bad = tmp_path / "bad"
    bad.mkdir()
    good_wav = (raw / "libri" / "clip.wav").read_bytes()
    (bad / "cut.wav").write_bytes(good_wav[:30])
    fl = bytearray((raw / "libri" / "book.flac").read_bytes())
    fl[len(fl) // 2] ^= 0x5A
    (bad / "flip.flac").write_bytes(bytes(fl))
    (bad / "junk.npy").write_bytes(b"not a numpy file")
    (bad / "ok.wav").write_bytes(good_wav)
    man_b, errs_b = audio_to_codes.encode_corpus(gen, str(bad), str(tmp_path / "codes_bad"), batch_size=16, ingest=HostIngest(gen))
    assert sorted(os.path.basename(e["file"]) for e in errs_b) == ["cut.wav", "flip.flac", "junk.npy"]
    assert [e.path for e in man_b] == ["ok"]
    out = tmp_path / "codes" / "MagiCodec-50Hz-Base" / "0.1s_2.0s" / "stereo"
    assert np.load(out / "fisher/fe_03_00001_c0.npy").shape == (1, 75)            # 1.5 s at 8 kHz -> 24 000 samples at 16 kHz
    assert np.load(out / "fisher/fe_03_00001_c1.npy").shape == (1, 75)
    assert np.load(out / "libri/book_c0.npy").shape == (1, 50)
    assert np.array_equal(np.load(out / "libri/book_c0.npy"), np.load(out / "libri/book_c1.npy"))   # mono file, --stereo
    assert sorted({e.path for e in man}) == ["fisher/fe_03_00001", "libri/book", "libri/clip"]
    # the SPHERE channel went through G.711 expansion + the polyphase resampler before the encoder
    ch0 = audio_io.load_audio(str(raw / "fisher" / "fe_03_00001.sph"), 16000, mono=False)[0]
    assert ch0.shape == (24000,)
    tok = pkg.AudioTokenizer(codec_model=gen.oracle, device="cpu")
    ref = tok.chunked_tokenize_audio(ch0, 0.1)
    assert np.array_equal(np.load(out / "fisher/fe_03_00001_c0.npy")[0], [ord(c) - tok.unicode_offset for c in ref])
    # duration-balanced sharding: the broken mp3 is small, the probe estimates it; every decodable file lands on one rank
    shards = [audio_to_codes.encode_corpus(gen, str(raw), str(tmp_path / "duo"), stereo=False, batch_size=16, rank=r, world_size=2,
                                           ingest=HostIngest(gen)) for r in range(2)]
    assert sorted(e.path for m, _ in shards for e in m) == ["fisher/fe_03_00001", "libri/book", "libri/clip"]
    assert sum(len(e) for _, e in shards) == 1


def test_tokenizer_resamples_with_the_soxr_class_filter():
    """_prep_audio_for_tokenization (audio_tokenizer.py:213-214) goes through audio_io.resample."""
    from realtime_codec_agent_b200 import audio_io
    gen = OracleGenerator(pkg.TINY_SPEC, pkg.init_random_weights(pkg.TINY_SPEC, seed=0))
    tok = pkg.AudioTokenizer(codec_model=gen, device="cpu")
    x = pkg.synth_audio(24000, file_id=8).numpy()
    got = tok._prep_audio_for_tokenization((24000, x))
    assert got.shape == (16000,) and np.array_equal(got, audio_io.resample(x, 24000, 16000))


def test_documented_command_line_resolves():
    """`python -m rca_b200_loader audio_to_codes ...` (the package directory name is not an identifier,
    so the repo-root loader module is the `-m` entry; torchrun takes the same `-m`)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "rca_b200_loader", "audio_to_codes", "--help"], cwd=root, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and "--audio_path" in r.stdout and "--stereo" in r.stdout
    r = subprocess.run([sys.executable, "-m", "rca_b200_loader", "no_such_module"], cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
