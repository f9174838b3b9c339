"""Generates tests/golden/golden_post.npz with the reference's OWN post-decode code, imported unmodified
from /root/reference (oracle/post_decode_oracle.py: load_reference_module):

* utils/audio_utils.py functions on seeded inputs (ramps, joins, pad/trim, RMS normalisation);
* the detokenize_output_chunk chain (realtime_agent_v2.py:556-579) enacted with those functions over
  a stub tokenizer that replays fixed decoder outputs — so the fixture pins the chain itself,
  independent of any network;
* ExternalTTSDuplexAligner (external_tts_duplex_aligner.py) driven through its real class with a
  stub tokenizer and a transformers config directory.

Run in the build container only:   python tests/golden/make_golden_post.py
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.post_decode_oracle import OracleOutputChain, load_reference_module  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SR, CHUNK, L = 16000, 1600, 320


class ReplayTokenizer:
    """detokenize_audio stub: returns prepared (wav, preroll) pairs in order."""
    sampling_rate = SR

    def __init__(self, decoded):
        self.decoded, self.i = decoded, 0

    def detokenize_audio(self, s, preroll_samples=0):
        wav = self.decoded[self.i]
        self.i += 1
        return (SR, wav), "", wav.shape[-1] - CHUNK


class AlignerStubTokenizer:
    def __init__(self, table, silence_codes):
        self.table, self.silence = table, silence_codes

    def get_codec_embeddings(self):
        return self.table

    def _encode_silence(self, secs):
        return self.silence[None, None]


def main():
    au = load_reference_module("utils.audio_utils")
    utils = {k: getattr(au, k) for k in ("create_crossfade_ramps", "smooth_join", "pad_or_trim", "normalize_audio_rms")}
    rng = np.random.default_rng(20261018)
    out = {}
    Lr, fi, fo = au.create_crossfade_ramps(SR, 0.02)
    out["ramp_L"], out["fade_in"], out["fade_out"] = np.int64(Lr), fi, np.ascontiguousarray(fo)
    a = rng.standard_normal(CHUNK).astype(np.float32) * 0.1
    b = rng.standard_normal(CHUNK + L).astype(np.float32) * 0.1
    out["join_a"], out["join_b"] = a, b
    out["join_out"] = au.smooth_join(a, b, Lr, fi, fo)
    out["pad_right"] = au.pad_or_trim(a[:1000], 1600)
    out["pad_left"] = au.pad_or_trim(a[:1000], 1600, pad_side="left")
    out["trim"] = au.pad_or_trim(b, 1600)
    out["norm_loud"] = au.normalize_audio_rms(b, target_rms=0.05)
    out["norm_silent"] = au.normalize_audio_rms(b * 1e-3, target_rms=0.05)
    # the chain: 12 chunks; chunk 0 arrives without preroll, chunk 5 is near-silent (below the RMS threshold)
    decoded = []
    for i in range(12):
        n = CHUNK if i == 0 else CHUNK + L
        w = (rng.standard_normal(n) * (0.002 if i == 5 else 0.08)).astype(np.float32)
        decoded.append(w)
    out["chain_decoded"] = np.concatenate(decoded)
    out["chain_lengths"] = np.array([d.shape[0] for d in decoded], dtype=np.int64)
    for tag, target in (("plain", 0.0), ("rms", 0.05)):
        chain = OracleOutputChain(ReplayTokenizer(decoded), 0.1, 0.02, target, utils=utils)
        emitted = [chain.step("") for _ in decoded]
        assert all(e.shape[-1] == CHUNK for e in emitted)
        out[f"chain_emitted_{tag}"] = np.stack(emitted)
        out[f"chain_history_{tag}"] = np.concatenate(chain.history)

    # aligner: real reference class
    ali = load_reference_module("external_tts_duplex_aligner")
    table = torch.from_numpy(rng.standard_normal((4096, 16)).astype(np.float32))
    silence = torch.from_numpy(rng.integers(0, 4096, size=500))
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "config.json"), "w") as f:
            json.dump({"model_type": "llama", "codec_vocab_start": 128256}, f)
        al = ali.ExternalTTSDuplexAligner(AlignerStubTokenizer(table, silence), d)
    tts = rng.integers(0, 4096, size=(6, 25)) + 128256
    dup = rng.integers(0, 4096, size=(6, 25)) + 128256
    out["ali_table"], out["ali_silence_codes"] = table.numpy(), silence.numpy()
    out["ali_silence_embedding"] = al.silence_embedding.numpy()
    out["ali_tts"], out["ali_duplex"] = tts, dup
    out["ali_scores"] = np.array([al.interrupt_score(t.tolist(), u.tolist()) for t, u in zip(tts, dup)], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "golden_post.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
