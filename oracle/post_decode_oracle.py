"""ORACLE (test infrastructure — only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import this): CPU restatement of the two rows that follow the decoder.

* ``ref_*``                numpy restatement of /root/reference/realtime_codec_agent/utils/audio_utils.py
                           (smooth_join :4-17, create_crossfade_ramps :19-23, pad_or_trim :25-37,
                           normalize_audio_rms :39-46)
* ``OracleOutputChain``    restatement of RealtimeAgent.detokenize_output_chunk, realtime_agent_v2.py:556-579,
                           over any tokenizer object with ``detokenize_audio``
* ``oracle_interrupt_score``  restatement of ExternalTTSDuplexAligner, external_tts_duplex_aligner.py:8-27

PINNED: in the build container ``load_reference_module`` imports the reference's own files UNMODIFIED
(by path, under a stand-in package name so that realtime_codec_agent/__init__.py:1-5 — which needs
llama_cpp — never runs); tests/golden/make_golden_post.py drives them to produce
tests/golden/golden_post.npz and tests/test_post_decode_host.py checks the restatement against both.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REFERENCE_PKG_DIR = "/root/reference/realtime_codec_agent"
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_STANDIN = "_reference_pkg"


def reference_available() -> bool:
    return os.path.isdir(REFERENCE_PKG_DIR)


def load_reference_module(dotted: str):
    """Import realtime_codec_agent.<dotted> from /root/reference without executing the package __init__."""
    if not reference_available():
        raise FileNotFoundError(REFERENCE_PKG_DIR)
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    if _STANDIN not in sys.modules:
        pkg = types.ModuleType(_STANDIN)
        pkg.__path__ = [REFERENCE_PKG_DIR]
        sys.modules[_STANDIN] = pkg
    return importlib.import_module(f"{_STANDIN}.{dotted}")


# ------------------------------------------------------------------ utils/audio_utils.py
def ref_create_crossfade_ramps(sr, fade_secs):                      # :19-23
    L = int(sr * fade_secs)
    fade_in = np.sin(0.5 * np.pi * np.linspace(0, 1, L, endpoint=False, dtype=np.float32))
    return L, fade_in, fade_in[::-1]


def ref_smooth_join(a, b, L, fade_in, fade_out):                    # :4-17
    if a.shape[-1] == 0:
        return b
    if L == 0:
        return np.concatenate((a, b), axis=-1)
    cross = a[..., -L:] * fade_out + b[..., :L] * fade_in
    return np.concatenate((a[..., :-L], cross, b[..., L:]), axis=-1)


def ref_pad_or_trim(x, n, pad_side="right"):                        # :25-37
    if x.ndim > 1:
        raise ValueError("Input chunk must be a 1D array.")
    if x.shape[-1] < n:
        w = n - x.shape[-1]
        return np.pad(x, (0, w) if pad_side == "right" else (w, 0), mode="constant")
    return x[..., :n] if x.shape[-1] > n else x


def ref_normalize_audio_rms(x, target_rms=0.05, silence_rms_threshold=0.003):   # :39-46
    rms = np.sqrt(np.mean(x ** 2))
    return x if rms < silence_rms_threshold else x * (target_rms / rms)


RESTATED_UTILS = {"create_crossfade_ramps": ref_create_crossfade_ramps, "smooth_join": ref_smooth_join,
                  "pad_or_trim": ref_pad_or_trim, "normalize_audio_rms": ref_normalize_audio_rms}


# ---------------------------------------------------- realtime_agent_v2.py:556-579
class OracleOutputChain:
    """State = audio_history_ch1 (:106); ``utils`` = dict of the four functions (restated or the reference's own)."""

    def __init__(self, audio_tokenizer, chunk_size_secs=0.1, chunk_fade_secs=0.02, target_volume_rms=0.0, utils=None):
        self.u = utils or RESTATED_UTILS
        self.tok = audio_tokenizer
        self.chunk = int(chunk_size_secs * audio_tokenizer.sampling_rate)                     # :129
        self.ramps = self.u["create_crossfade_ramps"](audio_tokenizer.sampling_rate, fade_secs=chunk_fade_secs)  # :131
        self.target = target_volume_rms
        self.history = []

    def step(self, out_chunk_str):
        L = self.ramps[0]
        (_, out), _, pre = self.tok.detokenize_audio(out_chunk_str, preroll_samples=L)      # :560
        out = self.u["pad_or_trim"](out, self.chunk + pre)                                     # :561
        if self.target > 0:
            out = self.u["normalize_audio_rms"](out, target_rms=self.target)                   # :562-563
        if len(self.history) > 0:                                                              # :564
            joined = self.u["smooth_join"](self.history[-1], out, *self.ramps)
            assert joined.shape[-1] == 2 * self.chunk
            self.history[-1] = joined[:self.chunk]
            self.history.append(joined[self.chunk:])
            return joined[-self.chunk - L:-L]                                                  # :574
        self.history.append(out)
        return self.u["pad_or_trim"](out[:-L], self.chunk, pad_side="left")                    # :578


# ---------------------------------------------- external_tts_duplex_aligner.py:8-27
def oracle_silence_embedding(codec_embeddings: torch.Tensor, silence_codes: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.embedding(silence_codes, codec_embeddings).mean(0)             # :13-15


def oracle_interrupt_score(codec_embeddings, silence_embedding, codec_vocab_start, tts_token_ids, duplex_token_ids):
    codes = torch.tensor([tts_token_ids, duplex_token_ids]) - codec_vocab_start               # :18-19
    embs = torch.nn.functional.embedding(codes, codec_embeddings)
    dist = torch.linalg.vector_norm(embs - silence_embedding, dim=-1).mean(dim=-1).tolist()   # :21-22
    return dist[0] / (dist[1] + 1e-5)                                                         # :26
