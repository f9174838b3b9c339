"""``ExternalTTSDuplexAligner`` on the B200 engine — same constructor and method as
/root/reference/realtime_codec_agent/external_tts_duplex_aligner.py:6-27.

The reference gathers rows of ``get_codec_embeddings()`` for the two token-id lists, takes each
row's distance to the mean embedding of 10 s of encoded silence and returns the ratio of the two
mean distances.  Here the gather + norm + mean is one kernel over the engine's cached projected
codebook (``mc_embed_distance``): two floats come back instead of a [2,n,16] tensor pipeline, and
the 131 072 x 16 table is never copied or re-projected.  Arithmetic is fp32 (the reference's GPU
path runs the same chain in bf16 under autocast; its CPU path is fp32 and is what the tests pin).
"""
from __future__ import annotations

from typing import List, Optional

import torch


class ExternalTTSDuplexAligner:
    def __init__(self, audio_tokenizer, duplex_model_dir: Optional[str] = None, codec_vocab_start: Optional[int] = None):
        if codec_vocab_start is None:
            from transformers import AutoConfig                     # external_tts_duplex_aligner.py:9-10
            codec_vocab_start = AutoConfig.from_pretrained(duplex_model_dir).codec_vocab_start
        self.codec_vocab_start = int(codec_vocab_start)
        self.audio_tokenizer = audio_tokenizer
        self._native = bool(getattr(audio_tokenizer, "_native", False))
        silence_codes = audio_tokenizer._encode_silence(10.0)[0, 0]              # :13
        if self._native:
            self._gen = audio_tokenizer.codec_model
            _, mean = self._gen.embed_distance(silence_codes[None], 0, None, want_mean=True)
            self.silence_embedding = mean[0]                                     # :14-15, fp32 on the device
            self.codec_embeddings = None                                         # never materialised
        else:
            self.codec_embeddings = audio_tokenizer.get_codec_embeddings()
            self.silence_embedding = torch.nn.functional.embedding(silence_codes, self.codec_embeddings).mean(0)

    def interrupt_score(self, tts_token_ids: List[int], duplex_token_ids: List[int]) -> float:
        """How many times further from silence the TTS prediction is than the duplex prediction (:17-27)."""
        ids = torch.tensor([tts_token_ids, duplex_token_ids])
        if self._native:
            dist = self._gen.embed_distance(ids, self.codec_vocab_start, self.silence_embedding)
            tts_dist, duplex_dist = dist.tolist()
        else:
            ids = ids.to(self.codec_embeddings.device) - self.codec_vocab_start
            rows = torch.nn.functional.embedding(ids, self.codec_embeddings)
            away = torch.linalg.vector_norm(rows - self.silence_embedding, dim=-1)
            tts_dist, duplex_dist = away.mean(dim=-1).tolist()
        return tts_dist / (duplex_dist + 1e-5)
