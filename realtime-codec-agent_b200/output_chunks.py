"""``OutputChunkEmitter`` — the agent's output-chunk step as one object.

Mirrors ``RealtimeAgent.detokenize_output_chunk`` (/root/reference/realtime_codec_agent/
realtime_agent_v2.py:556-579) from the point where the LM tokens have become an audio-code string:
``detokenize_audio(str, preroll_samples=L)`` -> ``pad_or_trim`` -> ``normalize_audio_rms`` (when
``target_volume_rms > 0``) -> ``smooth_join`` with the previous chunk -> the chunk to emit, shifted
left by the fade length; ``audio_history_ch1`` is maintained exactly as the agent maintains it
(:569-570,576), so ``np.concatenate(emitter.audio_history_ch1)`` (:744) gives the same recording.

With a ``B200Generator`` behind the tokenizer the whole chain is ONE C-ABI call
(``mc_stream_push_codes_emit``): decoder, RMS normalisation and crossfade replay as one CUDA graph
and 2*chunk + L floats come back.  Any other model object is driven through ``detokenize_audio``
and the numpy utilities, call for call like the reference.
"""
from __future__ import annotations

from typing import List

import numpy as np

from .audio_utils import create_crossfade_ramps, normalize_audio_rms, pad_or_trim, smooth_join


class OutputChunkEmitter:
    def __init__(self, audio_tokenizer, chunk_size_secs: float = 0.1, chunk_fade_secs: float = 0.02,
                 target_volume_rms: float = 0.0, silence_rms_threshold: float = 0.003):
        if audio_tokenizer.num_channels != 1:
            raise ValueError("the agent's output chain is mono (pad_or_trim rejects [C,T] arrays)")
        self.audio_tokenizer = audio_tokenizer
        self.target_volume_rms = float(target_volume_rms)
        self.silence_rms_threshold = float(silence_rms_threshold)
        sr = audio_tokenizer.sampling_rate
        self.chunk_size_samples = int(chunk_size_secs * sr)            # realtime_agent_v2.py:129
        self.crossfade_ramps = create_crossfade_ramps(sr, fade_secs=chunk_fade_secs)   # :131
        if not 0 < self.crossfade_ramps[0] <= self.chunk_size_samples:
            # L = 0 makes the reference's own slices `[:-L]` / `[-n-L:-L]` empty (its length asserts then fire);
            # L > chunk is rejected by RealtimeAgentConfig (realtime_agent_config.py:58)
            raise ValueError("chunk_fade_secs must give 0 < fade samples <= chunk samples")
        self.audio_history_ch1: List[np.ndarray] = []
        self._native = bool(getattr(audio_tokenizer, "_native", False))
        self._armed = False

    def reset(self) -> None:
        """Forget the previous chunk (the agent clears audio_history_ch1 together with reset_context)."""
        self.audio_history_ch1 = []
        self._armed = False

    def emit(self, out_chunk_str: str) -> np.ndarray:
        L = self.crossfade_ramps[0]
        n = self.chunk_size_samples
        if self._native:
            return self._emit_native(out_chunk_str)
        tok = self.audio_tokenizer
        (_, out_chunk), _, preroll = tok.detokenize_audio(out_chunk_str, preroll_samples=L)
        out_chunk = pad_or_trim(out_chunk, n + preroll)
        if self.target_volume_rms > 0:
            out_chunk = normalize_audio_rms(out_chunk, target_rms=self.target_volume_rms,
                                            silence_rms_threshold=self.silence_rms_threshold)
        if self.audio_history_ch1:
            joined = smooth_join(self.audio_history_ch1[-1], out_chunk, *self.crossfade_ramps)
            assert joined.shape[-1] == 2 * n, f"joined_ch1_chunks must have length {2 * n}, but got {joined.shape[-1]}"
            self.audio_history_ch1[-1] = joined[:n]
            self.audio_history_ch1.append(joined[n:])
            return joined[-n - L:-L]
        self.audio_history_ch1.append(out_chunk)
        return pad_or_trim(out_chunk[:-L], n, pad_side="left")

    def _emit_native(self, out_chunk_str: str) -> np.ndarray:
        L, fade_in, _ = self.crossfade_ramps
        n = self.chunk_size_samples
        tok = self.audio_tokenizer
        if not self._armed:
            tok.arm_emit(n, L, self.target_volume_rms, self.silence_rms_threshold, fade_in)
            self._armed = True
        block, had_prev = tok.detokenize_audio_emit(out_chunk_str)
        emitted, cross, fresh = block[:n], block[n:n + L], block[n + L:]
        if had_prev != bool(self.audio_history_ch1):
            raise RuntimeError("emit chain state diverged from audio_history_ch1 (tokenizer context reset without emitter.reset()?)")
        if had_prev:
            prev = self.audio_history_ch1[-1]
            self.audio_history_ch1[-1] = np.concatenate((prev[:n - L], cross))
        self.audio_history_ch1.append(fresh.copy())
        return emitted.copy()
