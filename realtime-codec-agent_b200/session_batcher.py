"""Multi-session batching for the streaming path.

The reference serves several live streams through one codec: ``tts_server.py:59,158`` tokenizes every 0.1 s chunk
of every active TTS request on ONE ``AudioTokenizer`` shared by Flask threads, ``realtime_agent_resources.py:41-49``
hands one model to two agents, ``inference_client_self_play.py:148-159`` runs two agents side by side.  Alone, each
session is a batch-1 pass that streams every weight of the network for 100 rows; here the rolling contexts of up to
``max_sessions`` sessions live in one device pool (``mc_pool_*`` in the C ABI) and the sessions that are due at the
same time go through the engine as ONE ``B = n * channels`` launch.

``SessionBatcher`` keeps, per session, exactly the host-side state of ``AudioTokenizer`` (context length, code
string context, the ``[-0:]`` / hanging-code quirks of audio_tokenizer.py:99-101,144,161-168), so every session's
outputs are what its own ``AudioTokenizer`` would have returned.  Sessions whose contexts differ in length (a stream
that has just started next to one in steady state) are grouped by length: one launch per group.

``ThreadedSessionBatcher`` adds the tts_server.py calling pattern: request threads call
``tokenize_audio(session, chunk)`` and block; a dispatcher thread collects whatever arrived within ``max_wait_ms``
and runs it as one batch.
"""
from __future__ import annotations

import ctypes as C
import threading
from concurrent.futures import Future
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _native as nat
from .codec_chars import UNICODE_OFFSET_LARGE, chars_to_codes, codes_to_chars


class SessionPool:
    """Thin binding of mc_pool_*: slots with device-resident audio / code contexts, pushed in batches."""

    def __init__(self, gen, channels: int, context_samples: int, max_sessions: int, max_chunk_samples: int = 0):
        self.gen, self.channels, self.max_sessions = gen, channels, max_sessions
        self.cap_samples = max(context_samples, max_chunk_samples)
        self.cap_frames = -(-self.cap_samples // gen.hop)
        gen.ensure_positions(self.cap_frames)
        self._p = C.c_void_p()
        rc = gen._lib.mc_pool_create(gen._handle, channels, context_samples, self.cap_samples, max_sessions, C.byref(self._p))
        nat.check(gen._lib, gen._handle, rc, "mc_pool_create")

    def reset(self, slot: int, audio: bool = True, codes: bool = True) -> None:
        self.gen._lib.mc_pool_reset(self._p, slot, int(audio), int(codes))

    def context_len(self, slot: int) -> Tuple[int, int]:
        a, c = C.c_int32(), C.c_int32()
        self.gen._lib.mc_pool_context_len(self._p, slot, C.byref(a), C.byref(c))
        return a.value, c.value

    def set_graphs(self, enabled: bool) -> None:
        self.gen._lib.mc_pool_set_graphs(self._p, int(enabled))

    def push_audio(self, slots, chunks: np.ndarray, keep_frames: int) -> np.ndarray:
        """chunks float32 [n, C, len] -> int64 codes [n, C, keep]."""
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        chunks = np.ascontiguousarray(chunks, dtype=np.float32).reshape(len(slots), self.channels, -1)
        out = np.empty((len(slots) * self.channels * self.cap_frames,), dtype=np.int64)
        got = C.c_int32(0)
        with self.gen._serial():
            rc = self.gen._lib.mc_pool_push_audio(self._p, slots.ctypes.data, len(slots), chunks.ctypes.data, chunks.shape[2],
                                                  keep_frames, out.ctypes.data, C.byref(got), self.gen._stream())
            nat.check(self.gen._lib, self.gen._handle, rc, "mc_pool_push_audio")
        return out[: len(slots) * self.channels * got.value].reshape(len(slots), self.channels, got.value)

    def push_codes(self, slots, codes: np.ndarray, keep_samples: int) -> np.ndarray:
        """codes int64 [n, C, len] -> float32 wav [n, C, keep]."""
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        codes = np.ascontiguousarray(codes, dtype=np.int64).reshape(len(slots), self.channels, -1)
        out = np.empty((len(slots) * self.channels * self.cap_samples,), dtype=np.float32)
        got = C.c_int32(0)
        with self.gen._serial():
            rc = self.gen._lib.mc_pool_push_codes(self._p, slots.ctypes.data, len(slots), codes.ctypes.data, codes.shape[2],
                                                  keep_samples, out.ctypes.data, C.byref(got), self.gen._stream())
            nat.check(self.gen._lib, self.gen._handle, rc, "mc_pool_push_codes")
        return out[: len(slots) * self.channels * got.value].reshape(len(slots), self.channels, got.value)

    def __del__(self):
        try:
            if getattr(self, "_p", None) and self.gen._handle:
                self.gen._lib.mc_pool_destroy(self._p)
                self._p = None
        except Exception:
            pass


class _SessionState:
    def __init__(self, slot: int):
        self.slot = slot
        self.audio_len = 0            # samples in the rolling audio context (len(tokenize_context[-1]))
        self.detok_context = ""       # the code-string context (detokenize_context)


class SessionBatcher:
    """Batched ``tokenize_audio`` / ``detokenize_audio`` over many independent sessions of one engine."""

    def __init__(self, codec_model, num_channels: int = 1, context_secs: float = 2.0, max_sessions: int = 8,
                 unicode_offset: int = UNICODE_OFFSET_LARGE):
        if not getattr(codec_model, "is_b200_native", False):
            raise RuntimeError("SessionBatcher needs the B200 engine (B200Generator); there is no fallback")
        self.gen = codec_model
        self.num_channels, self.context_secs, self.unicode_offset = num_channels, context_secs, unicode_offset
        self.sampling_rate = codec_model.sample_rate
        self.framerate = self.sampling_rate / codec_model.hop
        self.codebook_size = codec_model.codebook_size
        self.context_samples = int(context_secs * self.sampling_rate)
        self.context_frames = int(context_secs * self.framerate * num_channels)
        self.pool = SessionPool(codec_model, num_channels, self.context_samples, max_sessions)
        self._free = list(range(max_sessions - 1, -1, -1))
        self._sessions: Dict[int, _SessionState] = {}
        self._next_id = 0
        self._lock = threading.RLock()

    # ---- session lifetime
    def open_session(self) -> int:
        with self._lock:
            if not self._free:
                raise RuntimeError(f"all {self.pool.max_sessions} session slots are in use")
            slot = self._free.pop()
            self.pool.reset(slot)
            sid = self._next_id
            self._next_id += 1
            self._sessions[sid] = _SessionState(slot)
            return sid

    def close_session(self, sid: int) -> None:
        with self._lock:
            self._free.append(self._sessions.pop(sid).slot)

    def reset_context(self, sid: int) -> None:
        with self._lock:
            st = self._sessions[sid]
            self.pool.reset(st.slot)
            st.audio_len, st.detok_context = 0, ""

    # ---- encode
    def tokenize_audio(self, chunks: Dict[int, np.ndarray]) -> Dict[int, str]:
        """{session: float32/int16 chunk [T] or [C,T] at the codec's rate} -> {session: code string}, each exactly what
        that session's AudioTokenizer.tokenize_audio (audio_tokenizer.py:67-103) would return."""
        C_ = self.num_channels
        with self._lock:
            groups: Dict[Tuple[int, int], List[int]] = {}
            prepped = {}
            for sid, chunk in chunks.items():                    # validate everything before any state changes
                if sid not in self._sessions:
                    raise KeyError(f"unknown session {sid}")
                x = np.asarray(chunk)
                if x.dtype == np.int16:
                    x = x.astype("float32") / 32768.0
                if C_ == 1 and x.ndim > 1:
                    x = np.mean(x, axis=0)
                x = np.ascontiguousarray(x, dtype=np.float32).reshape(C_, -1)
                if not 0 < x.shape[1] <= self.pool.cap_samples:
                    raise ValueError(f"session {sid}: chunks must hold 1..{self.pool.cap_samples} samples")
                prepped[sid] = x
                groups.setdefault((self._sessions[sid].audio_len, x.shape[1]), []).append(sid)
            out: Dict[int, str] = {}
            for (_, n_new), sids in groups.items():
                n_chars = int(n_new / self.sampling_rate * self.framerate * C_)
                frames_needed = -(-n_chars // C_) if n_chars > 0 else 0                  # 0 -> all frames ([-0:])
                slots = [self._sessions[s].slot for s in sids]
                codes = self.pool.push_audio(slots, np.stack([prepped[s] for s in sids]), frames_needed)   # [n,C,k]
                for j, sid in enumerate(sids):
                    st = self._sessions[sid]
                    st.audio_len = min(st.audio_len + n_new, max(n_new, self.context_samples))
                    frame_major = np.ascontiguousarray(codes[j].T).reshape(1, -1)
                    text = codes_to_chars(frame_major, self.codebook_size, unicode_offset=self.unicode_offset)
                    out[sid] = text[-n_chars:]
            return out

    # ---- decode
    def detokenize_audio(self, strings: Dict[int, str], preroll_samples: int = 0):
        """{session: code string} -> {session: ((sr, wav), end_hanging, preroll_left)} like
        AudioTokenizer.detokenize_audio (audio_tokenizer.py:105-149)."""
        C_ = self.num_channels
        with self._lock:
            groups: Dict[Tuple[int, int], List[int]] = {}
            new_codes, hanging, wants, trimmed = {}, {}, {}, {}
            for sid, s in strings.items():                       # validate everything before any state changes
                if sid not in self._sessions:
                    raise KeyError(f"unknown session {sid}")
                extra = len(s) % C_
                if extra:
                    s = s[:-extra]
                    hanging[sid] = s[-extra:]
                else:
                    hanging[sid] = ""
                if not 0 < len(s) // C_ <= self.pool.cap_frames:
                    raise ValueError(f"session {sid}: strings must hold 1..{self.pool.cap_frames} frames")
                flat = chars_to_codes(s, 1, self.codebook_size, unicode_offset=self.unicode_offset)[0]
                new_codes[sid] = np.ascontiguousarray(np.asarray(flat).reshape(-1, C_).T)
                trimmed[sid] = s
            for sid, s in trimmed.items():
                st = self._sessions[sid]
                st.detok_context = (st.detok_context + s)[-max(len(s), self.context_frames):]
                wants[sid] = int(len(s) / (self.framerate * C_) * self.sampling_rate) + preroll_samples
                ctx_frames_before = self.pool.context_len(st.slot)[1]
                groups.setdefault((ctx_frames_before, len(s) // C_, wants[sid]), []).append(sid)
            out = {}
            for (_, _, want), sids in groups.items():
                slots = [self._sessions[s].slot for s in sids]
                wav = self.pool.push_codes(slots, np.stack([new_codes[s] for s in sids]), want)   # [n,C,k]
                for j, sid in enumerate(sids):
                    w = wav[j]
                    preroll_left = max(0, preroll_samples - want + w.shape[-1])
                    out[sid] = ((self.sampling_rate, (w[0] if C_ == 1 else w).copy()), hanging[sid], preroll_left)
            return out


class ThreadedSessionBatcher(SessionBatcher):
    """tts_server.py's calling pattern: every request thread calls ``tokenize_audio_one(session, chunk)`` and blocks;
    the dispatcher thread runs everything that arrived within ``max_wait_ms`` as one batch."""

    def __init__(self, *args, max_wait_ms: float = 2.0, **kw):
        super().__init__(*args, **kw)
        self.max_wait = max_wait_ms / 1e3
        self._cv = threading.Condition()
        self._pending: List[Tuple[int, np.ndarray, Future]] = []
        self._stop = False
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def tokenize_audio_one(self, sid: int, chunk: np.ndarray, timeout: Optional[float] = 30.0) -> str:
        fut: Future = Future()
        with self._cv:
            self._pending.append((sid, chunk, fut))
            self._cv.notify()
        return fut.result(timeout=timeout)

    def shutdown(self) -> None:
        with self._cv:
            self._stop = True
            self._cv.notify()
        self._thread.join(timeout=5)

    def _run(self) -> None:
        import time
        while True:
            with self._cv:
                while not self._pending and not self._stop:
                    self._cv.wait()
                if self._stop and not self._pending:
                    return
            time.sleep(self.max_wait)                      # let the other streams' chunks of this tick arrive
            with self._cv:
                batch, rest, seen = [], [], set()
                for item in self._pending:                 # one chunk per session per batch, in arrival order
                    (batch if item[0] not in seen else rest).append(item)
                    seen.add(item[0])
                self._pending = rest
            try:
                res = self.tokenize_audio({sid: chunk for sid, chunk, _ in batch})
                for sid, _, fut in batch:
                    fut.set_result(res[sid])
            except (ValueError, KeyError):
                # a malformed request (bad chunk length, closed session) is rejected BEFORE anything is pushed: serve the
                # others, fail only the offender
                for sid, chunk, fut in batch:
                    try:
                        fut.set_result(self.tokenize_audio({sid: chunk})[sid])
                    except Exception as ex:                # noqa: BLE001
                        fut.set_exception(ex)
            except Exception as ex:                        # noqa: BLE001  (engine failure: state unknown, everyone hears about it)
                for _, _, fut in batch:
                    if not fut.done():
                        fut.set_exception(ex)
