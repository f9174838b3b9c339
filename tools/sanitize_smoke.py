"""Small pass over every kernel family for compute-sanitizer (memcheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
TINY spec: batched encode / decode, a streaming session (split-K + fused-norm kernels, graphs off), the session pool,
device ingest (PCM conversion + resampler) and the emit chain."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rca_b200_loader  # noqa: F401
import realtime_codec_agent_b200 as pkg
from realtime_codec_agent_b200 import audio_io
from realtime_codec_agent_b200.session_batcher import SessionBatcher

spec = pkg.MID_SPEC if "mid" in sys.argv else pkg.TINY_SPEC
gen = pkg.B200Generator(spec, pkg.init_random_weights(spec, seed=0), device="cuda", max_positions=512)
wav = pkg.synth_audio(16000 * 6)
codes = gen.encode(wav[None, :32000].cuda())
gen.decode(codes)
gen.encode(torch.stack([wav[:32000], wav[16000:48000], wav[100:32100]]).cuda(), keep_last_frames=5)
tok = pkg.AudioTokenizer(codec_model=gen, device="cuda")
tok._stream_session().set_graphs(False)
for i in range(30):
    s = tok.tokenize_audio(wav[i * 1600:(i + 1) * 1600].numpy())
    tok.detokenize_audio(s, preroll_samples=320)
bat = SessionBatcher(gen, max_sessions=3)
bat.pool.set_graphs(False)
sids = [bat.open_session() for _ in range(3)]
for i in range(4):
    out = bat.tokenize_audio({sid: wav[(i + k) * 1600:(i + k + 1) * 1600].numpy() for k, sid in enumerate(sids)})
    bat.detokenize_audio(out, preroll_samples=320)
ing = audio_io.DeviceIngest(gen)
pcm = audio_io.PcmAudio(8000, 2, 4000, audio_io.PCM_ULAW, False, np.random.default_rng(0).integers(0, 255, 8000, dtype=np.uint8))
y = ing.to_device(pcm, mono=True)
torch.cuda.synchronize()
print("sanitize_smoke: ok", tuple(y.shape), gen.launch_count, "launches")
