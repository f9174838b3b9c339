"""GPU (B200), informative: what the reference's own GPU path would cost on this device — the oracle network run
as eager PyTorch under torch.autocast(bfloat16) (audio_tokenizer.py:24,78-82: cuBLAS linears, cuDNN convs, SDPA
attention, a materialised 131 072-wide distance matrix), default spec, the offline batch shape (256 windows x 2.0 s,
last 5 frames kept).  Printed next to the engine's time for the same batch (run pytest with -s; the log is committed
under profiles/).  The only assertion is that the engine is not slower.  SURVEY.md §8(d) "reference-on-GPU bar"."""
import pytest
import torch

import realtime_codec_agent_b200 as pkg
from oracle.magicodec_oracle import OracleGenerator

pytestmark = pytest.mark.gpu


def _time(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def test_engine_vs_eager_bf16_autocast_on_the_same_gpu():
    spec = pkg.DEFAULT_SPEC
    w = pkg.init_random_weights(spec, seed=0)
    B = 256
    wav = pkg.synth_audio(B * 1600 + 32000, device="cuda")
    windows = torch.stack([wav[b * 1600: b * 1600 + 32000] for b in range(B)])
    oracle = OracleGenerator(spec, w).cuda().eval()

    def eager():
        with torch.no_grad(), torch.autocast(device_type="cuda", dtype=torch.bfloat16):
            z = oracle.encoder(oracle.pad_audio(windows))
            return oracle.quantizer.inference(z[:, -5:].float())[1]

    gen = pkg.B200Generator(spec, w, device="cuda")

    def engine():
        return gen.encode(wav, keep_last_frames=5, row_stride=1600, num_windows=B, window_samples=32000)

    t_eager, t_engine = _time(eager, 3), _time(engine, 10)
    print(f"\n[reference-on-GPU bar] {B} windows x 2.0 s (25.6 s of new audio): eager bf16-autocast PyTorch {t_eager:.2f} ms "
          f"({B * 0.1 / t_eager * 1e3:.0f} audio-s/s) | engine {t_engine:.2f} ms ({B * 0.1 / t_engine * 1e3:.0f} audio-s/s) | "
          f"x{t_eager / t_engine:.1f}")
    assert t_engine < t_eager
