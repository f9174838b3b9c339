"""One batch of the bench workload (256 windows x 2.0 s, keep 5 frames) run twice: the command ncu wraps.
Also prints per-class device time from the engine's own event profile."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rca_b200_loader  # noqa: F401
import realtime_codec_agent_b200 as pkg

spec = pkg.DEFAULT_SPEC
gen = pkg.B200Generator(spec, pkg.init_random_weights(spec, seed=0), device="cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
wav = pkg.synth_audio(B * 1600 + 32000, device="cuda")
for it in range(2):
    gen.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    codes = gen.encode(wav, keep_last_frames=5, row_stride=1600, num_windows=B, window_samples=32000)
    e1.record()
    torch.cuda.synchronize()
    prof = gen.profile_end()
    print(f"iter {it}: {e0.elapsed_time(e1):.3f} ms for {B} windows; " +
          "; ".join(f"{k}: {v['ms']:.3f} ms / {v['launches']} launches" + (f" / {v['flops'] / v['ms'] / 1e9:.0f} TFLOP/s" if v['flops'] and v['ms'] else "")
                    for k, v in prof.items()))
print("checksum", int(codes.sum()), "audio_hash", float(wav.double().sum()), float(wav.double().abs().sum()))
again = gen.encode(wav, keep_last_frames=5, row_stride=1600, num_windows=B, window_samples=32000)
full = gen.encode(wav, keep_last_frames=0, row_stride=1600, num_windows=B, window_samples=32000)
print("repeat_equal", bool(torch.equal(codes, again)), "full_vs_keep5_equal", bool(torch.equal(full[:, -5:], codes)),
      "n_diff", int((full[:, -5:] != codes).sum()))
