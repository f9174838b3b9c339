"""Deterministic synthetic 16 kHz audio (SURVEY.md §8d "Synthetic audio").

Speech-like: a few harmonics of a slowly varying f0 under a syllabic AM envelope, plus low
level noise, with exact-zero silence gaps, so both active and silent frames are exercised.
Same (seed, file_id, channel) -> same bytes on any device of the same torch build when
generated on CPU; the bench generates on the GPU (values differ, statistics do not).
"""
from __future__ import annotations

import math

import torch


def synth_audio(num_samples: int, seed: int = 1234, file_id: int = 0, channel: int = 0,
                sample_rate: int = 16000, device: str | torch.device = "cpu") -> torch.Tensor:
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed((seed * 1000003 + file_id * 7919 + channel * 104729) & 0x7FFFFFFF)
    n = int(num_samples)
    t = torch.arange(n, device=dev, dtype=torch.float32) / sample_rate
    # control signals on a 100 Hz grid, linearly interpolated
    n_ctl = n // 160 + 2
    ctl = torch.rand((4, n_ctl), generator=g, device=dev)
    smooth = torch.nn.functional.avg_pool1d(ctl[None], 25, stride=1, padding=12, count_include_pad=False)[0]

    def up(x):
        return torch.nn.functional.interpolate(x[None, None], size=n, mode="linear", align_corners=True)[0, 0]

    f0 = 80.0 + 220.0 * up(smooth[0])
    phase = 2.0 * math.pi * torch.cumsum(f0, 0) / sample_rate
    amps = (1.0, 0.5, 0.33, 0.2, 0.12)
    sig = sum(a * torch.sin((k + 1) * phase + 0.7 * k) for k, a in enumerate(amps))
    syll = 0.5 * (1.0 + torch.sin(2.0 * math.pi * (3.0 + 3.0 * up(smooth[1])) * t))
    noise = torch.randn(n, generator=g, device=dev) * 0.03
    x = 0.35 * sig * syll + noise
    # silence gaps: ~40 % of 0.25 s segments are exact zeros (complementary across channels)
    seg = 4000
    n_seg = n // seg + 1
    gate = (torch.rand(n_seg, generator=g, device=dev) > 0.4)
    if channel % 2 == 1:
        gate = ~gate | (torch.rand(n_seg, generator=g, device=dev) > 0.8)
    gate = gate.float().repeat_interleave(seg)[:n]
    return (x * gate).clamp_(-1.0, 1.0).contiguous()
