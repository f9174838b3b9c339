"""Calibrates the stated parity tolerances: for each spec, encode windows of synthetic audio with
the CUDA engine and with the fp32 oracle (CPU), and report latent error, code agreement as a
function of the oracle's top-2 margin, and decode SNR.  Output: one JSON line per spec."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rca_b200_loader  # noqa: F401
import realtime_codec_agent_b200 as pkg
from oracle.magicodec_oracle import OracleGenerator


def sweep(name, spec, n_windows, few_rows=False, fuse_norm=1):
    """few_rows=True: every window alone through the streaming sessions' kernels (cluster split-K GEMMs, RMSNorms folded
    into GEMM epilogues; mc_set_option small_m_split_k = 2) instead of one batched launch."""
    torch.set_num_threads(os.cpu_count() or 1)
    w = pkg.init_random_weights(spec, seed=0)
    gen = pkg.B200Generator(spec, w, device="cuda", max_positions=512)
    oracle = OracleGenerator(spec, w)
    wav = torch.stack([pkg.synth_audio(32000, seed=77, file_id=i) for i in range(n_windows)])
    if few_rows:
        gen.set_option("small_m_split_k", 2)
        gen.set_option("fuse_norm", fuse_norm)
        name += f" [fuse_norm={fuse_norm}]"
        parts = [gen.encode(wav[i:i + 1].cuda(), return_margin=True, return_latents=True) for i in range(n_windows)]
        codes, margin_gpu, z_gpu = (torch.cat([p[k] for p in parts]) for k in range(3))
        name += " (few-rows kernels, one window per call)"
    else:
        codes, margin_gpu, z_gpu = gen.encode(wav.cuda(), return_margin=True, return_latents=True)
    with torch.no_grad():
        z_ref = oracle.encoder(oracle.pad_audio(wav))
        z_q, idx_ref, margin = oracle.quantizer.inference(z_ref, return_margin=True)
        rec_ref = oracle.decoder(z_q)[:, 0]
    if few_rows:
        rec = torch.cat([gen.decode(idx_ref[i:i + 1].cuda()) for i in range(n_windows)]).cpu()
    else:
        rec = gen.decode(idx_ref.cuda()).cpu()
    codes = codes.cpu()
    dz = (z_gpu.cpu() - z_ref)
    dis = codes != idx_ref
    m = margin.numpy().ravel()
    d = dis.numpy().ravel()
    out = {"spec": name, "frames": int(m.size), "z_err_max": float(dz.abs().max()), "z_err_rms": float(dz.pow(2).mean().sqrt()),
           "z_rms": float(z_ref.pow(2).mean().sqrt()), "agree_all": float(1 - d.mean()),
           "max_margin_of_disagreement": float(m[d].max()) if d.any() else 0.0,
           "margin_quantiles": {str(q): float(np.quantile(m, q)) for q in (0.01, 0.05, 0.1, 0.25, 0.5)},
           "near_tie_frac_at_eps": {str(e): float((m <= e).mean()) for e in (0.02, 0.05, 0.1, 0.2, 0.35)},
           "disagree_frac_above_eps": {str(e): float((d & (m > e)).mean()) for e in (0.02, 0.05, 0.1, 0.2, 0.35)},
           "decode_snr_db": float(10 * np.log10(rec_ref.pow(2).sum().item() / max((rec - rec_ref).pow(2).sum().item(), 1e-30))),
           "decode_max_abs_rel": float((rec - rec_ref).abs().max() / rec_ref.abs().max())}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    sweep("tiny", pkg.TINY_SPEC, 40)
    sweep("mid", pkg.MID_SPEC, 40)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    sweep("default", pkg.DEFAULT_SPEC, n)
    sweep("default", pkg.DEFAULT_SPEC, min(n, 100), few_rows=True)
    if "ablate" in sys.argv:
        sweep("default", pkg.DEFAULT_SPEC, min(n, 100), few_rows=True, fuse_norm=0)      # split-K kernels, standalone rmsnorm
        sweep("default", pkg.DEFAULT_SPEC, min(n, 100))                                   # the same 100 windows, batched path
