"""CPU: the packed weight layouts and the engine's buffer conventions (tests/packed_emulator.py
replays engine.cu with torch ops) reproduce the oracle when packed in fp32, and stay within bf16
round-off when packed the way the kernels consume them."""
import numpy as np
import pytest
import torch

import realtime_codec_agent_b200 as pkg
from oracle.magicodec_oracle import OracleGenerator
from realtime_codec_agent_b200.generator import pack_weights, split_bf16
from tests import packed_emulator as emu


@pytest.fixture(scope="module")
def setup():
    spec = pkg.TINY_SPEC
    w = pkg.init_random_weights(spec, seed=0)
    model = OracleGenerator(spec, w)
    wav = torch.stack([pkg.synth_audio(8000 + 37, file_id=5, channel=c) for c in range(2)])   # ragged: not a hop multiple
    return spec, w, model, wav


def test_fp32_packing_matches_oracle_exactly(setup):
    spec, w, model, wav = setup
    p = pack_weights(spec, w, max_positions=64, gemm_dtype=torch.float32)
    with torch.no_grad():
        z_ref = model.encoder(model.pad_audio(wav))
        z_q, idx = model.quantizer.inference(z_ref)
        rec_ref = model.decoder(z_q)[:, 0]
    z = emu.encode(spec, p, wav)
    assert z.shape == z_ref.shape
    assert torch.allclose(z, z_ref, atol=3e-5, rtol=1e-4)
    rec = emu.decode(spec, p, idx)
    assert torch.allclose(rec, rec_ref, atol=3e-5, rtol=1e-4)


def test_bf16_packing_is_within_roundoff(setup):
    spec, w, model, wav = setup
    p = pack_weights(spec, w, max_positions=64)
    with torch.no_grad():
        z_ref = model.encoder(model.pad_audio(wav))
    z = emu.encode(spec, p, wav)
    assert (z - z_ref).abs().max() < 0.08 * z_ref.abs().max()


def test_packed_vq_rows_reproduce_fp32_distances(setup):
    spec, w, model, wav = setup
    p = pack_weights(spec, w, max_positions=64)
    torch.manual_seed(1)
    z = torch.randn(64, 16)
    cb = p["vq.codebook"]
    exact = (cb.double().pow(2).sum(-1)[None] - 2.0 * z.double() @ cb.double().t())
    approx = emu.vq_scores_packed(p, z)
    assert (approx - exact).abs().max() < 1e-3            # dropped z_lo.c_lo term: ~2^-16 |z||c|
    assert torch.equal(approx.argmin(-1), exact.argmin(-1))
    with torch.no_grad():
        assert torch.equal(model.quantizer.projected(), cb)          # bit-identical projection


def test_split_bf16_reconstructs():
    x = torch.randn(1000) * 37.0
    a, b, c = split_bf16(x, 3)
    assert (a.float() + b.float() + c.float() - x).abs().max() <= 1e-6 * x.abs().max()


def test_c_abi_library_exports_every_declared_symbol():
    """The .so loads on a CPU-only box and exports each symbol of include/magicodec_b200.h."""
    import ctypes, os, re
    from realtime_codec_agent_b200 import _native as nat
    from realtime_codec_agent_b200.build import build
    path = build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "magicodec_b200.h")).read()
    declared = set(re.findall(r"\b(mc_[a-z_0-9]+)\s*\(", header))
    assert declared == set(nat.SYMBOLS), declared ^ set(nat.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mc_version() == 100
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            pkg.AudioTokenizer("MagiCodec-50Hz-Base")            # product path must fail loudly without a GPU


def test_checkpoint_roundtrip_is_weights_only(tmp_path):
    """save_checkpoint writes plain containers + tensors; load_checkpoint never unpickles code (weights_only=True)."""
    import pickle

    import pytest
    import torch

    import realtime_codec_agent_b200 as pkg

    spec = pkg.TINY_SPEC
    w = pkg.init_random_weights(spec, seed=3)
    path = tmp_path / "tiny.pt"
    pkg.save_checkpoint(str(path), spec, w)
    spec2, w2 = pkg.load_checkpoint(str(path))
    assert spec2 == spec and set(w2) == set(w)
    assert all(torch.equal(w[k], w2[k]) for k in w)

    class Evil:
        def __reduce__(self):
            return (print, ("code ran while loading a checkpoint",))

    bad = tmp_path / "evil.pt"
    torch.save({"spec": {}, "weights": {}, "x": Evil()}, str(bad))
    with pytest.raises((pickle.UnpicklingError, RuntimeError)):
        pkg.load_checkpoint(str(bad))
