"""Test helper: a CPU stand-in with B200Generator's fast-entry signature, backed by the oracle, so
host-side logic (corpus planning, sharding, manifests) can be exercised without a GPU."""
import torch


class OracleBackedGen:
    is_b200_native = True

    def __init__(self, oracle):
        self.oracle = oracle
        self.sample_rate = oracle.sample_rate
        self.codebook_size = oracle.codebook_size
        self.hop = oracle.hop
        self.device = torch.device("cpu")
        self.calls = []

    def encode(self, wav, keep_last_frames=0, row_stride=None, num_windows=None, window_samples=None, **kw):
        if row_stride is not None:
            flat = wav.reshape(-1)
            wav = torch.stack([flat[b * row_stride: b * row_stride + window_samples] for b in range(num_windows)])
        elif wav.dim() == 1:
            wav = wav[None]
        self.calls.append(tuple(wav.shape))
        with torch.no_grad():
            z = self.oracle.encoder(self.oracle.pad_audio(wav.float()))
            idx = self.oracle.quantizer.inference(z)[1]
        return idx[:, -keep_last_frames:] if keep_last_frames > 0 else idx
