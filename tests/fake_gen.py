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


class HostIngest:
    """The corpus pipeline's staging interface (audio_io.DeviceIngest) for the CPU tests: numpy / scipy mirrors of
    the ingest kernels around the oracle-backed model.  Test infrastructure only."""

    def __init__(self, gen):
        self.gen = gen

    def host_buffer(self, nbytes):
        return torch.empty(nbytes, dtype=torch.uint8)

    def upload(self, pcm, host_tensor=None):
        return pcm

    def wait_uploaded(self, staged):
        pass

    def convert(self, pcm, staged, mono):
        from realtime_codec_agent_b200 import audio_io
        x = audio_io.pcm_to_float(pcm, mono=mono)
        if pcm.sample_rate != self.gen.sample_rate:
            x = audio_io.resample(x, pcm.sample_rate, self.gen.sample_rate)
        return torch.from_numpy(x.copy())

    def codes_to_host(self, codes):
        return [c.clone() for c in codes], (lambda: None)
