"""Offline corpus encode CLI — flag- and layout-compatible stand-in for the third-party
``python -m codec_bpe.audio_to_codes`` that /root/reference/encode_audio_gpu_{1..4}.sh:1-8 and
encode_audio_stereo.sh:1-9 invoke:

    python -m realtime_codec_agent_b200.audio_to_codes --audio_path data/audio/raw \
        --codes_path data/audio/codes --chunk_size_secs 0.1 --context_secs 2.0 --batch_size 256 \
        --codec_model MagiCodec-50Hz-Base [--stereo] [--audio_filter CallFriend CallHome ...]

Output (what prep_lm_dataset_magicodec{,_stereo}.sh:2 and LMDatasetBuilder read):
    <codes_path>/<codec_model>/<chunk>s_<context>s/{mono|stereo}/<relative path>_c<channel>.npy
        int32 array (num_codebooks=1, T)            lm_dataset_builder.py:79 (name), :396-408 (rank)
    <...>/{mono|stereo}/codec_info.json             keys framerate / num_codebooks / codebook_size
                                                    (prep_lm_dataset.py:47-52, tools/total_duration_codes.py:6-8)
Files whose outputs already exist are skipped (resume).  Under torchrun every rank takes its
duration-balanced share of the files (the reference shards by hand across four scripts) and rank 0
writes the merged manifest gathered over NCCL.
"""
from __future__ import annotations

import argparse
import json
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import corpus

SUPPORTED_EXTENSIONS = (".wav", ".npy", ".flac", ".mp3", ".ogg", ".opus", ".m4a", ".sph")


def get_files(root: str, extensions: Sequence[str] = SUPPORTED_EXTENSIONS, filters: Optional[Sequence[str]] = None) -> List[str]:
    out = []
    for d, _, files in os.walk(root):
        for f in files:
            p = os.path.join(d, f)
            if f.lower().endswith(tuple(extensions)) and (not filters or any(s in p for s in filters)):
                out.append(p)
    return sorted(out)


def load_audio(path: str, target_sr: int, mono: bool) -> np.ndarray:
    """-> float32 [C, T] at target_sr.  wav via scipy, npy raw arrays; other containers need soundfile."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        arr, sr = np.load(path), target_sr
        arr = arr[None] if arr.ndim == 1 else arr
    elif ext == ".wav":
        from scipy.io import wavfile
        sr, arr = wavfile.read(path)
        arr = arr[:, None] if arr.ndim == 1 else arr
        if np.issubdtype(arr.dtype, np.integer):
            arr = arr.astype(np.float32) / float(np.iinfo(arr.dtype).max + 1)
        arr = arr.T
    else:
        try:
            import soundfile as sf
        except ImportError as ex:
            raise RuntimeError(f"{path}: decoding {ext} needs the 'soundfile' package, which is not installed") from ex
        arr, sr = sf.read(path, dtype="float32", always_2d=True)
        arr = arr.T
    arr = np.asarray(arr, dtype=np.float32)
    if mono and arr.shape[0] > 1:
        arr = arr.mean(axis=0, keepdims=True)
    if sr != target_sr:
        from .audio_tokenizer import _resample
        arr = _resample(arr, sr, target_sr)
    return np.ascontiguousarray(arr)


def output_dir(codes_path: str, codec_model: str, chunk_secs: float, context_secs: float, stereo: bool) -> str:
    return os.path.join(codes_path, codec_model, f"{chunk_secs}s_{context_secs}s", "stereo" if stereo else "mono")


def encode_corpus(gen, audio_path: str, codes_path: str, codec_model: str = "MagiCodec-50Hz-Base",
                  chunk_size_secs: float = 0.1, context_secs: float = 2.0, batch_size: int = 256, stereo: bool = False,
                  audio_filter: Optional[Sequence[str]] = None, rank: int = 0, world_size: int = 1,
                  files_per_group: int = 8, overwrite: bool = False, fuse_batches: int = 4) -> List[corpus.ManifestEntry]:
    out_root = output_dir(codes_path, codec_model, chunk_size_secs, context_secs, stereo)
    files = get_files(audio_path, filters=audio_filter)
    sizes = [float(os.path.getsize(f)) for f in files]
    mine = corpus.shard_by_duration(sizes, world_size)[rank]
    os.makedirs(out_root, exist_ok=True)
    if rank == 0:
        info = {"codec_model": codec_model, "framerate": gen.sample_rate / gen.hop, "num_codebooks": 1,
                "codebook_size": gen.codebook_size, "sampling_rate": gen.sample_rate,
                "chunk_size_secs": chunk_size_secs, "context_secs": context_secs, "stereo": stereo}
        with open(os.path.join(out_root, "codec_info.json"), "w") as f:
            json.dump(info, f, indent=2)
    manifest: List[corpus.ManifestEntry] = []
    pending: List[Tuple[int, int, str, torch.Tensor]] = []

    def flush():
        if not pending:
            return
        codes = corpus.encode_streams(gen, [p[3] for p in pending], chunk_size_secs, context_secs, batch_size, fuse_batches)
        for (fid, ch, dst, stream), c in zip(pending, codes):
            np.save(dst, corpus.codes_to_array(c))
            n_win = -(-int(stream.numel()) // int(chunk_size_secs * gen.sample_rate))
            manifest.append(corpus.manifest_entry(fid, ch, c, n_win, rank))
        pending.clear()

    for fid in mine:
        rel = os.path.splitext(os.path.relpath(files[fid], audio_path))[0]
        chans = 2 if stereo else 1
        dsts = [os.path.join(out_root, f"{rel}_c{c}.npy") for c in range(chans)]
        if not overwrite and all(os.path.isfile(d) for d in dsts):
            continue
        wav = load_audio(files[fid], gen.sample_rate, mono=not stereo)
        if stereo and wav.shape[0] == 1:
            wav = np.concatenate([wav, wav], axis=0)
        os.makedirs(os.path.dirname(dsts[0]), exist_ok=True)
        for c in range(chans):
            pending.append((fid, c, dsts[c], torch.from_numpy(wav[c]).to(gen.device)))
        if len(pending) >= files_per_group:
            flush()
    flush()
    return manifest


def main(argv: Optional[Sequence[str]] = None) -> None:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--audio_path", required=True)
    ap.add_argument("--codes_path", required=True)
    ap.add_argument("--chunk_size_secs", type=float, default=0.1)
    ap.add_argument("--context_secs", type=float, default=2.0)
    ap.add_argument("--batch_size", type=int, default=256)
    ap.add_argument("--codec_model", default="MagiCodec-50Hz-Base")
    ap.add_argument("--stereo", action="store_true")
    ap.add_argument("--audio_filter", nargs="+")
    ap.add_argument("--overwrite", action="store_true")
    args = ap.parse_args(argv)

    import torch.distributed as dist
    from .audio_tokenizer import load_magicodec_model

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gen = load_magicodec_model(args.codec_model, dev)[0]
    local_manifest = encode_corpus(gen, args.audio_path, args.codes_path, args.codec_model, args.chunk_size_secs,
                                   args.context_secs, args.batch_size, args.stereo, args.audio_filter, rank, world,
                                   overwrite=args.overwrite)
    merged = corpus.gather_manifests(local_manifest, dev)
    if rank == 0:
        out_root = output_dir(args.codes_path, args.codec_model, args.chunk_size_secs, args.context_secs, args.stereo)
        with open(os.path.join(out_root, "manifest.json"), "w") as f:
            json.dump([e.__dict__ for e in merged], f)
        print(f"encoded {len(merged)} streams -> {out_root}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
