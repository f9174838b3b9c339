"""Offline corpus encode CLI — flag- and layout-compatible stand-in for the third-party
``python -m codec_bpe.audio_to_codes`` that /root/reference/encode_audio_gpu_{1..4}.sh:1-8 and
encode_audio_stereo.sh:1-9 invoke:

    python -m rca_b200_loader audio_to_codes --audio_path data/audio/raw \
        --codes_path data/audio/codes --chunk_size_secs 0.1 --context_secs 2.0 --batch_size 256 \
        --codec_model MagiCodec-50Hz-Base [--stereo] [--audio_filter CallFriend CallHome ...]

Output (what prep_lm_dataset_magicodec{,_stereo}.sh:2 and LMDatasetBuilder read):
    <codes_path>/<codec_model>/<chunk>s_<context>s/{mono|stereo}/<relative path>_c<channel>.npy
        int32 array (num_codebooks=1, T)            lm_dataset_builder.py:79 (name), :396-408 (rank)
    <...>/{mono|stereo}/codec_info.json             keys framerate / num_codebooks / codebook_size
                                                    (prep_lm_dataset.py:47-52, tools/total_duration_codes.py:6-8)
    <...>/{mono|stereo}/manifest.json, errors.json  per-stream frames / windows / crc32; files that could not be decoded

The job is a pipeline, because one B200 encodes a 10-minute file in ~0.1 s:

    loader threads   container parsing / FLAC decode straight into pinned host buffers (audio_io.read_audio)
    copy stream      H2D of the raw PCM payload (1-2 bytes per sample) of file i+1 under the encode of file i
    main stream      mc_op_pcm_to_f32 + mc_op_resample (device ingest) -> corpus.encode_streams -> int32 codes
    writer thread    D2H through pinned memory, crc32, atomic .npy write (tmp + rename)

Files whose outputs already exist (and load) are skipped and re-enter the manifest from disk (resume).  Under
torchrun every rank takes its share of the files, balanced on DECODED duration x channels (header probe), and rank 0
writes the merged manifest gathered over NCCL.
"""
from __future__ import annotations

import argparse
import json
import os
import queue
import threading
import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import audio_io, corpus
from .audio_io import SUPPORTED_EXTENSIONS, UnsupportedAudio, load_audio  # noqa: F401  (re-exported: prep_channel_map.py:7-8)


def get_files(root: str, extensions: Sequence[str] = SUPPORTED_EXTENSIONS, filters: Optional[Sequence[str]] = None) -> List[str]:
    out = []
    for d, _, files in os.walk(root):
        for f in files:
            p = os.path.join(d, f)
            if f.lower().endswith(tuple(extensions)) and (not filters or any(s in p for s in filters)):
                out.append(p)
    return sorted(out)


def output_dir(codes_path: str, codec_model: str, chunk_secs: float, context_secs: float, stereo: bool) -> str:
    return os.path.join(codes_path, codec_model, f"{chunk_secs}s_{context_secs}s", "stereo" if stereo else "mono")


def _atomic_write(path: str, write_fn) -> None:
    tmp = f"{path}.tmp.{os.getpid()}"
    with open(tmp, "wb") as f:
        write_fn(f)
        f.flush()
        os.fsync(f.fileno())
    os.replace(tmp, path)


def _load_existing(dst: str) -> Optional[np.ndarray]:
    """A previous run's output, or None if it is missing / truncated / not a (1, T) integer array."""
    try:
        arr = np.load(dst, mmap_mode="r")
        if arr.ndim == 2 and arr.shape[0] == 1 and np.issubdtype(arr.dtype, np.integer):
            return np.asarray(arr)
    except Exception:                                               # noqa: BLE001
        pass
    return None


class _PinnedPool:
    """A few pinned host buffers recycled between the loader threads and the copy stream."""

    def __init__(self, count: int, make_buffer):
        self.free: "queue.Queue[Optional[torch.Tensor]]" = queue.Queue()
        self.make_buffer = make_buffer
        for _ in range(count):
            self.free.put(None)

    def acquire(self, nbytes: int, prev: Optional[torch.Tensor] = None) -> torch.Tensor:
        buf = prev if prev is not None else self.free.get()
        if buf is None or buf.numel() < nbytes:
            buf = self.make_buffer(max(nbytes + nbytes // 8, 1 << 20))
        return buf

    def release(self, buf: Optional[torch.Tensor]) -> None:
        self.free.put(buf)

    def drain(self) -> List[torch.Tensor]:
        out = []
        while True:
            try:
                buf = self.free.get_nowait()
            except queue.Empty:
                return out
            if buf is not None:
                out.append(buf)


def encode_corpus(gen, audio_path: str, codes_path: str, codec_model: str = "MagiCodec-50Hz-Base",
                  chunk_size_secs: float = 0.1, context_secs: float = 2.0, batch_size: int = 256, stereo: bool = False,
                  audio_filter: Optional[Sequence[str]] = None, rank: int = 0, world_size: int = 1,
                  overwrite: bool = False, fuse_batches: int = 4, loader_threads: int = 4, prefetch_files: int = 3,
                  stats: Optional[Dict[str, float]] = None, ingest=None) -> Tuple[List[corpus.ManifestEntry], List[Dict[str, str]]]:
    """Encodes this rank's share of the corpus.  Returns (manifest entries incl. resumed files, per-file errors).
    `ingest`: the staging / conversion interface (default audio_io.DeviceIngest(gen): CUDA streams, pinned memory and
    the ingest kernels; the CPU tests drive the same pipeline with a host-side stand-in around the oracle)."""
    out_root = output_dir(codes_path, codec_model, chunk_size_secs, context_secs, stereo)
    files = get_files(audio_path, filters=audio_filter)
    chans = 2 if stereo else 1
    sr = gen.sample_rate
    chunk_samples = int(chunk_size_secs * sr)

    # ---- balance ranks on decoded duration (seconds of audio the GPU will encode), not on bytes
    with ThreadPoolExecutor(max(1, loader_threads)) as ex:
        probes = list(ex.map(lambda p: audio_io.probe_audio(p, default_sr=sr), files))
    cost = [chans * frames / float(psr) for psr, _, frames in probes]
    mine = corpus.shard_by_duration(cost, world_size)[rank]
    os.makedirs(out_root, exist_ok=True)
    if rank == 0:
        info = {"codec_model": codec_model, "framerate": sr / gen.hop, "num_codebooks": 1,
                "codebook_size": gen.codebook_size, "sampling_rate": sr,
                "chunk_size_secs": chunk_size_secs, "context_secs": context_secs, "stereo": stereo}
        _atomic_write(os.path.join(out_root, "codec_info.json"), lambda f: f.write(json.dumps(info, indent=2).encode()))

    manifest: List[corpus.ManifestEntry] = []
    errors: List[Dict[str, str]] = []
    todo: List[Tuple[int, str, List[str]]] = []
    for fid in mine:
        rel = os.path.splitext(os.path.relpath(files[fid], audio_path))[0]
        dsts = [os.path.join(out_root, f"{rel}_c{c}.npy") for c in range(chans)]
        done = None if overwrite else [_load_existing(d) for d in dsts]
        if done is not None and all(a is not None for a in done):
            for c, arr in enumerate(done):                            # resume: the entry comes back from disk
                n = int(arr.shape[-1])
                manifest.append(corpus.ManifestEntry(fid, c, n, -(-n * gen.hop // chunk_samples),
                                                     zlib.crc32(arr.astype("<i4").tobytes()) & 0xFFFFFFFF, rank, rel))
            continue
        todo.append((fid, rel, dsts))

    # ---- pipeline
    if ingest is None:                                               # one per engine: its copy stream, filter taps and pinned staging
        ingest = getattr(gen, "_corpus_ingest", None)               # buffers outlive a run (a second run in the process reuses them)
        if ingest is None:
            ingest = audio_io.DeviceIngest(gen)
            try:
                gen._corpus_ingest = ingest
            except AttributeError:
                pass
    pool = _PinnedPool(prefetch_files + 1, ingest.host_buffer)
    write_q: "queue.Queue" = queue.Queue(maxsize=4 * (prefetch_files + 1))    # a slow disk throttles the encoder, not memory
    lock = threading.Lock()

    def load(item):
        fid, rel, dsts = item
        holder = {"buf": None}

        def alloc(nbytes: int) -> np.ndarray:
            holder["buf"] = pool.acquire(nbytes, holder["buf"])
            return holder["buf"].numpy()

        try:
            pcm = audio_io.read_audio(files[fid], alloc=alloc, default_sr=sr)
            if holder["buf"] is None:                                # zero-length payload never called alloc
                holder["buf"] = pool.acquire(1)
            return item, pcm, holder["buf"], None
        except Exception as ex_:                                      # noqa: BLE001  (corrupt / unreadable / unsupported:
            if holder["buf"] is None:                                # the file goes to errors.json, the run goes on)
                holder["buf"] = pool.acquire(1)
            msg = str(ex_) if isinstance(ex_, UnsupportedAudio) else f"{files[fid]}: {ex_!r}"
            return item, None, holder["buf"], msg

    def writer():
        while True:
            job = write_q.get()
            if job is None:
                return
            fid, rel, dsts, host_codes, wait, n_windows = job
            tmps = []
            try:
                wait()
                entries = []
                for c, dst in enumerate(dsts):
                    arr = host_codes[c].numpy().astype(np.int32)[None, :]   # (num_codebooks=1, T) int32
                    os.makedirs(os.path.dirname(dst), exist_ok=True)
                    tmp = f"{dst}.tmp.{os.getpid()}"
                    with open(tmp, "wb") as f:
                        np.save(f, arr)
                    tmps.append((tmp, dst))
                    entries.append(corpus.ManifestEntry(fid, c, int(arr.shape[-1]), n_windows,
                                                        zlib.crc32(arr.astype("<i4").tobytes()) & 0xFFFFFFFF, rank, rel))
                for tmp, dst in tmps:                                 # a file counts as done only with ALL its channels in place
                    os.replace(tmp, dst)
                with lock:
                    manifest.extend(entries)
                release = getattr(wait, "release", None)              # pinned code buffers go back to the ingest's pool
                if release is not None:
                    host_codes = None
                    release()
            except Exception as ex_:                                  # noqa: BLE001
                for tmp, _ in tmps:
                    try:
                        os.remove(tmp)
                    except OSError:
                        pass
                with lock:
                    errors.append({"file": files[fid], "error": f"write failed: {ex_!r}"})

    wt = threading.Thread(target=writer, daemon=True)
    wt.start()
    encoded_secs = 0.0
    ex = ThreadPoolExecutor(max(1, loader_threads))
    try:
        inflight = []
        it = iter(todo)
        for _ in range(prefetch_files + 1):                           # never more tasks in flight than pinned buffers
            nxt = next(it, None)
            if nxt is not None:
                inflight.append(ex.submit(load, nxt))
        while inflight:
            item, pcm, buf, err = inflight.pop(0).result()
            fid, rel, dsts = item
            if err is None and pcm.frames == 0:
                err = f"{files[fid]}: empty audio"
            if err is not None:
                with lock:
                    errors.append({"file": files[fid], "error": err})
                pool.release(buf)
            else:
                staged = ingest.upload(pcm, buf)                       # copy stream: under the previous file's encode
                wav = ingest.convert(pcm, staged, mono=not stereo)     # fp32 [C, T] at the codec's rate, main stream
                if stereo and wav.shape[0] == 1:
                    wav = torch.cat([wav, wav], dim=0)
                streams = [wav[c] for c in range(chans)]
                codes = corpus.encode_streams(gen, streams, chunk_size_secs, context_secs, batch_size, fuse_batches)
                host_codes, wait = ingest.codes_to_host(codes)
                n_windows = -(-int(wav.shape[1]) // chunk_samples)
                encoded_secs += chans * wav.shape[1] / float(sr)
                write_q.put((fid, rel, dsts, host_codes, wait, n_windows))
                ingest.wait_uploaded(staged)                          # the payload has left the pinned buffer (the GPU is
                pool.release(buf)                                     # still busy with the previous file's encode)
            nxt = next(it, None)
            if nxt is not None:
                inflight.append(ex.submit(load, nxt))
    finally:
        ex.shutdown(wait=True, cancel_futures=True)
        write_q.put(None)
        wt.join()
        if hasattr(ingest, "recycle"):
            for buf in pool.drain():
                ingest.recycle(buf)
    if stats is not None:
        stats["encoded_audio_secs"] = stats.get("encoded_audio_secs", 0.0) + encoded_secs
        stats["files_encoded"] = stats.get("files_encoded", 0) + len(todo) - len(errors)
    manifest.sort(key=lambda e: (e.file_id, e.channel))
    return manifest, errors


def main(argv: Optional[Sequence[str]] = None) -> None:
    ap = argparse.ArgumentParser(prog="python -m rca_b200_loader audio_to_codes", description=__doc__.split("\n\n")[0])
    ap.add_argument("--audio_path", required=True)
    ap.add_argument("--codes_path", required=True)
    ap.add_argument("--chunk_size_secs", type=float, default=0.1)
    ap.add_argument("--context_secs", type=float, default=2.0)
    ap.add_argument("--batch_size", type=int, default=256)
    ap.add_argument("--codec_model", default="MagiCodec-50Hz-Base")
    ap.add_argument("--stereo", action="store_true")
    ap.add_argument("--audio_filter", nargs="+")
    ap.add_argument("--overwrite", action="store_true")
    ap.add_argument("--loader_threads", type=int, default=4)
    ap.add_argument("--prefetch_files", type=int, default=3)
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
    local_manifest, local_errors = encode_corpus(gen, args.audio_path, args.codes_path, args.codec_model, args.chunk_size_secs,
                                                 args.context_secs, args.batch_size, args.stereo, args.audio_filter, rank, world,
                                                 overwrite=args.overwrite, loader_threads=args.loader_threads,
                                                 prefetch_files=args.prefetch_files)
    merged = corpus.gather_manifests(local_manifest, dev)
    all_errors = corpus.gather_objects(local_errors, dev)
    if rank == 0:
        out_root = output_dir(args.codes_path, args.codec_model, args.chunk_size_secs, args.context_secs, args.stereo)
        _atomic_write(os.path.join(out_root, "manifest.json"), lambda f: f.write(json.dumps([e.__dict__ for e in merged]).encode()))
        _atomic_write(os.path.join(out_root, "errors.json"), lambda f: f.write(json.dumps(all_errors, indent=1).encode()))
        print(f"encoded {len(merged)} streams -> {out_root}" + (f"; {len(all_errors)} file(s) skipped, see errors.json" if all_errors else ""))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
