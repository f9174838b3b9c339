"""Spawned by tests/test_corpus_host.py: one rank of a world_size-2 gloo group."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def manifest_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    import rca_b200_loader  # noqa: F401
    from realtime_codec_agent_b200 import corpus

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    durs = [10.0, 20.0, 30.0, 40.0, 50.0]
    mine = corpus.shard_by_duration(durs, world)[rank]
    local = [corpus.manifest_entry(f, 0, torch.arange(int(durs[f]) * 50) % 97, int(durs[f] * 10), rank) for f in mine]
    merged = corpus.gather_manifests(local)
    q.put((rank, [(e.file_id, e.n_frames, e.crc32, e.rank) for e in merged]))
    dist.destroy_process_group()
