#!/usr/bin/env python
"""bench.py — encode throughput of the MagiCodec tokenization path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference ...                      (the reference's CPU path, oracle port)

Workload (BASELINE.json configs[1], the prep_lm_dataset_magicodec shape): per rank, 1 h of synthetic
16 kHz mono audio as 6 x 10 min files, encoded the way encode_audio_gpu_N.sh does it — 0.1 s chunks,
2.0 s context, 256 windows per batch, random-init default-spec weights.  One "step" = one pass over
that hour.  `value` = audio-seconds per second over all ranks with the audio resident in HBM;
`e2e` = the same through the host-facing call with pinned HOST buffers (H2D of the audio and D2H of
the codes inside the timed region).  Weak scaling: every rank encodes its own hour (file-sharded, no
data-path collective); at N>1 each step ends with the NCCL all-gather of the per-rank manifests.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import rca_b200_loader  # noqa: F401
import realtime_codec_agent_b200 as pkg
from realtime_codec_agent_b200 import corpus

METRIC = "encode audio-sec/sec"
UNIT = "audio-s/s"
FILES_PER_RANK = 6
FILE_SECS = 600.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-size", type=int, default=256, help="windows per batch (the reference CLI's --batch_size)")
    ap.add_argument("--fuse-batches", type=int, default=4, help="consecutive batches issued as one engine launch "
                    "(results are independent of launch grouping; tests/test_gpu_parity.py checks this bit-exactly)")
    ap.add_argument("--file-secs", type=float, default=FILE_SECS)
    ap.add_argument("--files", type=int, default=FILES_PER_RANK)
    ap.add_argument("--cpu-sample-secs", type=float, default=0.0, help="0 = pick ~15 s of CPU work")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------- CPU reference arm
def cpu_reference_run(sample_secs: float, steps: int, warmup: int, cores: int):
    """The reference's CPU path: AudioTokenizer.chunked_tokenize_audio semantics (0.1 s chunks, 2.0 s
    context) around the fp32 oracle port of the network, all host threads.  The unmodified reference
    wrapper cannot travel to the GPU box (/root/reference is absent there), so the host side is this
    repo's AudioTokenizer, which the CPU tests prove string-identical to it."""
    from oracle.magicodec_oracle import OracleGenerator

    torch.set_num_threads(cores)
    spec = pkg.DEFAULT_SPEC
    model = OracleGenerator(spec, pkg.init_random_weights(spec, seed=0))
    tok = pkg.AudioTokenizer(codec_model=model, device="cpu")
    wav = pkg.synth_audio(int((2.0 + 60.0) * 16000), seed=1234, file_id=0).numpy()
    tok.tokenize_audio(wav[:32000])                          # fill the 2.0 s context
    pos = 32000
    t0 = time.perf_counter()
    tok.tokenize_audio(wav[pos:pos + 1600]); pos += 1600
    per_chunk = time.perf_counter() - t0
    if sample_secs <= 0:
        sample_secs = max(0.2, min(6.0, round(15.0 / max(per_chunk, 1e-3)) * 0.1 / max(1, steps + warmup)))
    n_chunks = max(1, int(round(sample_secs / 0.1)))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for _ in range(n_chunks):
            if pos + 1600 > wav.shape[0]:
                pos = 32000
            tok.tokenize_audio(wav[pos:pos + 1600]); pos += 1600
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": n_chunks * 0.1 * len(times) / total, "ms_per_step": 1e3 * total / len(times),
            "sample": f"{n_chunks} chunks of 0.1 s with a full 2.0 s context per step ({n_chunks * 0.1:.1f} s of audio), "
                      f"fp32 oracle port, default spec", "cores": cores}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py from the .ncu-rep); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        return t["gemm"]["dram_bytes_per_launch"], {"algorithmic_bytes_per_launch": t["gemm"]["algorithmic_bytes_per_launch"],
                                                    "launches_captured": t["gemm"]["launches"], "source": t["source"]}
    except Exception:
        return None, None


# --------------------------------------------------------------------------------- GPU arm
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    config = {"workload": f"1 h synthetic 16 kHz mono per GPU ({args.files} x {args.file_secs:.0f} s files), chunk 0.1 s, "
                          f"context 2.0 s, batch {args.batch_size} windows ({args.fuse_batches} batches fused per engine "
                          f"launch), MagiCodec default spec (8+8 layers, d=1024), random-init weights seed 0",
              "windows_per_step_per_gpu": int(args.files * args.file_secs * 10), "l2": "inputs_exceed_l2",
              "sharding": "files across ranks, no data-path collective; manifest all_gather per step at N>1"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(args.cpu_sample_secs, args.steps, max(args.warmup, 1), cores)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import torch.distributed as dist
    # Anything a library prints while the job runs (NCCL's version banner goes to stdout) must not land in front of
    # the ONE JSON line: fd 1 points at stderr until the line is printed.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec = pkg.DEFAULT_SPEC
    gen = pkg.B200Generator(spec, pkg.init_random_weights(spec, seed=0), device=dev)

    n_samp = int(args.file_secs * spec.sample_rate)
    file_ids = [rank * args.files + i for i in range(args.files)]
    dev_files = [pkg.synth_audio(n_samp, seed=1234, file_id=f, device=dev) for f in file_ids]
    host_files = [torch.empty(n_samp, dtype=torch.float32).pin_memory() for _ in file_ids]
    for h, d in zip(host_files, dev_files):
        h.copy_(d)
    staging = [torch.empty_like(d) for d in dev_files]
    audio_secs_per_step = args.files * args.file_secs

    def manifests(codes_list):
        local = [corpus.manifest_entry(f, 0, c, int(c.numel() // 5), rank) for f, c in zip(file_ids, codes_list)]
        return corpus.gather_manifests(local, dev)

    def step_device():
        codes = corpus.encode_streams(gen, dev_files, 0.1, 2.0, args.batch_size, args.fuse_batches)
        if world > 1:
            manifests(codes)
        return codes

    host_codes = [torch.empty(int(n_samp // 320), dtype=torch.int64).pin_memory() for _ in file_ids]

    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event() for _ in file_ids]

    def step_e2e():
        # what a corpus job does: the upload of file i+1 (pinned host -> HBM, side stream) runs under the encode of file i
        main = torch.cuda.current_stream()
        copy_stream.wait_stream(main)                     # staging buffers are free again
        with torch.cuda.stream(copy_stream):
            for s, h, ev in zip(staging, host_files, copied):
                s.copy_(h, non_blocking=True)
                ev.record(copy_stream)
        codes = []
        for s, hc, ev in zip(staging, host_codes, copied):
            main.wait_event(ev)
            c = corpus.encode_streams(gen, [s], 0.1, 2.0, args.batch_size, args.fuse_batches)[0]
            hc.copy_(c, non_blocking=True)
            codes.append(c)
        main.synchronize()
        if world > 1:
            manifests(codes)
        return codes

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = gen.launch_count
        if profile:
            gen.profile_begin()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        prof = gen.profile_end() if profile else None
        ms = max(e0.elapsed_time(e1), 0.0)
        t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), gen.launch_count - l0, prof

    for _ in range(max(args.warmup, 3)):
        codes = step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, _, launches, _ = timed(step_device, args.steps)
    # the same K steps once more with an event pair around every kernel launch: per-class device time for the
    # roofline object.  Kept out of the pass above so that `value` carries no instrumentation overhead.
    prof_ms, _, _, prof = timed(step_device, args.steps, profile=True)
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    _, e2e_wall_ms, _, _ = timed(step_e2e, args.steps)
    sync_all()

    value = world * audio_secs_per_step * args.steps / (dev_ms / 1e3)
    e2e_value = world * audio_secs_per_step * args.steps / (e2e_wall_ms / 1e3)

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else \
            "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        g = prof["gemm"]
        achieved = g["flops"] / (g["ms"] / 1e3) / 1e12 if g["ms"] > 0 else 0.0
        roofline = {"bound": "tensor", "kernel": "gemm_bf16_sm100_kernel (linears + implicit-GEMM convs)",
                    "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    "peak_source": peak_src, "traffic": ncu_traffic()[0], "traffic_detail": ncu_traffic()[1],
                    "launches": g["launches"], "avg_launch_ms": g["ms"] / max(1, g["launches"]),
                    "share_of_step": g["ms"] / prof_ms, "instrumented_ms_per_step": prof_ms / args.steps,
                    "other_classes_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
                    "hbm_kernels_achieved_GBps": (prof["elementwise"]["bytes"] / (prof["elementwise"]["ms"] / 1e3) / 1e9)
                    if prof["elementwise"]["ms"] > 0 else None,
                    "hbm_peak_GBps": peaks.get("hbm_gbs")}
        windows = config["windows_per_step_per_gpu"]
        exec_flops = sum(v["flops"] for v in prof.values()) / args.steps
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_wall_ms / args.steps,
                        "h2d_bytes_per_step": int(args.files * n_samp * 4), "d2h_bytes_per_step": int(sum(h.numel() for h in host_codes) * 8),
                        "api": "corpus.encode_streams per file -> B200Generator.encode -> mc_encode (pinned host audio in on a side stream, "
                               "pinned host codes out)"},
                "gpu_launches": int(launches),
                "roofline": roofline,
                "executed_tflop_per_step_per_gpu": exec_flops / 1e12,
                "reference_equivalent_tflop_per_step_per_gpu": windows * spec.encode_flops(100) / 1e12,
                "codes_checksum": int(sum(int(c.sum().item()) for c in codes) % (1 << 31))}
        if world == 1:
            try:                                      # the decode half of the path at the offline batch shape
                ctx = torch.stack([c[:25600] for c in codes]).reshape(-1, 100)[:1024].contiguous()    # 1024 windows x 100 codes
                gen.decode(ctx)
                torch.cuda.synchronize()
                d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                gen.profile_begin()
                d0.record()
                for _ in range(5):
                    gen.decode(ctx)
                d1.record()
                torch.cuda.synchronize()
                dprof = gen.profile_end()
                dms = d0.elapsed_time(d1) / 5
                line["decode"] = {"value": ctx.shape[0] * 2.0 / (dms / 1e3), "unit": "audio-s/s", "ms_per_launch": dms,
                                  "config": "mc_decode of 1024 windows x 100 codes -> 1024 x 2.0 s of waveform (every sample kept), "
                                            "codes resident in HBM",
                                  "gemm_TFLOPs": dprof["gemm"]["flops"] / (dprof["gemm"]["ms"] / 1e3) / 1e12 if dprof["gemm"]["ms"] else None,
                                  "class_ms_per_launch": {k: v["ms"] / 5 for k, v in dprof.items()}}
            except Exception as ex:
                line["decode"] = {"error": repr(ex)}
            try:                                      # BASELINE metric, second half: p50 streaming decode ms/frame
                from tools.bench_stream import run as stream_run, run_emit
                tok = pkg.AudioTokenizer(codec_model=gen, device=dev)
                st = stream_run(tok, 0.02, 600, 120)
                st100 = stream_run(tok, 0.1, 200, 30)
                emit = run_emit(tok, 200, 30)
                line["streaming"] = {"config": "batch-1, 20 ms frames, 2.0 s context, tokenize_audio + detokenize_audio per frame "
                                               "(device-resident context, CUDA-graph replay), wall clock around the Python call",
                                     "p50_decode_ms_per_frame": st["decode_wall_ms"]["p50"], "decode_wall_ms": st["decode_wall_ms"],
                                     "encode_wall_ms": st["encode_wall_ms"], "decode_cuda_ms": st["decode_cuda_ms"],
                                     "encode_cuda_ms": st["encode_cuda_ms"],
                                     "chunk_100ms": {"decode_wall_ms": st100["decode_wall_ms"], "encode_wall_ms": st100["encode_wall_ms"],
                                                     "note": "the agent's default chunk (realtime_agent_config.py): 5 frames per call"},
                                     "emit_chain_100ms": {"emit_wall_ms": emit["emit_wall_ms"],
                                                          "note": "OutputChunkEmitter.emit = decoder + pad_or_trim + normalize_audio_rms + "
                                                                  "smooth_join in one engine call (realtime_agent_v2.py:556-579)"}}
            except Exception as ex:                   # never lose the main line to the auxiliary metric
                line["streaming"] = {"error": repr(ex)}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(args.cpu_sample_secs, 1, 1, cores)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
