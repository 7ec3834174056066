"""The two BASELINE.json configurations that are not the headline bench line.

  latency  (config 4) batch-1 latency mode: per-frame wall time host-call -> samples-back over 10 000 frames, p50 / p99,
           (a) through the single-stream beatrice.h ABI exactly as ProcessorCore2::Process1 drives it
           (ExtractPhone1 + EstimatePitch1 + GenerateWaveform1, three synchronous calls per 10 ms frame), at the
           library's default arithmetic (split-bf16 on tcgen05), and
           (b) through the batched engine with one stream (BeatriceB200_Process48k, pinned host buffers, depth 1).
  sweep    (config 5) 1024 streams over 8 GPUs = 128 per GPU, per-stream speaker / pitch-shift / formant sweep:
           global stream g has speaker g % n, pitch shift (g % 25) - 12 st, formant ((g % 9) - 4) / 2, kNN-VQ on for
           every fourth stream, and every 100 frames each stream advances its speaker (set-speaker + the 4-hop
           key-value schedule inside the timed region).  One process per GPU under torchrun, streams sharded
           contiguously, no per-hop collective; timed with CUDA events, max over ranks.  Rank 0 checks PARITY inside
           the sweep: sampled streams' last 20 hops against the reference call site (oracle/_ref) over the CPU oracle
           fed the same input history and the same parameter events.

   python tools/config_bench.py latency [frames=10000] [--out profiles/x.json]
   python tools/config_bench.py sweep [frames=1000] [streams_per_gpu=128] [--depth 1|2] [--out profiles/x.json]
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
          tools/config_bench.py sweep 1000 128 --out profiles/x.json
Prints one JSON line per measurement (rank 0) and, with --out, writes them as a JSON list."""
import json
import os
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from beatrice_vst_b200 import batch as bbatch  # noqa: E402
from beatrice_vst_b200 import dist as bdist  # noqa: E402
from beatrice_vst_b200 import lib as blib  # noqa: E402
from beatrice_vst_b200 import model_spec, signals  # noqa: E402

RESULTS = []


def emit(d):
    RESULTS.append(d)
    print(json.dumps(d), flush=True)


def pct(a, p):
    return float(np.percentile(np.asarray(a, np.float64), p))


def stats(t, frames, what):
    t = np.asarray(t, np.float64)
    return {"config": what, "frames": frames, "p50_us": pct(t, 50), "p99_us": pct(t, 99), "p999_us": pct(t, 99.9),
            "max_us": float(t.max()), "mean_us": float(t.mean()), "jitter_us_p99_minus_p50": pct(t, 99) - pct(t, 50),
            "realtime_budget_us": 10000, "budget_used_p99": pct(t, 99) / 10000.0}


def latency(frames):
    product = blib.load_product()
    warm = 50
    with tempfile.TemporaryDirectory() as d:
        model_spec.write_model_dir(d, 8, 2, 0)
        # (a) the reference call site's three calls per frame, default precision (bf16x3 on tcgen05)
        s = blib.SingleStream(product, d)
        assert s.ok, s.errors
        s.set_pitch_range(1, 383)
        x = signals.voice_like((frames + warm) * 160, 16000.0, seed=1)
        t = []
        for i in range(frames + warm):
            f0 = time.perf_counter_ns()
            s.frame(x[i * 160:(i + 1) * 160])
            t.append((time.perf_counter_ns() - f0) * 1e-3)
        s.close()
        emit(stats(t[warm:], frames, "config 4: batch-1 latency, beatrice.h ABI (ExtractPhone1 + EstimatePitch1 + GenerateWaveform1 per 10 ms "
                                     "frame, host buffers, default precision bf16x3 / tcgen05)"))
        # (b) one stream through the batched 48 kHz entry
        eng = bbatch.Engine(product, 1, precision=2)
        assert eng.load(d) == 0
        x48 = signals.batch_48k(1, frames + warm, seed0=2)     # [hops][1][480]
        hin, hout = eng.pinned("in", (1, 480)), eng.pinned("out", (1, 480))
        t = []
        for i in range(frames + warm):
            hin[:] = x48[i]
            f0 = time.perf_counter_ns()
            eng.process_48k(hin, hout)
            t.append((time.perf_counter_ns() - f0) * 1e-3)
        eng.close()
        emit(stats(t[warm:], frames, "config 4: batch-1 latency, BeatriceB200_Process48k (1 call per hop incl. gain + 48 kHz FIRs on device, "
                                     "pinned host buffers, bf16x3, pipeline depth 1)"))


def stream_plan(g, frames, n_speakers, change_every=100):
    ev = [(0, "voice", g % n_speakers), (0, "pitch_shift", float((g % 25) - 12)), (0, "formant_shift", ((g % 9) - 4) / 2.0)]
    if g % 4 == 3 and os.environ.get("CFG5_NO_VQ") != "1":           # (diagnosis switches: what the sweep costs by part)
        ev.append((0, "vq_num_neighbors", 4))
    for h in range(change_every, frames if os.environ.get("CFG5_NO_CHANGE") != "1" else 0, change_every):
        ev.append((h, "voice", (g + h // change_every) % n_speakers))
    return ev


SETTER = dict(voice="TargetSpeaker", pitch_shift="PitchShift", formant_shift="FormantShift", vq_num_neighbors="VQNumNeighbors")


def sweep(frames, n, depth):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    product = blib.load_product()
    tmp = tempfile.TemporaryDirectory()
    images = None
    if rank == 0:
        model_spec.write_model_dir(tmp.name, 8, 2, 0)
        images = bdist.read_model_images(tmp.name)
    if world > 1:
        images = bdist.broadcast_model_images(images, src=0, device=torch.device("cuda", local_rank))
    eng = bbatch.Engine(product, n, device=local_rank, precision=2)
    assert eng.load_from_memory(images) == 0
    assert eng.set_pipeline_depth(depth) == 0
    ns = eng.n_speakers
    first, _ = bdist.shard_streams(n * world, world, rank)
    warm = 20
    total = warm + frames
    plans = [stream_plan(first + s, total, ns) for s in range(n)]
    by_hop = {}                                            # hop -> [(stream, setter, value)]: no per-hop scan of the plans
    for s in range(n):
        for (b, name, v) in plans[s]:
            by_hop.setdefault(b, []).append((s, name, int(v) if name in ("voice", "vq_num_neighbors") else float(v)))
    x48 = signals.batch_48k(min(n, 32), 64, seed0=5 + first)
    x48 = np.tile(x48, (1, (n + 31) // 32, 1))[:, :n, :]
    sampled = sorted(set([0, 5, 6, n // 2, n - 1])) if rank == 0 else []
    keep = 20
    # the 64 distinct input hops live in pinned host memory and the C entry point is called with their addresses, as a native
    # host would: no numpy copy / wrapper work per step inside the timed region (the setter calls stay inside it)
    hbank, hout = eng.pinned("in", (64, n, 480)), eng.pinned("out", (n, 480))
    hbank[:] = x48
    in_ptrs = [hbank[i].ctypes.data for i in range(64)]
    out_ptr = hout.ctypes.data
    call48, handle = eng.dll.BeatriceB200_Process48k, eng.h
    stream = torch.cuda.ExternalStream(eng.cuda_stream, device=torch.device("cuda", local_rank))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    changes, tail = 0, []

    def step(i):
        nonlocal changes
        for (s, name, val) in by_hop.get(i, ()):
            assert eng.set(SETTER[name], val, s) == 0
            changes += int(name == "voice" and i >= warm)
        if call48(handle, in_ptrs[i % 64], out_ptr) != 0:
            raise RuntimeError("BeatriceB200_Process48k failed")
        if sampled and i >= total - keep:
            tail.append(hout[sampled].copy())

    for i in range(warm):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev0.record(stream)
    for i in range(warm, total):
        step(i)
    ev1.record(stream)
    eng.synchronize()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms, wall * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, wall = float(t[0].item()), float(t[1].item()) * 1e-3
    if depth == 2 and sampled:
        tail = tail[1:] + [eng.drain()[sampled]]
    finite = bool(np.isfinite(hout).all())
    eng.close()
    parity = None
    if rank == 0:
        import callsite
        if callsite.available("oracle"):
            toml = os.path.join(tmp.name, "model.toml")
            got = np.stack(tail, axis=1)                     # [sampled][keep][480]

            def one(k):
                s = sampled[k]
                hist = np.concatenate([x48[i % 64][s] for i in range(total)])
                y, info = callsite.run("oracle", toml, hist, events=plans[s])
                assert info["load"] == 0 and info["last"] == 0
                ref = y[-keep * 480:]
                return float(np.sqrt(np.mean((ref.astype(np.float64) - got[k].reshape(-1)) ** 2))), float(ref.std())

            with ThreadPoolExecutor(max_workers=min(len(sampled), os.cpu_count() or 1)) as pool:
                res = list(pool.map(one, range(len(sampled))))
            parity = {"rms_worst": max(r[0] for r in res), "signal_rms_min": min(r[1] for r in res), "streams": len(res),
                      "hops_compared": keep, "hops_of_history": total,
                      "against": "reference call site (oracle/_ref) + CPU oracle, same inputs and parameter events"}
        emit({"config": f"config 5: {n * world} streams over {world} GPU(s) ({n} per GPU), per-stream speaker + pitch-shift (-12..+12 st) + "
                        "formant sweep, kNN-VQ on every 4th stream, every stream changes speaker every 100 frames (set-speaker + 4-hop "
                        "key-value schedule inside the timed region), BeatriceB200_Process48k with pinned host buffers, bf16x3",
              "n_gpus": world, "streams_per_gpu": n, "frames": frames, "pipeline_depth": depth,
              "frames_per_s": n * world * frames / (ms * 1e-3), "ms_per_hop": ms / frames,
              "frames_per_s_wall": n * world * frames / wall, "speaker_changes_rank0": changes, "output_finite": finite,
              "timing": "CUDA events on the engine stream around the whole loop (host-side setter calls included), max over ranks",
              "parity": parity})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    argv = sys.argv[1:]
    out = None
    depth = 1
    if "--out" in argv:
        k = argv.index("--out")
        out = argv[k + 1]
        del argv[k:k + 2]
    if "--depth" in argv:
        k = argv.index("--depth")
        depth = int(argv[k + 1])
        del argv[k:k + 2]
    mode = argv[0] if argv else "latency"
    if mode == "latency":
        latency(int(argv[1]) if len(argv) > 1 else 10000)
    else:
        sweep(int(argv[1]) if len(argv) > 1 else 1000, int(argv[2]) if len(argv) > 2 else 128, depth)
    if out and RESULTS:
        with open(out if os.path.isabs(out) else os.path.join(ROOT, out), "w") as f:
            json.dump(RESULTS, f, indent=1)
