"""CPU, world_size 2 over gloo: the N>1 plumbing of bench.py -- contiguous stream sharding and
the load-time broadcast of the model file images from rank 0 (the only communication step;
the per-hop path has no collective).  Each rank then runs ITS shard of streams through the
oracle ABI from the broadcast bytes and the union must equal the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from beatrice_vst_b200 import dist as bdist
from conftest import ROOT


def test_shard_streams_partitions_exactly():
    for total, world in [(2048, 8), (256, 1), (1024, 8), (10, 4), (3, 8)]:
        seen = []
        for r in range(world):
            first, count = bdist.shard_streams(total, world, r)
            seen += list(range(first, first + count))
        assert seen == list(range(total))
    assert bdist.shard_streams(2048, 8, 3) == (768, 256)      # stream -> gpu = id // 256


def _worker(rank, world, port, model_dir, out_dir):
    import tempfile

    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from beatrice_vst_b200 import lib as blib
    from beatrice_vst_b200 import signals
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    images = bdist.read_model_images(model_dir) if rank == 0 else None
    images = bdist.broadcast_model_images(images, src=0)
    # materialise the received bytes and run this rank's shard through the oracle ABI
    with tempfile.TemporaryDirectory() as d:
        for name, im in zip(bdist.MODEL_FILES, images):
            im.tofile(os.path.join(d, name))
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import loader as oracle_loader
        oracle = oracle_loader.load_oracle()
        first, count = bdist.shard_streams(4, world, rank)
        for s in range(first, first + count):
            st = blib.SingleStream(oracle, d, speaker=s % 8)
            _, q, _, w = st.run(signals.voice_like(160 * 3, 16000.0, seed=s))
            st.close()
            np.save(os.path.join(out_dir, f"w{s}.npy"), w)
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_and_sharded_streams_world2(model_dir, oracle, tmp_path):
    from beatrice_vst_b200 import lib as blib
    from beatrice_vst_b200 import signals
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, model_dir, str(tmp_path)), nprocs=2, join=True)
    for s in range(4):
        st = blib.SingleStream(oracle, model_dir, speaker=s % 8)
        _, _, _, w = st.run(signals.voice_like(160 * 3, 16000.0, seed=s))
        st.close()
        assert np.array_equal(np.load(tmp_path / f"w{s}.npy"), w)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_reference_arm_under_torchrun_world2():
    """`bench.py --impl reference` as the driver launches it for N = 2: rank 0 alone times the CPU arm (the
    reference call site over the oracle) and prints the one JSON line, the other rank leaves without work."""
    import json
    import subprocess

    import callsite
    if not callsite.available("oracle"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, lines                       # exactly one rank speaks
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["higher_is_better"] is True
    assert d["metric"].startswith("voice frames/s") and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
