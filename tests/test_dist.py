"""CPU-only: the N>1 plumbing of bench.py (frame sharding n mod k, max-over-ranks merge, reference arm on
rank 0 only) with world_size-2 gloo process groups.  No collective sits on the data path; these are the only
cross-rank operations the benchmark performs."""
import json
import os
import subprocess
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bench.frames_of_rank(rank, world, 5)
    # every rank reports its frame numbers; together they must tile 0..9 with n mod k routing
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    slowest = bench.max_over_ranks(10.0 + rank, world)       # rank 1 is "slower"
    dist.barrier()
    q.put((rank, mine, gathered, slowest))
    dist.destroy_process_group()


def test_frame_sharding_and_merge_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, mine, gathered, slowest in res:
        assert all(n % 2 == rank for n in mine)
        assert sorted(sum(gathered, [])) == list(range(10))
        assert slowest == 11.0


def test_reference_arm_rank0_only():
    """Under torchrun the reference arm runs and prints on rank 0 only; other ranks exit 0 silently."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_line_shape():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
