"""CPU, world_size 2 over gloo: the multi-rank plumbing of bench.py (the op itself never communicates:
images are independent, ranks only meet at the barriers / max-over-ranks reduction)."""
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


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from richsem_b200 import synthetic as syn

    ms = bench.max_over_ranks(10.0 + 5.0 * rank, world, "cpu")  # rank 1 is the slow one
    # each rank builds its own shard: different seeds -> different tensors, same shapes
    shapes = [(4, 6), (2, 3)]
    i = syn.make_inputs("U", 2, shapes, "cpu", seed=bench.rank_seed(1234, rank, 0), lq=5, m=2, d=4, p=2)
    sums = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(sums, i["value"].double().sum().reshape(1))
    dist.barrier()
    out[rank] = (ms, [float(s) for s in sums], tuple(i["value"].shape))
    dist.destroy_process_group()


def test_max_over_ranks_and_distinct_shards():
    world, port = 2, 29631
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][0] == out[1][0] == 15.0          # both ranks agree on the slowest rank's time
    assert out[0][1] == out[1][1] and out[0][1][0] != out[0][1][1]   # shards differ, gathered views agree
    assert out[0][2] == out[1][2]
    import bench

    assert bench.aggregate_qps(1000, 2, 15.0) == 2 * 1000 / 15e-3


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29633", str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "1"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["unit"] == "queries/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0


def test_reference_arm_does_not_load_the_product_library():
    """bench.py --impl reference (the CPU arm) must not map libmsda_b200.so: it loads the synthetic-input module by
    path and the oracle, never the richsem_b200 package (whose import dlopens the library)."""
    code = ("import sys, bench; from oracle.msda_oracle import core_pytorch_fwd_bwd; syn = bench.load_synthetic(); "
            "assert syn.level_shapes(800, 1333)[0] == (100, 167); "
            "assert not [m for m in sys.modules if m.startswith('richsem_b200')], sys.modules.keys(); "
            "maps = open('/proc/self/maps').read(); assert 'libmsda_b200' not in maps; print('clean')")
    r = subprocess.run([sys.executable, "-c", code], cwd=str(ROOT), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "clean" in r.stdout, r.stderr[-2000:]
