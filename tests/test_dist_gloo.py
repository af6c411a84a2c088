"""world_size-2 gloo test (CPU) of the multi-GPU host plumbing: id broadcast + shard ranges."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, out_dir):
    import torch.distributed as dist
    from annembed_b200.dist import broadcast_unique_id, exchange_layout_handles, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = broadcast_unique_id(lambda: (np.arange(128) * 7 % 251).astype(np.uint8), rank, world)
    lo, hi = shard_range(n, rank, world)

    class FakeCtx:                      # stands in for CudaContext: the exchange only moves 128 bytes per rank
        def export_layout(self):
            return np.full(128, 10 + rank, np.uint8)

        def import_layouts(self, allh):
            self.got = np.asarray(allh).copy()

    fc = FakeCtx()
    exchange_layout_handles(fc, rank, world)
    assert fc.got.shape == (world * 128,)
    assert all((fc.got[128 * r:128 * (r + 1)] == 10 + r).all() for r in range(world))
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([uid.astype(np.int64), [lo, hi]]))
    dist.barrier()
    dist.destroy_process_group()


def test_unique_id_broadcast_and_shards(tmp_path):
    world, n = 2, 1001
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (np.load(tmp_path / f"r{r}.npy") for r in range(world))
    np.testing.assert_array_equal(r0[:128], (np.arange(128) * 7 % 251))
    np.testing.assert_array_equal(r0[:128], r1[:128])
    assert r0[128] == 0 and r0[129] == r1[128] == 512 and r1[129] == n


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """bench.py --impl reference launched like the driver launches it for N > 1: rank 0 alone runs the CPU arm and prints
    ONE JSON line (same metric / unit / config object as the GPU arm's), the other rank exits 0 without work."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--nodes", "20000", "--steps", "1", "--warmup", "0", "--cpu-seconds", "0.2"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "edge_updates_per_s" and d["unit"] == "edge updates/s" and d["n_gpus"] == 2
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert cb["reference_layout"]["value"] > 0 and set(d["config"]) == {"workload", "l2", "cpu_arm"}
