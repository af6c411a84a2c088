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
