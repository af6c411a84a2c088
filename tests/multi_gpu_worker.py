"""torchrun worker for tests/test_gpu_multi.py: node-sharded embed over WORLD_SIZE GPUs, rank 0 saves the layout."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import annembed_b200 as A
from annembed_b200.dist import broadcast_unique_id, env_rank_world, exchange_layout_handles
from tests.conftest import random_graph


def main(out_path, n, d, fused, flags):
    rank, world, local = env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    row_ptr, col, dst = random_graph(n, 3, 9, seed=123)
    y0 = np.random.default_rng(5).uniform(-1, 1, size=(n, d)).astype(np.float32)
    params = A.EmbedderParams(asked_dim=d, nb_grad_batch=4, grad_step=1.0, seed=77, dmap_init=False, flags=flags)
    ctx = A.CudaContext(params, device=local)
    uid = broadcast_unique_id(ctx.unique_id, rank, world)
    ctx.comm_init(rank, world, uid)
    ctx.set_graph_csr(row_ptr, col, dst)
    if fused:
        exchange_layout_handles(ctx, rank, world)
    ctx.edge_weights(want_outputs=False)
    ctx.set_embedding(y0)
    ce0, ce1 = ctx.optimize()
    y = ctx.get_embedding()
    st = ctx.get_stats()
    tot = torch.tensor([float(st["positive_samples"])], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(tot)
    if rank == 0:
        np.savez(out_path, y=y, ce=np.array([ce0, ce1]), samples=np.array([tot.item()]), exchange_ms=st["exchange_ms"],
                 exchanges=st["exchanges"], launches=st["epoch_launches"], cross_rank_edges=st["cross_rank_edges"])
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]) != 0, int(sys.argv[5]) if len(sys.argv) > 5 else 0)
