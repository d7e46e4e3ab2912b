"""Developer check of the multi-GPU path on a box with >= 2 GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/tools/shard_check.py
every rank ticks its slice through sharding.ShardedMPC (library NCCL gather) and rank 0 checks the gathered batch against
the oracle and against single-GPU solves of the same records."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from hector_simulation_b200 import interface, scenarios, sharding  # noqa: E402


def main():
    world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    N, B = 10, 1000  # not a multiple of the world size: the tail slice is padded

    def bcast(b):
        box = [b]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    recs, _ = scenarios.make_batch(3, B, horizon=N, seed=4321)
    sh = sharding.ShardedMPC(B, N, rank, world, lambda bl: sharding.GpuBackend(bl, N, rank, world, lr, bcast), scenarios.UPDATE_DTYPE)
    print(f"rank {rank}: slice {sh.lo}:{sh.hi} of {B}, b_local {sh.b_local}", flush=True)
    mine = sh.local_slice(recs)
    for tick in range(4):
        w, s = sh.tick(mine)
        torch.cuda.synchronize()
        assert (interface.status_code(s) == 0).all()
    whole = sh.whole_batch()
    torch.cuda.synchronize()
    one = interface.BatchedMPC(B, N, device=lr)
    w1, s1 = one.solve_batch(recs)
    one.close()
    ok = np.array_equal(whole, w1.astype(np.float32))
    print(f"rank {rank}: gathered batch equals the single-GPU solve of all {B} records: {ok}", flush=True)
    assert ok
    sh.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
