"""world_size-2 gloo worker (CPU): host-side logic of the sharded map — block ownership and the mesh-order merge the
rank-0 gather uses. Launched by tests/test_multi_host.py through torch.distributed.run."""
import importlib
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def mesh_order(keys, bpc=8):
    keys = np.asarray(keys, np.int32).reshape(-1, 3)
    chunk = np.floor(keys.astype(np.float32) / np.float32(bpc)).astype(np.int64)        # block2chunk, tsdf.cu:256-260
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0], chunk[:, 2], chunk[:, 1], chunk[:, 0]))
    return keys[order]


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    vh = importlib.import_module("voxel-hashing-sdf_b200")
    rng = np.random.RandomState(7)                                   # same list on every rank
    keys = np.unique(rng.randint(-300, 300, size=(20000, 3)).astype(np.int32), axis=0)
    owner = np.array([vh.owner_of_block(*k, world, 1) for k in keys])
    assert owner.min() >= 0 and owner.max() < world
    mine = mesh_order(keys[owner == rank])
    # every rank owns a fair share
    assert abs(len(mine) - len(keys) / world) < 0.05 * len(keys)
    # default granularity: one owner per 8^3-block cube (the reference's chunk), so most +x/+y/+z neighbours share it
    own8 = np.array([vh.owner_of_block(*k, world) for k in keys])
    cube = np.floor(keys / 8.0).astype(np.int64)
    first = {}
    for c, o in zip(map(tuple, cube), own8):
        assert first.setdefault(c, o) == o, "blocks of one cube must share an owner"
    assert abs((own8 == rank).mean() - 1.0 / world) < 0.1
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    total = sum(len(g) for g in gathered)
    assert total == len(keys), "shards must partition the key set"
    assert len(np.unique(np.concatenate(gathered), axis=0)) == len(keys), "shards must be disjoint"
    if rank == 0:
        part, idx = vh.mesh_order_merge(gathered)
        merged = np.stack([gathered[p][i] for p, i in zip(part, idx)])
        assert np.array_equal(merged, mesh_order(keys)), "merge of the shards' lists must equal the global mesh order"
        # out-of-range coordinates have no owner
        assert vh.owner_of_block(1 << 21, 0, 0, world, 1) == -1
        print("MULTI_HOST_OK", len(keys), [len(g) for g in gathered], flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
