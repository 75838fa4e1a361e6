"""N > 1 path on CPU: two gloo processes shard a batch, run a stand-in calculation on their block, gather status arrays.
(The GPU calculation itself is covered by the -m gpu tests; here the partition / slicing / gather plumbing is checked.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_scn, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from pgm_b200 import structs
    from pgm_b200.distributed import calculate_power_flow_sharded, scenario_block

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dense = np.zeros((n_scn, 3), structs.UPDATE["sym_load"])
    dense["id"] = [7, 8, 9]
    dense["p_specified"] = np.arange(n_scn)[:, None] * 10.0 + np.arange(3)[None, :]
    counts = np.arange(n_scn) % 3
    indptr = np.concatenate([[0], np.cumsum(counts)])
    sparse = {"data": np.zeros(indptr[-1], structs.UPDATE["asym_load"]), "indptr": indptr}
    sparse["data"]["id"] = np.arange(indptr[-1])

    def fake_calculate(update, **kwargs):
        # stands in for PowerGridModel.calculate_power_flow: status = 1 where the first load's p is a multiple of 40
        p0 = update["sym_load"]["p_specified"][:, 0]
        assert np.array_equal(np.diff(update["asym_load"]["indptr"]), (p0 / 10).astype(int) % 3)
        return {"node": p0.copy()}, (p0 % 40 == 0).astype(np.int32), (p0 / 10).astype(np.int32)

    block, result, status, n_iter = calculate_power_flow_sharded(None, {"sym_load": dense, "asym_load": sparse}, dist=dist,
                                                                 calculate=fake_calculate)
    assert block == scenario_block(n_scn, rank, world)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([[block[0], block[1]], status, n_iter]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_scn", [10, 7, 1])
def test_two_rank_gloo_sharding(tmp_path, n_scn):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_scn, str(tmp_path)), nprocs=world, join=True)
    expected_status = (np.arange(n_scn) * 10 % 40 == 0).astype(int)
    blocks = []
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npy")
        blocks.append((int(d[0]), int(d[1])))
        assert np.array_equal(d[2:2 + n_scn], expected_status)          # every rank sees the whole status array
        assert np.array_equal(d[2 + n_scn:], np.arange(n_scn))
    assert blocks[0][0] == 0 and blocks[-1][1] == n_scn and blocks[0][1] == blocks[1][0]  # contiguous cover


def test_scenario_block_partition():
    sys.path.insert(0, ROOT)
    from pgm_b200.distributed import scenario_block

    for n in (0, 1, 7, 1000, 100000):
        for w in (1, 2, 4, 8):
            blocks = [scenario_block(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in blocks) <= -(-n // w) if n else True
