"""N > 1 host logic on CPU: world_size 2 and 3 over the `gloo` backend.

What the multi-GPU path does per rank (ba_solver.cu, lm_solver.cu): take the contiguous point range
sk_partition_points gives it, keep ALL cameras, compute its observations' contributions, and sum the
camera-block quantities (gradient, column norms, reduced rhs, SchurJacobi blocks, matvec result) and the
point-partitioned scalars (cost, model cost change, norms) across ranks.  Here every rank computes its
share with the CPU oracle / numpy and the sums travel over a real torch.distributed process group, so the
partition, the "cameras replicated / points sharded" bookkeeping and the rendezvous used by bench.py
(unique id broadcast from rank 0) are exercised without a GPU.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from skeres_b200 import _abi, api, synth
    d = synth.make_bal("small", seed=12)
    off = d.block_offsets()
    # the id broadcast bench.py performs before sk_comm_create
    ids = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    assert ids[0] == bytes(range(128))
    # this rank's point range (balanced by observation count) -> its residual blocks
    ptr = np.concatenate([[0], np.cumsum(np.bincount(d.point_index, minlength=d.num_points))]).astype(np.int64)
    begin = api.partition_points(ptr, world)
    mine = (d.point_index >= begin[rank]) & (d.point_index < begin[rank + 1])
    p = O.OracleProblem(d.parameters)
    p.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2)[mine], off[mine])
    cost, r, g, J = p.evaluate()
    nc = 9 * d.num_cameras
    # camera-block sums + point-partitioned scalars in ONE buffer (as the device path packs them)
    Jm = J.reshape(-1, 24)
    colsq = np.zeros(d.parameters.size)
    np.add.at(colsq, (off[mine][:, 0:1] + np.arange(9)).ravel(), (Jm[:, :18].reshape(-1, 2, 9) ** 2).sum(1).ravel())
    np.add.at(colsq, (off[mine][:, 1:2] + np.arange(3)).ravel(), (Jm[:, 18:].reshape(-1, 2, 3) ** 2).sum(1).ravel())
    buf = torch.from_numpy(np.concatenate([g[:nc], colsq[:nc], [cost, float(mine.sum())]]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    # the point part never leaves its rank; gather it only to check the result
    pts = [None] * world
    dist.all_gather_object(pts, (int(begin[rank]), int(begin[rank + 1]), g[nc:], colsq[nc:]))
    if rank == 0:
        full = O.OracleProblem(d.parameters)
        full.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), off)
        fc, fr, fg, fJ = full.evaluate()
        b = buf.numpy()
        assert np.allclose(b[:nc], fg[:nc], rtol=1e-12, atol=1e-9), "camera gradient"
        assert np.isclose(b[-2], fc, rtol=1e-13), "cost"
        assert int(b[-1]) == d.num_observations
        g_pts = np.zeros(3 * d.num_points)
        for lo, hi, gp, _ in pts:
            assert np.all(gp[:3 * lo] == 0) and np.all(gp[3 * hi:] == 0)       # a rank only touches its own points
            g_pts[3 * lo:3 * hi] = gp[3 * lo:3 * hi]
        assert np.allclose(g_pts, fg[nc:], rtol=1e-12, atol=1e-9), "point gradient"
        fJm = fJ.reshape(-1, 24)
        full_colsq = np.zeros(nc)
        np.add.at(full_colsq, (off[:, 0:1] + np.arange(9)).ravel(), (fJm[:, :18].reshape(-1, 2, 9) ** 2).sum(1).ravel())
        assert np.allclose(b[nc:2 * nc], full_colsq, rtol=1e-12)
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_point_partitioned_sums_over_gloo(world, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").exists()
