"""Multi-GPU correctness check (development helper): torchrun --nproc-per-node N tools/multi_gpu_check.py [shape]
Every rank solves the same problem through a point-partitioned communicator; rank 0 also solves it on a
single GPU and compares."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from skeres_b200 import _abi, api, synth

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
api.check(api.lib.sk_set_device(local))
ids = [api.Communicator.unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
comm = api.Communicator(ids[0], rank, world)
shape = sys.argv[1] if len(sys.argv) > 1 else "ladybug-49"
# "wide": more virtual blocks of 8 cameras than persistent CTAs -- the packed (three blocks per round) vector phases of the fused solve
d = synth.make_bal(n_cam=2500, n_pt=40000, n_obs=400000, seed=7) if shape == "wide" else synth.make_bal(shape, seed=1)

def solve(use_comm, local=False, data=None):
    bal = api.BalProblem.fromArrays(d if data is None else data)
    prob = bal.buildLocalProblem(rank, world) if local else bal.buildProblem()
    o = api.Solver.Options(); o.setLinearSolverType(_abi.ITERATIVE_SCHUR); o.setPreconditionerType(_abi.SCHUR_JACOBI)
    if len(sys.argv) > 2: o.setMaxNumIterations(int(sys.argv[2]))     # ill-conditioned shapes: compare the first rows only
    if use_comm: o.comm = comm
    o.residual_blocks_are_local = 1 if local else 0
    s = api.Solver.Summary()
    t = time.time(); api.ceres.solve(o, prob, s); dt = time.time() - t
    return bal.parameters.toArray(), s, dt

x, s, dt = solve(True)
rows = [(r.iteration, r.cost, r.linear_solver_iterations) for r in s.iterations]
print(f"rank {rank}: {s.message} final {s.final_cost:.9e} its {len(rows)} pcg {[r[2] for r in rows]} {dt:.3f}s gpus {s.num_gpus}", flush=True)
gathered = [None] * world
dist.all_gather_object(gathered, (x.tobytes(), rows))
# rank-local ingestion (each rank is handed only its own residual blocks): same bits for the trajectory, the cameras and
# this rank's points; foreign points are left as they were
xl, sl, dtl = solve(True, local=True)
rows_l = [(r.iteration, r.cost, r.linear_solver_iterations) for r in sl.iterations]
assert rows_l == rows, "rank-local ingestion changed the trajectory"
balr = api.BalProblem.fromArrays(d)
o0, o1 = balr.localRange(rank, world)
mine = np.unique(d.point_index[o0:o1])
nc9 = 9 * d.num_cameras
own = np.concatenate([np.arange(nc9), (nc9 + 3 * mine[:, None] + np.arange(3)).ravel()])
assert np.array_equal(xl[own], x[own]), "rank-local ingestion changed the solution"
other = np.setdiff1d(np.arange(x.size), own)
assert np.array_equal(xl[other], d.parameters[other]), "foreign points must stay untouched"
assert (sl.num_residual_blocks, sl.num_parameter_blocks) == (s.num_residual_blocks, s.num_parameter_blocks)
print(f"rank {rank}: rank-local ingestion OK ({o1 - o0} of {d.num_observations} observations ingested, {dtl:.3f}s vs {dt:.3f}s)", flush=True)
# failure injection: a NaN coordinate in a point only the LAST rank owns.  Its evaluation fails on that rank alone;
# every rank must report the same FAILURE and return -- a rank that terminated alone would leave the others in an allreduce.
import copy
dbad = copy.copy(d)
dbad.parameters = d.parameters.copy()
dbad.parameters[9 * d.num_cameras + 3 * (d.num_points - 1)] = float("nan")
for local_mode in (False, True):
    xb, sb, dtb = solve(True, local=local_mode, data=dbad)
    assert sb.termination_type == _abi.FAILURE and len(sb.iterations) == 0, (rank, sb.termination_type, sb.message)
print(f"rank {rank}: NaN on the last rank -> every rank terminates with FAILURE ({sb.message})", flush=True)
failed = None
if rank == 0:
  try:
      for g in gathered[1:]:
          assert g[0] == gathered[0][0], "ranks disagree on the solution"
          assert g[1] == gathered[0][1], "ranks disagree on the trajectory"
      x1, s1, dt1 = solve(False)
      rel = np.max(np.abs(x - x1) / np.maximum(np.abs(x1), 1e-2))
      print(f"single GPU: final {s1.final_cost:.9e} its {len(s1.iterations)} pcg {[r.linear_solver_iterations for r in s1.iterations]} {dt1:.3f}s")
      print(f"multi vs single: cost rel diff {abs(s.final_cost - s1.final_cost) / s1.final_cost:.2e}, max rel param diff {rel:.2e}")
      assert abs(s.final_cost - s1.final_cost) <= 1e-6 * s1.final_cost
      assert len(s.iterations) == len(s1.iterations)
      print("MULTI-GPU CHECK OK")
  except AssertionError as e:                  # rank 0 must still reach the barrier: the other ranks wait there
    import traceback
    traceback.print_exc()
    failed = e
dist.barrier()
dist.destroy_process_group()
if failed is not None:
    sys.exit(1)
