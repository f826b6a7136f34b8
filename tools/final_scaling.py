"""BASELINE.json configs[4]: ONE synthetic Final-13682-shaped bundle adjustment (13,682 cameras, 4,456,117 points, 28,987,644
observations), point-partitioned across the GPUs of one box -- strong scaling of a fixed problem.

    python tools/final_scaling.py --steps 6                                   # one GPU (the problem fits: 5.6 GB of Jacobian)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/final_scaling.py --steps 6

Every rank ingests only the residual blocks of its own points (residual_blocks_are_local), all cameras declared; the
per-PCG-iteration exchange goes through the NVLink peer window, the per-LM-iteration sums through NCCL.  Rank 0 prints one
JSON line: value = observations x LM iterations / s on the device (max over ranks), the per-family device times of an
instrumented pass, the device time of one PCG iteration, and -- N > 1 -- `multi_vs_single`: the same LM iterations on rank 0's
GPU alone (rows, PCG counts, costs).  --shape lets the same script run a smaller shape for development."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
from skeres_b200 import _abi, api, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--shape", default="final-13682")
    ap.add_argument("--no-single", action="store_true", help="skip the single-GPU comparison solve on rank 0")
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dist = comm = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api.check(api.lib.sk_set_device(local))
    if world > 1:
        ids = [api.Communicator.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = api.Communicator(ids[0], rank, world)
    t0 = time.time()
    if "," in args.shape:                                     # development: "cameras,points,observations" (e.g. one rank's share on one GPU)
        c_, p_, o_ = (int(v) for v in args.shape.split(","))
        d = synth.make_bal(n_cam=c_, n_pt=p_, n_obs=o_, seed=1)
    else:
        d = synth.make_bal(args.shape, seed=1)
    t_gen = time.time() - t0
    K, W = args.steps, args.warmup

    def options(max_it, profile, use_comm=True):
        o = api.Solver.Options()
        o.setLinearSolverType(_abi.ITERATIVE_SCHUR); o.setPreconditionerType(_abi.SCHUR_JACOBI); o.setMaxNumIterations(max_it)
        o.profile_kernels = profile
        if comm is not None and use_comm:
            o.comm = comm
            o.residual_blocks_are_local = 1
        return o

    def barrier():
        if dist is not None:
            dist.barrier()

    bal = api.BalProblem.fromArrays(d)
    x0 = api.DoubleArray.fromArray(d.parameters)
    t0 = time.time()
    problem = bal.buildLocalProblem(rank, world) if world > 1 else bal.buildProblem()
    solver = api.PreparedSolver(options(max(K, W), 2), problem)
    t_prep = time.time() - t0

    def run(slv, n):
        bal.parameters.copyFromArray(x0)
        return slv.minimize(max_num_iterations=n)

    if W > 0:
        run(solver, W)
    barrier()
    s = run(solver, K)
    barrier()
    its = max(s.data.num_iterations - 1, 0)
    dev_s = s.data.minimizer_device_time_in_seconds
    kms, kl = np.array(s.data.kernel_ms[:]), np.array(s.data.kernel_launches[:])
    solver.close()
    solver = api.PreparedSolver(options(max(K, W), 1), problem)
    sp = run(solver, K)
    fam_ms, fam_l = np.array(sp.data.kernel_ms[:]), np.array(sp.data.kernel_launches[:])
    barrier()
    solver.close()
    if dist is not None:
        import torch
        t = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s = float(t.item())
    mvs = None
    if rank == 0 and world > 1 and not args.no_single:
        bal1 = api.BalProblem.fromArrays(d)
        s1 = api.Solver.Summary()
        api.ceres.solve(options(K, 0, use_comm=False), bal1.buildProblem(), s1)
        pn, p1 = [r.linear_solver_iterations for r in s.iterations], [r.linear_solver_iterations for r in s1.iterations]
        mvs = {"lm_rows": [len(s.iterations), len(s1.iterations)], "pcg_iterations_n_gpus": pn, "pcg_iterations_1_gpu": p1, "pcg_counts_equal": pn == p1,
               "cost_rel_diff_per_row_max": max(abs(a.cost - b.cost) / b.cost for a, b in zip(s.iterations, s1.iterations)),
               "final_cost_rel_diff": abs(s.final_cost - s1.final_cost) / s1.final_cost,
               "single_gpu_device_s": s1.data.minimizer_device_time_in_seconds,
               "single_gpu_value": d.num_observations * max(s1.data.num_iterations - 1, 0) / s1.data.minimizer_device_time_in_seconds}
    barrier()
    if rank == 0:
        o_gpu, p_gpu = d.num_observations / world, d.num_points / world
        mv_bytes = o_gpu * 196 + p_gpu * 72 + d.num_cameras * 144
        fused = kl[9] > 0                                   # the PCG loop ran as one persistent kernel per linear solve (pcg_fused.cu)
        mv_ms = kms[3] / max(kl[3], 1)                      # fused: the product phases on the device clock of the first CTA
        pcg_ms = float(kms[9] / max(kl[3], 1)) if fused else float((fam_ms[3] + fam_ms[4] + fam_ms[8]) / max(fam_l[3], 1))
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        line = {"metric": "BA LM observations/s (n_obs x LM iterations per second; ITERATIVE_SCHUR + SCHUR_JACOBI)", "value": d.num_observations * its / dev_s,
                "unit": "obs*iter/s", "n_gpus": world, "steps": K, "warmup": W, "steps_timed": its, "ms_per_step": 1e3 * dev_s / max(its, 1), "scaling": "strong",
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"synthetic {args.shape}-shaped BAL ({d.num_cameras} cameras, {d.num_points} points, {d.num_observations} observations), "
                                       "ITERATIVE_SCHUR + SCHUR_JACOBI, trivial loss, seed 1", "parallelism": f"point-partitioned x{world}, cameras replicated, rank-local ingestion"},
                "pcg_iterations": [r.linear_solver_iterations for r in s.iterations], "final_cost": s.final_cost,
                "roofline": {"bound": "hbm", "kernel": "product phase of k_pcg_solve" if fused else "k_ba_matvec_tma", "avg_launch_ms": float(mv_ms), "launches": int(kl[3]), "algorithmic_bytes_per_launch": mv_bytes,
                             "achieved": mv_bytes / (mv_ms * 1e-3) / 1e9 if mv_ms > 0 else 0.0, "peak": hbm, "frac": (mv_bytes / (mv_ms * 1e-3) / 1e9 / hbm) if mv_ms > 0 else 0.0,
                             "pcg_iteration_ms": pcg_ms, "vector_phase_ms_per_product": float(kms[4] / max(kl[3], 1)) if fused else None,
                             "kernel_family_ms": {_abi.KF_NAMES[i]: float(fam_ms[i]) for i in range(_abi.KF_COUNT)},
                             "kernel_family_launches": {_abi.KF_NAMES[i]: int(fam_l[i]) for i in range(_abi.KF_COUNT)}},
                "multi_vs_single": mvs, "host_seconds": {"generate": t_gen, "ingest_and_prepare": t_prep}}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
