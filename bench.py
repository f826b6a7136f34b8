#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native LM solver core (contract: see DESIGN.md §Measurement).

A "step" is ONE Levenberg–Marquardt iteration (linear solve by ITERATIVE_SCHUR + SCHUR_JACOBI PCG,
candidate evaluation, step acceptance, and on acceptance the Jacobian re-evaluation) of a synthetic
Venice-1778-shaped bundle-adjustment problem (1,778 cameras, 993,923 points, 5,001,946 observations;
BASELINE.json configs[2]).  With N > 1 GPUs the problem is N Venice-sized shares (weak scaling, the
camera ring and the point set grow with N), point-partitioned across the ranks with NCCL allreduce
of the camera-block quantities.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU oracle on all host threads

value   = observations x LM iterations per second, timed on the device (CUDA events on the solver's
          stream) with the problem already resident in HBM; "lm_iterations_per_s" is in the same line.
e2e     = the same through the reference-facing API starting from HOST buffers: parameter upload,
          residual-block ingestion, preprocessing, K iterations, parameter download.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly ONE JSON line: NCCL's version banner (printed when the box sets NCCL_DEBUG) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from skeres_b200 import _abi, synth  # noqa: E402

METRIC = "BA LM observations/s (n_obs x LM iterations per second; ITERATIVE_SCHUR + SCHUR_JACOBI)"
UNIT = "obs*iter/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def shape_for(n_gpus, scale):
    c, p, o = synth.SHAPES["venice-1778"]
    f = n_gpus * scale
    return int(round(c * f)), int(round(p * f)), int(round(o * f))


def make_problem_data(n_gpus, scale, seed=1):
    c, p, o = shape_for(n_gpus, scale)
    return synth.make_bal(n_cam=c, n_pt=p, n_obs=o, seed=seed)


def matvec_algorithmic_bytes(n_obs, n_pts, n_cams):
    """one implicit-Schur product = O*(192 J + 4 idx) + P*72 + 2*C*72 bytes."""
    return n_obs * (192 + 4) + n_pts * 72 + 2 * n_cams * 72


# DRAM traffic of the dominant kernel per launch, from the committed `ncu --set full` capture of the SAME workload
# (Venice-1778 shape x 1 GPU): dram__bytes_read.sum + dram__bytes_write.sum of k_ba_matvec_tma, launch 0.
# bench.py cannot run ncu itself; the figure is only attached when the workload is the captured one.
MATVEC_NCU_TRAFFIC = {"bytes_per_launch": 1.058313e9 + 25.707008e6, "n_obs": 5001946,
                      "source": "profiles/r01_v9_matvec_tma_summary.md (ncu --set full, launch 0 of gpurun_out/prof_matvec_v9.ncu-rep)"}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows, self.proc, self.dev = [], None, device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_oracle(data, steps, warmup, threads, time_budget_s):
    """The CPU arm: the oracle (Ceres-algorithm restatement, NOT libceres) on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    p = O.OracleProblem(data.parameters)
    p.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, data.observations.reshape(-1, 2), data.block_offsets())
    o = _abi.default_options()
    o.linear_solver_type = _abi.ITERATIVE_SCHUR
    o.preconditioner_type = _abi.SCHUR_JACOBI
    o.max_num_iterations = steps
    o.max_solver_time_in_seconds = time_budget_s
    t = time.time()
    s = p.solve(o, threads=threads)
    wall = time.time() - t
    its = max(len(s.iterations) - 1, 0)
    t_steps = sum(r.iteration_time_in_seconds for r in s.iterations[1:])
    return its, t_steps, wall, s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the Venice-1778 shape per GPU (development only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="development: no per-kernel CUDA events in the timed region (roofline keys become meaningless)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = args.steps, max(args.warmup, 0)
    n_cam, n_pt, n_obs = shape_for(max(world, 1), args.scale)
    cfg = {"workload": f"synthetic Venice-1778-shaped BAL x{max(world, 1)} ({n_cam} cameras, {n_pt} points, {n_obs} observations), "
                       "ITERATIVE_SCHUR + SCHUR_JACOBI, trust-region LM, trivial loss, seed 1",
           "step": "one LM iteration", "parallelism": f"point-partitioned x{world}, cameras replicated" if world > 1 else "single GPU",
           "l2": "stored Jacobian (192 B/observation, %.2f GB per GPU) is larger than the 126 MB L2; no flush needed" % (192.0 * n_obs / max(world, 1) / 1e9)}

    # ------------------------------------------------------------------ reference (CPU oracle) arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        threads = os.cpu_count() or 1
        data = make_problem_data(max(world, 1), args.scale)
        its, t_steps, wall, s = run_oracle(data, K, W, threads, time_budget_s=150.0)
        val = n_obs * its / t_steps if t_steps > 0 else 0.0
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * t_steps / max(its, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg, "lm_iterations_per_s": its / t_steps if t_steps > 0 else 0.0,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "cpu_model": cpu_model(),
                                 "sample": f"first {its} of {K} LM iterations of the same full-size problem (150 s budget), "
                                           "CPU oracle = Ceres-algorithm restatement (libceres is not buildable here)"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "final_cost": s.final_cost, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ this repo's CUDA arm
    from skeres_b200 import api
    dist = None
    comm = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    api.check(api.lib.sk_set_device(local_rank))
    if world > 1:
        ids = [api.Communicator.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = api.Communicator(ids[0], rank, world)

    data = make_problem_data(max(world, 1), args.scale)
    bal = api.BalProblem.fromArrays(data)
    x0 = api.DoubleArray.fromArray(data.parameters)
    # N > 1: every rank ingests only the residual blocks of its own points (residual_blocks_are_local), all cameras declared
    problem = bal.buildLocalProblem(rank, world) if world > 1 else bal.buildProblem()
    opt = api.Solver.Options()
    opt.setLinearSolverType(_abi.ITERATIVE_SCHUR)
    opt.setPreconditionerType(_abi.SCHUR_JACOBI)
    opt.setMaxNumIterations(max(K, W, 1))
    opt.residual_blocks_are_local = 1 if world > 1 else 0
    # Events around EVERY launch cost ~10 % of the step at N = 2 (profiles/r01_multigpu_first_run.md), so the timed region
    # records them only around the dominant kernel (what the roofline needs); the full breakdown comes from a separate pass.
    opt.profile_kernels = 0 if args.no_profile else 2
    if comm is not None:
        opt.comm = comm
    solver = api.PreparedSolver(opt, problem)

    def barrier():
        if dist is not None:
            dist.barrier()

    def run_steps(n):
        """Exactly n LM iterations; a solve that terminates early is restarted from the start point."""
        done, dev_s, launches, kms, kl, last = 0, 0.0, 0, np.zeros(_abi.KF_COUNT), np.zeros(_abi.KF_COUNT), None
        while done < n:
            bal.parameters.copyFromArray(x0)
            s = solver.minimize(max_num_iterations=n - done)
            d = s.data
            it = max(d.num_iterations - 1, 0)
            dev_s += d.minimizer_device_time_in_seconds
            launches += d.num_kernel_launches
            kms += np.array(d.kernel_ms[:]); kl += np.array(d.kernel_launches[:])
            last = s
            if it == 0:
                break
            done += it
        return done, dev_s, launches, kms, kl, last

    if W > 0:
        run_steps(W)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_wall = time.time()
    done, dev_s, launches, kms, kl, last = run_steps(K)
    barrier()
    t_wall = time.time() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    # separate, NOT-timed-for-value pass with events around every family (same K steps) for the breakdown
    fam_ms, fam_launches = None, None
    if not args.no_profile:
        solver.close()
        opt.profile_kernels = 1
        solver = api.PreparedSolver(opt, problem)
        _, _, _, fam_ms, fam_launches, _ = run_steps(K)
        barrier()
    if dist is not None:
        import torch
        t = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s = float(t.item())
    value = n_obs * done / dev_s
    hbm_peak, peak_src = load_peaks()
    mv_launches = max(kl[3], 1)
    mv_ms = kms[3] / mv_launches
    per_gpu_obs, per_gpu_pts = n_obs / max(world, 1), n_pt / max(world, 1)
    mv_bytes = matvec_algorithmic_bytes(per_gpu_obs, per_gpu_pts, n_cam)
    achieved = mv_bytes / (mv_ms * 1e-3) / 1e9 if mv_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_ba_matvec_tma (implicit Schur product, PCG inner kernel)", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "frac_of_nominal_8000_GBps": achieved / 8000.0,
                "traffic": (MATVEC_NCU_TRAFFIC["bytes_per_launch"] if (world == 1 and n_obs == MATVEC_NCU_TRAFFIC["n_obs"]) else None),
                "traffic_source": MATVEC_NCU_TRAFFIC["source"], "peak_source": peak_src,
                "algorithmic_bytes_per_launch": mv_bytes, "avg_launch_ms": mv_ms, "launches": int(kl[3]),
                "share_of_step_device_time": float(kms[3] / (dev_s * 1e3)) if dev_s > 0 else None,
                "events_in_timed_region": "k_ba_matvec only",
                "kernel_family_ms": None if fam_ms is None else {_abi.KF_NAMES[i]: float(fam_ms[i]) for i in range(_abi.KF_COUNT)},
                "kernel_family_launches": None if fam_launches is None else {_abi.KF_NAMES[i]: int(fam_launches[i]) for i in range(_abi.KF_COUNT)},
                "kernel_family_note": "from a separate pass of the same K steps with events around every launch (not the pass `value` is timed on)"}

    # ------------------------------------------------------------------ end-to-end through the public API, host buffers
    e2e = None
    if not args.no_e2e:
        barrier()
        solver.close()
        del problem, bal
        host_params = data.parameters.copy()
        runs = []
        for rep in range(2):                                          # device allocation time varies 0.1-0.4 s from call to call:
            barrier()                                                 # two full passes, both reported, the faster one is `value`
            t0 = time.time()
            bal2 = api.BalProblem.fromArrays(data)                    # H2D: parameters
            problem2 = bal2.buildLocalProblem(rank, world) if world > 1 else bal2.buildProblem()   # residual-block ingestion (host)
            opt2 = api.Solver.Options()
            opt2.residual_blocks_are_local = 1 if world > 1 else 0
            opt2.setLinearSolverType(_abi.ITERATIVE_SCHUR)
            opt2.setPreconditionerType(_abi.SCHUR_JACOBI)
            opt2.setMaxNumIterations(K)
            if comm is not None:
                opt2.comm = comm
            summ = api.Solver.Summary()
            api.ceres.solve(opt2, problem2, summ)                     # preprocess + H2D layout/observations + K iterations
            out = bal2.parameters.toArray()                           # D2H: solution
            final_cost = summ.final_cost                              # D2H: summary
            barrier()
            runs.append((time.time() - t0, summ.preprocessor_time_in_seconds, max(summ.num_iterations - 1, 0), final_cost, out.nbytes))
            del problem2, bal2
        t1, pre_s, its, final_cost, out_bytes = min(runs)
        per_rank_obs = n_obs // max(world, 1)
        # observations (16 B) + tile-local ids (6 B) + per-tile metadata records of the prefetching matvec (~10 B) per observation,
        # point / camera tables, parameters
        h2d = host_params.nbytes + per_rank_obs * (16 + 6 + 10) + 4 * n_pt // max(world, 1) + 72 * n_cam
        d2h = out_bytes + 4096
        e2e = {"value": n_obs * its / t1, "unit": UNIT, "h2d_bytes_per_step": int(h2d / max(its, 1)), "d2h_bytes_per_step": int(d2h / max(its, 1)),
               "lm_iterations": its, "wall_s": t1, "wall_s_runs": [r[0] for r in runs], "preprocessor_s": pre_s, "final_cost": final_cost,
               "what": "DoubleArray upload + addResidualBlocks + ceres.solve (preprocess, layout upload, K LM iterations) + parameter download; "
                       "two full passes, the faster one reported" + ("; every rank ingests only the residual blocks of its own points "
                                                                    "(residual_blocks_are_local), all cameras declared" if world > 1 else "")}

    # ------------------------------------------------------------------ CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        its, t_steps, wall, s = run_oracle(data, 1, 0, threads, time_budget_s=25.0)
        cpu = {"value": n_obs * its / t_steps if t_steps > 0 else 0.0, "unit": UNIT, "cores": threads, "kind": "port", "cpu_model": cpu_model(),
               "sample": "first LM iteration (evaluate + SchurJacobi + PCG + candidate cost + re-evaluation) of the same full-size problem",
               "lm_iterations_per_s": its / t_steps if t_steps > 0 else 0.0}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": max(world, 1), "steps": K, "warmup": W,
                "ms_per_step": 1e3 * dev_s / max(done, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg, "lm_iterations_per_s": done / dev_s, "steps_timed": done,
                # CAUTION for scaling efficiency: weak scaling changes the PROBLEM with N and with it the number of PCG
                # iterations an LM iteration needs (measured: 742 executed matvecs per 10 LM iterations at N=1, 283 at
                # N=4), so value(N) / (N value(1)) mixes hardware scaling with a change of work per step (1.68 at N=4).
                # matvec_obs_per_s is affected too (the fixed per-LM-iteration work is amortised over fewer matvecs).
                # The quantity that IS comparable across N is the device time of one PCG iteration, pcg_iteration_ms.
                "matvec_obs_per_s": n_obs * float(kl[3]) / dev_s, "pcg_matvecs_timed": int(kl[3]),
                "pcg_iteration_ms": None if fam_ms is None or fam_launches[3] == 0 else
                    float((fam_ms[3] + fam_ms[4] + fam_ms[8]) / fam_launches[3]),
                "pcg_iteration_ms_note": "(k_ba_matvec + PCG vector kernels + allreduce) device time / executed matvecs, from the "
                                         "instrumented pass (N=1: 0.262 ms, profiles/r01_v10_bench_venice_n1.json)",
                "wall_s_timed_region": t_wall, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "clocks": clocks, "final_cost_last_solve": last.final_cost if last is not None else None,
                "pcg_iterations_last_solve": [r.linear_solver_iterations for r in last.iterations] if last is not None else None}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
