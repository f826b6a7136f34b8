#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native LM solver core (contract: see DESIGN.md §Measurement).

A "step" is ONE Levenberg–Marquardt iteration (linear solve by ITERATIVE_SCHUR + SCHUR_JACOBI PCG,
candidate evaluation, step acceptance, and on acceptance the Jacobian re-evaluation) of a synthetic
Venice-1778-shaped bundle-adjustment problem (1,778 cameras, 993,923 points, 5,001,946 observations;
BASELINE.json configs[2]).  With N > 1 GPUs the problem is ONE bundle adjustment made of N disjoint copies of that
scene (N x 1,778 cameras replicated on every rank, N x 993,923 points partitioned by point: every rank gets one
scene): its LM rows and PCG iteration counts are those of one scene by construction, so the work per step is
exactly N times the single-GPU step -- a weak-scaling measurement in which the algorithm does not change with N.
Everything multi-GPU still runs: the per-PCG-iteration exchange of the camera-sized vectors through the NVLink peer
window, the NCCL allreduces of gradient / reduced right-hand side / SchurJacobi blocks / scalars, and the PCG vector
kernels on all N x 1,778 replicated cameras.  After the timed region the N-GPU result is compared with a single-GPU
solve of one scene (`multi_vs_single`).  (BASELINE.json configs[4], one Final-13682-shaped problem strong-scaled
across the GPUs, is tools/final_scaling.py.)

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

# stdout carries exactly ONE JSON line.  NCCL's version banner (printed when the box sets NCCL_DEBUG) is written by native code
# straight to file descriptor 1 -- NCCL_DEBUG_FILE does not catch torch's bundled NCCL in every case (seen on a 2-GPU box) -- so
# descriptor 1 is pointed at stderr for the whole run and the JSON line goes to a duplicate of the original stdout.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.stdout.flush()
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


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


def make_problem_data(n_gpus, scale, seed=1, rank=0):
    """N = 1: the scene.  N > 1: this rank's share of N disjoint copies of the scene (synth.make_scene_share)."""
    if n_gpus <= 1:
        c, p, o = shape_for(1, scale)
        return synth.make_bal(n_cam=c, n_pt=p, n_obs=o, seed=seed)
    assert scale == 1.0, "--scale is a single-GPU development option"
    return synth.make_scene_share(rank, n_gpus, "venice-1778", seed=seed)


def path_algorithmic_bytes(n_obs, n_pts, n_cams, n_jac, n_cost, n_lin, n_matvec):
    """SURVEY.md section 8(d) summed over the per-iteration path (per GPU): Jacobian evaluations, residual-only evaluations,
    Schur set-ups (+ back-substitutions, one per linear solve) and implicit Schur products."""
    jac = 232 * n_obs + 72 * n_cams + 24 * n_pts
    cost = 40 * n_obs + 72 * n_cams + 24 * n_pts
    setup = 216 * n_obs + 96 * n_pts + 720 * n_cams
    back = 208 * n_obs + 72 * n_pts
    return n_jac * jac + n_cost * cost + n_lin * (setup + back) + n_matvec * matvec_algorithmic_bytes(n_obs, n_pts, n_cams)


def source_hash():
    """sha256 (first 16 hex digits) of the sources of the dominant kernel: the key of profiles/matvec_traffic.json."""
    import hashlib
    h = hashlib.sha256()
    for f in ("ba_product.cuh", "pcg_fused.cu", "pcg_device.cuh", "ba_tile_rec.h", "ba_kernels.cuh"):
        h.update(open(os.path.join(ROOT, "skeres_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(n_obs):
    """DRAM bytes PER PRODUCT of the dominant kernel from the committed `ncu --set full` capture of THIS kernel source on THIS
    workload (profiles/matvec_traffic.json, keyed by source_hash(): dram__bytes_read.sum + dram__bytes_write.sum of one
    k_pcg_solve launch divided by the products it executed); None -- with the reason -- when the kernel has changed since the
    last capture or the workload is another one.  bench.py cannot run ncu itself."""
    path = os.path.join(ROOT, "profiles", "matvec_traffic.json")
    key = source_hash()
    try:
        t = json.load(open(path)).get(key)
    except (OSError, ValueError):
        t = None
    if t is None:
        return None, f"no ncu capture committed for kernel source {key} (profiles/matvec_traffic.json)"
    if int(t["n_obs"]) != int(n_obs):
        return None, f"ncu capture is for n_obs = {t['n_obs']}"
    return float(t["bytes_per_product"]), t["source"]


def matvec_algorithmic_bytes(n_obs, n_pts, n_cams):
    """one implicit-Schur product = O*(192 J + 4 idx) + P*72 + 2*C*72 bytes."""
    return n_obs * (192 + 4) + n_pts * 72 + 2 * n_cams * 72


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
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the Venice-1778 shape (single-GPU development only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="development: no per-kernel CUDA events in the timed region (roofline keys become meaningless)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = args.steps, max(args.warmup, 0)
    N = max(world, 1)
    c1, p1, o1 = shape_for(1, args.scale)                    # one scene = one GPU's share
    n_cam, n_pt, n_obs = c1 * N, p1 * N, o1 * N
    cfg = {"workload": (f"synthetic Venice-1778-shaped BAL ({c1} cameras, {p1} points, {o1} observations)" if N == 1 else
                        f"ONE bundle adjustment of {N} disjoint copies of the synthetic Venice-1778-shaped scene ({n_cam} cameras replicated "
                        f"on every rank, {n_pt} points, {n_obs} observations; one scene per rank, so every camera is observed from one rank; "
                        f"LM rows and PCG counts equal one scene's by construction)") +
                       ", ITERATIVE_SCHUR + SCHUR_JACOBI, trust-region LM, trivial loss, seed 1",
           "step": "one LM iteration", "parallelism": f"point-partitioned x{world}, cameras replicated" if world > 1 else "single GPU",
           "l2": "stored Jacobian (192 B/observation, %.2f GB per GPU) is larger than the 126 MB L2; no flush needed" % (192.0 * o1 / 1e9)}

    # ------------------------------------------------------------------ reference (CPU oracle) arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = os.cpu_count() or 1
        data = make_problem_data(1, args.scale)              # one scene: N disjoint copies cost the CPU N times this, step by step
        budget = 170.0
        its, t_steps, wall, s = run_oracle(data, K, W, threads, time_budget_s=budget)
        # The GPU arm times exactly K LM iterations: one solve of this problem has fewer (13 from this start point), so it
        # restarts from the start point and times the first rows again.  The CPU arm prices THE SAME multiset of rows from the
        # per-row times of its one (bounded) solve instead of re-running identical work: rows 1..R once, rows 1..(K - R) again.
        row_s = [r.iteration_time_in_seconds for r in s.iterations[1:]]
        R = len(row_s)
        priced, k = [], 0
        while R > 0 and len(priced) < K:
            priced.append(row_s[k % R]); k += 1
        complete = s.termination_type == _abi.CONVERGENCE or R >= K    # False: the time budget cut the solve short, later rows are missing
        t_priced = sum(priced)
        val = N * o1 * len(priced) / (N * t_priced) if t_priced > 0 else 0.0    # N copies: N x the observations in N x the time
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * N * t_priced / max(len(priced), 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg, "lm_iterations_per_s": len(priced) / (N * t_priced) if t_priced > 0 else 0.0,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "cpu_model": cpu_model(),
                                 "sample": f"rows 1..{R} of one solve of one scene, each timed ({sum(row_s):.1f} s, budget {budget:.0f} s, "
                                           f"{'complete solve' if complete else 'cut by the budget'}); the {K}-step set of the GPU arm (a solve has {R} rows, "
                                           f"then it restarts) priced from those per-row times" + (f"; x{N}: N disjoint copies cost N times one scene" if N > 1 else "") +
                                           "; CPU oracle = Ceres-algorithm restatement (libceres is not buildable here)"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "rows_timed": R, "row_seconds": row_s, "final_cost": s.final_cost, "gpu_launches": 0}
        emit(line)
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

    data = make_problem_data(N, args.scale, rank=rank)

    def build(bal_):
        """N > 1: every rank ingests only the residual blocks of its own points (residual_blocks_are_local), all cameras declared."""
        return api.build_share_problem(bal_) if world > 1 else bal_.buildProblem()

    def options(max_it, profile):
        o = api.Solver.Options()
        o.setLinearSolverType(_abi.ITERATIVE_SCHUR)
        o.setPreconditionerType(_abi.SCHUR_JACOBI)
        o.setMaxNumIterations(max_it)
        o.residual_blocks_are_local = 1 if world > 1 else 0
        o.profile_kernels = profile
        if comm is not None:
            o.comm = comm
        return o

    bal = api.BalProblem.fromArrays(data)
    x0 = api.DoubleArray.fromArray(data.parameters)
    problem = build(bal)
    # Events around EVERY launch cost ~10 % of the step at N = 2 (profiles/r01_multigpu_first_run.md), so the timed region
    # records them only around the dominant kernel (what the roofline needs); the full breakdown comes from a separate pass.
    opt = options(max(K, W, 1), 0 if args.no_profile else 2)
    solver = api.PreparedSolver(opt, problem)

    def barrier():
        if dist is not None:
            dist.barrier()

    def run_steps(n, slv):
        """Exactly n LM iterations; a solve that terminates early is restarted from the start point."""
        acc = {"done": 0, "dev_s": 0.0, "launches": 0, "kms": np.zeros(_abi.KF_COUNT), "kl": np.zeros(_abi.KF_COUNT), "jac": 0, "cost": 0, "lin": 0, "last": None}
        while acc["done"] < n:
            bal.parameters.copyFromArray(x0)
            s = slv.minimize(max_num_iterations=n - acc["done"])
            d = s.data
            it = max(d.num_iterations - 1, 0)
            acc["dev_s"] += d.minimizer_device_time_in_seconds
            acc["launches"] += d.num_kernel_launches
            acc["kms"] += np.array(d.kernel_ms[:]); acc["kl"] += np.array(d.kernel_launches[:])
            # the Jacobian evaluation(s) of row 0 belong to the solve's start-up, not to an LM iteration: they ARE inside the
            # timed device time, so they are counted in the path's algorithmic bytes as well
            acc["jac"] += int(d.num_jacobian_evaluations) + (1 if opt.jacobi_scaling else 0)
            acc["cost"] += int(d.num_residual_evaluations); acc["lin"] += int(d.num_linear_solves)
            acc["last"] = s
            if it == 0:
                break
            acc["done"] += it
        return acc

    if W > 0:
        run_steps(W, solver)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_wall = time.time()
    A = run_steps(K, solver)
    barrier()
    t_wall = time.time() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    done, dev_s, launches, kms, kl, last = A["done"], A["dev_s"], A["launches"], A["kms"], A["kl"], A["last"]
    mv_alone_ms = None
    if world == 1 and not args.no_profile:
        mv_alone_ms = solver.timeSchurProduct(200)           # the dominant kernel back to back, nothing else on the GPU
    # separate, NOT-timed-for-value pass with events around every family (same K steps) for the breakdown
    fam_ms, fam_launches = None, None
    if not args.no_profile:
        solver.close()
        solver = api.PreparedSolver(options(max(K, W, 1), 1), problem)
        B = run_steps(K, solver)
        fam_ms, fam_launches = B["kms"], B["kl"]
        barrier()
    # ---- N-GPU result == single-GPU result: 3 LM iterations of the N-scene problem on N GPUs against one scene on one GPU
    multi_vs_single = None
    if world > 1:
        bal.parameters.copyFromArray(x0)
        sN = solver.minimize(max_num_iterations=3)
        barrier()
        solver.close()
        if rank == 0:
            one = synth.make_bal("venice-1778", seed=1)
            bal1 = api.BalProblem.fromArrays(one)
            o1_ = api.Solver.Options()
            o1_.setLinearSolverType(_abi.ITERATIVE_SCHUR); o1_.setPreconditionerType(_abi.SCHUR_JACOBI); o1_.setMaxNumIterations(3)
            s1 = api.Solver.Summary()
            api.ceres.solve(o1_, bal1.buildProblem(), s1)
            pc_n = [r.linear_solver_iterations for r in sN.iterations]
            pc_1 = [r.linear_solver_iterations for r in s1.iterations]
            row_rel = [abs(a.cost / N - b.cost) / b.cost for a, b in zip(sN.iterations, s1.iterations)]
            multi_vs_single = {"lm_rows": [len(sN.iterations), len(s1.iterations)], "pcg_iterations_n_gpus": pc_n, "pcg_iterations_1_gpu": pc_1,
                               "pcg_counts_equal": pc_n == pc_1, "cost_rel_diff_per_row_max": max(row_rel) if row_rel else None,
                               "final_cost_rel_diff": abs(sN.final_cost / N - s1.final_cost) / s1.final_cost,
                               "what": f"3 LM iterations of the {N}-scene problem on {N} GPUs (cost / {N}) against one scene on one GPU"}
            del bal1
        barrier()
    if dist is not None:
        import torch
        t = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s = float(t.item())
    value = n_obs * done / dev_s
    hbm_peak, peak_src = load_peaks()
    mv_bytes = matvec_algorithmic_bytes(o1, p1, n_cam)        # per GPU: its observations and points, every camera
    KF = {n: i for i, n in enumerate(_abi.KF_NAMES)}
    products = int(kl[KF["schur_matvec"]])                    # implicit-Schur products executed in the timed region
    fused = kl[KF["pcg_solve"]] > 0                           # the PCG loop ran as one persistent kernel per linear solve
    if fused:
        # dominant kernel = k_pcg_solve: CUDA events around each launch (one per linear solve); its algorithmic bytes are those of
        # the products it executed.  The product phase alone (first CTA's device clock, incl. the grid barrier that ends it) is
        # reported next to it.
        dom_ms, dom_launches = float(kms[KF["pcg_solve"]]), int(kl[KF["pcg_solve"]])
        dom_name = "k_pcg_solve (persistent fused PCG solve: implicit Schur products + vector phases behind grid barriers, one launch per linear solve)"
        mv_ms = dom_ms / max(products, 1)                     # per product, vector phases and barriers included
        product_phase_ms = float(kms[KF["schur_matvec"]]) / max(products, 1)
        vector_phase_ms = float(kms[KF["pcg_vector"]]) / max(products, 1)
    else:
        dom_ms, dom_launches = float(kms[KF["schur_matvec"]]), products
        dom_name = "k_ba_matvec_tma (implicit Schur product, PCG inner kernel)"
        mv_ms = dom_ms / max(products, 1)
        product_phase_ms = vector_phase_ms = None
    achieved = products * mv_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    path_bytes = path_algorithmic_bytes(o1, p1, n_cam, A["jac"], A["cost"], A["lin"], int(kl[3]))
    path_gbs = path_bytes / dev_s / 1e9 if dev_s > 0 else 0.0
    traffic_pp, traffic_src = ncu_traffic(o1) if world == 1 else (None, "single-GPU capture only")
    if fused:
        pcg_ms = mv_ms
    else:
        pcg_ms = None if fam_ms is None or fam_launches[3] == 0 else float((fam_ms[3] + fam_ms[4] + fam_ms[8]) / fam_launches[3])
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "frac_of_nominal_8000_GBps": achieved / 8000.0,
                # per launch like `achieved`: the capture's bytes per product x the products per launch of this run
                "traffic": None if traffic_pp is None else traffic_pp * products / max(dom_launches, 1),
                "traffic_per_product": traffic_pp, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": products * mv_bytes / max(dom_launches, 1), "avg_launch_ms": dom_ms / max(dom_launches, 1),
                "launches": dom_launches,
                "algorithmic_bytes_per_product": mv_bytes, "products": products, "ms_per_product": mv_ms,
                "product_phase_ms": product_phase_ms, "vector_phase_ms_per_product": vector_phase_ms,
                "product_phase_frac": None if not product_phase_ms else mv_bytes / (product_phase_ms * 1e-3) / 1e9 / hbm_peak,
                "phase_note": "inside k_pcg_solve: time its first CTA spent in the product passes (incl. the grid barrier that ends each) / in the "
                              "vector phases, on the device clock (globaltimer); CUDA events cannot split one launch" if fused else None,
                "avg_launch_ms_back_to_back": mv_alone_ms,
                "share_of_step_device_time": float(dom_ms / (dev_s * 1e3)) if dev_s > 0 else None,
                "events_in_timed_region": "k_pcg_solve only" if fused else "k_ba_matvec only",
                # the whole per-iteration path (evaluate + eliminate + PCG + back-substitution + LM control), per GPU:
                # algorithmic bytes of every evaluation / set-up / product executed in the timed region over its device time
                "path_frac": path_gbs / hbm_peak, "path_achieved_gbs": path_gbs, "path_algorithmic_bytes": path_bytes,
                "path_counts": {"jacobian_evaluations": A["jac"], "residual_evaluations": A["cost"], "linear_solves": A["lin"], "schur_products": int(kl[3])},
                "matvecs_per_step": float(kl[3]) / max(done, 1),
                "pcg_iteration_ms": pcg_ms,
                "pcg_iteration_ms_note": "(implicit Schur product + PCG vector phases + exchange) device time per executed product"
                                         + (" = k_pcg_solve event time / products, timed region; " if fused else ", instrumented pass; ") +
                                         "comparable across N because the N-scene workload executes the same products per step at every N",
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
            problem2 = build(bal2)                                    # residual-block ingestion (host)
            opt2 = options(K, 0)
            summ = api.Solver.Summary()
            api.ceres.solve(opt2, problem2, summ)                     # preprocess + H2D layout/observations + K iterations
            out = bal2.parameters.toArray()                           # D2H: solution
            final_cost = summ.final_cost                              # D2H: summary
            barrier()
            runs.append((time.time() - t0, summ.preprocessor_time_in_seconds, max(summ.num_iterations - 1, 0), final_cost, out.nbytes))
            del problem2, bal2
        t1, pre_s, its, final_cost, out_bytes = min(runs)
        # what actually crosses: the parameters, the residual-block table as it is (16 B of block offsets + 16 B of observation per
        # residual block: the layout is built from it on the device) and the tile table packed on the host (12 B per tile);
        # back: the point CSR the tile packing reads (4 B per point), the camera table, the solution, the summary
        h2d = host_params.nbytes + o1 * (16 + 16) + 12 * (o1 // 200)
        d2h = out_bytes + 4 * p1 + 12 * n_cam + 4096
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
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * dev_s / max(done, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg, "lm_iterations_per_s": done / dev_s, "steps_timed": done,
                "matvec_obs_per_s": n_obs * float(kl[3]) / dev_s, "pcg_matvecs_timed": int(kl[3]),
                "wall_s_timed_region": t_wall, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "multi_vs_single": multi_vs_single,
                "clocks": clocks, "final_cost_last_solve": last.final_cost if last is not None else None,
                "pcg_iterations_last_solve": [r.linear_solver_iterations for r in last.iterations] if last is not None else None}
        emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
