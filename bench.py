#!/usr/bin/env python
"""Benchmark of the B200-native refinement hot path.

  python bench.py --gpus N --steps K --warmup W            (our arm)
  python bench.py --impl reference --gpus N --steps K ...  (CPU arm: the oracle port on host cores)
For N > 1 launch through torch.distributed.run (one rank per GPU).

A "step" is one fused residual + Jacobian + J^T J pass (set-up kernels, K1,
assembly/reduction, and the NCCL allreduce of the per-camera blocks when N > 1)
over the whole workload with the observations resident in HBM.  The workload is
BASELINE.json configs[4]: optimize_bundle, 8 cameras x 100 000 views x 88 corners
(70.4 M observations, 2.25 GB of SoA observations >> the 126 MB L2, so no L2
flush is needed between steps); it is sharded by views across the N ranks
(strong scaling).  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_cams, n_poses, description)
    "c5": (8, 100000, "optimize_bundle<pinhole+BC5>: 8 cameras x 100000 views x 88 corners (70.4M observations), joint intrinsics + hand-eye + target pose"),
    "c4": (4, 5000, "optimize_bundle<pinhole+BC5>: 4 cameras x 5000 robot poses x 88 corners (1.76M observations)"),
    "mini": (8, 2000, "optimize_bundle<pinhole+BC5>: 8 cameras x 2000 views x 88 corners (smoke size)"),
}
OTHER_WORKLOADS = ("c1", "c2", "c3", "c4", "c4-axxb")   # BASELINE configs[0]-[3]: bench_workloads.py (single GPU)
CHUNK = 12500  # robot poses per generation chunk (8 chunks for c5)


K1_NCU_CSV = os.path.join("profiles", "r2_k1_fused_ncu_full_bench_c5.csv")   # ncu --set full of this command on the current kernel (tools/ncu_summary.py)


def k1_ncu_record(args, world):
    """What the committed `ncu --set full` capture of K1 inside this very command says (C5, one GPU, free intrinsics; None for any
    other shape): DRAM bytes per launch (`roofline.traffic`), the capture's own duration (cold, serialised: compare shares) and the
    executed FP64 instruction counts, from which the flop per observation of the FP64 roofline follow (DFMA = 2, DMUL / DADD = 1)."""
    if args.workload != "c5" or world != 1 or args.fixed_intrinsics:
        return None
    path = os.path.join(ROOT, K1_NCU_CSV)
    if not os.path.exists(path):
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ms": 1.0}
    rec = {}
    for line in open(path):
        parts = line.rstrip("\n").rsplit(",", 3)
        if len(parts) == 4 and parts[1] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
                                            "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
                                            "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
                                            "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "sm__cycles_elapsed.avg",
                                            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"):
            try:
                rec[parts[1]] = float(parts[3]) * scale.get(parts[2], 1.0)
            except ValueError:
                pass
    if "dram__bytes_read.sum" not in rec:
        return None
    n_obs = WORKLOADS["c5"][0] * WORKLOADS["c5"][1] * 88
    out = {"traffic": rec["dram__bytes_read.sum"] + rec.get("dram__bytes_write.sum", 0.0), "capture_ms": rec.get("gpu__time_duration.sum"),
           "fp64_pipe_active_pct": rec.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), "source": K1_NCU_CSV}
    cyc = rec.get("sm__cycles_elapsed.avg")   # thread-level instruction counts = rate per elapsed cycle x elapsed cycles
    dfma, dmul, dadd = ((rec.get(f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed") or 0.0) * (cyc or 0.0) for k in ("dfma", "dmul", "dadd"))
    if dfma:
        out["fp64_inst_per_obs"] = {"dfma": dfma / n_obs, "dmul": dmul / n_obs, "dadd": dadd / n_obs}
        out["flop_per_obs"] = (2 * dfma + dmul + dadd) / n_obs
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True); self._t.start(); return self

    def __exit__(self, *a):
        self._stop.set(); self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def shard_chunks(n_poses, rank, world):
    from calibration_b200 import sharding
    return sharding.chunk_shard(n_poses, CHUNK, rank, world)


def run_b200(args, rank, world, local_rank):
    # the contract is ONE JSON line on stdout: anything libraries print there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        out = _run_b200(args, rank, world, local_rank)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    if out is not None:
        print(json.dumps(out), flush=True)


_STAGE = {"name": "start", "t0": time.time()}


def stage(name, rank=0):
    """Progress marker on stderr (one line per stage and rank) + what the watchdog reports if a stage never ends."""
    _STAGE["name"] = name
    sys.stderr.write(f"[bench r{rank} +{time.time() - _STAGE['t0']:.1f}s] {name}\n"); sys.stderr.flush()


def start_watchdog(limit_s, rank):
    """A multi-rank run must never hang the box: past the limit the process says where it was and exits non-zero."""
    def run():
        time.sleep(limit_s)
        sys.stderr.write(f"[bench r{rank}] WATCHDOG: still in stage '{_STAGE['name']}' after {limit_s:.0f} s, aborting\n"); sys.stderr.flush()
        os._exit(3)
    threading.Thread(target=run, daemon=True).start()


def _run_b200(args, rank, world, local_rank):
    import torch
    start_watchdog(args.max_seconds, rank)
    stage("import / init", rank)
    from calibration_b200 import abi, capi, synth
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_cams, n_poses, desc = WORKLOADS[args.workload]
    chunks = shard_chunks(n_poses, rank, world)
    stage("generate synthetic shard", rank)
    t0 = time.time()
    prob, x0, xgt = synth.make_bundle(seed=137, n_cams=n_cams, n_poses=n_poses, chunk=CHUNK, chunks=chunks, pinned=True,
                                      optimize_intrinsics=not args.fixed_intrinsics)
    gen_s = time.time() - t0
    n_obs_local = int(prob.desc.n_obs)
    n_obs_total = n_obs_local * world
    obs_bytes_local = 32 * n_obs_local + 96 * int(prob.desc.n_blocks)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stage("communicator", rank)
    comm = None
    if dist is not None:  # process-lifetime communicator, like the CUDA context: created once, outside every timed region
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = capi.Comm(uid[0], rank, world, local_rank)
        if not args.no_peer_allreduce:   # NVLink peer-memory all-reduce for the 16 KB blocks; falls back to NCCL by itself

            def all_gather(b):
                out = [None] * world
                dist.all_gather_object(out, b)
                return out
            peer = comm.enable_peer(all_gather)
            stage(f"peer all-reduce {'enabled' if peer else 'unavailable (' + comm.peer_error + '), NCCL in use'}", rank)

    def make_handle(p=None):
        h = capi.RefineHandle(prob if p is None else p, device=local_rank)
        if comm is not None:
            h.attach_comm(comm)
        return h

    # the end-to-end arm uploads the observations in the shared-board form of cal_problem_desc (one board + u, v per
    # observation: what the C++ adapter sends when the views share a board, as every view of a calibration does); the
    # per-observation form is timed beside it by the N = 1 probe (e2e.shared_board.per_observation)
    prob_e2e = prob.with_shared_board()
    e2e_h2d_bytes = 16 * n_obs_local + 16 * int(prob_e2e.desc.board_n) + 96 * int(prob.desc.n_blocks)

    stage("create handle", rank)
    h = make_handle()
    info = h.layout_info()
    stage("warm-up + timed passes", rank)
    with ClockSampler(local_rank) as clk:   # clocks are sampled under load: warm-up, timed region and the cost passes
        for _ in range(max(args.warmup, 3)):
            h.bench_pass(x0, reps=1, jacobian=True)
        barrier()
        l0 = h.launch_count()
        ms_total, ms_k1, cost = h.bench_pass(x0, reps=args.steps, jacobian=True)
        launches_timed = h.launch_count() - l0
        ms_setup0, ms_between = h.bench_breakdown()
        barrier()
        ms_c, ms_ck, _ = h.bench_pass(x0, reps=args.steps, jacobian=False)
        t = torch.tensor([ms_total, ms_k1], dtype=torch.float64, device=f"cuda:{local_rank}")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_k1 = float(t[0]), float(t[1])
        # keep the same kernels running ~1.5 s more so the 200 ms clock sampler sees the GPU under this load; the
        # repeat count comes from the all-reduced time, so every rank issues the same number of collectives
        h.bench_pass(x0, reps=max(args.steps, int(1500.0 / max(ms_total / args.steps, 0.05))), jacobian=True)
    ms_step = ms_total / args.steps
    value = n_obs_total / (ms_step * 1e-3)

    stage("fp64 peak microbenchmark", rank)
    fp64_peak = capi.fp64_peak_tflops(local_rank)
    stage("end-to-end solve", rank)

    # ---- end to end through the C ABI from pinned HOST buffers: create (H2D of all observations +
    # layout), LM solve with covariance, results back to the host, destroy ----
    h.close()
    opts = abi.OptimOptions.default(compute_covariance=1)
    barrier()
    t0 = time.perf_counter()
    h2 = make_handle(prob_e2e)
    x_fin, res, cov = h2.solve(x0, opts)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    n_jac, n_cost = int(res.num_jac_evals), int(res.num_cost_evals)  # the covariance reuses the last Jacobian pass
    e2e_value = n_obs_total * n_jac / e2e_s
    launches_e2e = h2.launch_count()
    solve_err = float(np.abs(x_fin - xgt).max())
    h2.close()
    # the strictest reading of "end to end": upload ALL inputs and run ONE fused pass, result (cost, g, H) back on the host
    barrier()
    t0 = time.perf_counter()
    h3 = make_handle(prob_e2e)
    c1, g1, H1 = h3.eval(x0)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    one_pass_s = time.perf_counter() - t0
    h3.close()
    to = torch.tensor([one_pass_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(to, op=dist.ReduceOp.MAX)
    one_pass_s = float(to[0])
    stage("teardown", rank)
    if comm is not None:
        comm.close()
    if dist is not None:
        dist.destroy_process_group()

    if rank != 0:
        return None
    pk, pk_src = peaks()
    hbm_peak = float(pk["hbm_gbs"])
    k1_ms_launch = ms_k1 / args.steps
    achieved_gbs = obs_bytes_local / (k1_ms_launch * 1e-3) / 1e9
    # FP64 work of K1 per observation, from the SASS instruction mix of the committed ncu capture
    # (profiles/): DFMA counts 2 flops, DMUL/DADD 1; see DESIGN.md §5.
    ncu_rec = k1_ncu_record(args, world) or {}
    flop_per_obs = args.k1_flop_per_obs or ncu_rec.get("flop_per_obs") or 560.0
    k1_tflops = flop_per_obs * n_obs_local / (k1_ms_launch * 1e-3) / 1e12
    out = {
        "metric": "observations/s in residual+Jacobian+JtJ pass",
        "value": value, "unit": "observations/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "baseline_config": "configs[4]" if args.workload == "c5" else args.workload,
                   "observations_total": n_obs_total, "observations_per_gpu": n_obs_local,
                   "optimize_intrinsics": not args.fixed_intrinsics, "huber_delta": 1.0, "noise_px": 0.2,
                   "l2_policy": "inputs (2.25 GB SoA at c5) larger than the 126 MB L2; no flush",
                   "sharding": (f"views split contiguously over {world} rank(s); all-reduce of the per-camera blocks each pass: "
                                + ("one NVLink peer-memory kernel" if (comm is not None and getattr(comm, "peer", False)) else "NCCL")) if world > 1 else "single GPU",
                   "k1_segments": info["n_segments"], "k1_passes": info["k1_passes"], "local_entries": info["local_entries"]},
        "clocks": clk.summary(),
        "e2e": {"value": e2e_value, "unit": "observations/s", "h2d_bytes_per_step": e2e_h2d_bytes,
                "d2h_bytes_per_step": int(8 * (len(x_fin) + len(x_fin) ** 2)),
                "what": "cal_refine_create (H2D of all observations from pinned host memory in the shared-board form: one board + u, v per observation; layout) "
                        "+ cal_refine_solve (LM, covariance) + destroy; value = observations x fused passes executed / wall time",
                "wall_s": e2e_s, "lm_iterations": int(res.iterations), "jacobian_passes": n_jac, "cost_passes": n_cost,
                "lm_iteration_ms": 1e3 * e2e_s / max(int(res.iterations), 1), "final_cost": float(res.final_cost),
                "converged": bool(res.success), "max_abs_param_error_vs_ground_truth": solve_err, "gpu_launches": launches_e2e,
                "upload_plus_one_pass": {"value": n_obs_total / one_pass_s, "unit": "observations/s", "wall_s": one_pass_s,
                                         "what": "cal_refine_create (H2D of all observations) + ONE cal_refine_eval (fused pass, cost / g / H to the host) + destroy: "
                                                 "PCIe-bound, the bound a caller pays who uploads for a single evaluation"}},
        "gpu_launches": launches_timed,
        "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                     "traffic": ncu_rec.get("traffic"), "traffic_source": (K1_NCU_CSV + " (ncu --set full of this command on this kernel, not re-measured by this run; "
                                                                           f"the capture's own launch took {ncu_rec.get('capture_ms')} ms)") if ncu_rec else None,
                     "peak_source": pk_src, "kernel": "k1_kernel", "kernel_ms_per_launch": k1_ms_launch,
                     "kernel_share_of_step": ms_k1 / ms_total,
                     "rest_of_step_ms": {"first_setup_kernel": ms_setup0, "reduction_and_allreduce_plus_next_setup": ms_between,
                                         "note": "CUDA events on rank 0's stream; on several GPUs the reduction kernel waits for the slowest rank"},
                     "algorithmic_bytes_per_launch": obs_bytes_local,
                     "binding_roof": "fp64",
                     "fp64": {"achieved": k1_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": k1_tflops / fp64_peak,
                              "flop_per_observation": flop_per_obs, "peak_source": "DFMA-chain microbenchmark run in this process",
                              "flop_source": (K1_NCU_CSV + ": executed DFMA x 2 + DMUL + DADD per observation") if ncu_rec.get("flop_per_obs") and not args.k1_flop_per_obs
                              else "--k1-flop-per-obs" if args.k1_flop_per_obs else "default (237 DFMA + 44 DMUL + 42 DADD per observation)",
                              # the step loop issues one 64-bit shared-memory instruction per 6.2 FP64 instructions (the exchange of the Jacobian rows
                              # between the two threads that split a block's 136-entry system); on B200 such an instruction costs the FP64 pipe
                              # 1.6 DFMA issue slots at ANY occupancy (tools/ubench_fp64_mix.cu, profiles/r2_ubench_fp64.txt), so this mix cannot
                              # exceed 1 / (1 + 1.6 / 6.2) of the DFMA peak
                              "instruction_mix_bound_frac": 1.0 / (1.0 + 1.6 / 6.2)}},
        "cost_pass": {"ms_per_step": ms_c / args.steps, "kernel_ms": ms_ck / args.steps,
                      "value": n_obs_total / (ms_c / args.steps * 1e-3),
                      "hbm_frac": (obs_bytes_local / (ms_ck / args.steps * 1e-3) / 1e9) / hbm_peak},
        "setup": {"generate_s": gen_s},
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, steps=3)
    if world == 1 and not args.no_board_probe:
        stage("shared-board probe (child process)", rank)
        out["e2e"]["shared_board"] = shared_board_probe(args)
        stage("C++ adapter probe: optimize_bundle(std::vector<BundleObservation>) (child process)", rank)
        out["e2e"]["from_reference_types"] = cpp_adapter_probe(args)
    return out


def cpp_adapter_probe(args):
    """examples/cpp_bundle_e2e.cpp in a child process: calib::optimize_bundle from the reference's own AoS input type at this
    workload's size, AoS packing inside the timed region (no Python, no torch in that process)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import cpp_host_build as B
        exe, env = B.build_bundle_e2e()
        n_poses = WORKLOADS[args.workload][1]
        p = subprocess.run([exe, str(n_poses), "3"], capture_output=True, text=True, timeout=300, env=env)
        lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
        if not lines:
            return {"error": f"exit {p.returncode}: {(p.stderr or p.stdout).strip()[-300:]}"}
        rec = json.loads(lines[-1])
        walls = [r["wall_s"] for r in rec["runs"] if "wall_s" in r]
        if walls:
            rec["wall_s_best"] = min(walls)
            rec["wall_s_first"] = walls[0]
            rec["note"] = "the first run page-locks the staging pool; later runs reuse it (a long-running caller's steady state)"
        rec["exit_code"] = p.returncode
        return rec
    except Exception as e:  # noqa: BLE001 - a probe: its failure is recorded, never raised
        return {"error": repr(e)}


def run_board_probe(args):
    """Child process of the N = 1 run (`--probe-shared-board`): the same end-to-end solve with the observations in the
    shared-board form of cal_problem_desc (one board of 88 points + u, v per observation: 16 instead of 32 bytes per
    observation over PCIe), next to the per-observation form in the same process, and whether the two solves are
    bitwise identical.  A separate process so that nothing here can disturb the main record."""
    import torch
    from calibration_b200 import abi, capi, synth
    n_cams, n_poses, _ = WORKLOADS[args.workload]
    prob, x0, _ = synth.make_bundle(seed=137, n_cams=n_cams, n_poses=n_poses, chunk=CHUNK, chunks=shard_chunks(n_poses, 0, 1), pinned=True,
                                    optimize_intrinsics=not args.fixed_intrinsics)
    pb = prob.with_shared_board()
    opts = abi.OptimOptions.default(compute_covariance=1)
    n_obs = int(prob.desc.n_obs)

    def solve(p):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h = capi.RefineHandle(p)
        x, res, cov = h.solve(x0, opts)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        h.close()
        return dt, x, res, cov

    def one_pass(p):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h = capi.RefineHandle(p)
        c, g, H = h.eval(x0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        h.close()
        return dt, c

    solve(prob)                                  # warm-up: context, module load, memory-pool growth
    t_o, x_o, r_o, cov_o = solve(prob)
    t_b, x_b, r_b, cov_b = solve(pb)
    p_o, c_o = one_pass(prob)
    p_b, c_b = one_pass(pb)
    same = bool(np.array_equal(x_o, x_b) and np.array_equal(cov_o, cov_b) and r_o.iterations == r_b.iterations
                and r_o.final_cost == r_b.final_cost and c_o == c_b)
    n_jac = int(r_b.num_jac_evals)
    print(json.dumps({
        "what": "the e2e solve of this record with the observations in the shared-board form (cal_problem_desc.board_n = 88): "
                "create (H2D of one board + u, v) + LM solve with covariance + destroy; per_observation is the default form timed in the same process",
        "h2d_bytes": 16 * n_obs + 16 * int(pb.desc.board_n) + 96 * int(pb.desc.n_blocks),
        "wall_s": t_b, "value": n_obs * n_jac / t_b, "unit": "observations/s", "jacobian_passes": n_jac, "converged": bool(r_b.success),
        "upload_plus_one_pass_wall_s": p_b, "upload_plus_one_pass_value": n_obs / p_b,
        "per_observation": {"wall_s": t_o, "value": n_obs * int(r_o.num_jac_evals) / t_o, "upload_plus_one_pass_wall_s": p_o},
        "bitwise_identical_to_per_observation_form": same}), flush=True)


def shared_board_probe(args):
    """Runs run_board_probe in a child process and returns its record (or why there is none)."""
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), "--probe-shared-board", "--workload", args.workload]
    if args.fixed_intrinsics:
        cmd.append("--fixed-intrinsics")
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=150)
        lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
        if p.returncode != 0 or not lines:
            return {"error": f"exit {p.returncode}: {p.stderr.strip()[-300:]}"}
        return json.loads(lines[-1])
    except Exception as e:  # noqa: BLE001 - a probe: its failure is recorded, never raised
        return {"error": repr(e)}


def cpu_sample_problem(args):
    """Bounded sample of the workload for the CPU arm: the first 1/SAMPLE of the views."""
    from calibration_b200 import synth
    n_cams, n_poses, _ = WORKLOADS[args.workload]
    n_sample = max(64, n_poses // args.cpu_sample_div)
    prob, x0, xgt = synth.make_bundle(seed=137, n_cams=n_cams, n_poses=n_poses, chunk=n_sample, chunks=[0],
                                      optimize_intrinsics=not args.fixed_intrinsics)
    return prob, x0, n_sample


def cpu_baseline(args, steps):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    prob, x0, n_sample = cpu_sample_problem(args)
    cores = os.cpu_count() or 1
    dt, dta = time_cpu_passes(O, prob, x0, cores, steps)
    n = int(prob.desc.n_obs)
    # `value` is the FASTER of the two CPU passes (the conservative baseline: analytic derivatives); the dual-number
    # restatement — what the reference's AutoDiffCostFunction evaluates — is reported beside it
    return {"value": n / min(dt, dta), "unit": "observations/s", "cores": cores, "kind": "port",
            "sample": f"first {n_sample} views x {prob.desc.n_cams} cameras x 88 corners = {n} observations, {steps} fused passes, OpenMP over residual blocks",
            "what": "hand-derived Jacobians (oracle/analytic_pass.cpp)" if dta <= dt else "forward-mode duals (oracle/refine.cpp)",
            "ms_per_pass": min(dt, dta) * 1e3,
            "analytic_jacobians": {"value": n / dta, "ms_per_pass": dta * 1e3},
            "autodiff_duals": {"value": n / dt, "ms_per_pass": dt * 1e3, "note": "Jet width 24, as the reference's AutoDiffCostFunction<BundleReprojResidual>"},
            "lm_solve": cpu_lm_solve(O, prob, x0)}


def time_cpu_passes(O, prob, x0, cores, steps):
    """seconds per fused pass of the two CPU implementations on all host cores: (forward-mode duals, analytic Jacobians)"""
    O.refine_eval(prob, x0, jac=True, threads=cores)  # warm
    t0 = time.perf_counter()
    for _ in range(steps):
        O.refine_eval(prob, x0, jac=True, threads=cores)
    dt = (time.perf_counter() - t0) / steps
    O.analytic_bundle_eval(prob, x0, threads=cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.analytic_bundle_eval(prob, x0, threads=cores)
    dta = (time.perf_counter() - t0) / steps
    return dt, dta


def cpu_lm_solve(O, prob, x0):
    """The whole LM solve (with covariance) of the CPU restatement on the same bounded sample: the counterpart of the
    GPU arm's e2e.wall_s / lm_iteration_ms (SURVEY 8d asks for per-iteration and total solve time of the CPU path)."""
    t0 = time.perf_counter()
    _, r, _ = O.refine_solve(prob, x0)
    dt = time.perf_counter() - t0
    n_jac = int(r.num_jac_evals)
    return {"wall_s": dt, "iterations": int(r.iterations), "jacobian_passes": n_jac, "cost_passes": int(r.num_cost_evals),
            "ms_per_iteration": 1e3 * dt / max(int(r.iterations), 1), "converged": bool(r.success),
            "value": int(prob.desc.n_obs) * n_jac / dt, "unit": "observations/s (observations x fused passes / solve wall time, as e2e.value)"}


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference path on the host cores (Ceres cannot be built here)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    prob, x0, n_sample = cpu_sample_problem(args)
    cores = os.cpu_count() or 1
    dt_dual, dt_ana = time_cpu_passes(O, prob, x0, cores, args.steps)
    dt = min(dt_dual, dt_ana)   # the conservative (faster) CPU pass is the arm's value
    n_obs = int(prob.desc.n_obs)
    value = n_obs / dt
    n_cams, n_poses, desc = WORKLOADS[args.workload]
    cb = {"value": value, "unit": "observations/s", "cores": cores, "kind": "port",
          "sample": f"first {n_sample} views x {n_cams} cameras x 88 corners = {n_obs} observations per step",
          "what": "hand-derived Jacobians (oracle/analytic_pass.cpp)" if dt_ana <= dt_dual else "forward-mode duals (oracle/refine.cpp)",
          "analytic_jacobians": {"value": n_obs / dt_ana, "ms_per_pass": dt_ana * 1e3},
          "autodiff_duals": {"value": n_obs / dt_dual, "ms_per_pass": dt_dual * 1e3},
          "lm_solve": cpu_lm_solve(O, prob, x0)}
    print(json.dumps({
        "impl": "reference", "metric": "observations/s in residual+Jacobian+JtJ pass", "value": value, "unit": "observations/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "note": "CPU restatement (oracle) of the reference's Ceres path; Ceres/Eigen are absent from this image"},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "observations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(set(WORKLOADS) | set(OTHER_WORKLOADS) | {"c4-pass"}),
                    help="c5 (default, BASELINE configs[4]); c1 / c2 / c3 / c4 / c4-axxb: the other named shapes, single GPU")
    ap.add_argument("--ransac-problems", type=int, default=100000)
    ap.add_argument("--axxb-poses", type=int, default=5000)
    ap.add_argument("--fixed-intrinsics", action="store_true", help="BundleOptions default (optimize_intrinsics=false)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-peer-allreduce", action="store_true", help="use NCCL for the per-pass all-reduce instead of the NVLink peer-memory kernel")
    ap.add_argument("--max-seconds", type=float, default=420.0, help="watchdog: abort (exit 3) instead of hanging past this many seconds")
    ap.add_argument("--cpu-sample-div", type=int, default=64)
    ap.add_argument("--no-board-probe", action="store_true", help="skip the shared-board e2e probe (N = 1 only; runs in a child process)")
    ap.add_argument("--probe-shared-board", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--k1-flop-per-obs", type=float, default=0.0,
                    help="override the FP64 flop per observation of K1 (default: executed DFMA x 2 + DMUL + DADD of the committed ncu capture, "
                         "profiles/r2_k1_fused_ncu_full_bench_c5.csv, epilogue included; 560 when that file is absent)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.workload == "c4-pass":   # the bundle pass alone at C4 size through the C5 code path (strong-scaling runs)
        args.workload = "c4"
    elif args.workload in OTHER_WORKLOADS:
        import bench_workloads as W
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps(W.run_reference(args)), flush=True)
            return
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)   # one JSON line on stdout: library chatter goes to stderr
        try:
            out = W.run(args, ClockSampler, rank, world, local_rank)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        if out is not None:
            print(json.dumps(out), flush=True)
        return
    if args.probe_shared_board:
        run_board_probe(args)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
