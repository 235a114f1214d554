#!/usr/bin/env python
"""Benchmark of the Lp-Box ADMM hot path (contract: one JSON line on rank 0).

Workload (BASELINE.json configs[1]): a batch of synthetic combinatorial-auction instances, j=100 items, k=500 bids,
10,000 instances PER GPU (weak scaling: instances are independent, no data-path collective).  A "step" is one pass
of the hot path over the batch: ADMM_lp_iters_init + the plain Lp-Box ADMM loop to convergence for every instance.

  value : instances/s, inputs resident in HBM when the timed region starts (device time from CUDA events recorded on
          the library's stream, max over ranks)
  e2e   : the same metric through the public API (`lpbox.LPBatch` over the C ABI) with HOST buffers: H2D of the
          problem, init, solve, D2H of the log rows and packed binary solutions inside the timed region (+ the final
          NCCL gather of the packed solutions when N > 1)
  --impl reference : the reference's own compiled Eigen solver (oracle/_ref, else the oracle port) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))

import numpy as np  # noqa: E402

N_ITEMS, N_BIDS = 100, 500
MAX_ITERS = 20000           # LP.cpp:498
METRIC = "admm_instances_per_sec"
UNIT = "instances/s"


def alg_bytes(n, m, nnz, admm_iters, cg_iters):
    """SURVEY.md §8d streaming model (fp64 values + int32 indices, each operand once per logical operation):
    per CG iteration 24 nnz + 80 n + 16 m, per ADMM iteration outside CG 48 nnz + 112 n + 64 m."""
    return admm_iters * (48.0 * nnz + 112.0 * n + 64.0 * m) + cg_iters * (24.0 * nnz + 80.0 * n + 16.0 * m)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(len(r) >= 6 and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---- CPU side (reference arm / cpu_baseline) -------------------------------------------------------------------------
def _cpu_solve_one(args):
    kind, (m, n, colptr, rowidx, _v, b, _f) = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import scipy.sparse as sp
    t = time.perf_counter()
    if kind == "reference":
        import ref_harness as rh
        E = sp.csc_matrix((np.ones(len(rowidx)), rowidx, colptr), shape=(m, n)).tocsr(); E.sort_indices()
        rh.admm_linear_ineq((m, n, E.indptr, E.indices, E.data), b, np.ones(m), np.ones(n), rh.Hyper.lp())
    else:
        import oracle as orc
        o = orc.OracleLP(); o.set_problem_csc(m, n, colptr, rowidx, np.ones(len(rowidx)), b, np.ones(m)); o.solve_init(); o.solve_iter(0, MAX_ITERS)
    return time.perf_counter() - t


def cpu_kind():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_harness as rh
        if rh.available():
            return "reference"
    except Exception:
        pass
    return "port"


def cpu_run(problems, procs):
    import multiprocessing as mp
    kind = cpu_kind()
    if kind == "port":
        import oracle as orc
        orc.lib()   # make sure the .so is built before forking
    ctx = mp.get_context("fork")
    t = time.perf_counter()
    with ctx.Pool(procs) as pool:
        per = pool.map(_cpu_solve_one, [(kind, p) for p in problems], chunksize=1)
    wall = time.perf_counter() - t
    return kind, wall, per


def reference_arm(args, rank):
    if rank != 0:
        return
    import lpbox
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    sample = max(8, procs)              # one instance per worker per step (>= 8)
    probs = lpbox.gen_auctions(args.seed, sample, N_ITEMS, N_BIDS)
    kind = cpu_kind()
    times = []
    for s in range(args.warmup + args.steps):
        kind, wall, _ = cpu_run(probs, procs)
        if s >= args.warmup:
            times.append(wall)
    ms = 1e3 * sum(times) / len(times)
    val = sample / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"configs[1] plain Lp-Box ADMM, auctions j={N_ITEMS} k={N_BIDS}; each step = {sample} instances "
                                   f"solved to convergence by {procs} host processes"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": kind,
                             "sample": f"{sample} instances per step, one process per core, logging off"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# ---- GPU arm ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=10000, help="instances per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--l2f-batch", type=int, default=10000, help="instances of the batch also solved with learned early fixing (0 = skip)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import lpbox
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch
    probs = lpbox.gen_auctions(args.seed + 1000003 * rank, B, N_ITEMS, N_BIDS)
    batch = lpbox.LPBatch(probs, device=local, hist_cap=0)
    state_bytes = sum(8 * (8 * p[1] + 3 * p[0]) + 4 * len(p[3]) + 2 * (p[0] + p[1]) for p in probs)

    def step():
        batch.init()
        ms = batch.last_kernel_ms()
        log = batch.solve(MAX_ITERS)
        return ms + batch.last_kernel_ms(), log

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = batch.launch_count()
    t0 = time.perf_counter()
    dev_ms, logs = 0.0, []
    for _ in range(args.steps):
        ms, log = step()
        dev_ms += ms
        logs.append(log)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = batch.launch_count() - l0
    clocks = sampler.stop()

    log = logs[-1]
    admm_it, cg_it = int(log["iters"].sum()), int(log["cg_iters"].sum())
    abytes = sum(alg_bytes(p[1], p[0], len(p[3]), int(r["iters"]), int(r["cg_iters"])) for p, r in zip(probs, log))
    # the window kernel's own duration (CUDA events around it on its stream) of the last step
    kern_ms = batch.last_kernel_ms()

    # ---- e2e: host buffers -> public API -> host results -----------------------------------------------------------
    e2e_ms, h2d, d2h = 0.0, 0, 0
    for _ in range(args.e2e_steps):
        barrier()
        t1 = time.perf_counter()
        bb = lpbox.LPBatch(probs, device=local, hist_cap=0)
        bb.init()
        bb.solve(MAX_ITERS)
        elog, bits = bb.results()
        if dist is not None:   # the final gather of packed solutions over NVLink
            tb = torch.from_numpy(bits).cuda()
            out = torch.empty((world * tb.shape[0],) + tuple(tb.shape[1:]), dtype=tb.dtype, device=tb.device)
            dist.all_gather_into_tensor(out, tb)
            _ = out.cpu()
        barrier()
        e2e_ms += 1e3 * (time.perf_counter() - t1)
        h2d, d2h = bb.h2d_bytes(), bb.d2h_bytes()
        bb.close()
    e2e_ms /= max(args.e2e_steps, 1)

    # ---- early-fixing variant (configs[1] "with MHA early fixing"): device-resident window loop with the shipped policy ----
    l2f = None
    wpath = os.path.join(ROOT, "accelerated-lpbox-admm_b200", "lpbox", "weights", "lp_mha_policy.pt")
    if args.l2f_batch > 0 and os.path.exists(wpath) and rank == 0:
        from lpbox.policy import load_policy
        nb = min(args.l2f_batch, B)
        net = load_policy(wpath, device=f"cuda:{local}")

        from lpbox.policy_kernel import PolicyKernel
        score = PolicyKernel(net, device=local, chunk_rows=32768)      # bf16 tcgen05 kernels (csrc/policy_kernels.cu)
        lb = lpbox.LPBatch(probs[:nb], device=local, hist_cap=100)
        lb.init()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        llog, _, lstats = lpbox.solve_l2f(lb, score, ws=100, max_iter=10000)
        torch.cuda.synchronize(); l2f_s = time.perf_counter() - t2
        gap = (llog["obj"] - log["obj"][:nb]) / np.abs(log["obj"][:nb])
        l2f = {"instances": nb, "value": nb / l2f_s, "unit": UNIT, "policy": "GraphAttentionEncoder (reference recipe, 40 epochs) on the bf16 tcgen05 policy kernels", "policy_launches": int(score.launch_count()),
               "windows": lstats["windows"], "policy_rows": lstats["policy_rows"], "window_kernel_ms": lstats["window_ms"],
               "objective_gap_mean": float(gap.mean()), "infeasible_instances": int((llog["infeasible"] > 0).sum()),
               "mean_admm_iters": float(llog["iters"].mean()), "note": "wall clock incl. policy; not comparable with the reference arm (plain ADMM)"}
        lb.close()

    # ---- reduce over ranks -------------------------------------------------------------------------------------------
    t = torch.tensor([dev_ms, wall_ms, e2e_ms, kern_ms], dtype=torch.float64, device="cuda")
    c = torch.tensor([float(B), float(admm_it), float(cg_it), float(abytes), float(launches)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dev_ms, wall_ms, e2e_ms, kern_ms = [float(v) for v in t.tolist()]
    tot_B, tot_admm, tot_cg, tot_bytes, tot_launch = [float(v) for v in c.tolist()]

    if rank == 0:
        ms_per_step = dev_ms / args.steps
        value = tot_B / (ms_per_step / 1e3)
        peak, peak_src = measured_peak()
        # roofline of the dominant kernel (lp_admm_window_kernel) on ONE GPU: algorithmic bytes of one launch / its duration
        ach = (abytes / 1e9) / (kern_ms / 1e3)
        traffic = None                      # dram bytes per launch, from the committed ncu capture of the same launch (10,000 instances)
        try:
            with open(os.path.join(ROOT, "profiles", "r01_window_kernel_traffic.json")) as fh:
                tj = json.load(fh)
            if B == 10000:
                traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"])
        except (OSError, ValueError, KeyError):
            traffic = None
        alg_bytes_launch = int(abytes)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"configs[1]: {B} synthetic combinatorial-auction instances per GPU (j={N_ITEMS}, k={N_BIDS}), "
                                   "plain Lp-Box ADMM to convergence (ADMM_lp_iters_init + ADMM_lp_iters(0,2e4)), parity mode "
                                   "(bit-identical to the reference); the same batch with MHA early fixing is timed separately in the "
                                   "`l2f` object (approximate solutions, so it is not the headline)",
                       "instances_per_gpu": B, "n_items": N_ITEMS, "n_bids": N_BIDS,
                       "l2": f"inputs larger than L2 ({state_bytes / 1e6:.0f} MB of instance state per GPU vs 126 MB L2)",
                       "parallelism": f"instances sharded over {world} GPU(s), no data-path collective"},
            "admm_iters_per_sec": tot_admm / (ms_per_step / 1e3), "cg_iters_per_sec": tot_cg / (ms_per_step / 1e3),
            "wall_ms_per_step": wall_ms / args.steps,
            "e2e": ({"value": tot_B / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)} if e2e_ms > 0 else None),
            "gpu_launches": int(tot_launch),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                         "kernel": "lp_admm_window_kernel", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes_launch,
                         "note": "achieved = algorithmic bytes (SURVEY.md 8d streaming model) / kernel duration; the kernel keeps the "
                                 "iteration on chip, so physical DRAM traffic (`traffic`, bytes per launch from the ncu capture of this "
                                 "launch configuration, profiles/r01_window_kernel_traffic.json) is far below the algorithmic bytes"},
        }
        if l2f is not None:
            line["l2f"] = l2f
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            procs = max(1, min(cores, 64))
            sample = max(8, min(procs, 32))
            kind, wall, per = cpu_run(probs[:sample], procs)
            line["cpu_baseline"] = {"value": sample / wall, "unit": UNIT, "cores": procs, "kind": kind,
                                    "sample": f"first {sample} instances of the same batch, one process per core, logging off; "
                                              f"{sum(per) / len(per):.2f} s per instance per core"}
        _emit(line)
    batch.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def _emit(line):
    """The ONE JSON line of the contract goes to the real stdout; everything else this process (or a library such as NCCL,
    which prints its version banner to stdout) writes to fd 1 has been redirected to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    main()
