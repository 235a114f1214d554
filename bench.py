#!/usr/bin/env python
"""Benchmark of the Lp-Box ADMM hot path (contract: one JSON line on rank 0).

Headline (`--config lp_plain`, the default; BASELINE.json configs[1]): a batch of synthetic combinatorial-auction
instances, j=100 items, k=500 bids, 10,000 instances PER GPU (weak scaling: instances are independent, no data-path
collective; `--scaling strong` splits 10,000 instances over the ranks instead).  A "step" is one pass of the hot path
over the batch: ADMM_lp_iters_init + the plain Lp-Box ADMM loop to convergence for every instance, PARITY mode.

  value    : instances/s, inputs resident in HBM when the timed region starts (device time from CUDA events recorded on
             the library's stream, max over ranks)
  e2e      : the same metric through the public API (`lpbox.LPBatch` over the C ABI) with HOST buffers: H2D of the
             problem, init, solve, D2H of the log rows and packed binary solutions inside the timed region (+ the final
             NCCL gather of the packed solutions when N > 1)
  parity   : the GPU batch's relaxed iterates / binary solutions / objectives compared IN THIS RUN with the reference arm's
             results on the instances the `cpu_baseline` leg solves (a mismatch fails the run)
  --impl reference : the reference's own compiled Eigen solver (oracle/_ref, else the oracle port) on the host cores.

Other configurations (`--config`): lp_fast (opt-in fast mode, NOT bit-identical, with its accuracy vs parity mode),
lp_l2f (configs[1] with MHA early fixing), lp_large (configs[4] shape, j=400 k=2000), seg (configs[2]), sa (configs[3]),
policy (the early-fixing network alone).  Each prints the same kind of line with its own roofline / cpu_baseline / e2e.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "accelerated-lpbox-admm_b200")

import numpy as np  # noqa: E402

N_ITEMS, N_BIDS = 100, 500
MAX_ITERS = 20000           # LP.cpp:498
UNIT = "instances/s"


def alg_bytes(n, m, nnz, admm_iters, cg_iters):
    """SURVEY.md §8d streaming model (fp64 values + int32 indices, each operand once per logical operation):
    per CG iteration 24 nnz + 80 n + 16 m, per ADMM iteration outside CG 48 nnz + 112 n + 64 m."""
    return admm_iters * (48.0 * nnz + 112.0 * n + 64.0 * m) + cg_iters * (24.0 * nnz + 80.0 * n + 16.0 * m)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=float(j["hbm_gbs"]), tf=float(j.get("bf16_tflops_sustained", 1400.0)), src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf=1400.0, src="fallback (B200_PROFILING.md)")


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


# ---- workload generators that do NOT touch the product library (used by the reference arm) ------------------------------------
def gen_auctions_standalone(seed, count, n_items, n_bids, add_item_prob=0.7):
    """The benchmark's auction generator (csrc/auction_gen.cpp) through its own library oracle/liblpbox_gen.so: the same
    instances `lpbox.gen_auctions` produces, without mapping liblpbox_b200.so."""
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liblpbox_gen.so"))
    ip, dp = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)
    m_p, cp_p, ri_p, pr_p = ip(), ip(), ip(), dp()
    lib.lpbox_gen_auctions.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                       ctypes.POINTER(ip), ctypes.POINTER(ip), ctypes.POINTER(ip), ctypes.POINTER(dp)]
    rc = lib.lpbox_gen_auctions(int(seed), int(count), int(n_items), int(n_bids), float(add_item_prob), 0, ctypes.byref(m_p),
                                ctypes.byref(cp_p), ctypes.byref(ri_p), ctypes.byref(pr_p))
    if rc != 0:
        raise RuntimeError("lpbox_gen_auctions failed")
    ms = np.ctypeslib.as_array(m_p, shape=(count,)).copy()
    cps = np.ctypeslib.as_array(cp_p, shape=(count, n_bids + 1)).copy()
    tot = int(cps[:, -1].sum())
    ris = np.ctypeslib.as_array(ri_p, shape=(max(tot, 1),))[:tot].copy()
    prs = np.ctypeslib.as_array(pr_p, shape=(count, n_bids)).copy()
    libc = ctypes.CDLL(None)
    for p in (m_p, cp_p, ri_p, pr_p):
        libc.free(ctypes.cast(p, ctypes.c_void_p))
    out, o = [], 0
    for i in range(count):
        nz = int(cps[i, -1])
        out.append((int(ms[i]), n_bids, cps[i], ris[o:o + nz], None, -prs[i], None))
        o += nz
    return out


def synth_image(seed, nr, nc, blobs=5):
    """Soft blobs of intensity ~0.2*263 on a ~0.6*263 background + N(0, 8) noise."""
    rng = np.random.default_rng(seed)
    img = np.full((nr, nc), 0.6 * 263)
    yy, xx = np.mgrid[0:nr, 0:nc]
    for _ in range(blobs):
        cy, cx, r = rng.uniform(0.15 * nr, 0.85 * nr), rng.uniform(0.15 * nc, 0.85 * nc), rng.uniform(0.08, 0.2) * min(nr, nc)
        img[(yy - cy) ** 2 + (xx - cx) ** 2 < r * r] = 0.2 * 263
    return np.clip(img + rng.normal(0, 8, img.shape), 0, 255).astype(np.uint8)


# ---- CPU side (reference arm / cpu_baseline) -------------------------------------------------------------------------
def cpu_kind():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_harness as rh
        if rh.available():
            return "reference"
    except Exception:
        pass
    return "port"


def _cpu_solve_lp(args):
    """One auction instance, plain Lp-Box ADMM to convergence: the reference binary (kind "reference") or the C port."""
    kind, (m, n, colptr, rowidx, _v, b, _f) = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import scipy.sparse as sp
    t = time.perf_counter()
    if kind == "reference":
        import ref_harness as rh
        E = sp.csc_matrix((np.ones(len(rowidx)), rowidx, colptr), shape=(m, n)).tocsr(); E.sort_indices()
        res = rh.admm_linear_ineq((m, n, E.indptr, E.indices, E.data), b, np.ones(m), np.ones(n), rh.Hyper.lp())
        x = res["x"]
    else:
        import oracle as orc
        o = orc.OracleLP(); o.set_problem_csc(m, n, colptr, rowidx, np.ones(len(rowidx)), b, np.ones(m)); o.solve_init(); o.solve_iter(0, MAX_ITERS)
        x = o.state()["x"]
    return time.perf_counter() - t, x


def _cpu_solve_seg(args):
    """One image: graph construction + ADMM_bqp_unconstrained of the reference binary (or the C port)."""
    kind, img = args
    sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from seg_util import OracleSeg
    t = time.perf_counter()
    o = OracleSeg()
    rp, ci, va, b, c = o.build_graph(img)
    if kind == "reference":
        import ref_harness as rh
        res = rh.admm_unconstrained((rp, ci, va), b, np.zeros(len(b)), rh.Hyper.seg())
        x = res["x"]
    else:
        o.set_problem(rp, ci, va, b, c); o.init(); o.legacy()
        x = o.state()["x"]
    return time.perf_counter() - t, x


def cpu_pool(fn, items, procs):
    import multiprocessing as mp
    kind = cpu_kind()
    if kind == "port":
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as orc
        orc.lib()   # make sure the .so is built before forking
    ctx = mp.get_context("fork")
    t = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(fn, [(kind, p) for p in items], chunksize=1)
    wall = time.perf_counter() - t
    return kind, wall, res


def host_procs():
    return max(1, min(os.cpu_count() or 1, 64))


def reference_arm(args, rank):
    """The reference's own CPU implementation of the path on the host cores (rank 0 only).  Does not import the product."""
    if rank != 0:
        return
    procs = host_procs()
    cfg = args.config
    if cfg == "seg":
        sample = max(4, min(procs, 16))
        items = [synth_image(s, 375, 500) for s in range(sample)]
        fn, metric, unit, what = _cpu_solve_seg, "seg_images_per_sec", "images/s", f"{sample} synthetic 375x500 images (graph construction + ADMM_bqp_unconstrained) per step"
    elif cfg == "lp_large":
        sample = max(4, min(procs, 16))
        items = gen_auctions_standalone(args.seed + 11, sample, 400, 2000)
        fn, metric, unit, what = _cpu_solve_lp, "admm_instances_per_sec", UNIT, f"{sample} auctions j=400 k=2000 per step"
    elif cfg in ("lp_plain", "lp_fast", "lp_l2f"):
        sample = max(8, procs)              # one instance per worker per step (>= 8)
        items = gen_auctions_standalone(args.seed, sample, N_ITEMS, N_BIDS)
        fn, metric, unit, what = _cpu_solve_lp, "admm_instances_per_sec", UNIT, f"{sample} auctions j={N_ITEMS} k={N_BIDS} per step"
    else:
        _emit({"impl": "reference", "unavailable": f"--config {cfg}: the reference path is PyTorch on the GPU (see the `cpu_baseline` object of the ours line)"})
        return
    times, kind = [], cpu_kind()
    for s in range(args.warmup + args.steps):
        kind, wall, _ = cpu_pool(fn, items, procs)
        if s >= args.warmup:
            times.append(wall)
    ms = 1e3 * sum(times) / len(times)
    val = sample / (ms / 1e3)
    line = {"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{cfg}: plain Lp-Box ADMM on the host cores, {what}, solved to convergence by {procs} host processes "
                                   "(instances from the benchmark's generator via oracle/liblpbox_gen.so; the product library is not loaded)"},
            "cpu_baseline": {"value": val, "unit": unit, "cores": procs, "kind": kind,
                             "sample": f"{what}, one process per core, logging off"},
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# ---- GPU arm: shared plumbing -------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args):
        import torch
        self.torch = torch
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.peaks = measured_peaks()

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, maxes, sums):
        torch = self.torch
        t = torch.tensor(list(maxes), dtype=torch.float64, device="cuda")
        c = torch.tensor(list(sums), dtype=torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(c, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()], [float(v) for v in c.tolist()]

    def gather(self, vals):
        torch = self.torch
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        if self.dist is None:
            return [t.tolist()]
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [o.tolist() for o in out]

    def finish(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def base_line(ctx, metric, unit, value, ms_per_step, dtype, workload, extra_cfg=None):
    a = ctx.args
    cfg = {"workload": workload}
    cfg.update(extra_cfg or {})
    return {"metric": metric, "value": value, "unit": unit, "n_gpus": ctx.world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": dtype,
            "data": "synthetic", "config": cfg}


def onchip_roofline():
    """The binding on-chip resources of the window kernel, from the committed ncu capture of the current kernel (static)."""
    p = os.path.join(ROOT, "profiles", "r02_window_kernel_onchip.json")
    try:
        j = json.load(open(p))
        j["source"] = "static: profiles/r02_window_kernel_onchip.json (ncu --set full of tools/quick_bench.py 1036 60)"
        return j
    except (OSError, ValueError):
        return None


# ---- configs[1] plain / fast -------------------------------------------------------------------------------------------
def run_lp(ctx, fast=False, large=False):
    import lpbox
    a, torch = ctx.args, ctx.torch
    n_items, n_bids = (400, 2000) if large else (N_ITEMS, N_BIDS)
    default_B = 1184 if large else 10000
    B_total = a.batch if a.batch > 0 else default_B
    if a.scaling == "strong":
        lo = (B_total * ctx.rank) // ctx.world; hi = (B_total * (ctx.rank + 1)) // ctx.world
        B = hi - lo
        probs = lpbox.gen_auctions(a.seed + (11 if large else 0), B_total, n_items, n_bids)[lo:hi]
    else:
        B = B_total
        probs = lpbox.gen_auctions(a.seed + (11 if large else 0) + 1000003 * ctx.rank, B, n_items, n_bids)
    batch = lpbox.LPBatch(probs, device=ctx.local, hist_cap=0)
    if fast:
        batch.set_mode("fast")
    state_bytes = sum(8 * (8 * p[1] + 3 * p[0]) + 4 * len(p[3]) + 2 * (p[0] + p[1]) for p in probs)

    def step():
        batch.init()
        ms = batch.last_kernel_ms()
        log = batch.solve(MAX_ITERS)
        return ms + batch.last_kernel_ms(), log

    for _ in range(a.warmup):
        step()
    sampler = ClockSampler(ctx.local)
    ctx.barrier()
    sampler.start()
    l0 = batch.launch_count()
    t0 = time.perf_counter()
    dev_ms, logs = 0.0, []
    for _ in range(a.steps):
        ms, log = step()
        dev_ms += ms
        logs.append(log)
    ctx.barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = batch.launch_count() - l0
    clocks = sampler.stop()
    log = logs[-1]
    admm_it, cg_it = int(log["iters"].sum()), int(log["cg_iters"].sum())
    abytes = sum(alg_bytes(p[1], p[0], len(p[3]), int(r["iters"]), int(r["cg_iters"])) for p, r in zip(probs, log))
    kern_ms = batch.last_kernel_ms()        # the window kernel's own duration (CUDA events around it on its stream), last step

    # ---- e2e: host buffers -> public API -> host results -----------------------------------------------------------
    e2e_ms, h2d, d2h = 0.0, 0, 0
    for _ in range(a.e2e_steps):
        ctx.barrier()
        t1 = time.perf_counter()
        bb = lpbox.LPBatch(probs, device=ctx.local, hist_cap=0)
        if fast:
            bb.set_mode("fast")
        bb.init()
        bb.solve(MAX_ITERS, want_log=False)       # the log rows and packed solutions are read back once, by results()
        elog, bits = bb.results()
        if ctx.dist is not None:   # the final gather of packed solutions over NVLink
            tb = torch.from_numpy(bits).cuda()
            rows = torch.tensor([tb.shape[0]], dtype=torch.int64, device="cuda")
            ctx.dist.all_reduce(rows, op=ctx.dist.ReduceOp.MAX)           # shards differ by at most one instance under strong scaling
            pad = int(rows.item()) - tb.shape[0]
            if pad:
                tb = torch.cat([tb, torch.zeros((pad,) + tuple(tb.shape[1:]), dtype=tb.dtype, device=tb.device)])
            out = torch.empty((ctx.world * tb.shape[0],) + tuple(tb.shape[1:]), dtype=tb.dtype, device=tb.device)
            ctx.dist.all_gather_into_tensor(out, tb)
            _ = out.cpu()
        ctx.barrier()
        e2e_ms += 1e3 * (time.perf_counter() - t1)
        h2d, d2h = bb.h2d_bytes(), bb.d2h_bytes()
        bb.close()
    e2e_ms /= max(a.e2e_steps, 1)

    # ---- accuracy of the fast mode vs parity mode on the same instances (>= 1000 instances) ----------------------------
    fast_acc = None
    if fast and ctx.rank == 0:
        nb = min(B, 2000)
        pb = lpbox.LPBatch(probs[:nb], device=ctx.local, hist_cap=0); pb.init(); plog = pb.solve(MAX_ITERS); _, pbits = pb.results(); pb.close()
        fb = lpbox.LPBatch(probs[:nb], device=ctx.local, hist_cap=0); fb.set_mode("fast"); fb.init(); flog = fb.solve(MAX_ITERS); _, fbits = fb.results(); fb.close()
        flips = np.unpackbits(pbits ^ fbits, axis=1).sum(axis=1)
        gap = (flog["obj"] - plog["obj"]) / np.abs(plog["obj"])
        fast_acc = {"instances": nb, "identical_binary_solutions": int((flips == 0).sum()), "flipped_bits_mean": float(flips.mean()),
                    "flipped_bits_max": int(flips.max()), "objective_gap_mean": float(gap.mean()), "objective_gap_abs_max": float(np.abs(gap).max()),
                    "objective_gap_quantiles_1_50_99": [float(q) for q in np.quantile(gap, [0.01, 0.5, 0.99])],
                    "iters_ratio_mean": float((flog["iters"] / np.maximum(plog["iters"], 1)).mean()),
                    "infeasible_parity": int((plog["infeasible"] > 0).sum()), "infeasible_fast": int((flog["infeasible"] > 0).sum())}

    per_rank = ctx.gather([kern_ms, float(admm_it), float(B)])
    (dev_ms, wall_ms, e2e_ms, kern_ms_max), (tot_B, tot_admm, tot_cg, tot_bytes, tot_launch) = ctx.reduce(
        [dev_ms, wall_ms, e2e_ms, kern_ms], [float(B), float(admm_it), float(cg_it), float(abytes), float(launches)])

    line = None
    if ctx.rank == 0:
        ms_per_step = dev_ms / a.steps
        value = tot_B / (ms_per_step / 1e3)
        name = "lp_fast" if fast else ("lp_large" if large else "lp_plain")
        mode = ("FAST mode (tree reductions + FMA; NOT bit-identical to the reference -- see `fast_accuracy`)" if fast else
                "parity mode (bit-identical to the reference)")
        wl = (f"{name} / configs[{4 if large else 1}]: {int(tot_B)} synthetic combinatorial-auction instances (j={n_items}, k={n_bids}) over "
              f"{ctx.world} GPU(s), plain Lp-Box ADMM to convergence (ADMM_lp_iters_init + ADMM_lp_iters(0,2e4)), {mode}; instances come "
              "from the native restatement of the reference generator (csrc/auction_gen.cpp, own RNG: same distribution, different "
              "individual instances; tests/test_host_logic_cpu.py compares the distributions)")
        line = base_line(ctx, "admm_instances_per_sec", UNIT, value, ms_per_step, "f64", wl,
                         {"instances_per_gpu": B, "n_items": n_items, "n_bids": n_bids,
                          "l2": f"inputs larger than L2 ({state_bytes / 1e6:.0f} MB of instance state per GPU vs 126 MB L2)" if state_bytes > 126e6 else
                                f"{state_bytes / 1e6:.0f} MB of instance state per GPU; the window kernel keeps the iteration on chip, no reuse across steps is possible (every step restarts from x = 1)",
                          "parallelism": f"instances sharded over {ctx.world} GPU(s), no data-path collective"})
        ach = (abytes / 1e9) / (kern_ms / 1e3)
        traffic, tsrc = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_window_kernel_traffic.json")))
            if B == 10000 and not large and not fast:
                traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"]); tsrc = "static: profiles/r02_window_kernel_traffic.json (ncu capture of this launch configuration)"
        except (OSError, ValueError, KeyError):
            pass
        line.update({
            "admm_iters_per_sec": tot_admm / (ms_per_step / 1e3), "cg_iters_per_sec": tot_cg / (ms_per_step / 1e3),
            "wall_ms_per_step": wall_ms / a.steps,
            "e2e": ({"value": tot_B / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                     "steps": a.e2e_steps} if e2e_ms > 0 else None),
            "gpu_launches": int(tot_launch), "clocks": clocks,
            "per_rank": {"window_kernel_ms": [r[0] for r in per_rank], "admm_iters": [int(r[1]) for r in per_rank], "instances": [int(r[2]) for r in per_rank]},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": ctx.peaks["hbm"], "unit": "GB/s", "frac": ach / ctx.peaks["hbm"], "traffic": traffic,
                         "traffic_source": tsrc, "kernel": "lp_admm_window_kernel", "peak_source": ctx.peaks["src"],
                         "algorithmic_bytes_per_launch": int(abytes),
                         "note": "achieved = ALGORITHMIC bytes (SURVEY.md 8d streaming model) / kernel duration.  The kernel keeps the whole "
                                 "iteration on chip, so HBM does not bind it and this fraction is not a quality measure (it can exceed 1); "
                                 "the binding resources are in `onchip`",
                         "onchip": onchip_roofline()}})
        if fast_acc is not None:
            line["fast_accuracy"] = fast_acc
        if ctx.world == 1 and not a.no_cpu_baseline:
            procs = host_procs()
            sample = max(4, min(procs, 16)) if large else max(8, min(procs, 32))
            kind, wall, res = cpu_pool(_cpu_solve_lp, probs[:sample], procs)
            per = [r[0] for r in res]
            line["cpu_baseline"] = {"value": sample / wall, "unit": UNIT, "cores": procs, "kind": kind,
                                    "sample": f"first {sample} instances of the same batch, one process per core, logging off; "
                                              f"{sum(per) / len(per):.2f} s per instance per core"}
            # in-run parity: the GPU batch against what the reference arm just computed for the same instances
            ident_x = ident_b = 0
            gap = 0.0
            for i in range(sample):
                xr = res[i][1]
                xg = batch.state(i)["x"]
                ident_x += int(xr.shape == xg.shape and np.array_equal(xr, xg))
                br, bg = (xr >= 0.5), batch.x_sol(i) >= 0.5
                ident_b += int(np.array_equal(br, bg))
                b = np.asarray(probs[i][5])
                o_r, o_g = float(b @ br), float(b @ bg)
                gap = max(gap, abs(o_g - o_r) / max(abs(o_r), 1e-300))
            line["parity"] = {"checked": sample, "identical": ident_b, "identical_relaxed_iterates": ident_x, "max_rel_obj_gap": gap,
                              "against": f"cpu_baseline kind={kind} (same instances, same run)",
                              "bar": "fast mode: reported only" if fast else "bit-identical relaxed x and binary x"}
            if not fast and (ident_b != sample or ident_x != sample):
                line["parity"]["FAILED"] = True
    batch.close()
    return line


# ---- configs[1] with MHA early fixing ------------------------------------------------------------------------------------
def run_l2f(ctx):
    import lpbox
    a, torch = ctx.args, ctx.torch
    from lpbox.policy import load_policy
    from lpbox.policy_kernel import PolicyKernel
    B = a.batch if a.batch > 0 else 10000
    if a.scaling == "strong":
        lo = (B * ctx.rank) // ctx.world; hi = (B * (ctx.rank + 1)) // ctx.world
        probs = lpbox.gen_auctions(a.seed, B, N_ITEMS, N_BIDS)[lo:hi]
        B = hi - lo
    else:
        probs = lpbox.gen_auctions(a.seed + 1000003 * ctx.rank, B, N_ITEMS, N_BIDS)
    wpath = os.path.join(PKG, "lpbox", "weights", "lp_mha_policy.pt")
    net = load_policy(wpath, device=f"cuda:{ctx.local}")
    score = PolicyKernel(net, device=ctx.local, chunk_rows=131072)     # bf16 tcgen05 kernels (csrc/policy_kernels.cu)
    ref = lpbox.LPBatch(probs, device=ctx.local, hist_cap=0); ref.init(); plog = ref.solve(MAX_ITERS); ref.close()

    def step(guard=not a.l2f_no_guard, max_iter=a.l2f_max_iter):
        lb = lpbox.LPBatch(probs, device=ctx.local, hist_cap=100)
        if guard:
            lb.set_fix_guard(True)
        lb.init()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        llog, bits, st = lpbox.solve_l2f(lb, score, ws=100, max_iter=max_iter)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        h2d, d2h = lb.h2d_bytes(), lb.d2h_bytes()
        lb.close()
        return ms, llog, st, h2d, d2h

    for _ in range(max(1, min(a.warmup, 2))):
        step()
    sampler = ClockSampler(ctx.local)
    ctx.barrier(); sampler.start()
    l0 = score.launch_count()
    dev_ms = 0.0
    for _ in range(a.steps):
        ms, llog, st, h2d, d2h = step()
        dev_ms += ms
    ctx.barrier()
    clocks = sampler.stop()
    e2e_ms = 0.0
    for _ in range(a.e2e_steps):
        ctx.barrier(); t1 = time.perf_counter()
        step()
        ctx.barrier(); e2e_ms += 1e3 * (time.perf_counter() - t1)
    e2e_ms /= max(a.e2e_steps, 1)
    gap = (llog["obj"] - plog["obj"]) / np.abs(plog["obj"])
    # the reference driver's own settings for comparison (LP.trainer:510-545: 1e4 iterations, deter_fix_2 without a guard)
    rms, rlog, rst, _, _ = step(guard=False, max_iter=10000)
    rgap = (rlog["obj"] - plog["obj"]) / np.abs(plog["obj"])
    ref_settings = {"instances_per_sec_this_gpu": B / (rms / 1e3), "objective_gap_mean_vs_plain": float(rgap.mean()),
                    "infeasible_instances_this_gpu": int((rlog["infeasible"] > 0).sum()), "unconverged_instances_this_gpu": int((rlog["status"] == 0).sum()),
                    "note": "max_iter 1e4 as in LP.trainer:510, no feasibility guard: instances still running after 100 windows are rounded as they are"}
    (dev_ms, e2e_ms), (tot_B, infeas, gapsum, rows) = ctx.reduce([dev_ms, e2e_ms], [float(B), float((llog["infeasible"] > 0).sum()), float(gap.sum()), float(st["policy_rows"])])
    line = None
    if ctx.rank == 0:
        ms_per_step = dev_ms / a.steps
        line = base_line(ctx, "admm_instances_per_sec", UNIT, tot_B / (ms_per_step / 1e3), ms_per_step, "f64 (ADMM) + bf16 (policy)",
                         f"lp_l2f / configs[1]: {int(tot_B)} synthetic auctions (j={N_ITEMS}, k={N_BIDS}) over {ctx.world} GPU(s) with MHA early fixing: "
                         "windows of 100 iterations, GraphAttentionEncoder policy on the bf16 tcgen05 kernels, deter_fix_2 thresholds 0.9/0.1, "
                         "device-resident window -> policy -> compaction loop behind lpbox_batch_solve_l2f (approximate solutions: see `quality`)"
                         + f"; window loop capped at {a.l2f_max_iter} iterations" + ("" if a.l2f_no_guard else "; feasibility guard on fix-to-one decisions ON (extension, not in the reference)"),
                         {"instances_per_gpu": B, "parallelism": f"instances sharded over {ctx.world} GPU(s), no data-path collective"})
        flops = rows * 17.56e6
        line.update({"e2e": {"value": tot_B / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": a.e2e_steps,
                             "note": "wall clock incl. batch creation from host arrays and the read-back of the results"},
                     "gpu_launches": int(st["windows"] * 4 + (score.launch_count() - l0)), "clocks": clocks,
                     "quality": {"objective_gap_mean_vs_plain": gapsum / tot_B, "infeasible_instances": int(infeas), "windows": st["windows"],
                                 "mean_admm_iters": float(llog["iters"].mean()), "mean_admm_iters_plain": float(plog["iters"].mean()),
                                 "max_iter": a.l2f_max_iter, "fix_guard": not a.l2f_no_guard, "with_reference_driver_settings": ref_settings},
                     "roofline": {"bound": "tensor", "achieved": flops / 1e12 / (ms_per_step / 1e3) / ctx.world, "peak": ctx.peaks["tf"], "unit": "TFLOP/s",
                                  "frac": flops / 1e12 / (ms_per_step / 1e3) / ctx.world / ctx.peaks["tf"], "traffic": None,
                                  "note": "policy flops (17.56 MFLOP per variable-window) over the WHOLE step (ADMM windows included), per GPU; "
                                          "the policy kernels alone are measured by --config policy"}})
    return line


# ---- configs[2]: segmentation --------------------------------------------------------------------------------------------
def run_seg(ctx):
    import lpbox
    a, torch = ctx.args, ctx.torch
    B_total = a.batch if a.batch > 0 else 1024
    nr, nc = 375, 500
    if a.scaling == "strong":
        lo = (B_total * ctx.rank) // ctx.world; hi = (B_total * (ctx.rank + 1)) // ctx.world
        seeds = list(range(lo, hi))
    else:
        seeds = [1000003 * ctx.rank + s for s in range(B_total)]
    B = len(seeds)
    distinct = min(B, 64)       # 64 distinct synthetic images per rank (generation on the host is the slow part), cycled
    base = [synth_image(seeds[s], nr, nc) for s in range(distinct)]
    imgs = [base[i % distinct] for i in range(B)]
    n = nr * nc
    batch = lpbox.SegBatch(imgs, device=ctx.local)

    def step():
        batch.init()
        e = batch.solve()
        return batch.last_kernel_ms(), e

    for _ in range(a.warmup):
        step()
    sampler = ClockSampler(ctx.local)
    ctx.barrier(); sampler.start()
    l0 = batch.launch_count()
    dev_ms = 0.0
    for _ in range(a.steps):
        ms, energy = step()
        dev_ms += ms
    ctx.barrier()
    launches = batch.launch_count() - l0
    clocks = sampler.stop()
    log = batch.results()
    nnz = 7 * n
    abytes = float((log["iters"].astype(float) * (24.0 * nnz + 112.0 * n) + log["cg_iters"].astype(float) * (12.0 * nnz + 80.0 * n)).sum())
    kern_ms = batch.last_kernel_ms()
    e2e_ms = 0.0
    h2d = d2h = 0
    for _ in range(a.e2e_steps):
        ctx.barrier(); t1 = time.perf_counter()
        bb = lpbox.SegBatch(imgs, device=ctx.local); bb.init(); bb.solve(); bb.results()
        xs = bb.x_sol(0)
        ctx.barrier(); e2e_ms += 1e3 * (time.perf_counter() - t1)
        h2d, d2h = bb.h2d_bytes(), bb.d2h_bytes()
        bb.close()
    e2e_ms /= max(a.e2e_steps, 1)
    (dev_ms, e2e_ms, kern_ms), (tot_B, tot_it, tot_cg, tot_launch) = ctx.reduce(
        [dev_ms, e2e_ms, kern_ms], [float(B), float(log["iters"].sum()), float(log["cg_iters"].sum()), float(launches)])
    line = None
    if ctx.rank == 0:
        ms_per_step = dev_ms / a.steps
        line = base_line(ctx, "seg_images_per_sec", "images/s", tot_B / (ms_per_step / 1e3), ms_per_step, "f64",
                         f"seg / configs[2]: {int(tot_B)} synthetic {nr}x{nc} grey images over {ctx.world} GPU(s) ({distinct} distinct images per GPU, cycled), "
                         "device graph construction (6-neighbour Laplacian as the reference builds it) + ADMM_bqp_unconstrained_legacy to "
                         "convergence, parity mode", {"images_per_gpu": B, "n": n,
                                                      "l2": f"inputs larger than L2 ({B * 11 * n * 8 / 1e9:.1f} GB of iterate vectors per GPU vs 126 MB L2)"})
        ach = (abytes / 1e9) / (kern_ms / 1e3)
        line.update({"admm_iters_per_sec": tot_it / (ms_per_step / 1e3), "cg_iters_per_sec": tot_cg / (ms_per_step / 1e3),
                     "e2e": {"value": tot_B / (e2e_ms / 1e3), "unit": "images/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": a.e2e_steps},
                     "gpu_launches": int(tot_launch), "clocks": clocks,
                     "roofline": {"bound": "hbm", "achieved": ach, "peak": ctx.peaks["hbm"], "unit": "GB/s", "frac": ach / ctx.peaks["hbm"], "traffic": None,
                                  "traffic_source": None, "kernel": "seg_admm_kernel", "peak_source": ctx.peaks["src"],
                                  "algorithmic_bytes_per_launch": int(abytes),
                                  "note": "algorithmic bytes = SURVEY.md 8d streaming model (12 B per stored entry of A, every vector access counted); the kernel "
                                          "stores A in 3 B per entry and fuses passes, so its physical DRAM traffic (`traffic`) is below the model"}})
        tp = os.path.join(ROOT, "profiles", "r02_seg_kernel_traffic.json")
        if os.path.exists(tp) and int(tot_B) == 1024 * ctx.world and nr * nc == 187500:
            tj = json.load(open(tp))
            line["roofline"]["traffic"] = int(tj["dram_bytes_read"] + tj["dram_bytes_write"])
            line["roofline"]["traffic_source"] = "static: profiles/r02_seg_kernel_traffic.json (ncu metrics pass over this launch configuration)"
            line["roofline"]["physical_frac_in_capture"] = tj["dram_GBps_in_capture"] / ctx.peaks["hbm"]
        if ctx.world == 1 and not a.no_cpu_baseline:
            procs = host_procs()
            sample = max(4, min(procs, 16, distinct))
            kind, wall, res = cpu_pool(_cpu_solve_seg, base[:sample], procs)
            line["cpu_baseline"] = {"value": sample / wall, "unit": "images/s", "cores": procs, "kind": kind,
                                    "sample": f"first {sample} images of the batch, one process per core (graph construction + solve); "
                                              f"{sum(r[0] for r in res) / len(res):.1f} s per image per core"}
            ident = sum(int(np.array_equal(res[i][1], batch.state(i)["x"])) for i in range(sample))
            line["parity"] = {"checked": sample, "identical": ident, "against": f"cpu_baseline kind={kind} (same images, same run)", "bar": "bit-identical relaxed x"}
            if ident != sample:
                line["parity"]["FAILED"] = True
    batch.close()
    return line


# ---- configs[3]: sparse adversarial attack -------------------------------------------------------------------------------
def run_sa(ctx):
    a, torch = ctx.args, ctx.torch
    import torchvision
    from lpbox import sparse_attack as sa
    N = a.batch if a.batch > 0 else 4096
    K = a.sa_iters
    torch.manual_seed(1234 + ctx.rank)
    dev = torch.device("cuda", ctx.local)
    model = torchvision.models.resnet18(num_classes=10).to(dev).eval()
    for p in model.parameters():
        p.requires_grad_(False)
    seg = (torch.arange(32, device=dev).view(32, 1) // 4 * 8 + torch.arange(32, device=dev).view(1, 32) // 4)
    seg = seg.unsqueeze(0).expand(3, 32, 32).reshape(-1).to(torch.int32).contiguous()
    h_images = torch.rand(N, 3, 32, 32).pin_memory()
    images = h_images.to(dev)
    with torch.no_grad():
        target = (model(images - 0.5).argmax(1) + 1) % 10
    eps = 0.1 * torch.randn(N, 3, 32, 32, device=dev)
    G0 = torch.ones(N, 3, 32, 32, device=dev); nw = torch.ones_like(G0)

    def step(img):
        return sa.update_G(model, img, target, eps, G0.clone(), sa.init_params(), seg, nw, args={"maxIter_g": K})

    for _ in range(a.warmup):
        step(images)
    sampler = ClockSampler(ctx.local)
    ctx.barrier(); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        G, res = step(images)
    e1.record(); ctx.barrier()
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    e2e_ms = 0.0
    for _ in range(a.e2e_steps):
        ctx.barrier(); t1 = time.perf_counter()
        img = h_images.to(dev, non_blocking=True)
        G, res = step(img)
        l0 = (G >= 0.5).sum(dim=(1, 2, 3)).cpu()
        ctx.barrier(); e2e_ms += 1e3 * (time.perf_counter() - t1)
    e2e_ms /= max(a.e2e_steps, 1)
    (dev_ms, e2e_ms), (tot,) = ctx.reduce([dev_ms, e2e_ms], [float(N * K)])
    line = None
    if ctx.rank == 0:
        ms_per_step = dev_ms / a.steps
        line = base_line(ctx, "sa_image_iterations_per_sec", "image-iterations/s", tot / (ms_per_step / 1e3), ms_per_step, "f32",
                         f"sa / configs[3]: {N} synthetic 3x32x32 images per GPU against a random-init torchvision ResNet-18 (10 classes, eval, fp32), "
                         f"8x8 grid of 4x4 segments, {K} Lp-Box ADMM iterations of update_G per step (the reference runs 2000 per image; iterations/s "
                         "is size-independent), fused pre/post kernels around the PyTorch classifier", {"images_per_gpu": N, "iterations_per_step": K})
        line.update({"e2e": {"value": tot / (e2e_ms / 1e3), "unit": "image-iterations/s", "h2d_bytes_per_step": int(h_images.numel() * 4), "d2h_bytes_per_step": int(N * 8), "steps": a.e2e_steps},
                     "gpu_launches": int(2 * K * a.steps), "clocks": clocks,
                     "roofline": {"bound": "tensor", "achieved": None, "peak": ctx.peaks["tf"], "unit": "TFLOP/s", "frac": None, "traffic": None,
                                  "note": "the step is dominated by the PyTorch classifier forward/backward (library code, as in the reference); the two fused "
                                          "ADMM kernels (csrc/sa_kernels.cu) are elementwise, < 10 % of the step"}})
        if ctx.world == 1 and not a.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import sa_oracle
            Bm = torch.zeros(64, 3, 32, 32)
            sg = seg.view(3, 32, 32)[0].cpu()
            for s in range(64):
                Bm[s, :, sg == s] = 1
            mean = torch.full((1, 3, 1, 1), 0.5); std = torch.ones((1, 3, 1, 1))
            cm = torchvision.models.resnet18(num_classes=10).eval(); cm.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
            Kr = 20
            t1 = time.perf_counter()
            sa_oracle.update_G(cm, images[:1].cpu(), target[:1].cpu(), eps[:1].cpu(), G0[:1].cpu().clone(), sa_oracle.INIT, Bm, nw[:1].cpu(), Kr, mean=mean, std=std)
            dt = time.perf_counter() - t1
            line["cpu_baseline"] = {"value": Kr / dt, "unit": "image-iterations/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"the reference formulation (oracle/sa_oracle.py == main_ori.py:626-743: batch 1, PyTorch ops) on the host, {Kr} iterations of one image"}
    return line


# ---- the early-fixing policy network alone ---------------------------------------------------------------------------------
def run_policy(ctx):
    a, torch = ctx.args, ctx.torch
    from lpbox.policy import GraphAttentionEncoder
    from lpbox.policy_kernel import PolicyKernel
    rows = a.batch if a.batch > 0 else 500000
    torch.manual_seed(0)
    dev = torch.device("cuda", ctx.local)
    net = GraphAttentionEncoder(tokens=20).to(dev).eval()
    pk = PolicyKernel(net, device=ctx.local, chunk_rows=131072)   # 504 TFLOP/s; 462 at 32768, 427 at 16384 (tools/quick_bench_policy_chunks.py)
    h_x = torch.rand(rows, 20, 5).pin_memory()
    x = h_x.to(dev)
    for _ in range(a.warmup):
        pk(x)
    sampler = ClockSampler(ctx.local)
    ctx.barrier(); sampler.start()
    l0 = pk.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = pk(x)
    e1.record(); ctx.barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = pk.launch_count() - l0
    clocks = sampler.stop()
    e2e_ms = 0.0
    for _ in range(a.e2e_steps):
        ctx.barrier(); t1 = time.perf_counter()
        out = pk(h_x.to(dev, non_blocking=True)); _ = out.cpu()
        ctx.barrier(); e2e_ms += 1e3 * (time.perf_counter() - t1)
    e2e_ms /= max(a.e2e_steps, 1)
    macs = 20 * 2 * (128 * 384 + 128 * 128 + 128 * 512 * 2) + 2560 * 256 + 256 * 128 + 128 * 16 + 16 + 20 * 10 * 128 + 2 * 2 * 8 * 20 * 20 * 16
    (dev_ms, e2e_ms), (tot, tot_launch) = ctx.reduce([dev_ms, e2e_ms], [float(rows), float(launches)])
    line = None
    if ctx.rank == 0:
        ms_per_step = dev_ms / a.steps
        tf = 2.0 * macs * rows / (ms_per_step / 1e3) / 1e12
        line = base_line(ctx, "policy_variable_windows_per_sec", "variable-windows/s", tot / (ms_per_step / 1e3), ms_per_step, "bf16",
                         f"policy: GraphAttentionEncoder (2 layers, 8 heads, T=20 tokens x 5 iterates per variable, random init) forward on {rows} "
                         "variable-windows per GPU, bf16 tcgen05 kernels (csrc/policy_kernels.cu)", {"rows_per_gpu": rows})
        line.update({"e2e": {"value": tot / (e2e_ms / 1e3), "unit": "variable-windows/s", "h2d_bytes_per_step": int(h_x.numel() * 4), "d2h_bytes_per_step": int(rows * 4), "steps": a.e2e_steps},
                     "gpu_launches": int(tot_launch), "clocks": clocks,
                     "roofline": {"bound": "tensor", "achieved": tf, "peak": ctx.peaks["tf"], "unit": "TFLOP/s", "frac": tf / ctx.peaks["tf"], "traffic": None,
                                  "peak_source": ctx.peaks["src"] + " bf16_tflops_sustained", "flops_per_row": 2 * macs}})
        if ctx.world == 1 and not a.no_cpu_baseline:
            cn = GraphAttentionEncoder(tokens=20).eval(); cn.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})
            rs = 4096
            with torch.no_grad():
                cn(h_x[:256])
                t1 = time.perf_counter(); cn(h_x[:rs]); dt = time.perf_counter() - t1
            line["cpu_baseline"] = {"value": rs / dt, "unit": "variable-windows/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"the same network in PyTorch fp32 on the host, {rs} variable-windows"}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="lp_plain", choices=["lp_plain", "lp_fast", "lp_l2f", "lp_large", "seg", "sa", "policy"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--batch", type=int, default=0, help="units per GPU per step (weak) or in total (strong); 0 = the configuration's default")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--sa-iters", type=int, default=50)
    ap.add_argument("--l2f-no-guard", action="store_true", help="lp_l2f: switch the feasibility guard on fix-to-one decisions off (the guard is an extension, not in the reference)")
    ap.add_argument("--l2f-max-iter", type=int, default=20000, help="lp_l2f: iteration cap of the window loop (the solver's own cap, LP.cpp:498; the reference trainer stops at 1e4)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        reference_arm(args, int(os.environ.get("RANK", "0")))
        return

    sys.path.insert(0, PKG)
    ctx = Ctx(args)
    fn = {"lp_plain": lambda: run_lp(ctx), "lp_fast": lambda: run_lp(ctx, fast=True), "lp_large": lambda: run_lp(ctx, large=True),
          "lp_l2f": lambda: run_l2f(ctx), "seg": lambda: run_seg(ctx), "sa": lambda: run_sa(ctx), "policy": lambda: run_policy(ctx)}[args.config]
    line = fn()
    failed = False
    if ctx.rank == 0 and line is not None:
        _emit(line)
        failed = bool(line.get("parity", {}).get("FAILED"))
    ctx.finish()
    if failed:
        raise SystemExit("parity check failed: the GPU results differ from the reference arm's on the same instances")


def _emit(line):
    """The ONE JSON line of the contract goes to the real stdout; everything else this process (or a library such as NCCL,
    which prints its version banner to stdout) writes to fd 1 has been redirected to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    main()
