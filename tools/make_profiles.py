"""Turns the ncu captures / bench lines a GPU run left under gpurun_out/ into the tracked summaries under profiles/ (round 2)."""
import csv, json, os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def window_kernel():
    rep = os.path.join(G, "r02_lp_window.ncu-rep")
    if not os.path.exists(rep):
        return
    rows = ncu_csv(rep, "raw")
    m = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
    f = lambda k: float(m[k][1]) if k in m else float("nan")
    n_iter = 1036 * 60
    cycles = f("sm__cycles_elapsed.max")
    wf = f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
    want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
            "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum"]
    with open(os.path.join(P, "r02_window_kernel_ncu.txt"), "w") as fh:
        fh.write("# r02 lp_admm_window_kernel<128,4,1,0> (parity mode), FINAL kernel of round 2\n"
                 "# ncu --set full --import-source on --clock-control none -k regex:lp_admm_window -c 1 python tools/quick_bench.py 1036 60   (1036 instances x 60 ADMM iterations)\n")
        for k in want + sorted(h for h in m if "issue_stalled" in h and "per_issue_active" in h):
            if k in m:
                fh.write(f"{k:90s} {m[k][0]:16s} {m[k][1]}\n")
        # per-opcode view of the source page
        src = ncu_csv(rep, "source")
        hdr, data = src[1], src[2:]
        ix = {h: i for i, h in enumerate(hdr)}
        def g(r, k):
            try: return float(r[ix[k]])
            except Exception: return 0.0
        op = collections.Counter(); owf = collections.Counter(); oid = collections.Counter(); osm = collections.Counter()
        spills = 0
        for r in data:
            mm = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
            o = ".".join((mm.group(2) if mm else "?").split(".")[:2])
            op[o] += g(r, "Instructions Executed"); owf[o] += g(r, "L1 Wavefronts Shared"); oid[o] += g(r, "L1 Wavefronts Shared Ideal"); osm[o] += g(r, "# Samples")
            if "LDL" in r[ix["Source"]] or "STL" in r[ix["Source"]]:
                spills += g(r, "Instructions Executed")
        ts = sum(osm.values())
        fh.write(f"\n# per ADMM iteration (source page): {sum(op.values()) / n_iter:.0f} warp instructions, {sum(owf.values()) / n_iter:.0f} shared wavefronts, "
                 f"{spills / n_iter:.1f} local-memory (spill) instructions\n# opcode: warp-instructions / shared wavefronts (ideal) per ADMM iteration, share of stall samples\n")
        for o, c in op.most_common(18):
            fh.write(f"#   {o:20s} {c / n_iter:9.1f} {owf[o] / n_iter:9.1f} ({oid[o] / n_iter:9.1f}) {100 * osm[o] / ts:5.1f} %\n")
    onchip = {
        "kernel": "lp_admm_window_kernel<128,4,1,0>", "capture": "profiles/r02_window_kernel_ncu.txt (1036 instances x 60 iterations)",
        "smem_wavefronts_per_sm_cycle": wf / (cycles * 148), "lsu_data_pipe_pct_of_peak": f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "smem_conflict_frac": f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / wf,
        "fp64_pipe_active_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_slots_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": int(f("launch__registers_per_thread")), "ctas_per_sm": int(f("launch__occupancy_limit_registers")),
        "local_mem_instructions_per_admm_iteration": spills / n_iter,
        "smem_wavefronts_per_admm_iteration": wf / n_iter, "warp_instructions_per_admm_iteration": f("smsp__inst_executed.sum") / n_iter,
        "barrier_stall_per_issue": f("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    }
    json.dump(onchip, open(os.path.join(P, "r02_window_kernel_onchip.json"), "w"), indent=1)
    print("onchip:", onchip)


def traffic():
    p = os.path.join(G, "r02_window_traffic.csv")
    if not os.path.exists(p):
        return
    rows = [r for r in csv.reader(open(p)) if len(r) > 14]
    m = {r[12]: float(r[14]) for r in rows[1:]}
    out = {"kernel": "lp_admm_window_kernel<128,4,1,0>", "launch": f"tools/quick_bench.py 10000 20000 gen = the bench's own batch (10,000 generated instances to convergence, grid {rows[1][8]} x 128 threads, one launch = one step)",
           "command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,... --clock-control none -k regex:lp_admm_window -c 1 (tools/gpu_profiles.sh)",
           "dram_bytes_read": int(m["dram__bytes_read.sum"]), "dram_bytes_write": int(m["dram__bytes_write.sum"]), "lts_bytes": int(m["lts__t_bytes.sum"]),
           "smem_wavefronts": int(m["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]), "smem_bank_conflicts": int(m["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]),
           "local_load_bytes": int(m["l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum"]), "local_store_bytes": int(m["l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum"]),
           "warp_instructions": int(m["smsp__inst_executed.sum"]), "duration_ns": int(m["gpu__time_duration.sum"]), "sm_cycles": int(m["sm__cycles_elapsed.max"]),
           "fp64_pipe_active_pct": m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"], "source": "profiles/r02_window_kernel_traffic_b10000.csv"}
    out["smem_wavefronts_per_sm_cycle"] = out["smem_wavefronts"] / (out["sm_cycles"] * 148)
    out["smem_conflict_frac"] = out["smem_bank_conflicts"] / out["smem_wavefronts"]
    json.dump(out, open(os.path.join(P, "r02_window_kernel_traffic.json"), "w"), indent=1)
    open(os.path.join(P, "r02_window_kernel_traffic_b10000.csv"), "w").write(open(p).read())
    print("traffic:", out)


def launches():
    p = os.path.join(G, "r02_launches_bench.csv")
    if not os.path.exists(p):
        return
    rows = [r for r in csv.reader(open(p)) if len(r) > 14 and r[0].isdigit()]
    tot = collections.Counter(); cnt = collections.Counter()
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]); tot[name] += float(r[14]); cnt[name] += 1
    open(os.path.join(P, "r02_launches_bench_b2072.csv"), "w").write(open(p).read())
    s = sum(tot.values())
    with open(os.path.join(P, "r02_launches_bench_b2072_summary.txt"), "w") as fh:
        fh.write("# kernels launched by `python bench.py --batch 2072 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline` under\n"
                 "# ncu --metrics gpu__time_duration.sum --clock-control none (tools/gpu_profiles.sh): launches, total ns, share of device time\n")
        for k, v in tot.most_common():
            fh.write(f"{k:70s} {cnt[k]:4d} {v:16.0f} {100 * v / s:8.4f} %\n")
    print("launch list:", {k: round(100 * v / s, 4) for k, v in tot.most_common(4)})


def seg():
    rep = os.path.join(G, "r02_seg_kernel.ncu-rep")
    if not os.path.exists(rep):
        return
    rows = ncu_csv(rep, "raw")
    m = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
    f = lambda k: float(m[k][1]) if k in m else float("nan")
    unit = lambda k: {"Gbyte": 1e9, "Tbyte": 1e12, "Mbyte": 1e6, "byte": 1.0}.get(m[k][0], 1.0)
    dram = f("dram__bytes_read.sum") * unit("dram__bytes_read.sum") + f("dram__bytes_write.sum") * unit("dram__bytes_write.sum")
    dur = f("gpu__time_duration.sum") * {"s": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9}.get(m["gpu__time_duration.sum"][0], 1.0)
    with open(os.path.join(P, "r02_seg_kernel_ncu.txt"), "w") as fh:
        fh.write("# r02 seg_admm_kernel<compact, 256 threads, row image> -- FINAL kernel of round 2 (5 CTAs/SM, 8-entry row image, fused passes, L2 prefetch)\n"
                 "# ncu --set full --import-source on --clock-control none -k regex:seg_admm -c 1 python tools/quick_bench_seg.py 740 375 500 30   (740 images 375x500 = one full wave, 30 ADMM iterations)\n"
                 f"# physical DRAM traffic of this launch = read + write below = {dram / 1e12:.3f} TB in {dur:.4f} s -> {dram / dur / 1e12:.2f} TB/s = {100 * dram / dur / 6534.1e9:.0f} % of the measured 6.53 TB/s copy bandwidth\n")
        for k in want + sorted(h for h in m if "issue_stalled" in h and "per_issue_active" in h):
            if k in m:
                fh.write(f"{k:90s} {m[k][0]:16s} {m[k][1]}\n")
    print("seg written")


def seg_traffic():
    p = os.path.join(G, "r02_seg_traffic.csv")
    if not os.path.exists(p):
        return
    rows = [r for r in csv.reader(open(p)) if len(r) > 14]
    m = {r[12]: float(r[14]) for r in rows[1:]}
    out = {"kernel": rows[1][4].split("(")[0], "launch": f"tools/quick_bench_seg.py 1024 375 500 10000 64 = the launch of `bench.py --config seg` (1024 images 375x500 to convergence, 64 distinct images cycled, grid {rows[1][8]} x {rows[1][7]} threads)",
           "command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,... --clock-control none -k regex:seg_admm -c 1 (tools/gpu_profiles.sh)",
           "dram_bytes_read": int(m["dram__bytes_read.sum"]), "dram_bytes_write": int(m["dram__bytes_write.sum"]), "lts_bytes": int(m["lts__t_bytes.sum"]),
           "l2_hit_rate_pct": m.get("lts__t_sector_hit_rate.pct"), "duration_ns": int(m["gpu__time_duration.sum"]), "warp_instructions": int(m["smsp__inst_executed.sum"]),
           "source": "profiles/r02_seg_kernel_traffic_b1024.csv"}
    out["dram_GBps_in_capture"] = (out["dram_bytes_read"] + out["dram_bytes_write"]) / out["duration_ns"]
    json.dump(out, open(os.path.join(P, "r02_seg_kernel_traffic.json"), "w"), indent=1)
    open(os.path.join(P, "r02_seg_kernel_traffic_b1024.csv"), "w").write(open(p).read())
    print("seg traffic:", out)


def bench_lines():
    for f in sorted(os.listdir(G)):
        if f.startswith("r02_bench_") and f.endswith(".json") and os.path.getsize(os.path.join(G, f)) > 10:
            open(os.path.join(P, f), "w").write(open(os.path.join(G, f)).read())
            print("copied", f)


if __name__ == "__main__":
    window_kernel(); traffic(); launches(); seg(); seg_traffic(); bench_lines()
