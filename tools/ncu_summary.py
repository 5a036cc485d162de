#!/usr/bin/env python3
"""Summarise an ncu report (and the launch list of the same session) into profiles/<tag>_*.{md,csv}.
usage: tools/ncu_summary.py <tag>   (reads gpurun_out/prof_<tag>.ncu-rep, launches_<tag>.csv, bench_<tag>.json)"""
import collections, csv, io, json, os, subprocess, sys
tag = sys.argv[1]
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
g = lambda f: os.path.join(root, "gpurun_out", f)
out = os.path.join(root, "profiles")
os.makedirs(out, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.max.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
STALL = "smsp__average_warps_issue_stalled_"
lines = ["# ncu summary `%s`" % tag, ""]
rep = g("prof_%s.ncu-rep" % tag)
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for k, row in enumerate(rows[2:]):
        d = dict(zip(hdr, row)); u = dict(zip(hdr, units))
        lines += ["## kernel `%s` (launch %d of the capture; `ncu --set full --clock-control none`)" % (d.get("Kernel Name", "?"), k), "",
                  "| metric | value | unit |", "|---|---|---|"]
        for key in KEYS:
            if key in d:
                lines.append("| %s | %s | %s |" % (key, d[key], u[key]))
        if k == 0 and "lpcnet_sample_kernel" in d.get("Kernel Name", ""):
            def num(key):
                v = float(d[key].replace(",", "")); unit = u[key].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
            grid = int(float(d["launch__grid_size"].replace(",", "")))
            frames = int(os.environ.get("NCU_FRAMES", "3"))
            json.dump({"dram_bytes": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"), "grid": grid, "frames": frames,
                       "samples_in_launch": 4096 * frames * 160,
                       "l1_data_pipe_pct_of_peak_active_sms": float(d.get("l1tex__throughput.avg.pct_of_peak_sustained_active", "nan").replace(",", "")),
                       "issue_active_pct": float(d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "nan").replace(",", "")),
                       "note": "ncu --set full capture of lpcnet_sample_kernel at 4096 streams x %d frames" % frames},
                      open(os.path.join(out, "%s_traffic.json" % tag), "w"), indent=1)
        st = sorted(((float(d[h].replace(",", "")), h[len(STALL):-len("_per_issue_active.ratio")]) for h in hdr
                     if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and d[h]), reverse=True)
        lines += ["", "warp stall reasons (warps stalled per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in st[:8]), ""]
ll = g("launches_%s.csv" % tag)
if os.path.exists(ll):
    rows = [r for r in csv.reader(open(ll)) if len(r) > 5]
    hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    lines += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache & serialised: compare SHARES)", "",
              "| kernel | launches | total ms | avg us | share |", "|---|---|---|---|---|"]
    for k, v in agg.items():
        lines.append("| %s | %d | %.3f | %.1f | %.1f%% |" % (k.split("(")[0], len(v), sum(v) / 1e6, sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    import shutil
    shutil.copy(ll, os.path.join(out, "%s_launches.csv" % tag))
bj = g("bench_%s.json" % tag)
if os.path.exists(bj) and os.path.getsize(bj):
    b = json.loads(open(bj).read().strip().splitlines()[-1])
    lines += ["", "## bench line of the same session (not under a profiler)", "", "```json", json.dumps(b, indent=1), "```"]
    import shutil
    shutil.copy(bj, os.path.join(out, "%s_bench.json" % tag))
open(os.path.join(out, "%s_ncu_summary.md" % tag), "w").write("\n".join(lines) + "\n")
print("wrote profiles/%s_ncu_summary.md" % tag)
