"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + SASS hot spots.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [--sass N]"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("== kernel:", name[:90])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("  %-88s %s %s" % (k, r[i], units[i]))
    for i, h in enumerate(hdr):
        if h.startswith("l1tex__data_pipe_lsu_wavefronts") and h.endswith(".sum") and r[i] not in ("0", ""):
            print("  %-88s %s" % (h, r[i]))
        if h == "TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg" or h.endswith("l1tex__data_pipe_lsu_wavefronts.avg"):
            print("  %-88s %s" % (h, r[i]))
nsass = 0
if "--sass" in sys.argv:
    nsass = int(sys.argv[sys.argv.index("--sass") + 1])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
data = []
for r in rows:
    if r and r[0] == "Address":
        h = r
        continue
    if h and len(r) == len(h):
        data.append(r)
if h:
    iS, iE, iSm, iT = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Avg. Threads Executed")
    tot = sum(int(r[iE]) for r in data) or 1
    totS = sum(int(r[iSm]) for r in data) or 1
    ops, opsS = collections.Counter(), collections.Counter()
    for r in data:
        t = r[iS].split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        ops[op] += int(r[iE]); opsS[op] += int(r[iSm])
    print("== SASS: %d instructions, %d warp-instructions executed, %d samples" % (len(data), tot, totS))
    for op, c in ops.most_common(18):
        print("  %-18s %6.2f%% of executed  %6.2f%% of samples" % (op, 100 * c / tot, 100 * opsS[op] / totS))
    if nsass:
        print("== regions of %d SASS instructions" % nsass)
        for i in range(0, len(data), nsass):
            ch = data[i:i + nsass]
            e = sum(int(r[iE]) for r in ch); s = sum(int(r[iSm]) for r in ch)
            thr = sum(float(r[iT]) * int(r[iE]) for r in ch) / max(e, 1)
            if e * 200 > tot:
                print("  @%4d %5.1f%% executed %5.1f%% samples, %.1f threads/instr   %s" % (i, 100 * e / tot, 100 * s / totS, thr, ch[0][iS].strip()[:60]))
