"""Developer tool: text summary of an `ncu --set full --import-source on` report for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index] > profiles/rNN_ncu_<what>.txt
Prints the counters SURVEY.md 8d asks for (dram bytes, issue / tensor / shared-memory utilisation, stalls) and, when the
report carries the source page, the executed-instruction mix and the hottest stall sites."""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, row = rows[0], rows[1], rows[2 + kidx]
want = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sector_hit_rate.pct",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__cycles_elapsed.avg.per_second",
]
for w in want:
    if w in h:
        i = h.index(w)
        print(f"{w:72s} {row[i]} {units[i]}")
st = [(float(row[i] or 0), k) for i, k in enumerate(h)
      if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", k) and row[i]]
print("stalls per issue-active cycle: " + ", ".join(
    f"{k.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for v, k in sorted(st, reverse=True)[:10]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
if len(blocks) > 1 + kidx:
    r = list(csv.reader(io.StringIO('"Kernel Name"' + blocks[1 + kidx])))
    hh = r[1]
    if "Source" in hh and "Instructions Executed" in hh:
        iS, iE, iP = hh.index("Source"), hh.index("Instructions Executed"), hh.index("Warp Stall Sampling (All Samples)")
        ops, tot, lines = collections.Counter(), 0, []
        for x in r[2:]:
            if len(x) < len(hh):
                continue
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", x[iS].strip())
            e = int(x[iE] or 0)
            ops[m.group(2) if m else "?"] += e
            tot += e
            lines.append((int(x[iP] or 0), e, x[iS].strip()))
        print(f"executed warp instructions {tot}, SASS lines {len(lines)}")
        print("mix: " + ", ".join(f"{k} {v} ({100 * v / tot:.1f}%)" for k, v in ops.most_common(24)))
        print("hottest stall sites (samples, executions, instruction):")
        for s_, e, t in sorted(lines, reverse=True)[:16]:
            print(f"  {s_:5d} {e:9d}  {t}")
