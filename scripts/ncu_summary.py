"""Summarise an .ncu-rep (key metrics, stall reasons, opcode mix) as text for profiles/."""
import csv, io, subprocess, sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__cycles_elapsed.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "sm__cycles_elapsed.avg.per_second"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("== kernel:", d.get("Kernel Name", "?")[:110])
    for k in KEYS:
        if k in d:
            print("  %-62s %s %s" % (k, d[k], units[hdr.index(k)]))
    st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(v))
          for h, v in d.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v]
    print("  stall reasons (warps per issue):", ", ".join("%s=%.2f" % s for s in sorted(st, key=lambda x: -x[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
try:
    h = rows[1]
    iS, iN, iE = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    c, s = Counter(), Counter()
    for r in rows[2:]:
        if len(r) <= iE or not r[iE].isdigit():
            continue
        t = r[iS].split()
        op = t[1] if t[0].startswith("@") else t[0]
        c[op] += int(r[iE]); s[op] += int(r[iN])
    tot, tots = sum(c.values()), max(sum(s.values()), 1)
    print("  executed warp-instructions: %d; top opcodes (%% instr / %% stall samples):" % tot)
    print("   " + "  ".join("%s %.1f/%.1f" % (op, 100 * n / tot, 100 * s[op] / tots) for op, n in c.most_common(16)))
except Exception as e:  # noqa
    print("  (no source page: %s)" % e)
