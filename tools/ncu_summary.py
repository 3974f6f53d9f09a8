"""Summaries of an ncu capture for profiles/: the headline metrics of the raw page and, from the source page, the hot
loops (runs of instructions executed about once per warp-hour) with their instruction counts, FP64 share and
warp-stall samples by reason -- for the pair kernel that is one block per role and pass.

    ncu -i X.ncu-rep --page raw --csv > X_raw.csv ; ncu -i X.ncu-rep --page source --csv > X_src.csv
    python tools/ncu_summary.py X_raw.csv X_src.csv <warp-hours per launch> > profiles/...
"""
import csv
import sys

RAW = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
       "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
       "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warp_latency_per_inst_issued.ratio",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
       "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
STALLS = ["stall_wait", "stall_short_sb", "stall_long_sb", "stall_barrier", "stall_math", "stall_not_selected",
          "stall_selected", "stall_dispatch", "stall_branch_resolving", "stall_no_inst", "stall_lg", "stall_mio"]


def main(raw_csv, src_csv, warp_hours):
    rows = list(csv.reader(open(raw_csv)))
    d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
    print("# kernel:", d.get("Kernel Name", ("", "?"))[1])
    for m in RAW:
        if m in d:
            print(f"{m:88s} {d[m][0]:16s} {d[m][1]}")
    rows = list(csv.reader(open(src_csv)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[start]
    idx = {n: i for i, n in enumerate(hdr)}
    data = [r for r in rows[start + 1:] if len(r) >= len(hdr)]
    ex = [int(r[idx["Instructions Executed"]]) for r in data]
    total = sum(ex)
    print(f"\n# executed warp instructions {total} = {total / warp_hours:.1f} per warp-hour "
          f"({warp_hours:.0f} warp-hours per launch); warp-stall samples {sum(int(r[idx['# Samples']]) for r in data)}")
    # hot loops: maximal runs of instructions executed >= 0.45 x the most common large count
    big = sorted(e for e in ex if e > 0.2 * max(ex))
    per_hour = big[len(big) // 2]
    runs, cur = [], None
    for k, e in enumerate(ex):
        hot = e >= 0.9 * per_hour
        if hot and cur is None:
            cur = k
        if not hot and cur is not None:
            if k - cur > 20:
                runs.append((cur, k))
            cur = None
    print(f"# hot loops (instructions executed ~{per_hour} times = once per warp-hour of a pass):")
    fp64 = ("DFMA", "DMUL", "DADD", "DSETP")
    for a, b in runs:
        ops = [data[k][idx["Source"]].split()[0 if not data[k][idx["Source"]].strip().startswith("@") else 1] for k in range(a, b)]
        n_fp = sum(1 for o in ops if o.startswith(fp64))
        smp = sum(int(data[k][idx["# Samples"]]) for k in range(a, b))
        has_bar = any(o.startswith("BAR") for o in ops)
        role = "forcing warp (writes the ring: STS)" if has_bar and any(o.startswith("STS") for o in ops) else \
               ("state warp (reads the ring after the barrier)" if has_bar else "hour loop (one role: monolithic form)")
        st = {s: sum(int(data[k][idx[s]]) for k in range(a, b)) for s in STALLS if s in idx}
        st = {k: v for k, v in st.items() if v}
        print(f"  instr {a:5d}..{b:5d}: {b - a:4d} instructions / warp-hour, {n_fp:3d} FP64, {smp:8d} samples, {role}")
        print("      stalls: " + ", ".join(f"{k[6:]} {v} ({100.0 * v / max(smp, 1):.0f} %)" for k, v in sorted(st.items(), key=lambda kv: -kv[1])))
    # opcode mix over everything
    mix = {}
    for r, e in zip(data, ex):
        t = r[idx["Source"]].split()
        op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
        mix[op] = mix.get(op, 0) + e
    print("# opcode mix (executed warp instructions per warp-hour):")
    print("  " + "  ".join(f"{op} {n / warp_hours:.1f}" for op, n in sorted(mix.items(), key=lambda kv: -kv[1])[:24]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]))
