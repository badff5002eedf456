"""Summarise an .ncu-rep (raw + source pages) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/X.txt"""
import csv, subprocess, sys, io
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "gpc__cycles_elapsed.max", "smsp__cycles_elapsed.avg.per_second", "sm__cycles_active.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "smsp__inst_executed.sum"]
lines = []
for r in rows[2:]:
    lines.append("=" * 100)
    for h, u, v in zip(hdr, units, r):
        if h in keys or h.startswith("smsp__average_warps_issue_stalled") and "per_issue_active" in h and float(v or 0) > 0.05:
            lines.append(f"{h:95s} {u:14s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
for i, r in enumerate(srows):
    if r and r[0] == "Address":
        h2 = r; idx = {h: j for j, h in enumerate(h2)}; body = []
        for q in srows[i + 1:]:
            if not q or q[0] in ("Kernel Name", "Address"): break
            if len(q) == len(h2): body.append(q)
        tot = sum(int(q[idx["# Samples"]]) for q in body) or 1
        lines.append("-" * 100)
        lines.append(f"source page: {len(body)} SASS instructions, {tot} stall samples; top 15 by samples:")
        for q in sorted(body, key=lambda q: -int(q[idx["# Samples"]]))[:15]:
            st = {h.replace("stall_", ""): int(q[idx[h]]) for h in h2 if h.startswith("stall_") and "Not Issued" not in h and int(q[idx[h]]) > 0.1 * int(q[idx["# Samples"]])}
            lines.append(f"  {100.0 * int(q[idx['# Samples']]) / tot:5.1f}%  {q[idx['Source']].strip()[:58]:58s} {st}")
        if "L1 Wavefronts Shared Excessive" in idx:
            exc = sum(int(q[idx["L1 Wavefronts Shared Excessive"]] or 0) for q in body); wf = sum(int(q[idx["L1 Wavefronts Shared"]] or 0) for q in body)
            lines.append(f"shared wavefronts {wf}, excessive {exc}")
        if "L2 Theoretical Sectors Global Excessive" in idx:
            exg = sum(int(q[idx["L2 Theoretical Sectors Global Excessive"]] or 0) for q in body); sg = sum(int(q[idx["L2 Theoretical Sectors Global"]] or 0) for q in body)
            lines.append(f"L2 theoretical sectors global {sg}, excessive {exg}")
        mn = {}
        for q in body:
            op = q[idx["Source"]].strip().split()
            op = op[1] if op and op[0].startswith("@") else (op[0] if op else "")
            op = op.split(".")[0]
            mn[op] = mn.get(op, 0) + int(q[idx["Instructions Executed"]])
        lines.append("instructions executed by opcode: " + ", ".join(f"{k}={v}" for k, v in sorted(mn.items(), key=lambda kv: -kv[1])[:12]))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
