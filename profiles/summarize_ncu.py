"""Turn an .ncu-rep into the markdown summary committed in this directory.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep "title" >> profiles/rN_ncu_summary.md

Runs here (no GPU needed): `ncu -i rep --page raw --csv` for the launch metrics and `--page source --csv`
for the SASS opcode mix with the average number of active lanes per opcode."""
import collections
import csv
import io
import re
import subprocess
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main(rep, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[0]
    print(f"## {title}\n")
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "smsp__inst_executed_op_global_red.sum"]
    for r in rows[2:]:
        print(f"kernel: `{r[h.index('Kernel Name')]}` (launch id {r[0]})\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in keys:
            if k in h:
                print(f"| {k} | {r[h.index(k)]} | {rows[1][h.index(k)]} |")
        st = [c for c in h if c.startswith("smsp__average_warps_issue_stalled") and c.endswith("_per_issue_active.ratio")]
        vals = sorted(((num(r[h.index(c)]), c) for c in st), reverse=True)
        print("\nwarp stalls per issued instruction: " + ", ".join(
            f"{c.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, c in vals[:7]) + "\n")
    rows = list(csv.reader(io.StringIO(src)))
    hdr = next(r for r in rows if "Instructions Executed" in r)
    data = [dict(zip(hdr, x)) for x in rows if len(x) == len(hdr) and x != hdr]
    tot = sum(num(d["Instructions Executed"]) for d in data)
    tt = sum(num(d["Thread Instructions Executed"]) for d in data)
    print(f"SASS opcode mix (all captured launches; {tot:.4g} warp instructions, {tt / tot:.2f} active lanes on average):\n")
    print("| opcode | share of instructions | active lanes |\n|---|---:|---:|")
    ops = collections.defaultdict(lambda: [0.0, 0.0])
    for d in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", d["Source"])
        op = m.group(2) if m else "?"
        op = ".".join(op.split(".")[:2]) if op.startswith(("IMAD.WIDE", "MUFU", "ATOM", "RED", "LDG", "STG")) else op.split(".")[0]
        ops[op][0] += num(d["Instructions Executed"])
        ops[op][1] += num(d["Thread Instructions Executed"])
    for k, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:18]:
        print(f"| {k} | {v[0] / tot:.2%} | {v[1] / max(v[0], 1):.1f} |")
    print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
