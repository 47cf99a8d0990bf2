"""Per-source-line instruction profile of one kernel: joins `ncu --page source --csv` (per SASS address:
instructions executed, active lanes, stall samples) with `nvdisasm --print-line-info` of the profiled .so.

    python profiles/line_profile.py <rep.ncu-rep> <libimc_b200.so> <kernel-substring> [top]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main(rep, so, kname, top=40):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    lines_by_off = None
    for cubin in sorted(os.listdir(tmp)):
        out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
        if kname not in out:
            continue
        cur_fn, cur_line, table = None, None, {}
        for ln in out.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln) or re.match(r"\s*//-+ \.text\.(\S+)", ln)
            if m:
                cur_fn = m.group(1)
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur_fn and kname in cur_fn:
                table[int(m.group(1), 16)] = cur_line
        if table:
            lines_by_off = table
            break
    if not lines_by_off:
        raise SystemExit("kernel not found in " + so)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = next(r for r in rows if "Instructions Executed" in r)
    data = [dict(zip(hdr, x)) for x in rows if len(x) == len(hdr) and x != hdr]
    base = min(int(d["Address"], 16) if d["Address"].startswith("0x") else int(d["Address"]) for d in data)
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    tot = tt = ts = 0.0
    for d in data:
        a = int(d["Address"], 16) if d["Address"].startswith("0x") else int(d["Address"])
        key = lines_by_off.get(a - base, ("?", 0))
        ie, te, sm = num(d["Instructions Executed"]), num(d["Thread Instructions Executed"]), num(d["# Samples"])
        agg[key][0] += ie; agg[key][1] += te; agg[key][2] += sm
        tot += ie; tt += te; ts += sm
    srcs = {}
    print(f"kernel {kname}: {tot:.4g} warp instructions, {tt / tot:.2f} active lanes\n")
    print("| file:line | instr share | lanes | stall-sample share | source |\n|---|---:|---:|---:|---|")
    for (fn, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        for root in ("mixedprecisionimc.jl_b200/csrc", "include"):
            p = os.path.join(root, fn or "")
            if os.path.exists(p):
                srcs.setdefault(p, open(p).read().splitlines())
                if 0 < ln <= len(srcs[p]):
                    text = srcs[p][ln - 1].strip()[:90]
        print(f"| {fn}:{ln} | {v[0] / tot:.2%} | {v[1] / max(v[0], 1):.1f} | {v[2] / max(ts, 1):.2%} | `{text}` |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
