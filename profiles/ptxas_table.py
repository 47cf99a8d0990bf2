"""Turns `nvcc -Xptxas -v` logs (one per precision translation unit) into the register / spill / shared-memory table of
profiles/r2_ptxas.md.

    for p in f16 f32 f64; do nvcc <the flags of __graft_entry__.NVCC_FLAGS> -Xptxas -v -c mixedprecisionimc.jl_b200/csrc/imc_engine_$p.cu -o /dev/null 2> build/ptxas_$p.log; done
    python profiles/ptxas_table.py build/ptxas_f16.log build/ptxas_f32.log build/ptxas_f64.log > profiles/r2_ptxas.md
"""
import re
import subprocess
import sys

rows = []
for path in sys.argv[1:]:
    text = open(path).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?Function properties for \S+\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(.*?)\n", text, flags=re.S):
        name, stack, sst, sld, regs, rest = m.groups()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"^void imc::", "", dem)
        dem = re.sub(r"\(.*\)$", "", dem).replace("imc::", "")
        smem = re.search(r"(\d+) bytes smem", rest)
        rows.append((dem, int(regs), int(stack), int(sst), int(sld), int(smem.group(1)) if smem else 0))
print("# ptxas -v, sm_100a, kernels of the engine as built (registers per thread, stack frame / spill bytes, static shared memory)\n")
print("Template arguments of the tracking kernels: <precision, [geometry,] replay tape, tally kind (-1 run-time, 0 ATOMIC global, 1 ATOMIC shared,")
print("2 FIXED global, 3 FIXED shared)>.  The Float32 / Float16 tracking kernels are built for 4 blocks of 256 threads per SM (<= 64 registers),")
print("the Float64 ones for 3 (<= 85).\n")
print("| kernel | registers | stack | spill stores | spill loads | static smem |")
print("|---|---:|---:|---:|---:|---:|")
for r in sorted(rows):
    print(f"| `{r[0]}` | {r[1]} | {r[2]} | {r[3]} | {r[4]} | {r[5]} |")
