"""Build libimc_b200 variants for A/B runs on the GPU box:  python scratch/build_variant.py NAME [-Dflag ...] [--src DIR]
Output: variants/libimc_NAME.so (git-ignored, travels with gpurun).  bench.py / tests pick a variant with IMC_LIB=path."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = sys.argv[1]
args = sys.argv[2:]
src_root = ROOT
if "--src" in args:
    i = args.index("--src"); src_root = args[i + 1]; del args[i:i + 2]
csrc = os.path.join(src_root, "mixedprecisionimc.jl_b200", "csrc")
bdir = os.path.join(ROOT, "build", "var_" + name)
os.makedirs(bdir, exist_ok=True)
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
flags = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-mfma,-O2",
         "-I" + os.path.join(src_root, "include"), "-I" + csrc] + args
srcs = ["imc_exact.cu", "imc_selftest.cu", "imc_capi.cu", "imc_engine_f16.cu", "imc_engine_f32.cu", "imc_engine_f64.cu"]
def cc(s):
    o = os.path.join(bdir, s.replace(".cu", ".o"))
    r = subprocess.run(["nvcc"] + flags + ["-c", os.path.join(csrc, s), "-o", o], capture_output=True, text=True)
    if r.returncode: raise SystemExit(r.stderr)
    return o
with ThreadPoolExecutor(6) as ex: objs = list(ex.map(cc, srcs))
out = os.path.join(ROOT, "variants", f"libimc_{name}.so")
subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs, check=True)
r = subprocess.run(["cuobjdump", "--dump-resource-usage", os.path.join(bdir, "imc_engine_f32.o")], capture_output=True, text=True).stdout.splitlines()
for i, l in enumerate(r):
    if "k_track_refillINS_3F32ELi2ELb0ELi0" in l: print(name, r[i + 1].strip()[:60])
print(out)
