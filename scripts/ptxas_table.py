"""profiles/ptxas_<tag>.txt: registers / stack / spills of every kernel instantiation (nvcc -Xptxas -v, sm_100a).
    python scripts/ptxas_table.py [tag]        (compiles the library once more into /tmp, ~2.5 min)"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
src = os.path.join(ROOT, "generalized-icp_b200", "csrc", "gicp_b200.cu")
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC",
       "-shared", "-Xptxas", "-v", "-o", "/tmp/libgicp_ptxas.so", src, "-ldl"]
err = subprocess.run(cmd, capture_output=True, text=True, check=True).stderr
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Compiling entry function '(\S+)'", err)),
                       capture_output=True, text=True).stdout.splitlines()
blocks = re.split(r"ptxas info\s+: Compiling entry function ", err)[1:]
rows = []
for name, blk in zip(names, blocks):
    if "gicp::" not in name:
        continue
    short = re.sub(r"^void ", "", name).replace("gicp::", "").split("(")[0]
    used = re.search(r"Used (\d+ registers[^\n]*)", blk).group(1)
    used = re.sub(r", \d+ bytes cmem\[\d+\]", "", used)
    stack = re.search(r"(\d+ bytes stack frame, \d+ bytes spill stores, \d+ bytes spill loads)", blk).group(1)
    rows.append(f"{short:52s} Used {used}; {stack}")
with open(os.path.join(ROOT, "profiles", f"ptxas_{tag}.txt"), "w") as f:
    f.write("# nvcc -Xptxas -v, sm_100a, final code of the round (template arguments: <dim, storage type[, list capacity]>)\n")
    f.write("\n".join(sorted(rows)) + "\n")
print(len(rows), "kernels")
