"""Registers / stack / spills / static shared memory of every kernel of libpvacb.so, from the `-Xptxas -v` output the in-tree build
keeps per translation unit (pvac_hfhe_cppbyv_b200/_build/*.log).  python profiles/ptxas_summary.py > profiles/r02_ptxas.txt"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"Compiling entry function '([^']+)' for 'sm_100a'\n(?:ptxas info\s*:\s*Function properties for [^\n]+\n\s*(\d+) bytes stack frame, "
                 r"(\d+) bytes spill stores, (\d+) bytes spill loads\n)?ptxas info\s*:\s*Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?")

print("# nvcc 12.9 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xptxas -v; dynamic shared memory is not in this list")
print(f"{'translation unit':14s} {'kernel':58s} {'regs':>4s} {'stack':>5s} {'spill st/ld':>11s} {'static smem':>11s}")
for f in sorted(glob.glob(os.path.join(ROOT, "pvac_hfhe_cppbyv_b200", "_build", "*.log"))):
    for m in PAT.finditer(open(f).read()):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        if "cub::" in name:
            name = "cub::" + name.split("::")[2].split("<")[0] + " (library)"
        print(f"{os.path.basename(f)[:-4]:14s} {name[:58]:58s} {m.group(5):>4s} {m.group(2) or '0':>5s} {(m.group(3) or '0') + '/' + (m.group(4) or '0'):>11s} {m.group(7) or '0':>11s}")
