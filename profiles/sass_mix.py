"""Instruction mix of the hot kernels of libpvacb.so from `cuobjdump -sass` (static counts per kernel, and for the innermost loops the
counts between a backward branch target and the branch).  python profiles/sass_mix.py > profiles/r02_sass_mix.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pvac_hfhe_cppbyv_b200", "libpvacb.so")
HOT = ["sigma_fused_kernel", "prf_lpn_kernel", "concat_kernel", "dec_edges_kernel", "mul_pairs_kernel", "commit_kernel"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kernels, name = {}, None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1)
        kernels[name] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", ln)
    if m and name:
        kernels[name].append((int(m.group(1), 16), m.group(2), m.group(3)))


def family(op):
    base = op.split(".")[0]
    if base in ("LDG", "STG", "LDS", "STS", "LDC", "LDL", "STL", "ATOMS", "ATOMG", "RED", "LDSM", "UBLKCP", "SYNCS"):
        return base + ("." + ".".join(p for p in op.split(".")[1:] if p in ("128", "64", "U8", "U16", "CONSTANT")) if base in ("LDG", "STG", "LDS") else "")
    return base


print("# cuobjdump -sass of pvac_hfhe_cppbyv_b200/libpvacb.so (sm_100a). Static instruction counts; ALU pipe = LOP3/SHF/IADD3/PRMT/..., FMA pipe = IMAD/IADD via IMAD")
for mangled, ins in kernels.items():
    dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"\(.*", "", dem).replace("void ", "")
    if not any(h in short for h in HOT) or "cub::" in short:
        continue
    cnt = collections.Counter(family(op) for _, op, _ in ins)
    total = len(ins)
    print(f"\n## {short}: {total} instructions, {total * 16 / 1024:.1f} KB")
    print("   " + ", ".join(f"{k} {v}" for k, v in cnt.most_common(22)))
    # innermost loops: backward branches
    addr = {a: i for i, (a, _, _) in enumerate(ins)}
    loops = []
    for i, (a, op, rest) in enumerate(ins):
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m and int(m.group(1), 16) in addr and addr[int(m.group(1), 16)] < i:
                loops.append((addr[int(m.group(1), 16)], i))
    inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    for s, e in sorted(inner, key=lambda l: l[0] - l[1])[:4]:
        c = collections.Counter(family(op) for _, op, _ in ins[s:e + 1])
        print(f"   loop of {e - s + 1} instructions ({(e - s + 1) * 16 / 1024:.1f} KB): " + ", ".join(f"{k} {v}" for k, v in c.most_common(10)))
