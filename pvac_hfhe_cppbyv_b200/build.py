"""Builds libpvacb.so (CUDA, sm_100a only) in-tree with nvcc. No JIT, no torch extension machinery: the .so is a plain
C-ABI library (include/pvacb.h) that travels with the repository snapshot to the GPU box.

    python -m pvac_hfhe_cppbyv_b200.build [--force]
"""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libpvacb.so")

SOURCES = ["engine.cu", "boundary.cu", "group.cu", "prf.cu", "sigma.cu", "enc.cu", "arith.cu", "mul.cu", "compact.cu", "commit.cu", "text.cu", "recrypt.cu", "dec.cu", "extras.cu", "keygen.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(nvcc, src, obj, log):
    cmd = [nvcc] + NVCC_FLAGS + (["-x", "cu"] if src.endswith(".cpp") else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return src


def build(force=False, verbose=False, tuning=False):
    """tuning=True builds libpvacb_tuning.so with -DPVACB_TUNING: the experiment shapes of the sigma kernel and the probe's
    load flavours behind environment switches (profiles/README.md). The product library has none of them."""
    global OBJ, LIB, NVCC_FLAGS
    if tuning:
        OBJ, LIB = os.path.join(HERE, "_build_tuning"), os.path.join(HERE, "libpvacb_tuning.so")
        NVCC_FLAGS = NVCC_FLAGS + ["-DPVACB_TUNING"]
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "pvacb.h"))
    stamp = os.path.join(OBJ, "stamp")
    want = _digest(headers + [os.path.join(CSRC, s) for s in SOURCES])
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == want:
        return LIB
    nvcc = _nvcc()
    hdr_digest = _digest(headers)
    jobs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        for s in SOURCES:
            obj = os.path.join(OBJ, s + ".o")
            tag = os.path.join(OBJ, s + ".tag")
            d = _digest([os.path.join(CSRC, s)]) + hdr_digest
            if not force and os.path.exists(obj) and os.path.exists(tag) and open(tag).read() == d:
                continue
            jobs.append((ex.submit(_compile, nvcc, s, obj, os.path.join(OBJ, s + ".log")), tag, d))
        for fut, tag, d in jobs:
            name = fut.result()
            with open(tag, "w") as f:
                f.write(d)
            if verbose:
                print("compiled", name)
    objs = [os.path.join(OBJ, s + ".o") for s in SOURCES]
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as f:
        f.write(want)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, tuning="--tuning" in sys.argv))
