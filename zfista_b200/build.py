"""In-tree build of libzfista_b200.so with nvcc for sm_100a.

    python -m zfista_b200.build [--force] [--verbose]

Every .cu under zfista_b200/csrc is compiled (in parallel) with
``-gencode arch=compute_100a,code=sm_100a -lineinfo`` and linked into
``zfista_b200/libzfista_b200.so``.  The .so is git-ignored but travels to the GPU
box with the repo snapshot.
"""
from __future__ import annotations

import argparse
import fcntl
import hashlib
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
BUILD_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libzfista_b200.so")

# flags that define the build (hashed into the fingerprint); the include paths are added at
# compile time only: the repo is mounted at different absolute paths here and on the GPU box,
# and a path in the fingerprint would force a rebuild there on every run
NVCC_BASE_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
]
NVCC_FLAGS = NVCC_BASE_FLAGS + ["-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(d, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(" ".join(NVCC_BASE_FLAGS).encode())
    return h.hexdigest()


def _up_to_date(fp: str) -> bool:
    stamp = os.path.join(BUILD_DIR, "fingerprint.txt")
    if not (os.path.exists(LIB_PATH) and os.path.exists(stamp)):
        return False
    with open(stamp) as fh:
        return fh.read().strip() == fp


def build(force: bool = False, verbose: bool = False) -> str:
    """Build (if the sources changed) and return the library path.  Safe to call from several
    processes at once (one rank per GPU under torchrun): an exclusive file lock lets one of
    them build while the others wait, and the library is moved into place atomically so that
    nobody can ever map a half-written file."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    fp = _fingerprint()
    if not force and _up_to_date(fp):
        return LIB_PATH
    with open(os.path.join(BUILD_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _up_to_date(fp):      # another process built it meanwhile
                return LIB_PATH
            return _build_locked(fp, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(fp: str, verbose: bool) -> str:
    stamp = os.path.join(BUILD_DIR, "fingerprint.txt")
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, flush=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    fd, tmp = tempfile.mkstemp(prefix=".libzfista_b200.", suffix=".so.tmp", dir=PKG_DIR)
    os.close(fd)
    try:
        r = subprocess.run([nvcc, "-shared", "-o", tmp, *objs, "-lcudart"], capture_output=True,
                           text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.chmod(tmp, 0o755)
        os.replace(tmp, LIB_PATH)          # atomic: readers see the old or the new file
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)
    with open(stamp, "w") as fh:
        fh.write(fp)
    return LIB_PATH


SASS_MNEMONICS = ("UBLKCP", "UTMALDG", "DMMA", "STAS", "SYNCS", "UCGABAR", "DFMA", "LDGSTS")


def sass_counts(out_path: str | None = None) -> str:
    """Per-kernel counts of the SASS mnemonics that prove what the kernels are made of (bulk TMA
    UBLKCP, tensor-map TMA UTMALDG, FP64 tensor-core DMMA, st.async STAS, mbarrier SYNCS, cluster
    barrier UCGABAR, plain DFMA), from `cuobjdump -sass` of the objects build() made.  Returns
    the table as text and writes it to `out_path` (profiles/sass_counts.txt) when given."""
    import collections
    import re

    cuobjdump = os.path.join(os.path.dirname(_nvcc()), "cuobjdump") if os.path.isabs(_nvcc()) \
        else "cuobjdump"
    rows = []
    for obj in sorted(f for f in os.listdir(BUILD_DIR) if f.endswith(".o")):
        r = subprocess.run([cuobjdump, "-sass", os.path.join(BUILD_DIR, obj)],
                           capture_output=True, text=True)
        if r.returncode != 0:
            continue
        fn, arch = None, "?"
        counts = collections.OrderedDict()
        for ln in r.stdout.splitlines():
            m = re.match(r"\s*arch = (sm_\w+)", ln)
            if m:
                arch = m.group(1)
            m = re.match(r"\s*Function : (\S+)", ln)
            if m:
                fn = m.group(1)
                counts[fn] = collections.Counter(arch=arch)
                continue
            if fn is None:
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if m:
                op = m.group(1)
                counts[fn]["total"] += 1
                for mn in SASS_MNEMONICS:
                    if op.startswith(mn):
                        counts[fn][mn] += 1
        for fn, c in counts.items():
            r2 = subprocess.run(["c++filt", fn], capture_output=True, text=True)
            name = (r2.stdout.strip() or fn).split("(")[0].replace("void ", "")
            rows.append((obj, name[:70], c["arch"], c["total"], [c[mn] for mn in SASS_MNEMONICS]))
    head = f"{'object':28s} {'kernel':70s} {'arch':8s} {'instr':>7s} " + " ".join(
        f"{mn:>7s}" for mn in SASS_MNEMONICS)
    lines = ["# cuobjdump -sass of zfista_b200/build/*.o (written by __graft_entry__.build())", head]
    for obj, name, arch, total, cs in rows:
        lines.append(f"{obj:28s} {name:70s} {arch:8s} {total:7d} " + " ".join(f"{c:7d}" for c in cs))
    text = "\n".join(lines) + "\n"
    if out_path:
        with open(out_path, "w") as fh:
            fh.write(text)
    return text


def sass_stale(out_path: str) -> bool:
    """True if `out_path` is missing or older than the library (so build() rewrites it once)."""
    return (not os.path.exists(out_path)) or (
        os.path.exists(LIB_PATH) and os.path.getmtime(out_path) < os.path.getmtime(LIB_PATH))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
    sys.exit(0)
