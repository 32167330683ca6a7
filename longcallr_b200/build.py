"""Builds the two in-tree shared libraries (no JIT cache: the .so files travel with the repo snapshot).

  lib/liblcr_host.so         g++   csrc/host/*.cpp            (BAM/FASTA decode, regions, VCF, synthetic data)
  lib/liblongcallr_b200.so   nvcc  csrc/device/*.cu, *.cpp    (the C ABI + sm_100a kernels)

-fmad=false / -ffp-contract=off: include/lcr_contract.h promises bit-identical math on host and device.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
INC = os.path.join(ROOT, "include")
LIB = os.path.join(HERE, "lib")
OBJ = os.path.join(HERE, "build")
HOST_SRC = ["lcr_host.cpp", "vcf.cpp", "synth.cpp", "bam_io.cpp"]
DEV_SRC = ["api.cu", "pileup.cu", "fragments.cu", "phase.cu", "phase_enum.cu", "regions.cu", "params.cpp"]
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
EXTRA = os.environ.get("LCR_NVCC_EXTRA", "").split()
NVCC_FLAGS = EXTRA + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off", "-Xptxas", "-v", "--expt-relaxed-constexpr", "--extended-lambda"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, log=None):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def _headers():
    hs = [os.path.join(INC, f) for f in os.listdir(INC)]
    for d in ("host", "device"):
        p = os.path.join(HERE, "csrc", d)
        hs += [os.path.join(p, f) for f in os.listdir(p) if f.endswith(".h")]
    return hs


def build_host(force=False):
    os.makedirs(LIB, exist_ok=True)
    src = [os.path.join(HERE, "csrc", "host", f) for f in HOST_SRC]
    out = os.path.join(LIB, "liblcr_host.so")
    if force or _stale(out, src + _headers()):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wall", "-pthread", "-I" + INC,
              "-I" + os.path.join(HERE, "csrc", "host"), "-shared", "-o", out] + src + ["-lz"])
    return out


def build_device(force=False):
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    out = os.path.join(LIB, "liblongcallr_b200.so")
    hdrs = _headers()
    jobs = []
    for f in DEV_SRC:
        src = os.path.join(HERE, "csrc", "device", f)
        obj = os.path.join(OBJ, f + ".o")
        if force or _stale(obj, [src] + hdrs):
            jobs.append((src, obj))
    objs = [os.path.join(OBJ, f + ".o") for f in DEV_SRC]

    def compile_one(job):
        src, obj = job
        return _run([NVCC] + NVCC_FLAGS + ["-I" + INC, "-I" + os.path.join(HERE, "csrc", "device"), "-c", src, "-o", obj], log=obj + ".log")

    with concurrent.futures.ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(compile_one, jobs))
    if force or jobs or _stale(out, objs):
        _run([NVCC, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    return out


def build_all(force=False):
    return build_host(force), build_device(force)


if __name__ == "__main__":
    print(build_all("--force" in sys.argv))
