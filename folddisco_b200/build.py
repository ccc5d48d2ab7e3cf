"""Builds libfolddisco_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m folddisco_b200.build [--force] [--verbose]

-fmad=false and -ffp-contract=off are REQUIRED: the geometric hash must round every f32/f64 operation the
way the reference (Rust, no FMA contraction) does; see csrc/fd_math.cuh.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
SO = os.path.join(HERE, "libfolddisco_b200.so")
CLI = os.path.join(HERE, "folddisco-b200")  # `index` / `query` front end (csrc/host/fd_cli.cpp), links the .so
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CU = ["fd_ctx.cu", "fd_hash.cu", "fd_postings.cu", "fd_query.cu", "fd_edges.cu", "fd_kabsch.cu", "fd_metrics.cu", "fd_verify.cu",
      "fd_comm.cu"]
CPP = ["host/fd_host.cpp"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function,-pthread", "-Xptxas", "-v"]


def _deps():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files]
    out.append(os.path.join(os.path.dirname(HERE), "include", "folddisco_b200.h"))
    return out


def needs_build():
    if not os.path.exists(SO) or not os.path.exists(CLI):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in CU + CPP if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src).rsplit(".", 1)[0] + ".o")
        cmd = [NVCC] + FLAGS + (["-x", "cu"] if src.endswith(".cpp") else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return src, obj, r.returncode, r.stdout

    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(compile_one, srcs))
    log = []
    for src, obj, rc, out in results:
        log.append("==== %s ====\n%s" % (src, out))
        if rc != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % src)
    with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    # NCCL: the multi-GPU exchange lives inside the library (fd_comm.cu); libnccl.so.2 is bound with dlopen on the first
    # fd_comm_* call, so that a host process that already carries an NCCL (PyTorch bundles its own) keeps exactly one
    link = [NVCC, "-shared", "-o", SO] + [o for _, o, _, _ in results] + ["-Xcompiler", "-pthread", "-ldl", "-lz"]
    subprocess.check_call(link)
    cli = ["g++", "-O2", "-std=c++17", "-Wall", "-pthread", os.path.join(CSRC, "host", "fd_cli.cpp"), "-o", CLI,
           "-L" + HERE, "-lfolddisco_b200", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath-link," + os.path.dirname(NVCC) + "/../lib64"]
    subprocess.check_call(cli)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
