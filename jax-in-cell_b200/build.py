"""Build libjic_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "jaxincell_b200", "libjic_b200.so")


def nccl_include():
    try:
        import nvidia.nccl as m  # torch's bundled NCCL (headers only needed; the library is dlopen'ed at run time)
        for base in list(getattr(m, "__path__", [])):
            inc = os.path.join(base, "include")
            if os.path.exists(os.path.join(inc, "nccl.h")):
                return inc
    except Exception:
        pass
    for inc in glob.glob(os.path.join(sys.prefix, "lib", "python*", "site-packages", "nvidia", "nccl", "include")) + ["/usr/include"]:
        if os.path.exists(os.path.join(inc, "nccl.h")):
            return inc
    raise RuntimeError("nccl.h not found")


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    srcs = glob.glob(os.path.join(CSRC, "*")) + [os.path.join(HERE, "..", "include", "jic_b200.h"), __file__]
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False, out=None, defines=None):
    """defines: {macro: int} tuning knobs of the push kernel (also read from the environment); out: alternative .so path."""
    if out is None and not force and not needs_build():
        return OUT
    out = out or OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false".replace("=false", ""),
           "-shared", "-Xcompiler", "-fPIC", "-I", nccl_include(), "-o", out, os.path.join(CSRC, "jic_engine.cu"), "-ldl"]
    cmd = [c for c in cmd if c != "--use_fast_math"]  # IEEE arithmetic: parity with the reference matters more than a few percent
    for macro in ("JIC_PUSH_THREADS", "JIC_PUSH_MINBLOCKS", "JIC_PUSH_MINBLOCKS_F32", "JIC_PUSH_STAGES", "JIC_PUSH_STAGE_BLOCKS", "JIC_MAX_CHUNK"):  # tuning knobs of the binned push kernel
        val = (defines or {}).get(macro, os.environ.get(macro))
        if val:
            cmd.insert(1, f"-D{macro}={int(val)}")
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
