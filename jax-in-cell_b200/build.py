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
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-I", nccl_include(), "-o", out, os.path.join(CSRC, "jic_engine.cu"), "-ldl"]
    # (no --use_fast_math: IEEE arithmetic -- parity with the reference matters more than a few percent)
    for macro in ("JIC_PUSH_THREADS", "JIC_PUSH_MINBLOCKS", "JIC_PUSH_THREADS_F32", "JIC_PUSH_MINBLOCKS_F32", "JIC_PUSH_STAGES", "JIC_PUSH_STAGES_F32", "JIC_PUSH_STAGE_BLOCKS", "JIC_PUSH_STAGE_BLOCKS_F32", "JIC_PUSH_RUN", "JIC_ITEMS_PER_WARP", "JIC_TAIL_SPLIT", "JIC_MAX_CHUNK"):  # tuning knobs of the binned push kernel
        val = (defines or {}).get(macro, os.environ.get(macro))
        if val is not None and val != "":
            cmd.insert(1, f"-D{macro}={int(val)}")
    for item in filter(None, ((defines or {}).get("EXTRA") or os.environ.get("JIC_EXTRA_DEFINES", "")).split(",")):  # experiments: "A=1,B=2"
        cmd.insert(1, "-D" + item)
    maxreg = (defines or {}).get("MAXRREG", os.environ.get("JIC_MAXRREG"))
    if maxreg:
        cmd.insert(1, f"-maxrregcount={int(maxreg)}")
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


def build_ffi(out=None):
    """libjic_b200_ffi.so: the XLA FFI handlers of csrc/jic_xla_ffi.cc over libjic_b200.so.  Needs jax (for jax.ffi.include_dir());
    without it the source reduces to a stub and this function refuses to build a library that could not register anything."""
    try:
        import jax.ffi
        xla_include = jax.ffi.include_dir()
    except ImportError as e:
        raise RuntimeError("jax is not installed: the XLA FFI headers (jax.ffi.include_dir()) are unavailable") from e
    build()
    out = out or os.path.join(HERE, "jaxincell_b200", "libjic_b200_ffi.so")
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-I", xla_include, "-I", os.path.join(cuda_home, "include"),
           os.path.join(CSRC, "jic_xla_ffi.cc"), "-L", os.path.dirname(OUT), "-l:libjic_b200.so", "-L", os.path.join(cuda_home, "lib64"),
           "-lcudart", "-Wl,-rpath,$ORIGIN", "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    if "--ffi" in sys.argv:
        print(build_ffi())
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
