"""The C-ABI library loads on a CPU-only box and exports every symbol include/jic_b200.h declares (no compute calls)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("jic_build", os.path.join(ROOT, "jax-in-cell_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from jaxincell_b200 import _lib
    return _lib


def test_header_symbols_are_exported_and_bound(lib):
    header = open(os.path.join(ROOT, "include", "jic_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|void|const char\*)\s+(jic_[a-z_0-9]+)\s*\(", header, flags=re.M))
    assert len(declared) >= 14
    bound = {name for name, _, _ in lib.SYMBOLS}
    assert declared == bound, declared ^ bound
    handle = lib.load()
    for name in declared:
        assert hasattr(handle, name)
    assert handle.jic_abi_version() == 1


def test_struct_layout_matches_header(lib, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "jic_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(jic_params),sizeof(jic_species),sizeof(jic_outputs),offsetof(jic_params,filter_alpha),offsetof(jic_params,grid_first));return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(lib.Params), C.sizeof(lib.Species), C.sizeof(lib.Outputs), lib.Params.filter_alpha.offset, lib.Params.grid_first.offset]


def test_invalid_arguments_are_reported_without_a_gpu(lib):
    h = lib.load()
    ctx = C.c_void_p()
    p = lib.Params()
    sp = (lib.Species * 1)()
    assert h.jic_create(C.byref(p), sp, C.byref(ctx)) == -1  # struct_bytes == 0 -> ABI mismatch
    assert b"ABI mismatch" in h.jic_last_error(None)
    p.struct_bytes = C.sizeof(lib.Params)
    p.n_grid, p.n_species, p.length, p.dx, p.dt = 2, 1, 1.0, 0.5, 1e-9
    assert h.jic_create(C.byref(p), sp, C.byref(ctx)) == -1
    assert b"number_grid_points" in h.jic_last_error(None)
    assert h.jic_run(None, 1, None, None) == -1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from jaxincell_b200 import HotPath, JicError
    with pytest.raises(JicError, match="no CUDA device"):
        HotPath(species=[dict(count=1, q=1.0, m=1.0, qm=1.0)], length=1.0, G=8, dt=1e-9)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "jax-in-cell_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle"


# ---- the jax.ffi glue (csrc/jic_xla_ffi.cc): cannot run here (no jax / XLA headers), but it must at least parse and type-check ----
FFI_SRC = os.path.join(ROOT, "jax-in-cell_b200", "csrc", "jic_xla_ffi.cc")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")


def test_xla_ffi_glue_is_a_stub_without_the_xla_headers(tmp_path):
    so = str(tmp_path / "ffi_stub.so")
    subprocess.run(["g++", "-std=c++17", "-shared", "-fPIC", "-I", CUDA_INC, FFI_SRC, "-o", so], check=True)
    assert C.CDLL(so).jic_xla_ffi_available() == 0


def test_xla_ffi_glue_type_checks():
    """Against tests/ffi_mock (a compile-check mock of the names the glue uses, NOT XLA): the handler's 38 parameters must match the
    order and types of its binding, and every C-ABI call must match include/jic_b200.h."""
    res = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "ffi_mock"), "-I", CUDA_INC, FFI_SRC],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_jax_ffi_binding_fails_loudly_without_jax():
    from jaxincell_b200 import JicError, _jax_ffi
    try:
        import jax  # noqa: F401
        pytest.skip("jax present")
    except ImportError:
        pass
    with pytest.raises(JicError, match="needs jax"):
        _jax_ffi.register()


def test_jax_ffi_wrappers_marshal_the_reference_arguments(monkeypatch):
    """No jax here: a recording fake of `jax.ffi.ffi_call` checks that the two Python wrappers hand the handlers exactly the operands,
    result shapes and attribute names that csrc/jic_xla_ffi.cc binds (the names are read from the C++ source)."""
    import sys
    import types

    import numpy as np
    from jaxincell_b200 import _jax_ffi

    calls = []

    def ffi_call(target, results):
        def call(*operands, **attrs):
            calls.append((target, results, operands, attrs))
            return tuple(np.zeros(r.shape, r.dtype) for r in results)
        return call

    fake = types.ModuleType("jax")
    fake.ShapeDtypeStruct = lambda shape, dtype: types.SimpleNamespace(shape=tuple(shape), dtype=np.dtype(dtype))
    fake.ffi = types.SimpleNamespace(ffi_call=ffi_call)
    fake.numpy = types.ModuleType("jax.numpy")
    for name in ("asarray", "where", "float32", "uint8"):
        setattr(fake.numpy, name, getattr(np, name))
    monkeypatch.setitem(sys.modules, "jax", fake)
    monkeypatch.setitem(sys.modules, "jax.numpy", fake.numpy)
    monkeypatch.setattr(_jax_ffi, "_registered", True)

    src = open(FFI_SRC).read()

    def bound_attrs(symbol):
        body = src[src.index(f"    {symbol}, "):]
        body = body[:body.index("));") + 3]
        return re.findall(r'\.Attr<[^>]*>+\("(\w+)"\)', body), body.count(".Arg<"), body.count(".Ret<")

    N, G, T = 7, 5, 3
    species = [(4, -1.0, 2.0, -0.5), (3, 1.0, 9.0, 0.1)]
    solver = dict(filter_passes=5, filter_alpha=0.5, filter_strides=(1, 2, 4), relativistic=False, field_solver=0)
    grid = np.linspace(-0.4, 0.4, G)
    x = np.zeros((N, 3)); eE = np.zeros((G, 3), np.float32)
    hist, fields0, v_init = _jax_ffi.boris_run(x, x, eE, eE, species=species, n_steps=T, length=1.0, dx=0.2, dt=1e-10, grid=grid, solver=solver)
    target, results, operands, attrs = calls[-1]
    names, n_arg, n_ret = bound_attrs("jic_boris_run")
    assert target == "jic_boris_run" and list(attrs) == names and len(operands) == n_arg and len(results) == n_ret
    assert hist[0].shape == (T, N, 3) and hist[2].shape == (T, G, 3) and hist[5].shape == (T, G) and fields0[0].shape == (G, 3) and v_init.shape == (N, 3)

    q = np.array([-1.0] * 4 + [1.0] * 3).reshape(-1, 1)
    carry = (np.zeros((G, 3)), np.zeros((G, 3)), x, x, x, x, q, np.abs(q), q)
    ext = {"external_electric_field": eE, "external_magnetic_field": eE}
    new_carry, step_data = _jax_ffi.boris_step(carry, 0, solver, ext, 0.2, 1e-10, grid, (1.0, 1.0, 1.0), 0, 0, 0, 0, 0, species=species)
    target, results, operands, attrs = calls[-1]
    names, n_arg, n_ret = bound_attrs("jic_boris_step")
    assert target == "jic_boris_step" and list(attrs) == names and len(operands) == n_arg and len(results) == n_ret
    assert len(new_carry) == 9 and len(step_data) == 6 and new_carry[6].shape == q.shape
    assert not new_carry[6].any()  # the fake returned alive = 0 everywhere: every charge is zeroed
