"""The INDEXED engine's CUDA source executed on the CPU (tests/cuda_on_cpu): g++ compiles csrc/jic_device.cuh, jic_kernels.cuh and
jic_carry.cuh as host code against a fake cuda_runtime.h, and ONE emulated thread runs k_start, k_step, k_fields -- and the carry
loader k_load_carry / k_carry_fields -- in the order csrc/jic_engine.cu launches them.  The histories must match the golden vectors.

What this proves: the arithmetic and control flow of that source (gather, Boris, BCs, deposits, filter, Maxwell half steps, output
rows, carry reload) are right, without a GPU.  What it cannot see: races, memory spaces, launch configuration, the binned engine
(warp-level code).  The GPU tests remain the parity tests proper; this one guards the kernels while no GPU is at hand."""
import ctypes as C
import ctypes as C_
import glob
import os
import subprocess

import numpy as np
import pytest


def FUZZ(n):
    """Seeds per fuzz test; JIC_FUZZ_SCALE=10 runs ten times as many (a bug hunt, not the default suite)."""
    return int(n * float(os.environ.get("JIC_FUZZ_SCALE", "1")))

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU_DIR = os.path.join(HERE, "cuda_on_cpu")
CSRC = os.path.join(ROOT, "jax-in-cell_b200", "csrc")
KEYS = ("electric_field", "magnetic_field", "current_density", "charge_density", "positions", "velocities")
# explicit stepper without the per-step field_solver branch (that one needs k_gauss, a multi-CTA / warp-shuffle kernel)
GOLDEN = [f for f in sorted(glob.glob(os.path.join(HERE, "golden", "refsrc_*.npz"))) if "crank_nicolson" not in f and "field_solver" not in f]


class EmuParams(C.Structure):
    _fields_ = [("G", C.c_int), ("n_species", C.c_int), ("pbl", C.c_int), ("pbr", C.c_int), ("fbl", C.c_int), ("fbr", C.c_int),
                ("relativistic", C.c_int), ("unused", C.c_int), ("filter_passes", C.c_int), ("n_strides", C.c_int), ("strides", C.c_int * 8),
                ("L", C.c_double), ("Ly", C.c_double), ("Lz", C.c_double), ("dx", C.c_double), ("dt", C.c_double), ("grid_first", C.c_double),
                ("grid_last", C.c_double), ("filter_alpha", C.c_double), ("count", C.c_longlong * 8), ("q", C.c_double * 8), ("m", C.c_double * 8),
                ("qm", C.c_double * 8)]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libjic_emu.so")
    # hidden visibility + -Bsymbolic: the emulated kernels have the mangled names of the real library's launch stubs, and libjic_b200.so may
    # already be loaded RTLD_GLOBAL in this process -- the emulation must bind to its own definitions
    cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden", "-Wl,-Bsymbolic", "-I", os.path.join(EMU_DIR, "fake_cuda"), "-I", CSRC,
           os.path.join(EMU_DIR, "emulate_indexed.cpp"), "-o", so]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    lib = C.CDLL(so)
    lib.emu_run.restype = C.c_int
    for fn in (lib.emu_run, lib.emu_run_f32):
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(EmuParams)] + [C.c_void_p] * 4 + [C.c_int, C.c_int] + [C.c_void_p] * 8
    return lib


def run_emulated(lib, g, reload_at=-1, real=np.float64):
    G, T, N = int(g["G"]), int(g["T"]), len(g["x0"])
    ne, ni = int(g["n_e"]), int(g["n_i"])
    p = EmuParams()
    p.G, p.n_species = G, 2
    p.pbl, p.pbr, p.fbl, p.fbr = (int(b) for b in g["bcs"])
    p.relativistic, p.filter_passes, p.filter_alpha = int(g["relativistic"]), int(g["filter_passes"]), float(g["filter_alpha"])
    strides = [int(s) for s in g["filter_strides"]]
    p.n_strides = len(strides)
    for i, s in enumerate(strides):
        p.strides[i] = s
    L = float(g["length"])
    Ly, Lz = (float(b) for b in g["box_yz"]) if "box_yz" in g else (L, L)
    dx = L / G
    grid = np.linspace(-L / 2 + dx / 2, L / 2 - dx / 2, G)  # what _engine.make_params hands to jic_create
    p.L, p.Ly, p.Lz, p.dx, p.dt, p.grid_first, p.grid_last = L, Ly, Lz, dx, float(g["dt"]), float(grid[0]), float(grid[-1])
    for s, (n, o) in enumerate(((ne, 0), (ni, ne))):
        p.count[s], p.q[s], p.m[s], p.qm[s] = n, float(g["q"][o]), float(g["m"][o]), float(g["qm"][o])
    x0, v0 = np.ascontiguousarray(g["x0"], real), np.ascontiguousarray(g["v0"], real)
    eE, eB = np.ascontiguousarray(g["ext_E"], np.float32), np.ascontiguousarray(g["ext_B"], np.float32)
    out = dict(electric_field=np.zeros((T, G, 3), real), magnetic_field=np.zeros((T, G, 3), real), current_density=np.zeros((T, G, 3), real),
               charge_density=np.zeros((T, G), real), positions=np.zeros((T, N, 3), real), velocities=np.zeros((T, N, 3), real), E0=np.zeros((G, 3)),
               initial_velocities=np.zeros((N, 3), real))
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rc = (lib.emu_run if real == np.float64 else lib.emu_run_f32)(C.byref(p), ptr(x0), ptr(v0), ptr(eE), ptr(eB), T, reload_at, *[ptr(out[k]) for k in KEYS], ptr(out["E0"]), ptr(out["initial_velocities"]))
    assert rc == 0
    return out


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def fuzz_err(out, ref, k):
    """relerr, except that a self-consistent B at the round-off level of the E update (no transverse currents: 1e-10 of E/c and below --
    there even the two ORACLES differ by 1e-3 of it, seed 116 of the 3x fuzz) is measured against c B ~ 1e-4 E at least."""
    if k != "magnetic_field":
        return relerr(out[k], ref[k])
    floor = 1e-4 * np.abs(ref["electric_field"]).max() / 2.99792458e8
    return float(np.abs(out[k] - ref[k]).max() / max(np.abs(ref[k]).max(), floor))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(f)[:-4] for f in GOLDEN])
def test_indexed_engine_source_reproduces_the_reference_source_vectors(emu, path):
    g = dict(np.load(path))
    out = run_emulated(emu, g)
    for k in KEYS:
        assert relerr(out[k], g[k]) < 1e-9, k
    assert relerr(out["E0"], g["E0"]) < 1e-9
    assert relerr(out["initial_velocities"], g["initial_velocities"]) < 1e-14


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(f)[:-4] for f in GOLDEN])
@pytest.mark.parametrize("reload_at", [1, 4])
def test_carry_loader_source_continues_a_run(emu, path, reload_at):
    """jic_load_carry's kernels: after `reload_at` steps the state is wiped, the reference-shaped carry (E, B, x_{n-1/2}, x_n, x_{n+1/2}, v)
    goes back in through k_load_carry -> k_fields(init) -> k_carry_fields, and the run continues as if nothing had happened."""
    g = dict(np.load(path))
    straight = run_emulated(emu, g)
    reloaded = run_emulated(emu, g, reload_at=reload_at)
    for k in KEYS:
        assert relerr(reloaded[k], g[k]) < 1e-9, k
        assert relerr(reloaded[k], straight[k]) < 1e-12, k


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(f)[:-4] for f in GOLDEN])
def test_fp32_instantiation_within_the_fp32_tolerance(emu, path):
    """The float instantiation of the same source, carry reload included, within the 1e-3 the north star grants fp32 (fields 2e-3: the raw grid
    accumulates in float)."""
    g = dict(np.load(path))
    if float(g["dt"]) * 2.99792458e8 / (float(g["length"]) / int(g["G"])) > 1.0:
        pytest.skip("CFL > 1 in fp32: charge cancellation and cell-boundary decisions eat the 1e-3 budget (the GPU fp32 tests use CFL <= 1 as well)")
    out = run_emulated(emu, g, reload_at=3, real=np.float32)
    for k in KEYS:
        assert relerr(out[k].astype(np.float64), g[k]) < 2e-3, k


def _random_case(seed):
    from plasma import cfl_dt, two_species
    rng = np.random.default_rng(1000 + seed)
    G = int(rng.choice([3, 4, 5, 6, 7, 9, 16, 33]))
    length = float(rng.choice([0.01, 1.0]))
    bcs = tuple(int(b) for b in rng.integers(0, 3, 4))
    n_e, n_i = int(rng.integers(20, 90)), int(rng.integers(20, 90))
    p = two_species(n_e, n_i, length=length, G=G, seed=seed, vth_e=float(rng.choice([0.02, 0.1, 0.3])), vth_yz=float(rng.choice([0.0, 0.05])),
                    drift=float(rng.choice([0.0, 4e7])), plus_minus=bool(rng.integers(0, 2)), gpdl=0.6)
    passes = int(rng.choice([0, 1, 2, 5, 16, 17, 18, 25]))
    strides = tuple(int(s) for s in rng.choice([1, 2, 3, 4, 8], size=int(rng.integers(1, 4)), replace=False))
    box = (length * float(rng.choice([0.3, 1.0, 2.0])), length * float(rng.choice([0.5, 1.0])))
    ext = float(rng.choice([0.0, 1.0]))
    return dict(x0=p["x0"] * np.array([1.0, box[0] / length, box[1] / length]), v0=p["v0"], q=p["q"], m=p["m"], qm=p["qm"], n_e=n_e, n_i=n_i, length=length, G=G,
                dt=cfl_dt(length, G, float(rng.choice([0.5, 0.95, 2.5]))), T=int(rng.integers(3, 9)), bcs=np.array(bcs), filter_passes=passes,
                filter_alpha=float(rng.choice([0.3, 0.5, 0.8])), filter_strides=np.array(strides), relativistic=int(rng.integers(0, 2)), box_yz=np.array(box),
                ext_E=(ext * 1e3 * rng.standard_normal((G, 3))).astype(np.float32), ext_B=(ext * 1e-3 * rng.standard_normal((G, 3))).astype(np.float32))


# 68, 213 (and 32 below): G = 3, absorbing left wall, CFL 2.5 -- start-up positions x_{-1/2} parked on a cell border (deposit_jx_startup)
@pytest.mark.parametrize("seed", sorted(set(range(FUZZ(40))) | {68, 213}))
def test_indexed_engine_source_against_the_oracle_on_random_configurations(emu, seed):
    """Differential check over corners no fixture holds: every BC combination, grids from 3 cells, filter passes beyond the reference's cap
    of 17, strides larger than the grid, CFL 2.5 jumps, thin transverse boxes, relativistic or not, with and without external fields."""
    from oracle import closed_form as C
    g = _random_case(seed)
    pbl, pbr, fbl, fbr = (int(b) for b in g["bcs"])
    ref = C.run(g["x0"], g["v0"], g["q"], g["m"], g["qm"], length=g["length"], G=g["G"], dt=g["dt"], total_steps=g["T"], pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr,
                box_yz=tuple(g["box_yz"]), ext_E=g["ext_E"], ext_B=g["ext_B"],
                solver=dict(filter_passes=g["filter_passes"], filter_alpha=g["filter_alpha"], filter_strides=tuple(int(s) for s in g["filter_strides"]),
                            relativistic=bool(g["relativistic"])))
    out = run_emulated(emu, g, reload_at=2 if seed % 2 else -1)
    assert all(np.isfinite(ref[k]).all() for k in KEYS)
    if np.abs(ref["velocities"][..., 0]).max() * g["dt"] > g["length"]:
        # numerically unstable corners (few cells at CFL 2.5) blow the field up until particles cross more than a box length per step:
        # beyond the documented domain of the position-based "absorbed" test (known limits; seeds 181, 228 of the 6x fuzz)
        pytest.skip("particles move more than one box length per step")
    slack = None
    for k in KEYS:
        err = fuzz_err(out, ref, k)
        if err >= 1e-7 and slack is None:
            # a numerically unstable corner (few cells at CFL 2.5 amplify round-off by an order of magnitude per step)?  Then the two
            # oracles -- same semantics, different order of the arithmetic -- disagree as well, and nothing can be held tighter than that
            from oracle import literal as L
            lit = L.run(g["x0"], g["v0"], g["q"], g["m"], g["qm"], length=g["length"], G=g["G"], dt=g["dt"], total_steps=g["T"], pbl=pbl, pbr=pbr,
                        fbl=fbl, fbr=fbr, box_yz=tuple(g["box_yz"]), ext_E=g["ext_E"], ext_B=g["ext_B"],
                        solver=dict(filter_passes=g["filter_passes"], filter_alpha=g["filter_alpha"], filter_strides=tuple(int(s) for s in g["filter_strides"]),
                                    relativistic=bool(g["relativistic"])))
            slack = {kk: 20 * fuzz_err(lit, ref, kk) for kk in KEYS}
        assert err < 1e-7 + (slack[k] if slack else 0.0), (k, err, {kk: g[kk] for kk in ("G", "bcs", "filter_passes", "filter_strides", "relativistic", "T")})


# ---- warp-level kernels on the multi-threaded emulation (fake_cuda_mt): k_gauss (field_solver) and the Crank-Nicolson stepper ----------
@pytest.fixture(scope="module")
def emu_mt(tmp_path_factory):
    root = tmp_path_factory.mktemp("emu_mt")
    src = root / "x" / "csrc"
    src.mkdir(parents=True)
    (root / "include").mkdir()
    (root / "include" / "jic_b200.h").write_text(open(os.path.join(ROOT, "include", "jic_b200.h")).read())
    for name in ("jic_device.cuh", "jic_kernels.cuh", "jic_cn.cuh", "jic_cn_sorted.cuh", "jic_carry.cuh"):
        # the one textual change: dynamic shared arrays become plain externs (the harness defines them), so that `__shared__` can mean `static`
        (src / name).write_text(open(os.path.join(CSRC, name)).read().replace("extern __shared__", "extern"))
    so = str(root / "libjic_emu_mt.so")
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-fvisibility=hidden", "-Wl,-Bsymbolic", "-I",
           os.path.join(EMU_DIR, "fake_cuda_mt"), "-I", str(src), os.path.join(EMU_DIR, "emulate_mt.cpp"), "-o", so]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    lib = C.CDLL(so)
    lib.emu_fs_run.restype = C.c_int
    lib.emu_fs_run.argtypes = [C.POINTER(EmuParams), C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
    lib.emu_fs_run_chunked.restype = C.c_int
    lib.emu_fs_run_chunked.argtypes = lib.emu_fs_run.argtypes
    lib.emu_fs_run_two_ranks.restype = C.c_int
    lib.emu_fs_run_two_ranks.argtypes = [C.POINTER(EmuParams)] * 2 + [C.c_void_p] * 4 + [C.c_int, C.c_int] + [C.c_void_p] * 4
    lib.emu_cn_run.restype = C.c_int
    lib.emu_cn_run.argtypes = [C.POINTER(EmuParams), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int] + [C.c_void_p] * 7
    lib.emu_cn_sorted_run.restype = C.c_int
    lib.emu_cn_sorted_run.argtypes = lib.emu_cn_run.argtypes
    lib.emu_set_cn_sort_every.argtypes = [C.c_int]
    return lib


def _params_of(g):
    G = int(g["G"])
    ne, ni = int(g["n_e"]), int(g["n_i"])
    p = EmuParams()
    p.G, p.n_species = G, 2
    p.pbl, p.pbr, p.fbl, p.fbr = (int(b) for b in g["bcs"])
    p.relativistic, p.filter_passes, p.filter_alpha = int(g["relativistic"]), int(g["filter_passes"]), float(g["filter_alpha"])
    p.unused = int(g["field_solver"]) if "field_solver" in g else 0
    strides = [int(s) for s in g["filter_strides"]]
    p.n_strides = len(strides)
    for i, s in enumerate(strides):
        p.strides[i] = s
    L = float(g["length"])
    dx = L / G
    grid = np.linspace(-L / 2 + dx / 2, L / 2 - dx / 2, G)
    p.L, p.Ly, p.Lz, p.dx, p.dt, p.grid_first, p.grid_last = L, L, L, dx, float(g["dt"]), float(grid[0]), float(grid[-1])
    for s, (n, o) in enumerate(((ne, 0), (ni, ne))):
        p.count[s], p.q[s], p.m[s], p.qm[s] = n, float(g["q"][o]), float(g["m"][o]), float(g["qm"][o])
    return p


def _histories(g):
    G, T, N = int(g["G"]), int(g["T"]), len(g["x0"])
    return dict(electric_field=np.zeros((T, G, 3)), magnetic_field=np.zeros((T, G, 3)), current_density=np.zeros((T, G, 3)),
                charge_density=np.zeros((T, G)), positions=np.zeros((T, N, 3)), velocities=np.zeros((T, N, 3)))


FS_GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "refsrc_field_solver_*.npz")))
CN_GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "refsrc_crank_nicolson_*.npz")))


@pytest.mark.parametrize("reload_at", [-1, 3])
@pytest.mark.parametrize("path", FS_GOLDEN, ids=[os.path.basename(f)[:-4] for f in FS_GOLDEN])
def test_field_solver_source_on_the_threaded_emulation(emu_mt, path, reload_at):
    """k_step's face deposit, k_gauss_kernel, k_gauss (warp reductions, several CTAs) and k_fields' E_x replacement -- solvers 1, 2, 3, with
    and without walls -- against the reference-source vectors; with reload_at the carry loader is exercised on this branch too."""
    g = dict(np.load(path))
    out = _histories(g)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    x0, v0 = np.ascontiguousarray(g["x0"], np.float64), np.ascontiguousarray(g["v0"], np.float64)
    assert emu_mt.emu_fs_run(C.byref(_params_of(g)), ptr(x0), ptr(v0), int(g["T"]), reload_at, *[ptr(out[k]) for k in KEYS]) == 0
    for k in KEYS:
        assert relerr(out[k], g[k]) < 1e-9, k


@pytest.mark.parametrize("push", ["unsorted", "sorted"])
@pytest.mark.parametrize("reload_at", [-1, 3])
@pytest.mark.parametrize("path", CN_GOLDEN, ids=[os.path.basename(f)[:-4] for f in CN_GOLDEN])
def test_crank_nicolson_source_on_the_threaded_emulation(emu_mt, path, reload_at, push):
    """k_cn_start, k_cn_push, k_cn_fields (block reductions, device-side convergence flag), k_cn_record against the reference-source vectors,
    Picard iteration counts included; with reload_at the CN carry loader (k_cn_load, k_carry_copy_fields) continues a wiped run.
    push = sorted: the large-run variant (csrc/jic_cn_sorted.cuh: counting sort by cell every step, k_cn_push_sorted with its per-warp
    window, histories and exports through the permutation) on the same vectors."""
    g = dict(np.load(path))
    out = _histories(g)
    T = int(g["T"])
    picard = np.zeros(T, np.int64)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    x0, v0 = np.ascontiguousarray(g["x0"], np.float64), np.ascontiguousarray(g["v0"], np.float64)
    run = emu_mt.emu_cn_sorted_run if push == "sorted" else emu_mt.emu_cn_run
    emu_mt.emu_set_cn_sort_every(4 if reload_at > 0 else 1)  # (the engine's default is a sort every fourth step)
    rc = run(C.byref(_params_of(g)), ptr(x0), ptr(v0), T, int(g["cn_substeps"]), int(g["cn_max_iterations"]), float(g["cn_tolerance"]),
             reload_at, *[ptr(out[k]) for k in KEYS], ptr(picard))
    assert rc == 0
    for k in KEYS:
        assert relerr(out[k], g[k]) < 1e-9, k
    assert picard.tolist() == g["picard_iterations"].tolist()


@pytest.mark.parametrize("seed", sorted(set(range(FUZZ(24))) | {32}))
def test_field_solver_source_against_the_oracle_on_random_configurations(emu_mt, seed):
    """The corners of _random_case with a random field_solver on top (grids from 3 cells, every BC combination, CFL 2.5, filter passes beyond
    the cap), against the closed-form oracle; every other case also reloads its carry after two steps."""
    from oracle import closed_form as C
    g = _random_case(500 + seed)
    g["field_solver"] = 1 + seed % 3
    g["box_yz"] = np.array([g["length"], g["length"]])
    g["x0"] = np.clip(g["x0"], -g["length"] / 2, g["length"] / 2)
    pbl, pbr, fbl, fbr = (int(b) for b in g["bcs"])
    ref = C.run(g["x0"], g["v0"], g["q"], g["m"], g["qm"], length=g["length"], G=g["G"], dt=g["dt"], total_steps=g["T"], pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr,
                solver=dict(filter_passes=g["filter_passes"], filter_alpha=g["filter_alpha"], filter_strides=tuple(int(s) for s in g["filter_strides"]),
                            relativistic=bool(g["relativistic"]), field_solver=g["field_solver"]))
    if np.abs(ref["velocities"][..., 0]).max() * g["dt"] > g["length"]:
        # a numerically unstable corner (3 cells at CFL 2.5) whose field blows up until particles cross more than a box length per step:
        # beyond the documented domain of the position-based "absorbed" test (known limits; seed 42 of the 3x fuzz)
        pytest.skip("particles move more than one box length per step")
    out = _histories(g)
    ptr = lambda a: a.ctypes.data_as(C_.c_void_p)  # noqa: E731
    x0, v0 = np.ascontiguousarray(g["x0"], np.float64), np.ascontiguousarray(g["v0"], np.float64)
    assert emu_mt.emu_fs_run(C_.byref(_params_of(g)), ptr(x0), ptr(v0), int(g["T"]), 2 if seed % 2 else -1, *[ptr(out[k]) for k in KEYS]) == 0
    slack = None
    for k in KEYS:
        err = fuzz_err(out, ref, k)
        if err >= 1e-7 and slack is None:
            # fields at round-off level (a neutral start: E_x = solve of rho ~ 1e-11 V/m, seeds 119 and 134 of the 6x fuzz) or an unstable
            # corner: the two oracles disagree as well, and nothing can be held tighter than that
            from oracle import literal as L
            lit = L.run(g["x0"], g["v0"], g["q"], g["m"], g["qm"], length=g["length"], G=g["G"], dt=g["dt"], total_steps=g["T"], pbl=pbl, pbr=pbr, fbl=fbl,
                        fbr=fbr, solver=dict(filter_passes=g["filter_passes"], filter_alpha=g["filter_alpha"],
                                             filter_strides=tuple(int(s) for s in g["filter_strides"]), relativistic=bool(g["relativistic"]),
                                             field_solver=g["field_solver"]))
            slack = {kk: 20 * fuzz_err(lit, ref, kk) for kk in KEYS}
        assert err < 1e-7 + (slack[k] if slack else 0.0), (k, err, {kk: g[kk] for kk in ("G", "bcs", "filter_passes", "filter_strides", "relativistic", "T", "field_solver")})


def _two_rank_case(seed):
    from plasma import cfl_dt, two_species
    rng = np.random.default_rng(7000 + seed)
    G = int(rng.choice([5, 9, 12, 16]))
    # one periodic and one non-periodic particle wall (the face-fix configuration) twice out of three, anything else otherwise
    bcs = [(0, 1, 0, 1), (1, 0, 1, 0), (2, 0, 2, 0), (0, 2, 0, 2), (0, 0, 0, 0), (1, 2, 1, 2)][seed % 6]
    n_e, n_i = int(rng.integers(40, 120)), int(rng.integers(40, 120))
    length = 0.01
    p = two_species(n_e, n_i, length=length, G=G, seed=seed, vth_e=0.5, vth_yz=0.05, gpdl=0.6)
    return dict(x0=p["x0"], v0=p["v0"], q=p["q"], m=p["m"], qm=p["qm"], n_e=n_e, n_i=n_i, length=length, G=G, dt=cfl_dt(length, G, 2.5), T=5,
                bcs=np.array(bcs), filter_passes=2, filter_alpha=0.5, filter_strides=np.array([1, 2]), relativistic=0, field_solver=1 + seed % 3,
                species=p["species"])


@pytest.mark.parametrize("seed", range(8))
def test_start_up_kernels_chunk_by_chunk(emu_mt, seed):
    """jic_initialize_host enqueues k_start (and k_start_face_fix) once per uploaded chunk, on chunk-local pointers with a global offset:
    three chunks whose borders fall inside the species blocks give the histories of the oracle (face-fix configurations included)."""
    from oracle import closed_form as C
    g = _two_rank_case(seed)
    pbl, pbr, fbl, fbr = (int(b) for b in g["bcs"])
    ref = C.run(g["x0"], g["v0"], g["q"], g["m"], g["qm"], length=g["length"], G=g["G"], dt=g["dt"], total_steps=g["T"], pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr,
                solver=dict(filter_passes=2, filter_alpha=0.5, filter_strides=(1, 2), field_solver=g["field_solver"]))
    out = _histories(g)
    ptr = lambda a: a.ctypes.data_as(C_.c_void_p)  # noqa: E731
    x0, v0 = np.ascontiguousarray(g["x0"], np.float64), np.ascontiguousarray(g["v0"], np.float64)
    assert emu_mt.emu_fs_run_chunked(C_.byref(_params_of(g)), ptr(x0), ptr(v0), int(g["T"]), 3, *[ptr(out[k]) for k in KEYS]) == 0
    for k in KEYS:
        assert relerr(out[k], ref[k]) < 1e-7, (k, g["bcs"], g["field_solver"])


@pytest.mark.parametrize("seed", range(FUZZ(12)))
def test_field_solver_on_two_emulated_ranks(emu_mt, seed):
    """EngineT's multi-rank orchestration of the field_solver branch, restated on the emulation with the real kernels: particles sharded by
    jaxincell_b200.shard_particles, raw grids (face component included) summed before the field kernels, and -- the point -- the step-0
    face correction of k_start_face_fix, which the start-up reduction has already summed, kept on rank 0 only (EngineT::initialize_finish).
    Both ranks end with identical fields equal to the single-rank oracle; keeping the reduced correction on every rank (what the engine
    would do without that memset) is shown to be wrong whenever the correction is non-zero."""
    import importlib.util
    from oracle import closed_form as C
    spec = importlib.util.spec_from_file_location("jic_parallel", os.path.join(ROOT, "jax-in-cell_b200", "jaxincell_b200", "_parallel.py"))
    par = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(par)
    g = _two_rank_case(seed)
    pbl, pbr, fbl, fbr = (int(b) for b in g["bcs"])
    ref = C.run(g["x0"], g["v0"], g["q"], g["m"], g["qm"], length=g["length"], G=g["G"], dt=g["dt"], total_steps=g["T"], pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr,
                keep_particles=False, solver=dict(filter_passes=2, filter_alpha=0.5, filter_strides=(1, 2), field_solver=g["field_solver"]))
    ptr = lambda a: a.ctypes.data_as(C_.c_void_p)  # noqa: E731
    shards, params = [], []
    for rank in range(2):
        x, v, _ = par.shard_particles(g["x0"], g["v0"], g["species"], rank, 2)
        counts = [s["count"] for s in par.shard_species(g["species"], rank, 2)]
        local = dict(g, n_e=counts[0], n_i=counts[1], q=np.array([g["q"][0]] * counts[0] + [g["q"][-1]] * counts[1]),
                     m=np.array([g["m"][0]] * counts[0] + [g["m"][-1]] * counts[1]), qm=np.array([g["qm"][0]] * counts[0] + [g["qm"][-1]] * counts[1]))
        params.append(_params_of(local))
        shards.append((np.ascontiguousarray(x, np.float64), np.ascontiguousarray(v, np.float64)))
    fields = ("electric_field", "magnetic_field", "current_density", "charge_density")

    def run(keep_on_all_ranks):
        out = {k: v for k, v in _histories(g).items() if k in fields}
        rc = emu_mt.emu_fs_run_two_ranks(C_.byref(params[0]), C_.byref(params[1]), ptr(shards[0][0]), ptr(shards[0][1]), ptr(shards[1][0]), ptr(shards[1][1]),
                                         int(g["T"]), keep_on_all_ranks, *[ptr(out[k]) for k in fields])
        return rc, out

    rc, out = run(0)
    assert rc == 0
    for k in fields:
        assert relerr(out[k], ref[k]) < 1e-7, (k, g["bcs"], g["field_solver"])
    if pbl != pbr and 0 in (pbl, pbr):
        rc, wrong = run(1)
        # (rc == -2: the ranks disagree; otherwise they agree on a wrong E_x)  -- unless no particle needed the correction
        xp = g["x0"][:, 0] + g["dt"] / 2 * g["v0"][:, 0]
        crossed_periodic = (xp < -g["length"] / 2) if pbl == 0 else (xp > g["length"] / 2)
        if crossed_periodic.any():
            assert rc == -2 or relerr(wrong["electric_field"], ref["electric_field"]) > 1e-6


@pytest.mark.parametrize("push", ["unsorted", "sorted"])
@pytest.mark.parametrize("seed", range(FUZZ(24)))
def test_crank_nicolson_source_against_the_oracle_on_random_configurations(emu_mt, seed, push):
    """The implicit stepper's source over random boundary combinations, grids from 3 cells, 1-3 sub-steps, tight and loose Picard tolerances,
    against oracle/literal.py (pinned to the reference's CN_step by the refsrc vectors); odd seeds reload the CN carry after two steps.
    Both pushes: on these small grids the sorted one's warps span the whole box, so its window and its direct path both see traffic."""
    from oracle import literal as L
    g = _random_case(900 + seed)
    rng = np.random.default_rng(seed)
    n_sub, max_iter, tol = int(rng.integers(1, 4)), int(rng.choice([3, 8, 20])), float(rng.choice([1e-4, 1e-9, 1e-30]))
    g["dt"] = g["dt"] * 0.3  # (the implicit examples of the reference run at a fraction of the light CFL)
    g["box_yz"] = np.array([g["length"], g["length"]])
    g["x0"] = np.clip(g["x0"], -g["length"] / 2, g["length"] / 2)
    pbl, pbr, fbl, fbr = (int(b) for b in g["bcs"])
    ref = L.run_CN(g["x0"], g["v0"], g["q"], g["m"], g["qm"], length=g["length"], G=g["G"], dt=g["dt"], total_steps=g["T"], pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr,
                   solver={"max_number_of_Picard_iterations_implicit_CN": max_iter, "number_of_particle_substeps_implicit_CN": n_sub,
                           "tolerance_Picard_iterations_implicit_CN": tol,  # (the filter only enters the initial Gauss solve, _state_initialization.py:371-374)
                           "filter_passes": g["filter_passes"], "filter_alpha": g["filter_alpha"], "filter_strides": tuple(int(s) for s in g["filter_strides"])})
    out = _histories(g)
    T = int(g["T"])
    picard = np.zeros(T, np.int64)
    ptr = lambda a: a.ctypes.data_as(C_.c_void_p)  # noqa: E731
    x0, v0 = np.ascontiguousarray(g["x0"], np.float64), np.ascontiguousarray(g["v0"], np.float64)
    run = emu_mt.emu_cn_sorted_run if push == "sorted" else emu_mt.emu_cn_run
    emu_mt.emu_set_cn_sort_every(1 + seed % 4)
    assert run(C_.byref(_params_of(g)), ptr(x0), ptr(v0), T, n_sub, max_iter, tol, 2 if seed % 2 else -1, *[ptr(out[k]) for k in KEYS], ptr(picard)) == 0
    assert all(np.isfinite(ref[k]).all() for k in KEYS)
    info = {kk: g[kk] for kk in ("G", "bcs", "T")} | dict(n_sub=n_sub, max_iter=max_iter, tol=tol, push=push)
    for k in KEYS:
        assert relerr(out[k], ref[k]) < 1e-7, (k, info)
    # the Picard loop ends on `delta > tol`; with tolerances at round-off level the last iteration is decided by the summation order
    # (tol = 1e-30 means "until the iteration reproduces itself bit for bit": the count is then a property of the rounding, not of the scheme)
    if tol >= 1e-9:
        assert np.abs(picard - ref["picard_iterations"]).max() <= (0 if tol >= 1e-4 else 1), info
