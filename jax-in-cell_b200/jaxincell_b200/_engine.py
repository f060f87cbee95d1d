"""Thin host wrapper around the C ABI: torch supplies device memory, streams and (for N>1) the rendezvous only."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import JicError, Outputs, Params, Species

_ENGINES = {"indexed": _lib.ENGINE_INDEXED, "binned": _lib.ENGINE_BINNED}
_DEPOSITS = {"auto": _lib.DEPOSIT_AUTO, "global": _lib.DEPOSIT_GLOBAL_ATOMICS, "shared": _lib.DEPOSIT_SHARED_GRID}
_HIST = ("electric_field", "magnetic_field", "current_density", "charge_density", "positions", "velocities", "kinetic_energy")


def make_params(*, length, G, dt, n_species, length_y=0.0, length_z=0.0, pbl=0, pbr=0, fbl=0, fbr=0, filter_passes=5,
                filter_alpha=0.5, filter_strides=(1, 2, 4), relativistic=False, dtype=torch.float64, engine="indexed",
                track_yz=False, deposit="auto", steps_per_graph=0, device=-1, field_solver=0,
                time_evolution_algorithm=0, cn_substeps=2, cn_max_iterations=20, cn_tolerance=1e-6):
    """Fill a jic_params exactly as build_domain_state does (jaxincell/_state_initialization.py:27-49)."""
    p = Params()
    p.struct_bytes = C.sizeof(Params)
    p.dtype = _lib.F64 if dtype == torch.float64 else _lib.F32
    p.engine = _ENGINES[engine]
    p.device = device
    p.n_grid = int(G)
    p.n_species = int(n_species)
    dx = float(length) / int(G)
    grid = np.linspace(-float(length) / 2 + dx / 2, float(length) / 2 - dx / 2, int(G))
    p.length, p.length_y, p.length_z = float(length), float(length_y or 0.0), float(length_z or 0.0)
    p.dx, p.dt = dx, float(dt)
    p.grid_first, p.grid_last = float(grid[0]), float(grid[-1])
    p.particle_bc_left, p.particle_bc_right, p.field_bc_left, p.field_bc_right = int(pbl), int(pbr), int(fbl), int(fbr)
    strides = tuple(int(s) for s in filter_strides)
    if len(strides) > _lib.JIC_MAX_STRIDES:
        raise JicError(f"at most {_lib.JIC_MAX_STRIDES} filter strides")
    p.filter_passes, p.n_filter_strides, p.filter_alpha = int(filter_passes), len(strides), float(filter_alpha)
    for i, s in enumerate(strides):
        p.filter_strides[i] = s
    p.relativistic, p.track_yz = int(bool(relativistic)), int(bool(track_yz))
    p.deposit, p.steps_per_graph = _DEPOSITS[deposit], int(steps_per_graph)
    p.field_solver = int(field_solver)  # Boris_step's `field_solver` argument (jaxincell/_algorithms.py:20,69-78)
    # CN_step (jaxincell/_algorithms.py:100-241); defaults of _parameters/_solver_parameters.py:15-17
    p.time_evolution_algorithm, p.cn_substeps, p.cn_max_iterations = int(time_evolution_algorithm), int(cn_substeps), int(cn_max_iterations)
    p.cn_tolerance = float(cn_tolerance)
    return p, grid


def make_species(species):
    arr = (Species * len(species))()
    for i, s in enumerate(species):
        arr[i].count, arr[i].charge, arr[i].mass, arr[i].charge_to_mass = int(s["count"]), float(s["q"]), float(s["m"]), float(s["qm"])
    return arr


class HotPath:
    """One device context: the carry of jaxincell/_simulation.py:228-231 and the scan of :253, on one GPU."""

    def __init__(self, *, species, dtype=torch.float64, device=None, **kw):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise JicError("no CUDA device: jaxincell_b200 has no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = dtype
        self.params, self.grid = make_params(n_species=len(species), dtype=dtype, device=self.device.index or 0, **kw)
        self.species = make_species(species)
        self.N = int(sum(int(s["count"]) for s in species))
        self.G = self.params.n_grid
        self.n_species = len(species)
        self.ctx = C.c_void_p()
        _lib.check(self.lib.jic_create(C.byref(self.params), self.species, C.byref(self.ctx)))
        self._keep = []

    # ---- helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, rc):
        _lib.check(rc, self.ctx)

    def _ptr(self, t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()

    def _dev(self, a, dtype=None):
        t = torch.as_tensor(a)
        return t.to(device=self.device, dtype=dtype or self.dtype).contiguous()

    # ---- multi-GPU rendezvous (torch.distributed is only the courier of the 128-byte NCCL id)
    def comm_init_from_torch(self):
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        if world == 1:
            return
        from ._parallel import broadcast_bytes
        buf = C.create_string_buffer(128)
        if rank == 0:
            _lib.check(self.lib.jic_comm_unique_id(buf))
        raw = broadcast_bytes(buf.raw if rank == 0 else None, 128, 0, self.device)
        self._chk(self.lib.jic_comm_init(self.ctx, C.c_char_p(raw), rank, world))

    # ---- state in
    def set_external_fields(self, ext_E=None, ext_B=None):
        e = None if ext_E is None else self._dev(ext_E, torch.float32)
        b = None if ext_B is None else self._dev(ext_B, torch.float32)
        for name, a in (("ext_E", e), ("ext_B", b)):  # the library reads exactly G * 3 floats from each
            if a is not None and tuple(a.shape) != (self.G, 3):
                raise JicError(f"{name} must have shape ({self.G}, 3), got {tuple(a.shape)}")
        self._chk(self.lib.jic_set_external_fields(self.ctx, self._ptr(e), self._ptr(b), self._stream()))
        self._keep = [e, b]

    def initialize(self, x0, v0):
        x0, v0 = self._dev(x0), self._dev(v0)
        if tuple(x0.shape) != (self.N, 3) or tuple(v0.shape) != (self.N, 3):
            raise JicError(f"x0/v0 must have shape ({self.N}, 3)")
        self._chk(self.lib.jic_initialize(self.ctx, self._ptr(x0), self._ptr(v0), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()  # x0/v0 may be freed by the caller afterwards

    def load_carry(self, E, B, x_minus_half, x_n, x_plus_half, v_n):
        """jic_load_carry: the reference's scan carry (jaxincell/_simulation.py:228-231) instead of (x0, v0); `run(1)` is then exactly one
        Boris_step.  E, B (G,3); the particle arrays (N,3) in species-table order.  INDEXED engine."""
        E, B = self._dev(E), self._dev(B)
        parts = [self._dev(a) for a in (x_minus_half, x_n, x_plus_half, v_n)]
        if tuple(E.shape) != (self.G, 3) or tuple(B.shape) != (self.G, 3):
            raise JicError(f"E/B must have shape ({self.G}, 3)")
        if any(tuple(a.shape) != (self.N, 3) for a in parts):
            raise JicError(f"particle arrays of the carry must have shape ({self.N}, 3)")
        self._chk(self.lib.jic_load_carry(self.ctx, self._ptr(E), self._ptr(B), *[self._ptr(a) for a in parts], self._stream()))
        torch.cuda.current_stream(self.device).synchronize()  # the inputs may be freed by the caller afterwards

    def load_carry_cn(self, E, B, x_n, v_n, alive=None):
        """jic_load_carry_cn: the carry of CN_step (jaxincell/_simulation.py:237-240); `alive` (N,) is False where the carry's charge is 0."""
        E, B, x_n, v_n = self._dev(E), self._dev(B), self._dev(x_n), self._dev(v_n)
        a = None if alive is None else self._dev(np.asarray(alive).reshape(-1) != 0, torch.uint8)
        if tuple(E.shape) != (self.G, 3) or tuple(B.shape) != (self.G, 3) or tuple(x_n.shape) != (self.N, 3) or tuple(v_n.shape) != (self.N, 3):
            raise JicError(f"the CN carry is E, B ({self.G}, 3) and x, v ({self.N}, 3)")
        if a is not None and tuple(a.shape) != (self.N,):
            raise JicError(f"alive must have {self.N} entries")
        self._chk(self.lib.jic_load_carry_cn(self.ctx, self._ptr(E), self._ptr(B), self._ptr(x_n), self._ptr(v_n), self._ptr(a), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()

    def initialize_host(self, x0, v0):
        """x0, v0: HOST tensors / arrays (N,3), ideally pinned: jic_initialize_host uploads them in chunks overlapped with the
        start-up kernels (the device never holds a full copy).  Synchronises before returning."""
        x0 = torch.as_tensor(x0).to(dtype=self.dtype).contiguous()
        v0 = torch.as_tensor(v0).to(dtype=self.dtype).contiguous()
        if x0.is_cuda or v0.is_cuda:
            raise JicError("initialize_host takes host memory; use initialize() for device tensors")
        if tuple(x0.shape) != (self.N, 3) or tuple(v0.shape) != (self.N, 3):
            raise JicError(f"x0/v0 must have shape ({self.N}, 3)")
        with torch.cuda.device(self.device):
            self._chk(self.lib.jic_initialize_host(self.ctx, self._ptr(x0), self._ptr(v0), self._stream()))
            torch.cuda.current_stream(self.device).synchronize()

    # ---- stepping
    def alloc_outputs(self, n_steps, fields=True, particles=False, kinetic=False):
        """History tensors for `run`.  kinetic=True: "kinetic_energy" (T, n_species) float64 -- the per-species kinetic energy of every
        step, reduced on the device (one more pass over the velocities per step), for runs too large for the velocity history."""
        out = {}
        T, G, N = int(n_steps), self.G, self.N
        if kinetic:
            out["kinetic_energy"] = torch.zeros((T, self.n_species), dtype=torch.float64, device=self.device)
        if fields:
            for k in _HIST[:3]:
                out[k] = torch.empty((T, G, 3), dtype=self.dtype, device=self.device)
            out["charge_density"] = torch.empty((T, G), dtype=self.dtype, device=self.device)
        if particles:
            out["positions"] = torch.empty((T, N, 3), dtype=self.dtype, device=self.device)
            out["velocities"] = torch.empty((T, N, 3), dtype=self.dtype, device=self.device)
        return out

    def run(self, n_steps, outputs=None, fields=True, particles=False, kinetic=False):
        """Advance n_steps; returns the dict of history tensors (allocated here unless `outputs` is given)."""
        if outputs is None:
            outputs = self.alloc_outputs(n_steps, fields, particles, kinetic)
        o = Outputs()
        T = int(n_steps)
        for k in _HIST:
            if k in outputs:  # the library writes T rows of the full shape: check before handing over a raw pointer
                t = outputs[k]
                want = {"charge_density": (T, self.G), "positions": (T, self.N, 3), "velocities": (T, self.N, 3),
                        "kinetic_energy": (T, self.n_species)}.get(k, (T, self.G, 3))
                wdt = torch.float64 if k == "kinetic_energy" else self.dtype
                if tuple(t.shape) != want or t.dtype != wdt or not t.is_contiguous() or t.device != self.device:
                    raise JicError(f"output buffer '{k}' must be a contiguous {wdt} tensor of shape {want} on {self.device}")
            setattr(o, k, outputs[k].data_ptr() if k in outputs else None)
        self._chk(self.lib.jic_run(self.ctx, T, C.byref(o), self._stream()))
        return outputs

    # ---- state out
    def fields(self):
        G = self.G
        E, B, J = (torch.empty((G, 3), dtype=self.dtype, device=self.device) for _ in range(3))
        rho = torch.empty((G,), dtype=self.dtype, device=self.device)
        self._chk(self.lib.jic_get_fields(self.ctx, self._ptr(E), self._ptr(B), self._ptr(J), self._ptr(rho), self._stream()))
        return E, B, J, rho

    def initial(self, velocities=True):
        G = self.G
        E0, B0 = (torch.empty((G, 3), dtype=self.dtype, device=self.device) for _ in range(2))
        v = torch.empty((self.N, 3), dtype=self.dtype, device=self.device) if velocities else None
        self._chk(self.lib.jic_get_initial(self.ctx, self._ptr(E0), self._ptr(B0), self._ptr(v), self._stream()))
        return E0, B0, v

    def particles(self):
        x = torch.empty((self.N, 3), dtype=self.dtype, device=self.device)
        v = torch.empty((self.N, 3), dtype=self.dtype, device=self.device)
        alive = torch.empty((self.N,), dtype=torch.uint8, device=self.device)
        self._chk(self.lib.jic_get_particles(self.ctx, self._ptr(x), self._ptr(v), self._ptr(alive), self._stream()))
        return x, v, alive

    def kinetic_energy(self):
        ke = torch.zeros((1,), dtype=torch.float64, device=self.device)
        self._chk(self.lib.jic_kinetic_energy(self.ctx, self._ptr(ke), self._stream()))
        return ke

    def check_status(self):
        """Synchronise and raise JicError if a kernel set a sticky error flag (store capacity exhausted: particles dropped; a peer
        rank missed the fused reduction).  `run` only enqueues: call this before trusting the histories of a single `run`."""
        self._chk(self.lib.jic_check_status(self.ctx, self._stream()))

    def store_stats(self):
        """Counters of the binned particle store (jic_store_stats): dict with items, overflow (per buffer), error, general (entries of
        the last step's general-path list), slots_used, slots_per_buffer, absorbed."""
        a = (C.c_int64 * 8)()
        self._chk(self.lib.jic_store_stats(self.ctx, a, self._stream()))
        if self.params.time_evolution_algorithm == 1:   # no binned store: which Crank-Nicolson push the context runs
            return dict(cn_sorted=int(a[0]))
        return dict(items=a[0], overflow=(a[1], a[2]), error=a[3], general=a[4], slots_used=a[5], slots_per_buffer=a[6], absorbed=a[7])

    def profile_steps(self, n_steps):
        """(ms in the particle kernels, ms in all-reduce + field kernel) summed over n_steps real steps (CUDA events)."""
        a, b = C.c_double(0.0), C.c_double(0.0)
        self._chk(self.lib.jic_profile_steps(self.ctx, int(n_steps), C.byref(a), C.byref(b), self._stream()))
        return a.value, b.value

    def push_kernel_time(self, reset=True):
        """(summed device-side ms of the binned push kernel, number of launches) since the last reset -- %globaltimer stamps taken
        by the kernel itself, valid inside graph replays."""
        a, n = C.c_double(0.0), C.c_int64(0)
        self._chk(self.lib.jic_push_kernel_time(self.ctx, C.byref(a), C.byref(n), int(bool(reset)), self._stream()))
        return a.value, n.value

    def picard_iterations(self):
        """Crank-Nicolson: (Picard iterations of the last step, of all steps so far)."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._chk(self.lib.jic_get_picard_iterations(self.ctx, C.byref(a), C.byref(b), self._stream()))
        return a.value, b.value

    def comm_mode(self):
        """"single" | "nccl" (all-reduce in front of the field kernel) | "fused" (peer-memory reduction inside the field kernel)."""
        return ("single", "nccl", "fused")[int(self.lib.jic_comm_mode(self.ctx))]

    def launch_count(self):
        return int(self.lib.jic_launch_count(self.ctx))

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx.value:
            self.lib.jic_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def trim_memory():
    """Hand the device memory that the library's pools keep for the next context back to the driver (jic_trim_memory)."""
    _lib.load().jic_trim_memory()


def sample_particles(species_sampling, box_size, *, dtype=torch.float64, device=None, threefry_partitionable=True, rank=0, world=1):
    """jic_sample_particles[_slice]: initial (x0, v0) device tensors from the reference's formulas and jax.random's streams
    (jaxincell/_state_initialization.py:51-85).  `species_sampling`: dicts with count, seed_position, seed_velocity and per-axis
    triples random_positions, velocity_plus_minus, perturbation_amplitude, perturbation_wavenumber, vth_over_c, drift_speed.
    world > 1: only this rank's index slice of every species (`_parallel.shard_counts`), generated in place -- no scatter."""
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise JicError("no CUDA device: jaxincell_b200 has no CPU path")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    arr = (_lib.SpeciesSampling * len(species_sampling))()
    for i, s in enumerate(species_sampling):
        arr[i].count, arr[i].seed_position, arr[i].seed_velocity = int(s["count"]), int(s["seed_position"]), int(s["seed_velocity"])
        for a in range(3):
            arr[i].random_positions[a] = int(bool(s["random_positions"][a]))
            arr[i].velocity_plus_minus[a] = int(bool(s["velocity_plus_minus"][a]))
            arr[i].perturbation_amplitude[a] = float(s["perturbation_amplitude"][a])
            arr[i].perturbation_wavenumber[a] = float(s["perturbation_wavenumber"][a])
            arr[i].vth_over_c[a] = float(s["vth_over_c"][a])
            arr[i].drift_speed[a] = float(s["drift_speed"][a])
    from ._parallel import shard_counts
    firsts, locals_ = [], []
    for s in species_sampling:
        counts = shard_counts(int(s["count"]), world)
        firsts.append(sum(counts[:rank])); locals_.append(counts[rank])
    first = (C.c_int64 * len(firsts))(*firsts)
    local = (C.c_int64 * len(locals_))(*locals_)
    N = int(sum(locals_))
    x0 = torch.empty((N, 3), dtype=dtype, device=device)
    v0 = torch.empty((N, 3), dtype=dtype, device=device)
    box = (C.c_double * 3)(*[float(b) for b in box_size])
    with torch.cuda.device(device):
        _lib.check(lib.jic_sample_particles_slice(_lib.F64 if dtype == torch.float64 else _lib.F32, len(species_sampling), arr, first, local, box,
                                                  int(bool(threefry_partitionable)), C.c_void_p(x0.data_ptr()), C.c_void_p(v0.data_ptr()),
                                                  C.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
    return x0, v0


def simulate_host(*, species, x0, v0, n_steps, ext_E=None, ext_B=None, dtype=np.float64, fields=True, particles=False,
                  initial=False, out=None, **kw):
    """jic_simulate_host: HOST (NumPy) buffers in and out, every host<->device copy inside the call."""
    lib = _lib.load()
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    if particles:
        kw.setdefault("track_yz", True)
    kinetic = bool(kw.pop("kinetic", False))
    params, grid = make_params(n_species=len(species), dtype=tdt, **kw)
    sp = make_species(species)
    N = int(sum(int(s["count"]) for s in species))
    G, T = params.n_grid, int(n_steps)
    x0 = np.ascontiguousarray(x0, dtype=dtype)
    v0 = np.ascontiguousarray(v0, dtype=dtype)
    assert x0.shape == (N, 3) and v0.shape == (N, 3)
    res = {} if out is None else out
    if fields:
        for k in _HIST[:3]:
            res.setdefault(k, np.empty((T, G, 3), dtype=dtype))
        res.setdefault("charge_density", np.empty((T, G), dtype=dtype))
    if particles:
        res.setdefault("positions", np.empty((T, N, 3), dtype=dtype))
        res.setdefault("velocities", np.empty((T, N, 3), dtype=dtype))
    if kinetic:
        res.setdefault("kinetic_energy", np.zeros((T, len(species)), dtype=np.float64))
    o = Outputs()
    for k in _HIST:
        setattr(o, k, res[k].ctypes.data if k in res else None)
    eE = None if ext_E is None else np.ascontiguousarray(ext_E, dtype=np.float32)
    eB = None if ext_B is None else np.ascontiguousarray(ext_B, dtype=np.float32)
    for name, a in (("ext_E", eE), ("ext_B", eB)):  # the library reads exactly G * 3 floats from each
        if a is not None and a.shape != (G, 3):
            raise JicError(f"{name} must have shape ({G}, 3), got {a.shape}")
    for k in _HIST:  # caller-provided output buffers: the library writes T rows of the full shape into them
        if k in res:
            want = {"charge_density": (T, G), "positions": (T, N, 3), "velocities": (T, N, 3), "kinetic_energy": (T, len(species))}.get(k, (T, G, 3))
            wdt = np.float64 if k == "kinetic_energy" else np.dtype(dtype)
            if not isinstance(res[k], np.ndarray) or res[k].shape != want or res[k].dtype != wdt or not res[k].flags["C_CONTIGUOUS"]:
                raise JicError(f"output buffer '{k}' must be a C-contiguous {np.dtype(wdt).name} array of shape {want}")
    E0 = B0 = vi = None
    if initial:
        E0, B0 = np.empty((G, 3), dtype=dtype), np.empty((G, 3), dtype=dtype)
        # post-BC initial velocities: kept by the INDEXED store and by the Crank-Nicolson stepper (whatever the engine string says);
        # the binned store does not keep particle order
        if kw.get("engine", "indexed") == "indexed" or int(kw.get("time_evolution_algorithm", 0)) == 1:
            vi = np.empty((N, 3), dtype=dtype)
    vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p()
    _lib.check(lib.jic_simulate_host(C.byref(params), sp, vp(x0), vp(v0), vp(eE), vp(eB), T, C.byref(o), vp(E0), vp(B0), vp(vi)))
    res["grid"] = grid
    if initial:
        res["fields"] = (E0, B0)
        if vi is not None:
            res["initial_velocities"] = vi
    return res
