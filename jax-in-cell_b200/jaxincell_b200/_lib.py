"""ctypes binding of include/jic_b200.h.  There is no CPU fallback: a missing or stale library is an error."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JIC_B200_LIB") or os.path.join(_HERE, "libjic_b200.so")  # override: tuning variants of the same ABI

JIC_MAX_SPECIES = 8
JIC_MAX_STRIDES = 8
JIC_ABI_VERSION = 1
F64, F32 = 0, 1
ENGINE_INDEXED, ENGINE_BINNED = 0, 1
DEPOSIT_AUTO, DEPOSIT_GLOBAL_ATOMICS, DEPOSIT_SHARED_GRID = 0, 1, 2


class JicError(RuntimeError):
    pass


class Species(C.Structure):
    _fields_ = [("count", C.c_int64), ("charge", C.c_double), ("mass", C.c_double), ("charge_to_mass", C.c_double)]


class Params(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_uint32), ("dtype", C.c_int32), ("engine", C.c_int32), ("device", C.c_int32),
        ("n_grid", C.c_int32), ("n_species", C.c_int32),
        ("length", C.c_double), ("length_y", C.c_double), ("length_z", C.c_double),
        ("dx", C.c_double), ("dt", C.c_double), ("grid_first", C.c_double), ("grid_last", C.c_double),
        ("particle_bc_left", C.c_int32), ("particle_bc_right", C.c_int32), ("field_bc_left", C.c_int32), ("field_bc_right", C.c_int32),
        ("filter_passes", C.c_int32), ("n_filter_strides", C.c_int32), ("filter_strides", C.c_int32 * JIC_MAX_STRIDES),
        ("filter_alpha", C.c_double),
        ("relativistic", C.c_int32), ("track_yz", C.c_int32), ("deposit", C.c_int32), ("steps_per_graph", C.c_int32),
        ("field_solver", C.c_int32), ("time_evolution_algorithm", C.c_int32), ("cn_substeps", C.c_int32), ("cn_max_iterations", C.c_int32),
        ("cn_tolerance", C.c_double), ("reserved", C.c_int32 * 2),
    ]


class SpeciesSampling(C.Structure):
    _fields_ = [("count", C.c_int64), ("seed_position", C.c_int64), ("seed_velocity", C.c_int64),
                ("random_positions", C.c_int32 * 3), ("velocity_plus_minus", C.c_int32 * 3),
                ("perturbation_amplitude", C.c_double * 3), ("perturbation_wavenumber", C.c_double * 3),
                ("vth_over_c", C.c_double * 3), ("drift_speed", C.c_double * 3)]


class Outputs(C.Structure):
    _fields_ = [("electric_field", C.c_void_p), ("magnetic_field", C.c_void_p), ("current_density", C.c_void_p),
                ("charge_density", C.c_void_p), ("positions", C.c_void_p), ("velocities", C.c_void_p), ("kinetic_energy", C.c_void_p)]


# every symbol include/jic_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("jic_abi_version", C.c_int, []),
    ("jic_last_error", C.c_char_p, [_P]),
    ("jic_create", C.c_int, [C.POINTER(Params), C.POINTER(Species), C.POINTER(_P)]),
    ("jic_destroy", C.c_int, [_P]),
    ("jic_comm_unique_id", C.c_int, [_P]),
    ("jic_comm_init", C.c_int, [_P, _P, C.c_int, C.c_int]),
    ("jic_comm_mode", C.c_int, [_P]),
    ("jic_set_external_fields", C.c_int, [_P, _P, _P, _P]),
    ("jic_initialize", C.c_int, [_P, _P, _P, _P]),
    ("jic_initialize_host", C.c_int, [_P, _P, _P, _P]),
    ("jic_load_carry_cn", C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    ("jic_load_carry", C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    ("jic_run", C.c_int, [_P, C.c_int64, C.POINTER(Outputs), _P]),
    ("jic_get_fields", C.c_int, [_P, _P, _P, _P, _P, _P]),
    ("jic_get_initial", C.c_int, [_P, _P, _P, _P, _P]),
    ("jic_get_particles", C.c_int, [_P, _P, _P, _P, _P]),
    ("jic_kinetic_energy", C.c_int, [_P, _P, _P]),
    ("jic_profile_steps", C.c_int, [_P, C.c_int64, _P, _P, _P]),
    ("jic_get_picard_iterations", C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _P]),
    ("jic_check_status", C.c_int, [_P, _P]),
    ("jic_push_kernel_time", C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32, _P]),
    ("jic_store_stats", C.c_int, [_P, C.POINTER(C.c_int64), _P]),
    ("jic_trim_memory", None, []),
    ("jic_launch_count", C.c_int64, [_P]),
    ("jic_sample_particles", C.c_int, [C.c_int32, C.c_int32, C.POINTER(SpeciesSampling), C.POINTER(C.c_double), C.c_int32, _P, _P, _P]),
    ("jic_sample_particles_slice", C.c_int, [C.c_int32, C.c_int32, C.POINTER(SpeciesSampling), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                            C.POINTER(C.c_double), C.c_int32, _P, _P, _P]),
    ("jic_simulate_host", C.c_int, [C.POINTER(Params), C.POINTER(Species), _P, _P, _P, _P, C.c_int64, C.POINTER(Outputs), _P, _P, _P]),
]

_lib = None


def load():
    """Load libjic_b200.so (built by ``jax-in-cell_b200/build.py`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise JicError(f"{LIB_PATH} is missing: run `python jax-in-cell_b200/build.py` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.jic_abi_version() != JIC_ABI_VERSION:
        raise JicError(f"libjic_b200 ABI {lib.jic_abi_version()} != binding ABI {JIC_ABI_VERSION}")
    _lib = lib
    return lib


def check(rc, ctx=None):
    if rc != 0:
        msg = load().jic_last_error(ctx)
        raise JicError(f"libjic_b200 error {rc}: {msg.decode() if msg else '?'}")
