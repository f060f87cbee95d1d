"""GPU legs of the golden vectors that were added after this round's GPU budget was spent: every `refsrc_*` file (the reference's own
source on the NumPy stand-in, tests/golden/make_reference_golden.py) and the five newest cases of tests/golden/make_golden.py
(field_solver 3, field_solver 1 with walls, G = 5 and G = 3 grids, a box with its own length_y / length_z).

Same check as tests/test_golden.py::test_cuda_reproduces_golden (both CUDA engines through the C ABI, 1e-5 relative in fp64).  The file
sorts last on purpose: the driver runs `pytest -x`, these cases have not been on hardware yet, and a surprise here must not hide the
rest of the suite."""
import copy
import os

import numpy as np
import pytest

from test_golden import GPU_FILES_LATE, cuda_reproduces_golden
from test_simulation_driver import _DRV, DRIVER_CASES, REFSRC, RUN_CASE, Simulation


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["indexed", "binned"])
@pytest.mark.parametrize("path", GPU_FILES_LATE, ids=[os.path.basename(f)[:-4] for f in GPU_FILES_LATE])
def test_cuda_reproduces_late_golden(path, engine):
    cuda_reproduces_golden(path, engine)


@pytest.mark.gpu
def test_run_of_the_reference_source_is_reproduced_end_to_end():
    """`Simulation(parameters).run()` here vs the reference's own `Simulation(parameters).run()` (on the stand-in) for the same parameter
    dictionary: initial particles from the device Threefry sampler, 40 steps at CFL 4.5 (multi-cell jumps), every history."""
    ref = REFSRC[RUN_CASE]
    a = np.load(os.path.join(_DRV, "refsrc_driver_arrays.npz"))
    out = Simulation(copy.deepcopy(DRIVER_CASES[RUN_CASE])).run()
    assert set(ref["output_keys"]) <= set(out)
    np.testing.assert_allclose(out["plasma_frequency"], ref["plasma_frequency"], rtol=1e-14)
    np.testing.assert_allclose([out["time_array"][0], out["time_array"][1], out["time_array"][-1]], ref["time_array"][:3], rtol=1e-14)
    np.testing.assert_allclose(out["initial_positions"], a[f"{RUN_CASE}__positions"], rtol=0, atol=1e-15 * 0.01)
    for k in ("electric_field", "magnetic_field", "current_density", "charge_density", "positions", "velocities"):
        err = np.abs(np.asarray(out[k]) - a[f"run__{k}"]).max() / max(np.abs(a[f"run__{k}"]).max(), 1e-300)
        assert err < 1e-5, (k, err)
