#!/usr/bin/env python
"""Device-timed measurements of the SURVEY 8(f) rows next to the headline path (run on a B200: python profiles/measure_rows.py).
One JSON line per row: field_solver overhead on the binned engine, initial sampling, Crank-Nicolson iteration rate."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "jax-in-cell_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from jaxincell_b200 import HotPath, sample_particles  # noqa: E402

PEAK = bench.peaks()[0]


def timed(fn, reps=1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def crank_nicolson(dev):
    """Particle-iterations per second of the two Crank-Nicolson pushes on the same particles in the same process: the cell-sorted one
    (default from 200000 particles up) and the unsorted one (JIC_CN_SORTED_MIN raised above N), and how far their fields are apart."""
    class A:
        grid, particles = 4096, 20_000_000
    w = bench.workload(A, 1)
    x0, v0 = bench.make_particles(w, torch, dev, torch.float64, 1701, "random")
    S, max_it = 2, 6
    fields = {}
    for name, threshold in (("sorted", "0"), ("unsorted", str(1 << 40))):
        os.environ["JIC_CN_SORTED_MIN"] = threshold
        hp = HotPath(species=w["species"], length=w["length"], G=w["G"], dt=0.3 * w["dt"], time_evolution_algorithm=1, cn_substeps=S,
                     cn_max_iterations=max_it, cn_tolerance=1e-30)
        hp.set_external_fields(None, None)
        hp.initialize(x0, v0)
        outs = hp.alloc_outputs(10)
        hp.run(10, outputs=outs)
        fields[name] = outs["electric_field"].clone()
        it0 = hp.picard_iterations()[1]
        ms = timed(lambda: hp.run(10, outputs=outs))
        iters = hp.picard_iterations()[1] - it0
        N = hp.N
        bytes_per = (12 + 2 * S) * 8 + 1
        print(json.dumps({"row": f"8f-4 Crank-Nicolson, {name} push", "cn_sorted": hp.store_stats()["cn_sorted"], "particles": N, "substeps": S,
                          "picard_iterations": iters, "ms_per_iteration": ms / iters, "particle_iterations_per_s": N * iters / ms * 1e3,
                          "achieved_GBps": bytes_per * N * iters / ms / 1e6, "peak_GBps": PEAK, "frac": bytes_per * N * iters / ms / 1e6 / PEAK,
                          "finite": bool(torch.isfinite(outs["electric_field"][-1]).all()),
                          "note": f"algorithmic bytes per particle and iteration = {bytes_per} (x,v in, x,v out, S staggered positions in and out, "
                                  "alive byte); the sorted push pays one counting sort of the state per step on top (not counted as algorithmic)"}),
              flush=True)
        hp.close()
    del os.environ["JIC_CN_SORTED_MIN"]
    a, b = fields["sorted"], fields["unsorted"]
    print(json.dumps({"row": "8f-4 Crank-Nicolson, sorted vs unsorted push", "steps": 10,
                      "max_rel_diff_E": float((a - b).abs().max() / b.abs().max())}), flush=True)


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None

    class A:  # the bench workload
        grid, particles = 4096, 100_000_000
    w = bench.workload(A, 1)

    if only in (None, "cn"):
        crank_nicolson(dev)
    if only == "cn":
        return

    # ---- initial sampling (jic_sample_particles): 48 B written per particle in fp64
    n = 50_000_000
    sp = [dict(count=n, seed_position=1701, seed_velocity=1704, random_positions=[True, True, True], velocity_plus_minus=[True, False, False],
               perturbation_amplitude=[1e-7, 0, 0], perturbation_wavenumber=[8, 0, 0], vth_over_c=[0.05, 0.01, 0.01], drift_speed=[6e7, 0, 0])]
    sample_particles(sp, (w["length"],) * 3)
    ms = timed(lambda: sample_particles(sp, (w["length"],) * 3), 3)
    print(json.dumps({"row": "8f-2 initial sampling (k_sample_species)", "particles": n, "ms": ms, "particles_per_s": n / ms * 1e3,
                      "achieved_GBps": 48 * n / ms / 1e6, "peak_GBps": PEAK, "frac": 48 * n / ms / 1e6 / PEAK,
                      "note": "6 Threefry-2x32-20 blocks + erfinv per particle; algorithmic bytes = 48 B written per particle"}), flush=True)

    # ---- field_solver on the binned engine: same workload as bench.py, 5e7 particles
    A.particles = 50_000_000
    w = bench.workload(A, 1)
    x0, v0 = bench.make_particles(w, torch, dev, torch.float64, 1701, "random")
    res = {}
    for fs in (0, 1, 2):
        hp = HotPath(species=w["species"], length=w["length"], G=w["G"], dt=w["dt"], engine="binned", field_solver=fs)
        hp.set_external_fields(None, None)
        hp.initialize(x0, v0)
        outs = hp.alloc_outputs(50)
        hp.run(50, outputs=outs)
        ms = timed(lambda: hp.run(50, outputs=outs)) / 50
        a, b = hp.profile_steps(10)
        res[fs] = dict(ms_per_step=ms, push_ms=a / 10, grid_ms=b / 10, finite=bool(torch.isfinite(outs["electric_field"][-1]).all()))
        hp.close()
    print(json.dumps({"row": "8f-3 per-step electrostatic correction (binned engine, 5e7 particles, G=4096)", "field_solver": res,
                      "note": "push_ms includes the face deposit (STAG instantiation of k_push); grid_ms includes k_gauss (filter of the face "
                              "component + G^2 circular convolution) in front of the field kernel"}), flush=True)
    del x0, v0
    torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
