#!/usr/bin/env python
"""Headline benchmark: particle-steps/s of the fused PIC hot path on the synthetic scaling plasma.

  python bench.py --gpus 1 --steps K --warmup W            (torchrun launches N>1 ranks, one per GPU)
  python bench.py --impl reference ...                     (the CPU restatement of the reference on the host cores)

One JSON line on rank 0.  `value` is device-timed (CUDA events around the graph-captured time loop, barrier + sync on
both sides, max over ranks) with the particles resident in HBM; `e2e` is the same metric through the host-buffer
entry point (pinned host -> device copies of the initial particles and device -> host copies of the field histories
inside the timed region).  Weak scaling (default): every rank owns --particles macro-particles; --scaling strong: the
ranks share --total-particles.  Before anything is timed, `parity` steps a small plasma of the same shape on the same
ranks and compares it with the compiled oracle (checker only); `sustained` repeats the timed replay until the clocks
have settled under the power cap; `roofline.kernel_ms` is taken by the push kernel itself inside the timed replay.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "jax-in-cell_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

C_LIGHT = 2.99792458e8
EPS0 = 8.85418782e-12
QE = 1.60217663e-19
ME = 9.10938371e-31
MP = 1.67262193e-27
BYTES_PER_PARTICLE_STEP = {"f64": 64, "f32": 32}  # read + write of x_{n+1/2}, v_x, v_y, v_z


def workload(args, world):
    """G=4096, dx=0.01/70 m, CFL 1, periodic, filter 5/0.5/(1,2,4); electrons N/2 as two
    counter-streaming beams (+-0.2c, thermal 0.05c) with uniform-random x, ions N/2 cold (T_i/T_e = 1e-2)."""
    G = args.grid
    dx = 0.01 / 70
    length = G * dx
    dt = 1.0 * dx / C_LIGHT
    n_e = args.particles // 2
    n_i = args.particles - n_e
    vth_e = 0.05
    gpdl = 2.0

    def weight(n_global):  # jaxincell/_state_initialization.py:172-185 with the GLOBAL species count
        return EPS0 * ME * C_LIGHT ** 2 / QE ** 2 * G ** 2 / length / (2 * n_global) * vth_e ** 2 * gpdl ** 2

    we, wi = weight(n_e * world), weight(n_i * world)
    species = [dict(count=n_e, q=-QE * we, m=ME * we, qm=-QE / ME), dict(count=n_i, q=QE * wi, m=MP * wi, qm=QE / MP)]
    return dict(G=G, length=length, dt=dt, species=species, vth_e=vth_e, n_e=n_e, n_i=n_i)


def make_particles(w, torch, device, dtype, seed, order):
    """Synthetic particles generated ON the device (Philox), (N,3) row-major like the reference's arrays."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n_e, n_i, L = w["n_e"], w["n_i"], w["length"]
    N = n_e + n_i
    x0 = torch.empty((N, 3), dtype=dtype, device=device)
    v0 = torch.empty((N, 3), dtype=dtype, device=device)
    if order == "random":
        x0[:, 0].uniform_(-L / 2, L / 2, generator=g)
    else:  # the reference default (random_positions_x=False): linspace per species -> sorted, the atomics worst case
        x0[:n_e, 0] = torch.linspace(-L / 2, L / 2, n_e, dtype=dtype, device=device)
        x0[n_e:, 0] = torch.linspace(-L / 2, L / 2, n_i, dtype=dtype, device=device)
    x0[:, 1:].uniform_(-L / 2, L / 2, generator=g)
    s = C_LIGHT / 2 ** 0.5
    v0[:n_e, 0].normal_(0.0, 0.05 * s, generator=g)
    v0[:n_e, 0] += 0.2 * C_LIGHT
    v0[:n_e:2, 0] *= 1.0
    v0[1:n_e:2, 0] *= -1.0  # velocity_plus_minus_x: alternate sign by index
    v0[:n_e, 1:].normal_(0.0, 0.01 * s, generator=g)
    vthi = 0.05 * (1e-2 * ME / MP) ** 0.5
    v0[n_e:, 0].normal_(0.0, vthi * s, generator=g)
    v0[n_e:, 1:].normal_(0.0, 0.2 * vthi * s, generator=g)
    return x0, v0


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML in-process (a sample
    every few ms; the timed region is tens of ms), falling back to polling nvidia-smi."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        self.rows.append((mhz, self.max_mhz, {k for k, b in self.BITS.items() if mask & b}))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            c = [x.strip() for x in out.split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            self.rows.append((float(c[0]), float(c[1]), {n for n, v in zip(names, c[3:7]) if v.lower().startswith("active")}))

    def run(self):
        while not self._stop_evt.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.002 if self.nvml else 0.1)

    def mark(self):
        """Index of the next sample: lets the caller cut out the samples taken inside the timed region."""
        return len(self.rows)

    def stop(self, lo=0, hi=None):
        self._stop_evt.set()
        self.join(timeout=6)
        rows = self.rows[lo:hi] or self.rows
        sm = sorted(r[0] for r in rows)
        reasons = set()
        for r in rows:
            reasons |= r[2]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[1] for r in rows), default=None),
                "reasons": sorted(reasons), "samples": len(rows), "source": "nvml" if self.nvml else "nvidia-smi"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(dtype, engine, n_particles):
    """dram__bytes_read.sum + dram__bytes_write.sum of one push launch from the committed ncu capture (profiles/), scaled to this
    run's particle count (the kernel streams every particle once, so its traffic is linear in N); None when there is no capture."""
    path = os.path.join(ROOT, "profiles", "r02_push_traffic.json")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r01_push_traffic.json")
    if engine != "binned" or not os.path.exists(path):
        return None
    rec = json.load(open(path)).get(dtype)
    if not rec:
        return None
    return (rec["dram_bytes_read"] + rec["dram_bytes_write"]) * (n_particles / rec["particles"])


def sample_plasma(w, n_sample, np):
    """A bounded sample of the bench workload for the CPU legs: same geometry and distributions, charges rescaled to the same density."""
    rng = np.random.default_rng(1701)
    L = w["length"]
    n_e = n_i = n_sample // 2
    x0 = np.zeros((n_sample, 3)); v0 = np.zeros((n_sample, 3))
    x0[:, 0] = rng.uniform(-L / 2, L / 2, n_sample)
    s = C_LIGHT / 2 ** 0.5
    v0[:n_e, 0] = (0.2 * C_LIGHT + 0.05 * s * rng.standard_normal(n_e)) * (-1.0) ** np.arange(n_e)
    v0[:n_e, 1:] = 0.01 * s * rng.standard_normal((n_e, 2))
    sp = w["species"]
    scale = (w["n_e"] / n_e)
    q = np.concatenate([np.full(n_e, sp[0]["q"] * scale), np.full(n_i, sp[1]["q"] * scale)])
    m = np.concatenate([np.full(n_e, sp[0]["m"] * scale), np.full(n_i, sp[1]["m"] * scale)])
    qm = np.concatenate([np.full(n_e, sp[0]["qm"]), np.full(n_i, sp[1]["qm"])])
    return x0, v0, q, m, qm


def cpu_port_rate(w, seconds_target=15.0, threads=None, n_sample=4_000_000, steps=None, warmup=0):
    """The compiled oracle (oracle/c/jic_oracle.c: C + OpenMP restatement of the reference's step, one thread per host core, thread-private
    grids summed) over a bounded sample of the workload.  steps=None: as many steps as fit about `seconds_target`; otherwise `warmup`
    untimed + exactly `steps` timed steps.  Falls back to the NumPy port when the C oracle cannot be built or loaded."""
    import numpy as np
    try:
        from oracle import c_port as CP
        CP.load()
    except Exception as e:  # noqa: BLE001  (no gcc / no prebuilt library on this host)
        rate, cores, sample, s_per_step = cpu_numpy_port_rate(w, seconds_target, threads, min(n_sample, 400_000), steps, warmup)
        return rate, cores, sample + f" (C oracle unavailable: {type(e).__name__})", s_per_step
    threads = threads or len(os.sched_getaffinity(0)) or 1
    x0, v0, q, m, qm = sample_plasma(w, n_sample, np)
    kw = dict(length=w["length"], G=w["G"], dt=w["dt"], keep_particles=False, threads=threads,
              solver=dict(filter_passes=5, filter_alpha=0.5, filter_strides=(1, 2, 4)))
    if steps is None:  # calibrate on 2 steps, then fill the time budget
        probe = CP.run(x0, v0, q, m, qm, total_steps=2, **kw)["step_seconds"]
        warmup, steps = 1, int(max(3, min(400, seconds_target / max(float(probe[-1]), 1e-6))))
    secs = CP.run(x0, v0, q, m, qm, total_steps=warmup + steps, **kw)["step_seconds"][warmup:]
    el = float(secs.sum())
    return (n_sample * steps / el, threads,
            f"{n_sample} particles x {steps} steps, G={w['G']}, C/OpenMP closed-form port of the reference (oracle/c/jic_oracle.c), {threads} threads", el / steps)


def cpu_numpy_port_rate(w, seconds_target=15.0, threads=None, n_sample=400_000, steps=None, warmup=0):
    """Oracle port (NumPy closed form) on the host cores: particles split over threads, grids summed.
    steps=None: run for about `seconds_target`; otherwise `warmup` untimed + exactly `steps` timed steps over the sample."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import closed_form as CF
    threads = threads or os.cpu_count() or 1
    L, G, dt = w["length"], w["G"], w["dt"]
    x0, v0, q, m, qm = sample_plasma(w, n_sample, np)
    dom = CF.Domain(L, G, dt)
    solver = dict(filter_passes=5, filter_alpha=0.5, filter_strides=(1, 2, 4))
    chunks = np.array_split(np.arange(n_sample), threads)
    states = [CF.start(x0[c], v0[c], q[c], m[c], qm[c], dom, 0, 0, 0, 0, solver) for c in chunks]
    # the grid is shared: every chunk deposits, the sums are combined, one field solve (same as the multi-GPU scheme)
    from oracle import literal as LT

    def push_deposit(st, E, B):
        E_p, B_p = CF.gather_EB(st.x_half[:, 0], E, B, dom, 0, 0)
        x_pp, v_new = CF.push_boris(dt, st.x_half, st.v, st.qm, E_p, B_p)
        x_pp, v_new, qq, mm, qqm = LT.set_BC_particles(x_pp, v_new, st.q, st.m, st.qm, dom.dx, dom.grid, *dom.box, 0, 0)
        x_new = LT.set_BC_positions(x_pp - (dt / 2) * v_new, dom.dx, dom.grid, *dom.box, 0, 0)
        J = CF.deposit_current_raw(st.x_half[:, 0], x_new[:, 0], x_pp[:, 0], v_new, qq, dom, 0, 0)
        rho = CF.deposit_rho_raw(x_new[:, 0], qq, dom, 0, 0)
        st.x_half, st.v = x_pp, v_new
        return J, rho

    E = sum(s_.E for s_ in states); B = states[0].B
    J = sum(s_.J for s_ in states)  # start() filtered each chunk's J; the filter is linear
    pool = ThreadPoolExecutor(threads)

    def one_step(E, B, J):
        E, B = LT.field_update1(E, B, dom.dx, dt / 2, J, 0, 0)
        parts = list(pool.map(lambda st: push_deposit(st, E, B), states))
        J = LT.filter_vector_field(sum(p_[0] for p_ in parts), 5, 0.5, (1, 2, 4), 0, 0)
        LT.filter_scalar_field(sum(p_[1] for p_ in parts), 5, 0.5, (1, 2, 4), 0, 0)
        E, B = LT.field_update2(E, B, dom.dx, dt / 2, J, 0, 0)
        return E, B, J

    for _ in range(warmup):
        E, B, J = one_step(E, B, J)
    done, t0 = 0, time.perf_counter()
    while True:
        E, B, J = one_step(E, B, J)
        done += 1
        el = time.perf_counter() - t0
        if (steps is not None and done >= steps) or (steps is None and (el > seconds_target or done >= 200)):
            break
    pool.shutdown()
    return n_sample * done / el, threads, f"{n_sample} particles x {done} steps, G={G}, NumPy closed-form port of the reference, {threads} threads", el / done


def real_reference_status():
    """The unmodified reference lives in baseline/_ref when it could be installed.  It is pure Python on
    JAX; importing it needs jax, jax_tqdm and matplotlib, none of which exist in this image."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import jaxincell  # noqa: F401
        return None
    except Exception as e:  # noqa: BLE001
        return f"{type(e).__name__}: {e}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args, 1)
    why_not = real_reference_status()
    # a "step" of this arm = one PIC step over a bounded sample of the workload (4e6 particles); K and W as given, capped so that the
    # run ends within a few minutes on any host
    k = max(1, min(args.steps, 400))
    rate, cores, sample, s_per_step = cpu_port_rate(w, steps=k, warmup=min(max(args.warmup, 0), 20))
    line = {"impl": "reference", "metric": "particle-steps/sec", "value": rate, "unit": "particle-steps/s", "n_gpus": args.gpus,
            "steps": k, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic two-beam plasma (SURVEY 8d config 5): G={args.grid}, {args.particles} macro-particles per GPU, CFL 1, periodic, "
                                   f"filter 5/0.5/(1,2,4), x order {args.order}", "particles_per_gpu": args.particles, "grid": args.grid,
                       "cpu_sample": "the CPU arm steps a bounded sample of this workload (same geometry, distributions and density): " + sample},
            "cpu_baseline": {"value": rate, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference itself cannot run here (import jaxincell from baseline/_ref -> " + str(why_not) + "); "
                    "this is the oracle port of its algorithm (compiled C + OpenMP restatement, NumPy if no compiler) on the host cores"}
    emit(line)


def np_isfinite(a):
    import numpy as np
    return np.isfinite(np.asarray(a)).all()


def parity_check(args, torch, dist, device, rank, world, engine):
    """N-rank parity where the driver sees it: a small plasma of the bench's shape (same grid, geometry, distributions and filter),
    index-sharded over the ranks exactly like the timed run, stepped through the same kernels and the same grid reduction; rank 0
    compares every step's E, B, J, rho with the compiled oracle (checker only) and all ranks compare their final fields bit for bit
    with rank 0's."""
    import numpy as np
    from jaxincell_b200 import HotPath
    from jaxincell_b200._parallel import shard_particles, shard_species
    from oracle import c_port as CP
    n_total, T = args.parity_particles, args.parity_steps

    class A:
        grid, particles = args.grid, n_total
    w = workload(A, 1)
    x0, v0, q, m, qm = sample_plasma(w, n_total, np)
    ne = n_total // 2
    species = [dict(count=ne, q=float(q[0]), m=float(m[0]), qm=float(qm[0])), dict(count=n_total - ne, q=float(q[-1]), m=float(m[-1]), qm=float(qm[-1]))]
    xs, vs, _ = shard_particles(x0, v0, species, rank, world)
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    hp = HotPath(species=shard_species(species, rank, world), dtype=dtype, length=w["length"], G=w["G"], dt=w["dt"], engine=engine)
    if world > 1:
        hp.comm_init_from_torch()
    hp.set_external_fields(None, None)
    hp.initialize(torch.from_numpy(xs).to(device=device, dtype=dtype), torch.from_numpy(vs).to(device=device, dtype=dtype))
    out = hp.run(T)
    hp.check_status()
    mode = hp.comm_mode()
    keys = ("electric_field", "magnetic_field", "current_density", "charge_density")
    identical = True
    if world > 1:
        for k in keys:
            ref0 = out[k].clone()
            dist.broadcast(ref0, src=0)
            same = torch.tensor([1 if torch.equal(ref0, out[k]) else 0], device=device)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            identical = identical and bool(int(same.item()))
    res = None
    if rank == 0:
        CP.load()
        ref = CP.run(x0, v0, q, m, qm, length=w["length"], G=w["G"], dt=w["dt"], total_steps=T, keep_particles=False,
                     threads=len(os.sched_getaffinity(0)) or 1, solver=dict(filter_passes=5, filter_alpha=0.5, filter_strides=(1, 2, 4)))
        worst, per_key = 0.0, {}
        for k in keys:
            a, b = out[k].double().cpu().numpy(), np.asarray(ref[k])
            # per step (row-wise): max |a - b| over the grid / max |b| over the grid of that step
            den = np.maximum(np.abs(b).reshape(T, -1).max(axis=1), 1e-300)
            err = float((np.abs(a - b).reshape(T, -1).max(axis=1) / den).max())
            per_key[k] = err
            worst = max(worst, err)
        tol = 1e-5 if args.dtype == "f64" else 1e-3
        res = {"ok": bool(worst < tol and identical), "max_rel_err": worst, "per_key": per_key, "tolerance": tol, "ranks_identical": identical,
               "norm": "per step: max|cuda - oracle| over the grid / max|oracle| over the grid, worst step", "steps": T, "particles": n_total,
               "ranks": world, "grid_reduction": mode, "checker": "oracle/c/jic_oracle.c (CPU, test infrastructure)"}
    hp.close()
    return res


def timed_run(hp, torch, dist, world, K, outs, barrier):
    """K graph-replayed steps bracketed by barrier + synchronize; returns (ms, device-side ms of the push kernel inside the same replay)."""
    barrier()
    try:
        hp.push_kernel_time(reset=True)
        have_ktime = True
    except Exception:  # noqa: BLE001  (INDEXED engine or an older library)
        have_ktime = False
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    hp.run(K, outputs=outs)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    kms, kn = hp.push_kernel_time(reset=True) if have_ktime else (0.0, 0)
    return ms, (kms / kn if kn else None), kn


_JSON_FD = None


def keep_stdout_for_the_json_line():
    """Libraries write to stdout behind Python's back (NCCL prints its version line there at communicator creation): from here on
    file descriptor 1 is stderr, and the ONE JSON line goes to the descriptor that was stdout when the process started."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.buffer.write(data); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    keep_stdout_for_the_json_line()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=100_000_000, help="macro-particles PER GPU (weak scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--total-particles", type=int, default=100_000_000, help="macro-particles over ALL GPUs (--scaling strong)")
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--engine", default=os.environ.get("JIC_BENCH_ENGINE", "auto"))
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--order", default="random", choices=["random", "sorted"])
    ap.add_argument("--deposit", default="auto", choices=["auto", "global", "shared"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--no-f32", action="store_true")
    ap.add_argument("--parity-particles", type=int, default=2_000_000)
    ap.add_argument("--parity-steps", type=int, default=10)
    ap.add_argument("--sustained-seconds", type=float, default=1.5)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # the checker / CPU baseline is OpenMP code: its idle threads must sleep, not spin, while the GPU legs are timed
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
    import torch
    import torch.distributed as dist
    from jaxincell_b200 import HotPath, JicError

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU path); use --impl reference for the CPU restatement")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    # host side of the e2e leg: run (and allocate the pinned buffers) on the CPUs next to this GPU, as `numactl` would; a pinned
    # buffer on the far socket halves the host->device rate (measured: 26 vs 45 GB/s).  Undone before the CPU baseline.
    all_cpus = os.sched_getaffinity(0)
    affinity = "unchanged"
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        affinity = f"GPU-local CPUs (nvmlDeviceSetCpuAffinity): {len(os.sched_getaffinity(0))} of {len(all_cpus)}"
    except Exception as e:  # noqa: BLE001
        affinity = f"unchanged ({type(e).__name__})"
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    if args.scaling == "strong":  # fixed total work: every rank owns 1/world of --total-particles
        args.particles = args.total_particles // world
    w = workload(args, world)
    engine = args.engine
    if engine == "auto":
        engine = "binned"
        try:
            HotPath(species=[dict(count=8, q=1.0, m=1.0, qm=1.0)], length=1.0, G=8, dt=1e-9, engine="binned").close()
        except JicError:
            engine = "indexed"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity first (small, same shape, same sharding and reduction): a fast wrong answer is not a result
    parity = None
    if not args.no_parity:
        all_now = os.sched_getaffinity(0)
        if rank == 0:
            os.sched_setaffinity(0, all_cpus)  # the checker may use every host core
        parity = parity_check(args, torch, dist, device, rank, world, engine)
        if rank == 0:
            os.sched_setaffinity(0, all_now)
        barrier()

    hp = HotPath(species=w["species"], dtype=dtype, length=w["length"], G=w["G"], dt=w["dt"], engine=engine, deposit=args.deposit)
    if world > 1:
        hp.comm_init_from_torch()
    reduction = hp.comm_mode()
    x0, v0 = make_particles(w, torch, device, dtype, 1701 + rank, args.order)
    hp.set_external_fields(None, None)
    hp.initialize(x0, v0)
    N, G, K, W = hp.N, hp.G, args.steps, max(args.warmup, 3)

    outs = hp.alloc_outputs(K)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()  # samples clocks from here to the end of the profiled steps: the GPU is under load throughout
    hp.run(W, outputs=hp.alloc_outputs(W))
    hp.run(K, outputs=outs)  # untimed: instantiates the CUDA graphs the timed call replays
    barrier()
    l0 = hp.launch_count()
    m0 = sampler.mark() if sampler else 0
    ms, kernel_ms_graph, kernel_launches = timed_run(hp, torch, dist, world, K, outs, barrier)
    m1 = sampler.mark() if sampler else 0
    launches = hp.launch_count() - l0
    hp.check_status()
    # the same kernels outside the graph, CUDA events around the particle kernel(s) and the grid part of every step (kept as a
    # cross-check of the in-graph figure; its launches are not overlapped, so it reads a little higher)
    n_prof = min(K, 10)
    ms_push, ms_grid = hp.profile_steps(n_prof)
    barrier()
    clocks = sampler.stop(m0, max(m1, m0 + 1)) if sampler else None
    t = torch.tensor([ms, kernel_ms_graph if kernel_ms_graph is not None else ms_push / n_prof, ms_push / n_prof, ms_grid / n_prof],
                     dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, kernel_ms, ms_push_step, ms_grid_step = (float(v) for v in t.cpu())
    value = N * world * K / (ms * 1e-3)
    energy_ok = bool(torch.isfinite(outs["electric_field"][-1]).all().item())

    # ---- end to end: host buffers in, host buffers out, through the public API a user calls for every run on an existing
    #      context (HotPath.initialize + HotPath.run).  Context creation (cudaMalloc, NCCL communicator) is set-up, not part of it.
    e2e = None
    if not args.no_e2e:
        del outs
        hx = torch.empty((N, 3), dtype=dtype, pin_memory=True)
        hv = torch.empty((N, 3), dtype=dtype, pin_memory=True)
        hx.copy_(x0); hv.copy_(v0)
        del x0, v0
        torch.cuda.empty_cache()
        host_out = {k: torch.empty(s, dtype=dtype, pin_memory=True) for k, s in
                    (("electric_field", (K, G, 3)), ("magnetic_field", (K, G, 3)), ("current_density", (K, G, 3)), ("charge_density", (K, G)))}
        dev_out = hp.alloc_outputs(K)
        for timed in (False, True):  # one untimed pass first: staging buffers, copy stream and graphs of this path exist afterwards
            barrier()
            t0 = time.perf_counter()
            hp.initialize_host(hx, hv)  # pinned host -> device in chunks, overlapped with the start-up kernels; synchronises
            t1 = time.perf_counter()
            o2 = hp.run(K, outputs=dev_out)
            for k, h in host_out.items():
                h.copy_(o2[k], non_blocking=True)
            barrier()
            t2 = time.perf_counter()
        hp.check_status()
        el = t2 - t0
        tt = torch.tensor([el], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        el = float(tt.cpu()[0])
        es = dtype.itemsize
        e2e = {"value": N * world * K / el, "unit": "particle-steps/s", "h2d_bytes_per_step": int(world * N * 6 * es / K),
               "d2h_bytes_per_step": int(G * 10 * es), "seconds": el,
               "seconds_breakdown_rank0": {"upload_and_start_up": t1 - t0, "steps_and_download": t2 - t1},
               "host_cpu_affinity": affinity,
               "what": f"pinned-host->device copy of x0,v0 ({N * 6 * es / 1e9:.1f} GB per GPU, once per run, chunked and overlapped with the "
                       f"leap-frog start-up / binning kernels: HotPath.initialize_host) + {K} steps + device->host copy of the E,B,J,rho "
                       f"histories, on an existing context; second of two identical passes (the first one creates the staging buffers)"}
        ok2 = bool(torch.isfinite(host_out["electric_field"][-1]).all().item())
        energy_ok = energy_ok and ok2
        # what the link gives: one plain pinned -> device copy of the same x0 buffer, device-timed (the denominator of the upload time)
        sink = torch.empty((N, 3), dtype=dtype, device=device)
        sink.copy_(hx, non_blocking=True)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        c0.record()
        sink.copy_(hx, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        link = N * 3 * es / (c0.elapsed_time(c1) * 1e-3) / 1e9
        e2e["h2d_link_gbs_measured"] = link
        e2e["upload_gbs_achieved"] = N * 6 * es / (t1 - t0) / 1e9
        e2e["upload_note"] = ("upload_and_start_up moves x0,v0 at upload_gbs_achieved against h2d_link_gbs_measured for a bare cudaMemcpyAsync of "
                              "the same pinned buffer on this box; the rest of it is the start-up kernel of the last chunk and the first field solve")
        del sink
        del hx, hv, host_out, dev_out
    # ---- sustained: the same K-step replay repeated until the board has been under load for a while (power cap, clocks settle)
    sustained = None
    if not args.no_sustained:
        s2 = ClockSampler(local) if rank == 0 else None
        if s2:
            s2.start()
        outs = hp.alloc_outputs(K)
        hp.run(K, outputs=outs)  # (after the e2e leg re-initialised the context: the same state, graphs already instantiated)
        reps, t_start = [], time.perf_counter()
        budget = torch.tensor([0.0], device=device)
        while True:
            ms_r, k_r, _ = timed_run(hp, torch, dist, world, K, outs, barrier)
            tt = torch.tensor([ms_r, k_r if k_r is not None else 0.0], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            reps.append(tuple(float(v) for v in tt.cpu()))
            budget[0] = 1.0 if (time.perf_counter() - t_start) >= args.sustained_seconds or len(reps) >= 400 else 0.0
            if world > 1:
                dist.all_reduce(budget, op=dist.ReduceOp.MAX)
            if budget.item() > 0:
                break
        c2 = s2.stop() if s2 else None
        tail = sorted(r[0] for r in reps[len(reps) // 2:])  # the second half: after the clocks have settled
        med = tail[len(tail) // 2]
        ktail = sorted(r[1] for r in reps[len(reps) // 2:])
        sustained = {"value": N * world * K / (med * 1e-3), "unit": "particle-steps/s", "ms_per_step": med / K, "kernel_ms": ktail[len(ktail) // 2],
                     "replays": len(reps), "steps_per_replay": K, "seconds": time.perf_counter() - t_start,
                     "what": f"median over the second half of {len(reps)} back-to-back replays of the same {K}-step graph loop", "clocks": c2}

    hp.close()
    torch.cuda.empty_cache()

    # ---- the drop-in entry point itself: Simulation(parameters).run() from the parameter dictionary to the output dictionary
    #      (particles sampled on the device from the reference's seeds, context creation, K steps, histories to the host)
    e2e_sim = None
    if not args.no_e2e and engine == "binned":
        try:
            from jaxincell_b200 import Simulation
            par = {
                "domain_parameters": dict(total_steps=K, number_grid_points=G, length=w["length"], timestep_over_spatialstep_times_c=1.0),
                "species_parameters": {
                    "electrons": dict(number_pseudoparticles=w["n_e"] * world, vth_over_c_x=0.05, vth_over_c_y=0.01, vth_over_c_z=0.01,
                                      random_positions_x=True, velocity_plus_minus_x=True, drift_speed_x=0.2 * C_LIGHT,
                                      perturbation_amplitude_x=0.0, grid_points_per_Debye_length=2.0),
                    "ions": dict(number_pseudoparticles=w["n_i"] * world, random_positions_x=True, perturbation_amplitude_x=0.0,
                                 ion_temperature_over_electron_temperature_x=1e-2, ion_temperature_over_electron_temperature_y=1e-2,
                                 ion_temperature_over_electron_temperature_z=1e-2, grid_points_per_Debye_length=2.0)},
                "solver_parameters": dict(print_info=False, particle_history=False, dtype="float64" if args.dtype == "f64" else "float32",
                                          filter_passes=5, filter_alpha=0.5, filter_strides=(1, 2, 4))}
            times = []
            for _ in range(2):  # the second call is the one reported (the first pays one-time CUDA module loads)
                barrier()
                t0 = time.perf_counter()
                o_sim = Simulation(par).run()
                barrier()
                times.append(time.perf_counter() - t0)
            tt = torch.tensor([times[-1]], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            el = float(tt.cpu()[0])
            e2e_sim = {"value": N * world * K / el, "unit": "particle-steps/s", "seconds": el, "first_call_seconds": times[0],
                       "finite": bool(np_isfinite(o_sim["electric_field"][-1])),
                       "what": f"wall clock of Simulation(parameters).run() -- dict in, output dict out, particle_history=False: context creation "
                               f"(cudaMalloc of the particle store), Threefry sampling of {N * world} particles on the device"
                               + (f" ({world} ranks, each its own index slice)" if world > 1 else "")
                               + f", start-up, {K} steps, E,B,J,rho histories to the host; second of two calls"}
            del o_sim
        except Exception as e:  # noqa: BLE001
            e2e_sim = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()

    # ---- the same workload in fp32 (sub-record; the headline stays the reference's own precision)
    f32 = None
    if not args.no_f32 and args.dtype == "f64" and engine == "binned":
        try:
            hp32 = HotPath(species=w["species"], dtype=torch.float32, length=w["length"], G=w["G"], dt=w["dt"], engine=engine)
            if world > 1:
                hp32.comm_init_from_torch()
            xa, va = make_particles(w, torch, device, torch.float32, 1701 + rank, args.order)
            hp32.set_external_fields(None, None)
            hp32.initialize(xa, va)
            del xa, va
            o32 = hp32.alloc_outputs(K)
            hp32.run(W, outputs=hp32.alloc_outputs(W))
            hp32.run(K, outputs=o32)
            ms32, k32, _ = timed_run(hp32, torch, dist, world, K, o32, barrier)
            hp32.check_status()
            t32 = torch.tensor([ms32, k32 if k32 is not None else 0.0], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(t32, op=dist.ReduceOp.MAX)
            ms32, k32 = (float(v) for v in t32.cpu())
            f32 = {"value": N * world * K / (ms32 * 1e-3), "unit": "particle-steps/s", "ms_per_step": ms32 / K, "kernel_ms": k32,
                   "finite": bool(torch.isfinite(o32["electric_field"][-1]).all().item())}
            hp32.close()
        except Exception as e:  # noqa: BLE001
            f32 = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        peak, peak_src = peaks()
        bpp = BYTES_PER_PARTICLE_STEP[args.dtype]
        achieved = bpp * N / (kernel_ms * 1e-3) / 1e9
        in_graph = kernel_ms_graph is not None
        line = {
            "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic",
            "config": {"workload": f"synthetic two-beam plasma (SURVEY 8d config 5): G={G}, {N} macro-particles per GPU"
                                   + (f" ({N * world} in total, strong scaling)" if args.scaling == "strong" else "")
                                   + f", CFL 1, periodic, filter 5/0.5/(1,2,4), x order {args.order}", "engine": engine, "particles_per_gpu": N, "grid": G,
                       "grid_reduction": {"single": "none (one rank)", "nccl": "NCCL all-reduce of the raw grid before the field kernel",
                                          "fused": "fused into the field kernel: peers' raw grids read over NVLink (CUDA IPC), summed in rank order"}[reduction],
                       "l2": "particle state (>= 3.2 GB per GPU) is far larger than L2; no flush needed",
                       "untimed_steps_before_timing": W + K,
                       "timed_region": f"{K} steps = {ms:.1f} ms: a burst at boost clocks when K is small; see `sustained` for the figure under the power cap"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args.dtype, engine, N),
                         "traffic_unit": "bytes per launch; NOT measured by this run: dram__bytes_read + dram__bytes_write of one launch from the committed "
                                         "ncu capture (profiles/), scaled to this run's particle count",
                         "algorithmic_bytes_per_launch": bpp * N,
                         "peak_source": peak_src, "kernel": "k_step (fused gather+push+BC+deposit)" if engine == "indexed" else "k_push_binned",
                         "kernel_ms": kernel_ms,
                         "kernel_ms_source": ("device-side %globaltimer span (first CTA in, last CTA out) of k_push, averaged over the launches INSIDE the timed "
                                              "graph replay (jic_push_kernel_time)" if in_graph else "CUDA events around the kernel, separate un-graphed pass"),
                         "grid_part_ms": ms / K - kernel_ms if in_graph else ms_grid_step,
                         "ungraphed_cross_check": {"kernel_ms": ms_push_step, "grid_part_ms": ms_grid_step,
                                                   "how": "jic_profile_steps: CUDA events around the particle kernel(s) and the grid part, launches not overlapped"},
                         "algorithmic_bytes": f"{bpp} B per particle-step x {N} particles per launch"},
            "finite": energy_ok,
        }
        if parity is not None:
            line["parity"] = parity
        if sustained is not None:
            sustained["roofline_frac"] = (bpp * N / (sustained["kernel_ms"] * 1e-3) / 1e9 / peak) if sustained["kernel_ms"] else None
            line["sustained"] = sustained
        if f32 is not None:
            if "value" in f32:
                f32["roofline_frac"] = (32 * N / (f32["kernel_ms"] * 1e-3) / 1e9 / peak) if f32["kernel_ms"] else None
            line["extra"] = {"f32": f32}
        if clocks:
            line["clocks"] = clocks
        if e2e:
            line["e2e"] = e2e
        if e2e_sim:
            line["e2e_simulation"] = e2e_sim
        if not args.no_cpu:
            os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every host core again
            rate, cores, sample, _ = cpu_port_rate(w, seconds_target=12.0)
            line["cpu_baseline"] = {"value": rate, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
