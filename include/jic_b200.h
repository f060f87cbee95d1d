/*
 * jic_b200.h -- C ABI of the B200-native particle hot path that replaces JAX-in-Cell's explicit (Boris) time loop.
 *
 * The reference (uwplasma/JAX-in-Cell) has no FFI of its own: its hot path is Python traced under one jax.jit.
 * The seam this library plugs into is the scan at jaxincell/_simulation.py:228-257 (initial carry -> lax.scan of
 * jaxincell/_algorithms.py:17-95 `Boris_step` -> six stacked histories).  Each entry point below names the reference
 * code it replaces.  Everything is plain C: opaque handle, POD structs, raw pointers, sizes, `void*` CUDA streams.
 *
 * Conventions
 *   - return value: 0 = JIC_OK, negative = error; text via jic_last_error(ctx) (or jic_last_error(NULL) for
 *     failures of jic_create / jic_simulate_host).  Nothing throws across this boundary.
 *   - "real" = double when params.dtype == JIC_F64 (the reference's only precision, _simulation.py:32), float for JIC_F32.
 *   - device pointers unless the function name ends in _host.  Buffers are owned by the caller; the library keeps
 *     no reference to them after the call returns, except the history pointers of jic_run, which must stay valid
 *     until the stream has drained.
 *   - all work is enqueued on the given stream; no entry point except *_host, jic_create, jic_comm_init and
 *     jic_destroy synchronises.  Scratch memory, the NCCL communicator and CUDA graphs belong to the context.
 *   - a context is bound to one device and must not be used from two host threads at once; distinct contexts are
 *     independent (no global mutable state).
 *   - array layouts are the reference's: particles (N,3) row-major, fields (G,3) row-major, rho (G,).
 */
#ifndef JIC_B200_H
#define JIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JIC_ABI_VERSION 1
#define JIC_MAX_SPECIES 8
#define JIC_MAX_STRIDES 8

enum jic_status {
  JIC_OK = 0,
  JIC_ERR_INVALID_ARGUMENT = -1,
  JIC_ERR_CUDA = -2,
  JIC_ERR_BAD_STATE = -3,
  JIC_ERR_NCCL = -4,
  JIC_ERR_UNSUPPORTED = -5
};

enum jic_dtype { JIC_F64 = 0, JIC_F32 = 1 };

/* Particle store.
 *   INDEXED keeps particle p in slot p for the whole run (needed for the reference's (T,N,3) histories); one thread per
 *           particle, deposition through global atomics or a CTA-private shared-memory copy of the grid.
 *   BINNED  keeps particles binned by (species, cell) and re-bins them inside the push kernel every step, so that
 *           gather coefficients and deposition stencils are CTA-uniform and accumulate in registers (the fast path
 *           for >= ~1e6 particles).  Particle order is not preserved; particle histories are not available. */
enum jic_engine { JIC_ENGINE_INDEXED = 0, JIC_ENGINE_BINNED = 1 };

/* Boundary codes of jaxincell/_parameters/_domain_parameters.py:53-56 */
enum jic_bc { JIC_BC_PERIODIC = 0, JIC_BC_REFLECTIVE = 1, JIC_BC_ABSORBING = 2 };

enum jic_deposit { JIC_DEPOSIT_AUTO = 0, JIC_DEPOSIT_GLOBAL_ATOMICS = 1, JIC_DEPOSIT_SHARED_GRID = 2 };

/* One contiguous block of identical macro-particles (jaxincell/_state_initialization.py:242-261:
 * species are concatenated block-wise; charge and mass are already multiplied by the weight, q/m is not). */
typedef struct jic_species {
  int64_t count;          /* particles of this species ON THIS RANK */
  double charge;          /* q_s * w   [C]  */
  double mass;            /* m_s * w   [kg] */
  double charge_to_mass;  /* q_s / m_s      */
} jic_species;

/* Static description of a run: what build_domain_state (jaxincell/_state_initialization.py:27-49), the four BC
 * integers (_simulation.py:208-211) and solver_parameters (filter, relativistic) hand to Boris_step. */
typedef struct jic_params {
  uint32_t struct_bytes;  /* = sizeof(jic_params); checked */
  int32_t dtype;          /* jic_dtype */
  int32_t engine;         /* jic_engine */
  int32_t device;         /* CUDA ordinal, -1 = current device */
  int32_t n_grid;         /* G */
  int32_t n_species;      /* <= JIC_MAX_SPECIES */
  double length, length_y, length_z; /* box_size */
  double dx, dt;          /* as the host computed them: dx = L/G, dt = CFL*dx/c */
  double grid_first, grid_last; /* grid[0], grid[-1] of linspace(-L/2+dx/2, L/2-dx/2, G): edge tests use these bits */
  int32_t particle_bc_left, particle_bc_right, field_bc_left, field_bc_right;
  int32_t filter_passes;  /* >= 0 */
  int32_t n_filter_strides;
  int32_t filter_strides[JIC_MAX_STRIDES];
  double filter_alpha;
  int32_t relativistic;   /* solver_parameters["relativistic"] */
  int32_t track_yz;       /* also advance and wrap y,z (only needed for the full (N,3) position outputs) */
  int32_t deposit;        /* jic_deposit (INDEXED engine) */
  int32_t steps_per_graph;/* steps captured per CUDA graph, 0 = default */
  int32_t field_solver;   /* per-step electrostatic correction of jaxincell/_algorithms.py:69-78 (the `field_solver` argument of
                           * Boris_step): 0 = none, 1 = E_from_Gauss_1D_FFT, 2 = E_from_Gauss_1D_Cartesian, 3 = E_from_Poisson_1D_FFT
                           * (_fields.py:9-81).  E_x is replaced every step by the solve of rho(x_n) deposited on the faces. */
  int32_t time_evolution_algorithm; /* solver_parameters["time_evolution_algorithm"]: 0 = explicit Boris_step (the hot path),
                           * 1 = implicit Crank-Nicolson CN_step (jaxincell/_algorithms.py:100-241): Picard iterations over Faraday,
                           * a sub-stepped push and Ampere with J - <J>.  Periodic S2 gather / deposit, no filter, no external
                           * fields, non-relativistic, like the reference.  Particle arrays are kept in input order. */
  int32_t cn_substeps;    /* number_of_particle_substeps_implicit_CN (>= 1) */
  int32_t cn_max_iterations; /* max_number_of_Picard_iterations_implicit_CN (>= 1) */
  double cn_tolerance;    /* tolerance_Picard_iterations_implicit_CN */
  int32_t reserved[2];    /* must be zero */
} jic_params;

/* Where jic_run writes the per-step outputs (jaxincell/_algorithms.py:93, stacked at _simulation.py:256-257).
 * Row t of every history is the state after step t of THIS call.  NULL = not recorded. */
typedef struct jic_outputs {
  void* electric_field;   /* real (T,G,3) */
  void* magnetic_field;   /* real (T,G,3) */
  void* current_density;  /* real (T,G,3) */
  void* charge_density;   /* real (T,G)   */
  void* positions;        /* real (T,N,3), INDEXED engine with track_yz only */
  void* velocities;       /* real (T,N,3), INDEXED engine only */
  void* kinetic_energy;   /* double (T,n_species): sum over the LOCAL particles of a species of m v^2 / 2 after the step -- the kinetic
                           * energies of jaxincell/_diagnostics.py:131-138 evaluated on the device, for runs whose (T,N,3) velocity
                           * history would not fit (any engine; one more pass over the velocities per step when requested) */
} jic_outputs;

typedef struct jic_context jic_context;

int jic_abi_version(void);
const char* jic_last_error(const jic_context* ctx);

/* Allocate all device state for `params` (replaces the carry construction at _simulation.py:228-231). */
int jic_create(const jic_params* params, const jic_species* species, jic_context** out);
int jic_destroy(jic_context* ctx);

/* Multi-GPU: particles are sharded across ranks, fields replicated; one all-reduce of the raw [Jx,Jy,Jz,rho] grid per
 * step (NCCL, or fused into the field kernel over peer memory -- see jic_comm_mode).  Rank 0 creates an id (128 bytes), the host exchanges it (torch.distributed / MPI / files), every rank calls
 * jic_comm_init before jic_initialize.  No reference counterpart (the reference is single-device). */
int jic_comm_unique_id(void* id_128_bytes);
int jic_comm_init(jic_context* ctx, const void* id_128_bytes, int rank, int world_size);
/* How the per-step reduction of the raw grid runs: 0 = single rank, 1 = NCCL all-reduce in front of the field kernel,
 * 2 = fused into the field kernel (every rank reads the other ranks' raw grids over NVLink through CUDA-IPC mappings and sums
 * them in rank order while it loads its window; chosen by jic_comm_init when all ranks are processes of one host and every
 * mapping succeeds; JIC_P2P=0 in the environment forces 1). */
int jic_comm_mode(const jic_context* ctx);

/* External fields, float32 (G,3) as the reference stores them (_state_initialization.py:382-392); NULL = zeros. */
int jic_set_external_fields(jic_context* ctx, const float* external_E, const float* external_B, void* stream);

/* Initial particles at t=0: x0, v0 real (N,3).  Performs the leap-frog start-up of _simulation.py:217-225
 * (x_{+1/2} with the full particle BC, x_{-1/2} with the post-BC velocity), the initial charge deposit and
 * Gauss solve of _state_initialization.py:374-378, and the first current deposit of _algorithms.py:29-32. */
int jic_initialize(jic_context* ctx, const void* x0, const void* v0, void* stream);

/* The same with HOST buffers (pinned memory makes the copies asynchronous): the upload is cut into chunks that alternate between
 * two device staging buffers and the start-up kernel of a chunk runs while the next chunk is in flight, so the device never
 * holds a full copy of x0, v0 and the start-up work hides behind the PCIe transfer.  The host buffers must stay valid until the
 * stream has drained.  JIC_HOST_CHUNK=<particles> overrides the chunk size (default 2^23). */
int jic_initialize_host(jic_context* ctx, const void* x0_host, const void* v0_host, void* stream);

/* Step granularity: load the reference's scan carry instead of starting from t = 0.  The carry of jaxincell/_simulation.py:228-231 /
 * _algorithms.py:23-24 is (E^n, B^n, x_{n-1/2}, x_n, x_{n+1/2}, v_n, q, m, q/m): E, B real (G,3) at integer time, the four particle
 * arrays real (N,3) in the order of the species table.  Charges are the species values of jic_create; particles the reference has
 * absorbed carry q = 0 and sit parked outside the box, which is how this library recognises them too.  After the call
 * jic_run(ctx, 1, outputs) is exactly one Boris_step (_algorithms.py:17-95): `outputs` rows are its step_data, and the new carry is
 * (jic_get_fields E, B; the x_{n+1/2} passed in; outputs->positions; jic_get_particles x_half, v).  INDEXED engine, explicit stepper. */
int jic_load_carry(jic_context* ctx, const void* E, const void* B, const void* x_minus_half, const void* x_n, const void* x_plus_half,
                   const void* v_n, void* stream);

/* The same for the implicit stepper (time_evolution_algorithm = 1): the carry of CN_step is (E^n, B^n, x_n, v_n, q, m, q/m)
 * (jaxincell/_simulation.py:237-240, _algorithms.py:103-104).  alive[N] (uint8, NULL = all ones) is 0 where the carry's charge is zero
 * (particles absorbed by the start-up half step: their charge stays zero for the whole run).  jic_run(ctx, 1, outputs) is then one
 * CN_step; the new carry is (jic_get_fields E, B; outputs->positions; outputs->velocities; q, m, q/m unchanged). */
int jic_load_carry_cn(jic_context* ctx, const void* E, const void* B, const void* x_n, const void* v_n, const uint8_t* alive, void* stream);

/* Advance n_steps (the lax.scan of _simulation.py:253 over Boris_step).  Captured as CUDA graphs; no host sync. */
int jic_run(jic_context* ctx, int64_t n_steps, const jic_outputs* outputs, void* stream);

/* Current grid state: E, B (G,3) at integer time, filtered J (G,3) and rho (G,) of the last step.  Any may be NULL. */
int jic_get_fields(jic_context* ctx, void* E, void* B, void* J, void* rho, void* stream);
/* Initial fields (output key "fields") and post-BC initial velocities (key "initial_velocities", INDEXED only). */
int jic_get_initial(jic_context* ctx, void* E0, void* B0, void* initial_velocities, void* stream);
/* Current particles: x_{n+1/2} (what the pusher carries) and v_n, real (N,3); y,z are zero unless track_yz.
 * BINNED engine: bin order, not input order.  alive[N] (uint8, optional) is 0 for absorbed particles. */
int jic_get_particles(jic_context* ctx, void* x_half, void* v, uint8_t* alive, void* stream);
/* Sum over local particles of 0.5*m*v^2 (jaxincell/_diagnostics.py:131-138) into *kinetic_energy (device double). */
int jic_kinetic_energy(jic_context* ctx, double* kinetic_energy, void* stream);
/* Measurement aid: advance n_steps WITHOUT graphs or histories, bracketing the particle kernel(s) and the grid part
 * (all-reduce + field kernel) of every step with CUDA events on `stream`; returns the summed milliseconds of each.
 * Synchronises the stream. */
int jic_profile_steps(jic_context* ctx, int64_t n_steps, double* ms_particle_kernels, double* ms_grid_kernels, void* stream);
/* Crank-Nicolson only: Picard iterations of the last completed step and of all steps so far.  Synchronises the stream. */
int jic_get_picard_iterations(jic_context* ctx, int64_t* last_step, int64_t* total, void* stream);
/* Synchronises the stream and reports the sticky device-side error flags of the context: exhausted capacity of the BINNED store
 * (a bin grew faster than its head-room and the overflow list filled up: particles were dropped) and a peer rank that missed the
 * fused reduction's barrier.  jic_run only enqueues work, so a caller that reads the histories of a single jic_run must call this
 * (jic_simulate_host does) -- otherwise the flags are only seen by the NEXT jic_run / jic_get_particles.
 * The reference has no counterpart: its arrays cannot overflow (jaxincell/_simulation.py:228-257 scans fixed-shape carries). */
int jic_check_status(jic_context* ctx, void* stream);
/* Measurement aid, BINNED engine: device-side duration of the push kernel (k_push), summed over its launches since the last reset.
 * Every CTA stamps %globaltimer on entry and exit; the span first-in .. last-out of a launch is accumulated on the device, so the
 * figure is valid for launches inside CUDA-graph replays (the same replays jic_run's throughput is measured on), where events around
 * single kernels are not available.  Synchronises the stream.  reset != 0 zeroes the sum afterwards. */
int jic_push_kernel_time(jic_context* ctx, double* ms_sum, int64_t* n_launches, int32_t reset, void* stream);
/* Diagnostic aid, BINNED engine: counters of the particle store after the work queued so far (synchronises the stream):
 * out[0] work items of the next push, out[1] / out[2] entries in the overflow lists of the two buffers, out[3] sticky error flag
 * (0 ok, 1 overflow or general-path list full, 2 slot capacity exhausted), out[4] entries in the general-path list of the last step,
 * out[5] slots in use (holes included), out[6] slots per buffer, out[7] particles absorbed so far.
 * Crank-Nicolson contexts (no binned store): out[0] = 1 when the context runs the cell-sorted push (csrc/jic_cn_sorted.cuh; contexts of
 * at least JIC_CN_SORTED_MIN particles, an environment variable read at creation, default 200000), 0 for the unsorted one; the rest 0. */
int jic_store_stats(jic_context* ctx, int64_t out[8], void* stream);
/* The large device buffers of a context (>= 32 MiB each: particle stores, per-particle arrays, upload staging) come from a stream-ordered
 * memory pool the library owns, one per device, which keeps freed memory so that the next context of the process does not pay
 * cudaMalloc / cudaFree again (28 ms per 1e8-particle context).  jic_trim_memory hands everything the pools hold and nobody uses back to
 * the driver.  Environment JIC_POOL=0: no pool, plain cudaMalloc / cudaFree. */
void jic_trim_memory(void);
/* Number of kernel launches issued by this context so far (for the bench's gpu_launches figure). */
int64_t jic_launch_count(const jic_context* ctx);

/* Initial particles generated ON THE DEVICE with the reference's formulas and jax.random's streams
 * (jaxincell/_state_initialization.py:51-85 `initialize_species_phase_space`, seeds from `species_seed_pair` :87-96, the
 * 0.99c clip of :259-260).  One entry per species, in the order the reference concatenates them (electrons, then ions). */
typedef struct jic_species_sampling {
  int64_t count;                      /* number_pseudoparticles */
  int64_t seed_position, seed_velocity; /* axis a uses PRNGKey(seed_position + a + 1) and PRNGKey(seed_velocity + a + 4) */
  int32_t random_positions[3];        /* random_positions_{x,y,z}: uniform in the box, else linspace(-L/2, L/2, count) */
  int32_t velocity_plus_minus[3];     /* multiply by (-1)**arange(count) */
  double perturbation_amplitude[3];
  double perturbation_wavenumber[3];  /* as in the input: multiplied by 2 pi / box length inside */
  double vth_over_c[3];
  double drift_speed[3];
} jic_species_sampling;
/* x0, v0: device, real (sum of counts, 3).  threefry_partitionable: 1 = jax >= 0.5 default bit layout, 0 = the original one. */
int jic_sample_particles(int32_t dtype, int32_t n_species, const jic_species_sampling* species, const double box_size[3],
                         int32_t threefry_partitionable, void* x0, void* v0, void* stream);
/* The same streams, but only particles [first[s], first[s] + local_count[s]) of every species s: what one rank of an index-sharded
 * run owns.  jax.random's Threefry is counter based -- element i of a draw depends on (key, i, n) only -- so a rank
 * samples its slice without generating or receiving anybody else's.  x0, v0: device, real (sum of local counts, 3). */
int jic_sample_particles_slice(int32_t dtype, int32_t n_species, const jic_species_sampling* species, const int64_t* first,
                               const int64_t* local_count, const double box_size[3], int32_t threefry_partitionable, void* x0, void* v0,
                               void* stream);

/* Whole-simulation call with HOST buffers (the drop-in for Simulation.run's device part, _simulation.py:169-257):
 * host->device copies of x0, v0 and the external fields, jic_initialize, jic_run, device->host copies of the
 * requested histories and a final synchronise are all inside.  `host_outputs` members point to host memory. */
int jic_simulate_host(const jic_params* params, const jic_species* species, const void* x0_host, const void* v0_host,
                      const float* external_E_host, const float* external_B_host, int64_t n_steps,
                      const jic_outputs* host_outputs, void* fields0_E_host, void* fields0_B_host,
                      void* initial_velocities_host);

#ifdef __cplusplus
}
#endif
#endif /* JIC_B200_H */
