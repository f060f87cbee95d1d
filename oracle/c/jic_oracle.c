/*
 * TEST INFRASTRUCTURE ONLY -- C restatement (gcc, OpenMP over particles) of the explicit Boris PIC step of JAX-in-Cell.
 *
 * Only tests/, __graft_entry__.smoke()/build() and bench.py's cpu_baseline / --impl reference legs may build, load or call this.
 * The product path (jax-in-cell_b200/) never does.
 *
 * It is the O(N)+O(G) closed form of oracle/closed_form.py written per particle, same expressions in the same
 * order, so that the checker also runs at the full sizes of BASELINE.json (1e7..1e8 particles) and the CPU baseline is a compiled,
 * multi-threaded implementation rather than NumPy.  Parity status: as oracle/literal.py (pinned by the reference's unit-test vectors
 * and by the reference's own source run on tests/refshim, tests/golden/refsrc_*.npz) -- tests/test_oracle_c.py holds this file to
 * oracle/closed_form.py and to those vectors.  Compiled with -ffp-contract=off: no fused multiply-add, like NumPy.
 *
 * Reference lines (relative to the reference checkout, jaxincell/...):
 *   s2_cloud            _sources.py:43-110      (S2 weights, ghost fold by particle BC)
 *   deposit_current     _sources.py:156-237     (charge-conserving J_x over the 6-node window, J_y,z = rho(x_n) v)
 *   gather              _particles.py:8-45, _boundary_conditions.py:209-247, called as _algorithms.py:40-43
 *   push                _particles.py:68-127 (Boris), :132-200 (relativistic)
 *   particle BCs        _boundary_conditions.py:7-145
 *   filter              _filters.py:9-184
 *   Maxwell half steps  _fields.py:84-193, _boundary_conditions.py:148-207
 *   start-up            _simulation.py:216-225, _state_initialization.py:365-378
 *   step order, outputs _algorithms.py:17-95
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EPS0 8.85418782e-12 /* _constants.py:1-7 */
#define CLIGHT 2.99792458e8
#define MAX_FILTER_PASSES 16

typedef struct jo_params {
  int32_t G;
  int32_t pbl, pbr, fbl, fbr;
  int32_t filter_passes, n_strides, strides[8];
  int32_t relativistic, field_solver, n_threads;
  double L, Ly, Lz, dx, dt, filter_alpha;
} jo_params;

/* numpy's floor_divide for doubles (npy_divmod), which the NumPy oracles use where the reference has `//` */
static double floor_div(double a, double b) {
  {
    /* fast path: away from an integer quotient floor(a/b) is the same number; fmod only decides the ties */
    const double qt = a / b, f = floor(qt), fr = qt - f;
    if (fr > 1e-9 && fr < 1.0 - 1e-9) return f;
  }
  double mod = fmod(a, b), div = (a - mod) / b, fl;
  if (mod != 0.0) {
    if ((b < 0) != (mod < 0)) div -= 1.0;
  }
  if (div != 0.0) {
    fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
  } else {
    fl = copysign(0.0, a / b);
  }
  return fl;
}
/* numpy's float % (sign of the divisor) */
static double floor_mod(double a, double b) {
  if (a > 0.0 && a < b) return a; /* fmod(a, b) == a exactly */
  double mod = fmod(a, b);
  if (mod != 0.0) {
    if ((b < 0) != (mod < 0)) mod += b;
  } else {
    mod = copysign(0.0, b);
  }
  return mod;
}
static inline long clipl(long v, long lo, long hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline long pmod(long a, long n) { long r = a % n; return r < 0 ? r + n : r; }

/* five (node, value) entries: nodes c-1, c, c+1 (value 0 when off the grid) + folded ghost weight on the first / last node */
static void s2_cloud(const jo_params* p, const double* grid, double x, double q, int faces, long node[5], double val[5]) {
  const int G = p->G;
  const double dx = p->dx, L = p->L;
  const double g0 = faces ? grid[0] + dx / 2 : grid[0], gl = faces ? grid[G - 1] + dx / 2 : grid[G - 1];
  const double s = (x - g0) / dx;
  long c = (long)floor(s + 0.5);
  double d;
  if (faces) {
    d = s - (double)c;
  } else {
    const int inside = (x >= -L / 2) && (x <= L / 2);
    if (inside) c = clipl(c, 0, G - 1);
    d = inside ? (x - grid[clipl(c, 0, G - 1)]) / dx : s - (double)c;
  }
  const double w[3] = {0.5 * ((0.5 - d) * (0.5 - d)), 0.75 - d * d, 0.5 * ((0.5 + d) * (0.5 + d))};
  const double qd = q / dx;
  for (int k = 0; k < 3; ++k) {
    const long n = c + k - 1;
    const int ok = n >= 0 && n < G;
    node[k] = ok ? n : 0;
    val[k] = ok ? qd * w[k] : 0.0;
  }
  const double tl = 0.5 + (g0 - x) / dx, tr = 0.5 + (x - gl) / dx;
  const double ex_l = qd * (fabs(x - g0) <= dx / 2 ? 0.5 * (tl * tl) : 0.0);
  const double ex_r = qd * (fabs(x - gl) <= dx / 2 ? 0.5 * (tr * tr) : 0.0);
  node[3] = 0;
  val[3] = p->pbl == 0 ? ex_r : (p->pbl == 1 ? ex_l : 0.0);
  node[4] = G - 1;
  val[4] = p->pbr == 0 ? ex_l : (p->pbr == 1 ? ex_r : 0.0);
}

/* acc layout: [Jx(G) | Jy(G) | Jz(G) | rho(G) | rho_faces(G)] */
static void deposit_current(const jo_params* p, const double* grid, double x_old, double x_mid, double x_new, double vy, double vz, double q,
                            double* acc, int with_rho) {
  const int G = p->G, W = G < 6 ? G : 6;
  const double dx = p->dx, dt = p->dt, gs = grid[0] - dx / 2;
  const long cell = (long)floor_div(x_old - gs, dx);
  double win[6] = {0, 0, 0, 0, 0, 0}, old[6] = {0, 0, 0, 0, 0, 0};
  long node[5];
  double val[5];
  s2_cloud(p, grid, x_new, q, 0, node, val);
  for (int k = 0; k < 5; ++k) {
    const long rel = pmod(node[k] - (cell - 3), G);
    if (rel < W) win[rel] += val[k] / dt;
  }
  s2_cloud(p, grid, x_old, q, 0, node, val);
  for (int k = 0; k < 5; ++k) {
    const long rel = pmod(node[k] - (cell - 3), G);
    if (rel < W) old[rel] += -val[k] / dt;
  }
  double run = 0.0;
  for (int j = 0; j < W; ++j) {
    win[j] += old[j]; /* (sum of the new cloud) + (sum of the old cloud), the order of the NumPy form */
    run += -win[j] * dx;
    acc[pmod(cell - 3 + j, G)] += run;
  }
  s2_cloud(p, grid, x_mid, q, 0, node, val);
  for (int k = 0; k < 5; ++k) {
    acc[G + node[k]] += val[k] * vy;
    acc[2 * G + node[k]] += val[k] * vz;
    if (with_rho) acc[3 * G + node[k]] += val[k]; /* rho(x_{n+1}) of the step outputs: the same cloud (_algorithms.py:86-89) */
  }
}

static void deposit_rho(const jo_params* p, const double* grid, double x, double q, int faces, double* rho) {
  long node[5];
  double val[5];
  s2_cloud(p, grid, x, q, faces, node, val);
  for (int k = 0; k < 5; ++k) rho[node[k]] += val[k];
}

/* padded total field rows [L2, L1, f_0..f_{G-1}, R] x 3 components */
static void pad_field(const jo_params* p, const double* f, const double* ext, double* padded) {
  const int G = p->G;
  for (int i = 0; i < G; ++i)
    for (int c = 0; c < 3; ++c) padded[(i + 2) * 3 + c] = f[i * 3 + c] + ext[i * 3 + c];
  for (int c = 0; c < 3; ++c) {
    const double* t = padded + 2 * 3; /* total field row 0 */
    padded[0 * 3 + c] = p->fbl == 0 ? t[(G - 2) * 3 + c] : (p->fbl == 1 ? t[1 * 3 + c] : 0.0);
    padded[1 * 3 + c] = p->fbl == 0 ? t[(G - 1) * 3 + c] : (p->fbl == 1 ? t[0 * 3 + c] : 0.0);
    padded[(G + 2) * 3 + c] = p->fbr == 0 ? t[0 * 3 + c] : (p->fbr == 1 ? t[(G - 1) * 3 + c] : 0.0);
  }
}

static void gather_tab(const jo_params* p, const double* padded, const double* nodes /* G+1 */, double grid_start, double x, double out[3]) {
  const int G = p->G;
  const double dx = p->dx;
  const long i = (long)floor_div(x - grid_start + dx, dx);
  const double g = nodes[clipl(i, 0, G)];
  const double* f0 = padded + 3 * clipl(i, 0, G + 2);
  const double* f1 = padded + 3 * clipl(i + 1, 0, G + 2);
  const double* f2 = padded + 3 * clipl(i + 2, 0, G + 2);
  const double t = (g - x) / dx, u = (g - x) * (g - x) / (dx * dx);
  for (int c = 0; c < 3; ++c) out[c] = 0.5 * f0[c] * ((0.5 + t) * (0.5 + t)) + f1[c] * (0.75 - u) + 0.5 * f2[c] * ((0.5 - t) * (0.5 - t));
}

static void cross(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

static void push_boris(double dt, const double x[3], const double v[3], double qm, const double E[3], const double B[3], double xo[3], double vo[3]) {
  double vm[3], Rv[3], Bv[3], cr[3], c2[3];
  for (int k = 0; k < 3; ++k) vm[k] = v[k] + qm * E[k] * dt / 2;
  cross(vm, B, cr);
  for (int k = 0; k < 3; ++k) {
    Rv[k] = vm[k] + 0.5 * dt * qm * cr[k];
    Bv[k] = 0.5 * qm * dt * B[k];
  }
  cross(Rv, Bv, c2);
  const double rb = Rv[0] * Bv[0] + Rv[1] * Bv[1] + Rv[2] * Bv[2], bb = Bv[0] * Bv[0] + Bv[1] * Bv[1] + Bv[2] * Bv[2];
  for (int k = 0; k < 3; ++k) {
    const double vp = (c2[k] + rb * Bv[k] + Rv[k]) / (1 + bb);
    vo[k] = vp + qm * E[k] * dt / 2;
    xo[k] = x[k] + dt * vo[k];
  }
}

static void push_relativistic(double dt, const double x[3], const double v[3], double q, double m, const double E[3], const double B[3], double xo[3],
                              double vo[3]) {
  const double c = CLIGHT;
  double s = 0, pm[3], t[3], pxt[3], pn[3];
  for (int k = 0; k < 3; ++k) s += (v[k] / c) * (v[k] / c);
  const double gamma_n = 1 / sqrt(1.0 - s);
  double p2 = 0;
  for (int k = 0; k < 3; ++k) {
    pm[k] = gamma_n * m * v[k] + q * E[k] * dt / 2;
    p2 += pm[k] * pm[k];
  }
  const double gamma_minus = sqrt(1 + p2 / ((m * m) * (c * c)));
  for (int k = 0; k < 3; ++k) t[k] = (q * dt) / (2 * m * gamma_minus) * B[k];
  const double pdt = pm[0] * t[0] + pm[1] * t[1] + pm[2] * t[2], t2 = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
  cross(pm, t, pxt);
  double g2 = 0;
  for (int k = 0; k < 3; ++k) {
    const double pp = (pm[k] * (1 - t2) + 2 * (pdt * t[k] + pxt[k])) / (1 + t2);
    pn[k] = pp + q * E[k] * dt / 2;
    g2 += (pn[k] / (m * c)) * (pn[k] / (m * c));
  }
  const double gamma_new = sqrt(1.0 + g2);
  for (int k = 0; k < 3; ++k) {
    vo[k] = pn[k] / (gamma_new * m);
    xo[k] = x[k] + dt * vo[k];
  }
}

static double bc_x(const jo_params* p, const double* grid, double x) {
  const double L = p->L;
  if (x < -L / 2) return p->pbl == 0 ? floor_mod(x + L / 2, L) - L / 2 : (p->pbl == 1 ? -L - x : grid[0] - 1.5 * p->dx);
  if (x > L / 2) return p->pbr == 0 ? floor_mod(x + L / 2, L) - L / 2 : (p->pbr == 1 ? L - x : grid[p->G - 1] + 3 * p->dx);
  return x;
}
static void bc_positions(const jo_params* p, const double* grid, double x[3]) {
  x[0] = bc_x(p, grid, x[0]);
  x[1] = floor_mod(x[1] + p->Ly / 2, p->Ly) - p->Ly / 2;
  x[2] = floor_mod(x[2] + p->Lz / 2, p->Lz) - p->Lz / 2;
}
/* full BC: position, velocity flip / zero, charge and q/m zeroed on absorption (masses are never modified) */
static void bc_particle(const jo_params* p, const double* grid, double x[3], double v[3], double* q, double* qm) {
  const double L = p->L;
  const int left = x[0] < -L / 2, right = x[0] > L / 2;
  const int bc = left ? p->pbl : (right ? p->pbr : 0);
  bc_positions(p, grid, x);
  if ((left || right) && bc == 1) v[0] = v[0] * -1.0;
  if ((left || right) && bc == 2) {
    v[0] = v[1] = v[2] = 0.0;
    *q = 0.0;
    *qm = 0.0;
  }
}

/* ---- grid side ------------------------------------------------------------------------------------------------------------ */
static void filter_pass(const jo_params* p, const double* y, double* out, int ncomp, double alpha, int s) {
  const int G = p->G;
  const int periodic = p->fbl == 0 && p->fbr == 0;
  for (int j = 0; j < G; ++j) {
    const long jm = j - s, jp = j + s;
    for (int c = 0; c < ncomp; ++c) {
      double l, r;
      if (periodic) {
        l = y[pmod(jm, G) * ncomp + c];
        r = y[pmod(jp, G) * ncomp + c];
      } else {
        l = (jm < 0 && p->fbl == 2) ? 0.0 : y[clipl(jm, 0, G - 1) * ncomp + c];
        r = (jp >= G && p->fbr == 2) ? 0.0 : y[clipl(jp, 0, G - 1) * ncomp + c];
      }
      out[j * ncomp + c] = alpha * y[j * ncomp + c] + (1 - alpha) * 0.5 * (l + r);
    }
  }
}
static void filter_field(const jo_params* p, double* y, int ncomp, double* tmp) {
  if (p->filter_passes <= 0) return;
  const int passes = p->filter_passes, clamped = passes < MAX_FILTER_PASSES + 1 ? passes : MAX_FILTER_PASSES + 1;
  int regular = passes - 1 > 0 ? passes - 1 : 0;
  if (regular > MAX_FILTER_PASSES) regular = MAX_FILTER_PASSES;
  const double comp = clamped - p->filter_alpha * (clamped - 1);
  const size_t bytes = (size_t)p->G * ncomp * sizeof(double);
  for (int si = 0; si < p->n_strides; ++si) {
    for (int k = 0; k < regular; ++k) {
      filter_pass(p, y, tmp, ncomp, p->filter_alpha, p->strides[si]);
      memcpy(y, tmp, bytes);
    }
    filter_pass(p, y, tmp, ncomp, comp, p->strides[si]);
    memcpy(y, tmp, bytes);
  }
}

static void curl_E(const jo_params* p, const double* E, const double* B, double* out) {
  const int G = p->G;
  const double c = CLIGHT, dx = p->dx;
  double gl[3];
  if (p->fbl == 0) memcpy(gl, E + 3 * (G - 1), sizeof(gl));
  else if (p->fbl == 1) memcpy(gl, E, sizeof(gl));
  else { gl[0] = 0.0; gl[1] = -2 * c * B[2] - E[1]; gl[2] = 2 * c * B[1] - E[2]; }
  for (int i = 0; i < G; ++i) {
    const double* prev = i ? E + 3 * (i - 1) : gl;
    const double dFz = (E[3 * i + 2] - prev[2]) / dx, dFy = (E[3 * i + 1] - prev[1]) / dx;
    out[3 * i] = 0.0;
    out[3 * i + 1] = -dFz;
    out[3 * i + 2] = dFy;
  }
}
static void curl_B(const jo_params* p, const double* B, const double* E, double* out) {
  const int G = p->G;
  const double c = CLIGHT, dx = p->dx;
  double gr[3];
  if (p->fbr == 0) memcpy(gr, B, sizeof(gr));
  else if (p->fbr == 1) memcpy(gr, B + 3 * (G - 1), sizeof(gr));
  else { gr[0] = 0.0; gr[1] = -(2 / c) * E[3 * (G - 1) + 2] - B[3 * (G - 1) + 1]; gr[2] = (2 / c) * E[3 * (G - 1) + 1] - B[3 * (G - 1) + 2]; }
  for (int i = 0; i < G; ++i) {
    const double* next = i + 1 < G ? B + 3 * (i + 1) : gr;
    const double dFz = (next[2] - B[3 * i + 2]) / dx, dFy = (next[1] - B[3 * i + 1]) / dx;
    out[3 * i] = 0.0;
    out[3 * i + 1] = -dFz;
    out[3 * i + 2] = dFy;
  }
}
static void ampere(const jo_params* p, double* E, const double* B, const double* J, double h, double* tmp) {
  curl_B(p, B, E, tmp);
  for (int k = 0; k < 3 * p->G; ++k) E[k] = E[k] + h * ((CLIGHT * CLIGHT) * tmp[k] - (J[k] / EPS0));
}
static void faraday(const jo_params* p, const double* E, double* B, double h, double* tmp) {
  curl_E(p, E, B, tmp);
  for (int k = 0; k < 3 * p->G; ++k) B[k] = B[k] - h * tmp[k];
}

/* sum the thread-private raw grids in thread order */
static void reduce(double* dst, const double* priv, int n_threads, size_t len) {
  memset(dst, 0, len * sizeof(double));
  for (int t = 0; t < n_threads; ++t)
    for (size_t k = 0; k < len; ++k) dst[k] += priv[(size_t)t * len + k];
}

int jo_abi_version(void) { return 2; }

static double wall_seconds(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/*
 * x0, v0: (N,3); q, m, qm: (N,) weight-scaled charge / mass and q/m; grid: (G,) = numpy.linspace(-L/2+dx/2, L/2-dx/2, G);
 * ext_E, ext_B: (G,3) already widened from float32; gauss_h: (G,) circulant kernel (field_solver 1, 3) or NULL.
 * Histories: (T,G,3) x3, (T,G); hist_x / hist_v (T,N,3) or NULL.  E0, B0 (G,3), v_init (N,3) or NULL.
 * x_half_out, v_out (N,3) or NULL: the carried state after the last step.  step_seconds (T) or NULL: wall-clock seconds of each step
 * (bench.py's CPU legs).  Returns 0, or -1 on allocation failure.
 */
int jo_run(const jo_params* p, int64_t N, const double* x0, const double* v0, const double* q_in, const double* m, const double* qm_in,
           const double* grid, const double* ext_E, const double* ext_B, const double* gauss_h, int64_t T, double* hist_E, double* hist_B,
           double* hist_J, double* hist_rho, double* hist_x, double* hist_v, double* E0, double* B0, double* v_init, double* x_half_out,
           double* v_out, double* step_seconds /* (T) wall-clock of every step, or NULL */) {
  const int G = p->G;
  const double dx = p->dx, dt = p->dt;
  int n_threads = p->n_threads > 0 ? p->n_threads : 1;
#ifndef _OPENMP
  n_threads = 1;
#endif
  const size_t ACC = (size_t)5 * G;
  double* xh = malloc(sizeof(double) * 3 * (N ? N : 1));
  double* xn = malloc(sizeof(double) * 3 * (N ? N : 1));
  double* v = malloc(sizeof(double) * 3 * (N ? N : 1));
  double* q = malloc(sizeof(double) * (N ? N : 1));
  double* qm = malloc(sizeof(double) * (N ? N : 1));
  double* priv = calloc((size_t)n_threads * ACC, sizeof(double));
  double* acc = calloc(ACC, sizeof(double));
  double* E = calloc((size_t)3 * G, sizeof(double));
  double* B = calloc((size_t)3 * G, sizeof(double));
  double* J = calloc((size_t)3 * G, sizeof(double));
  double* rho = calloc(G, sizeof(double));
  double* tmp = calloc((size_t)3 * G, sizeof(double));
  double* padE = calloc((size_t)3 * (G + 3), sizeof(double));
  double* padB = calloc((size_t)3 * (G + 3), sizeof(double));
  double* nodesE = malloc(sizeof(double) * (G + 1));
  double* nodesB = malloc(sizeof(double) * (G + 1));
  double* zero3 = calloc((size_t)3 * G, sizeof(double));
  int rc = -1;
  if (!xh || !xn || !v || !q || !qm || !priv || !acc || !E || !B || !J || !rho || !tmp || !padE || !padB || !nodesE || !nodesB || !zero3) goto done;
  if (!ext_E) ext_E = zero3;
  if (!ext_B) ext_B = zero3;
  /* node tables of the gather: E on grid + dx/2 with grid_start = grid[0]; B on grid with grid_start = grid[0] - dx/2 (_algorithms.py:41-42) */
  nodesE[0] = (grid[0] + dx / 2) - dx;
  nodesB[0] = grid[0] - dx;
  for (int k = 0; k < G; ++k) {
    nodesE[k + 1] = grid[k] + dx / 2;
    nodesB[k + 1] = grid[k];
  }

  /* ---- start-up: rho_0 -> E_x, x_{+1/2} with the full BC, x_{-1/2} with the post-BC velocity, first J ---- */
#pragma omp parallel num_threads(n_threads)
  {
#ifdef _OPENMP
    double* a = priv + (size_t)omp_get_thread_num() * ACC;
#else
    double* a = priv;
#endif
    memset(a, 0, ACC * sizeof(double));
#pragma omp for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
      deposit_rho(p, grid, x0[3 * i], q_in[i], 0, a + 3 * G);
      double xp[3], vv[3], xm[3], qi = q_in[i], qmi = qm_in[i];
      for (int k = 0; k < 3; ++k) {
        xp[k] = x0[3 * i + k] + (dt / 2) * v0[3 * i + k];
        vv[k] = v0[3 * i + k];
      }
      bc_particle(p, grid, xp, vv, &qi, &qmi);
      for (int k = 0; k < 3; ++k) xm[k] = x0[3 * i + k] - (dt / 2) * vv[k];
      bc_positions(p, grid, xm);
      deposit_current(p, grid, xm[0], x0[3 * i], xp[0], vv[1], vv[2], qi, a, 0);
      for (int k = 0; k < 3; ++k) {
        xh[3 * i + k] = xp[k];
        xn[3 * i + k] = x0[3 * i + k];
        v[3 * i + k] = vv[k];
      }
      q[i] = qi;
      qm[i] = qmi;
      if (v_init) for (int k = 0; k < 3; ++k) v_init[3 * i + k] = vv[k];
    }
  }
  reduce(acc, priv, n_threads, ACC);
  memcpy(rho, acc + 3 * G, sizeof(double) * G);
  filter_field(p, rho, 1, tmp);
  {
    double run = 0.0;
    for (int i = 0; i < G; ++i) {
      run += rho[i];
      E[3 * i] = (dx / EPS0) * run;
    }
  }
  if (E0) memcpy(E0, E, sizeof(double) * 3 * G);
  if (B0) memcpy(B0, B, sizeof(double) * 3 * G);
  for (int i = 0; i < G; ++i)
    for (int c = 0; c < 3; ++c) J[3 * i + c] = acc[c * G + i];
  filter_field(p, J, 3, tmp);

  /* ---- steps ---- */
  for (int64_t t = 0; t < T; ++t) {
    const double t_start = wall_seconds();
    ampere(p, E, B, J, dt / 2, tmp);   /* field_update1: E then B */
    faraday(p, E, B, dt / 2, tmp);
    pad_field(p, E, ext_E, padE);
    pad_field(p, B, ext_B, padB);
    double* hx = hist_x ? hist_x + (size_t)t * 3 * N : NULL;
    double* hv = hist_v ? hist_v + (size_t)t * 3 * N : NULL;
#pragma omp parallel num_threads(n_threads)
    {
#ifdef _OPENMP
      double* a = priv + (size_t)omp_get_thread_num() * ACC;
#else
      double* a = priv;
#endif
      memset(a, 0, ACC * sizeof(double));
#pragma omp for schedule(static)
      for (int64_t i = 0; i < N; ++i) {
        double Ep[3], Bp[3], xpp[3], vn[3], xnew[3];
        const double* x = xh + 3 * i;
        gather_tab(p, padE, nodesE, grid[0], x[0], Ep);
        gather_tab(p, padB, nodesB, grid[0] - dx / 2, x[0], Bp);
        if (p->relativistic) push_relativistic(dt, x, v + 3 * i, q[i], m[i], Ep, Bp, xpp, vn);
        else push_boris(dt, x, v + 3 * i, qm[i], Ep, Bp, xpp, vn);
        double qi = q[i], qmi = qm[i];
        bc_particle(p, grid, xpp, vn, &qi, &qmi);
        for (int k = 0; k < 3; ++k) xnew[k] = xpp[k] - (dt / 2) * vn[k];
        bc_positions(p, grid, xnew);
        deposit_current(p, grid, x[0], xnew[0], xpp[0], vn[1], vn[2], qi, a, 1);
        if (p->field_solver) deposit_rho(p, grid, xn[3 * i], qi, 1, a + 4 * G); /* rho(x_n) on the faces, post-BC charge (_algorithms.py:69-72) */
        for (int k = 0; k < 3; ++k) {
          xh[3 * i + k] = xpp[k];
          xn[3 * i + k] = xnew[k];
          v[3 * i + k] = vn[k];
          if (hx) hx[3 * i + k] = xnew[k];
          if (hv) hv[3 * i + k] = vn[k];
        }
        q[i] = qi;
        qm[i] = qmi;
      }
    }
    reduce(acc, priv, n_threads, ACC);
    for (int i = 0; i < G; ++i)
      for (int c = 0; c < 3; ++c) J[3 * i + c] = acc[c * G + i];
    filter_field(p, J, 3, tmp);
    faraday(p, E, B, dt / 2, tmp);     /* field_update2: B then E */
    ampere(p, E, B, J, dt / 2, tmp);
    if (p->field_solver) {
      double* rf = acc + 4 * G;
      filter_field(p, rf, 1, tmp);
      if (p->field_solver == 2) {
        double run = 0.0;
        for (int i = 0; i < G; ++i) {
          run += rf[i];
          E[3 * i] = (dx / EPS0) * run;
        }
      } else {
        for (int i = 0; i < G; ++i) {
          double s = 0.0;
          for (int j = 0; j < G; ++j) s += gauss_h[pmod(i - j, G)] * rf[j];
          E[3 * i] = s;
        }
      }
    }
    memcpy(rho, acc + 3 * G, sizeof(double) * G);
    filter_field(p, rho, 1, tmp);
    if (hist_E) memcpy(hist_E + (size_t)t * 3 * G, E, sizeof(double) * 3 * G);
    if (hist_B) memcpy(hist_B + (size_t)t * 3 * G, B, sizeof(double) * 3 * G);
    if (hist_J) memcpy(hist_J + (size_t)t * 3 * G, J, sizeof(double) * 3 * G);
    if (hist_rho) memcpy(hist_rho + (size_t)t * G, rho, sizeof(double) * G);
    if (step_seconds) step_seconds[t] = wall_seconds() - t_start;
  }
  if (x_half_out) memcpy(x_half_out, xh, sizeof(double) * 3 * N);
  if (v_out) memcpy(v_out, v, sizeof(double) * 3 * N);
  rc = 0;
done:
  free(xh); free(xn); free(v); free(q); free(qm); free(priv); free(acc); free(E); free(B); free(J); free(rho); free(tmp);
  free(padE); free(padB); free(nodesE); free(nodesB); free(zero3);
  return rc;
}
