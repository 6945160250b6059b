/*
 * CPU restatement (plain C, fp64) of the reference's linearized contact subproblem solve.
 *
 * TEST INFRASTRUCTURE ONLY: this is the checker / CPU baseline ("kind": "port" in bench.py's
 * cpu_baseline).  Nothing on the product path links or calls it.
 * PARITY UNPINNED at the iterate level: the IP loop is RoboDojo 0.1.3's (not under /root/reference),
 * restated from SURVEY.md §3.4 exactly as in oracle/ip.py; see that file's header.
 *
 * Follows the reference function by function, keeping its algorithmic choices (explicit Dx⁻¹,
 * modified Gram-Schmidt QR of the ny×ny Schur complement at every iteration, all nθ sensitivity
 * columns), so that its timing is a fair stand-in for the Julia CPU path:
 *   RLin/RZLin/RθLin slicing          src/controller/linearized_solver.jl:92-120, 245-251, 346-348
 *   Schur ctor (Ai, CAi, CAiB)        src/solver/schur.jl:33-49
 *   rlin!                             src/controller/linearized_solver.jl:364-373
 *   rzlin! / schur_factorize!         linearized_solver.jl:378-399, schur.jl:80-88
 *   factorize!(SDMGSSolver)           src/solver/qr.jl:113-137
 *   qr_solve!                         src/solver/qr.jl:142-158
 *   schur_solve!                      src/solver/schur.jl:93-110
 *   linear_solve!(Δ) / (δz)           linearized_solver.jl:424-444, 451-479
 *   residual_/bilinear_violation      linearized_solver.jl:401-409
 *   general_correction_term!          linearized_solver.jl:411-418
 *   implicit_dynamics! cold start     src/controller/implicit_dynamics.jl:160-164, src/simulation/simulation.jl:59-63
 * Batch loop: OpenMP `parallel for` over independent subproblems (the reference's optional
 * `Threads.@threads`, implicit_dynamics.jl:166-171).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 32   /* max nx, ny   */
#define MAXT 64   /* max nθ       */

typedef struct {
  int nq, nu, nw, nc, nb, mode;
  int nx, ny, nz, nth, nd, ncol;
} dims_t;

typedef struct {
  double r_tol, kappa_tol, eps_min, kappa_reg, gamma_reg, undercut, ls_scale;
  int32_t max_iter, max_ls, diff_sol, reserved;
} ip_opts_t;

/* per-knot constants (RLin + RZLin + RθLin + Schur) */
typedef struct {
  double Dx[MAXN * MAXN], Dy1[MAXN * MAXN], Rx[MAXN * MAXN], Ry1[MAXN * MAXN], Ry2[MAXN];
  double rthd[MAXN * MAXT], rthr[MAXN * MAXT];
  double rdyn0[MAXN], rrst0[MAXN], x0[MAXN], y10[MAXN], y20[MAXN], th0[MAXT];
  double Ai[MAXN * MAXN], CAi[MAXN * MAXN], CAiB[MAXN * MAXN];
} knot_t;

typedef struct {
  dims_t d;
  int H;
  int use_lu; /* 0: reference MGS-QR (default); 1: LU with partial pivoting (accurate test variant) */
  knot_t* k;
} oracle_t;

static void set_dims(dims_t* d, int nq, int nu, int nw, int nc, int nb, int mode) {
  d->nq = nq; d->nu = nu; d->nw = nw; d->nc = nc; d->nb = nb; d->mode = mode;
  d->nx = nq; d->ny = 2 * nc + nb; d->nz = nq + 4 * nc + 2 * nb; d->nth = 2 * nq + nu + nw + 2;
  d->nd = mode ? nq + nc + nb : nq; d->ncol = 2 * nq + nu;
}

/* dense inverse by LU with partial pivoting (Julia `inv` → LAPACK getrf/getri) */
static int invert(int n, const double* A, double* Ai) {
  double M[MAXN * 2 * MAXN];
  const int w = 2 * n;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < w; ++j) M[i * w + j] = (j < n) ? A[i + j * n] : (j - n == i ? 1.0 : 0.0);
  for (int k = 0; k < n; ++k) {
    int p = k;
    for (int i = k + 1; i < n; ++i)
      if (fabs(M[i * w + k]) > fabs(M[p * w + k])) p = i;
    if (M[p * w + k] == 0.0) return 1;
    if (p != k)
      for (int j = 0; j < w; ++j) { double t = M[k * w + j]; M[k * w + j] = M[p * w + j]; M[p * w + j] = t; }
    const double inv = 1.0 / M[k * w + k];
    for (int j = 0; j < w; ++j) M[k * w + j] *= inv;
    for (int i = 0; i < n; ++i)
      if (i != k) {
        const double f = M[i * w + k];
        if (f != 0.0)
          for (int j = 0; j < w; ++j) M[i * w + j] -= f * M[k * w + j];
      }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Ai[i + j * n] = M[i * w + n + j];
  return 0;
}

/* C(m×n) = A(m×k) B(k×n), column-major */
static void matmul(int m, int k, int n, const double* A, const double* B, double* C) {
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) {
      double s = 0.0;
      for (int l = 0; l < k; ++l) s += A[i + l * m] * B[l + j * k];
      C[i + j * m] = s;
    }
}

/*
 * Arrays in Julia layout: z0 nz×H, th0 nθ×H, r0 nz×H, rz0 nz×nz×H, rth0 nz×nθ×H.
 */
void* oracle_create(int nq, int nu, int nw, int nc, int nb, int mode, int H, const double* z0,
                    const double* th0, const double* r0, const double* rz0, const double* rth0) {
  oracle_t* o = (oracle_t*)calloc(1, sizeof(oracle_t));
  set_dims(&o->d, nq, nu, nw, nc, nb, mode);
  const dims_t* d = &o->d;
  if (d->nx > MAXN || d->ny > MAXN || d->nth > MAXT) { free(o); return NULL; }
  o->H = H;
  o->k = (knot_t*)calloc((size_t)H, sizeof(knot_t));
  const int nx = d->nx, ny = d->ny, nz = d->nz, nth = d->nth;
  const int ox = 0, oy1 = nx, oy2 = nx + ny, odyn = 0, orst = nx;
  for (int t = 0; t < H; ++t) {
    knot_t* k = &o->k[t];
    const double* rz = rz0 + (size_t)t * nz * nz;
    const double* rt = rth0 + (size_t)t * nz * nth;
    for (int j = 0; j < nx; ++j) for (int i = 0; i < nx; ++i) k->Dx[i + j * nx] = rz[(odyn + i) + (size_t)(ox + j) * nz];
    for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) k->Dy1[i + j * nx] = rz[(odyn + i) + (size_t)(oy1 + j) * nz];
    for (int j = 0; j < nx; ++j) for (int i = 0; i < ny; ++i) k->Rx[i + j * ny] = rz[(orst + i) + (size_t)(ox + j) * nz];
    for (int j = 0; j < ny; ++j) for (int i = 0; i < ny; ++i) k->Ry1[i + j * ny] = rz[(orst + i) + (size_t)(oy1 + j) * nz];
    for (int i = 0; i < ny; ++i) k->Ry2[i] = rz[(orst + i) + (size_t)(oy2 + i) * nz];
    for (int j = 0; j < nth; ++j) for (int i = 0; i < nx; ++i) k->rthd[i + j * nx] = rt[(odyn + i) + (size_t)j * nz];
    for (int j = 0; j < nth; ++j) for (int i = 0; i < ny; ++i) k->rthr[i + j * ny] = rt[(orst + i) + (size_t)j * nz];
    for (int i = 0; i < nx; ++i) { k->rdyn0[i] = r0[(size_t)t * nz + odyn + i]; k->x0[i] = z0[(size_t)t * nz + ox + i]; }
    for (int i = 0; i < ny; ++i) {
      k->rrst0[i] = r0[(size_t)t * nz + orst + i];
      k->y10[i] = z0[(size_t)t * nz + oy1 + i];
      k->y20[i] = z0[(size_t)t * nz + oy2 + i];
    }
    for (int i = 0; i < nth; ++i) k->th0[i] = th0[(size_t)t * nth + i];
    invert(nx, k->Dx, k->Ai);                       /* schur.jl:39 */
    matmul(ny, nx, nx, k->Rx, k->Ai, k->CAi);        /* schur.jl:40 */
    matmul(ny, nx, ny, k->CAi, k->Dy1, k->CAiB);     /* schur.jl:41 */
  }
  return o;
}

void oracle_set_solver(void* h, int use_lu) {
  if (h) ((oracle_t*)h)->use_lu = use_lu;
}

void oracle_destroy(void* h) {
  oracle_t* o = (oracle_t*)h;
  if (!o) return;
  free(o->k);
  free(o);
}

/* per-thread solver state */
typedef struct {
  double qs[MAXN * MAXN];            /* qs[j*ny + i]: j-th orthonormal column            */
  double rs[MAXN * (MAXN + 1) / 2];  /* packed upper triangle, column-major (qr.jl:18-22) */
  double y1[MAXN], y2[MAXN];         /* RZLin.y1, y2 at the last rzlin!                   */
  double lu[MAXN * MAXN];            /* accurate variant: LU factors, row-major            */
  int perm[MAXN];
  double rdyn[MAXN], rrst[MAXN], rbil[MAXN];
} work_t;

#define TRIU(k, j) (((j) - 1) * (j) / 2 + (k) - 1) /* 1-based (k, j) */

/* factorize!(gs_solver, A): modified Gram-Schmidt, qr.jl:113-137 */
static void mgs_factorize(int n, const double* A, work_t* w) {
  int off = 0;
  for (int j = 0; j < n; ++j) {
    double* q = &w->qs[j * n];
    for (int i = 0; i < n; ++i) q[i] = A[i + j * n];
    for (int k = 0; k < j; ++k) {
      const double* qk = &w->qs[k * n];
      double r = 0.0;
      for (int i = 0; i < n; ++i) r += q[i] * qk[i];
      w->rs[off++] = r;
      for (int i = 0; i < n; ++i) q[i] -= qk[i] * r;
    }
    double nr = 0.0;
    for (int i = 0; i < n; ++i) nr += q[i] * q[i];
    nr = sqrt(nr);
    w->rs[off++] = nr;
    for (int i = 0; i < n; ++i) q[i] /= nr;
  }
}

/* qr_solve!: x = R⁻¹ Qᵀ b, qr.jl:142-158 */
static void mgs_solve(int n, const work_t* w, const double* b, double* x) {
  for (int j = 0; j < n; ++j) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += w->qs[j * n + i] * b[i];
    x[j] = s;
  }
  for (int j = n; j >= 1; --j) {
    for (int k = j + 1; k <= n; ++k) x[j - 1] -= w->rs[TRIU(j, k)] * x[k - 1];
    x[j - 1] /= w->rs[TRIU(j, j)];
  }
}

/* NOT in the reference: LU with partial pivoting, the accurate variant used by the tests to separate
 * implementation error from the reference's own MGS round-off (cond² · eps). */
static void lu_factorize(int n, const double* A, work_t* w) {
  for (int i = 0; i < n; ++i) { w->perm[i] = i; for (int j = 0; j < n; ++j) w->lu[i * n + j] = A[i + j * n]; }
  for (int k = 0; k < n; ++k) {
    int p = k;
    for (int i = k + 1; i < n; ++i) if (fabs(w->lu[i * n + k]) > fabs(w->lu[p * n + k])) p = i;
    if (p != k) {
      for (int j = 0; j < n; ++j) { double t = w->lu[k * n + j]; w->lu[k * n + j] = w->lu[p * n + j]; w->lu[p * n + j] = t; }
      int t = w->perm[k]; w->perm[k] = w->perm[p]; w->perm[p] = t;
    }
    const double inv = 1.0 / w->lu[k * n + k];
    for (int i = k + 1; i < n; ++i) {
      const double m = w->lu[i * n + k] * inv;
      w->lu[i * n + k] = m;
      for (int j = k + 1; j < n; ++j) w->lu[i * n + j] -= m * w->lu[k * n + j];
    }
  }
}
static void lu_solve_(int n, const work_t* w, const double* b, double* x) {
  for (int i = 0; i < n; ++i) {
    double s = b[w->perm[i]];
    for (int j = 0; j < i; ++j) s -= w->lu[i * n + j] * x[j];
    x[i] = s;
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = x[i];
    for (int j = i + 1; j < n; ++j) s -= w->lu[i * n + j] * x[j];
    x[i] = s / w->lu[i * n + i];
  }
}

/* rlin!, linearized_solver.jl:364-373 */
static void rlin(const dims_t* d, const knot_t* k, const double* alt, const double* z, const double* th,
                 double kappa, work_t* w) {
  const int nx = d->nx, ny = d->ny, nth = d->nth;
  const double *x = z, *y1 = z + nx, *y2 = z + nx + ny;
  double dx[MAXN], dy1[MAXN], dth[MAXT];
  for (int i = 0; i < nx; ++i) dx[i] = x[i] - k->x0[i];
  for (int i = 0; i < ny; ++i) dy1[i] = y1[i] - k->y10[i];
  for (int i = 0; i < nth; ++i) dth[i] = th[i] - k->th0[i];
  for (int i = 0; i < nx; ++i) {
    double s = k->rdyn0[i], a = 0.0, b = 0.0, c = 0.0;
    for (int j = 0; j < nx; ++j) a += k->Dx[i + j * nx] * dx[j];
    for (int j = 0; j < ny; ++j) b += k->Dy1[i + j * nx] * dy1[j];
    for (int j = 0; j < nth; ++j) c += k->rthd[i + j * nx] * dth[j];
    w->rdyn[i] = s + a + b + c;
  }
  for (int i = 0; i < ny; ++i) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int j = 0; j < nx; ++j) a += k->Rx[i + j * ny] * dx[j];
    for (int j = 0; j < ny; ++j) b += k->Ry1[i + j * ny] * dy1[j];
    for (int j = 0; j < nth; ++j) c += k->rthr[i + j * ny] * dth[j];
    w->rrst[i] = k->rrst0[i] + a + b + k->Ry2[i] * (y2[i] - k->y20[i]) + c + ((alt && i < d->nc) ? alt[i] : 0.0);
    w->rbil[i] = y1[i] * y2[i] - kappa;
  }
}

/* rzlin! + schur_factorize!, linearized_solver.jl:378-399, schur.jl:80-88 */
static void rzlin(const dims_t* d, const knot_t* k, const double* z, double reg, work_t* w, int use_lu) {
  const int nx = d->nx, ny = d->ny;
  double S[MAXN * MAXN];
  for (int i = 0; i < ny; ++i) { w->y1[i] = z[nx + i]; w->y2[i] = z[nx + ny + i]; }
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < ny; ++i) {
      double D = k->Ry1[i + j * ny];
      if (i == j) D -= k->Ry2[i] * fmax(w->y2[i], reg) / fmax(w->y1[i], reg);
      S[i + j * ny] = D - k->CAiB[i + j * ny];
    }
  if (use_lu) lu_factorize(ny, S, w); else mgs_factorize(ny, S, w);
}

/* schur_solve!, schur.jl:93-110: x = Ai (u + B t), y = −t, t = S⁻¹ (CAi u − v) */
static void schur_solve(const dims_t* d, const knot_t* k, const work_t* w, const double* u, const double* v,
                        double* x, double* y, int use_lu) {
  const int nx = d->nx, ny = d->ny;
  double rhs[MAXN], t[MAXN], tmp[MAXN];
  for (int i = 0; i < ny; ++i) {
    double s = 0.0;
    for (int j = 0; j < nx; ++j) s += k->CAi[i + j * ny] * u[j];
    rhs[i] = s - v[i];
  }
  if (use_lu) lu_solve_(ny, w, rhs, t); else mgs_solve(ny, w, rhs, t);
  for (int i = 0; i < nx; ++i) {
    double s = 0.0;
    for (int j = 0; j < ny; ++j) s += k->Dy1[i + j * nx] * t[j];
    tmp[i] = u[i] + s;
  }
  for (int i = 0; i < nx; ++i) {
    double s = 0.0;
    for (int j = 0; j < nx; ++j) s += k->Ai[i + j * nx] * tmp[j];
    x[i] = s;
  }
  for (int i = 0; i < ny; ++i) y[i] = -t[i];
}

/* linear_solve!(Δ, rz, r; reg), linearized_solver.jl:424-444 */
static void linear_solve(const dims_t* d, const knot_t* k, const work_t* w, double reg, double* delta, int use_lu) {
  const int nx = d->nx, ny = d->ny;
  double v[MAXN];
  for (int i = 0; i < ny; ++i) v[i] = w->rrst[i] - k->Ry2[i] * w->rbil[i] / fmax(reg, w->y1[i]);
  schur_solve(d, k, w, w->rdyn, v, delta, delta + nx, use_lu);
  for (int i = 0; i < ny; ++i)
    delta[nx + ny + i] = (w->rbil[i] - fmax(reg, w->y2[i]) * delta[nx + i]) / fmax(reg, w->y1[i]);
}

static double step_length(int ny, const double* y1, const double* y2, const double* d1, const double* d2,
                          double tau) {
  double a = 1.0;
  for (int i = 0; i < ny; ++i)
    if (d1[i] > 0.0) a = fmin(a, tau * y1[i] / d1[i]);
  for (int i = 0; i < ny; ++i)
    if (d2[i] > 0.0) a = fmin(a, tau * y2[i] / d2[i]);
  return a;
}

static void violations(const dims_t* d, const work_t* w, double* rv, double* kv) {
  double r = 0.0, k = 0.0;
  for (int i = 0; i < d->nx; ++i) r = fmax(r, fabs(w->rdyn[i]));
  for (int i = 0; i < d->ny; ++i) { r = fmax(r, fabs(w->rrst[i])); k = fmax(k, fabs(w->rbil[i])); }
  *rv = r; *kv = k;
}

/* interior_point_solve! (RoboDojo 0.1.3; oracle/ip.py) for one subproblem.
 * dz_out: nd × ncol column-major (the δq0, δq1, δu1 views of implicit_dynamics.jl:82-86). */
static void ip_solve_one(const oracle_t* o, const ip_opts_t* op, int knot, const double* th, const double* q2,
                         const double* alt, double* z, double* dz_out, uint8_t* status, int32_t* iters_out) {
  const dims_t* d = &o->d;
  const knot_t* k = &o->k[knot];
  const int nx = d->nx, ny = d->ny, nz = d->nz;
  work_t w;
  double daff[2 * MAXN + MAXN], dl[2 * MAXN + MAXN], zc[2 * MAXN + MAXN];
  for (int i = 0; i < nz; ++i) z[i] = 1.0;     /* z_initialize! */
  for (int i = 0; i < nx; ++i) z[i] = q2[i];
  rlin(d, k, alt, z, th, 0.0, &w);
  double r_vio, k_vio, reg = 0.0;
  violations(d, &w, &r_vio, &k_vio);
  int iters = 0;
  for (int it = 0; it < op->max_iter; ++it) {
    if (r_vio < op->r_tol && k_vio < op->kappa_tol) break;
    ++iters;
    reg = (k_vio < op->kappa_reg) ? k_vio * op->gamma_reg : 0.0;
    rzlin(d, k, z, reg, &w, o->use_lu);
    linear_solve(d, k, &w, reg, daff, o->use_lu);
    const double *y1 = z + nx, *y2 = z + nx + ny;
    const double a_aff = step_length(ny, y1, y2, daff + nx, daff + nx + ny, 1.0);
    double mu = 0.0, mu_aff = 0.0;
    for (int i = 0; i < ny; ++i) mu += y1[i] * y2[i];
    mu /= ny;
    for (int i = 0; i < ny; ++i) mu_aff += (y1[i] - a_aff * daff[nx + i]) * (y2[i] - a_aff * daff[nx + ny + i]);
    mu_aff /= ny;
    double sg = fmin(fmax(mu_aff / mu, 0.0), 1.0);
    sg = sg * sg * sg;
    rlin(d, k, alt, z, th, fmax(sg * mu, op->kappa_tol / op->undercut), &w);
    for (int i = 0; i < ny; ++i) w.rbil[i] += daff[nx + i] * daff[nx + ny + i]; /* general_correction_term! */
    linear_solve(d, k, &w, reg, dl, o->use_lu);
    const double vmax = fmax(r_vio, k_vio);
    const double tau = fmax(1.0 - op->eps_min, 1.0 - vmax * vmax);
    double alpha = step_length(ny, y1, y2, dl + nx, dl + nx + ny, tau);
    double rc = 0.0, kc = 0.0;
    for (int ls = 0; ls <= op->max_ls; ++ls) {
      for (int i = 0; i < nz; ++i) zc[i] = z[i] - alpha * dl[i];
      rlin(d, k, alt, zc, th, 0.0, &w);
      violations(d, &w, &rc, &kc);
      if (rc <= r_vio || kc <= k_vio || ls == op->max_ls) break;
      alpha *= op->ls_scale;
    }
    memcpy(z, zc, sizeof(double) * nz);
    r_vio = rc; k_vio = kc;
  }
  *status = (r_vio < op->r_tol && k_vio < op->kappa_tol) ? 1 : 0;
  *iters_out = iters;
  if (op->diff_sol && dz_out) {
    /* differentiate_solution!: all nθ columns are solved (linearized_solver.jl:467-477), only the
       consumed rows / columns are stored */
    const double reg_d = fmax(reg, op->kappa_tol * op->gamma_reg);
    rzlin(d, k, z, reg_d, &w, o->use_lu);
    double x[MAXN], y[MAXN];
    for (int c = 0; c < d->nth; ++c) {
      schur_solve(d, k, &w, &k->rthd[c * nx], &k->rthr[c * ny], x, y, o->use_lu);
      if (c < d->ncol) {
        for (int i = 0; i < nx; ++i) dz_out[c * d->nd + i] = -x[i];
        for (int i = 0; i < d->nd - nx; ++i) dz_out[c * d->nd + nx + i] = -y[i];
      }
    }
  }
}

/* Batched solve; arrays as in cimpc_ip_solve_batch_host.  threads <= 0: all OpenMP threads. */
int oracle_ip_solve_batch(void* h, int64_t n, const int32_t* knot, const double* theta, const double* q2,
                          const double* alt, const ip_opts_t* op, double* z_out, double* dz_out,
                          uint8_t* status, int32_t* iters, int threads) {
  const oracle_t* o = (const oracle_t*)h;
  if (!o) return 1;
  const dims_t* d = &o->d;
#ifdef _OPENMP
  omp_set_num_threads(threads > 0 ? threads : omp_get_num_procs());
#endif
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < n; ++i) {
    int kn = knot[i];
    if (kn < 0 || kn >= o->H) kn = 0;
    ip_solve_one(o, op, kn, theta + i * d->nth, q2 + i * d->nx, alt ? alt + i * d->nc : NULL,
                 z_out + i * d->nz, dz_out ? dz_out + i * (int64_t)d->nd * d->ncol : NULL, status + i, iters + i);
  }
  return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}
