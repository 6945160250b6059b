// C ABI (include/cimpc_b200.h): context, linearization upload + set-up kernel, batched solves.
#include <dlfcn.h>

#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "registry.cuh"

using namespace cimpc;

// ------------------------------------------------------------------------------------------
// Set-up kernel: one CTA per reference knot.  Slices r0 / rz0 / rθ0 into the RLin / RZLin / RθLin
// blocks (src/controller/linearized_solver.jl:92-120, 245-251, 346-348) and forms the constant
// products of the Schur constructor (src/solver/schur.jl:33-49): Ai = Dx⁻¹ (Gauss-Jordan with
// partial pivoting), CAi = Rx Ai, AiB = Ai Dy1, S0 = Ry1 − CAi Dy1, plus the θ-offset vectors and the
// sensitivity right-hand sides W = CAi Rθdyn − Rθrst, AR = Ai Rθdyn used by ip_solve_kernel.
// ------------------------------------------------------------------------------------------
struct PrepParams {
  LinLayout lay;
  const double* z0;    // nz × H
  const double* th0;   // nθ × H
  const double* r0;    // nz × H
  const double* rz0;   // nz × nz × H
  const double* rth0;  // nz × nθ × H
  double* lin;         // H × stride
  int* flag;           // set to 1 when a knot lacks the block structure the reduced Schur complement relies on
};

__global__ void __launch_bounds__(128) prep_kernel(const PrepParams p) {
  const LinLayout& a = p.lay;
  const int nx = a.nx, ny = a.ny, nz = a.nz, nth = a.nth, ncol = a.ncol, G = a.group;
  const int nr = a.nr, nrp = a.nrp, npsi = ny - nr;
  const int t = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const double* z0 = p.z0 + (size_t)t * nz;
  const double* th0 = p.th0 + (size_t)t * nth;
  const double* r0 = p.r0 + (size_t)t * nz;
  const double* rz = p.rz0 + (size_t)t * nz * nz;
  const double* rth = p.rth0 + (size_t)t * nz * nth;
  double* L = p.lin + (size_t)t * a.stride;
  // z = [x(nx); y1(ny); y2(ny)],  r = [dyn(nx); rst(ny); bil(ny)]   (index.jl:289-327)
  const int ox = 0, oy1 = nx, oy2 = nx + ny, odyn = 0, orst = nx;

  // dense column-major work copies in shared memory
  extern __shared__ double sm[];
  double* Dx = sm;                 // nx×nx
  double* Dy1 = Dx + nx * nx;      // nx×ny
  double* Rx = Dy1 + nx * ny;      // ny×nx
  double* Ry1 = Rx + ny * nx;      // ny×ny
  double* Ry2 = Ry1 + ny * ny;     // ny
  double* Rtd = Ry2 + ny;          // nx×nth
  double* Rtr = Rtd + nx * nth;    // ny×nth
  double* Ai = Rtr + ny * nth;     // nx×nx
  double* CAi = Ai + nx * nx;      // ny×nx
  double* AiB = CAi + ny * nx;     // nx×ny
  double* S0 = AiB + nx * ny;      // ny×ny
  double* M = S0 + ny * ny;        // nx×2nx (Gauss-Jordan work, row-major)
  __shared__ int s_piv;

  for (int e = tid; e < a.stride; e += nt) L[e] = 0.0;
  for (int e = tid; e < nx * nx; e += nt) Dx[e] = rz[(odyn + e % nx) + (size_t)(ox + e / nx) * nz];
  for (int e = tid; e < nx * ny; e += nt) Dy1[e] = rz[(odyn + e % nx) + (size_t)(oy1 + e / nx) * nz];
  for (int e = tid; e < ny * nx; e += nt) Rx[e] = rz[(orst + e % ny) + (size_t)(ox + e / ny) * nz];
  for (int e = tid; e < ny * ny; e += nt) Ry1[e] = rz[(orst + e % ny) + (size_t)(oy1 + e / ny) * nz];
  for (int e = tid; e < ny; e += nt) Ry2[e] = rz[(orst + e) + (size_t)(oy2 + e) * nz];
  for (int e = tid; e < nx * nth; e += nt) Rtd[e] = rth[(odyn + e % nx) + (size_t)(e / nx) * nz];
  for (int e = tid; e < ny * nth; e += nt) Rtr[e] = rth[(orst + e % ny) + (size_t)(e / ny) * nz];
  __syncthreads();

  // ---- Ai = Dx⁻¹ by Gauss-Jordan with partial pivoting on [Dx | I] ----
  const int w2 = 2 * nx;
  for (int e = tid; e < nx * w2; e += nt) {
    const int i = e / w2, j = e % w2;
    M[e] = (j < nx) ? Dx[i + j * nx] : ((j - nx == i) ? 1.0 : 0.0);
  }
  __syncthreads();
  for (int k = 0; k < nx; ++k) {
    if (tid == 0) {
      int best = k;
      double bv = fabs(M[k * w2 + k]);
      for (int i = k + 1; i < nx; ++i) {
        const double v = fabs(M[i * w2 + k]);
        if (v > bv) { bv = v; best = i; }
      }
      s_piv = best;
    }
    __syncthreads();
    const int pr = s_piv;
    if (pr != k) {
      for (int j = tid; j < w2; j += nt) {
        const double tmp = M[k * w2 + j];
        M[k * w2 + j] = M[pr * w2 + j];
        M[pr * w2 + j] = tmp;
      }
    }
    __syncthreads();
    const double pinv = 1.0 / M[k * w2 + k];
    __syncthreads();
    for (int j = tid; j < w2; j += nt) M[k * w2 + j] *= pinv;
    __syncthreads();
    for (int e = tid; e < nx * w2; e += nt) {
      const int i = e / w2, j = e % w2;
      if (i != k && j != k) M[e] = fma(-M[i * w2 + k], M[k * w2 + j], M[e]);
    }
    __syncthreads();
    for (int i = tid; i < nx; i += nt)
      if (i != k) M[i * w2 + k] = 0.0;
    __syncthreads();
  }
  for (int e = tid; e < nx * nx; e += nt) Ai[e] = M[(e % nx) * w2 + nx + e / nx];
  __syncthreads();

  // ---- constant products (schur.jl:39-41) ----
  for (int e = tid; e < ny * nx; e += nt) {  // CAi = Rx Ai
    const int i = e % ny, j = e / ny;
    double s = 0.0;
    for (int k = 0; k < nx; ++k) s = fma(Rx[i + k * ny], Ai[k + j * nx], s);
    CAi[e] = s;
  }
  for (int e = tid; e < nx * ny; e += nt) {  // AiB = Ai Dy1
    const int i = e % nx, j = e / nx;
    double s = 0.0;
    for (int k = 0; k < nx; ++k) s = fma(Ai[i + k * nx], Dy1[k + j * nx], s);
    AiB[e] = s;
  }
  __syncthreads();
  for (int e = tid; e < ny * ny; e += nt) {  // S0 = Ry1 − CAi Dy1
    const int i = e % ny, j = e / ny;
    double s = 0.0;
    for (int k = 0; k < nx; ++k) s = fma(CAi[i + k * ny], Dy1[k + j * nx], s);
    S0[e] = Ry1[e] - s;
  }
  __syncthreads();

  // ---- packed, lane-major output layout (dims.cuh) ----
  for (int e = tid; e < (nx + ny) * G; e += nt) {  // RES: {[Dx Dy1][l][j], [Rx Ry1][l][j]}
    const int l = e % G, j = e / G;
    double d = 0.0, r = 0.0;
    if (j < nx) {
      if (l < nx) d = Dx[l + j * nx];
      if (l < ny) r = Rx[l + j * ny];
    } else {
      if (l < nx) d = Dy1[l + (j - nx) * nx];
      if (l < ny) r = Ry1[l + (j - nx) * ny];
    }
    L[a.o_res + 2 * e] = d;
    L[a.o_res + 2 * e + 1] = r;
  }
  for (int e = tid; e < nx * G; e += nt) {  // CA2: {CAi[l][j], Ai[l][j]}; AIBR: AiB[i][l]
    const int l = e % G, j = e / G;
    L[a.o_ca2 + 2 * e] = (l < ny) ? CAi[l + j * ny] : 0.0;
    L[a.o_ca2 + 2 * e + 1] = (l < nx) ? Ai[l + j * nx] : 0.0;
    L[a.o_aibr + e] = (l < nr) ? AiB[j + l * nx] : 0.0;
  }
  // ---- reduced Schur complement (dims.cuh): S0 = [A0 B; C D0] over y1 = [γ1; b1 | ψ1], rows [imp; mdp | fri] ----
  // structure the elimination of ψ1 relies on (all exact zeros of the residual, simulation.jl:133-158)
  for (int e = tid; e < npsi * npsi; e += nt)
    if (S0[(nr + e % npsi) + (nr + e / npsi) * ny] != 0.0) *p.flag = 1;        // D0 = Ry1[fri, ψ] = 0
  for (int e = tid; e < nx * npsi; e += nt)
    if (AiB[e % nx + (nr + e / nx) * nx] != 0.0) *p.flag = 1;                  // Dy1[:, ψ] = 0
  for (int l = tid; l < nr; l += nt) {                                           // B: at most one non-zero per row
    int cnt = 0;
    for (int c = 0; c < npsi; ++c) cnt += (S0[l + (nr + c) * ny] != 0.0);
    if (cnt > 1) *p.flag = 1;
  }
  // c(l), b0(c): contact of every reduced row and the first row of every contact (shared scratch: reuse M)
  int* cidv = reinterpret_cast<int*>(M);  // [nr] c(l), then [npsi] b0(c)
  int* b0v = cidv + nr;
  __syncthreads();
  if (tid == 0) {
    for (int c = 0; c < npsi; ++c) b0v[c] = -1;
    for (int l = 0; l < nr; ++l) {
      cidv[l] = -1;
      for (int c = 0; c < npsi; ++c)
        if (S0[l + (nr + c) * ny] != 0.0) {
          cidv[l] = c;
          if (b0v[c] < 0) b0v[c] = l;
        }
    }
    for (int c = 0; c < npsi; ++c)
      if (b0v[c] < 0) *p.flag = 1;  // ψ_c must couple to at least one row
  }
  __syncthreads();
  for (int e = tid; e < nrp * G; e += nt) {  // AIBC, A0 (+ identity padding), A0ᵀ, BC, BCᵀ, AB, ABᵀ
    const int l = e % G, j = e / G;
    L[a.o_aibc + e] = (l < nx && j < nr) ? AiB[l + j * nx] : 0.0;
    const bool in = l < nr && j < nr;
    const double pad = (l == j && l >= nr && l < nrp) ? 1.0 : 0.0;
    L[a.o_s0 + e] = in ? S0[l + j * ny] : pad;
    L[a.o_s0t + e] = in ? S0[j + l * ny] : pad;
    double bc = 0.0, bct = 0.0, ab = 0.0, abt = 0.0;
    if (in) {
      const int cl = cidv[l], cj = cidv[j];
      if (cl >= 0 && b0v[cl] >= 0) {
        bc = S0[l + (nr + cl) * ny] * S0[(nr + cl) + j * ny];  // B[l,c]·C[c,j]
        ab = S0[b0v[cl] + j * ny];                              // A0[b0(l), j]
      }
      if (cj >= 0 && b0v[cj] >= 0) {
        bct = S0[j + (nr + cj) * ny] * S0[(nr + cj) + l * ny];
        abt = S0[b0v[cj] + l * ny];
      }
    }
    if (l >= nr && l < ny && j < nr && b0v[l - nr] >= 0) ab = S0[b0v[l - nr] + j * ny];  // ψ lanes: row b0 of their contact
    L[a.o_bc + e] = bc;
    L[a.o_bct + e] = bct;
    L[a.o_ab + e] = ab;
    L[a.o_abt + e] = abt;
  }
  for (int l = tid; l < G; l += nt) {  // CRW: the (at most NFR) non-zeros of row l − nr of C = S0[fri, γb]
    int k = 0;
    if (l >= nr && l < ny)
      for (int j = 0; j < nr; ++j) {
        const double v = S0[l + j * ny];
        if (v != 0.0) {
          if (k < a.nfr) {
            L[a.o_crw + (2 * k) * G + l] = v;
            L[a.o_crw + (2 * k + 1) * G + l] = (double)j;
          } else {
            *p.flag = 1;
          }
          ++k;
        }
      }
  }
  for (int e = tid; e < G; e += nt) {  // per-lane scalars
    L[a.o_ry2 + e] = (e < ny) ? Ry2[e] : 0.0;
    double bv = 0.0, cid = -1.0, b0 = -1.0, rho = 0.0, pb0 = -1.0, ibv0 = 0.0;
    if (e < nr && cidv[e] >= 0 && b0v[cidv[e]] >= 0) {
      const int c = cidv[e];
      bv = S0[e + (nr + c) * ny];
      cid = (double)c;
      b0 = (double)b0v[c];
      rho = bv / S0[b0v[c] + (nr + c) * ny];
      ibv0 = 1.0 / S0[b0v[c] + (nr + c) * ny];
    }
    if (e >= nr && e < ny && b0v[e - nr] >= 0) {
      pb0 = (double)b0v[e - nr];
      ibv0 = 1.0 / S0[b0v[e - nr] + e * ny];
    }
    L[a.o_bv + e] = bv;
    L[a.o_cid + e] = cid;
    L[a.o_b0 + e] = b0;
    L[a.o_rho + e] = rho;
    L[a.o_pb0 + e] = pb0;
    L[a.o_ibv0 + e] = ibv0;
  }
  for (int j = tid; j < nrp; j += nt) {  // per-row scalars at uniform addresses
    const bool has = j < nr && cidv[j] >= 0 && b0v[cidv[j]] >= 0;
    L[a.o_cidu + j] = has ? (double)cidv[j] : -1.0;
    L[a.o_rhou + j] = has ? S0[j + (nr + cidv[j]) * ny] / S0[b0v[cidv[j]] + (nr + cidv[j]) * ny] : 0.0;
    L[a.o_b0u + j] = has ? (double)b0v[cidv[j]] : (double)j;
  }
  for (int e = tid; e < ny * ncol; e += nt) {  // W = CAi Rθdyn − Rθrst (first ncol columns, rows γ1 b1), W[k][c] at c*NRP + k
    const int i = e % ny, j = e / ny;
    double s = 0.0;
    for (int k = 0; k < nx; ++k) s = fma(CAi[i + k * ny], Rtd[k + j * nx], s);
    s -= Rtr[i + j * ny];
    if (i < nr) L[a.o_w + j * nrp + i] = s;
    else if (s != 0.0) *p.flag = 1;  // the fri rows do not depend on (q0, q1, u1)
  }
  for (int e = tid; e < ncol * G; e += nt) {  // AR = Ai Rθdyn[:, 1:ncol], AR[l][c] at c*G + l
    const int l = e % G, j = e / G;
    double s = 0.0;
    if (l < nx)
      for (int k = 0; k < nx; ++k) s = fma(Ai[l + k * nx], Rtd[k + j * nx], s);
    L[a.o_ar + e] = s;
  }
  for (int e = tid; e < G; e += nt) {  // C0: {cdyn0, crst0}
    double cd = 0.0, cr = 0.0;
    if (e < nx) {  // cdyn = rdyn0 − Dx x0 − Dy1 y10 − Rθdyn θ0
      cd = r0[odyn + e];
      for (int k = 0; k < nx; ++k) cd = fma(-Dx[e + k * nx], z0[ox + k], cd);
      for (int k = 0; k < ny; ++k) cd = fma(-Dy1[e + k * nx], z0[oy1 + k], cd);
      for (int k = 0; k < nth; ++k) cd = fma(-Rtd[e + k * nx], th0[k], cd);
    }
    if (e < ny) {  // crst = rrst0 − Rx x0 − Ry1 y10 − Ry2∘y20 − Rθrst θ0
      cr = r0[orst + e];
      for (int k = 0; k < nx; ++k) cr = fma(-Rx[e + k * ny], z0[ox + k], cr);
      for (int k = 0; k < ny; ++k) cr = fma(-Ry1[e + k * ny], z0[oy1 + k], cr);
      cr = fma(-Ry2[e], z0[oy2 + e], cr);
      for (int k = 0; k < nth; ++k) cr = fma(-Rtr[e + k * ny], th0[k], cr);
    }
    L[a.o_c0 + 2 * e] = cd;
    L[a.o_c0 + 2 * e + 1] = cr;
  }
  for (int e = tid; e < nth * G; e += nt) {  // RTH: {Rθdyn[l][j], Rθrst[l][j]}
    const int l = e % G, j = e / G;
    L[a.o_rth + 2 * e] = (l < nx) ? Rtd[l + j * nx] : 0.0;
    L[a.o_rth + 2 * e + 1] = (l < ny) ? Rtr[l + j * ny] : 0.0;
  }
}

// ------------------------------------------------------------------------------------------
struct cimpc_ctx {
  int device = 0;
  int sm_count = 0;
  const ModelEntry* entry = nullptr;
  double* lin = nullptr;
  int32_t h_ref = 0;
  int64_t launches = 0;
  std::string cuda_err;
  // staging for the host entry point
  void* pin = nullptr;   size_t pin_bytes = 0;
  void* dev = nullptr;   size_t dev_bytes = 0;
  static constexpr int NS = 3;            // pipeline depth of the host entry point
  static constexpr int64_t CHUNK = 32768;  // subproblems per pipelined chunk
  cudaStream_t streams[NS] = {nullptr, nullptr, nullptr};
  // device Newton state (cimpc_newton_create)
  struct NewtonState {
    NewtonParams p{};
    void* arena = nullptr;
    cimpc_ip_opts ip{};
    cimpc_newton_opts no{};
    double *z = nullptr, *dz = nullptr, *ref_q = nullptr, *ref_u = nullptr, *w = nullptr, *obj_q = nullptr,
           *obj_u = nullptr;
    int32_t* window = nullptr;
    uint8_t* status = nullptr;
    int32_t* iters = nullptr;
    double* lscratch = nullptr;  // factor block columns of every rollout
    double *ref_y = nullptr, *obj_y = nullptr, *obj_v = nullptr;
    int variant = -1;  // −1: specialised kernel (:configuration, TrackingObjective); 0 / 1: general kernel without / with velocity cost
    int* h_active = nullptr;  // pinned
    int32_t last_sweeps = 0;
    bool ready = false;
    // one MPC step of every rollout = ONE graph launch: [zero control words] → newton_reset → WHILE { ip_solve →
    // newton_step → newton_advance } → newton_finish.  Per-call pointers travel through `call` (device), filled from a
    // pinned ring so that consecutive asynchronous calls never wait for each other.
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    NewtonCall* call = nullptr;          // device (arena)
    static constexpr int RING = 16;
    NewtonCall* ring = nullptr;          // pinned, RING slots
    cudaEvent_t ring_ev[RING] = {};
    int ring_pos = 0;
    int32_t* ctrl = nullptr;             // device: act_count[2], par, sweep_ctr
    int max_rounds = 0;
    // host buffers of cimpc_newton_solve_batch_host
    double* hq = nullptr;  size_t hq_doubles = 0;   // pinned: q0 | q1 | u
    double* dq = nullptr;                           // device: q0 | q1 | u
  } nw;
  double* sim_scratch = nullptr;
  size_t sim_scratch_doubles = 0;
  double* dense = nullptr;  // dense linearization arrays of the current reference (see alloc_dense)
  int32_t dense_h = 0;
  bool lin_bad_structure = false;
  void* nccl_comm = nullptr;  // communicator created by cimpc_comm_init (destroyed with the context)
  int nccl_world = 1, nccl_rank = 0;
  int* prep_flag = nullptr;  // device int written by prep_kernel (structure check of the uploaded linearization)
};

static int cuda_fail(cimpc_ctx* c, cudaError_t e, const char* where) {
  if (c) c->cuda_err = std::string(where) + ": " + cudaGetErrorString(e);
  return CIMPC_ERR_CUDA;
}
#define CK(call)                                            \
  do {                                                      \
    cudaError_t e_ = (call);                                \
    if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call); \
  } while (0)

static void newton_release(cimpc_ctx* ctx) {
  auto& nw = ctx->nw;
  if (nw.exec) cudaGraphExecDestroy(nw.exec);
  if (nw.graph) cudaGraphDestroy(nw.graph);
  nw.exec = nullptr; nw.graph = nullptr;
  if (nw.arena) cudaFree(nw.arena);
  nw.arena = nullptr;
  if (nw.ring) cudaFreeHost(nw.ring);
  nw.ring = nullptr;
  for (auto& e : nw.ring_ev)
    if (e) { cudaEventDestroy(e); e = nullptr; }
  if (nw.hq) cudaFreeHost(nw.hq);
  if (nw.dq) cudaFree(nw.dq);
  nw.hq = nullptr; nw.dq = nullptr; nw.hq_doubles = 0;
  nw.ready = false;
}

// First compiled instance with these sizes / mode (the nominal robots come first); with a name, the instance of that
// model — the reference's names (`get_simulation("centroidal_quadruped", …)`, `model_variable_name = "quadruped_payload"`)
// and the short tags of csrc/gen are both accepted.
static const ModelEntry* find_entry(const cimpc_model_desc& d, const char* model_name) {
  std::string want = model_name ? model_name : "";
  if (want == "hopper_2D") want = "hopper2d";
  if (want == "hopper_2D_piecewise") want = "hopper2d_piecewise";
  if (want == "centroidal_quadruped") want = "centroidal";
  if (want == "centroidal_quadruped_payload") want = "centroidal_payload";
#define CIMPC_SEARCH(tag_, nq, nu, nw, nc, nb)                                                   \
  {                                                                                              \
    int cnt = 0;                                                                                 \
    const ModelEntry* e = entries_##tag_(&cnt);                                                  \
    for (int i = 0; i < cnt; ++i)                                                                \
      if (std::memcmp(&e[i].desc, &d, sizeof(cimpc_model_desc)) == 0 && (want.empty() || want == e[i].name))      \
        return &e[i];                                                                            \
  }
  CIMPC_FOR_EACH_MODEL(CIMPC_SEARCH)
#undef CIMPC_SEARCH
  return nullptr;
}

__global__ void newton_finish_kernel(const NewtonParams p, const NewtonCall* __restrict__ call, int nq, int nu, int nyd,
                                     double r_tol, int len) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.R) return;
  double* u_out = call->u_out;
  double* q_out = call->q_out;
  double* y_out = call->y_out;
  int32_t* info = call->info;
  for (int k = 0; k < nu; ++k) u_out[(size_t)r * nu + k] = p.traj_u[(size_t)r * p.H * nu + k];
  if (q_out)
    for (int e = 0; e < (p.H + 2) * nq; ++e) q_out[(size_t)r * (p.H + 2) * nq + e] = p.traj_q[(size_t)r * (p.H + 2) * nq + e];
  if (y_out && nyd > 0)
    for (int e = 0; e < p.H * nyd; ++e) y_out[(size_t)r * p.H * nyd + e] = p.traj_y[(size_t)r * p.H * nyd + e];
  if (info) {
    info[4 * r + 0] = p.newton_it[r];
    info[4 * r + 1] = p.sweeps[r];
    info[4 * r + 2] = (p.r_norm[r] / (double)len < r_tol) ? 1 : 0;
    info[4 * r + 3] = p.phase[r];
  }
}

// Between two sweeps (one thread): retire the list just processed, flip to the other one, count the sweep.
__global__ void newton_advance_kernel(const NewtonParams p) {
  const int cur = *p.par;
  if (p.act_count[cur] == 0 && p.act_count[cur ^ 1] == 0) return;  // solve finished: keep the sweep count
  p.act_count[cur] = 0;
  *p.par = cur ^ 1;
  *p.sweep_ctr += 1;
}


// The kernels of one implicit_dynamics! sweep + Newton update, sizes read on the device.
static int newton_sweep_launch(cimpc_ctx* ctx, cudaStream_t s) {
  auto& nw = ctx->nw;
  NewtonParams& p = nw.p;
  IpParams ip;
  ip.n = (int64_t)p.H * p.R;  // sizes the grid; the kernel reads the live count through `par`
  ip.knot = p.knot; ip.theta = p.theta; ip.q2_init = p.q2; ip.alt = p.alt; ip.alt_by_rollout = 1;
  ip.lin = ctx->lin; ip.h_ref = ctx->h_ref; ip.z_out = nw.z; ip.dz_out = nw.dz; ip.status = nw.status;
  ip.iters = nw.iters; ip.o = nw.ip; ip.R = p.R;
  ip.act2 = p.act_list; ip.cnt2 = p.act_count; ip.par = p.par; ip.stages = p.H;
  cudaError_t e = ctx->entry->launch(ip, ctx->sm_count, s);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "ip_solve_kernel launch");
  e = nw.variant < 0 ? ctx->entry->newton_step(p, nw.lscratch, s)
      : (nw.variant < 2 ? ctx->entry->newton_step_g[nw.variant](p, nw.lscratch, s)
                        : ctx->entry->newton_step_d[nw.variant - 2](p, nw.lscratch, s));
  if (e != cudaSuccess) return cuda_fail(ctx, e, "newton_step_kernel launch");
  newton_advance_kernel<<<1, 1, 0, s>>>(p);
  CK(cudaGetLastError());
  return CIMPC_OK;
}

// One MPC step of every rollout as ONE graph launch, no host round trip between sweeps:
//   [zero control words] → newton_reset → max_rounds × { ip_solve (compacted to the active rollouts, sizes read on the
//   device) → newton_step → newton_advance (flips the lists) } → newton_finish
// The sweep count is bounded by the reference's own iteration caps (newton.jl:202-269: 1 + max_iter·8), so the loop is
// unrolled to that bound; sweeps after the last active rollout has finished find empty lists and exit at once
// (≈ 15 µs for the three launches).  A conditional WHILE node (CUDA 12.4) was built and measured first: correct, but
// 0.58 ms PER ITERATION slower on this driver (62.1 vs 43.6 ms per batch of 16 384 MPC steps), so the bounded unroll
// is what ships.
static int newton_build_graph(cimpc_ctx* ctx) {
  auto& nw = ctx->nw;
  NewtonParams& p = nw.p;
  const cimpc_model_desc& d = ctx->entry->desc;
  const int H = p.H, R = p.R, nyd = ctx->entry->lay.nd - d.nq;
  nw.max_rounds = 1 + p.max_iter * 8 + 1;  // one evaluation + ≤ 8 line-search sweeps per iteration (+ 1 slack)
  cudaStream_t s = ctx->streams[0];
  { int occ = 0; CK(ctx->entry->occupancy(&occ)); }  // opt-in attributes are set outside the capture
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
  CK(cudaMemsetAsync(nw.ctrl, 0, sizeof(int32_t) * 4, s));
  cudaError_t e = ctx->entry->newton_reset(p, nw.call, s);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "newton_reset_kernel capture");
  for (int round = 0; round < nw.max_rounds; ++round) {
    int rc = newton_sweep_launch(ctx, s);
    if (rc != CIMPC_OK) return rc;
  }
  const int len = H * (d.nu + d.nq + nyd + ctx->entry->lay.nd);
  newton_finish_kernel<<<(R + 127) / 128, 128, 0, s>>>(p, nw.call, d.nq, d.nu, nyd, p.r_tol, len);
  CK(cudaGetLastError());
  CK(cudaStreamEndCapture(s, &nw.graph));
  CK(cudaGraphInstantiate(&nw.exec, nw.graph, 0));
  return CIMPC_OK;
}

// ---- NCCL, bound at run time -------------------------------------------------------------------------------------
// The library does not link libnccl: single-GPU users never need it, and a host that already carries an NCCL (CUDA.jl's
// NCCL.jl artifact, torch's bundled copy) must keep using THAT copy — communicator handles are not portable between two
// loaded NCCL builds.  So: take the libnccl.so.2 already mapped into the process if there is one, else dlopen it.
namespace {
struct NcclApi {
  struct Id { char b[128]; };  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed BY VALUE to ncclCommInitRank
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*CommCount)(void*, int*) = nullptr;
  int (*CommUserRank)(void*, int*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    auto sym = [&](const char* n) { return dlsym(h, n); };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.CommCount = (decltype(api.CommCount))sym("ncclCommCount");
    api.CommUserRank = (decltype(api.CommUserRank))sym("ncclCommUserRank");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.CommCount && api.CommUserRank &&
             api.GroupStart && api.GroupEnd && api.Send && api.Recv;
  });
  return api;
}
constexpr int NCCL_FLOAT64 = 8;  // ncclDataType_t ncclFloat64 (nccl.h; stable since NCCL 2.0)
int nccl_fail(cimpc_ctx* c, int rc, const char* where) {
  NcclApi& a = nccl_api();
  if (c) c->cuda_err = std::string(where) + ": NCCL " + (a.GetErrorString ? a.GetErrorString(rc) : "error");
  return CIMPC_ERR_CUDA;
}
}  // namespace

extern "C" {

void cimpc_ip_opts_default(cimpc_ip_opts* o) {
  if (!o) return;
  o->r_tol = 1e-5; o->kappa_tol = 1e-5; o->eps_min = 0.05; o->kappa_reg = 1e-3; o->gamma_reg = 0.1;
  o->undercut = 5.0; o->ls_scale = 0.5; o->max_iter = 100; o->max_ls = 3; o->diff_sol = 0; o->reserved = 0;
}

const char* cimpc_status_string(int s) {
  switch (s) {
    case CIMPC_OK: return "ok";
    case CIMPC_ERR_INVALID_ARGUMENT: return "invalid argument";
    case CIMPC_ERR_UNSUPPORTED_MODEL: return "unsupported model dimensions (no compiled kernel instance)";
    case CIMPC_ERR_NOT_INITIALIZED: return "linearization not uploaded";
    case CIMPC_ERR_CUDA: return "CUDA error (see cimpc_last_cuda_error)";
    case CIMPC_ERR_NO_DEVICE: return "no usable CUDA device";
    default: return "unknown status";
  }
}

const char* cimpc_last_cuda_error(const cimpc_ctx* ctx) { return ctx ? ctx->cuda_err.c_str() : ""; }
int cimpc_version(void) { return CIMPC_B200_VERSION; }
int64_t cimpc_launch_count(const cimpc_ctx* ctx) { return ctx ? ctx->launches : 0; }

int cimpc_create(cimpc_ctx** out, int device, const cimpc_model_desc* desc) {
  return cimpc_create_named(out, device, desc, nullptr);
}

int cimpc_create_named(cimpc_ctx** out, int device, const cimpc_model_desc* desc, const char* model_name) {
  if (!out || !desc) return CIMPC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  const ModelEntry* e = find_entry(*desc, model_name);
  if (!e) return CIMPC_ERR_UNSUPPORTED_MODEL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    (void)cudaGetLastError();
    return CIMPC_ERR_NO_DEVICE;
  }
  cimpc_ctx* ctx = new (std::nothrow) cimpc_ctx();
  if (!ctx) return CIMPC_ERR_INVALID_ARGUMENT;
  ctx->device = device;
  ctx->entry = e;
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete ctx;
    return CIMPC_ERR_NO_DEVICE;
  }
  if (prop.major != 10) {  // sm_100a cubin only
    delete ctx;
    return CIMPC_ERR_NO_DEVICE;
  }
  ctx->sm_count = prop.multiProcessorCount;
  for (int i = 0; i < cimpc_ctx::NS; ++i)
    if (cudaStreamCreateWithFlags(&ctx->streams[i], cudaStreamNonBlocking) != cudaSuccess) {
      delete ctx;
      return CIMPC_ERR_CUDA;
    }
  *out = ctx;
  return CIMPC_OK;
}

int cimpc_destroy(cimpc_ctx* ctx) {
  if (!ctx) return CIMPC_OK;
  cudaSetDevice(ctx->device);
  if (ctx->lin) cudaFree(ctx->lin);
  newton_release(ctx);
  if (ctx->nccl_comm && nccl_api().ok) nccl_api().CommDestroy(ctx->nccl_comm);
  if (ctx->sim_scratch) cudaFree(ctx->sim_scratch);
  if (ctx->dense) cudaFree(ctx->dense);
  if (ctx->prep_flag) cudaFree(ctx->prep_flag);
  if (ctx->nw.h_active) cudaFreeHost(ctx->nw.h_active);
  if (ctx->dev) cudaFree(ctx->dev);
  if (ctx->pin) cudaFreeHost(ctx->pin);
  for (int i = 0; i < cimpc_ctx::NS; ++i)
    if (ctx->streams[i]) cudaStreamDestroy(ctx->streams[i]);
  delete ctx;
  return CIMPC_OK;
}

int cimpc_get_dims(const cimpc_ctx* ctx, cimpc_dims* d) {
  if (!ctx || !d) return CIMPC_ERR_INVALID_ARGUMENT;
  const LinLayout& l = ctx->entry->lay;
  d->nx = l.nx; d->ny = l.ny; d->nz = l.nz; d->ntheta = l.nth; d->nd = l.nd; d->ncol = l.ncol; d->group = l.group;
  return CIMPC_OK;
}

// Dense linearization arrays of the current reference on the device: [z0 | r0 | θ0 | rz0 | rθ0].
static cudaError_t alloc_dense(cimpc_ctx* ctx, int32_t H) {
  const LinLayout& l = ctx->entry->lay;
  const size_t nz = l.nz, nth = l.nth;
  const size_t tot = (2 * nz + nth + nz * nz + nz * nth) * (size_t)H;
  if (ctx->dense && ctx->dense_h == H) return cudaSuccess;
  if (ctx->dense) cudaFree(ctx->dense);
  ctx->dense = nullptr;
  ctx->dense_h = 0;
  cudaError_t e = cudaMalloc(&ctx->dense, tot * sizeof(double));
  if (e == cudaSuccess) ctx->dense_h = H;
  return e;
}
struct DensePtrs { double *z0, *r0, *t0, *rz, *rt; };
static DensePtrs dense_ptrs(const cimpc_ctx* ctx) {
  const LinLayout& l = ctx->entry->lay;
  const size_t nz = l.nz, nth = l.nth, H = ctx->dense_h;
  DensePtrs d;
  d.z0 = ctx->dense; d.r0 = d.z0 + nz * H; d.t0 = d.r0 + nz * H; d.rz = d.t0 + nth * H; d.rt = d.rz + nz * nz * H;
  return d;
}

// RLin / RZLin / RθLin / Schur constants of every knot from the dense arrays (prep_kernel).
static cudaError_t run_prep(cimpc_ctx* ctx, int32_t H, cudaStream_t s) {
  const LinLayout& l = ctx->entry->lay;
  cudaError_t e = cudaSuccess;
  if (ctx->h_ref != H || !ctx->lin) {
    if (ctx->lin) cudaFree(ctx->lin);
    ctx->lin = nullptr;
    ctx->h_ref = 0;
    e = cudaMalloc(&ctx->lin, (size_t)H * l.stride * sizeof(double));
    if (e != cudaSuccess) return e;
  }
  const DensePtrs d = dense_ptrs(ctx);
  if (!ctx->prep_flag) {
    e = cudaMalloc(&ctx->prep_flag, sizeof(int));
    if (e != cudaSuccess) return e;
  }
  e = cudaMemsetAsync(ctx->prep_flag, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  PrepParams pp{l, d.z0, d.t0, d.r0, d.rz, d.rt, ctx->lin, ctx->prep_flag};
  const size_t smem = sizeof(double) * ((size_t)4 * l.nx * l.nx + 4 * l.nx * l.ny + 2 * l.ny * l.ny + l.ny +
                                        (size_t)(l.nx + l.ny) * l.nth + 16);
  if (smem > 48 * 1024) {
    e = cudaFuncSetAttribute(prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  prep_kernel<<<H, 128, smem, s>>>(pp);
  ctx->launches++;
  e = cudaGetLastError();
  int flag = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, ctx->prep_flag, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) {
    ctx->h_ref = flag ? 0 : H;  // a linearization without the contact structure is not usable
    ctx->lin_bad_structure = flag != 0;
  }
  return e;
}

int cimpc_upload_linearization(cimpc_ctx* ctx, int32_t H, const double* z0, const double* th0,
                               const double* r0, const double* rz0, const double* rth0, void* stream) {
  if (!ctx || H <= 0 || !z0 || !th0 || !r0 || !rz0 || !rth0) return CIMPC_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  const LinLayout& l = ctx->entry->lay;
  const size_t nz = l.nz, nth = l.nth;
  CK(alloc_dense(ctx, H));
  const DensePtrs d = dense_ptrs(ctx);
  CK(cudaMemcpyAsync(d.z0, z0, nz * H * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d.r0, r0, nz * H * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d.t0, th0, nth * H * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d.rz, rz0, nz * nz * H * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d.rt, rth0, nz * nth * H * 8, cudaMemcpyHostToDevice, s));
  cudaError_t e = run_prep(ctx, H, s);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "cimpc_upload_linearization");
  if (ctx->lin_bad_structure) return CIMPC_ERR_UNSUPPORTED_MODEL;
  return CIMPC_OK;
}

int cimpc_linearize(cimpc_ctx* ctx, int32_t H, const double* z0, const double* th0, double kappa, void* stream) {
  if (!ctx || H <= 0 || !z0 || !th0) return CIMPC_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  const LinLayout& l = ctx->entry->lay;
  const size_t nz = l.nz, nth = l.nth;
  CK(alloc_dense(ctx, H));
  const DensePtrs d = dense_ptrs(ctx);
  CK(cudaMemcpyAsync(d.z0, z0, nz * H * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d.t0, th0, nth * H * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(d.rz, 0, (nz * nz + nz * nth) * (size_t)H * 8, s));  // rz0 and rθ0 are contiguous
  LinEvalParams lp{H, d.z0, d.t0, kappa, d.r0, d.rz, d.rt};
  cudaError_t e = ctx->entry->linearize(lp, s);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "linearize_kernel launch");
  ctx->launches++;
  e = run_prep(ctx, H, s);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "cimpc_linearize");
  if (ctx->lin_bad_structure) return CIMPC_ERR_UNSUPPORTED_MODEL;
  return CIMPC_OK;
}

int cimpc_get_linearization(cimpc_ctx* ctx, double* r0, double* rz0, double* rth0) {
  if (!ctx) return CIMPC_ERR_INVALID_ARGUMENT;
  if (!ctx->dense || ctx->dense_h <= 0 || ctx->dense_h != ctx->h_ref) return CIMPC_ERR_NOT_INITIALIZED;
  CK(cudaSetDevice(ctx->device));
  const LinLayout& l = ctx->entry->lay;
  const size_t nz = l.nz, nth = l.nth, H = ctx->dense_h;
  const DensePtrs d = dense_ptrs(ctx);
  if (r0) CK(cudaMemcpy(r0, d.r0, nz * H * 8, cudaMemcpyDeviceToHost));
  if (rz0) CK(cudaMemcpy(rz0, d.rz, nz * nz * H * 8, cudaMemcpyDeviceToHost));
  if (rth0) CK(cudaMemcpy(rth0, d.rt, nz * nth * H * 8, cudaMemcpyDeviceToHost));
  return CIMPC_OK;
}

static int check_solve_args(cimpc_ctx* ctx, int64_t n, const void* knot, const void* theta, const void* q2,
                            const cimpc_ip_opts* o, const void* z, const void* dz, const void* st,
                            const void* it) {
  if (!ctx || n < 0 || !o) return CIMPC_ERR_INVALID_ARGUMENT;
  if (!ctx->lin || ctx->h_ref <= 0) return CIMPC_ERR_NOT_INITIALIZED;
  if (n > 0 && (!knot || !theta || !q2 || !z || !st || !it)) return CIMPC_ERR_INVALID_ARGUMENT;
  if (n > 0 && o->diff_sol && !dz) return CIMPC_ERR_INVALID_ARGUMENT;
  if (o->max_iter < 0 || o->max_ls < 0) return CIMPC_ERR_INVALID_ARGUMENT;
  return CIMPC_OK;
}

int cimpc_ip_solve_batch(cimpc_ctx* ctx, int64_t n, const int32_t* knot, const double* theta,
                         const double* q2_init, const double* alt, const cimpc_ip_opts* opts,
                         double* z_out, double* dz_out, uint8_t* status, int32_t* iters, void* stream) {
  int rc = check_solve_args(ctx, n, knot, theta, q2_init, opts, z_out, dz_out, status, iters);
  if (rc != CIMPC_OK) return rc;
  if (n == 0) return CIMPC_OK;
  CK(cudaSetDevice(ctx->device));
  IpParams p;
  p.n = n; p.knot = knot; p.theta = theta; p.q2_init = q2_init; p.alt = alt; p.lin = ctx->lin;
  p.h_ref = ctx->h_ref; p.z_out = z_out; p.dz_out = dz_out; p.status = status; p.iters = iters; p.o = *opts;
  cudaError_t e = ctx->entry->launch(p, ctx->sm_count, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "ip_solve_kernel launch");
  ctx->launches++;
  return CIMPC_OK;
}

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

static bool is_pinned(const void* p) {
  if (!p) return true;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// Host-buffer entry point: the batch is cut into chunks that flow through NS streams, so the H2D copy
// of chunk c+1, the kernel of chunk c and the D2H copy of chunk c−1 overlap.  Buffers that are already
// page-locked (cudaHostAlloc / cudaHostRegister / CUDA.jl `pin`) are DMA'd in place; pageable ones are
// staged through the context's pinned slots.
int cimpc_ip_solve_batch_host(cimpc_ctx* ctx, int64_t n, const int32_t* knot, const double* theta,
                              const double* q2_init, const double* alt, const cimpc_ip_opts* opts,
                              double* z_out, double* dz_out, uint8_t* status, int32_t* iters) {
  return cimpc_ip_solve_batch_host_ex(ctx, n, knot, theta, q2_init, alt, opts, z_out, dz_out, status, iters,
                                      CIMPC_OUT_Z | CIMPC_OUT_DZ | CIMPC_OUT_STATUS);
}

int cimpc_ip_solve_batch_host_ex(cimpc_ctx* ctx, int64_t n, const int32_t* knot, const double* theta,
                                 const double* q2_init, const double* alt, const cimpc_ip_opts* opts,
                                 double* z_out, double* dz_out, uint8_t* status, int32_t* iters, uint32_t out_mask) {
  const bool want_zfull = (out_mask & CIMPC_OUT_Z) != 0, want_zhead = !want_zfull && (out_mask & CIMPC_OUT_ZHEAD) != 0;
  const bool want_dz = (out_mask & CIMPC_OUT_DZ) != 0, want_st = (out_mask & CIMPC_OUT_STATUS) != 0;
  // buffers that are not copied back need not be provided
  int rc = check_solve_args(ctx, n, knot, theta, q2_init, opts, (want_zfull || want_zhead) ? (void*)z_out : (void*)ctx,
                            (want_dz || !opts || !opts->diff_sol) ? (void*)dz_out : (void*)ctx,
                            want_st ? (void*)status : (void*)ctx, want_st ? (void*)iters : (void*)ctx);
  if (rc != CIMPC_OK) return rc;
  if (n == 0) return CIMPC_OK;
  for (int64_t i = 0; i < n; ++i)
    if (knot[i] < -1 || knot[i] >= ctx->h_ref) return CIMPC_ERR_INVALID_ARGUMENT;  // −1 = skip
  CK(cudaSetDevice(ctx->device));
  const LinLayout& l = ctx->entry->lay;
  const int nc = ctx->entry->desc.nc;
  const bool diff = opts->diff_sol != 0;
  const bool copy_dz = diff && want_dz;
  constexpr int NS = cimpc_ctx::NS;
  const int64_t C = n < cimpc_ctx::CHUNK ? n : cimpc_ctx::CHUNK;  // subproblems per chunk
  // per-slot layout (bytes)
  const size_t b_knot = align256(C * 4), b_th = align256(C * l.nth * 8), b_q2 = align256(C * l.nx * 8);
  const size_t b_alt = alt ? align256(C * nc * 8) : 0;
  const size_t b_z = align256(C * l.nz * 8), b_dz = diff ? align256(C * (size_t)l.nd * l.ncol * 8) : 0;
  const size_t b_st = align256(C), b_it = align256(C * 4);
  const size_t o_knot = 0, o_th = o_knot + b_knot, o_q2 = o_th + b_th, o_alt = o_q2 + b_q2;
  const size_t o_z = o_alt + b_alt, o_dz = o_z + b_z, o_st = o_dz + b_dz, o_it = o_st + b_st;
  const size_t slot = o_it + b_it;
  const bool pin_in = is_pinned(knot) && is_pinned(theta) && is_pinned(q2_init) && is_pinned(alt);
  const bool pin_out = (!(want_zfull || want_zhead) || is_pinned(z_out)) && (!copy_dz || is_pinned(dz_out)) &&
                       (!want_st || (is_pinned(status) && is_pinned(iters)));
  if (ctx->dev_bytes < slot * NS) {
    if (ctx->dev) cudaFree(ctx->dev);
    ctx->dev = nullptr; ctx->dev_bytes = 0;
    CK(cudaMalloc(&ctx->dev, slot * NS));
    ctx->dev_bytes = slot * NS;
  }
  if ((!pin_in || !pin_out) && ctx->pin_bytes < slot * NS) {
    if (ctx->pin) cudaFreeHost(ctx->pin);
    ctx->pin = nullptr; ctx->pin_bytes = 0;
    CK(cudaMallocHost(&ctx->pin, slot * NS));
    ctx->pin_bytes = slot * NS;
  }
  const int64_t nchunks = (n + C - 1) / C;
  auto drain = [&](int64_t c) {  // copy staged outputs of chunk c to the caller (pageable outputs only)
    const int64_t lo = c * C, m = (n - lo < C) ? (n - lo) : C;
    const char* hp = (const char*)ctx->pin + (size_t)(c % NS) * slot;
    if (want_zfull) std::memcpy(z_out + lo * l.nz, hp + o_z, m * l.nz * 8);
    if (want_zhead)
      for (int64_t i = 0; i < m; ++i) std::memcpy(z_out + (lo + i) * l.nz, hp + o_z + i * l.nz * 8, l.nd * 8);
    if (copy_dz) std::memcpy(dz_out + lo * (size_t)l.nd * l.ncol, hp + o_dz, m * (size_t)l.nd * l.ncol * 8);
    if (want_st) {
      std::memcpy(status + lo, hp + o_st, m);
      std::memcpy(iters + lo, hp + o_it, m * 4);
    }
  };
  for (int64_t c = 0; c < nchunks; ++c) {
    const int sl = (int)(c % NS);
    cudaStream_t s = ctx->streams[sl];
    if (c >= NS) {
      CK(cudaStreamSynchronize(s));
      if (!pin_out) drain(c - NS);
    }
    const int64_t lo = c * C, m = (n - lo < C) ? (n - lo) : C;
    char* dp = (char*)ctx->dev + (size_t)sl * slot;
    char* hp = ctx->pin ? (char*)ctx->pin + (size_t)sl * slot : nullptr;
    const void *s_knot = knot + lo, *s_th = theta + lo * l.nth, *s_q2 = q2_init + lo * l.nx;
    const void* s_alt = alt ? alt + lo * nc : nullptr;
    if (!pin_in) {
      std::memcpy(hp + o_knot, s_knot, m * 4); s_knot = hp + o_knot;
      std::memcpy(hp + o_th, s_th, m * l.nth * 8); s_th = hp + o_th;
      std::memcpy(hp + o_q2, s_q2, m * l.nx * 8); s_q2 = hp + o_q2;
      if (alt) { std::memcpy(hp + o_alt, s_alt, m * nc * 8); s_alt = hp + o_alt; }
    }
    CK(cudaMemcpyAsync(dp + o_knot, s_knot, m * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dp + o_th, s_th, m * l.nth * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dp + o_q2, s_q2, m * l.nx * 8, cudaMemcpyHostToDevice, s));
    if (alt) CK(cudaMemcpyAsync(dp + o_alt, s_alt, m * nc * 8, cudaMemcpyHostToDevice, s));
    rc = cimpc_ip_solve_batch(ctx, m, (const int32_t*)(dp + o_knot), (const double*)(dp + o_th),
                              (const double*)(dp + o_q2), alt ? (const double*)(dp + o_alt) : nullptr, opts,
                              (double*)(dp + o_z), diff ? (double*)(dp + o_dz) : nullptr, (uint8_t*)(dp + o_st),
                              (int32_t*)(dp + o_it), s);
    if (rc != CIMPC_OK) return rc;
    void* t_z = pin_out ? (void*)(z_out + lo * l.nz) : (void*)(hp + o_z);
    void* t_dz = !diff ? nullptr : (pin_out ? (void*)(dz_out + lo * (size_t)l.nd * l.ncol) : (void*)(hp + o_dz));
    void* t_st = pin_out ? (void*)(status + lo) : (void*)(hp + o_st);
    void* t_it = pin_out ? (void*)(iters + lo) : (void*)(hp + o_it);
    if (want_zfull) CK(cudaMemcpyAsync(t_z, dp + o_z, m * l.nz * 8, cudaMemcpyDeviceToHost, s));
    if (want_zhead)  // only z*[1:nd] of every subproblem (what `d` needs), same nz stride on both sides
      CK(cudaMemcpy2DAsync(t_z, l.nz * 8, dp + o_z, l.nz * 8, l.nd * 8, m, cudaMemcpyDeviceToHost, s));
    if (copy_dz) CK(cudaMemcpyAsync(t_dz, dp + o_dz, m * (size_t)l.nd * l.ncol * 8, cudaMemcpyDeviceToHost, s));
    if (want_st) {
      CK(cudaMemcpyAsync(t_st, dp + o_st, m, cudaMemcpyDeviceToHost, s));
      CK(cudaMemcpyAsync(t_it, dp + o_it, m * 4, cudaMemcpyDeviceToHost, s));
    }
  }
  for (int64_t c = (nchunks > NS ? nchunks - NS : 0); c < nchunks; ++c) {
    CK(cudaStreamSynchronize(ctx->streams[c % NS]));
    if (!pin_out) drain(c);
  }
  return CIMPC_OK;
}


// ------------------------------------------------------------------------------------------
// Device Newton
// ------------------------------------------------------------------------------------------
void cimpc_newton_opts_default(cimpc_newton_opts* o) {
  if (!o) return;
  o->r_tol = 1e-5; o->beta_init = 1e-5; o->max_iter = 10; o->reserved = 0;
}

int32_t cimpc_newton_last_sweeps(const cimpc_ctx* ctx) {
  if (!ctx || !ctx->nw.ready) return 0;
  int32_t v = 0;  // sweeps of the last solve: the loop counter of the graph (synchronises with the device)
  if (cudaSetDevice(ctx->device) != cudaSuccess) return 0;
  if (cudaMemcpy(&v, ctx->nw.ctrl + 3, sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  return v;
}

// Cholesky Q = L Lᵀ of one nq×nq column-major SPD matrix and E = L⁻ᵀ (upper triangular); false if not SPD.
static bool whitening_factor(int nq, const double* Q, double* E) {
  std::vector<double> L((size_t)nq * nq, 0.0), Li((size_t)nq * nq, 0.0);
  for (int j = 0; j < nq; ++j) {
    double d = Q[j + (size_t)j * nq];
    for (int k = 0; k < j; ++k) d -= L[j + (size_t)k * nq] * L[j + (size_t)k * nq];
    if (!(d > 0.0)) return false;
    const double ljj = std::sqrt(d);
    L[j + (size_t)j * nq] = ljj;
    for (int i = j + 1; i < nq; ++i) {
      double v = Q[i + (size_t)j * nq];
      for (int k = 0; k < j; ++k) v -= L[i + (size_t)k * nq] * L[j + (size_t)k * nq];
      L[i + (size_t)j * nq] = v / ljj;
    }
  }
  for (int j = 0; j < nq; ++j) {  // Li = L⁻¹ (lower triangular), column by column
    Li[j + (size_t)j * nq] = 1.0 / L[j + (size_t)j * nq];
    for (int i = j + 1; i < nq; ++i) {
      double v = 0.0;
      for (int k = j; k < i; ++k) v -= L[i + (size_t)k * nq] * Li[k + (size_t)j * nq];
      Li[i + (size_t)j * nq] = v / L[i + (size_t)i * nq];
    }
  }
  for (int i = 0; i < nq; ++i)
    for (int j = 0; j < nq; ++j) E[i + (size_t)j * nq] = Li[j + (size_t)i * nq];  // E = Liᵀ
  return true;
}

static int newton_create_impl(cimpc_ctx* ctx, int32_t H, int64_t R64, const double* obj_q, const double* obj_qd,
                              const double* obj_u, const double* obj_gamma, const double* obj_b, const double* obj_v,
                              const double* v_target, double kappa, const cimpc_newton_opts* nopts,
                              const cimpc_ip_opts* ip_opts);

int cimpc_newton_create_ex(cimpc_ctx* ctx, int32_t H, int64_t R64, const double* obj_q, const double* obj_u,
                           const double* obj_gamma, const double* obj_b, const double* obj_v, double kappa,
                           const cimpc_newton_opts* nopts, const cimpc_ip_opts* ip_opts) {
  if (!obj_q) return CIMPC_ERR_INVALID_ARGUMENT;
  return newton_create_impl(ctx, H, R64, obj_q, nullptr, obj_u, obj_gamma, obj_b, obj_v, nullptr, kappa, nopts, ip_opts);
}

int cimpc_newton_create_dense(cimpc_ctx* ctx, int32_t H, int64_t R64, const double* obj_q_dense, const double* obj_u,
                              const double* obj_v, const double* v_target, double kappa, const cimpc_newton_opts* nopts,
                              const cimpc_ip_opts* ip_opts) {
  if (!obj_q_dense) return CIMPC_ERR_INVALID_ARGUMENT;
  if (ctx && ctx->entry->desc.mode != 0) return CIMPC_ERR_UNSUPPORTED_MODEL;  // :configuration mode only
  return newton_create_impl(ctx, H, R64, nullptr, obj_q_dense, obj_u, nullptr, nullptr, obj_v, v_target, kappa, nopts,
                            ip_opts);
}

static int newton_create_impl(cimpc_ctx* ctx, int32_t H, int64_t R64, const double* obj_q, const double* obj_qd,
                              const double* obj_u, const double* obj_gamma, const double* obj_b, const double* obj_v,
                              const double* v_target, double kappa, const cimpc_newton_opts* nopts,
                              const cimpc_ip_opts* ip_opts) {
  if (!ctx || H < 1 || H > 64 || R64 < 1 || R64 > (1 << 30) / H || (!obj_q && !obj_qd) || !obj_u || !nopts || !ip_opts)
    return CIMPC_ERR_INVALID_ARGUMENT;
  if (!ctx->lin) return CIMPC_ERR_NOT_INITIALIZED;
  const LinLayout& l = ctx->entry->lay;
  const cimpc_model_desc& d = ctx->entry->desc;
  const int R = (int)R64, nq = d.nq, nu = d.nu, nw_ = d.nw, nd = l.nd, nth = l.nth, nz = l.nz, ncol = l.ncol;
  const int nyd = nd - nq;
  // kernel variant: −1 specialised (:configuration, TrackingObjective), 0 / 1 general without / with velocity cost,
  // 2 / 3 dense weights without / with velocity cost
  int variant = obj_qd ? (obj_v ? 3 : 2) : ((d.mode == 0 && !obj_v) ? -1 : (obj_v ? 1 : 0));
  double wmin = 1e300;
  std::vector<double> h_e;
  if (obj_qd) {
    h_e.resize((size_t)H * nq * nq);
    for (int t = 0; t < H; ++t)
      if (!whitening_factor(nq, obj_qd + (size_t)t * nq * nq, h_e.data() + (size_t)t * nq * nq))
        return CIMPC_ERR_INVALID_ARGUMENT;  // obj.q[t] must be symmetric positive definite (dual Schur form)
  } else {
    for (int e = 0; e < H * nq; ++e) wmin = obj_q[e] < wmin ? obj_q[e] : wmin;
  }
  for (int e = 0; e < H * nu; ++e) wmin = obj_u[e] < wmin ? obj_u[e] : wmin;
  if (!(wmin > 0.0)) return CIMPC_ERR_INVALID_ARGUMENT;  // Q must be positive (dual Schur form)
  if (obj_v)
    for (int e = 0; e < H * nq; ++e)
      if (!(obj_v[e] > 0.0)) return CIMPC_ERR_INVALID_ARGUMENT;
  if (nyd > 0) {
    if (!obj_gamma || !obj_b) return CIMPC_ERR_INVALID_ARGUMENT;
    // the force rows are eliminated as zero-weight rows (every example of the reference uses 1e-100): refuse weights
    // that are not negligible against the configuration / control weights
    for (int e = 0; e < H * d.nc; ++e)
      if (!(obj_gamma[e] >= 0.0) || obj_gamma[e] > 1e-30 * wmin) return CIMPC_ERR_UNSUPPORTED_MODEL;
    for (int e = 0; e < H * d.nb; ++e)
      if (!(obj_b[e] >= 0.0) || obj_b[e] > 1e-30 * wmin) return CIMPC_ERR_UNSUPPORTED_MODEL;
  }
  CK(cudaSetDevice(ctx->device));
  auto& nw = ctx->nw;
  newton_release(ctx);
  nw.variant = variant;
  const size_t n = (size_t)H * R;
  const size_t lsc = variant < 0 ? ctx->entry->newton_scratch(H)
                     : (variant < 2 ? ctx->entry->newton_scratch_g[variant](H) : ctx->entry->newton_scratch_d[variant - 2](H));
  // carve one arena
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_tq = take(sizeof(double) * R * (H + 2) * nq), o_tu = take(sizeof(double) * R * H * nu),
               o_nu = take(sizeof(double) * R * H * nd), o_cq = take(sizeof(double) * R * (H + 2) * nq),
               o_cu = take(sizeof(double) * R * H * nu), o_cnu = take(sizeof(double) * R * H * nd),
               o_dl = take(sizeof(double) * R * H * (nu + nq + nd + nyd)), o_rn = take(sizeof(double) * R),
               o_al = take(sizeof(double) * R), o_be = take(sizeof(double) * R), o_ls = take(sizeof(int) * R),
               o_ni = take(sizeof(int) * R), o_sw = take(sizeof(int) * R), o_ph = take(sizeof(int) * R),
               o_kn = take(sizeof(int32_t) * n), o_th = take(sizeof(double) * n * nth), o_q2 = take(sizeof(double) * n * nq),
               o_z = take(sizeof(double) * n * nz), o_dz = take(sizeof(double) * n * nd * ncol), o_st = take(n),
               o_it = take(sizeof(int32_t) * n), o_na = take(sizeof(int)), o_ac = take(sizeof(int32_t) * 4),
               o_al2 = take(sizeof(int32_t) * 2 * R), o_alt = take(sizeof(double) * R * d.nc), o_call = take(sizeof(NewtonCall)),
               o_rq = take(sizeof(double) * (H + 2) * nq),
               o_ru = take(sizeof(double) * H * nu), o_w = take(sizeof(double) * H * (nw_ > 0 ? nw_ : 1)),
               o_win = take(sizeof(int32_t) * (H + 2)), o_oq = take(sizeof(double) * H * nq),
               o_ou = take(sizeof(double) * H * nu),
               o_ty = take(sizeof(double) * R * H * (nyd > 0 ? nyd : 1)), o_cy = take(sizeof(double) * R * H * (nyd > 0 ? nyd : 1)),
               o_ry = take(sizeof(double) * H * (nyd > 0 ? nyd : 1)), o_oy = take(sizeof(double) * H * (nyd > 0 ? nyd : 1)),
               o_ov = take(sizeof(double) * H * nq),
               o_qd = take(sizeof(double) * (obj_qd ? (size_t)H * nq * nq : 1)),
               o_qe = take(sizeof(double) * (obj_qd ? (size_t)H * nq * nq : 1)),
               o_qt = take(sizeof(double) * H * nq), o_vt = take(sizeof(double) * H * nq),
               o_qi = take(sizeof(double) * H * nq), o_ui = take(sizeof(double) * H * nu),
               o_qsi = take(sizeof(double) * H * nq), o_usi = take(sizeof(double) * H * nu),
               o_lsc = take(sizeof(double) * R * lsc);
  CK(cudaMalloc(&nw.arena, off));
  CK(cudaMemset(nw.arena, 0, off));
  char* b = (char*)nw.arena;
  NewtonParams& p = nw.p;
  p = NewtonParams{};
  p.R = R; p.H = H;
  p.traj_q = (double*)(b + o_tq); p.traj_u = (double*)(b + o_tu); p.nu = (double*)(b + o_nu);
  p.cand_q = (double*)(b + o_cq); p.cand_u = (double*)(b + o_cu); p.cand_nu = (double*)(b + o_cnu);
  p.delta = (double*)(b + o_dl); p.r_norm = (double*)(b + o_rn); p.alpha = (double*)(b + o_al);
  p.beta = (double*)(b + o_be); p.ls_it = (int*)(b + o_ls); p.newton_it = (int*)(b + o_ni);
  p.sweeps = (int*)(b + o_sw); p.phase = (int*)(b + o_ph); p.knot = (int32_t*)(b + o_kn);
  p.theta = (double*)(b + o_th); p.q2 = (double*)(b + o_q2);
  nw.z = (double*)(b + o_z); nw.dz = (double*)(b + o_dz); nw.status = (uint8_t*)(b + o_st); nw.iters = (int32_t*)(b + o_it);
  p.z = nw.z; p.dz = nw.dz; p.n_active = (int*)(b + o_na);
  nw.ctrl = (int32_t*)(b + o_ac);
  p.act_count = nw.ctrl; p.par = nw.ctrl + 2; p.sweep_ctr = nw.ctrl + 3; p.act_list = (int32_t*)(b + o_al2);
  p.alt = (double*)(b + o_alt);
  nw.call = (NewtonCall*)(b + o_call);
  p.call = nw.call;
  nw.ref_q = (double*)(b + o_rq); nw.ref_u = (double*)(b + o_ru); nw.w = (double*)(b + o_w);
  nw.window = (int32_t*)(b + o_win); nw.obj_q = (double*)(b + o_oq); nw.obj_u = (double*)(b + o_ou);
  p.ref_q = nw.ref_q; p.ref_u = nw.ref_u; p.w = nw.w; p.window = nw.window; p.obj_q = nw.obj_q; p.obj_u = nw.obj_u;
  nw.ref_y = (double*)(b + o_ry); nw.obj_y = (double*)(b + o_oy); nw.obj_v = (double*)(b + o_ov);
  if (nyd > 0) {
    p.traj_y = (double*)(b + o_ty); p.cand_y = (double*)(b + o_cy); p.ref_y = nw.ref_y; p.obj_y = nw.obj_y;
    // obj_y = [obj.γ; obj.b] per stage
    CK(cudaMemcpy2D(nw.obj_y, sizeof(double) * nyd, obj_gamma, sizeof(double) * d.nc, sizeof(double) * d.nc, H, cudaMemcpyHostToDevice));
    CK(cudaMemcpy2D(nw.obj_y + d.nc, sizeof(double) * nyd, obj_b, sizeof(double) * d.nb, sizeof(double) * d.nb, H, cudaMemcpyHostToDevice));
  }
  if (obj_v) {
    p.obj_v = nw.obj_v;
    CK(cudaMemcpy(nw.obj_v, obj_v, sizeof(double) * H * nq, cudaMemcpyHostToDevice));
  }
  nw.lscratch = (double*)(b + o_lsc);
  p.kappa = kappa; p.r_tol = nopts->r_tol; p.beta_init = nopts->beta_init; p.max_iter = nopts->max_iter;
  if (obj_q) CK(cudaMemcpy(nw.obj_q, obj_q, sizeof(double) * H * nq, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(nw.obj_u, obj_u, sizeof(double) * H * nu, cudaMemcpyHostToDevice));
  {  // reciprocal weights (the specialised kernel reads Q⁻¹ from here)
    std::vector<double> qi((size_t)H * nq, 0.0), ui((size_t)H * nu);
    if (obj_q)
      for (int e = 0; e < H * nq; ++e) qi[e] = 1.0 / obj_q[e];
    for (int e = 0; e < H * nu; ++e) ui[e] = 1.0 / obj_u[e];
    p.obj_qi = (double*)(b + o_qi); p.obj_ui = (double*)(b + o_ui);
    CK(cudaMemcpy((void*)p.obj_qi, qi.data(), sizeof(double) * H * nq, cudaMemcpyHostToDevice));
    CK(cudaMemcpy((void*)p.obj_ui, ui.data(), sizeof(double) * H * nu, cudaMemcpyHostToDevice));
    for (double& v : qi) v = std::sqrt(v);
    for (double& v : ui) v = std::sqrt(v);
    p.obj_qsi = (double*)(b + o_qsi); p.obj_usi = (double*)(b + o_usi);
    CK(cudaMemcpy((void*)p.obj_qsi, qi.data(), sizeof(double) * H * nq, cudaMemcpyHostToDevice));
    CK(cudaMemcpy((void*)p.obj_usi, ui.data(), sizeof(double) * H * nu, cudaMemcpyHostToDevice));
  }
  if (obj_qd) {
    p.obj_qd = (double*)(b + o_qd); p.obj_e = (double*)(b + o_qe);
    CK(cudaMemcpy((void*)p.obj_qd, obj_qd, sizeof(double) * (size_t)H * nq * nq, cudaMemcpyHostToDevice));
    CK(cudaMemcpy((void*)p.obj_e, h_e.data(), sizeof(double) * (size_t)H * nq * nq, cudaMemcpyHostToDevice));
    bool any_vt = false;
    if (v_target)
      for (int e = 0; e < H * nq; ++e) any_vt = any_vt || v_target[e] != 0.0;
    if (any_vt) {  // q_target[t] = Σ_{s<t} v_target[s]  (objective.jl:36-44)
      std::vector<double> qt((size_t)H * nq, 0.0);
      for (int t = 1; t < H; ++t)
        for (int k = 0; k < nq; ++k) qt[(size_t)t * nq + k] = qt[(size_t)(t - 1) * nq + k] + v_target[(size_t)(t - 1) * nq + k];
      p.q_tgt = (double*)(b + o_qt); p.v_tgt = (double*)(b + o_vt);
      CK(cudaMemcpy((void*)p.q_tgt, qt.data(), sizeof(double) * H * nq, cudaMemcpyHostToDevice));
      CK(cudaMemcpy((void*)p.v_tgt, v_target, sizeof(double) * H * nq, cudaMemcpyHostToDevice));
    }
  }
  if (!nw.h_active) CK(cudaMallocHost(&nw.h_active, sizeof(int)));
  nw.ip = *ip_opts;
  nw.ip.diff_sol = 1;
  nw.no = *nopts;
  CK(cudaMallocHost(&nw.ring, sizeof(NewtonCall) * cimpc_ctx::NewtonState::RING));
  for (auto& ev : nw.ring_ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  nw.ring_pos = 0;
  int rc = newton_build_graph(ctx);
  if (rc != CIMPC_OK) return rc;
  nw.ready = true;
  return CIMPC_OK;
}

int cimpc_newton_create(cimpc_ctx* ctx, int32_t H, int64_t R64, const double* obj_q, const double* obj_u,
                        double kappa, const cimpc_newton_opts* nopts, const cimpc_ip_opts* ip_opts) {
  if (ctx && ctx->entry->desc.mode != 0) return CIMPC_ERR_UNSUPPORTED_MODEL;  // needs the γ, b weights: use _ex
  return cimpc_newton_create_ex(ctx, H, R64, obj_q, obj_u, nullptr, nullptr, nullptr, kappa, nopts, ip_opts);
}

int cimpc_newton_solve_batch(cimpc_ctx* ctx, const int32_t* window, const double* ref_q, const double* ref_u,
                             double mu, double h, const double* q0, const double* q1, const uint8_t* active,
                             int32_t warm_start, double* u_out, double* q_out, int32_t* info, void* stream) {
  return cimpc_newton_solve_batch_ex(ctx, window, ref_q, ref_u, nullptr, nullptr, mu, h, q0, q1, active, warm_start, u_out,
                                     q_out, nullptr, info, stream);
}

// Upload the shared per-call data (window, rotated reference), fill the NewtonCall slot and launch the solve graph.
static int newton_launch(cimpc_ctx* ctx, const int32_t* window, const double* ref_q, const double* ref_u,
                         const double* ref_gamma, const double* ref_b, const NewtonCall& call_in, cudaStream_t s) {
  auto& nw = ctx->nw;
  NewtonParams& p = nw.p;
  const cimpc_model_desc& d = ctx->entry->desc;
  const int H = p.H;
  const int nyd_ = ctx->entry->lay.nd - d.nq;
  for (int t = 0; t < H; ++t)
    if (window[t] < 0 || window[t] >= ctx->h_ref) return CIMPC_ERR_INVALID_ARGUMENT;
  CK(cudaMemcpyAsync(nw.window, window, sizeof(int32_t) * (H + 2), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(nw.ref_q, ref_q, sizeof(double) * (H + 2) * d.nq, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(nw.ref_u, ref_u, sizeof(double) * H * d.nu, cudaMemcpyHostToDevice, s));
  if (nyd_ > 0) {  // ref_y = [γ_ref; b_ref] per stage
    CK(cudaMemcpy2DAsync(nw.ref_y, sizeof(double) * nyd_, ref_gamma, sizeof(double) * d.nc, sizeof(double) * d.nc, H,
                         cudaMemcpyHostToDevice, s));
    CK(cudaMemcpy2DAsync(nw.ref_y + d.nc, sizeof(double) * nyd_, ref_b, sizeof(double) * d.nb, sizeof(double) * d.nb, H,
                         cudaMemcpyHostToDevice, s));
  }
  // per-call pointers: pinned ring slot → device struct (a slot is reused only after its copy has completed)
  const int slot = nw.ring_pos;
  nw.ring_pos = (nw.ring_pos + 1) % cimpc_ctx::NewtonState::RING;
  CK(cudaEventSynchronize(nw.ring_ev[slot]));
  nw.ring[slot] = call_in;
  CK(cudaMemcpyAsync(nw.call, &nw.ring[slot], sizeof(NewtonCall), cudaMemcpyHostToDevice, s));
  CK(cudaEventRecord(nw.ring_ev[slot], s));
  if (std::getenv("CIMPC_NEWTON_HOSTLOOP")) {
    // PROFILING AID (not a fallback: same kernels, same arithmetic): the sweeps of the graph driven from the host with a
    // 16-byte read-back per sweep, so that a launch list shows only the sweeps that did work.
    NewtonParams& p = nw.p;
    const cimpc_model_desc& d = ctx->entry->desc;
    const int H = p.H, R = p.R, nyd = ctx->entry->lay.nd - d.nq;
    CK(cudaMemsetAsync(nw.ctrl, 0, sizeof(int32_t) * 4, s));
    cudaError_t e = ctx->entry->newton_reset(p, nw.call, s);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "newton_reset_kernel launch");
    int32_t h_ctrl[4] = {1, 0, 0, 0};
    for (int round = 0; round < nw.max_rounds; ++round) {
      CK(cudaMemcpyAsync(h_ctrl, nw.ctrl, sizeof(h_ctrl), cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      if (h_ctrl[h_ctrl[2]] <= 0) break;
      int rc = newton_sweep_launch(ctx, s);
      if (rc != CIMPC_OK) return rc;
    }
    const int len = H * (d.nu + d.nq + nyd + ctx->entry->lay.nd);
    newton_finish_kernel<<<(R + 127) / 128, 128, 0, s>>>(p, nw.call, d.nq, d.nu, nyd, p.r_tol, len);
    CK(cudaGetLastError());
    return CIMPC_OK;
  }
  CK(cudaGraphLaunch(nw.exec, s));
  ctx->launches += 2 + 3 * nw.max_rounds;  // kernel nodes of the graph
  nw.last_sweeps = -1; // read back lazily
  return CIMPC_OK;
}

int cimpc_newton_solve_batch_ex2(cimpc_ctx* ctx, const int32_t* window, const double* ref_q, const double* ref_u,
                                 const double* ref_gamma, const double* ref_b, double mu, double h, const double* q0,
                                 const double* q1, const uint8_t* active, const double* alt, int32_t warm_start,
                                 double* u_out, double* q_out, double* y_out, int32_t* info, void* stream) {
  if (!ctx || !window || !ref_q || !ref_u || !q0 || !q1 || !u_out) return CIMPC_ERR_INVALID_ARGUMENT;
  auto& nw = ctx->nw;
  if (!nw.ready) return CIMPC_ERR_NOT_INITIALIZED;
  const int nyd_ = ctx->entry->lay.nd - ctx->entry->desc.nq;
  if (nyd_ > 0 && (!ref_gamma || !ref_b)) return CIMPC_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  NewtonCall c{};
  c.q0 = q0; c.q1 = q1; c.active = active; c.alt = alt; c.u_out = u_out; c.q_out = q_out; c.y_out = y_out; c.info = info;
  c.warm_start = warm_start ? 1 : 0; c.mu = mu; c.h = h;
  return newton_launch(ctx, window, ref_q, ref_u, ref_gamma, ref_b, c, (cudaStream_t)stream);
}

int cimpc_newton_solve_batch_ex(cimpc_ctx* ctx, const int32_t* window, const double* ref_q, const double* ref_u,
                                const double* ref_gamma, const double* ref_b, double mu, double h, const double* q0,
                                const double* q1, const uint8_t* active, int32_t warm_start, double* u_out,
                                double* q_out, double* y_out, int32_t* info, void* stream) {
  return cimpc_newton_solve_batch_ex2(ctx, window, ref_q, ref_u, ref_gamma, ref_b, mu, h, q0, q1, active, nullptr,
                                      warm_start, u_out, q_out, y_out, info, stream);
}

// HOST buffers in, host buffers out: the end-to-end boundary of the MPC path (policy.jl:118-120 hands `newton_solve!`
// two configurations per rollout and takes one control back: (2 nq + nu)·8 bytes per rollout cross the bus).
int cimpc_newton_solve_batch_host(cimpc_ctx* ctx, const int32_t* window, const double* ref_q, const double* ref_u,
                                  const double* ref_gamma, const double* ref_b, double mu, double h, const double* q0,
                                  const double* q1, const uint8_t* active, const double* alt, int32_t warm_start,
                                  double* u_out, int32_t* info) {
  if (!ctx || !window || !ref_q || !ref_u || !q0 || !q1 || !u_out) return CIMPC_ERR_INVALID_ARGUMENT;
  auto& nw = ctx->nw;
  if (!nw.ready) return CIMPC_ERR_NOT_INITIALIZED;
  const cimpc_model_desc& d = ctx->entry->desc;
  const int nyd_ = ctx->entry->lay.nd - d.nq;
  if (nyd_ > 0 && (!ref_gamma || !ref_b)) return CIMPC_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  const size_t R = (size_t)nw.p.R, nq = d.nq, nu = d.nu, nc = d.nc;
  // device block: q0 | q1 | u | alt | info(4 int32 = 2 doubles) | active (R bytes)
  const size_t n_d = R * (2 * nq + nu + nc + 2) + (R + 7) / 8;
  if (nw.hq_doubles < n_d) {
    if (nw.hq) cudaFreeHost(nw.hq);
    if (nw.dq) cudaFree(nw.dq);
    nw.hq = nullptr; nw.dq = nullptr; nw.hq_doubles = 0;
    CK(cudaMallocHost(&nw.hq, n_d * sizeof(double)));
    CK(cudaMalloc(&nw.dq, n_d * sizeof(double)));
    nw.hq_doubles = n_d;
  }
  cudaStream_t s = ctx->streams[0];
  double *d_q0 = nw.dq, *d_q1 = d_q0 + R * nq, *d_u = d_q1 + R * nq, *d_alt = d_u + R * nu;
  int32_t* d_info = (int32_t*)(d_alt + R * nc);
  uint8_t* d_act = (uint8_t*)(d_info + 4 * R);
  const bool pin_in = is_pinned(q0) && is_pinned(q1);
  const double *s_q0 = q0, *s_q1 = q1;
  if (!pin_in) {  // pageable caller buffers are staged through the context's pinned block
    std::memcpy(nw.hq, q0, R * nq * 8);
    std::memcpy(nw.hq + R * nq, q1, R * nq * 8);
    s_q0 = nw.hq; s_q1 = nw.hq + R * nq;
  }
  CK(cudaMemcpyAsync(d_q0, s_q0, R * nq * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_q1, s_q1, R * nq * 8, cudaMemcpyHostToDevice, s));
  if (alt) CK(cudaMemcpyAsync(d_alt, alt, R * nc * 8, cudaMemcpyHostToDevice, s));
  if (active) CK(cudaMemcpyAsync(d_act, active, R, cudaMemcpyHostToDevice, s));
  NewtonCall c{};
  c.q0 = d_q0; c.q1 = d_q1; c.active = active ? d_act : nullptr; c.alt = alt ? d_alt : nullptr; c.u_out = d_u;
  c.info = info ? d_info : nullptr; c.warm_start = warm_start ? 1 : 0; c.mu = mu; c.h = h;
  int rc = newton_launch(ctx, window, ref_q, ref_u, ref_gamma, ref_b, c, s);
  if (rc != CIMPC_OK) return rc;
  const bool pin_out = is_pinned(u_out);
  double* t_u = pin_out ? u_out : nw.hq + 2 * R * nq;
  CK(cudaMemcpyAsync(t_u, d_u, R * nu * 8, cudaMemcpyDeviceToHost, s));
  if (info) CK(cudaMemcpyAsync(info, d_info, R * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (!pin_out) std::memcpy(u_out, t_u, R * nu * 8);
  return CIMPC_OK;
}


int cimpc_sim_step_batch(cimpc_ctx* ctx, int64_t n, const double* q0, const double* q1, const double* u, const double* w,
                         const uint8_t* active, double mu, double h, const cimpc_ip_opts* opts, double* q2, double* gamma,
                         double* b,
                         uint8_t* status, int32_t* iters, void* stream) {
  return cimpc_sim_step_batch_ex(ctx, n, q0, q1, u, w, active, mu, h, opts, q2, gamma, b, nullptr, status, iters, stream);
}

static int sim_steps_impl(cimpc_ctx* ctx, int64_t n, int32_t n_steps, bool multi, const double* q0, const double* q1, const double* u,
                          const double* w, const uint8_t* active, double mu, double h, const cimpc_ip_opts* opts, double* q2,
                          double* gamma, double* b, double* phi, uint8_t* status, int32_t* iters, void* stream) {
  if (!ctx || n < 0 || n > (1 << 30) || !opts || n_steps < 1 || n_steps > 4096) return CIMPC_ERR_INVALID_ARGUMENT;
  if (n == 0) return CIMPC_OK;
  if (!q0 || !q1 || !u || !q2 || !gamma || !b || !status || !iters) return CIMPC_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  const size_t need = ctx->entry->sim_scratch((int)n);
  if (ctx->sim_scratch_doubles < need) {
    if (ctx->sim_scratch) cudaFree(ctx->sim_scratch);
    ctx->sim_scratch = nullptr; ctx->sim_scratch_doubles = 0;
    CK(cudaMalloc(&ctx->sim_scratch, need * sizeof(double)));
    ctx->sim_scratch_doubles = need;
  }
  SimParams p;
  p.R = (int)n; p.q0 = q0; p.q1 = q1; p.u = u; p.w = w; p.active = active; p.mu = mu; p.h = h; p.o = *opts;
  p.q2_out = q2; p.gamma_out = gamma; p.b_out = b; p.phi_out = phi; p.status = status; p.iters = iters; p.scratch = ctx->sim_scratch;
  p.nsteps = n_steps;
  p.multi = multi ? 1 : 0;
  cudaError_t e = ctx->entry->sim_step(p, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "sim_step_kernel launch");
  ctx->launches++;
  return CIMPC_OK;
}

int cimpc_sim_step_batch_ex(cimpc_ctx* ctx, int64_t n, const double* q0, const double* q1, const double* u, const double* w,
                            const uint8_t* active, double mu, double h, const cimpc_ip_opts* opts, double* q2,
                            double* gamma, double* b, double* phi, uint8_t* status, int32_t* iters, void* stream) {
  return sim_steps_impl(ctx, n, 1, false, q0, q1, u, w, active, mu, h, opts, q2, gamma, b, phi, status, iters, stream);
}

int cimpc_sim_steps_batch(cimpc_ctx* ctx, int64_t n, int32_t n_steps, const double* q0, const double* q1, const double* u,
                          const double* w, const uint8_t* active, double mu, double h, const cimpc_ip_opts* opts, double* q2,
                          double* gamma, double* b, double* phi, uint8_t* status, int32_t* iters, void* stream) {
  // (n_steps = 1 keeps the multi-step conventions: failed / inactive rollouts get defined outputs)
  return sim_steps_impl(ctx, n, n_steps, true, q0, q1, u, w, active, mu, h, opts, q2, gamma, b, phi, status, iters, stream);
}


// ------------------------------------------------------------------------------------------
// Multi-GPU: final trajectory gather over NCCL (the only collective of the path)
// ------------------------------------------------------------------------------------------
int cimpc_nccl_get_unique_id(uint8_t* id128) {
  if (!id128) return CIMPC_ERR_INVALID_ARGUMENT;
  NcclApi& a = nccl_api();
  if (!a.ok) return CIMPC_ERR_NO_DEVICE;
  return a.GetUniqueId(id128) == 0 ? CIMPC_OK : CIMPC_ERR_CUDA;
}

int cimpc_comm_init(cimpc_ctx* ctx, int32_t world, int32_t rank, const uint8_t* id128) {
  if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return CIMPC_ERR_INVALID_ARGUMENT;
  NcclApi& a = nccl_api();
  if (!a.ok) {
    ctx->cuda_err = "libnccl.so.2 not found";
    return CIMPC_ERR_NO_DEVICE;
  }
  CK(cudaSetDevice(ctx->device));
  if (ctx->nccl_comm) {
    a.CommDestroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  NcclApi::Id id;
  std::memcpy(id.b, id128, 128);
  int rc = a.CommInitRank(&ctx->nccl_comm, world, id, rank);
  if (rc != 0) return nccl_fail(ctx, rc, "ncclCommInitRank");
  ctx->nccl_world = world;
  ctx->nccl_rank = rank;
  return CIMPC_OK;
}

int cimpc_gather(cimpc_ctx* ctx, void* comm, const double* local, int64_t count, double* out, const int64_t* counts,
                 int32_t root, void* stream) {
  if (!ctx || count < 0 || (count > 0 && !local)) return CIMPC_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  NcclApi& a = nccl_api();
  void* c = comm ? comm : ctx->nccl_comm;
  int world = 1, rank = 0;
  if (c) {
    if (!a.ok) return CIMPC_ERR_NO_DEVICE;
    int rc = a.CommCount(c, &world);
    if (rc == 0) rc = a.CommUserRank(c, &rank);
    if (rc != 0) return nccl_fail(ctx, rc, "ncclCommCount");
  }
  if (root < 0 || root >= world) return CIMPC_ERR_INVALID_ARGUMENT;
  if (world == 1) {  // single rank: the gather is a copy
    if (!out) return CIMPC_ERR_INVALID_ARGUMENT;
    if (count > 0 && out != local) CK(cudaMemcpyAsync(out, local, sizeof(double) * count, cudaMemcpyDeviceToDevice, s));
    return CIMPC_OK;
  }
  if (rank == root && !out) return CIMPC_ERR_INVALID_ARGUMENT;
  // gather-to-root with point-to-point transfers inside one group: the root posts one receive per peer at that
  // peer's offset (NVLink / NVSwitch carry them concurrently); nobody but the root ever holds the other shards
  int rc = a.GroupStart();
  if (rc != 0) return nccl_fail(ctx, rc, "ncclGroupStart");
  if (rank == root) {
    int64_t off = 0;
    for (int r = 0; r < world; ++r) {
      const int64_t cr = counts ? counts[r] : count;
      if (cr < 0) { a.GroupEnd(); return CIMPC_ERR_INVALID_ARGUMENT; }
      if (r == rank) {
        if (cr != count) { a.GroupEnd(); return CIMPC_ERR_INVALID_ARGUMENT; }
        if (count > 0 && out + off != local)
          if (cudaMemcpyAsync(out + off, local, sizeof(double) * count, cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
            a.GroupEnd();
            return cuda_fail(ctx, cudaGetLastError(), "cimpc_gather local copy");
          }
      } else if (cr > 0) {
        rc = a.Recv(out + off, (size_t)cr, NCCL_FLOAT64, r, c, s);
        if (rc != 0) { a.GroupEnd(); return nccl_fail(ctx, rc, "ncclRecv"); }
      }
      off += cr;
    }
  } else if (count > 0) {
    rc = a.Send(local, (size_t)count, NCCL_FLOAT64, root, c, s);
    if (rc != 0) { a.GroupEnd(); return nccl_fail(ctx, rc, "ncclSend"); }
  }
  rc = a.GroupEnd();
  if (rc != 0) return nccl_fail(ctx, rc, "ncclGroupEnd");
  return CIMPC_OK;
}

}  // extern "C"
