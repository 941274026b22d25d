// csrc/gmres.cu -- device-resident restarted GMRES with per-iteration relaxation of the expansion order: the caller
// of the hot path (reference examples/BEM/GMRES.hpp:142-252 with examples/BEM/SolverOptions.hpp:25-38).
//
// Same algorithm and decisions as the reference: modified Gram-Schmidt, Givens rotations on host scalars,
// convergence on the rotated residual estimate |s[i+1]| / ||b||, and before every inner matvec
//     p = max(1, predict_p(|resid|));  kernel.set_p(p)
// (GMRES_Stokes.hpp:229: p = max(p_min, predict_p(|resid|) - 1), selected by fmmb_solver_options.p_min / p_offset)
// (the first matvec of a restart cycle runs at whatever order the kernel was left at, GMRES.hpp:174-175).
// What moves: the Krylov basis, the work vectors and every BLAS-1 operation live on the GPU, the matvec is fed
// from device-resident vectors (fmmb_plan_execute_device), and an inner iteration costs ONE host synchronisation:
// the Gram-Schmidt coefficients stay in device memory between the dot product that produces them and the axpy
// that consumes them, and come back to the host as one column when the iteration's norm is needed.
#include "common.cuh"
#include <cmath>
#include <cstring>

namespace fmmb {

// Solver workspace, kept on the plan between solves (a cudaMalloc / cudaFree pair per Krylov vector would cost more
// than the BLAS-1 work of a whole solve at BEM sizes).
struct GmresWorkspace {
  DevBuf<double> x, b, w, z, diag, scal, partial, basis, zbasis;   // (z)basis: vectors back to back, grown geometrically
  DevBuf<unsigned int> counter;
  DevBuf<double> part2;               // mgs_sweep_kernel: two rows of partial sums
  DevBuf<unsigned int> bar;           // its grid-barrier counter (never reset) and the value the host knows it has
  unsigned int bar_count = 0;
  bool mgs_cooperative = true;        // cleared when a cooperative launch is refused
};
void gmres_free(GmresWorkspace* w) { delete w; }

namespace {

constexpr int kDotBlocks = 32;

// Thread sums -> block sum (tree in shared memory) -> partial[block]; the last block to arrive adds the partials in
// index order (the result does not depend on which block is last).  The partials are fetched by as many threads in
// parallel and summed from shared memory: one thread reading them one after the other from L2 cost ~10 us per dot
// product, most of what a Gram-Schmidt step took at BEM sizes.
__device__ __forceinline__ void block_dot_finish(double s, double* sh, int* last, double* __restrict__ partial,
                                                 unsigned int* __restrict__ counter, double* __restrict__ out) {
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = sh[0];
    __threadfence();
    *last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (*last) {                                   // block-uniform
    if (threadIdx.x == 0) __threadfence();
    __syncthreads();
    if (threadIdx.x < gridDim.x) sh[threadIdx.x] = ((volatile double*)partial)[threadIdx.x];   // gridDim.x <= 256
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0;
      for (unsigned k = 0; k < gridDim.x; ++k) t += sh[k];
      *out = t;
      *counter = 0;
    }
  }
}

// out[slot] = sum a[i] b[i]: block partial sums, the last block to finish adds them in index order
// (deterministic: the result does not depend on which block is last)
__global__ void __launch_bounds__(256)
dot_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* __restrict__ partial,
           unsigned int* __restrict__ counter, double* __restrict__ out) {
  __shared__ double sh[256];
  __shared__ int last;
  double s = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += a[i] * b[i];
  block_dot_finish(s, sh, &last, partial, counter, out);
}
// One modified Gram-Schmidt step in one launch: w -= (*coef) v_prev (skipped for v_prev == nullptr), then
// out = <w, v_next> (v_next == nullptr: <w, w>).  Element for element the arithmetic of axpy_dev_kernel followed by
// dot_kernel (same grid, same strides, same order of the partial sums), so the Hessenberg entries keep their bits.
__global__ void __launch_bounds__(256)
axpy_dot_kernel(const double* __restrict__ vprev, const double* __restrict__ coef, double* __restrict__ w,
                const double* __restrict__ vnext, int64_t n, double* __restrict__ partial,
                unsigned int* __restrict__ counter, double* __restrict__ out) {
  __shared__ double sh[256];
  __shared__ int last;
  const double alpha = vprev ? -1.0 * *coef : 0.0;
  double s = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double wi = w[i];
    if (vprev) { wi = fma(alpha, vprev[i], wi); w[i] = wi; }
    s += wi * (vnext ? vnext[i] : wi);
  }
  block_dot_finish(s, sh, &last, partial, counter, out);
}
// The whole modified Gram-Schmidt sweep of one iteration in ONE launch (cooperative: all kDotBlocks blocks are
// resident): step k subtracts the projection on V[k-1] and takes the product with V[k] (k = nvec: |w|^2), the blocks
// meet at a grid barrier, every block adds the partial sums in index order, and at the end V[nvec] = w / |w|.  Same
// strides, same partial sums, same order as the chain of axpy_dot_kernel launches it replaces (same bits); what goes
// away is a launch + drain per step (~4.5 us against a ~1.5 us barrier).
//   bar:  arrival counter, never reset; the host passes the value it has before this launch (bar_base)
//   part: two rows of kDotBlocks partial sums, used alternately (a block that is one barrier ahead writes the other row)
// A barrier that does not complete within kGridBarrierTimeoutClocks writes NaN coefficients (the host turns that into an
// error) instead of spinning forever.
constexpr long long kGridBarrierTimeoutClocks = 6000000000ll;    // ~3 s of SM clock

__global__ void __launch_bounds__(256)
mgs_sweep_kernel(double* __restrict__ w, const double* __restrict__ basis, int nvec, int64_t n,
                 double* __restrict__ vnext, double* __restrict__ part, unsigned int* __restrict__ bar,
                 unsigned int bar_base, double* __restrict__ scal) {
  __shared__ double sh[256];
  __shared__ double h_sh;
  __shared__ int dead;
  if (threadIdx.x == 0) dead = 0;
  const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  double h_prev = 0.0;
  for (int k = 0; k <= nvec; ++k) {
    const double* vp = k ? basis + (size_t)n * (k - 1) : nullptr;
    const double* vn = k < nvec ? basis + (size_t)n * k : nullptr;
    const double alpha = -1.0 * h_prev;
    double s = 0;
    for (int64_t i = i0; i < n; i += stride) {
      double wi = w[i];
      if (vp) { wi = fma(alpha, vp[i], wi); w[i] = wi; }
      s += wi * (vn ? vn[i] : wi);
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int t = 128; t > 0; t >>= 1) {
      if ((int)threadIdx.x < t) sh[threadIdx.x] += sh[threadIdx.x + t];
      __syncthreads();
    }
    double* row = part + (size_t)(k & 1) * gridDim.x;
    if (threadIdx.x == 0) {
      row[blockIdx.x] = sh[0];
      __threadfence();
      atomicAdd(bar, 1u);
      const unsigned int target = bar_base + (unsigned)(k + 1) * gridDim.x;
      const long long t0 = clock64();
      while ((int)(*(volatile unsigned int*)bar - target) < 0) {
        if (clock64() - t0 > kGridBarrierTimeoutClocks) { dead = 1; break; }
      }
      __threadfence();
    }
    __syncthreads();
    if (threadIdx.x < gridDim.x) sh[threadIdx.x] = ((volatile double*)row)[threadIdx.x];   // gridDim.x <= 256
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0;
      for (unsigned b = 0; b < gridDim.x; ++b) t += sh[b];
      if (dead) t = __longlong_as_double(0x7ff8000000000000ll);
      h_sh = t;
      if (blockIdx.x == 0) scal[k] = t;
    }
    __syncthreads();
    h_prev = h_sh;
  }
  const double inv = 1.0 / sqrt(h_prev);
  for (int64_t i = i0; i < n; i += stride) vnext[i] = w[i] * inv;
}

// out = w / sqrt(*coef)
__global__ void scale_into_kernel(const double* __restrict__ w, double* __restrict__ out, int64_t n,
                                  const double* __restrict__ coef) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = w[i] * (1.0 / sqrt(*coef));
}
// y += alpha x with alpha = sign * (*coef) read from device memory (sqrt / reciprocal variants for the norms)
// mode 0: alpha = sign * coef;  1: y = y * (sign / sqrt(coef)) (x unused)
__global__ void axpy_dev_kernel(const double* __restrict__ x, double* __restrict__ y, int64_t n,
                                const double* __restrict__ coef, double sign, int mode) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mode == 0) y[i] = fma(sign * *coef, x[i], y[i]);
  else y[i] *= sign / sqrt(*coef);
}
__global__ void axpy_kernel(const double* __restrict__ x, double* __restrict__ y, int64_t n, double alpha) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) y[i] = fma(alpha, x[i], y[i]);
}
// z = d .* v (diagonal preconditioner) or z = v
__global__ void precond_kernel(const double* __restrict__ v, const double* __restrict__ d, double* __restrict__ z, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) z[i] = d ? d[i] * v[i] : v[i];
}
inline int nblk(int64_t n, int t) { return (int)((n + t - 1) / t); }

unsigned predict_p(const fmmb_solver_options& o, double eps) {      // SolverOptions::predict_p (:25-38)
  if (!o.variable_p) return o.max_p;
  if (o.relax_type == 0) {
    const double alpha = 1. / std::min(eps, 1.);
    const double nu = std::min(alpha * o.residual, 1.);
    return std::min((unsigned)std::ceil(-std::log2(nu)), o.max_p);
  }
  return std::min((unsigned)std::ceil(-std::log2(eps)), o.max_p);
}
inline void apply_rotation(double& dx, double& dy, double cs, double sn) {
  const double t = cs * dx + sn * dy;
  dy = -sn * dx + cs * dy;
  dx = t;
}
inline void make_rotation(double dx, double dy, double& cs, double& sn) {
  if (dy == 0.0) { cs = 1.0; sn = 0.0; }
  else if (std::fabs(dy) > std::fabs(dx)) { const double t = dx / dy; sn = 1.0 / std::sqrt(1.0 + t * t); cs = t * sn; }
  else { const double t = dy / dx; cs = 1.0 / std::sqrt(1.0 + t * t); sn = t * cs; }
}

}  // namespace

void run_matvec_for_solver(fmmb_plan* plan, const double* q, double* r);   // capi.cu

// Solver workspace of a plan ahead of its first solve (warm start of panel plans, capi.cu): the vectors a solve
// feeds the matvec from (z) and reads it into (w) keep their addresses, so the launch graphs captured on them at
// construction are the ones the solve replays.
void gmres_reserve(fmmb_plan* plan, double** z_out, double** w_out) {
  const int64_t n = plan->tree.n * plan->charge_dim;
  if (!plan->gmres_ws) plan->gmres_ws = new GmresWorkspace();
  GmresWorkspace& ws = *plan->gmres_ws;
  ws.x.resize(n); ws.b.resize(n); ws.w.resize(n); ws.z.resize(n); ws.diag.resize(n);
  ws.scal.resize(256 + 4);
  ws.partial.resize(kDotBlocks);
  ws.counter.resize(1);
  if (ws.basis.n < (size_t)n * 24) ws.basis.resize((size_t)n * 24);
  ws.z.zero(plan->stream);
  // the solver's kernels are loaded with the module's first use of each: do that now
  cudaFuncAttributes fa;
  FMMB_CUDA(cudaFuncGetAttributes(&fa, dot_kernel));
  FMMB_CUDA(cudaFuncGetAttributes(&fa, axpy_dot_kernel));
  FMMB_CUDA(cudaFuncGetAttributes(&fa, mgs_sweep_kernel));
  if (ws.bar.n < 1) { ws.bar.resize(1); ws.bar.zero(plan->stream); ws.bar_count = 0; }
  ws.part2.resize(2 * kDotBlocks);
  FMMB_CUDA(cudaFuncGetAttributes(&fa, scale_into_kernel));
  FMMB_CUDA(cudaFuncGetAttributes(&fa, axpy_dev_kernel));
  FMMB_CUDA(cudaFuncGetAttributes(&fa, axpy_kernel));
  FMMB_CUDA(cudaFuncGetAttributes(&fa, precond_kernel));
  *z_out = ws.z.p;
  *w_out = ws.w.p;
}

// An inner near-field solve as preconditioner (Preconditioners::LocalInnerSolver / BlockDiagonal: examples/BEM/
// LocalPC.hpp:26-59, LocalPC_Stokes.hpp:27-62, BlockDiagonalPC.hpp:16-60): z = GMRES(pc plan, rhs v, x0 = 0, opts).
struct InnerPC {
  fmmb_plan* plan;
  fmmb_solver_options opts;
};

static GmresWorkspace& workspace(fmmb_plan* plan) {
  if (!plan->gmres_ws) plan->gmres_ws = new GmresWorkspace();
  return *plan->gmres_ws;
}

// The solver on device vectors x (initial guess in, solution out) and b of n doubles.  diag: device vector or nullptr.
// flexible: the preconditioned vectors Z[j] are kept and the solution is updated from them (FGMRES of
// GMRES_Stokes.hpp:319-431); otherwise the update applies the preconditioner to V[j] again (GMRES.hpp:234-239).
static void gmres_core(fmmb_plan* plan, double* x, const double* b, const double* diag, const fmmb_solver_options& o,
                       bool flexible, const InnerPC* pc, fmmb_gmres_info* info, int32_t* p_sched, double* res_hist,
                       int cap) {
  const int64_t n = plan->tree.n * plan->charge_dim;
  const int R = std::max(1, o.restart);
  cudaStream_t s = plan->stream;
  GmresWorkspace& ws = workspace(plan);
  DevBuf<double>&w = ws.w, &z = ws.z, &scal = ws.scal, &partial = ws.partial;
  DevBuf<unsigned int>& counter = ws.counter;
  w.resize(n); z.resize(n);
  scal.resize(R + 4);                     // [0..R]: Gram-Schmidt column + |w|^2, [R+1]: norms of b and of the residual
  partial.resize(kDotBlocks);
  counter.resize(1);
  counter.zero(s);
  if (ws.bar.n < 1) { ws.bar.resize(1); ws.bar.zero(s); ws.bar_count = 0; }
  ws.part2.resize(2 * kDotBlocks);
  if (ws.basis.n < (size_t)n * 24) ws.basis.resize((size_t)n * 24);
  auto basis = [&](int k) -> double* {
    if (ws.basis.n < (size_t)n * (k + 1)) ws.basis.grow(std::max((size_t)n * (k + 1), 2 * ws.basis.n), s);   // keeps the vectors
    return ws.basis.p + (size_t)n * k;
  };
  auto zbasis = [&](int k) -> double* {
    if (ws.zbasis.n < (size_t)n * (k + 1)) ws.zbasis.grow(std::max((size_t)n * (k + 1), std::max((size_t)n * 8, 2 * ws.zbasis.n)), s);
    return ws.zbasis.p + (size_t)n * k;
  };
  const int g = nblk(n, 256);
  auto dot_to = [&](const double* a, const double* c, double* out) {
    dot_kernel<<<kDotBlocks, 256, 0, s>>>(a, c, n, partial.p, counter.p, out);
  };
  auto fetch = [&](const double* dev, double* host, int count) {
    FMMB_CUDA(cudaMemcpyAsync(host, dev, count * sizeof(double), cudaMemcpyDeviceToHost, s));
    FMMB_CUDA(cudaStreamSynchronize(s));
  };
  // zout = M(v)
  auto apply_M = [&](const double* v, double* zout) {
    if (pc) {
      fmmb_plan* ip = pc->plan;
      GmresWorkspace& iw = workspace(ip);
      iw.x.resize(n); iw.b.resize(n);
      FMMB_CUDA(cudaStreamSynchronize(s));                       // v is complete; the inner solve runs on its plan's stream
      FMMB_CUDA(cudaMemcpyAsync(iw.b.p, v, n * sizeof(double), cudaMemcpyDeviceToDevice, ip->stream));
      FMMB_CUDA(cudaMemsetAsync(iw.x.p, 0, n * sizeof(double), ip->stream));
      gmres_core(ip, iw.x.p, iw.b.p, nullptr, pc->opts, false, nullptr, nullptr, nullptr, nullptr, 0);
      FMMB_CUDA(cudaStreamSynchronize(ip->stream));
      FMMB_CUDA(cudaMemcpyAsync(zout, iw.x.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    } else {
      precond_kernel<<<g, 256, 0, s>>>(v, diag, zout, n);
    }
  };
  double normb2 = 0;
  dot_to(b, b, scal.p + R + 1);
  fetch(scal.p + R + 1, &normb2, 1);
  const double normb = std::sqrt(normb2);
  if (!(normb > 0)) {
    // b = 0 (or not finite): the reference divides by |b| and iterates on NaN residuals; the solution of A x = 0
    // by GMRES from any start is x = 0, which is what a caller gets here, with zero iterations
    if (normb == 0) {
      FMMB_CUDA(cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), s));
      FMMB_CUDA(cudaStreamSynchronize(s));
      if (info) { info->iterations = 0; info->n_records = 0; info->final_residual = 0.0; info->final_p = plan->p; }
      return;
    }
    throw StatusError{FMMB_ERR_INVALID, "fmmb_gmres: the right-hand side is not finite"};
  }

  std::vector<std::vector<double>> H;
  std::vector<double> sv(R + 1), cs(R), sn(R), col(R + 2);
  double resid = 0;
  int i = -1, iter = 0, rec = 0;
  do {
    // w = A x - b at the order the kernel currently has; V[0] = -w / |w|
    run_matvec_for_solver(plan, x, w.p);
    axpy_kernel<<<g, 256, 0, s>>>(b, w.p, n, -1.0);
    dot_to(w.p, w.p, scal.p + R + 1);
    double beta2 = 0;
    fetch(scal.p + R + 1, &beta2, 1);
    const double beta = std::sqrt(beta2);
    FMMB_CUDA(cudaMemcpyAsync(basis(0), w.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    axpy_dev_kernel<<<g, 256, 0, s>>>(nullptr, basis(0), n, scal.p + R + 1, -1.0, 1);
    H.clear();
    std::fill(sv.begin(), sv.end(), 0.0);
    sv[0] = beta;
    i = -1;
    resid = sv[0] / normb;
    do {
      NvtxRange r_it("fmmb_gmres: iteration (predict_p, matvec, Gram-Schmidt, Givens)");
      ++i; ++iter;
      // GMRES.hpp:195: max(1u, predict_p);  GMRES_Stokes.hpp:229: max(p_min, predict_p - 1);  :375 (FGMRES): max(5u, predict_p)
      const unsigned pp = predict_p(o, std::fabs(resid));
      const int p = (int)std::max(std::max(1u, o.p_min), pp > o.p_offset ? pp - o.p_offset : 0u);
      if (p > FMMB_MAX_P) throw StatusError{FMMB_ERR_INVALID, "predicted expansion order exceeds FMMB_MAX_P"};
      plan->p = p;
      double* vnext = basis(i + 1);             // (may grow the basis: before any pointer into it is used)
      apply_M(basis(i), z.p);
      if (flexible) FMMB_CUDA(cudaMemcpyAsync(zbasis(i), z.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
      run_matvec_for_solver(plan, z.p, w.p);
      // modified Gram-Schmidt: coefficient k is produced and consumed on the device; launch k subtracts the
      // projection on V[k-1] and takes the product with V[k], the last one leaves |w|^2 behind the coefficients
      bool swept = false;
      if (ws.mgs_cooperative) {
        double* w_p = w.p; const double* basis_p = ws.basis.p; int nvec = i + 1; int64_t nn = n;
        double* part_p = ws.part2.p; unsigned int* bar_p = ws.bar.p; unsigned int base = ws.bar_count; double* scal_p = scal.p;
        void* args[] = {&w_p, &basis_p, &nvec, &nn, &vnext, &part_p, &bar_p, &base, &scal_p};
        const cudaError_t e = cudaLaunchCooperativeKernel((const void*)mgs_sweep_kernel, dim3(kDotBlocks), dim3(256), args, 0, s);
        if (e == cudaSuccess) {
          ws.bar_count += (unsigned)(nvec + 1) * kDotBlocks;
          swept = true;
        } else {
          cudaGetLastError();                   // no cooperative launches here: the chain of launches below
          ws.mgs_cooperative = false;
        }
      }
      if (!swept) {
        for (int k = 0; k <= i + 1; ++k)
          axpy_dot_kernel<<<kDotBlocks, 256, 0, s>>>(k ? basis(k - 1) : nullptr, k ? scal.p + k - 1 : nullptr, w.p,
                                                    k <= i ? basis(k) : nullptr, n, partial.p, counter.p, scal.p + k);
        scale_into_kernel<<<g, 256, 0, s>>>(w.p, vnext, n, scal.p + i + 1);
      }
      // the one synchronisation of the iteration: the new Hessenberg column
      FMMB_CUDA(cudaMemcpyAsync(col.data(), scal.p, (i + 2) * sizeof(double), cudaMemcpyDeviceToHost, s));
      FMMB_CUDA(cudaStreamSynchronize(s));
      for (int k = 0; k < i + 2; ++k)
        if (std::isnan(col[k])) throw StatusError{FMMB_ERR_CUDA, "fmmb_gmres: Gram-Schmidt coefficients are not numbers (grid barrier timed out, or the operator returned NaN)"};
      H.push_back(std::vector<double>(col.begin(), col.begin() + i + 2));
      H[i][i + 1] = std::sqrt(H[i][i + 1]);
      for (int k = 0; k < i; ++k) apply_rotation(H[i][k], H[i][k + 1], cs[k], sn[k]);
      make_rotation(H[i][i], H[i][i + 1], cs[i], sn[i]);
      apply_rotation(H[i][i], H[i][i + 1], cs[i], sn[i]);
      apply_rotation(sv[i], sv[i + 1], cs[i], sn[i]);
      resid = sv[i + 1] / normb;
      if (rec < cap) {
        if (p_sched) p_sched[rec] = p;
        if (res_hist) res_hist[rec] = std::fabs(resid);
      }
      ++rec;
      if (std::fabs(resid) < o.residual) break;
      if (o.verbose) printf("it: %03d, res: %.3e, fmm_req_p: %01d\n", iter, std::fabs(resid), p);
    } while (i + 1 < R && i + 1 <= o.max_iters && std::fabs(resid) > o.residual);
    // back substitution on the host, solution update on the device
    for (int j = i; j >= 0; --j) {
      sv[j] /= H[j][j];
      for (int k = j - 1; k >= 0; --k) sv[k] -= H[j][k] * sv[j];
    }
    for (int j = 0; j <= i; ++j) {
      if (flexible) {
        axpy_kernel<<<g, 256, 0, s>>>(zbasis(j), x, n, sv[j]);
      } else {
        apply_M(basis(j), z.p);
        axpy_kernel<<<g, 256, 0, s>>>(z.p, x, n, sv[j]);
      }
    }
    if (o.verbose && iter % 10 == 0) printf("it: %04d, residual: %.3e\n", iter, std::fabs(resid));
  } while (std::fabs(resid) > o.residual && iter < o.max_iters);
  if (o.verbose) printf("Final residual: %.4e, after %d iterations\n", std::fabs(resid), iter);
  FMMB_CUDA(cudaGetLastError());
  FMMB_CUDA(cudaStreamSynchronize(s));
  if (info) {
    info->iterations = iter;
    info->n_records = rec;
    info->final_residual = std::fabs(resid);
    info->final_p = plan->p;
  }
}

static void check_solver_plan(fmmb_plan* plan) {
  if (plan->charge_dim != plan->result_dim || !(plan->bem || plan->sbem))
    throw StatusError{FMMB_ERR_UNSUPPORTED, "fmmb_gmres: BEM plans (results have the shape of the charges)"};
  if (plan->tree.nranks > 1) throw StatusError{FMMB_ERR_UNSUPPORTED, "fmmb_gmres: single-GPU plans"};
}

// fmmb_gmres: host vectors in and out.  Vec<3> unknowns (StokesSphericalBEM) are solved on the flat array of 3 n
// doubles, like GMRES_Stokes.hpp:85-105.
void gmres_solve(fmmb_plan* plan, const double* b_host, double* x_host, const double* diag_host,
                 const fmmb_solver_options& o, fmmb_gmres_info* info, int32_t* p_sched, double* res_hist, int cap) {
  check_solver_plan(plan);
  const int64_t n = plan->tree.n * plan->charge_dim;
  cudaStream_t s = plan->stream;
  GmresWorkspace& ws = workspace(plan);
  ws.x.from_host(x_host, n, s);
  ws.b.from_host(b_host, n, s);
  if (diag_host) ws.diag.from_host(diag_host, n, s);
  gmres_core(plan, ws.x.p, ws.b.p, diag_host ? ws.diag.p : nullptr, o, false, nullptr, info, p_sched, res_hist, cap);
  FMMB_CUDA(cudaMemcpyAsync(x_host, ws.x.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  FMMB_CUDA(cudaStreamSynchronize(s));
}

// fmmb_fgmres: flexible GMRES; pc_plan == nullptr: identity preconditioner, else an inner solve on pc_plan (a
// near-field-only plan over the same panels) per outer iteration.
void fgmres_solve(fmmb_plan* plan, fmmb_plan* pc_plan, const fmmb_solver_options* pc_opts, const double* b_host,
                  double* x_host, const fmmb_solver_options& o, fmmb_gmres_info* info, int32_t* p_sched,
                  double* res_hist, int cap) {
  check_solver_plan(plan);
  const int64_t n = plan->tree.n * plan->charge_dim;
  InnerPC pc{pc_plan, fmmb_solver_options{}};
  if (pc_plan) {
    check_solver_plan(pc_plan);
    if (pc_plan == plan || pc_plan->tree.n * pc_plan->charge_dim != n || pc_plan->device != plan->device)
      throw StatusError{FMMB_ERR_INVALID, "fmmb_fgmres: the preconditioner plan must be another plan over the same panels on the same device"};
    if (!pc_opts) throw StatusError{FMMB_ERR_INVALID, "fmmb_fgmres: a preconditioner plan needs its solver options"};
    pc.opts = *pc_opts;
    pc.opts.verbose = 0;
  }
  cudaStream_t s = plan->stream;
  GmresWorkspace& ws = workspace(plan);
  ws.x.from_host(x_host, n, s);
  ws.b.from_host(b_host, n, s);
  gmres_core(plan, ws.x.p, ws.b.p, nullptr, o, true, pc_plan ? &pc : nullptr, info, p_sched, res_hist, cap);
  FMMB_CUDA(cudaMemcpyAsync(x_host, ws.x.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  FMMB_CUDA(cudaStreamSynchronize(s));
}

}  // namespace fmmb
