// csrc/comm.cu -- the one collective of a sharded matvec: all-gather of the per-rank result slices
// over NCCL (NVLink / NVSwitch on a single node).  Slices have different lengths (ranges are balanced
// by work, not by count), so the all-gather is a group of broadcasts, one per owner.
#include "common.cuh"
#include <nccl.h>
#include <cstring>
#include <algorithm>

namespace fmmb {

#define FMMB_NCCL(call)                                                                         \
  do {                                                                                          \
    ncclResult_t r_ = (call);                                                                   \
    if (r_ != ncclSuccess)                                                                      \
      throw StatusError{FMMB_ERR_CUDA, std::string("NCCL error in " #call ": ") + ncclGetErrorString(r_)}; \
  } while (0)

static_assert(sizeof(ncclUniqueId) == 128, "fmmb.h promises a 128-byte id");

void comm_unique_id(unsigned char* id) {
  ncclUniqueId u;
  FMMB_NCCL(ncclGetUniqueId(&u));
  std::memcpy(id, &u, sizeof u);
}

void comm_init(fmmb_plan* plan, const unsigned char* id) {
  if (plan->tree.nranks <= 1) throw StatusError{FMMB_ERR_INVALID, "plan was not created with nranks > 1"};
  if (plan->comm) return;
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof u);
  ncclComm_t c;
  FMMB_NCCL(ncclCommInitRank(&c, plan->tree.nranks, u, plan->tree.rank));
  plan->comm = c;
}

void comm_destroy(fmmb_plan* plan) {
  if (plan->comm) ncclCommDestroy((ncclComm_t)plan->comm);
  plan->comm = nullptr;
}

// Result slices differ in length (ranges are balanced by work): one ncclAllGather of chunks padded to the
// longest slice into a staging buffer, then every slice is copied to its place in tree order.
__global__ void place_slices(const double4* __restrict__ stage, const long long* __restrict__ cuts, int nranks,
                             long long chunk, double4* __restrict__ tree) {
  const int q = blockIdx.y;
  const long long b0 = cuts[q], len = cuts[q + 1] - b0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x)
    tree[b0 + i] = stage[(size_t)q * chunk + i];
}

static void ensure_cuts(fmmb_plan* plan, cudaStream_t s) {
  Tree& T = plan->tree;
  if (!plan->cuts_ready) {
    std::vector<long long> h(T.body_cuts.begin(), T.body_cuts.end());
    plan->cuts_dev.resize(h.size());
    FMMB_CUDA(cudaMemcpyAsync(plan->cuts_dev.p, h.data(), h.size() * sizeof(long long), cudaMemcpyHostToDevice, s));
    FMMB_CUDA(cudaStreamSynchronize(s));
    plan->cuts_ready = true;
  }
}

// Sharded call: every rank contributes the charges of its own bodies (tree order); slices are padded to the
// longest one and gathered with one ncclAllGather (8 bytes per body instead of 32 for the results).
void allgather_charges(fmmb_plan* plan, const double* d_own, cudaStream_t s) {
  Tree& T = plan->tree;
  ncclComm_t c = (ncclComm_t)plan->comm;
  long long chunk = 0;
  for (int q = 0; q < T.nranks; ++q) chunk = std::max<long long>(chunk, T.body_cuts[q + 1] - T.body_cuts[q]);
  plan->chg_chunk = chunk;
  plan->chg_stage.resize((size_t)chunk * T.nranks);
  plan->chg_send.resize((size_t)chunk);
  ensure_cuts(plan, s);
  // the caller's slice is exactly own_n long: stage it so the padded send never reads past its end
  const long long own = T.own_b1 - T.own_b0;
  if (own) FMMB_CUDA(cudaMemcpyAsync(plan->chg_send.p, d_own, own * sizeof(double), cudaMemcpyDeviceToDevice, s));
  FMMB_NCCL(ncclAllGather(plan->chg_send.p, plan->chg_stage.p, (size_t)chunk, ncclDouble, c, s));
}

void allgather_results(fmmb_plan* plan, cudaStream_t s) {
  Tree& T = plan->tree;
  ncclComm_t c = (ncclComm_t)plan->comm;
  long long chunk = 0;
  for (int q = 0; q < T.nranks; ++q) chunk = std::max<long long>(chunk, T.body_cuts[q + 1] - T.body_cuts[q]);
  plan->res_stage.resize((size_t)chunk * T.nranks);
  ensure_cuts(plan, s);
  // send buffer = my slice inside res_tree (reads past the slice end stay inside res_tree or the pad)
  const double* send = reinterpret_cast<const double*>(plan->res_tree.p + T.own_b0);
  FMMB_NCCL(ncclAllGather(send, plan->res_stage.p, (size_t)chunk * 4, ncclDouble, c, s));
  dim3 grid(64, T.nranks);
  place_slices<<<grid, 256, 0, s>>>(plan->res_stage.p, plan->cuts_dev.p, T.nranks, chunk, plan->res_tree.p);
  FMMB_CUDA(cudaGetLastError());
}

// ---- generic result epilogue (BEM: 1, Stokes: 3, Yukawa: 4 doubles per body) ----------------------------------
namespace {
__global__ void gen_combine_slice(const double* __restrict__ near, const double* __restrict__ far, int64_t i0, int64_t i1,
                                  int rd, double* __restrict__ out) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < (i1 - i0) * rd) out[t] = near[i0 * rd + t] + far[i0 * rd + t];
}
__global__ void iota_kernel(unsigned* __restrict__ a, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = (unsigned)i;
}
__global__ void place_charge_slices(const double* __restrict__ stage, const long long* __restrict__ cuts, long long chunk,
                                    int cd, double* __restrict__ qtree) {
  const int q = blockIdx.y;
  const long long b0 = cuts[q] * cd, len = (cuts[q + 1] - cuts[q]) * cd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x)
    qtree[b0 + i] = stage[(size_t)q * chunk * cd + i];
}
__global__ void gen_combine_scatter(const double* __restrict__ near, const double* __restrict__ far,
                                    const unsigned* __restrict__ perm, int64_t i0, int64_t i1, int rd,
                                    double* __restrict__ out) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (i1 - i0) * rd) return;
  const int64_t i = i0 + t / rd;
  const int k = (int)(t % rd);
  out[(size_t)perm[i] * rd + k] = near[(size_t)i * rd + k] + far[(size_t)i * rd + k];
}
__global__ void gen_combine(const double* __restrict__ near, const double* __restrict__ far, int64_t i0, int64_t i1,
                            int rd, double* __restrict__ tree) {
  int64_t t = i0 * rd + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < i1 * rd) tree[t] = near[t] + far[t];
}
__global__ void gen_place(const double* __restrict__ stage, const long long* __restrict__ cuts, long long chunk,
                          int rd, double* __restrict__ tree) {
  const int q = blockIdx.y;
  const long long b0 = cuts[q] * rd, len = (cuts[q + 1] - cuts[q]) * rd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x)
    tree[b0 + i] = stage[(size_t)q * chunk * rd + i];
}
__global__ void gen_scatter(const double* __restrict__ tree, const unsigned* __restrict__ perm, int64_t n, int rd,
                            double* __restrict__ out) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n * rd) return;
  const int64_t i = t / rd;
  out[(size_t)perm[i] * rd + t % rd] = tree[t];
}
}  // namespace

// near + far (tree order, rd doubles per body, valid on the owned range) -> results in the caller's order.
// Single GPU / no communicator: the owned range is scattered.  With a communicator the owned slices are
// all-gathered (padded NCCL all-gather) so that every rank ends up with the full result vector.
void finish_results(fmmb_plan* plan, const double* near, const double* far, int rd, double* d_results, cudaStream_t s) {
  Tree& T = plan->tree;
  const int64_t n = T.n;
  auto nb = [](int64_t c, int t) { return (int)((c + t - 1) / t); };
  if (plan->call_sharded) {
    // results stay sharded by target: this rank's slice in tree order, no collective, no permutation
    if (T.own_b1 > T.own_b0)
      gen_combine_slice<<<nb((T.own_b1 - T.own_b0) * rd, 256), 256, 0, s>>>(near, far, T.own_b0, T.own_b1, rd, d_results);
    ++plan->launches;
    FMMB_CUDA(cudaGetLastError());
    return;
  }
  if (!(T.nranks > 1 && plan->comm)) {
    if (T.own_b1 > T.own_b0)
      gen_combine_scatter<<<nb((T.own_b1 - T.own_b0) * rd, 256), 256, 0, s>>>(near, far, T.perm.p, T.own_b0, T.own_b1,
                                                                             rd, d_results);
    ++plan->launches;
    FMMB_CUDA(cudaGetLastError());
    return;
  }
  ncclComm_t c = (ncclComm_t)plan->comm;
  long long chunk = 0;
  for (int q = 0; q < T.nranks; ++q) chunk = std::max<long long>(chunk, T.body_cuts[q + 1] - T.body_cuts[q]);
  plan->gen_tree.resize(((size_t)n + (size_t)chunk) * rd);     // + pad: the padded send reads a full chunk
  plan->gen_stage.resize((size_t)chunk * T.nranks * rd);
  ensure_cuts(plan, s);
  if (T.own_b1 > T.own_b0)
    gen_combine<<<nb((T.own_b1 - T.own_b0) * rd, 256), 256, 0, s>>>(near, far, T.own_b0, T.own_b1, rd, plan->gen_tree.p);
  FMMB_NCCL(ncclAllGather(plan->gen_tree.p + (size_t)T.own_b0 * rd, plan->gen_stage.p, (size_t)chunk * rd, ncclDouble, c, s));
  dim3 grid(64, T.nranks);
  gen_place<<<grid, 256, 0, s>>>(plan->gen_stage.p, plan->cuts_dev.p, chunk, rd, plan->gen_tree.p);
  gen_scatter<<<nb(n * rd, 256), 256, 0, s>>>(plan->gen_tree.p, T.perm.p, n, rd, d_results);
  plan->launches += 3;
  FMMB_CUDA(cudaGetLastError());
}

// Sharded call of the kernel classes that gather their charges through a permutation (BEM, Stokes, Yukawa): the
// per-rank charge slices (tree order, cd doubles per body) become one tree-ordered vector on every rank -- a padded
// ncclAllGather -- and the class's own gather kernel then runs with the identity permutation (exec_perm below).
const double* sharded_assemble_charges(fmmb_plan* plan, const double* d_own, cudaStream_t s) {
  Tree& T = plan->tree;
  const int cd = plan->charge_dim;
  auto nb = [](int64_t c, int t) { return (int)((c + t - 1) / t); };
  if (T.iota.n != (size_t)T.n) {
    if (plan->capturing) throw StatusError{FMMB_ERR_INVALID, "first sharded call inside a graph capture"};
    T.iota.resize(T.n);
    iota_kernel<<<nb(T.n, 256), 256, 0, s>>>(T.iota.p, T.n);
  }
  if (T.nranks == 1) return d_own;                          // the slice is the whole tree-ordered vector
  if (!plan->comm) throw StatusError{FMMB_ERR_INVALID, "call fmmb_plan_comm_init first"};
  ncclComm_t c = (ncclComm_t)plan->comm;
  long long chunk = 0;
  for (int q = 0; q < T.nranks; ++q) chunk = std::max<long long>(chunk, T.body_cuts[q + 1] - T.body_cuts[q]);
  plan->chg_stage.resize((size_t)chunk * T.nranks * cd);
  plan->chg_send.resize((size_t)chunk * cd);
  plan->q_tree.resize((size_t)T.n * cd);
  ensure_cuts(plan, s);
  const long long own = T.own_b1 - T.own_b0;
  if (own) FMMB_CUDA(cudaMemcpyAsync(plan->chg_send.p, d_own, own * cd * sizeof(double), cudaMemcpyDeviceToDevice, s));
  FMMB_NCCL(ncclAllGather(plan->chg_send.p, plan->chg_stage.p, (size_t)chunk * cd, ncclDouble, c, s));
  dim3 grid(64, T.nranks);
  place_charge_slices<<<grid, 256, 0, s>>>(plan->chg_stage.p, plan->cuts_dev.p, chunk, cd, plan->q_tree.p);
  FMMB_CUDA(cudaGetLastError());
  plan->launches += 2;
  return plan->q_tree.p;
}

namespace {
__global__ void pack_boxes(const int* __restrict__ list, int count, int xs, const double* __restrict__ M,
                           double* __restrict__ out) {
  int i = blockIdx.x;
  if (i >= count) return;
  const double* src = M + (size_t)list[i] * xs;
  for (int k = threadIdx.x; k < xs; k += blockDim.x) out[(size_t)i * xs + k] = src[k];
}
__global__ void unpack_boxes(const int* __restrict__ list, const int* __restrict__ off, int nranks, int me, int chunk,
                             int xs, const double* __restrict__ in, double* __restrict__ M) {
  // blockIdx.y = owner rank, blockIdx.x = box slot inside the owner's chunk
  int q = blockIdx.y, i = blockIdx.x;
  if (q == me || i >= off[q + 1] - off[q]) return;
  const double* src = in + ((size_t)q * chunk + i) * xs;
  double* dst = M + (size_t)list[off[q] + i] * xs;
  for (int k = threadIdx.x; k < xs; k += blockDim.x) dst[k] = src[k];
}
}  // namespace

// Owned upward pass: every rank holds the multipoles of the boxes inside its range; one padded
// ncclAllGather makes them all visible everywhere.
void exchange_multipoles(fmmb_plan* plan, cudaStream_t s) {
  Tree& T = plan->tree;
  ncclComm_t c = (ncclComm_t)plan->comm;
  const int P = plan->p, xs = (P * P + 1) & ~1;
  const size_t chunk = (size_t)T.xchg_max * xs;
  plan->xchg_off_dev.resize(T.nranks + 1);
  if (!plan->xchg_off_ready) {
    FMMB_CUDA(cudaMemcpyAsync(plan->xchg_off_dev.p, T.xchg_off.data(), (T.nranks + 1) * sizeof(int),
                              cudaMemcpyHostToDevice, s));
    plan->xchg_off_ready = true;
  }
  T.xchg_send.resize(chunk);
  T.xchg_recv.resize(chunk * T.nranks);
  const int mine = T.xchg_off[T.rank + 1] - T.xchg_off[T.rank];
  if (mine) pack_boxes<<<mine, 64, 0, s>>>(T.xchg_list.p + T.xchg_off[T.rank], mine, xs, plan->M.p, T.xchg_send.p);
  FMMB_NCCL(ncclAllGather(T.xchg_send.p, T.xchg_recv.p, chunk, ncclDouble, c, s));
  dim3 grid(T.xchg_max, T.nranks);
  unpack_boxes<<<grid, 64, 0, s>>>(T.xchg_list.p, plan->xchg_off_dev.p, T.nranks, T.rank, T.xchg_max, xs,
                                  T.xchg_recv.p, plan->M.p);
  FMMB_CUDA(cudaGetLastError());
  plan->launches += 2;
}


// ---- multipole exchange through peer memory (NVLink P2P stores, no NCCL, no staging) -----------------------------
// Every rank exports the allocation that holds its multipole array (cudaIpcMemHandle) and opens the peers'.  After
// its owned upward pass a rank PUSHES the multipoles of the boxes inside its range straight into every peer's
// array, 512-byte rows over NVLink, from the kernel that reads them (one launch instead of pack / ncclAllGather /
// unpack).  Synchronisation is two monotonic flag vectors per rank that the peers write with release stores at
// system scope (they live in the tail of the exported allocation):
//   pushed[q]  = e   rank q's multipoles of matvec e have landed here     (waited on before M2L reads them)
//   rdone[q]   = e   rank q has finished reading ITS array in matvec e     (waited on before pushing matvec e + 1,
//                                                                           so nobody's array changes under a reader)
//   qpushed[q] = e   rank q's charge slice of matvec e has landed here     (sharded call; waited on before it is used)
// The same allocation carries a tree-ordered charge vector that the peers fill in the sharded call, which replaces
// the NCCL all-gather of the charge slices.  The matvec counter lives in device memory (advanced by the last
// kernel of every matvec), so a captured CUDA graph replays correctly.
namespace {
constexpr int kPeerMaxRanks = 64;
constexpr int kPeerTail = 4 * kPeerMaxRanks;       // doubles behind the multipoles: pushed[64], rdone[64], qpushed[64]

struct PeerBlob {                                   // what travels between the ranks (128 bytes)
  cudaIpcMemHandle_t mem;                           // 64 bytes
  long long doubles;                                // size of the multipole part, for a consistency check
  int rank, device;
  char pad[128 - sizeof(cudaIpcMemHandle_t) - sizeof(long long) - 2 * sizeof(int)];
};
static_assert(sizeof(PeerBlob) == 128, "fmmb.h promises a 128-byte blob");

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Bounded wait on a flag a peer writes: a rank that died or fell out of step must not wedge this GPU inside a kernel.
// After kPeerTimeoutNs the wait gives up, raises the plan's timeout flag (peer_state[3], reported by the next
// fmmb_plan_sync / host-buffer call as FMMB_ERR_CUDA) and lets the matvec run to its end on whatever data is there.
constexpr unsigned long long kPeerTimeoutNs = 20ull * 1000 * 1000 * 1000;
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void wait_flag_at_least(const unsigned long long* flag, unsigned long long v,
                                                   unsigned long long* timeout_flag) {
  if (ld_acquire_sys(flag) >= v) return;
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(flag) < v) {
    __nanosleep(200);
    if (global_ns() - t0 > kPeerTimeoutNs) { atomicExch(timeout_flag, 1ull); return; }
  }
}

__global__ void __launch_bounds__(64)
peer_push_kernel(const int* __restrict__ list, int count, int xs, const double* __restrict__ M,
                 double* const* __restrict__ peerM, unsigned long long* const* __restrict__ peer_flags, int nranks, int me,
                 const unsigned long long* __restrict__ local_flags, unsigned long long* __restrict__ epoch,
                 unsigned int* __restrict__ counter) {
  __shared__ int s_last;
  const unsigned long long e = *epoch;              // matvecs completed so far; this one is e + 1
  // nobody may still be reading the previous multipoles out of the arrays this block is about to write
  if ((int)threadIdx.x < nranks) wait_flag_at_least(local_flags + kPeerMaxRanks + threadIdx.x, e, epoch + 3);
  __syncthreads();
  const int i = blockIdx.x;
  if (i < count) {
    const size_t o = (size_t)list[i] * xs;
    for (int q = 0; q < nranks; ++q) {
      if (q == me) continue;
      double* dst = peerM[q] + o;
      for (int k = threadIdx.x; k < xs; k += blockDim.x) dst[k] = M[o + k];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {                                     // every block's rows are out: tell the peers (and myself)
    __threadfence_system();
    if ((int)threadIdx.x < nranks) st_release_sys(peer_flags[threadIdx.x] + me, e + 1);
    if (threadIdx.x == 0) *counter = 0;
  }
}
__global__ void peer_wait_kernel(const unsigned long long* __restrict__ local_flags,
                                 unsigned long long* __restrict__ epoch, int nranks) {
  const unsigned long long e = *epoch + 1;
  if ((int)threadIdx.x < nranks) wait_flag_at_least(local_flags + threadIdx.x, e, epoch + 3);
}
// last kernel of a matvec: this rank no longer reads its multipole array; the matvec counter advances
__global__ void peer_read_done_kernel(unsigned long long* const* __restrict__ peer_flags, int nranks, int me,
                                      unsigned long long* __restrict__ epoch) {
  const unsigned long long e = *epoch + 1;
  if ((int)threadIdx.x < nranks) {
    __threadfence_system();
    st_release_sys(peer_flags[threadIdx.x] + kPeerMaxRanks + me, e);
  }
  __syncthreads();
  if (threadIdx.x == 0) *epoch = e;
}
// sharded call: my charge slice (tree order) goes straight into every rank's tree-ordered charge vector
__global__ void __launch_bounds__(256)
peer_push_charges_kernel(const double* __restrict__ own, long long b0, long long len, double* const* __restrict__ peerM,
                         size_t q_off, unsigned long long* const* __restrict__ peer_flags, int nranks, int me,
                         const unsigned long long* __restrict__ epoch, unsigned int* __restrict__ counter) {
  __shared__ int s_last;
  const unsigned long long e = *epoch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x) {
    const double v = own[i];
    for (int q = 0; q < nranks; ++q) peerM[q][q_off + b0 + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if ((int)threadIdx.x < nranks) st_release_sys(peer_flags[threadIdx.x] + 2 * kPeerMaxRanks + me, e + 1);
    if (threadIdx.x == 0) *counter = 0;
  }
}
// ... and, once every rank's slice has landed, into the charge slot of the bodies
__global__ void __launch_bounds__(256)
peer_place_charges_kernel(const unsigned long long* __restrict__ local_flags, unsigned long long* __restrict__ epoch,
                          int nranks, const double* __restrict__ qtree, long long n, double4* __restrict__ body) {
  const unsigned long long e = *epoch + 1;
  if ((int)threadIdx.x < nranks) wait_flag_at_least(local_flags + 2 * kPeerMaxRanks + threadIdx.x, e, epoch + 3);
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    body[i].w = qtree[i];
}
}  // namespace

void peer_export(fmmb_plan* plan, unsigned char* blob) {
  Tree& T = plan->tree;
  if (T.nranks <= 1) throw StatusError{FMMB_ERR_INVALID, "plan was not created with nranks > 1"};
  if (T.nranks > kPeerMaxRanks) throw StatusError{FMMB_ERR_UNSUPPORTED, "peer exchange is built for up to 64 ranks"};
  if (plan->kind != FMMB_LAPLACE_SPHERICAL) throw StatusError{FMMB_ERR_UNSUPPORTED, "peer exchange: LaplaceSpherical plans"};
  cudaStream_t s = plan->stream;
  // one allocation for good: multipoles at the largest batched order (P = 8: 64 doubles per box) + the flag tail
  const size_t md = (size_t)(T.nboxes + 1) * 64;
  if (!plan->peer_alloc) {
    FMMB_CUDA(cudaStreamSynchronize(s));
    plan->M.release();
    plan->M.resize(md + kPeerTail + (size_t)T.n);     // multipoles | flag vectors | tree-ordered charges
    plan->M.zero(s);
    plan->M.n = 0;
    plan->p_alloc = 0;
    plan->peer_alloc = true;
    plan->peer_state.resize(4);                       // [0] matvec counter, [1], [2] block counters of the push kernels,
                                                      // [3] timeout flag of the bounded flag waits
    plan->peer_state.zero(s);
    FMMB_CUDA(cudaStreamSynchronize(s));
  }
  PeerBlob b;
  std::memset(&b, 0, sizeof b);
  FMMB_CUDA(cudaIpcGetMemHandle(&b.mem, plan->M.p));
  b.doubles = (long long)md;
  b.rank = T.rank;
  b.device = plan->device;
  std::memcpy(blob, &b, sizeof b);
}

void peer_init(fmmb_plan* plan, const unsigned char* blobs) {
  Tree& T = plan->tree;
  if (!plan->peer_alloc) throw StatusError{FMMB_ERR_INVALID, "call fmmb_plan_peer_export first"};
  if (plan->peer_ready) return;
  const size_t md = (size_t)(T.nboxes + 1) * 64;
  std::vector<double*> pm(T.nranks);
  std::vector<unsigned long long*> pf(T.nranks);
  for (int q = 0; q < T.nranks; ++q) {
    PeerBlob b;
    std::memcpy(&b, blobs + 128 * (size_t)q, sizeof b);
    if (b.rank != q || b.doubles != (long long)md)
      throw StatusError{FMMB_ERR_INVALID, "peer blobs must be ordered by rank and come from plans on the same tree"};
    void* p = plan->M.p;
    if (q != T.rank) {
      FMMB_CUDA(cudaIpcOpenMemHandle(&p, b.mem, cudaIpcMemLazyEnablePeerAccess));
      plan->peer_opened.push_back(p);
    }
    pm[q] = (double*)p;
    pf[q] = (unsigned long long*)((double*)p + md);
  }
  plan->peer_M.from_host(pm.data(), pm.size(), plan->stream);
  plan->peer_flags.from_host(pf.data(), pf.size(), plan->stream);
  FMMB_CUDA(cudaStreamSynchronize(plan->stream));
  plan->peer_ready = true;
}

void peer_close(fmmb_plan* plan) {
  if (plan->peer_flag_host) { cudaFreeHost(plan->peer_flag_host); plan->peer_flag_host = nullptr; }
  for (void* p : plan->peer_opened) cudaIpcCloseMemHandle(p);
  plan->peer_opened.clear();
  plan->peer_ready = false;
}

void exchange_multipoles_peer(fmmb_plan* plan, cudaStream_t s) {
  Tree& T = plan->tree;
  const int P = plan->p, xs = (P * P + 1) & ~1;
  const int mine = T.xchg_off[T.rank + 1] - T.xchg_off[T.rank];
  unsigned long long* local_flags = (unsigned long long*)(plan->M.p + (size_t)(T.nboxes + 1) * 64);
  unsigned long long* st = plan->peer_state.p;
  peer_push_kernel<<<std::max(mine, 1), 64, 0, s>>>(T.xchg_list.p + T.xchg_off[T.rank], mine, xs, plan->M.p, plan->peer_M.p,
                                                   plan->peer_flags.p, T.nranks, T.rank, local_flags, st,
                                                   (unsigned int*)(st + 1));
  peer_wait_kernel<<<1, 64, 0, s>>>(local_flags, st, T.nranks);
  FMMB_CUDA(cudaGetLastError());
  plan->launches += 2;
}

// sharded call: d_own = the charges of this rank's bodies (tree order) -> body[].w on every rank
void peer_exchange_charges(fmmb_plan* plan, const double* d_own, cudaStream_t s) {
  Tree& T = plan->tree;
  const size_t md = (size_t)(T.nboxes + 1) * 64;
  unsigned long long* local_flags = (unsigned long long*)(plan->M.p + md);
  unsigned long long* st = plan->peer_state.p;
  const long long own = T.own_b1 - T.own_b0;
  const int blocks = (int)std::max<long long>(1, std::min<long long>(296, (own + 255) / 256));
  peer_push_charges_kernel<<<blocks, 256, 0, s>>>(d_own, T.own_b0, own, plan->peer_M.p, md + kPeerTail, plan->peer_flags.p,
                                                 T.nranks, T.rank, st, (unsigned int*)(st + 2));
  peer_place_charges_kernel<<<296, 256, 0, s>>>(local_flags, st, T.nranks, plan->M.p + md + kPeerTail, T.n, T.body.p);
  FMMB_CUDA(cudaGetLastError());
  plan->launches += 2;
}

// Did a bounded flag wait of the peer exchange give up?  The flag travels to a pinned host word behind the matvec
// (peer_flag_fetch, asynchronous, 8 bytes) and is looked at after the caller's synchronisation (peer_check_timeout):
// no extra blocking copy on the host-buffer call path.
void peer_flag_fetch(fmmb_plan* plan, cudaStream_t s) {
  if (!plan->peer_ready) return;
  if (!plan->peer_flag_host) {
    FMMB_CUDA(cudaHostAlloc((void**)&plan->peer_flag_host, sizeof(unsigned long long), cudaHostAllocDefault));
    *plan->peer_flag_host = 0;
  }
  FMMB_CUDA(cudaMemcpyAsync(plan->peer_flag_host, plan->peer_state.p + 3, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
}
void peer_check_timeout(fmmb_plan* plan) {
  if (!plan->peer_ready || !plan->peer_flag_host || !*plan->peer_flag_host) return;
  *plan->peer_flag_host = 0;
  FMMB_CUDA(cudaMemset(plan->peer_state.p + 3, 0, sizeof(unsigned long long)));
  throw StatusError{FMMB_ERR_CUDA, "peer-memory exchange timed out: a rank did not arrive within 20 s "
                                   "(all ranks must run the same sequence of matvecs); the last results are invalid"};
}

void peer_read_done(fmmb_plan* plan, cudaStream_t s) {
  Tree& T = plan->tree;
  peer_read_done_kernel<<<1, 64, 0, s>>>(plan->peer_flags.p, T.nranks, T.rank, plan->peer_state.p);
  FMMB_CUDA(cudaGetLastError());
  ++plan->launches;
}

}  // namespace fmmb
