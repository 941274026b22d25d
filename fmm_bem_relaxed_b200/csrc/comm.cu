// csrc/comm.cu -- the one collective of a sharded matvec: all-gather of the per-rank result slices
// over NCCL (NVLink / NVSwitch on a single node).  Slices have different lengths (ranges are balanced
// by work, not by count), so the all-gather is a group of broadcasts, one per owner.
#include "common.cuh"
#include <nccl.h>
#include <cstring>

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

void allgather_results(fmmb_plan* plan, cudaStream_t s) {
  Tree& T = plan->tree;
  ncclComm_t c = (ncclComm_t)plan->comm;
  double* base = reinterpret_cast<double*>(plan->res_tree.p);
  FMMB_NCCL(ncclGroupStart());
  for (int q = 0; q < T.nranks; ++q) {
    int64_t b0 = T.body_cuts[q], b1 = T.body_cuts[q + 1];
    if (b1 > b0) FMMB_NCCL(ncclBroadcast(base + 4 * b0, base + 4 * b0, (size_t)(4 * (b1 - b0)), ncclDouble, q, c, s));
  }
  FMMB_NCCL(ncclGroupEnd());
}

}  // namespace fmmb
