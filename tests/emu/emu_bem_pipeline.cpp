// tests/emu/emu_bem_pipeline.cpp -- TEST INFRASTRUCTURE ONLY (built and run by tests/test_cuda_emulation.py, CPU).
// The HOST functions bem_setup + bem_execute of fmm_bem_relaxed_b200/csrc/bem.cu as written (LaplaceSphericalBEM plan),
// under the emulation of tests/emu/cuda_emu.hpp: runtime calls as macros on host memory, every launch with its own
// configuration and guard bytes behind the dynamic shared segment.  Used for what has not run on hardware in that file:
// the Gauss rules above 4 points and the treecode branch (bem_m2p_kernel).  Translations: the per-pair kernels of
// csrc/laplace.cu when the file carries the far-field structures, nothing to do otherwise.
//   emu_bem_pipeline <file>      file layout: see tests/test_cuda_emulation.py; writes <file>.out (n doubles)
#include "cuda_emu.hpp"
#include "../../fmm_bem_relaxed_b200/csrc/common.cuh"
#include "../../fmm_bem_relaxed_b200/csrc/laplace_ops.cuh"
#include "../../fmm_bem_relaxed_b200/hostcxx/bem_math.hpp"

namespace fmmb {
#include "bem_whole.inc"
namespace emu_trans {          // the per-pair translation kernels of csrc/laplace.cu (the path behind P > 8 and m2l_mode 1)
using namespace ops;
#include "lap_trans.inc"
}
static int g_translation_calls = 0;
// csrc/laplace.cu::laplace_translations, per-pair path, launch configurations of :943-981 (see emu_stokes_bem.cpp)
void laplace_translations(fmmb_plan* plan, cudaStream_t) {
  ++g_translation_calls;
  Tree& T = plan->tree;
  if (T.n_lr == 0) return;
  const int P = plan->p, nc = P * (P + 1) / 2, pp = P * P, nb = T.nboxes;
  const size_t sh_mm = (size_t)(pp + nc) * sizeof(double2);
  for (int l = T.nlevels - 2; l >= 0; --l) {
    const int lo = T.level_off[l], hi = T.level_off[l + 1];
    emu::launch_cfg(hi - lo, 64, sh_mm, [&] {
      emu_trans::m2m_kernel(lo, hi, nullptr, T.key.p, T.cbegin.p, T.cend.p, T.center.p, P, plan->M.p);
    });
  }
  if (plan->opts.evaluator == FMMB_EVAL_TREECODE) return;
  static std::map<int, std::vector<double>> coeff;
  std::vector<double>& C = coeff[P];
  if (C.empty()) {
    C.resize((size_t)nc * pp);
    emu::launch_cfg((int)((C.size() + 255) / 256), 256, 0, [&] { emu_trans::m2l_coeff_kernel(P, C.data()); });
  }
  int threads = 128;
  while (threads < nc) threads += 32;
  size_t sh = (size_t)(5 * pp) * sizeof(double2);
  const size_t red = (size_t)(threads / nc) * nc * sizeof(double2);
  if (red > sh) sh = red;
  emu::launch_cfg(nb, threads, sh, [&] {
    emu_trans::m2l_pair_kernel(nb, nullptr, T.m2l_off.p, T.m2l_src.p, T.center.p, P, C.data(), plan->M.p, plan->L.p, 0);
  });
  for (int l = 1; l < T.nlevels; ++l) {
    const int lo = T.level_off[l], hi = T.level_off[l + 1];
    emu::launch_cfg(hi - lo, 64, sh_mm, [&] {
      emu_trans::l2l_kernel(lo, hi, T.parent.p, T.has_local.p, T.center.p, P, plan->L.p);
    });
  }
}
void laplace_prepare_expansions(fmmb_plan* plan) {          // csrc/laplace.cu: sizes plan->M / plan->L for the order
  const int xs = ops::xstride(plan->p);
  plan->M.resize((size_t)plan->tree.nboxes * xs);
  plan->L.resize((size_t)plan->tree.nboxes * xs);
  plan->M.zero(nullptr); plan->L.zero(nullptr);
}
void finish_results(fmmb_plan* plan, const double* near, const double* far, int rd, double* d_results, cudaStream_t) {
  Tree& T = plan->tree;                                   // csrc/comm.cu: gen_combine_scatter on one rank
  for (int64_t i = T.own_b0; i < T.own_b1; ++i)
    for (int c = 0; c < rd; ++c) d_results[(size_t)T.perm.p[i] * rd + c] = near[i * rd + c] + far[i * rd + c];
}
}  // namespace fmmb
using namespace fmmb;

template <class T> static const T* take(const char*& p, size_t n) { const T* r = (const T*)p; p += n * sizeof(T); return r; }

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<char> buf(sz);
  if (fread(buf.data(), 1, sz, f) != (size_t)sz) return 2;
  fclose(f);
  const char* p = buf.data();
  const long long* hd = take<long long>(p, 4);
  const long n = hd[0], nb = hd[1], ni = hd[2], ne = hd[3];
  const int* ip = take<int>(p, 4);
  const int K = ip[0], P = ip[1], treecode = ip[2];
  const double* verts = take<double>(p, 9 * n);
  const int* bc = take<int>(p, n);
  const unsigned* perm = take<unsigned>(p, n);
  const unsigned* bb = take<unsigned>(p, nb);
  const unsigned* be = take<unsigned>(p, nb);
  const int* off = take<int>(p, nb + 1);
  const int* src = take<int>(p, ne);
  const int4* items = (const int4*)take<int>(p, 4 * ni);
  const double* q = take<double>(p, n);
  const double* geom = take<double>(p, 4 * nb);
  const unsigned* parent = take<unsigned>(p, nb);
  const int* leaf = take<int>(p, nb);
  // optional far-field section: int64 n_lr, int32 nlevels, pad; key[nb], cbegin[nb], cend[nb] u32,
  // level_off[nlevels + 1], m2l_off[nb + 1], m2l_src[n_lr], has_local[nb] i32
  long n_lr = 0;
  int nlevels = 0;
  const unsigned *key = nullptr, *cbegin = nullptr, *cend = nullptr;
  const int *level_off = nullptr, *m2l_off = nullptr, *m2l_src = nullptr, *has_local_i = nullptr;
  if (p < buf.data() + buf.size()) {
    n_lr = (long)*take<long long>(p, 1);
    nlevels = take<int>(p, 2)[0];
    key = take<unsigned>(p, nb); cbegin = take<unsigned>(p, nb); cend = take<unsigned>(p, nb);
    level_off = take<int>(p, nlevels + 1);
    m2l_off = take<int>(p, nb + 1); m2l_src = take<int>(p, n_lr);
    has_local_i = take<int>(p, nb);
  }

  upload_laplace_tables();
  fmmb_plan plan;
  plan.bem_near_kernel = 0;   // the one-warp-per-item kernel: the split kernel's warp shuffles are not emulated
  plan.kind = FMMB_LAPLACE_SPHERICAL_BEM;
  plan.p = P;
  std::memset(&plan.opts, 0, sizeof plan.opts);
  plan.opts.evaluator = treecode ? FMMB_EVAL_TREECODE : FMMB_EVAL_FMM;
  plan.charge_dim = plan.result_dim = 1;
  Tree& T = plan.tree;
  T.n = n; T.nboxes = (int)nb; T.own_b0 = 0; T.own_b1 = n;
  T.perm.from_host(perm, n, nullptr);
  T.bbegin.from_host(bb, nb, nullptr); T.bend.from_host(be, nb, nullptr);
  T.parent.from_host(parent, nb, nullptr);
  T.p2p_off.from_host(off, nb + 1, nullptr); T.p2p_src.from_host(src, ne, nullptr);
  T.p2p_items.from_host(items, ni, nullptr); T.n_p2p_items = (int)ni;
  std::vector<double4> cen(nb), body(n);
  std::vector<int> leaves;
  for (long b = 0; b < nb; ++b) { cen[b] = make_double4(geom[4 * b], geom[4 * b + 1], geom[4 * b + 2], geom[4 * b + 3]); if (leaf[b]) leaves.push_back((int)b); }
  for (long i = 0; i < n; ++i) {                          // tree-ordered panel centres (what build_tree leaves in body)
    const double* v = verts + 9 * (size_t)perm[i];
    body[i] = make_double4(((v[0] + v[3]) + v[6]) / 3, ((v[1] + v[4]) + v[7]) / 3, ((v[2] + v[5]) + v[8]) / 3, 0.0);
  }
  T.center.from_host(cen.data(), nb, nullptr);
  T.body.from_host(body.data(), n, nullptr);
  T.leaves.from_host(leaves.data(), leaves.size(), nullptr); T.nleaves = (int)leaves.size();
  T.own_leaves.from_host(leaves.data(), leaves.size(), nullptr); T.n_own_leaves = (int)leaves.size();
  std::vector<unsigned char> hl(nb, 0);
  std::vector<int> zoff(nb + 1, 0);
  if (n_lr > 0) {
    for (long b = 0; b < nb; ++b) hl[b] = (unsigned char)has_local_i[b];
    T.n_lr = n_lr; T.nlevels = nlevels;
    T.level_off.assign(level_off, level_off + nlevels + 1);
    T.key.from_host(key, nb, nullptr); T.cbegin.from_host(cbegin, nb, nullptr); T.cend.from_host(cend, nb, nullptr);
    T.m2l_off.from_host(m2l_off, nb + 1, nullptr); T.m2l_src.from_host(m2l_src, n_lr, nullptr);
  } else {
    T.m2l_off.from_host(zoff.data(), nb + 1, nullptr); T.m2l_src.resize(1);
  }
  T.has_local.from_host(hl.data(), nb, nullptr);

  bem_setup(&plan, verts, bc, K, -1.0);
  std::vector<double> out(n, -11.0), out2(n, -12.0);
  bem_execute(&plan, q, out.data());
  if (n_lr == 0) bem_execute(&plan, q, out2.data()); else out2 = out;     // (far field: one matvec, the translations dominate)
  std::string o = std::string(argv[1]) + ".out";
  f = fopen(o.c_str(), "wb");
  fwrite(out.data(), 8, out.size(), f);
  fclose(f);
  printf("bem pipeline: n %ld launches %ld translation_calls %d guard_failures %ld repeatable %d nnz %lld\n", n, emu::launches,
         g_translation_calls, emu::guard_failures, (int)(out == out2), (long long)bem_nnz(plan.bem));
  bem_free(plan.bem);
  return 0;
}
