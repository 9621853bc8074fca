// Host engine + C ABI of libsblx.so.  One handle = one grid model resident on one B200.
// See include/sblx.h for the reference interfaces each entry point replaces and kernels.cuh for
// the device side.
#include "../../include/sblx.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <cuda_profiler_api.h>
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the library is loaded with dlopen (see NcclApi), nothing links against it

#include "kernels.cuh"
#include "symbolic.h"

namespace sblx {

static thread_local std::string g_create_error;

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
#define CK(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess) {                                                                                  \
      char buf_[512];                                                                                         \
      snprintf(buf_, sizeof buf_, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call); \
      throw CudaError(buf_);                                                                                  \
    }                                                                                                         \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0, cap = 0;
  // Grows only: a buffer that is large enough is reused (cudaMalloc / cudaFree cost tens of microseconds each, and a
  // small batch stages about two dozen of them - more than its whole solve).  release() really frees.
  void alloc(size_t count) {
    if (count > 0 && count <= cap && p) { n = count; return; }
    release();
    n = cap = count;
    if (count) CK(cudaMalloc(&p, count * sizeof(T)));
  }
  void upload(const std::vector<T>& v, cudaStream_t s = 0) {
    alloc(v.size());
    if (!v.empty()) CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = cap = 0;
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

static double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------------------------------
// host model (no CUDA): pattern, maps, symbolic analysis
// ------------------------------------------------------------------------------------------
struct HostModel {
  int n = 0, m = 0, slack = 0, nbr = 0;
  std::vector<int> y_rowptr, y_col;
  std::vector<double2> y_val;
  std::vector<int> node_of_bus, bus_of_node;
  std::vector<int64_t> jb_rowptr;
  std::vector<int> jb_row, jb_col, jb_ypos;
  Symbolic sym;
  // per bus
  std::vector<double2> S0, V0;
  std::vector<double> vset, qmin, qmax, qload;
  std::vector<unsigned char> type0, lockmask;
  // branches (0-based buses)
  std::vector<int> br_from, br_to;
  std::vector<double2> y11, y12, y21, y22;
  std::vector<double> rating;
  std::vector<unsigned char> br_status;
  // topology helpers
  std::vector<int> adj_ptr, adj_nb, adj_br;   // in-service branch adjacency
  std::vector<int> deg;
  std::vector<int> base_iso;   // buses without any in-service branch in the intact grid
  int base_islands = 1;
};

static void build_pattern(HostModel& H, int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                          int base, const int32_t* brf, const int32_t* brt, int64_t nbr) {
  // union pattern: Ybus entries, their transposes, the diagonal and every branch's four stamps
  std::vector<std::vector<int>> rows(n);
  for (int64_t j = 0; j < n; ++j)
    for (int64_t p = colptr[j] - base; p < colptr[j + 1] - base; ++p) {
      int64_t i = rowval[p] - base;
      if (i < 0 || i >= n) throw std::invalid_argument("sblx_create: Ybus row index out of range");
      rows[i].push_back((int)j);
      rows[j].push_back((int)i);
    }
  for (int64_t i = 0; i < n; ++i) rows[i].push_back((int)i);
  for (int64_t k = 0; k < nbr; ++k) {
    int f = brf[k] - base, t = brt[k] - base;
    if (f < 0 || f >= n || t < 0 || t >= n || f == t) throw std::invalid_argument("sblx_create: branch endpoint out of range");
    rows[f].push_back(t);
    rows[t].push_back(f);
  }
  H.y_rowptr.assign(n + 1, 0);
  H.y_col.clear();
  for (int64_t i = 0; i < n; ++i) {
    auto& r = rows[i];
    std::sort(r.begin(), r.end());
    r.erase(std::unique(r.begin(), r.end()), r.end());
    for (int c : r) H.y_col.push_back(c);
    H.y_rowptr[i + 1] = (int)H.y_col.size();
  }
  H.y_val.assign(H.y_col.size(), make_double2(0.0, 0.0));
  if (nzval) {
    for (int64_t j = 0; j < n; ++j)
      for (int64_t p = colptr[j] - base; p < colptr[j + 1] - base; ++p) {
        int i = (int)(rowval[p] - base);
        auto b = H.y_col.begin() + H.y_rowptr[i], e = H.y_col.begin() + H.y_rowptr[i + 1];
        auto it = std::lower_bound(b, e, (int)j);
        size_t pos = it - H.y_col.begin();
        H.y_val[pos].x += nzval[2 * p];
        H.y_val[pos].y += nzval[2 * p + 1];
      }
  }
}

static void build_nodes_and_symbolic(HostModel& H, const sblx_options& o) {
  const int n = H.n;
  H.node_of_bus.assign(n, -1);
  H.bus_of_node.clear();
  for (int i = 0; i < n; ++i)
    if (i != H.slack) {
      H.node_of_bus[i] = (int)H.bus_of_node.size();
      H.bus_of_node.push_back(i);
    }
  H.m = n - 1;
  H.jb_rowptr.assign(H.m + 1, 0);
  H.jb_row.clear();
  H.jb_col.clear();
  H.jb_ypos.clear();
  for (int a = 0; a < H.m; ++a) {
    int i = H.bus_of_node[a];
    for (int p = H.y_rowptr[i]; p < H.y_rowptr[i + 1]; ++p) {
      int j = H.y_col[p];
      if (j == H.slack) continue;
      H.jb_row.push_back(a);
      H.jb_col.push_back(H.node_of_bus[j]);
      H.jb_ypos.push_back(p);
    }
    H.jb_rowptr[a + 1] = (int64_t)H.jb_col.size();
  }
  SymbolicOptions so;
  so.ordering = o.ordering;
  so.relax_zero_frac = o.relax_zero_frac;
  if (o.relax_small > 0) so.relax_small = o.relax_small;
  if (o.panel_smem_kb >= 4.0 && o.panel_smem_kb <= 200.0) so.smem_budget_bytes = (int)(o.panel_smem_kb * 1024);
  if (const char* e = getenv("SBLX_ND_LEAF")) so.nd_leaf = std::max(4, atoi(e));   // tuning probes
  if (const char* e = getenv("SBLX_RELAX_MAXW")) so.relax_small_max_w = atoi(e);
  analyze(H.m, H.jb_rowptr, H.jb_col, so, H.sym);
  if (H.sym.max_w > 16) throw std::runtime_error("sblx: panel wider than 16 pivots");
}

// Plain C++ walk over the same panels / index maps as k_front + k_backsolve (symbolic self-test).
static double host_selftest(const HostModel& H, uint64_t seed) {
  const Symbolic& S = H.sym;
  const int m = H.m;
  const int64_t nnzB = (int64_t)H.jb_col.size();
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> ud(-1.0, 1.0);
  std::vector<double> Jv(nnzB * 4);
  std::vector<double> b(2 * m), x(2 * m, 0.0);
  for (int a = 0; a < m; ++a) {
    int cnt = (int)(H.jb_rowptr[a + 1] - H.jb_rowptr[a]);
    for (int64_t k = H.jb_rowptr[a]; k < H.jb_rowptr[a + 1]; ++k) {
      for (int q = 0; q < 4; ++q) Jv[4 * k + q] = ud(rng);
      if (H.jb_col[k] == a) {  // off-diagonal-dominant 2x2 pivot like the real Jacobian
        Jv[4 * k + 0] = 0.1 * ud(rng);
        Jv[4 * k + 3] = 0.1 * ud(rng);
        Jv[4 * k + 1] = 4.0 * cnt + ud(rng);
        Jv[4 * k + 2] = 4.0 * cnt + ud(rng);
      }
    }
  }
  for (auto& v : b) v = ud(rng);
  // rhs in elimination order
  std::vector<double> rhs(2 * m), yv(2 * m), xv(2 * m);
  for (int a = 0; a < m; ++a) {
    rhs[2 * S.iperm[a]] = b[2 * a];
    rhs[2 * S.iperm[a] + 1] = b[2 * a + 1];
  }
  std::vector<double> U((size_t)S.u_blocks * 4, 0.0), C((size_t)S.c_blocks * 4 + 8, 0.0);
  auto mm_sub = [](double* t, const double* a, const double* u) {
    t[0] -= a[0] * u[0] + a[1] * u[2];
    t[1] -= a[0] * u[1] + a[1] * u[3];
    t[2] -= a[2] * u[0] + a[3] * u[2];
    t[3] -= a[2] * u[1] + a[3] * u[3];
  };
  const int np = (int)S.panels.size();
  for (int l = 0; l < S.nlevels; ++l)
    for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
      const int pi = S.level_panels[q];
      const Panel& P = S.panels[pi];
      const int w = P.w, h = P.h, nf = w + h;
      std::vector<double> T((size_t)w * nf * 4, 0.0), Lp((size_t)h * w * 4, 0.0), b1(2 * w, 0.0);
      for (int64_t k = 0; k < P.asm_cnt; ++k) {
        int src = S.asm_src[P.asm_ptr + k], dst = S.asm_dst[P.asm_ptr + k];
        double* d = dst < w * nf ? &T[(size_t)dst * 4] : &Lp[(size_t)(dst - w * nf) * 4];
        for (int z = 0; z < 4; ++z) d[z] = Jv[4 * (size_t)src + z];
      }
      for (int p = 0; p < w; ++p) { b1[2 * p] = rhs[2 * (P.first + p)]; b1[2 * p + 1] = rhs[2 * (P.first + p) + 1]; }
      for (int ci = 0; ci < P.nchild; ++ci) {
        int c = S.child_list[P.child_ptr + ci];
        const Panel& Cp = S.panels[c];
        const int hc = Cp.h, kc = S.kpiv[c];
        const int* rel = &S.rel[S.rel_ptr[c]];
        const double* Cm = &C[(size_t)Cp.c_off * 4];
        const double* cv = Cm + (size_t)hc * hc * 4;
        for (int a = 0; a < hc; ++a)
          for (int bb = 0; bb < hc; ++bb) {
            if (a >= kc && bb >= kc) continue;
            int la = rel[a], lb = rel[bb];
            double* d = la < w ? &T[((size_t)la * nf + lb) * 4] : &Lp[((size_t)lb * h + (la - w)) * 4];
            for (int z = 0; z < 4; ++z) d[z] += Cm[((size_t)a * hc + bb) * 4 + z];
          }
        for (int a = 0; a < kc; ++a) { b1[2 * rel[a]] += cv[2 * a]; b1[2 * rel[a] + 1] += cv[2 * a + 1]; }
      }
      for (int p = 0; p < w; ++p) {
        double* d = &T[((size_t)p * nf + p) * 4];
        double det = d[0] * d[3] - d[1] * d[2];
        double D[4] = {d[3] / det, -d[1] / det, -d[2] / det, d[0] / det};
        for (int c = p + 1; c < nf; ++c) {
          double* u = &T[((size_t)p * nf + c) * 4];
          double t[4] = {D[0] * u[0] + D[1] * u[2], D[0] * u[1] + D[1] * u[3], D[2] * u[0] + D[3] * u[2], D[2] * u[1] + D[3] * u[3]};
          for (int z = 0; z < 4; ++z) u[z] = t[z];
        }
        double y0 = D[0] * b1[2 * p] + D[1] * b1[2 * p + 1], y1 = D[2] * b1[2 * p] + D[3] * b1[2 * p + 1];
        b1[2 * p] = y0; b1[2 * p + 1] = y1;
        for (int i = p + 1; i < w; ++i) {
          const double* a = &T[((size_t)i * nf + p) * 4];
          for (int c = p + 1; c < nf; ++c) mm_sub(&T[((size_t)i * nf + c) * 4], a, &T[((size_t)p * nf + c) * 4]);
          b1[2 * i] -= a[0] * y0 + a[1] * y1;
          b1[2 * i + 1] -= a[2] * y0 + a[3] * y1;
        }
        for (int a = 0; a < h; ++a) {
          const double* lv = &Lp[((size_t)p * h + a) * 4];
          for (int c = p + 1; c < w; ++c) mm_sub(&Lp[((size_t)c * h + a) * 4], lv, &T[((size_t)p * nf + c) * 4]);
        }
      }
      std::copy(T.begin(), T.end(), U.begin() + (size_t)P.u_off * 4);
      for (int p = 0; p < w; ++p) { yv[2 * (P.first + p)] = b1[2 * p]; yv[2 * (P.first + p) + 1] = b1[2 * p + 1]; }
      double* Co = &C[(size_t)P.c_off * 4];
      double* cvo = Co + (size_t)h * h * 4;
      for (int a = 0; a < h; ++a) {
        for (int bb = 0; bb < h; ++bb) {
          double acc[4] = {0, 0, 0, 0};
          for (int p = 0; p < w; ++p) mm_sub(acc, &Lp[((size_t)p * h + a) * 4], &T[((size_t)p * nf + w + bb) * 4]);
          for (int ci = 0; ci < P.nchild; ++ci) {
            int c = S.child_list[P.child_ptr + ci];
            const Panel& Cp = S.panels[c];
            const int* inv = &S.inv[S.inv_ptr[P.child_ptr + ci]];
            if (inv[a] >= 0 && inv[bb] >= 0)
              for (int z = 0; z < 4; ++z) acc[z] += C[(size_t)Cp.c_off * 4 + ((size_t)inv[a] * Cp.h + inv[bb]) * 4 + z];
          }
          for (int z = 0; z < 4; ++z) Co[((size_t)a * h + bb) * 4 + z] = acc[z];
        }
        double v0 = 0, v1 = 0;
        for (int p = 0; p < w; ++p) {
          const double* lv = &Lp[((size_t)p * h + a) * 4];
          v0 -= lv[0] * b1[2 * p] + lv[1] * b1[2 * p + 1];
          v1 -= lv[2] * b1[2 * p] + lv[3] * b1[2 * p + 1];
        }
        for (int ci = 0; ci < P.nchild; ++ci) {
          int c = S.child_list[P.child_ptr + ci];
          const Panel& Cp = S.panels[c];
          int ia = S.inv[S.inv_ptr[P.child_ptr + ci] + a];
          if (ia >= 0) {
            const double* cvc = &C[(size_t)Cp.c_off * 4 + (size_t)Cp.h * Cp.h * 4];
            v0 += cvc[2 * ia];
            v1 += cvc[2 * ia + 1];
          }
        }
        cvo[2 * a] = v0;
        cvo[2 * a + 1] = v1;
      }
    }
  (void)np;
  for (int l = S.nlevels - 1; l >= 0; --l)
    for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
      const Panel& P = S.panels[S.level_panels[q]];
      const int w = P.w, h = P.h, nf = w + h;
      const double* Up = &U[(size_t)P.u_off * 4];
      for (int p = w - 1; p >= 0; --p) {
        double a0 = yv[2 * (P.first + p)], a1 = yv[2 * (P.first + p) + 1];
        for (int c = p + 1; c < nf; ++c) {
          int g = c < w ? P.first + c : S.struct_nodes[P.struct_ptr + (c - w)];
          const double* u = &Up[((size_t)p * nf + c) * 4];
          a0 -= u[0] * xv[2 * g] + u[1] * xv[2 * g + 1];
          a1 -= u[2] * xv[2 * g] + u[3] * xv[2 * g + 1];
        }
        xv[2 * (P.first + p)] = a0;
        xv[2 * (P.first + p) + 1] = a1;
      }
    }
  for (int a = 0; a < m; ++a) { x[2 * a] = xv[2 * S.iperm[a]]; x[2 * a + 1] = xv[2 * S.iperm[a] + 1]; }
  // residual
  double rn = 0, bn = 0;
  for (int a = 0; a < m; ++a) {
    double r0 = -b[2 * a], r1 = -b[2 * a + 1];
    for (int64_t k = H.jb_rowptr[a]; k < H.jb_rowptr[a + 1]; ++k) {
      int c = H.jb_col[k];
      r0 += Jv[4 * k] * x[2 * c] + Jv[4 * k + 1] * x[2 * c + 1];
      r1 += Jv[4 * k + 2] * x[2 * c] + Jv[4 * k + 3] * x[2 * c + 1];
    }
    rn = std::max(rn, std::max(std::fabs(r0), std::fabs(r1)));
    bn = std::max(bn, std::max(std::fabs(b[2 * a]), std::fabs(b[2 * a + 1])));
  }
  return rn / std::max(bn, 1e-300);
}

// ------------------------------------------------------------------------------------------
struct LaunchGroup {
  int off, cnt, nt, smem;     // range in d_level_panels, threads, dynamic smem bytes
  int subst_chunks = 0, schur_tiles = 0, wmax = 0;   // split path: grid.z of k_subst / k_schur, widest panel
  int lanes = 0;              // k_front: lanes per scenario (< nt: several scenarios share one warp)
  int wide = 0;               // k_front<64>: pivot blocks of 9..16 buses
  int reg = 0;                // k_front_reg: tiny fronts held in registers
  mutable int pf_delta = -1;  // L2 prefetch distance in CTAs (lazily: prefetch multiplier x resident CTAs of this launch)
};

struct HostResults {
  std::vector<int> conv, status, iter, nsw, islands;   // status: -1 solved on device, -2 island-wise composite, >= 0 host verdict
  std::vector<double> fm, vmin, vmax, load;
  // island-wise solves (rectangular_network_solver.jl:1919-2243): scenario s = sub-scenarios [sub_first[s], +sub_count[s])
  std::vector<int> sub_first, sub_count;
  std::vector<std::vector<int>> sub_buses;   // per sub-scenario (index - B): buses of its island
  std::vector<int> parent;                   // per sub-scenario (index - B): the caller scenario it belongs to
  int64_t Bint = 0;                          // scenarios on the device = B + sub-scenarios
};

// NCCL, loaded on first use.  libsblx.so carries no link-time dependency on it: a CPU-only process can load the
// library and run the host-side entry points; the multi-GPU entries fail loudly if NCCL cannot be found.
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("SBLX_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) return;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))dlsym(api.lib, "ncclCommInitAll");
    api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
  });
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.AllGather || !api.CommDestroy)
    throw std::runtime_error("sblx: NCCL (libnccl.so.2) could not be loaded - set SBLX_NCCL_LIB to its path");
  return api;
}
#define NCK(call)                                                                                   \
  do {                                                                                              \
    ncclResult_t r_ = (call);                                                                       \
    if (r_ != ncclSuccess) {                                                                        \
      char buf_[512];                                                                               \
      NcclApi& a_ = nccl_api();                                                                     \
      snprintf(buf_, sizeof buf_, "NCCL error %s at %s:%d (%s)", a_.GetErrorString ? a_.GetErrorString(r_) : "?", __FILE__, __LINE__, #call); \
      throw CudaError(buf_);                                                                        \
    }                                                                                               \
  } while (0)
static void nccl_destroy(ncclComm_t c) {
  try { nccl_api().CommDestroy(c); } catch (...) {}
}

// what the caller wants back besides the per-scenario summary
struct Want {
  bool v = false, types = false, flows = false;
  int64_t viol_cap = 0, over_cap = 0, qlog_cap = 0;
};

}  // namespace sblx

using namespace sblx;

struct sblx_handle {
  sblx_options opt{};
  HostModel H;
  std::string err;
  std::mutex mu;
  int device = 0, sm_count = 148;
  double pf_mult = 1.0;       // SBLX_PF_MULT: prefetch distance in units of the resident CTA count
  cudaStream_t stream = nullptr;
  bool cuda_ready = false;
  // device model
  DevBuf<int> d_y_rowptr, d_y_col, d_node_of_bus, d_bus_of_node, d_elim_of_node, d_node_of_elim, d_jb_row, d_jb_col, d_jb_ypos;
  DevBuf<double2> d_y_val, d_S0, d_V0, d_y11, d_y12, d_y21, d_y22;
  DevBuf<double> d_vset, d_qmin, d_qmax, d_qload, d_rating;
  DevBuf<unsigned char> d_type0, d_lock, d_br_status;
  DevBuf<int> d_br_from, d_br_to;
  DevBuf<PanelDev> d_panels, d_panels_lv;     // by panel id / in level_panels order (contiguous per launch group)
  DevBuf<int> d_struct, d_child, d_rel, d_kpiv, d_inv, d_asm_src, d_asm_dst, d_level_panels, d_ing_ptr, d_ing_src, d_ing_pairs, d_reg_words, d_reg_vptr, d_reg_vsrc, d_ing_xptr, d_ing_xsrc, d_vec_ptr, d_vec_src;
  DevBuf<ChildDev> d_childdev;
  DevBuf<long long> d_rel_ptr, d_inv_ptr;
  std::vector<std::vector<LaunchGroup>> front_groups;   // per level
  std::vector<std::vector<LaunchGroup>> solve_groups;
  std::vector<LaunchGroup> level_all;                   // per level: all its panels (k_pivot)
  long long l_blocks = 1;
  DevModel M{};
  // slots
  int W = 0;
  size_t bytes_per_slot = 0;
  DevBuf<double2> s_V, s_I, s_S, s_rhs, s_yv, s_xv;
  DevBuf<double> s_qload, s_Jv, s_U, s_C, s_fm, s_Lw;
  DevBuf<double2> s_Dw;
  DevBuf<unsigned char> s_type, s_iso, s_evcnt, s_evact;
  DevBuf<short> s_evlast;
  DevBuf<int> s_scen, s_iter, s_state, s_changed, s_nev, s_singular, s_numerical, s_reason, s_viol, s_maxsw, s_rated, s_out_cnt, s_out_br;
  DevBuf<int> s_nvv, s_nov, s_tiny, s_tinyit, s_refine, s_nrefine, l_refine;
  DevBuf<double> s_fnorm;
  DevBuf<double2> s_xsave;
  const Lists* Lover = nullptr;   // list override of the factor / solve launches (refinement pass)
  DevBuf<unsigned long long> s_mis, s_mis2, s_pvres, s_vminb, s_vmaxb, s_loadb;
  DevBuf<int> l_active, l_fresh, l_step, l_fin, l_counts;
  Slots A{};
  Lists L{};
  // staged batch
  int64_t B = 0;
  bool staged = false, ran = false, sweep_q = false;
  DevBuf<int> b_queue, b_out_ptr, b_out_br, b_iso_ptr, b_iso_bus, b_ref_ptr, b_ref_bus, b_lk_ptr, b_lk_bus;
  DevBuf<int> b_parent, b_host_status, b_sub_first, b_sub_count, b_islands;
  DevBuf<double2> b_S, b_Vstart;
  DevBuf<double> b_ql, b_scale;
  DevBuf<int> r_conv, r_status, r_iter, r_nsw, r_nvv, r_nov, r_tiny;
  DevBuf<double> r_fm, r_vmin, r_vmax, r_load, r_over_pct;
  DevBuf<double2> r_V, r_flows;
  DevBuf<int2> r_viol_list, r_over_list;
  DevBuf<int4> r_qlog_list;
  int64_t qlog_cap = 0;
  DevBuf<unsigned long long> r_list_counts;
  DevBuf<unsigned char> r_types;
  DevBuf<SummaryRec> r_summary, g_summary;     // own records (padded to the common shard size) / gathered records of all ranks
  int64_t summary_cap = 0;
  int64_t viol_cap = 0, over_cap = 0;
  // device-side sweep parameters of the staged batch
  int sweep_mode = 0;
  uint64_t sweep_seed = 0;
  int64_t sweep_first = 0;
  double sweep_sigma = 0.0, sweep_lo = 0.0, sweep_hi = 0.0;
  // multi-GPU: communicator of this handle's rank (nullptr = single GPU)
  ncclComm_t comm = nullptr;
  bool comm_owned = false;
  int nranks = 1, rank = 0;
  Batch Bt{};
  HostResults hostres;       // host-decided scenarios (islanded / bad branch), status < 0 = solved on device
  int64_t nq = 0;
  int* pinned_counts = nullptr;
  void* pinned_stage[2] = {nullptr, nullptr};   // staging buffers of h2d_staged
  cudaEvent_t stage_ev[2] = {};
  sblx_stats stats{};
  cudaEvent_t ev[8] = {};
  cudaEvent_t ring_ev[4] = {};     // look-ahead ticks of small batches: "counters of tick k are on the host"

  ~sblx_handle() {
    if (cuda_ready) {
      cudaSetDevice(device);
      for (auto& e : ev)
        if (e) cudaEventDestroy(e);
      for (auto& e : ring_ev)
        if (e) cudaEventDestroy(e);
      if (pinned_counts) cudaFreeHost(pinned_counts);
      for (int k = 0; k < 2; ++k) {
        if (pinned_stage[k]) cudaFreeHost(pinned_stage[k]);
        if (stage_ev[k]) cudaEventDestroy(stage_ev[k]);
      }
      if (stream) cudaStreamDestroy(stream);
      if (comm && comm_owned) sblx::nccl_destroy(comm);
    }
  }
};

namespace sblx {

static void upload_model(sblx_handle* h) {
  HostModel& H = h->H;
  cudaStream_t s = h->stream;
  const Symbolic& S = H.sym;
  h->d_y_rowptr.upload(H.y_rowptr, s);
  h->d_y_col.upload(H.y_col, s);
  h->d_y_val.upload(H.y_val, s);
  h->d_node_of_bus.upload(H.node_of_bus, s);
  h->d_bus_of_node.upload(H.bus_of_node, s);
  h->d_elim_of_node.upload(S.iperm, s);
  h->d_node_of_elim.upload(S.perm, s);
  h->d_jb_row.upload(H.jb_row, s);
  h->d_jb_col.upload(H.jb_col, s);
  h->d_jb_ypos.upload(H.jb_ypos, s);
  h->d_S0.upload(H.S0, s);
  h->d_V0.upload(H.V0, s);
  h->d_vset.upload(H.vset, s);
  h->d_qmin.upload(H.qmin, s);
  h->d_qmax.upload(H.qmax, s);
  h->d_qload.upload(H.qload, s);
  h->d_type0.upload(H.type0, s);
  h->d_lock.upload(H.lockmask, s);
  h->d_br_from.upload(H.br_from, s);
  h->d_br_to.upload(H.br_to, s);
  h->d_y11.upload(H.y11, s);
  h->d_y12.upload(H.y12, s);
  h->d_y21.upload(H.y21, s);
  h->d_y22.upload(H.y22, s);
  h->d_rating.upload(H.rating, s);
  h->d_br_status.upload(H.br_status, s);
  std::vector<int> l_off(S.panels.size(), 0);
  h->l_blocks = 1;
  for (int l = 0; l < S.nlevels; ++l) {
    long long off = 0;
    for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
      const Panel& P = S.panels[S.level_panels[q]];
      l_off[S.level_panels[q]] = (int)off;
      off += (long long)P.h * P.w;
    }
    h->l_blocks = std::max(h->l_blocks, off);
  }
  std::vector<PanelDev> pd(S.panels.size());
  for (size_t i = 0; i < pd.size(); ++i) {
    const Panel& P = S.panels[i];
    pd[i] = PanelDev{P.first, P.w, P.h, (int)P.nchild, (int)P.struct_ptr, (int)P.child_ptr, (int)P.ing_base, l_off[i], (long long)P.u_off, (long long)P.c_off, (int)P.ingf_base, (int)P.reg_base, (int)P.reg_vbase, 0};
  }
  h->d_panels.upload(pd, s);
  h->d_struct.upload(S.struct_nodes, s);
  h->d_child.upload(S.child_list, s);
  h->d_rel.upload(S.rel, s);
  h->d_kpiv.upload(S.kpiv, s);
  h->d_inv.upload(S.inv, s);
  h->d_asm_src.upload(S.asm_src, s);
  h->d_asm_dst.upload(S.asm_dst, s);
  h->d_ing_ptr.upload(S.ing_ptr, s);
  h->d_ing_src.upload(S.ing_src, s);
  h->d_ing_pairs.upload(S.ing_pairs, s);
  h->d_reg_words.upload(S.reg_words.empty() ? std::vector<int>(1, 0) : S.reg_words, s);
  h->d_reg_vptr.upload(S.reg_vptr.empty() ? std::vector<int>(1, 0) : S.reg_vptr, s);
  h->d_reg_vsrc.upload(S.reg_vsrc.empty() ? std::vector<int>(1, 0) : S.reg_vsrc, s);
  h->d_ing_xptr.upload(S.ing_xptr, s);
  h->d_ing_xsrc.upload(S.ing_xsrc.empty() ? std::vector<int>(1, 0) : S.ing_xsrc, s);
  h->d_vec_ptr.upload(S.vec_ptr, s);
  h->d_vec_src.upload(S.vec_src, s);
  {
    std::vector<ChildDev> cd(S.child_list.size());
    for (size_t k = 0; k < cd.size(); ++k) {
      const Panel& C = S.panels[S.child_list[k]];
      cd[k] = ChildDev{(long long)C.c_off, C.h, (int)S.inv_ptr[k], S.kpiv[S.child_list[k]], 0};
    }
    h->d_childdev.upload(cd, s);
  }
  std::vector<long long> rp(S.rel_ptr.begin(), S.rel_ptr.end()), ip(S.inv_ptr.begin(), S.inv_ptr.end());
  h->d_rel_ptr.upload(rp, s);
  h->d_inv_ptr.upload(ip, s);
  // launch groups: per level, panels sorted by size class
  // size class by panel bytes; the two small classes factor the pivot block with an 8-pivot register column
  // Class 9: fronts of <= 8 buses in total, factored in registers (k_front_reg).  Classes 8/0/1 are the lane-group kernels (4 / 8 / 16 lanes per scenario: pivot blocks of <= 2 / <= 4 / <= 8
  // buses, 8 / 4 / 2 scenarios per warp); classes 2..6 give the whole CTA (32..512 threads) to one front; class 7
  // is the 64-thread kernel with a 16-row register column (small fronts with pivot blocks of 9..16 buses).
  int g8_bytes = 4 * 1024, g16_bytes = 8 * 1024, g4_bytes = 1024;
  const bool reg_class = getenv("SBLX_NO_REG_CLASS") == nullptr;                // tuning probe
  if (const char* e = getenv("SBLX_G4_KB10")) g4_bytes = atoi(e) * 102;        // tuning probe, tenths of a KB (0 disables the class)
  if (const char* e = getenv("SBLX_G8_KB")) g8_bytes = atoi(e) * 1024;       // tuning probes (0 disables the class)
  if (const char* e = getenv("SBLX_G16_KB")) g16_bytes = atoi(e) * 1024;
  g8_bytes = std::min(g8_bytes, 12 * 1024);
  g16_bytes = std::min(g16_bytes, 24 * 1024);
  int c2_bytes = 72 * 1024;
  if (const char* e = getenv("SBLX_C2_KB")) c2_bytes = std::min(atoi(e), 80) * 1024;
  auto front_class_b = [c2_bytes](int bytes) { return bytes <= 6 * 1024 ? 0 : bytes <= 24 * 1024 ? 1 : bytes <= c2_bytes ? 2 : bytes <= 110 * 1024 ? 3 : 4; };
  auto front_class_p = [&](const Panel& P) {
    const int bytes = panel_smem_bytes(P.w, P.h);
    if (reg_class && P.reg_base >= 0) return 9;       // w + h <= 8: factored entirely in registers (k_front_reg)
    if (P.w <= 2 && bytes <= g4_bytes) return 8;      // 4 lanes per scenario, 8 scenarios per warp
    if (P.w <= 4 && bytes <= g8_bytes) return 0;
    if (P.w <= 8 && bytes <= g16_bytes) return 1;
    int c = front_class_b(bytes);
    return (P.w > 8 && c < 2) ? 7 : 2 + c;          // 7: small front with a wide pivot block (64 threads, 16-row register column)
  };
  static const int front_nt[10] = {32, 32, 32, 64, 128, 256, 512, 64, 32, 32};
  static const int front_g[10] = {8, 16, 32, 64, 128, 256, 512, 64, 4, 8};
  static const int front_wide[10] = {0, 0, 0, 0, 0, 0, 0, 1, 0, 0};
  std::vector<int> lp;
  h->front_groups.assign(S.nlevels, {});
  h->solve_groups.assign(S.nlevels, {});
  for (int l = 0; l < S.nlevels; ++l) {
    std::vector<int> ids(S.level_panels.begin() + S.level_ptr[l], S.level_panels.begin() + S.level_ptr[l + 1]);
    std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) {
      return front_class_p(S.panels[a]) < front_class_p(S.panels[b]);
    });
    size_t i = 0;
    while (i < ids.size()) {
      int c = front_class_p(S.panels[ids[i]]);
      size_t j = i;
      int smem = 0, hs = 0;
      while (j < ids.size() && front_class_p(S.panels[ids[j]]) == c) {
        smem = std::max(smem, panel_smem_bytes(S.panels[ids[j]].w, S.panels[ids[j]].h));
        hs = std::max(hs, S.panels[ids[j]].h + S.panels[ids[j]].w);
        ++j;
      }
      LaunchGroup g{(int)lp.size() + (int)i, (int)(j - i), front_nt[c], smem};
      g.lanes = front_g[c];
      g.wide = front_wide[c];
      g.reg = c == 9;
      for (size_t q = i; q < j; ++q) {
        const Panel& P = S.panels[ids[q]];
        const int hpad = (P.h + 31) & ~31;
        g.subst_chunks = std::max(g.subst_chunks, (2 * hpad + 127) / 128);
        g.schur_tiles = std::max(g.schur_tiles, ((P.h + 31) / 32) * ((P.h + 63) / 64));
        g.wmax = std::max(g.wmax, P.w);
      }
      h->front_groups[l].push_back(g);
      LaunchGroup gs{g.off, g.cnt, hs <= 32 ? 32 : hs <= 96 ? 64 : 128, hs * 16 + 64 + g.wmax * g.wmax * 32};
      gs.lanes = g.lanes < 32 ? 8 : gs.nt;      // the lane-group front classes have pivot blocks of <= 8 buses
      h->solve_groups[l].push_back(gs);
      i = j;
    }
    LaunchGroup ga{(int)lp.size(), (int)ids.size(), 128, 0};
    if ((int)h->level_all.size() <= l) h->level_all.resize(l + 1);
    h->level_all[l] = ga;
    lp.insert(lp.end(), ids.begin(), ids.end());
  }
  h->d_level_panels.upload(lp, s);
  {
    std::vector<PanelDev> plv(lp.size());
    for (size_t i = 0; i < lp.size(); ++i) plv[i] = pd[lp[i]];
    h->d_panels_lv.upload(plv, s);
  }
  CK(cudaStreamSynchronize(s));
  // opt-in shared memory
  CK(cudaFuncSetAttribute(k_front<32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
  CK(cudaFuncSetAttribute(k_front<32, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
  CK(cudaFuncSetAttribute(k_front<32, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
  CK(cudaFuncSetAttribute(k_front<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1024));
  CK(cudaFuncSetAttribute(k_front<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
  CK(cudaFuncSetAttribute(k_front<64, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
  CK(cudaFuncSetAttribute(k_front<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
  CK(cudaFuncSetAttribute(k_front<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
  CK(cudaFuncSetAttribute(k_front<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CK(cudaFuncSetAttribute(k_subst, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024));
  CK(cudaFuncSetAttribute(k_schur, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CK(cudaFuncSetAttribute(k_backsolve<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CK(cudaFuncSetAttribute(k_backsolve<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CK(cudaFuncSetAttribute(k_backsolve<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));

  DevModel& M = h->M;
  M.n = H.n; M.m = H.m; M.slack = H.slack; M.nbr = H.nbr; M.nnzB = (int)H.jb_col.size(); M.W = 0;
  M.y_rowptr = h->d_y_rowptr.p; M.y_col = h->d_y_col.p; M.y_val = h->d_y_val.p;
  M.node_of_bus = h->d_node_of_bus.p; M.bus_of_node = h->d_bus_of_node.p;
  M.elim_of_node = h->d_elim_of_node.p; M.node_of_elim = h->d_node_of_elim.p;
  M.jb_row = h->d_jb_row.p; M.jb_col = h->d_jb_col.p; M.jb_ypos = h->d_jb_ypos.p;
  M.S0 = h->d_S0.p; M.V0 = h->d_V0.p; M.vset = h->d_vset.p; M.qmin = h->d_qmin.p; M.qmax = h->d_qmax.p;
  M.qload = h->d_qload.p; M.type0 = h->d_type0.p; M.lockmask = h->d_lock.p;
  M.br_from = h->d_br_from.p; M.br_to = h->d_br_to.p; M.y11 = h->d_y11.p; M.y12 = h->d_y12.p; M.y21 = h->d_y21.p; M.y22 = h->d_y22.p;
  M.rating = h->d_rating.p; M.br_status = h->d_br_status.p;
  M.panels = h->d_panels.p; M.struct_nodes = h->d_struct.p; M.child_list = h->d_child.p; M.rel = h->d_rel.p;
  M.rel_ptr = h->d_rel_ptr.p; M.kpiv = h->d_kpiv.p; M.inv = h->d_inv.p; M.inv_ptr = h->d_inv_ptr.p;
  M.asm_src = h->d_asm_src.p; M.asm_dst = h->d_asm_dst.p;
  M.ing_ptr = h->d_ing_ptr.p; M.ing_src = h->d_ing_src.p;
  M.ing_pairs = reinterpret_cast<const int2*>(h->d_ing_pairs.p); M.reg_words = h->d_reg_words.p; M.reg_vptr = h->d_reg_vptr.p; M.reg_vsrc = h->d_reg_vsrc.p; M.ing_xptr = h->d_ing_xptr.p; M.ing_xsrc = h->d_ing_xsrc.p; M.vec_ptr = h->d_vec_ptr.p; M.vec_src = h->d_vec_src.p;
  M.childdev = h->d_childdev.p;
  M.u_blocks = S.u_blocks; M.c_blocks = S.c_blocks + 2; M.l_blocks = h->l_blocks;
  const sblx_options& o = h->opt;
  M.tol = o.tol; M.damp = o.damp; M.hyst = o.q_hyst_pu; M.vm_min = o.vm_min_pu; M.vm_max = o.vm_max_pu; M.base_mva = o.base_mva;
  M.pivot_tol = o.pivot_tol_rel > 0.0 ? o.pivot_tol_rel : 0.0;
  M.max_iter = o.max_iter; M.qlim_enabled = o.qlimits_enabled; M.qlim_start = o.qlimit_start_iter; M.cooldown = o.cooldown_iters;
  M.guard = o.guard_max_switches; M.freeze = o.freeze_after_repeated_switching;
  M.allow_reenable = (o.cooldown_iters > 0) || (o.q_hyst_pu > 0.0);   // rectangular_network_solver.jl:663-665
  M.pf_mode = 0;
  if (const char* e = getenv("SBLX_PF_MODE")) M.pf_mode = atoi(e);     // tuning probes
  if (const char* e = getenv("SBLX_PF_MULT")) h->pf_mult = atof(e);
  CK(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device));
}

static size_t slot_bytes(const HostModel& H) {
  const Symbolic& S = H.sym;
  size_t n = H.n, m = H.m;
  return n * (16 * 3 + 8 + 4 + 2) + m * 16 * 3 + (size_t)H.jb_col.size() * 32 + (size_t)(S.u_blocks + S.c_blocks + 2) * 32 +
         (size_t)MAXOUT * 4 + 200;
}

static void ensure_cuda(sblx_handle* h) {
  if (h->cuda_ready) {
    CK(cudaSetDevice(h->device));
    return;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) throw CudaError("no CUDA device available (libsblx has no CPU fallback)");
  int dev = h->opt.device;
  if (dev < 0) CK(cudaGetDevice(&dev));
  if (dev >= ndev) throw std::invalid_argument("sblx: device ordinal out of range");
  CK(cudaSetDevice(dev));
  h->device = dev;
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  for (auto& ev : h->ev) CK(cudaEventCreate(&ev));
  for (auto& ev : h->ring_ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CK(cudaMallocHost(&h->pinned_counts, 8 * 5 * sizeof(int)));    // [0..8): the synchronous loop, then 4 ring entries
  h->cuda_ready = true;
  upload_model(h);
}

static void ensure_slots(sblx_handle* h, int64_t want) {
  HostModel& H = h->H;
  size_t freeb = 0, totalb = 0;
  CK(cudaMemGetInfo(&freeb, &totalb));
  h->bytes_per_slot = slot_bytes(H);
  // memory already held by the current slot set counts as available
  size_t avail = freeb + (size_t)h->W * h->bytes_per_slot;
  int64_t cap = (int64_t)((double)avail * h->opt.mem_fraction / (double)h->bytes_per_slot);
  cap = std::min<int64_t>(cap, 32768);
  if (h->opt.max_slots > 0) cap = std::min<int64_t>(cap, h->opt.max_slots);
  cap = std::max<int64_t>(cap, 1);
  int64_t W = std::min<int64_t>(cap, std::max<int64_t>(want, 1));
  W = std::min<int64_t>(cap, (W + 31) / 32 * 32);
  if (W <= h->W) return;
  const size_t n = H.n, m = H.m, nw = n * (size_t)W;
  h->s_V.alloc(nw); h->s_I.alloc(nw); h->s_S.alloc(nw); h->s_qload.alloc(nw);
  h->s_type.alloc(nw); h->s_iso.alloc(nw); h->s_evcnt.alloc(nw); h->s_evact.alloc(nw); h->s_evlast.alloc(nw);
  h->s_rhs.alloc(m * W); h->s_yv.alloc(m * W); h->s_xv.alloc(m * W);
  h->s_Jv.alloc((size_t)H.jb_col.size() * 4 * W);
  h->s_U.alloc((size_t)h->M.u_blocks * 4 * W);
  h->s_C.alloc((size_t)h->M.c_blocks * 4 * W);
  if (h->opt.factor_mode == 1) {
    h->s_Lw.alloc((size_t)h->M.l_blocks * 4 * W);
    h->s_Dw.alloc((size_t)2 * m * W);
  }
  h->s_scen.alloc(W); h->s_iter.alloc(W); h->s_state.alloc(W); h->s_changed.alloc(W); h->s_nev.alloc(W);
  h->s_singular.alloc(W); h->s_numerical.alloc(W); h->s_reason.alloc(W); h->s_viol.alloc(W); h->s_maxsw.alloc(W);
  h->s_rated.alloc(W); h->s_out_cnt.alloc(W); h->s_out_br.alloc((size_t)W * MAXOUT); h->s_fm.alloc(W);
  h->s_nvv.alloc(W); h->s_nov.alloc(W); h->s_tiny.alloc(W); h->s_tinyit.alloc(W);
  h->s_refine.alloc(W); h->s_nrefine.alloc(W); h->l_refine.alloc(W); h->s_fnorm.alloc(W); h->s_xsave.alloc(m * W);
  h->s_mis.alloc(W); h->s_mis2.alloc(W); h->s_pvres.alloc(W); h->s_vminb.alloc(W); h->s_vmaxb.alloc(W); h->s_loadb.alloc(W);
  h->l_active.alloc(W); h->l_fresh.alloc(W); h->l_step.alloc(W); h->l_fin.alloc(W); h->l_counts.alloc(8);
  h->W = (int)W;
  h->M.W = (int)W;
  Slots& A = h->A;
  A.V = h->s_V.p; A.I = h->s_I.p; A.S = h->s_S.p; A.qload_s = nullptr;
  A.type = h->s_type.p; A.iso = h->s_iso.p; A.evcnt = h->s_evcnt.p; A.evact = h->s_evact.p; A.evlast = h->s_evlast.p;
  A.rhs = h->s_rhs.p; A.Jv = h->s_Jv.p; A.U = h->s_U.p; A.C = h->s_C.p; A.yv = h->s_yv.p; A.xv = h->s_xv.p;
  A.Lw = h->s_Lw.p; A.Dw = h->s_Dw.p;
  A.scen = h->s_scen.p; A.iter = h->s_iter.p; A.state = h->s_state.p; A.changed = h->s_changed.p; A.nev = h->s_nev.p;
  A.singular = h->s_singular.p; A.numerical = h->s_numerical.p; A.reason = h->s_reason.p;
  A.mis_bits = h->s_mis.p; A.mis2_bits = h->s_mis2.p; A.fm = h->s_fm.p;
  A.pvres_bits = h->s_pvres.p; A.vmin_bits = h->s_vminb.p; A.vmax_bits = h->s_vmaxb.p; A.load_bits = h->s_loadb.p;
  A.viol = h->s_viol.p; A.maxsw = h->s_maxsw.p; A.rated = h->s_rated.p; A.out_cnt = h->s_out_cnt.p; A.out_br = h->s_out_br.p;
  A.nvv = h->s_nvv.p; A.nov = h->s_nov.p; A.tiny = h->s_tiny.p; A.tinyit = h->s_tinyit.p;
  A.refine = h->s_refine.p; A.nrefine = h->s_nrefine.p; A.fnorm = h->s_fnorm.p; A.xsave = h->s_xsave.p;
  Lists& L = h->L;
  L.active = h->l_active.p; L.fresh = h->l_fresh.p; L.step = h->l_step.p; L.fin = h->l_fin.p; L.counts = h->l_counts.p;
  L.refine = h->l_refine.p;
}

// ------------------------------------------------------------------------------------------
// host topology analysis per scenario: isolated buses (network.jl:1881-1930) and islands
// (ac_islands.jl:50-93, rectangular_network_solver.jl:1914-1935)
// ------------------------------------------------------------------------------------------
static void build_topology(HostModel& H) {
  const int n = H.n;
  H.deg.assign(n, 0);
  for (int k = 0; k < H.nbr; ++k)
    if (H.br_status[k]) { H.deg[H.br_from[k]]++; H.deg[H.br_to[k]]++; }
  H.adj_ptr.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) H.adj_ptr[i + 1] = H.adj_ptr[i] + H.deg[i];
  H.adj_nb.assign(H.adj_ptr[n], 0);
  H.adj_br.assign(H.adj_ptr[n], 0);
  std::vector<int> fill(H.adj_ptr.begin(), H.adj_ptr.end() - 1);
  for (int k = 0; k < H.nbr; ++k)
    if (H.br_status[k]) {
      int f = H.br_from[k], t = H.br_to[k];
      H.adj_nb[fill[f]] = t; H.adj_br[fill[f]++] = k;
      H.adj_nb[fill[t]] = f; H.adj_br[fill[t]++] = k;
    }
  H.base_iso.clear();
  for (int i = 0; i < n; ++i)
    if (H.deg[i] == 0) H.base_iso.push_back(i);
}

struct TopoScratch {
  std::vector<int> markA, markB, parent;
  int stamp = 0;
  std::vector<int> qa, qb;
};

// returns island count; fills iso (sorted). status_out: -1 = solve on device, else SBLX_ST_*
static int analyze_scenario(const HostModel& H, const int* out, int nout, TopoScratch& ts, std::vector<int>& iso, int& status_out,
                            std::vector<std::vector<int>>* islands_out = nullptr) {
  const int n = H.n;
  iso = H.base_iso;
  status_out = -1;
  auto removed = [&](int br) {
    for (int q = 0; q < nout; ++q)
      if (out[q] == br) return true;
    return false;
  };
  // isolated endpoints
  for (int q = 0; q < nout; ++q) {
    const int br = out[q];
    if (!H.br_status[br]) continue;
    for (int u : {H.br_from[br], H.br_to[br]}) {
      int rem = 0;
      for (int p = H.adj_ptr[u]; p < H.adj_ptr[u + 1]; ++p)
        if (!removed(H.adj_br[p])) ++rem;
      if (rem == 0 && std::find(iso.begin(), iso.end(), u) == iso.end()) iso.push_back(u);
    }
  }
  std::sort(iso.begin(), iso.end());
  auto is_iso = [&](int u) { return std::binary_search(iso.begin(), iso.end(), u); };
  bool need_full = H.base_islands != 1;
  if (!need_full) {
    if ((int)ts.markA.size() != n) { ts.markA.assign(n, 0); ts.markB.assign(n, 0); }
    for (int q = 0; q < nout && !need_full; ++q) {
      const int br = out[q];
      if (!H.br_status[br]) continue;
      const int u = H.br_from[br], v = H.br_to[br];
      if (is_iso(u) || is_iso(v)) {
        // a leaf bus that lost its only branch cannot disconnect anything else; a bus that lost SEVERAL branches
        // may have been the only link between its neighbours: full component analysis
        if ((is_iso(u) && H.deg[u] > 1) || (is_iso(v) && H.deg[v] > 1)) need_full = true;
        continue;
      }
      // bidirectional BFS u <-> v in the graph minus the outage set
      const int sa = ++ts.stamp, sb = ++ts.stamp;
      ts.qa.assign(1, u); ts.qb.assign(1, v);
      ts.markA[u] = sa; ts.markB[v] = sb;
      size_t ia = 0, ib = 0;
      bool met = false;
      while (!met) {
        const bool canA = ia < ts.qa.size(), canB = ib < ts.qb.size();
        if (!canA || !canB) break;   // one side exhausted its component without meeting the other
        const bool expandA = (ts.qa.size() - ia) <= (ts.qb.size() - ib);
        std::vector<int>& qx = expandA ? ts.qa : ts.qb;
        size_t& ix = expandA ? ia : ib;
        std::vector<int>& mine = expandA ? ts.markA : ts.markB;
        std::vector<int>& other = expandA ? ts.markB : ts.markA;
        const int sm = expandA ? sa : sb, so = expandA ? sb : sa;
        const int x = qx[ix++];
        for (int p = H.adj_ptr[x]; p < H.adj_ptr[x + 1]; ++p) {
          if (removed(H.adj_br[p])) continue;
          const int y = H.adj_nb[p];
          if (other[y] == so) { met = true; break; }
          if (mine[y] != sm) { mine[y] = sm; qx.push_back(y); }
        }
      }
      if (!met) need_full = true;
    }
    if (!need_full) {
      // the slack must not be isolated
      if (is_iso(H.slack)) { status_out = SBLX_ST_ISLANDED_NO_REF; }
      return 1;
    }
  }
  // full component analysis
  ts.parent.resize(n);
  for (int i = 0; i < n; ++i) ts.parent[i] = i;
  auto find = [&](int i) {
    while (ts.parent[i] != i) { ts.parent[i] = ts.parent[ts.parent[i]]; i = ts.parent[i]; }
    return i;
  };
  // buses without any in-service branch in the base grid are isolated too
  std::vector<char> isov(n, 0);
  for (int u : iso) isov[u] = 1;
  for (int i = 0; i < n; ++i)
    if (H.deg[i] == 0) { isov[i] = 1; }
  for (int k = 0; k < H.nbr; ++k) {
    if (!H.br_status[k] || removed(k)) continue;
    int a = find(H.br_from[k]), b = find(H.br_to[k]);
    if (a != b) { if (a < b) ts.parent[b] = a; else ts.parent[a] = b; }
  }
  std::vector<char> has_ref(n, 0);
  int comps = 0;
  for (int i = 0; i < n; ++i) {
    if (isov[i]) continue;
    int r = find(i);
    if (r == i) ++comps;
    if (H.type0[i] == T_SLACK || H.type0[i] == T_PV) has_ref[r] = 1;
  }
  if (comps > 1) {
    bool all_ref = true;
    for (int i = 0; i < n; ++i)
      if (!isov[i] && find(i) == i && !has_ref[i]) all_ref = false;
    status_out = all_ref ? SBLX_ST_ISLANDED_UNSUPPORTED : SBLX_ST_ISLANDED_NO_REF;
    if (all_ref && islands_out) {
      // islands ordered by their smallest member (ac_islands.jl:88-92); the union-find root IS the smallest member
      std::vector<int> idx_of_root(n, -1);
      islands_out->clear();
      for (int i = 0; i < n; ++i) {
        if (isov[i]) continue;
        const int r = find(i);
        if (idx_of_root[r] < 0) { idx_of_root[r] = (int)islands_out->size(); islands_out->emplace_back(); }
        (*islands_out)[idx_of_root[r]].push_back(i);
      }
      status_out = -2;
    }
  } else if (isov[H.slack]) {
    status_out = SBLX_ST_ISLANDED_NO_REF;
  }
  return comps;
}

// ------------------------------------------------------------------------------------------
static void alloc_results(sblx_handle* h, int64_t B, const Want& want) {
  const bool want_v = want.v, want_types = want.types;
  h->r_conv.alloc(B); h->r_status.alloc(B); h->r_iter.alloc(B); h->r_nsw.alloc(B);
  h->r_fm.alloc(B); h->r_vmin.alloc(B); h->r_vmax.alloc(B); h->r_load.alloc(B);
  h->r_nvv.alloc(B); h->r_nov.alloc(B); h->r_tiny.alloc(B);
  if (want_v) h->r_V.alloc((size_t)B * h->H.n); else h->r_V.release();
  if (want_types) h->r_types.alloc((size_t)B * h->H.n); else h->r_types.release();
  if (want.flows) h->r_flows.alloc((size_t)B * h->H.nbr * 2); else h->r_flows.release();
  h->viol_cap = want.viol_cap; h->over_cap = want.over_cap;
  if (want.viol_cap > 0) h->r_viol_list.alloc(want.viol_cap); else h->r_viol_list.release();
  if (want.over_cap > 0) { h->r_over_list.alloc(want.over_cap); h->r_over_pct.alloc(want.over_cap); } else { h->r_over_list.release(); h->r_over_pct.release(); }
  h->qlog_cap = want.qlog_cap;
  if (want.qlog_cap > 0) h->r_qlog_list.alloc(want.qlog_cap); else h->r_qlog_list.release();
  h->r_list_counts.alloc(4);
  CK(cudaMemsetAsync(h->r_list_counts.p, 0, 4 * sizeof(unsigned long long), h->stream));
  CK(cudaMemsetAsync(h->r_nvv.p, 0, B * sizeof(int), h->stream));
  CK(cudaMemsetAsync(h->r_nov.p, 0, B * sizeof(int), h->stream));
  CK(cudaMemsetAsync(h->r_tiny.p, 0, B * sizeof(int), h->stream));
  if (want.flows) CK(cudaMemsetAsync(h->r_flows.p, 0, (size_t)B * h->H.nbr * 2 * sizeof(double2), h->stream));
  CK(cudaMemsetAsync(h->r_conv.p, 0, B * sizeof(int), h->stream));
  CK(cudaMemsetAsync(h->r_status.p, 0, B * sizeof(int), h->stream));
  CK(cudaMemsetAsync(h->r_iter.p, 0, B * sizeof(int), h->stream));
  CK(cudaMemsetAsync(h->r_nsw.p, 0, B * sizeof(int), h->stream));
  CK(cudaMemsetAsync(h->r_fm.p, 0xff, B * sizeof(double), h->stream));
  CK(cudaMemsetAsync(h->r_vmin.p, 0xff, B * sizeof(double), h->stream));
  CK(cudaMemsetAsync(h->r_vmax.p, 0xff, B * sizeof(double), h->stream));
  CK(cudaMemsetAsync(h->r_load.p, 0xff, B * sizeof(double), h->stream));
  if (want_v) CK(cudaMemsetAsync(h->r_V.p, 0, (size_t)B * h->H.n * sizeof(double2), h->stream));
  if (want_types) CK(cudaMemsetAsync(h->r_types.p, 0, (size_t)B * h->H.n, h->stream));
}

static void set_batch(sblx_handle* h) {
  Batch& Bt = h->Bt;
  Bt.B = (int)h->hostres.Bint; Bt.nq = (int)h->nq; Bt.use_sweep_q = h->sweep_q;
  Bt.Btop = h->hostres.sub_first.empty() ? Bt.B : (int)h->hostres.sub_first.size();
  Bt.queue = h->b_queue.p; Bt.out_ptr = h->b_out_ptr.p; Bt.out_br = h->b_out_br.p;
  Bt.iso_ptr = h->b_iso_ptr.p; Bt.iso_bus = h->b_iso_bus.p;
  Bt.ref_ptr = h->b_ref_ptr.p; Bt.ref_bus = h->b_ref_bus.p;
  Bt.lk_ptr = h->b_lk_ptr.p; Bt.lk_bus = h->b_lk_bus.p;
  Bt.Ssweep = h->b_S.p; Bt.qlsweep = h->b_ql.p; Bt.Vstart = h->b_Vstart.p;
  Bt.parent = h->b_parent.p;
  Bt.sweep_mode = h->sweep_mode; Bt.sweep_seed = h->sweep_seed; Bt.sweep_first = h->sweep_first;
  Bt.sweep_sigma = h->sweep_sigma; Bt.sweep_lo = h->sweep_lo; Bt.sweep_hi = h->sweep_hi; Bt.scale = h->b_scale.p;
  Bt.r_nvv = h->r_nvv.p; Bt.r_nov = h->r_nov.p; Bt.r_tiny = h->r_tiny.p;
  Bt.viol_list = h->r_viol_list.p; Bt.over_list = h->r_over_list.p; Bt.over_pct = h->r_over_pct.p;
  Bt.list_counts = h->r_list_counts.p; Bt.viol_cap = h->viol_cap; Bt.over_cap = h->over_cap;
  Bt.qlog_list = h->r_qlog_list.p; Bt.qlog_cap = h->qlog_cap;
  Bt.r_flows = h->r_flows.p;
  Bt.r_conv = h->r_conv.p; Bt.r_status = h->r_status.p; Bt.r_iter = h->r_iter.p; Bt.r_nsw = h->r_nsw.p;
  Bt.r_fm = h->r_fm.p; Bt.r_vmin = h->r_vmin.p; Bt.r_vmax = h->r_vmax.p; Bt.r_load = h->r_load.p;
  Bt.r_V = h->r_V.p; Bt.r_types = h->r_types.p;
  h->A.qload_s = h->sweep_q ? h->s_qload.p : nullptr;
}

// Host -> device copy of a large caller array through two pinned staging buffers: the caller's memory is pageable (Julia /
// numpy arrays), a direct cudaMemcpy from it is staged by the driver in small pieces; here a host thread's memcpy into
// pinned chunk k + 1 overlaps the DMA of chunk k.
static void h2d_staged(sblx_handle* h, void* dst, const void* src, size_t bytes) {
  const size_t CH = 32u << 20;
  if (bytes <= CH) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return;
  }
  if (!h->pinned_stage[0]) {
    CK(cudaMallocHost(&h->pinned_stage[0], CH));
    CK(cudaMallocHost(&h->pinned_stage[1], CH));
    CK(cudaEventCreateWithFlags(&h->stage_ev[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->stage_ev[1], cudaEventDisableTiming));
  }
  size_t off = 0;
  int k = 0;
  bool used[2] = {false, false};
  while (off < bytes) {
    const size_t nb = std::min(CH, bytes - off);
    if (used[k]) CK(cudaEventSynchronize(h->stage_ev[k]));          // the DMA that last read this buffer is done
    memcpy(h->pinned_stage[k], (const char*)src + off, nb);
    CK(cudaMemcpyAsync((char*)dst + off, h->pinned_stage[k], nb, cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(h->stage_ev[k], h->stream));
    used[k] = true;
    off += nb;
    k ^= 1;
  }
  CK(cudaStreamSynchronize(h->stream));
}

static void stage_outages(sblx_handle* h, int64_t B, const int32_t* optr, const int32_t* obr, const Want& want) {
  ensure_cuda(h);
  h->sweep_mode = 0;
  h->b_scale.release();
  const double t0 = now_ms();
  HostModel& H = h->H;
  const int base = h->opt.index_base;
  h->B = B;
  h->sweep_q = false;
  HostResults& R = h->hostres;
  R.status.assign(B, -1);
  R.islands.assign(B, 1);
  std::vector<int> out_ptr(B + 1, 0), out_br;
  std::vector<int> iso_ptr(1, 0), iso_bus;
  std::vector<std::vector<int>> iso_per(B);
  std::vector<std::vector<std::vector<int>>> isl_per(B);
  if (optr) {
    if (optr[0] != 0) throw std::invalid_argument("sblx: outage_ptr[0] must be 0");
    for (int64_t s = 0; s < B; ++s)
      if (optr[s + 1] < optr[s]) throw std::invalid_argument("sblx: outage_ptr must be non-decreasing");
    if (optr[B] > 0 && !obr) throw std::invalid_argument("sblx: outage_branch is NULL");
    out_br.reserve(optr[B]);
    for (int64_t s = 0; s < B; ++s) {
      const size_t b0 = out_br.size();
      for (int q = optr[s]; q < optr[s + 1]; ++q) {
        int br = obr[q] - base;
        if (br < 0 || br >= H.nbr) { R.status[s] = SBLX_ST_BAD_BRANCH; continue; }
        out_br.push_back(br);
      }
      // a branch can only be removed once (removeBranch! on an already removed branch changes nothing): the device
      // subtracts the four stamps once per list entry, so the set is made unique here
      std::sort(out_br.begin() + b0, out_br.end());
      out_br.erase(std::unique(out_br.begin() + b0, out_br.end()), out_br.end());
      out_ptr[s + 1] = (int)out_br.size();
      if (out_ptr[s + 1] - out_ptr[s] > MAXOUT) R.status[s] = SBLX_ST_BAD_BRANCH;
    }
  }
  // topology analysis, threaded over scenarios
  {
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    if (B < 256) nt = 1;
    std::vector<std::thread> th;
    std::atomic<int64_t> next(0);
    auto work = [&]() {
      TopoScratch ts;
      std::vector<int> iso;
      for (;;) {
        int64_t s0 = next.fetch_add(64);
        if (s0 >= B) break;
        for (int64_t s = s0; s < std::min<int64_t>(B, s0 + 64); ++s) {
          if (R.status[s] == SBLX_ST_BAD_BRANCH) { R.islands[s] = 0; continue; }
          int st;
          R.islands[s] = analyze_scenario(H, out_br.data() + out_ptr[s], out_ptr[s + 1] - out_ptr[s], ts, iso, st, &isl_per[s]);
          R.status[s] = st;
          iso_per[s] = iso;
        }
      }
    };
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
  }
  // device scenario table: the B caller scenarios, then one sub-scenario per island of every split scenario
  // (the other islands are masked as isolated rows; an island without the slack promotes its lowest PV bus)
  std::vector<int> queue;
  queue.reserve(B);
  std::vector<int> ref_ptr(1, 0), ref_bus;
  // lock_pv_to_pq_buses in an island-wise solve: the reference hands the list UNCHANGED to every island sub-net
  // (rectangular_network_solver.jl:1991), where the indices address the sub-net's own bus numbering (the island's
  // buses in ascending order, ac_islands.jl:329-337; ids beyond the island's size are ignored, limits.jl:434-438)
  std::vector<int> lk_ptr(1, 0), lk_bus, lock_ids;
  for (int b = 0; b < H.n; ++b)
    if (H.lockmask[b]) lock_ids.push_back(b);
  R.sub_first.assign(B, 0);
  R.sub_count.assign(B, 0);
  R.sub_buses.clear();
  std::vector<int> parent;
  for (int64_t s = 0; s < B; ++s) {
    for (int b : iso_per[s]) iso_bus.push_back(b);
    iso_ptr.push_back((int)iso_bus.size());
    ref_ptr.push_back((int)ref_bus.size());
    if (R.status[s] == -1) queue.push_back((int)s);
  }
  int64_t Bint = B;
  for (int64_t s = 0; s < B; ++s) {
    if (R.status[s] != -2) continue;
    R.sub_first[s] = (int)Bint;
    R.sub_count[s] = (int)isl_per[s].size();
    for (auto& isl : isl_per[s]) {
      std::vector<char> in(H.n, 0);
      bool has_slack = false;
      int promoted = -1;
      for (int b : isl) {
        in[b] = 1;
        if (b == H.slack) has_slack = true;
        if (promoted < 0 && H.type0[b] == T_PV) promoted = b;     // ascending bus order: lowest-index PV
      }
      for (int b = 0; b < H.n; ++b)
        if (!in[b]) iso_bus.push_back(b);
      iso_ptr.push_back((int)iso_bus.size());
      if (!has_slack) ref_bus.push_back(promoted);
      ref_ptr.push_back((int)ref_bus.size());
      for (int j : lock_ids)
        if (j < (int)isl.size()) lk_bus.push_back(isl[j]);
      lk_ptr.push_back((int)lk_bus.size());
      // same outage list as the parent scenario
      const int ob = out_ptr[s], oe = out_ptr[s + 1];
      for (int q = ob; q < oe; ++q) out_br.push_back(out_br[q]);
      out_ptr.push_back((int)out_br.size());
      queue.push_back((int)Bint);
      R.sub_buses.push_back(isl);
      parent.push_back((int)s);
      ++Bint;
    }
  }
  R.Bint = Bint;
  h->nq = (int64_t)queue.size();
  const double t1 = now_ms();
  h->stats = sblx_stats{};
  h->stats.host_prep_ms = t1 - t0;
  ensure_slots(h, h->nq);
  h->b_queue.upload(queue, h->stream);
  h->b_out_ptr.upload(out_ptr, h->stream);
  h->b_out_br.upload(out_br, h->stream);
  h->b_iso_ptr.upload(iso_ptr, h->stream);
  h->b_iso_bus.upload(iso_bus, h->stream);
  h->b_ref_ptr.upload(ref_ptr, h->stream);
  h->b_ref_bus.upload(ref_bus, h->stream);
  h->b_lk_ptr.upload(lk_ptr, h->stream);
  h->b_lk_bus.upload(lk_bus.empty() ? std::vector<int>(1, 0) : lk_bus, h->stream);
  h->b_parent.upload(parent.empty() ? std::vector<int>(1, 0) : parent, h->stream);
  R.parent = parent;
  h->b_host_status.upload(R.status, h->stream);
  h->b_sub_first.upload(R.sub_first, h->stream);
  h->b_sub_count.upload(R.sub_count, h->stream);
  h->b_islands.upload(R.islands, h->stream);
  h->b_S.release(); h->b_ql.release(); h->b_Vstart.release();
  alloc_results(h, Bint, want);
  // summary records of the B caller scenarios: allocated here (never inside the timed run), sized for the common shard
  // size of a sharded solve so that the all-gather can send it as is
  h->summary_cap = std::max<int64_t>(h->summary_cap, std::max<int64_t>(B, 1));
  if ((int64_t)h->r_summary.n < h->summary_cap) h->r_summary.alloc(h->summary_cap);
  CK(cudaMemsetAsync(h->r_summary.p, 0, h->r_summary.n * sizeof(SummaryRec), h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->stats.h2d_bytes = (int64_t)((queue.size() + out_ptr.size() + out_br.size() + iso_ptr.size() + iso_bus.size()) * sizeof(int));
  h->stats.h2d_ms = now_ms() - t1;
  set_batch(h);
  h->staged = true;
  h->ran = false;
}

template <int NT, int G = NT, bool WIDE = false>
static void launch_front(sblx_handle* h, const LaunchGroup& g, int nstep) {
  constexpr int NG = NT / G;                       // scenarios per CTA
  const int stride = (g.smem + 15) / 16;           // double2 per scenario group
  dim3 grid((nstep + NG - 1) / NG, g.cnt);
  if (g.pf_delta < 0) {
    g.pf_delta = 0;
    if (h->M.pf_mode & 3) {
      int occ = 1;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_front<NT, G, WIDE>, NT, (size_t)stride * 16 * NG));
      g.pf_delta = std::max(1, (int)(h->pf_mult * h->sm_count * std::max(occ, 1)));
    }
  }
  k_front<NT, G, WIDE><<<grid, NT, (size_t)stride * 16 * NG, h->stream>>>(h->M, h->A, h->Lover ? *h->Lover : h->L, h->d_panels_lv.p + g.off, stride, g.pf_delta);
}
template <int NT>
static void launch_solve(sblx_handle* h, const LaunchGroup& g, int nstep) {
  dim3 grid(nstep, g.cnt);
  k_backsolve<NT><<<grid, NT, g.smem, h->stream>>>(h->M, h->A, h->Lover ? *h->Lover : h->L, h->d_panels_lv.p + g.off);
}
static void launch_solve_group(sblx_handle* h, const LaunchGroup& g, int nstep) {
  if (g.lanes == 8) {
    k_backsolve_small<<<dim3((nstep + 3) / 4, g.cnt), 32, 0, h->stream>>>(h->M, h->A, h->Lover ? *h->Lover : h->L, h->d_panels_lv.p + g.off);
    return;
  }
  switch (g.nt) {
    case 32: launch_solve<32>(h, g, nstep); break;
    case 64: launch_solve<64>(h, g, nstep); break;
    default: launch_solve<128>(h, g, nstep); break;
  }
}

// numeric factorisation of the step set: level by level
static void launch_factor(sblx_handle* h, int nstep) {
  sblx_stats& st = h->stats;
  cudaStream_t s = h->stream;
  for (size_t l = 0; l < h->front_groups.size(); ++l) {
    if (h->opt.factor_mode != 1) {
      for (const LaunchGroup& g : h->front_groups[l]) {
        switch (g.nt) {
          case 32:
            if (g.reg) k_front_reg<8, 8><<<dim3((nstep + 3) / 4, g.cnt), 32, 0, h->stream>>>(h->M, h->A, h->Lover ? *h->Lover : h->L, h->d_panels_lv.p + g.off);
            else if (g.lanes == 4) launch_front<32, 4>(h, g, nstep);
            else if (g.lanes == 8) launch_front<32, 8>(h, g, nstep);
            else if (g.lanes == 16) launch_front<32, 16>(h, g, nstep);
            else launch_front<32>(h, g, nstep);
            break;
          case 64:
            if (g.wide) launch_front<64, 64, true>(h, g, nstep);
            else launch_front<64>(h, g, nstep);
            break;
          case 128: launch_front<128>(h, g, nstep); break;
          case 256: launch_front<256>(h, g, nstep); break;
          default: launch_front<512>(h, g, nstep); break;
        }
        st.launches++;
        st.factor_launches++;
      }
    } else {
      const LaunchGroup& ga = h->level_all[l];
      k_pivot<<<dim3(ga.cnt, (nstep + 3) / 4), dim3(32, 4), 0, s>>>(h->M, h->A, h->L, h->d_level_panels.p + ga.off);
      st.launches++;
      st.factor_launches++;
      for (const LaunchGroup& g : h->front_groups[l]) {
        if (g.subst_chunks == 0) continue;
        const int sm_sub = (2 * g.wmax * g.wmax + 2 * g.wmax) * 16;
        k_subst<<<dim3(g.cnt, nstep, g.subst_chunks), 128, sm_sub, s>>>(h->M, h->A, h->L, h->d_level_panels.p + g.off);
        st.launches++;
        st.factor_launches++;
      }
      for (const LaunchGroup& g : h->front_groups[l]) {
        if (g.schur_tiles == 0) continue;
        const int sm_sch = g.wmax * (2 * 32 + 2 * 64) * 16;
        k_schur<<<dim3(g.cnt, nstep, g.schur_tiles), 128, sm_sch, s>>>(h->M, h->A, h->L, h->d_level_panels.p + g.off);
        st.launches++;
        st.factor_launches++;
      }
    }
  }
}

// Small batches (everything resident from the first tick, at most 1024 scenarios): the per-tick host round trip and the
// launch latency of ~20 tiny kernels dominate (IEEE-14 N-1: 31 ticks).  Every kernel sizes its own work from the device
// counters and exits early beyond them, so a tick can be enqueued from an UPPER BOUND of the running scenarios alone; the
// host therefore stays two ticks ahead of the device and consumes the counters of tick k while ticks k+1, k+2 run.  Cost:
// at most two empty ticks after the last scenario finished.  Falls back to the synchronous loop (returns with the state in
// running / qhead) as soon as a tiny pivot was seen, because the refinement pass needs a round trip of its own.
// Per-phase timings are not taken here (total_ms is).
static void run_lookahead_prefix(sblx_handle* h, int64_t& running, int64_t& qhead, bool& saw_tiny) {
  cudaStream_t s = h->stream;
  const HostModel& H = h->H;
  const int n = H.n, m = H.m, nnzB = (int)H.jb_col.size();
  sblx_stats& st = h->stats;
  const int64_t nq = h->nq;
  constexpr int R = 4, D = 2;
  const int maxticks = h->opt.max_iter + 4;
  int ub = (int)nq, enq = 0, consumed = 0;
  bool stop = false;
  auto enqueue_tick = [&](int assign, int ubound) {
    const int u = std::max(ubound, 1);
    k_assign<<<1, 1024, 0, s>>>(h->M, h->A, h->L, h->Bt);
    st.launches++;
    if (assign > 0) {
      k_init<<<dim3((n + 7) / 8, (unsigned)((assign + 31) / 32)), dim3(32, 8), 0, s>>>(h->M, h->A, h->L, h->Bt);
      k_init_iso<<<(unsigned)((assign + 127) / 128), 128, 0, s>>>(h->M, h->A, h->L, h->Bt);
      st.launches += 2;
    }
    const dim3 gb((n + 7) / 8, (unsigned)((u + 31) / 32));
    k_mismatch<<<gb, dim3(32, 8), 0, s>>>(h->M, h->A, h->L);
    st.launches++;
    if (h->opt.qlimits_enabled) {
      k_qlimit<<<gb, dim3(32, 8), 0, s>>>(h->M, h->A, h->L, h->Bt);
      k_norm<<<u, 256, 0, s>>>(h->M, h->A, h->L);
      st.launches += 2;
    }
    k_decide<<<1, 1024, 0, s>>>(h->M, h->A, h->L);
    k_jacobian<<<dim3((nnzB + 31) / 32, (u + 31) / 32), dim3(32, 8), 0, s>>>(h->M, h->A, h->L);
    st.launches += 2;
    launch_factor(h, u);
    for (int l = (int)h->solve_groups.size() - 1; l >= 0; --l)
      for (const LaunchGroup& g : h->solve_groups[l]) {
        launch_solve_group(h, g, u);
        st.launches++;
      }
    if (h->M.pivot_tol > 0.0) {
      k_tiny_guard<<<u, 256, 0, s>>>(h->M, h->A, h->L);
      st.launches++;
    }
    k_update<<<dim3((m + 7) / 8, (u + 31) / 32), dim3(32, 8), 0, s>>>(h->M, h->A, h->L);
    k_poststep<<<(u + 127) / 128, 128, 0, s>>>(h->M, h->A, h->L);
    k_fin_init<<<(u + 127) / 128, 128, 0, s>>>(h->M, h->A, h->L);
    k_fin_bus<<<gb, dim3(32, 8), 0, s>>>(h->M, h->A, h->L, h->Bt);
    if (H.nbr > 0) {
      k_fin_branch<<<dim3((H.nbr + 7) / 8, (u + 31) / 32), dim3(32, 8), 0, s>>>(h->M, h->A, h->L, h->Bt);
      st.launches++;
    }
    k_fin_write<<<(u + 127) / 128, 128, 0, s>>>(h->M, h->A, h->L, h->Bt);
    st.launches += 5;
  };
  // In the look-ahead phase consecutive ticks are the same ~25 launches with the same grids as long as the upper bound
  // does not move (a straggler iterating to the cap): the tick is captured into a CUDA graph once per bound and replayed
  // with one cudaGraphLaunch instead of ~25 kernel launches.
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  int graph_ub = -1;
  int64_t graph_launches = 0, graph_factor_launches = 0;
  const bool use_graph = getenv("SBLX_NO_GRAPH") == nullptr;
  auto drop_graph = [&]() {
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
    gexec = nullptr; graph = nullptr; graph_ub = -1;
  };
  auto replay_tick = [&](int ubound) -> bool {
    if (ubound != graph_ub) {
      drop_graph();
      const int64_t l0 = st.launches, f0 = st.factor_launches;
      if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return false; }
      enqueue_tick(0, ubound);
      const cudaError_t e1 = cudaStreamEndCapture(s, &graph);
      graph_launches = st.launches - l0;
      graph_factor_launches = st.factor_launches - f0;
      st.launches = l0; st.factor_launches = f0;          // counted per replay below
      if (e1 != cudaSuccess || !graph || cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) { cudaGetLastError(); drop_graph(); return false; }
      graph_ub = ubound;
    }
    if (cudaGraphLaunch(gexec, s) != cudaSuccess) { cudaGetLastError(); drop_graph(); return false; }
    st.launches += graph_launches;
    st.factor_launches += graph_factor_launches;
    return true;
  };
  for (;;) {
    // the first ticks are taken one at a time (most scenarios are done after four or five iterations: running ahead
    // would only add empty ticks); the look-ahead pays on the tail of stragglers that iterate on
    while (!stop && enq - consumed <= (consumed >= 4 ? D : 0) && enq < maxticks) {
      if (!(use_graph && consumed >= 4 && enq > 0 && replay_tick(ub))) enqueue_tick(enq == 0 ? (int)nq : 0, ub);
      CK(cudaMemcpyAsync(h->pinned_counts + 8 * (1 + enq % R), h->l_counts.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
      CK(cudaEventRecord(h->ring_ev[enq % R], s));
      ++enq;
    }
    if (consumed == enq) break;
    CK(cudaEventSynchronize(h->ring_ev[consumed % R]));
    const int* c = h->pinned_counts + 8 * (1 + consumed % R);
    ++consumed;
    const int nactive = c[0], nstep = c[2], nfin = c[3], ndrain = c[5];
    if (nactive + ndrain > 0) {
      st.ticks++;
      st.mismatch_evaluations += nactive;
      st.scenario_iterations += nstep;
    }
    qhead = c[4];
    running = (int64_t)nactive + ndrain - nfin;
    ub = std::min<int>(ub, (int)running);
    if (c[7] > 0) { saw_tiny = true; stop = true; }                    // drain what is in flight, then the synchronous loop takes over
    if (running == 0 && qhead >= nq) stop = true;
  }
  drop_graph();
  CK(cudaGetLastError());
}

static void run_staged(sblx_handle* h) {
  if (!h->staged) throw std::invalid_argument("sblx_run_staged: nothing staged");
  ensure_cuda(h);
  cudaStream_t s = h->stream;
  const HostModel& H = h->H;
  const int W = h->W, n = H.n, m = H.m, nnzB = (int)H.jb_col.size();
  sblx_stats& st = h->stats;
  // per-run counters (stage-time fields host_prep_ms / h2d_* are kept)
  st.launches = st.factor_launches = st.ticks = st.scenario_iterations = st.mismatch_evaluations = st.scenarios_solved = 0;
  st.refine_passes = 0;
  st.total_ms = st.factor_ms = st.solve_ms = st.mismatch_ms = st.jacobian_ms = st.other_ms = 0.0;
  // reset slots
  CK(cudaMemsetAsync(h->s_state.p, 0, W * sizeof(int), s));
  CK(cudaMemsetAsync(h->s_scen.p, 0xff, W * sizeof(int), s));
  CK(cudaMemsetAsync(h->l_counts.p, 0, 8 * sizeof(int), s));
  cudaEvent_t* ev = h->ev;
  CK(cudaEventRecord(ev[0], s));
  int64_t running = 0, qhead = 0;
  const int64_t nq = h->nq;
  double t_mis = 0, t_jac = 0, t_fac = 0, t_sol = 0, t_oth = 0;
  bool have_prev = false, tiny_mode = (h->M.pf_mode & 2048) != 0;   // test hook: refinement armed from the first tick
  // SBLX_PROFILE_TICK=k brackets tick k (1-based) of this run with cudaProfilerStart/Stop (ncu --profile-from-start off)
  const char* ptick_env = getenv("SBLX_PROFILE_TICK");
  const long ptick = (ptick_env && nq > 1) ? atol(ptick_env) : 0;
  if (nq <= std::min<int64_t>(W, 1024) && h->opt.factor_mode != 1 && ptick == 0 && (h->M.pf_mode & 3072) == 0 && !getenv("SBLX_NO_LOOKAHEAD"))
    run_lookahead_prefix(h, running, qhead, tiny_mode);
  while (running > 0 || qhead < nq) {
    if (ptick > 0 && st.ticks + 1 == ptick) cudaProfilerStart();
    const int64_t assign = std::min<int64_t>(W - running, nq - qhead);
    const int nactive_ub = (int)(running + assign);   // exact: DRAIN slots are excluded on the device side
    CK(cudaEventRecord(ev[1], s));
    k_assign<<<1, 1024, 0, s>>>(h->M, h->A, h->L, h->Bt);
    st.launches++;
    if (assign > 0) {
      dim3 g((n + 7) / 8, (unsigned)((assign + 31) / 32));
      k_init<<<g, dim3(32, 8), 0, s>>>(h->M, h->A, h->L, h->Bt);
      k_init_iso<<<(unsigned)((assign + 127) / 128), 128, 0, s>>>(h->M, h->A, h->L, h->Bt);
      st.launches += 2;
    }
    {
      dim3 g((n + 7) / 8, (unsigned)((nactive_ub + 31) / 32));
      k_mismatch<<<g, dim3(32, 8), 0, s>>>(h->M, h->A, h->L);
      st.launches++;
      if (h->opt.qlimits_enabled) {
        k_qlimit<<<g, dim3(32, 8), 0, s>>>(h->M, h->A, h->L, h->Bt);
        k_norm<<<nactive_ub, 256, 0, s>>>(h->M, h->A, h->L);
        st.launches += 2;
      }
      k_decide<<<1, 1024, 0, s>>>(h->M, h->A, h->L);
      st.launches++;
    }
    CK(cudaEventRecord(ev[2], s));
    CK(cudaMemcpyAsync(h->pinned_counts, h->l_counts.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (have_prev) {
      float a = 0;
      CK(cudaEventElapsedTime(&a, ev[3], ev[4])); t_jac += a;
      CK(cudaEventElapsedTime(&a, ev[4], ev[5])); t_fac += a;
      CK(cudaEventElapsedTime(&a, ev[5], ev[6])); t_sol += a;
      CK(cudaEventElapsedTime(&a, ev[6], ev[7])); t_oth += a;
    }
    {
      float a = 0;
      CK(cudaEventElapsedTime(&a, ev[1], ev[2]));
      t_mis += a;
    }
    const int nactive = h->pinned_counts[0], nstep = h->pinned_counts[2], nfin = h->pinned_counts[3];
    if (h->pinned_counts[7] > 0) tiny_mode = true;
    qhead = h->pinned_counts[4];
    running += assign;
    st.ticks++;
    st.mismatch_evaluations += nactive;
    st.scenario_iterations += nstep;
    CK(cudaEventRecord(ev[3], s));
    if (nstep > 0) {
      dim3 g((nnzB + 31) / 32, (nstep + 31) / 32);
      k_jacobian<<<g, dim3(32, 8), 0, s>>>(h->M, h->A, h->L);
      st.launches++;
    }
    CK(cudaEventRecord(ev[4], s));
    if (nstep > 0) launch_factor(h, nstep);
    CK(cudaEventRecord(ev[5], s));
    if (nstep > 0) {
      for (int l = (int)h->solve_groups.size() - 1; l >= 0; --l)
        for (const LaunchGroup& g : h->solve_groups[l]) {
          launch_solve_group(h, g, nstep);
          st.launches++;
        }
    }
    CK(cudaEventRecord(ev[6], s));
    if (nstep > 0) {
      if (h->M.pivot_tol > 0.0) {
        k_tiny_guard<<<nstep, 256, 0, s>>>(h->M, h->A, h->L);
        st.launches++;
        if (tiny_mode && h->opt.factor_mode != 1) {
          // a tiny pivot has been met in this run: from now on the refine list is read back right after the guard (one
          // extra round trip per tick) and flagged steps are repaired before they are applied
          CK(cudaMemcpyAsync(h->pinned_counts, h->l_counts.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
          CK(cudaStreamSynchronize(s));
          const int nref = h->pinned_counts[6];
          if (nref > 0) {
            // ONE refinement pass: the factor / solve launches run over the refine list (counts[2] of the override
            // aliases counts[6]); a step that is still bad afterwards is declared singular
            Lists Lr = h->L;
            Lr.step = h->L.refine;
            Lr.counts = h->L.counts + 4;
            k_refine_prep<<<nref, 256, 0, s>>>(h->M, h->A, h->L);
            h->Lover = &Lr;
            launch_factor(h, nref);
            for (int l = (int)h->solve_groups.size() - 1; l >= 0; --l)
              for (const LaunchGroup& g : h->solve_groups[l]) launch_solve_group(h, g, nref);
            h->Lover = nullptr;
            k_refine_post<<<nref, 256, 0, s>>>(h->M, h->A, h->L, 1);
            st.refine_passes += nref;
          }
        }
      }
      dim3 g((m + 7) / 8, (nstep + 31) / 32);
      k_update<<<g, dim3(32, 8), 0, s>>>(h->M, h->A, h->L);
      k_poststep<<<(nstep + 127) / 128, 128, 0, s>>>(h->M, h->A, h->L);
      st.launches += 2;
    }
    if (nfin > 0) {
      k_fin_init<<<(nfin + 127) / 128, 128, 0, s>>>(h->M, h->A, h->L);
      dim3 gb((n + 7) / 8, (nfin + 31) / 32);
      k_fin_bus<<<gb, dim3(32, 8), 0, s>>>(h->M, h->A, h->L, h->Bt);
      if (H.nbr > 0) {
        dim3 gr((H.nbr + 7) / 8, (nfin + 31) / 32);
        k_fin_branch<<<gr, dim3(32, 8), 0, s>>>(h->M, h->A, h->L, h->Bt);
        st.launches++;
      }
      k_fin_write<<<(nfin + 127) / 128, 128, 0, s>>>(h->M, h->A, h->L, h->Bt);
      st.launches += 3;
      running -= nfin;
    }
    CK(cudaEventRecord(ev[7], s));
    if (ptick > 0 && st.ticks == ptick) { cudaStreamSynchronize(s); cudaProfilerStop(); }
    have_prev = true;
    CK(cudaGetLastError());
    if (st.ticks > (int64_t)(nq + 8) * (h->opt.max_iter + 2)) throw std::runtime_error("sblx: tick loop did not terminate");
  }
  if (h->B > 0) {
    k_pack_summary<<<(unsigned)((h->B + 255) / 256), 256, 0, s>>>(h->Bt, h->b_host_status.p, h->b_sub_first.p, h->b_sub_count.p,
                                                                  h->b_islands.p, h->r_summary.p);
    st.launches++;
  }
  CK(cudaEventRecord(ev[1], s));
  CK(cudaStreamSynchronize(s));
  if (have_prev) {
    float a = 0;
    CK(cudaEventElapsedTime(&a, ev[3], ev[4])); t_jac += a;
    CK(cudaEventElapsedTime(&a, ev[4], ev[5])); t_fac += a;
    CK(cudaEventElapsedTime(&a, ev[5], ev[6])); t_sol += a;
    CK(cudaEventElapsedTime(&a, ev[6], ev[7])); t_oth += a;
  }
  float tot = 0;
  CK(cudaEventElapsedTime(&tot, ev[0], ev[1]));
  st.total_ms = tot; st.mismatch_ms = t_mis; st.jacobian_ms = t_jac; st.factor_ms = t_fac; st.solve_ms = t_sol; st.other_ms = t_oth;
  st.scenarios_solved = nq;
  h->ran = true;
}

static Want want_from(const sblx_results* res) {
  Want w;
  w.v = res->V != nullptr;
  w.types = res->bus_type_final != nullptr;
  w.flows = res->flows != nullptr;
  w.viol_cap = (res->viol_scenario && res->viol_bus && res->viol_capacity > 0) ? res->viol_capacity : 0;
  w.over_cap = (res->over_scenario && res->over_branch && res->over_capacity > 0) ? res->over_capacity : 0;
  w.qlog_cap = (res->qlog_scenario && res->qlog_iter && res->qlog_bus && res->qlog_side && res->qlog_capacity > 0) ? res->qlog_capacity : 0;
  return w;
}

// summary arrays of `cnt` scenarios from their records, written at res[...][dst0 + k]
static void fill_summary(sblx_results* res, const SummaryRec* recs, int64_t cnt, int64_t dst0) {
  for (int64_t k = 0; k < cnt; ++k) {
    const SummaryRec& r = recs[k];
    const int64_t d = dst0 + k;
    if (res->converged) res->converged[d] = (uint8_t)r.conv;
    if (res->status) res->status[d] = r.status;
    if (res->iterations) res->iterations[d] = r.iter;
    if (res->n_switch_events) res->n_switch_events[d] = r.nsw;
    if (res->island_count) res->island_count[d] = r.islands;
    if (res->final_mismatch) res->final_mismatch[d] = r.fm;
    if (res->vmin_pu) res->vmin_pu[d] = r.vmin;
    if (res->vmax_pu) res->vmax_pu[d] = r.vmax;
    if (res->max_loading_pct) res->max_loading_pct[d] = r.load;
    if (res->n_voltage_violations) res->n_voltage_violations[d] = r.nvv;
    if (res->n_overloaded) res->n_overloaded[d] = r.nov;
    if (res->n_tiny_pivots) res->n_tiny_pivots[d] = r.tiny;
  }
}

// Copies the results of the staged batch back.  `dst0` = position of this batch's scenario 0 in the caller's arrays
// (a shard of a larger batch writes its rows at its offset); `summary` = also fill the per-scenario summary arrays
// from this handle's own records (a sharded solve fills them from the gathered records instead).
static void fetch_results(sblx_handle* h, sblx_results* res, int64_t dst0 = 0, bool summary = true) {
  if (!h->ran) throw std::invalid_argument("sblx_fetch_results: no completed run");
  const double t0 = now_ms();
  const int64_t B = h->B;
  const HostResults& R = h->hostres;
  const int64_t Bi = R.Bint, nsub = Bi - B;
  const int n = h->H.n, nbr = h->H.nbr, base = h->opt.index_base;
  cudaStream_t s = h->stream;
  std::vector<SummaryRec> recs(B);
  CK(cudaMemcpyAsync(recs.data(), h->r_summary.p, B * sizeof(SummaryRec), cudaMemcpyDeviceToHost, s));
  int64_t bytes = B * (int64_t)sizeof(SummaryRec);
  std::vector<double2> subV, subF;
  std::vector<unsigned char> subT;
  const bool wantV = res->V && h->r_V.p, wantT = res->bus_type_final && h->r_types.p, wantF = res->flows && h->r_flows.p;
  double2* Vout = wantV ? reinterpret_cast<double2*>(res->V) + (size_t)dst0 * n : nullptr;
  uint8_t* Tout = wantT ? res->bus_type_final + (size_t)dst0 * n : nullptr;
  double2* Fout = wantF ? reinterpret_cast<double2*>(res->flows) + (size_t)dst0 * nbr * 2 : nullptr;
  if (wantV) {
    CK(cudaMemcpyAsync(Vout, h->r_V.p, (size_t)B * n * 16, cudaMemcpyDeviceToHost, s));
    bytes += B * n * 16;
    if (nsub > 0) { subV.resize((size_t)nsub * n); CK(cudaMemcpyAsync(subV.data(), h->r_V.p + (size_t)B * n, (size_t)nsub * n * 16, cudaMemcpyDeviceToHost, s)); }
  }
  if (wantT) {
    CK(cudaMemcpyAsync(Tout, h->r_types.p, (size_t)B * n, cudaMemcpyDeviceToHost, s));
    bytes += B * n;
    if (nsub > 0) { subT.resize((size_t)nsub * n); CK(cudaMemcpyAsync(subT.data(), h->r_types.p + (size_t)B * n, (size_t)nsub * n, cudaMemcpyDeviceToHost, s)); }
  }
  if (wantF) {
    CK(cudaMemcpyAsync(Fout, h->r_flows.p, (size_t)B * nbr * 32, cudaMemcpyDeviceToHost, s));
    bytes += B * (int64_t)nbr * 32;
    if (nsub > 0) { subF.resize((size_t)nsub * nbr * 2); CK(cudaMemcpyAsync(subF.data(), h->r_flows.p + (size_t)B * nbr * 2, (size_t)nsub * nbr * 32, cudaMemcpyDeviceToHost, s)); }
  }
  // compacted violation / overload lists
  unsigned long long lc[4] = {0, 0, 0, 0};
  std::vector<int2> vl, ol;
  std::vector<int4> ql;
  std::vector<double> op;
  const bool wantVL = h->viol_cap > 0 && res->viol_scenario, wantOL = h->over_cap > 0 && res->over_scenario;
  const bool wantQL = h->qlog_cap > 0 && res->qlog_scenario;
  if (wantVL || wantOL || wantQL) {
    CK(cudaMemcpyAsync(lc, h->r_list_counts.p, sizeof lc, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (wantVL) {
      vl.resize((size_t)std::min<unsigned long long>(lc[0], (unsigned long long)h->viol_cap));
      if (!vl.empty()) CK(cudaMemcpyAsync(vl.data(), h->r_viol_list.p, vl.size() * sizeof(int2), cudaMemcpyDeviceToHost, s));
    }
    if (wantOL) {
      ol.resize((size_t)std::min<unsigned long long>(lc[1], (unsigned long long)h->over_cap));
      op.resize(ol.size());
      if (!ol.empty()) {
        CK(cudaMemcpyAsync(ol.data(), h->r_over_list.p, ol.size() * sizeof(int2), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(op.data(), h->r_over_pct.p, ol.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
      }
    }
    if (wantQL) {
      ql.resize((size_t)std::min<unsigned long long>(lc[2], (unsigned long long)h->qlog_cap));
      if (!ql.empty()) CK(cudaMemcpyAsync(ql.data(), h->r_qlog_list.p, ql.size() * sizeof(int4), cudaMemcpyDeviceToHost, s));
    }
    bytes += (int64_t)(vl.size() * 8 + ol.size() * 16 + ql.size() * 16);
  }
  CK(cudaStreamSynchronize(s));
  if (summary) fill_summary(res, recs.data(), B, dst0);
  int64_t tiny = 0;
  for (int64_t k = 0; k < B; ++k) {
    tiny += recs[k].tiny;
    if (R.status[k] != -2) continue;
    // island-wise composite: buses outside every island are isolated - they keep 1 angle 0 and type "isolated" like in
    // a single solve; flows of a split scenario are the union of its islands' (disjoint) branch sets
    if (wantV)
      for (int b = 0; b < n; ++b) Vout[(size_t)k * n + b] = make_double2(1.0, 0.0);
    if (wantT)
      for (int b = 0; b < n; ++b) Tout[(size_t)k * n + b] = (uint8_t)T_ISO;
    if (wantF)
      for (int b = 0; b < 2 * nbr; ++b) Fout[(size_t)k * nbr * 2 + b] = make_double2(0.0, 0.0);
    for (int q = 0; q < R.sub_count[k]; ++q) {
      const int64_t u = R.sub_first[k] + q - B;
      for (int b : R.sub_buses[u]) {
        if (wantV) Vout[(size_t)k * n + b] = subV[(size_t)u * n + b];
        if (wantT) Tout[(size_t)k * n + b] = subT[(size_t)u * n + b];
      }
      if (wantF && recs[k].conv)
        for (int b = 0; b < 2 * nbr; ++b) {
          const double2 f = subF[(size_t)u * nbr * 2 + b];
          if (f.x != 0.0 || f.y != 0.0) Fout[(size_t)k * nbr * 2 + b] = f;
        }
    }
  }
  h->stats.tiny_pivots = tiny;
  // lists: device scenario id -> caller scenario (island sub-scenarios map to their parent), entries of cases that did
  // not converge as a whole are dropped, order restored
  auto finish_list = [&](std::vector<int2>& lst, std::vector<double>* pct, unsigned long long devcount, int64_t cap, int32_t* oscen,
                         int32_t* oelem, double* opct, int64_t* ocount) {
    std::vector<size_t> keep;
    keep.reserve(lst.size());
    for (size_t i = 0; i < lst.size(); ++i) {
      int sc = lst[i].x;
      if (sc >= B) sc = (sc - B) < (int64_t)R.parent.size() ? R.parent[sc - B] : -1;   // island sub-scenario -> its caller scenario
      if (sc < 0 || !recs[sc].conv) continue;
      lst[i].x = sc;
      keep.push_back(i);
    }
    std::sort(keep.begin(), keep.end(), [&](size_t a, size_t b) { return lst[a].x != lst[b].x ? lst[a].x < lst[b].x : lst[a].y < lst[b].y; });
    const bool overflow = devcount > (unsigned long long)cap;
    for (size_t j = 0; j < keep.size() && (int64_t)j < cap; ++j) {
      oscen[j] = (int32_t)(lst[keep[j]].x + dst0);
      oelem[j] = lst[keep[j]].y + base;
      if (opct && pct) opct[j] = (*pct)[keep[j]];
    }
    if (ocount) *ocount = overflow ? (int64_t)devcount : (int64_t)keep.size();
  };
  if (wantVL) finish_list(vl, nullptr, lc[0], h->viol_cap, res->viol_scenario, res->viol_bus, nullptr, res->viol_count);
  if (wantOL) finish_list(ol, &op, lc[1], h->over_cap, res->over_scenario, res->over_branch, res->over_loading_pct, res->over_count);
  if (wantQL) {
    // the event log of EVERY solved scenario (converged or not), island sub-scenarios under their caller scenario;
    // order = the reference's push order: by iteration, buses ascending within an iteration (limits.jl:439-524)
    std::vector<int4> keep;
    keep.reserve(ql.size());
    for (int4 e : ql) {
      if (e.x >= B) e.x = (e.x - B) < (int64_t)R.parent.size() ? R.parent[e.x - B] : -1;
      if (e.x >= 0) keep.push_back(e);
    }
    std::stable_sort(keep.begin(), keep.end(), [](const int4& a, const int4& b) {
      return a.x != b.x ? a.x < b.x : a.y != b.y ? a.y < b.y : a.z < b.z;
    });
    for (size_t j = 0; j < keep.size() && (int64_t)j < h->qlog_cap; ++j) {
      res->qlog_scenario[j] = (int32_t)(keep[j].x + dst0);
      res->qlog_iter[j] = keep[j].y;
      res->qlog_bus[j] = keep[j].z + base;
      res->qlog_side[j] = keep[j].w;
    }
    if (res->qlog_count) *res->qlog_count = lc[2] > (unsigned long long)h->qlog_cap ? (int64_t)lc[2] : (int64_t)keep.size();
  }
  h->stats.d2h_ms = now_ms() - t0;
  h->stats.d2h_bytes = bytes;
}

}  // namespace sblx

// ==========================================================================================
// C ABI
// ==========================================================================================
#define API_BEGIN(h) \
  try {              \
    std::lock_guard<std::mutex> lk_((h)->mu);
#define API_END(h)                                                                    \
  return SBLX_OK;                                                                     \
  }                                                                                   \
  catch (const CudaError& e) { (h)->err = e.what(); return SBLX_E_CUDA; }             \
  catch (const std::invalid_argument& e) { (h)->err = e.what(); return SBLX_E_ARG; }  \
  catch (const std::bad_alloc&) { (h)->err = "out of host memory"; return SBLX_E_NOMEM; } \
  catch (const std::exception& e) { (h)->err = e.what(); return SBLX_E_INTERNAL; }

extern "C" {

int sblx_version(void) { return SBLX_VERSION; }

void sblx_default_options(sblx_options* o) {
  if (!o) return;
  memset(o, 0, sizeof *o);
  o->struct_size = (int32_t)sizeof(sblx_options);
  o->index_base = 1;
  o->max_iter = 30;
  o->qlimits_enabled = 1;
  o->qlimit_start_iter = 2;
  o->cooldown_iters = 0;
  o->guard_max_switches = 10;
  o->freeze_after_repeated_switching = 1;
  o->device = -1;
  o->max_slots = 0;
  o->ordering = 0;
  o->relax_small = 18;
  o->factor_mode = 0;
  o->tol = 1e-8;
  o->damp = 1.0;
  o->q_hyst_pu = 0.0;
  o->vm_min_pu = 0.9;
  o->vm_max_pu = 1.1;
  o->base_mva = 100.0;
  o->mem_fraction = 0.6;
  o->relax_zero_frac = 0.08;
  o->panel_smem_kb = 200.0;
  o->pivot_tol_rel = 1e-10;
}

const char* sblx_last_error(const sblx_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static void fill_host_model(HostModel& H, const sblx_options& o, int64_t n, const int64_t* colptr, const int64_t* rowval,
                            const double* nzval, int64_t slack_idx, int64_t nbranch, const int32_t* br_from,
                            const int32_t* br_to) {
  if (n < 2) throw std::invalid_argument("sblx_create: need at least 2 buses");
  if (!colptr || !rowval) throw std::invalid_argument("sblx_create: Ybus pattern missing");
  const int base = o.index_base;
  if (base != 0 && base != 1) throw std::invalid_argument("sblx_create: index_base must be 0 or 1");
  H.n = (int)n;
  H.slack = (int)(slack_idx - base);
  if (H.slack < 0 || H.slack >= n) throw std::invalid_argument("sblx_create: slack_idx out of range");
  H.nbr = (int)nbranch;
  build_pattern(H, n, colptr, rowval, nzval, base, br_from, br_to, nbranch);
  build_nodes_and_symbolic(H, o);
}

int sblx_create(sblx_handle** out, int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                int64_t slack_idx, const uint8_t* bus_type, const double* vset, const double* s_spec, const double* v0,
                const double* qmin_pu, const double* qmax_pu, const double* qload_pu, int64_t nbranch,
                const int32_t* br_from, const int32_t* br_to, const double* br_y11, const double* br_y12,
                const double* br_y21, const double* br_y22, const double* br_rating, const uint8_t* br_status,
                const sblx_options* opts) {
  if (!out) return SBLX_E_ARG;
  *out = nullptr;
  sblx_handle* h = nullptr;
  try {
    h = new sblx_handle();
    if (opts) {
      if (opts->struct_size != (int32_t)sizeof(sblx_options)) throw std::invalid_argument("sblx_create: options struct_size mismatch");
      h->opt = *opts;
    } else {
      sblx_default_options(&h->opt);
    }
    if (!nzval || !bus_type || !vset || !s_spec || !v0) throw std::invalid_argument("sblx_create: required array is NULL");
    if (nbranch > 0 && (!br_from || !br_to || !br_y11 || !br_y12 || !br_y21 || !br_y22))
      throw std::invalid_argument("sblx_create: branch table incomplete");
    if (h->opt.max_iter < 1 || !(h->opt.tol > 0) || !(h->opt.damp > 0 && h->opt.damp <= 1.0) || h->opt.qlimit_start_iter < 1)
      throw std::invalid_argument("sblx_create: invalid solver options");
    HostModel& H = h->H;
    fill_host_model(H, h->opt, n, colptr, rowval, nzval, slack_idx, nbranch, br_from, br_to);
    const int base = h->opt.index_base;
    H.S0.resize(n); H.V0.resize(n); H.vset.assign(vset, vset + n);
    H.qmin.resize(n); H.qmax.resize(n); H.qload.resize(n); H.type0.resize(n); H.lockmask.assign(n, 0);
    int nslack = 0;
    for (int64_t i = 0; i < n; ++i) {
      H.S0[i] = make_double2(s_spec[2 * i], s_spec[2 * i + 1]);
      H.V0[i] = make_double2(v0[2 * i], v0[2 * i + 1]);
      // limits.jl:211-225: NaN = no limit; the +Inf/-Inf sentinels of buses without generators stay non-finite
      double lo = qmin_pu ? qmin_pu[i] : -INFINITY, hi = qmax_pu ? qmax_pu[i] : INFINITY;
      if (std::isnan(lo)) lo = -INFINITY;
      if (std::isnan(hi)) hi = INFINITY;
      H.qmin[i] = lo; H.qmax[i] = hi;
      H.qload[i] = qload_pu ? qload_pu[i] : 0.0;
      uint8_t t = bus_type[i];
      if (t > 3) throw std::invalid_argument("sblx_create: unknown bus type");
      if (t == SBLX_BUS_SLACK) ++nslack;
      H.type0[i] = t == SBLX_BUS_ISOLATED ? (unsigned char)T_PQ : t;
    }
    if (nslack != 1 || bus_type[H.slack] != SBLX_BUS_SLACK)
      throw std::invalid_argument("sblx_create: exactly one Slack bus at slack_idx is required");
    if (nbranch == 0) {
      // PFModel without a branch table (solvePf of a bare model): connectivity comes from the Ybus
      // pattern as pseudo two-ports with zero stamps (they cannot be outaged and carry no rating)
      for (int i = 0; i < H.n; ++i)
        for (int p = H.y_rowptr[i]; p < H.y_rowptr[i + 1]; ++p) {
          const int j = H.y_col[p];
          if (j > i && (H.y_val[p].x != 0.0 || H.y_val[p].y != 0.0)) {
            H.br_from.push_back(i); H.br_to.push_back(j);
            H.y11.push_back(make_double2(0, 0)); H.y12.push_back(make_double2(0, 0));
            H.y21.push_back(make_double2(0, 0)); H.y22.push_back(make_double2(0, 0));
            H.rating.push_back(std::nan("")); H.br_status.push_back(1);
          }
        }
      H.nbr = (int)H.br_from.size();
    } else {
    H.br_from.resize(nbranch); H.br_to.resize(nbranch); H.y11.resize(nbranch); H.y12.resize(nbranch);
    H.y21.resize(nbranch); H.y22.resize(nbranch); H.rating.resize(nbranch); H.br_status.resize(nbranch);
    for (int64_t k = 0; k < nbranch; ++k) {
      H.br_from[k] = br_from[k] - base; H.br_to[k] = br_to[k] - base;
      H.y11[k] = make_double2(br_y11[2 * k], br_y11[2 * k + 1]); H.y12[k] = make_double2(br_y12[2 * k], br_y12[2 * k + 1]);
      H.y21[k] = make_double2(br_y21[2 * k], br_y21[2 * k + 1]); H.y22[k] = make_double2(br_y22[2 * k], br_y22[2 * k + 1]);
      H.rating[k] = br_rating ? br_rating[k] : std::nan("");
      H.br_status[k] = br_status ? (br_status[k] != 0) : 1;
    }
    }
    build_topology(H);
    {
      // islands of the intact grid
      TopoScratch ts;
      std::vector<int> iso;
      int st;
      H.base_islands = 0;   // force the full analysis
      H.base_islands = analyze_scenario(H, nullptr, 0, ts, iso, st);
    }
    *out = h;
    return SBLX_OK;
  } catch (const std::invalid_argument& e) {
    g_create_error = e.what();
    delete h;
    return SBLX_E_ARG;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    delete h;
    return SBLX_E_INTERNAL;
  }
}

void sblx_destroy(sblx_handle* h) { delete h; }

int sblx_set_locked_pv(sblx_handle* h, int64_t nlock, const int32_t* buses) {
  if (!h) return SBLX_E_ARG;
  API_BEGIN(h)
  std::fill(h->H.lockmask.begin(), h->H.lockmask.end(), 0);
  for (int64_t k = 0; k < nlock; ++k) {
    int b = buses[k] - h->opt.index_base;
    if (b >= 0 && b < h->H.n) h->H.lockmask[b] = 1;
  }
  if (h->cuda_ready) {
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpy(h->d_lock.p, h->H.lockmask.data(), h->H.n, cudaMemcpyHostToDevice));
  }
  API_END(h)
}

int sblx_set_v0(sblx_handle* h, const double* v0) {
  if (!h || !v0) return SBLX_E_ARG;
  API_BEGIN(h)
  for (int i = 0; i < h->H.n; ++i) h->H.V0[i] = make_double2(v0[2 * i], v0[2 * i + 1]);
  if (h->cuda_ready) {
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpy(h->d_V0.p, h->H.V0.data(), (size_t)h->H.n * 16, cudaMemcpyHostToDevice));
  }
  API_END(h)
}

static void fill_info(const sblx_handle* h, sblx_info* info) {
  const HostModel& H = h->H;
  const Symbolic& S = H.sym;
  memset(info, 0, sizeof *info);
  info->n = H.n; info->m = H.m; info->nnz_y = (int64_t)H.y_col.size();
  info->nnz_j_blocks = (int64_t)H.jb_col.size();
  info->nnz_j_scalar = 4 * info->nnz_j_blocks;
  info->nnz_lu_scalar = S.nnzLU_scalar;
  info->n_panels = (int64_t)S.panels.size(); info->n_levels = S.nlevels;
  info->u_blocks = S.u_blocks; info->c_blocks = S.c_blocks;
  info->max_slots = h->W; info->bytes_per_slot = (int64_t)slot_bytes(H);
  info->block_ops = S.block_ops;
  info->algorithmic_bytes_per_iteration = 16.0 * info->nnz_j_scalar + 16.0 * info->nnz_lu_scalar + 128.0 * H.n;
  snprintf(info->ordering, sizeof info->ordering, "%s", S.ordering_name.c_str());
}

int sblx_get_info(const sblx_handle* h, sblx_info* info) {
  if (!h || !info) return SBLX_E_ARG;
  fill_info(h, info);
  return SBLX_OK;
}

int sblx_get_stats(const sblx_handle* h, sblx_stats* s) {
  if (!h || !s) return SBLX_E_ARG;
  *s = h->stats;
  return SBLX_OK;
}

int sblx_stage_outages(sblx_handle* h, int64_t B, const int32_t* outage_ptr, const int32_t* outage_branch, int want_v, int want_types) {
  if (!h || B < 0) return SBLX_E_ARG;
  API_BEGIN(h)
  if (B > 0 && !outage_ptr) throw std::invalid_argument("sblx_stage_outages: outage_ptr is NULL");
  Want w;
  w.v = want_v != 0;
  w.types = want_types != 0;
  stage_outages(h, B, outage_ptr, outage_branch, w);
  API_END(h)
}

int sblx_run_staged(sblx_handle* h) {
  if (!h) return SBLX_E_ARG;
  API_BEGIN(h)
  run_staged(h);
  API_END(h)
}

int sblx_fetch_results(sblx_handle* h, sblx_results* res) {
  if (!h || !res) return SBLX_E_ARG;
  API_BEGIN(h)
  fetch_results(h, res);
  API_END(h)
}

int sblx_device_summary(sblx_handle* h, void** dev_ptr, int64_t* nbytes) {
  if (!h || !dev_ptr || !nbytes) return SBLX_E_ARG;
  API_BEGIN(h)
  if (!h->ran) throw std::invalid_argument("sblx_device_summary: no completed run");
  *dev_ptr = h->r_summary.p;
  *nbytes = h->B * (int64_t)sizeof(SummaryRec);
  API_END(h)
}

int sblx_solve_outages(sblx_handle* h, int64_t B, const int32_t* outage_ptr, const int32_t* outage_branch, sblx_results* res) {
  if (!h || !res || B < 0) return SBLX_E_ARG;
  API_BEGIN(h)
  if (B == 0) return SBLX_OK;
  if (!outage_ptr) throw std::invalid_argument("sblx_solve_outages: outage_ptr is NULL");
  stage_outages(h, B, outage_ptr, outage_branch, want_from(res));
  run_staged(h);
  fetch_results(h, res);
  API_END(h)
}

// ------------------------------------------------------------------------------------------
// multi-GPU: contiguous scenario shards + one ncclAllGather of the per-scenario records
// ------------------------------------------------------------------------------------------
static void shard_bounds(int64_t B, int nranks, std::vector<int64_t>& b) {
  // contiguous ranges in input order like Iterators.partition(eachindex(cases), ...) (contingency.jl:304), balanced
  // so that shard sizes differ by at most one
  b.resize(nranks + 1);
  for (int r = 0; r <= nranks; ++r) b[r] = B * r / nranks;
}

static void validate_outages(int64_t B, const int32_t* optr, const int32_t* obr) {
  if (B > 0 && !optr) throw std::invalid_argument("sblx: outage_ptr is NULL");
  if (B > 0 && optr[0] != 0) throw std::invalid_argument("sblx: outage_ptr[0] must be 0");
  for (int64_t s = 0; s < B; ++s)
    if (optr[s + 1] < optr[s]) throw std::invalid_argument("sblx: outage_ptr must be non-decreasing");
  if (B > 0 && optr[B] > 0 && !obr) throw std::invalid_argument("sblx: outage_branch is NULL");
}

// COLLECTIVE over the handle's communicator: solves this rank's shard of the batch, gathers the records of all ranks
// and fills the summary arrays of ALL B scenarios; bulk outputs (V, types, flows, lists) only for the own shard.
static void solve_sharded(sblx_handle* h, int64_t B, const int32_t* optr, const int32_t* obr, sblx_results* res, sblx_results* res_bulk) {
  ensure_cuda(h);
  const int nr = h->comm ? h->nranks : 1, rk = h->comm ? h->rank : 0;
  std::vector<int64_t> b;
  shard_bounds(B, nr, b);
  const int64_t lo = b[rk], hi = b[rk + 1], cnt = hi - lo;
  int64_t per_max = 1;
  for (int r = 0; r < nr; ++r) per_max = std::max(per_max, b[r + 1] - b[r]);
  std::vector<int32_t> lp(cnt + 1, 0);
  for (int64_t k = 0; k <= cnt; ++k) lp[k] = optr[lo + k] - optr[lo];
  const int32_t* lb = obr ? obr + optr[lo] : nullptr;
  h->summary_cap = std::max(h->summary_cap, per_max);
  sblx_results* rb = res_bulk ? res_bulk : res;
  if (cnt > 0) {
    stage_outages(h, cnt, lp.data(), lb, want_from(rb));
    run_staged(h);
  } else {   // more ranks than scenarios: this rank only takes part in the gather
    h->B = 0; h->staged = false; h->ran = false; h->stats = sblx_stats{};
    if ((int64_t)h->r_summary.n < h->summary_cap) h->r_summary.alloc(h->summary_cap);
    CK(cudaMemsetAsync(h->r_summary.p, 0, h->r_summary.n * sizeof(SummaryRec), h->stream));
  }
  cudaStream_t s = h->stream;
  std::vector<SummaryRec> all((size_t)per_max * nr);
  float ag_ms = 0.f;
  if (h->comm && nr > 1) {
    if ((int64_t)h->g_summary.n < per_max * nr) h->g_summary.alloc((size_t)per_max * nr);
    CK(cudaEventRecord(h->ev[0], s));
    NCK(nccl_api().AllGather(h->r_summary.p, h->g_summary.p, (size_t)per_max * sizeof(SummaryRec), ncclChar, h->comm, s));
    CK(cudaEventRecord(h->ev[1], s));
    CK(cudaMemcpyAsync(all.data(), h->g_summary.p, all.size() * sizeof(SummaryRec), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&ag_ms, h->ev[0], h->ev[1]));
  } else {
    CK(cudaMemcpyAsync(all.data(), h->r_summary.p, (size_t)per_max * sizeof(SummaryRec), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  if (res)
    for (int r = 0; r < nr; ++r) fill_summary(res, all.data() + (size_t)r * per_max, b[r + 1] - b[r], b[r]);
  if (cnt > 0) fetch_results(h, rb, lo, false);
  h->stats.n_ranks = nr; h->stats.shard_lo = lo; h->stats.shard_hi = hi; h->stats.allgather_ms = ag_ms;
  h->stats.d2h_bytes += (int64_t)(all.size() * sizeof(SummaryRec));
}

int sblx_comm_unique_id(uint8_t id128[128]) {
  if (!id128) return SBLX_E_ARG;
  try {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    if (nccl_api().GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return SBLX_E_CUDA; }
    memcpy(id128, &id, 128);
    return SBLX_OK;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    return SBLX_E_INTERNAL;
  }
}

int sblx_comm_init_rank(sblx_handle* h, int32_t nranks, int32_t rank, const uint8_t id128[128]) {
  if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return SBLX_E_ARG;
  API_BEGIN(h)
  ensure_cuda(h);
  if (h->comm && h->comm_owned) nccl_destroy(h->comm);
  h->comm = nullptr;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t c = nullptr;
  NCK(nccl_api().CommInitRank(&c, nranks, id, rank));
  h->comm = c; h->comm_owned = true; h->nranks = nranks; h->rank = rank;
  API_END(h)
}

int sblx_comm_destroy(sblx_handle* h) {
  if (!h) return SBLX_E_ARG;
  API_BEGIN(h)
  if (h->comm && h->comm_owned) { CK(cudaSetDevice(h->device)); nccl_destroy(h->comm); }
  h->comm = nullptr; h->nranks = 1; h->rank = 0;
  API_END(h)
}

int sblx_solve_outages_sharded(sblx_handle* h, int64_t B, const int32_t* outage_ptr, const int32_t* outage_branch, sblx_results* res) {
  if (!h || !res || B < 0) return SBLX_E_ARG;
  API_BEGIN(h)
  if (B == 0) return SBLX_OK;
  validate_outages(B, outage_ptr, outage_branch);
  solve_sharded(h, B, outage_ptr, outage_branch, res, nullptr);
  API_END(h)
}

int sblx_shard_bounds(const sblx_handle* h, int64_t B, const int32_t* outage_ptr, const int32_t* outage_branch, int32_t nranks,
                      int64_t* bounds) {
  (void)h; (void)outage_ptr; (void)outage_branch;
  if (B < 0 || nranks < 1 || !bounds) return SBLX_E_ARG;
  std::vector<int64_t> b;
  shard_bounds(B, nranks, b);
  for (int r = 0; r <= nranks; ++r) bounds[r] = b[r];
  return SBLX_OK;
}

}  // extern "C" (reopened below)

// one process, several GPUs: ndev engines + an ncclCommInitAll communicator, one host thread per GPU inside the call
struct sblx_multi {
  std::vector<sblx_handle*> hs;
  std::string err;
  std::mutex mu;
};

extern "C" {

int sblx_create_multi(sblx_multi** out, int32_t ndev, const int32_t* devices, int64_t n, const int64_t* colptr, const int64_t* rowval,
                      const double* nzval, int64_t slack_idx, const uint8_t* bus_type, const double* vset, const double* s_spec,
                      const double* v0, const double* qmin_pu, const double* qmax_pu, const double* qload_pu, int64_t nbranch,
                      const int32_t* br_from, const int32_t* br_to, const double* br_y11, const double* br_y12, const double* br_y21,
                      const double* br_y22, const double* br_rating, const uint8_t* br_status, const sblx_options* opts) {
  if (!out || ndev < 1) return SBLX_E_ARG;
  *out = nullptr;
  sblx_multi* m = new sblx_multi();
  try {
    std::vector<int> devs(ndev);
    for (int i = 0; i < ndev; ++i) devs[i] = devices ? devices[i] : i;
    for (int i = 0; i < ndev; ++i) {
      sblx_options o;
      if (opts) o = *opts; else sblx_default_options(&o);
      o.device = devs[i];
      sblx_handle* h = nullptr;
      const int rc = sblx_create(&h, n, colptr, rowval, nzval, slack_idx, bus_type, vset, s_spec, v0, qmin_pu, qmax_pu, qload_pu, nbranch,
                                 br_from, br_to, br_y11, br_y12, br_y21, br_y22, br_rating, br_status, &o);
      if (rc != SBLX_OK) { const std::string msg = g_create_error; for (auto* q : m->hs) delete q; delete m; g_create_error = msg; return rc; }
      m->hs.push_back(h);
      ensure_cuda(h);
    }
    if (ndev > 1) {
      std::vector<ncclComm_t> comms(ndev);
      NCK(nccl_api().CommInitAll(comms.data(), ndev, devs.data()));
      for (int i = 0; i < ndev; ++i) { m->hs[i]->comm = comms[i]; m->hs[i]->comm_owned = true; m->hs[i]->nranks = ndev; m->hs[i]->rank = i; }
    }
    *out = m;
    return SBLX_OK;
  } catch (const CudaError& e) {
    g_create_error = e.what();
    for (auto* q : m->hs) delete q;
    delete m;
    return SBLX_E_CUDA;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    for (auto* q : m->hs) delete q;
    delete m;
    return SBLX_E_INTERNAL;
  }
}

void sblx_destroy_multi(sblx_multi* m) {
  if (!m) return;
  for (auto* h : m->hs) delete h;
  delete m;
}

const char* sblx_multi_last_error(const sblx_multi* m) { return m ? m->err.c_str() : g_create_error.c_str(); }

sblx_handle* sblx_multi_handle(sblx_multi* m, int32_t i) { return (m && i >= 0 && i < (int)m->hs.size()) ? m->hs[i] : nullptr; }

int sblx_multi_set_v0(sblx_multi* m, const double* v0) {
  if (!m) return SBLX_E_ARG;
  for (auto* h : m->hs) {
    const int rc = sblx_set_v0(h, v0);
    if (rc != SBLX_OK) { m->err = h->err; return rc; }
  }
  return SBLX_OK;
}

int sblx_multi_set_locked_pv(sblx_multi* m, int64_t nlock, const int32_t* buses) {
  if (!m) return SBLX_E_ARG;
  for (auto* h : m->hs) {
    const int rc = sblx_set_locked_pv(h, nlock, buses);
    if (rc != SBLX_OK) { m->err = h->err; return rc; }
  }
  return SBLX_OK;
}

int sblx_multi_get_stats(const sblx_multi* m, int32_t i, sblx_stats* st) {
  if (!m || !st || i < 0 || i >= (int)m->hs.size()) return SBLX_E_ARG;
  *st = m->hs[i]->stats;
  return SBLX_OK;
}

int sblx_multi_solve_outages(sblx_multi* m, int64_t B, const int32_t* outage_ptr, const int32_t* outage_branch, sblx_results* res) {
  if (!m || !res || B < 0) return SBLX_E_ARG;
  std::lock_guard<std::mutex> lk(m->mu);
  if (B == 0) return SBLX_OK;
  try {
    validate_outages(B, outage_ptr, outage_branch);   // argument errors surface before any rank enters the collective
  } catch (const std::exception& e) {
    m->err = e.what();
    return SBLX_E_ARG;
  }
  const int nd = (int)m->hs.size();
  // per-rank list buffers (the ranks append concurrently; merged in rank = input order afterwards)
  struct RankLists { std::vector<int32_t> vs, vb, os, ob, qs, qi, qb, qd; std::vector<double> op; int64_t vc = 0, oc = 0, qc = 0; };
  std::vector<RankLists> rl(nd);
  std::vector<sblx_results> rr(nd, *res);
  const bool wantVL = res->viol_scenario && res->viol_bus && res->viol_capacity > 0;
  const bool wantOL = res->over_scenario && res->over_branch && res->over_capacity > 0;
  const bool wantQL = res->qlog_scenario && res->qlog_iter && res->qlog_bus && res->qlog_side && res->qlog_capacity > 0;
  for (int i = 0; i < nd; ++i) {
    if (wantQL) {
      rl[i].qs.resize(res->qlog_capacity); rl[i].qi.resize(res->qlog_capacity); rl[i].qb.resize(res->qlog_capacity); rl[i].qd.resize(res->qlog_capacity);
      rr[i].qlog_scenario = rl[i].qs.data(); rr[i].qlog_iter = rl[i].qi.data(); rr[i].qlog_bus = rl[i].qb.data(); rr[i].qlog_side = rl[i].qd.data();
      rr[i].qlog_count = &rl[i].qc;
    }
    if (wantVL) { rl[i].vs.resize(res->viol_capacity); rl[i].vb.resize(res->viol_capacity); rr[i].viol_scenario = rl[i].vs.data(); rr[i].viol_bus = rl[i].vb.data(); rr[i].viol_count = &rl[i].vc; }
    if (wantOL) {
      rl[i].os.resize(res->over_capacity); rl[i].ob.resize(res->over_capacity); rl[i].op.resize(res->over_capacity);
      rr[i].over_scenario = rl[i].os.data(); rr[i].over_branch = rl[i].ob.data(); rr[i].over_loading_pct = rl[i].op.data(); rr[i].over_count = &rl[i].oc;
    }
  }
  std::vector<int> rcs(nd, SBLX_OK);
  std::vector<std::string> errs(nd);
  auto work = [&](int i) {
    sblx_handle* h = m->hs[i];
    std::lock_guard<std::mutex> hl(h->mu);
    try {
      CK(cudaSetDevice(h->device));
      solve_sharded(h, B, outage_ptr, outage_branch, i == 0 ? res : nullptr, &rr[i]);
    } catch (const CudaError& e) { errs[i] = e.what(); rcs[i] = SBLX_E_CUDA; }
    catch (const std::invalid_argument& e) { errs[i] = e.what(); rcs[i] = SBLX_E_ARG; }
    catch (const std::exception& e) { errs[i] = e.what(); rcs[i] = SBLX_E_INTERNAL; }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nd; ++i) th.emplace_back(work, i);
  work(0);
  for (auto& t : th) t.join();
  for (int i = 0; i < nd; ++i)
    if (rcs[i] != SBLX_OK) { m->err = errs[i]; return rcs[i]; }
  if (wantVL) {
    int64_t tot = 0, w = 0;
    for (int i = 0; i < nd; ++i) {
      const int64_t have = std::min<int64_t>(rl[i].vc, res->viol_capacity);
      for (int64_t j = 0; j < have && w < res->viol_capacity; ++j, ++w) { res->viol_scenario[w] = rl[i].vs[j]; res->viol_bus[w] = rl[i].vb[j]; }
      tot += rl[i].vc;
    }
    if (res->viol_count) *res->viol_count = tot;
  }
  if (wantOL) {
    int64_t tot = 0, w = 0;
    for (int i = 0; i < nd; ++i) {
      const int64_t have = std::min<int64_t>(rl[i].oc, res->over_capacity);
      for (int64_t j = 0; j < have && w < res->over_capacity; ++j, ++w) {
        res->over_scenario[w] = rl[i].os[j]; res->over_branch[w] = rl[i].ob[j];
        if (res->over_loading_pct) res->over_loading_pct[w] = rl[i].op[j];
      }
      tot += rl[i].oc;
    }
    if (res->over_count) *res->over_count = tot;
  }
  if (wantQL) {
    int64_t tot = 0, w = 0;
    for (int i = 0; i < nd; ++i) {
      const int64_t have = std::min<int64_t>(rl[i].qc, res->qlog_capacity);
      for (int64_t j = 0; j < have && w < res->qlog_capacity; ++j, ++w) {
        res->qlog_scenario[w] = rl[i].qs[j]; res->qlog_iter[w] = rl[i].qi[j]; res->qlog_bus[w] = rl[i].qb[j]; res->qlog_side[w] = rl[i].qd[j];
      }
      tot += rl[i].qc;
    }
    if (res->qlog_count) *res->qlog_count = tot;
  }
  return SBLX_OK;
}

// One chunk of an explicit injection sweep: scenarios [b0, b0 + nb) of the caller's matrices, results written at row b0.
static void solve_injections_chunk(sblx_handle* h, int64_t b0, int64_t nb, const double* s_spec, const double* qload_pu, sblx_results* res) {
  const int n = h->H.n;
  stage_outages(h, nb, nullptr, nullptr, want_from(res));
  const size_t cnt = (size_t)nb * n;
  const double t0 = now_ms();
  h->b_S.alloc(cnt);
  h2d_staged(h, h->b_S.p, s_spec + (size_t)b0 * n * 2, cnt * 16);
  h->stats.h2d_bytes += (int64_t)cnt * 16;
  if (qload_pu) {
    h->b_ql.alloc(cnt);
    h2d_staged(h, h->b_ql.p, qload_pu + (size_t)b0 * n, cnt * 8);
    h->stats.h2d_bytes += (int64_t)cnt * 8;
    h->sweep_q = true;
  }
  h->stats.h2d_ms += now_ms() - t0;
  set_batch(h);
  run_staged(h);
  fetch_results(h, res, b0);
}

int sblx_solve_injections(sblx_handle* h, int64_t B, const double* s_spec, const double* qload_pu, sblx_results* res) {
  if (!h || !res || B < 0) return SBLX_E_ARG;
  API_BEGIN(h)
  if (B == 0) return SBLX_OK;
  if (!s_spec) throw std::invalid_argument("sblx_solve_injections: s_spec is NULL");
  // The per-scenario injection matrices are streamed through the device in chunks of scenarios, so a large sweep never
  // needs B x n resident (16 + 8 bytes per bus and scenario: 2.4 GB per 10 000 scenarios of a 10k-bus grid).
  double budget_gb = 4.0;
  if (const char* e = getenv("SBLX_SWEEP_CHUNK_GB")) budget_gb = std::max(1e-6, atof(e));
  const int64_t per = (int64_t)h->H.n * (qload_pu ? 24 : 16);
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(B, (int64_t)(budget_gb * 1e9) / per));
  if (chunk >= B) {
    solve_injections_chunk(h, 0, B, s_spec, qload_pu, res);
  } else {
    // lists: every chunk appends behind the previous one; counts and the run statistics accumulate
    sblx_results rc = *res;
    int64_t vtot = 0, otot = 0, qtot = 0, vw = 0, ow = 0, qw = 0, cv = 0, co = 0, cq = 0;
    sblx_stats acc{};
    for (int64_t b0 = 0; b0 < B; b0 += chunk) {
      const int64_t nb = std::min(chunk, B - b0);
      if (res->viol_scenario) { rc.viol_scenario = res->viol_scenario + vw; rc.viol_bus = res->viol_bus + vw; rc.viol_capacity = res->viol_capacity - vw; rc.viol_count = &cv; }
      if (res->over_scenario) {
        rc.over_scenario = res->over_scenario + ow; rc.over_branch = res->over_branch + ow; rc.over_capacity = res->over_capacity - ow; rc.over_count = &co;
        rc.over_loading_pct = res->over_loading_pct ? res->over_loading_pct + ow : nullptr;
      }
      if (res->qlog_scenario) {
        rc.qlog_scenario = res->qlog_scenario + qw; rc.qlog_iter = res->qlog_iter + qw; rc.qlog_bus = res->qlog_bus + qw; rc.qlog_side = res->qlog_side + qw;
        rc.qlog_capacity = res->qlog_capacity - qw; rc.qlog_count = &cq;
      }
      cv = co = cq = 0;
      solve_injections_chunk(h, b0, nb, s_spec, qload_pu, &rc);
      vtot += cv; otot += co; qtot += cq;
      vw = std::min<int64_t>(res->viol_capacity, vw + cv); ow = std::min<int64_t>(res->over_capacity, ow + co); qw = std::min<int64_t>(res->qlog_capacity, qw + cq);
      const sblx_stats& st = h->stats;
      acc.total_ms += st.total_ms; acc.factor_ms += st.factor_ms; acc.solve_ms += st.solve_ms; acc.mismatch_ms += st.mismatch_ms;
      acc.jacobian_ms += st.jacobian_ms; acc.other_ms += st.other_ms; acc.h2d_ms += st.h2d_ms; acc.d2h_ms += st.d2h_ms;
      acc.host_prep_ms += st.host_prep_ms; acc.launches += st.launches; acc.factor_launches += st.factor_launches; acc.ticks += st.ticks;
      acc.scenario_iterations += st.scenario_iterations; acc.mismatch_evaluations += st.mismatch_evaluations;
      acc.scenarios_solved += st.scenarios_solved; acc.h2d_bytes += st.h2d_bytes; acc.d2h_bytes += st.d2h_bytes;
      acc.tiny_pivots += st.tiny_pivots; acc.refine_passes += st.refine_passes;
    }
    if (res->viol_count) *res->viol_count = vtot;
    if (res->over_count) *res->over_count = otot;
    if (res->qlog_count) *res->qlog_count = qtot;
    acc.n_ranks = 1; acc.shard_hi = B;
    h->stats = acc;
  }
  API_END(h)
}

static void solve_sweep(sblx_handle* h, int64_t B, int mode, int64_t first, uint64_t seed, double sigma, double lo, double hi,
                        const double* scale, sblx_results* res) {
  stage_outages(h, B, nullptr, nullptr, want_from(res));
  h->sweep_mode = mode; h->sweep_seed = seed; h->sweep_first = first; h->sweep_sigma = sigma; h->sweep_lo = lo; h->sweep_hi = hi;
  h->sweep_q = true;                       // the load's reactive part scales with it (the Q-limit bookkeeping reads qload)
  if (mode == 2) {
    h->b_scale.alloc(B);
    CK(cudaMemcpyAsync(h->b_scale.p, scale, B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->stats.h2d_bytes += B * (int64_t)sizeof(double);
  }
  set_batch(h);
  run_staged(h);
  fetch_results(h, res);
}

int sblx_solve_injection_sweep(sblx_handle* h, int64_t B, int64_t first, uint64_t seed, double sigma, double lo, double hi,
                               sblx_results* res) {
  if (!h || !res || B < 0) return SBLX_E_ARG;
  API_BEGIN(h)
  if (B == 0) return SBLX_OK;
  if (!(sigma >= 0.0) || !(lo <= hi)) throw std::invalid_argument("sblx_solve_injection_sweep: need sigma >= 0 and lo <= hi");
  solve_sweep(h, B, 1, first, seed, sigma, lo, hi, nullptr, res);
  API_END(h)
}

int sblx_solve_injection_scale(sblx_handle* h, int64_t B, const double* scale, sblx_results* res) {
  if (!h || !res || B < 0) return SBLX_E_ARG;
  API_BEGIN(h)
  if (B == 0) return SBLX_OK;
  if (!scale) throw std::invalid_argument("sblx_solve_injection_scale: scale is NULL");
  solve_sweep(h, B, 2, 0, 0, 0.0, 0.0, 0.0, scale, res);
  API_END(h)
}

int sblx_sweep_factors(sblx_handle* h, int64_t B, int64_t first, uint64_t seed, double sigma, double lo, double hi, double* out) {
  if (!h || !out || B < 0) return SBLX_E_ARG;
  API_BEGIN(h)
  ensure_cuda(h);
  const int n = h->H.n;
  const int64_t chunk = std::max<int64_t>(1, (64ll << 20) / ((int64_t)n * 8));
  DevBuf<double> buf;
  buf.alloc((size_t)std::min(chunk, std::max<int64_t>(B, 1)) * n);
  for (int64_t b0 = 0; b0 < B; b0 += chunk) {
    const int64_t nb = std::min(chunk, B - b0);
    const long long tot = (long long)nb * n;
    k_sweep_factors<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(h->M, first + b0, (int)nb, seed, sigma, lo, hi, buf.p);
    CK(cudaMemcpyAsync(out + (size_t)b0 * n, buf.p, (size_t)tot * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  API_END(h)
}

int sblx_solve_one(sblx_handle* h, const double* v_start, double* v_out, uint8_t* converged, int32_t* iters,
                   double* residual_inf, int32_t* status) {
  if (!h) return SBLX_E_ARG;
  API_BEGIN(h)
  int32_t optr[2] = {0, 0};
  Want w1;
  w1.v = v_out != nullptr;
  stage_outages(h, 1, optr, nullptr, w1);
  if (v_start) {
    std::vector<double2> vs(h->H.n);
    for (int i = 0; i < h->H.n; ++i) vs[i] = make_double2(v_start[2 * i], v_start[2 * i + 1]);
    h->b_Vstart.upload(vs, h->stream);
    CK(cudaStreamSynchronize(h->stream));
    set_batch(h);
  }
  run_staged(h);
  sblx_results r{};
  uint8_t c = 0;
  int32_t it = 0, stt = 0;
  double fm = 0;
  r.converged = &c; r.iterations = &it; r.status = &stt; r.final_mismatch = &fm; r.V = v_out;
  fetch_results(h, &r);
  if (converged) *converged = c;
  if (iters) *iters = it;
  if (residual_inf) *residual_inf = fm;
  if (status) *status = stt;
  API_END(h)
}

int sblx_debug_set_pf_mode(sblx_handle* h, int32_t mode) {
  if (!h) return SBLX_E_ARG;
  API_BEGIN(h)
  h->M.pf_mode = mode;
  for (auto& lv : h->front_groups)
    for (auto& g : lv) g.pf_delta = -1;
  API_END(h)
}

int sblx_debug_phase_cycles(double* out50) {
#ifdef SBLX_PHASE_TIMING
  unsigned long long hbuf[50];
  if (cudaMemcpyFromSymbol(hbuf, g_phase_cycles, sizeof hbuf) != cudaSuccess) return SBLX_E_CUDA;
  for (int i = 0; i < 50; ++i) out50[i] = (double)hbuf[i];
  return SBLX_OK;
#else
  (void)out50;
  return SBLX_E_ARG;
#endif
}

int sblx_debug_pattern(const sblx_handle* h, int64_t* rowptr, int32_t* col) {
  if (!h) return SBLX_E_ARG;
  const HostModel& H = h->H;
  if (rowptr) for (int a = 0; a <= H.m; ++a) rowptr[a] = H.jb_rowptr[a];
  if (col) for (size_t k = 0; k < H.jb_col.size(); ++k) col[k] = H.bus_of_node[H.jb_col[k]];
  return SBLX_OK;
}

int sblx_debug_newton_step(sblx_handle* h, int64_t nout, const int32_t* outage_branch, const uint8_t* bus_type,
                           const double* v_in, double* F, double* jblocks, double* dx, double* max_mismatch) {
  if (!h || !v_in) return SBLX_E_ARG;
  API_BEGIN(h)
  const HostModel& H = h->H;
  const int n = H.n, m = H.m, nnzB = (int)H.jb_col.size();
  int32_t optr[2] = {0, (int32_t)nout};
  stage_outages(h, 1, optr, outage_branch, Want{});
  if (h->nq != 1) throw std::invalid_argument("sblx_debug_newton_step: scenario is islanded / invalid");
  std::vector<double2> vs(n);
  for (int i = 0; i < n; ++i) vs[i] = make_double2(v_in[2 * i], v_in[2 * i + 1]);
  h->b_Vstart.upload(vs, h->stream);
  set_batch(h);
  cudaStream_t s = h->stream;
  CK(cudaMemsetAsync(h->s_state.p, 0, h->W * sizeof(int), s));
  CK(cudaMemsetAsync(h->s_scen.p, 0xff, h->W * sizeof(int), s));
  CK(cudaMemsetAsync(h->l_counts.p, 0, 8 * sizeof(int), s));
  k_assign<<<1, 1024, 0, s>>>(h->M, h->A, h->L, h->Bt);
  k_init<<<dim3((n + 7) / 8, 1), dim3(32, 8), 0, s>>>(h->M, h->A, h->L, h->Bt);
  k_init_iso<<<1, 128, 0, s>>>(h->M, h->A, h->L, h->Bt);
  if (bus_type) {
    // override the active set of slot 0 (interleaved stride W)
    std::vector<unsigned char> t(n);
    for (int i = 0; i < n; ++i) t[i] = bus_type[i] == SBLX_BUS_ISOLATED ? (unsigned char)T_PQ : bus_type[i];
    CK(cudaMemcpy2DAsync(h->s_type.p, h->W, t.data(), 1, 1, n, cudaMemcpyHostToDevice, s));
  }
  k_mismatch<<<dim3((n + 7) / 8, 1), dim3(32, 8), 0, s>>>(h->M, h->A, h->L);
  // force a step list of {slot 0}
  int one[8] = {1, 0, 1, 0, 1, 0, 0, 0};
  int zero = 0;
  CK(cudaMemcpyAsync(h->l_counts.p, one, sizeof one, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(h->l_step.p, &zero, sizeof(int), cudaMemcpyHostToDevice, s));
  k_jacobian<<<dim3((nnzB + 31) / 32, 1), dim3(32, 8), 0, s>>>(h->M, h->A, h->L);
  launch_factor(h, 1);
  for (int l = (int)h->solve_groups.size() - 1; l >= 0; --l)
    for (const LaunchGroup& g : h->solve_groups[l]) {
      launch_solve_group(h, g, 1);
    }
  CK(cudaGetLastError());
  std::vector<double2> rhs(m), xv(m);
  std::vector<double> J((size_t)nnzB * 4);
  unsigned long long mb = 0;
  CK(cudaMemcpyAsync(rhs.data(), h->s_rhs.p, m * 16, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(xv.data(), h->s_xv.p, m * 16, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(J.data(), h->s_Jv.p, (size_t)nnzB * 32, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&mb, h->s_mis.p, 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const Symbolic& S = H.sym;
  for (int a = 0; a < m; ++a) {
    const int e = S.iperm[a];
    if (F) { F[2 * a] = -rhs[e].x; F[2 * a + 1] = -rhs[e].y; }
    if (dx) { dx[2 * a] = xv[e].x; dx[2 * a + 1] = xv[e].y; }
  }
  if (jblocks) memcpy(jblocks, J.data(), J.size() * sizeof(double));
  if (max_mismatch) memcpy(max_mismatch, &mb, 8);
  h->staged = false;
  API_END(h)
}

int sblx_debug_host_selftest(const sblx_handle* h, uint64_t seed, double* rel_residual) {
  if (!h || !rel_residual) return SBLX_E_ARG;
  try {
    *rel_residual = host_selftest(h->H, seed);
    return SBLX_OK;
  } catch (const std::exception&) {
    return SBLX_E_INTERNAL;
  }
}

int sblx_debug_topology(int64_t n, int64_t slack_idx, const uint8_t* bus_type, int64_t nbranch, const int32_t* br_from,
                        const int32_t* br_to, const uint8_t* br_status, int32_t index_base, int64_t B, const int32_t* outage_ptr,
                        const int32_t* outage_branch, int32_t* island_count, int32_t* status, int32_t* n_isolated) {
  // host-only: the topology analysis stage_outages runs per scenario (isolated buses, islands, reference check)
  if (n <= 0 || nbranch < 0 || !bus_type || !br_from || !br_to || B < 0 || (B > 0 && !outage_ptr)) return SBLX_E_ARG;
  try {
    HostModel H;
    H.n = (int)n; H.nbr = (int)nbranch; H.slack = (int)(slack_idx - index_base);
    H.type0.assign(bus_type, bus_type + n);
    H.br_from.resize(nbranch); H.br_to.resize(nbranch); H.br_status.resize(nbranch);
    for (int64_t k = 0; k < nbranch; ++k) {
      H.br_from[k] = br_from[k] - index_base; H.br_to[k] = br_to[k] - index_base;
      H.br_status[k] = br_status ? (br_status[k] != 0) : 1;
      if (H.br_from[k] < 0 || H.br_from[k] >= n || H.br_to[k] < 0 || H.br_to[k] >= n) return SBLX_E_ARG;
    }
    build_topology(H);
    TopoScratch ts;
    std::vector<int> iso;
    int st;
    H.base_islands = 0;
    H.base_islands = analyze_scenario(H, nullptr, 0, ts, iso, st);
    std::vector<int> out;
    for (int64_t s = 0; s < B; ++s) {
      out.clear();
      bool bad = false;
      for (int q = outage_ptr[s]; q < outage_ptr[s + 1]; ++q) {
        const int br = outage_branch[q] - index_base;
        if (br < 0 || br >= nbranch) bad = true; else out.push_back(br);
      }
      std::vector<std::vector<int>> islands;
      int cnt = 0;
      st = SBLX_ST_BAD_BRANCH;
      if (!bad) cnt = analyze_scenario(H, out.data(), (int)out.size(), ts, iso, st, &islands);
      if (island_count) island_count[s] = bad ? 0 : cnt;
      if (status) status[s] = st;
      if (n_isolated) n_isolated[s] = bad ? 0 : (int)iso.size();
    }
    return SBLX_OK;
  } catch (...) {
    return SBLX_E_INTERNAL;
  }
}

int sblx_analyze_only(int64_t n, const int64_t* colptr, const int64_t* rowval, int64_t slack_idx, const sblx_options* opts,
                      sblx_info* info, double* selftest_residual) {
  try {
    sblx_handle h;
    if (opts) h.opt = *opts; else sblx_default_options(&h.opt);
    fill_host_model(h.H, h.opt, n, colptr, rowval, nullptr, slack_idx, 0, nullptr, nullptr);
    if (info) fill_info(&h, info);
    if (selftest_residual) *selftest_residual = host_selftest(h.H, 12345);
    return SBLX_OK;
  } catch (const std::invalid_argument& e) {
    g_create_error = e.what();
    return SBLX_E_ARG;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    return SBLX_E_INTERNAL;
  }
}

}  // extern "C"
