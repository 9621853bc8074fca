// sm_100a CUDA kernels of the batched rectangular Newton-Raphson engine.
//
// Layout in HBM (W = resident scenario slots):
//   * per-bus state is SCENARIO-INTERLEAVED  [bus][slot]  (V, I, S, type, Q-limit bookkeeping):
//     a warp = 32 scenarios of the same bus, so the shared Ybus pattern/value loads are warp
//     broadcasts and every state load/store is a fully coalesced 512 B (double2) access;
//   * everything the factorisation touches is SCENARIO-MAJOR  [slot][entry]  (rhs, Jacobian 2x2
//     blocks, stored U panels, update-matrix arena, y, x): one CTA factors one front of one
//     scenario and streams contiguous 32 B blocks.
// Reference semantics restated per kernel (paths relative to Sparlectra.jl/src):
//   k_mismatch  - calc_injections + mismatch_rectangular        solver_core.jl:35-47,
//                                                               powerflow_rectangular/rectangular_core_equations.jl:79-148
//   k_qlimit    - _handle_rectangular_qlimit_iteration! + active_set_q_limits!
//                                                               powerflow_rectangular/rectangular_qlimit_iteration.jl:25-210, limits.jl:398-568
//   k_decide    - NR loop control                               powerflow_rectangular/rectangular_network_solver.jl:746-852
//   k_jacobian  - _assemble_rectangular_jacobian!               powerflow_rectangular/rectangular_jacobian_builders.jl:212-429
//   k_front / k_backsolve - solve_linear(J, -F) (UMFPACK there) solver_core.jl:60-111
//   k_update    - _apply_rectangular_delta! + slack reset       powerflow_rectangular/rectangular_newton_step.jl:38-55, rectangular_network_solver.jl:919
//   k_fin_*     - final acceptance + contingency metrics        rectangular_network_solver.jl:939-1094, contingency.jl:149-178, losses.jl:53-85
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

// Bounds-checked build (-DSBLX_BOUNDS_CHECK, tools/sanitize_run.py): compute-sanitizer is closed on the GPU pool, so
// every symbolic index the factorisation kernels follow is range-checked here instead; a violation traps the kernel
// (the host call then fails with a CUDA error).  Compiled out of the product build.
#ifdef SBLX_BOUNDS_CHECK
#define SBLX_CHECK(cond)                                                                                      \
  do {                                                                                                        \
    if (!(cond)) {                                                                                            \
      printf("SBLX_CHECK failed: %s (%s:%d) block (%d,%d) thread %d\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x); \
      __trap();                                                                                               \
    }                                                                                                         \
  } while (0)
#else
#define SBLX_CHECK(cond)
#endif

namespace sblx {

constexpr int ING_NONE = INT32_MIN, ING_MULTI = 1 << 30;   // ingest word encoding (same constants in symbolic.h)
constexpr int MAXOUT = 8;          // outaged branches per scenario (N-k, k <= 8)
constexpr int ST_IDLE = 0, ST_RUN = 1, ST_DRAIN = 2;
constexpr int T_ISO = 0, T_PQ = 1, T_PV = 2, T_SLACK = 3;

struct PanelDev {
  int first, w, h, nchild;
  int struct_ptr, child_ptr, ing_base, l_off;   // l_off: level-local offset (blocks) of this panel's L21 in the Lw workspace
  long long u_off, c_off;
  int ingf_base, reg_base, reg_vbase, pad_;     // reg_*: register-resident class (w + h <= 8), else reg_base = -1
};
// per (parent, child) record used by the Schur gather
struct ChildDev {
  long long c_off;
  int hc, inv_off;
  int kc, pad_;      // kc: leading structure entries of the child that are pivots of the parent
};

struct DevModel {
  int n, m, slack, nbr, nnzB, W;
  const int* y_rowptr; const int* y_col; const double2* y_val;
  const int* node_of_bus; const int* bus_of_node; const int* elim_of_node; const int* node_of_elim;
  const int* jb_row; const int* jb_col; const int* jb_ypos;
  const double2* S0; const double2* V0; const double* vset; const double* qmin; const double* qmax;
  const double* qload; const unsigned char* type0; const unsigned char* lockmask;
  const int* br_from; const int* br_to; const double2* y11; const double2* y12; const double2* y21; const double2* y22;
  const double* rating; const unsigned char* br_status;
  const PanelDev* panels; const int* struct_nodes; const int* child_list; const int* rel; const long long* rel_ptr;
  const int* kpiv; const int* inv; const long long* inv_ptr; const int* asm_src; const int* asm_dst;
  const int* ing_ptr; const int* ing_src; const int2* ing_pairs; const int* reg_words; const int* reg_vptr; const int* reg_vsrc; const int* ing_xptr; const int* ing_xsrc; const int* vec_ptr; const int* vec_src; const ChildDev* childdev;
  long long u_blocks, c_blocks, l_blocks;
  double tol, damp, hyst, vm_min, vm_max, base_mva, pivot_tol;   // pivot_tol: |det| <= pivot_tol * max|entry|^2 of a 2x2 pivot block counts as a tiny pivot
  int max_iter, qlim_enabled, qlim_start, cooldown, guard, freeze, allow_reenable;
  int pf_mode;   // L2 prefetch experiment (SBLX_PF_MODE): 1 next CTA's pivot-row sources, 2 next CTA's whole child matrices, 4 own Schur sources
};

struct Slots {
  // [n][W]
  double2* V; double2* I; double2* S; double* qload_s;
  unsigned char* type; unsigned char* iso; unsigned char* evcnt; unsigned char* evact; short* evlast;
  // [W][...]
  double2* rhs; double* Jv; double* U; double* C; double2* yv; double2* xv;
  double* Lw; double2* Dw;   // split path: L21 workspace [W][l_blocks*4], inverse pivot blocks [W][2*m]
  // [W]
  int* scen; int* iter; int* state; int* changed; int* nev; int* singular; int* numerical; int* reason;
  unsigned long long* mis_bits; unsigned long long* mis2_bits; double* fm;
  unsigned long long* pvres_bits; unsigned long long* vmin_bits; unsigned long long* vmax_bits; unsigned long long* load_bits;
  int* viol; int* maxsw; int* rated; int* nvv; int* nov; int* tiny; int* tinyit;
  int* refine; int* nrefine; double* fnorm; double2* xsave;   // iterative refinement of steps whose factorisation met a tiny pivot   // nvv / nov: voltage-band violations / overloaded branches; tiny: near-singular pivot blocks
  int* out_cnt; int* out_br;   // [W][MAXOUT]
};

struct Lists {
  int* active; int* fresh; int* step; int* fin;
  int* counts;   // [0] nActive [1] nFresh [2] nStep [3] nFin [4] queue head [5] nDrain [6] steps to refine in this tick [7] ... so far in this run
  int* refine;   // slots whose step needs iterative refinement (filled by k_tiny_guard)
};

struct Batch {
  int B, nq, use_sweep_q;      // B: scenarios on the device (caller scenarios + island sub-scenarios)
  int Btop;                    // caller scenarios; ids >= Btop are island sub-scenarios
  const int* queue; const int* out_ptr; const int* out_br; const int* iso_ptr; const int* iso_bus;
  const int* ref_ptr; const int* ref_bus;     // per scenario: PV buses promoted to island reference (ac_islands.jl:228-234)
  const int* lk_ptr; const int* lk_bus;       // per ISLAND sub-scenario (index scen - Btop): its locked PV buses, see stage_outages
  const double2* Ssweep; const double* qlsweep; const double2* Vstart;
  const int* parent;           // per ISLAND sub-scenario (index scen - Btop): the caller scenario it belongs to
  // device-side injection sweep (mc_probabilistic_powerflow.jl:28-41): 1 = independent truncated-normal factor per
  // (scenario, load bus) from Philox4x32-10 keyed on the seed, counter = (scenario, bus); 2 = one factor per scenario
  int sweep_mode; unsigned long long sweep_seed; long long sweep_first; double sweep_sigma, sweep_lo, sweep_hi; const double* scale;
  // results
  int* r_conv; int* r_status; int* r_iter; int* r_nsw; double* r_fm; double* r_vmin; double* r_vmax; double* r_load;
  double2* r_V; unsigned char* r_types;
  int* r_nvv; int* r_nov; int* r_tiny;      // per scenario: voltage-band violations, overloaded branches (contingency.jl:149-178), tiny pivots
  // compacted (scenario, bus) / (scenario, branch) lists, appended with one atomic per entry (order restored on the host)
  int2* viol_list; int2* over_list; double* over_pct; unsigned long long* list_counts; long long viol_cap, over_cap;
  int4* qlog_list; long long qlog_cap;      // Q-limit event log {scenario, iteration, bus, side (+1 max, -1 min)} = net.qLimitLog (limits.jl:75-84)
  double2* r_flows;            // [B][nbr][2] complex branch flows S_from, S_to in p.u. (losses.jl:53-85), optional
};

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
// One 2x2 block (32 B, 32 B aligned) per instruction: sm_100 has 256-bit global loads / stores (LDG.256 / STG.256),
// so a block is one full DRAM sector per thread and instruction instead of two half-sector accesses.
__device__ __forceinline__ void ldg_blk(const double2* p, double2& r0, double2& r1) {
#ifndef SBLX_LDG_ALLOC   // streamed once: keep L1 for the index lists shared by the co-resident CTAs
  asm("ld.global.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r0.x), "=d"(r0.y), "=d"(r1.x), "=d"(r1.y) : "l"(__cvta_generic_to_global(p)));
#else
  asm("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r0.x), "=d"(r0.y), "=d"(r1.x), "=d"(r1.y) : "l"(__cvta_generic_to_global(p)));
#endif
}
// same load, pinned in program order (volatile): for requests that must be ISSUED before a long arithmetic loop whose
// result they are only added to afterwards - ptxas otherwise sinks the load to its first use
__device__ __forceinline__ void ldg_blk_early(const double2* p, double2& r0, double2& r1) {
  asm volatile("ld.global.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r0.x), "=d"(r0.y), "=d"(r1.x), "=d"(r1.y) : "l"(__cvta_generic_to_global(p)) : "memory");
}
__device__ __forceinline__ void stg_blk(double2* p, double2 r0, double2 r1) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(__cvta_generic_to_global(p)), "d"(r0.x), "d"(r0.y), "d"(r1.x), "d"(r1.y) : "memory");
}
// One bulk L2 prefetch (UBLKPF.L2: uniform operands, so a single thread issues it); chunked, 16 B granularity.
__device__ __forceinline__ void l2_prefetch_range(const void* p, size_t bytes) {
  const char* c = reinterpret_cast<const char*>(p);
  bytes = (bytes + 15) & ~(size_t)15;
  while (bytes > 0) {
    const unsigned n = bytes > 32768 ? 32768u : (unsigned)bytes;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(c)), "r"(n) : "memory");
    c += n;
    bytes -= n;
  }
}
// 16-byte asynchronous global -> shared copy (LDGSTS), no register staging
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(__cvta_generic_to_global(gmem)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ unsigned long long abs_bits(double x) { return (unsigned long long)__double_as_longlong(fabs(x)); }
__device__ __forceinline__ double bits_double(unsigned long long b) { return __longlong_as_double((long long)b); }

// exclusive scan of 0/1 flags over a block of 1024 threads; returns rank, writes block total
__device__ __forceinline__ int block_rank_1024(int flag, int* total) {
  __shared__ int wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned b = __ballot_sync(0xffffffffu, flag);
  int r = __popc(b & ((1u << lane) - 1u));
  if (lane == 0) wsum[warp] = __popc(b);
  __syncthreads();
  if (warp == 0) {
    int v = wsum[lane];
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    wsum[lane] = s - v;            // exclusive
    if (lane == 31) *total = s;
  }
  __syncthreads();
  r += wsum[warp];
  __syncthreads();
  return r;
}

// Philox4x32-10 (Salmon et al., SC'11): counter-based, so the factor of (scenario, bus) does not depend on which slot
// or wave the scenario lands in.  key = seed, counter = (scenario lo, scenario hi, bus, 0).
__host__ __device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned out[4]) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)M0 * c0, p1 = (unsigned long long)M1 * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// truncated-normal load factor f = clip(1 + sigma * z, lo, hi), z by Box-Muller from two 53-bit uniforms in (0, 1]
__host__ __device__ __forceinline__ double sweep_factor(unsigned long long seed, long long scen, int bus, double sigma, double lo, double hi) {
  unsigned r[4];
  philox4x32_10((unsigned)scen, (unsigned)((unsigned long long)scen >> 32), (unsigned)bus, 0u, (unsigned)seed, (unsigned)(seed >> 32), r);
  const double u1 = ((double)((((unsigned long long)r[0] << 32) | r[1]) >> 11) + 1.0) * (1.0 / 9007199254740992.0);
  const double u2 = ((double)((((unsigned long long)r[2] << 32) | r[3]) >> 11) + 1.0) * (1.0 / 9007199254740992.0);
  const double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
  const double f = 1.0 + sigma * z;
  return f < lo ? lo : f > hi ? hi : f;
}

// the factors k_init applies in sweep mode 1, for callers / parity tests: out[s * n + bus], 1.0 on non-load buses
__global__ void k_sweep_factors(DevModel M, long long first, int nb, unsigned long long seed, double sigma, double lo, double hi, double* out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nb * M.n) return;
  const int s = (int)(idx / M.n), bus = (int)(idx - (long long)s * M.n);
  out[idx] = M.S0[bus].x < 0.0 ? sweep_factor(seed, first + s, bus, sigma, lo, hi) : 1.0;
}

// ------------------------------------------------------------------------------------------
// tick bookkeeping: assign pending scenarios to idle slots, build active / drain lists
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_assign(DevModel M, Slots A, Lists L, Batch Bt) {
  __shared__ int s_total;
  __shared__ int s_qhead, s_nfresh, s_nact, s_ndrain;
  if (threadIdx.x == 0) { s_qhead = L.counts[4]; s_nfresh = 0; s_nact = 0; s_ndrain = 0; }
  __syncthreads();
  for (int base = 0; base < M.W; base += 1024) {
    const int slot = base + threadIdx.x;
    const bool valid = slot < M.W;
    int st = valid ? A.state[slot] : -1;
    int idle = (st == ST_IDLE);
    int rank = block_rank_1024(idle, &s_total);
    int qh = s_qhead, nf = s_nfresh;
    if (idle) {
      int q = qh + rank;
      if (q < Bt.nq) {
        A.scen[slot] = Bt.queue[q];
        A.state[slot] = ST_RUN;
        A.iter[slot] = 0;
        L.fresh[nf + rank] = slot;
        st = ST_RUN;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int take = min(s_total, max(Bt.nq - qh, 0));
      s_qhead = qh + take;
      s_nfresh = nf + take;
    }
    __syncthreads();
    int run = (st == ST_RUN);
    int r2 = block_rank_1024(run, &s_total);
    int na = s_nact;
    if (run) {
      L.active[na + r2] = slot;
      A.mis_bits[slot] = 0ull;
      A.mis2_bits[slot] = 0ull;
      A.changed[slot] = 0;
      A.tinyit[slot] = 0;
      A.refine[slot] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_nact = na + s_total;
    __syncthreads();
    int dr = (st == ST_DRAIN);
    int r3 = block_rank_1024(dr, &s_total);
    int nd = s_ndrain;
    if (dr) L.fin[nd + r3] = slot;
    __syncthreads();
    if (threadIdx.x == 0) s_ndrain = nd + s_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    L.counts[6] = 0;
    L.counts[0] = s_nact;
    L.counts[1] = s_nfresh;
    L.counts[4] = s_qhead;
    L.counts[5] = s_ndrain;
  }
}

// grid (ceil(n/8), ceil(nFresh/32)), block (32, 8)
__global__ void __launch_bounds__(256) k_init(DevModel M, Slots A, Lists L, Batch Bt) {
  const int nf = L.counts[1];
  const int li = blockIdx.y * 32 + threadIdx.x;
  const int bus = blockIdx.x * 8 + threadIdx.y;
  if (li >= nf || bus >= M.n) return;
  const int slot = L.fresh[li];
  const int scen = A.scen[slot];
  const size_t k = (size_t)bus * M.W + slot;
  A.V[k] = Bt.Vstart ? Bt.Vstart[bus] : M.V0[bus];
  // per-scenario injections are indexed by the CALLER scenario: an island sub-scenario (id >= Btop) uses its parent's
  const int pscen = (scen >= Bt.Btop && Bt.parent) ? Bt.parent[scen - Bt.Btop] : scen;
  double2 Sv = Bt.Ssweep ? Bt.Ssweep[(size_t)pscen * M.n + bus] : M.S0[bus];
  double qlv = Bt.qlsweep ? Bt.qlsweep[(size_t)pscen * M.n + bus] : M.qload[bus];
  if (Bt.sweep_mode != 0 && Sv.x < 0.0) {          // load buses only (net consumption), P and Q scaled together
    const double f = Bt.sweep_mode == 2 ? Bt.scale[pscen] : sweep_factor(Bt.sweep_seed, Bt.sweep_first + pscen, bus, Bt.sweep_sigma, Bt.sweep_lo, Bt.sweep_hi);
    Sv.x *= f; Sv.y *= f; qlv *= f;
  }
  A.S[k] = Sv;
  if (A.qload_s) A.qload_s[k] = qlv;
  A.type[k] = M.type0[bus];
  A.iso[k] = 0;
  A.evcnt[k] = 0;
  A.evact[k] = 0;
  A.evlast[k] = 0;
  A.I[k] = make_double2(0.0, 0.0);
  if (bus == 0) {
    A.nev[slot] = 0;
    A.singular[slot] = 0;
    A.numerical[slot] = 0;
    A.reason[slot] = 1;
    A.fm[slot] = 0.0;
    A.tiny[slot] = 0;
    A.nrefine[slot] = 0;
    int cnt = 0;
    if (Bt.out_ptr) {
      int b = Bt.out_ptr[scen], e = Bt.out_ptr[scen + 1];
      for (int q = b; q < e && cnt < MAXOUT; ++q) {
        int br = Bt.out_br[q];
        if (br >= 0 && br < M.nbr && M.br_status[br]) A.out_br[slot * MAXOUT + cnt++] = br;
      }
    }
    A.out_cnt[slot] = cnt;
  }
}

// isolated buses of the freshly assigned scenarios (host-detected: markIsolatedBuses!,
// network.jl:1881-1930): neutral PQ rows with S = 0, V = 1 (rectangular_network_solver.jl:508-512,
// network.jl:2318 `vm <= 0 => 1.0`).  grid ceil(nFresh/128)
__global__ void k_init_iso(DevModel M, Slots A, Lists L, Batch Bt) {
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= L.counts[1] || !Bt.iso_ptr) return;
  const int slot = L.fresh[li];
  const int scen = A.scen[slot];
  for (int q = Bt.iso_ptr[scen]; q < Bt.iso_ptr[scen + 1]; ++q) {
    const int bus = Bt.iso_bus[q];
    const size_t k = (size_t)bus * M.W + slot;
    A.iso[k] = 1;
    A.type[k] = T_PQ;
    A.S[k] = make_double2(0.0, 0.0);
    A.V[k] = make_double2(1.0, 0.0);
  }
  if (Bt.ref_ptr) {
    // promoted island reference: held like a slack at |V| = Vset with the start angle
    // (rectangular_network_solver.jl:523 on the island sub-net, ac_islands.jl:400-413)
    for (int q = Bt.ref_ptr[scen]; q < Bt.ref_ptr[scen + 1]; ++q) {
      const int bus = Bt.ref_bus[q];
      const size_t k = (size_t)bus * M.W + slot;
      A.type[k] = T_SLACK;
      const double2 v = A.V[k];
      const double vm = sqrt(v.x * v.x + v.y * v.y);
      const double vs = M.vset[bus];
      A.V[k] = vm > 0.0 ? make_double2(v.x * (vs / vm), v.y * (vs / vm)) : make_double2(vs, 0.0);
    }
  }
}

// ------------------------------------------------------------------------------------------
// K1: I = Ybus(s) V, S_calc = V conj(I), rhs = -F, per-scenario max |F|
// grid (ceil(n/8), ceil(nActive/32)), block (32, 8): lane = scenario, ty = bus
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 residual_rows(int type, double2 V, double2 Sc, double2 S, double vset) {
  // F = calc - spec ; PV second row is |V| - Vset (rectangular_core_equations.jl:116-128)
  double f0 = Sc.x - S.x;
  double f1 = (type == T_PV) ? (sqrt(V.x * V.x + V.y * V.y) - vset) : (Sc.y - S.y);
  return make_double2(f0, f1);
}

__global__ void __launch_bounds__(256) k_mismatch(DevModel M, Slots A, Lists L) {
  __shared__ unsigned long long smax[8][32];
  const int na = L.counts[0];
  const int li = blockIdx.y * 32 + threadIdx.x;
  const int bus = blockIdx.x * 8 + threadIdx.y;
  unsigned long long mybits = 0ull;
  int slot = -1;
  if (li < na && bus < M.n) {
    slot = L.active[li];
    const size_t W = M.W;
    const size_t k = (size_t)bus * W + slot;
    double2 I = make_double2(0.0, 0.0);
    const double2 Vi = A.V[k];
    if (!A.iso[k]) {
      for (int p = M.y_rowptr[bus]; p < M.y_rowptr[bus + 1]; ++p) {
        const double2 y = M.y_val[p];
        SBLX_CHECK(M.y_col[p] >= 0 && M.y_col[p] < M.n && slot >= 0 && slot < M.W);
        const double2 v = A.V[(size_t)M.y_col[p] * W + slot];
        I.x += y.x * v.x - y.y * v.y;
        I.y += y.x * v.y + y.y * v.x;
      }
      const int no = A.out_cnt[slot];
      for (int o = 0; o < no; ++o) {
        const int br = A.out_br[slot * MAXOUT + o];
        const int f = M.br_from[br], t = M.br_to[br];
        if (bus == f) {
          I = csub(I, cadd(cmul(M.y11[br], Vi), cmul(M.y12[br], A.V[(size_t)t * W + slot])));
        } else if (bus == t) {
          I = csub(I, cadd(cmul(M.y22[br], Vi), cmul(M.y21[br], A.V[(size_t)f * W + slot])));
        }
      }
    }
    A.I[k] = I;
    if (bus != M.slack) {
      const double2 Sc = cmul(Vi, cconj(I));
      const int ty = A.type[k];
      const double2 F = (ty == T_SLACK) ? make_double2(0.0, 0.0) : residual_rows(ty, Vi, Sc, A.S[k], M.vset[bus]);
      const int e = M.elim_of_node[M.node_of_bus[bus]];
      A.rhs[(size_t)slot * M.m + e] = make_double2(-F.x, -F.y);
      unsigned long long b0 = abs_bits(F.x), b1 = abs_bits(F.y);
      mybits = b0 > b1 ? b0 : b1;
    }
  }
  smax[threadIdx.y][threadIdx.x] = mybits;
  __syncthreads();
  if (threadIdx.y == 0 && slot >= 0) {
    unsigned long long v = smax[0][threadIdx.x];
#pragma unroll
    for (int r = 1; r < 8; ++r) {
      unsigned long long t = smax[r][threadIdx.x];
      v = t > v ? t : v;
    }
    atomicMax(&A.mis_bits[slot], v);   // NaN bit patterns sort above +Inf: NaN-propagating max
  }
}

// ------------------------------------------------------------------------------------------
// K5: Q-limit active set. Same mapping as k_mismatch.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_qlimit(DevModel M, Slots A, Lists L, Batch Bt) {
  const int na = L.counts[0];
  const int li = blockIdx.y * 32 + threadIdx.x;
  const int bus = blockIdx.x * 8 + threadIdx.y;
  if (li >= na || bus >= M.n || bus == M.slack) return;
  const int slot = L.active[li];
  const size_t k = (size_t)bus * M.W + slot;
  const int type0 = M.type0[bus];
  if (type0 != T_PV || A.iso[k]) return;       // only originally-PV buses ever take part
  const double mis = bits_double(A.mis_bits[slot]);
  const int it = A.iter[slot] + 1;
  const bool conv_now = mis <= M.tol;           // false for NaN
  const bool ready = (it >= M.qlim_start) || conv_now;   // rectangular_qlimit_iteration.jl:86,111
  if (!ready) return;
  const double qmin = M.qmin[bus], qmax = M.qmax[bus];
  const bool has_lo = isfinite(qmin), has_hi = isfinite(qmax);
  if (!(has_lo || has_hi)) return;
  int type = A.type[k];
  const double2 Vi = A.V[k];
  const double2 Sc = cmul(Vi, cconj(A.I[k]));
  const double ql = A.qload_s ? A.qload_s[k] : M.qload[bus];
  const double qreq = Sc.y + ql;
  bool touched = false;
  bool locked = M.lockmask[bus] != 0;
  const int scen = A.scen[slot];
  if (Bt.lk_ptr && scen >= Bt.Btop) {              // island sub-scenario: the lock list addresses the island's own numbering
    locked = false;
    for (int q = Bt.lk_ptr[scen - Bt.Btop]; q < Bt.lk_ptr[scen - Bt.Btop + 1]; ++q) locked |= Bt.lk_bus[q] == bus;
  }
  if (type == T_PV && !locked) {
    int side = 0;
    double clamp = 0.0;
    if (has_hi && qreq > qmax + M.hyst) { side = 1; clamp = qmax; }
    else if (has_lo && qreq < qmin - M.hyst) { side = -1; clamp = qmin; }
    if (side != 0 && !(M.freeze && A.evcnt[k] >= M.guard)) {
      type = T_PQ;                                          // make_pq! (rectangular_qlimit_iteration.jl:178-183)
      A.type[k] = T_PQ;
      double2 S = A.S[k];
      S.y = clamp - ql;
      A.S[k] = S;
      A.evcnt[k] = (unsigned char)min(255, A.evcnt[k] + 1);   // logQLimitHit!
      A.evact[k] = 1;
      A.evlast[k] = (short)it;
      atomicAdd(&A.nev[slot], 1);
      atomicOr(&A.changed[slot], 1);
      touched = true;
      if (Bt.qlog_list) {
        const unsigned long long q = atomicAdd(&Bt.list_counts[2], 1ull);
        if ((long long)q < Bt.qlog_cap) Bt.qlog_list[q] = make_int4(scen, it, bus, side);
      }
    }
  }
  if (M.allow_reenable && type != T_PV && A.evact[k] && A.evcnt[k] <= 1) {    // limits.jl:527-561
    const double lo = has_lo ? qmin + M.hyst : -INFINITY;
    const double hi = has_hi ? qmax - M.hyst : INFINITY;
    bool rdy = (qreq > lo) && (qreq < hi);
    if (rdy && M.cooldown > 0 && (it - (int)A.evlast[k]) < M.cooldown) rdy = false;
    if (rdy) {
      type = T_PV;
      A.type[k] = T_PV;
      A.evact[k] = 0;
      atomicOr(&A.changed[slot], 2);
      touched = true;
    }
  }
  if (touched) {
    const double2 F = residual_rows(type, Vi, Sc, A.S[k], M.vset[bus]);
    const int e = M.elim_of_node[M.node_of_bus[bus]];
    A.rhs[(size_t)slot * M.m + e] = make_double2(-F.x, -F.y);
  }
}

// mismatch norm after an active-set change (history[end] = max_mis, rectangular_qlimit_iteration.jl:195-204)
// grid nActive, block 256
__global__ void __launch_bounds__(256) k_norm(DevModel M, Slots A, Lists L) {
  __shared__ unsigned long long red[256];
  if ((int)blockIdx.x >= L.counts[0]) return;
  const int slot = L.active[blockIdx.x];
  if (!A.changed[slot]) return;
  const double2* r = A.rhs + (size_t)slot * M.m;
  unsigned long long v = 0ull;
  for (int e = threadIdx.x; e < M.m; e += 256) {
    double2 x = r[e];
    unsigned long long a = abs_bits(x.x), b = abs_bits(x.y);
    a = a > b ? a : b;
    v = a > v ? a : v;
  }
  red[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      unsigned long long t = red[threadIdx.x + s];
      if (t > red[threadIdx.x]) red[threadIdx.x] = t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) A.mis2_bits[slot] = red[0];
}

// NR loop control (rectangular_network_solver.jl:746-852). single block of 1024 threads.
__global__ void __launch_bounds__(1024) k_decide(DevModel M, Slots A, Lists L) {
  __shared__ int s_total, s_nstep, s_nfin;
  const int na = L.counts[0];
  if (threadIdx.x == 0) { s_nstep = 0; s_nfin = L.counts[5]; }
  __syncthreads();
  for (int base = 0; base < na; base += 1024) {
    const int li = base + threadIdx.x;
    int step = 0, fin = 0, slot = -1;
    if (li < na) {
      slot = L.active[li];
      const int it = A.iter[slot] + 1;
      A.iter[slot] = it;
      const double mis = bits_double(A.mis_bits[slot]);
      const int ch = A.changed[slot];
      A.fm[slot] = ch ? bits_double(A.mis2_bits[slot]) : mis;
      if (!isfinite(mis)) {
        A.numerical[slot] = 0; A.reason[slot] = 2; fin = 1;      // :nr_nonfinite
      } else if (mis <= M.tol && !ch) {
        A.numerical[slot] = 1; A.reason[slot] = 0; fin = 1;      // converged, active set stable
      } else {
        step = 1;
      }
    }
    int r = block_rank_1024(step, &s_total);
    int ns = s_nstep;
    if (step) L.step[ns + r] = slot;
    __syncthreads();
    if (threadIdx.x == 0) s_nstep = ns + s_total;
    __syncthreads();
    int r2 = block_rank_1024(fin, &s_total);
    int nf = s_nfin;
    if (fin) L.fin[nf + r2] = slot;
    __syncthreads();
    if (threadIdx.x == 0) s_nfin = nf + s_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) { L.counts[2] = s_nstep; L.counts[3] = s_nfin; }
}

// ------------------------------------------------------------------------------------------
// K2: Jacobian 2x2 blocks into the fixed superset pattern.
// grid (ceil(nnzB/32), ceil(nStep/32)), block (32, 8): one CTA = 32 scenarios x 32 block entries.
// Compute phase: lane = scenario (the per-bus state is scenario-interleaved: coalesced reads, shared Ybus
// values are warp broadcasts), 4 entries per thread, blocks parked in a padded shared-memory tile.
// Store phase: lane = entry, so each warp writes 1 KB of one scenario's contiguous Jacobian row
// (the Jacobian is scenario-major for the factorisation) with 256-bit stores.
// block = [dP/dVr dP/dVi ; dQ/dVr dQ/dVi]  (PV: second row = d|V|/dVr, d|V|/dVi on the diagonal)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_jacobian(DevModel M, Slots A, Lists L) {
  __shared__ double2 tile[32][33][2];                  // [scenario][entry (+1 pad)][block row]
  __shared__ int slots[32];
  const int ns = L.counts[2];
  const int li = blockIdx.y * 32 + threadIdx.x;
  const bool live = li < ns;
  const int slot = live ? L.step[li] : 0;
  if (threadIdx.y == 0) slots[threadIdx.x] = live ? slot : -1;
  const size_t W = M.W;
  const int ent0 = blockIdx.x * 32;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int e = threadIdx.y + 8 * q, ent = ent0 + e;
    if (!live || ent >= M.nnzB) continue;
    const int i = M.bus_of_node[M.jb_row[ent]], j = M.bus_of_node[M.jb_col[ent]];
    SBLX_CHECK(i >= 0 && i < M.n && j >= 0 && j < M.n && M.jb_ypos[ent] >= 0 && slot >= 0 && slot < M.W);
    const size_t ki = (size_t)i * W + slot;
    double b00 = 0, b01 = 0, b10 = 0, b11 = 0;
    const bool isoi = A.iso[ki] != 0, isoj = A.iso[(size_t)j * W + slot] != 0;
    if (isoi || isoj || A.type[ki] == T_SLACK) {
      if (i == j) { b00 = 1.0; b11 = 1.0; }              // neutral identity row (isolated bus / promoted island reference): delta = 0
    } else {
      double2 y = M.y_val[M.jb_ypos[ent]];
      const int no = A.out_cnt[slot];
      for (int o = 0; o < no; ++o) {
        const int br = A.out_br[slot * MAXOUT + o];
        const int f = M.br_from[br], t = M.br_to[br];
        if (i == j) {
          if (i == f) y = csub(y, M.y11[br]);
          else if (i == t) y = csub(y, M.y22[br]);
        } else if (i == f && j == t) y = csub(y, M.y12[br]);
        else if (i == t && j == f) y = csub(y, M.y21[br]);
      }
      const double2 Vi = A.V[ki];
      const double2 a = cmul(Vi, cconj(y));        // A = V_i conj(Y_ij)
      double2 dvr = a;                              // dS/dVr
      double2 dvi = make_double2(a.y, -a.x);        // dS/dVi = i*(-A)
      if (i == j) {
        const double2 I = A.I[ki];
        dvr.x += I.x; dvr.y -= I.y;                 // + conj(I)
        dvi.x += I.y; dvi.y += I.x;                 // + i*conj(I)
      }
      b00 = dvr.x; b01 = dvi.x;
      if (A.type[ki] == T_PV) {
        if (i == j) {
          const double vm = sqrt(Vi.x * Vi.x + Vi.y * Vi.y);
          const double d = vm > 0.0 ? vm : 1e-9;    // vm_eps (rectangular_jacobian_builders.jl:362)
          b10 = Vi.x / d; b11 = Vi.y / d;
        }
      } else {
        b10 = dvr.y; b11 = dvi.y;
      }
    }
    tile[threadIdx.x][e][0] = make_double2(b00, b01);
    tile[threadIdx.x][e][1] = make_double2(b10, b11);
  }
  __syncthreads();
  const int e = threadIdx.x, ent = ent0 + e;
  if (ent >= M.nnzB) return;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int sc = threadIdx.y + 8 * q;
    const int sl = slots[sc];
    if (sl < 0) continue;
    double2* out = reinterpret_cast<double2*>(A.Jv + ((size_t)sl * M.nnzB + ent) * 4);
    stg_blk(out, tile[sc][e][0], tile[sc][e][1]);
  }
}

// Phase ablation build (-DSBLX_ABLATE, tools only): bits of SBLX_PF_MODE >= 8 switch phases of k_front OFF so that their
// real cost (with all the overlap between co-resident CTAs) shows up as a time difference.  Results are garbage.
#ifdef SBLX_ABLATE
#define ABL(bit) ((M.pf_mode & (bit)) != 0)
#else
#define ABL(bit) false
#endif
#ifdef SBLX_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[5][10];   // [class][phase]; [9] = CTA count
#define PHASE_MARK(ph)                                                                             \
  do {                                                                                             \
    if (threadIdx.x == 0) {                                                                        \
      const long long now_ = clock64();                                                            \
      atomicAdd(&g_phase_cycles[NT == 32 ? 0 : NT == 64 ? 1 : NT == 128 ? 2 : NT == 256 ? 3 : 4][ph], (unsigned long long)(now_ - tmark_)); \
      tmark_ = now_;                                                                               \
    }                                                                                              \
  } while (0)
#else
#define PHASE_MARK(ph)
#endif

// ------------------------------------------------------------------------------------------
// K3: multifrontal panel factorisation, one CTA per (front, scenario).
// shared memory (double2 = one block row): T0/T1 [w][nf] pivot rows, L0/L1 [h][w] structure
// rows x pivot cols, b1 [w] rhs of the pivots.  T rows are stored pre-scaled by the inverse
// pivot block (U' = D^-1 U, y' = D^-1 y) so the back substitution has no division.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool inv2x2(double2 r0, double2 r1, double2& i0, double2& i1) {
  // inverse of [[a,b],[c,d]] with the determinant taken through the partially pivoted LU
  const double a = r0.x, b = r0.y, c = r1.x, d = r1.y;
  double det;
  if (fabs(a) >= fabs(c)) det = (a == 0.0) ? 0.0 : a * (d - (c / a) * b);
  else det = -c * (b - (a / c) * d);
  const bool ok = isfinite(det) && det != 0.0;
  const double s = ok ? 1.0 / det : 0.0;
  i0 = make_double2(d * s, -b * s);
  i1 = make_double2(-c * s, a * s);
  return ok;
}
// t -= a * u for 2x2 blocks (rows as double2): 8 DFMA, no separate multiplies/adds
__device__ __forceinline__ void bmsub(double2& t0, double2& t1, double2 a0, double2 a1, double2 u0, double2 u1) {
  t0.x = fma(-a0.x, u0.x, t0.x); t0.x = fma(-a0.y, u1.x, t0.x);
  t0.y = fma(-a0.x, u0.y, t0.y); t0.y = fma(-a0.y, u1.y, t0.y);
  t1.x = fma(-a1.x, u0.x, t1.x); t1.x = fma(-a1.y, u1.x, t1.x);
  t1.y = fma(-a1.x, u0.y, t1.y); t1.y = fma(-a1.y, u1.y, t1.y);
}

// Grid: x = scenario (fastest: co-resident CTAs run the SAME front of different scenarios, so the shared index
// lists hit L1 and the warps follow one code path), y = front of this launch group.
// G = lanes per scenario.  G == NT: the whole CTA works on one (front, scenario).  G < NT (one-warp CTAs only):
// the warp is split into NT/G lane groups, each factoring the same small front of a DIFFERENT scenario in its own
// slice of shared memory - tiny fronts (w <= G/2 pivots) leave most of a warp idle otherwise, and the per-front
// instruction overhead is shared by the groups.
// WIDE (64-thread class only): 16-row register column, for the few small fronts whose pivot block is wider than 8
// buses (dense root fronts of a minimum-degree tree) - they would otherwise fall into the 128-thread class.
template <int NT, int G = NT, bool WIDE = false>
#ifndef SBLX_INGEST_BATCH
#define SBLX_INGEST_BATCH 2
#endif
#ifndef SBLX_OCC128
#define SBLX_OCC128 3
#endif
#ifndef SBLX_OCC32G
#define SBLX_OCC32G 8
#endif
#ifndef SBLX_OCC64
#define SBLX_OCC64 10
#endif
#ifndef SBLX_OCC256
#define SBLX_OCC256 2
#endif
__global__ void __launch_bounds__(NT, (NT == 256 ? SBLX_OCC256 : NT == 128 ? SBLX_OCC128 : NT == 64 ? (WIDE ? 6 : SBLX_OCC64) : NT == 32 ? (G < 32 ? SBLX_OCC32G : 16) : 1))
k_front(DevModel M, Slots A, Lists L, const PanelDev* __restrict__ group_panels, int group_stride, int pf_delta) {
  static_assert(G == NT || NT == 32, "lane groups need a one-warp CTA");
  extern __shared__ double2 sm2_all[];
  constexpr int NG = NT / G;
  const int nact = L.counts[2];
  if ((int)blockIdx.x * NG >= nact) return;
  // a group past the end repeats the last scenario (identical values to identical addresses): no predication needed
  const int bscen = min((int)blockIdx.x * NG + (int)threadIdx.x / G, nact - 1);
  const int slot = L.step[bscen];
  const PanelDev P = group_panels[blockIdx.y];          // the launch group's panel records, contiguous: no id indirection
  const int w = P.w, h = P.h, nf = w + h;
  const int tid = G == NT ? (int)threadIdx.x : (int)threadIdx.x % G;    // thread index within the scenario's group
  double2* sm2 = sm2_all + (G == NT ? 0 : ((int)threadIdx.x / G) * group_stride);
  auto front_sync = [&]() { if (NT == 32) __syncwarp(); else __syncthreads(); };
#ifdef SBLX_PHASE_TIMING
  long long tmark_ = clock64();
#endif
  // L2 prefetch (one thread of the LAST warp - warp 0 owns the serial pivot chain): the update matrices this front
  // reads were written one launch earlier by thousands of other scenarios' CTAs and are long gone from L2, so every
  // ingest / gather load is a full HBM round trip at low occupancy.  The CTA that will follow this one on the SM
  // (pf_delta CTAs ahead in launch order: x = scenario fastest, then the next front) gets its children's matrices
  // pulled into L2 now with a few bulk prefetches; its loads then pay L2 latency instead of HBM latency.
  if (G == NT && M.pf_mode != 0 && (int)threadIdx.x == (NT > 32 ? NT - 32 : 0)) {
    if ((M.pf_mode & 3) && pf_delta > 0) {
      const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x + pf_delta;
      const int y2 = (int)(lin / gridDim.x), x2 = (int)(lin - (long long)y2 * gridDim.x);
      if (y2 < (int)gridDim.y && x2 < nact) {
        const PanelDev P2 = group_panels[y2];
        const int slot2 = L.step[x2];
        const double* Cs2 = A.C + (size_t)slot2 * M.c_blocks * 4;
        for (int ci = 0; ci < P2.nchild; ++ci) {
          const ChildDev Cd = M.childdev[P2.child_ptr + ci];
          const size_t nb = (M.pf_mode & 2) ? (size_t)Cd.hc * Cd.hc * 32 + (size_t)Cd.hc * 16 : (size_t)Cd.kc * Cd.hc * 32;
          l2_prefetch_range(Cs2 + Cd.c_off * 4, nb);
        }
      }
    }
    if (M.pf_mode & 4) {
      const double* Cs1 = A.C + (size_t)slot * M.c_blocks * 4;
      for (int ci = 0; ci < P.nchild; ++ci) {
        const ChildDev Cd = M.childdev[P.child_ptr + ci];
        l2_prefetch_range(Cs1 + (Cd.c_off + (long long)Cd.kc * Cd.hc) * 4, (size_t)(Cd.hc - Cd.kc) * Cd.hc * 32 + (size_t)Cd.hc * 16);
      }
    }
  }
  // T0/T1: pivot rows [w][nf] (block row 0 / block row 1 planes); L0/L1: structure rows x pivot columns stored
  // PIVOT-MAJOR [w][h] so that consecutive threads (consecutive structure rows) hit consecutive 16 B words.
  double2* T0 = sm2;
  double2* T1 = T0 + w * nf;
  double2* L0 = T1 + w * nf;
  double2* L1 = L0 + h * w;
  double2* b1 = L1 + h * w;
  const int nT = w * nf, nL = h * w;
  double* Cs = A.C + (size_t)slot * M.c_blocks * 4;
  // ---- gather-style ingest: every panel block sums its sources (its Jacobian block, then the matching
  // blocks of the children's update matrices, in child order) and is written to shared memory once.
  // Phase A (all threads): the w x w pivot block and the pivots' rhs.  Phase B: the rest of the panel - in a
  // multi-warp CTA it is done by warps 1.. WHILE warp 0 factors the pivot block in registers.
  const double2* Jv = reinterpret_cast<const double2*>(A.Jv + (size_t)slot * M.nnzB * 4);
  const double2* Ca = reinterpret_cast<const double2*>(Cs);
  const int2* ipairs = M.ing_pairs + P.ingf_base;       // (source word, shared-memory block index), in walk order
  auto store_block = [&](int k, double2 x0, double2 x1) {
    if (k < nT) { T0[k] = x0; T1[k] = x1; } else { L0[k - nT] = x0; L1[k - nT] = x1; }
  };
  // IB panel blocks per thread and step: the IB index pairs are loaded first, then all value loads are in flight
  // together (single-source blocks: one dependent load level); multi-source blocks (8 %) take the CSR side path.
  constexpr int IB = SBLX_INGEST_BATCH;
#ifdef SBLX_CPASYNC
  // asynchronous variant: single-source blocks (92 %) go global -> shared with two 16-byte LDGSTS each, nothing is staged
  // in registers and a thread keeps ALL its blocks in flight; multi-source blocks take the synchronous side path.
  // The caller waits (cp.async.wait_group) before the barrier that publishes the panel.
  auto ingest_range = [&](int lo, int hi, int t0, int tstep) {
    for (int t = lo + t0; t < hi; t += tstep) {
      const int2 e = ipairs[t];
      const int v = e.x, k = e.y;
      double2* d0 = k < nT ? T0 + k : L0 + (k - nT);
      double2* d1 = k < nT ? T1 + k : L1 + (k - nT);
      if (v == ING_NONE) {
        *d0 = make_double2(0.0, 0.0);
        *d1 = make_double2(0.0, 0.0);
      } else if (v < ING_MULTI) {
        const double2* ptr = v < 0 ? Jv + 2 * (size_t)(~v) : Ca + 2 * (size_t)v;
        cp_async16(d0, ptr);
        cp_async16(d1, ptr + 1);
      } else {
        const int x = v - ING_MULTI;
        double2 x0 = make_double2(0.0, 0.0), x1 = x0;
        for (int q = M.ing_xptr[x]; q < M.ing_xptr[x + 1]; ++q) {
          const int sidx = M.ing_xsrc[q];
          const double2* ptr = sidx < 0 ? Jv + 2 * (size_t)(~sidx) : Ca + 2 * (size_t)sidx;
          double2 s0, s1;
          ldg_blk(ptr, s0, s1);
          x0 = cadd(x0, s0); x1 = cadd(x1, s1);
        }
        *d0 = x0;
        *d1 = x1;
      }
    }
  };
#else
  auto ingest_range = [&](int lo, int hi, int t0, int tstep) {
    for (int t = lo + t0; t < hi; t += IB * tstep) {
      int2 e[IB];
#pragma unroll
      for (int i = 0; i < IB; ++i) e[i] = (t + i * tstep < hi) ? ipairs[t + i * tstep] : make_int2(ING_NONE, -1);
      double2 x0[IB], x1[IB];
#pragma unroll
      for (int i = 0; i < IB; ++i) {
        const int v = e[i].x;
        SBLX_CHECK(v == ING_NONE || v >= ING_MULTI || (v < 0 ? (~v) < M.nnzB : (long long)v < M.c_blocks));
        SBLX_CHECK(e[i].y < nT + nL);
        const double2* ptr = v < 0 ? Jv + 2 * (size_t)(~v) : Ca + 2 * (size_t)v;
        x0[i] = make_double2(0.0, 0.0);
        x1[i] = x0[i];
        if (v != ING_NONE && v < ING_MULTI) ldg_blk(ptr, x0[i], x1[i]);
      }
#pragma unroll
      for (int i = 0; i < IB; ++i)
        if (e[i].x >= ING_MULTI) {
          const int x = e[i].x - ING_MULTI;
          for (int q = M.ing_xptr[x]; q < M.ing_xptr[x + 1]; ++q) {
            const int sidx = M.ing_xsrc[q];
            SBLX_CHECK(sidx < 0 ? (~sidx) < M.nnzB : (long long)sidx < M.c_blocks);
            const double2* ptr = sidx < 0 ? Jv + 2 * (size_t)(~sidx) : Ca + 2 * (size_t)sidx;
            double2 s0, s1;
            ldg_blk(ptr, s0, s1);
            x0[i] = cadd(x0[i], s0); x1[i] = cadd(x1[i], s1);
          }
        }
#pragma unroll
      for (int i = 0; i < IB; ++i)
        if (e[i].y >= 0) store_block(e[i].y, x0[i], x1[i]);
    }
  };
#endif
  {
    ingest_range(0, w * w, tid, G);                     // the pivot block
#ifdef SBLX_CPASYNC
    cp_async_commit();
#endif
    const double2* r = A.rhs + (size_t)slot * M.m + P.first;
    const int* vptr = M.vec_ptr + P.first;
    for (int p = tid; p < w; p += G) {
      double2 v = r[p];
      for (int q = vptr[p]; q < vptr[p + 1]; ++q) {
        SBLX_CHECK(M.vec_src[q] >= 0 && (long long)M.vec_src[q] < 2 * M.c_blocks);
        v = cadd(v, Ca[M.vec_src[q]]);
      }
      b1[p] = v;
    }
  }
  // the rest of the panel: the pivot rows right of the pivot block, then all structure-row blocks
  auto ingest_rest = [&](int t0, int tstep) { if (!ABL(64)) ingest_range(w * w, nT + nL, t0, tstep); };
  // ---- panel factorisation, step 1: the pivot block (w <= 16 bus blocks = at most 32 x 32 scalars) is
  // factored entirely in the registers of warp 0: lane = scalar column, a[i] = scalar row i of that
  // column; per 2x2 pivot the inverse is formed once, the two pivot rows are scaled in place
  // (U' = D^-1 U), and the multipliers (kept unscaled in lanes 2p, 2p+1) are broadcast with
  // __shfl_sync for the trailing update - no shared-memory round trips on the serial chain.
  double2* Dv0 = b1 + w;
  double2* Dv1 = Dv0 + w;
#ifdef SBLX_CPASYNC
  // every thread (warp 0 too: the copies are asynchronous) queues its share of the rest of the panel behind the
  // pivot block, then waits for the pivot-block group only
  if (NT > 32) { ingest_rest(tid, NT); cp_async_commit(); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
  front_sync();
  PHASE_MARK(0);
#else
  front_sync();
  PHASE_MARK(0);   // ingest of the pivot block
  if (NT > 32 && tid >= 32) ingest_rest(tid - 32, NT - 32);
#endif
  if (NT == 32 || tid < 32) {
    const int lane = tid;                        // lane within the scenario's group: shuffles below use width G
    const int bc = lane >> 1, hi = lane & 1;
    constexpr int SW = G < 32 ? G : 32;          // shuffle width
    constexpr int WR = G < 32 ? G / 2 : (NT <= 64 && !WIDE) ? 8 : 16;   // register column (host class rule: w <= WR)
    double a[2 * WR];
#pragma unroll
    for (int bi = 0; bi < WR; ++bi) {
      double v0 = 0.0, v1 = 0.0;
      if (bi < w && bc < w) {
        const double2 t0 = T0[bi * nf + bc], t1 = T1[bi * nf + bc];
        v0 = hi ? t0.y : t0.x;
        v1 = hi ? t1.y : t1.x;
      }
      a[2 * bi] = v0;
      a[2 * bi + 1] = v1;
    }
    bool sing = false;
    int ntiny = 0;
    const unsigned full = 0xffffffffu;
    // Compact loop (fits the instruction cache): after pivot p the register column is shifted down by two
    // rows, so the pivot rows are always a[0], a[1] and every index below is a compile-time constant.
    // Each step every lane stores its two pivot-row entries: scaled U' right of the pivot, the (already
    // final) multipliers left of it.
#pragma unroll 1
    for (int p = 0; p < (ABL(512) ? 0 : w); ++p) {
      const int l0 = 2 * p, l1 = 2 * p + 1;
      __syncwarp(full);      // lanes may still be diverged after the data-dependent ingest loops / predicated stores
      const double p00 = __shfl_sync(full, a[0], l0, SW), p01 = __shfl_sync(full, a[0], l1, SW);
      const double p10 = __shfl_sync(full, a[1], l0, SW), p11 = __shfl_sync(full, a[1], l1, SW);
      // det = p00*p11 - p01*p10 with the rounding error of the product recovered by FMA (Kahan)
      const double wq = p01 * p10;
      const double det = fma(p00, p11, -wq) + fma(-p01, p10, wq);
      const bool ok = isfinite(det) && det != 0.0;
      sing |= !ok;
      const double pmx = fmax(fmax(fabs(p00), fabs(p01)), fmax(fabs(p10), fabs(p11)));
      ntiny += (fabs(det) <= M.pivot_tol * pmx * pmx) ? 1 : 0;     // static pivoting monitor: block nearly singular relative to its own scale
      const double sc = ok ? 1.0 / det : 0.0;
      const double d00 = p11 * sc, d01 = -p01 * sc, d10 = -p10 * sc, d11 = p00 * sc;
      if (lane == 0) { Dv0[p] = make_double2(d00, d01); Dv1[p] = make_double2(d10, d11); }
      const bool right = lane > l1;
      double u0 = 0.0, u1 = 0.0;
      if (right) {
        u0 = d00 * a[0] + d01 * a[1];
        u1 = d10 * a[0] + d11 * a[1];
        a[0] = u0;
        a[1] = u1;
      }
      if (bc < w) {
        reinterpret_cast<double*>(&T0[p * nf + bc])[hi] = a[0];
        reinterpret_cast<double*>(&T1[p * nf + bc])[hi] = a[1];
      }
      const int rem = 2 * (w - p);          // rows still alive in the register column (incl. the two pivot rows)
      __syncwarp(full);
#pragma unroll
      for (int j = 2; j < 2 * WR; ++j) {
        if (j < rem) {
          const double m0 = __shfl_sync(full, a[j], l0, SW), m1 = __shfl_sync(full, a[j], l1, SW);
          if (right) a[j] = fma(-m1, u1, fma(-m0, u0, a[j]));
        }
      }
#pragma unroll
      for (int j = 0; j < 2 * WR - 2; ++j) a[j] = a[j + 2];
      a[2 * WR - 2] = 0.0;
      a[2 * WR - 1] = 0.0;
    }
    if (sing && lane == 0) atomicOr(&A.singular[slot], 1);
    if (ntiny && lane == 0) { atomicAdd(&A.tiny[slot], ntiny); A.tinyit[slot] = 1; }
    PHASE_MARK(1);   // pivot-block chain (warp 0)
    if (NT == 32) ingest_rest(tid, G);
  }
#ifdef SBLX_CPASYNC
  cp_async_commit();
  cp_async_wait<0>();
#endif
  front_sync();
  PHASE_MARK(2);   // barrier after the chain: the other warps' share of the ingest
  // ---- step 2: independent substitution tasks, no barriers: U'12 columns (and the rhs) against the
  // multipliers of the pivot block, L21 rows against the scaled pivot rows.
  {
    constexpr int TP = G < 32 ? G : 32;             // task padding: a warp (or lane group) never mixes task kinds
    const int hpad = (h + 1 + TP - 1) & ~(TP - 1);  // column tasks: h structure columns + the rhs
    for (int t = tid; t < (ABL(128) ? 0 : 2 * hpad); t += G) {
      if (t < hpad) {
        const int b = t;
        if (b < h) {
          double2* c0 = T0 + w + b;
          double2* c1 = T1 + w + b;
          for (int p = 0; p < w; ++p) {
            double2 x0 = c0[p * nf], x1 = c1[p * nf];
            const double2* m0 = T0 + p * nf;
            const double2* m1 = T1 + p * nf;
            for (int q = 0; q < p; ++q) bmsub(x0, x1, m0[q], m1[q], c0[q * nf], c1[q * nf]);
            const double2 d0 = Dv0[p], d1 = Dv1[p];
            c0[p * nf] = make_double2(d0.x * x0.x + d0.y * x1.x, d0.x * x0.y + d0.y * x1.y);
            c1[p * nf] = make_double2(d1.x * x0.x + d1.y * x1.x, d1.x * x0.y + d1.y * x1.y);
          }
        } else if (b == h) {                         // forward elimination of the pivots' rhs: y' = D^-1 (b - M y')
          for (int p = 0; p < w; ++p) {
            double2 v = b1[p];
            const double2* m0 = T0 + p * nf;
            const double2* m1 = T1 + p * nf;
            for (int q = 0; q < p; ++q) {
              const double2 y = b1[q];
              v.x = fma(-m0[q].x, y.x, v.x); v.x = fma(-m0[q].y, y.y, v.x);
              v.y = fma(-m1[q].x, y.x, v.y); v.y = fma(-m1[q].y, y.y, v.y);
            }
            const double2 d0 = Dv0[p], d1 = Dv1[p];
            b1[p] = make_double2(d0.x * v.x + d0.y * v.y, d1.x * v.x + d1.y * v.y);
          }
        }
      } else {
        const int ar = t - hpad;
        if (ar < h) {
          double2* r0 = L0 + ar;
          double2* r1 = L1 + ar;
          for (int p = 1; p < w; ++p) {
            double2 x0 = r0[p * h], x1 = r1[p * h];
            for (int q = 0; q < p; ++q) bmsub(x0, x1, r0[q * h], r1[q * h], T0[q * nf + p], T1[q * nf + p]);
            r0[p * h] = x0;
            r1[p * h] = x1;
          }
        }
      }
    }
  }
  front_sync();
  PHASE_MARK(3);   // substitution tasks
  // ---- store U' panel and y'
  {
    double2* U = reinterpret_cast<double2*>(A.U + ((size_t)slot * M.u_blocks + P.u_off) * 4);
    SBLX_CHECK(P.u_off >= 0 && P.u_off + nT <= M.u_blocks && P.first >= 0 && P.first + w <= M.m && slot >= 0 && slot < M.W);
    for (int i = tid; i < (ABL(256) ? 0 : nT); i += G) stg_blk(U + 2 * i, T0[i], T1[i]);
    double2* y = A.yv + (size_t)slot * M.m + P.first;
    for (int p = tid; p < w; p += G) y[p] = b1[p];
  }
  PHASE_MARK(4);   // U store issue
  if (h == 0) return;
  double2* Co = reinterpret_cast<double2*>(Cs + P.c_off * 4);
  SBLX_CHECK(P.c_off >= 0 && P.c_off + (long long)h * h + (h + 1) / 2 <= M.c_blocks);
  // ---- Schur complement straight to HBM: C = gather(children) - L * U'
#ifdef SBLX_SCHUR_EARLY
  if (NT >= 128) {
    // Variant: warp-level work items of 4 structure rows x 32 structure columns, and the FIRST child's blocks of the item
    // are requested BEFORE the FMA loop (into registers that nothing touches until the loop is done), so that their HBM
    // round trip overlaps the arithmetic instead of following it.  16 accumulators + 16 prefetch registers (double2 x 8
    // each) instead of 32 + 32 at the gather; per pivot 2 U' loads (consecutive words) + 8 L broadcasts for 32 DFMA.
    const int lane = tid & 31, warp = tid >> 5, nwarp = NT / 32;
    const int ncg = (h + 31) >> 5, nrg = (h + 3) >> 2;
    const int nch = ABL(16) ? 0 : (int)P.nchild;
    for (int item = warp; item < nrg * ncg; item += nwarp) {
      const int rg = item / ncg, cg = item - rg * ncg;
      const int a0 = rg * 4;
      const int bA = cg * 32 + lane;
      const bool vA = bA < h;
      const int cA = vA ? bA : h - 1;
      int ar[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) ar[i] = min(a0 + i, h - 1);
      double2 g[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) { g[i][0] = make_double2(0.0, 0.0); g[i][1] = g[i][0]; }
      if (nch > 0) {
        const ChildDev Cd = M.childdev[P.child_ptr];
        const int hc = Cd.hc;
        const int* inv = M.inv + Cd.inv_off;
        const double2* Cm = reinterpret_cast<const double2*>(Cs + Cd.c_off * 4);
        const int ibA = vA ? inv[bA] : -1;
        SBLX_CHECK(ibA < hc && Cd.c_off >= 0 && Cd.c_off + (long long)hc * hc <= M.c_blocks);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ia = (a0 + i < h) ? inv[a0 + i] : -1;
          SBLX_CHECK(ia < hc);
          if (ia >= 0 && ibA >= 0) ldg_blk_early(Cm + 2 * ((size_t)ia * hc + ibA), g[i][0], g[i][1]);
        }
      }
      double2 acc[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[i][0] = make_double2(0.0, 0.0); acc[i][1] = acc[i][0]; }
      const double2* u0p = T0 + w + cA;
      const double2* u1p = T1 + w + cA;
#pragma unroll 4
      for (int p = 0; p < (ABL(8) ? 0 : w); ++p) {
        const double2 ua0 = u0p[p * nf], ua1 = u1p[p * nf];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double2 l0 = L0[p * h + ar[i]], l1 = L1[p * h + ar[i]];
          bmsub(acc[i][0], acc[i][1], l0, l1, ua0, ua1);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[i][0] = cadd(acc[i][0], g[i][0]); acc[i][1] = cadd(acc[i][1], g[i][1]); }
      for (int ci = 1; ci < nch; ++ci) {
        const ChildDev Cd = M.childdev[P.child_ptr + ci];
        const int hc = Cd.hc;
        const int* inv = M.inv + Cd.inv_off;
        const double2* Cm = reinterpret_cast<const double2*>(Cs + Cd.c_off * 4);
        const int ibA = vA ? inv[bA] : -1;
        SBLX_CHECK(ibA < hc && Cd.c_off >= 0 && Cd.c_off + (long long)hc * hc <= M.c_blocks);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ia = (a0 + i < h) ? inv[a0 + i] : -1;
          g[i][0] = make_double2(0.0, 0.0); g[i][1] = g[i][0];
          if (ia >= 0 && ibA >= 0) ldg_blk(Cm + 2 * ((size_t)ia * hc + ibA), g[i][0], g[i][1]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i][0] = cadd(acc[i][0], g[i][0]); acc[i][1] = cadd(acc[i][1], g[i][1]); }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (a0 + i >= h) break;
        if (ABL(32) && acc[i][0].x != 123.456) continue;
        if (vA) stg_blk(Co + 2 * ((a0 + i) * h + bA), acc[i][0], acc[i][1]);
      }
    }
  } else
#endif
  if (NT >= 128) {
    // warp-level work items: 4 structure rows x 64 structure columns; lane l owns columns (cg*64 + l) and
    // (cg*64 + 32 + l): L loads are warp broadcasts, U' loads are consecutive 16 B words (no bank conflicts),
    // 32 accumulators per thread (64 DFMA per 12 LDS.128).
    const int lane = tid & 31, warp = tid >> 5, nwarp = NT / 32;
    const int ncg = (h + 63) >> 6, nrg = (h + 3) >> 2;
    for (int item = warp; item < nrg * ncg; item += nwarp) {
      const int rg = item / ncg, cg = item - rg * ncg;
      const int a0 = rg * 4;
      const int bA = cg * 64 + lane, bB = bA + 32;
      const bool vA = bA < h, vB = bB < h;
      const int cA = vA ? bA : h - 1, cB = vB ? bB : h - 1;
      int ar[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) ar[i] = min(a0 + i, h - 1);
      double2 acc[4][2][2];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) { acc[i][j][0] = make_double2(0.0, 0.0); acc[i][j][1] = make_double2(0.0, 0.0); }
      const double2* u0p = T0 + w + cA;
      const double2* u1p = T1 + w + cA;
      const int dB = cB - cA;
#pragma unroll 2
      for (int p = 0; p < (ABL(8) ? 0 : w); ++p) {
        const double2 ua0 = u0p[p * nf], ua1 = u1p[p * nf];
        const double2 ub0 = u0p[p * nf + dB], ub1 = u1p[p * nf + dB];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double2 l0 = L0[p * h + ar[i]], l1 = L1[p * h + ar[i]];
          bmsub(acc[i][0][0], acc[i][0][1], l0, l1, ua0, ua1);
          bmsub(acc[i][1][0], acc[i][1][1], l0, l1, ub0, ub1);
        }
      }
      for (int ci = 0; ci < (ABL(16) ? 0 : (int)P.nchild); ++ci) {
        const ChildDev Cd = M.childdev[P.child_ptr + ci];
        const int hc = Cd.hc;
        const int* inv = M.inv + Cd.inv_off;
        const double2* Cm = reinterpret_cast<const double2*>(Cs + Cd.c_off * 4);
        const int ibA = vA ? inv[bA] : -1, ibB = vB ? inv[bB] : -1;
        SBLX_CHECK(ibA < hc && ibB < hc && Cd.c_off >= 0 && Cd.c_off + (long long)hc * hc <= M.c_blocks);
        // all (predicated) loads are issued before any accumulation: 16 independent 16 B loads in flight
        double2 g[4][2][2];
        int ia[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ia[i] = (a0 + i < h) ? inv[a0 + i] : -1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          SBLX_CHECK(ia[i] < hc);
          const bool okA = ia[i] >= 0 && ibA >= 0, okB = ia[i] >= 0 && ibB >= 0;
          const double2* spA = Cm + 2 * ((size_t)(okA ? ia[i] : 0) * hc + (okA ? ibA : 0));
          const double2* spB = Cm + 2 * ((size_t)(okB ? ia[i] : 0) * hc + (okB ? ibB : 0));
          const double2 z = make_double2(0.0, 0.0);
          g[i][0][0] = z; g[i][0][1] = z; g[i][1][0] = z; g[i][1][1] = z;
          if (okA) ldg_blk(spA, g[i][0][0], g[i][0][1]);
          if (okB) ldg_blk(spB, g[i][1][0], g[i][1][1]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i][0][0] = cadd(acc[i][0][0], g[i][0][0]); acc[i][0][1] = cadd(acc[i][0][1], g[i][0][1]);
          acc[i][1][0] = cadd(acc[i][1][0], g[i][1][0]); acc[i][1][1] = cadd(acc[i][1][1], g[i][1][1]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (a0 + i >= h) break;
        if (ABL(32) && acc[i][0][0].x != 123.456) continue;       // keep the accumulators live without storing
        if (vA) stg_blk(Co + 2 * ((a0 + i) * h + bA), acc[i][0][0], acc[i][0][1]);
        if (vB) stg_blk(Co + 2 * ((a0 + i) * h + bB), acc[i][1][0], acc[i][1][1]);
      }
    }
  } else {
    // small fronts: one 2x2-block register tile per thread
    const int th = (h + 1) >> 1;
    for (int tile = tid; tile < th * th; tile += G) {
      const int a0 = (tile / th) * 2, b0 = (tile % th) * 2;
      const bool va = a0 + 1 < h, vb = b0 + 1 < h;
      double2 acc[2][2][2];
#pragma unroll
      for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int y = 0; y < 2; ++y) { acc[x][y][0] = make_double2(0.0, 0.0); acc[x][y][1] = make_double2(0.0, 0.0); }
      const int a1i = va ? a0 + 1 : a0, b1i = vb ? b0 + 1 : b0;
#ifdef SBLX_SCHUR_EARLY_SMALL
      // the first child's four blocks are requested before the FMA loop (see the large-front variant)
      double2 g[2][2][2];
#pragma unroll
      for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int y = 0; y < 2; ++y) { g[x][y][0] = make_double2(0.0, 0.0); g[x][y][1] = g[x][y][0]; }
      if (P.nchild > 0) {
        const ChildDev Cd = M.childdev[P.child_ptr];
        const int hc = Cd.hc;
        const int* inv = M.inv + Cd.inv_off;
        const double2* Cm = reinterpret_cast<const double2*>(Cs + Cd.c_off * 4);
        const int ia0 = inv[a0], ia1 = va ? inv[a0 + 1] : -1;
        const int ib0 = inv[b0], ib1 = vb ? inv[b0 + 1] : -1;
        if (ia0 >= 0 && ib0 >= 0) ldg_blk_early(Cm + 2 * (ia0 * hc + ib0), g[0][0][0], g[0][0][1]);
        if (ia0 >= 0 && ib1 >= 0) ldg_blk_early(Cm + 2 * (ia0 * hc + ib1), g[0][1][0], g[0][1][1]);
        if (ia1 >= 0 && ib0 >= 0) ldg_blk_early(Cm + 2 * (ia1 * hc + ib0), g[1][0][0], g[1][0][1]);
        if (ia1 >= 0 && ib1 >= 0) ldg_blk_early(Cm + 2 * (ia1 * hc + ib1), g[1][1][0], g[1][1][1]);
      }
#endif
      for (int p = 0; p < w; ++p) {
        const double2 la0 = L0[p * h + a0], la1 = L1[p * h + a0];
        const double2 lb0 = L0[p * h + a1i], lb1 = L1[p * h + a1i];
        const double2 ua0 = T0[p * nf + w + b0], ua1 = T1[p * nf + w + b0];
        const double2 ub0 = T0[p * nf + w + b1i], ub1 = T1[p * nf + w + b1i];
        bmsub(acc[0][0][0], acc[0][0][1], la0, la1, ua0, ua1);
        bmsub(acc[0][1][0], acc[0][1][1], la0, la1, ub0, ub1);
        bmsub(acc[1][0][0], acc[1][0][1], lb0, lb1, ua0, ua1);
        bmsub(acc[1][1][0], acc[1][1][1], lb0, lb1, ub0, ub1);
      }
#ifdef SBLX_SCHUR_EARLY_SMALL
#pragma unroll
      for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int y = 0; y < 2; ++y) { acc[x][y][0] = cadd(acc[x][y][0], g[x][y][0]); acc[x][y][1] = cadd(acc[x][y][1], g[x][y][1]); }
      for (int ci = 1; ci < P.nchild; ++ci) {
#else
      for (int ci = 0; ci < P.nchild; ++ci) {
#endif
        const ChildDev Cd = M.childdev[P.child_ptr + ci];
        const int hc = Cd.hc;
        const int* inv = M.inv + Cd.inv_off;
        const double2* Cm = reinterpret_cast<const double2*>(Cs + Cd.c_off * 4);
        const int ia0 = inv[a0], ia1 = va ? inv[a0 + 1] : -1;
        const int ib0 = inv[b0], ib1 = vb ? inv[b0 + 1] : -1;
        SBLX_CHECK(ia0 < hc && ia1 < hc && ib0 < hc && ib1 < hc && Cd.c_off >= 0 && Cd.c_off + (long long)hc * hc <= M.c_blocks);
        if (ia0 >= 0 && ib0 >= 0) { double2 s0, s1; ldg_blk(Cm + 2 * (ia0 * hc + ib0), s0, s1); acc[0][0][0] = cadd(acc[0][0][0], s0); acc[0][0][1] = cadd(acc[0][0][1], s1); }
        if (ia0 >= 0 && ib1 >= 0) { double2 s0, s1; ldg_blk(Cm + 2 * (ia0 * hc + ib1), s0, s1); acc[0][1][0] = cadd(acc[0][1][0], s0); acc[0][1][1] = cadd(acc[0][1][1], s1); }
        if (ia1 >= 0 && ib0 >= 0) { double2 s0, s1; ldg_blk(Cm + 2 * (ia1 * hc + ib0), s0, s1); acc[1][0][0] = cadd(acc[1][0][0], s0); acc[1][0][1] = cadd(acc[1][0][1], s1); }
        if (ia1 >= 0 && ib1 >= 0) { double2 s0, s1; ldg_blk(Cm + 2 * (ia1 * hc + ib1), s0, s1); acc[1][1][0] = cadd(acc[1][1][0], s0); acc[1][1][1] = cadd(acc[1][1][1], s1); }
      }
      stg_blk(Co + 2 * (a0 * h + b0), acc[0][0][0], acc[0][0][1]);
      if (vb) stg_blk(Co + 2 * (a0 * h + b0 + 1), acc[0][1][0], acc[0][1][1]);
      if (va) {
        stg_blk(Co + 2 * ((a0 + 1) * h + b0), acc[1][0][0], acc[1][0][1]);
        if (vb) stg_blk(Co + 2 * ((a0 + 1) * h + b0 + 1), acc[1][1][0], acc[1][1][1]);
      }
    }
  }
  PHASE_MARK(5);   // Schur (thread 0's share)
  // ---- update vector: cv[a] = gather(children) - L[a][:] y'
  double2* cvo = Co + 2 * (size_t)h * h;
  for (int a = tid; a < h; a += G) {
    double2 v = make_double2(0.0, 0.0);
    for (int p = 0; p < w; ++p) {
      const double2 l0 = L0[p * h + a], l1 = L1[p * h + a], y = b1[p];
      v.x -= l0.x * y.x + l0.y * y.y;
      v.y -= l1.x * y.x + l1.y * y.y;
    }
    for (int ci = 0; ci < P.nchild; ++ci) {
      const ChildDev Cd = M.childdev[P.child_ptr + ci];
      const int ia = M.inv[Cd.inv_off + a];
      if (ia >= 0) {
        const double2* cvc = reinterpret_cast<const double2*>(Cs + Cd.c_off * 4) + 2 * (size_t)Cd.hc * Cd.hc;
        v = cadd(v, cvc[ia]);
      }
    }
    cvo[a] = v;
  }
  PHASE_MARK(6);
#ifdef SBLX_PHASE_TIMING
  front_sync();
  PHASE_MARK(7);   // tail: slowest warp
#endif
#ifdef SBLX_PHASE_TIMING
  if (threadIdx.x == 0) atomicAdd(&g_phase_cycles[NT == 32 ? 0 : NT == 64 ? 1 : NT == 128 ? 2 : NT == 256 ? 3 : 4][9], 1ull);
#endif
}

// ------------------------------------------------------------------------------------------
// K3r: tiny fronts (w + h <= NF = 8 buses) factored ENTIRELY IN REGISTERS: 8 lanes per scenario, 4 scenarios per
// one-warp CTA (a 16-lane / 12-bus instantiation needs 168 registers and measured slower than the shared-memory
// lane-group kernel).  Lane r holds block row r of the full front (pivot rows, structure rows incl. the Schur part) and
// its rhs / update-vector entry.  Per pivot: the 2x2 pivot block is broadcast, its owner scales its row, the
// scaled row is broadcast block by block and the lanes below update theirs - no shared memory, no barriers.
// Same recurrences as k_front (U' = D^-1 U, unscaled multipliers, y' = D^-1 (b - M y')).  On sparse
// transmission grids most fronts of a minimum-degree tree are in this class.
// ------------------------------------------------------------------------------------------
template <int G, int NF>
__global__ void __launch_bounds__(32, (NF <= 8 ? 16 : 10)) k_front_reg(DevModel M, Slots A, Lists L, const PanelDev* __restrict__ group_panels) {
  constexpr int NG = 32 / G;
  const int nact = L.counts[2];
  if ((int)blockIdx.x * NG >= nact) return;
  const int bscen = min((int)blockIdx.x * NG + (int)threadIdx.x / G, nact - 1);   // past the end: repeat the last one
  const int slot = L.step[bscen];
  const PanelDev P = group_panels[blockIdx.y];
  const int w = P.w, h = P.h, nf = w + h;
  const int r = threadIdx.x % G;
  const bool mine = r < nf;
  const unsigned full = 0xffffffffu;
  double* Cs = A.C + (size_t)slot * M.c_blocks * 4;
  const double2* Jv = reinterpret_cast<const double2*>(A.Jv + (size_t)slot * M.nnzB * 4);
  const double2* Ca = reinterpret_cast<const double2*>(Cs);
  double2 a0[NF], a1[NF];
  const int* words = M.reg_words + P.reg_base + r * nf;
#pragma unroll
  for (int c = 0; c < NF; ++c) {
    a0[c] = make_double2(0.0, 0.0);
    a1[c] = a0[c];
    if (c < nf && mine) {
      const int v = words[c];
      SBLX_CHECK(v == ING_NONE || v >= ING_MULTI || (v < 0 ? (~v) < M.nnzB : (long long)v < M.c_blocks));
      if (v != ING_NONE) {
        if (v < ING_MULTI) {
          ldg_blk(v < 0 ? Jv + 2 * (size_t)(~v) : Ca + 2 * (size_t)v, a0[c], a1[c]);
        } else {
          const int x = v - ING_MULTI;
          for (int q = M.ing_xptr[x]; q < M.ing_xptr[x + 1]; ++q) {
            const int sidx = M.ing_xsrc[q];
            double2 s0, s1;
            ldg_blk(sidx < 0 ? Jv + 2 * (size_t)(~sidx) : Ca + 2 * (size_t)sidx, s0, s1);
            a0[c] = cadd(a0[c], s0); a1[c] = cadd(a1[c], s1);
          }
        }
      }
    }
  }
  double2 b = make_double2(0.0, 0.0);
  if (mine) {
    if (r < w) b = A.rhs[(size_t)slot * M.m + P.first + r];
    const int* vp = M.reg_vptr + P.reg_vbase + r;
    for (int q = vp[0]; q < vp[1]; ++q) b = cadd(b, Ca[M.reg_vsrc[q]]);
  }
  bool sing = false;
  int ntiny = 0;
#pragma unroll
  for (int p = 0; p < NF; ++p) {
    if (p < w) {
      const double p00 = __shfl_sync(full, a0[p].x, p, G), p01 = __shfl_sync(full, a0[p].y, p, G);
      const double p10 = __shfl_sync(full, a1[p].x, p, G), p11 = __shfl_sync(full, a1[p].y, p, G);
      const double wq = p01 * p10;
      const double det = fma(p00, p11, -wq) + fma(-p01, p10, wq);      // Kahan: rounding error of the product recovered
      const bool ok = isfinite(det) && det != 0.0;
      sing |= !ok;
      const double pmx = fmax(fmax(fabs(p00), fabs(p01)), fmax(fabs(p10), fabs(p11)));
      ntiny += (fabs(det) <= M.pivot_tol * pmx * pmx) ? 1 : 0;
      const double sc = ok ? 1.0 / det : 0.0;
      const double d00 = p11 * sc, d01 = -p01 * sc, d10 = -p10 * sc, d11 = p00 * sc;
      if (r == p) {                                   // the pivot's owner scales its row and its rhs
#pragma unroll
        for (int c = 0; c < NF; ++c)
          if (c > p && c < nf) {
            const double2 t0 = a0[c], t1 = a1[c];
            a0[c] = make_double2(d00 * t0.x + d01 * t1.x, d00 * t0.y + d01 * t1.y);
            a1[c] = make_double2(d10 * t0.x + d11 * t1.x, d10 * t0.y + d11 * t1.y);
          }
        b = make_double2(d00 * b.x + d01 * b.y, d10 * b.x + d11 * b.y);
      }
      const bool below = r > p && mine;
      const double2 m0 = a0[p], m1 = a1[p];           // this lane's multiplier block A_rp (unscaled)
      const double yx = __shfl_sync(full, b.x, p, G), yy = __shfl_sync(full, b.y, p, G);
      if (below) {
        b.x = fma(-m0.x, yx, b.x); b.x = fma(-m0.y, yy, b.x);
        b.y = fma(-m1.x, yx, b.y); b.y = fma(-m1.y, yy, b.y);
      }
#pragma unroll
      for (int c = 0; c < NF; ++c)
        if (c > p && c < nf) {
          double2 u0, u1;
          u0.x = __shfl_sync(full, a0[c].x, p, G); u0.y = __shfl_sync(full, a0[c].y, p, G);
          u1.x = __shfl_sync(full, a1[c].x, p, G); u1.y = __shfl_sync(full, a1[c].y, p, G);
          if (below) bmsub(a0[c], a1[c], m0, m1, u0, u1);
        }
    }
  }
  if (sing && r == 0) atomicOr(&A.singular[slot], 1);
  if (ntiny && r == 0) { atomicAdd(&A.tiny[slot], ntiny); A.tinyit[slot] = 1; }
  // ---- outputs: U' rows and y' of the pivots, update matrix + update vector of the structure rows
  if (r < w) {
    double2* U = reinterpret_cast<double2*>(A.U + ((size_t)slot * M.u_blocks + P.u_off) * 4) + 2 * (size_t)(r * nf);
#pragma unroll
    for (int c = 0; c < NF; ++c)
      if (c < nf && c > r) stg_blk(U + 2 * c, a0[c], a1[c]);
    A.yv[(size_t)slot * M.m + P.first + r] = b;
  } else if (mine) {
    double2* Co = reinterpret_cast<double2*>(Cs + P.c_off * 4);
    double2* row = Co + 2 * (size_t)((r - w) * h);
#pragma unroll
    for (int c = 0; c < NF; ++c)
      if (c < nf && c >= w) stg_blk(row + 2 * (c - w), a0[c], a1[c]);
    Co[2 * (size_t)h * h + (r - w)] = b;
  }
}

// ==========================================================================================
// Split factorisation path: three streaming kernels per level instead of one fused CTA per front.
//   k_pivot : one WARP per (front, scenario): gathers the w x w pivot block + rhs straight into
//             registers, factors it (same register/shuffle chain as above), writes the factored
//             block (multipliers + scaled rows), the inverse pivots and y' to HBM.
//   k_subst : one THREAD per U'12 column / L21 row: gathers its column into registers (all loads in
//             flight at once), substitutes against the factored pivot block (shared memory, 8 KB),
//             writes U'12 into the U panel and L21 into a level-local workspace.
//   k_schur : one CTA per 32 x 64 tile of F22: loads the L / U' strips of the tile, register-tile
//             FMA loop, gathers the children's contributions, writes the update matrix once.
// Only small per-CTA shared memory -> many CTAs per SM; every kernel is a coalesced stream.
// ==========================================================================================
__device__ __forceinline__ const double2* ing_source(int sidx, const double2* Jv, const double2* Ca) {
  return sidx < 0 ? Jv + 2 * (size_t)(~sidx) : Ca + 2 * (size_t)sidx;
}

// grid (panels, ceil(nStep/4)), block (32, 4)
__global__ void __launch_bounds__(128) k_pivot(DevModel M, Slots A, Lists L, const int* __restrict__ panel_ids) {
  const int lane = threadIdx.x;
  const int si = blockIdx.y * 4 + threadIdx.y;
  if (si >= L.counts[2]) return;
  const int slot = L.step[si];
  const PanelDev P = M.panels[panel_ids[blockIdx.x]];
  const int w = P.w, nf = P.w + P.h;
  const double2* Jv = reinterpret_cast<const double2*>(A.Jv + (size_t)slot * M.nnzB * 4);
  const double2* Ca = reinterpret_cast<const double2*>(A.C + (size_t)slot * M.c_blocks * 4);
  const int* iptr = M.ing_ptr + P.ing_base;
  const int* isrc = M.ing_src;
  const int bc = lane >> 1, hi = lane & 1;
  const unsigned full = 0xffffffffu;
  constexpr int WR = 16;
  double a[2 * WR];
#pragma unroll
  for (int i = 0; i < 2 * WR; ++i) a[i] = 0.0;
  for (int r = 0; r < 64; ++r) {             // source rounds: all pivot-block loads of one round are independent
    bool any = false;
#pragma unroll
    for (int bi = 0; bi < WR; ++bi) {
      if (bi < w && bc < w) {
        const int k = bi * nf + bc;
        const int s0 = iptr[k] + r;
        if (s0 < iptr[k + 1]) {
          const double2* ptr = ing_source(isrc[s0], Jv, Ca);
          const double2 r0 = ptr[0], r1 = ptr[1];
          a[2 * bi] += hi ? r0.y : r0.x;
          a[2 * bi + 1] += hi ? r1.y : r1.x;
          any = true;
        }
      }
    }
    if (!__any_sync(full, any)) break;
  }
  // rhs: lane i holds scalar row i of the pivots' right-hand side
  double rb = 0.0;
  if (lane < 2 * w) {
    const int p = lane >> 1;
    double2 v = A.rhs[(size_t)slot * M.m + P.first + p];
    const int* vptr = M.vec_ptr + P.first;
    for (int q = vptr[p]; q < vptr[p + 1]; ++q) v = cadd(v, Ca[M.vec_src[q]]);
    rb = (lane & 1) ? v.y : v.x;
  }
  double2* Up = reinterpret_cast<double2*>(A.U + ((size_t)slot * M.u_blocks + P.u_off) * 4);
  double2* Dw = A.Dw + ((size_t)slot * M.m + P.first) * 2;
  bool sing = false;
#pragma unroll 1
  for (int p = 0; p < w; ++p) {
    const int l0 = 2 * p, l1 = 2 * p + 1;
    __syncwarp(full);
    const double p00 = __shfl_sync(full, a[0], l0), p01 = __shfl_sync(full, a[0], l1);
    const double p10 = __shfl_sync(full, a[1], l0), p11 = __shfl_sync(full, a[1], l1);
    const double wq = p01 * p10;
    const double det = fma(p00, p11, -wq) + fma(-p01, p10, wq);
    const bool ok = isfinite(det) && det != 0.0;
    sing |= !ok;
    const double sc = ok ? 1.0 / det : 0.0;
    const double d00 = p11 * sc, d01 = -p01 * sc, d10 = -p10 * sc, d11 = p00 * sc;
    if (lane == 0) { Dw[2 * p] = make_double2(d00, d01); Dw[2 * p + 1] = make_double2(d10, d11); }
    const double yr0 = __shfl_sync(full, rb, l0), yr1 = __shfl_sync(full, rb, l1);
    const double y0 = d00 * yr0 + d01 * yr1, y1 = d10 * yr0 + d11 * yr1;
    if (lane == l0) rb = y0;
    if (lane == l1) rb = y1;
    const bool right = lane > l1;
    double u0 = 0.0, u1 = 0.0;
    if (right) {
      u0 = d00 * a[0] + d01 * a[1];
      u1 = d10 * a[0] + d11 * a[1];
      a[0] = u0;
      a[1] = u1;
    }
    if (bc < w) {
      reinterpret_cast<double*>(&Up[2 * (p * nf + bc)])[hi] = a[0];
      reinterpret_cast<double*>(&Up[2 * (p * nf + bc) + 1])[hi] = a[1];
    }
    const int rem = 2 * (w - p);
    __syncwarp(full);
#pragma unroll
    for (int j = 2; j < 2 * WR; ++j) {
      if (j < rem) {
        const double m0 = __shfl_sync(full, a[j], l0), m1 = __shfl_sync(full, a[j], l1);
        if (right) a[j] = fma(-m1, u1, fma(-m0, u0, a[j]));
        if (lane == l0 + j) rb = fma(-m1, y1, fma(-m0, y0, rb));
      }
    }
#pragma unroll
    for (int j = 0; j < 2 * WR - 2; ++j) a[j] = a[j + 2];
    a[2 * WR - 2] = 0.0;
    a[2 * WR - 1] = 0.0;
  }
  if (lane < 2 * w) reinterpret_cast<double*>(&A.yv[(size_t)slot * M.m + P.first + (lane >> 1)])[lane & 1] = rb;
  if (sing && lane == 0) atomicOr(&A.singular[slot], 1);
}

// grid (panels, nStep, chunks of 128 tasks), block 128
__global__ void __launch_bounds__(128, 3) k_subst(DevModel M, Slots A, Lists L, const int* __restrict__ panel_ids) {
  extern __shared__ double2 sm2[];
  const PanelDev P = M.panels[panel_ids[blockIdx.x]];
  const int w = P.w, h = P.h, nf = w + h;
  const int hpad = (h + 31) & ~31;
  if ((int)blockIdx.z * 128 >= 2 * hpad) return;
  if ((int)blockIdx.y >= L.counts[2]) return;
  const int slot = L.step[blockIdx.y];
  const int tid = threadIdx.x;
  double2* M0 = sm2;            // factored pivot block, block-row planes [w][w]
  double2* M1 = M0 + w * w;
  double2* D0 = M1 + w * w;     // inverse pivots [w]
  double2* D1 = D0 + w;
  double2* Up = reinterpret_cast<double2*>(A.U + ((size_t)slot * M.u_blocks + P.u_off) * 4);
  {
    const double2* Dw = A.Dw + ((size_t)slot * M.m + P.first) * 2;
    for (int i = tid; i < w * w; i += 128) {
      const int r = i / w, c = i - r * w;
      M0[i] = Up[2 * (r * nf + c)];
      M1[i] = Up[2 * (r * nf + c) + 1];
    }
    for (int p = tid; p < w; p += 128) { D0[p] = Dw[2 * p]; D1[p] = Dw[2 * p + 1]; }
  }
  const int t = blockIdx.z * 128 + tid;
  const bool is_col = t < hpad;
  const int idx = is_col ? t : t - hpad;
  const bool live = idx < h;
  const double2* Jv = reinterpret_cast<const double2*>(A.Jv + (size_t)slot * M.nnzB * 4);
  const double2* Ca = reinterpret_cast<const double2*>(A.C + (size_t)slot * M.c_blocks * 4);
  const int* iptr = M.ing_ptr + P.ing_base;
  const int* isrc = M.ing_src;
  constexpr int WR = 16;
  double2 x0[WR], x1[WR];
#pragma unroll
  for (int p = 0; p < WR; ++p) { x0[p] = make_double2(0.0, 0.0); x1[p] = x0[p]; }
  if (live) {
    // panel block of entry p of this task: column task -> (p, w + idx) in the pivot rows, row task -> L[p][idx]
    const int kbase = is_col ? (w + idx) : (w * nf + idx);
    const int kstep = is_col ? nf : h;
    for (int r = 0; r < 64; ++r) {
      bool any = false;
#pragma unroll
      for (int p = 0; p < WR; ++p) {
        if (p < w) {
          const int k = kbase + p * kstep;
          const int s0 = iptr[k] + r;
          if (s0 < iptr[k + 1]) {
            const double2* ptr = ing_source(isrc[s0], Jv, Ca);
            x0[p] = cadd(x0[p], ptr[0]);
            x1[p] = cadd(x1[p], ptr[1]);
            any = true;
          }
        }
      }
      if (!any) break;
    }
  }
  __syncthreads();
  if (!live) return;
  if (is_col) {
    double2* out = Up + 2 * (w + idx);
#pragma unroll 1
    for (int q = 0; q < w; ++q) {
      const double2 d0 = D0[q], d1 = D1[q];
      const double2 u0 = make_double2(d0.x * x0[0].x + d0.y * x1[0].x, d0.x * x0[0].y + d0.y * x1[0].y);
      const double2 u1 = make_double2(d1.x * x0[0].x + d1.y * x1[0].x, d1.x * x0[0].y + d1.y * x1[0].y);
      out[2 * (q * nf)] = u0;
      out[2 * (q * nf) + 1] = u1;
#pragma unroll
      for (int j = 1; j < WR; ++j) {
        if (q + j < w) bmsub(x0[j], x1[j], M0[(q + j) * w + q], M1[(q + j) * w + q], u0, u1);
        x0[j - 1] = x0[j];
        x1[j - 1] = x1[j];
      }
    }
  } else {
    double2* out = reinterpret_cast<double2*>(A.Lw + ((size_t)slot * M.l_blocks + P.l_off) * 4) + 2 * idx;
#pragma unroll 1
    for (int q = 0; q < w; ++q) {
      const double2 l0 = x0[0], l1 = x1[0];
      out[2 * (q * h)] = l0;
      out[2 * (q * h) + 1] = l1;
#pragma unroll
      for (int j = 1; j < WR; ++j) {
        if (q + j < w) bmsub(x0[j], x1[j], l0, l1, M0[q * w + q + j], M1[q * w + q + j]);
        x0[j - 1] = x0[j];
        x1[j - 1] = x1[j];
      }
    }
  }
}

// grid (panels, nStep, tiles), block 128: tile = 32 structure rows x 64 structure columns of F22
__global__ void __launch_bounds__(128, 4) k_schur(DevModel M, Slots A, Lists L, const int* __restrict__ panel_ids) {
  extern __shared__ double2 sm2[];
  const PanelDev P = M.panels[panel_ids[blockIdx.x]];
  const int w = P.w, h = P.h, nf = w + h;
  if (h == 0) return;
  const int ntc = (h + 63) >> 6, ntr = (h + 31) >> 5;
  if ((int)blockIdx.z >= ntr * ntc) return;
  if ((int)blockIdx.y >= L.counts[2]) return;
  const int slot = L.step[blockIdx.y];
  const int tr = blockIdx.z / ntc, tc = blockIdx.z - tr * ntc;
  const int a_base = tr * 32, b_base = tc * 64;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double2* L0 = sm2;              // [w][32]
  double2* L1 = L0 + w * 32;
  double2* U0 = L1 + w * 32;      // [w][64]
  double2* U1 = U0 + w * 64;
  const double2* Lg = reinterpret_cast<const double2*>(A.Lw + ((size_t)slot * M.l_blocks + P.l_off) * 4);
  const double2* Ug = reinterpret_cast<const double2*>(A.U + ((size_t)slot * M.u_blocks + P.u_off) * 4);
  for (int i = tid; i < w * 32; i += 128) {
    const int p = i >> 5, r = i & 31;
    const int a = min(a_base + r, h - 1);
    L0[i] = Lg[2 * (p * h + a)];
    L1[i] = Lg[2 * (p * h + a) + 1];
  }
  for (int i = tid; i < w * 64; i += 128) {
    const int p = i >> 6, c = i & 63;
    const int b = min(b_base + c, h - 1);
    U0[i] = Ug[2 * (p * nf + w + b)];
    U1[i] = Ug[2 * (p * nf + w + b) + 1];
  }
  __syncthreads();
  double* Cs = A.C + (size_t)slot * M.c_blocks * 4;
  double2* Co = reinterpret_cast<double2*>(Cs + P.c_off * 4);
  const int bA = b_base + lane, bB = bA + 32;
  const bool vA = bA < h, vB = bB < h;
  for (int rg = warp; rg < 8; rg += 4) {
    const int a0 = a_base + rg * 4;
    if (a0 >= h) break;
    double2 acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) { acc[i][j][0] = make_double2(0.0, 0.0); acc[i][j][1] = make_double2(0.0, 0.0); }
    const double2* lp0 = L0 + rg * 4;
    const double2* lp1 = L1 + rg * 4;
    const double2* up0 = U0 + lane;
    const double2* up1 = U1 + lane;
#pragma unroll 2
    for (int p = 0; p < w; ++p) {
      const double2 ua0 = up0[p * 64], ua1 = up1[p * 64];
      const double2 ub0 = up0[p * 64 + 32], ub1 = up1[p * 64 + 32];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double2 l0 = lp0[p * 32 + i], l1 = lp1[p * 32 + i];
        bmsub(acc[i][0][0], acc[i][0][1], l0, l1, ua0, ua1);
        bmsub(acc[i][1][0], acc[i][1][1], l0, l1, ub0, ub1);
      }
    }
    for (int ci = 0; ci < P.nchild; ++ci) {
      const ChildDev Cd = M.childdev[P.child_ptr + ci];
      const int hc = Cd.hc;
      const int* inv = M.inv + Cd.inv_off;
      const double2* Cm = reinterpret_cast<const double2*>(Cs + Cd.c_off * 4);
      const int ibA = vA ? inv[bA] : -1, ibB = vB ? inv[bB] : -1;
      double2 g[4][2][2];
      int ia[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) ia[i] = (a0 + i < h) ? inv[a0 + i] : -1;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool okA = ia[i] >= 0 && ibA >= 0, okB = ia[i] >= 0 && ibB >= 0;
        const double2* spA = Cm + 2 * ((size_t)(okA ? ia[i] : 0) * hc + (okA ? ibA : 0));
        const double2* spB = Cm + 2 * ((size_t)(okB ? ia[i] : 0) * hc + (okB ? ibB : 0));
        const double2 z = make_double2(0.0, 0.0);
        g[i][0][0] = okA ? spA[0] : z; g[i][0][1] = okA ? spA[1] : z;
        g[i][1][0] = okB ? spB[0] : z; g[i][1][1] = okB ? spB[1] : z;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0][0] = cadd(acc[i][0][0], g[i][0][0]); acc[i][0][1] = cadd(acc[i][0][1], g[i][0][1]);
        acc[i][1][0] = cadd(acc[i][1][0], g[i][1][0]); acc[i][1][1] = cadd(acc[i][1][1], g[i][1][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (a0 + i >= h) break;
      if (vA) { Co[2 * ((size_t)(a0 + i) * h + bA)] = acc[i][0][0]; Co[2 * ((size_t)(a0 + i) * h + bA) + 1] = acc[i][0][1]; }
      if (vB) { Co[2 * ((size_t)(a0 + i) * h + bB)] = acc[i][1][0]; Co[2 * ((size_t)(a0 + i) * h + bB) + 1] = acc[i][1][1]; }
    }
  }
  // update vector of the tile's rows (first tile column only): cv[a] = gather(children) - L[a][:] y'
  if (tc == 0 && tid < 32) {
    const int a = a_base + tid;
    if (a < h) {
      const double2* y = A.yv + (size_t)slot * M.m + P.first;
      double2 v = make_double2(0.0, 0.0);
      for (int p = 0; p < w; ++p) {
        const double2 l0 = L0[p * 32 + tid], l1 = L1[p * 32 + tid], yy = y[p];
        v.x = fma(-l0.x, yy.x, v.x); v.x = fma(-l0.y, yy.y, v.x);
        v.y = fma(-l1.x, yy.x, v.y); v.y = fma(-l1.y, yy.y, v.y);
      }
      for (int ci = 0; ci < P.nchild; ++ci) {
        const ChildDev Cd = M.childdev[P.child_ptr + ci];
        const int ia = M.inv[Cd.inv_off + a];
        if (ia >= 0) {
          const double2* cvc = reinterpret_cast<const double2*>(Cs + Cd.c_off * 4) + 2 * (size_t)Cd.hc * Cd.hc;
          v = cadd(v, cvc[ia]);
        }
      }
      Co[2 * (size_t)h * h + a] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// K4: back substitution of one front: x_P = y' - U'_PP x_P(upper) - U'_PR x_R.  w <= 32.
// ------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) k_backsolve(DevModel M, Slots A, Lists L, const PanelDev* __restrict__ group_panels) {
  extern __shared__ double2 sm2[];
  if ((int)blockIdx.x >= L.counts[2]) return;          // grid: x = scenario (fastest), y = front
  const int slot = L.step[blockIdx.x];
  const PanelDev P = group_panels[blockIdx.y];
  const int w = P.w, h = P.h, nf = w + h;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double2* xs = sm2;        // [h]
  double2* t = xs + h;      // [w]
  double2* xv = A.xv + (size_t)slot * M.m;
  const double2* U = reinterpret_cast<const double2*>(A.U + ((size_t)slot * M.u_blocks + P.u_off) * 4);
  const double2* y = A.yv + (size_t)slot * M.m + P.first;
  const int* sn = M.struct_nodes + P.struct_ptr;
  for (int b = tid; b < h; b += NT) {
    SBLX_CHECK(sn[b] >= 0 && sn[b] < M.m);
    xs[b] = xv[sn[b]];
  }
  // the strictly upper part of the pivot block goes to shared memory now: the serial back substitution below must
  // not pay one HBM round trip per pivot
  double2* Ub = t + w;      // [w*w][2]
  for (int idx = tid; idx < w * w; idx += NT) {
    const int i = idx / w, j = idx - i * w;
    if (j > i) { Ub[2 * idx] = U[2 * (i * nf + j)]; Ub[2 * idx + 1] = U[2 * (i * nf + j) + 1]; }
  }
  __syncthreads();
  for (int p = warp; p < w; p += NT / 32) {
    double ax = 0.0, ay = 0.0;
    for (int b = lane; b < h; b += 32) {
      const double2 u0 = U[2 * (p * nf + w + b)], u1 = U[2 * (p * nf + w + b) + 1], x = xs[b];
      ax += u0.x * x.x + u0.y * x.y;
      ay += u1.x * x.x + u1.y * x.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ax += __shfl_xor_sync(0xffffffffu, ax, o);
      ay += __shfl_xor_sync(0xffffffffu, ay, o);
    }
    if (lane == 0) t[p] = make_double2(y[p].x - ax, y[p].y - ay);
  }
  __syncthreads();
  if (warp == 0) {
    double2 xc = make_double2(0.0, 0.0);
    for (int p = w - 1; p >= 0; --p) {
      double ax = 0.0, ay = 0.0;
      if (lane > p && lane < w) {
        const double2 u0 = Ub[2 * (p * w + lane)], u1 = Ub[2 * (p * w + lane) + 1];
        ax = u0.x * xc.x + u0.y * xc.y;
        ay = u1.x * xc.x + u1.y * xc.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, o);
        ay += __shfl_xor_sync(0xffffffffu, ay, o);
      }
      if (lane == p) xc = make_double2(t[p].x - ax, t[p].y - ay);
    }
    if (lane < w) xv[P.first + lane] = xc;
  }
}

// Back substitution for the tiny fronts (pivot blocks of <= 8 buses, the lane-group classes of k_front): 8 lanes
// per scenario, 4 scenarios per one-warp CTA.  Lane p owns pivot row p: it walks its U' row sequentially (one
// 256-bit load per block, no shuffle reductions), then the w x w triangle is resolved with one broadcast per pivot.
__global__ void __launch_bounds__(32) k_backsolve_small(DevModel M, Slots A, Lists L, const PanelDev* __restrict__ group_panels) {
  constexpr int G = 8, NG = 32 / G;
  const int nact = L.counts[2];
  if ((int)blockIdx.x * NG >= nact) return;
  const int bscen = min((int)blockIdx.x * NG + (int)threadIdx.x / G, nact - 1);   // past the end: repeat the last one
  const int slot = L.step[bscen];
  const PanelDev P = group_panels[blockIdx.y];
  const int w = P.w, h = P.h, nf = w + h;
  const int p = threadIdx.x % G;
  double2* xv = A.xv + (size_t)slot * M.m;
  const double2* U = reinterpret_cast<const double2*>(A.U + ((size_t)slot * M.u_blocks + P.u_off) * 4);
  const int* sn = M.struct_nodes + P.struct_ptr;
  double2 r = make_double2(0.0, 0.0);
  double2 tu0[G], tu1[G];                        // this row's part of the pivot block (columns right of the diagonal)
#pragma unroll
  for (int j = 0; j < G; ++j) { tu0[j] = make_double2(0.0, 0.0); tu1[j] = tu0[j]; }
  if (p < w) {
    r = A.yv[(size_t)slot * M.m + P.first + p];
#pragma unroll
    for (int j = 1; j < G; ++j)
      if (j > p && j < w) ldg_blk(U + 2 * (p * nf + j), tu0[j], tu1[j]);
    const double2* row = U + 2 * (size_t)(p * nf + w);
    for (int b = 0; b < h; ++b) {
      double2 u0, u1;
      ldg_blk(row + 2 * b, u0, u1);
      SBLX_CHECK(sn[b] >= 0 && sn[b] < M.m);
      const double2 x = xv[sn[b]];
      r.x -= u0.x * x.x + u0.y * x.y;
      r.y -= u1.x * x.x + u1.y * x.y;
    }
  }
#pragma unroll
  for (int j = G - 1; j >= 1; --j) {
    // lane j's r is final once every column right of it has been applied
    const double xjx = __shfl_sync(0xffffffffu, r.x, j, G), xjy = __shfl_sync(0xffffffffu, r.y, j, G);
    if (j < w && p < j) {
      r.x -= tu0[j].x * xjx + tu0[j].y * xjy;
      r.y -= tu1[j].x * xjx + tu1[j].y * xjy;
    }
  }
  if (p < w) xv[P.first + p] = r;
}

// Safety net of the static-pivoting LU (the reference's UMFPACK pivots across rows and falls back to QR / SVD,
// solver_core.jl:60-82).  A factorisation that met a tiny pivot block re-checks its own solve: r = b - J x with the stored
// Jacobian blocks.  If |r|_inf > 1e-6 |b|_inf the step is NOT applied as it is: the slot goes to the refine list and the
// host runs iterative refinement on it (second factor + solve pass on the residual, x += dx, at most two passes); a step
// that cannot be repaired ends the scenario as singular_newton_step instead of silently taking a bad step.
// r is left in the y vector (dead after the back substitution).  Block-wide; returns (|r|_inf, |b|_inf) to thread 0.
__device__ __forceinline__ void step_residual(const DevModel& M, const Slots& A, int slot, const double2* b, const double2* x, double2* r,
                                              double (*red)[256], double& mr_out, double& mb_out) {
  const double* Jv = A.Jv + (size_t)slot * M.nnzB * 4;
  for (int e = threadIdx.x; e < M.m; e += 256) r[e] = b[e];
  __syncthreads();
  for (int k = threadIdx.x; k < M.nnzB; k += 256) {
    const int er = M.elim_of_node[M.jb_row[k]], ec = M.elim_of_node[M.jb_col[k]];
    const double2 xc = x[ec];
    const double* q = Jv + 4 * (size_t)k;
    atomicAdd(&r[er].x, -(q[0] * xc.x + q[1] * xc.y));
    atomicAdd(&r[er].y, -(q[2] * xc.x + q[3] * xc.y));
  }
  __syncthreads();
  double mr = 0.0, mb = 0.0;
  bool bad = false;
  for (int e = threadIdx.x; e < M.m; e += 256) {
    const double2 a = r[e], f = b[e];
    bad |= !(isfinite(a.x) && isfinite(a.y));
    mr = fmax(mr, fmax(fabs(a.x), fabs(a.y)));
    mb = fmax(mb, fmax(fabs(f.x), fabs(f.y)));
  }
  red[0][threadIdx.x] = bad ? INFINITY : mr;
  red[1][threadIdx.x] = mb;
  __syncthreads();
  for (int s2 = 128; s2 > 0; s2 >>= 1) {
    if ((int)threadIdx.x < s2) {
      red[0][threadIdx.x] = fmax(red[0][threadIdx.x], red[0][threadIdx.x + s2]);
      red[1][threadIdx.x] = fmax(red[1][threadIdx.x], red[1][threadIdx.x + s2]);
    }
    __syncthreads();
  }
  mr_out = red[0][0];
  mb_out = red[1][0];
  __syncthreads();
}

// grid nStep, block 256: immediately out unless the scenario was flagged in THIS iteration
__global__ void __launch_bounds__(256) k_tiny_guard(DevModel M, Slots A, Lists L) {
  __shared__ double red[2][256];
  if ((int)blockIdx.x >= L.counts[2]) return;
  const int slot = L.step[blockIdx.x];
  if (!A.tinyit[slot] || A.singular[slot]) return;
  double mr, mb;
  step_residual(M, A, slot, A.rhs + (size_t)slot * M.m, A.xv + (size_t)slot * M.m, A.yv + (size_t)slot * M.m, red, mr, mb);
  if (threadIdx.x == 0) {
    const bool force = (M.pf_mode & 1024) != 0;              // test hook: refine every flagged step
    if (force || !(mr <= 1e-6 * fmax(mb, 1e-300))) {
      A.refine[slot] = 1;
      A.fnorm[slot] = mb;
      L.refine[atomicAdd(&L.counts[6], 1)] = slot;
      atomicAdd(&L.counts[7], 1);
    }
  }
}

// refinement pass, before the second factor + solve: keep x, make the residual the new right-hand side.  grid nRefine
__global__ void __launch_bounds__(256) k_refine_prep(DevModel M, Slots A, Lists L) {
  if ((int)blockIdx.x >= L.counts[6]) return;
  const int slot = L.refine[blockIdx.x];
  double2* xs = A.xsave + (size_t)slot * M.m;
  double2* rhs = A.rhs + (size_t)slot * M.m;
  const double2* x = A.xv + (size_t)slot * M.m;
  const double2* r = A.yv + (size_t)slot * M.m;
  for (int e = threadIdx.x; e < M.m; e += 256) { xs[e] = x[e]; rhs[e] = r[e]; }
}

// after it: x = x_saved + dx, r' = r - J dx; accept when |r'| <= 1e-6 |b|, else keep the slot for another pass
// (`last` = no pass left: the step is declared singular).  grid nRefine
__global__ void __launch_bounds__(256) k_refine_post(DevModel M, Slots A, Lists L, int last) {
  __shared__ double red[2][256];
  if ((int)blockIdx.x >= L.counts[6]) return;
  const int slot = L.refine[blockIdx.x];
  if (!A.refine[slot]) return;
  if (A.singular[slot]) { if (threadIdx.x == 0) A.refine[slot] = 0; return; }     // the second factorisation hit an exactly singular block
  double2* x = A.xv + (size_t)slot * M.m;
  double mr, mb;
  step_residual(M, A, slot, A.rhs + (size_t)slot * M.m, x, A.yv + (size_t)slot * M.m, red, mr, mb);
  const double2* xs = A.xsave + (size_t)slot * M.m;
  for (int e = threadIdx.x; e < M.m; e += 256) x[e] = cadd(x[e], xs[e]);
  if (threadIdx.x == 0) {
    A.nrefine[slot] += 1;
    if (mr <= 1e-6 * fmax(A.fnorm[slot], 1e-300)) A.refine[slot] = 0;
    else if (last) { A.refine[slot] = 0; A.singular[slot] = 1; }
  }
}

// V += damp * dx (slack untouched).  grid (ceil(m/8), ceil(nStep/32)), block (32, 8)
__global__ void __launch_bounds__(256) k_update(DevModel M, Slots A, Lists L) {
  const int ns = L.counts[2];
  const int li = blockIdx.y * 32 + threadIdx.x;
  const int node = blockIdx.x * 8 + threadIdx.y;
  if (li >= ns || node >= M.m) return;
  const int slot = L.step[li];
  if (A.singular[slot] || A.refine[slot]) return;    // the reference breaks before applying the step; a step waiting for refinement is withheld
  const int bus = M.bus_of_node[node];
  const double2 dx = A.xv[(size_t)slot * M.m + M.elim_of_node[node]];
  const size_t k = (size_t)bus * M.W + slot;
  double2 v = A.V[k];
  v.x += M.damp * dx.x;
  v.y += M.damp * dx.y;
  A.V[k] = v;
}

__global__ void k_poststep(DevModel M, Slots A, Lists L) {
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= L.counts[2]) return;
  const int slot = L.step[li];
  if (A.singular[slot]) { A.state[slot] = ST_DRAIN; A.numerical[slot] = 0; A.reason[slot] = 3; }
  else if (A.refine[slot]) { A.iter[slot] -= 1; }    // not refined in this tick (the host learns of it one tick late): the iteration is repeated
  else if (A.iter[slot] >= M.max_iter) { A.state[slot] = ST_DRAIN; A.numerical[slot] = 0; A.reason[slot] = 1; }
}

// ------------------------------------------------------------------------------------------
// finalisation of finished scenarios
// ------------------------------------------------------------------------------------------
__global__ void k_fin_init(DevModel M, Slots A, Lists L) {
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= L.counts[3]) return;
  const int slot = L.fin[li];
  A.pvres_bits[slot] = 0ull;
  A.vmin_bits[slot] = 0x7ff0000000000000ull;   // +Inf
  A.vmax_bits[slot] = 0ull;
  A.load_bits[slot] = 0ull;
  A.viol[slot] = 0;
  A.maxsw[slot] = 0;
  A.rated[slot] = 0;
  A.nvv[slot] = 0;
  A.nov[slot] = 0;
}

// grid (ceil(n/8), ceil(nFin/32)), block (32, 8)
__global__ void __launch_bounds__(256) k_fin_bus(DevModel M, Slots A, Lists L, Batch Bt) {
  const int nf = L.counts[3];
  const int li = blockIdx.y * 32 + threadIdx.x;
  const int bus = blockIdx.x * 8 + threadIdx.y;
  if (li >= nf || bus >= M.n) return;
  const int slot = L.fin[li];
  const int scen = A.scen[slot];
  const size_t k = (size_t)bus * M.W + slot;
  const double2 V = A.V[k];
  const int type = A.type[k];
  const bool iso = A.iso[k] != 0;
  SBLX_CHECK(scen >= 0 && scen < Bt.B);
  if (Bt.r_V) Bt.r_V[(size_t)scen * M.n + bus] = V;
  if (Bt.r_types) Bt.r_types[(size_t)scen * M.n + bus] = (unsigned char)(iso ? T_ISO : type);
  if (!A.numerical[slot]) return;
  const double vm = sqrt(V.x * V.x + V.y * V.y);
  if (!iso && isfinite(vm)) {
    atomicMin(&A.vmin_bits[slot], (unsigned long long)__double_as_longlong(vm));
    atomicMax(&A.vmax_bits[slot], (unsigned long long)__double_as_longlong(vm));
    if (vm < M.vm_min || vm > M.vm_max) {            // voltage_violations (contingency.jl:160-162)
      atomicAdd(&A.nvv[slot], 1);
      if (Bt.viol_list) {
        const unsigned long long q = atomicAdd(&Bt.list_counts[0], 1ull);
        if ((long long)q < Bt.viol_cap) Bt.viol_list[q] = make_int2(scen, bus);
      }
    }
  }
  if (bus == M.slack) return;
  if (type == T_PV) {
    if (!A.evact[k]) atomicMax(&A.pvres_bits[slot], abs_bits(vm - M.vset[bus]));   // rectangular_core_equations.jl:150-161
    const double qmin = M.qmin[bus], qmax = M.qmax[bus];
    if (M.qlim_enabled && isfinite(qmin) && isfinite(qmax)) {                       // rectangular_network_solver.jl:64-95
      const double2 Sc = cmul(V, cconj(A.I[k]));
      const double qc = Sc.y + (A.qload_s ? A.qload_s[k] : M.qload[bus]);
      if (qc < qmin - M.tol || qc > qmax + M.tol) atomicAdd(&A.viol[slot], 1);
    }
  }
  if (A.evcnt[k] >= max(M.guard, 1)) atomicOr(&A.maxsw[slot], 1);                  // :1027-1036
}

// branch flows and loading (losses.jl:53-85, contingency.jl:165-176). grid (ceil(nbr/8), ceil(nFin/32))
__global__ void __launch_bounds__(256) k_fin_branch(DevModel M, Slots A, Lists L, Batch Bt) {
  const int nf = L.counts[3];
  const int li = blockIdx.y * 32 + threadIdx.x;
  const int br = blockIdx.x * 8 + threadIdx.y;
  if (li >= nf || br >= M.nbr) return;
  const int slot = L.fin[li];
  const int scen = A.scen[slot];
  double2* fl = Bt.r_flows ? Bt.r_flows + ((size_t)scen * M.nbr + br) * 2 : nullptr;
  if (fl) { fl[0] = make_double2(0.0, 0.0); fl[1] = make_double2(0.0, 0.0); }   // out of service / not solved: zero flow
  if (!A.numerical[slot] || !M.br_status[br]) return;
  const int no = A.out_cnt[slot];
  for (int o = 0; o < no; ++o)
    if (A.out_br[slot * MAXOUT + o] == br) return;
  const int f = M.br_from[br], t = M.br_to[br];
  if (A.iso[(size_t)f * M.W + slot] || A.iso[(size_t)t * M.W + slot]) return;
  const double rt = M.rating[br];
  const bool rated = isfinite(rt) && rt > 0.0;
  if (!rated && !fl) return;
  const double2 Vf = A.V[(size_t)f * M.W + slot], Vt = A.V[(size_t)t * M.W + slot];
  const double2 If = cadd(cmul(M.y11[br], Vf), cmul(M.y12[br], Vt));
  const double2 It = cadd(cmul(M.y22[br], Vt), cmul(M.y21[br], Vf));
  const double2 Sf = cmul(Vf, cconj(If)), St = cmul(Vt, cconj(It));
  if (fl) { fl[0] = Sf; fl[1] = St; }
  if (!rated) return;
  const double sf = hypot(Sf.x, Sf.y), st = hypot(St.x, St.y);
  const double loading = 100.0 * fmax(sf, st) * M.base_mva / rt;
  atomicMax(&A.load_bits[slot], abs_bits(loading));
  A.rated[slot] = 1;
  if (loading > 100.0) {                              // overloaded_branches (contingency.jl:175)
    atomicAdd(&A.nov[slot], 1);
    if (Bt.over_list) {
      const unsigned long long q = atomicAdd(&Bt.list_counts[1], 1ull);
      if ((long long)q < Bt.over_cap) { Bt.over_list[q] = make_int2(scen, br); Bt.over_pct[q] = loading; }
    }
  }
}

__global__ void k_fin_write(DevModel M, Slots A, Lists L, Batch Bt) {
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= L.counts[3]) return;
  const int slot = L.fin[li];
  const int scen = A.scen[slot];
  int conv = A.numerical[slot];
  int reason = A.reason[slot];
  if (conv) {
    const double pvres = bits_double(A.pvres_bits[slot]);
    if (pvres > M.tol) { conv = 0; reason = 5; }
    if (M.qlim_enabled && A.viol[slot] > 0) { conv = 0; reason = 4; }
    if (M.freeze && A.maxsw[slot]) { conv = 0; reason = 6; }
  }
  Bt.r_conv[scen] = conv;
  Bt.r_status[scen] = reason;
  Bt.r_iter[scen] = A.iter[slot];
  Bt.r_nsw[scen] = A.nev[slot];
  Bt.r_fm[scen] = A.fm[slot];
  Bt.r_nvv[scen] = conv ? A.nvv[slot] : 0;
  Bt.r_nov[scen] = conv ? A.nov[slot] : 0;
  Bt.r_tiny[scen] = A.tiny[slot];
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  if (conv) {
    const double vmin = bits_double(A.vmin_bits[slot]), vmax = bits_double(A.vmax_bits[slot]);
    Bt.r_vmin[scen] = isfinite(vmin) ? vmin : nan;
    Bt.r_vmax[scen] = (vmax > 0.0) ? vmax : nan;
    Bt.r_load[scen] = A.rated[slot] ? bits_double(A.load_bits[slot]) : nan;
  } else {
    Bt.r_vmin[scen] = nan; Bt.r_vmax[scen] = nan; Bt.r_load[scen] = nan;
  }
  A.state[slot] = ST_IDLE;
  A.scen[slot] = -1;
}

// Per-scenario summary records of the B CALLER scenarios (what the NCCL all-gather moves and what fetch reads back):
// {i32 converged, status, iterations, n_switch; f64 final_mismatch, vmin, vmax, max_loading; i32 n_voltage_violations,
// n_overloaded, island_count, n_tiny_pivots} = 64 bytes.  Scenarios decided on the host (bad branch, islanded without
// reference) get their host verdict; island-wise composites are reduced from their sub-scenarios exactly like runpf!
// recombines the islands (iterations summed, the first failing island fails the case with 0 iterations).
struct SummaryRec {
  int conv, status, iter, nsw;
  double fm, vmin, vmax, load;
  int nvv, nov, islands, tiny;
};
__global__ void k_pack_summary(Batch Bt, const int* __restrict__ host_status, const int* __restrict__ sub_first,
                               const int* __restrict__ sub_count, const int* __restrict__ islands, SummaryRec* out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= Bt.Btop) return;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  SummaryRec r;
  r.islands = islands[s];
  const int hs = host_status[s];
  if (hs == -1) {
    r.conv = Bt.r_conv[s]; r.status = Bt.r_status[s]; r.iter = Bt.r_iter[s]; r.nsw = Bt.r_nsw[s];
    r.fm = Bt.r_fm[s]; r.vmin = Bt.r_vmin[s]; r.vmax = Bt.r_vmax[s]; r.load = Bt.r_load[s];
    r.nvv = Bt.r_nvv[s]; r.nov = Bt.r_nov[s]; r.tiny = Bt.r_tiny[s];
  } else if (hs == -2) {
    r.conv = 1; r.status = 0; r.iter = 0; r.nsw = 0; r.fm = 0.0; r.vmin = nan; r.vmax = nan; r.load = nan;
    r.nvv = 0; r.nov = 0; r.tiny = 0;
    bool any_ld = false;
    for (int q = 0; q < sub_count[s]; ++q) {
      const int u = sub_first[s] + q;
      r.tiny += Bt.r_tiny[u];
      if (!Bt.r_conv[u]) { r.conv = 0; r.status = Bt.r_status[u]; r.iter = 0; break; }
      r.iter += Bt.r_iter[u]; r.nsw += Bt.r_nsw[u];
      r.fm = fmax(r.fm, Bt.r_fm[u]);
      r.vmin = isnan(r.vmin) ? Bt.r_vmin[u] : fmin(r.vmin, Bt.r_vmin[u]);
      r.vmax = isnan(r.vmax) ? Bt.r_vmax[u] : fmax(r.vmax, Bt.r_vmax[u]);
      const double ld = Bt.r_load[u];
      if (!isnan(ld)) { r.load = any_ld ? fmax(r.load, ld) : ld; any_ld = true; }
      r.nvv += Bt.r_nvv[u]; r.nov += Bt.r_nov[u];
    }
    if (!r.conv) { r.fm = nan; r.vmin = nan; r.vmax = nan; r.load = nan; r.nsw = 0; r.nvv = 0; r.nov = 0; }
  } else {
    r.conv = 0; r.status = hs; r.iter = 0; r.nsw = 0; r.fm = nan; r.vmin = nan; r.vmax = nan; r.load = nan;
    r.nvv = 0; r.nov = 0; r.tiny = 0;
  }
  out[s] = r;
}

}  // namespace sblx
