// cpu_ref.cpp - reference-shaped multithreaded CPU baseline of the contingency hot path.
//
// TEST / MEASUREMENT INFRASTRUCTURE ONLY (like everything under oracle/): it is built by oracle/build_cpu_ref.sh,
// checked against the Python oracle in tests/test_cpu_ref.py, and timed by bench.py's `cpu_baseline` leg and
// `--impl reference`.  The product (sparlectra.jl_b200/) never links, loads or calls it.
//
// What it restates, per scenario and in the reference's order (paths relative to Sparlectra.jl/src):
//   * contingency fan-out in contiguous chunks over worker threads        contingency.jl:299-318
//   * removeBranch! + markIsolatedBuses! + fresh Ybus assembly            contingency.jl:183-190, network.jl:1881-1930,
//                                                                         equicircuit.jl:583-645, branch.jl:297-310
//   * rectangular NR loop with the Q-limit active set (defaults)          powerflow_rectangular/rectangular_network_solver.jl:746-922,
//                                                                         rectangular_qlimit_iteration.jl:25-210, limits.jl:398-568
//   * mismatch / Jacobian                                                 rectangular_core_equations.jl:79-148, rectangular_jacobian_builders.jl:212-429
//   * sparse direct solve, re-analysed EVERY iteration (variant A, the reference default `A \ b`, solver_core.jl:60-65)
//     or with the symbolic analysis reused (variant B, `umfpack_reuse`, rectangular_factorized_solver.jl:101-132)
//   * final acceptance + contingency metrics                              rectangular_network_solver.jl:939-1094, contingency.jl:149-178
// The linear solver is this file's own multifrontal LU (minimum-degree order on A+A^T of the bus graph, supernodes,
// dense fronts with partial pivoting inside the fully-summed block) - the role UMFPACK plays in the reference; Julia /
// SuiteSparse are not available in this image.  It omits deepcopy(Net), status logging and diagnostics, so it is
// FASTER than the real Julia path: a conservative denominator.  Not supported (reported as status 9, the tests skip
// them): outages that split the grid into several islands with branches; hysteresis / cooldown / lock lists.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <functional>
#include <queue>
#include <thread>
#include <vector>

namespace {

using cd = std::complex<double>;
constexpr int T_ISO = 0, T_PQ = 1, T_PV = 2, T_SLACK = 3;

struct Grid {
  int n = 0, slack = 0, nbr = 0;
  std::vector<int> rowptr, col;          // Ybus pattern, CSR (structurally symmetric superset)
  std::vector<cd> ysh;                   // bus shunts on the diagonal (Ybus_ii minus the branch stamps)
  std::vector<int> type0;
  std::vector<double> vset, qmin, qmax, qload, rating;
  std::vector<cd> S0, V0;
  std::vector<int> bf, bt, status;
  std::vector<cd> y11, y12, y21, y22;
  std::vector<int> p11, p12, p21, p22;   // CSR positions of every branch's four stamps
  double base_mva = 100.0;
  // non-slack numbering and the block pattern of the Jacobian
  std::vector<int> node_of_bus, bus_of_node;
  std::vector<int> jrow_ptr, jcol, jypos;  // block CSR over nodes, jypos = position in the Ybus CSR
  int m = 0;
};

// ------------------------------------------------------------------------------------------
// symbolic analysis: minimum degree (explicit elimination graph), postordered elimination tree, supernodes
// ------------------------------------------------------------------------------------------
struct Supernode {
  int first = 0, w = 0;
  std::vector<int> st;      // structure, elimination indices ascending
  int parent = -1;
  std::vector<int> asm_slot, asm_r, asm_c;   // Jacobian block slot -> local block row / column
};
struct Symbolic {
  std::vector<int> perm, iperm;   // elimination index -> node, node -> elimination index
  std::vector<Supernode> sn;
  std::vector<std::vector<int>> children;
  long long nnz_lu = 0;
};

// Approximate minimum degree on the quotient graph (Amestoy, Davis, Duff 1996, without supervariables): a pivot turns
// into an element whose variable list replaces the clique it would create; adjacent elements are absorbed; the degree
// of a variable i is bounded by |A_i| + |L_p \ i| + sum over its other elements e of |L_e \ L_p|, the last term from
// one counting pass (w[e]).  This is the job COLAMD/AMD does inside UMFPACK's symbolic phase.
void min_degree(const std::vector<std::vector<int>>& adj_in, std::vector<int>& order) {
  const int n = (int)adj_in.size();
  std::vector<std::vector<int>> A = adj_in, E(n), Le(n);
  std::vector<char> alive(n, 1), el_alive(n, 0);
  std::vector<int> deg(n), stamp(n, -1), wstamp(n, -1), w(n, 0);
  using QE = std::pair<int, int>;
  std::priority_queue<QE, std::vector<QE>, std::greater<QE>> heap;
  for (int i = 0; i < n; ++i) { deg[i] = (int)A[i].size(); heap.push({deg[i], i}); }
  order.clear();
  order.reserve(n);
  std::vector<int> Lp;
  for (int k = 0; k < n; ++k) {
    int p = -1;
    while (!heap.empty()) {
      const QE top = heap.top();
      heap.pop();
      if (alive[top.second] && top.first == deg[top.second]) { p = top.second; break; }
    }
    if (p < 0) break;
    alive[p] = 0;
    order.push_back(p);
    Lp.clear();
    stamp[p] = k;
    for (int v : A[p])
      if (alive[v] && stamp[v] != k) { stamp[v] = k; Lp.push_back(v); }
    for (int e : E[p])
      if (el_alive[e]) {
        for (int v : Le[e])
          if (alive[v] && stamp[v] != k) { stamp[v] = k; Lp.push_back(v); }
        el_alive[e] = 0;
        std::vector<int>().swap(Le[e]);
      }
    std::vector<int>().swap(A[p]);
    std::vector<int>().swap(E[p]);
    for (int i : Lp)
      for (int e : E[i])
        if (el_alive[e]) {
          if (wstamp[e] != k) { wstamp[e] = k; w[e] = (int)Le[e].size(); }
          --w[e];
        }
    const int lp = (int)Lp.size();
    for (int i : Lp) {
      auto& Ai = A[i];
      size_t o = 0;
      for (int v : Ai)
        if (alive[v] && stamp[v] != k) Ai[o++] = v;        // variables inside L_p are now reached through the element
      Ai.resize(o);
      auto& Ei = E[i];
      o = 0;
      long long d = (long long)Ai.size() + (lp - 1);
      for (int e : Ei) {
        if (!el_alive[e]) continue;
        if (w[e] == 0) continue;                            // L_e inside L_p: every member drops it here = absorbed
        Ei[o++] = e;
        d += w[e];
      }
      Ei.resize(o);
      Ei.push_back(p);
      deg[i] = (int)std::min<long long>(d, n - k - 1);
      heap.push({deg[i], i});
    }
    Le[p] = Lp;
    el_alive[p] = 1;
  }
}

void col_structures(const std::vector<std::vector<int>>& adj, const std::vector<int>& order, std::vector<std::vector<int>>& st,
                    std::vector<int>& parent) {
  const int n = (int)adj.size();
  std::vector<int> pos(n);
  for (int e = 0; e < n; ++e) pos[order[e]] = e;
  st.assign(n, {});
  parent.assign(n, -1);
  std::vector<std::vector<int>> kids(n);
  std::vector<int> tmp;
  for (int e = 0; e < n; ++e) {
    auto& s = st[e];
    for (int u : adj[order[e]])
      if (pos[u] > e) s.push_back(pos[u]);
    std::sort(s.begin(), s.end());
    for (int c : kids[e]) {
      tmp.clear();
      std::set_union(s.begin(), s.end(), st[c].begin(), st[c].end(), std::back_inserter(tmp));
      tmp.erase(std::remove(tmp.begin(), tmp.end(), e), tmp.end());
      s.swap(tmp);
    }
    if (!s.empty()) {
      parent[e] = s.front();
      kids[s.front()].push_back(e);
    }
  }
}

void analyze(const Grid& G, Symbolic& S) {
  const int m = G.m;
  std::vector<std::vector<int>> adj(m);
  for (int a = 0; a < m; ++a)
    for (int k = G.jrow_ptr[a]; k < G.jrow_ptr[a + 1]; ++k)
      if (G.jcol[k] != a) adj[a].push_back(G.jcol[k]);
  std::vector<int> order;
  min_degree(adj, order);
  std::vector<std::vector<int>> st;
  std::vector<int> parent;
  col_structures(adj, order, st, parent);
  // postorder of the elimination tree (children before parents, subtrees contiguous), then the structures again
  {
    std::vector<std::vector<int>> kids(m);
    std::vector<int> roots;
    for (int e = 0; e < m; ++e) (parent[e] < 0 ? roots : kids[parent[e]]).push_back(e);
    std::vector<int> post;
    post.reserve(m);
    std::vector<std::pair<int, size_t>> stack;
    for (int r : roots) {
      stack.push_back({r, 0});
      while (!stack.empty()) {
        auto& top = stack.back();
        if (top.second < kids[top.first].size()) {
          const int c = kids[top.first][top.second++];
          stack.push_back({c, 0});
        } else {
          post.push_back(top.first);
          stack.pop_back();
        }
      }
    }
    std::vector<int> order2(m);
    for (int e = 0; e < m; ++e) order2[e] = order[post[e]];
    order.swap(order2);
    col_structures(adj, order, st, parent);
  }
  S.perm = order;
  S.iperm.assign(m, 0);
  for (int e = 0; e < m; ++e) S.iperm[order[e]] = e;
  // fundamental supernodes
  S.sn.clear();
  std::vector<int> sn_of(m, -1);
  for (int e = 0; e < m; ++e) {
    const bool join = e > 0 && parent[e - 1] == e && st[e - 1].size() == st[e].size() + 1 && S.sn.back().w < 48;
    if (!join) {
      S.sn.emplace_back();
      S.sn.back().first = e;
    }
    S.sn.back().w++;
    sn_of[e] = (int)S.sn.size() - 1;
  }
  S.nnz_lu = 0;
  for (auto& s : S.sn) {
    s.st = st[s.first + s.w - 1];
    const long long w = s.w, h = (long long)s.st.size();
    S.nnz_lu += 4 * (w * w + 2 * w * h);
  }
  S.children.assign(S.sn.size(), {});
  for (size_t k = 0; k < S.sn.size(); ++k)
    if (!S.sn[k].st.empty()) {
      S.sn[k].parent = sn_of[S.sn[k].st.front()];
      S.children[S.sn[k].parent].push_back((int)k);
    }
  // assembly lists of the original Jacobian blocks
  std::vector<int> loc(m, -1);
  for (size_t k = 0; k < S.sn.size(); ++k) {
    Supernode& s = S.sn[k];
    for (int p = 0; p < s.w; ++p) loc[s.first + p] = p;
    for (size_t q = 0; q < s.st.size(); ++q) loc[s.st[q]] = s.w + (int)q;
    for (int p = 0; p < s.w; ++p) {
      const int node = S.perm[s.first + p];
      for (int z = G.jrow_ptr[node]; z < G.jrow_ptr[node + 1]; ++z) {
        const int ej = S.iperm[G.jcol[z]];
        if (ej < s.first) continue;                       // belongs to an earlier front
        // row entry (node, j): all columns >= first; the transposed position (j, node) is handled when j's row is
        // walked if j is a pivot here, else now
        s.asm_slot.push_back(z); s.asm_r.push_back(p); s.asm_c.push_back(loc[ej]);
      }
    }
    // column entries: rows outside the pivot set (structure rows) x pivot columns
    for (size_t q = 0; q < s.st.size(); ++q) {
      const int node = S.perm[s.st[q]];
      for (int z = G.jrow_ptr[node]; z < G.jrow_ptr[node + 1]; ++z) {
        const int ej = S.iperm[G.jcol[z]];
        if (ej >= s.first && ej < s.first + s.w) { s.asm_slot.push_back(z); s.asm_r.push_back(s.w + (int)q); s.asm_c.push_back(ej - s.first); }
      }
    }
    for (int p = 0; p < s.w; ++p) loc[s.first + p] = -1;
    for (size_t q = 0; q < s.st.size(); ++q) loc[s.st[q]] = -1;
  }
}

// ------------------------------------------------------------------------------------------
// numeric multifrontal LU + solve.  Fronts are dense scalar matrices of order 2 (w + h), row-major.
// ------------------------------------------------------------------------------------------
struct Factors {
  std::vector<std::vector<double>> LU;   // per supernode: rows 0..2w (U11 | U12) and rows 2w.. (L21), with L11 unit-lower in place
  std::vector<std::vector<int>> piv;     // row swaps inside the pivot block
  std::vector<std::vector<double>> upd;  // update matrices waiting for their parents
};

bool factor(const Symbolic& S, const std::vector<double>& J /* 4 per block slot */, Factors& F) {
  const size_t ns = S.sn.size();
  F.LU.assign(ns, {});
  F.piv.assign(ns, {});
  F.upd.assign(ns, {});
  std::vector<int> loc;
  int maxe = 0;
  for (auto& s : S.sn) maxe = std::max(maxe, s.st.empty() ? s.first + s.w : s.st.back() + 1);
  loc.assign(maxe + 1, -1);
  for (size_t k = 0; k < ns; ++k) {
    const Supernode& s = S.sn[k];
    const int w = s.w, h = (int)s.st.size(), nf = w + h, N = 2 * nf, P = 2 * w;
    std::vector<double>& A = F.LU[k];
    A.assign((size_t)N * N, 0.0);
    for (size_t q = 0; q < s.asm_slot.size(); ++q) {
      const double* b = &J[4 * (size_t)s.asm_slot[q]];
      const int r = 2 * s.asm_r[q], c = 2 * s.asm_c[q];
      A[(size_t)r * N + c] += b[0]; A[(size_t)r * N + c + 1] += b[1];
      A[(size_t)(r + 1) * N + c] += b[2]; A[(size_t)(r + 1) * N + c + 1] += b[3];
    }
    if (!S.children[k].empty()) {
      for (int p = 0; p < w; ++p) loc[s.first + p] = p;
      for (int q = 0; q < h; ++q) loc[s.st[q]] = w + q;
      for (int c : S.children[k]) {
        const Supernode& ch = S.sn[c];
        const int hc = (int)ch.st.size(), Nc = 2 * hc;
        const std::vector<double>& U = F.upd[c];
        for (int a = 0; a < hc; ++a) {
          const int la = 2 * loc[ch.st[a]];
          for (int b = 0; b < hc; ++b) {
            const int lb = 2 * loc[ch.st[b]];
            const double* u = &U[(size_t)(2 * a) * Nc + 2 * b];
            A[(size_t)la * N + lb] += u[0]; A[(size_t)la * N + lb + 1] += u[1];
            A[(size_t)(la + 1) * N + lb] += u[Nc]; A[(size_t)(la + 1) * N + lb + 1] += u[Nc + 1];
          }
        }
        std::vector<double>().swap(F.upd[c]);
      }
    }
    // panel factorisation with partial pivoting among the fully-summed rows
    std::vector<int>& pv = F.piv[k];
    pv.resize(P);
    for (int j = 0; j < P; ++j) {
      int best = j;
      double bv = std::fabs(A[(size_t)j * N + j]);
      for (int r = j + 1; r < P; ++r) {
        const double v = std::fabs(A[(size_t)r * N + j]);
        if (v > bv) { bv = v; best = r; }
      }
      pv[j] = best;
      if (!(bv > 0.0) || !std::isfinite(bv)) return false;
      if (best != j)
        for (int c = 0; c < N; ++c) std::swap(A[(size_t)j * N + c], A[(size_t)best * N + c]);
      const double inv = 1.0 / A[(size_t)j * N + j];
      const double* uj = &A[(size_t)j * N];
      for (int r = j + 1; r < N; ++r) {           // multipliers for ALL rows below, trailing update of the panel columns only
        double* ar = &A[(size_t)r * N];
        const double l = ar[j] * inv;
        ar[j] = l;
        if (l != 0.0) {
          const int cend = r < P ? N : P;            // pivot rows carry their whole row (U12 included); structure rows only the panel part
          for (int c = j + 1; c < cend; ++c) ar[c] -= l * uj[c];
        }
      }
    }
    // Schur complement of the structure rows: A22 -= L21 * U12  (rank-P update, inner loop contiguous)
    if (h > 0) {
      const int H = N - P;
      std::vector<double>& C = F.upd[k];
      C.resize((size_t)H * H);
      for (int r = 0; r < H; ++r) {
        double* cr = &C[(size_t)r * H];
        const double* ar = &A[(size_t)(P + r) * N];
        std::memcpy(cr, ar + P, sizeof(double) * H);
        for (int j = 0; j < P; ++j) {
          const double l = ar[j];
          if (l == 0.0) continue;
          const double* uj = &A[(size_t)j * N + P];
          for (int c = 0; c < H; ++c) cr[c] -= l * uj[c];
        }
      }
      for (int p = 0; p < w; ++p) loc[s.first + p] = -1;
      for (int q = 0; q < h; ++q) loc[s.st[q]] = -1;
    }
  }
  return true;
}

void solve(const Symbolic& S, const Factors& F, std::vector<double>& x /* 2 per elimination index, in: rhs, out: solution */) {
  const size_t ns = S.sn.size();
  for (size_t k = 0; k < ns; ++k) {
    const Supernode& s = S.sn[k];
    const int w = s.w, h = (int)s.st.size(), N = 2 * (w + h), P = 2 * w;
    const std::vector<double>& A = F.LU[k];
    double* xp = &x[2 * (size_t)s.first];
    for (int j = 0; j < P; ++j)                       // whole rows were swapped (multipliers included): all swaps first
      if (F.piv[k][j] != j) std::swap(xp[j], xp[F.piv[k][j]]);
    for (int j = 0; j < P; ++j) {
      const double v = xp[j];
      for (int r = j + 1; r < P; ++r) xp[r] -= A[(size_t)r * N + j] * v;
    }
    for (int q = 0; q < h; ++q) {
      double* xs = &x[2 * (size_t)s.st[q]];
      for (int z = 0; z < 2; ++z) {
        const double* ar = &A[(size_t)(P + 2 * q + z) * N];
        double acc = 0.0;
        for (int j = 0; j < P; ++j) acc += ar[j] * xp[j];
        xs[z] -= acc;
      }
    }
  }
  for (size_t kk = ns; kk-- > 0;) {
    const Supernode& s = S.sn[kk];
    const int w = s.w, h = (int)s.st.size(), N = 2 * (w + h), P = 2 * w;
    const std::vector<double>& A = F.LU[kk];
    double* xp = &x[2 * (size_t)s.first];
    for (int j = P - 1; j >= 0; --j) {
      const double* aj = &A[(size_t)j * N];
      double acc = xp[j];
      for (int c = j + 1; c < P; ++c) acc -= aj[c] * xp[c];
      for (int q = 0; q < h; ++q) {
        const double* xs = &x[2 * (size_t)s.st[q]];
        acc -= aj[P + 2 * q] * xs[0] + aj[P + 2 * q + 1] * xs[1];
      }
      xp[j] = acc / aj[j];
    }
  }
}

// ------------------------------------------------------------------------------------------
// one scenario
// ------------------------------------------------------------------------------------------
struct Opts {
  int variant = 0, max_iter = 30, qlim = 1, qlim_start = 2, guard = 10;
  double tol = 1e-8, vm_min = 0.9, vm_max = 1.1;
};
struct Out {
  int conv = 0, iters = 0, status = 1, nsw = 0;
  double vmin = NAN, vmax = NAN, load = NAN;
};

struct Worker {
  const Grid& G;
  const Opts& O;
  Symbolic sym;
  bool have_sym = false;
  Factors fac;
  std::vector<cd> Yv, V, I, S;
  std::vector<int> type, evcnt, iso;
  std::vector<char> evact;
  std::vector<double> J, rhs, x;
  explicit Worker(const Grid& g, const Opts& o) : G(g), O(o) {}

  void injections() {
    for (int i = 0; i < G.n; ++i) {
      cd acc = 0.0;
      for (int p = G.rowptr[i]; p < G.rowptr[i + 1]; ++p) acc += Yv[p] * V[G.col[p]];
      I[i] = acc;
    }
  }
  double mismatch() {   // rhs = -F in node order, returns max |F| (NaN-propagating)
    double mx = 0.0;
    bool bad = false;
    for (int a = 0; a < G.m; ++a) {
      const int i = G.bus_of_node[a];
      const cd sc = V[i] * std::conj(I[i]);
      const double f0 = sc.real() - S[i].real();
      const double f1 = type[i] == T_PV ? std::abs(V[i]) - G.vset[i] : sc.imag() - S[i].imag();
      rhs[2 * a] = -f0; rhs[2 * a + 1] = -f1;
      if (std::isnan(f0) || std::isnan(f1)) bad = true;
      mx = std::max(mx, std::max(std::fabs(f0), std::fabs(f1)));
    }
    return bad ? NAN : mx;
  }
  void jacobian() {
    for (int a = 0; a < G.m; ++a) {
      const int i = G.bus_of_node[a];
      for (int z = G.jrow_ptr[a]; z < G.jrow_ptr[a + 1]; ++z) {
        const int j = G.bus_of_node[G.jcol[z]];
        double* b = &J[4 * (size_t)z];
        if (iso[i] || iso[j]) { b[0] = b[3] = (i == j) ? 1.0 : 0.0; b[1] = b[2] = 0.0; continue; }
        const cd al = V[i] * std::conj(Yv[G.jypos[z]]);
        double dvr_x = al.real(), dvr_y = al.imag(), dvi_x = al.imag(), dvi_y = -al.real();
        if (i == j) { dvr_x += I[i].real(); dvr_y -= I[i].imag(); dvi_x += I[i].imag(); dvi_y += I[i].real(); }
        b[0] = dvr_x; b[1] = dvi_x;
        if (type[i] == T_PV) {
          if (i == j) { const double vm = std::abs(V[i]), d = vm > 0.0 ? vm : 1e-9; b[2] = V[i].real() / d; b[3] = V[i].imag() / d; }
          else { b[2] = 0.0; b[3] = 0.0; }
        } else { b[2] = dvr_y; b[3] = dvi_y; }
      }
    }
  }

  void run(const int* out, int nout, Out& R, cd* Vout, uint8_t* Tout) {
    const int n = G.n;
    // removeBranch! / markIsolatedBuses! / fresh Ybus (diagonal shunts + every in-service stamp)
    Yv.assign(G.col.size(), cd(0.0));
    std::vector<int> deg(n, 0);
    auto removed = [&](int k) { for (int q = 0; q < nout; ++q) if (out[q] == k) return true; return false; };
    for (int i = 0; i < n; ++i) Yv[G.rowptr[i] + (int)(std::lower_bound(G.col.begin() + G.rowptr[i], G.col.begin() + G.rowptr[i + 1], i) - (G.col.begin() + G.rowptr[i]))] = G.ysh[i];
    for (int k = 0; k < G.nbr; ++k) {
      if (!G.status[k] || removed(k)) continue;
      Yv[G.p11[k]] += G.y11[k]; Yv[G.p12[k]] += G.y12[k]; Yv[G.p21[k]] += G.y21[k]; Yv[G.p22[k]] += G.y22[k];
      deg[G.bf[k]]++; deg[G.bt[k]]++;
    }
    iso.assign(n, 0);
    for (int i = 0; i < n; ++i) iso[i] = deg[i] == 0;
    // islands: one BFS from the slack over in-service branches; anything else with branches = unsupported here
    {
      std::vector<std::vector<int>> adj(n);
      for (int k = 0; k < G.nbr; ++k)
        if (G.status[k] && !removed(k)) { adj[G.bf[k]].push_back(G.bt[k]); adj[G.bt[k]].push_back(G.bf[k]); }
      std::vector<char> seen(n, 0);
      std::vector<int> q;
      if (!iso[G.slack]) { q.push_back(G.slack); seen[G.slack] = 1; }
      for (size_t h = 0; h < q.size(); ++h)
        for (int u : adj[q[h]])
          if (!seen[u]) { seen[u] = 1; q.push_back(u); }
      bool split = false;
      for (int i = 0; i < n; ++i)
        if (!iso[i] && !seen[i]) split = true;
      if (iso[G.slack] || split) {
        bool anyref = true;   // coarse: report "islanded" - the oracle decides between island-wise solve and no reference
        (void)anyref;
        R.conv = 0; R.iters = 0; R.status = iso[G.slack] ? 7 : 9;
        return;
      }
    }
    type.assign(G.type0.begin(), G.type0.end());
    S = G.S0;
    V = G.V0;
    for (int i = 0; i < n; ++i)
      if (iso[i]) { type[i] = T_PQ; S[i] = 0.0; V[i] = cd(1.0, 0.0); }
    I.assign(n, cd(0.0));
    evcnt.assign(n, 0);
    evact.assign(n, 0);
    J.resize(4 * G.jcol.size());
    rhs.resize(2 * (size_t)G.m);
    x.resize(2 * (size_t)G.m);
    bool converged = false;
    int reason = 1, iters = 0, nsw = 0;
    for (int it = 1; it <= O.max_iter; ++it) {
      iters = it;
      injections();
      double mis = mismatch();
      if (!std::isfinite(mis)) { reason = 2; break; }
      bool changed = false;
      const bool conv_now = mis <= O.tol;
      if (O.qlim && (it >= O.qlim_start || conv_now)) {
        for (int i = 0; i < n; ++i) {
          if (type[i] != T_PV || iso[i]) continue;
          const bool hi = std::isfinite(G.qmax[i]), lo = std::isfinite(G.qmin[i]);
          if (!(hi || lo)) continue;
          const double qreq = (V[i] * std::conj(I[i])).imag() + G.qload[i];
          double clamp;
          if (hi && qreq > G.qmax[i]) clamp = G.qmax[i];
          else if (lo && qreq < G.qmin[i]) clamp = G.qmin[i];
          else continue;
          if (evcnt[i] >= O.guard) continue;
          type[i] = T_PQ;
          S[i] = cd(S[i].real(), clamp - G.qload[i]);
          evcnt[i]++; evact[i] = 1; nsw++;
          changed = true;
        }
        if (changed) mis = mismatch();
      }
      if (conv_now && !changed) { converged = true; reason = 0; break; }
      jacobian();
      if (O.variant == 0 || !have_sym) { analyze(G, sym); have_sym = true; }   // variant A: re-analyse every iteration
      // rhs into elimination order
      for (int a = 0; a < G.m; ++a) { x[2 * (size_t)sym.iperm[a]] = rhs[2 * a]; x[2 * (size_t)sym.iperm[a] + 1] = rhs[2 * a + 1]; }
      if (!factor(sym, J, fac)) { reason = 3; break; }
      solve(sym, fac, x);
      for (int a = 0; a < G.m; ++a) {
        const int i = G.bus_of_node[a];
        V[i] += cd(x[2 * (size_t)sym.iperm[a]], x[2 * (size_t)sym.iperm[a] + 1]);
      }
    }
    // final acceptance
    const bool numerical = converged;
    if (numerical) {
      double pvres = 0.0;
      for (int i = 0; i < n; ++i)
        if (type[i] == T_PV && !evact[i] && i != G.slack) pvres = std::max(pvres, std::fabs(std::abs(V[i]) - G.vset[i]));
      if (pvres > O.tol) { converged = false; reason = 5; }
      if (O.qlim) {
        injections();
        int viol = 0;
        for (int i = 0; i < n; ++i) {
          if (type[i] != T_PV || !(std::isfinite(G.qmin[i]) && std::isfinite(G.qmax[i]))) continue;
          const double qc = (V[i] * std::conj(I[i])).imag() + G.qload[i];
          if (qc < G.qmin[i] - O.tol || qc > G.qmax[i] + O.tol) viol++;
        }
        if (viol > 0) { converged = false; reason = 4; }
      }
      for (int i = 0; i < n; ++i)
        if (evcnt[i] >= std::max(O.guard, 1)) { converged = false; reason = 6; }
    }
    R.conv = converged; R.iters = iters; R.status = converged ? 0 : reason; R.nsw = nsw;
    if (converged) {
      double lo = INFINITY, hi = -INFINITY, ld = NAN;
      for (int i = 0; i < n; ++i) {
        if (iso[i]) continue;
        const double vm = std::abs(V[i]);
        if (!std::isfinite(vm)) continue;
        lo = std::min(lo, vm); hi = std::max(hi, vm);
      }
      for (int k = 0; k < G.nbr; ++k) {
        if (!G.status[k] || removed(k)) continue;
        const double rt = G.rating[k];
        if (!(std::isfinite(rt) && rt > 0.0)) continue;
        const cd vf = V[G.bf[k]], vt = V[G.bt[k]];
        const cd sf = vf * std::conj(G.y11[k] * vf + G.y12[k] * vt), st = vt * std::conj(G.y22[k] * vt + G.y21[k] * vf);
        const double l = 100.0 * std::max(std::abs(sf), std::abs(st)) * G.base_mva / rt;
        if (std::isnan(ld) || l > ld) ld = l;
      }
      R.vmin = std::isfinite(lo) ? lo : NAN; R.vmax = std::isfinite(hi) ? hi : NAN; R.load = ld;
    }
    if (Vout) for (int i = 0; i < n; ++i) Vout[i] = V[i];
    if (Tout) for (int i = 0; i < n; ++i) Tout[i] = (uint8_t)(iso[i] ? T_ISO : type[i]);
  }
};

}  // namespace

extern "C" {

// All indices 0-based.  Ybus: n x n CSC.  Returns 0 on success.  seconds_out[0] = wall time of the scenario loop,
// seconds_out[1] = nnz(L+U) of this solver's factor (scalars).  variant: 0 = symbolic analysis every iteration
// (reference default), 1 = symbolic once per worker thread (umfpack_reuse).
int cpuref_run(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval, int64_t slack, const uint8_t* bus_type,
               const double* vset, const double* s_spec, const double* v0, const double* qmin, const double* qmax, const double* qload,
               int64_t nbr, const int32_t* br_from, const int32_t* br_to, const double* y11, const double* y12, const double* y21,
               const double* y22, const double* rating, const uint8_t* br_status, double base_mva, int64_t B, const int32_t* optr,
               const int32_t* obr, int32_t variant, int32_t nthreads, int32_t max_iter, double tol, int32_t qlimits_enabled,
               int32_t qlimit_start_iter, uint8_t* converged, int32_t* iterations, int32_t* status, int32_t* n_switch, double* V,
               uint8_t* types, double* vmin, double* vmax, double* max_loading, double* seconds_out) {
  if (n < 2 || !colptr || !rowval || !nzval || B < 0) return -1;
  Grid G;
  G.n = (int)n; G.slack = (int)slack; G.nbr = (int)nbr; G.base_mva = base_mva;
  std::vector<std::vector<int>> rows(n);
  for (int64_t j = 0; j < n; ++j)
    for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) { rows[rowval[p]].push_back((int)j); rows[j].push_back((int)rowval[p]); }
  for (int64_t i = 0; i < n; ++i) rows[i].push_back((int)i);
  for (int64_t k = 0; k < nbr; ++k) { rows[br_from[k]].push_back(br_to[k]); rows[br_to[k]].push_back(br_from[k]); }
  G.rowptr.assign(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) {
    auto& r = rows[i];
    std::sort(r.begin(), r.end());
    r.erase(std::unique(r.begin(), r.end()), r.end());
    for (int c : r) G.col.push_back(c);
    G.rowptr[i + 1] = (int)G.col.size();
  }
  auto pos = [&](int i, int j) { return (int)(std::lower_bound(G.col.begin() + G.rowptr[i], G.col.begin() + G.rowptr[i + 1], j) - G.col.begin()); };
  G.ysh.assign(n, cd(0.0));
  for (int64_t j = 0; j < n; ++j)
    for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p)
      if (rowval[p] == j) G.ysh[j] += cd(nzval[2 * p], nzval[2 * p + 1]);
  G.bf.assign(br_from, br_from + nbr); G.bt.assign(br_to, br_to + nbr);
  G.status.resize(nbr); G.rating.resize(nbr);
  G.y11.resize(nbr); G.y12.resize(nbr); G.y21.resize(nbr); G.y22.resize(nbr);
  G.p11.resize(nbr); G.p12.resize(nbr); G.p21.resize(nbr); G.p22.resize(nbr);
  for (int64_t k = 0; k < nbr; ++k) {
    G.status[k] = br_status ? br_status[k] != 0 : 1;
    G.rating[k] = rating ? rating[k] : NAN;
    G.y11[k] = cd(y11[2 * k], y11[2 * k + 1]); G.y12[k] = cd(y12[2 * k], y12[2 * k + 1]);
    G.y21[k] = cd(y21[2 * k], y21[2 * k + 1]); G.y22[k] = cd(y22[2 * k], y22[2 * k + 1]);
    const int f = G.bf[k], t = G.bt[k];
    G.p11[k] = pos(f, f); G.p12[k] = pos(f, t); G.p21[k] = pos(t, f); G.p22[k] = pos(t, t);
    if (G.status[k]) { G.ysh[f] -= G.y11[k]; G.ysh[t] -= G.y22[k]; }   // what remains on the diagonal are the bus shunts
  }
  G.type0.resize(n); G.vset.assign(vset, vset + n); G.S0.resize(n); G.V0.resize(n);
  G.qmin.resize(n); G.qmax.resize(n); G.qload.resize(n);
  for (int64_t i = 0; i < n; ++i) {
    G.type0[i] = bus_type[i] == T_ISO ? T_PQ : bus_type[i];
    G.S0[i] = cd(s_spec[2 * i], s_spec[2 * i + 1]); G.V0[i] = cd(v0[2 * i], v0[2 * i + 1]);
    double lo = qmin ? qmin[i] : -INFINITY, hi = qmax ? qmax[i] : INFINITY;
    if (std::isnan(lo)) lo = -INFINITY;
    if (std::isnan(hi)) hi = INFINITY;
    G.qmin[i] = lo; G.qmax[i] = hi; G.qload[i] = qload ? qload[i] : 0.0;
  }
  G.node_of_bus.assign(n, -1);
  for (int i = 0; i < n; ++i)
    if (i != G.slack) { G.node_of_bus[i] = (int)G.bus_of_node.size(); G.bus_of_node.push_back(i); }
  G.m = (int)n - 1;
  G.jrow_ptr.assign(G.m + 1, 0);
  for (int a = 0; a < G.m; ++a) {
    const int i = G.bus_of_node[a];
    for (int p = G.rowptr[i]; p < G.rowptr[i + 1]; ++p)
      if (G.col[p] != G.slack) { G.jcol.push_back(G.node_of_bus[G.col[p]]); G.jypos.push_back(p); }
    G.jrow_ptr[a + 1] = (int)G.jcol.size();
  }
  Opts O;
  O.variant = variant; O.max_iter = max_iter; O.tol = tol; O.qlim = qlimits_enabled; O.qlim_start = qlimit_start_iter;
  const int nt = std::max(1, std::min<int>(nthreads, (int)std::max<int64_t>(B, 1)));
  const int64_t per = (B + nt - 1) / nt;          // contiguous chunks like Iterators.partition(eachindex(cases), cld(n, cap))
  std::vector<long long> nnz(nt, 0);
  const auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int t) {
    Worker W(G, O);
    for (int64_t s = t * per; s < std::min<int64_t>(B, (t + 1) * per); ++s) {
      Out R;
      W.run(obr + optr[s], optr[s + 1] - optr[s], R, V ? reinterpret_cast<cd*>(V) + (size_t)s * n : nullptr, types ? types + (size_t)s * n : nullptr);
      if (converged) converged[s] = (uint8_t)R.conv;
      if (iterations) iterations[s] = R.iters;
      if (status) status[s] = R.status;
      if (n_switch) n_switch[s] = R.nsw;
      if (vmin) vmin[s] = R.vmin;
      if (vmax) vmax[s] = R.vmax;
      if (max_loading) max_loading[s] = R.load;
    }
    nnz[t] = W.have_sym ? W.sym.nnz_lu : 0;
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  if (seconds_out) {
    seconds_out[0] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    seconds_out[1] = (double)*std::max_element(nnz.begin(), nnz.end());
  }
  return 0;
}

}  // extern "C"
