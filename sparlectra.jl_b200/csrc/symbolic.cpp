// Host-side symbolic analysis: fill-reducing ordering (minimum degree / BFS nested
// dissection), elimination tree, column structures, relaxed supernodes, panel fronts sized for
// one CTA's shared memory, assembly tree, level sets and all index maps.  See symbolic.h.
#include "symbolic.h"

#include <algorithm>
#include <cassert>
#include <cstdio>
#include <functional>
#include <numeric>
#include <queue>
#include <sstream>
#include <stdexcept>

namespace sblx {

using Adj = std::vector<std::vector<int>>;

// ---------------------------------------------------------------------------------------------
// minimum degree on an explicit elimination graph (exact external degree, lazy heap)
// ---------------------------------------------------------------------------------------------
static void order_md(const Adj& adj_in, std::vector<int>& order) {
  const int n = (int)adj_in.size();
  Adj adj = adj_in;
  for (auto& a : adj) std::sort(a.begin(), a.end());
  std::vector<char> alive(n, 1);
  using QE = std::pair<int, int>;
  std::priority_queue<QE, std::vector<QE>, std::greater<QE>> heap;
  for (int i = 0; i < n; ++i) heap.push({(int)adj[i].size(), i});
  order.clear();
  order.reserve(n);
  std::vector<int> tmp;
  while (!heap.empty()) {
    auto [d, v] = heap.top();
    heap.pop();
    if (!alive[v] || d != (int)adj[v].size()) continue;
    alive[v] = 0;
    order.push_back(v);
    const std::vector<int> nb = adj[v];
    for (int a : nb) {
      // adj[a] = (adj[a] U nb) \ {a, v}
      tmp.clear();
      std::set_union(adj[a].begin(), adj[a].end(), nb.begin(), nb.end(), std::back_inserter(tmp));
      tmp.erase(std::remove_if(tmp.begin(), tmp.end(), [&](int x) { return x == a || x == v; }), tmp.end());
      adj[a].swap(tmp);
      heap.push({(int)adj[a].size(), a});
    }
    std::vector<int>().swap(adj[v]);
  }
}

// ---------------------------------------------------------------------------------------------
// nested dissection by BFS level-structure separators (George's automatic ND), MD on the leaves
// ---------------------------------------------------------------------------------------------
struct NdCtx {
  const Adj* adj;
  std::vector<int> tag;    // subset membership stamp
  std::vector<int> lvl, lvl2, side;
  bool use_fband = true;
  int stamp = 0;
  int leaf = 48;
  std::vector<int>* out;
};

static void nd_leaf_md(NdCtx& c, const std::vector<int>& S) {
  // induced subgraph -> MD
  const int k = (int)S.size();
  std::vector<int> loc;  // via tag lookups: build map node->local using lvl as scratch
  for (int i = 0; i < k; ++i) c.lvl[S[i]] = i;
  const int st = ++c.stamp;
  for (int v : S) c.tag[v] = st;
  Adj sub(k);
  for (int i = 0; i < k; ++i)
    for (int u : (*c.adj)[S[i]])
      if (c.tag[u] == st) sub[i].push_back(c.lvl[u]);
  std::vector<int> ord;
  order_md(sub, ord);
  for (int i : ord) c.out->push_back(S[i]);
}

static void nd_rec(NdCtx& c, std::vector<int> S) {
  if ((int)S.size() <= c.leaf) {
    nd_leaf_md(c, S);
    return;
  }
  const Adj& adj = *c.adj;
  int st = ++c.stamp;
  for (int v : S) c.tag[v] = st;
  // connected components
  {
    std::vector<std::vector<int>> comps;
    const int seen = ++c.stamp;  // tag == st : unvisited member; tag == seen : visited
    for (int s : S) {
      if (c.tag[s] != st) continue;
      comps.emplace_back();
      auto& comp = comps.back();
      comp.push_back(s);
      c.tag[s] = seen;
      for (size_t q = 0; q < comp.size(); ++q)
        for (int u : adj[comp[q]])
          if (c.tag[u] == st) {
            c.tag[u] = seen;
            comp.push_back(u);
          }
    }
    if (comps.size() > 1) {
      for (auto& comp : comps) nd_rec(c, std::move(comp));
      return;
    }
  }
  st = ++c.stamp;
  for (int v : S) c.tag[v] = st;
  // BFS helper producing level numbers in c.lvl for members; returns order
  auto bfs = [&](int root, std::vector<int>& ord) {
    const int vis = ++c.stamp;
    ord.clear();
    ord.push_back(root);
    c.tag[root] = vis;
    c.lvl[root] = 0;
    for (size_t q = 0; q < ord.size(); ++q) {
      int v = ord[q];
      for (int u : adj[v])
        if (c.tag[u] == st) {
          c.tag[u] = vis;
          c.lvl[u] = c.lvl[v] + 1;
          ord.push_back(u);
        }
    }
    for (int v : ord) c.tag[v] = st;  // restore membership stamp
  };
  std::vector<int> ord;
  int root = S[0], far = S[0];
  int depth = -1;
  for (int rep = 0; rep < 4; ++rep) {  // pseudo-peripheral node
    bfs(root, ord);
    int last = ord.back();
    // among the last level pick the minimum-degree node
    int L = c.lvl[last];
    int best = last;
    for (int i = (int)ord.size() - 1; i >= 0 && c.lvl[ord[i]] == L; --i)
      if (adj[ord[i]].size() < adj[best].size()) best = ord[i];
    if (L <= depth) break;
    depth = L;
    far = root;
    root = best;
  }
  // Two separator families are tried and the smaller (balance-weighted, after trimming) one is kept:
  //  (1) a level of the BFS level structure rooted at the pseudo-peripheral node p (George's automatic ND);
  //  (2) a band {c, c+1} of f(v) = d(p, v) - d(q, v), q = the far end of the pseudo-diameter.  |f(u) - f(v)| <= 2
  //      along an edge, so two consecutive values always separate; on mesh-like grids the level sets of f run
  //      ACROSS the diameter and are much shorter than the ring-shaped BFS levels.
  const int total = (int)S.size();
  bfs(far, ord);
  std::vector<int> dq(total);
  for (int i = 0; i < total; ++i) c.lvl2[S[i]] = c.lvl[S[i]];
  bfs(root, ord);
  const int nl = c.lvl[ord.back()] + 1;
  if (nl < 3) {  // (near-)clique: no useful separator
    nd_leaf_md(c, S);
    return;
  }
  auto pick = [&](const std::vector<int>& cnt, int width, int& bestk, double& bestscore) {
    // cnt[k] = members with key k; separator = keys [k, k + width); returns the best k by size / balance
    const int nk = (int)cnt.size();
    bestk = -1;
    bestscore = 1e300;
    int cum = 0;
    for (int k = 0; k + width <= nk; ++k) {
      int sep = 0;
      for (int j = 0; j < width; ++j) sep += cnt[k + j];
      const int left = cum, right = total - cum - sep;
      cum += cnt[k];
      if (left == 0 || right == 0) continue;
      const double bal = (double)std::min(left, right) / std::max(1, std::max(left, right));
      if (bal < 0.35) continue;
      const double score = sep / (0.2 + bal);
      if (score < bestscore) {
        bestscore = score;
        bestk = k;
      }
    }
  };
  // key arrays: family 1 = level, family 2 = f shifted to >= 0
  std::vector<int> cnt1(nl, 0);
  for (int v : ord) cnt1[c.lvl[v]]++;
  int fmin = 1 << 30, fmax = -(1 << 30);
  for (int v : ord) {
    const int f = c.lvl[v] - c.lvl2[v];
    fmin = std::min(fmin, f);
    fmax = std::max(fmax, f);
  }
  std::vector<int> cnt2(fmax - fmin + 1, 0);
  for (int v : ord) cnt2[c.lvl[v] - c.lvl2[v] - fmin]++;
  int k1, k2;
  double s1, s2;
  pick(cnt1, 1, k1, s1);
  pick(cnt2, 2, k2, s2);
  if (k1 < 0) {
    int cum2 = 0;
    for (int l = 0; l < nl; ++l) {
      cum2 += cnt1[l];
      if (cum2 * 2 >= total) {
        k1 = std::min(std::max(l, 1), nl - 2);
        break;
      }
    }
  }
  // materialise a candidate: side[v] in c.side: 0 = A, 1 = separator, 2 = B; trimmed on both faces
  auto build = [&](int family, int k, std::vector<int>& A, std::vector<int>& B, std::vector<int>& Sep) -> double {
    A.clear(); B.clear(); Sep.clear();
    if (k < 0) return 1e300;
    const int width = family == 1 ? 1 : 2;
    for (int v : ord) {
      const int key = family == 1 ? c.lvl[v] : c.lvl[v] - c.lvl2[v] - fmin;
      c.side[v] = key < k ? 0 : key >= k + width ? 2 : 1;
    }
    for (int v : ord)      // separator nodes that do not touch B fall back to A
      if (c.side[v] == 1) {
        bool t = false;
        for (int u : adj[v])
          if (c.tag[u] == st && c.side[u] == 2) { t = true; break; }
        if (!t) c.side[v] = 0;
      }
    for (int v : ord)      // ... and those that do not touch A (any more) go to B
      if (c.side[v] == 1) {
        bool t = false;
        for (int u : adj[v])
          if (c.tag[u] == st && c.side[u] == 0) { t = true; break; }
        if (!t) c.side[v] = 2;
      }
    for (int v : ord) (c.side[v] == 0 ? A : c.side[v] == 2 ? B : Sep).push_back(v);
    if (A.empty() || B.empty()) return 1e300;
    const double bal = (double)std::min(A.size(), B.size()) / std::max(A.size(), B.size());
    return Sep.size() / (0.2 + bal);
  };
  std::vector<int> A, B, Sep, A2, B2, Sep2;
  const double f1 = build(1, k1, A, B, Sep);
  const double f2 = c.use_fband ? build(2, k2, A2, B2, Sep2) : 1e300;
  if (f2 < f1) {
    A.swap(A2);
    B.swap(B2);
    Sep.swap(Sep2);
  }
  if (A.empty() || B.empty()) {
    nd_leaf_md(c, S);
    return;
  }
  if (getenv("SBLX_ND_DEBUG") && total > 600) fprintf(stderr, "nd: |S|=%d -> A=%zu B=%zu Sep=%zu (level %.1f, fband %.1f)\n", total, A.size(), B.size(), Sep.size(), f1, f2);
  std::vector<int> sepcopy = Sep;
  nd_rec(c, std::move(A));
  nd_rec(c, std::move(B));
  for (int v : sepcopy) c.out->push_back(v);
}

static void order_nd(const Adj& adj, int leaf, std::vector<int>& order) {
  const int n = (int)adj.size();
  NdCtx c;
  c.adj = &adj;
  c.tag.assign(n, 0);
  c.lvl.assign(n, 0);
  c.lvl2.assign(n, 0);
  c.side.assign(n, 0);
  c.use_fband = getenv("SBLX_ND_NO_FBAND") == nullptr;
  c.leaf = leaf;
  order.clear();
  order.reserve(n);
  c.out = &order;
  std::vector<int> all(n);
  std::iota(all.begin(), all.end(), 0);
  nd_rec(c, std::move(all));
}

// ---------------------------------------------------------------------------------------------
// column structures + elimination tree for a given order
// ---------------------------------------------------------------------------------------------
struct ColStruct {
  std::vector<std::vector<int>> st;  // st[e] = sorted elimination indices > e in column e of L
  std::vector<int> parent;
};

static void column_structures(const Adj& adj, const std::vector<int>& order, ColStruct& cs) {
  const int n = (int)adj.size();
  std::vector<int> pos(n);
  for (int e = 0; e < n; ++e) pos[order[e]] = e;
  cs.st.assign(n, {});
  cs.parent.assign(n, -1);
  std::vector<std::vector<int>> children(n);
  std::vector<int> tmp;
  for (int e = 0; e < n; ++e) {
    std::vector<int>& s = cs.st[e];
    for (int u : adj[order[e]])
      if (pos[u] > e) s.push_back(pos[u]);
    std::sort(s.begin(), s.end());
    for (int c : children[e]) {
      tmp.clear();
      const auto& sc = cs.st[c];
      // union(s, sc \ {e})
      std::set_union(s.begin(), s.end(), sc.begin(), sc.end(), std::back_inserter(tmp));
      tmp.erase(std::remove(tmp.begin(), tmp.end(), e), tmp.end());
      s.swap(tmp);
    }
    if (!s.empty()) {
      cs.parent[e] = s.front();
      children[s.front()].push_back(e);
    }
  }
}

static double exact_block_ops(const ColStruct& cs, int64_t& nnzL) {
  double ops = 0;
  nnzL = 0;
  for (auto& s : cs.st) {
    double c = (double)s.size();
    ops += c * c;
    nnzL += (int64_t)s.size();
  }
  return ops;
}

// ---------------------------------------------------------------------------------------------
struct SNode {
  std::vector<int> piv;  // elimination indices (old numbering), in pivot order
  std::vector<int> st;   // structure (old numbering), sorted
  int parent = -1;
  std::vector<int> children;
  bool dead = false;
};

std::string Symbolic::summary() const {
  std::ostringstream o;
  o << "ordering=" << ordering_name << " m=" << m << " nnzB=" << nnzB << " nnzL_blocks=" << nnzL_blocks
    << " nnzLU_scalar=" << nnzLU_scalar << " panels=" << panels.size() << " levels=" << nlevels
    << " block_ops=" << block_ops << " (exact " << block_ops_exact << ")"
    << " u_blocks=" << u_blocks << " c_blocks=" << c_blocks << " max_w=" << max_w << " max_h=" << max_h
    << " max_smem=" << max_smem_bytes;
  return o.str();
}

static void build_from_order(int m, const std::vector<int64_t>& rowptr, const std::vector<int>& colidx, const Adj& adj,
                             const std::vector<int>& order, const SymbolicOptions& opt, Symbolic& S) {
  ColStruct cs;
  column_structures(adj, order, cs);
  S.m = m;
  S.block_ops_exact = exact_block_ops(cs, S.nnzL_blocks);
  S.nnzLU_scalar = 4 * (2 * S.nnzL_blocks + m);

  // ---- supernodes: e joins e-1 when parent[e-1]==e and struct(e-1) = {e} U struct(e)
  std::vector<SNode> sn;
  std::vector<int> sn_of(m, -1);
  for (int e = 0; e < m; ++e) {
    bool join = e > 0 && cs.parent[e - 1] == e && cs.st[e - 1].size() == cs.st[e].size() + 1;
    if (!join) sn.emplace_back();
    sn.back().piv.push_back(e);
    sn_of[e] = (int)sn.size() - 1;
  }
  for (auto& s : sn) s.st = cs.st[s.piv.back()];
  for (int k = 0; k < (int)sn.size(); ++k) {
    if (!sn[k].st.empty()) {
      sn[k].parent = sn_of[sn[k].st.front()];
      sn[sn[k].parent].children.push_back(k);
    }
  }
  // ---- relaxed amalgamation (bottom-up; children have smaller ids than parents)
  for (int p = 0; p < (int)sn.size(); ++p) {
    if (sn[p].children.empty()) continue;
    std::vector<int> ch = sn[p].children;
    std::sort(ch.begin(), ch.end(), [&](int a, int b) {
      return sn[a].piv.size() + sn[a].st.size() > sn[b].piv.size() + sn[b].st.size();
    });
    std::vector<int> kept;
    for (int c : ch) {
      double wc = (double)sn[c].piv.size(), hc = (double)sn[c].st.size();
      double wp = (double)sn[p].piv.size(), hp = (double)sn[p].st.size();
      double zeros = 2.0 * wc * (wp + hp - hc);
      double area = wc * (wc + 2 * hc) + wp * (wp + 2 * hp);
      // the unconditional small-front rule must not widen a pivot block past what the lean small-front kernels
      // take (8 buses): a tiny front that lands in a many-warp class costs far more than the launch it saves
      bool small = (wc + wp + hp) <= opt.relax_small && (wc + wp) <= opt.relax_small_max_w;
      bool ok = small || zeros <= opt.relax_zero_frac * area;
      if (ok && (int)(wc + wp) <= 4 * opt.max_panel_width) {
        // merge c into p: child's pivots first
        std::vector<int> np = sn[c].piv;
        np.insert(np.end(), sn[p].piv.begin(), sn[p].piv.end());
        sn[p].piv.swap(np);
        for (int g : sn[c].children) {
          sn[g].parent = p;
          kept.push_back(g);
        }
        sn[c].dead = true;
      } else {
        kept.push_back(c);
      }
    }
    sn[p].children.swap(kept);
  }
  // ---- final numbering: postorder over live supernodes
  std::vector<int> roots;
  for (int k = 0; k < (int)sn.size(); ++k)
    if (!sn[k].dead && sn[k].parent < 0) roots.push_back(k);
  std::vector<int> post;
  post.reserve(sn.size());
  {
    std::vector<std::pair<int, size_t>> stack;
    for (int r : roots) {
      stack.push_back({r, 0});
      while (!stack.empty()) {
        auto& [k, ci] = stack.back();
        if (ci < sn[k].children.size()) {
          int c = sn[k].children[ci++];
          stack.push_back({c, 0});
        } else {
          post.push_back(k);
          stack.pop_back();
        }
      }
    }
  }
  std::vector<int> newidx(m, -1);  // old elimination index -> final elimination index
  {
    int f = 0;
    for (int k : post)
      for (int e : sn[k].piv) newidx[e] = f++;
    assert(f == m);
  }
  S.perm.assign(m, 0);
  S.iperm.assign(m, 0);
  for (int e = 0; e < m; ++e) {
    S.perm[newidx[e]] = order[e];
    S.iperm[order[e]] = newidx[e];
  }
  // ---- panels
  S.panels.clear();
  S.struct_nodes.clear();
  std::vector<int> panel_of(m, -1);
  S.max_w = S.max_h = 0;
  for (int k : post) {
    const int w = (int)sn[k].piv.size();
    std::vector<int> st(sn[k].st.size());
    for (size_t i = 0; i < st.size(); ++i) st[i] = newidx[sn[k].st[i]];
    std::sort(st.begin(), st.end());
    const int h = (int)st.size();
    const int first = newidx[sn[k].piv[0]];
    // widest panel that fits at the START of the chain (largest structure)
    auto fits = [&](int wk, int hk) { return wk <= opt.max_panel_width && panel_smem_bytes(wk, hk) <= opt.smem_budget_bytes; };
    int done = 0;
    while (done < w) {
      int rem = w - done;
      int wk = std::min(rem, opt.max_panel_width);
      while (wk > 1 && !fits(wk, (rem - wk) + h)) --wk;
      if (!fits(wk, (rem - wk) + h))
        throw std::runtime_error("sblx symbolic: a front does not fit the shared-memory panel budget");
      // balance the remaining chain
      int np = (rem + wk - 1) / wk;
      wk = (rem + np - 1) / np;
      while (wk > 1 && !fits(wk, (rem - wk) + h)) --wk;
      Panel P;
      P.first = first + done;
      P.w = wk;
      P.h = (rem - wk) + h;
      P.struct_ptr = (int64_t)S.struct_nodes.size();
      for (int g = first + done + wk; g < first + w; ++g) S.struct_nodes.push_back(g);
      for (int g : st) S.struct_nodes.push_back(g);
      for (int g = P.first; g < P.first + P.w; ++g) panel_of[g] = (int)S.panels.size();
      S.max_w = std::max(S.max_w, P.w);
      S.max_h = std::max(S.max_h, P.h);
      S.panels.push_back(P);
      done += wk;
    }
  }
  const int np = (int)S.panels.size();
  // parents / children / levels
  std::vector<std::vector<int>> children(np);
  for (int p = 0; p < np; ++p) {
    Panel& P = S.panels[p];
    if (P.h > 0) {
      P.parent = panel_of[S.struct_nodes[P.struct_ptr]];
      assert(P.parent > p);
      children[P.parent].push_back(p);
    }
  }
  S.child_list.clear();
  for (int p = 0; p < np; ++p) {
    Panel& P = S.panels[p];
    P.child_ptr = (int64_t)S.child_list.size();
    P.nchild = (int64_t)children[p].size();
    int lv = 0;
    for (int c : children[p]) {
      S.child_list.push_back(c);
      lv = std::max(lv, S.panels[c].level + 1);
    }
    P.level = lv;
  }
  S.nlevels = 0;
  for (auto& P : S.panels) S.nlevels = std::max(S.nlevels, P.level + 1);
  // local index lookup inside a panel
  auto local_index = [&](const Panel& P, int g) -> int {
    if (g >= P.first && g < P.first + P.w) return g - P.first;
    const int* b = S.struct_nodes.data() + P.struct_ptr;
    const int* e = b + P.h;
    const int* it = std::lower_bound(b, e, g);
    if (it == e || *it != g) return -1;
    return P.w + (int)(it - b);
  };
  // rel maps (child structure -> parent local) and kpiv
  S.rel_ptr.assign(np + 1, 0);
  S.rel.clear();
  S.kpiv.assign(np, 0);
  for (int p = 0; p < np; ++p) {
    const Panel& P = S.panels[p];
    S.rel_ptr[p] = (int64_t)S.rel.size();
    if (P.parent >= 0) {
      const Panel& Q = S.panels[P.parent];
      int k = 0;
      for (int a = 0; a < P.h; ++a) {
        int li = local_index(Q, S.struct_nodes[P.struct_ptr + a]);
        if (li < 0) throw std::runtime_error("sblx symbolic: child structure not contained in parent front");
        if (li < Q.w) {
          if (a != k) throw std::runtime_error("sblx symbolic: parent pivots not leading in child structure");
          ++k;
        }
        S.rel.push_back(li);
      }
      S.kpiv[p] = k;
    }
  }
  S.rel_ptr[np] = (int64_t)S.rel.size();
  // inverse gather maps per (parent, child)
  S.inv_ptr.assign(S.child_list.size() + 1, 0);
  S.inv.clear();
  for (int p = 0; p < np; ++p) {
    const Panel& P = S.panels[p];
    for (int ci = 0; ci < P.nchild; ++ci) {
      int c = S.child_list[P.child_ptr + ci];
      const Panel& C = S.panels[c];
      S.inv_ptr[P.child_ptr + ci] = (int64_t)S.inv.size();
      size_t base = S.inv.size();
      S.inv.resize(base + P.h, -1);
      for (int a = S.kpiv[c]; a < C.h; ++a) {
        int li = S.rel[S.rel_ptr[c] + a];
        S.inv[base + (li - P.w)] = a;
      }
    }
  }
  S.inv_ptr[S.child_list.size()] = (int64_t)S.inv.size();
  // assembly lists of original Jacobian blocks
  {
    std::vector<std::vector<std::pair<int, int>>> lists(np);
    S.nnzB = rowptr[m];
    for (int i = 0; i < m; ++i) {
      for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        int j = colidx[k];
        int fi = S.iperm[i], fj = S.iperm[j];
        int owner = panel_of[std::min(fi, fj)];
        const Panel& P = S.panels[owner];
        int li = local_index(P, fi), lj = local_index(P, fj);
        if (li < 0 || lj < 0) throw std::runtime_error("sblx symbolic: Jacobian block outside its front");
        int nf = P.w + P.h;
        int dst = (li < P.w) ? (li * nf + lj) : (P.w * nf + lj * P.h + (li - P.w));   // structure rows: pivot-major [w][h]
        lists[owner].push_back({(int)k, dst});
      }
    }
    S.asm_src.clear();
    S.asm_dst.clear();
    for (int p = 0; p < np; ++p) {
      S.panels[p].asm_ptr = (int64_t)S.asm_src.size();
      S.panels[p].asm_cnt = (int64_t)lists[p].size();
      for (auto& pr : lists[p]) {
        S.asm_src.push_back(pr.first);
        S.asm_dst.push_back(pr.second);
      }
    }
  }
  // U offsets, op counts
  S.u_blocks = 0;
  S.block_ops = 0;
  S.max_smem_bytes = 0;
  for (auto& P : S.panels) {
    P.u_off = S.u_blocks;
    S.u_blocks += (int64_t)P.w * (P.w + P.h);
    const double w = P.w, h = P.h, nf = w + h;
    for (int p = 0; p < P.w; ++p) S.block_ops += (w - p - 1) * (nf - p - 1) + h * (w - p - 1);
    S.block_ops += h * h * w;
    S.max_smem_bytes = std::max(S.max_smem_bytes, panel_smem_bytes(P.w, P.h));
  }
  // level lists
  S.level_ptr.assign(S.nlevels + 1, 0);
  for (auto& P : S.panels) S.level_ptr[P.level + 1]++;
  for (int l = 0; l < S.nlevels; ++l) S.level_ptr[l + 1] += S.level_ptr[l];
  S.level_panels.assign(np, 0);
  {
    std::vector<int> fill(S.level_ptr.begin(), S.level_ptr.end() - 1);
    for (int p = 0; p < np; ++p) S.level_panels[fill[S.panels[p].level]++] = p;
  }
  // update-matrix arena: first-fit with liveness by level (children freed after the parent's level)
  {
    struct Free { int64_t off, len; };
    std::vector<Free> fl;
    int64_t top = 0;
    auto alloc = [&](int64_t len) -> int64_t {
      for (size_t i = 0; i < fl.size(); ++i)
        if (fl[i].len >= len) {
          int64_t o = fl[i].off;
          fl[i].off += len;
          fl[i].len -= len;
          if (fl[i].len == 0) fl.erase(fl.begin() + i);
          return o;
        }
      int64_t o = top;
      top += len;
      return o;
    };
    auto release = [&](int64_t off, int64_t len) {
      if (len == 0) return;
      size_t i = 0;
      while (i < fl.size() && fl[i].off < off) ++i;
      fl.insert(fl.begin() + i, Free{off, len});
      // coalesce
      if (i + 1 < fl.size() && fl[i].off + fl[i].len == fl[i + 1].off) {
        fl[i].len += fl[i + 1].len;
        fl.erase(fl.begin() + i + 1);
      }
      if (i > 0 && fl[i - 1].off + fl[i - 1].len == fl[i].off) {
        fl[i - 1].len += fl[i].len;
        fl.erase(fl.begin() + i);
      }
      if (!fl.empty() && fl.back().off + fl.back().len == top) {
        top = fl.back().off;
        fl.pop_back();
      }
    };
    auto csize = [](const Panel& P) -> int64_t { return (int64_t)P.h * P.h + (P.h + 1) / 2; };
    int64_t peak = 0;
    for (int l = 0; l < S.nlevels; ++l) {
      for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
        Panel& P = S.panels[S.level_panels[q]];
        P.c_off = alloc(csize(P));
        peak = std::max(peak, top);
      }
      for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
        const Panel& P = S.panels[S.level_panels[q]];
        for (int ci = 0; ci < P.nchild; ++ci) {
          const Panel& C = S.panels[S.child_list[P.child_ptr + ci]];
          release(C.c_off, csize(C));
        }
      }
    }
    S.c_blocks = std::max<int64_t>(peak, 1);
  }
  // gather-style ingest lists (need the final arena offsets)
  {
    S.ing_ptr.clear();
    S.ing_src.clear();
    S.ing_pairs.clear();
    S.ing_xptr.assign(1, 0);
    S.ing_xsrc.clear();
    S.vec_ptr.assign(m + 1, 0);
    S.vec_src.clear();
    S.reg_words.clear();
    S.reg_vptr.clear();
    S.reg_vsrc.clear();
    std::vector<std::vector<int>> vsrc(m);
    for (int p = 0; p < np; ++p) {
      Panel& P = S.panels[p];
      const int w = P.w, h = P.h, nf = w + h;
      const int nblk = w * nf + h * w;
      std::vector<std::vector<int>> src(nblk);
      for (int64_t k = 0; k < P.asm_cnt; ++k) src[S.asm_dst[P.asm_ptr + k]].push_back(~S.asm_src[P.asm_ptr + k]);
      for (int ci = 0; ci < P.nchild; ++ci) {
        const int c = S.child_list[P.child_ptr + ci];
        const Panel& C = S.panels[c];
        const int hc = C.h, kc = S.kpiv[c];
        const int* rel = &S.rel[S.rel_ptr[c]];
        for (int a = 0; a < hc; ++a)
          for (int b = 0; b < hc; ++b) {
            if (a >= kc && b >= kc) continue;
            const int la = rel[a], lb = rel[b];
            const int dst = (la < w) ? (la * nf + lb) : (w * nf + lb * h + (la - w));
            const int64_t off = C.c_off + (int64_t)a * hc + b;
            if (off > 0x7fffffffLL) throw std::runtime_error("sblx symbolic: update arena too large for 32-bit ingest offsets");
            src[dst].push_back((int)off);
          }
        const int64_t vbase = 2 * (C.c_off + (int64_t)hc * hc);
        for (int a = 0; a < kc; ++a) vsrc[P.first + rel[a]].push_back((int)(vbase + a));
      }
      P.reg_base = -1;
      if (nf <= 8) {       // register-resident class: sources of the full nf x nf front, row-major, and per-row vector sources
        std::vector<std::vector<int>> full((size_t)nf * nf), fv(nf);
        for (int k = 0; k < nblk; ++k) {
          int r, c2;
          if (k < w * nf) { r = k / nf; c2 = k - r * nf; }
          else { const int q = k - w * nf, lj = q / h, li = q - lj * h; r = w + li; c2 = lj; }
          full[(size_t)r * nf + c2] = src[k];
        }
        for (int ci = 0; ci < P.nchild; ++ci) {
          const int c = S.child_list[P.child_ptr + ci];
          const Panel& C = S.panels[c];
          const int hc = C.h, kc = S.kpiv[c];
          const int* rel = &S.rel[S.rel_ptr[c]];
          for (int a = kc; a < hc; ++a)
            for (int b = kc; b < hc; ++b) full[(size_t)rel[a] * nf + rel[b]].push_back((int)(C.c_off + (int64_t)a * hc + b));
          const int64_t vbase = 2 * (C.c_off + (int64_t)hc * hc);
          for (int a = 0; a < hc; ++a) fv[rel[a]].push_back((int)(vbase + a));
        }
        P.reg_base = (int64_t)S.reg_words.size();
        for (auto& f : full) {
          int word;
          if (f.empty()) word = ING_NONE;
          else if (f.size() == 1) word = f[0];
          else {
            word = ING_MULTI + (int)S.ing_xptr.size() - 1;
            for (int v : f) S.ing_xsrc.push_back(v);
            S.ing_xptr.push_back((int)S.ing_xsrc.size());
          }
          S.reg_words.push_back(word);
        }
        P.reg_vbase = (int64_t)S.reg_vptr.size();
        for (int r = 0; r < nf; ++r) {
          S.reg_vptr.push_back((int)S.reg_vsrc.size());
          for (int v : fv[r]) S.reg_vsrc.push_back(v);
        }
        S.reg_vptr.push_back((int)S.reg_vsrc.size());
      }
      P.ing_base = (int64_t)S.ing_ptr.size();
      P.ingf_base = (int64_t)S.ing_pairs.size() / 2;
      for (int k = 0; k < nblk; ++k) {
        S.ing_ptr.push_back((int)S.ing_src.size());
        for (int v : src[k]) S.ing_src.push_back(v);
      }
      auto emit = [&](int k) {
        int word;
        if (src[k].empty()) word = ING_NONE;
        else if (src[k].size() == 1) {
          if (src[k][0] >= ING_MULTI) throw std::runtime_error("sblx symbolic: update arena too large for the ingest encoding");
          word = src[k][0];
        } else {
          word = ING_MULTI + (int)S.ing_xptr.size() - 1;
          for (int v : src[k]) S.ing_xsrc.push_back(v);
          S.ing_xptr.push_back((int)S.ing_xsrc.size());
        }
        S.ing_pairs.push_back(word);
        S.ing_pairs.push_back(k);
      };
      for (int r = 0; r < w; ++r)
        for (int c2 = 0; c2 < w; ++c2) emit(r * nf + c2);                 // pivot block
      for (int r = 0; r < w; ++r)
        for (int j = 0; j < h; ++j) emit(r * nf + w + j);                 // pivot rows right of it
      for (int li = 0; li < h; ++li)
        for (int lj = 0; lj < w; ++lj) emit(w * nf + lj * h + li);       // structure rows, pivot column fastest
      S.ing_ptr.push_back((int)S.ing_src.size());
    }
    for (int e2 = 0; e2 < m; ++e2) {
      S.vec_ptr[e2] = (int)S.vec_src.size();
      for (int v : vsrc[e2]) S.vec_src.push_back(v);
    }
    S.vec_ptr[m] = (int)S.vec_src.size();
  }
}

void analyze(int m, const std::vector<int64_t>& rowptr, const std::vector<int>& colidx, const SymbolicOptions& opt,
             Symbolic& out) {
  Adj adj(m);
  for (int i = 0; i < m; ++i)
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (colidx[k] != i) adj[i].push_back(colidx[k]);
  std::vector<int> ord_md, ord_nd;
  Symbolic best;
  bool have = false;
  auto consider = [&](const std::vector<int>& ord, const char* name) {
    Symbolic S;
    build_from_order(m, rowptr, colidx, adj, ord, opt, S);
    S.ordering_name = name;
    // cost model: arithmetic + a per-level launch charge.  A level costs two launches (~10 us) shared by all the
    // scenarios in flight, a block multiply-add ~0.04 ns per scenario; an N-1 study of a grid with m buses keeps
    // about m scenarios in flight, hence ~2e5 / m block operations per level (measured: on a 2 869-bus
    // transmission-like grid minimum degree with 82 levels beats nested dissection with 12 by 27 %).
    const double level_cost = std::min(2.0e4, std::max(10.0, 2.0e5 / std::max(1, m)));
    auto cost = [level_cost](const Symbolic& s) { return s.block_ops + level_cost * s.nlevels; };
    if (!have || cost(S) < cost(best)) {
      best = std::move(S);
      have = true;
    }
  };
  if (opt.ordering == 0 || opt.ordering == 1) {
    order_md(adj, ord_md);
    consider(ord_md, "md");
  }
  if (opt.ordering == 0 || opt.ordering == 2) {
    order_nd(adj, opt.nd_leaf, ord_nd);
    consider(ord_nd, "nd");
  }
  out = std::move(best);
}

}  // namespace sblx

#ifdef SBLX_SYMBOLIC_MAIN
// stand-alone probe: tiled grid rows x cols
int main(int argc, char** argv) {
  int rows = argc > 1 ? atoi(argv[1]) : 100, cols = argc > 2 ? atoi(argv[2]) : 100;
  int ordering = argc > 3 ? atoi(argv[3]) : 0;
  int n = rows * cols;
  std::vector<std::vector<int>> a(n);
  auto id = [&](int r, int c) { return r * cols + c; };
  auto edge = [&](int x, int y) { a[x].push_back(y); a[y].push_back(x); };
  for (int r = 0; r < rows; ++r) for (int c = 0; c + 1 < cols; ++c) edge(id(r, c), id(r, c + 1));
  for (int r = 0; r + 1 < rows; ++r) for (int c = 0; c < cols; ++c) edge(id(r, c), id(r + 1, c));
  for (int r = 0; r + 1 < rows; ++r) for (int c = 0; c + 1 < cols; ++c) edge(id(r, c), id(r + 1, c + 1));
  // drop node 0 (slack)
  int m = n - 1;
  std::vector<int64_t> rp(m + 1, 0);
  std::vector<int> ci;
  for (int i = 1; i < n; ++i) {
    std::vector<int> row;
    row.push_back(i - 1);
    for (int u : a[i]) if (u != 0) row.push_back(u - 1);
    std::sort(row.begin(), row.end());
    for (int u : row) ci.push_back(u);
    rp[i] = (int64_t)ci.size();
  }
  sblx::SymbolicOptions opt;
  opt.ordering = ordering;
  if (argc > 4) opt.max_panel_width = atoi(argv[4]);
  if (argc > 5) opt.relax_zero_frac = atof(argv[5]);
  if (argc > 6) opt.relax_small = atoi(argv[6]);
  sblx::Symbolic S;
  sblx::analyze(m, rp, ci, opt, S);
  printf("%s\n", S.summary().c_str());
  // histogram of panels by smem bytes and op share
  double ops_by[4] = {0, 0, 0, 0};
  int cnt_by[4] = {0, 0, 0, 0};
  double cread = 0;
  for (auto& P : S.panels) {
    int b = sblx::panel_smem_bytes(P.w, P.h);
    int k = b <= 6144 ? 0 : b <= 24576 ? 1 : b <= 73728 ? 2 : 3;
    double w = P.w, h = P.h, nf = w + h, o = h * h * w;
    for (int p = 0; p < P.w; ++p) o += (w - p - 1) * (nf - p - 1) + h * (w - p - 1);
    ops_by[k] += o;
    cnt_by[k]++;
    cread += 2.0 * h * h;
  }
  for (int k = 0; k < 4; ++k) printf("class %d: panels %d  op share %.3f\n", k, cnt_by[k], ops_by[k] / S.block_ops);
  printf("C write+read blocks %.0f (%.1f MB)  U blocks %lld (%.1f MB)\n", cread, cread * 32 / 1e6, (long long)S.u_blocks,
         S.u_blocks * 32 / 1e6);
  std::vector<int> per_level(S.nlevels, 0);
  for (auto& P : S.panels) per_level[P.level]++;
  printf("panels per level:");
  for (int l = 0; l < S.nlevels; ++l) printf(" %d", per_level[l]);
  printf("\n");
  return 0;
}
#endif
