// Host-side symbolic analysis for the batched block-sparse LU (computed once per grid).
//
// The Newton system J dx = -F of the rectangular formulation
// (reference: src/powerflow_rectangular/rectangular_jacobian_builders.jl:212-429) has one
// 2x2 real block per Ybus non-zero between two non-slack buses.  All scenarios of one grid
// share one superset block pattern (PV<->PQ switching and branch outages only change
// values), so ordering, elimination tree, supernodes, panel fronts, level sets and every
// index map are computed here once and reused by every scenario and every iteration -
// what the reference's `umfpack_reuse` backend does per solve
// (rectangular_factorized_solver.jl:82-132) hoisted to "once per grid".
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sblx {

struct SymbolicOptions {
  int ordering = 0;            // 0 = auto (pick cheaper of ND / MD), 1 = minimum degree, 2 = nested dissection
  int smem_budget_bytes = 200 * 1024;  // panel budget per front (dynamic shared memory)
  int max_panel_width = 16;    // cap on pivots per panel (the pivot block is factored in one warp's registers: 32 scalar columns)
  int relax_small = 12;        // always merge a child when the merged front has <= this many nodes
  int relax_small_max_w = 4;   // ... as long as the merged pivot block stays within the small-front kernels' width
  double relax_zero_frac = 0.25;  // else merge when introduced explicit zeros <= frac * merged front area
  int nd_leaf = 48;            // nested dissection stops at subgraphs of this many nodes (ordered by MD)
};

// One panel front: `w` pivot nodes (contiguous in the final elimination numbering
// [first, first+w)) and `h` structure nodes (rows/cols of the Schur complement).
struct Panel {
  int first = 0, w = 0, h = 0;
  int parent = -1;
  int level = 0;
  int cls = 0;                 // size class (kernel configuration)
  int64_t u_off = 0;           // offset (in 2x2 blocks) of the stored U panel  w x (w+h)
  int64_t c_off = 0;           // offset (blocks) of the update matrix h x h (+ update vector) in the stack arena
  int64_t struct_ptr = 0;      // into struct_nodes (h entries, ascending final elimination index)
  int64_t child_ptr = 0, nchild = 0;  // into child_list
  int64_t asm_ptr = 0, asm_cnt = 0;   // original Jacobian blocks assembled here
  int64_t ing_base = 0;               // first entry of this panel in ing_ptr (w*(w+h) + h*w + 1 entries)
  int64_t ingf_base = 0;              // first (source, destination) pair of this panel in ing_pairs (w*(w+h) + h*w pairs)
  int64_t reg_base = -1;              // register-resident class (w + h <= 8): first word in reg_words ((w+h)^2 words), else -1
  int64_t reg_vbase = 0;              //   ... and first entry in reg_vptr (w + h + 1 entries)
};

struct Symbolic {
  int m = 0;                   // non-slack nodes
  int64_t nnzB = 0;            // Jacobian 2x2 blocks (pattern entries between non-slack nodes)
  std::vector<int> perm;       // final elimination index -> node
  std::vector<int> iperm;      // node -> final elimination index
  std::vector<Panel> panels;   // topological order (children before parents)
  std::vector<int> struct_nodes;      // concatenated structure lists (final elimination indices)
  std::vector<int> child_list;        // concatenated child panel ids
  // per (child panel): map child-structure position -> parent local index (0..w+h)
  std::vector<int64_t> rel_ptr;       // per panel: offset into rel (size h of that panel)
  std::vector<int> rel;               // child's structure position -> local index in ITS parent
  std::vector<int> kpiv;              // per panel: number of leading structure entries that are pivots of the parent
  // per (parent, child) gather map: parent structure position -> child structure position or -1
  std::vector<int64_t> inv_ptr;       // per entry of child_list: offset into inv (size h of the parent)
  std::vector<int> inv;
  // assembly of original entries: Jacobian block slot -> destination offset (in blocks) inside the shared-memory panel
  std::vector<int> asm_src, asm_dst;
  // gather-style ingest: per panel block (shared-memory order: pivot rows [w][nf], then structure rows pivot-major
  // [w][h]) the CSR list of its sources; src < 0: Jacobian slot ~src, src >= 0: block offset in the update arena
  std::vector<int> ing_ptr, ing_src;
  // the same lists in the form the fused front kernel reads: one (source, destination) pair per panel block IN THE
  // ORDER THE THREADS WALK (pivot block row-major; pivot rows right of it; structure rows with the pivot column
  // fastest = along the rows of the children's update matrices), so the kernel does no index arithmetic.
  // source: ING_NONE (explicit zero), the single source (75 % of the blocks: one dependent index load instead of
  // two), or ING_MULTI + x with the sources in ing_xsrc[ing_xptr[x] .. ing_xptr[x+1]); destination: block index in
  // the shared-memory panel
  std::vector<int> ing_pairs, ing_xptr, ing_xsrc;
  // register-resident tiny fronts (w + h <= 8): source word of EVERY block of the full front, row-major (pivot rows,
  // then structure rows incl. the Schur part = the children's contributions), same word encoding; and per front row
  // the update-vector / rhs contributions of the children (double2 indices into the arena)
  std::vector<int> reg_words, reg_vptr, reg_vsrc;
  // rhs of the pivots: per elimination index the CSR list of update-vector entries (double2 index in the arena)
  std::vector<int> vec_ptr, vec_src;
  // levels
  int nlevels = 0;
  std::vector<int> level_ptr;         // level -> range in level_panels
  std::vector<int> level_panels;      // panels sorted by (level, class)
  // sizes
  int64_t u_blocks = 0;        // total U storage per scenario (blocks)
  int64_t c_blocks = 0;        // peak update-matrix arena per scenario (blocks)
  int64_t nnzL_blocks = 0;     // strictly-lower blocks of the (unrelaxed) factor
  int64_t nnzLU_scalar = 0;    // scalar nnz(L+U) of the unrelaxed factor: 4 * (2*nnzL_blocks + m)
  double block_ops = 0;        // 2x2 block multiply-adds in the numeric factorisation (incl. relaxed zeros)
  double block_ops_exact = 0;  // same for the unrelaxed factor
  int max_w = 0, max_h = 0;
  int max_smem_bytes = 0;
  std::string ordering_name;
  std::string summary() const;
};

// Block pattern input: CSR over non-slack nodes (m rows), column indices ascending, diagonal
// present in every row. `slot` = position in that CSR = Jacobian block slot.
void analyze(int m, const std::vector<int64_t>& rowptr, const std::vector<int>& colidx, const SymbolicOptions& opt,
             Symbolic& out);

#ifndef __CUDACC__
constexpr int ING_NONE = INT32_MIN;    // (kernels.cuh defines the same pair for the device side)
constexpr int ING_MULTI = 1 << 30;
#endif

// shared-memory bytes a panel needs in the factor kernel
inline int panel_smem_bytes(int w, int h) { return ((w * (w + h) + h * w) * 4 + 6 * w + 10) * 8; }

}  // namespace sblx
