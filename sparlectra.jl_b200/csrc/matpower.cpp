// MATPOWER `.m` case reader behind the C ABI: lets a non-Julia host feed the engine from real case files.
// Text rules of the reference's reader (src/MatpowerIO.jl): `%` comments stripped to the end of the line (:344-365);
// `mpc.baseMVA = <number>;` (:367-371); a matrix block is everything between `<key> = [` and the next `];` (:391-421);
// rows are separated by `;`, CR or LF, blank rows skipped (:427, :495-514); `mpc.bus` / `mpc.gen` / `mpc.branch` are
// padded / truncated with zeros to the MATPOWER v2 widths 13 / 21 / 13 (:278-280); bus rows are sorted by BUS_I when
// they are not already (`legacy_sort_bus`, :174-182); raw TAP values are kept (TAP == 0 = line, :184-189).
// Host only, no CUDA.  Only what the batched solver needs is read: no gencost, dcline, names or script post-processing.
#include "../../include/sblx.h"

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

struct sblx_mpc {
  double base_mva = 0.0;
  std::vector<double> bus, gen, branch;      // row-major, widths 13 / 21 / 13
  int64_t nbus = 0, ngen = 0, nbranch = 0;
};

namespace {

thread_local std::string g_mpc_error;

std::string strip_comments(const std::string& txt) {
  std::string out;
  out.reserve(txt.size());
  bool skip = false;
  for (char c : txt) {
    if (c == '\n') { skip = false; out.push_back(c); continue; }
    if (c == '%') skip = true;
    if (!skip) out.push_back(c);
  }
  return out;
}

double parse_base_mva(const std::string& txt) {
  const std::string key = "mpc.baseMVA";
  size_t k = txt.find(key);
  while (k != std::string::npos) {
    size_t p = k + key.size();
    while (p < txt.size() && isspace((unsigned char)txt[p])) ++p;
    if (p < txt.size() && txt[p] == '=') {
      ++p;
      while (p < txt.size() && isspace((unsigned char)txt[p])) ++p;
      size_t q = p;
      while (q < txt.size() && (isdigit((unsigned char)txt[q]) || txt[q] == '.')) ++q;
      size_t r = q;
      while (r < txt.size() && isspace((unsigned char)txt[r])) ++r;
      if (q > p && r < txt.size() && txt[r] == ';') return atof(txt.substr(p, q - p).c_str());
    }
    k = txt.find(key, k + 1);
  }
  throw std::runtime_error("Could not find `mpc.baseMVA = ...;` in MATPOWER file.");
}

std::vector<double> parse_matrix(const std::string& txt, const std::string& key, int ncols, int64_t& nrows) {
  const size_t k = txt.find(key);
  if (k == std::string::npos) throw std::runtime_error("Could not find matrix block `" + key + " = [ ... ];`");
  size_t p = k + key.size();
  while (p < txt.size() && isspace((unsigned char)txt[p])) ++p;
  if (p >= txt.size() || txt[p] != '=') throw std::runtime_error("Could not find `= [` after matrix block key `" + key + "`.");
  ++p;
  while (p < txt.size() && isspace((unsigned char)txt[p])) ++p;
  if (p >= txt.size() || txt[p] != '[') throw std::runtime_error("Could not find `= [` after matrix block key `" + key + "`.");
  ++p;
  const size_t end = txt.find("];", p);
  if (end == std::string::npos) throw std::runtime_error("Could not find closing `];` for matrix block `" + key + "`.");
  std::vector<double> out;
  nrows = 0;
  size_t i = p;
  while (i <= end) {
    size_t j = i;
    while (j < end && txt[j] != ';' && txt[j] != '\n' && txt[j] != '\r') ++j;
    std::istringstream row(txt.substr(i, j - i));
    std::vector<double> vals;
    std::string tok;
    while (row >> tok) {
      char* e = nullptr;
      const double v = strtod(tok.c_str(), &e);
      if (e == tok.c_str() || *e != '\0') throw std::runtime_error("Matrix `" + key + "`: cannot parse number `" + tok + "`.");
      if ((int)vals.size() < ncols) vals.push_back(v);
    }
    if (!vals.empty() || false) {
      vals.resize(ncols, 0.0);
      out.insert(out.end(), vals.begin(), vals.end());
      ++nrows;
    }
    if (j >= end) break;
    i = j + 1;
  }
  if (nrows == 0) throw std::runtime_error("Matrix `" + key + "` appears empty or could not be parsed.");
  return out;
}

}  // namespace

extern "C" {

const char* sblx_mpc_last_error(void) { return g_mpc_error.c_str(); }

int sblx_mpc_parse(const char* text, sblx_mpc** out) {
  if (!text || !out) return SBLX_E_ARG;
  *out = nullptr;
  try {
    const std::string txt = strip_comments(text);
    sblx_mpc* m = new sblx_mpc();
    try {
      m->base_mva = parse_base_mva(txt);
      m->bus = parse_matrix(txt, "mpc.bus", 13, m->nbus);
      m->gen = parse_matrix(txt, "mpc.gen", 21, m->ngen);
      m->branch = parse_matrix(txt, "mpc.branch", 13, m->nbranch);
      bool sorted = true;
      for (int64_t i = 1; i < m->nbus; ++i)
        if (m->bus[13 * i] < m->bus[13 * (i - 1)]) sorted = false;
      if (!sorted) {                         // legacy_sort_bus: stable sort of the rows by BUS_I
        std::vector<int64_t> idx(m->nbus);
        for (int64_t i = 0; i < m->nbus; ++i) idx[i] = i;
        std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return m->bus[13 * a] < m->bus[13 * b]; });
        std::vector<double> b2(m->bus.size());
        for (int64_t i = 0; i < m->nbus; ++i) std::copy(m->bus.begin() + 13 * idx[i], m->bus.begin() + 13 * (idx[i] + 1), b2.begin() + 13 * i);
        m->bus.swap(b2);
      }
    } catch (...) {
      delete m;
      throw;
    }
    *out = m;
    return SBLX_OK;
  } catch (const std::exception& e) {
    g_mpc_error = e.what();
    return SBLX_E_ARG;
  }
}

int sblx_mpc_read(const char* path, sblx_mpc** out) {
  if (!path || !out) return SBLX_E_ARG;
  std::ifstream f(path, std::ios::binary);
  if (!f) { g_mpc_error = std::string("cannot open ") + path; *out = nullptr; return SBLX_E_ARG; }
  std::stringstream ss;
  ss << f.rdbuf();
  return sblx_mpc_parse(ss.str().c_str(), out);
}

int sblx_mpc_dims(const sblx_mpc* m, double* base_mva, int64_t* nbus, int64_t* ngen, int64_t* nbranch) {
  if (!m) return SBLX_E_ARG;
  if (base_mva) *base_mva = m->base_mva;
  if (nbus) *nbus = m->nbus;
  if (ngen) *ngen = m->ngen;
  if (nbranch) *nbranch = m->nbranch;
  return SBLX_OK;
}

int sblx_mpc_tables(const sblx_mpc* m, double* bus, double* gen, double* branch) {
  if (!m) return SBLX_E_ARG;
  if (bus) memcpy(bus, m->bus.data(), m->bus.size() * sizeof(double));
  if (gen) memcpy(gen, m->gen.data(), m->gen.size() * sizeof(double));
  if (branch) memcpy(branch, m->branch.data(), m->branch.size() * sizeof(double));
  return SBLX_OK;
}

void sblx_mpc_free(sblx_mpc* m) { delete m; }

}  // extern "C"
