// host.cpp -- section B of include/parafem_b200.h: CPU-side helpers that
// restate the ParaFEM library routines a p121/p123 host driver calls around
// the device path (partition arithmetic, p12meshgen cubes, steering arrays,
// deck readers, gather-table construction).  No CUDA in this file.
//
// Reference files followed (all under /root/reference/parafem/src):
//   modules/mpi/gather_scatter.f90  calc_nels_pp :146-257, calc_neq_pp :263-343,
//                                   make_ggl :1387-1780
//   modules/shared/geometry.f90     geometry_8bxz :70-169, geometry_20bxz :175-286,
//                                   cube_bc20 :425-500, cube_bc8 :589-650, box_bc8 :652-696
//   modules/mpi/loading.f90         load :36-142, load_p121 :386-546
//   modules/shared/new_library.f90  rearrange :3059-3112, find_g3 :3130-3212,
//                                   rearrange_2 :3118-3124, find_g4 :3249-3271,
//                                   abaqus2sg :3515-3682
//   modules/mpi/input.f90           read_p121 :3234-3396, read_p123 :3620-3806,
//                                   read_g_coord_pp :288-445, read_g_num_pp :935-1080,
//                                   read_rest :2570-2632, read_loads :2350-2411
//   tools/preprocessing/p12meshgen/p12meshgen.f90 :118-364 (p121), :658-819 (p123)
#include "parafem_b200.h"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// even split used for both elements and equations: the first `rem` ranks get
// one extra item (gather_scatter.f90:217-238 and :319-339 are the same formula)
void even_split(int64_t n, int npes, int numpe, int64_t *cnt, int64_t *start) {
  if (npes <= 1) { *cnt = n; *start = 1; return; }
  int64_t lo = n / npes, rem = n - lo * npes;
  int64_t hi = rem == 0 ? lo : lo + 1;
  if (numpe <= rem || rem == 0) {
    *cnt = hi; *start = (int64_t)(numpe - 1) * hi + 1;
  } else {
    *cnt = lo; *start = rem * hi + (int64_t)(numpe - rem - 1) * (hi - 1) + 1;
  }
}

// value as it survives a Fortran Ew.d text field (d significant digits)
double through_text(double x, int digits) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.*E", digits - 1, x);
  return strtod(buf, nullptr);
}

inline bool divisible(int64_t a, int64_t m) { return a % m == 0; }

void hex8_element(int64_t iel, int nxe, int nze, double aa, double bb, double cc,
                  int32_t *num, double *coord /* (8,3) col-major */) {
  // geometry_8bxz: elements run x fastest, then z (downwards), then y planes
  int64_t plane = (int64_t)nxe * nze;
  int64_t iq = (iel - 1) / plane + 1;
  int64_t ipl = iel - (iq - 1) * plane;
  int64_t is = (ipl - 1) / nxe + 1;
  int64_t ip = ipl - (is - 1) * nxe;
  int64_t row = nxe + 1, layer = (int64_t)(nxe + 1) * (nze + 1);
  int64_t n1 = (iq - 1) * layer + is * row + ip;
  int64_t n[8];
  n[0] = n1; n[1] = n1 - row; n[2] = n[1] + 1; n[3] = n1 + 1;
  n[4] = n1 + layer; n[5] = n[4] - row; n[6] = n[5] + 1; n[7] = n[4] + 1;
  for (int m = 0; m < 8; ++m) num[m] = (int32_t)n[m];
  double x0 = (ip - 1) * aa, x1 = ip * aa;
  double y0 = (iq - 1) * bb, y1 = iq * bb;
  // geometry_8bxz writes -is*cc and -(is-1)*cc with INTEGER is: the shipped 8-node decks (p124_demo.d) show the top
  // face at +0.0, i.e. the integer is negated before the product
  double zb = (double)(-is) * cc, zt = (double)(-(is - 1)) * cc;
  const int xhi[8] = {0, 0, 1, 1, 0, 0, 1, 1};
  const int yhi[8] = {0, 0, 0, 0, 1, 1, 1, 1};
  const int ztop[8] = {0, 1, 1, 0, 0, 1, 1, 0};
  for (int m = 0; m < 8; ++m) {
    coord[0 * 8 + m] = xhi[m] ? x1 : x0;
    coord[1 * 8 + m] = yhi[m] ? y1 : y0;
    coord[2 * 8 + m] = ztop[m] ? zt : zb;
  }
}

void hex20_element(int64_t iel, int nxe, int nze, double aa, double bb, double cc,
                   int32_t *num, double *coord /* (20,3) col-major */) {
  // geometry_20bxz: a "full" plane of corner+mid-edge nodes followed by a
  // mid-plane holding only the y-mid-edge nodes
  int64_t plane = (int64_t)nxe * nze;
  int64_t iq = (iel - 1) / plane + 1;
  int64_t ipl = iel - (iq - 1) * plane;
  int64_t is = (ipl - 1) / nxe + 1;
  int64_t ip = ipl - (is - 1) * nxe;
  int64_t per_y = (int64_t)(2 * nxe + 1) * (nze + 1) + (int64_t)(2 * nze + 1) * (nxe + 1);
  int64_t f1 = per_y * (iq - 1), f2 = per_y * iq;
  int64_t band = 3 * nxe + 2, mid = (int64_t)(nxe + 1) * (nze + 1);
  int64_t n[21];
  n[1] = f1 + band * is + 2 * ip - 1;
  n[2] = f1 + band * is - nxe + ip - 1;
  n[3] = n[1] - band;
  n[4] = n[3] + 1; n[5] = n[4] + 1; n[6] = n[2] + 1; n[7] = n[1] + 2; n[8] = n[1] + 1;
  n[9] = f2 - mid + (nxe + 1) * is + ip;
  n[10] = n[9] - nxe - 1; n[11] = n[10] + 1; n[12] = n[9] + 1;
  n[13] = f2 + band * is + 2 * ip - 1;
  n[14] = f2 + band * is - nxe + ip - 1;
  n[15] = n[13] - band;
  n[16] = n[15] + 1; n[17] = n[16] + 1; n[18] = n[14] + 1; n[19] = n[13] + 2; n[20] = n[13] + 1;
  for (int m = 0; m < 20; ++m) num[m] = (int32_t)n[m + 1];
  double *X = coord, *Y = coord + 20, *Z = coord + 40;
  auto c = [](double *a, int k) -> double & { return a[k - 1]; };  // 1-based view
  double x0 = (ip - 1) * aa, x1 = ip * aa;
  for (int k : {1, 2, 3, 9, 10, 13, 14, 15}) c(X, k) = x0;
  for (int k : {5, 6, 7, 11, 12, 17, 18, 19}) c(X, k) = x1;
  c(X, 4) = .5 * (c(X, 3) + c(X, 5)); c(X, 8) = .5 * (c(X, 1) + c(X, 7));
  c(X, 16) = .5 * (c(X, 15) + c(X, 17)); c(X, 20) = .5 * (c(X, 13) + c(X, 19));
  double y0 = (iq - 1) * bb, y1 = iq * bb;
  for (int k = 1; k <= 8; ++k) c(Y, k) = y0;
  for (int k = 13; k <= 20; ++k) c(Y, k) = y1;
  c(Y, 9) = .5 * (c(Y, 1) + c(Y, 13)); c(Y, 10) = .5 * (c(Y, 3) + c(Y, 15));
  c(Y, 11) = .5 * (c(Y, 5) + c(Y, 17)); c(Y, 12) = .5 * (c(Y, 7) + c(Y, 19));
  double zb = -(double)is * cc, zt = -(double)(is - 1) * cc;
  for (int k : {1, 7, 8, 9, 12, 13, 19, 20}) c(Z, k) = zb;
  for (int k : {3, 4, 5, 10, 11, 15, 16, 17}) c(Z, k) = zt;
  c(Z, 2) = .5 * (c(Z, 1) + c(Z, 3)); c(Z, 6) = .5 * (c(Z, 5) + c(Z, 7));
  c(Z, 14) = .5 * (c(Z, 13) + c(Z, 15)); c(Z, 18) = .5 * (c(Z, 17) + c(Z, 19));
}

struct RestWriter {
  int32_t *rest; int64_t nr; int ncol; int64_t count = 0; bool overflow = false;
  void add(int64_t node, int a, int b, int c) {
    if (count >= nr) { overflow = true; ++count; return; }
    rest[count] = (int32_t)node;
    int v[3] = {a, b, c};
    for (int k = 1; k < ncol; ++k) rest[(int64_t)k * nr + count] = v[k - 1];
    ++count;
  }
};

// tokenizer for list-directed decks: whitespace/comma separated, quotes kept off
bool read_tokens(const std::string &path, std::vector<std::string> &tok) {
  FILE *f = fopen(path.c_str(), "r");
  if (!f) return false;
  std::string cur; int ch; bool inq = false;
  auto flush = [&]() { if (!cur.empty()) { tok.push_back(cur); cur.clear(); } };
  while ((ch = fgetc(f)) != EOF) {
    if (ch == '\'' || ch == '"') { inq = !inq; if (!inq) { tok.push_back(cur); cur.clear(); } continue; }
    if (!inq && (ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r' || ch == ',')) { flush(); continue; }
    cur.push_back((char)ch);
  }
  flush(); fclose(f);
  return true;
}

bool is_number(const std::string &s) {
  if (s.empty()) return false;
  char *end = nullptr; strtod(s.c_str(), &end);
  return end && *end == '\0';
}

}  // namespace

extern "C" {

void pf_calc_nels_pp(int64_t nels, int npes, int numpe, int64_t *nels_pp, int64_t *iel_start) {
  even_split(nels, npes, numpe, nels_pp, iel_start);
}
void pf_calc_neq_pp(int64_t neq, int npes, int numpe, int64_t *neq_pp, int64_t *ieq_start) {
  even_split(neq, npes, numpe, neq_pp, ieq_start);
}

// calc_nels_pp partitioner 2 = read_nels_pp (input.f90:3108-3196): "<npes> n_1 ... n_npes"
int pf_read_psize(const char *job, int npes, int numpe, int64_t *nels_pp, int64_t *iel_start) {
  if (!job || npes < 1 || numpe < 1 || numpe > npes) return 2;
  FILE *f = fopen((std::string(job) + ".psize").c_str(), "r");
  if (!f) return 5;
  long long p = 0;
  if (fscanf(f, "%lld", &p) != 1 || p != npes) { fclose(f); return 6; }   // "Number of partitions is different ..."
  int64_t start = 1, mine = -1;
  for (int r = 1; r <= npes; ++r) {
    long long c = 0;
    if (fscanf(f, "%lld", &c) != 1 || c < 0) { fclose(f); return 7; }
    if (r < numpe) start += c;
    if (r == numpe) mine = c;
  }
  fclose(f);
  *nels_pp = mine; *iel_start = start;
  return 0;
}

int pf_p121_sizes(int nxe, int nye, int nze, int nod, int64_t *nn, int64_t *nr, int64_t *loaded) {
  int64_t X = nxe, Y = nye, Z = nze, nle = nxe / 5;
  if (nod == 20) {
    *nr = ((2 * X + 1) * (Z + 1) + (X + 1) * Z) * 2 + ((2 * Y - 1) * Z + (Y - 1) * Z) * 2 +
          (2 * Y - 1) * (X + 1) + (Y - 1) * X;
    *nn = (((2 * X + 1) * (Z + 1)) + ((X + 1) * Z)) * (Y + 1) + (X + 1) * (Z + 1) * Y;
    *loaded = 3 * nle * nle + 4 * nle + 1;
  } else if (nod == 8) {
    // p12meshgen.f90:181 writes ((nxe+1)*(nze+1))*2 + ((nye-1)*(nze+1))*2 + (nxe-1)*(nze-1), which equals
    // the number of rows cube_bc8 emits only when nxe == nye; count what cube_bc8 emits instead
    *nr = ((X + 1) * (Z + 1)) * 2 + (Y - 1) * (2 * Z + X + 1);
    *nn = (X + 1) * (Y + 1) * (Z + 1);
    *loaded = (nle + 1) * (nle + 1);
  } else return 1;
  return 0;
}

int pf_p123_sizes(int nxe, int nye, int nze, int64_t *nn, int64_t *nr, int64_t *nres) {
  int64_t X = nxe, Y = nye, Z = nze;
  *nr = (X + 1) * (Y + 1) + (X + 1) * Z + Y * Z;
  *nn = (X + 1) * (Y + 1) * (Z + 1);
  *nres = X * (Z - 1) + 1;
  return 0;
}

int pf_cube_elements(int nxe, int nze, int nod, double aa, double bb, double cc,
                     int64_t iel_start, int64_t nels_pp, int round_mode,
                     int32_t *g_num_pp, double *g_coord_pp) {
  if (nod != 8 && nod != 20) return 1;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nels_pp; ++e) {
    int32_t *num = g_num_pp + e * nod;
    double *co = g_coord_pp + e * nod * 3;
    if (nod == 20) hex20_element(iel_start + e, nxe, nze, aa, bb, cc, num, co);
    else hex8_element(iel_start + e, nxe, nze, aa, bb, cc, num, co);
    if (round_mode == 1)
      for (int k = 0; k < nod * 3; ++k) co[k] = through_text(co[k], 6);
  }
  return 0;
}

int pf_cube_rest(int kind, int nxe, int nye, int nze, int nod, int64_t nr, int32_t *rest) {
  int64_t X = nxe, Z = nze;
  if (kind == 1) {  // box_bc8 (p123): one column of flags, all zero
    RestWriter w{rest, nr, 2};
    int64_t face = (X + 1) * (Z + 1);
    for (int64_t i = 0; i < nye; ++i)
      for (int64_t j = i * face + 1; j <= (i + 1) * face; ++j)
        if (divisible(j, X + 1) || j < i * face + X + 1) w.add(j, 0, 0, 0);
    for (int64_t j = nye * face + 1; j <= (nye + 1) * face; ++j) w.add(j, 0, 0, 0);
    return (w.overflow || w.count != nr) ? 2 : 0;
  }
  RestWriter w{rest, nr, 4};
  if (nod == 20) {
    int64_t face1 = 3 * X * Z + 2 * (X + Z) + 1, face2 = (X + 1) * (Z + 1);
    int64_t face = face1 + face2, l = Z * (X + 1), m = 3 * X + 2, n = 3 * X * Z + 2 * Z;
    for (int64_t i = 0; i <= nye; ++i) {
      bool endplane = (i == 0 || i == nye);
      for (int64_t j = i * face + 1; j <= i * face + face1; ++j) {
        int64_t k = j - i * face;
        bool side = k <= n && (divisible(k + m - 1, m) || divisible(k + X + 1, m) ||
                               divisible(k + X, m) || divisible(k, m));
        if (endplane) {
          if (side) w.add(j, 0, 0, 1);
          else if (k <= n) w.add(j, 1, 0, 1);
          else w.add(j, 0, 0, 0);
        } else {
          if (side) w.add(j, 0, 1, 1);
          else if (k > n) w.add(j, 0, 0, 0);
        }
      }
      if (i < nye)
        for (int64_t j = i * face + face1 + 1; j <= (i + 1) * face; ++j) {
          int64_t k = j - (face1 + i * face);
          if (k <= l && (divisible(k + X, X + 1) || divisible(k, X + 1))) w.add(j, 0, 1, 1);
          else if (k > l) w.add(j, 0, 0, 0);
        }
    }
  } else if (nod == 8) {
    int64_t face = (X + 1) * (Z + 1), m = 2 * X + 2, n = (X + 1) * Z;
    for (int64_t i = 0; i <= nye; ++i) {
      bool endplane = (i == 0 || i == nye);
      for (int64_t j = i * face + 1; j <= i * face + face; ++j) {
        int64_t k = j - i * face;
        bool side = k <= n && (divisible(k + m - 1, m) || divisible(k + X + 1, m) ||
                               divisible(k + X, m) || divisible(k, m));
        if (endplane) {
          if (side) w.add(j, 0, 0, 1);
          else if (k <= n) w.add(j, 1, 0, 1);
          else w.add(j, 0, 0, 0);
        } else {
          if (side) w.add(j, 0, 1, 1);
          else if (k > n) w.add(j, 0, 0, 0);
        }
      }
    }
  } else return 1;
  return (w.overflow || w.count != nr) ? 2 : 0;
}

int pf_p121_loads(int nxe, int nze, int nod, double aa, double bb, int round_mode,
                  int32_t *node, double *val) {
  int64_t X = nxe, Z = nze, nle = nxe / 5, c = 0;
  std::vector<double> v;
  if (nod == 20) {
    int64_t f1 = (2 * X + 1) * (Z + 1) + (X + 1) * Z, f2 = (X + 1) * (Z + 1);
    auto edge_row = [&](int64_t base, double w_end, double w_even, double w_odd) {
      for (int64_t i = 1; i <= 2 * nle + 1; ++i) {
        node[c++] = (int32_t)(base + i);
        v.push_back((i == 1 || i == 2 * nle + 1) ? w_end : (i % 2 == 0 ? w_even : w_odd));
      }
    };
    edge_row(0, -1., 4., -2.);
    for (int64_t j = 0; j < nle; ++j) {
      for (int64_t i = 1; i <= nle + 1; ++i) {
        node[c++] = (int32_t)(i + f1 + j * (f1 + f2));
        v.push_back((i == 1 || i == nle + 1) ? 4. : 8.);
      }
      if (j != nle - 1) edge_row(f1 + (j + 1) * f2 + j * f1, -2., 8., -4.);
      else edge_row(f1 + (j + 1) * f2 + j * f1, -1., 4., -2.);
    }
    for (auto &x : v) x = -x * aa * bb * (25. / 12.);
  } else if (nod == 8) {
    int64_t f1 = (X + 1) * (Z + 1);
    for (int64_t i = 1; i <= nle + 1; ++i) {
      node[c++] = (int32_t)i;
      v.push_back((i == 1 || i == nle + 1) ? -6.25 : -12.5);
    }
    for (int64_t j = 0; j < nle; ++j)
      for (int64_t i = 1; i <= nle + 1; ++i) {
        node[c++] = (int32_t)(i + (j + 1) * f1);
        bool end = (i == 1 || i == nle + 1);
        if (j != nle - 1) v.push_back(end ? -12.5 : -25.);
        else v.push_back(end ? -6.25 : -12.5);
      }
    for (auto &x : v) x = x * aa * bb;
  } else return 1;
  for (int64_t i = 0; i < c; ++i) {
    val[3 * i + 0] = 0.; val[3 * i + 1] = 0.;
    val[3 * i + 2] = round_mode == 1 ? through_text(v[i], 8) : v[i];
  }
  return 0;
}

int pf_form_nf(int64_t nn, int nodof, int64_t nr, const int32_t *rest, int32_t *nf, int64_t *neq) {
  // free = 1 everywhere, then the rest rows overwrite; number free dofs in node order
  for (int64_t i = 0; i < nn * nodof; ++i) nf[i] = 1;
  for (int64_t i = 0; i < nr; ++i) {
    int64_t node = rest[i];
    if (node < 1 || node > nn) return 2;
    for (int k = 0; k < nodof; ++k) nf[(node - 1) * nodof + k] = rest[(int64_t)(k + 1) * nr + i];
  }
  int64_t m = 0;
  for (int64_t i = 0; i < nn * nodof; ++i)
    if (nf[i] != 0) nf[i] = (int32_t)(++m);
  *neq = m;
  return 0;
}

// status 5: a node number outside 1..nn (what a deck's .d / .lds / .fix files hold is never trusted as an index)
int pf_find_g(int nod, int nodof, int64_t nels_pp, int64_t nn, const int32_t *g_num_pp, const int32_t *nf,
              int32_t *g_g_pp) {
  int ntot = nod * nodof;
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (int64_t e = 0; e < nels_pp; ++e)
    for (int m = 0; m < nod; ++m) {
      int64_t node = g_num_pp[e * nod + m];
      if (node < 1 || node > nn) { bad |= 1; continue; }
      for (int k = 0; k < nodof; ++k) g_g_pp[e * ntot + m * nodof + k] = nf[(node - 1) * nodof + k];
    }
  return bad ? 5 : 0;
}

int pf_load(int nodof, int64_t loaded, int64_t nn, const int32_t *node, const double *val, const int32_t *nf,
            int64_t ieq_start, int64_t neq_pp, double *r_pp) {
  for (int64_t i = 0; i < neq_pp; ++i) r_pp[i] = 0.;
  for (int64_t i = 0; i < loaded; ++i)
    if (node[i] < 1 || node[i] > nn) return 5;
  for (int64_t i = 0; i < loaded; ++i)
    for (int k = 0; k < nodof; ++k) {
      int64_t eq = nf[(int64_t)(node[i] - 1) * nodof + k];
      if (eq >= ieq_start && eq < ieq_start + neq_pp) r_pp[eq - ieq_start] = val[i * nodof + k];
    }
  return 0;
}

int pf_abaqus2sg(int nod, int64_t nels, int32_t *g_num) {
  // new position m (0-based) takes old position perm[m] (0-based)
  static const int p20[20] = {3, 11, 0, 8, 1, 9, 2, 10, 19, 16, 17, 18, 7, 15, 4, 12, 5, 13, 6, 14};
  static const int p8[8] = {0, 4, 5, 1, 3, 7, 6, 2};
  static const int p4[4] = {0, 2, 1, 3};      // tetrahedron, new_library.f90:3647-3650
  const int *p = nod == 20 ? p20 : nod == 8 ? p8 : nod == 4 ? p4 : nullptr;
  if (!p) return 1;
  for (int64_t e = 0; e < nels; ++e) {
    int32_t t[20];
    for (int m = 0; m < nod; ++m) t[m] = g_num[e * nod + m];
    for (int m = 0; m < nod; ++m) g_num[e * nod + m] = t[p[m]];
  }
  return 0;
}

int pf_read_dat(const char *job, int program, pf_deck_info *info) {
  std::vector<std::string> t;
  if (!read_tokens(std::string(job) + ".dat", t)) return 1;
  memset(info, 0, sizeof *info);
  info->program = program;
  size_t k = 0;
  // xx3-style decks carry a leading program tag ('gpu') before 'hexahedron'
  while (k < t.size() && !is_number(t[k])) ++k;
  std::vector<double> v;
  for (; k < t.size(); ++k) if (is_number(t[k])) v.push_back(strtod(t[k].c_str(), nullptr));
  if (program == 121) {
    if (v.size() < 12) return 2;
    info->meshgen = (int)v[0]; info->partitioner = (int)v[1];
    info->nels = (int64_t)v[2]; info->nn = (int64_t)v[3]; info->nr = (int64_t)v[4];
    info->nip = (int)v[5]; info->nod = (int)v[6]; info->loaded = (int64_t)v[7];
    info->e = v[8]; info->v = v[9]; info->tol = v[10]; info->limit = (int)v[11];
  } else if (program == 123) {
    if (v.size() < 15) return 2;
    info->meshgen = (int)v[0]; info->partitioner = (int)v[1];
    info->nels = (int64_t)v[2]; info->nn = (int64_t)v[3]; info->nr = (int64_t)v[4];
    info->nip = (int)v[5]; info->nod = (int)v[6]; info->loaded = (int64_t)v[7];
    info->fixed = (int64_t)v[8];
    info->kx = v[9]; info->ky = v[10]; info->kz = v[11]; info->tol = v[12];
    info->limit = (int)v[13]; info->nres = (int64_t)v[14];
  } else if (program == 124) {
    // read_p124 (input.f90:3997-4169): element, mesh, partition, np_types / nels nn nr nip nod loaded fixed /
    // val0 / dtim nstep npri theta / tol limit nres   (p12meshgen.f90:975-987)
    if (v.size() < 18) return 2;
    info->meshgen = (int)v[0]; info->partitioner = (int)v[1]; info->np_types = (int)v[2];
    info->nels = (int64_t)v[3]; info->nn = (int64_t)v[4]; info->nr = (int64_t)v[5];
    info->nip = (int)v[6]; info->nod = (int)v[7]; info->loaded = (int64_t)v[8]; info->fixed = (int64_t)v[9];
    info->val0 = v[10]; info->dtim = v[11]; info->nstep = (int)v[12]; info->npri = (int)v[13]; info->theta = v[14];
    info->tol = v[15]; info->limit = (int)v[16]; info->nres = (int64_t)v[17];
  } else if (program == 125) {
    // read_p125: element, mesh, partition / nels nn nr nip nod loaded fixed / kx ky kz / dtim nstep npri / nres val0
    if (v.size() < 17) return 2;
    info->meshgen = (int)v[0]; info->partitioner = (int)v[1];
    info->nels = (int64_t)v[2]; info->nn = (int64_t)v[3]; info->nr = (int64_t)v[4];
    info->nip = (int)v[5]; info->nod = (int)v[6]; info->loaded = (int64_t)v[7]; info->fixed = (int64_t)v[8];
    info->kx = v[9]; info->ky = v[10]; info->kz = v[11]; info->dtim = v[12]; info->nstep = (int)v[13];
    info->npri = (int)v[14]; info->nres = (int64_t)v[15]; info->val0 = v[16];
  } else if (program == 2) {
    // read_xx2 (input.f90:5391-5562): element, mesh, partition, np_types, nels nn nr nip nod loaded_nodes
    // fixed_freedoms, tol limit
    if (v.size() < 12) return 2;
    info->meshgen = (int)v[0]; info->partitioner = (int)v[1]; info->np_types = (int)v[2];
    info->nels = (int64_t)v[3]; info->nn = (int64_t)v[4]; info->nr = (int64_t)v[5];
    info->nip = (int)v[6]; info->nod = (int)v[7]; info->loaded = (int64_t)v[8]; info->fixed = (int64_t)v[9];
    info->tol = v[10]; info->limit = (int)v[11];
  } else return 3;
  // sizes a caller allocates from: finite, non-negative and inside the 32-bit node / element numbering
  for (double x : v) if (!(x == x) || x > 9.0e18 || x < -9.0e18) return 4;
  const int64_t lim = 2147483647LL;
  if (info->nels < 1 || info->nels > lim || info->nn < 1 || info->nn > lim || info->nr < 0 || info->nr > info->nn ||
      info->loaded < 0 || info->loaded > 3 * info->nn || info->fixed < 0 || info->fixed > 3 * info->nn ||
      info->nod < 1 || info->nod > 20 || info->nip < 1 || info->nip > 27 || info->limit < 0 || info->np_types < 0 ||
      info->nstep < 0)
    return 4;
  return 0;
}

int pf_read_d(const char *job, int64_t nn, int64_t nels, int nod, double *g_coord, int32_t *g_num) {
  return pf_read_d_mat(job, nn, nels, nod, g_coord, g_num, nullptr);
}

// <job>.mat: read_material / read_materialValue (input.f90:3067-3102, 2789-2824) expect
// "*MATERIAL nmats nvals" / a name line / nmats lines "id v_1 .. v_nvals"; the shipped xx2-tiny.mat
// predates the header: a line with the material count, then the nmats lines.  Both are accepted.
// prop(nprops,np_types).
int pf_read_mat(const char *job, int nprops, int np_types, double *prop) {
  FILE *f = fopen((std::string(job) + ".mat").c_str(), "r");
  if (!f) return 1;
  char line[1024];
  int got = 0, rc = 0;
  bool header = false, first = true;
  while (got < np_types && fgets(line, sizeof line, f)) {
    char *q = line;
    while (*q == ' ' || *q == '\t') ++q;
    if (*q == '\n' || *q == 0) continue;
    if (first) {
      first = false;
      if (*q == '*') {                       // keyword nmats nvals, then one line to skip
        char kw[64]; int nm = 0, nv = 0;
        if (sscanf(q, "%63s %d %d", kw, &nm, &nv) != 3 || nm < np_types || nv != nprops) { rc = 2; break; }
        header = true;
        if (!fgets(line, sizeof line, f)) { rc = 3; break; }
        continue;
      }
    }
    char *end = nullptr;
    strtol(q, &end, 10);                     // material number
    if (end == q) { rc = 4; break; }
    {                                        // a lone integer on the line = the material count of the old format
      char *r = end;
      while (*r == ' ' || *r == '\t' || *r == '\r') ++r;
      if (*r == '\n' || *r == 0) continue;
    }
    for (int k = 0; k < nprops; ++k) {
      char *e2 = nullptr;
      const double val = strtod(end, &e2);
      if (e2 == end) { rc = 5; break; }
      prop[(size_t)got * nprops + k] = val; end = e2;
    }
    if (rc) break;
    ++got;
  }
  (void)header;
  fclose(f);
  if (!rc && got != np_types) rc = 6;
  return rc;
}

// read_elements (input.f90:1434-1583): as pf_read_d, and the material number of every element
// (last column of the element lines) into etype (may be NULL)
// ---- fast path: the whole <job>.d in memory, the node and element sections cut into one chunk per thread at line
// boundaries and parsed with from_chars.  At 2 M+ elements the reference's rank-1 READ + MPI_SEND loop dominates its
// wall time (p123 book run: 117.8 s total, 52.4 s solve); this reads ~1 GB/s.  Records are taken
// in line order, as the reference does (READ(10,*) bitBucket,g_coord(:,j), input.f90:388-390; the leading numbers of
// an element line are read into a dummy, :1053).  Anything unexpected makes the caller fall back to the
// record-by-record reader below.
}  // extern "C"
namespace {

inline const char *skip_blank(const char *p, const char *e) {
  while (p < e && (*p == ' ' || *p == '\t' || *p == '\r' || *p == ',')) ++p;
  return p;
}
inline bool take_double(const char *&p, const char *e, double &v) {
  p = skip_blank(p, e);
  if (p < e && *p == '+') ++p;
  auto r = std::from_chars(p, e, v);
  if (r.ec != std::errc()) return false;
  if (r.ptr < e && (*r.ptr == 'D' || *r.ptr == 'd')) {     // Fortran double-precision exponent letter
    char tmp[64];
    const char *q = r.ptr + 1;
    while (q < e && (*q == '+' || *q == '-' || (*q >= '0' && *q <= '9'))) ++q;
    const size_t n = (size_t)(q - p);
    if (n >= sizeof tmp) return false;
    memcpy(tmp, p, n);
    tmp[r.ptr - p] = 'E';
    auto r2 = std::from_chars(tmp, tmp + n, v);
    if (r2.ec != std::errc() || r2.ptr != tmp + n) return false;
    p = q;
    return true;
  }
  p = r.ptr;
  return true;
}
inline bool take_int(const char *&p, const char *e, long long &v) {
  p = skip_blank(p, e);
  if (p < e && *p == '+') ++p;
  auto r = std::from_chars(p, e, v);
  if (r.ec != std::errc()) return false;
  p = r.ptr;
  return true;
}
inline bool blank_line(const char *p, const char *e) { return skip_blank(p, e) == e; }

// calls fn(line_index, begin, end) for every non-blank line of [a, b), in parallel; false if the count != expect
template <class F>
bool for_lines(const char *a, const char *b, int64_t expect, F fn) {
  int nt = 1;
#ifdef _OPENMP
  nt = omp_get_max_threads();
#endif
  nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, (b - a) / 65536 + 1));
  std::vector<const char *> cut((size_t)nt + 1);
  cut[0] = a; cut[(size_t)nt] = b;
  for (int t = 1; t < nt; ++t) {
    const char *p = a + (b - a) * t / nt;
    const char *nl = static_cast<const char *>(memchr(p, '\n', (size_t)(b - p)));
    cut[(size_t)t] = nl ? nl + 1 : b;
  }
  std::vector<int64_t> first((size_t)nt + 1, 0);
  bool ok = true;
#pragma omp parallel num_threads(nt)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    int64_t n = 0;
    for (const char *p = cut[(size_t)t]; p < cut[(size_t)t + 1];) {
      const char *nl = static_cast<const char *>(memchr(p, '\n', (size_t)(cut[(size_t)t + 1] - p)));
      const char *le = nl ? nl : cut[(size_t)t + 1];
      if (!blank_line(p, le)) ++n;
      p = le + 1;
    }
    first[(size_t)t + 1] = n;
#pragma omp barrier
#pragma omp single
    {
      for (int q = 0; q < nt; ++q) first[(size_t)q + 1] += first[(size_t)q];
      if (first[(size_t)nt] != expect) ok = false;
    }
    if (ok) {
      int64_t line = first[(size_t)t];
      bool good = true;
      for (const char *p = cut[(size_t)t]; p < cut[(size_t)t + 1] && good;) {
        const char *nl = static_cast<const char *>(memchr(p, '\n', (size_t)(cut[(size_t)t + 1] - p)));
        const char *le = nl ? nl : cut[(size_t)t + 1];
        if (!blank_line(p, le)) good = fn(line++, p, le);
        p = le + 1;
      }
      if (!good) {
#pragma omp atomic write
        ok = false;
      }
    }
  }
  return ok;
}

int read_d_fast(const std::string &path, int64_t nn, int64_t nels, int nod, double *g_coord, int32_t *g_num, int32_t *etype) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return 1;
  fseek(f, 0, SEEK_END);
  const long size = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<char> buf((size_t)std::max<long>(size, 0));
  const bool got = size > 0 && fread(buf.data(), 1, (size_t)size, f) == (size_t)size;
  fclose(f);
  if (!got) return -1;
  const char *b0 = buf.data(), *be = b0 + size;
  // "*THREE_DIMENSIONAL" / "*NODES" / nn records / "*ELEMENTS" / nels records
  const char *p = static_cast<const char *>(memchr(b0, '\n', (size_t)size));
  if (!p) return -1;
  p = static_cast<const char *>(memchr(p + 1, '\n', (size_t)(be - p - 1)));
  if (!p) return -1;
  const char *nodes = p + 1;
  const char *star = static_cast<const char *>(memchr(nodes, '*', (size_t)(be - nodes)));
  if (!star) return -1;
  const char *el = static_cast<const char *>(memchr(star, '\n', (size_t)(be - star)));
  if (!el) return -1;
  ++el;
  if (!for_lines(nodes, star, nn, [&](int64_t i, const char *q, const char *e) {
        long long id; double x, y, z;
        if (!take_int(q, e, id) || !take_double(q, e, x) || !take_double(q, e, y) || !take_double(q, e, z)) return false;
        g_coord[i * 3] = x; g_coord[i * 3 + 1] = y; g_coord[i * 3 + 2] = z;
        return true;
      })) return -1;
  if (!for_lines(el, be, nels, [&](int64_t i, const char *q, const char *e) {
        long long id, a, b, c, v;
        if (!take_int(q, e, id) || !take_int(q, e, a) || !take_int(q, e, b) || !take_int(q, e, c) || b != nod) return false;
        for (int m = 0; m < nod; ++m) {
          if (!take_int(q, e, v) || v < 1 || v > nn) return false;     // the record reader reports which
          g_num[i * nod + m] = (int32_t)v;
        }
        if (!take_int(q, e, v)) return false;
        if (etype) etype[i] = (int32_t)v;
        return true;
      })) return -1;
  return 0;
}

}  // namespace
extern "C" {

int pf_read_d_mat(const char *job, int64_t nn, int64_t nels, int nod, double *g_coord, int32_t *g_num, int32_t *etype) {
  {
    const int fast = read_d_fast(std::string(job) + ".d", nn, nels, nod, g_coord, g_num, etype);
    if (fast >= 0) return fast;            // 0 = parsed, 1 = no such file; < 0: let the record reader decide
  }
  FILE *f = fopen((std::string(job) + ".d").c_str(), "r");
  if (!f) return 1;
  char word[256];
  int rc = 0;
  if (fscanf(f, "%255s", word) != 1 || fscanf(f, "%255s", word) != 1) rc = 2;  // *THREE_DIMENSIONAL *NODES
  for (int64_t i = 0; i < nn && !rc; ++i) {
    long long id; double x, y, z;
    if (fscanf(f, "%lld %lf %lf %lf", &id, &x, &y, &z) != 4) { rc = 3; break; }
    g_coord[i * 3 + 0] = x; g_coord[i * 3 + 1] = y; g_coord[i * 3 + 2] = z;      // line order (bitBucket, input.f90:389)
  }
  if (!rc && fscanf(f, "%255s", word) != 1) rc = 4;  // *ELEMENTS
  for (int64_t e = 0; e < nels && !rc; ++e) {
    long long id, a, b, c, v;
    if (fscanf(f, "%lld %lld %lld %lld", &id, &a, &b, &c) != 4 || b != nod) { rc = 5; break; }
    for (int m = 0; m < nod; ++m) {
      if (fscanf(f, "%lld", &v) != 1) { rc = 6; break; }
      if (v < 1 || v > nn) { rc = 8; break; }                         // node number outside 1..nn
      g_num[e * nod + m] = (int32_t)v;
    }
    if (!rc && fscanf(f, "%lld", &v) != 1) rc = 7;  // material id
    if (!rc && etype) etype[e] = (int32_t)v;
  }
  fclose(f);
  return rc;
}

int pf_read_bnd(const char *job, int64_t nr, int nodof, int32_t *rest) {
  FILE *f = fopen((std::string(job) + ".bnd").c_str(), "r");
  if (!f) return 1;
  int rc = 0;
  for (int64_t i = 0; i < nr && !rc; ++i)
    for (int k = 0; k <= nodof; ++k) {
      long long v;
      if (fscanf(f, "%lld", &v) != 1) { rc = 2; break; }
      rest[(int64_t)k * nr + i] = (int32_t)v;
    }
  fclose(f);
  return rc;
}

// read_fixed (input.f90:2483-2564): <job>.fix holds `fixed` lines "node sense value"
int pf_read_fix(const char *job, int64_t fixed, int32_t *node, int32_t *sense, double *valf) {
  FILE *f = fopen((std::string(job) + ".fix").c_str(), "r");
  if (!f) return 1;
  int rc = 0;
  for (int64_t i = 0; i < fixed; ++i) {
    long long n, s; double v;
    if (fscanf(f, "%lld %lld %lf", &n, &s, &v) != 3) { rc = 2; break; }
    node[i] = (int32_t)n; sense[i] = (int32_t)s; valf[i] = v;
  }
  fclose(f);
  return rc;
}

int pf_read_lds(const char *job, int64_t loaded, int nodof, int32_t *node, double *val) {
  // read_loads (input.f90:2350-2411): READ(10,*) node(i),val(:,i) -- one list-directed READ per record, so
  // values beyond the first nodof of a line are ignored (the xx11 decks carry three per line for nodof = 1)
  FILE *f = fopen((std::string(job) + ".lds").c_str(), "r");
  if (!f) return 1;
  int rc = 0;
  char line[4096];
  for (int64_t i = 0; i < loaded && !rc; ++i) {
    char *q = nullptr;
    for (;;) {                                       // next non-blank record
      if (!fgets(line, sizeof line, f)) { rc = 2; break; }
      q = line;
      while (*q == ' ' || *q == '\t' || *q == '\r' || *q == '\n') ++q;
      if (*q) break;
    }
    if (rc) break;
    char *end = nullptr;
    const long long n = strtoll(q, &end, 10);
    if (end == q) { rc = 2; break; }
    node[i] = (int32_t)n;
    for (int k = 0; k < nodof; ++k) {
      q = end;
      val[i * nodof + k] = strtod(q, &end);
      if (end == q) {                                // the record ended early: list-directed input continues on the next one
        if (!fgets(line, sizeof line, f)) { rc = 3; break; }
        end = line; --k;
      }
    }
  }
  fclose(f);
  return rc;
}

int pf_coords_pp(int nod, int64_t nels_pp, int64_t nn, const int32_t *g_num_pp, const double *g_coord,
                 double *g_coord_pp) {
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (int64_t e = 0; e < nels_pp; ++e)
    for (int m = 0; m < nod; ++m) {
      int64_t node = g_num_pp[e * nod + m] - 1;
      if (node < 0 || node >= nn) { bad |= 1; continue; }
      for (int d = 0; d < 3; ++d) g_coord_pp[e * nod * 3 + d * nod + m] = g_coord[node * 3 + d];
    }
  return bad ? 5 : 0;
}

// Fortran Ew.d edit descriptor (0.ddddE+xx), as dismsh_ensi_p's '(e12.5)' (output.f90:3050,3100)
static void fortran_e(char *out, size_t cap, double x, int w, int d) {
  // d significant digits, correctly rounded (to_chars, same digits as printf's %.{d-1}E), written as 0.DDDDE+XX
  char body[80];
  int n = 0;
  if (std::signbit(x)) body[n++] = '-';                   // gfortran prints -0.0 signed
  body[n++] = '0'; body[n++] = '.';
  if (x == 0.0) {
    for (int k = 0; k < d; ++k) body[n++] = '0';
    memcpy(body + n, "E+00", 4); n += 4;
  } else {
    char t[64];
    auto r = std::to_chars(t, t + sizeof t, std::fabs(x), std::chars_format::scientific, d - 1);
    *r.ptr = 0;
    const char *e = strchr(t, 'e');
    for (const char *q = t; q < e; ++q) if (*q != '.') body[n++] = *q;
    int ex = atoi(e + 1) + 1;
    body[n++] = 'E'; body[n++] = ex < 0 ? '-' : '+';
    if (ex < 0) ex = -ex;
    if (ex >= 100) { body[n - 2] = body[n - 1]; --n; body[n++] = (char)('0' + ex / 100); ex %= 100; }   // E+100 -> +100
    body[n++] = (char)('0' + ex / 10); body[n++] = (char)('0' + ex % 10);
  }
  int pad = w - n;
  size_t o = 0;
  for (; pad > 0 && o + 1 < cap; --pad) out[o++] = ' ';
  for (int k = 0; k < n && o + 1 < cap; ++k) out[o++] = body[k];
  out[o] = 0;
}

}  // extern "C"
namespace {
// n records formatted by `line(i, std::string&)` in parallel blocks and written in order
template <class F>
void write_records(FILE *f, int64_t n, F line) {
  const int64_t block = 1 << 12;
  const int64_t nblocks = (n + block - 1) / block;
  int nt = 1;
#ifdef _OPENMP
  nt = omp_get_max_threads();
#endif
  for (int64_t b0 = 0; b0 < nblocks; b0 += nt) {
    const int64_t nb = std::min<int64_t>(nt, nblocks - b0);
    std::vector<std::string> out((size_t)nb);
#pragma omp parallel for schedule(static, 1)
    for (int64_t b = 0; b < nb; ++b) {
      std::string &o = out[(size_t)b];
      const int64_t i0 = (b0 + b) * block, i1 = std::min<int64_t>(n, i0 + block);
      o.reserve((size_t)(i1 - i0) * 96);
      for (int64_t i = i0; i < i1; ++i) line(i, o);
    }
    for (auto &o : out) fwrite(o.data(), 1, o.size(), f);
  }
}
}  // namespace
extern "C" {

void pf_calc_nodes_pp(int64_t nn, int npes, int numpe, int64_t *nodes_pp, int64_t *node_start) {
  even_split(nn, npes, numpe, nodes_pp, node_start);
}

// calc_npes_pp (gather_scatter.f90:349-394): the reference's rough bound on the number of ranks a rank exchanges with,
// used to dimension toget / toput BEFORE make_ggl knows the answer ("causes the most execution failures, particularly
// for pathological cases", :371-372).  This library never needs it -- pf_setup_mesh counts the neighbours exactly
// (all-gather of the wanted-equation counts) -- but a driver that keeps its own make_ggl call still asks for it.
int pf_calc_npes_pp(int npes) {
  if (npes < 1) return 0;
  if (npes <= 15) return npes;
  if (npes <= 32) return npes / 2;
  if (npes <= 256) return npes / 4;
  if (npes <= 1024) return npes / 7;
  return npes / 12;
}

int pf_nodal_values(int nodof, int64_t nn, const int32_t *nf, int64_t ieq_start, int64_t neq_pp,
                    const double *x_pp, int64_t node_start, int64_t nodes_pp, double *out) {
  if (node_start < 1 || node_start + nodes_pp - 1 > nn) return 1;
  for (int64_t j = 0; j < nodes_pp; ++j)
    for (int k = 0; k < nodof; ++k) {
      const int64_t eq = nf[(node_start - 1 + j) * nodof + k];
      out[j * nodof + k] = (eq >= ieq_start && eq < ieq_start + neq_pp) ? x_pp[eq - ieq_start] : 0.0;
    }
  return 0;
}

int pf_write_ensi(const char *path, int numvar, int64_t nn, const double *values, int decimals) {
  FILE *f = fopen(path, "w");
  if (!f) return 1;
  fprintf(f, "Alya Ensight Gold --- %s per-node variable file\n", numvar == 1 ? "Scalar" : "Vector");
  fprintf(f, "part\n%s\ncoordinates\n", numvar == 1 ? "    1" : "     1");
  for (int c = 0; c < numvar; ++c)
    write_records(f, nn, [&](int64_t j, std::string &o) {
      char buf[64];
      fortran_e(buf, sizeof buf, values[j * numvar + c], 12, decimals);
      o += buf; o += '\n';
    });
  fclose(f);
  return 0;
}

// p12meshgen's output side for p121 (p12meshgen.f90:244-323): writes <job>.d/.bnd/.lds/.dat in the
// formats the reference's readers expect.  g_num is in S&G order; 20-node bricks are written in
// Abaqus order with meshgen = 2, 8-node bricks as they are with meshgen = 1.
int pf_write_deck_p121(const char *job, int nod, int64_t nels, int64_t nn, int64_t nr, int nip,
                       int64_t loaded, double e, double v, double tol, int limit, const double *g_coord,
                       const int32_t *g_num, const int32_t *rest, const int32_t *node, const double *val) {
  if (nod != 8 && nod != 20) return 1;
  static const int to_abaqus20[20] = {3, 5, 7, 1, 15, 17, 19, 13, 4, 6, 8, 2, 16, 18, 20, 14, 10, 11, 12, 9};
  char a[64], b[64], c[64];
  FILE *f = fopen((std::string(job) + ".d").c_str(), "w");
  if (!f) return 2;
  fprintf(f, "*THREE_DIMENSIONAL\n*NODES\n");
  write_records(f, nn, [&](int64_t i, std::string &o) {
    char x[64], y[64], z[64], ln[160];
    fortran_e(x, sizeof x, g_coord[i * 3], 14, 6); fortran_e(y, sizeof y, g_coord[i * 3 + 1], 14, 6);
    fortran_e(z, sizeof z, g_coord[i * 3 + 2], 14, 6);
    snprintf(ln, sizeof ln, "%12lld%s%s%s\n", (long long)(i + 1), x, y, z);
    o += ln;
  });
  fprintf(f, "*ELEMENTS\n");
  write_records(f, nels, [&](int64_t el, std::string &o) {
    char ln[32];
    snprintf(ln, sizeof ln, "%12lld %s", (long long)(el + 1), nod == 20 ? "3 20 1 " : "3 8 1 ");
    o += ln;
    for (int q = 0; q < nod; ++q) {
      const int m = nod == 20 ? to_abaqus20[q] - 1 : q;
      snprintf(ln, sizeof ln, "%12d", g_num[el * nod + m]);
      o += ln;
    }
    o += " 1\n";
  });
  fclose(f);
  f = fopen((std::string(job) + ".bnd").c_str(), "w");
  if (!f) return 2;
  for (int64_t i = 0; i < nr; ++i)
    fprintf(f, "%15d%6d%6d%6d\n", rest[i], rest[nr + i], rest[2 * nr + i], rest[3 * nr + i]);
  fclose(f);
  f = fopen((std::string(job) + ".lds").c_str(), "w");
  if (!f) return 2;
  for (int64_t i = 0; i < loaded; ++i) {
    fortran_e(a, sizeof a, val[i * 3], 16, 8); fortran_e(b, sizeof b, val[i * 3 + 1], 16, 8);
    fortran_e(c, sizeof c, val[i * 3 + 2], 16, 8);
    fprintf(f, "%12d%s%s%s\n", node[i], a, b, c);
  }
  fclose(f);
  f = fopen((std::string(job) + ".dat").c_str(), "w");
  if (!f) return 2;
  fortran_e(a, sizeof a, e, 12, 4); fortran_e(b, sizeof b, v, 12, 4); fortran_e(c, sizeof c, tol, 12, 4);
  fprintf(f, "'hexahedron'\n%s\n1\n%12lld%12lld%12lld%5d%5d%9lld\n%s%s%s%8d\n", nod == 8 ? "1" : "2",
          (long long)nels, (long long)nn, (long long)nr, nip, nod, (long long)loaded, a, b, c, limit);
  fclose(f);
  return 0;
}

// p12meshgen's output side for p123 / p124 / p125 (8-node bricks, one freedom per node)
int pf_write_deck_scalar(const char *job, const pf_deck_info *info, const double *g_coord, const int32_t *g_num,
                         const int32_t *rest) {
  if (!info || info->nod != 8 || (info->program != 123 && info->program != 124 && info->program != 125)) return 1;
  const int prog = info->program;
  const int64_t nn = info->nn, nels = info->nels, nr = info->nr;
  static const int to_abaqus8[8] = {1, 4, 8, 5, 2, 3, 7, 6};            // p12meshgen.f90:720-725, 887-891
  FILE *f = fopen((std::string(job) + ".d").c_str(), "w");
  if (!f) return 2;
  fprintf(f, "*THREE_DIMENSIONAL\n*NODES\n");
  write_records(f, nn, [&](int64_t i, std::string &o) {
    char x[64], y[64], z[64], ln[160];
    fortran_e(x, sizeof x, g_coord[i * 3], 14, 6); fortran_e(y, sizeof y, g_coord[i * 3 + 1], 14, 6);
    fortran_e(z, sizeof z, g_coord[i * 3 + 2], 14, 6);
    snprintf(ln, sizeof ln, "%12lld%s%s%s\n", (long long)(i + 1), x, y, z);
    o += ln;
  });
  fprintf(f, "*ELEMENTS\n");
  const int wnode = prog == 123 ? 10 : 12;                               // (I12,A,8I10,A) / (I12,A,8I12,A)
  write_records(f, nels, [&](int64_t el, std::string &o) {
    char ln[32];
    snprintf(ln, sizeof ln, "%12lld 3 8 1 ", (long long)(el + 1));
    o += ln;
    for (int q = 0; q < 8; ++q) {
      snprintf(ln, sizeof ln, "%*d", wnode, g_num[el * 8 + to_abaqus8[q] - 1]);
      o += ln;
    }
    o += " 1\n";
  });
  fclose(f);
  f = fopen((std::string(job) + ".bnd").c_str(), "w");
  if (!f) return 2;
  for (int64_t i = 0; i < nr; ++i) fprintf(f, "%*d%6d\n", prog == 123 ? 8 : 12, rest[i], rest[nr + i]);   // (I8,3I6) / (I12,3I6)
  fclose(f);
  char a[64], b[64], c[64], d[64], e[64];
  if (info->loaded > 0) {
    f = fopen((std::string(job) + ".lds").c_str(), "w");
    if (!f) return 2;
    fortran_e(a, sizeof a, 10.0, 16, 8);
    for (int64_t i = 0; i < info->loaded; ++i) fprintf(f, "%*lld%s\n", prog == 123 ? 10 : prog == 124 ? 12 : 11, (long long)info->nres, a);
    fclose(f);
  }
  if (info->fixed > 0) {
    f = fopen((std::string(job) + ".fix").c_str(), "w");
    if (!f) return 2;
    fortran_e(a, sizeof a, 100.0, 16, 8);
    for (int64_t i = 0; i < info->fixed; ++i) fprintf(f, "%*lld%s\n", prog == 123 ? 10 : 12, (long long)info->nres, a);
    fclose(f);
  }
  if (prog == 124) {
    f = fopen((std::string(job) + ".mat").c_str(), "w");
    if (!f) return 2;
    fortran_e(a, sizeof a, info->kx, 12, 4); fortran_e(b, sizeof b, info->ky, 12, 4); fortran_e(c, sizeof c, info->kz, 12, 4);
    fortran_e(d, sizeof d, info->rho, 12, 4); fortran_e(e, sizeof e, info->cp, 12, 4);
    fprintf(f, "*MATERIAL%5d%5d\n<edit material_name>\n 1%s%s%s%s%s\n", 1, 5, a, b, c, d, e);
    fclose(f);
  }
  f = fopen((std::string(job) + ".dat").c_str(), "w");
  if (!f) return 2;
  fprintf(f, "'hexahedron'\n2\n1\n");
  if (prog == 123) {
    fortran_e(a, sizeof a, info->kx, 12, 4); fortran_e(b, sizeof b, info->ky, 12, 4); fortran_e(c, sizeof c, info->kz, 12, 4);
    fortran_e(d, sizeof d, info->tol, 12, 4);
    fprintf(f, "%12lld%12lld%12lld%6d%6d%6lld%6lld\n%s%s%s%s%8d%8lld\n", (long long)nels, (long long)nn, (long long)nr, info->nip,
            info->nod, (long long)info->loaded, (long long)info->fixed, a, b, c, d, info->limit, (long long)info->nres);
  } else if (prog == 124) {
    fortran_e(a, sizeof a, info->val0, 12, 4); fortran_e(b, sizeof b, info->dtim, 12, 4); fortran_e(c, sizeof c, info->theta, 12, 4);
    fortran_e(d, sizeof d, info->tol, 12, 4);
    fprintf(f, "1\n%12lld%12lld%9lld%9d%9d%9lld%9lld\n%s\n%s%8d%8d%s\n%s%8d%8lld\n", (long long)nels, (long long)nn, (long long)nr,
            info->nip, info->nod, (long long)info->loaded, (long long)info->fixed, a, b, info->nstep, info->npri, c, d, info->limit,
            (long long)info->nres);
  } else {
    fortran_e(a, sizeof a, info->kx, 12, 4); fortran_e(b, sizeof b, info->ky, 12, 4); fortran_e(c, sizeof c, info->kz, 12, 4);
    fortran_e(d, sizeof d, info->dtim, 12, 4); fortran_e(e, sizeof e, info->val0, 12, 4);
    fprintf(f, "%9lld%9lld%9lld%9d%9d%9lld%9lld\n%s%s%s\n%s%8d%8d\n%8lld%s\n", (long long)nels, (long long)nn, (long long)nr, info->nip,
            info->nod, (long long)info->loaded, (long long)info->fixed, a, b, c, d, info->nstep, info->npri, (long long)info->nres, e);
  }
  fclose(f);
  return 0;
}

int pf_make_ggl(int ntot, int64_t nels_pp, const int32_t *g_g_pp, int64_t neq, int npes, int numpe,
                int32_t *ggl_pp, int64_t cap, int32_t *halo_eq, int64_t *halo_cnt, int64_t *nhalo) {
  int64_t neq_pp, ieq_start;
  even_split(neq, npes, numpe, &neq_pp, &ieq_start);
  int64_t lo = ieq_start, hi = ieq_start + neq_pp;  // owned: [lo, hi)
  int64_t total = nels_pp * ntot;
  // remote equations referenced by local elements, sorted unique; ascending
  // global number == ascending owner rank because ownership is contiguous
  std::vector<int32_t> remote;
  int bad = 0;
#pragma omp parallel
  {
    std::vector<int32_t> mine;
    int bad_t = 0;
#pragma omp for schedule(static) nowait
    for (int64_t i = 0; i < total; ++i) {
      int64_t g = g_g_pp[i];
      if (g < 0 || g > neq) bad_t = 1;
      else if (g != 0 && (g < lo || g >= hi)) mine.push_back((int32_t)g);
    }
    std::sort(mine.begin(), mine.end());
    mine.erase(std::unique(mine.begin(), mine.end()), mine.end());
#pragma omp critical
    {
      remote.insert(remote.end(), mine.begin(), mine.end());
      bad |= bad_t;
    }
  }
  if (bad) return 2;
  std::sort(remote.begin(), remote.end());
  remote.erase(std::unique(remote.begin(), remote.end()), remote.end());
  *nhalo = (int64_t)remote.size();
  if (halo_cnt) {
    for (int r = 0; r < npes; ++r) halo_cnt[r] = 0;
    int r = 1; int64_t cnt_r, st_r; even_split(neq, npes, r, &cnt_r, &st_r);
    for (int32_t g : remote) {
      while (g >= st_r + cnt_r) { ++r; even_split(neq, npes, r, &cnt_r, &st_r); }
      halo_cnt[r - 1]++;
    }
  }
  if (cap < (int64_t)remote.size() || !ggl_pp || !halo_eq) return cap == 0 ? 0 : 3;
  std::copy(remote.begin(), remote.end(), halo_eq);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < total; ++i) {
    int64_t g = g_g_pp[i];
    if (g == 0) ggl_pp[i] = 0;
    else if (g >= lo && g < hi) ggl_pp[i] = (int32_t)(g - lo + 1);
    else {
      auto it = std::lower_bound(remote.begin(), remote.end(), (int32_t)g);
      ggl_pp[i] = (int32_t)(neq_pp + 1 + (it - remote.begin()));
    }
  }
  return 0;
}


/* ---- binary decks: <job>.bin.ensi.geo, the EnSight Gold "C Binary" geometry file that
 * p12meshgenbin writes (mesh_ensi_geo_bin, input.f90:7986-8164) and read_g_coord_pp_be / read_g_num_pp_be
 * (input.f90:632-790, 1254-1420) read: nine 80-character records and two C ints of header, nn, the coordinates as
 * C floats component by component, the element type record, nels and the connectivity as C ints in EnSight's node
 * order.  Single-precision coordinates are what the format holds. ---- */
static const int kEnsi8[8] = {1, 4, 8, 5, 2, 3, 7, 6};                 /* input.f90:8099-8103 (S&G positions, 1-based) */
static const int kEnsi20[20] = {1, 7, 19, 13, 3, 5, 17, 15, 8, 12, 20, 9, 4, 11, 16, 10, 2, 6, 18, 14};   /* :8109-8119 */
static const int kEnsi4[4] = {1, 3, 2, 4};                             /* :8131-8132 */
static const int *ensi_order(int nod) { return nod == 8 ? kEnsi8 : nod == 20 ? kEnsi20 : nod == 4 ? kEnsi4 : nullptr; }

static void put80(FILE *f, const std::string &text) {
  char rec[80];
  memset(rec, ' ', sizeof rec);                                        /* Fortran CHARACTER(LEN=80): blank padded */
  memcpy(rec, text.data(), std::min<size_t>(text.size(), 80));
  fwrite(rec, 1, 80, f);
}

int pf_write_geo_bin(const char *job, int nod, int64_t nn, int64_t nels, const double *g_coord /*(3,nn)*/,
                     const int32_t *g_num_sg /*(nod,nels), S&G order*/) {
  const int *ord = ensi_order(nod);
  if (!ord || nn < 1 || nels < 1 || nn > 2147483647LL || nels > 2147483647LL) return 2;
  FILE *f = fopen((std::string(job) + ".bin.ensi.geo").c_str(), "wb");
  if (!f) return 1;
  std::string name(job);
  const size_t slash = name.find_last_of('/');
  if (slash != std::string::npos) name = name.substr(slash + 1);
  put80(f, "C Binary"); put80(f, "Problem name: " + name); put80(f, "Geometry files");
  put80(f, "node id off"); put80(f, "element id off"); put80(f, "part");
  const int32_t one = 1, nn32 = (int32_t)nn, nels32 = (int32_t)nels;
  fwrite(&one, 4, 1, f);
  put80(f, "Volume"); put80(f, "coordinates");
  fwrite(&nn32, 4, 1, f);
  std::vector<float> ordn((size_t)nn);
  for (int j = 0; j < 3; ++j) {
    for (int64_t i = 0; i < nn; ++i) ordn[(size_t)i] = (float)g_coord[i * 3 + j];
    fwrite(ordn.data(), 4, (size_t)nn, f);
  }
  put80(f, nod == 8 ? "hexa8" : nod == 20 ? "hexa20" : "tetra4");
  fwrite(&nels32, 4, 1, f);
  std::vector<int32_t> row((size_t)nod);
  for (int64_t e = 0; e < nels; ++e) {
    for (int m = 0; m < nod; ++m) row[(size_t)m] = g_num_sg[e * nod + ord[m] - 1];
    fwrite(row.data(), 4, (size_t)nod, f);
  }
  const bool ok = !ferror(f);
  fclose(f);
  return ok ? 0 : 3;
}

/* sizes in the header of a binary geometry file: nn, nels, nod (from the element type record) */
int pf_geo_bin_sizes(const char *job, int64_t *nn, int64_t *nels, int *nod) {
  FILE *f = fopen((std::string(job) + ".bin.ensi.geo").c_str(), "rb");
  if (!f) return 1;
  char rec[81]; rec[80] = 0;
  int32_t part = 0, nn32 = 0, nels32 = 0;
  int rc = 0;
  if (fread(rec, 1, 80, f) != 80 || strncmp(rec, "C Binary", 8) != 0) rc = 2;
  for (int k = 0; k < 5 && !rc; ++k) if (fread(rec, 1, 80, f) != 80) rc = 2;
  if (!rc && fread(&part, 4, 1, f) != 1) rc = 2;
  for (int k = 0; k < 2 && !rc; ++k) if (fread(rec, 1, 80, f) != 80) rc = 2;
  if (!rc && (fread(&nn32, 4, 1, f) != 1 || nn32 < 1)) rc = 2;
  if (!rc && fseek(f, (long)nn32 * 12, SEEK_CUR) != 0) rc = 2;
  if (!rc && fread(rec, 1, 80, f) != 80) rc = 2;
  if (!rc) {
    if (!strncmp(rec, "hexa20", 6)) *nod = 20;
    else if (!strncmp(rec, "hexa8", 5)) *nod = 8;
    else if (!strncmp(rec, "tetra4", 6)) *nod = 4;
    else rc = 3;
  }
  if (!rc && (fread(&nels32, 4, 1, f) != 1 || nels32 < 1)) rc = 2;
  fclose(f);
  if (!rc) { *nn = nn32; *nels = nels32; }
  return rc;
}

/* read_g_coord_pp_be + read_g_num_pp_be on global arrays: g_coord(3,nn) widened to double, g_num(nod,nels) in the
 * file's (EnSight) node order -- for 8-node bricks that is the order abaqus2sg expects (the drivers that use the
 * binary readers, xx12 / xx12_b, call it with meshgen = 2); pf_ensi2sg undoes it for every element type.  Node
 * numbers outside 1..nn are status 8. */
int pf_read_geo_bin(const char *job, int64_t nn, int64_t nels, int nod, double *g_coord, int32_t *g_num) {
  int64_t nn_f = 0, nels_f = 0; int nod_f = 0;
  int rc = pf_geo_bin_sizes(job, &nn_f, &nels_f, &nod_f);
  if (rc) return rc;
  if (nn_f != nn || nels_f != nels || nod_f != nod) return 4;
  FILE *f = fopen((std::string(job) + ".bin.ensi.geo").c_str(), "rb");
  if (!f) return 1;
  if (fseek(f, 6 * 80 + 4 + 2 * 80 + 4, SEEK_SET) != 0) { fclose(f); return 2; }
  std::vector<float> ordn((size_t)nn);
  for (int j = 0; j < 3 && !rc; ++j) {
    if (fread(ordn.data(), 4, (size_t)nn, f) != (size_t)nn) { rc = 2; break; }
    for (int64_t i = 0; i < nn; ++i) g_coord[i * 3 + j] = (double)ordn[(size_t)i];
  }
  if (!rc && fseek(f, 80 + 4, SEEK_CUR) != 0) rc = 2;
  if (!rc && fread(g_num, 4, (size_t)(nels * nod), f) != (size_t)(nels * nod)) rc = 2;
  fclose(f);
  if (!rc)
    for (int64_t i = 0; i < nels * nod; ++i) if (g_num[i] < 1 || g_num[i] > nn) return 8;
  return rc;
}

/* EnSight node order -> Smith & Griffiths order, in place (the inverse of mesh_ensi_geo_bin's permutation) */
int pf_ensi2sg(int nod, int64_t nels, int32_t *g_num) {
  const int *ord = ensi_order(nod);
  if (!ord) return 1;
  for (int64_t e = 0; e < nels; ++e) {
    int32_t t[20];
    for (int m = 0; m < nod; ++m) t[m] = g_num[e * nod + m];
    for (int m = 0; m < nod; ++m) g_num[e * nod + ord[m] - 1] = t[m];
  }
  return 0;
}

/* Tables of the exchanges fused into the PCG kernels of the peer transport (device.cu: k_pupdate stores every new p
 * value a peer's elements gather straight into that peer's halo segment; k_dot adds the received partial sums
 * chunk by chunk).  Pure index work on the halo tables of pf_setup_mesh, kept here so that it is testable on CPU. */
int pf_make_put_tables(int nranks, int64_t neq_pp, const int64_t *put_off /*nranks+1*/, const int32_t *put_slot,
                       const int64_t *fwd_dst_off /*nranks*/, uint32_t *bits /*(neq_pp+31)/32 + 1, zeroed here*/,
                       int32_t *slot0, uint32_t *ptr /*nput+1*/, int32_t *rank, int64_t *dst, int64_t *n_unique) {
  if (nranks < 1 || neq_pp < 0 || !put_off || !bits || !ptr || !n_unique) return 1;
  const int64_t nput = put_off[nranks];
  struct Ent { int32_t eq0, rank; int64_t dst; };
  std::vector<Ent> ents((size_t)nput);
  for (int r = 0; r < nranks; ++r)
    for (int64_t k = put_off[r]; k < put_off[r + 1]; ++k) {
      const int64_t slot = put_slot[k];
      if (slot < 1 || slot > neq_pp) return 2;
      ents[(size_t)k] = {(int32_t)(slot - 1), r, fwd_dst_off[r] + (k - put_off[r])};
    }
  std::stable_sort(ents.begin(), ents.end(), [](const Ent &a, const Ent &b) { return a.eq0 < b.eq0; });
  const int64_t nwords = (neq_pp + 31) / 32 + 1;
  for (int64_t w = 0; w < nwords; ++w) bits[w] = 0u;
  int64_t nu = 0;
  for (int64_t k = 0; k < nput; ++k) {
    if (k == 0 || ents[(size_t)k].eq0 != ents[(size_t)k - 1].eq0) { slot0[nu] = ents[(size_t)k].eq0; ptr[nu] = (uint32_t)k; ++nu; }
    bits[ents[(size_t)k].eq0 >> 5] |= 1u << (ents[(size_t)k].eq0 & 31);
    rank[k] = ents[(size_t)k].rank; dst[k] = ents[(size_t)k].dst;
  }
  ptr[nu] = (uint32_t)nput;
  *n_unique = nu;
  return 0;
}

/* accumulate entries per reduction chunk: entry k (owned slot acc_slot[k], ascending) belongs to chunk
 * (acc_slot[k]-1)/chunk; chunk_ptr[c] = first entry of chunk c, chunk_ptr[nchunks] = nacc */
int pf_make_acc_chunks(int64_t neq_pp, int chunk, int64_t nacc, const int32_t *acc_slot, uint32_t *chunk_ptr) {
  if (chunk < 1 || neq_pp < 0 || nacc < 0 || !chunk_ptr) return 1;
  const int64_t nchunks = (neq_pp + chunk - 1) / chunk;
  int64_t k = 0;
  for (int64_t c = 0; c < nchunks; ++c) {
    chunk_ptr[c] = (uint32_t)k;
    while (k < nacc && (int64_t)(acc_slot[k] - 1) / chunk == c) {
      if (k > 0 && acc_slot[k] <= acc_slot[k - 1]) return 2;
      ++k;
    }
  }
  chunk_ptr[nchunks] = (uint32_t)k;
  return k == nacc ? 0 : 2;
}

}  // extern "C"
