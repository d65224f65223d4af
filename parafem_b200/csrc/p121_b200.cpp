// p121_b200 -- C++ host driver for program p121 on one B200, written because this image has no
// Fortran compiler.  It makes the same sequence of calls a Fortran driver makes through
// fortran/parafem_gpu.f90 and prints the lines of <job>.res.
//
//   p121_b200 <job>               read the ParaFEM deck <job>.dat/.d/.bnd/.lds
//   p121_b200 --cube n nod        p12meshgen cube n^3 of nod-node bricks, generated in memory
//
// Flow mirrored: programs/5th_ed/p121/p121.f90:27-110 (single rank: numpe = npes = 1).
#include "parafem_b200.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Fortran Ew.d edit descriptor: 0.dddd E+xx (C's %E prints d.dddE+xx)
static std::string fe(double x, int w = 12, int d = 4) {
  char buf[64];
  if (x == 0.0) snprintf(buf, sizeof buf, "0.%0*dE+00", d, 0);
  else {
    snprintf(buf, sizeof buf, "%.*E", d - 1, x);          // d.ddd E+xx with d significant digits
    std::string m(buf);
    const size_t epos = m.find('E');
    int ex = atoi(m.c_str() + epos + 1) + 1;
    std::string digits;
    for (char c : m.substr(0, epos)) if (c >= '0' && c <= '9') digits.push_back(c);
    snprintf(buf, sizeof buf, "%s0.%sE%c%02d", x < 0 ? "-" : "", digits.c_str(), ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
  }
  std::string s(buf);
  if ((int)s.size() < w) s.insert(0, (size_t)w - s.size(), ' ');
  return s;
}

#define CHECK(call)                                                        \
  do {                                                                     \
    int st_ = (call);                                                      \
    if (st_ > 0) {                                                         \
      char buf[1024]; pf_last_error(h, buf, sizeof buf);                   \
      fprintf(stderr, "%s failed, status %d: %s\n", #call, st_, buf);     \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(int argc, char **argv) {
  pf_handle h = nullptr;
  if (argc < 2) { fprintf(stderr, "usage: %s <job> | --cube n nod\n", argv[0]); return 2; }
  const double t_start = now();
  const int nodof = 3;
  int nod = 20, nip = 8, limit = 2000;
  int64_t nels = 0, nn = 0, nr = 0, loaded = 0, neq = 0;
  double e = 100.0, v = 0.3, tol = 1e-5;
  std::vector<int32_t> g_num, rest, node, nf, g_g;
  std::vector<double> g_coord_pp, val;
  std::string res_path = "p121_b200.res";

  if (!strcmp(argv[1], "--cube")) {
    if (argc < 4) return 2;
    const int n = atoi(argv[2]); nod = atoi(argv[3]);
    const double aa = 10.0 / n;
    if (pf_p121_sizes(n, n, n, nod, &nn, &nr, &loaded)) return 2;
    nels = (int64_t)n * n * n;
    g_num.resize(nels * nod); g_coord_pp.resize(nels * nod * 3);
    pf_cube_elements(n, n, nod, aa, aa, aa, 1, nels, 0, g_num.data(), g_coord_pp.data());
    rest.assign(nr * 4, 0);
    if (pf_cube_rest(0, n, n, n, nod, nr, rest.data())) return 2;
    node.resize(loaded); val.resize(loaded * 3);
    pf_p121_loads(n, n, nod, aa, aa, 0, node.data(), val.data());
    limit = 20000;
  } else {
    const char *job = argv[1];
    res_path = std::string(job) + ".b200.res";
    pf_deck_info info;
    if (pf_read_dat(job, 121, &info)) { fprintf(stderr, "cannot read %s.dat\n", job); return 2; }
    nod = info.nod; nip = info.nip; limit = info.limit; nels = info.nels; nn = info.nn; nr = info.nr;
    loaded = info.loaded; e = info.e; v = info.v; tol = info.tol;
    std::vector<double> g_coord(nn * 3);
    g_num.resize(nels * nod);
    if (pf_read_d(job, nn, nels, nod, g_coord.data(), g_num.data())) { fprintf(stderr, "cannot read %s.d\n", job); return 2; }
    if (info.meshgen == 2) pf_abaqus2sg(nod, nels, g_num.data());
    g_coord_pp.resize(nels * nod * 3);
    if (pf_coords_pp(nod, nels, nn, g_num.data(), g_coord.data(), g_coord_pp.data())) { fprintf(stderr, "%s.d names a node outside 1..nn\n", job); return 2; }
    rest.assign(nr * 4, 0);
    if (pf_read_bnd(job, nr, nodof, rest.data())) { fprintf(stderr, "cannot read %s.bnd\n", job); return 2; }
    node.resize(loaded); val.resize(loaded * 3);
    if (pf_read_lds(job, loaded, nodof, node.data(), val.data())) { fprintf(stderr, "cannot read %s.lds\n", job); return 2; }
  }
  const double t_read = now() - t_start;

  // steering array and equations (p121.f90:44-48)
  const int ntot = nod * nodof;
  nf.resize(nn * nodof); g_g.resize(nels * ntot);
  if (pf_form_nf(nn, nodof, nr, rest.data(), nf.data(), &neq)) return 2;
  if (pf_find_g(nod, nodof, nels, nn, g_num.data(), nf.data(), g_g.data())) { fprintf(stderr, "connectivity names a node outside 1..nn\n"); return 2; }
  int64_t neq_pp, ieq_start;
  pf_calc_neq_pp(neq, 1, 1, &neq_pp, &ieq_start);

  // device: p121.f90:49-69,86
  CHECK(pf_init(0, 1, 0, nullptr, &h));
  CHECK(pf_setup_mesh(h, nod, nodof, nip, nels, g_coord_pp.data(), g_g.data(), neq, ieq_start, neq_pp));
  CHECK(pf_form_km_elastic(h, e, v));
  CHECK(pf_build_precon(h, 0, nullptr, 0.0));
  const double t_setup = now() - t_start;

  // starting r (p121.f90:79-85)
  std::vector<double> r(neq_pp), x(neq_pp);
  if (pf_load(nodof, loaded, nn, node.data(), val.data(), nf.data(), ieq_start, neq_pp, r.data())) { fprintf(stderr, "the load list names a node outside 1..nn\n"); return 2; }
  double q = 0.0;
  for (double t : r) q += t;

  // PCG (p121.f90:87-104)
  int iters = 0, converged = 0;
  const double t3 = now();
  CHECK(pf_pcg_solve(h, r.data(), tol, limit, x.data(), &iters, &converged));
  const double t_solve = now() - t3;
  double sigma[6];
  CHECK(pf_centroid_stress(h, 0, e, v, sigma));

  FILE *f = fopen(res_path.c_str(), "w");
  FILE *outs[2] = {stdout, f};
  for (FILE *o : outs) {
    if (!o) continue;
    fprintf(o, "This job ran on %7d processes\n", 1);
    fprintf(o, "There are %12lld nodes%12lld restrained and %12lld equations\n", (long long)nn, (long long)nr, (long long)neq);
    fprintf(o, "Time to read input is:%10.4f\n", t_read);
    fprintf(o, "Time after setup is:%10.4f\n", t_setup);
    fprintf(o, "The total load is:%s\n", fe(q).c_str());
    fprintf(o, "The number of iterations to convergence was %6d\n", iters);
    fprintf(o, "Time to solve equations was  :%10.4f\n", t_solve);
    fprintf(o, "The central nodal displacement is :%s\n", fe(x[0]).c_str());
    fprintf(o, "The Centroid point stresses for element 1 are\nPoint %5d\n", 1);
    for (double s : sigma) fprintf(o, "%s", fe(s).c_str());
    fprintf(o, "\nThis analysis took  :%10.4f\n", now() - t_start);
  }
  if (f) fclose(f);
  // displacements for ParaView (p121.f90:124-138)
  {
    int64_t nodes_pp, node_start;
    pf_calc_nodes_pp(nn, 1, 1, &nodes_pp, &node_start);
    std::vector<double> disp((size_t)nodes_pp * nodof);
    pf_nodal_values(nodof, nn, nf.data(), ieq_start, neq_pp, x.data(), node_start, nodes_pp, disp.data());
    const std::string ensi = res_path.substr(0, res_path.size() - 4) + ".ensi.DISPL-000001";
    pf_write_ensi(ensi.c_str(), nodof, nn, disp.data(), 5);
  }
  pf_finalize(h);
  return converged ? 0 : 3;
}
