// p1210_b200 -- C++ host driver for program p1210 (programs/5th_ed/p1210/p1210.f90: forced vibration of an
// elastic-plastic von Mises solid, lumped mass, explicit integration) on one B200:
//   p1210_b200 <job> [pload] [form] [nstep]
// reads <job>.dat / .d / .bnd / .lds as read_p1210 (input.f90:5107-5109) and the mesh / restraint / load readers do,
// also the 2010 layout of the one deck the reference ships (p1210_tiny.dat: no nres; `x dtim nstep npri` with the counts
// written as reals); `pload` overrides the deck's load multiplier (the shipped golden needs 2.0, tests/p1210_util.py),
// `form` 1 selects the tensor-core operator form, `nstep` shortens the run.  Writes <job>.b200.res with the lines of
// p1210.f90:55-60,113-114,152-153 and <job>.b200.dis in the layout of the shipped p1210_tiny.dis (*DISPLACEMENT / step /
// node x y z in 1PE12.4).  Same call sequence a Fortran driver makes through fortran/parafem_gpu.f90.
#include "parafem_b200.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

static double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// Fortran E12.4: 0.dddd E+xx
static std::string fe(double x) {
  char buf[64];
  if (x == 0.0) return "  0.0000E+00";
  int ex = (int)std::floor(std::log10(std::fabs(x))) + 1;
  double m = x / std::pow(10.0, ex);
  if (std::fabs(std::round(m * 1e4) / 1e4) >= 1.0) { m /= 10.0; ex += 1; }
  snprintf(buf, sizeof buf, "%s0.%04dE%c%02d", m < 0 ? "-" : "", (int)std::lround(std::fabs(m) * 1e4), ex < 0 ? '-' : '+', std::abs(ex));
  std::string s(buf);
  if (s.size() < 12) s.insert(0, 12 - s.size(), ' ');
  return s;
}
static double num(std::string t) {
  for (char &c : t) if (c == 'D' || c == 'd') c = 'E';
  return atof(t.c_str());
}
static bool is_int(const std::string &t) {
  size_t i = (t[0] == '+' || t[0] == '-') ? 1 : 0;
  if (i >= t.size()) return false;
  for (; i < t.size(); ++i) if (t[i] < '0' || t[i] > '9') return false;
  return true;
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
  if (argc < 2) { fprintf(stderr, "usage: %s <job> [pload] [form] [nstep]\n", argv[0]); return 2; }
  const double t_start = now();
  const std::string job = argv[1];
  const int nodof = 3;
  // ---- read_p1210 ----
  std::vector<std::string> tk;
  {
    std::ifstream in(job + ".dat");
    if (!in) { fprintf(stderr, "cannot read %s.dat\n", job.c_str()); return 2; }
    std::string t;
    while (in >> t) tk.push_back(t);
  }
  if (tk.size() < 17) { fprintf(stderr, "%s.dat: too few values for read_p1210\n", job.c_str()); return 2; }
  const int meshgen = atoi(tk[1].c_str());
  const int64_t nels = atoll(tk[3].c_str()), nn = atoll(tk[5].c_str()), nr = atoll(tk[6].c_str()), loaded = atoll(tk[8].c_str());
  const int nip = atoi(tk[4].c_str()), nod = atoi(tk[7].c_str());
  int64_t nres = 1;
  double rho, e, v, sbary, dtim, pload;
  int nstep, npri;
  if (tk.size() >= 18 && is_int(tk[9])) {          // the current layout: nres, rho e v sbary, dtim nstep npri pload
    nres = atoll(tk[9].c_str());
    rho = num(tk[10]); e = num(tk[11]); v = num(tk[12]); sbary = num(tk[13]); dtim = num(tk[14]);
    nstep = atoi(tk[15].c_str()); npri = atoi(tk[16].c_str()); pload = num(tk[17]);
  } else {                                         // the 2010 deck: rho e v sbary, x dtim nstep npri
    rho = num(tk[9]); e = num(tk[10]); v = num(tk[11]); sbary = num(tk[12]); pload = num(tk[13]); dtim = num(tk[14]);
    nstep = (int)std::lround(num(tk[15])); npri = (int)std::lround(num(tk[16]));
  }
  if (argc > 2) pload = atof(argv[2]);
  const int form = argc > 3 ? atoi(argv[3]) : 0;
  if (argc > 4) nstep = atoi(argv[4]);
  if (nels < 1 || nn < 1 || nr < 0 || nr > nn || nod != 20 || nip != 8 || loaded < 0 || loaded > nn || npri < 1 || nstep < 0) {
    fprintf(stderr, "%s.dat: sizes outside what p1210 takes (20-node hexahedra, nip = 8)\n", job.c_str());
    return 2;
  }
  // ---- mesh, restraints, loads (p1210.f90:33-41,106-111) ----
  std::vector<double> g_coord(nn * 3), g_coord_pp(nels * nod * 3), val(loaded * 3);
  std::vector<int32_t> g_num(nels * nod), rest(nr * 4, 0), node(loaded), nf(nn * nodof), g_g(nels * nod * nodof);
  if (pf_read_d(job.c_str(), nn, nels, nod, g_coord.data(), g_num.data())) { fprintf(stderr, "cannot read %s.d\n", job.c_str()); return 2; }
  if (meshgen == 2) pf_abaqus2sg(nod, nels, g_num.data());
  if (pf_coords_pp(nod, nels, nn, g_num.data(), g_coord.data(), g_coord_pp.data())) { fprintf(stderr, "%s.d names a node outside 1..nn\n", job.c_str()); return 2; }
  if (pf_read_bnd(job.c_str(), nr, nodof, rest.data())) { fprintf(stderr, "cannot read %s.bnd\n", job.c_str()); return 2; }
  if (loaded && pf_read_lds(job.c_str(), loaded, nodof, node.data(), val.data())) { fprintf(stderr, "cannot read %s.lds\n", job.c_str()); return 2; }
  int64_t neq = 0, neq_pp, ieq_start;
  if (pf_form_nf(nn, nodof, nr, rest.data(), nf.data(), &neq)) return 2;
  if (pf_find_g(nod, nodof, nels, nn, g_num.data(), nf.data(), g_g.data())) { fprintf(stderr, "connectivity names a node outside 1..nn\n"); return 2; }
  pf_calc_neq_pp(neq, 1, 1, &neq_pp, &ieq_start);
  std::vector<double> fext(neq_pp, 0.0), x1(neq_pp), d1(neq_pp), d2(neq_pp);
  if (loaded && pf_load(nodof, loaded, nn, node.data(), val.data(), nf.data(), ieq_start, neq_pp, fext.data())) { fprintf(stderr, "the load list names a node outside 1..nn\n"); return 2; }

  // ---- device: lumped mass, zero state, the explicit loop npri steps at a time ----
  CHECK(pf_init(0, 1, 0, nullptr, &h));
  CHECK(pf_setup_mesh(h, nod, nodof, nip, nels, g_coord_pp.data(), g_g.data(), neq, ieq_start, neq_pp));
  CHECK(pf_vm_explicit_begin(h, e, v, sbary, rho, dtim, pload, fext.data()));
  CHECK(pf_vm_explicit_set_form(h, form));
  FILE *res = fopen((job + ".b200.res").c_str(), "w"), *dis = fopen((job + ".b200.dis").c_str(), "w");
  if (!res || !dis) { fprintf(stderr, "cannot write beside %s\n", job.c_str()); return 2; }
  fprintf(res, "This job ran on %6d processes\n", 1);
  fprintf(res, "There are %12lld nodes %12lld restrained and %12lld equations\n", (long long)nn, (long long)nr, (long long)neq);
  fprintf(res, "Time after setup was:%10.4f\n", now() - t_start);
  fprintf(res, "  Time      Displacement  Velocity   Acceleration \n%s%s%s%s\n", fe(0.0).c_str(), fe(0.0).c_str(), fe(0.0).c_str(), fe(0.0).c_str());
  int64_t nodes_pp, node_start;
  pf_calc_nodes_pp(nn, 1, 1, &nodes_pp, &node_start);
  std::vector<double> disp((size_t)nodes_pp * nodof);
  double real_time = 0.0, step_ms = 0.0;
  for (int jj = npri; jj <= nstep; jj += npri) {
    double ms = 0.0;
    CHECK(pf_vm_explicit_steps(h, npri, &ms));
    step_ms += ms;
    for (int q = 0; q < npri; ++q) real_time = real_time + dtim;
    CHECK(pf_vm_explicit_get(h, x1.data(), d1.data(), d2.data(), nullptr));
    fprintf(res, "%s%s%s%s\n", fe(real_time).c_str(), fe(x1[nres - 1]).c_str(), fe(d1[nres - 1]).c_str(), fe(d2[nres - 1]).c_str());
    pf_nodal_values(nodof, nn, nf.data(), ieq_start, neq_pp, x1.data(), node_start, nodes_pp, disp.data());
    fprintf(dis, "*DISPLACEMENT                                     \n%13d\n", jj);
    for (int64_t i = 0; i < nn; ++i) fprintf(dis, "%8lld%12.4E%12.4E%12.4E\n", (long long)(i + 1), disp[3 * i], disp[3 * i + 1], disp[3 * i + 2]);
  }
  fprintf(res, "This analysis took:%10.4f\n", now() - t_start);
  fclose(res); fclose(dis);
  printf("p1210 %s: %lld equations, %d steps (form %d, load multiplier %g) in %.3f s on the device\n", job.c_str(), (long long)neq,
         nstep / npri * npri, form, pload, step_ms / 1e3);
  pf_finalize(h);
  return 0;
}
