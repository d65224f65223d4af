// p12x_b200 -- C++ host driver for the scalar (one freedom per node, 8-node brick) programs on one B200:
//   p12x_b200 p123 n [round]    steady conduction          programs/5th_ed/p123/p123.f90
//   p12x_b200 p124 n [round]    implicit transient, theta  programs/5th_ed/p124/p124.f90
//   p12x_b200 p125 n [round]    explicit transient         programs/5th_ed/p125/p125.f90
// on the p12meshgen box of n^3 elements with the data of the shipped <program>_*.mg files (unit cube;
// round = 1: coordinates as they survive the deck's E14.6).  Same call sequence a Fortran driver makes
// through fortran/parafem_gpu.f90; prints the lines of <job>.res (single rank).
#include "parafem_b200.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
  if (argc < 3) { fprintf(stderr, "usage: %s p123|p124|p125 n [round]\n", argv[0]); return 2; }
  const std::string prog = argv[1];
  const int n = atoi(argv[2]), round_mode = argc > 3 ? atoi(argv[3]) : 0;
  if (n < 2 || (prog != "p123" && prog != "p124" && prog != "p125")) return 2;
  const double t_start = now();
  const int nod = 8, nodof = 1, nip = 8;
  const double aa = 1.0 / n;
  int64_t nn, nr, nres, neq = 0;
  pf_p123_sizes(n, n, n, &nn, &nr, &nres);
  const int64_t nels = (int64_t)n * n * n;
  std::vector<int32_t> g_num(nels * nod), rest(nr * 2, 0), nf(nn), g_g(nels * nod);
  std::vector<double> g_coord_pp(nels * nod * 3);
  pf_cube_elements(n, n, nod, aa, aa, aa, 1, nels, round_mode, g_num.data(), g_coord_pp.data());
  if (pf_cube_rest(1, n, n, n, nod, nr, rest.data())) return 2;
  if (pf_form_nf(nn, nodof, nr, rest.data(), nf.data(), &neq)) return 2;     // rearrange_2 + find_g4
  if (pf_find_g(nod, nodof, nels, nn, g_num.data(), nf.data(), g_g.data())) { fprintf(stderr, "connectivity names a node outside 1..nn\n"); return 2; }
  int64_t neq_pp, ieq_start;
  pf_calc_neq_pp(neq, 1, 1, &neq_pp, &ieq_start);
  const double t_read = now() - t_start;

  CHECK(pf_init(0, 1, 0, nullptr, &h));
  CHECK(pf_setup_mesh(h, nod, nodof, nip, nels, g_coord_pp.data(), g_g.data(), neq, ieq_start, neq_pp));
  std::vector<double> x(neq_pp);
  printf("This job ran on %5d  processes\n", 1);

  if (prog == "p123") {                                  // p123_*.mg: kx ky kz 2.0, tol 1e-5, source 10 at nres
    CHECK(pf_form_kc_laplace(h, 2.0, 2.0, 2.0));
    CHECK(pf_build_precon(h, 0, nullptr, 1e20));
    const double t_setup = now() - t_start;
    std::vector<double> r(neq_pp, 0.0);
    r[nres - 1] = 10.0;
    int iters = 0, conv = 0;
    const double t3 = now();
    CHECK(pf_pcg_solve(h, r.data(), 1e-5, 10000, x.data(), &iters, &conv));
    printf("There are %12lld nodes%12lld restrained and   %12lld equations\n", (long long)nn, (long long)nr, (long long)neq);
    printf("Time after setup is %10.4f\n", t_setup);
    printf("The number of iterations to convergence was %5d\n", iters);
    printf("The total load is %s\n", fe(10.0).c_str());
    printf("The potentials are:\n Freedom       Potential\n");
    for (int i = 0; i < 4; ++i) printf("%8lld     %s\n", (long long)(nres + i), fe(x[nres - 1 + i]).c_str());
    printf("Time spent in the solver was %10.4f\n", now() - t3);
  } else if (prog == "p124") {                           // p124_*.mg: k rho cp 1, dtim .01, 150 steps, theta .5,
    const double dtim = 0.01, theta = 0.5, tol = 1e-4, val0 = 100.0;   // npri 10, tol 1e-4, limit 100, val0 100
    const int nstep = 150, npri = 10, limit = 100;
    CHECK(pf_form_k_transient(h, 1.0, 1.0, 1.0, 1.0, 1.0, theta, dtim));
    CHECK(pf_build_precon(h, 0, nullptr, 1e20));
    CHECK(pf_transient_start(h, val0, nullptr));
    printf("There are %12lld nodes%12lld restrained and   %12lld equations\n", (long long)nn, (long long)nr, (long long)neq);
    printf("Time after setup is %10.4f\n", now() - t_start);
    printf("  Time       Temperature  Iterations \n%s%s\n", fe(0.0).c_str(), fe(val0).c_str());
    double solve_ms = 0.0;
    for (int j = 1; j <= nstep; ++j) {
      int iters = 0, conv = 0; double ms = 0.0;
      CHECK(pf_transient_step(h, nullptr, tol, limit, &iters, &conv, &ms));
      solve_ms += ms;
      if (j / npri * npri == j) {
        CHECK(pf_pcg_get_x(h, x.data()));
        printf("%s%s%10d\n", fe(j * dtim).c_str(), fe(x[nres - 1]).c_str(), iters);
      }
    }
    printf("The solution phase took %10.4f\n", solve_ms / 1e3);
  } else {                                               // p125_*.mg: k 1, dtim 2e-4, 5000 steps, npri 500, val0 100
    const double dtim = 2e-4, val0 = 100.0;
    const int nstep = 5000, npri = 500;
    CHECK(pf_form_k_explicit(h, 1.0, 1.0, 1.0, dtim));
    CHECK(pf_explicit_start(h, val0));
    printf("There are %12lld nodes%12lld restrained and%12lld equations\n", (long long)nn, (long long)nr, (long long)neq);
    printf("Time to read input is:%10.4f\n", t_read);
    printf("Time after setup is:%10.4f\n", now() - t_start);
    printf("  Time        Pressure\n%s%s\n", fe(0.0).c_str(), fe(val0).c_str());
    double step_ms = 0.0;
    for (int j = npri; j <= nstep; j += npri) {
      double ms = 0.0;
      CHECK(pf_explicit_steps(h, npri, &ms));
      step_ms += ms;
      CHECK(pf_pcg_get_x(h, x.data()));
      printf("%s%s\n", fe(j * dtim).c_str(), fe(x[nres - 1]).c_str());
    }
    printf("Time stepping recursion took  :%10.4f\n", step_ms / 1e3);
  }
  printf("This analysis took  :%10.4f\n", now() - t_start);
  pf_finalize(h);
  return 0;
}
