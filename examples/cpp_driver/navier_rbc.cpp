// C++ driver over the C ABI of include/b200pde.h: the reference's examples/navier_rbc.rs
// (Navier2D::new_confined(129, 129, 1e5, 1, 0.01, 1, "rbc"), set_velocity / set_temperature, 100 steps) run through
// libb200pde.so without Python.  The host LAPACK setup of the confined Poisson solver (FdmaTensor::from_matrix,
// src/solver/fdma_tensor.rs:117-129: dgeev + inverses) is done here with LAPACKE from any LAPACK library given on the
// command line (dlopen; this image only has the OpenBLAS inside the numpy wheel, whose symbols carry a prefix / suffix).
//
//   g++ -O2 -std=c++17 -I include examples/cpp_driver/navier_rbc.cpp -o navier_rbc -L rustpde_mpi_b200 -lb200pde -ldl \
//       -Wl,-rpath,$PWD/rustpde_mpi_b200
//   ./navier_rbc <lapack .so> [nx ny steps periodic]
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <string>
#include <vector>
#include <dlfcn.h>
#include "b200pde.h"

#define CHECK(call)                                                                    \
  do {                                                                                 \
    int st_ = (call);                                                                  \
    if (st_ != B2_OK) { std::fprintf(stderr, "%s -> %d: %s\n", #call, st_, b2_last_error()); return 1; } \
  } while (0)

typedef int64_t lint;   // ILP64 LAPACKE (scipy_openblas64_); LP64 libraries: compile with -DLAPACK_LP64
#ifdef LAPACK_LP64
typedef int lapack_int_t;
static const char* SYM_GEEV = "LAPACKE_dgeev"; static const char* SYM_GESV = "LAPACKE_dgesv";
#else
typedef lint lapack_int_t;
static const char* SYM_GEEV = "scipy_LAPACKE_dgeev64_"; static const char* SYM_GESV = "scipy_LAPACKE_dgesv64_";
#endif
typedef lapack_int_t (*geev_fn)(int, char, char, lapack_int_t, double*, lapack_int_t, double*, double*, double*, lapack_int_t, double*, lapack_int_t);
typedef lapack_int_t (*gesv_fn)(int, lapack_int_t, lapack_int_t, double*, lapack_int_t, lapack_int_t*, double*, lapack_int_t);
static geev_fn p_geev; static gesv_fn p_gesv;
static const int ROW_MAJOR = 101;

// X = A^-1 B (row-major, n x n)
static std::vector<double> solve(std::vector<double> a, std::vector<double> b, int n) {
  std::vector<lapack_int_t> piv(n);
  if (p_gesv(ROW_MAJOR, n, n, a.data(), n, piv.data(), b.data(), n) != 0) { std::fprintf(stderr, "dgesv failed\n"); std::exit(2); }
  return b;
}

// (lam, fwd = Q^-1 C^-1, bwd = Q) of X = C^-1 A, eigenvalues sorted descending (src/solver/utils.rs:67-100), on the two parity
// blocks separately (A and C only couple indices of equal parity); singularity shift of src/solver/poisson.rs:84-86
static void poisson_eig(int kind, int n, double c, std::vector<double>& lam, std::vector<double>& fwd, std::vector<double>& bwd) {
  const int m = n - 2;
  std::vector<double> a0((size_t)m * m), cm((size_t)m * m);
  if (b2_host_poisson_matrices(kind, n, c, a0.data(), cm.data()) != B2_OK) { std::fprintf(stderr, "%s\n", b2_last_error()); std::exit(2); }
  std::vector<double> lam_u(m), q((size_t)m * m, 0.0), f((size_t)m * m, 0.0);
  for (int par = 0; par < 2; par++) {
    std::vector<int> idx;
    for (int i = par; i < m; i += 2) idx.push_back(i);
    const int k = (int)idx.size();
    std::vector<double> as((size_t)k * k), cs((size_t)k * k), eye((size_t)k * k, 0.0);
    for (int r = 0; r < k; r++) for (int s = 0; s < k; s++) { as[(size_t)r * k + s] = a0[(size_t)idx[r] * m + idx[s]]; cs[(size_t)r * k + s] = cm[(size_t)idx[r] * m + idx[s]]; }
    for (int r = 0; r < k; r++) eye[(size_t)r * k + r] = 1.0;
    std::vector<double> cinv = solve(cs, eye, k), x = solve(cs, as, k);
    std::vector<double> wr(k), wi(k), vr((size_t)k * k), vl(1);
    if (p_geev(ROW_MAJOR, 'N', 'V', k, x.data(), k, wr.data(), wi.data(), vl.data(), 1, vr.data(), k) != 0) { std::fprintf(stderr, "dgeev failed\n"); std::exit(2); }
    std::vector<int> ord(k);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int i, int j) { return wr[i] > wr[j]; });
    std::vector<double> qs((size_t)k * k);
    for (int r = 0; r < k; r++) for (int s = 0; s < k; s++) qs[(size_t)r * k + s] = vr[(size_t)r * k + ord[s]];
    std::vector<double> fp = solve(qs, cinv, k);   // Q^-1 C^-1
    for (int s = 0; s < k; s++) lam_u[idx[s]] = wr[ord[s]];
    for (int r = 0; r < k; r++) for (int s = 0; s < k; s++) { q[(size_t)idx[r] * m + idx[s]] = qs[(size_t)r * k + s]; f[(size_t)idx[r] * m + idx[s]] = fp[(size_t)r * k + s]; }
  }
  std::vector<int> perm(m);
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int i, int j) { return lam_u[i] > lam_u[j]; });
  lam.resize(m); fwd.assign((size_t)m * m, 0.0); bwd.assign((size_t)m * m, 0.0);
  for (int r = 0; r < m; r++) {
    lam[r] = lam_u[perm[r]];
    for (int s = 0; s < m; s++) { fwd[(size_t)r * m + s] = f[(size_t)perm[r] * m + s]; bwd[(size_t)s * m + r] = q[(size_t)s * m + perm[r]]; }
  }
  if (std::fabs(lam[0]) < 1e-10) for (auto& v : lam) v -= 1e-10;
}

int main(int argc, char** argv) {
  const char* lapack = argc > 1 ? argv[1] : "liblapacke.so";
  const int nx = argc > 2 ? std::atoi(argv[2]) : 129, ny = argc > 3 ? std::atoi(argv[3]) : 129;
  const int steps = argc > 4 ? std::atoi(argv[4]) : 100, periodic = argc > 5 ? std::atoi(argv[5]) : 0;
  const double ra = 1e5, pr = 1.0, dt = 0.01, aspect = 1.0, PI = 3.14159265358979323846;
  std::vector<double> lam, fwd, bwd;
  if (!periodic) {
    void* h = dlopen(lapack, RTLD_NOW | RTLD_LOCAL);
    if (!h) { std::fprintf(stderr, "dlopen %s: %s\n", lapack, dlerror()); return 2; }
    p_geev = reinterpret_cast<geev_fn>(dlsym(h, SYM_GEEV)); p_gesv = reinterpret_cast<gesv_fn>(dlsym(h, SYM_GESV));
    if (!p_geev || !p_gesv) { std::fprintf(stderr, "LAPACKE dgeev / dgesv not found in %s\n", lapack); return 2; }
    poisson_eig(B2_CHEB_NEUMANN, nx, 1.0 / (aspect * aspect), lam, fwd, bwd);
  }
  b2_ctx* ctx = nullptr; b2_navier* nav = nullptr;
  CHECK(b2_ctx_create(0, 0, 1, 0, &ctx));
  CHECK(b2_navier2d_create(ctx, nx, ny, ra, pr, dt, aspect, "rbc", periodic, periodic ? nullptr : lam.data(), periodic ? nullptr : fwd.data(),
                           periodic ? nullptr : bwd.data(), &nav));
  // examples/navier_rbc.rs:18-22: set_velocity(0.2, 1, 1), set_temperature(0.2, 1, 1) (functions.rs:85-125: unit coordinates)
  b2_field *temp, *velx, *vely;
  CHECK(b2_navier_field(nav, 0, &temp)); CHECK(b2_navier_field(nav, 1, &velx)); CHECK(b2_navier_field(nav, 2, &vely));
  std::vector<double> x(nx), y(ny), v((size_t)nx * ny);
  for (int i = 0; i < nx; i++) x[i] = periodic ? (double)i / (nx - 1) : (-std::cos(PI * i / (nx - 1)) + 1.0) / 2.0;   // (x - x0) / (x_last - x0)
  for (int j = 0; j < ny; j++) y[j] = (-std::cos(PI * j / (ny - 1)) + 1.0) / 2.0;
  const double amp = 0.2;
  for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) v[(size_t)i * ny + j] = amp * std::sin(PI * x[i]) * std::cos(PI * y[j]);
  CHECK(b2_field_set_v_host(velx, v.data(), v.size() * sizeof(double))); CHECK(b2_forward(velx));
  for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) v[(size_t)i * ny + j] = -amp * std::cos(PI * x[i]) * std::sin(PI * y[j]);
  CHECK(b2_field_set_v_host(vely, v.data(), v.size() * sizeof(double))); CHECK(b2_forward(vely));
  CHECK(b2_field_set_v_host(temp, v.data(), v.size() * sizeof(double))); CHECK(b2_forward(temp));
  // integrate (src/lib.rs:187-219): update until max_time, exit() on a NaN divergence
  double div = 0.0, t = 0.0;
  for (int s = 0; s < steps; s++) {
    CHECK(b2_navier_update(nav, 1));
    if ((s + 1) % 10 == 0 || s + 1 == steps) {
      CHECK(b2_navier_div_norm(nav, &div));
      if (std::isnan(div)) { std::fprintf(stderr, "divergence is NaN\n"); return 3; }
    }
  }
  CHECK(b2_navier_get_time(nav, &t));
  CHECK(b2_backward(temp));
  CHECK(b2_field_get_v_host(temp, v.data(), v.size() * sizeof(double)));
  double sum = 0.0, sq = 0.0;
  for (double a : v) { sum += a; sq += a * a; }
  std::printf("navier_rbc nx=%d ny=%d periodic=%d steps=%d time=%.6f div=%.12e temp_sum=%.12e temp_sumsq=%.12e\n", nx, ny, periodic, steps, t, div, sum, sq);
  CHECK(b2_navier_destroy(nav));
  CHECK(b2_ctx_destroy(ctx));
  return 0;
}
