// CPU baseline for bench.py: a C++17 / OpenMP restatement of the reference's Navier2D::update().
//
// TEST INFRASTRUCTURE ONLY (like everything under oracle/): used by bench.py's cpu_baseline / --impl reference legs
// and as a second, fast checker in tests/.  The product (rustpde_mpi_b200) never loads it.
//
// The Rust reference cannot be built in this image (no cargo/rustc; funspace is not vendored), so this file keeps the
// reference's PASS STRUCTURE instead: every reference call (to_ortho, gradient, forward, backward,
// solve, each array expression of navier_eq.rs) is its own pass over freshly written arrays, lanes are processed in
// parallel the way rayon's par lanes do (src/field.rs:104-128 -> OpenMP parallel for), axis-0 lanes are gathered /
// scattered through contiguous scratch like ndarray's non-contiguous lanes, and the two dense products of
// Poisson::solve_par (src/solver/poisson.rs:213-219,231-235) go to the OpenBLAS DGEMM that ships inside the numpy
// wheel (the reference links OpenBLAS through ndarray-linalg).  Algorithms follow oracle/rustpde_oracle.py, which is
// pinned to the reference's golden vectors; tests/test_cpu_restated.py checks this file against that oracle.
//
// Transform sizes: Chebyshev n-1 and Fourier n must be powers of two (the benchmark configurations).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <omp.h>
#include <string>
#include <type_traits>
#include <vector>

typedef std::complex<double> cd;
static const double PI = 3.14159265358979323846264338327950288;

// ------------------------------------------------------------------------------------------------
// OpenBLAS from the numpy wheel (ILP64 build: 64-bit integers, symbols carry the scipy_ prefix / 64_ suffix)
// ------------------------------------------------------------------------------------------------
typedef void (*dgemm_fn)(int order, int ta, int tb, int64_t m, int64_t n, int64_t k, double alpha, const double* a, int64_t lda,
                         const double* b, int64_t ldb, double beta, double* c, int64_t ldc);
typedef void (*setthr_fn)(int);
static dgemm_fn g_dgemm = nullptr;
static setthr_fn g_setthr = nullptr;
static std::string g_err;

static void gemm_fallback(int64_t m, int64_t n, int64_t k, const double* a, const double* b, double* c) {
  // blocked OpenMP C = A B (row-major), used only when no OpenBLAS could be loaded
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < m; i++) {
    double* ci = c + i * n;
    for (int64_t j = 0; j < n; j++) ci[j] = 0.0;
    for (int64_t p = 0; p < k; p++) {
      const double aip = a[i * k + p];
      const double* bp = b + p * n;
      for (int64_t j = 0; j < n; j++) ci[j] += aip * bp[j];
    }
  }
}
// C (m x n) = A (m x k) B (k x n), all row-major
static void gemm(int64_t m, int64_t n, int64_t k, const double* a, const double* b, double* c) {
  if (g_dgemm) g_dgemm(101, 111, 111, m, n, k, 1.0, a, k, b, n, 0.0, c, n);
  else gemm_fallback(m, n, k, a, b, c);
}

// ------------------------------------------------------------------------------------------------
// complex FFT, power of two, Stockham autosort radix 4 / 2
// ------------------------------------------------------------------------------------------------
struct Fft {
  int n = 0;
  std::vector<cd> tw;
  void init(int n_) {
    n = n_; tw.resize(n);
    for (int k = 0; k < n; k++) tw[k] = cd(std::cos(2 * PI * k / n), -std::sin(2 * PI * k / n));
  }
  // forward transform of a (length n); b is scratch of the same length; the result ends in a
  void run(cd* a, cd* b) const {
    cd* x = a; cd* y = b;
    int Ns = 1;
    while (Ns < n) {
      const int rem = n / Ns;
      if (rem >= 4) {
        const int st = n / 4, tws = n / (Ns * 4);
        for (int j = 0; j < st; j++) {
          const int k = j & (Ns - 1);
          const cd w1 = tw[k * tws], w2 = tw[2 * k * tws], w3 = tw[3 * k * tws];
          const cd v0 = x[j], v1 = x[j + st] * w1, v2 = x[j + 2 * st] * w2, v3 = x[j + 3 * st] * w3;
          const cd t0 = v0 + v2, t1 = v0 - v2, t2 = v1 + v3, d = v1 - v3, t3 = cd(d.imag(), -d.real());   // * (-i)
          const int j0 = (j - k) * 4 + k;
          y[j0] = t0 + t2; y[j0 + Ns] = t1 + t3; y[j0 + 2 * Ns] = t0 - t2; y[j0 + 3 * Ns] = t1 - t3;
        }
        Ns *= 4;
      } else {
        const int st = n / 2, tws = n / (Ns * 2);
        for (int j = 0; j < st; j++) {
          const int k = j & (Ns - 1);
          const cd v0 = x[j], v1 = x[j + st] * tw[k * tws];
          const int j0 = (j - k) * 2 + k;
          y[j0] = v0 + v1; y[j0 + Ns] = v0 - v1;
        }
        Ns *= 2;
      }
      std::swap(x, y);
    }
    if (x != a) memcpy(a, x, sizeof(cd) * n);
  }
};

// ------------------------------------------------------------------------------------------------
// bases (funspace 0.3.0 semantics; oracle/rustpde_oracle.py:42-212)
// ------------------------------------------------------------------------------------------------
enum { CH = 0, CD = 1, CN = 2, R2C = 4 };
static bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

struct Base {
  int kind = 0, n = 0, m = 0, N = 0;
  bool cheb = false, comp = false;
  std::vector<double> s2;            // stencil: ortho_{k+2} += s2[k] c_k
  std::vector<double> tl, td, tu;    // Thomas factors of S^T S (offsets -2, 0, +2)
  Fft fft;                           // length N (Chebyshev: n-1) or n/2 (Fourier)
  std::vector<cd> wh;                // exp(-i pi k / Nf), k = 0..Nf
  int init(int kind_, int n_) {
    kind = kind_; n = n_;
    cheb = kind != R2C; comp = kind == CD || kind == CN;
    m = cheb ? (comp ? n - 2 : n) : n / 2 + 1;
    N = cheb ? n - 1 : n / 2;
    if (!is_pow2(N)) return 1;
    fft.init(N);
    wh.resize(N + 1);
    for (int k = 0; k <= N; k++) wh[k] = cd(std::cos(PI * k / N), -std::sin(PI * k / N));
    if (comp) {
      s2.resize(m);
      for (int k = 0; k < m; k++) s2[k] = kind == CD ? -1.0 : -((double)k / (k + 2.0)) * ((double)k / (k + 2.0));
      // S^T S: dia 1 + s2^2, off-diagonals (k, k+2) = s2[k]; LU without pivoting along each parity
      tl.assign(m, 0.0); td.assign(m, 0.0); tu.assign(m, 0.0);
      for (int k = 0; k < m; k++) { td[k] = 1.0 + s2[k] * s2[k]; if (k + 2 < m) tu[k] = s2[k]; }
      for (int k = 2; k < m; k++) { tl[k] = s2[k - 2] / td[k - 2]; td[k] -= tl[k] * tu[k - 2]; }
    }
    return 0;
  }
  int scratch_len() const { return 2 * N + 8; }   // in cd units

  // ---- lane kernels (contiguous lanes) ----
  // DCT-I of x[0..N] through one complex FFT of N points on the even extension (length 2N): out[k] = X_k
  void dct1(const double* x, double* out, cd* z, cd* zs) const {
    for (int j = 0; j < N; j++) {
      const int a = 2 * j, b = 2 * j + 1;
      z[j] = cd(a <= N ? x[a] : x[2 * N - a], b <= N ? x[b] : x[2 * N - b]);
    }
    fft.run(z, zs);
    for (int k = 0; k <= N; k++) {
      const cd zk = z[k == N ? 0 : k], zc = std::conj(z[k == 0 ? 0 : N - k]);
      const cd e = 0.5 * (zk + zc), o = cd(0.0, -0.5) * (zk - zc);
      out[k] = (e + wh[k] * o).real();
    }
  }
  void cheb_fwd(const double* v, double* c, cd* z, cd* zs) const {   // values -> Chebyshev coefficients (A.1)
    dct1(v, c, z, zs);
    const double f = 1.0 / N;
    for (int k = 0; k <= N; k++) c[k] *= (k & 1) ? -f : f;
    c[0] *= 0.5; c[N] *= 0.5;
  }
  void cheb_bwd(const double* c, double* v, double* tmp, cd* z, cd* zs) const {
    for (int k = 0; k <= N; k++) tmp[k] = (k & 1) ? -c[k] : c[k];
    tmp[0] *= 2.0; tmp[N] *= 2.0;
    dct1(tmp, v, z, zs);
    for (int k = 0; k <= N; k++) v[k] *= 0.5;
  }
  void rfft(const double* v, cd* out, cd* z, cd* zs) const {   // unnormalised r2c, n/2+1 modes (A.4)
    for (int j = 0; j < N; j++) z[j] = cd(v[2 * j], v[2 * j + 1]);
    fft.run(z, zs);
    for (int k = 0; k <= N; k++) {
      const cd zk = z[k == N ? 0 : k], zc = std::conj(z[k == 0 ? 0 : N - k]);
      out[k] = 0.5 * (zk + zc) + wh[k] * (cd(0.0, -0.5) * (zk - zc));
    }
  }
  void irfft(const cd* in, double* v, cd* z, cd* zs) const {   // c2r with 1/n; Im of modes 0 and n/2 ignored
    for (int k = 0; k < N; k++) {
      cd yk = in[k], yc = std::conj(in[N - k]);
      if (k == 0) { yk = cd(in[0].real(), 0.0); yc = cd(in[N].real(), 0.0); }
      const cd e = 0.5 * (yk + yc), o = 0.5 * (yk - yc) * std::conj(wh[k]);
      z[k] = std::conj(e + cd(0.0, 1.0) * o);   // inverse through the forward transform: conj in, conj out
    }
    fft.run(z, zs);
    const double f = 1.0 / N;
    for (int j = 0; j < N; j++) { v[2 * j] = z[j].real() * f; v[2 * j + 1] = -z[j].imag() * f; }
  }
  template <class T> void to_ortho(const T* c, T* o) const {   // composite -> orthonormal (A.2)
    if (!comp) { for (int k = 0; k < m; k++) o[k] = c[k]; return; }
    o[0] = c[0]; o[1] = c[1];
    for (int k = 2; k < m; k++) o[k] = c[k] + s2[k - 2] * c[k - 2];
    o[m] = s2[m - 2] * c[m - 2]; o[m + 1] = s2[m - 1] * c[m - 1];
  }
  template <class T> void from_ortho(const T* o, T* c) const {   // c = (S^T S)^-1 S^T o
    if (!comp) { for (int k = 0; k < m; k++) c[k] = o[k]; return; }
    for (int k = 0; k < m; k++) c[k] = o[k] + s2[k] * o[k + 2];
    for (int k = 2; k < m; k++) c[k] -= tl[k] * c[k - 2];
    c[m - 1] /= td[m - 1]; c[m - 2] /= td[m - 2];
    for (int k = m - 3; k >= 0; k--) c[k] = (c[k] - tu[k] * c[k + 2]) / td[k];
  }
  template <class T> void diff_cheb(T* o, T* tmp, int d) const {   // A.3, in place on n ortho coefficients
    for (int rep = 0; rep < d; rep++) {
      for (int k = n - 1; k >= 0; k--) {
        T b = (k + 1 < n) ? 2.0 * (k + 1) * o[k + 1] : T(0.0);
        if (k + 2 < n) b += tmp[k + 2];
        tmp[k] = b;
      }
      tmp[0] *= 0.5;
      for (int k = 0; k < n; k++) o[k] = tmp[k];
    }
  }
  int ortho_len() const { return cheb ? n : m; }
};

// ------------------------------------------------------------------------------------------------
// 2-D arrays (row-major) and the lane driver
// ------------------------------------------------------------------------------------------------
// Arrays come from a size-keyed pool: after the first step no pass allocates (the reference's own passes allocate through
// ndarray; a pool is the friendlier reading of "one allocation-free pass per reference call").
static std::multimap<size_t, void*> g_pool;
static void* pool_get(size_t bytes) {
  auto it = g_pool.find(bytes);
  if (it != g_pool.end()) { void* p = it->second; g_pool.erase(it); return p; }
  void* p = aligned_alloc(64, (bytes + 63) / 64 * 64);
  memset(p, 0, bytes);   // first touch
  return p;
}
static void pool_put(size_t bytes, void* p) { if (p) g_pool.emplace(bytes, p); }

template <class T> struct Arr {
  int r = 0, c = 0;
  T* p = nullptr;
  Arr() {}
  Arr(const Arr& o) { *this = o; }
  Arr(Arr&& o) noexcept : r(o.r), c(o.c), p(o.p) { o.p = nullptr; o.r = o.c = 0; }
  ~Arr() { pool_put(bytes(), p); }
  size_t size() const { return (size_t)r * c; }
  size_t bytes() const { return size() * sizeof(T); }
  T* data() { return p; }
  const T* data() const { return p; }
  void shape(int r_, int c_) {   // contents unspecified: every pass overwrites its whole output
    if (p && (size_t)r_ * c_ == size()) { r = r_; c = c_; return; }
    pool_put(bytes(), p);
    r = r_; c = c_; p = static_cast<T*>(pool_get(bytes()));
  }
  void zero() { memset(p, 0, bytes()); }
  Arr& operator=(const Arr& o) {
    if (this == &o) return *this;
    shape(o.r, o.c);
    const size_t n = size();
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)((n + 4095) / 4096); i++) {
      const size_t a = (size_t)i * 4096, b = std::min(n, a + 4096);
      memcpy(p + a, o.p + a, (b - a) * sizeof(T));
    }
    return *this;
  }
  T* row(int i) { return p + (size_t)i * c; }
  const T* row(int i) const { return p + (size_t)i * c; }
};

struct Scratch {   // per thread
  std::vector<cd> z, zs;
  std::vector<double> t0, t1, t2, t3;
  std::vector<cd> gin, gout;   // gathered axis-0 lanes (as cd: large enough for either type)
  void size(int L) {
    z.resize(2 * L + 16); zs.resize(2 * L + 16); t0.resize(L + 16); t1.resize(L + 16); t2.resize(L + 16); t3.resize(L + 16);
    gin.resize((size_t)8 * (L + 16)); gout.resize((size_t)8 * (L + 16));
  }
};
static std::vector<Scratch> g_scr;
static Scratch& scr() { return g_scr[omp_get_thread_num()]; }

// out = f applied to every lane of `in` along `axis`; lane lengths Lin -> Lout.  One pass over the array.
// Axis-0 lanes are gathered 8 columns at a time into contiguous scratch (ndarray copies non-contiguous lanes too).
template <class Ti, class To, class F>
static void along(int axis, const Arr<Ti>& in, Arr<To>& out, int Lin, int Lout, F f) {
  if (axis == 1) {
    out.shape(in.r, Lout);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < in.r; i++) f(in.row(i), out.row(i), scr());
  } else {
    out.shape(Lout, in.c);
    const int B = 8, nb = (in.c + B - 1) / B;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nb; b++) {
      Scratch& s = scr();
      const int c0 = b * B, w = std::min(B, in.c - c0);
      Ti* gi = reinterpret_cast<Ti*>(s.gin.data());
      To* go = reinterpret_cast<To*>(s.gout.data());
      for (int i = 0; i < Lin; i++) { const Ti* src = in.row(i) + c0; for (int k = 0; k < w; k++) gi[(size_t)k * Lin + i] = src[k]; }
      for (int k = 0; k < w; k++) f(gi + (size_t)k * Lin, go + (size_t)k * Lout, s);
      for (int i = 0; i < Lout; i++) { To* dst = out.row(i) + c0; for (int k = 0; k < w; k++) dst[k] = go[(size_t)k * Lout + i]; }
    }
  }
}

// real lane transforms applied to the real and imaginary parts of complex lanes
static void cheb_fwd_lane(const Base& b, const double* v, double* c, Scratch& s) { b.cheb_fwd(v, c, s.z.data(), s.zs.data()); }
static void cheb_fwd_lane(const Base& b, const cd* v, cd* c, Scratch& s) {
  double *re = s.t0.data(), *im = s.t1.data(), *o = s.t2.data();
  for (int i = 0; i < b.n; i++) { re[i] = v[i].real(); im[i] = v[i].imag(); }
  b.cheb_fwd(re, o, s.z.data(), s.zs.data());
  for (int i = 0; i < b.n; i++) c[i] = cd(o[i], 0.0);
  b.cheb_fwd(im, o, s.z.data(), s.zs.data());
  for (int i = 0; i < b.n; i++) c[i] = cd(c[i].real(), o[i]);
}
static void cheb_bwd_lane(const Base& b, const double* c, double* v, Scratch& s) { b.cheb_bwd(c, v, s.t3.data(), s.z.data(), s.zs.data()); }
static void cheb_bwd_lane(const Base& b, const cd* c, cd* v, Scratch& s) {
  double *re = s.t0.data(), *im = s.t1.data(), *o = s.t2.data();
  for (int i = 0; i < b.n; i++) { re[i] = c[i].real(); im[i] = c[i].imag(); }
  b.cheb_bwd(re, o, s.t3.data(), s.z.data(), s.zs.data());
  for (int i = 0; i < b.n; i++) v[i] = cd(o[i], 0.0);
  b.cheb_bwd(im, o, s.t3.data(), s.z.data(), s.zs.data());
  for (int i = 0; i < b.n; i++) v[i] = cd(v[i].real(), o[i]);
}

// ------------------------------------------------------------------------------------------------
// Space2 (src/field.rs:81-129).  T = double (confined) or cd (periodic: axis 0 is Fourier r2c)
// ------------------------------------------------------------------------------------------------
template <class T> struct Space {
  Base b0, b1;
  static constexpr bool periodic = !std::is_same<T, double>::value;
  // forward: axis 1 first on the real data, then axis 0 (oracle Space2.forward)
  void forward(const Arr<double>& v, Arr<T>& vhat) const {
    Arr<double> t;
    along(1, v, t, b1.n, b1.m, [&](const double* in, double* out, Scratch& s) {
      double* o = s.t2.data();
      b1.cheb_fwd(in, o, s.z.data(), s.zs.data());
      b1.from_ortho(o, out);
    });
    if constexpr (periodic) {
      along(0, t, vhat, b0.n, b0.m, [&](const double* in, cd* out, Scratch& s) { b0.rfft(in, out, s.z.data(), s.zs.data()); });
    } else {
      along(0, t, vhat, b0.n, b0.m, [&](const double* in, double* out, Scratch& s) {
        double* o = s.t2.data();
        b0.cheb_fwd(in, o, s.z.data(), s.zs.data());
        b0.from_ortho(o, out);
      });
    }
  }
  // backward: axis 0 first, then axis 1
  void backward(const Arr<T>& vhat, Arr<double>& v) const {
    if constexpr (periodic) {
      Arr<double> t;
      along(0, vhat, t, b0.m, b0.n, [&](const cd* in, double* out, Scratch& s) { b0.irfft(in, out, s.z.data(), s.zs.data()); });
      along(1, t, v, b1.m, b1.n, [&](const double* in, double* out, Scratch& s) {
        double* o = s.t0.data();
        b1.to_ortho(in, o);
        b1.cheb_bwd(o, out, s.t3.data(), s.z.data(), s.zs.data());
      });
    } else {
      Arr<double> t;
      along(0, vhat, t, b0.m, b0.n, [&](const double* in, double* out, Scratch& s) {
        double* o = s.t0.data();
        b0.to_ortho(in, o);
        b0.cheb_bwd(o, out, s.t3.data(), s.z.data(), s.zs.data());
      });
      along(1, t, v, b1.m, b1.n, [&](const double* in, double* out, Scratch& s) {
        double* o = s.t0.data();
        b1.to_ortho(in, o);
        b1.cheb_bwd(o, out, s.t3.data(), s.z.data(), s.zs.data());
      });
    }
  }
  void to_ortho(const Arr<T>& vhat, Arr<T>& o) const {
    Arr<T> t;
    along(0, vhat, t, b0.m, b0.ortho_len(), [&](const T* in, T* out, Scratch&) { b0.to_ortho(in, out); });
    along(1, t, o, b1.m, b1.ortho_len(), [&](const T* in, T* out, Scratch&) { b1.to_ortho(in, out); });
  }
  void from_ortho(const Arr<T>& o, Arr<T>& vhat) const {
    Arr<T> t;
    along(0, o, t, b0.ortho_len(), b0.m, [&](const T* in, T* out, Scratch&) { b0.from_ortho(in, out); });
    along(1, t, vhat, b1.ortho_len(), b1.m, [&](const T* in, T* out, Scratch&) { b1.from_ortho(in, out); });
  }
  // gradient (src/field.rs:127-129): to_ortho, differentiate per axis, divide by scale^deriv
  void gradient(const Arr<T>& vhat, int d0, int d1, const double* scale, Arr<T>& out) const {
    Arr<T> o, t;
    to_ortho(vhat, o);
    const int L0 = b0.ortho_len(), L1 = b1.ortho_len();
    along(0, o, t, L0, L0, [&](const T* in, T* out_, Scratch& s) {
      for (int k = 0; k < L0; k++) out_[k] = in[k];
      if (d0 == 0) return;
      if constexpr (periodic) {
        for (int k = 0; k < L0; k++) { cd f(1.0, 0.0); for (int r = 0; r < d0; r++) f *= cd(0.0, (double)k); out_[k] *= f; }
      } else {
        b0.diff_cheb(out_, reinterpret_cast<T*>(s.z.data()), d0);
      }
    });
    along(1, t, out, L1, L1, [&](const T* in, T* out_, Scratch& s) {
      for (int k = 0; k < L1; k++) out_[k] = in[k];
      if (d1) b1.diff_cheb(out_, reinterpret_cast<T*>(s.z.data()), d1);
    });
    if (scale) {
      const double f = 1.0 / (std::pow(scale[0], d0) * std::pow(scale[1], d1));
#pragma omp parallel for schedule(static)
      for (int i = 0; i < out.r; i++) { T* p = out.row(i); for (int j = 0; j < out.c; j++) p[j] *= f; }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// solvers (src/solver/*.rs; oracle/rustpde_oracle.py:364-661)
// ------------------------------------------------------------------------------------------------
struct Diags { int n = 0; std::vector<double> low, dia, up1, up2; void size(int n_) { n = n_; low.assign(n, 0); dia.assign(n, 0); up1.assign(n, 0); up2.assign(n, 0); } };
static void sweep(Diags& a) {   // src/solver/fdma.rs:73-82
  for (int i = 2; i < a.n; i++) {
    a.low[i - 2] /= a.dia[i - 2];
    a.dia[i] -= a.low[i - 2] * a.up1[i - 2];
    if (i < a.n - 2) a.up1[i] -= a.low[i - 2] * a.up2[i - 2];
  }
}
template <class T> static void fdma(const Diags& a, T* x) {   // src/solver/fdma.rs:101-118
  const int n = a.n;
  for (int i = 2; i < n; i++) x[i] -= x[i - 2] * a.low[i - 2];
  x[n - 1] /= a.dia[n - 1];
  x[n - 2] /= a.dia[n - 2];
  x[n - 3] = (x[n - 3] - x[n - 1] * a.up1[n - 3]) / a.dia[n - 3];
  x[n - 4] = (x[n - 4] - x[n - 2] * a.up1[n - 4]) / a.dia[n - 4];
  for (int i = n - 5; i >= 0; i--) x[i] = (x[i] - x[i + 2] * a.up1[i] - x[i + 4] * a.up2[i]) / a.dia[i];
}
// B2 entries: pv(i, off) = (laplace_inv_eye . laplace_inv)[i, i + off] (src/field.rs:195-216)
static double pv(int n, int i, int off) {
  const int r = i + 2;
  if (off == 0) return r == 2 ? 0.25 : 1.0 / (4.0 * r * (r - 1.0));
  if (off == 2) return (r < n - 2) ? -1.0 / (2.0 * ((double)r * r - 1.0)) : 0.0;
  if (off == 4) return (r < n - 4) ? 1.0 / (4.0 * r * (r + 1.0)) : 0.0;
  return 0.0;
}
static Diags mat_a(const Base& b) {   // pinv . S
  Diags a; a.size(b.m);
  for (int i = 0; i < b.m; i++) {
    if (i >= 2) a.low[i - 2] = pv(b.n, i, 0) * b.s2[i - 2];
    a.dia[i] = pv(b.n, i, 0) + pv(b.n, i, 2) * b.s2[i];
    if (i + 2 < b.m) a.up1[i] = pv(b.n, i, 2) + pv(b.n, i, 4) * b.s2[i + 2];
    if (i + 4 < b.m) a.up2[i] = pv(b.n, i, 4);
  }
  return a;
}
static Diags mat_b(const Base& b) {   // peye . S
  Diags a; a.size(b.m);
  for (int i = 0; i < b.m; i++) { a.dia[i] = b.s2[i]; if (i + 2 < b.m) a.up1[i] = 1.0; }
  return a;
}
// MatVecFdma with pinv (src/solver/matvec.rs:161-228): n ortho coefficients -> m
template <class T> static void matvec_pinv(const Base& b, const T* x, T* y) {
  const int m = b.m, n = b.n;
  for (int i = 0; i < m; i++) {
    T v = pv(n, i, 0) * x[i];
    v += pv(n, i, 2) * x[i + 2];
    y[i] = v;
  }
  for (int i = 0; i + 4 < n; i++) y[i] += pv(n, i, 4) * x[i + 4];
}

template <class T> struct Hholtz {   // src/solver/hholtz_adi.rs:36-169
  const Space<T>* sp = nullptr;
  Diags lu[2];
  std::vector<double> sd;   // Fourier axis 0: 1 / (1 + c k^2)
  void init(const Space<T>& s, double c0, double c1) {
    sp = &s;
    const double c[2] = {c0, c1};
    const Base* bs[2] = {&s.b0, &s.b1};
    for (int ax = 0; ax < 2; ax++) {
      const Base& b = *bs[ax];
      if (b.comp) {
        Diags a = mat_a(b), bm = mat_b(b);
        lu[ax].size(b.m);
        for (int i = 0; i < b.m; i++) {
          lu[ax].low[i] = a.low[i] - bm.low[i] * c[ax]; lu[ax].dia[i] = a.dia[i] - bm.dia[i] * c[ax];
          lu[ax].up1[i] = a.up1[i] - bm.up1[i] * c[ax]; lu[ax].up2[i] = a.up2[i] - bm.up2[i] * c[ax];
        }
        sweep(lu[ax]);
      } else {
        sd.resize(b.m);
        for (int k = 0; k < b.m; k++) sd[k] = 1.0 + c[ax] * (double)k * k;   // mass - laplace * c, src/solver/sdma.rs:37-46
      }
    }
  }
  void solve(const Arr<T>& in, Arr<T>& out) const {
    const Base &b0 = sp->b0, &b1 = sp->b1;
    Arr<T> r0, r1, s0;
    const Arr<T>* cur = &in;
    if (b0.comp) { along(0, *cur, r0, b0.n, b0.m, [&](const T* x, T* y, Scratch&) { matvec_pinv(b0, x, y); }); cur = &r0; }
    along(1, *cur, r1, b1.n, b1.m, [&](const T* x, T* y, Scratch&) { matvec_pinv(b1, x, y); });
    if (b0.comp) along(0, r1, s0, b0.m, b0.m, [&](const T* x, T* y, Scratch&) { for (int i = 0; i < b0.m; i++) y[i] = x[i]; fdma(lu[0], y); });
    else along(0, r1, s0, b0.m, b0.m, [&](const T* x, T* y, Scratch&) { for (int i = 0; i < b0.m; i++) y[i] = x[i] / sd[i]; });
    along(1, s0, out, b1.m, b1.m, [&](const T* x, T* y, Scratch&) { for (int i = 0; i < b1.m; i++) y[i] = x[i]; fdma(lu[1], y); });
  }
};

template <class T> struct Poisson {   // src/solver/poisson.rs:42-236 + fdma_tensor.rs:236-290
  const Space<T>* sp = nullptr;
  std::vector<double> lam, fwd, bwd;
  Diags lap1, mass1;   // unswept (lap = mat_b * c, mass = mat_a) of axis 1
  void init(const Space<T>& s, double c0, double c1, const double* lam_, const double* fwd_, const double* bwd_) {
    sp = &s;
    const Base& b0 = s.b0;
    if (b0.comp) {
      const size_t mm = (size_t)b0.m * b0.m;
      lam.assign(lam_, lam_ + b0.m); fwd.assign(fwd_, fwd_ + mm); bwd.assign(bwd_, bwd_ + mm);
    } else {
      lam.resize(b0.m);
      for (int k = 0; k < b0.m; k++) lam[k] = -(double)k * k * c0;
      if (std::fabs(lam[0]) < 1e-10) for (auto& v : lam) v -= 1e-10;   // poisson.rs:84-86
    }
    mass1 = mat_a(s.b1);
    lap1 = mat_b(s.b1);
    for (int i = 0; i < s.b1.m; i++) { lap1.low[i] *= c1; lap1.dia[i] *= c1; lap1.up1[i] *= c1; lap1.up2[i] *= c1; }
  }
  void solve(const Arr<T>& in, Arr<T>& out) const {
    const Base &b0 = sp->b0, &b1 = sp->b1;
    Arr<T> r0, r1, g, y;
    const Arr<T>* cur = &in;
    if (b0.comp) { along(0, *cur, r0, b0.n, b0.m, [&](const T* x, T* yv, Scratch&) { matvec_pinv(b0, x, yv); }); cur = &r0; }
    along(1, *cur, r1, b1.n, b1.m, [&](const T* x, T* yv, Scratch&) { matvec_pinv(b1, x, yv); });
    const Arr<T>* rows = &r1;
    if constexpr (std::is_same<T, double>::value) {
      if (b0.comp) { g.shape(b0.m, b1.m); gemm(b0.m, b1.m, b0.m, fwd.data(), r1.data(), g.data()); rows = &g; }
    }
    y.shape(rows->r, rows->c);
    const int m1 = b1.m;
#pragma omp parallel
    {
      Diags f; f.size(m1);   // FdmaTensor::solve re-sweeps one system per row (fdma_tensor.rs:266-287)
#pragma omp for schedule(static)
      for (int i = 0; i < rows->r; i++) {
        const double l = lam[i];
        for (int k = 0; k < m1; k++) {
          f.low[k] = lap1.low[k] + mass1.low[k] * l; f.dia[k] = lap1.dia[k] + mass1.dia[k] * l;
          f.up1[k] = lap1.up1[k] + mass1.up1[k] * l; f.up2[k] = lap1.up2[k] + mass1.up2[k] * l;
        }
        sweep(f);
        T* yr = y.row(i);
        const T* xr = rows->row(i);
        for (int k = 0; k < m1; k++) yr[k] = xr[k];
        fdma(f, yr);
      }
    }
    if constexpr (std::is_same<T, double>::value) {
      if (b0.comp) { out.shape(b0.m, b1.m); gemm(b0.m, b1.m, b0.m, bwd.data(), y.data(), out.data()); return; }
    }
    out = y;
  }
};

// ------------------------------------------------------------------------------------------------
// Navier2D (src/navier_stokes/navier.rs:49-466, navier_eq.rs, functions.rs)
// ------------------------------------------------------------------------------------------------
template <class T> struct Field {
  Space<T> sp;
  Arr<double> v;
  Arr<T> vhat;
  void init(int k0, int n0, int k1, int n1) { sp.b0.init(k0, n0); sp.b1.init(k1, n1); v.shape(n0, n1); v.zero(); vhat.shape(sp.b0.m, sp.b1.m); vhat.zero(); }
};

template <class T> static void axpy(Arr<T>& y, double a, const Arr<T>& x) {   // y += a x, one pass
#pragma omp parallel for schedule(static)
  for (int i = 0; i < y.r; i++) { T* p = y.row(i); const T* q = x.row(i); for (int j = 0; j < y.c; j++) p[j] += a * q[j]; }
}

struct NavBase {
  virtual ~NavBase() {}
  virtual void update() = 0;
  virtual void set_v(int which, const double* v) = 0;
  virtual void get_vhat(int which, double* out) = 0;
  virtual void set_vhat(int which, const double* in) = 0;
  virtual double div_norm() = 0;
  virtual void vhat_shape(int which, int* r, int* c, int* cx) = 0;
  virtual void time_ops(int calls, double* sec) = 0;
};

template <class T> struct Navier : NavBase {
  static constexpr bool periodic = !std::is_same<T, double>::value;
  int nx, ny;
  double ra, pr, dt, nu, ka, scale[2], time = 0;
  Field<T> temp, velx, vely, pres, pseu, tempbc, field;
  Hholtz<T> hh[3];
  Poisson<T> pois;
  Arr<T> tbc_ortho;

  Field<T>* fld(int which) { Field<T>* f[] = {&temp, &velx, &vely, &pres, &pseu, &tempbc}; return f[which]; }
  int init(int nx_, int ny_, double ra_, double pr_, double dt_, double aspect, const double* lam, const double* fwd, const double* bwd) {
    nx = nx_; ny = ny_; ra = ra_; pr = pr_; dt = dt_; scale[0] = aspect; scale[1] = 1.0;
    const double h = scale[1] * 2.0;
    nu = std::sqrt(pr / (ra / (h * h * h)));
    ka = std::sqrt(1.0 / ((ra / (h * h * h)) * pr));
    const int kx_v = periodic ? R2C : CD, kx_t = periodic ? R2C : CN, kx_o = periodic ? R2C : CH, kx_p = periodic ? R2C : CN;
    if (!is_pow2(periodic ? nx : nx - 1) || !is_pow2(ny - 1)) return 1;
    velx.init(kx_v, nx, CD, ny); vely.init(kx_v, nx, CD, ny); temp.init(kx_t, nx, CD, ny);
    pres.init(kx_o, nx, CH, ny); pseu.init(kx_p, nx, CN, ny); tempbc.init(kx_o, nx, CH, ny); field.init(kx_o, nx, CH, ny);
    const double sx2 = scale[0] * scale[0], sy2 = scale[1] * scale[1];
    hh[0].init(velx.sp, dt * nu / sx2, dt * nu / sy2);
    hh[1].init(vely.sp, dt * nu / sx2, dt * nu / sy2);
    hh[2].init(temp.sp, dt * ka / sx2, dt * ka / sy2);
    pois.init(pseu.sp, 1.0 / sx2, 1.0 / sy2, lam, fwd, bwd);
    // tempbc: boundary_conditions.rs:18-36 / :143-161
    std::vector<double> y(ny);
    for (int j = 0; j < ny; j++) y[j] = -std::cos(PI * j / (ny - 1));
    const double x1 = y[0], x2 = y[ny - 1], y1 = 0.5, y2 = -0.5;
    const double m = (y2 - y1) / (x2 - x1), n = (y1 * x2 - y2 * x1) / (x2 - x1);
    for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) tempbc.v.row(i)[j] = m * y[j] + n;
    tempbc.sp.forward(tempbc.v, tempbc.vhat);
    tempbc.sp.backward(tempbc.vhat, tempbc.v);
    return 0;
  }
  // functions.rs:56-69: u * backward(gradient), accumulated into conv
  void conv_term(Arr<double>& conv, const Arr<double>& u, const Field<T>& f, int d0, int d1, bool first) {
    Arr<T> g; Arr<double> gv;
    f.sp.gradient(f.vhat, d0, d1, scale, g);
    field.sp.backward(g, gv);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < conv.r; i++) {
      double* c = conv.row(i); const double* a = u.row(i); const double* b = gv.row(i);
      if (first) for (int j = 0; j < conv.c; j++) c[j] = a[j] * b[j];
      else for (int j = 0; j < conv.c; j++) c[j] += a[j] * b[j];
    }
  }
  // navier_eq.rs:60-101 + forward + dealias (functions.rs:72-82)
  void conv(const Field<T>& f, const Arr<double>& ux, const Arr<double>& uy, bool with_bc, Arr<T>& out) {
    Arr<double> c; c.shape(nx, ny);
    conv_term(c, ux, f, 1, 0, true);
    conv_term(c, uy, f, 0, 1, false);
    if (with_bc) { conv_term(c, ux, tempbc, 1, 0, false); conv_term(c, uy, tempbc, 0, 1, false); }
    field.v = c;
    field.sp.forward(field.v, field.vhat);
    const int n_x = field.vhat.r * 2 / 3, n_y = field.vhat.c * 2 / 3;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < field.vhat.r; i++) {
      T* p = field.vhat.row(i);
      if (i >= n_x) for (int j = 0; j < field.vhat.c; j++) p[j] = T(0.0);
      else for (int j = n_y; j < field.vhat.c; j++) p[j] = T(0.0);
    }
    out = field.vhat;
  }
  void div(Arr<T>& d) {
    Arr<T> g;
    velx.sp.gradient(velx.vhat, 1, 0, scale, d);
    vely.sp.gradient(vely.vhat, 0, 1, scale, g);
    axpy(d, 1.0, g);
  }
  double div_norm() override {
    Arr<T> d; div(d);
    double s = 0;
    for (size_t i = 0; i < d.size(); i++) s += std::norm(d.p[i]);
    return std::sqrt(s);
  }
  void update() override {   // navier.rs:438-466
    Arr<T> that, t2, rhs, g, cv, dv, o;
    temp.sp.to_ortho(temp.vhat, that);
    tempbc.sp.to_ortho(tempbc.vhat, t2);
    axpy(that, 1.0, t2);
    velx.sp.backward(velx.vhat, velx.v);
    vely.sp.backward(vely.vhat, vely.v);
    Arr<double> ux = velx.v, uy = vely.v;
    // solve_velx (navier_eq.rs:176-187)
    velx.sp.to_ortho(velx.vhat, rhs);
    pres.sp.gradient(pres.vhat, 1, 0, scale, g); axpy(rhs, -dt, g);
    conv(velx, ux, uy, false, cv); axpy(rhs, -dt, cv);
    hh[0].solve(rhs, velx.vhat);
    // solve_vely (navier_eq.rs:190-203)
    vely.sp.to_ortho(vely.vhat, rhs);
    pres.sp.gradient(pres.vhat, 0, 1, scale, g); axpy(rhs, -dt, g);
    axpy(rhs, dt, that);
    conv(vely, ux, uy, false, cv); axpy(rhs, -dt, cv);
    hh[1].solve(rhs, vely.vhat);
    // projection (navier.rs:455-458)
    div(dv);
    pois.solve(dv, pseu.vhat);
    pseu.vhat.row(0)[0] = T(0.0);
    // correct_velocity (navier_eq.rs:117-125)
    pseu.sp.gradient(pseu.vhat, 1, 0, scale, g); velx.sp.from_ortho(g, o); axpy(velx.vhat, -1.0, o);
    pseu.sp.gradient(pseu.vhat, 0, 1, scale, g); vely.sp.from_ortho(g, o); axpy(vely.vhat, -1.0, o);
    // update_pres (navier_eq.rs:137-143)
    axpy(pres.vhat, -nu, dv);
    pseu.sp.to_ortho(pseu.vhat, o); axpy(pres.vhat, 1.0 / dt, o);
    // solve_temp (navier_eq.rs:209-224)
    temp.sp.to_ortho(temp.vhat, rhs);
    tempbc.sp.gradient(tempbc.vhat, 2, 0, scale, g); axpy(rhs, dt * ka, g);
    tempbc.sp.gradient(tempbc.vhat, 0, 2, scale, g); axpy(rhs, dt * ka, g);
    conv(temp, ux, uy, true, cv); axpy(rhs, -dt, cv);
    hh[2].solve(rhs, temp.vhat);
    time += dt;
  }
  void set_v(int which, const double* v) override {
    Field<T>* f = fld(which);
    memcpy(f->v.data(), v, f->v.bytes());
    f->sp.forward(f->v, f->vhat);
  }
  void vhat_shape(int which, int* r, int* c, int* cx) override { Field<T>* f = fld(which); *r = f->vhat.r; *c = f->vhat.c; *cx = periodic; }
  void get_vhat(int which, double* out) override { Field<T>* f = fld(which); memcpy(out, f->vhat.data(), f->vhat.bytes()); }
  void set_vhat(int which, const double* in) override { Field<T>* f = fld(which); memcpy(f->vhat.data(), in, f->vhat.bytes()); }
  // Seconds per call of the standalone operators on the temperature / pseudo-pressure spaces of this problem (the CPU side of
  // bench.py's ms / transform and ms / solve; the reference's harnesses are benches/benchmark_transform.rs, benchmark_solver.rs).
  // sec[8] = backward, forward, to_ortho, from_ortho, gradient (1,0), gradient (0,2), HholtzAdi solve, Poisson solve.  The state
  // of the run is not touched (copies of temp / pseu are used); one untimed call first.
  void time_ops(int calls, double* sec) override {
    Arr<T> vh = temp.vhat, o, g, sol, rhs_p, sol_p;
    Arr<double> v;
    temp.sp.to_ortho(vh, o);
    pseu.sp.to_ortho(pseu.vhat, rhs_p);
    auto timed = [&](auto fn) {
      fn();
      const double t0 = omp_get_wtime();
      for (int i = 0; i < calls; i++) fn();
      return (omp_get_wtime() - t0) / calls;
    };
    sec[0] = timed([&] { temp.sp.backward(vh, v); });
    sec[1] = timed([&] { temp.sp.forward(v, vh); });
    sec[2] = timed([&] { temp.sp.to_ortho(vh, o); });
    sec[3] = timed([&] { temp.sp.from_ortho(o, vh); });
    sec[4] = timed([&] { temp.sp.gradient(vh, 1, 0, nullptr, g); });
    sec[5] = timed([&] { temp.sp.gradient(vh, 0, 2, nullptr, g); });
    sec[6] = timed([&] { hh[2].solve(o, sol); });
    sec[7] = timed([&] { pois.solve(rhs_p, sol_p); });
  }
};

extern "C" {
const char* rc_last_error() { return g_err.c_str(); }
// blas_path: the OpenBLAS of the numpy wheel (may be null: blocked OpenMP fallback); nthreads <= 0: all cores
int rc_init(const char* blas_path, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_num_procs();
  omp_set_num_threads(nthreads);
  g_scr.assign(nthreads, Scratch());
  g_dgemm = nullptr; g_setthr = nullptr;
  if (blas_path && *blas_path) {
    void* h = dlopen(blas_path, RTLD_NOW | RTLD_LOCAL);
    if (h) {
      g_dgemm = reinterpret_cast<dgemm_fn>(dlsym(h, "scipy_cblas_dgemm64_"));
      g_setthr = reinterpret_cast<setthr_fn>(dlsym(h, "scipy_openblas_set_num_threads64_"));
      if (g_setthr) g_setthr(nthreads);
    }
  }
  return g_dgemm ? 1 : 0;
}
int rc_threads() { return (int)g_scr.size(); }
void* rc_navier_create(int nx, int ny, double ra, double pr, double dt, double aspect, int periodic, const double* lam,
                       const double* fwd, const double* bwd) {
  if (g_scr.empty()) rc_init(nullptr, 0);
  const int L = std::max(nx, ny) + 8;
  for (auto& s : g_scr) s.size(L);
  if (periodic) {
    auto* n = new Navier<cd>();
    if (n->init(nx, ny, ra, pr, dt, aspect, nullptr, nullptr, nullptr)) { delete n; g_err = "transform sizes must be powers of two"; return nullptr; }
    return static_cast<NavBase*>(n);
  }
  auto* n = new Navier<double>();
  if (n->init(nx, ny, ra, pr, dt, aspect, lam, fwd, bwd)) { delete n; g_err = "transform sizes must be powers of two"; return nullptr; }
  return static_cast<NavBase*>(n);
}
void rc_navier_destroy(void* h) { delete static_cast<NavBase*>(h); }
void rc_navier_update(void* h, int steps) { for (int i = 0; i < steps; i++) static_cast<NavBase*>(h)->update(); }
void rc_navier_set_v(void* h, int which, const double* v) { static_cast<NavBase*>(h)->set_v(which, v); }
void rc_navier_vhat_shape(void* h, int which, int* r, int* c, int* cx) { static_cast<NavBase*>(h)->vhat_shape(which, r, c, cx); }
void rc_navier_get_vhat(void* h, int which, double* out) { static_cast<NavBase*>(h)->get_vhat(which, out); }
void rc_navier_set_vhat(void* h, int which, const double* in) { static_cast<NavBase*>(h)->set_vhat(which, in); }
double rc_navier_div_norm(void* h) { return static_cast<NavBase*>(h)->div_norm(); }
void rc_navier_time_ops(void* h, int calls, double* sec) { static_cast<NavBase*>(h)->time_ops(calls < 1 ? 1 : calls, sec); }
}
