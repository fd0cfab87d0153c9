// Lane-program kernel: the one CUDA kernel family behind every per-axis operator
// of the Navier2D spectral hot path.
//
// A CTA owns one "lane group" = 4 neighbouring 1-D lanes (pencils) of a 2-D array
// and keeps them resident in shared memory while it interprets a short program of
// 1-D operators (load / banded mat-vec / Chebyshev recurrence / banded LU solve /
// DCT-I / real FFT / masks / store).  Arrays live in HBM in a 4x4 micro-tiled
// layout (128-byte tiles), so a lane group is ONE contiguous slab on the way in
// and full 128-byte lines on the way out, whether the store keeps the orientation
// or transposes it (that is how the x<->y pencil switch happens: every pass of a
// 2-D operator ends in a transposing store, locally or into a peer GPU's memory).
//
// sm_100a only.  No CPU fallback, no library calls in here.
#pragma once
#ifdef B2_EMU   // tests/emu: the same source compiled for the CPU SIMT emulator (test infrastructure only)
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define B2_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#define B2_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
#include <stdint.h>

#define B2_MAXOPS 24
#define B2_MAXPEERS 8

enum LaneOpCode {
  OP_LOAD = 1,     // W = [W +|*] a * src           i0=len  i2=flags(LD_*)       p0=src (p1 = stencil coefficients)
  OP_STORE = 2,    // dst = [dst +] a * W            i0=len  i2=flags(ST_*)       p0=dst  (p1=peer table)
  OP_BAND = 3,     // y_i = sum_m c_m[i] x_{i+o_m}   i0=len_out i1=packed offs i2=len_in  p0..p2 coef (null = 1, i1 byte=127: unused)
  OP_DERIV = 4,    // Chebyshev d/dx of i0 coeffs, i1 times, times a
  OP_FDMA = 5,     // banded LU solve (fwd elim + back subst)  i0=len i2=flags(FD_*) p0=fl p1=inv_dia p2=u1 p3=u2
  OP_DCT = 6,      // Chebyshev transform, i0=n (=N+1), i1: 0 fwd (values->coeffs) 1 bwd   p0=tw p1=tw2 p2=isin
  OP_RFFT = 7,     // Fourier r2c/c2r, i0=n, i1: 0 fwd 1 bwd                              p0=tw p1=tw2
  OP_FDIFF = 8,    // interleaved complex modes: c_k *= (i k)^{i1} * a,  i0 = number of modes
  OP_SCALEVEC = 9, // W[e] *= p0[e >> i1] for e < i0
  OP_ZEROTAIL = 10,// W[e] = 0 for e >= i0
  OP_LANEMASK = 11,// lanes >= i0 zeroed
  OP_ZEROELEM = 12,// W[lane i0][pos i1] = 0 (global lane index)
  OP_SCALE = 13,   // W *= a
};
enum { LD_ACC = 1, LD_PLAIN = 2, LD_MUL = 4, LD_STENCIL = 8 };  // LD_STENCIL: value = src[j] + p1[j] * src[j-2]
enum { ST_ACC = 1, ST_PLAIN = 2, ST_TRANS = 8, ST_PEER = 16 };
enum { FD_PERLANE = 1, FD_NOU2 = 2 };

struct LaneOp {
  int code, i0, i1, i2;
  double a, b;
  const void* p0;
  const void* p1;
  const void* p2;
  const void* p3;
};

struct LaneProg {
  int nops;
  int LP;         // shared-memory lane pitch in doubles (multiple of 4, >= every length used)
  int in_tiles;   // 4x4 tiles per lane of arrays in the orientation being read
  int out_tiles;  // tiles per row of the transposed orientation (= number of lane groups)
  int TPL;        // threads per lane (blockDim.x = 4*TPL)
  int C;          // pairs per thread per lane (= E+1 of the kernel instance), 2*C*TPL >= LP
  int group0;     // first lane group of this launch (multi-GPU slabs)
  int groups_per_rank;  // for ST_PEER: owner(J) = J / groups_per_rank (destination orientation)
  int rank;       // this GPU's rank (peer table index)
  int pad_;
  unsigned long long* prof;   // optional per-op cycle counters (64 entries), null in production
  int LN;         // lanes per CTA: 4 (a whole lane group) or 2 (half a group; grid = 2 x groups)
  int pad2_;
  LaneOp ops[B2_MAXOPS];
};

#ifdef B2_EMU
template <class T> static inline T ldg(const T* p) { return *p; }
#else
template <class T> __device__ __forceinline__ T ldg(const T* p) { return __ldg(p); }   // LDG.CONSTANT: read-only tables / coefficients
#endif
typedef double2 cplx;
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ cplx cmulmi(cplx a) { return make_double2(a.y, -a.x); }  // a * (-i)

// ---------------------------------------------------------------------------------------------
// Radix-R DFT in registers (forward, e^{-2 pi i jk/R}), natural order in and out.
// ---------------------------------------------------------------------------------------------
template <int R> struct Dft;
template <> struct Dft<1> { static __device__ __forceinline__ void run(cplx*) {} };
template <> struct Dft<2> {
  static __device__ __forceinline__ void run(cplx* v) {
    cplx a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};
template <> struct Dft<4> {
  static __device__ __forceinline__ void run(cplx* v) {
    cplx t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    cplx t2 = cadd(v[1], v[3]), t3 = cmulmi(csub(v[1], v[3]));
    v[0] = cadd(t0, t2);
    v[1] = cadd(t1, t3);
    v[2] = csub(t0, t2);
    v[3] = csub(t1, t3);
  }
};
template <> struct Dft<8> {
  static __device__ __forceinline__ void run(cplx* v) {
    cplx e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
    Dft<4>::run(e);
    Dft<4>::run(o);
    const double h = 0.70710678118654752440;
    cplx t1 = make_double2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));   // * e^{-i pi/4}
    cplx t2 = cmulmi(o[2]);                                                 // * (-i)
    cplx t3 = make_double2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));  // * e^{-3i pi/4}
    v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], t1);   v[5] = csub(e[1], t1);
    v[2] = cadd(e[2], t2);   v[6] = csub(e[2], t2);
    v[3] = cadd(e[3], t3);   v[7] = csub(e[3], t3);
  }
};
template <> struct Dft<16> {
  static __device__ __forceinline__ void run(cplx* v) {
    cplx e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    Dft<8>::run(e);
    Dft<8>::run(o);
    const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
    const cplx w[8] = {{1, 0}, {c1, -s1}, {h, -h}, {s1, -c1}, {0, -1}, {-s1, -c1}, {-h, -h}, {-c1, -s1}};
    v[0] = cadd(e[0], o[0]); v[8] = csub(e[0], o[0]);
#pragma unroll
    for (int k = 1; k < 8; k++) {
      cplx t = (k == 4) ? cmulmi(o[4]) : cmul(w[k], o[k]);
      v[k] = cadd(e[k], t);
      v[k + 8] = csub(e[k], t);
    }
  }
};

// ---------------------------------------------------------------------------------------------
// shuffle helpers for small structs of doubles
// ---------------------------------------------------------------------------------------------
template <int K> struct DVec { double d[K]; };
template <int K> __device__ __forceinline__ DVec<K> shfl_up(const DVec<K>& m, int delta, int width) {
  DVec<K> r;
#pragma unroll
  for (int i = 0; i < K; i++) r.d[i] = __shfl_up_sync(0xffffffffu, m.d[i], delta, width);
  return r;
}
template <int K> __device__ __forceinline__ DVec<K> shfl_down(const DVec<K>& m, int delta, int width) {
  DVec<K> r;
#pragma unroll
  for (int i = 0; i < K; i++) r.d[i] = __shfl_down_sync(0xffffffffu, m.d[i], delta, width);
  return r;
}

// Affine maps used by the lane recurrences.  "then(f, s)" = apply f first, then s.
// First order, two independent parities:  y -> A y + B.      d = {A0,B0,A1,B1}
struct Aff1 {
  typedef DVec<4> V;
  static __device__ __forceinline__ V identity() { V v; v.d[0] = 1; v.d[1] = 0; v.d[2] = 1; v.d[3] = 0; return v; }
  static __device__ __forceinline__ V then(const V& f, const V& s) {
    V r;
    r.d[0] = s.d[0] * f.d[0]; r.d[1] = fma(s.d[0], f.d[1], s.d[1]);
    r.d[2] = s.d[2] * f.d[2]; r.d[3] = fma(s.d[2], f.d[3], s.d[3]);
    return r;
  }
};
// Second order, two parities: state (u,w) -> P (u,w) + p.   d = {P00,P01,P10,P11,p0,p1} x 2
struct Aff2 {
  typedef DVec<12> V;
  static __device__ __forceinline__ V identity() {
    V v;
#pragma unroll
    for (int h = 0; h < 2; h++) { v.d[6*h+0] = 1; v.d[6*h+1] = 0; v.d[6*h+2] = 0; v.d[6*h+3] = 1; v.d[6*h+4] = 0; v.d[6*h+5] = 0; }
    return v;
  }
  static __device__ __forceinline__ V then(const V& f, const V& s) {
    V r;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const double* F = f.d + 6 * h; const double* S = s.d + 6 * h; double* R = r.d + 6 * h;
      R[0] = S[0] * F[0] + S[1] * F[2]; R[1] = S[0] * F[1] + S[1] * F[3];
      R[2] = S[2] * F[0] + S[3] * F[2]; R[3] = S[2] * F[1] + S[3] * F[3];
      R[4] = S[0] * F[4] + S[1] * F[5] + S[4];
      R[5] = S[2] * F[4] + S[3] * F[5] + S[5];
    }
    return r;
  }
};

// Exclusive scan of per-thread maps across the TPL threads of one lane.
// PREFIX: result = composition of the maps of threads q' < q (lowest applied first).
// SUFFIX: result = composition of the maps of threads q' > q (highest applied first).
// scratch: shared, >= 4 lanes * 8 warps entries of M::V.  All threads of the CTA must call.
template <class M, bool SUFFIX>
__device__ __forceinline__ typename M::V lane_scan_excl(typename M::V mine, int TPL, typename M::V* scratch) {
  typedef typename M::V V;
  const int tid = threadIdx.x;
  const int q = tid % TPL, lane = tid / TPL;
  const int width = TPL < 32 ? TPL : 32;
  const int qi = q % width;   // position inside the shuffle segment
  V inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    if (d < width) {
      V o = SUFFIX ? shfl_down(inc, d, width) : shfl_up(inc, d, width);
      bool take = SUFFIX ? (qi + d < width) : (qi >= d);
      if (take) inc = M::then(o, inc);
    }
  }
  V exc = SUFFIX ? shfl_down(inc, 1, width) : shfl_up(inc, 1, width);
  if (SUFFIX ? (qi == width - 1) : (qi == 0)) exc = M::identity();
  if (TPL > 32) {
    const int nw = TPL / 32, w = q / 32;
    if (SUFFIX ? (qi == 0) : (qi == 31)) scratch[lane * 8 + w] = inc;
    __syncthreads();
    V carry = M::identity();
    if (SUFFIX) { for (int k = nw - 1; k > w; k--) carry = M::then(carry, scratch[lane * 8 + k]); }
    else        { for (int k = 0; k < w; k++) carry = M::then(carry, scratch[lane * 8 + k]); }
    exc = M::then(carry, exc);
    __syncthreads();
  }
  return exc;
}

// sum over the TPL threads of a lane (result valid in every thread of the lane)
__device__ __forceinline__ double lane_sum(double v, int TPL, double* scratch) {
  const int tid = threadIdx.x, q = tid % TPL, lane = tid / TPL;
  const int width = TPL < 32 ? TPL : 32;
  for (int d = width >> 1; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d, width);
  if (TPL > 32) {
    const int nw = TPL / 32, w = q / 32;
    if ((q & 31) == 0) scratch[lane * 8 + w] = v;
    __syncthreads();
    double s = 0;
    for (int k = 0; k < nw; k++) s += scratch[lane * 8 + k];
    __syncthreads();
    v = s;
  }
  return v;
}

// ---------------------------------------------------------------------------------------------
// ops
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ void op_load(const LaneProg& P, const LaneOp& op, double* __restrict__ W, int gl) {
  const int T = blockDim.x, LP = P.LP;
  const int LN = P.LN, psh = LN == 4 ? 3 : 2, pm = 2 * LN - 1;   // 2*LN 16-byte pieces per tile belong to this CTA
  const int lb = (blockIdx.x & ((4 / LN) - 1)) * LN;             // first lane of the group handled by this CTA
  const int npieces = LN * LP / 2;
  const int len = op.i0;
  const double a = op.a;
  const bool acc = op.i2 & LD_ACC, mul = op.i2 & LD_MUL, plain = op.i2 & LD_PLAIN;
  const double2* src = reinterpret_cast<const double2*>(op.p0);
  const size_t slab = (size_t)gl * P.in_tiles * 8;  // in double2 units
  const double* sc = reinterpret_cast<const double*>(op.p1);
  const bool sten = op.i2 & LD_STENCIL;
  constexpr int U = 8;   // loads in flight per thread: all U global loads are issued before the first use
  for (int p0 = threadIdx.x; p0 < npieces; p0 += U * T) {
    double2 v[U], u[U];
#pragma unroll
    for (int k = 0; k < U; k++) {
      const int pidx = p0 + k * T;
      const int J = pidx >> psh, l = (pidx & pm) >> 1, j0 = 4 * J + (pidx & 1) * 2;
      v[k] = make_double2(0.0, 0.0); u[k] = make_double2(0.0, 0.0);
      if (pidx < npieces && J < P.in_tiles && j0 < len) {
        v[k] = plain ? src[((size_t)(4 * gl + lb + l) * P.in_tiles * 4 + j0) >> 1]
                     : src[slab + (size_t)J * 8 + (lb + l) * 2 + (pidx & 1)];
        if (sten && j0 >= 2) {   // composite -> orthonormal on the fly: + p1[j] * src[j-2]  (tiled sources only)
          const int jm = j0 - 2;
          u[k] = src[slab + ((size_t)(jm >> 2) * 16 + (lb + l) * 4 + (jm & 3)) / 2];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < U; k++) {
      const int pidx = p0 + k * T;
      if (pidx >= npieces) break;
      const int J = pidx >> psh, l = (pidx & pm) >> 1, j0 = 4 * J + (pidx & 1) * 2;
      double2 x = v[k];
      if (sten && j0 >= 2 && j0 < len) { x.x = fma(sc[j0], u[k].x, x.x); x.y = fma(sc[j0 + 1], u[k].y, x.y); }
      x.x *= a;
      x.y = (j0 + 1 < len) ? x.y * a : 0.0;
      double2* w = reinterpret_cast<double2*>(W + l * LP + j0);
      if (acc) { double2 o = *w; x.x += o.x; x.y += o.y; }
      else if (mul) { double2 o = *w; x.x *= o.x; x.y *= o.y; }
      *w = x;
    }
  }
  __syncthreads();
}

__device__ __noinline__ void op_store(const LaneProg& P, const LaneOp& op, const double* __restrict__ W, int g, int gl) {
  const int T = blockDim.x, LP = P.LP;
  const int LN = P.LN, psh = LN == 4 ? 3 : 2, pm = 2 * LN - 1, hl = LN >> 1;
  const int lb = (blockIdx.x & ((4 / LN) - 1)) * LN;
  const int npieces = P.in_tiles * 2 * LN;
  const int len = op.i0;
  const double a = op.a;
  const int flags = op.i2;
  double2* dst = reinterpret_cast<double2*>(const_cast<void*>(op.p0));
  if (flags & ST_TRANS) {
    double* const* peers = reinterpret_cast<double* const*>(op.p1);
    for (int pidx = threadIdx.x; pidx < npieces; pidx += T) {
      const int r = pidx & pm;
      int J = pidx >> psh, jl = r / hl, l0 = (r % hl) * 2, j = 4 * J + jl;
      double2 v = make_double2(0.0, 0.0);
      if (j < len) { v.x = a * W[l0 * LP + j]; v.y = a * W[(l0 + 1) * LP + j]; }
      double2* d = dst;
      int Jl = J;
      if (flags & ST_PEER) {  // row block J of the transposed array lives on rank J / groups_per_rank
        int owner = J / P.groups_per_rank;
        Jl = J - owner * P.groups_per_rank;
        d = reinterpret_cast<double2*>(reinterpret_cast<char*>(peers[owner]) +
                                       (reinterpret_cast<const char*>(op.p0) - reinterpret_cast<const char*>(peers[P.rank])));
      }
      // tiled: tile (Jl, g) holds [jl][l];  row-major ("plain", for the GEMM): row 4*Jl+jl, columns 4g+l
      size_t idx = (flags & ST_PLAIN) ? (((size_t)(4 * Jl + jl) * P.out_tiles * 4 + 4 * g + lb + l0) >> 1)
                                      : ((((size_t)Jl * P.out_tiles + g) * 16 + jl * 4 + lb + l0) >> 1);
      if (flags & ST_ACC) { double2 o = d[idx]; v.x += o.x; v.y += o.y; }
      d[idx] = v;
    }
  } else {
    const size_t slab = (size_t)gl * P.in_tiles * 8;
    for (int pidx = threadIdx.x; pidx < npieces; pidx += T) {
      int J = pidx >> psh, l = (pidx & pm) >> 1, j0 = 4 * J + (pidx & 1) * 2;
      double2 v = make_double2(0.0, 0.0);
      if (j0 < len) {
        double2 w = *reinterpret_cast<const double2*>(W + l * LP + j0);
        v.x = a * w.x;
        v.y = (j0 + 1 < len) ? a * w.y : 0.0;
      }
      size_t idx = (flags & ST_PLAIN) ? (((size_t)(4 * gl + lb + l) * P.in_tiles * 4 + j0) >> 1)
                                      : (slab + (size_t)J * 8 + (lb + l) * 2 + (pidx & 1));
      if (flags & ST_ACC) { double2 o = dst[idx]; v.x += o.x; v.y += o.y; }
      dst[idx] = v;
    }
  }
  __syncthreads();
}

// ---- chunked lane ops -------------------------------------------------------------------------
// Thread q of a lane owns CP consecutive PAIRS (2p, 2p+1), p = q*CP + t.  All banded operators of the
// path couple elements at even distance only, so every recurrence is a plain double2 recurrence over
// pairs (no parity bookkeeping), shared-memory accesses are 16-byte with stride CP = E+1 (odd: no bank
// conflicts) and coefficient vectors are stored "pair/scan" ordered, [t][q] as double2 (coalesced).
__device__ __forceinline__ double2 d2(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ double2 d2fma(double2 a, double2 b, double2 c) { return make_double2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }
__device__ __forceinline__ double2 d2mul(double2 a, double2 b) { return make_double2(a.x * b.x, a.y * b.y); }

template <int CP>
__device__ __noinline__ void op_band(const LaneProg& P, const LaneOp& op, double* __restrict__ W) {
  // y_i = sum_m c_m[i] x_{i+o_m}, o_m even.  Thread q of a lane takes the pairs p = q + t*TPL (coalesced
  // coefficient loads in natural order, conflict-free shared-memory reads); results are staged in
  // registers because the operation is in place.
  const int TPL = P.TPL, HP = P.LP >> 1;
  const int q = threadIdx.x % TPL, l = threadIdx.x / TPL;
  const int len_out = op.i0;
  const int h0 = (int)(signed char)(op.i1 & 0xff), h1 = (int)(signed char)((op.i1 >> 8) & 0xff), h2 = (int)(signed char)((op.i1 >> 16) & 0xff);
  const double2* __restrict__ c0 = (const double2*)op.p0;
  const double2* __restrict__ c1 = (const double2*)op.p1;
  const double2* __restrict__ c2 = (const double2*)op.p2;
  const double2* w2 = reinterpret_cast<const double2*>(W + l * P.LP);
  const double2 zero = d2(0.0, 0.0);
  double2 y[CP];
#pragma unroll
  for (int t = 0; t < CP; t++) y[t] = zero;
  auto term = [&](int h, const double2* __restrict__ c) {
    const int hp = h >> 1;
#pragma unroll
    for (int t = 0; t < CP; t++) {
      const int p = q + t * TPL, pp = p + hp;
      const bool ok = pp >= 0 && pp < HP && p < HP;
      const double2 x = w2[ok ? pp : 0];
      const double2 cc = c ? ldg(c + (p < HP ? p : 0)) : d2(1.0, 1.0);
      if (ok) y[t] = d2fma(cc, x, y[t]);
    }
  };
  if (h0 != 127) term(h0, c0);
  if (h1 != 127) term(h1, c1);
  if (h2 != 127) term(h2, c2);
  __syncthreads();
  double2* wo = reinterpret_cast<double2*>(W + l * P.LP);
#pragma unroll
  for (int t = 0; t < CP; t++) {
    const int p = q + t * TPL;
    double2 v = y[t];
    if (2 * p >= len_out) v.x = 0.0;
    if (2 * p + 1 >= len_out) v.y = 0.0;
    if (p < HP) wo[p] = v;
  }
  __syncthreads();
}

// Chebyshev derivative: b_k = S_{k+1},  S_m = 2 m a_m + S_{m+2};  b_0 *= 1/2;  result * scale.
// In pairs: S[p] = (2(2p) a_2p, 2(2p+1) a_2p+1) + S[p+1];  out[p] = (S[p].y, S[p+1].x).
template <int CP>
__device__ __noinline__ void op_deriv(const LaneProg& P, const LaneOp& op, double* __restrict__ W, void* scratch) {
  const int TPL = P.TPL, HP = P.LP >> 1;
  const int q = threadIdx.x % TPL, l = threadIdx.x / TPL;
  double2* w2 = reinterpret_cast<double2*>(W + l * P.LP);
  for (int rep = 0; rep < op.i1; rep++) {
    double2 tp[CP];
    double2 tot = d2(0.0, 0.0);
#pragma unroll
    for (int t = 0; t < CP; t++) {
      const int p = q * CP + t;
      double2 a = (p < HP) ? w2[p < HP ? p : 0] : d2(0.0, 0.0);
      tp[t] = d2(2.0 * (2 * p) * a.x, 2.0 * (2 * p + 1) * a.y);
      tot.x += tp[t].x; tot.y += tp[t].y;
    }
    Aff1::V m; m.d[0] = 1; m.d[1] = tot.x; m.d[2] = 1; m.d[3] = tot.y;
    Aff1::V inc = lane_scan_excl<Aff1, true>(m, TPL, (Aff1::V*)scratch);
    double2 S = d2(inc.d[1], inc.d[3]);   // S of the first pair of the next chunk
    const double sc = (rep == op.i1 - 1) ? op.a : 1.0;
#pragma unroll
    for (int t = CP - 1; t >= 0; t--) {
      const int p = q * CP + t;
      const double nx = S.x;
      S.x += tp[t].x; S.y += tp[t].y;
      double2 o = d2(S.y * sc, nx * sc);
      if (p == 0) o.x *= 0.5;
      if (p < HP) w2[p] = o;
    }
    __syncthreads();
  }
}

// In-place solve of the LU-factored 4-diagonal (-2,0,+2,+4) system (reference: src/solver/fdma.rs:101-118):
//   forward:  x_i -= fl_i x_{i-2}                     (fl_i = swept low_{i-2})
//   backward: x_i = (x_i - u1_i x_{i+2} - u2_i x_{i+4}) * id_i
// As pair recurrences: y_p = b_p - fl_p * y_{p-1};  x_p = (y_p - u1_p x_{p+1} - u2_p x_{p+2}) id_p.
// Each thread reduces its CP pairs to an affine map; maps are combined by a scan across the lane's threads.
template <int CP>
__device__ __noinline__ void op_fdma(const LaneProg& P, const LaneOp& op, double* __restrict__ W, int gl, void* scratch) {
  const int TPL = P.TPL, HP = P.LP >> 1;
  const int q = threadIdx.x % TPL, l = threadIdx.x / TPL;
  const int n = op.i0;
  double2* w2 = reinterpret_cast<double2*>(W + l * P.LP);
  const double2* __restrict__ cfl = (const double2*)op.p0; const double2* __restrict__ cid = (const double2*)op.p1;
  const double2* __restrict__ cu1 = (const double2*)op.p2; const double2* __restrict__ cu2 = (const double2*)op.p3;
  // shared vectors: [t][q]; per-lane arrays: [group][t][lane][q]  (double2 units, coalesced at every step)
  size_t base; int stride;
  if (op.i2 & FD_PERLANE) { const int lb = (blockIdx.x & ((4 / P.LN) - 1)) * P.LN; stride = 4 * TPL; base = ((size_t)gl * CP * 4 + lb + l) * TPL + q; }
  else { stride = TPL; base = q; }
  const bool nou2 = op.i2 & FD_NOU2;
  const double2 zero = d2(0.0, 0.0);
  const int p0 = q * CP;
  auto rd = [&](int t) -> double2 {   // right-hand side / intermediate at pair p0+t, zero outside [0, n)
    const int p = p0 + t;
    const bool ok = p < HP;
    double2 v = ok ? w2[ok ? p : 0] : zero;
    if (2 * p >= n) v.x = 0.0;
    if (2 * p + 1 >= n) v.y = 0.0;
    return v;
  };
  // ---- forward elimination: y_p = b_p - fl_p y_{p-1} ----
  {
    double2 A = d2(1.0, 1.0), B = zero;
#pragma unroll 6
    for (int t = 0; t < CP; t++) {
      const double2 f = ldg(cfl + base + (size_t)t * stride), b = rd(t);
      B = d2(fma(-f.x, B.x, b.x), fma(-f.y, B.y, b.y));
      A = d2(-f.x * A.x, -f.y * A.y);
    }
    Aff1::V m; m.d[0] = A.x; m.d[1] = B.x; m.d[2] = A.y; m.d[3] = B.y;
    Aff1::V inc = lane_scan_excl<Aff1, false>(m, TPL, (Aff1::V*)scratch);
    double2 y = d2(inc.d[1], inc.d[3]);   // y of the last pair before this chunk (the start state is 0)
#pragma unroll 6
    for (int t = 0; t < CP; t++) {
      const double2 f = ldg(cfl + base + (size_t)t * stride), b = rd(t);
      y = d2(fma(-f.x, y.x, b.x), fma(-f.y, y.y, b.y));
      if (p0 + t < HP) w2[p0 + t] = y;
    }
  }
  // every thread only touched its own chunk: no barrier needed before the back substitution
  // ---- back substitution: x_p = (y_p - u1_p x_{p+1} - u2_p x_{p+2}) id_p ----
  {
    Aff2::V m = Aff2::identity();
#pragma unroll 6
    for (int t = CP - 1; t >= 0; t--) {
      const size_t k = base + (size_t)t * stride;
      const double2 idv = ldg(cid + k), u1 = ldg(cu1 + k), u2 = nou2 ? zero : ldg(cu2 + k), y = rd(t);
      const double2 m0 = d2(-u1.x * idv.x, -u1.y * idv.y), m1 = d2(-u2.x * idv.x, -u2.y * idv.y), g0 = d2(y.x * idv.x, y.y * idv.y);
      double* M = m.d;   // compose onto the chunk map; state = (x_{p+1}, x_{p+2}) per component
      double r0 = m0.x * M[0] + m1.x * M[2], r1 = m0.x * M[1] + m1.x * M[3], rp = m0.x * M[4] + m1.x * M[5] + g0.x;
      M[2] = M[0]; M[3] = M[1]; M[5] = M[4]; M[0] = r0; M[1] = r1; M[4] = rp;
      M = m.d + 6;
      r0 = m0.y * M[0] + m1.y * M[2]; r1 = m0.y * M[1] + m1.y * M[3]; rp = m0.y * M[4] + m1.y * M[5] + g0.y;
      M[2] = M[0]; M[3] = M[1]; M[5] = M[4]; M[0] = r0; M[1] = r1; M[4] = rp;
    }
    Aff2::V inc = lane_scan_excl<Aff2, true>(m, TPL, (Aff2::V*)scratch);
    double2 s1 = d2(inc.d[4], inc.d[10]), s2 = d2(inc.d[5], inc.d[11]);   // x_{p+1}, x_{p+2} entering the chunk
#pragma unroll 6
    for (int t = CP - 1; t >= 0; t--) {
      const size_t k = base + (size_t)t * stride;
      const double2 idv = ldg(cid + k), u1 = ldg(cu1 + k), u2 = nou2 ? zero : ldg(cu2 + k), y = rd(t);
      double2 x = d2((y.x - u1.x * s1.x - u2.x * s2.x) * idv.x, (y.y - u1.y * s1.y - u2.y * s2.y) * idv.y);
      s2 = s1; s1 = x;
      if (p0 + t < HP) w2[p0 + t] = x;
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// complex FFT of Nc points per lane, in place in shared memory (Stockham autosort, register-staged:
// every thread reads its E points, the CTA syncs, then everything is written back).
// tw[t] = exp(-2 pi i t / Nc)
// ---------------------------------------------------------------------------------------------
template <int E, int R>
__device__ __forceinline__ void fft_stage(double* __restrict__ wl, int Nc, int Ns, int q, int TPL, const cplx* __restrict__ tw) {
  constexpr int NB = E / R;
  cplx v[E];
  const int stride = Nc / R;
#pragma unroll
  for (int b = 0; b < NB; b++) {
    int j = q + b * TPL;
#pragma unroll
    for (int r = 0; r < R; r++) v[b * R + r] = *reinterpret_cast<const cplx*>(wl + 2 * (j + r * stride));
  }
  __syncthreads();
#pragma unroll
  for (int b = 0; b < NB; b++) {
    int j = q + b * TPL;
    int k = j % Ns;
    if (Ns > 1) {
      int tstep = k * (Nc / (Ns * R));
#pragma unroll
      for (int r = 1; r < R; r++) v[b * R + r] = cmul(v[b * R + r], ldg(tw + r * tstep));
    }
    Dft<R>::run(v + b * R);
    int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; r++) *reinterpret_cast<cplx*>(wl + 2 * (j0 + r * Ns)) = v[b * R + r];
  }
  __syncthreads();
}

template <int E>
__device__ __forceinline__ void lane_fft(double* __restrict__ W, int LP, int Nc, int TPL, const cplx* __restrict__ tw) {
  const int q = threadIdx.x % TPL, l = threadIdx.x / TPL;
  double* wl = W + l * LP;
  int Ns = 1;
  // radix plan: as many radix-E passes as fit, then one pass with the remainder (1, 2, 4 or 8)
  while (Nc / Ns >= E) { fft_stage<E, E>(wl, Nc, Ns, q, TPL, tw); Ns *= E; }
  const int rem = Nc / Ns;
  if constexpr (E >= 16) { if (rem == 8) fft_stage<E, 8>(wl, Nc, Ns, q, TPL, tw); }
  if constexpr (E >= 8) { if (rem == 4) fft_stage<E, 4>(wl, Nc, Ns, q, TPL, tw); }
  if (rem == 2) fft_stage<E, 2>(wl, Nc, Ns, q, TPL, tw);
}

// Chebyshev transform (DCT-I of n = N+1 points on Gauss-Lobatto nodes x_j = -cos(pi j/N)) through ONE
// complex FFT of N/2 points:
//   mode 0 (forward):  c_k = (-1)^k X_k / N, c_0 and c_N halved,  X = DCT-I(v)
//   mode 1 (backward): v = DCT-I(y)/2, y_k = (-1)^k c_k, y_0 and y_N doubled
// tw: exp(-2 pi i t/(N/2)), tw2[j] = exp(-2 pi i j/N) (j <= N/2), isin[k] = 1/(4 sin(pi k/N))
template <int E>
__device__ __noinline__ void op_dct(const LaneProg& P, const LaneOp& op, double* __restrict__ W, double* scratch) {
  const int TPL = P.TPL, LP = P.LP;
  const int q = threadIdx.x % TPL, l = threadIdx.x / TPL;
  const int N = op.i0 - 1, M = N >> 1, mode = op.i1;
  const cplx* tw = (const cplx*)op.p0; const cplx* tw2 = (const cplx*)op.p1; const double* isin = (const double*)op.p2;
  double* w = W + l * LP;
  constexpr int NP = E / 2 + 1;
  const cplx* w2 = reinterpret_cast<const cplx*>(w);
  // ---- pre: x -> g (N/2 complex), pairs (j, M-j); branch-free, 16-byte shared-memory reads ----
  cplx gj[NP], gm[NP];
  double r0 = 0.0;
  const double sg = (mode == 1) ? -1.0 : 1.0, endf = (mode == 1) ? 2.0 : 1.0;   // backward: y_k = (-1)^k c_k, ends doubled
#pragma unroll
  for (int pi = 0; pi < NP; pi++) {
    const int j0 = q + pi * TPL;
    const int j = j0 <= M / 2 ? j0 : M / 2, jm = M - j;
    const cplx pj = w2[j], pjl = w2[j > 0 ? j - 1 : 0], pm = w2[jm], pml = w2[jm - 1];
    const double xo_p = sg * pj.y;                              // x_{2j+1}
    const double xo_m = (j == 0) ? xo_p : sg * pjl.y;           // x_{2j-1}, x_{-1} = x_1
    const double xm_m = sg * pml.y;                             // x_{2jm-1}
    const double xm_p = (jm == M) ? xm_m : sg * pm.y;           // x_{2jm+1}, x_{N+1} = x_{N-1}
    const cplx zj = make_double2(pj.x * (j == 0 ? endf : 1.0), xo_p - xo_m);
    const cplx zmc = make_double2(pm.x * (jm == M ? endf : 1.0), -(xm_p - xm_m));   // conj(z_{M-j})
    const cplx e = cadd(zj, zmc), d = cmul(csub(zj, zmc), ldg(tw2 + j));
    gj[pi] = make_double2(e.x - d.y, e.y + d.x);           // e + i d
    gm[pi] = make_double2(e.x + d.y, -e.y + d.x);          // conj(e) + i conj(d)
    r0 += (j0 < M / 2) ? (xo_p + xm_m) : 0.0;
  }
  r0 = 2.0 * lane_sum(r0, TPL, scratch);   // R_0 = 2 * sum of odd samples
  __syncthreads();
#pragma unroll
  for (int pi = 0; pi < NP; pi++) {
    const int j = q + pi * TPL;
    if (j <= M / 2) *reinterpret_cast<cplx*>(w + 2 * j) = gj[pi];
    if (j > 0 && j < M / 2) *reinterpret_cast<cplx*>(w + 2 * (M - j)) = gm[pi];
  }
  __syncthreads();
  lane_fft<E>(W, LP, M, TPL, tw);
  // ---- post: Z (N reals) -> X (N+1), pairs (k, N-k), 1 <= k <= M-1 in a branch-free unrolled loop ----
  const double fs = (mode == 0) ? 1.0 / N : 0.5;
#pragma unroll
  for (int pi = 0; pi < E + 1; pi++) {
    const int k0 = q + pi * TPL;
    const bool ok = k0 >= 1 && k0 <= M - 1;
    const int k = ok ? k0 : 1;
    const double zk = w[k], zn = w[N - k];
    const double A = 0.5 * (zk + zn), R = (zn - zk) * ldg(isin + k);
    const double sk = (mode == 0 && (k & 1)) ? -fs : fs;      // N is even: k and N-k have the same parity
    if (ok) { w[k] = (A + R) * sk; w[N - k] = (A - R) * sk; }
  }
  if (q == 0) {   // k = 0 (and N), k = M: untouched by the loop above
    const double z0 = w[0], e0 = (mode == 0) ? 0.5 * fs : fs;
    w[0] = (z0 + r0) * e0;
    w[N] = (z0 - r0) * e0;
    w[M] = w[M] * ((mode == 0 && (M & 1)) ? -fs : fs);
  }
  __syncthreads();
}

// Real FFT of n points along the lane (Fourier axis): forward r2c is unnormalised,
// n/2+1 interleaved complex modes; backward c2r carries 1/n and ignores Im of the k=0 and k=n/2 modes.
template <int E>
__device__ __noinline__ void op_rfft(const LaneProg& P, const LaneOp& op, double* __restrict__ W) {
  const int TPL = P.TPL, LP = P.LP;
  const int q = threadIdx.x % TPL, l = threadIdx.x / TPL;
  const int n = op.i0, M = n >> 1, mode = op.i1;
  const cplx* tw = (const cplx*)op.p0; const cplx* tw2 = (const cplx*)op.p1;
  double* w = W + l * LP;
  if (mode == 0) {
    lane_fft<E>(W, LP, M, TPL, tw);
    for (int k = q; k <= M / 2; k += TPL) {
      if (k == 0) {
        cplx z = *reinterpret_cast<cplx*>(w);
        *reinterpret_cast<cplx*>(w) = make_double2(z.x + z.y, 0.0);
        *reinterpret_cast<cplx*>(w + 2 * M) = make_double2(z.x - z.y, 0.0);
      } else {
        cplx zk = *reinterpret_cast<cplx*>(w + 2 * k), zm = cconj(*reinterpret_cast<cplx*>(w + 2 * (M - k)));
        cplx S = cadd(zk, zm), D = cmul(ldg(tw2 + k), csub(zk, zm));   // w_k D
        // X_k = (S - i wD)/2 ; X_{M-k} = conj((S + i wD)/2)
        *reinterpret_cast<cplx*>(w + 2 * k) = make_double2(0.5 * (S.x + D.y), 0.5 * (S.y - D.x));
        if (k != M - k) *reinterpret_cast<cplx*>(w + 2 * (M - k)) = make_double2(0.5 * (S.x - D.y), -0.5 * (S.y + D.x));
      }
    }
    __syncthreads();
  } else {
    for (int k = q; k <= M / 2; k += TPL) {
      if (k == 0) {
        double x0 = w[0], xm = w[2 * M];
        // Zc_0 = ((x0+xm) + i(x0-xm))/2 ; FFT input is conj(Zc)
        *reinterpret_cast<cplx*>(w) = make_double2(0.5 * (x0 + xm), -0.5 * (x0 - xm));
      } else {
        cplx xk = *reinterpret_cast<cplx*>(w + 2 * k), xm = cconj(*reinterpret_cast<cplx*>(w + 2 * (M - k)));
        cplx S = cadd(xk, xm), D = cmul(cconj(ldg(tw2 + k)), csub(xk, xm));  // conj(w_k) D'
        // Zc_k = (S + i cD)/2 ; Zc_{M-k} = conj((S - i cD)/2); store conjugates
        *reinterpret_cast<cplx*>(w + 2 * k) = make_double2(0.5 * (S.x - D.y), -0.5 * (S.y + D.x));
        if (k != M - k) *reinterpret_cast<cplx*>(w + 2 * (M - k)) = make_double2(0.5 * (S.x + D.y), 0.5 * (S.y - D.x));
      }
    }
    __syncthreads();
    lane_fft<E>(W, LP, M, TPL, tw);
    const double s = 1.0 / M;
    for (int e = q; e < LP; e += TPL) {
      double v = w[e];
      w[e] = (e < n) ? ((e & 1) ? -v * s : v * s) : 0.0;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void op_pointwise(const LaneProg& P, const LaneOp& op, double* __restrict__ W, int g) {
  const int TPL = P.TPL, LP = P.LP;
  const int lb = (blockIdx.x & ((4 / P.LN) - 1)) * P.LN;
  const int q = threadIdx.x % TPL, l = threadIdx.x / TPL;
  double* w = W + l * LP;
  switch (op.code) {
    case OP_FDIFF: {   // interleaved complex: (re, im) *= (i k)^d * a
      const int m = op.i0, d = op.i1 & 3;
      for (int k = q; k < m; k += TPL) {
        cplx c = *reinterpret_cast<cplx*>(w + 2 * k);
        double f = op.a;
        for (int t = 0; t < op.i1; t++) f *= (double)k;
        cplx r;
        if (d == 0) r = make_double2(c.x * f, c.y * f);
        else if (d == 1) r = make_double2(-c.y * f, c.x * f);
        else if (d == 2) r = make_double2(-c.x * f, -c.y * f);
        else r = make_double2(c.y * f, -c.x * f);
        *reinterpret_cast<cplx*>(w + 2 * k) = r;
      }
    } break;
    case OP_SCALEVEC: {
      const double* v = (const double*)op.p0;
      for (int e = q; e < op.i0; e += TPL) w[e] *= ldg(v + (e >> op.i1));
    } break;
    case OP_ZEROTAIL:
      for (int e = op.i0 + q; e < LP; e += TPL) w[e] = 0.0;
      break;
    case OP_LANEMASK:
      if (4 * g + lb + l >= op.i0) for (int e = q; e < LP; e += TPL) w[e] = 0.0;
      break;
    case OP_ZEROELEM:
      if (4 * g + lb + l == op.i0 && q == 0) w[op.i1] = 0.0;
      break;
    case OP_SCALE:
      for (int e = q; e < LP; e += TPL) w[e] *= op.a;
      break;
  }
  __syncthreads();
}

template <int E>
__global__ void __launch_bounds__(512) lane_kernel(const __grid_constant__ LaneProg P) {
  B2_DYN_SMEM(double, smem);
  double* W = smem;                       // [LN][LP]
  void* scratch = smem + P.LN * P.LP;     // 32 * sizeof(DVec<12>) = 3 KB
  const int gl = blockIdx.x / (4 / P.LN); // local lane group (addresses this GPU's slab)
  const int g = P.group0 + gl;            // global lane group (mode indices, transposed stores)
  for (int o = 0; o < P.nops; o++) {
    const LaneOp& op = P.ops[o];
    long long t0 = 0;
    if (P.prof) t0 = clock64();
    switch (op.code) {
      case OP_LOAD: op_load(P, op, W, gl); break;
      case OP_STORE: op_store(P, op, W, g, gl); break;
      case OP_BAND: op_band<E + 1>(P, op, W); break;
      case OP_DERIV: op_deriv<E + 1>(P, op, W, scratch); break;
      case OP_FDMA: op_fdma<E + 1>(P, op, W, gl, scratch); break;
      case OP_DCT: op_dct<E>(P, op, W, (double*)scratch); break;
      case OP_RFFT: op_rfft<E>(P, op, W); break;
      default: op_pointwise(P, op, W, g); break;
    }
    if (P.prof && threadIdx.x == 0) {
      atomicAdd(P.prof + op.code, (unsigned long long)(clock64() - t0));
      atomicAdd(P.prof + 32 + op.code, 1ull);
    }
  }
}
