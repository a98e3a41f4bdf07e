// Core of the pruned z-axis complex-to-real transform (float64 arithmetic).
//
// A real inverse FFT of length M = 2h whose Hermitian half spectrum has only the first
// Kz entries non-zero is computed as one complex Stockham FFT of length h:
//     Z[k] = (X[k] + conj X[h-k]) + i e^{2 pi i k/M} (X[k] - conj X[h-k]),   k < h
//     z    = sum_k Z[k] e^{2 pi i k n / h}  (un-normalised),   x[2n] = Re z[n], x[2n+1] = Im z[n]
// The FFT runs in multi-radix Stockham stages (radices <= 16); one stage step for "thread" i is
//     k = i & (p-1);  j = (i-k) R + k;  u_m = x[i + m h/R] e^{2 pi i k m /(p R)};
//     (Y_0..Y_{R-1}) = inverse DFT_R(u);  y[j + m p] = Y_m
// Everything here is plain C++ so that tests/ can compile it with g++ and check it against a
// naive DFT without a GPU (tests/test_zfft_core.py); the CUDA kernel (zpass.cu) supplies the
// thread mapping, shared memory and synchronisation.
#pragma once

#ifdef __CUDACC__
#define ZF_HD __host__ __device__ __forceinline__
#else
#define ZF_HD inline
#endif

namespace zfft {

struct alignas(16) cplx {
  double x, y;
};

ZF_HD cplx cadd(cplx a, cplx b) { return {a.x + b.x, a.y + b.y}; }
ZF_HD cplx csub(cplx a, cplx b) { return {a.x - b.x, a.y - b.y}; }
ZF_HD cplx cmul(cplx a, cplx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
ZF_HD cplx cconj(cplx a) { return {a.x, -a.y}; }
ZF_HD cplx cmul_i(cplx a) { return {-a.y, a.x}; }  // a * i

// e^{2 pi i q / 16}
#define ZF_C1 0.92387953251128673848   /* cos(pi/8) */
#define ZF_S1 0.38268343236508978178   /* sin(pi/8) */
#define ZF_R2 0.70710678118654752440   /* sqrt(1/2) */

template <int Q>
ZF_HD cplx mul_root16(cplx a) {  // a * e^{2 pi i Q/16}, Q in [0,8)
  if (Q == 0) return a;
  if (Q == 4) return cmul_i(a);
  if (Q == 2) return {ZF_R2 * (a.x - a.y), ZF_R2 * (a.x + a.y)};
  if (Q == 6) return {-ZF_R2 * (a.x + a.y), ZF_R2 * (a.x - a.y)};
  if (Q == 1) return cmul(a, cplx{ZF_C1, ZF_S1});
  if (Q == 3) return cmul(a, cplx{ZF_S1, ZF_C1});
  if (Q == 5) return cmul(a, cplx{-ZF_S1, ZF_C1});
  return cmul(a, cplx{-ZF_C1, ZF_S1});  // Q == 7
}

template <int R>
ZF_HD constexpr int bitrev(int m) {
  int r = 0;
  for (int b = 1; b < R; b <<= 1) {
    r = (r << 1) | (m & 1);
    m >>= 1;
  }
  return r;
}

// One decimation-in-frequency level on blocks of length LEN (compile-time recursion keeps every
// index and twiddle a constant, so v[] lives in registers).
template <int R, int LEN, int BLK, int J>
struct Level {
  static ZF_HD void run(cplx (&v)[R]) {
    const cplx a = v[BLK + J], b = v[BLK + J + LEN / 2];
    v[BLK + J] = cadd(a, b);
    v[BLK + J + LEN / 2] = mul_root16<J * (16 / LEN)>(csub(a, b));
    if (J + 1 < LEN / 2)
      Level<R, LEN, BLK, (J + 1 < LEN / 2 ? J + 1 : 0)>::run(v);
    else if (BLK + LEN < R)
      Level<R, LEN, (BLK + LEN < R ? BLK + LEN : 0), 0>::run(v);
    else if (LEN > 2)
      Level<R, (LEN > 2 ? LEN / 2 : 2), 0, 0>::run(v);
  }
};

// In-register un-normalised inverse DFT of R points (R in {2,4,8,16}):
// afterwards Y[m] = v[bitrev<R>(m)].
template <int R>
ZF_HD void dft_inverse_bitrev(cplx (&v)[R]) {
  Level<R, R, 0, 0>::run(v);
}

// radices of the Stockham stages for a complex length h (power of two, >= 2)
ZF_HD int next_radix(int remaining) {
  if (remaining % 16 == 0) return 16;
  if (remaining % 8 == 0) return 8;
  if (remaining % 4 == 0) return 4;
  return 2;
}

// pre-processing pair: given X[k] and X[h-k] and w = e^{2 pi i k/M}, produce Z[k] and Z[h-k]
ZF_HD void pack_pair(cplx xk, cplx xhk, cplx w, cplx& zk, cplx& zhk) {
  const cplx e = cadd(xk, cconj(xhk));
  const cplx o = cmul(csub(xk, cconj(xhk)), w);
  zk = cadd(e, cmul_i(o));
  // Z[h-k] = conj(e) + i * (-conj(w)) * (-conj(xk - conj xhk)) = conj(e) + i * conj(o)
  zhk = cadd(cconj(e), cmul_i(cconj(o)));
}

}  // namespace zfft
