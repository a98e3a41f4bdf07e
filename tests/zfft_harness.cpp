// CPU harness for bskit_b200/csrc/zfft_core.h: emulates the thread mapping of
// zpass_c2r_kernel (pack + multi-radix Stockham stages) and compares with a naive inverse DFT.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../bskit_b200/csrc/zfft_core.h"
using zfft::cplx;
typedef std::complex<double> cd;

template <int R>
static void stage(std::vector<cplx>& x, int H, int p, const std::vector<cd>& w2h) {
  const int T = H / R;
  std::vector<cplx> y(H);
  for (int i = 0; i < T; ++i) {
    const int k = i & (p - 1), j = (i - k) * R + k, step = (H / (p * R)) * k;
    cplx v[R];
    for (int m = 0; m < R; ++m) {
      cplx u = x[i + m * T];
      if (m > 0 && p > 1) {
        cd w = w2h[(2 * step * m) & (2 * H - 1)];
        u = zfft::cmul(u, cplx{w.real(), w.imag()});
      }
      v[m] = u;
    }
    zfft::dft_inverse_bitrev<R>(v);
    for (int m = 0; m < R; ++m) y[j + m * p] = v[zfft::bitrev<R>(m)];
  }
  x = y;
}

int main() {
  double worst = 0;
  for (int M : {64, 128, 256, 512, 1024, 2048})
    for (int Kz : {M / 2 + 1, 42 < M / 2 ? 42 : M / 4}) {
      const int H = M / 2;
      std::vector<cd> X(H + 1, cd(0, 0)), w(M);
      for (int j = 0; j < M; ++j) w[j] = std::polar(1.0, 2 * M_PI * j / M);
      srand(M + Kz);
      for (int k = 0; k < Kz; ++k) X[k] = cd(rand() / (double)RAND_MAX - .5, rand() / (double)RAND_MAX - .5);
      X[0] = X[0].real();
      if (Kz == H + 1) X[H] = X[H].real();
      std::vector<cplx> row(H + 1);
      for (int k = 0; k <= H; ++k) row[k] = cplx{X[k].real(), X[k].imag()};
      for (int k = 0; k <= H / 2; ++k) {
        cplx zk, zhk;
        zfft::pack_pair(row[k], row[H - k], cplx{w[k].real(), w[k].imag()}, zk, zhk);
        row[k] = zk;
        if (k != 0 && k != H - k) row[H - k] = zhk;
      }
      std::vector<cplx> z(row.begin(), row.begin() + H);
      int p = 1, rem = H;
      while (rem > 1) {
        int R = zfft::next_radix(rem);
        if (R == 16) stage<16>(z, H, p, w);
        else if (R == 8) stage<8>(z, H, p, w);
        else if (R == 4) stage<4>(z, H, p, w);
        else stage<2>(z, H, p, w);
        p *= R;
        rem /= R;
      }
      double err = 0, mag = 0;
      for (int n = 0; n < M; ++n) {   // naive un-normalised inverse of the Hermitian-extended spectrum
        cd acc = X[0];
        for (int k = 1; k < H; ++k) acc += 2.0 * (X[k] * std::polar(1.0, 2 * M_PI * k * n / (double)M)).real();
        acc += X[H] * ((n & 1) ? -1.0 : 1.0);
        double got = (n & 1) ? z[n / 2].y : z[n / 2].x;
        err = std::fmax(err, std::fabs(got - acc.real()));
        mag = std::fmax(mag, std::fabs(acc.real()));
      }
      printf("M=%d Kz=%d rel_err=%.3e\n", M, Kz, err / mag);
      worst = std::fmax(worst, err / mag);
    }
  printf("WORST %.3e\n", worst);
  return worst < 1e-12 ? 0 : 1;
}
