// TEST INFRASTRUCTURE ONLY -- CPU oracle for the fp16 C2C FFT hot path.
//
// Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library, and only
// as the checker.  The product library (libtfft.so) never links or calls it.
//
// What is restated here (all file:line citations are relative to /root/reference):
//   * the transform DEFINITION the reference implements and is checked against:
//     X[k] = (1/N) * sum_n x[n] * exp(-2*pi*i*n*k/N), planar RE/IM, natural order in
//     and out (src/base/ComputeFFT.h:1-16; 1/N from the "sequential scaling" 1/256 in
//     src/base/TensorFFT256.cu:167-171, 1/16 in src/base/TensorRadix16.cu:133-136 and
//     1/2 in src/base/Radix2.cu:64-76; the reference's own check divides cuFFT Z2Z by
//     N, src/testing/AccuracyCalculator.h:70-84).  oracle_dft_f64 / oracle_fft_f64.
//   * the reference's staged ALGORITHM (digit reversal src/base/TensorFFT256.cu:125-161,
//     256-point base kernel :167-305, radix-16 combine src/base/TensorRadix16.cu:101-213,
//     radix-2 combine src/base/Radix2.cu:20-77, launch sequence
//     src/base/ComputeFFT.h:54-151) in two flavours: exact double arithmetic
//     (structure check: must equal the definition) and with every fp16 rounding of the
//     reference emulated (estimates the reference's own fp16 output).
//   * the reference's test fixture (src/testing/TestingDataCreation.h:15-27,89-117)
//     and error statistics (src/testing/AccuracyCalculator.h:86-148).
//
// Pinning: tests/test_oracle.py checks these functions against numpy's pocketfft, and
// tests/golden/ holds outputs of the real reference kernels (built from
// /root/reference by oracle/Makefile into oracle/_ref/, run on a B200).
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const double kPi = 3.14159265358979323846264338327950288;

inline double round_h(double v) {  // round-to-nearest-even to IEEE binary16 and back
  return static_cast<double>(static_cast<_Float16>(v));
}
inline double round_hf(float v) { return static_cast<double>(static_cast<_Float16>(v)); }

// exact twiddle exp(-2*pi*i*e/n) for integer e (phase reduced in integers first)
inline std::complex<double> tw(int64_t e, int64_t n) {
  e %= n;
  if (e < 0) e += n;
  // octant reduction keeps the argument small and the symmetric values exact
  if (e == 0) return {1.0, 0.0};
  if (4 * e == n) return {0.0, -1.0};
  if (2 * e == n) return {-1.0, 0.0};
  if (4 * e == 3 * n) return {0.0, 1.0};
  double a = -2.0 * kPi * static_cast<double>(e) / static_cast<double>(n);
  return {std::cos(a), std::sin(a)};
}

void fft_inplace(std::vector<std::complex<double>>& a) {  // iterative radix-2, unscaled
  const size_t n = a.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  std::vector<std::complex<double>> w(n / 2);
  for (size_t i = 0; i < n / 2; ++i) w[i] = tw(static_cast<int64_t>(i), static_cast<int64_t>(n));
  for (size_t len = 2; len <= n; len <<= 1) {
    const size_t step = n / len;
    for (size_t i = 0; i < n; i += len)
      for (size_t j = 0; j < len / 2; ++j) {
        std::complex<double> u = a[i + j], v = a[i + j + len / 2] * w[j * step];
        a[i + j] = u + v;
        a[i + j + len / 2] = u - v;
      }
  }
}

int ilog2(int64_t n) {
  int l = 0;
  while ((int64_t(1) << l) < n) ++l;
  return l;
}

// digit reversal of the reference: output index o -> input index
// (src/base/TensorFFT256.cu:125-161, identical to src/base/Transposer.cu:65-93)
int64_t ref_input_index(int64_t o, int r16_steps, int r2_steps) {
  int64_t t = o;
  int64_t in = 16 * (t % 16);
  t /= 16;
  in += t % 16;
  for (int i = 1; i < r16_steps; ++i) {
    t /= 16;
    in = 16 * in + (t % 16);
  }
  if (r2_steps > 0) {
    t /= 16;
    in = 2 * in + (t % 2);
    for (int i = 1; i < r2_steps; ++i) {
      t /= 2;
      in = 2 * in + (t % 2);
    }
  }
  return in;
}

// The reference's launch sequence for ONE transform in Mode_256 (ComputeFFT.h:54-151).
// emulate_fp16 = false: exact double arithmetic.  true: every value the reference
// rounds to __half is rounded here (loads, __hdiv, twiddle casts, __hmul/__hsub/__hfma,
// HMMA with fp16 accumulators modelled as an fp32 dot product rounded to fp16).
void ref_algorithm(const double* in_re, const double* in_im, double* out_re, double* out_im, int64_t n,
                   bool emulate_fp16) {
  auto R = [&](double v) { return emulate_fp16 ? round_h(v) : v; };
  const int lg = ilog2(n);
  const int r16_steps = lg / 4 - 1;  // Plan.h:99
  const int r2_steps = lg % 4;       // Plan.h:100
  // 16x16 DFT matrix as the kernels build it (TensorFFT256.cu:56-76): cosf/-sinf -> half
  double Fr[16][16], Fi[16][16];
  for (int j = 0; j < 16; ++j)
    for (int c = 0; c < 16; ++c) {
      if (emulate_fp16) {
        float ph = (static_cast<float>(j * c) * static_cast<float>(kPi)) / 8.0f;
        Fr[j][c] = round_hf(cosf(ph));
        Fi[j][c] = round_hf(-sinf(ph));
      } else {
        auto w = tw(j * c, 16);
        Fr[j][c] = w.real();
        Fi[j][c] = w.imag();
      }
    }
  auto cmatmul = [&](const double (*are)[16], const double (*aim)[16], double (*yre)[16], double (*yim)[16]) {
    // TensorFFT256.cu:191-215: RE = A_re*F_re - A_im*F_im (two fp16 accumulators, then
    // an fp16 subtraction); IM = A_im*F_re + A_re*F_im (one fp16 accumulator)
    for (int r = 0; r < 16; ++r)
      for (int c = 0; c < 16; ++c) {
        double s1 = 0, s2 = 0, s3 = 0, s4 = 0;
        for (int k = 0; k < 16; ++k) {
          s1 += are[r][k] * Fr[k][c];
          s2 += aim[r][k] * Fi[k][c];
          s3 += aim[r][k] * Fr[k][c];
          s4 += are[r][k] * Fi[k][c];
        }
        if (emulate_fp16) {
          double re1 = round_hf(static_cast<float>(s1)), re2 = round_hf(static_cast<float>(s2));
          double im1 = round_hf(static_cast<float>(s3));
          double im = round_hf(static_cast<float>(im1 + s4));
          yre[r][c] = round_h(re1 - re2);
          yim[r][c] = im;
        } else {
          yre[r][c] = s1 - s2;
          yim[r][c] = s3 + s4;
        }
      }
  };
  auto cmul_tw = [&](double xr, double xi, double wr, double wi, double& yr, double& yi) {
    // __hsub(__hmul(re,wr), __hmul(im,wi)) ; __hfma(re, wi, __hmul(im, wr))
    if (emulate_fp16) {
      yr = round_h(round_h(xr * wr) - round_h(xi * wi));
      yi = round_h(xr * wi + round_h(xi * wr));
    } else {
      yr = xr * wr - xi * wi;
      yi = xr * wi + xi * wr;
    }
  };
  std::vector<double> a_re(n), a_im(n), b_re(n), b_im(n);
  // ---- base kernel: one "warp" per 256 outputs (TensorFFT256.cu:20-306)
  for (int64_t w = 0; w < n / 256; ++w) {
    double A_re[16][16], A_im[16][16], Y_re[16][16], Y_im[16][16], B_re[16][16], B_im[16][16];
    for (int j = 0; j < 16; ++j)
      for (int c = 0; c < 16; ++c) {
        int64_t o = c + 16 * j + 256 * w;
        int64_t idx = ref_input_index(o, r16_steps, r2_steps);
        A_re[j][c] = R(R(in_re[idx]) / 256.0);  // __hdiv(x, 256)
        A_im[j][c] = R(R(in_im[idx]) / 256.0);
      }
    cmatmul(A_re, A_im, Y_re, Y_im);
    for (int j = 0; j < 16; ++j)
      for (int c = 0; c < 16; ++c) {  // twiddle w_256^(c*j), written transposed (:225-254)
        double wr, wi;
        if (emulate_fp16) {
          float ph = (static_cast<float>(c * j) * static_cast<float>(kPi)) / 128.0f;
          wr = round_hf(cosf(ph));
          wi = round_hf(-sinf(ph));
        } else {
          auto t = tw(c * j, 256);
          wr = t.real();
          wi = t.imag();
        }
        cmul_tw(Y_re[j][c], Y_im[j][c], wr, wi, B_re[c][j], B_im[c][j]);
      }
    cmatmul(B_re, B_im, Y_re, Y_im);
    for (int j = 0; j < 16; ++j)
      for (int c = 0; c < 16; ++c) {  // store reverts the transpose (:295-305)
        a_re[c + 16 * j + 256 * w] = Y_re[c][j];
        a_im[c + 16 * j + 256 * w] = Y_im[c][j];
      }
  }
  // ---- radix-16 combine launches (ComputeFFT.h:105-120, TensorRadix16.cu:101-213)
  int64_t L = 256;
  for (int s = 1; s < r16_steps; ++s) {
    const int64_t comb = 16 * L;
    for (int64_t sub = 0; sub < n / comb; ++sub)
      for (int64_t i = 0; i < L; ++i) {
        double xr[16], xi[16];
        for (int j = 0; j < 16; ++j) {
          int64_t g = i + L * j + sub * comb;
          double vr = R(a_re[g] / 16.0), vi = R(a_im[g] / 16.0);
          double wr, wi;
          if (emulate_fp16) {
            float tmp = static_cast<float>(i * j) / static_cast<float>(comb);
            float ph = 2.0f * static_cast<float>(kPi) * tmp;
            wr = round_hf(cosf(ph));
            wi = round_hf(-sinf(ph));
          } else {
            auto t = tw(i * j, comb);
            wr = t.real();
            wi = t.imag();
          }
          cmul_tw(vr, vi, wr, wi, xr[j], xi[j]);
        }
        for (int k = 0; k < 16; ++k) {
          double s1 = 0, s2 = 0, s3 = 0, s4 = 0;
          for (int j = 0; j < 16; ++j) {
            s1 += xr[j] * Fr[j][k];
            s2 += xi[j] * Fi[j][k];
            s3 += xi[j] * Fr[j][k];
            s4 += xr[j] * Fi[j][k];
          }
          int64_t g = i + L * k + sub * comb;
          if (emulate_fp16) {
            double re1 = round_hf(static_cast<float>(s1)), re2 = round_hf(static_cast<float>(s2));
            double im1 = round_hf(static_cast<float>(s3));
            b_re[g] = round_h(re1 - re2);
            b_im[g] = round_hf(static_cast<float>(im1 + s4));
          } else {
            b_re[g] = s1 - s2;
            b_im[g] = s3 + s4;
          }
        }
      }
    a_re.swap(b_re);
    a_im.swap(b_im);
    L = comb;
  }
  // ---- radix-2 steps (ComputeFFT.h:123-145, Radix2.cu:20-77)
  for (int s = 0; s < r2_steps; ++s) {
    for (int64_t sub = 0; sub < n / (2 * L); ++sub)
      for (int64_t t = 0; t < L; ++t) {
        int64_t p1 = sub * 2 * L + t, p2 = p1 + L;
        double wr, wi;
        if (emulate_fp16) {
          float ph = static_cast<float>(kPi) * (static_cast<float>(t) / static_cast<float>(L));
          wr = round_hf(cosf(ph));
          wi = round_hf(-sinf(ph));
        } else {
          auto tt = tw(t, 2 * L);
          wr = tt.real();
          wi = tt.imag();
        }
        double mr, mi;
        cmul_tw(a_re[p2], a_im[p2], wr, wi, mr, mi);
        b_re[p1] = R(R(a_re[p1] + mr) * 0.5);
        b_im[p1] = R(R(a_im[p1] + mi) * 0.5);
        b_re[p2] = R(R(a_re[p1] - mr) * 0.5);
        b_im[p2] = R(R(a_im[p1] - mi) * 0.5);
      }
    a_re.swap(b_re);
    a_im.swap(b_im);
    L *= 2;
  }
  std::memcpy(out_re, a_re.data(), n * sizeof(double));
  std::memcpy(out_im, a_im.data(), n * sizeof(double));
}

}  // namespace

extern "C" {

// libstdc++ std::default_random_engine seeded through std::seed_seq{seed},
// uniform_real_distribution<float>(-1,1)  (src/testing/TestingDataCreation.h:15-27)
void oracle_random_weights(int count, int seed, float* out) {
  std::seed_seq sq = {seed};
  std::default_random_engine gen(sq);
  std::uniform_real_distribution<float> d(-1.0, 1.0);
  for (int i = 0; i < count; ++i) out[i] = d(gen);
}

// x[t] = sum_{i<cutoff} w[i]*sinf(2*pi*i*t/N), accumulated in double
// (src/testing/TestingDataCreation.h:89-117).  Output is NOT yet rounded to fp16.
void oracle_sine_fixture(int64_t n, int cutoff, const float* w_re, const float* w_im, double* re, double* im) {
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < n; ++t) {
    double a = 0, b = 0;
    for (int i = 0; i < cutoff; ++i) {
      float s = sinf(static_cast<float>((2 * kPi * i * t) / static_cast<double>(n)));
      a += w_re[i] * s;
      b += w_im[i] * s;
    }
    re[t] = a;
    im[t] = b;
  }
}

void oracle_round_to_half(const double* in, double* out, int64_t count) {
  for (int64_t i = 0; i < count; ++i) out[i] = round_h(in[i]);
}

// naive O(N^2) fp64 DFT with the reference's 1/N scale; batch transforms are
// contiguous planes of n doubles.  nthreads <= 0: all cores.
int oracle_dft_f64(const double* in_re, const double* in_im, double* out_re, double* out_im, int64_t n,
                   int64_t batch, int nthreads) {
  if (n <= 0 || batch <= 0) return -1;
  std::vector<double> cr(n), ci(n);
  for (int64_t e = 0; e < n; ++e) {
    auto w = tw(e, n);
    cr[e] = w.real();
    ci[e] = w.imag();
  }
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static) collapse(2)
  for (int64_t b = 0; b < batch; ++b)
    for (int64_t k = 0; k < n; ++k) {
      const double* xr = in_re + b * n;
      const double* xi = in_im + b * n;
      double sr = 0, si = 0;
      int64_t e = 0;
      for (int64_t j = 0; j < n; ++j) {
        sr += xr[j] * cr[e] - xi[j] * ci[e];
        si += xr[j] * ci[e] + xi[j] * cr[e];
        e += k;
        if (e >= n) e -= n;
      }
      out_re[b * n + k] = sr / static_cast<double>(n);
      out_im[b * n + k] = si / static_cast<double>(n);
    }
  return 0;
}

// O(N log N) fp64 FFT (power-of-two n), same convention as oracle_dft_f64.
int oracle_fft_f64(const double* in_re, const double* in_im, double* out_re, double* out_im, int64_t n,
                   int64_t batch, int nthreads) {
  if (n <= 0 || (n & (n - 1)) || batch <= 0) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    std::vector<std::complex<double>> a(n);
#pragma omp for schedule(static)
    for (int64_t b = 0; b < batch; ++b) {
      for (int64_t i = 0; i < n; ++i) a[i] = {in_re[b * n + i], in_im[b * n + i]};
      fft_inplace(a);
      for (int64_t i = 0; i < n; ++i) {
        out_re[b * n + i] = a[i].real() / static_cast<double>(n);
        out_im[b * n + i] = a[i].imag() / static_cast<double>(n);
      }
    }
  }
  return 0;
}

// 2-D transform (rows then columns), scale 1/(ny*nx): BASELINE config 5
int oracle_fft2_f64(const double* in_re, const double* in_im, double* out_re, double* out_im, int64_t ny,
                    int64_t nx, int64_t batch, int nthreads) {
  if ((nx & (nx - 1)) || (ny & (ny - 1))) return -1;
  std::vector<double> tr(ny * nx), ti(ny * nx), ur(ny * nx), ui(ny * nx);
  for (int64_t b = 0; b < batch; ++b) {
    const int64_t off = b * ny * nx;
    oracle_fft_f64(in_re + off, in_im + off, tr.data(), ti.data(), nx, ny, nthreads);
    for (int64_t y = 0; y < ny; ++y)
      for (int64_t x = 0; x < nx; ++x) {
        ur[x * ny + y] = tr[y * nx + x];
        ui[x * ny + y] = ti[y * nx + x];
      }
    oracle_fft_f64(ur.data(), ui.data(), tr.data(), ti.data(), ny, nx, nthreads);
    for (int64_t y = 0; y < ny; ++y)
      for (int64_t x = 0; x < nx; ++x) {
        out_re[off + y * nx + x] = tr[x * ny + y];
        out_im[off + y * nx + x] = ti[x * ny + y];
      }
  }
  return 0;
}

// The reference's own algorithm (Mode_256 default plan), one transform of length n >= 256.
int oracle_ref_algorithm(const double* in_re, const double* in_im, double* out_re, double* out_im, int64_t n,
                         int emulate_fp16) {
  if (n < 256 || (n & (n - 1))) return -1;
  ref_algorithm(in_re, in_im, out_re, out_im, n, emulate_fp16 != 0);
  return 0;
}

int64_t oracle_ref_input_index(int64_t o, int64_t n) {
  const int lg = ilog2(n);
  return ref_input_index(o, lg / 4 - 1, lg % 4);
}

// Error statistics.  stats[0] = relative L2 = ||a-b||_2/||b||_2 over the 2*count reals
// (BASELINE metric); stats[1..3] = the reference's triple max|a-b|, mean|a-b|, sample
// sigma of |a-b| about that mean (src/testing/AccuracyCalculator.h:86-148).
void oracle_error_stats(const double* a_re, const double* a_im, const double* b_re, const double* b_im,
                        int64_t count, double* stats) {
  long double num = 0, den = 0, sum = 0;
  double mx = 0;
  for (int64_t i = 0; i < count; ++i) {
    double dr = a_re[i] - b_re[i], di = a_im[i] - b_im[i];
    num += static_cast<long double>(dr) * dr + static_cast<long double>(di) * di;
    den += static_cast<long double>(b_re[i]) * b_re[i] + static_cast<long double>(b_im[i]) * b_im[i];
    sum += std::fabs(dr) + std::fabs(di);
    mx = std::fmax(mx, std::fmax(std::fabs(dr), std::fabs(di)));
  }
  const double avg = static_cast<double>(sum / (2.0L * count));
  long double var = 0;
  for (int64_t i = 0; i < count; ++i) {
    double dr = std::fabs(a_re[i] - b_re[i]) - avg, di = std::fabs(a_im[i] - b_im[i]) - avg;
    var += static_cast<long double>(dr) * dr + static_cast<long double>(di) * di;
  }
  stats[0] = den > 0 ? std::sqrt(static_cast<double>(num / den)) : 0.0;
  stats[1] = mx;
  stats[2] = avg;
  stats[3] = std::sqrt(static_cast<double>(var / (2.0L * count - 1.0L)));
}

int oracle_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
