// Log-mel front end on the device: the step UPSTREAM of the model.
//
// Reference: vad/acoustics/transforms/log_mel_spectrogram.py:19-32,
//   np.log(librosa.feature.melspectrogram(y, sr, n_mels, n_fft, hop_length, win_length) + 1e-6)
// transposed to [frames, n_mels] by vad/acoustics/feature_extractor.py:77-80, with librosa 0.8.0's
// defaults: periodic hann window of win_length zero-padded (centred) to n_fft, center=True with reflect
// padding of n_fft/2, rfft, power = |X|^2, Slaney mel scale with Slaney area normalisation, fmin = 0,
// fmax = sr/2.  librosa is neither vendored by the reference nor installed here: PARITY UNPINNED against
// librosa itself; the kernel is pinned to the NumPy restatement of that published algorithm
// (oracle/logmel_oracle.py) and reproduces its precision choices: the samples are windowed and
// transformed in float64 (numpy.fft computes in double), the spectrum is rounded to complex64, |X| is a
// float32 hypot, the filterbank product and the log are float32.
//
// CUDA-core work (640 B in, 320 B out per frame, ~13 k float64 operations): several frames per CTA, 64 threads per
// frame.  The n_fft real samples of a frame are transformed as ONE complex FFT of n_fft/2 points (z[n] = x[2n] +
// i x[2n+1]) followed by the split step of the real-input transform -- half the butterflies of the complex
// transform of round 1 -- in a radix-4 Stockham (autosort) formulation: log4(n_fft/2) stages that ping-pong between
// two shared-memory buffers (plus one radix-2 stage when n_fft/2 is not a power of four), one __syncthreads per
// stage, no bit reversal.  Twiddles and window come from host-built float64 tables staged in shared memory, the mel
// filterbank is applied as per-filter contiguous bin ranges (each Slaney triangle is a contiguous run of bins).
// Round 1's kernel (one CTA per frame, radix-2, 11 barriers per frame) took 780 us for 10 minutes of audio -- 55 % of
// the whole window-path model; the NumPy front end costs 0.3-1.9 s.
#include <math.h>

#include <vector>

#include "vadb_common.cuh"

namespace vadb {
namespace {

constexpr int LM_TPF = 64;          // threads per frame
// shared-memory index of complex point i: one 16-byte slot of padding per 8 points, so that the stride-4 / stride-16
// stores of the early Stockham stages spread over the banks (unpadded: 16-way conflicts on the first stage's stores)
#define LM_IDX(i) ((i) + ((i) >> 3))
__host__ __device__ constexpr size_t lm_buf_points(size_t N) { return N + (N >> 3) + 1; }

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// exp(-2 pi i m / n_fft) for m in [0, n_fft) from the half table tw[k] = exp(-2 pi i k / n_fft), k < n_fft / 2
__device__ __forceinline__ double2 twid(const double2* tw, int m, int half_n) {
  const double2 w = tw[m & (half_n - 1)];
  return (m & half_n) ? make_double2(-w.x, -w.y) : w;
}

__global__ void __launch_bounds__(256)
logmel_kernel(const float* __restrict__ audio, long n_samples, int n_fft, int fpc, int hop, long n_frames,
              const double* __restrict__ window,      // [n_fft] hann(win) zero-padded to n_fft
              const double2* __restrict__ twiddle,    // [n_fft/2] exp(-2 pi i k / n_fft)
              const int* __restrict__ fb_start,       // [n_mels] first bin of each filter
              const int* __restrict__ fb_len,         // [n_mels] number of bins
              const int* __restrict__ fb_off,         // [n_mels] offset into fb_w
              const float* __restrict__ fb_w,         // packed filter weights
              int n_mels, float* __restrict__ out /* [n_frames, n_mels] */) {
  extern __shared__ __align__(16) unsigned char lm_smem[];
  const int half_n = n_fft >> 1;                       // N: points of the complex transform
  double2* tw = reinterpret_cast<double2*>(lm_smem);                      // [N]
  double* win = reinterpret_cast<double*>(tw + half_n);                   // [n_fft]
  double2* buf = reinterpret_cast<double2*>(win + n_fft);                 // [fpc][2][N padded]
  const int bufp = (int)lm_buf_points((size_t)half_n);
  float* pw_all = reinterpret_cast<float*>(buf + (size_t)fpc * 2 * bufp);     // [fpc][N + 1]
  const int tid = threadIdx.x;
  for (int i = tid; i < half_n; i += blockDim.x) tw[i] = twiddle[i];
  for (int i = tid; i < n_fft; i += blockDim.x) win[i] = window[i];
  pdl_launch_dependents();
  pdl_wait();
  const int f = tid / LM_TPF, lt = tid % LM_TPF;       // frame slot of this thread, thread index inside the frame
  double2* b0 = buf + (size_t)f * 2 * bufp;
  double2* b1 = b0 + bufp;
  float* pw = pw_all + (size_t)f * (half_n + 1);
  for (long t0 = (long)blockIdx.x * fpc; t0 < n_frames; t0 += (long)gridDim.x * fpc) {
    const long t = t0 + f;
    const bool live = f < fpc && t < n_frames;
    __syncthreads();                                   // previous frames' buffers fully consumed (and the tables staged)
    if (live) {
      // z[n] = windowed samples 2n, 2n+1 of the centred, reflect-padded frame
      for (int n = lt; n < half_n; n += LM_TPF) {
        double v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = 2 * n + e;
          long m = t * hop + i - half_n;
          if (m < 0) m = -m;                           // np.pad(mode="reflect"): the edge sample is not repeated
          if (m >= n_samples) m = 2 * (n_samples - 1) - m;
          m = m < 0 ? 0 : (m >= n_samples ? n_samples - 1 : m);   // clips shorter than the padding
          v[e] = (double)audio[m] * win[i];
        }
        b0[LM_IDX(n)] = make_double2(v[0], v[1]);
      }
    }
    // Stockham autosort FFT of N points: stage with sub-transform length Ns reads src[j + r N/R], multiplies by
    // exp(-2 pi i r (j mod Ns) / (Ns R)), applies the radix-R butterfly and writes dst[(j div Ns) Ns R + (j mod Ns) + r Ns]
    double2* src = b0;
    double2* dst = b1;
    for (int Ns = 1; Ns < half_n;) {
      __syncthreads();
      const int R = (half_n / Ns) >= 4 ? 4 : 2;
      if (live) {
        const int nb = half_n / R;                     // butterflies of this stage
        for (int j = lt; j < nb; j += LM_TPF) {
          const int k = j & (Ns - 1);
          const int tstep = n_fft / (Ns * R);          // exp(-2 pi i r k / (Ns R)) = table[r k n_fft / (Ns R)]
          const int j0 = ((j - k) * R) + k;
          if (R == 4) {
            const double2 a0 = src[LM_IDX(j)];
            const double2 a1 = cmul(src[LM_IDX(j + nb)], twid(tw, k * tstep, half_n));
            const double2 a2 = cmul(src[LM_IDX(j + 2 * nb)], twid(tw, 2 * k * tstep, half_n));
            const double2 a3 = cmul(src[LM_IDX(j + 3 * nb)], twid(tw, 3 * k * tstep, half_n));
            const double2 s02 = make_double2(a0.x + a2.x, a0.y + a2.y), d02 = make_double2(a0.x - a2.x, a0.y - a2.y);
            const double2 s13 = make_double2(a1.x + a3.x, a1.y + a3.y), d13 = make_double2(a1.x - a3.x, a1.y - a3.y);
            dst[LM_IDX(j0)] = make_double2(s02.x + s13.x, s02.y + s13.y);
            dst[LM_IDX(j0 + Ns)] = make_double2(d02.x + d13.y, d02.y - d13.x);            // d02 - i d13
            dst[LM_IDX(j0 + 2 * Ns)] = make_double2(s02.x - s13.x, s02.y - s13.y);
            dst[LM_IDX(j0 + 3 * Ns)] = make_double2(d02.x - d13.y, d02.y + d13.x);        // d02 + i d13
          } else {
            const double2 a0 = src[LM_IDX(j)];
            const double2 a1 = cmul(src[LM_IDX(j + nb)], twid(tw, k * tstep, half_n));
            dst[LM_IDX(j0)] = make_double2(a0.x + a1.x, a0.y + a1.y);
            dst[LM_IDX(j0 + Ns)] = make_double2(a0.x - a1.x, a0.y - a1.y);
          }
        }
      }
      Ns *= R;
      double2* tmp = src; src = dst; dst = tmp;
    }
    __syncthreads();
    if (live) {
      // split step of the real-input transform, X[k] = (Z[k] + conj Z[N-k]) / 2 - i W^k (Z[k] - conj Z[N-k]) / 2, and the
      // power spectrum with the reference's roundings: complex64 spectrum, float32 |X|, squared
      for (int k = lt; k <= half_n; k += LM_TPF) {
        const double2 zk = src[LM_IDX(k & (half_n - 1))], zn = src[LM_IDX((half_n - k) & (half_n - 1))];
        const double2 e = make_double2(0.5 * (zk.x + zn.x), 0.5 * (zk.y - zn.y));      // even part
        const double2 o = make_double2(0.5 * (zk.x - zn.x), 0.5 * (zk.y + zn.y));      // (Z[k] - conj Z[N-k]) / 2
        const double2 w = twid(tw, k, half_n);                                          // exp(-2 pi i k / n_fft), k <= N
        const double2 wo = cmul(w, o);
        const float re = (float)(e.x + wo.y), im = (float)(e.y - wo.x);                // e - i (w o)
        const float mag = hypotf(re, im);
        pw[k] = mag * mag;
      }
    }
    __syncthreads();
    if (live) {
      for (int m = lt; m < n_mels; m += LM_TPF) {
        const int s0 = fb_start[m], len = fb_len[m];
        const float* w = fb_w + fb_off[m];
        float acc = 0.f;
        for (int k = 0; k < len; ++k) acc = fmaf(w[k], pw[s0 + k], acc);
        out[t * n_mels + m] = logf(acc + 1e-6f);
      }
    }
  }
}

// librosa.filters.mel's Slaney scale (htk=False)
double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}
// numpy.linspace(start, stop, num): start + i * step, last element exactly stop
std::vector<double> linspace(double start, double stop, int num) {
  std::vector<double> y(num);
  const double step = num > 1 ? (stop - start) / (num - 1) : 0.0;
  for (int i = 0; i < num; ++i) y[i] = start + i * step;
  if (num > 1) y[num - 1] = stop;
  return y;
}

}  // namespace

// Dense [n_mels, n_fft/2 + 1] float32 filterbank = librosa.filters.mel(sr, n_fft, n_mels, fmin=0,
// fmax=sr/2, htk=False, norm='slaney'), and the [n_fft] float64 window (periodic hann of `win` samples,
// zero-padded and centred).  Host-only; exported through the C ABI for the CPU tests.
void logmel_tables(int sr, int n_fft, int win, int n_mels, std::vector<float>* fb, std::vector<double>* window) {
  const int n_bins = n_fft / 2 + 1;
  const std::vector<double> fftfreqs = linspace(0.0, sr / 2.0, n_bins);
  std::vector<double> mel_f = linspace(hz_to_mel(0.0), hz_to_mel(sr / 2.0), n_mels + 2);
  for (double& m : mel_f) m = mel_to_hz(m);
  fb->assign((size_t)n_mels * n_bins, 0.f);
  for (int m = 0; m < n_mels; ++m) {
    const double fd0 = mel_f[m + 1] - mel_f[m], fd1 = mel_f[m + 2] - mel_f[m + 1];
    const double enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
    for (int k = 0; k < n_bins; ++k) {
      const double lower = -(mel_f[m] - fftfreqs[k]) / fd0;
      const double upper = (mel_f[m + 2] - fftfreqs[k]) / fd1;
      const double w = fmax(0.0, fmin(lower, upper));
      (*fb)[(size_t)m * n_bins + k] = (float)(w * enorm);
    }
  }
  window->assign(n_fft, 0.0);
  const int lpad = (n_fft - win) / 2;
  const double two_pi = 6.283185307179586476925286766559;
  for (int i = 0; i < win; ++i) (*window)[lpad + i] = 0.5 - 0.5 * cos(two_pi * i / win);   // scipy get_window("hann", win, fftbins=True)
}

size_t logmel_smem_bytes(int n_fft, int fpc) {
  const size_t N = (size_t)n_fft / 2;
  return N * sizeof(double2) + (size_t)n_fft * sizeof(double) + (size_t)fpc * 2 * lm_buf_points(N) * sizeof(double2) +
         (size_t)fpc * (N + 1) * sizeof(float);
}

cudaError_t launch_logmel(const float* audio, long n_samples, int n_fft, int hop, long n_frames, const double* window,
                          const double* twiddle, const int* fb_start, const int* fb_len, const int* fb_off,
                          const float* fb_w, int n_mels, float* out, cudaStream_t s) {
  if (n_frames <= 0) return cudaSuccess;
  int fpc = 4;                                       // frames per CTA (64 threads each)
  while (fpc > 1 && logmel_smem_bytes(n_fft, fpc) > 160 * 1024) fpc >>= 1;
  const size_t smem = logmel_smem_bytes(n_fft, fpc);
  static thread_local int attr_dev = -1;
  static thread_local size_t attr_smem = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if ((attr_dev != dev || smem > attr_smem) && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_dev = dev; attr_smem = smem;
  }
  const long groups = (n_frames + fpc - 1) / fpc;
  long blocks = groups < 148L * 8 ? groups : 148L * 8;     // a multiple of the SM count; CTAs stride over frame groups
  return launch_k(logmel_kernel, (unsigned)blocks, fpc * LM_TPF, smem, s, audio, n_samples, n_fft, fpc, hop, n_frames,
                  window, reinterpret_cast<const double2*>(twiddle), fb_start, fb_len, fb_off, fb_w, n_mels, out);
}

}  // namespace vadb
