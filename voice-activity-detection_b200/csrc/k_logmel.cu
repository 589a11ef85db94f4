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
// This is HBM-/latency-trivial CUDA-core work (640 B in, 320 B out per frame, ~25 kFLOP): one CTA of 128
// threads per frame, radix-2 decimation-in-time FFT in shared memory (n_fft = 512: 9 stages of 256
// butterflies), twiddles and window from host-built float64 tables, the mel filterbank as per-filter
// contiguous bin ranges (each Slaney triangle is a contiguous run of bins).  It exists so that the
// reference predictor's whole audio -> probabilities path stays on the device: the NumPy front end costs
// 1.9 s for 10 minutes of audio, the model 3 ms.
#include <math.h>

#include <vector>

#include "vadb_common.cuh"

namespace vadb {
namespace {

constexpr int LM_THREADS = 128;

__global__ void __launch_bounds__(LM_THREADS)
logmel_kernel(const float* __restrict__ audio, long n_samples, int n_fft, int log2n, int hop, long n_frames,
              const double* __restrict__ window,      // [n_fft] hann(win) zero-padded to n_fft
              const double2* __restrict__ twiddle,    // [n_fft/2] exp(-2 pi i k / n_fft)
              const int* __restrict__ fb_start,       // [n_mels] first bin of each filter
              const int* __restrict__ fb_len,         // [n_mels] number of bins
              const int* __restrict__ fb_off,         // [n_mels] offset into fb_w
              const float* __restrict__ fb_w,         // packed filter weights
              int n_mels, float* __restrict__ out /* [n_frames, n_mels] */) {
  extern __shared__ __align__(16) unsigned char lm_smem[];
  double2* data = reinterpret_cast<double2*>(lm_smem);                    // [n_fft]
  double2* tw = data + n_fft;                                             // [n_fft/2]
  float* pw = reinterpret_cast<float*>(tw + n_fft / 2);                   // [n_fft/2 + 1]
  const int tid = threadIdx.x;
  const int half_n = n_fft >> 1;
  for (int i = tid; i < half_n; i += LM_THREADS) tw[i] = twiddle[i];
  pdl_launch_dependents();
  pdl_wait();
  for (long t = blockIdx.x; t < n_frames; t += gridDim.x) {
    __syncthreads();                                   // previous frame's power / data fully consumed
    // windowed samples of the centred, reflect-padded frame, stored in bit-reversed order
    for (int i = tid; i < n_fft; i += LM_THREADS) {
      long m = t * hop + i - half_n;
      if (m < 0) m = -m;                               // np.pad(mode="reflect"): the edge sample is not repeated
      if (m >= n_samples) m = 2 * (n_samples - 1) - m;
      m = m < 0 ? 0 : (m >= n_samples ? n_samples - 1 : m);   // clips shorter than the padding
      const double x = (double)audio[m] * window[i];
      const unsigned r = __brev((unsigned)i) >> (32 - log2n);
      data[r] = make_double2(x, 0.0);
    }
    __syncthreads();
    for (int s = 1; s <= log2n; ++s) {
      const int hs = 1 << (s - 1);
      const int tstride = half_n >> (s - 1);           // twiddle index step: n_fft / (2 hs) / ... = (n/2)/hs
      for (int b = tid; b < half_n; b += LM_THREADS) {
        const int j = b & (hs - 1);
        const int i0 = ((b >> (s - 1)) << s) + j;
        const int i1 = i0 + hs;
        const double2 w = tw[j * tstride];
        const double2 u = data[i0], v = data[i1];
        const double vr = v.x * w.x - v.y * w.y, vi = v.x * w.y + v.y * w.x;
        data[i0] = make_double2(u.x + vr, u.y + vi);
        data[i1] = make_double2(u.x - vr, u.y - vi);
      }
      __syncthreads();
    }
    // power spectrum with the reference's roundings: complex64 spectrum, float32 |X|, squared
    for (int k = tid; k <= half_n; k += LM_THREADS) {
      const float re = (float)data[k].x, im = (float)data[k].y;
      const float mag = hypotf(re, im);
      pw[k] = mag * mag;
    }
    __syncthreads();
    for (int m = tid; m < n_mels; m += LM_THREADS) {
      const int s0 = fb_start[m], len = fb_len[m];
      const float* w = fb_w + fb_off[m];
      float acc = 0.f;
      for (int k = 0; k < len; ++k) acc = fmaf(w[k], pw[s0 + k], acc);
      out[t * n_mels + m] = logf(acc + 1e-6f);
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

size_t logmel_smem_bytes(int n_fft) {
  return (size_t)n_fft * sizeof(double2) + (size_t)(n_fft / 2) * sizeof(double2) + (size_t)(n_fft / 2 + 1) * sizeof(float);
}

cudaError_t launch_logmel(const float* audio, long n_samples, int n_fft, int hop, long n_frames, const double* window,
                          const double* twiddle, const int* fb_start, const int* fb_len, const int* fb_off,
                          const float* fb_w, int n_mels, float* out, cudaStream_t s) {
  if (n_frames <= 0) return cudaSuccess;
  int log2n = 0;
  while ((1 << log2n) < n_fft) ++log2n;
  const size_t smem = logmel_smem_bytes(n_fft);
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_dev != dev && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_dev = dev;
  }
  long blocks = n_frames < 148L * 16 ? n_frames : 148L * 16;     // a multiple of the SM count; CTAs stride over frames
  return launch_k(logmel_kernel, (unsigned)blocks, LM_THREADS, smem, s, audio, n_samples, n_fft, log2n, hop, n_frames,
                  window, reinterpret_cast<const double2*>(twiddle), fb_start, fb_len, fb_off, fb_w, n_mels, out);
}

}  // namespace vadb
