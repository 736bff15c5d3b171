// The two speaker encoders of the prompt path, written against a small backend interface:
//
//   style  : 16 kHz wave -> kaldi fbank (80 bins) -> minus time mean -> CAMPPlus -> style vector [192]
//            (`InferenceWrapper.calculate_style_vec`, evaluations/infer_arvc.py:179-211; modules/campplus/DTDNN.py,
//             modules/campplus/layers.py)
//   timbre : 16 kHz wave -> slaney mel magnitudes (128 bins, hop 320) -> ECAPA-TDNN trunk -> PerceiverResampler
//            (32 latents) -> FSQ 4^6 -> timbre latents [32][128]
//            (`InferenceWrapper.calculate_timbre_latent`, evaluations/infer_arvc.py:213-223; `SpeakerEncoder.tokenize_wav`,
//             modules/bicodec_speaker_encoder/speaker_encoder.py:136-144)
//
// Every dense layer with enough rows is a GEMM of the engine (GemmParams: 1x1 convs, k-tap dilated convs as row-offset
// taps, the stride-2 TDNN as overlapping rows with a_row_step = 2); everything else is an element-parallel functor: one
// thread computes one output element from global memory, no shared memory, no cooperation.  That keeps this SETUP path
// (it runs once per prompt, 0.3-0.5 GFLOP per second of reference audio) simple enough that the SAME source is compiled
//   * by nvcc into libsvanon_b200.so (speaker.cu: functors run as `pfor_kernel` launches, GEMMs are launch_gemm), and
//   * by g++ into a test-only host build (tests/hostemu/: functors run as loops, the GEMM is a three-loop reference)
// so the host orchestration (buffer shapes, margins, weight repacking, GEMM descriptors) is checked against the
// reference-generated fixtures on a machine without a GPU.  The host build is never loaded by the product.
#pragma once
#include <cmath>
#include <string>
#include <vector>

#include "common.cuh"

#if defined(__CUDACC__)
#define SV_HD __host__ __device__ __forceinline__
#else
#define SV_HD inline
#endif

namespace svanon {
namespace spk {

// ------------------------------------------------------------------------------------------------ functors
struct Fill {
  float* dst;
  float v;
  SV_HD void operator()(long long i) const { dst[i] = v; }
};

// mean of each 25 ms frame (kaldi `remove_dc_offset`); one thread per frame
struct FrameMean {
  const float* x;
  int win, shift;
  float* mean;
  SV_HD void operator()(long long f) const {
    const float* p = x + f * shift;
    double s = 0;
    for (int j = 0; j < win; ++j) s += p[j];
    mean[f] = (float)(s / win);
  }
};

// kaldi frame: (x - mean) - 0.97 * (previous sample - mean; the first sample against itself), povey window, zero-padded
// to `padded` samples.  One thread per (frame, sample).
struct KaldiFrame {
  const float* x;
  const float* mean;
  const float* window;   // [win]
  int win, shift, padded;
  float* out;            // [frames][padded]
  SV_HD void operator()(long long i) const {
    const long long f = i / padded;
    const int j = (int)(i - f * padded);
    float v = 0.f;
    if (j < win) {
      const float* p = x + f * shift;
      const float m = mean[f];
      const float a = p[j] - m;
      const float b = p[j > 0 ? j - 1 : 0] - m;
      v = (a - 0.97f * b) * window[j];
    }
    out[i] = v;
  }
};

// centred STFT frame of torchaudio's MelSpectrogram (center=True, reflect padding): sample f*hop + j - nfft/2,
// reflected at both ends, times the window (the 640-sample hann window centred in the 1024 frame).
struct ReflectFrame {
  const float* x;
  long long n;
  const float* window;   // [nfft]
  int nfft, hop;
  float* out;            // [frames][nfft]
  SV_HD void operator()(long long i) const {
    const long long f = i / nfft;
    const int j = (int)(i - f * nfft);
    long long idx = f * hop + j - nfft / 2;
    if (idx < 0) idx = -idx;
    if (idx >= n) idx = 2 * (n - 1) - idx;
    out[i] = x[idx] * window[j];
  }
};

// one DFT bin of one frame, fp64 accumulation against a [nfft] (cos, sin) table; power or magnitude
struct DftBin {
  const float* frames;   // [F][nfft]
  const double* tw;      // [nfft][2]
  int nfft, j0, j1, nbins, power;
  float* out;            // [F][nbins]
  SV_HD void operator()(long long i) const {
    const long long f = i / nbins;
    const int k = (int)(i - f * nbins);
    const float* fr = frames + f * nfft;
    double re = 0, im = 0;
    for (int j = j0; j < j1; ++j) {
      const int idx = (int)(((long long)j * k) & (nfft - 1));
      const double v = fr[j];
      re += v * tw[2 * idx];
      im -= v * tw[2 * idx + 1];
    }
    const double p = re * re + im * im;
    out[i] = power ? (float)p : (float)sqrt(p);
  }
};

// mel[f][m] = sum_k spec[f][k] * fb[k * fb_sk + m * fb_sm]  (optionally log(max(., eps))), written at out[f*o_sf + m*o_sm]
struct MelDot {
  const float* spec;
  int nbins, nmel;
  const float* fb;
  long long fb_sk, fb_sm;
  int do_log;
  float eps;
  float* out;
  long long o_sf, o_sm;
  SV_HD void operator()(long long i) const {
    const long long f = i / nmel;
    const int m = (int)(i - f * nmel);
    const float* s = spec + f * nbins;
    double acc = 0;
    for (int k = 0; k < nbins; ++k) acc += (double)s[k] * (double)fb[k * fb_sk + m * fb_sm];
    float v = (float)acc;
    if (do_log) v = logf(fmaxf(v, eps));
    out[f * o_sf + m * o_sm] = v;
  }
};

// mean over the T columns of row r of a [R][T] matrix
struct RowMean {
  const float* x;
  long long T;
  float* mean;
  SV_HD void operator()(long long r) const {
    const float* p = x + r * T;
    double s = 0;
    for (long long t = 0; t < T; ++t) s += p[t];
    mean[r] = (float)(s / (double)T);
  }
};
struct SubRowMean {
  float* x;
  long long T;
  const float* mean;
  SV_HD void operator()(long long i) const { x[i] -= mean[i / T]; }
};

// Conv2d k x k (k = 3: padding 1, k = 1: padding 0), stride (stride, 1) over [C][F][T], eval BatchNorm as a per-channel
// affine map, optional residual, optional ReLU.  One thread per output element.
struct Conv2dBn {
  const float* in;      // [Ci][Fi][T]
  const float* w;       // [Co][Ci][k][k]
  const float* scale;   // [Co]
  const float* shift;
  const float* res;     // [Co][Fo][T] or null
  float* out;
  int Ci, Fi, Fo, T, k, stride, relu;
  SV_HD void operator()(long long i) const {
    const int t = (int)(i % T);
    const long long r = i / T;
    const int fo = (int)(r % Fo);
    const int co = (int)(r / Fo);
    const int pad = (k - 1) / 2;
    float acc = 0.f;
    for (int ci = 0; ci < Ci; ++ci) {
      const float* wp = w + ((long long)co * Ci + ci) * k * k;
      for (int df = 0; df < k; ++df) {
        const int fi = fo * stride + df - pad;
        if (fi < 0 || fi >= Fi) continue;
        const float* ip = in + ((long long)ci * Fi + fi) * T;
        for (int dt = 0; dt < k; ++dt) {
          const int tt = t + dt - pad;
          if (tt < 0 || tt >= T) continue;
          acc += ip[tt] * wp[df * k + dt];
        }
      }
    }
    float v = acc * scale[co] + shift[co];
    if (res) v += res[i];
    if (relu) v = fmaxf(v, 0.f);
    out[i] = v;
  }
};

// [C][T] -> channels-last rows [T][C] starting `row0` rows into dst
struct ToChannelsLast {
  const float* src;
  int C;
  long long T;
  float* dst;
  long long row0;
  SV_HD void operator()(long long i) const {
    const long long t = i / C;
    const int c = (int)(i - t * C);
    dst[(t + row0) * C + c] = src[(long long)c * T + t];
  }
};

// mode 0: relu(x * scale + shift)  (BatchNorm then ReLU: campplus `get_nonlinear("batchnorm-relu")`, layers.py:10-23)
// mode 1: relu(x) * scale + shift  (ReLU then BatchNorm: Conv1dReluBn, ecapa_tdnn.py:69-90)
// mode 2: relu(x)
struct AffineAct {
  const float* src;
  long long lds;
  float* dst;
  long long ldd;
  int cols;
  const float* scale;
  const float* shift;
  int mode;
  SV_HD void operator()(long long i) const {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    float v = src[r * lds + c];
    if (mode == 0) v = fmaxf(v * scale[c] + shift[c], 0.f);
    else if (mode == 1) v = fmaxf(v, 0.f) * scale[c] + shift[c];
    else v = fmaxf(v, 0.f);
    dst[r * ldd + c] = v;
  }
};

// sums of column c over the rows of segment `seg` (segments of seg_len rows, the last one short)
struct SegSum {
  const float* x;
  long long ld;
  int C, T, seg_len;
  float* out;   // [nseg][C]
  SV_HD void operator()(long long i) const {
    const int seg = (int)(i / C), c = (int)(i % C);
    const int t0 = seg * seg_len, t1 = t0 + seg_len < T ? t0 + seg_len : T;
    double s = 0;
    for (int t = t0; t < t1; ++t) s += x[(long long)t * ld + c];
    out[i] = (float)s;
  }
};

// CAMLayer context (layers.py:84-123): hidden[seg][o] = relu(W1 (mean_T(h) + mean_seg(h)) + b1)
struct CamHidden {
  const float* segsum;   // [nseg][C]
  int C, T, seg_len, nseg, H;
  const float* w1;       // [H][C]
  const float* b1;
  float* hid;            // [nseg][H]
  SV_HD void operator()(long long i) const {
    const int seg = (int)(i / H), o = (int)(i % H);
    const int t0 = seg * seg_len, t1 = t0 + seg_len < T ? t0 + seg_len : T;
    const float inv_T = 1.f / (float)T, inv_len = 1.f / (float)(t1 - t0);
    float acc = b1[o];
    for (int c = 0; c < C; ++c) {
      float tot = 0.f;
      for (int s = 0; s < nseg; ++s) tot += segsum[s * C + c];
      const float ctx = tot * inv_T + segsum[seg * C + c] * inv_len;
      acc += w1[o * C + c] * ctx;
    }
    hid[i] = fmaxf(acc, 0.f);
  }
};

// y[t][o] *= sigmoid(W2 hidden[seg(t)] + b2)
struct CamMaskApply {
  float* y;
  long long ld;
  int N, H, seg_len;
  const float* hid;
  const float* w2;   // [N][H]
  const float* b2;
  SV_HD void operator()(long long i) const {
    const long long t = i / N;
    const int o = (int)(i - t * N);
    const float* h = hid + (t / seg_len) * H;
    float acc = b2[o];
    for (int k = 0; k < H; ++k) acc += w2[o * H + k] * h[k];
    y[t * ld + o] *= 1.f / (1.f + expf(-acc));
  }
};

// masked statistics pooling (layers.py:26-50): mean and unbiased std of column c over the first `len` rows
struct StatsPool {
  const float* x;
  long long ld;
  int C, len;
  float* out;   // [2C]: means then stds
  SV_HD void operator()(long long c) const {
    double s = 0;
    for (int t = 0; t < len; ++t) s += x[(long long)t * ld + c];
    const double m = s / len;
    double q = 0;
    for (int t = 0; t < len; ++t) {
      const double d = x[(long long)t * ld + c] - m;
      q += d * d;
    }
    out[c] = (float)m;
    out[C + c] = (float)sqrt(q / (len - 1));
  }
};

// out[r][o] = post( act( sum_k x[r][k] W[o][k] + b[o] ) ) + res[r][o];  act 0 none, 1 relu, 2 sigmoid
struct SmallLinear {
  const float* x;
  long long ldx;
  const float* W;
  const float* b;
  const float* res;
  long long ldr;
  float* out;
  long long ldo;
  int N, K, act;
  const float* post_scale;
  const float* post_shift;
  SV_HD void operator()(long long i) const {
    const long long r = i / N;
    const int o = (int)(i - r * N);
    const float* xp = x + r * ldx;
    const float* wp = W + (long long)o * K;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += xp[k] * wp[k];
    if (b) acc += b[o];
    if (act == 1) acc = fmaxf(acc, 0.f);
    else if (act == 2) acc = 1.f / (1.f + expf(-acc));
    if (post_scale) acc = acc * post_scale[o] + post_shift[o];
    if (res) acc += res[r * ldr + o];
    out[r * ldo + o] = acc;
  }
};

// out[r][c] = a[r][c] (+ b[r][c])
struct AddCols {
  const float* a;
  long long lda;
  const float* b;
  long long ldb;
  float* out;
  long long ldo;
  int cols;
  SV_HD void operator()(long long i) const {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    float v = a[r * lda + c];
    if (b) v += b[r * ldb + c];
    out[r * ldo + c] = v;
  }
};

// squeeze-excitation gate and residual (ecapa_tdnn.py:96-133): out = x + h * g[c]
struct GateResidual {
  const float* x;
  long long ldx;
  const float* h;
  long long ldh;
  const float* g;
  float* out;
  long long ldo;
  int cols;
  SV_HD void operator()(long long i) const {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    out[r * ldo + c] = x[r * ldx + c] + h[r * ldh + c] * g[c];
  }
};

struct ColMean {
  const float* x;
  long long ld;
  int T;
  float* out;
  SV_HD void operator()(long long c) const {
    double s = 0;
    for (int t = 0; t < T; ++t) s += x[(long long)t * ld + c];
    out[c] = (float)(s / T);
  }
};

// PerceiverResampler cross attention (perceiver_encoder.py:100-175): 32 latent queries, keys / values = the first
// `nkeys` rows of [latents ; context] (rows past the mask never contribute: their scores are -float max).  Two passes:
// scores[q][h][j] = (q . k_j) * scale, one thread per score; then one thread per output element does the softmax
// statistics of its (query, head) row and the weighted sum of its value column.  Head dim 64.
struct PerceiverScores {
  const float* q;    // [nq][heads*64]
  const float* kv;   // [rows][2*heads*64]: k then v
  int heads, nkeys;
  float scale;
  float* scores;     // [nq][heads][nkeys]
  SV_HD void operator()(long long i) const {
    const int j = (int)(i % nkeys);
    const long long r = i / nkeys;
    const int h = (int)(r % heads);
    const long long qi = r / heads;
    const int D = heads * 64;
    const float* qp = q + qi * D + h * 64;
    const float* kp = kv + (long long)j * 2 * D + h * 64;
    float s = 0.f;
    for (int d = 0; d < 64; ++d) s += qp[d] * kp[d];
    scores[i] = s * scale;
  }
};
struct PerceiverSoftmaxV {
  const float* scores;   // [nq][heads][nkeys]
  const float* kv;
  int heads, nkeys;
  float* out;            // [nq][heads*64]
  SV_HD void operator()(long long i) const {
    const int D = heads * 64;
    const int c = (int)(i % D);                  // h * 64 + d
    const long long qi = i / D;
    const float* sp = scores + (qi * heads + c / 64) * nkeys;
    float m = -3.4028234663852886e38f;
    for (int j = 0; j < nkeys; ++j) m = sp[j] > m ? sp[j] : m;
    float l = 0.f, acc = 0.f;
    for (int j = 0; j < nkeys; ++j) {
      const float p = expf(sp[j] - m);
      l += p;
      acc += p * kv[(long long)j * 2 * D + D + c];
    }
    out[i] = acc / l;
  }
};

// GEGLU (perceiver_encoder.py:207-216): u = [h | gate], out = gelu(gate) * h
struct Geglu {
  const float* u;
  int inner;
  float* out;
  SV_HD void operator()(long long i) const {
    const long long r = i / inner;
    const int c = (int)(i - r * inner);
    const float h = u[r * 2 * inner + c], g = u[r * 2 * inner + inner + c];
    out[i] = 0.5f * g * (1.f + erff(g * 0.70710678118654752f)) * h;
  }
};

// final RMSNorm of the resampler (perceiver_encoder.py:177-190): normalize(x) * sqrt(dim) * gamma
struct LatNorm {
  const float* x;
  int dim;
  const float* gamma;
  float* out;
  SV_HD void operator()(long long i) const {
    const long long r = i / dim;
    const int c = (int)(i - r * dim);
    const float* p = x + r * dim;
    float s = 0.f;
    for (int k = 0; k < dim; ++k) s += p[k] * p[k];
    const float nrm = fmaxf(sqrtf(s), 1e-12f);
    out[i] = p[c] / nrm * sqrtf((float)dim) * gamma[c];
  }
};

// FSQ levels 4^6 (fsq/finite_scalar_quantization.py:126-162): code = round(tanh(z + shift) * half_l - 0.5) / 2
struct FsqCode {
  const float* z;
  float half_l, shift;
  float* codes;
  SV_HD void operator()(long long i) const { codes[i] = rintf(tanhf(z[i] + shift) * half_l - 0.5f) * 0.5f; }
};
struct FsqIndex {
  const float* codes;   // [n][6]
  int* idx;
  SV_HD void operator()(long long i) const {
    int v = 0, basis = 1;
    for (int d = 0; d < 6; ++d) {
      v += (int)(codes[i * 6 + d] * 2.f + 2.f) * basis;
      basis *= 4;
    }
    idx[i] = v;
  }
};

// ------------------------------------------------------------------------------------------------ weights
struct Bn {
  const float* scale = nullptr;
  const float* shift = nullptr;
};

// eval-mode BatchNorm1d/2d -> y = x * scale + shift (eps 1e-5); affine == false: no weight / bias
template <class BK>
Bn make_bn(BK& bk, const std::string& key, int dim, bool affine = true) {
  auto mean = bk.fetch(key + ".running_mean", {dim});
  auto var = bk.fetch(key + ".running_var", {dim});
  std::vector<float> w, b;
  if (affine) {
    w = bk.fetch(key + ".weight", {dim});
    b = bk.fetch(key + ".bias", {dim});
  }
  std::vector<float> sc((size_t)dim), sh((size_t)dim);
  for (int i = 0; i < dim; ++i) {
    const float inv = 1.f / std::sqrt(var[i] + 1e-5f);
    const float s = affine ? w[i] * inv : inv;
    sc[i] = s;
    sh[i] = (affine ? b[i] : 0.f) - mean[i] * s;
  }
  Bn r;
  r.scale = bk.upload(sc);
  r.shift = bk.upload(sh);
  return r;
}

inline std::vector<double> twiddles(int nfft) {
  std::vector<double> t((size_t)2 * nfft);
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < nfft; ++i) {
    t[2 * i] = std::cos(2.0 * pi * i / nfft);
    t[2 * i + 1] = std::sin(2.0 * pi * i / nfft);
  }
  return t;
}

// Conv1d weight [Co][Ci][k] -> [Co][k][Ci] (one GEMM over k overlapping channels-last rows)
inline std::vector<float> conv_rows(const std::vector<float>& w, int Co, int Ci, int k) {
  std::vector<float> o((size_t)Co * k * Ci);
  for (int co = 0; co < Co; ++co)
    for (int ci = 0; ci < Ci; ++ci)
      for (int j = 0; j < k; ++j) o[((size_t)co * k + j) * Ci + ci] = w[((size_t)co * Ci + ci) * k + j];
  return o;
}
// Conv1d weight [Co][Ci][k] -> [k][Co][Ci] (one GEMM tap per kernel element)
inline std::vector<float> conv_taps(const std::vector<float>& w, int Co, int Ci, int k) {
  std::vector<float> o((size_t)Co * k * Ci);
  for (int co = 0; co < Co; ++co)
    for (int ci = 0; ci < Ci; ++ci)
      for (int j = 0; j < k; ++j) o[((size_t)j * Co + co) * Ci + ci] = w[((size_t)co * Ci + ci) * k + j];
  return o;
}

// ================================================================================================ style (CAMPPlus)
constexpr int FB_WIN = 400, FB_SHIFT = 160, FB_PAD = 512, FB_BINS = 257, FB_MEL = 80;
constexpr int CAM_SEG = 100;

struct StyleNet {
  const float *window = nullptr, *banks = nullptr;
  const double* tw = nullptr;
  struct C2 {
    const float* w = nullptr;
    Bn bn;
  };
  C2 conv1, conv2;
  struct Res {
    C2 a, b, sc;
    int stride = 1;
  } res[4];
  const float* tdnn_w = nullptr;
  Bn tdnn_bn;
  struct Dense {
    Bn bn1, bn2;
    const float *w1 = nullptr, *wl = nullptr, *c1w = nullptr, *c1b = nullptr, *c2w = nullptr, *c2b = nullptr;
    int cin = 0;
  };
  struct Block {
    std::vector<Dense> layers;
    int dilation = 1, c0 = 0, cmax = 0;
    Bn tbn;
    const float* tw = nullptr;
  } blk[3];
  Bn out_bn, dense_bn;
  const float* dense_w = nullptr;
  bool ready = false;
};

// number of fbank frames of an n-sample wave (snip_edges) and of TDNN output rows
inline long long style_frames(long long n) { return n < FB_WIN ? 0 : 1 + (n - FB_WIN) / FB_SHIFT; }
inline long long style_rows(long long frames) { return (frames - 1) / 2 + 1; }
inline size_t style_ws_floats(long long n) {
  const long long T = style_frames(n);
  return (size_t)T * 12000 + (size_t)style_rows(T) * 6400 + (1u << 20);
}

template <class BK>
void style_finalize(BK& bk, StyleNet& net) {
  // derived buffers supplied by the host shim with the reference's own torch formulas (povey window, kaldi mel banks)
  net.window = bk.dev("fbank.window", {FB_WIN});
  net.banks = bk.dev("fbank.mel_banks", {FB_MEL, FB_BINS});
  net.tw = bk.upload_d(twiddles(FB_PAD));
  auto c2 = [&](const std::string& conv, const std::string& bn, int ci, int k) {
    StyleNet::C2 c;
    c.w = bk.dev(conv + ".weight", {32, ci, k, k});
    c.bn = make_bn(bk, bn, 32);
    return c;
  };
  net.conv1 = c2("head.conv1", "head.bn1", 1, 3);
  for (int l = 0; l < 2; ++l)
    for (int b = 0; b < 2; ++b) {
      const std::string p = "head.layer" + std::to_string(l + 1) + "." + std::to_string(b);
      StyleNet::Res& r = net.res[l * 2 + b];
      r.stride = b == 0 ? 2 : 1;
      r.a = c2(p + ".conv1", p + ".bn1", 32, 3);
      r.b = c2(p + ".conv2", p + ".bn2", 32, 3);
      if (b == 0) r.sc = c2(p + ".shortcut.0", p + ".shortcut.1", 32, 1);
    }
  net.conv2 = c2("head.conv2", "head.bn2", 32, 3);
  net.tdnn_w = bk.upload(conv_rows(bk.fetch("xvector.tdnn.linear.weight", {128, 320, 5}), 128, 320, 5));
  net.tdnn_bn = make_bn(bk, "xvector.tdnn.nonlinear.batchnorm", 128);
  const int nl[3] = {12, 24, 16}, dil[3] = {1, 2, 2};
  int ch = 128;
  for (int i = 0; i < 3; ++i) {
    StyleNet::Block& B = net.blk[i];
    B.dilation = dil[i];
    B.c0 = ch;
    B.cmax = ch + 32 * nl[i];
    B.layers.resize(nl[i]);
    for (int j = 0; j < nl[i]; ++j) {
      const std::string p = "xvector.block" + std::to_string(i + 1) + ".tdnnd" + std::to_string(j + 1);
      StyleNet::Dense& d = B.layers[j];
      d.cin = ch + 32 * j;
      d.bn1 = make_bn(bk, p + ".nonlinear1.batchnorm", d.cin);
      d.w1 = bk.dev(p + ".linear1.weight", {128, d.cin, 1});
      d.bn2 = make_bn(bk, p + ".nonlinear2.batchnorm", 128);
      d.wl = bk.upload(conv_taps(bk.fetch(p + ".cam_layer.linear_local.weight", {32, 128, 3}), 32, 128, 3));
      d.c1w = bk.dev(p + ".cam_layer.linear1.weight", {64, 128, 1});
      d.c1b = bk.dev(p + ".cam_layer.linear1.bias", {64});
      d.c2w = bk.dev(p + ".cam_layer.linear2.weight", {32, 64, 1});
      d.c2b = bk.dev(p + ".cam_layer.linear2.bias", {32});
    }
    ch = B.cmax;
    const std::string t = "xvector.transit" + std::to_string(i + 1);
    B.tbn = make_bn(bk, t + ".nonlinear.batchnorm", ch);
    B.tw = bk.dev(t + ".linear.weight", {ch / 2, ch, 1});
    ch /= 2;
  }
  net.out_bn = make_bn(bk, "xvector.out_nonlinear.batchnorm", 512);
  net.dense_w = bk.dev("dense.linear.weight", {192, 1024, 1});
  net.dense_bn = make_bn(bk, "dense.nonlinear.batchnorm", 192, false);
  net.ready = true;
}

// `torchaudio.compliance.kaldi.fbank(wave, num_mel_bins=80, dither=0, sample_frequency=16000)` (call site
// evaluations/infer_arvc.py:186-191): wave [n] -> log-mel [80][T] (frequency-major), T = style_frames(n)
template <class BK>
void kaldi_fbank(BK& bk, const StyleNet& net, const float* wave, long long n, float* feat) {
  const long long T = style_frames(n);
  SV_CHECK(T >= 1, "kaldi fbank: the wave is shorter than one 25 ms frame (400 samples at 16 kHz)");
  SV_CHECK(T < (1 << 20), "kaldi fbank: wave too long");
  float* fmean = bk.alloc(T);
  bk.pfor(T, FrameMean{wave, FB_WIN, FB_SHIFT, fmean});
  float* frames = bk.alloc(T * FB_PAD);
  bk.pfor(T * FB_PAD, KaldiFrame{wave, fmean, net.window, FB_WIN, FB_SHIFT, FB_PAD, frames});
  float* power = bk.alloc(T * FB_BINS);
  bk.pfor(T * FB_BINS, DftBin{frames, net.tw, FB_PAD, 0, FB_WIN, FB_BINS, 1, power});
  bk.pfor(T * FB_MEL, MelDot{power, FB_BINS, FB_MEL, net.banks, 1, FB_BINS, 1, 1.1920928955078125e-07f, feat, 1, T});
}

// `CAMPPlus.forward(x, x_lens)` (modules/campplus/DTDNN.py:132-138) for one row: feat [80][T] (frequency-major) and the
// number of valid rows `len` after the stride-2 TDNN -> out [192]
template <class BK>
void campplus_forward(BK& bk, const StyleNet& net, const float* feat, long long T, int len, float* out) {
  SV_CHECK(T >= 4 && T < (1 << 20), "CAMPPlus: between 4 and 2^20 feature frames");
  // ---- FCM (DTDNN.py:13-48): frequency 80 -> 40 -> 20 -> 10, 32 channels
  const long long plane = (long long)32 * 80 * T;
  float* b0 = bk.alloc(plane);
  float* b1 = bk.alloc(plane);
  float* b2 = bk.alloc(plane);
  float* b3 = bk.alloc(plane);
  const int Ti = (int)T;
  auto conv = [&](const StyleNet::C2& c, const float* in, int Ci, int Fi, int k, int stride, const float* res, int relu,
                  float* o) {
    const int pad = (k - 1) / 2;
    const int Fo = (Fi + 2 * pad - k) / stride + 1;
    bk.pfor((long long)32 * Fo * T, Conv2dBn{in, c.w, c.bn.scale, c.bn.shift, res, o, Ci, Fi, Fo, Ti, k, stride, relu});
    return Fo;
  };
  int F = conv(net.conv1, feat, 1, 80, 3, 1, nullptr, 1, b0);
  float* cur = b0;
  float* spare[3] = {b1, b2, b3};
  for (int r = 0; r < 4; ++r) {
    const StyleNet::Res& R = net.res[r];
    float *a = spare[0], *o = spare[1], *sc = spare[2];
    const int Fo = conv(R.a, cur, 32, F, 3, R.stride, nullptr, 1, a);
    const float* shortcut = cur;
    if (R.stride != 1) {
      conv(R.sc, cur, 32, F, 1, R.stride, nullptr, 0, sc);
      shortcut = sc;
    }
    conv(R.b, a, 32, Fo, 3, 1, shortcut, 1, o);
    spare[1] = cur;
    cur = o;
    F = Fo;
  }
  float* fcm = spare[0];
  F = conv(net.conv2, cur, 32, F, 3, 2, nullptr, 1, fcm);
  SV_CHECK(F == 10, "FCM output height");
  // ---- TDNN: Conv1d 320 -> 128, k 5, stride 2, padding 2 == one GEMM over 5 overlapping channels-last rows
  const long long Tp = style_rows(T);
  const int Tpi = (int)Tp;
  float* xt = bk.alloc((T + 4) * 320);
  bk.pfor(2 * 320, Fill{xt, 0.f});
  bk.pfor(2 * 320, Fill{xt + (T + 2) * 320, 0.f});
  bk.pfor(T * 320, ToChannelsLast{fcm, 320, T, xt, 2});
  float* X = bk.alloc(Tp * net.blk[0].cmax);
  {
    GemmParams p;
    p.A = xt; p.W = net.tdnn_w; p.C = X; p.M = Tpi; p.N = 128; p.K = 5 * 320; p.lda = 320; p.a_row_step = 2;
    p.ldc = net.blk[0].cmax;
    bk.gemm(p);
    bk.pfor(Tp * 128, AffineAct{X, p.ldc, X, p.ldc, 128, net.tdnn_bn.scale, net.tdnn_bn.shift, 0});
  }
  // ---- three CAM dense TDNN blocks (DTDNN.py:63-95, layers.py:126-204)
  const int nseg = (Tpi + CAM_SEG - 1) / CAM_SEG;
  float* h0 = bk.alloc(Tp * 1024);
  float* segsum = bk.alloc((long long)nseg * 128);
  float* hid = bk.alloc((long long)nseg * 64);
  for (int bi = 0; bi < 3; ++bi) {
    const StyleNet::Block& B = net.blk[bi];
    const int d = B.dilation, ld = B.cmax;
    float* hm = bk.alloc((Tp + 2 * d) * 128);          // bottleneck activations with d zero rows on both sides
    bk.pfor((long long)d * 128, Fill{hm, 0.f});
    bk.pfor((long long)d * 128, Fill{hm + (Tp + d) * 128, 0.f});
    float* h = hm + (long long)d * 128;
    for (const StyleNet::Dense& L : B.layers) {
      bk.pfor(Tp * L.cin, AffineAct{X, ld, h0, L.cin, L.cin, L.bn1.scale, L.bn1.shift, 0});
      GemmParams p1;
      p1.A = h0; p1.W = L.w1; p1.C = h; p1.M = Tpi; p1.N = 128; p1.K = L.cin; p1.lda = L.cin; p1.ldc = 128;
      bk.gemm(p1);
      bk.pfor(Tp * 128, AffineAct{h, 128, h, 128, 128, L.bn2.scale, L.bn2.shift, 0});
      GemmParams p2;
      p2.A = h; p2.W = L.wl; p2.C = X + L.cin; p2.M = Tpi; p2.N = 32; p2.K = 128; p2.lda = 128; p2.ldc = ld;
      p2.taps = 3; p2.tap_off[0] = -d; p2.tap_off[1] = 0; p2.tap_off[2] = d;
      bk.gemm(p2);
      bk.pfor((long long)nseg * 128, SegSum{h, 128, 128, Tpi, CAM_SEG, segsum});
      bk.pfor((long long)nseg * 64, CamHidden{segsum, 128, Tpi, CAM_SEG, nseg, 64, L.c1w, L.c1b, hid});
      bk.pfor(Tp * 32, CamMaskApply{X + L.cin, ld, 32, 64, CAM_SEG, hid, L.c2w, L.c2b});
    }
    // transit (layers.py:183-204): BatchNorm ReLU 1x1 conv to half the channels -> first columns of the next buffer
    bk.pfor(Tp * ld, AffineAct{X, ld, h0, ld, ld, B.tbn.scale, B.tbn.shift, 0});
    const int ldn = bi < 2 ? net.blk[bi + 1].cmax : ld / 2;
    float* Xn = bk.alloc(Tp * ldn);
    GemmParams pt;
    pt.A = h0; pt.W = B.tw; pt.C = Xn; pt.M = Tpi; pt.N = ld / 2; pt.K = ld; pt.lda = ld; pt.ldc = ldn;
    bk.gemm(pt);
    X = Xn;
  }
  // ---- out BatchNorm ReLU, masked statistics pooling over lens = frames // 2 rows, dense 1024 -> 192, BatchNorm
  bk.pfor(Tp * 512, AffineAct{X, 512, X, 512, 512, net.out_bn.scale, net.out_bn.shift, 0});
  SV_CHECK(len >= 2 && len <= Tpi, "CAMPPlus: statistics pooling needs 2 <= x_lens <= rows after the stride-2 TDNN");
  float* stats = bk.alloc(1024);
  bk.pfor(512, StatsPool{X, 512, 512, len, stats});
  bk.pfor(192, SmallLinear{stats, 1024, net.dense_w, nullptr, nullptr, 0, out, 192, 192, 1024, 0, net.dense_bn.scale,
                           net.dense_bn.shift});
}

// `InferenceWrapper.calculate_style_vec` for one row (evaluations/infer_arvc.py:179-211): fbank of the wave minus its
// time mean, lens = frames // 2, CAMPPlus.  wave [n] at 16 kHz -> out [192]
template <class BK>
void style_forward(BK& bk, const StyleNet& net, const float* wave, long long n, float* out) {
  const long long T = style_frames(n);
  SV_CHECK(T >= 4, "style encoder: the reference wave is shorter than 4 fbank frames (55 ms at 16 kHz)");
  float* feat = bk.alloc(FB_MEL * T);
  kaldi_fbank(bk, net, wave, n, feat);
  float* rmean = bk.alloc(FB_MEL);
  bk.pfor(FB_MEL, RowMean{feat, T, rmean});
  bk.pfor(FB_MEL * T, SubRowMean{feat, T, rmean});
  campplus_forward(bk, net, feat, T, (int)(T / 2), out);
}

// the two entry points above in the reference's own layout, features [T][80] (time-major rows)
template <class BK>
void kaldi_fbank_rows(BK& bk, const StyleNet& net, const float* wave, long long n, float* feat_rows) {
  const long long T = style_frames(n);
  SV_CHECK(T >= 1, "kaldi fbank: the wave is shorter than one 25 ms frame (400 samples at 16 kHz)");
  float* feat = bk.alloc(FB_MEL * T);
  kaldi_fbank(bk, net, wave, n, feat);
  bk.pfor(T * FB_MEL, ToChannelsLast{feat, FB_MEL, T, feat_rows, 0});
}
template <class BK>
void campplus_forward_rows(BK& bk, const StyleNet& net, const float* feat_rows, long long T, int len, float* out) {
  SV_CHECK(T >= 4 && T < (1 << 20), "CAMPPlus: between 4 and 2^20 feature frames");
  float* feat = bk.alloc(FB_MEL * T);
  // [T][80] -> [80][T]: ToChannelsLast with the roles of rows and channels swapped
  bk.pfor(T * FB_MEL, ToChannelsLast{feat_rows, (int)T, FB_MEL, feat, 0});
  campplus_forward(bk, net, feat, T, len, out);
}

// ================================================================================================ timbre (BiCodec)
constexpr int TM_NFFT = 1024, TM_HOP = 320, TM_WIN = 640, TM_BINS = 513, TM_MEL = 128;
constexpr int TM_LATENTS = 32, TM_DIM = 128;

struct TimbreNet {
  const float *window = nullptr, *fb = nullptr;
  const double* tw = nullptr;
  struct Crb {                 // conv -> ReLU -> BatchNorm
    const float *w = nullptr, *b = nullptr;
    Bn bn;
  };
  Crb layer1;
  struct Se {
    Crb in, convs[7], out;
    const float *l1w = nullptr, *l1b = nullptr, *l2w = nullptr, *l2b = nullptr;
    int dilation = 1;
  } se[3];
  const float *cat_w = nullptr, *cat_b = nullptr;
  const float *latents = nullptr, *ctx_w = nullptr, *ctx_b = nullptr, *gamma = nullptr;
  struct Layer {
    const float *wq, *wkv, *wo, *f0w, *f0b, *f2w, *f2b;
  } layers[2];
  const float *pin_w = nullptr, *pin_b = nullptr, *pout_w = nullptr, *pout_b = nullptr;
  bool ready = false;
};

inline long long timbre_frames(long long n) { return n / TM_HOP + 1; }
inline size_t timbre_ws_floats(long long n) {
  const long long T = timbre_frames(n);
  return (size_t)T * (TM_NFFT + TM_BINS + 128 + 512 * 5 + 1536 * 2 + 64 + 128 + 1024 + 64 + 256) + (1u << 20);
}

template <class BK>
void timbre_finalize(BK& bk, TimbreNet& net) {
  net.window = bk.dev("mel.window", {TM_NFFT});              // hann(640, periodic) centred in the 1024-sample frame
  net.fb = bk.dev("mel.fb", {TM_BINS, TM_MEL});              // slaney filterbank (torchaudio melscale_fbanks)
  net.tw = bk.upload_d(twiddles(TM_NFFT));
  const std::string e = "speaker_encoder";
  auto crb = [&](const std::string& p, const std::string& bn, int co, int ci, int k, int how) {
    TimbreNet::Crb c;
    if (how == 0) c.w = bk.dev(p + ".weight", {co, ci, k});
    else if (how == 1) c.w = bk.upload(conv_rows(bk.fetch(p + ".weight", {co, ci, k}), co, ci, k));
    else c.w = bk.upload(conv_taps(bk.fetch(p + ".weight", {co, ci, k}), co, ci, k));
    c.b = bk.dev(p + ".bias", {co});
    c.bn = make_bn(bk, bn, co);
    return c;
  };
  net.layer1 = crb(e + ".layer1.conv", e + ".layer1.bn", 512, 128, 5, 1);
  for (int l = 0; l < 3; ++l) {
    const std::string p = e + ".layer" + std::to_string(l + 2) + ".se_res2block";
    TimbreNet::Se& s = net.se[l];
    s.dilation = l + 2;
    s.in = crb(p + ".0.conv", p + ".0.bn", 512, 512, 1, 0);
    for (int i = 0; i < 7; ++i)
      s.convs[i] = crb(p + ".1.convs." + std::to_string(i), p + ".1.bns." + std::to_string(i), 64, 64, 3, 2);
    s.out = crb(p + ".2.conv", p + ".2.bn", 512, 512, 1, 0);
    s.l1w = bk.dev(p + ".3.linear1.weight", {128, 512});
    s.l1b = bk.dev(p + ".3.linear1.bias", {128});
    s.l2w = bk.dev(p + ".3.linear2.weight", {512, 128});
    s.l2b = bk.dev(p + ".3.linear2.bias", {512});
  }
  net.cat_w = bk.dev(e + ".conv.weight", {1536, 1536, 1});
  net.cat_b = bk.dev(e + ".conv.bias", {1536});
  const std::string ps = "perceiver_sampler";
  net.latents = bk.dev(ps + ".latents", {TM_LATENTS, TM_DIM});
  net.ctx_w = bk.dev(ps + ".proj_context.weight", {TM_DIM, 1536});
  net.ctx_b = bk.dev(ps + ".proj_context.bias", {TM_DIM});
  net.gamma = bk.dev(ps + ".norm.gamma", {TM_DIM});
  for (int l = 0; l < 2; ++l) {
    const std::string a = ps + ".layers." + std::to_string(l);
    TimbreNet::Layer& L = net.layers[l];
    L.wq = bk.dev(a + ".0.to_q.weight", {512, TM_DIM});
    L.wkv = bk.dev(a + ".0.to_kv.weight", {1024, TM_DIM});
    L.wo = bk.dev(a + ".0.to_out.weight", {TM_DIM, 512});
    L.f0w = bk.dev(a + ".1.0.weight", {682, TM_DIM});
    L.f0b = bk.dev(a + ".1.0.bias", {682});
    L.f2w = bk.dev(a + ".1.2.weight", {TM_DIM, 341});
    L.f2b = bk.dev(a + ".1.2.bias", {TM_DIM});
  }
  net.pin_w = bk.dev("quantizer.project_in.weight", {6, TM_DIM});
  net.pin_b = bk.dev("quantizer.project_in.bias", {6});
  net.pout_w = bk.dev("quantizer.project_out.weight", {TM_DIM, 6});
  net.pout_b = bk.dev("quantizer.project_out.bias", {TM_DIM});
  net.ready = true;
}

// wave [n] at 16 kHz, of which the first wave_len samples are valid (a zero-padded batch row; wave_len = n for a single
// utterance) -> timbre latents out [32][128]; optional FSQ indices [32] and FSQ inputs z [32][6]
template <class BK>
void timbre_forward(BK& bk, const TimbreNet& net, const float* wave, long long n, long long wave_len, float* out,
                    int* indices = nullptr, float* z_out = nullptr) {
  SV_CHECK(n >= TM_NFFT, "timbre encoder: the reference wave is shorter than 1024 samples at 16 kHz");
  SV_CHECK(wave_len >= 0 && wave_len <= n, "timbre encoder: wave_len must lie in [0, n_samples]");
  const long long T = timbre_frames(n);
  SV_CHECK(T < (1 << 20), "timbre encoder: reference wave too long");
  const int Ti = (int)T;
  // ---- mel magnitudes [T][128] with 2 zero rows on both sides (layer1 is a k = 5, padding 2 conv)
  float* frames = bk.alloc(T * TM_NFFT);
  bk.pfor(T * TM_NFFT, ReflectFrame{wave, n, net.window, TM_NFFT, TM_HOP, frames});
  float* mag = bk.alloc(T * TM_BINS);
  const int j0 = (TM_NFFT - TM_WIN) / 2;
  bk.pfor(T * TM_BINS, DftBin{frames, net.tw, TM_NFFT, j0, j0 + TM_WIN, TM_BINS, 0, mag});
  float* melb = bk.alloc((T + 4) * TM_MEL);
  bk.pfor(2 * TM_MEL, Fill{melb, 0.f});
  bk.pfor(2 * TM_MEL, Fill{melb + (T + 2) * TM_MEL, 0.f});
  float* mel = melb + 2 * TM_MEL;
  bk.pfor(T * TM_MEL, MelDot{mag, TM_BINS, TM_MEL, net.fb, TM_MEL, 1, 0, 0.f, mel, TM_MEL, 1});
  // ---- ECAPA-TDNN trunk (ecapa_tdnn.py:150-209), channels-last
  float* x1 = bk.alloc(T * 512);
  {
    GemmParams p;
    p.A = melb; p.W = net.layer1.w; p.bias = net.layer1.b; p.C = x1; p.M = Ti; p.N = 512; p.K = 5 * TM_MEL; p.lda = TM_MEL;
    p.ldc = 512;
    bk.gemm(p);
    bk.pfor(T * 512, AffineAct{x1, 512, x1, 512, 512, net.layer1.bn.scale, net.layer1.bn.shift, 1});
  }
  float* cat = bk.alloc(T * 1536);
  float* H = bk.alloc(T * 512);
  float* R = bk.alloc(T * 512);
  float* H2 = bk.alloc(T * 512);
  float* sbuf = bk.alloc((T + 8) * 64);
  float* cm = bk.alloc(512);
  float* g1 = bk.alloc(128);
  float* g2 = bk.alloc(512);
  const float* xin = x1;
  long long ldx = 512;
  for (int l = 0; l < 3; ++l) {
    const TimbreNet::Se& S = net.se[l];
    const int d = S.dilation;
    GemmParams p;
    p.A = xin; p.W = S.in.w; p.bias = S.in.b; p.C = H; p.M = Ti; p.N = 512; p.K = 512; p.lda = ldx; p.ldc = 512;
    bk.gemm(p);
    bk.pfor(T * 512, AffineAct{H, 512, H, 512, 512, S.in.bn.scale, S.in.bn.shift, 1});
    // Res2Conv1dReluBn (ecapa_tdnn.py:12-63): 8 splits of 64 channels, 7 chained dilated k3 convs
    bk.pfor((long long)d * 64, Fill{sbuf, 0.f});
    bk.pfor((long long)d * 64, Fill{sbuf + (T + d) * 64, 0.f});
    float* s = sbuf + (long long)d * 64;
    for (int i = 0; i < 7; ++i) {
      bk.pfor(T * 64, AddCols{H + i * 64, 512, i ? R + (i - 1) * 64 : nullptr, 512, s, 64, 64});
      GemmParams q;
      q.A = s; q.W = S.convs[i].w; q.bias = S.convs[i].b; q.C = R + i * 64; q.M = Ti; q.N = 64; q.K = 64; q.lda = 64;
      q.ldc = 512; q.taps = 3; q.tap_off[0] = -d; q.tap_off[1] = 0; q.tap_off[2] = d;
      bk.gemm(q);
      bk.pfor(T * 64, AffineAct{R + i * 64, 512, R + i * 64, 512, 64, S.convs[i].bn.scale, S.convs[i].bn.shift, 1});
    }
    bk.pfor(T * 64, AddCols{H + 7 * 64, 512, nullptr, 0, R + 7 * 64, 512, 64});
    GemmParams o;
    o.A = R; o.W = S.out.w; o.bias = S.out.b; o.C = H2; o.M = Ti; o.N = 512; o.K = 512; o.lda = 512; o.ldc = 512;
    bk.gemm(o);
    bk.pfor(T * 512, AffineAct{H2, 512, H2, 512, 512, S.out.bn.scale, S.out.bn.shift, 1});
    // squeeze-excitation gate (ecapa_tdnn.py:96-111) and the block's residual
    bk.pfor(512, ColMean{H2, 512, Ti, cm});
    bk.pfor(128, SmallLinear{cm, 512, S.l1w, S.l1b, nullptr, 0, g1, 128, 128, 512, 1, nullptr, nullptr});
    bk.pfor(512, SmallLinear{g1, 128, S.l2w, S.l2b, nullptr, 0, g2, 512, 512, 128, 2, nullptr, nullptr});
    float* dst = cat + l * 512;
    bk.pfor(T * 512, GateResidual{xin, ldx, H2, 512, g2, dst, 1536, 512});
    xin = dst;
    ldx = 1536;
  }
  float* feat = bk.alloc(T * 1536);
  {
    GemmParams p;
    p.A = cat; p.W = net.cat_w; p.bias = net.cat_b; p.C = feat; p.M = Ti; p.N = 1536; p.K = 1536; p.lda = 1536; p.ldc = 1536;
    bk.gemm(p);
    bk.pfor(T * 1536, AffineAct{feat, 1536, feat, 1536, 1536, nullptr, nullptr, 2});
  }
  // ---- PerceiverResampler (perceiver_encoder.py:300-351): keys / values = [latents ; context]
  const int NL = TM_LATENTS;
  float* kvin = bk.alloc((T + NL) * TM_DIM);
  {
    GemmParams p;
    p.A = feat; p.W = net.ctx_w; p.bias = net.ctx_b; p.C = kvin + NL * TM_DIM; p.M = Ti; p.N = TM_DIM; p.K = 1536;
    p.lda = 1536; p.ldc = TM_DIM;
    bk.gemm(p);
  }
  const int nkeys = NL + (int)(wave_len / TM_HOP);          // the mask keeps the latents and the first wave_len // 320 frames
  float* lat = bk.alloc(NL * TM_DIM);
  bk.pfor(NL * TM_DIM, AddCols{net.latents, TM_DIM, nullptr, 0, lat, TM_DIM, TM_DIM});
  float* q = bk.alloc(NL * 512);
  float* kv = bk.alloc((T + NL) * 1024);
  float* att = bk.alloc(NL * 512);
  float* scores = bk.alloc((long long)NL * 8 * (T + NL));
  float* u = bk.alloc(NL * 682);
  float* gg = bk.alloc(NL * 341);
  for (int l = 0; l < 2; ++l) {
    const TimbreNet::Layer& L = net.layers[l];
    bk.pfor(NL * TM_DIM, AddCols{lat, TM_DIM, nullptr, 0, kvin, TM_DIM, TM_DIM});
    bk.pfor(NL * 512, SmallLinear{lat, TM_DIM, L.wq, nullptr, nullptr, 0, q, 512, 512, TM_DIM, 0, nullptr, nullptr});
    GemmParams p;
    p.A = kvin; p.W = L.wkv; p.C = kv; p.M = Ti + NL; p.N = 1024; p.K = TM_DIM; p.lda = TM_DIM; p.ldc = 1024;
    bk.gemm(p);
    bk.pfor((long long)NL * 8 * nkeys, PerceiverScores{q, kv, 8, nkeys, 0.125f, scores});
    bk.pfor(NL * 512, PerceiverSoftmaxV{scores, kv, 8, nkeys, att});
    bk.pfor(NL * TM_DIM, SmallLinear{att, 512, L.wo, nullptr, lat, TM_DIM, lat, TM_DIM, TM_DIM, 512, 0, nullptr, nullptr});
    bk.pfor(NL * 682, SmallLinear{lat, TM_DIM, L.f0w, L.f0b, nullptr, 0, u, 682, 682, TM_DIM, 0, nullptr, nullptr});
    bk.pfor(NL * 341, Geglu{u, 341, gg});
    bk.pfor(NL * TM_DIM, SmallLinear{gg, 341, L.f2w, L.f2b, lat, TM_DIM, lat, TM_DIM, TM_DIM, 341, 0, nullptr, nullptr});
  }
  float* ln = bk.alloc(NL * TM_DIM);
  bk.pfor(NL * TM_DIM, LatNorm{lat, TM_DIM, net.gamma, ln});
  // ---- FSQ 4^6 (fsq/residual_fsq.py:70-76 projections, finite_scalar_quantization.py:126-162)
  float* z = z_out ? z_out : bk.alloc(NL * 6);
  bk.pfor(NL * 6, SmallLinear{ln, TM_DIM, net.pin_w, net.pin_b, nullptr, 0, z, 6, 6, TM_DIM, 0, nullptr, nullptr});
  const float half_l = 3.0f * 1.001f / 2.0f;
  const float shift = std::atanh(0.5f / half_l);
  float* codes = bk.alloc(NL * 6);
  bk.pfor(NL * 6, FsqCode{z, half_l, shift, codes});
  if (indices) bk.pfor(NL, FsqIndex{codes, indices});
  bk.pfor(NL * TM_DIM, SmallLinear{codes, 6, net.pout_w, net.pout_b, nullptr, 0, out, TM_DIM, TM_DIM, 6, 0, nullptr, nullptr});
}

}  // namespace spk
}  // namespace svanon
