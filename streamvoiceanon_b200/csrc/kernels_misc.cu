// Row-wise / element-wise kernels of the E, A(prefill) and V stages.  All activations are channels-last
// ([time][channel], channel contiguous) so every warp access is coalesced over channels.
#include "common.cuh"

namespace svanon {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum for blockDim.x <= 1024; every thread gets the result
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// LayerNorm over C for each row.  Matches both F.layer_norm (channels_last) and the hand-written
// channels_first LayerNorm of firefly.py:366-371: mean, biased variance of (x-mean), (x-mean)/sqrt(var+eps).
// MAXPT values per thread are kept in registers: C <= 128*MAXPT... launched with blockDim = 128.
template <int MAXPT>
__global__ void layernorm_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ w,
                                 const float* __restrict__ b, int C, float eps, int seg_rows, long long x_seg,
                                 long long y_seg) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const long long row = blockIdx.x;
  const float* xr = x + seg_row_off(row, seg_rows, x_seg, C);
  float* yr = y + seg_row_off(row, seg_rows, y_seg, C);
  float v[MAXPT];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXPT; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    v[i] = (c < C) ? xr[c] : 0.f;
    s += v[i];
  }
  const float mean = block_sum(s, sh) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXPT; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    const float d = (c < C) ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float var = block_sum(q, sh) / C;
  const float inv = 1.f / sqrtf(var + eps);
#pragma unroll
  for (int i = 0; i < MAXPT; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c < C) yr[c] = (v[i] - mean) * inv * w[c] + b[c];
  }
}

// depthwise causal conv (k = 7, FishConvNet left pad 6) followed by LayerNorm(C); one CTA per time step.
template <int MAXPT>
__global__ void dwconv7_ln_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ dw_w,
                                  const float* __restrict__ dw_b, const float* __restrict__ ln_w,
                                  const float* __restrict__ ln_b, int C, float eps, int seg_rows, long long x_seg,
                                  float* __restrict__ y_lo) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const long long row = blockIdx.x;
  const float* xr = x + seg_row_off(row, seg_rows, x_seg, C);      // rows -6..-1 are this stream's margin/history
  float v[MAXPT];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXPT; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    float a = 0.f;
    if (c < C) {
      a = dw_b[c];
#pragma unroll
      for (int j = 0; j < 7; ++j) a = fmaf(dw_w[j * C + c], xr[(j - 6) * C + c], a);
    }
    v[i] = a;
    s += a;
  }
  const float mean = block_sum(s, sh) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXPT; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    const float d = (c < C) ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float var = block_sum(q, sh) / C;
  const float inv = 1.f / sqrtf(var + eps);
#pragma unroll
  for (int i = 0; i < MAXPT; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c < C) {
      const float o = (v[i] - mean) * inv * ln_w[c] + ln_b[c];
      y[row * C + c] = o;
      if (y_lo) y_lo[row * C + c] = tf32_lo(o);
    }
  }
}

template <int MAXPT>
__global__ void rmsnorm_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ w, int C,
                               float eps, long long x_ld, float* __restrict__ y_lo) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const long long row = blockIdx.x;
  x += row * (x_ld - C);                                            // input rows may be strided (x_ld >= C)
  float v[MAXPT];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXPT; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    v[i] = (c < C) ? x[row * C + c] : 0.f;
    s += v[i] * v[i];
  }
  const float inv = rsqrtf(block_sum(s, sh) / C + eps);
#pragma unroll
  for (int i = 0; i < MAXPT; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c < C) {
      const float o = v[i] * inv * w[c];
      y[row * C + c] = o;
      if (y_lo) y_lo[row * C + c] = tf32_lo(o);
    }
  }
}

// apply_rotary_emb (dual_ar_stream.py:1004-1016 / windowed_transformer.py:368-380): pairs (2i,2i+1),
// table entry [pos][i] = (cos, sin) already rounded to bf16 and widened back to fp32.
__global__ void rope_qk_kernel(float* __restrict__ qkv, const float* __restrict__ table, int rows, int heads, int pos0,
                               int seg_rows) {
  pdl_trigger();
  pdl_wait();
  const int D = heads * HEAD_DIM;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over rows * 2 * D/2
  const long long total = (long long)rows * D;                              // q pairs + k pairs = 2 * D/2 * rows
  if (idx >= total) return;
  const int row = idx / D;
  const int r = idx % D;                 // 0 .. D-1 : first D/2 -> q pairs, next D/2 -> k pairs
  const int which = r / (D / 2);
  const int pair = r % (D / 2);
  const int i = pair % (HEAD_DIM / 2);
  float* p = qkv + (long long)row * 3 * D + which * D + pair * 2;
  const int pos = pos0 + (seg_rows > 0 ? row % seg_rows : row);     // every stream's window starts at pos0
  const float c = table[((long long)pos * (HEAD_DIM / 2) + i) * 2 + 0];
  const float s = table[((long long)pos * (HEAD_DIM / 2) + i) * 2 + 1];
  const float x0 = p[0], x1 = p[1];
  rope_pair(x0, x1, c, s, p[0], p[1]);
}

__global__ void silu_mul_kernel(const float* __restrict__ h, float* __restrict__ out, long long rows, int I,
                                float* __restrict__ out_lo) {
  pdl_trigger();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * I) return;
  const long long r = idx / I;
  const int c = idx % I;
  const float a = h[r * 2 * I + c], b = h[r * 2 * I + I + c];
  const float o = (a / (1.f + expf(-a))) * b;
  out[idx] = o;
  if (out_lo) out_lo[idx] = tf32_lo(o);
}

// LinearSpectrogram magnitude, spectrogram.py:62: sqrt(re^2 + im^2 + 1e-6); pad columns are zero.
__global__ void magnitude_kernel(const float* __restrict__ spec, float* __restrict__ mag, int T, int ld_in) {
  pdl_trigger();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)T * N_FREQ_PAD) return;
  const int t = idx / N_FREQ_PAD, f = idx % N_FREQ_PAD;
  float v = 0.f;
  if (f < N_FREQ) {
    const float re = spec[(long long)t * ld_in + f], im = spec[(long long)t * ld_in + N_FREQ + f];
    v = sqrtf(re * re + im * im + 1e-6f);
  }
  mag[idx] = v;
}

// LFQ.forward inference arithmetic (bsq.py:330-369): project_in 512->13 (+bias), l2norm (positive scale: does
// not change signs, skipped), bit_i = proj_i > 0, id = sum bit_i << (12 - i).  One warp per token.
__global__ void bsq_kernel(const float* __restrict__ z, const float* __restrict__ w, const float* __restrict__ b,
                           long long* __restrict__ ids, int T) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= T) return;
  const float* zr = z + (long long)warp * ENC_DIM;
  float zv[ENC_DIM / 32];
#pragma unroll
  for (int i = 0; i < ENC_DIM / 32; ++i) zv[i] = zr[lane + 32 * i];
  long long id = 0;
  for (int bit = 0; bit < BSQ_BITS; ++bit) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ENC_DIM / 32; ++i) s = fmaf(zv[i], w[bit * ENC_DIM + lane + 32 * i], s);
    s = warp_sum(s) + b[bit];
    if (s > 0.f) id |= 1LL << (BSQ_BITS - 1 - bit);
  }
  if (lane == 0) ids[warp] = id;
}

// GroupedResidualFSQ.get_output_from_indices for 8 groups x 1 quantizer, levels (8,5,5,5)
// (vendored twin: finite_scalar_quantization.py:143-162, residual_fsq.py:112-156).
__global__ void fsq_lookup_kernel(const long long* __restrict__ codes, long long ld, const float* __restrict__ w,
                                  const float* __restrict__ b, float* __restrict__ z, int T, int seg_rows,
                                  long long codes_seg) {
  pdl_trigger();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)T * 512) return;
  const int t = idx / 512, c = idx % 512, g = c / 64, o = c % 64;
  const long long id = seg_rows > 0 ? codes[(t / seg_rows) * codes_seg + g * ld + (t % seg_rows)] : codes[g * ld + t];
  const int levels[4] = {8, 5, 5, 5};
  const int basis[4] = {1, 8, 40, 200};
  float acc = b[g * 64 + o];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int digit = (int)((id / basis[i]) % levels[i]);
    const int half = levels[i] / 2;
    const float code = (float)(digit - half) / (float)half;
    acc = fmaf(code, w[(g * 64 + o) * 4 + i], acc);
  }
  z[idx] = acc;
}

// GroupedResidualFSQ.forward -> indices, 8 groups x 1 quantizer, levels (8,5,5,5) (vendored twin:
// finite_scalar_quantization.py:126-156): per group project_in 64 -> 4 (+bias), bound = tanh(x + shift) * half_l - offset
// with half_l = (L-1) * (1 + 1e-3) / 2, offset = 0.5 for even L, shift = atanh(offset / half_l); round half to even;
// digit = round + L/2; index = sum digit * basis.  One warp per (token, group); codes [B][8][T] int32.
__global__ void fsq_encode_kernel(const float* __restrict__ z, const float* __restrict__ w, const float* __restrict__ b,
                                  int* __restrict__ codes, int B, int T) {
  pdl_trigger();
  pdl_wait();
  const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (item >= B * T * 8) return;
  const int g = item & 7, tok = item >> 3;
  const float* zr = z + (long long)tok * 512 + g * 64;
  const float z0 = zr[lane], z1 = zr[lane + 32];
  const int levels[4] = {8, 5, 5, 5};
  const int basis[4] = {1, 8, 40, 200};
  int idx = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float* wr = w + (g * 4 + i) * 64;
    float s = fmaf(z1, wr[lane + 32], z0 * wr[lane]);
    s = warp_sum(s) + b[g * 4 + i];
    const float half_l = (float)(levels[i] - 1) * 1.001f / 2.f;
    const float offset = (levels[i] % 2 == 0) ? 0.5f : 0.f;
    const float shift = atanhf(offset / half_l);
    const float bounded = tanhf(s + shift) * half_l - offset;
    idx += ((int)rintf(bounded) + levels[i] / 2) * basis[i];
  }
  if (lane == 0) codes[((long long)(tok / T) * 8 + g) * T + (tok % T)] = idx;
}

// activation_post (SiLU) + conv_post (16 -> 1, k = 13, causal) + tanh, firefly.py:289-291.  Four lanes share one
// output sample (taps part, part+4, part+8, part+12) and combine with two shuffles: 4x more threads in flight for a
// kernel that only has L = 2048 outputs per frame.
__global__ void conv_post_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                 float* __restrict__ out, int L, int seg_rows, long long x_seg) {
  pdl_trigger();
  pdl_wait();
  __shared__ float ws[13 * 16];
  for (int i = threadIdx.x; i < 13 * 16; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long t = gid >> 2;
  const int part = (int)(gid & 3);
  float acc = 0.f;
  if (t < L) {
    const float* xr = x + seg_row_off(t, seg_rows, x_seg, 16) - 12 * 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int tap = part + 4 * i;
      if (tap < 13) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 v = __ldg(reinterpret_cast<const float4*>(xr + tap * 16) + q);
          v.x = __fdividef(v.x, 1.f + __expf(-v.x)); v.y = __fdividef(v.y, 1.f + __expf(-v.y));
          v.z = __fdividef(v.z, 1.f + __expf(-v.z)); v.w = __fdividef(v.w, 1.f + __expf(-v.w));
          const float* wp = ws + tap * 16 + q * 4;
          acc = fmaf(v.x, wp[0], acc); acc = fmaf(v.y, wp[1], acc);
          acc = fmaf(v.z, wp[2], acc); acc = fmaf(v.w, wp[3], acc);
        }
      }
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (t < L && part == 0) out[t] = tanhf(acc + b[0]);
}

// Polyphase sinc resampler with torchaudio.functional.resample semantics (host audio boundary: 16 kHz / device rate <->
// the model's 44.1 kHz, evaluations/infer_arvc.py:274-278): out[f*new + ph] = sum_k kernel[ph][k] * x[f*orig + k - width]
// with zeros outside [0, n_in).  One thread per output sample; consecutive threads are consecutive phases of (mostly)
// one input frame, so the input reads are warp-broadcasts and the kernel-matrix reads hit L1/L2 (new x taps floats).
__global__ void resample_kernel(const float* __restrict__ x, long long n_in, const float* __restrict__ kern, int orig, int nw,
                                int width, int taps, float* __restrict__ out, long long n_out) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const long long f = i / nw;
  const int ph = (int)(i - f * nw);
  const float* kr = kern + (long long)ph * taps;
  const long long x0 = f * orig - width;
  float acc = 0.f;
  for (int k = 0; k < taps; ++k) {
    const long long xi = x0 + k;
    const float xv = (xi >= 0 && xi < n_in) ? __ldg(x + xi) : 0.f;
    acc = fmaf(__ldg(kr + k), xv, acc);
  }
  out[i] = acc;
}

__global__ void gather_rows_kernel(const float* __restrict__ table, const long long* __restrict__ idx,
                                   float* __restrict__ out, int C, long long out_ld) {
  pdl_trigger();
  pdl_wait();
  const long long row = blockIdx.x;
  const float* src = table + idx[row] * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) out[row * out_ld + c] = src[c];
}

__global__ void embed_codes_kernel(const float* __restrict__ table, const int* __restrict__ codes, long long ld,
                                   float* __restrict__ out, long long out_ld) {
  pdl_trigger();
  pdl_wait();
  const long long t = blockIdx.x;
  for (int c = threadIdx.x; c < AR_DIM; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < AR_CODEBOOKS; ++i)
      s += table[((long long)codes[i * ld + t] + i * AR_CB_SIZE) * AR_DIM + c];
    out[t * out_ld + c] = s;
  }
}

__global__ void copy_rows_kernel(const float* __restrict__ src, long long src_ld, float* __restrict__ dst,
                                 long long dst_ld, int C) {
  pdl_trigger();
  pdl_wait();
  const long long row = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) dst[row * dst_ld + c] = src[row * src_ld + c];
}

__global__ void scale_add3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                  float* __restrict__ out, long long n, float s, long long seg_n, long long out_seg) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[seg_n > 0 ? (i / seg_n) * out_seg + i % seg_n : i] = (a[i] + b[i] + c[i]) * s;
}

__global__ void fill_kernel(float* __restrict__ p, long long n, float v, long long seg_stride) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[blockIdx.y * seg_stride + i] = v;
}

// out[r][c] = c < a_n ? a[r][c] : b[r][c - a_n]   (int32 sources, int32 or int64 destination)
template <typename OutT>
__global__ void concat_cols_kernel(const int* __restrict__ a, long long a_ld, int a_n, const int* __restrict__ b,
                                   long long b_ld, int b_n, OutT* __restrict__ out, long long out_ld, int rows) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = a_n + b_n;
  if (idx >= rows * n) return;
  const int r = idx / n, c = idx % n;
  out[r * out_ld + c] = (OutT)(c < a_n ? a[r * a_ld + c] : b[r * b_ld + (c - a_n)]);
}

__global__ void append_codes_kernel(const int* __restrict__ codes, int* __restrict__ hist, long long ld, int col) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x < AR_CODEBOOKS) hist[threadIdx.x * ld + col] = codes[threadIdx.x];
}

__global__ void i64_to_i32_kernel(const long long* __restrict__ in, int* __restrict__ out, long long n) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)in[i];
}

inline unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

}  // namespace

void launch_layernorm(const float* x, float* y, const float* w, const float* b, int rows, int C, float eps,
                      cudaStream_t st, int seg_rows, long long x_seg, long long y_seg) {
  if (rows <= 0) return;
  SV_CHECK(C <= 2048, "layernorm C");
  if (C <= 512) launch_pdl(layernorm_kernel<4>, dim3(rows), dim3(128), 0, st, x, y, w, b, C, eps, seg_rows, x_seg, y_seg);
  else launch_pdl(layernorm_kernel<8>, dim3(rows), dim3(256), 0, st, x, y, w, b, C, eps, seg_rows, x_seg, y_seg);
  SV_LAUNCHED();
}

void launch_dwconv7_ln(const float* x, float* y, const float* dw_w, const float* dw_b, const float* ln_w,
                       const float* ln_b, int rows, int C, float eps, cudaStream_t st, int seg_rows, long long x_seg,
                       float* y_lo) {
  if (rows <= 0) return;
  SV_CHECK(C <= 512, "dwconv C");
  launch_pdl(dwconv7_ln_kernel<4>, dim3(rows), dim3(128), 0, st, x, y, dw_w, dw_b, ln_w, ln_b, C, eps, seg_rows, x_seg, y_lo);
  SV_LAUNCHED();
}

void launch_rmsnorm(const float* x, float* y, const float* w, int rows, int C, float eps, cudaStream_t st,
                    long long x_ld, float* y_lo) {
  if (rows <= 0) return;
  SV_CHECK(C <= 1024, "rmsnorm C");
  launch_pdl(rmsnorm_kernel<4>, dim3(rows), dim3(256), 0, st, x, y, w, C, eps, x_ld > 0 ? x_ld : (long long)C, y_lo);
  SV_LAUNCHED();
}

void launch_rope_qk(float* qkv, const float* table, int rows, int heads, int pos0, cudaStream_t st, int seg_rows) {
  if (rows <= 0) return;
  const long long total = (long long)rows * heads * HEAD_DIM;
  launch_pdl(rope_qk_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, st, qkv, table, rows, heads, pos0, seg_rows);
  SV_LAUNCHED();
}

void launch_silu_mul(const float* h13, float* out, int rows, int I, cudaStream_t st, float* out_lo) {
  if (rows <= 0) return;
  launch_pdl(silu_mul_kernel, dim3(blocks_for((long long)rows * I, 256)), dim3(256), 0, st, h13, out, (long long)rows, I, out_lo);
  SV_LAUNCHED();
}

void launch_magnitude(const float* spec, float* mag, int T, int ld_in, cudaStream_t st) {
  if (T <= 0) return;
  launch_pdl(magnitude_kernel, dim3(blocks_for((long long)T * N_FREQ_PAD, 256)), dim3(256), 0, st, spec, mag, T, ld_in);
  SV_LAUNCHED();
}

void launch_bsq(const float* z, const float* w, const float* b, long long* ids, int T, cudaStream_t st) {
  if (T <= 0) return;
  launch_pdl(bsq_kernel, dim3(blocks_for((long long)T * 32, 128)), dim3(128), 0, st, z, w, b, ids, T);
  SV_LAUNCHED();
}

void launch_fsq_lookup(const long long* codes, long long ld, const float* w, const float* b, float* z, int T,
                       cudaStream_t st, int seg_rows, long long codes_seg) {
  if (T <= 0) return;
  launch_pdl(fsq_lookup_kernel, dim3(blocks_for((long long)T * 512, 256)), dim3(256), 0, st, codes, ld, w, b, z, T, seg_rows,
             codes_seg);
  SV_LAUNCHED();
}

void launch_fsq_encode(const float* z, const float* w, const float* b, int* codes, int B, int T, cudaStream_t st) {
  if (B * T <= 0) return;
  launch_pdl(fsq_encode_kernel, dim3(blocks_for((long long)B * T * 8 * 32, 128)), dim3(128), 0, st, z, w, b, codes, B, T);
  SV_LAUNCHED();
}

void launch_conv_post(const float* x, const float* w, const float* b, float* out, int L, cudaStream_t st, int seg_rows,
                      long long x_seg) {
  if (L <= 0) return;
  launch_pdl(conv_post_kernel, dim3(blocks_for((long long)L * 4, 256)), dim3(256), 0, st, x, w, b, out, L, seg_rows, x_seg);
  SV_LAUNCHED();
}

void launch_resample(const float* x, long long n_in, const float* kern, int orig, int nw, int width, int taps, float* out,
                     long long n_out, cudaStream_t st) {
  if (n_out <= 0) return;
  launch_pdl(resample_kernel, dim3(blocks_for(n_out, 256)), dim3(256), 0, st, x, n_in, kern, orig, nw, width, taps, out, n_out);
  SV_LAUNCHED();
}

namespace {
constexpr int NMIX_THREADS = 512;

__device__ __forceinline__ double nmix_block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();                                    // red may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < NMIX_THREADS / 32; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(NMIX_THREADS)
noise_mix_kernel(const float* x, const float* __restrict__ noise, long long n, float alpha, float* out) {   // out may alias x
  pdl_trigger();
  pdl_wait();
  __shared__ double red[NMIX_THREADS / 32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += NMIX_THREADS) s += (double)x[i];
  const double mean = nmix_block_sum(s, red) / (double)n;
  double q = 0.0;
  for (long long i = threadIdx.x; i < n; i += NMIX_THREADS) {
    const double d = (double)x[i] - mean;
    q += d * d;
  }
  const double var = nmix_block_sum(q, red) / (double)(n - 1);      // n == 1: 0/0 = NaN, like torch.std
  const float meanf = (float)mean, stdf = (float)sqrt(var);
  for (long long i = threadIdx.x; i < n; i += NMIX_THREADS) {
    const float nz = noise[i] * stdf + meanf;
    out[i] = alpha * x[i] + (1.f - alpha) * nz;
  }
}
}  // namespace

void launch_noise_mix(const float* x, const float* noise, long long n, float alpha, float* out, cudaStream_t st) {
  if (n <= 0) return;
  launch_pdl(noise_mix_kernel, dim3(1), dim3(NMIX_THREADS), 0, st, x, noise, n, alpha, out);
  SV_LAUNCHED();
}

void launch_gather_rows(const float* table, const long long* idx, float* out, int rows, int C, long long out_ld,
                        cudaStream_t st) {
  if (rows <= 0) return;
  launch_pdl(gather_rows_kernel, dim3(rows), dim3(256), 0, st, table, idx, out, C, out_ld);
  SV_LAUNCHED();
}

void launch_embed_codes(const float* table, const int* codes, long long ld, float* out, int T, long long out_ld,
                        cudaStream_t st) {
  if (T <= 0) return;
  launch_pdl(embed_codes_kernel, dim3(T), dim3(256), 0, st, table, codes, ld, out, out_ld);
  SV_LAUNCHED();
}

void launch_copy_rows(const float* src, long long src_ld, float* dst, long long dst_ld, int rows, int C,
                      cudaStream_t st) {
  if (rows <= 0) return;
  launch_pdl(copy_rows_kernel, dim3(rows), dim3(256), 0, st, src, src_ld, dst, dst_ld, C);
  SV_LAUNCHED();
}

void launch_scale_add3(const float* a, const float* b, const float* c, float* out, long long n, float s,
                       cudaStream_t st, long long seg_n, long long out_seg) {
  if (n <= 0) return;
  launch_pdl(scale_add3_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, st, a, b, c, out, n, s, seg_n, out_seg);
  SV_LAUNCHED();
}

void launch_concat_cols(const int* a, long long a_ld, int a_n, const int* b, long long b_ld, int b_n, void* out,
                        long long out_ld, int rows, bool out_i64, cudaStream_t st) {
  const int n = rows * (a_n + b_n);
  if (n <= 0) return;
  if (out_i64) launch_pdl(concat_cols_kernel<long long>, dim3(blocks_for(n, 256)), dim3(256), 0, st, a, a_ld, a_n, b, b_ld, b_n, (long long*)out, out_ld, rows);
  else launch_pdl(concat_cols_kernel<int>, dim3(blocks_for(n, 256)), dim3(256), 0, st, a, a_ld, a_n, b, b_ld, b_n, (int*)out, out_ld, rows);
  SV_LAUNCHED();
}

void launch_append_codes(const int* codes, int* hist, long long ld, int col, cudaStream_t st) {
  launch_pdl(append_codes_kernel, dim3(1), dim3(32), 0, st, codes, hist, ld, col);
  SV_LAUNCHED();
}

void launch_i64_to_i32(const long long* in, int* out, long long n, cudaStream_t st) {
  if (n <= 0) return;
  launch_pdl(i64_to_i32_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, st, in, out, n);
  SV_LAUNCHED();
}

void launch_fill(float* p, long long n, float v, cudaStream_t st, int nseg, long long seg_stride) {
  if (n <= 0 || nseg <= 0) return;
  launch_pdl(fill_kernel, dim3(blocks_for(n, 256), nseg), dim3(256), 0, st, p, n, v, seg_stride);
  SV_LAUNCHED();
}

}  // namespace svanon
