// VC.pipeline glue kernels: IVF-Flat (nprobe=1) top-8 search + weighted reconstruction + index_rate blend,
// nearest x2 feature upsampling with the `protect` blend, the 160-tap box sum used to find quiet
// cut points, peak scan and int16 conversion.
#include "common.cuh"
#include "../../include/b200vc.h"

namespace b200vc {
namespace {

inline unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// row-wise argmin (first occurrence) of S[rows, n]
__global__ void argmin_rows_kernel(const float* __restrict__ S, int* __restrict__ out, int n, long long ld) {
  const int row = blockIdx.x;
  const float* s = S + (long long)row * ld;
  float best = INFINITY;
  int bi = 0x7fffffff;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float v = s[j];
    if (v < best) { best = v; bi = j; }
  }
  __shared__ float sv[32];
  __shared__ int si[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov < best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    best = threadIdx.x < nw ? sv[threadIdx.x] : INFINITY;
    bi = threadIdx.x < nw ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov < best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (threadIdx.x == 0) out[row] = bi;
  }
}

// ---------------------------------------------------------------------------
// One block per query: scan the assigned inverted list, keep the 8 nearest (squared L2, ties by list
// position), then  w = (1/d)^2 / sum ; npy = sum_k w_k v_k ; out = rate*npy + (1-rate)*q
// (vc_infer_pipeline.py:421-431).  Arithmetic is fp32 with separate mul/add like numpy.
// ---------------------------------------------------------------------------
constexpr int IVF_K = 8;
constexpr int IVF_WARPS = 8;

__global__ void __launch_bounds__(IVF_WARPS * 32)
ivf_scan_blend_kernel(const float* __restrict__ q, long long ldq, const int* __restrict__ assign,
                      const int* __restrict__ offsets, const long long* __restrict__ ids,
                      const float* __restrict__ vecs, float* __restrict__ out, long long ldo, int d, float rate,
                      float* __restrict__ out_score, long long* __restrict__ out_ids) {
  extern __shared__ float qs[];                    // d floats
  __shared__ float cd[IVF_WARPS][IVF_K];
  __shared__ int cp[IVF_WARPS][IVF_K];
  __shared__ float topd[IVF_K];
  __shared__ int topp[IVF_K];
  __shared__ float wnorm[IVF_K];
  const int t = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* qr = q + (long long)t * ldq;
  for (int c = threadIdx.x; c < d; c += blockDim.x) qs[c] = qr[c];
  if (lane < IVF_K) { cd[warp][lane] = INFINITY; cp[warp][lane] = 0x7fffffff; }
  __syncthreads();
  const int list = assign[t];
  const int beg = offsets[list], end = offsets[list + 1];
  for (int p = beg + warp; p < end; p += IVF_WARPS) {
    const float* v = vecs + (long long)p * d;
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float df = qs[c] - v[c];
      acc = fmaf(df, df, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      // insert (acc, p) into this warp's sorted top-K (ascending distance, then position)
      int pos = IVF_K;
      while (pos > 0 && (acc < cd[warp][pos - 1] || (acc == cd[warp][pos - 1] && p < cp[warp][pos - 1]))) --pos;
      if (pos < IVF_K) {
        for (int j = IVF_K - 1; j > pos; --j) { cd[warp][j] = cd[warp][j - 1]; cp[warp][j] = cp[warp][j - 1]; }
        cd[warp][pos] = acc;
        cp[warp][pos] = p;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // merge IVF_WARPS sorted lists
    int head[IVF_WARPS];
    for (int w = 0; w < IVF_WARPS; ++w) head[w] = 0;
    for (int k = 0; k < IVF_K; ++k) {
      float bd = INFINITY;
      int bp = 0x7fffffff, bw = -1;
      for (int w = 0; w < IVF_WARPS; ++w) {
        if (head[w] >= IVF_K) continue;
        const float dd = cd[w][head[w]];
        const int pp = cp[w][head[w]];
        if (pp == 0x7fffffff) continue;
        if (bw < 0 || dd < bd || (dd == bd && pp < bp)) { bd = dd; bp = pp; bw = w; }
      }
      if (bw >= 0) { ++head[bw]; topd[k] = bd; topp[k] = bp; }
      else { topd[k] = INFINITY; topp[k] = -1; }     // faiss pads short lists with +inf / label -1
    }
    float wsum = 0.f, w[IVF_K];
    for (int k = 0; k < IVF_K; ++k) {
      const float inv = __fdiv_rn(1.f, topd[k]);
      w[k] = __fmul_rn(inv, inv);
      wsum = __fadd_rn(wsum, w[k]);
    }
    for (int k = 0; k < IVF_K; ++k) wnorm[k] = __fdiv_rn(w[k], wsum);
    if (out_score)
      for (int k = 0; k < IVF_K; ++k) {
        out_score[(long long)t * IVF_K + k] = topd[k];
        out_ids[(long long)t * IVF_K + k] = topp[k] >= 0 ? ids[topp[k]] : -1;
      }
  }
  __syncthreads();
  const float one_minus = __fsub_rn(1.f, rate);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < IVF_K; ++k) {
      // short lists: big_npy[-1] * 0 in the reference -> contributes 0 for finite data
      const float v = topp[k] >= 0 ? vecs[(long long)topp[k] * d + c] : 0.f;
      acc = __fadd_rn(acc, __fmul_rn(v, wnorm[k]));
    }
    out[(long long)t * ldo + c] = __fadd_rn(__fmul_rn(acc, rate), __fmul_rn(one_minus, qs[c]));
  }
}

// out[p,c] = a*pf + b*(1-pf), a = feats[p/2,c], b = feats0[p/2,c]; pf from pitchf (vc_infer_pipeline.py:433-452)
__global__ void upsample2_protect_kernel(const float* __restrict__ feats, const float* __restrict__ feats0,
                                         const float* __restrict__ pitchf, float* __restrict__ out, long long P,
                                         int C, float protect, int do_protect) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P * C) return;
  const long long p = e / C;
  const int c = (int)(e % C);
  const float a = feats[(p >> 1) * C + c];
  if (!do_protect) { out[e] = a; return; }
  float pf = pitchf[p];
  if (pf > 0.f) pf = 1.f;
  if (pitchf[p] < 1.f) pf = protect;
  const float b = feats0[(p >> 1) * C + c];
  out[e] = __fadd_rn(__fmul_rn(a, pf), __fmul_rn(b, __fsub_rn(1.f, pf)));
}

// out[i] = sum_{j<window} x[i+j] accumulated in index order in fp64 (vc_infer_pipeline.py:517-519)
__global__ void boxsum_f64_kernel(const double* __restrict__ x, double* __restrict__ out, long long n, int window) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int j = 0; j < window; ++j) s += x[i + j];
  out[i] = s;
}


// Mono down-mix + band-limited resampling (Hann-windowed sinc, `zc` zero crossings each side), rate_in -> rate_out.
// Stand-in for the ffmpeg "-ac 1 -ar 16000" decode at my_utils.py:13-17 (ingest is SURVEY.md 8(f) rank 2).
__global__ void resample_sinc_mono_kernel(const float* __restrict__ x, long long n_in, int channels,
                                          float* __restrict__ out, long long n_out, double ratio /*in/out*/, int zc) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_out) return;
  const double scale = ratio > 1.0 ? 1.0 / ratio : 1.0;        // cutoff relative to the input Nyquist
  const double center = (double)m * ratio;
  const double halfw = (double)zc / scale;
  long long k0 = (long long)ceil(center - halfw), k1 = (long long)floor(center + halfw);
  float acc = 0.f;
  for (long long k = k0; k <= k1; ++k) {
    if (k < 0 || k >= n_in) continue;
    const float d = (float)((double)k - center);
    const float a = d * (float)scale;
    const float sinc = fabsf(a) < 1e-6f ? 1.f : sinf(3.14159265358979f * a) / (3.14159265358979f * a);
    const float win = 0.5f + 0.5f * cosf(3.14159265358979f * d / (float)halfw);
    float v = 0.f;
    for (int c = 0; c < channels; ++c) v += x[(long long)c * n_in + k];
    acc += (v / channels) * sinc * win * (float)scale;
  }
  out[m] = acc;
}

}  // namespace
}  // namespace b200vc

using namespace b200vc;

extern "C" {

int b200vc_argmin_rows(const float* S, int* out, int rows, int n, int64_t ld, void* stream) {
  B200VC_REQUIRE(S && out && rows > 0 && n > 0, "argmin_rows: bad args");
  argmin_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(S, out, n, ld);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_ivf_scan_blend(const float* q, int64_t ldq, const int* assign, const int* offsets, const int64_t* ids,
                          const float* vecs, float* out, int64_t ldo, int T, int d, float rate, float* out_score,
                          int64_t* out_ids, void* stream) {
  B200VC_REQUIRE(q && assign && offsets && ids && vecs && out && T > 0 && d > 0 && d * 4 <= 40 * 1024,
                 "ivf_scan_blend: bad args");
  ivf_scan_blend_kernel<<<(unsigned)T, IVF_WARPS * 32, d * sizeof(float), (cudaStream_t)stream>>>(
      q, ldq, assign, offsets, reinterpret_cast<const long long*>(ids), vecs, out, ldo, d, rate, out_score,
      reinterpret_cast<long long*>(out_ids));
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_upsample2_protect(const float* feats, const float* feats0, const float* pitchf, float* out, int64_t P,
                             int C, float protect, int do_protect, void* stream) {
  B200VC_REQUIRE(feats && out && P > 0 && C > 0 && (!do_protect || (feats0 && pitchf)), "upsample2_protect: bad args");
  upsample2_protect_kernel<<<blocks_for(P * C, 256), 256, 0, (cudaStream_t)stream>>>(feats, feats0, pitchf, out, P, C,
                                                                                  protect, do_protect);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_boxsum_f64(const double* x, double* out, int64_t n, int window, void* stream) {
  B200VC_REQUIRE(x && out && n > 0 && window > 0, "boxsum_f64: bad args");
  boxsum_f64_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, out, n, window);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_resample_sinc_mono(const float* x, int64_t n_in, int channels, float* out, int64_t n_out, double ratio,
                              int zero_crossings, void* stream) {
  B200VC_REQUIRE(x && out && n_in > 0 && n_out > 0 && channels > 0 && ratio > 0, "resample_sinc_mono: bad args");
  resample_sinc_mono_kernel<<<blocks_for(n_out, 256), 256, 0, (cudaStream_t)stream>>>(x, n_in, channels, out, n_out, ratio,
                                                                                    zero_crossings);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// change_rms + peak guard + int16 conversion on the device (vc_infer_pipeline.py:41-60, 645-649)
// ---------------------------------------------------------------------------------------------------------
namespace b200vc {
namespace {

// librosa.feature.rms(center=True, reflect): one block per frame, fp64 accumulation
template <typename T>
__global__ void frame_rms_kernel(const T* __restrict__ y, long long n, int frame, int hop, double* __restrict__ rms) {
  const long long f = blockIdx.x;
  const long long start = f * hop - frame / 2;
  double acc = 0.0;
  for (int i = threadIdx.x; i < frame; i += blockDim.x) {
    long long j = start + i;
    if (j < 0) j = -j;
    if (j >= n) j = 2 * (n - 1) - j;
    const double v = (j >= 0 && j < n) ? (double)y[j] : 0.0;
    acc += v * v;
  }
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    rms[f] = sqrt(t / frame);
  }
}

__device__ __forceinline__ double lerp_frames(const double* __restrict__ r, int nf, long long i, double scale) {
  // F.interpolate(mode="linear", align_corners=False)
  double src = ((double)i + 0.5) * scale - 0.5;
  if (src < 0.0) src = 0.0;
  int i0 = (int)src;
  if (i0 > nf - 1) i0 = nf - 1;
  const int i1 = i0 + (i0 < nf - 1 ? 1 : 0);
  const double l1 = src - (double)i0;
  return (1.0 - l1) * r[i0] + l1 * r[i1];
}

// data2[i] *= rms1(i)^(1-rate) * max(rms2(i),1e-6)^(rate-1)
__global__ void rms_mix_kernel(float* __restrict__ data2, long long n2, const double* __restrict__ rms1, int nf1,
                               const double* __restrict__ rms2, int nf2, double rate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  const double r1 = lerp_frames(rms1, nf1, i, (double)nf1 / (double)n2);
  float r2 = (float)lerp_frames(rms2, nf2, i, (double)nf2 / (double)n2);
  r2 = fmaxf(r2, 1e-6f);
  const double f = pow(r1, 1.0 - rate) * (double)powf(r2, (float)(rate - 1.0));
  data2[i] = (float)((double)data2[i] * f);
}

__global__ void absmax_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));   // m >= 0: int order == float order
}

// out = (int16)(x * (abs_max/0.99 > 1 ? 32768/(abs_max/0.99) : 32768))  with C truncation toward zero
__global__ void to_int16_kernel(const float* __restrict__ x, long long n, const float* __restrict__ absmax,
                                short* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float audio_max = absmax[0] / 0.99f;
  double max_int16 = 32768.0;
  if (audio_max > 1.f) max_int16 /= (double)audio_max;
  out[i] = (short)(int)((double)x[i] * max_int16);
}

}  // namespace
}  // namespace b200vc

extern "C" {

int b200vc_change_rms(const double* data1, int64_t n1, int sr1, float* data2, int64_t n2, int sr2, double rate,
                      double* scratch, void* stream) {
  B200VC_REQUIRE(data1 && data2 && scratch && n1 > 0 && n2 > 0 && sr1 > 1 && sr2 > 1, "change_rms: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  const int frame1 = sr1 / 2 * 2, hop1 = sr1 / 2, frame2 = sr2 / 2 * 2, hop2 = sr2 / 2;
  const int nf1 = 1 + (int)(n1 / hop1), nf2 = 1 + (int)(n2 / hop2);
  double* rms1 = scratch;
  double* rms2 = scratch + nf1;
  b200vc::frame_rms_kernel<double><<<nf1, 256, 0, s>>>(data1, n1, frame1, hop1, rms1);
  b200vc::frame_rms_kernel<float><<<nf2, 256, 0, s>>>(data2, n2, frame2, hop2, rms2);
  b200vc::rms_mix_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, s>>>(data2, n2, rms1, nf1, rms2, nf2, rate);
  b200vc::count_launch(3);
  B200VC_LAUNCH_CHECK();
  return b200vc::kOk;
}

int b200vc_to_int16_peak_guard(const float* x, int64_t n, float* scratch_absmax, int16_t* out, void* stream) {
  B200VC_REQUIRE(x && scratch_absmax && out && n > 0, "to_int16_peak_guard: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  B200VC_CHECK_CUDA(cudaMemsetAsync(scratch_absmax, 0, sizeof(float), s));
  b200vc::absmax_kernel<<<592, 256, 0, s>>>(x, n, scratch_absmax);
  b200vc::to_int16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, n, scratch_absmax, out);
  b200vc::count_launch(2);
  B200VC_LAUNCH_CHECK();
  return b200vc::kOk;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Zero-phase IIR in fp64, block-parallel, as a cascade of second-order sections (scipy.signal.sosfiltfilt semantics:
// odd extension, sosfilt_zi initial conditions).  Per section and direction:
//   pass 1: every thread runs the DF2T recurrence over its own block of L samples from a zero state
//   pass 2: one thread chains the block-end states:  s_{k+1} = zs_end_k + M_L s_k
//   pass 3: y[n] += h_{n mod L} . s_k   (response of the section to the carried-in state)
// Replaces `signal.filtfilt(bh, ah, audio)` (vc_infer_pipeline.py:22, 513; 5th-order Butterworth high-pass).  The
// reference's transfer-function form is itself only accurate to ~7e-7 here (scipy's filtfilt and sosfiltfilt of the same
// filter differ by that much); this cascade matches scipy.sosfiltfilt to ~4e-13.
// ---------------------------------------------------------------------------------------------------------
namespace b200vc {
namespace {

struct SosCoef { double b0, b1, b2, a1, a2; };

__global__ void odd_ext_kernel(const float* __restrict__ x, long long n, int pad, double* __restrict__ ext) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = n + 2 * pad;
  if (i >= total) return;
  double v;
  if (i < pad) v = 2.0 * (double)x[0] - (double)x[pad - i];
  else if (i < pad + n) v = (double)x[i - pad];
  else v = 2.0 * (double)x[n - 1] - (double)x[n - 2 - (i - pad - n)];
  ext[i] = v;
}

__global__ void sos_block_zs_kernel(const double* __restrict__ x, double* __restrict__ y, double* __restrict__ zs_end,
                                    long long n, int L, SosCoef c, int reverse) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nblk = (n + L - 1) / L;
  if (k >= nblk) return;
  double z0 = 0.0, z1 = 0.0;
  const long long beg = k * L, end = min(n, beg + (long long)L);
  for (long long t = beg; t < end; ++t) {
    const long long i = reverse ? (n - 1 - t) : t;
    const double xv = x[i];
    const double yv = c.b0 * xv + z0;
    z0 = c.b1 * xv + z1 - c.a1 * yv;
    z1 = c.b2 * xv - c.a2 * yv;
    y[i] = yv;
  }
  zs_end[k * 2 + 0] = z0;
  zs_end[k * 2 + 1] = z1;
}

__global__ void sos_scan_states_kernel(const double* __restrict__ zs_end, const double* __restrict__ ML,
                                       const double* __restrict__ zi, const double* __restrict__ scale_sample,
                                       double* __restrict__ states, long long nblk) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s0 = zi[0] * scale_sample[0], s1 = zi[1] * scale_sample[0];
  const double m00 = ML[0], m01 = ML[1], m10 = ML[2], m11 = ML[3];
  for (long long k = 0; k < nblk; ++k) {
    states[k * 2 + 0] = s0;
    states[k * 2 + 1] = s1;
    const double t0 = zs_end[k * 2 + 0] + m00 * s0 + m01 * s1;
    const double t1 = zs_end[k * 2 + 1] + m10 * s0 + m11 * s1;
    s0 = t0;
    s1 = t1;
  }
}

__global__ void sos_fix_kernel(double* __restrict__ y, const double* __restrict__ states, const double* __restrict__ H,
                               long long n, int L, int reverse) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long long k = t / L;
  const int j = (int)(t % L);
  const long long i = reverse ? (n - 1 - t) : t;
  y[i] += H[j * 2 + 0] * states[k * 2 + 0] + H[j * 2 + 1] * states[k * 2 + 1];
}

__global__ void copy_slice_f64_kernel(const double* __restrict__ src, double* __restrict__ dst, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
}  // namespace
}  // namespace b200vc

extern "C" {

int b200vc_sosfiltfilt_f64(const float* x, int64_t n, const double* sos_host, int nsec, const double* zi_dev,
                           const double* H_dev, const double* ML_dev, int L, int padlen, double* work, double* out,
                           void* stream) {
  using namespace b200vc;
  B200VC_REQUIRE(x && sos_host && zi_dev && H_dev && ML_dev && work && out && nsec >= 1 && nsec <= 8 && L >= 16 &&
                     padlen >= 0 && n > padlen + 1, "sosfiltfilt: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  const long long ne = n + 2 * padlen;
  const long long nblk = (ne + L - 1) / L;
  double* bufA = work;
  double* bufB = work + ne;
  double* edge = work + 2 * ne;           // 2 doubles: scale samples
  double* zs = edge + 2;
  double* st = zs + nblk * 2;
  const unsigned ge = (unsigned)((ne + 255) / 256), gb = (unsigned)((nblk + 63) / 64);
  odd_ext_kernel<<<ge, 256, 0, s>>>(x, n, padlen, bufA);
  int launches = 1;
  double* cur = bufA;
  double* nxt = bufB;
  for (int dir = 0; dir < 2; ++dir) {
    // all sections share the scale sample: first sample of the (direction's) input signal (scipy: zi * x_0)
    const double* scale_src = dir == 0 ? cur : cur + (ne - 1);
    B200VC_CHECK_CUDA(cudaMemcpyAsync(edge + dir, scale_src, sizeof(double), cudaMemcpyDeviceToDevice, s));
    for (int sec = 0; sec < nsec; ++sec) {
      const double* q = sos_host + sec * 6;
      SosCoef c{q[0] / q[3], q[1] / q[3], q[2] / q[3], q[4] / q[3], q[5] / q[3]};
      sos_block_zs_kernel<<<gb, 64, 0, s>>>(cur, nxt, zs, ne, L, c, dir);
      sos_scan_states_kernel<<<1, 32, 0, s>>>(zs, ML_dev + sec * 4, zi_dev + sec * 2, edge + dir, st, nblk);
      sos_fix_kernel<<<ge, 256, 0, s>>>(nxt, st, H_dev + (long long)sec * L * 2, ne, L, dir);
      launches += 3;
      double* t = cur; cur = nxt; nxt = t;
    }
  }
  copy_slice_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cur + padlen, out, n);
  count_launch(launches + 1);
  B200VC_LAUNCH_CHECK();
  return kOk;
}

}  // extern "C"
