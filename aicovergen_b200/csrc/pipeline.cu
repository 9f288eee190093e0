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

// out[ch,n] = ga * lerp(a_mono, n * ra) + gb * b[ch,n] + gc * c[ch,n]  — gain-and-sum stand-in for the pydub overlay
// at main.py:229-233 (a = converted vocals at its own rate, b = backup vocals, c = instrumental)
__global__ void mix3_kernel(const float* __restrict__ a, long long n_a, double ra, const float* __restrict__ b,
                            const float* __restrict__ c, float* __restrict__ out, long long n, float ga, float gb,
                            float gc) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 2 * n) return;
  const long long i = e % n;
  const double pos = (double)i * ra;
  const long long i0 = (long long)pos;
  const float fr = (float)(pos - (double)i0);
  float av = 0.f;
  if (i0 + 1 < n_a) av = a[i0] * (1.f - fr) + a[i0 + 1] * fr;
  else if (i0 < n_a) av = a[i0];
  out[e] = ga * av + gb * b[e] + gc * c[e];
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

int b200vc_mix3(const float* a_mono, int64_t n_a, double ratio_a, const float* b, const float* c, float* out, int64_t n,
                float ga, float gb, float gc, void* stream) {
  B200VC_REQUIRE(a_mono && b && c && out && n > 0 && n_a > 0, "mix3: bad args");
  mix3_kernel<<<blocks_for(2 * n, 256), 256, 0, (cudaStream_t)stream>>>(a_mono, n_a, ratio_a, b, c, out, n, ga, gb, gc);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

}  // extern "C"
