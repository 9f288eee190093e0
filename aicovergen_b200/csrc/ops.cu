// Coalesced fp32 row/elementwise kernels around the tap-GEMMs: normalisation, softmax
// (+ VITS relative-position terms), gates, gathers, NSF sine source, conv_post.
// All HBM-bound; one pass over the data each, float4 where rows allow.
#include <cuda_fp16.h>
#include "common.cuh"
#include "../../include/b200vc.h"

namespace b200vc {
namespace {

// ---------------------------------------------------------------------------
// LayerNorm over the last dim, optional residual add, one warp per row.
//   out[r,:] = LN(x[r,:] + res[r,:]) * gamma + beta       (modules.py:29-32, fairseq LayerNorm)
// ---------------------------------------------------------------------------
template <int MAXV>   // MAXV * 32 >= C
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float* __restrict__ out, long long rows, int C, long long ldx,
                                 long long ldr, long long ldo, float eps, int round_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* xr = x + (long long)warp * ldx;
  const float* rr = res ? res + (long long)warp * ldr : nullptr;
  float v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    float t = 0.f;
    if (c < C) {
      t = xr[c];
      if (rr) t += rr[c];
    }
    v[i] = t;
    s += t;
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    const float d = (c < C) ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  float* orow = out + (long long)warp * ldo;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < C) {
      float y = (v[i] - mean) * rstd * gamma[c] + beta[c];
      orow[c] = round_out ? round_tf32(y) : y;
    }
  }
}

// ---------------------------------------------------------------------------
// Row softmax in place, optional VITS relative-key bias (attentions.py:238-261).
// One block per (head,row); row staged in dynamic shared memory.
// ---------------------------------------------------------------------------
__global__ void softmax_rows_kernel(float* __restrict__ S, int T, long long ld, long long head_stride,
                                    int rows_per_head, const float* __restrict__ q, int ldq,
                                    const float* __restrict__ emb_k, int window, int dk, int round_out, int row0) {
  extern __shared__ float row[];
  __shared__ float red[32];
  __shared__ float rel[64];
  const int head = blockIdx.x / rows_per_head;
  const int i = blockIdx.x % rows_per_head;
  float* s = S + (long long)head * head_stride + (long long)i * ld;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (emb_k) {
    const int nrel = 2 * window + 1;
    for (int r = tid; r < nrel; r += nt) {
      const float* qi = q + (long long)(i + row0) * ldq + head * dk;
      const float* e = emb_k + r * dk;
      float a = 0.f;
      for (int d = 0; d < dk; ++d) a = fmaf(qi[d], e[d], a);
      rel[r] = a;
    }
    __syncthreads();
  }
  float m = -INFINITY;
  for (int j = tid; j < T; j += nt) {
    float v = s[j];
    if (emb_k) {
      const int r = j - (i + row0) + window;
      if (r >= 0 && r <= 2 * window) v += rel[r];
    }
    row[j] = v;
    m = fmaxf(m, v);
  }
  m = warp_max(m);
  if ((tid & 31) == 0) red[tid >> 5] = m;
  __syncthreads();
  if (tid < 32) {
    float t = (tid < (nt >> 5)) ? red[tid] : -INFINITY;
    t = warp_max(t);
    if (tid == 0) red[0] = t;
  }
  __syncthreads();
  m = red[0];
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j < T; j += nt) {
    const float e = expf(row[j] - m);
    row[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  if (tid < 32) {
    float t = (tid < (nt >> 5)) ? red[tid] : 0.f;
    t = warp_sum(t);
    if (tid == 0) red[0] = t;
  }
  __syncthreads();
  const float inv = 1.f / red[0];
  for (int j = tid; j < T; j += nt) {
    const float p = row[j] * inv;
    s[j] = round_out ? round_tf32(p) : p;
  }
}

// out[i, h*dk + d] += sum_r P[h, i, i + r - W] * emb_v[r, d]   (attentions.py:264-271)
__global__ void relpos_value_add_kernel(float* __restrict__ out, int ldo, const float* __restrict__ P,
                                        int T, long long ld, long long head_stride,
                                        const float* __restrict__ emb_v, int window, int dk, int heads, int row0) {
  const int i = blockIdx.x;
  for (int c = threadIdx.x; c < heads * dk; c += blockDim.x) {
    const int h = c / dk, d = c % dk;
    const float* p = P + (long long)h * head_stride + (long long)i * ld;
    float a = 0.f;
    for (int r = 0; r <= 2 * window; ++r) {
      const int j = i + row0 + r - window;
      if (j >= 0 && j < T) a = fmaf(p[j], emb_v[r * dk + d], a);
    }
    out[(long long)(i + row0) * ldo + c] += a;
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ table, const long long* __restrict__ idx,
                                   float* __restrict__ out, long long rows, int C) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * C) return;
  const long long r = e / C;
  const int c = (int)(e % C);
  out[e] = table[idx[r] * C + c];
}

// acts[t,c] = tanh(a[t,c]) * sigmoid(a[t,C+c])   (commons.py:105-112)
__global__ void gate_kernel(const float* __restrict__ a, float* __restrict__ out, long long rows, int C,
                            int round_out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * C) return;
  const long long r = e / C;
  const int c = (int)(e % C);
  const float x = a[r * 2 * C + c], y = a[r * 2 * C + C + c];
  const float v = tanhf(x) * (1.f / (1.f + expf(-y)));
  out[e] = round_out ? round_tf32(v) : v;
}

// z[t,c] = m[t,c] + exp(logs[t,c]) * noise[c*P + t] * scale   (models.py:748; stats = [m | logs] per row)
__global__ void zp_sample_kernel(const float* __restrict__ stats, const float* __restrict__ noise,
                                 float* __restrict__ z, long long P, int C, float scale) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P * C) return;
  const long long t = e / C;
  const int c = (int)(e % C);
  z[e] = stats[t * 2 * C + c] + expf(stats[t * 2 * C + C + c]) * noise[(long long)c * P + t] * scale;
}

__global__ void axpby_kernel(const float* __restrict__ a, const float* __restrict__ b,
                             float* __restrict__ out, long long n, float alpha, float beta) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[e] = alpha * a[e] + (b ? beta * b[e] : 0.f);
}

// ---------------------------------------------------------------------------
// NSF harmonic source (models.py:320-370, 414-419), harmonic_num = 0.
// The reference's sample-level cumsum only ever removes integers from the phase
// (SURVEY.md B2), so the phase is evaluated in closed form per frame in fp64:
//   phi(t*upp+i) = frac(upp * sum_{tau<t} r_tau + (i+1) * r_t),  r = (f0/sr) mod 1
// ---------------------------------------------------------------------------
__global__ void nsf_frame_prefix_kernel(const float* __restrict__ f0, double* __restrict__ cum,
                                        int T, float sr) {
  // single block: per-thread serial chunk + block scan of chunk totals (T <= ~1e5)
  extern __shared__ double part[];
  const int nt = blockDim.x, tid = threadIdx.x;
  const int per = (T + nt - 1) / nt;
  const int b = tid * per, e = min(T, b + per);
  double s = 0.0;
  for (int t = b; t < e; ++t) s += (double)fmodf(f0[t] / sr, 1.0f);
  part[tid] = s;
  __syncthreads();
  if (tid == 0) {
    double run = 0.0;
    for (int k = 0; k < nt; ++k) {
      const double v = part[k];
      part[k] = run;
      run += v;
    }
  }
  __syncthreads();
  double run = part[tid];
  for (int t = b; t < e; ++t) {
    cum[t] = run;   // exclusive prefix: sum_{tau < t}
    run += (double)fmodf(f0[t] / sr, 1.0f);
  }
}

__global__ void nsf_source_kernel(const float* __restrict__ f0, const double* __restrict__ cum,
                                  const float* __restrict__ noise, float* __restrict__ har,
                                  long long L, int upp, float sr, float lin_w, float lin_b) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= L) return;
  const int t = (int)(n / upp);
  const int i = (int)(n % upp);
  const float f = f0[t];
  const float r = fmodf(f / sr, 1.0f);
  double ph = (double)upp * cum[t] + (double)(i + 1) * (double)r;
  ph -= floor(ph);
  const float uv = f > 0.f ? 1.f : 0.f;
  const float sine = sinf((float)(ph * 6.283185307179586)) * 0.1f;
  const float namp = uv * 0.003f + (1.f - uv) * (0.1f / 3.f);
  const float s = sine * uv + namp * noise[n];
  har[n] = tanhf(lin_w * s + lin_b);
}

// Strided 1-D convolution FROM ONE input channel: out[t, c] = res[t, c] + bias[c] + sum_j w[c, j] * src[src_off + t*stride + j]
// (samples outside [0, n_src) read as zero); out2 = act2(out).  K <= a few hundred MACs per output, 8-12 bytes of HBM
// traffic per output: write-bound, so it is a row kernel (one thread = 4 channels of one frame), not a GEMM — as 128-row
// tensor-core tiles with K = 1..80 the NSF noise convolutions ran at 12 % of HBM bandwidth.
// Replaces Conv1d(1, C, k, stride) at infer_pack/models.py:477-486 (noise_convs, called :505-506) and the first
// feature-extractor convolution of HuBERT (Cin = 1, k = 10, stride 5).
__global__ void conv1d_from1_kernel(const float* __restrict__ src, long long n_src, const float* __restrict__ w,
                                    const float* __restrict__ bias, const float* __restrict__ res, float* __restrict__ out,
                                    float* __restrict__ out2, long long T, int C, int K, int stride, long long src_off,
                                    int act2, float act2_p, int round_out2) {
  extern __shared__ float wt[];                 // [K][C]: transposed so that a thread's 4 channels are one float4
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) {
    const int c = i / K, j = i - c * K;
    wt[j * C + c] = w[i];
  }
  __syncthreads();
  const int c4 = C >> 2;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= T * c4) return;
  const long long t = idx / c4;
  const int c = (int)(idx - t * c4) << 2;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long s0 = src_off + t * stride;
  for (int j = 0; j < K; ++j) {
    const long long si = s0 + j;
    const float x = (si >= 0 && si < n_src) ? __ldg(src + si) : 0.f;
    const float4 ww = *reinterpret_cast<const float4*>(&wt[j * C + c]);
    acc.x = fmaf(x, ww.x, acc.x); acc.y = fmaf(x, ww.y, acc.y); acc.z = fmaf(x, ww.z, acc.z); acc.w = fmaf(x, ww.w, acc.w);
  }
  if (bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c));
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
  }
  const long long o = t * C + c;
  if (res) {
    const float4 r = *reinterpret_cast<const float4*>(res + o);
    acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
  }
  *reinterpret_cast<float4*>(out + o) = acc;
  if (out2) {
    float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = apply_act(v[i], act2, act2_p);
      if (round_out2 & 1) v[i] = round_tf32(v[i]);
    }
    if (round_out2 & 2) {      // bit1: out2 holds fp16 (4 channels = one 8-byte store)
      const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out2) + o) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    } else {
      *reinterpret_cast<float4*>(out2 + o) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}


// out[t] = act(sum_k sum_c w[k,c] * x[t + k - pad, c])  — conv_post (models.py:486,514-515), Cout = 1
__global__ void conv1d_to1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                  float* __restrict__ out, long long T, int C, int K, int pad, int act) {
  extern __shared__ float ws[];
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const long long tt = t + k - pad;
    if (tt < 0 || tt >= T) continue;
    const float4* xr = reinterpret_cast<const float4*>(x + tt * C);
    const float* wk = ws + k * C;
#pragma unroll 4
    for (int c4 = 0; c4 < C / 4; ++c4) {
      const float4 v = __ldg(xr + c4);
      acc = fmaf(v.x, wk[c4 * 4 + 0], acc);
      acc = fmaf(v.y, wk[c4 * 4 + 1], acc);
      acc = fmaf(v.z, wk[c4 * 4 + 2], acc);
      acc = fmaf(v.w, wk[c4 * 4 + 3], acc);
    }
  }
  out[t] = apply_act(acc, act, 0.f);
}

// elementwise activation (+ optional second output), in/out may alias
__global__ void act_kernel(const float* __restrict__ x, float* __restrict__ out, long long n, int act,
                           float p, int round_out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float v = apply_act(x[e], act, p);
  if (round_out & 2) reinterpret_cast<__half*>(out)[e] = __float2half_rn(v);      // bit1: `out` holds fp16
  else out[e] = (round_out & 1) ? round_tf32(v) : v;
}


// ---------------------------------------------------------------------------
// GroupNorm with one channel per group over the time axis (fairseq HuBERT conv layer 0:
// nn.GroupNorm(512, 512)) + GELU.  Pass 1: per-channel sum / sum-of-squares (fp64 atomics);
// pass 2: normalise, affine, GELU.
// ---------------------------------------------------------------------------
__global__ void colstats_kernel(const float* __restrict__ x, double* __restrict__ stats, long long rows, int C,
                                int rows_per_block) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  double s = 0.0, q = 0.0;
  for (long long r = r0; r < r1; ++r) {
    const double v = (double)x[r * C + c];
    s += v;
    q += v * v;
  }
  atomicAdd(stats + c, s);
  atomicAdd(stats + C + c, q);
}

__global__ void groupnorm_apply_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float* __restrict__ out, long long rows, int C, float eps, int act,
                                       int round_out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * C) return;
  const int c = (int)(e % C);
  const double mean = stats[c] / (double)rows;
  const double var = stats[C + c] / (double)rows - mean * mean;
  const float rstd = (float)(1.0 / sqrt(fmax(var, 0.0) + (double)eps));
  float v = ((x[e] - (float)mean) * rstd) * gamma[c] + beta[c];
  v = apply_act(v, act, 0.f);
  out[e] = round_out ? round_tf32(v) : v;
}

inline unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

}  // namespace
}  // namespace b200vc

using namespace b200vc;

extern "C" {

int b200vc_layernorm(const float* x, const float* res, const float* gamma, const float* beta, float* out,
                     int64_t rows, int C, int64_t ldx, int64_t ldr, int64_t ldo, float eps, int round_out,
                     void* stream) {
  B200VC_RECORD(b200vc_layernorm(x, res, gamma, beta, out, rows, C, ldx, ldr, ldo, eps, round_out, stream));
  B200VC_REQUIRE(x && gamma && beta && out && rows > 0 && C > 0 && C <= 1024, "layernorm: bad args (C=%d)", C);
  cudaStream_t s = (cudaStream_t)stream;
  const int bs = 256;
  const unsigned grid = blocks_for(rows * 32, bs);
  if (C <= 256)
    layernorm_kernel<8><<<grid, bs, 0, s>>>(x, res, gamma, beta, out, rows, C, ldx, ldr, ldo, eps, round_out);
  else if (C <= 512)
    layernorm_kernel<16><<<grid, bs, 0, s>>>(x, res, gamma, beta, out, rows, C, ldx, ldr, ldo, eps, round_out);
  else
    layernorm_kernel<32><<<grid, bs, 0, s>>>(x, res, gamma, beta, out, rows, C, ldx, ldr, ldo, eps, round_out);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_softmax_rows(float* S, int heads, int rows_per_head, int T, int64_t ld, int64_t head_stride,
                        const float* q, int ldq, const float* emb_rel_k, int window, int dk,
                        int round_out, int row0, void* stream) {
  B200VC_RECORD(b200vc_softmax_rows(S, heads, rows_per_head, T, ld, head_stride, q, ldq, emb_rel_k, window, dk, round_out, row0, stream));
  B200VC_REQUIRE(S && heads > 0 && rows_per_head > 0 && T > 0, "softmax_rows: bad args");
  B200VC_REQUIRE(T * 4 <= 200 * 1024, "softmax_rows: row of %d floats does not fit shared memory", T);
  B200VC_REQUIRE(!emb_rel_k || (q && 2 * window + 1 <= 64), "softmax_rows: bad relative-position args");
  cudaStream_t s = (cudaStream_t)stream;
  const int smem = T * 4;
  static int configured = 0;
  if (smem > 48 * 1024 && configured < smem) {
    B200VC_CHECK_CUDA(cudaFuncSetAttribute(softmax_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  softmax_rows_kernel<<<(unsigned)(heads * rows_per_head), 256, smem, s>>>(
      S, T, ld, head_stride, rows_per_head, q, ldq, emb_rel_k, window, dk, round_out, row0);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_relpos_value_add(float* out, int ldo, const float* P, int rows, int T, int64_t ld, int64_t head_stride,
                            const float* emb_rel_v, int window, int dk, int heads, int row0, void* stream) {
  B200VC_RECORD(b200vc_relpos_value_add(out, ldo, P, rows, T, ld, head_stride, emb_rel_v, window, dk, heads, row0, stream));
  B200VC_REQUIRE(out && P && emb_rel_v && T > 0 && rows > 0 && row0 >= 0, "relpos_value_add: bad args");
  relpos_value_add_kernel<<<(unsigned)rows, 192, 0, (cudaStream_t)stream>>>(out, ldo, P, T, ld, head_stride,
                                                                           emb_rel_v, window, dk, heads, row0);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_gather_rows(const float* table, const int64_t* idx, float* out, int64_t rows, int C, void* stream) {
  B200VC_RECORD(b200vc_gather_rows(table, idx, out, rows, C, stream));
  B200VC_REQUIRE(table && idx && out && rows > 0 && C > 0, "gather_rows: bad args");
  gather_rows_kernel<<<blocks_for(rows * C, 256), 256, 0, (cudaStream_t)stream>>>(
      table, reinterpret_cast<const long long*>(idx), out, rows, C);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_gate_tanh_sigmoid(const float* a, float* out, int64_t rows, int C, int round_out, void* stream) {
  B200VC_RECORD(b200vc_gate_tanh_sigmoid(a, out, rows, C, round_out, stream));
  B200VC_REQUIRE(a && out && rows > 0 && C > 0, "gate: bad args");
  gate_kernel<<<blocks_for(rows * C, 256), 256, 0, (cudaStream_t)stream>>>(a, out, rows, C, round_out);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_zp_sample(const float* stats, const float* noise, float* z, int64_t P, int C, float scale, void* stream) {
  B200VC_RECORD(b200vc_zp_sample(stats, noise, z, P, C, scale, stream));
  B200VC_REQUIRE(stats && noise && z && P > 0 && C > 0, "zp_sample: bad args");
  zp_sample_kernel<<<blocks_for(P * C, 256), 256, 0, (cudaStream_t)stream>>>(stats, noise, z, P, C, scale);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_axpby(const float* a, const float* b, float* out, int64_t n, float alpha, float beta, void* stream) {
  B200VC_RECORD(b200vc_axpby(a, b, out, n, alpha, beta, stream));
  B200VC_REQUIRE(a && out && n > 0, "axpby: bad args");
  axpby_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, out, n, alpha, beta);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_act(const float* x, float* out, int64_t n, int act, float p, int round_out, void* stream) {
  B200VC_RECORD(b200vc_act(x, out, n, act, p, round_out, stream));
  B200VC_REQUIRE(x && out && n > 0, "act: bad args");
  act_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, out, n, act, p, round_out);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_nsf_source(const float* f0, const float* noise, float* har, double* scratch_cum, int T, int upp,
                      float sr, float lin_w, float lin_b, void* stream) {
  B200VC_RECORD(b200vc_nsf_source(f0, noise, har, scratch_cum, T, upp, sr, lin_w, lin_b, stream));
  B200VC_REQUIRE(f0 && noise && har && scratch_cum && T > 0 && upp > 0, "nsf_source: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  nsf_frame_prefix_kernel<<<1, 1024, 1024 * sizeof(double), s>>>(f0, scratch_cum, T, sr);
  const long long L = (long long)T * upp;
  nsf_source_kernel<<<blocks_for(L, 256), 256, 0, s>>>(f0, scratch_cum, noise, har, L, upp, sr, lin_w, lin_b);
  count_launch(2);
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_conv1d_from1(const float* src, int64_t n_src, const float* w, const float* bias, const float* res, float* out,
                        float* out2, int64_t T, int C, int K, int stride, int64_t src_off, int act2, float act2_p,
                        int round_out2, void* stream) {
  B200VC_RECORD(b200vc_conv1d_from1(src, n_src, w, bias, res, out, out2, T, C, K, stride, src_off, act2, act2_p, round_out2, stream));
  B200VC_REQUIRE(src && w && out && T > 0 && C > 0 && C % 4 == 0 && K > 0 && stride > 0 && (long long)K * C * 4 <= 160 * 1024,
                 "conv1d_from1: bad args (C=%d K=%d)", C, K);
  const size_t smem = (size_t)K * C * 4;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    B200VC_CHECK_CUDA(cudaFuncSetAttribute(conv1d_from1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = 160 * 1024;
  }
  conv1d_from1_kernel<<<blocks_for(T * (C / 4), 256), 256, smem, (cudaStream_t)stream>>>(
      src, n_src, w, bias, res, out, out2, T, C, K, stride, src_off, act2, act2_p, round_out2);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_conv1d_to1(const float* x, const float* w, float* out, int64_t T, int C, int K, int pad, int act,
                      void* stream) {
  B200VC_RECORD(b200vc_conv1d_to1(x, w, out, T, C, K, pad, act, stream));
  B200VC_REQUIRE(x && w && out && T > 0 && C % 4 == 0 && K * C * 4 <= 48 * 1024, "conv1d_to1: bad args");
  conv1d_to1_kernel<<<blocks_for(T, 256), 256, K * C * 4, (cudaStream_t)stream>>>(x, w, out, T, C, K, pad, act);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_groupnorm_time(const float* x, const float* gamma, const float* beta, float* out, double* stats,
                          int64_t rows, int C, float eps, int act, int round_out, void* stream) {
  B200VC_RECORD(b200vc_groupnorm_time(x, gamma, beta, out, stats, rows, C, eps, act, round_out, stream));
  B200VC_REQUIRE(x && gamma && beta && out && stats && rows > 0 && C > 0, "groupnorm_time: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  B200VC_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C, s));
  const int rpb = 256;
  dim3 grid(blocks_for(rows, rpb), blocks_for(C, 128));
  colstats_kernel<<<grid, 128, 0, s>>>(x, stats, rows, C, rpb);
  groupnorm_apply_kernel<<<blocks_for(rows * C, 256), 256, 0, s>>>(x, stats, gamma, beta, out, rows, C, eps, act, round_out);
  count_launch(2);
  B200VC_LAUNCH_CHECK();
  return kOk;
}

}  // extern "C"
