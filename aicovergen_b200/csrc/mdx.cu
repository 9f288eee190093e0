// MDX-Net pass helpers (src/mdx.py): chunk gathering with the pad_wave zero regions and the STFT reflect
// padding fused, NHWC <-> NHCW tiled transposes around the TDF (frequency-axis linear) GEMMs, and the
// iSTFT overlap-add + window-envelope normalisation + trim fused with the scatter into the song buffer.
#include <cuda_fp16.h>
#include "common.cuh"
#include "../../include/b200vc.h"

namespace b200vc {
namespace {

inline unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// out[b, ch, j] = sign * wave[ch, src_start[b] + reflect(j - half)]  if that song index lies in [lo[b], hi[b]) else 0
//   j in [0, chunk + 2*half).  Reproduces pad_wave's zero padding (mdx.py:156-171) followed by torch.stft's
//   reflect padding of each chunk (mdx.py:39).
__global__ void mdx_gather_chunks_kernel(const float* __restrict__ wave, long long n_song,
                                         const long long* __restrict__ src_start, const long long* __restrict__ lo,
                                         const long long* __restrict__ hi, float* __restrict__ out, int B, int chunk,
                                         int half, float sign, int round_out) {
  const long long plen = (long long)chunk + 2 * half;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)B * 2 * plen) return;
  const long long j = e % plen;
  const int ch = (int)((e / plen) % 2);
  const int b = (int)(e / (2 * plen));
  long long s = j - half;
  if (s < 0) s = -s;
  if (s >= chunk) s = 2 * ((long long)chunk - 1) - s;
  const long long idx = src_start[b] + s;
  float v = 0.f;
  if (idx >= lo[b] && idx < hi[b] && idx >= 0 && idx < n_song) v = sign * wave[(long long)ch * n_song + idx];
  out[e] = round_out ? round_tf32(v) : v;
}

// First 1x1 convolution of the TFC-TDF net (4 -> g channels, BatchNorm folded, ReLU), reading the DFT-friendly
// spectrogram layout spec[B][2 ch][T][F][2 ri] and writing NHWC x[B][T][F][g].  Pointwise and write-bound: one thread
// per (pixel, 4-channel group) so every store instruction covers contiguous 16-byte pieces of one NHWC row.
__global__ void mdx_first_conv_kernel(const float* __restrict__ spec, const float* __restrict__ w4,
                                      const float* __restrict__ bias, float* __restrict__ out, long long npix,
                                      long long TF, int g4, int round_out, int out_half) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= npix * g4) return;
  const long long pix = idx / g4;
  const int cg = (int)(idx - pix * g4);
  const long long b = pix / TF, tf = pix - b * TF;
  const float2 x0 = __ldg(reinterpret_cast<const float2*>(spec + ((b * 2 + 0) * TF + tf) * 2));
  const float2 x1 = __ldg(reinterpret_cast<const float2*>(spec + ((b * 2 + 1) * TF + tf) * 2));
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(w4) + cg * 4 + j);   // (ch0 re, ch0 im, ch1 re, ch1 im)
    float a = 0.f;
    a = fmaf(x0.x, w.x, a); a = fmaf(x0.y, w.y, a); a = fmaf(x1.x, w.z, a); a = fmaf(x1.y, w.w, a);
    a = fmaxf(a + __ldg(bias + cg * 4 + j), 0.f);
    v[j] = round_out ? round_tf32(a) : a;
  }
  if (out_half) {      // fp16 activation storage: 4 channels = one 8-byte store
    const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
    reinterpret_cast<uint2*>(out)[idx] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
  } else {
    reinterpret_cast<float4*>(out)[idx] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// Final 1x1 convolution (c -> 4 channels + bias): NHWC x[B][T][F][c] -> spec[B][2 ch][T][F][2 ri].  Read-bound: one
// thread per pixel streams its c contiguous floats (float4) and keeps the four dot products in registers.
__global__ void mdx_final_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                      float* __restrict__ spec, long long npix, long long TF, int c, int round_out, int x_half) {
  extern __shared__ float ws[];            // [4][c]
  for (int i = threadIdx.x; i < 4 * c; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const float4* xp = reinterpret_cast<const float4*>(x + pix * c);
  const uint2* xh = reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(x) + pix * c);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int k = 0; k < c; k += 4) {
    float4 t;
    if (x_half) {
      const uint2 u = __ldg(xh + (k >> 2));
      const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
      t = make_float4(f0.x, f0.y, f1.x, f1.y);
    } else {
      t = __ldg(xp + (k >> 2));
    }
    a0 = fmaf(t.x, ws[k], a0); a0 = fmaf(t.y, ws[k + 1], a0); a0 = fmaf(t.z, ws[k + 2], a0); a0 = fmaf(t.w, ws[k + 3], a0);
    a1 = fmaf(t.x, ws[c + k], a1); a1 = fmaf(t.y, ws[c + k + 1], a1); a1 = fmaf(t.z, ws[c + k + 2], a1); a1 = fmaf(t.w, ws[c + k + 3], a1);
    a2 = fmaf(t.x, ws[2 * c + k], a2); a2 = fmaf(t.y, ws[2 * c + k + 1], a2); a2 = fmaf(t.z, ws[2 * c + k + 2], a2); a2 = fmaf(t.w, ws[2 * c + k + 3], a2);
    a3 = fmaf(t.x, ws[3 * c + k], a3); a3 = fmaf(t.y, ws[3 * c + k + 1], a3); a3 = fmaf(t.z, ws[3 * c + k + 2], a3); a3 = fmaf(t.w, ws[3 * c + k + 3], a3);
  }
  a0 += __ldg(bias + 0); a1 += __ldg(bias + 1); a2 += __ldg(bias + 2); a3 += __ldg(bias + 3);
  if (round_out) { a0 = round_tf32(a0); a1 = round_tf32(a1); a2 = round_tf32(a2); a3 = round_tf32(a3); }
  const long long b = pix / TF, tf = pix - b * TF;
  reinterpret_cast<float2*>(spec)[(b * 2 + 0) * TF + tf] = make_float2(a0, a1);
  reinterpret_cast<float2*>(spec)[(b * 2 + 1) * TF + tf] = make_float2(a2, a3);
}

// x [R, W, C] -> out [R, C, W] with per-channel scale (R = B*H): 32x32 shared-memory tiles
__global__ void nhwc_to_nhcw_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                    float* __restrict__ out, int W, int C, int round_out) {
  __shared__ float tile[32][33];
  const long long r = blockIdx.z;
  const int w0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* xr = x + r * (long long)W * C;
  float* orow = out + r * (long long)W * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int w = w0 + i, c = c0 + threadIdx.x;
    if (w < W && c < C) tile[i][threadIdx.x] = xr[(long long)w * C + c] * (scale ? scale[c] : 1.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, w = w0 + threadIdx.x;
    if (w < W && c < C) {
      const float v = tile[threadIdx.x][i];
      orow[(long long)c * W + w] = round_out ? round_tf32(v) : v;
    }
  }
}

// out[r, w, c] = x[r, w, c] + t[r, c, w]
__global__ void nhcw_to_nhwc_add_kernel(const float* __restrict__ t, const float* __restrict__ x,
                                        float* __restrict__ out, int W, int C, int round_out) {
  __shared__ float tile[32][33];
  const long long r = blockIdx.z;
  const int w0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* tr = t + r * (long long)W * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, w = w0 + threadIdx.x;
    if (w < W && c < C) tile[i][threadIdx.x] = tr[(long long)c * W + w];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int w = w0 + i, c = c0 + threadIdx.x;
    if (w < W && c < C) {
      const long long o = r * (long long)W * C + (long long)w * C + c;
      const float v = x[o] + tile[threadIdx.x][i];
      out[o] = round_out ? round_tf32(v) : v;
    }
  }
}

// iSTFT tail (torch.istft, mdx.py:53): frames [B,2,T,n_fft] already hold irfft(X)*window.
//   y[s] = sum_t frames[t, s + half - t*hop] / env[s + half], s in [trim, chunk - trim)
// The kept part of chunk b is written at song[ch, dst_start[b] + (s - trim)] if that index is in
// [keep_lo[b], keep_hi[b]);  song = song * 1 + coef * y  (accumulate: the denoise pass adds two sweeps).
__global__ void mdx_ola_store_kernel(const float* __restrict__ frames, const float* __restrict__ env,
                                     const long long* __restrict__ dst_start, const long long* __restrict__ keep_lo,
                                     const long long* __restrict__ keep_hi, float* __restrict__ song, long long n_song,
                                     int B, int T, int n_fft, int hop, int chunk, int trim, float coef, int accumulate) {
  const int gen = chunk - 2 * trim;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)B * 2 * gen) return;
  const int k = (int)(e % gen);
  const int ch = (int)((e / gen) % 2);
  const int b = (int)(e / (2LL * gen));
  const long long dst = dst_start[b] + k;
  if (dst < keep_lo[b] || dst >= keep_hi[b] || dst < 0 || dst >= n_song) return;
  const int half = n_fft / 2;
  const int j = k + trim + half;                 // coordinate in the centre-padded signal
  // frames t with 0 <= j - t*hop < n_fft
  int t_hi = j / hop;
  if (t_hi > T - 1) t_hi = T - 1;
  int t_lo = (j - n_fft + hop) / hop;            // ceil((j - n_fft + 1)/hop) for positive numerators
  if (j - n_fft + 1 <= 0) t_lo = 0;
  const float* f = frames + ((long long)b * 2 + ch) * T * n_fft;
  float acc = 0.f;
  for (int t = t_lo; t <= t_hi; ++t) acc += f[(long long)t * n_fft + (j - t * hop)];
  const float y = acc / env[j];
  float* o = song + (long long)ch * n_song + dst;
  *o = accumulate ? (*o + coef * y) : coef * y;
}

// inverse stem (mdx.py:280): inv = -proc * compensation + wave_norm ; main *= peak happens before (mdx.py:267)
__global__ void mdx_finalize_kernel(float* __restrict__ proc, const float* __restrict__ wave_norm,
                                    float* __restrict__ inverse, long long n, float peak, float compensation) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float p = proc[e] * peak;
  proc[e] = p;
  if (inverse) inverse[e] = (-p * compensation) + wave_norm[e];
}

}  // namespace
}  // namespace b200vc

using namespace b200vc;

extern "C" {

int b200vc_mdx_gather_chunks(const float* wave, int64_t n_song, const int64_t* src_start, const int64_t* lo,
                             const int64_t* hi, float* out, int B, int chunk, int half, float sign, int round_out,
                             void* stream) {
  B200VC_RECORD(b200vc_mdx_gather_chunks(wave, n_song, src_start, lo, hi, out, B, chunk, half, sign, round_out, stream));
  B200VC_REQUIRE(wave && src_start && lo && hi && out && B > 0 && chunk > half && half >= 0, "mdx_gather_chunks: bad args");
  const long long n = (long long)B * 2 * ((long long)chunk + 2 * half);
  mdx_gather_chunks_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
      wave, n_song, reinterpret_cast<const long long*>(src_start), reinterpret_cast<const long long*>(lo),
      reinterpret_cast<const long long*>(hi), out, B, chunk, half, sign, round_out);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_mdx_first_conv(const float* spec, const float* w4, const float* bias, float* out, int B, int T, int F, int g,
                          int round_out, int out_half, void* stream) {
  B200VC_RECORD(b200vc_mdx_first_conv(spec, w4, bias, out, B, T, F, g, round_out, out_half, stream));
  B200VC_REQUIRE(spec && w4 && bias && out && B > 0 && T > 0 && F > 0 && g > 0 && g % 4 == 0, "mdx_first_conv: bad args (g=%d)", g);
  B200VC_REQUIRE(((uintptr_t)spec % 8 == 0) && ((uintptr_t)w4 % 16 == 0) && ((uintptr_t)out % 16 == 0), "mdx_first_conv: alignment");
  const long long npix = (long long)B * T * F;
  mdx_first_conv_kernel<<<blocks_for(npix * (g / 4), 256), 256, 0, (cudaStream_t)stream>>>(spec, w4, bias, out, npix, (long long)T * F,
                                                                                          g / 4, round_out, out_half);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_mdx_final_conv(const float* x, const float* w, const float* bias, float* spec, int B, int T, int F, int c,
                          int round_out, int x_half, void* stream) {
  B200VC_RECORD(b200vc_mdx_final_conv(x, w, bias, spec, B, T, F, c, round_out, x_half, stream));
  B200VC_REQUIRE(x && w && bias && spec && B > 0 && T > 0 && F > 0 && c > 0 && c % 4 == 0 && c <= 2048, "mdx_final_conv: bad args (c=%d)", c);
  B200VC_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)spec % 8 == 0), "mdx_final_conv: alignment");
  const long long npix = (long long)B * T * F;
  mdx_final_conv_kernel<<<blocks_for(npix, 128), 128, 4 * c * sizeof(float), (cudaStream_t)stream>>>(x, w, bias, spec, npix,
                                                                                                  (long long)T * F, c, round_out, x_half);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_nhwc_to_nhcw(const float* x, const float* scale, float* out, int64_t R, int W, int C, int round_out,
                        void* stream) {
  B200VC_RECORD(b200vc_nhwc_to_nhcw(x, scale, out, R, W, C, round_out, stream));
  B200VC_REQUIRE(x && out && R > 0 && R < 65536 && W > 0 && C > 0, "nhwc_to_nhcw: bad args (R=%lld)", (long long)R);
  dim3 grid((W + 31) / 32, (C + 31) / 32, (unsigned)R), block(32, 8);
  nhwc_to_nhcw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, scale, out, W, C, round_out);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_nhcw_to_nhwc_add(const float* t, const float* x, float* out, int64_t R, int W, int C, int round_out,
                            void* stream) {
  B200VC_RECORD(b200vc_nhcw_to_nhwc_add(t, x, out, R, W, C, round_out, stream));
  B200VC_REQUIRE(t && x && out && R > 0 && R < 65536 && W > 0 && C > 0, "nhcw_to_nhwc_add: bad args");
  dim3 grid((W + 31) / 32, (C + 31) / 32, (unsigned)R), block(32, 8);
  nhcw_to_nhwc_add_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(t, x, out, W, C, round_out);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_mdx_ola_store(const float* frames, const float* env, const int64_t* dst_start, const int64_t* keep_lo,
                         const int64_t* keep_hi, float* song, int64_t n_song, int B, int T, int n_fft, int hop,
                         int chunk, int trim, float coef, int accumulate, void* stream) {
  B200VC_RECORD(b200vc_mdx_ola_store(frames, env, dst_start, keep_lo, keep_hi, song, n_song, B, T, n_fft, hop, chunk, trim, coef, accumulate, stream));
  B200VC_REQUIRE(frames && env && dst_start && keep_lo && keep_hi && song && B > 0 && chunk > 2 * trim,
                 "mdx_ola_store: bad args");
  const long long n = (long long)B * 2 * (chunk - 2 * trim);
  mdx_ola_store_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
      frames, env, reinterpret_cast<const long long*>(dst_start), reinterpret_cast<const long long*>(keep_lo),
      reinterpret_cast<const long long*>(keep_hi), song, n_song, B, T, n_fft, hop, chunk, trim, coef, accumulate);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_mdx_finalize(float* proc, const float* wave_norm, float* inverse, int64_t n, float peak,
                        float compensation, void* stream) {
  B200VC_RECORD(b200vc_mdx_finalize(proc, wave_norm, inverse, n, peak, compensation, stream));
  B200VC_REQUIRE(proc && n > 0 && (!inverse || wave_norm), "mdx_finalize: bad args");
  mdx_finalize_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(proc, wave_norm, inverse, n, peak,
                                                                          compensation);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

}  // extern "C"
