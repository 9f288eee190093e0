// Vocal effects chain and final mix of the cover (reference src/main.py:206-233):
//   add_audio_effects: pedalboard HighpassFilter() -> Compressor(ratio 4, threshold -15 dB) -> Reverb(room, dry, wet, damping)
//   combine_audio    : pydub gain, overlay, export
// pedalboard wraps JUCE's dsp::IIR first-order high-pass, dsp::Compressor (ballistics filter + VCA) and juce::Reverb (Freeverb:
// 8 combs + 4 all-passes); pydub's sample arithmetic is CPython's audioop (mul, ratecv, tostereo, add).  All of it is
// sequential recursive code on the CPU; here each recursion is restructured so that a GPU can run it:
//   * high-pass + envelope follower: both forget their state geometrically -> the signal is cut into chunks, each chunk is
//     recomputed from `warm` samples earlier with zero state (error < 1e-10 of full scale) by its own thread; the compressor
//     gain (a pow per sample) is an element-wise pass over the stored envelope;
//   * comb filter j (delay D_j): y[t] = g x[t] + fb * last[t], last[t] = (1-damp) y[t-D] + damp last[t-1].  Inside one block
//     of D samples every y[t-D-k] is already known, and `last` forgets like damp^k: each run of 8 consecutive samples starts
//     `last` with the K-term Horner form of the one-pole low-pass (the same operation order as the sequential loop, started
//     K samples earlier) and then advances it sequentially; one thread block per comb walks the signal in steps of D samples;
//   * all-pass (delay D, gain 0.5): w[t] = x[t] + 0.5 w[t-D] = sum_m 0.5^m x[t-mD] — 32 terms, fully parallel;
//   * the mix is integer arithmetic per output sample; audioop.ratecv (linear interpolation on an integer phase accumulator)
//     has the closed form written at src_sample() below — bit-exact against CPython's audioop (tests/test_effects_cpu.py pins
//     the closed form, tests/test_effects_gpu.py the kernel).
// Every float operation that the C code performs unfused is written with __f*_rn so that nvcc cannot contract it.
#include <climits>

#include "../../include/b200vc.h"
#include "common.cuh"

namespace b200vc {
namespace {

// ---- int16 -> float (JUCE reader: sample * 2^-15), first-order high-pass (TDF-II), peak ballistics filter ----
struct HpfEnv {
  float b0, b1, a1, cte_at, cte_rl;
  float lv1, yold;
  __device__ __forceinline__ float step(float in, float& env_out) {
    const float out = __fadd_rn(__fmul_rn(in, b0), lv1);
    lv1 = __fsub_rn(__fmul_rn(in, b1), __fmul_rn(out, a1));
    const float a = fabsf(out);
    const float cte = (a > yold) ? cte_at : cte_rl;
    const float env = __fadd_rn(a, __fmul_rn(cte, __fsub_rn(yold, a)));
    yold = env;
    env_out = env;
    return out;
  }
};

// One thread per chunk (the two recurrences only: ~25 dependent cycles per sample).  chunk and warm are multiples of 8 and
// x / y / env are 16-byte aligned, so a thread walks its stream in groups of 8 samples: one 128-bit load (issued one group ahead
// of its use), 128-bit stores.  y = high-pass output, env = envelope; the VCA gain is applied by fx_gain_kernel.
__global__ void fx_hpf_env_kernel(const int16_t* __restrict__ x, float* __restrict__ y, float* __restrict__ envo, long long n,
                                  int chunk, int warm, HpfEnv f) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long lo = c * chunk;
  if (lo >= n) return;
  const long long hi = (lo + chunk < n) ? lo + chunk : n;
  long long t = (lo > warm) ? lo - warm : 0;
  f.lv1 = 0.f;
  f.yold = 0.f;
  const float sc = 1.0f / 32768.0f;
  uint4 nxt = make_uint4(0, 0, 0, 0);
  if (t + 8 <= hi) nxt = *reinterpret_cast<const uint4*>(x + t);
  for (; t + 8 <= hi; t += 8) {
    const uint4 cur = nxt;
    if (t + 16 <= hi) nxt = *reinterpret_cast<const uint4*>(x + t + 8);
    const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
    float o[8], e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int16_t sv = (int16_t)((i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xffffu));
      o[i] = f.step((float)sv * sc, e[i]);
    }
    if (t >= lo) {
      *reinterpret_cast<float4*>(y + t) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(y + t + 4) = make_float4(o[4], o[5], o[6], o[7]);
      *reinterpret_cast<float4*>(envo + t) = make_float4(e[0], e[1], e[2], e[3]);
      *reinterpret_cast<float4*>(envo + t + 4) = make_float4(e[4], e[5], e[6], e[7]);
    }
  }
  for (; t < hi; ++t) {                                     // tail of the last chunk
    float e;
    const float out = f.step((float)x[t] * sc, e);
    if (t >= lo) {
      y[t] = out;
      envo[t] = e;
    }
  }
}

// Compressor VCA: y *= env < thr ? 1 : (env / thr) ^ (1/ratio - 1)
__global__ void fx_gain_kernel(float* __restrict__ y, const float* __restrict__ env, long long n, float thr, float thr_inv,
                               float expo) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float e = env[t];
  const float g = (e < thr) ? 1.0f : powf(__fmul_rn(e, thr_inv), expo);
  y[t] = __fmul_rn(g, y[t]);
}

struct CombDelays {
  int d[8];
};

// One thread block per comb filter; Y[j][t] = the value the comb writes into its delay line at time t.  The last R samples of
// the delay line live in a shared-memory ring (R = power of two >= 2 D + K, zero-initialised = the empty delay line; one pad
// word per 32 so that the stride-8 accesses below are conflict-free).  The block walks the signal in steps of D samples with
// one barrier per step; inside a step thread j owns the 8 CONSECUTIVE samples t = s + 8j .. s + 8j + 7: the damping low-pass
// `last` is started with the K-term Horner form at the first of them and then advanced sequentially, exactly like the
// reference loop.  Inputs are loaded one step ahead.
constexpr int kCombRun = 8;
__device__ __forceinline__ unsigned ring_at(unsigned u, unsigned mask) {
  const unsigned r = u & mask;
  return r + (r >> 5);
}

__global__ void __launch_bounds__(256) fx_comb_kernel(const float* __restrict__ x, float* __restrict__ Y, long long n,
                                                      CombDelays dl, int ring_size, float gain, float damp, float omd, float fb,
                                                      int K) {
  extern __shared__ float ring[];
  const int D = dl.d[blockIdx.x];
  const unsigned mask = (unsigned)ring_size - 1u;
  float* Yj = Y + (long long)blockIdx.x * n;
  for (int i = threadIdx.x; i < ring_size + (ring_size >> 5); i += blockDim.x) ring[i] = 0.f;
  __syncthreads();
  const int i0 = threadIdx.x * kCombRun;                    // first sample of this thread inside a step
  float xn[kCombRun];
#pragma unroll
  for (int r = 0; r < kCombRun; ++r) xn[r] = (i0 + r < D && i0 + r < n) ? x[i0 + r] : 0.f;
  for (long long s = 0; s < n; s += D) {
    float xc[kCombRun];
#pragma unroll
    for (int r = 0; r < kCombRun; ++r) {
      xc[r] = xn[r];
      const long long nx = s + D + i0 + r;
      xn[r] = (i0 + r < D && nx < n) ? x[nx] : 0.f;
    }
    if (i0 < D && s + i0 < n) {
      const long long t0 = s + i0;
      const unsigned u0 = (unsigned)(t0 - D);               // ring arithmetic is modulo 2^32, ring_size divides it
      float last = 0.f;
      for (int k = K - 1; k >= 1; --k) last = __fadd_rn(__fmul_rn(ring[ring_at(u0 - (unsigned)k, mask)], omd), __fmul_rn(last, damp));
#pragma unroll
      for (int r = 0; r < kCombRun; ++r) {
        const long long t = t0 + r;
        if (i0 + r < D && t < n) {
          last = __fadd_rn(__fmul_rn(ring[ring_at(u0 + (unsigned)r, mask)], omd), __fmul_rn(last, damp));
          const float y = __fadd_rn(__fmul_rn(xc[r], gain), __fmul_rn(last, fb));
          ring[ring_at((unsigned)t, mask)] = y;
          Yj[t] = y;
        }
      }
    }
    __syncthreads();
  }
}

// sum of the eight comb outputs (comb j returns the delayed value Y[j][t - D_j]), accumulated in the reference's order
__global__ void fx_comb_sum_kernel(const float* __restrict__ Y, long long n, CombDelays dl, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long u = t - dl.d[j];
    acc = __fadd_rn(acc, (u >= 0) ? Y[(long long)j * n + u] : 0.f);
  }
  out[t] = acc;
}

__global__ void fx_allpass_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int D, int M) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float acc = 0.f;                                  // -> w[t - D]
  for (int m = M - 1; m >= 0; --m) {
    const long long u = t - D - (long long)m * D;
    if (u < 0) continue;                            // acc is still 0: nothing older exists
    acc = __fadd_rn(in[u], __fmul_rn(acc, 0.5f));
  }
  out[t] = __fsub_rn(acc, in[t]);
}

// wet/dry mix + JUCE's float -> 16-bit WAV conversion (float -> int32 full scale, round half even, keep the high 16 bits)
__global__ void fx_finish_kernel(const float* __restrict__ rev, const float* __restrict__ x, int16_t* __restrict__ out16,
                                 float* __restrict__ outf, long long n, float wet1, float dry) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float v = __fadd_rn(__fmul_rn(rev[t], wet1), __fmul_rn(x[t], dry));
  if (outf) outf[t] = v;
  const double d = (double)v;
  const int i32 = (d <= -1.0) ? INT_MIN : (d >= 1.0) ? INT_MAX : __double2int_rn(__dmul_rn(2147483647.0, d));
  out16[t] = (int16_t)(i32 >> 16);
}

// planar float stems -> interleaved 16-bit PCM as soundfile writes them (mdx.py:273,280): python-soundfile turns libsndfile's
// clipping on, which selects f2s_clip_array: scaled = x * 2^31 (float); >= 2^31 -> 0x7FFF; <= -2^31 -> -0x8000; else
// lrintf(scaled) >> 16
__global__ void pcm16_from_planar_kernel(const float* __restrict__ x, long long n, int channels, int16_t* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * channels) return;
  const long long j = e / channels;
  const int c = (int)(e % channels);
  const float s = __fmul_rn(x[(long long)c * n + j], 2147483648.0f);
  int q;
  if (s >= 2147483648.0f) q = 32767;
  else if (s <= -2147483648.0f) q = -32768;
  else if (s != s) q = 0;
  else q = (int)(__float2ll_rn(s) >> 16);
  out[e] = (int16_t)q;
}

// ---- pydub mix ----
struct MixSrc {
  const int16_t* x;
  long long n;        // frames in the file
  long long used;     // frames of the rate-converted segment that take part
  int ch, inr, outr;  // channels; rates divided by their gcd (inr == outr: no conversion)
  double g1, g2;      // the two apply_gain factors
};
struct MixParams {
  MixSrc s[3];
};

__device__ __forceinline__ int mul16(int v, double f) {       // audioop.mul: floor(clamp(v * f))
  double r = __dmul_rn((double)v, f);
  if (r > 32767.0) r = 32767.0;
  else if (r < -32767.0) r = -32768.0;
  return (int)floor(r);
}

__device__ __forceinline__ long long gained(const MixSrc& s, long long i, int c) {
  const int v = s.x[i * s.ch + ((c < s.ch) ? c : s.ch - 1)];   // audioop.tostereo(data, 2, 1, 1) copies the mono sample
  return mul16(mul16(v, s.g1), s.g2);
}

// audioop.ratecv(state=None, weightA=1, weightB=0): output j is emitted after input k = ceil(j*inr/outr) was read, with phase
// d = k*outr - j*inr: ((prev<<16)*d + (cur<<16)*(outr-d)) / outr, truncated toward zero, then >> 16
__device__ __forceinline__ int src_sample(const MixSrc& s, long long j, int c) {
  if (s.inr == s.outr) return (int)gained(s, j, c);
  const long long k = (j * s.inr + s.outr - 1) / s.outr;
  const long long d = k * s.outr - j * s.inr;
  const long long cur = gained(s, k, c), prev = (k > 0) ? gained(s, k - 1, c) : 0;
  const long long q = (65536LL * (prev * d + cur * (s.outr - d))) / s.outr;
  return (int)(q >> 16);
}

__device__ __forceinline__ int clip16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }

__global__ void pydub_mix_kernel(MixParams p, int16_t* __restrict__ out, long long n_out, int channels) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_out * channels) return;
  const long long j = e / channels;
  const int c = (int)(e % channels);
  int v = (j < p.s[0].used) ? src_sample(p.s[0], j, c) : 0;
  if (j < p.s[1].used) v = clip16(v + src_sample(p.s[1], j, c));      // audioop.add clips
  if (j < p.s[2].used) v = clip16(v + src_sample(p.s[2], j, c));
  out[e] = (int16_t)v;
}

}  // namespace
}  // namespace b200vc

using namespace b200vc;

extern "C" {

int b200vc_fx_hpf_comp(const int16_t* x, float* y, float* env_scratch, int64_t n, int chunk, int warm, float b0, float b1, float a1,
                       float cte_at, float cte_rl, float thr, float thr_inv, float expo, void* stream) {
  B200VC_RECORD(b200vc_fx_hpf_comp(x, y, env_scratch, n, chunk, warm, b0, b1, a1, cte_at, cte_rl, thr, thr_inv, expo, stream));
  B200VC_REQUIRE(x && y && env_scratch && n > 0 && chunk > 0 && warm >= 0, "fx_hpf_comp: bad args");
  B200VC_REQUIRE(chunk % 8 == 0 && warm % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(env_scratch) & 15) == 0,
                 "fx_hpf_comp: chunk / warm must be multiples of 8 and x / y 16-byte aligned");
  const long long chunks = (n + chunk - 1) / chunk;
  HpfEnv f{b0, b1, a1, cte_at, cte_rl, 0.f, 0.f};
  fx_hpf_env_kernel<<<(unsigned)((chunks + 31) / 32), 32, 0, (cudaStream_t)stream>>>(x, y, env_scratch, n, chunk, warm, f);
  fx_gain_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(y, env_scratch, n, thr, thr_inv, expo);
  count_launch(2);
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_fx_reverb_combs(const float* x, float* Y, float* comb_sum, int64_t n, const int* delays8, float gain, float damp,
                           float feedback, int terms, void* stream) {
  B200VC_REQUIRE(x && Y && comb_sum && delays8 && n > 0 && terms > 0, "fx_reverb_combs: bad args");
  CombDelays dl;
  for (int j = 0; j < 8; ++j) {
    B200VC_REQUIRE(delays8[j] > 0, "fx_reverb_combs: comb delay %d is not positive", j);
    dl.d[j] = delays8[j];
  }
  int dmax = 0;
  for (int j = 0; j < 8; ++j) dmax = dl.d[j] > dmax ? dl.d[j] : dmax;
  B200VC_REQUIRE(dmax <= 256 * kCombRun, "fx_reverb_combs: comb delay %d > 2048 samples (sample rate above 55 kHz)", dmax);
  B200VC_REQUIRE(terms <= 256, "fx_reverb_combs: %d low-pass terms (damping too close to 1)", terms);
  int ring = 1024;
  while (ring < 2 * dmax + terms) ring *= 2;
  B200VC_REQUIRE(ring * sizeof(float) <= 200 * 1024, "fx_reverb_combs: delay lines of %d samples do not fit shared memory", dmax);
  const size_t smem = (ring + ring / 32) * sizeof(float);
  if (smem > 48 * 1024)
    B200VC_CHECK_CUDA(cudaFuncSetAttribute(fx_comb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  auto launch = [=](cudaStream_t st) {
    fx_comb_kernel<<<8, 256, smem, st>>>(x, Y, n, dl, ring, gain, damp, 1.0f - damp, feedback, terms);
    fx_comb_sum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Y, n, dl, comb_sum);
    count_launch(2);
  };
  if (plan_recording()) {
    plan_push([=](void* s) -> int {
      launch((cudaStream_t)s);
      return cudaGetLastError() == cudaSuccess ? kOk : kErrCuda;
    });
    return kOk;
  }
  launch((cudaStream_t)stream);
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_fx_allpass(const float* in, float* out, int64_t n, int delay, int terms, void* stream) {
  B200VC_RECORD(b200vc_fx_allpass(in, out, n, delay, terms, stream));
  B200VC_REQUIRE(in && out && in != out && n > 0 && delay > 0 && terms > 0, "fx_allpass: bad args");
  fx_allpass_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, out, n, delay, terms);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_fx_finish(const float* reverb, const float* x, int16_t* out16, float* out_f, int64_t n, float wet1, float dry,
                     void* stream) {
  B200VC_RECORD(b200vc_fx_finish(reverb, x, out16, out_f, n, wet1, dry, stream));
  B200VC_REQUIRE(reverb && x && out16 && n > 0, "fx_finish: bad args");
  fx_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reverb, x, out16, out_f, n, wet1, dry);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_pcm16_from_planar(const float* x, int64_t n, int channels, int16_t* out, void* stream) {
  B200VC_RECORD(b200vc_pcm16_from_planar(x, n, channels, out, stream));
  B200VC_REQUIRE(x && out && n > 0 && channels > 0, "pcm16_from_planar: bad args");
  pcm16_from_planar_kernel<<<(unsigned)((n * channels + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, channels, out);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

int b200vc_pydub_mix(const b200vc_mix_source* src3, int16_t* out, int64_t n_out, int channels, void* stream) {
  B200VC_REQUIRE(src3 && out && n_out > 0 && (channels == 1 || channels == 2), "pydub_mix: bad args");
  MixParams p;
  for (int i = 0; i < 3; ++i) {
    const b200vc_mix_source& s = src3[i];
    B200VC_REQUIRE(s.x && s.n > 0 && s.channels >= 1 && s.channels <= channels && s.in_rate > 0 && s.out_rate > 0 &&
                       s.used >= 0,
                   "pydub_mix: bad source %d", i);
    long long a = s.in_rate, b = s.out_rate;
    while (b) {
      const long long t = a % b;
      a = b;
      b = t;
    }
    const long long avail = (s.in_rate == s.out_rate) ? s.n : ((s.n - 1) * (s.out_rate / a)) / (s.in_rate / a) + 1;
    B200VC_REQUIRE(s.used <= avail, "pydub_mix: source %d uses %lld frames, %lld exist after rate conversion", i,
                   (long long)s.used, avail);
    p.s[i] = MixSrc{s.x, s.n, s.used, s.channels, (int)(s.in_rate / a), (int)(s.out_rate / a), s.gain1, s.gain2};
  }
  const long long total = n_out * channels;
  if (plan_recording()) {
    plan_push([=](void* s) -> int {
      pydub_mix_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)s>>>(p, out, n_out, channels);
      count_launch();
      return cudaGetLastError() == cudaSuccess ? kOk : kErrCuda;
    });
    return kOk;
  }
  pydub_mix_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, out, n_out, channels);
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

}  // extern "C"
