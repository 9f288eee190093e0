// tcgen05 tap-GEMM v2: PERSISTENT, double-buffered TMEM accumulators, coalesced epilogue.
//
//   grid  = min(#tiles, #SMs) CTAs, each loops over output tiles (n-tile fastest so neighbouring CTAs share the A box in L2)
//   warp 0 : TMA producer — the shared-memory ring keeps filling across tile boundaries
//   warp 1 : TMEM allocator (2 x BN columns) + single-thread tcgen05.mma issuer; tcgen05.commit -> smem-empty / tmem-full
//   warps 2-17 : epilogue — tcgen05.ld of accumulator stage `it & 1` while the MMA warp already fills the other stage;
//               warp = (TMEM lane quarter, 32-column group mod 4), four warps per SM sub-partition, each thread
//               finishing its row's columns in registers (tg_store16)
//
// Same descriptor, same epilogue semantics as tapgemm.cuh (scalar form tg_epi1).
#include "tapgemm.cuh"
#include <cuda.h>
#include <cstdlib>

namespace b200vc {

int encode_map_f32(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims,
                   const cuuint64_t* strides_bytes, const cuuint32_t* box);   // tapgemm_tc.cu
int encode_map_f16(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims,
                   const cuuint64_t* strides_bytes, const cuuint32_t* box);   // tapgemm_tc.cu

namespace {

constexpr int KCHUNK = 32;
constexpr int EPI_WARPS = 16;                     // (TMEM lane quarter) x (32-column group mod 4): 4 warps per SM sub-partition
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// fp16 operands (kind::f16): same shared-memory geometry (128-byte rows, 32-byte K steps), twice the K per step
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
template <bool F16>
__device__ __forceinline__ void umma_any(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if constexpr (F16) umma_f16(tmem_d, adesc, bdesc, idesc, acc);
  else umma_tf32(tmem_d, adesc, bdesc, idesc, acc);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// MH = 128-row halves per tile.  MH == 2: one B tile feeds two MMAs (rows 0-127 and 128-255 of a 256-row A box), which
// cuts the L2->SM operand bytes per FLOP by 1.3-1.5x — the persistent kernel runs at the L2 throughput cap
// (~43 B/clk/SM, profiles/r01_ncu_kernels.md), so that ratio is its speed.
template <int BN, int STAGES, int MH, bool F16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tapgemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                   const __grid_constant__ TgParams p, int ntiles_n, int total_tiles, int tbw, int tbh, int bias_floats) {
  constexpr int KE = F16 ? 2 * KCHUNK : KCHUNK;                    // K elements per 128-byte chunk row
  constexpr int ES = F16 ? 2 : 4;                                  // operand element size
  constexpr int A_STAGE_BYTES = MH * TG_TILE_M * 128;
  constexpr int B_STAGE_BYTES = BN * 128;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr int ACC_COLS = MH * BN;                                // one accumulator stage: MH blocks of BN columns
  constexpr int NACC = (2 * ACC_COLS <= 512) ? 2 : 1;              // double-buffered when TMEM allows
  constexpr uint32_t TMEM_COLS = (NACC * ACC_COLS < 32) ? 32 : NACC * ACC_COLS;   // power of two for BN in {32..256}
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto a_stage = [&](int s) { return smem_base + s * STAGE_BYTES; };
  auto b_stage = [&](int s) { return smem_base + s * STAGE_BYTES + A_STAGE_BYTES; };
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  // per-column bias staged once per CTA (ncu r01e: the broadcast global bias loads in the epilogue cost 33 % of an
  // HBM-bound layer: four dependent L1 round trips per 16 columns on warps that have nothing else to overlap them with)
  float* bias_s = reinterpret_cast<float*>(smem_raw + (bar_base - smem_u32(smem_raw)) + 8 * (2 * STAGES + 6));
  for (int i = threadIdx.x; i < bias_floats; i += NUM_THREADS) bias_s[i] = i < p.N ? __ldg(p.bias + i) : 0.f;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform role index
  const int ntw = (p.OW + tbw - 1) / tbw;
  const int nth = (p.OH + tbh - 1) / tbh;
  const int kchunks = (p.Kc + KE - 1) / KE;
  const int nchunks = p.ntaps * kchunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), BN >= 128 ? EPI_WARPS : BN / 8);   // 4 warps per live 32-column group
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  auto tile_coords = [&](int tile, int& w0, int& h0, int& tb, int& n0) {
    const int nt = tile % ntiles_n;
    int mt = tile / ntiles_n;
    const int tw = mt % ntw; mt /= ntw;
    const int th = mt % nth; mt /= nth;
    tb = mt;
    w0 = tw * tbw;
    h0 = th * tbh;
    n0 = nt * BN;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    // This loop paces the whole kernel when it is long (ncu r01d: ~75 dependent uniform-datapath instructions per
    // k-chunk = ~600 cycles, more than the MMAs of a 128-wide chunk): no divisions, tap record loaded once per tap,
    // stage index / phase carried incrementally.
    {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int w0, h0, tb, n0;
        tile_coords(tile, w0, h0, tb, n0);
        const int wsel = tb * p.w_batch_step;
        for (int tap_i = 0; tap_i < p.ntaps; ++tap_i) {
          const TgTap tap = p.taps[tap_i];
          const int cw = w0 + tap.dw, ch = h0 + tap.dh, cp = tap.dp, widx = tap.widx + wsel;
          int ca = tap.c_off;
          for (int kc0 = 0; kc0 < p.Kc; kc0 += KE, ca += KE) {
            mbar_wait(empty_bar(s), ph ^ 1u);
            if (elect_one()) {
              mbar_expect_tx(full_bar(s), STAGE_BYTES);
              tma_load_5d(a_stage(s), &tmA, full_bar(s), ca, cw, ch, tb, cp);
              tma_load_3d(b_stage(s), &tmW, full_bar(s), kc0, n0, widx);
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // whole warp walks the loops (uniform control flow -> descriptors stay in uniform registers); one elected lane issues
    {
      int it = 0, s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it % NACC;
        const uint32_t use = (uint32_t)(it / NACC);
        mbar_wait(tempty_bar(acc), (use & 1u) ^ 1u);      // epilogue has drained this accumulator stage
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
        // instruction N = the live part of this n-tile rounded up to 16 (e.g. 48 for the MDX c=48 layers)
        const int n0_ = (tile % ntiles_n) * BN;
        const int nrem = p.N - n0_;
        const int n_ins = nrem >= BN ? BN : ((nrem + 15) & ~15);
        const uint32_t IDESC = F16 ? make_idesc_f16(128, n_ins) : make_idesc_tf32(128, n_ins);
        // the last k-chunk of a tap may be short (Kc = 48 -> 32 + 16): issue only the K=8 steps that hold data
        const int klast = ((p.Kc - (kchunks - 1) * KE) * ES + 31) >> 5;      // 32-byte K steps that hold data
        uint32_t accum = 0u;
        for (int chunk = 0, kc = 0; chunk < nchunks; ++chunk) {
          mbar_wait(full_bar(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int nk = (kc == kchunks - 1) ? klast : KCHUNK / 8;
          if (elect_one()) {
            const uint64_t adesc = make_smem_desc(a_stage(s));
            const uint64_t bdesc = make_smem_desc(b_stage(s));
            if (nk == KCHUNK / 8) {               // full chunk: straight-line issue
#pragma unroll
              for (int k = 0; k < KCHUNK / 8; ++k) {
#pragma unroll
                for (int mh = 0; mh < MH; ++mh)     // 128 rows x 128 B = 16 KB further into the A box: +1024 in the (addr >> 4) field
                  umma_any<F16>(d_tmem + (uint32_t)(mh * BN), adesc + (uint64_t)(mh * 1024 + 2 * k), bdesc + (uint64_t)(2 * k), IDESC, accum);
                accum = 1u;
              }
            } else {
              for (int k = 0; k < nk; ++k) {
#pragma unroll
                for (int mh = 0; mh < MH; ++mh)
                  umma_any<F16>(d_tmem + (uint32_t)(mh * BN), adesc + (uint64_t)(mh * 1024 + 2 * k), bdesc + (uint64_t)(2 * k), IDESC, accum);
                accum = 1u;
              }
            }
            umma_commit(empty_bar(s));
            if (chunk == nchunks - 1) umma_commit(tfull_bar(acc));
          }
          accum = 1u;
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
          if (++kc == kchunks) kc = 0;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    const int q = warp & 3;
    const int cpar = (warp - 2) >> 2;              // this warp takes the 32-column groups with index % 4 == cpar
    int it = 0;
    if (cpar * 32 < BN)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int w0, h0, tb, n0;
      tile_coords(tile, w0, h0, tb, n0);
      const int acc = it % NACC;
      const uint32_t use = (uint32_t)(it / NACC);
      mbar_wait(tfull_bar(acc), use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int mh = 0; mh < MH; ++mh) {
        const int m = mh * 128 + q * 32 + lane;      // row inside the (tbw x tbh) box, w fastest
        const TgRow r = tg_row(p, tb, h0 + m / tbw, w0 + m % tbw);
#pragma unroll 1
        for (int c0 = cpar * 32; c0 < BN; c0 += 128) {
          if (n0 + c0 >= p.N) break;                 // warp-uniform
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + mh * BN + c0), v);
          tg_store16(p, r, n0 + c0, v, bias_floats ? bias_s + n0 + c0 : nullptr, true);
          tg_store16(p, r, n0 + c0 + 16, v + 16, bias_floats ? bias_s + n0 + c0 + 16 : nullptr, true);
        }
      }
      // all TMEM reads of this accumulator stage are complete (tcgen05.wait::ld inside tmem_ld32)
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int BN, int STAGES, int MH, bool F16>
int launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmW, const TgParams& p, int ntiles_n, int total_tiles,
               int grid, int tbw, int tbh, cudaStream_t stream) {
  constexpr int smem_fixed = STAGES * (MH * TG_TILE_M * 128 + BN * 128) + 8 * (2 * STAGES + 6) + 1024;
  static_assert(smem_fixed <= 227 * 1024, "shared memory budget");
  // per-column bias in shared memory when it fits beside the operand ring (padded to whole n-tiles so every 16-column
  // group the epilogue touches is readable)
  const int want = (p.bias && !p.bias_per_row) ? ntiles_n * BN : 0;
  const int bias_floats = (want > 0 && smem_fixed + want * 4 + 16 <= 227 * 1024) ? want : 0;
  const int smem = smem_fixed + (bias_floats ? bias_floats * 4 + 16 : 0);
  static int configured = 0;
  if (configured < smem) {
    B200VC_CHECK_CUDA(cudaFuncSetAttribute(tapgemm_tc2_kernel<BN, STAGES, MH, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 227 * 1024;
  }
  tapgemm_tc2_kernel<BN, STAGES, MH, F16><<<grid, NUM_THREADS, smem, stream>>>(tmA, tmW, p, ntiles_n, total_tiles, tbw, tbh, bias_floats);
  return kOk;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace

static int g_rows256 = [] { const char* e = getenv("B200VC_TC2_ROWS256"); return (e && e[0] == '1') ? 1 : 0; }();
void tapgemm_tc2_set_rows256(int on) { g_rows256 = on ? 1 : 0; }

int tapgemm_tc2_launch(const TgParams& p, cudaStream_t stream) {
  B200VC_REQUIRE(tapgemm_tc_supported(p), "tapgemm_tc: operand alignment not TMA-compatible");
  const int BN = p.N > 128 ? 256 : (p.N > 64 ? 128 : (p.N > 32 ? 64 : 32));
  const int ntiles_n = ceil_div(p.N, BN);
  // 256-row tiles (two MMAs per B tile) halve the weight-tile L2->SM traffic.  Measured on B200 they do NOT pay: the
  // kernel is bound by shared-memory bandwidth (TMA writes + UMMA operand reads share 128 B/clk/SM, and every MMA
  // re-reads its B slice), not by L2 — C256 k11 662 -> 608 TFLOP/s, 768->3072 linear 306 -> 221, C128 k7 486 -> 509.
  // Kept behind b200vc_tapgemm_set_rows256(1) (initial value: env B200VC_TC2_ROWS256) for experiments and tests.
  const bool f16 = (p.dtype & TG_DT_AB) != 0;
  const unsigned long long es = f16 ? 2ull : 4ull;                 // operand element size
  const int epb = f16 ? 8 : 4;                                       // elements per 16 bytes (TMA stride granule)
  const bool rows256 = g_rows256 != 0 && !f16;
  int tbw = p.BW, tbh = p.BH, MH = 1;
  if (rows256) {
    const int bw2 = p.BH == 1 ? 2 * p.BW : p.BW, bh2 = p.BH == 1 ? 1 : 2 * p.BH;
    if (bw2 <= 256 && bh2 <= 256) {
      const long long t1 = (long long)ceil_div(p.OW, p.BW) * ceil_div(p.OH, p.BH) * p.OB;
      const long long t2 = (long long)ceil_div(p.OW, bw2) * ceil_div(p.OH, bh2) * p.OB;
      if (t2 * ntiles_n >= num_sms() && 2 * t2 * 100 <= t1 * 105) { tbw = bw2; tbh = bh2; MH = 2; }
    }
  }
  const int ntw = ceil_div(p.OW, tbw), nth = ceil_div(p.OH, tbh);
  const long long mtiles = (long long)ntw * nth * p.OB;
  const long long total = mtiles * ntiles_n;
  B200VC_REQUIRE(total > 0 && total < (1ll << 31), "tapgemm_tc: bad tile count %lld", total);

  CUtensorMap tmA, tmW;
  {
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5] = {(cuuint32_t)(f16 ? 2 * KCHUNK : KCHUNK), (cuuint32_t)tbw, (cuuint32_t)tbh, 1, 1};
    long long span = 1;
    for (int i = 0; i < 5; ++i) {
      dims[i] = (cuuint64_t)(p.a_dim[i] > 0 ? p.a_dim[i] : 1);
      if (i > 0) {
        long long st = p.a_stride[i];
        if (p.a_dim[i] <= 1) st = ((span + epb - 1) / epb) * epb;
        strides[i - 1] = (cuuint64_t)st * es;
        span = st * (long long)dims[i];
      } else {
        span = (long long)dims[0];
      }
    }
    int rc = f16 ? encode_map_f16(&tmA, p.A, 5, dims, strides, box) : encode_map_f32(&tmA, p.A, 5, dims, strides, box);
    if (rc) return rc;
  }
  {
    int max_widx = 0;
    for (int i = 0; i < p.ntaps; ++i) max_widx = p.taps[i].widx > max_widx ? p.taps[i].widx : max_widx;
    const int nw = max_widx + 1 + (p.OB - 1) * p.w_batch_step;
    cuuint64_t dims[3] = {(cuuint64_t)p.Kc, (cuuint64_t)p.N, (cuuint64_t)nw};
    long long wst = p.wstride;
    if (nw <= 1) wst = ((p.ldw * p.N + epb - 1) / epb) * epb;
    cuuint64_t strides[2] = {(cuuint64_t)p.ldw * es, (cuuint64_t)wst * es};
    cuuint32_t box[3] = {(cuuint32_t)(f16 ? 2 * KCHUNK : KCHUNK), (cuuint32_t)BN, 1};
    int rc = f16 ? encode_map_f16(&tmW, p.Wt, 3, dims, strides, box) : encode_map_f32(&tmW, p.Wt, 3, dims, strides, box);
    if (rc) return rc;
  }
  const int grid = (int)(total < num_sms() ? total : num_sms());
  int rc;
  if (f16) {
    switch (BN) {
      case 256: rc = launch_cfg<256, 4, 1, true>(tmA, tmW, p, ntiles_n, (int)total, grid, tbw, tbh, stream); break;
      case 128: rc = launch_cfg<128, 6, 1, true>(tmA, tmW, p, ntiles_n, (int)total, grid, tbw, tbh, stream); break;
      case 64:  rc = launch_cfg<64, 8, 1, true>(tmA, tmW, p, ntiles_n, (int)total, grid, tbw, tbh, stream); break;
      default:  rc = launch_cfg<32, 8, 1, true>(tmA, tmW, p, ntiles_n, (int)total, grid, tbw, tbh, stream); break;
    }
  } else if (MH == 2) {
    switch (BN) {
      case 256: rc = launch_cfg<256, 3, 2, false>(tmA, tmW, p, ntiles_n, (int)total, grid, tbw, tbh, stream); break;
      case 128: rc = launch_cfg<128, 4, 2, false>(tmA, tmW, p, ntiles_n, (int)total, grid, tbw, tbh, stream); break;
      case 64:  rc = launch_cfg<64, 5, 2, false>(tmA, tmW, p, ntiles_n, (int)total, grid, tbw, tbh, stream); break;
      default:  rc = launch_cfg<32, 6, 2, false>(tmA, tmW, p, ntiles_n, (int)total, grid, tbw, tbh, stream); break;
    }
  } else {
    switch (BN) {
      case 256: rc = launch_cfg<256, 4, 1, false>(tmA, tmW, p, ntiles_n, (int)total, grid, tbw, tbh, stream); break;
      case 128: rc = launch_cfg<128, 6, 1, false>(tmA, tmW, p, ntiles_n, (int)total, grid, tbw, tbh, stream); break;
      case 64:  rc = launch_cfg<64, 8, 1, false>(tmA, tmW, p, ntiles_n, (int)total, grid, tbw, tbh, stream); break;
      default:  rc = launch_cfg<32, 8, 1, false>(tmA, tmW, p, ntiles_n, (int)total, grid, tbw, tbh, stream); break;
    }
  }
  if (rc) return rc;
  count_launch();
  B200VC_LAUNCH_CHECK();
  return kOk;
}

}  // namespace b200vc
