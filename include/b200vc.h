/* b200vc — C ABI of the B200-native RVC / MDX hot path.
 *
 * The reference (SociallyIneptWeeb/AICoverGen) is pure Python; its "operator
 * boundary" for this path is the set of library calls it makes into
 * torch/cuDNN/cuBLAS/cuFFT/onnxruntime/faiss.  Every entry point below names
 * the reference call site(s) it replaces (paths relative to the reference
 * repo).  A maintainer binds them with ctypes (see INTEGRATION.md and
 * aicovergen_b200/_ffi.py): plain pointers and sizes only, no torch types.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; the message is
 *     retrievable with b200vc_last_error() (thread local).
 *   - all device pointers are caller-owned CUDA allocations (torch tensors
 *     passed by data_ptr()); fp32 unless stated; the library never frees or
 *     keeps them beyond the stream work it enqueues.
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *     calls are asynchronous unless documented otherwise.
 *   - activations are channels-last: 1-D signals [T, C], images [H, W, C].
 */
#ifndef B200VC_H_
#define B200VC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VC_MAX_TAPS 128

/* activation codes */
enum {
  B200VC_ACT_NONE = 0,
  B200VC_ACT_RELU = 1,
  B200VC_ACT_LRELU = 2,
  B200VC_ACT_GELU = 3,
  B200VC_ACT_TANH = 4,
  B200VC_ACT_SIGMOID = 5,
  B200VC_ACT_EXP = 6
};

/* which kernel family executes a tap-GEMM */
enum {
  B200VC_BACKEND_SIMT_FP32 = 0, /* exact fp32 FMA kernel                      */
  B200VC_BACKEND_TC_TF32 = 1    /* tcgen05.mma kind::tf32, TMA-fed, TMEM accum */
};

typedef struct b200vc_tap {
  int32_t c_off; /* channel offset inside A dim 0                     */
  int16_t dw;    /* offset along A dim 1                              */
  int16_t dh;    /* offset along A dim 2                              */
  int16_t dp;    /* coordinate along A dim 4                          */
  int16_t widx;  /* weight slice multiplied with this tap             */
} b200vc_tap;

/* Tap-GEMM problem descriptor:
 *   D[pixel, n] = sum_tap sum_{k<Kc} A[pixel + tap, tap.c_off + k] * W[tap.widx + b*w_batch_step, n, k]
 * Replaces F.conv1d / F.conv2d / F.conv_transpose1d/2d / F.linear / torch.matmul
 * calls at: infer_pack/modules.py:196,206,304,308 ; infer_pack/models.py:97,105,497-514 ;
 * infer_pack/attentions.py:217-223,233,263,392-398 ; rmvpe.py:27-49,147-155,241,245 ;
 * fairseq HubertModel (vc_infer_pipeline.py:405) ; the ONNX session at mdx.py:77 ;
 * the STFT/iSTFT DFTs at mdx.py:39,53 and rmvpe.py:305. */
typedef struct b200vc_tapgemm_params {
  const float* A;
  int32_t a_dim[5];     /* extents (c, w, h, b, p)                            */
  int64_t a_stride[5];  /* element strides, a_stride[0] == 1                  */
  const float* Wt;
  int64_t ldw, wstride; /* weight strides: (k, n, widx) = (1, ldw, wstride)   */
  int32_t w_batch_step;
  int32_t Kc, N, ntaps;
  int32_t OW, OH, OB;   /* output pixel space                                 */
  int32_t BW, BH;       /* tile box, BW*BH == 128                             */
  int32_t osh, osw, ooh, oow; /* out pixel = (h*osh+ooh, w*osw+oow)           */
  int32_t o_fh, o_fw;   /* full output extents: mapped pixel must lie in [0,o_fh) x [0,o_fw) */
  int64_t o_sb, o_sh, o_sw;   /* output element strides                       */
  int64_t r_sb, r_sh, r_sw;   /* residual #1 element strides (pixel b,h,w)    */
  const float* bias;
  int32_t bias_per_row;
  int32_t act_pre;
  float act_pre_p;
  const float* res;
  float scale;
  const float* res2;
  int32_t act_post;
  float act_post_p;
  float* out;
  float* out2;
  int32_t act2;
  float act2_p;
  int32_t vec4;         /* bit0: k-vectorised loads ok, bit1: n-vectorised epilogue ok */
  b200vc_tap taps[B200VC_MAX_TAPS];
} b200vc_tapgemm_params;

/* ---- library ---- */
const char* b200vc_version(void);
const char* b200vc_last_error(void);
/* number of kernels this library has launched in this process (bench.py: gpu_launches) */
int64_t b200vc_launch_count(void);
/* sizeof(b200vc_tapgemm_params) as compiled: bindings verify their struct mirror against it */
int64_t b200vc_sizeof_tapgemm_params(void);

/* ---- tap-GEMM ---- */
int b200vc_tapgemm(const b200vc_tapgemm_params* p, int backend, void* stream);
/* 1 when the descriptor satisfies the TMA alignment rules of the tcgen05 path */
int b200vc_tapgemm_tc_supported(const b200vc_tapgemm_params* p);

#ifdef __cplusplus
}
#endif
#endif /* B200VC_H_ */
