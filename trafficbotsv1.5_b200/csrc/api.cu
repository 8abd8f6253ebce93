// C-ABI glue: error strings, version, precision dispatch of tb_linear.
#include "common.cuh"

int tb_linear_f32(const float* X, int ldx, const float* W, const float* bias, int bias_group, float* Y, int ldy, int M,
                  int N, int K,
                  int relu, const uint8_t* mask_pre, const float* res, int ldr, const uint8_t* mask_post,
                  cudaStream_t st);
int tb_linear_tc(const void* X, int ldx, const void* W, int in_f16, const float* bias, int bias_group, float* Y,
                 int ldy, int M, int N, int K,
                 int relu, const uint8_t* mask_pre, const float* res, int ldr, const uint8_t* mask_post,
                 void* Yh, int ldyh, int colh, cudaStream_t st, const float* ln_g = nullptr,
                 const float* ln_b = nullptr, void* ln_out = nullptr, int ld_ln = 0);

extern "C" const char* tb_strerror(int code) {
  switch (code) {
    case TB_OK: return "ok";
    case TB_ERR_BAD_SHAPE: return "bad shape";
    case TB_ERR_KNN_RANGE: return "need 0 < K < T";
    case TB_ERR_UNSUPPORTED: return "unsupported size";
    case TB_ERR_MISALIGNED: return "misaligned pointer or leading dimension";
    case TB_ERR_NULL: return "null pointer";
    case TB_ERR_CUDA: return "CUDA launch failed";
    default: return "unknown error";
  }
}

extern "C" int tb_version(void) { return 200; }

unsigned int* tb_fp16_flag_ptr = nullptr;

extern "C" int tb_set_fp16_flag(unsigned int* d_flag) {
  tb_fp16_flag_ptr = d_flag;  // NULL switches the guard's flag write off (conversions still saturate)
  return TB_OK;
}

extern "C" int tb_linear(const void* X, int ldx, const void* W, const float* bias, int bias_group, float* Y, int ldy,
                         int M, int N, int K, int relu, const uint8_t* mask_pre, const float* res, int ldr, const uint8_t* mask_post,
                         int precision, void* Yh, int ldyh, int col_h, void* stream) {
  if (!X || !W) return TB_ERR_NULL;
  const int n32 = Yh ? col_h : N;  // columns written to the fp32 output
  if (M <= 0 || N <= 0 || K <= 0 || ldx < K || (res && ldr < N) || bias_group < 0) return TB_ERR_BAD_SHAPE;
  if (Yh && (col_h < 0 || col_h >= N || (col_h & 31) || ldyh < N - col_h)) return TB_ERR_BAD_SHAPE;
  if (n32 > 0 && (!Y || ldy < n32)) return Y ? TB_ERR_BAD_SHAPE : TB_ERR_NULL;
  if (Yh && ((reinterpret_cast<uintptr_t>(Yh) & 1) != 0)) return TB_ERR_MISALIGNED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == 0) {
    if (Yh) return TB_ERR_UNSUPPORTED;  // fp16 tables belong to the tensor-core mode
    return tb_linear_f32(static_cast<const float*>(X), ldx, static_cast<const float*>(W), bias, bias_group, Y, ldy, M, N,
                         K, relu, mask_pre, res, ldr, mask_post, st);
  }
  if (precision == 1 || precision == 2)
    return tb_linear_tc(X, ldx, W, precision == 2, bias, bias_group, Y, ldy, M, N, K, relu, mask_pre, res, ldr,
                        mask_post, Yh, ldyh, col_h, st);
  return TB_ERR_UNSUPPORTED;
}

// Projection + residual + LayerNorm of the result in one launch (tensor-core mode): Y = X W^T + bias (+ masks / residual
// as tb_linear) and ln_out = fp16(LayerNorm(Y) * gamma + beta), the operand rows of the next kind::f16 projection.
extern "C" int tb_linear_ln(const void* X, int ldx, const void* W, const float* bias, float* Y, int ldy, int M, int N,
                            int K, const uint8_t* mask_pre, const float* res, int ldr, const uint8_t* mask_post,
                            int precision, const float* ln_gamma, const float* ln_beta, void* ln_out, int ld_ln,
                            void* stream) {
  if (!X || !W || !Y || !ln_gamma || !ln_beta || !ln_out) return TB_ERR_NULL;
  if (M <= 0 || N <= 0 || K <= 0 || ldx < K || ldy < N || ld_ln < N || (res && ldr < N)) return TB_ERR_BAD_SHAPE;
  if (precision != 1 && precision != 2) return TB_ERR_UNSUPPORTED;
  return tb_linear_tc(X, ldx, W, precision == 2, bias, 0, Y, ldy, M, N, K, 0, mask_pre, res, ldr, mask_post, nullptr, 0, 0,
                      static_cast<cudaStream_t>(stream), ln_gamma, ln_beta, ln_out, ld_ln);
}
