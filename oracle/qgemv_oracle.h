/* TEST INFRASTRUCTURE — CPU restatement ("port" oracle) of the reference's
 * quantized mat-vec hot path.  See qgemv_oracle.c for the citations.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
 * this; the product path (llm_inference_b200/) never does. */
#ifndef QGEMV_ORACLE_H
#define QGEMV_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ggml tensor type ids (gguf.h:30-46) */
enum {
  ORC_F32 = 0, ORC_F16 = 1, ORC_Q4_0 = 2, ORC_Q5_0 = 6, ORC_Q8_0 = 8,
  ORC_Q4_K = 12, ORC_Q6_K = 14, ORC_BF16 = 30
};

float orc_f16_to_f32(uint16_t h);
uint16_t orc_f32_to_f16(float f);
float orc_bf16_to_f32(uint16_t h);
int orc_nearest_int(float v);

/* bytes per weight row for a format, or 0 if unsupported / K not divisible */
size_t orc_row_bytes(uint32_t ggml_type, size_t n_cols);

/* y: n/32 records of 34 bytes {u16 d; i8 qs[32]} */
void orc_quantize_row_q8_0(const float* x, uint8_t* y, size_t n);
/* y: n/256 records of 292 bytes {f32 d; i8 qs[256]; i16 bsums[16]} */
void orc_quantize_row_q8_k(const float* x, uint8_t* y, size_t n);

/* GEMVs.  `dots` (nullable) receives the per-block integer dot products that
 * must be bit-exact on the GPU:
 *   q4_0/q8_0: n_rows * (K/32) int32 (full 32-element block dot)
 *   q4_k     : n_rows * (K/32) int32 (the 32-element nibble·q8 sums)
 *   q6_k     : n_rows * (K/128) int32 (scale-weighted 128-element sums) */
void orc_gemv_q4_0(float* o, const uint8_t* w, const float* x, size_t n_rows,
                   size_t n_cols, int32_t* dots);
void orc_gemv_q8_0(float* o, const uint8_t* w, const float* x, size_t n_rows,
                   size_t n_cols, int32_t* dots);
void orc_gemv_q4_k(float* o, const uint8_t* w, const float* x, size_t n_rows,
                   size_t n_cols, int32_t* dots);
void orc_gemv_q6_k(float* o, const uint8_t* w, const float* x, size_t n_rows,
                   size_t n_cols, int32_t* dots);
void orc_gemv_q5_0(float* o, const uint8_t* w, const float* x, size_t n_rows,
                   size_t n_cols);
void orc_gemv_bf16(float* o, const uint16_t* w, const float* x, size_t n_rows,
                   size_t n_cols);
void orc_gemv_f16(float* o, const uint16_t* w, const float* x, size_t n_rows,
                  size_t n_cols);

/* The same per-block terms summed in the DEVICE's canonical order (chunks of 16 blocks, four chains per chunk,
 * (s0+s1)+(s2+s3), chunks left to right): GPU outputs must equal these bit for bit. */
void orc_gemv_q4_0_canonical(float* o, const uint8_t* w, const float* x, size_t n_rows, size_t n_cols);
void orc_gemv_q8_0_canonical(float* o, const uint8_t* w, const float* x, size_t n_rows, size_t n_cols);
void orc_gemv_q4_k_canonical(float* o, const uint8_t* w, const float* x, size_t n_rows, size_t n_cols);
void orc_gemv_q6_k_canonical(float* o, const uint8_t* w, const float* x, size_t n_rows, size_t n_cols);
void orc_gemv_q5_0_canonical(float* o, const uint8_t* w, const float* x, size_t n_rows, size_t n_cols);
void orc_gemv_f16_canonical(float* o, const uint16_t* w, const float* x, size_t n_rows, size_t n_cols);
void orc_gemv_bf16_canonical(float* o, const uint16_t* w, const float* x, size_t n_rows, size_t n_cols);

/* dispatcher: 0 ok, 1 unsupported type (the reference throws, ops.cpp:952-955) */
int orc_mat_vec_mul(uint32_t ggml_type, float* o, const uint8_t* w,
                    const float* x, size_t n_rows, size_t n_cols);

/* row dequantizers used by the embedding lookup (0 ok, 1 unsupported) */
int orc_dequantize_row(uint32_t ggml_type, const uint8_t* row, size_t n_cols,
                       float* out);

/* double-precision value of a row's exact mathematical result, for error
 * budgeting:  sum_abs receives sum |terms| (may be NULL).  Quantized formats
 * use the same quantized activations as the float path. */
double orc_row_exact(uint32_t ggml_type, const uint8_t* w_row, const float* x,
                     size_t n_cols, double* sum_abs);

#ifdef __cplusplus
}
#endif
#endif
