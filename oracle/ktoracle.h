/* TEST INFRASTRUCTURE ONLY — CPU oracle for the quantized-MoE decode hot path.
 *
 * Plain-C restatement of the reference's CPU arithmetic (llamafile/ggml path of
 * kt-kernel / archive cpuinfer_ext).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product (libktb200.so and the
 * ktransformers_b200 package) never does.
 *
 * Parity status: PINNED — tests/test_oracle_pinned.py checks every function here against
 * (a) oracle/_ref (the unmodified reference sources compiled by oracle/Makefile) where that
 * library is present and (b) the committed fixtures under tests/golden/ that were generated
 * from oracle/_ref by tests/golden/make_golden.py.
 */
#ifndef KTORACLE_H
#define KTORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ggml type ids (third_party/llama.cpp/ggml.h:349-380) */
enum {
    KTO_F32 = 0, KTO_F16 = 1, KTO_Q8_0 = 8, KTO_Q2_K = 10, KTO_Q3_K = 11, KTO_Q4_K = 12,
    KTO_Q5_K = 13, KTO_Q6_K = 14, KTO_Q8_K = 15, KTO_IQ4_XS = 23, KTO_BF16 = 30
};

long kto_type_size(int type);   /* bytes per block; 0 if unsupported */
long kto_blck_size(int type);   /* elements per block */
int  kto_vec_dot_type(int type);

float    kto_fp16_to_fp32(uint16_t h);
uint16_t kto_fp32_to_fp16(float f);
float    kto_bf16_to_fp32(uint16_t h);
uint16_t kto_fp32_to_bf16(float f);

void kto_to_float(const void* in, float* out, long n, int type);      /* dequantize / widen */
void kto_from_float(const float* in, void* out, long n, int type);    /* F32,F16,BF16,Q8_0,Q8_K only */

float kto_vec_dot(int wtype, long n, const void* w, const void* act); /* act is in kto_vec_dot_type(wtype) */

/* MOE::forward semantics (forward_one per token). ids are int64 (kt-kernel) — the archive's uint64 is the
 * same bits for valid ids. Ids outside [0,E) are skipped (kt-kernel common.hpp:255-258). */
void kto_moe_forward(int expert_num, int hidden, int inter, int use_silu,
                     const void* gate, const void* up, const void* down,
                     int gate_type, int up_type, int down_type, int hidden_type,
                     int qlen, int k, const int64_t* expert_ids, const float* weights,
                     const void* input, void* output);

/* Linear::forward / MLP::forward semantics (operators/llamafile/linear.cpp, mlp.cpp). */
void kto_linear_forward(int in_size, int out_size, const void* proj, int proj_type, int hidden_type,
                        int qlen, const void* input, void* output);
void kto_mlp_forward(int hidden, int inter, const void* gate, const void* up, const void* down,
                     int gate_type, int up_type, int down_type, int hidden_type,
                     int qlen, const void* input, void* output);

#ifdef __cplusplus
}
#endif
#endif
