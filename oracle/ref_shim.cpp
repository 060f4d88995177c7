// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// Thin extern "C" face over the UNMODIFIED reference CPU implementation of the
// quantized-MoE path, so that tests / bench.py's cpu_baseline leg can drive it with
// ctypes.  The reference sources are compiled from where they lie under
// /root/reference (see oracle/Makefile); nothing from the reference is copied here.
//
// What it exposes (reference entry point it forwards to):
//   ktref_moe_*     -> MOE::MOE / MOE::forward            archive/csrc/ktransformers_ext/operators/llamafile/moe.cpp:20,367
//   ktref_linear_*  -> Linear::Linear / Linear::forward   archive/csrc/ktransformers_ext/operators/llamafile/linear.cpp:12,65
//   ktref_mlp_*     -> MLP::MLP / MLP::forward            archive/csrc/ktransformers_ext/operators/llamafile/mlp.cpp
//   ktref_from_float / ktref_to_float -> conversion.h:18-36 (ggml type-traits from_float / to_float)
//   ktref_vec_dot   -> ggml type-traits vec_dot           third_party/llama.cpp/ggml.c:743-754
//   ktref_sgemm     -> llamafile_sgemm                    third_party/llamafile/sgemm.h:63
// The thread pool is the reference's own work-stealing Backend (cpu_backend/backend.cpp),
// created exactly like CPUInfer's constructor does (cpu_backend/cpuinfer.h:38-45), including
// the fp16 lookup-table initialisation the reference performs there.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "cpu_backend/backend.h"
#include "operators/llamafile/conversion.h"
#include "operators/llamafile/linear.h"
#include "operators/llamafile/mlp.h"
#include "operators/llamafile/moe.h"
#include "llama.cpp/ggml-impl.h"
#include "llama.cpp/ggml-quants.h"
#include "llama.cpp/ggml.h"
#include "llamafile/sgemm.h"

namespace {
Backend* g_backend = nullptr;
int g_threads = 0;
}

extern "C" {

// Mirrors CPUInfer::CPUInfer(thread_num) (cpuinfer.h:38-45).
int ktref_init(int thread_num) {
    if (g_backend) {
        if (thread_num == g_threads) return g_threads;
        delete g_backend;
        g_backend = nullptr;
    }
    if (thread_num < 1) thread_num = 1;
    g_backend = new Backend(thread_num - 1);
    g_threads = thread_num;
    for (int i = 0; i < (1 << 16); ++i) {
        ggml_table_f32_f16[i] = GGML_COMPUTE_FP16_TO_FP32(i);
    }
    return g_threads;
}

int ktref_threads(void) { return g_threads; }

const char* ktref_isa(void) {
#if defined(__AVX512F__) && defined(__AVX512VNNI__) && defined(__AVX512BF16__)
    return "avx512-vnni-bf16 (llamafile zen4 branch)";
#elif defined(__AVX512F__)
    return "avx512f";
#elif defined(__AVX2__)
    return "avx2";
#else
    return "generic";
#endif
}

long ktref_type_size(int t) { return (long)ggml_type_size((ggml_type)t); }
long ktref_blck_size(int t) { return (long)ggml_blck_size((ggml_type)t); }
int ktref_vec_dot_type(int t) { return (int)ggml_internal_get_type_traits((ggml_type)t).vec_dot_type; }

void ktref_from_float(const float* in, void* out, long n, int t) { from_float(in, out, (int)n, (ggml_type)t); }
void ktref_to_float(const void* in, float* out, long n, int t) { to_float(in, out, (int)n, (ggml_type)t); }

// Full (non-reference) quantiser used by the reference's own tests to mint weights:
// kt_kernel_ext.utils.from_float == type-traits from_float (the *_reference row quantisers).
float ktref_vec_dot(int t, long n, const void* x, const void* y) {
    float s = 0.f;
    ggml_internal_get_type_traits((ggml_type)t).vec_dot((int)n, &s, 0, x, 0, y, 0, 1);
    return s;
}

int ktref_sgemm(long m, long n, long k, const void* A, long lda, const void* B, long ldb, float* C, long ldc,
                int Atype, int Btype) {
    return llamafile_sgemm(m, n, k, A, lda, B, ldb, C, ldc, 0, 1, GGML_TASK_TYPE_COMPUTE, Atype, Btype,
                           GGML_TYPE_F32, GGML_PREC_DEFAULT)
               ? 1
               : 0;
}

void* ktref_moe_create(int expert_num, int routed_expert_num, int hidden_size, int intermediate_size, int stride,
                       int group_min_len, int group_max_len, int use_silu, void* gate, void* up, void* down,
                       int gate_type, int up_type, int down_type, int hidden_type) {
    if (!g_backend) ktref_init(1);
    MOEConfig cfg(expert_num, routed_expert_num, hidden_size, intermediate_size, stride, group_min_len,
                  group_max_len, use_silu != 0, gate, up, down, (ggml_type)gate_type, (ggml_type)up_type,
                  (ggml_type)down_type, (ggml_type)hidden_type);
    return new MOE(cfg);
}

void ktref_moe_warm_up(void* h) { ((MOE*)h)->warm_up(g_backend); }

// MOE::forward reads qlen from batch_size_tensor[0] (moe.cpp:368) and decrements it while chunking.
void ktref_moe_forward(void* h, int qlen, int k, const uint64_t* expert_ids, const float* weights,
                       const void* input, void* output) {
    int bsz = qlen;
    ((MOE*)h)->forward(qlen, k, expert_ids, weights, input, output, &bsz, g_backend);
}

void ktref_moe_destroy(void* h) { delete (MOE*)h; }

void* ktref_linear_create(int input_size, int output_size, int stride, int group_max_len, void* proj,
                          int proj_type, int hidden_type) {
    if (!g_backend) ktref_init(1);
    LinearConfig cfg(input_size, output_size, stride, group_max_len, proj, (ggml_type)proj_type,
                     (ggml_type)hidden_type);
    return new Linear(cfg);
}
void ktref_linear_forward(void* h, int qlen, const void* input, void* output) {
    ((Linear*)h)->forward(qlen, input, output, g_backend);
}
void ktref_linear_destroy(void* h) { delete (Linear*)h; }

void* ktref_mlp_create(int hidden_size, int intermediate_size, int stride, int group_max_len, void* gate, void* up,
                       void* down, int gate_type, int up_type, int down_type, int hidden_type) {
    if (!g_backend) ktref_init(1);
    MLPConfig cfg(hidden_size, intermediate_size, stride, group_max_len, gate, up, down, (ggml_type)gate_type,
                  (ggml_type)up_type, (ggml_type)down_type, (ggml_type)hidden_type);
    return new MLP(cfg);
}
void ktref_mlp_forward(void* h, int qlen, const void* input, void* output) {
    ((MLP*)h)->forward(qlen, input, output, g_backend);
}
void ktref_mlp_destroy(void* h) { delete (MLP*)h; }

}  // extern "C"
