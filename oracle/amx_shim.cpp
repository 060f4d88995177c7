// TEST INFRASTRUCTURE ONLY — extern "C" face over the UNMODIFIED kt-kernel AMX MoE backend (the "CPU-AMX" baseline BASELINE.json's
// north_star names): TP_MOE<AMX_MOE_TP<amx::GemmKernel224Int4>> == kt_kernel_ext.moe.AMXInt4_MOE
// (kt-kernel/ext_bindings.cpp:860; kt-kernel/operators/amx/moe.hpp, moe_base.hpp:455-654 forward_decode), driven like
// kt-kernel/examples/test_moe_amx.py:140-230: WorkerPool(threads) -> MOEConfig{pool, bf16 gate/up/down} -> load_weights()
// (online INT4 quantisation from bf16) -> forward(qlen_ptr, k, ids, weights, input bf16, output bf16).
// Compiled from the sources where they lie under /root/reference with oracle/amx_shim/{numa,hwloc}.h standing in for the
// missing libnuma / hwloc (single node): a SHIMMED build, say so wherever its numbers are reported.  Runs only on hosts with
// amx_tile + amx_int8 (Intel Sapphire Rapids and later).
#include <cstdint>
#include <cstdio>

#include "cpu_backend/worker_pool.h"
#include "llama.cpp/ggml-impl.h"
#include "operators/amx/moe.hpp"

namespace {
WorkerPool* g_pool = nullptr;
int g_threads = 0;
}
using AmxInt4 = TP_MOE<AMX_MOE_TP<amx::GemmKernel224Int4>>;

extern "C" {
int ktamx_init(int threads) {
    if (g_pool && threads == g_threads) return g_threads;
    if (g_pool) { delete g_pool; g_pool = nullptr; }
    g_pool = new WorkerPool(threads);
    g_threads = threads;
    for (int i = 0; i < (1 << 16); ++i) ggml_table_f32_f16[i] = GGML_COMPUTE_FP16_TO_FP32(i);   // cpuinfer.h:42-44
    return g_threads;
}
void* ktamx_moe_create(int E, int k, int H, int I, int max_len, void* gate_bf16, void* up_bf16, void* down_bf16) {
    GeneralMOEConfig c(E, k, H, I);
    c.max_len = max_len; c.gate_proj = gate_bf16; c.up_proj = up_bf16; c.down_proj = down_bf16; c.pool = g_pool;
    auto* m = new AmxInt4(c);
    m->load_weights();
    return m;
}
void ktamx_moe_forward(void* h, int qlen, int k, const int64_t* ids, const float* w, const void* in_bf16, void* out_bf16) {
    int q = qlen;
    ((AmxInt4*)h)->forward(&q, k, ids, w, in_bf16, out_bf16, false);
}
void ktamx_moe_destroy(void* h) { delete (AmxInt4*)h; }
}
