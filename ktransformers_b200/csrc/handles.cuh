// Internal (not part of the C-ABI): the opaque handles behind include/ktb200.h and the weight-format ids the
// dispatchers in moe.cu / moe_block.cu agree on.
#pragma once
#include "common.cuh"

namespace ktb {
enum FmtId { FMT_Q4K, FMT_Q5K, FMT_Q6K8, FMT_Q6K4T, FMT_GENK, FMT_NONE };
// how a Q6_K tensor was re-laid at load time
enum Q6Layout { LAYOUT_RAW = 0, LAYOUT_SOA8 = 1, LAYOUT_T4 = 2 };

static inline FmtId pick_fmt(int type, int layout) {
    if (type == KTB200_TYPE_Q4_K) return FMT_Q4K;
    if (type == KTB200_TYPE_Q5_K) return FMT_Q5K;
    if (type == KTB200_TYPE_Q6_K && layout == LAYOUT_SOA8) return FMT_Q6K8;
    if (type == KTB200_TYPE_Q6_K && layout == LAYOUT_T4) return FMT_Q6K4T;
    if (is_kquant(type)) return FMT_GENK;
    return FMT_NONE;
}
}  // namespace ktb

// ------------------------------------------------------------------------------------------
struct ktb200_mlp {
    int H, I, gate_type, up_type, down_type, hidden_type, group_max_len, device;
    const void *gate, *up, *down;
    bool loaded, gu_soa;
    int down_layout;   // Q6Layout of the down tensor
    float* inter;
};

struct ktb200_moe {
    ktb200_moe_config cfg;
    int device;
    bool loaded;
    bool gu_soa;
    int down_layout;   // Q6Layout of the down tensor
    float* inter;      // [group_max_len * k][I]
    // host-call staging
    int64_t* ids_d;
    float* w_d;
    void* in_d;
    void* out_d;
    // scratch of the persistent MoE-block kernel (moe_block.cu): router partial sums [8 tokens][8 splits][512 experts]
    // and two pairs of grid-barrier words used alternately (zero between launches)
    float* blk_partial;
    unsigned* blk_sync;
    unsigned blk_flip;
    const void* pf[3];       // ktb200_moe_block_prefetch_hint: ranges the block kernel pulls into L2 during its down phase
    size_t pf_bytes[3];
};

struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(0), ok(true) {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

