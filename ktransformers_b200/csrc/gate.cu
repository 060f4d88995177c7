// DeepSeek MoEGate routing on the GPU.  Replaces MoEGate.forward
// (archive/ktransformers/models/modeling_deepseek_v3.py:430-481 for V3 sigmoid/noaux_tc,
//  archive/ktransformers/models/modeling_deepseek.py:411-459 for V2 softmax greedy/group_limited_greedy),
// which the reference runs as ~10 ATen kernels per layer (fp32 F.linear, sigmoid, two topk, scatter,
// masked_fill, gather, sum, div, mul).
//
//   gate_logits_kernel : fp32 GEMV  logits[t][e] = x_t(fp32) . W[e]   (HBM-bound: E*H*4 bytes)
//                        grid (E, S): each CTA streams 1/S of one expert row with float4 loads and
//                        reuses it for every token; S partial sums per logit are added in fixed order
//                        by the selection kernel (deterministic, no atomics).
//   gate_select_kernel : one warp per token: scoring, bias, group top-2 / max, group top-k, expert
//                        top-k by iterative warp arg-max (ties -> lowest index), gather, normalise, scale.
#include "common.cuh"

namespace ktb {

constexpr int kGateThreads = 128;
constexpr int kGateTokTile = 8;

__global__ void __launch_bounds__(kGateThreads) gate_logits_kernel(const float* __restrict__ W, const void* __restrict__ x,
                                                                   int hidden_type, int E, int H, int T,
                                                                   float* __restrict__ partial /*[T][E][S]*/,
                                                                   const int* bsz) {
    const int e = blockIdx.x, s = blockIdx.y, S = gridDim.y;
    int Teff = T;
    if (bsz) Teff = min(T, *bsz);
    // column range of this split, in float4 units
    const int n4 = H / 4;
    const int c0 = (int)((long)n4 * s / S), c1 = (int)((long)n4 * (s + 1) / S);
    const float4* wrow = reinterpret_cast<const float4*>(W + (long)e * H);
    __shared__ float red[kGateTokTile][kGateThreads / 32];
    for (int t0 = 0; t0 < Teff; t0 += kGateTokTile) {
        float acc[kGateTokTile];
#pragma unroll
        for (int i = 0; i < kGateTokTile; i++) acc[i] = 0.f;
        for (int c = c0 + threadIdx.x; c < c1; c += kGateThreads) {
            const float4 w = __ldg(wrow + c);
#pragma unroll
            for (int i = 0; i < kGateTokTile; i++) {
                const int t = t0 + i;
                if (t < Teff) {
                    const long base = (long)t * H + 4L * c;
                    const float x0 = load_hidden(x, base, hidden_type), x1 = load_hidden(x, base + 1, hidden_type);
                    const float x2 = load_hidden(x, base + 2, hidden_type), x3 = load_hidden(x, base + 3, hidden_type);
                    acc[i] = fmaf(w.x, x0, acc[i]);
                    acc[i] = fmaf(w.y, x1, acc[i]);
                    acc[i] = fmaf(w.z, x2, acc[i]);
                    acc[i] = fmaf(w.w, x3, acc[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < kGateTokTile; i++) {
            const float v = warp_sum(acc[i]);
            if ((threadIdx.x & 31) == 0) red[i][threadIdx.x >> 5] = v;
        }
        __syncthreads();
        if (threadIdx.x < kGateTokTile) {
            const int t = t0 + threadIdx.x;
            if (t < Teff) {
                float v = 0.f;
                for (int w = 0; w < kGateThreads / 32; w++) v += red[threadIdx.x][w];
                partial[((long)t * E + e) * S + s] = v;
            }
        }
        __syncthreads();
    }
}

struct GateSelParams {
    int E, top_k, n_group, topk_group, scoring, topk_method, norm_topk_prob, S;
    float routed_scaling_factor;
    const float* bias;
    const float* partial;
    float* logits_out;  // optional [T][E]
    int64_t* idx;
    float* w;
    const int* bsz;
};

// one warp per token; dynamic smem: scores[E] | choice[E]
__global__ void __launch_bounds__(32) gate_select_kernel(const GateSelParams p) {
    extern __shared__ float sm[];
    const int t = blockIdx.x, lane = threadIdx.x;
    if (p.bsz && t >= *p.bsz) return;
    float* scores = sm;
    float* choice = sm + p.E;
    const int E = p.E;

    // logits (+ optional export), scoring
    float lmax = -INFINITY;
    for (int e = lane; e < E; e += 32) {
        float v = 0.f;
        const float* pp = p.partial + ((long)t * E + e) * p.S;
        for (int s = 0; s < p.S; s++) v += pp[s];
        if (p.logits_out) p.logits_out[(long)t * E + e] = v;
        scores[e] = v;
        lmax = fmaxf(lmax, v);
    }
    if (p.scoring == 0) {  // sigmoid
        for (int e = lane; e < E; e += 32) scores[e] = __fdiv_rn(1.0f, 1.0f + expf(-scores[e]));
    } else {               // softmax(dim=-1, fp32)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        float sum = 0.f;
        for (int e = lane; e < E; e += 32) { const float ex = expf(scores[e] - lmax); scores[e] = ex; sum += ex; }
        sum = warp_sum(sum);
        for (int e = lane; e < E; e += 32) scores[e] = __fdiv_rn(scores[e], sum);
    }
    for (int e = lane; e < E; e += 32) choice[e] = scores[e] + ((p.topk_method == 0 && p.bias) ? p.bias[e] : 0.f);
    __syncwarp();

    // group selection
    if (p.n_group > 1 && p.topk_method != 1) {
        const int gs = E / p.n_group;
        float gscore = -INFINITY;
        if (lane < p.n_group) {
            float m1 = -INFINITY, m2 = -INFINITY;
            for (int i = 0; i < gs; i++) {
                const float v = choice[lane * gs + i];
                if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) { m2 = v; }
            }
            gscore = (p.topk_method == 0) ? (m1 + m2) : m1;  // noaux_tc: top-2 sum ; group_limited_greedy: max
        }
        // rank among groups (higher first, ties -> lower index)
        int rank = 0;
        for (int g = 0; g < p.n_group; g++) {
            const float og = __shfl_sync(0xffffffffu, gscore, g);
            if (lane < p.n_group && (og > gscore || (og == gscore && g < lane))) rank++;
        }
        const unsigned sel = __ballot_sync(0xffffffffu, lane < p.n_group && rank < p.topk_group);
        const float fill = (p.topk_method == 0) ? -INFINITY : 0.0f;  // V3 masks with -inf, V2 with 0.0
        for (int e = lane; e < E; e += 32)
            if (!((sel >> (e / gs)) & 1u)) choice[e] = fill;
        __syncwarp();
    }

    // top-k by iterative arg-max; ties -> lowest expert index
    float wsum = 0.f;
    float myw = 0.f;  // lane i keeps weight i (top_k <= 32)
    long myidx = 0;
    for (int i = 0; i < p.top_k; i++) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int e = lane; e < E; e += 32) {
            const float v = choice[e];
            if (v > bv || (v == bv && e < bi)) { bv = v; bi = e; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (bi == 0x7fffffff) bi = 0;  // all -inf/NaN: degenerate
        // V3 gathers from the un-biased scores; V2 takes the (masked) score itself
        const float wv = (p.topk_method == 2) ? bv : scores[bi];
        if (lane == i) { myw = wv; myidx = bi; }
        wsum += wv;
        if (lane == 0) choice[bi] = -INFINITY;
        __syncwarp();
    }
    // normalisation / scaling: V3 (modeling_deepseek_v3.py:474-479): norm (if top_k>1 && norm) THEN always scale;
    // V2 (modeling_deepseek.py:455-459): norm XOR scale.
    // torch sums topk_weight in index order 0..k-1; wsum above is that same order.
    if (lane < p.top_k) {
        float w = myw;
        const bool do_norm = p.top_k > 1 && p.norm_topk_prob;
        if (do_norm) w = __fdiv_rn(w, wsum + 1e-20f);
        if (p.topk_method == 0 || !do_norm) w = w * p.routed_scaling_factor;
        p.idx[(long)t * p.top_k + lane] = myidx;
        p.w[(long)t * p.top_k + lane] = w;
    }
}

// per-device scratch for the S partial sums
static float* g_partial[64] = {nullptr};
static size_t g_partial_cap[64] = {0};

}  // namespace ktb

extern "C" int ktb200_moe_gate_forward(const ktb200_gate_config* c, int qlen, const void* x, int64_t* idx, float* w,
                                       float* logits, const int* bsz, void* stream) {
    using namespace ktb;
    if (!c || !x || !idx || !w) { set_error("null pointer"); return KTB200_EINVAL; }
    if (qlen <= 0) return KTB200_OK;
    if (c->n_experts <= 0 || c->hidden_size <= 0 || c->hidden_size % 4 || c->top_k <= 0 || c->top_k > 32 || c->top_k > c->n_experts) {
        set_error("gate: bad shape (E=%d H=%d top_k=%d; top_k<=32, H%%4==0)", c->n_experts, c->hidden_size, c->top_k);
        return KTB200_EINVAL;
    }
    if (c->n_group < 1 || c->n_group > 32 || c->n_experts % c->n_group || c->topk_group < 1 || c->topk_group > c->n_group) {
        set_error("gate: bad grouping (n_group=%d topk_group=%d)", c->n_group, c->topk_group);
        return KTB200_EINVAL;
    }
    if (c->scoring < 0 || c->scoring > 1 || c->topk_method < 0 || c->topk_method > 2) { set_error("gate: bad scoring/topk_method"); return KTB200_EINVAL; }
    if (!is_hidden_type(c->hidden_type) || !c->weight) { set_error("gate: bad hidden_type or null weight"); return KTB200_EINVAL; }
    int dev = 0;
    KTB_CUDA_CHECK(cudaGetDevice(&dev));
    cudaStream_t s = (cudaStream_t)stream;
    // splits so that E*S ~ 4 CTAs per SM
    int S = (4 * num_sms(dev) + c->n_experts - 1) / c->n_experts;
    if (S < 1) S = 1;
    if (S > 16) S = 16;
    while (S > 1 && c->hidden_size / 4 / S < 32) S--;
    const size_t need = (size_t)qlen * c->n_experts * S * sizeof(float);
    if (need > g_partial_cap[dev & 63]) {
        // grow-only scratch; allocation is NOT capturable, so warm up once with the largest qlen before graph capture
        if (g_partial[dev & 63]) cudaFree(g_partial[dev & 63]);
        size_t cap = need < (1u << 20) ? (1u << 20) : need;
        KTB_CUDA_CHECK(cudaMalloc(&g_partial[dev & 63], cap));
        g_partial_cap[dev & 63] = cap;
    }
    float* partial = g_partial[dev & 63];
    gate_logits_kernel<<<dim3(c->n_experts, S), kGateThreads, 0, s>>>(c->weight, x, c->hidden_type, c->n_experts,
                                                                     c->hidden_size, qlen, partial, bsz);
    KTB_LAUNCH_CHECK();
    GateSelParams p{c->n_experts, c->top_k, c->n_group, c->topk_group, c->scoring, c->topk_method, c->norm_topk_prob, S,
                    c->routed_scaling_factor, c->bias, partial, logits, idx, w, bsz};
    gate_select_kernel<<<qlen, 32, 2 * c->n_experts * sizeof(float), s>>>(p);
    KTB_LAUNCH_CHECK();
    return KTB200_OK;
}
