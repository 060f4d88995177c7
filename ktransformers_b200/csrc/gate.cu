// DeepSeek MoEGate routing on the GPU.  Replaces MoEGate.forward
// (archive/ktransformers/models/modeling_deepseek_v3.py:430-481 for V3 sigmoid/noaux_tc,
//  archive/ktransformers/models/modeling_deepseek.py:411-459 for V2 softmax greedy/group_limited_greedy),
// which the reference runs as ~10 ATen kernels per layer (fp32 F.linear, sigmoid, two topk, scatter,
// masked_fill, gather, sum, div, mul).
//
// ONE kernel:
//   phase 1  fp32 GEMV  logits[t][e] = x_t(fp32) . W[e]   (HBM-bound: E*H*4 bytes).  A CTA of 4 warps owns 4
//            expert rows x one column split; each warp streams its row slice with 16-byte loads (8 in flight
//            per lane) against the token slice staged in shared memory, and writes one partial sum per
//            (token, expert, split).  Partials are added in fixed split order: deterministic, no float atomics.
//   phase 2  the LAST CTA to finish (atomic ticket, self-resetting, graph-replay safe) selects, all 128 threads
//            per token — scoring, bias, group top-2 / max, group top-k, expert top-k by iterative arg-max with
//            REDUX max/min (ties -> lowest index), gather, normalise, scale.
#include "common.cuh"

namespace ktb {

constexpr int kGateWarps = 4;
constexpr int kGateThreads = kGateWarps * 32;
constexpr int kGateTokTile = 8;

struct GateParams {
    const float* W;
    const void* x;
    int hidden_type, E, H, T, S;
    int top_k, n_group, topk_group, scoring, topk_method, norm_topk_prob;
    float routed_scaling_factor;
    const float* bias;
    float* partial;      // [T][S][E]
    float* logits_out;   // optional [T][E]
    int64_t* idx;
    float* w;
    const int* bsz;
    unsigned* ticket;
};

// order-preserving float -> uint32 key
__device__ __forceinline__ unsigned fkey(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int kGateEPT = 4;   // experts per thread of the selecting CTA (E <= 512)

__device__ __forceinline__ float fkey_inv(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Selection for ONE token by the whole CTA (kGateThreads threads).  Shared scratch (floats):
//   scores[E] | choice[E] | gsc[32] | wbest[2*kGateWarps] | red[2*kGateWarps]
__device__ void gate_select_token_cta(const GateParams& p, int t, float* sm) {
    const int E = p.E, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* scores = sm;
    float* choice = sm + E;
    float* gsc = choice + E;
    unsigned* wbest = reinterpret_cast<unsigned*>(gsc + 32);
    float* red = reinterpret_cast<float*>(wbest + 2 * kGateWarps);

    // logits = sum of the S partials in fixed order; thread owns experts e = tid + 128*i
    float v[kGateEPT];
#pragma unroll
    for (int i = 0; i < kGateEPT; i++) v[i] = 0.f;
    for (int s = 0; s < p.S; s++) {
        const float* pp = p.partial + ((long)t * p.S + s) * E;
#pragma unroll
        for (int i = 0; i < kGateEPT; i++) {
            const int e = tid + kGateThreads * i;
            if (e < E) v[i] += __ldcg(pp + e);   // written by other SMs in this launch: read at L2
        }
    }
    if (p.logits_out) {
#pragma unroll
        for (int i = 0; i < kGateEPT; i++) {
            const int e = tid + kGateThreads * i;
            if (e < E) p.logits_out[(long)t * E + e] = v[i];
        }
    }
    if (p.scoring == 0) {  // sigmoid
#pragma unroll
        for (int i = 0; i < kGateEPT; i++) v[i] = __fdiv_rn(1.0f, 1.0f + expf(-v[i]));
    } else {               // softmax(dim=-1, fp32): block max, block sum
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < kGateEPT; i++) if (tid + kGateThreads * i < E) m = fmaxf(m, v[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) red[warp] = m;
        __syncthreads();
        m = red[0];
        for (int w = 1; w < kGateWarps; w++) m = fmaxf(m, red[w]);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < kGateEPT; i++) {
            v[i] = (tid + kGateThreads * i < E) ? expf(v[i] - m) : 0.f;
            sum += v[i];
        }
        sum = warp_sum(sum);
        if (lane == 0) red[kGateWarps + warp] = sum;
        __syncthreads();
        sum = 0.f;
        for (int w = 0; w < kGateWarps; w++) sum += red[kGateWarps + w];
#pragma unroll
        for (int i = 0; i < kGateEPT; i++) v[i] = __fdiv_rn(v[i], sum);
    }
    float c[kGateEPT];   // selection scores
#pragma unroll
    for (int i = 0; i < kGateEPT; i++) {
        const int e = tid + kGateThreads * i;
        c[i] = -INFINITY;
        if (e < E) {
            c[i] = v[i] + ((p.topk_method == 0 && p.bias) ? p.bias[e] : 0.f);
            scores[e] = v[i];
            choice[e] = c[i];
        }
    }
    __syncthreads();

    // group selection (noaux_tc: sum of the group's top-2 biased scores; group_limited_greedy: group max)
    if (p.n_group > 1 && p.topk_method != 1) {
        const int gs = E / p.n_group;
        for (int g = warp; g < p.n_group; g += kGateWarps) {   // one warp per group
            float a1 = -INFINITY, a2 = -INFINITY;
            for (int i = lane; i < gs; i += 32) {
                const float x = choice[g * gs + i];
                if (x > a1) { a2 = a1; a1 = x; } else if (x > a2) { a2 = x; }
            }
            const unsigned k1 = fkey(a1);
            const unsigned mx1 = __reduce_max_sync(0xffffffffu, k1);
            const int wl = __ffs(__ballot_sync(0xffffffffu, k1 == mx1)) - 1;
            const unsigned mx2 = __reduce_max_sync(0xffffffffu, lane == wl ? fkey(a2) : k1);
            if (lane == 0) gsc[g] = (p.topk_method == 0) ? (fkey_inv(mx1) + fkey_inv(mx2)) : fkey_inv(mx1);
        }
        __syncthreads();
        const float fill = (p.topk_method == 0) ? -INFINITY : 0.0f;  // V3 masks with -inf, V2 with 0.0
#pragma unroll
        for (int i = 0; i < kGateEPT; i++) {
            const int e = tid + kGateThreads * i;
            if (e < E) {
                const int g = e / gs;
                const float mine = gsc[g];
                int rank = 0;  // higher first, ties -> lower index
                for (int o = 0; o < p.n_group; o++) {
                    const float og = gsc[o];
                    rank += (og > mine || (og == mine && o < g)) ? 1 : 0;
                }
                if (rank >= p.topk_group) { c[i] = fill; choice[e] = fill; }
            }
        }
        __syncthreads();
    }

    // top-k by iterative arg-max; ties -> lowest expert index
    float wsum = 0.f, myw = 0.f;
    long myidx = 0;
    for (int it = 0; it < p.top_k; it++) {
        unsigned bk = 0;
        int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < kGateEPT; i++) {
            const int e = tid + kGateThreads * i;
            const unsigned kk = (e < E) ? fkey(c[i]) : 0u;
            if (kk > bk) { bk = kk; bi = e; }
        }
        const unsigned mx = __reduce_max_sync(0xffffffffu, bk);
        const int wi = __reduce_min_sync(0xffffffffu, (bk == mx) ? bi : 0x7fffffff);
        if (lane == 0) { wbest[2 * warp] = mx; wbest[2 * warp + 1] = (unsigned)wi; }
        __syncthreads();
        unsigned gk = wbest[0];
        int win = (int)wbest[1];
#pragma unroll
        for (int w = 1; w < kGateWarps; w++) {
            const unsigned kk = wbest[2 * w];
            const int ii = (int)wbest[2 * w + 1];
            if (kk > gk || (kk == gk && ii < win)) { gk = kk; win = ii; }
        }
        if (win == 0x7fffffff || win < 0 || win >= E) win = 0;  // degenerate (all NaN)
        // V3 gathers the weight from the un-biased scores; V2 group_limited takes the (masked) score itself
        const float wv = (p.topk_method == 2) ? choice[win] : scores[win];
#pragma unroll
        for (int i = 0; i < kGateEPT; i++)
            if (tid + kGateThreads * i == win) c[i] = -INFINITY;
        if (tid == it) { myw = wv; myidx = win; }
        wsum += wv;
        __syncthreads();   // wbest is rewritten in the next iteration
    }
    // V3 (modeling_deepseek_v3.py:474-479): normalise (if top_k>1 && norm_topk_prob) THEN always scale;
    // V2 (modeling_deepseek.py:455-459): normalise XOR scale.
    if (tid < p.top_k) {
        float w = myw;
        const bool do_norm = p.top_k > 1 && p.norm_topk_prob;
        if (do_norm) w = __fdiv_rn(w, wsum + 1e-20f);
        if (p.topk_method == 0 || !do_norm) w = w * p.routed_scaling_factor;
        p.idx[(long)t * p.top_k + tid] = myidx;
        p.w[(long)t * p.top_k + tid] = w;
    }
    __syncthreads();
}

template <int NT>
__device__ __forceinline__ void gate_dot(const GateParams& p, const float4* wrow, const float* xs, int nc4, int lane, int t0, int e,
                                         int s, int S, int nt = NT) {
    float acc[NT];
#pragma unroll
    for (int i = 0; i < NT; i++) acc[i] = 0.f;
    for (int cb = lane; cb < nc4; cb += 32 * 8) {
        float4 w[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int c = cb + 32 * u;
            if (c < nc4) w[u] = __ldg(wrow + c);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int c = cb + 32 * u;
            if (c < nc4) {
#pragma unroll
                for (int i = 0; i < NT; i++) {
                    const float4 xv = reinterpret_cast<const float4*>(xs + (size_t)(i < nt ? i : 0) * nc4 * 4)[c];
                    acc[i] = fmaf(w[u].x, xv.x, acc[i]);
                    acc[i] = fmaf(w[u].y, xv.y, acc[i]);
                    acc[i] = fmaf(w[u].z, xv.z, acc[i]);
                    acc[i] = fmaf(w[u].w, xv.w, acc[i]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NT; i++) {
        const float v = warp_sum(acc[i]);
        if (i < nt && lane == 0) p.partial[((long)(t0 + i) * S + s) * p.E + e] = v;
    }
}

__global__ void __launch_bounds__(kGateThreads) gate_kernel(const GateParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float* xs = reinterpret_cast<float*>(smem_raw);   // [tok tile][slice cols]  (phase 1) / scores+choice (phase 2)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int Teff = p.T;
    if (p.bsz) Teff = min(p.T, *p.bsz);
    const int s = blockIdx.y, S = p.S;
    const int n4 = p.H / 4;
    const int c0 = (int)((long)n4 * s / S), c1 = (int)((long)n4 * (s + 1) / S);   // float4 column range of this split
    const int nc4 = c1 - c0;
    const int e = blockIdx.x * kGateWarps + warp;
    const float4* wrow = reinterpret_cast<const float4*>(p.W + (long)(e < p.E ? e : 0) * p.H) + c0;

    for (int t0 = 0; t0 < Teff; t0 += kGateTokTile) {
        const int nt = min(kGateTokTile, Teff - t0);
        __syncthreads();
        for (int tt = 0; tt < nt; tt++)
            for (int c = threadIdx.x; c < nc4 * 4; c += kGateThreads)
                xs[tt * nc4 * 4 + c] = load_hidden(p.x, (long)(t0 + tt) * p.H + 4L * c0 + c, p.hidden_type);
        __syncthreads();
        if (e < p.E) {
            switch (nt) {   // compile-time token counts: no predicated-off FMAs for the common bs=1 decode
                case 1: gate_dot<1>(p, wrow, xs, nc4, lane, t0, e, s, S); break;
                case 2: gate_dot<2>(p, wrow, xs, nc4, lane, t0, e, s, S); break;
                case 3: case 4: gate_dot<4>(p, wrow, xs, nc4, lane, t0, e, s, S, nt); break;
                default: gate_dot<kGateTokTile>(p, wrow, xs, nc4, lane, t0, e, s, S, nt); break;
            }
        }
    }

    // ---- last CTA selects -------------------------------------------------------------------------------
    __shared__ unsigned s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y;
        const unsigned prev = atomicAdd(p.ticket, 1u);
        s_last = (prev == total - 1);
        if (s_last) *p.ticket = 0;   // self-reset: the next launch (or graph replay) starts from zero
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int t = 0; t < Teff; t++) gate_select_token_cta(p, t, xs);
}

// per-device scratch: partial sums + ticket
static float* g_partial[64] = {nullptr};
static size_t g_partial_cap[64] = {0};
static unsigned* g_ticket[64] = {nullptr};

}  // namespace ktb

extern "C" int ktb200_moe_gate_forward(const ktb200_gate_config* c, int qlen, const void* x, int64_t* idx, float* w,
                                       float* logits, const int* bsz, void* stream) {
    using namespace ktb;
    if (!c || !x || !idx || !w) { set_error("null pointer"); return KTB200_EINVAL; }
    if (qlen <= 0) return KTB200_OK;
    if (c->n_experts > kGateThreads * kGateEPT) { set_error("gate: at most %d experts", kGateThreads * kGateEPT); return KTB200_EINVAL; }
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
    const int d = dev & 63;
    cudaStream_t s = (cudaStream_t)stream;
    const int row_ctas = (c->n_experts + kGateWarps - 1) / kGateWarps;
    int S = (2 * num_sms(dev) + row_ctas - 1) / row_ctas;   // ~2 CTAs per SM
    if (S < 1) S = 1;
    if (S > 8) S = 8;
    while (S > 1 && c->hidden_size / 4 / S < 64) S--;
    const size_t need = (size_t)qlen * c->n_experts * S * sizeof(float);
    if (need > g_partial_cap[d] || !g_ticket[d]) {
        // grow-only scratch; allocation is NOT capturable: call once with the largest qlen before graph capture
        if (g_partial[d]) cudaFree(g_partial[d]);
        const size_t cap = need < (1u << 20) ? (1u << 20) : need;
        KTB_CUDA_CHECK(cudaMalloc(&g_partial[d], cap));
        g_partial_cap[d] = cap;
        if (!g_ticket[d]) {
            KTB_CUDA_CHECK(cudaMalloc(&g_ticket[d], sizeof(unsigned)));
            KTB_CUDA_CHECK(cudaMemset(g_ticket[d], 0, sizeof(unsigned)));
        }
    }
    GateParams p{c->weight, x, c->hidden_type, c->n_experts, c->hidden_size, qlen, S, c->top_k, c->n_group, c->topk_group,
                 c->scoring, c->topk_method, c->norm_topk_prob, c->routed_scaling_factor, c->bias, g_partial[d], logits,
                 idx, w, bsz, g_ticket[d]};
    const int nc4_max = c->hidden_size / 4 / S + 1;
    const int nt = qlen < kGateTokTile ? qlen : kGateTokTile;
    size_t smem = (size_t)nt * nc4_max * 16;
    const size_t smem_sel = ((size_t)2 * c->n_experts + 32 + 4 * kGateWarps) * sizeof(float);
    if (smem_sel > smem) smem = smem_sel;
    if (smem > 48 * 1024) KTB_CUDA_CHECK(cudaFuncSetAttribute(gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gate_kernel<<<dim3(row_ctas, S), kGateThreads, smem, s>>>(p);
    KTB_LAUNCH_CHECK();
    return KTB200_OK;
}
