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
//   phase 2  the LAST CTA to finish (atomic ticket, self-resetting, graph-replay safe) selects: one warp per
//            token — scoring, bias, group top-2 / max, group top-k, expert top-k by iterative arg-max with
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

constexpr int kGateEPL = 16;   // experts per lane held in registers (E <= 512)

__device__ void gate_select_token(const GateParams& p, int t, int lane, float* scores, float* choice) {
    const int E = p.E;
    // logits = sum of the S partials (fixed order), lane owns experts e = lane + 32*i
    float v[kGateEPL];
#pragma unroll
    for (int i = 0; i < kGateEPL; i++) v[i] = 0.f;
    for (int s = 0; s < p.S; s++) {
        const float* pp = p.partial + ((long)t * p.S + s) * E;
#pragma unroll
        for (int i = 0; i < kGateEPL; i++) {
            const int e = lane + 32 * i;
            if (e < E) v[i] += __ldcg(pp + e);   // written by other SMs in this launch: read at L2
        }
    }
    float lmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < kGateEPL; i++) {
        const int e = lane + 32 * i;
        if (e < E) {
            if (p.logits_out) p.logits_out[(long)t * E + e] = v[i];
            lmax = fmaxf(lmax, v[i]);
        }
    }
    if (p.scoring == 0) {  // sigmoid
#pragma unroll
        for (int i = 0; i < kGateEPL; i++) v[i] = __fdiv_rn(1.0f, 1.0f + expf(-v[i]));
    } else {               // softmax(dim=-1, fp32)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < kGateEPL; i++) {
            const int e = lane + 32 * i;
            v[i] = (e < E) ? expf(v[i] - lmax) : 0.f;
            sum += v[i];
        }
        sum = warp_sum(sum);
#pragma unroll
        for (int i = 0; i < kGateEPL; i++) v[i] = __fdiv_rn(v[i], sum);
    }
    float c[kGateEPL];   // selection scores
#pragma unroll
    for (int i = 0; i < kGateEPL; i++) {
        const int e = lane + 32 * i;
        c[i] = -INFINITY;
        if (e < E) {
            c[i] = v[i] + ((p.topk_method == 0 && p.bias) ? p.bias[e] : 0.f);
            scores[e] = v[i];
            choice[e] = c[i];
        }
    }
    __syncwarp();

    // group selection (noaux_tc: sum of the group's top-2 biased scores; group_limited_greedy: group max)
    if (p.n_group > 1 && p.topk_method != 1) {
        const int gs = E / p.n_group;
        float gscore = -INFINITY;
        const bool pow2 = (p.n_group & (p.n_group - 1)) == 0;
        if (pow2) {
            // 32/n_group lanes cooperate on one group, then merge their (max1, max2) pairs
            const int lpg = 32 / p.n_group, g = lane / lpg, sub = lane % lpg;
            float m1 = -INFINITY, m2 = -INFINITY;
            for (int i = sub; i < gs; i += lpg) {
                const float x = choice[g * gs + i];
                if (x > m1) { m2 = m1; m1 = x; } else if (x > m2) { m2 = x; }
            }
            for (int o = 1; o < lpg; o <<= 1) {
                const float o1 = __shfl_xor_sync(0xffffffffu, m1, o), o2 = __shfl_xor_sync(0xffffffffu, m2, o);
                const float hi = fmaxf(m1, o1), lo = fminf(m1, o1);
                m2 = fmaxf(lo, fmaxf(m2, o2));
                m1 = hi;
            }
            const float gsc = (p.topk_method == 0) ? (m1 + m2) : m1;
            gscore = __shfl_sync(0xffffffffu, gsc, (lane % p.n_group) * lpg);   // lane g (< n_group) gets group g's score
            if (lane >= p.n_group) gscore = -INFINITY;
        } else if (lane < p.n_group) {
            float m1 = -INFINITY, m2 = -INFINITY;
            for (int i = 0; i < gs; i++) {
                const float x = choice[lane * gs + i];
                if (x > m1) { m2 = m1; m1 = x; } else if (x > m2) { m2 = x; }
            }
            gscore = (p.topk_method == 0) ? (m1 + m2) : m1;
        }
        int rank = 0;  // higher first, ties -> lower index
        for (int g = 0; g < p.n_group; g++) {
            const float og = __shfl_sync(0xffffffffu, gscore, g);
            if (lane < p.n_group && (og > gscore || (og == gscore && g < lane))) rank++;
        }
        const unsigned sel = __ballot_sync(0xffffffffu, lane < p.n_group && rank < p.topk_group);
        const float fill = (p.topk_method == 0) ? -INFINITY : 0.0f;  // V3 masks with -inf, V2 with 0.0
#pragma unroll
        for (int i = 0; i < kGateEPL; i++) {
            const int e = lane + 32 * i;
            if (e < E && !((sel >> (e / gs)) & 1u)) c[i] = fill;
        }
    }

    // top-k by iterative arg-max over the register-resident scores; ties -> lowest expert index
    float wsum = 0.f, myw = 0.f;
    long myidx = 0;
    for (int it = 0; it < p.top_k; it++) {
        unsigned bk = 0;
        int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < kGateEPL; i++) {
            const int e = lane + 32 * i;
            const unsigned kk = (e < E) ? fkey(c[i]) : 0u;
            if (kk > bk) { bk = kk; bi = e; }
        }
        const unsigned mx = __reduce_max_sync(0xffffffffu, bk);
        int win = __reduce_min_sync(0xffffffffu, (bk == mx) ? bi : 0x7fffffff);
        if (win == 0x7fffffff) win = 0;  // degenerate (all NaN)
        // the winner's owner broadcasts its selection score; V3 gathers the weight from the un-biased scores,
        // V2 group_limited takes the (masked) score itself
        float cw = 0.f;
#pragma unroll
        for (int i = 0; i < kGateEPL; i++)
            if (lane + 32 * i == win) { cw = c[i]; c[i] = -INFINITY; }
        cw = __shfl_sync(0xffffffffu, cw, win & 31);
        const float wv = (p.topk_method == 2) ? cw : scores[win];
        if (lane == it) { myw = wv; myidx = win; }
        wsum += wv;
    }
    // V3 (modeling_deepseek_v3.py:474-479): normalise (if top_k>1 && norm_topk_prob) THEN always scale;
    // V2 (modeling_deepseek.py:455-459): normalise XOR scale.
    if (lane < p.top_k) {
        float w = myw;
        const bool do_norm = p.top_k > 1 && p.norm_topk_prob;
        if (do_norm) w = __fdiv_rn(w, wsum + 1e-20f);
        if (p.topk_method == 0 || !do_norm) w = w * p.routed_scaling_factor;
        p.idx[(long)t * p.top_k + lane] = myidx;
        p.w[(long)t * p.top_k + lane] = w;
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
            float acc[kGateTokTile];
#pragma unroll
            for (int i = 0; i < kGateTokTile; i++) acc[i] = 0.f;
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
                        for (int i = 0; i < kGateTokTile; i++) {
                            if (i < nt) {
                                const float4 xv = reinterpret_cast<const float4*>(xs + (size_t)i * nc4 * 4)[c];
                                acc[i] = fmaf(w[u].x, xv.x, acc[i]);
                                acc[i] = fmaf(w[u].y, xv.y, acc[i]);
                                acc[i] = fmaf(w[u].z, xv.z, acc[i]);
                                acc[i] = fmaf(w[u].w, xv.w, acc[i]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < kGateTokTile; i++) {
                if (i < nt) {
                    const float v = warp_sum(acc[i]);
                    if (lane == 0) p.partial[((long)(t0 + i) * S + s) * p.E + e] = v;
                }
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
    float* sc = xs;                 // per warp: scores[E] | choice[E]
    for (int t = warp; t < Teff; t += kGateWarps) gate_select_token(p, t, lane, sc + (size_t)warp * 2 * p.E, sc + (size_t)warp * 2 * p.E + p.E);
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
    if (c->n_experts > 32 * kGateEPL) { set_error("gate: at most %d experts", 32 * kGateEPL); return KTB200_EINVAL; }
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
    const size_t smem_sel = (size_t)kGateWarps * 2 * c->n_experts * sizeof(float);
    if (smem_sel > smem) smem = smem_sel;
    if (smem > 48 * 1024) KTB_CUDA_CHECK(cudaFuncSetAttribute(gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gate_kernel<<<dim3(row_ctas, S), kGateThreads, smem, s>>>(p);
    KTB_LAUNCH_CHECK();
    return KTB200_OK;
}
