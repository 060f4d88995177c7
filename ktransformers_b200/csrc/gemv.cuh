// Streaming integer GEMV kernels for quantised-MoE decode (sm_100a).
//
// Both kernels are HBM-bound byte streamers (3.1 FLOP/B at bs=1): every weight byte is read exactly
// once with 16-byte read-only loads that bypass L1, many loads are put in flight per lane before any
// is consumed, and the arithmetic is int8 dp4a against activations staged in shared memory.
// No tensor cores: at M = 1..8 rows per expert the MMA tile would be >87% padding and the kernel
// would still be bound by the same bytes.
//
//   rows_kernel  ("phase 1" of MOE::forward_one, moe.cpp:171-210; Linear::forward_many, linear.cpp:37-63)
//       out[slot][r] = epilogue( W0[e_slot][r,:] . xq_t , W1[e_slot][r,:] . xq_t )
//       one activation row per token shared by all of the token's slots.
//   reduce_kernel ("phase 2", moe.cpp:216-245; MLP down, mlp.cpp:99-117)
//       out[t][h] = sum_j w[t][j] * ( W[e_tj][h,:] . aq_tj )   — aq quantised in the prologue
//
// grid = (gx, T): blockIdx.y is the token, blockIdx.x splits that token's work into contiguous,
// equally sized ranges (gx is chosen so gx*T ~ 2 CTAs per SM).
#pragma once
#include "act_quant.cuh"
#include "formats.cuh"

namespace ktb {

constexpr int kGemvThreads = 256;

struct RowsParams {
    const void* w0;         // [E][rows][ncols] blocks
    const void* w1;         // second matrix (PAIR) or null
    int type0, type1;       // ggml types (FmtGenK reads them at run time)
    int n_experts;          // E behind w0/w1 (1 for a dense linear)
    int rows, ncols;        // per-expert matrix shape (out features, in features)
    int slots;              // slots per token (k for MoE, 1 for dense)
    const int64_t* ids;     // [T][slots] expert ids or null (dense: expert 0)
    int id_offset;          // expert-parallel shard offset
    const void* x;          // [T][ncols] hidden_type
    int hidden_type;
    int use_silu;
    float* out_f32;         // PAIR: [T*slots][rows] fp32 act(g)*u ; else optional fp32 out
    void* out_hidden;       // !PAIR: [T][rows] hidden_type (slots must be 1) or null
    const float* bias;      // !PAIR optional [rows]
    const int* bsz;         // optional device batch size
};

template <class Fmt, bool PAIR, int RW, int NB>
__global__ void __launch_bounds__(kGemvThreads, 2) rows_kernel(const RowsParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int t = blockIdx.y;
    if (p.bsz && t >= *p.bsz) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kGemvThreads / 32;
    const int nblk = p.ncols / QK_K;
    uint8_t* q8 = smem;                                                        // ncols
    float* dx = reinterpret_cast<float*>(smem + p.ncols);                      // nblk
    int16_t* bsums = reinterpret_cast<int16_t*>(smem + p.ncols + nblk * 4);    // ncols/16
    cta_quantize_q8k_rows<4>(p.x, (long)t * p.ncols, 0, p.hidden_type, 1, p.ncols, 0u, q8, dx, bsums);
    __syncthreads();
    const ActQ8K act{q8, dx, bsums};

    const int upr = nblk * Fmt::kUnitsPerBlock;  // units per row
    const long total = (long)p.slots * p.rows;
    const long u0 = total * blockIdx.x / gridDim.x, u1 = total * (blockIdx.x + 1) / gridDim.x;
    constexpr int NM = PAIR ? 2 : 1;

    for (long ub = u0 + (long)warp * RW; ub < u1; ub += (long)nwarps * RW) {
        typename Fmt::Row rp[RW][NM];
        bool valid[RW];
        float acc[RW][NM];
#pragma unroll
        for (int rw = 0; rw < RW; rw++) {
            const long u = ub + rw;
            valid[rw] = false;
#pragma unroll
            for (int m = 0; m < NM; m++) acc[rw][m] = 0.f;
            if (u < u1) {
                const int s = (int)(u / p.rows), r = (int)(u % p.rows);
                long e = p.ids ? (long)p.ids[(long)t * p.slots + s] - p.id_offset : 0;
                if (e >= 0 && e < p.n_experts) {
                    valid[rw] = true;
                    rp[rw][0] = Fmt::row(p.w0, e * p.rows + r, p.ncols, p.type0);
                    if (PAIR) rp[rw][NM - 1] = Fmt::row(p.w1, e * p.rows + r, p.ncols, p.type1);
                }
            }
        }
        for (int c0 = 0; c0 < upr; c0 += 32 * NB) {
            typename Fmt::Regs regs[RW][NM][NB];
#pragma unroll
            for (int rw = 0; rw < RW; rw++)
#pragma unroll
                for (int nb = 0; nb < NB; nb++) {
                    const int unit = c0 + nb * 32 + lane;
                    if (valid[rw] && unit < upr) {
#pragma unroll
                        for (int m = 0; m < NM; m++) Fmt::load(rp[rw][m], unit, regs[rw][m][nb]);
                    }
                }
#pragma unroll
            for (int nb = 0; nb < NB; nb++) {
                const int unit = c0 + nb * 32 + lane;
                if (unit < upr) {
                    typename Fmt::Act A;
                    Fmt::load_act(act, unit, A);
#pragma unroll
                    for (int rw = 0; rw < RW; rw++)
                        if (valid[rw]) {
#pragma unroll
                            for (int m = 0; m < NM; m++) acc[rw][m] += Fmt::dot(regs[rw][m][nb], A, unit);
                        }
                }
            }
        }
#pragma unroll
        for (int rw = 0; rw < RW; rw++) {
            const long u = ub + rw;
            if (u >= u1) continue;
            float g = warp_sum(acc[rw][0]);
            float uu = PAIR ? warp_sum(acc[rw][NM - 1]) : 0.f;
            if (lane == 0) {
                if (PAIR) {
                    const float a = valid[rw] ? (p.use_silu ? act_silu(g) : act_relu(g)) * uu : 0.f;
                    p.out_f32[(long)t * total + u] = a;
                } else {
                    if (!valid[rw]) g = 0.f;
                    if (p.bias) g += p.bias[u % p.rows];
                    if (p.out_f32) p.out_f32[(long)t * total + u] = g;
                    if (p.out_hidden) store_hidden(p.out_hidden, (long)t * total + u, p.hidden_type, g);
                }
            }
        }
    }
}

struct ReduceParams {
    const void* w;          // [E][rows][ncols]
    int type;
    int n_experts;
    int rows, ncols;        // rows = output features (H), ncols = reduction length (I)
    int slots;              // k
    const int64_t* ids;     // [T][slots] or null (dense)
    int id_offset;
    const float* weights;   // [T][slots] or null (all 1)
    const float* a;         // [T*slots][ncols] fp32 activations (phase-1 output)
    void* out;              // [T][rows] hidden_type
    int hidden_type;
    int accumulate;         // out = round(out + round(result)) in hidden_type (torch `y += y_` semantics)
    const int* bsz;
};

template <class Fmt, int RW, int NB>
__global__ void __launch_bounds__(kGemvThreads, 2) reduce_kernel(const ReduceParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int t = blockIdx.y;
    if (p.bsz && t >= *p.bsz) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kGemvThreads / 32;
    const int nblk = p.ncols / QK_K;
    const int k = p.slots;
    // per-slot staging: q8 [k][ncols] | dx [k][nblk] | bsums [k][ncols/16] | partial [rows_local][k]
    uint8_t* q8 = smem;
    float* dx = reinterpret_cast<float*>(smem + (size_t)k * p.ncols);
    int16_t* bsums = reinterpret_cast<int16_t*>(smem + (size_t)k * p.ncols + (size_t)k * nblk * 4);
    float* partial = reinterpret_cast<float*>(smem + (size_t)k * p.ncols + (size_t)k * nblk * 4 + (size_t)k * (p.ncols / 16) * 2);

    const int r0 = (int)((long)p.rows * blockIdx.x / gridDim.x), r1 = (int)((long)p.rows * (blockIdx.x + 1) / gridDim.x);
    const int nrows = r1 - r0;

    {
        // skipped experts contribute nothing (common.hpp:255-258): do not even read their activations
        unsigned skip = 0;
        for (int j = 0; j < k; j++) {
            long e = p.ids ? (long)p.ids[(long)t * k + j] - p.id_offset : 0;
            if (e < 0 || e >= p.n_experts) skip |= 1u << j;
        }
        cta_quantize_q8k_rows<8>(p.a, (long)t * k * p.ncols, p.ncols, KTB200_TYPE_F32, k, p.ncols, skip, q8, dx, bsums);
    }
    __syncthreads();

    const int upr = nblk * Fmt::kUnitsPerBlock;
    const int total = nrows * k;
    for (int ub = warp * RW; ub < total; ub += nwarps * RW) {
        typename Fmt::Row rp[RW];
        ActQ8K act[RW];
        bool valid[RW];
        float acc[RW];
#pragma unroll
        for (int rw = 0; rw < RW; rw++) {
            const int u = ub + rw;
            valid[rw] = false;
            acc[rw] = 0.f;
            if (u < total) {
                const int hl = u / k, j = u % k;
                long e = p.ids ? (long)p.ids[(long)t * k + j] - p.id_offset : 0;
                if (e >= 0 && e < p.n_experts) {
                    valid[rw] = true;
                    rp[rw] = Fmt::row(p.w, e * p.rows + r0 + hl, p.ncols, p.type);
                    act[rw] = ActQ8K{q8 + (size_t)j * p.ncols, dx + j * nblk, bsums + j * (p.ncols / 16)};
                }
            }
        }
        for (int c0 = 0; c0 < upr; c0 += 32 * NB) {
            typename Fmt::Regs regs[RW][NB];
#pragma unroll
            for (int rw = 0; rw < RW; rw++)
#pragma unroll
                for (int nb = 0; nb < NB; nb++) {
                    const int unit = c0 + nb * 32 + lane;
                    if (valid[rw] && unit < upr) Fmt::load(rp[rw], unit, regs[rw][nb]);
                }
#pragma unroll
            for (int rw = 0; rw < RW; rw++)
#pragma unroll
                for (int nb = 0; nb < NB; nb++) {
                    const int unit = c0 + nb * 32 + lane;
                    if (valid[rw] && unit < upr) {
                        typename Fmt::Act A;
                        Fmt::load_act(act[rw], unit, A);
                        acc[rw] += Fmt::dot(regs[rw][nb], A, unit);
                    }
                }
        }
#pragma unroll
        for (int rw = 0; rw < RW; rw++) {
            const int u = ub + rw;
            if (u >= total) continue;
            const float s = warp_sum(acc[rw]);
            if (lane == 0) partial[u] = valid[rw] ? s : 0.f;
        }
    }
    __syncthreads();
    // weighted accumulation over the k experts IN expert_ids ORDER (moe.cpp:222-236); the reference's
    // `out += d * w` is one fused multiply-add in every FMA-capable build.
    for (int hl = threadIdx.x; hl < nrows; hl += kGemvThreads) {
        float acc = 0.f;
        for (int j = 0; j < k; j++) {
            long e = p.ids ? (long)p.ids[(long)t * k + j] - p.id_offset : 0;
            if (e < 0 || e >= p.n_experts) continue;
            const float d = partial[hl * k + j];
            acc = p.weights ? __fmaf_rn(d, p.weights[(long)t * k + j], acc) : acc + d;
        }
        const long o = (long)t * p.rows + r0 + hl;
        if (p.accumulate) {
            // `y += y_` on hidden-type tensors (experts.py:1011): both operands are already rounded
            // to hidden_type, the sum is rounded once more.
            acc = load_hidden(p.out, o, p.hidden_type) + round_hidden(acc, p.hidden_type);
        }
        store_hidden(p.out, o, p.hidden_type, acc);
    }
}

}  // namespace ktb
