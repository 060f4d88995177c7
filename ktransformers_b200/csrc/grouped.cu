// Grouped expert GEMM for prefill-sized batches on the Blackwell tensor path — MOE::forward_many
// (archive/csrc/ktransformers_ext/operators/llamafile/moe.cpp:248-365 ≡ kt-kernel/operators/llamafile/moe.hpp:461-746):
//     count tokens per expert -> per-token Q8_K quantisation + scatter into per-expert contiguous order ->
//     per-expert GEMM (gate, up) -> silu * mul -> requantise -> per-expert GEMM (down) -> per-token weighted gather.
// The decode kernels stream every (token, expert) pair's weights; here an expert's weights are read ONCE per 32-token tile.
//
// Arithmetic.  The reference's dot is an exact integer: sum_j sc_j * sum_{32} q * x8 per super-block, scaled in fp32.  The
// tensor cores get operands that hold those integers EXACTLY in fp16: A = sc_j * q (<= 63 * 15 for Q4_K; sc * (q - 32) for
// Q6_K), B = the Q8_K activation bytes; tcgen05.mma kind::f16 accumulates their products in fp32 (every product and every
// partial sum of a 256-long block is an integer < 2^25: exact up to one final rounding), one accumulator per SUPER-BLOCK;
// the epilogue applies (d_w * d_x) * isum - (dmin_w * d_x) * msum in fp32 exactly like the decode kernels.  msum (Q4_K mins x
// activation block sums) is a second K = 16 MMA on [m_j] x [bsum16].  Q6_K: |sc * (q - 32)| can reach 4096 and fp16 holds
// integers exactly to 2048 (even ones to 4096): the rare odd product above 2048 moves one element by 2^-12 relative.
//
// Kernel (grouped_gemm_kernel): persistent CTAs, tile = (expert, 128 weight rows, 32 tokens), 13 warps:
//     warps 0-7   producers: weights (global, 16-byte loads) -> fp16 A tile in the K-major 128-byte-swizzle layout; activations
//                 (int8 SoA rows gathered through the sorted pair list) -> fp16 B tile; mins / block sums -> the small A2 / B2 tiles
//     warp  8     tcgen05 issuer: 16 MMAs (128 x 32 x 16) + 1 per super-block into TMEM, two stages
//     warps 9-12  epilogue: tcgen05.ld of the two accumulators, fp32 scale-and-add into registers, store at the end of the tile
#include <cuda_fp16.h>

#include "act_quant.cuh"
#include "common.cuh"
#include "handles.cuh"
#include "umma.cuh"

namespace ktb {

using namespace umma;

constexpr int kGM = 128, kGN = 32, kGThreads = 13 * 32, kGProd = 256;
constexpr int kGA = 4 * kGM * 128;        // 65,536: 4 swizzle atoms of 128 rows x 128 B per super-block
constexpr int kGB = 4 * kGN * 128;        // 16,384
constexpr int kGA2 = kGM * 32, kGB2 = kGN * 32;
constexpr int kOffB = 2 * kGA, kOffA2 = kOffB + 2 * kGB, kOffB2 = kOffA2 + 2 * kGA2, kOffMiscG = kOffB2 + 2 * kGB2;

struct GrpMisc {
    unsigned long long ab_full[2], d_full[2], stage_free[2];
    uint32_t tmem_base;
    float dxs[2][kGN];
    float2 rowsc[2][kGM];
};
constexpr int kGSmem = kOffMiscG + (int)sizeof(GrpMisc) + 1024;

struct GrpGemmParams {
    const uint8_t* w;          // expert weights
    long expert_bytes;         // bytes per expert
    int fmt;                   // 0: raw Q4_K rows, 1: Q6_K 4-row tiles (repack_q6k4t)
    int R, Kc;                 // rows per expert, reduction length
    const int8_t* xq;          // activations: int8 [rows][Kc]
    const float* xd;           // [rows][Kc / 256]
    const int16_t* xbs;        // [rows][Kc / 16]
    const int* rowmap;         // sorted position -> activation row (null: identity)
    const int* offsets;        // [E + 1] first sorted position of every expert
    const int* nt_prefix;      // [E + 1] 32-token tiles before every expert
    int E;
    float* out;                // [P][R] fp32
};

// byte b of a register array (b is a compile-time constant after unrolling: no local-memory byte addressing)
__device__ __forceinline__ int ub(const uint32_t* a, int b) { return (int)((a[b >> 2] >> ((b & 3) * 8)) & 0xffu); }
__device__ __forceinline__ int sb8(const uint32_t* a, int b) { return (int)(int8_t)((a[b >> 2] >> ((b & 3) * 8)) & 0xffu); }
__device__ __forceinline__ uint32_t h2(int a, int b) {
    const __half2 v = __halves2half2(__int2half_rn(a), __int2half_rn(b));
    return *reinterpret_cast<const uint32_t*>(&v);
}

__global__ void __launch_bounds__(kGThreads, 1) grouped_gemm_kernel(const GrpGemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    GrpMisc& misc = *reinterpret_cast<GrpMisc*>(smem + kOffMiscG);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = p.Kc / QK_K, MT = p.R / kGM;
    if (tid == 0) {
        for (int s = 0; s < 2; s++) { bar_init(smem_u32(&misc.ab_full[s]), 8); bar_init(smem_u32(&misc.d_full[s]), 1); bar_init(smem_u32(&misc.stage_free[s]), 4); }
        bar_fence_init();
    }
    if (warp == 8) tmem_alloc(smem_u32(&misc.tmem_base), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = misc.tmem_base;
    const int total_tiles = p.nt_prefix[p.E] * MT;
    unsigned it = 0;   // super-block iterations done by this CTA: stage = it & 1, use = it >> 1

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        // tile -> (expert, row tile, token tile); token tile fastest: concurrently running CTAs share the weight tile through L2
        int lo = 0, hi = p.E;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (p.nt_prefix[mid] * MT <= tile) lo = mid; else hi = mid;
        }
        const int e = lo;
        const int local = tile - p.nt_prefix[e] * MT, ntile_e = p.nt_prefix[e + 1] - p.nt_prefix[e];
        const int mt = local / ntile_e, nt = local - mt * ntile_e;
        const int m0 = mt * kGM;
        const int p0 = p.offsets[e] + nt * kGN;
        const int n_valid = min(kGN, p.offsets[e + 1] - p0);
        const uint8_t* we = p.w + (long)e * p.expert_bytes;

        if (warp < 8) {
            // ====================================================================== producers
            const int pt = tid;
            for (int sb = 0; sb < nblk; sb++, it++) {
                const int stage = it & 1;
                bar_wait(smem_u32(&misc.stage_free[stage]), ((it >> 1) & 1) ^ 1);
                uint8_t* As = smem + stage * kGA;
                uint8_t* Bs = smem + kOffB + stage * kGB;
                const int r = pt >> 1;
                if (p.fmt == 0) {
                    // Q4_K: thread = (row, two 64-element chunks); chunk c is exactly one 128-byte swizzle row of atom c
                    const uint8_t* blk = we + ((long)(m0 + r) * nblk + sb) * SZ_Q4_K;
                    const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(blk));
                    const uint32_t hw[4] = {hdr.x, hdr.y, hdr.z, hdr.w};   // bytes 4..15: the 12 packed 6-bit scales / mins
                    int sc[8], mn[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if (j < 4) { sc[j] = ub(hw, 4 + j) & 63; mn[j] = ub(hw, 8 + j) & 63; }
                        else { sc[j] = (ub(hw, 8 + j) & 0xF) | ((ub(hw, j) >> 6) << 4); mn[j] = (ub(hw, 8 + j) >> 4) | ((ub(hw, 4 + j) >> 6) << 4); }
                    }
#pragma unroll
                    for (int cc = 0; cc < 2; cc++) {
                        const int c = (pt & 1) * 2 + cc;
                        const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(blk + 16 + c * 32)), q1 = __ldg(reinterpret_cast<const uint4*>(blk + 32 + c * 32));
                        const uint32_t qw[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                        uint8_t* arow = As + c * (kGM * 128) + r * 128;
#pragma unroll
                        for (int i = 0; i < 8; i++) {   // piece i: elements 8i..8i+7 of the chunk; i < 4 low nibbles (sub-block 2c), else high (2c+1)
                            const int s_ = sc[2 * c + (i >> 2)];
                            const uint32_t w0 = qw[2 * (i & 3)], w1 = qw[2 * (i & 3) + 1];
                            const int sh = (i >> 2) * 4;
                            uint4 v;
                            v.x = h2(s_ * (int)((w0 >> sh) & 15), s_ * (int)((w0 >> (8 + sh)) & 15));
                            v.y = h2(s_ * (int)((w0 >> (16 + sh)) & 15), s_ * (int)((w0 >> (24 + sh)) & 15));
                            v.z = h2(s_ * (int)((w1 >> sh) & 15), s_ * (int)((w1 >> (8 + sh)) & 15));
                            v.w = h2(s_ * (int)((w1 >> (16 + sh)) & 15), s_ * (int)((w1 >> (24 + sh)) & 15));
                            *reinterpret_cast<uint4*>(arow + ((i ^ (r & 7)) << 4)) = v;
                        }
                    }
                    if ((pt & 1) == 0) {
                        const __half2 dm = *reinterpret_cast<const __half2*>(&hdr.x);
                        misc.rowsc[stage][r] = make_float2(__low2float(dm), __high2float(dm));
                    } else {   // A2 row: [m_0 m_0 m_1 m_1 ... m_7 m_7] against the sixteen 16-value activation sums
                        uint8_t* a2 = smem + kOffA2 + stage * kGA2 + (r >> 3) * 256 + (r & 7) * 16;
                        *reinterpret_cast<uint4*>(a2) = make_uint4(h2(mn[0], mn[0]), h2(mn[1], mn[1]), h2(mn[2], mn[2]), h2(mn[3], mn[3]));
                        *reinterpret_cast<uint4*>(a2 + 128) = make_uint4(h2(mn[4], mn[4]), h2(mn[5], mn[5]), h2(mn[6], mn[6]), h2(mn[7], mn[7]));
                    }
                } else {
                    // Q6_K in 4-row tiles: thread = (row, 128-element half)
                    const int hh = pt & 1;
                    const int row = m0 + r, rw = row & 3, nrb = 4 * nblk, f = rw * nblk + sb;
                    const uint8_t* item = we + (long)(row >> 2) * nrb * SZ_Q6_K;
                    uint32_t ql[16], qh[8];
#pragma unroll
                    for (int c = 0; c < 4; c++) *reinterpret_cast<uint4*>(ql + 4 * c) = __ldg(reinterpret_cast<const uint4*>(item + (long)(4 * hh + c) * nrb * 16 + f * 16));
#pragma unroll
                    for (int c = 0; c < 2; c++) *reinterpret_cast<uint4*>(qh + 4 * c) = __ldg(reinterpret_cast<const uint4*>(item + (long)nrb * 128 + (long)(2 * hh + c) * nrb * 16 + f * 16));
                    const uint4 scv = __ldg(reinterpret_cast<const uint4*>(item + (long)nrb * 192 + f * 16));
                    const uint32_t scw[2] = {hh ? scv.z : scv.x, hh ? scv.w : scv.y};
#pragma unroll
                    for (int gq = 0; gq < 4; gq++) {
#pragma unroll
                        for (int l0 = 0; l0 < 32; l0 += 8) {
                            const int s_ = sb8(scw, (l0 >> 4) + 2 * gq);
                            int v[8];
#pragma unroll
                            for (int x = 0; x < 8; x++) {
                                const int l = l0 + x;
                                const int lowq = ub(ql, (gq & 1) * 32 + l);
                                const int nib = (gq >> 1) ? (lowq >> 4) : (lowq & 15);
                                const int hi2 = (ub(qh, l) >> (2 * gq)) & 3;
                                v[x] = s_ * ((nib | (hi2 << 4)) - 32);
                            }
                            const int atom = 2 * hh + (gq >> 1), pi = (gq & 1) * 4 + (l0 >> 3);
                            *reinterpret_cast<uint4*>(As + atom * (kGM * 128) + r * 128 + ((pi ^ (r & 7)) << 4)) =
                                make_uint4(h2(v[0], v[1]), h2(v[2], v[3]), h2(v[4], v[5]), h2(v[6], v[7]));
                        }
                    }
                    if (hh == 0) {
                        const float d = __half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short*>(item + (long)nrb * 208 + f * 2))));
                        misc.rowsc[stage][r] = make_float2(d, 0.f);
                    }
                }
                // activations: 32 rows x 256 int8 -> fp16, 16 values per unit
#pragma unroll
                for (int uu = 0; uu < 2; uu++) {
                    const int u = pt + uu * kGProd, n = u >> 4, pc = u & 15;
                    uint4 lo4 = make_uint4(0, 0, 0, 0), hi4 = make_uint4(0, 0, 0, 0);
                    if (n < n_valid) {
                        const int row = p.rowmap ? p.rowmap[p0 + n] : p0 + n;
                        const uint4 xv = *reinterpret_cast<const uint4*>(p.xq + (long)row * p.Kc + sb * QK_K + pc * 16);
                        const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
                        lo4 = make_uint4(h2(sb8(xw, 0), sb8(xw, 1)), h2(sb8(xw, 2), sb8(xw, 3)), h2(sb8(xw, 4), sb8(xw, 5)), h2(sb8(xw, 6), sb8(xw, 7)));
                        hi4 = make_uint4(h2(sb8(xw, 8), sb8(xw, 9)), h2(sb8(xw, 10), sb8(xw, 11)), h2(sb8(xw, 12), sb8(xw, 13)), h2(sb8(xw, 14), sb8(xw, 15)));
                    }
                    const int kp = 2 * pc, atom = kp >> 3, pi = kp & 7;
                    uint8_t* brow = Bs + atom * (kGN * 128) + n * 128;
                    *reinterpret_cast<uint4*>(brow + ((pi ^ (n & 7)) << 4)) = lo4;
                    *reinterpret_cast<uint4*>(brow + (((pi + 1) ^ (n & 7)) << 4)) = hi4;
                }
                if (pt < 64) {   // B2: the sixteen 16-value sums of the block; dxs: the block scale
                    const int n = pt >> 1, kg = pt & 1;
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (n < n_valid) {
                        const int row = p.rowmap ? p.rowmap[p0 + n] : p0 + n;
                        const uint4 bv = *reinterpret_cast<const uint4*>(p.xbs + (long)row * (p.Kc / 16) + sb * 16 + kg * 8);
                        const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#define KTB_S16(w, hi) ((int)(short)((hi) ? ((w) >> 16) : ((w) & 0xffffu)))
                        v = make_uint4(h2(KTB_S16(bw[0], 0), KTB_S16(bw[0], 1)), h2(KTB_S16(bw[1], 0), KTB_S16(bw[1], 1)), h2(KTB_S16(bw[2], 0), KTB_S16(bw[2], 1)),
                                       h2(KTB_S16(bw[3], 0), KTB_S16(bw[3], 1)));
#undef KTB_S16
                        if (kg == 0) misc.dxs[stage][n] = p.xd[(long)row * nblk + sb];
                    } else if (kg == 0) misc.dxs[stage][n] = 0.f;
                    *reinterpret_cast<uint4*>(smem + kOffB2 + stage * kGB2 + (n >> 3) * 256 + kg * 128 + (n & 7) * 16) = v;
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) bar_arrive(smem_u32(&misc.ab_full[stage]));
            }
        } else if (warp == 8) {
            // ====================================================================== tensor-core issuer (converged warp)
            constexpr uint32_t idesc = instr_desc(1, 0, 0, 0, 0, kGM, kGN);   // f32 += f16 . f16, both K-major, 128 x 32
            for (int sb = 0; sb < nblk; sb++, it++) {
                const int stage = it & 1;
                bar_wait(smem_u32(&misc.ab_full[stage]), (it >> 1) & 1);
                tc_fence_after();
                const uint32_t a = base + stage * kGA, b = base + kOffB + stage * kGB;
                const uint32_t d1 = tmem + stage * 64, d2 = d1 + 32;
#pragma unroll
                for (int ks = 0; ks < 16; ks++)
                    mma_f16(d1, smem_desc(a + (ks >> 2) * (kGM * 128) + (ks & 3) * 32, 16, 1024, kLayoutSw128),
                            smem_desc(b + (ks >> 2) * (kGN * 128) + (ks & 3) * 32, 16, 1024, kLayoutSw128), idesc, ks != 0);
                if (p.fmt == 0)
                    mma_f16(d2, smem_desc(base + kOffA2 + stage * kGA2, 128, 256, kLayoutNone), smem_desc(base + kOffB2 + stage * kGB2, 128, 256, kLayoutNone), idesc, 0);
                mma_commit(smem_u32(&misc.d_full[stage]));
            }
        } else {
            // ====================================================================== epilogue (4 warps = 128 rows)
            const int sp = warp & 3, row = 32 * sp + lane;
            const uint32_t lane_base = (uint32_t)(32 * sp) << 16;
            float acc[kGN];
#pragma unroll
            for (int n = 0; n < kGN; n++) acc[n] = 0.f;
            for (int sb = 0; sb < nblk; sb++, it++) {
                const int stage = it & 1;
                bar_wait(smem_u32(&misc.d_full[stage]), (it >> 1) & 1);
                tc_fence_after();
                uint32_t d1[32], d2[32];
                tmem_ld32(tmem + lane_base + stage * 64, d1);
                if (p.fmt == 0) tmem_ld32(tmem + lane_base + stage * 64 + 32, d2);
                tmem_wait_ld();
                const float2 rs = misc.rowsc[stage][row];
                if (p.fmt == 0) {
#pragma unroll
                    for (int n = 0; n < kGN; n++) {
                        const float dx = misc.dxs[stage][n];
                        acc[n] += (rs.x * dx) * __uint_as_float(d1[n]) - (rs.y * dx) * __uint_as_float(d2[n]);
                    }
                } else {
#pragma unroll
                    for (int n = 0; n < kGN; n++) acc[n] += (rs.x * misc.dxs[stage][n]) * __uint_as_float(d1[n]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) bar_arrive(smem_u32(&misc.stage_free[stage]));
            }
#pragma unroll
            for (int n = 0; n < kGN; n++)
                if (n < n_valid) p.out[(long)(p0 + n) * p.R + m0 + row] = acc[n];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem, 128);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// bookkeeping kernels (moe.cpp:250-290: m_local_num_, m_local_pos_, prefix offsets)
__global__ void grp_count_kernel(const int64_t* ids, int npairs, int k, int id_offset, int n_local, const int* bsz, int t0, int* counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    if (bsz && t0 + i / k >= *bsz) return;
    const long e = (long)ids[i] - id_offset;
    if (e >= 0 && e < n_local) atomicAdd(counts + e, 1);
}
__global__ void grp_scan_kernel(const int* counts, int E, int* offsets, int* nt_prefix, int* cursor) {
    if (threadIdx.x == 0) {
        int o = 0, t = 0;
        for (int e = 0; e < E; e++) {
            offsets[e] = o; nt_prefix[e] = t; cursor[e] = o;
            o += counts[e];
            t += (counts[e] + kGN - 1) / kGN;
        }
        offsets[E] = o; nt_prefix[E] = t;
    }
}
__global__ void grp_scatter_kernel(const int64_t* ids, int npairs, int k, int id_offset, int n_local, const int* bsz, int t0, int* cursor, int* tokmap, int* pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const long e = (long)ids[i] - id_offset;
    if ((bsz && t0 + i / k >= *bsz) || e < 0 || e >= n_local) { pos[i] = -1; return; }
    const int s = atomicAdd(cursor + e, 1);   // order inside an expert is irrelevant: every pair's result is independent
    tokmap[s] = i / k;
    pos[i] = s;
}
// rows of `src` (hidden type, [n][ncols]) -> int8 SoA + block scales + 16-value sums; one warp per 256-block
__global__ void __launch_bounds__(256) grp_quant_x_kernel(const void* src, int hidden_type, int nrows, int ncols, int8_t* q, float* d, int16_t* bs) {
    const int lane = threadIdx.x & 31, gw = blockIdx.x * 8 + (threadIdx.x >> 5), nblk = ncols / QK_K;
    if (gw >= nrows * nblk) return;
    const int r = gw / nblk, b = gw - r * nblk;
    float x[8];
    load_block8(src, (long)r * ncols + (long)b * QK_K + lane * 8, hidden_type, x);
    warp_quantize_q8k_block(x, lane, reinterpret_cast<uint32_t*>(q + (long)r * ncols + (long)b * QK_K), d + (long)r * nblk + b, bs + (long)r * (ncols / 16) + b * 16);
}
// a = act(g) * u (fp32, sorted pair rows) -> Q8_K SoA      (moe.cpp:300-318: silu * mul, then from_float to vec_dot_type)
__global__ void __launch_bounds__(256) grp_act_quant_kernel(const float* g, const float* u, const int* offsets, int E, int ncols, int use_silu, int8_t* q, float* d, int16_t* bs) {
    const int lane = threadIdx.x & 31, gw = blockIdx.x * 8 + (threadIdx.x >> 5), nblk = ncols / QK_K;
    if (gw >= offsets[E] * nblk) return;
    const int r = gw / nblk, b = gw - r * nblk;
    const long o = (long)r * ncols + (long)b * QK_K + lane * 8;
    float x[8];
    const float4 g0 = *reinterpret_cast<const float4*>(g + o), g1 = *reinterpret_cast<const float4*>(g + o + 4);
    const float4 u0 = *reinterpret_cast<const float4*>(u + o), u1 = *reinterpret_cast<const float4*>(u + o + 4);
    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, uv[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = (use_silu ? act_silu(gv[i]) : act_relu(gv[i])) * uv[i];
    warp_quantize_q8k_block(x, lane, reinterpret_cast<uint32_t*>(q + (long)r * ncols + (long)b * QK_K), d + (long)r * nblk + b, bs + (long)r * (ncols / 16) + b * 16);
}
// out[t] = sum_j w[t][j] * down[pos[t][j]] in expert_ids order, one FMA per expert (moe.cpp:340-358), rounded like from_float
__global__ void __launch_bounds__(256) grp_combine_kernel(const float* dd, const int* pos, const float* weights, int T, int k, int H, const int* bsz, int t0, void* out,
                                                          int hidden_type) {
    const int t = blockIdx.y;
    if (bsz && t0 + t >= *bsz) return;
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    float acc = 0.f;
    for (int j = 0; j < k; j++) {
        const int s = pos[t * k + j];
        if (s >= 0) acc = __fmaf_rn(dd[(long)s * H + h], weights[t * k + j], acc);
    }
    store_hidden(out, (long)t * H + h, hidden_type, acc);
}

struct GrpScratch {
    size_t cap_pairs = 0, cap_x = 0, cap_a = 0, cap_d = 0;   // pairs, tokens * H, pairs * I, pairs * H
    int *counts = nullptr, *offsets = nullptr, *nt_prefix = nullptr, *cursor = nullptr, *tokmap = nullptr, *pos = nullptr;
    int8_t *xq = nullptr, *aq = nullptr;
    float *xd = nullptr, *ad = nullptr, *g = nullptr, *u = nullptr, *dd = nullptr;
    int16_t *xbs = nullptr, *abs16 = nullptr;
};
static GrpScratch g_grp[64];   // one arena per device, shared by every handle (calls on one device are stream-ordered by the caller)

static int grp_ensure(int dev, int tokens, int k, int H, int I) {
    GrpScratch& s = g_grp[dev & 63];
    size_t P = (size_t)tokens * k, nx = (size_t)tokens * H, na = P * I, nd = P * H;
    if (s.cap_pairs >= P && s.cap_x >= nx && s.cap_a >= na && s.cap_d >= nd) return KTB200_OK;
    P = P > s.cap_pairs ? P : s.cap_pairs; nx = nx > s.cap_x ? nx : s.cap_x; na = na > s.cap_a ? na : s.cap_a; nd = nd > s.cap_d ? nd : s.cap_d;   // grow only
    KTB_CUDA_CHECK(cudaDeviceSynchronize());   // earlier calls may still be using the arena
    cudaFree(s.counts); cudaFree(s.tokmap); cudaFree(s.pos); cudaFree(s.xq); cudaFree(s.xd); cudaFree(s.xbs); cudaFree(s.aq); cudaFree(s.ad); cudaFree(s.abs16);
    cudaFree(s.g); cudaFree(s.u); cudaFree(s.dd);
    s = GrpScratch();
    const size_t cp = P;
    KTB_CUDA_CHECK(cudaMalloc(&s.counts, (size_t)(4 * 1024 + 8) * sizeof(int)));
    s.offsets = s.counts + 1024; s.nt_prefix = s.counts + 2048 + 1; s.cursor = s.counts + 3072 + 2;
    KTB_CUDA_CHECK(cudaMalloc(&s.tokmap, cp * sizeof(int)));
    KTB_CUDA_CHECK(cudaMalloc(&s.pos, cp * sizeof(int)));
    KTB_CUDA_CHECK(cudaMalloc(&s.xq, nx));
    KTB_CUDA_CHECK(cudaMalloc(&s.xd, nx / 256 * sizeof(float)));
    KTB_CUDA_CHECK(cudaMalloc(&s.xbs, nx / 16 * sizeof(int16_t)));
    KTB_CUDA_CHECK(cudaMalloc(&s.aq, na));
    KTB_CUDA_CHECK(cudaMalloc(&s.ad, na / 256 * sizeof(float)));
    KTB_CUDA_CHECK(cudaMalloc(&s.abs16, na / 16 * sizeof(int16_t)));
    KTB_CUDA_CHECK(cudaMalloc(&s.g, na * sizeof(float)));
    KTB_CUDA_CHECK(cudaMalloc(&s.u, na * sizeof(float)));
    KTB_CUDA_CHECK(cudaMalloc(&s.dd, nd * sizeof(float)));
    s.cap_pairs = cp; s.cap_x = nx; s.cap_a = na; s.cap_d = nd;
    return KTB200_OK;
}

// true when ktb200_moe_forward may take the grouped tensor-core path for this handle
bool grouped_ok(const ktb200_moe* m, int k) {
    const ktb200_moe_config& c = m->cfg;
    const FmtId fd = pick_fmt(c.down_type, m->down_layout);
    return c.gate_type == KTB200_TYPE_Q4_K && c.up_type == KTB200_TYPE_Q4_K && (fd == FMT_Q6K4T || fd == FMT_Q4K) && c.hidden_size % 256 == 0 &&
           c.intermediate_size % 256 == 0 && c.hidden_size % kGM == 0 && c.intermediate_size % kGM == 0 && c.expert_num <= 1023 && k <= 32;
}

int moe_forward_grouped(ktb200_moe* m, int qlen, int k, const int64_t* ids, const float* weights, const void* input, void* output, const int* bsz, cudaStream_t s) {
    const ktb200_moe_config& c = m->cfg;
    const int E = c.expert_num, H = c.hidden_size, I = c.intermediate_size, dev = m->device;
    static const int chunk_cap = [] { const char* e = getenv("KTB200_GROUPED_CHUNK"); return e ? atoi(e) : 1024; }();
    const int Tc = qlen < chunk_cap ? qlen : chunk_cap;
    int rc = grp_ensure(dev, Tc, k, H, I);   // grow-only scratch: not capturable on first use (like ktb200_moe_gate_forward)
    if (rc) return rc;
    GrpScratch& g = g_grp[dev & 63];
    static bool attr[64] = {};
    if (!attr[dev & 63]) {
        KTB_CUDA_CHECK(cudaFuncSetAttribute(grouped_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSmem));
        attr[dev & 63] = true;
    }
    const FmtId fd = pick_fmt(c.down_type, m->down_layout);
    const size_t hb = type_size(c.hidden_type);
    const int grid = num_sms(dev);
    for (int t0 = 0; t0 < qlen; t0 += Tc) {
        const int T = qlen - t0 < Tc ? qlen - t0 : Tc, P = T * k;
        const int64_t* ids_c = ids + (size_t)t0 * k;
        const float* w_c = weights + (size_t)t0 * k;
        const uint8_t* x_c = reinterpret_cast<const uint8_t*>(input) + (size_t)t0 * H * hb;
        uint8_t* o_c = reinterpret_cast<uint8_t*>(output) + (size_t)t0 * H * hb;
        KTB_CUDA_CHECK(cudaMemsetAsync(g.counts, 0, (size_t)E * sizeof(int), s));
        grp_count_kernel<<<(P + 255) / 256, 256, 0, s>>>(ids_c, P, k, c.expert_id_offset, E, bsz, t0, g.counts);
        grp_scan_kernel<<<1, 32, 0, s>>>(g.counts, E, g.offsets, g.nt_prefix, g.cursor);
        grp_scatter_kernel<<<(P + 255) / 256, 256, 0, s>>>(ids_c, P, k, c.expert_id_offset, E, bsz, t0, g.cursor, g.tokmap, g.pos);
        grp_quant_x_kernel<<<(T * (H / 256) + 7) / 8, 256, 0, s>>>(x_c, c.hidden_type, T, H, g.xq, g.xd, g.xbs);
        GrpGemmParams gp{};
        gp.fmt = 0; gp.R = I; gp.Kc = H; gp.xq = g.xq; gp.xd = g.xd; gp.xbs = g.xbs; gp.rowmap = g.tokmap; gp.offsets = g.offsets; gp.nt_prefix = g.nt_prefix; gp.E = E;
        gp.expert_bytes = (long)I * (H / 256) * SZ_Q4_K;
        gp.w = reinterpret_cast<const uint8_t*>(c.gate_proj); gp.out = g.g;
        grouped_gemm_kernel<<<grid, kGThreads, kGSmem, s>>>(gp);
        gp.w = reinterpret_cast<const uint8_t*>(c.up_proj); gp.out = g.u;
        grouped_gemm_kernel<<<grid, kGThreads, kGSmem, s>>>(gp);
        grp_act_quant_kernel<<<(P * (I / 256) + 7) / 8, 256, 0, s>>>(g.g, g.u, g.offsets, E, I, c.use_silu, g.aq, g.ad, g.abs16);
        GrpGemmParams gd{};
        gd.fmt = fd == FMT_Q6K4T ? 1 : 0; gd.R = H; gd.Kc = I; gd.xq = g.aq; gd.xd = g.ad; gd.xbs = g.abs16; gd.rowmap = nullptr; gd.offsets = g.offsets; gd.nt_prefix = g.nt_prefix;
        gd.E = E; gd.expert_bytes = (long)H * (I / 256) * (fd == FMT_Q6K4T ? SZ_Q6_K : SZ_Q4_K);
        gd.w = reinterpret_cast<const uint8_t*>(c.down_proj); gd.out = g.dd;
        grouped_gemm_kernel<<<grid, kGThreads, kGSmem, s>>>(gd);
        grp_combine_kernel<<<dim3((H + 255) / 256, T), 256, 0, s>>>(g.dd, g.pos, w_c, T, k, H, bsz, t0, o_c, c.hidden_type);
        KTB_LAUNCH_CHECK();
        count_launch(8);
    }
    return KTB200_OK;
}

}  // namespace ktb
