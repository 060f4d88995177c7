// Absorbed-MLA paged decode attention on the Blackwell tensor path (tcgen05 + TMEM + TMA) and the paged latent KV write.
//
// Replaces MLAWrapper.run / flashinfer BatchMLAPagedAttentionWrapper
// (archive/ktransformers/operators/flashinfer_wrapper.py:117-161, attention.py:419-447) and the Triton split-KV decode
// (archive/ktransformers/operators/triton_attention.py:16-385).  Math (attention.py:395-478):
//     s[h,t] = (q_nope[h,:] . ckv[t,:] + q_pe[h,:] . k_pe[t,:]) * sm_scale        576-long dot, bf16 x bf16 -> fp32
//     p      = softmax_t(s)   fp32, online; P is cast to bf16 before P.V            (triton_attention.py:137-141)
//     out[h] = sum_t p[h,t] * ckv[t,:]                                             512 wide
//
// Work decomposition: CTA = (KV split, group of 64 heads, sequence); 6 warps with fixed roles
//     warp 0      TMA producer: paged KV tiles of 32 tokens x 576 columns, 9 boxes of [32 rows x 64 columns] per tile
//                 (cp.async.bulk.tensor, 128-byte swizzle) into a 4-stage ring, one mbarrier per stage
//     warp 1      tensor-core issuer (one lane): S = Q.K^T and O^T = V^T.P^T as tcgen05.mma with TMEM accumulators
//     warps 2..5  softmax / rescale / epilogue: tcgen05.ld of S, online softmax (thread = head), P -> shared memory
// Both products read the SAME shared-memory tile: for S it is the K-major B operand [32 tokens x 576], for O^T its first
// 512 columns are the MN-major A operand [latent x tokens] (the swizzle is a function of the shared-memory address only).
// The output is accumulated TRANSPOSED — O^T[latent 512][head 64] = 4 blocks of 128 TMEM lanes x 64 columns — because a
// [head][latent] accumulator for 64 heads would need 512 columns in the M=64 tcgen05 layout (half the lanes idle): the
// whole tensor memory.  TMEM map (512 columns allocated): [0,256) O^T, [256,384) / [384,512) the two S buffers, each FOUR
// partial accumulators of 32 columns: a 64 x 32 x 16 MMA is 16 cycles of work behind a pipeline several times as deep, so
// 36 of them chained through ONE accumulator run at the latency, not the throughput (measured: ~75 cycles each); the 36
// k-steps are dealt round-robin to four independent accumulators and the softmax warps add the four partial scores.
//
// Online softmax with a LAZY reference maximum: p = 2^(x - m_ref), m_ref is only raised (and O^T rescaled in TMEM, all
// four warps) when some head's running maximum exceeds it by more than 8 — p stays <= 256, exact in bf16/fp32 terms —
// so the steady-state tile costs no TMEM round trip of the accumulator.  Split-KV partials (fp32 O, base-2 LSE) are
// merged by mla_merge_kernel.
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"

namespace ktb {

using namespace umma;

constexpr int kDK = 576;             // 512 latent + 64 rope
constexpr int kDV = 512;
constexpr int kHG = 64;              // heads per CTA (UMMA M of S)
constexpr int kLT = 32;              // kv tokens per tile (UMMA N of S, K of O^T)
constexpr int kStages = 4;
constexpr int kChunks = kDK / 64;    // 9 column chunks of 64 bf16 = 128 B (one swizzle row)
constexpr int kMlaThreads = 192;
constexpr int kStageBytes = kLT * kDK * 2;         // 36,864 = 9 regions of 32 rows x 128 B
constexpr int kKRegion = kLT * 128;                // 4,096
constexpr int kQRegion = kHG * 128;                // 8,192
constexpr int kQBytes = kHG * kDK * 2;             // 73,728
constexpr int kPBytes = kHG * kLT * 2;             // 4,096: [8 head groups][4 token groups][8 heads][8 tokens]
constexpr int kOffQ = kStages * kStageBytes;       // 147,456
constexpr int kOffP = kOffQ + kQBytes;             // 221,184
constexpr int kOffMisc = kOffP + 2 * kPBytes;      // 229,376 (P is double-buffered)
constexpr int kTmemCols = 512;
constexpr int kColO = 0, kColS = 256;
constexpr int kSChains = 4;          // independent partial-score accumulators per S buffer
constexpr float kRescaleThreshold = 8.f;

struct MlaMisc {
    unsigned long long k_full[kStages], k_empty[kStages], s_full[2], s_empty[2], p_full, p_free[2];
    uint32_t tmem_base;
    int need[2][4];
    float alpha[kHG];     // per head: 2^(m_ref_old - m_ref_new) of the current rescale
    float l[kHG];         // final row sums
    float m[kHG];         // final reference maxima
};
constexpr int kMlaSmem = kOffMisc + (int)sizeof(MlaMisc) + 1024;   // + slack to align the base to 1024 B
static_assert(kMlaSmem <= 232448, "shared memory budget");

struct MlaKParams {
    const __nv_bfloat16* q_nope;   // [B][Hq][512]
    const __nv_bfloat16* q_pe;     // [B][Hq][64]
    const int* page_table;         // [B][max_pages]
    const int* kv_len;             // [B]
    int num_heads, page_size, max_pages, num_splits;
    float scale_log2;              // sm_scale * log2(e)
    float* o_part;                 // [B][splits][Hq][512]
    float* lse_part;               // [B][splits][Hq]  (base-2)
    float* debug;                  // optional: S of the first tile [64][32], P bytes, see ktb200_debug_mla
};

__device__ __forceinline__ float ex2(float x) {   // 2^x, one MUFU (x = -inf -> 0)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}

__global__ void __launch_bounds__(kMlaThreads, 1) mla_decode_tc_kernel(const __grid_constant__ CUtensorMap kv_map, const MlaKParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;           // 128-byte swizzle atoms are 1024-byte aligned
    uint8_t* smem = smem_raw + (base - raw);
    MlaMisc& misc = *reinterpret_cast<MlaMisc*>(smem + kOffMisc);
    const int split = blockIdx.x, hg = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h0 = hg * kHG;
    griddep_launch_dependents();
    griddep_wait();          // q, the newest cache row and kv_len come from the kernels before this one
    int L = p.kv_len[b];
    if (L > p.max_pages * p.page_size) L = p.max_pages * p.page_size;
    const int ntiles = (L + kLT - 1) / kLT;
    const int tiles_per = (ntiles + p.num_splits - 1) / p.num_splits;
    const int tile0 = split * tiles_per, tile1 = min(ntiles, tile0 + tiles_per);
    const int n = tile1 - tile0;
    float* o_out = p.o_part + (((long)b * p.num_splits + split) * p.num_heads + h0) * kDV;
    float* lse_out = p.lse_part + ((long)b * p.num_splits + split) * p.num_heads + h0;

    if (n <= 0) {   // empty split: the neutral element of the merge
        for (int i = tid; i < kHG * kDV; i += kMlaThreads)
            if (h0 + i / kDV < p.num_heads) o_out[i] = 0.f;
        if (tid < kHG && h0 + tid < p.num_heads) lse_out[tid] = -INFINITY;
        return;
    }

    const bool dbg_cta = p.debug && tid == 0 && split == 0 && hg == 0 && b == 0;
    if (dbg_cta) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        reinterpret_cast<unsigned long long*>(p.debug + 2048)[0] = t;
    }
    // ---- one-time setup -----------------------------------------------------------------------------------------
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&kv_map);
        for (int s = 0; s < kStages; s++) { bar_init(smem_u32(&misc.k_full[s]), 1); bar_init(smem_u32(&misc.k_empty[s]), 1); }
        for (int s = 0; s < 2; s++) { bar_init(smem_u32(&misc.s_full[s]), 1); bar_init(smem_u32(&misc.s_empty[s]), 4); }
        bar_init(smem_u32(&misc.p_full), 4);
        bar_init(smem_u32(&misc.p_free[0]), 1);
        bar_init(smem_u32(&misc.p_free[1]), 1);
        bar_fence_init();
        // the first tiles do not depend on anything set up below: request them now
        for (int j = 0; j < n && j < kStages; j++) {
            const int t_base = (tile0 + j) * kLT;
            const int page = p.page_table[(long)b * p.max_pages + t_base / p.page_size];
            const int row = page * p.page_size + t_base % p.page_size;
            const uint32_t bar = smem_u32(&misc.k_full[j]);
            bar_expect_tx(bar, kStageBytes);
#pragma unroll
            for (int c = 0; c < kChunks; c++) tma_load_2d(base + j * kStageBytes + c * kKRegion, &kv_map, bar, c * 64, row);
            if (p.debug && split == 0 && hg == 0 && b == 0) reinterpret_cast<unsigned long long*>(p.debug + 2048)[192 + j] = gtime();
        }
    }
    if (warp == 1) tmem_alloc(smem_u32(&misc.tmem_base), kTmemCols);
    // Q (64 heads x 576) -> shared memory in the K-major 128-byte-swizzle layout: chunk region c (64 columns) holds 64
    // rows of 128 B; the 16-byte piece j of row r sits at piece (j ^ (r & 7)).  Rows beyond num_heads are zero.
    {
        constexpr int kPieces = kHG * (kDK / 8), kIter = kPieces / kMlaThreads;   // 4608 = 24 x 192
        static_assert(kPieces % kMlaThreads == 0, "Q pieces");
        constexpr int kB = 12;
        static_assert(kIter % kB == 0, "Q batches");
#pragma unroll
        for (int i0 = 0; i0 < kIter; i0 += kB) {
            uint4 v[kB];
#pragma unroll
            for (int u = 0; u < kB; u++) {   // 12 independent 16-byte loads in flight per thread
                const int i = tid + (i0 + u) * kMlaThreads, r = i / (kDK / 8), j = i - r * (kDK / 8);
                v[u] = make_uint4(0, 0, 0, 0);
                if (h0 + r < p.num_heads) {
                    const long hrow = (long)b * p.num_heads + h0 + r;
                    v[u] = j < 64 ? __ldg(reinterpret_cast<const uint4*>(p.q_nope + hrow * kDV) + j) : __ldg(reinterpret_cast<const uint4*>(p.q_pe + hrow * 64) + (j - 64));
                }
            }
#pragma unroll
            for (int u = 0; u < kB; u++) {
                const int i = tid + (i0 + u) * kMlaThreads, r = i / (kDK / 8), j = i - r * (kDK / 8);
                const int c = j >> 3, jj = j & 7;
                *reinterpret_cast<uint4*>(smem + kOffQ + c * kQRegion + r * 128 + ((jj ^ (r & 7)) << 4)) = v[u];
            }
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = misc.tmem_base;
    if (dbg_cta) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        reinterpret_cast<unsigned long long*>(p.debug + 2048)[1] = t;
    }

    if (warp == 0) {
        // ================================================================ TMA producer
        if (lane == 0) {
            for (int j = kStages; j < n; j++) {
                const int s = j % kStages;
                bar_wait(smem_u32(&misc.k_empty[s]), ((j / kStages) & 1) ^ 1);
                const int t_base = (tile0 + j) * kLT;
                const int page = p.page_table[(long)b * p.max_pages + t_base / p.page_size];
                const int row = page * p.page_size + t_base % p.page_size;
                const uint32_t bar = smem_u32(&misc.k_full[s]);
                bar_expect_tx(bar, kStageBytes);
#pragma unroll
                for (int c = 0; c < kChunks; c++) tma_load_2d(base + s * kStageBytes + c * kKRegion, &kv_map, bar, c * 64, row);
                if (p.debug && split == 0 && hg == 0 && b == 0 && j < 60) reinterpret_cast<unsigned long long*>(p.debug + 2048)[192 + j] = gtime();
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================================ tensor-core issuer (converged warp, see mma_f16)
        {
            constexpr uint32_t idesc_qk = instr_desc(1, 1, 1, 0, 0, kHG, kLT);    // f32 += bf16 . bf16, A K-major, B K-major, 64 x 32
            constexpr uint32_t idesc_pv = instr_desc(1, 1, 1, 1, 0, 128, kHG);    // A MN-major (V^T from the [token][latent] tile), 128 x 64
            auto issue_qk = [&](int j) {
                const int s = j % kStages, buf = j & 1;
                bar_wait(smem_u32(&misc.s_empty[buf]), ((j >> 1) & 1) ^ 1);
                bar_wait(smem_u32(&misc.k_full[s]), (j / kStages) & 1);
                tc_fence_after();
                const uint32_t kb = base + s * kStageBytes, qb = base + kOffQ;
#pragma unroll
                for (int c = 0; c < kChunks; c++)
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        mma_f16(tmem + kColS + (buf * kSChains + (k & (kSChains - 1))) * kLT, smem_desc(qb + c * kQRegion + k * 32, 16, 1024, kLayoutSw128),
                                smem_desc(kb + c * kKRegion + k * 32, 16, 1024, kLayoutSw128), idesc_qk, c != 0);
                mma_commit(smem_u32(&misc.s_full[buf]));
            };
            issue_qk(0);
            for (int j = 0; j < n; j++) {
                if (j + 1 < n) issue_qk(j + 1);
                bar_wait(smem_u32(&misc.p_full), j & 1);
                tc_fence_after();
                const int s = j % kStages;
                const uint32_t kb = base + s * kStageBytes, pb = base + kOffP + (j & 1) * kPBytes;
#pragma unroll
                for (int m = 0; m < 4; m++)
#pragma unroll
                    for (int k = 0; k < 2; k++)
                        mma_f16(tmem + kColO + m * kHG, smem_desc(kb + 2 * m * kKRegion + k * 2048, kKRegion, 1024, kLayoutSw128),
                                smem_desc(pb + k * 256, 128, 512, kLayoutNone), idesc_pv, (j | k) != 0);
                mma_commit(smem_u32(&misc.k_empty[s]));
                mma_commit(smem_u32(&misc.p_free[j & 1]));
                if (p.debug && lane == 0 && split == 0 && hg == 0 && b == 0 && j < 60) reinterpret_cast<unsigned long long*>(p.debug + 2048)[128 + j] = gtime();
            }
        }
        __syncwarp();
    } else {
        // ================================================================ softmax / rescale / epilogue (4 warps)
        // S rows of a sub-partition live in its TMEM lanes 0..15 (M = 64 layout): lane l < 16 loads the 32 scores of head
        // 16 sp + l and hands tokens 16..31 to lane l + 16, so that all 32 lanes work: lane = (head, half of the tile).
        const int sp = warp & 3;                       // TMEM sub-partition of this warp
        const uint32_t lane_base = (uint32_t)(32 * sp) << 16;
        const int upper = lane >> 4;
        const int head = 16 * sp + (lane & 15);
        const int st = tid - 64;                       // 0..127 among the softmax threads
        float m_ref = -INFINITY, m_run = -INFINITY, l = 0.f;   // l: this lane's half of the row sum
        for (int j = 0; j < n; j++) {
            const int buf = j & 1;
            bar_wait(smem_u32(&misc.s_full[buf]), (j >> 1) & 1);
            tc_fence_after();
            if (p.debug && st == 0 && split == 0 && hg == 0 && b == 0 && j < 60) reinterpret_cast<unsigned long long*>(p.debug + 2048)[64 + j] = gtime();
            uint32_t sv[32];
            {   // S = sum of the four partial accumulators
                uint32_t t1[32], t2[32];
                const uint32_t sa = tmem + lane_base + kColS + buf * kSChains * kLT;
                tmem_ld32(sa, sv);
                tmem_ld32(sa + kLT, t1);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; i++) sv[i] = __float_as_uint(__uint_as_float(sv[i]) + __uint_as_float(t1[i]));
                tmem_ld32(sa + 2 * kLT, t1);
                tmem_ld32(sa + 3 * kLT, t2);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; i++) sv[i] = __float_as_uint(__uint_as_float(sv[i]) + (__uint_as_float(t1[i]) + __uint_as_float(t2[i])));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) bar_arrive(smem_u32(&misc.s_empty[buf]));
            if (p.debug && j == 0 && !upper && split == 0 && hg == 0 && b == 0)
                for (int i = 0; i < 32; i++) p.debug[head * 32 + i] = __uint_as_float(sv[i]);
            const int t_base = (tile0 + j) * kLT + 16 * upper;
            float x[16];
            float mt = -INFINITY;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const uint32_t hi = __shfl_sync(0xffffffffu, sv[16 + i], lane & 15);
                const float v = __uint_as_float(upper ? hi : sv[i]);
                x[i] = (t_base + i < L) ? v * p.scale_log2 : -INFINITY;
                mt = fmaxf(mt, x[i]);
            }
            mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 16));
            m_run = fmaxf(m_run, mt);
            // does any head need its reference raised?  (never on the first tile: O^T is overwritten by the first PV)
            bool need = false;
            if (j == 0) m_ref = m_run;
            else need = m_run - m_ref > kRescaleThreshold;
            const unsigned any_w = __ballot_sync(0xffffffffu, need);
            if (lane == 0) misc.need[buf][warp - 2] = any_w != 0;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const bool any = misc.need[buf][0] | misc.need[buf][1] | misc.need[buf][2] | misc.need[buf][3];
            if (any) {
                // O^T must be quiescent: the previous PV has to be complete (it also frees the other P buffer)
                bar_wait(smem_u32(&misc.p_free[(j - 1) & 1]), ((j - 1) >> 1) & 1);
                tc_fence_after();
                const float a = need ? ex2(m_ref - m_run) : 1.f;
                if (!upper) misc.alpha[head] = a;
                if (need) { l *= a; m_ref = m_run; }
                asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
                for (int q = 0; q < 8; q++) {   // 4 latent blocks x 2 halves of the 64 head columns
                    uint32_t ov[32];
                    const uint32_t ta = tmem + lane_base + kColO + q * 32;
                    tmem_ld32(ta, ov);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; i++) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * misc.alpha[(q & 1) * 32 + i]);
                    tmem_st32(ta, ov);
                }
                tmem_wait_st();
                tc_fence_before();
                asm volatile("bar.sync 1, 128;" ::: "memory");   // alpha[] / need[] may be rewritten only after everyone used them
            }
            // this tile's P buffer was last read by PV(j - 2)
            if (j >= 2) bar_wait(smem_u32(&misc.p_free[buf]), ((j >> 1) & 1) ^ 1);
            {
                float ps = 0.f;
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float p0 = ex2(x[i] - m_ref), p1 = ex2(x[i + 1] - m_ref);
                    ps += p0 + p1;
                    pk[i >> 1] = pack_bf16(p0, p1);
                }
                l += ps;
                // P[head][token] as the K-major no-swizzle B operand: core matrix (8 heads x 8 tokens) = 128 contiguous bytes
                uint8_t* prow = smem + kOffP + buf * kPBytes + (head >> 3) * 512 + (head & 7) * 16 + upper * 256;
                *reinterpret_cast<uint4*>(prow) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(prow + 128) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
            // rows of the tile beyond kv_len hold whatever the page contains: P is 0 there, but 0 * NaN is NaN -> zero the V rows
            if ((tile0 + j) * kLT + kLT > L) {
                const int valid = L - (tile0 + j) * kLT;
                uint8_t* kb = smem + (j % kStages) * kStageBytes;
                for (int i = st; i < (kLT - valid) * 64; i += 128) {      // 64 pieces of 16 B per row over the 8 latent chunks
                    const int r = valid + i / 64, pc = i % 64;
                    *reinterpret_cast<uint4*>(kb + (pc >> 3) * kKRegion + r * 128 + (pc & 7) * 16) = make_uint4(0, 0, 0, 0);
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bar_arrive(smem_u32(&misc.p_full));
            if (p.debug && st == 0 && split == 0 && hg == 0 && b == 0 && j < 60) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                reinterpret_cast<unsigned long long*>(p.debug + 2048)[4 + j] = t;
            }
        }
        // ---- epilogue: O^T / l -> fp32 partial output, base-2 LSE ---------------------------------------------------
        l += __shfl_xor_sync(0xffffffffu, l, 16);
        if (!upper) { misc.l[head] = l; misc.m[head] = m_ref; }
        bar_wait(smem_u32(&misc.p_free[(n - 1) & 1]), ((n - 1) >> 1) & 1);   // the last PV (and with it every earlier one) is complete
        tc_fence_after();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (p.debug && st == 0 && split == 0 && hg == 0 && b == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            reinterpret_cast<unsigned long long*>(p.debug + 2048)[2] = t;
        }
        if (st < kHG && h0 + st < p.num_heads) lse_out[st] = misc.m[st] + log2f(misc.l[st]);
        if (st < kHG) misc.alpha[st] = 1.f / misc.l[st];
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            float inv[32];
#pragma unroll
            for (int i = 0; i < 32; i++) inv[i] = misc.alpha[half * 32 + i];
            const int nh = min(32, p.num_heads - h0 - half * 32);   // valid heads of this half
#pragma unroll 1
            for (int blk = 0; blk < 4; blk++) {
                uint32_t ov[32];
                tmem_ld32(tmem + lane_base + kColO + blk * kHG + half * 32, ov);
                tmem_wait_ld();
                float* dst = o_out + (long)(half * 32) * kDV + blk * 128 + 32 * sp + lane;
#pragma unroll
                for (int i = 0; i < 32; i++)
                    if (i < nh) dst[(long)i * kDV] = __uint_as_float(ov[i]) * inv[i];
            }
        }
        tc_fence_before();
        if (p.debug && st == 0 && split == 0 && hg == 0 && b == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            reinterpret_cast<unsigned long long*>(p.debug + 2048)[3] = t;
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, kTmemCols);
    }
}

// out[b][h][:] = sum_s w_s * o_part[b][s][h][:],  w_s = 2^(lse_s - max) / sum ; lse (natural log) optional
__global__ void __launch_bounds__(128) mla_merge_kernel(const float* o_part, const float* lse_part, int num_splits, int num_heads,
                                                        __nv_bfloat16* out, float* lse_out) {
    __shared__ float ws[128];
    __shared__ float red[8];
    griddep_launch_dependents();
    griddep_wait();
    const int bh = blockIdx.x, b = bh / num_heads, h = bh % num_heads;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float my = tid < num_splits ? lse_part[((long)b * num_splits + tid) * num_heads + h] : -INFINITY;   // num_splits <= 128
    float mx = my;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    const int c = tid * 4;
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(out + ((long)b * num_heads + h) * kDV + c);
    if (mx == -INFINITY) {   // kv_len == 0 (e.g. a padded CUDA-graph batch slot): zeros, lse = -inf
        o2[0] = __floats2bfloat162_rn(0.f, 0.f);
        o2[1] = __floats2bfloat162_rn(0.f, 0.f);
        if (lse_out && tid == 0) lse_out[(long)b * num_heads + h] = -INFINITY;
        return;
    }
    const float e = tid < num_splits ? exp2f(my - mx) : 0.f;
    float den = warp_sum(e);
    if (lane == 0) red[4 + warp] = den;
    __syncthreads();
    den = (red[4] + red[5]) + (red[6] + red[7]);
    ws[tid] = e / den;
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    const float* src = o_part + ((long)b * num_splits * num_heads + h) * kDV + c;
    const long sstride = (long)num_heads * kDV;
    for (int s0 = 0; s0 < num_splits; s0 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (s0 + u < num_splits) v[u] = __ldcs(reinterpret_cast<const float4*>(src + (s0 + u) * sstride));
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (s0 + u < num_splits) {
                const float w = ws[s0 + u];
                acc.x += w * v[u].x; acc.y += w * v[u].y; acc.z += w * v[u].z; acc.w += w * v[u].w;
            }
    }
    o2[0] = __floats2bfloat162_rn(acc.x, acc.y);
    o2[1] = __floats2bfloat162_rn(acc.z, acc.w);
    if (lse_out && tid == 0) lse_out[(long)b * num_heads + h] = (mx + log2f(den)) * 0.6931471805599453f;
}

// StaticCache.update (archive/ktransformers/models/custom_cache.py:147-200): one CTA per token
__global__ void __launch_bounds__(72) mla_kv_write_kernel(__nv_bfloat16* kv, int page_size, const __nv_bfloat16* ckv,
                                                          const __nv_bfloat16* k_pe, const int* page_idx, const int* page_off) {
    const int t = blockIdx.x, i = threadIdx.x;   // 72 x 16 B = 1152 B
    __nv_bfloat16* dst = kv + ((long)page_idx[t] * page_size + page_off[t]) * kDK;
    const uint4 v = (i < 64) ? reinterpret_cast<const uint4*>(ckv + (long)t * kDV)[i] : reinterpret_cast<const uint4*>(k_pe + (long)t * 64)[i - 64];
    reinterpret_cast<uint4*>(dst)[i] = v;
}

static int pick_splits(int batch, int num_heads, int max_kv_tiles, int device) {
    const int groups = batch * ((num_heads + kHG - 1) / kHG);
    int s = (num_sms(device) + groups - 1) / groups;          // one CTA per SM (227 KB of shared memory each)
    const int cap = (max_kv_tiles + 3) / 4;                   // at least 4 tiles (128 tokens) per split
    if (s > cap) s = cap;
    if (s > 128) s = 128;
    if (s < 1) s = 1;
    return s;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libktb200.so links libcudart only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}

static float* g_mla_debug = nullptr;

}  // namespace ktb

extern "C" {

size_t ktb200_mla_workspace_bytes(int batch, int num_heads, int max_splits) {
    if (max_splits <= 0) max_splits = 128;
    return (size_t)batch * max_splits * num_heads * (ktb::kDV + 1) * sizeof(float);
}

int ktb200_mla_decode(const ktb200_mla_params* q, void* stream) {
    using namespace ktb;
    if (!q || !q->q_nope || !q->q_pe || !q->kv_cache || !q->page_table || !q->kv_len || !q->out || !q->workspace) { set_error("mla_decode: null pointer"); return KTB200_EINVAL; }
    if (q->batch <= 0) return KTB200_OK;
    if (q->num_heads <= 0 || q->page_size <= 0 || q->page_size % kLT || q->max_pages_per_seq <= 0) {
        set_error("mla_decode: page_size %d must be a positive multiple of %d", q->page_size, kLT);
        return KTB200_EINVAL;
    }
    if (((uintptr_t)q->kv_cache & 15) || ((uintptr_t)q->q_nope & 15) || ((uintptr_t)q->q_pe & 15)) { set_error("mla_decode: q / kv_cache must be 16-byte aligned"); return KTB200_EINVAL; }
    int dev = 0;
    KTB_CUDA_CHECK(cudaGetDevice(&dev));
    const int max_tiles = q->max_pages_per_seq * (q->page_size / kLT);
    int splits = q->num_kv_splits > 0 ? q->num_kv_splits : pick_splits(q->batch, q->num_heads, max_tiles, dev);
    if (splits > max_tiles) splits = max_tiles;
    const size_t need = (size_t)q->batch * splits * q->num_heads * (kDV + 1) * sizeof(float);
    if (need > q->workspace_bytes) { set_error("mla_decode: workspace too small (%zu < %zu bytes for %d splits)", q->workspace_bytes, need, splits); return KTB200_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;

    // the paged cache as a 2-D tensor [token rows][576 columns]; a tile never crosses a page (page_size % 32 == 0)
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { set_error("mla_decode: cuTensorMapEncodeTiled is not available from this driver"); return KTB200_ECUDA; }
    CUtensorMap map;
    const cuuint64_t rows = q->kv_cache_rows > 0 ? (cuuint64_t)q->kv_cache_rows : (cuuint64_t)1 << 31;
    const cuuint64_t gdim[2] = {(cuuint64_t)kDK, rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)kDK * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)kLT};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(q->kv_cache), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("mla_decode: cuTensorMapEncodeTiled failed (%d)", (int)cr); return KTB200_ECUDA; }

    MlaKParams p{};
    p.q_nope = (const __nv_bfloat16*)q->q_nope; p.q_pe = (const __nv_bfloat16*)q->q_pe;
    p.page_table = q->page_table; p.kv_len = q->kv_len; p.num_heads = q->num_heads; p.page_size = q->page_size;
    p.max_pages = q->max_pages_per_seq; p.num_splits = splits; p.scale_log2 = q->sm_scale * 1.4426950408889634f;
    p.o_part = (float*)q->workspace;
    p.lse_part = p.o_part + (size_t)q->batch * splits * q->num_heads * kDV;
    p.debug = g_mla_debug;
    static bool attr_set[64] = {};
    if (!attr_set[dev & 63]) {
        KTB_CUDA_CHECK(cudaFuncSetAttribute(mla_decode_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMlaSmem));
        attr_set[dev & 63] = true;
    }
    const int head_groups = (q->num_heads + kHG - 1) / kHG;
    KTB_CUDA_CHECK(launch_pdl(mla_decode_tc_kernel, dim3(splits, head_groups, q->batch), dim3(kMlaThreads), (size_t)kMlaSmem, s, map, p));
    count_launch();
    KTB_CUDA_CHECK(launch_pdl(mla_merge_kernel, dim3(q->batch * q->num_heads), dim3(128), 0, s, (const float*)p.o_part, (const float*)p.lse_part, splits, q->num_heads,
                              (__nv_bfloat16*)q->out, q->lse_out));
    count_launch();
    return KTB200_OK;
}

// Diagnostics: while set, CTA (split 0, head group 0, sequence 0) of ktb200_mla_decode writes the raw fp32 scores of its
// first tile (S[64 heads][32 tokens], before scaling) to debug_dev[0..2048) and %globaltimer stamps (uint64) to
// debug_dev + 2048: [0] kernel start, [1] setup done, [2] last PV complete, [3] epilogue stored, [4 + j] tile j's P ready
// [64 + j] tile j's S seen
// by the softmax warps, [128 + j] PV(j) issued, [192 + j] tile j's TMA issued (debug_dev must hold >= 2048 floats + 256 uint64).
void ktb200_debug_mla(float* debug_dev) { ktb::g_mla_debug = debug_dev; }

int ktb200_mla_kv_write(void* kv_cache, int page_size, const void* ckv, const void* k_pe, const int* page_idx,
                        const int* page_offset, int n_tokens, void* stream) {
    using namespace ktb;
    if (!kv_cache || !ckv || !k_pe || !page_idx || !page_offset) { set_error("mla_kv_write: null pointer"); return KTB200_EINVAL; }
    if (n_tokens <= 0) return KTB200_OK;
    mla_kv_write_kernel<<<n_tokens, 72, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)kv_cache, page_size, (const __nv_bfloat16*)ckv,
                                                                   (const __nv_bfloat16*)k_pe, page_idx, page_offset);
    KTB_LAUNCH_CHECK();
    return KTB200_OK;
}

}  // extern "C"
