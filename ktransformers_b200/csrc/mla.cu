// Absorbed-MLA paged decode attention + paged latent KV write (sm_100a).
//
// Replaces MLAWrapper.run / flashinfer BatchMLAPagedAttentionWrapper fa2
// (archive/ktransformers/operators/flashinfer_wrapper.py:117-161, third_party/custom_flashinfer/include/
// flashinfer/attention/mla.cuh:775-...) and the Triton split-KV decode (archive/ktransformers/operators/
// triton_attention.py:16-385).  Math (attention.py:395-478, Appendix A of SURVEY.md):
//     s[h,t] = (q_nope[h,:] . ckv[t,:] + q_pe[h,:] . k_pe[t,:]) * sm_scale          (576-long dot, bf16 MMA, fp32 acc)
//     p      = softmax_t(s)  (fp32, online), P cast to bf16 before P.V            (triton_attention.py:137-141)
//     out[h] = sum_t p[h,t] * ckv[t,:]                                             (512 wide)
//
// Round-1 kernel: flash-decoding with warp-level bf16 tensor-core MMA (mma.sync m16n8k16, fp32 accumulate),
// cp.async 3-stage paged KV pipeline in shared memory, split-KV across CTAs and an LSE merge kernel.
// CTA = 32 heads x one KV split; 8 warps: QK^T is tiled 2 (head tiles of 16) x 4 (8-token tiles) over the
// warps, P.V is tiled over the 512 latent dims (64 per warp).  K and V share ONE shared-memory tile (V is
// the first 512 columns of the same latent row) so every KV byte is fetched from HBM once per head tile and
// re-used for both products.  The tcgen05/TMEM version of the two products is the next step.
#include <cuda_bf16.h>

#include "common.cuh"

namespace ktb {

constexpr int kMlaThreads = 256;
constexpr int kHT = 32;              // heads per CTA
constexpr int kKT = 32;              // kv tokens per tile
constexpr int kDK = 576;             // 512 latent + 64 rope
constexpr int kDV = 512;
constexpr int kRowPad = kDK + 8;     // bf16 elements per smem row (1168 B: conflict-free ldmatrix)
constexpr int kPPad = kKT + 8;
constexpr int kStages = 3;

struct MlaSmem {
    __nv_bfloat16 q[kHT][kRowPad];
    __nv_bfloat16 k[kStages][kKT][kRowPad];
    __nv_bfloat16 p[kHT][kPPad];
    float max_part[kHT][4];
    float sum_part[kHT][4];
    float row_max[kHT];
    float row_sum[kHT];
    float alpha[kHT];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;   // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

struct MlaKParams {
    const __nv_bfloat16* q_nope;   // [B][Hq][512]
    const __nv_bfloat16* q_pe;     // [B][Hq][64]
    const __nv_bfloat16* kv;       // [pages][page_size][576]
    const int* page_table;         // [B][max_pages]
    const int* kv_len;             // [B]
    int num_heads, page_size, max_pages, num_splits;
    float scale_log2;              // sm_scale * log2(e)
    float* o_part;                 // [B][splits][Hq][512]
    float* lse_part;               // [B][splits][Hq]  (base-2)
};

__global__ void __launch_bounds__(kMlaThreads, 1) mla_decode_split_kernel(const MlaKParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    MlaSmem& sm = *reinterpret_cast<MlaSmem*>(smem_raw);
    const int split = blockIdx.x, ht = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h0 = ht * kHT;
    const int L = p.kv_len[b];
    // this split's token range, in whole tiles
    const int ntiles = (L + kKT - 1) / kKT;
    const int tiles_per = (ntiles + p.num_splits - 1) / p.num_splits;
    const int tile0 = split * tiles_per, tile1 = min(ntiles, tile0 + tiles_per);
    float* o_out = p.o_part + (((long)b * p.num_splits + split) * p.num_heads + h0) * kDV;
    float* lse_out = p.lse_part + ((long)b * p.num_splits + split) * p.num_heads + h0;

    if (tile0 >= tile1) {   // empty split: neutral element for the merge
        for (int i = tid; i < kHT * kDV; i += kMlaThreads)
            if (h0 + i / kDV < p.num_heads) o_out[i] = 0.f;
        if (tid < kHT && h0 + tid < p.num_heads) lse_out[tid] = -INFINITY;
        return;
    }

    // ---- Q tile -> smem (rows beyond num_heads are zero) ----------------------------------------------
    for (int i = tid; i < kHT * (kDK / 8); i += kMlaThreads) {
        const int r = i / (kDK / 8), c = (i % (kDK / 8)) * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (h0 + r < p.num_heads) {
            const long hrow = (long)b * p.num_heads + h0 + r;
            v = (c < kDV) ? *reinterpret_cast<const uint4*>(p.q_nope + hrow * kDV + c)
                          : *reinterpret_cast<const uint4*>(p.q_pe + hrow * 64 + (c - kDV));
        }
        *reinterpret_cast<uint4*>(&sm.q[r][c]) = v;
    }
    if (tid < kHT) { sm.row_max[tid] = -INFINITY; sm.row_sum[tid] = 0.f; }

    auto load_tile = [&](int tile, int stage) {
        const int t_base = tile * kKT;
        const int page = p.page_table[(long)b * p.max_pages + t_base / p.page_size];
        const __nv_bfloat16* src = p.kv + ((long)page * p.page_size + (t_base % p.page_size)) * kDK;
        for (int i = tid; i < kKT * (kDK / 8); i += kMlaThreads) {
            const int r = i / (kDK / 8), c = (i % (kDK / 8)) * 8;
            cp_async16(smem_u32(&sm.k[stage][r][c]), src + (long)r * kDK + c, t_base + r < L);
        }
    };

    // prologue: fill kStages-1 stages
#pragma unroll
    for (int s = 0; s < kStages - 1; s++) {
        if (tile0 + s < tile1) load_tile(tile0 + s, s);
        cp_async_commit();
    }

    const int mt = warp >> 2, nt = warp & 3;   // QK^T: head tile (16 rows) x token tile (8 tokens)
    float o[2][8][4];                          // P.V : 2 head tiles x 8 n-tiles of this warp's 64 latent dims
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int n = 0; n < 8; n++)
#pragma unroll
            for (int c = 0; c < 4; c++) o[a][n][c] = 0.f;

    for (int tile = tile0; tile < tile1; tile++) {
        const int stage = (tile - tile0) % kStages;
        cp_async_wait<kStages - 2>();
        __syncthreads();                       // tile `tile` landed; everyone is done with the stage refilled below
        if (tile + kStages - 1 < tile1) load_tile(tile + kStages - 1, (tile - tile0 + kStages - 1) % kStages);
        cp_async_commit();

        // ---- S = Q K^T for this warp's 16 x 8 tile ----------------------------------------------------
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        {
            const uint32_t a_base = smem_u32(&sm.q[16 * mt + (lane & 15)][(lane >> 4) * 8]);
            // B via ldmatrix.x4: matrices (k0, k0+8, k0+16, k0+24) of the 8 token rows
            const uint32_t b_base = smem_u32(&sm.k[stage][8 * nt + (lane & 7)][(lane >> 3) * 8]);
#pragma unroll 4
            for (int k0 = 0; k0 < kDK; k0 += 32) {
                uint32_t a0, a1, a2, a3, a4, a5, a6, a7, b0, b1, b2, b3;
                ldsm_x4(a_base + k0 * 2, a0, a1, a2, a3);
                ldsm_x4(a_base + (k0 + 16) * 2, a4, a5, a6, a7);
                ldsm_x4(b_base + k0 * 2, b0, b1, b2, b3);
                mma_bf16(s, a0, a1, a2, a3, b0, b1);
                mma_bf16(s, a4, a5, a6, a7, b2, b3);
            }
        }
        // ---- online softmax -----------------------------------------------------------------------------
        const int r_lo = 16 * mt + (lane >> 2), r_hi = r_lo + 8;
        const int col = 8 * nt + 2 * (lane & 3);
        const int t_base = tile * kKT;
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int t = t_base + col + (c & 1);
            v[c] = (t < L) ? s[c] * p.scale_log2 : -INFINITY;
        }
        float m_lo = fmaxf(v[0], v[1]), m_hi = fmaxf(v[2], v[3]);
        m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
        m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
        m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
        m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
        if ((lane & 3) == 0) { sm.max_part[r_lo][nt] = m_lo; sm.max_part[r_hi][nt] = m_hi; }
        __syncthreads();
        const float old_lo = sm.row_max[r_lo], old_hi = sm.row_max[r_hi];
        float new_lo = old_lo, new_hi = old_hi;
#pragma unroll
        for (int j = 0; j < 4; j++) { new_lo = fmaxf(new_lo, sm.max_part[r_lo][j]); new_hi = fmaxf(new_hi, sm.max_part[r_hi][j]); }
        // at least one valid token exists in every tile of the range, so new_* is finite
        float pr[4];
        pr[0] = exp2f(v[0] - new_lo); pr[1] = exp2f(v[1] - new_lo);
        pr[2] = exp2f(v[2] - new_hi); pr[3] = exp2f(v[3] - new_hi);
        *reinterpret_cast<__nv_bfloat162*>(&sm.p[r_lo][col]) = __floats2bfloat162_rn(pr[0], pr[1]);
        *reinterpret_cast<__nv_bfloat162*>(&sm.p[r_hi][col]) = __floats2bfloat162_rn(pr[2], pr[3]);
        float l_lo = pr[0] + pr[1], l_hi = pr[2] + pr[3];
        l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
        l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
        l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
        l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
        if ((lane & 3) == 0) { sm.sum_part[r_lo][nt] = l_lo; sm.sum_part[r_hi][nt] = l_hi; }
        if (nt == 0 && (lane & 3) == 0) {
            sm.alpha[r_lo] = exp2f(old_lo - new_lo);   // exp2(-inf) = 0 on the first tile
            sm.alpha[r_hi] = exp2f(old_hi - new_hi);
        }
        __syncthreads();
        if (tid < kHT) {   // fold the tile into the running row statistics
            const float a = sm.alpha[tid];
            float mx = sm.row_max[tid];
#pragma unroll
            for (int j = 0; j < 4; j++) mx = fmaxf(mx, sm.max_part[tid][j]);
            sm.row_sum[tid] = sm.row_sum[tid] * a + (sm.sum_part[tid][0] + sm.sum_part[tid][1] + sm.sum_part[tid][2] + sm.sum_part[tid][3]);
            sm.row_max[tid] = mx;
        }
        // ---- O = O * alpha + P V  (this warp: latent dims [64*warp, 64*warp+64)) -----------------------
#pragma unroll
        for (int a = 0; a < 2; a++) {
            const float al_lo = sm.alpha[16 * a + (lane >> 2)], al_hi = sm.alpha[16 * a + 8 + (lane >> 2)];
#pragma unroll
            for (int n = 0; n < 8; n++) { o[a][n][0] *= al_lo; o[a][n][1] *= al_lo; o[a][n][2] *= al_hi; o[a][n][3] *= al_hi; }
        }
#pragma unroll
        for (int k0 = 0; k0 < kKT; k0 += 16) {
            uint32_t pa[2][4];
#pragma unroll
            for (int a = 0; a < 2; a++)
                ldsm_x4(smem_u32(&sm.p[16 * a + (lane & 15)][k0 + (lane >> 4) * 8]), pa[a][0], pa[a][1], pa[a][2], pa[a][3]);
#pragma unroll
            for (int n2 = 0; n2 < 4; n2++) {   // two 8-wide n-tiles per ldmatrix.x4.trans
                uint32_t b0, b1, b2, b3;
                // matrices: (k0..+7, n), (k0+8.., n), (k0..+7, n+8), (k0+8.., n+8)
                const int krow = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int ncol = 64 * warp + 16 * n2 + (lane >> 4) * 8;
                ldsm_x4_t(smem_u32(&sm.k[stage][krow][ncol]), b0, b1, b2, b3);
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    mma_bf16(o[a][2 * n2], pa[a][0], pa[a][1], pa[a][2], pa[a][3], b0, b1);
                    mma_bf16(o[a][2 * n2 + 1], pa[a][0], pa[a][1], pa[a][2], pa[a][3], b2, b3);
                }
            }
        }
        // the __syncthreads at the top of the next iteration protects p / max_part / alpha / k[stage]
    }
    cp_async_wait<0>();
    __syncthreads();
    // ---- write normalised partial output + base-2 LSE ---------------------------------------------------
#pragma unroll
    for (int a = 0; a < 2; a++) {
        const int r_lo = 16 * a + (lane >> 2), r_hi = r_lo + 8;
        const float inv_lo = 1.f / sm.row_sum[r_lo], inv_hi = 1.f / sm.row_sum[r_hi];
#pragma unroll
        for (int n = 0; n < 8; n++) {
            const int c = 64 * warp + 8 * n + 2 * (lane & 3);
            if (h0 + r_lo < p.num_heads) *reinterpret_cast<float2*>(o_out + (long)r_lo * kDV + c) = make_float2(o[a][n][0] * inv_lo, o[a][n][1] * inv_lo);
            if (h0 + r_hi < p.num_heads) *reinterpret_cast<float2*>(o_out + (long)r_hi * kDV + c) = make_float2(o[a][n][2] * inv_hi, o[a][n][3] * inv_hi);
        }
    }
    if (tid < kHT && h0 + tid < p.num_heads) lse_out[tid] = sm.row_max[tid] + log2f(sm.row_sum[tid]);
}

// out[b][h][:] = sum_s w_s * o_part[b][s][h][:],  w_s = 2^(lse_s - max) / sum ; lse (natural log) optional
__global__ void __launch_bounds__(128) mla_merge_kernel(const float* o_part, const float* lse_part, int num_splits, int num_heads,
                                                        __nv_bfloat16* out, float* lse_out) {
    const int bh = blockIdx.x, b = bh / num_heads, h = bh % num_heads;
    float mx = -INFINITY;
    for (int s = 0; s < num_splits; s++) mx = fmaxf(mx, lse_part[((long)b * num_splits + s) * num_heads + h]);
    float den = 0.f;
    for (int s = 0; s < num_splits; s++) den += exp2f(lse_part[((long)b * num_splits + s) * num_heads + h] - mx);
    const int c = threadIdx.x * 4;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int s = 0; s < num_splits; s++) {
        const float w = exp2f(lse_part[((long)b * num_splits + s) * num_heads + h] - mx) / den;
        if (w != 0.f) {
            const float4 v = *reinterpret_cast<const float4*>(o_part + (((long)b * num_splits + s) * num_heads + h) * kDV + c);
            acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
        }
    }
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(out + ((long)b * num_heads + h) * kDV + c);
    o2[0] = __floats2bfloat162_rn(acc.x, acc.y);
    o2[1] = __floats2bfloat162_rn(acc.z, acc.w);
    if (lse_out && threadIdx.x == 0) lse_out[(long)b * num_heads + h] = (mx + log2f(den)) * 0.6931471805599453f;
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
    const int head_tiles = (num_heads + kHT - 1) / kHT;
    int s = (2 * num_sms(device) + batch * head_tiles - 1) / (batch * head_tiles);
    if (s > max_kv_tiles) s = max_kv_tiles;
    if (s > 128) s = 128;
    if (s < 1) s = 1;
    return s;
}

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
    if (q->num_heads <= 0 || q->page_size <= 0 || q->page_size % kKT || q->max_pages_per_seq <= 0) {
        set_error("mla_decode: page_size %d must be a positive multiple of %d", q->page_size, kKT);
        return KTB200_EINVAL;
    }
    int dev = 0;
    KTB_CUDA_CHECK(cudaGetDevice(&dev));
    const int max_tiles = q->max_pages_per_seq * (q->page_size / kKT);
    int splits = q->num_kv_splits > 0 ? q->num_kv_splits : pick_splits(q->batch, q->num_heads, max_tiles, dev);
    if (splits > max_tiles) splits = max_tiles;
    const size_t need = (size_t)q->batch * splits * q->num_heads * (kDV + 1) * sizeof(float);
    if (need > q->workspace_bytes) { set_error("mla_decode: workspace too small (%zu < %zu bytes for %d splits)", q->workspace_bytes, need, splits); return KTB200_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    MlaKParams p{};
    p.q_nope = (const __nv_bfloat16*)q->q_nope; p.q_pe = (const __nv_bfloat16*)q->q_pe; p.kv = (const __nv_bfloat16*)q->kv_cache;
    p.page_table = q->page_table; p.kv_len = q->kv_len; p.num_heads = q->num_heads; p.page_size = q->page_size;
    p.max_pages = q->max_pages_per_seq; p.num_splits = splits; p.scale_log2 = q->sm_scale * 1.4426950408889634f;
    p.o_part = (float*)q->workspace;
    p.lse_part = p.o_part + (size_t)q->batch * splits * q->num_heads * kDV;
    const size_t smem = sizeof(MlaSmem);
    KTB_CUDA_CHECK(cudaFuncSetAttribute(mla_decode_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int head_tiles = (q->num_heads + kHT - 1) / kHT;
    mla_decode_split_kernel<<<dim3(splits, head_tiles, q->batch), kMlaThreads, smem, s>>>(p);
    KTB_LAUNCH_CHECK();
    mla_merge_kernel<<<q->batch * q->num_heads, 128, 0, s>>>(p.o_part, p.lse_part, splits, q->num_heads, (__nv_bfloat16*)q->out, q->lse_out);
    KTB_LAUNCH_CHECK();
    return KTB200_OK;
}

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
