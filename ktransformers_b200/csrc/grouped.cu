// Grouped expert GEMM for prefill-sized batches on the Blackwell tensor path — MOE::forward_many
// (archive/csrc/ktransformers_ext/operators/llamafile/moe.cpp:248-365 ≡ kt-kernel/operators/llamafile/moe.hpp:461-746):
//     count tokens per expert -> per-token Q8_K quantisation + scatter into per-expert contiguous order ->
//     per-expert GEMM (gate, up) -> silu * mul -> requantise -> per-expert GEMM (down) -> per-token weighted gather.
// The decode kernels stream every (token, expert) pair's weights; here an expert's weights are read ONCE per 32-token tile.
//
// Arithmetic: the reference's dot is an exact integer per super-block — sum_j sc_j * (sum over sub-block j of q * x8) — scaled in
// fp32.  The integer tensor path (tcgen05.mma kind::i8, s32 accumulators in TMEM) computes the INNER sums exactly: one MMA of
// K = 32 per Q4_K sub-block (A = the raw 4-bit quants as u8, B = the Q8_K activation bytes), its own accumulator per sub-block;
// Q6_K sub-blocks are 16 long, so one K = 32 MMA covers two of them against a B operand of 64 rows — the 32 tokens with the odd
// sub-block zeroed, then the 32 tokens with the even one zeroed.  The epilogue multiplies every accumulator by its 6/8-bit
// sub-block scale in int32, converts ONCE per super-block and applies (d_w * d_x) * isum - (dmin_w * d_x) * msum in fp32, the
// decode kernels' formula; msum (Q4_K mins x activation block sums) is one K = 16 fp16 MMA of exact small integers.
// Nothing is rounded that the reference does not round.
//
// grouped_gemm_kernel: persistent CTAs, tile = (expert, 128 weight rows, 32 tokens), stage = half a super-block (128 of K):
//     warps 0-3    producers (one weight row per thread): weights (global, 16-byte loads, next stage prefetched in registers) -> int8 A tile in the K-major
//                  128-byte-swizzle layout (Q4_K: nibble split; Q6_K: 4 + 2 bit merge, -32); activation rows gathered through the
//                  sorted pair list -> B tile; row headers (scales, d, dmin) and token scales for the epilogue
//     warp  4      tcgen05 issuer: 4 integer MMAs per stage (+ the mins MMA on the second half), 4 shared-memory stages,
//                  2 TMEM buffers of 256 columns
//     warps 5-12   epilogue: tcgen05.ld, int32 scale-and-add, fp32 finish per super-block, stores at the end of the tile
#include <cuda_fp16.h>

#include "act_quant.cuh"
#include "common.cuh"
#include "handles.cuh"
#include "umma.cuh"

namespace ktb {

using namespace umma;

constexpr int kGM = 128, kGN = 32, kGStages = 3, kGRaw = 6;
constexpr int kGProdWarps = 4, kGEpiWarps = 8, kGThreads = (kGProdWarps + 1 + kGEpiWarps) * 32;   // 13 warps: at most 4 per scheduler, 128 registers each
constexpr int kGA = kGM * 128;            // 16,384: 128 rows x 128 int8 of K, one swizzle atom column
constexpr int kGB = 2 * kGN * 128;        //  8,192: 32 rows (Q4_K) or 64 rows (Q6_K even / odd variants)
constexpr int kGA2 = kGM * 32, kGB2 = kGN * 32;
constexpr int kRawPitch = 144, kRawSlot = kGM * kRawPitch;   // 9 x 16 bytes per producer thread and stage (odd pitch: conflict-free LDS.128)
constexpr int kOffB = kGStages * kGA, kOffA2 = kOffB + kGStages * kGB, kOffB2 = kOffA2 + kGStages * kGA2, kOffRaw = kOffB2 + kGStages * kGB2,
              kOffMiscG = kOffRaw + kGRaw * kRawSlot;

struct GrpMisc {
    unsigned long long ab_full[kGStages], smem_free[kGStages], tmem_full[2], tmem_free[2];
    uint32_t tmem_base, pad[3];
    float dxs[kGStages][kGN];
    uint4 hdr[kGStages][kGM];   // Q4_K: the block header (d, dmin, 12 scale bytes); Q6_K: 8 scales of the half, d as f32
};
constexpr int kGSmem = kOffMiscG + (int)sizeof(GrpMisc) + 1024;
static_assert(kGSmem <= 227 * 1024, "shared memory budget");

struct GrpGemmParams {
    const uint8_t* w;          // expert weights
    long expert_bytes;         // bytes per expert
    int R, Kc;                 // rows per expert, reduction length
    const int8_t* xq;          // activations: int8 [rows][Kc]
    const float* xd;           // [rows][Kc / 256]
    const int16_t* xbs;        // [rows][Kc / 16]
    const int* rowmap;         // sorted position -> activation row (null: identity)
    const int4* tinfo;         // [tiles] {expert, first weight row, first sorted position, valid tokens}
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
// the 6-bit (scale, min) pair j of a Q4_K header held as 4 words (get_scale_min_k4, ggml-quants.c)
__device__ __forceinline__ void q4k_scale_min(const uint32_t* hw, int j, int& sc, int& mn) {
    if (j < 4) { sc = ub(hw, 4 + j) & 63; mn = ub(hw, 8 + j) & 63; }
    else { sc = (ub(hw, 8 + j) & 0xF) | ((ub(hw, j) >> 6) << 4); mn = (ub(hw, 8 + j) >> 4) | ((ub(hw, 4 + j) >> 6) << 4); }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// tile table: tile -> (expert, row tile, token tile), token tile fastest so that CTAs running side by side share the weight tile
// through L2.  One thread per tile; tiles beyond the data-dependent total are left alone.
__global__ void grp_tiles_kernel(const int* nt_prefix, const int* offsets, int E, int MT, int4* tinfo) {
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= nt_prefix[E] * MT) return;
    int lo = 0, hi = E;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (nt_prefix[mid] * MT <= tile) lo = mid; else hi = mid;
    }
    const int local = tile - nt_prefix[lo] * MT, ntile_e = nt_prefix[lo + 1] - nt_prefix[lo];
    const int mt = local / ntile_e, nt = local - mt * ntile_e;
    const int p0 = offsets[lo] + nt * kGN;
    tinfo[tile] = make_int4(lo, mt * kGM, p0, min(kGN, offsets[lo + 1] - p0));
}

// The producer's share of one stage, as it sits in its raw-ring slot (16-byte units):
//   Q4_K: 0-3 qs (64 bytes of the half), 4 block header, 5 token scale (4 bytes), 6-7 activation pieces, 8 activation 16-sums
//   Q6_K: 0-3 ql, 4-5 qh, 6-7 activation pieces, 8 = 8 scales | d (2 of 4 bytes) | token scale
template <int FMT>
__global__ void __launch_bounds__(kGThreads, 1) grouped_gemm_kernel(const GrpGemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    GrpMisc& misc = *reinterpret_cast<GrpMisc*>(smem + kOffMiscG);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = p.Kc / QK_K, nst = 2 * nblk, MT = p.R / kGM;
    if (tid == 0) {
        for (int s = 0; s < kGStages; s++) { bar_init(smem_u32(&misc.ab_full[s]), kGProdWarps); bar_init(smem_u32(&misc.smem_free[s]), 1 + kGEpiWarps); }
        for (int b = 0; b < 2; b++) { bar_init(smem_u32(&misc.tmem_full[b]), 1); bar_init(smem_u32(&misc.tmem_free[b]), kGEpiWarps); }
        bar_fence_init();
    }
    if (warp == kGProdWarps) tmem_alloc(smem_u32(&misc.tmem_base), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = misc.tmem_base;
    const int total_tiles = p.nt_prefix[p.E] * MT;
    unsigned it = 0;   // stages done by this CTA: smem stage = it % kGStages, TMEM buffer = it & 1

    if (warp < kGProdWarps) {
        // ========================================================================== producers: thread = weight row r of the tile
        const int r = tid, sw = r & 7, pc = r & 7;
        const uint32_t raw_dst = base + kOffRaw + r * kRawPitch;
        const uint8_t* raw_src = smem + kOffRaw + r * kRawPitch;
        // fetch cursor: runs kGRaw stages ahead of the conversion, across tile boundaries
        int ftile = blockIdx.x, fst = 0, frow0 = -1, frow1 = -1, frow2 = -1, frw = 0;
        const uint8_t* fw = nullptr;
        auto enter_tile = [&]() {
            if (ftile >= total_tiles) return;
            const int4 ti = __ldg(p.tinfo + ftile);
            const int n0 = r >> 3, n1 = 16 + (r >> 3), n2 = r >> 1;
            frow0 = n0 < ti.w ? (p.rowmap ? __ldg(p.rowmap + ti.z + n0) : ti.z + n0) : -1;
            frow1 = n1 < ti.w ? (p.rowmap ? __ldg(p.rowmap + ti.z + n1) : ti.z + n1) : -1;
            frow2 = (r < 64 && n2 < ti.w) ? (p.rowmap ? __ldg(p.rowmap + ti.z + n2) : ti.z + n2) : -1;
            const int row = ti.y + r;
            frw = row & 3;
            fw = p.w + (long)ti.x * p.expert_bytes + (FMT == 0 ? (long)row * nblk * SZ_Q4_K : (long)(row >> 2) * 4 * nblk * SZ_Q6_K);
        };
        auto issue = [&](int slot) {
            if (ftile < total_tiles) {
                const uint32_t dst = raw_dst + slot * kRawSlot;
                const int sb = fst >> 1, hh = fst & 1;
                if (FMT == 0) {
                    const uint8_t* blk = fw + (long)sb * SZ_Q4_K;
                    cp_async16(dst + 64, blk);
#pragma unroll
                    for (int i = 0; i < 4; i++) cp_async16(dst + 16 * i, blk + 16 + hh * 64 + i * 16);
                } else {
                    const int nrb = 4 * nblk, fi = frw * nblk + sb;
#pragma unroll
                    for (int i = 0; i < 4; i++) cp_async16(dst + 16 * i, fw + (long)(4 * hh + i) * nrb * 16 + fi * 16);
#pragma unroll
                    for (int i = 0; i < 2; i++) cp_async16(dst + 64 + 16 * i, fw + (long)nrb * 128 + (long)(2 * hh + i) * nrb * 16 + fi * 16);
                    cp_async8(dst + 128, fw + (long)nrb * 192 + fi * 16 + hh * 8);
                    cp_async4(dst + 136, fw + (long)nrb * 208 + (fi >> 1) * 4);
                }
                if (frow0 >= 0) cp_async16(dst + 96, p.xq + (long)frow0 * p.Kc + fst * 128 + pc * 16);
                if (frow1 >= 0) cp_async16(dst + 112, p.xq + (long)frow1 * p.Kc + fst * 128 + pc * 16);
                if (hh == 1 && frow2 >= 0) {
                    if ((r & 1) == 0) cp_async4(dst + (FMT == 0 ? 80 : 140), p.xd + (long)frow2 * nblk + sb);
                    if (FMT == 0) cp_async16(dst + 128, p.xbs + (long)frow2 * (p.Kc / 16) + sb * 16 + (r & 1) * 8);
                }
                if (++fst == nst) { fst = 0; ftile += gridDim.x; enter_tile(); }
            }
            cp_async_commit();
        };
        enter_tile();
        for (int i = 0; i < kGRaw; i++) issue(i);
        unsigned cc = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int4 ti = __ldg(p.tinfo + tile);
            const int n_valid = ti.w;
            for (int st = 0; st < nst; st++, it++, cc++) {
                const int slot = cc % kGRaw, stage = it % kGStages, hh = st & 1;
                cp_async_wait<kGRaw - 1>();
                const uint4* rs = reinterpret_cast<const uint4*>(raw_src + slot * kRawSlot);
                uint4 f[9];
#pragma unroll
                for (int i = 0; i < 9; i++) f[i] = rs[i];
                issue(slot);   // refill the slot just read (thread-private bytes: no barrier involved)
                bar_wait(smem_u32(&misc.smem_free[stage]), ((it / kGStages) & 1) ^ 1);
                uint8_t* arow = smem + stage * kGA + r * 128;
                if (FMT == 0) {
                    // chunk c = 2 hh + part (32 bytes of qs): low nibbles = sub-block 2c (elements 64c .. 64c+31), high nibbles = sub-block 2c+1
#pragma unroll
                    for (int part = 0; part < 2; part++) {
                        const uint4 q0 = f[2 * part], q1 = f[2 * part + 1];
                        const int pi = 4 * part;
                        *reinterpret_cast<uint4*>(arow + (((pi + 0) ^ sw) << 4)) = make_uint4(q0.x & 0x0F0F0F0Fu, q0.y & 0x0F0F0F0Fu, q0.z & 0x0F0F0F0Fu, q0.w & 0x0F0F0F0Fu);
                        *reinterpret_cast<uint4*>(arow + (((pi + 1) ^ sw) << 4)) = make_uint4(q1.x & 0x0F0F0F0Fu, q1.y & 0x0F0F0F0Fu, q1.z & 0x0F0F0F0Fu, q1.w & 0x0F0F0F0Fu);
                        *reinterpret_cast<uint4*>(arow + (((pi + 2) ^ sw) << 4)) =
                            make_uint4((q0.x >> 4) & 0x0F0F0F0Fu, (q0.y >> 4) & 0x0F0F0F0Fu, (q0.z >> 4) & 0x0F0F0F0Fu, (q0.w >> 4) & 0x0F0F0F0Fu);
                        *reinterpret_cast<uint4*>(arow + (((pi + 3) ^ sw) << 4)) =
                            make_uint4((q1.x >> 4) & 0x0F0F0F0Fu, (q1.y >> 4) & 0x0F0F0F0Fu, (q1.z >> 4) & 0x0F0F0F0Fu, (q1.w >> 4) & 0x0F0F0F0Fu);
                    }
                    misc.hdr[stage][r] = f[4];
                    if (hh == 1) {   // A2 row: [m_0 m_0 m_1 m_1 ... m_7 m_7] against the sixteen 16-value activation sums
                        const uint32_t hw[4] = {f[4].x, f[4].y, f[4].z, f[4].w};
                        int sc, mn[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) q4k_scale_min(hw, j, sc, mn[j]);
                        uint8_t* a2 = smem + kOffA2 + stage * kGA2 + (r >> 3) * 256 + (r & 7) * 16;
                        *reinterpret_cast<uint4*>(a2) = make_uint4(h2(mn[0], mn[0]), h2(mn[1], mn[1]), h2(mn[2], mn[2]), h2(mn[3], mn[3]));
                        *reinterpret_cast<uint4*>(a2 + 128) = make_uint4(h2(mn[4], mn[4]), h2(mn[5], mn[5]), h2(mn[6], mn[6]), h2(mn[7], mn[7]));
                    }
                } else {
                    // element 32 g + l of the half (l = 16 part + 0..15): g = 0 ql[l] & 15 | (qh & 3) << 4, g = 1 ql[32 + l] & 15 | (qh >> 2 & 3) << 4,
                    // g = 2 ql[l] >> 4 | (qh >> 4 & 3) << 4, g = 3 ql[32 + l] >> 4 | (qh >> 6 & 3) << 4; stored as q - 32 in int8
#pragma unroll
                    for (int part = 0; part < 2; part++) {
                        const uint4 wa = f[part], wb = f[2 + part], wh = f[4 + part];
                        const uint32_t a[4] = {wa.x, wa.y, wa.z, wa.w}, b[4] = {wb.x, wb.y, wb.z, wb.w}, h[4] = {wh.x, wh.y, wh.z, wh.w};
                        uint32_t v[4][4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            v[0][i] = (a[i] & 0x0F0F0F0Fu) | ((h[i] << 4) & 0x30303030u);
                            v[1][i] = (b[i] & 0x0F0F0F0Fu) | ((h[i] << 2) & 0x30303030u);
                            v[2][i] = ((a[i] >> 4) & 0x0F0F0F0Fu) | (h[i] & 0x30303030u);
                            v[3][i] = ((b[i] >> 4) & 0x0F0F0F0Fu) | ((h[i] >> 2) & 0x30303030u);
                        }
#pragma unroll
                        for (int g = 0; g < 4; g++) {
#pragma unroll
                            for (int i = 0; i < 4; i++) {   // q - 32 per byte: flip bit 5, then copy it into bits 6 and 7 (no carries between bytes)
                                const uint32_t tt = v[g][i] ^ 0x20202020u;
                                v[g][i] = tt + (tt & 0x20202020u) * 6u;
                            }
                            *reinterpret_cast<uint4*>(arow + (((2 * g + part) ^ sw) << 4)) = make_uint4(v[g][0], v[g][1], v[g][2], v[g][3]);
                        }
                    }
                    const int fi = ((ti.y + r) & 3) * nblk + (st >> 1);
                    const uint32_t dbits = (fi & 1) ? (f[8].z >> 16) : (f[8].z & 0xffffu);
                    misc.hdr[stage][r] = make_uint4(f[8].x, f[8].y, __float_as_uint(fp16_bits_to_f32((uint16_t)dbits)), 0);
                }
                // activations: pieces (n, pc) and (16 + n, pc)
                uint8_t* Bs = smem + kOffB + stage * kGB;
                const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int n = (r >> 3) + 16 * u;
                    const uint4 bv = n < n_valid ? f[6 + u] : z;
                    if (FMT == 0) {
                        *reinterpret_cast<uint4*>(Bs + n * 128 + ((pc ^ (n & 7)) << 4)) = bv;
                    } else {   // a 16-byte piece is one Q6_K sub-block: rows 0-31 keep the even pieces, rows 32-63 the odd ones
                        *reinterpret_cast<uint4*>(Bs + n * 128 + ((pc ^ (n & 7)) << 4)) = (pc & 1) ? z : bv;
                        *reinterpret_cast<uint4*>(Bs + (kGN + n) * 128 + ((pc ^ (n & 7)) << 4)) = (pc & 1) ? bv : z;
                    }
                }
                if (hh == 1 && r < 64) {   // token scales, and (Q4_K) the sixteen 16-value sums of the super-block as fp16
                    const int n2 = r >> 1, kg = r & 1;
                    const bool ok = n2 < n_valid;
                    if (kg == 0) misc.dxs[stage][n2] = ok ? __uint_as_float(FMT == 0 ? f[5].x : f[8].w) : 0.f;
                    if (FMT == 0) {
                        uint4 vv = z;
                        if (ok) {
                            const uint32_t bw[4] = {f[8].x, f[8].y, f[8].z, f[8].w};
#define KTB_S16(w, hi) ((int)(short)((hi) ? ((w) >> 16) : ((w) & 0xffffu)))
                            vv = make_uint4(h2(KTB_S16(bw[0], 0), KTB_S16(bw[0], 1)), h2(KTB_S16(bw[1], 0), KTB_S16(bw[1], 1)), h2(KTB_S16(bw[2], 0), KTB_S16(bw[2], 1)),
                                            h2(KTB_S16(bw[3], 0), KTB_S16(bw[3], 1)));
#undef KTB_S16
                        }
                        *reinterpret_cast<uint4*>(smem + kOffB2 + stage * kGB2 + (n2 >> 3) * 256 + kg * 128 + (n2 & 7) * 16) = vv;
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) bar_arrive(smem_u32(&misc.ab_full[stage]));
            }
        }
        cp_async_wait<0>();
    } else if (warp == kGProdWarps) {
        // ========================================================================== tensor-core issuer (converged warp)
        constexpr uint32_t idesc_i8 = FMT == 0 ? instr_desc(2, 0, 1, 0, 0, kGM, kGN) : instr_desc(2, 1, 1, 0, 0, kGM, 2 * kGN);   // s32 += (u8 | s8) . s8
        constexpr uint32_t idesc_f16 = instr_desc(1, 0, 0, 0, 0, kGM, kGN);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            for (int st = 0; st < nst; st++, it++) {
                const int stage = it % kGStages, buf = it & 1;
                bar_wait(smem_u32(&misc.ab_full[stage]), (it / kGStages) & 1);
                bar_wait(smem_u32(&misc.tmem_free[buf]), ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t a = base + stage * kGA, b = base + kOffB + stage * kGB, d = tmem + buf * 256;
#pragma unroll
                for (int c = 0; c < 4; c++)
                    mma_i8(d + c * (FMT == 0 ? 32 : 64), smem_desc(a + c * 32, 16, 1024, kLayoutSw128), smem_desc(b + c * 32, 16, 1024, kLayoutSw128), idesc_i8, 0);
                if (FMT == 0 && (st & 1))
                    mma_f16(d + 128, smem_desc(base + kOffA2 + stage * kGA2, 128, 256, kLayoutNone), smem_desc(base + kOffB2 + stage * kGB2, 128, 256, kLayoutNone), idesc_f16, 0);
                mma_commit(smem_u32(&misc.smem_free[stage]));
                mma_commit(smem_u32(&misc.tmem_full[buf]));
            }
        }
    } else {
        // ========================================================================== epilogue: 8 warps = 128 rows x 2 column halves
        const int ew = warp - kGProdWarps - 1, sp = warp & 3, ch = ew >> 2, row = 32 * sp + lane;
        const uint32_t tbase = tmem + ((uint32_t)(32 * sp) << 16) + 16 * ch;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int4 ti = __ldg(p.tinfo + tile);
            float acc[16];
            int isum[16];
#pragma unroll
            for (int n = 0; n < 16; n++) { acc[n] = 0.f; isum[n] = 0; }
            for (int st = 0; st < nst; st++, it++) {
                const int stage = it % kGStages, buf = it & 1, hh = st & 1;
                bar_wait(smem_u32(&misc.tmem_full[buf]), (it >> 1) & 1);
                tc_fence_after();
                const uint4 hd = misc.hdr[stage][row];
                const uint32_t hw[4] = {hd.x, hd.y, hd.z, hd.w};
                const uint32_t d = tbase + buf * 256;
                if (FMT == 0) {
                    int sc[4], mn;
                    if (hh == 0) {
#pragma unroll
                        for (int c = 0; c < 4; c++) q4k_scale_min(hw, c, sc[c], mn);
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; c++) q4k_scale_min(hw, 4 + c, sc[c], mn);
                    }
#pragma unroll
                    for (int c = 0; c < 4; c += 2) {
                        uint32_t v0[16], v1[16];
                        tmem_ld16(d + 32 * c, v0);
                        tmem_ld16(d + 32 * c + 32, v1);
                        tmem_wait_ld();
#pragma unroll
                        for (int n = 0; n < 16; n++) isum[n] += sc[c] * (int)v0[n] + sc[c + 1] * (int)v1[n];
                    }
                    if (hh == 1) {
                        uint32_t ms[16];
                        tmem_ld16(d + 128, ms);
                        const __half2 dm = *reinterpret_cast<const __half2*>(&hw[0]);
                        const float dw = __low2float(dm), dmin = __high2float(dm);
                        tmem_wait_ld();
#pragma unroll
                        for (int n = 0; n < 16; n++) {
                            const float dx = misc.dxs[stage][16 * ch + n];
                            acc[n] += (dw * dx) * (float)isum[n] - (dmin * dx) * __uint_as_float(ms[n]);
                            isum[n] = 0;
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        uint32_t ve[16], vo[16];
                        tmem_ld16(d + 64 * c, ve);
                        tmem_ld16(d + 64 * c + 32, vo);
                        const int se = sb8(hw, 2 * c), so = sb8(hw, 2 * c + 1);
                        tmem_wait_ld();
#pragma unroll
                        for (int n = 0; n < 16; n++) isum[n] += se * (int)ve[n] + so * (int)vo[n];
                    }
                    if (hh == 1) {
                        const float dw = __uint_as_float(hw[2]);
#pragma unroll
                        for (int n = 0; n < 16; n++) {
                            acc[n] += (dw * misc.dxs[stage][16 * ch + n]) * (float)isum[n];
                            isum[n] = 0;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { bar_arrive(smem_u32(&misc.tmem_free[buf])); bar_arrive(smem_u32(&misc.smem_free[stage])); }
            }
#pragma unroll
            for (int n = 0; n < 16; n++)
                if (16 * ch + n < ti.w) p.out[(long)(ti.z + 16 * ch + n) * p.R + ti.y + row] = acc[n];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kGProdWarps) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
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
    int4 *tinfo_gu = nullptr, *tinfo_d = nullptr;
    size_t cap_tiles_gu = 0, cap_tiles_d = 0;
    int8_t *xq = nullptr, *aq = nullptr;
    float *xd = nullptr, *ad = nullptr, *g = nullptr, *u = nullptr, *dd = nullptr;
    int16_t *xbs = nullptr, *abs16 = nullptr;
};
static GrpScratch g_grp[64];   // one arena per device, shared by every handle (calls on one device are stream-ordered by the caller)

static int grp_ensure(int dev, int tokens, int k, int E, int H, int I) {
    GrpScratch& s = g_grp[dev & 63];
    size_t P = (size_t)tokens * k, nx = (size_t)tokens * H, na = P * I, nd = P * H;
    size_t tg = (P / kGN + E) * (size_t)(I / kGM), td = (P / kGN + E) * (size_t)(H / kGM);   // upper bounds of the tile counts
    if (s.cap_pairs >= P && s.cap_x >= nx && s.cap_a >= na && s.cap_d >= nd && s.cap_tiles_gu >= tg && s.cap_tiles_d >= td) return KTB200_OK;
    tg = tg > s.cap_tiles_gu ? tg : s.cap_tiles_gu; td = td > s.cap_tiles_d ? td : s.cap_tiles_d;
    P = P > s.cap_pairs ? P : s.cap_pairs; nx = nx > s.cap_x ? nx : s.cap_x; na = na > s.cap_a ? na : s.cap_a; nd = nd > s.cap_d ? nd : s.cap_d;   // grow only
    KTB_CUDA_CHECK(cudaDeviceSynchronize());   // earlier calls may still be using the arena
    cudaFree(s.counts); cudaFree(s.tokmap); cudaFree(s.pos); cudaFree(s.xq); cudaFree(s.xd); cudaFree(s.xbs); cudaFree(s.aq); cudaFree(s.ad); cudaFree(s.abs16);
    cudaFree(s.g); cudaFree(s.u); cudaFree(s.dd); cudaFree(s.tinfo_gu); cudaFree(s.tinfo_d);
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
    KTB_CUDA_CHECK(cudaMalloc(&s.tinfo_gu, tg * sizeof(int4)));
    KTB_CUDA_CHECK(cudaMalloc(&s.tinfo_d, td * sizeof(int4)));
    s.cap_pairs = cp; s.cap_x = nx; s.cap_a = na; s.cap_d = nd; s.cap_tiles_gu = tg; s.cap_tiles_d = td;
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
    int rc = grp_ensure(dev, Tc, k, E, H, I);   // grow-only scratch: not capturable on first use (like ktb200_moe_gate_forward)
    if (rc) return rc;
    GrpScratch& g = g_grp[dev & 63];
    static bool attr[64] = {};
    if (!attr[dev & 63]) {
        KTB_CUDA_CHECK(cudaFuncSetAttribute(grouped_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSmem));
        KTB_CUDA_CHECK(cudaFuncSetAttribute(grouped_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSmem));
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
        const int ub_gu = (P / kGN + E) * (I / kGM), ub_d = (P / kGN + E) * (H / kGM);
        grp_tiles_kernel<<<(ub_gu + 255) / 256, 256, 0, s>>>(g.nt_prefix, g.offsets, E, I / kGM, g.tinfo_gu);
        grp_tiles_kernel<<<(ub_d + 255) / 256, 256, 0, s>>>(g.nt_prefix, g.offsets, E, H / kGM, g.tinfo_d);
        grp_quant_x_kernel<<<(T * (H / 256) + 7) / 8, 256, 0, s>>>(x_c, c.hidden_type, T, H, g.xq, g.xd, g.xbs);
        GrpGemmParams gp{};
        gp.R = I; gp.Kc = H; gp.xq = g.xq; gp.xd = g.xd; gp.xbs = g.xbs; gp.rowmap = g.tokmap; gp.tinfo = g.tinfo_gu; gp.nt_prefix = g.nt_prefix; gp.E = E;
        gp.expert_bytes = (long)I * (H / 256) * SZ_Q4_K;
        gp.w = reinterpret_cast<const uint8_t*>(c.gate_proj); gp.out = g.g;
        grouped_gemm_kernel<0><<<grid, kGThreads, kGSmem, s>>>(gp);
        gp.w = reinterpret_cast<const uint8_t*>(c.up_proj); gp.out = g.u;
        grouped_gemm_kernel<0><<<grid, kGThreads, kGSmem, s>>>(gp);
        grp_act_quant_kernel<<<(P * (I / 256) + 7) / 8, 256, 0, s>>>(g.g, g.u, g.offsets, E, I, c.use_silu, g.aq, g.ad, g.abs16);
        GrpGemmParams gd{};
        gd.R = H; gd.Kc = I; gd.xq = g.aq; gd.xd = g.ad; gd.xbs = g.abs16; gd.rowmap = nullptr; gd.tinfo = g.tinfo_d; gd.nt_prefix = g.nt_prefix;
        gd.E = E; gd.expert_bytes = (long)H * (I / 256) * (fd == FMT_Q6K4T ? SZ_Q6_K : SZ_Q4_K);
        gd.w = reinterpret_cast<const uint8_t*>(c.down_proj); gd.out = g.dd;
        if (fd == FMT_Q6K4T) grouped_gemm_kernel<1><<<grid, kGThreads, kGSmem, s>>>(gd);
        else grouped_gemm_kernel<0><<<grid, kGThreads, kGSmem, s>>>(gd);
        grp_combine_kernel<<<dim3((H + 255) / 256, T), 256, 0, s>>>(g.dd, g.pos, w_c, T, k, H, bsz, t0, o_c, c.hidden_type);
        KTB_LAUNCH_CHECK();
        count_launch(9);   // + the one KTB_LAUNCH_CHECK counts = 10 launches per chunk
    }
    return KTB200_OK;
}

}  // namespace ktb
