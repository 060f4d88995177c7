// Fourth generation of the expert GEMV kernels: bulk-copy (TMA engine) staging + one lane per super-block.
//
// What ncu said about the cp.async generation (profiles/r01c_kernels.md): DRAM at ~50 % of peak, issue slots
// 40-50 % busy with only 12 warps per SM, and of the ~700 warp instructions a warp spent per 8 KB unit almost
// 230 were the copy itself (16 LDGSTS per lane, each with its own 64-bit address arithmetic) plus index
// divisions — LSU/MIO time shared with the LDS of the dot product.  Here:
//
//   * every ring slot is filled by ONE `cp.async.bulk.shared.global` issued by one lane and tracked by an
//     mbarrier (expect_tx / complete_tx): no per-lane copy instructions, no LSU time for the copy;
//   * slots are one weight row (Q4_K gate/up: nblk x 144 B) or one 4-row item (down) and a warp keeps SLOTS-1
//     of them in flight while it computes on one: smaller slots -> up to 18 warps per SM (was 12);
//   * unit -> (slot, row) bookkeeping is incremental (one division per token instead of one per unit);
//   * Q6_K down tensors are re-tiled once at load time into 4-row "chunk-major" items (`repack_q6k4t`, moe.cu):
//       item = rows 4q..4q+3 of one expert, f = rw*nb + blk (nrb = 4*nb (row, block) pairs)
//       [ql: c=0..7][f][16 B] | [qh: c=0..3][f][16 B] | [scales: f][16 B] | [d: f][2 B]       (= nrb * 210 B)
//     so an item is ONE contiguous bulk copy and lane f's 16-byte loads of chunk c sit next to lane f+1's:
//     bank-conflict free without padding.  Q4_K rows need no re-tiling (144-byte blocks, 36-word stride).
//
// Arithmetic is unchanged from the earlier generations (bit-exact Q8_K activations, integer dot products, fp32
// once per super-block): see DESIGN.md §2.
#pragma once
#include "gemv_pipe.cuh"

namespace ktb {

constexpr int kActBlkStride = QK_K + 16;   // int8 activation blocks padded to 272 B: 8 lanes x LDS.128 hit 32 distinct banks
constexpr int kBulkMaxWarps = 18;        // gate/up kernel (<= 96 registers per thread)
constexpr int kBulkMaxWarpsDown = 16;    // down kernel: 128 registers per thread, and shared memory caps it at 15 anyway

// ---------------------------------------------------------------------------------------------------------------
// mbarrier / bulk-copy PTX
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// global -> shared bulk copy (size and both addresses multiples of 16 B); completion is signalled on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// One Q4_K super-block (144 B at `wb`, shared memory, 16-byte aligned) against one padded int8 activation block.
// bsv = the block's eight 32-value activation sums (int16), dxb = the activation block scale.
// `act4(i)` returns the i-th 16-byte word of the block's 256 int8 activations (from shared memory or from registers).
template <class ActFn>
__device__ __forceinline__ float q4k_block_dot_t(const uint8_t* wb, ActFn&& act4, const uint4 bsv, const float dxb) {
    const uint4 hdr = *reinterpret_cast<const uint4*>(wb);
    const float2 dm = __half22float2(*reinterpret_cast<const __half2*>(&hdr.x));
    const uint32_t scl = hdr.y & 0x3f3f3f3fu;                                          // scales 0..3
    const uint32_t mnl = hdr.z & 0x3f3f3f3fu;                                          // mins   0..3
    const uint32_t sch = (hdr.w & 0x0f0f0f0fu) | ((hdr.y >> 2) & 0x30303030u);         // scales 4..7
    const uint32_t mnh = ((hdr.w >> 4) & 0x0f0f0f0fu) | ((hdr.z >> 2) & 0x30303030u);  // mins   4..7
    int isum = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const uint4 a0 = act4(4 * g), a1 = act4(4 * g + 1), a2 = act4(4 * g + 2), a3 = act4(4 * g + 3);
        const uint4 q0 = *reinterpret_cast<const uint4*>(wb + 16 + 32 * g);
        const uint4 q1 = *reinterpret_cast<const uint4*>(wb + 32 + 32 * g);
        int slo = 0, shi = 0, slo2 = 0, shi2 = 0;
        slo = dp4a_s8s8(q0.x & 0x0f0f0f0fu, a0.x, slo); slo2 = dp4a_s8s8(q0.y & 0x0f0f0f0fu, a0.y, slo2);
        slo = dp4a_s8s8(q0.z & 0x0f0f0f0fu, a0.z, slo); slo2 = dp4a_s8s8(q0.w & 0x0f0f0f0fu, a0.w, slo2);
        slo = dp4a_s8s8(q1.x & 0x0f0f0f0fu, a1.x, slo); slo2 = dp4a_s8s8(q1.y & 0x0f0f0f0fu, a1.y, slo2);
        slo = dp4a_s8s8(q1.z & 0x0f0f0f0fu, a1.z, slo); slo2 = dp4a_s8s8(q1.w & 0x0f0f0f0fu, a1.w, slo2);
        // high nibbles stay in place (x16): the sums are exact multiples of 16
        shi = dp4a_u8s8(q0.x & 0xf0f0f0f0u, a2.x, shi); shi2 = dp4a_u8s8(q0.y & 0xf0f0f0f0u, a2.y, shi2);
        shi = dp4a_u8s8(q0.z & 0xf0f0f0f0u, a2.z, shi); shi2 = dp4a_u8s8(q0.w & 0xf0f0f0f0u, a2.w, shi2);
        shi = dp4a_u8s8(q1.x & 0xf0f0f0f0u, a3.x, shi); shi2 = dp4a_u8s8(q1.y & 0xf0f0f0f0u, a3.y, shi2);
        shi = dp4a_u8s8(q1.z & 0xf0f0f0f0u, a3.z, shi); shi2 = dp4a_u8s8(q1.w & 0xf0f0f0f0u, a3.w, shi2);
        const uint32_t scw = (g < 2) ? scl : sch;
        const int sc0 = (int)((scw >> (16 * (g & 1))) & 0xff), sc1 = (int)((scw >> (16 * (g & 1) + 8)) & 0xff);
        isum += sc0 * (slo + slo2) + sc1 * ((shi + shi2) >> 4);
    }
    int msum = __dp2a_lo((int)bsv.x, (int)mnl, 0);
    msum = __dp2a_hi((int)bsv.y, (int)mnl, msum);
    msum = __dp2a_lo((int)bsv.z, (int)mnh, msum);
    msum = __dp2a_hi((int)bsv.w, (int)mnh, msum);
    return (dm.x * dxb) * (float)isum - (dm.y * dxb) * (float)msum;
}
__device__ __forceinline__ float q4k_block_dot(const uint8_t* wb, const uint8_t* aq, const uint4 bsv, const float dxb) {
    return q4k_block_dot_t(wb, [&](int i) { return *reinterpret_cast<const uint4*>(aq + 16 * i); }, bsv, dxb);
}

// ---------------------------------------------------------------------------------------------------------------
// Gate/up (PAIR) or dense (PAIR = false) rows of Q4_K tensors.  Per warp: a private ring of SLOTS row slots; the
// warp's stream of sub-units is g(u0), u(u0), g(u0+W), u(u0+W), ... (PAIR) and at any time SLOTS-1 rows are in
// flight behind the one being consumed.
//
// Tokens are processed in chunks of `tc` (launcher: as many as fit next to >= 12 rings).  Within a chunk ALL
// (token, slot) pairs this launch owns form ONE work list — expert-parallel shards and decode batches keep the ring
// streaming across tokens instead of draining it once per token — and all the chunk's activation rows are staged as
// Q8_K side by side (`act_tok` bytes each).
template <bool PAIR, int SLOTS>
__global__ void __launch_bounds__(kBulkMaxWarps * 32, 1) rows_bulk_q4k_kernel(const RowsParams p, int act_tok, int tc) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_np;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    int Teff = p.ntokens;
    if (p.bsz) Teff = min(Teff, *p.bsz);
    const int nblk = p.ncols / QK_K;
    const int row_bytes = nblk * SZ_Q4_K;
    constexpr int NM = PAIR ? 2 : 1;
    const int nslots = p.slots + (p.x0 ? 1 : 0);
    const int total_out = nslots * p.rows;
    // [tc activation rows: q8 [nblk][272] | bs32 [nblk][8] int16 | dx [nblk]] [pair list: tc*nslots ints] [mbarriers] [rings]
    int* pairs = reinterpret_cast<int*>(smem + (size_t)tc * act_tok);                 // (token in chunk) << 8 | slot
    const size_t off = ((size_t)tc * act_tok + (size_t)tc * nslots * 4 + 15) & ~(size_t)15;
    const int bar_bytes = (W * SLOTS * 8 + 15) & ~15;
    const uint32_t bar_u32 = (uint32_t)__cvta_generic_to_shared(smem + off) + warp * SLOTS * 8;
    uint8_t* ring = smem + off + bar_bytes + (size_t)warp * SLOTS * row_bytes;
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < SLOTS; s++) mbar_init(bar_u32 + 8 * s, 1);
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    int slot_i = 0, slot_u = 0;   // ring cursors (issue / use); they advance in lock step over the whole launch
    uint32_t phase = 0;           // bit s = parity the next use of slot s waits for

  for (int t0 = 0; t0 < Teff; t0 += tc) {
    const int nt = min(tc, Teff - t0);
    __syncthreads();   // previous chunk: everyone is done with the staging and the pair list (and the barriers are initialised)
    if (threadIdx.x == 0) {
        int np = 0;
        for (int tl = 0; tl < nt; tl++) {
            for (int s = 0; s < p.slots; s++) {
                const long e = p.ids ? (long)p.ids[(long)(t0 + tl) * p.slots + s] - p.id_offset : 0;
                if (e >= 0 && e < p.n_experts) pairs[np++] = (tl << 8) | s;
            }
            if (p.x0 && (p.shared_token < 0 || p.shared_token == t0 + tl)) pairs[np++] = (tl << 8) | p.slots;
        }
        s_np = np;
    }
    __syncthreads();
    const int total = s_np * p.rows;
    const int u0 = (int)((long)total * blockIdx.x / gridDim.x), u1 = (int)((long)total * (blockIdx.x + 1) / gridDim.x);
    int nu = u1 - u0 - warp;
    nu = nu > 0 ? (nu + W - 1) / W : 0;                     // units of this warp: u0 + warp + i*W
    const int nsub = nu * NM;
    // issue cursor: (pair index, row) of the next unit to request, and how many rows were requested
    int ipi = 0, irr = 0, isub = 0;
    if (nu > 0) { ipi = (u0 + warp) / p.rows; irr = (u0 + warp) - ipi * p.rows; }
    int cpi = ipi, crr = irr;                               // consume cursor

    auto issue_one = [&]() {
        if (isub < nsub) {
            if (lane == 0) {
                const int pr = pairs[ipi], s = pr & 0xff, tl = pr >> 8;
                const bool second = PAIR && (isub & 1);
                const uint8_t* src;
                if (s == p.slots) {
                    src = reinterpret_cast<const uint8_t*>(second ? p.x1 : p.x0) + (long)irr * row_bytes;
                } else {
                    const long e = p.ids ? (long)p.ids[(long)(t0 + tl) * p.slots + s] - p.id_offset : 0;
                    src = reinterpret_cast<const uint8_t*>(second ? p.w1 : p.w0) + (e * p.rows + irr) * row_bytes;
                }
                const uint32_t bar = bar_u32 + 8 * slot_i;
                mbar_expect_tx(bar, (uint32_t)row_bytes);
                bulk_g2s(ring_u32 + slot_i * row_bytes, src, (uint32_t)row_bytes, bar);
            }
            isub++;
            if (!PAIR || !(isub & 1)) {
                irr += W;
                while (irr >= p.rows) { irr -= p.rows; ipi++; }
            }
            slot_i = (slot_i + 1 == SLOTS) ? 0 : slot_i + 1;
        }
    };
#pragma unroll
    for (int s = 0; s < SLOTS; s++) issue_one();

    {   // quantise the chunk's activation rows into the padded layout: block g = (token in chunk, block of the row)
        float cur[8], nxt[8];
        const int totalb = nt * nblk;
        int g = warp;
        if (g < totalb) load_block8(p.x, (long)(t0 + g / nblk) * p.ncols + (long)(g % nblk) * QK_K + lane * 8, p.hidden_type, cur);
#pragma unroll 1
        while (g < totalb) {
            const int gn = g + W;
            if (gn < totalb) load_block8(p.x, (long)(t0 + gn / nblk) * p.ncols + (long)(gn % nblk) * QK_K + lane * 8, p.hidden_type, nxt);
            const int tl = g / nblk, b = g - tl * nblk;
            uint8_t* at = smem + (size_t)tl * act_tok;
            warp_quantize_q8k_block(cur, lane, reinterpret_cast<uint32_t*>(at + (size_t)b * kActBlkStride),
                                    reinterpret_cast<float*>(at + (size_t)nblk * (kActBlkStride + 16)) + b, nullptr,
                                    reinterpret_cast<int16_t*>(at + (size_t)nblk * kActBlkStride) + b * 8);
#pragma unroll
            for (int i = 0; i < 8; i++) cur[i] = nxt[i];
            g = gn;
        }
    }
    __syncthreads();

    float acc_first = 0.f;
    for (int n = 0; n < nsub; n++) {
        mbar_wait(bar_u32 + 8 * slot_u, (phase >> slot_u) & 1u);
        phase ^= 1u << slot_u;
        const uint8_t* row0 = ring + slot_u * row_bytes;
        const int pr = pairs[cpi];
        const uint8_t* at = smem + (size_t)(pr >> 8) * act_tok;
        const int16_t* bs32 = reinterpret_cast<const int16_t*>(at + (size_t)nblk * kActBlkStride);
        const float* dx = reinterpret_cast<const float*>(at + (size_t)nblk * (kActBlkStride + 16));
        float acc = 0.f;
        for (int blk = lane; blk < nblk; blk += 32)
            acc += q4k_block_dot(row0 + blk * SZ_Q4_K, at + (size_t)blk * kActBlkStride,
                                 *reinterpret_cast<const uint4*>(bs32 + blk * 8), dx[blk]);
        __syncwarp();                       // every lane is done reading the slot: hand it back to the copy engine
        slot_u = (slot_u + 1 == SLOTS) ? 0 : slot_u + 1;
        issue_one();
        if (PAIR && !(n & 1)) { acc_first = acc; continue; }
        float g = PAIR ? acc_first : acc, uu = acc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            g += __shfl_xor_sync(0xffffffffu, g, o);
            if (PAIR) uu += __shfl_xor_sync(0xffffffffu, uu, o);
        }
        if (lane == 0) {
            const int oidx = (pr & 0xff) * p.rows + crr;
            const long o = (long)(t0 + (pr >> 8)) * total_out + oidx;
            if (PAIR) {
                p.out_f32[o] = (p.use_silu ? act_silu(g) : act_relu(g)) * uu;
            } else {
                if (p.bias) g += p.bias[crr];
                if (p.out_f32) p.out_f32[o] = g;
                if (p.out_hidden) store_hidden(p.out_hidden, o, p.hidden_type, g);
            }
        }
        crr += W;
        while (crr >= p.rows) { crr -= p.rows; cpi++; }
    }
  }  // token chunks
}

// ---------------------------------------------------------------------------------------------------------------
// Item formats of the down-projection kernel.  An item is 4 consecutive rows x nb super-blocks, f = rw*nb + blk.
struct BulkQ4K {   // raw Q4_K rows: block f of the item at f*144
    static constexpr int kBlockBytes = SZ_Q4_K;
    static constexpr int kBs = 8;   // int16 activation sums per block (32-value groups)
    __device__ static __forceinline__ float dot(const uint8_t* sl, int f, int /*nrb*/, const uint8_t* aq, const int16_t* bs, float dxb) {
        return q4k_block_dot(sl + f * SZ_Q4_K, aq, *reinterpret_cast<const uint4*>(bs), dxb);
    }
};

struct BulkQ6K4T {   // chunk-major 4-row tiles (see the header comment)
    static constexpr int kBlockBytes = SZ_Q6_K;
    static constexpr int kBs = 16;  // int16 activation sums per block (16-value groups)
    __device__ static __forceinline__ float dot(const uint8_t* sl, int f, int nrb, const uint8_t* aq, const int16_t* bs, float dxb) {
        const uint8_t* ql = sl + f * 16;                  // chunk c at ql + c*nrb*16
        const uint8_t* qh = sl + nrb * 128 + f * 16;      // chunk c at qh + c*nrb*16
        const uint4 scv = *reinterpret_cast<const uint4*>(sl + nrb * 192 + f * 16);
        const float d = fp16_bits_to_f32(*reinterpret_cast<const uint16_t*>(sl + nrb * 208 + f * 2));
        const uint4 bs0 = *reinterpret_cast<const uint4*>(bs);
        const uint4 bs1 = *reinterpret_cast<const uint4*>(bs + 8);
        const uint32_t scw[4] = {scv.x, scv.y, scv.z, scv.w};
        const int cs = nrb * 16;
        int isum = 0;
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
            uint32_t a[8], b[8], h[8];
            *reinterpret_cast<uint4*>(a) = *reinterpret_cast<const uint4*>(ql + (4 * hh + 0) * cs);
            *reinterpret_cast<uint4*>(a + 4) = *reinterpret_cast<const uint4*>(ql + (4 * hh + 1) * cs);
            *reinterpret_cast<uint4*>(b) = *reinterpret_cast<const uint4*>(ql + (4 * hh + 2) * cs);
            *reinterpret_cast<uint4*>(b + 4) = *reinterpret_cast<const uint4*>(ql + (4 * hh + 3) * cs);
            *reinterpret_cast<uint4*>(h) = *reinterpret_cast<const uint4*>(qh + (2 * hh + 0) * cs);
            *reinterpret_cast<uint4*>(h + 4) = *reinterpret_cast<const uint4*>(qh + (2 * hh + 1) * cs);
            int s[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};   // [quarter i][l >= 16]
#pragma unroll
            for (int i = 0; i < 4; i++) {   // the four 32-value quarters of this 128-half
                uint32_t x[8];
                *reinterpret_cast<uint4*>(x) = *reinterpret_cast<const uint4*>(aq + 128 * hh + 32 * i);
                *reinterpret_cast<uint4*>(x + 4) = *reinterpret_cast<const uint4*>(aq + 128 * hh + 32 * i + 16);
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    // 6-bit value = 4 low bits from ql | 2 high bits from qh; 0..63 (the -32 is folded below)
                    uint32_t v;
                    if (i == 0) v = (a[w] & 0x0f0f0f0fu) | ((h[w] << 4) & 0x30303030u);
                    else if (i == 1) v = (b[w] & 0x0f0f0f0fu) | ((h[w] << 2) & 0x30303030u);
                    else if (i == 2) v = ((a[w] >> 4) & 0x0f0f0f0fu) | (h[w] & 0x30303030u);
                    else v = ((b[w] >> 4) & 0x0f0f0f0fu) | ((h[w] >> 2) & 0x30303030u);
                    s[i][w >> 2] = dp4a_s8s8(v, x[w], s[i][w >> 2]);
                }
            }
            // 16-value group g = 8*hh + 2*i + (l >= 16) carries scale byte g
            const uint32_t lo = scw[2 * hh], hi = scw[2 * hh + 1];
            isum += sext8(lo) * s[0][0] + sext8(lo >> 8) * s[0][1] + sext8(lo >> 16) * s[1][0] + sext8(lo >> 24) * s[1][1];
            isum += sext8(hi) * s[2][0] + sext8(hi >> 8) * s[2][1] + sext8(hi >> 16) * s[3][0] + sext8(hi >> 24) * s[3][1];
        }
        // sum (q-32) x = sum q x - 32 * sum_g sc_g * bsum_g   (dp2a: int16 bsums x int8 scales)
        int corr = __dp2a_lo((int)bs0.x, (int)scw[0], 0);
        corr = __dp2a_hi((int)bs0.y, (int)scw[0], corr);
        corr = __dp2a_lo((int)bs0.z, (int)scw[1], corr);
        corr = __dp2a_hi((int)bs0.w, (int)scw[1], corr);
        corr = __dp2a_lo((int)bs1.x, (int)scw[2], corr);
        corr = __dp2a_hi((int)bs1.y, (int)scw[2], corr);
        corr = __dp2a_lo((int)bs1.z, (int)scw[3], corr);
        corr = __dp2a_hi((int)bs1.w, (int)scw[3], corr);
        return (d * dxb) * (float)(isum - 32 * corr);
    }
};

// Down projection + weighted combine.  Work item of a warp = (pair, 4 consecutive output rows) = one bulk copy; every
// CTA owns a contiguous range of row quads for ALL pairs so the combine over experts stays in the CTA.
//
// Tokens are processed in chunks: as many consecutive tokens as have, together, at most `pcap` (token, slot) pairs
// owned by this launch (launcher: pcap >= slots + 1, so a chunk always holds at least one token).  Within a chunk all
// pairs form ONE work list (expert-parallel shards and decode batches do not drain the ring per token) and all their
// activation rows are staged as Q8_K side by side.
constexpr int kBulkMaxChunkTokens = 16;
template <class Fmt, int SLOTS>
__global__ void __launch_bounds__(kBulkMaxWarpsDown * 32, 1) reduce_bulk_kernel(const ReduceParams p, int nrows_max, int pcap) {
    constexpr int RW = 4;
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_np, s_nt;
    __shared__ int s_first[kBulkMaxChunkTokens + 1];   // first pair of every token of the chunk (+ end)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    int Teff = p.ntokens;
    if (p.bsz) Teff = min(Teff, *p.bsz);
    const int nb = p.ncols / QK_K;
    const int k = p.slots;
    const int ns = k + (p.xw ? 1 : 0);
    const int nrb = RW * nb;                                  // (row, block) pairs per item
    const int item_bytes = nrb * Fmt::kBlockBytes;
    // staging: q8 [pcap][nb][272] | bs [pcap][nb][kBs] int16 | dx [pcap][nb] | partial [nrows_max][pcap] | pair list [pcap] | mbarriers | rings
    uint8_t* q8 = smem;
    int16_t* bs = reinterpret_cast<int16_t*>(smem + (size_t)pcap * nb * kActBlkStride);
    float* dx = reinterpret_cast<float*>(smem + (size_t)pcap * nb * (kActBlkStride + 2 * Fmt::kBs));
    float* partial = dx + (size_t)pcap * nb;
    int* pairs = reinterpret_cast<int*>(partial + (size_t)nrows_max * pcap);   // (token in chunk) << 8 | slot
    size_t off = (size_t)pcap * nb * (kActBlkStride + 2 * Fmt::kBs + 4) + (size_t)nrows_max * pcap * 4 + (size_t)pcap * 4;
    off = (off + 15) & ~(size_t)15;
    const int bar_bytes = (W * SLOTS * 8 + 15) & ~15;
    const uint32_t bar_u32 = (uint32_t)__cvta_generic_to_shared(smem + off) + warp * SLOTS * 8;
    uint8_t* ring = smem + off + bar_bytes + (size_t)warp * SLOTS * item_bytes;
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < SLOTS; s++) mbar_init(bar_u32 + 8 * s, 1);
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    int slot_i = 0, slot_u = 0;
    uint32_t phase = 0;
    const int quads = p.rows / RW;
    const int q0 = (int)((long)quads * blockIdx.x / gridDim.x), q1 = (int)((long)quads * (blockIdx.x + 1) / gridDim.x);
    const int r0 = q0 * RW, nquads = q1 - q0, nrows = nquads * RW;

  for (int t0 = 0; t0 < Teff;) {
    __syncthreads();
    if (threadIdx.x == 0) {   // greedy chunk: tokens t0.. while their owned pairs fit
        int np = 0, nt = 0;
        while (t0 + nt < Teff && nt < kBulkMaxChunkTokens) {
            const bool sh_here = p.xw && (p.shared_token < 0 || p.shared_token == t0 + nt);
            int cnt = sh_here ? 1 : 0;
            for (int j = 0; j < k; j++) {
                const long e = p.ids ? (long)p.ids[(long)(t0 + nt) * k + j] - p.id_offset : 0;
                cnt += (e >= 0 && e < p.n_experts) ? 1 : 0;
            }
            if (nt > 0 && np + cnt > pcap) break;
            s_first[nt] = np;
            for (int j = 0; j < k; j++) {
                const long e = p.ids ? (long)p.ids[(long)(t0 + nt) * k + j] - p.id_offset : 0;
                if (e >= 0 && e < p.n_experts) pairs[np++] = (nt << 8) | j;
            }
            if (sh_here) pairs[np++] = (nt << 8) | k;
            nt++;
        }
        s_first[nt] = np;
        s_np = np;
        s_nt = nt;
    }
    __syncthreads();
    const int np = s_np, nt = s_nt;
    const int total = nquads * np;   // item = pair * nquads + quad
    int ni = total - warp;
    ni = ni > 0 ? (ni + W - 1) / W : 0;
    int ipi = 0, iq = 0, iss = 0;
    if (ni > 0) { ipi = warp / nquads; iq = warp - ipi * nquads; }
    int cpi = ipi, cq = iq;

    auto issue_one = [&]() {
        if (iss < ni) {
            if (lane == 0) {
                const int pr = pairs[ipi], j = pr & 0xff;
                long row = r0 + iq * RW;
                const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.w);
                if (j == k) wbase = reinterpret_cast<const uint8_t*>(p.xw);
                else row += (p.ids ? (long)p.ids[(long)(t0 + (pr >> 8)) * k + j] - p.id_offset : 0L) * p.rows;
                const uint8_t* src = wbase + (row >> 2) * item_bytes;
                const uint32_t bar = bar_u32 + 8 * slot_i;
                mbar_expect_tx(bar, (uint32_t)item_bytes);
                bulk_g2s(ring_u32 + slot_i * item_bytes, src, (uint32_t)item_bytes, bar);
            }
            iss++;
            iq += W;
            while (iq >= nquads) { iq -= nquads; ipi++; }
            slot_i = (slot_i + 1 == SLOTS) ? 0 : slot_i + 1;
        }
    };
#pragma unroll
    for (int s = 0; s < SLOTS; s++) issue_one();

    {   // quantise the pairs' activation rows (fp32 phase-1 output) into the padded layout: block g = (pair, block)
        float cur[8], nxt[8];
        const int totalb = np * nb;
        auto src_of = [&](int g) -> long {
            const int pi = g / nb, b = g - pi * nb, pr = pairs[pi];
            return ((long)(t0 + (pr >> 8)) * ns + (pr & 0xff)) * p.ncols + (long)b * QK_K + lane * 8;
        };
        int g = warp;
        if (g < totalb) load_block8(p.a, src_of(g), KTB200_TYPE_F32, cur);
#pragma unroll 1
        while (g < totalb) {
            const int gn = g + W;
            if (gn < totalb) load_block8(p.a, src_of(gn), KTB200_TYPE_F32, nxt);
            warp_quantize_q8k_block(cur, lane, reinterpret_cast<uint32_t*>(q8 + (size_t)g * kActBlkStride), dx + g,
                                    Fmt::kBs == 16 ? bs + g * 16 : nullptr, Fmt::kBs == 8 ? bs + g * 8 : nullptr);
#pragma unroll
            for (int i = 0; i < 8; i++) cur[i] = nxt[i];
            g = gn;
        }
    }
    __syncthreads();

    for (int n = 0; n < ni; n++) {
        mbar_wait(bar_u32 + 8 * slot_u, (phase >> slot_u) & 1u);
        phase ^= 1u << slot_u;
        const uint8_t* sl = ring + slot_u * item_bytes;
        float res;
        {
            float acc[RW] = {0.f, 0.f, 0.f, 0.f};
            for (int f = lane; f < nrb; f += 32) {   // (row, block) pairs of the tile; 4 x 8 = one per lane for I = 2048
                const int rw = f / nb, blk = f - rw * nb;
                const int ab = cpi * nb + blk;
                const float val = Fmt::dot(sl, f, nrb, q8 + (size_t)ab * kActBlkStride, bs + ab * Fmt::kBs, dx[ab]);
                acc[0] += rw == 0 ? val : 0.f; acc[1] += rw == 1 ? val : 0.f; acc[2] += rw == 2 ? val : 0.f; acc[3] += rw == 3 ? val : 0.f;
            }
            res = warp_reduce4(acc[0], acc[1], acc[2], acc[3], lane);
        }
        __syncwarp();
        slot_u = (slot_u + 1 == SLOTS) ? 0 : slot_u + 1;
        issue_one();
        if ((lane & 7) == 0) partial[(cq * RW + (lane >> 3)) * pcap + cpi] = res;
        cq += W;
        while (cq >= nquads) { cq -= nquads; cpi++; }
    }
    __syncthreads();
    // weighted accumulation over a token's experts IN expert_ids ORDER (moe.cpp:222-236), one FMA per expert
    for (int idx = threadIdx.x; idx < nrows * nt; idx += W * 32) {
        const int tl = idx / nrows, hl = idx - tl * nrows;
        const long t = t0 + tl;
        float acc = 0.f, shared = 0.f;
        for (int pi = s_first[tl]; pi < s_first[tl + 1]; pi++) {
            const int j = pairs[pi] & 0xff;
            const float dv = partial[hl * pcap + pi];
            if (j == k) shared = dv;
            else acc = p.weights ? __fmaf_rn(dv, p.weights[t * k + j], acc) : acc + dv;
        }
        const long o = t * p.rows + r0 + hl;
        bool has_sh = false;
        for (int pi = s_first[tl]; pi < s_first[tl + 1]; pi++) has_sh |= (pairs[pi] & 0xff) == k;
        if (p.xw_out) { if (has_sh) store_hidden(p.xw_out, o, p.xw_out_type, shared); }
        else if (p.xw) acc = round_hidden(acc, p.hidden_type) + round_hidden(shared, p.hidden_type);
        if (p.accumulate) acc = load_hidden(p.out, o, p.hidden_type) + round_hidden(acc, p.hidden_type);
        store_hidden(p.out, o, p.hidden_type, acc);
    }
    t0 += nt;
  }  // token chunks
}

}  // namespace ktb
