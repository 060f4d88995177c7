// Software-pipelined variant of rows_kernel for long raw-block rows (Q4_K / Q5_K gate+up of a 7168-wide model).
//
// ncu on rows_kernel shows every warp parked on long-scoreboard (global-load) stalls: the bytes a warp can keep in
// flight are bounded by the registers that hold them.  Here the weight rows travel global -> shared memory with
// cp.async (LDGSTS: no register staging): every warp owns a private 2-slot ring, streams the NEXT unit (one gate
// row + one up row, 8 KB) into one slot while it computes on the other, and reads weights back with conflict-free
// LDS.128.  No cross-warp synchronisation in the steady state; 12 warps x 8 KB = 96 KB in flight per SM instead
// of ~32 KB.  Arithmetic, lane mapping and epilogue are those of rows_kernel (same Fmt::dot).
#pragma once
#include "gemv.cuh"

namespace ktb {

__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <class Fmt, bool PAIR, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) rows_pipe_kernel(const RowsParams p, int act_bytes, int slot_bytes) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int t = blockIdx.y;
    if (p.bsz && t >= *p.bsz) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nblk = p.ncols / QK_K;
    const int row_bytes = nblk * Fmt::kBlockBytes;
    constexpr int NM = PAIR ? 2 : 1;
    uint8_t* q8 = smem;
    float* dx = reinterpret_cast<float*>(smem + p.ncols);
    int16_t* bsums = reinterpret_cast<int16_t*>(smem + p.ncols + ((nblk * 4 + 15) & ~15));   // 16-byte aligned
    uint8_t* ring = smem + act_bytes + (size_t)warp * 2 * slot_bytes;
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
    const ActQ8K act{q8, dx, bsums};
    const typename Fmt::Lane L = Fmt::lane(lane);

    const int nslots = p.slots + (p.x0 ? 1 : 0);
    const int total = nslots * p.rows;
    const int u0 = (int)((long)total * blockIdx.x / gridDim.x), u1 = (int)((long)total * (blockIdx.x + 1) / gridDim.x);

    auto unit_rows = [&](int u, const uint8_t* (&r)[NM]) -> bool {
        if (u >= u1) return false;
        const int s = u / p.rows, rr = u - s * p.rows;
        if (s == p.slots) {
            r[0] = reinterpret_cast<const uint8_t*>(p.x0) + (long)rr * row_bytes;
            if (PAIR) r[NM - 1] = reinterpret_cast<const uint8_t*>(p.x1) + (long)rr * row_bytes;
            return true;
        }
        const long e = p.ids ? (long)p.ids[(long)t * p.slots + s] - p.id_offset : 0;
        if (e < 0 || e >= p.n_experts) return false;
        r[0] = reinterpret_cast<const uint8_t*>(p.w0) + (e * p.rows + rr) * row_bytes;
        if (PAIR) r[NM - 1] = reinterpret_cast<const uint8_t*>(p.w1) + (e * p.rows + rr) * row_bytes;
        return true;
    };
    // stream one unit into ring slot `slot`; always commits a group so the per-thread group count stays in step
    auto issue = [&](int u, int slot) -> bool {
        const uint8_t* r[NM];
        const bool ok = unit_rows(u, r);
        if (ok) {
#pragma unroll
            for (int m = 0; m < NM; m++) {
                const uint32_t dst = ring_u32 + slot * slot_bytes + m * row_bytes;
                for (int c = lane * 16; c < row_bytes; c += 32 * 16) cp_async16_cg(dst + c, r[m] + c);
            }
        }
        cp_async_commit_group();
        return ok;
    };

    int u = u0 + warp;
    bool cur_ok = issue(u, 0);   // first unit streams in while the activation row is quantised
    cta_quantize_q8k_rows<4>(p.x, (long)t * p.ncols, 0, p.hidden_type, 1, p.ncols, 0u, q8, dx, bsums);
    __syncthreads();

    const int nsteps = (nblk + Fmt::kBlocksPerStep - 1) / Fmt::kBlocksPerStep;
    for (int it = 0; u < u1; u += WARPS, it++) {
        const int slot = it & 1;
        const bool next_ok = issue(u + WARPS, slot ^ 1);
        cp_async_wait_group<1>();   // everything but the group just committed has landed (this thread's copies)
        __syncwarp();               // ... and the other lanes' copies of this slot are visible
        float acc[NM];
#pragma unroll
        for (int m = 0; m < NM; m++) acc[m] = 0.f;
        if (cur_ok) {
            const uint8_t* row0 = ring + slot * slot_bytes;
#pragma unroll 2
            for (int s = 0; s < nsteps; s++) {
                const int blk = s * Fmt::kBlocksPerStep + L.blk;
                if (blk < nblk) {
                    typename Fmt::Regs R[NM];
#pragma unroll
                    for (int m = 0; m < NM; m++) Fmt::load_smem(row0 + m * row_bytes, blk, L, R[m]);
                    typename Fmt::Act A;
                    Fmt::load_act(act, blk, L, A);
#pragma unroll
                    for (int m = 0; m < NM; m++) acc[m] += Fmt::dot(R[m], A, L);
                }
            }
        }
        float g = warp_sum(acc[0]);
        const float uu = PAIR ? warp_sum(acc[NM - 1]) : 0.f;
        if (lane == 0) {
            if (PAIR) {
                p.out_f32[(long)t * total + u] = cur_ok ? (p.use_silu ? act_silu(g) : act_relu(g)) * uu : 0.f;
            } else {
                if (!cur_ok) g = 0.f;
                if (p.bias) g += p.bias[u % p.rows];
                if (p.out_f32) p.out_f32[(long)t * total + u] = g;
                if (p.out_hidden) store_hidden(p.out_hidden, (long)t * total + u, p.hidden_type, g);
            }
        }
        __syncwarp();               // all lanes are done with `slot` before the next iteration refills it
        cur_ok = next_ok;
    }
    cp_async_wait_group<0>();
}

}  // namespace ktb

namespace ktb {

// Pipelined reduce_kernel for Q6_K (8-row SoA) down projections.  Work item = (slot j, 4 consecutive rows):
// in the SoA layout the item is four contiguous pieces (ql 4x128nb | qh 4x64nb | scales 4x16nb | d 4x2nb),
// streamed into a warp-private 2-slot ring with cp.async while the previous item is reduced from shared memory.
// CTA row ranges are multiples of 4 rows; requires rows % 4 == 0 and nb even.
template <int WARPS, int SLOTS>
__global__ void __launch_bounds__(WARPS * 32, 1) reduce_pipe_q6k8_kernel(const ReduceParams p, int slot_bytes) {
    using Fmt = FmtQ6K8;
    constexpr int RW = 4;
    extern __shared__ __align__(16) uint8_t smem[];
    const int t = blockIdx.y;
    if (p.bsz && t >= *p.bsz) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = p.ncols / QK_K;
    const int k = p.slots;
    const int ns = k + (p.xw ? 1 : 0);
    // staging: q8 [ns][ncols] | dx [ns][nb] | bsums [ns][ncols/16] | partial [rows_local][ns] | ring
    uint8_t* q8 = smem;
    float* dx = reinterpret_cast<float*>(smem + (size_t)ns * p.ncols);
    int16_t* bsums = reinterpret_cast<int16_t*>(smem + (size_t)ns * p.ncols + (size_t)ns * nb * 4);
    float* partial = reinterpret_cast<float*>(smem + (size_t)ns * p.ncols + (size_t)ns * nb * 4 + (size_t)ns * (p.ncols / 16) * 2);
    const int quads = p.rows / RW;
    const int q0 = (int)((long)quads * blockIdx.x / gridDim.x), q1 = (int)((long)quads * (blockIdx.x + 1) / gridDim.x);
    const int r0 = q0 * RW, nrows = (q1 - q0) * RW;
    const int nquads = q1 - q0;
    size_t off = (size_t)ns * p.ncols + (size_t)ns * nb * 4 + (size_t)ns * (p.ncols / 16) * 2 + (size_t)(nrows > 0 ? nrows : 1) * ns * 4;
    off = (off + 15) & ~(size_t)15;
    uint8_t* ring = smem + off + (size_t)warp * SLOTS * slot_bytes;
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);

    unsigned skip = 0;
    for (int j = 0; j < k; j++) {
        const long e = p.ids ? (long)p.ids[(long)t * k + j] - p.id_offset : 0;
        if (e < 0 || e >= p.n_experts) skip |= 1u << j;
    }
    const int total = nquads * ns;   // item = j * nquads + quad
    const int s_ql = 4 * 128 * nb, s_qh = 4 * 64 * nb, s_sc = 4 * 16 * nb, s_d = 4 * 2 * nb;

    auto issue = [&](int item, int slot) -> bool {
        bool ok = item < total;
        if (ok) {
            const int j = item / nquads, quad = item - j * nquads;
            const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.w);
            long row = r0 + quad * RW;
            if (j == k) wbase = reinterpret_cast<const uint8_t*>(p.xw);
            else if ((skip >> j) & 1u) ok = false;
            else row += (p.ids ? (long)p.ids[(long)t * k + j] - p.id_offset : 0L) * p.rows;
            if (ok) {
                const long G = row >> 3, r8 = row & 7;
                const uint8_t* g = wbase + G * (8 * SZ_Q6_K) * nb;
                const uint32_t dst = ring_u32 + slot * slot_bytes;
                const uint8_t* src[4] = {g + r8 * 128 * nb, g + 1024L * nb + r8 * 64 * nb, g + 1536L * nb + r8 * 16 * nb, g + 1664L * nb + r8 * 2 * nb};
                const int len[4] = {s_ql, s_qh, s_sc, s_d};
                int o = 0;
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    for (int c = lane * 16; c < len[s]; c += 32 * 16) cp_async16_cg(dst + o + c, src[s] + c);
                    o += len[s];
                }
            }
        }
        cp_async_commit_group();
        return ok;
    };

    int item = warp;
    bool cur_ok = issue(item, 0);   // weights do not depend on phase 1: stream the first item during the prologue
    cta_quantize_q8k_rows<4>(p.a, (long)t * ns * p.ncols, p.ncols, KTB200_TYPE_F32, ns, p.ncols, skip, q8, dx, bsums);
    __syncthreads();

    const Fmt::Lane L = Fmt::lane(lane);
    const int nsteps = (nb + Fmt::kBlocksPerStep - 1) / Fmt::kBlocksPerStep;
    for (int it = 0; item < total; item += WARPS, it++) {
        const int slot = (SLOTS == 2) ? (it & 1) : 0;
        bool next_ok = false;
        if (SLOTS == 2) {
            next_ok = issue(item + WARPS, slot ^ 1);
            cp_async_wait_group<1>();
        } else {
            cp_async_wait_group<0>();
        }
        __syncwarp();
        const int j = item / nquads, quad = item - j * nquads;
        float res = 0.f;
        if (cur_ok) {
            const uint8_t* sl = ring + slot * slot_bytes;
            Fmt::Row rp[RW];
#pragma unroll
            for (int rw = 0; rw < RW; rw++)
                rp[rw] = Fmt::Row{sl + rw * 128 * nb, sl + s_ql + rw * 64 * nb, sl + s_ql + s_qh + rw * 16 * nb, sl + s_ql + s_qh + s_sc + rw * 2 * nb};
            const ActQ8K act{q8 + (size_t)j * p.ncols, dx + j * nb, bsums + j * (p.ncols / 16)};
            float acc[RW] = {0.f, 0.f, 0.f, 0.f};
            for (int s = 0; s < nsteps; s++) {
                const int blk = s * Fmt::kBlocksPerStep + L.blk;
                if (blk < nb) {
                    Fmt::Act A;
                    Fmt::load_act(act, blk, L, A);
#pragma unroll
                    for (int rw = 0; rw < RW; rw++) {
                        Fmt::Regs R;
                        Fmt::load_smem(rp[rw], blk, L, R);
                        acc[rw] += Fmt::dot(R, A, L);
                    }
                }
            }
            res = warp_reduce4(acc[0], acc[1], acc[2], acc[3], lane);
        }
        if ((lane & 7) == 0) partial[(quad * RW + (lane >> 3)) * ns + j] = res;
        __syncwarp();
        if (SLOTS == 2) cur_ok = next_ok;
        else cur_ok = issue(item + WARPS, 0);
    }
    cp_async_wait_group<0>();
    __syncthreads();
    for (int hl = threadIdx.x; hl < nrows; hl += WARPS * 32) {
        float acc = 0.f;
        for (int j = 0; j < k; j++) {
            if ((skip >> j) & 1u) continue;
            const float d = partial[hl * ns + j];
            acc = p.weights ? __fmaf_rn(d, p.weights[(long)t * k + j], acc) : acc + d;
        }
        const long o = (long)t * p.rows + r0 + hl;
        if (p.xw) acc = round_hidden(acc, p.hidden_type) + round_hidden(partial[hl * ns + k], p.hidden_type);
        if (p.accumulate) acc = load_hidden(p.out, o, p.hidden_type) + round_hidden(acc, p.hidden_type);
        store_hidden(p.out, o, p.hidden_type, acc);
    }
}

}  // namespace ktb

namespace ktb {

// Q4_K gate/up, third generation: rows staged in shared memory (same warp-private cp.async ring as rows_pipe_kernel)
// and consumed with ONE LANE PER SUPER-BLOCK.  A lane decodes its block's eight 6-bit (scale, min) pairs once
// (the ggml kmask trick, ggml-quants.c:7494-7499), runs the 64 dp4a of the block against the int8 activations,
// applies the sub-block scales as integers, folds the mins with 4 dp2a against pre-summed activations and touches
// fp32 once per block — the same granularity as the reference's scalar loop (one `d * sumi` per block) and ~2.5x
// fewer instructions than one lane per 16-byte chunk.  Bank-conflict-free by construction: weight blocks are 144 B
// (36 words) apart, activation blocks are padded to 272 B (68 words), so 8 lanes x LDS.128 hit 32 distinct banks.
constexpr int kActBlkStride = QK_K + 16;   // padded int8 activation block

// SLOTS = 2: every warp prefetches its next unit while it computes (12 warps fit); SLOTS = 1: no intra-warp overlap
// but twice the warps (24) — more eligible warps per scheduler, the in-flight bytes come from the warps that wait.
template <bool PAIR, int WARPS, int SLOTS>
__global__ void __launch_bounds__(WARPS * 32, 1) rows_pipe_q4k_blk_kernel(const RowsParams p, int act_bytes, int slot_bytes) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_vs[36];   // compacted list of the slots this launch actually computes for the current token
    __shared__ int s_nv;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int Teff = p.ntokens;
    if (p.bsz) Teff = min(Teff, *p.bsz);
    const int nblk = p.ncols / QK_K;
    const int row_bytes = nblk * SZ_Q4_K;
    constexpr int NM = PAIR ? 2 : 1;
    // activation staging: q8 [nblk][272] | bs32 [nblk][8] int16 | dx [nblk] float
    uint8_t* q8 = smem;
    int16_t* bs32 = reinterpret_cast<int16_t*>(smem + (size_t)nblk * kActBlkStride);
    float* dx = reinterpret_cast<float*>(smem + (size_t)nblk * kActBlkStride + (size_t)nblk * 16);
    uint8_t* ring = smem + act_bytes + (size_t)warp * SLOTS * slot_bytes;
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);

    const int nslots = p.slots + (p.x0 ? 1 : 0);
    const int total_out = nslots * p.rows;
  // One CTA per SM walks the tokens; per token the work list holds only the slots this shard owns (expert-parallel
  // shards skip most ids: a static split over all k slots would leave most CTAs idle).
  for (int t = 0; t < Teff; t++) {
    __syncthreads();   // previous token: everyone is done with the activation staging and s_vs
    if (threadIdx.x == 0) {
        int nv = 0;
        for (int s = 0; s < p.slots; s++) {
            const long e = p.ids ? (long)p.ids[(long)t * p.slots + s] - p.id_offset : 0;
            if (e >= 0 && e < p.n_experts) s_vs[nv++] = s;
        }
        if (p.x0) s_vs[nv++] = p.slots;
        s_nv = nv;
    }
    __syncthreads();
    const int total = s_nv * p.rows;
    const int u0 = (int)((long)total * blockIdx.x / gridDim.x), u1 = (int)((long)total * (blockIdx.x + 1) / gridDim.x);

    // unit u -> (real slot, row): returns the output index or -1
    auto unit_rows = [&](int u, const uint8_t* (&r)[NM]) -> int {
        if (u >= u1) return -1;
        const int vi = u / p.rows, rr = u - vi * p.rows;
        const int s = s_vs[vi];
        if (s == p.slots) {
            r[0] = reinterpret_cast<const uint8_t*>(p.x0) + (long)rr * row_bytes;
            if (PAIR) r[NM - 1] = reinterpret_cast<const uint8_t*>(p.x1) + (long)rr * row_bytes;
        } else {
            const long e = p.ids ? (long)p.ids[(long)t * p.slots + s] - p.id_offset : 0;
            r[0] = reinterpret_cast<const uint8_t*>(p.w0) + (e * p.rows + rr) * row_bytes;
            if (PAIR) r[NM - 1] = reinterpret_cast<const uint8_t*>(p.w1) + (e * p.rows + rr) * row_bytes;
        }
        return s * p.rows + rr;
    };
    auto issue = [&](int u, int slot) -> int {
        const uint8_t* r[NM];
        const int ok = unit_rows(u, r);
        if (ok >= 0) {
#pragma unroll
            for (int m = 0; m < NM; m++) {
                const uint32_t dst = ring_u32 + slot * slot_bytes + m * row_bytes;
                for (int c = lane * 16; c < row_bytes; c += 32 * 16) cp_async16_cg(dst + c, r[m] + c);
            }
        }
        cp_async_commit_group();
        return ok;
    };

    int u = u0 + warp;
    int cur_ok = issue(u, 0);
    {   // quantise the token's activation row into the padded layout (one warp per block, 4 blocks in flight)
        const int nwarps = WARPS;
        for (int g0 = warp; g0 < nblk; g0 += nwarps * 4) {
            float x[4][8];
            bool live[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int b = g0 + i * nwarps;
                live[i] = b < nblk;
                if (live[i]) load_block8(p.x, (long)t * p.ncols + (long)b * QK_K + lane * 8, p.hidden_type, x[i]);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int b = g0 + i * nwarps;
                if (live[i]) warp_quantize_q8k_block(x[i], lane, reinterpret_cast<uint32_t*>(q8 + (size_t)b * kActBlkStride), dx + b, nullptr, bs32 + b * 8);
            }
        }
    }
    __syncthreads();

    for (int it = 0; u < u1; u += WARPS, it++) {
        const int slot = (SLOTS == 2) ? (it & 1) : 0;
        int next_ok = -1;
        if (SLOTS == 2) {
            next_ok = issue(u + WARPS, slot ^ 1);
            cp_async_wait_group<1>();
        } else {
            cp_async_wait_group<0>();
        }
        __syncwarp();
        float acc[NM];
#pragma unroll
        for (int m = 0; m < NM; m++) acc[m] = 0.f;
        if (cur_ok >= 0) {
            const uint8_t* row0 = ring + slot * slot_bytes;
            for (int blk = lane; blk < nblk; blk += 32) {
                const uint8_t* aq = q8 + (size_t)blk * kActBlkStride;
                const uint4 bsv = *reinterpret_cast<const uint4*>(bs32 + blk * 8);
                const float dxb = dx[blk];
                uint32_t scl[NM], sch[NM], mnl[NM], mnh[NM];
                float2 dm[NM];
                int isum[NM];
#pragma unroll
                for (int m = 0; m < NM; m++) {
                    const uint4 hdr = *reinterpret_cast<const uint4*>(row0 + m * row_bytes + blk * SZ_Q4_K);
                    dm[m] = __half22float2(*reinterpret_cast<const __half2*>(&hdr.x));
                    scl[m] = hdr.y & 0x3f3f3f3fu;                                             // scales 0..3
                    mnl[m] = hdr.z & 0x3f3f3f3fu;                                             // mins   0..3
                    sch[m] = (hdr.w & 0x0f0f0f0fu) | ((hdr.y >> 2) & 0x30303030u);            // scales 4..7
                    mnh[m] = ((hdr.w >> 4) & 0x0f0f0f0fu) | ((hdr.z >> 2) & 0x30303030u);     // mins   4..7
                    isum[m] = 0;
                }
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const uint4 a0 = *reinterpret_cast<const uint4*>(aq + 64 * g);
                    const uint4 a1 = *reinterpret_cast<const uint4*>(aq + 64 * g + 16);
                    const uint4 a2 = *reinterpret_cast<const uint4*>(aq + 64 * g + 32);
                    const uint4 a3 = *reinterpret_cast<const uint4*>(aq + 64 * g + 48);
#pragma unroll
                    for (int m = 0; m < NM; m++) {
                        const uint8_t* qs = row0 + m * row_bytes + blk * SZ_Q4_K + 16 + 32 * g;
                        const uint4 q0 = *reinterpret_cast<const uint4*>(qs);
                        const uint4 q1 = *reinterpret_cast<const uint4*>(qs + 16);
                        int slo = 0, shi = 0;
                        slo = dp4a_s8s8(q0.x & 0x0f0f0f0fu, a0.x, slo); slo = dp4a_s8s8(q0.y & 0x0f0f0f0fu, a0.y, slo);
                        slo = dp4a_s8s8(q0.z & 0x0f0f0f0fu, a0.z, slo); slo = dp4a_s8s8(q0.w & 0x0f0f0f0fu, a0.w, slo);
                        slo = dp4a_s8s8(q1.x & 0x0f0f0f0fu, a1.x, slo); slo = dp4a_s8s8(q1.y & 0x0f0f0f0fu, a1.y, slo);
                        slo = dp4a_s8s8(q1.z & 0x0f0f0f0fu, a1.z, slo); slo = dp4a_s8s8(q1.w & 0x0f0f0f0fu, a1.w, slo);
                        shi = dp4a_u8s8(q0.x & 0xf0f0f0f0u, a2.x, shi); shi = dp4a_u8s8(q0.y & 0xf0f0f0f0u, a2.y, shi);
                        shi = dp4a_u8s8(q0.z & 0xf0f0f0f0u, a2.z, shi); shi = dp4a_u8s8(q0.w & 0xf0f0f0f0u, a2.w, shi);
                        shi = dp4a_u8s8(q1.x & 0xf0f0f0f0u, a3.x, shi); shi = dp4a_u8s8(q1.y & 0xf0f0f0f0u, a3.y, shi);
                        shi = dp4a_u8s8(q1.z & 0xf0f0f0f0u, a3.z, shi); shi = dp4a_u8s8(q1.w & 0xf0f0f0f0u, a3.w, shi);
                        const uint32_t scw = (g < 2) ? scl[m] : sch[m];
                        const int sc0 = (int)((scw >> (16 * (g & 1))) & 0xff), sc1 = (int)((scw >> (16 * (g & 1) + 8)) & 0xff);
                        isum[m] += sc0 * slo + sc1 * (shi >> 4);
                    }
                }
#pragma unroll
                for (int m = 0; m < NM; m++) {
                    int msum = __dp2a_lo((int)bsv.x, (int)mnl[m], 0);
                    msum = __dp2a_hi((int)bsv.y, (int)mnl[m], msum);
                    msum = __dp2a_lo((int)bsv.z, (int)mnh[m], msum);
                    msum = __dp2a_hi((int)bsv.w, (int)mnh[m], msum);
                    acc[m] += (dm[m].x * dxb) * (float)isum[m] - (dm[m].y * dxb) * (float)msum;
                }
            }
        }
        float g = warp_sum(acc[0]);
        const float uu = PAIR ? warp_sum(acc[NM - 1]) : 0.f;
        if (lane == 0 && cur_ok >= 0) {
            const long o = (long)t * total_out + cur_ok;
            if (PAIR) {
                p.out_f32[o] = (p.use_silu ? act_silu(g) : act_relu(g)) * uu;
            } else {
                if (p.bias) g += p.bias[cur_ok % p.rows];
                if (p.out_f32) p.out_f32[o] = g;
                if (p.out_hidden) store_hidden(p.out_hidden, o, p.hidden_type, g);
            }
        }
        __syncwarp();
        if (SLOTS == 2) cur_ok = next_ok;
        else cur_ok = issue(u + WARPS, 0);   // refill the single slot for the next round
    }
    cp_async_wait_group<0>();
  }  // tokens
}

}  // namespace ktb

namespace ktb {

// Q6_K down projection, one lane per super-block from the warp-private ring.  Work item = (slot j, 4 consecutive
// rows) as in reduce_pipe_q6k8_kernel; the copy scatters each block into a padded shared-memory layout
// (ql 128+16 B, qh 64+16 B) so that 8 lanes x LDS.128 never collide, and for nb = 8 (I = 2048) the 32 lanes are
// exactly 4 rows x 8 blocks.  Per block: 64 dp4a, the 16 int8 sub-scales applied as integers, the "-32" offset
// folded with 8 dp2a against the activation bsums, fp32 touched once.
constexpr int kQ6QlStride = 144, kQ6QhStride = 80;

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) reduce_pipe_q6k_blk_kernel(const ReduceParams p, int slot_bytes) {
    constexpr int RW = 4;
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_vs[36];
    __shared__ int s_nv;
    __shared__ unsigned s_skip;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int Teff = p.ntokens;
    if (p.bsz) Teff = min(Teff, *p.bsz);
    const int nb = p.ncols / QK_K;
    const int k = p.slots;
    const int ns = k + (p.xw ? 1 : 0);
    // staging: q8 [ns][nb][272] | bsums [ns][nb][16] int16 | dx [ns][nb] | partial [rows_local][ns] | ring
    uint8_t* q8 = smem;
    int16_t* bsums = reinterpret_cast<int16_t*>(smem + (size_t)ns * nb * kActBlkStride);
    float* dx = reinterpret_cast<float*>(smem + (size_t)ns * nb * kActBlkStride + (size_t)ns * nb * 32);
    float* partial = dx + (size_t)ns * nb;
    const int quads = p.rows / RW;
    const int q0 = (int)((long)quads * blockIdx.x / gridDim.x), q1 = (int)((long)quads * (blockIdx.x + 1) / gridDim.x);
    const int r0 = q0 * RW, nquads = q1 - q0, nrows = nquads * RW;
    size_t off = (size_t)ns * nb * (kActBlkStride + 32 + 4) + (size_t)(nrows > 0 ? nrows : 1) * ns * 4;
    off = (off + 15) & ~(size_t)15;
    uint8_t* ring = smem + off + (size_t)warp * 2 * slot_bytes;
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
    const int nrb = RW * nb;                                 // (row, block) pairs per item
    const int o_qh = nrb * kQ6QlStride, o_sc = o_qh + nrb * kQ6QhStride, o_d = o_sc + nrb * 16;

  for (int t = 0; t < Teff; t++) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned sk = 0;
        int nv = 0;
        for (int j = 0; j < k; j++) {
            const long e = p.ids ? (long)p.ids[(long)t * k + j] - p.id_offset : 0;
            if (e < 0 || e >= p.n_experts) sk |= 1u << j; else s_vs[nv++] = j;
        }
        if (p.xw) s_vs[nv++] = k;
        s_nv = nv;
        s_skip = sk;
    }
    __syncthreads();
    const unsigned skip = s_skip;
    const int total = nquads * s_nv;   // item = vi * nquads + quad over the VALID slots only

    auto issue = [&](int item, int slot) -> bool {
        bool ok = item < total;
        if (ok) {
            const int vi = item / nquads, quad = item - vi * nquads;
            const int j = s_vs[vi];
            const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.w);
            long row = r0 + quad * RW;
            if (j == k) wbase = reinterpret_cast<const uint8_t*>(p.xw);
            else row += (p.ids ? (long)p.ids[(long)t * k + j] - p.id_offset : 0L) * p.rows;
            {
                const long G = row >> 3, r8 = row & 7;
                const uint8_t* g = wbase + G * (8 * SZ_Q6_K) * nb;
                const uint32_t dst = ring_u32 + slot * slot_bytes;
                const uint8_t* s_ql = g + r8 * 128 * nb;
                const uint8_t* s_qh = g + 1024L * nb + r8 * 64 * nb;
                const uint8_t* s_sc = g + 1536L * nb + r8 * 16 * nb;
                const uint8_t* s_d = g + 1664L * nb + r8 * 2 * nb;
                for (int c = lane; c < nrb * 8; c += 32) cp_async16_cg(dst + (c >> 3) * kQ6QlStride + (c & 7) * 16, s_ql + c * 16);
                for (int c = lane; c < nrb * 4; c += 32) cp_async16_cg(dst + o_qh + (c >> 2) * kQ6QhStride + (c & 3) * 16, s_qh + c * 16);
                for (int c = lane; c < nrb; c += 32) cp_async16_cg(dst + o_sc + c * 16, s_sc + c * 16);
                for (int c = lane; c < nrb / 8; c += 32) cp_async16_cg(dst + o_d + c * 16, s_d + c * 16);
            }
        }
        cp_async_commit_group();
        return ok;
    };

    int item = warp;
    bool cur_ok = issue(item, 0);
    {   // quantise the ns activation rows (fp32 phase-1 output) into the padded layout
        const int totalb = ns * nb;
        for (int g0 = warp; g0 < totalb; g0 += WARPS * 4) {
            float x[4][8];
            bool live[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int gb = g0 + i * WARPS;
                live[i] = gb < totalb;
                if (live[i]) {
                    const int r = gb / nb, b = gb - r * nb;
                    live[i] = !((skip >> r) & 1u);
                    if (live[i]) load_block8(p.a, ((long)t * ns + r) * p.ncols + (long)b * QK_K + lane * 8, KTB200_TYPE_F32, x[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int gb = g0 + i * WARPS;
                if (live[i]) warp_quantize_q8k_block(x[i], lane, reinterpret_cast<uint32_t*>(q8 + (size_t)gb * kActBlkStride), dx + gb, bsums + gb * 16);
            }
        }
    }
    __syncthreads();

    for (int it = 0; item < total; item += WARPS, it++) {
        const int slot = it & 1;
        const bool next_ok = issue(item + WARPS, slot ^ 1);
        cp_async_wait_group<1>();
        __syncwarp();
        const int vi = item / nquads, quad = item - vi * nquads;
        const int j = s_vs[vi];
        float acc[RW] = {0.f, 0.f, 0.f, 0.f};
        if (cur_ok) {
            const uint8_t* sl = ring + slot * slot_bytes;
            for (int f = lane; f < nrb; f += 32) {
                const int rw = f / nb, blk = f - rw * nb;
                const uint8_t* ql = sl + f * kQ6QlStride;
                const uint8_t* qh = sl + o_qh + f * kQ6QhStride;
                const uint4 scv = *reinterpret_cast<const uint4*>(sl + o_sc + f * 16);
                const float d = fp16_bits_to_f32(*reinterpret_cast<const uint16_t*>(sl + o_d + f * 2));
                const int ab = j * nb + blk;
                const uint8_t* aq = q8 + (size_t)ab * kActBlkStride;
                const uint4 bs0 = *reinterpret_cast<const uint4*>(bsums + ab * 16);
                const uint4 bs1 = *reinterpret_cast<const uint4*>(bsums + ab * 16 + 8);
                const uint32_t scw[4] = {scv.x, scv.y, scv.z, scv.w};
                int isum = 0;
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    uint32_t a[8], b[8], h[8];
                    *reinterpret_cast<uint4*>(a) = *reinterpret_cast<const uint4*>(ql + 64 * hh);
                    *reinterpret_cast<uint4*>(a + 4) = *reinterpret_cast<const uint4*>(ql + 64 * hh + 16);
                    *reinterpret_cast<uint4*>(b) = *reinterpret_cast<const uint4*>(ql + 64 * hh + 32);
                    *reinterpret_cast<uint4*>(b + 4) = *reinterpret_cast<const uint4*>(ql + 64 * hh + 48);
                    *reinterpret_cast<uint4*>(h) = *reinterpret_cast<const uint4*>(qh + 32 * hh);
                    *reinterpret_cast<uint4*>(h + 4) = *reinterpret_cast<const uint4*>(qh + 32 * hh + 16);
                    int s[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};   // [quarter i][l >= 16]
#pragma unroll
                    for (int i = 0; i < 4; i++) {   // the four 32-value quarters of this 128-half
                        uint32_t x[8];
                        *reinterpret_cast<uint4*>(x) = *reinterpret_cast<const uint4*>(aq + 128 * hh + 32 * i);
                        *reinterpret_cast<uint4*>(x + 4) = *reinterpret_cast<const uint4*>(aq + 128 * hh + 32 * i + 16);
#pragma unroll
                        for (int w = 0; w < 8; w++) {
                            uint32_t v;
                            if (i == 0) v = (a[w] & 0x0f0f0f0fu) | ((h[w] << 4) & 0x30303030u);
                            else if (i == 1) v = (b[w] & 0x0f0f0f0fu) | ((h[w] << 2) & 0x30303030u);
                            else if (i == 2) v = ((a[w] >> 4) & 0x0f0f0f0fu) | (h[w] & 0x30303030u);
                            else v = ((b[w] >> 4) & 0x0f0f0f0fu) | ((h[w] >> 2) & 0x30303030u);
                            s[i][w >> 2] = dp4a_s8s8(v, x[w], s[i][w >> 2]);
                        }
                    }
                    // 16-value group g = 8*hh + 2*i + odd carries scale byte g
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
                const float val = (d * dx[ab]) * (float)(isum - 32 * corr);
                acc[0] += rw == 0 ? val : 0.f; acc[1] += rw == 1 ? val : 0.f; acc[2] += rw == 2 ? val : 0.f; acc[3] += rw == 3 ? val : 0.f;
            }
        }
        const float res = warp_reduce4(acc[0], acc[1], acc[2], acc[3], lane);
        if ((lane & 7) == 0) partial[(quad * RW + (lane >> 3)) * ns + j] = cur_ok ? res : 0.f;
        __syncwarp();
        cur_ok = next_ok;
    }
    cp_async_wait_group<0>();
    __syncthreads();
    for (int hl = threadIdx.x; hl < nrows; hl += WARPS * 32) {
        float acc = 0.f;
        for (int j = 0; j < k; j++) {
            if ((skip >> j) & 1u) continue;
            const float dv = partial[hl * ns + j];
            acc = p.weights ? __fmaf_rn(dv, p.weights[(long)t * k + j], acc) : acc + dv;
        }
        const long o = (long)t * p.rows + r0 + hl;
        if (p.xw) acc = round_hidden(acc, p.hidden_type) + round_hidden(partial[hl * ns + k], p.hidden_type);
        if (p.accumulate) acc = load_hidden(p.out, o, p.hidden_type) + round_hidden(acc, p.hidden_type);
        store_hidden(p.out, o, p.hidden_type, acc);
    }
  }  // tokens
}

}  // namespace ktb
