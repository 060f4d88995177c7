/* TEST INFRASTRUCTURE ONLY — see ktoracle.h.  Plain-C restatement of the reference CPU path.
 *
 * Every routine cites the reference function whose arithmetic it restates.  The integer part
 * (unpacking, int8 x int8 sums, bsums) is exact; the fp32 part follows the order of the
 * reference's scalar ("#else") code paths.  The reference's SIMD paths (iqk_mul_mat, AVX2,
 * AVX-512) re-associate those fp32 sums differently, so even two builds of the reference differ
 * from each other by ~1e-6 relative (tests/test_oracle_pinned.py measures this).
 *
 * Build: see oracle/Makefile (-ffp-contract=off so that the source order is what runs).
 */
#include "ktoracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define QK_K 256

/* ------------------------------------------------------------------ scalar conversions ----- */

/* fp16 -> fp32, exact (ggml-impl.h ggml_compute_fp16_to_fp32; the reference uses F16C or a table
 * filled with it in CPUInfer's ctor, cpuinfer.h:41-43). */
float kto_fp16_to_fp32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f;
    uint32_t man = h & 0x3ff;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal: normalise */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400));
            man &= 0x3ff;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

/* fp32 -> fp16 round-to-nearest-even (GGML_FP32_TO_FP16 == _cvtss_sh(x, 0) on F16C hosts). */
uint16_t kto_fp32_to_fp16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) { /* inf / nan */
        return (uint16_t)(sign | 0x7c00u | (ax > 0x7f800000u ? (0x200u | ((ax >> 13) & 0x3ff)) : 0));
    }
    if (ax >= 0x477ff000u) { /* rounds to >= 65520 -> inf */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (ax < 0x33000001u) { /* < 2^-25 (or == 2^-25 ties to even zero) */
        return (uint16_t)sign;
    }
    int e = (int)(ax >> 23) - 127;
    uint32_t m = (ax & 0x7fffffu) | 0x800000u;
    if (e < -14) { /* subnormal half */
        int shift = -14 - e + 13; /* bits to drop */
        uint32_t q = m >> shift;
        uint32_t rem = m & ((1u << shift) - 1);
        uint32_t half = 1u << (shift - 1);
        if (rem > half || (rem == half && (q & 1))) q++;
        return (uint16_t)(sign | q);
    }
    uint32_t q = ((uint32_t)(e + 15) << 10) | ((m >> 13) & 0x3ff);
    uint32_t rem = m & 0x1fff;
    if (rem > 0x1000 || (rem == 0x1000 && (q & 1))) q++;
    return (uint16_t)(sign | q);
}

/* ggml-impl.h:64-84 */
float kto_bf16_to_fp32(uint16_t h) {
    uint32_t bits = (uint32_t)h << 16;
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

/* ggml-impl.h:87-104: RNE, NaN quieted, subnormals flushed to (signed) zero. */
uint16_t kto_fp32_to_bf16(float f) {
    uint32_t i;
    memcpy(&i, &f, 4);
    if ((i & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((i >> 16) | 64);
    if (!(i & 0x7f800000u)) return (uint16_t)((i & 0x80000000u) >> 16);
    return (uint16_t)((i + (0x7fffu + ((i >> 16) & 1))) >> 16);
}

/* ggml-quants.c:1632-1637 — round-to-nearest-even through the 1.5*2^23 magic constant, applied to
 * the product a*b.  The reference source reads nearest_int(iscale*x[j]); once inlined this is
 * `iscale*x[j] + 12582912.f`, which every x86-64 build of the reference with FMA available
 * (-march=native; GNU C defaults to -ffp-contract=fast) contracts into ONE fused multiply-add, i.e.
 * the product is NOT rounded to fp32 before the magic add.  That changes the int8 result on exact
 * .5 ties (common for bf16-valued inputs) and a one-LSB activation flip moves the MoE output by
 * ~1e-3, so the as-built behaviour is what parity is pinned to (tests/test_oracle_pinned.py checks
 * byte equality with oracle/_ref).  The CUDA path uses __fmaf_rn for the same reason. */
static inline int nearest_int_mul(float a, float b) {
    float val = fmaf(a, b, 12582912.f);
    int i;
    memcpy(&i, &val, sizeof(int));
    return (i & 0x007fffff) - 0x00400000;
}

/* ------------------------------------------------------------------ block formats ---------- */
/* ggml-common.h: block_q8_0 :183-186, block_q2_K :210-220, block_q3_K :227-232,
 * block_q4_K :249-255, block_q5_K :256-267, block_q6_K :270-279, block_q8_K :283-287,
 * block_iq4_xs :365-370.  All are byte-packed; accessed here through byte offsets. */
#define SZ_Q8_0 34
#define SZ_Q2_K 84
#define SZ_Q3_K 110
#define SZ_Q4_K 144
#define SZ_Q5_K 176
#define SZ_Q6_K 210
#define SZ_Q8_K 292
#define SZ_IQ4_XS 136

static inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
static inline float rdf32(const uint8_t* p) { float v; memcpy(&v, p, 4); return v; }

long kto_type_size(int t) {
    switch (t) {
        case KTO_F32: return 4;
        case KTO_F16: case KTO_BF16: return 2;
        case KTO_Q8_0: return SZ_Q8_0;
        case KTO_Q2_K: return SZ_Q2_K;
        case KTO_Q3_K: return SZ_Q3_K;
        case KTO_Q4_K: return SZ_Q4_K;
        case KTO_Q5_K: return SZ_Q5_K;
        case KTO_Q6_K: return SZ_Q6_K;
        case KTO_Q8_K: return SZ_Q8_K;
        case KTO_IQ4_XS: return SZ_IQ4_XS;
        default: return 0;
    }
}
long kto_blck_size(int t) {
    switch (t) {
        case KTO_F32: case KTO_F16: case KTO_BF16: return 1;
        case KTO_Q8_0: return 32;
        case KTO_Q2_K: case KTO_Q3_K: case KTO_Q4_K: case KTO_Q5_K: case KTO_Q6_K: case KTO_Q8_K:
        case KTO_IQ4_XS: return QK_K;
        default: return 0;
    }
}
/* ggml.c type_traits[].vec_dot_type (:660-905) */
int kto_vec_dot_type(int t) {
    switch (t) {
        case KTO_F32: return KTO_F32;
        case KTO_F16: return KTO_F16;
        case KTO_BF16: return KTO_BF16;
        case KTO_Q8_0: return KTO_Q8_0;
        case KTO_Q2_K: case KTO_Q3_K: case KTO_Q4_K: case KTO_Q5_K: case KTO_Q6_K: case KTO_IQ4_XS:
            return KTO_Q8_K;
        default: return -1;
    }
}

static const int8_t kvalues_iq4nl[16] = {-127, -104, -83, -65, -49, -35, -22, -10, 1, 13, 25, 38, 53, 69, 89, 113};

/* get_scale_min_k4 (ggml-quants.c:1891-1899): 8 x (6-bit scale, 6-bit min) packed in 12 bytes. */
static inline void scale_min_k4(int j, const uint8_t* q, uint8_t* d, uint8_t* m) {
    if (j < 4) {
        *d = q[j] & 63;
        *m = q[j + 4] & 63;
    } else {
        *d = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4);
        *m = (q[j + 4] >> 4) | ((q[j - 0] >> 6) << 4);
    }
}

/* Uniform integer view of one 256-element K-quant super-block:
 *   value[i] = d * isc[i/16] * q[i]  -  dmin * imn[i/16]
 * q: signed 8-bit quant, isc: integer sub-scale per 16, imn: integer sub-min per 16.
 * This is exactly what each dequantize_row_* / vec_dot_* of the reference computes, written once. */
typedef struct {
    float d, dmin;
    int8_t q[QK_K];
    int isc[16];
    int imn[16];
} kblock_t;

static void unpack_block(int type, const uint8_t* b, kblock_t* o) {
    memset(o->imn, 0, sizeof(o->imn));
    o->dmin = 0.f;
    switch (type) {
        case KTO_Q4_K: { /* dequantize_row_q4_K ggml-quants.c:2548-2573 */
            o->d = kto_fp16_to_fp32(rd16(b));
            o->dmin = kto_fp16_to_fp32(rd16(b + 2));
            const uint8_t* sc = b + 4;
            const uint8_t* qs = b + 16;
            for (int j = 0; j < 8; j++) {
                uint8_t s, m;
                scale_min_k4(j, sc, &s, &m);
                o->isc[2 * j] = o->isc[2 * j + 1] = s;
                o->imn[2 * j] = o->imn[2 * j + 1] = m;
            }
            for (int g = 0; g < 4; g++)
                for (int l = 0; l < 32; l++) {
                    o->q[g * 64 + l] = (int8_t)(qs[g * 32 + l] & 0xF);
                    o->q[g * 64 + 32 + l] = (int8_t)(qs[g * 32 + l] >> 4);
                }
            break;
        }
        case KTO_Q5_K: { /* dequantize_row_q5_K ggml-quants.c:2756-2781 */
            o->d = kto_fp16_to_fp32(rd16(b));
            o->dmin = kto_fp16_to_fp32(rd16(b + 2));
            const uint8_t* sc = b + 4;
            const uint8_t* qh = b + 16;
            const uint8_t* ql = b + 48;
            for (int j = 0; j < 8; j++) {
                uint8_t s, m;
                scale_min_k4(j, sc, &s, &m);
                o->isc[2 * j] = o->isc[2 * j + 1] = s;
                o->imn[2 * j] = o->imn[2 * j + 1] = m;
            }
            for (int g = 0; g < 4; g++)
                for (int l = 0; l < 32; l++) {
                    int h1 = (qh[l] >> (2 * g)) & 1, h2 = (qh[l] >> (2 * g + 1)) & 1;
                    o->q[g * 64 + l] = (int8_t)((ql[g * 32 + l] & 0xF) + 16 * h1);
                    o->q[g * 64 + 32 + l] = (int8_t)((ql[g * 32 + l] >> 4) + 16 * h2);
                }
            break;
        }
        case KTO_Q6_K: { /* dequantize_row_q6_K ggml-quants.c:2970-3000 */
            const uint8_t* ql = b;
            const uint8_t* qh = b + 128;
            const int8_t* sc = (const int8_t*)(b + 192);
            o->d = kto_fp16_to_fp32(rd16(b + 208));
            for (int j = 0; j < 16; j++) o->isc[j] = sc[j];
            for (int n = 0; n < 2; n++)
                for (int l = 0; l < 32; l++) {
                    const uint8_t* L = ql + 64 * n;
                    uint8_t h = qh[32 * n + l];
                    int8_t* y = o->q + 128 * n;
                    y[l + 0] = (int8_t)((L[l + 0] & 0xF) | (((h >> 0) & 3) << 4)) - 32;
                    y[l + 32] = (int8_t)((L[l + 32] & 0xF) | (((h >> 2) & 3) << 4)) - 32;
                    y[l + 64] = (int8_t)((L[l + 0] >> 4) | (((h >> 4) & 3) << 4)) - 32;
                    y[l + 96] = (int8_t)((L[l + 32] >> 4) | (((h >> 6) & 3) << 4)) - 32;
                }
            break;
        }
        case KTO_Q2_K: { /* dequantize_row_q2_K ggml-quants.c:1972-2002 */
            const uint8_t* sc = b;
            const uint8_t* qs = b + 16;
            o->d = kto_fp16_to_fp32(rd16(b + 80));
            o->dmin = kto_fp16_to_fp32(rd16(b + 82));
            for (int j = 0; j < 16; j++) {
                o->isc[j] = sc[j] & 0xF;
                o->imn[j] = sc[j] >> 4;
            }
            for (int n = 0; n < 2; n++)
                for (int j = 0; j < 4; j++)
                    for (int l = 0; l < 32; l++)
                        o->q[128 * n + 32 * j + l] = (int8_t)((qs[32 * n + l] >> (2 * j)) & 3);
            break;
        }
        case KTO_Q3_K: { /* dequantize_row_q3_K ggml-quants.c:2320-2367 */
            const uint8_t* hm = b;
            const uint8_t* qs = b + 32;
            const uint8_t* s12 = b + 96;
            o->d = kto_fp16_to_fp32(rd16(b + 108));
            uint32_t aux[4], tmp;
            memcpy(aux, s12, 12);
            tmp = aux[2];
            aux[2] = ((aux[0] >> 4) & 0x0f0f0f0fu) | (((tmp >> 4) & 0x03030303u) << 4);
            aux[3] = ((aux[1] >> 4) & 0x0f0f0f0fu) | (((tmp >> 6) & 0x03030303u) << 4);
            aux[0] = (aux[0] & 0x0f0f0f0fu) | (((tmp >> 0) & 0x03030303u) << 4);
            aux[1] = (aux[1] & 0x0f0f0f0fu) | (((tmp >> 2) & 0x03030303u) << 4);
            const int8_t* sc = (const int8_t*)aux;
            for (int j = 0; j < 16; j++) o->isc[j] = sc[j] - 32;
            for (int n = 0; n < 2; n++)
                for (int j = 0; j < 4; j++)
                    for (int l = 0; l < 32; l++) {
                        int bit = (hm[l] >> (4 * n + j)) & 1;
                        o->q[128 * n + 32 * j + l] = (int8_t)(((qs[32 * n + l] >> (2 * j)) & 3) - (bit ? 0 : 4));
                    }
            break;
        }
        case KTO_IQ4_XS: { /* dequantize_row_iq4_xs ggml-quants.c:3568-3589 */
            o->d = kto_fp16_to_fp32(rd16(b));
            uint16_t sh = rd16(b + 2);
            const uint8_t* sl = b + 4;
            const uint8_t* qs = b + 8;
            for (int ib = 0; ib < 8; ib++) {
                int ls = ((sl[ib / 2] >> (4 * (ib % 2))) & 0xf) | (((sh >> (2 * ib)) & 3) << 4);
                o->isc[2 * ib] = o->isc[2 * ib + 1] = ls - 32;
                for (int j = 0; j < 16; j++) {
                    o->q[32 * ib + j] = kvalues_iq4nl[qs[16 * ib + j] & 0xf];
                    o->q[32 * ib + 16 + j] = kvalues_iq4nl[qs[16 * ib + j] >> 4];
                }
            }
            break;
        }
        default:
            memset(o, 0, sizeof(*o));
    }
}

/* ------------------------------------------------------------------ to_float / from_float -- */

void kto_to_float(const void* in, float* out, long n, int type) {
    const uint8_t* p = (const uint8_t*)in;
    switch (type) {
        case KTO_F32: memcpy(out, in, (size_t)n * 4); return;
        case KTO_F16: for (long i = 0; i < n; i++) out[i] = kto_fp16_to_fp32(rd16(p + 2 * i)); return;
        case KTO_BF16: for (long i = 0; i < n; i++) out[i] = kto_bf16_to_fp32(rd16(p + 2 * i)); return;
        case KTO_Q8_0: /* dequantize_row_q8_0 ggml-quants.c:1609-1624 */
            for (long b = 0; b < n / 32; b++) {
                float d = kto_fp16_to_fp32(rd16(p + b * SZ_Q8_0));
                const int8_t* q = (const int8_t*)(p + b * SZ_Q8_0 + 2);
                for (int j = 0; j < 32; j++) out[b * 32 + j] = q[j] * d;
            }
            return;
        case KTO_Q8_K: /* dequantize_row_q8_K ggml-quants.c:3632-3641 */
            for (long b = 0; b < n / QK_K; b++) {
                float d = rdf32(p + b * SZ_Q8_K);
                const int8_t* q = (const int8_t*)(p + b * SZ_Q8_K + 4);
                for (int j = 0; j < QK_K; j++) out[b * QK_K + j] = d * q[j];
            }
            return;
        default: break;
    }
    long ts = kto_type_size(type);
    kblock_t kb;
    for (long b = 0; b < n / QK_K; b++) {
        unpack_block(type, p + b * ts, &kb);
        for (int g = 0; g < 16; g++) {
            /* reference: dl = d*sc ; ml = dmin*m ; y = dl*q - ml   (two roundings before the fma-less sub) */
            float dl = kb.d * (float)kb.isc[g];
            float ml = kb.dmin * (float)kb.imn[g];
            for (int l = 0; l < 16; l++) out[b * QK_K + g * 16 + l] = dl * (float)kb.q[g * 16 + l] - ml;
        }
    }
}

static void quantize_row_q8_K(const float* x, uint8_t* y, long k) { /* ggml-quants.c:3593-3630 */
    for (long i = 0; i < k / QK_K; i++, x += QK_K, y += SZ_Q8_K) {
        float max = 0, amax = 0;
        for (int j = 0; j < QK_K; j++) {
            float ax = fabsf(x[j]);
            if (ax > amax) { amax = ax; max = x[j]; }
        }
        int8_t* qs = (int8_t*)(y + 4);
        if (!amax) {
            float z = 0;
            memcpy(y, &z, 4);
            memset(qs, 0, QK_K);
            /* the reference leaves bsums untouched here; zero them so the oracle is deterministic
             * (they are always multiplied by d == 0). */
            memset(y + 4 + QK_K, 0, 32);
            continue;
        }
        const float iscale = -127.f / max;
        for (int j = 0; j < QK_K; j++) {
            int v = nearest_int_mul(iscale, x[j]);
            qs[j] = (int8_t)(v < 127 ? v : 127);
        }
        for (int j = 0; j < 16; j++) {
            int sum = 0;
            for (int ii = 0; ii < 16; ii++) sum += qs[j * 16 + ii];
            int16_t s16 = (int16_t)sum;
            memcpy(y + 4 + QK_K + 2 * j, &s16, 2);
        }
        float d = 1 / iscale;
        memcpy(y, &d, 4);
    }
}

/* quantize_row_q8_0, as built on x86-64: the type-traits from_float for Q8_0 is the AVX/AVX2 branch
 * (ggml-quants.c:936-1000), not the scalar *_reference row quantiser (:841-864).  They differ:
 * id = 127/amax (not 1/(amax/127)) and rounding is round-to-nearest-EVEN (_mm256_round_ps
 * _MM_ROUND_NEAREST), not roundf's half-away-from-zero.  Pinned byte-for-byte against oracle/_ref. */
static void quantize_row_q8_0(const float* x, uint8_t* y, long k) {
    for (long i = 0; i < k / 32; i++) {
        float amax = 0.0f;
        for (int j = 0; j < 32; j++) {
            float v = fabsf(x[i * 32 + j]);
            if (v > amax) amax = v;
        }
        const float d = amax / 127.f;
        const float id = (amax != 0.0f) ? 127.f / amax : 0.0f;
        uint16_t hd = kto_fp32_to_fp16(d);
        memcpy(y + i * SZ_Q8_0, &hd, 2);
        int8_t* qs = (int8_t*)(y + i * SZ_Q8_0 + 2);
        for (int j = 0; j < 32; j++) qs[j] = (int8_t)nearbyintf(x[i * 32 + j] * id);
    }
}

void kto_from_float(const float* in, void* out, long n, int type) {
    uint8_t* p = (uint8_t*)out;
    switch (type) {
        case KTO_F32: memcpy(out, in, (size_t)n * 4); return;
        case KTO_F16: for (long i = 0; i < n; i++) { uint16_t h = kto_fp32_to_fp16(in[i]); memcpy(p + 2 * i, &h, 2); } return;
        case KTO_BF16: for (long i = 0; i < n; i++) { uint16_t h = kto_fp32_to_bf16(in[i]); memcpy(p + 2 * i, &h, 2); } return;
        case KTO_Q8_0: quantize_row_q8_0(in, p, n); return;
        case KTO_Q8_K: quantize_row_q8_K(in, p, n); return;
        default: abort(); /* weight quantisers are not part of the hot path */
    }
}

/* ------------------------------------------------------------------ dot products ----------- */

/* K-quant (or IQ4_XS) weights x Q8_K activations.
 * Restates ggml_vec_dot_{q2,q3,q4,q5,q6}_K_q8_K / iq4_xs_q8_K scalar branches
 * (ggml-quants.c:5603, 6181, 6962 (#else at 7471), 7530, 8167 (#else at 8772), 11292):
 *   per super-block  isum = sum_g isc[g] * sum_{16} q*q8   (int32, exact)
 *                    msum = sum_g imn[g] * bsums[g]        (int32, exact)
 *   acc += (d_w*d_x) * isum - (dmin_w*d_x) * msum          (fp32)
 * The scalar reference keeps 8 fp32 partial sums per lane position l (sums[l] += d*aux32[l]) and
 * adds them at the end; we keep that order. */
static float dot_k_q8k(int wtype, long n, const uint8_t* w, const uint8_t* a) {
    long ts = kto_type_size(wtype);
    float sums[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float sumf = 0;
    kblock_t kb;
    for (long i = 0; i < n / QK_K; i++) {
        unpack_block(wtype, w + i * ts, &kb);
        const uint8_t* yb = a + i * SZ_Q8_K;
        float yd = rdf32(yb);
        const int8_t* q8 = (const int8_t*)(yb + 4);
        int32_t aux32[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int32_t msum = 0;
        for (int g = 0; g < 16; g++) {
            int16_t bs;
            memcpy(&bs, yb + 4 + QK_K + 2 * g, 2);
            msum += (int32_t)bs * kb.imn[g];
            for (int l = 0; l < 16; l++) aux32[l & 7] += kb.isc[g] * ((int32_t)q8[g * 16 + l] * (int32_t)kb.q[g * 16 + l]);
        }
        const float d = kb.d * yd;
        for (int l = 0; l < 8; l++) sums[l] += d * (float)aux32[l];
        if (kb.dmin != 0.f || msum != 0) {
            const float dmin = kb.dmin * yd;
            sumf -= dmin * (float)msum;
        }
    }
    for (int l = 0; l < 8; l++) sumf += sums[l];
    return sumf;
}

/* ggml_vec_dot_q8_0_q8_0 scalar branch (ggml-quants.c:5585-5600). */
static float dot_q8_0(long n, const uint8_t* w, const uint8_t* a) {
    float sumf = 0;
    for (long i = 0; i < n / 32; i++) {
        const int8_t* x = (const int8_t*)(w + i * SZ_Q8_0 + 2);
        const int8_t* y = (const int8_t*)(a + i * SZ_Q8_0 + 2);
        int sumi = 0;
        for (int j = 0; j < 32; j++) sumi += x[j] * y[j];
        sumf += (float)sumi * (kto_fp16_to_fp32(rd16(w + i * SZ_Q8_0)) * kto_fp16_to_fp32(rd16(a + i * SZ_Q8_0)));
    }
    return sumf;
}

float kto_vec_dot(int wtype, long n, const void* w, const void* act) {
    switch (wtype) {
        case KTO_Q8_0: return dot_q8_0(n, (const uint8_t*)w, (const uint8_t*)act);
        case KTO_Q2_K: case KTO_Q3_K: case KTO_Q4_K: case KTO_Q5_K: case KTO_Q6_K: case KTO_IQ4_XS:
            return dot_k_q8k(wtype, n, (const uint8_t*)w, (const uint8_t*)act);
        case KTO_F32: {
            float s = 0; const float* x = (const float*)w; const float* y = (const float*)act;
            for (long i = 0; i < n; i++) s += x[i] * y[i];
            return s;
        }
        case KTO_F16: case KTO_BF16: {
            /* ggml_vec_dot_f16 / ggml_vec_dot_bf16: products of widened values, fp32 accumulate */
            float s = 0; const uint8_t* x = (const uint8_t*)w; const uint8_t* y = (const uint8_t*)act;
            for (long i = 0; i < n; i++) {
                float a = wtype == KTO_F16 ? kto_fp16_to_fp32(rd16(x + 2 * i)) : kto_bf16_to_fp32(rd16(x + 2 * i));
                float b = wtype == KTO_F16 ? kto_fp16_to_fp32(rd16(y + 2 * i)) : kto_bf16_to_fp32(rd16(y + 2 * i));
                s += a * b;
            }
            return s;
        }
        default: return NAN;
    }
}

/* ------------------------------------------------------------------ operators -------------- */

static inline long row_bytes(long n, int type) { return n / kto_blck_size(type) * kto_type_size(type); }

/* act_fn (moe.cpp:134-136) / act_fn_relu (:138-144) */
static inline float act_silu(float x) { return x / (1.0f + expf(-x)); }
static inline float act_relu(float x) { return x > 0.0f ? x : 0.0f; }

/* Convert one hidden-state row to the activation format a weight type wants (moe.cpp:147-170). */
static void prep_act(const void* in_row, int hidden_type, int wtype, long n, float* tmp_f32, uint8_t* out) {
    int vdt = kto_vec_dot_type(wtype);
    if (vdt == hidden_type) {
        memcpy(out, in_row, (size_t)row_bytes(n, hidden_type));
        return;
    }
    kto_to_float(in_row, tmp_f32, n, hidden_type);
    kto_from_float(tmp_f32, out, n, vdt);
}

void kto_moe_forward(int E, int H, int I, int use_silu, const void* gate, const void* up, const void* down,
                     int gate_type, int up_type, int down_type, int hidden_type, int qlen, int k,
                     const int64_t* ids, const float* weights, const void* input, void* output) {
    const long hrow = row_bytes(H, hidden_type);
    const long g_rb = row_bytes(H, gate_type), u_rb = row_bytes(H, up_type), d_rb = row_bytes(I, down_type);
    const int g_vdt = kto_vec_dot_type(gate_type), u_vdt = kto_vec_dot_type(up_type), d_vdt = kto_vec_dot_type(down_type);
    float* xf = (float*)malloc(sizeof(float) * (size_t)H);
    uint8_t* xg = (uint8_t*)malloc((size_t)row_bytes(H, g_vdt));
    uint8_t* xu = (uint8_t*)malloc((size_t)row_bytes(H, u_vdt));
    float* inter = (float*)malloc(sizeof(float) * (size_t)I * k);
    uint8_t* aq = (uint8_t*)malloc((size_t)row_bytes(I, d_vdt) * k);
    float* outf = (float*)malloc(sizeof(float) * (size_t)H);
    for (int t = 0; t < qlen; t++) {
        const uint8_t* in_row = (const uint8_t*)input + t * hrow;
        prep_act(in_row, hidden_type, gate_type, H, xf, xg);
        prep_act(in_row, hidden_type, up_type, H, xf, xu);
        /* phase 1: gate/up GEMV + activation (moe.cpp:171-210) */
        for (int j = 0; j < k; j++) {
            int64_t e = ids[(long)t * k + j];
            if (e < 0 || e >= E) continue;
            const uint8_t* gw = (const uint8_t*)gate + (size_t)e * I * g_rb;
            const uint8_t* uw = (const uint8_t*)up + (size_t)e * I * u_rb;
#pragma omp parallel for schedule(static)
            for (int r = 0; r < I; r++) {
                float g = kto_vec_dot(gate_type, H, gw + (size_t)r * g_rb, xg);
                float u = kto_vec_dot(up_type, H, uw + (size_t)r * u_rb, xu);
                inter[(long)j * I + r] = (use_silu ? act_silu(g) : act_relu(g)) * u;
            }
            /* down-input quantisation (moe.cpp:205-215) */
            kto_from_float(inter + (long)j * I, aq + (size_t)j * row_bytes(I, d_vdt), I, d_vdt);
        }
        /* phase 2: down GEMV, weighted accumulation in expert_ids order (moe.cpp:216-245) */
#pragma omp parallel for schedule(static)
        for (int h = 0; h < H; h++) {
            float acc = 0;
            for (int j = 0; j < k; j++) {
                int64_t e = ids[(long)t * k + j];
                if (e < 0 || e >= E) continue;
                const uint8_t* dw = (const uint8_t*)down + ((size_t)e * H + h) * d_rb;
                float dv = kto_vec_dot(down_type, I, dw, aq + (size_t)j * row_bytes(I, d_vdt));
                acc += dv * weights[(long)t * k + j];
            }
            outf[h] = acc;
        }
        kto_from_float(outf, (uint8_t*)output + t * hrow, H, hidden_type);
    }
    free(xf); free(xg); free(xu); free(inter); free(aq); free(outf);
}

/* Linear::forward_many (linear.cpp:37-63) */
void kto_linear_forward(int in_size, int out_size, const void* proj, int proj_type, int hidden_type, int qlen,
                        const void* input, void* output) {
    const int vdt = kto_vec_dot_type(proj_type);
    const long rb = row_bytes(in_size, proj_type);
    float* xf = (float*)malloc(sizeof(float) * (size_t)in_size);
    uint8_t* xq = (uint8_t*)malloc((size_t)row_bytes(in_size, vdt));
    float* of = (float*)malloc(sizeof(float) * (size_t)out_size);
    for (int t = 0; t < qlen; t++) {
        prep_act((const uint8_t*)input + t * row_bytes(in_size, hidden_type), hidden_type, proj_type, in_size, xf, xq);
#pragma omp parallel for schedule(static)
        for (int r = 0; r < out_size; r++) of[r] = kto_vec_dot(proj_type, in_size, (const uint8_t*)proj + (size_t)r * rb, xq);
        kto_from_float(of, (uint8_t*)output + t * row_bytes(out_size, hidden_type), out_size, hidden_type);
    }
    free(xf); free(xq); free(of);
}

/* MLP::forward_many (mlp.cpp:47-117) — always SiLU. */
void kto_mlp_forward(int H, int I, const void* gate, const void* up, const void* down, int gate_type, int up_type,
                     int down_type, int hidden_type, int qlen, const void* input, void* output) {
    const long g_rb = row_bytes(H, gate_type), u_rb = row_bytes(H, up_type), d_rb = row_bytes(I, down_type);
    const int g_vdt = kto_vec_dot_type(gate_type), u_vdt = kto_vec_dot_type(up_type), d_vdt = kto_vec_dot_type(down_type);
    float* xf = (float*)malloc(sizeof(float) * (size_t)H);
    uint8_t* xg = (uint8_t*)malloc((size_t)row_bytes(H, g_vdt));
    uint8_t* xu = (uint8_t*)malloc((size_t)row_bytes(H, u_vdt));
    float* inter = (float*)malloc(sizeof(float) * (size_t)I);
    uint8_t* aq = (uint8_t*)malloc((size_t)row_bytes(I, d_vdt));
    float* outf = (float*)malloc(sizeof(float) * (size_t)H);
    for (int t = 0; t < qlen; t++) {
        const uint8_t* in_row = (const uint8_t*)input + t * row_bytes(H, hidden_type);
        prep_act(in_row, hidden_type, gate_type, H, xf, xg);
        prep_act(in_row, hidden_type, up_type, H, xf, xu);
#pragma omp parallel for schedule(static)
        for (int r = 0; r < I; r++) {
            float g = kto_vec_dot(gate_type, H, (const uint8_t*)gate + (size_t)r * g_rb, xg);
            float u = kto_vec_dot(up_type, H, (const uint8_t*)up + (size_t)r * u_rb, xu);
            inter[r] = act_silu(g) * u;
        }
        kto_from_float(inter, aq, I, d_vdt);
#pragma omp parallel for schedule(static)
        for (int h = 0; h < H; h++) outf[h] = kto_vec_dot(down_type, I, (const uint8_t*)down + (size_t)h * d_rb, aq);
        kto_from_float(outf, (uint8_t*)output + t * row_bytes(H, hidden_type), H, hidden_type);
    }
    free(xf); free(xg); free(xu); free(inter); free(aq); free(outf);
}
