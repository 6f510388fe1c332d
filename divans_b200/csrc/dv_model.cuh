// dv_model.cuh -- the divANS model + rANS coder as warp-level device code (sm_100a).
//
// One 16-lane group owns one stream: lane i (0..15) of the group holds element i of whichever 16-entry adaptive
// CDF is being coded, so the reference's per-symbol work becomes
//   bin search  (probability/interface.rs:136-198)  -> per-lane compare + __ballot_sync + __ffs
//   start/freq  (probability/interface.rs:97-108)   -> per-lane exact (c<<15)/max, two __shfl_sync
//   blend       (probability/frequentist_cdf.rs:74-85) -> per-lane predicated add + uniform rescale branch
//   average     (probability/frequentist_cdf.rs:58-72) -> per-lane integer mad
// Everything that is scalar in the reference (rANS states, book-keeping) is replicated across the lanes of the
// group (uniform registers).  With LANES_PER_STREAM==32 the upper half-warp mirrors the lower one (same stream);
// with 16 the two half-warps run two different streams.
//
// The same code runs as decoder (ENC=false: symbols come out of the rANS state) and as encoder (ENC=true: symbols
// are given, (start,freq) pairs are recorded for the reverse rANS pass) -- exactly how the reference shares
// DivansCodec between ANSDecoder and ANSEncoder (codec/mod.rs:160-186).
#pragma once
#include "dv_common.cuh"

namespace dv {


struct Speed2 { int inc, lim; };

// ---------------------------------------------------------------------------------------------------------------
// rANS coder state (src/ans.rs:142-148 decoder; :254-301 encoder records (start,freq)).
// The decoder reads its byte stream as aligned 32-bit words from a CONTIGUOUS payload: the mux records of the two
// interleaved streams (mux.rs:384-444) are compacted by dv::demux_kernel before the stream kernel runs.
// ---------------------------------------------------------------------------------------------------------------
struct Coder {
    uint64_t a, b;
    const uint32_t *p;     // DEC: next payload word.  ENC: base of this coder's (start | freq<<16) log
    uint32_t left;         // DEC: payload words left. ENC: symbols logged so far
    uint32_t sym_count, need_a, need_b, underflow;
};

__device__ __forceinline__ void coder_init_dec(Coder &k, const uint32_t *payload, uint32_t n_words) {
    k.a = k.b = 0; k.sym_count = 0; k.need_a = 8; k.need_b = 0; k.underflow = 0;   // ans.rs:150-162
    k.p = payload; k.left = n_words;
}
__device__ __forceinline__ void coder_init_enc(Coder &k, uint32_t *log) {
    k.a = k.b = 0; k.sym_count = 0; k.need_a = 0; k.need_b = 0; k.underflow = 0;
    k.p = log; k.left = 0;
}
__device__ __forceinline__ void coder_fill(Coder &k) {
    // ans.rs:428-442 (push_data) and :173-189 (16-byte (re)initialisation at stream start / every 65536 symbols)
    if (k.need_a == 0) return;
    if (k.need_a == 1) {
        if (k.left >= 1) { uint32_t w = __ldg(k.p); k.p += 1; k.left -= 1; k.a = (k.a << 32) | (uint64_t)w; }
        else { k.underflow = 1; k.a <<= 32; }
    } else {
        if (k.left >= 4) {
            uint4 w = make_uint4(__ldg(k.p), __ldg(k.p + 1), __ldg(k.p + 2), __ldg(k.p + 3));
            k.p += 4; k.left -= 4;
            k.a = (uint64_t)w.x | ((uint64_t)w.y << 32);
            k.b = (uint64_t)w.z | ((uint64_t)w.w << 32);
        } else { k.underflow = 1; k.a = k.b = 0; k.left = 0; }
        k.sym_count = 0;
    }
    k.need_a = 0;
}
__device__ __forceinline__ void coder_advance(Coder &k, int start, int freq) {
    // ans.rs:230-244
    k.need_a = k.need_b | ((k.sym_count == NUM_SYMBOLS_BEFORE_FLUSH - 1) ? 8u : 0u);
    uint64_t x = (uint64_t)(int64_t)freq * (k.a >> 15) + (k.a & 0x7fff) - (uint64_t)(int64_t)start;
    k.sym_count = k.sym_count + 1;   // reset by the 16-byte re-initialisation that always follows symbol 65535
    k.need_b = (x < (1ull << 31)) ? 1u : 0u;
    k.a = k.b; k.b = x;
}

// exact floor((c<<15)/max): the reference divides through a reciprocal LUT that is asserted equal to integer
// division (probability/numeric.rs:26-31, make_div_lut.rs:37-39).  fp32 estimate + integer fix-up: c<<15 has <=16
// significant bits, so the estimate is within 1 of the true quotient (<= 2^16).
__device__ __forceinline__ int cdf_div(int c, int maxv) {
    uint32_t d = (uint32_t)maxv & 0xffffu;
    uint32_t n = (uint32_t)(c << 15);
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(__uint2float_rz(d)));   // 1 ulp, fixed up below
    float q = __uint2float_rz(n) * rc;
    uint32_t qi = __float2uint_rz(q);
    int32_t r = (int32_t)(n - qi * d);
    if (r < 0) { qi--; r += (int32_t)d; }
    if (r >= (int32_t)d) qi++;
    return (int)qi;
}

// cdf_div for the literal fast loops: operands are known to be valid adaptive values (0 <= c <= max < 2^15), which lets
// both conversions use the 32-bit ALU path (I2FP) instead of the 16-bit XU conversion.
__device__ __forceinline__ int cdf_div_pos(int c, int maxv) {
    const uint32_t d = (uint32_t)maxv;
    const uint32_t n = (uint32_t)c << 15;
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(__uint2float_rz(d)));
    uint32_t qi = __float2uint_rz(__uint2float_rz(n) * rc);
    int32_t r = (int32_t)(n - qi * d);
    if (r < 0) { qi--; r += (int32_t)d; }
    if (r >= (int32_t)d) qi++;
    return (int)qi;
}

// group-wide context handed around (all uniform except l16)
struct Grp {
    unsigned mask;     // participating lanes of this group
    int shift;         // 0 or 16: position of the group's 16 ballot bits
    int l16;           // lane & 15
    bool writer;       // lane that performs CDF stores for element l16 (false on the mirrored upper half)
    bool lane0;        // l16 == 0 (scalar work; true on lane 16 too when the upper half mirrors)
    bool store0;       // the single lane that performs scalar stores for the group
};

// freq only (for the mixing weights: codec/literal.rs:236-239)
__device__ __forceinline__ int cdf_freq(const Grp g, int c, int maxv, int sym) {
    int cum = cdf_div(c, maxv);
    int hi = __shfl_sync(g.mask, cum, sym, 16);
    int lo = __shfl_sync(g.mask, cum, (sym - 1) & 15, 16);
    if (sym == 0) lo = 0;
    return (int)(short)(hi - lo - 1);
}

__device__ __forceinline__ int cdf_blend(const Grp g, int c, int maxv, int sym, int inc, int lim) {
    // probability/frequentist_cdf.rs:74-85, i16 wrapping
    int c2 = (int)(short)(c + ((g.l16 >= sym) ? inc : 0));
    int nm = (int)(short)(maxv + inc);
    if (nm >= lim) {
        int t = (int)(short)(c2 + g.l16 + 1);
        c2 = (int)(short)(t - (t >> 2));
    }
    return c2;
}

// ---------------------------------------------------------------------------------------------------------------
// mixing weights (codec/weights.rs) -- uniform scalar math
// ---------------------------------------------------------------------------------------------------------------
struct Weights { int w0, w1; int norm; };   // norm kept as the u16 view used by literal.rs:230
__device__ __forceinline__ int weights_new(int prob, int weighted, int wi) {
    long long p1 = weighted, total = 1 << 15, p0 = total - p1;
    long long efficacy = total * (long long)prob - p1 * total;
    unsigned long long geo = (unsigned long long)(p1 * p0);
    int lg = geo ? 64 - __clzll((long long)geo) : 0;
    long long prod = (total - p1) * efficacy;
    long long adj = prod >> lg;
    int nw = (int)(unsigned int)(unsigned long long)((long long)wi + adj);
    return nw > 1 ? nw : 1;
}
__device__ __forceinline__ void weights_update(Weights &w, int f_cm, int f_nb, int weighted) {
    if (((w.w0 | w.w1) & 0x7f000000) != 0) {   // fix_weights, codec/weights.rs:64-79
        int ilog = 32 - min(__clz(w.w0), __clz(w.w1));
        if (ilog >= 24) { w.w0 >>= ilog - 24; w.w1 >>= ilog - 24; }
    }
    int n0 = weights_new(f_cm, weighted, w.w0);
    int n1 = weights_new(f_nb, weighted, w.w1);
    w.w0 = n0; w.w1 = n1;
    long long total = (long long)n0 + (long long)n1;        // compute_normalized_weight, :54-62
    int lz = __clzll(total);
    int shift = max(56 - lz, 0);
    unsigned d = (unsigned)(total >> shift) & 0xffu;
    int recip = d ? 1 + (1 << 24) / (int)d : 0;             // RECIPROCAL8 (div_lut.rs)
    unsigned num = ((unsigned)(n0 >> shift) << 8) & 0xffffu;
    int q = (int)(short)(((long long)recip * (long long)num) >> 24);
    w.norm = (int)(unsigned short)(q << 7);
}

// 32-bit restatement of weights_update for the literal fast path.  Identical results whenever the coded frequencies
// are in 1..32767 (always true for streams an encoder can produce): efficacy = 2^15 * (prob - p1) and
// prod = p0 * efficacy, so prod >> lg is (p0 * (prob - p1)) shifted by (15 - lg) -- no 64-bit arithmetic needed.
__device__ __forceinline__ int weights_new32(int prob, int p1, int wi) {
    const int p0 = (1 << 15) - p1;
    const int t = p0 * (prob - p1);
    const unsigned geo = (unsigned)(p1 * p0);
    const int lg = geo ? 32 - __clz((int)geo) : 0;
    const int adj = lg <= 15 ? (int)((unsigned)t << (15 - lg)) : (t >> (lg - 15));
    const int nw = (int)((unsigned)wi + (unsigned)adj);
    return nw > 1 ? nw : 1;
}
__device__ __forceinline__ void weights_update32(Weights &w, int f_cm, int f_nb, int weighted) {
    if (((w.w0 | w.w1) & 0x7f000000) != 0) {   // fix_weights, codec/weights.rs:64-79
        int ilog = 32 - min(__clz(w.w0), __clz(w.w1));
        if (ilog >= 24) { w.w0 >>= ilog - 24; w.w1 >>= ilog - 24; }
    }
    const int n0 = weights_new32(f_cm, weighted, w.w0);
    const int n1 = weights_new32(f_nb, weighted, w.w1);
    w.w0 = n0; w.w1 = n1;
    const unsigned total = (unsigned)n0 + (unsigned)n1;       // compute_normalized_weight, :54-62 (total < 2^32)
    const int shift = max(24 - __clz((int)total), 0);
    const unsigned d = (total >> shift) & 0xffu;
    const int recip = d ? 1 + cdf_div(512, (int)d) : 0;       // RECIPROCAL8[d] = 1 + 2^24 / d (div_lut.rs)
    const unsigned num = ((unsigned)(n0 >> shift) << 8) & 0xffffu;
    const int q = (int)(short)(((unsigned long long)(unsigned)recip * num) >> 24);
    w.norm = (int)(unsigned short)(q << 7);
}

}  // namespace dv
