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

__constant__ uint8_t c_ctx_lut[2048];   // mode*512 + {lut0[256], lut1[256]}   (codec/interface.rs:199-238)

struct Speed2 { int inc, lim; };

// ---------------------------------------------------------------------------------------------------------------
// input cursor: walks the mux records of ONE of the two interleaved byte streams in place (mux.rs:384-444), so the
// raw .divans bytes are consumed straight from HBM -- no host-side demux pass.
// ---------------------------------------------------------------------------------------------------------------
struct InCursor {
    const uint8_t *p;      // next payload byte of the current record
    const uint8_t *nxt;    // first byte after the current record (= next record header)
    const uint8_t *end;    // EOF marker position
    uint32_t rem;          // payload bytes left in the current record
    uint32_t sid;
    uint32_t underflow;
};

__device__ __forceinline__ void cur_next_record(InCursor &c) {
    while (true) {
        if (c.nxt >= c.end) { c.underflow = 1; c.rem = 0; return; }
        uint32_t b = __ldg(c.nxt);
        uint32_t len, hdr;
        if (b < 16) { len = ((uint32_t)__ldg(c.nxt + 1) | ((uint32_t)__ldg(c.nxt + 2) << 8)) + 1; hdr = 3; }
        else { len = 1024u << ((b >> 4) << 1); hdr = 1; }
        if ((b & 1) == c.sid) { c.p = c.nxt + hdr; c.rem = len; c.nxt = c.p + len; return; }
        c.nxt += hdr + len;
    }
}
__device__ __forceinline__ uint32_t cur_byte(InCursor &c) {
    while (c.rem == 0) { if (c.underflow) return 0; cur_next_record(c); if (c.underflow) return 0; }
    uint32_t v = __ldg(c.p); c.p++; c.rem--; return v;
}
__device__ __forceinline__ uint32_t ldg_u32_unaligned(const uint8_t *p) {
    // two aligned words + funnel shift; reads at most 3 bytes past p+4, always inside the stream (the 11 bytes of EOF
    // marker + trailer follow every record)
    const uint32_t *q = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
    uint32_t lo = __ldg(q), hi = __ldg(q + 1);
    return __funnelshift_r(lo, hi, ((uint32_t)reinterpret_cast<uintptr_t>(p) & 3u) * 8u);
}
__device__ __forceinline__ uint32_t cur_u32_slow(InCursor &c) {
    uint32_t v = cur_byte(c); v |= cur_byte(c) << 8; v |= cur_byte(c) << 16; v |= cur_byte(c) << 24;
    return v;
}

// ---------------------------------------------------------------------------------------------------------------
// rANS coder state (src/ans.rs:142-148 decoder; :254-301 encoder records (start,freq))
// ---------------------------------------------------------------------------------------------------------------
struct Coder {
    uint64_t a, b;
    uint32_t sym_count, need_a, need_b;
    InCursor in;
    // encoder
    uint32_t *sf;          // (start | freq<<16) log for this coder, whole stream
    uint32_t n_sf;
    uint32_t n_syms;
};

__device__ __forceinline__ void coder_init_dec(Coder &k, const uint8_t *body, const uint8_t *end, uint32_t sid) {
    k.a = k.b = 0; k.sym_count = 0; k.need_a = 8; k.need_b = 0;   // ans.rs:150-162
    k.in.p = body; k.in.nxt = body; k.in.end = end; k.in.rem = 0; k.in.sid = sid; k.in.underflow = 0;
    k.sf = nullptr; k.n_sf = 0; k.n_syms = 0;
}
// rare paths, by value so that the caller's Coder stays in registers: a word that straddles two mux records, and the
// 16-byte (re)initialisation at stream start / every 65536 symbols (ans.rs:173-189)
static __device__ __noinline__ Coder coder_fill_slow(Coder k) {
    if (k.need_a == 1) {
        uint32_t w = cur_u32_slow(k.in);
        k.a = (k.a << 32) | (uint64_t)w;
    } else {
        uint32_t w0 = cur_u32_slow(k.in), w1 = cur_u32_slow(k.in), w2 = cur_u32_slow(k.in), w3 = cur_u32_slow(k.in);
        k.a = (uint64_t)w0 | ((uint64_t)w1 << 32);
        k.b = (uint64_t)w2 | ((uint64_t)w3 << 32);
        k.sym_count = 0;
    }
    k.need_a = 0;
    return k;
}
__device__ __forceinline__ void coder_fill(Coder &k) {
    // ans.rs:428-442 (push_data)
    if (k.need_a == 0) return;
    if (k.need_a == 1 && k.in.rem >= 4) {
        uint32_t w = ldg_u32_unaligned(k.in.p);
        k.in.p += 4; k.in.rem -= 4;
        k.a = (k.a << 32) | (uint64_t)w;
        k.need_a = 0;
        return;
    }
    k = coder_fill_slow(k);
}
__device__ __forceinline__ void coder_advance(Coder &k, int start, int freq) {
    // ans.rs:230-244
    k.need_a = k.need_b | ((k.sym_count == NUM_SYMBOLS_BEFORE_FLUSH - 1) ? 8u : 0u);
    uint64_t x = (uint64_t)(int64_t)freq * (k.a >> 15) + (k.a & 0x7fff) - (uint64_t)(int64_t)start;
    k.sym_count = (k.sym_count + 1) & 0xffff;
    k.need_b = (x < (1ull << 31)) ? 1u : 0u;
    k.a = k.b; k.b = x;
    k.n_syms++;
}

// exact floor((c<<15)/max): the reference divides through a reciprocal LUT that is asserted equal to integer
// division (probability/numeric.rs:26-31, make_div_lut.rs:37-39).  fp32 estimate + integer fix-up: c<<15 has <=16
// significant bits, so the estimate is within 1 of the true quotient (<= 2^16).
__device__ __forceinline__ int cdf_div(int c, int maxv) {
    uint32_t d = (uint32_t)maxv & 0xffffu;
    uint32_t n = (uint32_t)(c << 15);
    if (d == 0) return (int)(n >> 1);           // RECIPROCAL[0] = (0,0) quirk (div_lut.rs)
    float q = __uint2float_rz(n) * __fdividef(1.0f, __uint2float_rz(d));   // rcp.approx: 1 ulp, fixed up below
    uint32_t qi = __float2uint_rz(q);
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

// Code one nibble against the CDF whose element l16 is `c` (max = maxv).  Returns the symbol; start/freq out.
template <bool ENC>
__device__ __forceinline__ int code_cdf(Coder &k, const Grp g, int c, int maxv, int sym_in, int &start, int &freq) {
    int sym;
    if (!ENC) {
        coder_fill(k);
        int off = (int)(k.a & 0x7fff);
        int r = (int)(short)((off * maxv) >> 15);                       // probability/interface.rs:140
        bool pred = (g.l16 == 15) || (r < c);
        unsigned bal = __ballot_sync(g.mask, pred);
        sym = __ffs((bal >> g.shift) & 0xffffu) - 1;
    } else {
        sym = sym_in;
    }
    int cum = cdf_div(c, maxv);
    int hi = __shfl_sync(g.mask, cum, sym, 16);
    int lo = __shfl_sync(g.mask, cum, (sym - 1) & 15, 16);
    if (sym == 0) lo = 0;
    start = (int)(short)(lo + 1);                                       // "major hax", probability/interface.rs:103-104
    freq = (int)(short)(hi - lo - 1);
    if (!ENC) {
        coder_advance(k, start, freq);
    } else {
        if (g.store0) k.sf[k.n_sf] = ((uint32_t)start & 0xffffu) | ((uint32_t)freq << 16);   // ans.rs:289-296
        k.n_sf++; k.n_syms++;
    }
    return sym;
}

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

// code a nibble against a prior stored in HBM and adapt it
template <bool ENC>
__device__ __forceinline__ int code_prior(Coder &k, const Grp g, int16_t *cdf, int sym_in, int inc, int lim) {
    int c = cdf[g.l16];
    int maxv = cdf[15];
    int start, freq;
    int sym = code_cdf<ENC>(k, g, c, maxv, sym_in, start, freq);
    int c2 = cdf_blend(g, c, maxv, sym, inc, lim);
    if (g.writer) cdf[g.l16] = (int16_t)c2;
    __syncwarp(g.mask);
    return sym;
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

}  // namespace dv
