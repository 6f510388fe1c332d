// dv8_core.cuh -- the 8-lane decode engine: FOUR streams per warp, lane j of a group holds CDF elements 2j and 2j+1
// packed in one 32-bit register (c[2j] | c[2j+1] << 16).
//
// Why 8 lanes: the decoder is bound by the per-nibble dependency chain of each stream (load the prior -> bin search ->
// next prior's address -> load), not by lanes.  Halving the lanes per stream halves the warp instructions issued per
// decoded nibble (the scalar part of the chain -- rANS state, context, addresses -- is replicated per lane anyway) and
// doubles the streams a B200 keeps resident (9472 at 16 warps per SM).  What the 16-lane engine (dv_core.cuh) does per
// element, this engine does on packed pairs:
//   bin search  (probability/interface.rs:136-198)  two compares per lane, two ballots, one popc over both
//   start/freq  (probability/interface.rs:97-108)   one reciprocal per lane (2^32 / max, under-estimated), two exact
//                                                   quotients by multiply-high + one fix-up, two shuffles of packed pairs
//   blend       (frequentist_cdf.rs:74-85)          one packed add, one packed rescale
// plus, in the literal loop:
//   * generation tags instead of initialisation: a literal prior whose tag byte differs from the stream's generation reads
//     as the default CDF and is tagged when first written (dv_common.cuh OFF_TAGS_*);
//   * the 16 candidate priors of the NEXT low nibble (one per possible high nibble) are 512 contiguous bytes
//     (lit_index_lo) and are prefetched into L1 as soon as the previous byte is known -- one dependent L2 round trip per
//     byte instead of two;
//   * literal context in one lookup (table T2 built per stream / block type, OFF_T2);
//   * the next payload word of the eager-refill coder is always in a register (the refill never waits for memory);
//   * decoded literals leave as aligned 8-byte stores of last_8_literals.
// Decode only; the encoder's model pass keeps the 16-lane engine.
#pragma once
#include "dv_engine_kernel.cuh"
#include "dv_kernels.h"

namespace dv {

constexpr int SMEM_BYTES_PER_GROUP8 = (int)((sizeof(Cold) + 15) / 16 * 16);

__device__ __forceinline__ uint32_t ld_u32(const void *p) { uint32_t v; asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ int ld_s16g(const void *p) { int v; asm volatile("ld.global.s16 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint32_t ld_u8g(const void *p) { uint32_t v; asm volatile("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// default CDF [4,8,...,64] (frequentist_cdf.rs:17-23), elements 2j and 2j+1 of lane j
__device__ __forceinline__ uint32_t default_pair(int li) { return (uint32_t)(8 * li + 4) | ((uint32_t)(8 * li + 8) << 16); }

// ---------------------------------------------------------------------------------------------------------------
// generic nibble core, 8 lanes per stream (command nibbles, literals outside the fast loop, dynamic context mixing).
// Each lane treats its two elements as separate ints with the reference's i16 semantics (wrap included), i.e. the
// arithmetic of dv_core.cuh nibble_core twice per lane.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int blend_elem(int c, int elem, int maxv, int sym, int inc, int lim) {   // frequentist_cdf.rs:74-85, i16 wrapping
    int c2 = (int)(short)(c + ((elem >= sym) ? inc : 0));
    const int nm = (int)(short)(maxv + inc);
    if (nm >= lim) { const int t = (int)(short)(c2 + elem + 1); c2 = (int)(short)(t - (t >> 2)); }
    return c2;
}
// cumulative value (c << 15) / max of elements `sym` and `sym - 1`, fetched from the lanes that own them
__device__ __forceinline__ void cum_pair(const int cum0, const int cum1, const int sym, int &hi, int &lo) {
    const int prev = (sym - 1) & 15;
    hi = __shfl_sync(FULL, (sym & 1) ? cum1 : cum0, sym >> 1, 8);
    lo = __shfl_sync(FULL, (prev & 1) ? cum1 : cum0, prev >> 1, 8);
    if (sym == 0) lo = 0;
}
__device__ __forceinline__ int first_true8(const unsigned b0, const unsigned b1, const int shift) {
    // index of the first element i with r < c[i] (else 15): element 2j is bit j of b0, element 2j+1 bit j of b1
    const int f0 = __ffs((b0 >> shift) & 0xffu), f1 = __ffs((b1 >> shift) & 0xffu);   // f1 >= 1: element 15 is forced
    const int e0 = f0 ? 2 * f0 - 2 : 99, e1 = 2 * f1 - 1;
    return min(e0, e1);
}

__device__ __forceinline__ int nibble_core8(St &s, const Next &nx, const G2 g) {
    const int li = g.l16;
    uint32_t cp = ld_u32(reinterpret_cast<const char *>(nx.cdf) + 4 * li);
    int maxv = ld_s16g(reinterpret_cast<const char *>(nx.cdf) + 30);
    bool fresh = false;                                  // the prior is tagged for an older stream: default CDF
    if (nx.tag != nullptr) { fresh = ld_u8g(nx.tag) != s.gen; if (fresh) { cp = default_pair(li); maxv = 64; } }
    __syncwarp();
    const int c0 = (int)(short)(cp & 0xffffu), c1 = (int)(short)(cp >> 16);
    const int inc = (int)(short)(nx.speed & 0xffff), lim = nx.speed >> 16;
    int sym, start, freq;
    if (!__any_sync(FULL, nx.cdf2 != nullptr)) {
        coder_fill(s.cur);
        const int off = (int)(s.cur.a & 0x7fff);
        const int r = (int)(short)((off * maxv) >> 15);                     // probability/interface.rs:140
        const unsigned b0 = __ballot_sync(FULL, r < c0), b1 = __ballot_sync(FULL, (li == 7) || (r < c1));
        sym = first_true8(b0, b1, g.shift);
        int hi, lo;
        cum_pair(cdf_div(c0, maxv), cdf_div(c1, maxv), sym, hi, lo);
        start = (int)(short)(lo + 1); freq = (int)(short)(hi - lo - 1);   // "major hax", probability/interface.rs:103-104
        coder_advance(s.cur, start, freq);
        const int n0 = blend_elem(c0, 2 * li, maxv, sym, inc, lim), n1 = blend_elem(c1, 2 * li + 1, maxv, sym, inc, lim);
        *reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(nx.cdf) + 4 * li) = ((uint32_t)n0 & 0xffffu) | ((uint32_t)n1 << 16);
        if (fresh && g.store0 && nx.speed != SPK_NONE) *nx.tag = (uint8_t)s.gen;
        return sym;
    }
    // ---- at least one group mixes two priors (dynamic context mixing >= 2, codec/literal.rs:219-243) ----
    const bool mixg = nx.cdf2 != nullptr;
    int cc0 = c0, cc1 = c1, mc = maxv;
    if (mixg) {
        const uint32_t q = ld_u32(reinterpret_cast<const char *>(nx.cdf2) + 4 * li);
        cc0 = (int)(short)(q & 0xffffu); cc1 = (int)(short)(q >> 16); mc = ld_s16g(reinterpret_cast<const char *>(nx.cdf2) + 30);
    }
    Weights w = nx.mix_hi ? s.c->w_hi : s.c->w_lo;
    const int prod = mc * maxv;
    int lz = prod == 0 ? 32 : __clz(prod); if (lz > 17) lz = 17;
    const int shift = 17 - lz;
    const int mixr = w.norm, inv = (1 << 15) - mixr;
    // frequentist_cdf.rs:58-72
    const int ca0 = (int)(short)((int)((unsigned)((cc0 * maxv) >> shift) * (unsigned)mixr + (unsigned)((c0 * mc) >> shift) * (unsigned)inv + 1u) >> 15);
    const int ca1 = (int)(short)((int)((unsigned)((cc1 * maxv) >> shift) * (unsigned)mixr + (unsigned)((c1 * mc) >> shift) * (unsigned)inv + 1u) >> 15);
    const int ma = __shfl_sync(FULL, ca1, 7, 8);
    const int cu0 = mixg ? ca0 : c0, cu1 = mixg ? ca1 : c1, mu = mixg ? ma : maxv;
    coder_fill(s.cur);
    const int off = (int)(s.cur.a & 0x7fff);
    const int r = (int)(short)((off * mu) >> 15);
    const unsigned b0 = __ballot_sync(FULL, r < cu0), b1 = __ballot_sync(FULL, (li == 7) || (r < cu1));
    sym = first_true8(b0, b1, g.shift);
    int hi, lo;
    cum_pair(cdf_div(cu0, mu), cdf_div(cu1, mu), sym, hi, lo);
    start = (int)(short)(lo + 1); freq = (int)(short)(hi - lo - 1);
    int h2, l2;
    cum_pair(cdf_div(cc0, mc), cdf_div(cc1, mc), sym, h2, l2);
    const int f_cm = (int)(short)(h2 - l2 - 1);
    cum_pair(cdf_div(c0, maxv), cdf_div(c1, maxv), sym, h2, l2);
    const int f_nb = (int)(short)(h2 - l2 - 1);
    coder_advance(s.cur, start, freq);
    if (mixg) {
        weights_update(w, f_cm, f_nb, freq);
        if (nx.mix_hi) s.c->w_hi = w; else s.c->w_lo = w;
        const int sp = nx.mix_hi ? s.c->ad_cm_hi : s.c->ad_cm_lo;
        const int ci = (int)(short)(sp & 0xffff), cl = sp >> 16;
        const int m0 = blend_elem(cc0, 2 * li, mc, sym, ci, cl), m1 = blend_elem(cc1, 2 * li + 1, mc, sym, ci, cl);
        *reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(nx.cdf2) + 4 * li) = ((uint32_t)m0 & 0xffffu) | ((uint32_t)m1 << 16);
    }
    const int n0 = blend_elem(c0, 2 * li, maxv, sym, inc, lim), n1 = blend_elem(c1, 2 * li + 1, maxv, sym, inc, lim);
    *reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(nx.cdf) + 4 * li) = ((uint32_t)n0 & 0xffffu) | ((uint32_t)n1 << 16);
    if (fresh && g.store0 && nx.speed != SPK_NONE) *nx.tag = (uint8_t)s.gen;
    return sym;
}

// ---------------------------------------------------------------------------------------------------------------
// literal context table: T2[byte * 8 + class] = literal_context_map[(btype << 6) + (lut0[byte] | class)], class = lut1[previous
// previous byte] in 0..7 (codec/literal.rs:87-117, codec/interface.rs:199-238).  Rebuilt by the group when the
// prediction mode, the context map or the literal block type changed.
// ---------------------------------------------------------------------------------------------------------------
static __device__ __noinline__ void build_t2(const G2 g, uint8_t *slot, const uint8_t *tables, uint32_t pred_mode, uint32_t btype_last) {
    const uint8_t *lut0 = tables + TB_CTX + 512 * pred_mode;
    const uint8_t *lcm = slot + OFF_LCM + (btype_last << 6);
    uint32_t *t2 = reinterpret_cast<uint32_t *>(slot + OFF_T2);
    for (uint32_t w = (uint32_t)g.l16; w < 512; w += (uint32_t)g.nl) {      // word w holds classes 4*(w&1) .. +3 of byte w >> 1
        const uint32_t a = lut0[w >> 1], k = (w & 1) * 4;
        t2[w] = (uint32_t)lcm[a | k] | ((uint32_t)lcm[a | (k + 1)] << 8) | ((uint32_t)lcm[a | (k + 2)] << 16) | ((uint32_t)lcm[a | (k + 3)] << 24);
    }
    __syncwarp(g.gmask);
}

// under-estimate of 2^32 / d for 16 <= d < 2^15 (relative error in (0, 2^-17]): quotients by multiply-high need ONE fix-up
__device__ __forceinline__ uint32_t recip32(const int d) {
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(__uint2float_rz((uint32_t)d)));
    return __float2uint_rz(rc * 4294965248.0f);   // 2^32 * (1 - 2^-21): rcp.approx is within 2^-22 relative
}
// floor((c << 15) / d) for 0 <= c <= d
__device__ __forceinline__ uint32_t divq(const uint32_t c, const uint32_t inv, const uint32_t d) {
    const uint32_t n = c << 15;
    uint32_t q = __umulhi(n, inv);
    if (n - q * d >= d) q++;
    return q;
}

// One literal nibble of the fast loop after its symbol is known: start/freq, rANS step with eager refill, packed blend,
// store.  `cpv`/`mv`: the (validated) prior; `st`: the rANS state that codes this nibble position.
__device__ __forceinline__ void finish8(uint64_t &st, const uint32_t cpv, const int mv, const int sym, const bool fresh, char *const p, uint8_t *const tagp,
                                        const uint32_t gen, const G2 g, const uint32_t *const wbase, uint32_t &wi, const uint32_t wmax, uint32_t &wnext,
                                        const uint32_t incp, const int inc, const int lim, const uint32_t kp) {
    const uint32_t inv = recip32(mv);
    const uint32_t cum = divq(cpv & 0xffffu, inv, (uint32_t)mv) | (divq(cpv >> 16, inv, (uint32_t)mv) << 16);   // element 15: 0x8000
    const int prev = (sym - 1) & 15;
    const uint32_t whi = __shfl_sync(FULL, cum, sym >> 1, 8), wlo = __shfl_sync(FULL, cum, prev >> 1, 8);
    const uint32_t hi = (sym & 1) ? (whi >> 16) : (whi & 0xffffu);
    uint32_t lo = (prev & 1) ? (wlo >> 16) : (wlo & 0xffffu);
    if (sym == 0) lo = 0;
    const uint32_t start = lo + 1, freq = hi - lo - 1;                       // "major hax", probability/interface.rs:103-104
    const uint32_t t = ((uint32_t)st & 0x7fffu) - start;                     // 0 <= t < freq: the search put the offset in this bin
    uint64_t x = (uint64_t)freq * (st >> 15) + (uint64_t)t;                  // ans.rs:230-244
    if (x < (1ull << 31)) {                                                  // eager refill (see dv_core.cuh literal_fast)
        x = (x << 32) | (uint64_t)wnext;
        wi = min(wi + 1, wmax);
        wnext = wbase[wi];                                                   // consumed by the NEXT refill
    }
    st = x;
    const int d = sym - 2 * g.l16;                                            // elements >= sym take the increment
    const uint32_t m = d <= 0 ? 0xffffffffu : (d == 1 ? 0xffff0000u : 0u);
    uint32_t c2 = cpv + (incp & m);
    if (mv + inc >= lim) { const uint32_t u = c2 + kp; c2 = u - ((u >> 2) & 0x3fff3fffu); }   // frequentist_cdf.rs:79-84 on both halves
    *reinterpret_cast<uint32_t *>(p + 4 * g.l16) = c2;
    if (fresh && g.store0) *tagp = (uint8_t)gen;
}

// bin search on a packed pair: number of elements with r < c[i] is 16 - sym for a monotone CDF whose last element is max
__device__ __forceinline__ int search8(const uint64_t st, const uint32_t cpv, const int mv, const uint32_t bsel) {
    const int rr = ((int)((uint32_t)st & 0x7fffu) * mv) >> 15;              // probability/interface.rs:140
    const unsigned b0 = __ballot_sync(FULL, rr < (int)(cpv & 0xffffu)), b1 = __ballot_sync(FULL, rr < (int)(cpv >> 16));
    return 16 - (__popc(__byte_perm(b0, b1, bsel)) >> 1);                    // bsel picks the group's byte of b0 and of b1, twice
}

// Converged literal fast path, four streams per warp: code_nibble_array (codec/literal.rs:261-394) for whole bytes.
// `active`: this group really is at the start of a literal byte.  A group that has run out of streams rides along as a
// dummy (it codes garbage against its own slot and stores no output) so that its warp-mates keep the fast loop.
__device__ __forceinline__ void literal_fast8(St &s, Next &nx, const G2 g, const bool active) {
    uint32_t n = active ? s.lit_left : 0xffffffffu;
    n = min(n, __shfl_xor_sync(FULL, n, 8)); n = min(n, __shfl_xor_sync(FULL, n, 16));
    // plain literals, one mixing value for the whole map (not the never-adapted flat prior), speeds that cannot wrap i16
    if (__all_sync(FULL, !active || (!s.mixing_trait && s.lit_cfg >= 0 && !(s.lit_cfg & 0x800) && s.speeds_small))) {
        if (active && s.c->t2_dirty) { build_t2(g, s.slot, s.tables, s.pred_mode, s.btype_last); s.c->t2_dirty = false; }
        __syncwarp();
        const int li = g.l16;
        const int cfg = active ? s.lit_cfg : mm_cfg(4);
        const uint32_t mm = (cfg & 0x100) ? 0xffu : 0u, o1 = (cfg & 0x200) ? 0xfu : 0u, fc = (cfg & 0x400) ? 0xfu : 0u;
        const uint32_t sh = (uint32_t)(cfg >> 2) & 63u, which = (uint32_t)cfg & 3u;
        const int inc = active ? (int)(short)(s.ad_stride & 0xffff) : 0x10, lim = active ? (s.ad_stride >> 16) : 0x2000;
        const uint32_t incp = (uint32_t)inc * 0x10001u, kp = (uint32_t)(2 * li + 1) | ((uint32_t)(2 * li + 2) << 16);
        const uint32_t defp = default_pair(li);
        const uint32_t bsel = (uint32_t)(g.shift >> 3) * 0x1111u + 0x4040u;   // PRMT selector: bytes [g, 4+g, g, 4+g] of (b0, b1)
        char *const hi_tab = reinterpret_cast<char *>(s.slot + OFF_LIT_HI) + (size_t)which * 65536 * 32;
        char *const lo_tab = reinterpret_cast<char *>(s.slot + OFF_LIT_LO) + (size_t)which * 65536 * 32;
        uint8_t *const hi_tag = s.slot + OFF_TAGS_HI + which * 65536, *const lo_tag = s.slot + OFF_TAGS_LO + which * 65536;
        const uint8_t *const t2 = s.slot + OFF_T2;
        const uint8_t *const lut1 = s.tables + TB_CTX + 512 * (active ? s.pred_mode : 0u) + 256;
        const uint32_t gen = s.gen;
        unsigned long long l8 = active ? s.l8 : 0ull;
        uint32_t ctx = active ? s.lit_ctx : 0u;
        uint32_t pcp = __ldg(lut1 + ((uint32_t)(l8 >> 56)));                 // class of the byte before the next one to decode
        uint8_t *const dst = s.out + s.out_pos;
        const uint32_t pos0 = s.out_pos + ((uint32_t)(uintptr_t)s.out & 7u);   // alignment of the 8-byte stores is that of the ADDRESS
        const bool st_lane = g.store0 && active;
        Coder k = s.cur;
        if (!active) { k.p = reinterpret_cast<const uint32_t *>(s.slot + OFF_T2); k.left = 0; k.need_a = 0; k.need_b = 0; k.sym_count = 0; k.a = k.b = 1ull << 40; }
        // ---- eager-refill coder (dv_core.cuh literal_fast): state a codes every high nibble, b every low nibble ----
        const uint32_t *const wbase = k.p;
        const uint32_t wmax = k.left + 1;
        uint32_t wi = 0;
        coder_fill(k);                                                        // pending refill / 16-byte (re)initialisation of `a`
        wi = (uint32_t)(k.p - wbase);
        if (k.need_b) { k.b = (k.b << 32) | (uint64_t)wbase[wi]; wi = min(wi + 1, wmax); k.need_b = 0; }
        uint32_t wnext = wbase[wi];
        uint32_t done = 0;
        while (done < n) {
            uint32_t m = n - done;
            if (k.sym_count >= NUM_SYMBOLS_BEFORE_FLUSH) {   // chunk restart, ans.rs:173-189
                if (wi + 5 <= wmax) { k.a = (uint64_t)wbase[wi] | ((uint64_t)wbase[wi + 1] << 32); k.b = (uint64_t)wbase[wi + 2] | ((uint64_t)wbase[wi + 3] << 32); wi += 4; }
                else { k.a = k.b = 0; wi = wmax; }
                wnext = wbase[wi];
                k.sym_count = 0;
            }
            m = min(m, (NUM_SYMBOLS_BEFORE_FLUSH - k.sym_count) >> 1);
            m = min(m, __shfl_xor_sync(FULL, m, 8)); m = min(m, __shfl_xor_sync(FULL, m, 16));
            if (m == 0) break;   // unreachable: the literal coder codes nibbles in pairs, sym_count stays even
            // bytes until the output cursor is 8-byte aligned leave one by one, then aligned 8-byte stores of l8, then a tail
            const uint32_t head = min(m, (8u - ((pos0 + done) & 7u)) & 7u);
            // ---- priors of the first byte ----
            uint32_t ssb = (uint32_t)(l8 >> sh) & 0xffu;
            uint32_t idx_h = ctx * 256u + (ssb & mm & (~o1 & 0xffu));
            uint32_t row_l = ((ctx & o1) << 12) | (((mm & ssb) | ((~mm & 0xffu) & ctx)) << 4);
            char *ph = hi_tab + (size_t)idx_h * 32u;
            __syncwarp();
            uint32_t cph = ld_u32(ph + 4 * li); int mh = ld_s16g(ph + 30); uint32_t tgh = ld_u8g(hi_tag + idx_h);
            if (li < 4) prefetch_l1(lo_tab + (size_t)row_l * 32u + 128u * li); else if (li == 4) prefetch_l1(lo_tag + row_l);
            for (uint32_t i = 0; i < m; i++) {
                // -- high nibble: search
                const bool fresh_h = tgh != gen;
                const uint32_t cph_v = fresh_h ? defp : cph; const int mh_v = fresh_h ? 64 : mh;
                const int h = search8(k.a, cph_v, mh_v, bsel);
                // -- low nibble: prior (one of the 16 prefetched candidates)
                const uint32_t idx_l = row_l + ((uint32_t)h & fc);
                char *const pl = lo_tab + (size_t)idx_l * 32u;
                __syncwarp();
                const uint32_t cpl = ld_u32(pl + 4 * li); const int ml = ld_s16g(pl + 30); const uint32_t tgl = ld_u8g(lo_tag + idx_l);
                // -- high nibble: finish
                finish8(k.a, cph_v, mh_v, h, fresh_h, ph, hi_tag + idx_h, gen, g, wbase, wi, wmax, wnext, incp, inc, lim, kp);
                // -- low nibble: search
                const bool fresh_l = tgl != gen;
                const uint32_t cpl_v = fresh_l ? defp : cpl; const int ml_v = fresh_l ? 64 : ml;
                const int l = search8(k.b, cpl_v, ml_v, bsel);
                const uint32_t cur = ((uint32_t)l | ((uint32_t)h << 4)) & 0xffu;
                l8 = (l8 >> 8) | ((unsigned long long)cur << 56);              // push_literal_byte, codec/interface.rs:280-284
                if (st_lane) {
                    if (i < head) dst[done + i] = (uint8_t)cur;
                    else if (((pos0 + done + i) & 7u) == 7u) *reinterpret_cast<unsigned long long *>(dst + done + i - 7) = l8;
                }
                // -- context and priors of the next byte (get_prev_word_context, codec/literal.rs:87-117, through T2)
                ctx = t2[cur * 8u + pcp];
                pcp = __ldg(lut1 + cur);
                ssb = (uint32_t)(l8 >> sh) & 0xffu;
                idx_h = ctx * 256u + (ssb & mm & (~o1 & 0xffu));
                ph = hi_tab + (size_t)idx_h * 32u;
                __syncwarp();
                cph = ld_u32(ph + 4 * li); mh = ld_s16g(ph + 30); tgh = ld_u8g(hi_tag + idx_h);   // speculative on the last byte: inside the slot
                const uint32_t row_n = ((ctx & o1) << 12) | (((mm & ssb) | ((~mm & 0xffu) & ctx)) << 4);
                if (li < 4) prefetch_l1(lo_tab + (size_t)row_n * 32u + 128u * li); else if (li == 4) prefetch_l1(lo_tag + row_n);
                // -- low nibble: finish
                finish8(k.b, cpl_v, ml_v, l, fresh_l, pl, lo_tag + idx_l, gen, g, wbase, wi, wmax, wnext, incp, inc, lim, kp);
                row_l = row_n;
            }
            // tail: the bytes after the last aligned 8-byte store are still only in l8
            if (st_lane) {
                const uint32_t end = pos0 + done + m;
                uint32_t tail = end & 7u;
                if (tail > m - head) tail = m - head;                          // (fewer than 8 bytes after the head: all of them)
                for (uint32_t t = 0; t < tail; t++) dst[done + m - tail + t] = (uint8_t)(l8 >> (8 * (8 - tail + t)));
            }
            done += m;
            if (active) k.sym_count += 2 * m;
        }
        if (!active) return;
        // back to the lazy representation the state machine uses
        if (wi >= wmax) { k.underflow = 1; wi = wmax - 1; }
        k.p = wbase + wi; k.left = wmax - 1 - wi;
        k.need_a = (k.sym_count >= NUM_SYMBOLS_BEFORE_FLUSH) ? 8u : 0u; k.need_b = 0;
        s.cur = k; s.l8 = l8; s.lit_ctx = ctx; s.out_pos += done; s.lit_left -= done;
        enter_lit_nibble<false, true, true>(s, nx);
        return;
    }
    // everything else (dynamic context mixing, per-context mixing values, the flat prior, wide speeds): the generic core
    // (a dummy group codes against its slot's dummy CDF, like an idle group of the main loop)
    for (uint32_t i = 0; i < n; i++) {
        __syncwarp();
        const int h = nibble_core8(s, nx, g);
        if (active) { s.lit_h = (uint32_t)h; enter_lit_nibble<false, false, true>(s, nx); }
        __syncwarp();
        const int l = nibble_core8(s, nx, g);
        if (active) {
            const uint32_t cur = ((uint32_t)l | ((uint32_t)h << 4)) & 0xff;
            s.l8 = (s.l8 >> 8) | ((unsigned long long)cur << 56);   // push_literal_byte, codec/interface.rs:280-284
            if (g.store0) s.out[s.out_pos] = (uint8_t)cur;
            s.out_pos++;
            s.lit_left--;
            lit_context(s);
            enter_lit_nibble<false, true, true>(s, nx);
        }
    }
}

}  // namespace dv
