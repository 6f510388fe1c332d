// dv_codec.cuh -- the command codec (reference L2: src/codec/*) run inside the warp loop, shared by decoder & encoder.
#pragma once
#include "dv_model.cuh"

namespace dv {

__device__ __forceinline__ uint32_t bitlen32(uint32_t v) { return v ? 32u - (uint32_t)__clz((int)v) : 0u; }
__device__ __forceinline__ uint32_t round_up_mod_4(uint32_t v) { return ((((v - 1u) & 0xffu) | 3u) + 1u) & 0xffu; }   // codec/interface.rs:180-182 (u8)

// Per-stream state.  Scalars are uniform across the group's lanes.
struct Stream {
    // arena pointers
    int16_t *lit_hi, *lit_lo, *lit_cm, *ctype_slabs, *dprior_slabs, *misc;
    uint8_t *lcm, *mix, *dcm;
    uint32_t *bitmaps;        // shared memory: [0..47] literal slabs, [48..55] ctype, [56..63] dprior, [64] flags
    uint8_t *scratch;         // shared memory, 64 B (dictionary word)
    const uint8_t *tables;
    // coders
    Coder cmd, lit;
    // CrossCommandBookKeeping (codec/interface.rs:142-168, ctor :348-402)
    uint32_t distance_lru[4];
    uint32_t btype_lru[3][2];
    uint32_t btype_max_seen[3];
    uint32_t last_dlen, last_clen, last_llen, last_4_states;
    int cmap_lru;             // lane i (<13) holds cmap_lru[i]  (context-map LRU, lane-parallel move-to-front)
    // LiteralBookKeeping (codec/interface.rs:125-140)
    unsigned long long last_8;
    uint32_t btype_last;
    uint32_t pred_mode;
    Speed2 adapt[4];
    Weights mw[2];
    uint32_t mixing_param;
    bool mixing_trait;
    bool lit_slabs_ready;
    // output ("ring buffer": the window is the output buffer itself, cmd_to_raw/mod.rs)
    uint8_t *out;
    uint64_t out_pos, out_cap;
    uint32_t ring_len;
    int status;
    // encoder wishes (CrossCommandBookKeeping desired_*)
    uint32_t desired_context_mixing, desired_prior_depth, desired_force_stride;
    bool desired_do_context_map, have_desired_adapt;
    Speed2 desired_adapt[4];
};

// The one worker every command-stream nibble goes through.  Deliberately NOT inlined: the command interpreter is
// branchy, cold relative to the literal loop, and keeping it out of line keeps the kernel (and ptxas time) small.
template <bool ENC>
static __device__ __noinline__ int cmd_nibble(Stream &s, const Grp g, int16_t *cdf, int sym_in, int inc, int lim) {
    Coder k = s.cmd;
    int sym = code_prior<ENC>(k, g, cdf, sym_in, inc, lim);
    s.cmd = k;
    return sym;
}

__device__ __forceinline__ void store_default_slab(const Grp g, int16_t *base, uint32_t n_cdfs) {
    // default CDF [4,8,...,64] (probability/frequentist_cdf.rs:17-23): two 16-byte halves
    const uint4 lo = make_uint4(0x00080004u, 0x0010000cu, 0x00180014u, 0x0020001cu);
    const uint4 hi = make_uint4(0x00280024u, 0x0030002cu, 0x00380034u, 0x0040003cu);
    uint4 *p = reinterpret_cast<uint4 *>(base);
    uint32_t n16 = n_cdfs * 2;
    if (g.writer) for (uint32_t i = g.l16; i < n16; i += 16) p[i] = (i & 1) ? hi : lo;
}
__device__ __forceinline__ bool bm_test(const uint32_t *bm, uint32_t i) { return (bm[i >> 5] >> (i & 31)) & 1u; }
__device__ __forceinline__ void bm_set(const Grp g, uint32_t *bm, uint32_t i) {
    __syncwarp(g.mask);
    if (g.store0) bm[i >> 5] |= 1u << (i & 31);
    __syncwarp(g.mask);
}
static __device__ __noinline__ int16_t *ctype_slab(Stream &s, const Grp g, uint32_t ctype) {
    int16_t *p = s.ctype_slabs + (size_t)ctype * 32 * 16;
    if (!bm_test(s.bitmaps + 48, ctype)) { store_default_slab(g, p, 32); bm_set(g, s.bitmaps + 48, ctype); }
    return p;
}
static __device__ __noinline__ int16_t *dprior_slab(Stream &s, const Grp g, uint32_t prior) {
    int16_t *p = s.dprior_slabs + (size_t)prior * 32 * 16;
    if (!bm_test(s.bitmaps + 56, prior)) { store_default_slab(g, p, 32); bm_set(g, s.bitmaps + 56, prior); }
    return p;
}
__device__ __forceinline__ uint32_t get_distance_prior(Stream &s, uint32_t copy_len) {
    // codec/interface.rs:426-430
    uint32_t dtype = s.btype_lru[2][0];
    uint32_t m = copy_len < 2 ? 2 : copy_len;
    m -= 2; if (m > 3) m = 3;
    return s.dcm[dtype * 4 + m];
}

// Make every literal-prior slab that the current context map / mixing mask can reach hold default CDFs.
// Run lazily before the first literal after stream start or after a PredictionMode command.
static __device__ __noinline__ void ensure_literal_slabs(Stream &s, const Grp g) {
    if (s.lit_slabs_ready) return;
    // max context value over the whole 16384-entry map (entries beyond the transmitted count keep stale values and
    // stay reachable through literal block switches: codec/interface.rs:296,309)
    uint32_t mx = 0;
    const uint4 *m4 = reinterpret_cast<const uint4 *>(s.lcm);
    for (uint32_t i = g.l16; i < 1024; i += 16) {
        uint4 v = m4[i];
        uint32_t t = __vmaxu4(__vmaxu4(v.x, v.y), __vmaxu4(v.z, v.w));
        mx = __vmaxu4(mx, t);
    }
    mx = max(max(mx & 0xff, (mx >> 8) & 0xff), max((mx >> 16) & 0xff, mx >> 24));
    uint32_t present = 0;   // bit v set if mixing value v occurs
    const uint4 *x4 = reinterpret_cast<const uint4 *>(s.mix);
    for (uint32_t i = g.l16; i < 512; i += 16) {
        uint4 v = x4[i];
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            present |= 1u << (w[j] & 15); present |= 1u << ((w[j] >> 8) & 15);
            present |= 1u << ((w[j] >> 16) & 15); present |= 1u << ((w[j] >> 24) & 15);
        }
    }
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) {
        mx = max(mx, __shfl_xor_sync(g.mask, mx, o, 16));
        present |= __shfl_xor_sync(g.mask, present, o, 16);
    }
    // which: mm 0,3 -> 0 ; mm 1 -> 2 ; everything else -> 1   (codec/literal.rs:184-208)
    bool w0 = present & 0x9u, w2 = present & 0x2u, w1 = present & ~0xBu;
    for (uint32_t which = 0; which < 3; which++) {
        if (!(which == 0 ? w0 : (which == 1 ? w1 : w2))) continue;
        uint32_t lo_max = which == 2 ? (min(mx, 15u) << 4 | 15u) : 15u;
        for (uint32_t c = 0; c <= mx; c++) {
            uint32_t id = which * 256 + c;
            if (!bm_test(s.bitmaps, id)) { store_default_slab(g, s.lit_hi + (size_t)id * 256 * 16, 256); bm_set(g, s.bitmaps, id); }
        }
        for (uint32_t c = 0; c <= lo_max; c++) {
            uint32_t id = 768 + which * 256 + c;
            if (!bm_test(s.bitmaps, id)) { store_default_slab(g, s.lit_lo + (size_t)(which * 256 + c) * 256 * 16, 256); bm_set(g, s.bitmaps, id); }
        }
    }
    if (s.mixing_trait && !(s.bitmaps[64] & 1u)) {   // lit_cm_priors allocated on first mixing>=2 (codec/interface.rs:322-329)
        store_default_slab(g, s.lit_cm, (uint32_t)LIT_CM_CDFS);
        __syncwarp(g.mask);
        if (g.store0) s.bitmaps[64] |= 1u;
    }
    __syncwarp(g.mask);
    s.lit_slabs_ready = true;
}

// ---- literal length (codec/literal.rs:561-661) ----
template <bool ENC>
static __device__ __noinline__ void code_literal_len(Stream &s, const Grp g, uint32_t &len_io, uint32_t &high_entropy_io) {
    uint32_t literal_len = len_io;
    uint32_t serialized_large = literal_len - 15u;
    uint32_t lllen = bitlen32(serialized_large);
    int16_t *slab = ctype_slab(s, g, s.btype_lru[1][0]);
    bool he_flag = false;
    for (;;) {
        uint32_t lm1 = literal_len - 1u;
        int nib = (int)(lm1 < 14 ? lm1 : 14);
        if (ENC && high_entropy_io && !he_flag) nib = 15;
        nib = cmd_nibble<ENC>(s, g, slab + CT_LL_COUNT_SMALL * 16, nib, DV_SPEED_MED);
        if (nib == 14) break;
        if (nib == 15) { high_entropy_io = 1; he_flag = true; if (!ENC && s.cmd.in.underflow) return; continue; }
        len_io = (uint32_t)nib + 1; s.last_llen = len_io;
        return;
    }
    int beg = (int)(lllen < 15 ? lllen : 15);
    beg = cmd_nibble<ENC>(s, g, slab + CT_LL_SIZE_BEG * 16, beg, DV_SPEED_MUD);
    uint32_t len_remaining, decoded;
    if (beg == 15) {
        int last = (int)((lllen - 15u) & 0xf);
        last = cmd_nibble<ENC>(s, g, slab + CT_LL_SIZE_LAST * 16, last, DV_SPEED_MUD);
        len_remaining = round_up_mod_4((uint32_t)last + 14);
        decoded = 1u << (last + 14);
    } else if (beg <= 1) {
        len_io = 15u + (uint32_t)beg;   // last_llen deliberately not updated (literal.rs:608-616)
        return;
    } else {
        len_remaining = round_up_mod_4((uint32_t)beg - 1);
        decoded = 1u << (beg - 1);
    }
    for (;;) {
        uint32_t next_rem = len_remaining - 4;
        int nib = (int)(((serialized_large ^ decoded) >> next_rem) & 0xff);
        if (ENC) nib &= 0xf;
        nib = cmd_nibble<ENC>(s, g, slab + CT_LL_MANTISSA * 16, nib, DV_SPEED_MUD);
        decoded |= (uint32_t)nib << next_rem;
        if (next_rem == 0) break;
        len_remaining = next_rem;
    }
    len_io = decoded + 15u; s.last_llen = len_io;
}

// ---- one literal nibble (codec/literal.rs:154-259) ----
// Hot-loop state lives in a struct of scalars (no arrays, address never escapes) so that it stays in registers.
struct LitCtx {
    Coder k;
    int16_t *lit_hi, *lit_lo, *lit_cm;
    const uint8_t *mix;
    Speed2 ad_stride, ad_cm_lo, ad_cm_hi;   // literal_adaptation[0], [2], [3]
    Weights w_lo, w_hi;                     // model_weights[0], [1]
};
template <bool ENC, bool MIX, bool HIGH>
__device__ __forceinline__ int code_lit_nibble(LitCtx &L, const Grp g, int nib_in, uint32_t ctx, uint32_t prev_byte,
                                               unsigned long long stride_bytes, uint32_t cur_byte_prior) {
    uint32_t mmi = ctx | (HIGH ? ((prev_byte >> 4) << 8) : (((cur_byte_prior & 0xf) << 8) | 4096u));
    uint32_t mm_opts = L.mix[mmi];
    uint32_t fast_cm_prior_mask = (mm_opts != 3) ? 0xffu : 0u;
    uint32_t mm = (mm_opts != 0 && mm_opts != 3) ? 0xffu : 0u;
    uint32_t opt_1_f_mask = (mm_opts == 1) ? 0xfu : 0u;
    uint32_t stride_offset = mm_opts < 4 ? 0u : (min(7u, mm_opts ^ 4u) << 3);
    uint32_t ssb = (uint32_t)(stride_bytes >> (0x38 - stride_offset)) & 0xffu;
    uint32_t index_b, index_c;
    if (HIGH) { index_b = ssb & mm & (~opt_1_f_mask & 0xffu); index_c = ctx; }
    else { index_b = (mm & ssb) | ((~mm & 0xffu) & ctx); index_c = (cur_byte_prior & fast_cm_prior_mask) | ((ctx & opt_1_f_mask) << 4); }
    uint32_t which = (mm >> 7) ^ (opt_1_f_mask >> 2);
    int16_t *np = (HIGH ? L.lit_hi : L.lit_lo) + ((size_t)((which * 256 + index_c) * 256 + index_b)) * 16;
    int sym, start, freq;
    if (MIX) {
        int16_t *cp = HIGH ? L.lit_cm + (size_t)ctx * 16 : L.lit_cm + (size_t)(256 + cur_byte_prior + 16 * ctx) * 16;
        int cs = np[g.l16], ms = np[15];      // stride prior
        int cc = cp[g.l16], mc = cp[15];      // context-map prior
        Weights &w = HIGH ? L.w_hi : L.w_lo;
        // average(cm, stride, w)  (probability/frequentist_cdf.rs:58-72)
        int prod = mc * ms;
        int lz = prod == 0 ? 32 : __clz(prod); if (lz > 17) lz = 17;
        int shift = 17 - lz;
        int mixr = w.norm, inv = (1 << 15) - mixr;
        int rs = (cc * ms) >> shift, ro = (cs * mc) >> shift;
        int ca = (int)(short)((int)((unsigned)rs * (unsigned)mixr + (unsigned)ro * (unsigned)inv + 1u) >> 15);
        int ma = __shfl_sync(g.mask, ca, 15, 16);
        sym = code_cdf<ENC>(L.k, g, ca, ma, nib_in, start, freq);
        int f_cm = cdf_freq(g, cc, mc, sym);
        int f_nb = cdf_freq(g, cs, ms, sym);
        weights_update(w, f_cm, f_nb, freq);
        const Speed2 sp = HIGH ? L.ad_cm_hi : L.ad_cm_lo;
        int c2 = cdf_blend(g, cc, mc, sym, sp.inc, sp.lim);
        if (g.writer) cp[g.l16] = (int16_t)c2;
        if (mm_opts != 2) {
            int s2 = cdf_blend(g, cs, ms, sym, L.ad_stride.inc, L.ad_stride.lim);
            if (g.writer) np[g.l16] = (int16_t)s2;
        }
    } else {
        if (mm_opts == 2) {   // flat default CDF, no adaptation (literal.rs:213-216,252-256)
            sym = code_cdf<ENC>(L.k, g, 4 * (g.l16 + 1), 64, nib_in, start, freq);
        } else {
            int c = np[g.l16], maxv = np[15];
            sym = code_cdf<ENC>(L.k, g, c, maxv, nib_in, start, freq);
            int c2 = cdf_blend(g, c, maxv, sym, L.ad_stride.inc, L.ad_stride.lim);
            if (g.writer) np[g.l16] = (int16_t)c2;
        }
    }
    __syncwarp(g.mask);
    return sym;
}

// last_8_literals re-seed (codec/decoder.rs:361-375, cmd_to_raw/mod.rs:69-86 incl. the order flip when index<8)
__device__ __forceinline__ unsigned long long reseed_last8(Stream &s) {
    uint32_t idx = (uint32_t)(s.out_pos & (s.ring_len - 1));
    unsigned long long v = 0;
    if (idx < 8) {
        for (uint32_t i = 0; i < 8; i++) {   // ret[i] = ring[(idx - i - 1) mod len]  -> packed ret[0] | ret[1]<<8 ...
            long long p = (long long)s.out_pos - 1 - (long long)i;
            unsigned long long b = 0;
            if (p >= 0) b = s.out[p];
            v |= b << (8 * i);
        }
    } else {
        for (uint32_t i = 0; i < 8; i++) v |= (unsigned long long)s.out[s.out_pos - 8 + i] << (8 * i);
    }
    return v;
}

// ---- literal content bytes (codec/literal.rs:261-394) ----
template <bool ENC, bool MIX>
static __device__ __noinline__ void code_literal_bytes(Stream &s, const Grp g, const uint8_t *src, uint32_t len) {
    unsigned long long l8 = reseed_last8(s);
    const uint8_t *lut = c_ctx_lut + 512 * (s.pred_mode & 3);
    const uint8_t *lcm = s.lcm + (s.btype_last << 6);
    uint8_t *dst = s.out + s.out_pos;
    LitCtx L;
    L.k = s.lit;
    L.lit_hi = s.lit_hi; L.lit_lo = s.lit_lo; L.lit_cm = s.lit_cm; L.mix = s.mix;
    L.ad_stride = s.adapt[0]; L.ad_cm_lo = s.adapt[2]; L.ad_cm_hi = s.adapt[3];
    L.w_lo = s.mw[0]; L.w_hi = s.mw[1];
    int st = ST_OK;
    for (uint32_t i = 0; i < len; i++) {
        uint32_t prev = (uint32_t)(l8 >> 56), pp = (uint32_t)(l8 >> 48) & 0xff;
        uint32_t ctx = lcm[lut[prev] | lut[256 + pp]];
        int byte_in = ENC ? (int)src[i] : 0;
        int h = code_lit_nibble<ENC, MIX, true>(L, g, byte_in >> 4, ctx, prev, l8, 0);
        int l = code_lit_nibble<ENC, MIX, false>(L, g, byte_in & 0xf, ctx, prev, l8, (uint32_t)h);
        uint32_t cur = (uint32_t)(l | (h << 4));
        l8 = (l8 >> 8) | ((unsigned long long)cur << 56);
        if (g.store0) dst[i] = (uint8_t)cur;
        if (!ENC && L.k.in.underflow) { st = ST_NEED_INPUT; break; }
    }
    __syncwarp(g.mask);
    s.lit = L.k; s.mw[0] = L.w_lo; s.mw[1] = L.w_hi;
    if (st != ST_OK) s.status = st;
    s.out_pos += len;
    s.last_8 = l8;
}

// ---- copy command fields (codec/copy.rs:50-286) ----
__device__ __forceinline__ void distance_from_mnemonic(const uint32_t lru[4], uint32_t code, uint32_t &dist, bool &ok) {
    // codec/interface.rs:979-1009
    if (code < 4) { dist = lru[code]; ok = true; return; }
    int us = (int)(code >> 2);
    int ss = us - (((-(int)(code & 1)) & us) << 1);
    int ret = (int)lru[(code & 2) >> 1] + ss;
    dist = (uint32_t)ret; ok = ret > 0;
}
template <bool ENC>
static __device__ __noinline__ void code_copy(Stream &s, const Grp g, uint32_t &dist_io, uint32_t &num_io) {
    uint32_t in_dist = dist_io, in_num = num_io;
    uint32_t dlen = bitlen32(in_dist), clen = bitlen32(in_num);
    if (ENC && dlen == 0) { s.status = ST_FAIL; return; }
    int16_t *cslab = ctype_slab(s, g, s.btype_lru[1][0]);
    uint32_t num_bytes, distance;
    {
        uint32_t ll = s.last_llen - 1u; if (ll > 3) ll = 3;
        uint32_t index = ((s.last_4_states >> 4) & 3u) + 4u * ll;
        int nib = (int)(in_num < 15 ? in_num : 15);
        nib = cmd_nibble<ENC>(s, g, cslab + (CT_CP_COUNT_SMALL + index) * 16, nib, DV_SPEED_MUD);
        if (nib != 15) {
            num_bytes = (uint32_t)nib; s.last_clen = bitlen32(num_bytes);
        } else {
            int beg = (int)min(15u, (clen - 4u) & 0xffu);
            beg = cmd_nibble<ENC>(s, g, cslab + CT_CP_COUNT_BEG * 16, beg, DV_SPEED_FAST);
            uint32_t len_remaining, decoded;
            if (beg == 15) {
                int last = (int)((clen - 19u) & 0xf);
                last = cmd_nibble<ENC>(s, g, cslab + CT_CP_COUNT_LAST * 16, last, DV_SPEED_FAST);
                s.last_clen = (uint32_t)last + 19;
                len_remaining = round_up_mod_4((uint32_t)last + 18);
                decoded = (last + 18) < 32 ? (1u << (last + 18)) : 0u;
            } else {
                s.last_clen = (uint32_t)beg + 4;
                len_remaining = round_up_mod_4((uint32_t)beg + 3);
                decoded = 1u << (beg + 3);
            }
            uint32_t len_decoded = 0;
            for (;;) {
                uint32_t next_rem = len_remaining - 4;
                int nb = (int)(((in_num ^ decoded) >> next_rem) & 0xff);
                if (ENC) nb &= 0xf;
                uint32_t index2 = len_decoded == 0 ? ((s.last_clen % 4) + 1) : 0u;
                nb = cmd_nibble<ENC>(s, g, cslab + (CT_CP_COUNT_MANT + index2) * 16, nb, DV_SPEED_SLOW);
                decoded |= (uint32_t)nb << next_rem;
                if (next_rem == 0) break;
                len_decoded += 4; len_remaining = next_rem;
            }
            num_bytes = decoded;
        }
    }
    {
        int beg = 15;
        if (ENC) {   // distance_mnemonic_code, codec/interface.rs:469-477
            for (uint32_t i = 0; i < 15; i++) {
                uint32_t d; bool ok; distance_from_mnemonic(s.distance_lru, i, d, ok);
                if (d == in_dist && ok) { beg = (int)i; break; }
            }
        }
        uint32_t ap = get_distance_prior(s, num_bytes);
        int16_t *dslab = dprior_slab(s, g, ap);
        beg = cmd_nibble<ENC>(s, g, dslab + (DP_MNEMONIC + (s.last_llen < 8 ? 1 : 0)) * 16, beg, DV_SPEED_SLOW);
        if (beg != 15) {
            bool ok; distance_from_mnemonic(s.distance_lru, (uint32_t)beg, distance, ok);
            s.last_dlen = bitlen32(distance);
            if (!ok) { s.status = ST_FAIL; return; }
        } else {
            int bn = (int)min(14u, (dlen - 1u) & 0xffu);
            if (ENC && (s.distance_lru[1] - 3u) == in_dist) bn = 15;
            uint32_t index = bitlen32(num_bytes) >> 2;
            bn = cmd_nibble<ENC>(s, g, dslab + (DP_DIST_BEG + index) * 16, bn, DV_SPEED_SLOW);
            if (bn == 15) {
                distance = s.distance_lru[1] - 3u;
                s.last_dlen = bitlen32(distance);
            } else if (bn == 0) {
                s.last_dlen = 1; distance = 1;
            } else {
                uint32_t start_rem, decoded;
                if (bn == 14) {
                    int last = (int)((dlen - 15u) & 0xf);
                    last = cmd_nibble<ENC>(s, g, dslab + DP_DIST_LAST * 16, last, DV_SPEED_ROCKET);
                    s.last_dlen = (uint32_t)last + 15;
                    start_rem = round_up_mod_4((uint32_t)last + 14);
                    decoded = (last + 14) < 32 ? (1u << (last + 14)) : 0u;
                } else {
                    s.last_dlen = (uint32_t)bn + 1;
                    start_rem = round_up_mod_4((uint32_t)bn);
                    decoded = 1u << bn;
                }
                uint32_t len_decoded = 0;
                for (int sr2 = (int)((start_rem + 3) >> 2) - 1; sr2 >= 0; sr2--) {
                    uint32_t next_rem = (uint32_t)sr2 << 2;
                    int nb = (int)(((in_dist ^ decoded) >> next_rem) & 0xff);
                    if (ENC) nb &= 0xf;
                    uint32_t index2 = len_decoded == 0 ? ((s.last_dlen & 3) + 1) : 0u;
                    int inc = 0x4 << ((index2 & 6) << ((index2 & 2) >> 1));
                    nb = cmd_nibble<ENC>(s, g, dslab + (DP_DIST_MANT + index2) * 16, nb, inc, 0x4000);
                    decoded |= (uint32_t)nb << next_rem;
                    len_decoded += 4;
                }
                distance = decoded;
            }
        }
    }
    dist_io = distance; num_io = num_bytes;
}

// obs_distance (codec/interface.rs:509-527)
__device__ __forceinline__ void obs_distance(Stream &s, uint32_t d) {
    uint32_t *l = s.distance_lru;
    if (d == l[1]) { l[1] = l[0]; l[0] = d; }
    else if (d == l[2]) { l[2] = l[1]; l[1] = l[0]; l[0] = d; }
    else if (d != l[0]) { l[3] = l[2]; l[2] = l[1]; l[1] = l[0]; l[0] = d; }
}
__device__ __forceinline__ void obs_btype(Stream &s, int which, uint32_t bt) {
    s.last_4_states >>= 2;   // codec/interface.rs:528-532
    s.btype_lru[which][1] = s.btype_lru[which][0];
    s.btype_lru[which][0] = bt;
    if (bt > s.btype_max_seen[which]) s.btype_max_seen[which] = bt;
}

// ---- replay into the window (cmd_to_raw/mod.rs:245-283).  out[pos+i] = out[pos-dist+(i mod dist)]: the copied region
// is periodic with period `dist` and its first period already exists, so all lanes copy independently. ----
static __device__ __noinline__ void replay_copy(Stream &s, const Grp g, uint32_t dist, uint32_t len) {
    if (dist == 0 || dist >= s.ring_len) { s.status = ST_FAIL; return; }   // DistanceGreaterRingBuffer & friends
    if ((uint64_t)len > s.out_cap - s.out_pos) { s.status = ST_NEED_OUTPUT; len = (uint32_t)(s.out_cap - s.out_pos); }
    long long base = (long long)s.out_pos - (long long)dist;
    uint8_t *dst = s.out + s.out_pos;
    uint32_t off = (uint32_t)g.l16 % dist;
    uint32_t step = 16u % dist;
    if (g.writer) {
        for (uint32_t i = (uint32_t)g.l16; i < len; i += 16) {
            long long sp = base + (long long)off;
            dst[i] = sp >= 0 ? s.out[sp] : (uint8_t)0;   // a fresh ring is zero-initialised (ffi/alloc_util.rs:70-99)
            off += step; if (off >= dist) off -= dist;
        }
    }
    __syncwarp(g.mask);
    s.out_pos += len;
}

// dictionary word + RFC 7932 transform (cmd_to_raw/mod.rs:284-309; brotli crate TransformDictionaryWord, NOT-IN-TREE)
__device__ __forceinline__ int dict_word_len(const uint8_t *tables, uint32_t word_size, uint32_t transform) {
    const uint8_t *tr = tables + TB_TRANSFORMS + 3 * transform;
    const uint16_t *psmap = reinterpret_cast<const uint16_t *>(tables + TB_PSMAP);
    int plen = tables[TB_PS + psmap[tr[0]]], slen = tables[TB_PS + psmap[tr[2]]];
    int t = tr[1], len = (int)word_size;
    int skip = t < 12 ? 0 : t - 11; if (skip > len) skip = len;
    len -= skip; if (t <= 9) len -= t; if (len < 0) len = 0;
    return plen + len + slen;
}
static __device__ __noinline__ void replay_dict(Stream &s, const Grp g, uint32_t word_size, uint32_t word_id, uint32_t transform) {
    const uint8_t *tb = s.tables;
    if (word_size < 4 || word_size > 24 || transform >= 121) { s.status = ST_FAIL; return; }
    uint64_t widx = (uint64_t)word_id * word_size + reinterpret_cast<const uint32_t *>(tb + TB_OFFSETS)[word_size];
    if (widx + word_size > TB_DICT_SIZE) { s.status = ST_FAIL; return; }
    int n = 0;
    if (g.lane0) {
        const uint8_t *word = tb + TB_DICT + widx;
        const uint8_t *tr = tb + TB_TRANSFORMS + 3 * transform;
        const uint16_t *psmap = reinterpret_cast<const uint16_t *>(tb + TB_PSMAP);
        const uint8_t *prefix = tb + TB_PS + psmap[tr[0]], *suffix = tb + TB_PS + psmap[tr[2]];
        uint8_t *o = s.scratch;
        int t = tr[1], len = (int)word_size;
        for (int i = 0; i < 64; i++) o[i] = 0;
        { int pl = *prefix++; while (pl--) o[n++] = *prefix++; }
        int skip = t < 12 ? 0 : t - 11; if (skip > len) skip = len;
        word += skip; len -= skip; if (t <= 9) len -= t;
        for (int i = 0; i < len; i++) o[n++] = word[i];
        if (len > 0 && (t == 10 || t == 11)) {
            uint8_t *up = o + n - len; int rem = t == 10 ? 1 : len;
            while (rem > 0) {
                int step;
                if (up[0] < 0xc0) { if (up[0] >= 'a' && up[0] <= 'z') up[0] ^= 32; step = 1; }
                else if (up[0] < 0xe0) { up[1] ^= 32; step = 2; }
                else { up[2] ^= 5; step = 3; }
                up += step; rem -= step;
                if (t == 10) break;
            }
        }
        { int sl = *suffix++; while (sl--) o[n++] = *suffix++; }
    }
    n = __shfl_sync(g.mask, n, 0, 16);
    __syncwarp(g.mask);
    if ((uint64_t)n > s.out_cap - s.out_pos) { s.status = ST_NEED_OUTPUT; n = (int)(s.out_cap - s.out_pos); }
    if (g.writer) for (int i = g.l16; i < n; i += 16) s.out[s.out_pos + i] = s.scratch[i];
    __syncwarp(g.mask);
    s.out_pos += (uint32_t)n;
}

// ---- dict command fields (codec/dict.rs:49-176) ----
template <bool ENC>
static __device__ __noinline__ void code_dict(Stream &s, const Grp g, uint32_t &id_io, uint32_t &size_io, uint32_t &tr_io) {
    uint32_t in_id = id_io, in_size = size_io & 0xff, in_tr = tr_io & 0xff;
    int16_t *cslab = ctype_slab(s, g, s.btype_lru[1][0]);
    int beg = (int)min(15u, (in_size - 4u) & 0xffu);
    beg = cmd_nibble<ENC>(s, g, cslab + CT_DC_SIZE_BEG * 16, beg, DV_SPEED_MUD);
    uint32_t word_size;
    if (beg == 15) {
        int b2 = (int)((in_size - 19u) & 0xf);
        b2 = cmd_nibble<ENC>(s, g, cslab + CT_DC_SIZE_LAST * 16, b2, DV_SPEED_MUD);
        word_size = (uint32_t)b2 + 19;
        if (word_size > 24) { s.status = ST_FAIL; return; }   // DictWordSizeTooLarge
    } else word_size = (uint32_t)beg + 4;
    uint32_t bits = s.tables[TB_SIZE_BITS + word_size];
    uint32_t len_remaining = round_up_mod_4(bits), decoded = 0, len_decoded = 0;
    for (;;) {
        uint32_t next_rem = len_remaining - 4;
        int nb = (int)(((in_id ^ decoded) >> next_rem) & 0xff);
        if (ENC) nb &= 0xf;
        uint32_t index = len_decoded == 0 ? ((bits % 4) + 1) : 0u;
        uint32_t ap = get_distance_prior(s, word_size);
        int16_t *dslab = dprior_slab(s, g, ap);
        nb = cmd_nibble<ENC>(s, g, dslab + (DP_DICT_INDEX + index) * 16, nb, DV_SPEED_MUD);
        decoded |= (uint32_t)nb << next_rem;
        if (next_rem == 0) break;
        len_decoded += 4; len_remaining = next_rem;
    }
    int hi = (int)(in_tr >> 4);
    hi = cmd_nibble<ENC>(s, g, s.misc + (MI_TRANSFORM + 0 + 2 * (word_size >> 1)) * 16, hi, DV_SPEED_FAST);
    int lo = (int)(in_tr & 0xf);
    lo = cmd_nibble<ENC>(s, g, s.misc + (MI_TRANSFORM + 1 + 2 * hi) * 16, lo, DV_SPEED_FAST);
    uint32_t tr = ((uint32_t)hi << 4) | (uint32_t)lo;
    if (tr >= 121) { s.status = ST_FAIL; return; }            // DictTransformIndexUndefined
    id_io = decoded; size_io = word_size; tr_io = tr;
}

// ---- block switch (codec/block_type.rs:27-110) ----
template <bool ENC>
static __device__ __noinline__ uint32_t code_btype(Stream &s, const Grp g, int which, uint32_t in_bt) {
    int varint;
    if (in_bt == s.btype_lru[which][1]) varint = 0;
    else if (in_bt == ((s.btype_max_seen[which] + 1) & 0xff)) varint = 1;
    else if (in_bt <= 12) varint = (int)in_bt + 2;
    else varint = 15;
    int16_t *bt = s.misc + MI_BTYPE * 16;
    varint = cmd_nibble<ENC>(s, g, bt + (BT_MNEMONIC + which) * 16, varint, DV_SPEED_SLOW);
    if (varint == 0) return s.btype_lru[which][1];
    if (varint == 1) return (s.btype_max_seen[which] + 1) & 0xff;
    if (varint != 15) return (uint32_t)varint - 2;
    int first = (int)(in_bt & 0xf), second = (int)(in_bt >> 4);
    first = cmd_nibble<ENC>(s, g, bt + (BT_FIRST + which) * 16, first, DV_SPEED_SLOW);
    second = cmd_nibble<ENC>(s, g, bt + (BT_SECOND + which) * 16, second, DV_SPEED_SLOW);
    return ((uint32_t)second << 4) | (uint32_t)first;
}

// ---- f8 speed codec (probability/interface.rs:566-585 and the brotli crate's u16 twins) ----
__device__ __forceinline__ int u8_to_speed(uint32_t data) {
    if (data < 8) return 0;
    uint32_t log_val = (data >> 3) - 1;
    int rem = (int)(short)((data & 7) << log_val);
    return (int)(short)((short)(1 << log_val) | (rem >> 3));
}
__device__ __forceinline__ uint32_t speed_to_u8_u16(uint32_t data) {   // brotli flavour, u16 wrapping
    data &= 0xffff;
    if (data == 0) return 0;
    uint32_t length = 32 - __clz((int)data);
    uint32_t rem = (data - (1u << (length - 1))) & 0xffff;
    uint32_t mant = (((rem << 3) & 0xffff) >> (length - 1)) & 0xff;
    return ((length << 3) | mant) & 0xff;
}
__device__ __forceinline__ uint32_t speed_to_u8_i16(int data) {        // divans flavour, i16
    uint32_t u = (uint32_t)data & 0xffff;
    uint32_t length = u ? 32 - __clz((int)u) : 0;
    uint32_t mant = 0;
    if (data != 0) {
        int rem = (int)(short)(data - (short)(1 << (length - 1)));
        mant = (uint32_t)(((int)(short)(rem << 3)) >> (length - 1)) & 0xff;
    }
    return ((length << 3) | mant) & 0xff;
}
__device__ __forceinline__ uint32_t u8_to_speed_u16(uint32_t data) {
    if (data < 8) return 0;
    uint32_t log_val = (data >> 3) - 1;
    uint32_t rem = ((data & 7) << log_val) & 0xffff;
    return ((1u << log_val) | (rem >> 3)) & 0xffff;
}

// Prediction-mode input as the encoder sees it (points into the DVCL blob)
struct PredModeIn {
    uint32_t pred_mode, is_adv, has_speeds;
    const uint16_t *speeds;      // cm[2][2], stride[2][2], combined[2][2]
    uint32_t lit_map_len, dist_map_len;
    const uint8_t *lit_map, *dist_map, *mixing;
};

// lane-parallel LRU-13 with move-to-front (codec/interface.rs:439-453): lane i (<13) holds entry i
__device__ __forceinline__ void cmap_lru_touch(Stream &s, const Grp g, uint32_t val) {
    unsigned hit = (__ballot_sync(g.mask, g.l16 < 13 && (uint32_t)s.cmap_lru == val) >> g.shift) & 0x1fffu;
    int found = hit ? __ffs(hit) - 1 : 12;
    int up = __shfl_up_sync(g.mask, s.cmap_lru, 1, 16);
    if (g.l16 >= 1 && g.l16 <= found) s.cmap_lru = up;
    if (g.l16 == 0) s.cmap_lru = (int)val;
}
__device__ __forceinline__ uint32_t cmap_lru_max(Stream &s, const Grp g) {
    uint32_t v = g.l16 < 13 ? (uint32_t)s.cmap_lru : 0u;
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) v = max(v, __shfl_xor_sync(g.mask, v, o, 16));
    return v;
}
template <bool ENC>
static __device__ __noinline__ void code_context_map(Stream &s, const Grp g, int is_distance, const uint8_t *in_map, uint32_t in_len, uint8_t *out_map,
                                 uint32_t out_cap) {
    int16_t *pp = s.misc + MI_PRED * 16;
    for (uint32_t index = 0;; index++) {
        int mn;
        if (!ENC || index >= in_len) mn = 14;
        else {
            uint32_t target = in_map[index];
            unsigned hit = (__ballot_sync(g.mask, g.l16 < 13 && (uint32_t)s.cmap_lru == target) >> g.shift) & 0x1fffu;
            mn = hit ? 31 - __clz((int)hit) : 15;    // "last match wins" (context_map.rs:281-285)
            if (target == ((cmap_lru_max(s, g) + 1) & 0xff)) mn = 13;
        }
        mn = cmd_nibble<ENC>(s, g, pp + (PM_MNEMONIC + is_distance) * 16, mn, DV_SPEED_MED);
        if (mn == 14) return;
        if (!ENC && s.cmd.in.underflow) { s.status = ST_NEED_INPUT; return; }
        uint32_t val;
        if (mn == 15) {
            int msn = (ENC && index < in_len) ? (in_map[index] >> 4) : 0;
            msn = cmd_nibble<ENC>(s, g, pp + (PM_FIRST_NIBBLE + is_distance) * 16, msn, DV_SPEED_MED);
            int lsn = (ENC && index < in_len) ? (in_map[index] & 0xf) : 0;
            lsn = cmd_nibble<ENC>(s, g, pp + (PM_SECOND_NIBBLE + is_distance) * 16, lsn, DV_SPEED_MED);
            val = ((uint32_t)msn << 4) | (uint32_t)lsn;
        } else if (mn == 13) {
            val = (cmap_lru_max(s, g) + 1) & 0xff;
        } else {
            val = (uint32_t)__shfl_sync(g.mask, s.cmap_lru, mn, 16);
        }
        if (index >= out_cap) { s.status = ST_FAIL; return; }   // IndexBeyondContextMapSize
        cmap_lru_touch(s, g, val);
        if (g.store0) out_map[index] = (uint8_t)val;
    }
}

// ---- prediction mode command (codec/context_map.rs:105-428 + obs_prediction_mode_context_map, codec/interface.rs:293-321) ----
template <bool ENC>
static __device__ __noinline__ void code_predmode(Stream &s, const Grp g, const PredModeIn *in) {
    int16_t *pp = s.misc + MI_PRED * 16;
    Speed2 desired[4] = {{DV_SPEED_MUD}, {DV_SPEED_MUD}, {DV_SPEED_MUD}, {DV_SPEED_MUD}};
    if (ENC) {
        if (in->has_speeds) {   // context_map.rs:123-146
            const uint16_t *cm = in->speeds, *st = in->speeds + (s.desired_context_mixing != 0 ? 8 : 4);
            for (int k = 0; k < 2; k++) {
                uint32_t a = speed_to_u8_u16(cm[k * 2]), b = speed_to_u8_u16(cm[k * 2 + 1]);
                if (a != 0 || b != 0) { desired[2 + k].inc = u8_to_speed(a); desired[2 + k].lim = u8_to_speed(b); }
                a = speed_to_u8_u16(st[k * 2]); b = speed_to_u8_u16(st[k * 2 + 1]);
                if (a != 0 || b != 0) { desired[k].inc = u8_to_speed(a); desired[k].lim = u8_to_speed(b); }
            }
        }
        if (s.have_desired_adapt) for (int k = 0; k < 4; k++) desired[k] = s.desired_adapt[k];
    }
    s.cmap_lru = g.l16;                                             // reset_context_map_lru
    if (g.writer) for (uint32_t i = g.l16; i < 1024; i += 16) s.dcm[i] = (uint8_t)(i & 3);   // reset_distance_context_map
    __syncwarp(g.mask);
    int pm = ENC ? (int)in->pred_mode : 0;
    pm = cmd_nibble<ENC>(s, g, pp + PM_ONLY * 16, pm, DV_SPEED_MED);
    int mixnib = ENC ? (int)(s.desired_context_mixing | (in->is_adv << 3)) : 0;
    mixnib = cmd_nibble<ENC>(s, g, pp + PM_SPEED_PALETTE * 16, mixnib, DV_SPEED_MED);   // aliases SpeedPalette[0]
    uint32_t mixing_math = (uint32_t)mixnib & 3;
    bool combine = mixnib != 0;
    int pd = ENC ? (int)s.desired_prior_depth : 0;
    pd = cmd_nibble<ENC>(s, g, pp + PM_SPEED_PALETTE * 16, pd, DV_SPEED_FAST);          // aliases SpeedPalette[0]
    (void)pd;
    uint32_t f8[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    for (uint32_t index = 0; index < 16; index++) {
        uint32_t si = index >> 2, pt = index & 3;
        int nib = 0;
        if (ENC) {
            uint32_t c0 = speed_to_u8_i16(desired[si].inc), c1 = speed_to_u8_i16(desired[si].lim);
            nib = (int)(pt == 0 ? ((c0 & 0x7f) >> 3) : pt == 1 ? (c0 & 7) : pt == 2 ? ((c1 & 0x7f) >> 3) : (c1 & 7));
        }
        nib = cmd_nibble<ENC>(s, g, pp + (PM_SPEED_PALETTE + pt) * 16, nib, DV_SPEED_FAST);
        if (pt == 0) f8[si][0] |= ((uint32_t)nib << 3) & 0xff;
        if (pt == 1) f8[si][0] |= (uint32_t)nib;
        if (pt == 2) f8[si][1] |= ((uint32_t)nib << 3) & 0xff;
        if (pt == 3) f8[si][1] |= (uint32_t)nib;
    }
    if (!ENC && s.cmd.in.underflow) { s.status = ST_NEED_INPUT; return; }
    bool do_cm = ENC ? s.desired_do_context_map : true;
    code_context_map<ENC>(s, g, 0, ENC ? in->lit_map : nullptr, (ENC && do_cm) ? in->lit_map_len : 0, s.lcm, 16384);
    if (s.status != ST_OK) return;
    s.cmap_lru = g.l16;
    code_context_map<ENC>(s, g, 1, ENC ? in->dist_map : nullptr, (ENC && do_cm) ? in->dist_map_len : 0, s.dcm, 1024);
    if (s.status != ST_OK) return;
    __syncwarp(g.mask);
    for (uint32_t index = 0; index < 8192; index++) {
        int mv = 0;
        if (ENC) mv = !s.desired_do_context_map ? 4 : (!combine ? 0 : (int)in->mixing[index]);
        uint32_t prior = index >= 256 ? (uint32_t)(s.mix[index - 256] & 0xf) : 16u;
        mv = cmd_nibble<ENC>(s, g, pp + (PM_MIXING_VALUE + prior) * 16, mv, DV_SPEED_PLANE);
        if (g.store0) s.mix[index] = (uint8_t)mv;
        if ((index & 255) == 255) { __syncwarp(g.mask); if (!ENC && s.cmd.in.underflow) { s.status = ST_NEED_INPUT; return; } }
    }
    __syncwarp(g.mask);
    // obs_prediction_mode_context_map
    s.mixing_param = mixing_math;
    s.mixing_trait = mixing_math > 1;
    if (pm > 3) { s.status = ST_FAIL; return; }                     // PredictionModeOutOfBounds
    s.pred_mode = (uint32_t)pm;
    for (int k = 0; k < 4; k++) {
        uint32_t a = speed_to_u8_u16(u8_to_speed_u16(f8[k][0])), b = speed_to_u8_u16(u8_to_speed_u16(f8[k][1]));
        s.adapt[k].inc = u8_to_speed(a); s.adapt[k].lim = u8_to_speed(b);
    }
    s.lit_slabs_ready = false;
}

// fresh per-stream state (CrossCommandBookKeeping::new codec/interface.rs:348-402, LiteralBookKeeping::new :246-264)
static __device__ __noinline__ void stream_reset(Stream &s, const Grp g) {
    s.distance_lru[0] = 4; s.distance_lru[1] = 11; s.distance_lru[2] = 15; s.distance_lru[3] = 16;
    for (int i = 0; i < 3; i++) { s.btype_lru[i][0] = 0; s.btype_lru[i][1] = 1; s.btype_max_seen[i] = 0; }
    s.last_dlen = 1; s.last_clen = 1; s.last_llen = 1; s.last_4_states = 3 << 4;
    s.cmap_lru = 0;
    s.last_8 = 0; s.btype_last = 0;
    s.pred_mode = 0;   // LiteralPredictionModeNibble::default() (brotli crate, NOT-IN-TREE): LSB6 assumed, unobservable when a
                       // PredictionMode command precedes the first literal -- "parity unpinned" for PM-less streams
    for (int k = 0; k < 4; k++) { s.adapt[k].inc = 0x10; s.adapt[k].lim = 0x2000; }
    for (int k = 0; k < 2; k++) { s.mw[k].w0 = 1; s.mw[k].w1 = 1; s.mw[k].norm = 1 << 14; }
    s.mixing_param = 1; s.mixing_trait = false; s.lit_slabs_ready = false;
    s.status = ST_OK;
    // zero maps (fresh allocations are zeroed: ffi/alloc_util.rs:70-99), clear slab bitmaps, default the small dense priors
    __syncwarp(g.mask);
    if (g.writer) {
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4 *p = reinterpret_cast<uint4 *>(s.lcm);
        for (uint32_t i = g.l16; i < (16384 + 8192 + 1024) / 16; i += 16) p[i] = z;   // lcm, mix, dcm are contiguous
        for (uint32_t i = g.l16; i < 65; i += 16) s.bitmaps[i] = 0;
    }
    store_default_slab(g, s.misc, (uint32_t)MISC_CDFS);
    __syncwarp(g.mask);
}

}  // namespace dv
