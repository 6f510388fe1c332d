// dv_kernels.cu -- sm_100a kernels of the divANS batch engine: framing/CRC pre-pass and the stream decoder.
#include "dv_codec.cuh"
#include "dv_kernels.h"

namespace dv {

// ---------------------------------------------------------------------------------------------------------------
// frame kernel: one thread per stream walks the 16-byte header and the mux record chain (mux.rs:384-444) to the
// EOF marker, checks the trailer magic and the CRC32C of header..EOF marker (codec/decoder.rs:186-213).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t crc_step(const uint32_t *tab, uint32_t crc, uint32_t byte) {
    return tab[(crc ^ byte) & 0xff] ^ (crc >> 8);
}
#if DV_LPS == 32
__global__ void __launch_bounds__(128) frame_kernel(FrameParams p) {
    __shared__ uint32_t tab[4][256];   // slice-by-4 tables for the Castagnoli polynomial (reflected 0x82F63B78)
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
        tab[0][i] = c;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = tab[0][i];
        for (int t = 1; t < 4; t++) { c = tab[0][c & 0xff] ^ (c >> 8); tab[t][i] = c; }
    }
    __syncthreads();
    uint32_t sidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx >= p.n_streams) return;
    const uint8_t *in = p.in + p.in_off[sidx];
    uint64_t n = p.in_len[sidx];
    int32_t st = ST_OK;
    uint32_t body_end = 0;
    if (n < 16) st = ST_NEED_INPUT;
    else if (in[0] != 0xff || in[1] != 0xe5 || in[2] != 0x8c || in[3] != 0x9f) st = ST_FAIL;   // MAGIC_NUMBER, src/interface.rs:164
    else if (in[5] < 10 || in[5] >= 25) st = ST_FAIL;                                             // BadWindowSize, divans_decompressor.rs:47-50
    else {
        uint64_t pos = 16;
        for (;;) {
            if (pos >= n) { st = ST_NEED_INPUT; break; }
            uint32_t b = in[pos];
            if (b == 0xff) {
                if (pos + 3 > n) { st = ST_NEED_INPUT; break; }
                if (in[pos + 1] != 0xfe || in[pos + 2] != 0xff) { st = ST_FAIL; break; }
                body_end = (uint32_t)pos;
                break;
            }
            uint64_t len, hdr;
            if (b < 16) { if (pos + 3 > n) { st = ST_NEED_INPUT; break; } len = ((uint64_t)in[pos + 1] | ((uint64_t)in[pos + 2] << 8)) + 1; hdr = 3; }
            else { uint32_t k = b >> 4; if (k > 3) { st = ST_FAIL; break; } len = 1024ull << (k << 1); hdr = 1; }
            if (pos + hdr + len > n) { st = ST_NEED_INPUT; break; }
            pos += hdr + len;
        }
        if (st == ST_OK) {
            uint64_t tr = (uint64_t)body_end + 3;
            if (tr + 8 > n) st = ST_NEED_INPUT;
            else {
                if (in[tr + 4] != 'a' || in[tr + 5] != 'n' || in[tr + 6] != 's' || in[tr + 7] != '~') st = ST_FAIL;
                if (st == ST_OK && !(p.flags & 3u)) {
                    uint32_t crc = 0xffffffffu;
                    uint64_t i = 0;
                    for (; i < tr && (((uintptr_t)(in + i)) & 3); i++) crc = crc_step(tab[0], crc, in[i]);
                    for (; i + 4 <= tr; i += 4) {
                        uint32_t w = *reinterpret_cast<const uint32_t *>(in + i) ^ crc;
                        crc = tab[3][w & 0xff] ^ tab[2][(w >> 8) & 0xff] ^ tab[1][(w >> 16) & 0xff] ^ tab[0][w >> 24];
                    }
                    for (; i < tr; i++) crc = crc_step(tab[0], crc, in[i]);
                    crc = ~crc;
                    uint32_t want = (uint32_t)in[tr] | ((uint32_t)in[tr + 1] << 8) | ((uint32_t)in[tr + 2] << 16) | ((uint32_t)in[tr + 3] << 24);
                    if (crc != want) st = ST_FAIL;   // BadChecksum
                }
            }
        }
    }
    p.body_end[sidx] = st == ST_OK ? body_end : 0;
    p.status[sidx] = st;
}

#endif  // DV_LPS == 32 (frame kernel)

// ---------------------------------------------------------------------------------------------------------------
// decode kernel: persistent lane-groups pull stream indices from a global counter.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SMEM_WORDS_PER_GROUP = 96;   // 65 bitmap words (padded to 80) + 64 B dictionary scratch

template <int LPS>   // lanes per stream: 32 (upper half mirrors) or 16 (two streams per warp)
__global__ void __launch_bounds__(DECODE_BLOCK_THREADS) decode_kernel(DecodeParams p) {
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    constexpr int GPW = 32 / LPS;
    const int group_in_warp = (LPS == 16) ? (lane >> 4) : 0;
    const int group_in_block = warp_in_block * GPW + group_in_warp;
    const uint32_t slot = blockIdx.x * (DECODE_BLOCK_THREADS / LPS) + group_in_block;
    Grp g;
    g.l16 = lane & 15;
    g.shift = (LPS == 16) ? (lane & 16) : 0;
    g.mask = (LPS == 16) ? (0xffffu << (lane & 16)) : 0xffffffffu;
    g.writer = (LPS == 16) ? true : (lane < 16);
    g.lane0 = (lane & 15) == 0;
    g.store0 = (LPS == 16) ? ((lane & 15) == 0) : (lane == 0);

    Stream s;
    uint8_t *slotp = p.arena + (uint64_t)slot * SLOT_STRIDE;
    s.lit_hi = reinterpret_cast<int16_t *>(slotp + OFF_LIT_HI);
    s.lit_lo = reinterpret_cast<int16_t *>(slotp + OFF_LIT_LO);
    s.lit_cm = reinterpret_cast<int16_t *>(slotp + OFF_LIT_CM);
    s.ctype_slabs = reinterpret_cast<int16_t *>(slotp + OFF_CTYPE);
    s.dprior_slabs = reinterpret_cast<int16_t *>(slotp + OFF_DPRIOR);
    s.misc = reinterpret_cast<int16_t *>(slotp + OFF_MISC);
    s.lcm = slotp + OFF_LCM; s.mix = slotp + OFF_MIX; s.dcm = slotp + OFF_DCM;
    s.bitmaps = smem + group_in_block * SMEM_WORDS_PER_GROUP;
    s.scratch = reinterpret_cast<uint8_t *>(s.bitmaps + 80);
    s.tables = p.tables;
    s.desired_context_mixing = 0; s.desired_prior_depth = 0; s.desired_force_stride = 9; s.desired_do_context_map = true;
    s.have_desired_adapt = false;

    uint64_t tot_cmd = 0, tot_lit = 0;
    for (;;) {
        uint32_t sidx = 0;
        if (g.store0) sidx = atomicAdd(p.work_counter, 1u);
        sidx = __shfl_sync(g.mask, sidx, 0, LPS);
        if (sidx >= p.n_streams) break;
        if (p.status[sidx] != ST_OK) { if (g.store0) p.out_len[sidx] = 0; continue; }   // framing / CRC failure
        const uint8_t *in = p.in + p.in_off[sidx];
        const uint32_t body_end = p.body_end[sidx];
        s.out = p.out + p.out_off[sidx]; s.out_cap = p.out_cap[sidx]; s.out_pos = 0;
        s.ring_len = 1u << in[5];
        stream_reset(s, g);
        coder_init_dec(s.cmd, in + 16, in + body_end, 0);
        coder_init_dec(s.lit, in + 16, in + body_end, 1);
        // DivansCodec::encode_or_decode_one_command (codec/mod.rs:652-1024) fused with
        // DivansDecoderCodec::decode_process_output (codec/decoder.rs:230-419)
        for (;;) {
            int t = cmd_nibble<false>(s, g, s.misc + (MI_CC + (s.last_4_states >> 4)) * 16, 0, DV_SPEED_ROCKET);
            if (s.cmd.in.underflow) { s.status = ST_NEED_INPUT; break; }
            if (t == 0xf) break;
            if (t == 1) {   // copy
                s.last_4_states = (s.last_4_states >> 2) | 64;
                uint32_t d = 0, nb = 0;
                code_copy<false>(s, g, d, nb);
                if (s.status != ST_OK) break;
                obs_distance(s, d);
                replay_copy(s, g, d, nb);
            } else if (t == 2) {   // dict
                s.last_4_states = (s.last_4_states >> 2) | 192;
                uint32_t id = 0, sz = 0, tr = 0;
                code_dict<false>(s, g, id, sz, tr);
                if (s.status != ST_OK) break;
                replay_dict(s, g, sz, id, tr);
            } else if (t == 3) {   // literal
                s.last_4_states = (s.last_4_states >> 2) | 128;
                uint32_t len = 0, he = 0;
                code_literal_len<false>(s, g, len, he);
                if (s.cmd.in.underflow) { s.status = ST_NEED_INPUT; break; }
                if ((uint64_t)len > s.out_cap - s.out_pos) { s.status = ST_NEED_OUTPUT; break; }
                ensure_literal_slabs(s, g);
                if (s.mixing_trait) code_literal_bytes<false, true>(s, g, nullptr, len);
                else code_literal_bytes<false, false>(s, g, nullptr, len);
            } else if (t == 4) {   // literal block switch
                uint32_t bt = code_btype<false>(s, g, 0, 0);
                int stride = cmd_nibble<false>(s, g, s.misc + (MI_BTYPE + BT_STRIDE) * 16, 0, DV_SPEED_SLOW);
                (void)stride;
                obs_btype(s, 0, bt);
                s.btype_last = bt;
            } else if (t == 5) {
                uint32_t bt = code_btype<false>(s, g, 1, 0); obs_btype(s, 1, bt);
            } else if (t == 6) {
                uint32_t bt = code_btype<false>(s, g, 2, 0); obs_btype(s, 2, bt);
            } else if (t == 7) {
                code_predmode<false>(s, g, nullptr);
            } else { s.status = ST_FAIL; }   // CommandCodeOutOfBounds
            if (s.status != ST_OK) break;
            if (s.cmd.in.underflow || s.lit.in.underflow) { s.status = ST_NEED_INPUT; break; }
        }
        tot_cmd += s.cmd.n_syms; tot_lit += s.lit.n_syms;
        if (g.store0) { p.out_len[sidx] = s.out_pos; p.status[sidx] = s.status; }
    }
    if (p.nibble_counts && g.store0) {
        atomicAdd((unsigned long long *)&p.nibble_counts[0], (unsigned long long)tot_cmd);
        atomicAdd((unsigned long long *)&p.nibble_counts[1], (unsigned long long)tot_lit);
    }
}

#if DV_LPS == 32
void launch_frame(const FrameParams &p, cudaStream_t st) {
    uint32_t blocks = (p.n_streams + 127) / 128;
    frame_kernel<<<blocks, 128, 0, st>>>(p);
}
// each translation unit has its own copy of the __constant__ LUT: upload to both
cudaError_t upload_ctx_lut32(const uint8_t *host2048) { return cudaMemcpyToSymbol(c_ctx_lut, host2048, 2048); }
#else
cudaError_t upload_ctx_lut16(const uint8_t *host2048) { return cudaMemcpyToSymbol(c_ctx_lut, host2048, 2048); }
#endif
#ifndef DV_LPS
#error "compile with -DDV_LPS=16 or 32 (one translation unit per instantiation keeps ptxas time in check)"
#endif
#if DV_LPS == 32
void launch_decode32(const DecodeParams &p, uint32_t n_blocks, cudaStream_t st) {
    size_t smem = (size_t)(DECODE_BLOCK_THREADS / 32) * SMEM_WORDS_PER_GROUP * 4;
    decode_kernel<32><<<n_blocks, DECODE_BLOCK_THREADS, smem, st>>>(p);
}
int decode_max_blocks_per_sm32() {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, decode_kernel<32>, DECODE_BLOCK_THREADS, (size_t)(DECODE_BLOCK_THREADS / 32) * SMEM_WORDS_PER_GROUP * 4);
    return nb;
}
#else
void launch_decode16(const DecodeParams &p, uint32_t n_blocks, cudaStream_t st) {
    size_t smem = (size_t)(DECODE_BLOCK_THREADS / 16) * SMEM_WORDS_PER_GROUP * 4;
    decode_kernel<16><<<n_blocks, DECODE_BLOCK_THREADS, smem, st>>>(p);
}
int decode_max_blocks_per_sm16() {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, decode_kernel<16>, DECODE_BLOCK_THREADS, (size_t)(DECODE_BLOCK_THREADS / 16) * SMEM_WORDS_PER_GROUP * 4);
    return nb;
}
#endif

}  // namespace dv
