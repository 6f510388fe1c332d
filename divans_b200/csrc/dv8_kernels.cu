// dv8_kernels.cu -- stream decoder of the 8-lane engine (dv8_core.cuh): persistent warps, FOUR streams per warp in lock step,
// work pulled from a global counter.  Same framing pre-pass (dv_kernels.cu) and the same command state machine
// (dv_engine_kernel.cuh, transition<false, true>) as the 16/32-lane decoders.
#include "dv8_core.cuh"

namespace dv {

constexpr int DEC8_BLOCK_THREADS = 32;          // one warp = 4 streams per block: blocks spread evenly over the SMs
constexpr int DEC8_MIN_BLOCKS = 16;             // 16 warps = 64 streams per SM (9472 per B200), <= 128 registers

// a slot's generation wrapped (every 255 streams): forget every tag
static __device__ __noinline__ void clear_tags(const G2 g, uint8_t *slot) {
    uint4 *p = reinterpret_cast<uint4 *>(slot + OFF_TAGS_HI);
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint32_t i = (uint32_t)g.l16; i < (uint32_t)(2 * LIT_TABLE_CDFS / 16); i += (uint32_t)g.nl) p[i] = z;
    __syncwarp(g.gmask);
}

__global__ void __launch_bounds__(DEC8_BLOCK_THREADS, DEC8_MIN_BLOCKS) decode_kernel8(DecodeParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int group_in_block = warp_in_block * 4 + (lane >> 3);
    const uint32_t slot = blockIdx.x * (DEC8_BLOCK_THREADS / 8) + group_in_block;
    G2 g;
    g.l16 = lane & 7;
    g.shift = lane & 24;
    g.gmask = 0xffu << (lane & 24);
    g.store0 = (lane & 7) == 0;
    g.nl = 8;

    St s;
    s.slot = p.arena + (uint64_t)slot * SLOT_STRIDE;
    s.c = reinterpret_cast<Cold *>(smem + group_in_block * SMEM_BYTES_PER_GROUP8);
    s.tables = p.tables;
    s.state = S_IDLE;
    s.c->desired_context_mixing = 0; s.c->desired_prior_depth = 0; s.c->desired_force_stride = 9; s.c->desired_do_context_map = true;
    s.c->have_desired_adapt = false; s.c->desired_adapt0 = s.c->desired_adapt1 = s.c->desired_adapt2 = s.c->desired_adapt3 = 0;
    s.c->in.cmds = nullptr; s.c->in.n_cmds = 0; s.c->in.pos = 0; s.c->in.n_pms = 0; s.c->in.pms = nullptr; s.c->in.lits = nullptr;
    s.c->model_rev = p.model_rev;
    s.c->sidx = 0; s.out = nullptr; s.out_pos = 0; s.c->out_cap = 0; s.c->ring_len = 1024;
    s.c->gen_ctr = *reinterpret_cast<const uint32_t *>(s.slot + OFF_HDR);   // generations survive from launch to launch
    s.gen = 0;
    st_reset(s);
    coder_init_dec(s.cur, nullptr, 0); s.cur.need_a = 0; coder_init_dec(s.c->oth, nullptr, 0);
    Next nx; nx.cdf = A_misc(s, MI_DUMMY); nx.cdf2 = nullptr; nx.speed = SPK_NONE; nx.sym = 0; nx.mix_hi = false; nx.tag = nullptr;
    store_default_cdfs(g, reinterpret_cast<int16_t *>(s.slot + OFF_MISC), (uint32_t)MISC_CDFS);   // incl. the dummy CDF
    bool exhausted = false;

    for (;;) {
        __syncwarp();
        // ---- fetch work for idle groups (converged; the broadcast shuffle is executed by every lane) ----
        const bool want = (s.state == S_IDLE) && !exhausted;
        if (__any_sync(FULL, want)) {
            uint32_t v = 0;
            if (want && g.store0) v = atomicAdd(p.work_counter, 1u);
            v = __shfl_sync(FULL, v, 0, 8);
            if (want) {
                if (v >= p.n_streams) exhausted = true;
                else if (p.status[v] != ST_OK) { if (g.store0) p.out_len[v] = 0; }   // framing / CRC failure: stay idle, fetch again
                else {
                    const uint8_t *in = p.in + p.in_off[v];
                    const uint32_t pay0 = p.frame[4 * v + 1], pay1 = p.frame[4 * v + 2];
                    const uint8_t *pl = p.payload + 16ull * p.frame[4 * v + 3];
                    s.c->sidx = v;
                    s.out = p.out + p.out_off[v];
                    uint64_t cap = p.out_cap[v];
                    s.c->out_cap = cap > 0xffffffffull ? 0xffffffffu : (uint32_t)cap; s.out_pos = 0;
                    s.c->ring_len = 1u << in[5];
                    reset_slot(g, s.slot, s.c->bitmaps);
                    st_reset(s);
                    // a new generation: every literal prior of the slot reads as the default CDF until this stream writes it
                    uint32_t ctr = s.c->gen_ctr + 1;
                    if ((ctr & 0xffu) == 0) { clear_tags(g, s.slot); ctr++; }
                    s.c->gen_ctr = ctr; s.gen = ctr & 0xffu;
                    coder_init_dec(s.cur, reinterpret_cast<const uint32_t *>(pl), pay0 >> 2);   // command stream (CMD_CODER, codec/interface.rs:49)
                    coder_init_dec(s.c->oth, reinterpret_cast<const uint32_t *>(pl + (((uint64_t)pay0 + 15) & ~15ull)), pay1 >> 2);   // literal stream (LIT_CODER, :50)
                    enter_cmd_type<false>(s, nx);
                }
            }
            if (__all_sync(FULL, exhausted && s.state == S_IDLE)) break;
            __syncwarp();
        }
        // ---- whole literal bytes while every group is at a byte boundary of a literal (or out of work) ----
        const bool lit = s.state == S_LIT_HI;
        if (__all_sync(FULL, lit || (exhausted && s.state == S_IDLE)) && __any_sync(FULL, lit)) {
            literal_fast8(s, nx, g, lit);
            if (lit) {
                if (s.cur.underflow) s.status = ST_NEED_INPUT;
                if (s.lit_left == 0 && s.status == ST_OK) { swap_coders(s); enter_cmd_type<false>(s, nx); }
                if (s.status != ST_OK) {
                    if (g.store0) { p.out_len[s.c->sidx] = s.out_pos; p.status[s.c->sidx] = s.status; }
                    s.state = S_IDLE; s.status = ST_OK;
                    nx.cdf = A_misc(s, MI_DUMMY); nx.cdf2 = nullptr; nx.speed = SPK_NONE; nx.tag = nullptr;
                    coder_init_dec(s.cur, nullptr, 0); s.cur.need_a = 0;
                }
            }
            continue;
        }
        // ---- one nibble per group ----
        const bool busy = s.state != S_IDLE;
        const int sym = nibble_core8(s, nx, g);
        // ---- per-group scalar state machines (divergent) ----
        if (busy) {
            if (s.cur.underflow) s.status = ST_NEED_INPUT;
            else transition<false, true>(s, nx, g, sym);
            if (s.status != ST_OK || s.state == S_IDLE) {
                if (s.status == ST_OK && s.c->oth.underflow) s.status = ST_NEED_INPUT;
                if (g.store0) { p.out_len[s.c->sidx] = s.out_pos; p.status[s.c->sidx] = s.status; }
                s.state = S_IDLE; s.status = ST_OK;
                nx.cdf = A_misc(s, MI_DUMMY); nx.cdf2 = nullptr; nx.speed = SPK_NONE; nx.tag = nullptr;
                coder_init_dec(s.cur, nullptr, 0); s.cur.need_a = 0;
            }
        }
    }
    if (g.store0) *reinterpret_cast<uint32_t *>(s.slot + OFF_HDR) = s.c->gen_ctr;
}

void launch_decode8(const DecodeParams &p, uint32_t n_blocks, cudaStream_t st) {
    size_t smem = (size_t)(DEC8_BLOCK_THREADS / 8) * SMEM_BYTES_PER_GROUP8;
    decode_kernel8<<<n_blocks, DEC8_BLOCK_THREADS, smem, st>>>(p);
}
int decode_max_blocks_per_sm8() {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, decode_kernel8, DEC8_BLOCK_THREADS, (size_t)(DEC8_BLOCK_THREADS / 8) * SMEM_BYTES_PER_GROUP8);
    return nb < DEC8_MIN_BLOCKS ? nb : DEC8_MIN_BLOCKS;
}
int decode_groups_per_block8() { return DEC8_BLOCK_THREADS / 8; }

}  // namespace dv
