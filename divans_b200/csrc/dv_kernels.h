// dv_kernels.h -- host-visible launch wrappers of the divANS kernels.
#pragma once
#include "dv_common.cuh"

namespace dv {
constexpr int DECODE_BLOCK_THREADS = 64;

// Resident blocks per SM of a stream kernel (register-limited).  (L1 and shared memory share 256 KB per SM and the driver gives
// these kernels a 100 KB carve-out although they use ~2.3 KB per block.  Asking for less through
// cudaFuncAttributePreferredSharedMemoryCarveout brought nothing: 16 % left the configuration ncu reports at 100 KB and the
// time unchanged, 28 % was 1-4 % slower -- profiles/r2_v8_sweep.txt -- so no preference is set.)
template <typename K> static inline int stream_kernel_blocks_per_sm(K kernel, int threads, size_t dyn_smem) {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, dyn_smem);
    return nb;
}

void launch_frame(const FrameParams &p, uint8_t *payload, uint64_t payload_cap_bytes, cudaStream_t st);   // frame + payload scan + demux (3 launches)
void launch_decode32(const DecodeParams &p, uint32_t n_blocks, cudaStream_t st);
void launch_decode16(const DecodeParams &p, uint32_t n_blocks, cudaStream_t st);
int decode_max_blocks_per_sm32();
// v2 engine (dv2_kernels.cu), lanes_per_stream = 16 (two streams per warp) or 8 (four)
void launch_decode_v2(int lanes_per_stream, bool prefetch, const DecodeParams &p, uint32_t n_blocks, cudaStream_t st);
int decode_max_blocks_per_sm_v2(int lanes_per_stream);
int decode_groups_per_block_v2(int lanes_per_stream);
int decode_max_blocks_per_sm16();
void launch_encode_model(const EncodeParams &p, uint32_t n_blocks, cudaStream_t st);   // groups of 16 lanes
void launch_encode_flush_mux(const EncodeParams &p, cudaStream_t st);                  // reverse rANS + mux/CRC (2 launches)
int encode_max_blocks_per_sm();
// the reference's feature="blend" probability model (dv_blend.cuh): 16 lanes per stream, generic nibble path only
void launch_decode16_blend(const DecodeParams &p, uint32_t n_blocks, cudaStream_t st);
int decode_max_blocks_per_sm16_blend();
void launch_encode_model_blend(const EncodeParams &p, uint32_t n_blocks, cudaStream_t st);
void launch_rcp15_init(uint64_t *tab, cudaStream_t st);
}  // namespace dv
