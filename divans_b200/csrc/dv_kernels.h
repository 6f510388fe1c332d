// dv_kernels.h -- host-visible launch wrappers of the divANS kernels.
#pragma once
#include "dv_common.cuh"

namespace dv {
constexpr int DECODE_BLOCK_THREADS = 64;

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
void launch_rcp15_init(uint64_t *tab, cudaStream_t st);
}  // namespace dv
