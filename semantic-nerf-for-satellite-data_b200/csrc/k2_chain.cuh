// K2 chained MLP kernel: ONE persistent launch runs a whole chain of layers (the forward trunk + heads,
// or the backward dgrad chain) instead of one launch per layer.
//
// Every SM pair (cta_group::2) owns 256-row blocks of the sample matrix and carries CHAIN_SLOTS of them,
// interleaved, through all layers of the chain: tile order (slot0,l,n0) (slot0,l,n1) (slot1,l,n0) ... then
// layer l+1.  A layer's input rows were written by the same CTA one layer-step earlier, so they are read
// back from L2 instead of HBM (the write-once streams - saved derivatives - carry evict_first hints),
// there is no per-layer launch / pipeline fill / drain, and the epilogue of one slot overlaps the MMAs of
// the other.  A per-slot, per-layer completion counter in shared memory (bumped by the epilogue leaders
// once their TMA stores have completed, polled by the A-producer) orders producer and consumer.
#pragma once
#include "k2_gemm.cuh"

namespace snb {

constexpr int CHAIN_SLOTS = 2;
constexpr int CHAIN_MAX_LAYERS = 16;
constexpr int CHAIN_MAX_COLSUM = 12;   // layers whose bias gradient (column sums) is accumulated in smem
constexpr int CHAIN_A_STAGES = 5;      // 16 KB slots
constexpr int CHAIN_B_STAGES = 3;      // 16 KB slots (weights: L2-resident, evict_last)
constexpr int CHAIN_SMEM_BYTES = (CHAIN_A_STAGES + CHAIN_B_STAGES) * 16384 + GEMM_NUM_STAGING * GEMM_STAGING +
                                 CHAIN_MAX_COLSUM * 512 * 4 + 1024 + 1024;

struct alignas(64) ChainLayer {
  CUtensorMap tmA[2];   // A K-segments, box {64, 128}
  CUtensorMap tmB;      // weights [N, K], box {64, 128} (each CTA of the pair loads half of the 256-wide N tile)
  CUtensorMap tmO0, tmO1, tmMul;
  int seg_kb[2];
  int nseg;
  int kb_total;
  int n_tiles;          // N / 256
  int epi;              // EPI_SIN / EPI_LINEAR / EPI_MUL
  int two_out;
  int cs_slot;          // >= 0: accumulate column sums of the bf16 output into colsum (bias gradient)
  int dep[2];           // chain layers whose output rows (same block) feed this layer's A; -1 = none
  float w0;
  const float* bias;
  float* colsum;
};

struct ChainArgs {
  ChainLayer layers[CHAIN_MAX_LAYERS];
  int n_layers;
  int M;
  int n_blocks;         // ceil(M / 256)
};

int chain_launch(const ChainArgs& a, cudaStream_t st);

}  // namespace snb
