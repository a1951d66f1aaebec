// K2 chained MLP kernel: ONE persistent launch runs a whole chain of layers (the forward trunk + heads,
// or the backward dgrad chain) instead of one launch per layer.
//
// Every SM pair (cta_group::2) owns 256-row blocks of the sample matrix and carries CHAIN_SLOTS of them,
// interleaved, through all layers of the chain: tile order (l, slot0, n0) (l, slot0, n1) (l, slot1, n0) ...
// then layer l+1.  A layer's input rows were written by the SAME CTA one layer-step earlier, so they are
// read back from L2 instead of HBM; inference keeps the inter-layer activations in a small per-pair scratch
// that never leaves L2 at all.  There is no per-layer launch / pipeline fill / drain, and the epilogue of
// one slot overlaps the MMAs of the other.  A per-slot mbarrier (arrived by the epilogue group leaders once
// the TMA stores of a layer have completed, waited by the A-producer) orders producer and consumer.
#pragma once
#include "k2_gemm.cuh"

namespace snb {

constexpr int CHAIN_SLOTS = 3;         // (a pair's 4 blocks as ONE group of 4 instead of 2 + 2 was measured: +1.5 % step time at
                                       // 1024 rays, profiles/r02j_tune.log)
constexpr int CHAIN_MAX_LAYERS = 28;   // both passes of a launch together (main forward 14 + solar forward 13); ~1 KB of
                                       // kernel parameters per layer, 32 764 bytes at most
constexpr int CHAIN_A_STAGES = 5;      // 16 KB slots (128 rows x 64 k)
constexpr int CHAIN_B_STAGES = 4;      // 16 KB slots (this CTA's 128 of the 256 weight rows x 64 k; weights are L2-resident)
constexpr int CHAIN_EPI_WARPS = 16;    // two groups of 8 warps; warps w and w+4 of a group split a 64-column chunk
constexpr int CHAIN_THREADS = 96 + CHAIN_EPI_WARPS * 32 + 32;   // + the store warp
constexpr int CHAIN_RING_BYTES = (CHAIN_A_STAGES + CHAIN_B_STAGES) * 16384;
constexpr int CHAIN_SMEM_BYTES = CHAIN_RING_BYTES + GEMM_NUM_STAGING * GEMM_STAGING +
                                 8192 /*per-warp bias slices*/ + 8192 /*sign-mask tiles*/ + 1024 /*barriers*/ + 1024 /*alignment*/;
static_assert(CHAIN_SMEM_BYTES <= 227 * 1024, "chained kernel: shared memory");

struct alignas(64) ChainMaps {   // only touched by the TMA unit
  CUtensorMap tmA[3];   // A K-segments, box {64 k, 128 rows}
  CUtensorMap tmB;      // weights [N, K], box {64 k, 128 rows}
  CUtensorMap tmO0;     // output, box {64 columns, 128 rows}
  CUtensorMap tmMul;    // EPI_MUL multiplicand, box {64 columns, 128 rows}
  CUtensorMap tmMask;   // sign mask (u32 words, not swizzled), box {8 words = the 256 columns of a tile, 128 rows}
};

struct ChainLayer {     // scalars every role reads once per tile: kept together so they stay in the constant cache
  int seg_kb[3];
  int a_scratch[3];     // segment lives in the per-pair scratch (row = (pair*SLOTS + slot)*256) instead of at the block's rows
  int nseg;
  int kb_total;
  int tail_k16;         // k = 16 MMA steps of the LAST k-block (4 = a full 64-column block): the 16-column aux / dPre16 K-segments
                        // are zero beyond their real columns, so their other three MMAs would only add zeros
  int n_tiles;          // N / 256
  int epi;              // EPI_SIN / EPI_LINEAR / EPI_MUL (256-column tiles) or EPI_HEADOUT (one 16-column tile, see `rows`)
  int relu;             // EPI_SIN layers: max(acc + bias, 0) instead of sin (vanilla NeRF); no sign mask
  int mul_siren;        // EPI_MUL: 2 = ReLU derivative [h > 0] of the saved activation; 1 = the multiplicand is the SIREN derivative rebuilt from the saved activation h = sin(y)
                        // (tmMul) and the sign mask: w0 * (-1)^bit * sqrt(1 - h^2); 0: multiply by the tmMul tensor itself
  int o_scratch;        // outputs go to the per-pair scratch
  int mask_ld;          // 32-bit words per row of `mask`
  float w0;
  int rows_mode;        // EPI_HEADOUT: the 16 head pre-activations are the sum of up to three of these layers, each placed
                        // right after the layer that produced its input (still in L2): 0 = store the partial sums to
                        // `part`, 1 = add to `part`, 2 = add `part` (if set) + bias, apply the head activations, write the
                        // packed (P, n_out) fp32 rows
  float* part;          // EPI_HEADOUT: (M, 16) fp32 partial sums
  uint32_t* mask;       // EPI_SIN: written (NULL = inference), one bit per element = [cos(w0*(acc+bias)) < 0]: with h it is
                        // all the backward pass needs of the derivative (1 bit instead of 16 per element of HBM write
                        // traffic); EPI_MUL + mul_siren: read.  Word layout: see chunk_math in k2_chain.cu
  const float* bias;
};

// A launch carries one or two PASSES: row ranges with their own layer sequence (a training step's main pass and its
// solar-correction pass).  Every SM pair first works through its blocks of pass 0, then through its blocks of pass 1; pass 1
// deals its blocks to the pairs rotated by `shift`, so the pairs that carry one block more than the others of pass 0 carry
// one less of pass 1: at 1024 rays x 64 samples (256 blocks per pass on 74 pairs) a pair runs 4 + 3 blocks instead of the
// 4 + 4 of two launches.
struct ChainPass {
  int layer0, n_layers; // layers [layer0, layer0 + n_layers) of ChainArgs::layers
  int M;                // rows
  int n_blocks;         // ceil(M / 256)
  int shift;            // block b belongs to pair (b + shift) % n_pairs
  int head_mask;
  // head output (EPI_HEADOUT, rows_mode 2): packed (M, n_out) fp32 [rgb 0:3 | sigma 3 | sun 4 | sky 5:8 | beta 8 | sem 9:]
  float* out_packed;
  const float* sky;     // (rays or points, 3) per-ray sky colour from K1
};

struct ChainArgs {
  ChainLayer layers[CHAIN_MAX_LAYERS];
  ChainPass pass[2];
  int n_passes;
  int n_layers;         // all passes together
  int n_out, rows_per_ray, n_classes, sem_sigmoid;
  int nerf;             // head output: the sun column is written as 1 (no lighting model: irradiance = 1 in K3)
  int beta_s;           // head output: pre-activation row 6 is the separate semantic uncertainty head (packed column 9, softplus);
                        // the class rows / columns follow it
  // the packed weight image (or NULL): every thread prefetches one or two of its 128-byte lines into L2 before the first
  // tile, so the first block of a pair does not pay an HBM round trip per layer for weights the activation traffic of the
  // previous pass has evicted (at 1024 rays a pair carries only 3-4 blocks per pass)
  const void* prefetch;
  unsigned prefetch_bytes;
  int exp;              // SNB_EXPERIMENTS builds only (env SNB_EXP): bit 0 skip every second B load, bit 1 every second A load,
                        // bit 2 skip the epilogue math, bit 3 skip the TMA stores, bit 4 skip the multiplicand loads
  ChainMaps maps[CHAIN_MAX_LAYERS];
};

// rows of per-pair scratch a chain launch may address: (SMs/2) * CHAIN_SLOTS * 256
int chain_scratch_rows();
int chain_launch(const ChainArgs& a, cudaStream_t st);

}  // namespace snb
