// K2 - the dense contraction: a persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] (+)= sum over K-segments  A_seg[M,K_seg] * B[N,K]^T        (bf16 in, fp32 accumulate in TMEM)
//
// One CTA per SM, 11 warps:  warp 0 = TMA producer for A, warp 1 = tcgen05.mma issuer (+ TMEM owner),
// warp 2 = TMA producer for B, warps 3-10 = two epilogue warpgroups (tcgen05.ld -> bias/activation -> bf16 -> swizzled smem -> TMA
// store) that split the column chunks of every tile.
// Tile 128 x block_n (<=256) x 64, decoupled A/B smem rings, two 256-column TMEM accumulators so the
// epilogue of tile i overlaps the MMAs of tile i+1.
//
// Operand layouts (all 128-byte swizzle, filled by TMA, consumed through UMMA smem descriptors):
//   K-major  : row-major [rows, K]   - forward (activations x weights) and dgrad (dY x W^T copy)
//   MN-major : row-major [K, rows]   - wgrad (dY^T x X: the reduction runs over samples)
// A may be split into up to 3 K-segments read from different tensors (skip connection
// cat(enc,h), cat(f,sun_d,t) heads, the packed head-output layer) - no concat is materialised.
#pragma once
#include "snb_common.cuh"

namespace snb {

enum GemmEpi {
  // bf16-output epilogues: served by the chained kernel (k2_chain.cu), also for single layers
  EPI_SIN = 0,      // h = sin(w0*(acc+bias)) -> bf16 ; optional sign mask of the derivative w0*cos(w0*(acc+bias))
  EPI_LINEAR = 1,   // acc + bias -> bf16
  EPI_MUL = 2,      // acc * mul[m,n] -> bf16, or acc * (SIREN derivative rebuilt from the saved activation + sign mask)
  EPI_HEADOUT = 3,  // N=16 head pre-activations -> packed (P, n_out) fp32 with the reference activations (chained kernel)
  // this kernel:
  EPI_F32ROWS = 4,  // N=16 raw fp32 rows -> f32rows[M,16]
  EPI_WGRAD = 5     // fp32 += acc (split-K): TMA reduce-add (block_n % 32 == 0, >= 32) or red.global
};

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_MAX_BLOCK_N = 256;
constexpr int GEMM_A_STAGE = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;      // 16 KB
constexpr int GEMM_B_STAGE = GEMM_MAX_BLOCK_N * GEMM_BLOCK_K * 2;  // 32 KB
// The A and B operands have independent smem rings (and producer warps): activations come from
// DRAM (~2500 clk under load) and need depth, weights come from L2 and do not.  The ring depths are
// per-launch parameters; together they may use GEMM_OPERAND_BYTES.
constexpr int GEMM_OPERAND_BYTES = 160 * 1024;
constexpr int GEMM_MAX_RING = 8;
constexpr int GEMM_STAGING = 16384;                                // one 128 x 128B swizzled chunk
constexpr int GEMM_NUM_STAGING = 4;
constexpr int GEMM_EPI_THREADS = 256;  // two epilogue warpgroups
constexpr int GEMM_THREADS = 96 + GEMM_EPI_THREADS;  // warp 0: A producer, 1: MMA, 2: B producer, 3-10: epilogue
constexpr int GEMM_ONES_BYTES = 2048;   // constant all-ones B tile (16 x 64 bf16) of the bias-gradient MMA
constexpr int GEMM_SIDE_STAGE = 4096;   // side operand: up to 32 rows x 64 k bf16 per CTA and k-block
constexpr int GEMM_LAYOUT_BYTES = GEMM_OPERAND_BYTES + GEMM_NUM_STAGING * GEMM_STAGING + GEMM_ONES_BYTES + 512;
// the dynamic shared memory is declared 1024-byte aligned; the kernel still rounds its base up and traps if the
// layout would not fit (512 bytes of slack are left below the 227 KB limit)
constexpr int GEMM_SMEM_BYTES = GEMM_LAYOUT_BYTES + 512;

struct GemmArgs {
  CUtensorMap tmA[3];
  CUtensorMap tmB;
  CUtensorMap tmO0;
  int seg_kb[3];
  int nseg;
  int kb_total;
  int M, N;
  int block_n;
  int m_tiles, n_tiles, splits;
  int a_mn, b_mn;
  unsigned a_bytes, b_bytes;
  int a_stages, b_stages;  // ring depths: a_stages*16 KB + b_stages*b_slot <= GEMM_OPERAND_BYTES
  int cta_group;           // 1, or 2 = SM pair per 256 x block_n tile (set before building tmB: its box is block_n/cta_group rows)
  unsigned b_slot;         // bytes per B ring slot
  const float* bias;
  float w0;
  // EPI_WGRAD: if set, colsum[m] += sum over the K (sample) dimension of A[:, m] - the bias gradient rides the weight
  // gradient GEMM as one extra N = 16 MMA per k-block against a constant all-ones tile (n-block 0 tiles only)
  float* colsum;
  // EPI_WGRAD, SM-pair mode: optional SIDE operand riding the same launch - side_out[M, side_n] += A^T x X2 for a narrow
  // second right-hand side X2 (the per-ray `aux` columns of the fused head layer, the encoding block of the skip layer),
  // given TRANSPOSED ([side_n, K] K-major, like a weight matrix) so each CTA of the pair loads side_n / 2 rows x 64 k per
  // k-block.  One extra N = side_n MMA per k-block into spare TMEM columns, shared between the n-tiles like the column
  // sums; replaces a separate launch that re-read the whole dY from HBM.  The side ring lives in the epilogue's staging
  // buffers (idle during the k-loop: a unit runs one tile when colsum / side are set).
  CUtensorMap tmB2;
  int side_n;              // 0, 16 or 64
  float* side_out;
  long long ld_side;
  // EPI_F32ROWS / small-N EPI_WGRAD
  float* f32out;
  long long ldo;
};

// ---- host side -------------------------------------------------------------------------------
// 2-D tensor map over a row-major [outer, inner] array, 128B swizzle, zero OOB fill.
int make_tmap_2d(CUtensorMap* map, const void* ptr, int elem_bytes, uint64_t inner, uint64_t outer,
                 uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);
// sign-mask tiles of the chained kernel: u32 words, [rows, words] row-major, box {8 words, 128 rows}, no swizzle
int make_tmap_mask(CUtensorMap* map, const void* ptr, uint64_t words, uint64_t rows, uint64_t row_stride_bytes);

// fills tile counts / byte counts from M, N, block_n, kb_total, splits, a_mn, b_mn
int gemm_pick_cta_group(int epi, long long M, int N, int block_n);
void gemm_finalize(GemmArgs& a);
int gemm_launch(const GemmArgs& a, int epi, cudaStream_t st);

}  // namespace snb
