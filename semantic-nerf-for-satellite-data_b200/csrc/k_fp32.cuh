// fp32 verification mode of K2 ("fp32 mode" of the parity contract): the MLP evaluated with fp32 inputs, fp32 weights
// (straight from the flat parameter buffer, nn.Linear layout) and fp32 FMA accumulation on the CUDA cores.
// B200 has no fp32 tensor-core path (kind::tf32 keeps 10 mantissa bits, fewer than this mode promises), so this mode
// trades the tensor pipe for exactness: it is for evaluation / parity runs, not for training throughput.
#pragma once
#include "snb_common.cuh"

namespace snb {

enum F32Act {
  F32_NONE = 0,      // y
  F32_SIN = 1,       // sin(w0 * y)                       (Siren, baseline/models/commons.py:27-38)
  F32_SIGMOID = 2,   // sigmoid(y)
  F32_SOFTPLUS = 3,  // softplus(y), beta = 1, threshold 20
  F32_RGB = 4,       // sigmoid(y) * 1.002 - 0.001        (rs_semantic.py:282-284)
  F32_RELU = 5       // max(y, 0)                         (vanilla NeRF, nerf.py:110)
};

struct F32Seg {      // one block of input columns: the reference's torch.cat([...], -1) operands, in order
  const float* a;    // row-major
  long long lda;
  int k;             // columns
  int row_div;       // 1: one row per point; S: one row per ray (row = (row_off + m) / row_div)
  long long row_off; // global index of the chunk's first point (per-ray segments only)
};

struct F32Gemm {     // C[M, N] = act(w0 * (sum_seg A_seg W[:, koff_seg : koff_seg + k_seg]^T + bias))
  F32Seg seg[2];
  int nseg;
  const float* w;    // [N, ldw] row-major (nn.Linear.weight)
  int ldw;
  const float* bias; // [N]
  int M, N;
  float w0;
  int act;
  float* c;
  long long ldc;
};

int f32_gemm_launch(const F32Gemm& g, cudaStream_t st);
// enc[m, :] = [sin(2^k x), cos(2^k x)]_{k < n_freq} (commons.py:68-74) for n_freq > 0, else a copy of xyz
int f32_posenc_launch(const float* xyz, long long M, int n_freq, float* enc, int ld_enc, cudaStream_t st);
// out[m, 5:8] = sky[(row_off + m) / row_div]
int f32_sky_launch(const float* sky, long long M, int row_div, long long row_off, float* out, int n_out, cudaStream_t st);

}  // namespace snb
