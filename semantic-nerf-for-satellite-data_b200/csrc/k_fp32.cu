// fp32 verification mode of K2 - see k_fp32.cuh.  Plain CUDA-core kernels: a 128 x 128 x 16 register-tiled SGEMM whose A
// operand is the concatenation of up to two column blocks (the reference's torch.cat inputs: skip connection,
// cat(f, sun_d), cat(f, t); per-ray blocks are broadcast by row index instead of repeat_interleave), with bias,
// the SIREN / head activation and a strided store (head outputs land directly in the packed (P, 9+C) tensor).
#include "k_fp32.cuh"

namespace snb {

namespace {

constexpr int FB = 128;   // tile rows / columns
constexpr int FK = 16;    // k-step
constexpr int FPAD = 4;

__device__ __forceinline__ float f32_act(float y, int act, float w0) {
  switch (act) {
    case F32_SIN: return sinf(w0 * y);
    case F32_SIGMOID: return 1.0f / (1.0f + expf(-y));
    case F32_SOFTPLUS: return y > 20.0f ? y : log1pf(expf(y));
    case F32_RGB: return (1.0f / (1.0f + expf(-y))) * 1.002f - 0.001f;
    case F32_RELU: return fmaxf(y, 0.f);
    default: return y;
  }
}

__global__ void __launch_bounds__(256) f32_gemm_kernel(const __grid_constant__ F32Gemm g) {
  __shared__ __align__(16) float As[FK][FB + FPAD];
  __shared__ __align__(16) float Bs[FK][FB + FPAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.y * FB;
  const int n0 = blockIdx.x * FB;
  const int lrow = tid >> 1, lk = (tid & 1) * 8;   // this thread's share of a tile load: one row, 8 consecutive k

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  int koff = 0;
  for (int s = 0; s < g.nseg; ++s) {
    const F32Seg sg = g.seg[s];
    const long long m = m0 + lrow;
    const float* arow = nullptr;
    if (m < g.M) arow = sg.a + (sg.row_div > 1 ? (sg.row_off + m) / sg.row_div : m) * sg.lda;
    const int n = n0 + lrow;
    const float* wrow = n < g.N ? g.w + (long long)n * g.ldw + koff : nullptr;
    for (int k0 = 0; k0 < sg.k; k0 += FK) {
      float av[8], bv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + lk + j;
        av[j] = (arow != nullptr && k < sg.k) ? __ldg(arow + k) : 0.f;
        bv[j] = (wrow != nullptr && k < sg.k) ? __ldg(wrow + k) : 0.f;
      }
      __syncthreads();   // the previous k-step has been consumed
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        As[lk + j][lrow] = av[j];
        Bs[lk + j][lrow] = bv[j];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < FK; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    koff += sg.k;
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + ty * 8 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tx * 8 + j;
      if (n >= g.N) continue;
      const float y = acc[i][j] + (g.bias ? __ldg(g.bias + n) : 0.f);
      g.c[m * g.ldc + n] = f32_act(y, g.act, g.w0);
    }
  }
}

__global__ void __launch_bounds__(256) f32_posenc_kernel(const float* __restrict__ xyz, long long M, int n_freq,
                                                         float* __restrict__ enc, int ld_enc) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float x[3] = {xyz[m * 3], xyz[m * 3 + 1], xyz[m * 3 + 2]};
  float* e = enc + m * ld_enc;
  if (n_freq == 0) {
    e[0] = x[0]; e[1] = x[1]; e[2] = x[2];
    return;
  }
  float f = 1.0f;
  for (int k = 0; k < n_freq; ++k, f *= 2.0f) {   // commons.py:68-74: for each frequency sin(3) then cos(3)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = f * x[c];   // exact (power of two)
      e[k * 6 + c] = sinf(v);
      e[k * 6 + 3 + c] = cosf(v);
    }
  }
}

__global__ void __launch_bounds__(256) f32_sky_kernel(const float* __restrict__ sky, long long M, int row_div, long long row_off,
                                                      float* __restrict__ out, int n_out) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long r = row_div > 1 ? (row_off + m) / row_div : row_off + m;
  out[m * n_out + 5] = sky[r * 3];
  out[m * n_out + 6] = sky[r * 3 + 1];
  out[m * n_out + 7] = sky[r * 3 + 2];
}

}  // namespace

int f32_gemm_launch(const F32Gemm& g, cudaStream_t st) {
  SNB_CHECK_ARG(g.M > 0 && g.N > 0 && g.nseg >= 1 && g.nseg <= 2 && g.w && g.c, SNB_ERR_INVALID, "fp32 gemm: bad arguments");
  dim3 grid((g.N + FB - 1) / FB, (g.M + FB - 1) / FB);
  f32_gemm_kernel<<<grid, 256, 0, st>>>(g);
  return launch_status("f32_gemm_kernel");
}

int f32_posenc_launch(const float* xyz, long long M, int n_freq, float* enc, int ld_enc, cudaStream_t st) {
  f32_posenc_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(xyz, M, n_freq, enc, ld_enc);
  return launch_status("f32_posenc_kernel");
}

int f32_sky_launch(const float* sky, long long M, int row_div, long long row_off, float* out, int n_out, cudaStream_t st) {
  f32_sky_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(sky, M, row_div, row_off, out, n_out);
  return launch_status("f32_sky_kernel");
}

}  // namespace snb
