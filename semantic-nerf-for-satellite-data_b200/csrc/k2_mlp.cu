// K2 - the SatNeRF / Semantic-NeRF MLP (8x512 SIREN trunk + heads) as chains of tcgen05 GEMMs.
//
// Follows baseline/models/satnerf.py:208-255 and semantic/models/rs_semantic.py:260-340.
// Algebraic restructuring (results identical up to bf16 rounding):
//   * skip connection cat(enc, h3) (satnerf.py:225-226): two K-segments into one accumulator;
//   * first trunk layer (w0 = 30): bf16x2 split of input and weight, three K-segments
//     hi*Whi + hi*Wlo + lo*Whi, so the phase 30*(W enc + b) keeps ~16 mantissa bits;
//   * rgb / beta / semantic / sun first layers share their input f: ONE GEMM with N = 768..1280;
//     their cat(f, sun_d) / cat(f, t) columns and biases ride in a 16-wide per-row K-segment `aux`;
//   * feats_from_xyz is LINEAR (satnerf.py:163-165 "no non-linearity here") and feeds only those first layers, so it is folded
//     into them: W' = W_h1 W_f, b' = b_h1 + W_h1 b_f are composed in fp32 at pack time and the layer f = W_f h7 + b_f is never
//     evaluated - not in the forward, not in the dgrad (dY7 = dY_hh W'), not as a weight-gradient GEMM: with G1 = dY_hh^T h7
//     (the head layer's wgrad taken against h7 instead of f), dW_h1 = G1 W_f^T + colsum(dY_hh) b_f^T, dW_f = W_h1^T G1,
//     db_f = W_h1^T colsum(dY_hh) are weight-sized fp32 products (< 1 GFLOP per step).  3 x 512 x 512 MACs per sample and
//     pass less (-9.3 % main pass, -10.9 % solar pass), 3 KB per sample less HBM traffic (f, dF);
//   * the 1..C-wide output layers (sigma, rgb.2, sun.6, beta.2, semantic.2) are one N=16 GEMM over
//     the K-segments [h7 | s3 | hh] whose epilogue applies softplus/sigmoid and writes the packed
//     (P, 9+C) fp32 tensor in the reference's column order (rs_semantic.py:291-311).
// Backward = dgrad GEMMs against transposed bf16 weight copies with the saved SIREN derivative
// multiplied in the epilogue, and split-K wgrad GEMMs (reduction over samples, MN-major operands)
// accumulated with TMA reduce-add; bias gradients are column sums taken in the dgrad epilogue that
// produces dY, the per-ray columns of the fused head layer are dY^T x aux.
#include <string.h>

#include <string>
#include <vector>

#include <stdlib.h>

#include "k2_chain.cuh"
#include "k2_gemm.cuh"

namespace snb {

constexpr int F = 512;    // fc_units (configs/pipelines/*.toml: fc_units = 512)
// feat_last (the width of the head hidden layers): fc_units / 2, or fc_units with fc_use_full_features - snb_model::fl;
// the functions below read it into a local FL
constexpr int PACK_SPLITS = 4;   // K-splits of the folded-layer product W' = W_h1 W_f at pack time
constexpr int LAYERS = 8; // fc_layers, skip at layer 4

struct TensorInfo {
  std::string name;
  int64_t offset;
  int rows, cols;  // cols == 0: bias vector of `rows`
};

struct PackJob {
  long long dst;  // element offset in the destination image
  long long src;  // element offset in the flat fp32 source
  int ldd, lds;
  int rows, cols;  // destination extent
  int transpose;   // dst[i, j] = src[j, i]
  int mode;        // 0 bf16(v)  1 bf16(v - bf16(v))  2 fp32 copy  (pack);   unpack: accumulate
};

}  // namespace snb

struct snb_model {
  int kind, n_classes, sem_sigmoid;
  int fl;   // feat_last: 256, or 512 with SNB_VARIANT_FULL_FEATURES
  int kin0;   // input features of the first trunk layer in the PARAMETERS: 6 x mapping_pos_n_freq (<= k0: K1 always writes the
              // 10-frequency row, the packed weights of the frequencies a model does not have are zero), or 3
  int k0, enc_ld, w0_ld, hhw, n_out, tau;   // enc_ld: K1 row width; w0_ld: packed K of the first layer
  int hh_rgb, hh_beta, hh_sem, hh_bs, hh_sun;
  int beta_s;   // use_separate_beta_for_s: a second uncertainty head (packed column 9, head pre-activation row 6)
  int kho;  // K of the head-output layer: 512 + 256 + hhw
  int relu, aux_ld, kdir;   // vanilla NeRF: ReLU activations, 32-column aux rows carrying the 24 encoded view-direction values
  int variant;              // SNB_VARIANT_* head-input variants of the semantic model (rs_semantic.py:186-215)
  int taux_lo, taux_cols;   // columns of the fused head layer's hidden rows whose blocks take the embedding t as an input
  std::vector<snb::TensorInfo> tensors;
  int64_t n_params;
  // packed bf16 image (element offsets) ----------------------------------------------------------
  long long wl[8], wf, wh1, ws2, ws4, who;           // forward  [N, Kp]
  long long tl[8], tf, th1, ts4, ts2, tho, taux;     // transposed copies for dgrad
  long long packed_bf16_elems;
  long long bias_off;  // fp32 section (byte offset = packed_bf16_elems*2), element offsets below
  long long bl[8], bfe, bs2, bs4, bho;
  long long wf32, wh1_32, bh1_32;   // fp32 copies of W_f [F,F], of the head first layers' f-columns [hhw,F] and biases [hhw]
  long long wp32;                   // fp32 W' = W_h1 W_f: PACK_SPLITS partial products [hhw,F] of the split-K pack product
  long long bias_elems;
  std::vector<snb::PackJob> pack_jobs;
  struct HeadBlock { int row; long long w, b; int kin; };   // hidden blocks of the fused head layer: flat offsets of W / bias
  std::vector<HeadBlock> head_blocks;
  // fp32 packed-gradient scratch (element offsets) ----------------------------------------------------
  long long gl[8], gl4e, gf, gh1, gh1w, gh1aux, gs2, gs4, ghot, gbl[8], gbf, gbs2, gbs4, gbho;
  long long gscratch_elems;
  std::vector<snb::PackJob> unpack_jobs;
  // gradient buckets in the order the backward pass completes them: [0] heads (+ feats, sigma), [1] trunk layers 4-7,
  // [2] trunk layers 0-3; flat element ranges [bucket_lo[b], bucket_hi[b])
  int64_t bucket_lo[3], bucket_hi[3];
  int64_t find(const char* name) const {
    for (auto& t : tensors)
      if (t.name == name) return t.offset;
    return -1;
  }
};

namespace snb {

// ---- pack / unpack kernels ----------------------------------------------------------------------
constexpr int MAX_JOBS = 80;
struct JobTable {
  PackJob jobs[MAX_JOBS];
  int n;
};

__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ JobTable tab, const float* __restrict__ src,
                                                   __nv_bfloat16* __restrict__ dst_bf16, float* __restrict__ dst_f32) {
  const PackJob& j = tab.jobs[blockIdx.y];
  const long long n = (long long)j.rows * j.cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / j.cols), c = (int)(i - (long long)r * j.cols);
    const float v = j.transpose ? src[j.src + (long long)c * j.lds + r] : src[j.src + (long long)r * j.lds + c];
    const long long d = j.dst + (long long)r * j.ldd + c;
    if (j.mode == 2) {
      dst_f32[d] = v;
    } else {
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      dst_bf16[d] = (j.mode == 0) ? hi : __float2bfloat16_rn(v - __bfloat162float(hi));
    }
  }
}

// grads[dst(r,c)] += scratch[src(r,c)]  (same job description, roles of src/dst swapped)
__global__ void __launch_bounds__(256) unpack_kernel(const __grid_constant__ JobTable tab,
                                                     const float* __restrict__ scratch, float* __restrict__ grads) {
  const PackJob& j = tab.jobs[blockIdx.y];
  const long long n = (long long)j.rows * j.cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / j.cols), c = (int)(i - (long long)r * j.cols);
    const float v = j.transpose ? scratch[j.src + (long long)c * j.lds + r] : scratch[j.src + (long long)r * j.lds + c];
    grads[j.dst + (long long)r * j.ldd + c] += v;
  }
}

// g_out (P, n_out) fp32 + out -> gradient w.r.t. the 16 head pre-activations (bf16) + their column sums
__global__ void __launch_bounds__(256)
head_grad_kernel(const float* __restrict__ out, const float* __restrict__ g_out, long long P, int n_out, int C,
                 int sem_sigmoid, int head_mask, int beta_s, __nv_bfloat16* __restrict__ dpre, float* __restrict__ gb_ho) {
  __shared__ float red[16];
  if (threadIdx.x < 16) red[threadIdx.x] = 0.f;
  __syncthreads();
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const float* o = out + p * n_out;
    const float* g = g_out + p * n_out;
    float d[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) d[j] = 0.f;
    if (head_mask & SNB_HEAD_RGB) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float s = (o[j] + 0.001f) * (1.0f / 1.002f);  // sigmoid value
        d[j] = g[j] * 1.002f * s * (1.0f - s);
      }
    }
    // softplus' = sigmoid(x) = 1 - exp(-softplus(x)); expm1 keeps full relative precision for the small outputs of empty space
    if (head_mask & SNB_HEAD_SIGMA) d[3] = g[3] * -expm1f(-o[3]);
    if (head_mask & SNB_HEAD_SUN) d[4] = g[4] * o[4] * (1.0f - o[4]);
    if (head_mask & SNB_HEAD_BETA) d[5] = g[8] * -expm1f(-o[8]);
    if (beta_s && (head_mask & SNB_HEAD_BETA)) d[6] = g[9] * -expm1f(-o[9]);   // semantic uncertainty head (softplus)
    if (head_mask & SNB_HEAD_SEM) {
      const int so = 9 + beta_s, ro = 6 + beta_s;
#pragma unroll
      for (int c = 0; c < 10; ++c)
        if (c < C && ro + c < 16) d[ro + c] = sem_sigmoid ? g[so + c] * o[so + c] * (1.0f - o[so + c]) : g[so + c];
    }
    __align__(16) __nv_bfloat16 row[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      row[j] = __float2bfloat16_rn(d[j]);
      acc[j] += d[j];
    }
    uint4* dst = reinterpret_cast<uint4*>(dpre + p * 16);
    dst[0] = reinterpret_cast<const uint4*>(row)[0];
    dst[1] = reinterpret_cast<const uint4*>(row)[1];
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float v = acc[j];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[j], v);
  }
  __syncthreads();
  if (threadIdx.x < 16) atomicAdd(gb_ho + threadIdx.x, red[threadIdx.x]);
}

// ---- weight-sized fp32 products of the folded feats layer ---------------------------------------------------------------
// C[m, n] (=) sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn]  (+ u[m*su] * (v ? v[n*sv] : 1)), strided operands so that A B, A B^T
// and A^T B all run through one kernel; 64 x 64 tiles, 16-deep k-steps, 256 threads with 4 x 4 outputs each.  Outputs: fp32 C
// and / or bf16 o16[m*ld16 + n] and its transpose o16t[n*ld16t + m] (the packed forward / dgrad copies of W').
struct SmallGemm {
  const float* A; long long sam, sak;
  const float* B; long long sbk, sbn;
  int M, N, K;
  float* C; long long ldc;
  const float* u; long long su; const float* v; long long sv;
  __nv_bfloat16* o16; long long ld16; __nv_bfloat16* o16t; long long ld16t;
  int splits;   // > 1: the K range is split over blockIdx.z and C is accumulated with atomics (C zeroed by the caller; fp32 C only)
                // 0: chosen by small_gemm so that the grid fills the SMs a few times over (these products are latency-bound:
                // one 64 x 64 tile per SM leaves every global load exposed)
  long long slice;   // != 0: split z STORES its partial product to C + z * slice instead (deterministic; the reader sums them)
};

__global__ void __launch_bounds__(256) small_gemm_kernel(const SmallGemm g) {
  __shared__ __align__(16) float As[2][16][64 + 4], Bs[2][16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int kper = ((g.K + g.splits - 1) / g.splits + 15) & ~15;
  const int kbeg = blockIdx.z * kper, kend = min(g.K, kbeg + kper);
  // the fastest thread index runs along the unit-stride dimension of each operand
  const bool a_kfast = g.sak == 1, b_kfast = g.sbk == 1;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kk = a_kfast ? (tid & 15) : ((tid >> 6) + 4 * r), mm = a_kfast ? ((tid >> 4) + 16 * r) : (tid & 63);
      ra[r] = (m0 + mm < g.M && k0 + kk < kend) ? __ldg(g.A + (long long)(m0 + mm) * g.sam + (long long)(k0 + kk) * g.sak) : 0.f;
      const int kb = b_kfast ? (tid & 15) : ((tid >> 6) + 4 * r), nn = b_kfast ? ((tid >> 4) + 16 * r) : (tid & 63);
      rb[r] = (n0 + nn < g.N && k0 + kb < kend) ? __ldg(g.B + (long long)(k0 + kb) * g.sbk + (long long)(n0 + nn) * g.sbn) : 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kk = a_kfast ? (tid & 15) : ((tid >> 6) + 4 * r), mm = a_kfast ? ((tid >> 4) + 16 * r) : (tid & 63);
      As[buf][kk][mm] = ra[r];
      const int kb = b_kfast ? (tid & 15) : ((tid >> 6) + 4 * r), nn = b_kfast ? ((tid >> 4) + 16 * r) : (tid & 63);
      Bs[buf][kb][nn] = rb[r];
    }
  };
  if (kbeg < kend) {
    fetch(kbeg);
    stash(0);
  }
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += 16, buf ^= 1) {
    const bool more = k0 + 16 < kend;
    if (more) fetch(k0 + 16);          // the next tile's global loads are in flight while this one is multiplied
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float c = acc[i][j];
      if (g.u != nullptr && blockIdx.z == 0) c = fmaf(g.u[(long long)m * g.su], g.v ? g.v[(long long)n * g.sv] : 1.0f, c);
      if (g.C != nullptr) {
        if (g.slice != 0) g.C[(long long)blockIdx.z * g.slice + (long long)m * g.ldc + n] = c;
        else if (g.splits > 1) atomicAdd(g.C + (long long)m * g.ldc + n, c);
        else g.C[(long long)m * g.ldc + n] = c;
      }
      if (g.o16 != nullptr) g.o16[(long long)m * g.ld16 + n] = __float2bfloat16_rn(c);
      if (g.o16t != nullptr) g.o16t[(long long)n * g.ld16t + m] = __float2bfloat16_rn(c);
    }
  }
}

static int small_gemm(SmallGemm g, cudaStream_t st) {
  if (g.C == nullptr || g.o16 != nullptr || g.o16t != nullptr) g.splits = 1;
  if (g.splits < 1) {
    const int tiles = ((g.N + 63) / 64) * ((g.M + 63) / 64);
    int s = (4 * 148 + tiles - 1) / tiles;
    const int max_s = g.K / 64 > 0 ? g.K / 64 : 1;      // at least four 16-deep k-steps per split
    g.splits = s < 1 ? 1 : (s > max_s ? max_s : s);
  }
  dim3 grid((g.N + 63) / 64, (g.M + 63) / 64, g.splits);
  small_gemm_kernel<<<grid, 256, 0, st>>>(g);
  return launch_status("small_gemm_kernel");
}

// y[m] (=) sum_k A[m*sam + k*sak] * x[k*sx] + (u ? u[m] : 0): one warp per row (the bias rows of the folded feats layer)
__global__ void __launch_bounds__(256) small_gemv_kernel(const float* __restrict__ A, long long sam, long long sak,
                                                         const float* __restrict__ x, long long sx, int M, int K,
                                                         const float* __restrict__ u, float* __restrict__ y, long long sy,
                                                         __nv_bfloat16* __restrict__ y16, long long sy16) {
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(A[(long long)m * sam + (long long)k * sak], x[(long long)k * sx], acc);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if (lane == 0) {
    if (u != nullptr) acc += u[m];
    if (y != nullptr) y[(long long)m * sy] = acc;
    if (y16 != nullptr) y16[(long long)m * sy16] = __float2bfloat16_rn(acc);
  }
}

// the same product for a matrix stored with m as the unit-stride index (A^T x): lanes run along m, 8 k-groups per block and
// blockIdx.y split the reduction, y is accumulated with atomics (y zeroed by the caller)
__global__ void __launch_bounds__(256) small_gemv_t_kernel(const float* __restrict__ A, long long sak, const float* __restrict__ x,
                                                           long long sx, int M, int K, float* __restrict__ y, long long sy) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, kg = threadIdx.x >> 5;
  const int m = blockIdx.x * 32 + lane;
  const int kper = (K + gridDim.y - 1) / gridDim.y;
  const int kbeg = blockIdx.y * kper, kend = min(K, kbeg + kper);
  float acc = 0.f;
  if (m < M)
    for (int k = kbeg + kg; k < kend; k += 8) acc = fmaf(A[(long long)k * sak + m], x[(long long)k * sx], acc);
  red[kg][lane] = acc;
  __syncthreads();
  if (kg == 0 && m < M) {
#pragma unroll
    for (int i = 1; i < 8; ++i) acc += red[i][lane];
    atomicAdd(y + (long long)m * sy, acc);
  }
}

// W' (fp32, [rows, F], the sum of `slices` partial products `slice` floats apart, added in a fixed order: the packed image
// is a deterministic function of the parameters) -> its bf16 copies: o16[r*ld16 + c] (forward rows) and o16t[c*ld16t + r]
// (dgrad rows), 32 x 32 tiles
__global__ void __launch_bounds__(256) wprime_bf16_kernel(const float* __restrict__ W, int slices, long long slice, int rows, int cols,
                                                          __nv_bfloat16* __restrict__ o16, long long ld16,
                                                          __nv_bfloat16* __restrict__ o16t, long long ld16t) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    float v = 0.f;
    if (r < rows && c < cols)
      for (int z = 0; z < slices; ++z) v += W[(long long)z * slice + (long long)r * cols + c];
    tile[ty + 8 * i][tx] = v;
    if (r < rows && c < cols) o16[(long long)r * ld16 + c] = __float2bfloat16_rn(v);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;
    if (r < rows && c < cols) o16t[(long long)c * ld16t + r] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
  }
}

static int small_gemv(const float* A, long long sam, long long sak, const float* x, long long sx, int M, int K, const float* u,
                      float* y, long long sy, __nv_bfloat16* y16, long long sy16, cudaStream_t st) {
  if (sam == 1 && u == nullptr && y16 == nullptr && y != nullptr) {
    dim3 grid((M + 31) / 32, K >= 512 ? 8 : (K >= 128 ? 4 : 1));
    small_gemv_t_kernel<<<grid, 256, 0, st>>>(A, sak, x, sx, M, K, y, sy);
    return launch_status("small_gemv_t_kernel");
  }
  small_gemv_kernel<<<(M + 7) / 8, 256, 0, st>>>(A, sam, sak, x, sx, M, K, u, y, sy, y16, sy16);
  return launch_status("small_gemv_kernel");
}

// ---- model description ----------------------------------------------------------------------------
static void add_tensor(snb_model* m, const std::string& name, int rows, int cols) {
  m->tensors.push_back({name, m->n_params, rows, cols});
  m->n_params += (int64_t)rows * (cols ? cols : 1);
}

static long long take(long long& cursor, long long elems) {
  long long o = cursor;
  cursor += (elems + 63) & ~63ll;  // keep every matrix 128-byte aligned
  return o;
}

static void build_layout(snb_model* m) {
  const int FL = m->fl;
  const int k0 = m->k0, kin0 = m->kin0, hhw = m->hhw, tau = m->tau, C = m->n_classes;
  const bool sem = m->kind == SNB_MODEL_SEMANTIC;
  const bool nerf = m->kind == SNB_MODEL_NERF;   // nerf.py:118-160: trunk, sigma, feats, rgb(f | dir) only
  const bool has_beta = m->kind == SNB_MODEL_SATNERF || sem;   // S-NeRF (snerf.py:161-186) and NeRF have no uncertainty head
  const bool enc60 = k0 == 60;                   // positional encoding input: the [hi | lo | 0] 128-column K1 row
  // head-input variants (semantic model only): the transient embedding t as an extra input of the semantic head
  // (`use_tj_for_s`, rs_semantic.py:207-211,330-338) and / or of the colour head (`use_tj_instead_of_beta`, :186-189,287-288):
  // like the uncertainty head's cat(f, t) (:251-252) they are 4 more weight columns against the aux K-segment [1, sun_d, t]
  const bool tj_s = sem && (m->variant & SNB_VARIANT_TJ_FOR_S);
  const bool tj_rgb = sem && (m->variant & SNB_VARIANT_TJ_INSTEAD_OF_BETA);
  const bool bs = m->beta_s != 0;   // semantic_beta_from_xyz (rs_semantic.py:228-237): cat(f, t) -> 256 -> 1, softplus
  // use_separate_tj_for_semantic (rs_semantic.py:300-301,334-335): the semantic head and the semantic uncertainty head read
  // a SECOND embedding t_s: aux columns 8..11 (the caller passes the two tables side by side as one (vocab, 8) table)
  const int ts_col = (sem && (m->variant & SNB_VARIANT_SEPARATE_TJ_S)) ? 4 + tau : 4;   // aux = [1 | sun_d | t | t_s]
  // ---- flat fp32 parameter order = reference state_dict order ----
  m->n_params = 0;
  for (int i = 0; i < LAYERS; ++i) {
    int kin = i == 0 ? m->kin0 : (i == 4 ? F + m->kin0 : F);
    add_tensor(m, "fc_net." + std::to_string(2 * i) + ".weight", F, kin);
    add_tensor(m, "fc_net." + std::to_string(2 * i) + ".bias", F, 0);
  }
  add_tensor(m, "sigma_from_xyz.0.weight", 1, F);
  add_tensor(m, "sigma_from_xyz.0.bias", 1, 0);
  add_tensor(m, "feats_from_xyz.weight", F, F);
  add_tensor(m, "feats_from_xyz.bias", F, 0);
  add_tensor(m, "rgb_from_xyzdir.0.weight", FL, F + (nerf ? m->kdir : 0) + (tj_rgb ? tau : 0));
  add_tensor(m, "rgb_from_xyzdir.0.bias", FL, 0);
  add_tensor(m, "rgb_from_xyzdir.2.weight", 3, FL);
  add_tensor(m, "rgb_from_xyzdir.2.bias", 3, 0);
  if (sem) {
    add_tensor(m, "semantic_prediction.0.weight", FL, F + (tj_s ? tau : 0));
    add_tensor(m, "semantic_prediction.0.bias", FL, 0);
    add_tensor(m, "semantic_prediction.2.weight", C, FL);
    add_tensor(m, "semantic_prediction.2.bias", C, 0);
  }
  if (!nerf) {
    add_tensor(m, "sun_v_net.0.weight", FL, F + 3);
    add_tensor(m, "sun_v_net.0.bias", FL, 0);
    add_tensor(m, "sun_v_net.2.weight", FL, FL);
    add_tensor(m, "sun_v_net.2.bias", FL, 0);
    add_tensor(m, "sun_v_net.4.weight", FL, FL);
    add_tensor(m, "sun_v_net.4.bias", FL, 0);
    add_tensor(m, "sun_v_net.6.weight", 1, FL);
    add_tensor(m, "sun_v_net.6.bias", 1, 0);
    add_tensor(m, "sky_color.0.weight", FL, 3);
    add_tensor(m, "sky_color.0.bias", FL, 0);
    add_tensor(m, "sky_color.2.weight", 3, FL);
    add_tensor(m, "sky_color.2.bias", 3, 0);
  }
  if (has_beta) {
    add_tensor(m, "beta_from_xyz.0.weight", FL, F + tau);
    add_tensor(m, "beta_from_xyz.0.bias", FL, 0);
    add_tensor(m, "beta_from_xyz.2.weight", 1, FL);
    add_tensor(m, "beta_from_xyz.2.bias", 1, 0);
  }
  if (bs) {
    add_tensor(m, "semantic_beta_from_xyz.0.weight", FL, F + tau);
    add_tensor(m, "semantic_beta_from_xyz.0.bias", FL, 0);
    add_tensor(m, "semantic_beta_from_xyz.2.weight", 1, FL);
    add_tensor(m, "semantic_beta_from_xyz.2.bias", 1, 0);
  }

  auto P = [&](const std::string& n) { return m->find(n.c_str()); };
  auto fcw = [&](int i) { return P("fc_net." + std::to_string(2 * i) + ".weight"); };
  auto fcb = [&](int i) { return P("fc_net." + std::to_string(2 * i) + ".bias"); };

  // ---- packed bf16 image ----
  long long cur = 0;
  const int kl4 = 64 + F;       // packed K of layer 4: [enc(64) | h3(512)]
  const int kh1 = F + 64;       // packed K of the fused head first layers: [f(512) | aux(64)]
  const int ktf = hhw + 64;     // packed K of the dgrad into h7: [dY_hh(hhw) | dPre16(64)] x [W' ; w_sigma] (feats folded in)
  m->kho = F + FL + hhw;        // [h7 | s3 | hh]
  for (int i = 0; i < LAYERS; ++i) m->wl[i] = take(cur, (long long)F * (i == 0 ? m->w0_ld : (i == 4 ? kl4 : F)));
  m->wf = -1;   // feats_from_xyz is folded into the head first layers (see the header): no forward copy of its own
  m->wh1 = take(cur, (long long)hhw * kh1);
  m->ws2 = take(cur, (long long)FL * FL);
  m->ws4 = take(cur, (long long)FL * FL);
  m->who = take(cur, 16ll * m->kho);
  for (int i = 1; i < LAYERS; ++i) m->tl[i] = take(cur, (long long)F * F);
  m->tl[0] = -1;
  m->tf = take(cur, (long long)F * ktf);
  m->th1 = -1;  // (its role - the dgrad through the head first layers - is the W' block of tf)
  m->ts4 = take(cur, (long long)FL * FL);
  m->ts2 = take(cur, (long long)FL * FL);
  m->tho = take(cur, (long long)(FL + hhw) * 16);
  // blocks with a t input: [rgb | beta | sem] are the first three 256-column blocks of the hidden rows
  m->taux_lo = tj_rgb ? 0 : FL;
  m->taux_cols = (bs ? 4 * FL : (tj_s ? 3 * FL : 2 * FL)) - m->taux_lo;
  m->taux = take(cur, 16ll * m->taux_cols);
  m->packed_bf16_elems = cur;
  long long bc = 0;
  for (int i = 0; i < LAYERS; ++i) m->bl[i] = take(bc, F);
  m->bfe = take(bc, F);
  m->bs2 = take(bc, FL);
  m->bs4 = take(bc, FL);
  m->bho = take(bc, 16);
  m->wf32 = take(bc, (long long)F * F);
  m->wh1_32 = take(bc, (long long)hhw * F);
  m->bh1_32 = take(bc, hhw);
  m->wp32 = take(bc, (long long)PACK_SPLITS * hhw * F);
  m->bias_elems = bc;

  auto& J = m->pack_jobs;
  auto job = [&](long long dst, int ldd, long long src, int lds, int rows, int cols, int tr, int mode) {
    J.push_back({dst, src, ldd, lds, rows, cols, tr, mode});
  };
  if (enc60) {
    // layer 0, semantic: K-segments [enc cols 0..127 = hi | lo | 0] x [W_hi | W_hi | 0] and [enc cols 0..63 = hi | lo(0:4)] x [W_lo | 0]
    // (a model with fewer than 10 frequencies fills the first kin0 = 6 L columns of each 60-wide slot; the rest stay zero)
    job(m->wl[0], m->w0_ld, fcw(0), kin0, F, kin0, 0, 0);
    job(m->wl[0] + k0, m->w0_ld, fcw(0), kin0, F, kin0, 0, 0);
    job(m->wl[0] + 128, m->w0_ld, fcw(0), kin0, F, kin0, 0, 1);
  } else {
    // layer 0, satnerf: [hi | lo | hi] against the input row [hi | hi | lo]
    job(m->wl[0], m->w0_ld, fcw(0), k0, F, k0, 0, 0);
    job(m->wl[0] + k0, m->w0_ld, fcw(0), k0, F, k0, 0, 1);
    job(m->wl[0] + 2 * k0, m->w0_ld, fcw(0), k0, F, k0, 0, 0);
  }
  for (int i = 1; i < LAYERS; ++i) {
    if (i == 4) {
      job(m->wl[4], kl4, fcw(4), F + kin0, F, kin0, 0, 0);            // enc columns (first kin0 of the 64-wide segment)
      job(m->wl[4] + 64, kl4, fcw(4) + kin0, F + kin0, F, F, 0, 0);   // h3 columns
      job(m->tl[4], F, fcw(4) + kin0, F + kin0, F, F, 1, 0);
    } else {
      job(m->wl[i], F, fcw(i), F, F, F, 0, 0);
      job(m->tl[i], F, fcw(i), F, F, F, 1, 0);
    }
  }
  // tf = [W'^T (composed at pack time, snb_model_pack) | dPre16 block]: dPre16 column 3 = sigma
  job(m->tf + hhw + 3, ktf, P("sigma_from_xyz.0.weight"), F, F, 1, 1, 0);
  job(m->wf32, F, P("feats_from_xyz.weight"), F, F, F, 0, 2);            // fp32 copy for the weight-sized gradient products
  // fused head first layers: rows [rgb | beta | (sem) | sun]  (NeRF: the rgb block only; the others stay zero)
  struct Blk { int row; const char* w; const char* b; int kin; bool t_in; int t_col; };
  std::vector<Blk> blks = {{m->hh_rgb, "rgb_from_xyzdir.0", "rgb_from_xyzdir.0", F + (nerf ? m->kdir : 0) + (tj_rgb ? tau : 0), tj_rgb, 4}};
  if (has_beta) blks.push_back({m->hh_beta, "beta_from_xyz.0", "beta_from_xyz.0", F + tau, true, 4});
  if (!nerf) blks.push_back({m->hh_sun, "sun_v_net.0", "sun_v_net.0", F + 3, false, 4});
  if (sem) blks.push_back({m->hh_sem, "semantic_prediction.0", "semantic_prediction.0", F + (tj_s ? tau : 0), tj_s, ts_col});
  if (bs) blks.push_back({m->hh_bs, "semantic_beta_from_xyz.0", "semantic_beta_from_xyz.0", F + tau, true, ts_col});
  for (auto& b : blks) {
    const long long w = P(std::string(b.w) + ".weight"), bb = P(std::string(b.b) + ".bias");
    // the f-columns W' = W_b W_f (+ their transpose in tf) and the bias column b' = b_b + W_b b_f (aux column 0 = 1) are
    // composed in fp32 by snb_model_pack; here only the fp32 copy of W_b's f-columns for the gradient products
    job(m->wh1_32 + (long long)b.row * F, F, w, b.kin, FL, F, 0, 2);
    job(m->bh1_32 + b.row, 1, bb, 1, FL, 1, 0, 2);
    m->head_blocks.push_back({b.row, w, bb, b.kin});
  }
  if (nerf) {   // aux columns 1..24 = the encoded view direction (cat(f, Mapping(dir)), nerf.py:197-199)
    job(m->wh1 + (long long)m->hh_rgb * kh1 + F + 1, kh1, P("rgb_from_xyzdir.0.weight") + F, F + m->kdir, FL, m->kdir, 0, 0);
  } else {
    job(m->wh1 + (long long)m->hh_sun * kh1 + F + 1, kh1, P("sun_v_net.0.weight") + F, F + 3, FL, 3, 0, 0);
    for (auto& b : blks)
      if (b.t_in) {   // aux columns 4..4+tau = t (8.. = t_s): the block's last tau weight columns; their transpose feeds the dgrad into t
        const long long w = P(std::string(b.w) + ".weight");
        job(m->wh1 + (long long)b.row * kh1 + F + b.t_col, kh1, w + F, b.kin, FL, tau, 0, 0);
        job(m->taux + (long long)b.t_col * m->taux_cols + (b.row - m->taux_lo), m->taux_cols, w + F, b.kin, tau, FL, 1, 0);
      }
    job(m->ws2, FL, P("sun_v_net.2.weight"), FL, FL, FL, 0, 0);
    job(m->ts2, FL, P("sun_v_net.2.weight"), FL, FL, FL, 1, 0);
    job(m->ws4, FL, P("sun_v_net.4.weight"), FL, FL, FL, 0, 0);
    job(m->ts4, FL, P("sun_v_net.4.weight"), FL, FL, FL, 1, 0);
  }
  // head output layer: rows [rgb0-2, sigma, sun, beta, sem...], K = [h7 | s3 | hh]
  const int kho = m->kho;
  job(m->who + 3ll * kho, kho, P("sigma_from_xyz.0.weight"), F, 1, F, 0, 0);
  if (!nerf) job(m->who + 4ll * kho + F, kho, P("sun_v_net.6.weight"), FL, 1, FL, 0, 0);
  job(m->who + 0ll * kho + F + FL + m->hh_rgb, kho, P("rgb_from_xyzdir.2.weight"), FL, 3, FL, 0, 0);
  if (has_beta) job(m->who + 5ll * kho + F + FL + m->hh_beta, kho, P("beta_from_xyz.2.weight"), FL, 1, FL, 0, 0);
  const int ro = 6 + (bs ? 1 : 0);   // first semantic row of the 16 head pre-activations (row 6 = the semantic uncertainty head)
  if (bs) job(m->who + 6ll * kho + F + FL + m->hh_bs, kho, P("semantic_beta_from_xyz.2.weight"), FL, 1, FL, 0, 0);
  if (sem) job(m->who + (long long)ro * kho + F + FL + m->hh_sem, kho, P("semantic_prediction.2.weight"), FL, C, FL, 0, 0);
  // transposed head output for dgrad: rows = [s3 | hh] features, 16 columns
  if (!nerf) job(m->tho + 4, 16, P("sun_v_net.6.weight"), FL, FL, 1, 1, 0);
  job(m->tho + (long long)(FL + m->hh_rgb) * 16 + 0, 16, P("rgb_from_xyzdir.2.weight"), FL, FL, 3, 1, 0);
  if (has_beta) job(m->tho + (long long)(FL + m->hh_beta) * 16 + 5, 16, P("beta_from_xyz.2.weight"), FL, FL, 1, 1, 0);
  if (bs) job(m->tho + (long long)(FL + m->hh_bs) * 16 + 6, 16, P("semantic_beta_from_xyz.2.weight"), FL, FL, 1, 1, 0);
  if (sem) job(m->tho + (long long)(FL + m->hh_sem) * 16 + ro, 16, P("semantic_prediction.2.weight"), FL, FL, C, 1, 0);
  // fp32 biases
  for (int i = 0; i < LAYERS; ++i) job(m->bl[i], 1, fcb(i), 1, F, 1, 0, 2);
  job(m->bfe, 1, P("feats_from_xyz.bias"), 1, F, 1, 0, 2);
  if (!nerf) job(m->bs2, 1, P("sun_v_net.2.bias"), 1, FL, 1, 0, 2);
  if (!nerf) job(m->bs4, 1, P("sun_v_net.4.bias"), 1, FL, 1, 0, 2);
  job(m->bho + 0, 1, P("rgb_from_xyzdir.2.bias"), 1, 3, 1, 0, 2);
  job(m->bho + 3, 1, P("sigma_from_xyz.0.bias"), 1, 1, 1, 0, 2);
  if (!nerf) job(m->bho + 4, 1, P("sun_v_net.6.bias"), 1, 1, 1, 0, 2);
  if (has_beta) job(m->bho + 5, 1, P("beta_from_xyz.2.bias"), 1, 1, 1, 0, 2);
  if (bs) job(m->bho + 6, 1, P("semantic_beta_from_xyz.2.bias"), 1, 1, 1, 0, 2);
  if (sem) job(m->bho + ro, 1, P("semantic_prediction.2.bias"), 1, C, 1, 0, 2);

  // ---- fp32 packed-gradient scratch + unpack jobs (grads[dst] += scratch[src]) ----
  long long gc = 0;
  for (int i = 0; i < LAYERS; ++i) m->gl[i] = take(gc, (long long)F * (i == 0 ? 64 : F));
  m->gl4e = take(gc, (long long)F * 64);
  m->gf = take(gc, (long long)F * F);
  m->gh1 = take(gc, (long long)hhw * F);    // G1 = dY_hh^T h7
  m->gh1w = take(gc, (long long)hhw * F);   // dW_h1 = G1 W_f^T + colsum(dY_hh) b_f^T
  const int gald = nerf ? 64 : 16;   // row length of the dY^T x aux block (the wgrad side operand's column count)
  m->gh1aux = take(gc, (long long)hhw * gald);
  m->gs2 = take(gc, (long long)FL * FL);
  m->gs4 = take(gc, (long long)FL * FL);
  m->ghot = take(gc, (long long)kho * 16);
  for (int i = 0; i < LAYERS; ++i) m->gbl[i] = take(gc, (long long)F * 16);
  m->gbf = take(gc, (long long)F * 16);
  m->gbs2 = take(gc, (long long)FL * 16);
  m->gbs4 = take(gc, (long long)FL * 16);
  m->gbho = take(gc, 16);
  m->gscratch_elems = gc;
  auto& U = m->unpack_jobs;
  auto ujob = [&](long long dst, int ldd, long long src, int lds, int rows, int cols, int tr) {
    U.push_back({dst, src, ldd, lds, rows, cols, tr, 0});
  };
  ujob(fcw(0), kin0, m->gl[0], 64, F, kin0, 0);
  for (int i = 1; i < LAYERS; ++i) {
    if (i == 4) {
      ujob(fcw(4), F + kin0, m->gl4e, 64, F, kin0, 0);
      ujob(fcw(4) + kin0, F + kin0, m->gl[4], F, F, F, 0);
    } else {
      ujob(fcw(i), F, m->gl[i], F, F, F, 0);
    }
  }
  for (int i = 0; i < LAYERS; ++i) ujob(fcb(i), 1, m->gbl[i], 1, F, 1, 0);
  ujob(P("feats_from_xyz.weight"), F, m->gf, F, F, F, 0);
  ujob(P("feats_from_xyz.bias"), 1, m->gbf, 1, F, 1, 0);
  for (auto& b : blks) {
    const long long w = P(std::string(b.w) + ".weight"), bb = P(std::string(b.b) + ".bias");
    ujob(w, b.kin, m->gh1w + (long long)b.row * F, F, FL, F, 0);
    ujob(bb, 1, m->gh1aux + (long long)b.row * gald, gald, FL, 1, 0);
  }
  if (nerf) {
    ujob(P("rgb_from_xyzdir.0.weight") + F, F + m->kdir, m->gh1aux + (long long)m->hh_rgb * gald + 1, gald, FL, m->kdir, 0);
  } else {
    ujob(P("sun_v_net.0.weight") + F, F + 3, m->gh1aux + (long long)m->hh_sun * 16 + 1, 16, FL, 3, 0);
    for (auto& b : blks)
      if (b.t_in) ujob(P(std::string(b.w) + ".weight") + F, b.kin, m->gh1aux + (long long)b.row * 16 + b.t_col, 16, FL, tau, 0);
    ujob(P("sun_v_net.2.weight"), FL, m->gs2, FL, FL, FL, 0);
    ujob(P("sun_v_net.2.bias"), 1, m->gbs2, 1, FL, 1, 0);
    ujob(P("sun_v_net.4.weight"), FL, m->gs4, FL, FL, FL, 0);
    ujob(P("sun_v_net.4.bias"), 1, m->gbs4, 1, FL, 1, 0);
  }
  // head output layer: scratch is transposed [K features, 16]
  ujob(P("sigma_from_xyz.0.weight"), F, m->ghot + 3, 16, 1, F, 1);
  if (!nerf) ujob(P("sun_v_net.6.weight"), FL, m->ghot + (long long)F * 16 + 4, 16, 1, FL, 1);
  ujob(P("rgb_from_xyzdir.2.weight"), FL, m->ghot + (long long)(F + FL + m->hh_rgb) * 16 + 0, 16, 3, FL, 1);
  if (has_beta) ujob(P("beta_from_xyz.2.weight"), FL, m->ghot + (long long)(F + FL + m->hh_beta) * 16 + 5, 16, 1, FL, 1);
  if (bs) ujob(P("semantic_beta_from_xyz.2.weight"), FL, m->ghot + (long long)(F + FL + m->hh_bs) * 16 + 6, 16, 1, FL, 1);
  if (sem) ujob(P("semantic_prediction.2.weight"), FL, m->ghot + (long long)(F + FL + m->hh_sem) * 16 + ro, 16, C, FL, 1);
  ujob(P("rgb_from_xyzdir.2.bias"), 1, m->gbho + 0, 1, 3, 1, 0);
  ujob(P("sigma_from_xyz.0.bias"), 1, m->gbho + 3, 1, 1, 1, 0);
  if (!nerf) ujob(P("sun_v_net.6.bias"), 1, m->gbho + 4, 1, 1, 1, 0);
  if (has_beta) ujob(P("beta_from_xyz.2.bias"), 1, m->gbho + 5, 1, 1, 1, 0);
  if (bs) ujob(P("semantic_beta_from_xyz.2.bias"), 1, m->gbho + 6, 1, 1, 1, 0);
  if (sem) ujob(P("semantic_prediction.2.bias"), 1, m->gbho + ro, 1, C, 1, 0);
  m->bucket_lo[2] = 0;
  m->bucket_hi[2] = m->bucket_lo[1] = fcw(4);
  m->bucket_hi[1] = m->bucket_lo[0] = P("sigma_from_xyz.0.weight");
  m->bucket_hi[0] = m->n_params;
}

// lo/hi: only the jobs whose flat-parameter offset (unpack: dst) lies in [lo, hi)
static int run_jobs(const std::vector<PackJob>& all, bool unpack, const float* src, void* dst_bf16, float* dst_f32,
                    cudaStream_t st, long long lo = 0, long long hi = -1) {
  std::vector<PackJob> sel;
  if (hi >= 0) {
    for (auto& j : all)
      if (j.dst >= lo && j.dst < hi) sel.push_back(j);
  }
  const std::vector<PackJob>& jobs = hi >= 0 ? sel : all;
  for (size_t base = 0; base < jobs.size(); base += MAX_JOBS) {
    JobTable tab;
    tab.n = (int)std::min<size_t>(MAX_JOBS, jobs.size() - base);
    for (int i = 0; i < tab.n; ++i) tab.jobs[i] = jobs[base + i];
    dim3 grid(256, tab.n);   // the largest jobs are 512 x 512: 4 elements per thread (64 blocks left most SMs idle for 14 us)
    if (unpack) unpack_kernel<<<grid, 256, 0, st>>>(tab, src, dst_f32);
    else pack_kernel<<<grid, 256, 0, st>>>(tab, src, (__nv_bfloat16*)dst_bf16, dst_f32);
    if (int r = launch_status(unpack ? "unpack_kernel" : "pack_kernel")) return r;
  }
  return 0;
}

// ---- workspace layout -----------------------------------------------------------------------------
struct Workspace {
  // all offsets in bytes from the workspace base; 0-sized members are absent
  size_t h[8], sg[8], hh, sghh, s2, sgs2, s3, sgs3;  // forward (sg*: sign masks of the SIREN derivatives, train only)
  size_t dpre, dy[8], dyhh, dys3, dys2;             // backward (dy[i] = gradient w.r.t. the pre-activation of trunk layer i)
  size_t scr_h[2], scr_s2;                          // inference: per-SM-pair scratch of the chained kernel (L2-resident)
  size_t hpart;                                     // (P, 16) fp32 partial sums of the head pre-activations
  size_t auxT, encT;                                // [16, ldt] / [64, ldt] K-major copies of aux / enc[:, :64] (wgrad side operands)
  size_t gscratch;                                  // fp32 packed gradients
  size_t total;
};

static int g_chain = -1;   // -1: from the environment (SNB_CHAIN, default on)
static bool use_chain() {
  if (g_chain < 0) {
    const char* e = getenv("SNB_CHAIN");
    g_chain = e ? (atoi(e) != 0) : 1;
  }
  return g_chain != 0;
}

static Workspace layout_workspace(const snb_model* m, int64_t P, int train) {
  const int FL = m->fl;
  Workspace w;
  memset(&w, 0, sizeof(w));
  size_t cur = 0;
  auto take_b = [&](size_t bytes) {
    size_t o = cur;
    cur += (bytes + 1023) & ~(size_t)1023;
    return o;
  };
  const size_t rowF = (size_t)P * F * 2, rowFL = (size_t)P * FL * 2, rowHH = (size_t)P * m->hhw * 2;
  if (train) {
    for (int i = 0; i < 8; ++i) w.h[i] = take_b(rowF);
    for (int i = 0; i < 8; ++i) w.sg[i] = take_b((size_t)P * (F / 8));   // 1 bit per element
  } else {
    size_t a = take_b(rowF), b = take_b(rowF);
    for (int i = 0; i < 8; ++i) w.h[i] = (i & 1) ? b : a;  // ping-pong (layer 4 reads h3, writes h4: distinct)
    // chained kernel: layers 0..6 and s2 live in a per-pair scratch instead (h7 / hh / s3 stay per-point:
    // the head-output kernel reads them)
    const size_t R = (size_t)chain_scratch_rows();
    w.scr_h[0] = take_b(R * F * 2);
    w.scr_h[1] = take_b(R * F * 2);
    w.scr_s2 = take_b(R * FL * 2);
  }
  w.hpart = take_b((size_t)P * 16 * 4);
  w.hh = take_b(rowHH);
  w.s2 = take_b(rowFL);
  w.s3 = take_b(rowFL);
  if (train) {
    w.sghh = take_b((size_t)P * (m->hhw / 8));
    w.sgs2 = take_b((size_t)P * (FL / 8));
    w.sgs3 = take_b((size_t)P * (FL / 8));
    w.dpre = take_b((size_t)P * 16 * 2);
    for (int i = 0; i < 8; ++i) w.dy[i] = take_b(rowF);
    w.dyhh = take_b(rowHH);
    w.dys3 = take_b(rowFL);
    w.dys2 = take_b(rowFL);
    const size_t ldt = ((size_t)P + 63) & ~(size_t)63;
    w.auxT = take_b(64 * ldt * 2);
    w.encT = take_b(64 * ldt * 2);
    w.gscratch = take_b((size_t)m->gscratch_elems * 4);
  }
  w.total = cur;
  return w;
}

// The same layout seen from row r on: every per-point member moved by r rows.  A training step keeps the points of the
// solar-correction pass in rows [P, P + Psc) of ONE workspace (snb_mlp_forward_with_solar): the two passes run their own
// chained launches over their row ranges, and the weight gradients of the layers both passes share run once over all rows.
static Workspace shift_rows(const snb_model* m, Workspace w, long long r) {
  const int FL = m->fl;
  const size_t n = (size_t)r;
  for (int i = 0; i < 8; ++i) {
    w.h[i] += n * F * 2;
    w.sg[i] += n * (F / 8);
    w.dy[i] += n * F * 2;
  }
  w.hh += n * m->hhw * 2;
  w.dyhh += n * m->hhw * 2;
  w.sghh += n * (m->hhw / 8);
  w.s2 += n * FL * 2;  w.s3 += n * FL * 2;  w.dys2 += n * FL * 2;  w.dys3 += n * FL * 2;
  w.sgs2 += n * (FL / 8);  w.sgs3 += n * (FL / 8);
  w.dpre += n * 16 * 2;
  w.hpart += n * 16 * 4;
  return w;
}

// ---- GEMM plan builders -----------------------------------------------------------------------------
struct Seg {
  const void* ptr;   // bf16, row-major, first column of the segment
  long long ld;      // leading dimension (elements)
  int cols;          // readable columns from ptr (TMA zero-fills beyond)
  int kb;            // 64-wide k-blocks this segment contributes
};

struct Plan {
  std::vector<GemmArgs> g;
  std::vector<int> epi;
  int rc = 0;
  GemmArgs& add(int e) {
    g.emplace_back();
    memset(&g.back(), 0, sizeof(GemmArgs));
    epi.push_back(e);
    return g.back();
  }
  void chk(int r) {
    if (r && !rc) rc = r;
  }
};

// D[M,16] = sum_seg A_seg[M,K] * B[16,Kp]^T  (K-major operands; the N = 16 fp32-row epilogue F32ROWS)
static GemmArgs& add_rows16(Plan& p, int epi, long long M, const Seg* segs, int nseg, const void* B, long long ldb, int b_cols,
                            const float* bias) {
  GemmArgs& a = p.add(epi);
  a.M = (int)M;
  a.N = 16;
  a.block_n = 16;
  a.cta_group = 1;
  a.nseg = nseg;
  a.kb_total = 0;
  for (int s = 0; s < nseg; ++s) {
    a.seg_kb[s] = segs[s].kb;
    a.kb_total += segs[s].kb;
    p.chk(make_tmap_2d(&a.tmA[s], segs[s].ptr, 2, (uint64_t)segs[s].cols, (uint64_t)M, (uint64_t)segs[s].ld * 2, 64,
                       GEMM_BLOCK_M));
  }
  p.chk(make_tmap_2d(&a.tmB, B, 2, (uint64_t)b_cols, (uint64_t)16, (uint64_t)ldb * 2, 64, 16));
  a.bias = bias;
  a.w0 = 1.0f;
  a.splits = 1;
  gemm_finalize(a);
  return a;
}

// G[Mf,Nf] += dY[P,Mf]^T * X[P,Nf]   (reduction over the P samples; MN-major operands; split-K)
// colsum (optional): [Mf] += column sums of dY over the samples = the bias gradient, computed by the same launch
// side (optional, SM-pair launches only): side_out[Mf, side_n] += dY^T * X2, X2 given transposed as XT[side_n, ldt] (bf16)
struct WgradSide {
  const void* xt;
  long long ldt;
  int n;        // 16 or 64 MMA columns
  float* out;
  long long ld;
  int n_real;   // rows XT really has (<= n; the tensor map zero-fills the rest)
};

static void add_wgrad(Plan& p, int Mf, int Nf, const void* dY, long long ld_dy, const void* X, long long ld_x,
                      long long P, float* G, long long ldg, int sms, float* colsum = nullptr,
                      const WgradSide* side = nullptr) {
  GemmArgs& a = p.add(EPI_WGRAD);
  a.M = Mf;
  a.N = Nf;
  a.block_n = Nf >= 256 ? 256 : Nf;
  a.cta_group = (Mf % 256 == 0) ? gemm_pick_cta_group(EPI_WGRAD, Mf, Nf, a.block_n) : 1;
  a.nseg = 1;
  a.kb_total = (int)((P + 63) / 64);
  a.seg_kb[0] = a.kb_total;
  a.a_mn = 1;
  a.b_mn = 1;
  p.chk(make_tmap_2d(&a.tmA[0], dY, 2, (uint64_t)Mf, (uint64_t)P, (uint64_t)ld_dy * 2, 64, 64));
  p.chk(make_tmap_2d(&a.tmB, X, 2, (uint64_t)Nf, (uint64_t)P, (uint64_t)ld_x * 2, 64, 64));
  if (a.block_n >= 32 && Nf % 32 == 0) {
    p.chk(make_tmap_2d(&a.tmO0, G, 4, (uint64_t)Nf, (uint64_t)Mf, (uint64_t)ldg * 4, 32, GEMM_BLOCK_M));
  } else {
    a.f32out = G;
    a.ldo = ldg;
  }
  const int tiles = ((Mf + 128 * a.cta_group - 1) / (128 * a.cta_group)) * ((Nf + a.block_n - 1) / a.block_n);
  // one wave: the largest split count whose tile total still fits the SMs / SM pairs (one more tile would double the time)
  int splits = (sms / a.cta_group) / tiles;
  if (splits < 1) splits = 1;
  // keep at least 4 k-blocks per split so the accumulate traffic stays small against the operand reads
  const int max_splits = a.kb_total / 4 > 0 ? a.kb_total / 4 : 1;
  a.splits = splits < max_splits ? splits : max_splits;
  a.colsum = colsum;
  if (side != nullptr) {
    if (a.cta_group != 2) {
      set_error("wgrad: a side operand needs the SM-pair form (M %d, N %d)", Mf, Nf);
      p.chk(SNB_ERR_UNSUPPORTED);
    }
    a.side_n = side->n;
    a.side_out = side->out;
    a.ld_side = side->ld;
    p.chk(make_tmap_2d(&a.tmB2, side->xt, 2, (uint64_t)P, (uint64_t)(side->n_real > 0 ? side->n_real : side->n),
                       (uint64_t)side->ldt * 2, 64, (uint32_t)(side->n / 2)));
  }
  gemm_finalize(a);
}

// dst[c, p] = src[p, c] for c < cols (bf16): the K-major copies of the narrow wgrad side operands
__global__ void __launch_bounds__(256) transpose_cols_kernel(const __nv_bfloat16* __restrict__ src, long long ld_src, int cols,
                                                             long long P, __nv_bfloat16* __restrict__ dst, long long ldt) {
  __shared__ __nv_bfloat16 tile[64][64 + 2];
  const long long p0 = (long long)blockIdx.x * 64;
  for (int i = threadIdx.x; i < 64 * cols; i += 256) {
    const int r = i / cols, c = i - r * cols;
    tile[r][c] = p0 + r < P ? src[(p0 + r) * ld_src + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * cols; i += 256) {
    const int c = i >> 6, r = i & 63;
    if (p0 + r < ldt) dst[(long long)c * ldt + p0 + r] = tile[r][c];
  }
}

// the same for cols = 16 / 32 / 64 (every caller): 16-byte loads along the rows, 4-byte stores of sample pairs along the
// samples, 128-sample tiles, grid-stride (the generic kernel moved 2 bytes per access: 167 us of an 8192-ray step, ~4x the
// HBM time of its 300 MB)
template <int COLS>
__global__ void __launch_bounds__(256) transpose_cols_vec_kernel(const __nv_bfloat16* __restrict__ src, long long ld_src, long long P,
                                                                 __nv_bfloat16* __restrict__ dst, long long ldt) {
  constexpr int TP = 128, CPR = COLS / 8;
  __shared__ __nv_bfloat16 tile[TP][COLS + 2];
  for (long long p0 = (long long)blockIdx.x * TP; p0 < ldt; p0 += (long long)gridDim.x * TP) {
    for (int i = threadIdx.x; i < TP * CPR; i += 256) {
      const int r = i / CPR, q = i % CPR;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (p0 + r < P) v = *reinterpret_cast<const uint4*>(src + (p0 + r) * ld_src + q * 8);
      const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
      for (int j = 0; j < 8; ++j) tile[r][q * 8 + j] = e[j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (TP / 2) * COLS; i += 256) {
      const int c = i / (TP / 2), r = (i % (TP / 2)) * 2;
      if (p0 + r < ldt) {   // ldt is a multiple of 64: a pair never straddles it
        __nv_bfloat162 pr;
        pr.x = tile[r][c];
        pr.y = tile[r + 1][c];
        *reinterpret_cast<__nv_bfloat162*>(dst + (long long)c * ldt + p0 + r) = pr;
      }
    }
    __syncthreads();
  }
}

static int transpose_cols(const void* src, long long ld_src, int cols, long long P, void* dst, long long ldt, cudaStream_t st) {
  if ((cols == 16 || cols == 32 || cols == 64) && ld_src % 8 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 3) == 0 &&
      ldt % 64 == 0) {
    long long blocks = (ldt + 127) / 128;
    if (blocks > 148 * 8) blocks = 148 * 8;
    const __nv_bfloat16* s = (const __nv_bfloat16*)src;
    __nv_bfloat16* d = (__nv_bfloat16*)dst;
    if (cols == 64) transpose_cols_vec_kernel<64><<<(unsigned)blocks, 256, 0, st>>>(s, ld_src, P, d, ldt);
    else if (cols == 32) transpose_cols_vec_kernel<32><<<(unsigned)blocks, 256, 0, st>>>(s, ld_src, P, d, ldt);
    else transpose_cols_vec_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(s, ld_src, P, d, ldt);
    return launch_status("transpose_cols_vec_kernel");
  }
  transpose_cols_kernel<<<(unsigned)((ldt + 63) / 64), 256, 0, st>>>((const __nv_bfloat16*)src, ld_src, cols, P,
                                                                     (__nv_bfloat16*)dst, ldt);
  return launch_status("transpose_cols_kernel");
}

static int run_plan(const Plan& p, cudaStream_t st) {
  if (p.rc) return p.rc;
  for (size_t i = 0; i < p.g.size(); ++i)
    if (int r = gemm_launch(p.g[i], p.epi[i], st)) return r;
  return 0;
}

// ---- chained-kernel plan ------------------------------------------------------------------------------
struct CSeg {
  const void* ptr;   // bf16, row-major, first column of the segment
  long long ld;
  int cols;          // readable columns (TMA zero-fills beyond)
  int kb;
  long long rows;    // rows of the tensor (P, or the scratch rows)
  int scratch;
};

struct ChainPlan {
  ChainArgs a;
  int rc = 0;
  bool per_layer;     // one launch per layer (snb_set_chained_mlp(0)) instead of one per pass
  bool relu = false;  // vanilla NeRF: EPI_SIN layers apply max(., 0), EPI_MUL layers with a saved activation multiply by [h > 0]
  cudaStream_t st;
  ChainPlan(long long P, bool chained, cudaStream_t stream) : per_layer(!chained), st(stream) {
    memset(&a, 0, sizeof(a));
    begin_pass(P);
  }
  // the following layers belong to a (second) pass over P rows of their own; both passes run in ONE launch (per-layer mode:
  // nothing to merge, the layers are launched one by one)
  void begin_pass(long long P) {
    if (per_layer || a.n_passes == 0 || a.n_layers == 0) {
      a.n_passes = 1;
      a.n_layers = 0;
    } else {
      close_pass();
      if (a.n_passes >= 2) {
        set_error("chain plan: more than two passes");
        chk(SNB_ERR_UNSUPPORTED);
        return;
      }
      // the pairs that carry one block more of pass 0 than the others carry one less of this one
      const int sms = num_sms();
      const int np = a.pass[0].n_blocks < sms / 2 ? a.pass[0].n_blocks : sms / 2;
      ++a.n_passes;
      memset(&a.pass[1], 0, sizeof(ChainPass));
      a.pass[1].shift = np > 0 ? a.pass[0].n_blocks % np : 0;
    }
    ChainPass& ps = a.pass[a.n_passes - 1];
    ps.layer0 = a.n_layers;
    ps.M = (int)P;
    ps.n_blocks = (int)((P + 255) / 256);
  }
  ChainPass& pass() { return a.pass[a.n_passes - 1]; }
  void close_pass() { pass().n_layers = a.n_layers - pass().layer0; }
  void chk(int r) {
    if (r && !rc) rc = r;
  }
  // D[rows, N] = epilogue(sum_seg A_seg * B[N, Kp]^T); N % 256 == 0.
  //   EPI_SIN:  out0 = sin(w0 (acc + bias)); mask (optional) receives the sign bits of the derivative
  //   EPI_MUL:  out0 = acc * mul, or (mask != NULL) acc * w0 * (+-)sqrt(1 - mul^2): `mul` is then the saved activation
  void add(int epi, int N, const CSeg* segs, int nseg, const void* B, long long ldb, int b_cols, void* out0, long long ldo,
           long long o_rows, int o_scratch, const void* mul, long long ldmul, uint32_t* mask, int mask_ld, const float* bias,
           float w0) {
    if (a.n_layers >= CHAIN_MAX_LAYERS || N % 256 != 0 || nseg > 3) {
      set_error("chain plan: layer %d unsupported (N %d, %d segments)", a.n_layers, N, nseg);
      chk(SNB_ERR_UNSUPPORTED);
      return;
    }
    ChainMaps& mp = a.maps[a.n_layers];
    ChainLayer& ly = a.layers[a.n_layers++];
    ly.epi = epi;
    ly.n_tiles = N / 256;
    ly.nseg = nseg;
    ly.kb_total = 0;
    for (int s = 0; s < nseg; ++s) {
      ly.seg_kb[s] = segs[s].kb;
      ly.a_scratch[s] = per_layer ? 0 : segs[s].scratch;
      ly.kb_total += segs[s].kb;
      chk(make_tmap_2d(&mp.tmA[s], segs[s].ptr, 2, (uint64_t)segs[s].cols, (uint64_t)segs[s].rows, (uint64_t)segs[s].ld * 2, 64,
                       GEMM_BLOCK_M));
    }
    {
      const CSeg& last = segs[nseg - 1];
      ly.tail_k16 = (last.kb == 1 && last.cols < 64) ? (last.cols + 15) / 16 : 4;
    }
    chk(make_tmap_2d(&mp.tmB, B, 2, (uint64_t)b_cols, (uint64_t)N, (uint64_t)ldb * 2, 64, 128));
    chk(make_tmap_2d(&mp.tmO0, out0, 2, (uint64_t)N, (uint64_t)o_rows, (uint64_t)ldo * 2, 64, GEMM_BLOCK_M));
    if (epi == EPI_MUL) chk(make_tmap_2d(&mp.tmMul, mul, 2, (uint64_t)N, (uint64_t)pass().M, (uint64_t)ldmul * 2, 64, GEMM_BLOCK_M));
    ly.mul_siren = (epi == EPI_MUL && mask != nullptr) ? (relu ? 2 : 1) : 0;
    ly.relu = (relu && epi == EPI_SIN) ? 1 : 0;
    if (epi == EPI_LINEAR || relu) mask = nullptr;   // ReLU needs no sign mask: its derivative is [h > 0]
    ly.mask = mask;
    ly.mask_ld = mask_ld;
    if (mask) chk(make_tmap_mask(&mp.tmMask, mask, (uint64_t)(N / 32), (uint64_t)pass().M, (uint64_t)mask_ld * 4));
    ly.o_scratch = o_scratch;
    ly.bias = bias;
    ly.w0 = w0;
    if (per_layer) flush();
  }
  // 16 head pre-activations: part (+)= A_seg * B[16, Kp]^T (mode 0 / 1), or (mode 2) the packed head outputs
  void add_rows16(const CSeg& seg, const void* B, long long ldb, int b_cols, int mode, float* part, const float* bias) {
    if (a.n_layers >= CHAIN_MAX_LAYERS) {
      chk(SNB_ERR_UNSUPPORTED);
      return;
    }
    ChainMaps& mp = a.maps[a.n_layers];
    ChainLayer& ly = a.layers[a.n_layers++];
    ly.epi = EPI_HEADOUT;
    ly.n_tiles = 1;
    ly.nseg = 1;
    ly.kb_total = ly.seg_kb[0] = seg.kb;
    ly.tail_k16 = 4;
    ly.a_scratch[0] = per_layer ? 0 : seg.scratch;
    chk(make_tmap_2d(&mp.tmA[0], seg.ptr, 2, (uint64_t)seg.cols, (uint64_t)seg.rows, (uint64_t)seg.ld * 2, 64, GEMM_BLOCK_M));
    chk(make_tmap_2d(&mp.tmB, B, 2, (uint64_t)b_cols, (uint64_t)16, (uint64_t)ldb * 2, 64, 8));
    ly.rows_mode = mode;
    ly.part = part;
    ly.bias = bias;
    ly.w0 = 1.0f;
    if (per_layer) flush();
  }
  void flush() {
    close_pass();
    if (!rc && a.n_layers > 0) chk(chain_launch(a, st));
    a.n_layers = 0;
    if (a.n_passes > 1) {   // (a flushed two-pass plan is finished)
      a.n_passes = 1;
      a.pass[0] = a.pass[1];
    }
    a.pass[0].layer0 = 0;
    a.pass[0].shift = 0;
  }
  int run() {
    flush();
    return rc;
  }
};

static bool mask_supported(int head_mask) {
  return head_mask == SNB_HEADS_ALL || head_mask == SNB_HEADS_SOLAR || head_mask == SNB_HEADS_DEPTH;
}

}  // namespace snb

using namespace snb;

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" int snb_model_create(snb_model** out, int model_kind, int n_classes, int semantic_sigmoid, int variant,
                                int t_embedding_tau, int mapping_pos_n_freq) {
  SNB_CHECK_ARG(out != nullptr, SNB_ERR_INVALID, "model_create: null out");
  SNB_CHECK_ARG(model_kind >= SNB_MODEL_SATNERF && model_kind <= SNB_MODEL_SNERF, SNB_ERR_INVALID, "model_create: bad kind %d",
                model_kind);
  if (model_kind != SNB_MODEL_SEMANTIC) n_classes = 0;
  SNB_CHECK_ARG((variant & ~(SNB_VARIANT_FULL_FEATURES | SNB_VARIANT_RELU)) == 0 || model_kind == SNB_MODEL_SEMANTIC,
                SNB_ERR_UNSUPPORTED, "model_create: head-input variants exist for the semantic model only");
  SNB_CHECK_ARG(!(variant & (SNB_VARIANT_FULL_FEATURES | SNB_VARIANT_RELU)) || model_kind == SNB_MODEL_SEMANTIC ||
                    model_kind == SNB_MODEL_SATNERF,
                SNB_ERR_UNSUPPORTED,
                "model_create: fc_use_full_features / the ReLU activation exist for SatNeRF and the semantic model (satnerf.py:123-127)");
  SNB_CHECK_ARG((variant & ~(SNB_VARIANT_TJ_FOR_S | SNB_VARIANT_TJ_INSTEAD_OF_BETA | SNB_VARIANT_SEPARATE_BETA_S |
                             SNB_VARIANT_SEPARATE_TJ_S | SNB_VARIANT_FULL_FEATURES | SNB_VARIANT_RELU)) == 0,
                SNB_ERR_UNSUPPORTED, "model_create: variant bits %d not implemented", variant);
  // the per-ray columns of the head inputs travel in the 16 aux columns [1 | sun_d (3) | t (tau) | t_s (tau)]
  if (t_embedding_tau <= 0) t_embedding_tau = 4;
  if (mapping_pos_n_freq <= 0) mapping_pos_n_freq = 10;
  SNB_CHECK_ARG(mapping_pos_n_freq <= 10 && (mapping_pos_n_freq == 10 || model_kind == SNB_MODEL_SEMANTIC), SNB_ERR_UNSUPPORTED,
                "model_create: mapping_pos_n_freq %d (1..10, semantic model: K1 writes the 10-frequency row)", mapping_pos_n_freq);
  SNB_CHECK_ARG(t_embedding_tau <= ((variant & SNB_VARIANT_SEPARATE_TJ_S) ? 6 : 12), SNB_ERR_UNSUPPORTED,
                "model_create: t_embedding_tau %d does not fit the 16 per-ray columns (max 12, 6 with a second embedding)",
                t_embedding_tau);
  SNB_CHECK_ARG(!(variant & SNB_VARIANT_SEPARATE_BETA_S) || n_classes <= 9, SNB_ERR_UNSUPPORTED,
                "model_create: the separate semantic uncertainty head leaves 9 of the 16 head rows for classes (n_classes %d)",
                n_classes);
  SNB_CHECK_ARG(n_classes >= 0 && n_classes <= 10 && (model_kind != SNB_MODEL_SEMANTIC || n_classes >= 1),
                SNB_ERR_UNSUPPORTED, "model_create: n_classes %d outside [1,10]", n_classes);
  snb_model* m = new snb_model();
  m->kind = model_kind;
  m->n_classes = n_classes;
  m->sem_sigmoid = semantic_sigmoid;
  m->variant = variant;
  m->beta_s = (variant & SNB_VARIANT_SEPARATE_BETA_S) ? 1 : 0;
  m->tau = t_embedding_tau;  // 4 in configs/pipelines/*.toml
  m->fl = (variant & SNB_VARIANT_FULL_FEATURES) ? F : F / 2;
  const int FL = m->fl;
  const bool enc60 = model_kind == SNB_MODEL_SEMANTIC || model_kind == SNB_MODEL_NERF;   // positional encoding of xyz (10 frequencies)
  m->k0 = enc60 ? 60 : 3;
  m->kin0 = model_kind == SNB_MODEL_SEMANTIC ? 6 * mapping_pos_n_freq : m->k0;
  m->enc_ld = enc60 ? 128 : 64;
  m->w0_ld = enc60 ? 192 : 64;
  m->relu = (model_kind == SNB_MODEL_NERF || (variant & SNB_VARIANT_RELU)) ? 1 : 0;
  m->aux_ld = model_kind == SNB_MODEL_NERF ? 32 : 16;
  m->kdir = 24;
  m->n_out = 9 + m->beta_s + n_classes;   // [rgb | sigma | sun | sky | beta | (beta_s) | sem]  (rs_semantic.py:291-311)
  // hidden block order of the fused head first layers: [rgb | beta | (sem) | sun]
  m->hh_rgb = 0;
  m->hh_beta = FL;
  m->hh_sem = 2 * FL;
  // NeRF: the rgb block only (no sun / uncertainty heads: their layers are not part of its plans)
  // S-NeRF: [rgb | sun] (no uncertainty block)
  m->hh_bs = 3 * FL;
  m->hhw = model_kind == SNB_MODEL_SEMANTIC ? (4 + m->beta_s) * FL
                                            : (model_kind == SNB_MODEL_NERF ? FL : (model_kind == SNB_MODEL_SNERF ? 2 * FL : 3 * FL));
  m->hh_sun = m->hhw - FL;
  if (model_kind == SNB_MODEL_SNERF || model_kind == SNB_MODEL_NERF) m->hh_beta = m->hh_sem = 0;   // absent blocks: never addressed
  build_layout(m);
  *out = m;
  return 0;
}

extern "C" void snb_model_destroy(snb_model* m) { delete m; }
extern "C" int snb_set_chained_mlp(int on) {
  const int prev = use_chain() ? 1 : 0;
  g_chain = on ? 1 : 0;
  return prev;
}
extern "C" int64_t snb_model_param_count(const snb_model* m) { return m ? m->n_params : -1; }
extern "C" int snb_model_num_tensors(const snb_model* m) { return m ? (int)m->tensors.size() : -1; }
extern "C" int snb_model_tensor_info(const snb_model* m, int i, const char** name, int64_t* offset, int* rows,
                                     int* cols) {
  SNB_CHECK_ARG(m && i >= 0 && i < (int)m->tensors.size(), SNB_ERR_INVALID, "tensor_info: bad index");
  if (name) *name = m->tensors[i].name.c_str();
  if (offset) *offset = m->tensors[i].offset;
  if (rows) *rows = m->tensors[i].rows;
  if (cols) *cols = m->tensors[i].cols;
  return 0;
}
extern "C" size_t snb_model_packed_bytes(const snb_model* m) {
  return m ? (size_t)m->packed_bf16_elems * 2 + (size_t)m->bias_elems * 4 : 0;
}

extern "C" int snb_model_pack(const snb_model* m, const float* params, void* packed, void* stream) {
  SNB_CHECK_ARG(m && params && packed, SNB_ERR_INVALID, "model_pack: null argument");
  SNB_CHECK_ARG(((uintptr_t)packed & 127) == 0, SNB_ERR_INVALID, "model_pack: packed image must be 128-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  SNB_CUDA(cudaMemsetAsync(packed, 0, snb_model_packed_bytes(m), st));
  float* f32 = reinterpret_cast<float*>(reinterpret_cast<char*>(packed) + (size_t)m->packed_bf16_elems * 2);
  if (int r = run_jobs(m->pack_jobs, false, params, packed, f32, st)) return r;
  // the folded feats layer: per hidden block b, W'_b = W_b[:, :F] W_f (forward rows of wh1 + their transpose in tf) and the
  // bias column b'_b = b_b + W_b[:, :F] b_f, composed in fp32 from the flat parameters
  __nv_bfloat16* pk = reinterpret_cast<__nv_bfloat16*>(packed);
  const long long wf = m->find("feats_from_xyz.weight"), bf = m->find("feats_from_xyz.bias");
  const int kh1 = F + 64, ktf = m->hhw + 64;
  {
    // all hidden blocks at once: their f-columns were just copied (fp32) into the contiguous [hhw, F] block wh1_32
    SmallGemm g;
    memset(&g, 0, sizeof(g));
    g.A = f32 + m->wh1_32; g.sam = F; g.sak = 1;
    g.B = params + wf; g.sbk = F; g.sbn = 1;
    g.M = m->hhw; g.N = F; g.K = F;
    g.C = f32 + m->wp32; g.ldc = F;
    g.splits = PACK_SPLITS; g.slice = (long long)m->hhw * F;   // partial products side by side, summed in order below
    if (int r = small_gemm(g, st)) return r;
    wprime_bf16_kernel<<<dim3(F / 32, (m->hhw + 31) / 32), 256, 0, st>>>(f32 + m->wp32, PACK_SPLITS, g.slice, m->hhw, F,
                                                                        pk + m->wh1, kh1, pk + m->tf, ktf);
    if (int r = launch_status("wprime_bf16_kernel")) return r;
  }
  // bias column (aux column 0 = 1): b' = b_h1 + W_h1 b_f, all hidden blocks in one launch (rows of absent blocks are zero)
  return small_gemv(f32 + m->wh1_32, F, 1, params + bf, 1, m->hhw, F, f32 + m->bh1_32, nullptr, 0, pk + m->wh1 + F, kh1, st);
}

extern "C" size_t snb_mlp_workspace_bytes(const snb_model* m, int64_t n_points, int train) {
  if (!m || n_points <= 0) return 0;
  return layout_workspace(m, n_points, train).total;
}

// plan: NULL = build and run this pass's chained launch; else the pass is appended to *plan (the caller runs it)
static int mlp_forward_rows(const snb_model* m, const void* packed, char* ws, const Workspace& w, long long P, const void* enc,
                            const void* aux, const float* sky, int rows_per_ray, int head_mask, int train, float* out,
                            void* stream, ChainPlan* plan = nullptr, bool plan_started = false);

extern "C" int snb_mlp_forward(const snb_model* m, const void* packed, void* workspace, size_t workspace_bytes,
                               int64_t n_points, const void* enc, const void* aux, const float* sky,
                               int rows_per_ray, int head_mask, int train, float* out, void* stream) {
  SNB_CHECK_ARG(m && packed && workspace && enc && out, SNB_ERR_INVALID, "mlp_forward: null argument");
  SNB_CHECK_ARG(n_points > 0 && n_points < (1ll << 31), SNB_ERR_INVALID, "mlp_forward: n_points out of range");
  SNB_CHECK_ARG(mask_supported(head_mask), SNB_ERR_UNSUPPORTED,
                "mlp_forward: head_mask %d (supported: ALL=63, SOLAR=5, DEPTH=1)", head_mask);
  SNB_CHECK_ARG(head_mask == SNB_HEADS_DEPTH || aux != nullptr, SNB_ERR_INVALID, "mlp_forward: aux required");
  SNB_CHECK_ARG(!(m->kind == SNB_MODEL_NERF && head_mask == SNB_HEADS_SOLAR), SNB_ERR_UNSUPPORTED,
                "mlp_forward: NeRF has no sun head - there is no solar-correction pass (baseline/components/rendering.py:103-118)");
  SNB_CHECK_ARG((((uintptr_t)workspace | (uintptr_t)packed | (uintptr_t)enc) & 127) == 0, SNB_ERR_INVALID,
                "mlp_forward: workspace/packed/enc must be 128-byte aligned");
  const Workspace w = layout_workspace(m, n_points, train);
  SNB_CHECK_ARG(workspace_bytes >= w.total, SNB_ERR_WORKSPACE, "mlp_forward: workspace %zu < required %zu",
                workspace_bytes, w.total);
  return mlp_forward_rows(m, packed, reinterpret_cast<char*>(workspace), w, n_points, enc, aux, sky, rows_per_ray, head_mask,
                          train, out, stream);
}

extern "C" int snb_mlp_forward_with_solar(const snb_model* m, const void* packed, void* workspace, size_t workspace_bytes,
                                          int64_t n_points, int64_t n_solar_points, const void* enc, const void* aux,
                                          const float* sky, int rows_per_ray, float* out, void* stream) {
  SNB_CHECK_ARG(m && packed && workspace && enc && aux && out, SNB_ERR_INVALID, "mlp_forward_with_solar: null argument");
  SNB_CHECK_ARG(n_points > 0 && n_solar_points >= 0 && n_solar_points <= n_points && n_points + n_solar_points < (1ll << 31),
                SNB_ERR_INVALID, "mlp_forward_with_solar: point counts out of range");
  SNB_CHECK_ARG(m->kind != SNB_MODEL_NERF || n_solar_points == 0, SNB_ERR_UNSUPPORTED,
                "mlp_forward_with_solar: NeRF has no sun head - there is no solar-correction pass");
  SNB_CHECK_ARG((((uintptr_t)workspace | (uintptr_t)packed | (uintptr_t)enc) & 127) == 0, SNB_ERR_INVALID,
                "mlp_forward_with_solar: workspace/packed/enc must be 128-byte aligned");
  const Workspace w = layout_workspace(m, n_points + n_solar_points, 1);
  SNB_CHECK_ARG(workspace_bytes >= w.total, SNB_ERR_WORKSPACE, "mlp_forward_with_solar: workspace %zu < required %zu",
                workspace_bytes, w.total);
  char* ws = reinterpret_cast<char*>(workspace);
  // both passes in ONE chained launch (ChainPass): every SM pair carries its blocks of the main pass, then of the solar pass
  ChainPlan cp(n_points, use_chain(), (cudaStream_t)stream);
  if (int r = mlp_forward_rows(m, packed, ws, w, n_points, enc, aux, sky, rows_per_ray, SNB_HEADS_ALL, 1, out, stream, &cp, true))
    return r;
  if (n_solar_points > 0) {
    const char* enc_s = reinterpret_cast<const char*>(enc) + (size_t)n_points * m->enc_ld * 2;
    if (int r = mlp_forward_rows(m, packed, ws, shift_rows(m, w, n_points), n_solar_points, enc_s, aux, nullptr, rows_per_ray,
                                 SNB_HEADS_SOLAR, 1, out + (size_t)n_points * m->n_out, stream, &cp, false))
      return r;
  }
  return cp.run();
}

static int mlp_forward_rows(const snb_model* m, const void* packed, char* ws, const Workspace& w, long long P, const void* enc,
                            const void* aux, const float* sky, int rows_per_ray, int head_mask, int train, float* out,
                            void* stream, ChainPlan* plan, bool plan_started) {
  const int FL = m->fl;
  const __nv_bfloat16* pk = reinterpret_cast<const __nv_bfloat16*>(packed);
  const float* pb = reinterpret_cast<const float*>(reinterpret_cast<const char*>(packed) + (size_t)m->packed_bf16_elems * 2);
  auto H = [&](int i) { return (void*)(ws + w.h[i]); };
  const int hhw = m->hhw;
  const bool need_f = head_mask != SNB_HEADS_DEPTH;
  const bool all = head_mask == SNB_HEADS_ALL;
  {
    // trunk + feats + head first layers + sun layers.  Chained (default): one persistent launch, inter-layer
    // activations are read back from L2 and, in inference, live in a per-SM-pair scratch that never reaches HBM.
    const bool chained = use_chain();
    ChainPlan own(P, chained, (cudaStream_t)stream);
    ChainPlan& cp = plan ? *plan : own;
    if (plan && !plan_started) cp.begin_pass(P);
    auto finish = [&]() -> int { return plan ? cp.rc : cp.run(); };
    cp.relu = m->relu != 0;
    cp.a.prefetch = packed;
    cp.a.prefetch_bytes = (unsigned)((size_t)m->packed_bf16_elems * 2);
    cp.a.nerf = m->kind == SNB_MODEL_NERF ? 1 : 0;
    const float w_first = m->relu ? 1.0f : 30.0f;   // Siren(w0 = 30) on the first trunk layer only (satnerf.py:146)
    const long long R = chain_scratch_rows();
    const bool scr = !train && chained;   // inference: layers 0..6 and s2 in the per-pair scratch
    auto hbuf = [&](int i) { return scr && i < 7 ? (void*)(ws + w.scr_h[i & 1]) : H(i); };
    auto hrows = [&](int i) { return scr && i < 7 ? R : P; };
    auto hscr = [&](int i) { return scr && i < 7 ? 1 : 0; };
    auto SG = [&](int i) { return train ? reinterpret_cast<uint32_t*>(ws + w.sg[i]) : (uint32_t*)nullptr; };
    {
      // semantic: the 128-column row is read twice ([hi|lo|0] then its first 64 columns again, against W_lo)
      CSeg s0[2] = {{enc, m->enc_ld, m->enc_ld, m->enc_ld / 64, P, 0}, {enc, m->enc_ld, 64, 1, P, 0}};
      cp.add(EPI_SIN, F, s0, m->k0 == 60 ? 2 : 1, pk + m->wl[0], m->w0_ld, m->w0_ld, hbuf(0), F, hrows(0), hscr(0),
             nullptr, 0, SG(0), F / 32, pb + m->bl[0], w_first);
    }
    for (int i = 1; i < LAYERS; ++i) {
      if (i == 4) {   // skip connection cat(enc, h3) as two K-segments
        CSeg s[2] = {{enc, m->enc_ld, 64, 1, P, 0}, {hbuf(3), F, F, F / 64, hrows(3), hscr(3)}};
        cp.add(EPI_SIN, F, s, 2, pk + m->wl[4], 64 + F, 64 + F, hbuf(4), F, hrows(4), hscr(4), nullptr, 0, SG(4), F / 32,
               pb + m->bl[4], 1.0f);
      } else {
        CSeg s[1] = {{hbuf(i - 1), F, F, F / 64, hrows(i - 1), hscr(i - 1)}};
        cp.add(EPI_SIN, F, s, 1, pk + m->wl[i], F, F, hbuf(i), F, hrows(i), hscr(i), nullptr, 0, SG(i), F / 32, pb + m->bl[i], 1.0f);
      }
    }
    // head outputs: three N = 16 layers ([h7 | s3 | hh] x W_out^T split by K-segment), each right after the layer that
    // produced its input so it is read from L2; the last one applies the head activations and writes `out`
    cp.pass().out_packed = out;
    cp.pass().sky = sky;
    cp.pass().head_mask = head_mask;
    cp.a.n_out = m->n_out;
    cp.a.rows_per_ray = rows_per_ray;
    cp.a.n_classes = m->n_classes;
    cp.a.beta_s = m->beta_s;
    cp.a.sem_sigmoid = m->sem_sigmoid;
    float* hpart = reinterpret_cast<float*>(ws + w.hpart);
    {
      CSeg s7 = {H(7), F, F, F / 64, P, 0};
      cp.add_rows16(s7, pk + m->who, m->kho, F, need_f ? 0 : 2, need_f ? hpart : nullptr, need_f ? nullptr : pb + m->bho);
    }
    if (need_f) {
      void* s2buf = scr ? (void*)(ws + w.scr_s2) : (void*)(ws + w.s2);
      const long long srows = scr ? R : P;
      const int sflag = scr ? 1 : 0;
      // fused head first layers (all blocks, or only the sun block for the solar pass), straight from h7: the linear
      // feats_from_xyz layer is folded into their weights (W' = W_h1 W_f, see the header)
      const int r0 = all ? 0 : m->hh_sun, n = all ? hhw : FL;
      CSeg s1[2] = {{H(7), F, F, F / 64, P, 0}, {aux, m->aux_ld, m->aux_ld, 1, P, 0}};
      cp.add(EPI_SIN, n, s1, 2, pk + m->wh1 + (long long)r0 * (F + 64), F + 64, F + 64, ws + w.hh + (size_t)r0 * 2, hhw, P, 0,
             nullptr, 0, train ? reinterpret_cast<uint32_t*>(ws + w.sghh) + r0 / 32 : nullptr, hhw / 32, nullptr, 1.0f);
      const bool nerf = m->kind == SNB_MODEL_NERF;   // no sun head: the hh rows are the last head-output layer
      if (all) {
        CSeg shh = {ws + w.hh, hhw, hhw, hhw / 64, P, 0};
        cp.add_rows16(shh, pk + m->who + F + FL, m->kho, hhw, nerf ? 2 : 1, hpart, nerf ? pb + m->bho : nullptr);
      }
      if (nerf) return finish();
      CSeg s2[1] = {{ws + w.hh + (size_t)m->hh_sun * 2, hhw, FL, FL / 64, P, 0}};
      cp.add(EPI_SIN, FL, s2, 1, pk + m->ws2, FL, FL, s2buf, FL, srows, sflag, nullptr, 0,
             train ? reinterpret_cast<uint32_t*>(ws + w.sgs2) : nullptr, FL / 32, pb + m->bs2, 1.0f);
      CSeg s3[1] = {{s2buf, FL, FL, FL / 64, srows, sflag}};
      cp.add(EPI_SIN, FL, s3, 1, pk + m->ws4, FL, FL, ws + w.s3, FL, P, 0, nullptr, 0,
             train ? reinterpret_cast<uint32_t*>(ws + w.sgs3) : nullptr, FL / 32, pb + m->bs4, 1.0f);
      CSeg ss3 = {ws + w.s3, FL, FL, FL / 64, P, 0};
      cp.add_rows16(ss3, pk + m->who + F, m->kho, FL, 2, hpart, pb + m->bho);
    }
    return finish();
  }
}

extern "C" int snb_model_grad_buckets(const snb_model* m, int64_t* lo3, int64_t* hi3) {
  SNB_CHECK_ARG(m && lo3 && hi3, SNB_ERR_INVALID, "grad_buckets: null argument");
  for (int b = 0; b < 3; ++b) {
    lo3[b] = m->bucket_lo[b];
    hi3[b] = m->bucket_hi[b];
  }
  return 0;
}

static int mlp_backward_rows(const snb_model* m, const void* packed, char* ws, const Workspace& w, long long P, long long Psc,
                             const void* enc, const void* aux, const float* out, const float* g_out, int head_mask, float* grads,
                             float* g_aux, void* const* bucket_events, cudaStream_t st);

extern "C" int snb_mlp_backward(const snb_model* m, const void* packed, void* workspace, size_t workspace_bytes,
                                int64_t n_points, const void* enc, const void* aux, const float* out,
                                const float* g_out, int head_mask, float* grads, float* g_aux, void* const* bucket_events,
                                void* stream) {
  SNB_CHECK_ARG(m && packed && workspace && enc && aux && out && g_out && grads, SNB_ERR_INVALID,
                "mlp_backward: null argument");
  SNB_CHECK_ARG(n_points > 0 && n_points < (1ll << 31), SNB_ERR_INVALID, "mlp_backward: n_points out of range");
  SNB_CHECK_ARG(mask_supported(head_mask), SNB_ERR_UNSUPPORTED, "mlp_backward: head_mask %d unsupported", head_mask);
  const Workspace w = layout_workspace(m, n_points, 1);
  SNB_CHECK_ARG(workspace_bytes >= w.total, SNB_ERR_WORKSPACE, "mlp_backward: workspace %zu < required %zu",
                workspace_bytes, w.total);
  return mlp_backward_rows(m, packed, reinterpret_cast<char*>(workspace), w, n_points, 0, enc, aux, out, g_out, head_mask, grads,
                           g_aux, bucket_events, (cudaStream_t)stream);
}

extern "C" int snb_mlp_backward_with_solar(const snb_model* m, const void* packed, void* workspace, size_t workspace_bytes,
                                           int64_t n_points, int64_t n_solar_points, const void* enc, const void* aux,
                                           const float* out, const float* g_out, float* grads, float* g_aux,
                                           void* const* bucket_events, void* stream) {
  SNB_CHECK_ARG(m && packed && workspace && enc && aux && out && g_out && grads, SNB_ERR_INVALID,
                "mlp_backward_with_solar: null argument");
  SNB_CHECK_ARG(n_points > 0 && n_solar_points >= 0 && n_solar_points <= n_points && n_points + n_solar_points < (1ll << 31),
                SNB_ERR_INVALID, "mlp_backward_with_solar: point counts out of range");
  SNB_CHECK_ARG(m->kind != SNB_MODEL_NERF || n_solar_points == 0, SNB_ERR_UNSUPPORTED,
                "mlp_backward_with_solar: NeRF has no sun head - there is no solar-correction pass");
  const Workspace w = layout_workspace(m, n_points + n_solar_points, 1);
  SNB_CHECK_ARG(workspace_bytes >= w.total, SNB_ERR_WORKSPACE, "mlp_backward_with_solar: workspace %zu < required %zu",
                workspace_bytes, w.total);
  return mlp_backward_rows(m, packed, reinterpret_cast<char*>(workspace), w, n_points, n_solar_points, enc, aux, out, g_out,
                           SNB_HEADS_ALL, grads, g_aux, bucket_events, (cudaStream_t)stream);
}

// P rows of the pass `head_mask`, followed (Psc > 0: head_mask is ALL) by Psc rows of the solar-correction pass
static int mlp_backward_rows(const snb_model* m, const void* packed, char* ws, const Workspace& w, long long P, long long Psc,
                             const void* enc, const void* aux, const float* out, const float* g_out, int head_mask, float* grads,
                             float* g_aux, void* const* bucket_events, cudaStream_t st) {
  const int FL = m->fl;
  const int sms = num_sms();
  if (sms <= 0) return SNB_ERR_NO_DEVICE;
  const __nv_bfloat16* pk = reinterpret_cast<const __nv_bfloat16*>(packed);
  float* gs = reinterpret_cast<float*>(ws + w.gscratch);
  auto H = [&](int i) { return (void*)(ws + w.h[i]); };
  const int hhw = m->hhw;
  const bool all = head_mask == SNB_HEADS_ALL;
  const bool depth = head_mask == SNB_HEADS_DEPTH;
  const bool nerf = m->kind == SNB_MODEL_NERF;
  void* dpre = ws + w.dpre;
  const Workspace wsc = shift_rows(m, w, P);   // the solar rows' view
  const long long Pt = P + Psc;                // rows of the layers both passes share

  SNB_CUDA(cudaMemsetAsync(gs, 0, (size_t)m->gscratch_elems * 4, st));
  auto head_grad = [&](const Workspace& v, long long n, long long row0, int mask) -> int {
    long long blocks = (n + 255) / 256;
    if (blocks > sms * 8) blocks = sms * 8;
    head_grad_kernel<<<(int)blocks, 256, 0, st>>>(out + row0 * m->n_out, g_out + row0 * m->n_out, n, m->n_out, m->n_classes,
                                                  m->sem_sigmoid, mask, m->beta_s, (__nv_bfloat16*)(ws + v.dpre), gs + m->gbho);
    return launch_status("head_grad_kernel");
  };
  if (int r = head_grad(w, P, 0, head_mask)) return r;
  if (Psc > 0)
    if (int r = head_grad(wsc, Psc, P, SNB_HEADS_SOLAR)) return r;
  auto DY = [&](int i) { return (void*)(ws + w.dy[i]); };
  Plan p;
  const int r0 = all ? 0 : m->hh_sun, nh = all ? hhw : FL;
  const char* dyhh_r0 = ws + w.dyhh + (size_t)r0 * 2;
  // ---- dgrad: the gradient w.r.t. every pre-activation, from the heads back to trunk layer 0 ---------------
  ChainPlan cp(P, use_chain(), st);   // both passes' dgrad chains in one launch
  cp.relu = m->relu != 0;
  cp.a.prefetch = packed;
  cp.a.prefetch_bytes = (unsigned)((size_t)m->packed_bf16_elems * 2);
  auto dgrad = [&](const Workspace& w, long long P, int head_mask) -> int {
    const bool all = head_mask == SNB_HEADS_ALL;
    const bool depth = head_mask == SNB_HEADS_DEPTH;
    const int r0 = all ? 0 : m->hh_sun, nh = all ? hhw : FL;
    const char* dyhh_r0 = ws + w.dyhh + (size_t)r0 * 2;
    void* dpre = ws + w.dpre;
    auto H = [&](int i) { return (void*)(ws + w.h[i]); };
    auto DY = [&](int i) { return (void*)(ws + w.dy[i]); };
    // chained (default): one persistent launch; each dY is read back from L2 by the next step of the same SM pair.
    // The SIREN derivative w0 cos(.) is rebuilt in the epilogue from the saved activation and its sign mask.
    auto SG = [&](int i) { return reinterpret_cast<uint32_t*>(ws + w.sg[i]); };
    uint32_t* sghh = reinterpret_cast<uint32_t*>(ws + w.sghh);
    CSeg cdpre[1] = {{dpre, 16, 16, 1, P, 0}};
    if (!depth) {
      // sun head: s3 <- head output, then back through sun.4, sun.2 into the sun block of hh  (NeRF has no sun head)
      if (!nerf)
        cp.add(EPI_MUL, FL, cdpre, 1, pk + m->tho, 16, 16, ws + w.dys3, FL, P, 0, ws + w.s3, FL,
               reinterpret_cast<uint32_t*>(ws + w.sgs3), FL / 32, nullptr, 1.0f);
      if (all)  // rgb / beta / sem blocks of hh (columns [0, hhw-256); NeRF: its single rgb block)
        cp.add(EPI_MUL, nerf ? hhw : hhw - FL, cdpre, 1, pk + m->tho + (long long)FL * 16, 16, 16, ws + w.dyhh, hhw, P, 0, ws + w.hh,
               hhw, sghh, hhw / 32, nullptr, 1.0f);
      if (!nerf) {
        CSeg c3[1] = {{ws + w.dys3, FL, FL, FL / 64, P, 0}};
        cp.add(EPI_MUL, FL, c3, 1, pk + m->ts4, FL, FL, ws + w.dys2, FL, P, 0, ws + w.s2, FL,
               reinterpret_cast<uint32_t*>(ws + w.sgs2), FL / 32, nullptr, 1.0f);
        CSeg c2[1] = {{ws + w.dys2, FL, FL, FL / 64, P, 0}};
        cp.add(EPI_MUL, FL, c2, 1, pk + m->ts2, FL, FL, ws + w.dyhh + (size_t)m->hh_sun * 2, hhw, P, 0,
               ws + w.hh + (size_t)m->hh_sun * 2, hhw, sghh + m->hh_sun / 32, hhw / 32, nullptr, 1.0f);
      }
      // dY7 = ([dY_hh | dPre16] * [W' ; w_sigma]) * c7   (W' = W_h1 W_f: the folded feats layer; the sun block is the last
      // hidden block, so the solar pass reads the tail [sun block | dPre16] of the same matrix)
      CSeg c7[2] = {{dyhh_r0, hhw, nh, nh / 64, P, 0}, {dpre, 16, 16, 1, P, 0}};
      cp.add(EPI_MUL, F, c7, 2, pk + m->tf + r0, hhw + 64, nh + 64, DY(7), F, P, 0, H(7), F, SG(7), F / 32, nullptr, 1.0f);
    } else {
      cp.add(EPI_MUL, F, cdpre, 1, pk + m->tf + hhw, hhw + 64, 64, DY(7), F, P, 0, H(7), F, SG(7), F / 32, nullptr, 1.0f);
    }
    for (int i = LAYERS - 1; i > 0; --i) {
      // dY_{i-1} = (dY_i W_i) * c_{i-1}
      CSeg c[1] = {{DY(i), F, F, F / 64, P, 0}};
      cp.add(EPI_MUL, F, c, 1, pk + m->tl[i], F, F, DY(i - 1), F, P, 0, H(i - 1), F, SG(i - 1), F / 32, nullptr,
             (i - 1 == 0 && !m->relu) ? 30.0f : 1.0f);
    }
    return cp.rc;
  };
  if (int r = dgrad(w, P, head_mask)) return r;
  if (Psc > 0) {
    cp.begin_pass(Psc);
    if (int r = dgrad(wsc, Psc, SNB_HEADS_SOLAR)) return r;
  }
  if (int r = cp.run()) return r;
  // ---- wgrad: every weight gradient is dY^T x (layer input), split-K over the samples -----------------------
  // (the layers both passes share - trunk, sun layers, head outputs - reduce over all Pt rows in one launch each)
  const long long ldt = (Pt + 63) & ~63ll;
  if (!depth)
    if (int r = transpose_cols(aux, m->aux_ld, m->aux_ld, P, ws + w.auxT, ldt, st)) return r;
  if (int r = transpose_cols(enc, m->enc_ld, 64, Pt, ws + w.encT, ldt, st)) return r;
  // the wgrads run heads first, then trunk layers 7..0; after each of the three gradient buckets (snb_model_grad_buckets)
  // its packed gradients are added into `grads` and its event (if any) is recorded: a data-parallel caller starts that
  // bucket's all-reduce on a side stream while the remaining wgrads still run
  const float* pbf = reinterpret_cast<const float*>(reinterpret_cast<const char*>(packed) + (size_t)m->packed_bf16_elems * 2);
  auto finish_bucket = [&](int b) -> int {
    if (int r = run_plan(p, st)) return r;
    p.g.clear();
    p.epi.clear();
    if (b == 0 && !depth) {
      // folded feats layer: dW_h1 = G1 W_f^T + s b_f^T, dW_f = W_h1^T G1, db_f = W_h1^T s  with G1 = dY_hh^T h7 and
      // s = colsum(dY_hh) (column 0 of the aux gradient block), over the hidden rows [r0, r0 + nh) this pass touched
      const int gald = m->aux_ld > 16 ? 64 : 16;
      const float* G1 = gs + m->gh1 + (long long)r0 * F;
      const float* sv = gs + m->gh1aux + (long long)r0 * gald;
      SmallGemm g;
      memset(&g, 0, sizeof(g));
      g.A = G1; g.sam = F; g.sak = 1;
      g.B = pbf + m->wf32; g.sbk = 1; g.sbn = F;           // B[k, j] = W_f[j, k]
      g.M = nh; g.N = F; g.K = F;
      g.u = sv; g.su = gald; g.v = pbf + m->bfe; g.sv = 1;
      g.C = gs + m->gh1w + (long long)r0 * F; g.ldc = F;
      if (int r = small_gemm(g, st)) return r;
      memset(&g, 0, sizeof(g));
      g.A = pbf + m->wh1_32 + (long long)r0 * F; g.sam = 1; g.sak = F;   // A[j, o] = W_h1[o, j]
      g.B = G1; g.sbk = F; g.sbn = 1;
      g.M = F; g.N = F; g.K = nh;
      g.C = gs + m->gf; g.ldc = F;
      // split reductions (gf and gh1w are zeroed with the gradient scratch)
      if (int r = small_gemm(g, st)) return r;
      if (int r = small_gemv(pbf + m->wh1_32 + (long long)r0 * F, 1, F, sv, gald, F, nh, nullptr, gs + m->gbf, 1, nullptr, 0, st))
        return r;
    }
    if (int r = run_jobs(m->unpack_jobs, true, gs, nullptr, grads, st, m->bucket_lo[b], m->bucket_hi[b])) return r;
    if (bucket_events != nullptr && bucket_events[b] != nullptr) SNB_CUDA(cudaEventRecord((cudaEvent_t)bucket_events[b], st));
    return 0;
  };
  add_wgrad(p, F, 16, H(7), F, dpre, 16, Pt, gs + m->ghot, 16, sms);
  if (!depth && !nerf) add_wgrad(p, FL, 16, ws + w.s3, FL, dpre, 16, Pt, gs + m->ghot + (long long)F * 16, 16, sms);
  if (all) add_wgrad(p, hhw, 16, ws + w.hh, hhw, dpre, 16, P, gs + m->ghot + (long long)(F + FL) * 16, 16, sms);
  if (!depth) {
    if (!nerf) {
      add_wgrad(p, FL, FL, ws + w.dys3, FL, ws + w.s2, FL, Pt, gs + m->gs4, FL, sms, gs + m->gbs4);
      add_wgrad(p, FL, FL, ws + w.dys2, FL, ws + w.hh + (size_t)m->hh_sun * 2, hhw, Pt, gs + m->gs2, FL, sms, gs + m->gbs2);
    }
    // fused head first layers: weight / bias / per-ray-column gradients
    // ... the bias / per-ray-column gradients dY^T x aux ride the same launch as a 16-column side operand
    const int gald = m->aux_ld > 16 ? 64 : 16;
    const WgradSide s_aux = {ws + w.auxT, ldt, gald, gs + m->gh1aux + (long long)r0 * gald, gald, m->aux_ld};
    // G1 = dY_hh^T h7 (against h7, not f: the feats layer is folded; see finish_bucket(0) for dW_h1 / dW_f / db_f)
    add_wgrad(p, nh, F, dyhh_r0, hhw, H(7), F, P, gs + m->gh1 + (long long)r0 * F, F, sms, nullptr, &s_aux);
    if (Psc > 0) {
      // the solar rows reach the fused head layer through its sun block only (aux row i belongs to solar row i as well)
      const WgradSide s_aux_sc = {ws + w.auxT, ldt, gald, gs + m->gh1aux + (long long)m->hh_sun * gald, gald, m->aux_ld};
      add_wgrad(p, FL, F, ws + wsc.dyhh + (size_t)m->hh_sun * 2, hhw, ws + wsc.h[7], F, Psc,
                gs + m->gh1 + (long long)m->hh_sun * F, F, sms, nullptr, &s_aux_sc);
    }
    if (g_aux && m->find("beta_from_xyz.0.weight") < 0) {
      SNB_CUDA(cudaMemsetAsync(g_aux, 0, (size_t)P * 16 * sizeof(float), st));   // no embedding-dependent head in this model
    } else if (all && g_aux) {
      // d aux = dY_beta * W_beta0[:, 512:]  -> embedding gradient (summed per ray by the caller-side kernel)
      // (over the hidden blocks that take t: the uncertainty block, plus the semantic / colour blocks of the head variants)
      Seg sb[1] = {{ws + w.dyhh + (size_t)m->taux_lo * 2, hhw, m->taux_cols, m->taux_cols / 64}};
      GemmArgs& a = add_rows16(p, EPI_F32ROWS, P, sb, 1, pk + m->taux, m->taux_cols, m->taux_cols, nullptr);
      a.f32out = g_aux;
      a.ldo = 16;
    }
  }
  if (int r = finish_bucket(0)) return r;
  for (int i = LAYERS - 1; i >= 0; --i) {
    if (i == 0) {
      add_wgrad(p, F, 64, DY(0), F, enc, m->enc_ld, Pt, gs + m->gl[0], 64, sms, gs + m->gbl[0]);
    } else {
      // skip layer: its encoding block dY4^T x enc[:, :64] is a 64-column side operand of the same launch
      const WgradSide s_enc = {ws + w.encT, ldt, 64, gs + m->gl4e, 64, 64};
      add_wgrad(p, F, F, DY(i), F, H(i - 1), F, Pt, gs + m->gl[i], F, sms, gs + m->gbl[i], i == 4 ? &s_enc : nullptr);
    }
    if (i == 4)
      if (int r = finish_bucket(1)) return r;
  }
  return finish_bucket(2);
}

// =====================================================================================================
// fp32 verification mode (k_fp32.cuh): the same forward on the CUDA cores, fp32 end to end
// =====================================================================================================
#include "k_fp32.cuh"

namespace {
constexpr long long F32_CHUNK = 65536;   // points per pass over the layers (bounds the fp32 activation scratch)
constexpr int F32_ENC_LD = 64;
// per-point scratch floats: enc | hA | hB | f | g1 | g2
static long long f32_row_floats(const snb_model* m) { return F32_ENC_LD + 3 * snb::F + 2 * m->fl; }
}  // namespace

extern "C" size_t snb_mlp_fp32_workspace_bytes(const snb_model* m, int64_t n_points) {
  if (!m || n_points <= 0) return 0;
  const long long rows = n_points < F32_CHUNK ? n_points : F32_CHUNK;
  return (size_t)rows * f32_row_floats(m) * sizeof(float);
}

extern "C" int snb_mlp_forward_fp32(const snb_model* m, const float* params, void* workspace, size_t workspace_bytes,
                                    int64_t n_points, const float* xyz, const float* sun_d, const float* t, const float* sky,
                                    int rows_per_ray, int head_mask, float* out, void* stream) {
  SNB_CHECK_ARG(m && params && workspace && xyz && out, SNB_ERR_INVALID, "mlp_forward_fp32: null argument");
  SNB_CHECK_ARG(n_points > 0 && n_points < (1ll << 31) && rows_per_ray >= 0, SNB_ERR_INVALID, "mlp_forward_fp32: bad sizes");
  SNB_CHECK_ARG(mask_supported(head_mask), SNB_ERR_UNSUPPORTED,
                "mlp_forward_fp32: head_mask %d (supported: ALL=63, SOLAR=5, DEPTH=1)", head_mask);
  const bool all = head_mask == SNB_HEADS_ALL, depth = head_mask == SNB_HEADS_DEPTH;
  const bool nerf = m->kind == SNB_MODEL_NERF;   // sun_d carries the ENCODED view direction (R, 24) there; t / sky unused
  SNB_CHECK_ARG(depth || sun_d != nullptr, SNB_ERR_INVALID, "mlp_forward_fp32: sun_d required");
  SNB_CHECK_ARG(!all || nerf || ((t != nullptr || m->find("beta_from_xyz.0.weight") < 0) && sky != nullptr), SNB_ERR_INVALID,
                "mlp_forward_fp32: t and sky required for all heads");
  SNB_CHECK_ARG(workspace_bytes >= snb_mlp_fp32_workspace_bytes(m, n_points), SNB_ERR_WORKSPACE,
                "mlp_forward_fp32: workspace %zu too small", workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int FL = m->fl;
  const bool sem = m->kind == SNB_MODEL_SEMANTIC;
  const int hid = m->relu ? F32_RELU : F32_SIN;   // `nl` of the model: every hidden activation (satnerf.py:127)
  const int k0 = m->kin0, tau = m->tau, n_out = m->n_out, C = m->n_classes;   // the model's own input width (6 x its frequencies)
  const bool by_ray = rows_per_ray > 1;   // per-ray sun_d / t / sky rows, broadcast over the ray's samples
  const int div = by_ray ? rows_per_ray : 1;
  auto W = [&](const char* name) { return params + m->find((std::string(name) + ".weight").c_str()); };
  auto B = [&](const char* name) { return params + m->find((std::string(name) + ".bias").c_str()); };
  SNB_CUDA(cudaMemsetAsync(out, 0, (size_t)n_points * n_out * sizeof(float), st));   // heads outside head_mask read as 0
  float* base = reinterpret_cast<float*>(workspace);
  for (long long r0 = 0; r0 < n_points; r0 += F32_CHUNK) {
    const int M = (int)(n_points - r0 < F32_CHUNK ? n_points - r0 : F32_CHUNK);
    float* enc = base;
    float* hA = enc + (long long)M * F32_ENC_LD;
    float* hB = hA + (long long)M * F;
    float* f = hB + (long long)M * F;
    float* g1 = f + (long long)M * F;
    float* g2 = g1 + (long long)M * FL;
    float* o = out + r0 * n_out;
    if (int r = f32_posenc_launch(xyz + r0 * 3, M, m->k0 == 60 ? k0 / 6 : 0, enc, F32_ENC_LD, st)) return r;
    auto gemm = [&](F32Seg s0, const F32Seg* s1, const float* w, int ldw, const float* bias, int N, int act, float w0, float* c,
                    long long ldc) {
      F32Gemm g;
      memset(&g, 0, sizeof(g));
      g.seg[0] = s0;
      g.nseg = 1;
      if (s1) {
        g.seg[1] = *s1;
        g.nseg = 2;
      }
      g.w = w; g.ldw = ldw; g.bias = bias; g.M = M; g.N = N; g.w0 = w0; g.act = act; g.c = c; g.ldc = ldc;
      return f32_gemm_launch(g, st);
    };
    auto rows = [&](const float* a, long long lda, int k) { return F32Seg{a, lda, k, 1, 0}; };
    auto per_ray = [&](const float* a, int k) {   // (rays, k) broadcast by row index, or (points, k)
      return by_ray ? F32Seg{a, (long long)k, k, div, r0} : F32Seg{a + r0 * k, (long long)k, k, 1, 0};
    };
    // use_separate_tj_for_semantic: `t` then carries the two embeddings side by side, (R, 2 tau) = [t | t_s]
    const bool sep_ts = (m->variant & SNB_VARIANT_SEPARATE_TJ_S) != 0;
    const int t_ld = sep_ts ? 2 * tau : tau;
    auto per_ray_t = [&](int col) {   // tau columns of the per-ray embedding rows starting at `col`
      return by_ray ? F32Seg{t + col, (long long)t_ld, tau, div, r0} : F32Seg{t + r0 * t_ld + col, (long long)t_ld, tau, 1, 0};
    };
    const F32Seg senc = rows(enc, F32_ENC_LD, k0);
    // trunk (satnerf.py:220-229): layer i reads cur, writes the other buffer; the skip layer reads cat(enc, h)
    float* cur = hA;
    float* nxt = hB;
    for (int i = 0; i < LAYERS; ++i) {
      const std::string nm = "fc_net." + std::to_string(2 * i);
      int rc;
      if (i == 0) {
        rc = gemm(senc, nullptr, W(nm.c_str()), k0, B(nm.c_str()), F, hid, m->relu ? 1.0f : 30.0f, cur, F);
      } else {
        const F32Seg sh = rows(cur, F, F);
        if (i == 4) rc = gemm(senc, &sh, W(nm.c_str()), k0 + F, B(nm.c_str()), F, hid, 1.0f, nxt, F);
        else rc = gemm(sh, nullptr, W(nm.c_str()), F, B(nm.c_str()), F, hid, 1.0f, nxt, F);
        float* tmp = cur; cur = nxt; nxt = tmp;
      }
      if (rc) return rc;
    }
    const F32Seg sh7 = rows(cur, F, F);
    if (int r = gemm(sh7, nullptr, W("sigma_from_xyz.0"), F, B("sigma_from_xyz.0"), 1, F32_SOFTPLUS, 1.0f, o + 3, n_out)) return r;
    if (depth) continue;
    if (int r = gemm(sh7, nullptr, W("feats_from_xyz"), F, B("feats_from_xyz"), F, F32_NONE, 1.0f, f, F)) return r;
    const F32Seg sf = rows(f, F, F);
    if (nerf) {   // rgb = sigmoid(W2 relu(W0 cat(f, Mapping(dir)) + b0) + b2) * 1.002 - 0.001  (nerf.py:197-203); sun column = 1
      const F32Seg sd = per_ray(sun_d, m->kdir);
      if (int r = gemm(sf, &sd, W("rgb_from_xyzdir.0"), F + m->kdir, B("rgb_from_xyzdir.0"), FL, F32_RELU, 1.0f, g1, FL)) return r;
      if (int r = gemm(rows(g1, FL, FL), nullptr, W("rgb_from_xyzdir.2"), FL, B("rgb_from_xyzdir.2"), 3, F32_RGB, 1.0f, o, n_out)) return r;
      continue;
    }
    {   // sun visibility: cat(f, sun_d) -> 3 x sin -> sigmoid  (satnerf.py:236-243)
      const F32Seg ss = per_ray(sun_d, 3);
      if (int r = gemm(sf, &ss, W("sun_v_net.0"), F + 3, B("sun_v_net.0"), FL, hid, 1.0f, g1, FL)) return r;
      if (int r = gemm(rows(g1, FL, FL), nullptr, W("sun_v_net.2"), FL, B("sun_v_net.2"), FL, hid, 1.0f, g2, FL)) return r;
      if (int r = gemm(rows(g2, FL, FL), nullptr, W("sun_v_net.4"), FL, B("sun_v_net.4"), FL, hid, 1.0f, g1, FL)) return r;
      if (int r = gemm(rows(g1, FL, FL), nullptr, W("sun_v_net.6"), FL, B("sun_v_net.6"), 1, F32_SIGMOID, 1.0f, o + 4, n_out)) return r;
    }
    if (!all) continue;
    {
      const bool tj_rgb = (m->variant & SNB_VARIANT_TJ_INSTEAD_OF_BETA) != 0;   // cat(f, t) -> colour head (rs_semantic.py:287-288)
      const F32Seg st_ = per_ray_t(0);
      if (int r = gemm(sf, tj_rgb ? &st_ : nullptr, W("rgb_from_xyzdir.0"), F + (tj_rgb ? tau : 0), B("rgb_from_xyzdir.0"), FL,
                       hid, 1.0f, g1, FL)) return r;
    }
    if (int r = gemm(rows(g1, FL, FL), nullptr, W("rgb_from_xyzdir.2"), FL, B("rgb_from_xyzdir.2"), 3, F32_RGB, 1.0f, o, n_out)) return r;
    if (m->find("beta_from_xyz.0.weight") >= 0) {
      const F32Seg st_ = per_ray_t(0);
      if (int r = gemm(sf, &st_, W("beta_from_xyz.0"), F + tau, B("beta_from_xyz.0"), FL, hid, 1.0f, g1, FL)) return r;
      if (int r = gemm(rows(g1, FL, FL), nullptr, W("beta_from_xyz.2"), FL, B("beta_from_xyz.2"), 1, F32_SOFTPLUS, 1.0f, o + 8, n_out)) return r;
    }
    if (m->beta_s) {   // semantic uncertainty head (rs_semantic.py:228-237,297-303): cat(f, t) -> sin -> softplus, column 9
      const F32Seg st_ = per_ray_t(sep_ts ? tau : 0);
      if (int r = gemm(sf, &st_, W("semantic_beta_from_xyz.0"), F + tau, B("semantic_beta_from_xyz.0"), FL, hid, 1.0f, g1, FL)) return r;
      if (int r = gemm(rows(g1, FL, FL), nullptr, W("semantic_beta_from_xyz.2"), FL, B("semantic_beta_from_xyz.2"), 1, F32_SOFTPLUS, 1.0f,
                       o + 9, n_out)) return r;
    }
    if (sem) {
      const bool tj_s = (m->variant & SNB_VARIANT_TJ_FOR_S) != 0;               // cat(f, t) -> semantic head (rs_semantic.py:330-338)
      const F32Seg st_ = per_ray_t(sep_ts ? tau : 0);
      if (int r = gemm(sf, tj_s ? &st_ : nullptr, W("semantic_prediction.0"), F + (tj_s ? tau : 0), B("semantic_prediction.0"), FL,
                       hid, 1.0f, g1, FL)) return r;
      if (int r = gemm(rows(g1, FL, FL), nullptr, W("semantic_prediction.2"), FL, B("semantic_prediction.2"), C,
                       m->sem_sigmoid ? F32_SIGMOID : F32_NONE, 1.0f, o + 9 + m->beta_s, n_out)) return r;
    }
    if (int r = f32_sky_launch(sky, M, div, r0, o, n_out, st)) return r;
  }
  return 0;
}
