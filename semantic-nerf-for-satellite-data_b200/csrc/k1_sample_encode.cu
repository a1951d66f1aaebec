// K1 - stratified sample generation along satellite rays + input encoding.
//
// Follows framework/components/rendering.py:84-116 (sample_rays; arithmetic kept in the
// reference's association with explicit round-to-nearest mul/add so no FMA contraction changes
// a bit), baseline/models/commons.py:58-74 (Mapping.forward: [sin(2^k x), cos(2^k x)]_k, no
// identity term), the per-ray broadcasts of semantic/models/rs_semantic.py:42-61 and the
// embedding lookup of semantic/components/rendering.py:35-45 (int cast done on device).
//
// Output rows are written as bf16 K-segments ready for TMA:
//   enc (P, enc_ld): hi + lo is the two-term bf16 split of the fp32 encoding; the first trunk layer
//        multiplies hi*W_hi + hi*W_lo + lo*W_hi, so it keeps ~16 mantissa bits through the tensor
//        pipe even though its SIREN frequency is 30.  Row formats: see EncRow below.
//   aux (P, 16):     [1, sun_d(3), t(tau), 0..]       - the per-ray columns of the head inputs
//        cat(f, sun_d) / cat(f, t) (satnerf.py:245,250) and the bias, as one extra K-segment.
// HBM-bound: 48 B read per ray, 4 + 2*enc_ld + 32 B written per sample (semantic: 292 B, + 256 B for the
// solar-correction row); every global store is a fully coalesced 16-byte-per-lane store.
#include "snb_common.cuh"

namespace snb {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// Philox4x32-10 keyed on the seed, counter = (sample, ray_lo, ray_hi, 0): results do not depend on
// the launch geometry or on how rays are sharded across ranks.
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t ray, uint32_t sample) {
  uint32_t c[4] = {sample, (uint32_t)ray, (uint32_t)(ray >> 32), 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return (float)(c[0] >> 8) * (1.0f / 16777216.0f);  // [0,1)
}

// ---- row formats -----------------------------------------------------------------------------------
// semantic (k0 = 60): 128 columns [hi(60) | lo(60) | 0(8)]; the first trunk layer reads the row as two
//   K-segments - columns 0..127 against [W_hi | W_hi | 0] and columns 0..63 against [W_lo | 0] - which is
//   hi*W_hi + lo*W_hi + hi*W_lo without storing hi twice (256 B per sample instead of 384 B).
// satnerf  (k0 = 3):   64 columns  [hi(3) | hi(3) | lo(3) | 0..] against [W_hi | W_lo | W_hi | 0].
template <int KIND>
struct EncRow {
  static constexpr int K0 = (KIND == SNB_MODEL_SEMANTIC) ? 60 : 3;
  static constexpr int LD = (KIND == SNB_MODEL_SEMANTIC) ? 128 : 64;
  static constexpr int CHUNKS = LD / 8;   // 16-byte chunks per row
};

__device__ __forceinline__ uint32_t bf16_bits(float v) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v)); }
__device__ __forceinline__ void split_bits(float v, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi = (uint32_t)__bfloat16_as_ushort(h);
  lo = bf16_bits(v - __bfloat162float(h));
}

// sin and cos of a (|a| <= ~2^12; here |a| <= 512 * |x|) to ~1 ulp: three-term Cody-Waite reduction by pi/2
// (the split of pi/2 into 24 + 24 + 24 significant bits keeps q * C exact enough for |q| < 2^13) and the
// classic single-precision minimax kernels on [-pi/4, pi/4].  The library sincosf costs ~4x as many
// instructions (it also carries a Payne-Hanek path that these arguments never take) and made K1
// instruction-bound instead of HBM-bound.
__device__ __forceinline__ void sincos_reduced(float a, float& sn, float& cs) {
  const float qf = rintf(a * 0.636619772367581343f);   // a * 2/pi
  const int q = (int)qf;
  float r = fmaf(qf, -1.57079601e+00f, a);
  r = fmaf(qf, -3.13916473e-07f, r);
  r = fmaf(qf, -5.39030253e-15f, r);
  const float s2 = r * r;
  float ps = fmaf(s2, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(ps, s2, -1.6666654611e-1f);
  ps = fmaf(ps * s2, r, r);                             // sin(r)
  float pc = fmaf(s2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(pc, s2, 4.166664568298827e-2f);
  pc = fmaf(pc * s2, s2, fmaf(s2, -0.5f, 1.0f));        // cos(r)
  const float a0 = (q & 1) ? pc : ps, b0 = (q & 1) ? ps : pc;
  sn = (q & 2) ? -a0 : a0;
  cs = ((q + 1) & 2) ? -b0 : b0;
}

// Rows are staged in shared memory (one thread = one sample writes its row with 16-byte stores, the
// 16-byte chunk index XOR-swizzled with the row so a quarter-warp hits all 32 banks) and then copied to
// global memory by the whole block with fully coalesced 16-byte stores.
template <int CHUNKS>
__device__ __forceinline__ uint32_t stage_off(int row, int chunk) {
  return (uint32_t)(row * CHUNKS + ((chunk & ~7) | ((chunk ^ row) & 7))) * 16u;
}

template <int KIND>
__device__ __forceinline__ void stage_enc_row(uint8_t* stage, int row, float x, float y, float z) {
  using R = EncRow<KIND>;
  uint32_t w[R::LD / 2];   // 32-bit words = bf16 column pairs; every index below is a compile-time constant
#pragma unroll
  for (int i = 0; i < R::LD / 2; ++i) w[i] = 0u;
  const float xyz[3] = {x, y, z};
  if (KIND == SNB_MODEL_SEMANTIC) {
    // commons.py:68-74: for k: [sin(2^k x)(3), cos(2^k x)(3)], no identity term.
    // sin / cos are evaluated to ~1 ulp at the frequencies k = 0, 3, 6, 9 and carried to k + 1, k + 2 by the double-angle
    // formulas (sin 2a = 2 sin a cos a, cos 2a = 1 - 2 sin^2 a): two doublings grow the absolute error to <= ~10 ulp = 6e-7,
    // an order of magnitude below the 2^-17 the two-term bf16 split keeps.  This halves the kernel's instruction count: with
    // 30 full sincos evaluations per row it was issue-bound (~1100 instructions per 256-byte row), not HBM-bound.
    float sn[10][3], cs[10][3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
#pragma unroll
      for (int k0 = 0; k0 < 10; k0 += 3) {
        sincos_reduced((float)(1 << k0) * xyz[ch], sn[k0][ch], cs[k0][ch]);  // power-of-two factor: the product is exact
#pragma unroll
        for (int k = k0 + 1; k < k0 + 3 && k < 10; ++k) {
          const float s = sn[k - 1][ch], c = cs[k - 1][ch];
          sn[k][ch] = (s + s) * c;
          cs[k][ch] = fmaf(-(s + s), s, 1.0f);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      uint32_t hi[6], lo[6];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        split_bits(sn[k][ch], hi[ch], lo[ch]);
        split_bits(cs[k][ch], hi[3 + ch], lo[3 + ch]);
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        w[k * 3 + j] = hi[2 * j] | (hi[2 * j + 1] << 16);
        w[30 + k * 3 + j] = lo[2 * j] | (lo[2 * j + 1] << 16);
      }
    }
  } else {
    uint32_t hi[3], lo[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) split_bits(xyz[ch], hi[ch], lo[ch]);
    // columns: hi0 hi1 | hi2 hi0 | hi1 hi2 | lo0 lo1 | lo2 0
    w[0] = hi[0] | (hi[1] << 16);
    w[1] = hi[2] | (hi[0] << 16);
    w[2] = hi[1] | (hi[2] << 16);
    w[3] = lo[0] | (lo[1] << 16);
    w[4] = lo[2];
  }
#pragma unroll
  for (int c = 0; c < R::CHUNKS; ++c)
    *reinterpret_cast<uint4*>(stage + stage_off<R::CHUNKS>(row, c)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}

// block-wide copy of `rows` staged rows to global memory (rows are contiguous there)
template <int CHUNKS>
__device__ __forceinline__ void flush_rows(const uint8_t* stage, __nv_bfloat16* __restrict__ dst, int rows) {
  uint4* out = reinterpret_cast<uint4*>(dst);
  const int total = rows * CHUNKS;
  for (int q = threadIdx.x; q < total; q += blockDim.x) {
    const int row = q / CHUNKS, chunk = q - row * CHUNKS;
    out[q] = *reinterpret_cast<const uint4*>(stage + stage_off<CHUNKS>(row, chunk));
  }
}

__device__ __forceinline__ void flush_aux(const uint8_t* stage, __nv_bfloat16* __restrict__ dst, int rows) {
  uint4* out = reinterpret_cast<uint4*>(dst);
  for (int q = threadIdx.x; q < rows * 2; q += blockDim.x) out[q] = reinterpret_cast<const uint4*>(stage)[q];
}

__device__ __forceinline__ void stage_aux_row(uint8_t* stage, int row, float sx, float sy, float sz,
                                              const float* __restrict__ t, int tau) {
  float v[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) v[i] = (t != nullptr && i < tau) ? t[i] : 0.f;
  const uint32_t one = bf16_bits(1.0f);
  uint4* dst = reinterpret_cast<uint4*>(stage + (size_t)row * 32);
  dst[0] = make_uint4(one | (bf16_bits(sx) << 16), bf16_bits(sy) | (bf16_bits(sz) << 16),
                      bf16_bits(v[0]) | (bf16_bits(v[1]) << 16), bf16_bits(v[2]) | (bf16_bits(v[3]) << 16));
  dst[1] = make_uint4(bf16_bits(v[4]) | (bf16_bits(v[5]) << 16), bf16_bits(v[6]) | (bf16_bits(v[7]) << 16),
                      bf16_bits(v[8]) | (bf16_bits(v[9]) << 16), bf16_bits(v[10]) | (bf16_bits(v[11]) << 16));
}

constexpr int K1_THREADS = 128;   // samples per block iteration

template <int KIND>
__global__ void __launch_bounds__(K1_THREADS, 6)   // <= 80 registers: six 36 KB blocks (24 warps) per SM
k1_sample_encode_kernel(const float* __restrict__ rays, const float* __restrict__ extras,
                        const float* __restrict__ u, uint64_t seed, const uint64_t* __restrict__ seed_dev,
                        uint64_t ray_offset, const float* __restrict__ t_steps, const float* __restrict__ t_table, int vocab,
                        int tau, int n_rays, int S, int z_given, float* __restrict__ z_vals,
                        __nv_bfloat16* __restrict__ enc, __nv_bfloat16* __restrict__ enc_sc,
                        __nv_bfloat16* __restrict__ aux) {
  using R = EncRow<KIND>;
  __shared__ __align__(16) uint8_t stage[K1_THREADS * R::LD * 2];
  __shared__ __align__(16) uint8_t stage_aux[K1_THREADS * 32];
  const long long P = (long long)n_rays * S;
  const int tid = threadIdx.x;
  if (seed_dev != nullptr) seed = *seed_dev;   // a captured CUDA graph replays with a fresh key each step
  for (long long p0 = (long long)blockIdx.x * K1_THREADS; p0 < P; p0 += (long long)gridDim.x * K1_THREADS) {
    const long long p = p0 + tid;
    const int rows = (int)min((long long)K1_THREADS, P - p0);
    const bool live = p < P;
    float ox = 0.f, oy = 0.f, oz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f, sx = 0.f, sy = 0.f, sz = 0.f, z = 0.f;
    if (live) {
      const int ray = (int)(p / S), s = (int)(p - (long long)ray * S);
      const float4 r0 = __ldg(reinterpret_cast<const float4*>(rays + (size_t)ray * 8));
      const float4 r1 = __ldg(reinterpret_cast<const float4*>(rays + (size_t)ray * 8) + 1);
      const float4 ex = __ldg(reinterpret_cast<const float4*>(extras + (size_t)ray * 4));
      ox = r0.x; oy = r0.y; oz = r0.z; dx = r0.w; dy = r1.x; dz = r1.y;
      sx = ex.x; sy = ex.y; sz = ex.z;
      if (z_given) {
        z = z_vals[p];
      } else {
        // rendering.py:95-110  z = near*(1-t) + far*t ; mid points ; lower + (upper-lower)*u
        const float near = r1.z, far = r1.w;
        auto z0 = [&](int i) {
          float t = __ldg(t_steps + i);
          return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, t)), __fmul_rn(far, t));
        };
        const float zc = z0(s);
        const float lower = (s == 0) ? zc : __fmul_rn(0.5f, __fadd_rn(z0(s - 1), zc));
        const float upper = (s == S - 1) ? zc : __fmul_rn(0.5f, __fadd_rn(zc, z0(s + 1)));
        const float uu = u ? __ldg(u + p) : philox_uniform(seed, ray_offset + (uint64_t)ray, (uint32_t)s);
        z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), uu));
        z_vals[p] = z;
      }
      if (aux) {
        // semantic/components/rendering.py:35-42: ts -> integer index -> embedding row (on device, no host sync)
        int ti = (int)ex.w;
        ti = min(max(ti, 0), vocab - 1);
        stage_aux_row(stage_aux, tid, sx, sy, sz, t_table ? t_table + (size_t)ti * tau : nullptr, tau);
      }
    }
    if (enc) {
      // rendering.py:113  xyz = o + d*z
      if (live)
        stage_enc_row<KIND>(stage, tid, __fadd_rn(ox, __fmul_rn(dx, z)), __fadd_rn(oy, __fmul_rn(dy, z)),
                            __fadd_rn(oz, __fmul_rn(dz, z)));
      __syncthreads();
      flush_rows<R::CHUNKS>(stage, enc + (size_t)p0 * R::LD, rows);
      if (aux) flush_aux(stage_aux, aux + (size_t)p0 * 16, rows);
      __syncthreads();
    } else if (aux) {
      __syncthreads();
      flush_aux(stage_aux, aux + (size_t)p0 * 16, rows);
      __syncthreads();
    }
    if (enc_sc) {
      // semantic/components/rendering.py:61-63  solar-correction points o + sun_d*z, same z
      if (live)
        stage_enc_row<KIND>(stage, tid, __fadd_rn(ox, __fmul_rn(sx, z)), __fadd_rn(oy, __fmul_rn(sy, z)),
                            __fadd_rn(oz, __fmul_rn(sz, z)));
      __syncthreads();
      flush_rows<R::CHUNKS>(stage, enc_sc + (size_t)p0 * R::LD, rows);
      __syncthreads();
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(K1_THREADS, 6)   // <= 80 registers: six 36 KB blocks (24 warps) per SM
k1_encode_points_kernel(const float* __restrict__ xyz, const float* __restrict__ sun_d,
                        const float* __restrict__ t, int tau, long long P, __nv_bfloat16* __restrict__ enc,
                        __nv_bfloat16* __restrict__ aux) {
  using R = EncRow<KIND>;
  __shared__ __align__(16) uint8_t stage[K1_THREADS * R::LD * 2];
  __shared__ __align__(16) uint8_t stage_aux[K1_THREADS * 32];
  const int tid = threadIdx.x;
  for (long long p0 = (long long)blockIdx.x * K1_THREADS; p0 < P; p0 += (long long)gridDim.x * K1_THREADS) {
    const long long p = p0 + tid;
    const int rows = (int)min((long long)K1_THREADS, P - p0);
    if (p < P) {
      if (enc) stage_enc_row<KIND>(stage, tid, xyz[p * 3], xyz[p * 3 + 1], xyz[p * 3 + 2]);
      if (aux) stage_aux_row(stage_aux, tid, sun_d[p * 3], sun_d[p * 3 + 1], sun_d[p * 3 + 2], t ? t + (size_t)p * tau : nullptr, tau);
    }
    __syncthreads();
    if (enc) flush_rows<R::CHUNKS>(stage, enc + (size_t)p0 * R::LD, rows);
    if (aux) flush_aux(stage_aux, aux + (size_t)p0 * 16, rows);
    __syncthreads();
  }
}

// sky_color(sun_d) = sigmoid(W2 relu(W1 sun_d + b1) + b2)  (satnerf.py:188-193,248): a function of
// the ray only.  One warp per row of `dirs` (stride floats between rows).
__global__ void __launch_bounds__(128)
k1_sky_kernel(const float* __restrict__ dirs, int stride, long long n, const float* __restrict__ w1,
              const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
              int hidden, float* __restrict__ sky) {
  const int lane = threadIdx.x & 31;
  for (long long i = (long long)blockIdx.x * 4 + (threadIdx.x >> 5); i < n; i += (long long)gridDim.x * 4) {
    const float sx = dirs[i * stride], sy = dirs[i * stride + 1], sz = dirs[i * stride + 2];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int h = lane; h < hidden; h += 32) {
      float y = fmaf(w1[h * 3 + 2], sz, fmaf(w1[h * 3 + 1], sy, fmaf(w1[h * 3], sx, b1[h])));
      y = fmaxf(y, 0.f);
      a0 = fmaf(w2[h], y, a0);
      a1 = fmaf(w2[hidden + h], y, a1);
      a2 = fmaf(w2[2 * hidden + h], y, a2);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, d);
      a1 += __shfl_xor_sync(0xffffffffu, a1, d);
      a2 += __shfl_xor_sync(0xffffffffu, a2, d);
    }
    if (lane == 0) {
      sky[i * 3 + 0] = 1.0f / (1.0f + expf(-(a0 + b2[0])));
      sky[i * 3 + 1] = 1.0f / (1.0f + expf(-(a1 + b2[1])));
      sky[i * 3 + 2] = 1.0f / (1.0f + expf(-(a2 + b2[2])));
    }
  }
}

static int sky_launch(const float* dirs, int stride, long long n, const float* w1, const float* b1,
                      const float* w2, const float* b2, int hidden, float* sky, cudaStream_t st) {
  if (!sky || n == 0) return 0;
  SNB_CHECK_ARG(w1 && b1 && w2 && b2 && hidden > 0, SNB_ERR_INVALID, "sky: null sky_color parameters");
  long long blocks = (n + 3) / 4;
  if (blocks > 148 * 32) blocks = 148 * 32;
  k1_sky_kernel<<<(int)blocks, 128, 0, st>>>(dirs, stride, n, w1, b1, w2, b2, hidden, sky);
  return launch_status("k1_sky_kernel");
}

}  // namespace snb

extern "C" int snb_sample_encode(const float* rays, const float* extras, const float* u, uint64_t seed,
                                 const uint64_t* seed_dev, uint64_t ray_offset, const float* t_steps, const float* t_table, int vocab,
                                 int tau, const float* sky_w1, const float* sky_b1, const float* sky_w2,
                                 const float* sky_b2, int sky_hidden, int n_rays, int n_samples,
                                 int model_kind, int z_given, float* z_vals, void* enc, void* enc_sc,
                                 void* aux, float* sky, void* stream) {
  using namespace snb;
  SNB_CHECK_ARG(rays && extras && z_vals, SNB_ERR_INVALID, "sample_encode: null rays/extras/z_vals");
  SNB_CHECK_ARG(z_given || t_steps, SNB_ERR_INVALID, "sample_encode: t_steps (linspace(0,1,S)) required");
  SNB_CHECK_ARG(n_rays >= 0 && n_samples >= 2, SNB_ERR_UNSUPPORTED, "sample_encode: need n_samples >= 2");
  SNB_CHECK_ARG(model_kind == SNB_MODEL_SATNERF || model_kind == SNB_MODEL_SEMANTIC, SNB_ERR_INVALID,
                "sample_encode: bad model_kind %d", model_kind);
  SNB_CHECK_ARG(tau >= 0 && tau <= 12 && (aux == nullptr || t_table == nullptr || vocab > 0), SNB_ERR_UNSUPPORTED,
                "sample_encode: tau %d unsupported (max 12)", tau);
  SNB_CHECK_ARG((((uintptr_t)enc | (uintptr_t)enc_sc | (uintptr_t)aux) & 15) == 0, SNB_ERR_INVALID,
                "sample_encode: enc/aux must be 16-byte aligned");
  if (n_rays == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const long long P = (long long)n_rays * n_samples;
  long long blocks = (P + K1_THREADS - 1) / K1_THREADS;
  if (blocks > 148 * 64) blocks = 148 * 64;
  if (model_kind == SNB_MODEL_SEMANTIC)
    k1_sample_encode_kernel<SNB_MODEL_SEMANTIC><<<(int)blocks, K1_THREADS, 0, st>>>(
        rays, extras, u, seed, seed_dev, ray_offset, t_steps, t_table, vocab, tau, n_rays, n_samples, z_given, z_vals,
        (__nv_bfloat16*)enc, (__nv_bfloat16*)enc_sc, (__nv_bfloat16*)aux);
  else
    k1_sample_encode_kernel<SNB_MODEL_SATNERF><<<(int)blocks, K1_THREADS, 0, st>>>(
        rays, extras, u, seed, seed_dev, ray_offset, t_steps, t_table, vocab, tau, n_rays, n_samples, z_given, z_vals,
        (__nv_bfloat16*)enc, (__nv_bfloat16*)enc_sc, (__nv_bfloat16*)aux);
  if (int r = launch_status("k1_sample_encode_kernel")) return r;
  return sky_launch(extras, 4, n_rays, sky_w1, sky_b1, sky_w2, sky_b2, sky_hidden, sky, st);
}

extern "C" int snb_encode_points(const float* xyz, const float* sun_d, const float* t, int tau,
                                 const float* sky_w1, const float* sky_b1, const float* sky_w2,
                                 const float* sky_b2, int sky_hidden, int n_points, int model_kind, void* enc,
                                 void* aux, float* sky, void* stream) {
  using namespace snb;
  SNB_CHECK_ARG(xyz && sun_d, SNB_ERR_INVALID, "encode_points: null xyz/sun_d");
  SNB_CHECK_ARG(model_kind == SNB_MODEL_SATNERF || model_kind == SNB_MODEL_SEMANTIC, SNB_ERR_INVALID,
                "encode_points: bad model_kind %d", model_kind);
  SNB_CHECK_ARG(tau >= 0 && tau <= 12, SNB_ERR_UNSUPPORTED, "encode_points: tau %d unsupported", tau);
  if (n_points <= 0) return n_points < 0 ? SNB_ERR_INVALID : 0;
  cudaStream_t st = (cudaStream_t)stream;
  long long blocks = ((long long)n_points + K1_THREADS - 1) / K1_THREADS;
  if (blocks > 148 * 64) blocks = 148 * 64;
  if (model_kind == SNB_MODEL_SEMANTIC)
    k1_encode_points_kernel<SNB_MODEL_SEMANTIC><<<(int)blocks, K1_THREADS, 0, st>>>(
        xyz, sun_d, t, tau, n_points, (__nv_bfloat16*)enc, (__nv_bfloat16*)aux);
  else
    k1_encode_points_kernel<SNB_MODEL_SATNERF><<<(int)blocks, K1_THREADS, 0, st>>>(
        xyz, sun_d, t, tau, n_points, (__nv_bfloat16*)enc, (__nv_bfloat16*)aux);
  if (int r = launch_status("k1_encode_points_kernel")) return r;
  return sky_launch(sun_d, 3, n_points, sky_w1, sky_b1, sky_w2, sky_b2, sky_hidden, sky, st);
}
