// K1 - stratified sample generation along satellite rays + input encoding.
//
// Follows framework/components/rendering.py:84-116 (sample_rays; arithmetic kept in the
// reference's association with explicit round-to-nearest mul/add so no FMA contraction changes
// a bit), baseline/models/commons.py:58-74 (Mapping.forward: [sin(2^k x), cos(2^k x)]_k, no
// identity term), the per-ray broadcasts of semantic/models/rs_semantic.py:42-61 and the
// embedding lookup of semantic/components/rendering.py:35-45 (int cast done on device).
//
// Output rows are written as bf16 K-segments ready for TMA:
//   enc (P, enc_ld): [hi(k0) | hi(k0) | lo(k0) | 0]  - hi + lo is the two-term bf16 split of the
//        fp32 encoding; against [W_hi | W_lo | W_hi] the first trunk layer keeps ~16 mantissa
//        bits through the tensor pipe even though its SIREN frequency is 30.
//   aux (P, 16):     [1, sun_d(3), t(tau), 0..]       - the per-ray columns of the head inputs
//        cat(f, sun_d) / cat(f, t) (satnerf.py:245,250) and the bias, as one extra K-segment.
// HBM-bound: 48 B read per ray, 4 + 2*enc_ld + 32 B written per sample.
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

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// encode one point into a bf16 row [hi | hi | lo | 0] of enc_ld columns, written with 16-byte stores
template <int KIND>
__device__ __forceinline__ void write_enc_row(__nv_bfloat16* __restrict__ row, float x, float y, float z) {
  constexpr int K0 = (KIND == SNB_MODEL_SEMANTIC) ? 60 : 3;
  constexpr int LD = (KIND == SNB_MODEL_SEMANTIC) ? 192 : 64;
  __align__(16) __nv_bfloat16 buf[LD];
  float xyz[3] = {x, y, z};
  if (KIND == SNB_MODEL_SEMANTIC) {
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const float f = (float)(1 << k);
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float s, c;
        sincosf(f * xyz[ch], &s, &c);  // f is a power of two: the product is exact
        __nv_bfloat16 hi, lo;
        split_bf16(s, hi, lo);
        buf[k * 6 + ch] = hi; buf[K0 + k * 6 + ch] = hi; buf[2 * K0 + k * 6 + ch] = lo;
        split_bf16(c, hi, lo);
        buf[k * 6 + 3 + ch] = hi; buf[K0 + k * 6 + 3 + ch] = hi; buf[2 * K0 + k * 6 + 3 + ch] = lo;
      }
    }
  } else {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      __nv_bfloat16 hi, lo;
      split_bf16(xyz[ch], hi, lo);
      buf[ch] = hi; buf[K0 + ch] = hi; buf[2 * K0 + ch] = lo;
    }
  }
#pragma unroll
  for (int i = 3 * K0; i < LD; ++i) buf[i] = __float2bfloat16_rn(0.f);
  uint4* dst = reinterpret_cast<uint4*>(row);
  const uint4* src = reinterpret_cast<const uint4*>(buf);
#pragma unroll
  for (int i = 0; i < LD / 8; ++i) dst[i] = src[i];
}

__device__ __forceinline__ void write_aux_row(__nv_bfloat16* __restrict__ row, float sx, float sy, float sz,
                                              const float* __restrict__ t, int tau) {
  __align__(16) __nv_bfloat16 buf[16];
  buf[0] = __float2bfloat16_rn(1.0f);
  buf[1] = __float2bfloat16_rn(sx);
  buf[2] = __float2bfloat16_rn(sy);
  buf[3] = __float2bfloat16_rn(sz);
#pragma unroll
  for (int i = 0; i < 12; ++i) buf[4 + i] = __float2bfloat16_rn((t != nullptr && i < tau) ? t[i] : 0.f);
  uint4* dst = reinterpret_cast<uint4*>(row);
  dst[0] = reinterpret_cast<const uint4*>(buf)[0];
  dst[1] = reinterpret_cast<const uint4*>(buf)[1];
}

template <int KIND>
__global__ void __launch_bounds__(128)
k1_sample_encode_kernel(const float* __restrict__ rays, const float* __restrict__ extras,
                        const float* __restrict__ u, uint64_t seed, uint64_t ray_offset,
                        const float* __restrict__ t_steps, const float* __restrict__ t_table, int vocab,
                        int tau, int n_rays, int S, int z_given, float* __restrict__ z_vals,
                        __nv_bfloat16* __restrict__ enc, __nv_bfloat16* __restrict__ enc_sc,
                        __nv_bfloat16* __restrict__ aux) {
  constexpr int LD = (KIND == SNB_MODEL_SEMANTIC) ? 192 : 64;
  const long long P = (long long)n_rays * S;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P;
       p += (long long)gridDim.x * blockDim.x) {
    const int ray = (int)(p / S), s = (int)(p - (long long)ray * S);
    const float* r = rays + (size_t)ray * 8;
    const float ox = __ldg(r + 0), oy = __ldg(r + 1), oz = __ldg(r + 2);
    const float* e = extras + (size_t)ray * 4;
    const float sx = __ldg(e + 0), sy = __ldg(e + 1), sz = __ldg(e + 2);
    float z;
    if (z_given) {
      z = z_vals[p];
    } else {
      // rendering.py:95-110  z = near*(1-t) + far*t ; mid points ; lower + (upper-lower)*u
      const float near = __ldg(r + 6), far = __ldg(r + 7);
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
    if (enc) {
      // rendering.py:113  xyz = o + d*z
      const float dx = __ldg(r + 3), dy = __ldg(r + 4), dz = __ldg(r + 5);
      write_enc_row<KIND>(enc + (size_t)p * LD, __fadd_rn(ox, __fmul_rn(dx, z)),
                          __fadd_rn(oy, __fmul_rn(dy, z)), __fadd_rn(oz, __fmul_rn(dz, z)));
    }
    if (enc_sc) {
      // semantic/components/rendering.py:61-63  solar-correction points o + sun_d*z, same z
      write_enc_row<KIND>(enc_sc + (size_t)p * LD, __fadd_rn(ox, __fmul_rn(sx, z)),
                          __fadd_rn(oy, __fmul_rn(sy, z)), __fadd_rn(oz, __fmul_rn(sz, z)));
    }
    if (aux) {
      // semantic/components/rendering.py:35-42: ts -> integer index -> embedding row (on device, no host sync)
      int ti = (int)__ldg(e + 3);
      ti = min(max(ti, 0), vocab - 1);
      write_aux_row(aux + (size_t)p * 16, sx, sy, sz, t_table ? t_table + (size_t)ti * tau : nullptr, tau);
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(128)
k1_encode_points_kernel(const float* __restrict__ xyz, const float* __restrict__ sun_d,
                        const float* __restrict__ t, int tau, long long P, __nv_bfloat16* __restrict__ enc,
                        __nv_bfloat16* __restrict__ aux) {
  constexpr int LD = (KIND == SNB_MODEL_SEMANTIC) ? 192 : 64;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P;
       p += (long long)gridDim.x * blockDim.x) {
    if (enc) write_enc_row<KIND>(enc + (size_t)p * LD, xyz[p * 3], xyz[p * 3 + 1], xyz[p * 3 + 2]);
    if (aux)
      write_aux_row(aux + (size_t)p * 16, sun_d[p * 3], sun_d[p * 3 + 1], sun_d[p * 3 + 2],
                    t ? t + (size_t)p * tau : nullptr, tau);
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
                                 uint64_t ray_offset, const float* t_steps, const float* t_table, int vocab,
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
  long long blocks = (P + 127) / 128;
  if (blocks > (1 << 20)) blocks = 1 << 20;
  if (model_kind == SNB_MODEL_SEMANTIC)
    k1_sample_encode_kernel<SNB_MODEL_SEMANTIC><<<(int)blocks, 128, 0, st>>>(
        rays, extras, u, seed, ray_offset, t_steps, t_table, vocab, tau, n_rays, n_samples, z_given, z_vals,
        (__nv_bfloat16*)enc, (__nv_bfloat16*)enc_sc, (__nv_bfloat16*)aux);
  else
    k1_sample_encode_kernel<SNB_MODEL_SATNERF><<<(int)blocks, 128, 0, st>>>(
        rays, extras, u, seed, ray_offset, t_steps, t_table, vocab, tau, n_rays, n_samples, z_given, z_vals,
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
  long long blocks = ((long long)n_points + 127) / 128;
  if (model_kind == SNB_MODEL_SEMANTIC)
    k1_encode_points_kernel<SNB_MODEL_SEMANTIC><<<(int)blocks, 128, 0, st>>>(
        xyz, sun_d, t, tau, n_points, (__nv_bfloat16*)enc, (__nv_bfloat16*)aux);
  else
    k1_encode_points_kernel<SNB_MODEL_SATNERF><<<(int)blocks, 128, 0, st>>>(
        xyz, sun_d, t, tau, n_points, (__nv_bfloat16*)enc, (__nv_bfloat16*)aux);
  if (int r = launch_status("k1_encode_points_kernel")) return r;
  return sky_launch(sun_d, 3, n_points, sky_w1, sky_b1, sky_w2, sky_b2, sky_hidden, sky, st);
}
