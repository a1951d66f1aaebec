// K3 - per-ray alpha compositing with the irradiance lighting model and semantic compositing,
// forward and backward.  One warp per ray; the transmittance cumprod is a warp product-scan
// (lane-local serial products over SPL consecutive samples + a 5-step shuffle scan), the
// cumprod gradient is the matching suffix-sum scan.
//
// Follows framework/util/rendering.py:4-34 (convert_sigmas) and the tail of `inference`
// (baseline/models/satnerf.py:73-96, semantic/models/rs_semantic.py:81-126,131-136).
//
// HBM-bound: forward reads (n_out + 1) floats per sample and writes 2 (weights, transparency);
// each ray's packed rows stream into shared memory with 16-byte cp.async copies, double-buffered so the
// next ray's loads are in flight while this one is composited.
#include "snb_common.cuh"

namespace snb {

constexpr int K3_WARPS = 4;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
  return v;
}

// exclusive product scan across lanes
__device__ __forceinline__ float warp_excl_prod(float v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float o = __shfl_up_sync(FULL, v, d);
    if (lane >= d) v *= o;
  }
  float e = __shfl_up_sync(FULL, v, 1);
  return lane == 0 ? 1.0f : e;
}

// exclusive suffix sum across lanes: sum over lanes > lane
__device__ __forceinline__ float warp_excl_suffix_sum(float v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float o = __shfl_down_sync(FULL, v, d);
    if (lane + d < 32) v += o;
  }
  float e = __shfl_down_sync(FULL, v, 1);
  return lane == 31 ? 0.0f : e;
}

struct SampleVals {
  float alpha, e, q, delta, sigma;
};

__device__ __forceinline__ SampleVals sample_alpha(const float* zs, const float* row, int s, int S) {
  SampleVals v;
  v.sigma = row[3];
  // framework/util/rendering.py:12-16: last delta is 1e10
  v.delta = (s < S - 1) ? __fsub_rn(zs[s + 1], zs[s]) : 1e10f;
  float x = __fmul_rn(v.delta, fmaxf(v.sigma, 0.0f));
  v.e = expf(-x);
  v.alpha = __fsub_rn(1.0f, v.e);                       // :24
  v.q = __fadd_rn(__fsub_rn(1.0f, v.alpha), 1e-10f);    // :25-27
  return v;
}

// CMAX: compile-time bound of the class loops (0: no semantic columns - SatNeRF / S-NeRF / NeRF and the solar pass; 6: the
// usual 5-6 classes; 10: the widest head).  With a single 10-wide predicated loop the kernels spent a third of their
// instructions on classes that do not exist (ncu: 600 - 1400 warp instructions per ray, issue-bound at 55 % of the HBM peak).
template <int SPL, bool BWD, int CMAX>
__global__ void __launch_bounds__(K3_WARPS * 32, SPL >= 8 ? 1 : (BWD ? 5 : 6))
k3_composite_kernel(const float* __restrict__ out, const float* __restrict__ z_vals, int n_rays, int S,
                    int n_out, int C, int flags,
                    // forward outputs
                    float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ weights,
                    float* __restrict__ transparency, float* __restrict__ sem_logits,
                    long long* __restrict__ sem_label,
                    // backward inputs / output
                    const float* __restrict__ g_rgb, const float* __restrict__ g_depth,
                    const float* __restrict__ g_weights, const float* __restrict__ g_transp,
                    const float* __restrict__ g_sem, const float* __restrict__ g_direct,
                    float* __restrict__ g_out) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_words = S * n_out;
  const int so = 9 + ((flags & SNB_COMPOSITE_BETA_S) ? 1 : 0);   // first class column (column 9 may hold the semantic uncertainty)
  // per warp: two input buffers {packed rows [S*n_out], z [S]} (the next ray streams in with cp.async while this
  // one is composited) + gradient rows in BWD; every region 16-byte aligned
  const int rw4 = (row_words + 3) & ~3, s4 = (S + 3) & ~3;
  const int in_words = rw4 + s4;
  const int per_warp = 2 * in_words + (BWD ? rw4 : 0);
  float* wbase = smem + warp * per_warp;
  float* grow = wbase + 2 * in_words;  // BWD only
  // 16-byte accesses when every ray's rows start on a 16-byte boundary (S*n_out % 4 == 0: true for the even sample
  // counts the pipelines use)
  const bool vec_rows = (row_words & 3) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                        (!BWD || (((reinterpret_cast<uintptr_t>(g_out) | reinterpret_cast<uintptr_t>(g_direct)) & 15) == 0));
  const bool vec_z = (S & 3) == 0 && ((reinterpret_cast<uintptr_t>(z_vals) & 15) == 0);
  const bool async_in = vec_rows && vec_z;

  // asynchronous global -> shared copy of one ray's inputs (one cp.async group per ray)
  auto fetch = [&](int ray, float* dst) {
    if (ray < n_rays) {
      const float4* s4p = reinterpret_cast<const float4*>(out + (size_t)ray * row_words);
      const uint32_t d = smem_u32(dst);
      for (int i = lane; i < row_words / 4; i += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + i * 16), "l"(s4p + i) : "memory");
      const float4* z4 = reinterpret_cast<const float4*>(z_vals + (size_t)ray * S);
      for (int i = lane; i < S / 4; i += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + (rw4 + i * 4) * 4), "l"(z4 + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int ray0 = blockIdx.x * K3_WARPS + warp, stride = gridDim.x * K3_WARPS;
  int cur = 0;
  if (async_in) fetch(ray0, wbase);
  for (int ray = ray0; ray < n_rays; ray += stride, cur ^= 1) {
    float* rows = wbase + (async_in ? cur * in_words : 0);
    float* zs = rows + rw4;
    if (async_in) {
      fetch(ray + stride, wbase + (cur ^ 1) * in_words);          // the other buffer was released by the __syncwarp below
      asm volatile("cp.async.wait_group 1;" ::: "memory");      // this ray's group has landed
    } else {
      const float* src = out + (size_t)ray * row_words;
      for (int i = lane; i < row_words; i += 32) rows[i] = __ldg(src + i);
      for (int i = lane; i < S; i += 32) zs[i] = __ldg(z_vals + (size_t)ray * S + i);
    }
    __syncwarp();

    // ---- forward pass over this lane's SPL consecutive samples --------------------------------
    float alpha[SPL], tloc[SPL], qv[SPL];
    float Q = 1.0f;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      int s = lane * SPL + j;
      tloc[j] = Q;
      if (s < S) {
        SampleVals v = sample_alpha(zs, rows + s * n_out, s, S);
        alpha[j] = v.alpha;
        qv[j] = v.q;
        Q *= v.q;
      } else {
        alpha[j] = 0.0f;
        qv[j] = 1.0f;
      }
    }
    const float prefix = warp_excl_prod(Q, lane);

    float acc_d = 0.f, acc_r = 0.f, acc_g = 0.f, acc_b = 0.f;
    float acc_s[CMAX > 0 ? CMAX : 1];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc_s[c] = 0.f;
    float T[SPL], w[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      int s = lane * SPL + j;
      T[j] = prefix * tloc[j];
      w[j] = alpha[j] * T[j];
      if (s < S) {
        const float* r = rows + s * n_out;
        float v = r[4];
        acc_d += w[j] * zs[s];
        // rs_semantic.py:101-102: irradiance = sun + (1 - sun) * sky ; rgb = sum w * albedo * irradiance
        acc_r += w[j] * r[0] * (v + (1.0f - v) * r[5]);
        acc_g += w[j] * r[1] * (v + (1.0f - v) * r[6]);
        acc_b += w[j] * r[2] * (v + (1.0f - v) * r[7]);
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) acc_s[c] += w[j] * r[so + c];
      }
    }
    if (!BWD) {
      // this lane's SPL consecutive samples: one 8 / 16-byte store per tensor when the row allows it
      const size_t o = (size_t)ray * S + (size_t)lane * SPL;
      if (SPL == 2 && (S & 1) == 0 && lane * SPL + 1 < S) {
        *reinterpret_cast<float2*>(weights + o) = make_float2(w[0], w[1 % SPL]);
        *reinterpret_cast<float2*>(transparency + o) = make_float2(T[0], T[1 % SPL]);
      } else if (SPL == 4 && (S & 3) == 0 && lane * SPL + 3 < S) {
        *reinterpret_cast<float4*>(weights + o) = make_float4(w[0], w[1 % SPL], w[2 % SPL], w[3 % SPL]);
        *reinterpret_cast<float4*>(transparency + o) = make_float4(T[0], T[1 % SPL], T[2 % SPL], T[3 % SPL]);
      } else {
#pragma unroll
        for (int j = 0; j < SPL; ++j)
          if (lane * SPL + j < S) {
            weights[o + j] = w[j];
            transparency[o + j] = T[j];
          }
      }
    }
    acc_d = warp_sum(acc_d);
    acc_r = warp_sum(acc_r);
    acc_g = warp_sum(acc_g);
    acc_b = warp_sum(acc_b);
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) acc_s[c] = warp_sum(acc_s[c]);

    if (!BWD) {
      if (lane == 0) {
        // rs_semantic.py:103 / satnerf.py:79 clamp the composited colour; NeRF's inference (nerf.py:73-86) does not
        const bool clamp = !(flags & SNB_COMPOSITE_NO_CLAMP);
        rgb[ray * 3 + 0] = clamp ? fminf(fmaxf(acc_r, 0.f), 1.f) : acc_r;
        rgb[ray * 3 + 1] = clamp ? fminf(fmaxf(acc_g, 0.f), 1.f) : acc_g;
        rgb[ray * 3 + 2] = clamp ? fminf(fmaxf(acc_b, 0.f), 1.f) : acc_b;
        depth[ray] = acc_d;
        if (C > 0) {
          int best = 0;
          float bv = acc_s[0];
#pragma unroll
          for (int c = 0; c < CMAX; ++c) {
            if (c < C) {
              sem_logits[(size_t)ray * C + c] = acc_s[c];
              if (acc_s[c] > bv) { bv = acc_s[c]; best = c; }
            }
          }
          sem_label[ray] = best;  // argmax(softmax(x)) == argmax(x), first maximum (rs_semantic.py:131-136)
        }
      }
    } else {
      // ---- backward -------------------------------------------------------------------------
      // clamp passes gradient where 0 <= raw <= 1
      float gr = g_rgb ? g_rgb[ray * 3 + 0] : 0.f, gg = g_rgb ? g_rgb[ray * 3 + 1] : 0.f,
            gb = g_rgb ? g_rgb[ray * 3 + 2] : 0.f;
      if (!(flags & SNB_COMPOSITE_NO_CLAMP)) {
        if (!(acc_r >= 0.f && acc_r <= 1.f)) gr = 0.f;
        if (!(acc_g >= 0.f && acc_g <= 1.f)) gg = 0.f;
        if (!(acc_b >= 0.f && acc_b <= 1.f)) gb = 0.f;
      }
      const float gd = g_depth ? g_depth[ray] : 0.f;
      float gs[CMAX > 0 ? CMAX : 1];
#pragma unroll
      for (int c = 0; c < CMAX; ++c) gs[c] = (g_sem && c < C) ? g_sem[(size_t)ray * C + c] : 0.f;

      float Gw[SPL], B[SPL];
      float Bsum = 0.f;
#pragma unroll
      for (int j = 0; j < SPL; ++j) {
        int s = lane * SPL + j;
        Gw[j] = 0.f;
        B[j] = 0.f;
        if (s < S) {
          const float* r = rows + s * n_out;
          float v = r[4];
          float g = g_weights ? g_weights[(size_t)ray * S + s] : 0.f;
          g += gd * zs[s];
          g += gr * r[0] * (v + (1.0f - v) * r[5]) + gg * r[1] * (v + (1.0f - v) * r[6]) +
               gb * r[2] * (v + (1.0f - v) * r[7]);
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) g += gs[c] * r[so + c];
          Gw[j] = g;
          float dT = (g_transp ? g_transp[(size_t)ray * S + s] : 0.f) + g * alpha[j];
          B[j] = dT * T[j];
          Bsum += B[j];
        }
      }
      // R_j = sum_{s > j} B_s : lanes above + later samples in this lane
      float R = warp_excl_suffix_sum(Bsum, lane);
#pragma unroll
      for (int j = SPL - 1; j >= 0; --j) {
        int s = lane * SPL + j;
        if (s < S) {
          const float* r = rows + s * n_out;
          float* go = grow + s * n_out;
          float dq = R / qv[j];
          float dalpha = Gw[j] * T[j] - dq;
          SampleVals sv = sample_alpha(zs, r, s, S);
          float dsig = (sv.sigma > 0.f) ? dalpha * sv.e * sv.delta : 0.f;
          float v = r[4];
          float wj = w[j];
          go[0] = gr * wj * (v + (1.0f - v) * r[5]);
          go[1] = gg * wj * (v + (1.0f - v) * r[6]);
          go[2] = gb * wj * (v + (1.0f - v) * r[7]);
          go[3] = dsig;
          go[4] = wj * (gr * r[0] * (1.0f - r[5]) + gg * r[1] * (1.0f - r[6]) + gb * r[2] * (1.0f - r[7]));
          go[5] = gr * wj * r[0] * (1.0f - v);
          go[6] = gg * wj * r[1] * (1.0f - v);
          go[7] = gb * wj * r[2] * (1.0f - v);
          go[8] = 0.f;
          if (so > 9) go[9] = 0.f;
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) go[so + c] = gs[c] * wj;
          for (int c = so + C; c < n_out; ++c) go[c] = 0.f;
        }
        R += B[j];
      }
      __syncwarp();
      float* dst = g_out + (size_t)ray * row_words;
      if (vec_rows) {
        float4* d4 = reinterpret_cast<float4*>(dst);
        const float4* g4 = reinterpret_cast<const float4*>(grow);
        if (g_direct) {
          const float4* gd4 = reinterpret_cast<const float4*>(g_direct + (size_t)ray * row_words);
          const int n4 = row_words / 4;
          for (int base = 0; base < n4; base += 256) {
            float4 t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int i = base + lane + 32 * k;
              if (i < n4) t[k] = __ldg(gd4 + i);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int i = base + lane + 32 * k;
              if (i < n4) {
                const float4 a = g4[i];
                d4[i] = make_float4(a.x + t[k].x, a.y + t[k].y, a.z + t[k].z, a.w + t[k].w);
              }
            }
          }
        } else {
#pragma unroll 8
          for (int i = lane; i < row_words / 4; i += 32) d4[i] = g4[i];
        }
      } else if (g_direct) {
        const float* gdir = g_direct + (size_t)ray * row_words;
        for (int i = lane; i < row_words; i += 32) dst[i] = grow[i] + __ldg(gdir + i);
      } else {
        for (int i = lane; i < row_words; i += 32) dst[i] = grow[i];
      }
    }
    __syncwarp();
  }
}

template <bool BWD>
static int launch_k3(const float* out, const float* z, int n_rays, int S, int n_out, int C, int flags, float* rgb,
                     float* depth, float* weights, float* transp, float* sem, long long* label,
                     const float* g_rgb, const float* g_depth, const float* g_w, const float* g_t,
                     const float* g_sem, const float* g_direct, float* g_out, cudaStream_t st) {
  const int spl = (S + 31) / 32;
  const size_t rw4 = (size_t)((S * n_out + 3) & ~3), s4 = (size_t)((S + 3) & ~3);
  const size_t smem = (size_t)K3_WARPS * (2 * (rw4 + s4) + (BWD ? rw4 : 0)) * sizeof(float);
  int sms = num_sms();
  if (sms <= 0) return SNB_ERR_NO_DEVICE;
  int blocks = (n_rays + K3_WARPS - 1) / K3_WARPS;
#define K3_LAUNCH(SPL_)                                                                               \
  do {                                                                                                \
    if (C == 0) K3_LAUNCH_C(SPL_, 0);                                                                 \
    else if (C <= 6) K3_LAUNCH_C(SPL_, 6);                                                            \
    else K3_LAUNCH_C(SPL_, 10);                                                                       \
  } while (0)
#define K3_LAUNCH_C(SPL_, CMAX_)                                                                      \
  do {                                                                                                \
    auto kfn = k3_composite_kernel<SPL_, BWD, CMAX_>;                                                 \
    /* one full wave of resident blocks, each looping over rays: a partial second wave would idle most SMs. */ \
    /* The attribute / occupancy queries cost several microseconds of host time each: once per shared-memory size. */ \
    static size_t cached_smem = ~(size_t)0;                                                           \
    static int cached_resident = 0;                                                                   \
    if (cached_smem != smem) {                                                                        \
      if (smem > 48 * 1024) SNB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      SNB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cached_resident, kfn, K3_WARPS * 32, smem)); \
      cached_smem = smem;                                                                             \
    }                                                                                                 \
    int resident = cached_resident;                                                                   \
    if (resident < 1) resident = 1;                                                                   \
    if (blocks > sms * resident) blocks = sms * resident;                                             \
    kfn<<<blocks, K3_WARPS * 32, smem, st>>>(out, z, n_rays, S, n_out, C, flags, rgb, depth, weights, transp, sem, label, \
                                             g_rgb, g_depth, g_w, g_t, g_sem, g_direct, g_out);       \
  } while (0)
  if (spl <= 1) K3_LAUNCH(1);
  else if (spl <= 2) K3_LAUNCH(2);
  else if (spl <= 4) K3_LAUNCH(4);
  else K3_LAUNCH(8);
#undef K3_LAUNCH
#undef K3_LAUNCH_C
  return launch_status("k3_composite_kernel");
}

static int check_k3(const float* out, const float* z, int n_rays, int S, int n_out, int C, int flags = 0) {
  SNB_CHECK_ARG(out && z, SNB_ERR_INVALID, "composite: null input");
  SNB_CHECK_ARG(n_rays >= 0, SNB_ERR_INVALID, "composite: n_rays < 0");
  // S < 2: the reference itself degenerates (framework/util/rendering.py:13 builds delta_inf from
  // an empty slice, so every per-sample tensor collapses to (N,0)); not supported here.
  SNB_CHECK_ARG(S >= 2 && S <= 256, SNB_ERR_UNSUPPORTED, "composite: n_samples %d outside [2,256]", S);
  // n_out > 9 + C is allowed: trailing columns are ignored (the solar pass composites no semantics)
  const int so = 9 + ((flags & SNB_COMPOSITE_BETA_S) ? 1 : 0);
  SNB_CHECK_ARG(C >= 0 && C <= 10 && n_out >= so + C && n_out <= 32, SNB_ERR_UNSUPPORTED,
                "composite: n_out %d / n_classes %d unsupported (need %d + C <= n_out <= 32, C <= 10)", n_out, C, so);
  return 0;
}

// =====================================================================================================================
// K3 + losses fused: composite forward, the training losses that sit directly on its outputs and the
// composite backward in ONE pass per ray - the per-sample weights / transparency / beta / sun_sc tensors and the ~40
// small PyTorch launches of the loss modules never exist.  Per ray the kernel forms the loss terms, their gradients
// w.r.t. the composited values analytically, and pushes them straight through the compositing backward into
// g_out (P, n_out), the gradient of the packed head outputs.
//   mode 0 (main pass):   colour loss - MSE (SNerfLoss, baseline/components/loss.py:71-94) or uncertainty-aware
//                         (SatNerfLoss, loss.py:16-27,50-68) - + semantic cross-entropy with ignore_index
//                         (semantic/components/loss.py:35-65) + car regularisation (loss.py:117-157)
//   mode 1 (solar pass):  solar-correction terms 2 and 3 (baseline/components/loss.py:4-13); transparency_sc / weights_sc
//                         are detached there, so only the sun column receives a gradient
//   mode 2 (depth batch): DepthLoss (baseline/components/loss.py:30-47)
//   mode 3 (statistics pre-pass of the uncertainty-weighted semantic loss, `use_beta_for_s`): SemanticUncertaintyLoss
//                         (semantic/components/loss.py:6-32,68-114) is lambda_s * CE_mean * mean_r(1 / (2 beta_r^2)) - a
//                         product of two batch means, so its gradient needs both means first.  The pre-pass accumulates
//                         loss_terms[0] += sum of the per-ray cross-entropies, loss_terms[1] += sum_r 1 / (2 beta_r^2) and
//                         writes no gradient; the caller hands them to the main pass as counts[4], counts[5].
// Means are over n_rays (inv_n), the CE mean over the non-ignored rays and the car term over the car rays: their counts
// come from a device buffer (counts[0], counts[1]) so no host synchronisation is needed.
// loss_terms (device float[8], accumulated): 0 colour, 1 log-beta (without the constant 3/2), 2 CE, 3 car, 4 sc term 2,
// 5 sc term 3, 6 depth.
// =====================================================================================================================
struct LossParams {
  int mode, color;
  float beta_min, inv_n;
  float lambda_s;
  int ignore_index;
  float lambda_c;
  int car_label;
  float lambda_sc, lambda_ds;
  int flags;
  int sem_unc;
};

template <int SPL, int CMAX>
__global__ void __launch_bounds__(K3_WARPS * 32, SPL >= 8 ? 1 : 5)
k3_loss_kernel(const float* __restrict__ out, const float* __restrict__ z_vals, int n_rays, int S, int n_out, int C,
               const float* __restrict__ gt_rgb, const long long* __restrict__ labels,
               const unsigned char* __restrict__ ray_mask, const float* __restrict__ depth_gt,
               const float* __restrict__ depth_w, const float* __restrict__ counts, const LossParams lp,
               float* __restrict__ g_out, float* __restrict__ loss_terms) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_words = S * n_out;
  // use_separate_beta_for_s: column 9 = the semantic uncertainty head; the uncertainty-weighted semantic loss then uses ITS
  // composited value and adds its own log-beta term (semantic/components/loss.py:15-31)
  const bool bsep = (lp.flags & SNB_COMPOSITE_BETA_S) != 0;
  const int so = 9 + (bsep ? 1 : 0);
  const int rw4 = (row_words + 3) & ~3, s4 = (S + 3) & ~3;
  const int in_words = rw4 + s4;
  const int per_warp = 2 * in_words + rw4;
  float* wbase = smem + warp * per_warp;
  float* grow = wbase + 2 * in_words;
  const bool vec_rows = (row_words & 3) == 0 && (((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(g_out)) & 15) == 0);
  const bool async_in = vec_rows && (S & 3) == 0 && ((reinterpret_cast<uintptr_t>(z_vals) & 15) == 0);
  auto fetch = [&](int ray, float* dst) {
    if (ray < n_rays) {
      const float4* s4p = reinterpret_cast<const float4*>(out + (size_t)ray * row_words);
      const uint32_t d = smem_u32(dst);
      for (int i = lane; i < row_words / 4; i += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + i * 16), "l"(s4p + i) : "memory");
      const float4* z4 = reinterpret_cast<const float4*>(z_vals + (size_t)ray * S);
      for (int i = lane; i < S / 4; i += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + (rw4 + i * 4) * 4), "l"(z4 + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const float inv_valid = counts ? 1.0f / fmaxf(counts[0], 1.0f) : 0.f;
  const float inv_car = counts ? 1.0f / fmaxf(counts[1], 1.0f) : 0.f;
  // uncertainty-weighted semantic loss: the two batch means of the statistics pre-pass (mode 3)
  const float unc_ce = (counts && lp.sem_unc) ? counts[4] * inv_valid : 0.f;      // CE_mean
  const float unc_M = (counts && lp.sem_unc) ? counts[5] * lp.inv_n : 1.f;        // mean_r 1 / (2 beta_r^2)
  float lsum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) lsum[i] = 0.f;

  const int ray0 = blockIdx.x * K3_WARPS + warp, stride = gridDim.x * K3_WARPS;
  int cur = 0;
  if (async_in) fetch(ray0, wbase);
  for (int ray = ray0; ray < n_rays; ray += stride, cur ^= 1) {
    float* rows = wbase + (async_in ? cur * in_words : 0);
    float* zs = rows + rw4;
    if (async_in) {
      fetch(ray + stride, wbase + (cur ^ 1) * in_words);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      const float* src = out + (size_t)ray * row_words;
      for (int i = lane; i < row_words; i += 32) rows[i] = __ldg(src + i);
      for (int i = lane; i < S; i += 32) zs[i] = __ldg(z_vals + (size_t)ray * S + i);
    }
    __syncwarp();

    // ---- composite forward (same arithmetic as k3_composite_kernel) ----------------------------------------
    float alpha[SPL], tloc[SPL], qv[SPL];
    float Q = 1.0f;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      int s = lane * SPL + j;
      tloc[j] = Q;
      if (s < S) {
        SampleVals v = sample_alpha(zs, rows + s * n_out, s, S);
        alpha[j] = v.alpha;
        qv[j] = v.q;
        Q *= v.q;
      } else {
        alpha[j] = 0.0f;
        qv[j] = 1.0f;
      }
    }
    const float prefix = warp_excl_prod(Q, lane);
    float acc_d = 0.f, acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_beta = 0.f, acc_bs = 0.f, acc_t2 = 0.f, acc_t3 = 0.f;
    float acc_s[CMAX > 0 ? CMAX : 1];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc_s[c] = 0.f;
    float T[SPL], w[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      int s = lane * SPL + j;
      T[j] = prefix * tloc[j];
      w[j] = alpha[j] * T[j];
      if (s < S) {
        const float* r = rows + s * n_out;
        const float v = r[4];
        acc_d += w[j] * zs[s];
        if (lp.mode == 3) {
          acc_beta += w[j] * r[bsep ? 9 : 8];
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) acc_s[c] += w[j] * r[so + c];
        } else if (lp.mode == 0) {
          acc_r += w[j] * r[0] * (v + (1.0f - v) * r[5]);
          acc_g += w[j] * r[1] * (v + (1.0f - v) * r[6]);
          acc_b += w[j] * r[2] * (v + (1.0f - v) * r[7]);
          acc_beta += w[j] * r[8];
          if (bsep) acc_bs += w[j] * r[9];
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) acc_s[c] += w[j] * r[so + c];
        } else if (lp.mode == 1) {
          acc_t2 += (T[j] - v) * (T[j] - v);
          acc_t3 += w[j] * v;
        }
      }
    }
    // gB / gBs: gradient w.r.t. the composited uncertainty (column 8 / the separate semantic one, column 9);
    // gB_w / gBs_w: the part that reaches only the compositing weights (`detach_beta_for_s` detaches the per-sample beta
    // values, not the weights they are summed with: semantic/components/loss.py:15-19)
    float gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f, gB = 0.f, gBs = 0.f, gB_w = 0.f, gBs_w = 0.f;
    float gs[CMAX > 0 ? CMAX : 1];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) gs[c] = 0.f;
    if (lp.mode == 3) {
      // statistics pre-pass: per-ray cross-entropy and 1 / (2 beta^2); no gradient rows
      acc_beta = warp_sum(acc_beta);
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) acc_s[c] = warp_sum(acc_s[c]);
      const float Bt = acc_beta + lp.beta_min;
      lsum[1] += 0.5f / (Bt * Bt);
      const bool on = labels != nullptr && C > 0 && (ray_mask == nullptr || ray_mask[ray] != 0);
      const int y = on ? (int)labels[ray] : -1;
      if (on && y != lp.ignore_index && y >= 0 && y < C) {
        float mx = acc_s[0];
#pragma unroll
        for (int c = 1; c < CMAX; ++c)
          if (c < C) mx = fmaxf(mx, acc_s[c]);
        float se = 0.f, ly = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) {
            se += expf(acc_s[c] - mx);
            if (c == y) ly = acc_s[c];
          }
        lsum[0] += mx + logf(se) - ly;
      }
      __syncwarp();
      continue;
    }
    if (lp.mode == 0) {
      acc_r = warp_sum(acc_r);
      acc_g = warp_sum(acc_g);
      acc_b = warp_sum(acc_b);
      acc_beta = warp_sum(acc_beta);
      if (bsep) acc_bs = warp_sum(acc_bs);
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) acc_s[c] = warp_sum(acc_s[c]);
      // colour loss on the clamped colour (rs_semantic.py:103); the clamp passes gradient inside [0, 1] only
      // (NeRF's inference does not clamp, nerf.py:73-86: SNB_COMPOSITE_NO_CLAMP)
      const bool clamp = !(lp.flags & SNB_COMPOSITE_NO_CLAMP);
      const float d0 = (clamp ? fminf(fmaxf(acc_r, 0.f), 1.f) : acc_r) - gt_rgb[ray * 3 + 0];
      const float d1 = (clamp ? fminf(fmaxf(acc_g, 0.f), 1.f) : acc_g) - gt_rgb[ray * 3 + 1];
      const float d2 = (clamp ? fminf(fmaxf(acc_b, 0.f), 1.f) : acc_b) - gt_rgb[ray * 3 + 2];
      const float sq = d0 * d0 + d1 * d1 + d2 * d2;
      const float k3n = lp.inv_n * (1.0f / 3.0f);
      if (lp.color == 1) {   // ((rgb - gt)^2 / (2 beta^2)).mean() + (3 + log(beta).mean()) / 2
        const float Bt = acc_beta + lp.beta_min;
        const float i2 = 1.0f / (Bt * Bt);
        lsum[0] += 0.5f * sq * i2 * k3n;
        lsum[1] += 0.5f * logf(Bt) * lp.inv_n;
        gr = d0 * i2 * k3n;
        gg = d1 * i2 * k3n;
        gb = d2 * i2 * k3n;
        gB = -sq * i2 / Bt * k3n + 0.5f * lp.inv_n / Bt;
      } else {               // mse_loss(rgb, gt)
        lsum[0] += sq * k3n;
        gr = 2.0f * d0 * k3n;
        gg = 2.0f * d1 * k3n;
        gb = 2.0f * d2 * k3n;
      }
      if (clamp) {
        if (!(acc_r >= 0.f && acc_r <= 1.f)) gr = 0.f;
        if (!(acc_g >= 0.f && acc_g <= 1.f)) gg = 0.f;
        if (!(acc_b >= 0.f && acc_b <= 1.f)) gb = 0.f;
      }
      // semantic_sparsity_mask (semantic/components/training_step.py:58-88): rays outside it take part in neither the
      // cross-entropy nor the car regularisation (loss.py:52-55,131-143); same predicate as snb_label_counts
      const bool sem_on = labels != nullptr && C > 0 && (ray_mask == nullptr || ray_mask[ray] != 0);
      if (sem_on) {
        const int y = (int)labels[ray];
        if (lp.lambda_s != 0.f && y != lp.ignore_index && y >= 0 && y < C) {
          float mx = acc_s[0];
#pragma unroll
          for (int c = 1; c < CMAX; ++c)
            if (c < C) mx = fmaxf(mx, acc_s[c]);
          float se = 0.f, ly = 0.f;
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) {
              se += expf(acc_s[c] - mx);
              if (c == y) ly = acc_s[c];
            }
          const float lse = mx + logf(se);
          // plain SemanticLoss: lambda_s * CE_mean.  Uncertainty-weighted (sem_unc): lambda_s * CE_mean * M with
          // M = mean_r 1 / (2 beta_r^2) from the pre-pass - the logits' gradient is scaled by M
          const float ks = lp.lambda_s * inv_valid * (lp.sem_unc ? unc_M : 1.0f);
          lsum[2] += ks * (lse - ly);
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) gs[c] = ks * (expf(acc_s[c] - lse) - (c == y ? 1.0f : 0.0f));
        }
        if (lp.lambda_c != 0.f && y == lp.car_label) {   // mse(1, sum w beta) over the car rays
          const float e = 1.0f - acc_beta;
          lsum[3] += lp.lambda_c * e * e * inv_car;
          gB += -2.0f * lp.lambda_c * e * inv_car;
        }
      }
      if (lp.sem_unc != 0 && lp.lambda_s != 0.f) {
        // d/d beta_r of lambda_s * CE_mean * mean_r 1 / (2 beta_r^2) = -lambda_s * CE_mean / (N beta_r^3) - for EVERY ray of
        // the batch, labelled or not: the mean over beta runs over all rays (loss.py:19-25).  With the separate head it is
        // beta_s, which also carries its own term lambda_s * (3 + mean log beta_s) / 2 (:27-30).  sem_unc == 2
        // (detach_beta_for_s): the per-sample values are detached, the weights are not - the gradient reaches sigma only
        const float Bt = (bsep ? acc_bs : acc_beta) + lp.beta_min;
        float g = -lp.lambda_s * unc_ce * lp.inv_n / (Bt * Bt * Bt);
        if (bsep) {
          lsum[7] += lp.lambda_s * 0.5f * logf(Bt) * lp.inv_n;
          g += lp.lambda_s * 0.5f * lp.inv_n / Bt;
        }
        const bool det = lp.sem_unc == 2;
        if (bsep) (det ? gBs_w : gBs) += g;
        else (det ? gB_w : gB) += g;
      }
    } else if (lp.mode == 1) {
      acc_t2 = warp_sum(acc_t2);
      acc_t3 = warp_sum(acc_t3);
      const float k = lp.lambda_sc * (1.0f / 3.0f) * lp.inv_n;
      lsum[4] += k * acc_t2;
      lsum[5] += k * (1.0f - acc_t3);
    } else {
      acc_d = warp_sum(acc_d);
      const float wt = depth_w ? depth_w[ray] : 1.0f;
      const float e = acc_d - depth_gt[ray];
      const float k = lp.lambda_ds * (1.0f / 3.0f) * lp.inv_n;
      lsum[6] += k * wt * e * e;
      gd = 2.0f * k * wt * e;
    }

    // ---- gradient of the packed rows ------------------------------------------------------------------------
    if (lp.mode == 1) {
      const float k = lp.lambda_sc * (1.0f / 3.0f) * lp.inv_n;
#pragma unroll
      for (int j = 0; j < SPL; ++j) {
        int s = lane * SPL + j;
        if (s < S) {
          float* go = grow + s * n_out;
          for (int c = 0; c < n_out; ++c) go[c] = 0.f;
          go[4] = k * (-2.0f * (T[j] - rows[s * n_out + 4]) - w[j]);   // T, w detached (loss.py:9-10)
        }
      }
    } else {
      float Gw[SPL], Bv[SPL];
      float Bsum = 0.f;
#pragma unroll
      for (int j = 0; j < SPL; ++j) {
        int s = lane * SPL + j;
        Gw[j] = 0.f;
        Bv[j] = 0.f;
        if (s < S) {
          const float* r = rows + s * n_out;
          const float v = r[4];
          float g = (gB + gB_w) * r[8] + gd * zs[s];
          if (bsep) g += (gBs + gBs_w) * r[9];
          g += gr * r[0] * (v + (1.0f - v) * r[5]) + gg * r[1] * (v + (1.0f - v) * r[6]) + gb * r[2] * (v + (1.0f - v) * r[7]);
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) g += gs[c] * r[so + c];
          Gw[j] = g;
          Bv[j] = g * alpha[j] * T[j];
          Bsum += Bv[j];
        }
      }
      float R = warp_excl_suffix_sum(Bsum, lane);
#pragma unroll
      for (int j = SPL - 1; j >= 0; --j) {
        int s = lane * SPL + j;
        if (s < S) {
          const float* r = rows + s * n_out;
          float* go = grow + s * n_out;
          const float dq = R / qv[j];
          const float dalpha = Gw[j] * T[j] - dq;
          SampleVals sv = sample_alpha(zs, r, s, S);
          const float v = r[4], wj = w[j];
          go[0] = gr * wj * (v + (1.0f - v) * r[5]);
          go[1] = gg * wj * (v + (1.0f - v) * r[6]);
          go[2] = gb * wj * (v + (1.0f - v) * r[7]);
          go[3] = (sv.sigma > 0.f) ? dalpha * sv.e * sv.delta : 0.f;
          go[4] = wj * (gr * r[0] * (1.0f - r[5]) + gg * r[1] * (1.0f - r[6]) + gb * r[2] * (1.0f - r[7]));
          go[5] = gr * wj * r[0] * (1.0f - v);
          go[6] = gg * wj * r[1] * (1.0f - v);
          go[7] = gb * wj * r[2] * (1.0f - v);
          go[8] = gB * wj;
          if (bsep) go[9] = gBs * wj;
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) go[so + c] = gs[c] * wj;
          for (int c = so + C; c < n_out; ++c) go[c] = 0.f;
        }
        R += Bv[j];
      }
    }
    __syncwarp();
    float* dst = g_out + (size_t)ray * row_words;
    if (vec_rows) {
      float4* d4 = reinterpret_cast<float4*>(dst);
      const float4* g4 = reinterpret_cast<const float4*>(grow);
#pragma unroll 8
      for (int i = lane; i < row_words / 4; i += 32) d4[i] = g4[i];
    } else {
      for (int i = lane; i < row_words; i += 32) dst[i] = grow[i];
    }
    __syncwarp();
  }
  // every lane of a warp holds the same per-ray sums: one atomic per warp and term
  if (lane < 8) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (lane == i) v = lsum[i];
    if (v != 0.f) atomicAdd(loss_terms + lane, v);
  }
}

static int launch_k3_loss(const float* out, const float* z, int n_rays, int S, int n_out, int C, const float* gt_rgb,
                          const long long* labels, const unsigned char* ray_mask, const float* depth_gt,
                          const float* depth_w, const float* counts,
                          const LossParams& lp, float* g_out, float* loss_terms, cudaStream_t st) {
  const int spl = (S + 31) / 32;
  const size_t rw4 = (size_t)((S * n_out + 3) & ~3), s4 = (size_t)((S + 3) & ~3);
  const size_t smem = (size_t)K3_WARPS * (2 * (rw4 + s4) + rw4) * sizeof(float);
  int sms = num_sms();
  if (sms <= 0) return SNB_ERR_NO_DEVICE;
  int blocks = (n_rays + K3_WARPS - 1) / K3_WARPS;
#define K3L_LAUNCH(SPL_)                                                                              \
  do {                                                                                                \
    if (C == 0) K3L_LAUNCH_C(SPL_, 0);                                                                \
    else if (C <= 6) K3L_LAUNCH_C(SPL_, 6);                                                           \
    else K3L_LAUNCH_C(SPL_, 10);                                                                      \
  } while (0)
#define K3L_LAUNCH_C(SPL_, CMAX_)                                                                     \
  do {                                                                                                \
    auto kfn = k3_loss_kernel<SPL_, CMAX_>;                                                           \
    static size_t cached_smem = ~(size_t)0;                                                           \
    static int cached_resident = 0;                                                                   \
    if (cached_smem != smem) {                                                                        \
      if (smem > 48 * 1024) SNB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      SNB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cached_resident, kfn, K3_WARPS * 32, smem)); \
      cached_smem = smem;                                                                             \
    }                                                                                                 \
    int resident = cached_resident;                                                                   \
    if (resident < 1) resident = 1;                                                                   \
    if (blocks > sms * resident) blocks = sms * resident;                                             \
    kfn<<<blocks, K3_WARPS * 32, smem, st>>>(out, z, n_rays, S, n_out, C, gt_rgb, labels, ray_mask, depth_gt, depth_w, counts, lp, \
                                             g_out, loss_terms);                                      \
  } while (0)
  if (spl <= 1) K3L_LAUNCH(1);
  else if (spl <= 2) K3L_LAUNCH(2);
  else if (spl <= 4) K3L_LAUNCH(4);
  else K3L_LAUNCH(8);
#undef K3L_LAUNCH
#undef K3L_LAUNCH_C
  return launch_status("k3_loss_kernel");
}

}  // namespace snb

// counts[0] += rays that enter the cross-entropy mean, counts[1] += rays that enter the car regularisation,
// counts[2] += labels outside [0, C) that are not ignore_index (PyTorch's CrossEntropyLoss raises for those; the caller checks)
namespace snb {
__global__ void __launch_bounds__(256)
label_counts_kernel(const long long* __restrict__ labels, const unsigned char* __restrict__ ray_mask, int n, int C,
                    int ignore_index, int car_label, float* __restrict__ counts) {
  int valid = 0, car = 0, oor = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const long long y = labels[i];
    const bool in = ray_mask == nullptr || ray_mask[i] != 0;
    const bool range = y >= 0 && y < C;
    if (in && y != ignore_index && range) ++valid;
    if (in && y == car_label) ++car;
    if (y != ignore_index && !range) ++oor;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    valid += __shfl_xor_sync(FULL, valid, d);
    car += __shfl_xor_sync(FULL, car, d);
    oor += __shfl_xor_sync(FULL, oor, d);
  }
  if ((threadIdx.x & 31) == 0) {
    if (valid) atomicAdd(counts + 0, (float)valid);
    if (car) atomicAdd(counts + 1, (float)car);
    if (oor) atomicAdd(counts + 2, (float)oor);
  }
}
}  // namespace snb

extern "C" int snb_label_counts(const int64_t* labels, const uint8_t* ray_mask, int n_rays, int n_classes, int ignore_index,
                                int car_label, float* counts, void* stream) {
  SNB_CHECK_ARG(labels && counts && n_rays >= 0 && n_classes >= 0, SNB_ERR_INVALID, "label_counts: bad argument");
  if (n_rays == 0) return 0;
  int blocks = (n_rays + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  snb::label_counts_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(labels), ray_mask, n_rays,
                                                                    n_classes, ignore_index, car_label, counts);
  return snb::launch_status("label_counts_kernel");
}

extern "C" int snb_composite_loss(const float* out, const float* z_vals, int n_rays, int n_samples, int n_out, int n_classes,
                                  const float* gt_rgb, const int64_t* labels, const uint8_t* ray_mask, const float* depth_gt,
                                  const float* depth_w, const float* counts, const snb_loss_params* p, float* g_out,
                                  float* loss_terms, void* stream) {
  if (int r = snb::check_k3(out, z_vals, n_rays, n_samples, n_out, n_classes, p ? p->flags : 0)) return r;
  SNB_CHECK_ARG(p && loss_terms && (g_out || p->mode == 3), SNB_ERR_INVALID, "composite_loss: null argument");
  SNB_CHECK_ARG(p->mode >= 0 && p->mode <= 3, SNB_ERR_INVALID, "composite_loss: mode %d", p->mode);
  SNB_CHECK_ARG(!(p->sem_unc != 0 && p->mode == 0) || (labels != nullptr && counts != nullptr), SNB_ERR_INVALID,
                "composite_loss: the uncertainty-weighted semantic loss needs labels and the pre-pass statistics in counts[4:6]");
  SNB_CHECK_ARG(p->mode != 0 || gt_rgb != nullptr, SNB_ERR_INVALID, "composite_loss: the colour loss needs gt_rgb");
  SNB_CHECK_ARG(p->mode != 2 || depth_gt != nullptr, SNB_ERR_INVALID, "composite_loss: the depth loss needs depth_gt");
  SNB_CHECK_ARG(labels == nullptr || counts != nullptr || p->mode != 0, SNB_ERR_INVALID,
                "composite_loss: labels need the device counts [n_valid, n_car]");
  if (n_rays == 0) return 0;
  snb::LossParams lp;
  lp.mode = p->mode; lp.color = p->color; lp.beta_min = p->beta_min; lp.inv_n = p->inv_n; lp.lambda_s = p->lambda_s;
  lp.ignore_index = p->ignore_index; lp.lambda_c = p->lambda_c; lp.car_label = p->car_label; lp.lambda_sc = p->lambda_sc;
  lp.lambda_ds = p->lambda_ds;
  lp.flags = p->flags;
  lp.sem_unc = p->sem_unc;
  return snb::launch_k3_loss(out, z_vals, n_rays, n_samples, n_out, n_classes, gt_rgb, reinterpret_cast<const long long*>(labels),
                             ray_mask, depth_gt, depth_w, counts, lp, g_out, loss_terms, (cudaStream_t)stream);
}

extern "C" int snb_composite_forward(const float* out, const float* z_vals, int n_rays, int n_samples,
                                     int n_out, int n_classes, int flags, float* rgb, float* depth, float* weights,
                                     float* transparency, float* sem_logits, int64_t* sem_label,
                                     void* stream) {
  if (int r = snb::check_k3(out, z_vals, n_rays, n_samples, n_out, n_classes, flags)) return r;
  SNB_CHECK_ARG(rgb && depth && weights && transparency, SNB_ERR_INVALID, "composite_forward: null output");
  SNB_CHECK_ARG(n_classes == 0 || (sem_logits && sem_label), SNB_ERR_INVALID,
                "composite_forward: semantic outputs required when n_classes > 0");
  if (n_rays == 0) return 0;
  return snb::launch_k3<false>(out, z_vals, n_rays, n_samples, n_out, n_classes, flags, rgb, depth, weights,
                               transparency, sem_logits, reinterpret_cast<long long*>(sem_label), nullptr,
                               nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int snb_composite_backward(const float* out, const float* z_vals, int n_rays, int n_samples,
                                      int n_out, int n_classes, int flags, const float* g_rgb, const float* g_depth,
                                      const float* g_weights, const float* g_transparency,
                                      const float* g_sem_logits, const float* g_out_direct, float* g_out,
                                      void* stream) {
  if (int r = snb::check_k3(out, z_vals, n_rays, n_samples, n_out, n_classes, flags)) return r;
  SNB_CHECK_ARG(g_out, SNB_ERR_INVALID, "composite_backward: null g_out");
  if (n_rays == 0) return 0;
  return snb::launch_k3<true>(out, z_vals, n_rays, n_samples, n_out, n_classes, flags, nullptr, nullptr, nullptr,
                              nullptr, nullptr, nullptr, g_rgb, g_depth, g_weights, g_transparency,
                              g_sem_logits, g_out_direct, g_out, (cudaStream_t)stream);
}
