// Small per-ray / per-parameter kernels around K2: sky_color and embedding parameter gradients,
// and the Adam step on the flat parameter buffer.
#include "snb_common.cuh"

struct snb_model;
extern "C" int64_t snb_model_param_count(const snb_model* m);
extern "C" int snb_model_num_tensors(const snb_model* m);
extern "C" int snb_model_tensor_info(const snb_model* m, int i, const char** name, int64_t* offset, int* rows, int* cols);

namespace snb {

// sky_color(sun_d) = sigmoid(W2 relu(W1 sun_d + b1) + b2) (satnerf.py:188-193) is evaluated once
// per ray; its parameter gradients are reduced here from the per-sample gradient of `out[:, 5:8]`.
// The embedding gradient (semantic/components/rendering.py:42: models["t"](ts)) is the per-ray sum of
// the gradient of the aux columns 4..4+tau, scattered by ts.
__global__ void __launch_bounds__(256)
ray_param_bwd_kernel(const float* __restrict__ extras, const float* __restrict__ sky, const float* __restrict__ g_out,
                     const float* __restrict__ g_aux, int n_rays, int S, int n_out, int tau, int vocab, int hidden,
                     const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                     float* __restrict__ gw1, float* __restrict__ gb1, float* __restrict__ gw2, float* __restrict__ gb2,
                     float* __restrict__ g_t_table) {
  __shared__ float part[8][16];
  __shared__ float fin[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nv = 3 + tau;
  // hidden unit(s) of this thread: tid (and tid + 256 with fc_use_full_features, hidden = 512)
  constexpr int HPT = 2;
  float a_w2[HPT][3], a_w1[HPT][3], a_b1[HPT], a_b2[3] = {0.f, 0.f, 0.f};
  float w1r[HPT][3], b1r[HPT], w2r[HPT][3];
#pragma unroll
  for (int u = 0; u < HPT; ++u) {
    const int h = tid + 256 * u;
    const bool on = sky && h < hidden;
    a_b1[u] = 0.f;
    b1r[u] = on ? b1[h] : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      a_w2[u][c] = a_w1[u][c] = 0.f;
      w1r[u][c] = on ? w1[h * 3 + c] : 0.f;
      w2r[u][c] = on ? w2[c * hidden + h] : 0.f;
    }
  }
  for (int ray = blockIdx.x; ray < n_rays; ray += gridDim.x) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0.f;
    for (int s = tid; s < S; s += 256) {
      const size_t p = (size_t)ray * S + s;
      if (sky) {
        v[0] += g_out[p * n_out + 5]; v[1] += g_out[p * n_out + 6]; v[2] += g_out[p * n_out + 7];
      }
      if (g_aux)
        for (int j = 0; j < tau; ++j) v[3 + j] += g_aux[p * 16 + 4 + j];
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < nv) {
        float x = v[j];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if (lane == 0) part[warp][j] = x;
      }
    }
    __syncthreads();
    if (tid < nv) {
      float x = 0.f;
      for (int wv = 0; wv < 8; ++wv) x += part[wv][tid];
      fin[tid] = x;
    }
    __syncthreads();
    const float* e = extras + (size_t)ray * 4;
    if (g_aux && tid < tau) {
      int ti = (int)e[3];
      ti = min(max(ti, 0), vocab - 1);
      atomicAdd(g_t_table + (size_t)ti * tau + tid, fin[3 + tid]);
    }
    if (sky) {
      const float sx = e[0], sy = e[1], sz = e[2];
      float d[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float o = sky[(size_t)ray * 3 + c];
        d[c] = fin[c] * o * (1.0f - o);
        if (tid == 0) a_b2[c] += d[c];
      }
#pragma unroll
      for (int u = 0; u < HPT; ++u) {
        if (tid + 256 * u < hidden) {
          const float y = fmaxf(fmaf(w1r[u][2], sz, fmaf(w1r[u][1], sy, fmaf(w1r[u][0], sx, b1r[u]))), 0.f);
          float dy = 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            a_w2[u][c] += d[c] * y;
            dy += d[c] * w2r[u][c];
          }
          if (y <= 0.f) dy = 0.f;
          a_w1[u][0] += dy * sx; a_w1[u][1] += dy * sy; a_w1[u][2] += dy * sz;
          a_b1[u] += dy;
        }
      }
    }
    __syncthreads();
  }
  if (sky) {
#pragma unroll
    for (int u = 0; u < HPT; ++u) {
      const int h = tid + 256 * u;
      if (h < hidden) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          atomicAdd(gw2 + c * hidden + h, a_w2[u][c]);
          atomicAdd(gw1 + h * 3 + c, a_w1[u][c]);
        }
        atomicAdd(gb1 + h, a_b1[u]);
      }
    }
    if (tid == 0)
#pragma unroll
      for (int c = 0; c < 3; ++c) atomicAdd(gb2 + c, a_b2[c]);
  }
}

// torch.optim.Adam (amsgrad=False, weight_decay=0, maximize=False)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float gscale,
            const int* __restrict__ step_dev) {
  if (step_dev != nullptr) {   // a captured CUDA graph replays with the step counter read from the device
    const float t = (float)*step_dev;
    bc1 = 1.0f - powf(b1, t);
    bc2_sqrt = sqrtf(1.0f - powf(b2, t));
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

}  // namespace snb

static int64_t tensor_offset(const snb_model* m, const char* want) {
  const int n = snb_model_num_tensors(m);
  for (int i = 0; i < n; ++i) {
    const char* name;
    int64_t off;
    int r, c;
    snb_model_tensor_info(m, i, &name, &off, &r, &c);
    if (strcmp(name, want) == 0) return off;
  }
  return -1;
}

extern "C" int snb_ray_param_backward(const snb_model* m, const float* params, const float* extras, const float* sky,
                                      const float* g_out, const float* g_aux, int n_rays, int n_samples, int n_out,
                                      int tau, int vocab, float* grads, float* g_t_table, void* stream) {
  using namespace snb;
  SNB_CHECK_ARG(m && params && extras && grads, SNB_ERR_INVALID, "ray_param_backward: null argument");
  SNB_CHECK_ARG(!sky || g_out, SNB_ERR_INVALID, "ray_param_backward: g_out required with sky");
  SNB_CHECK_ARG(!g_aux || g_t_table, SNB_ERR_INVALID, "ray_param_backward: g_t_table required with g_aux");
  SNB_CHECK_ARG(tau >= 0 && tau <= 12, SNB_ERR_UNSUPPORTED, "ray_param_backward: tau %d", tau);
  if (n_rays <= 0 || (!sky && !g_aux)) return 0;
  const int64_t w1 = tensor_offset(m, "sky_color.0.weight"), b1 = tensor_offset(m, "sky_color.0.bias"),
                w2 = tensor_offset(m, "sky_color.2.weight"), b2 = tensor_offset(m, "sky_color.2.bias");
  SNB_CHECK_ARG(w1 >= 0 && b1 >= 0 && w2 >= 0 && b2 >= 0, SNB_ERR_INVALID, "ray_param_backward: sky tensors missing");
  int sms = num_sms();
  if (sms <= 0) return SNB_ERR_NO_DEVICE;
  int hidden = 0, w1_cols = 0;
  {
    const int nt = snb_model_num_tensors(m);
    for (int i = 0; i < nt; ++i) {
      const char* name;
      int64_t off;
      snb_model_tensor_info(m, i, &name, &off, &hidden, &w1_cols);
      if (strcmp(name, "sky_color.0.weight") == 0) break;
    }
  }
  SNB_CHECK_ARG(hidden > 0 && hidden <= 512, SNB_ERR_UNSUPPORTED, "ray_param_backward: sky_color hidden width %d", hidden);
  int blocks = n_rays < sms * 4 ? n_rays : sms * 4;
  ray_param_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      extras, sky, g_out, g_aux, n_rays, n_samples, n_out, tau, vocab, hidden, params + w1, params + b1, params + w2,
      grads + w1, grads + b1, grads + w2, grads + b2, g_t_table);
  return launch_status("ray_param_bwd_kernel");
}

extern "C" int snb_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, int step, const int* step_dev, float grad_scale,
                             void* stream) {
  using namespace snb;
  SNB_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && n >= 0 && (step >= 1 || step_dev != nullptr), SNB_ERR_INVALID,
                "adam_step: bad argument");
  if (step < 1) step = 1;
  if (n == 0) return 0;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2 = 1.0f - powf(beta2, (float)step);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                             bc1, sqrtf(bc2), grad_scale, step_dev);
  return launch_status("adam_kernel");
}


// ---- vanilla NeRF: the per-sample aux row [1, Mapping(4, 3)(dir), 0...] (32 bf16) ---------------------------------------
namespace snb {
__global__ void __launch_bounds__(256) nerf_aux_kernel(const float* __restrict__ dirs, int stride, long long n_rays, int S,
                                                       __nv_bfloat16* __restrict__ aux) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_rays * S) return;
  const float* d = dirs + (p / S) * stride;
  float v[32];
  v[0] = 1.0f;
  float f = 1.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k, f *= 2.0f) {   // commons.py:68-74: per frequency sin(3) then cos(3)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = f * d[c];
      v[1 + k * 6 + c] = sinf(x);
      v[1 + k * 6 + 3 + c] = cosf(x);
    }
  }
#pragma unroll
  for (int j = 25; j < 32; ++j) v[j] = 0.f;
  uint4* dst = reinterpret_cast<uint4*>(aux + p * 32);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    dst[q] = make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                        pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
}
}  // namespace snb

extern "C" int snb_nerf_aux(const float* dirs, int stride, int n_rays, int n_samples, void* aux32, void* stream) {
  using namespace snb;
  SNB_CHECK_ARG(dirs && aux32 && stride >= 3 && n_rays >= 0 && n_samples >= 1, SNB_ERR_INVALID, "nerf_aux: bad argument");
  SNB_CHECK_ARG(((uintptr_t)aux32 & 15) == 0, SNB_ERR_INVALID, "nerf_aux: aux must be 16-byte aligned");
  const long long P = (long long)n_rays * n_samples;
  if (P == 0) return 0;
  nerf_aux_kernel<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dirs, stride, n_rays, n_samples,
                                                                                 (__nv_bfloat16*)aux32);
  return launch_status("nerf_aux_kernel");
}
