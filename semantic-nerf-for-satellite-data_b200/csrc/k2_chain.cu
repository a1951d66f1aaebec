// K2 chained MLP kernel - see k2_chain.cuh for the design.  sm_100a only (tcgen05 / TMEM / TMA).
//
// Roles (19 warps, one CTA per SM, clusters of 2 = SM pairs, cta_group::2 MMAs over 256 x 256 x 64):
//   warp 0      TMA producer for A (activations; waits for the producing layer of the same slot)
//   warp 1      tcgen05.mma issuer of the leader CTA (+ TMEM owner)
//   warp 2      TMA producer for B (weights)
//   warps 3-18  epilogue: two groups of 8 warps.  Group g takes the 64-column chunks {g, g+2} of every
//               256-column tile; inside a group, warp w and warp w+4 read the same TMEM lane quadrant and
//               split the chunk's columns 0-31 / 32-63, so every SM sub-partition has 4 epilogue warps to
//               hide the TMEM-load / MUFU / shared-memory latencies behind each other.
#include "k2_chain.cuh"
#include "k2_ptx.cuh"

#include <string.h>

namespace snb {

bool profile_gemm_begin(cudaStream_t st, double macs, int epi, int M, int N, int K, int cg, int splits);
void profile_gemm_end(cudaStream_t st);

namespace {

struct Cursor {   // position in this SM pair's tile sequence: group, layer, slot, n-tile
  int g, l, s, j;
  bool done;
};

struct Seq {
  int n_my;       // row blocks this pair owns: pair, pair + n_pairs, ...
  int n_layers;
  __device__ __forceinline__ int nslots(int g) const {
    const int r = n_my - g * CHAIN_SLOTS;
    return r < CHAIN_SLOTS ? r : CHAIN_SLOTS;
  }
};

__device__ __forceinline__ void cur_init(Cursor& c, const Seq& q) {
  c.g = c.l = c.s = c.j = 0;
  c.done = q.n_my <= 0;
}
__device__ __forceinline__ void cur_next(Cursor& c, const Seq& q, const ChainArgs& a) {
  if (++c.j < a.layers[c.l].n_tiles) return;
  c.j = 0;
  if (++c.s < q.nslots(c.g)) return;
  c.s = 0;
  if (++c.l < q.n_layers) return;
  c.l = 0;
  ++c.g;
  if (c.g * CHAIN_SLOTS >= q.n_my) c.done = true;
}

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }


// ---- epilogue math, specialised per mode so the inner loops are branch-free ---------------------------------
enum ChunkMode { CM_LINEAR = 0, CM_SIN = 1, CM_SIN_MASK = 2 };

// 32 accumulator columns of one row -> 16 packed bf16x2 words (+ the sign bits of the SIREN derivative).
// bsm: shared-memory address of this thread's 32 bias values, pre-multiplied by w0.
// Sign-mask layout (one 32-bit word per row and 32 columns): bit k = column 2k, bit 16 + k = column 2k + 1, so that
// the backward epilogue flips the signs of a packed bf16x2 pair with one shift and one logic op.
template <int MODE>
__device__ __forceinline__ void chunk_math(const uint32_t (&v)[32], uint32_t bsm, float w0, uint32_t (&outw)[16],
                                           uint32_t& mbits) {
  uint32_t me = 0, mo = 0;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    uint32_t b0, b1, b2, b3;
    ld_shared_v4(bsm + g * 16, b0, b1, b2, b3);
    float y[4];
    y[0] = fmaf(__uint_as_float(v[g * 4 + 0]), w0, __uint_as_float(b0));
    y[1] = fmaf(__uint_as_float(v[g * 4 + 1]), w0, __uint_as_float(b1));
    y[2] = fmaf(__uint_as_float(v[g * 4 + 2]), w0, __uint_as_float(b2));
    y[3] = fmaf(__uint_as_float(v[g * 4 + 3]), w0, __uint_as_float(b3));
    if (MODE == CM_SIN_MASK) {
      // y = n*pi + r, |r| <= pi/2: cos(y) = (-1)^n cos(r), so its sign is the parity of n = rint(y/pi), read off the
      // mantissa after adding 1.5 * 2^23; bits enter at the top of me / mo (even / odd columns) in column order
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t u = __float_as_uint(fmaf(y[j], 0.318309886183790672f, 12582912.0f));
        if (j & 1) mo = __funnelshift_r(mo, u, 1);
        else me = __funnelshift_r(me, u, 1);
      }
    }
    if (MODE != CM_LINEAR) {
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] = __sinf(y[j]);
    }
    outw[g * 2] = pack_bf16x2(y[0], y[1]);
    outw[g * 2 + 1] = pack_bf16x2(y[2], y[3]);
  }
  mbits = (me >> 16) | (mo & 0xffff0000u);
}

// dgrad: accumulator * multiplicand, in place in the staged multiplicand tile (this thread's 4 x 16 bytes of a row).
// SIREN: the multiplicand is the derivative w0 cos(.) rebuilt from the saved activation h = sin(.) (|h| <= 1 in bf16,
// so 1 - h^2 >= 0 exactly) and its sign bit: |cos| = sqrt(1 - h^2); the signs are applied to the packed bf16x2 pairs.
template <bool SIREN, bool W0ONE>
__device__ __forceinline__ void chunk_mul(const uint32_t (&v)[32], uint32_t buf, uint32_t row_off, uint32_t sw, int half,
                                          float w0, uint32_t mw) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint32_t off = row_off + ((((uint32_t)(half * 4 + g)) ^ sw) << 4);
    uint32_t q[4], o[4];
    ld_shared_v4(buf + off, q[0], q[1], q[2], q[3]);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float f0 = bf16_lo(q[p]), f1 = bf16_hi(q[p]);
      if (SIREN) {
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(f0) : "f"(fmaf(-f0, f0, 1.0f)));
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(f1) : "f"(fmaf(-f1, f1, 1.0f)));
        if (!W0ONE) { f0 *= w0; f1 *= w0; }
      }
      o[p] = pack_bf16x2(__uint_as_float(v[g * 8 + 2 * p]) * f0, __uint_as_float(v[g * 8 + 2 * p + 1]) * f1);
      if (SIREN) o[p] ^= (mw << (15 - (g * 4 + p))) & 0x80008000u;
    }
    st_shared_v4(buf + off, o[0], o[1], o[2], o[3]);
  }
}

}  // namespace

__global__ void __launch_bounds__(CHAIN_THREADS, 1) snb_chain_kernel(const __grid_constant__ ChainArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = sA + CHAIN_A_STAGES * 16384;
  uint8_t* sStg = smem + CHAIN_RING_BYTES;
  float* bias_smem = reinterpret_cast<float*>(sStg + GEMM_NUM_STAGING * GEMM_STAGING);   // [group][buffer][128]: the bias of a group's two chunks
  uint64_t* fullA = reinterpret_cast<uint64_t*>(bias_smem + 512);
  uint64_t* emptyA = fullA + 8;
  uint64_t* fullB = emptyA + 8;
  uint64_t* emptyB = fullB + 8;
  uint64_t* tfull = emptyB + 8;
  uint64_t* tempty = tfull + 2;
  uint64_t* mfull = tempty + 2;
  uint64_t* ready = mfull + GEMM_NUM_STAGING;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ready + CHAIN_SLOTS);

  const int warp = threadIdx.x >> 5;  // warp-uniform
  const int lane = threadIdx.x & 31;
  uint32_t cta_rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const bool lead_cta = (cta_rank == 0);
  const int pair = (int)(blockIdx.x >> 1);
  const int n_pairs = (int)(gridDim.x >> 1);
  Seq seq;
  seq.n_my = pair < args.n_blocks ? (args.n_blocks - pair + n_pairs - 1) / n_pairs : 0;
  seq.n_layers = args.n_layers;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) {
      mbar_init(&fullA[s], 2);   // the leader's expect_tx arrive + the peer's remote arrive
      mbar_init(&emptyA[s], 1);
      mbar_init(&fullB[s], 2);
      mbar_init(&emptyB[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 2 * CHAIN_EPI_WARPS * 32);
    }
    for (int s = 0; s < GEMM_NUM_STAGING; ++s) mbar_init(&mfull[s], 1);
    for (int s = 0; s < CHAIN_SLOTS; ++s) mbar_init(&ready[s], 2);   // the two epilogue group leaders
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // peer barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer: operand A ===========================
    uint32_t stage = 0, phase = 0, rbits = 0;
    Cursor c;
    for (cur_init(c, seq); !c.done; cur_next(c, seq, args)) {
      const ChainLayer& ly = args.layers[c.l];
      if (c.j == 0 && !(c.g == 0 && c.l == 0)) {
        // the previous tile-set of this slot (the layer that produced this layer's input rows) has been
        // stored completely; one barrier phase per tile-set keeps producer and epilogue in lock step
        mbar_wait(&ready[c.s], (rbits >> c.s) & 1u);
        rbits ^= 1u << c.s;
        fence_proxy_async_all();
      }
      const int blk = (c.g * CHAIN_SLOTS + c.s) * n_pairs + pair;
      const int row_real = blk * 256 + (int)cta_rank * GEMM_BLOCK_M;
      const int row_scr = (pair * CHAIN_SLOTS + c.s) * 256 + (int)cta_rank * GEMM_BLOCK_M;
      const int kb_total = ly.kb_total;
      int sg = 0, kk = 0;
      for (int kb = 0; kb < kb_total; ++kb, ++kk) {
        while (sg + 1 < ly.nseg && kk >= ly.seg_kb[sg]) {
          kk -= ly.seg_kb[sg];
          ++sg;
        }
        mbar_wait(&emptyA[stage], phase ^ 1);
        if (elect_one()) {
          if (lead_cta) mbar_expect_tx(&fullA[stage], 2 * 16384);
          else mbar_arrive_remote(&fullA[stage], 0);
          tma_load_2d_2sm(sA + stage * 16384, &args.maps[c.l].tmA[sg], &fullA[stage], kk * GEMM_BLOCK_K,
                          ly.a_scratch[sg] ? row_scr : row_real);
        }
        __syncwarp();
        if (++stage == CHAIN_A_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ===================================== TMA producer: operand B ===========================
    uint32_t stage = 0, phase = 0;
    Cursor c;
    for (cur_init(c, seq); !c.done; cur_next(c, seq, args)) {
      const ChainLayer& ly = args.layers[c.l];
      const bool rows16 = ly.epi == EPI_HEADOUT;   // N = 16: this CTA holds 8 of the 16 weight rows
      const int n_row = rows16 ? (int)cta_rank * 8 : c.j * 256 + (int)cta_rank * 128;
      const uint32_t b_bytes = rows16 ? 2u * 1024u : 2u * 16384u;
      const int kb_total = ly.kb_total;
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&emptyB[stage], phase ^ 1);
        if (elect_one()) {
          if (lead_cta) mbar_expect_tx(&fullB[stage], b_bytes);
          else mbar_arrive_remote(&fullB[stage], 0);
          tma_load_2d_2sm_hint(sB + stage * 16384, &args.maps[c.l].tmB, &fullB[stage], kb * GEMM_BLOCK_K, n_row, L2_EVICT_LAST);
        }
        __syncwarp();
        if (++stage == CHAIN_B_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA) ===========================
    if (lead_cta) {
      // cute::UMMA::InstrDescriptor: f32 accumulate, bf16 x bf16, both K-major, N = 256, M = 256 (SM pair)
      const uint32_t idesc256 = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((256u >> 4) << 24);
      const uint64_t desc0 = umma_desc(0, 16u, 1024u);   // K-major SW128: 8-row groups 1024 B apart
      const uint32_t a_base = smem_u32(sA) >> 4, b_base = smem_u32(sB) >> 4;
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
      Cursor c;
      for (cur_init(c, seq); !c.done; cur_next(c, seq, args), ++it) {
        const int kb_total = args.layers[c.l].kb_total;
        const uint32_t idesc = args.layers[c.l].epi == EPI_HEADOUT ? idesc16 : idesc256;
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&fullA[sa], pa);
          mbar_wait(&fullB[sb], pb);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = desc0 + (uint64_t)(a_base + sa * 1024u);
            const uint64_t bd = desc0 + (uint64_t)(b_base + sb * 1024u);
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
              tc_mma_bf16_2sm(d_tmem, ad + k * 2, bd + k * 2, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit_2sm(&emptyA[sa]);   // frees the slots in both CTAs once these MMAs have read them
            tc_commit_2sm(&emptyB[sb]);
          }
          __syncwarp();
          if (++sa == CHAIN_A_STAGES) { sa = 0; pa ^= 1; }
          if (++sb == CHAIN_B_STAGES) { sb = 0; pb ^= 1; }
        }
        if (elect_one()) tc_commit_2sm(&tfull[acc]);   // accumulator complete -> both CTAs' epilogues
        __syncwarp();
      }
    }
  } else {
    // ===================================== epilogue (2 groups x 256 threads) =================
    const int ewarp = warp - 3;               // 0..15
    const int grp = ewarp >> 3;               // epilogue group
    const int half = (ewarp >> 2) & 1;        // which 32 columns of a 64-column chunk
    const int quad = warp & 3;                // TMEM lane quadrant this warp may read (warp id % 4)
    const int row = quad * 32 + lane;         // row of the 128-row tile == TMEM lane
    const int gtid = (ewarp & 7) * 32 + lane; // 0..255 within the group
    const bool leader = (gtid == 0);
    const uint32_t stg0 = smem_u32(sStg) + grp * 2 * GEMM_STAGING;   // this group's two staging buffers
    uint8_t* stg_ptr = sStg + grp * 2 * GEMM_STAGING;
    uint64_t* gmfull = mfull + grp * 2;
    const uint32_t row_off = (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    const int bar_id = 1 + grp;
    auto gbar = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory"); };

    uint32_t it = 0;
    uint32_t cn = 0;          // chunks this group has committed (one bulk group per chunk); chunk cn stages in buffer cn & 1
    uint32_t mpar = 0;        // phase bits of the two mul-operand barriers
    bool cur_prefetched = false;
    // completion signals owed to the A producer: slot + the commit count that must have completed (leader only)
    int npend = 0, pend_slot0 = 0, pend_slot1 = 0;
    uint32_t pend_cn0 = 0, pend_cn1 = 0;

    auto confirm = [&](uint32_t completed) {   // leader only
      while (npend > 0 && pend_cn0 <= completed) {
        fence_proxy_async_all();
        mbar_arrive(&ready[pend_slot0]);
        pend_cn0 = pend_cn1;
        pend_slot0 = pend_slot1;
        --npend;
      }
    };
    // a tile-set's rows must be in L2 / HBM before the A producer may read them back: checked right before this
    // leader issues its next store, i.e. one chunk of compute after the stores in question were issued
    auto confirm_pending = [&]() {   // leader only
      if (npend > 0) {
        bulk_wait_done<0>();
        confirm(cn);
      }
    };
    // bias of this group's two chunks of a tile: column (grp + 2 * (i >> 6)) * 64 + (i & 63) for i < 128
    float* gbias = bias_smem + grp * 256;
    auto bias_fetch = [&](const Cursor& t) -> float {
      const float* bp = args.layers[t.l].epi == EPI_HEADOUT ? nullptr : args.layers[t.l].bias;
      // staged pre-multiplied by w0: the epilogue computes fma(acc, w0, w0 * bias)
      return (gtid < 128 && bp != nullptr) ? args.layers[t.l].w0 * __ldg(bp + t.j * 256 + (grp + 2 * (gtid >> 6)) * 64 + (gtid & 63)) : 0.f;
    };

    Cursor c, nx;
    cur_init(c, seq);
    nx = c;
    if (!nx.done) cur_next(nx, seq, args);
    if (!c.done) {
      const float b0 = bias_fetch(c);
      if (gtid < 128) gbias[gtid] = b0;
      gbar();
    }
    for (; !c.done; c = nx, cur_next(nx, seq, args), ++it) {
      const ChainLayer& ly = args.layers[c.l];
      const int epi = ly.epi;
      const float w0 = ly.w0;
      const uint32_t* mask = ly.mask;
      const int mask_ld = ly.mask_ld;
      const bool siren = ly.mul_siren != 0;
      const bool next_mul = !nx.done && args.layers[nx.l].epi == EPI_MUL;
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const int blk = (c.g * CHAIN_SLOTS + c.s) * n_pairs + pair;
      const int m_real = blk * 256 + (int)cta_rank * GEMM_BLOCK_M;
      const int m_scr = (pair * CHAIN_SLOTS + c.s) * 256 + (int)cta_rank * GEMM_BLOCK_M;
      const int m_out = ly.o_scratch ? m_scr : m_real;
      const int n0 = c.j * 256;
      const bool row_ok = m_real + row < args.M;
      const float* tbias = gbias + (it & 1) * 128;
      float nbias = 0.f;   // next tile's bias value staged by threads 0..127 of the group

      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * 256;

      if (epi == EPI_HEADOUT) {
        // ---- 16 head pre-activations per row: one warp per TMEM lane quadrant, straight from / to global memory ----
        if (grp == 0 && half == 0) {
          uint32_t v[16];
          tmem_ld16(taddr, v);
          tc_wait_ld();
          if (row_ok) {
            const long long grow = (long long)m_real + row;
            float x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(v[j]);
            if (ly.part != nullptr && ly.rows_mode >= 1) {
              const float4* pp = reinterpret_cast<const float4*>(ly.part + grow * 16);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 t4 = pp[j];
                x[4 * j] += t4.x; x[4 * j + 1] += t4.y; x[4 * j + 2] += t4.z; x[4 * j + 3] += t4.w;
              }
            }
            if (ly.rows_mode < 2) {
              float4* pp = reinterpret_cast<float4*>(ly.part + grow * 16);
#pragma unroll
              for (int j = 0; j < 4; ++j) pp[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
            } else {
              if (ly.bias != nullptr) {
#pragma unroll
                for (int j = 0; j < 16; ++j) x[j] += __ldg(ly.bias + j);
              }
              float* o = args.out_packed + grow * args.n_out;
              const int hm = args.head_mask;
              // rs_semantic.py:282-284: rgb = sigmoid(.) * (1 + 2*0.001) - 0.001
#pragma unroll
              for (int j = 0; j < 3; ++j) o[j] = (hm & SNB_HEAD_RGB) ? (1.0f / (1.0f + expf(-x[j]))) * 1.002f - 0.001f : 0.f;
              o[3] = (hm & SNB_HEAD_SIGMA) ? (x[3] > 20.f ? x[3] : log1pf(expf(x[3]))) : 0.f;
              o[4] = (hm & SNB_HEAD_SUN) ? 1.0f / (1.0f + expf(-x[4])) : 0.f;
              if ((hm & SNB_HEAD_SKY) && args.sky != nullptr) {
                const long long ray = args.rows_per_ray > 0 ? grow / args.rows_per_ray : grow;
                o[5] = __ldg(args.sky + ray * 3);
                o[6] = __ldg(args.sky + ray * 3 + 1);
                o[7] = __ldg(args.sky + ray * 3 + 2);
              } else {
                o[5] = o[6] = o[7] = 0.f;
              }
              o[8] = (hm & SNB_HEAD_BETA) ? (x[5] > 20.f ? x[5] : log1pf(expf(x[5]))) : 0.f;
#pragma unroll
              for (int cc = 0; cc < 10; ++cc) {
                if (cc < args.n_classes) {
                  const float sv = x[6 + cc];
                  o[9 + cc] = (hm & SNB_HEAD_SEM) ? (args.sem_sigmoid ? 1.0f / (1.0f + expf(-sv)) : sv) : 0.f;
                }
              }
            }
          }
        }
        // the next tile's bias still has to be staged (done inside the chunk loop otherwise)
        if (!nx.done) nbias = bias_fetch(nx);
        if (gtid < 128) gbias[((it + 1) & 1) * 128 + gtid] = nbias;
        gbar();
      } else {
#pragma unroll 1
      for (int ci = 0; ci < 2; ++ci) {
        const int ch = grp + 2 * ci;                 // 64-column chunk of the tile
        const uint32_t b = cn & 1;                   // staging buffer of this chunk
        const uint32_t buf0 = stg0 + b * GEMM_STAGING;
        // the chunk after this one (same tile, or the first chunk of the next tile)
        const bool nmul = (ci == 0) ? (epi == EPI_MUL) : next_mul;
        if (leader && (epi == EPI_MUL || nmul)) {
          // a mul-operand tile is TMA-loaded into the staging buffer its chunk will be multiplied in: the
          // store that last used that buffer must have drained first
          bulk_wait_read<0>();
          if (epi == EPI_MUL && !cur_prefetched) {
            mbar_expect_tx(&gmfull[b], GEMM_STAGING);
            tma_load_2d_hint(stg_ptr + b * GEMM_STAGING, &args.maps[c.l].tmMul, &gmfull[b], n0 + ch * 64, m_real, L2_EVICT_FIRST);
          }
          if (nmul) {
            const uint32_t nb = b ^ 1;
            mbar_expect_tx(&gmfull[nb], GEMM_STAGING);
            if (ci == 0) {
              tma_load_2d_hint(stg_ptr + nb * GEMM_STAGING, &args.maps[c.l].tmMul, &gmfull[nb], n0 + (ch + 2) * 64, m_real, L2_EVICT_FIRST);
            } else {
              const int nblk = (nx.g * CHAIN_SLOTS + nx.s) * n_pairs + pair;
              tma_load_2d_hint(stg_ptr + nb * GEMM_STAGING, &args.maps[nx.l].tmMul, &gmfull[nb], nx.j * 256 + grp * 64,
                               nblk * 256 + (int)cta_rank * GEMM_BLOCK_M, L2_EVICT_FIRST);
            }
          }
        }
        cur_prefetched = nmul;
        if (ci == 1 && !nx.done) nbias = bias_fetch(nx);   // in flight during this chunk
        const int colbase = n0 + ch * 64 + half * 32;
        uint32_t mw = 0;
        if (epi == EPI_MUL && siren && row_ok) mw = __ldg(mask + (size_t)(m_real + row) * mask_ld + (colbase >> 5));

        uint32_t v[32];
        tmem_ld32(taddr + ch * 64 + half * 32, v);
        tc_wait_ld();

        if (epi == EPI_MUL) {
          mbar_wait(&gmfull[b], (mpar >> b) & 1u);
          mpar ^= 1u << b;
          if (!siren) chunk_mul<false, true>(v, buf0, row_off, sw, half, w0, mw);
          else if (w0 == 1.0f) chunk_mul<true, true>(v, buf0, row_off, sw, half, w0, mw);
          else chunk_mul<true, false>(v, buf0, row_off, sw, half, w0, mw);
        } else {
          uint32_t outw[16];
          uint32_t mbits = 0;
          const uint32_t bsm = smem_u32(tbias + ci * 64 + half * 32);
          if (epi == EPI_LINEAR) chunk_math<CM_LINEAR>(v, bsm, w0, outw, mbits);
          else if (mask == nullptr) chunk_math<CM_SIN>(v, bsm, w0, outw, mbits);
          else chunk_math<CM_SIN_MASK>(v, bsm, w0, outw, mbits);
          if (epi == EPI_SIN && mask != nullptr && row_ok)
            const_cast<uint32_t*>(mask)[(size_t)(m_real + row) * mask_ld + (colbase >> 5)] = mbits;
          // results are in registers: the store that last used this buffer (two chunks ago) had the whole compute
          // phase above to drain; the previous chunk's store (other buffer) may still be in flight
          if (leader) bulk_wait_read<1>();
          gbar();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t off = row_off + ((((uint32_t)(half * 4 + g)) ^ sw) << 4);
            st_shared_v4(buf0 + off, outw[g * 4], outw[g * 4 + 1], outw[g * 4 + 2], outw[g * 4 + 3]);
          }
        }
        if (ci == 1 && gtid < 128) gbias[((it + 1) & 1) * 128 + gtid] = nbias;   // visible after the barrier below
        fence_proxy_async();
        gbar();
        if (leader) {
          confirm_pending();
          tma_store_2d(&args.maps[c.l].tmO0, buf0, n0 + ch * 64, m_out);
          bulk_commit();
        }
        ++cn;
      }
      }
      tc_fence_before();
      if (lead_cta) mbar_arrive(&tempty[acc]);
      else mbar_arrive_remote(&tempty[acc], 0);   // the leader CTA's MMA warp owns the accumulator hand-shake

      if (leader && c.j == ly.n_tiles - 1) {
        // this group's part of the tile-set (layer, slot) is committed; tell the A producer once it has landed.
        // Normally that is noticed before the next store (a tile of the other slot); when the very next tile-set
        // belongs to the same slot (single-slot tail) or nothing follows, wait here.
        if (npend == 0) { pend_cn0 = cn; pend_slot0 = c.s; }
        else { pend_cn1 = cn; pend_slot1 = c.s; }
        ++npend;
        // (a head-output tile-set issues no store of its own, so nothing later would notice it: release it now)
        if (nx.done || nx.s == c.s || epi == EPI_HEADOUT) {
          bulk_wait_done<0>();
          confirm(cn);
        }
      }
    }
    if (leader) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // neither CTA may exit while its peer can still signal it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ================================================================================================
// host side
// ================================================================================================
int chain_scratch_rows() {
  const int sms = num_sms();
  return sms <= 0 ? 0 : (sms / 2) * CHAIN_SLOTS * 256;
}

int chain_launch(const ChainArgs& a, cudaStream_t st) {
  SNB_CHECK_ARG(a.n_layers >= 1 && a.n_layers <= CHAIN_MAX_LAYERS, SNB_ERR_INVALID, "chain: %d layers", a.n_layers);
  SNB_CHECK_ARG(a.M >= 1 && a.n_blocks == (a.M + 255) / 256, SNB_ERR_INVALID, "chain: bad row count");
  double macs = 0.0;
  for (int l = 0; l < a.n_layers; ++l) {
    const ChainLayer& ly = a.layers[l];
    SNB_CHECK_ARG(ly.n_tiles >= 1 && ly.kb_total >= 1 && ly.nseg >= 1 && ly.nseg <= 3, SNB_ERR_INVALID, "chain: layer %d shape", l);
    SNB_CHECK_ARG(ly.epi == EPI_SIN || ly.epi == EPI_LINEAR || ly.epi == EPI_MUL || ly.epi == EPI_HEADOUT, SNB_ERR_UNSUPPORTED,
                  "chain: layer %d epilogue %d", l, ly.epi);
    if (ly.epi == EPI_HEADOUT)
      SNB_CHECK_ARG(ly.n_tiles == 1 && (ly.rows_mode == 2 ? a.out_packed != nullptr : ly.part != nullptr), SNB_ERR_INVALID,
                    "chain: head-output layer %d needs its destination", l);
    SNB_CHECK_ARG(!(ly.epi == EPI_MUL && ly.mul_siren) || (ly.mask != nullptr && ly.mask_ld > 0), SNB_ERR_INVALID,
                  "chain: layer %d needs the sign mask of the saved activation", l);
    macs += (double)a.n_blocks * 256.0 * (ly.epi == EPI_HEADOUT ? 16.0 : ly.n_tiles * 256.0) * ly.kb_total * GEMM_BLOCK_K;
  }
  const int sms = num_sms();
  if (sms <= 0) return SNB_ERR_NO_DEVICE;
  static bool attr_set = false;
  if (!attr_set) {
    SNB_CUDA(cudaFuncSetAttribute(snb_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHAIN_SMEM_BYTES));
    attr_set = true;
  }
  const int pairs = a.n_blocks < sms / 2 ? a.n_blocks : sms / 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(CHAIN_THREADS);
  cfg.dynamicSmemBytes = CHAIN_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const bool timed = profile_gemm_begin(st, macs, 100 + a.n_layers, a.M, 0, 0, 2, 1);
  SNB_CUDA(cudaLaunchKernelEx(&cfg, snb_chain_kernel, a));
  if (timed) profile_gemm_end(st);
  return launch_status("snb_chain_kernel");
}

}  // namespace snb
