// K2 chained MLP kernel - see k2_chain.cuh for the design.  sm_100a only (tcgen05 / TMEM / TMA).
//
// Roles (20 warps, one CTA per SM, clusters of 2 = SM pairs, cta_group::2 MMAs over 256 x 256 x 64):
//   warp 0      TMA producer for A (activations; waits for the producing layer of the same slot)
//   warp 1      tcgen05.mma issuer of the leader CTA (+ TMEM owner)
//   warp 2      TMA producer for B (weights)
//   warps 3-18  epilogue math: two groups of 8 warps.  Group g takes the 64-column chunks {g, g+2} of every
//               256-column tile; inside a group, warp w and warp w+4 read the same TMEM lane quadrant and
//               split the chunk's columns 0-31 / 32-63, so every SM sub-partition has 4 epilogue warps to
//               hide the TMEM-load / MUFU / shared-memory latencies behind each other.  They never wait for
//               each other: a chunk is handed over through mbarriers to
//   warp 19     the store warp: TMA-stores every staged chunk, re-arms the staging buffer once the store has
//               drained (for dgrad layers by TMA-loading the next multiplicand tile into it), and tells the
//               A producer when a layer's rows have landed.
#include "k2_chain.cuh"
#include "k2_ptx.cuh"

#include <stdlib.h>
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
  // Per pass p: this pair owns the row blocks pe, pe + n_pairs, ... (pe = the pair index rotated by the pass's shift).
  // The blocks are carried through the pass's layers in groups of up to CHAIN_SLOTS interleaved slots.  The groups are
  // BALANCED (28 blocks = 8 x 3 + 2 x 2, 4 blocks = 2 + 2, never ... + 1): a single-slot group has nothing to interleave
  // with, so every layer of it would wait for its own store -> load round trip through L2.
  // Groups are numbered through the passes: [0, g0) belong to pass 0, [g0, groups) to pass 1.
  int g0, groups;
  int base[2], rem[2], pe[2], l0[2], l1[2], n_pairs;
  __host__ __device__ __forceinline__ void init(const ChainArgs& a, int pair, int n_pairs_) {
    n_pairs = n_pairs_;
    int ng[2] = {0, 0};
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      base[p] = rem[p] = pe[p] = l0[p] = l1[p] = 0;
      if (p >= a.n_passes) continue;
      const ChainPass& ps = a.pass[p];
      pe[p] = (pair + n_pairs - ps.shift % n_pairs) % n_pairs;
      const int n_my = pe[p] < ps.n_blocks ? (ps.n_blocks - pe[p] + n_pairs - 1) / n_pairs : 0;
      ng[p] = (n_my + CHAIN_SLOTS - 1) / CHAIN_SLOTS;
      base[p] = ng[p] > 0 ? n_my / ng[p] : 0;
      rem[p] = ng[p] > 0 ? n_my - base[p] * ng[p] : 0;
      l0[p] = ps.layer0;
      l1[p] = ps.layer0 + ps.n_layers;
    }
    g0 = ng[0];
    groups = ng[0] + ng[1];
  }
  __host__ __device__ __forceinline__ int pass_of(int g) const { return g >= g0 ? 1 : 0; }
  // (selects, not indexed loads: the members stay in registers)
  __host__ __device__ __forceinline__ int nslots(int g) const {
    const bool p = g >= g0;
    const int gp = p ? g - g0 : g, b = p ? base[1] : base[0], r = p ? rem[1] : rem[0];
    return b + (gp < r ? 1 : 0);
  }
  // row block of slot s of group g
  __host__ __device__ __forceinline__ int block(int g, int s) const {
    const bool p = g >= g0;
    const int gp = p ? g - g0 : g, b = p ? base[1] : base[0], r = p ? rem[1] : rem[0];
    const int first = gp * b + (gp < r ? gp : r);   // first own block of the group
    return (first + s) * n_pairs + (p ? pe[1] : pe[0]);
  }
  __host__ __device__ __forceinline__ int layer_begin(int g) const { return g >= g0 ? l0[1] : l0[0]; }
  __host__ __device__ __forceinline__ int layer_end(int g) const { return g >= g0 ? l1[1] : l1[0]; }
};

__host__ __device__ __forceinline__ void cur_init(Cursor& c, const Seq& q) {
  c.g = c.s = c.j = 0;
  c.done = q.groups <= 0;
  c.l = c.done ? 0 : q.layer_begin(0);
}
__host__ __device__ __forceinline__ void cur_next(Cursor& c, const Seq& q, const ChainArgs& a) {
  if (++c.j < a.layers[c.l].n_tiles) return;
  c.j = 0;
  if (++c.s < q.nslots(c.g)) return;
  c.s = 0;
  if (++c.l < q.layer_end(c.g)) return;
  ++c.g;
  if (c.g >= q.groups) {
    c.done = true;
    c.l = 0;
  } else {
    c.l = q.layer_begin(c.g);
  }
}

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }


// ---- epilogue math, specialised per mode so the inner loops are branch-free ---------------------------------
enum ChunkMode { CM_LINEAR = 0, CM_SIN = 1, CM_SIN_MASK = 2, CM_RELU = 3 };

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
    if (MODE == CM_RELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] = fmaxf(y[j], 0.f);
    } else if (MODE != CM_LINEAR) {
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
// RELU: the multiplicand is the saved activation h = max(y, 0): derivative [h > 0]
template <bool SIREN, bool W0ONE, bool RELU = false>
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
      if (RELU) {
        f0 = f0 > 0.f ? 1.0f : 0.f;
        f1 = f1 > 0.f ? 1.0f : 0.f;
      }
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
  float* bias_smem = reinterpret_cast<float*>(sStg + GEMM_NUM_STAGING * GEMM_STAGING);   // [warp][tile parity][chunk][32]
  // [tile parity][128 rows][8 words]: the sign-mask words of a tile travel as ONE 4 KB TMA tile (full 32-byte sectors) instead
  // of one 4-byte access per thread and chunk, which cost a whole L2 sector operation each (ncu: a third of the write sectors)
  uint32_t* mask_smem = reinterpret_cast<uint32_t*>(bias_smem + 2048);
  uint64_t* fullA = reinterpret_cast<uint64_t*>(mask_smem + 2048);
  uint64_t* emptyA = fullA + 8;
  uint64_t* fullB = emptyA + 8;
  uint64_t* emptyB = fullB + 8;
  uint64_t* tfull = emptyB + 8;
  uint64_t* tempty = tfull + 2;
  uint64_t* rdy = tempty + 2;                 // [group][chunk]: staging buffer free / multiplicand tile landed
  uint64_t* stg = rdy + GEMM_NUM_STAGING;     // [group][chunk]: chunk staged by the group's 8 warps
  uint64_t* ready = stg + GEMM_NUM_STAGING;
  // sign-mask tile of a dgrad tile landed: one barrier per tile parity (= per mask buffer), one phase per chunked tile of
  // that parity - two barriers so that a tile's mask can be requested TWO tiles ahead without lapping a waiter
  uint64_t* mrdy = ready + CHAIN_SLOTS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mrdy + 2);

  const int warp = threadIdx.x >> 5;  // warp-uniform
  const int lane = threadIdx.x & 31;
  uint32_t cta_rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const bool lead_cta = (cta_rank == 0);
  const int pair = (int)(blockIdx.x >> 1);
  const int n_pairs = (int)(gridDim.x >> 1);
  Seq seq;
  seq.init(args, pair, n_pairs);
  if (args.prefetch != nullptr) {
    const char* base = reinterpret_cast<const char*>(args.prefetch);
    for (size_t off = ((size_t)blockIdx.x * CHAIN_THREADS + threadIdx.x) * 128; off < args.prefetch_bytes;
         off += (size_t)gridDim.x * CHAIN_THREADS * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
  }

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
    for (int s = 0; s < GEMM_NUM_STAGING; ++s) {
      mbar_init(&rdy[s], 1);
      mbar_init(&stg[s], 8);   // one arrival per warp of the group
    }
    for (int s = 0; s < CHAIN_SLOTS; ++s) mbar_init(&ready[s], 1);   // the store warp
    mbar_init(&mrdy[0], 1);
    mbar_init(&mrdy[1], 1);
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
    uint32_t owed = 0;   // bit s: the latest tile-set of slot s was stored through TMA and its phase of ready[s] is not consumed yet
    Cursor c;
    for (cur_init(c, seq); !c.done; cur_next(c, seq, args)) {
      const ChainLayer& ly = args.layers[c.l];
      if (c.j == 0) {
        if ((owed >> c.s) & 1u) {
          // the previous tile-set of this slot (the layer that produced this layer's input rows, or the last layer of the
          // slot's previous block) has been stored completely; one barrier phase per stored tile-set keeps producer and
          // store warp in lock step.  (A head-output tile-set stores nothing through TMA and feeds no later layer: no
          // phase for it.)
          mbar_wait(&ready[c.s], (rbits >> c.s) & 1u);
          rbits ^= 1u << c.s;
          fence_proxy_async_all();
        }
        if (ly.epi != EPI_HEADOUT) owed |= 1u << c.s;
        else owed &= ~(1u << c.s);
      }
      const int blk = seq.block(c.g, c.s);
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
#ifdef SNB_EXPERIMENTS
        if ((args.exp & 2) && (kb & 1)) {   // operand feed experiment: no load, the MMA reads whatever the stage holds
          if (elect_one()) {
            if (lead_cta) mbar_arrive(&fullA[stage]);
            else mbar_arrive_remote(&fullA[stage], 0);
          }
        } else
#endif
        if (elect_one()) {
          if (lead_cta) mbar_expect_tx(&fullA[stage], 2 * 16384);
          else mbar_arrive_remote(&fullA[stage], 0);
          // rows written by this pair one layer-step ago are kept in L2 (stores: evict-last) until their last read here
          tma_load_2d_2sm_hint(sA + stage * 16384, &args.maps[c.l].tmA[sg], &fullA[stage], kk * GEMM_BLOCK_K,
                               ly.a_scratch[sg] ? row_scr : row_real,
                               // (a head-output layer reads its rows BEFORE the wide layer that shares them: never the last use)
                               (c.j == ly.n_tiles - 1 && ly.epi != EPI_HEADOUT) ? L2_EVICT_FIRST : L2_EVICT_LAST);
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
#ifdef SNB_EXPERIMENTS
        if ((args.exp & 1) && (kb & 1)) {
          if (elect_one()) {
            if (lead_cta) mbar_arrive(&fullB[stage]);
            else mbar_arrive_remote(&fullB[stage], 0);
          }
        } else
#endif
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
        const int tail_k16 = args.layers[c.l].tail_k16;
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
            const int ksteps = kb == kb_total - 1 ? tail_k16 : GEMM_BLOCK_K / 16;
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
              if (k < ksteps) tc_mma_bf16_2sm(d_tmem, ad + k * 2, bd + k * 2, idesc, (kb > 0 || k > 0) ? 1u : 0u);
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
  } else if (warp == 3 + CHAIN_EPI_WARPS) {
    // ===================================== store warp =========================================
    // Staging buffer (group g, chunk ci of a tile) = sStg + (g * 2 + ci) * 16 KB; every chunked tile uses all four.
    uint32_t itc = 0;        // chunked tiles seen
    uint32_t cn = 0;         // stores committed
    int npend = 0, pend_slot0 = 0, pend_slot1 = 0;
    uint32_t pend_cn0 = 0, pend_cn1 = 0;
    auto confirm = [&](uint32_t completed) {   // elected lane only
      while (npend > 0 && pend_cn0 <= completed) {
        fence_proxy_async_all();
        mbar_arrive(&ready[pend_slot0]);
        pend_cn0 = pend_cn1;
        pend_slot0 = pend_slot1;
        --npend;
      }
    };
    // make buffer (g, ci) ready for tile t: dgrad layers get their multiplicand tile TMA-loaded into it
    auto arm = [&](const Cursor& t, int g, int ci, uint32_t par) {   // elected lane only; par = parity of the tile's index
      const ChainLayer& tl = args.layers[t.l];
      uint64_t* bar = &rdy[g * 2 + ci];
      const int trow = seq.block(t.g, t.s) * 256 + (int)cta_rank * GEMM_BLOCK_M;
#ifdef SNB_EXPERIMENTS
      if (tl.epi == EPI_MUL && (args.exp & 16)) {
        mbar_arrive(bar);
      } else
#endif
      if (tl.epi == EPI_MUL) {
        mbar_expect_tx(bar, GEMM_STAGING);
        tma_load_2d_hint(sStg + (g * 2 + ci) * GEMM_STAGING, &args.maps[t.l].tmMul, bar, t.j * 256 + (g + 2 * ci) * 64, trow,
                         L2_EVICT_FIRST);
      } else {
        mbar_arrive(bar);
      }
    };
    // the sign-mask words of tile t (the par-th chunked tile, parity par & 1) into mask buffer par & 1.  They come from HBM
    // (written by the forward pass) and are the first thing a dgrad tile's epilogue needs, so they are requested two tiles
    // ahead: ncu showed the 16 math warps waiting 12 % of their time for a request made half a tile ahead.
    auto arm_mask = [&](const Cursor& t, uint32_t par) {   // elected lane only
      const ChainLayer& tl = args.layers[t.l];
      uint64_t* bar = &mrdy[par & 1];
      if (tl.epi == EPI_MUL && tl.mul_siren == 1) {
        const int trow = seq.block(t.g, t.s) * 256 + (int)cta_rank * GEMM_BLOCK_M;
        mbar_expect_tx(bar, 4096);
        tma_load_2d_hint(mask_smem + (par & 1) * 1024, &args.maps[t.l].tmMask, bar, t.j * 8, trow, L2_EVICT_FIRST);
      } else {
        mbar_arrive(bar);
      }
    };
    const bool el = elect_one();
    Cursor c, nx;
    cur_init(c, seq);
    // skip to the first / next tile that stages chunks (head-output tiles write straight to global memory)
    auto next_chunked = [&](Cursor t) {
      do { cur_next(t, seq, args); } while (!t.done && args.layers[t.l].epi == EPI_HEADOUT);
      return t;
    };
    if (!c.done && el) {
      // NB: c may be a head-output tile only if the chain starts with one, which no plan does
      for (int k = 0; k < 4; ++k) arm(c, k >> 1, k & 1, 0);
      arm_mask(c, 0);
      const Cursor c1 = next_chunked(c);
      if (!c1.done) arm_mask(c1, 1);
    }
    int prev_g = -1, prev_ci = -1;   // the store issued before the current one: its buffer is re-armed once it has drained
    for (; !c.done;) {
      const ChainLayer& ly = args.layers[c.l];
      Cursor nx = c;
      cur_next(nx, seq, args);
      if (ly.epi != EPI_HEADOUT) {
        const Cursor nc = next_chunked(c);
        const Cursor nc2 = nc.done ? nc : next_chunked(nc);
        const int blk = seq.block(c.g, c.s);
        const int m_out = ly.o_scratch ? (pair * CHAIN_SLOTS + c.s) * 256 + (int)cta_rank * GEMM_BLOCK_M
                                       : blk * 256 + (int)cta_rank * GEMM_BLOCK_M;
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
          const int ci = k >> 1, g = k & 1;
          mbar_wait(&stg[g * 2 + ci], itc & 1);
          if (el) {
            if (npend > 0) {
              // a tile-set's rows must be in L2 / HBM before the A producer may read them back
              bulk_wait_done<0>();
              confirm(cn);
            }
#ifdef SNB_EXPERIMENTS
            if (!(args.exp & 8))
#endif
            tma_store_2d_hint(&args.maps[c.l].tmO0, smem_u32(sStg) + (g * 2 + ci) * GEMM_STAGING, c.j * 256 + (g + 2 * ci) * 64, m_out,
                              L2_EVICT_LAST);
            if (k == 3 && ly.epi == EPI_SIN && ly.mask != nullptr)   // all 16 warps have staged their words of the tile
              tma_store_2d(&args.maps[c.l].tmMask, smem_u32(mask_smem) + (itc & 1) * 4096, c.j * 8, blk * 256 + (int)cta_rank * GEMM_BLOCK_M);
            bulk_commit();
            ++cn;
            if (prev_g >= 0 && !nc.done) {
              // every store but the one just issued has drained: hand the previous chunk's buffer to the next tile
              bulk_wait_read<1>();
              arm(nc, prev_g, prev_ci, (itc + 1) & 1);
            }
            // both groups have staged their first chunk of this tile, so all 16 math warps hold its mask words in
            // registers: its mask buffer is free for the tile after next (forward tiles: only after the mask store below)
            if (k == 1 && !nc2.done && !(ly.epi == EPI_SIN && ly.mask != nullptr)) arm_mask(nc2, itc + 2);
            prev_g = g;
            prev_ci = ci;
          }
          __syncwarp();
        }
        // the buffer of the tile's last store: wait for its drain here rather than at the next tile's first store, so the
        // next tile's math never waits for it longer than the drain takes
        if (el && !nc.done) {
          bulk_wait_read<0>();
          arm(nc, 1, 1, (itc + 1) & 1);
        }
        if (el && !nc2.done && ly.epi == EPI_SIN && ly.mask != nullptr) arm_mask(nc2, itc + 2);   // (keeps the phases in step)
        prev_g = -1;
        ++itc;
      }
      if (el && c.j == ly.n_tiles - 1 && ly.epi != EPI_HEADOUT) {
        // the tile-set (layer, slot) is committed; tell the A producer once it has landed.  Normally that is noticed before
        // the next store; when the next tile-set is the same slot's (single-slot tail), nothing follows, or the next tiles
        // issue no store (head outputs: this warp does not take part in them, so nothing would notice in time), wait here.
        if (npend == 0) { pend_cn0 = cn; pend_slot0 = c.s; }
        else { pend_cn1 = cn; pend_slot1 = c.s; }
        ++npend;
        if (nx.done || nx.s == c.s || args.layers[nx.l].epi == EPI_HEADOUT) {
          bulk_wait_done<0>();
          confirm(cn);
        }
      }
      c = nx;
    }
    if (el) bulk_wait_all();
  } else {
    // ===================================== epilogue math (2 groups x 8 warps) =================
    const int ewarp = warp - 3;               // 0..15
    const int grp = ewarp >> 3;               // epilogue group
    const int half = (ewarp >> 2) & 1;        // which 32 columns of a 64-column chunk
    const int quad = warp & 3;                // TMEM lane quadrant this warp may read (warp id % 4)
    const int row = quad * 32 + lane;         // row of the 128-row tile == TMEM lane
    const uint32_t stg0 = smem_u32(sStg) + grp * 2 * GEMM_STAGING;   // this group's two staging buffers (chunk 0 / 1 of a tile)
    const uint32_t row_off = (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    float* wbias = bias_smem + ewarp * 128;   // [tile parity][chunk][32]: this warp's bias values (x w0)

    uint32_t it = 0, itc = 0;
    auto bias_fetch = [&](const Cursor& t, int ci) -> float {   // staged pre-multiplied by w0: y = fma(acc, w0, w0 * bias)
      const ChainLayer& tl = args.layers[t.l];
      return (tl.epi != EPI_HEADOUT && tl.bias != nullptr)
                 ? tl.w0 * __ldg(tl.bias + t.j * 256 + (grp + 2 * ci) * 64 + half * 32 + lane) : 0.f;
    };

    Cursor c, nx;
    cur_init(c, seq);
    nx = c;
    if (!nx.done) cur_next(nx, seq, args);
    if (!c.done) {
      wbias[lane] = bias_fetch(c, 0);
      wbias[32 + lane] = bias_fetch(c, 1);
      __syncwarp();
    }
    for (; !c.done; c = nx, cur_next(nx, seq, args), ++it) {
      const ChainLayer& ly = args.layers[c.l];
      const int epi = ly.epi;
      const float w0 = ly.w0;
      const uint32_t* mask = ly.mask;
      const int mask_ld = ly.mask_ld;
      const bool siren = ly.mul_siren == 1;
      const bool relu_bwd = ly.mul_siren == 2;
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const int blk = seq.block(c.g, c.s);
      const ChainPass& ps = args.pass[seq.pass_of(c.g)];
      const int m_real = blk * 256 + (int)cta_rank * GEMM_BLOCK_M;
      const int n0 = c.j * 256;
      const bool row_ok = m_real + row < ps.M;
      const uint32_t bsm = smem_u32(wbias + (it & 1) * 64);
      // in flight while this warp waits for the accumulator: the next tile's bias slice
      float nb0 = 0.f, nb1 = 0.f;
      if (!nx.done) { nb0 = bias_fetch(nx, 0); nb1 = bias_fetch(nx, 1); }

      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * 256;

      uint32_t* mtile = mask_smem + (itc & 1) * 1024 + row * 8 + grp * 2 + half;   // + 4 * ci: this thread's word of chunk ci
      uint32_t mw0 = 0, mw1 = 0;
      if (epi == EPI_MUL && siren) {
        mbar_wait(&mrdy[itc & 1], (itc >> 1) & 1);   // the tile's sign-mask words have landed
        mw0 = mtile[0];
        mw1 = mtile[4];
      }

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
              float* o = ps.out_packed + grow * args.n_out;
              const int hm = ps.head_mask;
              // rs_semantic.py:282-284: rgb = sigmoid(.) * (1 + 2*0.001) - 0.001
#pragma unroll
              for (int j = 0; j < 3; ++j) o[j] = (hm & SNB_HEAD_RGB) ? (1.0f / (1.0f + expf(-x[j]))) * 1.002f - 0.001f : 0.f;
              o[3] = (hm & SNB_HEAD_SIGMA) ? (x[3] > 20.f ? x[3] : log1pf(expf(x[3]))) : 0.f;
              o[4] = args.nerf ? 1.0f : ((hm & SNB_HEAD_SUN) ? 1.0f / (1.0f + expf(-x[4])) : 0.f);
              if ((hm & SNB_HEAD_SKY) && ps.sky != nullptr) {
                const long long ray = args.rows_per_ray > 0 ? grow / args.rows_per_ray : grow;
                o[5] = __ldg(ps.sky + ray * 3);
                o[6] = __ldg(ps.sky + ray * 3 + 1);
                o[7] = __ldg(ps.sky + ray * 3 + 2);
              } else {
                o[5] = o[6] = o[7] = 0.f;
              }
              o[8] = (hm & SNB_HEAD_BETA) ? (x[5] > 20.f ? x[5] : log1pf(expf(x[5]))) : 0.f;
              const int bs = args.beta_s;
              if (bs) o[9] = (hm & SNB_HEAD_BETA) ? (x[6] > 20.f ? x[6] : log1pf(expf(x[6]))) : 0.f;
#pragma unroll
              for (int cc = 0; cc < 10; ++cc) {
                if (cc < args.n_classes) {
                  const float sv = bs ? x[(7 + cc) & 15] : x[6 + cc];   // (n_classes <= 9 with the extra head: 7 + cc <= 15)
                  o[9 + bs + cc] = (hm & SNB_HEAD_SEM) ? (args.sem_sigmoid ? 1.0f / (1.0f + expf(-sv)) : sv) : 0.f;
                }
              }
            }
          }
        }
      } else {
#pragma unroll 1
        for (int ci = 0; ci < 2; ++ci) {
          const int ch = grp + 2 * ci;                 // 64-column chunk of the tile
          const uint32_t buf0 = stg0 + ci * GEMM_STAGING;
          uint64_t* brdy = &rdy[grp * 2 + ci];
          const int colbase = n0 + ch * 64 + half * 32;
          const uint32_t mw = ci ? mw1 : mw0;

          uint32_t v[32];
          tmem_ld32(taddr + ch * 64 + half * 32, v);
          tc_wait_ld();

#ifdef SNB_EXPERIMENTS
          if (args.exp & 4) {
            mbar_wait(brdy, itc & 1);
          } else
#endif
          if (epi == EPI_MUL) {
            mbar_wait(brdy, itc & 1);   // the multiplicand tile has landed in the staging buffer
            if (relu_bwd) chunk_mul<false, true, true>(v, buf0, row_off, sw, half, w0, mw);
            else if (!siren) chunk_mul<false, true>(v, buf0, row_off, sw, half, w0, mw);
            else if (w0 == 1.0f) chunk_mul<true, true>(v, buf0, row_off, sw, half, w0, mw);
            else chunk_mul<true, false>(v, buf0, row_off, sw, half, w0, mw);
          } else {
            uint32_t outw[16];
            uint32_t mbits = 0;
            const uint32_t bs = bsm + ci * 128;
            if (epi == EPI_LINEAR) chunk_math<CM_LINEAR>(v, bs, w0, outw, mbits);
            else if (ly.relu) chunk_math<CM_RELU>(v, bs, w0, outw, mbits);
            else if (mask == nullptr) chunk_math<CM_SIN>(v, bs, w0, outw, mbits);
            else chunk_math<CM_SIN_MASK>(v, bs, w0, outw, mbits);
            mbar_wait(brdy, itc & 1);   // the store that last used this buffer (previous tile) has drained
            if (epi == EPI_SIN && mask != nullptr) mtile[4 * ci] = mbits;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t off = row_off + ((((uint32_t)(half * 4 + g)) ^ sw) << 4);
              st_shared_v4(buf0 + off, outw[g * 4], outw[g * 4 + 1], outw[g * 4 + 2], outw[g * 4 + 3]);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&stg[grp * 2 + ci]);   // 8 arrivals = chunk staged -> store warp
        }
        ++itc;
      }
      tc_fence_before();
      if (lead_cta) mbar_arrive(&tempty[acc]);
      else mbar_arrive_remote(&tempty[acc], 0);   // the leader CTA's MMA warp owns the accumulator hand-shake
      wbias[((it + 1) & 1) * 64 + lane] = nb0;
      wbias[((it + 1) & 1) * 64 + 32 + lane] = nb1;
      __syncwarp();
    }
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
  SNB_CHECK_ARG(a.n_passes == 1 || a.n_passes == 2, SNB_ERR_INVALID, "chain: %d passes", a.n_passes);
  int covered = 0, max_blocks = 0;
  long long rows = 0;
  for (int p = 0; p < a.n_passes; ++p) {
    const ChainPass& ps = a.pass[p];
    SNB_CHECK_ARG(ps.M >= 1 && ps.n_blocks == (ps.M + 255) / 256 && ps.shift >= 0, SNB_ERR_INVALID, "chain: bad row count");
    SNB_CHECK_ARG(ps.layer0 == covered && ps.n_layers >= 1, SNB_ERR_INVALID, "chain: pass %d layer range", p);
    covered += ps.n_layers;
    max_blocks = ps.n_blocks > max_blocks ? ps.n_blocks : max_blocks;
    rows += ps.M;
  }
  SNB_CHECK_ARG(covered == a.n_layers, SNB_ERR_INVALID, "chain: the passes cover %d of %d layers", covered, a.n_layers);
  double macs = 0.0;
  for (int l = 0; l < a.n_layers; ++l) {
    const ChainLayer& ly = a.layers[l];
    const ChainPass& ps = a.pass[(a.n_passes == 2 && l >= a.pass[1].layer0) ? 1 : 0];
    SNB_CHECK_ARG(ly.n_tiles >= 1 && ly.kb_total >= 1 && ly.nseg >= 1 && ly.nseg <= 3, SNB_ERR_INVALID, "chain: layer %d shape", l);
    SNB_CHECK_ARG(ly.epi == EPI_SIN || ly.epi == EPI_LINEAR || ly.epi == EPI_MUL || ly.epi == EPI_HEADOUT, SNB_ERR_UNSUPPORTED,
                  "chain: layer %d epilogue %d", l, ly.epi);
    if (ly.epi == EPI_HEADOUT)
      SNB_CHECK_ARG(ly.n_tiles == 1 && (ly.rows_mode == 2 ? ps.out_packed != nullptr : ly.part != nullptr), SNB_ERR_INVALID,
                    "chain: head-output layer %d needs its destination", l);
    SNB_CHECK_ARG(!(ly.epi == EPI_MUL && ly.mul_siren == 1) || (ly.mask != nullptr && ly.mask_ld > 0), SNB_ERR_INVALID,
                  "chain: layer %d needs the sign mask of the saved activation", l);
    SNB_CHECK_ARG(ly.tail_k16 >= 1 && ly.tail_k16 <= 4, SNB_ERR_INVALID, "chain: layer %d tail_k16 %d", l, ly.tail_k16);
    macs += (double)ps.n_blocks * 256.0 * (ly.epi == EPI_HEADOUT ? 16.0 : ly.n_tiles * 256.0) *
            ((ly.kb_total - 1) * GEMM_BLOCK_K + ly.tail_k16 * 16);
  }
  const int sms = num_sms();
  if (sms <= 0) return SNB_ERR_NO_DEVICE;
  static bool attr_set = false;
  if (!attr_set) {
    SNB_CUDA(cudaFuncSetAttribute(snb_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHAIN_SMEM_BYTES));
    attr_set = true;
  }
  const int pairs = max_blocks < sms / 2 ? max_blocks : sms / 2;
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
  const bool timed = profile_gemm_begin(st, macs, 100 + a.n_layers, (int)rows, 0, 0, 2, 1);
#ifdef SNB_EXPERIMENTS
  {
    ChainArgs b = a;
    const char* e = getenv("SNB_EXP");
    b.exp = e ? atoi(e) : 0;
    SNB_CUDA(cudaLaunchKernelEx(&cfg, snb_chain_kernel, b));
  }
#else
  SNB_CUDA(cudaLaunchKernelEx(&cfg, snb_chain_kernel, a));
#endif
  if (timed) profile_gemm_end(st);
  return launch_status("snb_chain_kernel");
}

}  // namespace snb

// test hook (runs on the host, no device needed): the row blocks SM pair `pair` of `n_pairs` carries through a launch of one or
// two passes, in execution order, as (pass << 24 | block) words - the same Seq / Cursor code the kernel runs.
extern "C" int snb_chain_schedule(int n_blocks0, int n_blocks1, int shift1, int n_pairs, int pair, int* out, int cap) {
  using namespace snb;
  if (n_blocks0 < 1 || n_blocks1 < 0 || n_pairs < 1 || pair < 0 || pair >= n_pairs || !out) return -1;
  ChainArgs* a = new ChainArgs();
  memset(a, 0, sizeof(*a));
  a->n_passes = n_blocks1 > 0 ? 2 : 1;
  a->pass[0].n_blocks = n_blocks0;
  a->pass[0].layer0 = 0;
  a->pass[0].n_layers = 1;
  a->pass[1].n_blocks = n_blocks1;
  a->pass[1].layer0 = 1;
  a->pass[1].n_layers = 1;
  a->pass[1].shift = shift1;
  a->n_layers = a->n_passes;
  a->layers[0].n_tiles = a->layers[1].n_tiles = 1;
  Seq q;
  q.init(*a, pair, n_pairs);
  int n = 0;
  Cursor c;
  for (cur_init(c, q); !c.done; cur_next(c, q, *a)) {
    if (n < cap) out[n] = (q.pass_of(c.g) << 24) | q.block(c.g, c.s);
    ++n;
  }
  delete a;
  return n;
}
