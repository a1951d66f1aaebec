// K2 GEMM kernel - see k2_gemm.cuh for the design.  sm_100a only (tcgen05 / TMEM / TMA).
#include "k2_chain.cuh"
#include "k2_gemm.cuh"
#include "k2_ptx.cuh"

#include <stdlib.h>

#include <mutex>

namespace snb {

bool profile_gemm_begin(cudaStream_t st, double macs, int epi, int M, int N, int K, int cg, int splits);
void profile_gemm_end(cudaStream_t st);

// ================================================================================================
// the kernel
// ================================================================================================
struct TileCoord {
  int m_blk, n_blk, kb0, kb1;
};
__device__ __forceinline__ TileCoord decode_tile(const GemmArgs& a, int tile) {
  TileCoord t;
  t.n_blk = tile % a.n_tiles;
  int r = tile / a.n_tiles;
  t.m_blk = r % a.m_tiles;
  int split = r / a.m_tiles;
  t.kb0 = (int)(((long long)a.kb_total * split) / a.splits);
  t.kb1 = (int)(((long long)a.kb_total * (split + 1)) / a.splits);
  return t;
}

// CG = 1: one CTA per 128 x block_n tile.  CG = 2: an SM pair (cluster of 2) per 256 x block_n tile -
// each CTA holds its 128 rows of A and HALF of the B tile, the leader CTA issues cta_group::2 MMAs.
template <int EPI, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1) snb_gemm_kernel(const __grid_constant__ GemmArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  if ((smem - smem_raw) + GEMM_LAYOUT_BYTES > GEMM_SMEM_BYTES) {
    if (threadIdx.x == 0) printf("snb gemm: dynamic shared memory base misaligned by %d bytes\n", (int)(smem - smem_raw));
    __trap();
  }
  uint8_t* sA = smem;
  uint8_t* sB = sA + args.a_stages * GEMM_A_STAGE;
  uint8_t* sStg = smem + GEMM_OPERAND_BYTES;
  uint8_t* sOnes = sStg + GEMM_NUM_STAGING * GEMM_STAGING;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(sOnes + GEMM_ONES_BYTES);
  uint64_t* emptyA = fullA + GEMM_MAX_RING;
  uint64_t* fullB = emptyA + GEMM_MAX_RING;
  uint64_t* emptyB = fullB + GEMM_MAX_RING;
  uint64_t* tfull = emptyB + GEMM_MAX_RING;
  uint64_t* tempty = tfull + 2;
  uint64_t* mfull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mfull + GEMM_NUM_STAGING);
  const uint32_t a_stages = (uint32_t)args.a_stages, b_stages = (uint32_t)args.b_stages;

  const int warp = threadIdx.x >> 5;  // warp-uniform
  const int lane = threadIdx.x & 31;
  uint32_t cta_rank = 0;
  if constexpr (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const bool lead_cta = (cta_rank == 0);
  const int unit = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;    // tile-scheduling unit: CTA or CTA pair
  const int n_units = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const uint32_t b_slot = args.b_slot;

  if (threadIdx.x == 0) {
    for (int s = 0; s < args.nseg; ++s) tma_prefetch_desc(&args.tmA[s]);
    tma_prefetch_desc(&args.tmB);
    if (EPI == EPI_WGRAD && args.side_n > 0) tma_prefetch_desc(&args.tmB2);
    for (int s = 0; s < GEMM_MAX_RING; ++s) {
      mbar_init(&fullA[s], CG);   // CG == 2: the leader's expect_tx arrive + the peer's remote arrive
      mbar_init(&emptyA[s], 1);
      mbar_init(&fullB[s], CG);
      mbar_init(&emptyB[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], CG * GEMM_EPI_THREADS);
    }
    for (int s = 0; s < GEMM_NUM_STAGING; ++s) mbar_init(&mfull[s], 1);
    fence_barrier_init();
  }
  if constexpr (EPI == EPI_WGRAD) {
    // the all-ones operand of the bias-gradient MMA (read by the tensor core: async proxy)
    for (int i = threadIdx.x; i < GEMM_ONES_BYTES / 4; i += GEMM_THREADS) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
    fence_proxy_async();
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();   // peer barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total_tiles = args.m_tiles * args.n_tiles * args.splits;

  if (warp == 0) {
    // ===================================== TMA producer: operand A ===========================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = unit; tile < total_tiles; tile += n_units) {
        const TileCoord t = decode_tile(args, tile);
        const int m_row = t.m_blk * (GEMM_BLOCK_M * CG) + (int)cta_rank * GEMM_BLOCK_M;
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(&emptyA[stage], phase ^ 1);
          if (lead_cta) mbar_expect_tx(&fullA[stage], args.a_bytes * CG);
          else mbar_arrive_remote(&fullA[stage], 0);
          uint8_t* a_dst = sA + stage * GEMM_A_STAGE;
          if (!args.a_mn) {
            int s = 0, kk = kb;
            while (s + 1 < args.nseg && kk >= args.seg_kb[s]) {
              kk -= args.seg_kb[s];
              ++s;
            }
            tma_load_op<CG>(a_dst, &args.tmA[s], &fullA[stage], kk * GEMM_BLOCK_K, m_row);
          } else {
            // [64 samples x 64 features] boxes; feature atoms 8 KB apart (UMMA LBO)
            tma_load_op<CG>(a_dst, &args.tmA[0], &fullA[stage], m_row, kb * GEMM_BLOCK_K);
            tma_load_op<CG>(a_dst + 8192, &args.tmA[0], &fullA[stage], m_row + 64, kb * GEMM_BLOCK_K);
          }
          if (++stage == a_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ===================================== TMA producer: operand B ===========================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const int n_half = args.block_n / CG;   // B rows this CTA holds
      const int nb_boxes = (n_half + 63) / 64;
      for (int tile = unit; tile < total_tiles; tile += n_units) {
        const TileCoord t = decode_tile(args, tile);
        const int n_row = t.n_blk * args.block_n + (int)cta_rank * n_half;
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(&emptyB[stage], phase ^ 1);
          // side operand: only on the k-blocks whose side MMAs this n-tile issues (see the MMA warp)
          const bool side_kb = EPI == EPI_WGRAD && CG == 2 && args.side_n > 0 && (kb % args.n_tiles) == t.n_blk;
          const uint32_t side_bytes = side_kb ? (uint32_t)(args.side_n / CG) * (GEMM_BLOCK_K * 2) : 0u;
          if (lead_cta) mbar_expect_tx(&fullB[stage], (args.b_bytes + side_bytes) * CG);
          else mbar_arrive_remote(&fullB[stage], 0);
          if (side_kb)
            tma_load_op<CG>(sStg + stage * GEMM_SIDE_STAGE, &args.tmB2, &fullB[stage], kb * GEMM_BLOCK_K,
                            (int)cta_rank * (args.side_n / CG));
          uint8_t* b_dst = sB + stage * b_slot;
          if (!args.b_mn) {
            tma_load_op<CG>(b_dst, &args.tmB, &fullB[stage], kb * GEMM_BLOCK_K, n_row);
          } else {
            for (int j = 0; j < nb_boxes; ++j)
              tma_load_op<CG>(b_dst + j * 8192, &args.tmB, &fullB[stage], n_row + j * 64, kb * GEMM_BLOCK_K);
          }
          if (++stage == b_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ========================================
    // The whole warp runs the loop (warp-uniform control flow and operands, so the descriptors live in
    // uniform registers); one elected lane issues the tcgen05 instructions.
    if (lead_cta) {
      // cute::UMMA::InstrDescriptor: c_format f32 (1<<4), a/b format bf16 (1<<7, 1<<10),
      // a_major bit 15, b_major bit 16, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)args.a_mn << 15) |
                             ((uint32_t)args.b_mn << 16) | ((uint32_t)(args.block_n >> 3) << 17) |
                             ((uint32_t)((GEMM_BLOCK_M * CG) >> 4) << 24);
      // K-major SW128: 8-row groups 1024 B apart (SBO); LBO unused (1).  MN-major SW128: 8-k groups
      // 1024 B apart (SBO), 64-element MN atoms 8192 B apart (LBO).
      const uint32_t a_lbo = args.a_mn ? 8192u : 16u, b_lbo = args.b_mn ? 8192u : 16u;
      const uint32_t a_kstep = (args.a_mn ? 2048u : 32u) >> 4, b_kstep = (args.b_mn ? 2048u : 32u) >> 4;
      // descriptor = constant high part | (smem address >> 4): only the low word changes per k-step
      const uint64_t adesc0 = umma_desc(0, a_lbo, 1024u), bdesc0 = umma_desc(0, b_lbo, 1024u);
      const uint64_t bdesc_k0 = umma_desc(0, 16u, 1024u);   // K-major tile (side operand)
      const uint32_t a_base = smem_u32(sA) >> 4, b_base = smem_u32(sB) >> 4;
      const uint32_t b_slot16 = b_slot >> 4;
      // bias gradient: D2[M, 16] += A[M, k] * ones[16, k]^T into the last 16 TMEM columns (every unit runs at most one
      // tile when colsum is set, so the second accumulator's columns are free)
      const uint32_t idesc_cs = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)args.a_mn << 15) | ((16u >> 3) << 17) |
                                ((uint32_t)((GEMM_BLOCK_M * CG) >> 4) << 24);
      const uint64_t odesc = umma_desc(smem_u32(sOnes), 16u, 1024u);
      const uint32_t d_cs = tmem_base + 2 * GEMM_MAX_BLOCK_N - 16;
      // side operand (K-major [side_n / CG rows, 64 k] per CTA in the staging area): D3[M, side_n] in the columns below d_cs
      const uint32_t idesc_side = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)args.a_mn << 15) |
                                  ((uint32_t)(args.side_n >> 3) << 17) | ((uint32_t)((GEMM_BLOCK_M * CG) >> 4) << 24);
      const uint32_t side_base = smem_u32(sStg) >> 4;
      const uint32_t d_side = tmem_base + 2 * GEMM_MAX_BLOCK_N - 16 - 64;
      const bool has_side = EPI == EPI_WGRAD && CG == 2 && args.side_n > 0;
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
      for (int tile = unit; tile < total_tiles; tile += n_units, ++it) {
        const TileCoord t = decode_tile(args, tile);
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * GEMM_MAX_BLOCK_N;
        // the n-tiles of an m-block share the column-sum work: k-block kb belongs to n-tile kb % n_tiles (every tile of a
        // split owns the same k range, so together they cover it exactly once) - one tile doing all of it ran ~10 % longer
        const bool cs_tile = EPI == EPI_WGRAD && args.colsum != nullptr;
        bool cs_started = false;
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(&fullA[sa], pa);
          mbar_wait(&fullB[sb], pb);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = adesc0 + (uint64_t)(a_base + sa * (GEMM_A_STAGE >> 4));
            const uint64_t bd = bdesc0 + (uint64_t)(b_base + sb * b_slot16);
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
              if constexpr (CG == 2) tc_mma_bf16_2sm(d_tmem, ad + k * a_kstep, bd + k * b_kstep, idesc, (kb > t.kb0 || k > 0) ? 1u : 0u);
              else tc_mma_bf16(d_tmem, ad + k * a_kstep, bd + k * b_kstep, idesc, (kb > t.kb0 || k > 0) ? 1u : 0u);
            }
            if (cs_tile && (kb % args.n_tiles) == t.n_blk) {
#pragma unroll
              for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
                if constexpr (CG == 2) tc_mma_bf16_2sm(d_cs, ad + k * a_kstep, odesc + k * 2, idesc_cs, (cs_started || k > 0) ? 1u : 0u);
                else tc_mma_bf16(d_cs, ad + k * a_kstep, odesc + k * 2, idesc_cs, (cs_started || k > 0) ? 1u : 0u);
              }
            }
            if constexpr (CG == 2) {
              if (has_side && (kb % args.n_tiles) == t.n_blk) {
                const uint64_t sd = bdesc_k0 + (uint64_t)(side_base + sb * (GEMM_SIDE_STAGE >> 4));
#pragma unroll
                for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
                  tc_mma_bf16_2sm(d_side, ad + k * a_kstep, sd + k * 2, idesc_side, (cs_started || k > 0) ? 1u : 0u);
              }
            }
            // frees the smem slots (in both CTAs of a pair) once these MMAs have read them
            if constexpr (CG == 2) { tc_commit_2sm(&emptyA[sa]); tc_commit_2sm(&emptyB[sb]); }
            else { tc_commit(&emptyA[sa]); tc_commit(&emptyB[sb]); }
          }
          __syncwarp();
          if ((cs_tile || has_side) && (kb % args.n_tiles) == t.n_blk) cs_started = true;
          if (++sa == a_stages) { sa = 0; pa ^= 1; }
          if (++sb == b_stages) { sb = 0; pb ^= 1; }
        }
        if (elect_one()) {
          if constexpr (CG == 2) tc_commit_2sm(&tfull[acc]);   // accumulator complete -> both CTAs' epilogues
          else tc_commit(&tfull[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================== epilogue (2 x 128 threads) ========================
    // Two warpgroups share every tile: group g takes the column chunks c with (c & 1) == g, so each
    // SM sub-partition has two epilogue warps to hide the TMEM-load / MUFU latencies behind each other.
    const int ewarp = warp - 3;             // 0..7
    const int grp = ewarp >> 2;             // epilogue group
    const int quad = warp & 3;              // TMEM lane quadrant this warp may read (warp id % 4)
    const int row = quad * 32 + lane;       // row of the 128-row tile == TMEM lane
    const int gtid = (ewarp & 3) * 32 + lane;  // 0..127 within the group
    const bool leader = (gtid == 0);
    const uint32_t stg0 = smem_u32(sStg) + grp * 2 * GEMM_STAGING;   // this group's two staging buffers
    const uint32_t row_off = (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    const int bar_id = 1 + grp;
    auto gbar = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
    uint32_t it = 0;
    uint32_t cn = 0;  // this group's running chunk counter (position in its staging ring)

    for (int tile = unit; tile < total_tiles; tile += n_units, ++it) {
      const TileCoord t = decode_tile(args, tile);
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * GEMM_MAX_BLOCK_N;
      const int m0 = t.m_blk * (GEMM_BLOCK_M * CG) + (int)cta_rank * GEMM_BLOCK_M;
      const int n0 = t.n_blk * args.block_n;

      if constexpr (EPI == EPI_F32ROWS) {
        if (grp == 0) {
          uint32_t v[16];
          tmem_ld16(taddr, v);
          tc_wait_ld();
          const long long grow = (long long)m0 + row;
          if (grow < args.M) {
            float4* dst = reinterpret_cast<float4*>(args.f32out + grow * args.ldo);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                   __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          }
        }
      } else {  // EPI_WGRAD
        if (args.block_n >= 32 && (args.block_n & 31) == 0 && args.f32out == nullptr) {
          // fp32 tile -> 32-column swizzled chunks -> TMA reduce-add (split-K accumulation in L2)
          const int chunks = args.block_n / 32;
          for (int c = grp; c < chunks; c += 2, ++cn) {
            const uint32_t buf0 = stg0 + (cn & 1) * GEMM_STAGING;
            uint32_t v[32];
            tmem_ld32(taddr + c * 32, v);
            tc_wait_ld();
            if (leader) bulk_wait_read<1>();
            gbar();
#pragma unroll
            for (int g = 0; g < 8; ++g)
              st_shared_v4(buf0 + row_off + ((((uint32_t)g) ^ sw) << 4), v[4 * g], v[4 * g + 1], v[4 * g + 2],
                           v[4 * g + 3]);
            fence_proxy_async();
            gbar();
            if (leader) {
              tma_reduce_add_2d(&args.tmO0, buf0, n0 + c * 32, m0);
              bulk_commit();
            }
          }
        } else {
          // narrow outputs (bias / per-ray columns, N = 16 or 64): red.global.add.f32
          const long long grow = (long long)m0 + row;
          for (int c0 = grp * 16; c0 < args.block_n; c0 += 32) {
            uint32_t v[16];
            tmem_ld16(taddr + c0, v);
            tc_wait_ld();
            if (grow < args.M) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int col = n0 + c0 + j;
                if (col < args.N) atomicAdd(args.f32out + grow * args.ldo + col, __uint_as_float(v[j]));
              }
            }
          }
        }
      }
      if constexpr (EPI == EPI_WGRAD) {
        // (this tile issued column-sum MMAs iff its k range holds a k-block with kb % n_tiles == n_blk)
        if (args.colsum != nullptr && grp == 0 &&
            t.kb0 + ((t.n_blk - t.kb0 % args.n_tiles + args.n_tiles) % args.n_tiles) < t.kb1) {
          // bias gradient of this tile's 128 rows over its share of the k-blocks: column 0 of the all-ones product
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + 2 * GEMM_MAX_BLOCK_N - 16, v);
          tc_wait_ld();
          if (m0 + row < args.M) atomicAdd(args.colsum + m0 + row, __uint_as_float(v[0]));
        }
        if (CG == 2 && args.side_n > 0 && grp == 0 &&
            t.kb0 + ((t.n_blk - t.kb0 % args.n_tiles + args.n_tiles) % args.n_tiles) < t.kb1) {
          // this tile's share of side_out[M, side_n] = A^T x X2 (fp32 partial sums over its k-blocks)
          const long long grow = (long long)m0 + row;
          for (int c0 = 0; c0 < args.side_n; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + 2 * GEMM_MAX_BLOCK_N - 16 - 64 + c0, v);
            tc_wait_ld();
            if (grow < args.M) {
#pragma unroll
              for (int j = 0; j < 16; ++j) atomicAdd(args.side_out + grow * args.ld_side + c0 + j, __uint_as_float(v[j]));
            }
          }
        }
      }
      tc_fence_before();
      if (lead_cta) mbar_arrive(&tempty[acc]);
      else mbar_arrive_remote(&tempty[acc], 0);   // the leader CTA's MMA warp owns the accumulator hand-shake
    }
    if (leader) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();   // neither CTA may exit while its peer can still signal it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ================================================================================================
// host side
// ================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    // no link-time dependency on libcuda: resolved through the runtime
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d(CUtensorMap* map, const void* ptr, int elem_bytes, uint64_t inner, uint64_t outer,
                 uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  SNB_CHECK_ARG(fn != nullptr, SNB_ERR_NO_DEVICE, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  SNB_CHECK_ARG(ptr != nullptr && ((uintptr_t)ptr & 15) == 0, SNB_ERR_INVALID, "tensor map: base %p not 16B aligned", ptr);
  SNB_CHECK_ARG((row_stride_bytes & 15) == 0, SNB_ERR_INVALID, "tensor map: row stride %llu not a multiple of 16",
                (unsigned long long)row_stride_bytes);
  SNB_CHECK_ARG(box_inner * elem_bytes <= 128 && box_outer <= 256 && box_outer >= 1, SNB_ERR_INVALID,
                "tensor map: box %u x %u unsupported", box_inner, box_outer);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SNB_CHECK_ARG(r == CUDA_SUCCESS, SNB_ERR_INVALID,
                "cuTensorMapEncodeTiled failed (%d): inner %llu outer %llu stride %llu box %ux%u", (int)r,
                (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_inner,
                box_outer);
  return 0;
}

int make_tmap_mask(CUtensorMap* map, const void* ptr, uint64_t words, uint64_t rows, uint64_t row_stride_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  SNB_CHECK_ARG(fn != nullptr, SNB_ERR_NO_DEVICE, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  SNB_CHECK_ARG(ptr != nullptr && ((uintptr_t)ptr & 15) == 0 && (row_stride_bytes & 15) == 0 && words % 8 == 0, SNB_ERR_INVALID,
                "mask tensor map: base %p / stride %llu / %llu words not 16-byte granular", ptr,
                (unsigned long long)row_stride_bytes, (unsigned long long)words);
  cuuint64_t dims[2] = {words, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {8, GEMM_BLOCK_M};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SNB_CHECK_ARG(r == CUDA_SUCCESS, SNB_ERR_INVALID, "cuTensorMapEncodeTiled (mask) failed (%d): %llu words x %llu rows, stride %llu",
                (int)r, (unsigned long long)words, (unsigned long long)rows, (unsigned long long)row_stride_bytes);
  return 0;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

int gemm_pick_cta_group(int epi, long long M, int N, int block_n) {
  static const int allow = env_int("SNB_CG", 2);
  if (allow < 2) return 1;
  // SM pairs for the big tiles: 256 x 256, each CTA ingests 32 KB per k-block instead of 48 KB
  if (block_n == 256 && N % 256 == 0 && M >= 256 && epi != EPI_HEADOUT && epi != EPI_F32ROWS) return 2;
  return 1;
}

void gemm_finalize(GemmArgs& a) {
  if (a.cta_group != 2) a.cta_group = 1;
  const int cg = a.cta_group;
  a.m_tiles = (a.M + GEMM_BLOCK_M * cg - 1) / (GEMM_BLOCK_M * cg);
  a.n_tiles = (a.N + a.block_n - 1) / a.block_n;
  if (a.splits < 1) a.splits = 1;
  if (a.splits > a.kb_total) a.splits = a.kb_total;
  const int n_half = a.block_n / cg;
  a.a_bytes = GEMM_A_STAGE;
  a.b_bytes = a.b_mn ? (unsigned)(((n_half + 63) / 64) * 8192) : (unsigned)(n_half * GEMM_BLOCK_K * 2);
  a.b_slot = cg == 2 ? GEMM_B_STAGE / 2 : GEMM_B_STAGE;
  if (a.a_stages <= 0 || a.b_stages <= 0) {
    static int ov[6] = {0, 0, 0, 0, 0, 0};
    static bool parsed = false;
    if (!parsed) {
      parsed = true;  // tuning override: SNB_RINGS="aK,bK,aMN,bMN,a2,b2"
      if (const char* e = getenv("SNB_RINGS")) sscanf(e, "%d,%d,%d,%d,%d,%d", &ov[0], &ov[1], &ov[2], &ov[3], &ov[4], &ov[5]);
    }
    if (cg == 2) { a.a_stages = 5; a.b_stages = 5; }          // 16 KB slots each: 160 KB
    else if (!a.a_mn) { a.a_stages = 4; a.b_stages = 3; }
    else { a.a_stages = 4; a.b_stages = 3; }
    if (cg == 2 && ov[4] > 0 && ov[5] > 0) { a.a_stages = ov[4]; a.b_stages = ov[5]; }
    if (cg == 1 && !a.a_mn && ov[0] > 0 && ov[1] > 0) { a.a_stages = ov[0]; a.b_stages = ov[1]; }
    if (cg == 1 && a.a_mn && ov[2] > 0 && ov[3] > 0) { a.a_stages = ov[2]; a.b_stages = ov[3]; }
  }
}

template <int EPI, int CG>
static int launch_epi(const GemmArgs& a, int units, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    SNB_CUDA(cudaFuncSetAttribute(snb_gemm_kernel<EPI, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    attr_set = true;
  }
  if (CG == 1) {
    snb_gemm_kernel<EPI, CG><<<units, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(a);
  } else {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * units);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = GEMM_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SNB_CUDA(cudaLaunchKernelEx(&cfg, snb_gemm_kernel<EPI, CG>, a));
  }
  return launch_status("snb_gemm_kernel");
}

int gemm_launch(const GemmArgs& a, int epi, cudaStream_t st) {
  SNB_CHECK_ARG(a.block_n >= 16 && a.block_n <= GEMM_MAX_BLOCK_N && (a.block_n % 16) == 0, SNB_ERR_UNSUPPORTED,
                "gemm: block_n %d unsupported", a.block_n);
  SNB_CHECK_ARG(a.kb_total >= 1 && a.M >= 1 && a.N >= 1, SNB_ERR_INVALID, "gemm: empty problem");
  SNB_CHECK_ARG(a.a_stages >= 1 && a.b_stages >= 1 && a.a_stages <= GEMM_MAX_RING && a.b_stages <= GEMM_MAX_RING &&
                    a.a_stages * GEMM_A_STAGE + a.b_stages * (int)a.b_slot <= GEMM_OPERAND_BYTES,
                SNB_ERR_INVALID, "gemm: ring depths %d/%d exceed the operand smem", a.a_stages, a.b_stages);
  SNB_CHECK_ARG(a.splits == 1 || epi == EPI_WGRAD, SNB_ERR_INVALID, "gemm: split-K only for the accumulate epilogue");
  SNB_CHECK_ARG(epi == EPI_F32ROWS || epi == EPI_WGRAD, SNB_ERR_UNSUPPORTED,
                "gemm: epilogue %d is served by the chained kernel (k2_chain.cu)", epi);
  if (epi == EPI_F32ROWS)
    SNB_CHECK_ARG(a.block_n == 16 && a.n_tiles == 1, SNB_ERR_UNSUPPORTED, "gemm: the row epilogue needs N == 16");
  const int sms = num_sms();
  if (sms <= 0) return SNB_ERR_NO_DEVICE;
  const long long tiles = (long long)a.m_tiles * a.n_tiles * a.splits;
  const int max_units = a.cta_group == 2 ? sms / 2 : sms;
  const int grid = (int)(tiles < max_units ? tiles : max_units);
  SNB_CHECK_ARG(a.colsum == nullptr || (epi == EPI_WGRAD && tiles <= max_units && a.a_mn), SNB_ERR_UNSUPPORTED,
                "gemm: the fused column sum needs the wgrad form and at most one tile per scheduling unit");
  SNB_CHECK_ARG(a.b_stages * (int)a.b_slot + a.a_stages * GEMM_A_STAGE <= GEMM_OPERAND_BYTES, SNB_ERR_INVALID,
                "gemm: rings %d/%d exceed the operand smem", a.a_stages, a.b_stages);
  SNB_CHECK_ARG(a.side_n == 0 || (epi == EPI_WGRAD && a.cta_group == 2 && tiles <= max_units && a.a_mn && a.side_out != nullptr &&
                                  (a.side_n == 16 || a.side_n == 64) &&
                                  a.b_stages * GEMM_SIDE_STAGE <= GEMM_NUM_STAGING * GEMM_STAGING),
                SNB_ERR_UNSUPPORTED, "gemm: the side operand needs the SM-pair wgrad form, one tile per unit and 16 or 64 columns");
  const double macs = (double)a.m_tiles * GEMM_BLOCK_M * a.cta_group * (double)a.n_tiles * a.block_n * (double)a.kb_total * GEMM_BLOCK_K;
  const bool timed = profile_gemm_begin(st, macs, epi, a.M, a.N, a.kb_total * GEMM_BLOCK_K, a.cta_group, a.splits);
  int rc;
  const bool two = a.cta_group == 2;
  switch (epi) {
    case EPI_F32ROWS: rc = launch_epi<EPI_F32ROWS, 1>(a, grid, st); break;
    case EPI_WGRAD: rc = two ? launch_epi<EPI_WGRAD, 2>(a, grid, st) : launch_epi<EPI_WGRAD, 1>(a, grid, st); break;
    default: set_error("gemm: unknown epilogue %d", epi); rc = SNB_ERR_INVALID;
  }
  if (timed) profile_gemm_end(st);
  return rc;
}

}  // namespace snb

// ---- test hook ---------------------------------------------------------------------------------
extern "C" int snb_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int N, int K,
                             int a_mn, int b_mn, int epi, void* out0, void* out1, int64_t ldo, const void* mul,
                             const float* bias, float w0, int splits, void* stream) {
  using namespace snb;
  SNB_CHECK_ARG(A && B && out0, SNB_ERR_INVALID, "gemm_bf16: null operand");
  SNB_CHECK_ARG(M > 0 && N > 0 && K > 0 && M < (1ll << 31), SNB_ERR_INVALID, "gemm_bf16: bad shape");
  if (epi == EPI_SIN || epi == EPI_LINEAR || epi == EPI_MUL) {
    // bf16-output epilogues: a one-layer chain (out1 = the sign mask: written by EPI_SIN, read by EPI_MUL)
    SNB_CHECK_ARG(splits == 1, SNB_ERR_INVALID, "gemm_bf16: split-K only for the accumulate epilogue");
    SNB_CHECK_ARG(!a_mn && !b_mn && N % 256 == 0, SNB_ERR_UNSUPPORTED, "gemm_bf16: bf16 epilogues need K-major operands and N %% 256 == 0");
    SNB_CHECK_ARG(epi != EPI_MUL || mul != nullptr, SNB_ERR_INVALID, "gemm_bf16: mul operand required");
    ChainArgs* ca = new ChainArgs();
    memset(ca, 0, sizeof(*ca));
    ca->n_passes = 1;
    ca->pass[0].M = (int)M;
    ca->pass[0].n_blocks = (int)((M + 255) / 256);
    ca->pass[0].n_layers = 1;
    ca->n_layers = 1;
    ChainLayer& ly = ca->layers[0];
    ChainMaps& mp = ca->maps[0];
    ly.epi = epi;
    ly.n_tiles = N / 256;
    ly.nseg = 1;
    ly.kb_total = ly.seg_kb[0] = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
    ly.tail_k16 = 4;
    int r = make_tmap_2d(&mp.tmA[0], A, 2, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, 64, GEMM_BLOCK_M);
    if (!r) r = make_tmap_2d(&mp.tmB, B, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, 64, 128);
    if (!r) r = make_tmap_2d(&mp.tmO0, out0, 2, (uint64_t)N, (uint64_t)M, (uint64_t)ldo * 2, 64, GEMM_BLOCK_M);
    if (!r && epi == EPI_MUL) r = make_tmap_2d(&mp.tmMul, mul, 2, (uint64_t)N, (uint64_t)M, (uint64_t)ldo * 2, 64, GEMM_BLOCK_M);
    ly.mul_siren = (epi == EPI_MUL && out1 != nullptr) ? 1 : 0;
    ly.mask = epi == EPI_LINEAR ? nullptr : reinterpret_cast<uint32_t*>(out1);
    ly.mask_ld = N / 32;
    if (!r && ly.mask) r = make_tmap_mask(&mp.tmMask, ly.mask, (uint64_t)(N / 32), (uint64_t)M, (uint64_t)(N / 32) * 4);
    ly.bias = bias;
    ly.w0 = w0;
    if (!r) r = chain_launch(*ca, (cudaStream_t)stream);
    delete ca;
    return r;
  }
  GemmArgs a;
  memset(&a, 0, sizeof(a));
  a.M = (int)M;
  a.N = N;
  a.block_n = N >= 256 ? 256 : (N >= 128 ? 128 : (N >= 64 ? 64 : (N >= 32 && epi == EPI_WGRAD ? 32 : 16)));
  a.kb_total = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  a.cta_group = gemm_pick_cta_group(epi, M, N, a.block_n);
  const uint32_t b_rows = (uint32_t)(a.block_n / a.cta_group);
  a.nseg = 1;
  a.seg_kb[0] = a.kb_total;
  a.a_mn = a_mn;
  a.b_mn = b_mn;
  a.splits = splits;
  a.bias = bias;
  a.w0 = w0;
  a.ldo = ldo;
  int r;
  // K-major operand: [rows, K] box {64 k, rows};  MN-major: [K, rows] box {64 rows, 64 k}
  if (!a_mn) r = make_tmap_2d(&a.tmA[0], A, 2, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, 64, GEMM_BLOCK_M);
  else r = make_tmap_2d(&a.tmA[0], A, 2, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, 64);
  if (r) return r;
  if (!b_mn) r = make_tmap_2d(&a.tmB, B, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, 64, b_rows);
  else r = make_tmap_2d(&a.tmB, B, 2, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, 64);
  if (r) return r;
  if (epi == EPI_F32ROWS) {
    a.f32out = (float*)out0;
  } else if (epi == EPI_WGRAD) {
    if (a.block_n >= 32 && N % 32 == 0) {
      if ((r = make_tmap_2d(&a.tmO0, out0, 4, (uint64_t)N, (uint64_t)M, (uint64_t)ldo * 4, 32, GEMM_BLOCK_M))) return r;
    } else {
      a.f32out = (float*)out0;
    }
  } else {
    SNB_CHECK_ARG(false, SNB_ERR_UNSUPPORTED, "gemm_bf16: epilogue %d not available through the test hook", epi);
  }
  gemm_finalize(a);
  return gemm_launch(a, epi, (cudaStream_t)stream);
}
