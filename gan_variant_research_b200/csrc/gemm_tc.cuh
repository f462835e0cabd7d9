// netF head (north-star piece 3) on tcgen05:
//   Linear(C_l, nc) -> ReLU -> Linear(nc, nc) applied to the gathered patches of both sides, and its
//   backward (dH, dX, dW1, db1, dW2, db2).  Every matrix lives in HBM as bf16 hi(+lo) "operand
//   blobs" already tiled as the tensor core's no-swizzle canonical shared-memory image (8x8 core
//   matrices of 128 B), so one 1-D bulk copy (UBLKCP) lands an operand chunk ready for UTCHMMA:
//     row blob   [tile of 128 rows][col/8][16 row groups][8 rows][8 cols]   activations X, H, Y, dY, dH
//     weight blob          [k/8][n/8 row groups][8 n][8 k]                  W as the B operand (N x K)
//   * k_gemm_tc  : D(128 x N) = A_tile(128 x K) B^T, both K-major, K streamed in chunks of 16 through a
//                  4-slot ring, two CTAs per SM; warp-specialised like k_loss_tc (producer / MMA issuer / 4 epilogue
//                  warps, thread <-> row).  The epilogue mode decides what leaves the SM:
//                  H = relu(.+b1) as the next GEMM's A blob; Y = .+b2 directly in the operand formats
//                  k_loss_tc consumes (so the logits kernel runs unchanged on the head's output);
//                  dH = (.) * [H>0] as a blob; dX as the transposed fp32 rows the dense backward reads.
//   * k_wgrad_tc : dW(128 x N) = sum_m A[m,:]^T B[m,:], both operands read MN-MAJOR from the very same
//                  row blobs (no transposed copies), split-K over row tiles, fp32 partials + a
//                  deterministic reduce; the bias gradient is one extra N=16 MMA against a tile of ones.
// bf16x3 (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM) keeps the 1e-3 parity of the north star;
// single-pass bf16 is the fast mode.
#pragma once
#include "common.cuh"
#include "loss_tc.cuh"
#include "umma.cuh"

namespace pnce {

enum GemmMode { GM_H = 0, GM_YQ = 1, GM_YK = 2, GM_DH = 3, GM_DX = 4, GM_YROWS = 5 };

struct GemmProb {
  const __nv_bfloat16 *a_hi, *a_lo;       // row blob, K columns
  const __nv_bfloat16 *b_hi, *b_lo;       // weight blob, N x K
  const float* bias;                       // [N] or NULL
  int K, N, tiles, mode;                   // K, N multiples of 32, <= 256
  int P, Ppad, halves, C;                  // geometry: tile = b * halves + mh, row = patch mh*128 + i
  __nv_bfloat16 *o_hi, *o_lo;              // GM_H / GM_DH / GM_YQ: row blob with N columns
  __nv_bfloat16 *k_hi, *k_lo, *k2_hi, *k2_lo;   // GM_YK: the two K layouts of k_loss_tc
  float* ss;                               // GM_YQ / GM_YK: [B][N/32][Ppad] partial sums of squares
  float* outT;                             // GM_DX: dxT [B][C][Ppad], or row-major [B][Ppad][C] with rm != 0
  const __nv_bfloat16* mask_hi;            // GM_DH: H blob (hi part), same indexing as o_hi
  // GM_YROWS (PatchSampleF(use_mlp=True).forward, persistent kernel only): y / max(||y||, eps) leaves as fp32 rows
  // (B*P, N) in the CALLER's row order (row b*P + perm[slot]), with the norm bookkeeping of pnce_sample_fwd in inv_out
  float* rows_out;
  float* inv_out;
  const int* perm;                         // sorted slot -> original index (k_prep)
  int rm;                                  // GM_DX: row-major rows (channels-last maps, nhwc.cuh)
};

constexpr int kGemmMaxProb = 16;
struct GemmLaunch {
  GemmProb pr[kGemmMaxProb];
  long long start[kGemmMaxProb + 1];
  int n, x3;
  int* err;                                // protocol-timeout flag (nonfinite[1])
  int dbg;                                 // experiment build only: 1 = epilogue without global stores, 2 = empty epilogue
};

// ring: 4 slots x 24 KB (16 columns of K per stage: A hi 4K | A lo 4K | B hi 8K | B lo 8K) = 96 KB, so two
// CTAs fit one SM (one's epilogue overlaps the other's MMAs) and each keeps 3 bulk copies in flight
constexpr int kGemmSlots = 4;
constexpr int kGemmStageBytes = 24576;
constexpr int kGemmOffAlo = 4096, kGemmOffBhi = 8192, kGemmOffBlo = 16384;
constexpr int kGemmSmemBytes = kGemmSlots * kGemmStageBytes + 256;

struct GemmShared {
  uint64_t full[kGemmSlots], empty[kGemmSlots], dfull;
  uint32_t tmem_base;
  int dead;
};

__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    h[k] = bf16x2_bits(v[2 * k], v[2 * k + 1]);
    l[k] = bf16x2_bits(v[2 * k] - __uint_as_float(h[k] << 16), v[2 * k + 1] - __uint_as_float(h[k] & 0xffff0000u));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(kTcThreads, 2) k_gemm_tc(const __grid_constant__ GemmLaunch g) {
  pdl_launch();
  extern __shared__ __align__(1024) unsigned char smem[];
  using namespace umma;
  GemmShared* sh = reinterpret_cast<GemmShared*>(smem + kGemmSlots * kGemmStageBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int pi = 0;
  for (int i = 1; i < g.n; ++i)
    if ((long long)blockIdx.x >= g.start[i]) pi = i;
  const GemmProb& pr = g.pr[pi];
  const int tile = (int)(blockIdx.x - g.start[pi]);
  const int K = pr.K, N = pr.N, nstage = K >> 4, K8 = K >> 3, N8 = N >> 3;
  const bool x3 = g.x3 != 0;
  volatile int* dead = &sh->dead;
  if (tid == 0) {
    for (int k = 0; k < kGemmSlots; ++k) { mbar_init(&sh->full[k], 1); mbar_init(&sh->empty[k], 1); }
    mbar_init(&sh->dfull, 1);
    sh->dead = 0;
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(&sh->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                                                 // prologue done: now wait for the prerequisite grids
  const uint32_t tmem = sh->tmem_base;
  const uint32_t bbytes = (uint32_t)N * 32u;                // one B chunk: 2 slabs x N rows x 16 B
  const uint32_t lbo_b = (uint32_t)N * 16u;

  if (warp == 0) {
    if (lane == 0) {
      const unsigned char* ga_hi = reinterpret_cast<const unsigned char*>(pr.a_hi) + (size_t)tile * K8 * 2048;
      const unsigned char* ga_lo = reinterpret_cast<const unsigned char*>(pr.a_lo) + (size_t)tile * K8 * 2048;
      const unsigned char* gb_hi = reinterpret_cast<const unsigned char*>(pr.b_hi);
      const unsigned char* gb_lo = reinterpret_cast<const unsigned char*>(pr.b_lo);
      for (int s = 0; s < nstage; ++s) {
        const int slot = s % kGemmSlots;
        if (!mbar_wait(&sh->empty[slot], ((uint32_t)(s / kGemmSlots) & 1u) ^ 1u, dead)) break;
        unsigned char* st = smem + slot * kGemmStageBytes;
        mbar_expect_tx(&sh->full[slot], (4096u + bbytes) * (x3 ? 2u : 1u));
        bulk_g2s(st, ga_hi + (size_t)s * 4096, 4096u, &sh->full[slot]);
        if (x3) bulk_g2s(st + kGemmOffAlo, ga_lo + (size_t)s * 4096, 4096u, &sh->full[slot]);
        bulk_g2s(st + kGemmOffBhi, gb_hi + (size_t)s * bbytes, bbytes, &sh->full[slot]);
        if (x3) bulk_g2s(st + kGemmOffBlo, gb_lo + (size_t)s * bbytes, bbytes, &sh->full[slot]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(128, N, 0, 0);
      bool ok = true;
      for (int s = 0; s < nstage && ok; ++s) {
        const int slot = s % kGemmSlots;
        ok = mbar_wait(&sh->full[slot], (uint32_t)(s / kGemmSlots) & 1u, dead);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + slot * kGemmStageBytes);
        const uint64_t a_hi = smem_desc(st, 2048, 128);
        const uint64_t b_hi = smem_desc(st + kGemmOffBhi, lbo_b, 128);
        mma_bf16(tmem, a_hi, b_hi, idesc, s ? 1u : 0u);
        if (x3) {
          const uint64_t a_lo = smem_desc(st + kGemmOffAlo, 2048, 128);
          const uint64_t b_lo = smem_desc(st + kGemmOffBlo, lbo_b, 128);
          mma_bf16(tmem, a_hi, b_lo, idesc, 1u);
          mma_bf16(tmem, a_lo, b_hi, idesc, 1u);
        }
        mma_commit(&sh->empty[slot]);
      }
      mma_commit(&sh->dfull);
    }
  } else {
    // ===================== epilogue: thread <-> row of the tile =====================
    const int q = warp & 3, i = q * 32 + lane;
    const int b = tile / pr.halves, mh = tile - b * pr.halves;
    const int p = mh * 128 + i;                               // patch slot of this row
    const bool rowok = p < pr.P;
    const int Ppad = pr.Ppad, mode = pr.mode;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const size_t rowblob = ((size_t)tile * N8 * 16 + (size_t)(i >> 3)) * 64 + (size_t)(i & 7) * 8;   // + n8 * 1024
    mbar_wait(&sh->dfull, 0u, dead);
    tc_fence_after();
    const int nch = N >> 5;
    for (int ch = 0; ch < nch; ++ch) {
      uint32_t r[32];
      tmem_ld32(trow + ch * 32, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        float x = __uint_as_float(r[k]);
        if (pr.bias != nullptr) x += __ldg(pr.bias + ch * 32 + k);
        if (mode == GM_H) x = fmaxf(x, 0.f);
        v[k] = rowok ? x : 0.f;                               // padding rows stay exactly zero
      }
      if (mode == GM_DX) {
        // same L2 evict_last policy as the loss kernel's dxT stores (loss_tc.cuh, st_dx)
        const bool keep_dx = g_dx_evict_last != 0;
        uint64_t pol_dx = 0;
        if (keep_dx) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_dx));
        if (rowok) {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const int c = ch * 32 + k;
            if (c < pr.C) st_dx(pr.outT + ((size_t)b * pr.C + c) * Ppad + p, v[k], keep_dx, pol_dx);
          }
        }
        continue;
      }
      if (mode == GM_DH) {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const uint4 hm = *reinterpret_cast<const uint4*>(pr.mask_hi + rowblob + (size_t)(ch * 4 + g8) * 1024);
          const uint32_t w[4] = {hm.x, hm.y, hm.z, hm.w};
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) {
            if ((w[k2] & 0x7fffu) == 0u) v[g8 * 8 + 2 * k2] = 0.f;            // relu'(0) = 0
            if ((w[k2] & 0x7fff0000u) == 0u) v[g8 * 8 + 2 * k2 + 1] = 0.f;
          }
        }
      }
      if (mode == GM_YQ || mode == GM_YK) {
        float ssum = 0.f;
        bool bad = false;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          ssum = fmaf(v[k], v[k], ssum);
          bad |= !isfinite(v[k]);
        }
        pr.ss[((size_t)b * nch + ch) * Ppad + p] = bad ? __int_as_float(0x7fc00000) : ssum;
      }
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        float v8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v8[k] = v[g8 * 8 + k];
        uint4 hi, lo;
        split8(v8, hi, lo);
        const int n8 = ch * 4 + g8;
        if (mode == GM_YK) {
          // key blocks of <= 256 rows, one after the other (gather_tc_chunk's K layout; one block when Ppad <= 256)
          const int pb = p >> 8, nb8 = min(256, Ppad - pb * 256) >> 3;
          const size_t o1 = ((((size_t)b * Ppad + (size_t)pb * 256) * N8) / 8 + (size_t)n8 * nb8 + ((p & 255) >> 3)) * 64 + (size_t)(p & 7) * 8;
          const size_t o2 = (((size_t)b * (Ppad >> 3) + (p >> 3)) * N8 + n8) * 64 + (size_t)(p & 7) * 8;
          *reinterpret_cast<uint4*>(pr.k_hi + o1) = hi;
          *reinterpret_cast<uint4*>(pr.k2_hi + o2) = hi;
          if (x3) {
            *reinterpret_cast<uint4*>(pr.k_lo + o1) = lo;
            *reinterpret_cast<uint4*>(pr.k2_lo + o2) = lo;
          }
        } else {
          const size_t o = rowblob + (size_t)n8 * 1024;
          *reinterpret_cast<uint4*>(pr.o_hi + o) = hi;
          if (x3) *reinterpret_cast<uint4*>(pr.o_lo + o) = lo;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem);
  }
  if (tid == 0 && sh->dead && g.err != nullptr) *reinterpret_cast<volatile int*>(g.err) = 1;   // plain store: the flag may live in mapped host memory
}

// -------------------------------------------------------------------------------------------------
// Weight gradients.  dW(128 x N) = sum over row tiles of A_tile^T B_tile with A = dY or dH (the 128
// output rows are a 128-column slice of the A blob) and B = H or X, both read MN-major.
// -------------------------------------------------------------------------------------------------
struct WgradProb {
  const __nv_bfloat16 *a_hi, *a_lo;       // row blob with NA columns (gradient side)
  const __nv_bfloat16 *b_hi, *b_lo;       // row blob with N columns (activation side)
  int NA, N, tiles, slabs;                 // tiles = row tiles (K = 128 each); slabs = split-K factor
  float* partial;                          // [slabs][NA/128][128][N]
  float* pbias;                            // [slabs][NA]
};
struct WgradLaunch {
  WgradProb pr[kGemmMaxProb];
  long long start[kGemmMaxProb + 1];
  int n, x3;
  int* err;
};
// (Measured and rejected: two stages of HALF a row tile each -- 64 rows of K, 16 + N/8 bulk copies of 1 KB per part because
// the 64 rows of an 8-column slab are 1 KB inside the slab's 2 KB -- so that loads fly under the other half's MMAs:
// 117 us vs 95 us at B=64; the bulk-copy engine is built for few large copies, DESIGN.md 4.6.  Also rejected: the A tile
// resident and two buffers for the N/2-column HALVES of the B tile (only the A tile loaded with nothing to hide behind):
// 144 us vs 87 us -- an M=128 MMA costs ~130 cycles whatever N is, so halving N doubles the MMA time.)
constexpr int kWgOffAlo = 32768, kWgOffBhi = 65536, kWgOffBlo = 131072, kWgOffOnes = 196608;
constexpr int kWgSmemBytes = 196608 + 4096 + 256;

__global__ void __launch_bounds__(kTcThreads, 1) k_wgrad_tc(const __grid_constant__ WgradLaunch g) {
  pdl_launch();
  extern __shared__ __align__(1024) unsigned char smem[];
  using namespace umma;
  struct Sh { uint64_t full, empty, dfull; uint32_t tmem_base; int dead; };
  Sh* sh = reinterpret_cast<Sh*>(smem + kWgOffOnes + 4096);      // (the 4 KB at kWgOffOnes held a tile of ones: see the bias note below)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int pi = 0;
  for (int i = 1; i < g.n; ++i)
    if ((long long)blockIdx.x >= g.start[i]) pi = i;
  const WgradProb& pr = g.pr[pi];
  const int local = (int)(blockIdx.x - g.start[pi]);
  const int nhalf = pr.NA >> 7;
  const int ih = local % nhalf, slab = local / nhalf;
  const int N = pr.N, N8 = N >> 3, NA8 = pr.NA >> 3;
  const int t0 = (int)((long long)pr.tiles * slab / pr.slabs), t1 = (int)((long long)pr.tiles * (slab + 1) / pr.slabs);
  const bool x3 = g.x3 != 0;
  volatile int* dead = &sh->dead;
  if (tid == 0) {
    // empty: the MMA thread's commit + one arrival per epilogue warp (they read the A tile for the bias sums)
    mbar_init(&sh->full, 1); mbar_init(&sh->empty, 1 + 4); mbar_init(&sh->dfull, 1);
    sh->dead = 0;
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&sh->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                                                 // prologue done: now wait for the prerequisite grids
  const uint32_t tmem = sh->tmem_base;
  const uint32_t abytes = 16u * 2048u, bbytes = (uint32_t)N8 * 2048u;

  if (warp == 0) {
    if (lane == 0) {
      for (int t = t0; t < t1; ++t) {
        // the single stage cannot load ahead; pull the NEXT row tile's operands into L2 while this one's MMAs run
        // (both blobs come from DRAM otherwise, and that latency would sit in front of every step)
        if (t + 1 < t1) {
          const size_t an = ((size_t)(t + 1) * NA8 + (size_t)ih * 16) * 2048, bn = (size_t)(t + 1) * N8 * 2048;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const unsigned char*>(pr.a_hi) + an), "r"(abytes) : "memory");
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const unsigned char*>(pr.b_hi) + bn), "r"(bbytes) : "memory");
          if (x3) {
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const unsigned char*>(pr.a_lo) + an), "r"(abytes) : "memory");
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const unsigned char*>(pr.b_lo) + bn), "r"(bbytes) : "memory");
          }
        }
        if (!mbar_wait(&sh->empty, ((uint32_t)(t - t0) & 1u) ^ 1u, dead)) break;
        mbar_expect_tx(&sh->full, (abytes + bbytes) * (x3 ? 2u : 1u));
        const size_t ao = ((size_t)t * NA8 + (size_t)ih * 16) * 2048;
        const size_t bo = (size_t)t * N8 * 2048;
        bulk_g2s(smem, reinterpret_cast<const unsigned char*>(pr.a_hi) + ao, abytes, &sh->full);
        bulk_g2s(smem + kWgOffBhi, reinterpret_cast<const unsigned char*>(pr.b_hi) + bo, bbytes, &sh->full);
        if (x3) {
          bulk_g2s(smem + kWgOffAlo, reinterpret_cast<const unsigned char*>(pr.a_lo) + ao, abytes, &sh->full);
          bulk_g2s(smem + kWgOffBlo, reinterpret_cast<const unsigned char*>(pr.b_lo) + bo, bbytes, &sh->full);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(128, N, 1, 1);        // A and B both MN-major
      const uint32_t sa = smem_u32(smem), sal = smem_u32(smem + kWgOffAlo);
      const uint32_t sb = smem_u32(smem + kWgOffBhi), sbl = smem_u32(smem + kWgOffBlo);
      bool ok = true;
      for (int t = t0; t < t1 && ok; ++t) {
        ok = mbar_wait(&sh->full, (uint32_t)(t - t0) & 1u, dead);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {                     // 16 rows m per MMA: 2 row groups of 128 B
          // MN-major: LBO = stride between 8-row K groups (128 B), SBO = between 8-column MN chunks (2048 B)
          const uint64_t a_hi = smem_desc(sa + ks * 256, 128, 2048);
          const uint64_t b_hi = smem_desc(sb + ks * 256, 128, 2048);
          const uint32_t acc = (t > t0 || ks) ? 1u : 0u;
          mma_bf16(tmem, a_hi, b_hi, idesc, acc);
          if (x3) {
            const uint64_t a_lo = smem_desc(sal + ks * 256, 128, 2048);
            const uint64_t b_lo = smem_desc(sbl + ks * 256, 128, 2048);
            mma_bf16(tmem, a_hi, b_lo, idesc, 1u);
            mma_bf16(tmem, a_lo, b_hi, idesc, 1u);
          }
        }
        mma_commit(&sh->empty);
      }
      mma_commit(&sh->dfull);
    }
  } else {
    const int q = warp & 3, i = q * 32 + lane;                // output row = column ih*128 + i of the A blob
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    // ---- bias gradient = column sums of the A tile, on the CUDA cores while the MMAs run.  (It used to be two extra MMAs
    //      per 16 rows against a tile of ones, N = 16: an M = 128 MMA costs ~130 cycles whatever N is, so they were 16 of a
    //      row tile's 40 MMAs.)  Thread <-> (8-column slab et >> 3, row et & 7 of every 8-row group): 16 + 16 LDS.128 per
    //      row tile, 8 partial sums per thread, folded over the 8 rows of a group by shuffles at the end ----
    const int et = tid - 64;
    float bs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) bs[k] = 0.f;
    {
      const unsigned char* arow = smem + (size_t)(et >> 3) * 2048 + (size_t)(et & 7) * 16;
      for (int t = t0; t < t1; ++t) {
        if (!mbar_wait(&sh->full, (uint32_t)(t - t0) & 1u, dead)) break;
#pragma unroll 4
        for (int rg = 0; rg < 16; ++rg) {
          const uint4 h = *reinterpret_cast<const uint4*>(arow + rg * 128);
          const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) {
            bs[2 * k2] += __uint_as_float(hw[k2] << 16);
            bs[2 * k2 + 1] += __uint_as_float(hw[k2] & 0xffff0000u);
          }
          if (x3) {
            const uint4 l = *reinterpret_cast<const uint4*>(arow + kWgOffAlo + rg * 128);
            const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) {
              bs[2 * k2] += __uint_as_float(lw[k2] << 16);
              bs[2 * k2 + 1] += __uint_as_float(lw[k2] & 0xffff0000u);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->empty);                 // this warp is done with the stage
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        bs[k] += __shfl_xor_sync(0xffffffffu, bs[k], 1);
        bs[k] += __shfl_xor_sync(0xffffffffu, bs[k], 2);
        bs[k] += __shfl_xor_sync(0xffffffffu, bs[k], 4);
      }
    }
    mbar_wait(&sh->dfull, 0u, dead);
    tc_fence_after();
    float* out = pr.partial + (((size_t)slab * nhalf + ih) * 128 + i) * N;
    const bool any = t1 > t0;
    for (int ch = 0; ch < (N >> 5); ++ch) {
      uint32_t r[32];
      tmem_ld32(trow + ch * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        float4 o;
        o.x = any ? __uint_as_float(r[k4 * 4 + 0]) : 0.f;
        o.y = any ? __uint_as_float(r[k4 * 4 + 1]) : 0.f;
        o.z = any ? __uint_as_float(r[k4 * 4 + 2]) : 0.f;
        o.w = any ? __uint_as_float(r[k4 * 4 + 3]) : 0.f;
        *reinterpret_cast<float4*>(out + ch * 32 + k4 * 4) = o;
      }
    }
    if ((et & 7) == 0) {                                      // columns (et >> 3) * 8 ... + 7 of this CTA's 128
      float* pb = pr.pbias + (size_t)slab * pr.NA + ih * 128 + (et >> 3) * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) pb[k] = any ? bs[k] : 0.f;
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
  if (tid == 0 && sh->dead && g.err != nullptr) *reinterpret_cast<volatile int*>(g.err) = 1;   // plain store: the flag may live in mapped host memory
}

// dW = upstream * sum_slabs partial, db likewise; deterministic order; all layers in one launch.
struct WreduceJob {
  const float* partial; const float* pbias; float* dw; float* db;
  int NA, N, Nout, slabs;                  // dw is (NA, Nout) row-major, Nout <= N (channel padding dropped)
};
struct WreduceLaunch {
  WreduceJob job[2 * PNCE_MAX_LAYERS];
  int start[2 * PNCE_MAX_LAYERS + 1];      // block prefix
  int n;
  const float* grad_out;
};
__global__ void __launch_bounds__(kThreads) k_wreduce(const __grid_constant__ WreduceLaunch gl) {
  pdl_enter();
  int ji = 0;
  for (int i = 1; i < gl.n; ++i)
    if ((int)blockIdx.x >= gl.start[i]) ji = i;
  const WreduceJob& a = gl.job[ji];
  const float g = gl.grad_out ? __ldg(gl.grad_out) : 1.0f;
  const long long total = (long long)a.NA * a.N;
  const long long e = (long long)(blockIdx.x - gl.start[ji]) * kThreads + threadIdx.x;
  if (e < total) {
    const int i = (int)(e / a.N), n = (int)(e - (long long)i * a.N);
    if (n < a.Nout) {
      float s = 0.f;
      for (int k = 0; k < a.slabs; ++k) s += a.partial[(size_t)k * total + e];
      a.dw[(size_t)i * a.Nout + n] = s * g;
    }
  } else if (e - total < a.NA) {
    const int i = (int)(e - total);
    float s = 0.f;
    for (int k = 0; k < a.slabs; ++k) s += a.pbias[(size_t)k * a.NA + i];
    a.db[i] = s * g;
  }
}

// fp32 weights -> weight blobs, every blob of every layer in ONE launch.  w is (R, Cw) row-major; the
// blob holds B[n][k] = transpose ? w[k][n] : w[n][k] for n < N, k < K (zero padded).
// One block of 64 threads per 8x8 core matrix.
struct WprepJob {
  const float* w; __nv_bfloat16 *hi, *lo;
  int R, Cw, N, K, transpose;
};
struct WprepLaunch {
  WprepJob job[4 * PNCE_MAX_LAYERS];
  int start[4 * PNCE_MAX_LAYERS + 1];
  int n;
};
__global__ void __launch_bounds__(64) k_wprep(const __grid_constant__ WprepLaunch g) {
  pdl_enter();
  int ji = 0;
  for (int i = 1; i < g.n; ++i)
    if ((int)blockIdx.x >= g.start[i]) ji = i;
  const WprepJob& a = g.job[ji];
  const int blk = blockIdx.x - g.start[ji];
  const int N8 = a.N >> 3;
  const int n8 = blk % N8, k8 = blk / N8;
  const int n = n8 * 8 + (threadIdx.x >> 3), k = k8 * 8 + (threadIdx.x & 7);
  float v = 0.f;
  if (!a.transpose) { if (n < a.R && k < a.Cw) v = a.w[(size_t)n * a.Cw + k]; }
  else { if (k < a.R && n < a.Cw) v = a.w[(size_t)k * a.Cw + n]; }
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const size_t o = ((size_t)k8 * N8 + n8) * 64 + threadIdx.x;
  a.hi[o] = h;
  if (a.lo != nullptr) a.lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
}

}  // namespace pnce
