// netF head GEMMs, PERSISTENT variant of k_gemm_tc (gemm_tc.cuh): same operand blobs, same five epilogue
// modes, but one CTA per SM walks the tile list of the launch with a DOUBLE-BUFFERED TMEM accumulator
// (2 x 256 columns), so the MMAs of tile n+1 run under the epilogue of tile n:
//   warp 0      bulk-copy producer, 8-slot ring of 16-column K stages (192 KB): runs up to a whole tile ahead
//   warp 1      MMA issuer (waits accfree[buf], commits accfull[buf])
//   warps 2..9  epilogue, two warps per TMEM lane quadrant (= two per warp scheduler): the pair splits the
//               32-column chunks of its rows even/odd -- every chunk (bias, ReLU / mask, sums of squares,
//               hi/lo split, blob store) is self-contained, so the two never exchange anything
// Tiles are taken round-robin (CTA c: tiles c, c+G, ...): neighbouring SMs work on the same problem at the
// same time and share its weight blob in L2.  Launch, TMEM allocation and barrier set-up once per SM.
#pragma once
#include "gemm_tc.cuh"

namespace pnce {

constexpr int kGpEpiWarps = 16;            // four per TMEM lane quadrant = four per warp scheduler
constexpr int kGpThreads = 64 + 32 * kGpEpiWarps;
// stages of 32 columns of K (A hi 8K | A lo 8K | B hi 16K | B lo 16K = 48 KB), 4 slots: the same 192 KB in flight as 8
// slots of 16 columns, in half as many (twice as large) bulk copies
constexpr int kGpSlots = 4;
constexpr int kGpStageBytes = 2 * kGemmStageBytes;
constexpr int kGpOffAlo = 2 * kGemmOffAlo, kGpOffBhi = 2 * kGemmOffBhi, kGpOffBlo = 2 * kGemmOffBlo;
constexpr int kGpSmemBytes = kGpSlots * kGpStageBytes + 512 + 4096 + 2048;   // + GM_YROWS row-norm exchange [2][4][128] floats + bias [2][256]

struct GpShared {
  uint64_t full[kGpSlots], empty[kGpSlots], accfull[2], accfree[2];
  uint32_t tmem_base;
  int dead;
};
static_assert(sizeof(GpShared) <= 512, "GpShared must fit its slot");

struct GpTile {
  const GemmProb* pr;
  int tile, K, N, nstage;
};
__device__ __forceinline__ void gp_decode(const GemmLaunch& g, long long t, GpTile& o) {
  int pi = 0;
  for (int i = 1; i < g.n; ++i)
    if (t >= g.start[i]) pi = i;
  o.pr = &g.pr[pi];
  o.tile = (int)(t - g.start[pi]);
  o.K = o.pr->K; o.N = o.pr->N; o.nstage = o.K >> 5;
}

__global__ void __launch_bounds__(kGpThreads, 1) k_gemm_tc_p(const __grid_constant__ GemmLaunch g) {
  pdl_launch();
  extern __shared__ __align__(1024) unsigned char smem[];
  using namespace umma;
  GpShared* sh = reinterpret_cast<GpShared*>(smem + kGpSlots * kGpStageBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long total = g.start[g.n];
  const bool x3 = g.x3 != 0;
  volatile int* dead = &sh->dead;
  if (tid == 0) {
    for (int k = 0; k < kGpSlots; ++k) { mbar_init(&sh->full[k], 1); mbar_init(&sh->empty[k], 1); }
    for (int k = 0; k < 2; ++k) { mbar_init(&sh->accfull[k], 1); mbar_init(&sh->accfree[k], 32 * kGpEpiWarps); }
    sh->dead = 0;
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&sh->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                                                 // prologue done: now wait for the prerequisite grids
  const uint32_t tmem = sh->tmem_base;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      bool ok = true;
      for (long long t = blockIdx.x; t < total && ok; t += gridDim.x) {
        GpTile w;
        gp_decode(g, t, w);
        const GemmProb& pr = *w.pr;
        const uint32_t bbytes = (uint32_t)w.N * 64u;          // one B chunk: 4 slabs x N rows x 16 B
        const unsigned char* ga_hi = reinterpret_cast<const unsigned char*>(pr.a_hi) + (size_t)w.tile * (w.K >> 3) * 2048;
        const unsigned char* ga_lo = reinterpret_cast<const unsigned char*>(pr.a_lo) + (size_t)w.tile * (w.K >> 3) * 2048;
        const unsigned char* gb_hi = reinterpret_cast<const unsigned char*>(pr.b_hi);
        const unsigned char* gb_lo = reinterpret_cast<const unsigned char*>(pr.b_lo);
        // (measured and rejected: an L2 prefetch of the next tile's activations and ReLU mask here -- GEMM launches 228 ->
        //  260 us per step; the same prefetch does pay in k_wgrad_tc, whose single stage cannot load ahead)
        for (int s = 0; s < w.nstage; ++s, ++it) {
          const int slot = it % kGpSlots;
          ok = mbar_wait(&sh->empty[slot], ((it / kGpSlots) & 1u) ^ 1u, dead);
          if (!ok) break;
          unsigned char* st = smem + slot * kGpStageBytes;
          mbar_expect_tx(&sh->full[slot], (8192u + bbytes) * (x3 ? 2u : 1u));
          bulk_g2s(st, ga_hi + (size_t)s * 8192, 8192u, &sh->full[slot]);
          if (x3) bulk_g2s(st + kGpOffAlo, ga_lo + (size_t)s * 8192, 8192u, &sh->full[slot]);
          bulk_g2s(st + kGpOffBhi, gb_hi + (size_t)s * bbytes, bbytes, &sh->full[slot]);
          if (x3) bulk_g2s(st + kGpOffBlo, gb_lo + (size_t)s * bbytes, bbytes, &sh->full[slot]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0;
      int n = 0;
      bool ok = true;
      for (long long t = blockIdx.x; t < total && ok; t += gridDim.x, ++n) {
        GpTile w;
        gp_decode(g, t, w);
        const uint32_t idesc = idesc_bf16(128, w.N, 0, 0);
        const uint32_t lbo_b = (uint32_t)w.N * 16u;
        const uint32_t acc = tmem + (uint32_t)(n & 1) * 256u;
        // the accumulator buffer must have been read out by the epilogue of tile n-2
        ok = mbar_wait(&sh->accfree[n & 1], (((uint32_t)n >> 1) & 1u) ^ 1u, dead);
        tc_fence_after();
        for (int s = 0; s < w.nstage && ok; ++s, ++it) {
          const int slot = it % kGpSlots;
          ok = mbar_wait(&sh->full[slot], (it / kGpSlots) & 1u, dead);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + slot * kGpStageBytes);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {                     // 16 columns of K = 2 slabs per MMA
            const uint64_t a_hi = smem_desc(st + ks * 4096, 2048, 128);
            const uint64_t b_hi = smem_desc(st + kGpOffBhi + ks * 2 * lbo_b, lbo_b, 128);
            mma_bf16(acc, a_hi, b_hi, idesc, (s | ks) ? 1u : 0u);
            if (x3) {
              const uint64_t a_lo = smem_desc(st + kGpOffAlo + ks * 4096, 2048, 128);
              const uint64_t b_lo = smem_desc(st + kGpOffBlo + ks * 2 * lbo_b, lbo_b, 128);
              mma_bf16(acc, a_hi, b_lo, idesc, 1u);
              mma_bf16(acc, a_lo, b_hi, idesc, 1u);
            }
          }
          mma_commit(&sh->empty[slot]);
        }
        mma_commit(&sh->accfull[n & 1]);
      }
    }
  } else {
    // ===================== epilogue: two warps per quadrant, thread <-> row of the tile =====================
    const int q = warp & 3, i = q * 32 + lane;
    const int half = (warp - 2) >> 2;                        // this warp's 32-column chunks: half, half + 4, ...
    constexpr int kSub = kGpEpiWarps / 4;
    int n = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x, ++n) {
      GpTile w;
      gp_decode(g, t, w);
      const GemmProb& pr = *w.pr;
      const int tile = w.tile, N = w.N, N8 = N >> 3;
      const int b = tile / pr.halves, mh = tile - b * pr.halves;
      const int p = mh * 128 + i;                             // patch slot of this row
      const bool rowok = p < pr.P;
      const int Ppad = pr.Ppad, mode = pr.mode;
      const uint32_t trow = tmem + (uint32_t)(n & 1) * 256u + ((uint32_t)(q * 32) << 16);
      const size_t rowblob = ((size_t)tile * N8 * 16 + (size_t)(i >> 3)) * 64 + (size_t)(i & 7) * 8;   // + n8 * 1024
      // this tile's bias (or zeros) in shared memory: one load per epilogue thread instead of one per element and thread
      float* bias_s = reinterpret_cast<float*>(smem + kGpSlots * kGpStageBytes + 512 + 4096) + (n & 1) * 256;
      {
        const int et = tid - 64;
        if (et < 256) bias_s[et] = (pr.bias != nullptr && et < N) ? __ldg(pr.bias + et) : 0.f;
        asm volatile("bar.sync 5, %0;" ::"n"(32 * kGpEpiWarps) : "memory");
      }
      mbar_wait(&sh->accfull[n & 1], ((uint32_t)n >> 1) & 1u, dead);
      tc_fence_after();
      const int nch = N >> 5;
#ifdef PNCE_EXPERIMENTS
      if (g.dbg == 2) {                                        // experiment: what does the launch cost without any epilogue work?
        tc_fence_before();
        mbar_arrive(&sh->accfree[n & 1]);
        continue;
      }
#endif
      if (mode == GM_YROWS) {
        // ---- PatchSampleF(use_mlp=True) output: y = acc + b2, out = y / max(||y||, eps) as fp32 rows.  The row norm
        //      needs every column, and the two warps of a quadrant hold alternate chunks: pass 1 sums the squares and
        //      swaps the partial sums through shared memory, pass 2 re-reads the accumulator and writes the rows ----
        float* xss = reinterpret_cast<float*>(smem + kGpSlots * kGpStageBytes + 512) + (n & 1) * (kSub * 128);
        float ssum = 0.f;
        for (int ch = half; ch < nch; ch += kSub) {
          uint32_t r[32];
          tmem_ld32(trow + ch * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float x = __uint_as_float(r[k]) + bias_s[ch * 32 + k];
            ssum = fmaf(x, x, ssum);
          }
        }
        xss[half * 128 + i] = ssum;
        switch (q) {                                              // the two warps of this quadrant (immediate barrier ids)
          case 0: asm volatile("bar.sync 1, %0;" ::"n"(32 * kSub) : "memory"); break;
          case 1: asm volatile("bar.sync 2, %0;" ::"n"(32 * kSub) : "memory"); break;
          case 2: asm volatile("bar.sync 3, %0;" ::"n"(32 * kSub) : "memory"); break;
          default: asm volatile("bar.sync 4, %0;" ::"n"(32 * kSub) : "memory"); break;
        }
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < kSub; ++k) tot += xss[k * 128 + i];
        const float nrm = sqrtf(tot);
        const float den = (nrm == nrm) ? fmaxf(nrm, kNormEps) : nrm;      // clamp_min keeps NaN (F.normalize, patchnce_cut.py:77)
        const float scale = 1.0f / den;
        const size_t orow = ((size_t)b * pr.P + (rowok ? __ldg(pr.perm + p) : 0)) * N;
        if (rowok && half == 0 && pr.inv_out != nullptr)
          pr.inv_out[orow / N] = (nrm == nrm) ? (nrm < kNormEps ? -1.0f / kNormEps : 1.0f / nrm) : nrm;
        for (int ch = half; ch < nch; ch += kSub) {
          uint32_t r[32];
          tmem_ld32(trow + ch * 32, r);
          tmem_ld_wait();
          if (rowok) {
            float* o = pr.rows_out + orow + ch * 32;
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              float4 v4;
              v4.x = (__uint_as_float(r[k4 * 4 + 0]) + bias_s[ch * 32 + k4 * 4 + 0]) * scale;
              v4.y = (__uint_as_float(r[k4 * 4 + 1]) + bias_s[ch * 32 + k4 * 4 + 1]) * scale;
              v4.z = (__uint_as_float(r[k4 * 4 + 2]) + bias_s[ch * 32 + k4 * 4 + 2]) * scale;
              v4.w = (__uint_as_float(r[k4 * 4 + 3]) + bias_s[ch * 32 + k4 * 4 + 3]) * scale;
              *reinterpret_cast<float4*>(o + k4 * 4) = v4;
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&sh->accfree[n & 1]);
        continue;
      }
      // ---- every other mode: this warp's chunks ch = half, half + 4, ...  Sixteen epilogue warps (four per scheduler)
      //      hide the TMEM / shared / global latencies of one another: with eight, software-pipelined (the TMEM load of the
      //      next chunk in flight under the current one, 168 registers), the launch stayed issue-starved -- ncu: 1.2 warp
      //      instructions per cycle and SM, ~20 k of them and ~17 k cycles per 128 x 256 tile against 7 k of MMAs ----
      uint32_t ra[32];
      uint4 hm[4];                                             // GM_DH: ReLU mask of the NEXT chunk to process (single buffer,
                                                               // reloaded as soon as it has been applied)
      auto load_mask = [&](int ch) {
        if (mode == GM_DH && ch < nch) {
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8)
            hm[g8] = __ldcg(reinterpret_cast<const uint4*>(pr.mask_hi + rowblob + (size_t)(ch * 4 + g8) * 1024));
        }
      };
      // per-tile address bases of the blob stores (64-bit arithmetic once per tile, not once per 8-column slab)
      const int pb = p >> 8, nb8 = min(256, Ppad - pb * 256) >> 3;   // GM_YK: key blocks of <= 256 rows, one after the other
      const size_t yk1 = ((((size_t)b * Ppad + (size_t)pb * 256) * N8) / 8 + ((p & 255) >> 3)) * 64 + (size_t)(p & 7) * 8;   // + n8 * nb8 * 64
      const size_t yk2 = (((size_t)b * (Ppad >> 3) + (p >> 3)) * N8) * 64 + (size_t)(p & 7) * 8;                               // + n8 * 64
      const bool padded = (pr.P & 127) != 0;                   // only then does a tile hold rows past P
      auto chunk = [&](int ch, const uint32_t (&r)[32]) {
        float v[32];
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + ch * 32 + k4 * 4);
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int t4 = 0; t4 < 4; ++t4) {
            float x = __uint_as_float(r[k4 * 4 + t4]) + bb[t4];
            if (mode == GM_H) x = fmaxf(x, 0.f);
            v[k4 * 4 + t4] = x;
          }
        }
        if (padded && !rowok) {                                 // padding rows stay exactly zero
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = 0.f;
        }
        if (mode == GM_DX) {
          // same L2 evict_last policy as the loss kernel's dxT stores (loss_tc.cuh, st_dx): k_wgrad_tc runs between this
          // launch and the dense backward that reads these rows
          const bool keep_dx = g_dx_evict_last != 0;
          uint64_t pol_dx = 0;
          if (keep_dx) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_dx));
          if (rowok && pr.rm) {                               // channels-last maps: row-major rows, one 128-byte line per thread
            float* o = pr.outT + ((size_t)b * Ppad + p) * pr.C + ch * 32;
            if (ch * 32 + 32 <= pr.C && (pr.C & 7) == 0) {
#pragma unroll
              for (int k8 = 0; k8 < 4; ++k8) st_dx8(o + k8 * 8, v + k8 * 8, keep_dx, pol_dx);
            } else if (ch * 32 + 32 <= pr.C && (pr.C & 3) == 0) {
#pragma unroll
              for (int k4 = 0; k4 < 8; ++k4) st_dx4(o + k4 * 4, v[k4 * 4], v[k4 * 4 + 1], v[k4 * 4 + 2], v[k4 * 4 + 3], keep_dx, pol_dx);
            } else {
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (ch * 32 + k < pr.C) st_dx(o + k, v[k], keep_dx, pol_dx);
            }
          } else if (rowok) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int c = ch * 32 + k;
              if (c < pr.C) st_dx(pr.outT + ((size_t)b * pr.C + c) * Ppad + p, v[k], keep_dx, pol_dx);
            }
          }
          return;
        }
        if (mode == GM_DH) {
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const uint32_t wv[4] = {hm[g8].x, hm[g8].y, hm[g8].z, hm[g8].w};
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) {
              if ((wv[k2] & 0x7fffu) == 0u) v[g8 * 8 + 2 * k2] = 0.f;            // relu'(0) = 0
              if ((wv[k2] & 0x7fff0000u) == 0u) v[g8 * 8 + 2 * k2 + 1] = 0.f;
            }
          }
          load_mask(ch + kSub);                                 // flies under the split + stores below and the next TMEM wait
        }
        if (mode == GM_YQ || mode == GM_YK) {
          // a non-finite element makes the sum of squares non-finite; so does a finite row beyond 1.8e19, which the
          // reference would normalise to 0 -- the head's output is O(1), the per-element test is not worth 2 instructions
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            s0 = fmaf(v[k], v[k], s0);
            s1 = fmaf(v[k + 1], v[k + 1], s1);
          }
          const float ssum = s0 + s1;
          pr.ss[((size_t)b * nch + ch) * Ppad + p] = isfinite(ssum) ? ssum : __int_as_float(0x7fc00000);
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
            const size_t o1 = yk1 + (size_t)(n8 * nb8) * 64;
            const size_t o2 = yk2 + (size_t)n8 * 64;
#ifdef PNCE_EXPERIMENTS
            if (g.dbg == 1 && (hi.x | lo.x) != 0x7fc1u) continue;   // experiment: compute everything, store (almost) nothing
#endif
            *reinterpret_cast<uint4*>(pr.k_hi + o1) = hi;
            *reinterpret_cast<uint4*>(pr.k2_hi + o2) = hi;
            if (x3) {
              *reinterpret_cast<uint4*>(pr.k_lo + o1) = lo;
              *reinterpret_cast<uint4*>(pr.k2_lo + o2) = lo;
            }
          } else {
            const size_t o = rowblob + (size_t)n8 * 1024;
#ifdef PNCE_EXPERIMENTS
            if (g.dbg == 1 && (hi.x | lo.x) != 0x7fc1u) continue;
#endif
            *reinterpret_cast<uint4*>(pr.o_hi + o) = hi;
            if (x3) *reinterpret_cast<uint4*>(pr.o_lo + o) = lo;
          }
        }
      };
      load_mask(half);
      // (measured and rejected on top of the 16 warps: software pipelining of the TMEM loads -- a second 32-column buffer
      //  with 8 warps: 214 us per step of GEMM launches; 16-column steps with two 16-register buffers: 203 us; this loop: 201)
#pragma unroll 1
      for (int ch = half; ch < nch; ch += kSub) {
        tmem_ld32(trow + ch * 32, ra);
        tmem_ld_wait();
        chunk(ch, ra);
      }
      tc_fence_before();
      mbar_arrive(&sh->accfree[n & 1]);                       // every epilogue thread arrives: the buffer goes back to the MMA thread
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
  if (tid == 0 && sh->dead && g.err != nullptr) *reinterpret_cast<volatile int*>(g.err) = 1;   // plain store: the flag may live in mapped host memory
}

}  // namespace pnce
