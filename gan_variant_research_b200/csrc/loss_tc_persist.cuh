// Tensor-core path, launch 3, PERSISTENT variant for P <= 256 (one key block): the same math and the
// same operand blobs as k_loss_tc (loss_tc.cuh), but one CTA per SM walks a list of
// (layer, image, 128-row half) items so that consecutive items overlap:
//
//   producer  (warp 0)    P1 loads(n) ........ P2 loads(n) | P1 loads(n+1) ....
//   MMA       (warp 1)       Z(n) = Q K^T ...... dQ(n) = dZ K (chunk by chunk) | Z(n+1) ....
//   epilogue  (warps 2-9)          pass A(n) | pass B(n) -> dZ | dQ epilogue(n) | pass A(n+1) ...
//   norms     (warp 10)   1/||k_j||, ||q_i|| of item n+1 while item n is in flight
//
// * Z(n+1) is issued as soon as pass B(n) has read the last logits chunk (zfree), so the phase-1
//   loads + MMAs of the next item run under the dQ epilogue of the current one; the CTA launch,
//   TMEM allocation and barrier set-up are paid once per SM instead of once per item.
// * the row/key norms (a dependent chain of L2 reads in k_loss_tc's prologue) are computed one item
//   ahead by a warp of their own and handed over through shared memory.
// * EIGHT epilogue warps, two per TMEM lane quadrant (= two per warp scheduler): the pair splits the
//   32-column chunks of its 32 rows even/odd and exchanges the partial row sums through shared
//   memory.  One warp per scheduler (k_loss_tc) leaves the epilogue latency-bound: ~0.25 IPC in the
//   exp2 passes and ~2k cycles per 32-channel chunk of the dQ epilogue (measured with the stamps).
// * the raw q values of the dQ epilogue are prefetched into L2 before pass B.
// Static schedule: CTA c takes items c, c + G, c + 2G, ... of the heavy-first item order (the two
// halves of an image are adjacent items, i.e. run at the same time on neighbouring CTAs and share
// the K blob in L2).  Every mbarrier wait is bounded (umma.cuh): a protocol bug raises the
// library's timeout flag instead of hanging the GPU.
#pragma once
#include "loss_tc.cuh"

namespace pnce {

constexpr int kTpThreads = 352;            // 11 warps: producer, MMA, 8 epilogue, norms
constexpr int kTpSlots2 = 8;               // phase-2 ring: up to 8 slots of (32 keys x Cp channels, hi + lo) in the 96 KB ring
// slots in use for a layer with Cp channels: the narrower the layer, the more key chunks are in flight
// (with 3 slots a C = 64 layer had 24 KB in flight and its dQ MMAs waited on L2 latency)
__host__ __device__ constexpr int tp_slots2(int Cp) { return Cp <= 96 ? 8 : (Cp <= 128 ? 6 : (Cp <= 192 ? 4 : 3)); }
struct TpShared {
  uint64_t full1[kTcSlots1], empty1[kTcSlots1], full2[kTpSlots2], empty2[kTpSlots2];
  uint64_t zfull, zfree, dqfull, dzready[8], normfull, normfree;
  uint32_t tmem_base;
  int dead, flag, badk;
  float red[2][8];
  float xch[128];            // exchange between the two warps of a quadrant (odd-chunk warp -> even-chunk warp -> back):
                             // partial sum of exp2 after pass A, partial sum_j dZ_ij y_ij after pass B
};
constexpr int kTpSharedBytes = 1024;       // 227 KB minus the 1 KB the driver reserves per CTA is the budget
constexpr int kTpSmemBytes = 2 * kTcStageBytes + 2 * kTcDzBytes + 1024 /*invk*/ + kTpSharedBytes;
static_assert(sizeof(TpShared) <= kTpSharedBytes, "TpShared must fit its slot");
static_assert(kTpSmemBytes <= 227 * 1024, "shared memory budget");

struct TpItem {
  const LayerDev* L;
  int b, mh, halves, Ppad, P, C, Cp, Cp8, nstage, nj;
};
// Item order: the slots of the BlockMap run from the heaviest layer to the lightest (launch_loss_tc), ROTATED by
// m.start[PNCE_MAX_LAYERS + 1] items, so that the first third of the CTAs starts on an item of the lightest layers and
// the rest on a heavy one.  All CTAs start in lock-step on a cold L2 (the gather has just streamed 2 GB through it) and
// every code path is slow the first time it runs (first item 42-61 k cycles vs 17-25 k in steady state; first pass-B
// chunk 9 k vs 1.2 k -- per-item stamps, scratch/exp10.py; the kernel is 185 KB of SASS): two kinds of first item spread
// those bursts.  Measured (scratch/exp34.py, B=64): kernel 95 -> 89-91 us for 24-100 light items first, 94 us for a whole
// wave of light items, 97 us for two waves; the makespan model is indifferent to the rotation.
__device__ __forceinline__ void tp_decode(const Params& p, const BlockMap& m, long long item, TpItem& t) {
  const long long total = m.start[p.n_layers];
  item += m.start[PNCE_MAX_LAYERS + 1];
  if (item >= total) item -= total;
  const int slot = find_layer(m, item, p.n_layers);
  const LayerDev& L = p.L[m.layer[slot]];
  const int local = (int)(item - m.start[slot]);
  t.L = &L;
  t.halves = L.Ppad >> 7;
  t.mh = local % t.halves;
  t.b = p.b0 + local / t.halves;
  t.Ppad = L.Ppad; t.P = L.P; t.C = L.C; t.Cp = L.Cp; t.Cp8 = L.Cp >> 3; t.nstage = L.nchunk;
  t.nj = (L.P + 31) >> 5;
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// 1/max(||k_j||, eps) of key j and ||q|| of row gi from the gather's per-chunk partial sums of squares
// (NaN marks a non-finite row).  All loads are issued before the first use.
__device__ __forceinline__ float tp_key_weight(const LayerDev& L, int b, int nstage, int Ppad, int P, int j, bool& bad) {
  float ss[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) ss[s] = (s < nstage && j < P) ? __ldcg(L.kss + ((size_t)b * nstage + s) * Ppad + j) : 0.f;
  float t = 0.f;
#pragma unroll
  for (int s = 0; s < 8; ++s) t += ss[s];
  const float nrm = sqrtf(t);
  bad = !(nrm == nrm) && j < P;
  return (j < P && nrm == nrm) ? 1.0f / fmaxf(nrm, kNormEps) : 0.f;
}
__device__ __forceinline__ void tp_row_ss_load(const Params& p, const BlockMap& m, long long item, int i, float (&ss)[8]) {
  TpItem t;
  tp_decode(p, m, item, t);
  const int gi = t.mh * 128 + i;
#pragma unroll
  for (int s = 0; s < 8; ++s)
    ss[s] = (s < t.nstage && gi < t.P) ? __ldcg(t.L->qss + ((size_t)t.b * t.nstage + s) * t.Ppad + gi) : 0.f;
}

// dQ epilogue of the channel chunks s = half, half + 2, ... of this thread's row (the partner warp of the
// quadrant takes the others); q values are fetched two own-chunks ahead.
template <int NPC>
__device__ __forceinline__ void tp_dq_epilogue(uint32_t tacc, int half, int nstage, int C, const __nv_bfloat16* __restrict__ qh,
                                               const __nv_bfloat16* __restrict__ ql, float* __restrict__ dxrow, float c1,
                                               float c2, bool rowok, __nv_bfloat16* dyh, __nv_bfloat16* dyl,
                                               TcQChunk& qa, TcQChunk& qb) {
  using namespace umma;
  for (int s = half; s < nstage; s += 4) {
    uint32_t r[32];
    tmem_ld32(tacc + s * 32, r);
    tmem_ld_wait();
    tc_dq_chunk<NPC>(r, qa, dxrow + (size_t)s * 32 * NPC, C - s * 32, c1, c2, rowok,
                     dyh ? dyh + (size_t)s * 4096 : nullptr, dyl ? dyl + (size_t)s * 4096 : nullptr, 0);
    tc_q_load(qa, qh, ql, s + 4, nstage);
    if (s + 2 < nstage) {
      tmem_ld32(tacc + (s + 2) * 32, r);
      tmem_ld_wait();
      tc_dq_chunk<NPC>(r, qb, dxrow + (size_t)(s + 2) * 32 * NPC, C - (s + 2) * 32, c1, c2, rowok,
                       dyh ? dyh + (size_t)(s + 2) * 4096 : nullptr, dyl ? dyl + (size_t)(s + 2) * 4096 : nullptr, 0);
      tc_q_load(qb, qh, ql, s + 6, nstage);
    }
  }
}

__global__ void __launch_bounds__(kTpThreads, 1) k_loss_tc_p(const __grid_constant__ Params p,
                                                             const __grid_constant__ BlockMap m) {
  extern __shared__ __align__(1024) unsigned char smem[];
  using namespace umma;
  unsigned char* stage0 = smem;
  unsigned char* dzhi = smem + 2 * kTcStageBytes;
  unsigned char* dzlo = dzhi + kTcDzBytes;
  float* invk_s = reinterpret_cast<float*>(dzlo + kTcDzBytes);
  TpShared* sh = reinterpret_cast<TpShared*>(invk_s + 256);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long total = m.start[p.n_layers];
  const bool x3 = (p.math == PNCE_MATH_TC_BF16X3);
  volatile int* dead = &sh->dead;
  if (p.trace != nullptr && tid == 0) {                       // debug timeline: (sm id, start ns, end ns) per CTA
    unsigned sm; unsigned long long t;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[64 + 3 * (size_t)blockIdx.x] = sm;
    p.trace[64 + 3 * (size_t)blockIdx.x + 1] = (long long)t;
  }
  // debug stamps (pnce_debug_set key 3): clock64 per item and phase for CTA 0 and CTA grid/2, 16 slots per item,
  // at trace[64 + 3 * grid + (sel * 8 + n) * 16 + slot]
  long long* trs = nullptr;
  if (p.trace != nullptr && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2))
    trs = p.trace + 64 + 3 * (size_t)gridDim.x + (blockIdx.x ? 128 : 0);
#ifdef PNCE_EXPERIMENTS
#define PNCE_TS(n_, slot_) do { if (trs && (n_) < 8) trs[(n_) * 16 + (slot_)] = clock64(); } while (0)
#else
#define PNCE_TS(n_, slot_) do { } while (0)
#endif

  if (tid == 0) {
    for (int k = 0; k < kTcSlots1; ++k) { mbar_init(&sh->full1[k], 1); mbar_init(&sh->empty1[k], 1); }
    for (int k = 0; k < kTpSlots2; ++k) { mbar_init(&sh->full2[k], 1); mbar_init(&sh->empty2[k], 1); }
    mbar_init(&sh->zfull, 1); mbar_init(&sh->dqfull, 1); mbar_init(&sh->zfree, 256);
    mbar_init(&sh->normfull, 32); mbar_init(&sh->normfree, 256);
    for (int k = 0; k < 8; ++k) mbar_init(&sh->dzready[k], 128);
    sh->dead = 0;
    sh->badk = 0;
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&sh->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      uint32_t it1 = 0, e2par = 0xffu;                        // e2par bit s: parity to wait for on empty2[s]
      int n = 0;
      bool ok = true;
      for (long long item = blockIdx.x; item < total && ok; item += gridDim.x, ++n) {
        TpItem t;
        tp_decode(p, m, item, t);
        const LayerDev& L = *t.L;
        const size_t qoff = ((size_t)t.b * t.halves + t.mh) * t.Cp8 * 2048;
        const size_t koff = (size_t)t.b * t.Ppad * t.Cp * 2;
        const unsigned char* gq_hi = reinterpret_cast<const unsigned char*>(L.qhi) + qoff;
        const unsigned char* gq_lo = reinterpret_cast<const unsigned char*>(L.qlo) + qoff;
        const unsigned char* gk_hi = reinterpret_cast<const unsigned char*>(L.khi) + koff;
        const unsigned char* gk_lo = reinterpret_cast<const unsigned char*>(L.klo) + koff;
        const unsigned char* g2_hi = reinterpret_cast<const unsigned char*>(L.k2hi) + koff;
        const unsigned char* g2_lo = reinterpret_cast<const unsigned char*>(L.k2lo) + koff;
        const uint32_t kbytes = (uint32_t)t.Ppad * 64u;       // one K chunk: 4 slabs x Ppad/8 core matrices x 128 B
        const uint32_t k2bytes = (uint32_t)t.Cp * 64u;        // 32 keys x Cp channels x 2 B
        // the ring and the dZ region (phase-1 slots 2, 3) are free once the previous item's dQ MMAs are done
        if (n > 0) ok = mbar_wait(&sh->dqfull, (uint32_t)(n - 1) & 1u, dead);
        PNCE_TS(n, 10);
        for (int s = 0; s < t.nstage && ok; ++s, ++it1) {
          const int slot = it1 % kTcSlots1;
          ok = mbar_wait(&sh->empty1[slot], ((it1 / kTcSlots1) & 1u) ^ 1u, dead);
          if (!ok) break;
          unsigned char* st = stage0 + slot * kTcStageBytes;
          mbar_expect_tx(&sh->full1[slot], (8192u + kbytes) * (x3 ? 2u : 1u));
          bulk_g2s(st, gq_hi + (size_t)s * 8192, 8192u, &sh->full1[slot]);
          if (x3) bulk_g2s(st + kTcOffQlo, gq_lo + (size_t)s * 8192, 8192u, &sh->full1[slot]);
          bulk_g2s(st + kTcOffKhi, gk_hi + (size_t)s * kbytes, kbytes, &sh->full1[slot]);
          if (x3) bulk_g2s(st + kTcOffKlo, gk_lo + (size_t)s * kbytes, kbytes, &sh->full1[slot]);
        }
        PNCE_TS(n, 11);
        if (ok) ok = mbar_wait(&sh->zfull, (uint32_t)n & 1u, dead);    // phase-1 MMAs drained: ring changes geometry
        PNCE_TS(n, 12);
        const int ns2 = tp_slots2(t.Cp);
        for (int j = 0; j < t.nj && ok; ++j) {
          const int slot = j % ns2;
          ok = mbar_wait(&sh->empty2[slot], (e2par >> slot) & 1u, dead);
          e2par ^= 1u << slot;
          if (!ok) break;
          unsigned char* st = stage0 + (size_t)slot * 2u * k2bytes;    // slot = hi | lo, k2bytes each
          mbar_expect_tx(&sh->full2[slot], k2bytes * (x3 ? 2u : 1u));
          bulk_g2s(st, g2_hi + (size_t)j * k2bytes, k2bytes, &sh->full2[slot]);
          if (x3) bulk_g2s(st + k2bytes, g2_lo + (size_t)j * k2bytes, k2bytes, &sh->full2[slot]);
        }
        PNCE_TS(n, 13);
        // warm the L2 with the next item's phase-1 operands (its loads can only start when this item's
        // dQ MMAs have drained the ring)
        const long long nxt = item + gridDim.x;
        if (nxt < total && ok) {
          TpItem u;
          tp_decode(p, m, nxt, u);
          const LayerDev& N = *u.L;
          const size_t uq = ((size_t)u.b * u.halves + u.mh) * u.Cp8 * 2048, uk = (size_t)u.b * u.Ppad * u.Cp * 2;
          const uint32_t qb = (uint32_t)u.Cp8 * 2048u, kb = (uint32_t)u.Ppad * (uint32_t)u.Cp * 2u;
          bulk_prefetch_l2(reinterpret_cast<const unsigned char*>(N.qhi) + uq, qb);
          bulk_prefetch_l2(reinterpret_cast<const unsigned char*>(N.khi) + uk, kb);
          if (x3) {
            bulk_prefetch_l2(reinterpret_cast<const unsigned char*>(N.qlo) + uq, qb);
            bulk_prefetch_l2(reinterpret_cast<const unsigned char*>(N.klo) + uk, kb);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t dzh = smem_u32(dzhi), dzl = smem_u32(dzlo);
      uint32_t it1 = 0, f2par = 0, dzphase = 0;               // bit j: parity of the next completion of full2[j] / dzready[j]
      int n = 0;
      bool ok = true;
      for (long long item = blockIdx.x; item < total && ok; item += gridDim.x, ++n) {
        TpItem t;
        tp_decode(p, m, item, t);
        const uint32_t idesc1 = idesc_bf16(128, t.Ppad, 0, 0);
        const uint32_t idesc2 = idesc_bf16(128, t.Cp, 0, 1);  // dQ: B read MN-major (N = channel)
        const uint32_t lbo_k = (uint32_t)t.Ppad * 16u;        // slab (c/8) stride of a K chunk in smem
        const uint32_t lbo2 = (uint32_t)t.Cp8 * 128u;         // 8-key group stride of a key-major chunk
        // the previous logits tile must have been read out of TMEM by pass B
        if (n > 0) ok = mbar_wait(&sh->zfree, (uint32_t)(n - 1) & 1u, dead);
        tc_fence_after();
        PNCE_TS(n, 6);
        // ---- Z = Q_half K^T ----
        for (int s = 0; s < t.nstage && ok; ++s, ++it1) {
          const int slot = it1 % kTcSlots1;
          ok = mbar_wait(&sh->full1[slot], (it1 / kTcSlots1) & 1u, dead);
          tc_fence_after();
          const uint32_t st = smem_u32(stage0 + slot * kTcStageBytes);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {                   // 16 channels = 2 slabs per MMA
            const uint64_t a_hi = smem_desc(st + ks * 4096, 2048, 128);
            const uint64_t b_hi = smem_desc(st + kTcOffKhi + ks * 2 * lbo_k, lbo_k, 128);
            mma_bf16(tmem, a_hi, b_hi, idesc1, (s | ks) ? 1u : 0u);
            if (x3) {
              const uint64_t a_lo = smem_desc(st + kTcOffQlo + ks * 4096, 2048, 128);
              const uint64_t b_lo = smem_desc(st + kTcOffKlo + ks * 2 * lbo_k, lbo_k, 128);
              mma_bf16(tmem, a_hi, b_lo, idesc1, 1u);
              mma_bf16(tmem, a_lo, b_hi, idesc1, 1u);
            }
          }
          mma_commit(&sh->empty1[slot]);
        }
        mma_commit(&sh->zfull);
        PNCE_TS(n, 7);
        // ---- dQ = dZ K, one stage per 32 keys, released chunk by chunk by pass B.  The dQ accumulator
        //      of the previous item has been read: its epilogue precedes this item's pass B in the
        //      epilogue threads' program order (tcgen05 fences on both sides of dzready) ----
        const int ns2 = tp_slots2(t.Cp);
        const uint32_t k2bytes = (uint32_t)t.Cp * 64u;
        for (int j = 0; j < t.nj && ok; ++j) {
          const int slot = j % ns2;
          ok = mbar_wait(&sh->dzready[j], (dzphase >> j) & 1u, dead);
          dzphase ^= 1u << j;
          if (j == 0) PNCE_TS(n, 8);
          if (ok) ok = mbar_wait(&sh->full2[slot], (f2par >> slot) & 1u, dead);
          f2par ^= 1u << slot;
          tc_fence_after();
          const uint32_t st = smem_u32(stage0 + (size_t)slot * 2u * k2bytes);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {                   // 16 keys per MMA
            const uint64_t a_hi = smem_desc(dzh + (uint32_t)(j * 2 + ks) * 4096u, 2048, 128);
            const uint64_t b_hi = smem_desc(st + (uint32_t)ks * 2u * lbo2, lbo2, 128);
            mma_bf16(tmem + 256u, a_hi, b_hi, idesc2, (j | ks) ? 1u : 0u);
            if (x3) {
              const uint64_t a_lo = smem_desc(dzl + (uint32_t)(j * 2 + ks) * 4096u, 2048, 128);
              const uint64_t b_lo = smem_desc(st + k2bytes + (uint32_t)ks * 2u * lbo2, lbo2, 128);
              mma_bf16(tmem + 256u, a_lo, b_hi, idesc2, 1u);
              mma_bf16(tmem + 256u, a_hi, b_lo, idesc2, 1u);
            }
          }
          mma_commit(&sh->empty2[slot]);
        }
        mma_commit(&sh->dqfull);
        PNCE_TS(n, 9);
      }
    }
  } else if (warp == 10) {
    // ===================== key norms of the NEXT item (the first item's are computed by the epilogue warps) ======
    int n = 1;
    for (long long item = (long long)blockIdx.x + gridDim.x; item < total; item += gridDim.x, ++n) {
      TpItem t;
      tp_decode(p, m, item, t);
      const LayerDev& L = *t.L;
      float wk[8];
      bool anybad = false;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        bool bad;
        wk[u] = tp_key_weight(L, t.b, t.nstage, t.Ppad, t.P, lane + 32 * u, bad);
        anybad |= bad;
      }
      anybad = __any_sync(0xffffffffu, anybad);
      // hand-over: the epilogue warps must be done with the previous item's values (end of its pass B)
      mbar_wait(&sh->normfree, (uint32_t)(n - 1) & 1u, dead);
#pragma unroll
      for (int u = 0; u < 8; ++u) invk_s[lane + 32 * u] = wk[u];
      if (lane == 0) sh->badk = anybad ? 1 : 0;
      mbar_arrive(&sh->normfull);
      if (*dead) break;
    }
  } else {
    // ===================== epilogue: two warps per TMEM lane quadrant, thread <-> logits row =====================
    const int q = warp & 3;                                  // TMEM lane quadrant this warp may touch
    const int half = (warp - 2) >> 2;                        // 0: even 32-column chunks, 1: odd ones
    const int i = q * 32 + lane;                             // row inside the item's 128-row half
    const int et = tid - 64;                                 // 0..255
    const float inv_tau = 1.0f / p.tau;
    const float cl = kClamp * kLog2e;
    const bool need_clamp = inv_tau * 1.02f > kClamp;        // |cos| <= 1: the clamp cannot bind otherwise
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t rowoff = (uint32_t)(i >> 3) * 128u + (uint32_t)(i & 7) * 16u;
    if (et != 0) trs = nullptr;
    float ssq[8];                                            // this row's partial sums of squares, loaded one item ahead
    if ((long long)blockIdx.x < total) tp_row_ss_load(p, m, blockIdx.x, i, ssq);
    int n = 0;
    for (long long item = blockIdx.x; item < total; item += gridDim.x, ++n) {
      TpItem t;
      tp_decode(p, m, item, t);
      const LayerDev& L = *t.L;
      const int P = t.P, Ppad = t.Ppad, C = t.C, nstage = t.nstage, b = t.b, mh = t.mh;
      const int gi = mh * 128 + i;                           // sorted patch slot
      const bool rowok = gi < P;
      const uint32_t par = (uint32_t)n & 1u;
      if (n == 0) {
        // cold start: every epilogue thread takes one key so that all norm loads of the first item are in
        // flight at once, next to the first bulk copies
        bool bad;
        const float w = tp_key_weight(L, b, nstage, Ppad, P, et, bad);
        invk_s[et] = w;
        if (bad) sh->badk = 1;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      } else {
        mbar_wait(&sh->normfull, (uint32_t)(n - 1) & 1u, dead);
      }
      PNCE_TS(n, 0);
      float rownrm;
      {
        float tss = 0.f;
#pragma unroll
        for (int s8 = 0; s8 < 8; ++s8) tss += ssq[s8];
        rownrm = sqrtf(tss);                                  // NaN: the gather saw a non-finite element
        if (rowok && half == 0 && L.qinv != nullptr)
          L.qinv[(size_t)b * P + gi] = (rownrm == rownrm) ? (rownrm < kNormEps ? -1.0f / kNormEps : 1.0f / rownrm) : rownrm;
      }
      const bool badq = rowok && !(rownrm == rownrm);
      const float sc = (rowok && !badq) ? 1.0f / fmaxf(rownrm, kNormEps) : 0.f;     // 1 / max(||q_i||, eps)
      const bool noproj = rownrm < kNormEps;
      const bool badrow = badq || (sh->badk != 0);
      const float a = sc * inv_tau * kLog2e;                  // acc -> logit in log2 units (times 1/||k_j||)
      const float coef = rowok ? 1.0f / ((float)P * (float)p.B * (float)p.n_layers) : 0.f;
      const int chd = (gi & 255) >> 5;                        // chunk holding this row's diagonal (warp-uniform)
      const int nch = t.nj;                                   // chunks that hold real columns
      const bool pad = (nch * 32 != P);                       // the last chunk holds padding columns (masked in pass A)
      // ---- pass A: row sum of exp2 over this warp's chunks, diagonal ----
      mbar_wait(&sh->zfull, par, dead);
      tc_fence_after();
      PNCE_TS(n, 1);
      float se4[4] = {0.f, 0.f, 0.f, 0.f};
      float ydacc = 0.f;                                      // raw accumulator of the diagonal element
      for (int ch = half; ch < nch; ch += 2) {
        uint32_t r[32];
        tmem_ld32(trow + ch * 32, r);
        tmem_ld_wait();
        if (pad && ch == nch - 1) {
          if (need_clamp) tc_pass_a_masked<true>(r, invk_s + ch * 32, a, cl, se4);
          else tc_pass_a_masked<false>(r, invk_s + ch * 32, a, cl, se4);
        } else if (need_clamp) tc_pass_a<true>(r, invk_s + ch * 32, a, cl, se4);
        else tc_pass_a<false>(r, invk_s + ch * 32, a, cl, se4);
        if (ch == chd) {
#pragma unroll
          for (int k = 0; k < 32; ++k) ydacc = (k == lane) ? __uint_as_float(r[k]) : ydacc;
        }
      }
      float se = (se4[0] + se4[1]) + (se4[2] + se4[3]);
      if (half == 1) sh->xch[i] = se;
      // raw q for the dQ epilogue comes from this row's slice of the Q operand blob: pulled into L2 now
      // (the lines fly under pass B)
      const size_t qoff = (((size_t)b * t.halves + mh) * t.Cp8 * 16 + (size_t)(i >> 3)) * 64 + (size_t)(i & 7) * 8;
      const __nv_bfloat16* __restrict__ qh = L.qhi + qoff;
      const __nv_bfloat16* __restrict__ ql = (x3 && L.qlo != nullptr) ? L.qlo + qoff : nullptr;
      if ((i & 7) == 0) {                                      // 8 rows share each 128-byte line
        for (int s8 = half; s8 < nstage * 4; s8 += 2) {
          prefetch_l2(qh + (size_t)s8 * 1024);
          if (ql != nullptr) prefetch_l2(ql + (size_t)s8 * 1024);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 0) {
        se += sh->xch[i];
        sh->xch[i] = se;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 1) se = sh->xch[i];
      PNCE_TS(n, 2);
      // next item's row sums of squares: the loads fly under pass B
      if (item + gridDim.x < total) tp_row_ss_load(p, m, item + gridDim.x, i, ssq);
      // the warp that owns the diagonal's chunk (chd) owns the diagonal logit, the row loss and the "- I" fix-up
      const bool own = (chd & 1) == half;
      const float wd = invk_s[gi & 255];
      const float ydr = ydacc * a * wd;                       // unclamped diagonal logit (log2 units); owner only
      const float yd = need_clamp ? fminf(fmaxf(ydr, -cl), cl) : ydr;
      const float lse2 = lg2f(se);
      float rowloss = (rowok && own) ? (lse2 - yd) * kLn2 : 0.f;            // :94, labels = arange
      if (badrow && rowok && own) rowloss = __int_as_float(0x7fc00000);
      // ---- pass B: dZ (pre-divided by ||k_j||) -> smem A operand chunk by chunk, s_i ----
      float s2 = 0.f;
      TcDiag dg;
      dg.rowok = rowok;
      dg.pass = !need_clamp || fabsf(ydr) <= cl;
      dg.d_w = dg.pass ? (ex2f(yd - lse2) - 1.f) * coef * wd : 0.f;
      dg.off = (uint32_t)((gi & 255) >> 3) * 2048u + rowoff + (uint32_t)(gi & 7) * 2u;
      PNCE_TS(n, 14);
      for (int ch = half; ch < nch; ch += 2) {
        uint32_t r[32];
        tmem_ld32(trow + ch * 32, r);
        tmem_ld_wait();
        if (ch == half) PNCE_TS(n, 15);
        if (need_clamp) tc_pass_b<true>(r, invk_s + ch * 32, a, cl, lse2, coef, x3, dzhi, dzlo, (uint32_t)(ch * 4) * 2048u + rowoff, s2);
        else tc_pass_b<false>(r, invk_s + ch * 32, a, cl, lse2, coef, x3, dzhi, dzlo, (uint32_t)(ch * 4) * 2048u + rowoff, s2);
        if (ch == chd) tc_fix_diag(dg, x3, dzhi, dzlo);
        fence_proxy_async_smem();
        mbar_arrive(&sh->dzready[ch]);                         // 128 arrivals release chunk ch to the MMA thread
      }
      PNCE_TS(n, 3);
      tc_fence_before();
      mbar_arrive(&sh->zfree);                                 // the logits tile may be overwritten by the next item
      mbar_arrive(&sh->normfree);                              // and so may invk_s / qnrm / badk
      if (rowok && own && dg.pass) s2 = fmaf(-coef, ydr, s2);  // the diagonal's "- I" term of sum_j dZ_ij y_ij
      if (half == 1) sh->xch[i] = s2;
      // row losses: warp shuffle, then one partial per item (deterministic order)
      rowloss = warp_sum(rowloss);
      if (lane == 0) sh->red[n & 1][half * 4 + q] = rowloss;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et == 0) {
        const float* rd = sh->red[n & 1];
        L.partial[(size_t)b * L.nparts + mh] = ((rd[0] + rd[4]) + (rd[1] + rd[5])) + ((rd[2] + rd[6]) + (rd[3] + rd[7]));
      }
      if (half == 0) {
        s2 += sh->xch[i];
        sh->xch[i] = s2;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 1) s2 = sh->xch[i];
      const float s_i = s2 * kLn2;                              // sum_j dZ_ij z_ij
      // ---- dQ epilogue: dq/tau -> normalise backward -> dxT (coalesced: lane <-> consecutive slot) ----
      //   dx = dq*sc - q_raw * (sc^2 s_i)       with dq = acc / tau;  g / eps when ||q|| < eps
      const float c1 = inv_tau * sc;
      const float c2 = (noproj || p.rows_mode) ? 0.f : sc * sc * s_i;   // rows API: rows are used as given, no projection
      float* __restrict__ dxrow = L.dxT + (size_t)b * C * Ppad + (rowok ? gi : 0);   // dxpitch == Ppad on this path
      __nv_bfloat16* dyh = L.dyhi ? L.dyhi + qoff : nullptr;   // head mode: d loss / d (head output) as a row blob
      __nv_bfloat16* dyl = (L.dyhi && L.dylo) ? L.dylo + qoff : nullptr;
      TcQChunk qa, qb;                                         // this warp's first two channel chunks of raw q:
      tc_q_load(qa, qh, ql, half, nstage);                     // in flight while the dQ MMAs finish
      tc_q_load(qb, qh, ql, half + 2, nstage);
      mbar_wait(&sh->dqfull, par, dead);
      tc_fence_after();
      PNCE_TS(n, 4);
      if (p.rows_mode) {
        // module-split rows API: d loss / d q row-major, rows in the caller's order (the raw q values are not needed)
        float* __restrict__ drow = L.dq_rows + ((size_t)b * P + (rowok ? gi : 0)) * C;
        for (int s = half; s < nstage; s += 2) {
          uint32_t r[32];
          tmem_ld32(trow + 256u + s * 32, r);
          tmem_ld_wait();
          if (rowok) {
            const int nvalid = C - s * 32;
            if (nvalid >= 32 && (C & 3) == 0) {
#pragma unroll
              for (int k4 = 0; k4 < 8; ++k4)
                *reinterpret_cast<float4*>(drow + s * 32 + k4 * 4) =
                    make_float4(__uint_as_float(r[k4 * 4]) * c1, __uint_as_float(r[k4 * 4 + 1]) * c1,
                                __uint_as_float(r[k4 * 4 + 2]) * c1, __uint_as_float(r[k4 * 4 + 3]) * c1);
            } else {
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (k < nvalid) drow[s * 32 + k] = __uint_as_float(r[k]) * c1;
            }
          }
        }
      } else if (Ppad == 128) tp_dq_epilogue<128>(trow + 256u, half, nstage, C, qh, ql, dxrow, c1, c2, rowok, dyh, dyl, qa, qb);
      else tp_dq_epilogue<256>(trow + 256u, half, nstage, C, qh, ql, dxrow, c1, c2, rowok, dyh, dyl, qa, qb);
      tc_fence_before();
      PNCE_TS(n, 5);
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
  if (tid == 0 && sh->dead) {
    p.counter[1] = 1u;                                        // finalize_losses turns the loss into NaN (in-band)
    if (p.nonfinite != nullptr) *reinterpret_cast<volatile int*>(p.nonfinite + 1) = 1;   // protocol timeout flag (may be mapped host memory)
  }
  if (p.trace != nullptr && tid == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[64 + 3 * (size_t)blockIdx.x + 2] = (long long)t;
  }
  last_cta_finalize(p, &sh->flag);
}

}  // namespace pnce
