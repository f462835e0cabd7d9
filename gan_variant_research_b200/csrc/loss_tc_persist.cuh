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
  float xch[3][128];         // exchange between the two warps of a quadrant (odd-chunk warp -> even-chunk warp -> back):
                             // [0] partial sum of exp2, [1] partial sum_j e_ij y_ij (single pass) or sum_j dZ_ij y_ij
                             // (two passes), [2] the diagonal logit (known to the warp that owns the diagonal's chunk)
};
constexpr int kTpSharedBytes = 2048;       // the kernel has no static shared memory: all 227 KB are dynamic
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


// ---- the epilogue warps' inner loops, written COMPACT on purpose ------------------------------------------------
// The first version of this kernel unrolled every pass over a whole 32-column chunk, twice (register double buffers)
// and per template variant: 244 KB of SASS, ~50 KB of it executed per item by eleven warps in four roles -- far more
// than the 32 KB instruction cache level behind the schedulers' 6 KB L0s.  ncu: "no instruction" was the second
// largest stall of the kernel (1.5 warp-cycles per issued instruction) and phase stamps showed a few dozen
// instructions between two barriers taking 2-3 k cycles.  The loops below walk a row in steps of EIGHT columns
// (tcgen05.ld 32x32b.x8, two register buffers so that the load of step n+1 flies under the arithmetic of step n) and
// stay rolled: the body of a pass is ~250 instructions, that of the dQ epilogue ~400.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// Key-chunk order of an item.  Single-pass softmax (below) defers the release of the four chunks that hold the
// diagonals of the item's own 128 rows until the row sums are known, so those chunks go LAST: with two row halves
// the first half (rows 0..127, diagonals in chunks 0..3) visits 4..7 first; the second half's natural order already
// ends on its diagonals (4..7).  nd = number of chunks visited before chunk 0 (0: natural order).
__host__ __device__ __forceinline__ int tp_rot_count(bool rot, int nj) { return (rot && nj > 4) ? nj - 4 : 0; }
__device__ __forceinline__ int tp_chunk_all(int k, int nd) { return k < nd ? 4 + k : k - nd; }
// the same order restricted to the chunks of one parity (an epilogue warp takes every other chunk)
__device__ __forceinline__ int tp_chunk_at(int k, int half, int hi_cnt) {
  return k < hi_cnt ? 4 + half + 2 * k : half + 2 * (k - hi_cnt);
}

enum { kPassSingle = 0, kPassA = 1, kPassB = 2 };
struct TpPassArgs {
  const float* invk_s;       // 1 / ||k_j|| (0 for padding / non-finite keys)
  unsigned char *dzhi, *dzlo;
  uint64_t* dzready;
  float a, cl, lse2, coef;
  uint32_t rowoff;
  int chd, lane;
  bool x3, defer;            // defer: the warp that owns the diagonal's chunk releases it later (single pass)
};

// One step = 8 logits of this thread's row (chunk c, columns g8 * 8 ...).
//   kPassSingle  e = exp2(y) ONCE: unnormalised e / ||k_j|| -> dZ operand (bf16 hi + lo), se += e, s2 += e y.  No running
//                maximum is needed when the clamp cannot bind (|y| <= log2(e) / tau <= 71: exp2 and a sum of 1024 of
//                them fit fp32); the row factor coef / se commutes with dQ' = E K and is applied in the dQ epilogue,
//                the "- I" term is patched into the operand by the diagonal's owner once se is known.
//   kPassA / B   the two-pass form for temperatures where the +-50 clamp can bind (patchnce_cut.py:88): A = row sum
//                of exp2 of the clamped logits, B = dZ = (softmax - I) coef with the clamp's backward mask.
// Padding columns (w == 0 -> y = 0 -> exp2 = 1) are kept out of se by a select: subtracting their count cancels
// catastrophically when every real logit is far below zero (scratch/stress.py).  Replaces patchnce_cut.py:85-94 + autograd.
template <int MODE>
__device__ __forceinline__ void tp_pass_step(const uint32_t (&r)[8], int c, int g8, const TpPassArgs& A, float& se0,
                                             float& se1, float& s2, float& ydacc, const TcDiag& dg) {
  const float* wk = A.invk_s + c * 32 + g8 * 8;
  const float4 w0 = *reinterpret_cast<const float4*>(wk), w1 = *reinterpret_cast<const float4*>(wk + 4);
  const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
  float dd[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const float y = __uint_as_float(r[t]) * A.a * ww[t];
    if (MODE == kPassSingle) {
      const float e = ex2f(y);
      if (t & 1) se1 += (ww[t] != 0.f) ? e : 0.f;
      else se0 += (ww[t] != 0.f) ? e : 0.f;
      s2 = fmaf(e, y, s2);
      dd[t] = e * ww[t];
    } else if (MODE == kPassA) {
      const float e = ex2f(fminf(fmaxf(y, -A.cl), A.cl));
      if (t & 1) se1 += (ww[t] != 0.f) ? e : 0.f;
      else se0 += (ww[t] != 0.f) ? e : 0.f;
    } else {
      const float yc = fminf(fmaxf(y, -A.cl), A.cl);
      float d = ex2f(yc - A.lse2) * A.coef;
      d = (fabsf(y) <= A.cl) ? d : 0.f;                         // clamp backward mask (inclusive)
      s2 = fmaf(d, y, s2);
      dd[t] = d * ww[t];
    }
  }
  if (MODE != kPassA) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
      hw[k2] = bf16x2_bits(dd[2 * k2], dd[2 * k2 + 1]);
      lw[k2] = bf16x2_bits(dd[2 * k2] - __uint_as_float(hw[k2] << 16), dd[2 * k2 + 1] - __uint_as_float(hw[k2] & 0xffff0000u));
    }
    const uint32_t off = (uint32_t)(c * 4 + g8) * 2048u + A.rowoff;       // 8-column slab of the A operand
    *reinterpret_cast<uint4*>(A.dzhi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    if (A.x3) *reinterpret_cast<uint4*>(A.dzlo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
  if (MODE != kPassB && c == A.chd) {                           // raw accumulator of the row's diagonal element
    const int dsel = (g8 == (A.lane >> 3)) ? (A.lane & 7) : -1;
#pragma unroll
    for (int t = 0; t < 8; ++t) ydacc = (t == dsel) ? __uint_as_float(r[t]) : ydacc;
  }
  if (MODE != kPassA && g8 == 3) {                              // chunk complete: hand it to the MMA thread
    if (MODE == kPassB && c == A.chd) tc_fix_diag(dg, A.x3, A.dzhi, A.dzlo);
    umma::fence_proxy_async_smem();
    if (!(A.defer && c == A.chd)) umma::mbar_arrive(&A.dzready[c]);       // 128 arrivals release chunk c
  }
}

// nmine chunks of this warp's parity, in the item's chunk order; trow = TMEM address of the logits row block.
template <int MODE>
__device__ __forceinline__ void tp_pass(uint32_t trow, int nmine, int half, int hi_cnt, const TpPassArgs& A, float& se0,
                                        float& se1, float& s2, float& ydacc, const TcDiag& dg) {
  using namespace umma;
  const int nsteps = nmine * 4;
  if (nsteps == 0) return;
  uint32_t ra[8], rb[8];
  tmem_ld8(trow + tp_chunk_at(0, half, hi_cnt) * 32, ra);
#pragma unroll 1
  for (int st = 0; st < nsteps; st += 2) {
    const int c = tp_chunk_at(st >> 2, half, hi_cnt), g8 = st & 3;         // g8 = 0 or 2
    tmem_ld_wait();
    tmem_ld8(trow + c * 32 + (g8 + 1) * 8, rb);
    tp_pass_step<MODE>(ra, c, g8, A, se0, se1, s2, ydacc, dg);
    tmem_ld_wait();
    if (st + 2 < nsteps) tmem_ld8(trow + tp_chunk_at((st + 2) >> 2, half, hi_cnt) * 32 + ((st + 2) & 3) * 8, ra);
    tp_pass_step<MODE>(rb, c, g8 + 1, A, se0, se1, s2, ydacc, dg);
  }
}

__global__ void __launch_bounds__(kTpThreads, 1) k_loss_tc_p(const __grid_constant__ Params p,
                                                             const __grid_constant__ BlockMap m) {
  pdl_launch();
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
  const bool single_pass = !((1.0f / p.tau) * 1.02f > kClamp);  // |cos| <= 1: the +-50 clamp cannot bind (every CUT temperature)
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
    trs = p.trace + 64 + 3 * (size_t)gridDim.x + (blockIdx.x ? 256 : 0);
#ifdef PNCE_EXPERIMENTS
#define PNCE_TS(n_, slot_) do { if (trs && (n_) < 8) trs[(n_) * 32 + (slot_)] = clock64(); } while (0)
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
  pdl_wait();                                                 // prologue done: now wait for the prerequisite grids
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
        const int nd = tp_rot_count(single_pass && t.halves == 2 && t.mh == 0, t.nj);
        for (int k = 0; k < t.nj && ok; ++k) {
          const int slot = k % ns2, j = tp_chunk_all(k, nd);  // the MMA thread's chunk order
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
        const int nd = tp_rot_count(single_pass && t.halves == 2 && t.mh == 0, t.nj);
        for (int k = 0; k < t.nj && ok; ++k) {
          const int slot = k % ns2, j = tp_chunk_all(k, nd);  // diagonal chunks last (released after the row sums)
          ok = mbar_wait(&sh->dzready[j], (dzphase >> j) & 1u, dead);
          dzphase ^= 1u << j;
          if (k == 0) PNCE_TS(n, 8);
          if (ok) ok = mbar_wait(&sh->full2[slot], (f2par >> slot) & 1u, dead);
          f2par ^= 1u << slot;
          tc_fence_after();
          const uint32_t st = smem_u32(stage0 + (size_t)slot * 2u * k2bytes);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {                   // 16 keys per MMA
            const uint64_t a_hi = smem_desc(dzh + (uint32_t)(j * 2 + ks) * 4096u, 2048, 128);
            const uint64_t b_hi = smem_desc(st + (uint32_t)ks * 2u * lbo2, lbo2, 128);
            mma_bf16(tmem + 256u, a_hi, b_hi, idesc2, (k | ks) ? 1u : 0u);
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
    const bool need_clamp = !single_pass;
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
      const bool single = !need_clamp;
      const bool rot = single && t.halves == 2 && mh == 0;    // chunk order of the item: tp_rot_count
      const int nmine = (nch - half + 1) >> 1;                // chunks of this warp's parity
      const int hi_cnt = (rot && nch > 4 + half) ? (nch - 4 - half + 1) >> 1 : 0;
      // raw q of this row for the dQ epilogue comes from its slice of the Q operand blob
      const size_t qoff = (((size_t)b * t.halves + mh) * t.Cp8 * 16 + (size_t)(i >> 3)) * 64 + (size_t)(i & 7) * 8;
      const __nv_bfloat16* __restrict__ qh = L.qhi + qoff;
      const __nv_bfloat16* __restrict__ ql = (x3 && L.qlo != nullptr) ? L.qlo + qoff : nullptr;
      // the warp that owns the diagonal's chunk owns the diagonal logit, the row loss and the "- I" fix-up
      const bool own = (chd & 1) == half && chd < nch;
      const float wd = invk_s[gi & 255];                      // 1 / ||k_i|| of the row's own key
      TpPassArgs A;
      A.invk_s = invk_s; A.dzhi = dzhi; A.dzlo = dzlo; A.dzready = sh->dzready;
      A.a = a; A.cl = cl; A.lse2 = 0.f; A.coef = coef; A.rowoff = rowoff; A.chd = chd; A.lane = lane; A.x3 = x3;
      A.defer = single;
      TcDiag dg;
      dg.rowok = rowok; dg.pass = true; dg.d_w = 0.f;
      dg.off = (uint32_t)((gi & 255) >> 3) * 2048u + rowoff + (uint32_t)(gi & 7) * 2u;
      mbar_wait(&sh->zfull, par, dead);
      tc_fence_after();
      PNCE_TS(n, 1);
      if ((i & 7) == 0) {                                      // 8 rows share each 128-byte line: pull raw q into L2
        for (int s8 = half; s8 < nstage * 4; s8 += 2) {        // (the lines fly under the pass)
          prefetch_l2(qh + (size_t)s8 * 1024);
          if (ql != nullptr) prefetch_l2(ql + (size_t)s8 * 1024);
        }
      }
      float se0 = 0.f, se1 = 0.f, s2 = 0.f, ydacc = 0.f;
      if (single) {
        // ---- ONE pass: exp2 once, unnormalised operand, row sums alongside ----
        tp_pass<kPassSingle>(trow, nmine, half, hi_cnt, A, se0, se1, s2, ydacc, dg);
        PNCE_TS(n, 3);
        tc_fence_before();
        mbar_arrive(&sh->zfree);                               // the logits tile may be overwritten by the next item
        mbar_arrive(&sh->normfree);                            // and so may invk_s / badk (wd, badrow were read above)
      } else {
        // ---- pass A: row sum of exp2 of the clamped logits ----
        tp_pass<kPassA>(trow, nmine, half, 0, A, se0, se1, s2, ydacc, dg);
      }
      float se = se0 + se1;
      if (half == 1) { sh->xch[0][i] = se; sh->xch[1][i] = s2; }
      if (own) sh->xch[2][i] = ydacc * a * wd;                 // diagonal logit (log2 units), unclamped
      // next item's row sums of squares: the loads fly under the exchange and what follows
      if (item + gridDim.x < total) tp_row_ss_load(p, m, item + gridDim.x, i, ssq);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 0) {
        se += sh->xch[0][i]; sh->xch[0][i] = se;
        if (single) { s2 += sh->xch[1][i]; sh->xch[1][i] = s2; }
      }
      const float ydr = sh->xch[2][i];
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 1) { se = sh->xch[0][i]; if (single) s2 = sh->xch[1][i]; }
      PNCE_TS(n, 2);
      const float lse2 = lg2f(se);
      const float yd = single ? ydr : fminf(fmaxf(ydr, -cl), cl);
      float rowloss = (rowok && own) ? (lse2 - yd) * kLn2 : 0.f;            // :94, labels = arange
      if (badrow && rowok && own) rowloss = __int_as_float(0x7fc00000);
      if (single) {
        if (own) {
          // the "- I" term: dQ' = sum_j e_ij k^_j - se_i k^_i, i.e. the diagonal entry of the operand becomes
          // (e_ii - se_i) / ||k_i||; this warp's 32 arrivals then release the chunk
          dg.d_w = (ex2f(ydr) - se) * wd;
          tc_fix_diag(dg, x3, dzhi, dzlo);
          fence_proxy_async_smem();
          mbar_arrive(&sh->dzready[chd]);
        }
        s2 = coef * (s2 / se - ydr);                            // sum_j dZ_ij y_ij with dZ = coef (e / se - I)
      } else {
        // ---- pass B: dZ (pre-divided by ||k_j||) -> smem A operand chunk by chunk, s_i ----
        dg.pass = fabsf(ydr) <= cl;
        dg.d_w = dg.pass ? (ex2f(yd - lse2) - 1.f) * coef * wd : 0.f;
        A.lse2 = lse2;
        PNCE_TS(n, 14);
        s2 = 0.f;
        tp_pass<kPassB>(trow, nmine, half, 0, A, se0, se1, s2, ydacc, dg);
        PNCE_TS(n, 3);
        tc_fence_before();
        mbar_arrive(&sh->zfree);
        mbar_arrive(&sh->normfree);
        if (rowok && own && dg.pass) s2 = fmaf(-coef, ydr, s2);  // the diagonal's "- I" term of sum_j dZ_ij y_ij
        if (half == 1) sh->xch[1][i] = s2;
      }
      // row losses: warp shuffle, then one partial per item (deterministic order)
      rowloss = warp_sum(rowloss);
      if (lane == 0) sh->red[n & 1][half * 4 + q] = rowloss;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et == 0) {
        const float* rd = sh->red[n & 1];
        L.partial[(size_t)b * L.nparts + mh] = ((rd[0] + rd[4]) + (rd[1] + rd[5])) + ((rd[2] + rd[6]) + (rd[3] + rd[7]));
      }
      if (!single) {
        if (half == 0) { s2 += sh->xch[1][i]; sh->xch[1][i] = s2; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (half == 1) s2 = sh->xch[1][i];
      }
      const float s_i = s2 * kLn2;                              // sum_j dZ_ij z_ij
      // ---- dQ epilogue: dq/tau -> normalise backward -> dxT (coalesced: lane <-> consecutive slot) ----
      //   dx = dq*sc - q_raw * (sc^2 s_i)       with dq = acc / tau;  g / eps when ||q|| < eps
      //   single pass: acc holds the unnormalised sum, dq = (coef / se) acc / tau
      const float c1 = single ? inv_tau * sc * (coef / se) : inv_tau * sc;
      const float c2 = (noproj || p.rows_mode) ? 0.f : sc * sc * s_i;   // rows API: rows are used as given, no projection
      float* __restrict__ dxrow = L.dxT + (size_t)b * C * Ppad + (rowok ? gi : 0);   // dxpitch == Ppad on this path
      const bool rm = p.nhwc != 0;                             // channels-last maps: row-major rows [slot][C]
      if (rm) dxrow = L.dxT + ((size_t)b * Ppad + (rowok ? gi : 0)) * C;
      __nv_bfloat16* dyh = L.dyhi ? L.dyhi + qoff : nullptr;   // head mode: d loss / d (head output) as a row blob
      __nv_bfloat16* dyl = (L.dyhi && L.dylo) ? L.dylo + qoff : nullptr;
      TcQChunk qn;                                             // this warp's first channel chunk of raw q: in flight
      if (!p.rows_mode) tc_q_load(qn, qh, ql, half, nstage);   // while the dQ MMAs finish
      mbar_wait(&sh->dqfull, par, dead);
      tc_fence_after();
      PNCE_TS(n, 4);
      if (p.rows_mode) {
        // module-split rows API: d loss / d q row-major, rows in the caller's order (the raw q values are not needed)
        float* __restrict__ drow = L.dq_rows + ((size_t)b * P + (rowok ? gi : 0)) * C;
#pragma unroll 1
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
      } else {
        // this warp's channel chunks s = half, half + 2, ... (the partner warp of the quadrant takes the others);
        // raw q is fetched one own-chunk ahead
#pragma unroll 1
        for (int s = half; s < nstage; s += 2) {
          uint32_t r[32];
          tmem_ld32(trow + 256u + s * 32, r);
          const TcQChunk qc = qn;
          tc_q_load(qn, qh, ql, s + 2, nstage);
          tmem_ld_wait();
          tc_dq_chunk<0>(r, qc, dxrow + (rm ? (size_t)s * 32 : (size_t)s * 32 * Ppad), C - s * 32, c1, c2, rowok,
                         dyh ? dyh + (size_t)s * 4096 : nullptr, dyl ? dyl + (size_t)s * 4096 : nullptr, rm ? C : Ppad, rm);
        }
      }
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
  last_cta_finalize(p, &sh->flag, smem);                       // every CTA-wide phase is over (the __syncthreads above): the ring is free scratch
}

}  // namespace pnce
