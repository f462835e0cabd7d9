// Tensor-core path, launch 2: fused logits + diagonal CE forward/backward on tcgen05.
// One CTA per (layer, image, 128-row half of the P x P logits); 192 threads, warp-specialised:
//   warp 0      bulk-copy producer (TMA engine, 1-D, operand blobs are pre-tiled by k_gather_tc)
//   warp 1      TMEM owner + the single MMA-issuing thread
//   warps 2..5  epilogue: one thread per logits row (TMEM lane), tcgen05.ld -> registers
//
//   phase 1   Z(128 x N) = Q_half K^T           kind::f16 bf16, fp32 accumulate in TMEM cols [0,N)
//             streamed over C in chunks of 32 channels through a 2-slot smem ring.
//             bf16x3 mode issues hi*hi + hi*lo + lo*hi into the same accumulator (~2^-17 operand error).
//   epilogue  z_ij = acc * (1/||q_i||)(1/||k_j||)/tau, clamp +-50, row sum of exp, diagonal pick,
//             row loss (warp-shuffle reduced), dZ = (softmax - I) * mask / (P B L),
//             s_i = sum_j dZ_ij z_ij (= q_hat . dq, no second reduction needed),
//             dZ_ij / ||k_j|| split hi/lo -> shared memory as the next MMA's A operand.
//   phase 2   dQ(128 x C) = dZ K               K re-streamed in the same chunks and read MN-major from
//             the very same shared-memory image; accumulate in TMEM cols [256, 256+C)
//   epilogue  dq/tau, normalise backward, dxT[b][c][rank[p]] (unit upstream gradient).
// The logits, softmax and dZ never leave the SM.  Replaces patchnce_cut.py:83-110 and the autograd
// backward of :77-94.  Shapes: P <= 256, C <= 256.
#pragma once
#include "common.cuh"
#include "loss_simt.cuh"
#include "umma.cuh"

namespace pnce {

constexpr int kTcThreads = 192;
constexpr int kTcStageBytes = 49152;        // Qhi 8K | Qlo 8K | Khi 16K | Klo 16K   (32 channels)
constexpr int kTcOffQlo = 8192, kTcOffKhi = 16384, kTcOffKlo = 32768;
constexpr int kTcDzBytes = 65536;           // 128 x 256 bf16
constexpr int kTcSmemBytes = 2 * kTcStageBytes + 2 * kTcDzBytes + 1024 /*invk*/ + 256 /*barriers, misc*/;

struct TcShared {
  uint64_t full[2], empty[2], zfull, dzready, dqfull;
  uint32_t tmem_base;
  int dead;
  int flag;
  float red[4];
};

__global__ void __launch_bounds__(kTcThreads, 1) k_loss_tc(const __grid_constant__ Params p,
                                                           const __grid_constant__ BlockMap m) {
  extern __shared__ __align__(1024) unsigned char smem[];
  using namespace umma;
  unsigned char* stage0 = smem;
  unsigned char* dzhi = smem + 2 * kTcStageBytes;
  unsigned char* dzlo = dzhi + kTcDzBytes;
  float* invk_s = reinterpret_cast<float*>(dzlo + kTcDzBytes);
  TcShared* sh = reinterpret_cast<TcShared*>(invk_s + 256);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int l = find_layer(m, blockIdx.x, p.n_layers);
  const LayerDev& L = p.L[l];
  const int local = (int)(blockIdx.x - m.start[l]);
  const int halves = L.Ppad >> 7;
  const int mh = local % halves, b = local / halves;
  const int N = L.Ppad, P = L.P, C = L.C, Cp8 = L.Cp >> 3, nstage = L.nchunk;
  const bool x3 = (p.math == PNCE_MATH_TC_BF16X3);
  volatile int* dead = &sh->dead;

  if (tid == 0) {
    mbar_init(&sh->full[0], 1); mbar_init(&sh->full[1], 1);
    mbar_init(&sh->empty[0], 1); mbar_init(&sh->empty[1], 1);
    mbar_init(&sh->zfull, 1); mbar_init(&sh->dzready, 128); mbar_init(&sh->dqfull, 1);
    sh->dead = 0;
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&sh->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;
  const uint32_t kbytes = (uint32_t)N * 64u;                 // one K chunk: 4 slabs x N/8 core matrices x 128 B
  const uint32_t lbo_k = (uint32_t)N * 16u;                  // slab (c/8) stride of a K chunk in smem

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      const unsigned char* gq_hi = reinterpret_cast<const unsigned char*>(L.qhi) +
                                   ((size_t)b * halves + mh) * Cp8 * 2048;
      const unsigned char* gq_lo = reinterpret_cast<const unsigned char*>(L.qlo) +
                                   ((size_t)b * halves + mh) * Cp8 * 2048;
      const unsigned char* gk_hi = reinterpret_cast<const unsigned char*>(L.khi) + (size_t)b * Cp8 * (N * 16);
      const unsigned char* gk_lo = reinterpret_cast<const unsigned char*>(L.klo) + (size_t)b * Cp8 * (N * 16);
      uint32_t it = 0;
      for (int ph = 0; ph < 2; ++ph) {
        for (int s = 0; s < nstage; ++s, ++it) {
          const int slot = it & 1;
          const uint32_t par = (it >> 1) & 1u;
          if (!mbar_wait(&sh->empty[slot], par ^ 1u, dead)) break;
          unsigned char* st = stage0 + slot * kTcStageBytes;
          const uint32_t tx = (ph == 0 ? 8192u : 0u) * (x3 ? 2u : 1u) + kbytes * (x3 ? 2u : 1u);
          mbar_expect_tx(&sh->full[slot], tx);
          if (ph == 0) {
            bulk_g2s(st, gq_hi + (size_t)s * 8192, 8192u, &sh->full[slot]);
            if (x3) bulk_g2s(st + kTcOffQlo, gq_lo + (size_t)s * 8192, 8192u, &sh->full[slot]);
          }
          bulk_g2s(st + kTcOffKhi, gk_hi + (size_t)s * kbytes, kbytes, &sh->full[slot]);
          if (x3) bulk_g2s(st + kTcOffKlo, gk_lo + (size_t)s * kbytes, kbytes, &sh->full[slot]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc1 = idesc_bf16(128, N, 0, 0);
      const uint32_t idesc2 = idesc_bf16(128, 32, 0, 1);     // B read MN-major (N = channel)
      uint32_t it = 0;
      bool ok = true;
      // phase 1: Z = Q K^T
      for (int s = 0; s < nstage && ok; ++s, ++it) {
        const int slot = it & 1;
        ok = mbar_wait(&sh->full[slot], (it >> 1) & 1u, dead);
        tc_fence_after();
        const uint32_t st = smem_u32(stage0 + slot * kTcStageBytes);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {                     // 16 channels = 2 slabs per MMA
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
        mma_commit(&sh->empty[slot]);
      }
      mma_commit(&sh->zfull);
      // phase 2: dQ = dZ K
      if (ok) ok = mbar_wait(&sh->dzready, 0u, dead);
      tc_fence_after();
      const uint32_t dzh = smem_u32(dzhi), dzl = smem_u32(dzlo);
      for (int s = 0; s < nstage && ok; ++s, ++it) {
        const int slot = it & 1;
        ok = mbar_wait(&sh->full[slot], (it >> 1) & 1u, dead);
        tc_fence_after();
        const uint32_t st = smem_u32(stage0 + slot * kTcStageBytes);
        const uint32_t d = tmem + 256u + (uint32_t)s * 32u;
        const int ksteps = N >> 4;
        for (int ks = 0; ks < ksteps; ++ks) {                // 16 key rows j per MMA
          const uint64_t a_hi = smem_desc(dzh + ks * 4096, 2048, 128);
          const uint64_t b_hi = smem_desc(st + kTcOffKhi + ks * 256, 128, lbo_k);
          mma_bf16(d, a_hi, b_hi, idesc2, ks ? 1u : 0u);
          if (x3) {
            const uint64_t a_lo = smem_desc(dzl + ks * 4096, 2048, 128);
            const uint64_t b_lo = smem_desc(st + kTcOffKlo + ks * 256, 128, lbo_k);
            mma_bf16(d, a_lo, b_hi, idesc2, 1u);
            mma_bf16(d, a_hi, b_lo, idesc2, 1u);
          }
        }
        mma_commit(&sh->empty[slot]);
      }
      mma_commit(&sh->dqfull);
    }
  } else {
    // ===================== epilogue: thread <-> logits row =====================
    const int q = warp & 3;                                  // TMEM lane quadrant this warp may touch
    const int i = q * 32 + lane;                             // row inside the half
    const int gi = mh * 128 + i;                             // patch index
    const bool rowok = gi < P;
    const int et = tid - 64;                                 // 0..127
    // 1/||k_j|| for the N key rows, 1/||q_i|| for this row
    for (int j = et; j < N; j += 128) {
      float ss = 0.f;
      if (j < P)
        for (int s = 0; s < nstage; ++s) ss += L.kss[((size_t)b * nstage + s) * N + j];
      const float nrm = sqrtf(ss);
      float inv = 1.0f / fmaxf(nrm, kNormEps);
      if (!(nrm == nrm)) inv = nrm;
      invk_s[j] = (j < P) ? inv : 0.f;
    }
    float sc = 0.f;                                           // 1 / max(||q_i||, eps)
    bool noproj = false;
    if (rowok) {
      float ss = 0.f;
      for (int s = 0; s < nstage; ++s) ss += L.qss[((size_t)b * nstage + s) * N + gi];
      const float nrm = sqrtf(ss);
      sc = 1.0f / fmaxf(nrm, kNormEps);
      if (!(nrm == nrm)) sc = nrm;
      noproj = nrm < kNormEps;
      L.qinv[(size_t)b * P + gi] = (nrm == nrm) ? (noproj ? -1.0f / kNormEps : 1.0f / nrm) : nrm;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const float inv_tau = 1.0f / p.tau;
    const float zscale = sc * inv_tau;
    const float coef = 1.0f / ((float)P * (float)p.B * (float)p.n_layers);
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    bool ok = mbar_wait(&sh->zfull, 0u, dead);
    tc_fence_after();
    // ---- pass A: row sum of exp, diagonal ----
    float se = 0.f, zd = 0.f;
    const int nch = N >> 5;
    for (int ch = 0; ch < nch; ++ch) {
      uint32_t r[32];
      tmem_ld32(trow + ch * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int j = ch * 32 + k;
        if (j < P) {
          const float zc = clamp_nan(__uint_as_float(r[k]) * zscale * invk_s[j], kClamp);   // :85, :88
          se += __expf(zc);
          if (j == gi) zd = zc;
        }
      }
    }
    const float lse = logf(se);
    float rowloss = rowok ? (lse - zd) : 0.f;                 // :94, labels = arange
    // ---- pass B: dZ (pre-divided by ||k_j||) -> smem A operand, s_i ----
    float s_i = 0.f;
    for (int ch = 0; ch < nch; ++ch) {
      uint32_t r[32];
      tmem_ld32(trow + ch * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
          float dd[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k = g8 * 8 + k2 * 2 + h;
            const int j = ch * 32 + k;
            float d = 0.f;
            if (rowok && j < P) {
              const float ik = invk_s[j];
              const float zraw = __uint_as_float(r[k]) * zscale * ik;
              const float pj = __expf(clamp_nan(zraw, kClamp) - lse);
              const bool pass = (zraw >= -kClamp) && (zraw <= kClamp);
              d = pass ? (pj - (j == gi ? 1.f : 0.f)) * coef : 0.f;
              s_i = fmaf(d, zraw, s_i);
              d *= ik;
            }
            dd[h] = d;
          }
          const __nv_bfloat16 ah = __float2bfloat16_rn(dd[0]), bh = __float2bfloat16_rn(dd[1]);
          hw[k2] = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
          const __nv_bfloat16 al = __float2bfloat16_rn(dd[0] - __bfloat162float(ah));
          const __nv_bfloat16 bl = __float2bfloat16_rn(dd[1] - __bfloat162float(bh));
          lw[k2] = (uint32_t)__bfloat16_as_ushort(al) | ((uint32_t)__bfloat16_as_ushort(bl) << 16);
        }
        const int j8 = ch * 4 + g8;
        const uint32_t off = (uint32_t)(j8 * 16 + (i >> 3)) * 128u + (uint32_t)(i & 7) * 16u;
        *reinterpret_cast<uint4*>(dzhi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        if (x3) *reinterpret_cast<uint4*>(dzlo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(&sh->dzready);
    // row losses: warp shuffle, then one partial per CTA (deterministic order)
    rowloss = warp_sum(rowloss);
    if (lane == 0) sh->red[q] = rowloss;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (et == 0) L.partial[(size_t)b * L.nparts + mh] = sh->red[0] + sh->red[1] + sh->red[2] + sh->red[3];
    // ---- dQ epilogue: dq/tau -> normalise backward -> dxT ----
    if (ok) ok = mbar_wait(&sh->dqfull, 0u, dead);
    tc_fence_after();
    const int slot_out = rowok ? L.rank[gi] : 0;
    const float* qrow = L.qT + (size_t)b * C * N + gi;
    float* dxrow = L.dxT + (size_t)b * C * P + slot_out;
    for (int s = 0; s < nstage; ++s) {
      uint32_t r[32];
      tmem_ld32(trow + 256u + s * 32, r);
      tmem_ld_wait();
      if (rowok) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const int c = s * 32 + k;
          if (c < C) {
            const float dq = __uint_as_float(r[k]) * inv_tau;
            const float qh = qrow[(size_t)c * N] * sc;
            // F.normalize backward: (g - x^(x^.g)) / n when n >= eps, else g / eps
            dxrow[(size_t)c * P] = noproj ? dq * sc : (dq - qh * s_i) * sc;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
  if (tid == 0 && sh->dead && p.nonfinite != nullptr) atomicExch(p.nonfinite + 1, 1);   // protocol timeout flag
  last_cta_finalize(p, &sh->flag);
}

// -------------------------------------------------------------------------------------------------
// Self-test kernel: D(128 x N) = A(128 x K) * B, operands given as pre-tiled blobs with the
// descriptor parameters supplied by the host.  Exercises bulk copy, mbarrier tx, TMEM alloc,
// tcgen05.mma (K-major or MN-major B), commit and tcgen05.ld exactly as k_loss_tc uses them.
// -------------------------------------------------------------------------------------------------
struct ProbeArgs {
  const void* a_blob; const void* b_blob; float* d_out; int* err;
  uint32_t a_bytes, b_bytes;
  uint32_t a_lbo, a_sbo, a_kstep;      // bytes; a_kstep = start-address advance per K=16 step
  uint32_t b_lbo, b_sbo, b_kstep;
  int n, k, b_mn_major;
};

__global__ void __launch_bounds__(128, 1) k_umma_probe(const __grid_constant__ ProbeArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  using namespace umma;
  __shared__ uint64_t full, done;
  __shared__ uint32_t tmem_base;
  __shared__ int dead;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* sa = smem;
  unsigned char* sb = smem + ((a.a_bytes + 1023u) & ~1023u);
  if (tid == 0) {
    mbar_init(&full, 1); mbar_init(&done, 1);
    dead = 0;
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(&tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    mbar_expect_tx(&full, a.a_bytes + a.b_bytes);
    bulk_g2s(sa, a.a_blob, a.a_bytes, &full);
    bulk_g2s(sb, a.b_blob, a.b_bytes, &full);
    if (mbar_wait(&full, 0u, &dead)) {
      tc_fence_after();
      const uint32_t idesc = idesc_bf16(128, a.n, 0, a.b_mn_major);
      for (int ks = 0; ks < a.k / 16; ++ks) {
        const uint64_t da = smem_desc(smem_u32(sa) + ks * a.a_kstep, a.a_lbo, a.a_sbo);
        const uint64_t db = smem_desc(smem_u32(sb) + ks * a.b_kstep, a.b_lbo, a.b_sbo);
        mma_bf16(tmem, da, db, idesc, ks ? 1u : 0u);
      }
    }
    mma_commit(&done);
  }
  __syncwarp();
  const bool ok = mbar_wait(&done, 0u, &dead);
  tc_fence_after();
  for (int ch = 0; ch < a.n / 32; ++ch) {
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + ch * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 32; ++k) a.d_out[(size_t)(warp * 32 + lane) * a.n + ch * 32 + k] = __uint_as_float(r[k]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem);
  }
  if (tid == 0 && (!ok || dead)) *a.err = 1;
}

}  // namespace pnce
