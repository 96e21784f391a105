// K1 + K2, fast mode, DECOUPLED cta_group::2 variant (OMBO_FAST_MODE=4) -- see the banner below and
// DESIGN.md section 4.  Same contract as k_posterior_fast (posterior_fast.cu).
#include "umma.cuh"

// =================================================================================================
// Decoupled cta_group::2 kernel (the default).
//
// What bounds the single-CTA variants above is shared-memory bandwidth: per 256-column unit the UMMAs
// read 48 KB of A and 96 KB of B, TMA writes 64 KB of B and the generators 32 KB of A -- more than the
// 128 B/clk an SM's shared memory delivers in the 1536 cycles the tensor core needs (measured with
// scripts/umma_bench.cu: the UMMA itself runs at N/2 cycles, SS or TS).  A CTA pair halves the B traffic
// per SM (each CTA holds half of every B tile), which leaves the tensor core as the bound -- provided
// the K1 generators never sit on the MMA's critical path.  So here they do not touch the operand ring:
//
//   generators (warps 8-15)  K* block kb of the tile -> bf16 hi/lo -> straight into the per-CTA L2
//                            cache (STG, already in the UMMA shared-memory image), free-running up to a
//                            whole tile ahead; they only wait for "slot kb was last read" (`landed`)
//   A feeder   (warp 3)      for every (pass, kb) in MMA order: wait `produced`, wait the stage, 2-SM
//                            tensor load of the cached 32 KB stage (bytes counted on the leader)
//   B producer (warp 0)      2-SM tensor loads of this CTA's 128 rows of the 256-row B tile
//   MMA issuer (warp 1 of the leader)  12 x tcgen05.mma.cta_group::2 (M 256, N 256) per unit; publishes
//                            `landed` (fills whose data reached shared memory) to both CTAs
//   epilogue   (warps 4-7)   as above
// =================================================================================================
__device__ __forceinline__ uint32_t ld_volatile_smem(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.volatile.shared::cta.u32 %0, [%1];\n" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_smem(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];\n" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_smem(uint32_t *p, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;\n" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.relaxed.cluster.shared::cluster.u32 [%0], %1;\n" ::"r"(addr), "r"(v) : "memory");
}

template <int DP, int R, bool F16>
__global__ void __launch_bounds__(16 * 32, 1)
k_posterior_fast_dc(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                    const __grid_constant__ CUtensorMap map_kc, const FastParams prm) {
  constexpr int NSTA = 3, NSTB = 3;
  constexpr int GEN_THREADS = 256;
  constexpr int CW = 256, NSLOT = 2, KSH = 2;
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(CW >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d = prm.gp.d, np = prm.gp.n_pad;
  const int nkb = np / FK;
  const int n_chunks = (np + CW - 1) / CW;
  const int n_pass = (n_chunks + NSLOT - 1) / NSLOT;
  const long long n_tiles = (prm.m + FM - 1) / FM;
  const long long n_iter = (n_tiles + gridDim.x - 1) / gridDim.x;
  const uint32_t crank = cluster_rank();
  const bool leader = crank == 0;
  auto last_kb = [&](int c) { return min(((c + 1) << KSH) - 1, nkb - 1); };
  auto pass_kb_end = [&](int p) { return last_kb(min(NSLOT * p + NSLOT - 1, n_chunks - 1)) + 1; };

  unsigned char *sA = fast_smem + ((1024u - (smem_u32(fast_smem) & 1023u)) & 1023u);
  unsigned char *sB = sA + NSTA * STAGE_BYTES;
  float *xc = (float *)(sB + NSTB * STAGE_BYTES);
  float *xt = xc + (size_t)DP * FM;
  uint64_t *bars = (uint64_t *)(((uintptr_t)(xt + 3 * (size_t)(DP + 2) * FK) + 15) & ~(uintptr_t)15);   // xt: 3 train-slice buffers
  uint64_t *a_full = bars, *a_empty = bars + NSTA, *b_full = bars + 2 * NSTA, *b_empty = b_full + NSTB;
  uint64_t *t_full = b_empty + NSTB, *t_empty = t_full + 4;
  uint32_t *tmem_slot = (uint32_t *)(t_empty + 4);
  uint32_t *produced = tmem_slot + 1;     // K* blocks of this CTA written to the cache (running count)
  uint32_t *landed = tmem_slot + 2;       // operand-ring fills whose data has reached shared memory (running count)
  double *inv_ell = (double *)(bars + 24);  // 1 / lengthscale: the per-tile coordinate prep multiplies (FP64 division
                                            // is ~14 instructions on a pipe the B200 barely has)

  if (tid == 0) {
    for (int s = 0; s < NSTA; ++s) { mbar_init(smem_u32(&a_full[s]), 1); mbar_init(smem_u32(&a_empty[s]), 1); }
    for (int s = 0; s < NSTB; ++s) { mbar_init(smem_u32(&b_full[s]), 1); mbar_init(smem_u32(&b_empty[s]), 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(smem_u32(&t_full[s]), 1); mbar_init(smem_u32(&t_empty[s]), 8); }
    *produced = 0; *landed = 0;
    for (int j = 0; j < DP; ++j) inv_ell[j] = j < d ? 1.0 / prm.gp.ell[j] : 0.0;
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const bool mo = prm.mean_only != 0;
  const bool pon = prm.prof != nullptr;
  // fills per tile, and the offset of the last pass (the last read of every cache slot of a tile)
  int fpt = 0;
  for (int p = 0; p < n_pass; ++p) fpt += pass_kb_end(p);
  const int off_last = fpt - nkb;

  if (warp == 0) {
    // =============================== B producer ===========================================
    if (!mo && elect_one()) {
      uint32_t st = 0, ph = 0;
      long long w_bempty = 0; const long long t_start = clock64();
      for (long long it = 0; it < n_iter; ++it)
        for (int p = 0; p < n_pass; ++p) {
          const int c_first = NSLOT * p, c_last = min(c_first + NSLOT - 1, n_chunks - 1);
          const int kb_end = last_kb(c_last) + 1;
          for (int kb = 0; kb < kb_end; ++kb)
            for (int c = max(c_first, kb >> KSH); c <= c_last; ++c) {
              mbar_wait_prof(smem_u32(&b_empty[st]), ph ^ 1, 64, w_bempty, pon);
              const uint32_t dst = smem_u32(sB + st * STAGE_BYTES);
              const uint32_t full = mapa_rank(smem_u32(&b_full[st]), 0);
              if (leader) mbar_expect_tx(smem_u32(&b_full[st]), 2 * STAGE_BYTES);   // both halves
              tma_load_2d_2sm(dst, &map_hi, full, kb * FK, c * CW + (int)crank * 128);
              tma_load_2d_2sm(dst + PLANE_BYTES, &map_lo, full, kb * FK, c * CW + (int)crank * 128);
              if (++st == NSTB) { st = 0; ph ^= 1; }
            }
        }
      if (pon) { prm.prof[blockIdx.x * 16 + 0] = w_bempty; prm.prof[blockIdx.x * 16 + 1] = clock64() - t_start; }
    }
  } else if (warp == 3) {
    // =============================== A feeder =============================================
    if (!mo && elect_one()) {
      uint32_t sa = 0, pa = 0;
      long long w_prod = 0, w_aempty = 0;
      for (long long it = 0; it < n_iter; ++it)
        for (int p = 0; p < n_pass; ++p) {
          const int kb_end = pass_kb_end(p);
          const int kb_new = p == 0 ? 0 : pass_kb_end(p - 1);        // blocks first needed by this pass
          for (int kb = 0; kb < kb_end; ++kb) {
            if (kb >= kb_new) {
              const uint32_t need = (uint32_t)(it * nkb + kb + 1);
              const long long t0 = pon ? clock64() : 0;
              while (ld_acquire_smem(produced) < need) __nanosleep(64);
              if (pon) w_prod += clock64() - t0;
              asm volatile("fence.proxy.async.global;\n" ::: "memory");     // generic-proxy STGs -> async-proxy read
            }
            mbar_wait_prof(smem_u32(&a_empty[sa]), pa ^ 1, 32, w_aempty, pon);
            const uint32_t full = mapa_rank(smem_u32(&a_full[sa]), 0);
            if (leader) mbar_expect_tx(smem_u32(&a_full[sa]), 2 * STAGE_BYTES);
            tma_load_2d_2sm(smem_u32(sA + sa * STAGE_BYTES), &map_kc, full, 0, (int)((blockIdx.x * nkb + kb) * 256));
            if (++sa == NSTA) { sa = 0; pa ^= 1; }
          }
        }
      if (pon) { prm.prof[blockIdx.x * 16 + 8] = w_prod; prm.prof[blockIdx.x * 16 + 9] = w_aempty; }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader) ==================================
    if (!mo && leader && elect_one()) {
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, tph = 0, fills = 0;
      constexpr uint32_t IDESC_RT = IDESC | (F16 ? 0u : ((1u << 7) | (1u << 10)));   // bf16 / fp16 planes
      const uint32_t landed_peer = mapa_rank(smem_u32(landed), 1);
      long long w_afull = 0, w_bfull = 0, w_tempty = 0;
      for (long long it = 0; it < n_iter; ++it)
        for (int p = 0; p < n_pass; ++p) {
          const int c_first = NSLOT * p, c_last = min(c_first + NSLOT - 1, n_chunks - 1);
          const int kb_end = last_kb(c_last) + 1;
          for (int kb = 0; kb < kb_end; ++kb) {
            mbar_wait_prof(smem_u32(&a_full[sa]), pa, MMA_SLEEP_NS, w_afull, pon);
            tc_fence_after();
            ++fills;                                   // both CTAs' loads of this fill have landed:
            *(volatile uint32_t *)landed = fills;      // its cache slot may be overwritten (next tile)
            st_cluster_u32(landed_peer, fills);
            const uint32_t a_hi = smem_u32(sA + sa * STAGE_BYTES), a_lo = a_hi + PLANE_BYTES;
            for (int c = max(c_first, kb >> KSH); c <= c_last; ++c) {
              const int slot = c & (NSLOT - 1);
              if (kb == 0) {
                mbar_wait_prof(smem_u32(&t_empty[slot]), ((tph >> slot) & 1) ^ 1, MMA_SLEEP_NS, w_tempty, pon);
                tph ^= (1u << slot);
                tc_fence_after();
              }
              mbar_wait_prof(smem_u32(&b_full[sb]), pb, MMA_SLEEP_NS, w_bfull, pon);
              tc_fence_after();
              const uint32_t b_hi = smem_u32(sB + sb * STAGE_BYTES), b_lo = b_hi + PLANE_BYTES;
              const uint32_t dcol = tmem_base + (uint32_t)(slot * CW);
              if (!(prm.dbg & 1)) {
#pragma unroll
                for (int ks = 0; ks < FK / 16; ++ks) {
                  const uint64_t dah = make_sdesc(a_hi + ks * 32), dal = make_sdesc(a_lo + ks * 32);
                  const uint64_t dbh = make_sdesc(b_hi + ks * 32), dbl = make_sdesc(b_lo + ks * 32);
                  umma_bf16_2sm(dcol, dah, dbh, IDESC_RT, (kb > 0 || ks > 0) ? 1u : 0u);
                  umma_bf16_2sm(dcol, dal, dbh, IDESC_RT, 1u);
                  umma_bf16_2sm(dcol, dah, dbl, IDESC_RT, 1u);
                }
              }
              umma_commit_2sm(smem_u32(&b_empty[sb]));
              if (kb == last_kb(c)) umma_commit_2sm(smem_u32(&t_full[slot]));
              if (++sb == NSTB) { sb = 0; pb ^= 1; }
            }
            umma_commit_2sm(smem_u32(&a_empty[sa]));
            if (++sa == NSTA) { sa = 0; pa ^= 1; }
          }
        }
      if (pon) { prm.prof[blockIdx.x * 16 + 2] = w_afull; prm.prof[blockIdx.x * 16 + 3] = w_bfull; prm.prof[blockIdx.x * 16 + 4] = w_tempty; }
    }
  } else if (warp >= 4 && warp < 8) {
    // =============================== epilogue =============================================
    const int quad = warp - 4;
    const int row = quad * 32 + lane;
    uint32_t fph = 0;
    for (long long it = 0; it < (mo ? 0 : n_iter); ++it) {
      const long long tile = blockIdx.x + it * gridDim.x;
      double ss = 0.0;
      for (int c = 0; c < n_chunks; ++c) {
        const int slot = c & (NSLOT - 1);
        mbar_wait_sleep(smem_u32(&t_full[slot]), (fph >> slot) & 1, 200);
        fph ^= (1u << slot);
        tc_fence_after();
        float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < CW / 32; ++q) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(slot * CW + q * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) { float f = __uint_as_float(v[e]); part[e & 3] = fmaf(f, f, part[e & 3]); }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&t_empty[slot]), 0));
        ss += (double)((part[0] + part[1]) + (part[2] + part[3]));
      }
      const long long cg = tile * FM + row;
      if (cg < prm.m) {
        double v = prm.gp.sigma_f2 - ss * prm.gp.bscale[2];      // ss = s^2 ||L^-1 k*||^2
        prm.var_out[cg] = fmax(v, prm.gp.var_floor) + prm.gp.sigma_n2;
      }
    }
  } else if (warp >= 8) {
    // =============================== K1 generators ========================================
    // lane = (row_sub, chunk): the 8 lanes of a row group own the eight 16-byte operand chunks of one row,
    // so a warp-wide STG.128 writes 4 complete 128-byte rows (512 contiguous bytes) of the stage image.
    // Each block has two phases: distances (FFMA2 on the FMA pipe) and kernel function + bf16 split (MUFU
    // and ALU).  There is ONE barrier per block.  Warps 0-3 run their distance phase BEFORE it, warps 4-7
    // (the other warp of the same scheduler) AFTER it, so between two barriers one warp of every scheduler
    // is in its FMA phase while the other is in its MUFU phase.
    const int gt = tid - 8 * 32;
    const int q = lane & 7;
    const int row0 = 16 * (gt >> 5) + (lane >> 3);       // this thread's rows: row0 + 4 i, i = 0..3
    const bool matern = (prm.gp.kernel == OMBO_KERNEL_MATERN52);
    constexpr bool f16 = F16, direct = F16;         // fp16 planes (ill-conditioned GP): fp16 split + direct distances
    constexpr int RB = 128 / (32 * R);
    const bool ahead = (RB == 1) && ((gt >> 5) < 4);
    constexpr int XT_STRIDE = (DP + 2) * FK;
    constexpr int LD_ROWS = GEN_THREADS / 16;
    const int ld_j = gt >> 4, ld_o = (gt & 15) * 4;
    constexpr int LD_SWEEPS = (DP + 2 + LD_ROWS - 1) / LD_ROWS;
    auto prefetch_slice = [&](float *dst, int kb) {
#pragma unroll
      for (int sw = 0; sw < LD_SWEEPS; ++sw) {
        const int jj = ld_j + LD_ROWS * sw;
        if (jj <= DP + 1) {
          const float *src = (jj < DP) ? prm.gp.xs32 + (size_t)jj * np + kb * FK + ld_o
                                       : (jj == DP ? prm.gp.alpha32 : prm.gp.b2_32) + kb * FK + ld_o;
          const uint32_t sd = smem_u32(dst + jj * FK + ld_o);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(sd), "l"(src) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    float mu_acc[4];
    // distance phase: r^2 = |a|^2 + |b|^2 + sum_j (-2 a_j) b_j on centred scaled inputs
    auto dist = [&](const float *xs, const float (&x)[R][DP], const float (&a2)[R], float2 (&r2)[R][4]) __attribute__((always_inline)) {
      if (direct) {
        // direct differences: (b_j - a_j)^2 summed, no |a|^2 + |b|^2 - 2 a.b cancellation (ill-conditioned GPs)
#pragma unroll
        for (int rr = 0; rr < R; ++rr)
#pragma unroll
          for (int e = 0; e < 4; ++e) r2[rr][e] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < DP; ++j) {
          const float4 t0 = *(const float4 *)(xs + j * FK);
          const float4 t1 = *(const float4 *)(xs + j * FK + 4);
#pragma unroll
          for (int rr = 0; rr < R; ++rr) {
            const float2 xx = make_float2(x[rr][j], x[rr][j]), hh = make_float2(0.5f, 0.5f);   // x = -2 a_j
            const float2 d0 = __ffma2_rn(xx, hh, make_float2(t0.x, t0.y)), d1 = __ffma2_rn(xx, hh, make_float2(t0.z, t0.w));
            const float2 d2 = __ffma2_rn(xx, hh, make_float2(t1.x, t1.y)), d3 = __ffma2_rn(xx, hh, make_float2(t1.z, t1.w));
            r2[rr][0] = __ffma2_rn(d0, d0, r2[rr][0]);
            r2[rr][1] = __ffma2_rn(d1, d1, r2[rr][1]);
            r2[rr][2] = __ffma2_rn(d2, d2, r2[rr][2]);
            r2[rr][3] = __ffma2_rn(d3, d3, r2[rr][3]);
          }
        }
        return;
      }
      const float4 n0 = *(const float4 *)(xs + (DP + 1) * FK);
      const float4 n1 = *(const float4 *)(xs + (DP + 1) * FK + 4);
#pragma unroll
      for (int rr = 0; rr < R; ++rr) {
        const float2 aa = make_float2(a2[rr], a2[rr]);
        r2[rr][0] = __fadd2_rn(aa, make_float2(n0.x, n0.y));
        r2[rr][1] = __fadd2_rn(aa, make_float2(n0.z, n0.w));
        r2[rr][2] = __fadd2_rn(aa, make_float2(n1.x, n1.y));
        r2[rr][3] = __fadd2_rn(aa, make_float2(n1.z, n1.w));
      }
#pragma unroll
      for (int j = 0; j < DP; ++j) {
        const float4 t0 = *(const float4 *)(xs + j * FK);
        const float4 t1 = *(const float4 *)(xs + j * FK + 4);
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
          const float2 xx = make_float2(x[rr][j], x[rr][j]);
          r2[rr][0] = __ffma2_rn(xx, make_float2(t0.x, t0.y), r2[rr][0]);
          r2[rr][1] = __ffma2_rn(xx, make_float2(t0.z, t0.w), r2[rr][1]);
          r2[rr][2] = __ffma2_rn(xx, make_float2(t1.x, t1.y), r2[rr][2]);
          r2[rr][3] = __ffma2_rn(xx, make_float2(t1.z, t1.w), r2[rr][3]);
        }
      }
    };
    // kernel-function phase: k' = k(r)/sigma_f2, mean partial sums, bf16 hi/lo split, STG into the cache
    auto func = [&](const float *xs, int rb, unsigned char *st_hi, const float2 (&r2)[R][4]) __attribute__((always_inline)) {
      const float4 al0 = *(const float4 *)(xs + DP * FK);
      const float4 al1 = *(const float4 *)(xs + DP * FK + 4);
#pragma unroll
      for (int rr = 0; rr < R; ++rr) {
        float2 kv[4];
        if (matern) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 rad, ex;
            // |r^2|: rounding can leave -1e-7 where the distance is 0 (MUFU takes the modifier for free)
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad.x) : "f"(fabsf(r2[rr][e].x)));
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad.y) : "f"(fabsf(r2[rr][e].y)));
            const float2 arg = __fmul2_rn(rad, make_float2(-3.2259955597f, -3.2259955597f));   // -sqrt5 log2(e)
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.x) : "f"(arg.x));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.y) : "f"(arg.y));
            const float2 poly = __ffma2_rn(rad, make_float2(2.2360679775f, 2.2360679775f),
                                           __ffma2_rn(r2[rr][e], make_float2(1.6666666667f, 1.6666666667f),
                                                      make_float2(1.0f, 1.0f)));
            kv[e] = __fmul2_rn(poly, ex);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 arg = __fmul2_rn(r2[rr][e], make_float2(-0.7213475204f, -0.7213475204f));  // -0.5 log2(e)
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(kv[e].x) : "f"(arg.x));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(kv[e].y) : "f"(arg.y));
          }
        }
        {
          float2 m2 = __fmul2_rn(kv[0], make_float2(al0.x, al0.y));
          m2 = __ffma2_rn(kv[1], make_float2(al0.z, al0.w), m2);
          m2 = __ffma2_rn(kv[2], make_float2(al1.x, al1.y), m2);
          m2 = __ffma2_rn(kv[3], make_float2(al1.z, al1.w), m2);
          const float ms = m2.x + m2.y;          // static register indices (no local-memory array)
          if (RB == 1 || rb == 0) mu_acc[rr] += ms; else mu_acc[(R + rr) & 3] += ms;
        }
        if (!mo) {
          uint32_t hi[4], lo[4];
          if (f16) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __half2 h = __float22half2_rn(kv[e]);
              __half2 l = __float22half2_rn(__ffma2_rn(__half22float2(h), make_float2(-1.0f, -1.0f), kv[e]));   // k' - hi, exact
              hi[e] = *reinterpret_cast<uint32_t *>(&h);
              lo[e] = *reinterpret_cast<uint32_t *>(&l);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat162 h = __float22bfloat162_rn(kv[e]);
              const uint32_t hb = *reinterpret_cast<uint32_t *>(&h);
              const float2 hf = make_float2(__uint_as_float(hb << 16), __uint_as_float(hb & 0xffff0000u));
              __nv_bfloat162 l = __float22bfloat162_rn(__ffma2_rn(hf, make_float2(-1.0f, -1.0f), kv[e]));
              hi[e] = hb;
              lo[e] = *reinterpret_cast<uint32_t *>(&l);
            }
          }
          const int row = row0 + 4 * (R * rb + rr);
          const uint32_t off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + (((q ^ (row & 7)) & 7) << 4));
          *(uint4 *)(st_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);                  // STG: the cache IS the
          *(uint4 *)(st_hi + PLANE_BYTES + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);    // shared-memory image
        }
      }
    };
    auto load_rows = [&](int rb, float (&x)[R][DP], float (&a2)[R]) __attribute__((always_inline)) {
#pragma unroll
      for (int rr = 0; rr < R; ++rr) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < DP; ++j) { x[rr][j] = xc[j * FM + row0 + 4 * (R * rb + rr)]; acc = fmaf(x[rr][j], x[rr][j], acc); }
        a2[rr] = 0.25f * acc;
      }
    };
    // slices of blocks 0 and 1; from then on the slice of block g+2 is fetched right after barrier g
    prefetch_slice(xt, 0);
    prefetch_slice(xt + XT_STRIDE, 1 % nkb);
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    unsigned char *kc = prm.kcache + (size_t)blockIdx.x * nkb * STAGE_BYTES;
    int xb = 0;                                    // buffer (0..2) of the current block's slice
    for (long long it = 0; it < n_iter; ++it) {
      const long long tile = blockIdx.x + it * gridDim.x;
      asm volatile("bar.sync 1, %0;\n" ::"n"(GEN_THREADS));
      for (int e = gt; e < FM * DP; e += GEN_THREADS) {
        const int r_ = e & (FM - 1), j = e >> 7;
        const long long cg = tile * FM + r_;
        // -2 * centred scaled coordinate
        xc[j * FM + r_] = (cg < prm.m && j < d)
                              ? -2.0f * (float)((ombo_pool_coord(prm.pool, cg, j) - prm.gp.center[j]) * inv_ell[j])
                              : 0.f;
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(GEN_THREADS));
#pragma unroll
      for (int r4 = 0; r4 < 4; ++r4) mu_acc[r4] = 0.f;
      float x[RB == 1 ? R : 1][RB == 1 ? DP : 1], a2[RB == 1 ? R : 1];   // rows in registers for the whole tile
      float2 r2[RB == 1 ? R : 1][4];
      if constexpr (RB == 1) load_rows(0, x, a2);
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t g = (uint32_t)(it * nkb + kb);                  // running block index of this CTA
        const float *xs = xt + xb * XT_STRIDE + 8 * q;
        const bool work = !(prm.dbg & 2);
        if constexpr (RB == 1) { if (ahead && work) dist(xs, x, a2, r2); }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        // the STGs of block g-2 were issued two phases ago: the proxy fence finds them complete
        if (!mo) asm volatile("fence.proxy.async.global;\n" ::: "memory");
        asm volatile("bar.sync 1, %0;\n" ::"n"(GEN_THREADS));
        if (!mo && gt == 0 && g >= 2) st_release_smem(produced, g - 1);        // blocks 0 .. g-2
        {
          const int kn = kb + 2 < nkb ? kb + 2 : kb + 2 - nkb;
          prefetch_slice(xt + (xb == 0 ? 2 : xb - 1) * XT_STRIDE, kn);       // buffer of slice g-1 (dead)
        }
        if (!mo && it > 0) {
          // cache slot kb still holds the previous tile's block until its last read (last pass) has landed
          const uint32_t need = (uint32_t)((it - 1) * fpt + off_last + kb + 1);
          while (ld_volatile_smem(landed) < need) __nanosleep(64);
        }
        unsigned char *st_hi = kc + (size_t)kb * STAGE_BYTES;
        if (work) {
          if constexpr (RB == 1) {
            if (!ahead) dist(xs, x, a2, r2);
            func(xs, 0, st_hi, r2);
          } else {
#pragma unroll 1
            for (int rb = 0; rb < RB; ++rb) {
              float xr[R][DP], ar[R];
              float2 rr2[R][4];
              load_rows(rb, xr, ar);
              dist(xs, xr, ar, rr2);
              func(xs, rb, st_hi, rr2);
            }
          }
        }
        xb = xb == 2 ? 0 : xb + 1;
      }
      // mean: the 8 lanes of a row group hold the partial sums over the eight operand chunks
#pragma unroll
      for (int r4 = 0; r4 < 4; ++r4) {
        float v = mu_acc[r4];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        const long long cg = tile * FM + row0 + 4 * r4;
        if (q == 0 && cg < prm.m) {
          prm.mu_out[cg] = (double)v;
          if (mo) prm.var_out[cg] = nan("");
        }
      }
    }
    // the last two blocks are still unpublished
    if (!mo) {
      asm volatile("fence.proxy.async.global;\n" ::: "memory");
      asm volatile("bar.sync 1, %0;\n" ::"n"(GEN_THREADS));
      if (gt == 0) st_release_smem(produced, (uint32_t)(n_iter * nkb));
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u));
}

template <int DP, int R, bool F16>
static int launch_fast_dc(ombo_ctx *ctx, const CUtensorMap &map_hi, const CUtensorMap &map_lo, const CUtensorMap &map_kc,
                          const FastParams &prm, int grid, cudaStream_t s) {
  const size_t smem = (size_t)6 * STAGE_BYTES + (size_t)DP * FM * 4 + 3 * (size_t)(DP + 2) * FK * 4 +
                      16 + 24 * 8 + 24 * 8 + 1024;     // barriers + counters, 1/lengthscale, alignment slack
  OMBO_CUDA(cudaFuncSetAttribute(k_posterior_fast_dc<DP, R, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per (function, device)
  ProfScope prof(ctx, s);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(16 * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  OMBO_CUDA(cudaLaunchKernelEx(&cfg, k_posterior_fast_dc<DP, R, F16>, map_hi, map_lo, map_kc, prm));
  return OMBO_OK;
}

int ombo_launch_fast_dc(ombo_ctx *ctx, const CUtensorMap &map_hi, const CUtensorMap &map_lo, const CUtensorMap &map_kc,
                        const FastParams &prm, int grid, cudaStream_t s) {
  // 4 candidate rows x d coordinates live in registers: d <= 12 (larger d runs the single-CTA kernel)
  const int d = prm.gp.d;
  const bool f16 = (prm.gp.flags & OMBO_GP_FP16_PLANES) != 0;
#define DC_CASE(DPV)                                                                                     \
  if (d <= DPV) return f16 ? launch_fast_dc<DPV, 4, true>(ctx, map_hi, map_lo, map_kc, prm, grid, s)      \
                           : launch_fast_dc<DPV, 4, false>(ctx, map_hi, map_lo, map_kc, prm, grid, s)
  DC_CASE(2); DC_CASE(4); DC_CASE(6); DC_CASE(8); DC_CASE(10); DC_CASE(12);
#undef DC_CASE
  ombo_set_error("internal: decoupled fast kernel is instantiated for d <= 12 only");
  return OMBO_ERR_UNSUPPORTED;
}
