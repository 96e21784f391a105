// K1 + K2, fast mode, "f8c" operand format: one fp16 product plus two e4m3 correction products on a CTA pair.
// Same contract as k_posterior_fast (posterior_fast.cu): mu and var per candidate of one GP, tolerance rtol 1e-3.
//
//   V = K* . (s sigma_f2 L^-1)^T,  A = K* = k'(r) in (0,1] generated on chip,  B = s sigma_f2 L^-1 (planes from K3)
//   A = A_h + A_l   (A_h = fp16(A), exact residual A_l, |A_l| <= 2^-12)
//   B = B_h + B_l   (B_h = fp16(B), exact residual B_l, |B_l| <= 4 with max |B| scaled into [8192, 16384))
//   V ~= A_h B_h^T                              kind::f16      (K = 16 per instruction)         1   tensor unit
//      + e4m3(2^12 A_l) . e4m3(2^-12 B)^T       kind::f8f6f4   (K = 32: twice the MAC rate)     1/2
//      + e4m3(A) . e4m3(B_l)^T                  kind::f8f6f4                                    1/2
// All three accumulate into the same FP32 TMEM tile (every scale is a power of two, so the three products carry
// the same factor s).  The correction operands only need the ~2^-5 relative precision e4m3 has: they multiply
// residuals that are already 2^-12 of the operands.  2.0 tensor units per algorithmic MAC instead of the 3.0 of the
// 16-bit x3 split; measured sigma error against FP64 in DESIGN.md section 4 (the refresh selects this format only
// while the conditioning proxy of the factor stays below the measured limit).
//
// One CTA PAIR (cluster of 2, tcgen05 cta_group::2) owns 256 candidates: UMMA M = 256 (128 rows per CTA),
// N <= 256, each CTA TMA-loads only ITS half of every B tile, which halves the B traffic through shared memory
// and L2 per candidate -- the two limits the single-CTA kernel runs into.
//   warp 0        B producer (both CTAs): this CTA's rows of the unit's three planes, bytes counted on the leader
//   warp 1        MMA issuer (leader CTA): per (256-column chunk, 64-deep K-block) 4 x f16 + 2 x 2 x f8f6f4
//   warp 2        TMEM allocator (512 columns = two 256-column accumulator chunks per CTA)
//   warps 4-7     epilogue: tcgen05.ld, square + row sum, var = sigma_f2 - sum V^2 / s^2
//   warps 8..     K1 generators (8 or 16 warps): distances in FP32 (packed FFMA2), Matern-5/2 / RBF via MUFU,
//                 fp16 / e4m3 planes written straight into the UMMA operand layouts (SWIZZLE_128B / SWIZZLE_64B)
// L^-1 is lower triangular: chunk c needs K-blocks kb <= 4c+3 only, and on the four diagonal K-blocks of a chunk
// the rows above the diagonal band are zero, so N shrinks to 192 / 128 / 64 -- in pair mode by letting each CTA
// fetch a different row range of the band (CTA r loads rows r0 + r N/2 .. of the chunk) and offsetting the
// accumulator columns by r0.  TMEM holds two chunks; n_pad > 512 takes several passes with the K* blocks of
// earlier passes streamed back from a per-CTA cache in L2 (as in posterior_fast.cu).
#include <cuda_fp8.h>

#include "umma.cuh"

#define F8_HI_BYTES 16384   // 128 rows x 64 fp16 (SWIZZLE_128B)
#define F8_P8_BYTES 8192    // 128 rows x 64 e4m3 (SWIZZLE_64B)
#define F8_OFF_C1 F8_HI_BYTES                  // A: e4m3(2^12 A_l)   B: e4m3(2^-12 B)
#define F8_OFF_C2 (F8_HI_BYTES + F8_P8_BYTES)  // A: e4m3(A)          B: e4m3(B_l)

// UMMA shared-memory descriptor, K-major, SWIZZLE_64B: rows of 64 B, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t make_sdesc64(uint32_t saddr) {
  const uint32_t lo = ((saddr & 0x3FFFF) >> 4) | (1u << 16);
  const uint32_t hi = 32u | (1u << 14) | (4u << 29);                   // SBO = 512 B, version 1, SWIZZLE_64B
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void umma_f8_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

struct F8Maps {            // tensor maps of the three B planes (128-row and 32-row boxes) and of the K* cache
  CUtensorMap hi128, hi32, c1_128, c1_32, c2_128, c2_32, kc;
};

template <int DP, int R, int GW>
__global__ void __launch_bounds__((8 + GW) * 32, 1)
k_posterior_fast8(const __grid_constant__ F8Maps maps, const FastParams prm) {
  constexpr int NSTA = 3, NSTB = 3;
  constexpr int GEN_WARPS = GW, GEN_THREADS = GW * 32;
  constexpr int CW = 256, NSLOT = 2, KSH = 2;
  constexpr int PREP_THREADS = 64;             // warps 2 and 3: candidate coordinates of the NEXT tile
  // D = f32, A = B = fp16 / e4m3 (format code 0 in both kinds), K-major, M = 256 (pair); N is patched per unit
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(256 >> 4) << 24);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d = prm.gp.d, np = prm.gp.n_pad;
  const int nkb = np / FK;
  const int n_chunks = (np + CW - 1) / CW;
  const int n_pass = (n_chunks + NSLOT - 1) / NSLOT;
  const long long n_tiles = (prm.m + FM - 1) / FM;
  const long long n_iter = (n_tiles + gridDim.x - 1) / gridDim.x;   // both CTAs of a pair walk the same sequence
  const uint32_t crank = cluster_rank();
  const bool leader = crank == 0;
  const bool use_cache = prm.kcache != nullptr;
  auto last_kb = [&](int c) { return min(((c + 1) << KSH) - 1, nkb - 1); };
  // K-blocks of pass p: [0, kb_cached) were generated by earlier passes and come back from the L2 cache,
  // [kb_cached, kb_end) are generated now
  auto pass_kb_end = [&](int p) { return last_kb(min(NSLOT * p + NSLOT - 1, n_chunks - 1)) + 1; };
  auto pass_kb_cached = [&](int p) { return (use_cache && p > 0) ? pass_kb_end(p - 1) : 0; };
  // Order in which a pass walks its K-blocks: fresh and cached blocks ALTERNATE (F0 C0 F1 C1 ...), so that the
  // generators work on the next fresh block while the tensor core multiplies a reloaded one.  In block order
  // (all cached blocks first) the generators idle through the reloads and the MMA then starves on the fresh run.
  auto seq_kb = [&](int kb_cached, int kb_end, int i) {
    const int nf = kb_end - kb_cached, nc = kb_cached, nmin = min(nf, nc);
    if (i < 2 * nmin) return (i & 1) ? (i >> 1) : kb_cached + (i >> 1);
    return nf > nc ? kb_cached + (i - nc) : (i - nf);
  };
  // position of the last K-block of chunk c in that order (the chunk's accumulator is complete after it)
  auto last_pos = [&](int kb_cached, int kb_end, int c) {
    const int lk = last_kb(c);
    int i = kb_end - 1;
    while (i > 0 && seq_kb(kb_cached, kb_end, i) > lk) --i;
    return i;
  };
  // columns of chunk c that K-block kb touches: [r0, r0 + ncols) of the chunk (rows of L^-1 above the diagonal band
  // and beyond n_pad are zero)
  auto unit_r0 = [&](int c, int kb) { return max(0, kb - 4 * c) * 64; };
  auto unit_ncols = [&](int c, int kb) { return min(CW, np - c * CW) - unit_r0(c, kb); };

  unsigned char *sA = fast_smem + ((1024u - (smem_u32(fast_smem) & 1023u)) & 1023u);
  unsigned char *sB = sA + NSTA * STAGE_BYTES;
  float *xc = (float *)(sB + NSTB * STAGE_BYTES);              // [2][DP][128] scaled candidate coords (double buffered)
  float *xt = xc + 2 * (size_t)DP * FM;                        // [2][(DP+2)][64] train slice (+ alpha, |b|^2 rows)
  float *mu_sm = xt + 2 * (size_t)(DP + 2) * FK;               // [8][128]
  uint64_t *bars = (uint64_t *)(((uintptr_t)(mu_sm + 8 * FM) + 15) & ~(uintptr_t)15);
  uint64_t *a_full = bars, *a_empty = bars + NSTA, *b_full = bars + 2 * NSTA, *b_empty = b_full + NSTB;
  uint64_t *t_full = b_empty + NSTB, *t_empty = t_full + 4;
  uint32_t *tmem_slot = (uint32_t *)(t_empty + 4);
  double *inv_ell = (double *)(bars + 24);

  if (tid == 0) {
    for (int j = 0; j < DP; ++j) inv_ell[j] = j < d ? 1.0 / prm.gp.ell[j] : 0.0;
    for (int s = 0; s < NSTA; ++s) { mbar_init(smem_u32(&a_full[s]), 2 * GEN_WARPS); mbar_init(smem_u32(&a_empty[s]), 1); }
    for (int s = 0; s < NSTB; ++s) { mbar_init(smem_u32(&b_full[s]), 1); mbar_init(smem_u32(&b_empty[s]), 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(smem_u32(&t_full[s]), 1); mbar_init(smem_u32(&t_empty[s]), 8); }
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer arrives on / loads into this CTA's barriers and smem
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool pon = prm.prof != nullptr;

  if (warp == 0) {
    // =============================== B producer (both CTAs) ===============================
    if (elect_one()) {
      uint32_t st = 0, ph = 0;
      long long w_bempty = 0; const long long t_start = clock64();
      for (long long it = 0; it < n_iter; ++it) {
        for (int p = 0; p < n_pass; ++p) {
          const int c_first = NSLOT * p, c_last = min(c_first + NSLOT - 1, n_chunks - 1);
          const int kb_end = pass_kb_end(p), kb_cached = pass_kb_cached(p);
          for (int i = 0; i < kb_end; ++i) {
            const int kb = seq_kb(kb_cached, kb_end, i);
            for (int c = max(c_first, kb >> KSH); c <= c_last; ++c) {
              const int ncols = unit_ncols(c, kb), rows = ncols >> 1;          // rows of the band this CTA holds
              const int row0 = c * CW + unit_r0(c, kb) + (int)crank * rows;
              mbar_wait_prof(smem_u32(&b_empty[st]), ph ^ 1, 64, w_bempty, pon);
              const uint32_t dst = smem_u32(sB + st * STAGE_BYTES);
              // the leader alone arms its barrier, with the bytes of BOTH halves (256 B per band row over the planes)
              const uint32_t full = mapa_rank(smem_u32(&b_full[st]), 0);
              if (leader) mbar_expect_tx(smem_u32(&b_full[st]), (uint32_t)(ncols * 256));
              if (rows == 128) {
                tma_load_2d_2sm(dst, &maps.hi128, full, kb * FK, row0);
                tma_load_2d_2sm(dst + F8_OFF_C1, &maps.c1_128, full, kb * FK, row0);
                tma_load_2d_2sm(dst + F8_OFF_C2, &maps.c2_128, full, kb * FK, row0);
              } else {
                for (int rr = 0; rr < rows; rr += 32) {
                  tma_load_2d_2sm(dst + (uint32_t)(rr * 128), &maps.hi32, full, kb * FK, row0 + rr);
                  tma_load_2d_2sm(dst + F8_OFF_C1 + (uint32_t)(rr * 64), &maps.c1_32, full, kb * FK, row0 + rr);
                  tma_load_2d_2sm(dst + F8_OFF_C2 + (uint32_t)(rr * 64), &maps.c2_32, full, kb * FK, row0 + rr);
                }
              }
              if (++st == NSTB) { st = 0; ph ^= 1; }
            }
          }
        }
      }
      if (pon) { prm.prof[blockIdx.x * 16 + 0] = w_bempty; prm.prof[blockIdx.x * 16 + 1] = clock64() - t_start; }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA) ==============================
    if (leader && elect_one()) {
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, tph = 0;
      long long w_afull = 0, w_bfull = 0, w_tempty = 0;
      for (long long it = 0; it < n_iter; ++it) {
        for (int p = 0; p < n_pass; ++p) {
          const int c_first = NSLOT * p, c_last = min(c_first + NSLOT - 1, n_chunks - 1);
          const int kb_end = pass_kb_end(p), kb_cached = pass_kb_cached(p);
          int lastp[NSLOT];
          for (int c = c_first; c <= c_last; ++c) lastp[c - c_first] = last_pos(kb_cached, kb_end, c);
          for (int i = 0; i < kb_end; ++i) {
            const int kb = seq_kb(kb_cached, kb_end, i);
            mbar_wait_prof(smem_u32(&a_full[sa]), pa, MMA_SLEEP_NS, w_afull, pon);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(sA + sa * STAGE_BYTES);
            for (int c = max(c_first, kb >> KSH); c <= c_last; ++c) {
              const int slot = c & (NSLOT - 1);
              if (i == 0) {                                  // first touch of this accumulator slot (every chunk of the
                                                             // pass takes part in the pass's first block)
                mbar_wait_prof(smem_u32(&t_empty[slot]), ((tph >> slot) & 1) ^ 1, MMA_SLEEP_NS, w_tempty, pon);
                tph ^= (1u << slot);
                tc_fence_after();
              }
              const int r0 = unit_r0(c, kb), ncols = unit_ncols(c, kb);
              const uint32_t dcol = tmem_base + (uint32_t)(slot * CW + r0);
              const uint32_t idesc = IDESC | ((uint32_t)(ncols >> 3) << 17);
              mbar_wait_prof(smem_u32(&b_full[sb]), pb, MMA_SLEEP_NS, w_bfull, pon);
              tc_fence_after();
              const uint32_t b_hi = smem_u32(sB + sb * STAGE_BYTES);
              if (!(prm.dbg & 1)) {
#pragma unroll
                for (int ks = 0; ks < FK / 16; ++ks)
                  umma_bf16_2sm(dcol, make_sdesc(a_hi + ks * 32), make_sdesc(b_hi + ks * 32), idesc, (i > 0 || ks > 0) ? 1u : 0u);
                if (!(prm.dbg & 4)) {
#pragma unroll
                  for (int ks = 0; ks < FK / 32; ++ks)
                    umma_f8_2sm(dcol, make_sdesc64(a_hi + F8_OFF_C1 + ks * 32), make_sdesc64(b_hi + F8_OFF_C1 + ks * 32), idesc, 1u);
#pragma unroll
                  for (int ks = 0; ks < FK / 32; ++ks)
                    umma_f8_2sm(dcol, make_sdesc64(a_hi + F8_OFF_C2 + ks * 32), make_sdesc64(b_hi + F8_OFF_C2 + ks * 32), idesc, 1u);
                }
              }
              umma_commit_2sm(smem_u32(&b_empty[sb]));
              if (i == lastp[c - c_first]) umma_commit_2sm(smem_u32(&t_full[slot]));      // chunk complete
              if (++sb == NSTB) { sb = 0; pb ^= 1; }
            }
            umma_commit_2sm(smem_u32(&a_empty[sa]));
            if (++sa == NSTA) { sa = 0; pa ^= 1; }
          }
        }
      }
      if (pon) { prm.prof[blockIdx.x * 16 + 2] = w_afull; prm.prof[blockIdx.x * 16 + 3] = w_bfull; prm.prof[blockIdx.x * 16 + 4] = w_tempty; }
    }
  } else if (warp == 2 || warp == 3) {
    // =============================== candidate coordinates of the next tile ===============
    // -2 x the centred, length-scaled coordinates of tile `it` into xc[it & 1] while the generators work on tile
    // it - 1 (FP64 counter generator / conversion: ~10 k cycles per tile that used to sit on the generators'
    // critical path).  Named barriers 2 + b ("xc[b] filled") and 4 + b ("xc[b] read").
    const int pt = tid - 64;
    for (long long it = 0; it < n_iter; ++it) {
      const int b = (int)(it & 1);
      const long long tile = blockIdx.x + it * gridDim.x;
      if (it >= 2) named_bar_sync(4 + b, PREP_THREADS + GEN_THREADS);
      float *xcb = xc + (size_t)b * DP * FM;
      for (int e = pt; e < FM * DP; e += PREP_THREADS) {
        const int r_ = e & (FM - 1), j = e >> 7;
        const long long cg = tile * FM + r_;
        xcb[j * FM + r_] = (cg < prm.m && j < d)
                               ? -2.0f * (float)((ombo_pool_coord(prm.pool, cg, j) - prm.gp.center[j]) * inv_ell[j])
                               : 0.f;
      }
      named_bar_arrive(2 + b, PREP_THREADS + GEN_THREADS);
    }
  } else if (warp >= 4 && warp < 8) {
    // =============================== epilogue (both CTAs: own 128 candidates) =============
    const int quad = warp - 4;
    const int row = quad * 32 + lane;
    uint32_t fph = 0;
    const uint32_t t_empty_leader = mapa_rank(smem_u32(&t_empty[0]), 0);
    for (long long it = 0; it < n_iter; ++it) {
      const long long tile = blockIdx.x + it * gridDim.x;
      double ss = 0.0;
      for (int c = 0; c < n_chunks; ++c) {
        const int slot = c & (NSLOT - 1);
        mbar_wait_sleep(smem_u32(&t_full[slot]), (fph >> slot) & 1, 200);
        fph ^= (1u << slot);
        tc_fence_after();
        float part[4] = {0.f, 0.f, 0.f, 0.f};
        const int ncw = min(CW, np - c * CW);
        for (int q = 0; q < ncw / 32; ++q) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(slot * CW + q * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) { float f = __uint_as_float(v[e]); part[e & 3] = fmaf(f, f, part[e & 3]); }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(t_empty_leader + (uint32_t)(slot * 8));
        ss += (double)((part[0] + part[1]) + (part[2] + part[3]));
      }
      const long long cg = tile * FM + row;
      if (cg < prm.m) {
        double v = prm.gp.sigma_f2 - ss * prm.gp.bscale[2];
        prm.var_out[cg] = fmax(v, prm.gp.var_floor) + prm.gp.sigma_n2;
      }
    }
  } else if (warp >= 8) {
    // =============================== K1 generators ========================================
    // Warp q (of 8 per row half) owns K columns 8q..8q+7 of the block; lane l owns candidate rows l + 32 rr of
    // its row half; the candidate coordinates stay in registers for the whole tile.
    const int gt = tid - 8 * 32;
    const int q = (gt >> 5) & 7;
    const int rh = gt >> 8;                      // row half (GW = 16 only)
    const bool matern = (prm.gp.kernel == OMBO_KERNEL_MATERN52);
    constexpr int ROWS = R;                       // rows per lane; R * 32 * (GW / 8) == 128
    static_assert(R * 32 * (GW / 8) == 128, "generator rows must cover the tile");
    constexpr int XT_STRIDE = (DP + 2) * FK;
    uint32_t sa = 0, pa = 0;
    int xbuf = 0;
    long long w_aempty = 0;
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
    prefetch_slice(xt, 0);
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    asm volatile("bar.sync 1, %0;\n" ::"n"(GEN_THREADS));
    const uint32_t a_full_leader0 = mapa_rank(smem_u32(&a_full[0]), 0);
    for (long long it = 0; it < n_iter; ++it) {
      const long long tile = blockIdx.x + it * gridDim.x;
      const int b = (int)(it & 1);
      named_bar_sync(2 + b, PREP_THREADS + GEN_THREADS);       // xc[b] holds this tile
      const float *xcb = xc + (size_t)b * DP * FM;
      float mu_acc[ROWS];
      float x[ROWS][DP], a2[ROWS];
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) {
        float acc = 0.f;
        mu_acc[rr] = 0.f;
#pragma unroll
        for (int j = 0; j < DP; ++j) { x[rr][j] = xcb[j * FM + lane + 32 * (ROWS * rh + rr)]; acc = fmaf(x[rr][j], x[rr][j], acc); }
        a2[rr] = 0.25f * acc;
      }
      if (it + 2 < n_iter) named_bar_arrive(4 + b, PREP_THREADS + GEN_THREADS);   // xc[b] may be refilled
      for (int p = 0; p < n_pass; ++p) {
        const int kb_end = pass_kb_end(p), kb_cached = pass_kb_cached(p);
        const bool store_cache = use_cache && (p + 1 < n_pass);
        unsigned char *kc = prm.kcache + (size_t)blockIdx.x * nkb * STAGE_BYTES;
        if (kb_cached > 0 && gt == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");   // stores landed
        for (int i = 0; i < kb_end; ++i) {
          const int kb = seq_kb(kb_cached, kb_end, i);
          const uint32_t full_leader = a_full_leader0 + (uint32_t)(sa * 8);
          if (kb < kb_cached) {
            // a K* block of an earlier pass: one thread streams the cached 32 KB stage back (2-SM load: the bytes are
            // counted on the leader's barrier) and performs all of this CTA's arrivals.  EVERY generator thread waits
            // for the stage's release first: a warp that skipped ahead through a run of cached blocks would get more
            // than one barrier phase ahead of the MMA, and a parity wait cannot tell phases two apart.
            if (gt == 0) {
              // a cache store issued from this stage three blocks ago must have finished reading it
              if (store_cache) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
              mbar_wait_prof(smem_u32(&a_empty[sa]), pa ^ 1, 32, w_aempty, pon);
              mbar_expect_tx_remote(full_leader, STAGE_BYTES);
              mbar_arrive_n_remote(full_leader, GEN_WARPS - 1);
              tma_load_2d_2sm(smem_u32(sA + sa * STAGE_BYTES), &maps.kc, full_leader, 0, (int)((blockIdx.x * nkb + kb) * 256));
            } else {
              mbar_wait_sleep(smem_u32(&a_empty[sa]), pa ^ 1, 128);
            }
            __syncwarp();
            if (++sa == NSTA) { sa = 0; pa ^= 1; }
            continue;
          }
          // with the cache every block is generated once per tile, otherwise the mean is taken in the last pass
          const bool do_mu = use_cache || (p == n_pass - 1);
          // the next block this CTA generates (fresh blocks come in increasing order; across passes / tiles it wraps)
          prefetch_slice(xt + (xbuf ^ 1) * XT_STRIDE, (kb + 1 < kb_end) ? kb + 1 : (p + 1 < n_pass ? pass_kb_cached(p + 1) : 0));
          const float *xs = xt + xbuf * XT_STRIDE + 8 * q;
          unsigned char *st_hi = sA + sa * STAGE_BYTES;
          float2 r2[ROWS][4];
          if (!(prm.dbg & 2)) {
            const float4 n0 = *(const float4 *)(xs + (DP + 1) * FK);      // |b_i|^2
            const float4 n1 = *(const float4 *)(xs + (DP + 1) * FK + 4);
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr) {
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
              for (int rr = 0; rr < ROWS; ++rr) {
                const float2 xx = make_float2(x[rr][j], x[rr][j]);            // -2 a_j
                r2[rr][0] = __ffma2_rn(xx, make_float2(t0.x, t0.y), r2[rr][0]);
                r2[rr][1] = __ffma2_rn(xx, make_float2(t0.z, t0.w), r2[rr][1]);
                r2[rr][2] = __ffma2_rn(xx, make_float2(t1.x, t1.y), r2[rr][2]);
                r2[rr][3] = __ffma2_rn(xx, make_float2(t1.z, t1.w), r2[rr][3]);
              }
            }
          } else {
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
              for (int e = 0; e < 4; ++e) r2[rr][e] = make_float2(1.f, 1.f);
          }
          const float4 al0 = *(const float4 *)(xs + DP * FK);
          const float4 al1 = *(const float4 *)(xs + DP * FK + 4);
          mbar_wait_prof(smem_u32(&a_empty[sa]), pa ^ 1, 32, w_aempty, pon);   // stage released by the MMA
#pragma unroll
          for (int rr = 0; rr < ROWS; ++rr) {
            float2 kv[4];
            if (prm.dbg & 2) {
#pragma unroll
              for (int e = 0; e < 4; ++e) kv[e] = r2[rr][e];
            } else if (matern) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 rad, ex;
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
            if (do_mu) {
              float2 m2 = __fmul2_rn(kv[0], make_float2(al0.x, al0.y));
              m2 = __ffma2_rn(kv[1], make_float2(al0.z, al0.w), m2);
              m2 = __ffma2_rn(kv[2], make_float2(al1.x, al1.y), m2);
              m2 = __ffma2_rn(kv[3], make_float2(al1.z, al1.w), m2);
              mu_acc[rr] += m2.x + m2.y;
            }
            uint32_t hi[4];
            uint32_t c1[2], c2[2];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __half2 h = __float22half2_rn(kv[e]);
              const float2 hf = __half22float2(h);
              // 2^12 (k' - hi): the residual is exact in FP32, the scaling too
              const float2 res = __ffma2_rn(hf, make_float2(-4096.0f, -4096.0f), __fmul2_rn(kv[e], make_float2(4096.0f, 4096.0f)));
              hi[e] = *reinterpret_cast<const uint32_t *>(&h);
              const uint32_t l8 = (uint32_t)__nv_cvt_float2_to_fp8x2(res, __NV_SATFINITE, __NV_E4M3);
              const uint32_t a8 = (uint32_t)__nv_cvt_float2_to_fp8x2(hf, __NV_SATFINITE, __NV_E4M3);
              if (e & 1) { c1[e >> 1] |= l8 << 16; c2[e >> 1] |= a8 << 16; }
              else { c1[e >> 1] = l8; c2[e >> 1] = a8; }
            }
            const int row = lane + 32 * (ROWS * rh + rr);
            const uint32_t off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + (((q ^ (row & 7)) & 7) << 4));
            *(uint4 *)(st_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            const uint32_t off8 = (uint32_t)((row >> 3) * 512 + (row & 7) * 64 + ((((q >> 1) ^ (row >> 1)) & 3) << 4) + (q & 1) * 8);
            *(uint2 *)(st_hi + F8_OFF_C1 + off8) = make_uint2(c1[0], c1[1]);
            *(uint2 *)(st_hi + F8_OFF_C2 + off8) = make_uint2(c2[0], c2[1]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(full_leader);            // the leader's MMA consumes both halves
          const uint32_t stage_just_written = smem_u32(sA + sa * STAGE_BYTES);
          if (++sa == NSTA) { sa = 0; pa ^= 1; }
          asm volatile("cp.async.wait_group 0;\n" ::: "memory");          // next slice has landed
          if (store_cache && gt == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
          asm volatile("bar.sync 1, %0;\n" ::"n"(GEN_THREADS));
          if (store_cache && gt == 0) bulk_store(kc + (size_t)kb * STAGE_BYTES, stage_just_written, STAGE_BYTES);
          xbuf ^= 1;
        }
      }
      if (pon && gt == 0) prm.prof[blockIdx.x * 16 + 5] = w_aempty;
      // mean: one partial sum per (chunk warp, row)
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) mu_sm[q * FM + lane + 32 * (ROWS * rh + rr)] = mu_acc[rr];
      asm volatile("bar.sync 1, %0;\n" ::"n"(GEN_THREADS));
      if (gt < FM) {
        const long long cg = tile * FM + gt;
        if (cg < prm.m) {
          float acc = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) acc += mu_sm[w * FM + gt];
          prm.mu_out[cg] = (double)acc;
        }
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(GEN_THREADS));       // mu_sm is rewritten by the next tile
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // nobody exits while the peer may still signal it
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u));
}

// ---- host side ------------------------------------------------------------------------------
static int make_plane_map(CUtensorMap *map, const void *base, int n_pad, int esz, int box_rows) {
  PFN_encodeTiled_t enc = ombo_get_encode_tiled();
  if (!enc) { ombo_set_error("cuTensorMapEncodeTiled is not available from the driver"); return OMBO_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)n_pad, (cuuint64_t)n_pad};
  cuuint64_t strides[1] = {(cuuint64_t)n_pad * esz};
  cuuint32_t box[2] = {FK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   esz == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { ombo_set_error("cuTensorMapEncodeTiled (f8c plane) failed (%d)", (int)r); return OMBO_ERR_CUDA; }
  return OMBO_OK;
}

template <int DP, int R, int GW>
static int launch_fast8(ombo_ctx *ctx, const F8Maps &maps, const FastParams &prm, int grid, cudaStream_t s) {
  const size_t smem = (size_t)6 * STAGE_BYTES + 2 * (size_t)DP * FM * 4 + 2 * (size_t)(DP + 2) * FK * 4 +
                      8 * FM * 4 + 16 + 24 * 8 + 32 * 8 + 1024;
  // per device, not per process: the attribute belongs to the (function, device) pair
  OMBO_CUDA(cudaFuncSetAttribute(k_posterior_fast8<DP, R, GW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(ctx, s);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3((8 + GW) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  OMBO_CUDA(cudaLaunchKernelEx(&cfg, k_posterior_fast8<DP, R, GW>, maps, prm));
  return OMBO_OK;
}

// mu / var of one GP whose planes are in the f8c format (gp.flags & OMBO_GP_F8C_PLANES), d <= 12
int ombo_posterior_fast8(ombo_ctx *ctx, const GpDev &gp, const PoolDev &pool, long long m, double *mu, double *var,
                         cudaStream_t s) {
  const long long tiles = (m + FM - 1) / FM;
  F8Maps maps;
  const unsigned char *c1 = (const unsigned char *)gp.blo, *c2 = c1 + (size_t)gp.n_pad * gp.n_pad;
  int rc;
  if ((rc = make_plane_map(&maps.hi128, gp.bhi, gp.n_pad, 2, 128))) return rc;
  if ((rc = make_plane_map(&maps.hi32, gp.bhi, gp.n_pad, 2, 32))) return rc;
  if ((rc = make_plane_map(&maps.c1_128, c1, gp.n_pad, 1, 128))) return rc;
  if ((rc = make_plane_map(&maps.c1_32, c1, gp.n_pad, 1, 32))) return rc;
  if ((rc = make_plane_map(&maps.c2_128, c2, gp.n_pad, 1, 128))) return rc;
  if ((rc = make_plane_map(&maps.c2_32, c2, gp.n_pad, 1, 32))) return rc;
  maps.kc = maps.hi128;
  FastParams prm;
  prm.gp = gp; prm.pool = pool; prm.m = m; prm.mu_out = mu; prm.var_out = var; prm.mean_only = 0;
  prm.dbg = ctx->knobs.fast_dbg; prm.trim_b = 1; prm.kcache = nullptr;
  const bool want_prof = ctx->knobs.fast_profile != 0;
  if (want_prof && !ctx->prof_dev) OMBO_CUDA(cudaMalloc(&ctx->prof_dev, 16 * 256 * sizeof(long long)));
  prm.prof = want_prof ? ctx->prof_dev : nullptr;
  if (want_prof) OMBO_CUDA(cudaMemsetAsync(ctx->prof_dev, 0, 16 * 256 * sizeof(long long), s));
  int grid = (int)(tiles < ctx->num_sms ? tiles : ctx->num_sms);
  grid = (grid + 1) / 2 * 2;                    // whole pairs; a surplus CTA runs a dummy tile
  if (grid > ctx->num_sms) grid = ctx->num_sms / 2 * 2;
  if (gp.n_pad > 512) {                          // more than one TMEM pass: K* blocks are cached in L2 between passes
    rc = ombo_ws_reserve(&ctx->ws_scratch, &ctx->ws_scratch_bytes, (size_t)grid * (gp.n_pad / FK) * STAGE_BYTES);
    if (rc) return rc;
    prm.kcache = (unsigned char *)ctx->ws_scratch;
    rc = ombo_make_linear_map(&maps.kc, prm.kcache, (size_t)grid * (gp.n_pad / FK) * 256);
    if (rc) return rc;
  }
  const int d = gp.d;
  const bool gw16 = ctx->knobs.fast_gen_warps == 16;
#define F8_DISPATCH(DPV)                                                                     \
  rc = gw16 ? launch_fast8<DPV, 2, 16>(ctx, maps, prm, grid, s) : launch_fast8<DPV, 4, 8>(ctx, maps, prm, grid, s)
  if (d <= 2) { F8_DISPATCH(2); }
  else if (d <= 4) { F8_DISPATCH(4); }
  else if (d <= 6) { F8_DISPATCH(6); }
  else if (d <= 8) { F8_DISPATCH(8); }
  else if (d <= 10) { F8_DISPATCH(10); }
  else { F8_DISPATCH(12); }
#undef F8_DISPATCH
  if (rc) return rc;
  ctx->launches += 1;
  OMBO_CUDA(cudaGetLastError());
  if (want_prof) {
    static long long h[16 * 256];
    OMBO_CUDA(cudaStreamSynchronize(s));
    OMBO_CUDA(cudaMemcpy(h, ctx->prof_dev, sizeof(h), cudaMemcpyDeviceToHost));
    double a[8] = {0};
    for (int b = 0; b < grid; ++b) for (int k = 0; k < 8; ++k) a[k] += (double)h[b * 16 + k] / grid;
    fprintf(stderr, "[fast8 prof] per-CTA cycles (leader + peer averaged): tma.wait_b_empty %.0f / tma.total %.0f | mma.wait_a_full %.0f "
            "wait_b_full %.0f wait_t_empty %.0f | gen.wait_a_empty %.0f  (tiles/CTA %.1f)\n",
            a[0], a[1], 2 * a[2], 2 * a[3], 2 * a[4], a[5], (double)tiles / grid);
  }
  return OMBO_OK;
}
