// block_staged.cu -- weight-id-major block-diagonal aggregation with TMA-STAGED gathers (sm_100a).
//
// Same arithmetic and the same work items as k_block_rel (rgcn_kernels.cu): a warp owns one (work item,
// column slab) = up to item_max messages of ONE weight id, keeps that weight block slab in REGISTERS,
// pre-sums runs of messages with the same accumulation row and adds each run's transformed row into
// out[row] with 128-bit vector reductions that resolve in L2.  What changes is the data path of the gathered
// rows.  ncu on the HBM-bound shape (d = 512, s = 8, profiles/r2_ncu_synthetic.md) showed k_block_rel stalled
// on its own gathers (long-scoreboard 3.2 of 3.5 resident warps per scheduler, 16 warps/SM at 128 registers,
// DRAM 38 %, L2 45 %, issue 50 %): the rows in flight lived in registers, so occupancy bounded the bytes in
// flight.  Here every warp drives a private ring of DEPTH row slots in shared memory: one elected lane issues
// cp.async.bulk (TMA bulk copy, global -> shared, mbarrier complete_tx) for the row of message m + DEPTH the
// moment message m has been read out of its slot, so DEPTH-1 rows per warp are always in flight without
// holding a single register, the consumer reads its quads with conflict-free LDS.128, and the former
// register -> shared -> register bounce (STS.128 + 8 scalar LDS per run) disappears: the lanes that share a
// block exchange their quads with shuffles.
//
// Reference arithmetic: encoders/message_gcns/gcn_basis_concat.py:35-52 (messages) and :69-75 (the two
// normalised SpMMs); FUSE_DW also produces the block weight gradient in the same walk (what tf.gradients
// derives for W_forward / W_backward, optimization/abstract.py:117-118).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"

#define FULL 0xffffffffu

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_a(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// TMA bulk copy global -> shared (1-D, 16 B granular); completion is signalled on `bar` as transaction bytes
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {  // gathered rows stream through L2 once: do not let them
  uint64_t p;                                               // evict the accumulation window the reductions live in
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// Ampere-style asynchronous copy, 16 B per lane (LDGSTS): one instruction moves a 512 B row slab; src_bytes = 0
// zero-fills the destination (lanes past the row end)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes, uint64_t policy) {
  // No L2::cache_hint operand here.  ptxas 12.9 either drops it for this zero-fill form (identical LDGSTS encoding with
  // and without) or, when the shared address is partly uniform, emits `LDGSTS [R+UR0], desc[UR1]` with UR0/UR1 never
  // written -- an illegal instruction at run time (compute-sanitizer; scripts/check_sass_ur.py scans for it).
  (void)policy;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void red4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void fma4(float4& a, float s, const float4& x) {
  a.x = fmaf(s, x.x, a.x);
  a.y = fmaf(s, x.y, a.y);
  a.z = fmaf(s, x.z, a.z);
  a.w = fmaf(s, x.w, a.w);
}
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 shfl_xor4(const float4& v, int m) {
  return make_float4(__shfl_xor_sync(FULL, v.x, m), __shfl_xor_sync(FULL, v.y, m), __shfl_xor_sync(FULL, v.z, m),
                     __shfl_xor_sync(FULL, v.w, m));
}
__device__ __forceinline__ float comp(const float4& v, int c) {  // c is a compile-time constant after unrolling
  return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w));
}

// S  : block size (4, 8 or 16): S/4 consecutive lanes share a block and exchange their quads with shuffles
// NV : quads per lane; a column slab is NV*128 columns (NV*512 bytes per gathered row)
// NW : warps per CTA (every warp is its own producer: no cross-warp synchronisation anywhere)
// NG : group buffers per warp.  A GROUP is GS = 8 consecutive messages: its rows are requested together -- lane g
//      issues the bulk copy of the group's g-th row -- and land on ONE mbarrier (expect_tx = the group's bytes), so
//      the issue and wait overheads are paid once per 8 messages; group i + NG - 1 is requested when group i starts
//      to be consumed, i.e. 8 (NG - 2) .. 8 (NG - 1) rows per warp are in flight.
// TAIL : the last slab is narrower than NV*128 columns (d % (NV*128) != 0): lanes past the row end read zeros
// GS : messages per group (8 or 4)
// MODE : 0 = TMA bulk copies (cp.async.bulk, one per row, issued by the group's lanes through the uniform datapath:
//            ~10 issue slots per copy), 1 = cp.async (LDGSTS.128: the whole warp copies a row slab with ONE
//            instruction per 512 B; completion by commit/wait groups).  The fused backward is issue-bound
//            (profiles/r2_ncu_staged.md: 69 % issue-active), where the cheaper request path of mode 1 pays.
template <int S, int NV, bool FUSE_DW, bool TAIL, int NW, int NG, int GS, int MODE>
__global__ void __launch_bounds__(NW * 32, 1)
    k_block_stg(const WorkItem* __restrict__ items, int n_items, int n_slabs, const int32_t* __restrict__ r_row,
                const int32_t* __restrict__ r_nbr, const float* __restrict__ r_norm, const float* __restrict__ X, int ldx,
                int d, const float* __restrict__ Wt, float* __restrict__ out, const float* __restrict__ Hrow, int ldh,
                float* __restrict__ dWt, int* __restrict__ next_unit) {
  static_assert(S == 4 || S == 8 || S == 16, "block size");
  static_assert(GS == 4 || GS == 8, "group size");
  static_assert(NG >= 2 && (NG - 1) * GS <= 32, "the fetch cursor must stay within the next index batch");
  constexpr int G = S / 4;                      // lanes per block
  constexpr int SLAB_B = NV * 512;              // bytes of one row slab
  constexpr int RING_B = NG * GS * SLAB_B;      // bytes of one ring (gathered rows; FUSE_DW: a second one for H rows)
  constexpr int WARP_B = RING_B * (FUSE_DW ? 2 : 1);
  extern __shared__ __align__(128) uint8_t smem[];
  // shfl from lane 0 / __reduce_*_sync results are provably warp-uniform for the compiler: without them every
  // shuffle and vote of the row loop is wrapped in divergence guards (BRA.DIV / BSSY / WARPSYNC, ~20 % of the loop)
  const int warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t gring_a = smem_u32(smem) + (uint32_t)warp * WARP_B;
  const uint32_t hring_a = gring_a + RING_B;
  const uint32_t full_a = smem_u32(smem) + (uint32_t)NW * WARP_B + (uint32_t)warp * NG * 8;
  if (lane == 0) {
    for (int s = 0; s < NG; ++s) mbar_init_a(full_a + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const uint64_t pol = policy_evict_first();
  const int rel = lane & (G - 1);  // position of this lane's quad inside its block
  uint32_t gcnt = 0;               // groups requested so far by this warp (buffer = gcnt % NG, phase = gcnt / NG)
  uint32_t ccnt = 0;               // groups consumed so far

  // DYNAMIC work distribution: a warp takes the next (item, slab) unit from a global counter when it is done with
  // its current one.  The units are ordered by (supertile, weight id), and the vector reductions only resolve in L2
  // while all warps of the chip work inside the same few supertiles.  A static round-robin keeps them together on a
  // small graph, but at the full benchmark size (2.4 M items, 170 ms) per-SM speed differences let the warps drift
  // hundreds of supertiles apart and the reduction window falls out of L2 (measured: 0.85 ns instead of 0.46 ns per
  // message, and FASTER on 110 SMs than on 148).
  const int64_t n_units = (int64_t)n_items * n_slabs;
  for (;;) {
    int unit_l = 0;
    if (lane == 0) unit_l = atomicAdd(next_unit, 1);
    const int64_t unit = __shfl_sync(FULL, unit_l, 0);
    if (unit >= n_units) break;
    const int item = (int)(unit / n_slabs);
    const int c0 = (int)(unit % n_slabs) * (NV * 128);
    const int4 itv = __ldg(reinterpret_cast<const int4*>(items) + item);
    const int beg = __reduce_max_sync(FULL, itv.x), n = __reduce_max_sync(FULL, itv.y - itv.x),
              w = __reduce_max_sync(FULL, itv.z);
    const int ng = (n + GS - 1) / GS;
    const uint32_t vb = (uint32_t)min(NV * 128, d - c0) * 4u;  // valid bytes of this slab (d % 4 == 0)
    const float* Xc = X + c0;
    const float* Hc = FUSE_DW ? Hrow + c0 : nullptr;

    bool ok[NV];
    int col[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      col[k] = c0 + 4 * (lane + 32 * k);
      ok[k] = col[k] < d;
    }

    // ---- index batches: `cur` = the 32 messages [b0, b0 + 32) being consumed, `nxt` = the following 32 (requests
    //      run at most (NG-1)*8 <= 32 messages ahead).  start bit i: message i opens a run (row differs from its
    //      predecessor's); a run is transformed once, at its last message
    int b0 = 0;
    int cur_row = 0, cur_nbr = 0, nxt_row = 0, nxt_nbr = 0;
    float cur_nm = 0.f, nxt_nm = 0.f;
    if (lane < n) {
      cur_row = __ldg(r_row + beg + lane);
      cur_nbr = __ldg(r_nbr + beg + lane);
      cur_nm = __ldg(r_norm + beg + lane);
    }
    if (32 + lane < n) {
      nxt_row = __ldg(r_row + beg + 32 + lane);
      nxt_nbr = __ldg(r_nbr + beg + 32 + lane);
      nxt_nm = __ldg(r_norm + beg + 32 + lane);
    }
    auto start_bits = [&](int row_reg, int prev_row_last) {
      int p = __shfl_up_sync(FULL, row_reg, 1);
      if (lane == 0) p = prev_row_last;
      return __ballot_sync(FULL, row_reg != p);
    };
    uint32_t cur_start = start_bits(cur_row, -1);
    uint32_t nxt_start = start_bits(nxt_row, __shfl_sync(FULL, cur_row, 31));

    // request group gi (0 <= gi < ng, warp-uniform)
    auto request = [&](int gi) {
      const int q_rel = gi * GS - b0;           // first message of the group relative to the current batch: 0 .. 56
      const bool in_cur = q_rel < 32;           // a group never straddles batches (GS | 32)
      const int cnt = min(GS, n - gi * GS);
      const uint32_t buf = gcnt % NG;
      const uint32_t sb = ((in_cur ? cur_start : nxt_start) >> (q_rel & 31)) & ((1u << cnt) - 1u);  // run starts
      if (MODE == 0) {  // TMA: lane g < cnt issues the bulk copy of message gi*GS + g; all land on one mbarrier
        const int sl = (q_rel + (lane & (GS - 1))) & 31;
        const int src = __shfl_sync(FULL, in_cur ? cur_nbr : nxt_nbr, sl);
        uint32_t tx = (uint32_t)cnt * vb;
        int hrow = 0;
        bool st = false;
        if (FUSE_DW) {
          hrow = __shfl_sync(FULL, in_cur ? cur_row : nxt_row, sl);
          st = (sb >> (lane & (GS - 1))) & 1u;
          tx += (uint32_t)__popc(sb) * vb;
        }
        if (lane == 0) mbar_expect_tx_a(full_a + buf * 8, tx);
        __syncwarp();
        if (lane < cnt) {
          const uint32_t off = (buf * GS + lane) * SLAB_B;
          bulk_g2s_a(gring_a + off, Xc + (size_t)(uint32_t)src * (uint32_t)ldx, vb, full_a + buf * 8, pol);
          if (FUSE_DW && st)
            bulk_g2s_a(hring_a + off, Hc + (size_t)(uint32_t)hrow * (uint32_t)ldh, vb, full_a + buf * 8, pol);
        }
      } else {          // cp.async: every lane copies its own quads of every row of the group
        const int nbr_sel = in_cur ? cur_nbr : nxt_nbr;
        const int row_sel = in_cur ? cur_row : nxt_row;
        const uint32_t gdst = gring_a + buf * GS * SLAB_B + lane * 16, hdst = hring_a + buf * GS * SLAB_B + lane * 16;
#pragma unroll
        for (int j = 0; j < GS; ++j) {
          if (j < cnt) {
            const int src = __shfl_sync(FULL, nbr_sel, (q_rel + j) & 31);
            const float* xr = X + (size_t)(uint32_t)src * (uint32_t)ldx;
#pragma unroll
            for (int k = 0; k < NV; ++k)
              cp_async16(gdst + j * SLAB_B + k * 512, xr + (ok[k] ? col[k] : c0), ok[k] ? 16u : 0u, pol);
            if (FUSE_DW && ((sb >> j) & 1u)) {
              const int hrow = __shfl_sync(FULL, row_sel, (q_rel + j) & 31);
              const float* hr = Hrow + (size_t)(uint32_t)hrow * (uint32_t)ldh;
#pragma unroll
              for (int k = 0; k < NV; ++k)
                cp_async16(hdst + j * SLAB_B + k * 512, hr + (ok[k] ? col[k] : c0), ok[k] ? 16u : 0u, pol);
            }
          }
        }
        cp_async_commit();
      }
      ++gcnt;
    };
    // prologue: NG - 1 groups in flight (mode 1 keeps the commit-group count uniform with empty groups)
    for (int gi = 0; gi < NG - 1; ++gi) {
      if (gi < ng)
        request(gi);
      else if (MODE == 1)
        cp_async_commit();
    }

    // ---- weights of this (weight id, slab) -> registers, pre-arranged by exchange distance r = lane ^ source lane
    float4 wsel[G][4][NV];
    float4 acc[FUSE_DW ? G : 1][4][NV];
    const float* wr = Wt + (size_t)w * S * d;
#pragma unroll
    for (int r = 0; r < G; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          wsel[r][j][k] = ok[k] ? ldg4(wr + (size_t)(4 * (rel ^ r) + j) * d + col[k]) : zero4();
          if (FUSE_DW) acc[FUSE_DW ? r : 0][j][k] = zero4();
        }

    float4 xs[NV], hq[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) xs[k] = hq[k] = zero4();

    // one transform per run: y = T . xs over the block (quads of the lanes sharing the block arrive by shuffle);
    // the same products feed the weight gradient (xs = summed upstream gradient, hq = the run's own input row)
    auto flush = [&](int row) {
      float4 y[NV];
#pragma unroll
      for (int k = 0; k < NV; ++k) y[k] = zero4();
#pragma unroll
      for (int r = 0; r < G; ++r) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const float4 v = (r == 0) ? xs[k] : shfl_xor4(xs[k], r);  // quad of the lane at exchange distance r
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float xv = comp(v, j);
            fma4(y[k], xv, wsel[r][j][k]);
            if (FUSE_DW) fma4(acc[FUSE_DW ? r : 0][j][k], xv, hq[k]);
          }
        }
      }
      float* po = out + (size_t)(uint32_t)row * (uint32_t)d;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        if (!TAIL || ok[k]) red4(po + col[k], y[k]);
        xs[k] = zero4();                             // the next run accumulates from zero
      }
    };

    uint32_t ends_mask = 0;
    auto make_ends = [&]() {  // message t of the current batch closes its run when its successor opens one / the item ends
      const int nb = min(32, n - b0);
      ends_mask = (cur_start >> 1) | ((b0 + nb < n) ? ((nxt_start & 1u) << 31) : (1u << (nb - 1)));
    };
    make_ends();

    for (int gi = 0; gi < ng; ++gi) {
      __syncwarp();                                // every lane is done reading the buffer of group gi - 1 ...
      if (gi + NG - 1 < ng)
        request(gi + NG - 1);                      // ... which the group requested now reuses
      else if (MODE == 1)
        cp_async_commit();
      const uint32_t buf = ccnt % NG;
      if (MODE == 0) {
        mbar_wait_a(full_a + buf * 8, (ccnt / NG) & 1u);
      } else {
        cp_async_wait<NG - 1>();                   // this lane's copies of group gi have landed ...
        __syncwarp();                              // ... and so have everyone else's
      }
      ++ccnt;
      const uint32_t gbase = gring_a + buf * GS * SLAB_B + lane * 16;
      const uint32_t hbase = hring_a + buf * GS * SLAB_B + lane * 16;
      const int cnt = min(GS, n - gi * GS);
      const int tb0 = gi * GS - b0;
      for (int t = 0; t < cnt; ++t) {
        const int tb = tb0 + t;
        const float nm = __shfl_sync(FULL, cur_nm, tb);
        const bool st = (cur_start >> tb) & 1u;
        float4 x[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          x[k] = lds4(gbase + t * SLAB_B + k * 512);
          if (MODE == 0 && TAIL && !ok[k]) x[k] = zero4();  // TMA copies stop at the row end (cp.async zero-fills)
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) fma4(xs[k], nm, x[k]);   // xs is zero at the start of a run (flush clears it)
        if (FUSE_DW && st) {  // the run's own input row
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            hq[k] = lds4(hbase + t * SLAB_B + k * 512);
            if (MODE == 0 && TAIL && !ok[k]) hq[k] = zero4();
          }
        }
        if ((ends_mask >> tb) & 1u) flush(__shfl_sync(FULL, cur_row, tb));
      }
      if (tb0 + GS == 32 && gi + 1 < ng) {  // the batch is used up: rotate
        b0 += 32;
        cur_row = nxt_row;
        cur_nbr = nxt_nbr;
        cur_nm = nxt_nm;
        cur_start = nxt_start;
        nxt_row = nxt_nbr = 0;
        nxt_nm = 0.f;
        if (b0 + 32 + lane < n) {
          nxt_row = __ldg(r_row + beg + b0 + 32 + lane);
          nxt_nbr = __ldg(r_nbr + beg + b0 + 32 + lane);
          nxt_nm = __ldg(r_norm + beg + b0 + 32 + lane);
        }
        nxt_start = start_bits(nxt_row, __shfl_sync(FULL, cur_row, 31));
        make_ends();
      }
    }
    if (FUSE_DW) {
#pragma unroll
      for (int r = 0; r < G; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float* pw = dWt + ((size_t)w * S + 4 * (rel ^ r) + j) * d;
#pragma unroll
          for (int k = 0; k < NV; ++k)
            if (!TAIL || ok[k]) red4(pw + col[k], acc[FUSE_DW ? r : 0][j][k]);
        }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------
// TEAM variant: the T = n_slabs warps that own the column slabs of one work item share ONE staged copy of every
// gathered row.  At the full benchmark size (10 M nodes: H is 20 GB) the per-warp kernel above slowed down 1.9x
// against its small-graph rate: every 512 B slab request is its own random access into a 20 GB table, i.e. its
// own address translation, and 4 warps request the 4 slabs of a row at different times.  Here the team's leader
// warp requests each row ONCE (one 2 KB TMA bulk copy: one translation, one descriptor, a quarter of the request
// instructions) into a ring all T warps read; a group's buffer returns to the leader through an `empty` mbarrier
// the T warps arrive on (the classic multi-consumer TMA pipeline: full = transaction barrier, empty = count T).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int S, int NV, bool FUSE_DW, bool TAIL, int T, int NTEAMS, int NG, int GS>
__global__ void __launch_bounds__(NTEAMS * T * 32, 1)
    k_block_team(const WorkItem* __restrict__ items, int n_items, const int32_t* __restrict__ r_row,
                 const int32_t* __restrict__ r_nbr, const float* __restrict__ r_norm, const float* __restrict__ X, int ldx,
                 int d, const float* __restrict__ Wt, float* __restrict__ out, const float* __restrict__ Hrow, int ldh,
                 float* __restrict__ dWt, int* __restrict__ next_unit) {
  static_assert(S == 4 || S == 8 || S == 16, "block size");
  static_assert(GS == 4 || GS == 8, "group size");
  static_assert(NG >= 2 && (NG - 1) * GS <= 32, "the fetch cursor must stay within the next index batch");
  constexpr int G = S / 4;                      // lanes per block
  constexpr int ROW_B = T * NV * 512;           // bytes of one staged row slot (>= d * 4)
  __shared__ int next_item[NTEAMS][2];          // the team's current item, double-buffered across iterations
  constexpr int RING_B = NG * GS * ROW_B;
  constexpr int TEAM_B = RING_B * (FUSE_DW ? 2 : 1);
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // shfl: provably warp-uniform
  const int team = warp / T, sw = warp % T;     // sw = this warp's column slab; the sw == 0 warp is the team's leader
  const uint32_t gring_a = smem_u32(smem) + (uint32_t)team * TEAM_B;
  const uint32_t hring_a = gring_a + RING_B;
  const uint32_t full_a = smem_u32(smem) + (uint32_t)NTEAMS * TEAM_B + (uint32_t)team * NG * 16;
  const uint32_t empty_a = full_a + NG * 8;
  if (sw == 0 && lane == 0) {
    for (int s = 0; s < NG; ++s) {
      mbar_init_a(full_a + s * 8, 1);
      mbar_init_a(empty_a + s * 8, T);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint64_t pol = policy_evict_first();
  const int rel = lane & (G - 1);
  const int c0 = sw * (NV * 128);
  const uint32_t rowb = (uint32_t)d * 4u;       // bytes of one gathered row
  uint32_t gcnt = 0;                            // groups requested so far by this team (leader's counter)
  uint32_t ccnt = 0;                            // groups consumed so far by this warp

  bool ok[NV];
  int col[NV];
  char* outb[NV];                               // this lane's column of row 0 of `out`: row r is outb + r * rowb
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    col[k] = c0 + 4 * (lane + 32 * k);
    ok[k] = col[k] < d;
    outb[k] = reinterpret_cast<char*>(out) + (size_t)(ok[k] ? col[k] : 0) * 4u;
  }

  // dynamic work distribution (see k_block_stg): the leader takes the team's next item from the global counter, the
  // T warps meet at one named barrier per item (slot it & 1 cannot be overwritten before every warp has passed the
  // following barrier)
  for (int it = 0;; ++it) {
    if (sw == 0 && lane == 0) next_item[team][it & 1] = atomicAdd(next_unit, 1);
    asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(T * 32) : "memory");
    // __reduce_*_sync returns in a uniform register: the compiler then knows the item loop and everything keyed on
    // (beg, n, w) is warp-convergent and drops the divergence guards around every shuffle / vote of the row loop
    const int item = __reduce_max_sync(FULL, next_item[team][it & 1]);
    if (item >= n_items) break;
    const int4 itv = __ldg(reinterpret_cast<const int4*>(items) + item);
    const int beg = __reduce_max_sync(FULL, itv.x), n = __reduce_max_sync(FULL, itv.y - itv.x),
              w = __reduce_max_sync(FULL, itv.z);
    const int ng = (n + GS - 1) / GS;

    int b0 = 0;
    int cur_row = 0, cur_nbr = 0, nxt_row = 0, nxt_nbr = 0;
    float cur_nm = 0.f, nxt_nm = 0.f;
    if (lane < n) {
      cur_row = __ldg(r_row + beg + lane);
      cur_nbr = __ldg(r_nbr + beg + lane);
      cur_nm = __ldg(r_norm + beg + lane);
    }
    if (32 + lane < n) {
      nxt_row = __ldg(r_row + beg + 32 + lane);
      nxt_nbr = __ldg(r_nbr + beg + 32 + lane);
      nxt_nm = __ldg(r_norm + beg + 32 + lane);
    }
    auto start_bits = [&](int row_reg, int prev_row_last) {
      int p = __shfl_up_sync(FULL, row_reg, 1);
      if (lane == 0) p = prev_row_last;
      return __ballot_sync(FULL, row_reg != p);
    };
    uint32_t cur_start = start_bits(cur_row, -1);
    uint32_t nxt_start = start_bits(nxt_row, __shfl_sync(FULL, cur_row, 31));

    // leader only: request group gi -- lane g < cnt issues ONE bulk copy of the whole row of message gi*GS + g
    auto request = [&](int gi) {
      const int q_rel = gi * GS - b0;
      const bool in_cur = q_rel < 32;
      const int cnt = min(GS, n - gi * GS);
      const uint32_t buf = gcnt % NG;
      const uint32_t sb = ((in_cur ? cur_start : nxt_start) >> (q_rel & 31)) & ((1u << cnt) - 1u);
      const int sl = (q_rel + (lane & (GS - 1))) & 31;
      const int src = __shfl_sync(FULL, in_cur ? cur_nbr : nxt_nbr, sl);
      uint32_t tx = (uint32_t)cnt * rowb;
      int hrow = 0;
      bool st = false;
      if (FUSE_DW) {
        hrow = __shfl_sync(FULL, in_cur ? cur_row : nxt_row, sl);
        st = (sb >> (lane & (GS - 1))) & 1u;
        tx += (uint32_t)__popc(sb) * rowb;
      }
      if (gcnt >= NG) mbar_wait_a(empty_a + buf * 8, ((gcnt / NG) - 1u) & 1u);  // all T warps left the buffer
      if (lane == 0) mbar_expect_tx_a(full_a + buf * 8, tx);
      __syncwarp();
      if (lane < cnt) {
        const uint32_t off = (buf * GS + lane) * ROW_B;
        bulk_g2s_a(gring_a + off, X + (size_t)(uint32_t)src * (uint32_t)ldx, rowb, full_a + buf * 8, pol);
        if (FUSE_DW && st)
          bulk_g2s_a(hring_a + off, Hrow + (size_t)(uint32_t)hrow * (uint32_t)ldh, rowb, full_a + buf * 8, pol);
      }
      ++gcnt;
    };
    if (sw == 0)
      for (int gi = 0; gi < min(NG - 1, ng); ++gi) request(gi);

    float4 wsel[G][4][NV];
    float4 acc[FUSE_DW ? G : 1][4][NV];
    const float* wr = Wt + (size_t)w * S * d;
#pragma unroll
    for (int r = 0; r < G; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          wsel[r][j][k] = ok[k] ? ldg4(wr + (size_t)(4 * (rel ^ r) + j) * d + col[k]) : zero4();
          if (FUSE_DW) acc[FUSE_DW ? r : 0][j][k] = zero4();
        }

    float4 xs[NV], hq[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) xs[k] = hq[k] = zero4();

    auto flush = [&](int row) {
      float4 y[NV];
#pragma unroll
      for (int k = 0; k < NV; ++k) y[k] = zero4();
#pragma unroll
      for (int r = 0; r < G; ++r) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const float4 v = (r == 0) ? xs[k] : shfl_xor4(xs[k], r);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float xv = comp(v, j);
            fma4(y[k], xv, wsel[r][j][k]);
            if (FUSE_DW) fma4(acc[FUSE_DW ? r : 0][j][k], xv, hq[k]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        if (!TAIL || ok[k]) red4(reinterpret_cast<float*>(outb[k] + (size_t)(uint32_t)row * rowb), y[k]);
        xs[k] = zero4();                              // the next run accumulates from zero
      }
    };

    uint32_t ends_mask = 0;
    auto make_ends = [&]() {
      const int nb = min(32, n - b0);
      ends_mask = (cur_start >> 1) | ((b0 + nb < n) ? ((nxt_start & 1u) << 31) : (1u << (nb - 1)));
    };
    make_ends();

    for (int gi = 0; gi < ng; ++gi) {
      if (sw == 0 && gi + NG - 1 < ng) request(gi + NG - 1);
      const uint32_t buf = ccnt % NG;
      mbar_wait_a(full_a + buf * 8, (ccnt / NG) & 1u);
      ++ccnt;
      const uint32_t gbase = gring_a + buf * GS * ROW_B + (uint32_t)c0 * 4u + lane * 16;
      const uint32_t hbase = hring_a + buf * GS * ROW_B + (uint32_t)c0 * 4u + lane * 16;
      const int cnt = min(GS, n - gi * GS);
      const int tb0 = gi * GS - b0;
      for (int t = 0; t < cnt; ++t) {
        const int tb = tb0 + t;
        const float nm = __shfl_sync(FULL, cur_nm, tb);
        const bool st = (cur_start >> tb) & 1u;
        float4 x[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          x[k] = lds4(gbase + t * ROW_B + k * 512);
          if (TAIL && !ok[k]) x[k] = zero4();  // bytes past the row end were not copied
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) fma4(xs[k], nm, x[k]);   // xs is zero at the start of a run (flush clears it)
        if (FUSE_DW && st) {
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            hq[k] = lds4(hbase + t * ROW_B + k * 512);
            if (TAIL && !ok[k]) hq[k] = zero4();
          }
        }
        if ((ends_mask >> tb) & 1u) flush(__shfl_sync(FULL, cur_row, tb));
      }
      __syncwarp();                                   // every lane has consumed its loads of this buffer
      if (lane == 0) mbar_arrive_a(empty_a + buf * 8);
      if (tb0 + GS == 32 && gi + 1 < ng) {
        b0 += 32;
        cur_row = nxt_row;
        cur_nbr = nxt_nbr;
        cur_nm = nxt_nm;
        cur_start = nxt_start;
        nxt_row = nxt_nbr = 0;
        nxt_nm = 0.f;
        if (b0 + 32 + lane < n) {
          nxt_row = __ldg(r_row + beg + b0 + 32 + lane);
          nxt_nbr = __ldg(r_nbr + beg + b0 + 32 + lane);
          nxt_nm = __ldg(r_norm + beg + b0 + 32 + lane);
        }
        nxt_start = start_bits(nxt_row, __shfl_sync(FULL, cur_row, 31));
        make_ends();
      }
    }
    if (FUSE_DW) {
#pragma unroll
      for (int r = 0; r < G; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float* pw = dWt + ((size_t)w * S + 4 * (rel ^ r) + j) * d;
#pragma unroll
          for (int k = 0; k < NV; ++k)
            if (!TAIL || ok[k]) red4(pw + col[k], acc[FUSE_DW ? r : 0][j][k]);
        }
    }
  }
}

// One zeroed int per launch for the dynamic work distribution: a ring of 256 counters per device, the launch takes the
// next slot and clears it on its own stream (stream-ordered before the kernel).  Launches of this library that are
// concurrently in flight on different streams therefore never share a counter unless more than 256 are pending.
int* next_counter_slot(cudaStream_t st) {
  static int* ring[64] = {nullptr};
  static unsigned cursor[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!ring[dev] && cudaMalloc((void**)&ring[dev], 256 * sizeof(int)) != cudaSuccess) {
    ring[dev] = nullptr;
    return nullptr;
  }
  int* slot = ring[dev] + (cursor[dev]++ & 255u);
  if (cudaMemsetAsync(slot, 0, sizeof(int), st) != cudaSuccess) return nullptr;
  return slot;
}

template <int S, int NV, bool FUSE, bool TAIL, int T, int NTEAMS, int NG, int GS>
int launch_team_t(const WorkItem* items, int n_items, const int32_t* r_row, const int32_t* r_nbr, const float* r_norm,
                  const float* X, int ldx, int d, const float* Wt, float* out, const float* Hrow, int ldh, float* dWt,
                  cudaStream_t st) {
  constexpr int ROW_B = T * NV * 512;
  constexpr int smem = NTEAMS * NG * GS * ROW_B * (FUSE ? 2 : 1) + NTEAMS * NG * 16;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  auto kern = k_block_team<S, NV, FUSE, TAIL, T, NTEAMS, NG, GS>;
  static bool attr_set = false;
  if (!attr_set) {
    int rc = rgcn_check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                             "cudaFuncSetAttribute(team smem)");
    if (rc) return rc;
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (const char* e = std::getenv("RGCN_STG_SMS")) sms = std::max(1, std::min(sms, std::atoi(e)));  // experiment knob
  int grid = std::min((n_items + NTEAMS - 1) / NTEAMS, sms);  // persistent: one CTA per SM
  if (grid < 1) grid = 1;
  int* counter = next_counter_slot(st);
  if (!counter) {
    rgcn_set_error("staged block kernel: cannot allocate the work counter");
    return RGCN_ERR_CUDA;
  }
  kern<<<grid, NTEAMS * T * 32, smem, st>>>(items, n_items, r_row, r_nbr, r_norm, X, ldx, d, Wt, out, Hrow, ldh, dWt,
                                            counter);
  ++g_rgcn_launches;
  return rgcn_check_cuda(cudaGetLastError(), "k_block_team");
}

template <int S, int NV, bool FUSE, bool TAIL, int NW, int NG, int GS, int MODE>
int launch_stg_t(const WorkItem* items, int n_items, const int32_t* r_row, const int32_t* r_nbr, const float* r_norm,
                 const float* X, int ldx, int d, const float* Wt, float* out, const float* Hrow, int ldh, float* dWt,
                 cudaStream_t st) {
  constexpr int SLAB_B = NV * 512;
  constexpr int smem = NW * NG * GS * SLAB_B * (FUSE ? 2 : 1) + NW * NG * 8;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  auto kern = k_block_stg<S, NV, FUSE, TAIL, NW, NG, GS, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    int rc = rgcn_check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                             "cudaFuncSetAttribute(staged smem)");
    if (rc) return rc;
    attr_set = true;
  }
  const int n_slabs = (d + NV * 128 - 1) / (NV * 128);
  const int64_t units = (int64_t)n_items * n_slabs;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (const char* e = std::getenv("RGCN_STG_SMS")) sms = std::max(1, std::min(sms, std::atoi(e)));  // experiment knob
  int grid = (int)std::min<int64_t>((units + NW - 1) / NW, sms);  // persistent: one CTA per SM
  if (grid < 1) grid = 1;
  int* counter = next_counter_slot(st);
  if (!counter) {
    rgcn_set_error("staged block kernel: cannot allocate the work counter");
    return RGCN_ERR_CUDA;
  }
  kern<<<grid, NW * 32, smem, st>>>(items, n_items, n_slabs, r_row, r_nbr, r_norm, X, ldx, d, Wt, out, Hrow, ldh, dWt,
                                    counter);
  ++g_rgcn_launches;
  return rgcn_check_cuda(cudaGetLastError(), "k_block_stg");
}

}  // namespace

bool block_stg_supported(int d, int s) { return (s == 4 || s == 8 || s == 16) && d % s == 0 && d % 4 == 0; }

int launch_block_stg(const WorkItem* items, int n_items, const int32_t* r_row, const int32_t* r_nbr,
                     const float* r_norm, const float* X, int ldx, int d, int s, const float* Wt, float* out,
                     const float* Hrow, int ldh, float* dWt, cudaStream_t st) {
  if (n_items == 0) return RGCN_OK;
  if (!block_stg_supported(d, s) || (ldx % 4) != 0 || (dWt && (ldh % 4) != 0)) {
    rgcn_set_error("staged block kernel: unsupported shape");
    return RGCN_ERR_INVALID;
  }
  const bool fuse = dWt != nullptr;
#define ARGS items, n_items, r_row, r_nbr, r_norm, X, ldx, d, Wt, out, Hrow, ldh, dWt, st
#define STG(S_, NV_, FUSE_, NW_, NG_, GS_, MODE_)                                                  \
  do {                                                                                             \
    if (d % ((NV_)*128) != 0) return launch_stg_t<S_, NV_, FUSE_, true, NW_, NG_, GS_, MODE_>(ARGS); \
    return launch_stg_t<S_, NV_, FUSE_, false, NW_, NG_, GS_, MODE_>(ARGS);                         \
  } while (0)
  // Configurations (one persistent CTA per SM; registers from -Xptxas -v, budgeted by ptxas for the block rounded up
  // to 4 warps; shared memory = warps x groups x GS rows x slab bytes, twice that for the fused backward).
  // RGCN_STG_FWD / RGCN_STG_BWD pick among the measured variants of the s = 8 kernels (A/B knobs):
  //   forward  0: 2 quads/lane, 12 warps x 2 groups of 8, TMA     2: 1 quad/lane, 16 warps x 3 groups, TMA
  //            3: as 2 with cp.async (default when the team kernel does not apply)
  //   backward 0: 1 quad/lane, 12 warps x 2 groups of 8, TMA      1: as 0 with cp.async (default)
  //            3: 2 quads/lane, 8 warps x 2 groups of 4, TMA
  // NOT offered: 2 quads/lane with cp.async.  ptxas 12.9 miscompiles those instantiations -- the LDGSTS with an L2
  // cache hint comes out as `[R0+UR0], desc[UR1]` with UR0/UR1 never written, and faults with "illegal
  // instruction" (compute-sanitizer, round 2); scripts/check_sass_ur.py and tests/test_cabi_host.py scan the built
  // library for that pattern.
  int fwd_cfg = -1, bwd_cfg = -1;
  if (const char* e = std::getenv("RGCN_STG_FWD")) fwd_cfg = std::atoi(e);
  if (const char* e = std::getenv("RGCN_STG_BWD")) bwd_cfg = std::atoi(e);
  // TEAM kernels (one staged copy of a row shared by the 4 slab warps of an item): rows of 4 x 128 columns.
  // Default for those widths; RGCN_STG_TEAM=0 falls back to the per-warp rings (A/B knob).
  bool team = d > 384 && d <= 512 && (s == 4 || s == 8) && ldx % 4 == 0;
  if (const char* e = std::getenv("RGCN_STG_TEAM")) team = team && std::atoi(e) != 0;
  if (team && !(fuse ? bwd_cfg >= 0 : fwd_cfg >= 0)) {
#define TEAM(S_, FUSE_, NTEAMS_, NG_)                                                                   \
  do {                                                                                                  \
    if (d % 128 != 0) return launch_team_t<S_, 1, FUSE_, true, 4, NTEAMS_, NG_, 8>(ARGS);                \
    return launch_team_t<S_, 1, FUSE_, false, 4, NTEAMS_, NG_, 8>(ARGS);                                 \
  } while (0)
    if (s == 8) {
      if (!fuse) TEAM(8, false, 4, 3);   // 16 warps, 4 teams x 3 groups x 8 rows x 2 KB = 192 KB
      TEAM(8, true, 3, 2);               // 12 warps, 3 teams x 2 rings x 2 groups x 8 rows x 2 KB = 192 KB
    } else {
      if (!fuse) TEAM(4, false, 4, 3);
      TEAM(4, true, 3, 2);
    }
#undef TEAM
  }
  if (s == 4) {
    if (!fuse) {
      STG(4, 1, false, 16, 3, 8, 1);
    }
    STG(4, 1, true, 12, 2, 8, 1);
  } else if (s == 8) {
    if (!fuse) {
      if (d <= 128 || fwd_cfg == 2) STG(8, 1, false, 16, 3, 8, 0);
      if (fwd_cfg == 0) STG(8, 2, false, 12, 2, 8, 0);
      STG(8, 1, false, 16, 3, 8, 1);
    }
    if (d > 128 && bwd_cfg == 3) STG(8, 2, true, 8, 2, 4, 0);
    if (bwd_cfg == 0) STG(8, 1, true, 12, 2, 8, 0);
    STG(8, 1, true, 12, 2, 8, 1);
  } else {
    if (!fuse) STG(16, 1, false, 16, 3, 8, 1);
    STG(16, 1, true, 8, 2, 8, 1);
  }
#undef ARGS
#undef STG
#undef STG
}
