// gemm_tf32x3.cu -- fp32-accurate dense GEMM on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
//   C[M,N] (+)= A[M,K] * B^T      A row-major (K contiguous), B given K-major as Bt[N,K]
//
// Used for the self-loop terms of the R-GCN layer (H @ W_self, dS @ W_self^T; reference:
// gcn_basis.py:70-71 / gcn_basis_concat.py:65-66 `tf.matmul`).  The 1e-4 parity bar rules out a
// single TF32 pass (10-bit mantissa), so every fp32 operand is split a = a_hi + a_lo with both parts exact
// TF32 values (the streamed operand by truncation in the producers, the small pre-split operand with
// cvt.rna.tf32.f32) and three MMAs are issued per K-step:  a_hi*b_hi + a_hi*b_lo + a_lo*b_hi
// (the a_lo*b_lo term is below fp32 rounding).
//
// Structure of the NT kernel (PERSISTENT: min(tiles, SMs) CTAs of 416 threads, 1 CTA/SM, 128x128 tiles):
//   warps 0-7   producers: global fp32 -> registers -> (hi, lo) -> st.shared in the UMMA canonical
//               K-major SWIZZLE_128B layout for A; cp.async of the pre-split Bt_hi / Bt_lo tiles;
//               fence.proxy.async + mbarrier arrive.  They run straight on into the next tile.
//   warp 8      one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=128, K=8), 12 per 32-wide
//               K block, frees smem stages and publishes the tile's accumulators with tcgen05.commit.
//   warps 9-12  epilogue: tcgen05.ld of the accumulator rows (one TMEM lane quarter each) -> global
//               stores (or the rank-counting epilogue), then `drained` so the set can be reused.
//   3-stage smem ring (64 KB per stage); TMEM = two accumulator sets (tile parity) x {big, small} x 128
//   columns, so the MMAs of tile t+1 overlap the epilogue of tile t.
// The TN kernel (below) keeps one tile (x split-K) per CTA, 288 threads, four accumulators.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 32;  // BK floats = 128 bytes = one swizzle row
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;     // 16 KB (A_hi, A_lo, B_hi, B_lo each)
constexpr int STAGE_BYTES = 4 * TILE_BYTES; // 64 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + alignment slack
constexpr int N_PRODUCERS = 256;  // 8 producer/epilogue warps + 1 MMA warp
constexpr int N_THREADS = N_PRODUCERS + 32;
constexpr int MMA_WARP = N_PRODUCERS / 32;
// TMEM budget of both kernels: 4 x 128 columns.  The big hi*hi products and the small cross terms are always
// accumulated separately and summed in fp32 registers in the epilogue (the tensor core adds into its accumulator
// with truncation, so same-magnitude additions per accumulator keep the result at SGEMM-level accuracy).  The TN
// kernel additionally alternates between two accumulators of each kind per K block (N_ACC = 4 for one tile); the
// persistent NT kernel uses one of each per tile and the other half of TMEM for the next tile.
constexpr int N_ACC = 4;
constexpr int TMEM_COLS = N_ACC * BN;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes)
               : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1)
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B) | [46,48) version = 1 (sm_100)
//   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 (bits 4-5 = 1), A=B=TF32 (bits 7-9,
// 10-12 = 2), both K-major (bits 15, 16 = 0), N>>3 at [17,23), M>>4 at [24,29).
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// a = hi + lo with hi = RN_tf32(a) and lo = RN_tf32(a - hi): both exactly representable in TF32, so
// the tensor core's own fp32->tf32 conversion is exact and the split error is <= 2^-22 |a|, unbiased.
// Producer-side split by TRUNCATION: hi = a with the 13 low mantissa bits cleared, lo = (a - hi) (exact in fp32) with
// its low bits cleared.  Three integer/FP instructions per element instead of two cvt.rna.tf32 (which sm_100 expands
// to ~5 instructions each: the round-to-nearest split made the producers execute ~360 instructions per thread and
// k-block and capped the tensor pipe at 35 % active, ncu round 2).  Both parts are exact TF32 values; the split error
// is |a - hi - lo| < 2^-20 |a| (round-to-nearest: 2^-22), still two orders below the 1e-4 bar of the layer and inside
// the 1e-5 bar of tests/test_gpu_gemm.py.  The small K-major operand is still split with round-to-nearest (once).
__device__ __forceinline__ void split_tf32_trunc(float a, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
  lo = __uint_as_float(__float_as_uint(a - hi) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32(float a, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(a));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(a - hi));
  lo = __uint_as_float(l);
}

__device__ __forceinline__ uint32_t swz(int r, int c) {  // byte offset of 16 B chunk c of tile row r
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// EPI = 1: RANKING epilogue (all-entity scoring of the DistMult decoder, decoders/bilinear_diag.py:51-61, fused with
// the rank counts of common/evaluation.py:148-159): the C tile is never written.  Row m of A is a query (e1*r or r*e2),
// row n of Bt an entity code; each energy goes through the reference's float32 sigmoid and is compared with the gold
// entity's score; a 32-column chunk yields a 32-bit "score >= gold" mask whose popcount is the chunk's contribution to
// the raw rank, and popcount(mask & known bits) its contribution to the filtered correction.
struct RankEpi {
  const float* gold_sig;       // [M] sigmoid(energy of the gold entity)
  const int32_t* gold_col;     // [M] gold entity id (always counted: score >= itself)
  const uint32_t* known;       // [M, words] bit v = entity v is a known true answer (or nullptr)
  int words;                   // ceil(N / 32)
  int32_t* raw_cnt;            // [M] += #{v : score_v >= gold}
  int32_t* known_cnt;          // [M] += #{known v : score_v >= gold}
};
__device__ __forceinline__ float sigmoid_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

// PERSISTENT, three roles: 8 producer warps, 1 MMA warp, 4 epilogue warps.  A CTA walks the tiles blockIdx.x,
// blockIdx.x + gridDim.x, ... (N tiles fastest, see below); the shared-memory stages and their barrier phases run on
// across tile boundaries, so the producers fill the pipeline of tile t+1 (global-load latency included) while the MMAs
// of tile t drain and the epilogue warps move its accumulators out of TMEM.  (Round 2 profile of the one-tile-per-CTA
// version on the 10 M x 512 x 512 self-loop GEMM: tensor pipe 35 % active, producers parked on their first loads --
// with K = 512 a tile is only 16 k-blocks, and every tile paid pipeline fill + epilogue with the tensor core idle.)
// TMEM holds TWO accumulator sets (tile parity): a tile uses one big (hi*hi) and one small (cross terms) accumulator,
// 2 x 128 columns, so the MMAs of tile t+1 run while the epilogue warps move tile t out of the other set; the MMA warp
// only waits for `drained[set]` of tile t-2.  (A first persistent version kept round 1's four accumulators per tile --
// even/odd k-blocks separately -- in a single set: the next tile's MMAs then waited for the whole epilogue and the
// kernel was 12 % SLOWER than one tile per CTA.)
constexpr int N_EPI = 128;                                  // 4 epilogue warps: one per TMEM lane quarter
constexpr int N_THREADS_NT = N_PRODUCERS + 32 + N_EPI;      // 416
constexpr int EPI_WARP0 = MMA_WARP + 1;

template <int EPI>
__global__ void __launch_bounds__(N_THREADS_NT, 1)
    k_gemm_tf32x3(const float* __restrict__ A, int64_t lda, const float* __restrict__ Bhi,
                  const float* __restrict__ Blo, int64_t ldb, float* __restrict__ C, int64_t ldc,
                  int M, int N, int K, int accumulate, int n_tiles, RankEpi re) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar[2], drained_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B: 1024 B aligned
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Tile order, N tiles fastest: the CTAs that share an A tile (all N tiles of one M tile) run at the same time,
  // so the big operand (V x K, 20 GB at the full benchmark size) is read from HBM once and from L2 afterwards; the
  // small K-major operand (<= a few MB of hi/lo planes) lives in L2 throughout.  (Round 1 launched the M tiles
  // fastest: every N tile re-read its A tile from HBM -- ncu: 1.64 GB of DRAM reads for a 0.41 GB operand.)
  const int tn = (N + BN - 1) / BN;
  const int num_kb = (K + BK - 1) / BK;
  const int n_mine = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA (>= 1)
  const int total_g = n_mine * num_kb;                                                   // its k-blocks, all tiles

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], N_PRODUCERS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&accum_bar[b], 1);
      mbar_init(&drained_bar[b], N_EPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {  // TMEM allocation by one full warp; the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_smem)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp < MMA_WARP) {
    // ================= producers =================
    // Software-pipelined over the CTA's whole k-block sequence g = 0 .. total_g-1 (tile g / num_kb, block g % num_kb):
    // while block g is converted and stored, the global loads of A(g+1), A(g+2) are already in flight in registers
    // and the async copies of B(g+1) in the next smem stage -- also when g+1 belongs to the NEXT tile.
    // All addresses are carried incrementally (row pointers advance by BK floats per k-block and are rebuilt once
    // per tile): recomputing (row * ld + column) with its bounds checks for every request made the producers execute
    // ~330 instructions per thread and k-block, of which 16 were the loads and copies themselves.
    const int c = tid & 7;        // 16 B chunk within the 128 B K-row
    const int rbase = tid >> 3;   // 0..31; this thread handles tile rows rbase + 32 i
    const uint32_t soff = swz(rbase, c);  // swizzled offset of (row rbase, chunk c); row rbase + 32 i is 4096 i further
    constexpr int NR = BM / 32;
    static_assert(BM == BN, "the producers use one row schedule for both operands");

    // ---- B request stream (async copies, one block ahead)
    int b_kb = 0, b_tile = (int)blockIdx.x, b_stage = 0;
    const float* pbh[NR];
    const float* pbl[NR];
    bool bok[NR];
    auto b_set_tile = [&]() {
      const int n0 = (b_tile % tn) * BN;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int r = n0 + rbase + 32 * i;
        bok[i] = r < N;
        const size_t off = bok[i] ? ((size_t)r * ldb + c * 4) : 0;
        pbh[i] = Bhi + off;
        pbl[i] = Blo + off;
      }
    };
    b_set_tile();
    auto issue_b = [&]() {        // requests the stream's current block, then advances it
      const uint32_t b_hi = smem_base + b_stage * STAGE_BYTES + 2 * TILE_BYTES + soff, b_lo = b_hi + TILE_BYTES;
      const bool col_ok = b_kb * BK + c * 4 < K;  // K % 4 == 0: a 16 B chunk is entirely valid or entirely padding
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const bool ok = col_ok && bok[i];
        cp_async16(b_hi + 4096 * i, ok ? pbh[i] : Bhi, ok ? 16u : 0u);
        cp_async16(b_lo + 4096 * i, ok ? pbl[i] : Blo, ok ? 16u : 0u);
        pbh[i] += BK;
        pbl[i] += BK;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      b_stage = b_stage + 1 == STAGES ? 0 : b_stage + 1;
      if (++b_kb == num_kb) {
        b_kb = 0;
        b_tile += (int)gridDim.x;
        b_set_tile();
      }
    };
    // ---- A request stream (register prefetch, PF blocks ahead)
    int a_kb = 0, a_tile = (int)blockIdx.x;
    const float* pa[NR];
    bool aok[NR];
    auto a_set_tile = [&]() {
      const int m0 = (a_tile / tn) * BM;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int r = m0 + rbase + 32 * i;
        aok[i] = r < M;
        pa[i] = A + (aok[i] ? ((size_t)r * lda + c * 4) : 0);
      }
    };
    a_set_tile();
    auto load_a = [&](float4 (&v)[NR]) {   // requests the stream's current block, then advances it
      const bool col_ok = a_kb * BK + c * 4 < K;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        v[i] = (col_ok && aok[i]) ? __ldg(reinterpret_cast<const float4*>(pa[i])) : make_float4(0.f, 0.f, 0.f, 0.f);
        pa[i] += BK;
      }
      if (++a_kb == num_kb) {
        a_kb = 0;
        a_tile += (int)gridDim.x;
        a_set_tile();
      }
    };
    // PF k-blocks of A are in flight in registers (and of B as asynchronous copies) while block g is converted and
    // stored.  (The asynchronous copies of B stay ONE block ahead: they need the shared-memory stage of block g + 1,
    // and asking for the stage of g + 2 would make the producers wait for the MMAs of block g - 1 before storing block
    // g -- measured 8 % slower.  The A rows only need registers, so they run two blocks ahead.)
    constexpr int PF = 2;                    // register prefetch distance of A in k-blocks (PF + 1 buffers)
    float4 vbuf[PF + 1][NR];
    mbar_wait(&empty_bar[0], 1u);            // first use of a stage: the "previous phase" is complete
    issue_b();
    for (int p = 0; p < PF; ++p)
      if (p < total_g) load_a(vbuf[p]);
    int s = 0;                               // stage of block g
    uint32_t ph_next = 0;                    // phase bit of block g + 1's stage use
    for (int g0 = 0; g0 < total_g; g0 += PF + 1) {
#pragma unroll
      for (int u = 0; u <= PF; ++u) {
        const int g = g0 + u;
        if (g >= total_g) break;
        const bool more = g + 1 < total_g;
        const int s1 = s + 1 == STAGES ? 0 : s + 1;
        if (s1 == 0) ph_next ^= 1u;          // block g + 1 starts a new round of the stage ring
        if (more) {
          mbar_wait(&empty_bar[s1], ph_next ^ 1u);
          issue_b();
        }
        if (g + PF < total_g) load_a(vbuf[(u + PF) % (PF + 1)]);
        const uint32_t a_hi = smem_base + s * STAGE_BYTES + soff, a_lo = a_hi + TILE_BYTES;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          float4 hi, lo;
          split_tf32_trunc(vbuf[u][i].x, hi.x, lo.x);
          split_tf32_trunc(vbuf[u][i].y, hi.y, lo.y);
          split_tf32_trunc(vbuf[u][i].z, hi.z, lo.z);
          split_tf32_trunc(vbuf[u][i].w, hi.w, lo.w);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + 4096 * i), "f"(hi.x), "f"(hi.y),
                       "f"(hi.z), "f"(hi.w)
                       : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + 4096 * i), "f"(lo.x), "f"(lo.y),
                       "f"(lo.z), "f"(lo.w)
                       : "memory");
        }
        if (more)
          asm volatile("cp.async.wait_group 1;" ::: "memory");  // B(g) has landed, B(g+1) may still fly
        else
          asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor core
        mbar_arrive(&full_bar[s]);
        s = s1;
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ================= epilogue warps =================
    // a warp may only touch TMEM lanes 32*(warp % 4)..+31: the four epilogue warps take one lane quarter each and
    // all BN columns of it
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int ti = 0; ti < n_mine; ++ti) {
      const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
      const int m0 = (tile / tn) * BM, n0 = (tile % tn) * BN;
      const int row = m0 + (warp & 3) * 32 + lane;  // TMEM lane == accumulator row
      int rank_raw = 0, rank_known = 0;
      float gold_s = 0.f;
      int gold_c = -1;
      if (EPI == 1 && row < M) {
        gold_s = __ldg(re.gold_sig + row);
        gold_c = __ldg(re.gold_col + row);
      }
      const int set = ti & 1;                       // accumulator set of this tile: columns [set*2*BN, +2*BN)
      mbar_wait(&accum_bar[set], (uint32_t)(ti >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int cb = 0; cb < BN; cb += 32) {
        uint32_t r[32];
        float sum[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) sum[q] = 0.f;
        // small (cross-term) accumulator first, then the big one
        for (int a = 1; a >= 0; --a) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
                "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
                "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(taddr + (uint32_t)((2 * set + a) * BN + cb)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int q = 0; q < 32; ++q) sum[q] += __uint_as_float(r[q]);
        }
        if (cb == BN - 32) {  // this thread's last read of the accumulators: the MMAs of the next tile may overwrite
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(&drained_bar[set]);
        }
        if (EPI == 1) {
          if (row < M && n0 + cb < N) {
            uint32_t bits = 0;
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (n0 + cb + q < N && sigmoid_ref(sum[q]) >= gold_s) bits |= 1u << q;
            const int gq = gold_c - (n0 + cb);
            if (gq >= 0 && gq < 32) bits |= 1u << gq;   // the gold entity always scores >= itself
            rank_raw += __popc(bits);
            if (re.known) rank_known += __popc(bits & __ldg(re.known + (size_t)row * re.words + ((n0 + cb) >> 5)));
          }
          continue;
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) r[q] = __float_as_uint(sum[q]);
        if (row < M) {
          float* crow = C + (size_t)row * ldc + n0 + cb;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int ncol = n0 + cb + 4 * q;
            if (ncol < N) {  // N % 4 == 0
              float4 o = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                     __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
              if (accumulate) {
                const float4 old = *reinterpret_cast<const float4*>(crow + 4 * q);
                o.x += old.x;
                o.y += old.y;
                o.z += old.z;
                o.w += old.w;
              }
              *reinterpret_cast<float4*>(crow + 4 * q) = o;
            }
          }
        }
      }
      if (EPI == 1 && row < M) {
        if (rank_raw) atomicAdd(re.raw_cnt + row, rank_raw);
        if (rank_known) atomicAdd(re.known_cnt + row, rank_known);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // ================= MMA issuer (one elected lane) =================
    if (lane == 0) {
      int g = 0;
      for (int ti = 0; ti < n_mine; ++ti) {
        const int set = ti & 1;
        if (ti >= 2) {  // the epilogue warps have read tile ti - 2 out of this accumulator set
          mbar_wait(&drained_bar[set], (uint32_t)((ti >> 1) - 1) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          const int s = g % STAGES;
          const uint32_t ph = (uint32_t)(g / STAGES) & 1u;
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = smem_base + s * STAGE_BYTES, a_lo = a_hi + TILE_BYTES;
          const uint32_t b_hi = a_lo + TILE_BYTES, b_lo = b_hi + TILE_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {  // UMMA K = 8 tf32 = 32 bytes along the swizzled row
            const uint64_t dah = make_desc(a_hi + kk * 32), dal = make_desc(a_lo + kk * 32);
            const uint64_t dbh = make_desc(b_hi + kk * 32), dbl = make_desc(b_lo + kk * 32);
            const uint32_t acc_big = tmem_base + (uint32_t)(2 * set * BN);
            const uint32_t acc_small = acc_big + (uint32_t)BN;
            const uint32_t first = (kb == 0 && kk == 0) ? 0u : 1u;  // first touch overwrites
            umma_tf32(acc_small, dal, dbh, first);
            umma_tf32(acc_small, dah, dbl, 1);
            umma_tf32(acc_big, dah, dbh, first);
          }
          umma_commit(&empty_bar[s]);  // frees the smem stage when these MMAs have read it
        }
        umma_commit(&accum_bar[set]);  // this tile's accumulators are complete
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "n"(TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// C[M,N] += A^T B with A [K,M] and B [K,N] row-major: the V-long reductions of the backward pass
// (dW_self = H^T dS, basis dV = Agg^T G).  Both operands are "MN-major" for the tensor core (the
// contraction index is the slow one in memory).  For 32-bit MN-major operands the only shared-memory
// layout the tensor core accepts is SWIZZLE_128B_BASE32B (cutlass sm100_common.inl:92): a 128-byte row
// holds 32 consecutive M (or N) values of one k, FOUR k-rows form a 512-byte atom in which the 32-byte
// chunk index is XORed with (k & 3) (Swizzle<2,5,2> on byte addresses); SBO = 512 B between k-atoms
// (two per K = 8 MMA step), LBO = 4 KB between 32-wide MN blocks.  Both operands are split hi/lo in
// registers by the producers.  Split-K over gridDim.z; partial tiles are added with red.global.add.v4.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t IDESC_TN = IDESC | (1u << 15) | (1u << 16);  // a_major = b_major = MN

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(4096 >> 4) << 16) |
         ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);  // layout type 1 = SWIZZLE_128B_BASE32B
}
__device__ __forceinline__ void umma_tf32_tn(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(IDESC_TN), "r"(accumulate)
      : "memory");
}

template <int PF>
__global__ void __launch_bounds__(N_THREADS, 1)
    k_gemm_tn_tf32x3(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                     float* __restrict__ C, int64_t ldc, int M, int N, int K, int kb_per_split) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kb_total = (K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int num_kb = min(kb_per_split, kb_total - kb_begin);
  if (num_kb <= 0) return;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], N_PRODUCERS);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_smem)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp < MMA_WARP) {
    // producers: thread -> (k-row tid / 8, 16 B chunk tid % 8) of MN block i (i = 0..3).  Software-
    // pipelined like the NT kernel: the loads of block kb+1 are in flight while block kb is split and stored.
    auto load_ab = [&](int kbi, float4 (&va)[4], float4 (&vb)[4]) {
      const int k0 = (kb_begin + kbi) * BK;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // chunk id: MN block (32 floats) = i, k-row = tid / 8, 16 B chunk within the 128 B row = tid % 8:
        // a warp touches 4 consecutive k-rows x 128 B = one contiguous, conflict-free 512 B of the tile
        const int kr = tid >> 3, cm = 8 * i + (tid & 7);
        const int krow = k0 + kr;
        const bool kok = krow < K;
        va[i] = (kok && (m0 + 4 * cm < M))
                    ? __ldg(reinterpret_cast<const float4*>(A + (size_t)krow * lda + m0 + 4 * cm))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
        vb[i] = (kok && (n0 + 4 * cm < N))
                    ? __ldg(reinterpret_cast<const float4*>(B + (size_t)krow * ldb + n0 + 4 * cm))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    // PF k-blocks of both operands in flight in registers (see the NT kernel: one block of prefetch left the
    // producers latency-bound)
    float4 va[PF + 1][4], vb[PF + 1][4];
    for (int p = 0; p < PF; ++p)
      if (p < num_kb) load_ab(p, va[p], vb[p]);
    for (int kb0 = 0; kb0 < num_kb; kb0 += PF + 1) {
#pragma unroll
      for (int u = 0; u <= PF; ++u) {
        const int kbi = kb0 + u;
        if (kbi >= num_kb) break;
        const int s = kbi % STAGES;
        const uint32_t ph = (uint32_t)(kbi / STAGES) & 1u;
        if (kbi + PF < num_kb) load_ab(kbi + PF, va[(u + PF) % (PF + 1)], vb[(u + PF) % (PF + 1)]);
        mbar_wait(&empty_bar[s], ph ^ 1u);
        const uint32_t a_hi = smem_base + s * STAGE_BYTES, a_lo = a_hi + TILE_BYTES;
        const uint32_t b_hi = a_lo + TILE_BYTES, b_lo = b_hi + TILE_BYTES;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kr = tid >> 3, cm = 8 * i + (tid & 7);
          const int c8 = cm & 7;  // 16 B chunk within the 128 B row: 32 B chunk (c8 >> 1) is swizzled with k & 3
          const uint32_t off = (uint32_t)((cm >> 3) * 4096 + (kr >> 2) * 512 + (kr & 3) * 128 +
                                          ((((c8 >> 1) ^ (kr & 3)) << 5) | ((c8 & 1) << 4)));
          float4 hi, lo;
          split_tf32_trunc(va[u][i].x, hi.x, lo.x);
          split_tf32_trunc(va[u][i].y, hi.y, lo.y);
          split_tf32_trunc(va[u][i].z, hi.z, lo.z);
          split_tf32_trunc(va[u][i].w, hi.w, lo.w);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "f"(hi.x), "f"(hi.y),
                       "f"(hi.z), "f"(hi.w) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + off), "f"(lo.x), "f"(lo.y),
                       "f"(lo.z), "f"(lo.w) : "memory");
          split_tf32_trunc(vb[u][i].x, hi.x, lo.x);
          split_tf32_trunc(vb[u][i].y, hi.y, lo.y);
          split_tf32_trunc(vb[u][i].z, hi.z, lo.z);
          split_tf32_trunc(vb[u][i].w, hi.w, lo.w);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(b_hi + off), "f"(hi.x), "f"(hi.y),
                       "f"(hi.z), "f"(hi.w) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(b_lo + off), "f"(lo.x), "f"(lo.y),
                       "f"(lo.z), "f"(lo.w) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&full_bar[s]);
      }
    }
    // epilogue: add this split's tile into C
    mbar_wait(&accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = m0 + (warp & 3) * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int cb_begin = (warp >> 2) * (BN / 2);
    const int n_pair = num_kb > 1 ? 2 : 1;
#pragma unroll
    for (int cb = cb_begin; cb < cb_begin + BN / 2; cb += 32) {
      uint32_t r[32];
      float sum[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) sum[q] = 0.f;
      for (int a = N_ACC - 1; a >= 0; --a) {
        if ((a & 1) >= n_pair) continue;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
              "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
              "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
              "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
              "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr + (uint32_t)(a * BN + cb)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 32; ++q) sum[q] += __uint_as_float(r[q]);
      }
      if (row < M) {
        float* crow = C + (size_t)row * ldc + n0 + cb;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (n0 + cb + 4 * q < N)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow + 4 * q),
                         "f"(sum[4 * q]), "f"(sum[4 * q + 1]), "f"(sum[4 * q + 2]), "f"(sum[4 * q + 3])
                         : "memory");
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    if (lane == 0) {
      for (int kbi = 0; kbi < num_kb; ++kbi) {
        const int s = kbi % STAGES;
        const uint32_t ph = (uint32_t)(kbi / STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_base + s * STAGE_BYTES, a_lo = a_hi + TILE_BYTES;
        const uint32_t b_hi = a_lo + TILE_BYTES, b_lo = b_hi + TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {  // two 512 B k-atoms (8 k-rows) per MMA K-step
          const uint64_t dah = make_desc_mn(a_hi + kk * 1024), dal = make_desc_mn(a_lo + kk * 1024);
          const uint64_t dbh = make_desc_mn(b_hi + kk * 1024), dbl = make_desc_mn(b_lo + kk * 1024);
          const uint32_t acc_big = tmem_base + (uint32_t)((kbi & 1) * BN);
          const uint32_t acc_small = tmem_base + (uint32_t)((2 + (kbi & 1)) * BN);
          const uint32_t first = (kbi < 2 && kk == 0) ? 0u : 1u;
          umma_tf32_tn(acc_small, dal, dbh, first);
          umma_tf32_tn(acc_small, dah, dbl, 1);
          umma_tf32_tn(acc_big, dah, dbh, first);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&accum_bar);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "n"(TMEM_COLS)
                 : "memory");
  }
}

// Bt_hi/Bt_lo[n][k] from B: transposed = 0: B is [N,K] row-major already (K-major);
//                                  transposed = 1: B is [K,N] row-major -> transpose while splitting
__global__ void k_split_b(const float* __restrict__ B, int64_t ldb, int N, int K, int transposed,
                          float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t total = (int64_t)N * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / K), k = (int)(i % K);
    const float v = transposed ? __ldg(B + (size_t)k * ldb + n) : __ldg(B + (size_t)n * ldb + k);
    float h, l;
    split_tf32(v, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

}  // namespace

int launch_gemm_split_b(const float* B, int64_t ldb, int N, int K, int transposed, float* hi, float* lo,
                        cudaStream_t st) {
  const int64_t total = (int64_t)N * K;
  if (total == 0) return RGCN_OK;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_split_b<<<blocks, 256, 0, st>>>(B, ldb, N, K, transposed, hi, lo);
  ++g_rgcn_launches;
  return rgcn_check_cuda(cudaGetLastError(), "k_split_b");
}

// grid of the persistent NT kernel: the SM count of the current device ($RGCN_GEMM_CTAS overrides; a value >= the
// tile count gives a one-tile-per-CTA schedule)
static int64_t nt_grid_cap() {
  if (const char* e = std::getenv("RGCN_GEMM_CTAS")) return std::max<int64_t>(1, std::atoll(e));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

// register prefetch distance of the producers (k-blocks of the streamed operand in flight): $RGCN_GEMM_PF = 2 | 3 | 4
static int gemm_prefetch_distance() {
  static int pf = 0;
  if (!pf) {
    const char* e = std::getenv("RGCN_GEMM_PF");
    pf = e ? std::atoi(e) : 3;
    if (pf < 2 || pf > 4) pf = 3;
  }
  return pf;
}

int launch_gemm_tf32x3(const float* A, int64_t lda, const float* Bt_hi, const float* Bt_lo, int64_t ldb,
                       float* C, int64_t ldc, int M, int N, int K, int accumulate, cudaStream_t st) {
  if (M == 0 || N == 0) return RGCN_OK;
  if (K % 4 != 0 || N % 4 != 0 || lda % 4 != 0 || ldb % 4 != 0 || ldc % 4 != 0) {
    rgcn_set_error("gemm_tf32x3: K, N and leading dimensions must be multiples of 4");
    return RGCN_ERR_INVALID;
  }
  if (K == 0) {   // empty contraction: C = 0 (or unchanged); the kernel would publish accumulators no MMA ever wrote
    if (accumulate) return RGCN_OK;
    return rgcn_check_cuda(cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, st), "memset(C)");
  }
  static bool attr_set = false;
  if (!attr_set) {
    int rc = rgcn_check_cuda(cudaFuncSetAttribute(k_gemm_tf32x3<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  SMEM_BYTES),
                             "cudaFuncSetAttribute(gemm smem)");
    if (rc) return rc;
    attr_set = true;
  }
  const int64_t tiles = (int64_t)((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  if (tiles > 0x7fffffffLL) {
    rgcn_set_error("gemm_tf32x3: too many tiles");
    return RGCN_ERR_INVALID;
  }
  dim3 grid((unsigned)std::min<int64_t>(tiles, nt_grid_cap()));   // persistent: at most one CTA per SM
  k_gemm_tf32x3<0><<<grid, N_THREADS_NT, SMEM_BYTES, st>>>(A, lda, Bt_hi, Bt_lo, ldb, C, ldc, M, N, K, accumulate,
                                                            (int)tiles, RankEpi{});
  ++g_rgcn_launches;
  return rgcn_check_cuda(cudaGetLastError(), "k_gemm_tf32x3");
}

// Scoring GEMM with the ranking epilogue: queries Q [M,K] against the pre-split entity codes Bt [N,K]; the counts
// accumulate (+=) into raw_cnt / known_cnt (zeroed by the caller).
int launch_gemm_rank_tf32x3(const float* Q, int64_t ldq, const float* Bt_hi, const float* Bt_lo, int64_t ldb, int M,
                            int N, int K, const float* gold_sig, const int32_t* gold_col, const uint32_t* known,
                            int words, int32_t* raw_cnt, int32_t* known_cnt, cudaStream_t st) {
  if (M == 0 || N == 0) return RGCN_OK;
  if (K <= 0 || K % 4 != 0 || ldq % 4 != 0 || ldb % 4 != 0) {
    rgcn_set_error("gemm_rank_tf32x3: K > 0; K and leading dimensions must be multiples of 4");
    return RGCN_ERR_INVALID;
  }
  static bool attr_set = false;
  if (!attr_set) {
    int rc = rgcn_check_cuda(cudaFuncSetAttribute(k_gemm_tf32x3<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  SMEM_BYTES),
                             "cudaFuncSetAttribute(gemm rank smem)");
    if (rc) return rc;
    attr_set = true;
  }
  RankEpi re{gold_sig, gold_col, known, words, raw_cnt, known_cnt};
  const int64_t tiles = (int64_t)((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  if (tiles > 0x7fffffffLL) {
    rgcn_set_error("gemm_rank_tf32x3: too many tiles");
    return RGCN_ERR_INVALID;
  }
  dim3 grid((unsigned)std::min<int64_t>(tiles, nt_grid_cap()));
  k_gemm_tf32x3<1><<<grid, N_THREADS_NT, SMEM_BYTES, st>>>(Q, ldq, Bt_hi, Bt_lo, ldb, nullptr, 0, M, N, K, 0,
                                                            (int)tiles, re);
  ++g_rgcn_launches;
  return rgcn_check_cuda(cudaGetLastError(), "k_gemm_tf32x3<rank>");
}

// C[M,N] (+)= A^T B, A [K,M] row-major, B [K,N] row-major (see k_gemm_tn_tf32x3)
int launch_gemm_tn_tf32x3(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                          int M, int N, int K, int accumulate, cudaStream_t st) {
  if (M == 0 || N == 0) return RGCN_OK;
  if (M % 4 != 0 || N % 4 != 0 || lda % 4 != 0 || ldb % 4 != 0 || ldc % 4 != 0) {
    rgcn_set_error("gemm_tn_tf32x3: M, N and leading dimensions must be multiples of 4");
    return RGCN_ERR_INVALID;
  }
  if (!accumulate) {
    int rc = rgcn_check_cuda(cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, st),
                             "memset(C)");
    if (rc) return rc;
  }
  if (K == 0) return RGCN_OK;
  static bool attr_set = false;
  if (!attr_set) {
    int rc = rgcn_check_cuda(cudaFuncSetAttribute(k_gemm_tn_tf32x3<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  SMEM_BYTES),
                             "cudaFuncSetAttribute(gemm tn smem)");
    if (!rc) rc = rgcn_check_cuda(cudaFuncSetAttribute(k_gemm_tn_tf32x3<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       SMEM_BYTES), "cudaFuncSetAttribute(gemm tn smem)");
    if (rc) return rc;
    attr_set = true;
  }
  const int tm = (M + BM - 1) / BM, tn = (N + BN - 1) / BN;
  const int kb_total = (K + BK - 1) / BK;
  int splits = (2 * 148 + tm * tn - 1) / (tm * tn);  // about two waves of CTAs
  if (splits > kb_total / 4) splits = kb_total / 4;  // at least 4 K blocks per split
  if (splits < 1) splits = 1;
  const int kb_per_split = (kb_total + splits - 1) / splits;
  splits = (kb_total + kb_per_split - 1) / kb_per_split;
  dim3 grid(tm, tn, splits);
  if (gemm_prefetch_distance() >= 3)
    k_gemm_tn_tf32x3<3><<<grid, N_THREADS, SMEM_BYTES, st>>>(A, lda, B, ldb, C, ldc, M, N, K, kb_per_split);
  else
    k_gemm_tn_tf32x3<2><<<grid, N_THREADS, SMEM_BYTES, st>>>(A, lda, B, ldb, C, ldc, M, N, K, kb_per_split);
  ++g_rgcn_launches;
  return rgcn_check_cuda(cudaGetLastError(), "k_gemm_tn_tf32x3");
}
