// b2h_ppo_tma.cuh — the TMA-fed GEMM of the PPO update and the operand format it reads (included by b2h_ppo.cu).
//
// The staged kernel of b2h_ppo.cu (gemm_kernel) spends its time in seven producer warps that pull fp32 operands through
// registers, split them into tf32 hi / lo and scatter them into the tensor cores' core-matrix order: ~8000 warp
// instructions per K chunk, tensor pipe 16-20 % active (profiles/r02_ppo_gemm_staged.txt).  Here no thread touches an operand:
//
//   * every matrix a GEMM of the update reads is kept in HBM already split: two row-major planes (hi, lo), rows padded to
//     128 and columns to 32.  The kernels that PRODUCE those matrices write them that way (the gather of the minibatch, the
//     GEMM epilogues, the loss kernel, the weight split after Adam), so the split costs no extra pass;
//   * the same planes serve both roles a matrix plays in the update.  As a K-major operand (contraction over its columns:
//     forward, input gradient) a K chunk is a 2-D TMA box {32 columns, 128 or N rows} with the 128-byte swizzle, read by
//     tcgen05.mma through a SWIZZLE_128B descriptor.  As an MN-major operand (contraction over its rows: weight gradients, W
//     in the input gradient) a K chunk is a 3-D box {32 columns, 32 rows, M or N / 32 column groups} with the 32-byte-atom
//     128-byte swizzle -- the one layout tcgen05 accepts for MN-major tf32 (SWIZZLE_128B_BASE32B) -- and the instruction
//     descriptor's major bits are set.  Nothing is ever transposed in memory; tile tails are zero-filled by the TMA unit;
//   * roles: warp 0 issues tcgen05.mma (hi*hi + hi*lo + lo*hi), warp 1 issues the TMA loads (mbarrier expect_tx), warps
//     4-7 run the epilogue out of TMEM -- bias / ReLU / ReLU-mask, then either the next GEMM's operand planes (split into
//     hi / lo on the way out), a plain row-major tile (head outputs), or red.global.add into the flat gradient (weight
//     gradients, split over the SMs along the minibatch).
#pragma once

#ifndef B2H_GEMM_TK
#define B2H_GEMM_TK 16
#endif
constexpr int TK = B2H_GEMM_TK;              // K per chunk: 32 (K-major: one 128-byte swizzle row) or 16 (64-byte swizzle rows: half-size stages, twice as
                                             // many -- with two CTAs per SM each CTA then has a two-stage ring of its own: 0.284 -> 0.266 ms per minibatch; the default) or 8
                                             // (32-byte swizzle, four stages: measured slower, 0.298 ms)
static_assert(TK == 32 || TK == 16 || TK == 8, "K chunk");
constexpr int T_EPI_PART = 128 * 32;         // floats of one plane of an epilogue staging buffer: 128 rows x 32 columns (16 KB)
#ifndef B2H_GEMM_CTAS
#define B2H_GEMM_CTAS 2
#endif
// CTAs per SM: 1 = a 192 KB ring (two stages at N = 256), 16 warps; 2 = two CTAs of 96 KB (one stage at N = 256) and 8 warps
// each share an SM and its 512 TMEM columns, so that one's prologue / epilogue runs under the other's main loop and the 256
// tiles of a forward GEMM are one wave (measured: 0.305 -> 0.295 ms per minibatch; the default)
constexpr int T_CTAS_DEFAULT = B2H_GEMM_CTAS;
// per kernel instantiation (template parameter CTAS): stages at the widest N (narrower tiles get more), threads
__host__ __device__ constexpr int t_ns(int ctas) { return (ctas == 1 ? 2 : 1) * (32 / B2H_GEMM_TK); }
__host__ __device__ constexpr int t_threads(int ctas) { return ctas == 1 ? 512 : 256; }
constexpr int TNS_MAX = 8;
constexpr int T_A_PART = 128 * TK;           // floats of one plane of an A chunk (16 KB)
constexpr int T_B_PART = 256 * TK;           // ... of a B chunk at the widest N (32 KB)
constexpr int T_STAGE = 2 * T_A_PART + 2 * T_B_PART;   // A hi | A lo | B hi | B lo: 96 KB

struct alignas(64) TMaps { CUtensorMap m[2][6]; };      // [problem][A hi, A lo, B hi, B lo, C hi, C lo (operand-plane results)]

struct TProblem {
  int a_mn, b_mn;            // operand use: 0 = K-major (contraction over the matrix's columns), 1 = MN-major (over its rows)
  int m_tiles, n_tiles;      // tiles of 128 rows / nw columns
  int nw;                    // UMMA N of a tile (multiple of 16, <= 256)
  int chunks;                // K chunks of 32
  int M, N;                  // valid rows / columns of the result (plain epilogues)
  int epi;                   // 0: result as operand planes (hi / lo), 1: plain row-major store, 2: red.global.add (weight gradient)
  float *c_hi, *c_lo;        // epi 0: result planes
  int c_ld;                  //        their row stride (padded columns)
  float* C;                  // epi 1 / 2
  int ldc, transpose_c;
  const float* bias;         // [N] or null
  int relu;
  const uint32_t* bits_in;   // epi 0: sign bits of the activation that gates the result (ReLU backward): [rows][8] words, or null
  uint32_t* bits_out;        // epi 0: sign bits of this result (x > 0) for the backward pass, or null
  float* colsum;             // epi 0: column sums of the result are added here (bias gradient), or null
};
#ifdef B2H_GEMM_CLK
#define GCLK(i) do { if (threadIdx.x == (i == 4 || i == 5 ? 64 : 0)) a.clk[(size_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) * 8 + (i)] = clock64(); } while (0)
#else
#define GCLK(i) do { } while (0)
#endif
struct TArgs {
  long long* clk;            // B2H_GEMM_CLK builds: eight time stamps per CTA
  TProblem p[2];
  int nsplit, chunks_per_split, precise;
  int cluster;               // 1, or 2: pairs of CTAs along the M tiles share every B chunk (each loads half of it and multicasts)
  int* error;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// shared-memory matrix descriptor with a swizzled layout: layout type at bits 61-63 (2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B)
__device__ __forceinline__ uint64_t umma_desc_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  return umma_desc(saddr, lbo, sbo) | ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
// the same loads delivered to every CTA of the cluster named in `mask` (same shared-memory offset, same mbarrier offset in each)
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;\n"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3, %4}], [%5], %6;\n"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "h"(mask) : "memory");
}
// named barrier of an epilogue part (128 threads); immediate ids, so that the kernel claims only the barriers it uses (two
// CTAs share an SM's sixteen)
template <int NPART>
__device__ __forceinline__ void part_barrier(int part) {
  if (part == 0) asm volatile("bar.sync 1, 128;\n" ::: "memory");
  else if (part == 1) asm volatile("bar.sync 2, 128;\n" ::: "memory");
  else if (NPART > 2 && part == 2) asm volatile("bar.sync 3, 128;\n" ::: "memory");
  else if (NPART > 2) asm volatile("bar.sync 4, 128;\n" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// instruction descriptor kind::tf32 with the operands' major bits (15: A, 16: B; 1 = MN-major)
__device__ __forceinline__ uint32_t umma_idesc_tf32_major(int m, int n, int a_mn, int b_mn) {
  return umma_idesc_tf32(m, n) | ((uint32_t)(a_mn != 0) << 15) | ((uint32_t)(b_mn != 0) << 16);
}

template <int CTAS>
__global__ void __launch_bounds__(t_threads(CTAS), CTAS) gemm_t_kernel(const __grid_constant__ TMaps maps, TArgs a) {
  constexpr int TNS = t_ns(CTAS), T_THREADS = t_threads(CTAS);   // two roles during the main loop, then epilogue parts of four warps each
  constexpr int T_NPART = T_THREADS / 128, T_ROUNDS = 8 / T_NPART;
  extern __shared__ __align__(1024) unsigned char smem_t[];
  float* stage0 = reinterpret_cast<float*>(smem_t);
  __shared__ __align__(8) unsigned long long bar_storage[2 * TNS_MAX + 1];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int prob = blockIdx.z / a.nsplit, split = blockIdx.z - prob * a.nsplit;
  const TProblem P = prob ? a.p[1] : a.p[0];
  GCLK(0);
  if ((int)blockIdx.x >= P.m_tiles || (int)blockIdx.y >= P.n_tiles) return;
  const int row0 = blockIdx.x * 128, col0 = blockIdx.y * P.nw;
  const int c_begin = split * a.chunks_per_split, c_end = min(P.chunks, c_begin + a.chunks_per_split);
  const int nchunk = c_end - c_begin;
  if (nchunk <= 0) return;
  const bool precise = a.precise != 0;
  const int nw = P.nw;
  // a stage holds A hi | A lo | B hi | B lo of one K chunk; its size follows the N tile, the ring always fills the 192 KB
  const int b_part = nw * TK;                                       // floats of one plane of a B chunk (a multiple of 1 KB: nw % 8 == 0)
  const int stage_floats = 2 * T_A_PART + 2 * ((b_part + 255) & ~255);
  const int ns = min(TNS_MAX, (TNS * T_STAGE) / stage_floats);
  const uint32_t full0 = smem_u32(&bar_storage[0]), empty0 = smem_u32(&bar_storage[TNS_MAX]);
  const uint32_t accbar = smem_u32(&bar_storage[2 * TNS_MAX]);
  // Cluster of two CTAs along the M tiles (same B tile): each loads half of every B chunk and multicasts it to both, so a B
  // chunk crosses L2 -> SM once per pair.  A stage is then only free when BOTH CTAs' MMAs have read it: the commits are
  // multicast to the pair's empty barriers (count 2).  Both CTAs of a pair take the same early returns (same y / z, m_tiles even).
  const int csize = a.cluster;
  uint32_t crank = 0;
  if (csize > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(crank));
  const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
  if (threadIdx.x == 0) {
    for (int s = 0; s < ns; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(full0 + 8 * s));    // the TMA thread's arrive.expect_tx + the bytes
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(empty0 + 8 * s), "r"(csize));   // tcgen05.commit of every CTA of the cluster
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(accbar));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (csize > 1) cluster_sync_all();   // the peer's barriers are initialised before anything of ours can reach them
  const uint32_t tmem = tmem_base_s;
  bool ok = true;
  GCLK(1);

  if (warp == 0) {
    if (lane == 0) {   // ---- MMA issuer
      const uint32_t idesc = umma_idesc_tf32_major(128, nw, P.a_mn, P.b_mn);
      // K-major (SWIZZLE_128B): rows of 128 bytes, 8-row swizzle atoms 1024 B apart (SBO), LBO unused (1); a K step of 8 tf32
      // is 32 bytes further along the row.  MN-major (SWIZZLE_128B_BASE32B): per group of 32 columns, 32 K rows of 128 bytes;
      // atoms of 4 K rows 512 B apart (SBO), column groups 4096 B apart (LBO); a K step of 8 = two atoms = 1024 B.
      // (K chunks of 16: the K-major rows are 64 bytes -> SWIZZLE_64B, atoms of 8 rows 512 B apart; MN-major chunks hold 16 rows per column group)
      constexpr uint32_t kSbo = (uint32_t)TK * 32u, kLt = TK == 32 ? 2u : (TK == 16 ? 4u : 6u), mnLbo = (uint32_t)TK * 128u;   // 8 rows of TK floats; SWIZZLE_128B / _64B / _32B
      const uint32_t lboA = P.a_mn ? mnLbo : 16u, sboA = P.a_mn ? 512u : kSbo, stepA = P.a_mn ? 1024u : 32u, ltA = P.a_mn ? 1u : kLt;
      const uint32_t lboB = P.b_mn ? mnLbo : 16u, sboB = P.b_mn ? 512u : kSbo, stepB = P.b_mn ? 1024u : 32u, ltB = P.b_mn ? 1u : kLt;
      int s = 0, use = 0;
      for (int c = 0; c < nchunk && ok; c++) {
        ok = mbar_wait(full0 + 8 * s, use & 1);
        if (!ok) break;
        if (c == 0) GCLK(2);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t A_hi = smem_u32(stage0 + s * stage_floats), A_lo = A_hi + T_A_PART * 4, B_hi = A_lo + T_A_PART * 4,
                       B_lo = B_hi + (uint32_t)(stage_floats - 2 * T_A_PART) * 2;
#pragma unroll
        for (int ks = 0; ks < TK / 8; ks++) {
          const uint64_t ah = umma_desc_sw(A_hi + ks * stepA, lboA, sboA, ltA), bh = umma_desc_sw(B_hi + ks * stepB, lboB, sboB, ltB);
          umma_tf32(tmem, ah, bh, idesc, (c | ks) != 0);
          if (precise) {
            const uint64_t al = umma_desc_sw(A_lo + ks * stepA, lboA, sboA, ltA), bl = umma_desc_sw(B_lo + ks * stepB, lboB, sboB, ltB);
            umma_tf32(tmem, ah, bl, idesc, 1);
            umma_tf32(tmem, al, bh, idesc, 1);
          }
        }
        if (csize > 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
                       ::"r"(empty0 + 8 * s), "h"(cmask) : "memory");
        else
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(empty0 + 8 * s) : "memory");
        if (++s == ns) { s = 0; use++; }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(accbar) : "memory");
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {   // ---- TMA producer
      const CUtensorMap* mp = maps.m[prob ? 1 : 0];
      const uint32_t bytes = (uint32_t)(128 * TK * 4 + nw * TK * 4) * (precise ? 2u : 1u);
      int s = 0, use = 0;
      for (int c = 0; c < nchunk && ok; c++) {
        const int kc = c_begin + c;
        const uint32_t fullb = full0 + 8 * s;
        if (use > 0) ok = mbar_wait(empty0 + 8 * s, (use - 1) & 1);
        if (!ok) break;
        const uint32_t A_hi = smem_u32(stage0 + s * stage_floats), A_lo = A_hi + T_A_PART * 4, B_hi = A_lo + T_A_PART * 4,
                       B_lo = B_hi + (uint32_t)(stage_floats - 2 * T_A_PART) * 2;
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(fullb), "r"(bytes) : "memory");
        // K-major: 2-D box at {first column of the chunk, first row of the tile}; MN-major: 3-D box at {0, first row of the
        // chunk, first column group of the tile}
        if (P.a_mn) { tma_load_3d(A_hi, mp + 0, 0, kc * TK, row0 / 32, fullb); if (precise) tma_load_3d(A_lo, mp + 1, 0, kc * TK, row0 / 32, fullb); }
        else        { tma_load_2d(A_hi, mp + 0, kc * TK, row0, fullb);        if (precise) tma_load_2d(A_lo, mp + 1, kc * TK, row0, fullb); }
        if (csize > 1) {   // this CTA's half of the B chunk, delivered to both CTAs of the pair (the half-box maps are built for it)
          if (P.b_mn) {
            const uint32_t off = crank * (uint32_t)(nw / 64) * (uint32_t)(TK * 128);
            tma_load_3d_mc(B_hi + off, mp + 2, 0, kc * TK, col0 / 32 + (int)crank * (nw / 64), fullb, cmask);
            if (precise) tma_load_3d_mc(B_lo + off, mp + 3, 0, kc * TK, col0 / 32 + (int)crank * (nw / 64), fullb, cmask);
          } else {
            const uint32_t off = crank * (uint32_t)(nw / 2) * (uint32_t)(TK * 4);
            tma_load_2d_mc(B_hi + off, mp + 2, kc * TK, col0 + (int)crank * (nw / 2), fullb, cmask);
            if (precise) tma_load_2d_mc(B_lo + off, mp + 3, kc * TK, col0 + (int)crank * (nw / 2), fullb, cmask);
          }
        } else if (P.b_mn) { tma_load_3d(B_hi, mp + 2, 0, kc * TK, col0 / 32, fullb); if (precise) tma_load_3d(B_lo, mp + 3, 0, kc * TK, col0 / 32, fullb); }
        else               { tma_load_2d(B_hi, mp + 2, kc * TK, col0, fullb);        if (precise) tma_load_2d(B_lo, mp + 3, kc * TK, col0, fullb); }
        if (++s == ns) { s = 0; use++; }
      }
    }
    __syncwarp();
  }
  {   // ---- epilogue on all warps: TMEM lane quadrant = warp % 4 (thread = row of the tile); the four warps with the
      // same warp / 4 form a PART that owns every fourth group of 32 columns and synchronises only with itself
    ok = mbar_wait(accbar, 0) && ok;
    __syncwarp();
    GCLK(4);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int quad = warp & 3, part = warp >> 2;
    const int r = quad * 32 + lane;
    const int grow = row0 + r;
    const bool first = split == 0;
    if (ok && P.epi == 0) {
      // ---- the result is the next GEMM's operand: bias, ReLU / ReLU mask, split into hi / lo, staged in the (now idle)
      // operand ring as 128-byte swizzled rows and written with TMA stores, 32 columns at a time; the column sums of the
      // tile (bias gradient) are taken from the staged copy, the ReLU pattern travels as one bit per element
      // staging buffers of 32 columns x 128 rows x (hi + lo) = 32 KB in the idle operand ring (six fit): parts 0 and 1 use
      // buffers {0, 4} and {1, 5} for their two groups, parts 2 and 3 reuse buffer 2 / 3 once the first store has read it
      static_assert((TNS * T_STAGE) / (2 * T_EPI_PART) >= (T_NPART == 4 ? 6 : 2), "staging buffers");
      const CUtensorMap* cmap = maps.m[prob ? 1 : 0] + 4;
      uint32_t in_bits[8], out_bits[T_ROUNDS];
      if (P.bits_in) {
        const uint4* bp = reinterpret_cast<const uint4*>(P.bits_in + (size_t)grow * 8);
        const uint4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
        in_bits[0] = b0.x; in_bits[1] = b0.y; in_bits[2] = b0.z; in_bits[3] = b0.w; in_bits[4] = b1.x; in_bits[5] = b1.y; in_bits[6] = b1.z; in_bits[7] = b1.w;
      }
#pragma unroll
      for (int gp = 0; gp < T_ROUNDS; gp++) out_bits[gp] = 0u;
#pragma unroll
      for (int gp = 0; gp < T_ROUNDS; gp++) {
        const int g = T_NPART * gp + part;
        const bool active = g * 32 < nw;
        if (!active) break;                                        // uniform over the part
        uint32_t v[32];
#pragma unroll
        for (int hh = 0; hh < 2; hh++)
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
              : "=r"(v[16 * hh + 0]), "=r"(v[16 * hh + 1]), "=r"(v[16 * hh + 2]), "=r"(v[16 * hh + 3]), "=r"(v[16 * hh + 4]), "=r"(v[16 * hh + 5]),
                "=r"(v[16 * hh + 6]), "=r"(v[16 * hh + 7]), "=r"(v[16 * hh + 8]), "=r"(v[16 * hh + 9]), "=r"(v[16 * hh + 10]), "=r"(v[16 * hh + 11]),
                "=r"(v[16 * hh + 12]), "=r"(v[16 * hh + 13]), "=r"(v[16 * hh + 14]), "=r"(v[16 * hh + 15])
              : "r"(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * 32 + 16 * hh)));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        const int colb = col0 + g * 32;
        float x[32];
        uint32_t word = 0u;
        if (P.bias) {                                              // N is a multiple of 32 for operand-plane results
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(P.bias + colb) + q);
            v[4 * q + 0] = __float_as_uint(__uint_as_float(v[4 * q + 0]) + b4.x); v[4 * q + 1] = __float_as_uint(__uint_as_float(v[4 * q + 1]) + b4.y);
            v[4 * q + 2] = __float_as_uint(__uint_as_float(v[4 * q + 2]) + b4.z); v[4 * q + 3] = __float_as_uint(__uint_as_float(v[4 * q + 3]) + b4.w);
          }
        }
        uint32_t gate = 0xFFFFFFFFu;
        if (P.bits_in) {
          if constexpr (T_NPART == 4) {
            const uint32_t lo2 = part & 1 ? in_bits[4 * gp + 1] : in_bits[4 * gp], hi2 = part & 1 ? in_bits[4 * gp + 3] : in_bits[4 * gp + 2];
            gate = part & 2 ? hi2 : lo2;
          } else {
            gate = part & 1 ? in_bits[2 * gp + 1] : in_bits[2 * gp];
          }
        }
#pragma unroll
        for (int q = 0; q < 32; q++) {
          float y = __uint_as_float(v[q]);
          if (P.relu) y = fmaxf(y, 0.f);
          y = (gate >> q) & 1u ? y : 0.f;
          word |= (y > 0.f ? 1u : 0u) << q;
          x[q] = y;
        }
        out_bits[gp] = word;
        // four parts, six buffers: parts 0 / 1 use {0, 4} / {1, 5}, parts 2 / 3 reuse 2 / 3; two parts: one buffer each, reused every round
        const int b = T_NPART == 4 ? (part < 2 ? part + 4 * gp : part) : part;
        float* s_hi = stage0 + b * (2 * T_EPI_PART);
        float* s_lo = s_hi + T_EPI_PART;
        if (T_NPART == 4 ? (gp == 1 && part >= 2) : gp >= 1) {   // the TMA store of this part's previous group must have finished reading the buffer
          if (r == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
          part_barrier<T_NPART>(part);
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
          float4 hi, lo;
          split1(x[4 * q + 0], hi.x, lo.x); split1(x[4 * q + 1], hi.y, lo.y); split1(x[4 * q + 2], hi.z, lo.z); split1(x[4 * q + 3], hi.w, lo.w);
          const int off = r * 32 + ((q ^ (r & 7)) << 2);             // 128-byte swizzle: 16-byte chunk index ^ (row % 8)
          *reinterpret_cast<float4*>(s_hi + off) = hi;
          *reinterpret_cast<float4*>(s_lo + off) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        part_barrier<T_NPART>(part);
        if (r == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];\n"
                       ::"l"(cmap + 0), "r"(colb), "r"(row0), "r"(smem_u32(s_hi)) : "memory");
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];\n"
                       ::"l"(cmap + 1), "r"(colb), "r"(row0), "r"(smem_u32(s_lo)) : "memory");
          asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
        if (P.colsum) {           // thread = (column of the group, quarter of the rows)
          const int col = r & 31, rq = r >> 5;
          float acc = 0.f;
#pragma unroll 8
          for (int i = 0; i < 32; i++) {
            const int rr = rq * 32 + i;
            const int off = rr * 32 + ((((col >> 2) ^ (rr & 7)) << 2) | (col & 3));
            acc += s_hi[off] + s_lo[off];
          }
          if (colb + col < P.N) atomicAdd(P.colsum + colb + col, acc);
        }
      }
      if (P.bits_out) {           // each part owns the words of its groups
        uint32_t* bp = P.bits_out + (size_t)grow * 8;
#pragma unroll
        for (int gp = 0; gp < T_ROUNDS; gp++)
          if ((T_NPART * gp + part) * 32 < nw) bp[T_NPART * gp + part] = out_bits[gp];
      }
      GCLK(5);
      if (r == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // shared memory stays valid until the stores have read it
    } else if (ok) {
      for (int c0 = part * 16; c0 < nw; c0 += 16 * T_NPART) {
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        const int colb = col0 + c0;
        float x[16];
#pragma unroll
        for (int q = 0; q < 16; q++) x[q] = __uint_as_float(v[q]);
          if (grow >= P.M || colb >= P.N) continue;
          if (P.bias && first) {
#pragma unroll
            for (int q = 0; q < 16; q++) x[q] += colb + q < P.N ? __ldg(P.bias + colb + q) : 0.f;
          }
          if (P.transpose_c) {
#pragma unroll
            for (int q = 0; q < 16; q++)
              if (colb + q < P.N) {
                float* dst = P.C + (size_t)(colb + q) * P.ldc + grow;
                if (P.epi == 2) atomicAdd(dst, x[q]); else *dst = x[q];
              }
          } else {
            float* crow = P.C + (size_t)grow * P.ldc + colb;
            if (colb + 16 <= P.N && (P.ldc & 3) == 0 && ((uintptr_t)P.C & 15) == 0) {
#pragma unroll
              for (int q = 0; q < 4; q++) {
                if (P.epi == 2) red_add_v4(crow + 4 * q, x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
                else *reinterpret_cast<float4*>(crow + 4 * q) = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
              }
            } else {
#pragma unroll
              for (int q = 0; q < 16; q++)
                if (colb + q < P.N) { if (P.epi == 2) atomicAdd(crow + q, x[q]); else crow[q] = x[q]; }
            }
          }
      }
    }
  }
  if (!ok) atomicExch(a.error, 1);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  GCLK(6);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tmem) : "memory");
  if (csize > 1) cluster_sync_all();   // nobody leaves while the peer's commits / multicasts may still target its shared memory
}

// ------------------------------------------------------------------------------------------------ producers of operand planes
struct GatherTArgs {
  const float *obs, *actions, *old_logp, *adv, *ret;
  const int64_t* idx;
  long long row_start;
  float *x_hi, *x_lo;        // [rows padded to 128][ld]: the minibatch observations, split
  float *act, *olp, *a, *r;  // plain minibatch vectors
  double* scratch;
  int n_rows, rows_pad, obs_dim, ld, act_dim;
};
// One warp per row.  Rows beyond the minibatch and columns beyond obs_dim are zeros (they enter the weight-gradient sums).
__global__ void __launch_bounds__(256) gather_t_kernel(GatherTArgs g) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= g.rows_pad) return;
  const bool live = row < g.n_rows;
  const long long src = live ? (g.idx ? (long long)g.idx[row] : g.row_start + row) : 0;
  const float* xs = g.obs + (size_t)src * g.obs_dim;
  const bool vec = (g.obs_dim & 3) == 0;
  for (int col = lane * 4; col < g.ld; col += 128) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      if (vec && col + 3 < g.obs_dim) v = __ldg(reinterpret_cast<const float4*>(xs + col));
      else {
        if (col + 0 < g.obs_dim) v.x = __ldg(xs + col + 0);
        if (col + 1 < g.obs_dim) v.y = __ldg(xs + col + 1);
        if (col + 2 < g.obs_dim) v.z = __ldg(xs + col + 2);
        if (col + 3 < g.obs_dim) v.w = __ldg(xs + col + 3);
      }
    }
    float4 hi, lo;
    split1(v.x, hi.x, lo.x); split1(v.y, hi.y, lo.y); split1(v.z, hi.z, lo.z); split1(v.w, hi.w, lo.w);
    *reinterpret_cast<float4*>(g.x_hi + (size_t)row * g.ld + col) = hi;
    *reinterpret_cast<float4*>(g.x_lo + (size_t)row * g.ld + col) = lo;
  }
  if (live) {
    for (int i = lane; i < g.act_dim; i += 32) g.act[(size_t)row * g.act_dim + i] = __ldg(g.actions + (size_t)src * g.act_dim + i);
    if (lane == 0) { g.olp[row] = __ldg(g.old_logp + src); g.a[row] = __ldg(g.adv + src); g.r[row] = __ldg(g.ret + src); }
  }
}

// Weights W [R, F] (nn.Linear layout, inside the flat parameter vector) -> split planes [rows_pad][ld], zero padded
struct PackTJob { const float* W; float *hi, *lo; int R, F, rows_pad, ld; };
struct PackTJobs { PackTJob j[6]; };
__global__ void __launch_bounds__(256) pack_t_kernel(PackTJobs jobs) {
  const PackTJob jb = jobs.j[blockIdx.y];
  const int ld4 = jb.ld / 4, total = jb.rows_pad * ld4;   // float4 items
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int row = i / ld4, col = (i - row * ld4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < jb.R) {
      const float* src = jb.W + (size_t)row * jb.F + col;
      if (col + 0 < jb.F) v.x = src[0];
      if (col + 1 < jb.F) v.y = src[1];
      if (col + 2 < jb.F) v.z = src[2];
      if (col + 3 < jb.F) v.w = src[3];
    }
    float4 hi, lo;
    split1(v.x, hi.x, lo.x); split1(v.y, hi.y, lo.y); split1(v.z, hi.z, lo.z); split1(v.w, hi.w, lo.w);
    *reinterpret_cast<float4*>(jb.hi + (size_t)row * jb.ld + col) = hi;
    *reinterpret_cast<float4*>(jb.lo + (size_t)row * jb.ld + col) = lo;
  }
}
