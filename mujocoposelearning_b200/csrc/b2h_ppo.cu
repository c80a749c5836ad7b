// b2h_ppo.cu — the PPO update on the device with hand-written kernels, sm_100a.
//
// What it replaces: SB3 2.3.2 PPO.train as the reference drives it (train_sb3.py:208-231, kwargs config.py:17-32): per
// minibatch the forward of the two MlpPolicy trunks (obs -> 256 -> 256 -> 21 / 1, ReLU: main.py:99-105), the clipped
// surrogate + value MSE (+ entropy bonus) with per-minibatch advantage normalisation, the backward pass, grad-norm clipping
// and Adam.  The reference runs this through PyTorch autograd on the CPU; round 1-2 here ran it through autograd + library
// GEMMs.  Now every step of a minibatch is a kernel of this file or of b2h_ppo_tma.cuh.
//
// The path the update runs on (b2h_ppo_tma.cuh):
//   pack_t_kernel        the weights, split into tf32 hi / lo planes
//   gather_t_kernel      minibatch rows of the rollout buffer -> split observation planes + plain vectors (b2h_ppo_train gathers
//                        the next minibatch on a side stream while the previous apply runs); adv_moments_kernel: their advantage moments
//   gemm_t_kernel        C = A . B^T on the tcgen05 tensor cores, fp32-faithful (hi*hi + hi*lo + lo*hi into TMEM), every operand
//                        streamed by TMA from planes its producer kernel already split; K-major (SWIZZLE_64B boxes) or MN-major
//                        (32-byte-atom swizzled 3-D boxes) use of the same planes gives the forward, input-gradient and
//                        weight-gradient products without a transpose; two CTAs per SM share the TMEM; results that feed later GEMMs
//                        leave through TMA stores (+ bias-gradient column sums, ReLU sign bits), weight gradients through red.global.add
//   ppo_loss_kernel      log-probability, ratio, clipped surrogate, value loss, their gradients with respect to the action mean /
//                        value / log_std, head bias gradients, loss statistics (eight lanes per sample)
//   apply_kernel         one launch: [sum of the ranks' gradients by peer loads over NVLink] -> sum of squares -> grid barrier ->
//                        grad-norm clip (max_grad_norm) + torch.optim.Adam's update rule on the flat parameter vector (or the
//                        caller all-reduces the flat gradient itself between b2h_ppo_minibatch_grad and b2h_ppo_apply)
//
// The first version of the GEMM stays in this file for A/B measurements, hidden widths that are not a multiple of 32 and the
// generic b2h_gemm entry (staged_operands = 1):
//   gather_kernel        minibatch rows -> contiguous row-major operands
//   gemm_kernel          the same three products with plain row-major fp32 operands: seven producer warps read either [rows, K]
//                        (K contiguous) or [K, rows] (K strided) memory, split into hi / lo and stage the canonical no-swizzle
//                        K-major core-matrix layout; split-K partial tiles added with red.global.add
//   colsum_kernel        hidden-layer bias gradients
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>

#include "../../include/b2h.h"

namespace {

thread_local std::string g_err_ppo;

// ------------------------------------------------------------------------------------------------------------------ GEMM
constexpr int GM = 128;          // rows of C per CTA (UMMA M)
constexpr int GN = 256;          // columns of C per CTA (UMMA N <= 256, multiple of 16)
constexpr int GK = 32;           // K columns per chunk (four MMA K-steps)
constexpr int GNS = 2;           // operand stages
constexpr int GTHREADS = 256;
constexpr int GPROD = GTHREADS - 32;                            // producer threads: warps 1..7
constexpr int GA_PART = GM * GK, GB_PART = GN * GK;             // floats of the hi (or lo) part of a chunk
constexpr int GSTAGE = 2 * GA_PART + 2 * GB_PART;               // floats: A hi | A lo | B hi | B lo  (96 KB)
constexpr int GITEMS = ((GM + GN) * (GK / 4) + GPROD - 1) / GPROD;   // 16-byte items per producer per chunk (14)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared-memory matrix descriptor, no swizzle, K-major: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}
// instruction descriptor kind::tf32: D fp32 (bit 4), A / B tf32 (2 at bits 7 and 10), both K-major, N >> 3 at 17, M >> 4 at 24
__device__ __forceinline__ uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint32_t mbar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 22); spin++) {   // bounded: a lost arrive must not hang the GPU
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

struct GemmProblem {
  const float *A, *B, *bias, *mask;
  float* C;
  int lda, ldb, ldc, ldmask;
  int M, N, K;
  int a_kstrided, b_kstrided;   // 0: the operand is [rows, K] row-major (K contiguous); 1: [K, rows] row-major
  int relu;                     // epilogue: max(x, 0) (after the bias)
  int atomic;                   // epilogue: C += x (split-K partial sums, accumulation into an existing C)
  int transpose_c;              // epilogue: the tile is written as C[col * ldc + row]
};
struct GemmArgs {
  GemmProblem p[2];
  int nsplit, k_per_split, precise;
  int* error;
};

struct Chunk { float4 v[GITEMS]; };

// Item g of an operand with `rows` rows -> (row r, 16-byte K group k4 of the chunk).
//   K contiguous: eight consecutive threads take eight consecutive rows of one K group (one 128-byte core matrix per
//   quarter-warp on the shared side, 64 contiguous bytes of each of 8 rows per warp on the global side).
//   K strided: consecutive threads take consecutive rows of one K group (the four K values of an item are four loads, each
//   coalesced over the warp; the shared side is 32 consecutive 16-byte units).
__device__ __forceinline__ void item_kc(int g, int& r, int& k4) { r = (g & 7) | ((g >> 6) << 3); k4 = (g >> 3) & 7; }
// g / rows by a multiply: magic = ceil(2^32 / rows) is exact for g < 2^32 / rows (here g < 2048, rows <= 256)
__device__ __forceinline__ void item_ks(int g, int rows, uint32_t magic, int& r, int& k4) { k4 = (int)__umulhi((uint32_t)g, magic); r = g - k4 * rows; }

__device__ __forceinline__ float4 load_item(const float* __restrict__ base, int ld, bool kstrided, bool vec, int row, int row_lim, int k,
                                            int kend) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= row_lim || k >= kend) return v;
  if (!kstrided) {
    const float* src = base + (size_t)row * ld + k;
    if (vec && k + 3 < kend) return __ldg(reinterpret_cast<const float4*>(src));
    v.x = __ldg(src);
    if (k + 1 < kend) v.y = __ldg(src + 1);
    if (k + 2 < kend) v.z = __ldg(src + 2);
    if (k + 3 < kend) v.w = __ldg(src + 3);
  } else {
    const float* src = base + (size_t)k * ld + row;
    v.x = __ldg(src);
    if (k + 1 < kend) v.y = __ldg(src + ld);
    if (k + 2 < kend) v.z = __ldg(src + 2 * (size_t)ld);
    if (k + 3 < kend) v.w = __ldg(src + 3 * (size_t)ld);
  }
  return v;
}

__device__ __forceinline__ void load_chunk(Chunk& c, int pt, const GemmProblem& P, int row0, int col0, int nw, uint32_t magic, int k0, int kend) {
  constexpr int NA = GM * (GK / 4);
  const int nitems = NA + nw * (GK / 4);
  const bool veca = (P.lda & 3) == 0 && ((uintptr_t)P.A & 15) == 0, vecb = (P.ldb & 3) == 0 && ((uintptr_t)P.B & 15) == 0;
#pragma unroll
  for (int i = 0; i < GITEMS; i++) {
    const int f = pt + i * GPROD;
    int r, k4;
    if (f < NA) {
      if (P.a_kstrided) { k4 = f >> 7; r = f & (GM - 1); } else item_kc(f, r, k4);
      c.v[i] = load_item(P.A, P.lda, P.a_kstrided != 0, veca, row0 + r, P.M, k0 + 4 * k4, kend);
    } else if (f < nitems) {
      if (P.b_kstrided) item_ks(f - NA, nw, magic, r, k4); else item_kc(f - NA, r, k4);
      c.v[i] = load_item(P.B, P.ldb, P.b_kstrided != 0, vecb, col0 + r, P.N, k0 + 4 * k4, kend);
    } else {
      c.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}
// tf32 hi = the value rounded to 10 mantissa bits (exact in the tensor core's operand format), lo = the exact remainder
__device__ __forceinline__ void split1(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}
__device__ __forceinline__ void store_chunk(const Chunk& c, int pt, const GemmProblem& P, float* A_hi, float* B_hi, int nw, uint32_t magic, bool precise) {
  constexpr int NA = GM * (GK / 4);
  const int nitems = NA + nw * (GK / 4);
#pragma unroll
  for (int i = 0; i < GITEMS; i++) {
    const int f = pt + i * GPROD;
    if (f >= nitems) continue;
    const bool isA = f < NA;
    int r, k4;
    if (isA) { if (P.a_kstrided) { k4 = f >> 7; r = f & (GM - 1); } else item_kc(f, r, k4); }
    else     { if (P.b_kstrided) item_ks(f - NA, nw, magic, r, k4); else item_kc(f - NA, r, k4); }
    const int groups = (isA ? GM : nw) >> 3;
    const float4 v = c.v[i];
    float4 hi, lo;
    split1(v.x, hi.x, lo.x); split1(v.y, hi.y, lo.y); split1(v.z, hi.z, lo.z); split1(v.w, hi.w, lo.w);
    const int off = ((k4 * groups + (r >> 3)) * 32) + (r & 7) * 4;   // floats; a core matrix is 8 rows x 16 bytes
    float* dst = isA ? A_hi : B_hi;
    *reinterpret_cast<float4*>(dst + off) = hi;
    if (precise) *reinterpret_cast<float4*>(dst + (isA ? GA_PART : GB_PART) + off) = lo;
  }
}

__global__ void __launch_bounds__(GTHREADS, 1) gemm_kernel(GemmArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* stage0 = reinterpret_cast<float*>(smem);
  __shared__ __align__(8) unsigned long long bar_storage[2 * GNS + 1];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int prob = blockIdx.z / a.nsplit, split = blockIdx.z - prob * a.nsplit;
  const GemmProblem P = prob ? a.p[1] : a.p[0];
  const int row0 = blockIdx.x * GM, col0 = blockIdx.y * GN;
  if (row0 >= P.M || col0 >= P.N) return;                       // the two problems of a launch share the larger tile grid
  const int nw = min(GN, ((P.N - col0) + 15) & ~15);            // UMMA N of this tile (rows of B beyond N are zero-filled)
  const int k_begin = split * a.k_per_split, k_end = min(P.K, k_begin + a.k_per_split);
  const int nchunk = (k_end - k_begin + GK - 1) / GK;
  if (nchunk <= 0) return;
  const bool precise = a.precise != 0;
  uint32_t full[GNS], empty[GNS];
#pragma unroll
  for (int s = 0; s < GNS; s++) { full[s] = smem_u32(&bar_storage[s]); empty[s] = smem_u32(&bar_storage[GNS + s]); }
  const uint32_t accbar = smem_u32(&bar_storage[2 * GNS]);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < GNS; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(full[s]), "r"(GPROD));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(empty[s]));
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
  const uint32_t tmem = tmem_base_s;
  bool ok = true;

  if (warp == 0) {
    if (lane == 0) {   // ---- MMA issuer
      const uint32_t idesc = umma_idesc_tf32(GM, nw);
      const uint32_t lboA = (GM / 8) * 128, lboB = (uint32_t)(nw / 8) * 128;
      for (int c = 0; c < nchunk && ok; c++) {
        const int s = c % GNS, use = c / GNS;
        ok = mbar_wait(full[s], use & 1);
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t A_hi = smem_u32(stage0 + s * GSTAGE), A_lo = A_hi + GA_PART * 4, B_hi = A_lo + GA_PART * 4, B_lo = B_hi + GB_PART * 4;
#pragma unroll
        for (int ks = 0; ks < GK / 8; ks++) {
          const uint64_t ah = umma_desc(A_hi + ks * 2 * lboA, lboA, 128), bh = umma_desc(B_hi + ks * 2 * lboB, lboB, 128);
          umma_tf32(tmem, ah, bh, idesc, (c | ks) != 0);
          if (precise) {
            const uint64_t al = umma_desc(A_lo + ks * 2 * lboA, lboA, 128), bl = umma_desc(B_lo + ks * 2 * lboB, lboB, 128);
            umma_tf32(tmem, ah, bl, idesc, 1);
            umma_tf32(tmem, al, bh, idesc, 1);
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(empty[s]) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(accbar) : "memory");
    }
    __syncwarp();
  } else {               // ---- producers: both operands, global -> registers -> hi / lo -> core-matrix layout
    const int pt = threadIdx.x - 32;
    const uint32_t magic = (uint32_t)((0x100000000ULL + (uint32_t)nw - 1) / (uint32_t)nw);
    Chunk cur, nxt;
    load_chunk(cur, pt, P, row0, col0, nw, magic, k_begin, k_end);
    for (int c = 0; c < nchunk && ok; c++) {
      if (c + 1 < nchunk) load_chunk(nxt, pt, P, row0, col0, nw, magic, k_begin + (c + 1) * GK, k_end);
      const int s = c % GNS, use = c / GNS;
      if (use > 0) ok = mbar_wait(empty[s], (use - 1) & 1);
      if (!ok) break;
      float* A_hi = stage0 + s * GSTAGE;
      store_chunk(cur, pt, P, A_hi, A_hi + 2 * GA_PART, nw, magic, precise);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      mbar_arrive(full[s]);
      cur = nxt;
    }
  }
  // ---- epilogue (warps 4-7: TMEM lane quadrant = warp % 4, thread = row of the tile)
  if (warp >= 4 && ok) {
    ok = mbar_wait(accbar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int quad = warp & 3;
    const int grow = row0 + (threadIdx.x - 128);
    const bool rowok = grow < P.M;
    const bool first = split == 0;
    if (ok) {
      for (int c0 = 0; c0 < nw; c0 += 16) {
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        const int colb = col0 + c0;
        if (!rowok || colb >= P.N) continue;
        float x[16];
#pragma unroll
        for (int q = 0; q < 16; q++) x[q] = __uint_as_float(v[q]);
        const bool full16 = colb + 16 <= P.N;
        if (P.bias && first) {
#pragma unroll
          for (int q = 0; q < 16; q++) if (full16 || colb + q < P.N) x[q] += __ldg(P.bias + colb + q);
        }
        if (P.relu) {
#pragma unroll
          for (int q = 0; q < 16; q++) x[q] = fmaxf(x[q], 0.f);
        }
        if (P.mask) {
          const float* mrow = P.mask + (size_t)grow * P.ldmask + colb;
          if (full16 && (P.ldmask & 3) == 0 && ((uintptr_t)P.mask & 15) == 0) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const float4 m = __ldg(reinterpret_cast<const float4*>(mrow) + q);
              x[4 * q + 0] = m.x > 0.f ? x[4 * q + 0] : 0.f; x[4 * q + 1] = m.y > 0.f ? x[4 * q + 1] : 0.f;
              x[4 * q + 2] = m.z > 0.f ? x[4 * q + 2] : 0.f; x[4 * q + 3] = m.w > 0.f ? x[4 * q + 3] : 0.f;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 16; q++) if (colb + q < P.N) x[q] = __ldg(mrow + q) > 0.f ? x[q] : 0.f;
          }
        }
        if (P.transpose_c) {
#pragma unroll
          for (int q = 0; q < 16; q++)
            if (colb + q < P.N) {
              float* dst = P.C + (size_t)(colb + q) * P.ldc + grow;
              if (P.atomic) atomicAdd(dst, x[q]); else *dst = x[q];
            }
        } else {
          float* crow = P.C + (size_t)grow * P.ldc + colb;
          const bool vec = full16 && (P.ldc & 3) == 0 && ((uintptr_t)P.C & 15) == 0;
          if (vec) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
              if (P.atomic) red_add_v4(crow + 4 * q, x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
              else *reinterpret_cast<float4*>(crow + 4 * q) = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 16; q++)
              if (colb + q < P.N) { if (P.atomic) atomicAdd(crow + q, x[q]); else crow[q] = x[q]; }
          }
        }
      }
    }
  }
  if (!ok) atomicExch(a.error, 1);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tmem) : "memory");
}

int g_sm_count = 0;

// Launch one or two problems of the same tile-grid class.  split_k <= 0: chosen so that the launch fills the SMs once.
int launch_gemm(const GemmProblem* probs, int nprob, int precise, int split_k, int* error_dev, cudaStream_t stream) {
  static bool attr_set = false;
  const size_t smem = (size_t)GNS * GSTAGE * sizeof(float);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { g_err_ppo = cudaGetErrorString(e); return B2H_ECUDA; }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
    attr_set = true;
  }
  GemmArgs a;
  int mt = 0, nt = 0, kmax = 0;
  for (int i = 0; i < nprob; i++) {
    const GemmProblem& p = probs[i];
    if (!p.A || !p.B || !p.C || p.M <= 0 || p.N <= 0 || p.K <= 0) { g_err_ppo = "gemm: bad argument"; return B2H_EINVAL; }
    a.p[i] = p;
    mt = std::max(mt, (p.M + GM - 1) / GM);
    nt = std::max(nt, (p.N + GN - 1) / GN);
    kmax = std::max(kmax, p.K);
  }
  if (nprob == 1) a.p[1] = a.p[0];
  const int chunks = (kmax + GK - 1) / GK;
  if (split_k <= 0) split_k = std::max(1, std::min(chunks, g_sm_count / std::max(1, mt * nt * nprob)));
  split_k = std::min(split_k, chunks);
  for (int i = 0; i < nprob; i++)
    if (split_k > 1 && (!probs[i].atomic || probs[i].relu || probs[i].mask)) { g_err_ppo = "gemm: split-K needs an accumulating linear epilogue"; return B2H_EINVAL; }
  a.k_per_split = ((chunks + split_k - 1) / split_k) * GK;
  a.nsplit = (kmax + a.k_per_split - 1) / a.k_per_split;
  a.precise = precise;
  a.error = error_dev;
  dim3 grid(mt, nt, nprob * a.nsplit);
  gemm_kernel<<<grid, GTHREADS, smem, stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_err_ppo = cudaGetErrorString(e); return B2H_ECUDA; }
  return B2H_OK;
}

// --------------------------------------------------------------------------------------------------------- small kernels
struct GatherArgs {
  const float *obs, *actions, *old_logp, *adv, *ret;   // flattened rollout buffer [n, ...]
  const int64_t* idx;                                   // minibatch rows (null: rows row_start ... row_start + n_rows)
  long long row_start;
  float *X, *act, *olp, *a, *r;                         // contiguous minibatch
  double* scratch;                                      // zeroed here: [0..3] loss statistics, [4] gradient sum of squares
  int n_rows, obs_dim, act_dim;
};
__global__ void __launch_bounds__(256) gather_kernel(GatherArgs g) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x < 8) g.scratch[threadIdx.x] = 0.0;
  if (warp >= g.n_rows) return;
  const long long src = g.idx ? (long long)g.idx[warp] : g.row_start + warp;
  const float* xs = g.obs + (size_t)src * g.obs_dim;
  float* xd = g.X + (size_t)warp * g.obs_dim;
  if ((g.obs_dim & 3) == 0) {
    for (int i = lane; i < g.obs_dim / 4; i += 32) reinterpret_cast<float4*>(xd)[i] = __ldg(reinterpret_cast<const float4*>(xs) + i);
  } else {
    for (int i = lane; i < g.obs_dim; i += 32) xd[i] = __ldg(xs + i);
  }
  for (int i = lane; i < g.act_dim; i += 32) g.act[(size_t)warp * g.act_dim + i] = __ldg(g.actions + (size_t)src * g.act_dim + i);
  if (lane == 0) { g.olp[warp] = __ldg(g.old_logp + src); g.a[warp] = __ldg(g.adv + src); g.r[warp] = __ldg(g.ret + src); }
}

constexpr int OUT_LD = 32;      // row stride of the head outputs / their gradients (act_dim <= 32, zero padded)
constexpr int MAX_ACT = 32;

struct LossArgs {
  const float *mean, *value;          // [n, OUT_LD] head outputs of the policy / value network (value: column 0)
  const float *act, *olp, *adv, *ret, *log_std;
  float *dmean, *dvalue;              // [n, OUT_LD] gradients of the loss with respect to the head outputs (row-major; staged GEMM)
  float *dt_hi[2], *dt_lo[2];         // the same as split planes [rows padded to 128][32] for the TMA GEMM (dmean null then)
  int rows_pad;
  float *g_b3_pi, *g_b3_vf, *g_log_std;
  double* stats;                      // [0] policy loss, [1] value loss, [2] clip fraction (sums over the minibatch / n), [3] approx kl
  const double* moments;              // advantage mean and unbiased std of the minibatch (adv_moments_kernel)
  int n, act_dim;
  float clip, ent_coef, vf_coef;
  int normalize;
};

__device__ __forceinline__ double warp_sum_d(double v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Advantage moments of the minibatch, once: mean and torch.std (unbiased) in double -> moments[0], moments[1].  One CTA; it
// runs on the side stream while the forward GEMMs run (the loss kernel is the first to need it).
__global__ void __launch_bounds__(1024) adv_moments_kernel(const float* __restrict__ adv, int n, double* moments) {
  __shared__ double red[64];
  double s = 0.0, q = 0.0;
  const int n4 = (((uintptr_t)adv & 15) == 0) ? n / 4 : 0;
  for (int i = threadIdx.x; i < n4; i += 1024) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(adv) + i);
    s += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
    q += ((double)v.x * v.x + (double)v.y * v.y) + ((double)v.z * v.z + (double)v.w * v.w);
  }
  for (int i = 4 * n4 + threadIdx.x; i < n; i += 1024) { const double v = (double)__ldg(adv + i); s += v; q += v * v; }
  s = warp_sum_d(s); q = warp_sum_d(q);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = s; red[32 + (threadIdx.x >> 5)] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int w = 0; w < 32; w++) { ts += red[w]; tq += red[32 + w]; }
    const double m = ts / n;
    moments[0] = m;
    moments[1] = n > 1 ? sqrt(fmax(tq - ts * m, 0.0) / (n - 1)) : 1.0;
  }
}

// Eight lanes per sample, four action dimensions per lane (one float4 of the head output): coalesced reads of the means and
// writes of the gradient planes, three shuffles for the log-probability, nine values per thread to reduce for the head-bias /
// log_std gradients.  Lane 0 of a sample carries the value-network terms and the statistics.
constexpr int LOSS_NT = 256, LOSS_ROWS = LOSS_NT / 8;
__global__ void __launch_bounds__(LOSS_NT) ppo_loss_kernel(LossArgs L) {
  __shared__ float fred[LOSS_NT / 32][8][9];
  __shared__ double dred[LOSS_NT / 32][4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, l8 = tid & 7;
  const int i = blockIdx.x * LOSS_ROWS + (tid >> 3);
  const bool live = i < L.n;
  const int j0 = 4 * l8;
  float ls[4], istd[4], z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 4; q++) { ls[q] = j0 + q < L.act_dim ? __ldg(L.log_std + j0 + q) : 0.f; istd[q] = __expf(-ls[q]); }
  float part = 0.f;
  if (live) {
    const float4 m4 = *reinterpret_cast<const float4*>(L.mean + (size_t)i * OUT_LD + j0);
    const float mu[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (j0 + q < L.act_dim) {
        z[q] = (L.act[(size_t)i * L.act_dim + j0 + q] - mu[q]) * istd[q];
        part += -0.5f * z[q] * z[q] - ls[q] - 0.9189385332046727f;
      }
  }
  part += __shfl_xor_sync(0xffffffffu, part, 1); part += __shfl_xor_sync(0xffffffffu, part, 2); part += __shfl_xor_sync(0xffffffffu, part, 4);
  const float logp = part;
  float pl = 0.f, vl = 0.f, cf = 0.f, kl = 0.f, dv = 0.f, glp = 0.f;
  if (live) {
    const float inv_n = 1.f / (float)L.n;
    float a = L.adv[i];
    if (L.normalize && L.n > 1) a = (float)((double)a - L.moments[0]) / ((float)L.moments[1] + 1e-8f);   // (adv - mean) / (std + 1e-8)
    const float lr = logp - L.olp[i];
    const float ratio = expf(lr);
    const float s1 = a * ratio, s2 = a * fminf(fmaxf(ratio, 1.f - L.clip), 1.f + L.clip);
    const bool inside = ratio >= 1.f - L.clip && ratio <= 1.f + L.clip;
    // d(-min(s1, s2)) / d logp: s1 carries a * ratio; s2 carries it only inside the clip range (ties split evenly, same total)
    glp = (inside || s1 < s2) ? -a * ratio * inv_n : 0.f;
    if (l8 == 0) {
      pl = -fminf(s1, s2);
      cf = fabsf(ratio - 1.f) > L.clip ? 1.f : 0.f;
      kl = (ratio - 1.f) - lr;
      const float diff = L.value[(size_t)i * OUT_LD] - L.ret[i];
      vl = diff * diff;
      dv = L.vf_coef * 2.f * diff * inv_n;
    }
  }
  float dm[4], dls[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const bool on = live && j0 + q < L.act_dim;
    dm[q] = on ? glp * z[q] * istd[q] : 0.f;
    dls[q] = on ? glp * (z[q] * z[q] - 1.f) : 0.f;
  }
  // gradients with respect to the head outputs: rows beyond the minibatch (up to the tile boundary) are zeros
  if (L.dmean) {
    if (live) {
      *reinterpret_cast<float4*>(L.dmean + (size_t)i * OUT_LD + j0) = make_float4(dm[0], dm[1], dm[2], dm[3]);
      *reinterpret_cast<float4*>(L.dvalue + (size_t)i * OUT_LD + j0) = make_float4(l8 == 0 ? dv : 0.f, 0.f, 0.f, 0.f);
    }
  } else if (i < L.rows_pad) {
    const size_t off = (size_t)i * OUT_LD + j0;
    float4 hi, lo;
    split1(dm[0], hi.x, lo.x); split1(dm[1], hi.y, lo.y); split1(dm[2], hi.z, lo.z); split1(dm[3], hi.w, lo.w);
    *reinterpret_cast<float4*>(L.dt_hi[0] + off) = hi;
    *reinterpret_cast<float4*>(L.dt_lo[0] + off) = lo;
    float4 vh = make_float4(0.f, 0.f, 0.f, 0.f), vlo = vh;
    if (l8 == 0) split1(dv, vh.x, vlo.x);
    *reinterpret_cast<float4*>(L.dt_hi[1] + off) = vh;
    *reinterpret_cast<float4*>(L.dt_lo[1] + off) = vlo;
  }
  // ---- reductions.  Column sums: the four samples of a warp that share a lane position (xor 8, 16), then over the warps
  float red9[9] = {dm[0], dm[1], dm[2], dm[3], dls[0], dls[1], dls[2], dls[3], dv};
#pragma unroll
  for (int k = 0; k < 9; k++) { red9[k] += __shfl_xor_sync(0xffffffffu, red9[k], 8); red9[k] += __shfl_xor_sync(0xffffffffu, red9[k], 16); }
  if (lane < 8) {
#pragma unroll
    for (int k = 0; k < 9; k++) fred[warp][lane][k] = red9[k];
  }
  const double st[4] = {warp_sum_d((double)pl), warp_sum_d((double)vl), warp_sum_d((double)cf), warp_sum_d((double)kl)};
  if (lane == 0) { dred[warp][0] = st[0]; dred[warp][1] = st[1]; dred[warp][2] = st[2]; dred[warp][3] = st[3]; }
  __syncthreads();
  if (tid < 72) {               // (lane position, value): 8 x 9
    const int lp = tid / 9, k = tid - 9 * lp;
    float t = 0.f;
    for (int w = 0; w < LOSS_NT / 32; w++) t += fred[w][lp][k];
    if (k < 4) { if (4 * lp + k < L.act_dim) atomicAdd(L.g_b3_pi + 4 * lp + k, t); }
    else if (k < 8) { if (4 * lp + k - 4 < L.act_dim) atomicAdd(L.g_log_std + 4 * lp + k - 4, t - (blockIdx.x == 0 ? L.ent_coef : 0.f)); }   // entropy = sum(log_std) + const
    else if (lp == 0) atomicAdd(L.g_b3_vf, t);
  } else if (tid >= 96 && tid < 100) {
    double t = 0.0;
    for (int w = 0; w < LOSS_NT / 32; w++) t += dred[w][tid - 96];
    atomicAdd(L.stats + (tid - 96), t / L.n);
  }
}

struct ColsumArgs { const float* src[4]; float* dst[4]; int n, width, rows_per_cta; };
__global__ void __launch_bounds__(256) colsum_kernel(ColsumArgs c) {
  const float* __restrict__ s = c.src[blockIdx.y];
  const int r0 = blockIdx.x * c.rows_per_cta, r1 = min(c.n, r0 + c.rows_per_cta);
  for (int col = threadIdx.x; col < c.width; col += 256) {
    float acc0 = 0.f, acc1 = 0.f;
    int r = r0;
    for (; r + 1 < r1; r += 2) { acc0 += __ldg(s + (size_t)r * c.width + col); acc1 += __ldg(s + (size_t)(r + 1) * c.width + col); }
    if (r < r1) acc0 += __ldg(s + (size_t)r * c.width + col);
    atomicAdd(c.dst[blockIdx.y] + col, acc0 + acc1);
  }
}

// ---- gradient all-reduce over NVLink peer memory, fused with the sum of squares of the reduced gradient --------------------
// Every rank keeps its flat gradient in a buffer the other ranks of the node have mapped (CUDA IPC).  One kernel per minibatch:
//   1. tell every peer "my gradient of epoch e is complete" (st.release.sys into the peer's flag row), wait until all peers
//      have said so (ld.acquire.sys on the local flag row);
//   2. each thread sums its float4 slice over the ranks IN RANK ORDER (peer loads through NVLink / NVSwitch), so every rank
//      computes bit-identical sums: the replicas cannot drift;
//   3. the reduced gradient goes to a local buffer and its sum of squares (the grad-norm clip needs it) to the accumulator.
// The gradient buffers alternate between two copies (epoch parity): a rank overwrites copy e % 2 again at epoch e + 2, i.e. after
// it has passed the barrier of epoch e + 1, which every peer only signals after its reduction of epoch e has finished -- no
// second barrier is needed.  Waits are bounded; on a timeout the error flag is raised and the kernel carries on.
constexpr int P2P_MAX_RANKS = 16;
struct ApplyArgs {
  // clip + Adam
  float *p, *g, *m, *v;                 // parameters, gradient (world > 1: receives the sum over the ranks), Adam moments
  long long n;
  double* sumsq;                        // accumulator (zeroed by the caller): sum of squares of g (before grad_scale)
  double* norm_out;                     // receives the gradient norm (after grad_scale, before clipping)
  unsigned* arrived;                    // grid barrier counter (zeroed by the caller)
  int* error;
  float grad_scale;                     // 1 / world size when g holds a sum over ranks
  float max_norm, lr, beta1, beta2, eps;
  float bc1, bc2_sqrt;                  // 1 - beta1^t, sqrt(1 - beta2^t)
  // cross-rank sum (world > 1)
  const float* grad[P2P_MAX_RANKS];     // this epoch's gradient copy on every rank (index = rank; own entry is local memory)
  uint32_t* flags[P2P_MAX_RANKS];       // flag rows [P2P_MAX_RANKS] on every rank
  int rank, world;
  uint32_t epoch;
};
// One launch: [cross-rank gradient sum] -> sum of squares -> grid barrier -> torch.nn.utils.clip_grad_norm_ (coef = max_norm /
// (norm + 1e-6), clamped to 1) + torch.optim.Adam's update rule.  The grid is one CTA per SM (all co-resident), every thread
// owns the same slice in both phases.
__global__ void __launch_bounds__(256) apply_kernel(ApplyArgs a) {
  __shared__ double red[8];
  __shared__ float s_coef;
  const long long n4 = a.n / 4;         // the flat vectors are padded to a multiple of 4 floats
  if (a.world > 1) {
    if (threadIdx.x < a.world) {
      const int j = threadIdx.x;
      if (blockIdx.x == 0) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(a.flags[j] + a.rank), "r"(a.epoch) : "memory");
      }
      const uint32_t* mine = a.flags[a.rank] + j;
      bool seen = false;
      for (int spin = 0; spin < (1 << 22) && !seen; spin++) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(mine) : "memory");
        seen = (int32_t)(v - a.epoch) >= 0;
        if (!seen) __nanosleep(64);
      }
      if (!seen) atomicExch(a.error, 2);
    }
    __syncthreads();
  }
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 acc;
    if (a.world > 1) {
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j0 = 0; j0 < a.world; j0 += 8) {   // eight peer loads in flight (one asm block: ptxas must not interleave the sums), then the sum in rank order
        float4 w[8];
        // ld.volatile: the peers' buffers change between launches, nothing may come from a non-coherent cache; entries beyond
        // `world` alias this rank's own buffer and are dropped below
        asm volatile(
            "ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%32];\n\t"
            "ld.volatile.global.v4.f32 {%4, %5, %6, %7}, [%33];\n\t"
            "ld.volatile.global.v4.f32 {%8, %9, %10, %11}, [%34];\n\t"
            "ld.volatile.global.v4.f32 {%12, %13, %14, %15}, [%35];\n\t"
            "ld.volatile.global.v4.f32 {%16, %17, %18, %19}, [%36];\n\t"
            "ld.volatile.global.v4.f32 {%20, %21, %22, %23}, [%37];\n\t"
            "ld.volatile.global.v4.f32 {%24, %25, %26, %27}, [%38];\n\t"
            "ld.volatile.global.v4.f32 {%28, %29, %30, %31}, [%39];\n"
            : "=f"(w[0].x), "=f"(w[0].y), "=f"(w[0].z), "=f"(w[0].w), "=f"(w[1].x), "=f"(w[1].y), "=f"(w[1].z), "=f"(w[1].w),
              "=f"(w[2].x), "=f"(w[2].y), "=f"(w[2].z), "=f"(w[2].w), "=f"(w[3].x), "=f"(w[3].y), "=f"(w[3].z), "=f"(w[3].w),
              "=f"(w[4].x), "=f"(w[4].y), "=f"(w[4].z), "=f"(w[4].w), "=f"(w[5].x), "=f"(w[5].y), "=f"(w[5].z), "=f"(w[5].w),
              "=f"(w[6].x), "=f"(w[6].y), "=f"(w[6].z), "=f"(w[6].w), "=f"(w[7].x), "=f"(w[7].y), "=f"(w[7].z), "=f"(w[7].w)
            : "l"(reinterpret_cast<const float4*>(a.grad[j0 + 0]) + i), "l"(reinterpret_cast<const float4*>(a.grad[j0 + 1]) + i),
              "l"(reinterpret_cast<const float4*>(a.grad[j0 + 2]) + i), "l"(reinterpret_cast<const float4*>(a.grad[j0 + 3]) + i),
              "l"(reinterpret_cast<const float4*>(a.grad[j0 + 4]) + i), "l"(reinterpret_cast<const float4*>(a.grad[j0 + 5]) + i),
              "l"(reinterpret_cast<const float4*>(a.grad[j0 + 6]) + i), "l"(reinterpret_cast<const float4*>(a.grad[j0 + 7]) + i)
            : "memory");
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (j0 + j < a.world) { acc.x += w[j].x; acc.y += w[j].y; acc.z += w[j].z; acc.w += w[j].w; }
      }
      reinterpret_cast<float4*>(a.g)[i] = acc;
    } else {
      acc = reinterpret_cast<const float4*>(a.g)[i];
    }
    s += (double)acc.x * acc.x + (double)acc.y * acc.y + (double)acc.z * acc.z + (double)acc.w * acc.w;
  }
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += red[w];
    atomicAdd(a.sumsq, t);
    __threadfence();
    atomicAdd(a.arrived, 1u);
    bool all = false;
    for (int spin = 0; spin < (1 << 24) && !all; spin++) {   // grid barrier: one CTA per SM, all co-resident; bounded all the same
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(a.arrived) : "memory");
      all = v >= gridDim.x;
    }
    if (!all) atomicExch(a.error, 3);
    double total;
    asm volatile("ld.volatile.global.f64 %0, [%1];\n" : "=d"(total) : "l"(a.sumsq) : "memory");
    const float norm = (float)sqrt(total) * a.grad_scale;
    s_coef = a.max_norm > 0.f ? fminf(a.max_norm / (norm + 1e-6f), 1.f) : 1.f;
    if (blockIdx.x == 0 && a.norm_out) *a.norm_out = (double)norm;
  }
  __syncthreads();
  const float gs = a.grad_scale * s_coef;
  const float step_size = a.lr / a.bc1;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 g4 = reinterpret_cast<const float4*>(a.g)[i];
    float4 m4 = reinterpret_cast<float4*>(a.m)[i], v4 = reinterpret_cast<float4*>(a.v)[i], p4 = reinterpret_cast<float4*>(a.p)[i];
    auto upd = [&](float g, float& m, float& v, float& p) {
      g *= gs;
      m = a.beta1 * m + (1.f - a.beta1) * g;
      v = a.beta2 * v + (1.f - a.beta2) * g * g;
      p -= step_size * (m / (sqrtf(v) / a.bc2_sqrt + a.eps));
    };
    upd(g4.x, m4.x, v4.x, p4.x); upd(g4.y, m4.y, v4.y, p4.y); upd(g4.z, m4.z, v4.z, p4.z); upd(g4.w, m4.w, v4.w, p4.w);
    reinterpret_cast<float4*>(a.m)[i] = m4; reinterpret_cast<float4*>(a.v)[i] = v4; reinterpret_cast<float4*>(a.p)[i] = p4;
  }
}

int al4(long long x) { return (int)((x + 3) & ~3LL); }

#include "b2h_ppo_tma.cuh"

}  // namespace

// ------------------------------------------------------------------------------------------------------------- C-ABI
struct B2HPpo {
  B2HPpoConfig cfg;
  int64_t off[13];           // flat offsets: pi W1 b1 W2 b2 W3 b3, vf W1 .. b3, log_std
  int64_t nflat;
  int device;
  // workspace (device)
  float *X, *act, *olp, *adv, *ret;
  float *h1[2], *h2[2], *out[2], *dout[2], *dh2[2], *dh1[2];
  double* scratch;           // [8]: 0..3 statistics, 4 gradient sum of squares, 5 gradient norm
  int* error;                // tensor pipeline timeout flag
  // TMA path (b2h_ppo_tma.cuh): every GEMM operand as two row-major planes (tf32 hi / lo), rows padded to 128, columns to 32
  bool tma;
  float* tbase;
  struct TMat { float *hi, *lo; int ld, rows_pad; };
  TMat tX, th1[2], th2[2], tdh2[2], tdh1[2], tdout[2], tW1[2], tW2[2], tW3[2];
  TMaps maps[8];             // fwd1 fwd2 fwd3 dW3 dh2 dW2 dh1 dW1, each [network][A hi, A lo, B hi, B lo, C hi, C lo]
  uint32_t *bits1[2], *bits2[2];   // sign bits of h1 / h2 (one bit per element, [rows][8] words): the ReLU masks of the backward pass
  int nw_obs;                // N tile of the first-layer weight gradient (obs_dim split into equal tiles <= 256)
  int cluster[8];            // per GEMM: 2 = pairs of CTAs share every B chunk through TMA multicast (its B maps hold half boxes)
  cudaStream_t side;         // the weight-gradient GEMMs of the head and of layer 2 run here, beside the input-gradient chain
  cudaEvent_t ev_fork[2], ev_join, ev_moments;
  double* moments;           // advantage mean / std of the minibatch (device)
  const int64_t* gathered_idx; int gathered_n;   // the minibatch whose rows b2h_ppo_train has already gathered beside the previous apply
  cudaEvent_t ev_gather;
  // peer-memory gradient reduction: [grad copy 0 | grad copy 1 | reduced | flags] in one IPC-exported allocation
  float* comm;
  size_t comm_floats;
  int rank, world;
  uint32_t epoch;
  void* peer_base[P2P_MAX_RANKS];
};

namespace {

int roundup(int x, int m) { return (x + m - 1) / m * m; }

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// K-major use of a plane: 2-D map {columns, rows}, box {32 columns, `rows` rows}, 128-byte swizzle
bool encode_k(EncodeFn encode, CUtensorMap* m, float* plane, int ld, int rows_pad, int rows, int kcols = TK) {
  const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows_pad};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kcols, (cuuint32_t)rows};
  const cuuint32_t estr[2] = {1, 1};
  return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, plane, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                kcols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : (kcols == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// MN-major use: 3-D map {32 columns of a group, rows, column groups}, box {32, 32 rows, `cols` / 32 groups}, 128-byte swizzle
// with 32-byte atoms (what tcgen05's SWIZZLE_128B_BASE32B reads)
bool encode_mn(EncodeFn encode, CUtensorMap* m, float* plane, int ld, int rows_pad, int cols) {
  const cuuint64_t dims[3] = {32, (cuuint64_t)rows_pad, (cuuint64_t)(ld / 32)};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
  const cuuint32_t box[3] = {32, (cuuint32_t)TK, (cuuint32_t)(cols / 32)};
  const cuuint32_t estr[3] = {1, 1, 1};
  return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, plane, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// operand maps of one GEMM: `rows` = 128 for A, the N tile for B
bool operand_maps(EncodeFn enc, CUtensorMap* hi_lo, const B2HPpo::TMat& t, int mn, int rows) {
  if (mn) return encode_mn(enc, hi_lo + 0, t.hi, t.ld, t.rows_pad, rows) && encode_mn(enc, hi_lo + 1, t.lo, t.ld, t.rows_pad, rows);
  return encode_k(enc, hi_lo + 0, t.hi, t.ld, t.rows_pad, rows) && encode_k(enc, hi_lo + 1, t.lo, t.ld, t.rows_pad, rows);
}

cudaError_t gemm_t_attributes() {
  cudaError_t e = cudaFuncSetAttribute(gemm_t_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(t_ns(1) * T_STAGE * sizeof(float)));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_t_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(t_ns(2) * T_STAGE * sizeof(float)));
  return e;
}
// Which instantiation a GEMM runs.  Measured: sending the narrow N tiles (the heads, N = 32: bound by TMA latency, not bytes) to
// the one-CTA-per-SM instantiation, whose 192 KB ring holds four stages of their small chunks, is SLOWER (0.294 against 0.285 ms
// per minibatch): their 256 tiles become 1.7 waves again and cannot share SMs with the 96 KB CTAs of the GEMM on the other stream.
int ctas_for(int nw) { (void)nw; return T_CTAS_DEFAULT; }
cudaError_t launch_gemm_t_kernel(int ctas, dim3 grid, int cluster, cudaStream_t s, const TMaps& maps, const TArgs& a) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(t_threads(ctas)); cfg.dynamicSmemBytes = t_ns(ctas) * T_STAGE * sizeof(float); cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = ctas == 1 ? cudaLaunchKernelEx(&cfg, gemm_t_kernel<1>, maps, a) : cudaLaunchKernelEx(&cfg, gemm_t_kernel<2>, maps, a);
  if (e == cudaSuccess) e = cudaGetLastError();
  return e;
}

int setup_tma(B2HPpo* h) {
  const B2HPpoConfig& c = h->cfg;
  const int H = c.hidden, D = c.obs_dim;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
    g_err_ppo = "cuTensorMapEncodeTiled is not available from this driver";
    return B2H_ECUDA;
  }
  EncodeFn enc = reinterpret_cast<EncodeFn>(fn);
  const int Bp = roundup(c.max_batch, 256), Hp = roundup(H, 128), Dp = roundup(D, 32);   // row tiles come in pairs (clusters of two)
  const int tiles_obs = (Dp + 255) / 256;
  h->nw_obs = roundup((Dp + tiles_obs - 1) / tiles_obs, 32);       // MN-major tiles come in groups of 32 columns
  size_t total = 0;
  auto plan = [&](B2HPpo::TMat& t, int rows_pad, int cols_pad) { t.rows_pad = rows_pad; t.ld = cols_pad; total += 2 * (size_t)rows_pad * cols_pad; };
  plan(h->tX, Bp, std::max(Dp, h->nw_obs * tiles_obs));
  for (int n = 0; n < 2; n++) {
    plan(h->th1[n], Bp, Hp); plan(h->th2[n], Bp, Hp); plan(h->tdh2[n], Bp, Hp); plan(h->tdh1[n], Bp, Hp); plan(h->tdout[n], Bp, OUT_LD);
    plan(h->tW1[n], roundup(H, 128), Dp); plan(h->tW2[n], roundup(H, 128), Hp); plan(h->tW3[n], 128, Hp);
  }
  total += 4 * (size_t)Bp * 8;                                      // the four sign-bit arrays
  if (cudaMalloc(&h->tbase, total * sizeof(float)) != cudaSuccess) { g_err_ppo = "b2h_ppo_create: out of device memory (operand planes)"; return B2H_ENOMEM; }
  cudaMemset(h->tbase, 0, total * sizeof(float));
  float* p = h->tbase;
  auto place = [&](B2HPpo::TMat& t) { const size_t n = (size_t)t.rows_pad * t.ld; t.hi = p; t.lo = p + n; p += 2 * n; };
  place(h->tX);
  for (int n = 0; n < 2; n++) {
    place(h->th1[n]); place(h->th2[n]); place(h->tdh2[n]); place(h->tdh1[n]); place(h->tdout[n]);
    place(h->tW1[n]); place(h->tW2[n]); place(h->tW3[n]);
  }
  for (int n = 0; n < 2; n++) { h->bits1[n] = reinterpret_cast<uint32_t*>(p); p += (size_t)Bp * 8; h->bits2[n] = reinterpret_cast<uint32_t*>(p); p += (size_t)Bp * 8; }
  bool ok = true;
  for (int n = 0; n < 2 && ok; n++) {
    CUtensorMap(*m)[6] = nullptr;
    auto gemm_maps = [&](int g, const B2HPpo::TMat& A, int a_mn, const B2HPpo::TMat& Bm, int b_mn, int nw, const B2HPpo::TMat* Cm) {
      m = &h->maps[g].m[n];
      // M tiles in pairs that share the B tile (batch rows are padded to 256; the weight gradients have H / 128 tiles): each CTA of
      // a pair loads half of every B chunk and multicasts it, so the B maps hold half boxes
      // Measured (B200, minibatch 16384): 0.315 ms with the pairs against 0.305 ms without -- what bounds the main loops is
      // each SM's own ingest (96 KB per K chunk per 0.78 us of MMA = 63 B/clk), which multicast does not lower, not the
      // aggregate L2 traffic it halves; the pairing adds a cross-CTA hand-off per stage.  So it is off unless B2H_PPO_CLUSTER=1
      // (the test entry b2h_gemm_tma keeps exercising the multicast path).
      const bool batch_rows_on_m = !a_mn;
      static const bool want = getenv("B2H_PPO_CLUSTER") && atoi(getenv("B2H_PPO_CLUSTER")) != 0;
      h->cluster[g] = (want && nw % 64 == 0 && (batch_rows_on_m || (roundup(H, 128) / 128) % 2 == 0)) ? 2 : 1;
      // results that are operands of later GEMMs leave through TMA stores of {32 columns, 128 rows} swizzled boxes
      return operand_maps(enc, &(*m)[0], A, a_mn, 128) && operand_maps(enc, &(*m)[2], Bm, b_mn, nw / h->cluster[g]) &&
             (!Cm || (encode_k(enc, &(*m)[4], Cm->hi, Cm->ld, Cm->rows_pad, 128, 32) && encode_k(enc, &(*m)[5], Cm->lo, Cm->ld, Cm->rows_pad, 128, 32)));
    };
    ok = ok && gemm_maps(0, h->tX, 0, h->tW1[n], 0, H, &h->th1[n]);           // fwd1  h1 = relu(X W1^T + b1)
    ok = ok && gemm_maps(1, h->th1[n], 0, h->tW2[n], 0, H, &h->th2[n]);       // fwd2  h2 = relu(h1 W2^T + b2)
    ok = ok && gemm_maps(2, h->th2[n], 0, h->tW3[n], 0, 32, nullptr);         // fwd3  out = h2 W3^T + b3
    ok = ok && gemm_maps(3, h->th2[n], 1, h->tdout[n], 1, 32, nullptr);       // dW3^T = h2^T dout
    ok = ok && gemm_maps(4, h->tdout[n], 0, h->tW3[n], 1, H, &h->tdh2[n]);    // dh2 = dout W3 . (h2 > 0)
    ok = ok && gemm_maps(5, h->tdh2[n], 1, h->th1[n], 1, H, nullptr);         // dW2 = dh2^T h1
    ok = ok && gemm_maps(6, h->tdh2[n], 0, h->tW2[n], 1, H, &h->tdh1[n]);     // dh1 = dh2 W2 . (h1 > 0)
    ok = ok && gemm_maps(7, h->tdh1[n], 1, h->tX, 1, h->nw_obs, nullptr);     // dW1 = dh1^T X
  }
  if (!ok) { g_err_ppo = "cuTensorMapEncodeTiled failed"; return B2H_ECUDA; }
  cudaError_t e = gemm_t_attributes();
  if (e != cudaSuccess) { g_err_ppo = cudaGetErrorString(e); return B2H_ECUDA; }
  if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&h->ev_fork[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fork[1], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_moments, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&h->ev_gather, cudaEventDisableTiming) != cudaSuccess) {
    g_err_ppo = "stream / event creation failed";
    return B2H_ECUDA;
  }
  return B2H_OK;
}

#ifdef B2H_GEMM_CLK
long long* g_clkbuf = nullptr;
#endif
int launch_gemm_t(const B2HPpo* h, int g, TProblem* pr, int precise, bool split, cudaStream_t s) {
  if (g_sm_count <= 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  TArgs a;
  a.p[0] = pr[0]; a.p[1] = pr[1];
  const int tiles = pr[0].m_tiles * pr[0].n_tiles * 2, chunks = pr[0].chunks;
  const int ctas = ctas_for(pr[0].nw);
#ifndef B2H_SPLIT_PER_SM
#define B2H_SPLIT_PER_SM ctas
#endif
  int nsplit = split ? std::max(1, std::min(chunks, g_sm_count * (B2H_SPLIT_PER_SM) / std::max(1, tiles))) : 1;
  a.chunks_per_split = (chunks + nsplit - 1) / nsplit;
  a.nsplit = (chunks + a.chunks_per_split - 1) / a.chunks_per_split;
  a.precise = precise;
  a.error = h->error;
  a.clk = nullptr;
#ifdef B2H_GEMM_CLK
  {   // measurement build: one buffer of time stamps per GEMM of the minibatch (g), read back by tools/gemm_clocks.py
    static long long* clkbuf = nullptr;
    if (!clkbuf) { cudaMalloc(&clkbuf, 8 * 8 * 4096 * sizeof(long long)); cudaMemset(clkbuf, 0, 8 * 8 * 4096 * sizeof(long long)); }
    a.clk = clkbuf + (size_t)g * 8 * 4096;
    g_clkbuf = clkbuf;
  }
#endif
  a.cluster = h->cluster[g];
  if (a.cluster > 1 && pr[0].m_tiles % 2) { g_err_ppo = "internal: odd number of M tiles in a clustered GEMM"; return B2H_EINVAL; }
  cudaError_t e = launch_gemm_t_kernel(ctas, dim3(pr[0].m_tiles, pr[0].n_tiles, 2 * a.nsplit), a.cluster, s, h->maps[g], a);
  if (e != cudaSuccess) { g_err_ppo = cudaGetErrorString(e); return B2H_ECUDA; }
  return B2H_OK;
}

// Minibatch rows -> split observation planes + plain action / log-probability / advantage / return vectors, then the advantage
// moments (one CTA on the side stream).  `s` is the stream the gather runs on: the caller's, or the side stream when
// b2h_ppo_train prefetches the next minibatch beside the apply kernel (nothing of the previous minibatch reads these buffers then).
int gather_stage(B2HPpo* h, const float* obs, const float* actions, const float* old_logp, const float* adv, const float* ret,
                 const int64_t* idx, int64_t row_start, int n, cudaStream_t s) {
  const B2HPpoConfig& c = h->cfg;
  const int rows_pad = roundup(n, 256);
  GatherTArgs g;
  g.obs = obs; g.actions = actions; g.old_logp = old_logp; g.adv = adv; g.ret = ret; g.idx = idx; g.row_start = row_start;
  g.x_hi = h->tX.hi; g.x_lo = h->tX.lo; g.act = h->act; g.olp = h->olp; g.a = h->adv; g.r = h->ret; g.scratch = h->scratch;
  g.n_rows = n; g.rows_pad = rows_pad; g.obs_dim = c.obs_dim; g.ld = h->tX.ld; g.act_dim = c.act_dim;
  gather_t_kernel<<<(rows_pad * 32 + 255) / 256, 256, 0, s>>>(g);
  if (s != h->side) {
    if (cudaEventRecord(h->ev_fork[0], s) != cudaSuccess || cudaStreamWaitEvent(h->side, h->ev_fork[0], 0) != cudaSuccess) { g_err_ppo = "event failed"; return B2H_ECUDA; }
  }
  adv_moments_kernel<<<1, 1024, 0, h->side>>>(h->adv, n, h->moments);
  if (cudaEventRecord(h->ev_moments, h->side) != cudaSuccess) { g_err_ppo = "event failed"; return B2H_ECUDA; }
  return B2H_OK;
}

int minibatch_grad_tma(B2HPpo* h, const float* obs, const float* actions, const float* old_logp, const float* adv, const float* ret,
                       const int64_t* idx, int64_t row_start, int n, const float* P, float* G, cudaStream_t s) {
  const B2HPpoConfig& c = h->cfg;
  const int H = c.hidden, D = c.obs_dim, A = c.act_dim;
  const int64_t* o = h->off;
  const int rows_pad = roundup(n, 256), m_tiles_b = rows_pad / 128, kchunks_b = roundup(n, 32) / TK;
  const int nout[2] = {A, 1};
  // weights -> T-format (they change with every Adam step; 0.3 M floats)
  PackTJobs pj;
  int max_items = 0;
  for (int k = 0; k < 2; k++) {
    const B2HPpo::TMat* t[3] = {&h->tW1[k], &h->tW2[k], &h->tW3[k]};
    const int R[3] = {H, H, nout[k]}, F[3] = {D, H, H};
    for (int l = 0; l < 3; l++) {
      PackTJob& j = pj.j[3 * k + l];
      j.W = P + o[6 * k + 2 * l]; j.hi = t[l]->hi; j.lo = t[l]->lo; j.R = R[l]; j.F = F[l]; j.rows_pad = t[l]->rows_pad; j.ld = t[l]->ld;
      max_items = std::max(max_items, j.rows_pad * j.ld / 4);
    }
  }
  pack_t_kernel<<<dim3((max_items + 255) / 256, 6), 256, 0, s>>>(pj);
  if (cudaMemsetAsync(h->scratch, 0, 4 * sizeof(double), s) != cudaSuccess) { g_err_ppo = "memset failed"; return B2H_ECUDA; }   // the statistics
  if (h->gathered_idx && h->gathered_idx == idx && h->gathered_n == n) {
    // b2h_ppo_train gathered this minibatch on the side stream while the previous minibatch's apply ran
    if (cudaStreamWaitEvent(s, h->ev_gather, 0) != cudaSuccess) { g_err_ppo = "event failed"; return B2H_ECUDA; }
  } else {
    int rc = gather_stage(h, obs, actions, old_logp, adv, ret, idx, row_start, n, s);
    if (rc < 0) return rc;
  }
  h->gathered_idx = nullptr;

  auto base = [&](int a_mn, int b_mn, int m_tiles, int n_tiles, int nw, int chunks) {
    TProblem p;
    p.a_mn = a_mn; p.b_mn = b_mn; p.m_tiles = m_tiles; p.n_tiles = n_tiles; p.nw = nw; p.chunks = chunks; p.M = 0; p.N = 0; p.epi = 0;
    p.c_hi = p.c_lo = nullptr; p.c_ld = 0; p.C = nullptr; p.ldc = 0; p.transpose_c = 0; p.bias = nullptr; p.relu = 0; p.bits_in = nullptr; p.bits_out = nullptr; p.colsum = nullptr;
    return p;
  };
  TProblem pr[2];
  int rc;
  const int kc_obs = roundup(D, 32) / TK, kc_h = roundup(H, 32) / TK, h_tiles = roundup(H, 128) / 128;
  // ---- forward
  for (int k = 0; k < 2; k++) {
    pr[k] = base(0, 0, m_tiles_b, 1, H, kc_obs);
    pr[k].c_hi = h->th1[k].hi; pr[k].c_lo = h->th1[k].lo; pr[k].c_ld = h->th1[k].ld; pr[k].bias = P + o[6 * k + 1]; pr[k].relu = 1; pr[k].N = H; pr[k].bits_out = h->bits1[k];
  }
  if ((rc = launch_gemm_t(h, 0, pr, c.precise, false, s)) < 0) return rc;
  for (int k = 0; k < 2; k++) {
    pr[k] = base(0, 0, m_tiles_b, 1, H, kc_h);
    pr[k].c_hi = h->th2[k].hi; pr[k].c_lo = h->th2[k].lo; pr[k].c_ld = h->th2[k].ld; pr[k].bias = P + o[6 * k + 3]; pr[k].relu = 1; pr[k].N = H; pr[k].bits_out = h->bits2[k];
  }
  if ((rc = launch_gemm_t(h, 1, pr, c.precise, false, s)) < 0) return rc;
  for (int k = 0; k < 2; k++) {
    pr[k] = base(0, 0, m_tiles_b, 1, 32, kc_h);
    pr[k].epi = 1; pr[k].C = h->out[k]; pr[k].ldc = OUT_LD; pr[k].M = n; pr[k].N = nout[k]; pr[k].bias = P + o[6 * k + 5];
  }
  if ((rc = launch_gemm_t(h, 2, pr, c.precise, false, s)) < 0) return rc;
  // ---- loss
  LossArgs L;
  L.mean = h->out[0]; L.value = h->out[1]; L.act = h->act; L.olp = h->olp; L.adv = h->adv; L.ret = h->ret; L.log_std = P + o[12];
  L.dmean = nullptr; L.dvalue = nullptr; L.rows_pad = rows_pad;
  for (int k = 0; k < 2; k++) { L.dt_hi[k] = h->tdout[k].hi; L.dt_lo[k] = h->tdout[k].lo; }
  L.g_b3_pi = G + o[5]; L.g_b3_vf = G + o[11]; L.g_log_std = G + o[12]; L.stats = h->scratch;
  L.n = n; L.act_dim = A; L.clip = c.clip_range; L.ent_coef = c.ent_coef; L.vf_coef = c.vf_coef; L.normalize = c.normalize_advantage;
  L.moments = h->moments;
  if (cudaStreamWaitEvent(s, h->ev_moments, 0) != cudaSuccess) { g_err_ppo = "event failed"; return B2H_ECUDA; }
  ppo_loss_kernel<<<(rows_pad + LOSS_ROWS - 1) / LOSS_ROWS, LOSS_NT, 0, s>>>(L);
  // ---- backward
  for (int k = 0; k < 2; k++) {   // dW3^T [H, nout] = h2^T dout, written transposed into the [nout, H] gradient
    pr[k] = base(1, 1, h_tiles, 1, 32, kchunks_b);
    pr[k].epi = 2; pr[k].C = G + o[6 * k + 4]; pr[k].ldc = H; pr[k].transpose_c = 1; pr[k].M = H; pr[k].N = nout[k];
  }
  // the weight gradients of the head and of layer 2 only feed the flat gradient: they run on a side stream beside the
  // input-gradient chain (dh2 -> dh1 -> dW1) and fill the SMs its partial waves leave idle
  cudaStream_t s2 = h->side;
  if (cudaEventRecord(h->ev_fork[0], s) != cudaSuccess || cudaStreamWaitEvent(s2, h->ev_fork[0], 0) != cudaSuccess) { g_err_ppo = "event failed"; return B2H_ECUDA; }
  if ((rc = launch_gemm_t(h, 3, pr, c.precise, true, s2)) < 0) return rc;
  for (int k = 0; k < 2; k++) {   // dh2 = dout W3 . (h2 > 0)
    pr[k] = base(0, 1, m_tiles_b, 1, H, 32 / TK);
    pr[k].c_hi = h->tdh2[k].hi; pr[k].c_lo = h->tdh2[k].lo; pr[k].c_ld = h->tdh2[k].ld; pr[k].bits_in = h->bits2[k]; pr[k].colsum = G + o[6 * k + 3]; pr[k].N = H;
  }
  if ((rc = launch_gemm_t(h, 4, pr, c.precise, false, s)) < 0) return rc;
  for (int k = 0; k < 2; k++) {   // dW2 = dh2^T h1
    pr[k] = base(1, 1, h_tiles, 1, H, kchunks_b);
    pr[k].epi = 2; pr[k].C = G + o[6 * k + 2]; pr[k].ldc = H; pr[k].M = H; pr[k].N = H;
  }
  if (cudaEventRecord(h->ev_fork[1], s) != cudaSuccess || cudaStreamWaitEvent(s2, h->ev_fork[1], 0) != cudaSuccess) { g_err_ppo = "event failed"; return B2H_ECUDA; }
  if ((rc = launch_gemm_t(h, 5, pr, c.precise, true, s2)) < 0) return rc;
  if (cudaEventRecord(h->ev_join, s2) != cudaSuccess) { g_err_ppo = "event failed"; return B2H_ECUDA; }
  for (int k = 0; k < 2; k++) {   // dh1 = dh2 W2 . (h1 > 0)
    pr[k] = base(0, 1, m_tiles_b, 1, H, kc_h);
    pr[k].c_hi = h->tdh1[k].hi; pr[k].c_lo = h->tdh1[k].lo; pr[k].c_ld = h->tdh1[k].ld; pr[k].bits_in = h->bits1[k]; pr[k].colsum = G + o[6 * k + 1]; pr[k].N = H;
  }
  if ((rc = launch_gemm_t(h, 6, pr, c.precise, false, s)) < 0) return rc;
  const int obs_tiles = (roundup(D, 32) + 255) / 256;
  for (int k = 0; k < 2; k++) {   // dW1 = dh1^T X
    pr[k] = base(1, 1, h_tiles, obs_tiles, h->nw_obs, kchunks_b);
    pr[k].epi = 2; pr[k].C = G + o[6 * k + 0]; pr[k].ldc = D; pr[k].M = H; pr[k].N = D;
  }
  if ((rc = launch_gemm_t(h, 7, pr, c.precise, true, s)) < 0) return rc;
  if (cudaStreamWaitEvent(s, h->ev_join, 0) != cudaSuccess) { g_err_ppo = "event failed"; return B2H_ECUDA; }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_err_ppo = cudaGetErrorString(e); return B2H_ECUDA; }
  return B2H_OK;
}

}  // namespace

extern "C" {

const char* b2h_ppo_last_error(void) { return g_err_ppo.c_str(); }
size_t b2h_sizeof_ppo_config(void) { return sizeof(B2HPpoConfig); }

int64_t b2h_ppo_param_layout(int obs_dim, int hidden, int act_dim, int64_t offsets[13]) {
  const long long sizes[13] = {(long long)hidden * obs_dim, hidden, (long long)hidden * hidden, hidden, (long long)act_dim * hidden, act_dim,
                               (long long)hidden * obs_dim, hidden, (long long)hidden * hidden, hidden, hidden, 1, act_dim};
  long long o = 0;
  for (int i = 0; i < 13; i++) {
    if (offsets) offsets[i] = o;
    o = al4(o + sizes[i]);
  }
  return o;
}

int b2h_gemm(const float* a_dev, int lda, int a_kstrided, const float* b_dev, int ldb, int b_kstrided, float* c_dev, int ldc, int transpose_c,
             const float* bias_dev, const float* mask_dev, int ldmask, int m, int n, int k, int relu, int precise, int split_k, int accumulate,
             int* error_flag_dev, void* stream) {
  if (!a_dev || !b_dev || !c_dev || !error_flag_dev || m <= 0 || n <= 0 || k <= 0 || split_k < 0) { g_err_ppo = "b2h_gemm: bad argument"; return B2H_EINVAL; }
  if ((a_kstrided ? lda < m : lda < k) || (b_kstrided ? ldb < n : ldb < k) || (transpose_c ? ldc < m : ldc < n) || (mask_dev && ldmask < n)) {
    g_err_ppo = "b2h_gemm: leading dimension too small";
    return B2H_EINVAL;
  }
  GemmProblem p;
  p.A = a_dev; p.B = b_dev; p.bias = bias_dev; p.mask = mask_dev; p.C = c_dev; p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.ldmask = ldmask;
  p.M = m; p.N = n; p.K = k; p.a_kstrided = a_kstrided; p.b_kstrided = b_kstrided; p.relu = relu; p.transpose_c = transpose_c;
  p.atomic = (accumulate || split_k != 1) ? 1 : 0;
  return launch_gemm(&p, 1, precise, split_k, error_flag_dev, (cudaStream_t)stream);
}

// Test / measurement entry of the TMA-fed GEMM: plain row-major operands are packed into T-format here (device work + two
// temporary allocations; synchronises), then C (+)= A . B^T runs through gemm_t_kernel.
int b2h_gemm_tma(const float* a_dev, int a_mn, const float* b_dev, int b_mn, float* c_dev, int ldc, int transpose_c, const float* bias_dev,
                 int m, int n, int k, int precise, int split_k, int* error_flag_dev, void* stream) {
  if (!a_dev || !b_dev || !c_dev || !error_flag_dev || m <= 0 || n <= 0 || k <= 0) { g_err_ppo = "b2h_gemm_tma: bad argument"; return B2H_EINVAL; }
  cudaStream_t s = (cudaStream_t)stream;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
    g_err_ppo = "cuTensorMapEncodeTiled is not available from this driver";
    return B2H_ECUDA;
  }
  EncodeFn enc = reinterpret_cast<EncodeFn>(fn);
  const int n_tiles = (roundup(n, 32) + 255) / 256, nw = roundup((roundup(n, 32) + n_tiles - 1) / n_tiles, 32);
  // A is [m, k] (K-major use) or [k, m] (MN-major use); likewise B with n
  const int ar = a_mn ? k : m, ac = a_mn ? m : k, br = b_mn ? k : n, bc = b_mn ? n : k;
  B2HPpo::TMat ta, tb;
  ta.rows_pad = roundup(ar, 128); ta.ld = roundup(std::max(ac, a_mn ? 128 : 32), 32);
  tb.rows_pad = roundup(br, 256); tb.ld = roundup(std::max(bc, b_mn ? nw * n_tiles : 32), 32);
  const size_t na = (size_t)ta.rows_pad * ta.ld, nb = (size_t)tb.rows_pad * tb.ld;
  float* buf = nullptr;
  if (cudaMalloc(&buf, 2 * (na + nb) * sizeof(float)) != cudaSuccess) { g_err_ppo = "out of device memory"; return B2H_ENOMEM; }
  ta.hi = buf; ta.lo = buf + na; tb.hi = buf + 2 * na; tb.lo = tb.hi + nb;
  PackTJobs pj;
  pj.j[0] = PackTJob{a_dev, ta.hi, ta.lo, ar, ac, ta.rows_pad, ta.ld};
  pj.j[1] = PackTJob{b_dev, tb.hi, tb.lo, br, bc, tb.rows_pad, tb.ld};
  const int items = (int)std::max(na, nb) / 4;
  pack_t_kernel<<<dim3((items + 255) / 256, 2), 256, 0, s>>>(pj);
  B2HPpo hh;
  const int cluster = (nw % 64 == 0 && ((m + 127) / 128) % 2 == 0) ? 2 : 1;   // pairs of M tiles share the B chunks (TMA multicast)
  bool ok = operand_maps(enc, &hh.maps[0].m[0][0], ta, a_mn, 128) && operand_maps(enc, &hh.maps[0].m[0][2], tb, b_mn, nw / cluster);
  for (int i = 0; i < 4; i++) hh.maps[0].m[1][i] = hh.maps[0].m[0][i];
  for (int i = 4; i < 6; i++) { hh.maps[0].m[0][i] = hh.maps[0].m[0][0]; hh.maps[0].m[1][i] = hh.maps[0].m[0][0]; }
  int rc = B2H_OK;
  if (!ok) { g_err_ppo = "cuTensorMapEncodeTiled failed"; rc = B2H_ECUDA; }
  if (rc == B2H_OK && gemm_t_attributes() != cudaSuccess) {
    g_err_ppo = "cudaFuncSetAttribute failed"; rc = B2H_ECUDA;
  }
  if (rc == B2H_OK) {
    TProblem p;
    p.a_mn = a_mn; p.b_mn = b_mn; p.m_tiles = (m + 127) / 128; p.n_tiles = n_tiles; p.nw = nw; p.chunks = (k + TK - 1) / TK; p.M = m; p.N = n;
    p.epi = split_k == 1 ? 1 : 2; p.c_hi = p.c_lo = nullptr; p.c_ld = 0; p.C = c_dev; p.ldc = ldc; p.transpose_c = transpose_c; p.bias = bias_dev; p.relu = 0;
    p.bits_in = nullptr; p.bits_out = nullptr; p.colsum = nullptr;
    TArgs a;
    a.p[0] = p; a.p[1] = p;
    int nsplit = split_k == 1 ? 1 : (split_k > 1 ? std::min(split_k, p.chunks) : std::max(1, std::min(p.chunks, 148 / (p.m_tiles * p.n_tiles))));
    a.chunks_per_split = (p.chunks + nsplit - 1) / nsplit;
    a.nsplit = (p.chunks + a.chunks_per_split - 1) / a.chunks_per_split;
    a.precise = precise; a.error = error_flag_dev; a.cluster = cluster; a.clk = nullptr;
    cudaError_t le = launch_gemm_t_kernel(ctas_for(nw), dim3(p.m_tiles, p.n_tiles, a.nsplit), cluster, s, hh.maps[0], a);
    if (le != cudaSuccess || cudaGetLastError() != cudaSuccess) { g_err_ppo = std::string("launch failed: ") + cudaGetErrorString(le); rc = B2H_ECUDA; }
  }
  if (cudaStreamSynchronize(s) != cudaSuccess && rc == B2H_OK) { g_err_ppo = cudaGetErrorString(cudaGetLastError()); rc = B2H_ECUDA; }
  cudaFree(buf);
  return rc;
}

int b2h_ppo_create(const B2HPpoConfig* cfg, B2HPpo** out) {
  if (!cfg || !out) { g_err_ppo = "null argument"; return B2H_EINVAL; }
  if (cfg->obs_dim <= 0 || cfg->hidden <= 0 || cfg->hidden % 4 || cfg->act_dim <= 0 || cfg->act_dim > MAX_ACT || cfg->max_batch <= 0) {
    g_err_ppo = "b2h_ppo_create: need hidden % 4 == 0, 0 < act_dim <= 32, max_batch > 0";
    return B2H_EINVAL;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_err_ppo = "no CUDA device (this library has no CPU path)"; return B2H_ECUDA; }
  B2HPpo* h = new B2HPpo();
  h->cfg = *cfg;
  h->nflat = b2h_ppo_param_layout(cfg->obs_dim, cfg->hidden, cfg->act_dim, h->off);
  cudaGetDevice(&h->device);
  const size_t B = (size_t)cfg->max_batch, H = (size_t)cfg->hidden;
  h->tma = !cfg->staged_operands && cfg->hidden % 32 == 0 && cfg->hidden <= 256;
  // plain row-major workspace: the minibatch vectors and head outputs always; the activations only for the staged GEMM (the TMA
  // path keeps them as operand planes, setup_tma)
  const size_t staged = h->tma ? 0 : B * cfg->obs_dim + 2 * (4 * B * H + B * OUT_LD);
  const size_t floats = staged + B * cfg->act_dim + 3 * B + 2 * B * OUT_LD + 1024;
  float* base = nullptr;
  if (cudaMalloc(&base, floats * sizeof(float)) != cudaSuccess) {
    g_err_ppo = "b2h_ppo_create: out of device memory";
    delete h;
    return B2H_ENOMEM;
  }
  cudaMemset(base, 0, floats * sizeof(float));
  float* p = base;
  auto take = [&](size_t n) { float* r = p; p += (n + 3) & ~(size_t)3; return r; };
  h->scratch = reinterpret_cast<double*>(take(32));
  h->moments = h->scratch + 8;
  h->error = reinterpret_cast<int*>(take(4));
  h->act = take(B * cfg->act_dim); h->olp = take(B); h->adv = take(B); h->ret = take(B);
  h->X = nullptr;
  for (int n = 0; n < 2; n++) {
    h->out[n] = take(B * OUT_LD);
    h->h1[n] = h->h2[n] = h->dh2[n] = h->dh1[n] = h->dout[n] = nullptr;
  }
  if (!h->tma) {
    h->X = take(B * cfg->obs_dim);
    for (int n = 0; n < 2; n++) { h->h1[n] = take(B * H); h->h2[n] = take(B * H); h->dh2[n] = take(B * H); h->dh1[n] = take(B * H); h->dout[n] = take(B * OUT_LD); }
  }
  h->comm = nullptr; h->comm_floats = 0; h->rank = 0; h->world = 1; h->epoch = 0;
  h->side = nullptr; h->ev_fork[0] = h->ev_fork[1] = h->ev_join = h->ev_moments = h->ev_gather = nullptr;
  h->gathered_idx = nullptr; h->gathered_n = 0;
  for (int i = 0; i < P2P_MAX_RANKS; i++) h->peer_base[i] = nullptr;
  h->tbase = nullptr;
  if (h->tma) {
    const int rc = setup_tma(h);
    if (rc < 0) { b2h_ppo_destroy(h); return rc; }
  }
  *out = h;
  return B2H_OK;
}

void b2h_ppo_destroy(B2HPpo* h) {
  if (!h) return;
  cudaFree(h->scratch);   // the first allocation of the block
  if (h->tbase) cudaFree(h->tbase);
  if (h->ev_join) { cudaEventDestroy(h->ev_fork[0]); cudaEventDestroy(h->ev_fork[1]); cudaEventDestroy(h->ev_join); }
  if (h->ev_moments) cudaEventDestroy(h->ev_moments);
  if (h->ev_gather) cudaEventDestroy(h->ev_gather);
  if (h->side) cudaStreamDestroy(h->side);
  for (int i = 0; i < P2P_MAX_RANKS; i++)
    if (h->peer_base[i] && i != h->rank) cudaIpcCloseMemHandle(h->peer_base[i]);
  if (h->comm) cudaFree(h->comm);
  delete h;
}

int b2h_ppo_minibatch_grad(B2HPpo* h, const float* obs_dev, const float* actions_dev, const float* old_log_probs_dev,
                           const float* advantages_dev, const float* returns_dev, const int64_t* idx_dev, int64_t row_start, int n_rows,
                           const float* params_dev, float* grad_dev, void* stream_) {
  if (!h || !obs_dev || !actions_dev || !old_log_probs_dev || !advantages_dev || !returns_dev || !params_dev || !grad_dev) {
    g_err_ppo = "null argument";
    return B2H_EINVAL;
  }
  if (n_rows <= 0 || n_rows > h->cfg.max_batch) { g_err_ppo = "b2h_ppo_minibatch_grad: n_rows outside (0, max_batch]"; return B2H_EINVAL; }
  cudaStream_t s = (cudaStream_t)stream_;
  const B2HPpoConfig& c = h->cfg;
  const int H = c.hidden, D = c.obs_dim, A = c.act_dim, n = n_rows;
  if (cudaMemsetAsync(grad_dev, 0, (size_t)h->nflat * sizeof(float), s) != cudaSuccess) { g_err_ppo = "memset failed"; return B2H_ECUDA; }
  if (h->tma)
    return minibatch_grad_tma(h, obs_dev, actions_dev, old_log_probs_dev, advantages_dev, returns_dev, idx_dev, row_start, n, params_dev, grad_dev, s);
  GatherArgs g;
  g.obs = obs_dev; g.actions = actions_dev; g.old_logp = old_log_probs_dev; g.adv = advantages_dev; g.ret = returns_dev; g.idx = idx_dev;
  g.row_start = row_start; g.X = h->X; g.act = h->act; g.olp = h->olp; g.a = h->adv; g.r = h->ret; g.scratch = h->scratch;
  g.n_rows = n; g.obs_dim = D; g.act_dim = A;
  gather_kernel<<<(n * 32 + 255) / 256, 256, 0, s>>>(g);

  const float* P = params_dev;
  float* G = grad_dev;
  const int64_t* o = h->off;
  const int nout[2] = {A, 1};
  auto prob = [](const float* Am, int lda, int ak, const float* Bm, int ldb, int bk, float* C, int ldc, int M, int N, int K) {
    GemmProblem p;
    p.A = Am; p.B = Bm; p.C = C; p.bias = nullptr; p.mask = nullptr; p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.ldmask = 0; p.M = M; p.N = N; p.K = K;
    p.a_kstrided = ak; p.b_kstrided = bk; p.relu = 0; p.atomic = 0; p.transpose_c = 0;
    return p;
  };
  GemmProblem pr[2];
  int rc;
  // ---- forward: h1 = relu(X W1^T + b1), h2 = relu(h1 W2^T + b2), out = h2 W3^T + b3
  for (int k = 0; k < 2; k++) { pr[k] = prob(h->X, D, 0, P + o[6 * k + 0], D, 0, h->h1[k], H, n, H, D); pr[k].bias = P + o[6 * k + 1]; pr[k].relu = 1; }
  if ((rc = launch_gemm(pr, 2, c.precise, 1, h->error, s)) < 0) return rc;
  for (int k = 0; k < 2; k++) { pr[k] = prob(h->h1[k], H, 0, P + o[6 * k + 2], H, 0, h->h2[k], H, n, H, H); pr[k].bias = P + o[6 * k + 3]; pr[k].relu = 1; }
  if ((rc = launch_gemm(pr, 2, c.precise, 1, h->error, s)) < 0) return rc;
  for (int k = 0; k < 2; k++) { pr[k] = prob(h->h2[k], H, 0, P + o[6 * k + 4], H, 0, h->out[k], OUT_LD, n, nout[k], H); pr[k].bias = P + o[6 * k + 5]; }
  if ((rc = launch_gemm(pr, 2, c.precise, 1, h->error, s)) < 0) return rc;
  // ---- loss and its gradient with respect to the head outputs
  LossArgs L;
  L.mean = h->out[0]; L.value = h->out[1]; L.act = h->act; L.olp = h->olp; L.adv = h->adv; L.ret = h->ret; L.log_std = P + o[12];
  L.dmean = h->dout[0]; L.dvalue = h->dout[1]; L.rows_pad = 0; L.dt_hi[0] = L.dt_hi[1] = L.dt_lo[0] = L.dt_lo[1] = nullptr; L.g_b3_pi = G + o[5]; L.g_b3_vf = G + o[11]; L.g_log_std = G + o[12]; L.stats = h->scratch;
  L.n = n; L.act_dim = A; L.clip = c.clip_range; L.ent_coef = c.ent_coef; L.vf_coef = c.vf_coef; L.normalize = c.normalize_advantage;
  L.moments = h->moments;
  adv_moments_kernel<<<1, 1024, 0, s>>>(h->adv, n, h->moments);
  ppo_loss_kernel<<<(n + LOSS_ROWS - 1) / LOSS_ROWS, LOSS_NT, 0, s>>>(L);
  // ---- backward.  Head: dW3 = dout^T h2 (computed transposed: the 256 hidden features ride on the M side), dh2 = dout W3 . (h2 > 0)
  for (int k = 0; k < 2; k++) {
    pr[k] = prob(h->h2[k], H, 1, h->dout[k], OUT_LD, 1, G + o[6 * k + 4], H, H, nout[k], n);
    pr[k].atomic = 1; pr[k].transpose_c = 1;
  }
  if ((rc = launch_gemm(pr, 2, c.precise, 0, h->error, s)) < 0) return rc;
  for (int k = 0; k < 2; k++) { pr[k] = prob(h->dout[k], OUT_LD, 0, P + o[6 * k + 4], H, 1, h->dh2[k], H, n, H, nout[k]); pr[k].mask = h->h2[k]; pr[k].ldmask = H; }
  if ((rc = launch_gemm(pr, 2, c.precise, 1, h->error, s)) < 0) return rc;
  // layer 2: dW2 = dh2^T h1, dh1 = dh2 W2 . (h1 > 0)
  for (int k = 0; k < 2; k++) { pr[k] = prob(h->dh2[k], H, 1, h->h1[k], H, 1, G + o[6 * k + 2], H, H, H, n); pr[k].atomic = 1; }
  if ((rc = launch_gemm(pr, 2, c.precise, 0, h->error, s)) < 0) return rc;
  for (int k = 0; k < 2; k++) { pr[k] = prob(h->dh2[k], H, 0, P + o[6 * k + 2], H, 1, h->dh1[k], H, n, H, H); pr[k].mask = h->h1[k]; pr[k].ldmask = H; }
  if ((rc = launch_gemm(pr, 2, c.precise, 1, h->error, s)) < 0) return rc;
  // layer 1: dW1 = dh1^T X
  for (int k = 0; k < 2; k++) { pr[k] = prob(h->dh1[k], H, 1, h->X, D, 1, G + o[6 * k + 0], D, H, D, n); pr[k].atomic = 1; }
  if ((rc = launch_gemm(pr, 2, c.precise, 0, h->error, s)) < 0) return rc;
  // hidden-layer biases
  ColsumArgs cs;
  cs.src[0] = h->dh2[0]; cs.dst[0] = G + o[3]; cs.src[1] = h->dh1[0]; cs.dst[1] = G + o[1];
  cs.src[2] = h->dh2[1]; cs.dst[2] = G + o[9]; cs.src[3] = h->dh1[1]; cs.dst[3] = G + o[7];
  cs.n = n; cs.width = H; cs.rows_per_cta = 128;
  colsum_kernel<<<dim3((n + 127) / 128, 4), 256, 0, s>>>(cs);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_err_ppo = cudaGetErrorString(e); return B2H_ECUDA; }
  return B2H_OK;
}

namespace {
int launch_apply(B2HPpo* h, ApplyArgs& a, int64_t step, cudaStream_t s) {
  const B2HPpoConfig& c = h->cfg;
  if (g_sm_count <= 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  if (cudaMemsetAsync(h->scratch + 4, 0, 3 * sizeof(double), s) != cudaSuccess) { g_err_ppo = "memset failed"; return B2H_ECUDA; }
  a.n = h->nflat; a.sumsq = h->scratch + 4; a.norm_out = h->scratch + 5; a.arrived = reinterpret_cast<unsigned*>(h->scratch + 6); a.error = h->error;
  a.max_norm = c.max_grad_norm; a.lr = c.lr; a.beta1 = c.beta1; a.beta2 = c.beta2; a.eps = c.adam_eps;
  a.bc1 = (float)(1.0 - pow((double)c.beta1, (double)step));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)c.beta2, (double)step));
  const int blocks = (int)std::max<long long>(1, std::min<long long>(g_sm_count, (h->nflat / 4 + 255) / 256));
  apply_kernel<<<blocks, 256, 0, s>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_err_ppo = cudaGetErrorString(e); return B2H_ECUDA; }
  return B2H_OK;
}
}  // namespace

int b2h_ppo_apply(B2HPpo* h, float* params_dev, float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t step, float grad_scale,
                  void* stream_) {
  if (!h || !params_dev || !grad_dev || !exp_avg_dev || !exp_avg_sq_dev || step < 1) { g_err_ppo = "b2h_ppo_apply: bad argument"; return B2H_EINVAL; }
  ApplyArgs a;
  a.p = params_dev; a.g = grad_dev; a.m = exp_avg_dev; a.v = exp_avg_sq_dev; a.grad_scale = grad_scale;
  a.rank = 0; a.world = 1; a.epoch = 0;
  for (int j = 0; j < P2P_MAX_RANKS; j++) { a.grad[j] = grad_dev; a.flags[j] = nullptr; }
  return launch_apply(h, a, step, (cudaStream_t)stream_);
}

// ---- peer-memory gradient reduction (one node, one process per GPU)
int b2h_ppo_p2p_export(B2HPpo* h, void* ipc_handle_out64) {
  if (!h || !ipc_handle_out64) { g_err_ppo = "null argument"; return B2H_EINVAL; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  if (!h->comm) {
    h->comm_floats = 3 * (size_t)h->nflat + 64;
    if (cudaMalloc(&h->comm, h->comm_floats * sizeof(float)) != cudaSuccess) { g_err_ppo = "out of device memory (gradient exchange buffers)"; return B2H_ENOMEM; }
    cudaMemset(h->comm, 0, h->comm_floats * sizeof(float));
    cudaDeviceSynchronize();
  }
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, h->comm);
  if (e != cudaSuccess) { g_err_ppo = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e); return B2H_ECUDA; }
  memcpy(ipc_handle_out64, &hd, 64);
  return B2H_OK;
}

int b2h_ppo_p2p_attach(B2HPpo* h, int rank, int world, const void* ipc_handles) {
  if (!h || !ipc_handles || !h->comm || world < 1 || world > P2P_MAX_RANKS || rank < 0 || rank >= world) { g_err_ppo = "b2h_ppo_p2p_attach: bad argument (export first)"; return B2H_EINVAL; }
  h->rank = rank; h->world = world;
  for (int j = 0; j < world; j++) {
    if (j == rank) { h->peer_base[j] = h->comm; continue; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, static_cast<const char*>(ipc_handles) + 64 * j, 64);
    cudaError_t e = cudaIpcOpenMemHandle(&h->peer_base[j], hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { g_err_ppo = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); h->peer_base[j] = nullptr; return B2H_ECUDA; }
  }
  return B2H_OK;
}

float* b2h_ppo_p2p_grad(B2HPpo* h) { return h && h->comm ? h->comm + (size_t)((h->epoch + 1) & 1) * h->nflat : nullptr; }

int b2h_ppo_apply_p2p(B2HPpo* h, float* params_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t step, void* stream_) {
  if (!h || !h->comm || !h->peer_base[h->world - 1] || !params_dev || !exp_avg_dev || !exp_avg_sq_dev || step < 1) { g_err_ppo = "b2h_ppo_apply_p2p: not attached / bad argument"; return B2H_EINVAL; }
  h->epoch++;
  ApplyArgs a;
  const size_t copy = (size_t)(h->epoch & 1) * h->nflat;
  for (int j = 0; j < h->world; j++) {
    float* base = static_cast<float*>(h->peer_base[j]);
    a.grad[j] = base + copy;
    a.flags[j] = reinterpret_cast<uint32_t*>(base + 3 * (size_t)h->nflat);
  }
  for (int j = h->world; j < P2P_MAX_RANKS; j++) { a.grad[j] = a.grad[h->rank]; a.flags[j] = a.flags[h->rank]; }
  a.p = params_dev; a.g = h->comm + 2 * (size_t)h->nflat; a.m = exp_avg_dev; a.v = exp_avg_sq_dev; a.grad_scale = 1.f / (float)h->world;
  a.rank = h->rank; a.world = h->world; a.epoch = h->epoch;
  return launch_apply(h, a, step, (cudaStream_t)stream_);
}

int b2h_ppo_train(B2HPpo* h, const float* obs_dev, const float* actions_dev, const float* old_log_probs_dev, const float* advantages_dev,
                  const float* returns_dev, const int64_t* perm_dev, int64_t n_samples, int n_epochs, int batch_size, float* params_dev,
                  float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t* step_inout, void* stream) {
  if (!h || !perm_dev || !step_inout || n_samples <= 0 || n_epochs <= 0 || batch_size <= 0) { g_err_ppo = "b2h_ppo_train: bad argument"; return B2H_EINVAL; }
  const bool p2p = h->comm && h->world > 1 && h->peer_base[h->world - 1];     // attached: gradients summed over the ranks by peer loads
  h->gathered_idx = nullptr;   // a prefetch left behind by a call that failed half-way must not be taken for this call's first minibatch
  for (int e = 0; e < n_epochs; e++)
    for (int64_t i = 0; i < n_samples; i += batch_size) {
      const int n = (int)std::min<int64_t>(batch_size, n_samples - i);
      float* g = p2p ? b2h_ppo_p2p_grad(h) : grad_dev;
      int rc = b2h_ppo_minibatch_grad(h, obs_dev, actions_dev, old_log_probs_dev, advantages_dev, returns_dev, perm_dev + (size_t)e * n_samples + i,
                                      0, n, params_dev, g, stream);
      if (rc < 0) return rc;
      // the next minibatch's rows are gathered on the side stream while this one's apply kernel runs: after the last GEMM nothing
      // reads the observation planes or the minibatch vectors any more, and the gather does not depend on the weights
      int64_t ni = i + batch_size;
      int ne = e;
      if (ni >= n_samples) { ni = 0; ne = e + 1; }
      if (h->tma && ne < n_epochs) {
        cudaStream_t s = (cudaStream_t)stream;
        const int nn = (int)std::min<int64_t>(batch_size, n_samples - ni);
        const int64_t* nidx = perm_dev + (size_t)ne * n_samples + ni;
        if (cudaEventRecord(h->ev_fork[0], s) != cudaSuccess || cudaStreamWaitEvent(h->side, h->ev_fork[0], 0) != cudaSuccess) { g_err_ppo = "event failed"; return B2H_ECUDA; }
        rc = gather_stage(h, obs_dev, actions_dev, old_log_probs_dev, advantages_dev, returns_dev, nidx, 0, nn, h->side);
        if (rc < 0) return rc;
        if (cudaEventRecord(h->ev_gather, h->side) != cudaSuccess) { g_err_ppo = "event failed"; return B2H_ECUDA; }
        h->gathered_idx = nidx; h->gathered_n = nn;
      }
      rc = p2p ? b2h_ppo_apply_p2p(h, params_dev, exp_avg_dev, exp_avg_sq_dev, ++*step_inout, stream)
               : b2h_ppo_apply(h, params_dev, grad_dev, exp_avg_dev, exp_avg_sq_dev, ++*step_inout, 1.f, stream);
      if (rc < 0) return rc;
    }
  return B2H_OK;
}

int b2h_ppo_stats(B2HPpo* h, double stats_host[8], int* error_host, void* stream) {
  if (!h || !stats_host) { g_err_ppo = "null argument"; return B2H_EINVAL; }
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemcpyAsync(stats_host, h->scratch, 8 * sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
      (error_host && cudaMemcpyAsync(error_host, h->error, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess) ||
      cudaStreamSynchronize(s) != cudaSuccess) {
    g_err_ppo = cudaGetErrorString(cudaGetLastError());
    return B2H_ECUDA;
  }
  return B2H_OK;
}

#ifdef B2H_GEMM_CLK
int b2h_ppo_gemm_clocks(long long* out_host) {   // [8 GEMMs][4096 CTAs][8 stamps]
  if (!g_clkbuf) return B2H_EINVAL;
  cudaDeviceSynchronize();
  return cudaMemcpy(out_host, g_clkbuf, 8 * 8 * 4096 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? B2H_OK : B2H_ECUDA;
}
#endif
const double* b2h_ppo_stats_dev(const B2HPpo* h) { return h ? h->scratch : nullptr; }
const int* b2h_ppo_error_dev(const B2HPpo* h) { return h ? h->error : nullptr; }

}  // extern "C"
