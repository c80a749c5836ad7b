// b2h_mlp.cu — policy / value MLP forward of the rollout (SB3 MlpPolicy, train_sb3.py:208-214; main.py:99-105:
// separate pi / vf trunks, in -> 256 -> 256 -> out, ReLU) on the 5th-generation tensor cores, sm_100a.
//
// One CTA pushes a 128-row tile of the batch through all three layers of one network without leaving the SM:
//   * operands are staged into shared memory in the canonical no-swizzle K-major UMMA layout (8x16-byte core
//     matrices) by all 256 threads; weights come from global memory (L2-resident, 0.6 MB per network), the
//     activations of layers 2 and 3 from the previous layer's epilogue in shared memory;
//   * one elected thread issues tcgen05.mma kind::tf32 (M=128, N=256 / 32, K=8) with the accumulator in TMEM;
//     completion is signalled through tcgen05.commit -> mbarrier; the operand stage is a two-deep ring, so the
//     tensor core works on chunk k while the threads stage chunk k+1 (a stage is only reused after the MMAs that
//     read it have committed);
//   * the epilogue reads the accumulator with tcgen05.ld (each warp its 32-lane quadrant), adds the bias, applies
//     ReLU and writes the next layer's input (or the network output).
// precise = 1 splits every operand into tf32 hi + lo parts and issues hi*hi + hi*lo + lo*hi, which recovers
// fp32-level accuracy (the reference policy runs in fp32 on the CPU); precise = 0 is a single tf32 pass.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>

#include "../../include/b2h.h"

namespace {

constexpr int TILE_M = 128;   // batch rows per CTA (UMMA M)
#ifndef B2H_MLP_KC
#define B2H_MLP_KC 16
#endif
#ifndef B2H_MLP_NSTAGE
#define B2H_MLP_NSTAGE 1
#endif
#ifndef B2H_MLP_PF
#define B2H_MLP_PF 1
#endif
constexpr int KC = B2H_MLP_KC;          // K columns staged per round (KC / 8 MMA K-steps)
constexpr int PF = B2H_MLP_PF;          // chunks prefetched into registers ahead of the staging (multiple of NSTAGE)
static_assert(B2H_MLP_PF % B2H_MLP_NSTAGE == 0, "the stage of an unrolled ring slot must be a compile-time constant");
constexpr int NSTAGE = B2H_MLP_NSTAGE;  // operand stages (KC 32 x 1 stage or KC 16 x 2 stages fill the 227 KB next to H)
constexpr int MAXH = 256;     // hidden width (UMMA N) supported
constexpr int HSTRIDE = MAXH + 4;  // fp32 row stride of the activation buffer (16-byte shift per row: conflict-free float4)
#ifndef B2H_MLP_THREADS
#define B2H_MLP_THREADS 256
#endif
constexpr int NTHREADS = B2H_MLP_THREADS;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle, K-major: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}
// instruction descriptor kind::tf32: D fp32 (bit 4), A/B tf32 (2 at bits 7 and 10), both K-major, N>>3 at 17, M>>4 at 24
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

struct MlpArgs {     // up to two networks (blockIdx.y) over the same input: policy and value trunks of MlpPolicy
  const float* x;
  const float *w[2][3], *b[2][3];
  float* y[2];
  int out_dim[2];
  int n_rows, in_dim, hidden, precise;
  int* error;
};

// One K-chunk (KC columns) of an operand travels global/shared -> registers -> shared (UMMA core-matrix layout,
// split into tf32 hi and lo parts).  The loads of chunk k+1 are issued before the wait on the MMAs of chunk k.
template <int NV>
struct ChunkRegs { float4 v[NV]; };

// Work item f -> (row r, 16-byte K group k4): eight consecutive threads take eight consecutive rows of the same K
// group, i.e. one whole 128-byte core matrix per quarter-warp (conflict-free shared stores; the global side still
// reads 64 contiguous bytes of each of 8 rows per warp), then the K groups, then the next 8 rows.
__device__ __forceinline__ void chunk_item(int f, int& r, int& k4) {
  constexpr int G = KC / 4;
  r = (f & 7) | ((f / (8 * G)) << 3);
  k4 = (f >> 3) % G;
}
template <int NV>
__device__ __forceinline__ void load_chunk(ChunkRegs<NV>& c, const float* src, int ld, int rows, int valid_rows, int k0, int row_base) {
  const int nvec = rows * (KC / 4);
#pragma unroll
  for (int i = 0; i < NV; i++) {
    int f = threadIdx.x + i * NTHREADS, r, k4;
    chunk_item(f, r, k4);
    c.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f < nvec && r < valid_rows) c.v[i] = *reinterpret_cast<const float4*>(src + (size_t)(row_base + r) * ld + k0 + 4 * k4);
  }
}
template <int NV>
__device__ __forceinline__ void store_chunk(const ChunkRegs<NV>& c, float* dst_hi, float* dst_lo, int rows) {
  const int nvec = rows * (KC / 4), groups = rows >> 3;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    int f = threadIdx.x + i * NTHREADS, r, k4;
    if (f >= nvec) continue;
    chunk_item(f, r, k4);
    float4 v = c.v[i], hi, lo;
    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); lo.x = v.x - hi.x;
    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); lo.y = v.y - hi.y;
    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); lo.z = v.z - hi.z;
    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); lo.w = v.w - hi.w;
    int off = ((k4 * groups + (r >> 3)) * 32) + (r & 7) * 4;   // in floats: core matrix = 32 floats (128 B)
    *reinterpret_cast<float4*>(dst_hi + off) = hi;
    if (dst_lo) *reinterpret_cast<float4*>(dst_lo + off) = lo;
  }
}

constexpr int STAGE_FLOATS = 2 * TILE_M * KC + 2 * MAXH * KC;   // A hi/lo + W hi/lo of one K chunk

__global__ void __launch_bounds__(NTHREADS, 1) mlp_forward_kernel(MlpArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* stage0 = reinterpret_cast<float*>(smem);               // two operand stages: [A_hi | A_lo | W_hi | W_lo] each
  float* H = stage0 + NSTAGE * STAGE_FLOATS;                         // TILE_M x HSTRIDE activations
  __shared__ __align__(8) unsigned long long mbar_storage[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int net = blockIdx.y;
  const int row0 = blockIdx.x * TILE_M;
  const int valid = min(TILE_M, a.n_rows - row0);
  const uint32_t mbar[2] = {smem_u32(&mbar_storage[0]), smem_u32(&mbar_storage[1])};

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(mbar[0]));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(mbar[1]));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {  // TMEM: 256 fp32 accumulator columns, allocated and freed by warp 0
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  uint32_t parity = 0;    // bit st: phase stage st's mbarrier completes next
  uint32_t pending = 0;   // bit st: a commit was issued on stage st and not yet waited for
  bool ok = true;

  for (int layer = 0; layer < 3 && ok; layer++) {
    const int K = layer == 0 ? a.in_dim : a.hidden;
    const int nout = layer == 2 ? a.out_dim[net] : a.hidden;  // real output features
    const int N = layer == 2 ? 32 : a.hidden;                // UMMA N (last layer padded to 32)
    const float* W = a.w[net][layer];
    const float* Asrc = layer == 0 ? a.x : H;
    const int Ald = layer == 0 ? a.in_dim : HSTRIDE, Avalid = layer == 0 ? valid : TILE_M, Abase = layer == 0 ? row0 : 0;
    const uint32_t idesc = umma_idesc_tf32(TILE_M, N);
    const uint32_t lboA = (TILE_M / 8) * 128, lboW = (N / 8) * 128;
    // Register prefetch ring: the loads of chunk k + PF are issued as soon as chunk k's registers have been staged, so
    // an L2 / HBM round trip is covered by PF rounds of staging + MMA (one CTA per SM: registers are plentiful).
    ChunkRegs<TILE_M * (KC / 4) / NTHREADS> ra[PF];
    ChunkRegs<MAXH * (KC / 4) / NTHREADS> rw[PF];
#pragma unroll
    for (int p = 0; p < PF; p++)
      if (p * KC < K) {
        load_chunk(ra[p], Asrc, Ald, TILE_M, Avalid, p * KC, Abase);
        load_chunk(rw[p], W, K, N, nout, p * KC, 0);
      }
    for (int kb = 0; kb < K && ok; kb += PF * KC) {
#pragma unroll
      for (int p = 0; p < PF; p++) {
        const int k0 = kb + p * KC;
        if (k0 >= K || !ok) break;
        const int st = p % NSTAGE;   // compile-time after unrolling (PF is a multiple of NSTAGE)
        float* A_hi = stage0 + st * STAGE_FLOATS;
        float* A_lo = A_hi + TILE_M * KC;
        float* W_hi = A_lo + TILE_M * KC;
        float* W_lo = W_hi + MAXH * KC;
        if (pending & (1u << st)) {   // the MMAs that read this stage NSTAGE chunks ago must have finished before it is overwritten
          ok = mbar_wait(mbar[st], (parity >> st) & 1u) && ok;
          parity ^= 1u << st; pending &= ~(1u << st);
          if (!ok) break;
        }
        store_chunk(ra[p], A_hi, a.precise ? A_lo : nullptr, TILE_M);
        store_chunk(rw[p], W_hi, a.precise ? W_lo : nullptr, N);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy stores -> visible to the MMA
        __syncthreads();
        if (threadIdx.x == 0) {
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
          for (int ks = 0; ks < KC / 8; ks++) {
            uint64_t ah = umma_desc(smem_u32(A_hi) + ks * 2 * lboA, lboA, 128), wh = umma_desc(smem_u32(W_hi) + ks * 2 * lboW, lboW, 128);
            umma_tf32(tmem, ah, wh, idesc, (k0 | ks) != 0);
            if (a.precise) {
              uint64_t al = umma_desc(smem_u32(A_lo) + ks * 2 * lboA, lboA, 128), wl = umma_desc(smem_u32(W_lo) + ks * 2 * lboW, lboW, 128);
              umma_tf32(tmem, ah, wl, idesc, 1);
              umma_tf32(tmem, al, wh, idesc, 1);
            }
          }
          // commit: the mbarrier fires when every MMA issued so far has finished reading shared memory / writing TMEM
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar[st]) : "memory");
        }
        pending |= 1u << st;
        if (k0 + PF * KC < K) {
          load_chunk(ra[p], Asrc, Ald, TILE_M, Avalid, k0 + PF * KC, Abase);
          load_chunk(rw[p], W, K, N, nout, k0 + PF * KC, 0);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < NSTAGE; b++)   // drain: the accumulator is complete once every stage's commit has fired
      if (ok && (pending & (1u << b))) { ok = mbar_wait(mbar[b], (parity >> b) & 1u) && ok; parity ^= 1u << b; pending &= ~(1u << b); }
    if (!ok) break;
    // ---- epilogue: TMEM -> registers -> bias (+ReLU) -> shared activations / global output
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int quad = warp & 3, cgroup = warp >> 2;             // TMEM lane quadrant of the warp, column group
    const int r = quad * 32 + lane;
    const int ncols = max(16, N / (NTHREADS / 128));          // the warps of a quadrant split the columns between them
    const float* bias = a.b[net][layer];
    for (int c0 = cgroup * ncols; c0 < min(N, (cgroup + 1) * ncols); c0 += 16) {
      uint32_t v[16];
      uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      if (layer < 2) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          float4 o;
          o.x = fmaxf(__uint_as_float(v[4 * q + 0]) + __ldg(bias + c0 + 4 * q + 0), 0.f);
          o.y = fmaxf(__uint_as_float(v[4 * q + 1]) + __ldg(bias + c0 + 4 * q + 1), 0.f);
          o.z = fmaxf(__uint_as_float(v[4 * q + 2]) + __ldg(bias + c0 + 4 * q + 2), 0.f);
          o.w = fmaxf(__uint_as_float(v[4 * q + 3]) + __ldg(bias + c0 + 4 * q + 3), 0.f);
          *reinterpret_cast<float4*>(H + r * HSTRIDE + c0 + 4 * q) = o;
        }
      } else if (r < valid) {
#pragma unroll
        for (int q = 0; q < 16; q++)
          if (c0 + q < nout) a.y[net][(size_t)(row0 + r) * nout + c0 + q] = __uint_as_float(v[q]) + __ldg(bias + c0 + q);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();   // accumulator drained and H complete before the next layer overwrites TMEM / reads H
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }
  if (!ok && threadIdx.x == 0) atomicExch(a.error, 1);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tmem) : "memory");
}

// ------------------------------------------------------------------------------------------------ warp-specialised variant
// Same math and layouts as mlp_forward_kernel, but the K loop has no CTA barrier: warp 0 only issues tcgen05.mma, the
// other seven warps only stage operands, and they meet on mbarriers -
//   full[s]  (224 arrivals): the producers have written stage s and fenced it for the async proxy,
//   empty[s] (tcgen05.commit): the MMAs that read stage s have finished, the producers may overwrite it,
//   acc      (tcgen05.commit): the layer's accumulator is complete, everybody runs the TMEM epilogue.
// The CTA only synchronises once per layer (activations H complete / accumulator drained).
#ifndef B2H_MLP_WS_PF
#define B2H_MLP_WS_PF 3
#endif
constexpr int WS_PF = B2H_MLP_WS_PF;              // chunks a producer keeps in flight in registers
constexpr int WS_NS = 2;                          // operand stages
constexpr int WS_NPROD = NTHREADS - 32;           // producer threads (warps 1..7)
constexpr int WS_ITEMS = (TILE_M * (KC / 4) + MAXH * (KC / 4) + WS_NPROD - 1) / WS_NPROD;   // float4 per producer per chunk

__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(mbar) : "memory");
}

struct WsChunk { float4 v[WS_ITEMS]; };

// item f of a chunk: the first TILE_M * KC/4 items are rows of A, the rest rows of W (chunk_item order inside each)
__device__ __forceinline__ void ws_load(WsChunk& c, int pt, const float* Asrc, int Ald, int Avalid, int Abase, const float* W, int K,
                                        int N, int nout, int k0) {
  constexpr int NA = TILE_M * (KC / 4);
  const int nitems = NA + N * (KC / 4);
#pragma unroll
  for (int i = 0; i < WS_ITEMS; i++) {
    int f = pt + i * WS_NPROD, r, k4;
    c.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f < NA) {
      chunk_item(f, r, k4);
      if (r < Avalid) c.v[i] = *reinterpret_cast<const float4*>(Asrc + (size_t)(Abase + r) * Ald + k0 + 4 * k4);
    } else if (f < nitems) {
      chunk_item(f - NA, r, k4);
      if (r < nout) c.v[i] = *reinterpret_cast<const float4*>(W + (size_t)r * K + k0 + 4 * k4);
    }
  }
}
__device__ __forceinline__ void ws_store(const WsChunk& c, int pt, float* A_hi, float* A_lo, float* W_hi, float* W_lo, int N, bool precise) {
  constexpr int NA = TILE_M * (KC / 4);
  const int nitems = NA + N * (KC / 4);
#pragma unroll
  for (int i = 0; i < WS_ITEMS; i++) {
    int f = pt + i * WS_NPROD, r, k4;
    if (f >= nitems) continue;
    const bool isA = f < NA;
    chunk_item(isA ? f : f - NA, r, k4);
    const int groups = (isA ? TILE_M : N) >> 3;
    float4 v = c.v[i], hi, lo;
    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); lo.x = v.x - hi.x;
    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); lo.y = v.y - hi.y;
    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); lo.z = v.z - hi.z;
    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); lo.w = v.w - hi.w;
    int off = ((k4 * groups + (r >> 3)) * 32) + (r & 7) * 4;
    *reinterpret_cast<float4*>((isA ? A_hi : W_hi) + off) = hi;
    if (precise) *reinterpret_cast<float4*>((isA ? A_lo : W_lo) + off) = lo;
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) mlp_forward_ws_kernel(MlpArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* stage0 = reinterpret_cast<float*>(smem);
  float* H = stage0 + WS_NS * STAGE_FLOATS;
  __shared__ __align__(8) unsigned long long bar_storage[2 * WS_NS + 1];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int net = blockIdx.y;
  const int row0 = blockIdx.x * TILE_M;
  const int valid = min(TILE_M, a.n_rows - row0);
  uint32_t full[WS_NS], empty[WS_NS];
#pragma unroll
  for (int s = 0; s < WS_NS; s++) { full[s] = smem_u32(&bar_storage[s]); empty[s] = smem_u32(&bar_storage[WS_NS + s]); }
  const uint32_t accbar = smem_u32(&bar_storage[2 * WS_NS]);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < WS_NS; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(full[s]), "r"(WS_NPROD));
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
  int g = 0;   // chunks since the start of the kernel: stage g % WS_NS, use g / WS_NS (all roles count the same sequence)

  for (int layer = 0; layer < 3; layer++) {
    const int K = layer == 0 ? a.in_dim : a.hidden;
    const int nout = layer == 2 ? a.out_dim[net] : a.hidden;
    const int N = layer == 2 ? 32 : a.hidden;
    const float* W = a.w[net][layer];
    const float* Asrc = layer == 0 ? a.x : H;
    const int Ald = layer == 0 ? a.in_dim : HSTRIDE, Avalid = layer == 0 ? valid : TILE_M, Abase = layer == 0 ? row0 : 0;
    const int nchunk = K / KC;
    if (warp == 0) {
      if (lane == 0) {   // ---- MMA issuer
        const uint32_t idesc = umma_idesc_tf32(TILE_M, N);
        const uint32_t lboA = (TILE_M / 8) * 128, lboW = (N / 8) * 128;
        for (int c = 0; c < nchunk && ok; c++) {
          const int s = (g + c) % WS_NS, use = (g + c) / WS_NS;
          ok = mbar_wait(full[s], use & 1);
          if (!ok) break;
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint32_t A_hi = smem_u32(stage0 + s * STAGE_FLOATS), A_lo = A_hi + TILE_M * KC * 4, W_hi = A_lo + TILE_M * KC * 4,
                         W_lo = W_hi + MAXH * KC * 4;
#pragma unroll
          for (int ks = 0; ks < KC / 8; ks++) {
            uint64_t ah = umma_desc(A_hi + ks * 2 * lboA, lboA, 128), wh = umma_desc(W_hi + ks * 2 * lboW, lboW, 128);
            umma_tf32(tmem, ah, wh, idesc, (c | ks) != 0);
            if (a.precise) {
              uint64_t al = umma_desc(A_lo + ks * 2 * lboA, lboA, 128), wl = umma_desc(W_lo + ks * 2 * lboW, lboW, 128);
              umma_tf32(tmem, ah, wl, idesc, 1);
              umma_tf32(tmem, al, wh, idesc, 1);
            }
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(empty[s]) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(accbar) : "memory");
      }
      __syncwarp();      // lanes 1..31 park here instead of polling the accumulator barrier next to the issuing lane
    } else {             // ---- producers
      const int pt = threadIdx.x - 32;
      WsChunk regs[WS_PF];   // register prefetch ring: chunk c + WS_PF is requested as soon as chunk c has been staged
#pragma unroll
      for (int p = 0; p < WS_PF; p++)
        if (p < nchunk) ws_load(regs[p], pt, Asrc, Ald, Avalid, Abase, W, K, N, nout, p * KC);
      for (int cb = 0; cb < nchunk && ok; cb += WS_PF) {
#pragma unroll
        for (int p = 0; p < WS_PF; p++) {
          const int c = cb + p;
          if (c >= nchunk || !ok) break;
          const int s = (g + c) % WS_NS, use = (g + c) / WS_NS;
          if (use > 0) ok = mbar_wait(empty[s], (use - 1) & 1);
          if (!ok) break;
          float* A_hi = stage0 + s * STAGE_FLOATS;
          float* A_lo = A_hi + TILE_M * KC;
          float* W_hi = A_lo + TILE_M * KC;
          float* W_lo = W_hi + MAXH * KC;
          ws_store(regs[p], pt, A_hi, A_lo, W_hi, W_lo, N, a.precise != 0);
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
          mbar_arrive(full[s]);
          if (c + WS_PF < nchunk) ws_load(regs[p], pt, Asrc, Ald, Avalid, Abase, W, K, N, nout, (c + WS_PF) * KC);
        }
      }
    }
    g += nchunk;
    // ---- everybody: wait for the accumulator, epilogue
    ok = mbar_wait(accbar, layer & 1) && ok;
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (ok) {
      const int quad = warp & 3, cgroup = warp >> 2;
      const int r = quad * 32 + lane;
      const int ncols = max(16, N / (NTHREADS / 128));
      const float* bias = a.b[net][layer];
      for (int c0 = cgroup * ncols; c0 < min(N, (cgroup + 1) * ncols); c0 += 16) {
        uint32_t v[16];
        uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (layer < 2) {
#pragma unroll
          for (int q = 0; q < 4; q++) {
            float4 o;
            o.x = fmaxf(__uint_as_float(v[4 * q + 0]) + __ldg(bias + c0 + 4 * q + 0), 0.f);
            o.y = fmaxf(__uint_as_float(v[4 * q + 1]) + __ldg(bias + c0 + 4 * q + 1), 0.f);
            o.z = fmaxf(__uint_as_float(v[4 * q + 2]) + __ldg(bias + c0 + 4 * q + 2), 0.f);
            o.w = fmaxf(__uint_as_float(v[4 * q + 3]) + __ldg(bias + c0 + 4 * q + 3), 0.f);
            *reinterpret_cast<float4*>(H + r * HSTRIDE + c0 + 4 * q) = o;
          }
        } else if (r < valid) {
#pragma unroll
          for (int q = 0; q < 16; q++)
            if (c0 + q < nout) a.y[net][(size_t)(row0 + r) * nout + c0 + q] = __uint_as_float(v[q]) + __ldg(bias + c0 + q);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();   // accumulator drained and H complete before the next layer overwrites TMEM / reads H
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (!ok) break;
  }
  if (!ok && lane == 0) atomicExch(a.error, 1);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tmem) : "memory");
}

// ------------------------------------------------------------------------------------------------ packed-weight TMA pipeline (v2)
// The round-2 kernel.  What changed against mlp_forward_ws_kernel, and why (profiles/r01_mlp_kernel_tcgen05.txt: tensor
// pipe 25.6 % active, 64 us per 128-row tile, the time in 54 operand hand-offs):
//   * the weights are split into tf32 hi / lo and laid out in the canonical UMMA core-matrix order ONCE per policy update
//     (pack_weights_kernel), not by every CTA on every call; a K chunk of W then is one contiguous 32 KB block that the
//     TMA engine drops into shared memory (cp.async.bulk.tensor through a 2-D tensor map: rows = 128-byte core matrices)
//     while the threads do something else;
//   * K chunks are 32 wide (four MMA K-steps, 12 MMAs in precise mode): 11 + 8 + 8 hand-offs per tile instead of 54;
//   * there is no activation buffer: the epilogue of layer l reads the accumulator 32 columns at a time, applies bias +
//     ReLU, splits into hi / lo and writes the result straight into the operand ring as the next K chunk of layer l + 1,
//     whose MMAs (into the other half of TMEM) start while the rest of the epilogue is still running;
//   * roles: warp 0 issues tcgen05.mma, warp 1 issues the TMA loads, warps 4-7 (one thread per tile row, TMEM lane
//     quadrant = warp % 4) produce every A chunk -- from global memory for layer 0, from TMEM afterwards.
// Accumulators: layer 0 -> TMEM columns [0, 256), layer 1 -> [256, 512), layer 2 -> [0, 32).
// One operand ring of two stages; a stage holds a whole K chunk of 32 columns: A hi + lo (2 x 16 KB) and W hi + lo (2 x 32 KB),
// and ONE mbarrier says it is full (128 producer arrivals + the TMA thread's expect_tx bytes).  Measured alternatives
// (profiles/r02_mlp_v2_ring_variants.txt): separate A / W barriers +4 us per tile, a third A stage +8 us, W in half chunks
// through a five-deep ring +10 us -- every extra hand-off costs the single MMA-issuing thread more than the deeper
// prefetch returns; with a single tf32 pass instead of three the tile still takes 31 us, i.e. the tile is bound by ~27
// hand-off round trips of ~1 us, not by the tensor pipe.
constexpr int KC2 = 32;                                 // K columns of a chunk
constexpr int V2_NS = 2;                                // stages
constexpr int V2_A_PART = TILE_M * KC2;                 // floats of the hi (or lo) part of an A chunk: 16 KB
constexpr int V2_W_PART = MAXH * KC2;                   // ... of a W chunk: 32 KB
constexpr int V2_STAGE_FLOATS = 2 * V2_A_PART + 2 * V2_W_PART;
constexpr int V2_SMEM_FLOATS = V2_NS * V2_STAGE_FLOATS; // 192 KB
constexpr int V2_NPROD = 128;                           // A producers: warps 4..7

struct PackedNet {         // one trunk + head: offsets (floats) of the packed layers inside the buffer, chunk counts
  int chunks[3];           // K chunks of 32 per layer
  int n[3];                // UMMA N per layer (hidden, hidden, 32)
  int nout;                // real output features of the head
};
struct V2Args {
  const float* x;
  const float* bias[2][3];
  float* y[2];
  PackedNet net[2];
  int n_rows, in_dim, precise;
  int* error;
};
struct alignas(64) V2Maps { CUtensorMap m[2][3]; };

__device__ __forceinline__ void split_tf32(float4 v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); lo.x = v.x - hi.x;
  hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); lo.y = v.y - hi.y;
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); lo.z = v.z - hi.z;
  hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); lo.w = v.w - hi.w;
}

// W [nout, K] (nn.Linear layout) -> per K chunk of 32: [hi | lo] x [k4 group 0..7] x [row group 0..N/8) x 8 rows x 4 floats
struct PackJob { const float* W; float* dst; int nout, K, N, chunks; };
struct PackJobs { PackJob j[6]; };                       // both networks' three layers in one launch (blockIdx.y)
__global__ void pack_weights_kernel(PackJobs jobs) {
  const PackJob jb = jobs.j[blockIdx.y];
  const float* __restrict__ W = jb.W;
  float* __restrict__ dst = jb.dst;
  const int nout = jb.nout, K = jb.K, N = jb.N, halves = jb.chunks;
  constexpr int G = KC2 / 4;                             // 16-byte K groups per chunk
  const int total = halves * G * N;                      // float4 items per part
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ri = i & 7, rg = (i >> 3) % (N >> 3), k4 = (i / N) % G, c = i / (G * N);
    const int row = rg * 8 + ri, col = c * KC2 + k4 * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < nout) {
      const float* src = W + (size_t)row * K + col;
      if (col + 0 < K) v.x = src[0];
      if (col + 1 < K) v.y = src[1];
      if (col + 2 < K) v.z = src[2];
      if (col + 3 < K) v.w = src[3];
    }
    float4 hi, lo;
    split_tf32(v, hi, lo);
    const size_t part = (size_t)G * N * 4;               // floats per part of one chunk
    float* base = dst + (size_t)c * 2 * part + ((size_t)(k4 * (N >> 3) + rg) * 32 + ri * 4);
    *reinterpret_cast<float4*>(base) = hi;
    *reinterpret_cast<float4*>(base + part) = lo;
  }
}

__device__ __forceinline__ void v2_store_row_chunk(const float4 (&v)[8], int r, float* A_hi, float* A_lo, bool precise) {
#pragma unroll
  for (int k4 = 0; k4 < 8; k4++) {
    float4 hi, lo;
    split_tf32(v[k4], hi, lo);
    const int off = ((k4 * (TILE_M / 8) + (r >> 3)) * 32) + (r & 7) * 4;
    *reinterpret_cast<float4*>(A_hi + off) = hi;
    if (precise) *reinterpret_cast<float4*>(A_lo + off) = lo;
  }
}

__global__ void __launch_bounds__(256, 1) mlp_forward_v2_kernel(const __grid_constant__ V2Maps maps, V2Args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* stage0 = reinterpret_cast<float*>(smem);
  __shared__ __align__(8) unsigned long long bar_storage[2 * V2_NS + 3];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int net = blockIdx.y;
  const int row0 = blockIdx.x * TILE_M;
  const int valid = min(TILE_M, a.n_rows - row0);
  const PackedNet pn = a.net[net];
  const bool precise = a.precise != 0;
  uint32_t full[V2_NS], empty[V2_NS], acc[3];
#pragma unroll
  for (int s = 0; s < V2_NS; s++) { full[s] = smem_u32(&bar_storage[s]); empty[s] = smem_u32(&bar_storage[V2_NS + s]); }
#pragma unroll
  for (int l = 0; l < 3; l++) acc[l] = smem_u32(&bar_storage[2 * V2_NS + l]);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < V2_NS; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(full[s]), "r"(V2_NPROD + 1));   // 128 A producers + the TMA thread
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(empty[s]));
    }
#pragma unroll
    for (int l = 0; l < 3; l++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(acc[l]));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  bool ok = true;
  const int c0end = pn.chunks[0], c1end = c0end + pn.chunks[1], c2end = c1end + pn.chunks[2];   // global chunk sequence

  if (warp == 0) {
    if (lane == 0) {   // ---------------------------------------------------------------- MMA issuer
      for (int g = 0; g < c2end && ok; g++) {
        const int layer = g < c0end ? 0 : (g < c1end ? 1 : 2);
        const int c = g - (layer == 0 ? 0 : (layer == 1 ? c0end : c1end));
        const int N = pn.n[layer];
        const int s = g % V2_NS, use = g / V2_NS;
        ok = mbar_wait(full[s], use & 1);
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t idesc = umma_idesc_tf32(TILE_M, N);
        const uint32_t lboA = (TILE_M / 8) * 128, lboW = (uint32_t)(N / 8) * 128;
        const uint32_t A_hi = smem_u32(stage0 + s * V2_STAGE_FLOATS), A_lo = A_hi + V2_A_PART * 4, W_hi = A_lo + V2_A_PART * 4,
                       W_lo = W_hi + V2_W_PART * 4;
        const uint32_t d = tmem + (layer == 1 ? 256u : 0u);
#pragma unroll
        for (int ks = 0; ks < KC2 / 8; ks++) {
          uint64_t ah = umma_desc(A_hi + ks * 2 * lboA, lboA, 128), wh = umma_desc(W_hi + ks * 2 * lboW, lboW, 128);
          umma_tf32(d, ah, wh, idesc, (c | ks) != 0);
          if (precise) {
            uint64_t al = umma_desc(A_lo + ks * 2 * lboA, lboA, 128), wl = umma_desc(W_lo + ks * 2 * lboW, lboW, 128);
            umma_tf32(d, ah, wl, idesc, 1);
            umma_tf32(d, al, wh, idesc, 1);
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(empty[s]) : "memory");
        if (g + 1 == c0end || g + 1 == c1end || g + 1 == c2end)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(acc[layer]) : "memory");
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {   // ---------------------------------------------------------------- TMA: weight chunks
      for (int g = 0; g < c2end && ok; g++) {
        const int layer = g < c0end ? 0 : (g < c1end ? 1 : 2);
        const int c = g - (layer == 0 ? 0 : (layer == 1 ? c0end : c1end));
        const int N = pn.n[layer];
        const int s = g % V2_NS, use = g / V2_NS;
        if (use > 0) ok = mbar_wait(empty[s], (use - 1) & 1);
        if (!ok) break;
        const uint32_t part_bytes = (uint32_t)N * KC2 * 4;
        const uint32_t W_hi = smem_u32(stage0 + s * V2_STAGE_FLOATS + 2 * V2_A_PART), W_lo = W_hi + V2_W_PART * 4;
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(full[s]),
                     "r"(precise ? 2 * part_bytes : part_bytes) : "memory");
        const CUtensorMap* tm = &maps.m[net][layer];
        const int row_hi = c * 2 * N, row_lo = row_hi + N;      // rows of the map = 128-byte core matrices
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                     ::"r"(W_hi), "l"(tm), "r"(0), "r"(row_hi), "r"(full[s]) : "memory");
        if (precise)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                       ::"r"(W_lo), "l"(tm), "r"(0), "r"(row_lo), "r"(full[s]) : "memory");
      }
    }
    __syncwarp();
  } else if (warp >= 4) {   // ------------------------------------------------------------- A producers / epilogue (thread = tile row)
    const int r = threadIdx.x - 128;
    const int quad = warp & 3;
    int g = 0;
    // layer 0: x[row0 + r, :] -> chunks (zero rows beyond the batch, zero columns beyond in_dim)
    {
      const float* xr = a.x + (size_t)(row0 + (r < valid ? r : 0)) * a.in_dim;
      const bool vec = (a.in_dim & 3) == 0;
      float4 cur[8], nxt[8];
      auto load = [&](float4 (&v)[8], int c) {
#pragma unroll
        for (int k4 = 0; k4 < 8; k4++) {
          const int col = c * KC2 + 4 * k4;
          v[k4] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < valid) {
            if (vec && col + 3 < a.in_dim) v[k4] = *reinterpret_cast<const float4*>(xr + col);
            else {
              if (col + 0 < a.in_dim) v[k4].x = xr[col + 0];
              if (col + 1 < a.in_dim) v[k4].y = xr[col + 1];
              if (col + 2 < a.in_dim) v[k4].z = xr[col + 2];
              if (col + 3 < a.in_dim) v[k4].w = xr[col + 3];
            }
          }
        }
      };
      load(cur, 0);
      for (int c = 0; c < pn.chunks[0] && ok; c++, g++) {
        if (c + 1 < pn.chunks[0]) load(nxt, c + 1);
        const int s = g % V2_NS, use = g / V2_NS;
        if (use > 0) ok = mbar_wait(empty[s], (use - 1) & 1);
        if (!ok) break;
        float* A_hi = stage0 + s * V2_STAGE_FLOATS;
        v2_store_row_chunk(cur, r, A_hi, A_hi + V2_A_PART, precise);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        mbar_arrive(full[s]);
#pragma unroll
        for (int k4 = 0; k4 < 8; k4++) cur[k4] = nxt[k4];
      }
    }
    // layers 1 and 2: the previous layer's accumulator, 32 columns at a time -> bias + ReLU -> next K chunk
    for (int layer = 1; layer < 3 && ok; layer++) {
      ok = mbar_wait(acc[layer - 1], 0);
      if (!ok) break;
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const float* bias = a.bias[net][layer - 1];
      const uint32_t dsrc = tmem + ((uint32_t)(quad * 32) << 16) + (layer == 2 ? 256u : 0u);
      for (int c = 0; c < pn.chunks[layer] && ok; c++, g++) {
        uint32_t v[32];
#pragma unroll
        for (int h = 0; h < 2; h++)
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
              : "=r"(v[16 * h + 0]), "=r"(v[16 * h + 1]), "=r"(v[16 * h + 2]), "=r"(v[16 * h + 3]), "=r"(v[16 * h + 4]), "=r"(v[16 * h + 5]),
                "=r"(v[16 * h + 6]), "=r"(v[16 * h + 7]), "=r"(v[16 * h + 8]), "=r"(v[16 * h + 9]), "=r"(v[16 * h + 10]), "=r"(v[16 * h + 11]),
                "=r"(v[16 * h + 12]), "=r"(v[16 * h + 13]), "=r"(v[16 * h + 14]), "=r"(v[16 * h + 15])
              : "r"(dsrc + (uint32_t)(c * KC2 + 16 * h)));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        float4 o[8];
#pragma unroll
        for (int k4 = 0; k4 < 8; k4++) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c * KC2 + 4 * k4));
          o[k4].x = fmaxf(__uint_as_float(v[4 * k4 + 0]) + b4.x, 0.f);
          o[k4].y = fmaxf(__uint_as_float(v[4 * k4 + 1]) + b4.y, 0.f);
          o[k4].z = fmaxf(__uint_as_float(v[4 * k4 + 2]) + b4.z, 0.f);
          o[k4].w = fmaxf(__uint_as_float(v[4 * k4 + 3]) + b4.w, 0.f);
        }
        const int s = g % V2_NS, use = g / V2_NS;
        if (use > 0) ok = mbar_wait(empty[s], (use - 1) & 1);
        if (!ok) break;
        float* A_hi = stage0 + s * V2_STAGE_FLOATS;
        v2_store_row_chunk(o, r, A_hi, A_hi + V2_A_PART, precise);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");   // the TMEM reads above precede whoever this releases
        mbar_arrive(full[s]);
      }
    }
    // output head: accumulator columns [0, 32) + bias -> global
    if (ok) ok = mbar_wait(acc[2], 0);
    if (ok) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const float* bias = a.bias[net][2];
      const int nout = pn.nout;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        if (16 * h >= nout) break;
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(16 * h)));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (r < valid) {
#pragma unroll
          for (int q = 0; q < 16; q++)
            if (16 * h + q < nout) a.y[net][(size_t)(row0 + r) * nout + 16 * h + q] = __uint_as_float(v[q]) + __ldg(bias + 16 * h + q);
        }
      }
    }
  }
  if (!ok) atomicExch(a.error, 1);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

// ---- DiagGaussian sampling of SB3 (common/distributions.py): a = mean + exp(log_std) * eps, log_prob summed over
// action dims, clipped copy for the env (collect_rollouts clips to the Box bounds, the buffer keeps the raw action).
__device__ __forceinline__ void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
// Eight lanes per row, one lane per group of four action dimensions (one Philox block each): 4096 rows are 256 CTAs instead
// of the 32 a thread-per-row mapping gives (8.0 -> ~3 us); the log-probability is reduced over the row's lanes with shuffles.
__global__ void __launch_bounds__(128) policy_sample_kernel(const float* __restrict__ mean, const float* __restrict__ log_std, int n_rows,
                                     int act_dim, unsigned long long seed, unsigned long long step, int row_offset, int deterministic,
                                     float* __restrict__ actions, float* __restrict__ clipped, float* __restrict__ log_prob,
                                     const unsigned long long* __restrict__ step_dev) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = t >> 3, sub = t & 7;
  if (step_dev) step += *step_dev;   // counter kept on the device (graph-captured rollout loops)
  float lp = 0.f;
  if (row < n_rows) {
    for (int j0 = 4 * sub; j0 < act_dim; j0 += 32) {
      uint32_t c[4] = {(uint32_t)(row + row_offset), (uint32_t)step, (uint32_t)(step >> 32), (uint32_t)(j0 >> 2) ^ 0x504F4C49u};
      philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
      float eps[4];
      for (int h = 0; h < 2; h++) {  // Box-Muller, two normals per pair of words
        float u1 = ((c[2 * h] >> 8) + 1) * (1.0f / 16777216.0f), u2 = (c[2 * h + 1] >> 8) * (1.0f / 16777216.0f);
        float rad = sqrtf(-2.0f * logf(u1)), s, co;
        sincospif(2.0f * u2, &s, &co);
        eps[2 * h] = rad * co; eps[2 * h + 1] = rad * s;
      }
      for (int q = 0; q < 4 && j0 + q < act_dim; q++) {
        int j = j0 + q;
        float ls = log_std[j], m = mean[(size_t)row * act_dim + j];
        float e = deterministic ? 0.f : eps[q];
        float av = m + expf(ls) * e;
        actions[(size_t)row * act_dim + j] = av;
        clipped[(size_t)row * act_dim + j] = fminf(fmaxf(av, -1.f), 1.f);
        lp += -0.5f * e * e - ls - 0.91893853320467274f;   // -(a-mu)^2/(2 sigma^2) - log sigma - 0.5 log(2 pi)
      }
    }
  }
  lp += __shfl_xor_sync(0xffffffffu, lp, 4);
  lp += __shfl_xor_sync(0xffffffffu, lp, 2);
  lp += __shfl_xor_sync(0xffffffffu, lp, 1);
  if (row < n_rows && sub == 0) log_prob[row] = lp;
}

thread_local std::string g_err_mlp;

}  // namespace

extern "C" {

const char* b2h_mlp_last_error(void) { return g_err_mlp.c_str(); }

static int launch_mlp(MlpArgs& a, int nnets, void* stream) {
  if (a.n_rows <= 0 || a.in_dim <= 0 || a.in_dim % KC || a.hidden % KC || a.hidden < 16 || a.hidden > MAXH) {
    g_err_mlp = "unsupported MLP shape (in_dim and hidden must be multiples of 16, hidden <= 256)";
    return B2H_EUNSUPPORTED;
  }
  for (int n = 0; n < nnets; n++)
    if (a.out_dim[n] < 1 || a.out_dim[n] > 32) { g_err_mlp = "out_dim must be in [1, 32]"; return B2H_EUNSUPPORTED; }
  static int use_ws = -1;   // warp-specialised pipeline (default) or the barrier-per-chunk kernel (B2H_MLP_WS=0)
  if (use_ws < 0) { const char* e = getenv("B2H_MLP_WS"); use_ws = e ? atoi(e) != 0 : 1; }
  size_t smem = (size_t)((use_ws ? WS_NS : NSTAGE) * STAGE_FLOATS + TILE_M * HSTRIDE) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = use_ws ? cudaFuncSetAttribute(mlp_forward_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                           : cudaFuncSetAttribute(mlp_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { g_err_mlp = cudaGetErrorString(e); return B2H_ECUDA; }
    attr_set = true;
  }
  dim3 grid((a.n_rows + TILE_M - 1) / TILE_M, nnets);
  if (use_ws) mlp_forward_ws_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(a);
  else mlp_forward_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_err_mlp = cudaGetErrorString(e); return B2H_ECUDA; }
  return B2H_OK;
}

int b2h_mlp_forward(const float* x_dev, const float* w1_dev, const float* b1_dev, const float* w2_dev, const float* b2_dev,
                    const float* w3_dev, const float* b3_dev, float* y_dev, int n_rows, int in_dim, int hidden, int out_dim,
                    int precise, int* error_flag_dev, void* stream) {
  if (!x_dev || !w1_dev || !b1_dev || !w2_dev || !b2_dev || !w3_dev || !b3_dev || !y_dev || !error_flag_dev) { g_err_mlp = "null argument"; return B2H_EINVAL; }
  MlpArgs a = {};
  a.x = x_dev; a.w[0][0] = w1_dev; a.w[0][1] = w2_dev; a.w[0][2] = w3_dev; a.b[0][0] = b1_dev; a.b[0][1] = b2_dev; a.b[0][2] = b3_dev;
  a.y[0] = y_dev; a.out_dim[0] = out_dim; a.n_rows = n_rows; a.in_dim = in_dim; a.hidden = hidden; a.precise = precise; a.error = error_flag_dev;
  return launch_mlp(a, 1, stream);
}

int b2h_policy_forward(const float* x_dev, const float* const pi_dev[6], const float* const vf_dev[6], float* mean_dev, float* value_dev,
                       int n_rows, int in_dim, int hidden, int act_dim, int precise, int* error_flag_dev, void* stream) {
  if (!x_dev || !pi_dev || !vf_dev || !mean_dev || !value_dev || !error_flag_dev) { g_err_mlp = "null argument"; return B2H_EINVAL; }
  MlpArgs a = {};
  a.x = x_dev;
  for (int l = 0; l < 3; l++) {
    a.w[0][l] = pi_dev[2 * l]; a.b[0][l] = pi_dev[2 * l + 1]; a.w[1][l] = vf_dev[2 * l]; a.b[1][l] = vf_dev[2 * l + 1];
    if (!a.w[0][l] || !a.b[0][l] || !a.w[1][l] || !a.b[1][l]) { g_err_mlp = "null argument"; return B2H_EINVAL; }
  }
  a.y[0] = mean_dev; a.y[1] = value_dev; a.out_dim[0] = act_dim; a.out_dim[1] = 1;
  a.n_rows = n_rows; a.in_dim = in_dim; a.hidden = hidden; a.precise = precise; a.error = error_flag_dev;
  return launch_mlp(a, 2, stream);
}

// ---- packed-weight pipeline (mlp_forward_v2_kernel)
struct B2HPolicyPacked {
  int in_dim, hidden, act_dim;
  float* buf;
  size_t off[2][3];        // floats
  PackedNet net[2];
  V2Maps maps;
};

void b2h_policy_packed_destroy(B2HPolicyPacked* p) {
  if (!p) return;
  if (p->buf) cudaFree(p->buf);
  delete p;
}

int b2h_policy_packed_create(int in_dim, int hidden, int act_dim, B2HPolicyPacked** out) {
  if (!out || in_dim < 1 || hidden < 32 || hidden > MAXH || hidden % 32 || act_dim < 1 || act_dim > 32) {
    g_err_mlp = "unsupported MLP shape (hidden a multiple of 32 in [32, 256], act_dim <= 32)";
    return B2H_EUNSUPPORTED;
  }
  B2HPolicyPacked* p = new B2HPolicyPacked();
  p->in_dim = in_dim; p->hidden = hidden; p->act_dim = act_dim; p->buf = nullptr;
  size_t total = 0;
  for (int n = 0; n < 2; n++) {
    const int K[3] = {in_dim, hidden, hidden}, N[3] = {hidden, hidden, 32};
    for (int l = 0; l < 3; l++) {
      p->net[n].chunks[l] = (K[l] + KC2 - 1) / KC2;
      p->net[n].n[l] = N[l];
      p->off[n][l] = total;
      total += (size_t)p->net[n].chunks[l] * 2 * N[l] * KC2;      // hi + lo, K padded to whole chunks
    }
    p->net[n].nout = n == 0 ? act_dim : 1;
  }
  if (cudaMalloc(&p->buf, total * sizeof(float)) != cudaSuccess) { g_err_mlp = "cudaMalloc failed"; delete p; return B2H_ENOMEM; }
  // the driver entry point is looked up at run time: the library must load (and export its symbols) on a machine without
  // libcuda.so, e.g. the CPU-only build container
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeFn encode = nullptr;
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
      g_err_mlp = "cuTensorMapEncodeTiled is not available from this driver";
      b2h_policy_packed_destroy(p);
      return B2H_ECUDA;
    }
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  cudaMemset(p->buf, 0, total * sizeof(float));
  for (int n = 0; n < 2; n++)
    for (int l = 0; l < 3; l++) {
      // the packed layer as a 2-D tensor: rows = 128-byte core matrices (32 floats), one box = one part (hi or lo) of a K chunk
      const cuuint64_t dims[2] = {32, (cuuint64_t)p->net[n].chunks[l] * 2 * p->net[n].n[l]};
      const cuuint64_t strides[1] = {128};
      const cuuint32_t box[2] = {32, (cuuint32_t)p->net[n].n[l]};
      const cuuint32_t estr[2] = {1, 1};
      CUresult r = encode(&p->maps.m[n][l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p->buf + p->off[n][l], dims, strides, box, estr,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { g_err_mlp = "cuTensorMapEncodeTiled failed"; b2h_policy_packed_destroy(p); return B2H_ECUDA; }
    }
  const size_t smem = (size_t)V2_SMEM_FLOATS * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(mlp_forward_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { g_err_mlp = cudaGetErrorString(e); b2h_policy_packed_destroy(p); return B2H_ECUDA; }
  *out = p;
  return B2H_OK;
}

int b2h_policy_pack(B2HPolicyPacked* p, const float* const pi_dev[6], const float* const vf_dev[6], void* stream) {
  if (!p || !pi_dev || !vf_dev) { g_err_mlp = "null argument"; return B2H_EINVAL; }
  PackJobs jobs;
  int max_items = 0;
  for (int n = 0; n < 2; n++) {
    const float* const* w = n == 0 ? pi_dev : vf_dev;
    const int K[3] = {p->in_dim, p->hidden, p->hidden}, nout[3] = {p->hidden, p->hidden, p->net[n].nout};
    for (int l = 0; l < 3; l++) {
      if (!w[2 * l]) { g_err_mlp = "null argument"; return B2H_EINVAL; }
      PackJob& jb = jobs.j[3 * n + l];
      jb.W = w[2 * l]; jb.dst = p->buf + p->off[n][l]; jb.nout = nout[l]; jb.K = K[l]; jb.N = p->net[n].n[l]; jb.chunks = p->net[n].chunks[l];
      const int items = jb.chunks * (KC2 / 4) * jb.N;
      if (items > max_items) max_items = items;
    }
  }
  pack_weights_kernel<<<dim3((max_items + 255) / 256, 6), 256, 0, (cudaStream_t)stream>>>(jobs);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_err_mlp = cudaGetErrorString(e); return B2H_ECUDA; }
  return B2H_OK;
}

int b2h_policy_forward_packed(B2HPolicyPacked* p, const float* x_dev, const float* const pi_dev[6], const float* const vf_dev[6],
                              float* mean_dev, float* value_dev, int n_rows, int precise, int* error_flag_dev, void* stream) {
  if (!p || !x_dev || !pi_dev || !vf_dev || !error_flag_dev || n_rows <= 0 || (!mean_dev && !value_dev)) { g_err_mlp = "bad argument"; return B2H_EINVAL; }
  V2Args a = {};
  a.x = x_dev; a.n_rows = n_rows; a.in_dim = p->in_dim; a.precise = precise; a.error = error_flag_dev;
  // blockIdx.y walks the requested networks: both, or only the value trunk (predict_values) / only the policy trunk
  V2Maps maps;
  int nn = 0;
  for (int n = 0; n < 2; n++) {
    float* y = n == 0 ? mean_dev : value_dev;
    if (!y) continue;
    const float* const* w = n == 0 ? pi_dev : vf_dev;
    for (int l = 0; l < 3; l++) {
      if (!w[2 * l + 1]) { g_err_mlp = "null argument"; return B2H_EINVAL; }
      a.bias[nn][l] = w[2 * l + 1];
      maps.m[nn][l] = p->maps.m[n][l];
    }
    a.y[nn] = y; a.net[nn] = p->net[n];
    nn++;
  }
  const size_t smem = (size_t)V2_SMEM_FLOATS * sizeof(float);
  dim3 grid((n_rows + TILE_M - 1) / TILE_M, nn);
  mlp_forward_v2_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(maps, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_err_mlp = cudaGetErrorString(e); return B2H_ECUDA; }
  return B2H_OK;
}

int b2h_policy_sample(const float* mean_dev, const float* log_std_dev, int n_rows, int act_dim, uint64_t seed, uint64_t step,
                      int row_offset, int deterministic, float* actions_dev, float* clipped_dev, float* log_prob_dev, void* stream) {
  if (!mean_dev || !log_std_dev || !actions_dev || !clipped_dev || !log_prob_dev || n_rows <= 0 || act_dim <= 0) { g_err_mlp = "bad argument"; return B2H_EINVAL; }
  policy_sample_kernel<<<(8 * n_rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mean_dev, log_std_dev, n_rows, act_dim, seed, step,
                                                                                row_offset, deterministic, actions_dev, clipped_dev, log_prob_dev, nullptr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_err_mlp = cudaGetErrorString(e); return B2H_ECUDA; }
  return B2H_OK;
}

int b2h_policy_sample_dev(const float* mean_dev, const float* log_std_dev, int n_rows, int act_dim, uint64_t seed, const uint64_t* step_dev,
                          uint64_t step_offset, int row_offset, int deterministic, float* actions_dev, float* clipped_dev,
                          float* log_prob_dev, void* stream) {
  if (!mean_dev || !log_std_dev || !actions_dev || !clipped_dev || !log_prob_dev || !step_dev || n_rows <= 0 || act_dim <= 0) { g_err_mlp = "bad argument"; return B2H_EINVAL; }
  policy_sample_kernel<<<(8 * n_rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mean_dev, log_std_dev, n_rows, act_dim, seed, step_offset,
      row_offset, deterministic, actions_dev, clipped_dev, log_prob_dev, reinterpret_cast<const unsigned long long*>(step_dev));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_err_mlp = cudaGetErrorString(e); return B2H_ECUDA; }
  return B2H_OK;
}

}  // extern "C"
