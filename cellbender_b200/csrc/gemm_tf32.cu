// K3/K4: the large contractions of the encoder / decoder MLPs on the 5th-generation tensor cores.
//
// Replaces the cuBLAS SGEMMs behind nn.Linear in EncodeZ / EncodeNonZLatents / Decoder (reference
// vae/encoder.py:97-125, 211-228, 279-298; vae/decoder.py:41-47) and their autograd backward.
//
//   C[M, N] (=, +=)  A[M, K] * B[N, K]^T  (+ bias[N])        fp32 in, TF32 multiply, fp32 accumulate in TMEM
//
// Blackwell execution model (sm_100a): one warp issues TMA (cp.async.bulk.tensor) into a ring of
// 128B-swizzled shared-memory stages, one elected thread issues tcgen05.mma.kind::tf32 with the accumulator in
// TMEM (128 lanes x up to 512 columns), four warps drain TMEM with tcgen05.ld and write the result.
// mbarriers hand stages back and forth; tcgen05.commit releases a stage when its MMAs have been consumed.
//
// Operand layouts.  Both operands may be "K-major" (stored [rows = M or N, cols = K], the nn.Linear forward
// case) or "MN-major" (stored [rows = K, cols = M or N], which is what the weight-gradient products
// dW = dY^T X and the input-gradient products dX = dY W see when nothing is transposed in memory).  The TMA
// boxes are always 32 fp32 (128 B) wide in the contiguous dimension, so the two cases only differ in the
// shared-memory descriptor (canonical UMMA layouts ((8,m),(T,2)):((8T,SBO),(1,T)) vs ((T,8,m),(4,k)):((1,T,LBO),(8T,SBO)) with the
// 32-byte-atom 128B swizzle that tf32 MN-major operands require)
// and in the a_major / b_major bits of the instruction descriptor.  Out-of-bounds parts of a box (K tail, the
// 255-row minibatch) are zero-filled by TMA, so no padding copies are needed.
//
// Skinny shapes (M = 255 with K = 33 538) are split along K across CTAs; partial tiles go to a workspace with
// plain stores and a second kernel sums them in a fixed order (deterministic; no floating-point atomics).
#include <cstdlib>
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"

namespace cb {

constexpr int kGemmThreads = 192;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..5: epilogue
constexpr int kTilePitch = 132;    // floats per staged gradient-tile row (128 + 4: 16-byte rows, conflict-free float4 phases)
constexpr int kBlockM = 128;
constexpr int kBlockK = 32;  // 32 fp32 = 128 B = one swizzle row
constexpr int kUmmaK = 8;    // tf32: 32 B per MMA along K

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// the same box delivered to the same shared-memory offset (and mbarrier) of every CTA of the cluster named in cta_mask
__device__ __forceinline__ void tma_load_2d_multicast(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "h"(cta_mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once all previously issued tcgen05 ops of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 64-bit shared-memory matrix descriptor, descriptor version 1 (Blackwell).
// layout_type 2 = SWIZZLE_128B (K-major operands); 1 = SWIZZLE_128B_BASE32B, the only layout the tensor core
// accepts for MN-major 32-bit (tf32) operands: 128 B rows, 32 B chunks XOR-ed with (row & 3), atoms of 4 K-rows.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(layout_type) << 61;
  return d;
}

struct GemmParams {
  float* C;              // direct output (split_k == 1) [M, ldc]
  float* workspace;      // partial tiles (split_k > 1)  [split_k][Mpad][Npad]
  const float* bias;     // [N] or nullptr (direct path only; the reduce kernel adds it on the split path)
  long long ldc;
  int M, N, K;
  int k_blocks_total;    // ceil(K / 32) (+ ceil(K2 / 32) with a second operand pair)
  int k_blocks_first;    // k-blocks of the first operand pair
  int k_blocks_per_split;
  int split_k;
  int accumulate;        // C += ... (direct path)
  long long Mpad, Npad;
  // optional fused optimizer epilogue (weight-gradient products only): instead of storing the gradient tile, apply the
  // ClippedAdam update (csrc/adam.cu) to the parameter / moment tiles that share C's indexing -- the 68.7 MB gradient
  // of a G x 512 layer is then never written to, nor read back from, HBM.
  float* adam_p;         // nullptr = plain store
  float* adam_m;
  float* adam_v;
  const float* lr_table;
  const float* beta1_table;
  long long table_len;
  const long long* step_ptr;
  float beta2, eps, clip;
};

struct AdamCoef {
  float b1, omb1, b2, omb2, step_size, eps, clip;
};
__device__ __forceinline__ AdamCoef adam_coef(const GemmParams& p) {
  const long long t0 = *p.step_ptr;
  const long long ti = t0 < p.table_len ? t0 : p.table_len - 1;
  const float lr = p.lr_table[ti], b1 = p.beta1_table[ti];
  const double t = double(t0 + 1);
  AdamCoef c;
  c.b1 = b1, c.omb1 = 1.0f - b1, c.b2 = p.beta2, c.omb2 = 1.0f - p.beta2;
  c.step_size = float(double(lr) * sqrt(1.0 - pow(double(p.beta2), t)) / (1.0 - pow(double(b1), t)));
  c.eps = p.eps, c.clip = p.clip;
  return c;
}
// the arithmetic of clipped_adam_kernel (csrc/adam.cu), element for element
__device__ __forceinline__ void adam_update(float g, float& pw, float& m, float& v, const AdamCoef& c) {
  const float gc = fminf(fmaxf(g, -c.clip), c.clip);
  m = c.b1 * m + c.omb1 * gc;
  v = c.b2 * v + c.omb2 * gc * gc;
  pw -= c.step_size * m / (sqrtf(v) + c.eps);
}

// MT = accumulators of 128 rows per CTA, BN = tile width; MT * BN <= 512 TMEM columns.
// MINB = CTAs per SM the launch is sized for: short-K products (weight gradients: K = 255 droplets; the hidden layers)
// run two CTAs per SM with a 2-stage ring (2 x 256 TMEM columns, 2 x 97 KB of shared memory), so one CTA's epilogue
// overlaps the other's loads -- each CTA lives for only ~8 k-blocks.
// CLUSTER = 4: the four CTAs that compute the four 128-wide column tiles of the same rows form a thread-block cluster and
// share the A (activation) tile: each CTA fetches one quarter of it and TMA multicasts that quarter into the shared
// memory of all four, so the activation matrix crosses the L2 -> SM fabric once instead of four times (the skinny
// forward products are bound by exactly that traffic).  A stage is released to the producers of all four CTAs by a
// multicast tcgen05.commit; empty barriers therefore count four arrivals.
template <int MT, int BN, bool A_MN, bool B_MN, int STAGES, int MINB, int CLUSTER = 1>
__global__ void __launch_bounds__(kGemmThreads, MINB)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_a2, const __grid_constant__ CUtensorMap tmap_b2, GemmParams p) {
  constexpr uint32_t kABytes = MT * kBlockM * kBlockK * 4;  // per stage
  constexpr uint32_t kBBytes = BN * kBlockK * 4;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = (MT * BN <= 32) ? 32 : (MT * BN <= 64) ? 64 : (MT * BN <= 128) ? 128 : (MT * BN <= 256) ? 256 : 512;
  static_assert(MT * BN <= 512, "accumulators exceed TMEM");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * (MT * kBlockM);
  const int split = blockIdx.z;
  const int kb_begin = split * p.k_blocks_per_split;
  const int kb_end = min(p.k_blocks_total, kb_begin + p.k_blocks_per_split);
  const int n_kb = max(kb_end - kb_begin, 0);

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CLUSTER);
    }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  static_assert(CLUSTER == 1 || (!A_MN && (MT * kBlockM) % CLUSTER == 0), "multicast is written for K-major A tiles");
  constexpr uint16_t kClusterMask = uint16_t((1u << CLUSTER) - 1);
  uint32_t cta_rank = 0;
  if constexpr (CLUSTER > 1) {
    cta_rank = cluster_cta_rank();
    cluster_sync_all();  // every CTA's barriers are initialised before any peer multicasts into them
  }

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      for (int i = 0; i < n_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = smem + s * kStageBytes;
        uint8_t* b_dst = a_dst + kABytes;
        mbar_expect_tx(&full_bar[s], kStageBytes);
        // k-blocks [0, k_blocks_first) come from operand pair 1, the rest from pair 2 (C = A1 B1^T + A2 B2^T)
        const int kb = kb_begin + i;
        const bool second = kb >= p.k_blocks_first;
        const CUtensorMap* ma = second ? &tmap_a2 : &tmap_a;
        const CUtensorMap* mb = second ? &tmap_b2 : &tmap_b;
        const int k0 = (second ? kb - p.k_blocks_first : kb) * kBlockK;
        if constexpr (CLUSTER > 1) {
          // this CTA's quarter of the A tile (rows cta_rank * 64 ..), delivered to all CTAs of the cluster
          constexpr int kRows = MT * kBlockM / CLUSTER;
          tma_load_2d_multicast(ma, &full_bar[s], a_dst + cta_rank * (kRows * kBlockK * 4), k0, m0 + int(cta_rank) * kRows,
                                kClusterMask);
        } else if constexpr (!A_MN) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) tma_load_2d(ma, &full_bar[s], a_dst + mt * 16384, k0, m0 + mt * kBlockM);
        } else {
#pragma unroll
          for (int j = 0; j < MT * 4; ++j) tma_load_2d(ma, &full_bar[s], a_dst + j * 4096, m0 + j * 32, k0);
        }
        if (!B_MN) {
#pragma unroll
          for (int j = 0; j < BN / 128; ++j) tma_load_2d(mb, &full_bar[s], b_dst + j * 16384, k0, n0 + j * 128);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) tma_load_2d(mb, &full_bar[s], b_dst + j * 4096, n0 + j * 32, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B tf32, majors, N >> 3, M >> 4
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(A_MN) << 15) | (uint32_t(B_MN) << 16) |
                             (uint32_t(BN >> 3) << 17) | (uint32_t(kBlockM >> 4) << 24);
      for (int i = 0; i < n_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + s * kStageBytes);
        const uint32_t b_base = a_base + kABytes;
#pragma unroll
        for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4) {
          const uint64_t bdesc = B_MN ? make_smem_desc(b_base + k4 * 1024, 4096, 512, 1) : make_smem_desc(b_base + k4 * 32, 16, 1024, 2);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t adesc = A_MN ? make_smem_desc(a_base + mt * 16384 + k4 * 1024, 4096, 512, 1)
                                        : make_smem_desc(a_base + mt * 16384 + k4 * 32, 16, 1024, 2);
            umma_tf32(tmem_base + mt * BN, adesc, bdesc, idesc, (i > 0 || k4 > 0) ? 1u : 0u);
          }
        }
        if (CLUSTER > 1) umma_commit_multicast(&empty_bar[s], kClusterMask);  // every peer writes into this stage
        else umma_commit(&empty_bar[s]);  // stage free once these MMAs have read it
      }
      umma_commit(&accum_bar);  // accumulator complete
    }
  } else {
    // ======================= epilogue: TMEM -> registers (-> shared) -> global =======================
    // tcgen05.ld hands every thread one accumulator ROW (32 consecutive columns).
    //  * plain output: each thread stores its 128 contiguous bytes (8 x st.v4, fire-and-forget);
    //  * split-K partial tiles and the fused optimizer update: every warp first transposes its 32 x 32 sub-tile through
    //    a private staging buffer (carved out of the now idle operand ring: all MMAs have completed), after which one
    //    warp instruction covers 4 rows x 128 contiguous bytes -- the read-modify-write of param / exp_avg / exp_avg_sq
    //    then moves whole 128-byte lines, and all its loads are issued before the first store.
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const bool to_ws = p.split_k > 1;
    const bool staged = to_ws;
    constexpr int kStg = 36;  // staging row pitch in floats: 16-byte aligned rows, conflict-free float4 phases
    float* stg = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * kStg);
    const int cc = 4 * (lane & 7), rsub = lane >> 3;
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        float v[32];
        if (n_kb > 0) {
          tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(mt * BN + c), v);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        if (!staged) {
          const int row = m0 + mt * kBlockM + q * 32 + lane;
          const int col = n0 + c;
          if (row < p.M && col < p.N) {
            float* dst = p.C + (long long)row * p.ldc + col;
            const int nv = min(32, p.N - col);
            if (nv == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                if (p.bias != nullptr) {
                  const float4 b = *reinterpret_cast<const float4*>(p.bias + col + i);
                  o.x += b.x, o.y += b.y, o.z += b.z, o.w += b.w;
                }
                if (p.accumulate) {
                  const float4 old = *reinterpret_cast<const float4*>(dst + i);
                  o.x += old.x, o.y += old.y, o.z += old.z, o.w += old.w;
                }
                *reinterpret_cast<float4*>(dst + i) = o;
              }
            } else {
              for (int i = 0; i < nv; ++i) {
                float o = v[i];
                if (p.bias != nullptr) o += p.bias[col + i];
                if (p.accumulate) o += dst[i];
                dst[i] = o;
              }
            }
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(stg + lane * kStg + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const int col = n0 + c + cc;
        const int row0 = m0 + mt * kBlockM + q * 32 + rsub;  // this lane's rows: row0 + 4 rr
        if (to_ws) {
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            const float4 o = *reinterpret_cast<const float4*>(stg + (rr * 4 + rsub) * kStg + cc);
            *reinterpret_cast<float4*>(p.workspace + ((long long)split * p.Mpad + row0 + 4 * rr) * p.Npad + col) = o;
          }
        }
        __syncwarp();  // the staging buffer is rewritten by the next chunk
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if constexpr (CLUSTER > 1) cluster_sync_all();  // no CTA leaves while peers may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Persistent weight-gradient GEMM with the ClippedAdam sweep fused behind it (cb_gemm_tf32_dw_adam).
//   G[M, N] = A^T B for MN-major operands (A = dY stored [K, M], B = X stored [K, N], K = minibatch rows <= 512),
//   then  param / exp_avg / exp_avg_sq [M, ldc]  <-  ClippedAdam(G)   -- G never reaches HBM.
// One CTA per SM loops over 256 x 128 output tiles.  The accumulators are DOUBLE-BUFFERED in TMEM (2 x 256 columns):
// while the twelve epilogue warps drain tile i (TMEM -> shared-memory staging tile) and sweep it over the optimizer
// state (HBM-bound: 24 bytes moved per element), the TMA producer and the MMA issuer already run tile i + 1
// (L2-bound) -- the two phases that a one-tile-per-CTA kernel executes back to back overlap here.
//   warp 0: TMA producer   warp 1: MMA issuer   warps 2..5: TMEM drain + sweep   warps 6..13: sweep
constexpr int kPersistThreads = 448;
constexpr int kPersistStages = 3;
constexpr int kPersistEpiThreads = kPersistThreads - 64;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kPersistThreads, 1)
gemm_dw_adam_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, GemmParams p) {
  constexpr int MT = 2, BN = 128, STAGES = kPersistStages;
  constexpr uint32_t kABytes = MT * kBlockM * kBlockK * 4;
  constexpr uint32_t kBBytes = BN * kBlockK * 4;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kAccCols = MT * BN;  // TMEM columns of one accumulator buffer

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* tile = reinterpret_cast<float*>(smem + STAGES * kStageBytes);  // [128][kTilePitch] gradient staging tile
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (p.N + BN - 1) / BN, tiles_m = (p.M + MT * kBlockM - 1) / (MT * kBlockM);
  const int n_tiles = tiles_n * tiles_m;
  const int n_kb = p.k_blocks_total;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_full[0], 1), mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 4), mbar_init(&acc_empty[1], 4);  // one arrival per drain warp
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 2 * kAccCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ======================= TMA producer =======================
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m0 = (t / tiles_n) * (MT * kBlockM), n0 = (t % tiles_n) * BN;
      if (lane == 0) {
        for (int kb = 0; kb < n_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* a_dst = smem + s * kStageBytes;
          uint8_t* b_dst = a_dst + kABytes;
          mbar_expect_tx(&full_bar[s], kStageBytes);
          const int k0 = kb * kBlockK;
#pragma unroll
          for (int j = 0; j < MT * 4; ++j) tma_load_2d(&tmap_a, &full_bar[s], a_dst + j * 4096, m0 + j * 32, k0);
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) tma_load_2d(&tmap_b, &full_bar[s], b_dst + j * 4096, n0 + j * 32, k0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | (uint32_t(BN >> 3) << 17) |
                             (uint32_t(kBlockM >> 4) << 24);
      int it = 0, ti = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
        const int buf = ti & 1;
        mbar_wait(&acc_empty[buf], ((ti >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator buffer
        tc_fence_after();
        for (int kb = 0; kb < n_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + s * kStageBytes);
          const uint32_t b_base = a_base + kABytes;
#pragma unroll
          for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4) {
            const uint64_t bdesc = make_smem_desc(b_base + k4 * 1024, 4096, 512, 1);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint64_t adesc = make_smem_desc(a_base + mt * 16384 + k4 * 1024, 4096, 512, 1);
              umma_tf32(tmem_base + buf * kAccCols + mt * BN, adesc, bdesc, idesc, (kb > 0 || k4 > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ======================= epilogue: drain + optimizer sweep (warps 2..13) =======================
    const int ew = warp - 2, n_ew = kPersistEpiThreads / 32;
    const bool drain = ew < 4;
    const int q = warp & 3;  // TMEM lane quarter of a drain warp (warps 2,3,4,5 -> 2,3,0,1)
    const AdamCoef ac = adam_coef(p);
    int ti = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
      const int m0 = (t / tiles_n) * (MT * kBlockM), n0 = (t % tiles_n) * BN;
      const int buf = ti & 1;
      const int col = n0 + 4 * lane;
      const bool vec = (col + 3 < p.N) && ((p.ldc & 3) == 0) &&
                       (((reinterpret_cast<uintptr_t>(p.adam_p) | reinterpret_cast<uintptr_t>(p.adam_m) |
                          reinterpret_cast<uintptr_t>(p.adam_v)) & 15) == 0);
      if (drain) {
        mbar_wait(&acc_full[buf], (ti >> 1) & 1);
        tc_fence_after();
      }
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        if (drain) {
#pragma unroll 1
          for (int c = 0; c < BN; c += 32) {
            float v[32];
            tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(buf * kAccCols + mt * BN + c), v);
            float* dst = tile + (q * 32 + lane) * kTilePitch + c;
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (mt == MT - 1) {  // both halves of this accumulator buffer are in registers / shared memory: release it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
          }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(kPersistEpiThreads) : "memory");  // staging tile written
        const int row_base = m0 + mt * kBlockM;
#pragma unroll 1
        for (int r0 = ew; r0 < kBlockM; r0 += 4 * n_ew) {
          if (vec) {
            float4 w4[4], m4[4], v4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {  // all loads first
              const int r = r0 + u * n_ew;
              if (r < kBlockM && row_base + r < p.M) {
                const long long off = (long long)(row_base + r) * p.ldc + col;
                w4[u] = *reinterpret_cast<const float4*>(p.adam_p + off);
                m4[u] = *reinterpret_cast<const float4*>(p.adam_m + off);
                v4[u] = *reinterpret_cast<const float4*>(p.adam_v + off);
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = r0 + u * n_ew;
              if (r < kBlockM && row_base + r < p.M) {
                const float4 g4 = *reinterpret_cast<const float4*>(tile + r * kTilePitch + 4 * lane);
                adam_update(g4.x, w4[u].x, m4[u].x, v4[u].x, ac);
                adam_update(g4.y, w4[u].y, m4[u].y, v4[u].y, ac);
                adam_update(g4.z, w4[u].z, m4[u].z, v4[u].z, ac);
                adam_update(g4.w, w4[u].w, m4[u].w, v4[u].w, ac);
                const long long off = (long long)(row_base + r) * p.ldc + col;
                *reinterpret_cast<float4*>(p.adam_p + off) = w4[u];
                *reinterpret_cast<float4*>(p.adam_m + off) = m4[u];
                *reinterpret_cast<float4*>(p.adam_v + off) = v4[u];
              }
            }
          } else if (col < p.N) {  // ragged right edge / unaligned buffers
            const int nv = min(4, p.N - col);
            for (int u = 0; u < 4; ++u) {
              const int r = r0 + u * n_ew;
              if (r >= kBlockM || row_base + r >= p.M) continue;
              const long long off = (long long)(row_base + r) * p.ldc + col;
              for (int i = 0; i < nv; ++i)
                adam_update(tile[r * kTilePitch + 4 * lane + i], p.adam_p[off + i], p.adam_m[off + i], p.adam_v[off + i], ac);
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(kPersistEpiThreads) : "memory");  // staging tile consumed
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kAccCols);
  }
}

// out[m, n] (=, +=) sum_s ws[s][m][n] + bias[n]   -- fixed summation order
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int split_k, long long Mpad, long long Npad,
                                                            float* __restrict__ C, long long ldc, int M, int N,
                                                            const float* __restrict__ bias, int accumulate) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)M * N;
  if (idx >= total) return;
  const int m = int(idx / N), n = int(idx % N);
  float acc = 0.f;
  for (int s = 0; s < split_k; ++s) acc += ws[((long long)s * Mpad + m) * Npad + n];
  if (bias != nullptr) acc += bias[n];
  float* dst = C + (long long)m * ldc + n;
  *dst = accumulate ? (*dst + acc) : acc;
}

// ------------------------------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// 2-D fp32 tensor [rows, cols] row-major with leading dimension ld; box = 32 cols x box_rows, 128B swizzle, zero OOB fill
static int make_tmap(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_rows,
                     bool mn_major) {
  auto enc = get_encode();
  if (enc == nullptr) return 1;
  cuuint64_t gdim[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  cuuint64_t gstride[1] = {cuuint64_t(ld) * 4};
  cuuint32_t box[2] = {32u, cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

template <int MT, int BN, bool A_MN, bool B_MN, int STAGES, int MINB>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const CUtensorMap& tb2,
                      const GemmParams& p, cudaStream_t st) {
  constexpr int kStage = (MT * kBlockM + BN) * kBlockK * 4;
  const int smem = STAGES * kStage + 1024;
  auto kern = gemm_tf32_kernel<MT, BN, A_MN, B_MN, STAGES, MINB>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 3;
    configured = true;
  }
  dim3 grid((p.N + BN - 1) / BN, (p.M + MT * kBlockM - 1) / (MT * kBlockM), p.split_k);
  kern<<<grid, kGemmThreads, smem, st>>>(ta, tb, ta2, tb2, p);
  return 0;
}

// cluster-of-4 variant (A tile multicast): the split-K forward shape, grid.x == 4
template <bool B_MN>
static int launch_cluster4(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const CUtensorMap& tb2,
                           const GemmParams& p, cudaStream_t st) {
  constexpr int MT = 2, BN = 128, STAGES = 4;
  constexpr int kStage = (MT * kBlockM + BN) * kBlockK * 4;
  const int smem = STAGES * kStage + 1024;
  auto kern = gemm_tf32_kernel<MT, BN, false, B_MN, STAGES, 1, 4>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 3;
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((p.N + BN - 1) / BN, (p.M + MT * kBlockM - 1) / (MT * kBlockM), p.split_k);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, ta, tb, ta2, tb2, p) == cudaSuccess ? 0 : 4;
}

template <int MT, int BN, bool A_MN, bool B_MN>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const CUtensorMap& tb2,
                  const GemmParams& p, cudaStream_t st) {
  constexpr int kStage = (MT * kBlockM + BN) * kBlockK * 4;
  constexpr int kDeep = (kStage <= 48 * 1024) ? 4 : 3;
  if (p.k_blocks_per_split <= 16 && MT * BN <= 256)  // short K: two co-resident CTAs per SM, 2-stage ring each
    return launch_cfg<MT, BN, A_MN, B_MN, 2, 2>(ta, tb, ta2, tb2, p, st);
  return launch_cfg<MT, BN, A_MN, B_MN, kDeep, 1>(ta, tb, ta2, tb2, p, st);
}

}  // namespace cb

// workspace elements needed for a given problem (0 when split_k == 1)
extern "C" long long cb_gemm_tf32_workspace_elems(int M, int N, int split_k) {
  if (split_k <= 1) return 0;
  const long long Mpad = (M + 255) / 256 * 256, Npad = (N + 255) / 256 * 256;
  return (long long)split_k * Mpad * Npad;
}

// Suggested split along K so that about one CTA per SM streams the operands (skinny-M products).
extern "C" int cb_gemm_tf32_suggest_split(int M, int N, int K) {
  const int mt = (M > 128) ? 2 : 1;
  const int bn = (N >= 256 && mt == 1) ? 256 : 128;
  const long long tiles = (long long)((M + mt * 128 - 1) / (mt * 128)) * ((N + bn - 1) / bn);
  const int kb = (K + 31) / 32;
  if (tiles >= 148 || kb < 16) return 1;
  long long s = 148 / tiles;
  if (s > kb / 4) s = kb / 4;
  return int(s < 1 ? 1 : s);
}

namespace cb {
struct AdamArgs {
  float *p, *m, *v;
  const float *lr_table, *beta1_table;
  long long table_len;
  const long long* step_ptr;
  float beta2, eps, clip;
};
}  // namespace cb

static int gemm_tf32_impl(const float* A, long long lda, const float* B, long long ldb, int K, const float* A2,
                          long long lda2, const float* B2, long long ldb2, int K2, int a_mn, int b_mn, float* C,
                          long long ldc, int M, int N, const float* bias, int accumulate, int split_k,
                          float* workspace, long long workspace_elems, const cb::AdamArgs* adam, void* stream) {
  using namespace cb;
  CB_REQUIRE(M > 0 && N > 0 && K > 0 && K2 >= 0, "cb_gemm_tf32: bad shape M=%d N=%d K=%d K2=%d", M, N, K, K2);
  CB_REQUIRE((lda & 3) == 0 && (ldb & 3) == 0 && (lda2 & 3) == 0 && (ldb2 & 3) == 0,
             "cb_gemm_tf32: leading dimensions must be multiples of 4 (TMA needs 16-byte rows)");
  CB_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(A2) |
               reinterpret_cast<uintptr_t>(B2)) & 15) == 0, "cb_gemm_tf32: operands must be 16-byte aligned");
  if (split_k < 1) split_k = 1;
  const int kb_first = (K + kBlockK - 1) / kBlockK;
  const int kb_total = kb_first + (K2 > 0 ? (K2 + kBlockK - 1) / kBlockK : 0);
  if (split_k > kb_total) split_k = kb_total;
  GemmParams p;
  p.C = C, p.workspace = workspace, p.bias = bias, p.ldc = ldc, p.M = M, p.N = N, p.K = K;
  p.k_blocks_total = kb_total;
  p.k_blocks_first = kb_first;
  p.k_blocks_per_split = (kb_total + split_k - 1) / split_k;
  split_k = (kb_total + p.k_blocks_per_split - 1) / p.k_blocks_per_split;  // every split owns >= 1 k-block
  p.split_k = split_k;
  p.accumulate = accumulate;
  p.adam_p = p.adam_m = p.adam_v = nullptr;
  p.lr_table = p.beta1_table = nullptr, p.table_len = 0, p.step_ptr = nullptr, p.beta2 = p.eps = p.clip = 0.f;
  if (adam != nullptr) {
    CB_REQUIRE(split_k <= 1 && bias == nullptr && !accumulate,
               "cb_gemm_tf32_dw_adam: the fused optimizer epilogue needs split_k = 1, no bias, no accumulation");
    p.adam_p = adam->p, p.adam_m = adam->m, p.adam_v = adam->v, p.lr_table = adam->lr_table, p.beta1_table = adam->beta1_table;
    p.table_len = adam->table_len, p.step_ptr = adam->step_ptr, p.beta2 = adam->beta2, p.eps = adam->eps, p.clip = adam->clip;
  }
  p.Mpad = (M + 255) / 256 * 256, p.Npad = (N + 255) / 256 * 256;
  if (split_k > 1) {
    CB_REQUIRE(workspace != nullptr && workspace_elems >= (long long)split_k * p.Mpad * p.Npad,
               "cb_gemm_tf32: workspace too small (%lld < %lld)", workspace_elems, (long long)split_k * p.Mpad * p.Npad);
    p.bias = nullptr;
  }
  CUtensorMap ta, tb, ta2, tb2;
  // Skinny split-K products with four column tiles (the 512-wide encoder layers and the decoder's dh) can run as
  // clusters of four CTAs with the A tile multicast (each CTA fetches a 64-row quarter).  Measured on B200 this is
  // SLOWER for the training step (1.35-1.38 ms against 1.33 ms per step: the products are bound by HBM latency per
  // CTA, not by L2 -> SM traffic, and 37 four-CTA clusters place worse than 148 free CTAs), so it is opt-in:
  // CB_GEMM_CLUSTER=1.  Read per call so tests can switch it.
  const char* cluster_env = getenv("CB_GEMM_CLUSTER");
  const bool cluster_enabled = cluster_env != nullptr && cluster_env[0] == '1';
  const bool cluster4 = cluster_enabled && adam == nullptr && !a_mn && M > 128 && M <= 256 && split_k > 1 && (N + 127) / 128 == 4 &&
                        p.k_blocks_per_split > 16;
  const int a_box = cluster4 ? 64 : 128;
  // K-major operand: tensor [rows = M|N, cols = K], box 32 x 128 rows.  MN-major: tensor [rows = K, cols = M|N], box 32 x 32 rows.
  int e = a_mn ? make_tmap(&ta, A, K, M, lda, 32, true) : make_tmap(&ta, A, M, K, lda, a_box, false);
  CB_REQUIRE(e == 0, "cb_gemm_tf32: cuTensorMapEncodeTiled failed for A (code %d)", e);
  e = b_mn ? make_tmap(&tb, B, K, N, ldb, 32, true) : make_tmap(&tb, B, N, K, ldb, 128, false);
  CB_REQUIRE(e == 0, "cb_gemm_tf32: cuTensorMapEncodeTiled failed for B (code %d)", e);
  ta2 = ta, tb2 = tb;
  if (K2 > 0) {
    e = a_mn ? make_tmap(&ta2, A2, K2, M, lda2, 32, true) : make_tmap(&ta2, A2, M, K2, lda2, a_box, false);
    CB_REQUIRE(e == 0, "cb_gemm_tf32: cuTensorMapEncodeTiled failed for A2 (code %d)", e);
    e = b_mn ? make_tmap(&tb2, B2, K2, N, ldb2, 32, true) : make_tmap(&tb2, B2, N, K2, ldb2, 128, false);
    CB_REQUIRE(e == 0, "cb_gemm_tf32: cuTensorMapEncodeTiled failed for B2 (code %d)", e);
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool two = M > 128;
  // tile choice: MT = 2 accumulators when M > 128; BN = 256 only with a single accumulator
  const bool wide = (!two) && N >= 256;
  int rc = 0;
#define CB_DISPATCH(MT, BN)                                                              \
  do {                                                                                   \
    if (!a_mn && !b_mn) rc = launch<MT, BN, false, false>(ta, tb, ta2, tb2, p, st);       \
    else if (!a_mn && b_mn) rc = launch<MT, BN, false, true>(ta, tb, ta2, tb2, p, st);    \
    else if (a_mn && !b_mn) rc = launch<MT, BN, true, false>(ta, tb, ta2, tb2, p, st);    \
    else rc = launch<MT, BN, true, true>(ta, tb, ta2, tb2, p, st);                        \
  } while (0)
  if (adam != nullptr) {
    // fused optimizer epilogue: the weight-gradient layout (both operands stored with the reduction dimension as
    // rows), short K; persistent kernel, one CTA per SM
    CB_REQUIRE(a_mn && b_mn && K2 == 0 && p.k_blocks_total <= 16,
               "cb_gemm_tf32_dw_adam: needs a_mn = b_mn = 1 and K <= 512 (got K=%d)", K);
    constexpr int kSmem = kPersistStages * (2 * kBlockM + 128) * kBlockK * 4 + kBlockM * kTilePitch * 4 + 1024;
    static bool configured = false;
    if (!configured) {
      CB_REQUIRE(cudaFuncSetAttribute(gemm_dw_adam_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) ==
                     cudaSuccess, "cb_gemm_tf32_dw_adam: cannot reserve %d bytes of shared memory", kSmem);
      configured = true;
    }
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int n_tiles = ((N + 127) / 128) * ((M + 255) / 256);
    gemm_dw_adam_persistent_kernel<<<n_tiles < n_sm ? n_tiles : n_sm, kPersistThreads, kSmem, st>>>(ta, tb, p);
  } else if (cluster4) {
    rc = b_mn ? launch_cluster4<true>(ta, tb, ta2, tb2, p, st) : launch_cluster4<false>(ta, tb, ta2, tb2, p, st);
  } else if (two) CB_DISPATCH(2, 128);
  else if (wide) CB_DISPATCH(1, 256);
  else CB_DISPATCH(1, 128);
#undef CB_DISPATCH
  CB_REQUIRE(rc == 0, "cb_gemm_tf32: kernel configuration failed (code %d)", rc);
  if (split_k > 1) {
    const long long total = (long long)M * N;
    splitk_reduce_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(workspace, split_k, p.Mpad, p.Npad, C, ldc, M, N, bias, accumulate);
  }
  return cb::check_launch("cb_gemm_tf32");
}

extern "C" int cb_gemm_tf32_pair(const float* A, long long lda, const float* B, long long ldb, int K, const float* A2,
                                 long long lda2, const float* B2, long long ldb2, int K2, int a_mn, int b_mn, float* C,
                                 long long ldc, int M, int N, const float* bias, int accumulate, int split_k,
                                 float* workspace, long long workspace_elems, void* stream) {
  return gemm_tf32_impl(A, lda, B, ldb, K, A2, lda2, B2, ldb2, K2, a_mn, b_mn, C, ldc, M, N, bias, accumulate, split_k, workspace,
                        workspace_elems, nullptr, stream);
}

extern "C" int cb_gemm_tf32(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn, float* C,
                            long long ldc, int M, int N, int K, const float* bias, int accumulate, int split_k,
                            float* workspace, long long workspace_elems, void* stream) {
  return gemm_tf32_impl(A, lda, B, ldb, K, A, lda, B, ldb, 0, a_mn, b_mn, C, ldc, M, N, bias, accumulate, split_k, workspace,
                        workspace_elems, nullptr, stream);
}

// Weight-gradient product with the optimizer fused into the epilogue:  G = A B^T (never stored), then the
// ClippedAdam update of cb_clipped_adam_step applied to param / exp_avg / exp_avg_sq, which are indexed like C
// ([M, ldc] with the same leading dimension).  *step_ptr is read, not advanced.
extern "C" int cb_gemm_tf32_dw_adam(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn,
                                    long long ldc, int M, int N, int K, float* param, float* exp_avg, float* exp_avg_sq,
                                    const float* lr_table, const float* beta1_table, long long table_len,
                                    const long long* step_ptr, float beta2, float eps, float clip, void* stream) {
  CB_REQUIRE(param != nullptr && exp_avg != nullptr && exp_avg_sq != nullptr && lr_table != nullptr && beta1_table != nullptr &&
                 step_ptr != nullptr && table_len > 0, "cb_gemm_tf32_dw_adam: optimizer state pointers must not be NULL");
  cb::AdamArgs a{param, exp_avg, exp_avg_sq, lr_table, beta1_table, table_len, step_ptr, beta2, eps, clip};
  return gemm_tf32_impl(A, lda, B, ldb, K, A, lda, B, ldb, 0, a_mn, b_mn, param, ldc, M, N, nullptr, 0, 1, nullptr, 0, &a, stream);
}
