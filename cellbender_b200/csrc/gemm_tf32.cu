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

// ---- TMA store (shared -> global through a tensor map; out-of-bounds rows / columns of the box are clipped) -----------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)),
               "r"(c0), "r"(c1)
               : "memory");
}
// C += tile (element-wise fp32 add performed by the TMA unit at the L2)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

struct GemmParams {
  const float* bias;     // [N] or nullptr (direct path only; the reduce kernel adds it on the split path)
  int M, N, K;
  int k_blocks_total;    // ceil(K / 32) (+ ceil(K2 / 32) with a second operand pair)
  int k_blocks_first;    // k-blocks of the first operand pair
  int k_blocks_per_split;
  int split_k;
  int accumulate;        // C += ... (direct path: TMA reduce-add)
  long long Mpad;        // rows of one split's partial tile in the workspace
};

// MT = accumulators of 128 rows per CTA, BN = tile width; MT * BN <= 512 TMEM columns.
// MINB = CTAs per SM the launch is sized for: short-K products (weight gradients: K = 255 droplets; the hidden layers)
// run two CTAs per SM with a 2-stage ring (2 x 256 TMEM columns, 2 x 97 KB of shared memory), so one CTA's epilogue
// overlaps the other's loads -- each CTA lives for only ~8 k-blocks.
//
// Epilogue: each of the four epilogue warps drains its 32 TMEM lanes 32 columns at a time (tcgen05.ld), adds the bias,
// writes the 32 x 32 fp32 sub-tile into a 128B-swizzled staging buffer carved out of the idle operand ring (conflict-free
// float4 phases) and hands it to the TMA unit (cp.async.bulk.tensor store, or cp.reduce...add for C +=): every global
// write is a full 128-byte line, ragged edges (M = 255, N % 32 != 0) are clipped by the tensor map, and the store of
// sub-tile i overlaps the TMEM drain of sub-tile i + 1 (two staging buffers per warp).  tmap_c describes C [M, N]
// (split_k == 1) or the split-K workspace [split_k * Mpad, Npad].
template <int MT, int BN, bool A_MN, bool B_MN, int STAGES, int MINB>
__global__ void __launch_bounds__(kGemmThreads, MINB)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_a2, const __grid_constant__ CUtensorMap tmap_b2,
                 const __grid_constant__ CUtensorMap tmap_c, GemmParams p) {
  constexpr uint32_t kABytes = MT * kBlockM * kBlockK * 4;  // per stage
  constexpr uint32_t kBBytes = BN * kBlockK * 4;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = (MT * BN <= 32) ? 32 : (MT * BN <= 64) ? 64 : (MT * BN <= 128) ? 128 : (MT * BN <= 256) ? 256 : 512;
  static_assert(MT * BN <= 512, "accumulators exceed TMEM");
  static_assert(STAGES * kStageBytes >= 4 * 2 * 4096, "operand ring too small to host the epilogue staging buffers");

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
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

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
        if constexpr (!A_MN) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) tma_load_2d(ma, &full_bar[s], a_dst + mt * 16384, k0, m0 + mt * kBlockM);
        } else {
#pragma unroll
          for (int j = 0; j < MT * 4; ++j) tma_load_2d(ma, &full_bar[s], a_dst + j * 4096, m0 + j * 32, k0);
        }
        if constexpr (!B_MN) {
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
          // K-major: 128-row groups of B sit 16 KB apart (one TMA box each); MN-major: 32-column groups sit 4 KB apart
          const uint64_t bdesc = B_MN ? make_smem_desc(b_base + k4 * 1024, 4096, 512, 1) : make_smem_desc(b_base + k4 * 32, 16, 1024, 2);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t adesc = A_MN ? make_smem_desc(a_base + mt * 16384 + k4 * 1024, 4096, 512, 1)
                                        : make_smem_desc(a_base + mt * 16384 + k4 * 32, 16, 1024, 2);
            umma_tf32(tmem_base + mt * BN, adesc, bdesc, idesc, (i > 0 || k4 > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);  // stage free once these MMAs have read it
      }
      umma_commit(&accum_bar);  // accumulator complete
    }
  } else {
    // ======================= epilogue: TMEM -> registers -> swizzled staging -> TMA store =======================
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint8_t* stg = smem + (warp - 2) * 8192;  // two 4 KB buffers per warp, 1024-byte aligned (128B swizzle atom)
    const bool to_ws = p.split_k > 1;
    const int row_base = (to_ws ? int(split * p.Mpad) : 0) + m0 + q * 32;
    const int sw = lane & 7;
    int it = 0;
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      if (!to_ws && m0 + mt * kBlockM + q * 32 >= p.M) continue;  // the whole sub-tile row band is outside C
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        if (n0 + c >= p.N) break;
        float v[32];
        if (n_kb > 0) {
          tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(mt * BN + c), v);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        if (p.bias != nullptr) {
          const int col = n0 + c;
          if (col + 32 <= p.N) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + col + i);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col + i < p.N) v[i] += __ldg(p.bias + col + i);
          }
        }
        uint8_t* buf = stg + (it & 1) * 4096;
        if (it >= 2) {  // the store issued two sub-tiles ago has finished reading this buffer
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (p.accumulate && !to_ws) tma_reduce_add_2d(&tmap_c, buf, n0 + c, row_base + mt * kBlockM);
          else tma_store_2d(&tmap_c, buf, n0 + c, row_base + mt * kBlockM);
          tma_store_commit();
        }
        ++it;
      }
    }
    if (lane == 0) tma_store_wait_read<0>();  // the staging buffers must outlive the reads of the TMA unit
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// out[m, n] (=, +=) sum_s ws[s][m][n] + bias[n]   -- fixed summation order; 4 columns per thread
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int split_k, long long Mpad, long long Npad,
                                                            float* __restrict__ C, long long ldc, int M, int N,
                                                            const float* __restrict__ bias, int accumulate) {
  const int n4 = (N + 3) >> 2;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)M * n4) return;
  const int m = int(idx / n4), n = int(idx % n4) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < split_k; ++s) {
    const float4 w = *reinterpret_cast<const float4*>(ws + ((long long)s * Mpad + m) * Npad + n);  // Npad % 4 == 0
    acc.x += w.x, acc.y += w.y, acc.z += w.z, acc.w += w.w;
  }
  const float a[4] = {acc.x, acc.y, acc.z, acc.w};
  float* dst = C + (long long)m * ldc + n;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (n + i >= N) break;
    float o = a[i];
    if (bias != nullptr) o += bias[n + i];
    dst[i] = accumulate ? (dst[i] + o) : o;
  }
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

// tile_cfg values of cb_gemm_tf32_ex (0 = the heuristic below)
enum TileCfg { kAuto = 0, k2x128_s2 = 1, k2x128_deep = 2, k2x256_deep = 3, k1x256_s2 = 4, k1x256_deep = 5, k1x128_s2 = 6, k1x128_deep = 7 };

template <int MT, int BN, bool A_MN, bool B_MN, int STAGES, int MINB>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const CUtensorMap& tb2,
                      const CUtensorMap& tc, const GemmParams& p, cudaStream_t st) {
  constexpr int kStage = (MT * kBlockM + BN) * kBlockK * 4;
  const int smem = STAGES * kStage + 1024;
  auto kern = gemm_tf32_kernel<MT, BN, A_MN, B_MN, STAGES, MINB>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 3;
    configured = true;
  }
  dim3 grid((p.N + BN - 1) / BN, (p.M + MT * kBlockM - 1) / (MT * kBlockM), p.split_k);
  kern<<<grid, kGemmThreads, smem, st>>>(ta, tb, ta2, tb2, tc, p);
  return 0;
}

template <int MT, int BN, bool A_MN, bool B_MN>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const CUtensorMap& tb2,
                  const CUtensorMap& tc, const GemmParams& p, bool deep, cudaStream_t st) {
  constexpr int kStage = (MT * kBlockM + BN) * kBlockK * 4;
  constexpr int kDeep = (kStage <= 48 * 1024) ? 4 : 3;
  if constexpr (MT * BN <= 256) {
    if (!deep) return launch_cfg<MT, BN, A_MN, B_MN, 2, 2>(ta, tb, ta2, tb2, tc, p, st);  // two co-resident CTAs per SM
  }
  return launch_cfg<MT, BN, A_MN, B_MN, kDeep, 1>(ta, tb, ta2, tb2, tc, p, st);
}

// Measured on B200 (tools/gemm_bench.py, profiles/r02_gemm_bench.txt; PBMC8k shapes, L2 flushed):
//   dW = dY^T X (M = 512, N = G, K = 255)      256x256 deep 36.9 us | 256x128 s2x2 39.0 | 256x128 deep 44.0
//   dX pair (M = 255, N = G, K = 2 x 512)      256x256 deep 48.1 us | 256x128 s2x2 51.1
//   dWd = dlogits^T h (M = G, N = 512, K = 255) 128x256 s2x2 32.8 us | 256x128 s2x2 34.8
//   decoder forward (M = 255, N = G, K = 512)  256x128 s2x2 33.8 us | 256x256 deep 36.9
// The 256-wide tiles win where the small operand (dY, 0.5-1 MB) is re-read from L2 by every column tile: they halve that
// traffic; the K-major forward products prefer two co-resident CTAs per SM.
static int tile_heuristic(int M, int N, const GemmParams& p, bool a_mn, bool b_mn, bool pair) {
  if (M > 128 && N >= 4096 && p.split_k == 1 && (a_mn || pair)) return k2x256_deep;
  if (a_mn && b_mn && M >= 4096 && N >= 256 && p.split_k == 1 && p.k_blocks_per_split <= 16) return k1x256_s2;
  const bool two = M > 128;
  const bool wide = (!two) && N >= 256;
  // short K: two co-resident CTAs per SM with a 2-stage ring each (one CTA's epilogue overlaps the other's loads) -- but
  // only when the grid is large enough to fill the SMs twice; a small grid is latency-bound per CTA and wants the deep ring
  const long long ctas = (long long)((N + (wide ? 255 : 127)) / (wide ? 256 : 128)) * ((M + (two ? 255 : 127)) / (two ? 256 : 128)) * p.split_k;
  const bool shallow = p.k_blocks_per_split <= 16 && ctas > sm_count();
  if (two) return shallow ? k2x128_s2 : k2x128_deep;
  if (wide) return shallow ? k1x256_s2 : k1x256_deep;
  return shallow ? k1x128_s2 : k1x128_deep;
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
  const int n_sm = cb::sm_count();
  if (tiles >= n_sm || kb < 16) return 1;
  long long s = n_sm / tiles;
  if (s > kb / 4) s = kb / 4;
  return int(s < 1 ? 1 : s);
}

static int gemm_tf32_impl(const float* A, long long lda, const float* B, long long ldb, int K, const float* A2,
                          long long lda2, const float* B2, long long ldb2, int K2, int a_mn, int b_mn, float* C,
                          long long ldc, int M, int N, const float* bias, int accumulate, int split_k,
                          float* workspace, long long workspace_elems, int tile_cfg, void* stream) {
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
  p.bias = bias, p.M = M, p.N = N, p.K = K;
  p.k_blocks_total = kb_total;
  p.k_blocks_first = kb_first;
  p.k_blocks_per_split = (kb_total + split_k - 1) / split_k;
  split_k = (kb_total + p.k_blocks_per_split - 1) / p.k_blocks_per_split;  // every split owns >= 1 k-block
  p.split_k = split_k;
  p.accumulate = accumulate;
  p.Mpad = (M + 255) / 256 * 256;
  const long long Npad = (N + 255) / 256 * 256;
  CUtensorMap ta, tb, ta2, tb2, tc;
  int e;
  if (split_k > 1) {
    CB_REQUIRE(workspace != nullptr && workspace_elems >= (long long)split_k * p.Mpad * Npad,
               "cb_gemm_tf32: workspace too small (%lld < %lld)", workspace_elems, (long long)split_k * p.Mpad * Npad);
    CB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "cb_gemm_tf32: workspace must be 16-byte aligned");
    p.bias = nullptr;
    e = make_tmap(&tc, workspace, (long long)split_k * p.Mpad, Npad, Npad, 32, false);
  } else {
    CB_REQUIRE((ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0,
               "cb_gemm_tf32: C must be 16-byte aligned with a leading dimension that is a multiple of 4 (TMA store)");
    e = make_tmap(&tc, C, M, N, ldc, 32, false);
  }
  CB_REQUIRE(e == 0, "cb_gemm_tf32: cuTensorMapEncodeTiled failed for C (code %d)", e);
  // K-major operand: tensor [rows = M|N, cols = K], box 32 x 128 rows.  MN-major: tensor [rows = K, cols = M|N], box 32 x 32 rows.
  e = a_mn ? make_tmap(&ta, A, K, M, lda, 32, true) : make_tmap(&ta, A, M, K, lda, 128, false);
  CB_REQUIRE(e == 0, "cb_gemm_tf32: cuTensorMapEncodeTiled failed for A (code %d)", e);
  e = b_mn ? make_tmap(&tb, B, K, N, ldb, 32, true) : make_tmap(&tb, B, N, K, ldb, 128, false);
  CB_REQUIRE(e == 0, "cb_gemm_tf32: cuTensorMapEncodeTiled failed for B (code %d)", e);
  ta2 = ta, tb2 = tb;
  if (K2 > 0) {
    e = a_mn ? make_tmap(&ta2, A2, K2, M, lda2, 32, true) : make_tmap(&ta2, A2, M, K2, lda2, 128, false);
    CB_REQUIRE(e == 0, "cb_gemm_tf32: cuTensorMapEncodeTiled failed for A2 (code %d)", e);
    e = b_mn ? make_tmap(&tb2, B2, K2, N, ldb2, 32, true) : make_tmap(&tb2, B2, N, K2, ldb2, 128, false);
    CB_REQUIRE(e == 0, "cb_gemm_tf32: cuTensorMapEncodeTiled failed for B2 (code %d)", e);
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (tile_cfg == kAuto) tile_cfg = tile_heuristic(M, N, p, a_mn != 0, b_mn != 0, K2 > 0);
  CB_REQUIRE(tile_cfg >= k2x128_s2 && tile_cfg <= k1x128_deep, "cb_gemm_tf32_ex: tile_cfg=%d not in [0, 7]", tile_cfg);
  const bool deep = tile_cfg == k2x128_deep || tile_cfg == k2x256_deep || tile_cfg == k1x256_deep || tile_cfg == k1x128_deep;
  int rc = 0;
#define CB_DISPATCH(MT, BN)                                                                    \
  do {                                                                                         \
    if (!a_mn && !b_mn) rc = launch<MT, BN, false, false>(ta, tb, ta2, tb2, tc, p, deep, st);   \
    else if (!a_mn && b_mn) rc = launch<MT, BN, false, true>(ta, tb, ta2, tb2, tc, p, deep, st); \
    else if (a_mn && !b_mn) rc = launch<MT, BN, true, false>(ta, tb, ta2, tb2, tc, p, deep, st); \
    else rc = launch<MT, BN, true, true>(ta, tb, ta2, tb2, tc, p, deep, st);                     \
  } while (0)
  if (tile_cfg == k2x128_s2 || tile_cfg == k2x128_deep) CB_DISPATCH(2, 128);
  else if (tile_cfg == k2x256_deep) CB_DISPATCH(2, 256);
  else if (tile_cfg == k1x256_s2 || tile_cfg == k1x256_deep) CB_DISPATCH(1, 256);
  else CB_DISPATCH(1, 128);
#undef CB_DISPATCH
  CB_REQUIRE(rc == 0, "cb_gemm_tf32: kernel configuration failed (code %d)", rc);
  if (split_k > 1) {
    const long long total = (long long)M * ((N + 3) / 4);
    splitk_reduce_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(workspace, split_k, p.Mpad, Npad, C, ldc, M, N, bias, accumulate);
  }
  return cb::check_launch("cb_gemm_tf32");
}

extern "C" int cb_gemm_tf32_pair(const float* A, long long lda, const float* B, long long ldb, int K, const float* A2,
                                 long long lda2, const float* B2, long long ldb2, int K2, int a_mn, int b_mn, float* C,
                                 long long ldc, int M, int N, const float* bias, int accumulate, int split_k,
                                 float* workspace, long long workspace_elems, void* stream) {
  return gemm_tf32_impl(A, lda, B, ldb, K, A2, lda2, B2, ldb2, K2, a_mn, b_mn, C, ldc, M, N, bias, accumulate, split_k, workspace,
                        workspace_elems, 0, stream);
}

extern "C" int cb_gemm_tf32(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn, float* C,
                            long long ldc, int M, int N, int K, const float* bias, int accumulate, int split_k,
                            float* workspace, long long workspace_elems, void* stream) {
  return gemm_tf32_impl(A, lda, B, ldb, K, A, lda, B, ldb, 0, a_mn, b_mn, C, ldc, M, N, bias, accumulate, split_k, workspace,
                        workspace_elems, 0, stream);
}

// Same product with an explicit tile configuration (0 = heuristic; 1..7 see TileCfg): the tuning / benchmarking entry.
extern "C" int cb_gemm_tf32_ex(const float* A, long long lda, const float* B, long long ldb, int K, const float* A2,
                               long long lda2, const float* B2, long long ldb2, int K2, int a_mn, int b_mn, float* C,
                               long long ldc, int M, int N, const float* bias, int accumulate, int split_k,
                               float* workspace, long long workspace_elems, int tile_cfg, void* stream) {
  return gemm_tf32_impl(A, lda, B, ldb, K, K2 > 0 ? A2 : A, K2 > 0 ? lda2 : lda, K2 > 0 ? B2 : B, K2 > 0 ? ldb2 : ldb, K2, a_mn,
                        b_mn, C, ldc, M, N, bias, accumulate, split_k, workspace, workspace_elems, tile_cfg, stream);
}
