// K8: fused multi-tensor ClippedAdam + OneCycleLR step over ONE flat parameter buffer.
//
// Replaces the reference's ~40 per-tensor pyro ClippedAdam optimizers, each wrapped in its own torch OneCycleLR
// (reference run.py:593-629; stepped at train.py:56-60): per tensor that is clamp_, mul_, add_, addcmul_, sqrt, add_,
// addcdiv_ = 7 kernels x 40 tensors per step, and 7 passes over the 278 MB of parameters/moments.  Here all
// parameters live in one contiguous fp32 buffer (nn.Parameters are views into it) and one kernel makes the single
// read-modify-write pass: 28 bytes / parameter = 1.94 GB per step at the PBMC8k shape, the dominant HBM term of a step.
//
// Semantics (pyro.optim.clipped_adam.ClippedAdam.step, pyro-ppl 1.8.x, restated from its published source):
//   g = clamp(g, -clip, clip);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
//   p -= lr * sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps);   lr *= lrd (lrd = 1 by default)
// with lr and b1 taken from torch.optim.lr_scheduler.OneCycleLR (cycle_momentum: b1 cycles 0.95 -> 0.85 -> 0.95),
// precomputed on the host into per-step tables; the step counter lives on the device so that the whole training
// step can be replayed as a CUDA graph.
#include "common.cuh"

namespace cb {

// Scheduling: the sweep is HBM-bound and runs on a side stream WHILE the latency-bound chain of backward kernels runs on
// others (optim.early_block_update).  A persistent grid (r01: 2 048 threads/SM for the whole 75 us) left no thread slots
// -- and, capped at 1 024 threads/SM, no registers -- for the small kernels of that chain, which then waited up to 50 us
// for a slot (r02 timelines).  Hence short-lived CTAs: each one owns ONE tile of 256 x 4 float4 per array (16 KB each,
// all 16 loads of a thread issued before the first use), lives ~2 us and frees its slot, so CTAs of concurrent
// (higher-priority) kernels are placed within microseconds while the sweep still keeps > 150 KB per SM in flight.
// The per-step scalars come from host-built tables (lr, beta1, and the bias-corrected step size
// lr * sqrt(1 - b2^t) / (1 - b1^t) in float64); without a step-size table (constant learning rate: t is unbounded)
// the kernel evaluates it itself.
constexpr int kAdamUnroll = 4;
constexpr int kAdamTile = 256 * kAdamUnroll;  // float4 per CTA

struct StepCoef {
  float b1, omb1, step_size;
};
__device__ __forceinline__ StepCoef step_coef(const float* __restrict__ lr_table, const float* __restrict__ beta1_table,
                                              const float* __restrict__ step_table, long long table_len,
                                              const long long* __restrict__ step_ptr, float beta2) {
  const long long t0 = *step_ptr;  // number of optimizer steps already taken
  const long long ti = t0 < table_len ? t0 : table_len - 1;
  StepCoef c;
  c.b1 = beta1_table[ti];
  c.omb1 = 1.0f - c.b1;
  if (step_table != nullptr && t0 < table_len) {
    c.step_size = step_table[ti];
  } else {
    const double t = double(t0 + 1);
    c.step_size = float(double(lr_table[ti]) * sqrt(1.0 - pow(double(beta2), t)) / (1.0 - pow(double(c.b1), t)));
  }
  return c;
}

__global__ void __launch_bounds__(256) clipped_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                           float* __restrict__ m, float* __restrict__ v, long long n,
                                                           const float* __restrict__ lr_table,
                                                           const float* __restrict__ beta1_table,
                                                           const float* __restrict__ step_table, long long table_len,
                                                           const long long* __restrict__ step_ptr, float beta2, float eps,
                                                           float clip, float grad_scale) {
  const StepCoef sc = step_coef(lr_table, beta1_table, step_table, table_len, step_ptr, beta2);
  const float b1 = sc.b1, omb1 = sc.omb1, step_size = sc.step_size, omb2 = 1.0f - beta2;
  const long long n4 = n >> 2;
  const long long i0 = blockIdx.x * (long long)kAdamTile + threadIdx.x;
  float4 pp[kAdamUnroll], gg[kAdamUnroll], mm[kAdamUnroll], vv[kAdamUnroll];
#pragma unroll
  for (int u = 0; u < kAdamUnroll; ++u) {  // all loads first
    const long long i = i0 + u * 256;
    if (i < n4) {
      pp[u] = reinterpret_cast<float4*>(p)[i];
      gg[u] = ldg_stream4(g + 4 * i);
      mm[u] = reinterpret_cast<float4*>(m)[i];
      vv[u] = reinterpret_cast<float4*>(v)[i];
    }
  }
#pragma unroll
  for (int u = 0; u < kAdamUnroll; ++u) {
    const long long i = i0 + u * 256;
    if (i >= n4) break;
    float* pa = reinterpret_cast<float*>(&pp[u]);
    const float* ga = reinterpret_cast<const float*>(&gg[u]);
    float* ma = reinterpret_cast<float*>(&mm[u]);
    float* va = reinterpret_cast<float*>(&vv[u]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gc = fminf(fmaxf(ga[k] * grad_scale, -clip), clip);
      ma[k] = b1 * ma[k] + omb1 * gc;
      va[k] = beta2 * va[k] + omb2 * gc * gc;
      pa[k] -= step_size * ma[k] / (sqrtf(va[k]) + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pp[u];
    reinterpret_cast<float4*>(m)[i] = mm[u];
    reinterpret_cast<float4*>(v)[i] = vv[u];
  }
  if (blockIdx.x == gridDim.x - 1) {  // the < 4 trailing elements
    for (long long i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) {
      const float gc = fminf(fmaxf(g[i] * grad_scale, -clip), clip);
      const float mk = b1 * m[i] + omb1 * gc;
      const float vk = beta2 * v[i] + omb2 * gc * gc;
      m[i] = mk, v[i] = vk;
      p[i] -= step_size * mk / (sqrtf(vk) + eps);
    }
  }
}

__global__ void advance_step_kernel(long long* step_ptr) { *step_ptr += 1; }

// Data-parallel exchange fused into the optimizer sweep (one kernel = reduce-scatter + ClippedAdam + all-gather over
// NVLink peer memory).  Every rank owns one contiguous shard of the flat buffers.  For its shard it reads the gradient
// from EVERY rank's flat gradient buffer (its own from HBM, the peers' through NVLink P2P loads), sums them in rank
// order, applies the update to its param / exp_avg / exp_avg_sq, and stores the new parameter values into its own
// flat parameter buffer AND into every peer's (P2P stores).  The gradient and parameter buffers are symmetric-memory
// allocations (torch.distributed._symmetric_memory); a device-side barrier before the kernel (all ranks finished
// backward) and one after it (all shards delivered) are the only other traffic.  NVLink carries 4 bytes per element in
// and 4 (n_ranks - 1) out while HBM carries the sweep's 28: at two ranks both take about the same time, so the exchange
// costs no extra pass over memory at all.
constexpr int kMaxRanks = 16;
struct P2PPointers {
  const float* grad[kMaxRanks];  // this shard inside every rank's gradient buffer, in rank order
  float* peer_param[kMaxRanks];  // this shard inside every OTHER rank's parameter buffer
  int n_grad, n_peer;
  int own;  // index in grad[] of this rank's own (local HBM) gradient buffer
};

// ---- the same sweep with the peers' gradient shards fetched by TMA bulk copies ------------------------------------
// Plain P2P loads keep only ~20 KB per SM in flight (one float4 per thread), far too little for the NVLink round trip:
// the kernel above reaches ~320 GB/s per direction.  Here one elected thread per CTA streams the remote gradient tiles
// (4 KB per peer per tile) into a shared-memory ring with cp.async.bulk (global -> shared, completion on an mbarrier),
// several stages ahead of the threads that consume them, so each SM keeps STAGES x n_peers x 4 KB x (resident CTAs) in
// flight.  Local operands (own gradient, param, exp_avg, exp_avg_sq) are ordinary coalesced loads issued before the
// wait; the updated parameters go to the peers as posted P2P stores.
__device__ __forceinline__ uint32_t adam_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void adam_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(adam_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void adam_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(adam_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void adam_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(adam_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void adam_bulk_load(void* dst_smem, const void* src_global, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   adam_smem_u32(dst_smem)),
               "l"(src_global), "r"(bytes), "r"(adam_smem_u32(bar))
               : "memory");
}

constexpr int kP2PTile = 256;  // float4 per tile = one per thread (4 KB per gradient source)

template <int NG, int STAGES>
__global__ void __launch_bounds__(256) clipped_adam_p2p_bulk_kernel(float* __restrict__ p, float* __restrict__ m,
                                                                    float* __restrict__ v, long long n, P2PPointers q,
                                                                    const float* __restrict__ lr_table,
                                                                    const float* __restrict__ beta1_table,
                                                                    long long table_len, const long long* __restrict__ step_ptr,
                                                                    float beta2, float eps, float clip) {
  extern __shared__ __align__(128) unsigned char p2p_ring[];  // [STAGES][n_remote][kP2PTile] float4
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  const long long t0 = *step_ptr;
  const long long ti = t0 < table_len ? t0 : table_len - 1;
  const float lr = lr_table[ti], b1 = beta1_table[ti];
  const double t = double(t0 + 1);
  const float step_size = float(double(lr) * sqrt(1.0 - pow(double(beta2), t)) / (1.0 - pow(double(b1), t)));
  const float omb1 = 1.0f - b1, omb2 = 1.0f - beta2;
  const long long n4 = n >> 2;
  const long long n_tiles = (n4 + kP2PTile - 1) / kP2PTile;
  const int n_remote = q.n_grad - 1;
  const int tid = threadIdx.x;
  float4* ring = reinterpret_cast<float4*>(p2p_ring);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) adam_mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](long long tile, int s) {  // elected thread: fetch every peer's piece of this tile
    const long long f0 = tile * kP2PTile;
    const long long left = n4 - f0;
    const uint32_t bytes = uint32_t(left < kP2PTile ? left : kP2PTile) * 16u;
    adam_mbar_expect_tx(&full_bar[s], bytes * uint32_t(n_remote));
    int r = 0;
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      if (k < q.n_grad && k != q.own) {
        adam_bulk_load(ring + (size_t(s) * n_remote + r) * kP2PTile, reinterpret_cast<const float4*>(q.grad[k]) + f0, bytes,
                       &full_bar[s]);
        ++r;
      }
    }
  };

  if (tid == 0 && n_remote > 0) {
    for (int s = 0; s < STAGES; ++s) {
      const long long tile = blockIdx.x + (long long)s * gridDim.x;
      if (tile < n_tiles) issue(tile, s);
    }
  }
  long long it = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int s = int(it % STAGES);
    const uint32_t parity = uint32_t(it / STAGES) & 1u;
    const long long i = tile * kP2PTile + tid;
    const bool active = i < n4;
    float4 gl = make_float4(0.f, 0.f, 0.f, 0.f), pp = gl, mm = gl, vv = gl;
    if (active) {
      gl = __ldcg(reinterpret_cast<const float4*>(q.grad[q.own]) + i);
      pp = reinterpret_cast<float4*>(p)[i];
      mm = reinterpret_cast<float4*>(m)[i];
      vv = reinterpret_cast<float4*>(v)[i];
    }
    if (n_remote > 0) adam_mbar_wait(&full_bar[s], parity);
    if (active) {
      // sum in rank order (the result does not depend on which rank owns the shard)
      float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
      int r = 0;
#pragma unroll
      for (int k = 0; k < NG; ++k) {
        if (k < q.n_grad) {
          float4 x;
          if (k == q.own) x = gl;
          else x = ring[(size_t(s) * n_remote + r++) * kP2PTile + tid];
          if (k == 0) gg = x;
          else gg.x += x.x, gg.y += x.y, gg.z += x.z, gg.w += x.w;
        }
      }
      float* pa = reinterpret_cast<float*>(&pp);
      const float* ga = reinterpret_cast<const float*>(&gg);
      float* ma = reinterpret_cast<float*>(&mm);
      float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gc = fminf(fmaxf(ga[k], -clip), clip);
        ma[k] = b1 * ma[k] + omb1 * gc;
        va[k] = beta2 * va[k] + omb2 * gc * gc;
        pa[k] -= step_size * ma[k] / (sqrtf(va[k]) + eps);
      }
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
#pragma unroll
      for (int k = 0; k < NG - 1; ++k)
        if (k < q.n_peer) __stcg(reinterpret_cast<float4*>(q.peer_param[k]) + i, pp);
    }
    __syncthreads();  // every thread has read stage s: it can be refilled
    const long long next = tile + (long long)STAGES * gridDim.x;
    if (tid == 0 && n_remote > 0 && next < n_tiles) issue(next, s);
  }
}

// ---- the same sweep through the NVSwitch (NVLS): in-switch gradient reduction, multicast parameter store -----------
// mc_grad / mc_param are MULTICAST addresses of the symmetric gradient / parameter buffers (one address that names
// the same offset on every rank).  multimem.ld_reduce returns the sum over all ranks computed inside the switch, so a
// rank receives 16 bytes per float4 however many ranks there are (plain P2P loads: 16 (R - 1)); multimem.st delivers
// the updated parameters to every rank with ONE outgoing store.  Per rank and direction the links then carry about
// one parameter set per step for any R, where the plain P2P kernel carries 2 (R - 1) / R of it.
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc_ptr) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc_ptr)
               : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st(float* mc_ptr, const float4& x) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_ptr), "f"(x.x), "f"(x.y), "f"(x.z),
               "f"(x.w)
               : "memory");
}

// kNvlsUnroll in-switch reductions are issued back to back before the first one is consumed: the round trip through the
// NVSwitch is several microseconds, and one 16-byte request per thread (r01) left the links latency-bound.
constexpr int kNvlsUnroll = 4;
struct NvlsPeers {
  float* param[kMaxRanks];  // this shard inside every OTHER rank's parameter buffer (ST_P2P only)
  int n_peer;
};
// LD: 0 = multimem.ld_reduce of the gradient (in-switch sum), 1 = this rank's gradient only (diagnostic: isolates the
// store path).  ST: 0 = multimem.st (one multicast store), 1 = plain store to the local buffer + posted P2P stores to
// every peer, 2 = local store only (diagnostic: isolates the reduction path).
template <int LD, int ST>
__global__ void __launch_bounds__(256) clipped_adam_nvls_kernel(float* p, float* __restrict__ m, float* __restrict__ v, long long n,
                                                                const float* mc_grad, float* mc_param, const float* local_grad,
                                                                NvlsPeers peers, const float* __restrict__ lr_table,
                                                                const float* __restrict__ beta1_table, long long table_len,
                                                                const long long* __restrict__ step_ptr, float beta2, float eps,
                                                                float clip) {
  const long long t0 = *step_ptr;
  const long long ti = t0 < table_len ? t0 : table_len - 1;
  const float lr = lr_table[ti], b1 = beta1_table[ti];
  const double t = double(t0 + 1);
  const float step_size = float(double(lr) * sqrt(1.0 - pow(double(beta2), t)) / (1.0 - pow(double(b1), t)));
  const float omb1 = 1.0f - b1, omb2 = 1.0f - beta2;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < n4; i0 += stride * kNvlsUnroll) {
    float4 gg[kNvlsUnroll], pp[kNvlsUnroll], mm[kNvlsUnroll], vv[kNvlsUnroll];
#pragma unroll
    for (int u = 0; u < kNvlsUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i < n4) {
        if (LD == 0) gg[u] = multimem_ld_reduce_add(mc_grad + 4 * i);
        else gg[u] = __ldcg(reinterpret_cast<const float4*>(local_grad) + i);
      }
    }
#pragma unroll
    for (int u = 0; u < kNvlsUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i < n4) {
        pp[u] = reinterpret_cast<const float4*>(p)[i];
        mm[u] = reinterpret_cast<const float4*>(m)[i];
        vv[u] = reinterpret_cast<const float4*>(v)[i];
      }
    }
#pragma unroll
    for (int u = 0; u < kNvlsUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i >= n4) break;
      float* pa = reinterpret_cast<float*>(&pp[u]);
      const float* ga = reinterpret_cast<const float*>(&gg[u]);
      float* ma = reinterpret_cast<float*>(&mm[u]);
      float* va = reinterpret_cast<float*>(&vv[u]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gc = fminf(fmaxf(ga[k], -clip), clip);
        ma[k] = b1 * ma[k] + omb1 * gc;
        va[k] = beta2 * va[k] + omb2 * gc * gc;
        pa[k] -= step_size * ma[k] / (sqrtf(va[k]) + eps);
      }
      reinterpret_cast<float4*>(m)[i] = mm[u];
      reinterpret_cast<float4*>(v)[i] = vv[u];
      if (ST == 0) {
        multimem_st(mc_param + 4 * i, pp[u]);  // every rank's parameter buffer, this rank's included
      } else {
        reinterpret_cast<float4*>(p)[i] = pp[u];
        if (ST == 1) {
#pragma unroll
          for (int k = 0; k < kMaxRanks - 1; ++k)
            if (k < peers.n_peer) __stcg(reinterpret_cast<float4*>(peers.param[k]) + i, pp[u]);
        }
      }
    }
  }
}

}  // namespace cb

static int adam_range(float* p, const float* g, float* m, float* v, long long n, const float* lr_table,
                      const float* beta1_table, const float* step_table, long long table_len, long long* step_ptr, float beta2,
                      float eps, float clip, float grad_scale, int advance, void* stream);

extern "C" int cb_clipped_adam_step(float* p, const float* g, float* m, float* v, long long n, const float* lr_table,
                                    const float* beta1_table, const float* step_size_table, long long table_len,
                                    long long* step_ptr, float beta2, float eps, float clip, float grad_scale, void* stream) {
  return adam_range(p, g, m, v, n, lr_table, beta1_table, step_size_table, table_len, step_ptr, beta2, eps, clip, grad_scale, 1,
                    stream);
}

// One contiguous range of the flat buffers without advancing the step counter (n = 0: nothing to update); with
// advance != 0 the counter is incremented after the range.  Used when parts of the buffer are updated elsewhere
// (cb_gemm_tf32_dw_adam) and the rest is swept range by range.
extern "C" int cb_clipped_adam_range(float* p, const float* g, float* m, float* v, long long n, const float* lr_table,
                                     const float* beta1_table, const float* step_size_table, long long table_len,
                                     long long* step_ptr, float beta2, float eps, float clip, int advance, void* stream) {
  return adam_range(p, g, m, v, n, lr_table, beta1_table, step_size_table, table_len, step_ptr, beta2, eps, clip, 1.0f, advance,
                    stream);
}


static int adam_range(float* p, const float* g, float* m, float* v, long long n, const float* lr_table,
                      const float* beta1_table, const float* step_table, long long table_len, long long* step_ptr, float beta2,
                      float eps, float clip, float grad_scale, int advance, void* stream) {
  CB_REQUIRE(n >= 0 && table_len > 0, "cb_clipped_adam_step: bad sizes n=%lld table_len=%lld", n, table_len);
  if (n == 0) {
    if (advance) cb::advance_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_ptr);
    return cb::check_launch("cb_clipped_adam_step");
  }
  for (const void* q : {(const void*)p, (const void*)g, (const void*)m, (const void*)v})
    CB_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0, "cb_clipped_adam_step: buffers must be 16-byte aligned");
  const long long n4 = (n + 3) / 4;
  const long long blocks = (n4 + cb::kAdamTile - 1) / cb::kAdamTile;  // one short-lived CTA per tile
  CB_REQUIRE(blocks < (1LL << 31), "cb_clipped_adam_step: n=%lld too large for one launch", n);
  cb::clipped_adam_kernel<<<unsigned(blocks), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr_table, beta1_table, step_table,
                                                                              table_len, step_ptr, beta2, eps, clip, grad_scale);
  if (advance) cb::advance_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_ptr);
  return cb::check_launch("cb_clipped_adam_step");
}

// ClippedAdam on this rank's shard with the data-parallel exchange fused in (see clipped_adam_p2p_kernel).
// grad_ptrs[n_grad]: the shard's position inside every rank's gradient buffer, rank order (host array of device / peer
// pointers); peer_param_ptrs[n_peer]: the shard's position inside the other ranks' parameter buffers.  n % 4 == 0.
template <int NG, int STAGES>
static int launch_p2p_bulk(float* p, float* m, float* v, long long n, const cb::P2PPointers& q, const float* lr_table,
                           const float* beta1_table, long long table_len, const long long* step_ptr, float beta2, float eps,
                           float clip, cudaStream_t st) {
  auto kern = cb::clipped_adam_p2p_bulk_kernel<NG, STAGES>;
  const int n_remote = q.n_grad - 1;
  const int smem = STAGES * (n_remote > 0 ? n_remote : 1) * cb::kP2PTile * 16;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 3;
  int per_sm = 0;
  const int n_sm = cb::sm_count();
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem) != cudaSuccess || per_sm < 1) return 4;
  const long long n_tiles = (n / 4 + cb::kP2PTile - 1) / cb::kP2PTile;
  long long blocks = (long long)n_sm * per_sm;  // persistent: every CTA resident, tiles strided over them
  if (blocks > n_tiles) blocks = n_tiles;
  kern<<<unsigned(blocks), 256, smem, st>>>(p, m, v, n, q, lr_table, beta1_table, table_len, step_ptr, beta2, eps, clip);
  return 0;
}

extern "C" int cb_clipped_adam_p2p(float* p, float* m, float* v, long long n, const float* const* grad_ptrs, int n_grad,
                                   int own, float* const* peer_param_ptrs, int n_peer, const float* lr_table,
                                   const float* beta1_table, long long table_len, long long* step_ptr, float beta2,
                                   float eps, float clip, int advance, void* stream) {
  CB_REQUIRE(n >= 0 && (n & 3) == 0 && table_len > 0, "cb_clipped_adam_p2p: n=%lld must be a non-negative multiple of 4", n);
  CB_REQUIRE(n_grad >= 1 && n_grad <= cb::kMaxRanks && n_peer >= 0 && n_peer < n_grad + (n_grad == 1),
             "cb_clipped_adam_p2p: %d gradient sources / %d peers (at most %d ranks)", n_grad, n_peer, cb::kMaxRanks);
  CB_REQUIRE(own >= 0 && own < n_grad, "cb_clipped_adam_p2p: own=%d outside [0, %d)", own, n_grad);
  cb::P2PPointers q{};
  q.n_grad = n_grad, q.n_peer = n_peer, q.own = own;
  for (int k = 0; k < n_grad; ++k) q.grad[k] = grad_ptrs[k];
  for (int k = 0; k < n_peer; ++k) q.peer_param[k] = peer_param_ptrs[k];
  for (const void* x : {(const void*)p, (const void*)m, (const void*)v})
    CB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "cb_clipped_adam_p2p: buffers must be 16-byte aligned");
  for (int k = 0; k < n_grad; ++k)
    CB_REQUIRE(q.grad[k] != nullptr && (reinterpret_cast<uintptr_t>(q.grad[k]) & 15) == 0, "cb_clipped_adam_p2p: bad gradient pointer %d", k);
  for (int k = 0; k < n_peer; ++k)
    CB_REQUIRE(q.peer_param[k] != nullptr && (reinterpret_cast<uintptr_t>(q.peer_param[k]) & 15) == 0, "cb_clipped_adam_p2p: bad peer pointer %d", k);
  cudaStream_t st = (cudaStream_t)stream;
  if (n > 0) {  // TMA bulk prefetch of the peers' gradient tiles into a shared-memory ring
    int rc = 0;
    if (n_grad <= 2) rc = launch_p2p_bulk<2, 8>(p, m, v, n, q, lr_table, beta1_table, table_len, step_ptr, beta2, eps, clip, st);
    else if (n_grad <= 4) rc = launch_p2p_bulk<4, 8>(p, m, v, n, q, lr_table, beta1_table, table_len, step_ptr, beta2, eps, clip, st);
    else if (n_grad <= 8) rc = launch_p2p_bulk<8, 4>(p, m, v, n, q, lr_table, beta1_table, table_len, step_ptr, beta2, eps, clip, st);
    else rc = launch_p2p_bulk<16, 3>(p, m, v, n, q, lr_table, beta1_table, table_len, step_ptr, beta2, eps, clip, st);
    CB_REQUIRE(rc == 0, "cb_clipped_adam_p2p: kernel configuration failed (code %d)", rc);
  }
  if (advance) cb::advance_step_kernel<<<1, 1, 0, st>>>(step_ptr);
  return cb::check_launch("cb_clipped_adam_p2p");
}

// The NVLS form of cb_clipped_adam_p2p: mc_grad / mc_param are the multicast addresses of this rank's shard inside the
// symmetric gradient / parameter buffers; p / m / v the local shard.  ld_mode / st_mode: see clipped_adam_nvls_kernel
// (0 / 0 = in-switch reduction + multicast store); local_grad: this rank's own gradient shard (ld_mode 1);
// peer_param_ptrs[n_peer]: the shard inside the other ranks' parameter buffers (st_mode 1); blocks_per_sm: grid cap.
extern "C" int cb_clipped_adam_nvls_ex(float* p, float* m, float* v, long long n, const float* mc_grad, float* mc_param,
                                       const float* local_grad, float* const* peer_param_ptrs, int n_peer, int ld_mode,
                                       int st_mode, int blocks_per_sm, const float* lr_table, const float* beta1_table,
                                       long long table_len, long long* step_ptr, float beta2, float eps, float clip,
                                       int advance, void* stream) {
  CB_REQUIRE(n >= 0 && (n & 3) == 0 && table_len > 0, "cb_clipped_adam_nvls: n=%lld must be a non-negative multiple of 4", n);
  CB_REQUIRE(ld_mode >= 0 && ld_mode <= 1 && st_mode >= 0 && st_mode <= 2 && blocks_per_sm >= 1,
             "cb_clipped_adam_nvls: bad mode %d / %d / %d", ld_mode, st_mode, blocks_per_sm);
  CB_REQUIRE((ld_mode != 0 || mc_grad != nullptr) && (st_mode != 0 || mc_param != nullptr),
             "cb_clipped_adam_nvls: multicast pointers must not be NULL");
  CB_REQUIRE(ld_mode != 1 || local_grad != nullptr, "cb_clipped_adam_nvls: ld_mode 1 needs the local gradient");
  CB_REQUIRE(n_peer >= 0 && n_peer < cb::kMaxRanks && (st_mode != 1 || n_peer == 0 || peer_param_ptrs != nullptr),
             "cb_clipped_adam_nvls: bad peer list (%d)", n_peer);
  for (const void* x : {(const void*)p, (const void*)m, (const void*)v, (const void*)mc_grad, (const void*)mc_param,
                        (const void*)local_grad})
    CB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "cb_clipped_adam_nvls: buffers must be 16-byte aligned");
  cb::NvlsPeers peers{};
  peers.n_peer = st_mode == 1 ? n_peer : 0;
  for (int k = 0; k < peers.n_peer; ++k) {
    peers.param[k] = peer_param_ptrs[k];
    CB_REQUIRE(peers.param[k] != nullptr && (reinterpret_cast<uintptr_t>(peers.param[k]) & 15) == 0,
               "cb_clipped_adam_nvls: bad peer pointer %d", k);
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (n > 0) {
    long long blocks = (n / 4 + 256 * cb::kNvlsUnroll - 1) / (256 * cb::kNvlsUnroll);
    const long long cap = (long long)blocks_per_sm * cb::sm_count();
    if (blocks > cap) blocks = cap;
#define CB_NVLS_LAUNCH(LD, ST)                                                                                              \
  cb::clipped_adam_nvls_kernel<LD, ST><<<unsigned(blocks), 256, 0, st>>>(p, m, v, n, mc_grad, mc_param, local_grad, peers,   \
                                                                         lr_table, beta1_table, table_len, step_ptr, beta2, \
                                                                         eps, clip)
    if (ld_mode == 0 && st_mode == 0) CB_NVLS_LAUNCH(0, 0);
    else if (ld_mode == 0 && st_mode == 1) CB_NVLS_LAUNCH(0, 1);
    else if (ld_mode == 0 && st_mode == 2) CB_NVLS_LAUNCH(0, 2);
    else if (ld_mode == 1 && st_mode == 0) CB_NVLS_LAUNCH(1, 0);
    else if (ld_mode == 1 && st_mode == 1) CB_NVLS_LAUNCH(1, 1);
    else CB_NVLS_LAUNCH(1, 2);
#undef CB_NVLS_LAUNCH
  }
  if (advance) cb::advance_step_kernel<<<1, 1, 0, st>>>(step_ptr);
  return cb::check_launch("cb_clipped_adam_nvls");
}

extern "C" int cb_clipped_adam_nvls(float* p, float* m, float* v, long long n, const float* mc_grad, float* mc_param,
                                    const float* lr_table, const float* beta1_table, long long table_len,
                                    long long* step_ptr, float beta2, float eps, float clip, int advance, void* stream) {
  return cb_clipped_adam_nvls_ex(p, m, v, n, mc_grad, mc_param, nullptr, nullptr, 0, 0, 0, 6, lr_table, beta1_table, table_len,
                                 step_ptr, beta2, eps, clip, advance, stream);
}
