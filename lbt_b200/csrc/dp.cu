// The data-parallel end of a training step as ONE kernel over NVLink peer memory (the reference is
// single-device, so this replaces what a port would do with two NCCL all-reduces + three small kernels):
//
//   barrier A (all replicas have finished their backward)                          flags over peer stores
//   reduce-scatter:  g[i] = grad_0[i] + grad_1[i] + ... (fixed rank order) for the slice this replica owns   peer loads
//   momentum SGD on the slice (trainer.py:81-82)  ->  all-gather: the updated weights are stored into every replica   peer stores
//   overflow counters summed over the replicas, range controller (dfxp:84-94) — every replica computes the same ranges
//   barrier B (everybody is done reading my gradients / counters and writing my weights), counters zeroed, step += 1
//
// The gradient never makes a round trip: each element is read once per replica by its owner and only the updated weight
// travels back (2·(N-1)/N · 4 B per parameter over NVLink, the minimum of a reduce-scatter + all-gather), the optimizer
// runs on 1/N of the parameters per replica, and all replicas hold bit-identical weights and ranges by construction
// (one owner computes each value; sums are taken in rank order, not arrival order).  Momentum lives only at the owner.
//
// Every cross-GPU wait is bounded by a %globaltimer watchdog that raises pad[LBT_DP_PAD_ERROR] instead of hanging.
#include "common.cuh"

#include <cstring>

namespace lbt {
namespace {

constexpr unsigned long long kTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;   // first-step skew between ranks can be seconds

struct DpArgs {
  lbt_dp_peers peers;
  float* accum;
  size_t n;
  float lr;
  const float* dev_lr;
  float momentum;
  int shard;
  int32_t* ranges;
  const int32_t* bits;
  const float* target;
  size_t n_sites;
  unsigned long long* dev_step;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {   // strong load: served by the owner's L2, never a stale L1 line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_peer_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Threads 0..world-1 of the calling CTA each wait for one replica's flag to reach `epoch`.  Returns false (CTA-uniform)
// when a wait timed out or another CTA already raised the sticky error word: the caller must then leave the step
// untouched (no optimizer, no all-gather, no controller) — summing a slow replica's half-written gradients and storing
// the result into every peer would corrupt all replicas silently.
__device__ __forceinline__ bool wait_flags(uint32_t* pad, int slot0, int world, uint32_t epoch) {
  int bad = 0;
  if ((int)threadIdx.x < world) {
    const uint32_t* f = pad + slot0 + threadIdx.x;
    const unsigned long long t0 = globaltimer();
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
      if ((++spins & 0xff) == 0 && *reinterpret_cast<volatile uint32_t*>(pad + LBT_DP_PAD_ERROR) != 0u) {
        bad = 1;
        break;
      }
      if (globaltimer() - t0 > kTimeoutNs) {
        atomicCAS(pad + LBT_DP_PAD_ERROR, 0u, 1u + (uint32_t)slot0 + threadIdx.x);
        bad = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  return __syncthreads_or(bad) == 0;
}

int g_dp_max_blocks = 0;   // lbt_dp_tune: grid cap per replica (0 = 4 CTAs per SM)

__device__ __forceinline__ void dp_step_body(const DpArgs& p) {
  const lbt_dp_peers& pe = p.peers;
  const int world = pe.world, rank = pe.rank;
  uint32_t* pad = pe.pad[rank];
  // the error word is sticky: once a cross-replica wait has timed out this replica's state is no longer trustworthy and
  // every later step is a no-op (Trainer raises on DpExchange.error())
  if (*reinterpret_cast<volatile uint32_t*>(pad + LBT_DP_PAD_ERROR) != 0u) return;
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(pad + LBT_DP_PAD_EPOCH) + 1u;

  // ---- barrier A: my gradients and counters are final (kernel boundary), tell everybody; wait for everybody ----
  if (world > 1) {
    if (blockIdx.x == 0 && (int)threadIdx.x < world) {
      __threadfence_system();
      st_release_sys(pe.pad[threadIdx.x] + LBT_DP_PAD_READY + rank, epoch);
    }
    if (!wait_flags(pad, LBT_DP_PAD_READY, world, epoch)) return;   // nothing of this step is applied
  }

  // ---- owned slice: sum the replicas' gradients in rank order, momentum SGD, publish the new weights ----
  const float lr = p.dev_lr ? *p.dev_lr : p.lr;
  const float scale = 1.0f / (float)world;
  const size_t nv = p.n / 4;
  size_t lo = 0, hi = nv;
  if (p.shard && world > 1) {
    const size_t chunk = (nv + world - 1) / world;
    lo = (size_t)rank * chunk;
    hi = lo + chunk < nv ? lo + chunk : nv;
    if (lo > nv) lo = nv;
  }
  float* w = pe.w[rank];
  for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (size_t)gridDim.x * blockDim.x) {
    float4 t[LBT_DP_MAX_WORLD];                       // all replicas' loads in flight before the first add
#pragma unroll
    for (int r = 0; r < LBT_DP_MAX_WORLD; ++r)
      if (r < world) t[r] = ld_peer_f4(pe.grad[r] + 4 * i);
    float4 g = t[0];
#pragma unroll
    for (int r = 1; r < LBT_DP_MAX_WORLD; ++r)
      if (r < world) {
        g.x = __fadd_rn(g.x, t[r].x);
        g.y = __fadd_rn(g.y, t[r].y);
        g.z = __fadd_rn(g.z, t[r].z);
        g.w = __fadd_rn(g.w, t[r].w);
      }
    float4 wv = reinterpret_cast<float4*>(w)[i], av = reinterpret_cast<float4*>(p.accum)[i];
    av.x = __fadd_rn(__fmul_rn(p.momentum, av.x), __fmul_rn(scale, g.x));
    av.y = __fadd_rn(__fmul_rn(p.momentum, av.y), __fmul_rn(scale, g.y));
    av.z = __fadd_rn(__fmul_rn(p.momentum, av.z), __fmul_rn(scale, g.z));
    av.w = __fadd_rn(__fmul_rn(p.momentum, av.w), __fmul_rn(scale, g.w));
    wv.x = __fsub_rn(wv.x, __fmul_rn(lr, av.x));
    wv.y = __fsub_rn(wv.y, __fmul_rn(lr, av.y));
    wv.z = __fsub_rn(wv.z, __fmul_rn(lr, av.z));
    wv.w = __fsub_rn(wv.w, __fmul_rn(lr, av.w));
    reinterpret_cast<float4*>(p.accum)[i] = av;
    reinterpret_cast<float4*>(w)[i] = wv;
    if (p.shard)
      for (int r = 0; r < world; ++r)
        if (r != rank) reinterpret_cast<float4*>(pe.w[r])[i] = wv;
  }

  // ---- overflow counters of the global batch -> range controller (every replica computes the same decision) ----
  if (blockIdx.x == gridDim.x - 1) {
    for (size_t i = threadIdx.x; i < p.n_sites; i += blockDim.x) {
      unsigned long long over = 0, half = 0, numel = 0;
      for (int r = 0; r < world; ++r) {
        const unsigned long long* c = reinterpret_cast<const unsigned long long*>(pe.counters[r]) + i * LBT_CNT_WORDS;
        over += ld_peer_u64(c + LBT_CNT_OVER);
        half += ld_peer_u64(c + LBT_CNT_OVER_HALF);
        numel += ld_peer_u64(c + LBT_CNT_NUMEL);
      }
      if (numel != 0ull) {
        const float nn = (float)numel;
        const float r1 = __fdiv_rn((float)over, nn), r2 = __fdiv_rn((float)half, nn);
        const float t = p.target ? p.target[i] : 0.f;
        const int delta = (r1 > t) ? 1 : ((r2 <= t) ? -1 : 0);
        p.ranges[i] = min(p.bits[i] - 1, p.ranges[i] + delta);
      }
    }
  }

  // ---- the last CTA to finish runs barrier B and closes the step ----
  // (bar.sync orders the CTA's accesses before thread 0's fence, which is cumulative: one fence per CTA, not per thread)
  __shared__ uint32_t s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (world > 1) __threadfence_system(); else __threadfence();
    const uint32_t t = atomicAdd(pad + LBT_DP_PAD_TICKET, 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
    if (world > 1) __threadfence_system(); else __threadfence();
  }
  __syncthreads();
  if (!s_last) return;
  if (world > 1) {
    if ((int)threadIdx.x < world) st_release_sys(pe.pad[threadIdx.x] + LBT_DP_PAD_DONE + rank, epoch);
    if (!wait_flags(pad, LBT_DP_PAD_DONE, world, epoch)) return;    // peers may still be using my buffers: do not close the step
  }
  unsigned long long* mine = reinterpret_cast<unsigned long long*>(const_cast<uint64_t*>(pe.counters[rank]));
  for (size_t i = threadIdx.x; i < p.n_sites * LBT_CNT_WORDS; i += blockDim.x) mine[i] = 0ull;
  if (threadIdx.x == 0) {
    pad[LBT_DP_PAD_TICKET] = 0u;
    pad[LBT_DP_PAD_EPOCH] = epoch;
    if (p.dev_step) *p.dev_step += 1ull;
  }
}

__global__ void __launch_bounds__(256) dp_step_kernel(const DpArgs p) {
  pdl_trigger();   // programmatic dependent launch: the next kernel may be scheduled now ...
  pdl_wait();      // ... and this one touches global memory only after its predecessor has completed
  dp_step_body(p);
}

// All replicas of a SIMULATED world in ONE cooperative launch (tests on a single GPU): grid.y = replica, every slice of the
// grid runs the unmodified body against its own arguments, and the cooperative launch guarantees what the protocol needs —
// that all replicas' CTAs are resident at the same time.  (Separate launches that spin on each other's flags have no such
// guarantee on one GPU.)
__global__ void __launch_bounds__(256) dp_step_multi_kernel(const DpArgs* all) { dp_step_body(all[blockIdx.y]); }

typedef int (*GetAddressRangeFn)(unsigned long long*, size_t*, unsigned long long);

}  // namespace
}  // namespace lbt

using namespace lbt;

static int dp_make_args(DpArgs& a, size_t& blocks, const lbt_dp_peers* peers, float* accum, size_t n, float lr, const float* dev_lr,
                        float momentum, int shard, int32_t* ranges, const int32_t* bits, const float* target, size_t n_sites,
                        uint64_t* dev_step) {
  if (!peers || !accum) return LBT_EINVAL;
  const int world = peers->world, rank = peers->rank;
  if (world < 1 || world > LBT_DP_MAX_WORLD || rank < 0 || rank >= world) return LBT_EINVAL;
  if (n_sites && (!ranges || !bits)) return LBT_EINVAL;
  for (int r = 0; r < world; ++r) {
    if (!peers->grad[r] || !peers->pad[r] || (n_sites && !peers->counters[r])) return LBT_EINVAL;
    if ((shard || r == rank) && !peers->w[r]) return LBT_EINVAL;
    if ((reinterpret_cast<uintptr_t>(peers->grad[r]) | reinterpret_cast<uintptr_t>(peers->w[r])) & 15) return LBT_EUNSUPPORTED;
  }
  if ((n & 3) || (reinterpret_cast<uintptr_t>(accum) & 15)) return LBT_EUNSUPPORTED;   // flat buffers are padded to float4
  LBT_REQUIRE_ARCH();
  const DeviceInfo& di = device_info();
  a.peers = *peers;
  a.accum = accum;
  a.n = n;
  a.lr = lr;
  a.dev_lr = dev_lr;
  a.momentum = momentum;
  a.shard = (shard && world > 1) ? 1 : 0;
  a.ranges = ranges;
  a.bits = bits;
  a.target = target;
  a.n_sites = n_sites;
  a.dev_step = reinterpret_cast<unsigned long long*>(dev_step);
  const size_t mine = a.shard ? (n / 4 + world - 1) / world : n / 4;
  blocks = (mine + 255) / 256;
  size_t cap = (size_t)di.sm_count * 4;     // peer loads need many requests in flight; all CTAs co-resident
  if (g_dp_max_blocks > 0 && (size_t)g_dp_max_blocks < cap) cap = (size_t)g_dp_max_blocks;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return LBT_OK;
}

extern "C" int lbt_dp_step(const lbt_dp_peers* peers, float* accum, size_t n, float lr, const float* dev_lr, float momentum,
                           int shard, int32_t* ranges, const int32_t* bits, const float* target, size_t n_sites,
                           uint64_t* dev_step, void* stream) {
  DpArgs a;
  size_t blocks = 1;
  const int rc = dp_make_args(a, blocks, peers, accum, n, lr, dev_lr, momentum, shard, ranges, bits, target, n_sites, dev_step);
  if (rc) return rc;
  launch_pdl(dp_step_kernel, (unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream), a);
  return check_launch("lbt_dp_step");
}

// Test entry point (not in lbt.h): the steps of `world` replicas that all live on THIS GPU as one cooperative launch.
// peers[r] / accum[r] / ranges[r] / dev_step[r] are replica r's arguments of lbt_dp_step; `scratch` is caller-owned device
// memory of at least lbt_dp_emulate_scratch_bytes(world) bytes.
extern "C" size_t lbt_dp_emulate_scratch_bytes(int world) { return (size_t)(world > 0 ? world : 0) * sizeof(DpArgs); }

extern "C" int lbt_dp_step_emulate(int world, const lbt_dp_peers* peers, float* const* accum, size_t n, float lr,
                                   const float* const* dev_lr, float momentum, int shard, int32_t* const* ranges, const int32_t* bits,
                                   const float* target, size_t n_sites, uint64_t* const* dev_step, void* scratch, void* stream) {
  if (world < 1 || world > LBT_DP_MAX_WORLD || !peers || !accum || !scratch) return LBT_EINVAL;
  DpArgs all[LBT_DP_MAX_WORLD];
  size_t blocks = 1;
  for (int r = 0; r < world; ++r) {
    if (peers[r].world != world || peers[r].rank != r) return LBT_EINVAL;
    size_t b = 1;
    const int rc = dp_make_args(all[r], b, &peers[r], accum[r], n, lr, dev_lr ? dev_lr[r] : nullptr, momentum, shard,
                                ranges ? ranges[r] : nullptr, bits, target, n_sites, dev_step ? dev_step[r] : nullptr);
    if (rc) return rc;
    blocks = b;   // same n, world, shard for every replica
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dp_step_multi_kernel, 256, 0) != cudaSuccess || occ < 1) occ = 1;
  const size_t resident = (size_t)occ * (size_t)device_info().sm_count;
  if (blocks * (size_t)world > resident) blocks = resident / (size_t)world;
  if (blocks < 1) return LBT_EUNSUPPORTED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemcpyAsync(scratch, all, sizeof(DpArgs) * world, cudaMemcpyHostToDevice, st) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "lbt_dp_step_emulate(memcpy)");
    return LBT_ECUDA;
  }
  const DpArgs* dev_all = reinterpret_cast<const DpArgs*>(scratch);
  void* kargs[] = {(void*)&dev_all};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)dp_step_multi_kernel, dim3((unsigned)blocks, (unsigned)world, 1), dim3(256, 1, 1),
                                              kargs, 0, st);
  if (e != cudaSuccess) {
    set_cuda_error(e, "lbt_dp_step_emulate");
    (void)cudaGetLastError();
    return LBT_ECUDA;
  }
  return check_launch("lbt_dp_step_emulate");
}

// Tuning / test knob (not in lbt.h): cap the grid of lbt_dp_step.  Replicas wait for each other inside the kernel, so N
// replicas simulated on ONE GPU (tests) need all their CTAs co-resident.
extern "C" int lbt_dp_tune(int max_blocks) {
  g_dp_max_blocks = max_blocks > 0 ? max_blocks : 0;
  return LBT_OK;
}

// Export a caller-owned device buffer to the other replicas (one process per GPU): the CUDA IPC handle of the allocation
// that contains `ptr` and ptr's byte offset inside it.  The caller ships (handle, offset) over its own control plane.
extern "C" int lbt_dp_export(const void* ptr, void* handle64, size_t* offset) {
  if (!ptr || !handle64 || !offset) return LBT_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == LBT_DP_HANDLE_BYTES, "IPC handle size");
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !f) {
    set_cuda_error(cudaErrorNotSupported, "cuMemGetAddressRange entry point");
    return LBT_ECUDA;
  }
  unsigned long long base = 0;
  size_t size = 0;
  if (reinterpret_cast<GetAddressRangeFn>(f)(&base, &size, (unsigned long long)reinterpret_cast<uintptr_t>(ptr)) != 0) {
    set_cuda_error(cudaErrorInvalidValue, "cuMemGetAddressRange");
    return LBT_ECUDA;
  }
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(base)));
  if (e != cudaSuccess) {
    set_cuda_error(e, "cudaIpcGetMemHandle");
    (void)cudaGetLastError();
    return LBT_ECUDA;
  }
  memcpy(handle64, &h, sizeof(h));
  *offset = (size_t)(reinterpret_cast<uintptr_t>(ptr) - base);
  return LBT_OK;
}

// Map a peer replica's exported allocation into this process (enables peer access over NVLink); *base is the peer
// allocation's base — add the exported offset.  lbt_dp_close unmaps it.
extern "C" int lbt_dp_open(const void* handle64, void** base) {
  if (!handle64 || !base) return LBT_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_cuda_error(e, "cudaIpcOpenMemHandle");
    (void)cudaGetLastError();
    return LBT_ECUDA;
  }
  return LBT_OK;
}

extern "C" int lbt_dp_close(void* base) {
  if (!base) return LBT_EINVAL;
  cudaError_t e = cudaIpcCloseMemHandle(base);
  if (e != cudaSuccess) {
    set_cuda_error(e, "cudaIpcCloseMemHandle");
    (void)cudaGetLastError();
    return LBT_ECUDA;
  }
  return LBT_OK;
}
