// Multi-GPU residual sharding: one process per GPU, correspondences block-partitioned over the ranks,
// and the packed normal equations [J^T J | J^T r | cost] summed across ranks once per linearisation.
//
// The reduction is a one-shot all-reduce over NVLink peer memory inside our own kernel (no host round trip, no library
// call on the iteration path): every rank's linearisation kernels accumulate into that rank's exchange buffer, which is
// exported with cudaIpcGetMemHandle and mapped by all peers.  comm_allreduce then
//   1. publishes "my partial for epoch e is complete" by storing e into flag[my_rank] of every peer (remote st.global),
//   2. waits until all of its own flags reached e (local spin, bounded by a clock64 timeout so a dead peer can never
//      hang the GPU),
//   3. sums the partials of ranks 0..G-1 in rank order with peer ld.volatile.global loads — the same order on every
//      rank, so all ranks hold bitwise identical normal equations and take bitwise identical LM steps (the replicated
//      lm_* kernels then need no broadcast).
// Partials are double-buffered by epoch parity: a rank can start epoch e+2 only after every peer signalled e+1, i.e.
// after every peer finished reading epoch e, so one flag round per reduction suffices.
#include "wc_ctx.h"

void wc_solve_exchange_views(wc_ctx* c, int which, double** H, double** g, double** cost, int* N, void** state);

namespace {

struct CommArgs {
  unsigned long long* my_flags;        // local: flags[rank of writer]
  unsigned long long* peer_flags[8];   // remote flag arrays
  const double*       part[8];         // partial buffers of this epoch's parity, by rank (own entry local)
  double*             my_other;        // own partial of the other parity: zeroed here for the linearisation after next
  double*             H[2];
  double*             g[2];
  double*             cost[2];
  const int*          state;           // LMState: cur @0, done @1, ..., step_valid @4 (see wc_solve.cu)
  int*                err;
  unsigned long long  epoch;
  int                 rank, world, N, at_candidate;
  long long           timeout_cycles;
};

__global__ void __launch_bounds__(256) comm_allreduce(CommArgs a) {
  const int cur = a.state[0], done = a.state[1];
  if (done) return;  // replicated state: every rank takes the same branch.  (After an invalid step the linearisation wrote
                     // nothing: the reduction then sums zeros, which keeps the flag / parity protocol in step.)
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();
    *((volatile unsigned long long*)&a.peer_flags[threadIdx.x][a.rank]) = a.epoch;
  }
  if (threadIdx.x < a.world) {
    const long long t0 = clock64();
    while (*((volatile unsigned long long*)&a.my_flags[threadIdx.x]) < a.epoch) {
      if (clock64() - t0 > a.timeout_cycles) {
        *a.err = WC_ECOMM;
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
  const int    buf   = a.at_candidate ? 1 - cur : cur;
  const size_t nH    = (size_t)a.N * a.N;
  // only the lower triangle of J^T J is ever accumulated or read: packed index k -> (row, col), then gradient and cost
  const size_t nL    = (size_t)a.N * (a.N + 1) / 2;
  const size_t total = nL + a.N + 1;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    size_t i;
    if (k < nL) {
      size_t row = (size_t)((sqrt(8.0 * (double)k + 1.0) - 1.0) * 0.5);
      while (row * (row + 1) / 2 > k) --row;
      while ((row + 1) * (row + 2) / 2 <= k) ++row;
      i = row * a.N + (k - row * (row + 1) / 2);
    } else {
      i = nH + (k - nL);
    }
    double s = 0.0;
    for (int r = 0; r < a.world; ++r) s += *((const volatile double*)&a.part[r][i]);
    if (i < nH) a.H[buf][i] = s;
    else if (i < nH + a.N) a.g[buf][i - nH] = s;
    else *a.cost[buf] = s;
    // every peer has signalled this epoch, i.e. finished reading the previous one: the other-parity partial is free and
    // is cleared here for the next linearisation (no separate zeroing launch on the iteration path)
    a.my_other[i] = 0.0;
  }
}

// start of a sharded solve: one flag round, then both own partials are cleared (nobody can still be reading them)
__global__ void __launch_bounds__(256) comm_barrier_zero(CommArgs a, double* p0, double* p1, size_t n) {
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();
    *((volatile unsigned long long*)&a.peer_flags[threadIdx.x][a.rank]) = a.epoch;
  }
  if (threadIdx.x < a.world) {
    const long long t0 = clock64();
    while (*((volatile unsigned long long*)&a.my_flags[threadIdx.x]) < a.epoch) {
      if (clock64() - t0 > a.timeout_cycles) {
        *a.err = WC_ECOMM;
        break;
      }
    }
  }
  __syncthreads();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p0[i] = 0.0, p1[i] = 0.0;
}

// All-gather of per-query rows (the k-NN index lists of the sharded matcher): rank r computed rows [row0[r], row0[r+1])
// into ITS exported gather region; after one flag round every rank copies all slices into its local array.
struct GatherArgs {
  unsigned long long* my_flags;       // local: gather flags[rank of writer]
  unsigned long long* peer_flags[8];
  const int*          region[8];      // gather regions of this epoch's parity, by rank
  int*                dst;            // local full array
  int*                err;
  unsigned long long  epoch;
  int                 rank, world, width;  // ints per row
  int                 row0[9];
  long long           timeout_cycles;
};

__global__ void __launch_bounds__(256) comm_allgather_rows(GatherArgs a) {
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();
    *((volatile unsigned long long*)&a.peer_flags[threadIdx.x][a.rank]) = a.epoch;
  }
  if (threadIdx.x < a.world) {
    const long long t0 = clock64();
    while (*((volatile unsigned long long*)&a.my_flags[threadIdx.x]) < a.epoch) {
      if (clock64() - t0 > a.timeout_cycles) {
        *a.err = WC_ECOMM;
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
  const size_t total = (size_t)a.row0[a.world] * a.width;
  auto rank_of = [&](size_t i) {
    const int row = (int)(i / a.width);
    int       r   = 0;
    while (row >= a.row0[r + 1]) ++r;
    return r;
  };
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if ((a.width & 1) == 0) {
    // even row width: every slice starts on an 8-byte boundary; 8-byte peer loads, four in flight per thread
    const size_t n2 = total >> 1;
    int2*        d2 = reinterpret_cast<int2*>(a.dst);
    for (size_t i0 = tid; i0 < n2; i0 += 4 * nth) {
      unsigned long long v[4];  // (the flag round and the fences above order these loads after the peers' writes)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t i = i0 + u * nth;
        if (i < n2) v[u] = *((const volatile unsigned long long*)(reinterpret_cast<const unsigned long long*>(a.region[rank_of(2 * i)]) + i));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t i = i0 + u * nth;
        if (i < n2) reinterpret_cast<unsigned long long*>(d2)[i] = v[u];
      }
    }
  } else {
    for (size_t i = tid; i < total; i += nth) a.dst[i] = *((const volatile int*)&a.region[rank_of(i)][i]);
  }
}

// flag round of the sharded sweep upload: "my slab of this epoch is in my raw area" -> every peer; wait for all peers
struct ReadyArgs {
  unsigned long long* my_flags;
  unsigned long long* peer_flags[8];
  int*                err;
  unsigned long long  epoch;
  int                 rank, world;
  long long           timeout_cycles;
};
__global__ void comm_raw_ready(ReadyArgs a) {
  if (threadIdx.x < a.world) {
    __threadfence_system();
    *((volatile unsigned long long*)&a.peer_flags[threadIdx.x][a.rank]) = a.epoch;
    const long long t0 = clock64();
    while (*((volatile unsigned long long*)&a.my_flags[threadIdx.x]) < a.epoch) {
      if (clock64() - t0 > a.timeout_cycles) {
        *a.err = WC_ECOMM;
        break;
      }
    }
    __threadfence_system();
  }
}

}  // namespace

static size_t part_doubles(const wc_ctx* c) {
  const size_t N = 12 * (size_t)c->prm.max_samples;
  return N * N + N + 2;
}

static size_t gather_ints(const wc_ctx* c) { return (size_t)c->prm.max_surfels * 16; }
static size_t gather_off(const wc_ctx* c) { return 256 + 2 * part_doubles(c) * 8; }  // byte offset of the gather flags

static size_t raw_off(const wc_ctx* c) { return (gather_off(c) + 256 + 2 * gather_ints(c) * 4 + 255) & ~(size_t)255; }  // byte offset of the raw flags
static size_t raw_area_bytes(const wc_ctx* c) { return (size_t)c->prm.max_points * 48; }

static wc_status comm_alloc(wc_ctx* c) {
  if (c->d_xchg) return WC_OK;
  // [256 B reduce flags][2 x partial normal equations][256 B gather flags][2 x gather region of max_surfels x 16 ints]
  // [256 B raw flags][2 x raw sweep area of max_points x 48 B]
  c->xchg_bytes = raw_off(c) + 256 + 2 * raw_area_bytes(c);
  WC_CUDA(c, cudaMalloc(&c->d_xchg, c->xchg_bytes));
  WC_CUDA(c, cudaMemset(c->d_xchg, 0, c->xchg_bytes));
  WC_CUDA(c, cudaMalloc(&c->d_comm_err, 4));
  WC_CUDA(c, cudaMemset(c->d_comm_err, 0, 4));
  return WC_OK;
}

void wc_comm_free(wc_ctx* c) {
  if (c->comm_ready) {
    for (int r = 0; r < c->world; ++r)
      if (r != c->rank && c->peer_xchg[r]) cudaIpcCloseMemHandle(c->peer_xchg[r]);
  }
  if (c->d_xchg) cudaFree(c->d_xchg);
  if (c->d_comm_err) cudaFree(c->d_comm_err);
  c->d_xchg = nullptr, c->d_comm_err = nullptr, c->comm_ready = 0, c->world = 1, c->rank = 0, c->shard_upload = 0;
}

extern "C" wc_status wc_comm_export(wc_ctx* c, uint8_t handle[WC_IPC_HANDLE_BYTES]) {
  if (!c || !handle) return WC_EINVAL;
  wc_status s = comm_alloc(c);
  if (s) return s;
  static_assert(sizeof(cudaIpcMemHandle_t) == WC_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  WC_CUDA(c, cudaIpcGetMemHandle(&h, c->d_xchg));
  memcpy(handle, &h, sizeof(h));
  return WC_OK;
}

extern "C" wc_status wc_comm_connect(wc_ctx* c, int rank, int world, const uint8_t* all_handles) {
  if (!c || !all_handles || world < 1 || world > 8 || rank < 0 || rank >= world) return WC_EINVAL;
  wc_status s = comm_alloc(c);
  if (s) return s;
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      c->peer_xchg[r] = c->d_xchg;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, all_handles + (size_t)r * WC_IPC_HANDLE_BYTES, sizeof(h));
    void*       p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) WC_FAIL(c, WC_ECOMM, "cudaIpcOpenMemHandle(rank %d) -> %s", r, cudaGetErrorString(e));
    c->peer_xchg[r] = (double*)p;
  }
  c->rank = rank, c->world = world, c->comm_ready = 1, c->comm_epoch = 0, c->gather_epoch = 0, c->raw_epoch = 0;
  return WC_OK;
}

// Sweep uploads (wc_points_upload / wc_points_prefetch) become collective calls: every rank passes the same sweep, copies
// only its 1 / world slab over its own PCIe link, and reads the other slabs from its peers over NVLink.
extern "C" wc_status wc_comm_shard_upload(wc_ctx* c, int on) {
  if (!c) return WC_EINVAL;
  if (on && (!c->comm_ready || c->world < 2)) WC_FAIL(c, WC_ECOMM, "wc_comm_connect has not been called");
  cudaStreamSynchronize(c->stream);
  c->prefetch_src = nullptr, c->defer_src = nullptr;  // a pending prefetch belongs to the other mode
  c->shard_upload = on ? 1 : 0;
  return WC_OK;
}

extern "C" wc_status wc_comm_disconnect(wc_ctx* c) {
  if (!c) return WC_EINVAL;
  cudaStreamSynchronize(c->stream);
  wc_comm_free(c);
  return WC_OK;
}

// where this rank's linearisation kernels must accumulate for the next reduction (world > 1)
void wc_comm_partial_views(wc_ctx* c, double** H, double** g, double** cost) {
  const size_t N   = 12 * c->K;
  const int    par = (int)((c->comm_epoch + 1) & 1);
  double*      p   = (double*)((char*)c->d_xchg + 256) + (size_t)par * part_doubles(c);
  *H = p, *g = p + N * N, *cost = p + N * N + N;
}

wc_status wc_comm_allreduce(wc_ctx* c, int at_candidate) {
  if (c->world <= 1) return WC_OK;
  if (!c->comm_ready) WC_FAIL(c, WC_ECOMM, "wc_comm_connect has not been called");
  c->comm_epoch += 1;
  CommArgs a;
  memset(&a, 0, sizeof(a));
  const int par = (int)(c->comm_epoch & 1);
  a.my_flags    = (unsigned long long*)c->d_xchg;
  for (int r = 0; r < c->world; ++r) {
    a.peer_flags[r] = (unsigned long long*)c->peer_xchg[r];
    a.part[r]       = (const double*)((char*)c->peer_xchg[r] + 256) + (size_t)par * part_doubles(c);
  }
  a.my_other = (double*)((char*)c->d_xchg + 256) + (size_t)(1 - par) * part_doubles(c);
  void* state = nullptr;
  wc_solve_exchange_views(c, 0, a.H, a.g, a.cost, &a.N, &state);
  a.state = (const int*)state, a.err = c->d_comm_err;
  a.epoch = c->comm_epoch, a.rank = c->rank, a.world = c->world, a.at_candidate = at_candidate;
  a.timeout_cycles = 4000000000ll;  // ~2 s at 2 GHz
  const size_t total = (size_t)a.N * (a.N + 1) / 2 + a.N + 1;
  int          grid  = (int)((total + 255) / 256);
  if (grid > 2 * c->num_sms) grid = 2 * c->num_sms;
  { ++c->n_launches; comm_allreduce<<<grid, 256, 0, c->stream>>>(a); }
  WC_CUDA(c, cudaGetLastError());
  return WC_OK;
}

// start of a sharded solve: barrier over the ranks + clear both partial buffers of this rank
wc_status wc_comm_begin_solve(wc_ctx* c) {
  if (c->world <= 1) return WC_OK;
  if (!c->comm_ready) WC_FAIL(c, WC_ECOMM, "wc_comm_connect has not been called");
  c->comm_epoch += 1;
  CommArgs a;
  memset(&a, 0, sizeof(a));
  a.my_flags = (unsigned long long*)c->d_xchg;
  for (int r = 0; r < c->world; ++r) a.peer_flags[r] = (unsigned long long*)c->peer_xchg[r];
  a.err = c->d_comm_err, a.epoch = c->comm_epoch, a.rank = c->rank, a.world = c->world;
  a.timeout_cycles = 4000000000ll;
  double* p0 = (double*)((char*)c->d_xchg + 256);
  { ++c->n_launches; comm_barrier_zero<<<64, 256, 0, c->stream>>>(a, p0, p0 + part_doubles(c), (size_t)(12 * c->K) * (12 * c->K) + 12 * c->K + 1); }
  WC_CUDA(c, cudaGetLastError());
  return WC_OK;
}

// where this rank writes its rows for the next all-gather (indexed like the full array), world > 1
int* wc_comm_gather_region(wc_ctx* c) {
  const int par = (int)((c->gather_epoch + 1) & 1);
  return (int*)((char*)c->d_xchg + gather_off(c) + 256) + (size_t)par * gather_ints(c);
}

// rows [row0[r], row0[r+1]) of `width` ints were written by rank r into its gather region; fills dst on every rank
wc_status wc_comm_allgather_rows(wc_ctx* c, int* dst, const int* row0, int width) {
  if (c->world <= 1) return WC_OK;
  if (!c->comm_ready) WC_FAIL(c, WC_ECOMM, "wc_comm_connect has not been called");
  if ((size_t)row0[c->world] * width > gather_ints(c)) WC_FAIL(c, WC_ECAPACITY, "gather region too small");
  c->gather_epoch += 1;
  GatherArgs a;
  memset(&a, 0, sizeof(a));
  const int par = (int)(c->gather_epoch & 1);
  a.my_flags    = (unsigned long long*)((char*)c->d_xchg + gather_off(c));
  for (int r = 0; r < c->world; ++r) {
    a.peer_flags[r] = (unsigned long long*)((char*)c->peer_xchg[r] + gather_off(c));
    a.region[r]     = (const int*)((char*)c->peer_xchg[r] + gather_off(c) + 256) + (size_t)par * gather_ints(c);
  }
  for (int r = 0; r <= c->world; ++r) a.row0[r] = row0[r];
  a.dst = dst, a.err = c->d_comm_err, a.epoch = c->gather_epoch, a.rank = c->rank, a.world = c->world, a.width = width;
  a.timeout_cycles = 4000000000ll;
  { ++c->n_launches; comm_allgather_rows<<<c->num_sms, 256, 0, c->stream>>>(a); }
  WC_CUDA(c, cudaGetLastError());
  return WC_OK;
}

// ---- sharded sweep upload (upload half): every rank copies only ITS slab of the raw sweep over PCIe into
// its exported raw area; after one flag round the repack kernel of every rank reads each point from its owner's area
// over NVLink (wc_extract.cu).  Areas are double-buffered by epoch parity like the reduction partials: a rank can write
// its area of parity p again only after every peer signalled the epoch in between, i.e. finished reading it.
// this rank's raw area of the given epoch's parity (device pointer, record 0 of the sweep at offset 0)
void* wc_comm_raw_area(wc_ctx* c, unsigned long long epoch) {
  return (char*)c->d_xchg + raw_off(c) + 256 + (size_t)(epoch & 1) * raw_area_bytes(c);
}
void wc_comm_raw_areas(wc_ctx* c, unsigned long long epoch, const void* areas[8]) {
  for (int r = 0; r < 8; ++r)
    areas[r] = r < c->world ? (const void*)((char*)c->peer_xchg[r] + raw_off(c) + 256 + (size_t)(epoch & 1) * raw_area_bytes(c)) : nullptr;
}
wc_status wc_comm_raw_ready(wc_ctx* c, unsigned long long epoch) {
  if (!c->comm_ready) WC_FAIL(c, WC_ECOMM, "wc_comm_connect has not been called");
  ReadyArgs a;
  memset(&a, 0, sizeof(a));
  a.my_flags = (unsigned long long*)((char*)c->d_xchg + raw_off(c));
  for (int r = 0; r < c->world; ++r) a.peer_flags[r] = (unsigned long long*)((char*)c->peer_xchg[r] + raw_off(c));
  a.err = c->d_comm_err, a.epoch = epoch, a.rank = c->rank, a.world = c->world;
  a.timeout_cycles = 4000000000ll;
  { ++c->n_launches; comm_raw_ready<<<1, 32, 0, c->stream>>>(a); }
  WC_CUDA(c, cudaGetLastError());
  return WC_OK;
}

wc_status wc_comm_check(wc_ctx* c) {
  if (c->world <= 1 || !c->d_comm_err) return WC_OK;
  int e = 0;
  WC_CUDA(c, cudaMemcpy(&e, c->d_comm_err, 4, cudaMemcpyDeviceToHost));
  if (e) {
    cudaMemset(c->d_comm_err, 0, 4);
    WC_FAIL(c, WC_ECOMM, "peer exchange timed out (a rank did not reach the reduction)");
  }
  return WC_OK;
}

// Timing hook for the exchange step alone (tools/allreduce_compare.py puts it beside ncclAllReduce on the same buffer
// size): `reps` reductions of the packed normal equations of a K-pose window, back to back on the ctx stream.  Every
// rank must call it with the same arguments.
wc_status wc_solve_exchange_reset(wc_ctx* c, size_t K);  // wc_solve.cu
extern "C" wc_status wc_comm_bench(wc_ctx* c, size_t K, int reps, double* ms_per_call) {
  if (!c || !ms_per_call || reps < 1) return WC_EINVAL;
  if (c->world <= 1 || !c->comm_ready) WC_FAIL(c, WC_ECOMM, "wc_comm_connect has not been called");
  wc_status s = wc_solve_exchange_reset(c, K);
  if (s) return s;
  if ((s = wc_comm_begin_solve(c))) return s;
  for (int i = 0; i < 3; ++i)
    if ((s = wc_comm_allreduce(c, 0))) return s;
  WC_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
  for (int i = 0; i < reps; ++i)
    if ((s = wc_comm_allreduce(c, 0))) return s;
  WC_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  WC_CUDA(c, cudaStreamSynchronize(c->stream));
  float ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  *ms_per_call = ms / reps;
  return wc_comm_check(c);
}
