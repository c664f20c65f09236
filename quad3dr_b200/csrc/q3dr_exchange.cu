// q3dr_exchange.cu -- the result exchange between GPUs behind the C ABI (include/q3dr_raycast.h, "result exchange").
//
// The path shards by viewpoint and needs no collective during compute; what remains is the
// gather of the per-viewpoint hit lists and scores.  It is overlapped with the compute: the scoring call hands every
// packed chunk to exchange_push_chunk, which enqueues copy-engine pushes (cudaMemcpyAsync into peer memory, one side
// stream per target rank so that pushes to different peers run on different copy engines at once) behind the chunk's
// "packed" event and returns at once; the SMs are already casting the next chunk.  The raycast kernel is issue-bound
// and owns every register of every SM, so an SM-driven collective could not overlap with it -- copy engines can.
// Completion is signalled through memory: behind its data every rank pushes one 8-byte flag per peer on that peer's
// stream (copies of one stream complete in order), and end_step waits for all flags with a single-warp kernel on the
// caller's stream.  No NCCL call, no host barrier; works across processes (CUDA IPC) and
// across threads of one process (peer access).
#include "q3dr_internal.cuh"

namespace {

constexpr uint64_t kAlign = 256;
inline uint64_t up(uint64_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

struct SlotHeader {
  uint64_t n_viewpoints, n_entries, step, status;
};

// One warp; lane r watches the flag of source rank r (+32, +64 ...).  Exits when every flag carries `step`, or when
// `timeout_ns` has passed (status 2), or when some rank reported a failure (status 1).
__global__ void exchange_wait_kernel(const volatile uint64_t* flags, uint32_t n_ranks, uint64_t step, uint64_t timeout_ns,
                                     uint32_t* status) {
  uint64_t t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  uint32_t st = 0;
  for (uint32_t r = threadIdx.x; r < n_ranks; r += 32) {
    while (true) {
      const uint64_t f = flags[r];
      if ((f >> 1) >= step) {
        if (f & 1ull) st = 1;
        break;
      }
      uint64_t t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {
        st = 2;
        break;
      }
      __nanosleep(500);
    }
  }
  st = __reduce_max_sync(0xFFFFFFFFu, st);
  if (threadIdx.x == 0) *status = st;
  __threadfence_system();
}

}  // namespace

struct q3dr_exchange {
  q3dr_map* map = nullptr;
  int device = 0;
  uint32_t n_ranks = 0, rank = 0;
  uint64_t max_vp = 0, max_entries = 0;
  q3dr_exchange_layout lay{};
  char* base[2] = {nullptr, nullptr};                // this rank's receive buffers
  std::vector<char*> peer[2];                        // every rank's receive buffers as seen from this device
  std::vector<bool> opened[2];                       // mapped through CUDA IPC (to be closed)
  std::vector<cudaStream_t> side;                    // one push stream per target rank: pushes to different peers overlap
  cudaEvent_t ev_tail = nullptr;                     // "header and flag of this step are staged in device memory"
  SlotHeader* h_header = nullptr;                    // pinned staging [2]
  uint64_t* h_flag = nullptr;                        // pinned staging [2]
  char* d_stage = nullptr;                           // device staging [2] x {SlotHeader, flag}: peers are written D2D only
  SlotHeader* h_all = nullptr;                       // pinned: the headers of all source ranks after end_step
  uint32_t* d_status = nullptr;
  uint32_t* h_status = nullptr;
  uint32_t cur = 1;
  uint64_t step = 0;
  bool connected = false, armed = false, valid = false;
  int push_error = Q3DR_OK;
  std::string push_message;
  uint64_t pushed_entries = 0;
};

namespace q3dr_detail {

int exchange_push_chunk(q3dr_exchange* ex, cudaEvent_t ready, const int32_t* d_leaf, const float* d_info,
                        uint64_t first_entry, uint64_t n_entries) {
  if (!ex->armed || ex->push_error != Q3DR_OK) return Q3DR_OK;  // a failed step is reported by end_step, on all ranks
  if (first_entry + n_entries > ex->max_entries) {
    ex->push_error = Q3DR_ERR_INVALID_ARG;
    ex->push_message = "result larger than the receive slot (max_entries of q3dr_exchange_create)";
    return Q3DR_OK;
  }
  if (n_entries == 0) return Q3DR_OK;
  cudaError_t e = cudaSuccess;
  const uint64_t slot = (uint64_t)ex->rank * ex->lay.slot_bytes;
  for (uint32_t k = 0; k < ex->n_ranks && e == cudaSuccess; ++k) {
    const uint32_t p = (ex->rank + k) % ex->n_ranks;  // different ranks start with different targets
    char* dst = ex->peer[ex->cur][p] + slot;
    e = cudaStreamWaitEvent(ex->side[p], ready, 0);
    if (e == cudaSuccess) {
      e = cudaMemcpyAsync(dst + ex->lay.leaf_off + 4 * first_entry, d_leaf + first_entry, 4 * n_entries,
                          cudaMemcpyDeviceToDevice, ex->side[p]);
    }
    if (e == cudaSuccess) {
      e = cudaMemcpyAsync(dst + ex->lay.info_off + 4 * first_entry, d_info + first_entry, 4 * n_entries,
                          cudaMemcpyDeviceToDevice, ex->side[p]);
    }
  }
  if (e != cudaSuccess) {
    ex->push_error = Q3DR_ERR_CUDA;
    ex->push_message = std::string("peer push: ") + cudaGetErrorString(e);
    cudaGetLastError();
  }
  ex->pushed_entries = first_entry + n_entries;
  return Q3DR_OK;
}

}  // namespace q3dr_detail

extern "C" {

int q3dr_exchange_layout_of(uint32_t n_ranks, uint64_t max_viewpoints, uint64_t max_entries, q3dr_exchange_layout* out) {
  if (!out) return fail(Q3DR_ERR_INVALID_ARG, "out is NULL");
  if (n_ranks == 0 || n_ranks > 1024) return fail(Q3DR_ERR_INVALID_ARG, "n_ranks out of range");
  q3dr_exchange_layout l;
  l.header_off = 0;
  l.offsets_off = up(sizeof(SlotHeader));
  l.total_off = up(l.offsets_off + 8 * (max_viewpoints + 1));
  l.leaf_off = up(l.total_off + 4 * max_viewpoints);
  l.info_off = up(l.leaf_off + 4 * max_entries);
  l.slot_bytes = up(l.info_off + 4 * max_entries);
  l.flags_off = l.slot_bytes * n_ranks;
  l.buffer_bytes = up(l.flags_off + 8 * (uint64_t)n_ranks);
  *out = l;
  return Q3DR_OK;
}

int q3dr_exchange_create(q3dr_map* map, uint32_t n_ranks, uint32_t rank, uint64_t max_viewpoints, uint64_t max_entries,
                         q3dr_exchange** out) {
  if (!out) return fail(Q3DR_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  Q3DR_TRY(q3dr_detail::check_map(map));
  if (rank >= n_ranks) return fail(Q3DR_ERR_INVALID_ARG, "rank >= n_ranks");
  if (map->exchange) return fail(Q3DR_ERR_INVALID_ARG, "the map already has an exchange");
  q3dr_exchange_layout lay;
  Q3DR_TRY(q3dr_exchange_layout_of(n_ranks, max_viewpoints, max_entries, &lay));
  Q3DR_CUDA(cudaSetDevice(map->device));
  q3dr_exchange* ex = new (std::nothrow) q3dr_exchange;
  if (!ex) return fail(Q3DR_ERR_OUT_OF_MEMORY, "host allocation failed");
  ex->map = map;
  ex->device = map->device;
  ex->n_ranks = n_ranks;
  ex->rank = rank;
  ex->max_vp = max_viewpoints;
  ex->max_entries = max_entries;
  ex->lay = lay;
  cudaError_t e = cudaSuccess;
  for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
    e = cudaMalloc(&ex->base[b], lay.buffer_bytes);
    if (e == cudaSuccess) e = cudaMemset(ex->base[b], 0, lay.buffer_bytes);
    ex->peer[b].assign(n_ranks, nullptr);
    ex->opened[b].assign(n_ranks, false);
    if (e == cudaSuccess) ex->peer[b][rank] = ex->base[b];
  }
  ex->side.assign(n_ranks, nullptr);
  for (uint32_t p = 0; p < n_ranks && e == cudaSuccess; ++p) e = cudaStreamCreateWithFlags(&ex->side[p], cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ex->ev_tail, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaHostAlloc(&ex->h_header, 2 * sizeof(SlotHeader), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc(&ex->h_flag, 2 * sizeof(uint64_t), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc(&ex->h_all, n_ranks * sizeof(SlotHeader), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc(&ex->h_status, sizeof(uint32_t), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaMalloc(&ex->d_status, sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&ex->d_stage, 2 * kAlign);
  if (e != cudaSuccess) {
    const std::string msg = std::string("exchange setup: ") + cudaGetErrorString(e);
    cudaGetLastError();
    q3dr_exchange_destroy(ex);
    return fail(e == cudaErrorMemoryAllocation ? Q3DR_ERR_OUT_OF_MEMORY : Q3DR_ERR_CUDA, msg);
  }
  ex->connected = (n_ranks == 1);
  map->exchange = ex;
  *out = ex;
  return Q3DR_OK;
}

int q3dr_exchange_disconnect(q3dr_exchange* ex) {
  if (!ex) return Q3DR_OK;
  cudaSetDevice(ex->device);
  for (cudaStream_t st : ex->side) {
    if (st) cudaStreamSynchronize(st);
  }
  for (int b = 0; b < 2; ++b) {
    for (uint32_t p = 0; p < ex->peer[b].size(); ++p) {
      if (ex->opened[b][p] && ex->peer[b][p]) cudaIpcCloseMemHandle(ex->peer[b][p]);
      ex->opened[b][p] = false;
      if (p != ex->rank) ex->peer[b][p] = nullptr;
    }
  }
  cudaGetLastError();
  ex->connected = (ex->n_ranks == 1);
  ex->armed = false;
  return Q3DR_OK;
}

int q3dr_exchange_destroy(q3dr_exchange* ex) {
  if (!ex) return Q3DR_OK;
  q3dr_exchange_disconnect(ex);
  if (ex->map && ex->map->exchange == ex) ex->map->exchange = nullptr;
  for (int b = 0; b < 2; ++b) {
    if (ex->base[b]) cudaFree(ex->base[b]);
  }
  for (cudaStream_t st : ex->side) {
    if (st) cudaStreamDestroy(st);
  }
  if (ex->ev_tail) cudaEventDestroy(ex->ev_tail);
  if (ex->h_header) cudaFreeHost(ex->h_header);
  if (ex->h_flag) cudaFreeHost(ex->h_flag);
  if (ex->h_all) cudaFreeHost(ex->h_all);
  if (ex->h_status) cudaFreeHost(ex->h_status);
  if (ex->d_status) cudaFree(ex->d_status);
  if (ex->d_stage) cudaFree(ex->d_stage);
  cudaGetLastError();
  delete ex;
  return Q3DR_OK;
}

int q3dr_exchange_export(q3dr_exchange* ex, void* handles_out) {
  if (!ex || !handles_out) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == Q3DR_IPC_HANDLE_BYTES, "CUDA IPC handle size");
  Q3DR_CUDA(cudaSetDevice(ex->device));
  for (int b = 0; b < 2; ++b) {
    cudaIpcMemHandle_t h;
    Q3DR_CUDA(cudaIpcGetMemHandle(&h, ex->base[b]));
    std::memcpy(static_cast<char*>(handles_out) + b * Q3DR_IPC_HANDLE_BYTES, &h, sizeof(h));
  }
  return Q3DR_OK;
}

int q3dr_exchange_import(q3dr_exchange* ex, const void* all_handles) {
  if (!ex || !all_handles) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  Q3DR_CUDA(cudaSetDevice(ex->device));
  for (uint32_t p = 0; p < ex->n_ranks; ++p) {
    if (p == ex->rank) continue;
    for (int b = 0; b < 2; ++b) {
      if (ex->peer[b][p]) continue;
      cudaIpcMemHandle_t h;
      std::memcpy(&h, static_cast<const char*>(all_handles) + ((size_t)p * 2 + b) * Q3DR_IPC_HANDLE_BYTES, sizeof(h));
      void* q = nullptr;
      Q3DR_CUDA(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
      ex->peer[b][p] = static_cast<char*>(q);
      ex->opened[b][p] = true;
    }
  }
  ex->connected = true;
  return Q3DR_OK;
}

int q3dr_exchange_connect(q3dr_exchange* const* all, uint32_t n_ranks) {
  if (!all || n_ranks == 0) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  for (uint32_t r = 0; r < n_ranks; ++r) {
    if (!all[r] || all[r]->n_ranks != n_ranks || all[r]->rank != r) {
      return fail(Q3DR_ERR_INVALID_ARG, "exchange r must have rank r of n_ranks");
    }
  }
  for (uint32_t r = 0; r < n_ranks; ++r) {
    q3dr_exchange* ex = all[r];
    Q3DR_CUDA(cudaSetDevice(ex->device));
    for (uint32_t p = 0; p < n_ranks; ++p) {
      if (p == r) continue;
      if (all[p]->device != ex->device) {
        int can = 0;
        Q3DR_CUDA(cudaDeviceCanAccessPeer(&can, ex->device, all[p]->device));
        if (!can) return fail(Q3DR_ERR_CUDA, "no peer access between two devices of the exchange");
        const cudaError_t e = cudaDeviceEnablePeerAccess(all[p]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
          return fail(Q3DR_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        }
        cudaGetLastError();
      }
      for (int b = 0; b < 2; ++b) ex->peer[b][p] = all[p]->base[b];
    }
    ex->connected = true;
  }
  return Q3DR_OK;
}

int q3dr_exchange_begin_step(q3dr_exchange* ex) {
  if (!ex) return fail(Q3DR_ERR_INVALID_ARG, "exchange is NULL");
  if (!ex->connected) return fail(Q3DR_ERR_INVALID_ARG, "exchange not connected (q3dr_exchange_import / _connect)");
  ex->cur ^= 1u;
  ex->step += 1;
  ex->armed = true;
  ex->valid = false;
  ex->push_error = Q3DR_OK;
  ex->push_message.clear();
  ex->pushed_entries = 0;
  return Q3DR_OK;
}

int q3dr_exchange_end_step(q3dr_exchange* ex, int local_status, void* stream_v) {
  if (!ex) return fail(Q3DR_ERR_INVALID_ARG, "exchange is NULL");
  if (!ex->armed) return fail(Q3DR_ERR_INVALID_ARG, "q3dr_exchange_end_step without q3dr_exchange_begin_step");
  ex->armed = false;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  Q3DR_CUDA(cudaSetDevice(ex->device));
  q3dr_map* m = ex->map;
  const uint64_t nv = m->last_n_vp, ne = m->last_n_entries;
  int failed = (local_status != Q3DR_OK) ? local_status : ex->push_error;
  std::string why = ex->push_message;
  if (failed == Q3DR_OK && nv > ex->max_vp) {
    failed = Q3DR_ERR_INVALID_ARG;
    why = "more viewpoints than the receive slot holds (max_viewpoints of q3dr_exchange_create)";
  }
  if (failed == Q3DR_OK && ne != ex->pushed_entries) {
    failed = Q3DR_ERR_INTERNAL;
    why = "the scoring call of this step did not push all its entries";
  }
  // headers + flags behind the data pushes
  const uint32_t b = ex->cur;
  SlotHeader* hh = ex->h_header + b;
  hh->n_viewpoints = failed ? 0 : nv;
  hh->n_entries = failed ? 0 : ne;
  hh->step = ex->step;
  hh->status = (uint64_t)failed;
  ex->h_flag[b] = (ex->step << 1) | (failed ? 1ull : 0ull);
  const uint64_t slot = (uint64_t)ex->rank * ex->lay.slot_bytes;
  char* stage = ex->d_stage + b * kAlign;  // {SlotHeader, flag} of this step in local device memory
  cudaStream_t own = ex->side[ex->rank];
  cudaError_t e = cudaMemcpyAsync(stage, hh, sizeof(SlotHeader), cudaMemcpyHostToDevice, own);
  if (e == cudaSuccess) e = cudaMemcpyAsync(stage + sizeof(SlotHeader), ex->h_flag + b, 8, cudaMemcpyHostToDevice, own);
  if (e == cudaSuccess) e = cudaEventRecord(ex->ev_tail, own);
  // per target rank, on that target's stream (behind the data pushes to it): offsets, totals, header, and the flag last --
  // a peer that sees the flag sees everything this rank pushed to it before
  for (uint32_t k = 0; k < ex->n_ranks && e == cudaSuccess; ++k) {
    const uint32_t p = (ex->rank + k) % ex->n_ranks;
    char* dst = ex->peer[b][p];
    cudaStream_t st = ex->side[p];
    if (p != ex->rank) e = cudaStreamWaitEvent(st, ex->ev_tail, 0);
    if (e == cudaSuccess && !failed && nv > 0) {
      e = cudaMemcpyAsync(dst + slot + ex->lay.offsets_off, m->d_offsets.p, 8 * (nv + 1), cudaMemcpyDeviceToDevice, st);
      if (e == cudaSuccess) {
        e = cudaMemcpyAsync(dst + slot + ex->lay.total_off, m->d_total.p, 4 * nv, cudaMemcpyDeviceToDevice, st);
      }
    }
    if (e == cudaSuccess) {
      e = cudaMemcpyAsync(dst + slot + ex->lay.header_off, stage, sizeof(SlotHeader), cudaMemcpyDeviceToDevice, st);
    }
    if (e == cudaSuccess) {
      e = cudaMemcpyAsync(dst + ex->lay.flags_off + 8 * (uint64_t)ex->rank, stage + sizeof(SlotHeader), 8, cudaMemcpyDeviceToDevice, st);
    }
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(Q3DR_ERR_CUDA, std::string("exchange header push: ") + cudaGetErrorString(e));
  }
  // wait for everybody's flag on the device, then fetch the headers
  uint64_t timeout_ms = 20000;
  if (const char* env = std::getenv("Q3DR_EXCHANGE_TIMEOUT_MS")) timeout_ms = (uint64_t)std::max<long>(1, std::strtol(env, nullptr, 10));
  exchange_wait_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<const volatile uint64_t*>(ex->base[b] + ex->lay.flags_off), ex->n_ranks,
                                            ex->step, timeout_ms * 1000000ull, ex->d_status);
  Q3DR_CUDA(cudaGetLastError());
  Q3DR_CUDA(cudaMemcpyAsync(ex->h_status, ex->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  Q3DR_CUDA(cudaMemcpy2DAsync(ex->h_all, sizeof(SlotHeader), ex->base[b] + ex->lay.header_off, ex->lay.slot_bytes, sizeof(SlotHeader),
                              ex->n_ranks, cudaMemcpyDeviceToHost, stream));
  Q3DR_CUDA(cudaStreamSynchronize(stream));
  for (cudaStream_t st : ex->side) Q3DR_CUDA(cudaStreamSynchronize(st));  // own pushes done: the map's pools may be reused by the next call
  if (failed != Q3DR_OK) return fail(failed, "exchange: this rank failed: " + why);
  if (*ex->h_status == 2u) return fail(Q3DR_ERR_INTERNAL, "exchange: timed out waiting for the peers' completion flags");
  if (*ex->h_status == 1u) return fail(Q3DR_ERR_INTERNAL, "exchange: another rank reported a failure in this step");
  for (uint32_t r = 0; r < ex->n_ranks; ++r) {
    if (ex->h_all[r].step != ex->step) return fail(Q3DR_ERR_INTERNAL, "exchange: header of a peer is from another step");
  }
  ex->valid = true;
  return Q3DR_OK;
}

int q3dr_exchange_view(const q3dr_exchange* ex, uint32_t src, q3dr_result* out) {
  if (!ex || !out) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  if (src >= ex->n_ranks) return fail(Q3DR_ERR_INVALID_ARG, "source rank out of range");
  if (!ex->valid) return fail(Q3DR_ERR_INVALID_ARG, "no completed step to view");
  const char* slot = ex->base[ex->cur] + (uint64_t)src * ex->lay.slot_bytes;
  out->n_viewpoints = ex->h_all[src].n_viewpoints;
  out->n_entries = ex->h_all[src].n_entries;
  out->offsets = reinterpret_cast<const uint64_t*>(slot + ex->lay.offsets_off);
  out->total = reinterpret_cast<const float*>(slot + ex->lay.total_off);
  out->leaf = reinterpret_cast<const int32_t*>(slot + ex->lay.leaf_off);
  out->information = reinterpret_cast<const float*>(slot + ex->lay.info_off);
  out->first_pixel = nullptr;
  out->dist_sq = nullptr;
  out->hit_count = nullptr;
  return Q3DR_OK;
}

}  // extern "C"
