// C-ABI implementation (include/pusch_dec_cuda.h): context, streams ("queues"), pinned staging, HARQ arena and kernel
// launches. Single translation unit: the kernels are included so that they share the constant-memory tables.
#include <algorithm>
#include "pdc_device.cuh"
#include "tables.cuh"
#include "rate_dematch.cuh"
#include "ldpc_decode_scalar.cuh"
#include "ldpc_decode_h2.cuh"
#include "tb_assemble.cuh"
#include "ulsch_demux.cuh"
#include "demod.cuh"
#include "ldpc_encode.cuh"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include <list>
#include <memory>
#include <mutex>

// ---- device allocations ---------------------------------------------------------------------------------------------
#ifdef PDC_DEBUG_BOUNDS
// Debug build: 256 canary bytes behind every device allocation; pdc_debug_canaries_ok() reads them back.
constexpr size_t  CANARY_BYTES = 256;
constexpr uint8_t CANARY_VALUE = 0xA5;
struct CanaryEntry {
  const uint8_t* base;
  size_t         bytes;
  int            device;
};
struct CanaryRegistry {
  std::mutex               m;
  std::vector<CanaryEntry> entries;
};
static CanaryRegistry& canaries()
{
  static CanaryRegistry r;
  return r;
}
static cudaError_t canary_malloc(void** p, size_t bytes)
{
  cudaError_t e = cudaMalloc(p, bytes + CANARY_BYTES);
  if (e == cudaSuccess) {
    e = cudaMemset(static_cast<uint8_t*>(*p) + bytes, CANARY_VALUE, CANARY_BYTES);
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(canaries().m);
    canaries().entries.push_back(CanaryEntry{static_cast<const uint8_t*>(*p), bytes, dev});
  }
  return e;
}
static cudaError_t canary_free(const void* p)
{
  if (p != nullptr) {
    std::lock_guard<std::mutex> lock(canaries().m);
    auto&                       v = canaries().entries;
    for (size_t i = 0; i != v.size(); ++i) {
      if (v[i].base == p) {
        v[i] = v.back();
        v.pop_back();
        break;
      }
    }
  }
  return cudaFree(const_cast<void*>(p));
}
// Number of allocations whose canary was overwritten (0 = intact), -1 on a CUDA error.
static int canary_check(int device)
{
  std::lock_guard<std::mutex> lock(canaries().m);
  int                         bad = 0;
  std::vector<uint8_t>        h(CANARY_BYTES);
  for (const CanaryEntry& c : canaries().entries) {
    if (c.device != device) {
      continue;
    }
    if (cudaMemcpy(h.data(), c.base + c.bytes, CANARY_BYTES, cudaMemcpyDeviceToHost) != cudaSuccess) {
      return -1;
    }
    for (uint8_t b : h) {
      if (b != CANARY_VALUE) {
        fprintf(stderr, "[pusch_dec_cuda] canary behind a %zu-byte device allocation was overwritten\n", c.bytes);
        ++bad;
        break;
      }
    }
  }
  return bad;
}
#define PDC_MALLOC(p, bytes) canary_malloc(reinterpret_cast<void**>(p), (bytes))
#define PDC_FREE(p) canary_free(p)
#else
#define PDC_MALLOC(p, bytes) cudaMalloc(reinterpret_cast<void**>(p), (bytes))
#define PDC_FREE(p) cudaFree(p)
#endif

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* what, cudaError_t e = cudaSuccess)
{
  g_last_error = what;
  if (e != cudaSuccess) {
    g_last_error += ": ";
    g_last_error += cudaGetErrorString(e);
  }
  return code;
}

#define PDC_CUDA(expr)                                                                                                 \
  do {                                                                                                                 \
    cudaError_t e__ = (expr);                                                                                          \
    if (e__ != cudaSuccess) {                                                                                          \
      return fail(PDC_ERR_CUDA, #expr, e__);                                                                           \
    }                                                                                                                  \
  } while (0)

// Soft demapper stage (grow-only buffers). The call table travels in the kernel parameters: nothing to upload.
struct DemodStage {
  float2*              d_sym = nullptr;
  size_t               sym_cap = 0;
  float*               d_nv = nullptr;
  size_t               nv_cap = 0;
  std::vector<pdc::DemodCall> calls;
  const float2*        k_sym = nullptr;
  const float*         k_nv = nullptr;
  int8_t*              k_llrs = nullptr;
  bool                 kernel_pending = false;
};

// Buffers of the codeword front end (grow-only; the plan is staged in pinned memory).
struct FrontEnd {
  DemodStage           dm;
  pdc::UlschPlan       plan;
  int8_t*              d_raw = nullptr;
  size_t               raw_cap = 0;
  uint32_t*            d_seq = nullptr;
  size_t               seq_cap = 0;
  int8_t*              d_uci = nullptr;
  int8_t*              h_uci = nullptr;
  size_t               uci_cap = 0;
  unsigned char*       d_plan = nullptr;
  size_t               plan_cap = 0;
  // The plan is staged in a ring of pinned buffers: a slot is reused once its upload (an event) has completed, so
  // back-to-back launches on one stream do not overwrite a plan that is still waiting to be copied.
  static constexpr int RING = 4;
  unsigned char*       h_plan[RING]   = {nullptr, nullptr, nullptr, nullptr};
  size_t               h_plan_cap[RING] = {0, 0, 0, 0};
  cudaEvent_t          plan_ev[RING]  = {nullptr, nullptr, nullptr, nullptr};
  int                  ring_pos = 0;
  // Deferred descrambling: the last front end left codewords without UCI scrambled in d_in for the rate dematcher of
  // the next batch that reads its LLRs from d_sch.
  bool                 deferred = false;
  uint32_t             deferred_n_cw = 0;
  const int8_t*        deferred_in = nullptr;
  const int8_t*        deferred_sch = nullptr;
  uint4*               d_cb_scr = nullptr;
  size_t               cb_scr_cap = 0;
  // Kernels of the last plan (launched right away, or - for the queued interface - after the descriptor uploads of the
  // decode batch so that all host-to-device copies of a slot are issued back to back).
  pdc::UlschArgs       args = {};
  uint32_t             k_n_cw = 0, k_max_in = 0, k_max_sch = 0, k_max_uci = 0;
  bool                 k_all_deferred = false, k_generate_seq = false, kernels_pending = false;
  int8_t*              uci_copy_dst = nullptr; // host destination of the UCI soft bits once the kernels are queued
  // Pending results of the queue.
  int8_t*              u_uci = nullptr;
  size_t               uci_bytes = 0;
  uint32_t             sch_len = 0; // soft bits of the demultiplexed UL-SCH space waiting for pdc_submit
  bool                 pending = false;
};

// A large batch of pdc_submit is pipelined (see submit_pipelined): its soft bits go to the device in groups on a copy
// stream, and the kernels of a group start on one of a few "lane" streams as soon as its group has arrived.
constexpr int    PIPE_LANES      = 4;
constexpr int    PIPE_MAX_GROUPS = 16;
constexpr int    PIPE_GROUPS     = 8;               // groups a batch is cut into (when large enough)
constexpr size_t PIPE_MIN_BYTES  = 5u << 19;        // 2.5 MB: batches below this go in one piece (a 1.36 MB slot gains
                                                    // nothing: the extra launches and events cost what the overlap saves)
constexpr size_t PIPE_MIN_GROUP  = 1u << 20;

struct Queue {
  FrontEnd     fe;
  cudaStream_t stream = nullptr;
  cudaEvent_t  done   = nullptr;
  // Pipelined submission (batch queues only).
  cudaStream_t copy_stream = nullptr;
  cudaStream_t lane[PIPE_LANES] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t  ev_copy[PIPE_MAX_GROUPS] = {};
  cudaEvent_t  ev_lane[PIPE_LANES] = {};
  // PDC_PIPE_TRACE=1 (measurement aid): device timestamps of the pipelined submission - start, end of every group's copy,
  // end of every group's kernels and copies out, end of the batch - printed by pdc_wait.
  cudaEvent_t  tr_start = nullptr, tr_done = nullptr, tr_copy[PIPE_MAX_GROUPS] = {}, tr_group[PIPE_MAX_GROUPS] = {},
               tr_kern[PIPE_MAX_GROUPS] = {};
  int          tr_groups = 0;
  double       tr_host_us = 0;
  // Descriptors {transport blocks, codeblocks} and results {transport blocks, codeblocks} are one block each, on the
  // device and in page-locked host memory: one copy in and one copy out per batch instead of two (a small copy costs a
  // few microseconds of latency whatever its size). d_tbs / d_cbs / d_tb_res / d_cb_res point into the blocks.
  uint8_t* d_desc = nullptr;
  uint8_t* h_desc = nullptr;
  uint8_t* d_res  = nullptr;
  uint8_t* h_res  = nullptr;
  size_t   tb_desc_area = 0, tb_res_area = 0; // bytes in front of the codeblock part
  // Device staging.
  pdc_cb_desc*   d_cbs     = nullptr;
  int8_t*        d_llrs    = nullptr;
  pdc_tb_desc*   d_tbs     = nullptr;
  pdc_cb_result* d_cb_res  = nullptr;
  uint8_t*       d_cb_bits = nullptr;
  pdc_tb_result* d_tb_res  = nullptr;
  uint8_t*       d_tb_out  = nullptr;
  uint32_t*      d_tb_sync = nullptr; // {CRC accumulator, arrival counter} per TB for the TB assembly kernel
  // Pinned host mirrors.
  pdc_cb_desc*   h_cbs     = nullptr;
  pdc_tb_desc*   h_tbs     = nullptr;
  pdc_cb_result* h_cb_res  = nullptr;
  uint8_t*       h_cb_bits = nullptr;
  pdc_tb_result* h_tb_res  = nullptr;
  uint8_t*       h_tb_out  = nullptr;
  // Pending batch.
  bool           busy         = false;
  uint32_t       n_cb         = 0;
  uint32_t       n_tb         = 0;
  size_t         tb_out_bytes = 0;
  pdc_cb_result* u_cb_res     = nullptr;
  uint8_t*       u_cb_bits    = nullptr;
  pdc_tb_result* u_tb_res     = nullptr;
  uint8_t*       u_tb_out     = nullptr;
};

} // namespace

struct pdc_ctx {
  pdc_config           cfg;
  int                  sm_count = 0, cc_major = 0, cc_minor = 0;
  int8_t*              d_harq      = nullptr; // (harq_entries + 1) x PDC_MAX_CB_SOFT; the extra entry is scratch
  uint8_t*             d_harq_data = nullptr; // (harq_entries + 1) x PDC_MAX_CB_BYTES
  uint32_t*            d_tb_sync = nullptr;   // per TB {CRC accumulator, arrival counter} of the TB assembly kernel
  uint32_t             tb_sync_entries = 0;
  int32_t*             d_harq_last = nullptr; // (harq_entries + 1) x DM_MAX_PARTS: pdc::harq_last_pack records, -1 unknown
  int8_t*              d_scratch_llr = nullptr;
  size_t               scratch_llr_bytes = 0;
  float*               d_demod_tables = nullptr;  // piecewise-linear LLR tables of the soft demapper
  std::vector<Queue>   queues;
  std::atomic<uint64_t> launches{0};
  bool                 force_scalar = false;      // PDC_FORCE_SCALAR=1: use the general kernel for every batch
  bool                 no_pipeline  = false;      // PDC_NO_PIPELINE=1: every pdc_submit batch in one piece
  // Compressed check-to-variable messages of the resident CTAs + the work counter of the persistent decoder, ONE SET
  // PER CUDA STREAM: decoder launches of different streams may overlap at their tails and must not share scratch.
  struct DecodeScratch {
    cudaStream_t stream  = nullptr;
    uint32_t*    d_state = nullptr;
    size_t       words   = 0;
    uint32_t*    d_counter = nullptr;
    uint32_t     counter_base = 0; // value of *d_counter once the launches queued so far have run
  };
  // One entry per queue stream is created (and fully sized) by pdc_create, so nothing is allocated on the launch path of
  // the queued interface; streams of pdc_launch_device callers get theirs on first use. A list: entries never move, so
  // a launch may keep a pointer to its entry while another thread appends one; the lookup / append is under the mutex.
  std::list<DecodeScratch> decode_scratch;
  std::mutex               decode_scratch_mutex;
  FrontEnd             fe_sync;                   // buffers of the synchronous front-end calls
  // The synchronous single-object calls (pdc_ldpc_decode, pdc_rate_dematch, pdc_crc, pdc_demodulate_soft,
  // pdc_ulsch_demux, pdc_scrambling_sequence, pdc_encode, pdc_ldpc_encode) share fe_sync, d_scratch_llr, the scratch
  // HARQ entry and this private one-codeblock queue; they may be called from any number of threads (the reference runs
  // a pool of decoder / dematcher objects concurrently) and serialise on the mutex. They never touch the batch queues,
  // so they cannot collide with a pdc_submit in flight.
  Queue                sync_q;
  std::mutex           sync_mutex;
  // One mutex per batch queue: pdc_submit .. pdc_wait of a queue may come from different threads one after the other,
  // and two threads that (against the contract) drive one queue at once get PDC_ERR_CAPACITY instead of corrupting it.
  std::unique_ptr<std::mutex[]> queue_mutex;
  // Downlink twin (synchronous calls): grow-only device staging.
  unsigned char*       d_enc_cbs = nullptr;
  size_t               enc_cbs_cap = 0;
  unsigned char*       d_enc_msgs = nullptr;
  size_t               enc_msgs_cap = 0;
  unsigned char*       d_enc_out = nullptr;
  size_t               enc_out_cap = 0;
  int8_t*              d_sch_sync = nullptr;
  size_t               sch_sync_cap = 0;
};

namespace {

struct BatchShape {
  int  max_Z   = 0;
  bool any_bg1 = false;
  bool any_decode = false, any_dematch = false;
  // every codeblock is a first transmission (rv 0) whose soft bits end inside the first 24 variable nodes: four base-graph
  // rows in use (ldpc_decoder_impl.cpp:96-114) - selects the decoder instantiation that is fastest on those
  bool high_rate = false;
};

BatchShape scan_batch(const pdc_cb_desc* cbs, uint32_t n)
{
  BatchShape s;
  bool       high = true;
  for (uint32_t i = 0; i != n; ++i) {
    if (cbs[i].lifting_size <= pdc::MAX_Z) {
      s.max_Z = std::max<int>(s.max_Z, cbs[i].lifting_size);
    }
    s.any_bg1 |= (cbs[i].base_graph == 1);
    s.any_decode |= (cbs[i].flags & PDC_CB_DECODE) != 0;
    s.any_dematch |= (cbs[i].flags & PDC_CB_DEMATCH) != 0;
    const uint8_t first = PDC_CB_DECODE | PDC_CB_DEMATCH | PDC_CB_NEW_DATA;
    high = high && (cbs[i].flags & first) == first && cbs[i].rv == 0 &&
           (uint64_t)cbs[i].rm_length + cbs[i].nof_filler <=
               ((cbs[i].base_graph == 1) ? 24u : 12u) * (uint64_t)cbs[i].lifting_size;
  }
  s.high_rate = high && n != 0;
  if (s.max_Z < 2) {
    s.max_Z = 2;
  }
  return s;
}

template <typename T>
static cudaError_t grow_device(T** p, size_t* cap, size_t need)
{
  if (need <= *cap) {
    return cudaSuccess;
  }
  PDC_FREE(*p);
  *p   = nullptr;
  *cap = 0;
  size_t      want = need + need / 4 + 64;
  cudaError_t e    = PDC_MALLOC(p, want * sizeof(T));
  if (e == cudaSuccess) {
    *cap = want;
  }
  return e;
}

// Row-state scratch + ticket counter of the persistent decoder for stream s, holding at least `need` words. Queue
// streams find their entry pre-sized (pdc_create); a foreign stream (pdc_launch_device) allocates on first use.
cudaError_t decode_scratch_for(pdc_ctx* ctx, cudaStream_t s, size_t need, pdc_ctx::DecodeScratch** out)
{
  pdc_ctx::DecodeScratch* sc = nullptr;
  {
    std::lock_guard<std::mutex> lock(ctx->decode_scratch_mutex);
    for (pdc_ctx::DecodeScratch& c : ctx->decode_scratch) {
      if (c.stream == s) {
        sc = &c;
      }
    }
    if (sc == nullptr) {
      ctx->decode_scratch.emplace_back();
      sc         = &ctx->decode_scratch.back();
      sc->stream = s;
    }
  }
  // From here on the entry belongs to the caller: a stream is driven by one thread at a time.
  if (sc->d_counter == nullptr) {
    cudaError_t e = PDC_MALLOC(&sc->d_counter, sizeof(uint32_t));
    if (e == cudaSuccess) {
      e = cudaMemsetAsync(sc->d_counter, 0, sizeof(uint32_t), s);
    }
    if (e != cudaSuccess) {
      return e;
    }
    sc->counter_base = 0;
  }
  if (need > sc->words) {
    PDC_FREE(sc->d_state); // synchronises the device: only on the first batch of a foreign stream
    sc->d_state   = nullptr;
    sc->words     = 0;
    cudaError_t e = PDC_MALLOC(&sc->d_state, need * sizeof(uint32_t));
    if (e != cudaSuccess) {
      return e;
    }
    sc->words = need;
  }
  *out = sc;
  return cudaSuccess;
}

// Largest row-state scratch any batch can ask for on this device: the launch plan of every lifting size with all
// resident CTAs in use.
size_t decode_scratch_max_words(int sm_count)
{
  size_t words = 0;
  for (int i = 0; i != NR_LDPC_NOF_LIFTING_SIZES; ++i) {
    pdc::H2Plan plan;
    if (pdc::h2_plan(pdc::NR_LDPC_LIFTING_SIZES[i], true, 0x7fffffffu, sm_count, plan) == cudaSuccess) {
      words = std::max(words, plan.scratch_words_per_cta * (size_t)plan.grid);
    }
  }
  return words;
}

// Launches the kernels of one batch on stream s. All pointers are device pointers.
int launch_batch(pdc_ctx*             ctx,
                 const pdc_cb_desc*   d_cbs,
                 uint32_t             n_cb,
                 const int8_t*        d_llrs,
                 const pdc_tb_desc*   d_tbs,
                 uint32_t             n_tb,
                 pdc_cb_result*       d_cb_res,
                 uint8_t*             d_cb_bits,
                 pdc_tb_result*       d_tb_res,
                 uint8_t*             d_tb_out,
                 const BatchShape&    shape,
                 const int8_t*        direct_in,
                 uint32_t             direct_n,
                 cudaStream_t         s,
                 uint32_t*            tb_sync = nullptr,
                 FrontEnd*            fe      = nullptr,
                 const pdc_cb_desc*   tb_cbs_base = nullptr,   // what the TB descriptors' first_cb indexes (default: d_cbs)
                 const pdc_cb_result* tb_res_base = nullptr,
                 const uint8_t*       tb_bits_base = nullptr)
{
  pdc::BatchParams p;
  p.cb_scr = nullptr;
  p.seq    = nullptr;
  p.raw    = nullptr;
  if (fe != nullptr && fe->deferred && shape.any_dematch && d_llrs == fe->deferred_sch) {
    // Codewords whose descrambling was deferred to this batch: map every codeblock to its codeword (one-shot).
    PDC_CUDA(grow_device(&fe->d_cb_scr, &fe->cb_scr_cap, (size_t)n_cb));
    PDC_CUDA(pdc::launch_pdl(pdc::cb_descramble_map_kernel, dim3((n_cb + 255) / 256), dim3(256), 0, s, d_cbs, n_cb,
                             reinterpret_cast<const pdc::UlschCodeword*>(fe->d_plan), fe->deferred_n_cw, fe->d_cb_scr));
    ctx->launches++;
    p.cb_scr     = fe->d_cb_scr;
    p.seq        = fe->d_seq;
    p.raw        = fe->deferred_in;
    fe->deferred = false;
  }
  p.cbs          = d_cbs;
  p.n_cb         = n_cb;
  p.llrs         = d_llrs;
  p.harq         = ctx->d_harq;
  p.harq_entries = ctx->cfg.harq_entries + 1;
  p.results      = d_cb_res;
  p.cb_bits      = d_cb_bits;
  p.harq_data    = ctx->d_harq_data;
  p.harq_last    = ctx->d_harq_last;
  p.scale_mode   = ctx->cfg.scale_mode;
  p.simd_width   = ctx->cfg.combine_simd_width;
  if (shape.any_dematch) {
    PDC_CUDA(pdc::launch_rate_dematch(p, ctx->sm_count, s));
    ctx->launches++;
  }
  if (shape.any_decode) {
    if (direct_in == nullptr && ctx->cfg.scale_mode != PDC_SCALE_NEON && !ctx->force_scalar) {
      // Throughput kernel: two codeblocks per CTA, half-precision packed arithmetic.
      pdc::H2Plan plan;
      PDC_CUDA(pdc::h2_plan(shape.max_Z, shape.any_bg1, n_cb, ctx->sm_count, plan, ctx->cfg.scale_mode, shape.high_rate));
      size_t need = plan.scratch_words_per_cta * (size_t)plan.grid;
      pdc_ctx::DecodeScratch* sc = nullptr;
      PDC_CUDA(decode_scratch_for(ctx, s, need, &sc));
      // Pairs are handed out dynamically (codeblocks that stop early free their CTA for the next pair).
      {
        cudaError_t e = pdc::launch_ldpc_decode_h2(p, plan, sc->d_state, sc->d_counter, sc->counter_base, s);
        if (e != cudaSuccess) {
          // Re-arm the ticket counter for whatever comes next.
          cudaMemsetAsync(sc->d_counter, 0, sizeof(uint32_t), s);
          sc->counter_base = 0;
          return fail(PDC_ERR_CUDA, "launch_ldpc_decode_h2", e);
        }
        sc->counter_base += (n_cb + 1) / 2;
      }
    } else {
      PDC_CUDA(pdc::launch_ldpc_decode(p, shape.max_Z, shape.any_bg1, direct_in, direct_n, s));
    }
    ctx->launches++;
  }
  if (n_tb != 0) {
    pdc::TbParams t;
    t.tbs        = d_tbs;
    t.n_tb       = n_tb;
    t.cbs        = tb_cbs_base ? tb_cbs_base : d_cbs;
    t.cb_results = tb_res_base ? tb_res_base : d_cb_res;
    t.cb_bits    = tb_bits_base ? tb_bits_base : d_cb_bits;
    t.tb_results = d_tb_res;
    t.tb_bytes   = d_tb_out;
    if (n_tb > ctx->tb_sync_entries) {
      return fail(PDC_ERR_CAPACITY, "launch_batch: more transport blocks than the context was created for");
    }
    PDC_CUDA(pdc::launch_tb_assemble(t, ctx->d_harq_data, tb_sync ? tb_sync : ctx->d_tb_sync, s));
    ctx->launches++;
  }
  return PDC_OK;
}

template <typename T>
cudaError_t dev_alloc(T** p, size_t n)
{
  return PDC_MALLOC(p, std::max<size_t>(n, 1) * sizeof(T));
}
template <typename T>
cudaError_t host_alloc(T** p, size_t n)
{
  return cudaMallocHost(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T));
}

// Generic CRC of a packed bit string (compatibility API, not a hot path).
__constant__ uint32_t c_xpow2[pdc::CRC_KINDS][24]; // x^(2^i) mod P

__global__ void crc_kernel(const uint8_t* packed, uint32_t nbits, int kind, uint32_t* out)
{
  __shared__ uint32_t sh;
  if (threadIdx.x == 0) {
    sh = 0;
  }
  __syncthreads();
  const uint32_t poly   = pdc::crc_poly_any(kind);
  const int      order  = pdc::crc_order_any(kind);
  const uint32_t n_full = nbits / 32u;
  const uint32_t rem    = nbits & 31u;
  const uint32_t T      = n_full + (rem ? 1u : 0u);
  uint32_t       acc    = 0;
  for (uint32_t t = threadIdx.x; t < T; t += blockDim.x) {
    uint32_t w = 0;
    for (int k = 0; k != 4; ++k) {
      uint32_t byte = 4u * t + k;
      uint32_t v    = (8u * byte < nbits) ? packed[byte] : 0u;
      w             = (w << 8) | v;
    }
    uint32_t e;
    if (t < n_full) {
      e = nbits - 32u * (t + 1u) + order;
    } else {
      w >>= (32u - rem);
      e = order;
    }
    uint32_t xp = 1;
    for (int i = 0; e != 0; ++i, e >>= 1) {
      if (e & 1u) {
        xp = pdc::gf2_mulmod(xp, c_xpow2[kind - 1][i], poly, order);
      }
    }
    acc ^= pdc::gf2_mulmod(w, xp, poly, order);
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
  }
  if ((threadIdx.x & 31) == 0 && acc) {
    atomicXor(&sh, acc);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *out = sh;
  }
}

// Integer-pipe throughput probe: the decoder is bound by 32-bit integer/logic instruction issue (no tensor cores, no
// HBM), so the roofline denominator is measured on the device it runs on. mode 0: LOP3 + IADD3 only (ALU pipe);
// mode 1: LOP3/IADD3 interleaved with IMAD (ALU + FMA pipes).
template <int MODE>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t* out, int iters)
{
  uint32_t a[8];
  uint32_t b = threadIdx.x * 2654435761u + 1u, c = blockIdx.x * 40503u + 7u;
#pragma unroll
  for (int k = 0; k != 8; ++k) {
    a[k] = b + k;
  }
  for (int i = 0; i != iters; ++i) {
#pragma unroll
    for (int r = 0; r != 8; ++r) {
#pragma unroll
      for (int k = 0; k != 8; ++k) {
        if (MODE == 0 || (k & 1)) {
          a[k] = (a[k] ^ b) + c; // LOP3 + IADD3
        } else {
          a[k] = a[k] * 5u + b;  // IMAD
          a[k] = a[k] * 3u + c;  // IMAD
        }
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k != 8; ++k) {
    r ^= a[k];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

cudaError_t upload_crc_tables()
{
  uint32_t h[pdc::CRC_KINDS][24];
  for (int k = 0; k != pdc::CRC_KINDS; ++k) {
    uint32_t poly  = pdc::crc_poly_any(k + 1);
    int      order = pdc::crc_order_any(k + 1);
    uint32_t x     = 2; // x^1
    for (int i = 0; i != 24; ++i) {
      h[k][i] = x;
      x       = pdc::gf2_mulmod(x, x, poly, order);
    }
  }
  return cudaMemcpyToSymbol(c_xpow2, h, sizeof(h));
}

// Streams, device staging and page-locked mirrors of one queue. full_scratch: pre-size the decoder's row-state scratch
// for the largest batch (batch queues); the one-codeblock queue of the synchronous calls sizes its own on first use.
cudaError_t create_queue(pdc_ctx* ctx, Queue& q, uint32_t max_cbs, size_t max_llrs, uint32_t max_tbs, size_t max_tb_bytes,
                         bool full_scratch)
{
#define PDC_Q(expr)                                                                                                    \
  do {                                                                                                                 \
    cudaError_t e__ = (expr);                                                                                          \
    if (e__ != cudaSuccess) {                                                                                          \
      return e__;                                                                                                      \
    }                                                                                                                  \
  } while (0)
  PDC_Q(cudaStreamCreateWithFlags(&q.stream, cudaStreamNonBlocking));
  {
    pdc_ctx::DecodeScratch* sc = nullptr;
    PDC_Q(decode_scratch_for(ctx, q.stream, full_scratch ? decode_scratch_max_words(ctx->sm_count) : 0, &sc));
  }
  PDC_Q(cudaEventCreateWithFlags(&q.done, cudaEventDisableTiming));
  if (full_scratch) {
    // Streams and events of the pipelined submission; every lane decodes with its own row-state scratch.
    PDC_Q(cudaStreamCreateWithFlags(&q.copy_stream, cudaStreamNonBlocking));
    for (int l = 0; l != PIPE_LANES; ++l) {
      PDC_Q(cudaStreamCreateWithFlags(&q.lane[l], cudaStreamNonBlocking));
      PDC_Q(cudaEventCreateWithFlags(&q.ev_lane[l], cudaEventDisableTiming));
      pdc_ctx::DecodeScratch* sc = nullptr;
      PDC_Q(decode_scratch_for(ctx, q.lane[l], decode_scratch_max_words(ctx->sm_count), &sc));
    }
    for (int g = 0; g != PIPE_MAX_GROUPS; ++g) {
      PDC_Q(cudaEventCreateWithFlags(&q.ev_copy[g], cudaEventDisableTiming));
    }
  }
  q.tb_desc_area = (sizeof(pdc_tb_desc) * max_tbs + 15) & ~(size_t)15;
  q.tb_res_area  = (sizeof(pdc_tb_result) * max_tbs + 15) & ~(size_t)15;
  PDC_Q(dev_alloc(&q.d_desc, q.tb_desc_area + sizeof(pdc_cb_desc) * max_cbs));
  PDC_Q(dev_alloc(&q.d_res, q.tb_res_area + sizeof(pdc_cb_result) * max_cbs));
  q.d_tbs    = reinterpret_cast<pdc_tb_desc*>(q.d_desc);
  q.d_cbs    = reinterpret_cast<pdc_cb_desc*>(q.d_desc + q.tb_desc_area);
  q.d_tb_res = reinterpret_cast<pdc_tb_result*>(q.d_res);
  q.d_cb_res = reinterpret_cast<pdc_cb_result*>(q.d_res + q.tb_res_area);
  PDC_Q(dev_alloc(&q.d_llrs, (size_t)max_llrs + 16));
  PDC_Q(dev_alloc(&q.d_cb_bits, (size_t)max_cbs * PDC_MAX_CB_BYTES));
  PDC_Q(dev_alloc(&q.d_tb_out, (size_t)max_tb_bytes + 16));
  PDC_Q(dev_alloc(&q.d_tb_sync, 2 * (size_t)ctx->tb_sync_entries));
  PDC_Q(cudaMemset(q.d_tb_sync, 0, 2 * (size_t)ctx->tb_sync_entries * sizeof(uint32_t)));
  PDC_Q(host_alloc(&q.h_desc, q.tb_desc_area + sizeof(pdc_cb_desc) * max_cbs));
  PDC_Q(host_alloc(&q.h_res, q.tb_res_area + sizeof(pdc_cb_result) * max_cbs));
  q.h_tbs    = reinterpret_cast<pdc_tb_desc*>(q.h_desc);
  q.h_cbs    = reinterpret_cast<pdc_cb_desc*>(q.h_desc + q.tb_desc_area);
  q.h_tb_res = reinterpret_cast<pdc_tb_result*>(q.h_res);
  q.h_cb_res = reinterpret_cast<pdc_cb_result*>(q.h_res + q.tb_res_area);
  PDC_Q(host_alloc(&q.h_cb_bits, (size_t)max_cbs * PDC_MAX_CB_BYTES));
  PDC_Q(host_alloc(&q.h_tb_out, (size_t)max_tb_bytes + 16));
#undef PDC_Q
  return cudaSuccess;
}

} // namespace

extern "C" {

void pdc_default_config(pdc_config* cfg)
{
  memset(cfg, 0, sizeof(*cfg));
  cfg->device             = 0;
  cfg->max_cbs            = 4096;
  cfg->max_llrs           = 4096u * PDC_MAX_CB_SOFT;
  cfg->harq_entries       = 4096;
  cfg->max_tbs            = 256;
  cfg->max_tb_bytes       = 4u << 20;
  cfg->scale_mode         = PDC_SCALE_X86;
  cfg->combine_simd_width = 64;
  cfg->nof_streams        = 2;
  cfg->demod_mode         = PDC_DEMOD_X86;
}

const char* pdc_last_error(void)
{
  return g_last_error.c_str();
}

int pdc_create(const pdc_config* cfg, pdc_ctx** out)
{
  if (!cfg || !out || cfg->max_cbs == 0 || cfg->harq_entries == 0 || cfg->nof_streams == 0 ||
      cfg->scale_mode < PDC_SCALE_X86 || cfg->scale_mode > PDC_SCALE_NEON ||
      (cfg->demod_mode != PDC_DEMOD_X86 && cfg->demod_mode != PDC_DEMOD_SCALAR)) {
    return fail(PDC_ERR_INVALID, "pdc_create: invalid configuration");
  }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0 || cfg->device >= n_dev) {
    return fail(PDC_ERR_NO_DEVICE, "pdc_create: no usable CUDA device (this library has no CPU fallback)");
  }
  PDC_CUDA(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  PDC_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) {
    return fail(PDC_ERR_NO_DEVICE, "pdc_create: device is not sm_100 (the kernels are built for sm_100a only)");
  }
  pdc_ctx* ctx = new (std::nothrow) pdc_ctx;
  if (!ctx) {
    return fail(PDC_ERR_INVALID, "pdc_create: out of host memory");
  }
  ctx->cfg      = *cfg;
  {
    const char* fs    = getenv("PDC_FORCE_SCALAR");
    ctx->force_scalar = fs && fs[0] == '1';
    const char* np    = getenv("PDC_NO_PIPELINE");
    ctx->no_pipeline  = np && np[0] == '1';
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cc_major = prop.major;
  ctx->cc_minor = prop.minor;
  *out          = ctx;
#define PDC_CREATE(expr)                                                                                               \
  do {                                                                                                                 \
    cudaError_t e__ = (expr);                                                                                          \
    if (e__ != cudaSuccess) {                                                                                          \
      pdc_destroy(ctx);                                                                                                \
      *out = nullptr;                                                                                                  \
      return fail(PDC_ERR_CUDA, #expr, e__);                                                                           \
    }                                                                                                                  \
  } while (0)
  PDC_CREATE(pdc::upload_tables());
  PDC_CREATE(pdc::upload_h2_images());
  PDC_CREATE(pdc::upload_tb_tables());
  PDC_CREATE(pdc::upload_prg_tables());
  PDC_CREATE(upload_crc_tables());
  size_t entries = (size_t)cfg->harq_entries + 1;
  PDC_CREATE(dev_alloc(&ctx->d_harq, entries * PDC_MAX_CB_SOFT));
  PDC_CREATE(cudaMemset(ctx->d_harq, 0, entries * PDC_MAX_CB_SOFT));
  PDC_CREATE(dev_alloc(&ctx->d_harq_data, entries * PDC_MAX_CB_BYTES));
  PDC_CREATE(cudaMemset(ctx->d_harq_data, 0, entries * PDC_MAX_CB_BYTES));
  ctx->tb_sync_entries = std::max<uint32_t>(cfg->max_tbs, 1);
  PDC_CREATE(dev_alloc(&ctx->d_tb_sync, 2 * (size_t)ctx->tb_sync_entries));
  PDC_CREATE(cudaMemset(ctx->d_tb_sync, 0, 2 * (size_t)ctx->tb_sync_entries * sizeof(uint32_t)));
  PDC_CREATE(dev_alloc(&ctx->d_harq_last, entries * pdc::DM_MAX_PARTS));
  {
    // all-zero entries: nothing non-zero anywhere in the slot
    std::vector<int32_t> fresh(entries * pdc::DM_MAX_PARTS, pdc::harq_last_pack(0, PDC_MAX_CB_SOFT));
    PDC_CREATE(cudaMemcpy(ctx->d_harq_last, fresh.data(), fresh.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  PDC_CREATE(dev_alloc(&ctx->d_demod_tables, (size_t)pdc::demod::TABLE_FLOATS));
  pdc::demod::demod_tables_kernel<<<1, 128>>>(ctx->d_demod_tables);
  PDC_CREATE(cudaGetLastError());
  ctx->scratch_llr_bytes = 35u * PDC_MAX_CB_BYTES * 8u; // MAX_CODEBLOCK_RM_SIZE (ldpc.h:122)
  PDC_CREATE(dev_alloc(&ctx->d_scratch_llr, ctx->scratch_llr_bytes));
  // Kernel attributes are per device.
  PDC_CREATE(pdc::h2_configure_device());
  PDC_CREATE(pdc::scalar_configure_device());
  PDC_CREATE(cudaFuncSetAttribute(pdc::ldpc_encode_rm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)pdc::enc_smem_bytes(1, pdc::MAX_Z)));
  // The decoder's row state carries an evict-last policy. PDC_L2_PERSIST_MB=<n> additionally sets n MB of L2 aside for
  // such lines (0 / unset: the limit is left alone; values above the device maximum are clamped).
  {
    const char* lp = getenv("PDC_L2_PERSIST_MB");
    const long  mb = lp ? atol(lp) : 0;
    if (mb > 0 && prop.persistingL2CacheMaxSize > 0) {
      const size_t want = std::min((size_t)mb << 20, (size_t)prop.persistingL2CacheMaxSize);
      PDC_CREATE(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
    }
  }
  ctx->queues.resize(cfg->nof_streams);
  ctx->queue_mutex.reset(new std::mutex[cfg->nof_streams]);
  for (Queue& q : ctx->queues) {
    PDC_CREATE(create_queue(ctx, q, cfg->max_cbs, cfg->max_llrs, cfg->max_tbs, cfg->max_tb_bytes, true));
  }
  PDC_CREATE(create_queue(ctx, ctx->sync_q, 1, 16, 1, 64, false));
#undef PDC_CREATE
  return PDC_OK;
}

static void free_front_end(FrontEnd& fe)
{
  PDC_FREE(fe.dm.d_sym);
  PDC_FREE(fe.dm.d_nv);
  PDC_FREE(fe.d_raw);
  PDC_FREE(fe.d_seq);
  PDC_FREE(fe.d_uci);
  PDC_FREE(fe.d_plan);
  PDC_FREE(fe.d_cb_scr);
  cudaFreeHost(fe.h_uci);
  for (int k = 0; k != FrontEnd::RING; ++k) {
    cudaFreeHost(fe.h_plan[k]);
    if (fe.plan_ev[k]) {
      cudaEventDestroy(fe.plan_ev[k]);
    }
  }
  fe = FrontEnd();
}

static void destroy_queue(Queue& q)
{
  if (q.stream) {
    cudaStreamSynchronize(q.stream);
    cudaStreamDestroy(q.stream);
  }
  if (q.done) {
    cudaEventDestroy(q.done);
  }
  if (q.copy_stream) {
    cudaStreamSynchronize(q.copy_stream);
    cudaStreamDestroy(q.copy_stream);
  }
  for (int l = 0; l != PIPE_LANES; ++l) {
    if (q.lane[l]) {
      cudaStreamSynchronize(q.lane[l]);
      cudaStreamDestroy(q.lane[l]);
    }
    if (q.ev_lane[l]) {
      cudaEventDestroy(q.ev_lane[l]);
    }
  }
  for (int g = 0; g != PIPE_MAX_GROUPS; ++g) {
    if (q.ev_copy[g]) {
      cudaEventDestroy(q.ev_copy[g]);
    }
  }
  PDC_FREE(q.d_desc);
  PDC_FREE(q.d_res);
  PDC_FREE(q.d_llrs);
  PDC_FREE(q.d_cb_bits);
  PDC_FREE(q.d_tb_out);
  PDC_FREE(q.d_tb_sync);
  free_front_end(q.fe);
  cudaFreeHost(q.h_desc);
  cudaFreeHost(q.h_res);
  cudaFreeHost(q.h_cb_bits);
  cudaFreeHost(q.h_tb_out);
}

void pdc_destroy(pdc_ctx* ctx)
{
  if (!ctx) {
    return;
  }
  cudaSetDevice(ctx->cfg.device);
  free_front_end(ctx->fe_sync);
  PDC_FREE(ctx->d_sch_sync);
  for (Queue& q : ctx->queues) {
    destroy_queue(q);
  }
  destroy_queue(ctx->sync_q);
  PDC_FREE(ctx->d_harq);
  PDC_FREE(ctx->d_harq_data);
  PDC_FREE(ctx->d_harq_last);
  PDC_FREE(ctx->d_tb_sync);
  PDC_FREE(ctx->d_scratch_llr);
  PDC_FREE(ctx->d_demod_tables);
  PDC_FREE(ctx->d_enc_cbs);
  PDC_FREE(ctx->d_enc_msgs);
  PDC_FREE(ctx->d_enc_out);
  for (pdc_ctx::DecodeScratch& c : ctx->decode_scratch) {
    PDC_FREE(c.d_state);
    PDC_FREE(c.d_counter);
  }
  delete ctx;
}

int pdc_device_info(pdc_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor)
{
  if (!ctx) {
    return fail(PDC_ERR_INVALID, "pdc_device_info: null context");
  }
  if (sm_count) {
    *sm_count = ctx->sm_count;
  }
  if (cc_major) {
    *cc_major = ctx->cc_major;
  }
  if (cc_minor) {
    *cc_minor = ctx->cc_minor;
  }
  return PDC_OK;
}

uint64_t pdc_launch_count(pdc_ctx* ctx)
{
  return ctx ? ctx->launches.load() : 0;
}

int pdc_debug_canaries_ok(pdc_ctx* ctx)
{
#ifdef PDC_DEBUG_BOUNDS
  if (!ctx || cudaSetDevice(ctx->cfg.device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    return 0;
  }
  return canary_check(ctx->cfg.device) == 0 ? 1 : 0;
#else
  (void)ctx;
  return -1; // not a debug build: nothing to check
#endif
}

void* pdc_host_alloc(size_t bytes)
{
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    return nullptr;
  }
  return p;
}

void* pdc_host_alloc_input(size_t bytes)
{
  // Write-combined and portable: the CPU only ever WRITES soft bits here (reads would be slow), the copy engine reads
  // them without snooping the CPU caches, and every context of the process (one per GPU) may copy from it.
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocWriteCombined | cudaHostAllocPortable) != cudaSuccess) {
    return nullptr;
  }
  return p;
}

void pdc_host_free(void* p)
{
  if (p) {
    cudaFreeHost(p);
  }
}

static int front_end_kernels(pdc_ctx* ctx, FrontEnd& fe, cudaStream_t s);
static int demod_kernel_launch(pdc_ctx* ctx, DemodStage& dm, cudaStream_t s);

// Page-locked host memory (pdc_host_alloc, cudaHostAlloc, cudaHostRegister) can be the target of an asynchronous copy.
static bool is_pinned_host(const void* p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}


// ---- pipelined submission -----------------------------------------------------------------------------------------------
//
// One large batch in one piece costs copy-in + kernels + copy-out in series (a 16-cell slot: 440 + 330 + 50 us). Here
// the soft bits travel in groups of whole transport blocks (or, inside one large transport block, of codeblock pairs) on
// a copy stream, each group's rate dematcher / decoder / TB assembly run on one of PIPE_LANES streams as soon as the
// group has arrived, and its results leave as soon as they exist; the batch completes one group of kernels after its
// last byte came in. Everything still belongs to the one queue: pdc_wait sees it complete when every lane has.
struct PipeGroup {
  uint32_t cb0, cb1; // codeblocks [cb0, cb1)
  uint32_t tb0, tb1; // transport blocks that lie entirely in the group
  size_t   lo, hi;   // soft-bit bytes [lo, hi)
};
struct PipePlan {
  int       n_groups = 0;
  PipeGroup g[PIPE_MAX_GROUPS];
  bool      tail_tbs = false; // transport blocks that span groups: assembled after the lanes have joined
};

// false: the batch goes in one piece (small, or its soft bits are not laid out in codeblock order).
static bool plan_pipeline(const pdc_cb_desc* cbs, uint32_t n_cb, size_t n_llrs, const pdc_tb_desc* tbs, uint32_t n_tb,
                          PipePlan& plan)
{
  static const size_t min_bytes = [] {
    const char* e = getenv("PDC_PIPE_MIN_KB"); // measurement aid: smallest batch that is pipelined
    return e ? (size_t)atol(e) << 10 : PIPE_MIN_BYTES;
  }();
  if (n_llrs < min_bytes || n_cb < 4) {
    return false;
  }
  // Soft bits in codeblock order without overlap (what a demodulator emits), every codeblock dematched from them.
  size_t end = 0;
  for (uint32_t i = 0; i != n_cb; ++i) {
    if (!(cbs[i].flags & PDC_CB_DEMATCH) || cbs[i].llr_offset < end) {
      return false;
    }
    end = (size_t)cbs[i].llr_offset + cbs[i].rm_length;
  }
  // Transport blocks in order, each a run of consecutive codeblocks.
  uint32_t next_cb = 0;
  for (uint32_t t = 0; t != n_tb; ++t) {
    if (tbs[t].first_cb < next_cb) {
      return false;
    }
    next_cb = tbs[t].first_cb + tbs[t].nof_cb;
  }
  // Group sizes fall towards the end of the batch (5 4 3 2 1 1 sixteenths): what is left to do when the last byte has
  // arrived is the kernel chain of the LAST group, so that one is the smallest; every copy costs about 10 us on top of
  // its bytes, so there are no more groups than that takes (16-cell slot, p50: 3 3 3 2 2 1 1 1 -> 569 us, 4 4 3 2 2 1 ->
  // 559, 5 4 3 2 1 1 -> 554, 6 5 3 1 1 -> 558, 4 4 4 3 1 -> 570, eight equal groups -> 574, 6 6 3 1 -> 586).
  // (PDC_PIPE_WEIGHTS="5,4,3,2,1,1": measurement aid, sixteenths per group, the last one repeated.)
  static const struct Weights {
    int w[PIPE_GROUPS] = {5, 4, 3, 2, 1, 1, 1, 1};
    Weights()
    {
      const char* e = getenv("PDC_PIPE_WEIGHTS");
      if (e != nullptr) {
        int k = 0, last = 1;
        while (*e != 0 && k != PIPE_GROUPS) {
          last   = std::max(1, atoi(e));
          w[k++] = last;
          while (*e != 0 && *e != ',') {
            ++e;
          }
          e += (*e == ',') ? 1 : 0;
        }
        for (; k != PIPE_GROUPS; ++k) {
          w[k] = last;
        }
      }
    }
  } weights;
  const int* weight = weights.w;
  // Cut points: between transport blocks when there are several, else between codeblock pairs.
  uint32_t cb0 = 0, t_next = 0;
  size_t   lo = 0;
  while (cb0 != n_cb && plan.n_groups != PIPE_MAX_GROUPS) {
    uint32_t cb1 = cb0;
    size_t   hi  = lo;
    const bool   last_slot = plan.n_groups == PIPE_MAX_GROUPS - 1;
    const size_t target    = std::max(PIPE_MIN_GROUP, n_llrs * (size_t)weight[std::min(plan.n_groups, PIPE_GROUPS - 1)] / 16);
    while (cb1 != n_cb && (last_slot || hi - lo < target)) {
      if (n_tb > 1 && t_next != n_tb && tbs[t_next].first_cb <= cb1) {
        // take the whole transport block
        cb1 = std::max(cb1, tbs[t_next].first_cb + tbs[t_next].nof_cb);
        ++t_next;
      } else if (n_tb > 1 && t_next != n_tb) {
        cb1 = tbs[t_next].first_cb; // codeblocks outside any transport block, up to the next one
      } else {
        cb1 = std::min(n_cb, cb1 + 2);
      }
      hi = (size_t)cbs[cb1 - 1].llr_offset + cbs[cb1 - 1].rm_length;
    }
    PipeGroup& g = plan.g[plan.n_groups++];
    g.cb0 = cb0, g.cb1 = cb1, g.lo = lo, g.hi = (cb1 == n_cb) ? n_llrs : hi;
    cb0 = cb1;
    lo  = g.hi;
  }
  if (plan.n_groups < 2) {
    return false;
  }
  // Transport blocks inside one group are assembled there; the others after the join.
  uint32_t t = 0;
  for (int k = 0; k != plan.n_groups; ++k) {
    PipeGroup& g = plan.g[k];
    while (t != n_tb && tbs[t].first_cb + tbs[t].nof_cb <= g.cb0) {
      plan.tail_tbs = true; // (a block that ended before this group without lying inside an earlier one)
      ++t;
    }
    g.tb0 = t;
    while (t != n_tb && tbs[t].first_cb >= g.cb0 && tbs[t].first_cb + tbs[t].nof_cb <= g.cb1) {
      ++t;
    }
    g.tb1 = t;
    if (t != n_tb && tbs[t].first_cb < g.cb1) {
      // this block continues in the next group: everything from here on is assembled after the join
      plan.tail_tbs = true;
      for (int k2 = k + 1; k2 != plan.n_groups; ++k2) {
        plan.g[k2].tb0 = plan.g[k2].tb1 = 0;
      }
      break;
    }
  }
  if (t != n_tb) {
    plan.tail_tbs = true;
  }
  return true;
}

static int submit_pipelined(pdc_ctx*           ctx,
                            Queue&             q,
                            const PipePlan&    plan,
                            const BatchShape&  shape,
                            const pdc_cb_desc* cbs,
                            uint32_t           n_cb,
                            const int8_t*      llrs,
                            size_t             n_llrs,
                            const pdc_tb_desc* tbs,
                            uint32_t           n_tb,
                            pdc_cb_result*     cb_results,
                            uint8_t*           cb_bits,
                            pdc_tb_result*     tb_results,
                            uint8_t*           tb_bytes,
                            size_t             tb_out_bytes)
{
  (void)n_llrs;
  memcpy(q.h_cbs, cbs, sizeof(pdc_cb_desc) * n_cb);
  cudaStream_t cs = q.copy_stream;
  if (n_tb != 0) {
    memcpy(q.h_tbs, tbs, sizeof(pdc_tb_desc) * n_tb);
    PDC_CUDA(cudaMemcpyAsync(q.d_desc, q.h_desc, q.tb_desc_area + sizeof(pdc_cb_desc) * n_cb, cudaMemcpyHostToDevice, cs));
  } else {
    PDC_CUDA(cudaMemcpyAsync(q.d_cbs, q.h_cbs, sizeof(pdc_cb_desc) * n_cb, cudaMemcpyHostToDevice, cs));
  }
  const bool direct_bits = cb_bits && is_pinned_host(cb_bits);
  const bool direct_tb   = tb_bytes && n_tb != 0 && is_pinned_host(tb_bytes);
  uint8_t*   bits_dst    = cb_bits ? (direct_bits ? cb_bits : q.h_cb_bits) : nullptr;
  // (The TB assembly kernels write their 4-byte RESULTS straight into the page-locked host mirror: one small copy less
  // behind the last kernel of a lane. The transport-block bytes take a copy: see pdc_submit.)
  uint8_t*   tb_dst      = (tb_bytes && n_tb != 0) ? (direct_tb ? tb_bytes : q.h_tb_out) : nullptr;
  // All copies first (the copy engine then runs without gaps while the host queues the kernels of the groups behind
  // them: a launch costs a few microseconds of host time, a group's copy tens).
  static const bool trace = [] {
    const char* e = getenv("PDC_PIPE_TRACE");
    return e != nullptr && e[0] == '1';
  }();
  const auto host_t0 = std::chrono::steady_clock::now();
  if (trace) {
    if (q.tr_start == nullptr) {
      cudaEventCreate(&q.tr_start);
      cudaEventCreate(&q.tr_done);
      for (int k = 0; k != PIPE_MAX_GROUPS; ++k) {
        cudaEventCreate(&q.tr_copy[k]);
        cudaEventCreate(&q.tr_group[k]);
        cudaEventCreate(&q.tr_kern[k]);
      }
    }
    cudaEventRecord(q.tr_start, cs);
  }
  for (int k = 0; k != plan.n_groups; ++k) {
    const PipeGroup& g = plan.g[k];
    PDC_CUDA(cudaMemcpyAsync(q.d_llrs + g.lo, llrs + g.lo, g.hi - g.lo, cudaMemcpyHostToDevice, cs));
    PDC_CUDA(cudaEventRecord(q.ev_copy[k], cs));
    if (trace) {
      cudaEventRecord(q.tr_copy[k], cs);
    }
  }
  for (int k = 0; k != plan.n_groups; ++k) {
    const PipeGroup& g = plan.g[k];
    cudaStream_t ls = q.lane[k % PIPE_LANES];
    PDC_CUDA(cudaStreamWaitEvent(ls, q.ev_copy[k], 0));
    const uint32_t   n_g  = g.cb1 - g.cb0, n_t = g.tb1 - g.tb0;
    const BatchShape gs   = scan_batch(cbs + g.cb0, n_g);
    int rc = launch_batch(ctx, q.d_cbs + g.cb0, n_g, q.d_llrs, n_t ? q.d_tbs + g.tb0 : nullptr, n_t, q.d_cb_res + g.cb0,
                          q.d_cb_bits + (size_t)g.cb0 * PDC_MAX_CB_BYTES, n_t ? q.h_tb_res + g.tb0 : nullptr, q.d_tb_out,
                          gs, nullptr, 0, ls, q.d_tb_sync + 2 * (size_t)g.tb0, nullptr, q.d_cbs, q.d_cb_res, q.d_cb_bits);
    if (rc != PDC_OK) {
      cudaDeviceSynchronize();
      return rc;
    }
    if (trace) {
      cudaEventRecord(q.tr_kern[k], ls);
    }
    // What the group produced leaves right away.
    PDC_CUDA(cudaMemcpyAsync(q.h_cb_res + g.cb0, q.d_cb_res + g.cb0, sizeof(pdc_cb_result) * n_g, cudaMemcpyDeviceToHost,
                             ls));
    if (bits_dst) {
      PDC_CUDA(cudaMemcpyAsync(bits_dst + (size_t)g.cb0 * PDC_MAX_CB_BYTES, q.d_cb_bits + (size_t)g.cb0 * PDC_MAX_CB_BYTES,
                               (size_t)n_g * PDC_MAX_CB_BYTES, cudaMemcpyDeviceToHost, ls));
    }
    if (tb_dst && n_t != 0) {
      size_t lo = (size_t)-1, hi = 0;
      for (uint32_t t = g.tb0; t != g.tb1; ++t) {
        lo = std::min(lo, (size_t)tbs[t].out_offset);
        hi = std::max(hi, (size_t)tbs[t].out_offset + ((size_t)tbs[t].tbs_bits + 24 + 31) / 32 * 4);
      }
      PDC_CUDA(cudaMemcpyAsync(tb_dst + lo, q.d_tb_out + lo, hi - lo, cudaMemcpyDeviceToHost, ls));
    }
    if (trace) {
      cudaEventRecord(q.tr_group[k], ls);
    }
  }
  // Join: the queue's own stream continues when every lane is through.
  for (int l = 0; l != std::min(plan.n_groups, PIPE_LANES); ++l) {
    PDC_CUDA(cudaEventRecord(q.ev_lane[l], q.lane[l]));
    PDC_CUDA(cudaStreamWaitEvent(q.stream, q.ev_lane[l], 0));
  }
  if (plan.tail_tbs) {
    // Transport blocks whose codeblocks were spread over groups (one large block, typically): assembled now. The blocks
    // assembled inside a group keep their results (their descriptors are simply skipped).
    std::vector<char> in_group(n_tb, 0);
    for (int k = 0; k != plan.n_groups; ++k) {
      for (uint32_t t = plan.g[k].tb0; t != plan.g[k].tb1; ++t) {
        in_group[t] = 1;
      }
    }
    uint32_t t = 0;
    while (t != n_tb) {
      if (in_group[t]) {
        ++t;
        continue;
      }
      uint32_t t1 = t;
      while (t1 != n_tb && !in_group[t1]) {
        ++t1;
      }
      pdc::TbParams tp;
      tp.tbs        = q.d_tbs + t;
      tp.n_tb       = t1 - t;
      tp.cbs        = q.d_cbs;
      tp.cb_results = q.d_cb_res;
      tp.cb_bits    = q.d_cb_bits;
      tp.tb_results = q.h_tb_res + t;
      tp.tb_bytes   = q.d_tb_out;
      PDC_CUDA(pdc::launch_tb_assemble(tp, ctx->d_harq_data, q.d_tb_sync + 2 * (size_t)t, q.stream));
      ctx->launches++;
      if (tb_dst) {
        size_t lo = (size_t)-1, hi = 0;
        for (uint32_t u = t; u != t1; ++u) {
          lo = std::min(lo, (size_t)tbs[u].out_offset);
          hi = std::max(hi, (size_t)tbs[u].out_offset + ((size_t)tbs[u].tbs_bits + 24 + 31) / 32 * 4);
        }
        PDC_CUDA(cudaMemcpyAsync(tb_dst + lo, q.d_tb_out + lo, hi - lo, cudaMemcpyDeviceToHost, q.stream));
      }
      t = t1;
    }
  }
  PDC_CUDA(cudaEventRecord(q.done, q.stream));
  if (trace) {
    cudaEventRecord(q.tr_done, q.stream);
    q.tr_groups  = plan.n_groups;
    q.tr_host_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - host_t0).count();
  }
  q.busy         = true;
  q.n_cb         = n_cb;
  q.n_tb         = n_tb;
  q.tb_out_bytes = tb_out_bytes;
  q.u_cb_res     = cb_results;
  q.u_cb_bits    = direct_bits ? nullptr : cb_bits;
  q.u_tb_res     = tb_results;
  q.u_tb_out     = direct_tb ? nullptr : tb_bytes;
  (void)shape;
  return PDC_OK;
}

int pdc_submit(pdc_ctx*           ctx,
               uint32_t           stream,
               const pdc_cb_desc* cbs,
               uint32_t           n_cb,
               const int8_t*      llrs,
               size_t             n_llrs,
               const pdc_tb_desc* tbs,
               uint32_t           n_tb,
               pdc_cb_result*     cb_results,
               uint8_t*           cb_bits,
               pdc_tb_result*     tb_results,
               uint8_t*           tb_bytes)
{
  if (!ctx || stream >= ctx->queues.size() || !cbs || n_cb == 0 || !cb_results || (n_tb != 0 && (!tbs || !tb_results))) {
    return fail(PDC_ERR_INVALID, "pdc_submit: invalid argument");
  }
  if (n_cb > ctx->cfg.max_cbs || n_llrs > ctx->cfg.max_llrs || n_tb > ctx->cfg.max_tbs) {
    return fail(PDC_ERR_CAPACITY, "pdc_submit: batch exceeds the capacity of the context");
  }
  // One thread at a time per queue; a second thread arriving while a submit is being queued gets an error instead of a
  // corrupted descriptor block.
  std::unique_lock<std::mutex> queue_lock(ctx->queue_mutex[stream], std::try_to_lock);
  if (!queue_lock.owns_lock()) {
    return fail(PDC_ERR_CAPACITY, "pdc_submit: queue is being used by another thread");
  }
  Queue& q = ctx->queues[stream];
  if (q.busy) {
    return fail(PDC_ERR_CAPACITY, "pdc_submit: queue busy (call pdc_wait first)");
  }
  if (llrs == nullptr && n_llrs == 0 && q.fe.pending) {
    // The rate-matched LLRs are the UL-SCH soft bits pdc_submit_codewords left on the device.
    n_llrs = q.fe.sch_len;
  } else if (llrs == nullptr && n_llrs != 0) {
    return fail(PDC_ERR_INVALID, "pdc_submit: no LLR buffer");
  }
  // Validate what the kernels index with before anything is queued.
  for (uint32_t i = 0; i != n_cb; ++i) {
    if ((size_t)cbs[i].llr_offset + cbs[i].rm_length > n_llrs && (cbs[i].flags & PDC_CB_DEMATCH)) {
      return fail(PDC_ERR_INVALID, "pdc_submit: codeblock LLR range outside the batch");
    }
    if (cbs[i].harq_id >= ctx->cfg.harq_entries) {
      return fail(PDC_ERR_INVALID, "pdc_submit: harq_id outside the arena");
    }
    if (cbs[i].tb_index != 0xffff && cbs[i].tb_index >= n_tb) {
      return fail(PDC_ERR_INVALID, "pdc_submit: codeblock refers to a transport block outside the batch");
    }
  }
  size_t tb_out_bytes = 0;
  for (uint32_t i = 0; i != n_tb; ++i) {
    size_t need = ((size_t)tbs[i].tbs_bits + 24 + 31) / 32 * 4;
    if ((tbs[i].out_offset & 3u) || (size_t)tbs[i].out_offset + need > ctx->cfg.max_tb_bytes ||
        (size_t)tbs[i].first_cb + tbs[i].nof_cb > n_cb || tbs[i].nof_cb == 0 || need > 4u * 65535u) {
      return fail(PDC_ERR_INVALID, "pdc_submit: invalid transport block descriptor");
    }
    tb_out_bytes = std::max(tb_out_bytes, (size_t)tbs[i].out_offset + need);
    // The assembly kernel walks the transport block in steps of n_data = K - 24 - F bits per codeblock: every
    // codeblock of the block must have the same, valid shape and together they must hold payload + TB checksum
    // (ldpc_segmenter_impl.cpp:254-306 produces exactly such sets).
    const pdc_cb_desc& c0 = cbs[tbs[i].first_cb];
    const int          Z0 = c0.lifting_size;
    if ((c0.base_graph != 1 && c0.base_graph != 2) || Z0 < 2 || Z0 > pdc::MAX_Z ||
        pdc::host_tables().set_index[Z0] == 0xff) {
      return fail(PDC_ERR_INVALID, "pdc_submit: transport block with an invalid base graph / lifting size");
    }
    const uint32_t K0 = ((c0.base_graph == 1) ? 22u : 10u) * (uint32_t)Z0;
    for (uint32_t k = 1; k < tbs[i].nof_cb; ++k) {
      const pdc_cb_desc& ck = cbs[tbs[i].first_cb + k];
      if (ck.base_graph != c0.base_graph || ck.lifting_size != c0.lifting_size || ck.nof_filler != c0.nof_filler) {
        return fail(PDC_ERR_INVALID, "pdc_submit: the codeblocks of a transport block differ in shape");
      }
    }
    if (tbs[i].nof_cb == 1) {
      // a single codeblock carries the TB checksum as its own (CRC16 or CRC24A)
      const uint32_t crc_bits = (c0.crc_kind == PDC_CRC16) ? 16u : 24u;
      if ((uint32_t)c0.nof_filler + crc_bits >= K0 || tbs[i].tbs_bits + crc_bits > K0 - c0.nof_filler) {
        return fail(PDC_ERR_INVALID, "pdc_submit: transport block larger than its codeblock");
      }
    } else if ((uint32_t)c0.nof_filler + 24u >= K0 ||
               (uint64_t)tbs[i].tbs_bits + 24u > (uint64_t)tbs[i].nof_cb * (K0 - 24u - c0.nof_filler)) {
      return fail(PDC_ERR_INVALID, "pdc_submit: transport block larger than its codeblocks");
    }
  }
  PDC_CUDA(cudaSetDevice(ctx->cfg.device));
  BatchShape shape = scan_batch(cbs, n_cb);
  {
    PipePlan plan;
    if (llrs != nullptr && !q.fe.pending && q.copy_stream != nullptr && !ctx->no_pipeline &&
        plan_pipeline(cbs, n_cb, n_llrs, tbs, n_tb, plan)) {
      return submit_pipelined(ctx, q, plan, shape, cbs, n_cb, llrs, n_llrs, tbs, n_tb, cb_results, cb_bits, tb_results,
                              tb_bytes, tb_out_bytes);
    }
  }
  memcpy(q.h_cbs, cbs, sizeof(pdc_cb_desc) * n_cb);
  if (n_tb != 0) {
    // Transport-block and codeblock descriptors in one copy (they are one block: {TB area, codeblocks}).
    memcpy(q.h_tbs, tbs, sizeof(pdc_tb_desc) * n_tb);
    PDC_CUDA(cudaMemcpyAsync(q.d_desc, q.h_desc, q.tb_desc_area + sizeof(pdc_cb_desc) * n_cb, cudaMemcpyHostToDevice,
                             q.stream));
  } else {
    PDC_CUDA(cudaMemcpyAsync(q.d_cbs, q.h_cbs, sizeof(pdc_cb_desc) * n_cb, cudaMemcpyHostToDevice, q.stream));
  }
  if (n_llrs != 0 && llrs != nullptr) {
    PDC_CUDA(cudaMemcpyAsync(q.d_llrs, llrs, n_llrs, cudaMemcpyHostToDevice, q.stream));
  }
  // Codeblocks that are not decoded report "not run": the rate dematcher clears the result array when it runs.
  if (!shape.any_dematch) {
    PDC_CUDA(cudaMemsetAsync(q.d_cb_res, 0, sizeof(pdc_cb_result) * n_cb, q.stream));
  }
  if (q.fe.pending) {
    int rc_fe = front_end_kernels(ctx, q.fe, q.stream);
    if (rc_fe != PDC_OK) {
      return rc_fe;
    }
  }
  int rc = launch_batch(ctx, q.d_cbs, n_cb, q.d_llrs, q.d_tbs, n_tb, q.d_cb_res, q.d_cb_bits, q.d_tb_res, q.d_tb_out,
                        shape, nullptr, 0, q.stream, q.d_tb_sync, q.fe.pending ? &q.fe : nullptr);
  if (rc != PDC_OK) {
    return rc;
  }
  if (n_tb != 0) {
    // Transport-block and codeblock results in one copy.
    PDC_CUDA(cudaMemcpyAsync(q.h_res, q.d_res, q.tb_res_area + sizeof(pdc_cb_result) * n_cb, cudaMemcpyDeviceToHost, q.stream));
  } else {
    PDC_CUDA(cudaMemcpyAsync(q.h_cb_res, q.d_cb_res, sizeof(pdc_cb_result) * n_cb, cudaMemcpyDeviceToHost, q.stream));
  }
  // Outputs go straight to page-locked caller buffers; pageable ones are filled from the pinned staging in pdc_wait.
  // (Letting the TB assembly kernel write the transport-block bytes into the page-locked host buffer itself was
  // measured: its per-thread runs of words make small PCIe writes - a 16-cell slot 580 -> 1210 us, one cell 108 -> 150.)
  const bool direct_bits = cb_bits && is_pinned_host(cb_bits);
  const bool direct_tb   = tb_bytes && n_tb != 0 && is_pinned_host(tb_bytes);
  if (cb_bits) {
    PDC_CUDA(cudaMemcpyAsync(direct_bits ? cb_bits : q.h_cb_bits, q.d_cb_bits, (size_t)n_cb * PDC_MAX_CB_BYTES,
                             cudaMemcpyDeviceToHost, q.stream));
  }
  if (n_tb != 0) {
    if (tb_bytes) {
      PDC_CUDA(cudaMemcpyAsync(direct_tb ? tb_bytes : q.h_tb_out, q.d_tb_out, tb_out_bytes, cudaMemcpyDeviceToHost,
                               q.stream));
    }
  }
  PDC_CUDA(cudaEventRecord(q.done, q.stream));
  q.busy         = true;
  q.n_cb         = n_cb;
  q.n_tb         = n_tb;
  q.tb_out_bytes = tb_out_bytes;
  q.u_cb_res     = cb_results;
  q.u_cb_bits    = direct_bits ? nullptr : cb_bits;
  q.u_tb_res     = tb_results;
  q.u_tb_out     = direct_tb ? nullptr : tb_bytes;
  return PDC_OK;
}


// ---- codeword front end ------------------------------------------------------------------------------------------------

// Queues the kernels of the plan front_end_launch prepared (and the copy of the UCI soft bits to the host).
static int front_end_kernels(pdc_ctx* ctx, FrontEnd& fe, cudaStream_t s)
{
  if (!fe.kernels_pending) {
    return PDC_OK;
  }
  fe.kernels_pending       = false;
  if (fe.dm.kernel_pending) {
    // The codewords come from the soft demapper (pdc_submit_symbols).
    int rc = demod_kernel_launch(ctx, fe.dm, s);
    if (rc != PDC_OK) {
      return rc;
    }
  }
  const pdc::UlschArgs& a  = fe.args;
  const uint32_t        n_cw = fe.k_n_cw;
  if (fe.k_generate_seq) {
    // One polynomial jump per thread: few sequence words per thread while the batch cannot fill the GPU (the kernel is
    // then bound by the latency of one thread's chain), more once it can (the jumps are then what costs).
    const uint32_t words = (fe.k_max_in + 31) / 32;
    const bool     small = (size_t)words * n_cw / pdc::PRG_WORDS_PER_THREAD < (size_t)256 * ctx->sm_count;
    const uint32_t per_cta = 128 * (small ? pdc::PRG_WORDS_PER_THREAD_SMALL : pdc::PRG_WORDS_PER_THREAD);
    dim3           grid((words + per_cta - 1) / per_cta, n_cw);
    if (small) {
      pdc::prg_kernel<pdc::PRG_WORDS_PER_THREAD_SMALL><<<grid, 128, 0, s>>>(a.cws, fe.d_seq);
    } else {
      pdc::prg_kernel<pdc::PRG_WORDS_PER_THREAD><<<grid, 128, 0, s>>>(a.cws, fe.d_seq);
    }
    PDC_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  if (!fe.k_all_deferred) {
    // Enough CTAs to fill the GPU a few times over; each thread steps through the chunks of its codeword.
    const uint32_t chunks = (fe.k_max_sch + 15) / 16;
    uint32_t       gx = std::max(1u, std::min((chunks + 255) / 256, (uint32_t)(8 * ctx->sm_count + n_cw - 1) / n_cw));
    PDC_CUDA(pdc::launch_pdl(pdc::ulsch_sch_kernel, dim3(gx, n_cw), dim3(256), 0, s, a));
    ctx->launches++;
  }
  if (fe.k_max_uci != 0) {
    PDC_CUDA(pdc::launch_pdl(pdc::ulsch_uci_kernel, dim3((fe.k_max_uci + 255) / 256, n_cw), dim3(256), 0, s, a));
    ctx->launches++;
  }
  if (fe.uci_copy_dst != nullptr && fe.uci_bytes != 0) {
    PDC_CUDA(cudaMemcpyAsync(fe.uci_copy_dst, fe.d_uci, fe.uci_bytes, cudaMemcpyDeviceToHost, s));
  }
  fe.uci_copy_dst = nullptr;
  return PDC_OK;
}

// Plans the codewords, uploads the plan and queues the kernels on stream s:
//   d_in (raw soft bits, already on the device) -> d_sch (UL-SCH space) and fe.d_uci (UCI area of n_uci bytes).
// seq_words != nullptr: caller-supplied scrambling sequence (fe.plan layout), else generated from c_init.
static int front_end_launch(pdc_ctx*           ctx,
                            FrontEnd&          fe,
                            const pdc_cw_desc* cws,
                            uint32_t           n_cw,
                            size_t             n_in,
                            const int8_t*      d_in,
                            int8_t*            d_sch,
                            size_t             sch_capacity,
                            size_t             uci_capacity,
                            const uint8_t*     seq_bits_packed,
                            pdc_cw_result*     results,
                            cudaStream_t       s,
                            int8_t*            d_uci_user = nullptr,
                            bool               launch_now = true)
{
  pdc::UlschPlan& plan = fe.plan;
  plan.clear();
  size_t sch_end = 0, uci_end = 0;
  bool   any_deferred = false, all_deferred = true;
  for (uint32_t i = 0; i != n_cw; ++i) {
    if (!pdc::ulsch_plan_codeword(cws[i], plan)) {
      return fail(PDC_ERR_INVALID, "codeword front end: inconsistent codeword description");
    }
    pdc::UlschCodeword& cw = plan.cws.back();
    if ((cws[i].flags & PDC_CW_DEFER_DESCRAMBLING) && (cws[i].flags & PDC_CW_SCRAMBLED) && cw.n_out[0] == cw.n_in) {
      // No UCI: the UL-SCH stream is the input itself; leave it scrambled for the rate dematcher.
      cw.flags |= pdc::ULSCH_CW_DEFERRED;
      any_deferred = true;
    } else {
      all_deferred = false;
    }
    if ((cws[i].sch_offset & 3u) || (size_t)cw.in_off + cw.n_in > n_in) {
      return fail(PDC_ERR_INVALID, "codeword front end: codeword outside the input or misaligned UL-SCH offset");
    }
    results[i].n_sch       = cw.n_out[0];
    results[i].n_harq_ack  = cw.n_out[1];
    results[i].n_csi_part1 = cw.n_out[2];
    results[i].n_csi_part2 = cw.n_out[3];
    sch_end = std::max(sch_end, (size_t)cw.sch_off + cw.n_out[0]);
    uci_end = std::max(uci_end, (size_t)cw.uci_off + cw.n_out[1] + cw.n_out[2] + cw.n_out[3]);
  }
  if (sch_end > sch_capacity || uci_end > uci_capacity) {
    return fail(PDC_ERR_CAPACITY, "codeword front end: output buffer too small");
  }
  // Plan image: codewords | symbols | element lists.
  const size_t b_cws  = plan.cws.size() * sizeof(pdc::UlschCodeword);
  const size_t b_syms = plan.syms.size() * sizeof(pdc::UlschSymbol);
  const size_t b_list = ((plan.lists.size() * sizeof(uint16_t)) + 15) & ~(size_t)15;
  const size_t bytes  = b_cws + b_syms + b_list;
  PDC_CUDA(grow_device(&fe.d_plan, &fe.plan_cap, bytes));
  const int slot = fe.ring_pos;
  fe.ring_pos    = (fe.ring_pos + 1) % FrontEnd::RING;
  if (fe.plan_ev[slot] == nullptr) {
    PDC_CUDA(cudaEventCreateWithFlags(&fe.plan_ev[slot], cudaEventDisableTiming));
  } else {
    PDC_CUDA(cudaEventSynchronize(fe.plan_ev[slot]));
  }
  if (bytes > fe.h_plan_cap[slot]) {
    cudaFreeHost(fe.h_plan[slot]);
    fe.h_plan[slot]     = nullptr;
    fe.h_plan_cap[slot] = 0;
    PDC_CUDA(cudaMallocHost(reinterpret_cast<void**>(&fe.h_plan[slot]), bytes + bytes / 4 + 64));
    fe.h_plan_cap[slot] = bytes + bytes / 4 + 64;
  }
  unsigned char* h_plan = fe.h_plan[slot];
  memcpy(h_plan, plan.cws.data(), b_cws);
  memcpy(h_plan + b_cws, plan.syms.data(), b_syms);
  memcpy(h_plan + b_cws + b_syms, plan.lists.data(), plan.lists.size() * sizeof(uint16_t));
  PDC_CUDA(cudaMemcpyAsync(fe.d_plan, h_plan, bytes, cudaMemcpyHostToDevice, s));
  PDC_CUDA(cudaEventRecord(fe.plan_ev[slot], s));
  PDC_CUDA(grow_device(&fe.d_seq, &fe.seq_cap, (size_t)plan.seq_words + 2 * pdc::PRG_WORDS_PER_THREAD));
  if (d_uci_user == nullptr && uci_end > fe.uci_cap) {
    cudaFreeHost(fe.h_uci);
    fe.h_uci = nullptr;
    PDC_CUDA(grow_device(&fe.d_uci, &fe.uci_cap, uci_end));
    PDC_CUDA(cudaMallocHost(reinterpret_cast<void**>(&fe.h_uci), fe.uci_cap));
  }
  pdc::UlschArgs& a = fe.args;
  a.cws   = reinterpret_cast<const pdc::UlschCodeword*>(fe.d_plan);
  a.syms  = reinterpret_cast<const pdc::UlschSymbol*>(fe.d_plan + b_cws);
  a.lists = reinterpret_cast<const uint16_t*>(fe.d_plan + b_cws + b_syms);
  a.seq   = fe.d_seq;
  a.in    = d_in;
  a.sch   = d_sch;
  a.uci   = d_uci_user ? d_uci_user : fe.d_uci;
  uint32_t max_in = 0, max_sch = 0, max_uci = 0;
  for (const pdc::UlschCodeword& cw : plan.cws) {
    max_in  = std::max(max_in, cw.n_in);
    max_sch = std::max(max_sch, cw.n_out[0]);
    max_uci = std::max(max_uci, cw.n_out[1] + cw.n_out[2] + cw.n_out[3]);
  }
  if (seq_bits_packed != nullptr) {
    // Caller-supplied sequence (MSB-first bit string indexed like the input): repack per codeword, element k in bit k.
    std::vector<uint32_t> words(plan.seq_words + 2 * pdc::PRG_WORDS_PER_THREAD, 0u);
    for (const pdc::UlschCodeword& cw : plan.cws) {
      for (uint32_t i = 0; i != cw.n_in; ++i) {
        const size_t b = (size_t)cw.in_off + i;
        if ((seq_bits_packed[b >> 3] >> (7u - (b & 7u))) & 1u) {
          words[cw.seq_word_off + (i >> 5)] |= 1u << (i & 31u);
        }
      }
    }
    PDC_CUDA(cudaMemcpyAsync(fe.d_seq, words.data(), words.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    PDC_CUDA(cudaStreamSynchronize(s)); // `words` goes out of scope
  }
  fe.k_n_cw          = n_cw;
  fe.k_max_in        = max_in;
  fe.k_max_sch       = max_sch;
  fe.k_max_uci       = max_uci;
  fe.k_all_deferred  = all_deferred;
  fe.k_generate_seq  = seq_bits_packed == nullptr;
  fe.kernels_pending = true;
  fe.uci_bytes     = uci_end;
  fe.sch_len       = (uint32_t)sch_end;
  fe.deferred      = any_deferred;
  fe.deferred_n_cw = n_cw;
  fe.deferred_in   = d_in;
  fe.deferred_sch  = d_sch;
  if (launch_now) {
    return front_end_kernels(ctx, fe, s);
  }
  return PDC_OK;
}

int pdc_submit_codewords(pdc_ctx*           ctx,
                         uint32_t           stream,
                         const pdc_cw_desc* cws,
                         uint32_t           n_cw,
                         const int8_t*      raw_llrs,
                         size_t             n_raw,
                         int8_t*            uci_out,
                         size_t             uci_capacity,
                         pdc_cw_result*     results)
{
  if (!ctx || stream >= ctx->queues.size() || !cws || n_cw == 0 || !raw_llrs || !results || n_cw > 65535u) {
    return fail(PDC_ERR_INVALID, "pdc_submit_codewords: invalid argument");
  }
  Queue& q = ctx->queues[stream];
  if (q.busy || q.fe.pending) {
    return fail(PDC_ERR_CAPACITY, "pdc_submit_codewords: queue busy (call pdc_wait first)");
  }
  PDC_CUDA(cudaSetDevice(ctx->cfg.device));
  PDC_CUDA(grow_device(&q.fe.d_raw, &q.fe.raw_cap, n_raw + 16));
  PDC_CUDA(cudaMemcpyAsync(q.fe.d_raw, raw_llrs, n_raw, cudaMemcpyHostToDevice, q.stream));
  // The kernels are queued by the pdc_submit that follows (after its descriptor uploads) or, failing that, by pdc_wait:
  // a kernel between two host-to-device copies of one stream would hold back the copies of the next slot.
  int rc = front_end_launch(ctx, q.fe, cws, n_cw, n_raw, q.fe.d_raw, q.d_llrs, ctx->cfg.max_llrs,
                            uci_out ? uci_capacity : (size_t)-1, nullptr, results, q.stream, nullptr, false);
  if (rc != PDC_OK) {
    return rc;
  }
  q.fe.u_uci        = nullptr;
  q.fe.uci_copy_dst = nullptr;
  if (uci_out && q.fe.uci_bytes != 0) {
    const bool direct = is_pinned_host(uci_out);
    q.fe.uci_copy_dst = direct ? uci_out : q.fe.h_uci;
    q.fe.u_uci        = direct ? nullptr : uci_out;
  }
  q.fe.pending = true;
  return PDC_OK;
}

// ---- soft demapper -----------------------------------------------------------------------------------------------------

static int demod_kernel_launch(pdc_ctx* ctx, DemodStage& dm, cudaStream_t s)
{
  dm.kernel_pending = false;
  // Up to DEMOD_MAX_CALLS calls per launch (a 16-cell slot has 16 x 13).
  static thread_local pdc::DemodArgs args;
  args.tables     = ctx->d_demod_tables;
  args.symbols    = dm.k_sym;
  args.noise_vars = dm.k_nv;
  args.llrs       = dm.k_llrs;
  args.flags      = (ctx->cfg.demod_mode == PDC_DEMOD_SCALAR) ? pdc::DEMOD_SCALAR_ONLY : 0u;
  for (size_t c0 = 0; c0 < dm.calls.size(); c0 += pdc::DEMOD_MAX_CALLS) {
    const uint32_t n = (uint32_t)std::min<size_t>(pdc::DEMOD_MAX_CALLS, dm.calls.size() - c0);
    uint32_t       tiles = 0;
    for (uint32_t c = 0; c != n; ++c) {
      args.calls[c]      = dm.calls[c0 + c];
      args.tile_start[c] = tiles;
      tiles += (args.calls[c].n_sym + pdc::DEMOD_TILE - 1) / pdc::DEMOD_TILE;
    }
    args.tile_start[n] = tiles;
    args.n_calls       = n;
    if (tiles == 0) {
      continue;
    }
    pdc::demod_kernel<<<tiles, pdc::DEMOD_THREADS, 0, s>>>(args);
    PDC_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  return PDC_OK;
}

// Validates dm.calls and prepares (or queues) the kernel: d_sym / d_nv (n_sym symbols, on the device) -> d_llrs.
static int demod_prepare(pdc_ctx*      ctx,
                         DemodStage&   dm,
                         const float2* d_sym,
                         const float*  d_nv,
                         size_t        n_sym,
                         int8_t*       d_llrs,
                         size_t        llr_capacity,
                         cudaStream_t  s,
                         bool          launch_now)
{
  for (const pdc::DemodCall& call : dm.calls) {
    const uint32_t m = call.mod;
    if (!(m == PDC_MOD_PI_2_BPSK || m == PDC_MOD_BPSK || m == PDC_MOD_QPSK || m == PDC_MOD_QAM16 ||
          m == PDC_MOD_QAM64 || m == PDC_MOD_QAM256)) {
      return fail(PDC_ERR_INVALID, "soft demapper: invalid modulation");
    }
    const size_t qm = (m == PDC_MOD_PI_2_BPSK) ? 1 : m;
    if ((size_t)call.sym_off + call.n_sym > n_sym || (size_t)call.llr_off + (size_t)call.n_sym * qm > llr_capacity) {
      return fail(PDC_ERR_INVALID, "soft demapper: call outside the symbol or soft-bit buffer");
    }
  }
  dm.k_sym          = d_sym;
  dm.k_nv           = d_nv;
  dm.k_llrs         = d_llrs;
  dm.kernel_pending = true;
  return launch_now ? demod_kernel_launch(ctx, dm, s) : PDC_OK;
}

// Soft bits of a codeword: the resource elements of its OFDM symbols times the bits per element
// (pusch_demodulator_impl.cpp:143-174). 0 for a description the plan will reject.
static size_t codeword_soft_bits(const pdc_cw_desc& d)
{
  if (d.nof_symbols == 0 || d.start_symbol_index + d.nof_symbols > 14 || (d.dmrs_type != 1 && d.dmrs_type != 2)) {
    return 0;
  }
  const int per_prb_dmrs = d.nof_cdm_groups_without_data * ((d.dmrs_type == 1) ? 6 : 4);
  if (per_prb_dmrs > 12) {
    return 0;
  }
  size_t n_re = 0;
  for (unsigned l = d.start_symbol_index; l != (unsigned)d.start_symbol_index + d.nof_symbols; ++l) {
    n_re += (size_t)d.nof_prb * (((d.dmrs_symbol_mask >> l) & 1u) ? (12 - per_prb_dmrs) : 12);
  }
  return n_re * d.qm * d.nof_layers;
}

int pdc_submit_symbols(pdc_ctx*           ctx,
                       uint32_t           stream,
                       const pdc_cw_desc* cws,
                       uint32_t           n_cw,
                       const uint32_t*    sym_offsets,
                       const float*       symbols,
                       const float*       noise_vars,
                       size_t             n_sym,
                       int8_t*            uci_out,
                       size_t             uci_capacity,
                       pdc_cw_result*     results)
{
  if (!ctx || stream >= ctx->queues.size() || !cws || n_cw == 0 || n_cw > 65535u || !sym_offsets || !symbols ||
      !noise_vars || n_sym == 0 || n_sym > 0xffffffffu || !results) {
    return fail(PDC_ERR_INVALID, "pdc_submit_symbols: invalid argument");
  }
  Queue& q = ctx->queues[stream];
  if (q.busy || q.fe.pending) {
    return fail(PDC_ERR_CAPACITY, "pdc_submit_symbols: queue busy (call pdc_wait first)");
  }
  size_t n_raw = 0;
  for (uint32_t i = 0; i != n_cw; ++i) {
    if (!(cws[i].flags & PDC_CW_SCRAMBLED)) {
      return fail(PDC_ERR_INVALID, "pdc_submit_symbols: the demapper's soft bits are scrambled (PDC_CW_SCRAMBLED)");
    }
    n_raw = std::max(n_raw, (size_t)cws[i].in_offset + codeword_soft_bits(cws[i]));
  }
  if (n_raw > 0xfffffff0u) {
    return fail(PDC_ERR_CAPACITY, "pdc_submit_symbols: too many soft bits");
  }
  FrontEnd& fe = q.fe;
  PDC_CUDA(cudaSetDevice(ctx->cfg.device));
  PDC_CUDA(grow_device(&fe.d_raw, &fe.raw_cap, n_raw + 16));
  PDC_CUDA(grow_device(&fe.dm.d_sym, &fe.dm.sym_cap, n_sym));
  PDC_CUDA(grow_device(&fe.dm.d_nv, &fe.dm.nv_cap, n_sym));
  PDC_CUDA(cudaMemcpyAsync(fe.dm.d_sym, symbols, n_sym * sizeof(float2), cudaMemcpyHostToDevice, q.stream));
  PDC_CUDA(cudaMemcpyAsync(fe.dm.d_nv, noise_vars, n_sym * sizeof(float), cudaMemcpyHostToDevice, q.stream));
  int rc = front_end_launch(ctx, fe, cws, n_cw, n_raw, fe.d_raw, q.d_llrs, ctx->cfg.max_llrs,
                            uci_out ? uci_capacity : (size_t)-1, nullptr, results, q.stream, nullptr, false);
  if (rc != PDC_OK) {
    fe.kernels_pending = false;
    return rc;
  }
  // One demodulate_soft call per OFDM symbol of each codeword (pusch_demodulator_impl.cpp:231-247).
  fe.dm.calls.clear();
  for (uint32_t i = 0; i != n_cw; ++i) {
    const pdc::UlschCodeword& cw = fe.plan.cws[i];
    const uint32_t            qm = cw.qm;
    const uint32_t            mod =
        (qm == 1) ? ((cws[i].flags & PDC_CW_PLAIN_BPSK) ? PDC_MOD_BPSK : PDC_MOD_PI_2_BPSK) : qm;
    for (uint32_t k = 0; k != cw.n_sym; ++k) {
      const pdc::UlschSymbol& sy   = fe.plan.syms[cw.sym_first + k];
      const uint32_t          next = (k + 1 != cw.n_sym) ? fe.plan.syms[cw.sym_first + k + 1].in_off : cw.n_in;
      fe.dm.calls.push_back(
          pdc::DemodCall{sym_offsets[i] + sy.in_off / qm, (next - sy.in_off) / qm, cw.in_off + sy.in_off, mod});
    }
  }
  rc = demod_prepare(ctx, fe.dm, fe.dm.d_sym, fe.dm.d_nv, n_sym, fe.d_raw, n_raw, q.stream, false);
  if (rc != PDC_OK) {
    fe.kernels_pending = false;
    return rc;
  }
  fe.u_uci        = nullptr;
  fe.uci_copy_dst = nullptr;
  if (uci_out && fe.uci_bytes != 0) {
    const bool direct = is_pinned_host(uci_out);
    fe.uci_copy_dst   = direct ? uci_out : fe.h_uci;
    fe.u_uci          = direct ? nullptr : uci_out;
  }
  fe.pending = true;
  return PDC_OK;
}

int pdc_launch_demod_device(pdc_ctx*              ctx,
                            const pdc_demod_call* calls,
                            uint32_t              n_calls,
                            const void*           d_symbols,
                            const void*           d_noise_vars,
                            size_t                n_sym,
                            void*                 d_llrs,
                            size_t                llr_capacity,
                            void*                 cuda_stream)
{
  if (!ctx || !calls || n_calls == 0 || !d_symbols || !d_noise_vars || !d_llrs) {
    return fail(PDC_ERR_INVALID, "pdc_launch_demod_device: invalid argument");
  }
  PDC_CUDA(cudaSetDevice(ctx->cfg.device));
  DemodStage& dm = ctx->fe_sync.dm;
  dm.calls.clear();
  for (uint32_t i = 0; i != n_calls; ++i) {
    dm.calls.push_back(pdc::DemodCall{calls[i].sym_offset, calls[i].n_sym, calls[i].llr_offset, calls[i].modulation});
  }
  return demod_prepare(ctx, dm, static_cast<const float2*>(d_symbols), static_cast<const float*>(d_noise_vars), n_sym,
                       static_cast<int8_t*>(d_llrs), llr_capacity, static_cast<cudaStream_t>(cuda_stream), true);
}

int pdc_demodulate_soft(pdc_ctx*     ctx,
                        int8_t*      llrs,
                        const float* symbols,
                        const float* noise_vars,
                        uint32_t     n,
                        int          modulation)
{
  if (!ctx || !llrs || !symbols || !noise_vars) {
    return fail(PDC_ERR_INVALID, "pdc_demodulate_soft: invalid argument");
  }
  if (n == 0) {
    return PDC_OK;
  }
  std::lock_guard<std::mutex> sync_lock(ctx->sync_mutex);
  PDC_CUDA(cudaSetDevice(ctx->cfg.device));
  FrontEnd&    fe = ctx->fe_sync;
  const size_t qm = (modulation == PDC_MOD_PI_2_BPSK) ? 1 : (size_t)modulation;
  PDC_CUDA(grow_device(&fe.d_raw, &fe.raw_cap, (size_t)n * 8 + 16));
  PDC_CUDA(grow_device(&fe.dm.d_sym, &fe.dm.sym_cap, (size_t)n));
  PDC_CUDA(grow_device(&fe.dm.d_nv, &fe.dm.nv_cap, (size_t)n));
  PDC_CUDA(cudaMemcpyAsync(fe.dm.d_sym, symbols, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, nullptr));
  PDC_CUDA(cudaMemcpyAsync(fe.dm.d_nv, noise_vars, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, nullptr));
  fe.dm.calls.assign(1, pdc::DemodCall{0u, n, 0u, (uint32_t)modulation});
  int rc = demod_prepare(ctx, fe.dm, fe.dm.d_sym, fe.dm.d_nv, n, fe.d_raw, (size_t)n * 8, nullptr, true);
  if (rc != PDC_OK) {
    cudaStreamSynchronize(nullptr);
    return rc;
  }
  PDC_CUDA(cudaMemcpyAsync(llrs, fe.d_raw, (size_t)n * qm, cudaMemcpyDeviceToHost, nullptr));
  PDC_CUDA(cudaStreamSynchronize(nullptr));
  return PDC_OK;
}

int pdc_wait(pdc_ctx* ctx, uint32_t stream)
{
  if (!ctx || stream >= ctx->queues.size()) {
    return fail(PDC_ERR_INVALID, "pdc_wait: invalid argument");
  }
  std::lock_guard<std::mutex> queue_lock(ctx->queue_mutex[stream]);
  Queue& q = ctx->queues[stream];
  if (!q.busy && !q.fe.pending) {
    return PDC_OK;
  }
  // A front end without a decode batch behind it: its kernels are still to be queued, and there is no event.
  if (q.fe.pending && q.fe.kernels_pending) {
    cudaSetDevice(ctx->cfg.device);
    front_end_kernels(ctx, q.fe, q.stream);
  }
  cudaError_t e = q.busy ? cudaEventSynchronize(q.done) : cudaStreamSynchronize(q.stream);
  if (q.busy && q.tr_groups != 0 && e == cudaSuccess) {
    // PDC_PIPE_TRACE: where the device time of the pipelined batch went (us from the start of the first copy).
    static int printed = 0;
    if (printed++ % 50 == 49) {
      float t = 0;
      fprintf(stderr, "pipe trace: host enqueue %.0f us;", q.tr_host_us);
      for (int k = 0; k != q.tr_groups; ++k) {
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, q.tr_start, q.tr_copy[k]);
        cudaEventElapsedTime(&b, q.tr_start, q.tr_group[k]);
        cudaEventElapsedTime(&c, q.tr_start, q.tr_kern[k]);
        fprintf(stderr, " g%d in %.0f kernels %.0f out %.0f;", k, a * 1e3f, c * 1e3f, b * 1e3f);
      }
      cudaEventElapsedTime(&t, q.tr_start, q.tr_done);
      fprintf(stderr, " batch %.0f us\n", t * 1e3f);
    }
    q.tr_groups = 0;
  }
  if (q.fe.pending) {
    q.fe.pending = false;
    if (e == cudaSuccess && q.fe.u_uci != nullptr) {
      memcpy(q.fe.u_uci, q.fe.h_uci, q.fe.uci_bytes);
    }
    q.fe.u_uci = nullptr;
    if (!q.busy) {
      return (e == cudaSuccess) ? PDC_OK : fail(PDC_ERR_CUDA, "pdc_wait", e);
    }
  }
  q.busy        = false;
  if (e != cudaSuccess) {
    // A failed batch reports CRC failure with the maximum iteration count, like a dropped accelerator operation
    // (hw_accelerator_pusch_dec_acc100_impl.cpp:233-247).
    for (uint32_t i = 0; i != q.n_cb; ++i) {
      q.u_cb_res[i].crc_ok = 0;
      q.u_cb_res[i].status = 2;
    }
    return fail(PDC_ERR_CUDA, "pdc_wait", e);
  }
  memcpy(q.u_cb_res, q.h_cb_res, sizeof(pdc_cb_result) * q.n_cb);
  if (q.u_cb_bits) {
    memcpy(q.u_cb_bits, q.h_cb_bits, (size_t)q.n_cb * PDC_MAX_CB_BYTES);
  }
  if (q.n_tb != 0) {
    memcpy(q.u_tb_res, q.h_tb_res, sizeof(pdc_tb_result) * q.n_tb);
    if (q.u_tb_out) {
      memcpy(q.u_tb_out, q.h_tb_out, q.tb_out_bytes);
    }
  }
  return PDC_OK;
}

int pdc_poll(pdc_ctx* ctx, uint32_t stream, int* done)
{
  if (!ctx || stream >= ctx->queues.size() || !done) {
    return fail(PDC_ERR_INVALID, "pdc_poll: invalid argument");
  }
  Queue& q = ctx->queues[stream];
  *done    = !q.busy || (cudaEventQuery(q.done) == cudaSuccess);
  return PDC_OK;
}

int pdc_launch_device(pdc_ctx*    ctx,
                      const void* d_cbs,
                      uint32_t    n_cb,
                      const void* d_llrs,
                      const void* d_tbs,
                      uint32_t    n_tb,
                      void*       d_cb_results,
                      void*       d_cb_bits,
                      void*       d_tb_results,
                      void*       d_tb_bytes,
                      uint32_t    max_lifting_size,
                      uint32_t    flags_union,
                      int         any_bg1,
                      void*       cuda_stream)
{
  if (!ctx || !d_cbs || n_cb == 0 || !d_cb_results || !d_cb_bits || max_lifting_size < 2 ||
      max_lifting_size > (uint32_t)pdc::MAX_Z) {
    return fail(PDC_ERR_INVALID, "pdc_launch_device: invalid argument");
  }
  BatchShape shape;
  shape.max_Z       = (int)max_lifting_size;
  shape.any_bg1     = any_bg1 != 0;
  shape.any_decode  = (flags_union & PDC_CB_DECODE) != 0;
  shape.any_dematch = (flags_union & PDC_CB_DEMATCH) != 0;
  shape.high_rate   = (flags_union & PDC_LAUNCH_HIGH_RATE) != 0;
  return launch_batch(ctx, static_cast<const pdc_cb_desc*>(d_cbs), n_cb, static_cast<const int8_t*>(d_llrs),
                      static_cast<const pdc_tb_desc*>(d_tbs), n_tb, static_cast<pdc_cb_result*>(d_cb_results),
                      static_cast<uint8_t*>(d_cb_bits), static_cast<pdc_tb_result*>(d_tb_results),
                      static_cast<uint8_t*>(d_tb_bytes), shape, nullptr, 0, static_cast<cudaStream_t>(cuda_stream), nullptr,
                      &ctx->fe_sync);
}

int pdc_harq_read(pdc_ctx* ctx, uint32_t harq_id, int8_t* soft, uint32_t n)
{
  if (!ctx || harq_id >= ctx->cfg.harq_entries || !soft || n > PDC_MAX_CB_SOFT) {
    return fail(PDC_ERR_INVALID, "pdc_harq_read: invalid argument");
  }
  PDC_CUDA(cudaSetDevice(ctx->cfg.device));
  PDC_CUDA(cudaMemcpy(soft, ctx->d_harq + (size_t)harq_id * PDC_MAX_CB_SOFT, n, cudaMemcpyDeviceToHost));
  return PDC_OK;
}

int pdc_harq_write(pdc_ctx* ctx, uint32_t harq_id, const int8_t* soft, uint32_t n)
{
  if (!ctx || harq_id >= ctx->cfg.harq_entries || !soft || n > PDC_MAX_CB_SOFT) {
    return fail(PDC_ERR_INVALID, "pdc_harq_write: invalid argument");
  }
  PDC_CUDA(cudaSetDevice(ctx->cfg.device));
  PDC_CUDA(cudaMemcpy(ctx->d_harq + (size_t)harq_id * PDC_MAX_CB_SOFT, soft, n, cudaMemcpyHostToDevice));
  // Contents changed behind the kernels: position of the last non-zero soft bit unknown.
  PDC_CUDA(cudaMemset(ctx->d_harq_last + (size_t)harq_id * pdc::DM_MAX_PARTS, 0xff, pdc::DM_MAX_PARTS * sizeof(int32_t)));
  return PDC_OK;
}

int pdc_harq_free(pdc_ctx* ctx, uint32_t harq_id)
{
  if (!ctx || harq_id >= ctx->cfg.harq_entries) {
    return fail(PDC_ERR_INVALID, "pdc_harq_free: invalid argument");
  }
  return PDC_OK;
}

void* pdc_harq_device_ptr(pdc_ctx* ctx)
{
  return ctx ? ctx->d_harq : nullptr;
}

int pdc_ldpc_decode(pdc_ctx*      ctx,
                    int           base_graph,
                    int           lifting_size,
                    const int8_t* llrs,
                    uint32_t      n_llrs,
                    uint32_t      nof_filler,
                    int           crc_kind,
                    int           max_iter,
                    uint8_t*      out,
                    int*          iters)
{
  if (!ctx || !llrs || !out || (base_graph != 1 && base_graph != 2) || lifting_size < 2 ||
      lifting_size > pdc::MAX_Z || max_iter < 1 || max_iter > 255 || crc_kind < PDC_CRC_NONE || crc_kind > PDC_CRC24B) {
    return fail(PDC_ERR_INVALID, "pdc_ldpc_decode: invalid argument");
  }
  const uint32_t Z = (uint32_t)lifting_size;
  const uint32_t K = ((base_graph == 1) ? 22u : 10u) * Z;
  const uint32_t N = ((base_graph == 1) ? 66u : 50u) * Z;
  // Contract of ldpc_decoder_impl::decode (ldpc_decoder_impl.cpp:69-83).
  if (n_llrs < K + 2 * Z || n_llrs > N || nof_filler >= K) {
    return fail(PDC_ERR_INVALID, "pdc_ldpc_decode: input length outside [K + 2Z, N]");
  }
  std::lock_guard<std::mutex> sync_lock(ctx->sync_mutex);
  Queue&                      q = ctx->sync_q;
  PDC_CUDA(cudaSetDevice(ctx->cfg.device));
  pdc_cb_desc d;
  memset(&d, 0, sizeof(d));
  d.rm_length    = n_llrs;
  d.harq_id      = ctx->cfg.harq_entries; // scratch entry
  d.lifting_size = (uint16_t)Z;
  d.nof_filler   = (uint16_t)nof_filler;
  d.base_graph   = (uint8_t)base_graph;
  d.qm           = 1;
  d.crc_kind     = (uint8_t)crc_kind;
  d.max_iter     = (uint8_t)max_iter;
  d.flags        = PDC_CB_DECODE | ((crc_kind != PDC_CRC_NONE) ? PDC_CB_EARLY_STOP : 0);
  d.tb_index     = 0xffff;
  q.h_cbs[0]     = d;
  PDC_CUDA(cudaMemcpyAsync(q.d_cbs, q.h_cbs, sizeof(d), cudaMemcpyHostToDevice, q.stream));
  PDC_CUDA(cudaMemcpyAsync(ctx->d_scratch_llr, llrs, n_llrs, cudaMemcpyHostToDevice, q.stream));
  PDC_CUDA(cudaMemsetAsync(q.d_cb_res, 0, sizeof(pdc_cb_result), q.stream));
  BatchShape shape;
  shape.max_Z      = (int)Z;
  shape.any_bg1    = base_graph == 1;
  shape.any_decode = true;
  int rc = launch_batch(ctx, q.d_cbs, 1, nullptr, nullptr, 0, q.d_cb_res, q.d_cb_bits, nullptr, nullptr, shape,
                        ctx->d_scratch_llr, n_llrs, q.stream);
  if (rc != PDC_OK) {
    return rc;
  }
  PDC_CUDA(cudaMemcpyAsync(q.h_cb_res, q.d_cb_res, sizeof(pdc_cb_result), cudaMemcpyDeviceToHost, q.stream));
  PDC_CUDA(cudaMemcpyAsync(q.h_cb_bits, q.d_cb_bits, (K + 7) / 8, cudaMemcpyDeviceToHost, q.stream));
  PDC_CUDA(cudaStreamSynchronize(q.stream));
  pdc_cb_result r = q.h_cb_res[0];
  if (r.status == 2) {
    return fail(PDC_ERR_INVALID, "pdc_ldpc_decode: descriptor rejected by the kernel");
  }
  // All-zero input with a CRC calculator: the reference leaves the output untouched (ldpc_decoder_impl.cpp:88-94).
  if (!(r.status == 1 && crc_kind != PDC_CRC_NONE)) {
    memcpy(out, q.h_cb_bits, (K + 7) / 8);
  }
  if (iters) {
    *iters = (crc_kind != PDC_CRC_NONE && r.crc_ok) ? r.iters : 0;
  }
  return PDC_OK;
}

int pdc_rate_dematch(pdc_ctx*      ctx,
                     int8_t*       buffer,
                     uint32_t      N,
                     const int8_t* llrs,
                     uint32_t      E,
                     int           new_data,
                     int           rv,
                     int           qm,
                     uint32_t      nref,
                     uint32_t      nof_filler)
{
  if (!ctx || !buffer || !llrs || E == 0 || rv < 0 || rv > 3 ||
      !(qm == 1 || qm == 2 || qm == 4 || qm == 6 || qm == 8) || (E % (uint32_t)qm) != 0) {
    return fail(PDC_ERR_INVALID, "pdc_rate_dematch: invalid argument");
  }
  int bg = (N % 66u == 0) ? 1 : ((N % 50u == 0) ? 2 : 0);
  if (bg == 0 || N > PDC_MAX_CB_SOFT || E > ctx->scratch_llr_bytes) {
    return fail(PDC_ERR_INVALID, "pdc_rate_dematch: invalid buffer length");
  }
  uint32_t Z = N / ((bg == 1) ? 66u : 50u);
  std::lock_guard<std::mutex> sync_lock(ctx->sync_mutex);
  Queue&                      q = ctx->sync_q;
  PDC_CUDA(cudaSetDevice(ctx->cfg.device));
  pdc_cb_desc d;
  memset(&d, 0, sizeof(d));
  d.rm_length    = E;
  d.harq_id      = ctx->cfg.harq_entries;
  d.nref         = nref;
  d.lifting_size = (uint16_t)Z;
  d.nof_filler   = (uint16_t)nof_filler;
  d.base_graph   = (uint8_t)bg;
  d.qm           = (uint8_t)qm;
  d.rv           = (uint8_t)rv;
  d.flags        = PDC_CB_DEMATCH | (new_data ? PDC_CB_NEW_DATA : 0);
  d.tb_index     = 0xffff;
  q.h_cbs[0]     = d;
  int8_t* entry  = ctx->d_harq + (size_t)ctx->cfg.harq_entries * PDC_MAX_CB_SOFT;
  PDC_CUDA(cudaMemcpyAsync(q.d_cbs, q.h_cbs, sizeof(d), cudaMemcpyHostToDevice, q.stream));
  PDC_CUDA(cudaMemcpyAsync(entry, buffer, N, cudaMemcpyHostToDevice, q.stream));
  // (the scratch entry now holds the caller's buffer: whatever was recorded about its previous contents is void)
  PDC_CUDA(cudaMemsetAsync(ctx->d_harq_last + (size_t)ctx->cfg.harq_entries * pdc::DM_MAX_PARTS, 0xff,
                           pdc::DM_MAX_PARTS * sizeof(int32_t), q.stream));
  PDC_CUDA(cudaMemcpyAsync(ctx->d_scratch_llr, llrs, E, cudaMemcpyHostToDevice, q.stream));
  BatchShape shape;
  shape.max_Z       = (int)Z;
  shape.any_bg1     = bg == 1;
  shape.any_dematch = true;
  int rc = launch_batch(ctx, q.d_cbs, 1, ctx->d_scratch_llr, nullptr, 0, q.d_cb_res, q.d_cb_bits, nullptr, nullptr,
                        shape, nullptr, 0, q.stream);
  if (rc != PDC_OK) {
    return rc;
  }
  PDC_CUDA(cudaMemcpyAsync(buffer, entry, N, cudaMemcpyDeviceToHost, q.stream));
  PDC_CUDA(cudaStreamSynchronize(q.stream));
  return PDC_OK;
}

int pdc_measure_int_peak(pdc_ctx* ctx, int mode, double* lane_ops_per_s)
{
  if (!ctx || !lane_ops_per_s || mode < 0 || mode > 1) {
    return fail(PDC_ERR_INVALID, "pdc_measure_int_peak: invalid argument");
  }
  PDC_CUDA(cudaSetDevice(ctx->cfg.device));
  std::lock_guard<std::mutex> sync_lock(ctx->sync_mutex);
  Queue&    q      = ctx->sync_q;
  const int blocks = ctx->sm_count * 8, threads = 256, iters = 2000;
  uint32_t* d_out  = nullptr;
  PDC_CUDA(dev_alloc(&d_out, (size_t)blocks * threads));
  cudaEvent_t e0, e1;
  PDC_CUDA(cudaEventCreate(&e0));
  PDC_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep != 4; ++rep) {
    PDC_CUDA(cudaEventRecord(e0, q.stream));
    if (mode == 0) {
      int_peak_kernel<0><<<blocks, threads, 0, q.stream>>>(d_out, iters);
    } else {
      int_peak_kernel<1><<<blocks, threads, 0, q.stream>>>(d_out, iters);
    }
    PDC_CUDA(cudaEventRecord(e1, q.stream));
    PDC_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    PDC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) {
      best = ms;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  PDC_FREE(d_out);
  // 64 statements per inner iteration, 2 integer instructions each.
  double ops      = (double)blocks * threads * (double)iters * 64.0 * 2.0;
  *lane_ops_per_s = ops / (best * 1e-3);
  return PDC_OK;
}


int pdc_launch_codewords_device(pdc_ctx*           ctx,
                                const pdc_cw_desc* cws,
                                uint32_t           n_cw,
                                const void*        d_raw_llrs,
                                size_t             n_raw,
                                void*              d_sch,
                                size_t             sch_capacity,
                                void*              d_uci,
                                size_t             uci_capacity,
                                pdc_cw_result*     results,
                                void*              cuda_stream)
{
  if (!ctx || !cws || n_cw == 0 || n_cw > 65535u || !d_raw_llrs || !d_sch || !results) {
    return fail(PDC_ERR_INVALID, "pdc_launch_codewords_device: invalid argument");
  }
  PDC_CUDA(cudaSetDevice(ctx->cfg.device));
  return front_end_launch(ctx, ctx->fe_sync, cws, n_cw, n_raw, static_cast<const int8_t*>(d_raw_llrs),
                          static_cast<int8_t*>(d_sch), sch_capacity, d_uci ? uci_capacity : (size_t)-1, nullptr, results,
                          static_cast<cudaStream_t>(cuda_stream), d_uci ? static_cast<int8_t*>(d_uci) : nullptr);
}

int pdc_ulsch_demux(pdc_ctx*           ctx,
                    const pdc_cw_desc* cws,
                    uint32_t           n_cw,
                    const int8_t*      llrs,
                    size_t             n_llrs,
                    const uint8_t*     seq_bits,
                    int8_t*            sch_out,
                    size_t             sch_capacity,
                    int8_t*            uci_out,
                    size_t             uci_capacity,
                    pdc_cw_result*     results)
{
  if (!ctx || !cws || n_cw == 0 || n_cw > 65535u || !llrs || !sch_out || !results) {
    return fail(PDC_ERR_INVALID, "pdc_ulsch_demux: invalid argument");
  }
  std::lock_guard<std::mutex> sync_lock(ctx->sync_mutex);
  PDC_CUDA(cudaSetDevice(ctx->cfg.device));
  FrontEnd& fe = ctx->fe_sync;
  PDC_CUDA(grow_device(&fe.d_raw, &fe.raw_cap, n_llrs + 16));
  PDC_CUDA(grow_device(&ctx->d_sch_sync, &ctx->sch_sync_cap, sch_capacity + 16));
  PDC_CUDA(cudaMemcpyAsync(fe.d_raw, llrs, n_llrs, cudaMemcpyHostToDevice, nullptr));
  int rc = front_end_launch(ctx, fe, cws, n_cw, n_llrs, fe.d_raw, ctx->d_sch_sync, sch_capacity,
                            uci_out ? uci_capacity : (size_t)-1, seq_bits, results, nullptr);
  if (rc != PDC_OK) {
    cudaStreamSynchronize(nullptr);
    return rc;
  }
  // Only the codewords' own ranges are written back: the caller's buffer may hold other data in between.
  for (const pdc::UlschCodeword& cw : fe.plan.cws) {
    if (cw.n_out[0] != 0) {
      PDC_CUDA(cudaMemcpyAsync(sch_out + cw.sch_off, ctx->d_sch_sync + cw.sch_off, cw.n_out[0], cudaMemcpyDeviceToHost,
                               nullptr));
    }
    const uint32_t n_uci = cw.n_out[1] + cw.n_out[2] + cw.n_out[3];
    if (uci_out && n_uci != 0) {
      PDC_CUDA(cudaMemcpyAsync(uci_out + cw.uci_off, fe.d_uci + cw.uci_off, n_uci, cudaMemcpyDeviceToHost, nullptr));
    }
  }
  PDC_CUDA(cudaStreamSynchronize(nullptr));
  return PDC_OK;
}

int pdc_scrambling_sequence(pdc_ctx* ctx, uint32_t c_init, uint32_t offset, uint32_t n, uint8_t* packed)
{
  if (!ctx || !packed || n == 0 || (uint64_t)offset + n + pdc::PRG_NC >= (1ull << 21)) {
    return fail(PDC_ERR_INVALID, "pdc_scrambling_sequence: invalid argument (sequence positions below 2^21 - 1600)");
  }
  std::lock_guard<std::mutex> sync_lock(ctx->sync_mutex);
  PDC_CUDA(cudaSetDevice(ctx->cfg.device));
  FrontEnd&          fe = ctx->fe_sync;
  pdc::UlschCodeword cw = {};
  cw.c_init             = c_init;
  cw.n_in               = n;
  cw.prg_offset         = offset;
  const uint32_t n_words = (n + 31) / 32 + 2 * pdc::PRG_WORDS_PER_THREAD;
  PDC_CUDA(grow_device(&fe.d_seq, &fe.seq_cap, (size_t)n_words));
  PDC_CUDA(grow_device(&fe.d_plan, &fe.plan_cap, sizeof(cw)));
  PDC_CUDA(cudaMemcpyAsync(fe.d_plan, &cw, sizeof(cw), cudaMemcpyHostToDevice, nullptr));
  const uint32_t per_cta = 128 * pdc::PRG_WORDS_PER_THREAD;
  pdc::prg_kernel<pdc::PRG_WORDS_PER_THREAD><<<dim3(((n + 31) / 32 + per_cta - 1) / per_cta, 1), 128>>>(
      reinterpret_cast<const pdc::UlschCodeword*>(fe.d_plan), fe.d_seq);
  PDC_CUDA(cudaGetLastError());
  ctx->launches++;
  std::vector<uint32_t> words(n_words);
  PDC_CUDA(cudaMemcpy(words.data(), fe.d_seq, (size_t)((n + 31) / 32) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  memset(packed, 0, (n + 7) / 8);
  for (uint32_t i = 0; i != n; ++i) {
    if ((words[i >> 5] >> (i & 31u)) & 1u) {
      packed[i >> 3] |= (uint8_t)(0x80u >> (i & 7u));
    }
  }
  return PDC_OK;
}

int pdc_crc(pdc_ctx* ctx, int crc_kind, const uint8_t* packed, uint32_t nbits, uint32_t* checksum)
{
  if (!ctx || !packed || !checksum || crc_kind < PDC_CRC16 || crc_kind > PDC_CRC6 ||
      (nbits + 7) / 8 > ctx->scratch_llr_bytes) {
    return fail(PDC_ERR_INVALID, "pdc_crc: invalid argument");
  }
  std::lock_guard<std::mutex> sync_lock(ctx->sync_mutex);
  Queue&                      q = ctx->sync_q;
  PDC_CUDA(cudaSetDevice(ctx->cfg.device));
  uint8_t*  d_in  = reinterpret_cast<uint8_t*>(ctx->d_scratch_llr);
  uint32_t* d_out = reinterpret_cast<uint32_t*>(q.d_cb_res);
  if (nbits != 0) {
    PDC_CUDA(cudaMemcpyAsync(d_in, packed, (nbits + 7) / 8, cudaMemcpyHostToDevice, q.stream));
  }
  crc_kernel<<<1, 256, 0, q.stream>>>(d_in, nbits, crc_kind, d_out);
  PDC_CUDA(cudaGetLastError());
  ctx->launches++;
  PDC_CUDA(cudaMemcpyAsync(q.h_cb_res, d_out, sizeof(uint32_t), cudaMemcpyDeviceToHost, q.stream));
  PDC_CUDA(cudaStreamSynchronize(q.stream));
  memcpy(checksum, q.h_cb_res, sizeof(uint32_t));
  return PDC_OK;
}

} // extern "C"

// ---- downlink twin: LDPC encoding + rate matching ------------------------------------------------------------------

static int launch_encode(pdc_ctx* ctx, const pdc::EncodeParams& p, int max_Z, bool any_bg1, int mode, cudaStream_t s)
{
  const size_t  smem = pdc::enc_smem_bytes(any_bg1 ? 1 : 2, max_Z);
  const int threads = std::min(pdc::ENC_MAX_THREADS, ((max_Z + 31) / 32) * 32);
  pdc::ldpc_encode_rm_kernel<<<p.n_cb, threads, smem, s>>>(p, mode);
  PDC_CUDA(cudaGetLastError());
  ctx->launches++;
  return PDC_OK;
}

static bool enc_desc_valid(const pdc_enc_desc& d, int mode, size_t msg_bytes, size_t out_capacity)
{
  const int bg = d.base_graph, Z = d.lifting_size;
  if ((bg != 1 && bg != 2) || Z < 2 || Z > pdc::MAX_Z || pdc::host_tables().set_index[Z] == 0xff) {
    return false;
  }
  const int kb = (bg == 1) ? 22 : 10, N = ((bg == 1) ? 66 : 50) * Z, K = kb * Z, K_sys = (kb - 2) * Z;
  if ((size_t)d.msg_offset + (size_t)(K + 7) / 8 > msg_bytes || d.nof_filler >= K || d.rv > 3) {
    return false;
  }
  if (mode == 1) {
    return (size_t)d.out_offset + (size_t)N <= out_capacity;
  }
  const int qm  = d.qm;
  const int Ncb = (d.nref > 0) ? std::min<int>((int)d.nref, N) : N;
  const size_t out_bytes = (d.flags & PDC_ENC_PACKED) ? ((size_t)d.rm_length + 7) / 8 : (size_t)d.rm_length;
  if (!(qm == 1 || qm == 2 || qm == 4 || qm == 6 || qm == 8) || d.rm_length % qm != 0 || (d.flags & ~PDC_ENC_PACKED) != 0 ||
      (size_t)d.out_offset + out_bytes > out_capacity) {
    return false;
  }
  // A circular buffer that ends inside or before the filler bits is not a configuration the standard produces.
  return !(d.nof_filler != 0 && Ncb < K_sys) && Ncb > (int)d.nof_filler;
}

static int encode_sync(pdc_ctx*            ctx,
                       const pdc_enc_desc* cbs,
                       uint32_t            n_cb,
                       const uint8_t*      msgs,
                       size_t              msg_bytes,
                       uint8_t*            out,
                       size_t              out_capacity,
                       int                 mode)
{
  int  max_Z = 0;
  bool any_bg1 = false;
  for (uint32_t i = 0; i != n_cb; ++i) {
    if (!enc_desc_valid(cbs[i], mode, msg_bytes, out_capacity)) {
      return fail(PDC_ERR_INVALID, "pdc_encode: invalid codeblock descriptor");
    }
    max_Z   = std::max<int>(max_Z, cbs[i].lifting_size);
    any_bg1 = any_bg1 || cbs[i].base_graph == 1;
  }
  std::lock_guard<std::mutex> sync_lock(ctx->sync_mutex);
  PDC_CUDA(cudaSetDevice(ctx->cfg.device));
  PDC_CUDA(grow_device(&ctx->d_enc_cbs, &ctx->enc_cbs_cap, sizeof(pdc_enc_desc) * n_cb));
  PDC_CUDA(grow_device(&ctx->d_enc_msgs, &ctx->enc_msgs_cap, msg_bytes + 16));
  PDC_CUDA(grow_device(&ctx->d_enc_out, &ctx->enc_out_cap, out_capacity + 16));
  PDC_CUDA(cudaMemcpyAsync(ctx->d_enc_cbs, cbs, sizeof(pdc_enc_desc) * n_cb, cudaMemcpyHostToDevice, nullptr));
  PDC_CUDA(cudaMemcpyAsync(ctx->d_enc_msgs, msgs, msg_bytes, cudaMemcpyHostToDevice, nullptr));
  pdc::EncodeParams p;
  p.cbs          = reinterpret_cast<const pdc_enc_desc*>(ctx->d_enc_cbs);
  p.n_cb         = n_cb;
  p.msgs         = ctx->d_enc_msgs;
  p.out          = ctx->d_enc_out;
  p.out_capacity = (uint32_t)out_capacity;
  int rc = launch_encode(ctx, p, max_Z, any_bg1, mode, nullptr);
  if (rc != PDC_OK) {
    return rc;
  }
  // Only the codeblocks' own ranges are written back.
  for (uint32_t i = 0; i != n_cb; ++i) {
    const size_t n = (mode == 1) ? (size_t)((cbs[i].base_graph == 1) ? 66 : 50) * cbs[i].lifting_size
                     : (cbs[i].flags & PDC_ENC_PACKED) ? ((size_t)cbs[i].rm_length + 7) / 8
                                                       : (size_t)cbs[i].rm_length;
    if (n != 0) {
      PDC_CUDA(cudaMemcpyAsync(out + cbs[i].out_offset, ctx->d_enc_out + cbs[i].out_offset, n, cudaMemcpyDeviceToHost,
                               nullptr));
    }
  }
  PDC_CUDA(cudaStreamSynchronize(nullptr));
  return PDC_OK;
}

int pdc_encode(pdc_ctx*            ctx,
               const pdc_enc_desc* cbs,
               uint32_t            n_cb,
               const uint8_t*      msgs,
               size_t              msg_bytes,
               uint8_t*            out,
               size_t              out_capacity)
{
  if (!ctx || !cbs || n_cb == 0 || !msgs || !out || out_capacity > 0xfffffff0u) {
    return fail(PDC_ERR_INVALID, "pdc_encode: invalid argument");
  }
  return encode_sync(ctx, cbs, n_cb, msgs, msg_bytes, out, out_capacity, 0);
}

int pdc_ldpc_encode(pdc_ctx* ctx, int base_graph, int lifting_size, const uint8_t* msg_packed, uint8_t* codeblock_bits)
{
  if (!ctx || !msg_packed || !codeblock_bits || (base_graph != 1 && base_graph != 2) || lifting_size < 2 ||
      lifting_size > pdc::MAX_Z) {
    return fail(PDC_ERR_INVALID, "pdc_ldpc_encode: invalid argument");
  }
  pdc_enc_desc d = {};
  d.base_graph   = (uint8_t)base_graph;
  d.lifting_size = (uint16_t)lifting_size;
  d.qm           = 1;
  const size_t K = (size_t)((base_graph == 1) ? 22 : 10) * lifting_size, N = (size_t)((base_graph == 1) ? 66 : 50) * lifting_size;
  return encode_sync(ctx, &d, 1, msg_packed, (K + 7) / 8, codeblock_bits, N, 1);
}

int pdc_launch_encode_device(pdc_ctx*    ctx,
                             const void* d_cbs,
                             uint32_t    n_cb,
                             const void* d_msgs,
                             void*       d_out,
                             size_t      out_capacity,
                             uint32_t    max_lifting_size,
                             int         any_bg1,
                             void*       cuda_stream)
{
  if (!ctx || !d_cbs || n_cb == 0 || !d_msgs || !d_out || max_lifting_size < 2 || max_lifting_size > (uint32_t)pdc::MAX_Z) {
    return fail(PDC_ERR_INVALID, "pdc_launch_encode_device: invalid argument");
  }
  PDC_CUDA(cudaSetDevice(ctx->cfg.device));
  pdc::EncodeParams p;
  p.cbs          = static_cast<const pdc_enc_desc*>(d_cbs);
  p.n_cb         = n_cb;
  p.msgs         = static_cast<const uint8_t*>(d_msgs);
  p.out          = static_cast<uint8_t*>(d_out);
  p.out_capacity = (uint32_t)out_capacity;
  return launch_encode(ctx, p, (int)max_lifting_size, any_bg1 != 0, 0, static_cast<cudaStream_t>(cuda_stream));
}

#ifdef H2_PHASE_TIMING
// Timing build only (tools/phase_probe.py): the global-timer stamps CTA 0 of the last decode launch left at its phase
// boundaries.
extern "C" int pdc_debug_read_phases(unsigned long long* out16)
{
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out16, pdc::h2::g_h2_phase, sizeof(unsigned long long) * 16) == cudaSuccess ? PDC_OK
                                                                                                          : PDC_ERR_CUDA;
}
#endif
