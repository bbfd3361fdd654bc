// ccp_project.cu — the persistent lane-refill projection kernel (the hot path) and its launchers.
// Compiled once per (K arms, PANDA link-code mode) pair: -DCCP_TU_K=2|3 -DCCP_TU_PANDA=0|1|2, so the six
// translation units build in parallel.  Each exports ccp_launch_project_K<k>_P<p>().
//
// Replaces KinematicChainConstraint::project (ConstraintFunction.h:57-82) for a whole batch.
#include <stdlib.h>

#include <atomic>
#include <type_traits>

#include "ccp_device.cuh"
#include "ccp_internal.h"

// (sin, cos) of the 7K joints of this thread's sample, in shared memory: [slot][thread] so that a warp's
// access to one slot is 32 consecutive doubles (conflict-free), and ~56 registers are freed.
template <int K, int BLOCK>
struct ccp_sc_smem {
  double* base;  // &smem[threadIdx.x]
  __device__ __forceinline__ double& at(int a, int i, int k) const { return base[((a * CCPC_DOF + i) * 4 + k) * BLOCK]; }
  __device__ __forceinline__ double& s(int a, int i) { return at(a, i, 0); }
  __device__ __forceinline__ double& c(int a, int i) { return at(a, i, 1); }
  __device__ __forceinline__ double& rx(int a, int i) { return at(a, i, 2); }
  __device__ __forceinline__ double& ry(int a, int i) { return at(a, i, 3); }
  __device__ __forceinline__ double s(int a, int i) const { return at(a, i, 0); }
  __device__ __forceinline__ double c(int a, int i) const { return at(a, i, 1); }
  __device__ __forceinline__ double rx(int a, int i) const { return at(a, i, 2); }
  __device__ __forceinline__ double ry(int a, int i) const { return at(a, i, 3); }
};

// hybrid: (sin, cos) stay in registers, only the lever arms (rx, ry) — written once by the forward pass, read once by
// the gradient pass — go to shared memory ([slot][thread], conflict-free): 28 doubles less per thread for K = 2
template <int K, int BLOCK>
struct ccp_sc_hybrid {
  double v[K][CCPC_DOF][2];
  double* base;  // &smem[threadIdx.x]
  __device__ __forceinline__ double& s(int a, int i) { return v[a][i][0]; }
  __device__ __forceinline__ double& c(int a, int i) { return v[a][i][1]; }
  __device__ __forceinline__ double& rx(int a, int i) { return base[((a * CCPC_DOF + i) * 2 + 0) * BLOCK]; }
  __device__ __forceinline__ double& ry(int a, int i) { return base[((a * CCPC_DOF + i) * 2 + 1) * BLOCK]; }
  __device__ __forceinline__ double s(int a, int i) const { return v[a][i][0]; }
  __device__ __forceinline__ double c(int a, int i) const { return v[a][i][1]; }
  __device__ __forceinline__ double rx(int a, int i) const { return base[((a * CCPC_DOF + i) * 2 + 0) * BLOCK]; }
  __device__ __forceinline__ double ry(int a, int i) const { return base[((a * CCPC_DOF + i) * 2 + 1) * BLOCK]; }
};

// the Jacobian rows (28 (K-1) doubles) and the state x (7K doubles) can live there too
template <int K, int BLOCK>
struct ccp_jac_smem {
  double* base;
  __device__ __forceinline__ double& a(int p, int r, int i) { return base[(((p * 2 + r) * 2 + 0) * CCPC_DOF + i) * BLOCK]; }
  __device__ __forceinline__ double& z(int p, int r, int i) { return base[(((p * 2 + r) * 2 + 1) * CCPC_DOF + i) * BLOCK]; }
  __device__ __forceinline__ double a(int p, int r, int i) const { return base[(((p * 2 + r) * 2 + 0) * CCPC_DOF + i) * BLOCK]; }
  __device__ __forceinline__ double z(int p, int r, int i) const { return base[(((p * 2 + r) * 2 + 1) * CCPC_DOF + i) * BLOCK]; }
};
template <int BLOCK>
struct ccp_x_smem {
  double* base;
  __device__ __forceinline__ double& operator[](int j) { return base[j * BLOCK]; }
  __device__ __forceinline__ double operator[](int j) const { return base[j * BLOCK]; }
};
// what goes to shared memory: bit 0 = sincos, bit 1 = Jacobian rows, bit 2 = state x
#define CCP_SM_SC 1
#define CCP_SM_J 2
#define CCP_SM_X 4
#define CCP_SM_NOSTAGE 8   // no seed staging buffers (lanes load their seeds directly)
#define CCP_SM_RXY 16      // lever arms in shared memory, (sin, cos) in registers (ccp_sc_hybrid)
// doubles in front of the seed staging area: max(staging arrays, tail exchange) + state x
template <int K, int BLOCK, int SM>
__host__ __device__ constexpr int ccp_proj_stage_offset() {
  constexpr int stage = ((SM & CCP_SM_SC) ? 4 * CCPC_DOF * K : 0) + ((SM & CCP_SM_J) ? 4 * CCPC_DOF * (K - 1) : 0) +
                        ((SM & CCP_SM_RXY) ? 2 * CCPC_DOF * K : 0);
  constexpr int exch = CCPC_DOF * K + 1;
  return BLOCK * ((stage > exch ? stage : exch) + ((SM & CCP_SM_X) ? CCPC_DOF * K : 0));
}
template <int K, int BLOCK, int SM>
constexpr size_t ccp_proj_smem_bytes() {
  // + per warp two seed buffers of one claim chunk (32 states) each
  return sizeof(double) * ((size_t)ccp_proj_stage_offset<K, BLOCK, SM>() +
                           ((SM & CCP_SM_NOSTAGE) ? 0 : (size_t)(BLOCK / 32) * 2 * 32 * CCPC_DOF * K));
}

// ------------------------------------------------------------------------------------------
// work distribution
// ------------------------------------------------------------------------------------------
// The work items of a launch are numbered u = 0 .. total-1: first the samples adopted from the previous
// (pipelined) launch, then this launch's own seeds.  A warp owns a private CHUNK of work numbers (shared
// memory: [next, end)).  Lanes whose sample just finished take the next numbers from it; only when it runs dry
// does the warp's leader touch the global work counter (one atomic per CCP_CLAIM_CHUNK samples instead of one
// per refill) and start the bulk copy of the new chunk's seeds into shared memory.
#define CCP_MAX_DEVICES 64
#define CCP_CLAIM_CHUNK 32u

// how many trips pass between two tail rendezvous of the block
#define CCP_TAIL_PERIOD 4

struct ccp_work {
  unsigned total;          // adopted + own
  unsigned n_adopt;
  unsigned first_dynamic;  // work numbers below this are handed out statically
};

// Per-warp chunk state (shared memory).  The seeds of a chunk are STAGED in shared memory by one bulk-copy (TMA,
// cp.async.bulk + mbarrier) issued by the warp's leader the moment the chunk is claimed, so a refill reads its seed
// with a shared-memory load instead of waiting ~1 us on HBM/L2 with the other 31 lanes of the warp stalled behind
// it.  Two buffers alternate (chunk q uses buffer q & 1): lanes that took the last numbers of the old chunk still
// read it while the copy of the new one is in flight.
struct ccp_warp_chunk {
  unsigned next, end;    // private work numbers [next, end)
  unsigned seq;          // chunks claimed so far
  unsigned base[2];      // first work number of the chunk in buffer b
  unsigned staged[2];    // 1: buffer b holds (or is receiving) that chunk's seeds
  unsigned parity[2];    // mbarrier phase parity that completes buffer b's current copy
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// leader only: start the bulk copy of work numbers [base, end) into buffer b, if they are plain seeds of this launch
// and the copy meets TMA's 16-byte rules
template <int K, bool SOA>
__device__ __forceinline__ void stage_chunk(ccp_warp_chunk* wc, int b, unsigned base, unsigned end, double* buf,
                                            unsigned long long* mbar, const ccp_project_args& A, const ccp_work& W) {
  constexpr int n = CCPC_DOF * K;
  wc->base[b] = base;
  wc->staged[b] = 0u;
  if (!A.stage_seeds || buf == nullptr || end <= base || base < W.n_adopt) return;
  const unsigned i0 = base - W.n_adopt, cnt = end - base;
  const unsigned dst = smem_u32(buf), bar = smem_u32(mbar);
  if (!SOA) {
    const unsigned bytes = cnt * n * 8u;
    if (((i0 * (unsigned)(n * 8)) | bytes) & 15u) return;
    const double* src = A.seeds + (size_t)i0 * n;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
  } else {
    if ((i0 | cnt | (unsigned)A.seed_stride) & 1u) return;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cnt * 8u * n) : "memory");
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
      const double* src = A.seeds + (size_t)j * (size_t)A.seed_stride + i0;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       dst + (unsigned)j * CCP_CLAIM_CHUNK * 8u),
                   "l"(src), "r"(cnt * 8u), "r"(bar)
                   : "memory");
    }
  }
  wc->staged[b] = 1u;
  wc->parity[b] ^= 1u;  // the phase this copy completes
}

// Lanes whose sample just finished take the next work numbers.  Returns the number (>= total: none left) and, in
// `buf_sel`, which of the warp's two seed buffers belongs to the chunk the number came from.
// Fast path: one shared-memory atomic on the warp's chunk cursor per lane.  Only the lanes that overflow the chunk
// go on to claim a new one together (leader: global work counter, bulk copy of the seeds, new cursor).
template <int K, bool SOA>
__device__ __forceinline__ unsigned claim_chunked(ccp_warp_chunk* wc, double* stage, unsigned long long* mbar,
                                                   const ccp_project_args& A, const ccp_work& W, volatile int* s_tail,
                                                   int& buf_sel) {
  constexpr int n = CCPC_DOF * K;
  // every lane of the event reads the chunk's end and sequence number BEFORE any of them may replace the chunk
  const unsigned u0 = atomicAdd(&wc->next, 1u);
  const unsigned end0 = *(volatile unsigned*)&wc->end;
  const unsigned seq0 = *(volatile unsigned*)&wc->seq;
  if (u0 < end0) {
    buf_sel = (int)(seq0 & 1u);
    return u0;
  }
  const unsigned mask = __activemask();  // the lanes that found the chunk empty
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(mask) - 1;
  const unsigned need = __popc(mask);
  const unsigned rank = __popc(mask & ((1u << lane) - 1u));
  unsigned base1 = W.total;
  if (lane == leader) {
    unsigned end1 = W.total;
    if (!*s_tail) {
      base1 = W.first_dynamic + atomicAdd((unsigned int*)A.counter, CCP_CLAIM_CHUNK);
      if (base1 > W.total) base1 = W.total;
      end1 = (W.total - base1 < CCP_CLAIM_CHUNK) ? W.total : base1 + CCP_CLAIM_CHUNK;
      if (end1 - base1 < CCP_CLAIM_CHUNK) *s_tail = 1;  // the work has run dry: the block enters its tail
    }
    const unsigned seq = seq0 + 1u;
    const int b = (int)(seq & 1u);
    stage_chunk<K, SOA>(wc, b, base1, end1, stage + b * (int)CCP_CLAIM_CHUNK * n, mbar + b, A, W);
    wc->end = end1;
    wc->seq = seq;
    wc->next = (end1 - base1 < need) ? end1 : base1 + need;  // need <= 32 = CCP_CLAIM_CHUNK
  }
  base1 = __shfl_sync(mask, base1, leader);
  buf_sel = (int)((seq0 + 1u) & 1u);
  // a number past the chunk's end is >= total: no work for the lane
  return base1 + rank;
}

// Work number u -> the lane's sample: state x, index within its launch, iteration count | launch slot << 16.
template <int K, bool SOA, class XT>
__device__ __forceinline__ void load_sample(const ccp_model& M, const ccp_project_args& A, const ccp_work& W, unsigned u,
                                            const ccp_warp_chunk* wc, const double* stage, unsigned long long* mbar,
                                            int buf_sel, XT& x, unsigned& idx, int& it) {
  constexpr int n = CCPC_DOF * K;
  if (u >= W.total) {
    idx = CCP_NO_SAMPLE;
    return;
  }
  if (u < W.n_adopt) {
    const ccp_park_rec* r = A.adopt + u;
    idx = __ldcg(&r->idx);
    it = __ldcg(&r->it_slot);
#pragma unroll
    for (int j = 0; j < n; ++j) x[j] = __ldcg(&r->x[j]);
    return;
  }
  idx = u - W.n_adopt;
  it = (int)(A.slot << 16);
  if (wc->staged[buf_sel]) {
    // the chunk's seeds are (being) copied to shared memory: wait for the copy's mbarrier phase, then read
    const unsigned bar = smem_u32(mbar + buf_sel), parity = wc->parity[buf_sel];
    unsigned done = 0;
    do {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar), "r"(parity)
          : "memory");
    } while (!done);
    const double* src = stage + buf_sel * (int)CCP_CLAIM_CHUNK * n;
    const unsigned k = u - wc->base[buf_sel];
#pragma unroll
    for (int j = 0; j < n; ++j) x[j] = SOA ? src[j * (int)CCP_CLAIM_CHUNK + k] : src[k * n + j];
    return;
  }
#pragma unroll
  for (int j = 0; j < n; ++j) x[j] = ld_elem<SOA>(A.seeds, idx, j, A.seed_stride, n);
}

// ---- epilogue of one sample (ConstraintFunction.h:75-81): flags, outputs, compaction, completion count ----
// F needs only e2, sv2 and d[.][0] (what ccp_converged and ccp_residual read).
template <int K, bool SOA, class XT>
__device__ __forceinline__ void write_result(const ccp_model& M, const ccp_project_args& A, XT& x, unsigned idx, int it,
                                             const ccp_fwd<K>& F, unsigned* s_done) {
  constexpr int n = CCPC_DOF * K, m = 2 * (K - 1);
  const bool cv = ccp_converged<K>(M, F);
  const bool okk = cv && ccp_joint_valid<K>(M, x);
  // the sample reports into the arrays of the launch it was submitted with
  ccp_out_desc D;
  if (((unsigned)it >> 16) == A.slot) {
    D.x_out = A.x_out; D.ok = A.ok; D.conv = A.conv; D.iters = A.iters; D.resid = A.resid;
    D.count = A.out_stride; D.wrap = A.wrap;
    D.n_ok = A.own_n_ok; D.compact = A.own_compact; D.compact_idx = A.own_compact_idx;
    D.compact_cap = A.own_compact_cap; D.idx_base = A.idx_base;
  } else {
    const ccp_out_desc* T = A.desc_table + ((unsigned)it >> 16);
    D.x_out = T->x_out; D.ok = T->ok; D.conv = T->conv; D.iters = T->iters; D.resid = T->resid;
    D.count = T->count; D.wrap = T->wrap;
    D.n_ok = T->n_ok; D.compact = T->compact; D.compact_idx = T->compact_idx;
    D.compact_cap = T->compact_cap; D.idx_base = T->idx_base;
  }
  if (D.wrap) {  // the sampler's enforceBounds (KinematicChain.h:118-130); one out-of-line copy of the fmod code
#pragma unroll
    for (int j = 0; j < n; ++j) x[j] = ccp_wrap_pi_call(x[j]);
  }
  if (D.x_out) {
#pragma unroll
    for (int j = 0; j < n; ++j) st_elem<SOA>(D.x_out, idx, j, D.count, n, x[j]);
  }
  if (D.ok) D.ok[idx] = okk;
  if (D.conv) D.conv[idx] = cv;
  if (D.iters) D.iters[idx] = it & 0xffff;
  if (A.done) atomicAdd(&s_done[(unsigned)it >> 16], 1u);
  if (D.resid) {
    double fv[m];
    ccp_residual<K>(F, fv, nullptr);
#pragma unroll
    for (int k = 0; k < m; ++k) st_elem<SOA>(D.resid, idx, k, D.count, m, fv[k]);
  }
  if (D.n_ok) {
    // per-batch compaction (host path): the state joins the packed rows of the batch it was submitted with
    if (okk) {
      const unsigned long long slot = atomicAdd(D.n_ok, 1ULL);
      if ((long long)slot < D.compact_cap) {
        if (D.compact) {
#pragma unroll
          for (int j = 0; j < n; ++j) D.compact[slot * n + j] = x[j];
        }
        if (D.compact_idx) D.compact_idx[slot] = (int32_t)(D.idx_base + idx);
      }
    }
    return;
  }
  // the compacted stream is a stream: a state is appended to the buffer of the launch it finished in
  if (A.n_ok && okk) {
    const unsigned long long slot = atomicAdd(A.n_ok, 1ULL);
    if (A.compact) {
#pragma unroll
      for (int j = 0; j < n; ++j) A.compact[slot * n + j] = x[j];
    }
    // fused all-gather: the state goes straight into this rank's rows of every peer's pool (NVLink P2P
    // stores, fire and forget; visible to the peers when this kernel has completed)
    if (A.peer_world > 0 && (long long)slot < A.peer_cap) {
      const long long row_off = (A.peer_row0 + (long long)slot) * n;
      if (A.peer_mc) {
        // ONE store per 16 bytes into the pool's MULTICAST mapping: the NVSwitch fans it out to every rank's pool
        // (n / 2 stores per state whatever the world size, instead of n per peer)
        double* row = A.peer_mc + row_off;
        if constexpr ((n & 1) == 0) {
#pragma unroll
          for (int j = 0; j < n; j += 2)
            asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + j), "r"(__double2loint(x[j])),
                         "r"(__double2hiint(x[j])), "r"(__double2loint(x[j + 1])), "r"(__double2hiint(x[j + 1]))
                         : "memory");
        } else {  // 21 doubles: rows are only 8-byte aligned
#pragma unroll
          for (int j = 0; j < n; ++j) asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(row + j), "d"(x[j]) : "memory");
        }
      } else {
        for (int p = 0; p < A.peer_world; ++p) {
          double* row = A.peer_pool[p] + row_off;
          if constexpr ((n & 1) == 0) {  // 112-byte rows of a 16-byte aligned pool: 16-byte stores
#pragma unroll
            for (int j = 0; j < n; j += 2) *reinterpret_cast<double2*>(row + j) = make_double2(x[j], x[j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < n; ++j) row[j] = x[j];
          }
        }
      }
    }
  }
}

// PANDA: structured stock-Panda link code (ccp_core.h).  SM: which per-sample arrays are staged in
// shared memory (CCP_SM_* bits) instead of registers.
//
// Loop structure.  One trip = one Newton iteration of the lane's current sample (ConstraintFunction.h:68-73).
// A lane whose sample leaves the loop writes its result and takes the next work number (lane refill), so the
// 0..250-iteration spread does not idle the warp while there is work.  When the work runs dry the block enters
// its TAIL:
//   * complete mode (A.park == nullptr): every CCP_TAIL_PERIOD trips the warps meet at a barrier, and as soon as
//     the samples still iterating fit in fewer warps they are packed into the lowest warps through shared memory
//     ((x, it, index) is the whole state of a sample between trips).  Emptied warps wait at the barrier and issue
//     nothing, so the last samples run at one-warp-per-scheduler latency instead of sharing the FP64 pipe with
//     warps that carry one or two live lanes each.
//   * pipelined mode (A.park != nullptr): the warp writes its live samples to the park buffer and exits; the next
//     launch adopts them as its first work items.  No lane ever idles on a straggler.
// Which lane (or launch) ran a sample never affects its result.
template <int K, int PANDA, bool SOA, int BLOCK, int MINB, int SM>
__global__ void __launch_bounds__(BLOCK, MINB)
ccp_project_kernel(const __grid_constant__ ccp_model M, const __grid_constant__ ccp_project_args A) {
  constexpr int n = CCPC_DOF * K, m = 2 * (K - 1);
  constexpr int NW = BLOCK / 32;
  extern __shared__ __align__(16) double ccp_smem[];
  __shared__ int s_tail;
  __shared__ unsigned s_exited;              // pipelined mode: warps that have left
  __shared__ unsigned s_done[CCP_NUM_DESC];  // samples this block finished since its last publication, by launch slot
  __shared__ int s_wcnt[NW];
  __shared__ ccp_warp_chunk s_chunk[NW];
  __shared__ __align__(8) unsigned long long s_mbar[NW][2];
  double* sm_next = ccp_smem + threadIdx.x;
  typename std::conditional<(SM & CCP_SM_SC) != 0, ccp_sc_smem<K, BLOCK>,
                            typename std::conditional<(SM & CCP_SM_RXY) != 0, ccp_sc_hybrid<K, BLOCK>, ccp_sc_local<K>>::type>::type S;
  if constexpr ((SM & CCP_SM_SC) != 0) {
    S.base = sm_next;
    sm_next += 4 * CCPC_DOF * K * BLOCK;
  } else if constexpr ((SM & CCP_SM_RXY) != 0) {
    S.base = sm_next;
    sm_next += 2 * CCPC_DOF * K * BLOCK;
  }
  typename std::conditional<(SM & CCP_SM_J) != 0, ccp_jac_smem<K, BLOCK>, ccp_jac<K>>::type J;
  if constexpr ((SM & CCP_SM_J) != 0) {
    J.base = sm_next;
    sm_next += 4 * CCPC_DOF * (K - 1) * BLOCK;
  }
  typename std::conditional<(SM & CCP_SM_X) != 0, ccp_x_smem<BLOCK>, double[n]>::type x;
  if constexpr ((SM & CCP_SM_X) != 0) {
    constexpr int stage = ((SM & CCP_SM_SC) ? 4 * CCPC_DOF * K : 0) + ((SM & CCP_SM_J) ? 4 * CCPC_DOF * (K - 1) : 0);
    constexpr int exch = CCPC_DOF * K + 1;
    x.base = ccp_smem + threadIdx.x + (stage > exch ? stage : exch) * BLOCK;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  ccp_work W;
  W.n_adopt = A.adopt ? __ldcg(A.adopt_count) : 0u;
  W.total = W.n_adopt + (unsigned)A.count;
  // The first chunk of every warp is static and interleaved over the blocks (chunk c -> block c % grid, warp
  // c / grid), so a batch smaller than the machine spreads one warp per SM before any SM gets a second one; the
  // global counter hands out what lies beyond those gridDim.x * NW chunks.
  const unsigned static_chunks = gridDim.x * NW;
  W.first_dynamic = (W.total / CCP_CLAIM_CHUNK < static_chunks) ? W.total : static_chunks * CCP_CLAIM_CHUNK;
  if (threadIdx.x < CCP_NUM_DESC) s_done[threadIdx.x] = 0u;
  if (threadIdx.x == 0) {
    s_exited = 0u;
    s_tail = (W.first_dynamic >= W.total) ? 1 : 0;
    if (blockIdx.x == 0 && A.park) {  // successors find this launch's output arrays by slot
      ccp_out_desc d;
      d.x_out = A.x_out; d.ok = A.ok; d.conv = A.conv; d.iters = A.iters; d.resid = A.resid;
      d.count = A.out_stride; d.wrap = A.wrap; d.idx_base = A.idx_base;
      d.compact = A.own_compact; d.n_ok = A.own_n_ok; d.compact_idx = A.own_compact_idx; d.compact_cap = A.own_compact_cap;
      A.desc_table[A.slot] = d;
    }
  }
  // seed staging area: [NW][2][CCP_CLAIM_CHUNK * n] doubles behind the exchange / staging arrays
  ccp_warp_chunk* wc = &s_chunk[warp];
  unsigned long long* mbar = s_mbar[warp];
  double* stage = ccp_smem + ccp_proj_stage_offset<K, BLOCK, SM>() + warp * (2 * (int)CCP_CLAIM_CHUNK * n);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar + 1)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    wc->parity[0] = wc->parity[1] = 1u;  // stage_chunk flips it: the first copy into a buffer completes phase 0
    wc->staged[0] = wc->staged[1] = 0u;
  }
  __syncthreads();
  if (lane == 0) {
    const unsigned long long b0 = (unsigned long long)(warp * gridDim.x + blockIdx.x) * CCP_CLAIM_CHUNK;
    const unsigned lo = b0 < W.total ? (unsigned)b0 : W.total;
    const unsigned hi = (W.total - lo < CCP_CLAIM_CHUNK) ? W.total : lo + CCP_CLAIM_CHUNK;
    wc->next = lo;
    wc->end = hi;
    wc->seq = 0u;
    stage_chunk<K, SOA>(wc, 0, lo, hi, stage, mbar, A, W);
  }
  __syncthreads();
  int it = 0;
  unsigned idx = CCP_NO_SAMPLE;
  {
    int bsel;
    const unsigned u = claim_chunked<K, SOA>(wc, stage, mbar, A, W, &s_tail, bsel);
    load_sample<K, SOA>(M, A, W, u, wc, stage, mbar, bsel, x, idx, it);
  }
  bool tail = false;
  int since = 0;
  for (;;) {
    // Every lane of the warp meets here (lanes never leave the loop on their own).  `tail` must be the same for all of
    // them — the tail handling below is warp-collective — so it is agreed on by a vote, not read lane by lane: a lane
    // coming out of the epilogue branch may itself have just set s_tail.
    __syncwarp();
    if (!tail) tail = __any_sync(0xffffffffu, *(volatile int*)&s_tail != 0);
    if (tail) {
      if (A.park) {
        // ---- pipelined mode: once few enough of the warp's lanes still carry a sample, park what the warp
        // holds (live samples, then the unstarted rest of its private chunk) and leave.  A launch that outgrew
        // its static allotment parks the moment its work runs dry (every other warp is still full, nothing is
        // gained by waiting); a launch smaller than the machine keeps iterating until half the lanes are free,
        // so it makes progress and the parked set stays bounded (<= 16 per warp) however many follow.
        if (idx == CCP_NO_SAMPLE) {
          int bsel;
          const unsigned u = claim_chunked<K, SOA>(wc, stage, mbar, A, W, &s_tail, bsel);
          load_sample<K, SOA>(M, A, W, u, wc, stage, mbar, bsel, x, idx, it);
        }
        __syncwarp();
        const int park_at = (W.first_dynamic < W.total) ? 32 : 16;
        if (__popc(__ballot_sync(0xffffffffu, idx != CCP_NO_SAMPLE)) <= park_at) {
          // A sample carried through A.max_age launches is not parked again: its lane finishes it here (the warp
          // stays).  So launch c + max_age completes launch c's outputs (what the host path's D2H copies wait for),
          // and a launch's slot in the 64-entry descriptor ring is never recycled under one of its samples.
          for (;;) {
            const bool live = idx != CCP_NO_SAMPLE;
            const bool old = live && ((A.slot - ((unsigned)it >> 16)) & (CCP_NUM_DESC - 1u)) >= A.max_age;
            const unsigned bal = __ballot_sync(0xffffffffu, live && !old);
            if (bal != 0u) {
              unsigned base = 0;
              if (lane == 0) base = atomicAdd(A.park_count, (unsigned)__popc(bal));
              base = __shfl_sync(0xffffffffu, base, 0);
              if (live && !old) {
                ccp_park_rec* r = A.park + base + __popc(bal & ((1u << lane) - 1u));
                r->idx = idx;
                r->it_slot = it;
#pragma unroll
                for (int j = 0; j < n; ++j) r->x[j] = x[j];
                idx = CCP_NO_SAMPLE;
              }
            }
            if (idx == CCP_NO_SAMPLE) {  // the unstarted rest of the private chunk
              int bsel;
              const unsigned u = claim_chunked<K, SOA>(wc, stage, mbar, A, W, &s_tail, bsel);
              load_sample<K, SOA>(M, A, W, u, wc, stage, mbar, bsel, x, idx, it);
            }
            __syncwarp();
            const bool young = idx != CCP_NO_SAMPLE && ((A.slot - ((unsigned)it >> 16)) & (CCP_NUM_DESC - 1u)) < A.max_age;
            if (__ballot_sync(0xffffffffu, young) == 0u) break;
          }
          if (__ballot_sync(0xffffffffu, idx != CCP_NO_SAMPLE) == 0u) {
            if (A.done) {
              // The last warp out publishes how many samples of each launch this block finished.  Fence, then
              // count: whoever sees the count sees the results (the D2H stream waits on A.done[slot]).
              __threadfence();
              unsigned gone = 0;
              if (lane == 0) gone = atomicAdd(&s_exited, 1u);
              gone = __shfl_sync(0xffffffffu, gone, 0);
              if (gone == NW - 1) {
                __threadfence();
                for (int sl = lane; sl < (int)CCP_NUM_DESC; sl += 32) {
                  const unsigned c = *(volatile unsigned*)&s_done[sl];
                  if (c) atomicAdd(A.done + sl, c);
                }
              }
            }
            return;
          }
        }
      } else if (since == 0) {
        // ---- complete mode: tail rendezvous ----
        if (idx == CCP_NO_SAMPLE) {  // what is left of the warp's private chunk
          int bsel;
          const unsigned u = claim_chunked<K, SOA>(wc, stage, mbar, A, W, &s_tail, bsel);
          load_sample<K, SOA>(M, A, W, u, wc, stage, mbar, bsel, x, idx, it);
        }
        __syncwarp();
        const bool active = idx != CCP_NO_SAMPLE;
        if (A.done) __threadfence();  // the results of the samples counted in s_done, before the count goes out
        const int total = __syncthreads_count(active);
        unsigned fin = 0;
        if (A.done && threadIdx.x < CCP_NUM_DESC) {  // no thread is inside an epilogue between these two barriers
          fin = s_done[threadIdx.x];
          s_done[threadIdx.x] = 0u;
        }
        if (fin) {
          __threadfence();  // cumulative: everything the barrier made visible to this thread precedes the count
          atomicAdd(A.done + threadIdx.x, fin);
        }
        if (total == 0) break;
        const unsigned bal = __ballot_sync(0xffffffffu, active);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, nonempty = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          const int c = s_wcnt[w];
          nonempty += c > 0;
          before += (w < warp) ? c : 0;
        }
        if (nonempty > (total + 31) / 32) {
          // pack the live samples into the lowest warps
          double* ex = ccp_smem;  // [n + 1][BLOCK]; aliases the staging arrays, which are dead between trips
          if (active) {
            const int slot = before + __popc(bal & ((1u << lane) - 1u));
#pragma unroll
            for (int j = 0; j < n; ++j) ex[j * BLOCK + slot] = x[j];
            ex[n * BLOCK + slot] = __hiloint2double(it, (int)idx);
          }
          __syncthreads();
          if ((int)threadIdx.x < total) {
#pragma unroll
            for (int j = 0; j < n; ++j) x[j] = ex[j * BLOCK + threadIdx.x];
            const double pk = ex[n * BLOCK + threadIdx.x];
            it = __double2hiint(pk);
            idx = (unsigned)__double2loint(pk);
          } else {
            idx = CCP_NO_SAMPLE;
          }
          if constexpr (SM != 0) __syncthreads();  // the staging arrays are about to be written again
        }
        since = CCP_TAIL_PERIOD;
      }
      --since;
    }
    if (idx != CCP_NO_SAMPLE) {
      ccp_fwd<K> F;
      ccp_forward<K, PANDA>(M, x, S, F);
      const bool cont = ccp_needs_step<K>(M, F) && (it & 0xffff) < M.max_iter;
      if (cont) {
        ++it;
        ccp_jacobian<K, PANDA>(M, S, F, J);
        ccp_newton_step<K>(M, F, J, x);
        if (M.clamp) ccp_clamp_to_limits<K>(M, x);
      } else {
        // ---- the sample is finished (ConstraintFunction.h:75-81): results out, then refill the lane ----
        write_result<K, SOA>(M, A, x, idx, it, F, s_done);
        // Has the global counter run dry?  Private chunks keep a warp supplied for ~50 more trips, so without
        // this look a block would notice the end of the work long after the blocks around it, and a pipelined
        // launch would wait on it with SMs idle.  (Complete launches pack their tail anyway and skip the look.)
        const unsigned handed_out = (tail || !A.park) ? 0u : __ldcg((const unsigned*)A.counter);
        int bsel;
        const unsigned u = claim_chunked<K, SOA>(wc, stage, mbar, A, W, &s_tail, bsel);
        load_sample<K, SOA>(M, A, W, u, wc, stage, mbar, bsel, x, idx, it);
        if (!tail && A.park && handed_out >= W.total - W.first_dynamic) s_tail = 1;
      }
    }
  }
}


// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
// projection launch configuration: BLOCK threads, MINB resident blocks per SM (persistent grid), and
// which per-sample arrays live in shared memory.  CCP_PROJ_VARIANT (environment, read once) selects
// among the compiled configurations for tuning; the default is the best one measured on B200.
template <int K, int PANDA, bool SOA, int BLOCK, int MINB, int SM>
static cudaError_t launch_project_v(int sm_count, const ccp_model& M, const ccp_project_args& A, cudaStream_t st) {
  // persistent grid: MINB blocks per SM; a small batch is spread one warp's worth (32 samples) per block so that
  // it runs at one-warp-per-scheduler latency on many SMs instead of crowding a few
  long long need = (A.count + 31) / 32;
  long long cap = (long long)sm_count * MINB;
  if (A.adopt) need = cap;  // how many samples the previous launch parked is only known on the device
  int grid = (int)(need < cap ? need : cap);
  if (grid < 1) grid = 1;
  constexpr size_t smem = ccp_proj_smem_bytes<K, BLOCK, SM>();
  auto kern = ccp_project_kernel<K, PANDA, SOA, BLOCK, MINB, SM>;
  // Function attributes are per DEVICE: one flag per (instantiation, device), so that handles on several GPUs of one
  // process each raise the dynamic shared-memory limit on their own device.  (Relaxed atomics: setting it twice is harmless.)
  static std::atomic<unsigned char> attr_done[CCP_MAX_DEVICES];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= CCP_MAX_DEVICES || !attr_done[dev].load(std::memory_order_relaxed)) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < CCP_MAX_DEVICES) attr_done[dev].store(1, std::memory_order_relaxed);
  }
  if constexpr ((SM & CCP_SM_NOSTAGE) != 0) {
    ccp_project_args A2 = A;
    A2.stage_seeds = 0;  // no staging buffers in this configuration: the lanes load their seeds directly
    kern<<<grid, BLOCK, smem, st>>>(M, A2);
  } else {
    kern<<<grid, BLOCK, smem, st>>>(M, A);
  }
  return cudaGetLastError();
}

static int proj_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CCP_PROJ_VARIANT");
    v = e ? atoi(e) : 0;
  }
  return v;
}

template <int K, int PANDA, bool SOA>
static cudaError_t launch_project_g(int sm_count, const ccp_model& M, const ccp_project_args& A, cudaStream_t st) {
#ifdef CCP_TUNE
  if constexpr (K == 2) {
    switch (proj_variant()) {
      case 1: return launch_project_v<K, PANDA, SOA, 128, 3, 0>(sm_count, M, A, st);
      case 2: return launch_project_v<K, PANDA, SOA, 192, 2, 0>(sm_count, M, A, st);
      case 3: return launch_project_v<K, PANDA, SOA, 256, 1, 0>(sm_count, M, A, st);
      case 4: return launch_project_v<K, PANDA, SOA, 384, 1, CCP_SM_X>(sm_count, M, A, st);
      case 5: return launch_project_v<K, PANDA, SOA, 448, 1, CCP_SM_X>(sm_count, M, A, st);
      case 6: return launch_project_v<K, PANDA, SOA, 512, 1, CCP_SM_RXY | CCP_SM_NOSTAGE>(sm_count, M, A, st);
      case 7: return launch_project_v<K, PANDA, SOA, 448, 1, CCP_SM_RXY | CCP_SM_NOSTAGE>(sm_count, M, A, st);
      case 8: return launch_project_v<K, PANDA, SOA, 448, 1, CCP_SM_RXY>(sm_count, M, A, st);
      case 9: return launch_project_v<K, PANDA, SOA, 416, 1, CCP_SM_RXY>(sm_count, M, A, st);
      case 10: return launch_project_v<K, PANDA, SOA, 384, 1, CCP_SM_RXY>(sm_count, M, A, st);
      default: break;
    }
  } else {
    switch (proj_variant()) {
      case 1: return launch_project_v<K, PANDA, SOA, 320, 1, CCP_SM_SC>(sm_count, M, A, st);
      case 2: return launch_project_v<K, PANDA, SOA, 192, 1, CCP_SM_SC>(sm_count, M, A, st);
      case 3: return launch_project_v<K, PANDA, SOA, 128, 2, CCP_SM_SC>(sm_count, M, A, st);
      case 4: return launch_project_v<K, PANDA, SOA, 256, 1, CCP_SM_SC>(sm_count, M, A, st);
      case 5: return launch_project_v<K, PANDA, SOA, 288, 1, 0>(sm_count, M, A, st);
      case 6: return launch_project_v<K, PANDA, SOA, 320, 1, 0>(sm_count, M, A, st);
      case 7: return launch_project_v<K, PANDA, SOA, 320, 1, CCP_SM_RXY>(sm_count, M, A, st);
      case 8: return launch_project_v<K, PANDA, SOA, 384, 1, CCP_SM_RXY | CCP_SM_NOSTAGE>(sm_count, M, A, st);
      case 9: return launch_project_v<K, PANDA, SOA, 288, 1, CCP_SM_RXY>(sm_count, M, A, st);
      default: break;
    }
  }
#endif
  // One block per SM so the tail packing sees every live sample of the SM.  K = 2: 12 warps, everything in
  // registers (168, no spills); K = 3: 8 warps at 255 registers.  Both measured on B200 (profiles/)
  // against shared-memory staging and other block shapes.
  if constexpr (K == 2) return launch_project_v<K, PANDA, SOA, 384, 1, 0>(sm_count, M, A, st);
  else return launch_project_v<K, PANDA, SOA, 256, 1, 0>(sm_count, M, A, st);
}

#define CCP_CAT_(a, b, c, d) a##b##c##d
#define CCP_CAT(a, b, c, d) CCP_CAT_(a, b, c, d)
cudaError_t CCP_CAT(ccp_launch_project_K, CCP_TU_K, _P, CCP_TU_PANDA)(int sm_count, const ccp_model& M,
                                                                     const ccp_project_args& A, bool soa,
                                                                     cudaStream_t st) {
  constexpr int K = CCP_TU_K;
  constexpr int PANDA = CCP_TU_PANDA;  // 0 generic links, 1 structured alpha, 2 stock (ccp_core.h)
  if (soa) return launch_project_g<K, PANDA, true>(sm_count, M, A, st);
  return launch_project_g<K, PANDA, false>(sm_count, M, A, st);
}

