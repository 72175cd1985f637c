// sparse_score.cu — K4: sparse impact scoring + exact top-k over an inverted index.
//
// Replaces the Anserini / Lucene impact search of AnseriniSearch.retrieve_with_emb
// (reference retriever/anserini_search.py:143-216; io.anserini:anserini:0.25.0 SearchCollection -impact
// -pretokenized).  Score definition: scripts/asymmetric_sparse_infer.ipynb:207-228
//     score(q, d) = sum_{t in q ∩ d} count_q(t) * impact_d(t)        (integer arithmetic)
//
// HBM-bound gather-accumulate.  Documents are processed in blocks of 16384 consecutive ids whose int32
// accumulators live in shared memory (never in HBM).  A unit = (query, range of document blocks).  Per block:
//   1. every query term looks up its posting sub-range [blockptr[t][b], blockptr[t][b+1]) (no search);
//   2. the sub-ranges are flattened and all 256 threads stream (doc, impact) pairs with independent,
//      coalesced loads and atomicAdd count*impact into shared memory;
//   3. only the accumulators touched in this block are visited (the first atomicAdd that finds 0 records the
//      slot in a shared-memory list; > 4096 touches fall back to a full scan); entries beating the unit's running
//      threshold (a full 64-bit (score, ~id) key) are appended to a shared-memory candidate list; when the list
//      would overflow, a block-wide 64-bit radix select cuts list ∪ block back to the exact top-k and raises the
//      threshold.
// The unit's list goes to the workspace and lr_topk_merge produces the sorted result.
#include "common.cuh"

namespace lr {

constexpr int SS_THREADS = 256;
constexpr int SS_BLOCK_DOCS = 16384;
constexpr int SS_TERM_CHUNK = 256;
constexpr int SS_TOUCH_CAP = 4096;  // touched-accumulator list; a block that touches more falls back to a full scan

struct SSParams {
  const int32_t* q_indptr;
  const int32_t* q_tok;
  const int32_t* q_cnt;
  int64_t Q;
  const int64_t* post_indptr;
  const int32_t* post_doc;
  const uint16_t* post_imp;
  const uint32_t* blockptr;
  int64_t V, N;
  int nblk, S, k, cap;
  uint64_t* cand;   // [S][Q][cap]
  int32_t* counts;  // [S][Q]
};

__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
    if (lane >= off) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  int before = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SS_THREADS / 32; ++w) {
    const int t = warp_sums[w];
    if (w < warp) before += t;
    tot += t;
  }
  total = tot;
  __syncthreads();
  return before + incl - v;
}

__global__ void __launch_bounds__(SS_THREADS)
sparse_score_kernel(const SSParams p) {
  extern __shared__ __align__(16) uint8_t ss_smem[];
  int32_t* acc = reinterpret_cast<int32_t*>(ss_smem);
  uint64_t* list = reinterpret_cast<uint64_t*>(ss_smem + size_t(SS_BLOCK_DOCS) * 4);
  uint64_t* other = list + p.cap;
  int64_t* t_start = reinterpret_cast<int64_t*>(other + p.cap);
  int32_t* t_pre = reinterpret_cast<int32_t*>(t_start + SS_TERM_CHUNK);
  int32_t* t_w = t_pre + SS_TERM_CHUNK;
  uint16_t* touched = reinterpret_cast<uint16_t*>(t_w + SS_TERM_CHUNK);  // [SS_TOUCH_CAP] accumulators written this block
  __shared__ uint32_t hist[256];
  __shared__ int warp_sums[SS_THREADS / 32];
  __shared__ uint32_t s_n, s_nt, s_bin, s_rem, s_take;
  const int tid = threadIdx.x;

  for (int i = tid; i < SS_BLOCK_DOCS; i += SS_THREADS) acc[i] = 0;
  if (tid == 0) s_nt = 0;
  __syncthreads();

  const int64_t n_units = p.Q * p.S;
  for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
    const int64_t q = u / p.S;
    const int s = int(u % p.S);
    const int b0 = int((int64_t(s) * p.nblk) / p.S);
    const int b1 = int((int64_t(s + 1) * p.nblk) / p.S);
    const int qt0 = p.q_indptr[q];
    const int nterms = p.q_indptr[q + 1] - qt0;
    uint32_t n = 0;                    // entries in `list` (uniform copy of s_n between blocks)
    uint64_t thr = 0xFFFFFFFFull;      // candidates need key > thr; every score-0 key is <= this
    if (tid == 0) s_n = 0;
    // fast path (<= 256 query terms, the normal case): every thread owns one term and walks its block pointers with a
    // one-block-ahead prefetch, so the term set-up of a block costs no exposed global latency
    const bool one_chunk = nterms <= SS_TERM_CHUNK;
    const uint32_t* bp_row = nullptr;
    int64_t post_base = 0;
    int my_w = 0;
    uint32_t lo = 0, hi = 0, nxt = 0;
    if (one_chunk && tid < nterms) {
      const int t = p.q_tok[qt0 + tid];
      if (t >= 0 && t < p.V) {
        bp_row = p.blockptr + int64_t(t) * (p.nblk + 1);
        post_base = p.post_indptr[t];
        my_w = p.q_cnt[qt0 + tid];
        lo = bp_row[b0];
        hi = bp_row[b0 + 1];
        nxt = (b0 + 2 <= p.nblk) ? bp_row[b0 + 2] : hi;
      }
    }
    __syncthreads();

    for (int b = b0; b < b1; ++b) {
      const int64_t d0 = int64_t(b) * SS_BLOCK_DOCS;
      // ---- 1+2: accumulate the block's postings
      for (int tc = 0; tc < nterms; tc += SS_TERM_CHUNK) {
        int cnt = 0;
        if (one_chunk) {
          if (bp_row) {
            t_start[tid] = post_base + lo;
            cnt = int(hi - lo);
            t_w[tid] = my_w;
            lo = hi;
            hi = nxt;
            if (b + 3 <= p.nblk && b + 1 < b1) nxt = bp_row[b + 3];  // consumed two blocks from now
          }
        } else {
          const int i = tc + tid;
          if (i < nterms) {
            const int t = p.q_tok[qt0 + i];
            if (t >= 0 && t < p.V) {
              const uint32_t* bp = p.blockptr + int64_t(t) * (p.nblk + 1) + b;
              const uint32_t l2 = bp[0], h2 = bp[1];
              t_start[tid] = p.post_indptr[t] + l2;
              cnt = int(h2 - l2);
              t_w[tid] = p.q_cnt[qt0 + i];
            }
          }
        }
        int total;
        const int pre = block_exclusive_scan(cnt, warp_sums, total);
        t_pre[tid] = pre;
        __syncthreads();
        const int nt = min(SS_TERM_CHUNK, nterms - tc);
        for (int j = tid; j < total; j += SS_THREADS) {
          // largest i in [0, nt) with t_pre[i] <= j
          int l = 0, h = nt - 1;
          while (l < h) {
            const int mid = (l + h + 1) >> 1;
            if (t_pre[mid] <= j) l = mid; else h = mid - 1;
          }
          const int64_t idx = t_start[l] + (j - t_pre[l]);
          const int doc = __ldg(p.post_doc + idx);
          const int add = t_w[l] * int(__ldg(p.post_imp + idx));
          const int slot = doc - int(d0);
          // first touch of this accumulator in this block -> remember the slot (one shared-counter atomic per warp)
          const bool first = add != 0 && atomicAdd(&acc[slot], add) == 0;
          const unsigned am = __activemask();
          const unsigned fm = __ballot_sync(am, first);
          if (fm) {
            const int leader = __ffs(fm) - 1;
            uint32_t tbase = 0;
            if ((tid & 31) == leader) tbase = atomicAdd(&s_nt, uint32_t(__popc(fm)));
            tbase = __shfl_sync(am, tbase, leader);
            if (first) {
              const uint32_t tpos = tbase + __popc(fm & ((1u << (tid & 31)) - 1u));
              if (tpos < SS_TOUCH_CAP) touched[tpos] = uint16_t(slot);
            }
          }
        }
        __syncthreads();
      }
      // ---- 3: visit the touched accumulators (or all of them when the touched list overflowed)
      const uint32_t nt_all = s_nt;
      const bool dense = nt_all > SS_TOUCH_CAP;
      const int n_visit = dense ? SS_BLOCK_DOCS : int(nt_all);
      auto slot_of = [&](int i) -> int { return dense ? i : int(touched[i]); };
      int c = 0;
      for (int i = tid; i < n_visit; i += SS_THREADS) {
        const int sl = slot_of(i);
        const int sc = acc[sl];
        if (sc > 0 && make_key(uint32_t(sc), uint32_t(d0 + sl)) > thr) ++c;
      }
      uint32_t base = c ? atomicAdd(&s_n, uint32_t(c)) : 0u;
      __syncthreads();
      const uint32_t total_n = s_n;
      if (total_n <= uint32_t(p.cap)) {
        for (int i = tid; i < n_visit; i += SS_THREADS) {
          const int sl = slot_of(i);
          const int sc = acc[sl];
          if (sc != 0) {
            acc[sl] = 0;
            const uint64_t key = make_key(uint32_t(sc), uint32_t(d0 + sl));
            if (sc > 0 && key > thr) list[base++] = key;
          }
        }
        n = total_n;
      } else {
        // ---- overflow: exact k-th largest key of list[0..n) ∪ {block candidates}, then rebuild
        uint64_t prefix = 0;
        uint32_t remaining = uint32_t(p.k);
        uint64_t kstar = 0;
        for (int pass = 7; pass >= 0; --pass) {
          const int shift = pass * 8;
          hist[tid] = 0;
          __syncthreads();
          for (uint32_t i = tid; i < n; i += SS_THREADS) {
            const uint64_t key = list[i];
            if (pass == 7 || (key >> (shift + 8)) == prefix) atomicAdd(&hist[(key >> shift) & 0xFFu], 1u);
          }
          for (int i = tid; i < n_visit; i += SS_THREADS) {
            const int sl = slot_of(i);
            const int sc = acc[sl];
            if (sc > 0) {
              const uint64_t key = make_key(uint32_t(sc), uint32_t(d0 + sl));
              if (key > thr && (pass == 7 || (key >> (shift + 8)) == prefix))
                atomicAdd(&hist[(key >> shift) & 0xFFu], 1u);
            }
          }
          __syncthreads();
          if (tid < 32) {
            uint32_t bin, rem2;
            bool take_all;
            warp_find_bin_desc(hist, remaining, bin, rem2, take_all);
            if (tid == 0) {
              s_bin = bin;
              s_rem = rem2;
              s_take = take_all ? 1u : 0u;
            }
          }
          __syncthreads();
          prefix = (prefix << 8) | uint64_t(s_bin);
          remaining = s_rem;
          const bool take_all = s_take != 0;
          __syncthreads();
          if (take_all || pass == 0) {
            kstar = prefix << shift;
            break;
          }
        }
        if (tid == 0) s_n = 0;
        __syncthreads();
        for (uint32_t i = tid; i < n; i += SS_THREADS) {
          const uint64_t key = list[i];
          if (key >= kstar) other[atomicAdd(&s_n, 1u)] = key;
        }
        for (int i = tid; i < n_visit; i += SS_THREADS) {
          const int sl = slot_of(i);
          const int sc = acc[sl];
          if (sc != 0) {
            acc[sl] = 0;
            const uint64_t key = make_key(uint32_t(sc), uint32_t(d0 + sl));
            if (sc > 0 && key > thr && key >= kstar) other[atomicAdd(&s_n, 1u)] = key;
          }
        }
        __syncthreads();
        uint64_t* tmp = list;
        list = other;
        other = tmp;
        n = s_n;  // == k
        thr = kstar;
      }
      __syncthreads();
      if (tid == 0) s_nt = 0;
      // (the next block's first __syncthreads orders this reset before any new touch)
    }
    // ---- unit result
    uint64_t* dst = p.cand + (int64_t(s) * p.Q + q) * p.cap;
    for (uint32_t i = tid; i < n; i += SS_THREADS) dst[i] = list[i];
    if (tid == 0) p.counts[int64_t(s) * p.Q + q] = int32_t(n);
    __syncthreads();
  }
}

__global__ void build_blockptr_kernel(const int64_t* __restrict__ post_indptr, const int32_t* __restrict__ post_doc,
                                      int64_t V, int nblk, uint32_t* __restrict__ blockptr) {
  const int64_t total = V * int64_t(nblk + 1);
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    const int64_t t = e / (nblk + 1);
    const int b = int(e % (nblk + 1));
    const int64_t lo0 = post_indptr[t], hi0 = post_indptr[t + 1];
    const int64_t bound = int64_t(b) * SS_BLOCK_DOCS;
    int64_t lo = lo0, hi = hi0;  // first index with doc >= bound
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (int64_t(post_doc[mid]) < bound) lo = mid + 1; else hi = mid;
    }
    blockptr[e] = uint32_t(lo - lo0);
  }
}

struct SSPlan {
  int nblk, S, cap, grid;
  size_t smem, off_counts, off_cand, total_bytes;
};

static SSPlan ss_plan(int64_t Q, int64_t N, int k) {
  SSPlan pl{};
  pl.nblk = int((N + SS_BLOCK_DOCS - 1) / SS_BLOCK_DOCS);
  int cap = 2 * k > k + 256 ? 2 * k : k + 256;
  pl.cap = (cap + 63) / 64 * 64;
  pl.smem = size_t(SS_BLOCK_DOCS) * 4 + size_t(pl.cap) * 16 + SS_TERM_CHUNK * (8 + 4 + 4) + SS_TOUCH_CAP * 2;
  const int G = sm_count();
  const int ctas_per_sm = int((size_t(227) * 1024) / (pl.smem + 2048));
  const int slots = G * (ctas_per_sm < 1 ? 1 : ctas_per_sm);
  // enough units to fill the machine a few times over, at most one unit per document block
  int64_t S = (int64_t(4) * slots + Q - 1) / Q;
  if (S > pl.nblk) S = pl.nblk;
  if (S < 1) S = 1;
  pl.S = int(S);
  const int64_t units = Q * S;
  pl.grid = int(units < slots ? units : slots);
  auto align = [](size_t x) { return (x + 255) / 256 * 256; };
  pl.off_counts = 0;
  pl.off_cand = align(size_t(pl.S) * Q * 4);
  pl.total_bytes = align(pl.off_cand + size_t(pl.S) * Q * pl.cap * 8);
  return pl;
}

}  // namespace lr

using namespace lr;

extern "C" int lr_sparse_block_docs(void) { return SS_BLOCK_DOCS; }

extern "C" int lr_sparse_build_blockptr(const int64_t* post_indptr, const int32_t* post_doc, int64_t V, int64_t N,
                                        uint32_t* blockptr, void* stream) {
  LR_CHECK_ARG(post_indptr && blockptr && V >= 1 && N >= 1, "build_blockptr: bad arguments");
  const int nblk = int((N + SS_BLOCK_DOCS - 1) / SS_BLOCK_DOCS);
  const int64_t total = V * int64_t(nblk + 1);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  build_blockptr_kernel<<<unsigned(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(post_indptr, post_doc, V, nblk,
                                                                                        blockptr);
  LR_LAUNCH_CHECK();
  return LR_OK;
}

extern "C" size_t lr_sparse_score_workspace_bytes(int64_t Q, int64_t N, int k) {
  if (Q < 1 || N < 1 || k < 1) return 0;
  return ss_plan(Q, N, k).total_bytes;
}

extern "C" int lr_sparse_score_topk(const int32_t* q_indptr, const int32_t* q_tok, const int32_t* q_cnt, int64_t Q,
                                    const int64_t* post_indptr, const int32_t* post_doc, const uint16_t* post_imp,
                                    const uint32_t* blockptr, int64_t V, int64_t N, int64_t id_offset, int k,
                                    float* out_scores, int64_t* out_ids, uint64_t* out_keys, void* workspace,
                                    size_t ws_bytes, void* stream) {
  LR_CHECK_ARG(q_indptr && post_indptr && blockptr, "sparse_score: null pointer");
  LR_CHECK_ARG(Q >= 1 && N >= 1 && V >= 1, "sparse_score: Q, N, V must be >= 1");
  LR_CHECK_ARG(N < (int64_t(1) << 31), "sparse_score: N must be < 2^31 per shard");
  LR_CHECK_ARG(k >= 1 && k <= 1024, "sparse_score: k (%d) must be in [1, 1024]", k);
  LR_CHECK_ARG(id_offset >= 0 && id_offset + N <= (int64_t(1) << 32) - 2, "sparse_score: id_offset + N must stay below 2^32");
  LR_CHECK_ARG(out_scores || out_ids || out_keys, "sparse_score: no output requested");
  SSPlan pl = ss_plan(Q, N, k);
  if (!workspace || ws_bytes < pl.total_bytes || (uintptr_t(workspace) & 255)) {
    set_error("sparse_score: workspace too small or misaligned (%zu given, %zu needed)", ws_bytes, pl.total_bytes);
    return LR_EWORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  SSParams p{};
  p.q_indptr = q_indptr; p.q_tok = q_tok; p.q_cnt = q_cnt; p.Q = Q;
  p.post_indptr = post_indptr; p.post_doc = post_doc; p.post_imp = post_imp; p.blockptr = blockptr;
  p.V = V; p.N = N; p.nblk = pl.nblk; p.S = pl.S; p.k = k; p.cap = pl.cap;
  p.counts = reinterpret_cast<int32_t*>(ws + pl.off_counts);
  p.cand = reinterpret_cast<uint64_t*>(ws + pl.off_cand);
  cudaError_t e = cudaFuncSetAttribute(sparse_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl.smem));
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(smem=%zu) failed: %s", pl.smem, cudaGetErrorString(e));
    return LR_ECUDA;
  }
  sparse_score_kernel<<<pl.grid, SS_THREADS, pl.smem, st>>>(p);
  LR_LAUNCH_CHECK();
  return lr_topk_merge(p.cand, p.counts, pl.S, Q, Q, pl.cap, k, LR_SCORE_U32, id_offset, out_scores, out_ids, out_keys,
                       stream);
}
