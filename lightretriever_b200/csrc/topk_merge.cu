// topk_merge.cu — exact top-k of the union of L candidate lists per query.
//
// Replaces the Python heap merge HybridSearch._add_to_heap (reference
// retriever/hybrid_search.py:182-205; twin in retriever/faiss_search.py:272-284) and the
// host-side shard merge of Faiss IndexShards (retriever/faiss_index.py:65-68).
//
// One CTA per query.  Candidates are unique u64 keys (score key << 32 | ~id), so the k-th largest
// key is found by an MSB-first 8-bit radix select over the lists (HBM/L2-bound: each pass streams the
// lists once; it stops as soon as the remaining bucket is taken whole), the survivors are gathered
// into shared memory, sorted by a bitonic network and decoded.
#include "common.cuh"

namespace lr {

constexpr int MERGE_THREADS = 256;
constexpr int MERGE_MAX_LISTS_SMEM = 512;  // per-list counts / offsets kept in shared memory (more lists: global passes)
constexpr int MERGE_STAGE_KEYS = 6144;     // candidates staged in shared memory (48 KB) when every SM has a query to merge
constexpr int MERGE_STAGE_KEYS_MAX = 24576;  // ... and up to 192 KB when there are fewer CTAs than SMs (online shapes)

struct MergeParams {
  const uint64_t* keys;
  const int32_t* counts;
  int L;
  int64_t Q, q_stride;
  int cap, k, kpad, score_kind;
  int64_t out_key_stride;  // row pitch of out_keys (>= k)
  int64_t id_offset;
  float* out_scores;
  int64_t* out_ids;
  uint64_t* out_keys;
  // two-level merge: the grid is Q x groups; group g reduces lists [g*lists_per_group, ...) of a query to a sorted top-k
  // written to out_keys laid out [groups][Q][out_key_stride] (level 1; out_scores / out_ids unused)
  int groups, lists_per_group;
  int stage_keys;  // capacity of the shared-memory staging area; groups with more candidates stream from L2
};

// One CTA per (query, group of lists).  The lists' counts are read in parallel, the candidates are pulled into shared
// memory with ONE flat, fully parallel sweep (no per-list dependent loads), and the radix passes, the gather and the sort
// run on shared memory.  Groups whose candidates do not fit fall back to streaming the lists from L2 in every pass.
__global__ void __launch_bounds__(MERGE_THREADS) topk_merge_kernel(const MergeParams p) {
  extern __shared__ uint64_t msm[];  // sel [kpad] | stage [stage_keys]
  uint64_t* sel = msm;
  uint64_t* stage = msm + p.kpad;
  __shared__ uint32_t hist[256];
  __shared__ int32_t s_off[MERGE_MAX_LISTS_SMEM + 1];
  __shared__ uint32_t s_bin, s_remaining, s_take_all, s_total, s_nsel;
  const int64_t q = int64_t(blockIdx.x) / p.groups;
  const int g = int(int64_t(blockIdx.x) % p.groups);
  const int l0 = g * p.lists_per_group;
  const int nl = min(p.lists_per_group, p.L - l0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- counts -> exclusive offsets
  bool staged = nl <= MERGE_MAX_LISTS_SMEM;
  if (tid == 0) {
    s_total = 0;
    s_nsel = 0;
  }
  __syncthreads();
  if (staged) {
    for (int l = tid; l < nl; l += MERGE_THREADS) {
      const int64_t list = int64_t(l0 + l) * p.q_stride + q;
      int n = p.counts ? p.counts[list] : p.cap;
      s_off[l + 1] = n < 0 ? 0 : (n > p.cap ? p.cap : n);
    }
    if (tid == 0) s_off[0] = 0;
    __syncthreads();
    if (warp == 0) {  // inclusive scan of s_off[1..nl] in chunks of 32
      int carry = 0;
      for (int base = 0; base < nl; base += 32) {
        const int i = base + lane;
        int v = i < nl ? s_off[i + 1] : 0;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int t = __shfl_up_sync(0xFFFFFFFFu, v, off);
          if (lane >= off) v += t;
        }
        if (i < nl) s_off[i + 1] = v + carry;
        carry += __shfl_sync(0xFFFFFFFFu, v, 31);
      }
      if (lane == 0) s_total = uint32_t(carry);
    }
    __syncthreads();
    staged = s_total <= uint32_t(p.stage_keys);
    // every thread has read s_total before anyone may overwrite it below (the unstaged path resets it)
    __syncthreads();
  }
  // iterate over every candidate of this (query, group); f(key).  Empty slots (key 0) are skipped.
  auto for_each_candidate = [&](auto&& f) {
    if (staged) {
      const uint32_t tot = s_total;
      for (uint32_t i = tid; i < tot; i += MERGE_THREADS) {
        const uint64_t key = stage[i];
        if (key != 0ull) f(key);
      }
    } else {
      for (int l = warp; l < nl; l += MERGE_THREADS / 32) {
        const int64_t list = int64_t(l0 + l) * p.q_stride + q;
        int n = p.counts ? p.counts[list] : p.cap;
        n = n < 0 ? 0 : (n > p.cap ? p.cap : n);  // same clamp as the staged path: a bad count never reads past its list
        const uint64_t* src = p.keys + list * p.cap;
        for (int i = lane; i < n; i += 32) {
          const uint64_t key = src[i];
          if (key != 0ull) f(key);
        }
      }
    }
  };
  if (staged) {
    // flat sweep: entry e of the concatenated lists -> (list by binary search on the offsets, index)
    const uint32_t tot = s_total;
    for (uint32_t e = tid; e < tot; e += MERGE_THREADS) {
      int lo = 0, hi = nl - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (uint32_t(s_off[mid]) <= e) lo = mid; else hi = mid - 1;
      }
      const int64_t list = int64_t(l0 + lo) * p.q_stride + q;
      stage[e] = p.keys[list * p.cap + (e - uint32_t(s_off[lo]))];
    }
    __syncthreads();
  }
  uint32_t total;
  {
    uint32_t local = 0;
    for_each_candidate([&](uint64_t) { ++local; });
    __syncthreads();
    if (tid == 0) s_nsel = 0;  // reuse as the non-empty counter
    __syncthreads();
    if (local) atomicAdd(&s_nsel, local);
    __syncthreads();
    total = s_nsel;
    __syncthreads();
    if (tid == 0) s_nsel = 0;
    __syncthreads();
  }
  const uint32_t kk = total < uint32_t(p.k) ? total : uint32_t(p.k);

  // ---- radix select of the kk-th largest key
  uint64_t threshold = 0;  // select keys >= threshold (0 -> everything)
  if (kk > 0 && kk < total) {
    uint64_t prefix = 0;
    uint32_t remaining = kk;
    for (int pass = 7; pass >= 0; --pass) {
      const int shift = pass * 8;
      hist[tid] = 0;
      __syncthreads();
      for_each_candidate([&](uint64_t key) {
        if (pass == 7 || (key >> (shift + 8)) == prefix) atomicAdd(&hist[(key >> shift) & 0xFFu], 1u);
      });
      __syncthreads();
      if (tid < 32) {
        uint32_t bin, rem2;
        bool take_all;
        warp_find_bin_desc(hist, remaining, bin, rem2, take_all);
        if (tid == 0) {
          s_bin = bin;
          s_remaining = rem2;
          s_take_all = take_all ? 1u : 0u;
        }
      }
      __syncthreads();
      prefix = (prefix << 8) | uint64_t(s_bin);
      remaining = s_remaining;
      const bool take_all = s_take_all != 0;
      __syncthreads();
      if (take_all || pass == 0) {
        threshold = prefix << shift;  // every key with this prefix (and all larger ones) is selected
        break;
      }
    }
  }

  // ---- gather the selected keys
  if (kk > 0) {
    for_each_candidate([&](uint64_t key) {
      if (key >= threshold) {
        const uint32_t pos = atomicAdd(&s_nsel, 1u);
        if (pos < uint32_t(p.kpad)) sel[pos] = key;
      }
    });
  }
  __syncthreads();
  const uint32_t nsel = s_nsel < uint32_t(p.kpad) ? s_nsel : uint32_t(p.kpad);
  for (int i = tid; i < p.kpad; i += MERGE_THREADS)
    if (uint32_t(i) >= nsel) sel[i] = 0ull;
  __syncthreads();

  // ---- bitonic sort, descending
  for (int size = 2; size <= p.kpad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < p.kpad / 2; t += MERGE_THREADS) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const uint64_t a = sel[lo], b = sel[hi];
        if ((a < b) == desc) {
          sel[lo] = b;
          sel[hi] = a;
        }
      }
      __syncthreads();
    }
  }

  // ---- decode
  for (int i = tid; i < p.k; i += MERGE_THREADS) {
    const uint64_t key = (i < p.kpad) ? sel[i] : 0ull;
    const bool valid = uint32_t(i) < kk && key != 0ull;
    const int64_t o = q * p.k + i;
    const uint32_t hi = key_hi(key);
    const int64_t id = p.id_offset + int64_t(key_id(key));
    if (p.out_scores)
      p.out_scores[o] = valid ? (p.score_kind == LR_SCORE_F32 ? key_to_f32(hi) : float(hi)) : -INFINITY;
    if (p.out_ids) p.out_ids[o] = valid ? id : int64_t(-1);
    if (p.out_keys) p.out_keys[(int64_t(g) * p.Q + q) * p.out_key_stride + i] = valid ? make_key(hi, uint32_t(id)) : 0ull;
  }
}

__global__ void encode_keys_kernel(const float* scores, const int64_t* ids, int64_t n, uint64_t* keys) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t id = ids[i];
  keys[i] = (id < 0 || id >= 0xFFFFFFFFll) ? 0ull : make_key(f32_to_key(scores[i]), uint32_t(id));
}

static int merge_launch(MergeParams& p, cudaStream_t st) {
  int kpad = 2;
  while (kpad < p.k) kpad <<= 1;
  p.kpad = kpad;
  // staging area: what the longest group can hold, within 48 KB when the grid fills the machine, 192 KB otherwise
  const int64_t group_max = int64_t(p.lists_per_group) * p.cap;
  int64_t limit = (p.Q * p.groups >= sm_count()) ? MERGE_STAGE_KEYS : MERGE_STAGE_KEYS_MAX;
  // sel [kpad] + stage + ~3 KB of static shared memory must stay within the 227 KB a CTA may own (k > 2048 with the
  // large staging area would not)
  const int64_t room = (int64_t(227) * 1024 - 4096 - int64_t(kpad) * 8) / 8;
  if (limit > room) limit = room;
  p.stage_keys = int(group_max < limit ? group_max : limit);
  const size_t smem = (size_t(kpad) + size_t(p.stage_keys)) * 8;
  // the most this kernel ever asks for (sel + the large staging area, within what a CTA may own), set once per device
  int rc = ensure_dyn_smem(reinterpret_cast<const void*>(topk_merge_kernel), 227 * 1024 - 4096);
  if (rc) return rc;
  topk_merge_kernel<<<unsigned(p.Q * p.groups), MERGE_THREADS, smem, st>>>(p);
  LR_LAUNCH_CHECK();
  return LR_OK;
}

// internal entry: like lr_topk_merge, with a row pitch for out_keys (a merged list can be written straight into a
// slot of another candidate-list array)
int topk_merge_strided(const uint64_t* keys, const int32_t* counts, int L, int64_t Q, int64_t q_stride, int cap, int k,
                       int score_kind, int64_t id_offset, float* out_scores, int64_t* out_ids, uint64_t* out_keys,
                       int64_t out_key_stride, cudaStream_t st) {
  MergeParams p{};
  p.keys = keys; p.counts = counts; p.L = L; p.Q = Q; p.q_stride = q_stride; p.cap = cap; p.k = k;
  p.score_kind = score_kind; p.id_offset = id_offset;
  p.out_scores = out_scores; p.out_ids = out_ids; p.out_keys = out_keys;
  p.out_key_stride = out_key_stride;
  p.groups = 1; p.lists_per_group = L;
  return merge_launch(p, st);
}

// Scratch for topk_merge_two_level: [groups][Q][k] keys of the first level (0 when one level is used).
size_t topk_merge_scratch_bytes(int L, int64_t Q, int cap, int k) {
  const int sms = sm_count();
  if (Q >= sms || L < 8) return 0;
  int64_t groups = (int64_t(L) * cap + MERGE_STAGE_KEYS_MAX - 1) / MERGE_STAGE_KEYS_MAX;
  if (groups < 4) groups = 4;  // also spreads the per-list latency of short lists
  if (groups > sms / Q) groups = sms / Q;
  if (groups > L / 4) groups = L / 4;
  return groups > 1 ? size_t(groups) * size_t(Q) * size_t(k) * 8 : 0;
}

// Few queries, many (possibly long) lists — the online shapes: one CTA per query would walk every list alone.  Level 1
// gives each query `groups` CTAs, each reducing a slice of the lists to a sorted top-k in `scratch`; level 2 merges those.
int topk_merge_two_level(const uint64_t* keys, const int32_t* counts, int L, int64_t Q, int64_t q_stride, int cap, int k,
                         int score_kind, int64_t id_offset, float* out_scores, int64_t* out_ids, uint64_t* out_keys,
                         int64_t out_key_stride, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  const size_t need = topk_merge_scratch_bytes(L, Q, cap, k);
  if (need == 0 || !scratch || scratch_bytes < need)
    return topk_merge_strided(keys, counts, L, Q, q_stride, cap, k, score_kind, id_offset, out_scores, out_ids, out_keys,
                              out_key_stride, st);
  const int groups = int(need / (size_t(Q) * size_t(k) * 8));
  MergeParams p{};
  p.keys = keys; p.counts = counts; p.L = L; p.Q = Q; p.q_stride = q_stride; p.cap = cap; p.k = k;
  p.score_kind = score_kind; p.id_offset = 0;
  p.out_keys = static_cast<uint64_t*>(scratch);
  p.out_key_stride = k;
  p.groups = groups; p.lists_per_group = (L + groups - 1) / groups;
  int rc = merge_launch(p, st);
  if (rc) return rc;
  return topk_merge_strided(static_cast<const uint64_t*>(scratch), nullptr, groups, Q, Q, k, k, score_kind, id_offset,
                            out_scores, out_ids, out_keys, out_key_stride, st);
}

}  // namespace lr

using namespace lr;

extern "C" int lr_topk_merge(const uint64_t* keys, const int32_t* counts, int L, int64_t Q, int64_t q_stride, int cap,
                             int k, int score_kind, int64_t id_offset, float* out_scores, int64_t* out_ids,
                             uint64_t* out_keys, void* stream) {
  LR_CHECK_ARG(keys, "topk_merge: null keys");
  LR_CHECK_ARG(L >= 1 && Q >= 1 && cap >= 1 && q_stride >= Q, "topk_merge: bad sizes L=%d Q=%lld cap=%d q_stride=%lld", L,
               (long long)Q, cap, (long long)q_stride);
  LR_CHECK_ARG(k >= 1 && k <= 4096, "topk_merge: k (%d) must be in [1, 4096]", k);
  LR_CHECK_ARG(score_kind == LR_SCORE_F32 || score_kind == LR_SCORE_U32, "topk_merge: bad score_kind %d", score_kind);
  LR_CHECK_ARG(id_offset >= 0, "topk_merge: negative id_offset");
  return topk_merge_strided(keys, counts, L, Q, q_stride, cap, k, score_kind, id_offset, out_scores, out_ids, out_keys, k,
                            static_cast<cudaStream_t>(stream));
}

extern "C" int lr_encode_keys(const float* scores, const int64_t* ids, int64_t n, uint64_t* keys, void* stream) {
  LR_CHECK_ARG(scores && ids && keys && n >= 0, "encode_keys: bad arguments");
  if (n == 0) return LR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  encode_keys_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(scores, ids, n, keys);
  LR_LAUNCH_CHECK();
  return LR_OK;
}
