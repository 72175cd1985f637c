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
};

// iterate over every valid candidate of query q; f(key)
template <class F>
__device__ __forceinline__ void for_each_candidate(const MergeParams& p, int64_t q, F&& f) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW = MERGE_THREADS / 32;
  for (int l = warp; l < p.L; l += NW) {
    const int64_t list = int64_t(l) * p.q_stride + q;
    const int n = p.counts ? p.counts[list] : p.cap;
    const uint64_t* src = p.keys + list * p.cap;
    for (int i = lane; i < n; i += 32) {
      const uint64_t key = src[i];
      if (key != 0ull) f(key);
    }
  }
}

__global__ void __launch_bounds__(MERGE_THREADS) topk_merge_kernel(const MergeParams p) {
  extern __shared__ uint64_t sel[];  // [kpad]
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_bin, s_remaining, s_take_all, s_total, s_nsel;
  const int64_t q = blockIdx.x;
  const int tid = threadIdx.x;

  if (tid == 0) {
    s_total = 0;
    s_nsel = 0;
  }
  __syncthreads();
  {
    uint32_t local = 0;
    for_each_candidate(p, q, [&](uint64_t) { ++local; });
    if (local) atomicAdd(&s_total, local);
  }
  __syncthreads();
  const uint32_t total = s_total;
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
      for_each_candidate(p, q, [&](uint64_t key) {
        if (pass == 7 || (key >> (shift + 8)) == prefix) atomicAdd(&hist[(key >> shift) & 0xFFu], 1u);
      });
      __syncthreads();
      if (tid < 32) {
        // warp 0: descending scan over the 256 bins, lane 0 owns the top 8
        const int base = 8 * (31 - tid);
        uint32_t c[8], sum = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          c[i] = hist[base + 7 - i];
          sum += c[i];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
          if (tid >= off) incl += t;
        }
        const uint32_t excl = incl - sum;
        if (excl < remaining && remaining <= incl) {
          uint32_t acc = excl;
          bool done = false;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (!done && acc + c[i] >= remaining) {
              s_bin = uint32_t(base + 7 - i);
              s_remaining = remaining - acc;
              s_take_all = (c[i] == remaining - acc) ? 1u : 0u;
              done = true;
            }
            if (!done) acc += c[i];
          }
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
    for_each_candidate(p, q, [&](uint64_t key) {
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
    if (p.out_keys) p.out_keys[q * p.out_key_stride + i] = valid ? make_key(hi, uint32_t(id)) : 0ull;
  }
}

__global__ void encode_keys_kernel(const float* scores, const int64_t* ids, int64_t n, uint64_t* keys) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t id = ids[i];
  keys[i] = (id < 0 || id >= 0xFFFFFFFFll) ? 0ull : make_key(f32_to_key(scores[i]), uint32_t(id));
}

// internal entry: like lr_topk_merge, with a row pitch for out_keys (a merged list can be written straight into a
// slot of another candidate-list array)
int topk_merge_strided(const uint64_t* keys, const int32_t* counts, int L, int64_t Q, int64_t q_stride, int cap, int k,
                       int score_kind, int64_t id_offset, float* out_scores, int64_t* out_ids, uint64_t* out_keys,
                       int64_t out_key_stride, cudaStream_t st) {
  MergeParams p{};
  p.keys = keys; p.counts = counts; p.L = L; p.Q = Q; p.q_stride = q_stride; p.cap = cap; p.k = k;
  int kpad = 2;
  while (kpad < k) kpad <<= 1;
  p.kpad = kpad;
  p.score_kind = score_kind; p.id_offset = id_offset;
  p.out_scores = out_scores; p.out_ids = out_ids; p.out_keys = out_keys;
  p.out_key_stride = out_key_stride;
  topk_merge_kernel<<<unsigned(Q), MERGE_THREADS, size_t(kpad) * 8, st>>>(p);
  LR_LAUNCH_CHECK();
  return LR_OK;
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
