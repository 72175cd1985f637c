// fuse_topk.cu — hybrid score fusion of two top-k lists per query, on device.
//
// Replaces the Python dict loops of fuse_scores_linear / fuse_scores_rrf
// (reference retriever/score_fuse_utils.py:48-90, 3-46; caller HybridSearch._fuse_results
// retriever/hybrid_search.py:207-232):
//   linear: per query and per system, scores are min-max normalised  (s - min) / (max - min + eps)  over the entries
//           that system returned, multiplied by the system's weight and summed over the union of document ids;
//   rrf:    1 / (k_rrf + rank) summed over the systems, rank = 1-based position in the system's sorted list.
// Arithmetic is float64 in the reference's operation order, so the fused scores are bit-exact with the numpy code.
//
// One CTA per query.  The two lists (k entries each, id -1 = padding) are concatenated in shared memory, sorted by id
// (bitonic), duplicates (a document returned by both systems) are combined, and the union is sorted by
// (fused score desc, id asc).  Output: [Q, 2k] ids (int64, -1 padded), fused scores (float64, -inf padded), counts.
#include "common.cuh"

namespace lr {

constexpr int FUSE_THREADS = 256;
constexpr int FUSE_MAX_SLOTS = 16;  // npad <= 4096

template <typename ScoreT> struct FuseParams {
  const ScoreT* s0;
  const int64_t* i0;
  const ScoreT* s1;
  const int64_t* i1;
  int64_t Q;
  int k0, k1, npad;  // npad = power of two >= k0 + k1
  int method;        // 0 = linear, 1 = rrf
  double w0, w1, eps, k_rrf;
  int64_t* out_ids;
  double* out_scores;
  int32_t* out_counts;
};

__device__ __forceinline__ bool fused_before(double sa, int64_t ia, double sb, int64_t ib) {
  // descending fused score, ascending id; padding (id < 0) last
  if ((ia < 0) != (ib < 0)) return ib < 0;
  if (sa != sb) return sa > sb;
  return ia < ib;
}

template <class Less>
__device__ __forceinline__ void bitonic_sort_pairs(int64_t* ids, double* val, int n, Less less) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < n / 2; t += FUSE_THREADS) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;  // this sub-sequence is sorted "less first"
        const bool swap = up ? less(ids[hi], val[hi], ids[lo], val[lo]) : less(ids[lo], val[lo], ids[hi], val[hi]);
        if (swap) {
          const int64_t ti = ids[lo]; ids[lo] = ids[hi]; ids[hi] = ti;
          const double tv = val[lo]; val[lo] = val[hi]; val[hi] = tv;
        }
      }
      __syncthreads();
    }
  }
}

// ScoreT = float (the searchers' result arrays) or double (dict-shaped callers: Python floats are float64)
template <typename ScoreT>
__global__ void __launch_bounds__(FUSE_THREADS) fuse_topk_kernel(const FuseParams<ScoreT> p) {
  extern __shared__ __align__(16) uint8_t fuse_smem[];
  int64_t* ids = reinterpret_cast<int64_t*>(fuse_smem);
  double* val = reinterpret_cast<double*>(ids + p.npad);
  __shared__ ScoreT s_min[2], s_max[2];
  __shared__ int s_cnt;
  const int64_t q = blockIdx.x;
  const int tid = threadIdx.x;
  const ScoreT* sc[2] = {p.s0 + q * p.k0, p.s1 + q * p.k1};
  const int64_t* id[2] = {p.i0 + q * p.k0, p.i1 + q * p.k1};
  const int kk[2] = {p.k0, p.k1};

  // ---- per-system min / max over the returned entries (np.min / np.max, score_fuse_utils.py:76-77)
  if (tid < 2) {
    s_min[tid] = INFINITY;
    s_max[tid] = -INFINITY;
  }
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  __shared__ ScoreT w_mn[FUSE_THREADS / 32], w_mx[FUSE_THREADS / 32];
  for (int sys = 0; sys < 2; ++sys) {
    ScoreT mn = INFINITY, mx = -INFINITY;
    for (int i = tid; i < kk[sys]; i += FUSE_THREADS)
      if (id[sys][i] >= 0) {
        mn = min(mn, sc[sys][i]);
        mx = max(mx, sc[sys][i]);
      }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, off));
      mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, off));
    }
    if ((tid & 31) == 0) {
      w_mn[tid >> 5] = mn;
      w_mx[tid >> 5] = mx;
    }
    __syncthreads();
    if (tid == 0) {
      ScoreT a = INFINITY, b = -INFINITY;
      for (int w = 0; w < FUSE_THREADS / 32; ++w) {
        a = min(a, w_mn[w]);
        b = max(b, w_mx[w]);
      }
      s_min[sys] = a;
      s_max[sys] = b;
    }
    __syncthreads();
  }

  // ---- contributions
  for (int i = tid; i < p.npad; i += FUSE_THREADS) {
    int64_t did = -1;
    double c = 0.0;
    int sys = -1, pos = 0;
    if (i < p.k0) { sys = 0; pos = i; }
    else if (i < p.k0 + p.k1) { sys = 1; pos = i - p.k0; }
    if (sys >= 0 && id[sys][pos] >= 0) {
      did = id[sys][pos];
      const double w = sys == 0 ? p.w0 : p.w1;
      if (p.method == 0) {
        const double s = double(sc[sys][pos]), mn = double(s_min[sys]), mx = double(s_max[sys]);
        c = (s - mn) / (mx - mn + p.eps) * w;      // scores_normed * weight (score_fuse_utils.py:78-79)
      } else {
        c = 1.0 / (p.k_rrf + double(pos + 1));     // 1 / (k + rank) (score_fuse_utils.py:38)
      }
    }
    ids[i] = did;
    val[i] = c;
  }
  __syncthreads();

  // ---- sort by id (padding last), combine the two contributions of a shared document
  bitonic_sort_pairs(ids, val, p.npad, [](int64_t ia, double, int64_t ib, double) {
    if ((ia < 0) != (ib < 0)) return ib < 0;
    return ia < ib;
  });
  // a slot is a head when its id differs from its left neighbour's; a document returned by both systems occupies two
  // adjacent slots.  All reads happen before any write (each thread owns <= FUSE_MAX_SLOTS strided slots).
  int64_t hid[FUSE_MAX_SLOTS];
  double hval[FUSE_MAX_SLOTS];
  int heads = 0;
#pragma unroll
  for (int sidx = 0; sidx < FUSE_MAX_SLOTS; ++sidx) {
    const int i = sidx * FUSE_THREADS + tid;
    hid[sidx] = -1;
    hval[sidx] = -INFINITY;
    if (i < p.npad) {
      const int64_t my = ids[i];
      const bool head = my >= 0 && (i == 0 || ids[i - 1] != my);
      if (head) {
        hid[sidx] = my;
        hval[sidx] = (i + 1 < p.npad && ids[i + 1] == my) ? val[i] + val[i + 1] : val[i];
        ++heads;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int sidx = 0; sidx < FUSE_MAX_SLOTS; ++sidx) {
    const int i = sidx * FUSE_THREADS + tid;
    if (i < p.npad) {
      ids[i] = hid[sidx];
      val[i] = hval[sidx];
    }
  }
  if (heads) atomicAdd(&s_cnt, heads);
  __syncthreads();

  // ---- final order: fused desc, id asc
  bitonic_sort_pairs(ids, val, p.npad, [](int64_t ia, double sa, int64_t ib, double sb) {
    return fused_before(sa, ia, sb, ib);
  });
  const int nout = p.k0 + p.k1;
  for (int i = tid; i < nout; i += FUSE_THREADS) {
    p.out_ids[q * nout + i] = ids[i];
    p.out_scores[q * nout + i] = ids[i] >= 0 ? val[i] : -INFINITY;
  }
  if (tid == 0 && p.out_counts) p.out_counts[q] = s_cnt;
}

}  // namespace lr

using namespace lr;

template <typename ScoreT>
static int fuse_launch(const ScoreT* scores0, const int64_t* ids0, int k0, const ScoreT* scores1, const int64_t* ids1, int k1,
                       int64_t Q, int method, double w0, double w1, double eps, double k_rrf, int64_t* out_ids,
                       double* out_scores, int32_t* out_counts, void* stream) {
  LR_CHECK_ARG(scores0 && ids0 && scores1 && ids1 && out_ids && out_scores, "fuse_topk: null pointer");
  LR_CHECK_ARG(Q >= 1 && k0 >= 1 && k1 >= 1 && k0 + k1 <= 4096, "fuse_topk: need Q >= 1 and 2 <= k0 + k1 <= 4096");
  LR_CHECK_ARG(method == 0 || method == 1, "fuse_topk: method must be 0 (linear) or 1 (rrf)");
  FuseParams<ScoreT> p{};
  p.s0 = scores0; p.i0 = ids0; p.s1 = scores1; p.i1 = ids1; p.Q = Q; p.k0 = k0; p.k1 = k1;
  int npad = 2;
  while (npad < k0 + k1) npad <<= 1;
  p.npad = npad;
  p.method = method; p.w0 = w0; p.w1 = w1; p.eps = eps; p.k_rrf = k_rrf;
  p.out_ids = out_ids; p.out_scores = out_scores; p.out_counts = out_counts;
  const size_t smem = size_t(npad) * 16;
  int rc = ensure_dyn_smem(reinterpret_cast<const void*>(fuse_topk_kernel<ScoreT>), 4096 * 16);  // the k0 + k1 <= 4096 maximum, once
  if (rc) return rc;
  fuse_topk_kernel<ScoreT><<<unsigned(Q), FUSE_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(p);
  LR_LAUNCH_CHECK();
  return LR_OK;
}

extern "C" int lr_fuse_topk(const float* scores0, const int64_t* ids0, int k0, const float* scores1, const int64_t* ids1,
                            int k1, int64_t Q, int method, double w0, double w1, double eps, double k_rrf,
                            int64_t* out_ids, double* out_scores, int32_t* out_counts, void* stream) {
  return fuse_launch<float>(scores0, ids0, k0, scores1, ids1, k1, Q, method, w0, w1, eps, k_rrf, out_ids, out_scores,
                            out_counts, stream);
}

extern "C" int lr_fuse_topk_f64(const double* scores0, const int64_t* ids0, int k0, const double* scores1,
                                const int64_t* ids1, int k1, int64_t Q, int method, double w0, double w1, double eps,
                                double k_rrf, int64_t* out_ids, double* out_scores, int32_t* out_counts, void* stream) {
  return fuse_launch<double>(scores0, ids0, k0, scores1, ids1, k1, Q, method, w0, w1, eps, k_rrf, out_ids, out_scores,
                             out_counts, stream);
}
