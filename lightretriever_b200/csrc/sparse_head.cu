// sparse_head.cu — K3: document sparse head.
//
//   lr_sparse_head_max     out[b, v] = log1p(relu(max_{t valid} (h[b,t,:] . W[v,:] + bias[v])))
//       replaces aggregate()/max_linear_mapping  (reference finetune/sparse_pooling.py:244-278,
//       utils/max_linear_map.py:10-90: a Python loop over seq positions, one GEMM + 4 elementwise launches
//       per token) and relu_/log1p_ (finetune/modeling_hybrid.py:183-187).
//       One umma_gemm_kernel<EPI_MAXTOK> launch: rows = vocabulary (A = lm_head.weight [V, d]), columns = tokens
//       (B = hidden [B*S, d]); the per-token logits never leave TMEM/registers.
//   lr_sparsify_quantize   top_k_sampling (finetune/sparse_pooling.py:89-106) + the quantiser of
//       convert_sparse_reps_to_json_pt (finetune/sparse_converter_mixin.py:103-160) -> CSR (indptr, tok, impact).
#include "umma_gemm.cuh"

namespace lr {

constexpr int SP_THREADS = 256;

// ---- per document: threshold = k_eff-th largest value (radix select over f32 keys), count of surviving non-zeros
__global__ void __launch_bounds__(SP_THREADS)
sparsify_select_kernel(const float* __restrict__ reps, int64_t V, int k_eff, float quant, float* thr_out,
                       int32_t* cnt_out) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_bin, s_remaining;
  __shared__ int s_cnt;
  const int64_t b = blockIdx.x;
  const float* x = reps + b * V;
  const int tid = threadIdx.x;
  uint32_t thr_key = 0;  // keep everything
  if (k_eff > 0) {
    uint32_t prefix = 0, remaining = uint32_t(k_eff);
    for (int pass = 3; pass >= 0; --pass) {
      const int shift = pass * 8;
      hist[tid] = 0;
      __syncthreads();
      for (int64_t i = tid; i < V; i += SP_THREADS) {
        const uint32_t h = f32_to_key(x[i]);
        if (pass == 3 || (h >> (shift + 8)) == prefix) atomicAdd(&hist[(h >> shift) & 0xFFu], 1u);
      }
      __syncthreads();
      if (tid < 32) {
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
              done = true;
            }
            if (!done) acc += c[i];
          }
        }
      }
      __syncthreads();
      prefix = (prefix << 8) | s_bin;
      remaining = s_remaining;
      __syncthreads();
    }
    thr_key = prefix;
  }
  const float thr = k_eff > 0 ? key_to_f32(thr_key) : -INFINITY;
  // torch: indices_to_remove = scores < kth  ->  keep x >= thr (all ties kept)
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  int local = 0;
  for (int64_t i = tid; i < V; i += SP_THREADS) {
    const float v = x[i];
    if (v >= thr) {
      const float qv = rintf(fmaxf(v, 0.0f) * quant);  // torch.round = round half to even
      if (qv >= 1.0f) ++local;
    }
  }
  if (local) atomicAdd(&s_cnt, local);
  __syncthreads();
  if (tid == 0) {
    thr_out[b] = thr;
    cnt_out[b] = s_cnt;
  }
}

__global__ void exclusive_scan_kernel(const int32_t* cnt, int64_t B, int32_t* indptr) {
  // single thread block, sequential over chunks (B is a batch size)
  if (threadIdx.x == 0) {
    long long run = 0;
    for (int64_t b = 0; b < B; ++b) {
      indptr[b] = int32_t(run);
      run += cnt[b];
    }
    indptr[B] = int32_t(run);
  }
}

// ---- per document: ordered write of (token id, impact)
__global__ void __launch_bounds__(SP_THREADS)
sparsify_write_kernel(const float* __restrict__ reps, int64_t V, float quant, const float* thr_in,
                      const int32_t* __restrict__ indptr, int32_t* tok, uint16_t* impact, int64_t cap) {
  __shared__ int warp_tot[SP_THREADS / 32];
  __shared__ int s_base;
  const int64_t b = blockIdx.x;
  const float* x = reps + b * V;
  const float thr = thr_in[b];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = indptr[b];
  __syncthreads();
  for (int64_t i0 = 0; i0 < V; i0 += SP_THREADS) {
    const int64_t i = i0 + tid;
    float qv = 0.0f;
    if (i < V) {
      const float v = x[i];
      if (v >= thr) qv = rintf(fmaxf(v, 0.0f) * quant);
    }
    const bool keep = qv >= 1.0f;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, keep);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SP_THREADS / 32; ++w) {
      const int t = warp_tot[w];
      if (w < warp) before += t;
      total += t;
    }
    const int base = s_base;
    if (keep) {
      const int64_t pos = int64_t(base) + before + __popc(m & ((1u << lane) - 1u));
      if (pos < cap) {
        tok[pos] = int32_t(i);
        impact[pos] = uint16_t(fminf(qv, 65535.0f));
      }
    }
    __syncthreads();
    if (tid == 0) s_base = base + total;
    __syncthreads();
  }
}

}  // namespace lr

using namespace lr;

extern "C" int lr_sparse_head_max(const void* hidden, const void* W, const float* bias, const uint8_t* mask,
                                  int64_t B, int64_t S, int64_t d, int64_t V, int relu, int log1p, float* out,
                                  void* stream) {
  LR_CHECK_ARG(hidden && W && mask && out, "sparse_head: null pointer");
  LR_CHECK_ARG(B >= 1 && S >= 1 && V >= 1, "sparse_head: B, S, V must be >= 1");
  LR_CHECK_ARG(d >= 8 && d % 8 == 0, "sparse_head: d (%lld) must be a positive multiple of 8", (long long)d);
  LR_CHECK_ARG((uintptr_t(hidden) & 15) == 0 && (uintptr_t(W) & 15) == 0, "sparse_head: hidden/W must be 16-byte aligned");
  LR_CHECK_ARG(B * S < (int64_t(1) << 31) - BN && V < (int64_t(1) << 31) - BM, "sparse_head: B*S or V too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tmA, tmB;
  int rc;
  // a unit covers whole documents and at least ~2 column tiles
  int sps = int((2 * BN + S - 1) / S);
  if (sps < 1) sps = 1;
  sps = env_int("LR_SPARSE_HEAD_DOCS_PER_UNIT", sps);
  if (sps > B) sps = int(B);
  const int splits = int((B + sps - 1) / sps);
  // Vocabulary bands x document splits.  With enough of both, fixed teams of clusters walk the splits of a band in step
  // (umma_gemm.cuh for_each_unit): a hidden-state tile is fetched from HBM once per team and the band's lm_head rows
  // stay L2-resident; the team schedule runs as cta_group::2 pairs.  Otherwise round robin over 12-tile bands.
  int mode = env_int("LR_SPARSE_HEAD_CLUSTER", 0);
  int team_band = 0;
  if (env_int("LR_SPARSE_HEAD_SCHED", 1) != 0 && mode != 1 && V > 8 * 2 * BM && splits >= 2) {
    const int m_groups = int((V + 2 * BM - 1) / (2 * BM));
    team_band = env_int("LR_SPARSE_HEAD_TEAM_BAND", plan_team_band(m_groups, sm_count() / 2, splits, 0.85));
    if (team_band > m_groups) team_band = m_groups;
    if (team_band > 0 && mode == 0) mode = 3;
  }
  const GemmGeometry geo = plan_geometry(V, env_int("LR_SPARSE_HEAD_BAND", 12), mode);
  if ((rc = make_tmap(&tmA, W, V, d, d, BM))) return rc;
  if ((rc = make_tmap(&tmB, hidden, B * S, d, d, BN / geo.cl))) return rc;
  GemmParams prm{};
  prm.rows = V; prm.cols = B * S;
  prm.m_tiles = geo.m_tiles; prm.m_groups = geo.m_groups;
  prm.row_pad = int64_t(prm.m_tiles) * BM;
  prm.kblocks = int((d + BK - 1) / BK);
  prm.seg_len = S; prm.n_segs = B;
  prm.segs_per_split = sps;
  prm.splits = splits;
  prm.band_size = geo.band_size; prm.n_bands = geo.n_bands;
  prm.units = prm.m_groups * prm.splits;
  prm.bias = bias; prm.mask = mask; prm.out = out; prm.relu = relu; prm.log1p = log1p;
  prm.policy_a = l2_policy(env_int("LR_SPARSE_HEAD_POLICY_A", 0));
  prm.policy_b = l2_policy(env_int("LR_SPARSE_HEAD_POLICY_B", 0));
  int clusters = prm.units < geo.n_clusters ? prm.units : geo.n_clusters;
  uint32_t* team_ctr = nullptr;
  if (team_band > 0 && geo.cl == 2) {
    prm.sched = 1;
    prm.team_window = env_int("LR_SPARSE_HEAD_TEAM_WINDOW", 1);
    prm.band_size = team_band;
    prm.n_bands = (prm.m_groups + team_band - 1) / team_band;
    clusters = geo.n_clusters;
    // progress counters of the teams, one per (band, team): stream-ordered scratch, released after the launch
    const size_t ctr_bytes = size_t(prm.n_bands) * size_t(geo.n_clusters) * 4;
    LR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&team_ctr), ctr_bytes, st));
    cudaError_t e = cudaMemsetAsync(team_ctr, 0, ctr_bytes, st);
    if (e != cudaSuccess) {
      cudaFreeAsync(team_ctr, st);
      set_error("sparse_head: cudaMemsetAsync failed: %s", cudaGetErrorString(e));
      return LR_ECUDA;
    }
    prm.team_ctr = team_ctr;
  }
  const int grid = clusters * geo.cl;
  if (geo.pair) rc = launch_umma_gemm<EPI_MAXTOK, 2, true>(tmA, tmB, prm, grid, st);
  else rc = geo.cl == 2 ? launch_umma_gemm<EPI_MAXTOK, 2>(tmA, tmB, prm, grid, st)
                        : launch_umma_gemm<EPI_MAXTOK, 1>(tmA, tmB, prm, grid, st);
  if (team_ctr) cudaFreeAsync(team_ctr, st);
  return rc;
}

extern "C" size_t lr_sparsify_scratch_bytes(int64_t B, int64_t V) {
  (void)V;
  return size_t(B > 0 ? B : 0) * 8 + 256;
}

extern "C" int lr_sparsify_quantize(const float* reps, int64_t B, int64_t V, int top_k, int min_keep, float quant,
                                    int32_t* indptr, int32_t* tok, uint16_t* impact, int64_t cap, void* scratch,
                                    void* stream) {
  LR_CHECK_ARG(reps && indptr && scratch, "sparsify: null pointer");
  LR_CHECK_ARG(B >= 1 && V >= 1 && V < (int64_t(1) << 31), "sparsify: bad sizes");
  LR_CHECK_ARG(cap >= 0 && (cap == 0 || (tok && impact)), "sparsify: null tok/impact");
  LR_CHECK_ARG(quant > 0.0f, "sparsify: quantization factor must be > 0");
  int k_eff = 0;  // top_k <= 0 disables the filter (sparse_pooling.py:98-99)
  if (top_k > 0) {
    k_eff = top_k > min_keep ? top_k : min_keep;
    if (k_eff > V) k_eff = int(V);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* thr = static_cast<float*>(scratch);
  int32_t* cnt = reinterpret_cast<int32_t*>(thr + B);
  sparsify_select_kernel<<<unsigned(B), SP_THREADS, 0, st>>>(reps, V, k_eff, quant, thr, cnt);
  LR_LAUNCH_CHECK();
  exclusive_scan_kernel<<<1, 32, 0, st>>>(cnt, B, indptr);
  LR_LAUNCH_CHECK();
  sparsify_write_kernel<<<unsigned(B), SP_THREADS, 0, st>>>(reps, V, quant, thr, indptr, tok, impact, cap);
  LR_LAUNCH_CHECK();
  // nnz check (one small D2H): the caller sized `cap`; report overflow instead of silently truncating
  int32_t nnz = 0;
  LR_CUDA(cudaMemcpyAsync(&nnz, indptr + B, 4, cudaMemcpyDeviceToHost, st));
  LR_CUDA(cudaStreamSynchronize(st));
  if (int64_t(nnz) > cap) {
    set_error("sparsify: nnz %d exceeds capacity %lld", nnz, (long long)cap);
    return LR_EWORKSPACE;
  }
  return LR_OK;
}
