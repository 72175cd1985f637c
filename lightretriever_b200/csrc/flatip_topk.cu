// flatip_topk.cu — K2: exact flat inner-product top-k on B200 (host side: plan, tensor maps, launch).
//
// Replaces faiss.IndexFlatIP.search as called by FaissIndex.search
// (reference retriever/faiss_index.py:27-40) plus the per-chunk heap merge
// (retriever/hybrid_search.py:182-205).  The kernel is umma_gemm_kernel<EPI_TOPK> in umma_gemm.cuh:
// rows = queries, columns = documents, the epilogue keeps per-(split, query) candidate lists which
// lr_topk_merge reduces to the final sorted top-k.
#include "umma_gemm.cuh"

namespace lr {

struct FlatipPlan {
  int cl, m_groups;
  int m_tiles, n_tiles, splits, band_size, n_bands, cap, grid, units, rounds;
  int64_t q_pad;
  size_t off_gthr, off_counts, off_cand, total_bytes;
};

static FlatipPlan make_plan(int64_t Q, int64_t N, int k) {
  FlatipPlan pl{};
  const GemmGeometry geo = plan_geometry(Q, env_int("LR_FLATIP_BAND", 32), env_int("LR_FLATIP_CLUSTER", 0));
  const int G = geo.n_clusters;  // clusters that run concurrently
  pl.cl = geo.cl; pl.m_groups = geo.m_groups;
  pl.m_tiles = geo.m_tiles;
  pl.n_tiles = int((N + BN - 1) / BN);
  pl.q_pad = int64_t(pl.m_tiles) * BM;
  // Candidate-list capacity.  The running threshold of a row only rises when its list is cut back to the top-k, so
  // between two cuts exactly cap-k entries are admitted and the number of documents consumed grows by cap/k per cut;
  // measured on B200 (profiles/): 2k..2.5k beats both smaller and larger lists.
  int cap = 2 * k > k + 64 ? 2 * k : k + 64;
  if (k <= 128 && cap < 256) cap = 256;
  cap = env_int("LR_FLATIP_CAP", cap);
  if (cap < k + 64) cap = k + 64;
  pl.cap = (cap + 63) / 64 * 64;
  // corpus splits: minimise rounds * tiles-per-unit; accept a larger split count only for a >0.5% gain
  const int64_t list_bytes = pl.q_pad * int64_t(pl.cap) * 8;
  int64_t s_max = int64_t(4) * G;
  const int64_t mem_cap = (int64_t(12) << 30) / (list_bytes > 0 ? list_bytes : 1);
  if (s_max > mem_cap) s_max = mem_cap;
  if (s_max > pl.n_tiles) s_max = pl.n_tiles;
  if (s_max < 1) s_max = 1;
  const int forced = env_int("LR_FLATIP_SPLITS", 0);
  double best = 1e300;
  int best_s = 1;
  for (int64_t s = 1; s <= s_max; ++s) {
    const int64_t units = int64_t(pl.m_groups) * s;
    const int64_t rounds = (units + G - 1) / G;
    const int64_t tiles = (pl.n_tiles + s - 1) / s;
    const double cost = double(rounds) * double(tiles + 1);  // +1: per-unit start-up
    if (cost < best * (1.0 - 0.005)) {
      best = cost;
      best_s = int(s);
    }
  }
  if (forced > 0 && forced <= pl.n_tiles) best_s = forced;
  pl.splits = best_s;
  pl.band_size = geo.band_size; pl.n_bands = geo.n_bands;
  pl.units = pl.m_groups * pl.splits;
  int clusters = pl.units < G ? pl.units : G;
  if (clusters < 1) clusters = 1;
  pl.grid = clusters * pl.cl;
  pl.rounds = (pl.units + clusters - 1) / clusters;
  auto align = [](size_t x) { return (x + 255) / 256 * 256; };
  pl.off_gthr = 0;
  pl.off_counts = align(pl.off_gthr + size_t(pl.q_pad) * 4);
  pl.off_cand = align(pl.off_counts + size_t(pl.splits) * pl.q_pad * 4);
  pl.total_bytes = align(pl.off_cand + size_t(pl.splits) * pl.q_pad * pl.cap * 8);
  return pl;
}

static thread_local FlatipPlan g_last_plan{};

static int check_flatip_args(const void* q, int64_t ldq, const void* corpus, int64_t ldc, int64_t Q, int64_t N,
                             int64_t d_used) {
  LR_CHECK_ARG(q && corpus, "flatip: null q or corpus");
  LR_CHECK_ARG(Q >= 1 && N >= 1, "flatip: Q (%lld) and N (%lld) must be >= 1", (long long)Q, (long long)N);
  LR_CHECK_ARG(N < (int64_t(1) << 31) - BN, "flatip: N (%lld) must be < 2^31 per shard", (long long)N);
  LR_CHECK_ARG(Q < (int64_t(1) << 31) - BM, "flatip: Q (%lld) too large", (long long)Q);
  LR_CHECK_ARG(d_used >= 8 && d_used % 8 == 0, "flatip: d_used (%lld) must be a positive multiple of 8", (long long)d_used);
  LR_CHECK_ARG(ldq >= d_used && ldc >= d_used, "flatip: row pitch (ldq=%lld, ldc=%lld) < d_used (%lld)",
               (long long)ldq, (long long)ldc, (long long)d_used);
  LR_CHECK_ARG((ldq * 2) % 16 == 0 && (ldc * 2) % 16 == 0, "flatip: row pitch must be a multiple of 16 bytes");
  LR_CHECK_ARG((uintptr_t(q) & 15) == 0 && (uintptr_t(corpus) & 15) == 0, "flatip: q/corpus must be 16-byte aligned");
  return LR_OK;
}

static void fill_params(GemmParams& prm, const FlatipPlan& pl, int64_t Q, int64_t N, int64_t d_used) {
  prm.rows = Q; prm.cols = N; prm.row_pad = pl.q_pad;
  prm.kblocks = int((d_used + BK - 1) / BK);
  prm.m_tiles = pl.m_tiles; prm.m_groups = pl.m_groups; prm.n_tiles = pl.n_tiles; prm.splits = pl.splits;
  prm.band_size = pl.band_size; prm.n_bands = pl.n_bands; prm.units = pl.units;
  prm.policy_a = l2_policy(env_int("LR_FLATIP_POLICY_A", 0));
  prm.policy_b = l2_policy(env_int("LR_FLATIP_POLICY_B", 0));
  prm.debug_flags = env_int("LR_FLATIP_DEBUG", 0);
}

}  // namespace lr

using namespace lr;

extern "C" size_t lr_flatip_workspace_bytes(int64_t Q, int64_t N, int k) {
  if (Q < 1 || N < 1 || k < 1) return 0;
  return make_plan(Q, N, k).total_bytes;
}

extern "C" int lr_flatip_last_plan(int64_t* out8) {
  const FlatipPlan& pl = g_last_plan;
  out8[0] = pl.m_tiles; out8[1] = pl.n_tiles; out8[2] = pl.splits; out8[3] = pl.band_size * pl.cl;
  out8[4] = pl.cap; out8[5] = pl.grid; out8[6] = pl.units; out8[7] = pl.rounds;
  return LR_OK;
}

extern "C" int lr_flatip_topk(const void* q, int64_t ldq, const void* corpus, int64_t ldc, int64_t Q, int64_t N,
                              int64_t d_used, const float* q_scale, const float* c_scale, int64_t id_offset, int k,
                              float* out_scores, int64_t* out_ids, uint64_t* out_keys, void* workspace,
                              size_t ws_bytes, void* stream) {
  int rc = check_flatip_args(q, ldq, corpus, ldc, Q, N, d_used);
  if (rc) return rc;
  LR_CHECK_ARG(k >= 1 && k <= 2048, "flatip: k (%d) must be in [1, 2048]", k);
  LR_CHECK_ARG(id_offset >= 0 && id_offset + N <= (int64_t(1) << 32) - 2, "flatip: id_offset + N must stay below 2^32");
  LR_CHECK_ARG(out_scores || out_ids || out_keys, "flatip: no output requested");
  FlatipPlan pl = make_plan(Q, N, k);
  if (!workspace || ws_bytes < pl.total_bytes || (uintptr_t(workspace) & 255)) {
    set_error("flatip: workspace too small or misaligned (%zu given, %zu needed, 256-byte aligned)", ws_bytes,
              pl.total_bytes);
    return LR_EWORKSPACE;
  }
  g_last_plan = pl;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tmA, tmB;
  if ((rc = make_tmap(&tmA, q, Q, d_used, ldq, BM))) return rc;
  if ((rc = make_tmap(&tmB, corpus, N, d_used, ldc, BN / pl.cl))) return rc;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  GemmParams prm{};
  fill_params(prm, pl, Q, N, d_used);
  prm.k = k; prm.cap = pl.cap;
  prm.q_scale = q_scale; prm.c_scale = c_scale;
  prm.gthr = reinterpret_cast<uint32_t*>(ws + pl.off_gthr);
  prm.counts = reinterpret_cast<int32_t*>(ws + pl.off_counts);
  prm.cand = reinterpret_cast<uint64_t*>(ws + pl.off_cand);
  if (!(prm.debug_flags & 2))  // debug bit 1: keep the previous call's thresholds (perfect-threshold timing experiment)
    LR_CUDA(cudaMemsetAsync(prm.gthr, 0, size_t(pl.q_pad) * 4, st));
  rc = pl.cl == 2 ? launch_umma_gemm<EPI_TOPK, 2>(tmA, tmB, prm, pl.grid, st)
                  : launch_umma_gemm<EPI_TOPK, 1>(tmA, tmB, prm, pl.grid, st);
  if (rc) return rc;
  return lr_topk_merge(prm.cand, prm.counts, pl.splits, Q, pl.q_pad, pl.cap, k, LR_SCORE_F32, id_offset, out_scores,
                       out_ids, out_keys, stream);
}

extern "C" int lr_flatip_scores(const void* q, int64_t ldq, const void* corpus, int64_t ldc, int64_t Q, int64_t N,
                                int64_t d_used, float* out_scores, void* stream) {
  int rc = check_flatip_args(q, ldq, corpus, ldc, Q, N, d_used);
  if (rc) return rc;
  LR_CHECK_ARG(out_scores, "flatip_scores: null output");
  FlatipPlan pl = make_plan(Q, N, 1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tmA, tmB;
  if ((rc = make_tmap(&tmA, q, Q, d_used, ldq, BM))) return rc;
  if ((rc = make_tmap(&tmB, corpus, N, d_used, ldc, BN / pl.cl))) return rc;
  GemmParams prm{};
  fill_params(prm, pl, Q, N, d_used);
  prm.dbg_scores = out_scores;
  return pl.cl == 2 ? launch_umma_gemm<EPI_STORE, 2>(tmA, tmB, prm, pl.grid, st)
                    : launch_umma_gemm<EPI_STORE, 1>(tmA, tmB, prm, pl.grid, st);
}
