// Host-side check of the persistent kernels' tile schedules (dinox_b200/csrc/gemm_core.cuh TileWalker): for a grid of
// problem shapes and cluster counts, the union of what every cluster walks must be every tile exactly once, for
// the strided walk, the contiguous-run walk (resident A) and the column walk (resident B).  No GPU needed.
#include <cstdio>
#include <vector>
#include "gemm_core.cuh"
using namespace dinox::gemm;

static int check(int num_m, int num_n, int outer, bool split, int m_fastest, int ncl, int mode) {
  CoreParams p{};
  p.num_m_tiles = num_m; p.num_n_tiles = num_n; p.m_fastest = m_fastest;
  p.batches = split ? 1 : outer; p.splits = split ? outer : 1;
  const int num_super = num_m * num_n * outer;
  std::vector<int> seen((size_t)num_super, 0);
  long visited = 0;
  int max_per_cluster = 0, min_per_cluster = 1 << 30;
  for (int cid = 0; cid < ncl; ++cid) {
    TileWalker w;
    if (mode == 0) w.init(p, num_m, cid, ncl, num_super);
    else if (mode == 1) {
      const int per = (num_super + ncl - 1) / ncl, first = cid * per;
      const int count = per < num_super - first ? per : num_super - first;
      w.init_range(p, num_m, first, count);
    } else w.init_columns(p, num_m, cid, ncl);
    int n = 0, prev_slow = -1, changes = 0;
    for (; w.valid(); w.next(), ++n) {
      const TileCoord c = w.coord(p, 1, 0);
      const int o = split ? c.split : c.batch;
      if (c.m_tile < 0 || c.m_tile >= num_m || c.n_tile < 0 || c.n_tile >= num_n || o < 0 || o >= outer) return 1;
      ++seen[((size_t)o * num_m + c.m_tile) * num_n + c.n_tile];
      const int slow = mode == 2 ? c.n_tile : (m_fastest ? c.n_tile : c.m_tile) + o * 100000;
      if (slow != prev_slow) { ++changes; prev_slow = slow; }
    }
    visited += n;
    if (n > max_per_cluster) max_per_cluster = n;
    if (n < min_per_cluster) min_per_cluster = n;
    // a contiguous run changes its slow index at most ceil(n / fast extent) + 1 times
    if (mode == 1 && n > 0) {
      const int nfast = m_fastest ? num_m : num_n;
      if (changes > (n + nfast - 1) / nfast + 1) return 2;
    }
  }
  if (visited != num_super) return 3;
  for (int v : seen) if (v != 1) return 4;
  // balance: nobody carries more than one tile (strided / runs) or two (columns: rounding of the tail) above the mean
  const int mean_up = (num_super + ncl - 1) / ncl;
  if (max_per_cluster > mean_up + (mode == 2 ? 2 : 0)) return 5;
  (void)min_per_cluster;
  return 0;
}

int main() {
  const int ms[] = {1, 2, 3, 16, 17, 32, 63, 256}, ns[] = {1, 2, 3, 9, 34, 67, 74, 75, 157, 256}, cls[] = {1, 3, 74, 148};
  long cases = 0;
  for (int m : ms) for (int n : ns) for (int ncl_max : cls) for (int mode = 0; mode < 3; ++mode)
    for (int outer : {1, 3}) for (int split = 0; split < 2; ++split) for (int mf = 0; mf < 2; ++mf) {
      if (mode == 2 && (outer != 1 || split)) continue;        // the column walk has no batch / split-K dimension
      if (split && outer == 1) continue;
      const long super = (long)m * n * outer;
      const int ncl = super < ncl_max ? (int)super : ncl_max;    // launch_grid(): never more clusters than tiles
      const int rc = check(m, n, outer, split != 0, mf, ncl, mode);
      ++cases;
      if (rc) { std::printf("FAIL rc=%d m=%d n=%d outer=%d split=%d mf=%d ncl=%d mode=%d\n", rc, m, n, outer, split, mf, ncl, mode); return 1; }
    }
  std::printf("OK %ld schedules\n", cases);
  return 0;
}
