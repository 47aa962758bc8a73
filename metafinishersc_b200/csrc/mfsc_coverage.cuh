// mfsc_coverage.cuh -- stage 5: per-contig coverage profile from alignment intervals.
//
// Replaces the per-base Python loop `cov[contig][i] += 1 for i in range(S1-1, E1)` of
// misassemblyFixerLib/merger.py:126-140 (assignCoverage) by a difference array and one prefix sum:
// every selected row adds +1 at off[contig]+S1-1 and -1 at off[contig]+E1.  Intervals never cross
// a contig end, so the prefix sum over the concatenated contigs restarts from zero at every
// contig start by itself -- the segments are implicit and the result is bit-exact integers.
// The scan is a single-pass chained scan: a warp takes the next tile from a ticket counter, scans
// it in registers with shuffles, publishes its aggregate and looks back over the preceding tiles'
// published (flag,value) words, 32 tiles per step.  Traffic: 4 B read + 4 B written per base.
//
// The profile can stay resident (mfsc_cov_build): its consumer, the group C-test of the misassembly fixer
// (misassemblyFixerLib/mergerHelper.py:137-163 formatIntervalAndFilter, `np.sum(coverageData[start:end])` per
// candidate interval), only needs sums over intervals.  k_cov_tile_sums reduces every 512-base tile to one 64-bit
// sum (4 B read per base, 8 B written per tile); after an exclusive scan of those, k_cov_interval_sums answers a
// query with two prefix look-ups and at most two partial tiles.
#pragma once
#include "mfsc_platform.cuh"

#define SCAN_ITEMS 16                       // per lane
#define SCAN_TILE (SCAN_ITEMS * MFSC_LANES) // per warp

MFSC_KERNEL k_cov_diff(const int* ref_id, const int* s1, const int* e1, long long n, const long long* contig_off,
                       long long total, int* diff) {
  for (long long r = thread_global(); r < n; r += threads_total()) {
    const long long base = contig_off[ref_id[r]];
    atomic_add_i32(&diff[base + s1[r] - 1], 1);
    const long long end = base + e1[r];
    if (end < total) atomic_add_i32(&diff[end], -1);
  }
}

#ifdef MFSC_EMU
static inline int wscan_incl_i32(int v) { return v; }
#else
MFSC_DEV int wscan_incl_i32(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(MFSC_FULL, v, o); if (lane_id() >= o) v += t; }
  return v;
}
#endif

// state word: high 32 bits = flag (0 empty, 1 aggregate, 2 inclusive prefix), low 32 bits = value
MFSC_KERNEL k_cov_scan(const int* diff, long long n, int* out, unsigned long long* state, unsigned* ticket) {
  const int lane = lane_id();
  const long long n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  for (;;) {
    long long tile = 0;
    if (lane == 0) tile = (long long)atomic_add_u32(ticket, 1u);
    tile = wbcast_i64(tile, 0);
    if (tile >= n_tiles) break;
    const long long base = tile * SCAN_TILE + (long long)lane * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    if (base + SCAN_ITEMS <= n) {
      const int4* p = reinterpret_cast<const int4*>(diff + base);
#pragma unroll
      for (int t = 0; t < SCAN_ITEMS / 4; t++) {
        int4 x = p[t];
        v[4 * t] = x.x; v[4 * t + 1] = x.y; v[4 * t + 2] = x.z; v[4 * t + 3] = x.w;
      }
    } else {
#pragma unroll
      for (int t = 0; t < SCAN_ITEMS; t++) v[t] = (base + t < n) ? diff[base + t] : 0;
    }
    int sum = 0;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; t++) { sum += v[t]; v[t] = sum; }
    const int incl = wscan_incl_i32(sum);
    const int agg = wbcast_i32(incl, MFSC_LANES - 1);
    int prefix = 0;
    if (tile == 0) {
      if (lane == 0) st_volatile_u64(&state[0], (2ull << 32) | (unsigned)agg);
    } else {
      if (lane == 0) st_volatile_u64(&state[tile], (1ull << 32) | (unsigned)agg);
      long long pos = tile - 1;
      for (;;) {
        const long long idx = pos - lane;
        unsigned long long w = (3ull << 32);   // "beyond the start": treated as ready, value 0
        if (idx >= 0) w = ld_volatile_u64(&state[idx]);
        const unsigned flag = (unsigned)(w >> 32);
        if (wballot(flag == 0u)) continue;       // a predecessor has not published yet
        const unsigned incl_mask = wballot(flag >= 2u);
        int val = (int)(unsigned)(w & 0xffffffffull);
        if (flag == 3u) val = 0;
        if (incl_mask) {
          const int first = dev_ffs32(incl_mask) - 1;   // nearest tile holding an inclusive prefix
          prefix += wsum_i32(lane <= first ? val : 0);
          break;
        }
        prefix += wsum_i32(val);
        pos -= MFSC_LANES;
      }
      if (lane == 0) st_volatile_u64(&state[tile], (2ull << 32) | (unsigned)(prefix + agg));
    }
    const int lane_off = prefix + incl - sum;
    if (base + SCAN_ITEMS <= n) {
      int4* q = reinterpret_cast<int4*>(out + base);
#pragma unroll
      for (int t = 0; t < SCAN_ITEMS / 4; t++) {
        int4 x;
        x.x = v[4 * t] + lane_off; x.y = v[4 * t + 1] + lane_off; x.z = v[4 * t + 2] + lane_off; x.w = v[4 * t + 3] + lane_off;
        q[t] = x;
      }
    } else {
#pragma unroll
      for (int t = 0; t < SCAN_ITEMS; t++) if (base + t < n) out[base + t] = v[t] + lane_off;
    }
  }
}

// 64-bit sum of every scan tile of the coverage profile
MFSC_KERNEL k_cov_tile_sums(const int* cov, long long n, long long* tile_sum) {
  const int lane = lane_id();
  const long long n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  for (long long tile = warp_global(); tile < n_tiles; tile += warps_total()) {
    const long long base = tile * SCAN_TILE;
    long long s = 0;
    if (base + SCAN_TILE <= n) {
      const int4* p = reinterpret_cast<const int4*>(cov + base);
#pragma unroll
      for (int t = 0; t < SCAN_ITEMS / 4; t++) { const int4 x = p[t * MFSC_LANES + lane]; s += (long long)x.x + x.y + x.z + x.w; }
    } else {
      for (long long i = base + lane; i < n; i += MFSC_LANES) s += cov[i];
    }
    s = wsum_i64(s);
    if (lane == 0) tile_sum[tile] = s;
  }
}

// One warp per query: sum of cov[off[c] + a, off[c] + b) with a, b normalised like a Python slice of the contig's list.
MFSC_KERNEL k_cov_interval_sums(const int* cov, const long long* tile_pref, const long long* contig_off, const int* contig,
                                const long long* start, const long long* end, long long n_q, long long* out) {
  const int lane = lane_id();
  for (long long q = warp_global(); q < n_q; q += warps_total()) {
    const long long o = contig_off[contig[q]], len = contig_off[contig[q] + 1] - o;
    long long a = start[q], b = end[q];
    if (a < 0) a += len;
    if (b < 0) b += len;
    a = a < 0 ? 0 : (a > len ? len : a);
    b = b < 0 ? 0 : (b > len ? len : b);
    long long s = 0;
    if (b > a) {
      a += o; b += o;
      const long long ta = (a + SCAN_TILE - 1) / SCAN_TILE, tb = b / SCAN_TILE;   // whole tiles [ta, tb)
      if (ta <= tb) {
        for (long long i = a + lane; i < ta * SCAN_TILE; i += MFSC_LANES) s += cov[i];
        for (long long i = tb * SCAN_TILE + lane; i < b; i += MFSC_LANES) s += cov[i];
        s = wsum_i64(s) + (tile_pref[tb] - tile_pref[ta]);
      } else {
        for (long long i = a + lane; i < b; i += MFSC_LANES) s += cov[i];
        s = wsum_i64(s);
      }
    }
    if (lane == 0) out[q] = s;
  }
}
