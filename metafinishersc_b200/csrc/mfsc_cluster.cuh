// mfsc_cluster.cuh -- stage 3: mgaps-style clustering and chaining of the matches of one query strand
// (nucmer pipes `mummer | mgaps -l mincluster -s maxgap -d diagdiff -f diagfactor`; reference call
// site alignerRobot.py:46-64, defaults 65 / 90 / 5 / 0.12).
//
// One warp owns the sorted match list of one (read, strand).  Steps, all in the list's own slice
// of a few scratch arrays in HBM (a strand's list is short and stays in L1/L2):
//   1. overlap filter on identical diagonals / identical starts      (window scan over the lanes, hits in list order)
//   2. union-find over pairs closer than maxgap on near diagonals     (window scan over the lanes, unions by lane 0)
//   3. components ordered by their smallest member                    (counting sort, lane 0)
//   4. per component, repeatedly: chain DP whose inner "best predecessor" scan is spread over
//      the lanes and reduced with shuffles, trace back, emit the chain if it covers at least
//      mincluster bases, remove it.
// Emitted chains are split where consecutive matches lie in different reference records.
#pragma once
#include "mfsc_seed.cuh"

#define CL_WIDE 12   // widest window a single lane scans on its own
#define CL_K 4      // members per lane of a component whose chain DP runs as a register wavefront

struct ClusterScratch {
  // per match (slice = the strand's range in the sorted match arrays)
  unsigned* s1; int* s2; int* len; int* rec;  // filtered working copy
  int* score; int* adj; int* from; int* uf; int* ord; int* tmp;
  unsigned char* flag;                         // bit0 good, bit1 tentative
};

struct ClusterOut {
  unsigned* s1; int* s2; int* len;             // cluster matches, written into the strand's slice
  int* pc_first; int* pc_n; int* pc_rec;       // per split cluster: first match (absolute index), count, record
  int* n_pc;                                   // [n_qs] split clusters of the strand
  int* n_cm;                                   // [n_qs] cluster matches of the strand
};

MFSC_DEV int uf_find(int* uf, int a) {
  int r = a;
  while (uf[r] >= 0) r = uf[r];
  while (uf[a] >= 0) { int nx = uf[a]; uf[a] = r; a = nx; }
  return r;
}

// find while other lanes hook roots concurrently: parents only ever move to an ancestor (path halving), roots change
// only through the compare-and-swap of the caller
MFSC_DEV int uf_find_shared(int* uf, int a) {
  for (;;) {
    const int p = ld_volatile_i32(&uf[a]);
    if (p < 0) return a;
    const int g = ld_volatile_i32(&uf[p]);
    if (g < 0) return p;
    uf[a] = g;
    a = g;
  }
}

MFSC_DEV long long ll_abs(long long x) { return x < 0 ? -x : x; }

#ifndef CL_MINBLOCKS
#define CL_MINBLOCKS 10       // 51 registers; measured 3.49 ms (10 blocks) vs 3.64 ms (6 or 8) on cfg 1
#endif
MFSC_KERNEL_LB(128, CL_MINBLOCKS) k_cluster(const unsigned* m_s1, const int* m_s2, const int* m_len, const int* m_rec, const int* seg,
                      int n_qs, DevParams P, ClusterScratch W, ClusterOut O, unsigned* ticket, unsigned per_ticket) {
  const int lane = lane_id();
  // strands are handed out by a ticket counter, per_ticket at a time (8 on large batches, 1 when there are fewer strands
  // than resident warps): the forward strands carry nearly all matches of a reads-vs-contigs batch, so a static
  // round-robin leaves every other warp without work
  for (;;) {
    unsigned q0 = 0;
    if (lane == 0) q0 = atomic_add_u32(ticket, per_ticket);
    q0 = (unsigned)wbcast_i32((int)q0, 0);
    if (q0 >= (unsigned)n_qs) break;
    const unsigned q1 = q0 + per_ticket < (unsigned)n_qs ? q0 + per_ticket : (unsigned)n_qs;
  for (unsigned qs = q0; qs < q1; qs++) {
    const int b = seg[qs], e = seg[qs + 1];
    const int n = e - b;
    if (n <= 0) {
      if (lane == 0) { O.n_pc[qs] = 0; O.n_cm[qs] = 0; }
      continue;
    }
    unsigned* s1 = W.s1 + b; int* s2 = W.s2 + b; int* len = W.len + b; int* rec = W.rec + b;
    int* score = W.score + b; int* adj = W.adj + b; int* from = W.from + b; int* uf = W.uf + b;
    int* ord = W.ord + b; int* tmp = W.tmp + b;
    unsigned char* flag = W.flag + b;
    // The window scans of steps 1 and 2 (every match against the matches that start inside it / within maxgap
    // after it: hundreds of candidates per match when many reference records overlap the read, as in all-vs-all read
    // alignment) are spread over the lanes; the rare candidates that satisfy a condition are then handled one at a time
    // in list order with the sequential rule, so the result is that of the sequential algorithm.
    const unsigned all_lanes = MFSC_LANES == 32 ? 0xffffffffu : ((1u << (MFSC_LANES & 31)) - 1u);
    for (int i = lane; i < n; i += MFSC_LANES) { s1[i] = m_s1[b + i]; s2[i] = m_s2[b + i]; len[i] = m_len[b + i]; rec[i] = m_rec[b + i]; flag[i] = 1; }
    wsync();
    // ---- 1. filter ----
    // The filter only acts on pairs (i, j > i, s2[j] <= end of i) that share a diagonal, a reference start or a query start.
    // On reads against contigs there is usually no such pair at all: the lanes look for one, each over the windows of its
    // own matches, and the sequential rule below runs only when one exists.
    bool any_pair = false, wide = false;
    for (int i = lane; i < n - 1 && !wide; i += MFSC_LANES) {
      const long long i_s1 = s1[i], i_s2 = s2[i];
      const long long i_diag = i_s2 - i_s1, i_end = i_s2 + len[i];
      int j = i + 1;
      for (; j < n && (long long)s2[j] <= i_end + P.maxgap; j++) {
        if (j - i > CL_WIDE) { wide = true; break; }
        if ((long long)s2[j] > i_end) continue;
        const long long j_s1 = s1[j], j_s2 = s2[j];
        if ((j_s2 - j_s1 == i_diag) || j_s1 == i_s1 || j_s2 == i_s2) any_pair = true;
      }
    }
    // wide: some match has more than CL_WIDE later matches within maxgap of its end (many reference records overlap the
    // read, as in all-vs-all read alignment): the windows are then scanned by the whole warp, 32 matches at a time;
    // otherwise every lane walks the short windows of its own matches.
    wide = wballot(wide) != 0u;
    if (wide) any_pair = true;
    int N = n;
    if (wballot(any_pair) != 0u) {
    for (int i = 0; i < n - 1; i++) {
      unsigned fi = flag[i];
      if (!(fi & 1)) continue;
      const long long i_s1 = s1[i], i_s2 = s2[i];
      const long long i_diag = i_s2 - i_s1;
      int li = len[i];
      long long i_end = i_s2 + li;
      bool changed = false;
      int j0 = i + 1;
      while (j0 < n) {
        const int jj = j0 + lane;
        const bool in = jj < n && (long long)s2[jj] <= i_end;
        bool cand = false;
        if (in && (flag[jj] & 1)) {
          const long long j_s1 = s1[jj], j_s2 = s2[jj];
          cand = (j_s2 - j_s1 == i_diag) || j_s1 == i_s1 || j_s2 == i_s2;
        }
        const unsigned m = wballot(cand);
        if (m == 0u) {
          if (wballot(in) != all_lanes) break;       // the window ends inside this chunk
          j0 += MFSC_LANES;
          continue;
        }
        const int j = j0 + dev_ffs32(m) - 1;          // first candidate; everything before it is settled
        const long long j_s1 = s1[j], j_s2 = s2[j];
        const int lj = len[j];
        unsigned fj = flag[j];
        bool stop = false;
        if (j_s2 - j_s1 == i_diag) {
          const long long j_extent = (long long)lj + j_s2 - i_s2;
          if (j_extent > li) { li = (int)j_extent; i_end = i_s2 + j_extent; changed = true; }
          fj &= ~1u;
        } else {
          const long long olap = (i_s1 == j_s1) ? (i_s2 + li - j_s2) : (i_s1 + li - j_s1);
          if (li < lj) {
            if (olap >= li / 2) { fi &= ~1u; stop = true; }
          } else if (lj < li) {
            if (olap >= lj / 2) fj &= ~1u;
          } else {
            if (olap >= li / 2) {
              fj |= 2u;
              if (fi & 2u) { fi &= ~1u; stop = true; }
            }
          }
        }
        wsync();
        if (lane == 0) { flag[j] = (unsigned char)fj; flag[i] = (unsigned char)fi; if (changed) len[i] = li; }
        wsync();
        if (stop) break;
        j0 = j + 1;                                    // rescan behind the candidate: the window may have grown
      }
    }
    wsync();
    // compaction of the surviving matches, in order
    N = 0;
    for (int i0 = 0; i0 < n; i0 += MFSC_LANES) {
      const int i = i0 + lane;
      const bool keepm = i < n && (flag[i] & 1);
      unsigned v1 = 0; int v2 = 0, v3 = 0, v4 = 0;
      if (keepm) { v1 = s1[i]; v2 = s2[i]; v3 = len[i]; v4 = rec[i]; }
      const unsigned m = wballot(keepm);
      wsync();
      if (keepm) {
        const int w = N + dev_popc32(m & ((1u << lane) - 1u));
        s1[w] = v1; s2[w] = v2; len[w] = v3; rec[w] = v4;
      }
      N += dev_popc32(m);
      wsync();
    }
    }
    // ---- 2. union-find ----
    // Pairs closer than maxgap on near diagonals are united; a component's root is its smallest member, whatever the order
    // of the unions, so the lanes unite concurrently: each takes the windows of its own matches and hooks the larger root
    // under the smaller one with a compare-and-swap.
    for (int i = lane; i < N; i += MFSC_LANES) uf[i] = -1;
    wsync();
    if (!wide) {
      for (int i = lane; i < N - 1; i += MFSC_LANES) {
        const long long i_end = (long long)s2[i] + len[i], i_diag = (long long)s2[i] - (long long)s1[i];
        for (int j = i + 1; j < N; j++) {
          const long long sep = (long long)s2[j] - i_end;
          if (sep > P.maxgap) break;
          const long long dd = ll_abs(((long long)s2[j] - (long long)s1[j]) - i_diag);
          long long lim = (long long)(P.diagfactor * (double)sep);
          if (lim < P.diagdiff) lim = P.diagdiff;
          if (dd > lim) continue;
          int a = i, c = j;
          for (;;) {
            a = uf_find_shared(uf, a); c = uf_find_shared(uf, c);
            if (a == c) break;
            const int hi = a > c ? a : c, lo = a > c ? c : a;
            if (atomic_cas_i32(&uf[hi], -1, lo) == -1) break;
            a = hi; c = lo;                       // somebody hooked hi meanwhile: walk on from there
          }
        }
      }
    } else {
      for (int i = 0; i < N - 1; i++) {
        const long long i_end = (long long)s2[i] + len[i], i_diag = (long long)s2[i] - (long long)s1[i];
        for (int j0 = i + 1; j0 < N; j0 += MFSC_LANES) {
          const int jj = j0 + lane;
          bool in = false, hit = false;
          if (jj < N) {
            const long long sep = (long long)s2[jj] - i_end;
            in = sep <= P.maxgap;
            if (in) {
              const long long dd = ll_abs(((long long)s2[jj] - (long long)s1[jj]) - i_diag);
              long long lim = (long long)(P.diagfactor * (double)sep);
              if (lim < P.diagdiff) lim = P.diagdiff;
              hit = dd <= lim;
            }
          }
          unsigned m = wballot(hit);
          if (m) {
            if (lane == 0) {
              while (m) {
                const int j = j0 + dev_ffs32(m) - 1;
                m &= m - 1u;
                const int a = uf_find(uf, i), c = uf_find(uf, j);
                if (a != c) { if (a < c) uf[c] = a; else uf[a] = c; }
              }
            }
            wsync();
          }
          if (wballot(in) != all_lanes) break;
        }
      }
    }
    wsync();
    // ---- 3. members grouped by component, components by smallest member ----
    for (int i = lane; i < N; i += MFSC_LANES) {       // roots, read-only walk (no path compression while lanes share uf)
      int r = i;
      while (uf[r] >= 0) r = uf[r];
      from[i] = r; score[i] = r; tmp[i] = 0;           // from[] = root for now; score[] keeps the roots
    }
    wsync();
    // stable counting sort of the members by root, 32 members at a time: lanes with the same root find each other
    // with match.any, the lowest of them updates the root's counter, ranks inside the group follow lane order
    const unsigned lt_mask = MFSC_LANES == 32 ? ((1u << lane) - 1u) : 0u;
    for (int i0 = 0; i0 < N; i0 += MFSC_LANES) {
      const int i = i0 + lane;
      const int r = i < N ? from[i] : -1 - lane;
      const unsigned m = wmatch_any_i32(r);
      if (i < N && !(m & lt_mask)) tmp[r] += dev_popc32(m);
      wsync();
    }
    {
      int acc = 0;                                                             // tmp[root] = first slot: exclusive scan
      for (int i0 = 0; i0 < N; i0 += MFSC_LANES) {
        const int i = i0 + lane;
        const int c = i < N ? tmp[i] : 0;
        int incl = c;
#if MFSC_LANES == 32
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(MFSC_FULL, incl, o); if (lane >= o) incl += t; }
#endif
        if (i < N) tmp[i] = acc + incl - c;
        acc += wbcast_i32(incl, MFSC_LANES - 1);
      }
    }
    wsync();
    for (int i0 = 0; i0 < N; i0 += MFSC_LANES) {
      const int i = i0 + lane;
      const int r = i < N ? from[i] : -1 - lane;
      const unsigned m = wmatch_any_i32(r);
      int base = 0;
      if (i < N) base = tmp[r];
      wsync();
      if (i < N) {
        ord[base + dev_popc32(m & lt_mask)] = i;
        if (!(m & lt_mask)) tmp[r] = base + dev_popc32(m);
      }
      wsync();
    }
    // after this, component with root r occupies ord[tmp[r]-count .. tmp[r])
    wsync();
    // ---- 4. chains ----
    int n_pc = 0, n_cm = 0;  // warp-uniform
    int c0 = 0;
    while (c0 < N) {
      // component = maximal run in ord[] sharing a root (roots are smallest members, runs are ascending)
      const int root = score[ord[c0]];
      const int c1 = tmp[root];                 // the counting sort left the end slot of every component at its root
      wsync();
      // members of the component live in ord[c0..c1); work on indices into s1/s2/len
      int cn = c1 - c0;
      int* mem = ord + c0;
      // chain scores reuse score[] of the members: save nothing, roots are no longer needed for them
      while (cn > 0) {
        if (cn == 1) {
          int a = mem[0];
          if (len[a] >= P.mincluster) {
            if (lane == 0) {
              int w = b + n_cm;
              O.s1[w] = s1[a]; O.s2[w] = s2[a]; O.len[w] = len[a];
              O.pc_first[b + n_pc] = w; O.pc_n[b + n_pc] = 1; O.pc_rec[b + n_pc] = rec[a];
            }
            n_cm++; n_pc++;
          }
          cn = 0;
          break;
        }
        // Small component (at most CL_K members per lane) whose coordinates, taken relative to its first member, fit 28 bits:
        // the chain DP as a register wavefront in 32-bit arithmetic.  Lane l holds members l, l + 32, ...; the members are
        // finished in list order, member j's final score and its (end in the reference, end in the read, diagonal) are
        // broadcast from its lane with shuffles and every later member takes j as a candidate -- j ascending and a strict
        // comparison keep the first best, like the list scan.  (|cand| stays below 2^31: score and lengths < 2^28, overlap
        // < 2^29, diagonal difference < 2^30.)
        bool small = cn <= MFSC_LANES * CL_K;
        int m_s1[CL_K], m_s2[CL_K], m_len[CL_K], m_a[CL_K], bv[CL_K], bj[CL_K], badj[CL_K];
        if (small) {
          const long long base1 = s1[mem[0]], base2 = s2[mem[0]];
          bool fits = true;
#pragma unroll
          for (int t = 0; t < CL_K; t++) {
            const int i = lane + MFSC_LANES * t;
            m_a[t] = -1; m_s1[t] = 0; m_s2[t] = 0; m_len[t] = 0; bv[t] = 0; bj[t] = -1; badj[t] = 0;
            if (i < cn) {
              const int a = mem[i];
              const long long r1 = (long long)s1[a] - base1, r2 = (long long)s2[a] - base2;
              const int l = len[a];
              fits = fits && r1 > -(1 << 28) && r1 < (1 << 28) && r2 > -(1 << 28) && r2 < (1 << 28) && l < (1 << 28);
              m_a[t] = a; m_s1[t] = (int)r1; m_s2[t] = (int)r2; m_len[t] = l; bv[t] = l;
            }
          }
          small = wballot(!fits) == 0u;
        }
        if (small) {
#pragma unroll
          for (int tj = 0; tj < CL_K; tj++) {
            const int jbase = tj * MFSC_LANES;
            if (jbase < cn - 1) {
              const int jend = cn - 1 - jbase < MFSC_LANES ? cn - 1 - jbase : MFSC_LANES;
              const int my_e1 = m_s1[tj] + m_len[tj], my_e2 = m_s2[tj] + m_len[tj], my_dg = m_s2[tj] - m_s1[tj];
              for (int o = 0; o < jend; o++) {
                const int j = jbase + o;
                const int sc_j = wbcast_i32(bv[tj], o);
                const int j_e1 = wbcast_i32(my_e1, o), j_e2 = wbcast_i32(my_e2, o), j_dg = wbcast_i32(my_dg, o);
#pragma unroll
                for (int t = tj; t < CL_K; t++) {
                  const int i = lane + MFSC_LANES * t;
                  if (i > j && i < cn) {
                    const int olap1 = j_e1 - m_s1[t], olap2 = j_e2 - m_s2[t];
                    int olap = olap1 > 0 ? olap1 : 0;
                    olap = olap2 > olap ? olap2 : olap;
                    const int dd = (m_s2[t] - m_s1[t]) - j_dg;
                    const int cand = sc_j + m_len[t] - olap - (dd < 0 ? -dd : dd);
                    const int cand32 = cand < -0x40000000 ? -0x40000000 : cand;
                    if (cand32 > bv[t]) { bv[t] = cand32; bj[t] = j; badj[t] = olap; }
                  }
                }
              }
            }
          }
#pragma unroll
          for (int t = 0; t < CL_K; t++)
            if (m_a[t] >= 0) { const int a = m_a[t]; score[a] = bv[t]; from[a] = bj[t]; adj[a] = badj[t]; flag[a] = 0; }
          wsync();
        } else {
        // chain DP over the members in list order
        for (int i = 0; i < cn; i++) {
          const int a = mem[i];
          const long long a_s1 = s1[a], a_s2 = s2[a];
          const int a_len = len[a];
          int bestv = a_len;                    // candidates must beat the match alone
          int bestj = 0x7fffffff, bestadj = 0;
          for (int j = lane; j < i; j += MFSC_LANES) {
            const int c = mem[j];
            long long olap1 = (long long)s1[c] + len[c] - a_s1;
            long long olap = olap1 > 0 ? olap1 : 0;
            long long olap2 = (long long)s2[c] + len[c] - a_s2;
            if (olap2 > olap) olap = olap2;
            long long pen = olap + ll_abs((a_s2 - a_s1) - ((long long)s2[c] - (long long)s1[c]));
            long long cand = (long long)score[c] + a_len - pen;
            // a candidate only matters when it beats a_len > 0, so clamping it from below keeps the comparison exact
            const int cand32 = cand < -0x40000000ll ? -0x40000000 : (int)cand;
            if (cand32 > bestv) { bestv = cand32; bestj = j; bestadj = (int)olap; }   // first best within the lane
          }
          // first best across lanes: max value, then smallest j; the lane that holds it writes the result
          const int mv = wmax_i32(bestv);
          const int mj = wmin_i32(bestv == mv ? bestj : 0x7fffffff);
          if (mj != 0x7fffffff) {
            if (bestv == mv && bestj == mj) { score[a] = mv; from[a] = mj; adj[a] = bestadj; flag[a] = 0; }
          } else {
            if (lane == 0) { score[a] = a_len; from[a] = -1; adj[a] = 0; flag[a] = 0; }
          }
          wsync();
        }
        }
        // best chain end: first maximum
        long long key = -1;
        for (int i = lane; i < cn; i += MFSC_LANES) {
          long long kk = ((long long)score[mem[i]] << 32) | (unsigned)(0x7fffffff - i);
          if (kk > key) key = kk;
        }
        key = wmax_i64(key);
        int best = 0x7fffffff - (int)(key & 0xffffffffll);
        long long total = 0;
        if (lane == 0) {
          for (int i = best; i >= 0; i = from[mem[i]]) { flag[mem[i]] = 1; total += len[mem[i]]; }
        }
        wsync();
        total = wbcast_i64(total, 0);
        if (total >= P.mincluster) {
          // emit, splitting at record changes; overlaps are trimmed off the later match.  32 members at a time: the chain
          // members of a chunk find their output slots with ballots, a member whose record differs from the previous emitted
          // one opens a split cluster and counts its run inside the chunk, a run that goes on in the next chunk is added there.
          int add_cm = 0, add_pc = 0, lastrec = -1, open_pc = -1;
          bool seen_first = false;
          for (int i0 = 0; i0 < cn; i0 += MFSC_LANES) {
            const int i = i0 + lane;
            const int a = i < cn ? mem[i] : 0;
            const bool kept = i < cn && flag[a];
            const unsigned mk = wballot(kept);
            if (mk == 0u) continue;
            const int first_here = seen_first ? -1 : dev_ffs32(mk) - 1;     // the first member of the chain keeps its whole length
            seen_first = true;
            const int ad = kept ? (lane == first_here ? 0 : adj[a]) : 0;
            const int l = kept ? len[a] - ad : 0;
            const bool emit = kept && l > 0;
            const unsigned me = wballot(emit);
            if (me == 0u) continue;
            const int myrec = emit ? rec[a] : 0;
            const unsigned below = me & lt_mask;
            const int prevlane = below ? 31 - dev_clz32(below) : -1;
            const int prevrec_sh = wshfl_i32(myrec, prevlane < 0 ? 0 : prevlane);
            const int prevrec = prevlane >= 0 ? prevrec_sh : lastrec;
            const bool newpc = emit && myrec != prevrec;
            const unsigned mp = wballot(newpc);
            const int w = b + n_cm + add_cm + dev_popc32(me & lt_mask);
            if (emit) { O.s1[w] = s1[a] + (unsigned)ad; O.s2[w] = s2[a] + ad; O.len[w] = l; }
            if (newpc) {
              const int pcw = b + n_pc + add_pc + dev_popc32(mp & lt_mask);
              const unsigned above = mp & ~lt_mask & ~(1u << lane);
              const int end = above ? dev_ffs32(above) - 1 : MFSC_LANES;
              const unsigned upto = end >= 32 ? 0xffffffffu : ((1u << end) - 1u);
              O.pc_first[pcw] = w; O.pc_rec[pcw] = myrec; O.pc_n[pcw] = dev_popc32(me & upto & ~lt_mask);
            }
            const int firstp = mp ? dev_ffs32(mp) - 1 : MFSC_LANES;
            const int carry = dev_popc32(me & (firstp >= 32 ? 0xffffffffu : ((1u << firstp) - 1u)));
            if (carry && lane == 0) O.pc_n[open_pc] += carry;
            wsync();
            if (mp) open_pc = b + n_pc + add_pc + dev_popc32(mp) - 1;
            add_pc += dev_popc32(mp); add_cm += dev_popc32(me);
            lastrec = wbcast_i32(myrec, 31 - dev_clz32(me));
          }
          n_cm += add_cm;
          n_pc += add_pc;
        }
        // remove the chain: in-order compaction of the members left, 32 at a time
        int w = 0;
        for (int i0 = 0; i0 < cn; i0 += MFSC_LANES) {
          const int i = i0 + lane;
          int v = 0; bool keepm = false;
          if (i < cn) { v = mem[i]; keepm = !flag[v]; }
          const unsigned m = wballot(keepm);
          wsync();
          if (keepm) mem[w + dev_popc32(m & lt_mask)] = v;
          w += dev_popc32(m);
          wsync();
        }
        cn = w;
      }
      c0 = c1;
    }
    if (lane == 0) { O.n_pc[qs] = n_pc; O.n_cm[qs] = n_cm; }
  }
  }
}
