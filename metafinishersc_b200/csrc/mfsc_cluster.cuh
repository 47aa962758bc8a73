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

MFSC_DEV long long ll_abs(long long x) { return x < 0 ? -x : x; }

MFSC_KERNEL k_cluster(const unsigned* m_s1, const int* m_s2, const int* m_len, const int* m_rec, const int* seg,
                      int n_qs, DevParams P, ClusterScratch W, ClusterOut O, unsigned* ticket) {
  const int lane = lane_id();
  // strands are handed out by a ticket counter, 8 at a time: the forward strands carry nearly all matches of a
  // reads-vs-contigs batch, so a static round-robin leaves every other warp without work
  for (;;) {
    unsigned q0 = 0;
    if (lane == 0) q0 = atomic_add_u32(ticket, 8u);
    q0 = (unsigned)wbcast_i32((int)q0, 0);
    if (q0 >= (unsigned)n_qs) break;
    const unsigned q1 = q0 + 8u < (unsigned)n_qs ? q0 + 8u : (unsigned)n_qs;
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
    int N = 0;
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
    // ---- 2. union-find ----
    for (int i = lane; i < N; i += MFSC_LANES) uf[i] = -1;
    wsync();
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
      int root = score[ord[c0]];
      int c1 = c0 + 1;
      while (c1 < N && score[ord[c1]] == root) c1++;
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
          // emit (lane 0), splitting at record changes; overlaps are trimmed off the later match
          int add_cm = 0, add_pc = 0;
          if (lane == 0) {
            bool firstm = true;
            int lastrec = -1;
            for (int i = 0; i < cn; i++) {
              const int a = mem[i];
              if (!flag[a]) continue;
              int ad = firstm ? 0 : adj[a];
              firstm = false;
              int l = len[a] - ad;
              if (l <= 0) continue;
              int w = b + n_cm + add_cm;
              O.s1[w] = s1[a] + (unsigned)ad; O.s2[w] = s2[a] + ad; O.len[w] = l;
              if (rec[a] != lastrec) {
                int pcw = b + n_pc + add_pc;
                O.pc_first[pcw] = w; O.pc_n[pcw] = 0; O.pc_rec[pcw] = rec[a];
                add_pc++;
                lastrec = rec[a];
              }
              O.pc_n[b + n_pc + add_pc - 1]++;
              add_cm++;
            }
          }
          n_cm += wbcast_i32(add_cm, 0);
          n_pc += wbcast_i32(add_pc, 0);
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
