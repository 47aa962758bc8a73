// mfsc_extend.cuh -- stage 4: postnuc-style extension of the clusters of one read into alignments
// (nucmer's last stage, `postnuc -b breaklen`; reference call site alignerRobot.py:46-64, breaklen
// 200, or 50 with houseKeeper.globalFast).
//
// One warp owns one read (both strands).  The control flow that walks clusters, picks merge
// targets and drops shadowed clusters is scalar and warp-uniform; the work is the banded
// anti-diagonal DP (nucleotide scores +3 / -7, affine gaps -7 open / -4 extend, three states per
// cell) whose band is spread over the lanes.  Per cell the DP also carries the number of
// alignment columns and errors of the path that produced each state, so no traceback matrix is
// written: the row arithmetic show-coords needs (S1 E1 S2 E2, errors, columns) falls out of the
// finish cell.  Band state lives in shared memory, 10 KB per warp, updated in place.
#pragma once
#include "mfsc_cluster.cuh"

#define DP_GOOD 3
#define DP_BAD (-7)
#define DP_GOPEN (-7)
#define DP_GCONT (-4)
#define DP_NEG (-(1 << 29))
#define DP_MAX_ALN 10000
#define DP_SLOTS 256            // circular band storage; band_cap must stay below DP_SLOTS - 2
#define DP_SMASK (DP_SLOTS - 1)
#define DP_SMEM_INTS (10 * DP_SLOTS)
#define AUX_GAP 0x10001         // +1 column, +1 error
#define AUX_COL 0x10000         // +1 column

struct DpOut { int reached, fi, fj, cols, errs; };

struct ExtStats { unsigned long long* cells; unsigned long long* calls; int* max_band; };

// View of one sequence: 1-based position i lives at padded position base + i - 1
struct SeqView { unsigned base; int n; };

// Best of three states with the tie rule of the reference DP: D only when strictly above both,
// I only when strictly above M and not below D.
MFSC_DEV void best3(int d, int i, int m, int ad, int ai, int am, int& v, int& a) {
  if (d > i) { if (d > m) { v = d; a = ad; } else { v = m; a = am; } }
  else if (i > m) { v = i; a = ai; } else { v = m; a = am; }
}

// Cell (i,j): i bases of A and j bases of B consumed, anti-diagonal d = i + j.  A base t (1..N) is
// A(a0 + dirn*(t-1)), likewise B.  forced: the best cell is re-chosen on every diagonal (the DP
// never gives up); optimal: reaching the far corner only counts if the best score is there.
MFSC_DEV DpOut dp_engine(const SeqStore& RS, SeqView Av, int a0, const SeqStore& QS, SeqView Bv, int b0, int dirn,
                         int N, int M, bool forced, bool optimal, const DevParams& P, int* sm, ExtStats st) {
  const int lane = lane_id();
  int* sD = sm; int* sI = sm + DP_SLOTS; int* sM = sm + 2 * DP_SLOTS;
  int* aD = sm + 3 * DP_SLOTS; int* aI = sm + 4 * DP_SLOTS; int* aM = sm + 5 * DP_SLOTS;
  int* bB[2] = {sm + 6 * DP_SLOTS, sm + 7 * DP_SLOTS};
  int* bA[2] = {sm + 8 * DP_SLOTS, sm + 9 * DP_SLOTS};
  const int max_diff = DP_GOOD * P.breaklen;
  wsync();
  if (lane == 0) {
    sD[0] = DP_NEG; sI[0] = DP_NEG; sM[0] = 0; aD[0] = 0; aI[0] = 0; aM[0] = 0;
    bB[0][0] = 0; bA[0][0] = 0;
  }
  wsync();
  int high = 0, FinD = 0, FinJ = 0, finAux = 0, lastAux = 0;
  int p1lo = 0, p1hi = 0, p2lo = 1, p2hi = 0;   // bands of diagonals d-1 and d-2
  int nlo = (1 - N) > 0 ? (1 - N) : 0, nhi = M < 1 ? M : 1;
  unsigned long long cells = 0;
  int maxw = 0;
  int d;
  for (d = 1; d <= N + M && (d - FinD) <= P.breaklen && nlo <= nhi; d++) {
    const int w = nhi - nlo + 1;
    cells += (unsigned long long)w;
    if (w > maxw) maxw = w;
    int* cB = bB[d & 1]; int* cA = bA[d & 1];     // holds diagonal d-2, receives diagonal d
    long long lkey = -(1ll << 62); int laux = 0;  // lane-local best of this diagonal (value, then larger j)
    for (int top = nhi; top >= nlo; top -= MFSC_LANES) {
      const int j = top - lane;
      const bool act = j >= nlo;
      const int i = d - j;
      int pD = DP_NEG, pI = DP_NEG, pM = DP_NEG, paD = 0, paI = 0, paM = 0;   // parent (i, j-1)
      int qD = DP_NEG, qI = DP_NEG, qM = DP_NEG, qaD = 0, qaI = 0, qaM = 0;   // parent (i-1, j)
      int rB = DP_NEG, rA = 0;                                                // parent (i-1, j-1): best state
      bool hasD = false, hasI = false, hasM = false;
      if (act) {
        if (j - 1 >= p1lo && j - 1 <= p1hi) {
          const int y = (j - 1) & DP_SMASK; hasD = true;
          pD = sD[y]; pI = sI[y]; pM = sM[y]; paD = aD[y]; paI = aI[y]; paM = aM[y];
        }
        if (j >= p1lo && j <= p1hi) {
          const int y = j & DP_SMASK; hasI = true;
          qD = sD[y]; qI = sI[y]; qM = sM[y]; qaD = aD[y]; qaI = aI[y]; qaM = aM[y];
        }
        if (j - 1 >= p2lo && j - 1 <= p2hi && i >= 1) {
          const int y = (j - 1) & DP_SMASK; hasM = true;
          rB = cB[y]; rA = cA[y];
        }
      }
      wsync();
      if (act) {
        int Dv = DP_NEG, Iv = DP_NEG, Mv = DP_NEG, xD = 0, xI = 0, xM = 0;
        if (hasD) {
          int del = pD > DP_NEG ? pD + DP_GCONT : DP_NEG;
          int ins = pI > DP_NEG ? pI + DP_GOPEN : DP_NEG;
          int mat = pM > DP_NEG ? pM + DP_GOPEN : DP_NEG;
          best3(del, ins, mat, paD, paI, paM, Dv, xD);
          xD += AUX_GAP;
          if (Dv <= DP_NEG) Dv = DP_NEG;
        }
        if (hasI) {
          int del = qD > DP_NEG ? qD + DP_GOPEN : DP_NEG;
          int ins = qI > DP_NEG ? qI + DP_GCONT : DP_NEG;
          int mat = qM > DP_NEG ? qM + DP_GOPEN : DP_NEG;
          best3(del, ins, mat, qaD, qaI, qaM, Iv, xI);
          xI += AUX_GAP;
          if (Iv <= DP_NEG) Iv = DP_NEG;
        }
        if (hasM && rB > DP_NEG) {
          const int ca = code_at(RS, Av.base + (unsigned)(a0 + dirn * (i - 1) - 1));
          const int cb = code_at(QS, Bv.base + (unsigned)(b0 + dirn * (j - 1) - 1));
          Mv = rB + ((ca <= 3 && ca == cb) ? DP_GOOD : DP_BAD);
          xM = rA + AUX_COL + (ca != cb ? 1 : 0);
        }
        const int y = j & DP_SMASK;
        sD[y] = Dv; sI[y] = Iv; sM[y] = Mv; aD[y] = xD; aI[y] = xI; aM[y] = xM;
        int bv, ba;
        best3(Dv, Iv, Mv, xD, xI, xM, bv, ba);
        cB[y] = bv; cA[y] = ba;
        const long long key = (long long)bv * 4294967296ll + (long long)j;
        if (key > lkey) { lkey = key; laux = ba; }
      }
      wsync();
    }
    // diagonal maximum, ties to the larger j
    const long long dkey = wmax_i64(lkey);
    const int daux = wmax_i32(lkey == dkey ? laux : -1);
    const int dmax = (int)((dkey - (dkey & 0xffffffffll)) / 4294967296ll);
    const int dj = (int)(dkey & 0xffffffffll);
    lastAux = daux;
    if (forced || dmax >= high) { high = dmax; FinD = d; FinJ = dj; finAux = daux; }
    // trim cells that fell more than max_diff below the best (narrows the next diagonal only)
    int lo = 0x7fffffff, hi = -0x7fffffff;
    for (int j = nlo + lane; j <= nhi; j += MFSC_LANES) {
      if (high - cB[j & DP_SMASK] <= max_diff) { if (j < lo) lo = j; if (j > hi) hi = j; }
    }
    lo = wmin_i32(lo); hi = wmax_i32(hi);
    if (lo > hi) { lo = nhi + 1; hi = nlo - 1; }
    while (lo <= hi && hi - lo + 2 > P.band_cap) { if (cB[lo & DP_SMASK] < cB[hi & DP_SMASK]) lo++; else hi--; }
    p2lo = p1lo; p2hi = p1hi; p1lo = nlo; p1hi = nhi;
    if (lo > hi) { nlo = 1; nhi = 0; }
    else {
      nlo = lo > d + 1 - N ? lo : d + 1 - N;
      int t = d + 1 < M ? d + 1 : M;
      nhi = hi + 1 < t ? hi + 1 : t;
    }
  }
  const int dlast = d - 1;
  DpOut o; o.reached = 0;
  if (dlast == N + M) {
    if (!optimal) { o.reached = 1; FinD = N + M; FinJ = M; finAux = lastAux; }
    else if (FinD == dlast) o.reached = 1;
  }
  o.fi = FinD - FinJ; o.fj = FinJ; o.cols = finAux >> 16; o.errs = finAux & 0xffff;
  if (lane == 0) {
    atomic_add_u64(st.cells, cells);
    atomic_add_u64(st.calls, 1ull);
    if (maxw > *st.max_band) atomic_max_i32(st.max_band, maxw);
  }
  return o;
}

// ------------------------------------------------------------------------------------------
// cluster walking
// ------------------------------------------------------------------------------------------
struct RowsOut {
  // alignments of a read are written into the read's slice (first match index of its forward strand)
  int* ref; int* qry; int* dir; int* sA; int* eA; int* sB; int* eB; int* errs; int* cols;
  int* n_rows;  // [n_reads]
};

struct ExtCtx {
  SeqStore RS, QS; DevParams P; int* sm; ExtStats st;
  SeqView A; SeqView B[2];
  RowsOut R; int al0, al1;       // alignments of the current (record, read) group: rows [al0, al1)
  const int* ord; int nC;        // clusters of the group in walking order -> split cluster id
  int pc_rc0;                    // split cluster ids >= pc_rc0 belong to the reverse strand
  const int* pc_first; const int* pc_n; unsigned char* fused;
  const unsigned* cm_s1; const int* cm_s2; const int* cm_len;
  unsigned a_ostart;             // separator-joined coordinate of A's first base
  int ref_id, qry_id, lane;
};

MFSC_DEV int c_dir(const ExtCtx& E, int c) { return E.ord[c] >= E.pc_rc0 ? 1 : 0; }
MFSC_DEV int c_nm(const ExtCtx& E, int c) { return E.pc_n[E.ord[c]]; }
MFSC_DEV int c_s1(const ExtCtx& E, int c, int t) { return (int)(E.cm_s1[E.pc_first[E.ord[c]] + t] - E.a_ostart) + 1; }
MFSC_DEV int c_s2(const ExtCtx& E, int c, int t) { return E.cm_s2[E.pc_first[E.ord[c]] + t] + 1; }
MFSC_DEV int c_len(const ExtCtx& E, int c, int t) { return E.cm_len[E.pc_first[E.ord[c]] + t]; }

MFSC_DEV bool ext_forward(ExtCtx& E, int ci, int tA, int tB, bool forced, bool optimal) {
  const int eA = E.R.eA[ci], eB = E.R.eB[ci], dir = E.R.dir[ci];
  bool overflow = false;
  if (tA - eA + 1 > DP_MAX_ALN) { tA = eA + DP_MAX_ALN - 1; overflow = true; optimal = true; }
  if (tB - eB + 1 > DP_MAX_ALN) { tB = eB + DP_MAX_ALN - 1; overflow = true; optimal = true; }
  DpOut o = dp_engine(E.RS, E.A, eA, E.QS, E.B[dir], eB, +1, tA - eA + 1, tB - eB + 1, forced, optimal, E.P, E.sm, E.st);
  // the DP re-aligns the alignment's last column
  wsync();
  if (E.lane == 0) {
    E.R.cols[ci] += o.cols - 1; E.R.errs[ci] += o.errs;
    E.R.eA[ci] = eA + o.fi - 1; E.R.eB[ci] = eB + o.fj - 1;
  }
  wsync();
  return o.reached && !overflow;
}

MFSC_DEV bool ext_backward(ExtCtx& E, int ci, int ti) {
  const int sA = E.R.sA[ci], sB = E.R.sB[ci], dir = E.R.dir[ci];
  int tA, tB; bool optimal = false, overflow = false;
  if (ti >= 0) { tA = E.R.eA[ti]; tB = E.R.eB[ti]; } else { tA = 1; tB = 1; optimal = true; }
  if (sA - tA + 1 > DP_MAX_ALN) { tA = sA - DP_MAX_ALN + 1; overflow = true; optimal = true; }
  if (sB - tB + 1 > DP_MAX_ALN) { tB = sB - DP_MAX_ALN + 1; overflow = true; optimal = true; }
  DpOut o = dp_engine(E.RS, E.A, sA, E.QS, E.B[dir], sB, -1, sA - tA + 1, sB - tB + 1, false, optimal, E.P, E.sm, E.st);
  bool reached = o.reached;
  if (overflow || ti < 0) reached = false;
  if (reached) {
    ext_forward(E, ti, sA, sB, true, false);
    // the target now ends on this alignment's first column; append the rest and drop this alignment
    if (E.lane == 0) {
      E.R.cols[ti] += E.R.cols[ci] - 1; E.R.errs[ti] += E.R.errs[ci];
      E.R.eA[ti] = E.R.eA[ci]; E.R.eB[ti] = E.R.eB[ci];
    }
    wsync();
    E.al1--;
  } else {
    const int nsA = sA - o.fi + 1, nsB = sB - o.fj + 1;
    DpOut o2 = dp_engine(E.RS, E.A, nsA, E.QS, E.B[dir], nsB, +1, sA - nsA + 1, sB - nsB + 1, true, false, E.P, E.sm, E.st);
    wsync();
    if (E.lane == 0) {
      E.R.cols[ci] += o2.cols - 1; E.R.errs[ci] += o2.errs;
      E.R.sA[ci] = nsA; E.R.sB[ci] = nsB;
    }
    wsync();
  }
  return reached;
}

MFSC_DEV bool ext_shadowed(const ExtCtx& E, int c, int ap) {
  if (E.al1 == E.al0) return false;
  const int nm = c_nm(E, c), dir = c_dir(E, c);
  const int sA = c_s1(E, c, 0), eA = c_s1(E, c, nm - 1) + c_len(E, c, nm - 1) - 1;
  const int sB = c_s2(E, c, 0), eB = c_s2(E, c, nm - 1) + c_len(E, c, nm - 1) - 1;
  for (int k = ap; k >= E.al0; k--)
    if (E.R.dir[k] == dir && E.R.eA[k] >= eA && E.R.eB[k] >= eB && E.R.sA[k] <= sA && E.R.sB[k] <= sB) return true;
  return false;
}

MFSC_DEV int ext_forward_target(const ExtCtx& E, int cur, int& tA, int& tB) {
  const int nm = c_nm(E, cur), dir = c_dir(E, cur);
  const int sA = c_s1(E, cur, nm - 1) + c_len(E, cur, nm - 1) - 1, sB = c_s2(E, cur, nm - 1) + c_len(E, cur, nm - 1) - 1;
  int dist = (tA - sA) < (tB - sB) ? (tA - sA) : (tB - sB);
  int res = -1;
  for (int k = cur + 1; k < E.nC; k++) {
    if (c_dir(E, k) != dir) continue;
    const int km = c_nm(E, k);
    int eA = c_s1(E, k, 0), eB = c_s2(E, k, 0);
    if ((eA < sA || eB < sB) && c_s1(E, k, km - 1) >= sA && c_s2(E, k, km - 1) >= sB) {
      for (int t = 0; t < km && (eA < sA || eB < sB); t++) { eA = c_s1(E, k, t); eB = c_s2(E, k, t); }
    }
    if (eA >= sA && eB >= sB) {
      int greater, lesser;
      if (eA - sA > eB - sB) { greater = eA - sA; lesser = eB - sB; } else { lesser = eA - sA; greater = eB - sB; }
      if (greater < E.P.breaklen || (long long)lesser * DP_GOOD + (long long)(greater - lesser) * DP_GCONT >= 0) { res = k; tA = eA; tB = eB; break; }
      else if (2ll * greater - lesser < dist) { res = k; tA = eA; tB = eB; dist = (int)(2ll * greater - lesser); }
    }
  }
  return res;
}

MFSC_DEV int ext_reverse_target(const ExtCtx& E, int cur) {
  const int sA = E.R.sA[cur], sB = E.R.sB[cur], dir = E.R.dir[cur];
  int dist = sA < sB ? sA : sB;
  int res = -1;
  for (int k = cur - 1; k >= E.al0; k--) {
    if (E.R.dir[k] != dir) continue;
    const int eA = E.R.eA[k], eB = E.R.eB[k];
    if (eA <= sA && eB <= sB) {
      int greater, lesser;
      if (sA - eA > sB - eB) { greater = sA - eA; lesser = sB - eB; } else { lesser = sA - eA; greater = sB - eB; }
      if (greater < E.P.breaklen || (long long)lesser * DP_GOOD + (long long)(greater - lesser) * DP_GCONT >= 0) { res = k; break; }
      else if (2ll * greater - lesser < dist) { res = k; dist = (int)(2ll * greater - lesser); }
    }
  }
  return res;
}

// Walk the clusters of one (record, read) group in order; returns nothing, alignments are rows [al0, al1)
MFSC_DEV void ext_clusters(ExtCtx& E) {
  bool target_reached = false;
  int tA = 0, tB = 0;
  int prev = 0, cur = 0, curA = -1, targetC = -1;
  while (cur < E.nC) {
    if (!target_reached) {
      if (E.fused[E.ord[cur]] || (E.P.simplify && ext_shadowed(E, cur, curA))) {
        if (E.lane == 0) E.fused[E.ord[cur]] = 1;
        wsync();
        cur = ++prev;
        continue;
      }
    }
    const int nm = c_nm(E, cur), dir = c_dir(E, cur);
    for (int mi = 0; mi < nm; mi++) {
      const int ms1 = c_s1(E, cur, mi), ms2 = c_s2(E, cur, mi), mlen = c_len(E, cur, mi);
      if (target_reached) {
        if (E.R.eA[curA] != ms1 || E.R.eB[curA] != ms2) continue;   // walk to the match the DP landed on
        wsync();
        if (E.lane == 0) { E.R.eA[curA] += mlen - 1; E.R.eB[curA] += mlen - 1; E.R.cols[curA] += mlen - 1; }
        wsync();
      } else {
        curA = E.al1++;
        wsync();
        if (E.lane == 0) {
          E.R.ref[curA] = E.ref_id; E.R.qry[curA] = E.qry_id; E.R.dir[curA] = dir;
          E.R.sA[curA] = ms1; E.R.sB[curA] = ms2; E.R.eA[curA] = ms1 + mlen - 1; E.R.eB[curA] = ms2 + mlen - 1;
          E.R.cols[curA] = mlen; E.R.errs[curA] = 0;
        }
        wsync();
        const int ti = ext_reverse_target(E, curA);
        if (ext_backward(E, curA, ti)) curA = ti;
      }
      if (mi < nm - 1) {
        tA = c_s1(E, cur, mi + 1); tB = c_s2(E, cur, mi + 1);
        target_reached = ext_forward(E, curA, tA, tB, false, false);
      } else {
        tA = E.A.n; tB = E.B[0].n;
        targetC = ext_forward_target(E, cur, tA, tB);
        target_reached = ext_forward(E, curA, tA, tB, false, targetC < 0);
      }
    }
    if (targetC < 0) target_reached = false;
    if (E.lane == 0) E.fused[E.ord[cur]] = 1;
    wsync();
    if (!target_reached) cur = ++prev; else cur = targetC;
  }
}

struct ExtIn {
  const int* seg;                       // [2*n_reads+1] match ranges of the strands
  const int* n_pc;                      // [2*n_reads]
  const int* pc_first; const int* pc_n; const int* pc_rec;
  const unsigned* cm_s1; const int* cm_s2; const int* cm_len;
  const unsigned* r_ostart;             // [n_ref]
  int* ord; int* key_tmp; unsigned char* fused;   // scratch (per-match slices)
};

#ifdef MFSC_EMU
static thread_local int emu_dp_smem[DP_SMEM_INTS];
#endif

MFSC_KERNEL k_extend(SeqStore RS, SeqStore QS, ExtIn I, int n_reads, DevParams P, RowsOut R, ExtStats st) {
#ifdef MFSC_EMU
  int* sm = emu_dp_smem;
#else
  extern __shared__ int dp_smem[];
  int* sm = dp_smem + warp_in_block() * DP_SMEM_INTS;
#endif
  const int lane = lane_id();
  for (long long q = warp_global(); q < n_reads; q += warps_total()) {
    const int b0 = I.seg[2 * q], b1 = I.seg[2 * q + 1];
    const int n0 = I.n_pc[2 * q], n1 = I.n_pc[2 * q + 1];
    const int nC = n0 + n1;
    if (nC == 0) { if (lane == 0) R.n_rows[q] = 0; continue; }
    int* ord = I.ord + b0;
    int* tmp = I.key_tmp + b0;
    // split cluster ids of the read: forward strand b0..b0+n0, reverse strand b1..b1+n1.
    // Walking order: record, start in the record, strand, start in the read, emission order.
    for (int x = lane; x < nC; x += MFSC_LANES) {
      const int pc = x < n0 ? b0 + x : b1 + (x - n0);
      tmp[x] = pc;
      I.fused[pc] = 0;
    }
    wsync();
    for (int x = lane; x < nC; x += MFSC_LANES) {
      const int pc = tmp[x];
      const int rec = I.pc_rec[pc]; const unsigned s1 = I.cm_s1[I.pc_first[pc]]; const int dir = pc >= b1 ? 1 : 0;
      const int s2 = I.cm_s2[I.pc_first[pc]];
      int rank = 0;
      for (int y = 0; y < nC; y++) {
        const int pd = tmp[y];
        const int rec2 = I.pc_rec[pd]; const unsigned t1 = I.cm_s1[I.pc_first[pd]]; const int dir2 = pd >= b1 ? 1 : 0;
        const int t2 = I.cm_s2[I.pc_first[pd]];
        bool before;
        if (rec2 != rec) before = rec2 < rec;
        else if (t1 != s1) before = t1 < s1;
        else if (dir2 != dir) before = dir2 < dir;
        else if (t2 != s2) before = t2 < s2;
        else before = y < x;
        rank += before ? 1 : 0;
      }
      ord[rank] = pc;
    }
    wsync();
    ExtCtx E;
    E.RS = RS; E.QS = QS; E.P = P; E.sm = sm; E.st = st; E.R = R; E.lane = lane;
    E.B[0].base = QS.start[2 * q]; E.B[0].n = QS.len[2 * q];
    E.B[1].base = QS.start[2 * q + 1]; E.B[1].n = QS.len[2 * q + 1];
    E.pc_rc0 = b1; E.pc_first = I.pc_first; E.pc_n = I.pc_n; E.fused = I.fused;
    E.cm_s1 = I.cm_s1; E.cm_s2 = I.cm_s2; E.cm_len = I.cm_len;
    E.qry_id = (int)q;
    int row0 = b0, rows = 0;
    int g0 = 0;
    while (g0 < nC) {
      const int rec = I.pc_rec[ord[g0]];
      int g1 = g0 + 1;
      while (g1 < nC && I.pc_rec[ord[g1]] == rec) g1++;
      E.ord = ord + g0; E.nC = g1 - g0;
      E.A.base = RS.start[rec]; E.A.n = RS.len[rec]; E.a_ostart = I.r_ostart[rec]; E.ref_id = rec;
      E.al0 = row0 + rows; E.al1 = E.al0;
      ext_clusters(E);
      rows = E.al1 - row0;
      g0 = g1;
    }
    wsync();
    // reverse-strand rows are reported on the forward read: position p -> qlen - p + 1
    const int qlen = QS.len[2 * q];
    for (int x = lane; x < rows; x += MFSC_LANES) {
      if (R.dir[row0 + x]) { R.sB[row0 + x] = qlen - R.sB[row0 + x] + 1; R.eB[row0 + x] = qlen - R.eB[row0 + x] + 1; }
    }
    if (lane == 0) R.n_rows[q] = rows;
  }
}

// rows of all reads, compacted in read order: dst offsets = exclusive scan of n_rows
MFSC_KERNEL k_compact_rows(RowsOut S, const int* seg, const int* row_off, int n_reads, RowsOut D) {
  const int lane = lane_id();
  for (long long q = warp_global(); q < n_reads; q += warps_total()) {
    const int src = seg[2 * q], dst = row_off[q], n = S.n_rows[q];
    for (int x = lane; x < n; x += MFSC_LANES) {
      D.ref[dst + x] = S.ref[src + x]; D.qry[dst + x] = S.qry[src + x]; D.dir[dst + x] = S.dir[src + x];
      D.sA[dst + x] = S.sA[src + x]; D.eA[dst + x] = S.eA[src + x]; D.sB[dst + x] = S.sB[src + x]; D.eB[dst + x] = S.eB[src + x];
      D.errs[dst + x] = S.errs[src + x]; D.cols[dst + x] = S.cols[src + x];
    }
  }
}
