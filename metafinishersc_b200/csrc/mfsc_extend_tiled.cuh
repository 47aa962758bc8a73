// mfsc_extend_tiled.cuh -- stage 4 with several reads per warp.
//
// Same arithmetic as mfsc_extend.cuh (it reuses DP_CELL / DP_PASS and the band layout), different work
// decomposition.  In the one-read-per-warp kernel about 40 % of the ALU instructions are warp-uniform
// bookkeeping per anti-diagonal (maximum, finish cell, trim, next band) and a 97-cell band fills 25 of 32
// lanes.  Here a warp is split into XT tiles of XL lanes; every tile owns its own read, its own band state in
// shared memory and its own position in the cluster walk.  The warp runs one loop:
//
//     control : tiles that are not inside a DP advance their cluster walk (a small state machine that
//               replaces the recursion of ext_clusters / ext_backward / ext_forward) until they need a DP
//     diagonal: all tiles that are inside a DP compute ONE anti-diagonal together -- the bookkeeping
//               instructions are issued once for XT reads, the lanes of narrow bands are filled by the
//               other tiles
//
// A tile whose DP ends leaves to the control step while the others keep going, so tiles never wait for
// each other's DP calls.  Tile-level reductions use sub-warp masks.
#pragma once
#include "mfsc_extend.cuh"

#ifndef MFSC_EXT_TILES
#define MFSC_EXT_TILES 2
#endif
#if MFSC_LANES == 1
#define XT 1
#else
#define XT MFSC_EXT_TILES
#endif
#define XL (MFSC_LANES / XT)

#ifdef MFSC_EMU
static inline void tsync(unsigned) {}
static inline int tmax_i32(unsigned, int v) { return v; }
static inline int tmin_i32(unsigned, int v) { return v; }
static inline int tbcast_i32(unsigned, int v, int) { return v; }
static inline bool wany(bool p) { return p; }
#else
MFSC_DEV void tsync(unsigned m) { __syncwarp(m); }
MFSC_DEV int tmax_i32(unsigned m, int v) { return __reduce_max_sync(m, v); }
MFSC_DEV int tmin_i32(unsigned m, int v) { return __reduce_min_sync(m, v); }
MFSC_DEV int tbcast_i32(unsigned m, int v, int src_lane) { return __shfl_sync(m, v, src_lane); }
MFSC_DEV bool wany(bool p) { return __any_sync(MFSC_FULL, p); }
#endif

enum {
  XS_NEXT_READ = 0, XS_GROUP_BEGIN, XS_CLUSTER_BEGIN, XS_MATCH_BEGIN, XS_FWD_SETUP, XS_CLUSTER_END, XS_GROUP_END, XS_READ_END,
  XS_AFTER_BACK_SEARCH, XS_AFTER_BACK_MERGE_FWD, XS_AFTER_BACK_FORCED, XS_AFTER_FWD,
  XS_DP = 32, XS_DONE = 33
};

// per-tile cluster-walk state (shared memory; written by the tile's first lane, read by all after a tile sync)
struct TileCtl {
  int q, b0, b1, nC, row0, rows, g0, g1, qlen;
  int rec, a_n, al0, al1, gC;
  unsigned a_base, a_ostart, b_base[2];
  int cur, prev, curA, targetC, target_reached, mi, nm, dir;
  int ci, ti, after, overflow, sv_a, sv_b;      // pending DP: continuation, saved coordinates
  int pad[2];
};
#define XT_CTL_INTS 40
#define XT_TILE_INTS (DP_SMEM_INTS + XT_CTL_INTS)

struct TileDp {            // registers, uniform inside a tile
  int d, nlo, nhi, high, FinD, FinJ, finAux, lastAux, fillA, fillB, N, M, a0, b0, dirn, dir;
  bool forced, optimal;
  unsigned cells;
  // result of the last finished DP
  int o_reached, o_fi, o_fj, o_cols, o_errs;
};

struct TileEnv {
  SeqStore RS, QS; ExtIn I; RowsOut R; DevParams P;
  int tl; unsigned tmask; int* sm;     // lane in tile, tile mask, tile's shared memory
  TileCtl* C;
};

MFSC_DEV int x_ord(const TileEnv& V, int c) { return V.I.ord[V.C->b0 + V.C->g0 + c]; }
MFSC_DEV int x_dir(const TileEnv& V, int c) { return x_ord(V, c) >= V.C->b1 ? 1 : 0; }
MFSC_DEV int x_nm(const TileEnv& V, int c) { return V.I.pc_n[x_ord(V, c)]; }
MFSC_DEV int x_s1(const TileEnv& V, int c, int t) { return (int)(V.I.cm_s1[V.I.pc_first[x_ord(V, c)] + t] - V.C->a_ostart) + 1; }
MFSC_DEV int x_s2(const TileEnv& V, int c, int t) { return V.I.cm_s2[V.I.pc_first[x_ord(V, c)] + t] + 1; }
MFSC_DEV int x_len(const TileEnv& V, int c, int t) { return V.I.cm_len[V.I.pc_first[x_ord(V, c)] + t]; }

MFSC_DEV bool x_shadowed(const TileEnv& V, int c, int ap) {
  const TileCtl* C = V.C;
  if (C->al1 == C->al0) return false;
  const int nm = x_nm(V, c), dir = x_dir(V, c);
  const int sA = x_s1(V, c, 0), eA = x_s1(V, c, nm - 1) + x_len(V, c, nm - 1) - 1;
  const int sB = x_s2(V, c, 0), eB = x_s2(V, c, nm - 1) + x_len(V, c, nm - 1) - 1;
  for (int k = ap; k >= C->al0; k--)
    if (V.R.dir[k] == dir && V.R.eA[k] >= eA && V.R.eB[k] >= eB && V.R.sA[k] <= sA && V.R.sB[k] <= sB) return true;
  return false;
}

MFSC_DEV int x_forward_target(const TileEnv& V, int cur, int& tA, int& tB) {
  const int nm = x_nm(V, cur), dir = x_dir(V, cur);
  const int sA = x_s1(V, cur, nm - 1) + x_len(V, cur, nm - 1) - 1, sB = x_s2(V, cur, nm - 1) + x_len(V, cur, nm - 1) - 1;
  int dist = (tA - sA) < (tB - sB) ? (tA - sA) : (tB - sB);
  int res = -1;
  for (int k = cur + 1; k < V.C->gC; k++) {
    if (x_dir(V, k) != dir) continue;
    const int km = x_nm(V, k);
    int eA = x_s1(V, k, 0), eB = x_s2(V, k, 0);
    if ((eA < sA || eB < sB) && x_s1(V, k, km - 1) >= sA && x_s2(V, k, km - 1) >= sB) {
      for (int t = 0; t < km && (eA < sA || eB < sB); t++) { eA = x_s1(V, k, t); eB = x_s2(V, k, t); }
    }
    if (eA >= sA && eB >= sB) {
      int greater, lesser;
      if (eA - sA > eB - sB) { greater = eA - sA; lesser = eB - sB; } else { lesser = eA - sA; greater = eB - sB; }
      if (greater < V.P.breaklen || (long long)lesser * DP_GOOD + (long long)(greater - lesser) * DP_GCONT >= 0) { res = k; tA = eA; tB = eB; break; }
      else if (2ll * greater - lesser < dist) { res = k; tA = eA; tB = eB; dist = (int)(2ll * greater - lesser); }
    }
  }
  return res;
}

MFSC_DEV int x_reverse_target(const TileEnv& V, int cur) {
  const int sA = V.R.sA[cur], sB = V.R.sB[cur], dir = V.R.dir[cur];
  int dist = sA < sB ? sA : sB;
  int res = -1;
  for (int k = cur - 1; k >= V.C->al0; k--) {
    if (V.R.dir[k] != dir) continue;
    const int eA = V.R.eA[k], eB = V.R.eB[k];
    if (eA <= sA && eB <= sB) {
      int greater, lesser;
      if (sA - eA > sB - eB) { greater = sA - eA; lesser = sB - eB; } else { lesser = sA - eA; greater = sB - eB; }
      if (greater < V.P.breaklen || (long long)lesser * DP_GOOD + (long long)(greater - lesser) * DP_GCONT >= 0) { res = k; break; }
      else if (2ll * greater - lesser < dist) { res = k; dist = (int)(2ll * greater - lesser); }
    }
  }
  return res;
}

// start a DP for the tile: registers, diagonal 0 in shared memory, first window fill
MFSC_DEV void x_dp_begin(const TileEnv& V, TileDp& D, int dir, int a0, int b0, int dirn, int N, int M, bool forced, bool optimal) {
  int* sm = V.sm;
  int* nD = sm; int* naD = sm + DP_SLOTS; int* nI = sm + 2 * DP_SLOTS; int* naI = sm + 3 * DP_SLOTS; int* bBase = sm + 4 * DP_SLOTS;
  unsigned char* wA = (unsigned char*)(sm + 8 * DP_SLOTS);
  unsigned char* wB = wA + DP_WIN;
  D.dir = dir; D.a0 = a0; D.b0 = b0; D.dirn = dirn; D.N = N; D.M = M; D.forced = forced; D.optimal = optimal;
  D.d = 1; D.high = 0; D.FinD = 0; D.FinJ = 0; D.finAux = 0; D.lastAux = 0;
  D.nlo = (1 - N) > 0 ? (1 - N) : 0; D.nhi = M < 1 ? M : 1;
  tsync(V.tmask);
  if (V.tl == 0) {
    int* b0p = bBase; int* b1p = bBase + DP_SLOTS;
    nD[0] = DP_GOPEN; naD[0] = 0; nI[0] = DP_GOPEN; naI[0] = 0;
    nD[DP_SMASK] = DP_NEG; nI[1] = DP_NEG;
    b0p[0] = 0; b0p[2 * DP_SLOTS] = 0; b0p[DP_SMASK] = DP_NEG; b0p[1] = DP_NEG;
    b1p[DP_SMASK] = DP_NEG; b1p[0] = DP_NEG; b1p[1] = DP_NEG; b1p[2] = DP_NEG;
  }
  SeqView Av; Av.base = V.C->a_base; Av.n = V.C->a_n;
  SeqView Bv; Bv.base = V.C->b_base[dir]; Bv.n = V.C->qlen;
  {
    const int to = DP_FILL < N + 8 ? DP_FILL : N + 8;
    for (int t = V.tl; t < to; t += XL) { int c = 4; if (t < N) c = code_at(V.RS, Av.base + (unsigned)(a0 + dirn * t - 1)); wA[t & DP_WMASK] = (unsigned char)c; }
    const int tob = DP_FILL < M + 8 ? DP_FILL : M + 8;
    for (int t = V.tl; t < tob; t += XL) { int c = 4; if (t < M) c = code_at(V.QS, Bv.base + (unsigned)(b0 + dirn * t - 1)); wB[(t + 1) & DP_WMASK] = (unsigned char)c; }
  }
  D.fillA = DP_FILL; D.fillB = DP_FILL;
  tsync(V.tmask);
}

// forward extension request of alignment ci towards (tA, tB) (ext_forward of the one-read kernel)
MFSC_DEV void x_request_forward(const TileEnv& V, TileDp& D, int ci, int tA, int tB, bool forced, bool optimal, int after) {
  TileCtl* C = V.C;
  const int eA = V.R.eA[ci], eB = V.R.eB[ci], dir = V.R.dir[ci];
  bool overflow = false;
  if (tA - eA + 1 > DP_MAX_ALN) { tA = eA + DP_MAX_ALN - 1; overflow = true; optimal = true; }
  if (tB - eB + 1 > DP_MAX_ALN) { tB = eB + DP_MAX_ALN - 1; overflow = true; optimal = true; }
  tsync(V.tmask);      // every lane has read the walk state before it changes
  if (V.tl == 0) { C->ci = ci; C->after = after; C->overflow = overflow ? 1 : 0; C->sv_a = eA; C->sv_b = eB; }
  x_dp_begin(V, D, dir, eA, eB, +1, tA - eA + 1, tB - eB + 1, forced, optimal);
}

// result of a forward extension applied to alignment ci (the DP re-aligns the alignment's last column)
MFSC_DEV bool x_apply_forward(const TileEnv& V, const TileDp& D) {
  TileCtl* C = V.C;
  const int ci = C->ci, eA = C->sv_a, eB = C->sv_b;
  const bool ok = D.o_reached && !C->overflow;
  tsync(V.tmask);
  if (V.tl == 0) {
    V.R.cols[ci] += D.o_cols - 1; V.R.errs[ci] += D.o_errs;
    V.R.eA[ci] = eA + D.o_fi - 1; V.R.eB[ci] = eB + D.o_fj - 1;
  }
  tsync(V.tmask);
  return ok;
}

// Advance the tile's cluster walk until it needs a DP (returns XS_DP) or has no read left (XS_DONE).
MFSC_DEV int x_control(const TileEnv& V, TileDp& D, int st, int n_reads, unsigned* ticket, unsigned long long& calls) {
  TileCtl* C = V.C;
  const int tl = V.tl;
  const ExtIn& I = V.I;
  while (st != XS_DP && st != XS_DONE) {
    tsync(V.tmask);
    switch (st) {
      case XS_NEXT_READ: {
        int q = 0;
        if (tl == 0) q = (int)atomic_add_u32(ticket, 1u);
        q = tbcast_i32(V.tmask, q, (lane_id() / XL) * XL);
        if (q >= n_reads) { st = XS_DONE; break; }
        const int b0 = I.seg[2 * q], b1 = I.seg[2 * q + 1];
        const int n0 = I.n_pc[2 * q], n1 = I.n_pc[2 * q + 1];
        const int nC = n0 + n1;
        if (nC == 0) { if (tl == 0) V.R.n_rows[q] = 0; break; }
        int* ord = I.ord + b0; int* tmp = I.key_tmp + b0;
        for (int x = tl; x < nC; x += XL) { const int pc = x < n0 ? b0 + x : b1 + (x - n0); tmp[x] = pc; I.fused[pc] = 0; }
        tsync(V.tmask);
        for (int x = tl; x < nC; x += XL) {
          const int rec = I.pc_rec[tmp[x]];
          int f = x;
          for (int y = 0; y < x; y++) if (I.pc_rec[tmp[y]] == rec) { f = y; break; }
          I.grp[b0 + x] = f;
        }
        tsync(V.tmask);
        for (int x = tl; x < nC; x += XL) {
          const int pc = tmp[x];
          const int grp = I.grp[b0 + x]; const unsigned s1 = I.cm_s1[I.pc_first[pc]]; const int dir = pc >= b1 ? 1 : 0;
          const int s2 = I.cm_s2[I.pc_first[pc]];
          int rank = 0;
          for (int y = 0; y < nC; y++) {
            const int pd = tmp[y];
            const int grp2 = I.grp[b0 + y]; const unsigned t1 = I.cm_s1[I.pc_first[pd]]; const int dir2 = pd >= b1 ? 1 : 0;
            const int t2 = I.cm_s2[I.pc_first[pd]];
            bool before;
            if (grp2 != grp) before = grp2 < grp;
            else if (t1 != s1) before = t1 < s1;
            else if (dir2 != dir) before = dir2 < dir;
            else if (t2 != s2) before = t2 < s2;
            else before = y < x;
            rank += before ? 1 : 0;
          }
          ord[rank] = pc;
        }
        if (tl == 0) {
          C->q = q; C->b0 = b0; C->b1 = b1; C->nC = nC; C->row0 = b0; C->rows = 0; C->g0 = 0; C->qlen = V.QS.len[2 * q];
          C->b_base[0] = V.QS.start[2 * q]; C->b_base[1] = V.QS.start[2 * q + 1];
        }
        st = XS_GROUP_BEGIN;
        break;
      }
      case XS_GROUP_BEGIN: {
        const int g0 = C->g0, nC = C->nC;
        if (g0 >= nC) { st = XS_READ_END; break; }
        const int* ord = I.ord + C->b0;
        const int rec = I.pc_rec[ord[g0]];
        int g1 = g0 + 1;
        while (g1 < nC && I.pc_rec[ord[g1]] == rec) g1++;
        if (tl == 0) {
          C->g1 = g1; C->gC = g1 - g0; C->rec = rec; C->a_base = V.RS.start[rec]; C->a_n = V.RS.len[rec]; C->a_ostart = I.r_ostart[rec];
          C->al0 = C->row0 + C->rows; C->al1 = C->row0 + C->rows;
          C->cur = 0; C->prev = 0; C->curA = -1; C->targetC = -1; C->target_reached = 0;
        }
        st = XS_CLUSTER_BEGIN;
        break;
      }
      case XS_CLUSTER_BEGIN: {
        const int cur = C->cur;
        if (cur >= C->gC) { st = XS_GROUP_END; break; }
        if (!C->target_reached) {
          const int pc = x_ord(V, cur);
          if (I.fused[pc] || (V.P.simplify && x_shadowed(V, cur, C->curA))) {
            tsync(V.tmask);
            if (tl == 0) { I.fused[pc] = 1; C->prev = C->prev + 1; C->cur = C->prev; }
            break;
          }
        }
        const int nm = x_nm(V, cur), dir = x_dir(V, cur);
        if (tl == 0) { C->nm = nm; C->dir = dir; C->mi = 0; }
        st = XS_MATCH_BEGIN;
        break;
      }
      case XS_MATCH_BEGIN: {
        const int mi = C->mi, nm = C->nm, cur = C->cur;
        if (mi >= nm) { st = XS_CLUSTER_END; break; }
        const int ms1 = x_s1(V, cur, mi), ms2 = x_s2(V, cur, mi), mlen = x_len(V, cur, mi);
        if (C->target_reached) {
          const int curA = C->curA;
          if (V.R.eA[curA] != ms1 || V.R.eB[curA] != ms2) {        // walk to the match the DP landed on
            tsync(V.tmask);
            if (tl == 0) C->mi = mi + 1;
            break;
          }
          tsync(V.tmask);
          if (tl == 0) { V.R.eA[curA] += mlen - 1; V.R.eB[curA] += mlen - 1; V.R.cols[curA] += mlen - 1; }
          st = XS_FWD_SETUP;
          break;
        }
        const int curA = C->al1, dir = C->dir;
        tsync(V.tmask);
        if (tl == 0) {
          C->al1 = curA + 1; C->curA = curA;
          V.R.ref[curA] = C->rec; V.R.qry[curA] = C->q; V.R.dir[curA] = dir;
          V.R.sA[curA] = ms1; V.R.sB[curA] = ms2; V.R.eA[curA] = ms1 + mlen - 1; V.R.eB[curA] = ms2 + mlen - 1;
          V.R.cols[curA] = mlen; V.R.errs[curA] = 0;
        }
        tsync(V.tmask);
        const int ti = x_reverse_target(V, curA);
        // backward search from the alignment's first column towards the target alignment or the sequence start
        int tA, tB; bool optimal = false, overflow = false;
        if (ti >= 0) { tA = V.R.eA[ti]; tB = V.R.eB[ti]; } else { tA = 1; tB = 1; optimal = true; }
        if (ms1 - tA + 1 > DP_MAX_ALN) { tA = ms1 - DP_MAX_ALN + 1; overflow = true; optimal = true; }
        if (ms2 - tB + 1 > DP_MAX_ALN) { tB = ms2 - DP_MAX_ALN + 1; overflow = true; optimal = true; }
        if (tl == 0) { C->ci = curA; C->ti = ti; C->after = XS_AFTER_BACK_SEARCH; C->overflow = overflow ? 1 : 0; C->sv_a = ms1; C->sv_b = ms2; }
        x_dp_begin(V, D, dir, ms1, ms2, -1, ms1 - tA + 1, ms2 - tB + 1, false, optimal);
        calls++;
        st = XS_DP;
        break;
      }
      case XS_AFTER_BACK_SEARCH: {
        const int ci = C->ci, ti = C->ti, sA = C->sv_a, sB = C->sv_b;
        bool reached = D.o_reached != 0;
        if (C->overflow || ti < 0) reached = false;
        if (reached) {
          x_request_forward(V, D, ti, sA, sB, true, false, XS_AFTER_BACK_MERGE_FWD);
          if (tl == 0) C->ti = ci;                  // remember the alignment to be absorbed (ci of the request is the target)
        } else {
          const int nsA = sA - D.o_fi + 1, nsB = sB - D.o_fj + 1;
          const int dir = V.R.dir[ci];
          tsync(V.tmask);
          if (tl == 0) { C->after = XS_AFTER_BACK_FORCED; C->sv_a = nsA; C->sv_b = nsB; }
          x_dp_begin(V, D, dir, nsA, nsB, +1, sA - nsA + 1, sB - nsB + 1, true, false);
        }
        calls++;
        st = XS_DP;
        break;
      }
      case XS_AFTER_BACK_MERGE_FWD: {
        x_apply_forward(V, D);
        const int t = C->ci, c = C->ti;           // target alignment, absorbed alignment
        // the target now ends on the absorbed alignment's first column; append the rest and drop it
        if (tl == 0) {
          V.R.cols[t] += V.R.cols[c] - 1; V.R.errs[t] += V.R.errs[c];
          V.R.eA[t] = V.R.eA[c]; V.R.eB[t] = V.R.eB[c];
          C->al1 = C->al1 - 1; C->curA = t;
        }
        st = XS_FWD_SETUP;
        break;
      }
      case XS_AFTER_BACK_FORCED: {
        const int ci = C->ci;
        if (tl == 0) {
          V.R.cols[ci] += D.o_cols - 1; V.R.errs[ci] += D.o_errs;
          V.R.sA[ci] = C->sv_a; V.R.sB[ci] = C->sv_b;
        }
        st = XS_FWD_SETUP;
        break;
      }
      case XS_FWD_SETUP: {
        const int mi = C->mi, nm = C->nm, cur = C->cur, curA = C->curA;
        int tA, tB; bool fopt = false;
        if (mi < nm - 1) { tA = x_s1(V, cur, mi + 1); tB = x_s2(V, cur, mi + 1); }
        else {
          tA = C->a_n; tB = C->qlen;
          const int targetC = x_forward_target(V, cur, tA, tB);
          tsync(V.tmask);
          if (tl == 0) C->targetC = targetC;
          fopt = targetC < 0;
        }
        x_request_forward(V, D, curA, tA, tB, false, fopt, XS_AFTER_FWD);
        calls++;
        st = XS_DP;
        break;
      }
      case XS_AFTER_FWD: {
        const bool ok = x_apply_forward(V, D);
        if (tl == 0) { C->target_reached = ok ? 1 : 0; C->mi = C->mi + 1; }
        st = XS_MATCH_BEGIN;
        break;
      }
      case XS_CLUSTER_END: {
        const int cur = C->cur;
        int tr = C->target_reached;
        if (C->targetC < 0) tr = 0;
        const int pc = x_ord(V, cur);
        tsync(V.tmask);
        if (tl == 0) {
          C->target_reached = tr;
          I.fused[pc] = 1;
          if (!tr) { C->prev = C->prev + 1; C->cur = C->prev; } else C->cur = C->targetC;
        }
        st = XS_CLUSTER_BEGIN;
        break;
      }
      case XS_GROUP_END: {
        if (tl == 0) { C->rows = C->al1 - C->row0; C->g0 = C->g1; }
        st = XS_GROUP_BEGIN;
        break;
      }
      case XS_READ_END: {
        // reverse-strand rows are reported on the forward read: position p -> qlen - p + 1
        const int qlen = C->qlen, row0 = C->row0, rows = C->rows;
        for (int x = tl; x < rows; x += XL)
          if (V.R.dir[row0 + x]) { V.R.sB[row0 + x] = qlen - V.R.sB[row0 + x] + 1; V.R.eB[row0 + x] = qlen - V.R.eB[row0 + x] + 1; }
        if (tl == 0) V.R.n_rows[C->q] = rows;
        st = XS_NEXT_READ;
        break;
      }
      default: st = XS_DONE; break;
    }
  }
  tsync(V.tmask);
  return st;
}

#ifdef MFSC_EMU
static thread_local int emu_xt_smem[XT_TILE_INTS];
#endif

#ifndef MFSC_XT_MINBLOCKS
#define MFSC_XT_MINBLOCKS 3
#endif

MFSC_KERNEL_LB(128, MFSC_XT_MINBLOCKS) k_extend_tiled(SeqStore RS, SeqStore QS, ExtIn I, int n_reads, DevParams P, RowsOut R, ExtStats stt,
                                                       unsigned* ticket) {
#ifdef MFSC_EMU
  int* sm = emu_xt_smem;
  const int tile = 0;
#else
  extern __shared__ int4 dp_smem4[];
  const int tile = lane_id() / XL;
  int* sm = reinterpret_cast<int*>(dp_smem4) + (warp_in_block() * XT + tile) * XT_TILE_INTS;
#endif
  const int tl = lane_id() % XL;
  const int lane = tl;                 // DP_PASS spreads the groups of a band over `lane`: here the lane inside the tile
  TileEnv V;
  V.RS = RS; V.QS = QS; V.I = I; V.R = R; V.P = P; V.tl = tl; V.sm = sm;
#ifdef MFSC_EMU
  V.tmask = 1u;
#else
  V.tmask = (XL == 32) ? 0xffffffffu : (((1u << XL) - 1u) << (tile * XL));
#endif
  V.C = reinterpret_cast<TileCtl*>(sm + DP_SMEM_INTS);
  (void)tile;
  int* nD = sm; int* naD = sm + DP_SLOTS; int* nI = sm + 2 * DP_SLOTS; int* naI = sm + 3 * DP_SLOTS; int* bBase = sm + 4 * DP_SLOTS;
  unsigned char* wA = (unsigned char*)(sm + 8 * DP_SLOTS);
  unsigned char* wB = wA + DP_WIN;
  const unsigned* wA32 = (const unsigned*)wA; const unsigned* wB32 = (const unsigned*)wB;
  const int breaklen = P.breaklen, band_cap = P.band_cap, one = P.one;
  const int max_diff = DP_GOOD * breaklen;
  const int k256 = one << 8;
  TileDp D;
  D.cells = 0; D.o_reached = 0; D.o_fi = 0; D.o_fj = 0; D.o_cols = 0; D.o_errs = 0;
  D.d = 1; D.nlo = 1; D.nhi = 0; D.high = 0; D.FinD = 0; D.FinJ = 0; D.finAux = 0; D.lastAux = 0; D.fillA = 0; D.fillB = 0;
  D.N = 0; D.M = 0; D.a0 = 0; D.b0 = 0; D.dirn = 1; D.dir = 0; D.forced = false; D.optimal = false;
  unsigned long long t_cells = 0, t_calls = 0;
  int st = XS_NEXT_READ;
  for (;;) {
    if (st != XS_DP && st != XS_DONE) st = x_control(V, D, st, n_reads, ticket, t_calls);
    wsync();
    const bool in_dp = st == XS_DP;
    if (!wany(in_dp)) break;
    // ---------------- one anti-diagonal for every tile that is inside a DP ----------------
    const int d = D.d, N = D.N, M = D.M;
    const int nlo = in_dp ? D.nlo : 1, nhi = in_dp ? D.nhi : 0;
    const int w = nhi - nlo + 1;
    if (in_dp) D.cells += (unsigned)w;
    // window refill (tile-local, rare)
    if (in_dp && (d - nlo > D.fillA || nhi > D.fillB)) {
      SeqView Av; Av.base = V.C->a_base; Av.n = V.C->a_n;
      SeqView Bv; Bv.base = V.C->b_base[D.dir]; Bv.n = V.C->qlen;
      if (d - nlo > D.fillA) {
        const int from = D.fillA, to = from + DP_FILL < N + 8 ? from + DP_FILL : N + 8;
        for (int t = from + tl; t < to; t += XL) { int c = 4; if (t < N) c = code_at(RS, Av.base + (unsigned)(D.a0 + D.dirn * t - 1)); wA[t & DP_WMASK] = (unsigned char)c; }
        D.fillA += DP_FILL;
      }
      if (nhi > D.fillB) {
        const int from = D.fillB, to = from + DP_FILL < M + 8 ? from + DP_FILL : M + 8;
        for (int t = from + tl; t < to; t += XL) { int c = 4; if (t < M) c = code_at(QS, Bv.base + (unsigned)(D.b0 + D.dirn * t - 1)); wB[(t + 1) & DP_WMASK] = (unsigned char)c; }
        D.fillB += DP_FILL;
      }
    }
    wsync();
    int* cB = bBase + (d & 1) * DP_SLOTS; int* cA = cB + 2 * DP_SLOTS;
    int lkey = -0x7fffffff;
    const int glo = in_dp ? (nlo >> 2) : (1 << 20), ghi = in_dp ? (nhi >> 2) : 0;   // tiles outside a DP own no group
    const int npass = in_dp ? (ghi - glo) / XL + 1 : 0;
    const int wpass = wmax_i32(npass);
    for (int p = 0; p < wpass; p++) {
      int e0, e1, e2, e3, ej;
      DP_PASS(ghi - p * XL, e0, e1, e2, e3, ej)
      (void)e0; (void)e1; (void)e2; (void)e3; (void)ej;
    }
    // diagonal maximum (ties to the larger j), finish cell
    const int dkey = tmax_i32(V.tmask, lkey);
    const int dmax = dkey >> 8;
    const int dj = nlo + (dkey & 0xff);
    const int daux = cA[dj & DP_SMASK];
    const bool upd = in_dp && (D.forced || dmax >= D.high);
    if (in_dp) D.lastAux = daux;
    if (upd) { D.high = dmax; D.FinD = d; D.FinJ = dj; D.finAux = daux; }
    // halo
    if (in_dp) {
#if MFSC_LANES == 1
      nD[(nlo - 1) & DP_SMASK] = DP_NEG; nI[(nhi + 1) & DP_SMASK] = DP_NEG; cB[(nlo - 1) & DP_SMASK] = DP_NEG; cB[(nhi + 1) & DP_SMASK] = DP_NEG;
#else
      const int hidx = ((tl & 1) ? (nhi + 1) : (nlo - 1)) & DP_SMASK;
      int* hb = (tl & 2) ? cB : ((tl & 1) ? nI : nD);
      if (tl < 4) hb[hidx] = DP_NEG;
#endif
    }
    wsync();
    // trim
    int lo = 0x7fffffff, hi = -0x7fffffff;
    const int thr = D.high - max_diff;
    for (int p = 0; p < wpass; p++) {
      const int g = glo + p * XL + tl;
      if (in_dp && g <= ghi) {
        const int4 v = ld4(cB + ((g * 4) & DP_SMASK));
        const int j0 = g * 4;
        const unsigned m = (v.x >= thr ? 1u : 0u) | (v.y >= thr ? 2u : 0u) | (v.z >= thr ? 4u : 0u) | (v.w >= thr ? 8u : 0u);
        if (m) {
          const int a = j0 + dev_ffs32(m) - 1, z = j0 + 31 - dev_clz32(m);
          if (a < lo) lo = a;
          if (z > hi) hi = z;
        }
      }
    }
    lo = tmin_i32(V.tmask, lo); hi = tmax_i32(V.tmask, hi);
    if (in_dp) {
      const bool none = lo > hi;
      if (!none && hi - lo + 2 > band_cap) {
        while (lo <= hi && hi - lo + 2 > band_cap) { if (cB[lo & DP_SMASK] < cB[hi & DP_SMASK]) lo++; else hi--; }
      }
      const int a = d + 1 - N, t = d + 1 < M ? d + 1 : M;
      D.nlo = none ? 1 : (lo > a ? lo : a);
      D.nhi = none ? 0 : (hi + 1 < t ? hi + 1 : t);
      D.d = d + 1;
      // DP over?
      if (!(D.d <= N + M && (D.d - D.FinD) <= breaklen && D.nlo <= D.nhi)) {
        const int dlast = D.d - 1;
        int reached = 0;
        if (dlast == N + M) {
          if (!D.optimal) { reached = 1; D.FinD = N + M; D.FinJ = M; D.finAux = D.lastAux; }
          else if (D.FinD == dlast) reached = 1;
        }
        D.o_reached = reached; D.o_fi = D.FinD - D.FinJ; D.o_fj = D.FinJ;
        D.o_cols = D.FinD - (D.finAux >> 16); D.o_errs = (D.finAux & 0xffff) + D.FinD - 2 * (D.finAux >> 16);
        t_cells += D.cells; D.cells = 0;
        st = V.C->after;
      }
    }
  }
  if (tl == 0 && t_calls) {
    atomic_add_u64(stt.cells, t_cells);
    atomic_add_u64(stt.calls, t_calls);
  }
}
