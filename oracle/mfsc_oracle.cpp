// oracle/mfsc_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY -- CPU restatement of the hot path, used as the
// parity checker by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs.  The product (metafinishersc_b200/)
// must never import, link or execute anything in this directory.
//
// PARITY UNPINNED against real MUMmer: the arithmetic of the path lives in
// MUMmer 3.23 (nucmer = prenuc | mummer | mgaps | postnuc, and show-coords),
// an external download that is NOT under /root/reference (reference
// README.md:41-45, .travis.yml:11-13, reproduce.py:20-22) and is not on this
// machine.  This file restates MUMmer 3.23's published algorithm (Kurtz et al.
// 2004, "Versatile and open software for comparing large genomes"; the
// MUMmer 3 manual) from its documented behaviour:
//   mummer -maxmatch -b -l L -n   : all maximal exact matches >= L, both strands
//   mgaps  -l C -s G -d D -f F    : union-find clustering + chain DP
//   postnuc -b B                  : cluster extension, anti-diagonal banded DP
//   show-coords -r                : row arithmetic (LEN, %IDY)
// It is anchored on the reference's own call sites
// (srcRefactor/repeatPhaserLib/finisherSCCoreLib/alignerRobot.py:46-72,
//  srcRefactor/misassemblyFixerLib/adaptorFix.py:21-24) and on the
// structure-determined show-coords tables in the reference README.md:111-129,
// which tests/test_oracle_golden.py reproduces.
//
// Everything here is single-threaded per call and deterministic.

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

typedef int64_t i64;

struct Params {
  int minmatch;      // nucmer -l   (alignerRobot.py:46-64: 20, or 10 with refinedVersion)
  int mincluster;    // nucmer -c   65
  int maxgap;        // nucmer -g   90
  int breaklen;      // nucmer -b   200 (50 with houseKeeper.globalFast)
  int diagdiff;      // nucmer -D   5
  double diagfactor; // nucmer -d   0.12
  int maxmatch;      // 1: --maxmatch, 0: default -mumreference (adaptorFix.py:21)
  int simplify;      // 1: default (--simplify), shadowed clusters are dropped
  int band_cap;      // safety cap on the anti-diagonal band (cells); never reached in practice
};

const int GOOD = 3, BAD = -7, GOPEN = -7, GCONT = -4;  // nucleotide scoring of the DP
const i64 MAX_ALN = 10000;                             // longest single DP extension
const int NEG = -(1 << 29);

inline int code_of(char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;  // any other character never seeds a match and scores as a mismatch
  }
}

struct SeqSet {
  std::vector<uint8_t> t;      // codes; records separated by one code-4 separator (prenuc-style concatenation)
  std::vector<i64> start;      // start[k] = global position of record k's first base
  std::vector<i64> len;        // len[k]
  int n() const { return (int)len.size(); }
  int rec_of(i64 g) const {
    int k = (int)(std::upper_bound(start.begin(), start.end(), g) - start.begin()) - 1;
    return k;
  }
};

void load(SeqSet& s, const char* bases, const i64* off, int n) {
  s.start.resize(n); s.len.resize(n);
  i64 total = 0;
  for (int k = 0; k < n; k++) total += off[k + 1] - off[k] + 1;
  s.t.assign(total + 1, 4);
  i64 g = 0;
  for (int k = 0; k < n; k++) {
    s.start[k] = g; s.len[k] = off[k + 1] - off[k];
    for (i64 p = off[k]; p < off[k + 1]; p++) s.t[g++] = (uint8_t)code_of(bases[p]);
    s.t[g++] = 4;
  }
}

// ----------------------------------------------------------------------------
// Stage 1+2: maximal exact matches (mummer -maxmatch / -mumreference, -b, -n)
// ----------------------------------------------------------------------------
struct Mem { int q, strand; i64 s1, s2, len; };  // s1: global ref pos (0-based), s2: 0-based on the strand

struct KIndex {
  int k; std::vector<uint32_t> head; std::vector<uint32_t> pos;
};

void build_index(const SeqSet& R, int k, KIndex& ix) {
  ix.k = k;
  size_t nb = (size_t)1 << (2 * k);
  ix.head.assign(nb + 1, 0);
  i64 n = (i64)R.t.size();
  uint64_t mask = nb - 1, km = 0; int valid = 0;
  for (int pass = 0; pass < 2; pass++) {
    km = 0; valid = 0;
    for (i64 p = 0; p < n; p++) {
      int c = R.t[p];
      if (c > 3) { valid = 0; km = 0; continue; }
      km = ((km << 2) | (uint64_t)c) & mask; valid++;
      if (valid >= k) {
        if (pass == 0) ix.head[km + 1]++;
        else ix.pos[ix.head[km]++] = (uint32_t)(p - k + 1);
      }
    }
    if (pass == 0) {
      for (size_t b = 0; b < nb; b++) ix.head[b + 1] += ix.head[b];
      ix.pos.resize(ix.head[nb]);
    } else {
      for (size_t b = nb; b > 0; b--) ix.head[b] = ix.head[b - 1];
      ix.head[0] = 0;
    }
  }
}

// all MEMs of one strand of one query against R
void find_mems(const SeqSet& R, const KIndex& ix, const std::vector<uint8_t>& q, int qid, int strand,
               const Params& P, std::vector<Mem>& out) {
  int k = ix.k; i64 m = (i64)q.size();
  uint64_t mask = ((uint64_t)1 << (2 * k)) - 1, km = 0; int valid = 0;
  size_t first = out.size();
  for (i64 p = 0; p < m; p++) {
    int c = q[p];
    if (c > 3) { valid = 0; km = 0; continue; }
    km = ((km << 2) | (uint64_t)c) & mask; valid++;
    if (valid < k) continue;
    i64 qs = p - k + 1;
    for (uint32_t h = ix.head[km]; h < ix.head[km + 1]; h++) {
      i64 rs = ix.pos[h];
      // left-maximal: the pair (rs-1, qs-1) must not match
      if (qs > 0 && rs > 0 && q[qs - 1] <= 3 && R.t[rs - 1] == q[qs - 1]) continue;
      i64 l = k;
      while (qs + l < m && q[qs + l] <= 3 && R.t[rs + l] == q[qs + l]) l++;
      if (l >= P.minmatch) { Mem x = {qid, strand, rs, qs, l}; out.push_back(x); }
    }
  }
  if (!P.maxmatch) {
    // -mumreference: keep a match only if its string occurs exactly once in the reference
    size_t w = first;
    for (size_t i = first; i < out.size(); i++) {
      const Mem& x = out[i];
      uint64_t kk = 0;
      for (int j = 0; j < k; j++) kk = (kk << 2) | q[x.s2 + j];
      bool unique = true;
      for (uint32_t h = ix.head[kk]; h < ix.head[kk + 1] && unique; h++) {
        i64 rs = ix.pos[h];
        if (rs == x.s1) continue;
        i64 l = k;
        while (l < x.len && R.t[rs + l] == q[x.s2 + l]) l++;
        if (l >= x.len) unique = false;
      }
      if (unique) out[w++] = x;
    }
    out.resize(w);
  }
  std::sort(out.begin() + first, out.end(), [](const Mem& a, const Mem& b) {
    if (a.s2 != b.s2) return a.s2 < b.s2;
    return a.s1 < b.s1;
  });
}

// ----------------------------------------------------------------------------
// Stage 3: mgaps -- filter, union-find components, repeated best-chain DP
// ----------------------------------------------------------------------------
struct CMatch { i64 s1, s2, len; };
struct Cluster { int q, strand; std::vector<CMatch> m; };

struct GM { i64 s1, s2, len, score, adj; int from; bool good, tent; };

void filter_matches(std::vector<GM>& A) {
  int N = (int)A.size();
  for (int i = 0; i < N; i++) { A[i].good = true; A[i].tent = false; }
  for (int i = 0; i < N - 1; i++) {
    if (!A[i].good) continue;
    i64 i_diag = A[i].s2 - A[i].s1;
    i64 i_end = A[i].s2 + A[i].len;
    for (int j = i + 1; j < N && A[j].s2 <= i_end; j++) {
      if (!A[j].good) continue;
      i64 j_diag = A[j].s2 - A[j].s1;
      if (i_diag == j_diag) {
        i64 j_extent = A[j].len + A[j].s2 - A[i].s2;
        if (j_extent > A[i].len) { A[i].len = j_extent; i_end = A[i].s2 + j_extent; }
        A[j].good = false;
      } else if (A[i].s1 == A[j].s1 || A[i].s2 == A[j].s2) {
        i64 olap = (A[i].s1 == A[j].s1) ? (A[i].s2 + A[i].len - A[j].s2) : (A[i].s1 + A[i].len - A[j].s1);
        if (A[i].len < A[j].len) {
          if (olap >= A[i].len / 2) { A[i].good = false; break; }
        } else if (A[j].len < A[i].len) {
          if (olap >= A[j].len / 2) A[j].good = false;
        } else {
          if (olap >= A[i].len / 2) {
            A[j].tent = true;
            if (A[i].tent) { A[i].good = false; break; }
          }
        }
      }
    }
  }
  int w = 0;
  for (int i = 0; i < N; i++) if (A[i].good) A[w++] = A[i];
  A.resize(w);
}

int uf_find(std::vector<int>& uf, int a) {
  int r = a; while (uf[r] >= 0) r = uf[r];
  while (uf[a] >= 0) { int nx = uf[a]; uf[a] = r; a = nx; }
  return r;
}

void process_component(std::vector<GM> A, int qid, int strand, const Params& P, std::vector<Cluster>& out) {
  while (!A.empty()) {
    int N = (int)A.size();
    for (int i = 0; i < N; i++) {
      A[i].score = A[i].len; A[i].adj = 0; A[i].from = -1; A[i].good = false;
      for (int j = 0; j < i; j++) {
        i64 olap1 = A[j].s1 + A[j].len - A[i].s1;
        i64 olap = std::max<i64>(0, olap1);
        i64 olap2 = A[j].s2 + A[j].len - A[i].s2;
        olap = std::max(olap, olap2);
        i64 pen = olap + std::llabs((A[i].s2 - A[i].s1) - (A[j].s2 - A[j].s1));
        if (A[j].score + A[i].len - pen > A[i].score) {
          A[i].from = j; A[i].score = A[j].score + A[i].len - pen; A[i].adj = olap;
        }
      }
    }
    int best = 0;
    for (int i = 1; i < N; i++) if (A[i].score > A[best].score) best = i;
    i64 total = 0;
    for (int i = best; i >= 0; i = A[i].from) { A[i].good = true; total += A[i].len; }
    if (total >= P.mincluster) {
      Cluster c; c.q = qid; c.strand = strand;
      bool firstm = true;
      for (int i = 0; i < N; i++) if (A[i].good) {
        i64 adj = firstm ? 0 : A[i].adj;
        CMatch mm = {A[i].s1 + adj, A[i].s2 + adj, A[i].len - adj};
        firstm = false;
        if (mm.len > 0) c.m.push_back(mm);
      }
      if (!c.m.empty()) out.push_back(c);
    }
    int w = 0;
    for (int i = 0; i < N; i++) if (!A[i].good) A[w++] = A[i];
    A.resize(w);
  }
}

void mgaps(const Mem* mems, int n, int qid, int strand, const Params& P, std::vector<Cluster>& out) {
  if (n <= 0) return;
  std::vector<GM> A(n);
  for (int i = 0; i < n; i++) { A[i].s1 = mems[i].s1; A[i].s2 = mems[i].s2; A[i].len = mems[i].len; }
  // mems arrive sorted by (s2, s1)
  filter_matches(A);
  int N = (int)A.size();
  std::vector<int> uf(N, -1);
  for (int i = 0; i < N - 1; i++) {
    i64 i_end = A[i].s2 + A[i].len, i_diag = A[i].s2 - A[i].s1;
    for (int j = i + 1; j < N; j++) {
      i64 sep = A[j].s2 - i_end;
      if (sep > P.maxgap) break;
      i64 dd = std::llabs((A[j].s2 - A[j].s1) - i_diag);
      i64 lim = std::max<i64>(P.diagdiff, (i64)(P.diagfactor * (double)sep));
      if (dd <= lim) {
        int a = uf_find(uf, i), b = uf_find(uf, j);
        if (a != b) { if (a < b) uf[b] = a; else uf[a] = b; }  // root = smallest member index
      }
    }
  }
  // components in order of their smallest member; members keep (s2, s1) order
  std::vector<int> root(N);
  for (int i = 0; i < N; i++) root[i] = uf_find(uf, i);
  std::vector<std::vector<GM> > comp;
  std::vector<int> slot(N, -1);
  for (int i = 0; i < N; i++) {
    int r = root[i];
    if (slot[r] < 0) { slot[r] = (int)comp.size(); comp.push_back(std::vector<GM>()); }
    comp[slot[r]].push_back(A[i]);
  }
  for (size_t c = 0; c < comp.size(); c++) process_component(comp[c], qid, strand, P, out);
}

// ----------------------------------------------------------------------------
// Stage 4: postnuc -- cluster extension with the anti-diagonal banded DP
// ----------------------------------------------------------------------------
struct Aln {
  i64 sA, eA, sB, eB;  // 1-based, B on the cluster's strand
  int dir;             // 0 forward, 1 reverse
  std::string ops;     // one char per alignment column: 'M' pair, 'I' A-only (delta > 0), 'D' B-only (delta < 0)
};

struct Stats { i64 cells; i64 max_band; i64 dp_calls; i64 cap_hits; };   // cap_hits: diagonals on which band_cap trimmed the band

// base accessors: A = one reference record (1-based local), B = query strand (1-based)
struct View {
  const uint8_t* p; i64 n;   // p[0..n-1]
  int at(i64 i) const { return (i >= 1 && i <= n) ? p[i - 1] : 4; }
};

inline int sub_score(int a, int b) { return (a <= 3 && a == b) ? GOOD : BAD; }

struct EngOut { bool reached; i64 fi, fj; };

// One run of the DP engine.  Cell (i,j) = i bases of A and j bases of B consumed, anti-diagonal d=i+j.
// A base k (1..N) is Av.at(a0 + dirn*(k-1)), likewise B.
//   search : no traceback is produced
//   forced : per-diagonal best (never terminates early); used to realise a path to a known end point
//   optimal: reaching the corner only counts if the best score is there
EngOut engine(const View& Av, i64 a0, const View& Bv, i64 b0, int dirn, i64 N, i64 M,
              bool search, bool forced, bool optimal, const Params& P, std::string* ops, Stats& st) {
  st.dp_calls++;
  const i64 max_diff = (i64)GOOD * P.breaklen;
  struct Row { i64 lo, hi; std::vector<int> D, I, M; std::vector<uint8_t> tb; };
  std::vector<Row> rows;  // all rows when traceback is needed, else a rolling window (kept simple: all rows' scores freed)
  rows.reserve(1024);
  Row r0; r0.lo = r0.hi = 0; r0.D.assign(1, NEG); r0.I.assign(1, NEG); r0.M.assign(1, 0); r0.tb.assign(1, (uint8_t)(0x3F | (2 << 6)));
  rows.push_back(r0);
  i64 high = 0, FinD = 0, FinJ = 0;
  i64 nlo = std::max<i64>(0, 1 - N), nhi = std::min<i64>(1, M);
  i64 d;
  for (d = 1; d <= N + M && (d - FinD) <= P.breaklen && nlo <= nhi; d++) {
    Row cur; cur.lo = nlo; cur.hi = nhi;
    i64 w = nhi - nlo + 1;
    if (w > st.max_band) st.max_band = w;
    st.cells += w;
    cur.D.resize(w); cur.I.resize(w); cur.M.resize(w); cur.tb.resize(w);
    const Row& p1 = rows[d - 1];
    const Row* p2 = d >= 2 ? &rows[d - 2] : (const Row*)0;
    if (forced) high = NEG * 2;
    std::vector<int> mx(w);
    for (i64 j = nlo; j <= nhi; j++) {
      i64 i = d - j; i64 x = j - nlo;
      int Dv = NEG, Iv = NEG, Mv = NEG; int uD = 3, uI = 3, uM = 3;
      // D: B-only column, parent (i, j-1) on diagonal d-1
      if (j - 1 >= p1.lo && j - 1 <= p1.hi) {
        i64 y = j - 1 - p1.lo;
        int del = p1.D[y] > NEG ? p1.D[y] + GCONT : NEG;
        int ins = p1.I[y] > NEG ? p1.I[y] + GOPEN : NEG;
        int mat = p1.M[y] > NEG ? p1.M[y] + GOPEN : NEG;
        if (del > ins) { if (del > mat) { Dv = del; uD = 0; } else { Dv = mat; uD = 2; } }
        else if (ins > mat) { Dv = ins; uD = 1; } else { Dv = mat; uD = 2; }
        if (Dv <= NEG) { Dv = NEG; uD = 3; }
      }
      // I: A-only column, parent (i-1, j) on diagonal d-1
      if (j >= p1.lo && j <= p1.hi) {
        i64 y = j - p1.lo;
        int del = p1.D[y] > NEG ? p1.D[y] + GOPEN : NEG;
        int ins = p1.I[y] > NEG ? p1.I[y] + GCONT : NEG;
        int mat = p1.M[y] > NEG ? p1.M[y] + GOPEN : NEG;
        if (del > ins) { if (del > mat) { Iv = del; uI = 0; } else { Iv = mat; uI = 2; } }
        else if (ins > mat) { Iv = ins; uI = 1; } else { Iv = mat; uI = 2; }
        if (Iv <= NEG) { Iv = NEG; uI = 3; }
      }
      // M: pair column, parent (i-1, j-1) on diagonal d-2
      if (p2 && j - 1 >= p2->lo && j - 1 <= p2->hi && i >= 1) {
        i64 y = j - 1 - p2->lo;
        int del = p2->D[y], ins = p2->I[y], mat = p2->M[y];
        int v;
        if (del > ins) { if (del > mat) { v = del; uM = 0; } else { v = mat; uM = 2; } }
        else if (ins > mat) { v = ins; uM = 1; } else { v = mat; uM = 2; }
        if (v > NEG) Mv = v + sub_score(Av.at(a0 + dirn * (i - 1)), Bv.at(b0 + dirn * (j - 1)));
        else uM = 3;
      }
      cur.D[x] = Dv; cur.I[x] = Iv; cur.M[x] = Mv;
      int best, bs;
      if (Dv > Iv) { if (Dv > Mv) { best = Dv; bs = 0; } else { best = Mv; bs = 2; } }
      else if (Iv > Mv) { best = Iv; bs = 1; } else { best = Mv; bs = 2; }
      mx[x] = best;
      cur.tb[x] = (uint8_t)(uD | (uI << 2) | (uM << 4) | (bs << 6));
      if (best >= high) { high = best; FinD = d; FinJ = j; }
    }
    // trim hopeless cells from both edges (this only narrows the NEXT diagonal)
    i64 lo = nlo, hi = nhi;
    while (lo <= nhi && high - mx[lo - nlo] > max_diff) lo++;
    while (hi >= nlo && high - mx[hi - nlo] > max_diff) hi--;
    if (lo <= hi && hi - lo + 2 > P.band_cap) st.cap_hits++;
    while (lo <= hi && hi - lo + 2 > P.band_cap) { if (mx[lo - nlo] < mx[hi - nlo]) lo++; else hi--; }
    // Dead tail.  Once the band touches the end of a sequence it can happen that no cell is able to reach the best score
    // again: cell (i, j) can gain at most GOOD * min(N - i, M - j) more.  Checked on every 8th anti-diagonal over the cells just
    // computed; when no cell can, and the walk cannot get to the far corner before the break length runs out
    // (FinD + breaklen < N + M), the anti-diagonals still to come would change nothing that is reported -- FinD, FinJ and the
    // finish cell are fixed, `reached` stays false -- so the engine stops here.  This is not MUMmer's loop but an early exit
    // with the same result; it only lowers the number of cells evaluated (on accurate reads the ~200 anti-diagonals an
    // extension runs past the end of the read).  The CUDA engines apply the same rule, so `cells` stays comparable.
    if (!forced && (d & 7) == 0 && FinD + P.breaklen < N + M && (cur.hi == M || d - cur.lo == N)) {
      i64 ub = (i64)NEG * 4;
      for (i64 j = cur.lo; j <= cur.hi; j++) {
        const int v = mx[j - cur.lo];
        if (v <= NEG) continue;
        ub = std::max<i64>(ub, (i64)v + GOOD * std::min<i64>(N - (d - j), M - j));
      }
      if (ub < high) { rows.push_back(std::move(cur)); d++; break; }
    }
    nlo = std::max<i64>(lo, d + 1 - N);
    nhi = std::min<i64>(hi + 1, std::min<i64>(d + 1, M));
    if (lo > hi) { nlo = 1; nhi = 0; }
    rows.push_back(std::move(cur));
    if (d >= 2) {  // scores of diagonal d-2 are no longer needed; its traceback bits are (unless searching)
      Row& old = rows[d - 2]; std::vector<int>().swap(old.D); std::vector<int>().swap(old.I); std::vector<int>().swap(old.M);
      if (search) std::vector<uint8_t>().swap(old.tb);
    }
  }
  i64 dlast = d - 1;
  EngOut o; o.reached = false;
  if (dlast == N + M) {
    if (!optimal) { o.reached = true; FinD = N + M; FinJ = M; }
    else if (FinD == dlast) o.reached = true;
  }
  o.fi = FinD - FinJ; o.fj = FinJ;
  if (!search && ops) {
    // traceback from the finish cell's best state
    std::string rev;
    i64 cd = FinD, cj = FinJ;
    int state = (rows[cd].tb[cj - rows[cd].lo] >> 6) & 3;
    while (cd > 0) {
      const Row& r = rows[cd]; uint8_t tb = r.tb[cj - r.lo];
      if (state == 0) { rev.push_back('D'); state = tb & 3; cd -= 1; cj -= 1; }
      else if (state == 1) { rev.push_back('I'); state = (tb >> 2) & 3; cd -= 1; }
      else { rev.push_back('M'); state = (tb >> 4) & 3; cd -= 2; cj -= 1; }
    }
    ops->assign(rev.rbegin(), rev.rend());
  }
  return o;
}

struct PCluster { std::vector<CMatch> m; int dir; bool fused; };  // matches in record-local 1-based coords

struct Ext {
  const Params& P; View A, B[2]; std::vector<Aln> al; Stats& st;
  Ext(const Params& p, Stats& s) : P(p), st(s) {}

  // forward extension of alignment ci towards (tA,tB); returns target_reached
  bool extend_forward(int ci, i64 tA, i64 tB, bool forced, bool optimal) {
    Aln& a = al[ci];
    bool overflow = false;
    if (tA - a.eA + 1 > MAX_ALN) { tA = a.eA + MAX_ALN - 1; overflow = true; optimal = true; }
    if (tB - a.eB + 1 > MAX_ALN) { tB = a.eB + MAX_ALN - 1; overflow = true; optimal = true; }
    std::string ops;
    EngOut o = engine(A, a.eA, B[a.dir], a.eB, +1, tA - a.eA + 1, tB - a.eB + 1, false, forced, optimal, P, &ops, st);
    bool reached = o.reached && !overflow;
    // the DP re-aligns the alignment's last column; replace it by the new path
    if (!a.ops.empty()) a.ops.erase(a.ops.size() - 1);
    a.ops += ops;
    a.eA = a.eA + o.fi - 1; a.eB = a.eB + o.fj - 1;
    return reached;
  }

  // backward extension of the freshly created alignment ci towards alignment ti (or the sequence start)
  bool extend_backward(int ci, int ti) {
    Aln& c = al[ci];
    i64 tA, tB; bool optimal = false, overflow = false;
    if (ti >= 0) { tA = al[ti].eA; tB = al[ti].eB; } else { tA = 1; tB = 1; optimal = true; }
    if (c.sA - tA + 1 > MAX_ALN) { tA = c.sA - MAX_ALN + 1; overflow = true; optimal = true; }
    if (c.sB - tB + 1 > MAX_ALN) { tB = c.sB - MAX_ALN + 1; overflow = true; optimal = true; }
    EngOut o = engine(A, c.sA, B[c.dir], c.sB, -1, c.sA - tA + 1, c.sB - tB + 1, true, false, optimal, P, 0, st);
    bool reached = o.reached;
    if (overflow || ti < 0) reached = false;
    if (reached) {
      extend_forward(ti, c.sA, c.sB, true, false);
      Aln& t = al[ti];
      // t now ends on c's first column; append the rest of c
      t.ops += c.ops.substr(1);
      t.eA = c.eA; t.eB = c.eB;
      al.pop_back();
    } else {
      i64 nsA = c.sA - o.fi + 1, nsB = c.sB - o.fj + 1;
      std::string ops;
      engine(A, nsA, B[c.dir], nsB, +1, c.sA - nsA + 1, c.sB - nsB + 1, false, true, false, P, &ops, st);
      c.ops = ops + c.ops.substr(1);
      c.sA = nsA; c.sB = nsB;
    }
    return reached;
  }

  bool shadowed(const PCluster& c, int ap) {
    if (al.empty()) return false;
    i64 sA = c.m.front().s1, eA = c.m.back().s1 + c.m.back().len - 1;
    i64 sB = c.m.front().s2, eB = c.m.back().s2 + c.m.back().len - 1;
    for (int k = ap; k >= 0; k--)
      if (al[k].dir == c.dir && al[k].eA >= eA && al[k].eB >= eB && al[k].sA <= sA && al[k].sB <= sB) return true;
    return false;
  }

  int forward_target(std::vector<PCluster>& C, int cur, i64& tA, i64& tB) {
    const PCluster& c = C[cur];
    i64 sA = c.m.back().s1 + c.m.back().len - 1, sB = c.m.back().s2 + c.m.back().len - 1;
    i64 dist = std::min(tA - sA, tB - sB);
    int res = -1;
    for (int k = cur + 1; k < (int)C.size(); k++) {
      if (C[k].dir != c.dir) continue;
      i64 eA = C[k].m.front().s1, eB = C[k].m.front().s2;
      if ((eA < sA || eB < sB) && C[k].m.back().s1 >= sA && C[k].m.back().s2 >= sB) {
        for (size_t t = 0; t < C[k].m.size() && (eA < sA || eB < sB); t++) { eA = C[k].m[t].s1; eB = C[k].m[t].s2; }
      }
      if (eA >= sA && eB >= sB) {
        i64 greater, lesser;
        if (eA - sA > eB - sB) { greater = eA - sA; lesser = eB - sB; } else { lesser = eA - sA; greater = eB - sB; }
        if (greater < P.breaklen || lesser * GOOD + (greater - lesser) * GCONT >= 0) { res = k; tA = eA; tB = eB; break; }
        else if ((greater << 1) - lesser < dist) { res = k; tA = eA; tB = eB; dist = (greater << 1) - lesser; }
      }
    }
    return res;
  }

  int reverse_target(int cur) {
    const Aln& c = al[cur];
    i64 sA = c.sA, sB = c.sB;
    i64 dist = std::min(sA, sB);
    int res = -1;
    for (int k = cur - 1; k >= 0; k--) {
      if (al[k].dir != c.dir) continue;
      i64 eA = al[k].eA, eB = al[k].eB;
      if (eA <= sA && eB <= sB) {
        i64 greater, lesser;
        if (sA - eA > sB - eB) { greater = sA - eA; lesser = sB - eB; } else { lesser = sA - eA; greater = sB - eB; }
        if (greater < P.breaklen || lesser * GOOD + (greater - lesser) * GCONT >= 0) { res = k; break; }
        else if ((greater << 1) - lesser < dist) { res = k; dist = (greater << 1) - lesser; }
      }
    }
    return res;
  }

  void extend_clusters(std::vector<PCluster>& C) {
    bool target_reached = false;
    i64 tA = 0, tB = 0;
    int prev = 0, cur = 0, curA = -1, targetC = -1;
    int nC = (int)C.size();
    while (cur < nC) {
      if (!target_reached) {
        if (C[cur].fused || (P.simplify && shadowed(C[cur], curA))) {
          C[cur].fused = true; cur = ++prev; continue;
        }
      }
      PCluster& c = C[cur];
      int nm = (int)c.m.size();
      for (int mi = 0; mi < nm; mi++) {
        const CMatch& mt = c.m[mi];
        if (target_reached) {
          if (al[curA].eA != mt.s1 || al[curA].eB != mt.s2) {
            if (mi >= nm - 1) { fprintf(stderr, "oracle: target match does not exist\n"); abort(); }
            continue;
          }
          al[curA].eA += mt.len - 1; al[curA].eB += mt.len - 1;
          al[curA].ops.append((size_t)(mt.len - 1), 'M');
        } else {
          Aln a; a.sA = mt.s1; a.sB = mt.s2; a.eA = mt.s1 + mt.len - 1; a.eB = mt.s2 + mt.len - 1; a.dir = c.dir;
          a.ops.assign((size_t)mt.len, 'M');
          al.push_back(a); curA = (int)al.size() - 1;
          int ti = reverse_target(curA);
          if (extend_backward(curA, ti)) curA = ti;
        }
        if (mi < nm - 1) {
          tA = c.m[mi + 1].s1; tB = c.m[mi + 1].s2;
          target_reached = extend_forward(curA, tA, tB, false, false);
        } else {
          tA = A.n; tB = B[0].n;
          targetC = forward_target(C, cur, tA, tB);
          target_reached = extend_forward(curA, tA, tB, false, targetC < 0);
        }
      }
      if (targetC < 0) target_reached = false;
      c.fused = true;
      if (!target_reached) cur = ++prev; else cur = targetC;
    }
  }
};

struct Row {
  int ref, qry, dir; i64 sA, eA, sB, eB; i64 errors, cols;
  i64 delta_begin, delta_end;
};

struct Result {
  std::vector<Mem> mems;
  std::vector<Cluster> clusters;
  std::vector<Row> rows;
  std::vector<i64> deltas;
  Stats st;
};

void finish_alignment(const Aln& a, const View& A, const View& B, int ref, int qry, i64 qlen, Result& R) {
  Row r; r.ref = ref; r.qry = qry; r.dir = a.dir; r.sA = a.sA; r.eA = a.eA;
  i64 ia = a.sA, ib = a.sB, errors = 0, cols = 0, run = 0;
  r.delta_begin = (i64)R.deltas.size();
  for (size_t k = 0; k < a.ops.size(); k++) {
    char op = a.ops[k]; cols++; run++;
    if (op == 'M') { if (A.at(ia) != B.at(ib)) errors++; ia++; ib++; }
    else if (op == 'I') { errors++; ia++; R.deltas.push_back(run); run = 0; }
    else { errors++; ib++; R.deltas.push_back(-run); run = 0; }
  }
  r.delta_end = (i64)R.deltas.size();
  if (ia != a.eA + 1 || ib != a.eB + 1) { fprintf(stderr, "oracle: inconsistent alignment path\n"); abort(); }
  r.errors = errors; r.cols = cols;
  if (a.dir == 0) { r.sB = a.sB; r.eB = a.eB; } else { r.sB = qlen - a.sB + 1; r.eB = qlen - a.eB + 1; }
  R.rows.push_back(r);
}

void run(const SeqSet& Rf, const KIndex& ix, const SeqSet& Q, const Params& P, int stage_limit, Result& R) {
  R.st.cells = 0; R.st.max_band = 0; R.st.dp_calls = 0; R.st.cap_hits = 0;
  std::vector<uint8_t> fw, rc;
  for (int q = 0; q < Q.n(); q++) {
    i64 L = Q.len[q];
    fw.assign(Q.t.begin() + Q.start[q], Q.t.begin() + Q.start[q] + L);
    rc.resize(L);
    for (i64 i = 0; i < L; i++) { int c = fw[L - 1 - i]; rc[i] = (uint8_t)(c <= 3 ? 3 - c : 4); }
    size_t m0 = R.mems.size();
    find_mems(Rf, ix, fw, q, 0, P, R.mems);
    size_t m1 = R.mems.size();
    find_mems(Rf, ix, rc, q, 1, P, R.mems);
    size_t m2 = R.mems.size();
    if (stage_limit < 2) continue;
    size_t c0 = R.clusters.size();
    mgaps(R.mems.data() + m0, (int)(m1 - m0), q, 0, P, R.clusters);
    mgaps(R.mems.data() + m1, (int)(m2 - m1), q, 1, P, R.clusters);
    if (stage_limit < 3) continue;
    // postnuc: split clusters at reference-record boundaries, group per reference record
    // postnuc: split clusters at reference-record boundaries and group them per reference record; the
    // groups are walked in order of first appearance in the cluster stream (forward strand first), which
    // is the order of the `>ref qry` blocks in the delta file (pinned by the reference README.md:111-129,
    // whose un-sorted show-coords tables list Segkk1/Segkk1 before Segkk0/Segkk1).
    std::vector<std::vector<PCluster> > syn(Rf.n());
    std::vector<int> rec_order;
    for (size_t c = c0; c < R.clusters.size(); c++) {
      const Cluster& cl = R.clusters[c];
      int lastrec = -1;
      for (size_t t = 0; t < cl.m.size(); t++) {
        int rec = Rf.rec_of(cl.m[t].s1);
        if (rec != lastrec) {
          PCluster pc; pc.dir = cl.strand; pc.fused = false;
          if (syn[rec].empty()) rec_order.push_back(rec);
          syn[rec].push_back(pc); lastrec = rec;
        }
        CMatch mm = {cl.m[t].s1 - Rf.start[rec] + 1, cl.m[t].s2 + 1, cl.m[t].len};
        syn[rec].back().m.push_back(mm);
      }
    }
    for (size_t ro = 0; ro < rec_order.size(); ro++) {
      int rec = rec_order[ro];
      std::vector<PCluster>& C = syn[rec];
      std::stable_sort(C.begin(), C.end(), [](const PCluster& a, const PCluster& b) {
        if (a.m.front().s1 != b.m.front().s1) return a.m.front().s1 < b.m.front().s1;
        if (a.dir != b.dir) return a.dir < b.dir;
        return a.m.front().s2 < b.m.front().s2;
      });
      Ext E(P, R.st);
      E.A.p = Rf.t.data() + Rf.start[rec]; E.A.n = Rf.len[rec];
      E.B[0].p = fw.data(); E.B[0].n = L; E.B[1].p = rc.data(); E.B[1].n = L;
      E.extend_clusters(C);
      for (size_t a = 0; a < E.al.size(); a++) finish_alignment(E.al[a], E.A, E.B[E.al[a].dir], rec, q, L, R);
    }
  }
}

}  // namespace

// ----------------------------------------------------------------------------
// C ABI for ctypes (tests only)
// ----------------------------------------------------------------------------
extern "C" {

struct orc_params {
  int32_t minmatch, mincluster, maxgap, breaklen, diagdiff, maxmatch, simplify, band_cap;
  double diagfactor;
};

// stage_limit: 1 = seeds only, 2 = +clusters, 3 = +alignments
void* orc_run(const char* ref, const int64_t* ref_off, int n_ref, const char* qry, const int64_t* qry_off, int n_qry,
              const orc_params* p, int stage_limit) {
  Params P; P.minmatch = p->minmatch; P.mincluster = p->mincluster; P.maxgap = p->maxgap; P.breaklen = p->breaklen;
  P.diagdiff = p->diagdiff; P.diagfactor = p->diagfactor; P.maxmatch = p->maxmatch; P.simplify = p->simplify;
  P.band_cap = p->band_cap;
  SeqSet Rf, Q; load(Rf, ref, ref_off, n_ref); load(Q, qry, qry_off, n_qry);
  KIndex ix; build_index(Rf, std::min(P.minmatch, 12), ix);
  Result* R = new Result();
  run(Rf, ix, Q, P, stage_limit, *R);
  return R;
}

// The reference set loaded and indexed once, then shared (read-only) by any number of threads calling orc_run_ix:
// what the timed CPU arm of bench.py uses, so that no thread rebuilds the k-mer index per call.
struct RefHandle { SeqSet Rf; KIndex ix; };
void* orc_ref_build(const char* ref, const int64_t* ref_off, int n_ref, int minmatch) {
  RefHandle* h = new RefHandle();
  load(h->Rf, ref, ref_off, n_ref);
  build_index(h->Rf, std::min(minmatch, 12), h->ix);
  return h;
}
void orc_ref_free(void* h) { delete (RefHandle*)h; }
void* orc_run_ix(void* ref_handle, const char* qry, const int64_t* qry_off, int n_qry, const orc_params* p, int stage_limit) {
  RefHandle* h = (RefHandle*)ref_handle;
  Params P; P.minmatch = p->minmatch; P.mincluster = p->mincluster; P.maxgap = p->maxgap; P.breaklen = p->breaklen;
  P.diagdiff = p->diagdiff; P.diagfactor = p->diagfactor; P.maxmatch = p->maxmatch; P.simplify = p->simplify;
  P.band_cap = p->band_cap;
  if (h->ix.k > P.minmatch) return 0;
  SeqSet Q; load(Q, qry, qry_off, n_qry);
  Result* R = new Result();
  run(h->Rf, h->ix, Q, P, stage_limit, *R);
  return R;
}

// One run of the DP engine on two plain sequences (test hook for oracle/bruteforce.py): a[0..n), b[0..m) as characters,
// forward direction from (1, 1).  out: reached, fi, fj; ops (at most n + m + 1 chars, NUL-terminated): 'M' pair, 'I' a-only, 'D' b-only.
void orc_dp(const char* a, int n, const char* b, int m, int breaklen, int band_cap, int forced, int optimal, int64_t* out, char* ops_out) {
  std::vector<uint8_t> ca(n), cb(m);
  for (int i = 0; i < n; i++) ca[i] = (uint8_t)code_of(a[i]);
  for (int i = 0; i < m; i++) cb[i] = (uint8_t)code_of(b[i]);
  Params P; P.minmatch = 20; P.mincluster = 65; P.maxgap = 90; P.breaklen = breaklen; P.diagdiff = 5; P.diagfactor = 0.12;
  P.maxmatch = 1; P.simplify = 1; P.band_cap = band_cap;
  View A, B; A.p = ca.data(); A.n = n; B.p = cb.data(); B.n = m;
  Stats st; st.cells = 0; st.max_band = 0; st.dp_calls = 0; st.cap_hits = 0;
  std::string ops;
  EngOut o = engine(A, 1, B, 1, +1, n, m, false, forced != 0, optimal != 0, P, &ops, st);
  out[0] = o.reached ? 1 : 0; out[1] = o.fi; out[2] = o.fj; out[3] = st.cells; out[4] = st.max_band;
  memcpy(ops_out, ops.c_str(), ops.size() + 1);
}

void orc_free(void* h) { delete (Result*)h; }

int64_t orc_n_mems(void* h) { return (int64_t)((Result*)h)->mems.size(); }
// columns: q, strand, s1 (global ref pos, 0-based, one separator between records), s2 (0-based on strand), len
void orc_get_mems(void* h, int64_t* out) {
  Result* R = (Result*)h;
  for (size_t i = 0; i < R->mems.size(); i++) {
    const Mem& m = R->mems[i];
    out[5 * i] = m.q; out[5 * i + 1] = m.strand; out[5 * i + 2] = m.s1; out[5 * i + 3] = m.s2; out[5 * i + 4] = m.len;
  }
}

int64_t orc_n_clusters(void* h) { return (int64_t)((Result*)h)->clusters.size(); }
int64_t orc_n_cluster_matches(void* h) {
  Result* R = (Result*)h; int64_t n = 0;
  for (size_t i = 0; i < R->clusters.size(); i++) n += (int64_t)R->clusters[i].m.size();
  return n;
}
// cl: q, strand, first match index, n matches ; cm: s1, s2, len
void orc_get_clusters(void* h, int64_t* cl, int64_t* cm) {
  Result* R = (Result*)h; int64_t w = 0;
  for (size_t i = 0; i < R->clusters.size(); i++) {
    const Cluster& c = R->clusters[i];
    cl[4 * i] = c.q; cl[4 * i + 1] = c.strand; cl[4 * i + 2] = w; cl[4 * i + 3] = (int64_t)c.m.size();
    for (size_t t = 0; t < c.m.size(); t++) { cm[3 * w] = c.m[t].s1; cm[3 * w + 1] = c.m[t].s2; cm[3 * w + 2] = c.m[t].len; w++; }
  }
}

int64_t orc_n_rows(void* h) { return (int64_t)((Result*)h)->rows.size(); }
int64_t orc_n_deltas(void* h) { return (int64_t)((Result*)h)->deltas.size(); }
// rows: ref, qry, dir, sA, eA, sB, eB, errors, cols, delta_begin, delta_end   (delta-file order)
void orc_get_rows(void* h, int64_t* out, int64_t* deltas) {
  Result* R = (Result*)h;
  for (size_t i = 0; i < R->rows.size(); i++) {
    const Row& r = R->rows[i]; int64_t* o = out + 11 * i;
    o[0] = r.ref; o[1] = r.qry; o[2] = r.dir; o[3] = r.sA; o[4] = r.eA; o[5] = r.sB; o[6] = r.eB;
    o[7] = r.errors; o[8] = r.cols; o[9] = r.delta_begin; o[10] = r.delta_end;
  }
  if (deltas) for (size_t i = 0; i < R->deltas.size(); i++) deltas[i] = R->deltas[i];
}
void orc_get_stats(void* h, int64_t* out) {
  Result* R = (Result*)h; out[0] = R->st.cells; out[1] = R->st.max_band; out[2] = R->st.dp_calls; out[3] = R->st.cap_hits;
}

// Coverage profile, restating srcRefactor/misassemblyFixerLib/merger.py:126-140 (assignCoverage):
// cov[contig][i] += 1 for i in [S1-1, E1).
void orc_coverage(const int32_t* ref_id, const int32_t* s1, const int32_t* e1, int64_t n,
                  const int64_t* contig_off, int n_contig, int32_t* cov) {
  (void)n_contig;
  for (int64_t r = 0; r < n; r++) {
    int64_t base = contig_off[ref_id[r]];
    for (int64_t i = s1[r] - 1; i < e1[r]; i++) cov[base + i] += 1;
  }
}

}  // extern "C"
