// mfsc_extend.cuh -- stage 4: postnuc-style extension of the clusters of one read into alignments
// (nucmer's last stage, `postnuc -b breaklen`; reference call site alignerRobot.py:46-64, breaklen
// 200, or 50 with houseKeeper.globalFast).
//
// One warp owns one read (both strands).  The control flow that walks clusters, picks merge
// targets and drops shadowed clusters is scalar and warp-uniform; the work is the banded
// anti-diagonal DP (nucleotide scores +3 / -7, affine gaps -7 open / -4 extend, three states per
// cell) whose band is spread over the lanes.  Per cell the DP also carries the number of
// diagonal columns and mismatches of the path that produced each state, so no traceback matrix is
// needed for the rows: the arithmetic show-coords needs (S1 E1 S2 E2, errors, columns) falls out of the
// finish cell.  Only when the caller wants the delta indel lists does the TB = true instantiation record
// one byte per cell in HBM and walk it back (see TbCtx below).
//
// Band state lives in shared memory (9 KB per warp), indexed circularly by the B coordinate j and
// updated in place.  A cell does not store its three states: it stores what its two gap children
// will need -- the best "open or extend" value towards (i, j+1) and towards (i+1, j) -- and its
// best state for the diagonal child, each with the carried (columns, errors) word.  A cell then
// costs 6 shared loads, 6 shared stores and three 3-way selects.  Slots just outside the band
// hold a large negative sentinel so no cell tests band membership.  The bases of both sequences
// are decoded from the 2-bit text into two small circular byte windows 128 positions at a time.
#pragma once
#include "mfsc_cluster.cuh"

#define DP_GOOD 3
#define DP_BAD (-7)
#define DP_GOPEN (-7)
#define DP_GCONT (-4)
#define DP_NEG (-(1 << 22))     // "unreachable"; never clamped: real scores stay above -2^18, and (score << 8) must fit an int
#define DP_MAX_ALN 10000
#define DP_SLOTS 256            // circular band storage; band_cap must stay below DP_SLOTS - 6
#define DP_SMASK (DP_SLOTS - 1)
#define DP_WIN 512              // decoded-base window per sequence (bytes)
#define DP_WMASK (DP_WIN - 1)
#define DP_FILL 128             // bases decoded per refill
#define DP_SMEM_INTS (8 * DP_SLOTS + 2 * DP_WIN / 4)
// carried word of a path: (diagonal columns << 16) | mismatching diagonal columns.  Gap columns add nothing;
// at the finish cell (i, j): columns = i + j - diag, errors = mismatches + i + j - 2*diag.
#define AUX_PAIR 0x10000
// integer add issued as a multiply-add with a run-time 1: goes to the FMA pipe instead of the (saturated) ALU pipe
#define FADD(a, b) ((a) * one + (b))

struct DpOut { int reached, fi, fj, cols, errs, n_ops; };

struct ExtStats { unsigned long long* cells; unsigned long long* calls; int* max_band; unsigned long long* caps; unsigned long long* gaps; };   // caps: anti-diagonals on which band_cap trimmed the band

// View of one sequence: 1-based position i lives at padded position base + i - 1
struct SeqView { unsigned base; int n; };

struct RowsOut {
  // alignments of a read are written into the read's slice (first match index of its forward strand)
  int* ref; int* qry; int* dir; int* sA; int* eA; int* sB; int* eB; int* errs; int* cols;
  int* n_rows;  // [n_reads]
  // delta mode only: token list of each alignment (arena offsets) and the slice of the delta buffer it was written to
  int* head; int* tail; int* dl_begin; int* dl_end;
};

// Delta (indel offset list) support.  The DP itself never needs a traceback (see the header); when the caller
// asks for the delta lists a second instantiation of the kernel (TB = true) also writes one byte per cell --
// best state, "D extended", "I extended" -- into a per-warp buffer in HBM, walks it back from the finish cell
// after each DP that contributes columns to an alignment, and keeps every alignment as a linked list of token
// chunks (run of diagonal columns >= 0, -1 = column consuming A only, -2 = column consuming B only).
#define TB_DIAGS (2 * DP_MAX_ALN + 8)
struct TbCtx {
  unsigned char* tb;    // [TB_DIAGS * DP_SLOTS] codes of the current DP, row d, column j & DP_SMASK
  unsigned char* rev;   // [TB_DIAGS] columns of the last traceback, last column first (0 diagonal, 1 A only, 2 B only)
  int* arena; int cap; int top; int bad;      // token arena of the read being extended (cap ints)
};
struct DeltaOut { int* buf; unsigned long long* used; long long cap; unsigned* n_bad; };

struct ExtCtx {
  SeqStore RS, QS; DevParams P; int* sm; ExtStats st;
  mutable unsigned long long acc_cells, acc_calls; mutable int acc_maxw, acc_caps, acc_gaps;   // per-warp statistics, flushed once per kernel
  mutable TbCtx T;
  mutable int pk_fail;           // packed engine only: a field did not fit, the read goes to the general engine
  // The alignment the walk is working on, kept in registers (the same values on every lane): a row in HBM that lane 0 has just
  // written is a trip to L2 for the next read of it, and the walk touches the current row several times per match.
  mutable int c_idx, c_sA, c_sB, c_eA, c_eB, c_cols, c_errs;
  SeqView A; SeqView B[2];
  RowsOut R; int al0, al1;       // alignments of the current (record, read) group: rows [al0, al1)
  const int* ord; int nC;        // clusters of the group in walking order -> split cluster id
  int pc_rc0;                    // split cluster ids >= pc_rc0 belong to the reverse strand
  const int* pc_first; const int* pc_n; unsigned char* fused;
  const unsigned* cm_s1; const int* cm_s2; const int* cm_len;
  unsigned a_ostart;             // separator-joined coordinate of A's first base
  int ref_id, qry_id, lane;
  const int4* gap_res;           // stage 4a results per cluster match (mfsc_gaps.cuh), or nullptr
};

// Best of three states with the tie rule of the reference DP: D only when strictly above both,
// I only when strictly above M and not below D.
MFSC_DEV void best3(int d, int i, int m, int ad, int ai, int am, int& v, int& a) {
  const bool p = d > i;
  const int x = p ? d : i, ax = p ? ad : ai;
  const bool q = x > m;
  v = q ? x : m; a = q ? ax : am;
}

// decode local positions [from, from + DP_FILL) of a sequence walk (position t -> seq pos s0 + dirn*t, 1-based);
// the code of position t is stored at window byte t + off
MFSC_DEV void dp_fill(const SeqStore& S, SeqView V, int s0, int dirn, int n_local, int from, int off, unsigned char* win, int lane) {
  const int to = from + DP_FILL < n_local + 8 ? from + DP_FILL : n_local + 8;   // a few invalid codes past the end, never more
  for (int t = from + lane; t < to; t += MFSC_LANES) {
    int c = 4;
    if (t < n_local) c = code_at(S, V.base + (unsigned)(s0 + dirn * t - 1));
    win[(t + off) & DP_WMASK] = (unsigned char)c;
  }
}

MFSC_DEV int4 ld4(const int* p) { return *reinterpret_cast<const int4*>(p); }
MFSC_DEV void st4(int* p, int a, int b, int c, int d) { int4 v; v.x = a; v.y = b; v.z = c; v.w = d; *reinterpret_cast<int4*>(p) = v; }

// one DP cell; Dv/Iv: gap states handed over by the parents, rB/rA: best state of the diagonal parent
// The three 3-way selections of a cell share their winners.  Tie rule of the reference DP: the diagonal state
// wins ties, then the A-gap state (I), then the B-gap state (D).
//   best        : I vs M -> (vIM, aIM);  D wins iff D > vIM                          -> (bv, ba), pb = "D won"
//   towards j+1 : D extends (GCONT) or the best state opens (GOPEN): D wins iff D+GCONT > bv+GOPEN.  (When the best
//                 state is D itself its "open" candidate D+GOPEN is below D+GCONT, so using bv is exact.)
//   towards i+1 : I extends or the best state opens: I wins iff I+GCONT > bv+GOPEN, or on equality when the best
//                 state is D (I beats D on ties but loses to M).
#define DP_CELL(C, DV, XD, IV, XI, RB, RA, OB, OA, OD, OAD, OI, OAI, OT)                               \
  {                                                                                                      \
    const bool nm_ = (ym & (0xffu << (8 * C))) != 0u, ne_ = (xm & (0xffu << (8 * C))) != 0u;             \
    const int Mv_ = FADD(RB, nm_ ? DP_BAD : DP_GOOD);                                                    \
    const int xM_ = RA + (ne_ ? AUX_PAIR + 1 : AUX_PAIR);                                                \
    const bool p1_ = IV > Mv_;                                                                           \
    const int vIM_ = p1_ ? IV : Mv_, aIM_ = p1_ ? XI : xM_;                                              \
    const bool pb_ = DV > vIM_;                                                                          \
    const int bv_ = pb_ ? DV : vIM_, ba_ = pb_ ? XD : aIM_;                                              \
    OB = bv_; OA = ba_;                                                                                  \
    const int c_ = FADD(bv_, DP_GOPEN), e_ = FADD(DV, DP_GCONT), f_ = FADD(IV, DP_GCONT);                \
    const bool pd_ = e_ > c_;                                                                            \
    OD = pd_ ? e_ : c_; OAD = pd_ ? XD : ba_;                                                            \
    const bool pi_ = (f_ > c_) || (pb_ && f_ == c_);                                                     \
    OI = pi_ ? f_ : c_; OAI = pi_ ? XI : ba_;                                                            \
    /* traceback code: bits 0-1 best state (0 D, 1 I, 2 M), bit 2 D extended towards j+1, bit 3 I extended towards i+1 */ \
    OT = (pb_ ? 0u : (p1_ ? 1u : 2u)) | (pd_ ? 4u : 0u) | (pi_ ? 8u : 0u);                              \
  }

// One pass of the lanes over 32 groups of four cells, highest group GT on lane 0.  B0..B3/BJ receive the lane's
// four best-state values (DP_NEG outside the band) and the group's first j.
#define DP_PASS(GT, B0, B1, B2, B3, BJ)                                                                                 \
  {                                                                                                                      \
    const bool actg = (GT) - lane >= glo;                                                                                \
    const int g = actg ? (GT) - lane : glo;          /* idle lanes shadow the lowest group (loads only) */               \
    const int j0 = g * 4;                                                                                                \
    const int y = j0 & DP_SMASK, y1 = (j0 - 1) & DP_SMASK;                                                               \
    const int4 vD = ld4(nD + y), vaD = ld4(naD + y), vI = ld4(nI + y), vaI = ld4(naI + y), vB = ld4(cB + y), vA = ld4(cA + y); \
    const int sD = nD[y1], saD = naD[y1], sB = cB[y1], sA = cA[y1];                                                      \
    const int as = d - j0 - 4;                      /* lowest A window byte of the group (cell 3) */                     \
    const unsigned w0 = wA32[(as >> 2) & (DP_WIN / 4 - 1)], w1 = wA32[((as >> 2) + 1) & (DP_WIN / 4 - 1)];               \
    const unsigned aw = dev_funnel_r(w0, w1, (unsigned)(as & 3) * 8u);                                                   \
    const unsigned bw = wB32[(j0 >> 2) & (DP_WIN / 4 - 1)];                                                              \
    /* per byte c (= cell c): xm != 0 where the two bases differ, ym != 0 where they do not score as a match */          \
    const unsigned awr = dev_byte_rev(aw);                                                                               \
    const unsigned xm = awr ^ bw, ym = xm | ((awr | bw) & 0x04040404u);                                                  \
    wsync();                                                                                                             \
    if (actg) {                                                                                                          \
      int o0b, o0a, o0d, o0ad, o0i, o0ai, o1b, o1a, o1d, o1ad, o1i, o1ai, o2b, o2a, o2d, o2ad, o2i, o2ai, o3b, o3a, o3d, o3ad, o3i, o3ai; \
      unsigned o0t, o1t, o2t, o3t;                                                                                       \
      DP_CELL(3, vD.z, vaD.z, vI.w, vaI.w, vB.z, vA.z, o3b, o3a, o3d, o3ad, o3i, o3ai, o3t)                              \
      DP_CELL(2, vD.y, vaD.y, vI.z, vaI.z, vB.y, vA.y, o2b, o2a, o2d, o2ad, o2i, o2ai, o2t)                              \
      DP_CELL(1, vD.x, vaD.x, vI.y, vaI.y, vB.x, vA.x, o1b, o1a, o1d, o1ad, o1i, o1ai, o1t)                              \
      DP_CELL(0, sD, saD, vI.x, vaI.x, sB, sA, o0b, o0a, o0d, o0ad, o0i, o0ai, o0t)                                      \
      /* cells of the group outside [nlo, nhi] hold junk nobody reads (the halo is rewritten below), except in the */    \
      /* best buffer, which the trim and the maximum look at */                                                          \
      const int r0 = j0 - nlo;                                                                                           \
      if ((unsigned)(r0 + 3) >= (unsigned)w) o3b = DP_NEG;                                                               \
      if ((unsigned)(r0 + 2) >= (unsigned)w) o2b = DP_NEG;                                                               \
      if ((unsigned)(r0 + 1) >= (unsigned)w) o1b = DP_NEG;                                                               \
      if ((unsigned)r0 >= (unsigned)w) o0b = DP_NEG;                                                                     \
      st4(cB + y, o0b, o1b, o2b, o3b); st4(cA + y, o0a, o1a, o2a, o3a);                                                  \
      st4(nD + y, o0d, o1d, o2d, o3d); st4(naD + y, o0ad, o1ad, o2ad, o3ad);                                             \
      st4(nI + y, o0i, o1i, o2i, o3i); st4(naI + y, o0ai, o1ai, o2ai, o3ai);                                             \
      int k = o3b * k256 + (r0 + 3);                                                                                     \
      lkey = k > lkey ? k : lkey;                                                                                        \
      k = o2b * k256 + (r0 + 2); lkey = k > lkey ? k : lkey;                                                             \
      k = o1b * k256 + (r0 + 1); lkey = k > lkey ? k : lkey;                                                             \
      k = o0b * k256 + r0; lkey = k > lkey ? k : lkey;                                                                   \
      B0 = o0b; B1 = o1b; B2 = o2b; B3 = o3b; BJ = j0;                                                                   \
      if (TB) *reinterpret_cast<unsigned*>(tbp + (size_t)d * DP_SLOTS + y) = o0t | (o1t << 8) | (o2t << 16) | (o3t << 24); \
    }                                                                                                                    \
    wsync();                                                                                                             \
  }

// Narrow diagonals (at most one cell per lane): the same cell arithmetic without the group-of-four machinery.
// Lane l owns j = nhi - l.  B0/BJ receive the lane's best-state value (DP_NEG when the lane is idle) and its j.
#define DP_PASS1(B0, BJ)                                                                                                 \
  {                                                                                                                      \
    const bool act1 = lane < w;                                                                                          \
    const int j1 = act1 ? nhi - lane : nhi;          /* idle lanes shadow the top cell (loads only) */                   \
    const int y = j1 & DP_SMASK, y1 = (j1 - 1) & DP_SMASK;                                                               \
    const int sD = nD[y1], saD = naD[y1], sI = nI[y], saI = naI[y], sB = cB[y1], sA = cA[y1];                            \
    const unsigned ca1 = wA[(d - j1 - 1) & DP_WMASK], cb1 = wB[j1 & DP_WMASK];                                           \
    const unsigned xm = ca1 ^ cb1, ym = xm | ((ca1 | cb1) & 4u);                                                         \
    wsync();                                                                                                             \
    if (act1) {                                                                                                          \
      int ob, oa, od, oad, oi, oai; unsigned ot;                                                                         \
      DP_CELL(0, sD, saD, sI, saI, sB, sA, ob, oa, od, oad, oi, oai, ot)                                                 \
      cB[y] = ob; cA[y] = oa; nD[y] = od; naD[y] = oad; nI[y] = oi; naI[y] = oai;                                        \
      const int k = ob * k256 + (j1 - nlo);                                                                              \
      lkey = k > lkey ? k : lkey;                                                                                        \
      B0 = ob; BJ = j1;                                                                                                  \
      if (TB) tbp[(size_t)d * DP_SLOTS + y] = (unsigned char)ot;                                                         \
    }                                                                                                                    \
    wsync();                                                                                                             \
  }

// Walk the traceback codes of the current DP back from the finish cell; the columns land in E.T.rev, last column
// first.  Returns their number (all lanes).
MFSC_DEV int dp_traceback(const ExtCtx& E, int FinD, int FinJ) {
  const unsigned char* tbp = E.T.tb;
  wsync();
  int n = 0;
  if (E.lane == 0) {
    unsigned char* rev = E.T.rev;
    int dd = FinD, jj = FinJ;
    int s = dd > 0 ? (int)(tbp[(size_t)dd * DP_SLOTS + (jj & DP_SMASK)] & 3u) : 2;
    while (dd > 0 && n < TB_DIAGS) {
      if (s == 2) {
        rev[n++] = 0; dd -= 2; jj -= 1;
        if (dd > 0) s = (int)(tbp[(size_t)dd * DP_SLOTS + (jj & DP_SMASK)] & 3u);
      } else if (s == 0) {
        rev[n++] = 2; dd -= 1; jj -= 1;
        const unsigned pc = tbp[(size_t)dd * DP_SLOTS + (jj & DP_SMASK)];
        s = (pc & 4u) ? 0 : (int)(pc & 3u);
      } else {
        rev[n++] = 1; dd -= 1;
        const unsigned pc = tbp[(size_t)dd * DP_SLOTS + (jj & DP_SMASK)];
        s = (pc & 8u) ? 1 : (int)(pc & 3u);
      }
    }
    if (dd != 0) E.T.bad = 1;
  }
  n = wbcast_i32(n, 0);
  wsync();
  return n;
}

// Cell (i,j): i bases of A and j bases of B consumed, anti-diagonal d = i + j.  A base t (1..N) is
// A(a0 + dirn*(t-1)), likewise B.  forced: the best cell is re-chosen on every diagonal (the DP
// never gives up); optimal: reaching the far corner only counts if the best score is there.
// Each lane owns a group of four consecutive j (aligned to 4) per step: 128-bit shared loads and
// stores, four independent cells of instruction-level parallelism.
template <bool TB>
MFSC_NOINLINE DpOut dp_engine(const ExtCtx& E, int dir, int a0, int b0, int dirn, int N_, int M_, bool forced_, bool optimal, bool want_ops) {
  // Every per-diagonal quantity (band limits, best score, finish cell) is the same on all lanes.  Passing the
  // inputs through a warp reduction makes that visible to the compiler (CREDUX writes a uniform register), so
  // the per-diagonal bookkeeping is scheduled on the uniform datapath instead of the saturated vector ALU pipe.
  const int N = wmax_i32(N_), M = wmax_i32(M_);
  const bool forced = wmax_i32(forced_ ? 1 : 0) != 0;
  const SeqStore& RS = E.RS; const SeqStore& QS = E.QS; const SeqView Av = E.A; const SeqView Bv = E.B[dir];
  const DevParams& P = E.P;
#ifdef MFSC_EMU
  int* sm = E.sm;
#else
  extern __shared__ int4 dp_smem4[];
  int* sm = reinterpret_cast<int*>(dp_smem4) + warp_in_block() * DP_SMEM_INTS;   // re-derived here so accesses compile to LDS/STS
#endif
  const int lane = lane_id();
  unsigned char* const tbp = TB ? E.T.tb : nullptr;
  if (TB && lane == 0) tbp[0] = 2;                    // cell (0,0): diagonal state
  int* nD = sm; int* naD = sm + DP_SLOTS;            // value / aux handed to cell (i, j+1), stored at j
  int* nI = sm + 2 * DP_SLOTS; int* naI = sm + 3 * DP_SLOTS;   // handed to cell (i+1, j), stored at j
  int* bBase = sm + 4 * DP_SLOTS;                    // best state by diagonal parity: value at +0/+DP_SLOTS, aux at +2/+3*DP_SLOTS
  unsigned char* wA = (unsigned char*)(sm + 8 * DP_SLOTS);
  unsigned char* wB = wA + DP_WIN;                   // shifted by one: byte j holds the base consumed by column j
  const unsigned* wA32 = (const unsigned*)wA; const unsigned* wB32 = (const unsigned*)wB;
  const int breaklen = wmax_i32(P.breaklen), band_cap = wmax_i32(P.band_cap), one = P.one;
  const int max_diff = DP_GOOD * breaklen;
  const int k256 = one << 8;
  wsync();
  // diagonal 0 is the single cell (0,0) with M = 0: its gap children open a gap, its diagonal child reads best = 0
  if (lane == 0) {
    int* b0p = bBase; int* b1p = bBase + DP_SLOTS;
    nD[0] = DP_GOPEN; naD[0] = 0; nI[0] = DP_GOPEN; naI[0] = 0;
    nD[DP_SMASK] = DP_NEG; nI[1] = DP_NEG;                      // halo of diagonal 0
    b0p[0] = 0; b0p[2 * DP_SLOTS] = 0; b0p[DP_SMASK] = DP_NEG; b0p[1] = DP_NEG;
    b1p[DP_SMASK] = DP_NEG; b1p[0] = DP_NEG; b1p[1] = DP_NEG; b1p[2] = DP_NEG;
  }
  int fillA = 0, fillB = 0;
  dp_fill(RS, Av, a0, dirn, N, 0, 0, wA, lane); fillA = DP_FILL;
  dp_fill(QS, Bv, b0, dirn, M, 0, 1, wB, lane); fillB = DP_FILL;
  wsync();
  int high = 0, FinD = 0, FinJ = 0, finAux = 0, lastAux = 0;
  int nlo = (1 - N) > 0 ? (1 - N) : 0, nhi = M < 1 ? M : 1;
  unsigned long long cells = 0;
  int maxw = 0;
  int d;
  for (d = 1; d <= N + M && (d - FinD) <= breaklen && nlo <= nhi; d++) {
    const int w = nhi - nlo + 1;
    cells += (unsigned long long)w;
    if (w > maxw) maxw = w;
    // bases needed on this diagonal: A local index i-1 <= d-nlo-1, B local index j-1 <= nhi-1
    if (d - nlo > fillA) { dp_fill(RS, Av, a0, dirn, N, fillA, 0, wA, lane); fillA += DP_FILL; wsync(); }
    if (nhi > fillB) { dp_fill(QS, Bv, b0, dirn, M, fillB, 1, wB, lane); fillB += DP_FILL; wsync(); }
    int* cB = bBase + (d & 1) * DP_SLOTS; int* cA = cB + 2 * DP_SLOTS;   // holds diagonal d-2, receives diagonal d
    // lane-local best of this diagonal as one key: (score << 8) | (j - nlo); out-of-band cells never enter
    int lkey = -0x7fffffff;
    const int glo = nlo >> 2, ghi = nhi >> 2;
#ifdef MFSC_EMU
    if (w <= MFSC_LANES) {
      int e0, ej;
      DP_PASS1(e0, ej)
      (void)e0; (void)ej;
    } else for (int gt = ghi; gt >= glo; gt -= MFSC_LANES) {
      int e0, e1, e2, e3, ej;
      DP_PASS(gt, e0, e1, e2, e3, ej)
      (void)e0; (void)e1; (void)e2; (void)e3; (void)ej;
    }
#else
    // a band of at most 248 cells touches at most 63 groups: two passes of 32 lanes, the second one rarely needed.
    // The best-state values of both passes stay in registers for the trim below.
    int t0 = DP_NEG, t1 = DP_NEG, t2 = DP_NEG, t3 = DP_NEG, tj = 0, l0 = DP_NEG, l1 = DP_NEG, l2 = DP_NEG, l3 = DP_NEG, lj = 0;
    if (w <= MFSC_LANES) { DP_PASS1(t0, tj) }
    else {
      DP_PASS(ghi, t0, t1, t2, t3, tj)
      if (ghi - glo >= MFSC_LANES) { DP_PASS(ghi - MFSC_LANES, l0, l1, l2, l3, lj) }
    }
#endif
    // diagonal maximum, ties to the larger j; its carried word is read back from the best buffer
    const int dkey = wmax_i32(lkey);
    const int dmax = dkey >> 8;
    const int dj = nlo + (dkey & 0xff);
    const int daux = cA[dj & DP_SMASK];
    lastAux = daux;
    const bool upd = forced || dmax >= high;
    high = upd ? dmax : high; FinD = upd ? d : FinD; FinJ = upd ? dj : FinJ; finAux = upd ? daux : finAux;
    // halo: the next diagonal reads nD[nlo-1 .. nhi] and nI[nlo .. nhi+1]; diagonal d+2 reads best[nlo-1 .. nhi+1]
    {
      const int hidx = ((lane & 1) ? (nhi + 1) : (nlo - 1)) & DP_SMASK;
      int* hb = (lane & 2) ? cB : ((lane & 1) ? nI : nD);
      if (lane < 4 && lane < MFSC_LANES) hb[hidx] = DP_NEG;
#if MFSC_LANES == 1
      nI[(nhi + 1) & DP_SMASK] = DP_NEG; cB[(nlo - 1) & DP_SMASK] = DP_NEG; cB[(nhi + 1) & DP_SMASK] = DP_NEG;
#endif
    }
    // trim cells that fell more than max_diff below the best (narrows the next diagonal only)
    int lo = 0x7fffffff, hi = -0x7fffffff;
    const int thr = high - max_diff;
#ifdef MFSC_EMU
    wsync();
    for (int g = glo + lane; g <= ghi; g += MFSC_LANES) {
      const int4 v = ld4(cB + ((g * 4) & DP_SMASK));
      const int j0 = g * 4;
      unsigned m = (v.x >= thr ? 1u : 0u) | (v.y >= thr ? 2u : 0u) | (v.z >= thr ? 4u : 0u) | (v.w >= thr ? 8u : 0u);
      for (int c = 0; c < 4; c++) if (j0 + c < nlo || j0 + c > nhi) m &= ~(1u << c);   // the narrow pass leaves its neighbours untouched
      if (m) {
        const int a = j0 + dev_ffs32(m) - 1, z = j0 + 31 - dev_clz32(m);
        if (a < lo) lo = a;
        if (z > hi) hi = z;
      }
    }
#else
    {
      const unsigned mt = (t0 >= thr ? 1u : 0u) | (t1 >= thr ? 2u : 0u) | (t2 >= thr ? 4u : 0u) | (t3 >= thr ? 8u : 0u);
      const unsigned ml = (l0 >= thr ? 1u : 0u) | (l1 >= thr ? 2u : 0u) | (l2 >= thr ? 4u : 0u) | (l3 >= thr ? 8u : 0u);
      // the low pass (if any) holds smaller j than the top pass
      const int lo_t = mt ? tj + dev_ffs32(mt) - 1 : 0x7fffffff, lo_l = ml ? lj + dev_ffs32(ml) - 1 : 0x7fffffff;
      const int hi_t = mt ? tj + 31 - dev_clz32(mt) : -0x7fffffff, hi_l = ml ? lj + 31 - dev_clz32(ml) : -0x7fffffff;
      lo = lo_l < lo_t ? lo_l : lo_t;
      hi = hi_t > hi_l ? hi_t : hi_l;
    }
#endif
    lo = wmin_i32(lo); hi = wmax_i32(hi);
    wsync();
    const bool none = lo > hi;
    if (!none && hi - lo + 2 > band_cap) {
      int l2 = lo, h2 = hi;
      while (l2 <= h2 && h2 - l2 + 2 > band_cap) { if (cB[l2 & DP_SMASK] < cB[h2 & DP_SMASK]) l2++; else h2--; }
      lo = wmax_i32(l2); hi = wmax_i32(h2);      // values read from memory: hand them back as uniform
      if (lane == 0) E.acc_caps++;
    }
    // Dead tail (same rule as the oracle's engine and the packed engine, see mfsc_extend_pk.cuh)
    if ((d & 7) == 0 && !forced && FinD + breaklen < N + M && (nhi == M || d - nlo == N)) {
      int ub = -0x40000000;
      for (int j = nlo + lane; j <= nhi; j += MFSC_LANES) {
        const int c = cB[j & DP_SMASK];
        if (c <= DP_NEG) continue;
        const int rem_a = N - (d - j), rem_b = M - j;
        const int v = c + DP_GOOD * (rem_a < rem_b ? rem_a : rem_b);
        ub = v > ub ? v : ub;
      }
      if (wmax_i32(ub) < high) { d++; break; }
    }
    {
      const int a = d + 1 - N, t = d + 1 < M ? d + 1 : M;
      nlo = none ? 1 : (lo > a ? lo : a);
      nhi = none ? 0 : (hi + 1 < t ? hi + 1 : t);
    }
  }
  const int dlast = d - 1;
  DpOut o; o.reached = 0;
  if (dlast == N + M) {
    if (!optimal) { o.reached = 1; FinD = N + M; FinJ = M; finAux = lastAux; }
    else if (FinD == dlast) o.reached = 1;
  }
  o.fi = FinD - FinJ; o.fj = FinJ;
  o.cols = FinD - (finAux >> 16); o.errs = (finAux & 0xffff) + FinD - 2 * (finAux >> 16);
  if (lane == 0) {
    E.acc_cells += cells; E.acc_calls += 1ull;
    if (maxw > E.acc_maxw) E.acc_maxw = maxw;
  }
  o.n_ops = (TB && want_ops) ? dp_traceback(E, FinD, FinJ) : 0;
  return o;
}

#include "mfsc_extend_pk.cuh"

template <bool TB, bool PK>
MFSC_DEV DpOut dp_run(const ExtCtx& E, int dir, int a0, int b0, int dirn, int N, int M, bool forced, bool optimal, bool want_ops) {
  if (PK) return dp_engine_pk<TB>(E, dir, a0, b0, dirn, N, M, forced, optimal, want_ops);
  return dp_engine<TB>(E, dir, a0, b0, dirn, N, M, forced, optimal, want_ops);
}

// ------------------------------------------------------------------------------------------
// token lists of the alignments (delta mode, first lane only)
// ------------------------------------------------------------------------------------------
MFSC_DEV int tok_alloc(const ExtCtx& E, int n) {
  if (E.T.top + 2 + n > E.T.cap) { E.T.bad = 1; return -1; }
  const int c = E.T.top;
  E.T.top += 2 + n;
  E.T.arena[c] = -1; E.T.arena[c + 1] = n;
  return c;
}
MFSC_DEV int tok_run(const ExtCtx& E, int len) {
  const int c = tok_alloc(E, 1);
  if (c >= 0) E.T.arena[c + 2] = len;
  return c;
}
// chunk from the columns of the last traceback (E.T.rev holds them last column first)
MFSC_DEV int tok_from_rev(const ExtCtx& E, int n) {
  const unsigned char* rev = E.T.rev;
  int cnt = 0, run = 0;
  for (int k = n - 1; k >= 0; k--) { if (rev[k] == 0) run++; else { if (run) cnt++; cnt++; run = 0; } }
  if (run) cnt++;
  if (cnt == 0) return -1;
  const int c = tok_alloc(E, cnt);
  if (c < 0) return -1;
  int* t = E.T.arena + c + 2;
  int w = 0; run = 0;
  for (int k = n - 1; k >= 0; k--) { if (rev[k] == 0) run++; else { if (run) t[w++] = run; t[w++] = -(int)rev[k]; run = 0; } }
  if (run) t[w++] = run;
  return c;
}
MFSC_DEV void tok_drop_last_column(const ExtCtx& E, int ci) {      // the DP re-aligns the alignment's last column
  const int t = E.R.tail[ci];
  if (t < 0) { E.T.bad = 1; return; }
  int* last = E.T.arena + t + 2 + E.T.arena[t + 1] - 1;
  if (*last >= 1) *last -= 1; else E.T.bad = 1;
}
MFSC_DEV void tok_drop_first_column(const ExtCtx& E, int ci) {
  const int h = E.R.head[ci];
  if (h < 0) { E.T.bad = 1; return; }
  if (E.T.arena[h + 2] >= 1) E.T.arena[h + 2] -= 1; else E.T.bad = 1;
}
MFSC_DEV void tok_append(const ExtCtx& E, int ci, int chunk) {
  if (chunk < 0) return;
  const int t = E.R.tail[ci];
  if (t < 0) { E.R.head[ci] = chunk; E.R.tail[ci] = chunk; return; }
  E.T.arena[t] = chunk; E.R.tail[ci] = chunk;
}
MFSC_DEV void tok_prepend(const ExtCtx& E, int ci, int chunk) {
  if (chunk < 0) return;
  E.T.arena[chunk] = E.R.head[ci]; E.R.head[ci] = chunk;
  if (E.R.tail[ci] < 0) E.R.tail[ci] = chunk;
}
// delta values of alignment ci: walk the chunks; returns the count, writes when out != nullptr
MFSC_DEV int tok_deltas(const ExtCtx& E, int ci, int* out) {
  int n = 0, run = 0;
  for (int c = E.R.head[ci]; c >= 0; c = E.T.arena[c]) {
    const int m = E.T.arena[c + 1];
    for (int k = 0; k < m; k++) {
      const int t = E.T.arena[c + 2 + k];
      if (t >= 0) run += t;
      else { if (out) out[n] = (t == -1) ? (run + 1) : -(run + 1); n++; run = 0; }
    }
  }
  return n;
}

// ------------------------------------------------------------------------------------------
// cluster walking
// ------------------------------------------------------------------------------------------
MFSC_DEV int c_dir(const ExtCtx& E, int c) { return E.ord[c] >= E.pc_rc0 ? 1 : 0; }
MFSC_DEV int c_nm(const ExtCtx& E, int c) { return E.pc_n[E.ord[c]]; }
MFSC_DEV int c_s1(const ExtCtx& E, int c, int t) { return (int)(E.cm_s1[E.pc_first[E.ord[c]] + t] - E.a_ostart) + 1; }
MFSC_DEV int c_s2(const ExtCtx& E, int c, int t) { return E.cm_s2[E.pc_first[E.ord[c]] + t] + 1; }
MFSC_DEV int c_len(const ExtCtx& E, int c, int t) { return E.cm_len[E.pc_first[E.ord[c]] + t]; }

// the cached row goes back to HBM (all lanes hold the same values; lane 0 writes)
MFSC_DEV void rc_flush(const ExtCtx& E) {
  if (E.c_idx < 0) return;
  if (E.lane == 0) {
    const int k = E.c_idx;
    E.R.sA[k] = E.c_sA; E.R.sB[k] = E.c_sB; E.R.eA[k] = E.c_eA; E.R.eB[k] = E.c_eB; E.R.cols[k] = E.c_cols; E.R.errs[k] = E.c_errs;
  }
  wsync();
  E.c_idx = -1;
}
// make row ci the cached one
MFSC_DEV void rc_bind(const ExtCtx& E, int ci) {
  if (E.c_idx == ci) return;
  rc_flush(E);
  E.c_idx = ci;
  E.c_sA = E.R.sA[ci]; E.c_sB = E.R.sB[ci]; E.c_eA = E.R.eA[ci]; E.c_eB = E.R.eB[ci]; E.c_cols = E.R.cols[ci]; E.c_errs = E.R.errs[ci];
}
MFSC_DEV int r_sA(const ExtCtx& E, int k) { return k == E.c_idx ? E.c_sA : E.R.sA[k]; }
MFSC_DEV int r_sB(const ExtCtx& E, int k) { return k == E.c_idx ? E.c_sB : E.R.sB[k]; }
MFSC_DEV int r_eA(const ExtCtx& E, int k) { return k == E.c_idx ? E.c_eA : E.R.eA[k]; }
MFSC_DEV int r_eB(const ExtCtx& E, int k) { return k == E.c_idx ? E.c_eB : E.R.eB[k]; }

template <bool TB, bool PK>
MFSC_DEV bool ext_forward(ExtCtx& E, int ci, int tA, int tB, bool forced, bool optimal) {
  rc_bind(E, ci);
  const int eA = E.c_eA, eB = E.c_eB, dir = E.R.dir[ci];
  bool overflow = false;
  if (tA - eA + 1 > DP_MAX_ALN) { tA = eA + DP_MAX_ALN - 1; overflow = true; optimal = true; }
  if (tB - eB + 1 > DP_MAX_ALN) { tB = eB + DP_MAX_ALN - 1; overflow = true; optimal = true; }
  DpOut o = dp_run<TB, PK>(E, dir, eA, eB, +1, tA - eA + 1, tB - eB + 1, forced, optimal, true);
  // the DP re-aligns the alignment's last column
  E.c_cols += o.cols - 1; E.c_errs += o.errs;
  E.c_eA = eA + o.fi - 1; E.c_eB = eB + o.fj - 1;
  if (TB) {
    wsync();
    if (E.lane == 0) { tok_drop_last_column(E, ci); tok_append(E, ci, tok_from_rev(E, o.n_ops)); }
    wsync();
  }
  return o.reached && !overflow;
}

template <bool TB, bool PK>
MFSC_DEV bool ext_backward(ExtCtx& E, int ci, int ti) {
  rc_bind(E, ci);
  const int sA = E.c_sA, sB = E.c_sB, dir = E.R.dir[ci];
  int tA, tB; bool optimal = false, overflow = false;
  if (ti >= 0) { tA = r_eA(E, ti); tB = r_eB(E, ti); } else { tA = 1; tB = 1; optimal = true; }
  if (sA - tA + 1 > DP_MAX_ALN) { tA = sA - DP_MAX_ALN + 1; overflow = true; optimal = true; }
  if (sB - tB + 1 > DP_MAX_ALN) { tB = sB - DP_MAX_ALN + 1; overflow = true; optimal = true; }
  DpOut o = dp_run<TB, PK>(E, dir, sA, sB, -1, sA - tA + 1, sB - tB + 1, false, optimal, false);
  bool reached = o.reached;
  if (overflow || ti < 0) reached = false;
  if (reached) {
    ext_forward<TB, PK>(E, ti, sA, sB, true, false);      // binds ti (row ci went back to HBM)
    // the target now ends on this alignment's first column; append the rest and drop this alignment
    E.c_cols += E.R.cols[ci] - 1; E.c_errs += E.R.errs[ci];
    E.c_eA = E.R.eA[ci]; E.c_eB = E.R.eB[ci];
    if (TB) {
      wsync();
      if (E.lane == 0) {
        tok_drop_first_column(E, ci);
        if (E.R.tail[ti] >= 0 && E.R.head[ci] >= 0) { E.T.arena[E.R.tail[ti]] = E.R.head[ci]; E.R.tail[ti] = E.R.tail[ci]; } else E.T.bad = 1;
      }
      wsync();
    }
    E.al1--;
  } else {
    const int nsA = sA - o.fi + 1, nsB = sB - o.fj + 1;
    DpOut o2 = dp_run<TB, PK>(E, dir, nsA, nsB, +1, sA - nsA + 1, sB - nsB + 1, true, false, true);
    rc_bind(E, ci);
    E.c_cols += o2.cols - 1; E.c_errs += o2.errs;
    E.c_sA = nsA; E.c_sB = nsB;
    if (TB) {
      wsync();
      if (E.lane == 0) { tok_drop_first_column(E, ci); tok_prepend(E, ci, tok_from_rev(E, o2.n_ops)); }
      wsync();
    }
  }
  return reached;
}

MFSC_DEV bool ext_shadowed(const ExtCtx& E, int c, int ap) {
  if (E.al1 == E.al0) return false;
  const int nm = c_nm(E, c), dir = c_dir(E, c);
  const int sA = c_s1(E, c, 0), eA = c_s1(E, c, nm - 1) + c_len(E, c, nm - 1) - 1;
  const int sB = c_s2(E, c, 0), eB = c_s2(E, c, nm - 1) + c_len(E, c, nm - 1) - 1;
  for (int k = ap; k >= E.al0; k--)
    if (E.R.dir[k] == dir && r_eA(E, k) >= eA && r_eB(E, k) >= eB && r_sA(E, k) <= sA && r_sB(E, k) <= sB) return true;
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
  const int sA = r_sA(E, cur), sB = r_sB(E, cur), dir = E.R.dir[cur];
  int dist = sA < sB ? sA : sB;
  int res = -1;
  for (int k = cur - 1; k >= E.al0; k--) {
    if (E.R.dir[k] != dir) continue;
    const int eA = r_eA(E, k), eB = r_eB(E, k);
    if (eA <= sA && eB <= sB) {
      int greater, lesser;
      if (sA - eA > sB - eB) { greater = sA - eA; lesser = sB - eB; } else { lesser = sA - eA; greater = sB - eB; }
      if (greater < E.P.breaklen || (long long)lesser * DP_GOOD + (long long)(greater - lesser) * DP_GCONT >= 0) { res = k; break; }
      else if (2ll * greater - lesser < dist) { res = k; dist = (int)(2ll * greater - lesser); }
    }
  }
  return res;
}

// Gaps between consecutive matches of a cluster that stage 4a (mfsc_gaps.cuh) aligns as independent tasks
#define GAP_MIN 32                    // one of the sides longer than this
MFSC_DEV bool gap_is_task(int gN, int gM) {
  return gN >= 1 && gM >= 1 && (gN > GAP_MIN || gM > GAP_MIN) && gN <= DP_MAX_ALN && gM <= DP_MAX_ALN;
}
// A stage 4a result, as stored per start match: x = 1 | reached << 1 | fi << 2 | fj << 17, y = columns | errors << 16,
// z = cells | widest band << 23, w = the target it was computed for (gap_target_key).
MFSC_DEV int gap_target_key(int tA, int tB) { return tA * 31 + tB; }
// The walk stands on row ci and is about to extend forward from (sA, sB) = the end of cluster match m towards (tA, tB):
// if stage 4a has run exactly that DP, its outcome is applied to the row.  Returns 0 (nothing taken), 1 (taken, target
// not reached) or 2 (taken, target reached).
MFSC_DEV int gap_take(ExtCtx& E, int ci, int m, int sA, int sB, int tA, int tB) {
  if (E.gap_res == nullptr || E.pk_fail) return 0;
  rc_bind(E, ci);
  if (E.c_eA != sA || E.c_eB != sB) return 0;
  const int4 r = E.gap_res[m];
  if (!(r.x & 1) || r.w != gap_target_key(tA, tB)) return 0;
  const int fi = (r.x >> 2) & 0x7fff, fj = (r.x >> 17) & 0x7fff;
  E.c_cols += (r.y & 0xffff) - 1; E.c_errs += (r.y >> 16) & 0xffff;
  E.c_eA = sA + fi - 1; E.c_eB = sB + fj - 1;
  if (E.lane == 0) {
    E.acc_cells += (unsigned long long)(r.z & 0x7fffff); E.acc_calls += 1ull; E.acc_gaps += 1;
    const int mw = (r.z >> 23) & 0xff;
    if (mw > E.acc_maxw) E.acc_maxw = mw;
  }
  return (r.x & 2) ? 2 : 1;
}

// Walk the clusters of one (record, read) group in order; returns nothing, alignments are rows [al0, al1)
template <bool TB, bool PK, bool GAPS>
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
    const int first = E.pc_first[E.ord[cur]];         // the cluster's matches: cm_*[first .. first + nm)
    int tl_ok = 0, tl_cols = 0, tl_errs = 0, tl_chunk = -1;          // this lane's small rectangle of the current chunk of gaps
    for (int mi = 0; mi < nm; mi++) {
      const int ms1 = (int)(E.cm_s1[first + mi] - E.a_ostart) + 1, ms2 = E.cm_s2[first + mi] + 1, mlen = E.cm_len[first + mi];
      if (target_reached) {
        rc_bind(E, curA);
        if (E.c_eA != ms1 || E.c_eB != ms2) continue;   // walk to the match the DP landed on
        E.c_eA += mlen - 1; E.c_eB += mlen - 1; E.c_cols += mlen - 1;
        if (TB && mlen > 1) {
          wsync();
          if (E.lane == 0) tok_append(E, curA, tok_run(E, mlen - 1));
          wsync();
        }
      } else {
        rc_flush(E);
        curA = E.al1++;
        wsync();
        if (E.lane == 0) {
          E.R.ref[curA] = E.ref_id; E.R.qry[curA] = E.qry_id; E.R.dir[curA] = dir;
          if (TB) { const int c = tok_run(E, mlen); E.R.head[curA] = c; E.R.tail[curA] = c; }
        }
        wsync();
        E.c_idx = curA; E.c_sA = ms1; E.c_sB = ms2; E.c_eA = ms1 + mlen - 1; E.c_eB = ms2 + mlen - 1; E.c_cols = mlen; E.c_errs = 0;
        const int ti = ext_reverse_target(E, curA);
        if (ext_backward<TB, PK>(E, curA, ti)) curA = ti;
      }
      if (mi < nm - 1) {
        tA = (int)(E.cm_s1[first + mi + 1] - E.a_ostart) + 1; tB = E.cm_s2[first + mi + 1] + 1;
        bool taken = false;
        if (PK && !TB) {
          // the gaps of this cluster, 32 at a time, one small rectangle per lane (dp_tiny_lane)
          if (tl_chunk != mi / MFSC_LANES) {          // (a chunk's first gap may have been skipped by the walk: go by chunk, not by mi)
            tl_chunk = mi / MFSC_LANES;
            const int g = tl_chunk * MFSC_LANES + E.lane;
            tl_ok = 0;
            if (g < nm - 1) {
              const int gl = E.cm_len[first + g];
              const int gA = (int)(E.cm_s1[first + g] - E.a_ostart) + gl, gB = E.cm_s2[first + g] + gl;
              const int gN = (int)(E.cm_s1[first + g + 1] - E.a_ostart) + 1 - gA + 1, gM = E.cm_s2[first + g + 1] + 1 - gB + 1;
              if (gN >= 1 && gM >= 1 && gN <= TL_MAX && gM <= TL_MAX && gN + gM <= pk_tiny_limit(E.P.breaklen, E.P.band_cap))
                tl_ok = dp_tiny_lane(E, dir, gA, gB, gN, gM, &tl_cols, &tl_errs) ? 1 : 0;
            }
          }
          const int src = mi % MFSC_LANES;
          const int ok = wbcast_i32(tl_ok, src), gcols = wbcast_i32(tl_cols, src), gerrs = wbcast_i32(tl_errs, src);
          rc_bind(E, curA);
          const int eA = E.c_eA, eB = E.c_eB;
          if (ok && !E.pk_fail && eA == ms1 + mlen - 1 && eB == ms2 + mlen - 1) {
            const int gN = tA - eA + 1, gM = tB - eB + 1;
            E.c_cols += gcols - 1; E.c_errs += gerrs;
            E.c_eA = tA; E.c_eB = tB;
            if (E.lane == 0) {
              E.acc_cells += (unsigned long long)(gN + 1) * (unsigned long long)(gM + 1) - 1ull; E.acc_calls += 1ull;
              const int mw = (gN < gM ? gN : gM) + 1;
              if (mw > E.acc_maxw) E.acc_maxw = mw;
            }
            target_reached = true;
            taken = true;
          }
        }
        if (GAPS && PK && !TB && !taken && gap_is_task(tA - (ms1 + mlen - 1) + 1, tB - (ms2 + mlen - 1) + 1)) {
          const int got = gap_take(E, curA, first + mi, ms1 + mlen - 1, ms2 + mlen - 1, tA, tB);
          if (got) { target_reached = got > 1; taken = true; }
        }
        if (!taken) target_reached = ext_forward<TB, PK>(E, curA, tA, tB, false, false);
      } else {
        tA = E.A.n; tB = E.B[0].n;
        targetC = ext_forward_target(E, cur, tA, tB);
        int got = 0;
        if (GAPS && PK && !TB) got = gap_take(E, curA, first + mi, ms1 + mlen - 1, ms2 + mlen - 1, tA, tB);
        if (got) target_reached = got > 1;
        else target_reached = ext_forward<TB, PK>(E, curA, tA, tB, false, targetC < 0);
      }
    }
    if (targetC < 0) target_reached = false;
    if (E.lane == 0) E.fused[E.ord[cur]] = 1;
    wsync();
    if (!target_reached) cur = ++prev; else cur = targetC;
  }
  rc_flush(E);
}

struct ExtIn {
  const int* seg;                       // [2*n_reads+1] match ranges of the strands
  const int* n_pc;                      // [2*n_reads]
  const int* pc_first; const int* pc_n; const int* pc_rec;
  const unsigned* cm_s1; const int* cm_s2; const int* cm_len;
  const unsigned* r_ostart;             // [n_ref]
  int* ord; int* key_tmp; int* grp; unsigned char* fused;   // scratch (per-match slices)
  // reads to extend: all of them (list == nullptr) or list[0 .. *n_list).  The packed-engine instantiation appends the
  // reads it has to hand over to the general engine to retry[0 .. *n_retry).
  const int* list; const unsigned* n_list; int* retry; unsigned* n_retry;
  const int4* gap_res;                  // stage 4a results (mfsc_gaps.cuh) for the walk to pick up, or nullptr
  int ord_ready;                        // stage 4a has already written the walking order (ext_order) of every read
};

// Walking order of the clusters of one read (both strands), into I.ord[b0 .. b0 + n0 + n1): reference records in order of first
// appearance in the cluster stream (the order of the delta file's `>ref qry` blocks), then start in the record, strand,
// start in the read, emission order.  Also clears the clusters' `fused` flags.
MFSC_DEV void ext_order(const ExtIn& I, int b0, int b1, int n0, int n1, int lane) {
  const int nC = n0 + n1;
  int* ord = I.ord + b0;
  int* tmp = I.key_tmp + b0;
  // split cluster ids of the read: forward strand b0..b0+n0, reverse strand b1..b1+n1.
  // Walking order: reference records in order of first appearance in that stream (the order of the
  // delta file's `>ref qry` blocks), then start in the record, strand, start in the read, emission order.
  for (int x = lane; x < nC; x += MFSC_LANES) {
    const int pc = x < n0 ? b0 + x : b1 + (x - n0);
    tmp[x] = pc;
    I.fused[pc] = 0;
  }
  wsync();
  for (int x = lane; x < nC; x += MFSC_LANES) {
    const int rec = I.pc_rec[tmp[x]];
    int f = x;
    for (int y = 0; y < x; y++) if (I.pc_rec[tmp[y]] == rec) { f = y; break; }
    I.grp[b0 + x] = f;
  }
  wsync();
  for (int x = lane; x < nC; x += MFSC_LANES) {
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
  wsync();
}

#ifdef MFSC_EMU
static thread_local int emu_dp_smem[DP_SMEM_INTS];
#endif

#ifndef MFSC_EXT_MINBLOCKS
#define MFSC_EXT_MINBLOCKS 6
#endif
#ifndef MFSC_EXT_MINBLOCKS_PK
#define MFSC_EXT_MINBLOCKS_PK 7
#endif
struct TbPool { unsigned char* tb; unsigned char* rev; int* arena; int arena_ints; };

template <bool TB, bool PK, bool GAPS = false>
MFSC_KERNEL_LB(128, PK ? MFSC_EXT_MINBLOCKS_PK : MFSC_EXT_MINBLOCKS) k_extend(SeqStore RS, SeqStore QS, ExtIn I, int n_reads, DevParams P, RowsOut R, ExtStats st,
                                                   unsigned* ticket, TbPool pool, DeltaOut DO) {
#ifdef MFSC_EMU
  int* sm = emu_dp_smem;
#else
  extern __shared__ int4 dp_smem4[];
  int* sm = reinterpret_cast<int*>(dp_smem4) + warp_in_block() * (PK ? PK_SMEM_INTS : DP_SMEM_INTS);
#endif
  const int lane = lane_id();
  unsigned long long w_cells = 0, w_calls = 0, w_caps = 0, w_gaps = 0; int w_maxw = 0;
  const int n_work = I.list ? (int)*I.n_list : n_reads;
  // reads are handed out by a ticket counter: a warp takes the next read when it is done with its own
  for (;;) {
    int q = 0;
    if (lane == 0) q = (int)atomic_add_u32(ticket, 1u);
    q = wbcast_i32(q, 0);
    if (q >= n_work) break;
    if (I.list) q = I.list[q];
    const int b0 = I.seg[2 * q], b1 = I.seg[2 * q + 1];
    const int n0 = I.n_pc[2 * q], n1 = I.n_pc[2 * q + 1];
    const int nC = n0 + n1;
    if (nC == 0) { if (lane == 0) R.n_rows[q] = 0; continue; }
    int* ord = I.ord + b0;
    if (!GAPS || !I.ord_ready) ext_order(I, b0, b1, n0, n1, lane);
    ExtCtx E;
    E.RS = RS; E.QS = QS; E.P = P; E.sm = sm; E.st = st; E.R = R; E.lane = lane;
    E.acc_cells = 0; E.acc_calls = 0; E.acc_maxw = 0; E.acc_caps = 0; E.acc_gaps = 0;
    E.T.tb = nullptr; E.T.rev = nullptr; E.T.arena = nullptr; E.T.cap = 0; E.T.top = 0; E.T.bad = 0;
    E.pk_fail = 0;
    E.c_idx = -1; E.c_sA = E.c_sB = E.c_eA = E.c_eB = E.c_cols = E.c_errs = 0;
    if (TB) {
      const long long wg = warp_global();
      E.T.tb = pool.tb + (size_t)wg * TB_DIAGS * DP_SLOTS;
      E.T.rev = pool.rev + (size_t)wg * TB_DIAGS;
      E.T.arena = pool.arena + (size_t)wg * pool.arena_ints; E.T.cap = pool.arena_ints;
    }
    E.B[0].base = QS.start[2 * q]; E.B[0].n = QS.len[2 * q];
    E.B[1].base = QS.start[2 * q + 1]; E.B[1].n = QS.len[2 * q + 1];
    E.pc_rc0 = b1; E.pc_first = I.pc_first; E.pc_n = I.pc_n; E.fused = I.fused;
    E.gap_res = GAPS ? I.gap_res : nullptr;
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
      ext_clusters<TB, PK, GAPS>(E);
      rows = E.al1 - row0;
      g0 = g1;
    }
    wsync();
    if (PK && E.pk_fail) {
      // a DP of this read did not fit the packed words: nothing of it is kept, the general engine extends it again
      if (lane == 0) { R.n_rows[q] = 0; I.retry[atomic_add_u32(I.n_retry, 1u)] = q; }
      continue;
    }
    // reverse-strand rows are reported on the forward read: position p -> qlen - p + 1
    const int qlen = QS.len[2 * q];
    for (int x = lane; x < rows; x += MFSC_LANES) {
      if (R.dir[row0 + x]) { R.sB[row0 + x] = qlen - R.sB[row0 + x] + 1; R.eB[row0 + x] = qlen - R.eB[row0 + x] + 1; }
    }
    if (lane == 0) R.n_rows[q] = rows;
    if (TB) {
      // delta values of the read's alignments go to the batch-wide buffer
      wsync();
      if (lane == 0) {
        for (int x = 0; x < rows; x++) {
          const int r = row0 + x;
          const int n = tok_deltas(E, r, nullptr);
          const long long off = (long long)atomic_add_u64(DO.used, (unsigned long long)n);
          if (E.T.bad || off + n > DO.cap) { R.dl_begin[r] = -1; R.dl_end[r] = -1; atomic_add_u32(DO.n_bad, 1u); }
          else { tok_deltas(E, r, DO.buf + off); R.dl_begin[r] = (int)off; R.dl_end[r] = (int)(off + n); }
        }
      }
      wsync();
    }
    w_cells += E.acc_cells; w_calls += E.acc_calls; w_caps += (unsigned long long)E.acc_caps; w_gaps += (unsigned long long)E.acc_gaps;
    if (E.acc_maxw > w_maxw) w_maxw = E.acc_maxw;
  }
  if (lane == 0 && w_calls) {
    atomic_add_u64(st.cells, w_cells);
    atomic_add_u64(st.calls, w_calls);
    atomic_max_i32(st.max_band, w_maxw);
    if (w_caps) atomic_add_u64(st.caps, w_caps);
    if (GAPS && w_gaps) atomic_add_u64(st.gaps, w_gaps);
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
      if (S.dl_begin) { D.dl_begin[dst + x] = S.dl_begin[src + x]; D.dl_end[dst + x] = S.dl_end[src + x]; }
    }
  }
}
