// mfsc_extend_pk.cuh -- the packed DP engine of stage 4 (included by mfsc_extend.cuh).
//
// Same recurrence, band rule, tie rule and result as dp_engine in mfsc_extend.cuh, with every DP state held in ONE
// 32-bit word so that a three-way selection is one integer maximum instead of a compare and two selects:
//
//     bits 31..19  score (signed, relative: the running best is kept near PK_H0 by periodic rebasing)
//     bits 18..17  state priority, only meaningful inside a comparison (M 2, I 1, D 0: the tie rule of the
//                  reference DP -- candidates of one comparison never share (score, priority), so the low bits
//                  never decide)
//     bits 16..7   u  = columns of the path that consume B only (relative to the rebasing offset)
//     bits  6..0   mm = mismatching diagonal columns of the path (relative)
//
// At the finish cell (i, j): diagonal columns = j - u, so columns = i + u and errors = mm + i - j + 2u.
// Every PK_PERIOD anti-diagonals the live words are rebased: the score field by the movement of the running best,
// u and mm by their minimum over the live band (a path can add at most PK_PERIOD to u and PK_PERIOD/2 to mm in
// between).  If the spread of u or mm over the band, or the depth of an in-band score below the best, does not
// fit its field, the engine gives up: the read is then extended again by the general engine (k_extend<TB, false>),
// so the packing never costs exactness.  The same happens when a DP meets a base that is not ACGT: two such bases
// score as a mismatch without counting as an error, which this engine's single per-cell flag does not express.
// Band state: 4 KB + 1 KB of base windows per warp.
#pragma once

#define PK_SB 19
#define PK_K (1 << PK_SB)                       // one score unit
#define PK_P (1 << 17)                          // one priority unit
#define PK_U1 (1 << 7)                          // one B-only column
#define PK_CLEAN (~(3 << 17))                   // clears the priority bits
#define PK_SCORE ((int)0xfff80000)              // the score field
#define PK_UF(w) (((w) >> 7) & 0x3ff)
#define PK_MF(w) ((w) & 0x7f)
#define PK_H0 3000                              // field value the running best is rebased to
#define PK_SENT (-3500 * PK_K)                  // "unreachable"
#define PK_LIVE (-3000 * PK_K)                  // a word of the previous best buffer below this is a halo sentinel
#define PK_FLOOR (-2000)                        // in-band cells must be at least this (field value) at every rebase
#define PK_PERIOD 64
#define PK_U_SPAN (1023 - PK_PERIOD - 8)        // widest spread of u / mm over the live band a rebase accepts
#define PK_M_SPAN (127 - PK_PERIOD / 2 - 4)
#define PK_MATCH (DP_GOOD * PK_K + 2 * PK_P)
#define PK_MISM (DP_BAD * PK_K + 2 * PK_P + 1)   // a mismatching pair: score, priority of M, one more mismatch
#define PK_SMEM_INTS (4 * DP_SLOTS + 2 * DP_WIN / 4 + 4)     // state arrays, base windows, finish-cell record

MFSC_DEV int pk_max(int a, int b) { return a > b ? a : b; }

// Ordering of the band arrays inside one anti-diagonal.  The lanes of a warp run this loop converged: every branch in it is on a
// warp-uniform condition, per-lane differences are predicated stores, and each anti-diagonal passes a warp vote.  A converged warp
// issues its shared-memory instructions in program order for all lanes, so between the loads and the stores of a pass the compiler
// must be kept from reordering, but no warp barrier has to be issued (__syncwarp costs three issue slots here, five times per
// anti-diagonal: the divergence test, its mask and the barrier).  -DMFSC_PK_HARDSYNC=1 restores the barriers.
#if defined(MFSC_EMU) || MFSC_PK_HARDSYNC
#define PK_ORDER() wsync()
#else
#define PK_ORDER() asm volatile("" ::: "memory")
#endif

// dp_fill that also reports whether this lane decoded a base of the sequence proper that is not ACGT
MFSC_DEV bool pk_fill(const SeqStore& S, SeqView V, int s0, int dirn, int n_local, int from, int off, unsigned char* win, int lane) {
  const int to = from + DP_FILL < n_local + 8 ? from + DP_FILL : n_local + 8;
  bool bad = false;
  for (int t = from + lane; t < to; t += MFSC_LANES) {
    int c = 4;
    if (t < n_local) { c = code_at(S, V.base + (unsigned)(s0 + dirn * t - 1)); bad |= c > 3; }
    win[(t + off) & DP_WMASK] = (unsigned char)c;
  }
  return bad;
}

// one cell: DV = D state handed over by (i, j-1) (priority 0), IV = I state handed over by (i-1, j) (priority 1),
// RB = best state of (i-1, j-1) (clean).  OB best state (clean), OD / OI what the two gap children receive.
//   best        : max3(D, I, M)                               M wins ties, then I
//   towards j+1 : max(D + cont, best + open), one more B-only column;  the open candidate carries priority >= 1
//   towards i+1 : max(I + cont, best + open);  the open candidate keeps the priority of the best state, so on a
//                 tie the extension wins exactly when the best state is D
#define PK_CELL(C, DV, IV, RB, OB, OD, OI, OT)                                                             \
  {                                                                                                        \
    int m_ = FADD(RB, PK_MATCH);                                                                           \
    if ((xm & (0xffu << (8 * C))) != 0u) m_ = FADD(m_, PK_MISM - PK_MATCH);                                \
    const int bu_ = pk_max(pk_max(DV, IV), m_);                                                            \
    OB = bu_ & kclean;                                                                                     \
    const int c_ = FADD(bu_, DP_GOPEN * PK_K);                                                             \
    const int cd_ = FADD(c_, PK_U1 + PK_P);                                                                \
    const int e_ = FADD(DV, DP_GCONT * PK_K + PK_U1);                                                      \
    const int f_ = FADD(IV, DP_GCONT * PK_K);                                                              \
    const int du_ = pk_max(e_, cd_), iu_ = pk_max(f_, c_);                                                 \
    OD = du_ & kclean; OI = (iu_ & kclean) | kprio;                                                        \
    /* traceback code as in DP_CELL: bits 0-1 best state, bit 2 D extended, bit 3 I extended */           \
    if (TB) OT = (((unsigned)bu_ >> 17) & 3u) | ((((unsigned)du_ >> 17) & 3u) == 0u ? 4u : 0u) | ((((unsigned)iu_ >> 17) & 3u) == 1u ? 8u : 0u); \
  }

// One pass of the lanes over 32 groups of four cells, highest group GT on lane 0.  Lanes below the band shadow the
// lowest group: they compute what its lane computes and only their stores are suppressed, so the pass is branch-free.
#define PK_PASS(GT, B0, B1, B2, B3, BJ)                                                                                  \
  {                                                                                                                      \
    const bool actg = (GT) - lane >= glo;                                                                                \
    const int g = actg ? (GT) - lane : glo;                                                                              \
    const int j0 = g * 4;                                                                                                \
    const int y = j0 & DP_SMASK, y1 = (j0 - 1) & DP_SMASK;                                                               \
    const int4 vD = ld4(nD + y), vI = ld4(nI + y), vB = ld4(cB + y);                                                     \
    const int sD = nD[y1], sB = cB[y1];                                                                                  \
    const int as = d - j0 - 4;                                                                                           \
    const unsigned w0 = wA32[(as >> 2) & (DP_WIN / 4 - 1)], w1 = wA32[((as >> 2) + 1) & (DP_WIN / 4 - 1)];               \
    const unsigned aw = dev_funnel_r(w0, w1, (unsigned)(as & 3) * 8u);                                                   \
    const unsigned bw = wB32[(j0 >> 2) & (DP_WIN / 4 - 1)];                                                              \
    const unsigned awr = dev_byte_rev(aw);                                                                               \
    const unsigned xm = awr ^ bw;               /* per byte (= cell): non-zero where the two bases differ */             \
    PK_ORDER();                                                                                                             \
    int o0b, o0d, o0i, o1b, o1d, o1i, o2b, o2d, o2i, o3b, o3d, o3i;                                                      \
    unsigned o0t = 0, o1t = 0, o2t = 0, o3t = 0;                                                                         \
    PK_CELL(3, vD.z, vI.w, vB.z, o3b, o3d, o3i, o3t)                                                                     \
    PK_CELL(2, vD.y, vI.z, vB.y, o2b, o2d, o2i, o2t)                                                                     \
    PK_CELL(1, vD.x, vI.y, vB.x, o1b, o1d, o1i, o1t)                                                                     \
    PK_CELL(0, sD, vI.x, sB, o0b, o0d, o0i, o0t)                                                                         \
    const int r0 = j0 - nlo;                                                                                             \
    if ((unsigned)(r0 + 3) >= (unsigned)w) o3b = PK_SENT;                                                                \
    if ((unsigned)(r0 + 2) >= (unsigned)w) o2b = PK_SENT;                                                                \
    if ((unsigned)(r0 + 1) >= (unsigned)w) o1b = PK_SENT;                                                                \
    if ((unsigned)r0 >= (unsigned)w) o0b = PK_SENT;                                                                      \
    if (actg) {                                                                                                          \
      st4(cB + y, o0b, o1b, o2b, o3b); st4(nD + y, o0d, o1d, o2d, o3d); st4(nI + y, o0i, o1i, o2i, o3i);                 \
      if (TB) *reinterpret_cast<unsigned*>(tbp + (size_t)d * DP_SLOTS + y) = o0t | (o1t << 8) | (o2t << 16) | (o3t << 24); \
    }                                                                                                                    \
    /* lane-local best as (score, cell): ties to the larger j */                                                         \
    const int kl_ = pk_max(pk_max(pk_max((o3b & kscore) | 3, (o2b & kscore) | 2), (o1b & kscore) | 1), o0b & kscore);    \
    lkey = pk_max(lkey, kl_ + r0);                                                                                       \
    B0 = o0b; B1 = o1b; B2 = o2b; B3 = o3b; BJ = j0;                                                                     \
    PK_ORDER();                                                                                                             \
  }

// One pass of one cell per lane over the CNT cells below and including TOP (CNT <= 32, the lowest one is nlo).
#define PK_PASS1(TOP, CNT, B0, BJ)                                                                                       \
  {                                                                                                                      \
    const bool act1 = lane < (CNT);                                                                                      \
    const int j1 = act1 ? (TOP) - lane : nlo;        /* idle lanes shadow the lowest cell */                             \
    const int y = j1 & DP_SMASK, y1 = (j1 - 1) & DP_SMASK;                                                               \
    const int sD = nD[y1], sI = nI[y], sB = cB[y1];                                                                      \
    const unsigned ca1 = wA[(d - j1 - 1) & DP_WMASK], cb1 = wB[j1 & DP_WMASK];                                           \
    const unsigned xm = ca1 ^ cb1;                                                                                       \
    PK_ORDER();                                                                                                             \
    int ob, od, oi; unsigned ot = 0;                                                                                     \
    PK_CELL(0, sD, sI, sB, ob, od, oi, ot)                                                                               \
    if (act1) {                                                                                                          \
      cB[y] = ob; nD[y] = od; nI[y] = oi;                                                                                \
      if (TB) tbp[(size_t)d * DP_SLOTS + y] = (unsigned char)ot;                                                         \
    }                                                                                                                    \
    lkey = pk_max(lkey, (ob & kscore) + (j1 - nlo));                                                                     \
    B0 = ob; BJ = j1;                                                                                                    \
    PK_ORDER();                                                                                                             \
  }

// ------------------------------------------------------------------------------------------
// small rectangles
// ------------------------------------------------------------------------------------------
// A DP over N x M bases with N + M <= pk_tiny_limit(...) can neither trim its band (no cell can fall 3 * breaklen below
// the best: the best is at most 3 * floor(d/2) on anti-diagonal d and every cell is at least -6 - 4d, two gaps), nor stop
// early (fewer anti-diagonals than breaklen), nor meet the band cap: the band is the full rectangle, [max(0, d - N),
// min(d, M)] on anti-diagonal d.  When the caller takes the far corner wherever the best score is (every DP that is not
// `optimal`, and every `forced` one), nothing has to be tracked per anti-diagonal at all: the engine below fills the
// rectangle, one cell per lane and pass, on arrays pre-filled with sentinels (so the borders need no halo), and reads
// the corner word.  Cells and widest band follow in closed form.  This is the DP between consecutive matches of a
// cluster on accurate reads (a few bases each), where the per-diagonal bookkeeping of the banded engine is most of the work.
MFSC_DEV int pk_tiny_limit(int breaklen, int band_cap) {
  int t = ((DP_GOOD * breaklen - 6) * 2) / 11;
  if (t > breaklen) t = breaklen;
  if (t > band_cap - 3) t = band_cap - 3;
  if (t > DP_FILL - 8) t = DP_FILL - 8;          // one window fill per sequence, and fewer slots than the arrays hold
  return t;
}

template <bool TB>
MFSC_DEV DpOut dp_engine_tiny(const ExtCtx& E, int dir, int a0, int b0, int dirn, int N, int M, bool want_ops, int* sm) {
  DpOut o; o.reached = 1; o.fi = N; o.fj = M; o.cols = 0; o.errs = 0; o.n_ops = 0;
  const int lane = lane_id();
  const DevParams& P = E.P;
  unsigned char* const tbp = TB ? E.T.tb : nullptr;
  if (TB && lane == 0) tbp[0] = 2;
  int* nD = sm; int* nI = sm + DP_SLOTS; int* bBase = sm + 2 * DP_SLOTS;
  unsigned char* wA = (unsigned char*)(sm + 4 * DP_SLOTS);
  unsigned char* wB = wA + DP_WIN;
  const int one = P.one, kclean = P.pk_clean, kprio = P.pk_prio;
  wsync();
  // every slot a border cell may read: j in [-1, M + 1] of all four arrays
  for (int j = -1 + lane; j <= M + 1; j += MFSC_LANES) {
    const int y = j & DP_SMASK;
    nD[y] = PK_SENT; nI[y] = PK_SENT; bBase[y] = PK_SENT; bBase[DP_SLOTS + y] = PK_SENT;
  }
  bool bad = pk_fill(E.RS, E.A, a0, dirn, N, 0, 0, wA, lane);
  bad |= pk_fill(E.QS, E.B[dir], b0, dirn, M, 0, 1, wB, lane);
  if (wmax_i32(bad ? 1 : 0)) { E.pk_fail = 1; o.reached = 0; o.fi = 1; o.fj = 1; o.cols = 1; return o; }
  wsync();
  if (lane == 0) {
    nD[0] = (PK_H0 + DP_GOPEN) * PK_K + PK_U1; nI[0] = ((PK_H0 + DP_GOPEN) * PK_K) | PK_P;
    bBase[0] = PK_H0 * PK_K;
  }
  wsync();
  for (int d = 1; d <= N + M; d++) {
    const int nlo = d - N > 0 ? d - N : 0, nhi = d < M ? d : M;
    int* cB = bBase + (d & 1) * DP_SLOTS;
    for (int top = nhi; top >= nlo; top -= MFSC_LANES) {
      const int j1r = top - lane;
      const bool act1 = j1r >= nlo;
      const int j1 = act1 ? j1r : nlo;                 // idle lanes shadow the lowest cell
      const int y = j1 & DP_SMASK, y1 = (j1 - 1) & DP_SMASK;
      const int sD = nD[y1], sI = nI[y], sB = cB[y1];
      const unsigned ca1 = wA[(d - j1 - 1) & DP_WMASK], cb1 = wB[j1 & DP_WMASK];
      const unsigned xm = ca1 ^ cb1;
      wsync();
      int ob, od, oi; unsigned ot = 0;
      PK_CELL(0, sD, sI, sB, ob, od, oi, ot)
      if (act1) {
        cB[y] = ob; nD[y] = od; nI[y] = oi;
        if (TB) tbp[(size_t)d * DP_SLOTS + y] = (unsigned char)ot;
      }
      wsync();
    }
  }
  const int cw = (bBase + ((N + M) & 1) * DP_SLOTS)[M & DP_SMASK];
  const int u = PK_UF(cw), mm = PK_MF(cw);
  o.cols = N + u; o.errs = mm + N - M + 2 * u;
  if (lane == 0) {
    E.acc_cells += (unsigned long long)(N + 1) * (unsigned long long)(M + 1) - 1ull; E.acc_calls += 1ull;
    const int mw = (N < M ? N : M) + 1;
    if (mw > E.acc_maxw) E.acc_maxw = mw;
  }
  if (TB && want_ops) o.n_ops = dp_traceback(E, N + M, M);
  return o;
}

// ------------------------------------------------------------------------------------------
// small rectangles, one per LANE
// ------------------------------------------------------------------------------------------
// On accurate reads most DP calls are the gaps between consecutive matches of a cluster: one or two differing bases, a
// rectangle of a few bases a side.  Their start cell is the end of the match before them and their target the start of the
// next match -- both known from the cluster alone -- so the gaps of a cluster do not wait for one another: the lanes fill 32
// of them at a time, each lane its own rectangle with the same packed cell, the same initial words and the same sentinel
// borders as dp_engine_tiny, in registers / local memory.  The walk (ext_clusters) then takes the finished (columns, errors)
// where it would have called the warp-wide engine; anything that does not fit (sides over TL_MAX, a base that is not ACGT,
// delta mode) takes the normal call.
#define TL_MAX 7
#define TL_SLOTS (TL_MAX + 3)

MFSC_DEV bool dp_tiny_lane(const ExtCtx& E, int dir, int a0, int b0, int N, int M, int* cols, int* errs) {
  constexpr bool TB = false;
  const DevParams& P = E.P;
  const int one = P.one, kclean = P.pk_clean, kprio = P.pk_prio;
  unsigned long long ac = 0, bc = 0;      // 3 bits per base
  bool bad = false;
  for (int t = 0; t < N; t++) { const int c = code_at(E.RS, E.A.base + (unsigned)(a0 + t - 1)); bad |= c > 3; ac |= (unsigned long long)c << (3 * t); }
  for (int t = 0; t < M; t++) { const int c = code_at(E.QS, E.B[dir].base + (unsigned)(b0 + t - 1)); bad |= c > 3; bc |= (unsigned long long)c << (3 * t); }
  if (bad) return false;
  int nD[TL_SLOTS], nI[TL_SLOTS], B0[TL_SLOTS], B1[TL_SLOTS];     // slot j + 1 holds column j, j = -1 .. M + 1
#pragma unroll
  for (int y = 0; y < TL_SLOTS; y++) { nD[y] = PK_SENT; nI[y] = PK_SENT; B0[y] = PK_SENT; B1[y] = PK_SENT; }
  nD[1] = (PK_H0 + DP_GOPEN) * PK_K + PK_U1; nI[1] = ((PK_H0 + DP_GOPEN) * PK_K) | PK_P;
  B0[1] = PK_H0 * PK_K;
  for (int d = 1; d <= N + M; d++) {
    const int lo = d - N > 0 ? d - N : 0, hi = d < M ? d : M;
    int* cB = (d & 1) ? B1 : B0;
    for (int j = hi; j >= lo; j--) {          // descending: every input is read before the cell that overwrites it
      const int i = d - j;
      const int sD = nD[j], sI = nI[j + 1], sB = cB[j];
      const unsigned ca = i >= 1 ? (unsigned)((ac >> (3 * (i - 1))) & 7ull) : 4u, cb = j >= 1 ? (unsigned)((bc >> (3 * (j - 1))) & 7ull) : 5u;
      const unsigned xm = ca ^ cb;
      int ob, od, oi; unsigned ot = 0;
      PK_CELL(0, sD, sI, sB, ob, od, oi, ot)
      (void)ot;
      cB[j + 1] = ob; nD[j + 1] = od; nI[j + 1] = oi;
    }
  }
  const int cw = (((N + M) & 1) ? B1 : B0)[M + 1];
  const int u = PK_UF(cw), mm = PK_MF(cw);
  *cols = N + u; *errs = mm + N - M + 2 * u;
  return true;
}

// Returns what dp_engine returns.  When a rebase finds a field out of range E.pk_fail is set and the result is void.
template <bool TB>
MFSC_NOINLINE DpOut dp_engine_pk(const ExtCtx& E, int dir, int a0, int b0, int dirn, int N_, int M_, bool forced_, bool optimal, bool want_ops) {
  DpOut o; o.reached = 0; o.fi = 1; o.fj = 1; o.cols = 1; o.errs = 0; o.n_ops = 0;
  if (E.pk_fail) return o;
  const int N = wmax_i32(N_), M = wmax_i32(M_);
  const bool forced = wmax_i32(forced_ ? 1 : 0) != 0;
  const SeqStore& RS = E.RS; const SeqStore& QS = E.QS; const SeqView Av = E.A; const SeqView Bv = E.B[dir];
  const DevParams& P = E.P;
#ifdef MFSC_EMU
  int* sm = E.sm;
#else
  extern __shared__ int4 dp_smem4[];
  int* sm = reinterpret_cast<int*>(dp_smem4) + warp_in_block() * PK_SMEM_INTS;
#endif
  const int lane = lane_id();
  unsigned char* const tbp = TB ? E.T.tb : nullptr;
  if (TB && lane == 0) tbp[0] = 2;
  int* nD = sm; int* nI = sm + DP_SLOTS;             // handed to cell (i, j+1) / (i+1, j), stored at j
  int* bBase = sm + 2 * DP_SLOTS;                    // best state by diagonal parity
  unsigned char* wA = (unsigned char*)(sm + 4 * DP_SLOTS);
  unsigned char* wB = wA + DP_WIN;
  const unsigned* wA32 = (const unsigned*)wA; const unsigned* wB32 = (const unsigned*)wB;
  const int breaklen = wmax_i32(P.breaklen), band_cap = wmax_i32(P.band_cap), one = P.one;
  if (N + M <= pk_tiny_limit(breaklen, band_cap) && (forced || !optimal)) return dp_engine_tiny<TB>(E, dir, a0, b0, dirn, N, M, want_ops, sm);
  const int u_span = wmax_i32(P.pk_u_span), m_span = wmax_i32(P.pk_m_span);
  const int kclean = P.pk_clean, kprio = P.pk_prio, kscore = P.pk_score;     // PK_CLEAN, PK_P, PK_SCORE as run-time values: one LOP3 for (x & clean) | prio
  const int max_diff = DP_GOOD * breaklen;
  wsync();
  if (lane == 0) {
    int* b0p = bBase; int* b1p = bBase + DP_SLOTS;
    nD[0] = (PK_H0 + DP_GOPEN) * PK_K + PK_U1; nI[0] = ((PK_H0 + DP_GOPEN) * PK_K) | PK_P;
    nD[DP_SMASK] = PK_SENT; nI[1] = PK_SENT;
    b0p[0] = PK_H0 * PK_K; b0p[DP_SMASK] = PK_SENT; b0p[1] = PK_SENT;
    b1p[DP_SMASK] = PK_SENT; b1p[0] = PK_SENT; b1p[1] = PK_SENT; b1p[2] = PK_SENT;
  }
  int fillA = 0, fillB = 0;
  bool bad = pk_fill(RS, Av, a0, dirn, N, 0, 0, wA, lane); fillA = DP_FILL;
  bad |= pk_fill(QS, Bv, b0, dirn, M, 0, 1, wB, lane); fillB = DP_FILL;
  if (wmax_i32(bad ? 1 : 0)) { E.pk_fail = 1; return o; }
  wsync();
  // The finish cell is recorded in shared memory whenever the best score is matched or beaten: (j, word, rebasing
  // offsets at that time).  The word of the last diagonal's best cell is read back from the best buffer at the end.
  int* rec = sm + 4 * DP_SLOTS + 2 * DP_WIN / 4;
  if (lane == 0) { rec[0] = 0; rec[1] = PK_H0 * PK_K; rec[2] = 0; rec[3] = 0; }
  // halo writer lanes: lane 0 nD[nlo-1], lane 1 nI[nhi+1], lanes 2 and 3 best[nlo-1] and best[nhi+1]
  int* const halo_base = lane < 2 ? (lane == 0 ? nD : nI) : bBase;
  const int halo_par = lane >= 2 ? DP_SLOTS : 0, halo_hi = lane & 1;
  int highF = PK_H0, FinD = 0;
  int accU = 0, accM = 0;                          // what the rebases have subtracted from the u / mm fields so far
  int nlo = (1 - N) > 0 ? (1 - N) : 0, nhi = M < 1 ? M : 1;
  unsigned cells = 0;                              // at most 2 * DP_MAX_ALN diagonals of < 256 cells
  int maxw = 0;
  int d;
  // the walk ends after anti-diagonal dEnd = min(N + M, FinD + breaklen): kept up to date where FinD changes, so the loop tests one
  // limit; the two window-refill tests are one countdown (the window fronts move by at most one position per anti-diagonal)
  int dEnd = N + M < breaklen ? N + M : breaklen;
  int safe = (fillA - (1 - nlo)) < (fillB - nhi) ? (fillA - (1 - nlo)) : (fillB - nhi);
  for (d = 1; d <= dEnd && nlo <= nhi; d++) {
    const int w = nhi - nlo + 1;
    cells += (unsigned)w;
    if (w > maxw) maxw = w;
    if (safe < 0) {
      if (d - nlo > fillA) {
        if (wmax_i32(pk_fill(RS, Av, a0, dirn, N, fillA, 0, wA, lane) ? 1 : 0)) { E.pk_fail = 1; return o; }
        fillA += DP_FILL; wsync();
      }
      if (nhi > fillB) {
        if (wmax_i32(pk_fill(QS, Bv, b0, dirn, M, fillB, 1, wB, lane) ? 1 : 0)) { E.pk_fail = 1; return o; }
        fillB += DP_FILL; wsync();
      }
      safe = (fillA - (d - nlo)) < (fillB - nhi) ? (fillA - (d - nlo)) : (fillB - nhi);
    }
    safe--;
    int* cB = bBase + (d & 1) * DP_SLOTS;
    int lkey = -0x7fffffff;
    const int glo = nlo >> 2, ghi = nhi >> 2;
#ifdef MFSC_EMU
    if (w <= MFSC_LANES) {
      int e0, ej;
      PK_PASS1(nhi, w, e0, ej)
      (void)e0; (void)ej;
    } else for (int gt = ghi; gt >= glo; gt -= MFSC_LANES) {
      int e0, e1, e2, e3, ej;
      PK_PASS(gt, e0, e1, e2, e3, ej)
      (void)e0; (void)e1; (void)e2; (void)e3; (void)ej;
    }
#else
    int t0 = PK_SENT, t1 = PK_SENT, t2 = PK_SENT, t3 = PK_SENT, tj = 0, l0 = PK_SENT, l1 = PK_SENT, l2 = PK_SENT, l3 = PK_SENT, lj = 0;
    // a band of at most 248 cells touches at most 63 groups: one pass of 32 lanes x 4 cells, and for what is left below
    // it a pass of one cell per lane when that is at most 32 cells (bands up to ~160 cells), else a second four-cell pass
    const bool two = ghi - glo >= MFSC_LANES;
    const int low_top = 4 * (ghi - MFSC_LANES + 1) - 1, low_cnt = low_top - nlo + 1;
    const bool low1 = low_cnt <= MFSC_LANES;
    if (w <= MFSC_LANES) { PK_PASS1(nhi, w, t0, tj) }
    else {
      PK_PASS(ghi, t0, t1, t2, t3, tj)
      if (two) {
        if (low1) { PK_PASS1(low_top, low_cnt, l0, lj) }
        else { PK_PASS(ghi - MFSC_LANES, l0, l1, l2, l3, lj) }
      }
    }
#endif
    // A new best needs a cell at or above the best so far.  Most anti-diagonals have none (the best path advances two
    // anti-diagonals per matching pair, and noisy reads spend long stretches below their best): one vote decides, and the
    // maximum reduction, the finish-cell record and the updates run only when something can change.
    if (forced || wballot((lkey >> PK_SB) >= highF) != 0u) {
      const int dkey = wmax_i32(lkey);
      const int dj = nlo + (dkey & 0xff);
      highF = dkey >> PK_SB; FinD = d;
      dEnd = d + breaklen < N + M ? d + breaklen : N + M;
      const int dW = cB[dj & DP_SMASK];
      if (lane == (MFSC_LANES == 1 ? 0 : 4)) st4(rec, dj, dW, accU, accM);
    }
    // lanes 0..3 write the halo: one predicated store each, no branch.
    // halo: the next diagonal reads nD[nlo-1 .. nhi] and nI[nlo .. nhi+1]; diagonal d+2 reads best[nlo-1 .. nhi+1]
    {
#if MFSC_LANES == 1
      nD[(nlo - 1) & DP_SMASK] = PK_SENT; nI[(nhi + 1) & DP_SMASK] = PK_SENT; cB[(nlo - 1) & DP_SMASK] = PK_SENT; cB[(nhi + 1) & DP_SMASK] = PK_SENT;
#else
      int* hp = halo_base + (d & 1) * halo_par + ((nlo - 1 + halo_hi * (w + 1)) & DP_SMASK);
      if (lane < 4) *hp = PK_SENT;
#endif
    }
    // trim cells that fell more than max_diff below the best (narrows the next diagonal only)
    int lo = 0x7fffffff, hi = -0x7fffffff;
    const int thr = (highF - max_diff) * PK_K;      // a clean word is >= thr exactly when its score is
#ifdef MFSC_EMU
    wsync();
    for (int g = glo + lane; g <= ghi; g += MFSC_LANES) {
      const int4 v = ld4(cB + ((g * 4) & DP_SMASK));
      const int j0 = g * 4;
      unsigned m = (v.x >= thr ? 1u : 0u) | (v.y >= thr ? 2u : 0u) | (v.z >= thr ? 4u : 0u) | (v.w >= thr ? 8u : 0u);
      for (int c = 0; c < 4; c++) if (j0 + c < nlo || j0 + c > nhi) m &= ~(1u << c);
      if (m) {
        const int a = j0 + dev_ffs32(m) - 1, z = j0 + 31 - dev_clz32(m);
        if (a < lo) lo = a;
        if (z > hi) hi = z;
      }
    }
#else
    {
      // first and last cell of the lane at or above the threshold, as select chains on the four comparisons
      const bool p0 = t0 >= thr;
      if (w <= MFSC_LANES) {          // one cell per lane
        lo = p0 ? tj : 0x40000000;
        hi = p0 ? tj : -0x40000000;
      } else {
        const bool p1 = t1 >= thr, p2 = t2 >= thr, p3 = t3 >= thr;
        lo = tj + (p0 ? 0 : (p1 ? 1 : (p2 ? 2 : (p3 ? 3 : 0x40000000))));
        hi = tj + (p3 ? 3 : (p2 ? 2 : (p1 ? 1 : (p0 ? 0 : -0x40000000))));
      }
      if (two) {   // the low pass holds smaller j than the top pass
        const bool q0 = l0 >= thr;
        int lo_l, hi_l;
        if (low1) {
          lo_l = q0 ? lj : 0x40000000; hi_l = q0 ? lj : -0x40000000;
        } else {
          const bool q1 = l1 >= thr, q2 = l2 >= thr, q3 = l3 >= thr;
          lo_l = lj + (q0 ? 0 : (q1 ? 1 : (q2 ? 2 : (q3 ? 3 : 0x40000000))));
          hi_l = lj + (q3 ? 3 : (q2 ? 2 : (q1 ? 1 : (q0 ? 0 : -0x40000000))));
        }
        lo = lo_l < lo ? lo_l : lo;
        hi = hi > hi_l ? hi : hi_l;
      }
    }
#endif
    lo = wmin_i32(lo); hi = wmax_i32(hi);
    PK_ORDER();
    const bool none = lo > hi;
    if (!none && hi - lo + 2 > band_cap) {
      int l2 = lo, h2 = hi;
      while (l2 <= h2 && h2 - l2 + 2 > band_cap) { if ((cB[l2 & DP_SMASK] & PK_SCORE) < (cB[h2 & DP_SMASK] & PK_SCORE)) l2++; else h2--; }
      lo = wmax_i32(l2); hi = wmax_i32(h2);
      if (lane == 0) E.acc_caps++;
    }
    if ((d & 7) == 0) {
    if ((d & (PK_PERIOD - 1)) == 0) {
      // Rebase the live words: what this diagonal wrote (nD, nI, its best buffer over [nlo, nhi]) and the best buffer
      // of the previous diagonal over [nlo-1, nhi], the slots the next diagonal reads: each of those is either a cell
      // of the previous band or one of its two halo sentinels, told apart by value.  The other slots in reach are
      // halo and are rewritten as sentinels; slots further out hold leftovers nobody reads before they are written again.
      int* cP = bBase + ((d + 1) & 1) * DP_SLOTS;
      int mnU = 0x7fffffff, mxU = 0, mnM = 0x7fffffff, mxM = 0, mnS = 0x7fffffff;
      for (int j = nlo - 1 + lane; j <= nhi + 1; j += MFSC_LANES) {
        const int y = j & DP_SMASK;
        const bool inD = j >= nlo && j <= nhi;
        const int a[4] = {nD[y], nI[y], cB[y], cP[y]};
        const bool inP = j <= nhi && a[3] >= PK_LIVE;
#pragma unroll
        for (int t = 0; t < 4; t++)
          if (t < 3 ? inD : inP) {
            const int u_ = PK_UF(a[t]), m_ = PK_MF(a[t]);
            mnU = u_ < mnU ? u_ : mnU; mxU = u_ > mxU ? u_ : mxU; mnM = m_ < mnM ? m_ : mnM; mxM = m_ > mxM ? m_ : mxM;
          }
        // the handed-over gap states lie at most a gap penalty below the best states checked here
        if (inD) { const int s = a[2] >> PK_SB; mnS = s < mnS ? s : mnS; }
      }
      mnU = wmin_i32(mnU); mxU = wmax_i32(mxU); mnM = wmin_i32(mnM); mxM = wmax_i32(mxM); mnS = wmin_i32(mnS);
      if (mxU - mnU > u_span || mxM - mnM > m_span || mnS < PK_FLOOR) { E.pk_fail = 1; return o; }
      const int delta = (highF - PK_H0) * PK_K + mnU * PK_U1 + mnM;
      accU += mnU; accM += mnM; highF = PK_H0;
      wsync();
      for (int j = nlo - 1 + lane; j <= nhi + 1; j += MFSC_LANES) {
        const int y = j & DP_SMASK;
        const bool inD = j >= nlo && j <= nhi;
        const int vP = cP[y];
        nD[y] = inD ? nD[y] - delta : PK_SENT;
        nI[y] = inD ? nI[y] - delta : PK_SENT;
        cB[y] = inD ? cB[y] - delta : PK_SENT;
        cP[y] = (j <= nhi && vP >= PK_LIVE) ? vP - delta : PK_SENT;
      }
      wsync();
    }
    // Dead tail (same rule as the oracle's engine): when the band touches the end of a sequence and no cell of this
    // anti-diagonal can reach the best score again (a cell gains at most DP_GOOD per remaining base pair), and the far corner is
    // out of reach within the break length, nothing that is reported can change any more.
    if (!forced && FinD + breaklen < N + M && (nhi == M || d - nlo == N)) {
      int ub = -0x40000000;
      for (int j = nlo + lane; j <= nhi; j += MFSC_LANES) {
        const int rem_a = N - (d - j), rem_b = M - j;
        const int v = (cB[j & DP_SMASK] >> PK_SB) + DP_GOOD * (rem_a < rem_b ? rem_a : rem_b);
        ub = v > ub ? v : ub;
      }
      if (wmax_i32(ub) < highF) { d++; break; }
    }
    }
    {
      const int a = d + 1 - N, t = d + 1 < M ? d + 1 : M;
      nlo = none ? 1 : (lo > a ? lo : a);
      nhi = none ? 0 : (hi + 1 < t ? hi + 1 : t);
    }
  }
  const int dlast = d - 1;
  wsync();
  int FinJ = rec[0], finW = rec[1], finU = PK_UF(finW) + rec[2], finM = PK_MF(finW) + rec[3];
  if (dlast == N + M) {
    if (!optimal) {
      // the far corner counts wherever the best score is: its cell is the only one of the last diagonal
      const int lastW = (bBase + (dlast & 1) * DP_SLOTS)[M & DP_SMASK];
      o.reached = 1; FinD = N + M; FinJ = M; finU = PK_UF(lastW) + accU; finM = PK_MF(lastW) + accM;
    } else if (FinD == dlast) o.reached = 1;
  }
  o.fi = FinD - FinJ; o.fj = FinJ;
  o.cols = o.fi + finU; o.errs = finM + o.fi - o.fj + 2 * finU;
  if (lane == 0) {
    E.acc_cells += (unsigned long long)cells; E.acc_calls += 1ull;
    if (maxw > E.acc_maxw) E.acc_maxw = maxw;
  }
  if (TB && want_ops) o.n_ops = dp_traceback(E, FinD, FinJ);
  return o;
}
