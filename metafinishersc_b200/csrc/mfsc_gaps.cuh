// mfsc_gaps.cuh -- stage 4a: the forward extensions whose start does not depend on the walk, as independent tasks.
//
// The walk of stage 4 (k_extend, mfsc_extend.cuh) handles one read per warp.  Nearly all of its DP cells belong to
// forward extensions that start on the end of a cluster match and aim at a target fixed by the clusters alone
// (postnuc.cc extendClusters):
//   A  from the end of match k of a cluster to the start of match k + 1 (when one side is longer than GAP_MIN; the
//      short ones are the per-lane rectangles of the walk);
//   B  from the end of a cluster's last match to the cluster ext_forward_target picks, or to the end of the sequences.
// Such a DP is a function of the cluster set, provided the walk arrives with its alignment ending on that match -- the
// normal case.  So it does not have to wait for the walk: k_gap_plan lists these extensions for all reads (and leaves
// the walking order of every read in I.ord), k_gaps runs them one per warp from a ticket counter, and the walk picks a
// result up where it would have called the engine (gap_take), falling back to the in-line DP whenever there is no usable
// result (packed fields exceeded, the walk stands elsewhere, another target).  What stays in the walk: backward
// extensions, merges (forced re-alignment), short gaps.  A 10 kbp read at 15 % error is ~15-20 ms of one warp when the
// walk does everything; as ~10 tasks of ~1 ms plus a short walk, a batch of a few hundred reads fills the GPU, and a
// large batch does not end on a tail of half-idle SMs.  Results are those of the in-line DP by construction (same engine,
// same arguments); cells, calls and widest band are carried along so that the statistics are the same too.
#pragma once

#ifndef MFSC_GAP_MIN_MATCHES_PER_READ
#define MFSC_GAP_MIN_MATCHES_PER_READ 1000   // stage 4a runs when a batch has at least this many matches per read (mfsc_api.cu)
#endif

struct GapPlan {
  int4* task;                         // (start match m, strand index 2 q + dir, tA, tB; tB negated: `optimal` extension)
  int* task_rec;                      // reference record of the task
  unsigned* n_task;
  int4* res;                          // per start match, see gap_take (mfsc_extend.cuh)
};

MFSC_DEV void gap_emit(const GapPlan& G, int m, int qs, int rec, int tA, int tB, bool optimal) {
  G.res[m] = make_int4(0, 0, 0, 0);
  const unsigned slot = atomic_add_u32(G.n_task, 1u);
  G.task[slot] = make_int4(m, qs, tA, optimal ? -tB : tB);
  G.task_rec[slot] = rec;
}

// One warp per read: the walking order of its clusters (ext_order), then per cluster (one lane each) its kind A and
// kind B extensions.
MFSC_KERNEL k_gap_plan(SeqStore RS, SeqStore QS, ExtIn I, int n_reads, DevParams P, GapPlan G) {
  const int lane = lane_id();
  for (long long q = warp_global(); q < n_reads; q += warps_total()) {
    const int b0 = I.seg[2 * q], b1 = I.seg[2 * q + 1];
    const int n0 = I.n_pc[2 * q], n1 = I.n_pc[2 * q + 1];
    const int nC = n0 + n1;
    if (nC == 0) continue;
    ext_order(I, b0, b1, n0, n1, lane);
    const int* ord = I.ord + b0;
    ExtCtx E;                                      // what the cluster accessors and ext_forward_target read
    E.P = P; E.lane = lane; E.pc_rc0 = b1; E.pc_first = I.pc_first; E.pc_n = I.pc_n;
    E.cm_s1 = I.cm_s1; E.cm_s2 = I.cm_s2; E.cm_len = I.cm_len;
    E.B[0].n = QS.len[2 * q];
    int g0 = 0;
    while (g0 < nC) {
      const int rec = I.pc_rec[ord[g0]];
      int g1 = g0 + 1;
      while (g1 < nC && I.pc_rec[ord[g1]] == rec) g1++;
      E.ord = ord + g0; E.nC = g1 - g0; E.A.n = RS.len[rec]; E.a_ostart = I.r_ostart[rec];
      for (int cur = lane; cur < E.nC; cur += MFSC_LANES) {
        const int pc = E.ord[cur], first = I.pc_first[pc], nm = I.pc_n[pc], qs = 2 * (int)q + (pc >= b1 ? 1 : 0);
        for (int t = 0; t + 1 < nm; t++) {
          const int m = first + t, gl = I.cm_len[m];
          const int sA = (int)(I.cm_s1[m] - E.a_ostart) + gl, sB = I.cm_s2[m] + gl;
          const int tA = (int)(I.cm_s1[m + 1] - E.a_ostart) + 1, tB = I.cm_s2[m + 1] + 1;
          if (gap_is_task(tA - sA + 1, tB - sB + 1)) gap_emit(G, m, qs, rec, tA, tB, false);
        }
        int tA = E.A.n, tB = E.B[0].n;
        const int targetC = ext_forward_target(E, cur, tA, tB);
        gap_emit(G, first + nm - 1, qs, rec, tA, tB, targetC < 0);
      }
      g0 = g1;
    }
    wsync();
  }
}

template <bool PK>
MFSC_KERNEL_LB(128, PK ? MFSC_EXT_MINBLOCKS_PK : MFSC_EXT_MINBLOCKS) k_gaps(SeqStore RS, SeqStore QS, ExtIn I, DevParams P, GapPlan G, unsigned* ticket) {
#ifdef MFSC_EMU
  int* sm = emu_dp_smem;
#else
  extern __shared__ int4 dp_smem4[];
  int* sm = reinterpret_cast<int*>(dp_smem4) + warp_in_block() * (PK ? PK_SMEM_INTS : DP_SMEM_INTS);
#endif
  const int lane = lane_id();
  const unsigned n_task = *G.n_task;
  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomic_add_u32(ticket, 1u);
    t = wbcast_u32(t, 0);
    if (t >= n_task) break;
    const int4 T = G.task[t];
    const int m = T.x, q = T.y >> 1, dir = T.y & 1;
    const int rec = G.task_rec[t];
    ExtCtx E;
    E.RS = RS; E.QS = QS; E.P = P; E.sm = sm; E.lane = lane;
    E.acc_cells = 0; E.acc_calls = 0; E.acc_maxw = 0; E.acc_caps = 0; E.acc_gaps = 0;
    E.T.tb = nullptr; E.T.rev = nullptr; E.T.arena = nullptr; E.T.cap = 0; E.T.top = 0; E.T.bad = 0;
    E.pk_fail = 0;
    E.c_idx = -1; E.c_sA = E.c_sB = E.c_eA = E.c_eB = E.c_cols = E.c_errs = 0;
    E.A.base = RS.start[rec]; E.A.n = RS.len[rec];
    E.B[0].base = QS.start[2 * q]; E.B[0].n = QS.len[2 * q];
    E.B[1].base = QS.start[2 * q + 1]; E.B[1].n = QS.len[2 * q + 1];
    const unsigned ao = I.r_ostart[rec];
    const int gl = I.cm_len[m];
    const int sA = (int)(I.cm_s1[m] - ao) + gl, sB = I.cm_s2[m] + gl;                   // the last base of match m (1-based)
    int tA = T.z, tB = T.w < 0 ? -T.w : T.w;
    const int key = gap_target_key(tA, tB);
    // as ext_forward: an extension longer than the DP limit is cut there, counts as not reached and stops where its score peaks
    bool optimal = T.w < 0, overflow = false;
    if (tA - sA + 1 > DP_MAX_ALN) { tA = sA + DP_MAX_ALN - 1; overflow = true; optimal = true; }
    if (tB - sB + 1 > DP_MAX_ALN) { tB = sB + DP_MAX_ALN - 1; overflow = true; optimal = true; }
    const DpOut o = dp_run<false, PK>(E, dir, sA, sB, +1, tA - sA + 1, tB - sB + 1, false, optimal, true);
    wsync();
    if (lane == 0) {
      int4 r = make_int4(0, 0, 0, 0);
      if (!E.pk_fail && E.acc_caps == 0 && E.acc_calls == 1ull && E.acc_cells < (1ull << 23) && E.acc_maxw < 256 &&
          o.fi >= 0 && o.fi < 32768 && o.fj >= 0 && o.fj < 32768 && o.cols >= 0 && o.cols < 65536 && o.errs >= 0 && o.errs < 32768)
        r = make_int4(1 | ((o.reached && !overflow) ? 2 : 0) | (o.fi << 2) | (o.fj << 17), o.cols | (o.errs << 16),
                      (int)E.acc_cells | (E.acc_maxw << 23), key);
      G.res[m] = r;
    }
  }
}
