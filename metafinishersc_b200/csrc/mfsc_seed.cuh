// mfsc_seed.cuh -- stage 2: maximal exact matches (what `mummer -maxmatch -b -c -l L -n` streams to
// mgaps inside nucmer; reference call sites alignerRobot.py:46-64, adaptorFix.py:21).
//
// The reference text is indexed at every position with k-mers (k <= L).  A query strand is probed
// only at positions p = 0, s, 2s, ... with s = L - k + 1: every match of length >= L contains a
// probe-aligned k-mer.  A hit (r, p) is extended to the left by at most s bases; if all s match,
// the probe at p - s lies inside the same match and reports it, so this hit is dropped (that is
// the left-maximality de-duplication -- no hash set, no atomics).  Otherwise the match is
// extended to the right 32 bases per step on the packed text and emitted when it is >= L.
#pragma once
#include "mfsc_seq.cuh"

struct DevParams {
  int minmatch, mincluster, maxgap, breaklen, diagdiff, maxmatch, simplify, band_cap;
  double diagfactor;
  int k, stride;
  int one;   // the integer 1, opaque to the compiler (see FADD in mfsc_extend.cuh)
  int pk_u_span, pk_m_span;   // packed DP engine: widest field spreads a rebase accepts (mfsc_extend_pk.cuh)
  int pk_clean, pk_prio, pk_score;   // PK_CLEAN, PK_P, PK_SCORE, opaque to the compiler like `one`
};

struct MemOut {
  int* qs; unsigned* s1; int* s2; int* len; int* rec;
  unsigned long long* counter;   // matches found (64-bit: it keeps counting past the capacity, the host then retries with a larger buffer)
  unsigned long long capacity;
};

MFSC_DEV unsigned long long mem_slot(unsigned long long* ctr) {
#if defined(MFSC_EMU)
  return atomic_add_u64(ctr, 1ull);
#else
  unsigned m = __activemask();
  int leader = __ffs((int)m) - 1;
  unsigned long long base = 0;
  if (lane_id() == leader) base = atomicAdd(ctr, (unsigned long long)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + (unsigned long long)__popc(m & ((1u << lane_id()) - 1u));
#endif
}

// ------------------------------------------------------------------------------------------
// Stage 2 kernel.  A warp owns one tile of a query strand (SEED_TILE_WORDS packed words = 2048 bases):
//   a. the tile's packed text and invalid mask are staged in shared memory with 128-bit / 64-bit coalesced loads
//      (one load per lane), and every probe window is then cut out of shared memory;
//   b. all lanes probe together, SEED_U probes per lane and round: k-mer -> directory entry of its bucket word (occupancy
//      bits + rank, 8 bytes per 32 buckets, resident in L2) -> only occupied buckets go on to the bucket heads,
//      SEED_U independent loads in flight per lane;
//   c. the probes that found candidates are compacted into a shared-memory list (ballot + popc) and dealt to the
//      lanes one probe each, so the lanes verifying candidates are the ones that have any; buckets with more than
//      SEED_HEAVY positions (repeats) are walked by the whole warp, 32 positions at a time.  Verification is two
//      phases, each on dense lanes: the left test (most candidates end there), then the survivors, compacted
//      again, are extended to the right;
//   d. a match that is still running after SEED_LONG bases is finished by the whole warp, 32 words (1024 bases)
//      per step -- whole-contig matches of a self-alignment (reference call shapes #2/#3/#5/#7 of SURVEY appendix A).
// Matches are appended with a warp-aggregated atomic; their order is fixed by the sort that follows.
// ------------------------------------------------------------------------------------------
#define SEED_TILE_WORDS 64
#define SEED_TILE_BASES (SEED_TILE_WORDS * 32)
#define SEED_STAGE (SEED_TILE_WORDS + 2)      // words staged: a window may start in the tile's last word
#define SEED_U 4
#define SEED_HEAVY 8
#define SEED_LONG 256
#define SEED_WARPS 4                          // warps per block
#ifndef SEED_ROUNDS_HELD
#define SEED_ROUNDS_HELD 1        // holding the list over two rounds was measured slower (cfg 1: 4.8 vs 2.7 ms): the larger shared-memory footprint costs L1
#endif
#ifndef SEED_BLOCKS_PER_SM
#define SEED_BLOCKS_PER_SM 12
#endif
#define SEED_LIST (SEED_ROUNDS_HELD * MFSC_LANES * SEED_U)    // the probes with candidates of up to SEED_ROUNDS_HELD rounds wait here, so that the lanes that verify are full

struct SeedTiles {
  const unsigned* tile_off;   // [n_qs + 1] first tile of every query strand (a strand of len bases has ceil(len / 2048) tiles)
  unsigned n_tiles;
  unsigned* ticket;
};

struct SeedShared {
  unsigned long long txt[SEED_STAGE];
  unsigned inv[SEED_STAGE];
  unsigned l_b0[SEED_LIST]; unsigned l_b1[SEED_LIST]; int l_p[SEED_LIST];     // probes with candidates: light buckets from the bottom, heavy ones from the top
  unsigned g_r[MFSC_LANES * 2]; int g_p[MFSC_LANES * 2]; int g_left[MFSC_LANES * 2]; int g_right[MFSC_LANES * 2];   // long matches
};

MFSC_DEV unsigned long long sm_win64(const unsigned long long* txt, int pos) {
  const int w = pos >> 5; const unsigned sh = (unsigned)(pos & 31) * 2u;
  const unsigned long long lo = txt[w];
  if (sh == 0) return lo;
  return (lo >> sh) | (txt[w + 1] << (64u - sh));
}
MFSC_DEV unsigned sm_win32(const unsigned* inv, int pos) {
  const int w = pos >> 5; const unsigned sh = (unsigned)(pos & 31);
  const unsigned lo = inv[w];
  if (sh == 0) return lo;
  return (lo >> sh) | (inv[w + 1] << (32u - sh));
}

struct SeedCtx {
  SeqStore R, Q; const unsigned* r_ostart; KmerIndex X; DevParams P; MemOut out;
  unsigned qstart; int qlen, qs;
};

// left test of one candidate (reference position r of a k-mer equal to the probe at p): the number of matching bases to the
// left, at most s; s means "an earlier probe owns this match"
MFSC_DEV int seed_left(const SeedCtx& C, unsigned r, int p) {
  const int s = C.P.stride;
  return match_left(C.R, r, C.Q, C.qstart + (unsigned)p, p < s ? p : s);
}

// right extension of a candidate that passed the left test, the first SEED_LONG bases: returns 1 when the match is still
// running (the warp finishes it), 2 when it ended
MFSC_DEV int seed_right(const SeedCtx& C, unsigned r, int p, int* right_o) {
  const int k = C.P.k;
  const int lim = C.qlen - p - k;
  const int first = lim < SEED_LONG ? lim : SEED_LONG;
  const int right = match_right(C.R, r + k, C.Q, C.qstart + (unsigned)(p + k), first);
  *right_o = right;
  return (right >= first && first < lim) ? 1 : 2;
}

MFSC_DEV void seed_emit(const SeedCtx& C, unsigned r, int p, int left, int right) {
  const int k = C.P.k;
  const int total = left + k + right;
  if (total < C.P.minmatch) return;
  const unsigned s1p = r - (unsigned)left;
  const int s2 = p - left;
  if (!C.P.maxmatch) {
    // -mumreference: the matched string must occur exactly once in the reference
    const unsigned long long kmask = (1ull << (2 * k)) - 1ull;
    const unsigned long long km0 = win64(C.Q.txt, C.qstart + (unsigned)s2) & kmask;
    bool unique = true;
    unsigned g0 = 0, g1 = 0;
    kmer_bucket(C.X, km0, g0, g1);
    for (unsigned g = g0; g < g1 && unique; g++) {
      const unsigned r2 = C.X.pos[g];
      if (r2 == s1p) continue;
      if (match_right(C.R, r2 + k, C.Q, C.qstart + (unsigned)s2 + k, total - k) >= total - k) unique = false;
    }
    if (!unique) return;
  }
  const unsigned long long slot = mem_slot(C.out.counter);
  if (slot < C.out.capacity) {
    const int rec = seq_of(C.R.start, C.R.n_seq, s1p);
    st_stream(&C.out.qs[slot], C.qs);
    st_stream(&C.out.s1[slot], s1p - C.R.start[rec] + C.r_ostart[rec]);
    st_stream(&C.out.s2[slot], s2);
    st_stream(&C.out.len[slot], total);
    st_stream(&C.out.rec[slot], rec);
  }
}

// the whole warp finishes one long match: words of 32 bases, one per lane and step
MFSC_DEV int seed_finish_long(const SeedCtx& C, unsigned r, int p, int right, int lane) {
  const int k = C.P.k;
  const unsigned rp = r + (unsigned)k, qp = C.qstart + (unsigned)(p + k);
  const int lim = C.qlen - p - k;
  int n = right;
  for (;;) {
    const int off = n + 32 * lane;
    int m = 32;
    if (off < lim) {
      const unsigned long long x = win64(C.R.txt, rp + off) ^ win64(C.Q.txt, qp + off);
      const unsigned long long mm = (x | (x >> 1)) & MFSC_EVEN;
      const unsigned ia = win32(C.R.inv, rp + off) | win32(C.Q.inv, qp + off);
      const int t = mm ? ((dev_ffs64(mm) - 1) >> 1) : 32;
      const int u = ia ? (dev_ffs32(ia) - 1) : 32;
      m = t < u ? t : u;
      if (off + m > lim) m = lim - off;
    } else m = 0;
    const unsigned stop = wballot(m < 32);
    if (stop) {
      const int first = dev_ffs32(stop) - 1;
      return n + 32 * first + wbcast_i32(m, first);
    }
    n += 32 * MFSC_LANES;
  }
}

// R: reference store; r_ostart[rec] = position of record rec's first base in the separator-joined
// coordinate system the clustering stage works in (off[rec] + rec).  X: the k-mer index (mfsc_seq.cuh).
MFSC_KERNEL k_seeds(SeqStore R, const unsigned* r_ostart, KmerIndex X, SeqStore Q, SeedTiles T, DevParams P, MemOut out) {
#ifdef MFSC_EMU
  static thread_local SeedShared sm_all[1];
  SeedShared& S = sm_all[0];
#else
  __shared__ SeedShared sm_all[SEED_WARPS];
  SeedShared& S = sm_all[warp_in_block()];
#endif
  const int lane = lane_id();
  const int k = P.k, s = P.stride;
  const unsigned long long kmask = (1ull << (2 * k)) - 1ull;
  const unsigned vmask = (1u << k) - 1u;
  const unsigned lt_mask = MFSC_LANES == 32 ? ((1u << lane) - 1u) : 0u;
  SeedCtx C; C.R = R; C.Q = Q; C.r_ostart = r_ostart; C.X = X; C.P = P; C.out = out;
  const unsigned* pos = X.pos;
  const unsigned TICKET = 4;                     // consecutive tiles per ticket: the strand is looked up once per ticket
  for (;;) {
    unsigned t0 = 0;
    if (lane == 0) t0 = atomic_add_u32(T.ticket, TICKET);
    t0 = (unsigned)wbcast_i32((int)t0, 0);
    if (t0 >= T.n_tiles) break;
    const unsigned t1 = t0 + TICKET < T.n_tiles ? t0 + TICKET : T.n_tiles;
    int qs = seq_of(T.tile_off, Q.n_seq, t0);
    for (unsigned t = t0; t < t1; t++) {
      while (T.tile_off[qs + 1] <= t) qs++;
      const int tile = (int)(t - T.tile_off[qs]);
      const unsigned qstart = Q.start[qs];
      const int qlen = Q.len[qs];
      C.qstart = qstart; C.qlen = qlen; C.qs = qs;
      const int base0 = tile * SEED_TILE_BASES;              // first base of the tile on the strand
      const int np = qlen >= k ? (qlen - k) / s + 1 : 0;     // probes of the strand: positions 0, s, 2s, ...
      const int pb = (base0 + s - 1) / s;
      int pe = (base0 + SEED_TILE_BASES + s - 1) / s;
      if (pe > np) pe = np;
      if (pb >= pe) continue;
      // a. stage the tile: lane l takes words 2l, 2l+1 (one 128-bit load of text, one 64-bit load of the mask);
      //    two more words for the windows that start in the last word.  Words past the strand's end are padding
      //    inside the buffer (layout in mfsc_seq.cuh) or clamped to its last word.
      wsync();
      {
        const unsigned w0 = (qstart >> 5) + (unsigned)tile * SEED_TILE_WORDS;
        const unsigned wmax = Q.n_words - 1u;
        for (int x = 2 * lane; x < SEED_STAGE; x += 2 * MFSC_LANES) {
          const unsigned wa = w0 + (unsigned)x < wmax ? w0 + (unsigned)x : wmax, wb = w0 + (unsigned)x + 1u < wmax ? w0 + (unsigned)x + 1u : wmax;
#if !defined(MFSC_EMU)
          if (wb == wa + 1u && (wa & 1u) == 0u) {
            // the reads stream through once: evict-first, so that they do not push the index (directory, heads, positions,
            // reference text -- re-read by every probe) out of L2
            const ulonglong2 v = __ldcs(reinterpret_cast<const ulonglong2*>(Q.txt + wa));
            const uint2 iv = __ldcs(reinterpret_cast<const uint2*>(Q.inv + wa));
            S.txt[x] = v.x; S.txt[x + 1] = v.y; S.inv[x] = iv.x; S.inv[x + 1] = iv.y;
          } else
#endif
          { S.txt[x] = Q.txt[wa]; S.txt[x + 1] = Q.txt[wb]; S.inv[x] = Q.inv[wa]; S.inv[x + 1] = Q.inv[wb]; }
        }
      }
      wsync();
      int n_light = 0, n_heavy = 0;                // probes waiting in the shared list (all of this tile)
      for (int i0 = pb; i0 < pe; i0 += MFSC_LANES * SEED_U) {
        // b. SEED_U probes per lane: k-mer, occupancy bit, bucket bounds of the occupied ones
        unsigned long long km[SEED_U]; bool live[SEED_U]; unsigned b0[SEED_U], b1[SEED_U];
#pragma unroll
        for (int u = 0; u < SEED_U; u++) {
          const int i = i0 + u * MFSC_LANES + lane;
          const int lp = i * s - base0;                       // probe position inside the staged tile
          live[u] = i < pe;
          km[u] = 0;
          if (live[u]) {
            live[u] = (sm_win32(S.inv, lp) & vmask) == 0u;
            km[u] = sm_win64(S.txt, lp) & kmask;
          }
        }
        // directory entries first (SEED_U independent loads in flight), then the heads of the occupied buckets
        uint2 dw[SEED_U];
#pragma unroll
        for (int u = 0; u < SEED_U; u++) { dw[u].x = 0; dw[u].y = 0; if (live[u]) dw[u] = ldg(&X.dir[km[u] >> 5]); }
#pragma unroll
        for (int u = 0; u < SEED_U; u++) {
          const unsigned bit = (unsigned)(km[u] & 31ull);
          live[u] = ((dw[u].x >> bit) & 1u) != 0u;
          b0[u] = 0; b1[u] = 0;
          if (live[u]) {
            const unsigned r = dw[u].y + (unsigned)dev_popc32(dw[u].x & ((1u << bit) - 1u));
            b0[u] = ldg(&X.heads[r]); b1[u] = ldg(&X.heads[r + 1]);
          }
        }
        // c. probes with candidates -> shared list (light buckets from the bottom, heavy ones from the top); the list is
        //    worked off when another round might not fit, and at the end of the tile
#pragma unroll
        for (int u = 0; u < SEED_U; u++) {
          const unsigned cnt = b1[u] - b0[u];
          const bool lt = cnt > 0u && cnt <= SEED_HEAVY, hv = cnt > SEED_HEAVY;
          const unsigned ml = wballot(lt), mh = wballot(hv);
          const int p = (i0 + u * MFSC_LANES + lane) * s;
          if (lt) { const int w = n_light + dev_popc32(ml & lt_mask); S.l_b0[w] = b0[u]; S.l_b1[w] = b1[u]; S.l_p[w] = p; }
          if (hv) { const int w = SEED_LIST - 1 - n_heavy - dev_popc32(mh & lt_mask); S.l_b0[w] = b0[u]; S.l_b1[w] = b1[u]; S.l_p[w] = p; }
          n_light += dev_popc32(ml); n_heavy += dev_popc32(mh);
        }
        if (n_light + n_heavy == 0) continue;
        if (i0 + MFSC_LANES * SEED_U < pe && n_light + n_heavy <= SEED_LIST - MFSC_LANES * SEED_U) continue;
        wsync();
        // Verification in two phases, so that each runs on dense lanes.  Phase 1, one candidate per lane: position, left test.
        // Most candidates end here (inside a match an earlier probe owns, or a chance k-mer hit).  The survivors -- the left
        // ends of matches -- are compacted into a second list, and phase 2 extends them to the right, one per lane.
        int n_v = 0;                                 // survivors waiting in S.g_*
        // phase 2 + 3 over the survivor list
#define SEED_EXTEND_SURVIVORS()                                                                                       \
        {                                                                                                              \
          wsync();                                                                                                     \
          for (int x0 = 0; x0 < n_v; x0 += MFSC_LANES) {                                                               \
            const int x = x0 + lane;                                                                                   \
            int kind = 0, right = 0; unsigned r = 0; int p = 0, left = 0;                                              \
            if (x < n_v) { r = S.g_r[x]; p = S.g_p[x]; left = S.g_left[x]; kind = seed_right(C, r, p, &right); }       \
            if (kind == 2) seed_emit(C, r, p, left, right);                                                            \
            if (x < n_v) S.g_right[x] = kind == 1 ? right : -1;                                                        \
          }                                                                                                            \
          wsync();                                                                                                     \
          for (int g = 0; g < n_v; g++) {                  /* matches still running: the whole warp finishes each */   \
            const int gright = S.g_right[g];                                                                           \
            if (gright < 0) continue;                                                                                  \
            const unsigned gr = S.g_r[g]; const int gp = S.g_p[g], gl = S.g_left[g];                                   \
            const int tot = seed_finish_long(C, gr, gp, gright, lane);                                                 \
            if (lane == 0) seed_emit(C, gr, gp, gl, tot);                                                              \
          }                                                                                                            \
          n_v = 0;                                                                                                     \
          wsync();                                                                                                     \
        }
        // light buckets: one probe per lane
        for (int x0 = 0; x0 < n_light; x0 += MFSC_LANES) {
          const int x = x0 + lane;
          unsigned h = 0, he = 0; int p = 0;
          if (x < n_light) { h = S.l_b0[x]; he = S.l_b1[x]; p = S.l_p[x]; }
          for (int step = 0; step < SEED_HEAVY; step++) {
            int left = s; unsigned r = 0;
            const bool have = h < he;
            if (have) r = pos[h];
            bool owned = false;
#if MFSC_LANES == 32
            if (step == 0 && k >= s) {
              // The list holds the probes in order, so the lane below usually holds the previous probe (p - s).  When that probe's
              // first candidate is r - s, the s bases left of (r, p) lie inside its exact k-mer match (k >= s): this candidate
              // extends at least s bases to the left and belongs to an earlier probe -- dropped without touching the reference.
              const unsigned r_prev = __shfl_up_sync(MFSC_FULL, r, 1);
              const int p_prev = __shfl_up_sync(MFSC_FULL, p, 1);
              const bool have_prev = __shfl_up_sync(MFSC_FULL, have ? 1 : 0, 1) != 0;
              owned = have && lane > 0 && have_prev && p_prev == p - s && r_prev == r - (unsigned)s;
            }
#endif
            if (have) { if (!owned) left = seed_left(C, r, p); h++; }
            const bool surv = have && left < s;
            const unsigned mv = wballot(surv);
            if (mv) {
              if (surv) { const int w = n_v + dev_popc32(mv & lt_mask); S.g_r[w] = r; S.g_p[w] = p; S.g_left[w] = left; }
              n_v += dev_popc32(mv);
              if (n_v > MFSC_LANES) SEED_EXTEND_SURVIVORS()       // keep room for another 32
            }
            if (wballot(h < he) == 0u) break;
          }
        }
        // heavy buckets: the warp walks one bucket at a time
        for (int y = 0; y < n_heavy; y++) {
          const int e = SEED_LIST - 1 - y;
          const unsigned hb = S.l_b0[e], he = S.l_b1[e]; const int p = S.l_p[e];
          for (unsigned h0 = hb; h0 < he; h0 += MFSC_LANES) {
            const unsigned h = h0 + (unsigned)lane;
            int left = s; unsigned r = 0;
            if (h < he) { r = pos[h]; left = seed_left(C, r, p); }
            const bool surv = h < he && left < s;
            const unsigned mv = wballot(surv);
            if (mv) {
              if (surv) { const int w = n_v + dev_popc32(mv & lt_mask); S.g_r[w] = r; S.g_p[w] = p; S.g_left[w] = left; }
              n_v += dev_popc32(mv);
              if (n_v > MFSC_LANES) SEED_EXTEND_SURVIVORS()
            }
          }
        }
        if (n_v > 0) SEED_EXTEND_SURVIVORS()
#undef SEED_EXTEND_SURVIVORS
        n_light = 0; n_heavy = 0;
        wsync();
      }
    }
  }
}

// seg[q] = first index i with qs[i] >= q, for q = 0..n_qs  (qs ascending)
MFSC_KERNEL k_seg_bounds(const int* qs, long long n, int n_qs, int* seg) {
  for (long long i = thread_global(); i <= n; i += threads_total()) {
    int prev = (i == 0) ? -1 : qs[i - 1];
    int cur = (i == n) ? n_qs : qs[i];
    for (int q = prev + 1; q <= cur; q++) seg[q] = (int)i;
  }
}

// permutation gather used after the two stable radix passes that order matches by (qs, s2, s1)
MFSC_KERNEL k_gather_mems(const unsigned* perm, long long n, const int* qs, const unsigned* s1, const int* s2,
                          const int* len, const int* rec, int* oqs, unsigned* os1, int* os2, int* olen, int* orec) {
  for (long long i = thread_global(); i < n; i += threads_total()) {
    unsigned j = perm[i];
    oqs[i] = qs[j]; os1[i] = s1[j]; os2[i] = s2[j]; olen[i] = len[j]; orec[i] = rec[j];
  }
}

MFSC_KERNEL k_iota(unsigned* a, long long n) {
  for (long long i = thread_global(); i < n; i += threads_total()) a[i] = (unsigned)i;
}

MFSC_KERNEL k_sort_key(const int* qs, const int* s2, const unsigned* perm, long long n, unsigned long long* key) {
  for (long long i = thread_global(); i < n; i += threads_total()) {
    unsigned j = perm[i];
    key[i] = ((unsigned long long)(unsigned)qs[j] << 32) | (unsigned)s2[j];
  }
}

MFSC_KERNEL k_sort_key1(const unsigned* s1, long long n, unsigned* key) {
  for (long long i = thread_global(); i < n; i += threads_total()) key[i] = s1[i];
}
