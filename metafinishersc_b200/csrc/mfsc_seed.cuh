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

struct ProbeChunk { int qs, pb, pe; };

struct MemOut {
  int* qs; unsigned* s1; int* s2; int* len; int* rec;
  unsigned* counter;       // [0] = matches found (may exceed capacity), [1] = probes with >=1 candidate, [2..3] = candidates (u64)
  unsigned capacity;
};

MFSC_DEV unsigned mem_slot(unsigned* ctr) {
#if defined(MFSC_EMU)
  return atomic_add_u32(ctr, 1u);
#else
  unsigned m = __activemask();
  int leader = __ffs((int)m) - 1;
  unsigned base = 0;
  if (lane_id() == leader) base = atomicAdd(ctr, (unsigned)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + (unsigned)__popc(m & ((1u << lane_id()) - 1u));
#endif
}

// R: reference store; r_ostart[rec] = position of record rec's first base in the separator-joined
// coordinate system the clustering stage works in (off[rec] + rec).
MFSC_KERNEL k_seeds(SeqStore R, const unsigned* r_ostart, const unsigned* H, const unsigned* pos, SeqStore Q,
                    const ProbeChunk* chunks, int n_chunks, DevParams P, MemOut out) {
  const int k = P.k, s = P.stride, L = P.minmatch;
  const unsigned long long kmask = (1ull << (2 * k)) - 1ull;
  const unsigned vmask = (1u << k) - 1u;
  for (int c = (int)blockIdx.x; c < n_chunks; c += (int)gridDim.x) {
    const ProbeChunk ch = chunks[c];
    const unsigned qstart = Q.start[ch.qs];
    const int qlen = Q.len[ch.qs];
    for (int i = ch.pb + (int)threadIdx.x; i < ch.pe; i += (int)blockDim.x) {
      const int p = i * s;
      const unsigned qp = qstart + (unsigned)p;
      if (win32(Q.inv, qp) & vmask) continue;
      const unsigned long long km = win64(Q.txt, qp) & kmask;
      // the bucket heads (1/4 GB and up) are touched once per probe at random: stream them through L2 so that the
      // position table and the packed reference text, which every candidate re-reads, stay resident
      const unsigned b0 = ld_stream(&H[km]), b1 = ld_stream(&H[km + 1]);
      for (unsigned h = b0; h < b1; h++) {
        const unsigned r = pos[h];
        const int maxl = p < s ? p : s;
        const int left = match_left(R, r, Q, qp, maxl);
        if (left >= s) continue;  // an earlier probe owns this match
        const int right = match_right(R, r + k, Q, qp + k, qlen - p - k);
        const int total = left + k + right;
        if (total < L) continue;
        const unsigned s1p = r - (unsigned)left;
        const int s2 = p - left;
        if (!P.maxmatch) {
          // -mumreference: the matched string must occur exactly once in the reference
          const unsigned long long km0 = win64(Q.txt, qstart + (unsigned)s2) & kmask;
          bool unique = true;
          for (unsigned g = H[km0]; g < H[km0 + 1] && unique; g++) {
            const unsigned r2 = pos[g];
            if (r2 == s1p) continue;
            if (match_right(R, r2 + k, Q, qstart + (unsigned)s2 + k, total - k) >= total - k) unique = false;
          }
          if (!unique) continue;
        }
        const unsigned slot = mem_slot(&out.counter[0]);
        if (slot < out.capacity) {
          const int rec = seq_of(R.start, R.n_seq, s1p);
          out.qs[slot] = ch.qs;
          out.s1[slot] = s1p - R.start[rec] + r_ostart[rec];
          out.s2[slot] = s2;
          out.len[slot] = total;
          out.rec[slot] = rec;
        }
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
