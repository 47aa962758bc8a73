// mfsc_seq.cuh -- 2-bit packed sequence store in HBM and the exact-match primitives on it.
//
// Layout (both the contig index text and the read batches use it):
//   txt : uint64 words, 32 bases per word, base b of a word in bits [2b, 2b+1]  (A=0 C=1 G=2 T=3)
//   inv : uint32 words, 1 bit per base, same word index as txt; 1 = "never matches"
//         (non-ACGT character, padding between records, padding at both ends)
//   Every sequence starts on a word boundary; word 0 and the tail of every record's last word
//   are invalid padding (a full extra word when the length is a multiple of 32), so a maximal
//   match can never run across a record end, exactly as `mummer -n` behaves on the
//   concatenated reference (reference call site: alignerRobot.py:46-64 `nucmer --maxmatch`).
//   Positions are uint32 (max reference named by BASELINE.json: 2e9 bases).
#pragma once
#include "mfsc_platform.cuh"

struct SeqStore {
  const unsigned long long* txt;
  const unsigned* inv;
  const unsigned* start;  // [n_seq]  padded position of each sequence's first base (multiple of 32)
  const int* len;         // [n_seq]
  int n_seq;
  unsigned n_words;
};

MFSC_DEV unsigned long long win64(const unsigned long long* txt, unsigned pos) {
  unsigned w = pos >> 5, sh = (pos & 31u) * 2u;
  unsigned long long lo = txt[w];
  if (sh == 0) return lo;
  unsigned long long hi = txt[w + 1];
  return (lo >> sh) | (hi << (64u - sh));
}

MFSC_DEV unsigned win32(const unsigned* inv, unsigned pos) {
  unsigned w = pos >> 5, sh = pos & 31u;
  unsigned lo = inv[w];
  if (sh == 0) return lo;
  unsigned hi = inv[w + 1];
  return (lo >> sh) | (hi << (32u - sh));
}

// base code 0..3, or 4 when invalid
MFSC_DEV int code_at(const SeqStore& S, unsigned pos) {
  unsigned long long w = S.txt[pos >> 5];
  int c = (int)((w >> ((pos & 31u) * 2u)) & 3ull);
  if ((S.inv[pos >> 5] >> (pos & 31u)) & 1u) c = 4;
  return c;
}

#define MFSC_EVEN 0x5555555555555555ull

// number of consecutive matching valid bases at (rp+t, qp+t), t = 0.., capped at lim
MFSC_DEV int match_right(const SeqStore& R, unsigned rp, const SeqStore& Q, unsigned qp, int lim) {
  int n = 0;
  while (n < lim) {
    unsigned long long x = win64(R.txt, rp + n) ^ win64(Q.txt, qp + n);
    unsigned long long mm = (x | (x >> 1)) & MFSC_EVEN;
    unsigned ia = win32(R.inv, rp + n) | win32(Q.inv, qp + n);
    int t = mm ? ((dev_ffs64(mm) - 1) >> 1) : 32;
    int u = ia ? (dev_ffs32(ia) - 1) : 32;
    int m = t < u ? t : u;
    n += m;
    if (m < 32) break;
  }
  return n < lim ? n : lim;
}

// number of consecutive matching valid bases at (rp-1-t, qp-1-t), t = 0.., capped at maxn
MFSC_DEV int match_left(const SeqStore& R, unsigned rp, const SeqStore& Q, unsigned qp, int maxn) {
  int n = 0;
  while (n < maxn) {
    unsigned long long x = win64(R.txt, rp - n - 32) ^ win64(Q.txt, qp - n - 32);
    unsigned long long mm = (x | (x >> 1)) & MFSC_EVEN;
    unsigned ia = win32(R.inv, rp - n - 32) | win32(Q.inv, qp - n - 32);
    int t = mm ? ((dev_clz64(mm) - 1) >> 1) : 32;
    int u = ia ? dev_clz32(ia) : 32;
    int m = t < u ? t : u;
    n += m;
    if (m < 32) break;
  }
  return n < maxn ? n : maxn;
}

// index of the sequence containing padded position pos (start[] ascending)
MFSC_DEV int seq_of(const unsigned* start, int n, unsigned pos) {
  int lo = 0, hi = n;  // first index with start > pos
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (start[mid] > pos) hi = mid; else lo = mid + 1;
  }
  return lo - 1;
}

// ------------------------------------------------------------------------------------------
// ASCII -> packed store.  One thread per output word.
//   in_off[n_rec+1] : offsets of the records in `bases`
//   wstart[n_seq+1] : first word of each stored sequence (n_seq = n_rec, or 2*n_rec when
//                     both strands are stored: sequence 2r = record r forward, 2r+1 = its
//                     reverse complement, which is what `mummer -b -c` streams per query)
// ------------------------------------------------------------------------------------------
MFSC_DEV int ascii_code(unsigned char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
  }
}

MFSC_KERNEL k_pack(const unsigned char* bases, const long long* in_off, const unsigned* wstart, int n_seq,
                   int both_strands, unsigned n_words, unsigned long long* txt, unsigned* inv) {
  for (long long w = thread_global(); w < (long long)n_words; w += threads_total()) {
    unsigned long long t = 0;
    unsigned iv = 0xffffffffu;
    if (n_seq > 0 && w >= (long long)wstart[0] && w < (long long)wstart[n_seq]) {
      int k = seq_of(wstart, n_seq, (unsigned)w);
      int r = both_strands ? (k >> 1) : k;
      int rc = both_strands ? (k & 1) : 0;
      long long o = in_off[r];
      int len = (int)(in_off[r + 1] - o);
      int i0 = (int)(w - wstart[k]) * 32;
      iv = 0;
      for (int b = 0; b < 32; b++) {
        int i = i0 + b, c = 4;
        if (i < len) {
          c = ascii_code(rc ? bases[o + (len - 1 - i)] : bases[o + i]);
          if (rc && c < 4) c = 3 - c;
        }
        if (c > 3) iv |= 1u << b; else t |= (unsigned long long)c << (2 * b);
      }
    }
    txt[w] = t;
    inv[w] = iv;
  }
}

// ------------------------------------------------------------------------------------------
// k-mer bucket index of the reference text (stage 1).
//
// What the seed kernel reads (KmerIndex): most of the 4^k buckets are empty (k is chosen so that 4^k >= 4 R), so the
// index keeps a DIRECTORY of one 8-byte entry per 32 buckets -- {occupancy bits, number of occupied buckets before this
// word} -- and a bucket-start offset (`heads`) for the OCCUPIED buckets only:
//     bucket b occupied  <=>  bit (b & 31) of dir[b >> 5].x
//     its positions      =    pos[heads[r] .. heads[r + 1]),  r = dir[b >> 5].y + popc(dir[b >> 5].x & ((1 << (b & 31)) - 1))
// For a 5-10 Mbp reference at k = 13 that is 16 MB of directory + 4 B per occupied bucket instead of 268 MB of dense
// bucket heads: directory, heads, positions and text together fit the 126 MB L2, and four times less has to cross
// NVLink when the index is replicated.
//
// Built from a dense histogram that is scratch: count -> inclusive scan -> fill -> directory bits -> scan of the
// per-word counts -> heads.  One thread per text word (32 start positions).  The kernels take a bucket range
// [b_lo, b_hi): a build can be restricted to a slice of the k-mer space (mfsc_index_build_sharded: every GPU of a
// box builds one slice, the slices are exchanged with NCCL all-gathers).
// ------------------------------------------------------------------------------------------
struct KmerIndex {
  const uint2* dir;         // [4^k / 32 (+1)]
  const unsigned* heads;    // [occupied buckets + 1]
  const unsigned* pos;      // [indexed positions]
};

// bucket of k-mer km: false when it is empty
MFSC_DEV bool kmer_bucket(const KmerIndex& X, unsigned long long km, unsigned& b0, unsigned& b1) {
  const uint2 w = X.dir[km >> 5];
  const unsigned bit = (unsigned)(km & 31ull);
  if (!((w.x >> bit) & 1u)) return false;
  const unsigned r = w.y + (unsigned)dev_popc32(w.x & ((1u << bit) - 1u));
  b0 = X.heads[r]; b1 = X.heads[r + 1];
  return true;
}

MFSC_KERNEL k_index_count(const unsigned long long* txt, const unsigned* inv, unsigned n_words, int k, unsigned long long b_lo,
                          unsigned long long b_hi, unsigned* H) {
  const unsigned long long kmask = (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1ull);
  const unsigned vmask = (k >= 32) ? ~0u : ((1u << k) - 1u);
  for (long long w = thread_global(); w + 1 < (long long)n_words; w += threads_total()) {
    unsigned i0 = inv[w];
    if (i0 == 0xffffffffu) continue;
    unsigned i1 = inv[w + 1];
    unsigned long long t0 = txt[w], t1 = txt[w + 1];
    for (int b = 0; b < 32; b++) {
      unsigned iw = b ? ((i0 >> b) | (i1 << (32 - b))) : i0;
      if (iw & vmask) continue;
      unsigned long long tw = b ? ((t0 >> (2 * b)) | (t1 << (64 - 2 * b))) : t0;
      const unsigned long long km = tw & kmask;
      if (km < b_lo || km >= b_hi) continue;
      atomic_add_u32(&H[(km - b_lo) + 2], 1u);
    }
  }
}

// after the inclusive scan H[x + 1] = first slot of bucket b_lo + x; the fill advances it to the bucket's end, so that
// afterwards H[x] = first slot of bucket b_lo + x (and H[n] = total)
MFSC_KERNEL k_index_fill(const unsigned long long* txt, const unsigned* inv, unsigned n_words, int k, unsigned long long b_lo,
                         unsigned long long b_hi, unsigned* H, unsigned* pos) {
  const unsigned long long kmask = (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1ull);
  const unsigned vmask = (k >= 32) ? ~0u : ((1u << k) - 1u);
  for (long long w = thread_global(); w + 1 < (long long)n_words; w += threads_total()) {
    unsigned i0 = inv[w];
    if (i0 == 0xffffffffu) continue;
    unsigned i1 = inv[w + 1];
    unsigned long long t0 = txt[w], t1 = txt[w + 1];
    for (int b = 0; b < 32; b++) {
      unsigned iw = b ? ((i0 >> b) | (i1 << (32 - b))) : i0;
      if (iw & vmask) continue;
      unsigned long long tw = b ? ((t0 >> (2 * b)) | (t1 << (64 - 2 * b))) : t0;
      const unsigned long long km = tw & kmask;
      if (km < b_lo || km >= b_hi) continue;
      unsigned slot = atomic_add_u32(&H[(km - b_lo) + 1], 1u);
      pos[slot] = (unsigned)(w * 32 + b);
    }
  }
}

// directory bits of the words of a bucket slice (n_buckets a multiple of 32), and the number of occupied buckets per word
MFSC_KERNEL k_index_dir_bits(const unsigned* H, long long n_buckets, unsigned* bits, unsigned* cnt) {
  const long long n_words = n_buckets / 32;
  for (long long w = thread_global(); w < n_words; w += threads_total()) {
    unsigned m = 0;
    unsigned prev = H[w * 32];
    for (int b = 0; b < 32; b++) {
      const unsigned nx = H[w * 32 + b + 1];
      if (nx > prev) m |= 1u << b;
      prev = nx;
    }
    bits[w] = m;
    cnt[w] = (unsigned)dev_popc32(m);
  }
}

// rank[w] = exclusive scan of cnt (done by the caller).  Writes dir[w] = {bits, rank_base + rank[w]} and the start offset
// (pos_base + H[b]) of every occupied bucket into heads[rank[w] ...].
MFSC_KERNEL k_index_heads(const unsigned* H, long long n_buckets, const unsigned* bits, const unsigned* rank, unsigned rank_base,
                          unsigned pos_base, uint2* dir, unsigned* heads) {
  const long long n_words = n_buckets / 32;
  for (long long w = thread_global(); w < n_words; w += threads_total()) {
    unsigned m = bits[w];
    unsigned r = rank[w];
    uint2 e; e.x = m; e.y = rank_base + r;
    dir[w] = e;
    while (m) {
      const int b = dev_ffs32(m) - 1;
      m &= m - 1u;
      heads[r++] = pos_base + H[w * 32 + b];
    }
  }
}

// adds a constant to a range of directory ranks / heads (slices of a sharded build become global)
MFSC_KERNEL k_add_u32(unsigned* a, long long n, unsigned v) {
  for (long long i = thread_global(); i < n; i += threads_total()) a[i] += v;
}

// rank[w] = incl[w] - popc(bits[w]): the exclusive scan from the inclusive one
MFSC_KERNEL k_excl_from_incl(const unsigned* incl, const unsigned* bits, long long n, unsigned* rank) {
  for (long long i = thread_global(); i < n; i += threads_total()) rank[i] = incl[i] - (unsigned)dev_popc32(bits[i]);
}

// slice of a sharded build into the global arrays: directory words with their ranks made global, heads with their offsets made global
MFSC_KERNEL k_dir_place(const uint2* src, long long n, unsigned rank_base, uint2* dst) {
  for (long long i = thread_global(); i < n; i += threads_total()) { uint2 e = src[i]; e.y += rank_base; dst[i] = e; }
}
MFSC_KERNEL k_copy_add_u32(const unsigned* src, long long n, unsigned v, unsigned* dst) {
  for (long long i = thread_global(); i < n; i += threads_total()) dst[i] = src[i] + v;
}
