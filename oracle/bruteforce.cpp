// oracle/bruteforce.cpp
//
// TEST INFRASTRUCTURE ONLY -- a second opinion on the DP of oracle/mfsc_oracle.cpp (and, through it, of the CUDA
// kernels, which are bit-exact against that oracle).  Written independently of it: no shared code, full unbanded
// matrices filled row by row, nothing trimmed, no anti-diagonals.  It restates the textbook three-state affine-gap
// recurrence with the nucleotide scores the oracle attributes to MUMmer 3.23's postnuc (match +3, anything else -7,
// first gap column -7, every further one -4; `nucmer` call sites: alignerRobot.py:46-64 of the reference) and answers
//   bf_global      : best score of an alignment of all of a with all of b
//   bf_extend      : the cell an X-drop extension ends on when nothing is ever trimmed (best score over the
//                    anti-diagonals seen before `breaklen` of them pass without reaching the best again; ties go to
//                    the later anti-diagonal, then to the larger j)
//   bf_score_ops   : score, error columns and consumed lengths of a given column string
//   bf_edit_distance: unit-cost edit distance (lower bound on the error columns of ANY alignment of two strings)
// tests/test_bruteforce.py compares the oracle's engine with these on windows of up to 2 kbp.
// PARITY UNPINNED against real MUMmer, like the oracle: this file checks the oracle's arithmetic, not MUMmer's.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {
const int MATCH = 3, OTHER = -7, GAP_FIRST = -7, GAP_NEXT = -4;
const int NINF = -(1 << 28);

inline int norm(char c) {
  switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; }
  return 9;
}
inline int pair_score(char x, char y) { int a = norm(x), b = norm(y); return (a < 4 && a == b) ? MATCH : OTHER; }
inline int max3(int a, int b, int c) { return std::max(a, std::max(b, c)); }

// full matrices, (n+1) x (m+1); X = ends with a pair column, U = ends with a column that consumes a only,
// L = ends with a column that consumes b only
struct Full {
  int n, m; std::vector<int> X, U, L;
  int& x(int i, int j) { return X[(size_t)i * (m + 1) + j]; }
  int& u(int i, int j) { return U[(size_t)i * (m + 1) + j]; }
  int& l(int i, int j) { return L[(size_t)i * (m + 1) + j]; }
  int best(int i, int j) { return max3(x(i, j), u(i, j), l(i, j)); }
};

void fill(Full& F, const char* a, int n, const char* b, int m) {
  F.n = n; F.m = m;
  size_t cells = (size_t)(n + 1) * (m + 1);
  F.X.assign(cells, NINF); F.U.assign(cells, NINF); F.L.assign(cells, NINF);
  F.x(0, 0) = 0;
  for (int i = 0; i <= n; i++)
    for (int j = 0; j <= m; j++) {
      if (i == 0 && j == 0) continue;
      if (i > 0 && j > 0) { int p = F.best(i - 1, j - 1); if (p > NINF / 2) F.x(i, j) = p + pair_score(a[i - 1], b[j - 1]); }
      if (i > 0) {
        int v = NINF;
        if (F.u(i - 1, j) > NINF / 2) v = std::max(v, F.u(i - 1, j) + GAP_NEXT);
        if (F.x(i - 1, j) > NINF / 2) v = std::max(v, F.x(i - 1, j) + GAP_FIRST);
        if (F.l(i - 1, j) > NINF / 2) v = std::max(v, F.l(i - 1, j) + GAP_FIRST);
        F.u(i, j) = v;
      }
      if (j > 0) {
        int v = NINF;
        if (F.l(i, j - 1) > NINF / 2) v = std::max(v, F.l(i, j - 1) + GAP_NEXT);
        if (F.x(i, j - 1) > NINF / 2) v = std::max(v, F.x(i, j - 1) + GAP_FIRST);
        if (F.u(i, j - 1) > NINF / 2) v = std::max(v, F.u(i, j - 1) + GAP_FIRST);
        F.l(i, j) = v;
      }
    }
}
}  // namespace

extern "C" {

int bf_global(const char* a, int n, const char* b, int m) {
  Full F; fill(F, a, n, b, m);
  return F.best(n, m);
}

// out: [0] i, [1] j of the end cell, [2] its score, [3] 1 when the walk reached the last anti-diagonal
void bf_extend(const char* a, int n, const char* b, int m, int breaklen, int64_t* out) {
  Full F; fill(F, a, n, b, m);
  int high = 0, fin_d = 0, fin_j = 0, d = 1;
  for (; d <= n + m && d - fin_d <= breaklen; d++) {
    int jlo = std::max(0, d - n), jhi = std::min(d, m);
    for (int j = jlo; j <= jhi; j++) {
      int v = F.best(d - j, j);
      if (v > NINF / 2 && v >= high) { high = v; fin_d = d; fin_j = j; }
    }
  }
  out[0] = fin_d - fin_j; out[1] = fin_j; out[2] = high; out[3] = (d - 1 == n + m) ? 1 : 0;
}

// ops: 'M' pair column, 'I' column consuming a only, 'D' column consuming b only.  out: [0] score, [1] error columns,
// [2] bases of a consumed, [3] bases of b consumed, [4] columns
void bf_score_ops(const char* a, const char* b, const char* ops, int64_t* out) {
  int64_t score = 0, err = 0, i = 0, j = 0, cols = 0; char prev = 'M';
  for (const char* p = ops; *p; p++, cols++) {
    if (*p == 'M') { int s = pair_score(a[i], b[j]); score += s; if (norm(a[i]) != norm(b[j])) err++; i++; j++; }
    else if (*p == 'I') { score += (prev == 'I') ? GAP_NEXT : GAP_FIRST; err++; i++; }
    else { score += (prev == 'D') ? GAP_NEXT : GAP_FIRST; err++; j++; }
    prev = *p;
  }
  out[0] = score; out[1] = err; out[2] = i; out[3] = j; out[4] = cols;
}

int bf_edit_distance(const char* a, int n, const char* b, int m) {
  std::vector<int> prev(m + 1), cur(m + 1);
  for (int j = 0; j <= m; j++) prev[j] = j;
  for (int i = 1; i <= n; i++) {
    cur[0] = i;
    int ai = norm(a[i - 1]);
    for (int j = 1; j <= m; j++) {
      int bj = norm(b[j - 1]);
      int sub = prev[j - 1] + ((ai < 4 && ai == bj) ? 0 : 1);
      cur[j] = std::min(sub, std::min(prev[j], cur[j - 1]) + 1);
    }
    prev.swap(cur);
  }
  return prev[m];
}

}  // extern "C"
