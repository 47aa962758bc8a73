// mfsc_rowflags.cuh -- the per-row predicates that the reference's consumers evaluate over every alignment row in
// Python loops, one thread per row over the SoA row buffer: 28 B read + 4 B written per row.
//   MFSC_RF_HEADTAIL      abunHouseKeeper.headTailMatch   repeatPhaserLib/abunHouseKeeper.py:177-202  (thres 10)
//   MFSC_RF_IDENTICAL     abunHouseKeeper.identicalItem   repeatPhaserLib/abunHouseKeeper.py:162-175  (thres 20)
//   MFSC_RF_ASSOCIATED    associatedReadFinder.checkSatisfy  repeatPhaserLib/associatedReadFinder.py:16-43 (threshold 50;
//                         the reference tests columns 0-1 against the length of the name in column -1 and columns 2-3
//                         against the length of the name in column -2, which is kept as written)
//   MFSC_RF_EMBEDDED_LR   merger.completelyEmbedSR2TR, LR mode  misassemblyFixerLib/merger.py:113-123 (thres 20)
//   MFSC_RF_REF_COVERED   readContigGraphFormer.formExtraEdges  repeatPhaserLib/readContigGraphFormer.py:173-176 (LEN1 > len - 300)
//   MFSC_RF_REF_HEAD      readContigGraphFormer.addDataToList   repeatPhaserLib/readContigGraphFormer.py:14-30  (S1 < 50)
//   MFSC_RF_SHORT_REF / MFSC_RF_SHORT_QRY / MFSC_RF_DIFF_NAME   nonRedundantResolver.py:100-112 (|len - LEN| < 10, name1 != name2)
//   MFSC_RF_END_SPAN      gapFiller.formRelatedReadsFile  finisherSCCoreLib/gapFiller.py:116-118,177 (S1 < 10 and E1 > 400 - 10)
//   MFSC_RF_OVERLAP       IORobot.align  finisherSCCoreLib/IORobot.py:529-538 (E1 > len(ref) - 10 and S2 < 10)
#pragma once
#include "mfsc_platform.cuh"
#include "../../include/mfsc.h"

MFSC_KERNEL k_row_flags(long long n, const int* ref_id, const int* qry_id, const int* s1, const int* e1, const int* s2, const int* e2,
                        const int* ref_len, const int* qry_len, const int* ref_gid, const int* qry_gid, unsigned* flags) {
  for (long long r = thread_global(); r < n; r += threads_total()) {
    const int a = ref_id[r], b = qry_id[r];
    const int S1 = s1[r], E1 = e1[r], S2 = s2[r], E2 = e2[r];
    const int l1 = ref_len[a], l2 = qry_len[b];
    const int LEN1 = (E1 > S1 ? E1 - S1 : S1 - E1) + 1, LEN2 = (E2 > S2 ? E2 - S2 : S2 - E2) + 1;
    const int lo2 = S2 < E2 ? S2 : E2, hi2 = S2 < E2 ? E2 : S2;
    const bool diff = ref_gid[a] != qry_gid[b];
    unsigned f = 0;
    if (diff && !(S2 > E2) && ((S1 <= 10 && E2 >= l2 - 10) || (E1 >= l1 - 10 && S2 <= 10))) f |= MFSC_RF_HEADTAIL;
    {
      const int d1 = S1 - S2, endPt1 = l1 - E1, endPt2 = l2 - E2, d2 = endPt1 - endPt2;
      if ((d1 < 0 ? -d1 : d1) < 20 && (d2 < 0 ? -d2 : d2) < 20) f |= MFSC_RF_IDENTICAL;
    }
    if ((S1 < 50 || E1 > l2 - 50) && (lo2 < 50 || hi2 > l1 - 50)) f |= MFSC_RF_ASSOCIATED;
    if (lo2 < 20 && hi2 >= l2 - 20) f |= MFSC_RF_EMBEDDED_LR;
    if (LEN1 > l1 - 300) f |= MFSC_RF_REF_COVERED;
    if (S1 < 50) f |= MFSC_RF_REF_HEAD;
    { const int x = l1 - LEN1; if ((x < 0 ? -x : x) < 10) f |= MFSC_RF_SHORT_REF; }
    { const int x = l2 - LEN2; if ((x < 0 ? -x : x) < 10) f |= MFSC_RF_SHORT_QRY; }
    if (diff) f |= MFSC_RF_DIFF_NAME;
    if (S1 < 10 && E1 > 390) f |= MFSC_RF_END_SPAN;
    if (E1 > l1 - 10 && S2 < 10) f |= MFSC_RF_OVERLAP;
    flags[r] = f;
  }
}
