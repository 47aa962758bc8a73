/* mfsc.h -- C ABI of libmfsc_b200.so, the B200 (sm_100a) alignment + coverage library that stands in
 * for the `nucmer --maxmatch` / `show-coords -r` processes MetaFinisherSC shells out to, and for the
 * Python per-base coverage loop that consumes their rows.
 *
 * Boundary replaced (reference kakitone/MetaFinisherSC, paths relative to its root):
 *   srcRefactor/repeatPhaserLib/finisherSCCoreLib/alignerRobot.py:46-64   nucmerMummer   -> mfsc_index_build + mfsc_align_batch
 *   srcRefactor/repeatPhaserLib/finisherSCCoreLib/alignerRobot.py:66-72   showCoorMummer -> mfsc_write_coords
 *   srcRefactor/repeatPhaserLib/finisherSCCoreLib/alignerRobot.py:7-40    extractMumData -> mfsc_result_rows (rows served from memory)
 *   srcRefactor/misassemblyFixerLib/adaptorFix.py:21-22                   nucmer default -> params.maxmatch = 0
 *   srcRefactor/misassemblyFixerLib/merger.py:126-140                     assignCoverage -> mfsc_coverage
 * The reference has no FFI of its own (it calls os.system); INTEGRATION.md shows the ctypes stub a
 * maintainer would add.  Plain pointers and sizes only; every function returns 0 on success and a
 * negative code otherwise, mfsc_last_error() describes the failure.  There is no CPU fallback: without
 * a CUDA device every entry point that computes returns MFSC_ERR_CUDA.
 *
 * Threading: a handle may be used by one thread at a time; distinct handles are independent.
 */
#ifndef MFSC_H
#define MFSC_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFSC_OK 0
#define MFSC_ERR_ARG (-1)
#define MFSC_ERR_CUDA (-2)
#define MFSC_ERR_NOMEM (-3)
#define MFSC_ERR_IO (-4)

/* nucmer options (alignerRobot.py:46-64: defaults, `-b 50` with globalFast, `-l 10` with refinedVersion) */
typedef struct mfsc_params {
  int32_t minmatch;    /* -l  20 */
  int32_t mincluster;  /* -c  65 */
  int32_t maxgap;      /* -g  90 */
  int32_t breaklen;    /* -b  200 */
  int32_t diagdiff;    /* -D  5 */
  int32_t maxmatch;    /* 1: --maxmatch, 0: default (matches unique in the reference) */
  int32_t simplify;    /* 1: default, drop shadowed clusters */
  int32_t band_cap;    /* widest DP band kept, <= 248 */
  double diagfactor;   /* -d  0.12 */
  int32_t kmer;        /* index k-mer length, 0 = choose from the reference size */
  int32_t engine;      /* stage-4 DP engine: 0 = packed 32-bit states first, reads whose fields do not fit are extended
                          again by the general engine (same rows either way); 1 = general engine only.  With the packed
                          engine the forward extensions that do not depend on the walk run as independent tasks ahead
                          of it (stage 4a) when the batch has at least 1000 matches per read (reads against reads,
                          repeats); 2 = packed, stage 4a always; 3 = packed, never.  The rows are the same in every mode */
  int32_t pk_u_span;   /* packed engine: widest spread of the gap-column / mismatch counters over a band that a */
  int32_t pk_m_span;   /* rebase accepts; 0 = the field capacity.  Smaller values only force more reads to engine 1 */
} mfsc_params;

typedef struct mfsc_index mfsc_index;     /* contig index resident in HBM */
typedef struct mfsc_reads mfsc_reads;     /* read batch resident in HBM */
typedef struct mfsc_result mfsc_result;   /* rows (and optional intermediate stages) of one batch */

const char* mfsc_last_error(void);
int mfsc_version(void);
int mfsc_device_count(int32_t* n);
int mfsc_set_device(int32_t device);
/* The library keeps one stream per calling thread (created by the thread's first call).  A worker thread that is about to
 * end releases it with this call; the next call of the thread would create a new one. */
int mfsc_thread_release(void);
/* Scratch of finished calls stays in the device's stream-ordered memory pool for the next call (no release threshold).  This
 * gives it back to the driver: between two phases of very different batch sizes, or before handing the GPU to someone else. */
int mfsc_trim_pool(void);
void mfsc_default_params(mfsc_params* p);

/* pinned host memory for the buffers handed to the calls below (optional, speeds up the copies) */
int mfsc_host_alloc(void** p, int64_t bytes);
int mfsc_host_free(void* p);

/* stage 1.  bases: concatenated record characters (ACGT any case; anything else never matches),
 * rec_off[n_rec+1]: offsets of the records in bases. */
int mfsc_index_build(const uint8_t* bases, const int64_t* rec_off, int32_t n_rec, const mfsc_params* p,
                     mfsc_index** out);
/* mfsc_index_build with its round-1 signature: build_tables must be 1 (empty shells for a broadcast through other means are no
 * longer made here -- replicas come from mfsc_index_broadcast / mfsc_index_build_sharded below); n_positions is ignored.
 * mfsc_index_info: [0] k, [1] device, [2] text words, [3] occupied buckets, [4] positions, [5] bytes in HBM, [6] records, [7] bases. */
int mfsc_index_create(const uint8_t* bases, const int64_t* rec_off, int32_t n_rec, const mfsc_params* p,
                      int32_t build_tables, int64_t n_positions, mfsc_index** out);
int mfsc_index_info(const mfsc_index* ix, int64_t info[8]);
/* ptr/bytes[5]: packed text, invalid mask, bucket directory (8 B per 32 buckets), positions, heads of the occupied buckets */
int mfsc_index_buffers(const mfsc_index* ix, void* ptr[5], int64_t bytes[5]);
/* ms[0]: the whole build call (upload, packing, tables); ms[1]: the table kernels alone (count, scan, fill, occupancy) */
int mfsc_index_build_ms(const mfsc_index* ix, double ms[2]);
int mfsc_index_free(mfsc_index* ix);

/* Index replication over the GPUs of one box (the reference's analogue is the fan-out of its 20 jobs over
 * Pool(nProc) / MPI ranks, alignerRobot.py:114-166, each of which rebuilds the suffix tree of the whole reference): the
 * index is built ONCE and every other GPU receives it with one ncclBroadcast per buffer over NVLink / NVSwitch.
 *   one process, several GPUs (alignerRobot.useMummerAlignBatch with houseKeeper.globalGPUs > 1):
 *     mfsc_comm_init_local(devices, n, &c);  ix[root] = the built index (it lives on devices[root]);
 *     mfsc_index_broadcast(c, root, ix, &ms) creates ix[i] on devices[i] for every i != root and fills it.
 *   one process per GPU (torchrun): rank 0 calls mfsc_comm_unique_id, the launcher ships the 128 bytes to the other ranks,
 *     every rank calls mfsc_comm_init_rank after mfsc_set_device; mfsc_index_broadcast(c, root, &ix, &ms) with ix = the built
 *     index on the root rank and NULL elsewhere (created there).
 * NCCL is bound at run time (libnccl.so.2); without it these calls return MFSC_ERR_CUDA.  Replicas are freed with
 * mfsc_index_free from any thread. */
typedef struct mfsc_comm mfsc_comm;
int mfsc_comm_unique_id(uint8_t id[128]);
int mfsc_comm_init_rank(const uint8_t id[128], int32_t rank, int32_t world, mfsc_comm** out);
int mfsc_comm_init_local(const int32_t* devices, int32_t n, mfsc_comm** out);
int mfsc_index_broadcast(mfsc_comm* c, int32_t root, mfsc_index** ix, double* ms);
int mfsc_comm_free(mfsc_comm* c);
/* Sharded build, one process per GPU: every rank calls it with the same contigs; rank r builds the tables of its 1/N of the
 * k-mer space and the slices are exchanged with grouped NCCL broadcasts (one root per slice).  Every rank ends with the
 * complete index.  ms (optional): [0] upload + packing, [1] own slice, [2] exchange, [3] total (host clock).  With a
 * communicator of one rank this is mfsc_index_build. */
int mfsc_index_build_sharded(mfsc_comm* c, const uint8_t* bases, const int64_t* rec_off, int32_t n_rec, const mfsc_params* p,
                             mfsc_index** out, double ms[4]);

/* read batch upload (host -> HBM, packs both strands) */
int mfsc_reads_upload(const uint8_t* bases, const int64_t* read_off, int32_t n_reads, mfsc_reads** out);
int mfsc_reads_free(mfsc_reads* r);

#define MFSC_KEEP_MEMS 1       /* keep the sorted maximal exact matches for mfsc_result_mems */
#define MFSC_KEEP_CLUSTERS 2   /* keep the clusters for mfsc_result_clusters */
#define MFSC_STOP_SEEDS 4      /* stop after stage 2 */
#define MFSC_STOP_CLUSTERS 8   /* stop after stage 3 */
#define MFSC_WANT_DELTAS 16    /* also produce the delta (indel offset) list of every row: runs the traceback-recording
                                  instantiation of the extension kernel, about 2x slower and 5 MB of HBM per resident warp */

/* stages 2-4 on a resident batch */
int mfsc_align_reads(mfsc_index* ix, mfsc_reads* reads, const mfsc_params* p, int32_t flags, mfsc_result** out);
/* upload + align + rows back on the host: the call that replaces one nucmer + show-coords job */
int mfsc_align_batch(mfsc_index* ix, const uint8_t* bases, const int64_t* read_off, int32_t n_reads,
                     const mfsc_params* p, int32_t flags, mfsc_result** out);

#define MFSC_N_ROWS 0
#define MFSC_N_MEMS 1
#define MFSC_N_CLUSTERS 2
#define MFSC_N_CLUSTER_MATCHES 3
#define MFSC_N_DELTAS 4
int64_t mfsc_result_count(const mfsc_result* r, int32_t what);
/* rows in delta order (read, then contig, then cluster walk); s2/e2 on the forward read, s2 > e2 for
 * reverse-strand rows; errors/cols give %IDY = 100*(cols-errors)/cols as show-coords computes it */
int mfsc_result_rows(const mfsc_result* r, int32_t* ref_id, int32_t* qry_id, int32_t* s1, int32_t* e1,
                     int32_t* s2, int32_t* e2, int32_t* errors, int32_t* cols);
/* matches sorted by (read, strand, s2, s1): read, strand, s1 (0-based in the separator-joined
 * reference), s2 (0-based on the strand), len */
int mfsc_result_mems(const mfsc_result* r, int32_t* qry_id, int32_t* strand, int64_t* s1, int32_t* s2, int32_t* len);
/* clusters (split at record changes) in emission order: read, strand, record, first match, count;
 * matches: s1, s2, len as above */
int mfsc_result_clusters(const mfsc_result* r, int32_t* qry_id, int32_t* strand, int32_t* rec, int32_t* first,
                         int32_t* count, int64_t* m_s1, int32_t* m_s2, int32_t* m_len);
/* delta lists (MFSC_WANT_DELTAS): row i owns deltas[begin[i] .. end[i]), MUMmer's signed distances to the next
 * indel (positive: reference base without partner, negative: read base without partner) */
int mfsc_result_deltas(const mfsc_result* r, int64_t* begin, int64_t* end, int32_t* deltas);
/* stats[16]: 0 DP cells, 1 DP calls, 2 widest band, 3 probes, 4 matches, 5 kernel launches,
 * 6 algorithmic bytes of the seed kernel, 7 h2d bytes, 8 d2h bytes, 9 reads handed from the packed to the general DP engine,
 * 10 anti-diagonals on which band_cap trimmed the band (0: the cap never bound, the rows are those of an uncapped DP),
 * 11 gaps between cluster matches aligned as independent tasks ahead of the walk, 12 how many of those results the walk used;
 * ms[16]: 0 upload+pack, 1 seeds, 2 sort, 3 cluster, 4 extend, 5 compact+download, 6 total */
int mfsc_result_stats(const mfsc_result* r, int64_t stats[16], double ms[16]);
int mfsc_result_free(mfsc_result* r);

/* stage 5.  rows selected by the caller (merger.py:171-187): +1 over [s1-1, e1) of contig ref_id.
 * contig_off[n_contig+1]; cov_out[contig_off[n_contig]]. ms (optional): [0] kernels, [1] total. */
int mfsc_coverage(const int32_t* ref_id, const int32_t* s1, const int32_t* e1, int64_t n_rows,
                  const int64_t* contig_off, int32_t n_contig, int32_t* cov_out, double ms[2]);

/* The same profile kept resident in HBM for its consumer, the group C-test of the misassembly fixer
 * (misassemblyFixerLib/mergerHelper.py:137-163 formatIntervalAndFilter: `np.sum(coverageData[start:end])` per candidate
 * interval; merger.py:189-190 -> mergerHelper.py:84-85 is the 10^7-integer JSON round trip this avoids).
 * mfsc_cov_interval_sums: sums[q] = sum(cov[contig[q]][start[q]:end[q]]) with Python slice semantics (negative indices
 * count from the contig end, out-of-range indices are clamped, an empty slice sums to 0), 64-bit.
 * mfsc_cov_fetch: the profile of one contig (contig >= 0) or of all contigs concatenated (contig = -1). */
typedef struct mfsc_cov mfsc_cov;
int mfsc_cov_build(const int32_t* ref_id, const int32_t* s1, const int32_t* e1, int64_t n_rows,
                   const int64_t* contig_off, int32_t n_contig, mfsc_cov** out, double ms[2]);
int mfsc_cov_fetch(const mfsc_cov* c, int32_t contig, int32_t* cov_out);
int mfsc_cov_interval_sums(const mfsc_cov* c, const int32_t* contig, const int64_t* start, const int64_t* end, int64_t n,
                           int64_t* sums);
int mfsc_cov_free(mfsc_cov* c);

/* Row predicates of the reference's consumers (file:line in csrc/mfsc_rowflags.cuh), evaluated for every row at once:
 * flags[r] = OR of the MFSC_RF_* bits that hold for row r.  ref_len / qry_len: record lengths; ref_gid / qry_gid: an
 * id per record such that two records carry the same name exactly when their ids are equal (the reference compares
 * the name strings of columns -2 and -1). */
#define MFSC_RF_HEADTAIL 1u
#define MFSC_RF_IDENTICAL 2u
#define MFSC_RF_ASSOCIATED 4u
#define MFSC_RF_EMBEDDED_LR 8u
#define MFSC_RF_REF_COVERED 16u
#define MFSC_RF_REF_HEAD 32u
#define MFSC_RF_SHORT_REF 64u
#define MFSC_RF_SHORT_QRY 128u
#define MFSC_RF_DIFF_NAME 256u
#define MFSC_RF_END_SPAN 512u
#define MFSC_RF_OVERLAP 1024u
int mfsc_row_flags(int64_t n_rows, const int32_t* ref_id, const int32_t* qry_id, const int32_t* s1, const int32_t* e1,
                   const int32_t* s2, const int32_t* e2, const int32_t* ref_len, int32_t n_ref, const int32_t* qry_len,
                   int32_t n_qry, const int32_t* ref_gid, const int32_t* qry_gid, uint32_t* flags);

/* show-coords -r text (5 header lines, rows "%8d %8d  | %8d %8d  | %8d %8d  | %8.2f  | %s\t%s") and the
 * delta file of a row set.  names: NUL-separated, one per record.  The delta carries the alignment
 * headers (`>ref qry lenR lenQ`, `sR eR sQ eQ errors simErrors 0`) followed by the indel offset list when the
 * batch was aligned with MFSC_WANT_DELTAS.  Without it (the default, traceback-free DP) the file is NOT a MUMmer delta
 * and says so in a `#MFSC rows-only` third line: every alignment is its header alone, with the alignment column count
 * in the 7th field (MUMmer writes stop codons there, always 0 for nucmer) so that %IDY stays reproducible from the file. */
int mfsc_write_coords(const char* path, const char* ref_path, const char* qry_path, int64_t n_rows,
                      const int32_t* ref_id, const int32_t* qry_id, const int32_t* s1, const int32_t* e1,
                      const int32_t* s2, const int32_t* e2, const int32_t* errors, const int32_t* cols,
                      const char* ref_names, int32_t n_ref, const char* qry_names, int32_t n_qry, int32_t sort_by_ref);
int mfsc_write_delta(const char* path, const char* ref_path, const char* qry_path, int64_t n_rows,
                     const int32_t* ref_id, const int32_t* qry_id, const int32_t* s1, const int32_t* e1,
                     const int32_t* s2, const int32_t* e2, const int32_t* errors, const int32_t* cols,
                     const char* ref_names, const int32_t* ref_len, int32_t n_ref,
                     const char* qry_names, const int32_t* qry_len, int32_t n_qry,
                     const int64_t* d_begin, const int64_t* d_end, const int32_t* deltas /* all three NULL: headers only */);

/* FASTA ingest and part writing on the host (the O(bases) work either side of the aligner: fasta-splitter.pl:135-183,
 * IORobot.py:8-30).  mfsc_fasta_parse: data[n] = file bytes; bases_out[<= n] receives the concatenated sequence
 * characters, off_out[max_rec+1] the record offsets, name_span_out[2*max_rec] the [begin, end) of every header text
 * (after '>') inside data.  Blank lines are skipped, '\r' is stripped, sequence lines before the first header are
 * ignored.  Returns MFSC_ERR_ARG when there are more than max_rec records (n_rec_out then holds the count needed).
 * mfsc_fasta_write: records [rec_begin, rec_end) as ">name\n" + sequence wrapped at line_len columns (0: one line). */
int mfsc_fasta_parse(const uint8_t* data, int64_t n, uint8_t* bases_out, int64_t* off_out, int64_t* name_span_out,
                     int32_t max_rec, int32_t* n_rec_out);
int mfsc_fasta_count(const uint8_t* data, int64_t n, int32_t* n_rec_out);
int mfsc_fasta_write(const char* path, const char* names, const int64_t* name_off, const uint8_t* bases, const int64_t* off,
                     int32_t rec_begin, int32_t rec_end, int32_t line_len);

#ifdef __cplusplus
}
#endif
#endif
