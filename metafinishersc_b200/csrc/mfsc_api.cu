// mfsc_api.cu -- host side of libmfsc_b200.so: HBM residency, stage orchestration on one stream,
// and the C ABI declared in include/mfsc.h.  See DESIGN.md for the data layout and per-kernel
// rooflines.  (tests/emu builds this same file with -DMFSC_EMU for CPU-side logic tests.)
#include "../../include/mfsc.h"
#include "mfsc_extend.cuh"
#include "mfsc_gaps.cuh"
#include "mfsc_coverage.cuh"
#include "mfsc_rowflags.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <string>
#include <thread>
#include <vector>

#ifndef MFSC_EMU
#include <cub/cub.cuh>
#endif

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#ifdef MFSC_EMU
thread_local emu_dim3 emu_threadIdx, emu_blockIdx, emu_blockDim, emu_gridDim;
#endif

// ------------------------------------------------------------------------------------------
// device runtime wrappers
// ------------------------------------------------------------------------------------------
namespace {

#ifdef MFSC_EMU
struct Dev {
  cudaStream_t stream = 1;
  long long launches = 0;
  int device = 0;
  bool ok() { return true; }
  void* alloc(size_t bytes) { return calloc(bytes ? bytes : 1, 1); }
  void free_(void* p) { free(p); }
  void h2d(void* d, const void* h, size_t n) { if (n) memcpy(d, h, n); }
  void d2h(void* h, const void* d, size_t n) { if (n) memcpy(h, d, n); }
  void zero(void* d, size_t n) { if (n) memset(d, 0, n); }
  void sync() {}
  int n_sm() { return 4; }
};
static bool dev_check(std::string&) { return true; }
struct Stamp { std::chrono::steady_clock::time_point t; };
static void stamp(Dev&, Stamp& s) { s.t = std::chrono::steady_clock::now(); }
static double elapsed_ms(Dev&, Stamp& a, Stamp& b) { return std::chrono::duration<double, std::milli>(b.t - a.t).count(); }
static void stamp_free(Stamp&) {}
#else
struct Dev {
  cudaStream_t stream = 0;
  long long launches = 0;
  int sm = 0;
  int device = -1;
  bool ok() {
    if (stream) return true;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return false; }
    if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return false; }
    int d = 0; cudaGetDevice(&d);
    device = d;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, d);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    return true;
  }
  void* alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocAsync(&p, bytes ? bytes : 16, stream) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
  }
  void free_(void* p) { if (p) cudaFreeAsync(p, stream); }
  // large uploads go in 8 MB pieces: the copy engine arbitrates between streams per operation, so a small upload of another
  // stream (the contigs of an index build next to a 1.5 GB read batch) waits for one piece, not for the whole batch
  void h2d(void* d, const void* h, size_t n) {
    const size_t piece = (size_t)8 << 20;
    for (size_t o = 0; o < n; o += piece) cudaMemcpyAsync((char*)d + o, (const char*)h + o, n - o < piece ? n - o : piece, cudaMemcpyHostToDevice, stream);
  }
  void d2h(void* h, const void* d, size_t n) { if (n) { cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, stream); cudaStreamSynchronize(stream); } }
  void zero(void* d, size_t n) { if (n) cudaMemsetAsync(d, 0, n, stream); }
  void sync() { cudaStreamSynchronize(stream); }
  int n_sm() { return sm > 0 ? sm : 148; }
};
static bool dev_check(std::string& msg) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { msg = cudaGetErrorString(e); return false; }
  return true;
}
struct Stamp { cudaEvent_t e = nullptr; };
static void stamp(Dev& D, Stamp& s) { if (!s.e) cudaEventCreate(&s.e); cudaEventRecord(s.e, D.stream); }
static double elapsed_ms(Dev&, Stamp& a, Stamp& b) { float ms = 0; cudaEventSynchronize(b.e); cudaEventElapsedTime(&ms, a.e, b.e); return ms; }
static void stamp_free(Stamp& s) { if (s.e) cudaEventDestroy(s.e); s.e = nullptr; }
#endif

static thread_local Dev g_dev;

#ifdef MFSC_EMU
struct DeviceGuard { explicit DeviceGuard(int) {} };
static void dev_free_plain(void* p) { free(p); }
#else
// switches to a device for the lifetime of the object (frees and copies that concern another device than the thread's own)
struct DeviceGuard {
  int prev = -1; bool sw = false;
  explicit DeviceGuard(int dev) {
    if (dev < 0) return;
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; return; }
    if (prev != dev) { sw = cudaSetDevice(dev) == cudaSuccess; if (!sw) cudaGetLastError(); }
  }
  ~DeviceGuard() { if (sw && prev >= 0) cudaSetDevice(prev); }
};
static void dev_free_plain(void* p) { if (p && cudaFree(p) != cudaSuccess) cudaGetLastError(); }   // also valid for stream-ordered allocations (synchronises)
#endif

template <class T> T* dalloc(Dev& D, size_t n) { return (T*)D.alloc(n * sizeof(T)); }

// grid sizes: a multiple of the SM count (148 on B200)
static unsigned grid_for(Dev& D, long long work_items, int block, int per_sm) {
  long long need = (work_items + block - 1) / block;
  long long cap = (long long)D.n_sm() * per_sm;
  if (need < 1) need = 1;
  return (unsigned)(need < cap ? need : cap);
}

// stable sort of (key, value) pairs, ascending, on bits [0, end_bit)
static bool sort_pairs_u32(Dev& D, unsigned* key, unsigned* val, unsigned* key_alt, unsigned* val_alt, long long n, int end_bit,
                           unsigned** key_out, unsigned** val_out) {
#ifdef MFSC_EMU
  (void)end_bit;
  std::vector<unsigned> idx(n);
  for (long long i = 0; i < n; i++) idx[i] = (unsigned)i;
  std::stable_sort(idx.begin(), idx.end(), [&](unsigned a, unsigned b) { return key[a] < key[b]; });
  for (long long i = 0; i < n; i++) { key_alt[i] = key[idx[i]]; val_alt[i] = val[idx[i]]; }
  *key_out = key_alt; *val_out = val_alt;
  return true;
#else
  cub::DoubleBuffer<unsigned> k(key, key_alt), v(val, val_alt);
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k, v, (int)n, 0, end_bit, D.stream);
  void* tmp = D.alloc(tmp_bytes);
  if (!tmp) return false;
  cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k, v, (int)n, 0, end_bit, D.stream);
  D.free_(tmp);
  D.launches += 4;
  *key_out = k.Current(); *val_out = v.Current();
  return true;
#endif
}

static bool sort_pairs_u64(Dev& D, unsigned long long* key, unsigned* val, unsigned long long* key_alt, unsigned* val_alt, long long n,
                           int end_bit, unsigned** val_out) {
#ifdef MFSC_EMU
  (void)end_bit;
  std::vector<unsigned> idx(n);
  for (long long i = 0; i < n; i++) idx[i] = (unsigned)i;
  std::stable_sort(idx.begin(), idx.end(), [&](unsigned a, unsigned b) { return key[a] < key[b]; });
  for (long long i = 0; i < n; i++) { key_alt[i] = key[idx[i]]; val_alt[i] = val[idx[i]]; }
  *val_out = val_alt;
  return true;
#else
  cub::DoubleBuffer<unsigned long long> k(key, key_alt);
  cub::DoubleBuffer<unsigned> v(val, val_alt);
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k, v, (int)n, 0, end_bit, D.stream);
  void* tmp = D.alloc(tmp_bytes);
  if (!tmp) return false;
  cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k, v, (int)n, 0, end_bit, D.stream);
  D.free_(tmp);
  D.launches += 8;
  *val_out = v.Current();
  return true;
#endif
}

static bool inclusive_sum_u32(Dev& D, unsigned* a, long long n) {
#ifdef MFSC_EMU
  unsigned acc = 0;
  for (long long i = 0; i < n; i++) { acc += a[i]; a[i] = acc; }
  return true;
#else
  size_t tmp_bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, a, a, (int)n, D.stream);
  void* tmp = D.alloc(tmp_bytes);
  if (!tmp) return false;
  cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, a, a, (int)n, D.stream);
  D.free_(tmp);
  D.launches += 2;
  return true;
#endif
}

static bool exclusive_sum_i32(Dev& D, const int* in, int* out, long long n) {
#ifdef MFSC_EMU
  int acc = 0;
  for (long long i = 0; i < n; i++) { int v = in[i]; out[i] = acc; acc += v; }
  return true;
#else
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (int)n, D.stream);
  void* tmp = D.alloc(tmp_bytes);
  if (!tmp) return false;
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (int)n, D.stream);
  D.free_(tmp);
  D.launches += 2;
  return true;
#endif
}

static bool exclusive_sum_i64(Dev& D, const long long* in, long long* out, long long n) {
#ifdef MFSC_EMU
  long long acc = 0;
  for (long long i = 0; i < n; i++) { long long v = in[i]; out[i] = acc; acc += v; }
  return true;
#else
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (int)n, D.stream);
  void* tmp = D.alloc(tmp_bytes);
  if (!tmp) return false;
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (int)n, D.stream);
  D.free_(tmp);
  D.launches += 2;
  return true;
#endif
}

static int bits_for(unsigned long long v) { int b = 1; while (b < 64 && (v >> b)) b++; return b; }

// word layout of a set of sequences (see mfsc_seq.cuh)
static void layout(const int64_t* off, int n_rec, int both, std::vector<unsigned>& wstart, std::vector<unsigned>& start,
                   std::vector<int>& len, unsigned& n_words, bool& too_big) {
  const int n_seq = both ? 2 * n_rec : n_rec;
  wstart.resize(n_seq + 1); start.resize(n_seq); len.resize(n_seq);
  unsigned long long w = 1;
  for (int k = 0; k < n_seq; k++) {
    int r = both ? (k >> 1) : k;
    long long l = off[r + 1] - off[r];
    wstart[k] = (unsigned)w; start[k] = (unsigned)(w * 32); len[k] = (int)l;
    w += (unsigned long long)(l / 32 + 1);
  }
  wstart[n_seq] = (unsigned)w;
  too_big = (w + 2) * 32ull >= (1ull << 32);
  n_words = (unsigned)(w + 2);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------
struct SeqDev {
  unsigned long long* txt = nullptr; unsigned* inv = nullptr; unsigned* start = nullptr; int* len = nullptr;
  int n_seq = 0; unsigned n_words = 0; long long n_bases = 0;
  std::vector<unsigned> h_start; std::vector<int> h_len;
  SeqStore view() const { SeqStore s; s.txt = txt; s.inv = inv; s.start = start; s.len = len; s.n_seq = n_seq; s.n_words = n_words; return s; }
};

static bool seq_upload(Dev& D, const uint8_t* bases, const int64_t* off, int n_rec, int both, bool pack, SeqDev& S, std::string& err) {
  std::vector<unsigned> wstart; bool too_big = false;
  layout(off, n_rec, both, wstart, S.h_start, S.h_len, S.n_words, too_big);
  if (too_big) { err = "sequence set exceeds 2^32 packed positions"; return false; }
  S.n_seq = both ? 2 * n_rec : n_rec;
  S.n_bases = off[n_rec] - off[0];
  S.txt = dalloc<unsigned long long>(D, S.n_words);
  S.inv = dalloc<unsigned>(D, S.n_words);
  S.start = dalloc<unsigned>(D, S.n_seq + 1);
  S.len = dalloc<int>(D, S.n_seq + 1);
  if (!S.txt || !S.inv || !S.start || !S.len) { err = "out of device memory (sequence store)"; return false; }
  D.h2d(S.start, S.h_start.data(), sizeof(unsigned) * S.n_seq);
  D.h2d(S.len, S.h_len.data(), sizeof(int) * S.n_seq);
  if (pack) {
    unsigned char* d_bases = dalloc<unsigned char>(D, (size_t)S.n_bases + 1);
    long long* d_off = dalloc<long long>(D, n_rec + 1);
    unsigned* d_wstart = dalloc<unsigned>(D, S.n_seq + 1);
    if (!d_bases || !d_off || !d_wstart) { err = "out of device memory (upload)"; return false; }
    std::vector<long long> rel(n_rec + 1);
    for (int i = 0; i <= n_rec; i++) rel[i] = off[i] - off[0];
    D.h2d(d_bases, bases + off[0], (size_t)S.n_bases);
    D.h2d(d_off, rel.data(), sizeof(long long) * (n_rec + 1));
    D.h2d(d_wstart, wstart.data(), sizeof(unsigned) * (S.n_seq + 1));
    MFSC_LAUNCH(k_pack, grid_for(D, S.n_words, 256, 16), 256, 0, D.stream, d_bases, d_off, d_wstart, S.n_seq, both, S.n_words, S.txt, S.inv);
    D.launches++;
    D.sync();   // host staging vectors go out of scope
    D.free_(d_bases); D.free_(d_off); D.free_(d_wstart);
  }
  return true;
}

static void seq_free(Dev& D, SeqDev& S) {
  D.free_(S.txt); D.free_(S.inv); D.free_(S.start); D.free_(S.len);
  S.txt = nullptr; S.inv = nullptr; S.start = nullptr; S.len = nullptr;
}

struct mfsc_index {
  SeqDev ref;
  unsigned* ostart = nullptr;   // separator-joined coordinate of each record
  std::vector<unsigned> h_ostart;
  uint2* dir = nullptr;         // [n_buckets / 32 + 1] {occupancy bits, occupied buckets before the word}  (mfsc_seq.cuh)
  unsigned* heads = nullptr;    // [n_heads + 1] first position slot of every occupied bucket, then n_pos
  unsigned* pos = nullptr;      // [n_pos] text positions grouped by bucket
  int k = 0; long long n_buckets = 0, n_pos = 0, n_heads = 0;
  KmerIndex view() const { KmerIndex X; X.dir = dir; X.heads = heads; X.pos = pos; return X; }
  double build_ms = 0, tables_ms = 0;   // whole mfsc_index_create call / the table kernels alone (count, scan, fill, occupancy)
  int device = -1;              // device the buffers live on
  bool plain = false;           // buffers come from cudaMalloc (replicas made by mfsc_index_broadcast), not from the stream-ordered pool
};

struct mfsc_reads {
  SeqDev q;
  unsigned* tile_off = nullptr;   // [n_seq + 1] first seed tile of every strand (mfsc_seed.cuh)
  unsigned n_tiles = 0;
  int device = -1;
  int n_reads = 0;
  double upload_ms = 0;
  long long h2d_bytes = 0;
};

struct mfsc_result {
  // rows (host)
  std::vector<int> ref, qry, s1, e1, s2, e2, errs, cols;
  // mems (host, optional)
  std::vector<int> m_q, m_strand, m_s2, m_len; std::vector<long long> m_s1;
  // clusters (host, optional)
  std::vector<int> c_q, c_strand, c_rec, c_first, c_n, cm_s2, cm_len; std::vector<long long> cm_s1;
  // delta lists (host, optional): per row [begin, end) into deltas, -1 when unavailable
  std::vector<int> dl_b, dl_e, deltas;
  long long n_mems = 0;
  int64_t stats[16]; double ms[16];
};

template <class T, class U> static void copy_out(U* dst, const std::vector<T>& v) { if (dst) for (size_t i = 0; i < v.size(); i++) dst[i] = (U)v[i]; }

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

const char* mfsc_last_error(void) { return g_err.c_str(); }
int mfsc_version(void) { return 100; }

int mfsc_device_count(int32_t* n) {
#ifdef MFSC_EMU
  *n = 1; return MFSC_OK;
#else
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); c = 0; }
  *n = c; return MFSC_OK;
#endif
}

int mfsc_set_device(int32_t device) {
#ifdef MFSC_EMU
  (void)device; return MFSC_OK;
#else
  if (g_dev.stream) return fail(MFSC_ERR_ARG, "mfsc_set_device must be called before any other call of the thread");
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return fail(MFSC_ERR_CUDA, "cudaSetDevice failed"); }
  return MFSC_OK;
#endif
}

int mfsc_thread_release(void) {
#ifndef MFSC_EMU
  Dev& D = g_dev;
  if (D.stream) { cudaStreamSynchronize(D.stream); cudaStreamDestroy(D.stream); D.stream = 0; D.device = -1; }
#endif
  return MFSC_OK;
}

int mfsc_trim_pool(void) {
#ifndef MFSC_EMU
  Dev& D = g_dev;
  if (!D.ok()) return fail(MFSC_ERR_CUDA, "no CUDA device");
  cudaDeviceSynchronize();
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, D.device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
  cudaGetLastError();
#endif
  return MFSC_OK;
}

void mfsc_default_params(mfsc_params* p) {
  memset(p, 0, sizeof(*p));
  p->minmatch = 20; p->mincluster = 65; p->maxgap = 90; p->breaklen = 200; p->diagdiff = 5; p->maxmatch = 1; p->simplify = 1;
  p->band_cap = 248; p->diagfactor = 0.12; p->kmer = 0;
}

int mfsc_host_alloc(void** p, int64_t bytes) {
#ifdef MFSC_EMU
  *p = malloc((size_t)bytes); return *p ? MFSC_OK : MFSC_ERR_NOMEM;
#else
  if (!g_dev.ok()) return fail(MFSC_ERR_CUDA, "no CUDA device");
  if (cudaMallocHost(p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); return fail(MFSC_ERR_NOMEM, "cudaMallocHost failed"); }
  return MFSC_OK;
#endif
}
int mfsc_host_free(void* p) {
#ifdef MFSC_EMU
  free(p);
#else
  cudaFreeHost(p);
#endif
  return MFSC_OK;
}

static int check_params(const mfsc_params* p) {
  if (!p) return fail(MFSC_ERR_ARG, "params is NULL");
  if (p->minmatch < 4 || p->minmatch > 4096) return fail(MFSC_ERR_ARG, "minmatch out of range");
  if (p->band_cap < 8 || p->band_cap > DP_SLOTS - 8) return fail(MFSC_ERR_ARG, "band_cap must be in [8, 248]");
  if (p->breaklen < 1 || p->mincluster < 1 || p->maxgap < 0) return fail(MFSC_ERR_ARG, "bad clustering / extension parameter");
  return MFSC_OK;
}

static int choose_k(const mfsc_params* p, long long n_bases) {
  int k = p->kmer;
  if (k <= 0) {
    k = 8;
    // about a quarter position per bucket or fewer, up to 4^15 buckets (4.3 GB of bucket heads for references beyond
    // 67 Mbp: on a 500 Mbp reference k = 15 makes stage 2 1.7x faster than k = 14, 8.1 vs 13.7 ms per 300 Mbp of reads)
    while (k < 15 && (1ll << (2 * k)) < 4 * n_bases) k++;
  }
  if (k > p->minmatch) k = p->minmatch;
  if (k > 15) k = 15;
  if (k < 4) k = 4;
  return k;
}

// every failure path releases what it had allocated
static void index_release(Dev& D, mfsc_index* ix) {
  if (!ix) return;
  if (ix->plain || !D.stream || (ix->device >= 0 && ix->device != D.device)) {
    // buffers of another device (or of a replica made with cudaMalloc): free them synchronously on their own device
    DeviceGuard g(ix->device);
    dev_free_plain(ix->ref.txt); dev_free_plain(ix->ref.inv); dev_free_plain(ix->ref.start); dev_free_plain(ix->ref.len);
    dev_free_plain(ix->ostart); dev_free_plain(ix->dir); dev_free_plain(ix->heads); dev_free_plain(ix->pos);
  } else {
    seq_free(D, ix->ref);
    D.free_(ix->ostart); D.free_(ix->dir); D.free_(ix->heads); D.free_(ix->pos);
  }
  delete ix;
}

// Tables of the bucket slice [b_lo, b_hi) of the k-mer space (both multiples of 32) from the packed text: a dense
// histogram as scratch (zero, count, inclusive scan, fill), then the compact form -- directory words with ranks relative
// to the slice, heads relative to the slice's own position array.
struct TableSlice { uint2* dir = nullptr; unsigned* heads = nullptr; unsigned* pos = nullptr; long long n_words = 0, n_occ = 0, n_pos = 0; };

static void slice_free(Dev& D, TableSlice& S) { D.free_(S.dir); D.free_(S.heads); D.free_(S.pos); S.dir = nullptr; S.heads = nullptr; S.pos = nullptr; }

static bool build_slice(Dev& D, const SeqDev& ref, int k, unsigned long long b_lo, unsigned long long b_hi, TableSlice& S, std::string& err) {
  const long long nb = (long long)(b_hi - b_lo);
  S.n_words = nb / 32;
  unsigned* H = dalloc<unsigned>(D, nb + 2);
  unsigned* bits = dalloc<unsigned>(D, S.n_words + 1);
  unsigned* cnt = dalloc<unsigned>(D, S.n_words + 2);
  S.dir = dalloc<uint2>(D, S.n_words + 1);
  auto scratch_free = [&]() { D.free_(H); D.free_(bits); D.free_(cnt); };
  if (!H || !bits || !cnt || !S.dir) { scratch_free(); slice_free(D, S); err = "out of device memory (bucket histogram)"; return false; }
  D.zero(H, sizeof(unsigned) * (nb + 2));
  const unsigned g = grid_for(D, ref.n_words, 256, 16);
  MFSC_LAUNCH(k_index_count, g, 256, 0, D.stream, ref.txt, ref.inv, ref.n_words, k, b_lo, b_hi, H);
  if (!inclusive_sum_u32(D, H, nb + 2)) { scratch_free(); slice_free(D, S); err = "out of device memory (scan)"; return false; }
  unsigned total = 0;
  D.d2h(&total, H + nb + 1, sizeof(unsigned));
  S.n_pos = total;
  S.pos = dalloc<unsigned>(D, (size_t)total + 1);
  if (!S.pos) { scratch_free(); slice_free(D, S); err = "out of device memory (position table)"; return false; }
  MFSC_LAUNCH(k_index_fill, g, 256, 0, D.stream, ref.txt, ref.inv, ref.n_words, k, b_lo, b_hi, H, S.pos);
  const unsigned gw = grid_for(D, S.n_words, 256, 16);
  MFSC_LAUNCH(k_index_dir_bits, gw, 256, 0, D.stream, H, nb, bits, cnt);
  D.zero(cnt + S.n_words, sizeof(unsigned));
  // exclusive scan of the per-word counts = inclusive scan shifted by one word: scan in place, read the total, shift by using cnt - 1
  if (!inclusive_sum_u32(D, cnt, S.n_words + 1)) { scratch_free(); slice_free(D, S); err = "out of device memory (scan)"; return false; }
  unsigned occ = 0;
  D.d2h(&occ, cnt + S.n_words, sizeof(unsigned));
  S.n_occ = occ;
  S.heads = dalloc<unsigned>(D, (size_t)occ + 2);
  unsigned* rank = dalloc<unsigned>(D, S.n_words + 1);
  if (!S.heads || !rank) { D.free_(rank); scratch_free(); slice_free(D, S); err = "out of device memory (bucket heads)"; return false; }
  MFSC_LAUNCH(k_excl_from_incl, gw, 256, 0, D.stream, cnt, bits, S.n_words, rank);
  MFSC_LAUNCH(k_index_heads, gw, 256, 0, D.stream, H, nb, bits, rank, 0u, 0u, S.dir, S.heads);
  D.launches += 5;
  D.sync();
  D.free_(rank);
  scratch_free();
  return true;
}

int mfsc_index_create(const uint8_t* bases, const int64_t* rec_off, int32_t n_rec, const mfsc_params* p, int32_t build_tables,
                      int64_t n_positions, mfsc_index** out) {
  (void)n_positions;
  if (!rec_off || n_rec <= 0 || !out || !bases) return fail(MFSC_ERR_ARG, "mfsc_index_create: bad arguments");
  if (!build_tables) return fail(MFSC_ERR_ARG, "mfsc_index_create: empty shells are no longer made here; replicas come from mfsc_index_broadcast");
  int rc = check_params(p); if (rc) return rc;
  Dev& D = g_dev;
  if (!D.ok()) return fail(MFSC_ERR_CUDA, "no CUDA device (libmfsc_b200 has no CPU fallback)");
  mfsc_index* ix = new mfsc_index();
  ix->device = D.device;
  std::string err;
  Stamp t0, t1;
  stamp(D, t0);
  if (!seq_upload(D, bases, rec_off, n_rec, 0, true, ix->ref, err)) { index_release(D, ix); return fail(MFSC_ERR_NOMEM, err); }
  ix->h_ostart.resize(n_rec);
  for (int r = 0; r < n_rec; r++) ix->h_ostart[r] = (unsigned)(rec_off[r] - rec_off[0] + r);
  ix->ostart = dalloc<unsigned>(D, n_rec);
  ix->k = choose_k(p, ix->ref.n_bases);
  ix->n_buckets = 1ll << (2 * ix->k);
  if (!ix->ostart) { index_release(D, ix); return fail(MFSC_ERR_NOMEM, "out of device memory (record table)"); }
  D.h2d(ix->ostart, ix->h_ostart.data(), sizeof(unsigned) * n_rec);
  Stamp tk;
  stamp(D, tk);
  TableSlice S;
  if (!build_slice(D, ix->ref, ix->k, 0ull, (unsigned long long)ix->n_buckets, S, err)) { index_release(D, ix); return fail(MFSC_ERR_NOMEM, err); }
  ix->dir = S.dir; ix->heads = S.heads; ix->pos = S.pos; ix->n_pos = S.n_pos; ix->n_heads = S.n_occ;
  {
    const unsigned np = (unsigned)S.n_pos;
    D.h2d(ix->heads + S.n_occ, &np, sizeof(unsigned));      // terminal entry
    D.sync();
  }
  stamp(D, t1);
  D.sync();
  ix->build_ms = elapsed_ms(D, t0, t1);
  ix->tables_ms = elapsed_ms(D, tk, t1);
  stamp_free(t0); stamp_free(t1); stamp_free(tk);
  std::string msg;
  if (!dev_check(msg)) { index_release(D, ix); return fail(MFSC_ERR_CUDA, "index build: " + msg); }
  *out = ix;
  return MFSC_OK;
}

int mfsc_index_build(const uint8_t* bases, const int64_t* rec_off, int32_t n_rec, const mfsc_params* p, mfsc_index** out) {
  return mfsc_index_create(bases, rec_off, n_rec, p, 1, 0, out);
}

int mfsc_index_info(const mfsc_index* ix, int64_t info[8]) {
  if (!ix) return fail(MFSC_ERR_ARG, "index is NULL");
  info[0] = ix->k; info[1] = ix->device; info[2] = ix->ref.n_words; info[3] = ix->n_heads; info[4] = ix->n_pos;
  info[5] = (int64_t)ix->ref.n_words * 12 + (ix->n_buckets / 32 + 1) * 8 + (ix->n_heads + 2) * 4 + (ix->n_pos + 1) * 4;
  info[6] = ix->ref.n_seq; info[7] = ix->ref.n_bases;
  return MFSC_OK;
}

int mfsc_index_buffers(const mfsc_index* ix, void* ptr[5], int64_t bytes[5]) {
  if (!ix) return fail(MFSC_ERR_ARG, "index is NULL");
  ptr[0] = ix->ref.txt; bytes[0] = (int64_t)ix->ref.n_words * 8;
  ptr[1] = ix->ref.inv; bytes[1] = (int64_t)ix->ref.n_words * 4;
  ptr[2] = ix->dir; bytes[2] = (ix->n_buckets / 32 + 1) * 8;
  ptr[3] = ix->pos; bytes[3] = (ix->n_pos + 1) * 4;
  ptr[4] = ix->heads; bytes[4] = (ix->n_heads + 2) * 4;
  return MFSC_OK;
}

int mfsc_index_build_ms(const mfsc_index* ix, double ms[2]) {
  if (!ix) return fail(MFSC_ERR_ARG, "index is NULL");
  ms[0] = ix->build_ms; ms[1] = ix->tables_ms;
  return MFSC_OK;
}

int mfsc_index_free(mfsc_index* ix) {
  if (!ix) return MFSC_OK;
  index_release(g_dev, ix);
  return MFSC_OK;
}

// ------------------------------------------------------------------------------------------
// index replication over the GPUs of one box: one NCCL broadcast per index buffer over NVLink.
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch has already loaded in a torchrun rank, the
// system one otherwise), so the library itself loads on a machine without it; the calls below then fail loudly.
// ------------------------------------------------------------------------------------------
#ifndef MFSC_EMU
}  // extern "C"
#include <dlfcn.h>
namespace {
typedef struct ncclComm* nccl_comm_t;
struct nccl_uid { char internal[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(nccl_uid*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_uid, int) = nullptr;
  int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
  bool load() {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) { err = std::string("NCCL not found: ") + dlerror(); return false; }
#define MFSC_NCCL_SYM(field, sym) field = (decltype(field))dlsym(lib, sym); if (!field) { err = std::string("NCCL symbol missing: ") + sym; lib = nullptr; return false; }
    MFSC_NCCL_SYM(GetUniqueId, "ncclGetUniqueId") MFSC_NCCL_SYM(CommInitRank, "ncclCommInitRank") MFSC_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    MFSC_NCCL_SYM(CommDestroy, "ncclCommDestroy") MFSC_NCCL_SYM(Broadcast, "ncclBroadcast") MFSC_NCCL_SYM(AllGather, "ncclAllGather") MFSC_NCCL_SYM(GroupStart, "ncclGroupStart")
    MFSC_NCCL_SYM(GroupEnd, "ncclGroupEnd") MFSC_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef MFSC_NCCL_SYM
    return true;
  }
};
NcclApi g_nccl;
const int NCCL_U8 = 1;   // ncclUint8
}  // namespace
extern "C" {
#endif

struct mfsc_comm {
  int world = 0, rank = -1;          // rank >= 0: one process per GPU; rank < 0: one process, `world` devices
  std::vector<int> devices;          // local clique: device of every member
#ifndef MFSC_EMU
  std::vector<nccl_comm_t> comms;    // local clique: one per device; ranks: one
  std::vector<cudaStream_t> streams;
#endif
};

int mfsc_comm_unique_id(uint8_t id[128]) {
#ifdef MFSC_EMU
  memset(id, 0, 128); return MFSC_OK;
#else
  if (!id) return fail(MFSC_ERR_ARG, "mfsc_comm_unique_id: id is NULL");
  if (!g_nccl.load()) return fail(MFSC_ERR_CUDA, g_nccl.err);
  nccl_uid u;
  int rc = g_nccl.GetUniqueId(&u);
  if (rc) return fail(MFSC_ERR_CUDA, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc));
  memcpy(id, u.internal, 128);
  return MFSC_OK;
#endif
}

int mfsc_comm_init_rank(const uint8_t id[128], int32_t rank, int32_t world, mfsc_comm** out) {
  if (!id || !out || world < 1 || rank < 0 || rank >= world) return fail(MFSC_ERR_ARG, "mfsc_comm_init_rank: bad arguments");
  Dev& D = g_dev;
  if (!D.ok()) return fail(MFSC_ERR_CUDA, "no CUDA device (libmfsc_b200 has no CPU fallback)");
  mfsc_comm* c = new mfsc_comm();
  c->world = world; c->rank = rank; c->devices.push_back(D.device);
#ifndef MFSC_EMU
  if (!g_nccl.load()) { delete c; return fail(MFSC_ERR_CUDA, g_nccl.err); }
  nccl_uid u; memcpy(u.internal, id, 128);
  nccl_comm_t nc = nullptr;
  int rc = g_nccl.CommInitRank(&nc, world, u, rank);
  if (rc) { delete c; return fail(MFSC_ERR_CUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc)); }
  c->comms.push_back(nc); c->streams.push_back(D.stream);
#endif
  *out = c;
  return MFSC_OK;
}

int mfsc_comm_init_local(const int32_t* devices, int32_t n, mfsc_comm** out) {
  if (!devices || n < 1 || !out) return fail(MFSC_ERR_ARG, "mfsc_comm_init_local: bad arguments");
  Dev& D = g_dev;
  if (!D.ok()) return fail(MFSC_ERR_CUDA, "no CUDA device (libmfsc_b200 has no CPU fallback)");
  mfsc_comm* c = new mfsc_comm();
  c->world = n; c->rank = -1; c->devices.assign(devices, devices + n);
#ifndef MFSC_EMU
  if (!g_nccl.load()) { delete c; return fail(MFSC_ERR_CUDA, g_nccl.err); }
  c->comms.resize(n); c->streams.resize(n);
  std::vector<int> devs(devices, devices + n);
  int rc = g_nccl.CommInitAll(c->comms.data(), n, devs.data());
  if (rc) { delete c; return fail(MFSC_ERR_CUDA, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(rc)); }
  for (int i = 0; i < n; i++) {
    DeviceGuard g(devs[i]);
    if (cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); delete c; return fail(MFSC_ERR_CUDA, "cudaStreamCreate failed"); }
  }
#endif
  *out = c;
  return MFSC_OK;
}

int mfsc_comm_free(mfsc_comm* c) {
  if (!c) return MFSC_OK;
#ifndef MFSC_EMU
  for (size_t i = 0; i < c->comms.size(); i++) if (c->comms[i]) g_nccl.CommDestroy(c->comms[i]);
  if (c->rank < 0) for (size_t i = 0; i < c->streams.size(); i++) { DeviceGuard g(c->devices[i]); if (c->streams[i]) cudaStreamDestroy(c->streams[i]); }
#endif
  delete c;
  return MFSC_OK;
}

#ifndef MFSC_EMU
// an index of the root's shape on `device`, buffers from cudaMalloc, contents to be received
static mfsc_index* index_shell(int device, int k, unsigned n_words, long long n_pos, long long n_heads, const std::vector<unsigned>& h_start,
                               const std::vector<int>& h_len, const std::vector<unsigned>& h_ostart, long long n_bases) {
  DeviceGuard g(device);
  mfsc_index* ix = new mfsc_index();
  ix->device = device; ix->plain = true;
  ix->k = k; ix->n_buckets = 1ll << (2 * k); ix->n_pos = n_pos; ix->n_heads = n_heads;
  ix->ref.n_seq = (int)h_len.size(); ix->ref.n_words = n_words; ix->ref.n_bases = n_bases;
  ix->ref.h_start = h_start; ix->ref.h_len = h_len; ix->h_ostart = h_ostart;
  const size_t ns = h_len.size();
  bool ok = cudaMalloc(&ix->ref.txt, sizeof(unsigned long long) * (size_t)n_words) == cudaSuccess &&
            cudaMalloc(&ix->ref.inv, sizeof(unsigned) * (size_t)n_words) == cudaSuccess &&
            cudaMalloc(&ix->ref.start, sizeof(unsigned) * (ns + 1)) == cudaSuccess &&
            cudaMalloc(&ix->ref.len, sizeof(int) * (ns + 1)) == cudaSuccess &&
            cudaMalloc(&ix->ostart, sizeof(unsigned) * (ns + 1)) == cudaSuccess &&
            cudaMalloc(&ix->dir, sizeof(uint2) * (size_t)(ix->n_buckets / 32 + 1)) == cudaSuccess &&
            cudaMalloc(&ix->pos, sizeof(unsigned) * (size_t)(n_pos + 1)) == cudaSuccess &&
            cudaMalloc(&ix->heads, sizeof(unsigned) * (size_t)(n_heads + 2)) == cudaSuccess;
  if (ok) ok = cudaMemcpy(ix->ref.start, h_start.data(), sizeof(unsigned) * ns, cudaMemcpyHostToDevice) == cudaSuccess &&
               cudaMemcpy(ix->ref.len, h_len.data(), sizeof(int) * ns, cudaMemcpyHostToDevice) == cudaSuccess &&
               cudaMemcpy(ix->ostart, h_ostart.data(), sizeof(unsigned) * ns, cudaMemcpyHostToDevice) == cudaSuccess;
  if (!ok) { cudaGetLastError(); Dev none; none.stream = 0; index_release(none, ix); return nullptr; }
  return ix;
}
#endif

int mfsc_index_broadcast(mfsc_comm* c, int32_t root, mfsc_index** ix, double* ms) {
  if (!c || !ix || root < 0 || root >= c->world) return fail(MFSC_ERR_ARG, "mfsc_index_broadcast: bad arguments");
  if (ms) *ms = 0;
#ifdef MFSC_EMU
  // host emulation (tests only): "devices" are the one host; replicas are plain copies
  if (c->rank >= 0) return ix[0] ? MFSC_OK : fail(MFSC_ERR_ARG, "mfsc_index_broadcast: the host emulation has one rank");
  const mfsc_index* r = ix[root];
  if (!r) return fail(MFSC_ERR_ARG, "mfsc_index_broadcast: the root index is NULL");
  for (int i = 0; i < c->world; i++) {
    if (i == root) continue;
    mfsc_index* n = new mfsc_index(*r);
    auto dup = [](const void* src, size_t bytes) { void* d = calloc(bytes ? bytes : 1, 1); memcpy(d, src, bytes); return d; };
    const size_t ns = (size_t)r->ref.n_seq;
    n->ref.txt = (unsigned long long*)dup(r->ref.txt, 8 * (size_t)r->ref.n_words); n->ref.inv = (unsigned*)dup(r->ref.inv, 4 * (size_t)r->ref.n_words);
    n->ref.start = (unsigned*)dup(r->ref.start, 4 * ns); n->ref.len = (int*)dup(r->ref.len, 4 * ns); n->ostart = (unsigned*)dup(r->ostart, 4 * ns);
    n->dir = (uint2*)dup(r->dir, 8 * (size_t)(r->n_buckets / 32 + 1)); n->pos = (unsigned*)dup(r->pos, 4 * (size_t)(r->n_pos + 1));
    n->heads = (unsigned*)dup(r->heads, 4 * (size_t)(r->n_heads + 2));
    ix[i] = n;
  }
  return MFSC_OK;
#else
  Dev& D = g_dev;
  if (!D.ok()) return fail(MFSC_ERR_CUDA, "no CUDA device (libmfsc_b200 has no CPU fallback)");
  const auto t0 = std::chrono::steady_clock::now();
  auto bcast = [&](std::vector<void*>& send_recv, size_t bytes) -> int {
    // one ncclBroadcast per member of this process, grouped
    int rc = g_nccl.GroupStart();
    for (size_t i = 0; i < c->comms.size() && !rc; i++) {
      const int member = c->rank >= 0 ? c->rank : (int)i;
      (void)member;
      rc = g_nccl.Broadcast(send_recv[i], send_recv[i], bytes, NCCL_U8, root, c->comms[i], c->streams[i]);
    }
    int rc2 = g_nccl.GroupEnd();
    return rc ? rc : rc2;
  };
  auto sync_all = [&]() {
    for (size_t i = 0; i < c->streams.size(); i++) { DeviceGuard g(c->devices[i]); cudaStreamSynchronize(c->streams[i]); }
  };
  if (c->rank < 0) {
    // ---- one process: ix[root] is built, ix[i] is created on devices[i] ----
    mfsc_index* r = ix[root];
    if (!r || r->device != c->devices[root]) return fail(MFSC_ERR_ARG, "mfsc_index_broadcast: the root index does not live on devices[root]");
    for (int i = 0; i < c->world; i++) {
      if (i == root) continue;
      ix[i] = index_shell(c->devices[i], r->k, r->ref.n_words, r->n_pos, r->n_heads, r->ref.h_start, r->ref.h_len, r->h_ostart, r->ref.n_bases);
      if (!ix[i]) { for (int j = 0; j < i; j++) if (j != root) { mfsc_index_free(ix[j]); ix[j] = nullptr; } return fail(MFSC_ERR_NOMEM, "out of device memory (index replica)"); }
    }
    void* rp[5]; int64_t rb[5];
    mfsc_index_buffers(r, rp, rb);
    for (int b = 0; b < 5; b++) {
      std::vector<void*> bufs(c->world);
      for (int i = 0; i < c->world; i++) { void* p[5]; int64_t nb[5]; mfsc_index_buffers(ix[i], p, nb); bufs[i] = p[b]; }
      int rc = bcast(bufs, (size_t)rb[b]);
      if (rc) return fail(MFSC_ERR_CUDA, std::string("ncclBroadcast: ") + g_nccl.GetErrorString(rc));
    }
    sync_all();
  } else {
    // ---- one process per GPU: shape first, then the record tables, then the five buffers ----
    long long meta[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    mfsc_index* r = ix[0];
    if (c->rank == root) {
      if (!r) return fail(MFSC_ERR_ARG, "mfsc_index_broadcast: the root rank has no index");
      meta[0] = r->k; meta[1] = r->ref.n_words; meta[2] = r->n_pos; meta[3] = r->ref.n_seq; meta[4] = r->ref.n_bases; meta[5] = r->n_heads;
    }
    long long* d_meta = dalloc<long long>(D, 8);
    if (!d_meta) return fail(MFSC_ERR_NOMEM, "out of device memory (broadcast)");
    D.h2d(d_meta, meta, sizeof(meta)); D.sync();
    std::vector<void*> one(1);
    one[0] = d_meta;
    int rc = bcast(one, sizeof(meta));
    if (rc) { D.free_(d_meta); return fail(MFSC_ERR_CUDA, std::string("ncclBroadcast: ") + g_nccl.GetErrorString(rc)); }
    D.d2h(meta, d_meta, sizeof(meta));
    D.free_(d_meta);
    const size_t ns = (size_t)meta[3];
    std::vector<unsigned> tab(3 * ns + 1);
    if (c->rank == root) {
      memcpy(tab.data(), r->ref.h_start.data(), 4 * ns); memcpy(tab.data() + ns, r->ref.h_len.data(), 4 * ns);
      memcpy(tab.data() + 2 * ns, r->h_ostart.data(), 4 * ns);
    }
    unsigned* d_tab = dalloc<unsigned>(D, 3 * ns + 1);
    if (!d_tab) return fail(MFSC_ERR_NOMEM, "out of device memory (broadcast)");
    D.h2d(d_tab, tab.data(), 4 * 3 * ns); D.sync();
    one[0] = d_tab;
    rc = bcast(one, 4 * 3 * ns);
    if (rc) { D.free_(d_tab); return fail(MFSC_ERR_CUDA, std::string("ncclBroadcast: ") + g_nccl.GetErrorString(rc)); }
    D.d2h(tab.data(), d_tab, 4 * 3 * ns);
    D.free_(d_tab);
    if (c->rank != root) {
      std::vector<unsigned> hs(tab.begin(), tab.begin() + ns), ho(tab.begin() + 2 * ns, tab.begin() + 3 * ns);
      std::vector<int> hl(ns);
      memcpy(hl.data(), tab.data() + ns, 4 * ns);
      r = index_shell(D.device, (int)meta[0], (unsigned)meta[1], meta[2], meta[5], hs, hl, ho, meta[4]);
      if (!r) return fail(MFSC_ERR_NOMEM, "out of device memory (index replica)");
      ix[0] = r;
    }
    void* p[5]; int64_t nb[5];
    mfsc_index_buffers(r, p, nb);
    for (int b = 0; b < 5; b++) {
      one[0] = p[b];
      rc = bcast(one, (size_t)nb[b]);
      if (rc) return fail(MFSC_ERR_CUDA, std::string("ncclBroadcast: ") + g_nccl.GetErrorString(rc));
    }
    D.sync();
  }
  if (ms) *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  std::string msg;
  if (!dev_check(msg)) return fail(MFSC_ERR_CUDA, "index broadcast: " + msg);
  return MFSC_OK;
#endif
}

// Sharded build (one process per GPU): every rank packs the contigs, builds the tables of ITS slice of the k-mer space
// (count / fill touch only k-mers of the slice, the dense scratch histogram is 4^k / N entries), and the slices are
// exchanged with grouped NCCL broadcasts, one root per slice (an all-gather of unequal pieces) -- the build time divides by
// N and what crosses NVLink is the compact index once.
int mfsc_index_build_sharded(mfsc_comm* c, const uint8_t* bases, const int64_t* rec_off, int32_t n_rec, const mfsc_params* p, mfsc_index** out,
                             double ms[4]) {
  if (!c || !bases || !rec_off || n_rec <= 0 || !out) return fail(MFSC_ERR_ARG, "mfsc_index_build_sharded: bad arguments");
  if (ms) ms[0] = ms[1] = ms[2] = ms[3] = 0;
  if (c->rank < 0 || c->world == 1) {
    if (c->rank < 0 && c->world > 1) return fail(MFSC_ERR_ARG, "mfsc_index_build_sharded: needs a one-process-per-GPU communicator (mfsc_comm_init_rank)");
    return mfsc_index_build(bases, rec_off, n_rec, p, out);
  }
#ifdef MFSC_EMU
  return mfsc_index_build(bases, rec_off, n_rec, p, out);
#else
  int rc = check_params(p); if (rc) return rc;
  Dev& D = g_dev;
  if (!D.ok()) return fail(MFSC_ERR_CUDA, "no CUDA device (libmfsc_b200 has no CPU fallback)");
  const auto t0 = std::chrono::steady_clock::now();
  auto since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
  mfsc_index* ix = new mfsc_index();
  ix->device = D.device;
  std::string err;
  const int N = c->world, me = c->rank;
  // the contigs cross PCIe once: rank 0 uploads and packs them, the packed text (0.375 B per base) goes to the other ranks over
  // NVLink -- eight ranks uploading the same 50 MB next to their read batches compete for the host's memory bandwidth
  if (!seq_upload(D, bases, rec_off, n_rec, 0, me == 0, ix->ref, err)) { index_release(D, ix); return fail(MFSC_ERR_NOMEM, err); }
  {
    int nrc0 = g_nccl.GroupStart();
    if (!nrc0) nrc0 = g_nccl.Broadcast(ix->ref.txt, ix->ref.txt, sizeof(unsigned long long) * (size_t)ix->ref.n_words, NCCL_U8, 0, c->comms[0], c->streams[0]);
    if (!nrc0) nrc0 = g_nccl.Broadcast(ix->ref.inv, ix->ref.inv, sizeof(unsigned) * (size_t)ix->ref.n_words, NCCL_U8, 0, c->comms[0], c->streams[0]);
    { int e2 = g_nccl.GroupEnd(); if (!nrc0) nrc0 = e2; }
    if (nrc0) { index_release(D, ix); return fail(MFSC_ERR_CUDA, std::string("ncclBroadcast: ") + g_nccl.GetErrorString(nrc0)); }
  }
  ix->h_ostart.resize(n_rec);
  for (int r = 0; r < n_rec; r++) ix->h_ostart[r] = (unsigned)(rec_off[r] - rec_off[0] + r);
  ix->ostart = dalloc<unsigned>(D, n_rec);
  if (!ix->ostart) { index_release(D, ix); return fail(MFSC_ERR_NOMEM, "out of device memory (record table)"); }
  D.h2d(ix->ostart, ix->h_ostart.data(), sizeof(unsigned) * n_rec);
  ix->k = choose_k(p, ix->ref.n_bases);
  ix->n_buckets = 1ll << (2 * ix->k);
  const long long words = ix->n_buckets / 32;
  auto word_lo = [&](int r) { return words * r / N; };
  if (ms) ms[0] = since(t0);
  const auto t1 = std::chrono::steady_clock::now();
  TableSlice S;
  if (!build_slice(D, ix->ref, ix->k, (unsigned long long)word_lo(me) * 32ull, (unsigned long long)word_lo(me + 1) * 32ull, S, err)) {
    index_release(D, ix); return fail(MFSC_ERR_NOMEM, err);
  }
  if (ms) ms[1] = since(t1);
  const auto t2 = std::chrono::steady_clock::now();
  // sizes of all slices
  std::vector<long long> sizes(2 * N, 0);
  sizes[2 * me] = S.n_occ; sizes[2 * me + 1] = S.n_pos;
  long long* d_sizes = dalloc<long long>(D, 2 * N);
  if (!d_sizes) { slice_free(D, S); index_release(D, ix); return fail(MFSC_ERR_NOMEM, "out of device memory (sizes)"); }
  D.h2d(d_sizes, sizes.data(), sizeof(long long) * 2 * N);
  int nrc = g_nccl.GroupStart();
  for (int r = 0; r < N && !nrc; r++) nrc = g_nccl.Broadcast(d_sizes + 2 * r, d_sizes + 2 * r, 16, NCCL_U8, r, c->comms[0], c->streams[0]);
  { int e2 = g_nccl.GroupEnd(); if (!nrc) nrc = e2; }
  if (nrc) { D.free_(d_sizes); slice_free(D, S); index_release(D, ix); return fail(MFSC_ERR_CUDA, std::string("ncclBroadcast: ") + g_nccl.GetErrorString(nrc)); }
  D.d2h(sizes.data(), d_sizes, sizeof(long long) * 2 * N);
  D.free_(d_sizes);
  std::vector<long long> occ_off(N + 1, 0), pos_off(N + 1, 0);
  for (int r = 0; r < N; r++) { occ_off[r + 1] = occ_off[r] + sizes[2 * r]; pos_off[r + 1] = pos_off[r] + sizes[2 * r + 1]; }
  ix->n_heads = occ_off[N]; ix->n_pos = pos_off[N];
  if ((unsigned long long)ix->n_pos >= 0xffffffffull) { slice_free(D, S); index_release(D, ix); return fail(MFSC_ERR_ARG, "reference too large"); }
  ix->dir = dalloc<uint2>(D, words + 1); ix->heads = dalloc<unsigned>(D, ix->n_heads + 2); ix->pos = dalloc<unsigned>(D, ix->n_pos + 1);
  if (!ix->dir || !ix->heads || !ix->pos) { slice_free(D, S); index_release(D, ix); return fail(MFSC_ERR_NOMEM, "out of device memory (index)"); }
  // own slice into place, ranks and position offsets made global
  if (S.n_words) {
    MFSC_LAUNCH(k_dir_place, grid_for(D, S.n_words, 256, 16), 256, 0, D.stream, S.dir, S.n_words, (unsigned)occ_off[me], ix->dir + word_lo(me));
    D.launches++;
  }
  if (S.n_occ) {
    MFSC_LAUNCH(k_copy_add_u32, grid_for(D, S.n_occ, 256, 16), 256, 0, D.stream, S.heads, (long long)S.n_occ, (unsigned)pos_off[me], ix->heads + occ_off[me]);
    D.launches++;
  }
  if (S.n_pos) cudaMemcpyAsync(ix->pos + pos_off[me], S.pos, sizeof(unsigned) * (size_t)S.n_pos, cudaMemcpyDeviceToDevice, D.stream);
  {
    const unsigned np = (unsigned)ix->n_pos;
    D.h2d(ix->heads + ix->n_heads, &np, sizeof(unsigned));
  }
  // exchange: slice r comes from rank r.  The directory slices have equal size: one in-place ncclAllGather.  Heads and positions
  // are (nearly) equal pieces of different size: each is gathered with one ncclAllGather into a staging buffer of N x the
  // largest piece and copied into place.
  {
    const long long wpr = words / N;
    if (words % N == 0 && wpr > 0) {
      nrc = g_nccl.AllGather(ix->dir + word_lo(me), ix->dir, sizeof(uint2) * (size_t)wpr, NCCL_U8, c->comms[0], c->streams[0]);
    } else {
      nrc = g_nccl.GroupStart();
      for (int r = 0; r < N && !nrc; r++) {
        const size_t bytes = sizeof(uint2) * (size_t)(word_lo(r + 1) - word_lo(r));
        if (bytes) nrc = g_nccl.Broadcast(ix->dir + word_lo(r), ix->dir + word_lo(r), bytes, NCCL_U8, r, c->comms[0], c->streams[0]);
      }
      int e2 = g_nccl.GroupEnd(); if (!nrc) nrc = e2;
    }
    long long max_occ = 0, max_pos = 0;
    for (int r = 0; r < N; r++) { max_occ = std::max(max_occ, sizes[2 * r]); max_pos = std::max(max_pos, sizes[2 * r + 1]); }
    unsigned* st_h = dalloc<unsigned>(D, (size_t)max_occ * N + 1);
    unsigned* st_p = dalloc<unsigned>(D, (size_t)max_pos * N + 1);
    if (!st_h || !st_p) { D.free_(st_h); D.free_(st_p); slice_free(D, S); index_release(D, ix); return fail(MFSC_ERR_NOMEM, "out of device memory (exchange)"); }
    if (!nrc && max_occ) {
      cudaMemcpyAsync(st_h + (size_t)max_occ * me, ix->heads + occ_off[me], sizeof(unsigned) * (size_t)sizes[2 * me], cudaMemcpyDeviceToDevice, D.stream);
      nrc = g_nccl.AllGather(st_h + (size_t)max_occ * me, st_h, sizeof(unsigned) * (size_t)max_occ, NCCL_U8, c->comms[0], c->streams[0]);
      for (int r = 0; r < N; r++)
        if (r != me && sizes[2 * r]) cudaMemcpyAsync(ix->heads + occ_off[r], st_h + (size_t)max_occ * r, sizeof(unsigned) * (size_t)sizes[2 * r], cudaMemcpyDeviceToDevice, D.stream);
    }
    if (!nrc && max_pos) {
      cudaMemcpyAsync(st_p + (size_t)max_pos * me, ix->pos + pos_off[me], sizeof(unsigned) * (size_t)sizes[2 * me + 1], cudaMemcpyDeviceToDevice, D.stream);
      nrc = g_nccl.AllGather(st_p + (size_t)max_pos * me, st_p, sizeof(unsigned) * (size_t)max_pos, NCCL_U8, c->comms[0], c->streams[0]);
      for (int r = 0; r < N; r++)
        if (r != me && sizes[2 * r + 1]) cudaMemcpyAsync(ix->pos + pos_off[r], st_p + (size_t)max_pos * r, sizeof(unsigned) * (size_t)sizes[2 * r + 1], cudaMemcpyDeviceToDevice, D.stream);
    }
    {
      const unsigned np = (unsigned)ix->n_pos;       // the terminal entry again: rank N-1's piece may have been copied over it
      D.h2d(ix->heads + ix->n_heads, &np, sizeof(unsigned));
    }
    D.sync();
    D.free_(st_h); D.free_(st_p);
  }
  D.sync();
  slice_free(D, S);
  if (nrc) { index_release(D, ix); return fail(MFSC_ERR_CUDA, std::string("ncclBroadcast: ") + g_nccl.GetErrorString(nrc)); }
  if (ms) { ms[2] = since(t2); ms[3] = since(t0); }
  ix->build_ms = since(t0); ix->tables_ms = ms ? ms[1] : 0;
  std::string msg;
  if (!dev_check(msg)) { index_release(D, ix); return fail(MFSC_ERR_CUDA, "sharded index build: " + msg); }
  *out = ix;
  return MFSC_OK;
#endif
}

static void reads_release(Dev& D, mfsc_reads* r) {
  if (!r) return;
  if (!D.stream || (r->device >= 0 && r->device != D.device)) {
    DeviceGuard g(r->device);
    dev_free_plain(r->q.txt); dev_free_plain(r->q.inv); dev_free_plain(r->q.start); dev_free_plain(r->q.len); dev_free_plain(r->tile_off);
  } else {
    seq_free(D, r->q);
    D.free_(r->tile_off);
  }
  delete r;
}

int mfsc_reads_upload(const uint8_t* bases, const int64_t* read_off, int32_t n_reads, mfsc_reads** out) {
  if (!bases || !read_off || n_reads <= 0 || !out) return fail(MFSC_ERR_ARG, "mfsc_reads_upload: bad arguments");
  Dev& D = g_dev;
  if (!D.ok()) return fail(MFSC_ERR_CUDA, "no CUDA device (libmfsc_b200 has no CPU fallback)");
  mfsc_reads* r = new mfsc_reads();
  r->device = D.device;
  std::string err;
  Stamp t0, t1;
  stamp(D, t0);
  if (!seq_upload(D, bases, read_off, n_reads, 1, true, r->q, err)) { reads_release(D, r); return fail(MFSC_ERR_NOMEM, err); }
  // seed tiles: a strand of len bases is probed in tiles of SEED_TILE_BASES bases (mfsc_seed.cuh)
  {
    std::vector<unsigned> toff(r->q.n_seq + 1);
    unsigned long long acc = 0;
    for (int k = 0; k < r->q.n_seq; k++) { toff[k] = (unsigned)acc; acc += (unsigned long long)((r->q.h_len[k] + SEED_TILE_BASES - 1) / SEED_TILE_BASES); }
    toff[r->q.n_seq] = (unsigned)acc;
    r->n_tiles = (unsigned)acc;
    r->tile_off = dalloc<unsigned>(D, r->q.n_seq + 1);
    if (!r->tile_off) { reads_release(D, r); return fail(MFSC_ERR_NOMEM, "out of device memory (seed tiles)"); }
    D.h2d(r->tile_off, toff.data(), sizeof(unsigned) * (r->q.n_seq + 1));
    D.sync();
  }
  stamp(D, t1);
  D.sync();
  r->upload_ms = elapsed_ms(D, t0, t1);
  stamp_free(t0); stamp_free(t1);
  r->n_reads = n_reads;
  r->h2d_bytes = r->q.n_bases + (long long)sizeof(long long) * (n_reads + 1) + 12ll * r->q.n_seq + 4ll * (r->q.n_seq + 1);
  std::string msg;
  if (!dev_check(msg)) { reads_release(D, r); return fail(MFSC_ERR_CUDA, "reads upload: " + msg); }
  *out = r;
  return MFSC_OK;
}

int mfsc_reads_free(mfsc_reads* r) {
  reads_release(g_dev, r);
  return MFSC_OK;
}

int mfsc_align_reads(mfsc_index* ix, mfsc_reads* rd, const mfsc_params* p, int32_t flags, mfsc_result** out) {
  if (!ix || !rd || !out) return fail(MFSC_ERR_ARG, "mfsc_align_reads: bad arguments");
  int rc = check_params(p); if (rc) return rc;
  Dev& D = g_dev;
  if (!D.ok()) return fail(MFSC_ERR_CUDA, "no CUDA device (libmfsc_b200 has no CPU fallback)");
  if (ix->k > p->minmatch) return fail(MFSC_ERR_ARG, "index k-mer is longer than minmatch; rebuild the index with these params");
  const long long launches0 = D.launches;
  mfsc_result* res = new mfsc_result();
  memset(res->stats, 0, sizeof(res->stats)); memset(res->ms, 0, sizeof(res->ms));
  std::vector<void*> garbage;
  auto keep = [&](void* ptr) { garbage.push_back(ptr); return ptr; };
  auto cleanup = [&]() { for (void* g : garbage) D.free_(g); garbage.clear(); };
#define DA(T, name, count) T* name = (T*)keep(D.alloc(sizeof(T) * (size_t)((count) + 1))); if (!name) { cleanup(); delete res; return fail(MFSC_ERR_NOMEM, "out of device memory (" #name ")"); }

  DevParams P;
  P.minmatch = p->minmatch; P.mincluster = p->mincluster; P.maxgap = p->maxgap; P.breaklen = p->breaklen; P.diagdiff = p->diagdiff;
  P.maxmatch = p->maxmatch; P.simplify = p->simplify; P.band_cap = p->band_cap; P.diagfactor = p->diagfactor;
  P.k = ix->k; P.stride = p->minmatch - ix->k + 1; P.one = 1;
  P.pk_u_span = p->pk_u_span > 0 && p->pk_u_span < PK_U_SPAN ? p->pk_u_span : PK_U_SPAN;
  P.pk_clean = PK_CLEAN; P.pk_prio = PK_P; P.pk_score = PK_SCORE;
  P.pk_m_span = p->pk_m_span > 0 && p->pk_m_span < PK_M_SPAN ? p->pk_m_span : PK_M_SPAN;
  const SeqStore RS = ix->ref.view(), QS = rd->q.view();
  const int n_reads = rd->n_reads, n_qs = 2 * n_reads;
  Stamp ts[8];
  stamp(D, ts[0]);

  // ---- stage 2: probes -> maximal exact matches ----
  long long n_probes = 0;
  for (int qs = 0; qs < n_qs; qs++) {
    int len = rd->q.h_len[qs];
    n_probes += len >= P.k ? (len - P.k) / P.stride + 1 : 0;
  }
  DA(unsigned, d_counter, 8);
  DA(unsigned long long, d_found, 1);
  long long cap = std::max<long long>(1 << 16, rd->q.n_bases / 16);
  MemOut mo; unsigned long long found = 0;
  int *m_qs = nullptr, *m_s2 = nullptr, *m_len = nullptr, *m_rec = nullptr; unsigned* m_s1 = nullptr;
  SeedTiles ST; ST.tile_off = rd->tile_off; ST.n_tiles = rd->n_tiles; ST.ticket = d_counter + 4;
  for (int attempt = 0; attempt < 3; attempt++) {
    m_qs = (int*)D.alloc(sizeof(int) * (cap + 1)); m_s1 = (unsigned*)D.alloc(sizeof(unsigned) * (cap + 1));
    m_s2 = (int*)D.alloc(sizeof(int) * (cap + 1)); m_len = (int*)D.alloc(sizeof(int) * (cap + 1)); m_rec = (int*)D.alloc(sizeof(int) * (cap + 1));
    if (!m_qs || !m_s1 || !m_s2 || !m_len || !m_rec) {
      D.free_(m_qs); D.free_(m_s1); D.free_(m_s2); D.free_(m_len); D.free_(m_rec);
      cleanup(); delete res; return fail(MFSC_ERR_NOMEM, "out of device memory (matches)");
    }
    D.zero(d_counter, sizeof(unsigned) * 8);
    D.zero(d_found, sizeof(unsigned long long));
    mo.qs = m_qs; mo.s1 = m_s1; mo.s2 = m_s2; mo.len = m_len; mo.rec = m_rec; mo.counter = d_found; mo.capacity = (unsigned long long)cap;
    if (rd->n_tiles > 0) {
      // persistent grid: warps take tiles of the query strands from a ticket counter
      MFSC_LAUNCH(k_seeds, grid_for(D, ((long long)rd->n_tiles + 3) / 4 * MFSC_LANES, MFSC_LANES * SEED_WARPS, SEED_BLOCKS_PER_SM), MFSC_LANES * SEED_WARPS, 0, D.stream,
                  RS, ix->ostart, ix->view(), QS, ST, P, mo);
      D.launches++;
    }
    D.d2h(&found, d_found, sizeof(unsigned long long));
    if (found >= 0x7fffffffull) {
      D.free_(m_qs); D.free_(m_s1); D.free_(m_s2); D.free_(m_len); D.free_(m_rec);
      cleanup(); delete res; return fail(MFSC_ERR_ARG, "more than 2^31 matches in one batch: split the reads into smaller batches");
    }
    if ((long long)found <= cap) break;
    D.free_(m_qs); D.free_(m_s1); D.free_(m_s2); D.free_(m_len); D.free_(m_rec);
    m_qs = m_s2 = m_len = m_rec = nullptr; m_s1 = nullptr;
    cap = (long long)found + 1024;
  }
  if (!m_qs) { cleanup(); delete res; return fail(MFSC_ERR_NOMEM, "match buffer: the retry did not converge"); }
  keep(m_qs); keep(m_s1); keep(m_s2); keep(m_len); keep(m_rec);
  if (found >= 0x7fffffffull) { cleanup(); delete res; return fail(MFSC_ERR_ARG, "more than 2^31 matches in one batch: split the reads into smaller batches"); }
  const long long M = found;
  res->n_mems = M;
  res->stats[3] = n_probes; res->stats[4] = M;
  stamp(D, ts[1]);

  // ---- order by (read strand, s2, s1): two stable radix passes over a permutation ----
  DA(unsigned, perm, M); DA(unsigned, perm_alt, M); DA(unsigned, key1, M); DA(unsigned, key1_alt, M);
  DA(unsigned long long, key2, M); DA(unsigned long long, key2_alt, M);
  DA(int, s_qs, M); DA(unsigned, s_s1, M); DA(int, s_s2, M); DA(int, s_len, M); DA(int, s_rec, M);
  DA(int, seg, n_qs + 1);
  if (M > 0) {
    const unsigned g = grid_for(D, M, 256, 16);
    MFSC_LAUNCH(k_iota, g, 256, 0, D.stream, perm, M);
    MFSC_LAUNCH(k_sort_key1, g, 256, 0, D.stream, m_s1, M, key1);
    unsigned *k1o = nullptr, *p1o = nullptr;
    long long max_pos = (long long)ix->h_ostart.back() + ix->ref.h_len.back() + 1;
    if (!sort_pairs_u32(D, key1, perm, key1_alt, perm_alt, M, bits_for((unsigned long long)max_pos), &k1o, &p1o)) { cleanup(); delete res; return fail(MFSC_ERR_NOMEM, "out of device memory (sort)"); }
    unsigned* p1_other = (p1o == perm) ? perm_alt : perm;
    MFSC_LAUNCH(k_sort_key, g, 256, 0, D.stream, m_qs, m_s2, p1o, M, key2);
    unsigned* p2o = nullptr;
    if (!sort_pairs_u64(D, key2, p1o, key2_alt, p1_other, M, 32 + bits_for((unsigned long long)n_qs), &p2o)) { cleanup(); delete res; return fail(MFSC_ERR_NOMEM, "out of device memory (sort)"); }
    MFSC_LAUNCH(k_gather_mems, g, 256, 0, D.stream, p2o, M, m_qs, m_s1, m_s2, m_len, m_rec, s_qs, s_s1, s_s2, s_len, s_rec);
    D.launches += 4;
  }
  MFSC_LAUNCH(k_seg_bounds, grid_for(D, M + 1, 256, 16), 256, 0, D.stream, s_qs, M, n_qs, seg);
  D.launches++;
  stamp(D, ts[2]);

  if (flags & MFSC_KEEP_MEMS) {
    std::vector<unsigned> hs1(M);
    res->m_q.resize(M); res->m_strand.resize(M); res->m_s2.resize(M); res->m_len.resize(M); res->m_s1.resize(M);
    D.d2h(res->m_q.data(), s_qs, sizeof(int) * M); D.d2h(hs1.data(), s_s1, sizeof(unsigned) * M);
    D.d2h(res->m_s2.data(), s_s2, sizeof(int) * M); D.d2h(res->m_len.data(), s_len, sizeof(int) * M);
    for (long long i = 0; i < M; i++) { res->m_strand[i] = res->m_q[i] & 1; res->m_q[i] >>= 1; res->m_s1[i] = hs1[i]; }
  }

  long long n_rows = 0;
  if (!(flags & MFSC_STOP_SEEDS)) {
    // ---- stage 3: clusters ----
    ClusterScratch W; ClusterOut O;
    DA(unsigned, w_s1, M); DA(int, w_s2, M); DA(int, w_len, M); DA(int, w_rec, M); DA(int, w_score, M); DA(int, w_adj, M);
    DA(int, w_from, M); DA(int, w_uf, M); DA(int, w_ord, M); DA(int, w_tmp, M); DA(unsigned char, w_flag, M);
    DA(unsigned, o_s1, M); DA(int, o_s2, M); DA(int, o_len, M); DA(int, o_pcf, M); DA(int, o_pcn, M); DA(int, o_pcr, M);
    DA(int, o_npc, n_qs); DA(int, o_ncm, n_qs);
    W.s1 = w_s1; W.s2 = w_s2; W.len = w_len; W.rec = w_rec; W.score = w_score; W.adj = w_adj; W.from = w_from; W.uf = w_uf;
    W.ord = w_ord; W.tmp = w_tmp; W.flag = w_flag;
    O.s1 = o_s1; O.s2 = o_s2; O.len = o_len; O.pc_first = o_pcf; O.pc_n = o_pcn; O.pc_rec = o_pcr; O.n_pc = o_npc; O.n_cm = o_ncm;
    {
      // strands per ticket: 8 when there are many more strands than resident warps, fewer on small batches so that every warp slot has work
      const long long resident = (long long)D.n_sm() * CL_MINBLOCKS * 4;
      long long per_ticket = n_qs / (resident * 4);
      per_ticket = per_ticket < 1 ? 1 : (per_ticket > 8 ? 8 : per_ticket);
      MFSC_LAUNCH(k_cluster, grid_for(D, ((long long)n_qs + per_ticket - 1) / per_ticket * MFSC_LANES, 128, 16), 128, 0, D.stream, s_s1, s_s2, s_len, s_rec, seg,
                  n_qs, P, W, O, d_counter + 5, (unsigned)per_ticket);
    }
    D.launches++;
    stamp(D, ts[3]);

    if (flags & MFSC_KEEP_CLUSTERS) {
      std::vector<int> hseg(n_qs + 1), hnpc(n_qs), hpcf(M), hpcn(M), hpcr(M), hs2(M), hlen(M);
      std::vector<unsigned> hs1(M);
      D.d2h(hseg.data(), seg, sizeof(int) * (n_qs + 1)); D.d2h(hnpc.data(), o_npc, sizeof(int) * n_qs);
      D.d2h(hpcf.data(), o_pcf, sizeof(int) * M); D.d2h(hpcn.data(), o_pcn, sizeof(int) * M); D.d2h(hpcr.data(), o_pcr, sizeof(int) * M);
      D.d2h(hs1.data(), o_s1, sizeof(unsigned) * M); D.d2h(hs2.data(), o_s2, sizeof(int) * M); D.d2h(hlen.data(), o_len, sizeof(int) * M);
      for (int qs = 0; qs < n_qs; qs++)
        for (int c = 0; c < hnpc[qs]; c++) {
          int pc = hseg[qs] + c;
          res->c_q.push_back(qs >> 1); res->c_strand.push_back(qs & 1); res->c_rec.push_back(hpcr[pc]);
          res->c_first.push_back((int)res->cm_s1.size()); res->c_n.push_back(hpcn[pc]);
          for (int t = 0; t < hpcn[pc]; t++) { int w = hpcf[pc] + t; res->cm_s1.push_back(hs1[w]); res->cm_s2.push_back(hs2[w]); res->cm_len.push_back(hlen[w]); }
        }
    }

    if (!(flags & MFSC_STOP_CLUSTERS)) {
      // ---- stage 4: extension ----
      RowsOut R0, R1;
      const bool want_deltas = (flags & MFSC_WANT_DELTAS) != 0;
      memset(&R0, 0, sizeof(R0)); memset(&R1, 0, sizeof(R1));
      DA(int, r_ref, M); DA(int, r_qry, M); DA(int, r_dir, M); DA(int, r_sA, M); DA(int, r_eA, M); DA(int, r_sB, M); DA(int, r_eB, M);
      DA(int, r_errs, M); DA(int, r_cols, M); DA(int, r_n, n_reads); DA(int, r_off, n_reads + 1);
      DA(int, x_ord, M); DA(int, x_tmp, M); DA(int, x_grp, M); DA(unsigned char, x_fused, M);
      DA(unsigned long long, d_stats, 8);
      D.zero(d_stats, sizeof(unsigned long long) * 8);
      R0.ref = r_ref; R0.qry = r_qry; R0.dir = r_dir; R0.sA = r_sA; R0.eA = r_eA; R0.sB = r_sB; R0.eB = r_eB; R0.errs = r_errs; R0.cols = r_cols; R0.n_rows = r_n;
      DA(int, x_retry, n_reads);
      ExtIn I;
      I.seg = seg; I.n_pc = o_npc; I.pc_first = o_pcf; I.pc_n = o_pcn; I.pc_rec = o_pcr; I.cm_s1 = o_s1; I.cm_s2 = o_s2; I.cm_len = o_len;
      I.r_ostart = ix->ostart; I.ord = x_ord; I.key_tmp = x_tmp; I.grp = x_grp; I.fused = x_fused;
      I.list = nullptr; I.n_list = nullptr; I.retry = x_retry; I.n_retry = (unsigned*)(d_stats + 4);
      I.gap_res = nullptr; I.ord_ready = 0;
      ExtStats st; st.cells = d_stats; st.calls = d_stats + 1; st.max_band = (int*)(d_stats + 2); st.caps = d_stats + 5; st.gaps = d_stats + 7;
      const int ext_block = 128;
      const size_t ext_smem = (size_t)(ext_block / MFSC_LANES) * DP_SMEM_INTS * sizeof(int);
      const size_t pk_smem = (size_t)(ext_block / MFSC_LANES) * PK_SMEM_INTS * sizeof(int);
      // persistent grids: as many blocks as are resident at once, reads handed out by ticket
      unsigned* ext_ticket = (unsigned*)(d_stats + 3);
      auto resident_grid = [&](const void* fn, size_t smem, int fallback) {
        unsigned g = grid_for(D, (long long)n_reads * MFSC_LANES, ext_block, fallback);
#ifndef MFSC_EMU
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, ext_block, smem) == cudaSuccess && occ > 0)
          g = grid_for(D, (long long)n_reads * MFSC_LANES, ext_block, occ);
#else
        (void)fn; (void)smem;
#endif
        return g;
      };
      TbPool pool; pool.tb = nullptr; pool.rev = nullptr; pool.arena = nullptr; pool.arena_ints = 0;
      DeltaOut DO; DO.buf = nullptr; DO.used = nullptr; DO.cap = 0; DO.n_bad = nullptr;
      if (!want_deltas) {
        const unsigned ext_grid = resident_grid((const void*)k_extend<false, false>, ext_smem, 4);
        if (p->engine == 1) {
          MFSC_LAUNCH((k_extend<false, false>), ext_grid, ext_block, ext_smem, D.stream, RS, QS, I, n_reads, P, R0, st, ext_ticket, pool, DO);
        } else {
          // stage 4a (mfsc_gaps.cuh): the forward extensions that do not depend on the walk, as independent tasks ahead of it.
          // It pays where the work per read is large and uneven -- reads against reads, repeats: thousands of matches per read,
          // measured -12 % on stage 4 -- and costs a few per cent where a read is one or two alignments (the walk kernel that
          // can pick results up is the slower instantiation), so the match count per read decides; engine 2 / 3 force it on / off.
          static const bool gap_env = getenv("MFSC_NO_GAP_TASKS") == nullptr;
          const bool gap_tasks = gap_env && p->engine != 3 && (p->engine == 2 || M >= (long long)MFSC_GAP_MIN_MATCHES_PER_READ * n_reads);
          const unsigned pk_grid = resident_grid((const void*)k_extend<false, true>, pk_smem, 7);
          if (gap_tasks) {
            DA(int4, g_res, M); DA(int4, g_task, M); DA(int, g_rec, M);
            GapPlan G; G.res = g_res; G.task = g_task; G.task_rec = g_rec; G.n_task = (unsigned*)(d_stats + 6);
            MFSC_LAUNCH(k_gap_plan, grid_for(D, (long long)n_reads * MFSC_LANES, 128, 16), 128, 0, D.stream, RS, QS, I, n_reads, P, G);
            // (the task count is only known on the device: as many blocks as are resident, whatever the number of reads)
            unsigned gap_grid = resident_grid((const void*)k_gaps<true>, pk_smem, 7);
#ifndef MFSC_EMU
            {
              int occ = 0;
              if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k_gaps<true>, ext_block, pk_smem) == cudaSuccess && occ > 0)
                gap_grid = (unsigned)(D.n_sm() * occ);
            }
#endif
            MFSC_LAUNCH((k_gaps<true>), gap_grid, ext_block, pk_smem, D.stream, RS, QS, I, P, G, G.n_task + 1);
            D.launches += 2;
            I.gap_res = g_res; I.ord_ready = 1;
            const unsigned pkg_grid = resident_grid((const void*)k_extend<false, true, true>, pk_smem, 7);
            MFSC_LAUNCH((k_extend<false, true, true>), pkg_grid, ext_block, pk_smem, D.stream, RS, QS, I, n_reads, P, R0, st, ext_ticket, pool, DO);
          } else {
            // packed engine over every read, then the general engine over the reads it handed back (usually none)
            MFSC_LAUNCH((k_extend<false, true>), pk_grid, ext_block, pk_smem, D.stream, RS, QS, I, n_reads, P, R0, st, ext_ticket, pool, DO);
          }
          D.launches++;
          ExtIn I2 = I;
          I2.gap_res = nullptr; I2.ord_ready = 0;
          I2.list = x_retry; I2.n_list = I.n_retry; I2.retry = nullptr; I2.n_retry = nullptr;
          MFSC_LAUNCH((k_extend<false, false>), ext_grid, ext_block, ext_smem, D.stream, RS, QS, I2, n_reads, P, R0, st, ext_ticket + 1, pool, DO);
        }
      } else {
        // delta mode: the kernel instantiations that record one traceback byte per cell (mfsc_extend.cuh), packed
        // engine first and the general one over the reads it hands back, like the rows-only path above
        const bool tb_packed = p->engine != 1;
#ifdef MFSC_EMU
        const int tb_block = 32; const size_t tb_smem = 0, tbpk_smem = 0;
        const unsigned tb_grid = grid_for(D, (long long)n_reads * MFSC_LANES, tb_block, 1), tbpk_grid = tb_grid;
#else
        const int tb_block = ext_block; const size_t tb_smem = ext_smem, tbpk_smem = pk_smem;
        const unsigned tb_grid = resident_grid((const void*)k_extend<true, false>, tb_smem, 4);
        const unsigned tbpk_grid = tb_packed ? resident_grid((const void*)k_extend<true, true>, tbpk_smem, 6) : 0;
#endif
        const size_t n_warps = (size_t)std::max(tb_grid, tbpk_grid) * (tb_block / MFSC_LANES);
        // token arena per warp: a read of L bases has at most ~2L columns per alignment, two tokens per indel
        int max_qlen = 0;
        for (int r = 0; r < n_reads; r++) max_qlen = std::max(max_qlen, rd->q.h_len[2 * r]);
        const int arena_ints = std::max(1 << 16, 4 * max_qlen + 4096);
        DA(unsigned char, p_tb, n_warps * TB_DIAGS * DP_SLOTS); DA(unsigned char, p_rev, n_warps * TB_DIAGS); DA(int, p_arena, n_warps * (size_t)arena_ints);
        pool.arena_ints = arena_ints;
        DA(int, x_head, M); DA(int, x_tail, M); DA(int, x_dlb, M); DA(int, x_dle, M);
        // delta buffer: sized for noisy reads (one indel per two bases); a batch that needs more (all-vs-all read alignment:
        // many alignments per read) reports how much, and the extension is run once more with that much
        long long dcap = rd->q.n_bases / 2 + (1 << 20);
        DA(unsigned long long, d_used, 2);
        pool.tb = p_tb; pool.rev = p_rev; pool.arena = p_arena;
        R0.head = x_head; R0.tail = x_tail; R0.dl_begin = x_dlb; R0.dl_end = x_dle;
        for (int attempt = 0; attempt < 2; attempt++) {
          int* d_delta = (int*)keep(D.alloc(sizeof(int) * (size_t)(dcap + 1)));
          if (!d_delta) { cleanup(); delete res; return fail(MFSC_ERR_NOMEM, "out of device memory (delta lists)"); }
          D.zero(d_used, sizeof(unsigned long long) * 2);
          D.zero(d_stats, sizeof(unsigned long long) * 8);
          DO.buf = d_delta; DO.used = d_used; DO.cap = dcap; DO.n_bad = (unsigned*)(d_used + 1);
          if (!tb_packed) {
            MFSC_LAUNCH((k_extend<true, false>), tb_grid, tb_block, tb_smem, D.stream, RS, QS, I, n_reads, P, R0, st, ext_ticket, pool, DO);
          } else {
            MFSC_LAUNCH((k_extend<true, true>), tbpk_grid, tb_block, tbpk_smem, D.stream, RS, QS, I, n_reads, P, R0, st, ext_ticket, pool, DO);
            D.launches++;
            ExtIn I2 = I;
            I2.list = x_retry; I2.n_list = I.n_retry; I2.retry = nullptr; I2.n_retry = nullptr;
            MFSC_LAUNCH((k_extend<true, false>), tb_grid, tb_block, tb_smem, D.stream, RS, QS, I2, n_reads, P, R0, st, ext_ticket + 1, pool, DO);
          }
          unsigned long long used0 = 0;
          D.d2h(&used0, d_used, sizeof(used0));
          if ((long long)used0 <= dcap) break;
          dcap = (long long)used0 + (1 << 16);
          D.launches++;
        }
      }
      D.launches++;
      stamp(D, ts[4]);
      // ---- compact rows in read order and bring them back ----
      if (!exclusive_sum_i32(D, r_n, r_off, n_reads + 1)) { cleanup(); delete res; return fail(MFSC_ERR_NOMEM, "out of device memory (scan)"); }
      int total = 0;
      D.d2h(&total, r_off + n_reads, sizeof(int));
      n_rows = total;
      // the eight row columns sit in one buffer, so that they come back with one copy
      DA(int, c_all, 8 * (size_t)n_rows); DA(int, c_dir, n_rows);
      int* c_ref = c_all; int* c_qry = c_all + n_rows; int* c_sA = c_all + 2 * (size_t)n_rows; int* c_eA = c_all + 3 * (size_t)n_rows;
      int* c_sB = c_all + 4 * (size_t)n_rows; int* c_eB = c_all + 5 * (size_t)n_rows; int* c_errs = c_all + 6 * (size_t)n_rows; int* c_cols = c_all + 7 * (size_t)n_rows;
      int *c_dlb = nullptr, *c_dle = nullptr;
      if (want_deltas) {
        c_dlb = (int*)keep(D.alloc(sizeof(int) * (size_t)(n_rows + 1))); c_dle = (int*)keep(D.alloc(sizeof(int) * (size_t)(n_rows + 1)));
        if (!c_dlb || !c_dle) { cleanup(); delete res; return fail(MFSC_ERR_NOMEM, "out of device memory (delta ranges)"); }
      }
      R1.dl_begin = c_dlb; R1.dl_end = c_dle;
      R1.ref = c_ref; R1.qry = c_qry; R1.dir = c_dir; R1.sA = c_sA; R1.eA = c_eA; R1.sB = c_sB; R1.eB = c_eB; R1.errs = c_errs; R1.cols = c_cols; R1.n_rows = r_n;
      MFSC_LAUNCH(k_compact_rows, grid_for(D, (long long)n_reads * MFSC_LANES, 128, 16), 128, 0, D.stream, R0, seg, r_off, n_reads, R1);
      D.launches++;
      res->ref.resize(n_rows); res->qry.resize(n_rows); res->s1.resize(n_rows); res->e1.resize(n_rows); res->s2.resize(n_rows);
      res->e2.resize(n_rows); res->errs.resize(n_rows); res->cols.resize(n_rows);
      {
        std::vector<int> host_rows(8 * (size_t)n_rows);
        D.d2h(host_rows.data(), c_all, sizeof(int) * 8 * (size_t)n_rows);
        std::vector<int>* dst[8] = {&res->ref, &res->qry, &res->s1, &res->e1, &res->s2, &res->e2, &res->errs, &res->cols};
        for (int c = 0; c < 8; c++) memcpy(dst[c]->data(), host_rows.data() + (size_t)c * n_rows, sizeof(int) * (size_t)n_rows);
      }
      if (want_deltas) {
        unsigned long long used[2] = {0, 0};
        D.d2h(used, DO.used, sizeof(used));
        if ((unsigned)(used[1] & 0xffffffffull) != 0 || (long long)used[0] > DO.cap) { cleanup(); delete res; return fail(MFSC_ERR_NOMEM, "delta lists: token arena or delta buffer too small for this batch"); }
        res->deltas.resize(used[0]); res->dl_b.resize(n_rows); res->dl_e.resize(n_rows);
        D.d2h(res->deltas.data(), DO.buf, sizeof(int) * used[0]);
        D.d2h(res->dl_b.data(), c_dlb, sizeof(int) * n_rows); D.d2h(res->dl_e.data(), c_dle, sizeof(int) * n_rows);
      }
      unsigned long long hst[8];
      D.d2h(hst, d_stats, sizeof(hst));
      res->stats[9] = (int64_t)(hst[4] & 0xffffffffull);
      res->stats[10] = (int64_t)hst[5];
      res->stats[11] = (int64_t)(hst[6] & 0xffffffffull); res->stats[12] = (int64_t)hst[7];
      res->stats[0] = (int64_t)hst[0]; res->stats[1] = (int64_t)hst[1]; res->stats[2] = (int64_t)(int)(hst[2] & 0xffffffffull);
      res->stats[8] = n_rows * 32 + 8;
      stamp(D, ts[5]);
    } else { stamp(D, ts[4]); stamp(D, ts[5]); }
  } else { stamp(D, ts[3]); stamp(D, ts[4]); stamp(D, ts[5]); }
  D.sync();
  res->ms[1] = elapsed_ms(D, ts[0], ts[1]); res->ms[2] = elapsed_ms(D, ts[1], ts[2]); res->ms[3] = elapsed_ms(D, ts[2], ts[3]);
  res->ms[4] = elapsed_ms(D, ts[3], ts[4]); res->ms[5] = elapsed_ms(D, ts[4], ts[5]); res->ms[6] = elapsed_ms(D, ts[0], ts[5]);
  for (int i = 0; i < 8; i++) stamp_free(ts[i]);
  res->stats[5] = D.launches - launches0;
  // algorithmic bytes of the seed kernel: packed query once, one bucket descriptor per probe,
  // one position per candidate (bounded below by the matches), packed reference under every reported match
  res->stats[6] = rd->q.n_bases / 2 + 8 * n_probes + 4 * M;
  cleanup();
#undef DA
  std::string msg;
  if (!dev_check(msg)) { delete res; return fail(MFSC_ERR_CUDA, "align: " + msg); }
  *out = res;
  return MFSC_OK;
}

int mfsc_align_batch(mfsc_index* ix, const uint8_t* bases, const int64_t* read_off, int32_t n_reads, const mfsc_params* p, int32_t flags,
                     mfsc_result** out) {
  mfsc_reads* rd = nullptr;
  int rc = mfsc_reads_upload(bases, read_off, n_reads, &rd);
  if (rc) return rc;
  rc = mfsc_align_reads(ix, rd, p, flags, out);
  if (rc == MFSC_OK) {
    (*out)->ms[0] = rd->upload_ms; (*out)->ms[6] += rd->upload_ms; (*out)->stats[7] = rd->h2d_bytes;
    (*out)->stats[5] += 1;
  }
  mfsc_reads_free(rd);
  return rc;
}

int64_t mfsc_result_count(const mfsc_result* r, int32_t what) {
  if (!r) return -1;
  switch (what) {
    case MFSC_N_ROWS: return (int64_t)r->ref.size();
    case MFSC_N_MEMS: return (int64_t)r->m_q.size();
    case MFSC_N_CLUSTERS: return (int64_t)r->c_q.size();
    case MFSC_N_CLUSTER_MATCHES: return (int64_t)r->cm_s1.size();
    case MFSC_N_DELTAS: return (int64_t)r->deltas.size();
  }
  return -1;
}


int mfsc_result_rows(const mfsc_result* r, int32_t* ref_id, int32_t* qry_id, int32_t* s1, int32_t* e1, int32_t* s2, int32_t* e2,
                     int32_t* errors, int32_t* cols) {
  if (!r) return fail(MFSC_ERR_ARG, "result is NULL");
  copy_out(ref_id, r->ref); copy_out(qry_id, r->qry); copy_out(s1, r->s1); copy_out(e1, r->e1); copy_out(s2, r->s2); copy_out(e2, r->e2);
  copy_out(errors, r->errs); copy_out(cols, r->cols);
  return MFSC_OK;
}

int mfsc_result_mems(const mfsc_result* r, int32_t* qry_id, int32_t* strand, int64_t* s1, int32_t* s2, int32_t* len) {
  if (!r) return fail(MFSC_ERR_ARG, "result is NULL");
  copy_out(qry_id, r->m_q); copy_out(strand, r->m_strand); copy_out(s1, r->m_s1); copy_out(s2, r->m_s2); copy_out(len, r->m_len);
  return MFSC_OK;
}

int mfsc_result_clusters(const mfsc_result* r, int32_t* qry_id, int32_t* strand, int32_t* rec, int32_t* first, int32_t* count, int64_t* m_s1,
                         int32_t* m_s2, int32_t* m_len) {
  if (!r) return fail(MFSC_ERR_ARG, "result is NULL");
  copy_out(qry_id, r->c_q); copy_out(strand, r->c_strand); copy_out(rec, r->c_rec); copy_out(first, r->c_first); copy_out(count, r->c_n);
  copy_out(m_s1, r->cm_s1); copy_out(m_s2, r->cm_s2); copy_out(m_len, r->cm_len);
  return MFSC_OK;
}

int mfsc_result_deltas(const mfsc_result* r, int64_t* begin, int64_t* end, int32_t* deltas) {
  if (!r) return fail(MFSC_ERR_ARG, "result is NULL");
  if (r->dl_b.size() != r->ref.size()) return fail(MFSC_ERR_ARG, "the batch was aligned without MFSC_WANT_DELTAS");
  copy_out(begin, r->dl_b); copy_out(end, r->dl_e); copy_out(deltas, r->deltas);
  return MFSC_OK;
}

int mfsc_result_stats(const mfsc_result* r, int64_t stats[16], double ms[16]) {
  if (!r) return fail(MFSC_ERR_ARG, "result is NULL");
  if (stats) memcpy(stats, r->stats, sizeof(r->stats));
  if (ms) memcpy(ms, r->ms, sizeof(r->ms));
  return MFSC_OK;
}

int mfsc_result_free(mfsc_result* r) { delete r; return MFSC_OK; }

struct mfsc_cov {
  int* cov = nullptr;              // [total] coverage of the concatenated contigs
  long long* tile_pref = nullptr;  // [n_tiles + 1] exclusive 64-bit prefix of the tile sums
  long long* off = nullptr;        // [n_contig + 1] device copy of the contig offsets
  std::vector<long long> h_off;
  long long total = 0, n_tiles = 0;
  int n_contig = 0;
};

int mfsc_cov_build(const int32_t* ref_id, const int32_t* s1, const int32_t* e1, int64_t n_rows, const int64_t* contig_off, int32_t n_contig,
                   mfsc_cov** out, double ms[2]) {
  if (n_rows < 0 || !contig_off || n_contig <= 0 || !out) return fail(MFSC_ERR_ARG, "mfsc_cov_build: bad arguments");
  Dev& D = g_dev;
  if (!D.ok()) return fail(MFSC_ERR_CUDA, "no CUDA device (libmfsc_b200 has no CPU fallback)");
  const long long total = contig_off[n_contig];
  for (int64_t r = 0; r < n_rows; r++) {
    if (ref_id[r] < 0 || ref_id[r] >= n_contig) return fail(MFSC_ERR_ARG, "mfsc_coverage: contig id out of range");
    long long len = contig_off[ref_id[r] + 1] - contig_off[ref_id[r]];
    if (s1[r] < 1 || e1[r] > len || s1[r] > e1[r] + 1) return fail(MFSC_ERR_ARG, "mfsc_coverage: interval outside its contig");
  }
  Stamp t0, t1, t2;
  stamp(D, t0);
  const long long n_tiles = (total + SCAN_TILE - 1) / SCAN_TILE;
  mfsc_cov* c = new mfsc_cov();
  c->total = total; c->n_tiles = n_tiles; c->n_contig = n_contig; c->h_off.assign(contig_off, contig_off + n_contig + 1);
  int* d_ref = dalloc<int>(D, n_rows + 1); int* d_s1 = dalloc<int>(D, n_rows + 1); int* d_e1 = dalloc<int>(D, n_rows + 1);
  c->off = dalloc<long long>(D, n_contig + 1);
  int* d_diff = dalloc<int>(D, total + 16); c->cov = dalloc<int>(D, total + 16);
  unsigned long long* d_state = dalloc<unsigned long long>(D, n_tiles + 1);
  long long* d_tsum = dalloc<long long>(D, n_tiles + 2); c->tile_pref = dalloc<long long>(D, n_tiles + 2);
  unsigned* d_ticket = dalloc<unsigned>(D, 4);
  auto release = [&]() { D.free_(d_ref); D.free_(d_s1); D.free_(d_e1); D.free_(d_diff); D.free_(d_state); D.free_(d_tsum); D.free_(d_ticket); };
  if (!d_ref || !d_s1 || !d_e1 || !c->off || !d_diff || !c->cov || !d_state || !d_tsum || !c->tile_pref || !d_ticket) {
    release(); mfsc_cov_free(c); return fail(MFSC_ERR_NOMEM, "out of device memory (coverage)");
  }
  D.h2d(d_ref, ref_id, sizeof(int) * n_rows); D.h2d(d_s1, s1, sizeof(int) * n_rows); D.h2d(d_e1, e1, sizeof(int) * n_rows);
  D.h2d(c->off, contig_off, sizeof(long long) * (n_contig + 1));
  stamp(D, t1);
  D.zero(d_diff, sizeof(int) * (total + 16)); D.zero(d_state, sizeof(unsigned long long) * (n_tiles + 1)); D.zero(d_ticket, sizeof(unsigned) * 4);
  D.zero(d_tsum, sizeof(long long) * (n_tiles + 2));
  if (n_rows > 0) MFSC_LAUNCH(k_cov_diff, grid_for(D, n_rows, 256, 16), 256, 0, D.stream, d_ref, d_s1, d_e1, (long long)n_rows, c->off, total, d_diff);
  if (total > 0) {
    MFSC_LAUNCH(k_cov_scan, grid_for(D, n_tiles * MFSC_LANES, 256, 8), 256, 0, D.stream, d_diff, total, c->cov, d_state, d_ticket);
    MFSC_LAUNCH(k_cov_tile_sums, grid_for(D, n_tiles * MFSC_LANES, 256, 8), 256, 0, D.stream, c->cov, total, d_tsum);
  }
  D.launches += 3;
  if (!exclusive_sum_i64(D, d_tsum, c->tile_pref, n_tiles + 1)) { release(); mfsc_cov_free(c); return fail(MFSC_ERR_NOMEM, "out of device memory (scan)"); }
  stamp(D, t2);
  D.sync();
  if (ms) { ms[0] = elapsed_ms(D, t1, t2); ms[1] = elapsed_ms(D, t0, t2); }
  stamp_free(t0); stamp_free(t1); stamp_free(t2);
  release();
  std::string msg;
  if (!dev_check(msg)) { mfsc_cov_free(c); return fail(MFSC_ERR_CUDA, "coverage: " + msg); }
  *out = c;
  return MFSC_OK;
}

int mfsc_cov_fetch(const mfsc_cov* c, int32_t contig, int32_t* cov_out) {
  if (!c || !cov_out || contig < -1 || contig >= c->n_contig) return fail(MFSC_ERR_ARG, "mfsc_cov_fetch: bad arguments");
  Dev& D = g_dev;
  const long long b = contig < 0 ? 0 : c->h_off[contig], e = contig < 0 ? c->total : c->h_off[contig + 1];
  D.d2h(cov_out, c->cov + b, sizeof(int) * (size_t)(e - b));
  std::string msg;
  if (!dev_check(msg)) return fail(MFSC_ERR_CUDA, "coverage fetch: " + msg);
  return MFSC_OK;
}

int mfsc_cov_interval_sums(const mfsc_cov* c, const int32_t* contig, const int64_t* start, const int64_t* end, int64_t n, int64_t* sums) {
  if (!c || n < 0 || (n > 0 && (!contig || !start || !end || !sums))) return fail(MFSC_ERR_ARG, "mfsc_cov_interval_sums: bad arguments");
  for (int64_t q = 0; q < n; q++)
    if (contig[q] < 0 || contig[q] >= c->n_contig) return fail(MFSC_ERR_ARG, "mfsc_cov_interval_sums: contig id out of range");
  if (n == 0) return MFSC_OK;
  Dev& D = g_dev;
  if (!D.ok()) return fail(MFSC_ERR_CUDA, "no CUDA device (libmfsc_b200 has no CPU fallback)");
  int* d_c = dalloc<int>(D, n); long long* d_a = dalloc<long long>(D, n); long long* d_b = dalloc<long long>(D, n); long long* d_s = dalloc<long long>(D, n);
  if (!d_c || !d_a || !d_b || !d_s) { D.free_(d_c); D.free_(d_a); D.free_(d_b); D.free_(d_s); return fail(MFSC_ERR_NOMEM, "out of device memory (interval sums)"); }
  D.h2d(d_c, contig, sizeof(int) * n); D.h2d(d_a, start, sizeof(long long) * n); D.h2d(d_b, end, sizeof(long long) * n);
  MFSC_LAUNCH(k_cov_interval_sums, grid_for(D, n * MFSC_LANES, 256, 8), 256, 0, D.stream, c->cov, c->tile_pref, c->off, d_c, d_a, d_b, (long long)n, d_s);
  D.launches++;
  D.d2h(sums, d_s, sizeof(long long) * n);
  D.free_(d_c); D.free_(d_a); D.free_(d_b); D.free_(d_s);
  std::string msg;
  if (!dev_check(msg)) return fail(MFSC_ERR_CUDA, "interval sums: " + msg);
  return MFSC_OK;
}

int mfsc_cov_free(mfsc_cov* c) {
  if (!c) return MFSC_OK;
  Dev& D = g_dev;
  D.free_(c->cov); D.free_(c->tile_pref); D.free_(c->off);
  delete c;
  return MFSC_OK;
}

int mfsc_coverage(const int32_t* ref_id, const int32_t* s1, const int32_t* e1, int64_t n_rows, const int64_t* contig_off, int32_t n_contig,
                  int32_t* cov_out, double ms[2]) {
  if (!cov_out) return fail(MFSC_ERR_ARG, "mfsc_coverage: bad arguments");
  mfsc_cov* c = nullptr;
  double bms[2] = {0, 0};
  int rc = mfsc_cov_build(ref_id, s1, e1, n_rows, contig_off, n_contig, &c, bms);
  if (rc) return rc;
  const auto t0 = std::chrono::steady_clock::now();
  rc = mfsc_cov_fetch(c, -1, cov_out);
  if (ms) { ms[0] = bms[0]; ms[1] = bms[1] + std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
  mfsc_cov_free(c);
  return rc;
}

int mfsc_row_flags(int64_t n_rows, const int32_t* ref_id, const int32_t* qry_id, const int32_t* s1, const int32_t* e1, const int32_t* s2,
                   const int32_t* e2, const int32_t* ref_len, int32_t n_ref, const int32_t* qry_len, int32_t n_qry, const int32_t* ref_gid,
                   const int32_t* qry_gid, uint32_t* flags) {
  if (n_rows < 0 || n_ref < 0 || n_qry < 0) return fail(MFSC_ERR_ARG, "mfsc_row_flags: bad arguments");
  if (n_rows == 0) return MFSC_OK;
  if (!ref_id || !qry_id || !s1 || !e1 || !s2 || !e2 || !ref_len || !qry_len || !ref_gid || !qry_gid || !flags)
    return fail(MFSC_ERR_ARG, "mfsc_row_flags: bad arguments");
  for (int64_t r = 0; r < n_rows; r++)
    if (ref_id[r] < 0 || ref_id[r] >= n_ref || qry_id[r] < 0 || qry_id[r] >= n_qry) return fail(MFSC_ERR_ARG, "mfsc_row_flags: record id out of range");
  Dev& D = g_dev;
  if (!D.ok()) return fail(MFSC_ERR_CUDA, "no CUDA device (libmfsc_b200 has no CPU fallback)");
  const size_t n = (size_t)n_rows;
  int* d_rows = dalloc<int>(D, 6 * n); int* d_rl = dalloc<int>(D, 2 * (size_t)n_ref + 1); int* d_ql = dalloc<int>(D, 2 * (size_t)n_qry + 1);
  unsigned* d_f = dalloc<unsigned>(D, n);
  auto release = [&]() { D.free_(d_rows); D.free_(d_rl); D.free_(d_ql); D.free_(d_f); };
  if (!d_rows || !d_rl || !d_ql || !d_f) { release(); return fail(MFSC_ERR_NOMEM, "out of device memory (row flags)"); }
  const int32_t* cols[6] = {ref_id, qry_id, s1, e1, s2, e2};
  for (int c = 0; c < 6; c++) D.h2d(d_rows + c * n, cols[c], sizeof(int) * n);
  D.h2d(d_rl, ref_len, sizeof(int) * n_ref); D.h2d(d_rl + n_ref, ref_gid, sizeof(int) * n_ref);
  D.h2d(d_ql, qry_len, sizeof(int) * n_qry); D.h2d(d_ql + n_qry, qry_gid, sizeof(int) * n_qry);
  MFSC_LAUNCH(k_row_flags, grid_for(D, n_rows, 256, 8), 256, 0, D.stream, (long long)n_rows, d_rows, d_rows + n, d_rows + 2 * n, d_rows + 3 * n,
              d_rows + 4 * n, d_rows + 5 * n, d_rl, d_ql, d_rl + n_ref, d_ql + n_qry, d_f);
  D.launches++;
  D.d2h(flags, d_f, sizeof(unsigned) * n);
  release();
  std::string msg;
  if (!dev_check(msg)) return fail(MFSC_ERR_CUDA, "row flags: " + msg);
  return MFSC_OK;
}

// ------------------------------------------------------------------------------------------
// show-coords / delta text (host)
// ------------------------------------------------------------------------------------------
static std::vector<const char*> split_names(const char* blob, int n) {
  std::vector<const char*> v(n);
  const char* p = blob;
  for (int i = 0; i < n; i++) { v[i] = p; p += strlen(p) + 1; }
  return v;
}

int mfsc_write_coords(const char* path, const char* ref_path, const char* qry_path, int64_t n_rows, const int32_t* ref_id, const int32_t* qry_id,
                      const int32_t* s1, const int32_t* e1, const int32_t* s2, const int32_t* e2, const int32_t* errors, const int32_t* cols,
                      const char* ref_names, int32_t n_ref, const char* qry_names, int32_t n_qry, int32_t sort_by_ref) {
  if (!path || n_rows < 0) return fail(MFSC_ERR_ARG, "mfsc_write_coords: bad arguments");
  std::vector<const char*> rn = split_names(ref_names, n_ref), qn = split_names(qry_names, n_qry);
  std::vector<int64_t> order(n_rows);
  for (int64_t i = 0; i < n_rows; i++) order[i] = i;
  if (sort_by_ref) {
    // show-coords -r: by reference id (string compare), then reference start; ties keep delta order
    std::vector<int> rank(n_ref);
    std::vector<int> ids(n_ref);
    for (int i = 0; i < n_ref; i++) ids[i] = i;
    std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { return strcmp(rn[a], rn[b]) < 0; });
    for (int i = 0; i < n_ref; i++) rank[ids[i]] = i;
    for (int i = 1; i < n_ref; i++) if (strcmp(rn[ids[i]], rn[ids[i - 1]]) == 0) rank[ids[i]] = rank[ids[i - 1]];
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
      if (rank[ref_id[a]] != rank[ref_id[b]]) return rank[ref_id[a]] < rank[ref_id[b]];
      return s1[a] < s1[b];
    });
  }
  FILE* f = fopen(path, "w");
  if (!f) return fail(MFSC_ERR_IO, std::string("cannot open ") + path);
  fprintf(f, "%s %s\nNUCMER\n\n    [S1]     [E1]  |     [S2]     [E2]  |  [LEN 1]  [LEN 2]  |  [%% IDY]  | [TAGS]\n", ref_path, qry_path);
  fprintf(f, "=====================================================================================\n");
  for (int64_t x = 0; x < n_rows; x++) {
    int64_t i = order[x];
    int len1 = abs(e1[i] - s1[i]) + 1, len2 = abs(e2[i] - s2[i]) + 1;
    float tot = (float)cols[i];
    float idy = (tot - (float)errors[i]) / tot * 100.0f;
    fprintf(f, "%8d %8d  | %8d %8d  | %8d %8d  | %8.2f  | %s\t%s\n", s1[i], e1[i], s2[i], e2[i], len1, len2, (double)idy, rn[ref_id[i]], qn[qry_id[i]]);
  }
  fclose(f);
  return MFSC_OK;
}

int mfsc_write_delta(const char* path, const char* ref_path, const char* qry_path, int64_t n_rows, const int32_t* ref_id, const int32_t* qry_id,
                     const int32_t* s1, const int32_t* e1, const int32_t* s2, const int32_t* e2, const int32_t* errors, const int32_t* cols, const char* ref_names,
                     const int32_t* ref_len, int32_t n_ref, const char* qry_names, const int32_t* qry_len, int32_t n_qry,
                     const int64_t* d_begin, const int64_t* d_end, const int32_t* deltas) {
  if (!path || n_rows < 0) return fail(MFSC_ERR_ARG, "mfsc_write_delta: bad arguments");
  std::vector<const char*> rn = split_names(ref_names, n_ref), qn = split_names(qry_names, n_qry);
  FILE* f = fopen(path, "w");
  if (!f) return fail(MFSC_ERR_IO, std::string("cannot open ") + path);
  std::vector<char> iobuf(1 << 20);
  setvbuf(f, iobuf.data(), _IOFBF, iobuf.size());
  fprintf(f, "%s %s\nNUCMER\n", ref_path, qry_path);
  if (!deltas)   // not a MUMmer delta: say so where any reader of the file meets it first
    fprintf(f, "#MFSC rows-only: alignment headers without indel lists, field 7 = alignment columns; "
               "alignerRobot.fullDelta = True writes MUMmer delta records\n");
  int last_r = -1, last_q = -1;
  for (int64_t i = 0; i < n_rows; i++) {
    if (ref_id[i] != last_r || qry_id[i] != last_q) {
      fprintf(f, ">%s %s %d %d\n", rn[ref_id[i]], qn[qry_id[i]], ref_len[ref_id[i]], qry_len[qry_id[i]]);
      last_r = ref_id[i]; last_q = qry_id[i];
    }
    if (deltas && d_begin && d_begin[i] >= 0) {
      // a real MUMmer delta record: header, signed distances to the next indel, terminating 0
      fprintf(f, "%d %d %d %d %d %d 0\n", s1[i], e1[i], s2[i], e2[i], errors[i], errors[i]);
      // tens of millions of short lines: format by hand into a block buffer
      char blk[1 << 16];
      size_t w = 0;
      for (int64_t k = d_begin[i]; k < d_end[i]; k++) {
        if (w + 16 > sizeof(blk)) { fwrite(blk, 1, w, f); w = 0; }
        int v = deltas[k];
        unsigned u = v < 0 ? (unsigned)(-(long long)v) : (unsigned)v;
        if (v < 0) blk[w++] = '-';
        char tmp[12]; int n = 0;
        do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
        while (n) blk[w++] = tmp[--n];
        blk[w++] = '\n';
      }
      blk[w++] = '0'; blk[w++] = '\n';
      fwrite(blk, 1, w, f);
    } else {
      // header only (the default traceback-free DP produces no indel list); the 7th field carries the column count
      fprintf(f, "%d %d %d %d %d %d %d\n0\n", s1[i], e1[i], s2[i], e2[i], errors[i], errors[i], cols[i]);
    }
  }
  fclose(f);
  return MFSC_OK;
}

// ------------------------------------------------------------------------------------------
// FASTA ingest / part writing (host)
// ------------------------------------------------------------------------------------------
// header lines of data[b, e), b at the start of a line
static int64_t fasta_count_lines(const uint8_t* data, int64_t b, int64_t e) {
  int64_t c = 0, i = b;
  while (i < e) {
    if (data[i] == '>') c++;
    const void* nl = memchr(data + i, '\n', (size_t)(e - i));
    if (!nl) break;
    i = (const uint8_t*)nl - data + 1;
  }
  return c;
}

static int fasta_threads(int64_t n) {
  if (n <= (int64_t)(8 << 20)) return 1;
  return (int)std::min<int64_t>(16, std::max(1u, std::thread::hardware_concurrency()));
}

int mfsc_fasta_count(const uint8_t* data, int64_t n, int32_t* n_rec_out) {
  if (!data || n < 0 || !n_rec_out) return fail(MFSC_ERR_ARG, "mfsc_fasta_count: bad arguments");
  // chunks that begin at the start of a line, counted by a few host threads
  const int T = fasta_threads(n);
  std::vector<int64_t> cb(1, 0);
  for (int t = 1; t < T; t++) {
    int64_t p = n * t / T;
    if (p <= cb.back()) continue;
    const void* nl = memchr(data + p, '\n', (size_t)(n - p));
    if (!nl) break;
    p = (const uint8_t*)nl - data + 1;
    if (p < n && p > cb.back()) cb.push_back(p);
  }
  cb.push_back(n);
  const int C = (int)cb.size() - 1;
  std::vector<int64_t> cnt(C, 0);
  {
    std::vector<std::thread> th;
    for (int c = 1; c < C; c++) th.emplace_back([&, c]() { cnt[c] = fasta_count_lines(data, cb[c], cb[c + 1]); });
    cnt[0] = fasta_count_lines(data, cb[0], cb[1]);
    for (auto& x : th) x.join();
  }
  int64_t total = 0;
  for (int c = 0; c < C; c++) total += cnt[c];
  if (total > 0x7fffffff) return fail(MFSC_ERR_ARG, "mfsc_fasta_count: more than 2^31 records");
  *n_rec_out = (int32_t)total;
  return MFSC_OK;
}

// one chunk of a FASTA buffer that starts at a header line (or at byte 0): either counts its records and bases
// (bases_out == nullptr) or writes them at the given record / base offsets
static void fasta_chunk(const uint8_t* data, int64_t b, int64_t e, bool first_chunk, int64_t rec0, int64_t w0, uint8_t* bases_out, int64_t* off_out,
                        int64_t* name_span_out, int64_t* n_rec, int64_t* n_bases) {
  int64_t rec = 0, w = w0, i = b;
  bool in_rec = !first_chunk;   // later chunks begin with a header by construction
  (void)in_rec;
  bool seen = false;
  while (i < e) {
    const void* nl = memchr(data + i, '\n', (size_t)(e - i));
    const int64_t le0 = nl ? (const uint8_t*)nl - data : e;
    int64_t le = le0;
    while (le > i && (data[le - 1] == '\r')) le--;
    if (le > i) {
      if (data[i] == '>') {
        if (bases_out) { off_out[rec0 + rec] = w; name_span_out[2 * (rec0 + rec)] = i + 1; name_span_out[2 * (rec0 + rec) + 1] = le; }
        rec++; seen = true;
      } else if (seen) {
        if (bases_out) memcpy(bases_out + w, data + i, (size_t)(le - i));
        w += le - i;
      }
    }
    i = le0 + 1;
  }
  *n_rec = rec; *n_bases = w - w0;
}

int mfsc_fasta_parse(const uint8_t* data, int64_t n, uint8_t* bases_out, int64_t* off_out, int64_t* name_span_out, int32_t max_rec,
                     int32_t* n_rec_out) {
  if (!data || n < 0 || !bases_out || !off_out || !name_span_out || !n_rec_out) return fail(MFSC_ERR_ARG, "mfsc_fasta_parse: bad arguments");
  // chunks that begin at header lines, parsed by a few host threads: count, prefix sums, copy
  const int T = fasta_threads(n);
  std::vector<int64_t> cb(1, 0);
  for (int t = 1; t < T; t++) {
    int64_t p = n * t / T;
    if (p <= cb.back()) continue;
    // next "\n>" at or after p
    while (p < n) {
      const void* nl = memchr(data + p, '\n', (size_t)(n - p));
      if (!nl) { p = n; break; }
      p = (const uint8_t*)nl - data + 1;
      if (p < n && data[p] == '>') break;
    }
    if (p < n && p > cb.back()) cb.push_back(p);
  }
  cb.push_back(n);
  const int C = (int)cb.size() - 1;
  std::vector<int64_t> cr(C, 0), cw(C, 0);
  {
    std::vector<std::thread> th;
    for (int c = 1; c < C; c++) th.emplace_back([&, c]() { fasta_chunk(data, cb[c], cb[c + 1], false, 0, 0, nullptr, nullptr, nullptr, &cr[c], &cw[c]); });
    fasta_chunk(data, cb[0], cb[1], true, 0, 0, nullptr, nullptr, nullptr, &cr[0], &cw[0]);
    for (auto& x : th) x.join();
  }
  int64_t rec = 0, w = 0;
  std::vector<int64_t> r0(C), w0(C);
  for (int c = 0; c < C; c++) { r0[c] = rec; w0[c] = w; rec += cr[c]; w += cw[c]; }
  if (rec > max_rec) { *n_rec_out = (int32_t)rec; return fail(MFSC_ERR_ARG, "mfsc_fasta_parse: more records than max_rec"); }
  {
    std::vector<std::thread> th;
    int64_t dr, dw;
    for (int c = 1; c < C; c++) th.emplace_back([&, c]() { int64_t a, b2; fasta_chunk(data, cb[c], cb[c + 1], false, r0[c], w0[c], bases_out, off_out, name_span_out, &a, &b2); });
    fasta_chunk(data, cb[0], cb[1], true, r0[0], w0[0], bases_out, off_out, name_span_out, &dr, &dw);
    for (auto& x : th) x.join();
  }
  if (rec == 0) off_out[0] = 0;
  off_out[rec] = w;
  *n_rec_out = (int32_t)rec;
  return MFSC_OK;
}

int mfsc_fasta_write(const char* path, const char* names, const int64_t* name_off, const uint8_t* bases, const int64_t* off, int32_t rec_begin,
                     int32_t rec_end, int32_t line_len) {
  if (!path || !names || !name_off || !bases || !off || rec_begin < 0 || rec_end < rec_begin) return fail(MFSC_ERR_ARG, "mfsc_fasta_write: bad arguments");
  FILE* f = fopen(path, "wb");
  if (!f) return fail(MFSC_ERR_IO, std::string("cannot open ") + path);
  // formatted by hand into a block buffer (a wrapped 10 kb read is 167 lines: no stdio call per line)
  std::vector<char> buf((size_t)4 << 20);
  size_t w = 0;
  auto flush = [&](size_t need) { if (w + need > buf.size()) { fwrite(buf.data(), 1, w, f); w = 0; } };
  for (int32_t r = rec_begin; r < rec_end; r++) {
    const size_t nl = (size_t)(name_off[r + 1] - name_off[r]);
    if (nl + 2 > buf.size()) { fwrite(buf.data(), 1, w, f); w = 0; fputc('>', f); fwrite(names + name_off[r], 1, nl, f); fputc('\n', f); }
    else { flush(nl + 2); buf[w++] = '>'; memcpy(&buf[w], names + name_off[r], nl); w += nl; buf[w++] = '\n'; }
    const int64_t b = off[r], e = off[r + 1];
    if (line_len <= 0) {
      if ((size_t)(e - b) + 1 > buf.size()) { fwrite(buf.data(), 1, w, f); w = 0; fwrite(bases + b, 1, (size_t)(e - b), f); fputc('\n', f); }
      else { flush((size_t)(e - b) + 1); memcpy(&buf[w], bases + b, (size_t)(e - b)); w += (size_t)(e - b); buf[w++] = '\n'; }
    } else {
      for (int64_t p = b; p < e; p += line_len) {
        const size_t m = (size_t)std::min<int64_t>(line_len, e - p);
        flush(m + 1);
        memcpy(&buf[w], bases + p, m); w += m; buf[w++] = '\n';
      }
    }
  }
  fwrite(buf.data(), 1, w, f);
  fclose(f);
  return MFSC_OK;
}

}  // extern "C"
