/*
 * kfcount.h -- C ABI of libkfcount.so, the B200-native k-mer frequency engine.
 *
 * This is the drop-in boundary for kf2vec's k-mer frequency step.  The reference has no FFI of its
 * own: its boundary is the Python callable get_frequencies(args) (kf2vec/main.py:250-373), which
 * shells out to `jellyfish count -m k -s 100M -t p -C` (main.py:308-311) and `jellyfish dump -c`
 * (main.py:317-319) per input file and post-processes with pandas (main.py:323-357).  Every entry
 * point below names the reference lines it replaces.  INTEGRATION.md shows the ctypes stub a
 * kf2vec maintainer would add.
 *
 * Conventions: plain C types only; caller allocates every output; return 0 (KF_OK) or a negative
 * KF_ERR_* code; no exceptions cross the boundary; "d_" parameters are device pointers on the
 * device selected by kf_init, everything else is host memory.  There is no CPU fallback: without a
 * usable CUDA device every compute entry point returns KF_ERR_NO_DEVICE.
 */
#ifndef KFCOUNT_H
#define KFCOUNT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KF_ABI_VERSION 1

/* error codes */
#define KF_OK 0
#define KF_ERR_ARG (-1)          /* bad argument (k out of range, null pointer, ...) */
#define KF_ERR_NO_DEVICE (-2)    /* no CUDA device / kf_init not called / wrong architecture */
#define KF_ERR_CUDA (-3)         /* a CUDA runtime call failed; kf_last_cuda_error() has the text */
#define KF_ERR_IO (-4)           /* file could not be read / written */
#define KF_ERR_FORMAT (-5)       /* first byte is neither '>' nor '@' (jellyfish: "unsupported format") */
#define KF_ERR_FASTQ (-6)        /* FASTQ is not in 4-line layout: a line after a sequence line does not start with '+' */
#define KF_ERR_NOMEM (-7)        /* host or device allocation failed */
#define KF_ERR_LAYOUT (-8)       /* device arena violates the layout contract of kf_count_device */
#define KF_ERR_EMPTY (-9)        /* zero-length input (jellyfish fails on it; the reference then crashes) */
#define KF_ERR_UNSUPPORTED (-10)  /* valid input that this entry point does not take (FASTQ in the sparse path) */

/* flags (bit set) */
#define KF_FLAG_PSEUDOCOUNT 1u   /* main.py:332-334  counts += 0.5 before normalising           */
#define KF_FLAG_RAW_CNT 2u       /* main.py:340-342  skip the normalisation                      */
#define KF_FLAG_FORCE_WALKER 4u  /* debug: disable the vectorised fast path (byte walker only)   */
#define KF_FLAG_NO_LINEGRID 8u   /* debug: generic kernels only (no fixed-line-width kernel at k = 7, no partitioned kernel at k = 8..10) */
#define KF_FLAG_PART_ALL 16u     /* debug: k = 8..10 partitioned kernel for files of any size (default: >= 256 KiB) */

/* limits */
#define KF_MIN_K 1
#define KF_MAX_K 12              /* dense canonical output up to k=12 (8,390,656 columns)        */
#define KF_SPARSE_MIN_K 6        /* sparse (code, count) output: kf_sparse_*                     */
#define KF_SPARSE_MAX_K 31       /* 2-bit codes in 64 bits; the reference's -k goes to 31 (main.py:81-82) */
#define KF_MAX_K_SMEM 7          /* 4^k u32 bins privatised in shared memory up to here          */
#define KF_CHUNK 512             /* arena alignment unit: one warp-load of 32 x 16 bytes         */
#define KF_TAIL_PAD 4096         /* NUL bytes required after the last file of a device arena     */

/* ---- lifecycle ------------------------------------------------------------------------------- */
/* Selects the CUDA device, checks it is sm_100, creates the library's stream and workspace.
 * Replaces nothing in the reference (Jellyfish needs no init); one call per process/rank. */
int kf_init(int device);
int kf_shutdown(void);
int kf_device(void);                       /* device chosen by kf_init, or KF_ERR_NO_DEVICE       */
const char *kf_strerror(int code);
const char *kf_last_cuda_error(void);
int kf_abi_version(void);

/* ---- vocabulary: kf2vec/data/<vocab file> + main.py:278-296 ----------------------------------- */
/* Number of canonical k-mers = number of columns of a .kf row: (4^k + [k even] 4^(k/2)) / 2. */
int64_t kf_vocab_size(int k);
/* Writes the sorted canonical k-mers, one per line ("AAAAAAA\n..."), V*(k+1) bytes, to out. */
int kf_vocab(int k, char *out, size_t out_len);
/* canonical 2-bit codes (A0 C1 G2 T3, first base most significant) in column order */
int kf_vocab_codes(int k, uint32_t *out, size_t n_out);

/* ---- counting, host buffers in / host rows out (the end-to-end call) -------------------------- */
/* Replaces, per input i, `jellyfish count -C` + `jellyfish dump -c` + vocabulary merge + pseudocount
 * + normalise (main.py:308-342).  bufs[i] holds the raw bytes of one .fa/.fna/.fasta/.fq/.fastq file.
 *   counts_out [n][V] uint64  canonical counts in vocabulary order              (may be NULL)
 *   freq_out   [n][V] double  c/sum(c), or c (+0.5) with KF_FLAG_RAW_CNT         (may be NULL)
 *   totals_out [n]    uint64  number of valid k-mers                             (may be NULL)
 *   status_out [n]    int     KF_OK or the per-file error                        (required)
 * Host->device copies, all kernels and device->host copies happen inside the call. */
int kf_count_buffers(const uint8_t *const *bufs, const size_t *lens, int n, int k, uint32_t flags,
                     uint64_t *counts_out, double *freq_out, uint64_t *totals_out, int *status_out);

/* Same, reading the files itself (the loop body of main.py:301-357 without the text formatting). */
int kf_count_files(const char *const *paths, int n, int k, uint32_t flags,
                   uint64_t *counts_out, double *freq_out, uint64_t *totals_out, int *status_out);

/* The whole loop of get_frequencies (kf2vec/main.py:301-370: per file jellyfish count + dump, merge, pseudocount,
 * normalise, write <sample>.kf) for n files, as a three-stage pipeline over batches of at most batch_bytes (0: 256 MiB):
 * `threads` host threads (0: all) read batch b+1 into pinned memory while the GPU counts batch b and the threads format
 * and write the rows of batch b-1.  out_paths[i] receives the row of in_paths[i] labelled samples[i] (text exactly as
 * kf_write_kf); files whose status is not KF_OK get no output.  totals_out [n] (may be NULL); stage_seconds [4] (may be
 * NULL): host time spent waiting for reads, in the GPU stage, waiting for writes, and in total. */
int kf_files_to_kf(const char *const *in_paths, const char *const *out_paths, const char *const *samples, int n, int k,
                   uint32_t flags, int threads, size_t batch_bytes, int *status_out, uint64_t *totals_out,
                   double *stage_seconds);

/* Files on disk -> the [n, V] float32 feature matrix in DEVICE memory (row i for in_paths[i]: fp32(freq * 1e4), the
 * tensor train_classifier_model.py:144-150,323 / classify.py:102-114 build from the .kf files), through the same
 * pipelined reads and GPU stage, without the text round trip.  Rows of files whose status is not KF_OK are undefined. */
int kf_files_to_device(const char *const *in_paths, int n, int k, uint32_t flags, int threads, size_t batch_bytes,
                       float *d_feat_out, int *status_out, uint64_t *totals_out, double *stage_seconds);

/* ---- chunked-genome mode: one row per sliding window (get_chunks, main.py:813-881) -------------------------- */
/* Replaces, per 10-kbp chunk, `seqkit sliding` + `seqkit split` + the jellyfish count/dump pair the reference runs
 * on every chunk file (main.py:824-838, 869-881).  seq is the linearised, N-collapsed, gap-stripped sequence of one
 * or more contigs (main.py:732-753, prepared by the host); window i is seq[win_off[i] .. win_off[i]+win_len[i]) and
 * may overlap its neighbours.  Every byte that is not A/C/G/T (either case) breaks the k-mer window; k-mers never
 * extend beyond their window.  Outputs as kf_count_buffers, one row per window (any may be NULL). */
int kf_count_windows(const uint8_t *seq, size_t seq_len, const uint64_t *win_off, const uint32_t *win_len, int n, int k,
                     uint32_t flags, uint64_t *counts_out, double *freq_out, uint64_t *totals_out);

/* ---- counting, device-resident arena (kernel-only path; trainer hand-off) --------------------- */
/* Arena layout contract: file i occupies d_arena[offsets[i] .. offsets[i]+lens[i]); offsets[i] is a
 * multiple of KF_CHUNK; files are in increasing offset order and do not overlap; every byte of the
 * arena that belongs to no file is 0; at least KF_TAIL_PAD zero bytes follow the last file
 * (arena_bytes says how much is allocated).  formats[i] is '>' or '@' (first byte of file i).
 * Outputs are device pointers (any may be NULL): d_counts [n][V] uint64, d_freq [n][V] double,
 * d_feat [n][V] float = fp32(freq * 1e4) (train_classifier_model.py:149,323), d_totals [n] uint64.
 * Work is enqueued on `stream` (a cudaStream_t; NULL = the library's stream) and NOT synchronised. */
int kf_count_device(const uint8_t *d_arena, size_t arena_bytes, const uint64_t *offsets,
                    const uint64_t *lens, const uint8_t *formats, int n, int k, uint32_t flags,
                    uint64_t *d_counts, double *d_freq, float *d_feat, uint64_t *d_totals,
                    void *stream);
/* Per-file status of the last kf_count_device call on this arena (waits for the device): KF_OK, KF_ERR_EMPTY,
 * KF_ERR_FORMAT (first byte neither '>' nor '@'; such files are skipped and yield all-zero rows) or
 * KF_ERR_FASTQ.  jellyfish reports these through its exit code, which main.py:309-311 ignores. */
int kf_last_file_status(const uint8_t *d_arena, const uint64_t *offsets, const uint64_t *lens, const uint8_t *formats, int n,
                        int *status_out);
/* Number of kernels kf_count_device launched in its last call (for bench.py's gpu_launches). */
int kf_last_launch_count(void);
/* Size the persistent counting kernels for n_sms SMs instead of all of them (0 = all): leaves SMs free for a kernel
 * that runs beside them, e.g. the NCCL all-gather of the previous batch's rows (one CTA per SM with ~205 KB of shared
 * memory cannot share an SM, and a CTA that has to wait for one delays the whole launch).  Returns the SM count in
 * effect, or a negative error code.  (No reference counterpart: jellyfish -t p, main.py:309, is the nearest knob.) */
int kf_set_sm_limit(int n_sms);
/* Device time of the counting kernel(s) of the last kf_count_device / kf_count_buffers call, from CUDA
 * events recorded on the launching stream (waits for them).  Used for the roofline figure. */
int kf_last_count_kernel_ms(float *ms);
/* Durations (ms, CUDA events on the launching stream) of the counting kernels of the last n calls, oldest first, at most
 * 64; returns how many were written.  Waits for those calls: meant to be read AFTER a timed loop, so that the loop itself
 * runs without host synchronisation (bench.py: roofline.achieved over the timed region). */
int kf_count_kernel_ms_history(float *ms_out, int n);

/* ---- sparse counting for large k: observed canonical k-mers only (get_kmers, main.py:135-172; jellyfish -k up to 31) ---- */
/* Replaces `jellyfish count -m k -s 100M -C` + `jellyfish dump -c` (main.py:135-145, :308-319) where a dense row of 4^k bins
 * is not sensible: per file, the OBSERVED canonical k-mers (2-bit codes A0 C1 G2 T3, first base most significant -- the
 * order of kf_vocab_codes) with their counts, ascending by code (Jellyfish lists them in hash order).  Sort-and-run-length
 * on the device: radix partition of the canonical codes by their leading bits, per-bucket sort in shared memory, run
 * lengths.  KF_SPARSE_MIN_K <= k <= KF_SPARSE_MAX_K.  FASTA inputs; a FASTQ input gets status KF_ERR_UNSUPPORTED (use the
 * dense entry points up to k = 12).  The result stays in device memory until the next kf_sparse_count* call or
 * kf_sparse_release; n_distinct_out [n] = entries per file, totals_out [n] = valid k-mers per file (either may be NULL). */
int kf_sparse_count(const uint8_t *const *bufs, const size_t *lens, int n, int k, uint64_t *n_distinct_out,
                    uint64_t *totals_out, int *status_out);
/* Same on a device-resident arena (layout contract of kf_count_device); enqueued on `stream` and synchronised before
 * returning (the output size is only known then).  status_out may be NULL. */
int kf_sparse_count_device(const uint8_t *d_arena, size_t arena_bytes, const uint64_t *offsets, const uint64_t *lens,
                           const uint8_t *formats, int n, int k, uint64_t *n_distinct_out, uint64_t *totals_out,
                           int *status_out, void *stream);
/* Entries of the last result over all files, and the copy to host arrays: codes_out / counts_out [total] hold the files'
 * entries back to back in file order, row_off_out [n + 1] where each file's begin (any may be NULL). */
int64_t kf_sparse_total_entries(void);
int kf_sparse_fetch(uint64_t *codes_out, uint32_t *counts_out, uint64_t *row_off_out);
/* The FSW fork's feature matrix of one file of the last result (kf2vec/main.py:147-169, get_kmers): out_host
 * [n_distinct][k + 1] float32, row = the k-mer's k bases as codes A0 T1 C2 G3 (main.py:118) followed by count / divisor in
 * fp32 -- the reference divides the float32 counts by their float32 sum (main.py:165-169), which the caller passes.  Rows in
 * ascending code order (the reference: Jellyfish's hash order).  Expanded on the device, one copy to the host. */
int kf_sparse_kmer_matrix(int file, float divisor, float *out_host);
/* Device view of the last result: it is held as one chunk per internal sub-batch of files [file0, file1), whose entries
 * are the global entries [first_entry, first_entry + n_entries). */
int kf_sparse_chunk_count(void);
int kf_sparse_chunk(int i, const uint64_t **d_codes, const uint32_t **d_counts, uint64_t *n_entries, uint64_t *first_entry,
                    int *file0, int *file1);
int kf_sparse_release(void);

/* ---- .kf writer: main.py:344-357 --------------------------------------------------------------- */
/* Formats one row exactly as pandas `astype(str)` + ",".join does (Python repr of float64: shortest
 * round-trip digits, exponent form when exp10 < -4 or >= 16, "nan" for 0/0).  int_mode != 0 prints
 * integers ("5" not "5.0"): the reference's dtype quirk for -raw_cnt rows with no missing k-mer.
 * Returns the number of bytes written (excluding the NUL) or a negative error. */
int64_t kf_format_row(const char *sample, const double *row, int64_t V, int int_mode, char *out,
                      size_t out_len);
int kf_write_kf(const char *out_path, const char *sample, const double *row, int64_t V, int int_mode,
                int append);

/* All chunk rows of one genome into one file with one open (main.py:895-915).  labels: n NUL-terminated strings back to
 * back; int_modes[i] as int_mode above (may be NULL = all 0). */
int kf_write_kf_rows(const char *out_path, const char *labels, const double *rows, int64_t n, int64_t V, const uint8_t *int_modes,
                     int append);

/* ---- chunked-genome text preparation: seqtk seq -l 0 (main.py:732), awk N-run collapse (:740), seqkit seq -g -m (:753) ---- */
/* One pass over a FASTA buffer: linearise every record, collapse each run of 'N' / 'n' / '|' to one 'N', remove the gap
 * characters '-', '.', ' ', keep the records of at least min_len bytes.  Kept sequences land back to back in seq_out
 * (capacity seq_cap >= len is always enough); record r is seq_out[seq_off[r] .. +seq_len[r]) and its header line (without
 * '>') is data[id_off[r] .. +id_len[r]).  Returns the number of kept records (only the first max_records are stored). */
int64_t kf_linearise_fasta(const uint8_t *data, size_t len, uint64_t min_len, uint8_t *seq_out, size_t seq_cap, uint64_t *seq_off,
                           uint64_t *seq_len, uint64_t *id_off, uint32_t *id_len, int64_t max_records);

/* ---- .kf reader: utils.py:436-437 (my_read_csv), classify.py:102-114, query.py:148-158 ---------------------- */
/* Parses "label,v1,...,vV\n" rows from a text buffer (one or many .kf files concatenated, as query.py:153 does with
 * `cat`).  out [rows][V] double (may be NULL), feat_out [rows][V] float = float(v * 1e4) as the trainers build it
 * (train_classifier_model.py:149,323; may be NULL), label_off/label_len locate each label inside text (may be NULL).
 * Conversion is correctly rounded (repr()-written values round-trip bit-exactly).  Returns the number of rows in
 * the text (only the first max_rows are stored), or KF_ERR_FORMAT (wrong column count / unparsable value). */
int64_t kf_parse_kf(const char *text, size_t len, int64_t V, int64_t max_rows, double *out, float *feat_out,
                    int64_t *label_off, int32_t *label_len);

#ifdef __cplusplus
}
#endif
#endif /* KFCOUNT_H */
