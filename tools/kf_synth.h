/* kf_synth.h -- synthetic input generators of tools/libkfsynth.so (bench / test infrastructure, not product). */
#ifndef KF_SYNTH_H
#define KF_SYNTH_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- synthetic inputs (bench / tests) ---------------------- */
/* Deterministic bacterial-size FASTA: GC ~ U(0.30,0.70) from seed, n_bases split into 1..50 contigs,
 * 10 N-runs of 1..100, upper case, line_width-column lines, LF, header ">g<id>_c<j> synthetic".
 * Call with out == NULL to get the exact size.  Returns bytes written or a negative error. */
int64_t kf_synth_fasta(uint64_t seed, int64_t genome_id, int64_t n_bases, int line_width, uint8_t *out,
                       size_t out_len);
/* Same with the contig count capped at max_contigs and n_runs N-runs (kf_synth_fasta = 50, 10). */
int64_t kf_synth_fasta_ex(uint64_t seed, int64_t genome_id, int64_t n_bases, int line_width, int max_contigs,
                          int n_runs, uint8_t *out, size_t out_len);
/* 4-line FASTQ: n_reads x read_len sampled from a seed-derived genome of genome_len bases, random
 * strand, per-base N 0.2 %, 1 % of reads with an N-run, qualities '!'..'J' (may start with '@'/'+'). */
int64_t kf_synth_fastq(uint64_t seed, int64_t sample_id, int64_t genome_len, int64_t n_reads,
                       int read_len, uint8_t *out, size_t out_len);

#ifdef __cplusplus
}
#endif
#endif
