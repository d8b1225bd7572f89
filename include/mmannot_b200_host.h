/*
 * mmannot_b200 host front-end -- C view of the C++ host layer (libmmannot_host.so).
 *
 * These functions produce the packed buffers that cross the device boundary declared in
 * mmannot_b200.h: the flattened element table (Config, mmannot.cpp:219-471), the typed
 * intervals in reference order (IntervalList constructor, mmannot.cpp:1094-1290) and the
 * hits decoded from SAM/BAM (Reader/SamReader/BamReader/Read, mmannot.cpp:846-903,
 * 1339-1650).  They run on host cores only and never annotate anything.
 */
#ifndef MMANNOT_B200_HOST_H
#define MMANNOT_B200_HOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mmh_config mmh_config;
typedef struct mmh_annotation mmh_annotation;
typedef struct mmh_reader mmh_reader;

/* All functions returning int: 0 = ok, -1 = error (message from mmh_last_error, thread-local). */
const char *mmh_last_error(void);

int mmh_config_load(const char *path, mmh_config **out);
void mmh_config_free(mmh_config *c);
uint32_t mmh_config_n_elements(const mmh_config *c);
/* fills line[E] (uint16), strand[E], vicinity[E] -- the arrays of mma_params */
void mmh_config_tables(const mmh_config *c, uint16_t *line, uint8_t *strand, uint8_t *vicinity);
/* element display name ("source:type (+)"), NUL-terminated into buf; returns its length */
size_t mmh_config_name(const mmh_config *c, uint32_t element, char *buf, size_t cap);
size_t mmh_config_order_echo(const mmh_config *c, char *buf, size_t cap);

int mmh_annotation_build(const mmh_config *c, const char *gtf_path, uint64_t upstream, uint64_t downstream, mmh_annotation **out);
void mmh_annotation_free(mmh_annotation *a);
uint32_t mmh_annotation_n(const mmh_annotation *a);
uint32_t mmh_annotation_n_chr(const mmh_annotation *a);
uint64_t mmh_annotation_n_genes(const mmh_annotation *a);
uint64_t mmh_annotation_n_lines(const mmh_annotation *a);
const uint32_t *mmh_annotation_chr(const mmh_annotation *a);
const uint32_t *mmh_annotation_start(const mmh_annotation *a);
const uint32_t *mmh_annotation_end(const mmh_annotation *a);
const uint8_t *mmh_annotation_type(const mmh_annotation *a);
const uint8_t *mmh_annotation_strand(const mmh_annotation *a);
const char *mmh_annotation_id(const mmh_annotation *a, uint32_t i);
const char *mmh_annotation_chr_name(const mmh_annotation *a, uint32_t chr);
const char *mmh_annotation_warnings(const mmh_annotation *a);

/* format: 0 guess from suffix, 1 SAM, 2 BAM.  strandedness: 'U', 'F' or 'R' (-s). */
int mmh_reader_open(const mmh_annotation *a, const char *path, int format, char strandedness, mmh_reader **out);
void mmh_reader_close(mmh_reader *r);
/* decodes up to `cap` hits into the five arrays; returns the number written (0 = end of file) */
uint64_t mmh_reader_next(mmh_reader *r, uint64_t cap, uint32_t *start, uint32_t *end, uint32_t *meta, uint32_t *nh, uint64_t *read_key);
uint64_t mmh_reader_records(const mmh_reader *r);
size_t mmh_reader_warnings(mmh_reader *r, char *buf, size_t cap);
/* Read-key verification (the reference keys reads by the name string, mmannot.cpp:1656-1662; the device by mmh_name_key of it):
 * "" while every pair of neighbouring records with one key had one name, else the first pair of different names sharing a key. */
size_t mmh_reader_key_collision(mmh_reader *r, char *buf, size_t cap);

uint64_t mmh_name_key(const char *name, size_t len);

/* ---- synthetic inputs of the benchmark shapes (BASELINE.json configs 2-5) ---- */
typedef struct mmh_synth mmh_synth;
typedef struct mmh_synth_reads {
  uint32_t max_nh;
  int32_t paired;       /* two records per placement, mate flags set */
  int32_t flip_mate2;   /* store mate 2 with the strand bit flipped (-s FR as -s F for the reference) */
  int32_t rna_seq;      /* read length 50..150 instead of 18..30 */
  double p_in_feature;  /* probability that a placement falls inside a gene */
  double p_same_class;  /* probability that all hits of a multi-mapper fall into one gene class */
} mmh_synth_reads;
/* shape: "tair10" | "hs38" | "flybase6"; gene_scale 1.0 = full-size annotation */
int mmh_synth_create(const char *shape, uint64_t seed, double gene_scale, mmh_synth **out);
void mmh_synth_free(mmh_synth *s);
uint64_t mmh_synth_n_genes(const mmh_synth *s);
int mmh_synth_write_annotation(const mmh_synth *s, const char *path);
/* coordinate_sorted: bit 0 = records sorted by (chromosome, position); bit 1 = no BAM header (a part to be appended to another BGZF file);
 * bit 2 = records may straddle BGZF members (htslib never writes such files; the device BAM decoder hands them to the host decoder) */
int mmh_synth_write_bam(const mmh_synth *s, const char *path, uint64_t first_read, uint64_t n_reads, const mmh_synth_reads *spec, int coordinate_sorted);
uint64_t mmh_synth_count_hits(const mmh_synth *s, uint64_t first_read, uint64_t n_reads, const mmh_synth_reads *spec);
/* packed hits of reads [first_read, first_read + n_reads), identical to decoding the BAM written for them */
uint64_t mmh_synth_fill_hits(const mmh_synth *s, const mmh_annotation *a, char strandedness, uint64_t first_read, uint64_t n_reads,
                             const mmh_synth_reads *spec, uint64_t cap, uint32_t *start, uint32_t *end, uint32_t *meta, uint32_t *nh, uint64_t *read_key);

#ifdef __cplusplus
}
#endif
#endif
