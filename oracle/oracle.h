/*
 * TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
 *
 * CPU restatement (plain C) of mmannot's read-annotation hot path, used only as the
 * checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Nothing
 * under mmannot_b200/ may include, link or call this.
 *
 * Parity status: PINNED.  The reference ships no golden vectors or tests
 * (SURVEY.md section 4), so this restatement is pinned against the reference itself:
 * oracle/build_ref.sh compiles /root/reference/mmannot.cpp where it lies into
 * oracle/_ref/, tests/golden/make_golden.py runs it and commits its tables/statistics,
 * and tests/test_oracle_golden.py checks this file against those outputs.
 *
 * It operates on the same packed buffers as the device path (include/mmannot_b200.h)
 * and follows the reference literally: forward linear walk over the chromosome's
 * intervals with per-element (overlap, distance) slots, a by-name map with the NH
 * countdown, and the end-of-file flush.
 */
#ifndef MMANNOT_ORACLE_H
#define MMANNOT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_params {
  int32_t strategy;        /* 0 default, 1 unique, 2 random, 3 ratio */
  float overlap;           /* -l */
  float rescue_threshold;  /* -e / 100 */
  int32_t read_stats;      /* -m given (rescue acts only then) */
  uint32_t n_elements;
  const uint16_t *elem_line;
  const uint8_t *elem_strand;
  const uint8_t *elem_vicinity;
  uint32_t rand_seed;
} orc_params;

typedef struct orc_features {
  uint32_t n, n_chr;
  const uint32_t *chr, *start, *end;
  const uint8_t *type, *strand;
} orc_features;

typedef struct orc_hits {
  uint64_t n;
  const uint32_t *start, *end, *meta, *nh;
  const uint64_t *read_key;
} orc_hits;

typedef struct orc_result {
  uint64_t n_hits, n_reads, n_unique, n_ambiguous, n_multiple, n_unassigned, n_rescued;
  uint64_t n_rows;
  uint64_t *row_mask;   /* sorted ascending */
  double *row_value;    /* the reference's regionCounts value (file-order double accumulation) */
  uint64_t *hit_mask;   /* per-hit element set (n entries) when requested, else NULL */
} orc_result;

/* Per-hit annotation only: element bitmask of every hit (mm:1291-1332, 1018-1076). */
void orc_annotate(const orc_params *p, const orc_features *f, const orc_hits *h, uint64_t *hit_mask);

/* Whole path for one sample.  Returns 0, or -1 on allocation failure / bad input. */
int orc_run(const orc_params *p, const orc_features *f, const orc_hits *h, int want_hit_masks, orc_result *out);
void orc_free(orc_result *r);

/* glibc rand() (TYPE_3 additive feedback) restated; fills out[0..n) after srand(seed). */
void orc_glibc_rand(uint32_t seed, uint64_t n, uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif
