/*
 * TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH (see oracle.h).
 *
 * Plain-C restatement of the reference hot path.  Each function cites the lines of
 * /root/reference/mmannot.cpp ("mm:") it follows.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXE 64
#define CHR_MASK 0x00FFFFFFu
#define NOLINE 0xFFFFFFFFu

typedef uint64_t pos_t; /* Position = unsigned long, mm:54 */

/* ------------------------------------------------------------------ per hit */

/* Interval::overlaps, mm:632-636 */
static pos_t iv_overlaps(pos_t s1, pos_t e1, pos_t s2, pos_t e2) {
  pos_t s = s1 > s2 ? s1 : s2, e = e1 < e2 ? e1 : e2;
  if (s >= e) return 0;
  return e - s;
}
/* Interval::getDistance, mm:661-665 (called on the read) */
static pos_t iv_distance(pos_t rs, pos_t re, pos_t p) {
  if (p < rs) return rs - p;
  if (p > re) return p - re;
  return 0;
}
/* intervalInclusion / intervalOverlapPc / intervalOverlap, mm:992-1002, chosen as at mm:1972-1977 */
static pos_t score(float ovl, pos_t is, pos_t ie, pos_t rs, pos_t re) {
  if (ovl < 0.0f) return (rs >= is && re <= ie) ? 1 : 0;
  pos_t o = iv_overlaps(is, ie, rs, re);
  if (ovl < 1.0f) {
    pos_t size = re - rs + 1; /* Read::getSize, mm:901 */
    return ((float)size * ovl <= (float)o) ? o : 0;
  }
  return ((float)o >= ovl) ? o : 0;
}
/* Config::checkStrand, mm:438-443.  fs: 1 = F, 2 = R; rs: read strand bool */
static int strand_ok(uint8_t es, uint8_t fs, int rs) {
  if (es == 0) return 1;
  if (es == 1) return ((fs == 1) && rs) || ((fs == 2) && !rs);
  return ((fs == 1) && !rs) || ((fs == 2) && rs);
}

#define BIN_SIZE 16384u /* binSize, mm:67 */

typedef struct {
  const orc_params *p;
  const orc_features *f;
  uint32_t *chr_start; /* n_chr + 1 */
  uint32_t *bin_first; /* per chromosome: bins[b] = first interval (in order) whose end / binSize >= b, mm:1277-1284 */
  uint64_t *bin_base;  /* n_chr + 1 offsets into bin_first */
} actx;

/* IntervalList::scan + EvaluationStructure, mm:1291-1332, 1018-1076: start at the bin of the read start
 * (mm:1303-1305), skip the intervals that end before the read (mm:1306-1308), then walk while the
 * interval does not start after the read (mm:1311). */
static uint64_t annotate_one(const actx *a, uint32_t hstart, uint32_t hend, uint32_t meta) {
  const orc_params *p = a->p;
  const orc_features *f = a->f;
  uint32_t chr = meta & CHR_MASK;
  if (chr >= f->n_chr) return 0;
  uint32_t cs = a->chr_start[chr], ce = a->chr_start[chr + 1];
  if (cs == ce) return 0;
  int rstrand = (meta >> 31) & 1;
  pos_t rs = hstart, re = (hend == 0xFFFFFFFFu) ? ~(pos_t)0 : (pos_t)hend; /* empty CIGAR at position 0 wraps, mm:874 */
  pos_t ov[MAXE], di[MAXE];
  uint32_t E = p->n_elements, nset = 0;
  for (uint32_t i = 0; i < E; ++i) ov[i] = di[i] = 0;
  uint32_t v0 = cs;
  if (a->bin_first) {
    uint64_t nb = a->bin_base[chr + 1] - a->bin_base[chr];
    uint64_t bin = rs / BIN_SIZE;
    if (bin > nb - 1) bin = nb - 1;
    v0 = a->bin_first[a->bin_base[chr] + bin];
  }
  while (v0 < ce && (pos_t)f->end[v0] < rs) ++v0; /* isBefore(read), mm:1306-1308 */
  for (uint32_t v = v0; v < ce && !((pos_t)f->start[v] > re); ++v) { /* ! isAfter(read), mm:1311 */
    uint32_t t = f->type[v];
    if (!strand_ok(p->elem_strand[t], f->strand[v], rstrand)) continue;
    pos_t o = score(p->overlap, f->start[v], f->end[v], rs, re);
    if (o == 0) continue;
    pos_t d = 0;
    if (p->elem_vicinity[t] == 1) d = iv_distance(rs, re, f->end[v]);        /* upstream, mm:1317-1319 */
    else if (p->elem_vicinity[t] == 2) d = iv_distance(rs, re, f->start[v]); /* downstream, mm:1320-1322 */
    ov[t] = o; /* later intervals overwrite earlier ones, mm:1023-1028 */
    di[t] = d;
    ++nset;
  }
  if (nset == 0) return 0;
  /* getFirst, mm:1029-1076 */
  uint32_t good = NOLINE, nsel = 0, sel[MAXE];
  pos_t maxo = 0;
  for (uint32_t i = 0; i < E; ++i) {
    uint32_t line = p->elem_line[i];
    if (good != NOLINE && line != good) break;
    if (ov[i] > 0) {
      good = line;
      if (ov[i] > maxo) { nsel = 0; sel[nsel++] = i; maxo = ov[i]; }
      else if (ov[i] == maxo) sel[nsel++] = i;
    }
  }
  if (nsel == 0) return 0;
  if (nsel == 1) return 1ull << sel[0];
  uint64_t out = 0;
  pos_t mind = ~(pos_t)0;
  for (uint32_t k = 0; k < nsel; ++k) {
    uint32_t i = sel[k];
    if (di[i] < mind) { mind = di[i]; out = 1ull << i; }
    else if (di[i] == mind) out |= 1ull << i;
  }
  return out;
}

static int build_chr_start(const orc_features *f, uint32_t **out) {
  uint32_t *cs = (uint32_t *)calloc((size_t)f->n_chr + 2, sizeof(uint32_t));
  if (!cs) return -1;
  /* features are sorted by chromosome id (mm:1267); count then prefix */
  for (uint32_t i = 0; i < f->n; ++i) {
    if (f->chr[i] >= f->n_chr) { free(cs); return -1; }
    cs[f->chr[i] + 1]++;
  }
  for (uint32_t c = 0; c < f->n_chr; ++c) cs[c + 1] += cs[c];
  *out = cs;
  return 0;
}

/* the bins of mm:1277-1284 */
static int build_bins(actx *a) {
  const orc_features *f = a->f;
  a->bin_base = (uint64_t *)calloc((size_t)f->n_chr + 1, sizeof(uint64_t));
  if (!a->bin_base) return -1;
  for (uint32_t c = 0; c < f->n_chr; ++c) {
    uint64_t nb = 0;
    for (uint32_t i = a->chr_start[c]; i < a->chr_start[c + 1]; ++i) {
      uint64_t b = f->end[i] / BIN_SIZE;
      if (nb <= b) nb = b + 1;
    }
    a->bin_base[c + 1] = a->bin_base[c] + nb;
  }
  a->bin_first = (uint32_t *)malloc((a->bin_base[f->n_chr] ? a->bin_base[f->n_chr] : 1) * sizeof(uint32_t));
  if (!a->bin_first) return -1;
  for (uint32_t c = 0; c < f->n_chr; ++c) {
    uint64_t size = 0;
    uint32_t *bins = a->bin_first + a->bin_base[c];
    for (uint32_t i = a->chr_start[c]; i < a->chr_start[c + 1]; ++i) {
      uint64_t b = f->end[i] / BIN_SIZE;
      while (size <= b) bins[size++] = i;
    }
  }
  return 0;
}
static void free_actx(actx *a) { free(a->chr_start); free(a->bin_first); free(a->bin_base); }

void orc_annotate(const orc_params *p, const orc_features *f, const orc_hits *h, uint64_t *hit_mask) {
  actx a;
  memset(&a, 0, sizeof(a));
  a.p = p; a.f = f;
  if (build_chr_start(f, &a.chr_start)) return;
  if (build_bins(&a)) { free_actx(&a); return; }
  for (uint64_t i = 0; i < h->n; ++i) hit_mask[i] = annotate_one(&a, h->start[i], h->end[i], h->meta[i]);
  free_actx(&a);
}

/* ------------------------------------------------------- small hash maps */

typedef struct node {
  uint64_t key;
  struct node *next;
  /* open multi-mapping read (readCounts / rawCounts, mm:1656-1657) */
  uint32_t remaining, raw;
  uint32_t mult[MAXE]; /* the concatenated element list, as multiplicities */
  /* -y random bookkeeping (chosenId / numberSeen / seen, mm:1661-1662) */
  uint8_t has_open, has_chosen, seen;
  uint32_t chosen, number_seen;
  uint64_t first_order; /* insertion order, used to make the EOF flush deterministic */
} node;

typedef struct {
  node **bucket;
  uint64_t nbucket, count;
} nmap;

static uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
static int nmap_init(nmap *m, uint64_t nb) {
  m->nbucket = nb; m->count = 0;
  m->bucket = (node **)calloc(nb, sizeof(node *));
  return m->bucket ? 0 : -1;
}
static node *nmap_find(nmap *m, uint64_t key) {
  for (node *n = m->bucket[mix64(key) & (m->nbucket - 1)]; n; n = n->next)
    if (n->key == key) return n;
  return NULL;
}
static void nmap_grow(nmap *m) {
  uint64_t nb = m->nbucket * 2;
  node **b = (node **)calloc(nb, sizeof(node *));
  if (!b) return;
  for (uint64_t i = 0; i < m->nbucket; ++i) {
    node *n = m->bucket[i];
    while (n) {
      node *nx = n->next;
      uint64_t s = mix64(n->key) & (nb - 1);
      n->next = b[s]; b[s] = n;
      n = nx;
    }
  }
  free(m->bucket);
  m->bucket = b; m->nbucket = nb;
}
static node *nmap_get(nmap *m, uint64_t key) {
  node *n = nmap_find(m, key);
  if (n) return n;
  if (m->count > m->nbucket) nmap_grow(m);
  n = (node *)calloc(1, sizeof(node));
  if (!n) return NULL;
  n->key = key;
  uint64_t s = mix64(key) & (m->nbucket - 1);
  n->next = m->bucket[s]; m->bucket[s] = n;
  m->count++;
  return n;
}
static void nmap_free(nmap *m) {
  for (uint64_t i = 0; i < m->nbucket; ++i) {
    node *n = m->bucket[i];
    while (n) { node *nx = n->next; free(n); n = nx; }
  }
  free(m->bucket);
}

/* regionCounts, mm:1658: element set -> double */
typedef struct {
  uint64_t *key; double *val; uint64_t cap, n;
} cmap;
static int cmap_init(cmap *c) {
  c->cap = 1024; c->n = 0;
  c->key = (uint64_t *)calloc(c->cap, sizeof(uint64_t));
  c->val = (double *)calloc(c->cap, sizeof(double));
  return (c->key && c->val) ? 0 : -1;
}
static double *cmap_slot(cmap *c, uint64_t key) { /* key != 0 */
  if (c->n * 2 > c->cap) {
    uint64_t ncap = c->cap * 2;
    uint64_t *nk = (uint64_t *)calloc(ncap, sizeof(uint64_t));
    double *nv = (double *)calloc(ncap, sizeof(double));
    for (uint64_t i = 0; i < c->cap; ++i)
      if (c->key[i]) {
        uint64_t s = mix64(c->key[i]) & (ncap - 1);
        while (nk[s]) s = (s + 1) & (ncap - 1);
        nk[s] = c->key[i]; nv[s] = c->val[i];
      }
    free(c->key); free(c->val);
    c->key = nk; c->val = nv; c->cap = ncap;
  }
  uint64_t s = mix64(key) & (c->cap - 1);
  while (c->key[s] && c->key[s] != key) s = (s + 1) & (c->cap - 1);
  if (!c->key[s]) { c->key[s] = key; c->val[s] = 0.0; c->n++; }
  return &c->val[s];
}

/* ------------------------------------------------------------- glibc rand */

void orc_glibc_rand(uint32_t seed, uint64_t n, uint32_t *out) {
  /* glibc random_r.c, TYPE_3 (degree 31, separation 3): the stream rand() yields after
   * srand(seed); the reference never seeds, i.e. seed 1 (mm:1711). */
  uint64_t total = 344 + n;
  uint32_t *r = (uint32_t *)malloc(total * sizeof(uint32_t));
  if (!r) return;
  if (seed == 0) seed = 1;
  r[0] = seed;
  for (int i = 1; i < 31; ++i) {
    int32_t hi = (int32_t)r[i - 1] / 127773, lo = (int32_t)r[i - 1] % 127773;
    int32_t word = 16807 * lo - 2836 * hi;
    if (word < 0) word += 2147483647;
    r[i] = (uint32_t)word;
  }
  for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
  for (uint64_t i = 34; i < total; ++i) r[i] = r[i - 31] + r[i - 3];
  for (uint64_t k = 0; k < n; ++k) out[k] = r[k + 344] >> 1;
  free(r);
}

/* ------------------------------------------------------------- per read */

typedef struct {
  const orc_params *p;
  cmap counts;
  nmap names;
  orc_result *out;
  uint32_t rand_buf[4096];
  uint64_t rand_pos; /* index of the next rand() draw */
  uint64_t order;
} rctx;

static uint32_t next_rand(rctx *r) {
  /* regenerate a window of the stream when exhausted; simple and exact */
  uint64_t k = r->rand_pos++;
  uint64_t w = k & 4095;
  if (w == 0 || k == 0) {
    uint32_t *tmp = (uint32_t *)malloc((k + 4096) * sizeof(uint32_t));
    orc_glibc_rand(r->p->rand_seed, k + 4096, tmp);
    memcpy(r->rand_buf, tmp + k, 4096 * sizeof(uint32_t));
    free(tmp);
  }
  return r->rand_buf[w];
}

static uint64_t mult_mask(const uint32_t *mult, uint32_t E) {
  uint64_t m = 0;
  for (uint32_t i = 0; i < E; ++i) if (mult[i]) m |= 1ull << i;
  return m;
}

/* rescue(), mm:497-509, applied to the sorted list (printReadStats sorts first, mm:475).
 * Only reachable when -m is given and -e < 100 (mm:491, 2001, 2025). */
static uint64_t apply_rescue(const orc_params *p, const uint32_t *mult, uint64_t mask) {
  if (!p->read_stats || !(p->rescue_threshold < 1.0f)) return mask;
  uint64_t n = 0;
  for (uint32_t i = 0; i < p->n_elements; ++i) n += mult[i];
  if (n == 1) return mask;
  size_t t = (size_t)ceilf((float)n * p->rescue_threshold);
  for (uint32_t i = 0; i < p->n_elements; ++i)
    if (mult[i] && mult[i] >= t) return 1ull << i;
  return mask;
}

static int popcount64(uint64_t x) { int c = 0; while (x) { x &= x - 1; ++c; } return c; }

/* Counter::addCount, mm:1665-1739 */
static int add_count(rctx *r, uint64_t key, uint64_t mask, uint32_t nh) {
  const orc_params *p = r->p;
  orc_result *o = r->out;
  int nreg = popcount64(mask);
  if (nreg == 0) o->n_unassigned++;
  else if (nreg > 1) o->n_ambiguous++;
  else if (nh == 1) o->n_unique++;
  if (nh > 1 && p->strategy == 0) {
    o->n_multiple++;
    node *n = nmap_get(&r->names, key);
    if (!n) return -1;
    if (!n->has_open) {
      n->has_open = 1;
      n->remaining = nh - 1;
      n->raw = nh;
      n->first_order = r->order++;
      memset(n->mult, 0, sizeof(n->mult));
      for (uint32_t i = 0; i < p->n_elements; ++i) if (mask >> i & 1) n->mult[i]++;
      o->n_reads++;
    } else {
      n->remaining--;
      for (uint32_t i = 0; i < p->n_elements; ++i) if (mask >> i & 1) n->mult[i]++;
      if (n->remaining == 0) {
        uint64_t m = mult_mask(n->mult, p->n_elements);
        if (m) {
          m = apply_rescue(p, n->mult, m);
          *cmap_slot(&r->counts, m) += 1;
          if (popcount64(m) == 1) o->n_rescued++;
        }
        n->has_open = 0; /* readCounts.erase, mm:1698 */
      }
    }
  } else {
    if (mask) {
      int output = 0;
      if (p->strategy == 2) {
        node *n = nmap_get(&r->names, key);
        if (!n) return -1;
        if (!n->seen) {
          if (!n->has_chosen) {
            uint32_t draw = next_rand(r); /* rand() is evaluated before the modulo (mm:1711); NH = 0 traps there in the reference */
            n->chosen = nh ? draw % nh : 0;
            n->has_chosen = 1;
            n->number_seen = 0;
          } else {
            n->number_seen++;
          }
          if (n->number_seen == n->chosen) {
            output = 1;
            n->has_chosen = 0;
            n->seen = 1;
          }
        }
      }
      if (p->strategy != 2 || output) {
        uint32_t mult[MAXE];
        memset(mult, 0, sizeof(mult));
        for (uint32_t i = 0; i < p->n_elements; ++i) if (mask >> i & 1) mult[i] = 1;
        uint64_t m = apply_rescue(p, mult, mask);
        *cmap_slot(&r->counts, m) += (p->strategy == 3) ? 1.0 / nh : 1; /* mm:1730 */
      }
    }
    o->n_reads++;
  }
  return 0;
}

static int cmp_u64(const void *a, const void *b) {
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return x < y ? -1 : x > y;
}

int orc_run(const orc_params *p, const orc_features *f, const orc_hits *h, int want_hit_masks, orc_result *out) {
  if (p->n_elements > MAXE) return -1;
  memset(out, 0, sizeof(*out));
  actx a;
  memset(&a, 0, sizeof(a));
  a.p = p; a.f = f;
  if (build_chr_start(f, &a.chr_start)) return -1;
  if (build_bins(&a)) { free_actx(&a); return -1; }
  rctx r;
  memset(&r, 0, sizeof(r));
  r.p = p; r.out = out;
  if (cmap_init(&r.counts) || nmap_init(&r.names, 1 << 16)) return -1;
  if (want_hit_masks) out->hit_mask = (uint64_t *)calloc(h->n ? h->n : 1, sizeof(uint64_t));
  /* Counter::read main loop, mm:1772-1781 */
  for (uint64_t i = 0; i < h->n; ++i) {
    uint32_t nh = h->nh[i];
    if (p->strategy == 1 && nh != 1) continue; /* unique, mm:1773 */
    out->n_hits++;
    uint64_t m = annotate_one(&a, h->start[i], h->end[i], h->meta[i]);
    if (out->hit_mask) out->hit_mask[i] = m;
    if (add_count(&r, h->read_key[i], m, nh)) return -1;
  }
  /* end-of-file flush of the still-open multi-mapping reads, mm:1783-1792 (only the default
   * strategy ever fills readCounts, so the unique/ratio sub-cases there are dead) */
  for (uint64_t b = 0; b < r.names.nbucket; ++b)
    for (node *n = r.names.bucket[b]; n; n = n->next)
      if (n->has_open) {
        uint64_t m = mult_mask(n->mult, p->n_elements);
        if (m) {
          m = apply_rescue(p, n->mult, m);
          *cmap_slot(&r.counts, m) += 1;
          if (n->raw > 1 && popcount64(m) == 1) out->n_rescued++;
        }
      }
  /* rows, sorted by mask */
  out->n_rows = r.counts.n;
  out->row_mask = (uint64_t *)malloc((out->n_rows ? out->n_rows : 1) * sizeof(uint64_t));
  out->row_value = (double *)malloc((out->n_rows ? out->n_rows : 1) * sizeof(double));
  uint64_t k = 0;
  for (uint64_t s = 0; s < r.counts.cap; ++s) if (r.counts.key[s]) out->row_mask[k++] = r.counts.key[s];
  qsort(out->row_mask, out->n_rows, sizeof(uint64_t), cmp_u64);
  for (k = 0; k < out->n_rows; ++k) out->row_value[k] = *cmap_slot(&r.counts, out->row_mask[k]);
  free(r.counts.key); free(r.counts.val);
  nmap_free(&r.names);
  free_actx(&a);
  return 0;
}

void orc_free(orc_result *r) {
  free(r->row_mask); free(r->row_value); free(r->hit_mask);
  memset(r, 0, sizeof(*r));
}
