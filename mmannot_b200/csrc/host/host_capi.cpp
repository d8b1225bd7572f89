// C view of the host front-end (include/mmannot_b200_host.h).
#include "mmannot_b200_host.h"

#include <cstring>
#include <string>

#include "annotation.hpp"
#include "config.hpp"
#include "synth.hpp"
#include "xam.hpp"

using namespace mmb;

struct mmh_config { Config config; };
struct mmh_annotation { FeatureTable table; std::string warnings; };
struct mmh_reader { XamReader *reader; };
struct mmh_synth { SynthGenome *genome; };

namespace {
thread_local std::string g_error;
size_t copyOut(const std::string &s, char *buf, size_t cap) {
  if (buf && cap) {
    size_t n = std::min(s.size(), cap - 1);
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return s.size();
}
}  // namespace

extern "C" {

const char *mmh_last_error(void) { return g_error.c_str(); }

int mmh_config_load(const char *path, mmh_config **out) {
  mmh_config *c = new mmh_config();
  std::string err;
  if (!c->config.parse(path, err)) { g_error = err; delete c; return -1; }
  *out = c;
  return 0;
}
void mmh_config_free(mmh_config *c) { delete c; }
uint32_t mmh_config_n_elements(const mmh_config *c) { return static_cast<uint32_t>(c->config.getNElements()); }
void mmh_config_tables(const mmh_config *c, uint16_t *line, uint8_t *strand, uint8_t *vicinity) {
  std::vector<uint16_t> l; std::vector<uint8_t> s, v;
  c->config.deviceTables(l, s, v);
  for (size_t i = 0; i < l.size(); ++i) { line[i] = l[i]; strand[i] = s[i]; vicinity[i] = v[i]; }
}
size_t mmh_config_name(const mmh_config *c, uint32_t element, char *buf, size_t cap) { return copyOut(c->config.getName(element), buf, cap); }
size_t mmh_config_order_echo(const mmh_config *c, char *buf, size_t cap) { return copyOut(c->config.orderEcho(), buf, cap); }

int mmh_annotation_build(const mmh_config *c, const char *gtf_path, uint64_t upstream, uint64_t downstream, mmh_annotation **out) {
  mmh_annotation *a = new mmh_annotation();
  AnnotationOptions opt;
  opt.upstreamSize = upstream;
  opt.downstreamSize = downstream;
  std::string err;
  if (!buildFeatureTable(gtf_path, c->config, opt, a->table, err, a->warnings)) { g_error = err; delete a; return -1; }
  *out = a;
  return 0;
}
void mmh_annotation_free(mmh_annotation *a) { delete a; }
uint32_t mmh_annotation_n(const mmh_annotation *a) { return static_cast<uint32_t>(a->table.size()); }
uint32_t mmh_annotation_n_chr(const mmh_annotation *a) { return static_cast<uint32_t>(a->table.chromosomes.size()); }
uint64_t mmh_annotation_n_genes(const mmh_annotation *a) { return a->table.nGenes; }
uint64_t mmh_annotation_n_lines(const mmh_annotation *a) { return a->table.nLines; }
const uint32_t *mmh_annotation_chr(const mmh_annotation *a) { return a->table.chr.data(); }
const uint32_t *mmh_annotation_start(const mmh_annotation *a) { return a->table.start.data(); }
const uint32_t *mmh_annotation_end(const mmh_annotation *a) { return a->table.end.data(); }
const uint8_t *mmh_annotation_type(const mmh_annotation *a) { return a->table.type.data(); }
const uint8_t *mmh_annotation_strand(const mmh_annotation *a) { return a->table.strand.data(); }
const char *mmh_annotation_id(const mmh_annotation *a, uint32_t i) { return a->table.id[i].c_str(); }
const char *mmh_annotation_chr_name(const mmh_annotation *a, uint32_t chr) { return a->table.chromosomes[chr].c_str(); }
const char *mmh_annotation_warnings(const mmh_annotation *a) { return a->warnings.c_str(); }

int mmh_reader_open(const mmh_annotation *a, const char *path, int format, char strandedness, mmh_reader **out) {
  ReadsFormat f = format == 1 ? ReadsFormat::SAM : format == 2 ? ReadsFormat::BAM : ReadsFormat::UNKNOWN;
  Strandedness s;
  if (strandedness == 'U') s = Strandedness::U;
  else if (strandedness == 'F') s = Strandedness::F;
  else if (strandedness == 'R') s = Strandedness::R;
  else { g_error = std::string("Do not understand strandedness ") + strandedness; return -1; }
  XamReader *r = new XamReader(path, f, s, a->table);
  std::string err;
  if (!r->open(err)) { g_error = err; delete r; return -1; }
  *out = new mmh_reader{r};
  return 0;
}
void mmh_reader_close(mmh_reader *r) {
  if (r) { delete r->reader; delete r; }
}
uint64_t mmh_reader_next(mmh_reader *r, uint64_t cap, uint32_t *start, uint32_t *end, uint32_t *meta, uint32_t *nh, uint64_t *read_key) {
  HitBuffers b;
  b.start = start; b.end = end; b.meta = meta; b.nh = nh; b.key = read_key; b.capacity = cap;
  return r->reader->nextBatch(b, nullptr);
}
uint64_t mmh_reader_records(const mmh_reader *r) { return r->reader->recordsRead(); }
size_t mmh_reader_warnings(mmh_reader *r, char *buf, size_t cap) { return copyOut(r->reader->takeWarnings(), buf, cap); }
size_t mmh_reader_key_collision(mmh_reader *r, char *buf, size_t cap) { return copyOut(r->reader->keyCollision(), buf, cap); }

uint64_t mmh_name_key(const char *name, size_t len) { return name_key(name, len); }

static SynthReadSpec toSpec(const mmh_synth_reads *r) {
  SynthReadSpec s;
  if (r) {
    s.maxNH = r->max_nh; s.paired = r->paired != 0; s.flipMate2 = r->flip_mate2 != 0; s.rnaSeq = r->rna_seq != 0;
    s.pInFeature = r->p_in_feature; s.pSameClass = r->p_same_class;
  }
  return s;
}
static Strandedness toStrand(char c) { return c == 'U' ? Strandedness::U : c == 'R' ? Strandedness::R : Strandedness::F; }

int mmh_synth_create(const char *shape, uint64_t seed, double gene_scale, mmh_synth **out) {
  SynthGenome *g = new SynthGenome(shape, seed, gene_scale);
  if (!g->ok()) { g_error = std::string("unknown synthetic shape '") + shape + "'"; delete g; return -1; }
  *out = new mmh_synth{g};
  return 0;
}
void mmh_synth_free(mmh_synth *s) {
  if (s) { delete s->genome; delete s; }
}
uint64_t mmh_synth_n_genes(const mmh_synth *s) { return s->genome->genes.size(); }
int mmh_synth_write_annotation(const mmh_synth *s, const char *path) { s->genome->writeAnnotation(path); return 0; }
int mmh_synth_write_bam(const mmh_synth *s, const char *path, uint64_t first_read, uint64_t n_reads, const mmh_synth_reads *spec, int coordinate_sorted) {
  if (!s->genome->writeBam(path, first_read, n_reads, toSpec(spec), (coordinate_sorted & 1) != 0, (coordinate_sorted & 2) != 0, (coordinate_sorted & 4) != 0)) { g_error = std::string("cannot write '") + path + "'"; return -1; }
  return 0;
}
uint64_t mmh_synth_count_hits(const mmh_synth *s, uint64_t first_read, uint64_t n_reads, const mmh_synth_reads *spec) {
  return s->genome->countHits(first_read, n_reads, toSpec(spec));
}
uint64_t mmh_synth_fill_hits(const mmh_synth *s, const mmh_annotation *a, char strandedness, uint64_t first_read, uint64_t n_reads,
                             const mmh_synth_reads *spec, uint64_t cap, uint32_t *start, uint32_t *end, uint32_t *meta, uint32_t *nh, uint64_t *read_key) {
  HitBuffers b;
  b.start = start; b.end = end; b.meta = meta; b.nh = nh; b.key = read_key; b.capacity = cap;
  return s->genome->fillHits(a->table, toStrand(strandedness), first_read, n_reads, toSpec(spec), b);
}

}  // extern "C"
