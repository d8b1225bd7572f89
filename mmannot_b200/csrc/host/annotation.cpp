#include "annotation.hpp"

#include <algorithm>
#include <fstream>
#include <limits>
#include <sstream>
#include <unordered_map>
#include <unordered_set>

namespace mmb {

namespace {

typedef uint64_t Pos;
const Pos UNSET = std::numeric_limits<Pos>::max();

struct Span {
  Pos s = UNSET, e = UNSET;
  Span() {}
  Span(Pos a, Pos b) : s(a), e(b) {}
  bool isSet() const { return s != UNSET && e != UNSET; }
  void cover(const Span &o) { s = std::min(s, o.s); e = std::max(e, o.e); }
};

// Ordered set of disjoint exon blocks + the span that encloses them (reference: Transcript, mm:708-829).
struct Blocks {
  Span span;
  std::vector<Span> blocks;

  // mm:719-739: append, order by start, fuse blocks that are not strictly before the next one
  void add(const Span &x) {
    span.cover(x);
    blocks.push_back(x);
    std::sort(blocks.begin(), blocks.end(), [](const Span &a, const Span &b) { return a.s < b.s; });
    std::vector<Span> fused;
    Span cur;
    for (const Span &b : blocks) {
      if (!cur.isSet()) cur = b;
      else if (cur.e < b.s) { fused.push_back(cur); cur = b; }
      else cur.cover(b);
    }
    fused.push_back(cur);
    blocks.swap(fused);
  }

  // mm:787-803: clip every block to `w`, dropping the ones that vanish
  void clip(const Span &w) {
    std::vector<Span> kept;
    for (Span b : blocks) {
      if (!b.isSet()) continue;
      b.s = std::max(b.s, w.s);
      b.e = std::min(b.e, w.e);
      if (b.s > b.e) continue;
      kept.push_back(b);
    }
    blocks.swap(kept);
    if (blocks.empty()) span = Span();
    else span = Span(blocks.front().s, blocks.back().e);
  }
};

struct GeneModel {
  Span span;
  uint8_t strand;  // 1 '+', 2 other
  uint32_t chr;
  std::string id, source, type;
  Blocks merged;   // union of exon + CDS lines
  Span cdsSpan;    // hull of the CDS lines (mm:926-930)
};

struct GtfLine {
  std::string chromosome, source, type;
  Pos start = 0, end = 0;
  uint8_t strand = 2;
  std::vector<std::pair<std::string, std::string> > tags;  // tag -> FIRST comma piece of its value; later duplicates win

  bool has(const char *t) const {
    for (const auto &p : tags) if (p.first == t) return true;
    return false;
  }
  const std::string &get(const char *t) const {
    static const std::string empty;
    const std::string *r = &empty;
    for (const auto &p : tags) if (p.first == t) r = &p.second;
    return *r;
  }
};

// Column 9 grammar of the reference (mm:533-567): `tag value;` or `tag=value;`, value quoted or not.
void parseAttributes(const std::string &col, GtfLine &out) {
  // cursor version of: rest = trimmed(col); repeatedly cut `tag`, `value`, skip to after the next ';' (no string erases)
  const char *p = col.data(), *e = p + col.size();
  while (p < e && is_space(*p)) ++p;
  while (e > p && is_space(e[-1])) --e;
  auto findc = [](const char *a, const char *z, char c) { while (a < z && *a != c) ++a; return a; };  // z when absent
  while (p < e) {
    const char *pSpace = findc(p, e, ' '), *pEq = findc(p, e, '=');
    const char *cut = std::min(pSpace, pEq);  // e when neither is present
    const char *tagEnd = cut;
    while (tagEnd > p && is_space(tagEnd[-1])) --tagEnd;
    const char *tagBegin = p;
    if (cut != e) p = cut + 1;  // (without a separator the tag is the whole rest, which is then read again as the value)
    while (p < e && is_space(*p)) ++p;
    const char *vb, *ve;
    if (p < e && *p == '"') {
      ++p;
      vb = p;
      ve = findc(p, e, '"');
      if (ve != e) p = ve + 1;
    } else {
      vb = p;
      ve = findc(p, e, ';');
      while (ve > vb && is_space(ve[-1])) --ve;
    }
    const char *comma = findc(vb, ve, ',');
    out.tags.emplace_back(std::string(tagBegin, tagEnd), std::string(vb, comma));
    const char *semi = findc(p, e, ';');
    if (semi == e) break;
    p = semi + 1;
    while (p < e && is_space(*p)) ++p;
  }
}

bool parseGtfLine(const std::string &line, GtfLine &out, std::string &err) {
  // getline()-style split on tabs into pieces that keep their storage from line to line (a trailing empty piece is dropped)
  static thread_local std::vector<std::string> f(12);
  size_t nf = 0;
  for (size_t pos = 0; pos < line.size();) {
    size_t q = line.find('\t', pos);
    if (q == std::string::npos) q = line.size();
    if (nf < f.size()) f[nf].assign(line, pos, q - pos);
    ++nf;
    pos = q + 1;
  }
  if (nf != 9) {
    err = "Error, annotation line does not have 9 tab-separated fields: '" + line + "'";
    return false;
  }
  bool okS, okE;
  out.chromosome = f[0];
  out.source = f[1];
  out.type = f[2];
  out.start = parse_ulong(f[3], okS);
  out.end = parse_ulong(f[4], okE);
  if (!okS || !okE) {
    err = "Error, cannot read the coordinates of annotation line '" + line + "'";
    return false;
  }
  out.strand = (f[6] == "+") ? 1 : 2;
  parseAttributes(f[8], out);
  return true;
}

// id precedence of the reference's Gene(GtfLineParser&) constructor, mm:918
std::string geneIdOf(const GtfLine &l) {
  if (l.has("gene_id")) return l.get("gene_id");
  if (l.has("ID")) return l.get("ID");
  if (l.has("transcript_id")) return l.get("transcript_id");
  const std::string &p = l.get("Parent");
  size_t dot = p.find('.');
  return dot == std::string::npos ? p : p.substr(0, dot);
}

GeneModel makeGene(const GtfLine &l, uint32_t chr) {
  GeneModel g;
  g.span = Span(l.start, l.end);
  g.strand = l.strand;
  g.chr = chr;
  g.id = geneIdOf(l);
  g.source = l.source;
  g.type = l.type;
  g.merged.span = g.span;
  return g;
}

void geneAddExon(GeneModel &g, const Span &e) {
  g.span.cover(e);
  g.merged.add(e);
}
void geneAddCds(GeneModel &g, const Span &c) {
  geneAddExon(g, c);
  if (g.cdsSpan.isSet()) g.cdsSpan.cover(c);
  else g.cdsSpan = c;
}

struct RawFeature {
  Pos s, e;
  uint32_t chr;
  uint8_t type, strand;
  std::string id;
};

struct SortKey {
  uint32_t chr;
  Pos start;
  uint32_t src;
};

}  // namespace

bool buildFeatureTable(const std::string &gtfFile, const Config &config, const AnnotationOptions &opt,
                       FeatureTable &out, std::string &err, std::string &warnings) {
  std::ifstream file(gtfFile.c_str());
  if (!file.good()) {
    err = "Error, Annotation file '" + gtfFile + "' does not exists!";
    return false;
  }
  if (config.getNElements() > 64) {
    err = "Error, the device path supports at most 64 elements in the 'Order' section.";
    return false;
  }
  out = FeatureTable();
  std::unordered_map<std::string, size_t> geneOf;   // id -> gene index, per chromosome block (mm:1100, mm:1113)
  std::unordered_set<std::string> unused;           // ids seen on lines that are not in Order (mm:1101, mm:1217-1221)
  std::vector<GeneModel> genes;
  std::string line, currentChr;
  bool haveChr = false;
  uint32_t chrId = 0;
  size_t lineNo = 0;
  std::ostringstream warn;
  for (; std::getline(file, line); ++lineNo) {
    if (line.empty() || line[0] == '#') continue;
    GtfLine l;
    if (!parseGtfLine(line, l, err)) return false;
    l.source = config.translate(l.source);
    l.type = config.translate(l.type);
    if (!haveChr || l.chromosome != currentChr) {
      geneOf.clear();
      unused.clear();
      currentChr = l.chromosome;
      haveChr = true;
      size_t k = 0;
      while (k < out.chromosomes.size() && out.chromosomes[k] != currentChr) ++k;
      if (k == out.chromosomes.size()) out.chromosomes.push_back(currentChr);
      chrId = static_cast<uint32_t>(k);
    }
    const Span here(l.start, l.end);
    if (l.type == "gene") {
      std::string gid;
      if (l.has("ID")) gid = l.get("ID");
      else if (l.has("gene_id")) gid = l.get("gene_id");
      else warn << "Warning, cannot deduce gene id at line " << lineNo << ": '" << line << "'.\n";
      geneOf[gid] = genes.size();
      genes.push_back(makeGene(l, chrId));
    } else if (l.type == "transcript") {
      std::string tid, gid;
      if (l.has("ID")) tid = l.get("ID");
      else if (l.has("transcript_id")) tid = l.get("transcript_id");
      else warn << "Warning, cannot deduce transcript id at line " << lineNo << ": '" << line << "'.\n";
      if (l.has("Parent")) gid = l.get("Parent");
      else if (l.has("gene_id")) gid = l.get("gene_id");
      else warn << "Warning, cannot deduce transcript parent id at line " << lineNo << ": '" << line << "'.\n";
      if (!unused.count(gid)) {
        auto it = geneOf.find(gid);
        if (it != geneOf.end()) { size_t g = it->second; geneOf[tid] = g; }
      }
    } else if (l.type == "exon") {
      std::string gid;
      if (l.has("Parent")) gid = l.get("Parent");
      else if (l.has("gene_id")) gid = l.get("gene_id");
      else if (l.has("transcript_id")) gid = l.get("transcript_id");
      else warn << "Warning, cannot deduce exon id at line " << lineNo << ": '" << line << "'.\n";
      if (!unused.count(gid)) {
        auto it = geneOf.find(gid);
        if (it == geneOf.end()) {
          GeneModel g = makeGene(l, chrId);
          geneAddExon(g, here);
          geneOf[gid] = genes.size();
          genes.push_back(g);
        } else {
          geneAddExon(genes[it->second], here);
        }
      }
    } else if (l.type == "CDS") {
      std::string gid;
      if (l.has("gene_id")) gid = l.get("gene_id");
      else if (l.has("Parent")) gid = l.get("Parent");
      else if (l.has("transcript_id")) gid = l.get("transcript_id");
      else warn << "Warning, cannot deduce CDS parent id at line " << lineNo << ": '" << line << "'.\n";
      auto it = geneOf.find(gid);
      if (it == geneOf.end()) {
        GeneModel g = makeGene(l, chrId);
        geneAddCds(g, here);
        geneOf[gid] = genes.size();
        genes.push_back(g);
      } else {
        geneAddCds(genes[it->second], here);
      }
    } else if (l.type == "5'UTR" || l.type == "3'UTR") {
      // UTRs are derived from CDS and exons, the lines themselves are ignored (mm:1197-1202)
    } else if (config.getOrder(l.source, l.type) != NO_ID) {
      std::string fid;
      if (l.has("ID")) fid = l.get("ID");
      else if (l.has("gene_id")) fid = l.get("gene_id");
      else if (l.has("transcript_id")) fid = l.get("transcript_id");
      else if (l.has("Parent")) fid = l.get("Parent") + "_" + l.type;
      else warn << "Warning, cannot deduce id at line " << lineNo << ": '" << line << "'.\n";
      geneOf[fid] = genes.size();
      genes.push_back(makeGene(l, chrId));
    } else {
      if (l.has("gene_id")) unused.insert(l.get("gene_id"));
      if (l.has("transcript_id")) unused.insert(l.get("transcript_id"));
      if (l.has("ID")) unused.insert(l.get("ID"));
    }
  }
  out.nLines = lineNo;
  out.nGenes = genes.size();

  // Gene model -> typed intervals, in the reference's emission order (mm:1227-1266).
  std::vector<RawFeature> raw;
  auto emit = [&raw](const Span &s, size_t type, const GeneModel &g, const std::string &id) {
    raw.push_back(RawFeature{s.s, s.e, g.chr, static_cast<uint8_t>(type), g.strand, id});
  };
  for (GeneModel &g : genes) {
    // structure (mm:743-751, 955-962)
    if (g.merged.blocks.empty()) g.merged.blocks.push_back(g.merged.span);
    std::vector<Span> introns;
    for (size_t k = 1; k < g.merged.blocks.size(); ++k)
      introns.push_back(Span(g.merged.blocks[k - 1].e + 1, g.merged.blocks[k].s - 1));
    g.span = g.merged.span;
    Blocks cds, utr5, utr3;
    if (g.cdsSpan.isSet()) {
      cds = g.merged;
      cds.clip(g.cdsSpan);
      if (cds.span.isSet()) {
        utr5 = g.merged;
        utr3 = g.merged;
        utr5.clip(Span(g.span.s, cds.span.s - 1));
        utr3.clip(Span(cds.span.e + 1, g.span.e));
        if (g.strand == 2) std::swap(utr5, utr3);
      }
    }
    Span up, down;
    if (g.strand == 1) {
      up = Span(g.span.s <= opt.upstreamSize ? 1 : g.span.s - opt.upstreamSize, g.span.s - 1);
      down = Span(g.span.e + 1, g.span.e + opt.downstreamSize);
    } else {
      down = Span(g.span.s <= opt.downstreamSize ? 1 : g.span.s - opt.downstreamSize, g.span.s - 1);
      up = Span(g.span.e + 1, g.span.e + opt.upstreamSize);
    }
    size_t t;
    if ((t = config.getOrder(g.source, "CDS")) != NO_ID)
      for (const Span &b : cds.blocks) emit(b, t, g, g.id + "-CDS");
    if ((t = config.getOrder(g.source, "5'UTR")) != NO_ID)
      for (const Span &b : utr5.blocks) emit(b, t, g, g.id + "-5UTR");
    if ((t = config.getOrder(g.source, "3'UTR")) != NO_ID)
      for (const Span &b : utr3.blocks) emit(b, t, g, g.id + "-3UTR");
    if ((t = config.checkIntrons(g.source, g.type)) != NO_ID)
      for (const Span &b : introns) emit(b, t, g, g.id + "-intron");
    if ((t = config.checkUpstream(g.source, g.type)) != NO_ID) emit(up, t, g, g.id + "-upstream");
    if ((t = config.checkDownstream(g.source, g.type)) != NO_ID) emit(down, t, g, g.id + "-downstream");
    if ((t = config.getOrder(g.source, g.type)) != NO_ID)
      for (const Span &b : g.merged.blocks) emit(b, t, g, g.id);
  }

  // Same algorithm (std::sort), same comparator (chr, start) and same initial order as
  // mm:1267 => same permutation, including the order of equal-start intervals, which is
  // observable through "last matching interval wins" (mm:1023-1028).
  std::vector<SortKey> keys(raw.size());
  for (size_t i = 0; i < raw.size(); ++i) keys[i] = SortKey{raw[i].chr, raw[i].s, static_cast<uint32_t>(i)};
  std::sort(keys.begin(), keys.end(), [](const SortKey &a, const SortKey &b) {
    return (a.chr < b.chr) || ((a.chr == b.chr) && (a.start < b.start));
  });

  if (raw.empty()) {
    err = "Error, the annotation file has not been parsed properly!\nPlease check that your annotation file is not empty, and that your configuration file matches your annotation file.\nIf you have trouble designing a configuration file, please use the companion tool 'createConfigFile'.";
    warnings = warn.str();
    return false;
  }
  const Pos LIMIT = 0xFFFFFFFEull;
  out.chrHasFeatures.assign(out.chromosomes.size(), 0);
  for (const SortKey &k : keys) {
    const RawFeature &r = raw[k.src];
    if (r.s > LIMIT || r.e > LIMIT) {
      err = "Error, annotation coordinates beyond 4294967294 are not supported by the device path (interval '" + r.id + "').";
      return false;
    }
    out.chr.push_back(r.chr);
    out.start.push_back(static_cast<uint32_t>(r.s));
    out.end.push_back(static_cast<uint32_t>(r.e));
    out.type.push_back(r.type);
    out.strand.push_back(r.strand);
    out.id.push_back(r.id);
    out.chrHasFeatures[r.chr] = 1;
  }
  warnings = warn.str();
  return true;
}

}  // namespace mmb
