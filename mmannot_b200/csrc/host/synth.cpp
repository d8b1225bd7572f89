#include "synth.hpp"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

namespace mmb {

namespace {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
  uint64_t below(uint64_t n) { return n ? next() % n : 0; }
  uint64_t range(uint64_t lo, uint64_t hi) { return lo + below(hi - lo + 1); }  // inclusive
};

uint64_t mixSeed(uint64_t a, uint64_t b) {
  Rng r(a ^ (b * 0xD1342543DE82EF95ull) ^ 0x2545F4914F6CDD1Dull);
  r.next();
  return r.next();
}

// ---- BGZF / BAM writing -------------------------------------------------------------------

class BgzfWriter {
 public:
  explicit BgzfWriter(const std::string &path) : f_(std::fopen(path.c_str(), "wb")) { buf_.reserve(0xff00); }
  bool ok() const { return f_ != nullptr; }
  void write(const void *p, size_t n) {
    const unsigned char *c = static_cast<const unsigned char *>(p);
    while (n) {
      size_t take = std::min(n, static_cast<size_t>(0xff00) - buf_.size());
      buf_.insert(buf_.end(), c, c + take);
      c += take; n -= take;
      if (buf_.size() == 0xff00) flush();
    }
  }
  // htslib never lets an alignment record straddle two members (bgzf_flush_try before every record): the same here, unless
  // the caller wants a file that does (tests of the decoders' handling of it)
  void writeRecord(const void *p, size_t n, bool mayStraddle) {
    if (!mayStraddle && n <= 0xff00 && buf_.size() + n > 0xff00) flush();
    write(p, n);
  }
  void endOfHeader() { flush(); }
  void close() {
    if (!f_) return;
    flush();
    static const unsigned char eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    std::fwrite(eof, 1, 28, f_);
    std::fclose(f_);
    f_ = nullptr;
  }
  ~BgzfWriter() { close(); }

 private:
  void flush() {
    if (buf_.empty()) return;
    unsigned char out[0x10000 + 64];
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    deflateInit2(&zs, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = buf_.data(); zs.avail_in = static_cast<uInt>(buf_.size());
    zs.next_out = out + 18; zs.avail_out = sizeof(out) - 18 - 8;
    deflate(&zs, Z_FINISH);
    size_t clen = zs.total_out;
    deflateEnd(&zs);
    const unsigned char head[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    std::memcpy(out, head, 16);
    uint32_t bsize = static_cast<uint32_t>(clen + 18 + 8 - 1);
    out[16] = bsize & 0xff; out[17] = (bsize >> 8) & 0xff;
    uint32_t crc = static_cast<uint32_t>(crc32(crc32(0L, Z_NULL, 0), buf_.data(), static_cast<uInt>(buf_.size())));
    uint32_t isize = static_cast<uint32_t>(buf_.size());
    unsigned char *t = out + 18 + clen;
    for (int i = 0; i < 4; ++i) { t[i] = (crc >> (8 * i)) & 0xff; t[4 + i] = (isize >> (8 * i)) & 0xff; }
    std::fwrite(out, 1, clen + 26, f_);
    buf_.clear();
  }
  std::FILE *f_;
  std::vector<unsigned char> buf_;
};

void put32(std::vector<unsigned char> &v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back((x >> (8 * i)) & 0xff); }

int reg2bin(int64_t beg, int64_t end) {  // SAM spec 5.3
  --end;
  if (beg >> 14 == end >> 14) return static_cast<int>(((1 << 15) - 1) / 7 + (beg >> 14));
  if (beg >> 17 == end >> 17) return static_cast<int>(((1 << 12) - 1) / 7 + (beg >> 17));
  if (beg >> 20 == end >> 20) return static_cast<int>(((1 << 9) - 1) / 7 + (beg >> 20));
  if (beg >> 23 == end >> 23) return static_cast<int>(((1 << 6) - 1) / 7 + (beg >> 23));
  if (beg >> 26 == end >> 26) return static_cast<int>(((1 << 3) - 1) / 7 + (beg >> 26));
  return 0;
}

void bamRecord(std::vector<unsigned char> &v, const std::string &name, const SynthRecord &r) {
  v.clear();
  const uint32_t lName = static_cast<uint32_t>(name.size()) + 1;
  const int64_t pos0 = static_cast<int64_t>(r.pos) - 1;
  put32(v, 0);  // block_size, patched below
  put32(v, r.chr);
  put32(v, static_cast<uint32_t>(pos0));
  put32(v, (static_cast<uint32_t>(reg2bin(pos0, pos0 + r.len)) << 16) | (255u << 8) | lName);
  const uint32_t flag = r.flag | (r.forward ? 0u : 0x10u);
  put32(v, (flag << 16) | 1u);
  put32(v, r.len);
  put32(v, 0xFFFFFFFFu); put32(v, 0xFFFFFFFFu); put32(v, 0);
  v.insert(v.end(), name.begin(), name.end());
  v.push_back(0);
  put32(v, (r.len << 4) | 0u);  // <len>M
  for (uint32_t i = 0; i < (r.len + 1) / 2; ++i) v.push_back(0x12);  // "AC" pairs
  for (uint32_t i = 0; i < r.len; ++i) v.push_back(0xff);
  v.push_back('N'); v.push_back('M'); v.push_back('C'); v.push_back(0);
  v.push_back('N'); v.push_back('H');
  if (r.nh < 256) { v.push_back('C'); v.push_back(static_cast<unsigned char>(r.nh)); }
  else { v.push_back('S'); v.push_back(r.nh & 0xff); v.push_back((r.nh >> 8) & 0xff); }
  const uint32_t bs = static_cast<uint32_t>(v.size()) - 4;
  for (int i = 0; i < 4; ++i) v[i] = (bs >> (8 * i)) & 0xff;
}

}  // namespace

// ---- genome -------------------------------------------------------------------------------

SynthGenome::SynthGenome(const std::string &shape, uint64_t seed, double geneScale) : shape_(shape), seed_(seed) {
  uint32_t nGenes = 0;
  double meanExons = 5, intronMean = 200, nestedFrac = 0.05;
  uint32_t intronMax = 3000;
  if (shape == "tair10") {
    gff3_ = true; source_ = "TAIR10"; nGenes = 33000; meanExons = 5; intronMean = 160; intronMax = 3000;
    const char *n[] = {"Chr1", "Chr2", "Chr3", "Chr4", "Chr5", "ChrC", "ChrM"};
    const uint64_t l[] = {30427671, 19698289, 23459830, 18585056, 26975502, 154478, 366924};
    for (int i = 0; i < 7; ++i) { chrNames.push_back(n[i]); chrLen.push_back(l[i]); }
    classes_ = {{"mRNA", 27000, true, true, 300, 6000, true}, {"miRNA", 330, false, false, 80, 400, true}, {"tRNA", 690, false, false, 70, 90, true},
                {"snoRNA", 71, false, false, 80, 250, true}, {"snRNA", 13, false, false, 100, 200, true}, {"rRNA", 15, false, false, 120, 3500, true},
                {"transposable_element_gene", 3900, false, false, 500, 6000, true}, {"pseudogenic_transcript", 920, false, true, 300, 3000, true}};
  } else if (shape == "hs38") {
    gff3_ = false; source_ = ""; nGenes = 60000; meanExons = 9; intronMean = 5000; intronMax = 200000; nestedFrac = 0.12;
    const uint64_t l[] = {248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422, 135086622,
                          133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468};
    for (int i = 0; i < 22; ++i) { chrNames.push_back(std::to_string(i + 1)); chrLen.push_back(l[i]); }
    chrNames.push_back("X"); chrLen.push_back(156040895);
    chrNames.push_back("Y"); chrLen.push_back(57227415);
    chrNames.push_back("MT"); chrLen.push_back(16569);
    Rng r(mixSeed(seed, 77));
    for (int i = 0; i < 169; ++i) { chrNames.push_back("KI270" + std::to_string(302 + i) + ".1"); chrLen.push_back(r.range(20000, 200000)); }
    classes_ = {{"protein_coding", 20000, true, true, 1000, 100000, true}, {"lincRNA", 7500, false, true, 500, 50000, true},
                {"miRNA", 3000, false, false, 60, 140, true}, {"snRNA", 1900, false, false, 90, 200, true}, {"snoRNA", 1450, false, false, 70, 250, true},
                {"misc_RNA", 2000, false, false, 90, 350, true}, {"rRNA", 520, false, false, 100, 160, true},
                {"processed_pseudogene", 10000, false, false, 200, 2500, true}, {"unprocessed_pseudogene", 2600, false, true, 500, 30000, true},
                {"pseudogene", 500, false, false, 200, 2000, true}, {"processed_transcript", 500, false, true, 500, 30000, true},
                {"antisense", 5000, false, true, 500, 30000, false}, {"sense_intronic", 700, false, false, 300, 5000, false},
                {"transcribed_processed_pseudogene", 400, false, false, 300, 3000, true}, {"polymorphic_pseudogene", 50, true, true, 1000, 20000, true}};
  } else if (shape == "flybase6") {
    gff3_ = true; source_ = "FlyBase"; nGenes = 17500; meanExons = 4; intronMean = 900; intronMax = 30000; nestedFrac = 0.10;
    const char *n[] = {"2L", "2R", "3L", "3R", "4", "X", "Y", "mitochondrion_genome"};
    const uint64_t l[] = {23513712, 25286936, 28110227, 32079331, 1348131, 23542271, 3667352, 19524};
    for (int i = 0; i < 8; ++i) { chrNames.push_back(n[i]); chrLen.push_back(l[i]); }
    Rng r(mixSeed(seed, 78));
    for (int i = 0; i < 24; ++i) { chrNames.push_back("211000022" + std::to_string(278000 + i * 37)); chrLen.push_back(r.range(5000, 90000)); }
    classes_ = {{"mRNA", 13900, true, true, 400, 40000, true}, {"miRNA", 260, false, false, 20, 30, true}, {"tRNA", 310, false, false, 70, 90, true},
                {"snoRNA", 290, false, false, 70, 300, true}, {"snRNA", 30, false, false, 100, 200, true}, {"rRNA", 115, false, false, 120, 2000, true},
                {"ncRNA", 2400, false, true, 300, 10000, true}, {"pre_miRNA", 260, false, false, 60, 120, true}, {"pseudogene", 330, false, true, 300, 4000, true}};
  } else {
    return;
  }
  nGenes = std::max<uint32_t>(static_cast<uint32_t>(nGenes * geneScale), 8);
  chrCum_.assign(chrLen.size() + 1, 0);
  for (size_t c = 0; c < chrLen.size(); ++c) chrCum_[c + 1] = chrCum_[c] + chrLen[c];
  genomeLen_ = chrCum_.back();
  double wsum = 0;
  for (const GeneClass &g : classes_) wsum += g.weight;
  genesOfClass_.assign(classes_.size(), std::vector<uint32_t>());
  Rng rng(mixSeed(seed, 1));
  uint64_t serial = 0;
  for (size_t c = 0; c < chrLen.size(); ++c) {
    uint64_t quota = static_cast<uint64_t>(std::llround(static_cast<double>(nGenes) * chrLen[c] / genomeLen_));
    if (quota == 0 && chrLen[c] >= 15000) quota = 1;
    if (quota == 0) continue;
    const uint64_t pitch = std::max<uint64_t>(chrLen[c] / (quota + 1), 50);
    uint64_t prevStart = 0, prevEnd = 0;
    for (uint64_t k = 0; k < quota; ++k) {
      SynthGene g;
      g.chr = static_cast<uint32_t>(c);
      double x = rng.uniform() * wsum;
      uint32_t cls = 0;
      while (cls + 1 < classes_.size() && x >= classes_[cls].weight) { x -= classes_[cls].weight; ++cls; }
      const GeneClass &gc = classes_[cls];
      g.cls = cls;
      g.forward = (rng.next() & 1) != 0;
      uint64_t start = 1 + k * pitch + rng.below(pitch / 2 + 1);
      if (k > 0 && rng.uniform() < nestedFrac && prevEnd > prevStart + 200) start = prevStart + rng.below((prevEnd - prevStart) / 2);  // nested / overlapping
      // exon chain
      uint32_t nEx = 1;
      if (gc.spliced) { while (nEx < 40 && rng.uniform() < 1.0 - 1.0 / meanExons) ++nEx; }
      uint64_t cursor = start;
      const uint64_t budget = std::max<uint64_t>(gc.minLen, std::min<uint64_t>(gc.maxLen, pitch * 3));
      for (uint32_t e = 0; e < nEx; ++e) {
        uint64_t el = gc.spliced ? rng.range(40, 400) : rng.range(gc.minLen, std::max(gc.minLen, std::min<uint32_t>(gc.maxLen, static_cast<uint32_t>(budget))));
        if (e + 1 == nEx && gc.spliced) el += rng.range(50, 600);
        g.exons.push_back(std::make_pair(cursor, cursor + el - 1));
        cursor += el;
        if (e + 1 < nEx) {
          double u = rng.uniform();
          uint64_t il = static_cast<uint64_t>(-std::log(1.0 - u * 0.999) * intronMean) + 20;
          il = std::min<uint64_t>(il, intronMax);
          cursor += il;
        }
        if (cursor - start > budget && e + 1 < nEx) { nEx = e + 2; }
      }
      g.start = start;
      g.end = g.exons.back().second;
      if (g.end + 2000 >= chrLen[c]) break;
      if (gc.coding) {
        const auto &fe = g.exons.front();
        const auto &le = g.exons.back();
        uint64_t a = fe.first + rng.below(std::max<uint64_t>((fe.second - fe.first) / 2, 1));
        uint64_t b = le.second - rng.below(std::max<uint64_t>((le.second - le.first) / 2, 1));
        if (a < b) { g.cdsStart = a; g.cdsEnd = b; }
      }
      g.nTranscripts = (gc.spliced && g.exons.size() >= 3) ? static_cast<uint32_t>(1 + rng.below(3)) : 1;
      g.serial = ++serial;
      prevStart = g.start; prevEnd = g.end;
      genesOfClass_[cls].push_back(static_cast<uint32_t>(genes.size()));
      genes.push_back(g);
    }
  }
  ok_ = !genes.empty();
}

void SynthGenome::writeAnnotation(const std::string &path) const {
  std::FILE *f = std::fopen(path.c_str(), "w");
  if (!f) return;
  std::fprintf(f, "##synthetic %s-shaped annotation, seed %llu\n", shape_.c_str(), static_cast<unsigned long long>(seed_));
  // genes are generated chromosome by chromosome; order them by start inside each chromosome
  std::vector<uint32_t> order(genes.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = static_cast<uint32_t>(i);
  std::stable_sort(order.begin(), order.end(), [this](uint32_t a, uint32_t b) {
    return genes[a].chr != genes[b].chr ? genes[a].chr < genes[b].chr : genes[a].start < genes[b].start;
  });
  char gid[64], tid[80];
  for (uint32_t gi : order) {
    const SynthGene &g = genes[gi];
    const GeneClass &gc = classes_[g.cls];
    const char *chr = chrNames[g.chr].c_str();
    const char sd = g.forward ? '+' : '-';
    const unsigned long long S = g.start, E = g.end;
    if (gff3_) {
      const bool tair = shape_ == "tair10";
      if (tair) std::snprintf(gid, sizeof(gid), "AT%uG%05llu", g.chr + 1, static_cast<unsigned long long>(g.serial));
      else std::snprintf(gid, sizeof(gid), "FBgn%07llu", static_cast<unsigned long long>(g.serial));
      const std::string &child = gc.biotype;
      const char *src = source_.c_str();
      if (tair && child == "transposable_element_gene") {
        std::fprintf(f, "%s\t%s\ttransposable_element_gene\t%llu\t%llu\t.\t%c\t.\tID=%s;Note=transposable_element_gene;Name=%s\n", chr, src, S, E, sd, gid, gid);
      } else if (tair && child == "pseudogenic_transcript") {
        std::fprintf(f, "%s\t%s\tpseudogene\t%llu\t%llu\t.\t%c\t.\tID=%s;Note=pseudogene;Name=%s\n", chr, src, S, E, sd, gid, gid);
      } else {
        std::fprintf(f, "%s\t%s\tgene\t%llu\t%llu\t.\t%c\t.\tID=%s;Note=%s;Name=%s\n", chr, src, S, E, sd, gid,
                     gc.coding ? "protein_coding_gene" : child.c_str(), gid);
      }
      for (uint32_t t = 0; t < g.nTranscripts; ++t) {
        if (tair) std::snprintf(tid, sizeof(tid), "%s.%u", gid, t + 1);
        else std::snprintf(tid, sizeof(tid), "FBtr%07llu", static_cast<unsigned long long>(g.serial * 4 + t));
        const char *tType = (tair && child == "transposable_element_gene") ? "mRNA" : child.c_str();
        std::fprintf(f, "%s\t%s\t%s\t%llu\t%llu\t.\t%c\t.\tID=%s;Parent=%s;Name=%s\n", chr, src, tType, S, E, sd, tid, gid, tid);
        if (tair && gc.coding && g.cdsStart)
          std::fprintf(f, "%s\t%s\tprotein\t%llu\t%llu\t.\t%c\t.\tID=%s-Protein;Name=%s;Derives_from=%s\n", chr, src,
                       static_cast<unsigned long long>(g.cdsStart), static_cast<unsigned long long>(g.cdsEnd), sd, tid, tid, tid);
        const char *exonType = (tair && child == "pseudogenic_transcript") ? "pseudogenic_exon" : "exon";
        for (size_t e = 0; e < g.exons.size(); ++e) {
          if (t > 0 && e > 0 && e + 1 < g.exons.size() && (e % (t + 1)) == 1) continue;  // alternative transcripts skip some internal exons
          const unsigned long long es = g.exons[e].first, ee = g.exons[e].second;
          std::fprintf(f, "%s\t%s\t%s\t%llu\t%llu\t.\t%c\t.\tParent=%s\n", chr, src, exonType, es, ee, sd, tid);
          if (gc.coding && g.cdsStart) {
            const char *u5 = tair ? "five_prime_UTR" : "5UTR", *u3 = tair ? "three_prime_UTR" : "3UTR";
            if (es < g.cdsStart)
              std::fprintf(f, "%s\t%s\t%s\t%llu\t%llu\t.\t%c\t.\tParent=%s\n", chr, src, g.forward ? u5 : u3, es,
                           std::min<unsigned long long>(ee, g.cdsStart - 1), sd, tid);
            const unsigned long long cs = std::max<unsigned long long>(es, g.cdsStart), ce = std::min<unsigned long long>(ee, g.cdsEnd);
            if (cs <= ce) {
              if (tair) std::fprintf(f, "%s\t%s\tCDS\t%llu\t%llu\t.\t%c\t0\tParent=%s,%s-Protein;\n", chr, src, cs, ce, sd, tid, tid);
              else std::fprintf(f, "%s\t%s\tCDS\t%llu\t%llu\t.\t%c\t0\tParent=%s\n", chr, src, cs, ce, sd, tid);
            }
            if (ee > g.cdsEnd)
              std::fprintf(f, "%s\t%s\t%s\t%llu\t%llu\t.\t%c\t.\tParent=%s\n", chr, src, g.forward ? u3 : u5,
                           std::max<unsigned long long>(es, g.cdsEnd + 1), ee, sd, tid);
          }
        }
      }
    } else {  // Ensembl-76 style GTF: column 2 carries the biotype
      std::snprintf(gid, sizeof(gid), "ENSG%011llu", static_cast<unsigned long long>(g.serial));
      const char *bt = gc.biotype.c_str();
      std::fprintf(f, "%s\t%s\tgene\t%llu\t%llu\t.\t%c\t.\tgene_id \"%s\"; gene_name \"G%llu\"; gene_source \"ensembl\"; gene_biotype \"%s\";\n", chr, bt, S, E, sd,
                   gid, static_cast<unsigned long long>(g.serial), bt);
      for (uint32_t t = 0; t < g.nTranscripts; ++t) {
        std::snprintf(tid, sizeof(tid), "ENST%011llu", static_cast<unsigned long long>(g.serial * 4 + t));
        // secondary transcripts of coding genes carry another biotype in column 2, like Ensembl does
        const char *tbt = (t > 0 && gc.coding) ? (t == 1 ? "retained_intron" : "processed_transcript") : bt;
        std::fprintf(f, "%s\t%s\ttranscript\t%llu\t%llu\t.\t%c\t.\tgene_id \"%s\"; transcript_id \"%s\"; gene_name \"G%llu\"; gene_biotype \"%s\";\n", chr, tbt,
                     S, E, sd, gid, tid, static_cast<unsigned long long>(g.serial), bt);
        const bool codingTx = gc.coding && g.cdsStart && t == 0;
        uint32_t exonNo = 0;
        for (size_t e = 0; e < g.exons.size(); ++e) {
          if (t > 0 && e > 0 && e + 1 < g.exons.size() && (e % (t + 1)) == 1) continue;
          ++exonNo;
          const unsigned long long es = g.exons[e].first, ee = g.exons[e].second;
          std::fprintf(f, "%s\t%s\texon\t%llu\t%llu\t.\t%c\t.\tgene_id \"%s\"; transcript_id \"%s\"; exon_number \"%u\"; gene_biotype \"%s\"; exon_id \"ENSE%011llu\";\n",
                       chr, tbt, es, ee, sd, gid, tid, exonNo, bt, static_cast<unsigned long long>(g.serial * 64 + e));
          if (codingTx) {
            const unsigned long long cs = std::max<unsigned long long>(es, g.cdsStart), ce = std::min<unsigned long long>(ee, g.cdsEnd);
            if (cs <= ce) {
              std::fprintf(f, "%s\t%s\tCDS\t%llu\t%llu\t.\t%c\t0\tgene_id \"%s\"; transcript_id \"%s\"; exon_number \"%u\"; protein_id \"ENSP%011llu\";\n", chr, tbt, cs,
                           ce, sd, gid, tid, exonNo, static_cast<unsigned long long>(g.serial));
              if (cs == g.cdsStart)
                std::fprintf(f, "%s\t%s\t%s\t%llu\t%llu\t.\t%c\t0\tgene_id \"%s\"; transcript_id \"%s\"; exon_number \"%u\";\n", chr, tbt,
                             g.forward ? "start_codon" : "stop_codon", cs, std::min(cs + 2, ce), sd, gid, tid, exonNo);
              if (ce == g.cdsEnd)
                std::fprintf(f, "%s\t%s\t%s\t%llu\t%llu\t.\t%c\t0\tgene_id \"%s\"; transcript_id \"%s\"; exon_number \"%u\";\n", chr, tbt,
                             g.forward ? "stop_codon" : "start_codon", std::max(ce >= 2 ? ce - 2 : cs, cs), ce, sd, gid, tid, exonNo);
            }
          }
        }
        if (codingTx) {
          if (g.exons.front().first < g.cdsStart)
            std::fprintf(f, "%s\t%s\tUTR\t%llu\t%llu\t.\t%c\t.\tgene_id \"%s\"; transcript_id \"%s\";\n", chr, tbt,
                         static_cast<unsigned long long>(g.exons.front().first), static_cast<unsigned long long>(g.cdsStart - 1), sd, gid, tid);
          if (g.exons.back().second > g.cdsEnd)
            std::fprintf(f, "%s\t%s\tUTR\t%llu\t%llu\t.\t%c\t.\tgene_id \"%s\"; transcript_id \"%s\";\n", chr, tbt,
                         static_cast<unsigned long long>(g.cdsEnd + 1), static_cast<unsigned long long>(g.exons.back().second), sd, gid, tid);
        }
      }
    }
  }
  std::fclose(f);
}

// ---- reads --------------------------------------------------------------------------------

void SynthGenome::readRecords(uint64_t r, const SynthReadSpec &spec, std::string &name, std::vector<SynthRecord> &out) const {
  out.clear();
  Rng rng(mixSeed(seed_ ^ 0xABCDEF12345ull, r));
  char nm[48];
  std::snprintf(nm, sizeof(nm), "sy%llu.%llu", static_cast<unsigned long long>(seed_ % 1000), static_cast<unsigned long long>(r));
  name = nm;
  uint32_t len;
  if (spec.rnaSeq) len = static_cast<uint32_t>(rng.range(50, 150));
  else {
    const double u = rng.uniform();
    if (u < 0.35) len = 21;
    else if (u < 0.65) len = 24;
    else len = static_cast<uint32_t>(rng.range(18, 30));
  }
  uint32_t nh = 1;
  if (rng.uniform() >= 0.5 && spec.maxNH > 1) {
    nh = 2;
    if (spec.maxNH > 20 && rng.uniform() < 0.1) nh = static_cast<uint32_t>(rng.range(2, spec.maxNH));
    else while (nh < spec.maxNH && rng.uniform() < 0.55) ++nh;
  }
  const bool sameClass = nh > 1 && rng.uniform() < spec.pSameClass;
  uint32_t cls = 0;
  if (sameClass) {
    do { cls = static_cast<uint32_t>(rng.below(classes_.size())); } while (genesOfClass_[cls].empty());
  }
  for (uint32_t k = 0; k < nh; ++k) {
    SynthRecord rec;
    rec.len = len; rec.nh = nh; rec.flag = 0;
    rec.forward = (rng.next() & 1) != 0;
    const SynthGene *g = nullptr;
    if (sameClass) g = &genes[genesOfClass_[cls][rng.below(genesOfClass_[cls].size())]];
    else if (rng.uniform() < spec.pInFeature) g = &genes[rng.below(genes.size())];
    if (g) {
      rec.chr = g->chr;
      uint64_t lo = g->start, hi = g->end;
      if (g->exons.size() > 1 && rng.uniform() < 0.7) { const auto &e = g->exons[rng.below(g->exons.size())]; lo = e.first; hi = e.second; }
      const uint64_t jitter = rng.below(8);  // a few placements straddle the borders
      rec.pos = (hi >= lo + len) ? rng.range(lo, hi - len + 1) : lo;
      if (jitter == 0 && rec.pos > len) rec.pos -= rng.below(len);
    } else {
      const uint64_t x = rng.below(genomeLen_);
      size_t c = std::upper_bound(chrCum_.begin(), chrCum_.end(), x) - chrCum_.begin() - 1;
      rec.chr = static_cast<uint32_t>(c);
      rec.pos = 1 + (x - chrCum_[c]);
      if (rec.pos + len > chrLen[c]) rec.pos = chrLen[c] > len ? chrLen[c] - len : 1;
    }
    if (!spec.paired) {
      out.push_back(rec);
    } else {
      SynthRecord m1 = rec, m2 = rec;
      m1.flag = 0x1 | 0x2 | 0x40 | (rec.forward ? 0x20 : 0);
      m2.flag = 0x1 | 0x2 | 0x80 | (rec.forward ? 0 : 0x20);
      m2.forward = spec.flipMate2 ? rec.forward : !rec.forward;
      m2.pos = std::min<uint64_t>(rec.pos + rng.range(100, 400), chrLen[rec.chr] > len ? chrLen[rec.chr] - len : 1);
      out.push_back(m1);
      out.push_back(m2);
    }
  }
}

uint64_t SynthGenome::countHits(uint64_t first, uint64_t nReads, const SynthReadSpec &spec) const {
  std::string name;
  std::vector<SynthRecord> recs;
  uint64_t n = 0;
  for (uint64_t r = first; r < first + nReads; ++r) { readRecords(r, spec, name, recs); n += recs.size(); }
  return n;
}

uint64_t SynthGenome::fillHits(const FeatureTable &features, Strandedness s, uint64_t first, uint64_t nReads, const SynthReadSpec &spec,
                               const HitBuffers &dst) const {
  std::vector<uint32_t> chrMap(chrNames.size(), HIT_CHR_NONE);
  for (size_t c = 0; c < chrNames.size(); ++c)
    for (size_t k = 0; k < features.chromosomes.size(); ++k)
      if (features.chromosomes[k] == chrNames[c] && features.chrHasFeatures[k]) chrMap[c] = static_cast<uint32_t>(k);
  std::string name;
  std::vector<SynthRecord> recs;
  uint64_t n = 0;
  for (uint64_t r = first; r < first + nReads; ++r) {
    readRecords(r, spec, name, recs);
    const uint64_t key = name_key(name.data(), name.size());
    for (const SynthRecord &rec : recs) {
      if (n >= dst.capacity) return n;
      const bool fwd = rec.forward;
      const bool rs = s == Strandedness::F ? fwd : s == Strandedness::R ? !fwd : true;
      dst.start[n] = static_cast<uint32_t>(rec.pos);
      dst.end[n] = static_cast<uint32_t>(rec.pos + rec.len - 1);
      dst.meta[n] = (chrMap[rec.chr] & HIT_CHR_MASK) | (rs ? HIT_STRAND_BIT : 0u);
      dst.nh[n] = rec.nh;
      dst.key[n] = key;
      ++n;
    }
  }
  return n;
}

bool SynthGenome::writeBam(const std::string &path, uint64_t first, uint64_t nReads, const SynthReadSpec &spec, bool coordinateSorted, bool headerless, bool straddle) const {
  BgzfWriter w(path);
  if (!w.ok()) return false;
  std::vector<unsigned char> buf;
  if (!headerless) {
  std::string text = std::string("@HD\tVN:1.0\tSO:") + (coordinateSorted ? "coordinate" : "unsorted") + "\n";
  for (size_t c = 0; c < chrNames.size(); ++c) text += "@SQ\tSN:" + chrNames[c] + "\tLN:" + std::to_string(chrLen[c]) + "\n";
  buf.insert(buf.end(), {'B', 'A', 'M', 1});
  put32(buf, static_cast<uint32_t>(text.size()));
  buf.insert(buf.end(), text.begin(), text.end());
  put32(buf, static_cast<uint32_t>(chrNames.size()));
  for (size_t c = 0; c < chrNames.size(); ++c) {
    put32(buf, static_cast<uint32_t>(chrNames[c].size() + 1));
    buf.insert(buf.end(), chrNames[c].begin(), chrNames[c].end());
    buf.push_back(0);
    put32(buf, static_cast<uint32_t>(chrLen[c]));
  }
  w.write(buf.data(), buf.size());
  if (!straddle) w.endOfHeader();  // (htslib flushes after the header: the records start a member)
  }
  std::string name;
  std::vector<SynthRecord> recs;
  if (!coordinateSorted) {
    for (uint64_t r = first; r < first + nReads; ++r) {
      readRecords(r, spec, name, recs);
      for (const SynthRecord &rec : recs) {
        bamRecord(buf, name, rec);
        w.writeRecord(buf.data(), buf.size(), straddle);
      }
    }
  } else {
    struct Item { SynthRecord rec; uint64_t read; };
    std::vector<Item> all;
    for (uint64_t r = first; r < first + nReads; ++r) {
      readRecords(r, spec, name, recs);
      for (const SynthRecord &rec : recs) all.push_back(Item{rec, r});
    }
    std::stable_sort(all.begin(), all.end(), [](const Item &a, const Item &b) {
      return a.rec.chr != b.rec.chr ? a.rec.chr < b.rec.chr : a.rec.pos < b.rec.pos;
    });
    std::vector<SynthRecord> dummy;
    for (const Item &it : all) {
      char nm[48];
      std::snprintf(nm, sizeof(nm), "sy%llu.%llu", static_cast<unsigned long long>(seed_ % 1000), static_cast<unsigned long long>(it.read));
      bamRecord(buf, nm, it.rec);
      w.writeRecord(buf.data(), buf.size(), straddle);
    }
  }
  w.close();
  return true;
}

}  // namespace mmb
