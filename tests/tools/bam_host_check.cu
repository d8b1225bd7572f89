// CPU check of the device BAM decoder's inflate and record functions (they are __host__ __device__): every BGZF member of a BAM
// file inflated by inflateMember must equal zlib's output, and the hits parsed from the records must equal the host decoder's
// (passed in as a binary dump by tests/test_bam_decoder_host.py).  The record functions are driven member by member through
// bamCountMember / bamParseMember, the bodies of k_bam_count / k_bam_parse.  Usage: bam_host_check file.bam hits.bin strandedness
// Built with -DMMA_NAME_KEY_MASK=0x..ull -DEXPECT_COLLISIONS the read keys are cut down to a few bits and the tool checks the
// read-key verification instead: exactly the members with a colliding pair of neighbours must raise BAM_KEY_COLLISION.
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mma_bam.cuh"

using namespace mma;

static uint32_t rd32(const unsigned char *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint32_t rd16(const unsigned char *p) { return p[0] | (p[1] << 8); }

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<unsigned char> file;
  unsigned char tmp[1 << 16];
  size_t n;
  while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) file.insert(file.end(), tmp, tmp + n);
  fclose(f);
  // inflate every member twice: zlib and the decoder under test
  std::vector<unsigned char> all;
  std::vector<uint32_t> outOff(1, 0);
  size_t at = 0, members = 0;
  static Huff lit;
  static HuffDist dist;
  while (at + 18 <= file.size()) {
    const unsigned char *p = &file[at];
    const size_t xlen = rd16(p + 10), hdr = 12 + xlen, total = rd16(p + 16) + 1u;
    const uint32_t isize = rd32(p + total - 4);
    std::vector<unsigned char> a(isize + 1), b(isize + 1);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    inflateInit2(&zs, -15);
    zs.next_in = const_cast<unsigned char *>(p + hdr); zs.avail_in = (uInt)(total - hdr - 8);
    zs.next_out = a.data(); zs.avail_out = isize;
    const int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (isize && rc != Z_STREAM_END) { printf("zlib failed on member %zu\n", members); return 1; }
    if (isize && !inflateMember(p + hdr, (u32)(total - hdr - 8), b.data(), isize, lit, dist)) { printf("inflateMember failed on member %zu (isize %u)\n", members, isize); return 1; }
    if (memcmp(a.data(), b.data(), isize) != 0) { printf("member %zu differs\n", members); return 1; }
    all.insert(all.end(), b.begin(), b.begin() + isize);
    outOff.push_back((uint32_t)all.size());
    at += total;
    ++members;
  }
  // header
  const size_t lText = rd32(&all[4]);
  size_t p = 8 + lText;
  const uint32_t nRef = rd32(&all[p]);
  p += 4;
  for (uint32_t i = 0; i < nRef; ++i) p += 4 + rd32(&all[p]) + 4;
  // expected hits: n, then start/end/meta/nh (u32 each) and key (u64) arrays, then the chromosome of every BAM reference
  uint64_t nHits = 0;
  std::vector<uint32_t> es, ee, em, en, refToChr(nRef, 0x00FFFFFFu);
  std::vector<uint64_t> ek;
#if !defined(EXPECT_COLLISIONS)
  FILE *h = fopen(argv[2], "rb");
  if (!h) return 2;
  if (fread(&nHits, 8, 1, h) != 1) return 2;
  es.resize(nHits); ee.resize(nHits); em.resize(nHits); en.resize(nHits); ek.resize(nHits);
  if (fread(es.data(), 4, nHits, h) != nHits || fread(ee.data(), 4, nHits, h) != nHits || fread(em.data(), 4, nHits, h) != nHits ||
      fread(en.data(), 4, nHits, h) != nHits || fread(ek.data(), 8, nHits, h) != nHits || fread(refToChr.data(), 4, nRef, h) != nRef) return 2;
  fclose(h);
#endif
  std::vector<unsigned long long> refFirst(nRef + 1, ~0ull);
  u32 flagsWord = 0;
  BamView v;
  memset(&v, 0, sizeof(v));
  v.out = all.data(); v.refToChr = refToChr.data(); v.nRef = nRef; v.strandedness = (u32)atoi(argv[3]); v.flags = &flagsWord; v.refFirst = refFirst.data();
  // the members as the kernels see them: from the member that holds the first record (skipFirst = header bytes in it)
  size_t m0 = 0;
  while (m0 + 1 < outOff.size() && outOff[m0 + 1] <= p) ++m0;
  std::vector<uint32_t> off(outOff.begin() + m0, outOff.end());
  v.outOff = off.data(); v.nMembers = (u32)(off.size() - 1); v.skipFirst = (u32)(p - off[0]);
  // k_bam_count + k_bam_scan + k_bam_parse, member by member
  std::vector<u32> hitOff(v.nMembers + 1, 0);
  for (u32 m = 0; m < v.nMembers; ++m) {
    u32 cnt;
    if (!bamCountMember(v, m, cnt)) { printf("member %u does not hold whole records\n", m); return 1; }
    hitOff[m + 1] = hitOff[m] + cnt;
  }
  const uint64_t k = hitOff[v.nMembers];
  std::vector<uint32_t> gs(k + 1), ge(k + 1), gm(k + 1), gn(k + 1);
  std::vector<unsigned long long> gk(k + 1);
  HitOut o{gs.data(), ge.data(), gm.data(), gn.data(), gk.data()};
  u32 flags = 0;
  std::vector<u32> memberFlags(v.nMembers, 0);
  for (u32 m = 0; m < v.nMembers; ++m) flags |= (memberFlags[m] = bamParseMember(v, m, hitOff[m], hitOff[m + 1], 0, o));
#if defined(EXPECT_COLLISIONS)
  // keys cut down by MMA_NAME_KEY_MASK: the members that must carry BAM_KEY_COLLISION, from one plain walk over all records
  // (a record with its predecessor's key and another name marks the member of the PREDECESSOR: when the two lie in different
  // members the border is checked by the thread of the earlier one)
  {
    std::vector<u32> want(v.nMembers, 0);
    size_t q = p;
    u32 member = 0, prevMember = 0;
    std::string prevName;
    u64 prevKey = 0;
    bool have = false;
    size_t pairs = 0, nBorder = 0;
    while (q + 4 <= all.size()) {
      while (q >= off[member + 1]) ++member;
      const uint32_t bs = rd32(&all[q]);
      const unsigned char *nm = &all[q + 36];
      const u32 lrn = all[q + 12];
      const std::string name((const char *)nm, strnlen((const char *)nm, lrn));
      const u64 key = nameKey((const unsigned char *)name.data(), (u32)name.size()) & MMA_NAME_KEY_MASK;
      if (have && key == prevKey && name != prevName) {
        want[prevMember] = BAM_KEY_COLLISION;
        ++pairs;
        if (prevMember != member) ++nBorder;  // (reported so that the test can insist that the case occurred)
      }
      prevName = name; prevKey = key; prevMember = member; have = true;
      q += 4 + bs;
    }
    size_t nWant = 0;
    for (u32 m = 0; m < v.nMembers; ++m) {
      if (memberFlags[m] != want[m]) { printf("member %u: flags %u, expected %u\n", m, memberFlags[m], want[m]); return 1; }
      nWant += want[m] != 0;
    }
    printf("ok: %zu members, %zu colliding pairs, %zu members flagged, %zu pairs across a member border\n", members, pairs, nWant, nBorder);
    return 0;
  }
#endif
  if (flags) { printf("flags %u\n", flags); return 1; }
  if (k != nHits) { printf("records %llu != hits %llu\n", (unsigned long long)k, (unsigned long long)nHits); return 1; }
  for (uint64_t i = 0; i < nHits; ++i)
    if (gs[i] != es[i] || ge[i] != ee[i] || gm[i] != em[i] || gn[i] != en[i] || gk[i] != ek[i]) {
      printf("hit %llu differs: start %u/%u end %u/%u meta %x/%x nh %u/%u key %llx/%llx\n", (unsigned long long)i, gs[i], es[i], ge[i], ee[i], gm[i], em[i], gn[i], en[i],
             (unsigned long long)gk[i], (unsigned long long)ek[i]);
      return 1;
    }
  printf("ok: %zu members, %llu hits\n", members, (unsigned long long)nHits);
  return 0;
}
