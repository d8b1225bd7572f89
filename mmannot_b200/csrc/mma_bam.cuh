// BAM decode on the device: BGZF inflate + record parse -> the five hit arrays of the batch kernels, in HBM.
//
// The reference reads its BAM through gzread and parses one record at a time on one core (BamReader, mm:1487-1649, ~27 % of its
// time in zlib).  A BGZF file is a series of independent gzip members of at most 64 KB, so the compressed bytes can cross PCIe
// as they lie in the file (~16 B per record instead of 24 B per decoded hit) and be inflated side by side:
//
//   k_bam_inflate   one THREAD per member, MMA_BAM_LANES members per warp (divergent lanes serialise, so a warp is only given a
//                   few members and the parallelism comes from the number of warps): a complete inflate -- stored, fixed and
//                   dynamic Huffman blocks, canonical codes decoded through a 9-bit table with a bit-serial tail -- into the
//                   member's slot of one contiguous output buffer (slots from the ISIZE fields, summed on the host).
//   k_bam_count     one thread per member walks its records (block_size chain).  htslib never lets a record straddle two members,
//                   so every member is expected to start at a record and end at one; a member that does not raises a flag and the
//                   host decodes the file itself.
//   k_bam_parse     the same walk, writing one hit per record at the member's offset (exclusive scan of the counts): start =
//                   pos + 1, end = start + sum(M, D, =, X) - 1 (Read::parseCigar, mm:852-875), chromosome through the refID table
//                   the host resolved against the annotation, strand by -s (mm:836-844), NH from the record's NH tag with the
//                   reference's type rule (only unsigned types carry a value, mm:1596-1618), read key = the host's 64-bit name
//                   hash (common.hpp name_key) restated here.  The key stands for the name string the reference keys by
//                   (mm:1656-1662): whenever two neighbouring records carry the same key their names are compared byte by byte
//                   (inside a member, and the member's last record against the first record of the next member of the chunk);
//                   a pair that differs raises BAM_KEY_COLLISION.
//
// What this route does not do is flagged and left to the host decoder, for the whole file: XA tags (alternative hits need the
// text parser and the NM carried from record to record, mm:1360-1399), CIGAR operations the reference warns about ('N' and
// unknown codes, mm:871), unknown aux types, records that straddle members, corrupt deflate data.
#pragma once
#include "mma_device.cuh"

namespace mma {

#ifndef MMA_BAM_LANES
#define MMA_BAM_LANES 4  // members per warp in k_bam_inflate (measured on B200, 0.83 GB BAM: 2 or 4 -> 185 ms, 8 -> 230 ms, 16 -> 280 ms, 32 -> 447 ms)
#endif

enum BamFlag : u32 {
  BAM_BAD_DEFLATE = 1u,    // corrupt / truncated deflate stream, or ISIZE mismatch
  BAM_STRADDLE = 2u,       // a record crosses a member border (or the walk does not end at the member's end)
  BAM_HAS_XA = 4u,         // an XA tag: alternative hits, host parser
  BAM_ODD_CIGAR = 8u,      // 'N' or an unknown CIGAR operation: the reference prints a warning per occurrence
  BAM_ODD_AUX = 16u,       // unknown aux type / malformed aux area
  BAM_MALFORMED = 32u,     // record shorter than its fixed part / fields beyond the record
  BAM_KEY_COLLISION = 64u, // two neighbouring records share a read key but not the name (the host decoder names them and refuses the file)
};

struct BamView {
  const unsigned char *comp;  // whole BGZF members back to back
  const u32 *memberOff;       // [n + 1] byte offset of every member in comp
  const u32 *outOff;          // [n + 1] byte offset of every member's inflated bytes in out
  unsigned char *out;
  u32 nMembers;
  u32 skipFirst;              // bytes at the start of the FIRST member that are not records (rest of the BAM header)
  const u32 *refToChr;        // [nRef] annotation chromosome of every BAM reference (MMA_HIT_CHR_NONE: not in the annotation)
  u32 nRef;
  u32 strandedness;           // 0 U, 1 F, 2 R (mm:836-844)
  u32 uniqueOnly;             // -y unique: only hits with NH = 1 are looked at (mm:1773), so only they can reveal an unknown chromosome
  u32 *flags;                 // OR of BamFlag
  unsigned long long *refFirst;  // [nRef] ordinal of the first record seen on each reference (~0: none): for the host's warnings
};

// ---------------------------------------------------------------------------------------------- inflate

// (the inflate and record functions are __host__ __device__: tests/tools/bam_host_check.cu runs them on the CPU against zlib and
// the host decoder)
#define MMA_HD __host__ __device__
MMA_HD inline u32 bitReverse32(u32 x) {
#ifdef __CUDA_ARCH__
  return __brev(x);
#else
  u32 r = 0;
  for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
  return r;
#endif
}

struct BitReader {
  const unsigned char *p, *end;
  u64 buf;
  int cnt;
  bool over;
  MMA_HD __forceinline__ void refill() {
    if (cnt <= 32 && p + 4 <= end && ((reinterpret_cast<uintptr_t>(p) & 3) == 0)) {
      buf |= (u64)(*reinterpret_cast<const u32 *>(p)) << cnt;
      p += 4; cnt += 32;
      return;
    }
    while (cnt <= 56) {
      if (p >= end) { if (cnt < 0) cnt = 0; return; }
      buf |= (u64)(*p++) << cnt;
      cnt += 8;
    }
  }
  MMA_HD __forceinline__ u32 peek(int n) {  // n <= 24; missing bits read as zero (the caller checks `over` at the end)
    if (cnt < n) refill();
    return (u32)buf & ((1u << n) - 1u);
  }
  MMA_HD __forceinline__ void drop(int n) {
    if (n > cnt) { over = true; buf = 0; cnt = 0; return; }
    buf >>= n; cnt -= n;
  }
  MMA_HD __forceinline__ u32 bits(int n) { const u32 v = peek(n); drop(n); return v; }
};

// canonical Huffman code: counts per length + symbols in code order (bit-serial decode), and a first-level table of FB bits
// (9 for the literal / length code; 6 for the 30 distance symbols: the tables of the members in flight share the SM's shared
// memory, and what the distance table gives up lets 7 blocks run per SM instead of 5)
template <int NSYM, int FB>
struct HuffT {
  unsigned short count[16];
  unsigned short symbol[NSYM];
  unsigned short fast[1 << FB];  // (length << 9) | symbol for codes of at most FB bits, 0 = longer / invalid
};
typedef HuffT<288, 9> Huff;      // literal / length code (and the code-length code while a dynamic block's header is read)
typedef HuffT<32, 6> HuffDist;   // distance code

template <int NSYM, int FB>
MMA_HD __noinline__ bool huffBuild(HuffT<NSYM, FB> &h, const unsigned char *length, int n) {
  for (int i = 0; i < 16; ++i) h.count[i] = 0;
  for (int i = 0; i < n; ++i) h.count[length[i]]++;
  for (int i = 0; i < (1 << FB); ++i) h.fast[i] = 0;
  if (h.count[0] == n) return true;  // no codes: legal for an unused distance code
  int left = 1;
  for (int len = 1; len < 16; ++len) {
    left <<= 1;
    left -= h.count[len];
    if (left < 0) return false;  // over-subscribed
  }
  unsigned short offs[16];
  offs[1] = 0;
  for (int len = 1; len < 15; ++len) offs[len + 1] = offs[len] + h.count[len];
  for (int i = 0; i < n; ++i)
    if (length[i]) h.symbol[offs[length[i]]++] = (unsigned short)i;
  // first-level table: canonical codes are assigned in (length, symbol) order; deflate sends them most significant bit first,
  // the bit reader delivers the first bit in bit 0, so the index is the code reversed
  u32 code = 0;
  int idx = 0;
  for (int len = 1; len <= FB; ++len) {
    for (int k = 0; k < h.count[len]; ++k, ++idx, ++code) {
      const u32 rev = bitReverse32(code) >> (32 - len);
      const unsigned short e = (unsigned short)((len << 9) | h.symbol[idx]);
      for (u32 f = rev; f < (1u << FB); f += (1u << len)) h.fast[f] = e;
    }
    code <<= 1;
  }
  return true;
}

template <int NSYM, int FB>
MMA_HD __forceinline__ int huffDecode(const HuffT<NSYM, FB> &h, BitReader &br) {
  const u32 look = br.peek(15);
  const unsigned short e = h.fast[look & ((1u << FB) - 1u)];
  if (e) { br.drop(e >> 9); return e & 511; }
  // bit-serial tail (codes longer than FB bits)
  int code = 0, first = 0, index = 0;
  for (int len = 1; len < 16; ++len) {
    code |= (int)((look >> (len - 1)) & 1u);
    const int count = h.count[len];
    if (code - count < first) { br.drop(len); return h.symbol[index + (code - first)]; }
    index += count; first += count;
    first <<= 1; code <<= 1;
  }
  return -1;
}

// length / distance code tables of RFC 1951 (constant memory on the device, plain arrays for the CPU check)
#define MMA_DEFLATE_TABLES(Q)                                                                                                                  \
  Q unsigned short kLenBaseT[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258}; \
  Q unsigned char kLenExtraT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};                          \
  Q unsigned short kDistBaseT[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577}; \
  Q unsigned char kDistExtraT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};              \
  Q unsigned char kClOrderT[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
namespace devtab { MMA_DEFLATE_TABLES(__device__ __constant__) }
namespace hosttab { MMA_DEFLATE_TABLES(static const) }
#ifdef __CUDA_ARCH__
#define MMA_TAB(name) devtab::name
#else
#define MMA_TAB(name) hosttab::name
#endif
MMA_HD inline u32 kLenBase(int i) { return MMA_TAB(kLenBaseT)[i]; }
MMA_HD inline u32 kLenExtra(int i) { return MMA_TAB(kLenExtraT)[i]; }
MMA_HD inline u32 kDistBase(int i) { return MMA_TAB(kDistBaseT)[i]; }
MMA_HD inline u32 kDistExtra(int i) { return MMA_TAB(kDistExtraT)[i]; }
MMA_HD inline u32 kClOrder(int i) { return MMA_TAB(kClOrderT)[i]; }

// one raw deflate stream (RFC 1951) from [src, src + srcLen) into dst[0 .. dstLen); true when it ends exactly at dstLen
MMA_HD __noinline__ bool inflateMember(const unsigned char *src, u32 srcLen, unsigned char *dst, u32 dstLen, Huff &lit, HuffDist &dist) {
  BitReader br{src, src + srcLen, 0ull, 0, false};
  u32 out = 0;
  unsigned char lengths[320];
  for (;;) {
    const u32 last = br.bits(1), type = br.bits(2);
    if (type == 0) {  // stored
      br.drop(br.cnt & 7);
      const u32 len = br.bits(16), nlen = br.bits(16);
      if ((len ^ 0xFFFFu) != nlen || out + len > dstLen) return false;
      for (u32 i = 0; i < len; ++i) dst[out++] = (unsigned char)br.bits(8);
    } else if (type == 1 || type == 2) {
      if (type == 1) {
        int i = 0;
        for (; i < 144; ++i) lengths[i] = 8;
        for (; i < 256; ++i) lengths[i] = 9;
        for (; i < 280; ++i) lengths[i] = 7;
        for (; i < 288; ++i) lengths[i] = 8;
        huffBuild(lit, lengths, 288);
        for (i = 0; i < 30; ++i) lengths[i] = 5;
        huffBuild(dist, lengths, 30);
      } else {
        const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
        if (nlen > 286 || ndist > 30) return false;
        for (int i = 0; i < 19; ++i) lengths[i] = 0;
        for (int i = 0; i < ncode; ++i) lengths[kClOrder(i)] = (unsigned char)br.bits(3);
        if (!huffBuild(lit, lengths, 19)) return false;  // (the code-length code, built in the literal table for now)
        int idx = 0;
        while (idx < nlen + ndist) {
          int sym = huffDecode(lit, br);
          if (sym < 0) return false;
          if (sym < 16) lengths[idx++] = (unsigned char)sym;
          else {
            int prev = 0, rep;
            if (sym == 16) { if (idx == 0) return false; prev = lengths[idx - 1]; rep = 3 + (int)br.bits(2); }
            else if (sym == 17) rep = 3 + (int)br.bits(3);
            else rep = 11 + (int)br.bits(7);
            if (idx + rep > nlen + ndist) return false;
            while (rep--) lengths[idx++] = (unsigned char)prev;
          }
        }
        if (lengths[256] == 0) return false;
        // the distance lengths are used first: building the literal table overwrites nothing they need (separate arrays)
        if (!huffBuild(dist, lengths + nlen, ndist)) return false;
        if (!huffBuild(lit, lengths, nlen)) return false;
      }
      for (;;) {
        const int sym = huffDecode(lit, br);
        if (sym < 0) return false;
        if (sym < 256) {
          if (out >= dstLen) return false;
          dst[out++] = (unsigned char)sym;
        } else if (sym == 256) {
          break;
        } else {
          const int s = sym - 257;
          if (s >= 29) return false;
          const u32 len = kLenBase(s) + br.bits((int)kLenExtra(s));
          const int ds = huffDecode(dist, br);
          if (ds < 0 || ds >= 30) return false;
          const u32 d = kDistBase(ds) + br.bits((int)kDistExtra(ds));
          if (d > out || out + len > dstLen) return false;
          // The source lies d bytes behind the destination.  With d >= 8 the copy goes by aligned 32-bit words: destination word k
          // is cut out of two aligned source words by a funnel shift, and the second of them -- reused as the first of word
          // k + 1 -- ends at most 6 bytes behind the source position, i.e. before anything this copy has yet to write (the
          // byte-wise copy was 35 % of the kernel's issue slots: the decoder runs one lane per instruction).  A run-length
          // match (d < 8) repeats its d bytes one by one.
          unsigned char *to = dst + out;
          const unsigned char *from = to - d;
          u32 i = 0;
          if (d >= 8 && len >= 8) {
            const u32 head = (4u - (u32)((size_t)to & 3u)) & 3u;
            for (; i < head; ++i) to[i] = from[i];
            const size_t sa = (size_t)(from + i);
            const u32 sh = (u32)(sa & 3u) * 8u;
            const u32 *sw = reinterpret_cast<const u32 *>(sa & ~(size_t)3);
            u32 *dw = reinterpret_cast<u32 *>(to + i);
            if (sh == 0) {
              for (; i + 4 <= len; i += 4) *dw++ = *sw++;
            } else {
              u32 lo = *sw++;
              for (; i + 4 <= len; i += 4) {
                const u32 hi = *sw++;
#ifdef __CUDA_ARCH__
                *dw++ = __funnelshift_r(lo, hi, sh);
#else
                *dw++ = (lo >> sh) | (hi << (32u - sh));
#endif
                lo = hi;
              }
            }
          }
          for (; i < len; ++i) to[i] = from[i];
          out += len;
        }
        if (br.over) return false;
      }
    } else {
      return false;
    }
    if (br.over) return false;
    if (last) break;
  }
  return out == dstLen;
}

MMA_HD __forceinline__ u32 ld32u(const unsigned char *p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }
MMA_HD __forceinline__ u32 ld16u(const unsigned char *p) { return (u32)p[0] | ((u32)p[1] << 8); }

// one thread per member, MMA_BAM_LANES members per warp.  The grid is what the GPU holds at once (the Huffman tables in shared
// memory bound it at ~80 members per SM) and every lane takes the next member from a counter when it is done with one, so the
// size of a launch does not matter.  (The kernel is bound by issue slots, not by its tail: 1.02 threads per warp instruction at
// 57 % issue utilisation -- profiles/r02b_inflate.txt.)
#define BAM_INFLATE_THREADS 128
__global__ void __launch_bounds__(BAM_INFLATE_THREADS) k_bam_inflate(BamView v, u32 *next) {
  // the Huffman tables of the members in flight live in shared memory (30 KB per block: every symbol is a table lookup, and
  // local memory behind L1 missed 6 % of them)
  __shared__ Huff litTab[BAM_INFLATE_THREADS / 32][MMA_BAM_LANES];
  __shared__ HuffDist distTab[BAM_INFLATE_THREADS / 32][MMA_BAM_LANES];
  const u32 lane = threadIdx.x & 31u;
  if (lane >= MMA_BAM_LANES) return;
  for (u32 m = atomicAdd(next, 1u); m < v.nMembers; m = atomicAdd(next, 1u)) {
    const unsigned char *p = v.comp + v.memberOff[m];
    const u32 total = v.memberOff[m + 1] - v.memberOff[m];
    const u32 want = v.outOff[m + 1] - v.outOff[m];
    // gzip member: 10 fixed bytes, XLEN + extra field (FLG.FEXTRA is set in BGZF), deflate data, CRC32, ISIZE
    bool ok = total >= 28 && p[0] == 0x1f && p[1] == 0x8b && p[2] == 8 && (p[3] & 4);
    u32 hdr = 0;
    if (ok) {
      hdr = 12 + ld16u(p + 10);
      if (p[3] & ~4u) ok = false;  // (name / comment / header CRC fields: not BGZF)
      if (hdr + 8 > total) ok = false;
    }
    if (ok && ld32u(p + total - 4) != want) ok = false;
    if (ok && want) ok = inflateMember(p + hdr, total - hdr - 8, v.out + v.outOff[m], want, litTab[threadIdx.x >> 5][lane], distTab[threadIdx.x >> 5][lane]);
    if (!ok) atomicOr(v.flags, (u32)BAM_BAD_DEFLATE);
  }
}

// ---------------------------------------------------------------------------------------------- records

// the host's name_key (csrc/host/common.hpp), restated: the read key of every hit must be the one the host decoder would give
MMA_HD __forceinline__ u64 nameKey(const unsigned char *p, u32 n) {
  u64 h = 0x9E3779B97F4A7C15ull ^ ((u64)n * 0xD6E8FEB86659FD93ull);
  while (n >= 8) {
    const u64 w = (u64)ld32u(p) | ((u64)ld32u(p + 4) << 32);
    h = (h ^ w) * 0xFF51AFD7ED558CCDull;
    h ^= h >> 32;
    p += 8; n -= 8;
  }
  u64 w = 0;
  for (u32 i = 0; i < n; ++i) w |= (u64)p[i] << (8 * i);
  h = (h ^ w) * 0xC4CEB9FE1A85EC53ull;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return h;
}

// records of member m: [begin, end) inside v.out
MMA_HD __forceinline__ void memberRange(const BamView &v, u32 m, u32 &begin, u32 &end) {
  begin = v.outOff[m] + (m == 0 ? v.skipFirst : 0u);
  end = v.outOff[m + 1];
}

// number of records of member m; false when the member does not start at a record and end at one
MMA_HD __forceinline__ bool bamCountMember(const BamView &v, u32 m, u32 &n) {
  u32 pos, end;
  memberRange(v, m, pos, end);
  n = 0;
  bool bad = pos > end;
  while (!bad && pos < end) {
    if (pos + 4 > end) { bad = true; break; }
    const u32 bs = ld32u(v.out + pos);
    if (bs < 32 || (u64)pos + 4 + bs > end) { bad = true; break; }
    pos += 4 + bs;
    ++n;
  }
  if (bad) n = 0;
  return !bad;
}

__global__ void k_bam_count(BamView v, u32 *count) {
  const u32 m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= v.nMembers) return;
  u32 n;
  if (!bamCountMember(v, m, n)) atomicOr(v.flags, (u32)BAM_STRADDLE);
  count[m] = n;
}

struct HitOut {
  u32 *start, *end, *meta, *nh;
  u64 *key;
};

// read name of a record as the key was formed from it (bytes up to the first NUL, mm:1545); p = nullptr: no record
struct RecName {
  const unsigned char *p;
  u32 len;
  u64 key;
};
MMA_HD __forceinline__ bool sameName(const RecName &a, const RecName &b) {
  if (a.len != b.len) return false;
  for (u32 i = 0; i < a.len; ++i)
    if (a.p[i] != b.p[i]) return false;
  return true;
}
#ifndef MMA_NAME_KEY_MASK
#define MMA_NAME_KEY_MASK (~0ull)  // (tests/tools/bam_host_check.cu cuts the keys down to a few bits to stage collisions)
#endif

// one BAM alignment record (the bytes after its block_size field) -> one hit; like XamReader::parseBamRecord without XA
MMA_HD __forceinline__ u32 bamRecord(const BamView &v, const unsigned char *p, u32 blockSize, u64 ordinal, const HitOut &o, u32 at, RecName &name) {
  u32 flags = 0;
  name.p = nullptr; name.len = 0; name.key = 0;
  const unsigned char *recEnd = p + blockSize;
  const int refId = (int)ld32u(p), pos = (int)ld32u(p + 4);
  const u32 lReadName = p[8];
  const u32 flagNc = ld32u(p + 12), flag = flagNc >> 16, nCigar = flagNc & 0xffffu, lSeq = ld32u(p + 16);
  const unsigned char *q = p + 32;
  if ((u64)lReadName + 4ull * nCigar + ((u64)lSeq + 1) / 2 + lSeq + 32 > blockSize) {
    // malformed: the host decoder skips such a record without a hit; the counts would no longer match
    return (u32)BAM_MALFORMED;
  }
  u32 nameLen = 0;
  while (nameLen < lReadName && q[nameLen]) ++nameLen;  // up to the first NUL (mm:1545)
  const u64 key = nameKey(q, nameLen) & MMA_NAME_KEY_MASK;
  name.p = q; name.len = nameLen; name.key = key;
  q += lReadName;
  const u64 start = (u64)((long long)pos + 1);  // ++pos then widened (mm:1536-1537)
  u64 end = start;
  for (u32 i = 0; i < nCigar; ++i, q += 4) {
    const u32 c = ld32u(q), op = c & 15u, len = c >> 4;
    // MIDNSHP=X: M D = X advance the reference (mm:852-875); I S H P do not; N and unknown codes draw a warning
    if (op == 0 || op == 2 || op == 7 || op == 8) end += (u64)(long long)(int)len;
    else if (op == 3 || op > 8) flags |= BAM_ODD_CIGAR;
  }
  end -= 1;  // (an empty CIGAR gives end = start - 1)
  q += (lSeq + 1) / 2 + lSeq;
  u32 nHits = 1;
  while (q + 3 <= recEnd) {  // aux fields (mm:1563-1648); unsigned integer types only feed NH (mm:1596-1618)
    const unsigned char t0 = q[0], t1 = q[1], ty = q[2];
    q += 3;
    u32 vU = 0;
    bool bad = false;
    switch (ty) {
      case 'A': case 'c': q += 1; break;
      case 'C': if (q + 1 <= recEnd) vU = q[0]; q += 1; break;
      case 's': q += 2; break;
      case 'S': if (q + 2 <= recEnd) vU = ld16u(q); q += 2; break;
      case 'i': case 'f': q += 4; break;
      case 'I': if (q + 4 <= recEnd) vU = ld32u(q); q += 4; break;
      case 'Z': case 'H': { while (q < recEnd && *q) ++q; ++q; break; }
      case 'B': {
        if (q + 5 > recEnd) { bad = true; break; }
        const unsigned char sub = q[0];
        const u64 cnt = ld32u(q + 1), width = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
        if (cnt * width > blockSize) { bad = true; break; }
        q += 5 + cnt * width;
        break;
      }
      default: flags |= BAM_ODD_AUX; bad = true;
    }
    if (bad || q > recEnd) break;
    if (t0 == 'N' && t1 == 'H') nHits = vU;
    else if (t0 == 'X' && t1 == 'A') flags |= BAM_HAS_XA;
  }
  // chromosome: an unmapped record / out-of-range refID is "*": no hit on any feature and no warning
  u32 chr = 0x00FFFFFFu;
  if (refId >= 0 && (u32)refId < v.nRef) {
    chr = v.refToChr[refId];
    // (only the warnings about chromosomes the annotation does not know use this; an atomic per record on a handful of addresses
    // tripled the time of this kernel)
    if ((chr & 0x00FFFFFFu) == 0x00FFFFFFu && (!v.uniqueOnly || nHits == 1)) {
#ifdef __CUDA_ARCH__
      if (ordinal < *(volatile const unsigned long long *)&v.refFirst[refId]) atomicMin(&v.refFirst[refId], (unsigned long long)ordinal);
#else
      if (ordinal < v.refFirst[refId]) v.refFirst[refId] = ordinal;
#endif
    }
  }
  // XamReader::pushRecordHits emit(): coordinates beyond the 32-bit range cannot touch any feature
  u64 s = start, e = end;
  if (s > 0xFFFFFFFEull) { chr = 0x00FFFFFFu; s = 0; e = 0; }
  else if (e == ~0ull) e = 0xFFFFFFFFull;
  else if (e > 0xFFFFFFFEull) e = 0xFFFFFFFEull;
  const bool fwd = (flag & 0x10u) == 0;
  const bool strand = v.strandedness == 1 ? fwd : v.strandedness == 2 ? !fwd : true;
  o.start[at] = (u32)s; o.end[at] = (u32)e; o.meta[at] = (chr & 0x00FFFFFFu) | (strand ? 0x80000000u : 0u); o.nh[at] = nHits; o.key[at] = key;
  return flags;
}

// the records of member m -> hits [at, stop); returns the OR of the records' flags.  Read-key verification on the way: a record
// with its predecessor's key must have its predecessor's name, and so must the first record of the next member that holds one
// (members without records lie between the parts of a concatenated file and before the end of every BGZF file).
MMA_HD __forceinline__ u32 bamParseMember(const BamView &v, u32 m, u32 at, u32 stop, u64 ordBase, const HitOut &o) {
  u32 pos, end;
  memberRange(v, m, pos, end);
  u32 flags = 0;
  RecName prev{nullptr, 0, 0};
  while (pos + 4 <= end && at < stop) {
    const u32 bs = ld32u(v.out + pos);
    RecName cur;
    flags |= bamRecord(v, v.out + pos + 4, bs, ordBase + at, o, at, cur);
    if (cur.p) {
      if (prev.p && cur.key == prev.key && !sameName(prev, cur)) flags |= BAM_KEY_COLLISION;
      prev = cur;
    }
    pos += 4 + bs;
    ++at;
  }
  if (prev.p) {
    for (u32 nx = m + 1; nx < v.nMembers; ++nx) {
      u32 nb, ne;
      memberRange(v, nx, nb, ne);
      if (nb >= ne) continue;  // no record in this member
      // (a member that does not start with a whole record was flagged by k_bam_count: only what is read here must lie inside it)
      if ((u64)nb + 36 <= ne) {
        const unsigned char *q = v.out + nb + 36;
        const u32 lReadName = v.out[nb + 12];
        if ((u64)nb + 36 + lReadName <= ne) {
          RecName first{q, 0, 0};
          while (first.len < lReadName && q[first.len]) ++first.len;
          first.key = nameKey(q, first.len) & MMA_NAME_KEY_MASK;
          if (first.key == prev.key && !sameName(prev, first)) flags |= BAM_KEY_COLLISION;
        }
      }
      break;
    }
  }
  return flags;
}

__global__ void k_bam_parse(BamView v, const u32 *__restrict__ hitOff, u64 ordBase, HitOut o) {
  const u32 m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= v.nMembers) return;
  const u32 flags = bamParseMember(v, m, hitOff[m], hitOff[m + 1], ordBase, o);
  if (flags) atomicOr(v.flags, flags);
}

// exclusive prefix sum of the per-member record counts (single block: a chunk holds tens of thousands of members)
__global__ void k_bam_scan(const u32 *__restrict__ count, u32 n, u32 *off) {
  __shared__ u32 warpSum[32];
  __shared__ u32 carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (u32 base = 0; base < n; base += blockDim.x) {
    const u32 i = base + threadIdx.x;
    const u32 x = (i < n) ? count[i] : 0u;
    u32 val = x;
    for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, val, d); if (lane >= (u32)d) val += t; }
    if (lane == 31) warpSum[warp] = val;
    __syncthreads();
    if (warp == 0) {
      u32 w = (lane < (blockDim.x >> 5)) ? warpSum[lane] : 0u;
      for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, w, d); if (lane >= (u32)d) w += t; }
      warpSum[lane] = w;
    }
    __syncthreads();
    const u32 incl = val + carry + (warp > 0 ? warpSum[warp - 1] : 0u);
    if (i < n) off[i] = incl - x;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[n] = carry;
}

}  // namespace mma
