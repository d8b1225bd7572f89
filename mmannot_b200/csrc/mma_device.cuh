// Device-side data model and kernels of the read-annotation hot path (sm_100a).
//
// K1  index build : packed features, running max-end, position bins (first-start + spanning lists), and the
//                   SEGMENT ANSWER TABLE: the chromosome cut at every feature boundary; inside one segment the
//                   set of covering features is constant, so the element set of a read that stays inside a
//                   segment (or crosses exactly one boundary, inclusion mode) is precomputed per strand.
// K2-K4 k_batch   : ONE kernel per hit batch.  Per tile of 1024 hits: 16-byte vector loads of the packed hit
//                   arrays -> per-hit element set (segment table, 1-2 L2 gathers; hits the table cannot answer
//                   are compacted through shared memory and evaluated against the feature index) -> per-read
//                   NH countdown over the run of records sharing a read key, in shared memory -> counts in a
//                   block-private shared-memory table flushed once per block into the device hash table.
//     k_batch_close : end-of-batch bookkeeping; when the batch left an unfinished read in its middle, re-routes
//                   the other runs of that read name to the deferred list (undoing what k_batch counted for them).
//     k_slow_*    : deferred (name-sorted) resolution at end of sample.
//
// Reference semantics are cited per function as mm:LINE (= /root/reference/mmannot.cpp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mma {

typedef unsigned long long u64;
typedef unsigned int u32;

// ----------------------------------------------------------------------------- layouts

// feature meta word (built by k_pack_features)
//   bits 0..5   flattened Order element index
//   bits 6..21  Order line of that element (priority rank)
//   bits 22..23 element strand (0 all, 1 F, 2 R)
//   bits 24..25 element vicinity (0 none, 1 upstream, 2 downstream)
//   bit  26     feature lies on the '+' strand
#define FM_TYPE(m) ((m)&63u)
#define FM_LINE(m) (((m) >> 6) & 0xFFFFu)
#define FM_ESTRAND(m) (((m) >> 22) & 3u)
#define FM_VIC(m) (((m) >> 24) & 3u)
#define FM_FWD(m) (((m) >> 26) & 1u)

struct IndexView {
  const uint4 *feat;     // {start, end, meta, running max end within the chromosome}
  const uint2 *chrInfo;  // per chromosome {first bin entry, number of bins}
  const uint2 *bins;     // per bin entry {first feature starting at/after the bin start, offset into spanIdx}
  const u32 *spanIdx;    // features that start before a bin and reach into it, in feature order
  u32 nChr, shift;
};

// Segment answer table (only built when E <= 30: bits 30 and 31 of an answer word are flags)
//   The chromosomes are cut at every feature boundary; segment i of a chromosome spans [start_i, end_i].
//   position map  bm[chromosome base + (pos >> shift)] = {bits, rank}: the bin of 2^shift positions is cut into 32 granules of
//               2^gshift positions (shift = gshift + 5).  rank = index of the segment holding the bin's first position; bit p =
//               "a segment starts inside granule p" (the bin's first position itself excluded).  With gshift = 0 (one position
//               per bit: annotations up to ~128 Mb) the segment holding position x is EXACTLY rank + popc(bits up to x's bit);
//               with coarser granules it is a lower bound, corrected by stepping right (a boundary in x's own granule, or two in
//               one granule).
//   segment i   seg[2i]   = {end_i, answer F, answer R, len_{i+1} | len_{i+2} << 16}        read inside the segment
//               seg[2i+1] = {cross answer F, cross answer R, triple answer F, triple answer R}
//               (inclusion mode) cross: read = tail of i + head of i+1; triple: tail of i + all of i+1 + head of i+2.
//               len = length of the segment, saturated at 65535 (a saturated length only vouches for reads ending within
//               65534 positions of the previous boundary); 0 when the chromosome has no such segment.
//               tie[2i + {0 F, 1 R}] = tie point of an ANS_VICPAIR in-segment answer
// "answer F" is for a read whose strand bit is set (MMA_HIT_STRAND_BIT), "answer R" for the other one.
// An answer word is the element set (E <= 30), or carries one of two flags:
//   ANS_VICPAIR  the winning Order line matched exactly one upstream and one downstream element (bits 0..29 hold both): the
//                pick goes to the nearer one (mm:1066-1075).  With U = end of the upstream feature and D = start of the
//                downstream one, the read [s, e] is at distance U - e from the first and s - D from the second, so the pick
//                is upstream iff s + e > U + D, downstream iff s + e < U + D, both on a tie: "tie point" = U + D.  Only
//                in-segment answers carry the flag (a cross / triple answer of that kind is stored as ANS_GENERAL).
//   ANS_GENERAL  any other position-dependent pick: the table cannot answer
#define ANS_VICPAIR 0x80000000u
#define ANS_GENERAL 0x40000000u
//   bin entries   (annotations up to ~160 Mb: TAIR10, FlyBase) replace the position map: ONE 16-byte gather answers the common read.
//               ent[e] describes the 64 positions [64b, 64b + 63] of a chromosome (e = chromosome base + b):
//                 x, y   bits: bit p (1..63) = "a segment starts at 64b + p"
//                 z      lenZ (16 bits) | lenZ1 (8 bits) << 16      Z = the segment holding the bin's LAST position; lenZ = how far
//                        Z reaches beyond the bin (saturated at 65535), lenZ1 = length of the segment after Z (saturated at 255;
//                        0 when lenZ is saturated or there is no such segment)
//                 w      idZ (10 bits) | idZX << 10 | idA << 20      indices into the PAIR DICTIONARY dict[] = {answer F, answer R}:
//                        idZ the answers of a read inside Z, idZX of a read over Z and Z+1 (inclusion mode), idA of a read inside
//                        A = the segment holding the bin's FIRST position
//               rank[e] = index of A (the segment table index of a position p of the bin is rank + popc(bits up to p): exact).
//               A read [s, e] with no boundary after s inside its start bin starts in Z; with x = e - (s | 63) it lies inside Z
//               when x <= lenZ and over Z and Z+1 when 0 < x - lenZ <= lenZ1.  A read with no boundary up to s in its bin starts in
//               A and lies inside A when e stays before the bin's first boundary.  An answer the dictionary cannot give (position
//               dependent pick, or more than ENT_DICT distinct pairs) is ENT_NONE: the read then takes the segment record like
//               every other read.  The last bin of a chromosome is always empty, so a read starting beyond the annotated extent
//               finds the last segment there.  dict[0] = {0, 0} (also the answer of the dummy bin of unknown chromosomes),
//               dict[1] = {ENT_NONE, ENT_NONE}; nDict entries are in use.
#define ENT_NONE 0xFFFFFFFFu
#define ENT_DICT 1024u
struct FastView {
  const uint2 *bm;
  const uint4 *seg;
  const u32 *tie;
  const uint2 *chrInfo;  // per chromosome {first entry of bm, number of bins}
  u32 nChr, shift, gshift, enabled;
  u32 upMask, downMask;  // upstream / downstream elements (Config::isUpstream / isDownstream, mm:463-470)
  const uint4 *ent;      // bin entries (null: position map `bm` instead); then chrInfo = {first entry, number of 64-position bins},
                         // chrInfo[nChr] = an empty dummy bin whose answers are 0, shift = 6, gshift = 0
  const u32 *rank;       // bin entries: segment index of the first position of each bin
  const uint2 *dict;     // bin entries: answer pairs
  u32 nDict;             // bin entries: pairs in use (<= ENT_DICT)
};

struct HitView {
  const u32 *start, *end, *meta, *nh;
  const u64 *key;
  u32 n;
  u32 vec;  // all five arrays are 16-byte aligned: the tile loads may use 128-bit accesses
};

// open-addressing table: combination key -> count.  key 0 = empty (an empty element set is never counted)
struct TableView {
  u64 *keys;
  u64 *vals;
  u32 capMask;
  u32 *overflow;
};

enum StatSlot { ST_HITS = 0, ST_READS, ST_UNIQUE, ST_AMBIGUOUS, ST_MULTIPLE, ST_UNASSIGNED, ST_RESCUED, ST_N = 8 };

// an open multi-mapping read whose run of records reaches the end of a batch
struct Carry {
  u64 key;        // normalised read key
  u64 ord;        // ordinal of the read's first record
  u64 gm;         // union of the element sets seen so far
  u32 remaining;  // records still expected (mm:1673)
  u32 valid;
};

// control block of one sample (device memory)
struct SampleCtl {
  u64 stats[ST_N];   // Counter's counters (mm:1663)
  u64 ordBase;       // ordinal of the first hit of the batch in flight
  Carry carry[2];    // [batchSeq & 1] = carried into the batch in flight, the other one = carried out of it
  u32 batchSeq;      // number of batches closed so far
  u32 slowCount;     // deferred records
  u32 openCount;     // entries in the open-key set
  u32 dirty;         // the batch in flight left an unfinished read in its middle
  u32 overflow;      // any capacity problem (tables, deferred list, key set)
  u32 closeCount;    // blocks of k_batch_close that are done
  u32 fastMiss;      // hits the segment table could not answer (diagnostics)
  u32 walkCount;     // k_batch_fast: runs that needed the serial walker, or (GROUPS variant) groups beyond the first of a run
  u32 deferSegs;     // deferred pass: read names it resolved ...
  u32 deferContig;   // ... of which all records were adjacent in the file (name-grouped input that did not need deferring)
};

struct SlowView {  // deferred records: resolved after a (key, ordinal) sort at end of sample
  u64 *key, *ord, *mask;
  u32 *nh;
  u32 cap;
};

struct KeySetView {  // read names whose multi-mapping records must all take the deferred path
  u64 *keys;
  u32 *seq;  // batch in which the key was inserted
  u32 capMask;
};

#define KEY_EMPTY 0xFFFFFFFFFFFFFFFFull
#define NH_SHIFT 40  // -y ratio: combination key = element mask | NH << 40

struct Rules {
  int strategy;   // MMA_STRATEGY_*
  int mode;       // 0 inclusion, 1 fraction, 2 nucleotides (mm:1974-1976)
  float overlap;  // Globals::overlap
  int rescue;     // rescue() active (mm:491, 2025)
  float rescueThreshold;
  u32 nElements;
};

// ----------------------------------------------------------------------------- helpers

__device__ __forceinline__ u64 mix64(u64 x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
__device__ __forceinline__ u32 mix32(u64 x) {
  u32 h = (u32)x ^ ((u32)(x >> 32) * 0x9E3779B1u);
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
  return h;
}
__device__ __forceinline__ u64 normKey(u64 k) { return k == KEY_EMPTY ? KEY_EMPTY - 1 : k; }

__device__ __forceinline__ void tableAdd(const TableView &t, u64 ckey, u64 add) {
  u32 slot = (u32)mix64(ckey) & t.capMask;
  for (u32 probe = 0; probe <= t.capMask; ++probe) {
    u64 old = t.keys[slot];
    if (old != ckey) {
      if (old != 0) { slot = (slot + 1) & t.capMask; continue; }
      old = atomicCAS(&t.keys[slot], 0ull, ckey);
      if (old != 0 && old != ckey) { slot = (slot + 1) & t.capMask; continue; }
    }
    atomicAdd(&t.vals[slot], add);
    return;
  }
  atomicExch(t.overflow, 1u);
}

// Block-private table in shared memory (K4 privatisation); flushed once per block.
template <int SLOTS>
struct BlockTable {
  u64 keys[SLOTS];
  u32 cnt[SLOTS];
  __device__ void init() {
    for (int i = threadIdx.x; i < SLOTS; i += blockDim.x) { keys[i] = 0; cnt[i] = 0; }
  }
  __device__ __noinline__ void add(u64 ckey, u32 n, const TableView &g) {
    u32 slot = mix32(ckey) & (SLOTS - 1);
#pragma unroll 1
    for (int probe = 0; probe < 8; ++probe) {
      u64 old = keys[slot];
      if (old != ckey) {
        if (old == 0) old = atomicCAS(&keys[slot], 0ull, ckey);
        if (old != 0 && old != ckey) { slot = (slot + 1) & (SLOTS - 1); continue; }
      }
      atomicAdd(&cnt[slot], n);
      return;
    }
    tableAdd(g, ckey, n);  // crowded block table: go to the device table directly
  }
  __device__ void flush(const TableView &g) {
    for (int i = threadIdx.x; i < SLOTS; i += blockDim.x)
      if (cnt[i]) tableAdd(g, keys[i], cnt[i]);
  }
};

// every lane of the warp calls this (ckey = 0 for lanes with nothing to count)
template <int SLOTS>
__device__ __forceinline__ void warpCount(BlockTable<SLOTS> &bt, u64 ckey, const TableView &g) {
  if (__ballot_sync(0xffffffffu, ckey != 0) == 0) return;
  const u32 peers = __match_any_sync(0xffffffffu, ckey);
  if (ckey != 0 && (u32)(__ffs(peers) - 1) == (threadIdx.x & 31u)) bt.add(ckey, __popc(peers), g);
}

// returns 0 = absent, 1 = present from an earlier batch, 2 = inserted during batch `curSeq` (or later)
__device__ __forceinline__ int keySetLookup(const KeySetView &s, u64 key, u32 curSeq) {
  u32 slot = (u32)mix64(key) & s.capMask;
  for (u32 probe = 0; probe <= s.capMask; ++probe) {
    const u64 k = s.keys[slot];
    if (k == key) return (*(volatile const u32 *)&s.seq[slot] < curSeq) ? 1 : 2;
    if (k == KEY_EMPTY) return 0;
    slot = (slot + 1) & s.capMask;
  }
  return 0;
}
__device__ __forceinline__ void keySetInsert(const KeySetView &s, u64 key, u32 curSeq, SampleCtl *ctl) {
  u32 slot = (u32)mix64(key) & s.capMask;
  for (u32 probe = 0; probe <= s.capMask; ++probe) {
    u64 k = s.keys[slot];
    if (k == key) return;
    if (k == KEY_EMPTY) {
      k = atomicCAS(&s.keys[slot], KEY_EMPTY, key);
      if (k == KEY_EMPTY) {
        *(volatile u32 *)&s.seq[slot] = curSeq;
        atomicAdd(&ctl->openCount, 1u);
        return;
      }
      if (k == key) return;
    }
    slot = (slot + 1) & s.capMask;
  }
  atomicExch(&ctl->overflow, 1u);
}

__device__ __forceinline__ void slowAppend(const SlowView &s, SampleCtl *ctl, u64 key, u64 ord, u64 mask, u32 nh) {
  const u32 at = atomicAdd(&ctl->slowCount, 1u);
  if (at >= s.cap) { atomicExch(&ctl->overflow, 1u); return; }
  s.key[at] = key; s.ord[at] = ord; s.mask[at] = mask; s.nh[at] = nh;
}

// ----------------------------------------------------------------------------- per-hit annotation against the feature index

struct HitEval {  // running best of the winning Order line (EvaluationStructure::getFirst, mm:1029-1076)
  u64 seen;       // element types already decided (we walk the candidates backwards: the first
                  // passing interval met for a type is the LAST one in feature order, mm:1023-1028)
  u64 chosen;
  u64 lineAll;    // TRACK only: every matched element of the winning line
  u32 bestLine, bestOv, bestDist;
  u32 pUp, pDown; // TRACK only: reference coordinate of the line's upstream / downstream match (mm:1316-1322)
};

template <int MODE, bool TRACK>
__device__ __forceinline__ void evalCandidate(const uint4 f, u32 rs, u32 re, u32 rstrand, float ovl, HitEval &ev) {
  if (f.x > re) return;  // the reference stops its walk at the first interval starting after the read (mm:1311)
  // ... and starts it at the first interval, from the bin of the read start on, that does not end before the read
  // (mm:1303-1308).  Everything before that bin ends before the read too, so an interval is walked iff the running
  // maximum of the ends up to it reaches the read start.  Only matters for an empty CIGAR (end = start - 1, mm:874)
  // against a feature ending exactly at that end: it is "included" (mm:638) unless the walk skipped it.
  if (f.w < rs) return;
  const u32 m = f.z;
  const u32 t = FM_TYPE(m);
  if ((ev.seen >> t) & 1ull) return;
  const u32 es = FM_ESTRAND(m);
  if (es != 0) {  // Config::checkStrand, mm:438-443
    const bool same = (FM_FWD(m) == rstrand);
    if ((es == 1) != same) return;
  }
  u32 sc;
  if (MODE == 0) {  // intervalInclusion, mm:992-994
    sc = (rs >= f.x && re <= f.y) ? 1u : 0u;
  } else {
    const u32 s = max(f.x, rs), e = min(f.y, re);  // Interval::overlaps, mm:632-636
    const u32 o = (s >= e) ? 0u : e - s;
    if (MODE == 1) {  // intervalOverlapPc, mm:995-998 (fp32 product and compare)
      const u32 size = re - rs + 1u;
      sc = (__fmul_rn((float)size, ovl) <= (float)o) ? o : 0u;
    } else {  // intervalOverlap, mm:999-1002
      sc = ((float)o >= ovl) ? o : 0u;
    }
  }
  if (sc == 0) return;
  const u32 vic = FM_VIC(m);  // distance to the gene-side border, mm:1316-1322 + Interval::getDistance mm:661-665
  u32 d = 0;
  if (vic != 0) {
    const u32 p = (vic == 1) ? f.y : f.x;
    d = (p < rs) ? rs - p : (p > re) ? p - re : 0u;
  }
  ev.seen |= 1ull << t;
  const u32 line = FM_LINE(m);
  const u64 bit = 1ull << t;
  if (line < ev.bestLine) {
    ev.bestLine = line; ev.bestOv = sc; ev.bestDist = d; ev.chosen = bit;
    if (TRACK) { ev.lineAll = bit; ev.pUp = f.y; ev.pDown = f.x; }
  } else if (line == ev.bestLine) {
    if (TRACK) { ev.lineAll |= bit; if (vic == 1) ev.pUp = f.y; if (vic == 2) ev.pDown = f.x; }
    if (sc > ev.bestOv) { ev.bestOv = sc; ev.bestDist = d; ev.chosen = bit; }
    else if (sc == ev.bestOv) {
      if (d < ev.bestDist) { ev.bestDist = d; ev.chosen = bit; }
      else if (d == ev.bestDist) ev.chosen |= bit;
    }
  }
}

// IntervalList::scan, mm:1291-1332, as an index lookup: the candidates are the features that reach
// into the bin of the read start (spanning list) plus those that start in the bins the read covers.
struct EvalTrack { u64 lineAll; u32 pUp, pDown; };

template <int MODE, bool TRACK>
__device__ __forceinline__ u64 annotateEval(const IndexView &ix, u32 rs, u32 re, u32 meta, float ovl, EvalTrack *track) {
  const u32 chr = meta & 0x00FFFFFFu;
  if (TRACK) { track->lineAll = 0; track->pUp = 0; track->pDown = 0; }
  if (chr >= ix.nChr) return 0;
  const uint2 ci = __ldg(&ix.chrInfo[chr]);
  const u32 lastBin = ci.y - 1;
  const u32 qlo = min(rs, re), qhi = max(rs, re);  // re < rs only for an empty CIGAR (mm:874)
  const u32 b0 = min(qlo >> ix.shift, lastBin), b1 = min(qhi >> ix.shift, lastBin);
  const uint2 e0 = __ldg(&ix.bins[ci.x + b0]);
  const uint2 e1 = __ldg(&ix.bins[ci.x + b0 + 1]);
  const u32 hi = (b1 == b0) ? e1.x : __ldg(&ix.bins[ci.x + b1 + 1]).x;
  const u32 rstrand = meta >> 31;
  HitEval ev;
  ev.seen = 0; ev.chosen = 0; ev.lineAll = 0; ev.bestLine = 0xFFFFFFFFu; ev.bestOv = 0; ev.bestDist = 0; ev.pUp = 0; ev.pDown = 0;
  for (u32 i = hi; i-- > e0.x;) evalCandidate<MODE, TRACK>(__ldg(&ix.feat[i]), rs, re, rstrand, ovl, ev);
  for (u32 k = e1.y; k-- > e0.y;) evalCandidate<MODE, TRACK>(__ldg(&ix.feat[__ldg(&ix.spanIdx[k])]), rs, re, rstrand, ovl, ev);
  if (TRACK) { track->lineAll = ev.lineAll; track->pUp = ev.pUp; track->pDown = ev.pDown; }
  return ev.chosen;
}

template <int MODE>
__device__ __noinline__ u64 annotateHit(const IndexView &ix, u32 rs, u32 re, u32 meta, float ovl) {
  return annotateEval<MODE, false>(ix, rs, re, meta, ovl, nullptr);
}

// Segment-table lookup.
//   * a read inside one segment is covered by exactly the features covering the segment; they all score alike
//     (1 under inclusion, end - start under the overlap modes), so the pick is the precomputed one -- provided the
//     read is long enough to match anything at all under -l (mm:995-1002);
//   * under inclusion a read over two adjacent segments is included in exactly the features covering both.
// Two dependent gathers: the position map entry of the read start, then the 32-byte record of its segment.  What the
// table cannot answer (read over three or more segments, position-dependent picks other than the upstream/downstream
// tie, degenerate intervals) is evaluated against the feature index; bit 31 of the result says so (statistics).
#define FAST_MISS 0x80000000u

// segment index (lower bound, exact when gshift == 0) of position rs on chromosome info ci
__device__ __forceinline__ u32 fastSegIndex(const FastView &fx, uint2 ci, u32 rs) {
  if (fx.ent) {  // bin entries: exact
    const u32 at = ci.x + min(rs >> 6, ci.y - 1u);
    const uint4 e = __ldg(&fx.ent[at]);
    const u64 bits = ((u64)e.y << 32) | e.x;
    return __ldg(&fx.rank[at]) + __popcll(bits & ((2ull << (rs & 63u)) - 1ull));  // (the clamped last bin is empty)
  }
  const u32 bRaw = rs >> fx.shift;
  const uint2 en = __ldg(&fx.bm[ci.x + min(bRaw, ci.y - 1u)]);
  const u32 p = (bRaw < ci.y) ? ((rs >> fx.gshift) & 31u) : 31u;
  const u32 m = (fx.gshift == 0) ? (0xFFFFFFFFu >> (31u - p)) : ((1u << p) - 1u);
  return en.y + __popc(en.x & m);
}

// A read ends d >= 1 positions after the end of its start segment; lens = len_{i+1} | len_{i+2} << 16 (saturated at 65535).
// 1: it ends inside the next segment, 2: inside the one after, 0: further away or unknown
__device__ __forceinline__ int segmentsAhead(u32 d, u32 lens) {
  const u32 len1 = lens & 0xFFFFu, len2 = lens >> 16;
  if (d <= len1 && (len1 != 65535u || d <= 65534u)) return 1;
  if (len1 == 65535u) return 0;
  const u32 d2 = d - len1;  // > 0 here
  if (d2 <= len2 && (len2 != 65535u || d2 <= 65534u)) return 2;
  return 0;
}

// the pick of an ANS_VICPAIR answer for the read [rs, re]
__device__ __forceinline__ u32 vicPick(const FastView &fx, u32 a, u32 tie, u32 rs, u32 re) {
  const u64 sum = (u64)rs + (u64)re;
  a &= ~ANS_VICPAIR;
  if (sum > (u64)tie) a &= fx.upMask;
  else if (sum < (u64)tie) a &= fx.downMask;
  return a;
}

template <int MODE>
__device__ __noinline__ u32 fastMissEval(const IndexView &ix, u32 rs, u32 re, u32 meta, float ovl) {
  return (u32)annotateEval<MODE, false>(ix, rs, re, meta, ovl, nullptr) | FAST_MISS;
}

template <int MODE>
__device__ __forceinline__ u32 fastAnnotate(const FastView &fx, const IndexView &ix, u32 rs, u32 re, u32 meta, float ovl) {
  const u32 chr = meta & 0x00FFFFFFu;
  if (chr >= fx.nChr) return 0;
  if (re < rs || re >= 0xFFFFFFF0u) return fastMissEval<MODE>(ix, rs, re, meta, ovl);
  if (MODE != 0) {  // no feature can overlap the read by more than end - start
    const u32 o = re - rs;
    if (o == 0) return 0;
    if (MODE == 1) { if (!(__fmul_rn((float)(o + 1u), ovl) <= (float)o)) return 0; }
    else { if (!((float)o >= ovl)) return 0; }
  }
  u32 i = fastSegIndex(fx, __ldg(&fx.chrInfo[chr]), rs);
  uint4 t = __ldg(&fx.seg[2u * i]);  // {end, answer F, answer R, end of the next segment}
#pragma unroll 1
  for (int g = 0; rs > t.x; ++g) {   // coarse granules only
    if (g == 3) return fastMissEval<MODE>(ix, rs, re, meta, ovl);
    ++i;
    t = __ldg(&fx.seg[2u * i]);
  }
  const bool fwd = (meta >> 31) != 0;
  u32 a;
  if (re <= t.x) {
    a = fwd ? t.y : t.z;
    if (a & ANS_VICPAIR) a = vicPick(fx, a, __ldg(&fx.tie[2u * i + (fwd ? 0u : 1u)]), rs, re);
  } else if (MODE == 0) {
    const int which = segmentsAhead(re - t.x, t.w);  // 1: ends in the next segment, 2: in the one after, 0: further
    if (which == 0) return fastMissEval<MODE>(ix, rs, re, meta, ovl);
    const uint4 x = __ldg(&fx.seg[2u * i + 1u]);  // {cross answer F, cross answer R, triple answer F, triple answer R}
    a = (which == 1) ? (fwd ? x.x : x.y) : (fwd ? x.z : x.w);
  } else {
    return fastMissEval<MODE>(ix, rs, re, meta, ovl);
  }
  if (a & ANS_GENERAL) return fastMissEval<MODE>(ix, rs, re, meta, ovl);
  return a;
}

// ----------------------------------------------------------------------------- the batch kernel

// Work decomposition: a WARP TILE is 128 consecutive hits, 4 per lane (one 16-byte load per array and lane).  Every
// warp of the grid owns one contiguous chunk of warp tiles and walks it front to back WITHOUT any block-wide barrier:
// the state of the read that is still open at the end of a tile stays in registers for the next tile, so only the read
// open at the end of a chunk (one per ~16 tiles) has to be finished by a serial walk through global memory.
#define BATCH_THREADS 256
#define BATCH_WARPS (BATCH_THREADS / 32)
#define WT_HITS 128
#ifndef MMA_BLOCKS_PER_SM
#define MMA_BLOCKS_PER_SM 3  // resident k_batch blocks per SM the register allocation is tuned for (measured: 3 > 4 > 5)
#endif
#define BT_SLOTS 256
#define HIST_ROWS 32

// rescue(), mm:497-509, on a single ambiguous hit (mm:1728 -> 491): every multiplicity is 1
__device__ __forceinline__ u64 rescueSingle(const Rules &r, u64 mask) {
  const int nreg = __popcll(mask);
  if (r.rescue && nreg > 1) {
    const u32 t = (u32)ceilf(__fmul_rn((float)nreg, r.rescueThreshold));
    if (t <= 1u) return mask & (0 - mask);  // lowest element reaches the threshold first
  }
  return mask;
}

// rescue(), mm:497-509, from per-element multiplicities gathered by `multiplicity(bit)`
template <typename F>
__device__ __forceinline__ u64 rescueFromCounts(const Rules &r, u64 gm, u32 total, F multiplicity) {
  if (total == 1) return gm;
  const u32 t = (u32)ceilf(__fmul_rn((float)total, r.rescueThreshold));
  for (u64 rest = gm; rest; rest &= rest - 1) {
    const u64 bit = rest & (0 - rest);
    if (multiplicity(bit) >= t) return bit;
  }
  return gm;
}

template <int MODE, bool FAST>
struct Annotator {
  const IndexView &ix;
  const FastView &fx;
  float ovl;
  __device__ __forceinline__ u64 operator()(u32 rs, u32 re, u32 meta) const {
    if (FAST) return fastAnnotate<MODE>(fx, ix, rs, re, meta, ovl) & ~FAST_MISS;
    return annotateHit<MODE>(ix, rs, re, meta, ovl);
  }
};

// Shared memory of one block: only the count tables
template <bool HIST>
struct BatchSmem {
  BlockTable<BT_SLOTS> bt;  // element sets of two or more elements (and everything under -y ratio / wide sets)
  // reads counted per single-element set: one private column per thread, no atomics.  A thread adds at most 4 per
  // warp tile and sees < 2^14 warp tiles of one batch: no overflow
  unsigned short hist[HIST ? HIST_ROWS : 1][BATCH_THREADS];
  u32 walkQ[4][BATCH_THREADS];  // per thread: first records of the runs of the current tile that need the serial walk
  u32 stat[ST_N];
};

// The NH countdown of Counter::addCount (mm:1669-1702) over one run of records sharing a read key, as a serial walk
// through global memory (element sets recomputed).  A read opens at a record with NH > 1, takes the following NH-1
// records of its name (NH <= 1 records are reads of their own and do not count down), then its element set is
// counted.  This is the path of every run that is not the clean "n records carrying NH = n" and of the one run per
// chunk that crosses into the next warp's chunk.  What happens to a read still open at the end of its run:
//   run ends inside the batch : the read cannot be finished here (its remaining records may come later in the file,
//                               or never): its records and key go to the deferred path, the batch is marked dirty;
//   run reaches the batch end : the open state is carried into the next batch (ctl->carry).
template <int MODE, bool FAST, typename Count>
struct RunWalker {
  const HitView &h;
  const Rules &r;
  const Annotator<MODE, FAST> &annot;
  SampleCtl *ctl;
  const SlowView &slow;
  const KeySetView &open;
  Count &count;
  u32 seq;
  u32 nReads, nRescued;

  __device__ __forceinline__ u64 maskAt(u32 j) const { return annot(h.start[j], h.end[j], h.meta[j]); }

  __device__ u64 rescueGroup(u32 first, u32 end, u64 gm) const {
    u32 total = 0;
    for (u32 j = first; j < end; ++j)
      if (h.nh[j] > 1) total += __popcll(maskAt(j));
    return rescueFromCounts(r, gm, total, [&](u64 bit) {
      u32 c = 0;
      for (u32 j = first; j < end; ++j)
        if (h.nh[j] > 1 && (maskAt(j) & bit)) ++c;
      return c;
    });
  }

  // i = first record of the run inside this batch, k = its key; `cin` = state carried into the batch (or null)
  __device__ __noinline__ void walk(u32 i, u64 k, const Carry *cin) {
    const bool routed = (ctl->openCount != 0) && keySetLookup(open, k, seq) == 1;
    bool isOpen = false, fromCarry = false;
    u32 remaining = 0, first = i;
    u64 gm = 0;
    if (cin) { isOpen = true; fromCarry = true; remaining = cin->remaining; gm = cin->gm; }
    u32 j = i;
    for (; j < h.n; ++j) {
      if (j != i && normKey(h.key[j]) != k) break;  // end of the run
      const u32 nhj = h.nh[j];
      if (!(nhj > 1)) continue;
      const u64 mj = maskAt(j);
      if (routed) { slowAppend(slow, ctl, k, ctl->ordBase + j, mj, nhj); continue; }
      if (!isOpen) { isOpen = true; fromCarry = false; remaining = nhj - 1; gm = mj; first = j; ++nReads; }
      else { --remaining; gm |= mj; }
      if (remaining == 0) {
        if (gm != 0) {
          if (r.rescue) gm = rescueGroup(first, j + 1, gm);  // never a carried read: rescue mode does not carry
          count(gm);
          if (__popcll(gm) == 1) ++nRescued;
        }
        isOpen = false;
      }
    }
    if (!isOpen) return;
    if (routed) return;  // (a carried read is never routed: its name would have been deferred instead of carried)
    if (j >= h.n && !r.rescue) {  // the run reaches the end of the batch: carry the open read over
      Carry &c = ctl->carry[(seq + 1) & 1];
      c.key = k; c.gm = gm; c.remaining = remaining;
      c.ord = fromCarry ? cin->ord : ctl->ordBase + first;
      __threadfence();
      c.valid = 1;
      return;
    }
    // unfinished read: the deferred path opens it again
    --nReads;
    if (fromCarry) slowAppend(slow, ctl, k, cin->ord, cin->gm, cin->remaining + 1);  // stands for the records of earlier batches
    for (u32 q = fromCarry ? i : first; q < j; ++q) {
      const u32 nhq = h.nh[q];
      if (nhq > 1) slowAppend(slow, ctl, k, ctl->ordBase + q, maskAt(q), nhq);
    }
    keySetInsert(open, k, seq, ctl);
    ctl->dirty = 1;
  }
};

// STRAT: MMA_STRATEGY_* (0 default, 1 unique, 2 random, 3 ratio)
template <int MODE, int STRAT, bool FAST, typename MaskT>
__global__ void __launch_bounds__(BATCH_THREADS, (sizeof(MaskT) == 4) ? MMA_BLOCKS_PER_SM : 2)
k_batch(IndexView ix, FastView fx, HitView h, Rules r, TableView table, SampleCtl *ctl, SlowView slow, KeySetView open) {
  constexpr bool HIST = (sizeof(MaskT) == 4) && STRAT != 2 && STRAT != 3;  // single-element sets go to the private histogram
  __shared__ BatchSmem<HIST> sm;
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  sm.bt.init();
  if (HIST) {
#pragma unroll
    for (int e = 0; e < HIST_ROWS; ++e) sm.hist[e][tid] = 0;
  }
  if (tid < ST_N) sm.stat[tid] = 0;
  __syncthreads();
  const Annotator<MODE, FAST> annot{ix, fx, r.overlap};
  const u32 seq = ctl->batchSeq;
  // parallel countdown only when no read name is known as unfinished (then nothing has to be routed to the deferred
  // path) and rescue() is off (it needs multiplicities); the value is uniform over the warp
  const bool parallelRuns = (STRAT == 0) && (sizeof(MaskT) == 4) && !r.rescue && (__shfl_sync(0xffffffffu, ctl->openCount, 0) == 0);
  // per-thread counters, two 16-bit fields per register (a thread sees < 2^15 hits of one batch: 4 per warp tile,
  // at most 2^32 / 128 / (gridDim.x * 8) tiles with gridDim.x >= 148 * 2 for such a batch)
  u32 pUnasAmbi = 0;   // unassigned | ambiguous << 16      (mm:1666-1667)
  u32 pUniqMult = 0;   // unique | multiple << 16           (mm:1668, 1670)
  u32 pHitsMiss = 0;   // hits looked at | segment-table misses << 16
  u32 pClosResc = 0;   // multi-mapping reads closed by the parallel countdown | of which rescued << 16

  // one read counted for the element set `ckey` (0 = nothing)
  auto count = [&](u64 ckey) {
    if (HIST) {
      const u32 c = (u32)ckey;
      if (c == 0) return;
      if (c & (c - 1)) sm.bt.add(ckey, 1, table);
      else sm.hist[__ffs(c) - 1][tid] += 1;
    } else if (ckey) {
      sm.bt.add(ckey, 1, table);
    }
  };
  RunWalker<MODE, FAST, decltype(count)> w{h, r, annot, ctl, slow, open, count, seq, 0u, 0u};

  // this warp's chunk of warp tiles
  const u32 nWT = (h.n + WT_HITS - 1) / WT_HITS;
  const u32 nWarps = gridDim.x * BATCH_WARPS;
  const u32 per = (nWT + nWarps - 1) / nWarps;
  const u32 t0 = min(nWT, (blockIdx.x * BATCH_WARPS + warp) * per), t1 = min(nWT, t0 + per);

  // the run that is open at the end of the previous tile of the chunk (all lanes hold the same values)
  bool cCont = false;   // ... continues in the next tile
  bool cValid = false;  // ... and it starts inside this chunk (else the previous chunk's warp finishes it)
  u32 cStart = 0;       // its first record
  u32 cTot = 0;         // union of its element sets so far | bit 31: NH changed inside it
  u32 cNh = 0;          // NH of the last record of the previous tile
  u64 cKey = KEY_EMPTY; // key of the last record of the previous tile

  for (u32 t = t0; t < t1; ++t) {
    const u32 base = t * WT_HITS + lane * 4;
    u32 rs[4], re[4], meta[4], nh[4];
    u64 key[4];
    bool valid[4];
    // ---- load (128-bit accesses on full tiles of aligned arrays)
    if (h.vec && (t + 1) * WT_HITS <= h.n) {
      const uint4 a = __ldcs(reinterpret_cast<const uint4 *>(h.start + base));
      const uint4 b = __ldcs(reinterpret_cast<const uint4 *>(h.end + base));
      const uint4 c = __ldcs(reinterpret_cast<const uint4 *>(h.meta + base));
      const uint4 d = __ldcs(reinterpret_cast<const uint4 *>(h.nh + base));
      rs[0] = a.x; rs[1] = a.y; rs[2] = a.z; rs[3] = a.w;
      re[0] = b.x; re[1] = b.y; re[2] = b.z; re[3] = b.w;
      meta[0] = c.x; meta[1] = c.y; meta[2] = c.z; meta[3] = c.w;
      nh[0] = d.x; nh[1] = d.y; nh[2] = d.z; nh[3] = d.w;
      if (STRAT == 0) {
        const ulonglong2 k0 = __ldcs(reinterpret_cast<const ulonglong2 *>(h.key + base));
        const ulonglong2 k1 = __ldcs(reinterpret_cast<const ulonglong2 *>(h.key + base + 2));
        key[0] = normKey(k0.x); key[1] = normKey(k0.y); key[2] = normKey(k1.x); key[3] = normKey(k1.y);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) valid[j] = true;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const u32 i = base + j;
        valid[j] = i < h.n;
        rs[j] = 0; re[j] = 0; meta[j] = 0x00FFFFFFu; nh[j] = 1; key[j] = KEY_EMPTY;
        if (valid[j]) {
          rs[j] = __ldcs(&h.start[i]); re[j] = __ldcs(&h.end[i]); meta[j] = __ldcs(&h.meta[i]); nh[j] = __ldcs(&h.nh[i]);
          if (STRAT == 0) key[j] = normKey(__ldcs(&h.key[i]));
        }
      }
    }
    // ---- run starts (the keys are not needed after this)
    u32 hbits = 0, F = 0;
    const Carry *carryIn = nullptr;
    u64 nextKey = KEY_EMPTY;
    u32 tileEndsRun = 1;  // lane 31: the record after the tile's last one starts another run (or the batch ends there)
    if (STRAT == 0) {
      const u32 nextTile = (t + 1) * WT_HITS;
      if (lane == 31u && nextTile < h.n) tileEndsRun = (normKey(__ldg(&h.key[nextTile])) != key[3]) ? 1u : 0u;
      u64 prev = __shfl_up_sync(0xffffffffu, key[3], 1);
      if (lane == 0) {
        if (t != t0) prev = cKey;
        else if (base == 0) {
          const Carry &c = ctl->carry[seq & 1];
          prev = KEY_EMPTY;
          if (c.valid) { carryIn = &c; prev = c.key; }
        } else prev = normKey(h.key[base - 1]);
      }
      // a record whose read key differs from the previous record's starts a run
      hbits = ((key[0] != prev || !valid[0]) ? 1u : 0u) | ((key[1] != key[0] || !valid[1]) ? 2u : 0u) |
              ((key[2] != key[1] || !valid[2]) ? 4u : 0u) | ((key[3] != key[2] || !valid[3]) ? 8u : 0u);
      F = __ballot_sync(0xffffffffu, hbits != 0);  // lanes in which a run starts
      nextKey = __shfl_sync(0xffffffffu, key[3], 31);
    }
    // ---- element set of every hit that is looked at (unique: only NH == 1 is, mm:1773)
    bool visited[4];
    MaskT m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      visited[j] = valid[j] && !(STRAT == 1 && nh[j] != 1);
      m[j] = 0;
      if (!visited[j]) continue;
      if (FAST) {
        const u32 a = fastAnnotate<MODE>(fx, ix, rs[j], re[j], meta[j], r.overlap);
        m[j] = (MaskT)(a & ~FAST_MISS);
        pHitsMiss += (a >> 31) << 16;
      } else {
        m[j] = (MaskT)annotateHit<MODE>(ix, rs[j], re[j], meta[j], r.overlap);
      }
    }
    // ---- per-hit counters (mm:1666-1668) and the reads that are their own group (mm:1703-1738)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nreg = __popcll((u64)m[j]);
      const bool multi = visited[j] && STRAT == 0 && nh[j] > 1;  // joins the by-name countdown (mm:1669)
      pHitsMiss += visited[j] ? 1u : 0u;  // a hit that is not multi is a read of its own (mm:1737)
      pUnasAmbi += (visited[j] && nreg == 0 ? 1u : 0u) | (visited[j] && nreg > 1 ? 0x10000u : 0u);
      pUniqMult += (visited[j] && nreg == 1 && nh[j] == 1 ? 1u : 0u) | (multi ? 0x10000u : 0u);
      if (visited[j] && !multi && m[j] != 0) {
        if (STRAT == 2) {
          slowAppend(slow, ctl, normKey(h.key[base + j]), ctl->ordBase + base + j, (u64)m[j], nh[j]);  // order-dependent draw: deferred
        } else {
          u64 ckey = rescueSingle(r, (u64)m[j]);
          if (STRAT == 3) {
            if (nh[j] >= (1u << (64 - NH_SHIFT))) atomicExch(&ctl->overflow, 1u);
            ckey |= (u64)nh[j] << NH_SHIFT;
          }
          count(ckey);
        }
      }
    }
    if (STRAT != 0) continue;

    // ---- per-read countdown (mm:1669-1702)
    if (carryIn) {  // lane 0 of the batch's first tile: the read carried into this batch
      if (!(hbits & 1u)) w.walk(0, carryIn->key, carryIn);
      else {  // its name does not continue: unfinished
        --w.nReads;
        slowAppend(slow, ctl, carryIn->key, carryIn->ord, carryIn->gm, carryIn->remaining + 1);
        keySetInsert(open, carryIn->key, seq, ctl);
        ctl->dirty = 1;
      }
    }
    u32 nWalk = 0;
    if (!parallelRuns) {
      // serial walk of every run from its first record (rescue mode, wide element sets, or some read name is unfinished)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (valid[j] && ((hbits >> j) & 1u)) sm.walkQ[nWalk++][tid] = base + j;
#pragma unroll 1
      for (u32 q = 0; q < nWalk; ++q) { const u32 i0 = sm.walkQ[q][tid]; w.walk(i0, normKey(h.key[i0]), nullptr); }
      cKey = nextKey;  // the next tile of the chunk compares its first record with this tile's last one
      continue;
    }
    // Parallel countdown for the regular case: a run of n records that all carry NH = n (> 1) is one read; its
    // element set is the union over the run: a segmented OR scan over the 128 hits of the warp tile (bit 31 of the
    // scanned word = "NH changes inside the run"), seeded with the state carried from the previous tile.  The lane
    // owning the run's LAST record closes it.  Runs of another shape take the serial walk (RunWalker), started by the
    // same lane.
    u32 prevNh = __shfl_up_sync(0xffffffffu, nh[3], 1);
    if (lane == 0) prevNh = cNh;
    u32 pre[4], acc = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool isHead = (hbits >> j) & 1u;
      const bool bad = !isHead && nh[j] != (j ? nh[j > 0 ? j - 1 : 0] : prevNh);
      const u32 x = (u32)m[j] | (bad ? 0x80000000u : 0u);
      acc = isHead ? x : (acc | x);
      pre[j] = acc;
    }
    u32 inc = acc;  // inclusive scan over lanes: OR of `acc` from the nearest lane with a run start up to this lane
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u32 tt = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= (u32)d && ((F >> (lane - d + 1)) & ((1u << d) - 1u)) == 0) inc |= tt;
    }
    u32 X = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) X = 0;
    const u32 before = F & ((1u << lane) - 1u);
    const u32 lastHeadPos = base + (31 - __clz(hbits | 1u));
    const u32 sPrev = __shfl_sync(0xffffffffu, lastHeadPos, before ? (31 - __clz(before)) : 0);
    const u32 nextHead0 = __shfl_down_sync(0xffffffffu, hbits & 1u, 1);
    // bit j: the next record starts another run, i.e. this record ends its run (the tile's last record: from the peek)
    const u32 lastBits = (hbits >> 1) | ((lane < 31u ? nextHead0 : tileEndsRun) << 3);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!(valid[j] && ((lastBits >> j) & 1u))) continue;
      const u32 hbLe = hbits & ((2u << j) - 1u);
      u32 tot, runStart;
      if (hbLe) { tot = pre[j]; runStart = base + (31 - __clz(hbLe)); }
      else if (before) { tot = X | pre[j]; runStart = sPrev; }
      else if (cValid) { tot = cTot | X | pre[j]; runStart = cStart; }
      else continue;  // the run starts in another warp's chunk: that warp finishes it
      if (tot & 0x80000000u) { sm.walkQ[nWalk++][tid] = runStart; continue; }
      if (!(nh[j] > 1)) continue;  // a run of reads that are their own group
      if (nh[j] != base + j + 1 - runStart) { sm.walkQ[nWalk++][tid] = runStart; continue; }
      const u32 gm = tot & 0x7FFFFFFFu;
      pClosResc += 1u | (__popc(gm) == 1 ? 0x10000u : 0u);
      count(gm);
    }
#pragma unroll 1
    for (u32 q = 0; q < nWalk; ++q) { const u32 i0 = sm.walkQ[q][tid]; w.walk(i0, normKey(h.key[i0]), nullptr); }
    // the run still open at the end of the tile
    {
      const u32 incLast = __shfl_sync(0xffffffffu, inc, 31);
      if (F) {
        cTot = incLast;
        cStart = __shfl_sync(0xffffffffu, lastHeadPos, 31 - __clz(F));
        cValid = true;
      } else if (cValid) {
        cTot |= incLast;
      }
      cNh = __shfl_sync(0xffffffffu, nh[3], 31);
      cKey = nextKey;
      cCont = __shfl_sync(0xffffffffu, tileEndsRun, 31) == 0;
    }
  }
  // ---- the read open at the end of the chunk continues in another warp's chunk: finish it by the serial walk.  (At the
  //      end of the batch the tile's last record closed its run above.)
  if (STRAT == 0 && parallelRuns && cValid && cCont && t1 > t0 && lane == 0) w.walk(cStart, cKey, nullptr);
  u32 cHits = pHitsMiss & 0xFFFFu, cMiss = pHitsMiss >> 16, cUnassigned = pUnasAmbi & 0xFFFFu, cAmbiguous = pUnasAmbi >> 16;
  u32 cUnique = pUniqMult & 0xFFFFu, cMultiple = pUniqMult >> 16;
  u32 cReads = cHits - cMultiple + (pClosResc & 0xFFFFu) + w.nReads, cRescued = (pClosResc >> 16) + w.nRescued;

  // ---- block epilogue: counters, the private histogram columns and the private table
  cHits = __reduce_add_sync(0xffffffffu, cHits); cUnassigned = __reduce_add_sync(0xffffffffu, cUnassigned);
  cAmbiguous = __reduce_add_sync(0xffffffffu, cAmbiguous); cUnique = __reduce_add_sync(0xffffffffu, cUnique);
  cMultiple = __reduce_add_sync(0xffffffffu, cMultiple); cReads = __reduce_add_sync(0xffffffffu, cReads);
  cRescued = __reduce_add_sync(0xffffffffu, cRescued); cMiss = __reduce_add_sync(0xffffffffu, cMiss);
  if (lane == 0) {
    atomicAdd(&sm.stat[ST_HITS], cHits); atomicAdd(&sm.stat[ST_UNASSIGNED], cUnassigned); atomicAdd(&sm.stat[ST_AMBIGUOUS], cAmbiguous);
    atomicAdd(&sm.stat[ST_UNIQUE], cUnique); atomicAdd(&sm.stat[ST_MULTIPLE], cMultiple); atomicAdd(&sm.stat[ST_READS], cReads);
    atomicAdd(&sm.stat[ST_RESCUED], cRescued); atomicAdd(&sm.stat[7], cMiss);
  }
  __syncthreads();
  sm.bt.flush(table);
  if (HIST) {
    for (u32 e = warp; e < HIST_ROWS; e += BATCH_WARPS) {
      u32 v = 0;
#pragma unroll
      for (int q = 0; q < BATCH_WARPS; ++q) v += sm.hist[e][lane + 32 * q];
      v = __reduce_add_sync(0xffffffffu, v);
      if (lane == 0 && v) tableAdd(table, 1ull << e, v);
    }
  }
  if (tid < 7) {
    const int sv = (int)sm.stat[tid];  // reads / rescued can be negative within a block (unfinished reads)
    if (sv) atomicAdd(&ctl->stats[tid], (u64)(long long)sv);
  }
  if (tid == 7 && sm.stat[7]) atomicAdd(&ctl->fastMiss, sm.stat[7]);
}

// IntervalList::scan alone (mm:1291-1332): the element set of every hit, nothing counted
template <int MODE, bool FAST>
__global__ void __launch_bounds__(256)
k_annotate_only(IndexView ix, FastView fx, HitView h, Rules r, u64 *__restrict__ out) {
  const Annotator<MODE, FAST> annot{ix, fx, r.overlap};
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < h.n; i += gridDim.x * blockDim.x)
    out[i] = annot(h.start[i], h.end[i], h.meta[i]);
}

// -M (mm:1077-1081, 1328-1330): the intervals behind the element set of a hit = every candidate of a chosen element that
// passes the strand rule and scores > 0 (EvaluationStructure::set appends each of them to ids[type], mm:1023-1028).  The
// reference concatenates them element by element, but every consumer sorts the list (mm:1693, 1732, 1795), so they are
// emitted here in candidate order.  `emit(featureIndex)` is called once per such interval; returns their number.
template <int MODE, typename Emit>
__device__ __forceinline__ u32 chosenIntervals(const IndexView &ix, u32 rs, u32 re, u32 meta, float ovl, u64 chosen, Emit emit) {
  const u32 chr = meta & 0x00FFFFFFu;
  if (chosen == 0 || chr >= ix.nChr) return 0;
  const uint2 ci = __ldg(&ix.chrInfo[chr]);
  const u32 lastBin = ci.y - 1;
  const u32 qlo = min(rs, re), qhi = max(rs, re);
  const u32 b0 = min(qlo >> ix.shift, lastBin), b1 = min(qhi >> ix.shift, lastBin);
  const uint2 e0 = __ldg(&ix.bins[ci.x + b0]);
  const uint2 e1 = __ldg(&ix.bins[ci.x + b0 + 1]);
  const u32 hi = (b1 == b0) ? e1.x : __ldg(&ix.bins[ci.x + b1 + 1]).x;
  const u32 rstrand = meta >> 31;
  u32 n = 0;
  auto visit = [&](u32 i) {
    const uint4 f = __ldg(&ix.feat[i]);
    if (f.x > re || f.w < rs) return;  // outside the reference's walk (see evalCandidate)
    const u32 m = f.z;
    if (!((chosen >> FM_TYPE(m)) & 1ull)) return;
    const u32 es = FM_ESTRAND(m);
    if (es != 0 && ((es == 1) != (FM_FWD(m) == rstrand))) return;  // Config::checkStrand, mm:438-443
    u32 sc;
    if (MODE == 0) {
      sc = (rs >= f.x && re <= f.y) ? 1u : 0u;
    } else {
      const u32 s = max(f.x, rs), e = min(f.y, re);
      const u32 o = (s >= e) ? 0u : e - s;
      if (MODE == 1) sc = (__fmul_rn((float)(re - rs + 1u), ovl) <= (float)o) ? o : 0u;
      else sc = ((float)o >= ovl) ? o : 0u;
    }
    if (sc == 0) return;
    emit(i);
    ++n;
  };
  for (u32 k = e0.y; k < e1.y; ++k) visit(__ldg(&ix.spanIdx[k]));  // spanning features come first in feature order
  for (u32 i = e0.x; i < hi; ++i) visit(i);
  return n;
}

// pass 1: element set and number of intervals per hit; pass 2 (after an exclusive scan of the counts): the interval ids
template <int MODE>
__global__ void __launch_bounds__(256)
k_intervals_count(IndexView ix, HitView h, Rules r, u64 *__restrict__ masks, u32 *__restrict__ counts) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < h.n; i += gridDim.x * blockDim.x) {
    const u32 rs = h.start[i], re = h.end[i], meta = h.meta[i];
    const u64 c = annotateEval<MODE, false>(ix, rs, re, meta, r.overlap, nullptr);
    masks[i] = c;
    counts[i] = chosenIntervals<MODE>(ix, rs, re, meta, r.overlap, c, [](u32) {});
  }
}
template <int MODE>
__global__ void __launch_bounds__(256)
k_intervals_fill(IndexView ix, HitView h, Rules r, const u64 *__restrict__ masks, const u64 *__restrict__ offsets, u32 *__restrict__ ids) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < h.n; i += gridDim.x * blockDim.x) {
    u64 at = offsets[i];
    chosenIntervals<MODE>(ix, h.start[i], h.end[i], h.meta[i], r.overlap, masks[i], [&](u32 f) { ids[at++] = f; });
  }
}
__global__ void k_widen_counts(const u32 *__restrict__ in, u64 *__restrict__ out, u32 n) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// Compact transfer format -> the five hit arrays (see mma_packed_batch in mmannot_b200.h).  One block per tile of
// PACK_TILE hits, 4 consecutive hits per thread; the read key of a hit is the key of its run, found from the number of run
// starts up to the hit (tile base from the host + a block-wide prefix sum of the run-start bits).
#define PACK_TILE 1024
struct PackedView {
  const u32 *start, *packed, *tileRunBase, *escIndex, *escEnd, *escNh;
  const u64 *runKey;
  u32 n, nEsc, nRuns;
};
__global__ void __launch_bounds__(PACK_TILE / 4)
k_expand_packed(PackedView pv, u32 *__restrict__ end, u32 *__restrict__ meta, u32 *__restrict__ nh, u64 *__restrict__ key) {
  __shared__ u32 warpSum[PACK_TILE / 4 / 32];
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const u32 base = blockIdx.x * PACK_TILE + tid * 4;
  u32 p[4], st[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const u32 i = base + j;
    p[j] = (i < pv.n) ? pv.packed[i] : 0u;
    st[j] = (i < pv.n) ? pv.start[i] : 0u;
  }
  const u32 mine = ((p[0] >> 30) & 1u) + ((p[1] >> 30) & 1u) + ((p[2] >> 30) & 1u) + ((p[3] >> 30) & 1u);
  u32 inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (u32)d) inc += o; }
  if (lane == 31) warpSum[warp] = inc;
  __syncthreads();
  u32 before = pv.tileRunBase[blockIdx.x] + inc - mine;
  for (u32 w = 0; w < warp; ++w) before += warpSum[w];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const u32 i = base + j;
    if (i >= pv.n) break;
    before += (p[j] >> 30) & 1u;
    u32 len = p[j] & 255u, n = (p[j] >> 8) & 255u;
    u32 e = st[j] + len - 1u;
    if (len == 255u || n == 255u) {  // escaped: full values by binary search of the hit index
      u32 lo = 0, hi = pv.nEsc;
      while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (pv.escIndex[mid] < i) lo = mid + 1; else hi = mid; }
      if (lo < pv.nEsc && pv.escIndex[lo] == i) { e = pv.escEnd[lo]; n = pv.escNh[lo]; }
    }
    const u32 chr = (p[j] >> 16) & 0x3FFFu;
    end[i] = e;
    nh[i] = n;
    meta[i] = (chr == 0x3FFFu ? 0x00FFFFFFu : chr) | (p[j] & 0x80000000u);
    // (the struct is the host's word: run-start bits that disagree with tileRunBase / nRuns give wrong keys, never a read
    // outside runKey -- `before` = 0 wraps and is clamped as well)
    key[i] = pv.runKey[min(before - 1u, pv.nRuns - 1u)];
  }
}

// End of batch.  When the batch is dirty, every run (of this batch) of a read name that became unfinished DURING the
// batch was resolved by k_batch as if the name had no open read; such runs are walked again here, what k_batch counted
// for them is taken back and their multi-mapping records are handed to the deferred path, which replays the name's
// records in file order.  The last block to finish advances the batch state.
template <int MODE, bool FAST>
__global__ void __launch_bounds__(256)
k_batch_close(IndexView ix, FastView fx, HitView h, Rules r, TableView table, SampleCtl *ctl, SlowView slow, KeySetView open) {
  const Annotator<MODE, FAST> annot{ix, fx, r.overlap};
  const u32 seq = ctl->batchSeq;
  if (ctl->dirty) {
    const Carry cin = ctl->carry[seq & 1];
    Carry &cout = ctl->carry[(seq + 1) & 1];
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < h.n; i += gridDim.x * blockDim.x) {
      const u64 k = normKey(h.key[i]);
      const bool continues = (i == 0) && cin.valid && cin.key == k;
      if (i > 0 && normKey(h.key[i - 1]) == k) continue;  // not the first record of a run
      if (keySetLookup(open, k, seq) != 2) continue;      // only names that became unfinished during this batch
      // pass A: replay what k_batch did with this run (it was not routed: the name was not known as unfinished)
      bool isOpen = continues, fromCarry = continues;
      u32 remaining = continues ? cin.remaining : 0, first = i;
      u64 gm = continues ? cin.gm : 0;
      long long dReads = 0, dRescued = 0;
      u32 j = i;
      for (; j < h.n && (j == i || normKey(h.key[j]) == k); ++j) {
        const u32 nhj = h.nh[j];
        if (!(nhj > 1)) continue;
        const u64 mj = annot(h.start[j], h.end[j], h.meta[j]);
        if (!isOpen) { isOpen = true; fromCarry = false; remaining = nhj - 1; gm = mj; first = j; }
        else { --remaining; gm |= mj; }
        if (remaining == 0) {
          if (gm != 0) {
            u64 g = gm;
            if (r.rescue) {
              u32 total = 0;
              for (u32 q = first; q <= j; ++q)
                if (h.nh[q] > 1) total += __popcll(annot(h.start[q], h.end[q], h.meta[q]));
              g = rescueFromCounts(r, gm, total, [&](u64 bit) {
                u32 c = 0;
                for (u32 q = first; q <= j; ++q)
                  if (h.nh[q] > 1 && (annot(h.start[q], h.end[q], h.meta[q]) & bit)) ++c;
                return c;
              });
            }
            tableAdd(table, g, (u64)(-1ll));  // take the count back
            if (__popcll(g) == 1) --dRescued;
          }
          --dReads;
          isOpen = false;
        }
      }
      // the read still open at the end of the run: carried out (not yet deferred) or already deferred by k_batch
      u32 appendEnd = j;
      bool synth = continues;
      if (isOpen) {
        if (j >= h.n && !r.rescue) {  // k_batch carried it out of the batch: defer it instead
          cout.valid = 0;
          --dReads;
        } else {  // k_batch deferred records [first, j) (and the carried part when the read came from the carry)
          appendEnd = fromCarry ? i : first;
          if (fromCarry) synth = false;
        }
      }
      // pass B: defer every multi-mapping record of the run that k_batch did not
      if (synth) slowAppend(slow, ctl, k, cin.ord, cin.gm, cin.remaining + 1);
      for (u32 q = i; q < appendEnd; ++q) {
        const u32 nhq = h.nh[q];
        if (nhq > 1) slowAppend(slow, ctl, k, ctl->ordBase + q, annot(h.start[q], h.end[q], h.meta[q]), nhq);
      }
      if (dReads) atomicAdd(&ctl->stats[ST_READS], (u64)dReads);
      if (dRescued) atomicAdd(&ctl->stats[ST_RESCUED], (u64)dRescued);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const u32 done = atomicAdd(&ctl->closeCount, 1u) + 1;
    if (done == gridDim.x) {
      __threadfence();
      ctl->carry[seq & 1].valid = 0;  // consumed by this batch
      ctl->ordBase += h.n;
      ctl->dirty = 0;
      ctl->closeCount = 0;
      __threadfence();
      ctl->batchSeq = seq + 1;
    }
  }
}

// end of sample: a read still carried becomes a deferred record (it is flushed with the other open reads, mm:1783-1792)
__global__ void k_flush_carry(SampleCtl *ctl, SlowView slow) {
  Carry &c = ctl->carry[ctl->batchSeq & 1];
  if (!c.valid) return;
  slowAppend(slow, ctl, c.key, c.ord, c.gm, c.remaining + 1);
  ctl->stats[ST_READS] -= 1;  // the deferred path counts the read again when it opens it
  c.valid = 0;
}

// ----------------------------------------------------------------------------- end of sample: deferred records

// sorted by (key, ordinal): perm[p] = index into the deferred arrays.  One thread per key segment.
// default: countdown over all the records of the name in file order + end-of-file flush (mm:1783-1792).
__global__ void k_slow_default(const u32 *__restrict__ perm, u32 n, SlowView s, Rules r, TableView table, SampleCtl *ctl) {
  const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const u64 k = s.key[perm[p]];
  if (p > 0 && s.key[perm[p - 1]] == k) return;
  bool isOpen = false;
  u32 remaining = 0, first = 0, nReads = 0, nRescued = 0;
  u64 gm = 0;
  u32 q = p;
  for (; q < n; ++q) {
    const u32 id = perm[q];
    if (s.key[id] != k) break;
    const u64 mq = s.mask[id];
    if (!isOpen) { isOpen = true; remaining = s.nh[id] - 1; gm = mq; first = q; ++nReads; }
    else { --remaining; gm |= mq; }
    if (remaining == 0) {
      if (gm != 0) {
        if (r.rescue) {
          u32 cnt = 0;
          for (u32 z = first; z <= q; ++z) cnt += __popcll(s.mask[perm[z]]);
          gm = rescueFromCounts(r, gm, cnt, [&](u64 bit) {
            u32 c = 0;
            for (u32 z = first; z <= q; ++z) if (s.mask[perm[z]] & bit) ++c;
            return c;
          });
        }
        tableAdd(table, gm, 1);
        if (__popcll(gm) == 1) ++nRescued;
      }
      isOpen = false;
    }
  }
  if (isOpen && gm != 0) {  // flush at end of file
    if (r.rescue) {
      u32 cnt = 0;
      for (u32 z = first; z < q; ++z) cnt += __popcll(s.mask[perm[z]]);
      gm = rescueFromCounts(r, gm, cnt, [&](u64 bit) {
        u32 c = 0;
        for (u32 z = first; z < q; ++z) if (s.mask[perm[z]] & bit) ++c;
        return c;
      });
    }
    tableAdd(table, gm, 1);
    if (__popcll(gm) == 1) ++nRescued;
  }
  if (nReads) atomicAdd(&ctl->stats[ST_READS], (u64)nReads);
  if (nRescued) atomicAdd(&ctl->stats[ST_RESCUED], (u64)nRescued);
}

// The same for a list sorted by key ONLY (one radix sort instead of two): the records of a name are taken in ordinal order by
// repeated selection inside the name's segment (a name has at most a few hundred records; k_slow_maxlen tells the host when that
// does not hold and the fully sorted route has to be taken).
// (`s` holds the list in key order: gathered through the permutation of the sort, so that a name's records are adjacent in memory)
__global__ void k_slow_gather(const u32 *__restrict__ perm, u32 n, SlowView from, SlowView to) {
  const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const u32 id = perm[p];
  to.ord[p] = from.ord[id]; to.mask[p] = from.mask[id]; to.nh[p] = from.nh[id];
}
__global__ void k_slow_maxlen(u32 n, SlowView s, u32 *maxLen) {
  const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const u64 k = s.key[p];
  if (p > 0 && s.key[p - 1] == k) return;
  u32 q = p + 1;
  while (q < n && s.key[q] == k) ++q;
  if (q - p > 64) atomicMax(maxLen, q - p);
}
// One thread per read name (the thread of its first record in key order).  Counts go through a block-private table and the
// counters through shared memory: with one global atomic per name on a handful of addresses this pass was bound by same-address
// atomics (8 ms per 16 M records).
__global__ void __launch_bounds__(256) k_slow_default_byord(u32 n, SlowView s, Rules r, TableView table, SampleCtl *ctl) {
  __shared__ BlockTable<1024> bt;
  __shared__ u32 shStat[4];  // reads, rescued, names, names whose records were adjacent in the file
  bt.init();
  if (threadIdx.x < 4) shStat[threadIdx.x] = 0;
  __syncthreads();
  const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  u32 nReads = 0, nRescued = 0, nSegs = 0, nContig = 0;
  if (p < n && (p == 0 || s.key[p - 1] != s.key[p])) {
    const u64 k = s.key[p];
    u32 q = p + 1;
    u64 lo = s.ord[p], hi = lo;
    const u32 nh0 = s.nh[p];
    u64 orAll = s.mask[p];
    bool sameNh = true;
    for (; q < n; ++q) {
      const u32 id = q;
      if (s.key[id] != k) break;
      const u64 o = s.ord[id];
      lo = min(lo, o); hi = max(hi, o);
      orAll |= s.mask[id];
      sameNh &= s.nh[id] == nh0;
    }
    const u32 len = q - p;
    nSegs = 1;
    nContig = (hi - lo + 1 == (u64)len) ? 1u : 0u;
    if (sameNh && nh0 == len && !r.rescue) {
      // the usual shape -- NH records, all carrying that NH: one read whatever the order of its records (the countdown opens at
      // the first and closes at the last of them), its element set the union over all of them
      nReads = 1;
      if (orAll != 0) {
        bt.add(orAll, 1, table);
        if (__popcll(orAll) == 1) nRescued = 1;
      }
    } else {
      bool isOpen = false;
      u32 remaining = 0;
      u64 gm = 0, firstOrd = 0, last = 0;
      auto rescued = [&](u64 g, u64 o0, u64 o1) {  // rescue() over the records of the read: ordinals o0..o1 of this name
        u32 cnt = 0;
        for (u32 z = p; z < q; ++z) { const u64 o = s.ord[z]; if (o >= o0 && o <= o1) cnt += __popcll(s.mask[z]); }
        return rescueFromCounts(r, g, cnt, [&](u64 bit) {
          u32 c = 0;
          for (u32 z = p; z < q; ++z) { const u64 o = s.ord[z]; if (o >= o0 && o <= o1 && (s.mask[z] & bit)) ++c; }
          return c;
        });
      };
      for (u32 step = 0; step < len; ++step) {
        // the record with the smallest ordinal not taken yet (ordinals are distinct)
        u32 best = 0;
        u64 bestOrd = ~0ull;
        for (u32 z = p; z < q; ++z) {
          const u64 o = s.ord[z];
          if ((step == 0 || o > last) && o < bestOrd) { bestOrd = o; best = z; }
        }
        last = bestOrd;
        const u64 mq = s.mask[best];
        if (!isOpen) { isOpen = true; remaining = s.nh[best] - 1; gm = mq; firstOrd = bestOrd; ++nReads; }
        else { --remaining; gm |= mq; }
        if (remaining == 0) {
          if (gm != 0) {
            if (r.rescue) gm = rescued(gm, firstOrd, bestOrd);
            bt.add(gm, 1, table);
            if (__popcll(gm) == 1) ++nRescued;
          }
          isOpen = false;
        }
      }
      if (isOpen && gm != 0) {  // flush at end of file
        if (r.rescue) gm = rescued(gm, firstOrd, last);
        bt.add(gm, 1, table);
        if (__popcll(gm) == 1) ++nRescued;
      }
    }
  }
  nReads = __reduce_add_sync(0xffffffffu, nReads); nRescued = __reduce_add_sync(0xffffffffu, nRescued);
  nSegs = __reduce_add_sync(0xffffffffu, nSegs); nContig = __reduce_add_sync(0xffffffffu, nContig);
  if ((threadIdx.x & 31u) == 0) {
    if (nReads) atomicAdd(&shStat[0], nReads);
    if (nRescued) atomicAdd(&shStat[1], nRescued);
    if (nSegs) atomicAdd(&shStat[2], nSegs);
    if (nContig) atomicAdd(&shStat[3], nContig);
  }
  __syncthreads();
  bt.flush(table);
  if (threadIdx.x == 0 && shStat[0]) atomicAdd(&ctl->stats[ST_READS], (u64)shStat[0]);
  if (threadIdx.x == 1 && shStat[1]) atomicAdd(&ctl->stats[ST_RESCUED], (u64)shStat[1]);
  if (threadIdx.x == 2 && shStat[2]) atomicAdd(&ctl->deferSegs, shStat[2]);
  if (threadIdx.x == 3 && shStat[3]) atomicAdd(&ctl->deferContig, shStat[3]);
}

// random (mm:1706-1726).  A name draws i = rand() % NH at its first annotated hit and its i-th annotated hit (0-based, counted
// over ALL the input files of the run: Counter::clear keeps seen / chosenId / numberSeen, mm:1742-1747) is the one counted;
// afterwards the name is "seen" and every further hit of it is ignored.  The state that outlives an input file sits in a device
// map read key -> RND_SEEN | (chosen index << 32 | annotated hits so far), kept only when the context has several samples.
#define RND_SEEN 0xFFFFFFFFFFFFFFFFull
struct RndMap {
  u64 *keys, *vals;  // open addressing, KEY_EMPTY = free
  u32 capMask;       // 0 with keys == nullptr: no map (single input file)
};
__device__ __forceinline__ bool rndFind(const RndMap &m, u64 key, u64 &val) {
  if (!m.keys) return false;
  u32 slot = (u32)mix64(key) & m.capMask;
  for (u32 probe = 0; probe <= m.capMask; ++probe) {
    const u64 k = m.keys[slot];
    if (k == key) { val = m.vals[slot]; return true; }
    if (k == KEY_EMPTY) return false;
    slot = (slot + 1) & m.capMask;
  }
  return false;
}
__device__ __forceinline__ void rndStore(const RndMap &m, u64 key, u64 val) {  // (one thread per key: no two writers of a key)
  if (!m.keys) return;
  u32 slot = (u32)mix64(key) & m.capMask;
  for (u32 probe = 0; probe <= m.capMask; ++probe) {
    u64 k = m.keys[slot];
    if (k == KEY_EMPTY) k = atomicCAS(&m.keys[slot], KEY_EMPTY, key), k = (k == KEY_EMPTY) ? key : k;
    if (k == key) { m.vals[slot] = val; return; }
    slot = (slot + 1) & m.capMask;
  }
}
__global__ void k_rnd_rehash(RndMap from, RndMap to) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > from.capMask) return;
  const u64 k = from.keys[i];
  if (k != KEY_EMPTY) rndStore(to, k, from.vals[i]);
}
// ... segment heads of names that have not drawn yet publish the ordinal of the name's first annotated hit ...
__global__ void k_slow_random_heads(const u32 *__restrict__ perm, u32 n, SlowView s, RndMap map, u64 *headOrd, u32 *nHeads) {
  const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const u64 k = s.key[perm[p]];
  if (p > 0 && s.key[perm[p - 1]] == k) return;
  u64 state;
  if (rndFind(map, k, state)) return;  // seen, or drew in an earlier file
  headOrd[atomicAdd(nHeads, 1u)] = s.ord[perm[p]];
}
// ... and once those ordinals are sorted, the rank of a name's first annotated hit is the index of its
// rand() draw; the drawn-th annotated hit of the name (if the name has that many) is the one counted.
__global__ void k_slow_random_pick(const u32 *__restrict__ perm, u32 n, SlowView s, const u64 *__restrict__ sortedHeadOrd, u32 nHeads,
                                   const u32 *__restrict__ randStream, Rules r, TableView table, RndMap map) {
  const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const u32 id0 = perm[p];
  const u64 k = s.key[id0];
  if (p > 0 && s.key[perm[p - 1]] == k) return;
  u32 len = 1;  // annotated hits of the name in this file
  while (p + len < n && s.key[perm[p + len]] == k) ++len;
  u64 state;
  u64 chosen, before = 0;  // index of the hit to count; annotated hits of the name in earlier files
  if (rndFind(map, k, state)) {
    if (state == RND_SEEN) return;
    chosen = state >> 32; before = state & 0xFFFFFFFFull;
  } else {
    const u64 ord0 = s.ord[id0];
    u32 lo = 0, hi = nHeads;
    while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (sortedHeadOrd[mid] < ord0) lo = mid + 1; else hi = mid; }
    const u32 nh0 = s.nh[id0];
    chosen = nh0 ? randStream[lo] % nh0 : 0u;
  }
  if (chosen >= before && chosen - before < len) {
    tableAdd(table, rescueSingle(r, s.mask[perm[p + (u32)(chosen - before)]]), 1);
    rndStore(map, k, RND_SEEN);
  } else {
    rndStore(map, k, (chosen << 32) | ((before + len) & 0xFFFFFFFFull));
  }
}

__global__ void k_gather_keys(const u32 *__restrict__ perm, u32 n, const u64 *__restrict__ src, u64 *dst) {
  const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) dst[p] = src[perm[p]];
}
__global__ void k_iota(u32 *dst, u32 n) {
  const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) dst[p] = p;
}
__global__ void k_fill_u64(u64 *dst, u64 v, u64 n) {
  const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) dst[p] = v;
}
__global__ void k_fill_u32(u32 *dst, u32 v, u64 n) {
  const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) dst[p] = v;
}

// end of sample: the non-empty rows of the combination table, compacted (any order), with the control block in front:
// one small device->host copy instead of the whole table.  out = [SampleCtl | row count | rows {key, count}]
struct alignas(16) TableDump {  // (the rows behind it are read and written 16 bytes at a time)
  SampleCtl ctl;
  u64 nRows;
};
static_assert(sizeof(TableDump) % 16 == 0, "rows follow the head at a 16-byte boundary");
__global__ void k_table_compact(TableView t, const SampleCtl *ctl, TableDump *head, ulonglong2 *rows, u32 cap) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) head->ctl = *ctl;
  if (i > t.capMask) return;
  const u64 k = t.keys[i], v = t.vals[i];
  if (k == 0 || v == 0) return;  // a count taken back by k_batch_close can leave an empty row
  const u32 at = (u32)atomicAdd(&head->nRows, 1ull);
  if (at < cap) rows[at] = make_ulonglong2(k, v);
}

// multi-GPU merge on the device: the compacted tables of all ranks ([TableDump | rows] each, `stride` bytes apart, as left by
// an all-gather) are added into this sample's (cleared) table and counters
__global__ void k_table_import(TableView t, SampleCtl *ctl, const char *bufs, u32 nBufs, size_t stride, u32 cap) {
  const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u32 b = (u32)(idx / cap), i = (u32)(idx % cap);
  if (b >= nBufs) return;
  const TableDump *head = reinterpret_cast<const TableDump *>(bufs + (size_t)b * stride);
  const ulonglong2 *rows = reinterpret_cast<const ulonglong2 *>(bufs + (size_t)b * stride + sizeof(TableDump));
  if (i < ST_N && head->ctl.stats[i]) atomicAdd(&ctl->stats[i], head->ctl.stats[i]);
  if (i == 0 && ((head->ctl.overflow & 1u) || head->nRows > cap)) atomicOr(&ctl->overflow, 1u);  // (a dump cut short by the exchange counts as an overflow)
  if (i == 0 && head->ctl.slowCount) atomicOr(&ctl->overflow, 2u);  // a shard exported without its deferred records resolved (mma_export_table_async)
  if (i < head->nRows) { const ulonglong2 r = rows[i]; tableAdd(t, r.x, r.y); }
}

// dense read-out for the cross-GPU sum: out[i] = count of key ckey[i]
__global__ void k_dense_counts(TableView t, const u64 *__restrict__ ckey, u64 n, u64 *out) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 k = ckey[i];
  u64 v = 0;
  if (k != 0) {
    u32 slot = (u32)mix64(k) & t.capMask;
    for (u32 probe = 0; probe <= t.capMask; ++probe) {
      const u64 o = t.keys[slot];
      if (o == k) { v = t.vals[slot]; break; }
      if (o == 0) break;
      slot = (slot + 1) & t.capMask;
    }
  }
  out[i] = v;
}

// ----------------------------------------------------------------------------- K1: index build

struct BuildView {
  const u32 *chr, *start, *end;
  const uint8_t *type, *strand;
  const uint16_t *elemLine;
  const uint8_t *elemStrand, *elemVic;
  const u32 *chrStart;    // [nChr + 1]
  const u32 *chrBinBase;  // [nChr + 1] first bin entry of each chromosome (entries = bins + 1 sentinel)
  u32 nFeat, nChr, shift, nEntries;
};

__global__ void k_pack_features(BuildView b, uint4 *feat) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.nFeat) return;
  const u32 t = b.type[i];
  const u32 meta = t | ((u32)b.elemLine[t] << 6) | ((u32)b.elemStrand[t] << 22) | ((u32)b.elemVic[t] << 24) | ((b.strand[i] == 1 ? 1u : 0u) << 26);
  feat[i] = make_uint4(b.start[i], b.end[i], meta, 0u);
}

// running maximum of the interval ends inside each chromosome (one block per chromosome)
__global__ void k_prefix_max_end(BuildView b, uint4 *feat) {
  __shared__ u32 warpMax[32];
  __shared__ u32 carry;
  const u32 c = blockIdx.x;
  const u32 lo = b.chrStart[c], hi = b.chrStart[c + 1];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (u32 base = lo; base < hi; base += blockDim.x) {
    const u32 i = base + threadIdx.x;
    u32 v = (i < hi) ? feat[i].y : 0u;
    for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xffffffffu, v, d); if (lane >= (u32)d) v = max(v, o); }
    if (lane == 31) warpMax[warp] = v;
    __syncthreads();
    if (warp == 0) {
      u32 w = (lane < (blockDim.x >> 5)) ? warpMax[lane] : 0u;
      for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xffffffffu, w, d); if (lane >= (u32)d) w = max(w, o); }
      warpMax[lane] = w;
    }
    __syncthreads();
    u32 pre = carry;
    if (warp > 0) pre = max(pre, warpMax[warp - 1]);
    v = max(v, pre);
    if (i < hi) feat[i].w = v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = v;
    __syncthreads();
  }
}

__device__ __forceinline__ u32 chrOfEntry(const u32 *chrBinBase, u32 nChr, u32 e) {
  u32 lo = 0, hi = nChr;  // last c with chrBinBase[c] <= e
  while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (chrBinBase[mid] <= e) lo = mid; else hi = mid; }
  return lo;
}

// per bin entry: first feature starting at or after the bin start, and the number of earlier features
// that reach into the bin (found from the running max end).  fill = 0 counts, fill = 1 writes spanIdx.
__global__ void k_build_bins(BuildView b, const uint4 *__restrict__ feat, uint2 *bins, u32 *spanCount, u32 *spanIdx, int fill) {
  const u32 e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= b.nEntries) return;
  const u32 c = chrOfEntry(b.chrBinBase, b.nChr, e);
  const u32 bin = e - b.chrBinBase[c];
  const u32 nBins = b.chrBinBase[c + 1] - b.chrBinBase[c] - 1;
  const u32 cs = b.chrStart[c], ce = b.chrStart[c + 1];
  if (bin >= nBins) {  // sentinel entry
    if (!fill) { bins[e].x = ce; spanCount[e] = 0; }
    return;
  }
  const u64 binStart = (u64)bin << b.shift;
  u32 lo = cs, hi = ce;  // first feature with start >= binStart
  while (lo < hi) { const u32 mid = (lo + hi) >> 1; if ((u64)feat[mid].x < binStart) lo = mid + 1; else hi = mid; }
  const u32 firstIn = lo;
  lo = cs; hi = firstIn;  // first feature whose running max end reaches the bin
  while (lo < hi) { const u32 mid = (lo + hi) >> 1; if ((u64)feat[mid].w < binStart) lo = mid + 1; else hi = mid; }
  if (!fill) {
    u32 cnt = 0;
    for (u32 i = lo; i < firstIn; ++i) cnt += ((u64)feat[i].y >= binStart) ? 1u : 0u;
    bins[e].x = firstIn;
    spanCount[e] = cnt;
  } else {
    u32 at = bins[e].y;
    for (u32 i = lo; i < firstIn; ++i)
      if ((u64)feat[i].y >= binStart) spanIdx[at++] = i;
  }
}

// offsets of the spanning lists (exclusive prefix sum of spanCount, formed by the library scan the index build uses anyway) into bins[].y
__global__ void k_set_span_offsets(const u32 *__restrict__ spanOff, const u32 *__restrict__ spanCount, uint2 *bins, u32 n, u32 *total) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bins[i].y = spanOff[i];
  if (i == n - 1) *total = spanOff[i] + spanCount[i];
}

// ---- segment answer table

// boundary keys (chromosome << 32 | position): position 0 of every chromosome, every feature start, every end + 1
__global__ void k_seg_keys(BuildView b, u64 *keys) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < b.nFeat) {
    const u64 c = (u64)b.chr[i] << 32;
    keys[2 * i] = c | b.start[i];
    keys[2 * i + 1] = c | ((u64)b.end[i] + 1ull);
  }
  if (i < b.nChr) keys[2ull * b.nFeat + i] = (u64)i << 32;
}

// one answer word from the evaluation of a representative read (see FastView); *tie = U + D for an ANS_VICPAIR answer
__device__ __forceinline__ u32 answerWord(u64 chosen, const EvalTrack &tr, u32 upMask, u32 downMask, u32 *tie) {
  const u32 all = (u32)tr.lineAll, U = all & upMask, D = all & downMask;
  *tie = 0;
  // the pick between several matched elements of one line depends on distances to the read (mm:1066-1075) only when one
  // of them is an upstream / downstream element
  if (__popc(all) <= 1 || (U | D) == 0) return (u32)chosen;
  if ((all & ~(U | D)) == 0 && __popc(U) == 1 && __popc(D) == 1) {
    const u64 sum = (u64)tr.pUp + (u64)tr.pDown;
    if (sum <= 0xFFFFFFFFull) { *tie = (u32)sum; return ANS_VICPAIR | all; }
  }
  return ANS_GENERAL;
}

// one thread per segment: the answers of a read inside it and of a read over it and its right neighbour, per
// strand, evaluated with the SAME candidate walk as any hit (inclusion scoring; see fastAnnotate for why that
// also serves the overlap modes)
__global__ void k_seg_eval(IndexView ix, const u64 *__restrict__ segKey, u32 nSeg, u32 upMask, u32 downMask, uint4 *seg, u32 *tieOut) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nSeg) return;
  const u64 k = segKey[i];
  const u32 chr = (u32)(k >> 32), start = (u32)k;
  const bool hasNext = (i + 1 < nSeg) && (u32)(segKey[i + 1] >> 32) == chr;
  const bool hasNext2 = hasNext && (i + 2 < nSeg) && (u32)(segKey[i + 2] >> 32) == chr;
  const bool hasNext3 = hasNext2 && (i + 3 < nSeg) && (u32)(segKey[i + 3] >> 32) == chr;
  const u32 end = hasNext ? (u32)segKey[i + 1] - 1u : 0xFFFFFFFEu;
  const u32 end1 = hasNext ? (hasNext2 ? (u32)segKey[i + 2] - 1u : 0xFFFFFFFEu) : 0u;   // end of segment i+1
  const u32 end2 = hasNext2 ? (hasNext3 ? (u32)segKey[i + 3] - 1u : 0xFFFFFFFEu) : 0u;  // end of segment i+2
  const u32 len1 = hasNext ? min(end1 - end, 65535u) : 0u, len2 = hasNext2 ? min(end2 - end1, 65535u) : 0u;
  u32 in[2], cross[2], triple[2], tie[2];  // index 0: strand bit set ("F"), 1: not set
  for (u32 s = 0; s < 2; ++s) {
    const u32 meta = chr | (s == 0 ? 0x80000000u : 0u);
    EvalTrack tr;
    u32 otherTie;
    u64 a = annotateEval<0, true>(ix, start, start, meta, -1.0f, &tr);
    in[s] = answerWord(a, tr, upMask, downMask, &tie[s]);
    cross[s] = triple[s] = ANS_GENERAL;
    if (hasNext) {
      a = annotateEval<0, true>(ix, end, end + 1u, meta, -1.0f, &tr);
      cross[s] = answerWord(a, tr, upMask, downMask, &otherTie);
      if (cross[s] & ANS_VICPAIR) cross[s] = ANS_GENERAL;  // only in-segment answers have a tie point: left to the index walk
    }
    if (hasNext2) {
      a = annotateEval<0, true>(ix, end, end1 + 1u, meta, -1.0f, &tr);
      triple[s] = answerWord(a, tr, upMask, downMask, &otherTie);
      if (triple[s] & ANS_VICPAIR) triple[s] = ANS_GENERAL;
    }
  }
  seg[2 * i] = make_uint4(end, in[0], in[1], len1 | (len2 << 16));
  seg[2 * i + 1] = make_uint4(cross[0], cross[1], triple[0], triple[1]);
  tieOut[2 * i] = tie[0];
  tieOut[2 * i + 1] = tie[1];
}

// one thread per entry of the position map
__global__ void k_fast_bitmap(const u64 *__restrict__ segKey, u32 nSeg, const u32 *__restrict__ chrBinBase, u32 nChr, u32 shift, u32 gshift,
                              u32 nEntries, uint2 *bm) {
  const u32 e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nEntries) return;
  const u32 c = chrOfEntry(chrBinBase, nChr, e);
  const u64 pos = (u64)(e - chrBinBase[c]) << shift;
  const u64 cap = 0xFFFFFFFEull;
  const u64 first = pos > cap ? cap : pos;
  const u64 want = ((u64)c << 32) | first;
  u32 lo = 0, hi = nSeg;  // last segment whose key <= want (every chromosome has a segment starting at 0)
  while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (segKey[mid] <= want) lo = mid; else hi = mid; }
  u32 bits = 0;
  const u64 binEnd = pos + (1ull << shift);  // exclusive
  for (u32 k = lo + 1; k < nSeg; ++k) {
    const u64 sk = segKey[k];
    if ((u32)(sk >> 32) != c) break;
    const u64 sp = sk & 0xFFFFFFFFull;
    if (sp >= binEnd) break;
    bits |= 1u << (u32)((sp - pos) >> gshift);
  }
  bm[e] = make_uint2(bits, lo);
}

// ---- bin entries (see FastView): one thread per bin, two passes around the numbering of the pair dictionary
#define DICT_SLOTS 4096u
struct BinBuild {
  const u64 *segKey;
  const u32 *chrBinBase;
  const uint4 *seg;
  u32 nSeg, nChr, nEntries, nElements;
  u64 *hashKey;  // [DICT_SLOTS] answer pair + 1 (0 = empty)
  u32 *hashId;   // [DICT_SLOTS] dictionary index of the slot's pair
};
__device__ __forceinline__ u32 dictSlot(u64 pair) { return (u32)(mix64(pair) & (DICT_SLOTS - 1)); }
// pass 0: registers the pair; pass 1: its dictionary index (ENT_DICT - 1 when it has none)
__device__ __forceinline__ u32 dictPair(const BinBuild &b, u32 f, u32 r, int pass) {
  const u32 flags = ANS_VICPAIR | ANS_GENERAL;
  if (f & flags) f = ENT_NONE;
  if (r & flags) r = ENT_NONE;
  const u64 pair = (((u64)r << 32) | f) + 1ull;
  if (pair == 1ull) return 0;                       // {0, 0}
  if (pair == 0ull) return 1;                       // {none, none}
  u32 slot = dictSlot(pair);
  for (u32 probe = 0; probe < 64; ++probe) {
    u64 k = b.hashKey[slot];
    if (k == 0 && pass == 0) k = atomicCAS(&b.hashKey[slot], 0ull, pair), k = k ? k : pair;
    if (k == pair) return pass ? b.hashId[slot] : 0u;
    if (k == 0) break;
    slot = (slot + 1) & (DICT_SLOTS - 1);
  }
  return 1;
}
__global__ void k_bin_entries(BinBuild b, int pass, uint4 *ent, u32 *rank) {
  const u32 e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e > b.nEntries) return;
  if (e == b.nEntries) {  // the dummy bin of hits on a chromosome the annotation does not know: no element
    if (pass) { ent[e] = make_uint4(0u, 0u, 65535u, 0u); rank[e] = 0; }
    return;
  }
  const u32 c = chrOfEntry(b.chrBinBase, b.nChr, e);
  const u64 pos = (u64)(e - b.chrBinBase[c]) << 6;
  const u64 cap = 0xFFFFFFFEull;
  const u64 want = ((u64)c << 32) | (pos > cap ? cap : pos);
  u32 lo = 0, hi = b.nSeg;  // last segment whose key <= want (every chromosome has a segment starting at 0)
  while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (b.segKey[mid] <= want) lo = mid; else hi = mid; }
  u64 bits = 0;
  u32 z = lo;
  for (u32 k = lo + 1; k < b.nSeg; ++k) {
    const u64 sk = b.segKey[k];
    if ((u32)(sk >> 32) != c) break;
    const u64 sp = sk & 0xFFFFFFFFull;
    if (sp >= pos + 64ull) break;
    bits |= 1ull << (u32)(sp - pos);
    z = k;
  }
  const uint4 tz = b.seg[2 * z], xz = b.seg[2 * z + 1], ta = b.seg[2 * lo];
  const u64 binEnd = pos + 63ull;
  const u64 beyond = (u64)tz.x > binEnd ? (u64)tz.x - binEnd : 0ull;  // (Z holds the bin's last position: its end is not before it)
  const u32 lenZ = (u32)(beyond < 65535ull ? beyond : 65535ull);
  u32 lenZ1 = (lenZ < 65535u) ? min(tz.w & 0xFFFFu, 255u) : 0u;  // (a saturated 16-bit length vouches for 65534 positions: more than 255)
  const u32 idZ = dictPair(b, tz.y, tz.z, pass), idA = dictPair(b, ta.y, ta.z, pass);
  u32 idZX = 1;
  if (lenZ1) idZX = dictPair(b, xz.x, xz.y, pass);
  if (pass) {
    ent[e] = make_uint4((u32)bits, (u32)(bits >> 32), lenZ | (lenZ1 << 16), idZ | (idZX << 10) | (idA << 20));
    rank[e] = lo;
  }
}
// numbers the registered pairs (single block; indices 0 and 1 are reserved); *nUsed = entries of the dictionary in use
__global__ void k_dict_number(u64 *hashKey, u32 *hashId, uint2 *dict, u32 *nUsed) {
  __shared__ u32 next;
  if (threadIdx.x == 0) next = 2;
  for (u32 i = threadIdx.x; i < ENT_DICT; i += blockDim.x) dict[i] = (i == 0) ? make_uint2(0u, 0u) : make_uint2(ENT_NONE, ENT_NONE);
  __syncthreads();
  for (u32 sIdx = threadIdx.x; sIdx < DICT_SLOTS; sIdx += blockDim.x) {
    const u64 k = hashKey[sIdx];
    if (k == 0) continue;
    const u32 id = atomicAdd(&next, 1u);
    if (id < ENT_DICT) {
      const u64 pair = k - 1ull;
      hashId[sIdx] = id;
      dict[id] = make_uint2((u32)pair, (u32)(pair >> 32));
    } else {
      hashId[sIdx] = 1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *nUsed = min(next, ENT_DICT);
}

// adjacent-duplicate removal of the sorted boundary keys: flags, then a scatter through their prefix sums
__global__ void k_seg_flag(const u64 *__restrict__ sorted, u32 n, u32 *flag) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || sorted[i] != sorted[i - 1]) ? 1u : 0u;
}
__global__ void k_seg_scatter(const u64 *__restrict__ sorted, const u32 *__restrict__ flag, const u32 *__restrict__ pos, u32 n, u64 *out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) out[pos[i]] = sorted[i];
}

}  // namespace mma
