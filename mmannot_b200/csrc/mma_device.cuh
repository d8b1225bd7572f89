// Device-side data model and kernels of the read-annotation hot path (sm_100a).
//
// K1  index build      : packed features, running max-end, position bins (first-start + spanning lists)
// K2  k_annotate       : per hit -> element set (strand / -l test / Order priority), hit statistics,
//                        immediate counting of the reads that need no grouping
// K3  k_resolve        : per read -> NH countdown over the run of records sharing a read key
// K4  table kernels    : block-privatised shared-memory tables flushed into a device hash table,
//                        batch delta merge, deferred (name-sorted) resolution at end of sample
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

struct HitView {
  const u32 *start, *end, *meta, *nh;
  const u64 *key;
  u32 n;
};

// open-addressing table: combination key -> count.  key 0 = empty (an empty element set is never counted)
struct TableView {
  u64 *keys;
  u64 *vals;
  u32 capMask;
  u32 *overflow;
};

enum StatSlot { ST_HITS = 0, ST_READS, ST_UNIQUE, ST_AMBIGUOUS, ST_MULTIPLE, ST_UNASSIGNED, ST_RESCUED, ST_N = 8 };

// control block of one sample (device memory)
struct SampleCtl {
  u64 stats[ST_N];       // committed counters
  u64 dstats[ST_N];      // counters of the batch in flight (reads / rescued of multi-mapping reads)
  u64 ordBase;           // ordinal of the first hit of the batch in flight
  u32 slowCount;         // deferred records
  u32 slowCountAtBatch;  // value when the batch in flight started
  u32 openCount;         // entries in the open-key set
  u32 dirty;             // the batch in flight found an unfinished read
  u32 overflow;          // any capacity problem (tables, deferred list, key set)
  u32 pad;
};

struct SlowView {  // deferred records: resolved after a (key, ordinal) sort at end of sample
  u64 *key, *ord, *mask;
  u32 *nh;
  u32 cap;
};

struct KeySetView {  // keys whose records must all take the deferred path
  u64 *keys;
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
  __device__ void add(u64 ckey, u32 n, const TableView &g) {
    u32 slot = (u32)mix64(ckey) & (SLOTS - 1);
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
  u32 peers = __match_any_sync(0xffffffffu, ckey);
  if (ckey != 0 && (u32)(__ffs(peers) - 1) == (threadIdx.x & 31u)) bt.add(ckey, __popc(peers), g);
}

__device__ __forceinline__ bool keySetContains(const KeySetView &s, u64 key) {
  u32 slot = (u32)mix64(key) & s.capMask;
  for (u32 probe = 0; probe <= s.capMask; ++probe) {
    u64 k = s.keys[slot];
    if (k == key) return true;
    if (k == KEY_EMPTY) return false;
    slot = (slot + 1) & s.capMask;
  }
  return false;
}
__device__ __forceinline__ void keySetInsert(const KeySetView &s, u64 key, SampleCtl *ctl) {
  u32 slot = (u32)mix64(key) & s.capMask;
  for (u32 probe = 0; probe <= s.capMask; ++probe) {
    u64 k = s.keys[slot];
    if (k == key) return;
    if (k == KEY_EMPTY) {
      k = atomicCAS(&s.keys[slot], KEY_EMPTY, key);
      if (k == KEY_EMPTY) { atomicAdd(&ctl->openCount, 1u); return; }
      if (k == key) return;
    }
    slot = (slot + 1) & s.capMask;
  }
  atomicExch(&ctl->overflow, 1u);
}

__device__ __forceinline__ void slowAppend(const SlowView &s, SampleCtl *ctl, u64 key, u64 ord, u64 mask, u32 nh) {
  u32 at = atomicAdd(&ctl->slowCount, 1u);
  if (at >= s.cap) { atomicExch(&ctl->overflow, 1u); return; }
  s.key[at] = key; s.ord[at] = ord; s.mask[at] = mask; s.nh[at] = nh;
}

// ----------------------------------------------------------------------------- K2: per-hit annotation

struct HitEval {  // running best of the winning Order line (EvaluationStructure::getFirst, mm:1029-1076)
  u64 seen;       // element types already decided (we walk the candidates backwards: the first
                  // passing interval met for a type is the LAST one in feature order, mm:1023-1028)
  u64 chosen;
  u32 bestLine, bestOv, bestDist;
};

template <int MODE>
__device__ __forceinline__ void evalCandidate(const uint4 f, u32 rs, u32 re, u32 rstrand, float ovl, HitEval &ev) {
  if (f.x > re) return;  // the reference stops its walk at the first interval starting after the read (mm:1311)
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
  if (line < ev.bestLine) { ev.bestLine = line; ev.bestOv = sc; ev.bestDist = d; ev.chosen = bit; }
  else if (line == ev.bestLine) {
    if (sc > ev.bestOv) { ev.bestOv = sc; ev.bestDist = d; ev.chosen = bit; }
    else if (sc == ev.bestOv) {
      if (d < ev.bestDist) { ev.bestDist = d; ev.chosen = bit; }
      else if (d == ev.bestDist) ev.chosen |= bit;
    }
  }
}

// IntervalList::scan, mm:1291-1332, as an index lookup: the candidates are the features that reach
// into the bin of the read start (spanning list) plus those that start in the bins the read covers.
template <int MODE>
__device__ __forceinline__ u64 annotateHit(const IndexView &ix, u32 rs, u32 re, u32 meta, float ovl) {
  const u32 chr = meta & 0x00FFFFFFu;
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
  ev.seen = 0; ev.chosen = 0; ev.bestLine = 0xFFFFFFFFu; ev.bestOv = 0; ev.bestDist = 0;
  for (u32 i = hi; i-- > e0.x;) evalCandidate<MODE>(__ldg(&ix.feat[i]), rs, re, rstrand, ovl, ev);
  for (u32 k = e1.y; k-- > e0.y;) evalCandidate<MODE>(__ldg(&ix.feat[__ldg(&ix.spanIdx[k])]), rs, re, rstrand, ovl, ev);
  return ev.chosen;
}

#define ANNOTATE_THREADS 256
#define BT_SLOTS 256

// One thread per hit.  Writes the element set of every hit (default strategy only, for K3),
// accumulates the per-hit counters of Counter::addCount (mm:1666-1668) and counts at once the
// reads that are their own group (mm:1703-1738): NH <= 1 under `default`, every visited hit under
// `unique` / `ratio`.  Under `random` the annotated hits are deferred (order-dependent draw).
template <int MODE, typename MaskT>
__global__ void __launch_bounds__(ANNOTATE_THREADS)
k_annotate(IndexView ix, HitView h, Rules r, MaskT *__restrict__ outMask, TableView table, SampleCtl *ctl, SlowView slow) {
  __shared__ BlockTable<BT_SLOTS> bt;
  __shared__ u32 sstat[ST_N];
  bt.init();
  if (threadIdx.x < ST_N) sstat[threadIdx.x] = 0;
  __syncthreads();
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < h.n;
  u32 rs = 0, re = 0, meta = 0x00FFFFFFu, nh = 1;
  if (valid) { rs = __ldcs(&h.start[i]); re = __ldcs(&h.end[i]); meta = __ldcs(&h.meta[i]); nh = __ldcs(&h.nh[i]); }
  const bool visited = valid && !(r.strategy == 1 && nh != 1);  // unique: only NH == 1 is looked at (mm:1773)
  u64 mask = 0;
  if (visited) mask = annotateHit<MODE>(ix, rs, re, meta, r.overlap);
  const int nreg = __popcll(mask);
  const bool multi = visited && r.strategy == 0 && nh > 1;  // joins the by-name countdown (mm:1669)
  if (r.strategy == 0 && valid) outMask[i] = (MaskT)mask;
  // per-hit counters, one shared atomic per warp and counter
  const u32 lane = threadIdx.x & 31u;
  const u32 bVisited = __ballot_sync(0xffffffffu, visited);
  const u32 bUnassigned = __ballot_sync(0xffffffffu, visited && nreg == 0);
  const u32 bAmbiguous = __ballot_sync(0xffffffffu, visited && nreg > 1);
  const u32 bUnique = __ballot_sync(0xffffffffu, visited && nreg == 1 && nh == 1);
  const u32 bMulti = __ballot_sync(0xffffffffu, multi);
  if (lane == 0) {
    if (bVisited) atomicAdd(&sstat[ST_HITS], __popc(bVisited));
    if (bUnassigned) atomicAdd(&sstat[ST_UNASSIGNED], __popc(bUnassigned));
    if (bAmbiguous) atomicAdd(&sstat[ST_AMBIGUOUS], __popc(bAmbiguous));
    if (bUnique) atomicAdd(&sstat[ST_UNIQUE], __popc(bUnique));
    if (bMulti) atomicAdd(&sstat[ST_MULTIPLE], __popc(bMulti));
    const u32 own = bVisited & ~bMulti;  // each of these is a read of its own (mm:1737)
    if (own) atomicAdd(&sstat[ST_READS], __popc(own));
  }
  u64 ckey = 0;
  if (visited && !multi && mask != 0) {
    if (r.strategy == 2) {
      slowAppend(slow, ctl, normKey(__ldcs(&h.key[i])), ctl->ordBase + i, mask, nh);
    } else {
      u64 m = mask;
      if (r.rescue && nreg > 1) {  // rescue() on a single ambiguous hit (mm:1728 -> 491 -> 497-509): every multiplicity is 1
        const u32 t = (u32)ceilf(__fmul_rn((float)nreg, r.rescueThreshold));
        if (t <= 1u) m = m & (0 - m);  // lowest element reaches the threshold first
      }
      ckey = m;
      if (r.strategy == 3) {
        if (nh >= (1u << (64 - NH_SHIFT))) atomicExch(&ctl->overflow, 1u);
        ckey |= (u64)nh << NH_SHIFT;
      }
    }
  }
  warpCount(bt, ckey, table);
  __syncthreads();
  bt.flush(table);
  if (threadIdx.x < ST_N && sstat[threadIdx.x]) atomicAdd(&ctl->stats[threadIdx.x], (u64)sstat[threadIdx.x]);
}

// ----------------------------------------------------------------------------- K3: per-read resolution

// rescue(), mm:497-509, for a group of records [first, end) of one read (only reachable with -m and -e < 100)
template <typename MaskT>
__device__ u64 rescueGroup(const Rules &r, const MaskT *mask, const u32 *nh, const u64 *key, u64 k, u32 first, u32 end, u64 gm) {
  u32 n = 0;
  for (u32 j = first; j < end; ++j)
    if (normKey(key[j]) == k && nh[j] > 1) n += __popcll((u64)mask[j]);
  if (n == 1) return gm;
  const u32 t = (u32)ceilf(__fmul_rn((float)n, r.rescueThreshold));
  for (u64 rest = gm; rest; rest &= rest - 1) {
    const u64 bit = rest & (0 - rest);
    u32 c = 0;
    for (u32 j = first; j < end; ++j)
      if (normKey(key[j]) == k && nh[j] > 1 && ((u64)mask[j] & bit)) ++c;
    if (c >= t) return bit;
  }
  return gm;
}

#define RESOLVE_THREADS 256

// One thread per hit; the thread of the first record of a run of equal read keys walks the run and
// applies the NH countdown of Counter::addCount (mm:1669-1702): a read opens at a record with NH > 1,
// takes the following NH-1 records of its name (NH <= 1 records are reads of their own and do not
// count down), then its element set is counted.  A read still open at the end of its run cannot be
// finished here (its remaining records may come later in the file, or never): its records and key go
// to the deferred path, and the batch is marked dirty so that pass 1 re-routes every run of such keys.
//   pass 0: normal.   pass 1: only runs if pass 0 left the batch dirty (the delta was discarded).
template <typename MaskT>
__global__ void __launch_bounds__(RESOLVE_THREADS)
k_resolve(HitView h, Rules r, const MaskT *__restrict__ mask, TableView delta, SampleCtl *ctl, SlowView slow, KeySetView open, int pass) {
  if (pass == 1 && ctl->dirty == 0) return;
  __shared__ BlockTable<BT_SLOTS> bt;
  __shared__ u32 sReads, sRescued;
  bt.init();
  if (threadIdx.x == 0) { sReads = 0; sRescued = 0; }
  __syncthreads();
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  u64 commit[1];
  commit[0] = 0;
  u32 nReads = 0, nRescued = 0;
  if (i < h.n) {
    const u64 k = normKey(h.key[i]);
    const bool head = (i == 0) || (normKey(h.key[i - 1]) != k);
    if (head) {
      // find the run and whether it holds any multi-mapping record at all
      const bool anyOpenKeys = (ctl->openCount != 0) || pass == 1;
      bool routed = false;
      if (anyOpenKeys && keySetContains(open, k)) routed = true;
      bool isOpen = false;
      u32 remaining = 0, first = 0;
      u64 gm = 0;
      u32 j = i;
      for (; j < h.n && normKey(h.key[j]) == k; ++j) {
        const u32 nhj = h.nh[j];
        if (!(nhj > 1)) continue;
        const u64 mj = (u64)mask[j];
        if (routed) { slowAppend(slow, ctl, k, ctl->ordBase + j, mj, nhj); continue; }
        if (!isOpen) { isOpen = true; remaining = nhj - 1; gm = mj; first = j; ++nReads; }
        else { --remaining; gm |= mj; }
        if (isOpen && remaining == 0) {
          if (gm != 0) {
            if (r.rescue) gm = rescueGroup(r, mask, h.nh, h.key, k, first, j + 1, gm);
            if (commit[0] == 0) commit[0] = gm;
            else bt.add(gm, 1, delta);  // more than one read closed inside one run: rare
            if (__popcll(gm) == 1) ++nRescued;
          }
          isOpen = false;
        }
      }
      if (isOpen) {  // unfinished read
        --nReads;    // it will be opened again on the deferred path
        for (u32 q = first; q < j; ++q) {
          const u32 nhq = h.nh[q];
          if (nhq > 1) slowAppend(slow, ctl, k, ctl->ordBase + q, (u64)mask[q], nhq);
        }
        keySetInsert(open, k, ctl);
        ctl->dirty = 1;
      }
    }
  }
  warpCount(bt, commit[0], delta);
  if (nReads) atomicAdd(&sReads, nReads);
  if (nRescued) atomicAdd(&sRescued, nRescued);
  __syncthreads();
  bt.flush(delta);
  if (threadIdx.x == 0) {
    if (sReads) atomicAdd(&ctl->dstats[ST_READS], (u64)sReads);
    if (sRescued) atomicAdd(&ctl->dstats[ST_RESCUED], (u64)sRescued);
  }
}

// between pass 0 and pass 1: a dirty batch throws its delta away and rolls the deferred list back
__global__ void k_batch_mid(TableView delta, SampleCtl *ctl) {
  if (ctl->dirty == 0) return;
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= delta.capMask) { delta.keys[i] = 0; delta.vals[i] = 0; }
  if (i == 0) {
    ctl->dstats[ST_READS] = 0; ctl->dstats[ST_RESCUED] = 0;
    ctl->slowCount = ctl->slowCountAtBatch;
  }
}

// end of batch: delta -> sample table, batch counters -> sample counters
__global__ void k_batch_merge(TableView delta, TableView table, SampleCtl *ctl, u32 nHits) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= delta.capMask) {
    const u64 k = delta.keys[i];
    if (k != 0) {
      tableAdd(table, k, delta.vals[i]);
      delta.keys[i] = 0; delta.vals[i] = 0;
    }
  }
  if (i == 0) {
    ctl->stats[ST_READS] += ctl->dstats[ST_READS];
    ctl->stats[ST_RESCUED] += ctl->dstats[ST_RESCUED];
    ctl->dstats[ST_READS] = 0; ctl->dstats[ST_RESCUED] = 0;
    ctl->dirty = 0;
    ctl->slowCountAtBatch = ctl->slowCount;
    ctl->ordBase += nHits;
  }
}

// batches that need no K3 still have to advance the ordinal base
__global__ void k_batch_advance(SampleCtl *ctl, u32 nHits) {
  ctl->slowCountAtBatch = ctl->slowCount;
  ctl->ordBase += nHits;
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
          if (cnt != 1) {
            const u32 t = (u32)ceilf(__fmul_rn((float)cnt, r.rescueThreshold));
            for (u64 rest = gm; rest; rest &= rest - 1) {
              const u64 bit = rest & (0 - rest);
              u32 c = 0;
              for (u32 z = first; z <= q; ++z) if (s.mask[perm[z]] & bit) ++c;
              if (c >= t) { gm = bit; break; }
            }
          }
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
      if (cnt != 1) {
        const u32 t = (u32)ceilf(__fmul_rn((float)cnt, r.rescueThreshold));
        for (u64 rest = gm; rest; rest &= rest - 1) {
          const u64 bit = rest & (0 - rest);
          u32 c = 0;
          for (u32 z = first; z < q; ++z) if (s.mask[perm[z]] & bit) ++c;
          if (c >= t) { gm = bit; break; }
        }
      }
    }
    tableAdd(table, gm, 1);
    if (__popcll(gm) == 1) ++nRescued;
  }
  if (nReads) atomicAdd(&ctl->stats[ST_READS], (u64)nReads);
  if (nRescued) atomicAdd(&ctl->stats[ST_RESCUED], (u64)nRescued);
}

// random (mm:1706-1726): segment heads publish the ordinal of the name's first annotated hit ...
__global__ void k_slow_random_heads(const u32 *__restrict__ perm, u32 n, SlowView s, u64 *headOrd, u32 *nHeads) {
  const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const u64 k = s.key[perm[p]];
  if (p > 0 && s.key[perm[p - 1]] == k) return;
  headOrd[atomicAdd(nHeads, 1u)] = s.ord[perm[p]];
}
// ... and once those ordinals are sorted, the rank of a name's first annotated hit is the index of its
// rand() draw; the drawn-th annotated hit of the name (if the name has that many) is the one counted.
__global__ void k_slow_random_pick(const u32 *__restrict__ perm, u32 n, SlowView s, const u64 *__restrict__ sortedHeadOrd, u32 nHeads,
                                   const u32 *__restrict__ randStream, Rules r, TableView table) {
  const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const u32 id0 = perm[p];
  const u64 k = s.key[id0];
  if (p > 0 && s.key[perm[p - 1]] == k) return;
  const u64 ord0 = s.ord[id0];
  u32 lo = 0, hi = nHeads;
  while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (sortedHeadOrd[mid] < ord0) lo = mid + 1; else hi = mid; }
  const u32 nh0 = s.nh[id0];
  const u32 pick = nh0 ? randStream[lo] % nh0 : 0u;
  const u32 q = p + pick;
  if (q < n && q >= p) {
    const u32 id = perm[q];
    if (s.key[id] == k) {
      u64 m = s.mask[id];
      const int nreg = __popcll(m);
      if (r.rescue && nreg > 1) {
        const u32 t = (u32)ceilf(__fmul_rn((float)nreg, r.rescueThreshold));
        if (t <= 1u) m = m & (0 - m);
      }
      tableAdd(table, m, 1);
    }
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

__device__ __forceinline__ u32 chrOfEntry(const BuildView &b, u32 e) {
  u32 lo = 0, hi = b.nChr;  // last c with chrBinBase[c] <= e
  while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (b.chrBinBase[mid] <= e) lo = mid; else hi = mid; }
  return lo;
}

// per bin entry: first feature starting at or after the bin start, and the number of earlier features
// that reach into the bin (found from the running max end).  fill = 0 counts, fill = 1 writes spanIdx.
__global__ void k_build_bins(BuildView b, const uint4 *__restrict__ feat, uint2 *bins, u32 *spanCount, u32 *spanIdx, int fill) {
  const u32 e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= b.nEntries) return;
  const u32 c = chrOfEntry(b, e);
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

// exclusive prefix sum of spanCount into bins[].y (single block, tiles with a running carry)
__global__ void k_scan_spans(const u32 *__restrict__ spanCount, uint2 *bins, u32 n, u32 *total) {
  __shared__ u32 warpSum[32];
  __shared__ u32 carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (u32 base = 0; base < n; base += blockDim.x) {
    const u32 i = base + threadIdx.x;
    const u32 x = (i < n) ? spanCount[i] : 0u;
    u32 v = x;
    for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xffffffffu, v, d); if (lane >= (u32)d) v += o; }
    if (lane == 31) warpSum[warp] = v;
    __syncthreads();
    if (warp == 0) {
      u32 w = (lane < (blockDim.x >> 5)) ? warpSum[lane] : 0u;
      for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xffffffffu, w, d); if (lane >= (u32)d) w += o; }
      warpSum[lane] = w;
    }
    __syncthreads();
    const u32 incl = v + carry + (warp > 0 ? warpSum[warp - 1] : 0u);
    if (i < n) bins[i].y = incl - x;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

}  // namespace mma
