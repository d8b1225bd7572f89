// k_batch_lean: the batch kernel (K2 + K3 + K4) when the annotation has BIN ENTRIES (FastView::ent, annotations up to ~160 Mb):
// 32-bit element sets (E <= 30), -y default / unique / ratio.  Same decomposition as k_batch / k_batch_fast -- a warp owns a
// contiguous chunk of 128-hit warp tiles, 4 consecutive hits per lane, no block-wide barrier in the loop -- rebuilt around
// three changes:
//
//   stream    the five hit arrays of a tile are brought into a per-warp shared-memory ring by the TMA unit: one elected lane
//             issues five 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx, L2 evict-first) for the tile TWO tiles
//             ahead, the warp waits on the stage's mbarrier and reads its hits with conflict-free 128-bit shared loads, each
//             array only when it is needed (registers).  No global load of the stream goes through the LSU, and the first key
//             of the next tile (needed to close a run that ends on the tile border) is simply read from the next stage.
//   lookup    ONE 16-byte gather per hit: the bin entry of the read start names, in a dictionary of answer pairs kept in shared
//             memory, the answers of the segment that covers the end of the bin (in-segment and over the next boundary) and of
//             the one that covers its start.  Reads the entry cannot answer (5-10 %) are compacted over the warp and take the
//             segment record, one hit per lane.
//   per read  the segmented OR scan over the runs of a tile tests "distance to the nearest run start <= d" per round instead of
//             re-deriving it from the ballot; slots past the end of the batch are padding (no special cases in the loop);
//             rescued reads are counted from the histogram columns at the end instead of per record.
//
// Everything that is not the regular shape (-m rescue, unfinished read names, runs whose NH disagrees with their length, runs cut
// by a chunk border, the read carried into the batch) takes the serial RunWalker of mma_device.cuh exactly as in k_batch -- but
// not here: the lane that finds such a run sets the bit of its first record in a bitmap, and k_batch_walk (below, launched behind
// this kernel) walks the marked runs, one thread per run.  The tile loop holds no call, which is what lets it live in its
// register budget, and the walks of a messy batch spread over the whole GPU instead of serialising inside a warp.
#pragma once
#include "mma_batch_fast.cuh"

namespace mma {

#ifndef MMA_LEAN_THREADS
#define MMA_LEAN_THREADS 640
#endif
#ifndef MMA_LEAN_MAXREG
#define MMA_LEAN_MAXREG 96  // 20 warps per SM (registers are handed out per SM quarter: 5 warps x 32 lanes x 96 <= 16384; warps per block in fours)
#endif
#ifndef MMA_LEAN_BLOCKS_PER_SM
#define MMA_LEAN_BLOCKS_PER_SM 1  // one block per SM: the block-wide tables exist once, the shared memory they do not take stays L1
#endif
#define LEAN_THREADS MMA_LEAN_THREADS
#define LEAN_WARPS (LEAN_THREADS / 32)
#define LEAN_STAGES 2
#define LEAN_STAGE_BYTES 3072  // start[128] | end[128] | meta[128] | nh[128] | key[128]
#define LEAN_BT_SLOTS 1024

// ---- mbarrier / bulk copy (TMA 1-D) primitives
__device__ __forceinline__ u32 smemAddr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbarExpectTx(u32 bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(u32 bar, u32 parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ u64 policyEvictFirst() {
  u64 pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulkLoad(u32 dst, const void *src, u32 bytes, u32 bar, u64 pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar), "l"(pol)
               : "memory");
}
__device__ __forceinline__ uint4 lds128(u32 addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ u64 lds64(u32 addr) {
  u64 v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64x2(u32 addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

struct LeanWarp {  // private to one warp
  alignas(128) unsigned char ring[LEAN_STAGES][LEAN_STAGE_BYTES];
  u64 bar[LEAN_STAGES];
  u32 scratch[WT_HITS];          // answers of the compacted hits; then the first records of the runs to walk
  unsigned char slowQ[WT_HITS];
};
#define LEAN_BAR_OFF (LEAN_STAGES * LEAN_STAGE_BYTES)
// Shared memory of a block: this fixed part, then (sized at launch, so that what the block does not need stays L1 cache)
//   uint2 dict[nDict] | uint2 chrInfo[nChr + 1] | unsigned short hist[nElements + 1][LEAN_THREADS]   (histogram only when HIST)
template <bool HIST>
struct LeanSmem {
  LeanWarp w[LEAN_WARPS];
  typename BlockTableOf<HIST, LEAN_BT_SLOTS>::type bt;
  u32 stat[ST_N + 1];  // [ST_N]: reads of their own counted for a single element (see the epilogue)
  alignas(16) unsigned char var[16];
};
template <bool HIST>
inline size_t leanSmemBytes(u32 nDict, u32 nChr, u32 nElements) {
  return offsetof(LeanSmem<HIST>, var) + 8ull * nDict + 8ull * (nChr + 1) + (HIST ? 2ull * (nElements + 1) * LEAN_THREADS : 0ull) + 16;
}

// A hit the bin entry could not answer: the record of its segment (index = rank of the bin + boundaries up to the read start),
// then like fastAnnotate (which this replaces on the hot path: no chromosome check, no stepping).
template <int MODE>
__device__ __forceinline__ u32 recordAnnotate(const FastView &fx, const IndexView &ix, uint2 ci, u32 rs, u32 re, u32 meta, float ovl) {
  if (re < rs || re >= 0xFFFFFFF0u) return fastMissEval<MODE>(ix, rs, re, meta, ovl);
  const u32 at = ci.x + min(rs >> 6, ci.y - 1u);
  const uint4 e = __ldg(&fx.ent[at]);
  const u32 rank = __ldg(&fx.rank[at]);
  const u64 bits = ((u64)e.y << 32) | e.x;
  const u32 i = rank + __popcll(bits & ((2ull << (rs & 63u)) - 1ull));
  uint4 t, x;
  ldRecord(&fx.seg[2u * i], t, x);
  const bool fwd = (int)meta < 0;
  u32 a;
  if (re <= t.x) {
    a = fwd ? t.y : t.z;
    if (a & ANS_VICPAIR) a = vicPick(fx, a, __ldg(&fx.tie[2u * i + (fwd ? 0u : 1u)]), rs, re);
  } else if (MODE == 0) {
    const int which = segmentsAhead(re - t.x, t.w);
    if (which == 0) return fastMissEval<MODE>(ix, rs, re, meta, ovl);
    a = (which == 1) ? (fwd ? x.x : x.y) : (fwd ? x.z : x.w);
  } else {
    return fastMissEval<MODE>(ix, rs, re, meta, ovl);
  }
  if (a & ANS_GENERAL) return fastMissEval<MODE>(ix, rs, re, meta, ovl);
  return a;
}

// RUNS: how the records of multi-mapping reads are resolved (-y default)
//   0  a run of n records carrying NH = n is one read, closed by the scan
//   1  GROUPS: runs of k x NH records (paired-end data) are resolved in parallel as k reads (see k_batch_fast)
//   2  DEFER: input in which the records of a read are not adjacent (coordinate-sorted files).  No countdown here at all: every
//      record with NH > 1 goes to the deferred list with its element set (one reservation per warp tile), and the end-of-sample
//      pass groups them by read key.  Chosen by the host for a whole sample (sticky) when an earlier batch or sample left most of
//      its multi-mapping reads unfinished.
template <int MODE, int STRAT, int RUNS>
__global__ void __maxnreg__(MMA_LEAN_MAXREG)
k_batch_lean(const __grid_constant__ IndexView ix, const __grid_constant__ FastView fx, const __grid_constant__ HitView h, const __grid_constant__ Rules r,
             const __grid_constant__ TableView table, SampleCtl *ctl, const __grid_constant__ SlowView slow, u32 *walkMap) {
  constexpr bool HIST = (STRAT != 3);
  constexpr bool GROUPS = RUNS == 1, DEFER = (RUNS == 2) && STRAT == 0;
  constexpr u32 FULL = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char leanSmemRaw[];
  LeanSmem<HIST> &sm = *reinterpret_cast<LeanSmem<HIST> *>(leanSmemRaw);
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  uint2 *const smDict = reinterpret_cast<uint2 *>(sm.var);
  uint2 *const smChr = smDict + fx.nDict;
  unsigned short *const myHist = reinterpret_cast<unsigned short *>(smChr + fx.nChr + 1) + tid;  // this thread's column: row e at e * LEAN_THREADS
  const u32 nRows = r.nElements + 1;  // the last row only ever receives zeros
  sm.bt.init();
  if (HIST) {
    for (u32 e = 0; e < nRows; ++e) myHist[e * LEAN_THREADS] = 0;
  }
  if (tid <= ST_N) sm.stat[tid] = 0;
  for (u32 c = tid; c <= fx.nChr; c += LEAN_THREADS) smChr[c] = fx.chrInfo[c];  // launched only when nChr <= CHR_SMEM
  for (u32 c = tid; c < fx.nDict; c += LEAN_THREADS) smDict[c] = fx.dict[c];
  const u32 wbase = smemAddr(&sm.w[warp]);  // ring stage s at wbase + s * LEAN_STAGE_BYTES, its barrier at wbase + LEAN_BAR_OFF + 8 s
  const u32 dict0 = smemAddr(smDict);  // (the chromosome table follows the dictionary)
  LeanWarp &mine = sm.w[warp];
  if (lane == 0) {
    mbarInit(wbase + LEAN_BAR_OFF, 1);
    mbarInit(wbase + LEAN_BAR_OFF + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const u32 seq = ctl->batchSeq;
  // every run takes the serial walker when rescue() needs multiplicities or some read name is known as unfinished
  const bool forceWalk = (STRAT == 0) && (r.rescue || __shfl_sync(FULL, ctl->openCount, 0) != 0);
  // per-lane counters, two 16-bit fields per register (a lane sees at most 4 hits per tile and the host keeps a warp's chunk
  // below 2^14 tiles, see launchBatchKernels):
  u32 pAsgUniq = 0;   // assigned hits | hits of NH = 1 with an element (corrected for ambiguous ones) << 16        (mm:1666, 1668)
  u32 pMultAmbi = 0;  // hits joining the by-name countdown | ambiguous hits << 16                                 (mm:1670, 1667)
  u32 pOwnClos = 0;   // reads of their own counted for one element | multi-mapping reads closed by the scan << 16
  u32 pMissResc = 0;  // segment-table misses | (rescue() active only) closed reads resolved to one element << 16
  u32 pWalks = 0, cVis = 0;

  // a run the scan cannot close is left to k_batch_walk (launched behind this kernel): bit i of walkMap = "walk the run from record i"
  auto queueWalk = [&](u32 i0) {
    if (i0 < h.n) atomicOr(&walkMap[i0 >> 5], 1u << (i0 & 31u));  // (a run of padding slots has nothing to walk)
  };

  const u32 nWT = (h.n + WT_HITS - 1) / WT_HITS;
  const u32 nWarps = gridDim.x * LEAN_WARPS;
  const u32 per = (nWT + nWarps - 1) / nWarps;
  const u32 t0 = min(nWT, (blockIdx.x * LEAN_WARPS + warp) * per), t1 = min(nWT, t0 + per);

  // ---- staging: the i-th tile of the chunk goes to stage i & 1.  Full tiles by bulk copies; the (one) partial tile at the end
  //      of the batch by the warp itself, padded with hits that count for nothing (unknown chromosome, NH = 1, the reserved key).
  auto stage = [&](u32 t, u32 s) {
    const size_t base = (size_t)t * WT_HITS;
    if ((t + 1) * WT_HITS <= h.n) {
      if (lane == 0) {
        const u32 dst = wbase + s * LEAN_STAGE_BYTES, bar = wbase + LEAN_BAR_OFF + s * 8;
        const u64 pol = policyEvictFirst();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the warp's reads of the stage come before the unit's writes
        mbarExpectTx(bar, (STRAT == 0) ? LEAN_STAGE_BYTES : 2048u);
        bulkLoad(dst, h.start + base, 512, bar, pol);
        bulkLoad(dst + 512, h.end + base, 512, bar, pol);
        bulkLoad(dst + 1024, h.meta + base, 512, bar, pol);
        bulkLoad(dst + 1536, h.nh + base, 512, bar, pol);
        if (STRAT == 0) bulkLoad(dst + 2048, h.key + base, 1024, bar, pol);
      }
    } else {
      u32 *d32 = reinterpret_cast<u32 *>(&mine.ring[s][0]);
      u64 *d64 = reinterpret_cast<u64 *>(&mine.ring[s][2048]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const u32 o = lane * 4 + j;
        const size_t i = base + o;
        const bool v = i < h.n;
        d32[o] = v ? h.start[i] : 0u;
        d32[128 + o] = v ? h.end[i] : 0u;
        d32[256 + o] = v ? h.meta[i] : 0x00FFFFFFu;
        d32[384 + o] = v ? h.nh[i] : 1u;
        if (STRAT == 0) d64[o] = v ? normKey(h.key[i]) : KEY_EMPTY;
      }
    }
  };

  bool cValid = false, cCont = false;  // cCont: the run open at the end of the tile continues in the next tile
  u32 cStart = 0, cTot = 0, cNh = 0;
  // lane 0: does the chunk's first record start a run?  (The state carried into the batch counts as the record before it.)
  // The read carried into the batch itself is k_batch_walk's.
  bool headFirst = true;
  if (STRAT == 0 && !DEFER && lane == 0 && t0 < t1) {
    const u64 first = normKey(__ldg(&h.key[(size_t)t0 * WT_HITS]));
    if (t0 == 0) {
      const Carry &c = ctl->carry[seq & 1];
      if (c.valid) headFirst = c.key != first;
    } else {
      headFirst = normKey(__ldg(&h.key[(size_t)t0 * WT_HITS - 1])) != first;
    }
  }
  if (t0 < t1) stage(t0, 0);
  if (t0 + 1 < t1) stage(t0 + 1, 1);
  __syncwarp();

  for (u32 t = t0, it = 0; t < t1; ++t, ++it) {
    const u32 s = it & 1u;
    const u32 base = t * WT_HITS + lane * 4;
    const bool fullTile = (t + 1) * WT_HITS <= h.n;
    const u32 st = wbase + s * LEAN_STAGE_BYTES + lane * 16;
    if (fullTile) mbarWait(wbase + LEAN_BAR_OFF + s * 8, (it >> 1) & 1u);
    // ---- run starts
    u32 hbits = 0, F = 0;
    u64 nextKey = KEY_EMPTY;
    if (STRAT == 0 && !DEFER) {
      u64 key[4];
      const uint4 k0 = lds128(st + 2048 + lane * 16), k1 = lds128(st + 2048 + lane * 16 + 16);
      key[0] = ((u64)k0.y << 32) | k0.x; key[1] = ((u64)k0.w << 32) | k0.z; key[2] = ((u64)k1.y << 32) | k1.x; key[3] = ((u64)k1.w << 32) | k1.z;
      // the all-ones key is reserved (normKey): only a tile that holds a key with all-ones upper half needs the fix-up (the
      // partial tile was normalised when it was staged: its padding keeps the reserved key)
      const u32 hiMax = max(max(k0.y, k0.w), max(k1.y, k1.w));
      if (fullTile && __any_sync(FULL, hiMax == 0xFFFFFFFFu)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) key[j] = normKey(key[j]);
      }
      const u64 prev = __shfl_up_sync(FULL, key[3], 1);
      // lane 0: the tile's first record starts a run unless the previous tile's last run continues (first tile: see headFirst)
      const bool head0 = (lane == 0) ? (it == 0 ? headFirst : !cCont) : (key[0] != prev);
      hbits = (head0 ? 1u : 0u) | ((key[1] != key[0]) ? 2u : 0u) | ((key[2] != key[1]) ? 4u : 0u) | ((key[3] != key[2]) ? 8u : 0u);
      F = __ballot_sync(FULL, hbits != 0);
      nextKey = __shfl_sync(FULL, key[3], 31);
    }
    // ---- lookup: the bin entries of the read starts
    u32 m[4];
    u32 slowBits = 0;
    {
      u32 rs[4], meta[4];
      uint4 en[4];
      const uint4 a = lds128(st), c = lds128(st + 1024);
      rs[0] = a.x; rs[1] = a.y; rs[2] = a.z; rs[3] = a.w;
      meta[0] = c.x; meta[1] = c.y; meta[2] = c.z; meta[3] = c.w;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint2 ci = lds64x2(dict0 + 8u * fx.nDict + 8u * min(meta[j] & 0x00FFFFFFu, fx.nChr));
        en[j] = __ldg(&fx.ent[ci.x + min(rs[j] >> 6, ci.y - 1u)]);
      }
      const uint4 b = lds128(st + 512);
      const u32 re[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const u64 bits = ((u64)en[j].y << 32) | en[j].x;
        const bool inZ = ((bits >> (rs[j] & 63u)) >> 1) == 0;  // no boundary after the read start inside its bin
        const u32 binEnd = rs[j] | 63u;
        const u32 x = max(re[j], binEnd) - binEnd;               // how far the read reaches beyond its bin
        const u32 lenZ = en[j].z & 0xFFFFu, lenZ1 = en[j].z >> 16;
        const bool inside = x <= lenZ;
        const bool cross = MODE == 0 && !inside && x - lenZ <= lenZ1;
        const u32 id = inside ? (en[j].w & 1023u) : ((en[j].w >> 10) & 1023u);
        const uint2 pair = lds64x2(dict0 + 8u * id);
        const u32 a = ((int)meta[j] < 0) ? pair.x : pair.y;
        bool ok = inZ && (inside || cross) && a != ENT_NONE && re[j] >= rs[j];
        bool pass = true;
        if (MODE != 0) {  // no feature can overlap the read by more than end - start
          const u32 o = re[j] - rs[j];
          if (MODE == 1) pass = (o != 0) && (__fmul_rn((float)(o + 1u), r.overlap) <= (float)o);
          else pass = (o != 0) && ((float)o >= r.overlap);
          if (re[j] < rs[j]) pass = true;  // (an empty interval, end = start - 1: left to the index walk)
        }
        m[j] = (ok && pass) ? a : 0u;
        if (!ok && pass) slowBits |= 1u << j;
      }
    }
    u32 nh[4];
    {
      const uint4 d = lds128(st + 1536);
      nh[0] = d.x; nh[1] = d.y; nh[2] = d.z; nh[3] = d.w;
    }
    if (STRAT == 1) {  // unique: only NH == 1 is looked at (mm:1773)
      const u32 skip = (nh[0] != 1 ? 1u : 0u) | (nh[1] != 1 ? 2u : 0u) | (nh[2] != 1 ? 4u : 0u) | (nh[3] != 1 ? 8u : 0u);
      slowBits &= ~skip;
      cVis += 4u - __popc(skip);
#pragma unroll
      for (int j = 0; j < 4; ++j) if ((skip >> j) & 1u) m[j] = 0;
    }
    // ---- the rest, compacted over the warp: one hit per lane against the segment record (and, behind it, the feature index)
    if (__any_sync(FULL, slowBits != 0)) {
      u32 off[4], total = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const u32 b = __ballot_sync(FULL, (slowBits >> j) & 1u);
        off[j] = total + __popc(b & ((1u << lane) - 1u));
        total += __popc(b);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if ((slowBits >> j) & 1u) mine.slowQ[off[j]] = (unsigned char)(lane * 4 + j);
      __syncwarp();
      const u32 *st32 = reinterpret_cast<const u32 *>(&mine.ring[s][0]);
      for (u32 k = lane; k < total; k += 32) {
        const u32 id = mine.slowQ[k];
        const u32 qs = st32[id], qe = st32[128 + id], qm = st32[256 + id];
        const u32 chr = qm & 0x00FFFFFFu;
        mine.scratch[id] = (chr < fx.nChr) ? recordAnnotate<MODE>(fx, ix, smChr[chr], qs, qe, qm, r.overlap) : 0u;
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if ((slowBits >> j) & 1u) {
          const u32 a = mine.scratch[lane * 4 + j];
          m[j] = a & ~FAST_MISS;
          pMissResc += a >> 31;
        }
    }
    if (DEFER) {
      // every record with NH > 1 to the deferred list: {key, ordinal, element set, NH}
      const u32 mb = (nh[0] > 1 ? 1u : 0u) | (nh[1] > 1 ? 2u : 0u) | (nh[2] > 1 ? 4u : 0u) | (nh[3] > 1 ? 8u : 0u);
      u32 mine4 = __popc(mb), incl = mine4;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(FULL, incl, d); if (lane >= (u32)d) incl += o; }
      const u32 total = __shfl_sync(FULL, incl, 31);
      if (total) {
        u32 at0 = 0;
        if (lane == 0) at0 = atomicAdd(&ctl->slowCount, total);
        at0 = __shfl_sync(FULL, at0, 0);
        if (at0 + total > slow.cap) {
          if (lane == 0) atomicExch(&ctl->overflow, 1u);
        } else if (mb) {
          u32 at = at0 + incl - mine4;
          const u64 ord0 = ctl->ordBase + base;
          const uint4 k0 = lds128(st + 2048 + lane * 16), k1 = lds128(st + 2048 + lane * 16 + 16);
          const u64 kk[4] = {((u64)k0.y << 32) | k0.x, ((u64)k0.w << 32) | k0.z, ((u64)k1.y << 32) | k1.x, ((u64)k1.w << 32) | k1.z};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if ((mb >> j) & 1u) {
              slow.key[at] = fullTile ? normKey(kk[j]) : kk[j];  // (the partial tile was normalised when it was staged)
              slow.ord[at] = ord0 + j; slow.mask[at] = m[j]; slow.nh[at] = nh[j];
              ++at;
            }
        }
      }
    }
    __syncwarp();
    // the stage is free: bring in the tile two tiles ahead
    if (t + 2 < t1) stage(t + 2, s);
    // ---- per-hit counters (mm:1666-1668); a hit that does not join the by-name countdown is a read of its own.  (Padding
    //      slots -- and, under -y unique, the hits that are not looked at -- hold m = 0; what the padding adds to the hit count is
    //      taken back after the loop.)
    u32 ev[4];  // the element set counted at this hit slot (0 = none)
    u32 ambOr = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool multi = STRAT == 0 && nh[j] > 1;
      if (m[j] != 0) pAsgUniq += (nh[j] == 1) ? 0x10001u : 1u;
      pMultAmbi += multi ? 1u : 0u;
      ambOr |= m[j] & (m[j] - 1u);
      ev[j] = multi ? 0u : m[j];
      if (r.rescue) ev[j] = (u32)rescueSingle(r, (u64)ev[j]);
    }
    if (ambOr) {  // some hit of this lane matched several elements (mm:1667): it is ambiguous, and not "unique"
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (m[j] & (m[j] - 1u)) {
          pMultAmbi += 0x10000u;
          if (nh[j] == 1) pAsgUniq -= 0x10000u;
        }
    }
    u32 nWalk = 0, closeBits = 0;
    u32 inc = 0, lastHeadPos = 0, F2 = 0;
    bool serialTile = false;
    u32 tileEndsRun = 1;  // the record after the tile's last one starts another run (or the batch ends there)
    if (STRAT == 0 && !DEFER) {
      // first key of the next tile: from the next stage of the ring, or (last tile of the chunk) from global memory
      if (t + 1 < t1) {
        const bool nextFull = (t + 2) * WT_HITS <= h.n;
        if (nextFull) mbarWait(wbase + LEAN_BAR_OFF + (s ^ 1u) * 8, ((it + 1) >> 1) & 1u);
        u64 pk = lds64(wbase + (s ^ 1u) * LEAN_STAGE_BYTES + 2048);
        if (nextFull) pk = normKey(pk);
        tileEndsRun = (pk != nextKey) ? 1u : 0u;
      } else if ((size_t)t1 * WT_HITS < h.n) {
        tileEndsRun = (normKey(__ldg(&h.key[(size_t)t1 * WT_HITS])) != nextKey) ? 1u : 0u;
      }
      // ---- per-read countdown (mm:1669-1702)
      const u32 before = F & ((1u << lane) - 1u);
      u32 prevNh = __shfl_up_sync(FULL, nh[3], 1);
      if (lane == 0) prevNh = cNh;
      lastHeadPos = base + (31 - __clz(hbits | 1u));
      const u32 sPrev = __shfl_sync(FULL, lastHeadPos, before ? (31 - __clz(before)) : 0);
      const u32 nextHead0 = __shfl_down_sync(FULL, hbits & 1u, 1);
      // bit j: the next record starts another run, i.e. this record ends its run (the tile's last record: from the peek)
      const u32 lastBits = (hbits >> 1) | ((lane < 31u ? nextHead0 : tileEndsRun) << 3);
      // a run that starts before this lane's hits: its first record, and whether this warp owns it at all
      const u32 inStart = before ? sPrev : cStart;
      const bool inMine = before || cValid;  // else the run starts in another warp's chunk: that warp finishes it
      u32 walkBits = 0;  // slots whose run (GROUPS: whose group) has to be walked serially
      u32 off[4];        // GROUPS: position of the record inside its group
      if (GROUPS) {
        // A read opens at a record with NH = n > 1 and takes the n - 1 records of its name that follow, so a run of records
        // sharing a read key and carrying the same NH = n is a sequence of GROUPS of n records, one read each (one group for
        // single-end data, two -- the two mates -- for paired-end data, mm:1673-1681).  The element set of a read is the
        // union over its group: a segmented OR scan over the 128 hits of the warp tile, segments starting at run starts and
        // at every n-th record of a run, seeded with the state carried from the previous tile; the lane owning the LAST
        // record of a group counts the read.  A run that ends inside a group leaves an unfinished read: serial walk
        // (RunWalker) from the start of that group.  A tile in which NH changes inside a run -- or any tile while rescue() needs
        // multiplicities or some read name is known as unfinished -- is resolved serially: one walk per run (serialTile).
        u32 badBits = ((!(hbits & 1u) && nh[0] != prevNh) ? 1u : 0u) | ((!(hbits & 2u) && nh[1] != nh[0]) ? 2u : 0u) |
                      ((!(hbits & 4u) && nh[2] != nh[1]) ? 4u : 0u) | ((!(hbits & 8u) && nh[3] != nh[2]) ? 8u : 0u);
        // (only in runs this warp owns: the run reaching into the chunk's first tile is the previous chunk's, NH carried or not)
        if (!inMine) badBits &= (hbits & 1u) ? 0xEu : (hbits & 2u) ? 0xCu : (hbits & 4u) ? 0x8u : 0u;
        serialTile = forceWalk || __any_sync(FULL, badBits != 0);
        if (!serialTile) {
          u32 hb2 = hbits, endBits = 0;
          bool longRun = false;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const u32 hbLe = hbits & ((2u << j) - 1u);
            const u32 runStart = hbLe ? base + (31 - __clz(hbLe)) : inStart;
            const bool own = (hbLe || inMine) && nh[j] > 1;
            u32 o = base + j - runStart;
            if (o >= nh[j]) o -= nh[j];
            if (own && o >= nh[j]) longRun = true;  // third group or later: exact remainder below
            off[j] = o;
          }
          if (__any_sync(FULL, longRun)) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (nh[j] > 1) off[j] %= nh[j];
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const u32 hbLe = hbits & ((2u << j) - 1u);
            const bool own = (hbLe || inMine) && nh[j] > 1;
            if (own && off[j] == 0) hb2 |= 1u << j;                                  // first record of a group
            if (own && off[j] + 1 == nh[j]) endBits |= 1u << j;                      // last record of a group
            else if (own && ((lastBits >> j) & 1u)) walkBits |= 1u << j;             // the run ends inside a group
          }
          F2 = __ballot_sync(FULL, hb2 != 0);
          pWalks += __popc(hb2 & ~hbits);  // groups beyond the first of their run: what the other variant would walk serially
          u32 pre[4], acc = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc = ((hb2 >> j) & 1u) ? m[j] : (acc | m[j]);
            pre[j] = acc;
          }
          inc = acc;  // inclusive scan over lanes: OR of `acc` from the nearest lane with a group start up to this lane
          const u32 upTo2 = F2 & (0xFFFFFFFFu >> (31u - lane));
          const u32 hd2 = upTo2 ? (u32)__clz(upTo2) - (31u - lane) : 64u;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const u32 tt = __shfl_up_sync(FULL, inc, d);
            if ((u32)d <= hd2) inc |= tt;
          }
          u32 X = __shfl_up_sync(FULL, inc, 1);
          if (lane == 0) X = 0;
          const u32 inTot = (F2 & ((1u << lane) - 1u)) ? X : (cTot | X);  // union so far of a group that starts before this lane's hits
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if ((endBits >> j) & 1u) {
              ev[j] = (hb2 & ((2u << j) - 1u)) ? pre[j] : (inTot | pre[j]);
              closeBits |= 1u << j;
            }
          if (walkBits) {  // rare
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if ((walkBits >> j) & 1u) { queueWalk(base + j - off[j]); ++nWalk; }
          }
        } else {
          // serial tile: the run carried into the tile (from the start of its open group) and every run that starts in it
          if (lane == 0 && cValid && !(hbits & 1u)) {
            const u32 o = base - cStart;
            queueWalk(base - ((cNh > 1) ? o % cNh : 0u)); ++nWalk;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if ((hbits >> j) & 1u) { queueWalk(base + j); ++nWalk; }
        }
      } else {
        // A run of n records that all carry NH = n (> 1) is one read; its element set is the union over the run: a
        // segmented OR scan over the 128 hits of the warp tile (bit 31 of the scanned word = "irregular": NH changes inside
        // the run, or every run has to be walked), seeded with the state carried from the previous tile.  The lane owning
        // the run's LAST record closes it; irregular runs are marked for the serial walk (k_batch_walk) by the same lane.
        u32 pre[4], acc = 0;
        const u32 force = forceWalk ? 0x80000000u : 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool isHead = (hbits >> j) & 1u;
          const bool bad = !isHead && nh[j] != (j ? nh[j > 0 ? j - 1 : 0] : prevNh);
          const u32 x = m[j] | (bad ? 0x80000000u : 0u) | force;
          acc = isHead ? x : (acc | x);
          pre[j] = acc;
        }
        inc = acc;  // inclusive scan over lanes: OR of `acc` from the nearest lane with a run start up to this lane
        const u32 upTo = F & (0xFFFFFFFFu >> (31u - lane));
        const u32 hd = upTo ? (u32)__clz(upTo) - (31u - lane) : 64u;  // lanes back to the nearest one in which a run starts
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const u32 tt = __shfl_up_sync(FULL, inc, d);
          if ((u32)d <= hd) inc |= tt;
        }
        u32 X = __shfl_up_sync(FULL, inc, 1);
        if (lane == 0) X = 0;
        u32 carryTot = before ? X : (cTot | X);  // union so far of a run that starts before this lane's hits ...
        u32 runLen = base - inStart;             // ... and its length so far
        bool own = inMine;                       // ... and whether this warp owns it
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool isHead = (hbits >> j) & 1u;
          if (isHead) { carryTot = 0; own = true; }
          runLen = isHead ? 1u : runLen + 1u;
          const u32 tot = carryTot | pre[j];  // union of the run's element sets up to this record, wherever the run starts
          // the run's last record, in a run this warp owns: closed here when it is regular, else walked (a run of reads that are
          // their own group, NH <= 1 throughout, has nothing to close).  Branch-free: the lanes disagree on almost every tile.
          const bool last = ((lastBits >> j) & 1u) && own;
          const bool regular = (int)tot >= 0 && nh[j] > 1 && nh[j] == runLen;
          const bool irregular = !regular && ((int)tot < 0 || nh[j] > 1);
          ev[j] = (last && regular) ? tot : ev[j];
          closeBits |= (last && regular) ? (1u << j) : 0u;
          walkBits |= (last && irregular) ? (1u << j) : 0u;
        }
        if (walkBits) {  // rare
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if ((walkBits >> j) & 1u) {
              const u32 hbLe = hbits & ((2u << j) - 1u);
              queueWalk(hbLe ? base + (31 - __clz(hbLe)) : inStart); ++nWalk;
            }
        }
      }
    }
    // ---- counting: single-element sets into the lane's histogram column, the rest into the block table
    {
      u32 pend = 0, singleBits = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const u32 c = ev[j];
        const bool single = c != 0 && (c & (c - 1)) == 0;
        if (HIST) {
          const u32 row = single ? (u32)(31 - __clz(c)) : r.nElements;  // (the last row only ever receives zeros)
          myHist[row * LEAN_THREADS] += single ? 1 : 0;
          if (single) singleBits |= 1u << j;
          else if (c != 0) pend |= 1u << j;
        } else {
          if (c != 0) pend |= 1u << j;
        }
      }
      // rescue() active: closed reads resolved to one element, one by one; else: reads of their own counted for one element
      // (the closed ones then follow from the histogram columns, see the epilogue)
      if (r.rescue) pMissResc += __popc(singleBits & closeBits) << 16;
      else pOwnClos += __popc(singleBits & ~closeBits);
      pOwnClos += __popc(closeBits) << 16;
      while (__any_sync(FULL, pend != 0)) {
        if (pend) {
          const int j = __ffs(pend) - 1;
          pend &= pend - 1;
          const u32 c = (j == 0) ? ev[0] : (j == 1) ? ev[1] : (j == 2) ? ev[2] : ev[3];
          u64 ckey = c;
          if (STRAT == 3) {
            const u32 n = (j == 0) ? nh[0] : (j == 1) ? nh[1] : (j == 2) ? nh[2] : nh[3];
            if (n >= (1u << (64 - NH_SHIFT))) atomicExch(&ctl->overflow, 1u);
            ckey |= (u64)n << NH_SHIFT;
          }
          sm.bt.add(ckey, 1, table);
        }
      }
    }
    if (STRAT == 0 && !DEFER) {
      pWalks += nWalk;
      const u32 incLast = __shfl_sync(FULL, inc, 31);
      if (GROUPS) {
        // the run (and, inside it, the group) still open at the end of the tile
        if (!serialTile) {
          if (F2) cTot = incLast;
          else cTot |= incLast;
          if (F) {
            cStart = __shfl_sync(FULL, lastHeadPos, 31 - __clz(F));
            cValid = true;
          }
        } else {
          cValid = false;  // every run reaching into or starting in the tile has been walked to its end
          cTot = 0;
        }
      } else {
        // the run still open at the end of the tile
        if (F) {
          cTot = incLast;
          cStart = __shfl_sync(FULL, lastHeadPos, 31 - __clz(F));
          cValid = true;
        } else if (cValid) {
          cTot |= incLast;
        }
      }
      cNh = __shfl_sync(FULL, nh[3], 31);
      cCont = tileEndsRun == 0;
    }
  }
  // ---- a run open at the end of the chunk continues in another warp's chunk (which leaves it alone: it does not own its start).
  //      This warp finishes it: the records of the run in the NEXT tile are annotated, 4 per lane, and lane 0 takes the countdown
  //      through them (GROUPS: from the group open at the border through every later group of the run).  No per-hit counter moves:
  //      the hits belong to the other chunk.  A run that is irregular, that reaches beyond that tile or (GROUPS) that ends inside
  //      a group is marked for k_batch_walk from the run's (the open group's) first record.  (A run ending exactly at the chunk's
  //      last record was closed above, like the last run of the batch.)
  if (STRAT == 0 && !DEFER && cValid && cCont && t1 > t0) {
    const u32 next = t1 * WT_HITS;
    const u32 o = next - cStart;                                              // records of the run inside this chunk
    const u32 inGroup = (GROUPS && cNh > 1) ? o % cNh : 0u;                   // GROUPS: ... of which in the group open at the border
    const u32 walkFrom = GROUPS ? next - inGroup : cStart;
    const u64 k = normKey(lds64(wbase + ((t1 - t0 - 1u) & 1u) * LEAN_STAGE_BYTES + 2048 + 127 * 8));  // the chunk's last record
    u32 tn[4], tm[4], eq = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const size_t i = (size_t)next + lane * 4 + j;
      const bool v = i < h.n;
      const u64 kk = v ? normKey(__ldg(&h.key[i])) : ~k;
      const u32 qs = v ? __ldg(&h.start[i]) : 0u, qe = v ? __ldg(&h.end[i]) : 0u, qm = v ? __ldg(&h.meta[i]) : 0x00FFFFFFu;
      tn[j] = v ? __ldg(&h.nh[i]) : 0u;
      tm[j] = 0;
      if (kk == k) {
        eq |= 1u << j;
        const u32 chr = qm & 0x00FFFFFFu;
        bool pass = true;
        if (MODE != 0 && qe >= qs) {  // (as in the loop: no feature can overlap the read by more than end - start)
          const u32 ol = qe - qs;
          if (MODE == 1) pass = (ol != 0) && (__fmul_rn((float)(ol + 1u), r.overlap) <= (float)ol);
          else pass = (ol != 0) && ((float)ol >= r.overlap);
        }
        if (chr < fx.nChr && pass) tm[j] = recordAnnotate<MODE>(fx, ix, smChr[chr], qs, qe, qm, r.overlap) & ~FAST_MISS;
      }
    }
    const u32 lead = (eq == 0xFu) ? 4u : (u32)__ffs(~eq) - 1u;  // this lane's records that continue the run, if every lane before is all run
    const u32 fullLanes = __ballot_sync(FULL, eq == 0xFu);
    const u32 fp = (fullLanes == FULL) ? 32u : (u32)__ffs(~fullLanes) - 1u;  // first lane that is not all run
    const u32 mineN = (lane < fp) ? 4u : (lane == fp) ? lead : 0u;
    const u32 tailLen = __reduce_add_sync(FULL, mineN);
    u32 prevNh = __shfl_up_sync(FULL, tn[3], 1);
    if (lane == 0) prevNh = cNh;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((u32)j < mineN) {
        bad |= tn[j] != (j ? tn[j > 0 ? j - 1 : 0] : prevNh);
        mine.scratch[lane * 4 + j] = tm[j];
      }
    const bool anyBad = __any_sync(FULL, bad) || forceWalk || (int)cTot < 0 || fp == 32u;
    __syncwarp();
    if (lane == 0) {
      if (anyBad) {
        queueWalk(walkFrom);
      } else if (cNh > 1) {  // (a run of reads that are their own group has nothing to close)
        // one read closed by the countdown, counted like the closes of the loop
        auto closeRead = [&](u32 c) {
          pOwnClos += 1u << 16;
          if (c == 0) return;
          if (HIST && (c & (c - 1)) == 0) myHist[(31 - __clz(c)) * LEAN_THREADS] += 1;
          else sm.bt.add((u64)c, 1, table);
        };
        if (!GROUPS) {
          if (o + tailLen == cNh) {
            u32 acc = cTot;
            for (u32 q = 0; q < tailLen; ++q) acc |= mine.scratch[q];
            closeRead(acc);
          } else {
            queueWalk(walkFrom);
          }
        } else if ((o + tailLen) % cNh != 0) {  // the run ends inside a group
          queueWalk(walkFrom);
        } else {
          u32 acc = inGroup ? cTot : 0u, pos = inGroup;
          for (u32 q = 0; q < tailLen; ++q) {
            acc |= mine.scratch[q];
            if (++pos == cNh) { closeRead(acc); acc = 0; pos = 0; }
          }
        }
      }
    }
  }
  const u32 cAsg = pAsgUniq & 0xFFFFu, cClosed = pOwnClos >> 16, cResc = pMissResc >> 16;
  u32 cMulti = pMultAmbi & 0xFFFFu, cUniq = pAsgUniq >> 16, cAmbi = pMultAmbi >> 16, cOwn1 = pOwnClos & 0xFFFFu, cMiss = pMissResc & 0xFFFFu;
  // hits looked at: 4 per lane and tile (-y unique: those with NH = 1), minus the padding of the batch's last tile
  u32 cHits = 0;
  if (t1 > t0) {
    cHits = (STRAT == 1) ? cVis : 4u * (t1 - t0);
    if (t1 == nWT) {
      const u32 b = (nWT - 1) * WT_HITS + lane * 4;
      const u32 real = (h.n > b) ? min(h.n - b, 4u) : 0u;
      cHits -= 4u - real;
    }
  }
  u32 cUnassigned = cHits - cAsg;
  u32 cReads = cHits - cMulti + cClosed, cRescued = cResc;  // (the reads k_batch_walk opens and closes are added there)

  // ---- block epilogue: counters, the private histogram columns and the private table
  cHits = __reduce_add_sync(FULL, cHits); cUnassigned = __reduce_add_sync(FULL, cUnassigned);
  cAmbi = __reduce_add_sync(FULL, cAmbi); cUniq = __reduce_add_sync(FULL, cUniq);
  cMulti = __reduce_add_sync(FULL, cMulti); cReads = __reduce_add_sync(FULL, cReads);
  cRescued = __reduce_add_sync(FULL, cRescued); cMiss = __reduce_add_sync(FULL, cMiss);
  cOwn1 = __reduce_add_sync(FULL, cOwn1);
  if (STRAT == 0 && !forceWalk && !DEFER) {
    pWalks = __reduce_add_sync(FULL, pWalks);
    if (lane == 0 && pWalks) atomicAdd(&ctl->walkCount, pWalks);
  }
  if (lane == 0) {
    atomicAdd(&sm.stat[ST_HITS], cHits); atomicAdd(&sm.stat[ST_UNASSIGNED], cUnassigned); atomicAdd(&sm.stat[ST_AMBIGUOUS], cAmbi);
    atomicAdd(&sm.stat[ST_UNIQUE], cUniq); atomicAdd(&sm.stat[ST_MULTIPLE], cMulti); atomicAdd(&sm.stat[ST_READS], cReads);
    atomicAdd(&sm.stat[ST_RESCUED], cRescued); atomicAdd(&sm.stat[7], cMiss); atomicAdd(&sm.stat[ST_N], cOwn1);
  }
  __syncthreads();
  sm.bt.flush(table);
  if (HIST) {
    u32 singles = 0;
    for (u32 e = warp; e < r.nElements; e += LEAN_WARPS) {
      u32 v = 0;
#pragma unroll
      for (int q = 0; q < LEAN_WARPS; ++q) v += (myHist - tid)[e * LEAN_THREADS + lane + 32 * q];
      v = __reduce_add_sync(FULL, v);
      if (lane == 0 && v) tableAdd(table, 1ull << e, v);
      singles += v;
    }
    // multi-mapping reads resolved to one element (mm:1691) = reads counted for a single element - those that were a read of
    // their own (with rescue() active they are counted one by one instead: a rescued single hit is neither)
    if (STRAT == 0 && !r.rescue && lane == 0 && singles) atomicAdd(&ctl->stats[ST_RESCUED], (u64)singles);
  }
  if (tid < 7) {
    int sv = (int)sm.stat[tid];  // reads / rescued can be negative within a block (unfinished reads)
    if (tid == ST_RESCUED && STRAT == 0 && !r.rescue) sv -= (int)sm.stat[ST_N];
    if (sv) atomicAdd(&ctl->stats[tid], (u64)(long long)sv);
  }
  if (tid == 7 && sm.stat[7]) atomicAdd(&ctl->fastMiss, sm.stat[7]);
}

// RunWalker::walk (mma_device.cuh) for a whole warp: the same sequence of decisions, taken by all 32 lanes on identical state, with
// the records of the run fetched and annotated 32 at a time (lane l holds record base + l) -- the serial walk is a chain of dependent
// loads per record, and the runs left to k_batch_walk (one per chunk border on clean input) would otherwise be a tail of tens of
// microseconds behind every batch.  Side effects (counts, deferred records, the carry, the open-name set) are lane 0's.
template <int MODE>
struct WarpWalker {
  const HitView &h;
  const Rules &r;
  const Annotator<MODE, true> &annot;
  SampleCtl *ctl;
  const SlowView &slow;
  const KeySetView &open;
  const TableView &table;
  u32 seq, lane;
  int nReads, nRescued;  // (identical on all lanes)

  __device__ __forceinline__ u64 maskAt(u32 j) const { return annot(h.start[j], h.end[j], h.meta[j]); }

  __device__ u64 rescueGroup(u32 first, u32 end, u64 gm) const {  // as RunWalker::rescueGroup (every lane computes the same)
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

  // i = first record of the run inside this batch, k = its key; `cin` = state carried into the batch (or null); all lanes alike
  __device__ __forceinline__ void walk(u32 i, u64 k, const Carry *cin) {  // (inlined: the walker then lives in registers)
    constexpr u32 FULL = 0xffffffffu;
    u32 routedL = 0;
    if (lane == 0) routedL = ((ctl->openCount != 0) && keySetLookup(open, k, seq) == 1) ? 1u : 0u;
    const bool routed = __shfl_sync(FULL, routedL, 0) != 0;
    bool isOpen = false, fromCarry = false;
    u32 remaining = 0, first = i;
    u64 gm = 0;
    if (cin) { isOpen = true; fromCarry = true; remaining = cin->remaining; gm = cin->gm; }
    u32 wb = 0;  // the window: records wb .. wb + 31
    bool inRun = false;
    u32 wNh = 0;
    u64 wMask = 0;
    u32 j = i;
    for (; j < h.n; ++j) {
      if (j == i || j - wb >= 32u) {
        wb = j;
        const u32 q = j + lane;
        inRun = q < h.n && (q == i || normKey(h.key[q]) == k);
        wNh = inRun ? h.nh[q] : 0u;
        wMask = (inRun && wNh > 1) ? maskAt(q) : 0ull;
      }
      const u32 src = j - wb;
      if (!__shfl_sync(FULL, inRun ? 1u : 0u, src)) break;  // end of the run
      const u32 nhj = __shfl_sync(FULL, wNh, src);
      if (!(nhj > 1)) continue;
      const u64 mj = __shfl_sync(FULL, wMask, src);
      if (routed) {
        if (lane == 0) slowAppend(slow, ctl, k, ctl->ordBase + j, mj, nhj);
        continue;
      }
      if (!isOpen) { isOpen = true; fromCarry = false; remaining = nhj - 1; gm = mj; first = j; ++nReads; }
      else { --remaining; gm |= mj; }
      if (remaining == 0) {
        if (gm != 0) {
          if (r.rescue) gm = rescueGroup(first, j + 1, gm);  // never a carried read: rescue mode does not carry
          if (lane == 0) tableAdd(table, gm, 1);
          if (__popcll(gm) == 1) ++nRescued;
        }
        isOpen = false;
      }
    }
    if (!isOpen) return;
    if (routed) return;  // (a carried read is never routed: its name would have been deferred instead of carried)
    if (j >= h.n && !r.rescue) {  // the run reaches the end of the batch: carry the open read over
      if (lane == 0) {
        Carry &c = ctl->carry[(seq + 1) & 1];
        c.key = k; c.gm = gm; c.remaining = remaining;
        c.ord = fromCarry ? cin->ord : ctl->ordBase + first;
        __threadfence();
        c.valid = 1;
      }
      return;
    }
    // unfinished read: the deferred path opens it again
    --nReads;
    if (lane == 0) {
      if (fromCarry) slowAppend(slow, ctl, k, cin->ord, cin->gm, cin->remaining + 1);  // stands for the records of earlier batches
      for (u32 q = fromCarry ? i : first; q < j; ++q) {
        const u32 nhq = h.nh[q];
        if (nhq > 1) slowAppend(slow, ctl, k, ctl->ordBase + q, maskAt(q), nhq);
      }
      keySetInsert(open, k, seq, ctl);
      ctl->dirty = 1;
    }
  }
};

// The runs k_batch_lean marked (walkMap: bit i = "the run whose first record to look at is i"), one WARP per run, and the read
// carried into the batch (warp 0).  Counts go straight to the device table; the words of the bitmap are cleared as they are read
// (four per lane and round; the map is allocated with the padding for that).
template <int MODE>
__global__ void __launch_bounds__(256)
k_batch_walk(const __grid_constant__ IndexView ix, const __grid_constant__ FastView fx, const __grid_constant__ HitView h, const __grid_constant__ Rules r,
             const __grid_constant__ TableView table, SampleCtl *ctl, const __grid_constant__ SlowView slow, const __grid_constant__ KeySetView open,
             u32 *walkMap, int defer) {  // (__grid_constant__: the walker holds references to the views -- no per-thread copies on the stack)
  constexpr u32 FULL = 0xffffffffu;
  const Annotator<MODE, true> annot{ix, fx, r.overlap};
  const u32 seq = ctl->batchSeq;
  const u32 lane = threadIdx.x & 31u;
  WarpWalker<MODE> w{h, r, annot, ctl, slow, open, table, seq, lane, 0, 0};
  const u32 gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
  if (gwarp == 0 && h.n) {
    const Carry &c = ctl->carry[seq & 1];
    if (c.valid) {
      if (defer) {  // a read carried out of a batch that still ran the countdown: it joins the deferred list like its later records
        --w.nReads;
        if (lane == 0) slowAppend(slow, ctl, c.key, c.ord, c.gm, c.remaining + 1);
      } else if (c.key == normKey(h.key[0])) {
        w.walk(0, c.key, &c);
      } else {  // its name does not continue: unfinished
        --w.nReads;
        if (lane == 0) {
          slowAppend(slow, ctl, c.key, c.ord, c.gm, c.remaining + 1);
          keySetInsert(open, c.key, seq, ctl);
          ctl->dirty = 1;
        }
      }
    }
  }
  const u32 nQuads = defer ? 0u : ((h.n + 31u) / 32u + 3u) / 4u;  // (the DEFER variant marks nothing)
  uint4 *const map4 = reinterpret_cast<uint4 *>(walkMap);
  for (u32 q0 = gwarp * 32u; q0 < nQuads; q0 += nWarps * 32u) {
    const u32 q = q0 + lane;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (q < nQuads) v = map4[q];
    const bool any = (v.x | v.y | v.z | v.w) != 0;
    if (any) map4[q] = make_uint4(0, 0, 0, 0);
    u32 lanes = __ballot_sync(FULL, any);
    while (lanes) {
      const int src = __ffs(lanes) - 1;
      lanes &= lanes - 1;
      const u32 word0 = (q0 + (u32)src) * 4u;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        u32 bits = __shfl_sync(FULL, c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w, src);
        while (bits) {
          const u32 i0 = (word0 + c) * 32u + (u32)(__ffs(bits) - 1);
          bits &= bits - 1;
          w.walk(i0, normKey(h.key[i0]), nullptr);
        }
      }
    }
  }
  if (lane == 0) {
    if (w.nReads) atomicAdd(&ctl->stats[ST_READS], (u64)(long long)w.nReads);
    if (w.nRescued) atomicAdd(&ctl->stats[ST_RESCUED], (u64)(long long)w.nRescued);
  }
}

}  // namespace mma
