// k_batch_lean: the batch kernel (K2 + K3 + K4) when the annotation has BIN ENTRIES (FastView::ent, annotations up to ~160 Mb):
// 32-bit element sets (E <= 30), -y default / unique / ratio.  Same decomposition as k_batch / k_batch_fast -- a warp owns a
// contiguous chunk of 128-hit warp tiles, 4 consecutive hits per lane, no block-wide barrier in the loop -- rebuilt around
// three changes:
//
//   stream    the five hit arrays of a tile are brought into a per-warp shared-memory ring by the TMA unit: one elected lane
//             issues five 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx, L2 evict-first) for the tile TWO tiles
//             ahead, the warp waits on the stage's mbarrier and reads its hits with conflict-free 128-bit shared loads.  No
//             global load of the stream goes through the LSU, and the first key of the next tile (needed to close a run that
//             ends on the tile border) is simply read from the next stage.
//   lookup    ONE 32-byte gather per hit: the bin entry of the read start holds the answers of the segment that covers the end
//             of the bin (in-segment and over the next boundary).  Reads that start before a boundary of their bin, or need an
//             answer the entry does not hold (~10 %), are compacted over the warp and take the segment record, one hit per lane.
//   per read  the segmented OR scan over the runs of a tile tests "distance to the nearest run start <= d" per round instead of
//             re-deriving it from the ballot; counters are kept per lane in plain 32-bit registers.
//
// Everything that is not the regular shape (-m rescue, unfinished read names, runs whose NH disagrees with their length, runs cut
// by a chunk border) takes the serial RunWalker of mma_device.cuh exactly as in k_batch: results are identical by construction
// and checked against the oracle by the same tests.
#pragma once
#include "mma_batch_fast.cuh"

namespace mma {

#ifndef MMA_LEAN_THREADS
#define MMA_LEAN_THREADS 352
#endif
#ifndef MMA_LEAN_BLOCKS_PER_SM
#define MMA_LEAN_BLOCKS_PER_SM 2
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
__device__ __forceinline__ void ldEntry(const uint4 *p, uint4 &lo, uint4 &hi) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
               : "l"(p));
}

template <bool HIST>
struct LeanSmem {
  alignas(128) unsigned char ring[LEAN_WARPS][LEAN_STAGES][LEAN_STAGE_BYTES];
  alignas(8) u64 bar[LEAN_WARPS][LEAN_STAGES];
  typename BlockTableOf<HIST, LEAN_BT_SLOTS>::type bt;
  unsigned short hist[HIST ? HIST_ROWS : 1][LEAN_THREADS];
  uint2 chrInfo[CHR_SMEM + 1];
  u32 scratch[LEAN_WARPS][WT_HITS];  // answers of the compacted hits; then the first records of the runs to walk
  unsigned char slowQ[LEAN_WARPS][WT_HITS];
  u32 stat[ST_N];
};

template <bool HIST>
struct LeanCount {  // one read counted for an element set, from divergent code (the serial walker)
  LeanSmem<HIST> &sm;
  const TableView &table;
  u32 tid;
  __device__ __forceinline__ void operator()(u64 ckey) const {
    if (HIST) {
      const u32 c = (u32)ckey;
      if (c == 0) return;
      if (c & (c - 1)) sm.bt.add(ckey, 1, table);
      else sm.hist[__ffs(c) - 1][tid] += 1;
    } else if (ckey) {
      sm.bt.add(ckey, 1, table);
    }
  }
};

// GROUPS: runs of k x NH records (paired-end data) are resolved in parallel as k reads (see k_batch_fast)
template <int MODE, int STRAT, bool GROUPS>
__global__ void __launch_bounds__(LEAN_THREADS, MMA_LEAN_BLOCKS_PER_SM)
k_batch_lean(const __grid_constant__ IndexView ix, const __grid_constant__ FastView fx, const __grid_constant__ HitView h, const __grid_constant__ Rules r,
             const __grid_constant__ TableView table, SampleCtl *ctl, const __grid_constant__ SlowView slow, const __grid_constant__ KeySetView open) {
  constexpr bool HIST = (STRAT != 3);
  constexpr u32 FULL = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char leanSmemRaw[];
  LeanSmem<HIST> &sm = *reinterpret_cast<LeanSmem<HIST> *>(leanSmemRaw);
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  sm.bt.init();
  if (HIST) {
#pragma unroll
    for (int e = 0; e < HIST_ROWS; ++e) sm.hist[e][tid] = 0;
  }
  if (tid < ST_N) sm.stat[tid] = 0;
  for (u32 c = tid; c <= fx.nChr; c += LEAN_THREADS) sm.chrInfo[c] = fx.chrInfo[c];  // launched only when nChr <= CHR_SMEM
  const u32 bar0 = smemAddr(&sm.bar[warp][0]), ring0 = smemAddr(&sm.ring[warp][0][0]);
  if (lane == 0) {
    mbarInit(bar0, 1);
    mbarInit(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const Annotator<MODE, true> annot{ix, fx, r.overlap};
  const u32 seq = ctl->batchSeq;
  // every run takes the serial walker when rescue() needs multiplicities or some read name is known as unfinished
  const bool forceWalk = (STRAT == 0) && (r.rescue || __shfl_sync(FULL, ctl->openCount, 0) != 0);
  u32 cAsg = 0, cUniq = 0, cMulti = 0, cAmbi = 0, cHits = 0, cMiss = 0, cClosed = 0, cResc = 0, pWalks = 0;

  LeanCount<HIST> count{sm, table, tid};
  RunWalker<MODE, true, LeanCount<HIST>> w{h, r, annot, ctl, slow, open, count, seq, 0u, 0u};

  const u32 nWT = (h.n + WT_HITS - 1) / WT_HITS;
  const u32 nWarps = gridDim.x * LEAN_WARPS;
  const u32 per = (nWT + nWarps - 1) / nWarps;
  const u32 t0 = min(nWT, (blockIdx.x * LEAN_WARPS + warp) * per), t1 = min(nWT, t0 + per);
  const u32 nMax = fx.nChr;
  const u64 pol = policyEvictFirst();

  // ---- staging: tile t of the chunk goes to stage (t - t0) & 1.  Full tiles by bulk copies, the (one) partial tile at the end
  //      of the batch by the warp itself.
  auto stage = [&](u32 t) {
    const u32 s = (t - t0) & 1u;
    const u32 dst = ring0 + s * LEAN_STAGE_BYTES;
    const size_t base = (size_t)t * WT_HITS;
    if ((t + 1) * WT_HITS <= h.n) {
      if (lane == 0) {
        const u32 bar = bar0 + s * 8;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the warp's reads of the stage come before the unit's writes
        mbarExpectTx(bar, (STRAT == 0) ? LEAN_STAGE_BYTES : 2048u);
        bulkLoad(dst, h.start + base, 512, bar, pol);
        bulkLoad(dst + 512, h.end + base, 512, bar, pol);
        bulkLoad(dst + 1024, h.meta + base, 512, bar, pol);
        bulkLoad(dst + 1536, h.nh + base, 512, bar, pol);
        if (STRAT == 0) bulkLoad(dst + 2048, h.key + base, 1024, bar, pol);
      }
    } else {
      u32 *d32 = reinterpret_cast<u32 *>(&sm.ring[warp][s][0]);
      u64 *d64 = reinterpret_cast<u64 *>(&sm.ring[warp][s][2048]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const u32 o = lane * 4 + j;
        const size_t i = base + o;
        const bool v = i < h.n;
        d32[o] = v ? h.start[i] : 0u;
        d32[128 + o] = v ? h.end[i] : 0u;
        d32[256 + o] = v ? h.meta[i] : 0x00FFFFFFu;
        d32[384 + o] = v ? h.nh[i] : 1u;
        if (STRAT == 0) d64[o] = v ? h.key[i] : KEY_EMPTY;
      }
    }
  };

  bool cValid = false, cCont = false;  // cCont: the run open at the end of the tile continues in the next tile
  u32 cStart = 0, cTot = 0, cNh = 0;
  u64 cKey = KEY_EMPTY;
  u64 chunkPeek = KEY_EMPTY;  // first key after the chunk
  if (STRAT == 0 && t1 > t0 && (size_t)t1 * WT_HITS < h.n) chunkPeek = __ldg(&h.key[(size_t)t1 * WT_HITS]);
  if (t0 < t1) stage(t0);
  if (t0 + 1 < t1) stage(t0 + 1);
  __syncwarp();

  for (u32 t = t0; t < t1; ++t) {
    const u32 s = (t - t0) & 1u, use = (t - t0) >> 1;
    const u32 base = t * WT_HITS + lane * 4;
    const bool fullTile = (t + 1) * WT_HITS <= h.n;
    if (fullTile) mbarWait(bar0 + s * 8, use & 1u);
    u32 rs[4], re[4], meta[4], nh[4];
    u64 key[4];
    u32 validBits = 15u;
    {
      const unsigned char *st = &sm.ring[warp][s][0];
      const uint4 a = *reinterpret_cast<const uint4 *>(st + lane * 16);
      const uint4 b = *reinterpret_cast<const uint4 *>(st + 512 + lane * 16);
      const uint4 c = *reinterpret_cast<const uint4 *>(st + 1024 + lane * 16);
      const uint4 d = *reinterpret_cast<const uint4 *>(st + 1536 + lane * 16);
      rs[0] = a.x; rs[1] = a.y; rs[2] = a.z; rs[3] = a.w;
      re[0] = b.x; re[1] = b.y; re[2] = b.z; re[3] = b.w;
      meta[0] = c.x; meta[1] = c.y; meta[2] = c.z; meta[3] = c.w;
      nh[0] = d.x; nh[1] = d.y; nh[2] = d.z; nh[3] = d.w;
      if (STRAT == 0) {
        const ulonglong2 k0 = *reinterpret_cast<const ulonglong2 *>(st + 2048 + lane * 32);
        const ulonglong2 k1 = *reinterpret_cast<const ulonglong2 *>(st + 2048 + lane * 32 + 16);
        key[0] = k0.x; key[1] = k0.y; key[2] = k1.x; key[3] = k1.y;
        // the all-ones key is reserved (normKey): only a tile that holds a key with all-ones upper half needs the fix-up
        const u32 hiMax = max(max((u32)(k0.x >> 32), (u32)(k0.y >> 32)), max((u32)(k1.x >> 32), (u32)(k1.y >> 32)));
        if (__any_sync(FULL, hiMax == 0xFFFFFFFFu)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) key[j] = normKey(key[j]);
        }
      }
      if (!fullTile) {
        const u32 left = (h.n > base) ? min(h.n - base, 4u) : 0u;
        validBits = (1u << left) - 1u;
      }
    }
    // ---- run starts
    u32 hbits = 0, F = 0;
    const Carry *carryIn = nullptr;
    u64 nextKey = KEY_EMPTY;
    if (STRAT == 0) {
      u64 prev = __shfl_up_sync(FULL, key[3], 1);
      if (lane == 0) {
        if (t != t0) prev = cKey;
        else if (base == 0) {
          const Carry &c = ctl->carry[seq & 1];
          prev = KEY_EMPTY;
          if (c.valid) { carryIn = &c; prev = c.key; }
        } else prev = normKey(h.key[base - 1]);
      }
      hbits = ((key[0] != prev) ? 1u : 0u) | ((key[1] != key[0]) ? 2u : 0u) | ((key[2] != key[1]) ? 4u : 0u) | ((key[3] != key[2]) ? 8u : 0u);
      hbits |= ~validBits & 15u;  // (slots past the end of the batch count as run starts)
      F = __ballot_sync(FULL, hbits != 0);
      nextKey = __shfl_sync(FULL, key[3], 31);
    }
    // ---- which hits are looked at (unique: only NH == 1, mm:1773)
    u32 visBits = validBits;
    if (STRAT == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (nh[j] != 1) visBits &= ~(1u << j);
    }
    // ---- lookup: the bin entry of the read start
    u32 m[4];
    u32 slowBits = 0;
    {
      uint4 e0[4], e1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint2 ci = sm.chrInfo[min(meta[j] & 0x00FFFFFFu, nMax)];
        ldEntry(&fx.ent[2u * (ci.x + min(rs[j] >> 6, ci.y - 1u))], e0[j], e1[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool fwd = (int)meta[j] < 0;
        const u64 bits = ((u64)e0[j].y << 32) | e0[j].x;
        const bool inZ = ((bits >> (rs[j] & 63u)) >> 1) == 0;  // no boundary after the read start inside its bin
        u32 a = fwd ? e1[j].x : e1[j].y;
        if (re[j] > e0[j].w) {  // over the end of Z: the cross answer, when the read ends inside the next segment
          const u32 len = (e1[j].z >> 28) | ((e1[j].w >> 28) << 4);
          const u32 x = (fwd ? e1[j].z : e1[j].w) & ENT_XNONE;
          a = (MODE == 0 && re[j] - e0[j].w <= len && x != ENT_XNONE) ? x : ENT_NONE;
        }
        bool pass = true;
        if (MODE != 0) {  // no feature can overlap the read by more than end - start
          const u32 o = re[j] - rs[j];
          if (MODE == 1) pass = (o != 0) && (__fmul_rn((float)(o + 1u), r.overlap) <= (float)o);
          else pass = (o != 0) && ((float)o >= r.overlap);
        }
        const bool degen = re[j] < rs[j] || re[j] >= 0xFFFFFFF0u;
        const bool vis = (visBits >> j) & 1u;
        const bool ok = inZ && a != ENT_NONE && !degen;
        m[j] = (vis && pass && ok) ? a : 0u;
        if (vis && !ok && (pass || degen)) slowBits |= 1u << j;
      }
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
        if ((slowBits >> j) & 1u) sm.slowQ[warp][off[j]] = (unsigned char)(lane * 4 + j);
      __syncwarp();
      const u32 *st32 = reinterpret_cast<const u32 *>(&sm.ring[warp][s][0]);
      for (u32 k = lane; k < total; k += 32) {
        const u32 id = sm.slowQ[warp][k];
        sm.scratch[warp][id] = slowAnnotate<MODE>(fx, ix, st32[id], st32[128 + id], st32[256 + id], r.overlap);
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if ((slowBits >> j) & 1u) {
          const u32 a = sm.scratch[warp][lane * 4 + j];
          m[j] = a & ~FAST_MISS;
          cMiss += a >> 31;
        }
    }
    __syncwarp();
    // the stage is free: bring in the tile two tiles ahead
    if (t + 2 < t1) stage(t + 2);
    // ---- per-hit counters (mm:1666-1668); a visited hit that does not join the by-name countdown is a read of its own
    u32 ev[4];  // the element set counted at this hit slot (0 = none)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool multi = STRAT == 0 && nh[j] > 1;
      const bool amb = (m[j] & (m[j] - 1u)) != 0;
      cAsg += m[j] != 0 ? 1u : 0u;
      cUniq += (m[j] != 0 && !amb && nh[j] == 1) ? 1u : 0u;
      cMulti += multi ? 1u : 0u;
      cAmbi += amb ? 1u : 0u;
      ev[j] = multi ? 0u : m[j];
      if (r.rescue) ev[j] = (u32)rescueSingle(r, (u64)ev[j]);
    }
    cHits += __popc(visBits);
    u32 nWalk = 0, closeBits = 0;
    u32 inc = 0, lastHeadPos = 0, F2 = 0;
    bool serialTile = false;
    u32 tileEndsRun = 1;  // the record after the tile's last one starts another run (or the batch ends there)
    if (STRAT == 0) {
      // first key of the next tile: from the next stage of the ring, or (last tile of the chunk) fetched when the chunk began
      const u32 nextTile = (t + 1) * WT_HITS;
      if (nextTile < h.n) {
        u64 pk = chunkPeek;
        if (t + 1 < t1) {
          if ((t + 2) * WT_HITS <= h.n) mbarWait(bar0 + (s ^ 1u) * 8, ((t + 1 - t0) >> 1) & 1u);
          pk = *reinterpret_cast<const u64 *>(&sm.ring[warp][s ^ 1u][2048]);
        }
        tileEndsRun = (normKey(pk) != nextKey) ? 1u : 0u;
      }
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
      const u32 before = F & ((1u << lane) - 1u);
      // distance (in lanes) to the nearest lane, this one included, in which a run starts
      const u32 upTo = F & (0xFFFFFFFFu >> (31u - lane));
      u32 prevNh = __shfl_up_sync(FULL, nh[3], 1);
      if (lane == 0) prevNh = cNh;
      lastHeadPos = base + (31 - __clz(hbits | 1u));
      const u32 sPrev = __shfl_sync(FULL, lastHeadPos, before ? (31 - __clz(before)) : 0);
      const u32 nextHead0 = __shfl_down_sync(FULL, hbits & 1u, 1);
      // bit j: the next record starts another run, i.e. this record ends its run (the tile's last record: from the peek)
      const u32 lastBits = ((hbits >> 1) | ((lane < 31u ? nextHead0 : tileEndsRun) << 3)) & validBits;
      // a run that starts before this lane's hits: its first record, and whether this warp owns it at all
      const u32 inStart = before ? sPrev : cStart;
      const bool inMine = before || cValid;  // else the run starts in another warp's chunk: that warp finishes it
      if (GROUPS) {
        // A read opens at a record with NH = n > 1 and takes the n - 1 records of its name that follow, so a run of records
        // sharing a read key and carrying the same NH = n is a sequence of GROUPS of n records, one read each (one group for
        // single-end data, two -- the two mates -- for paired-end data, mm:1673-1681).  The element set of a read is the
        // union over its group: a segmented OR scan over the 128 hits of the warp tile, segments starting at run starts and
        // at every n-th record of a run, seeded with the state carried from the previous tile; the lane owning the LAST
        // record of a group counts the read.  A run that ends inside a group leaves an unfinished read: serial walk
        // (RunWalker) from the start of that group.  A tile in which NH changes inside a run -- or any tile while rescue() needs
        // multiplicities or some read name is known as unfinished -- is resolved serially: one walk per run (serialTile).
        const u32 badBits = ((!(hbits & 1u) && nh[0] != prevNh) ? 1u : 0u) | ((!(hbits & 2u) && nh[1] != nh[0]) ? 2u : 0u) |
                            ((!(hbits & 4u) && nh[2] != nh[1]) ? 4u : 0u) | ((!(hbits & 8u) && nh[3] != nh[2]) ? 8u : 0u);
        serialTile = forceWalk || __any_sync(FULL, (badBits & validBits) != 0);
        if (!serialTile) {
          u32 off[4];  // position of the record inside its group
          u32 hb2 = hbits, endBits = 0, tailBits = 0;
          bool longRun = false;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const u32 hbLe = hbits & ((2u << j) - 1u);
            const u32 runStart = hbLe ? base + (31 - __clz(hbLe)) : inStart;
            const bool mine = (hbLe || inMine) && ((validBits >> j) & 1u) && nh[j] > 1;
            u32 o = base + j - runStart;
            if (o >= nh[j]) o -= nh[j];
            if (mine && o >= nh[j]) longRun = true;  // third group or later: exact remainder below
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
            const bool mine = (hbLe || inMine) && ((validBits >> j) & 1u) && nh[j] > 1;
            if (mine && off[j] == 0) hb2 |= 1u << j;                                  // first record of a group
            if (mine && off[j] + 1 == nh[j]) endBits |= 1u << j;                      // last record of a group
            else if (mine && ((lastBits >> j) & 1u)) tailBits |= 1u << j;             // the run ends inside a group
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
          for (int j = 0; j < 4; ++j) {
            if ((endBits >> j) & 1u) {
              ev[j] = (hb2 & ((2u << j) - 1u)) ? pre[j] : (inTot | pre[j]);
              closeBits |= 1u << j;
            }
            if ((tailBits >> j) & 1u) sm.scratch[warp][lane * 4 + nWalk++] = base + j - off[j];
          }
        } else {
          // serial tile: the run carried into the tile (from the start of its open group) and every run that starts in it
          if (lane == 0 && cValid && !(hbits & 1u) && (validBits & 1u)) {
            const u32 o = base - cStart;
            sm.scratch[warp][lane * 4 + nWalk++] = base - ((cNh > 1) ? o % cNh : 0u);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (((hbits & validBits) >> j) & 1u) sm.scratch[warp][lane * 4 + nWalk++] = base + j;
        }
      } else {
        // A run of n records that all carry NH = n (> 1) is one read; its element set is the union over the run: a
        // segmented OR scan over the 128 hits of the warp tile (bit 31 of the scanned word = "irregular": NH changes inside
        // the run, or every run has to be walked), seeded with the state carried from the previous tile.  The lane owning
        // the run's LAST record closes it; irregular runs take the serial walk (RunWalker), started by the same lane.
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
        const u32 hd = upTo ? (u32)__clz(upTo) - (31u - lane) : 64u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const u32 tt = __shfl_up_sync(FULL, inc, d);
          if ((u32)d <= hd) inc |= tt;
        }
        u32 X = __shfl_up_sync(FULL, inc, 1);
        if (lane == 0) X = 0;
        const u32 inTot = before ? X : (cTot | X);  // union so far of a run that starts before this lane's hits
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool last = (lastBits >> j) & 1u;
          const u32 hbLe = hbits & ((2u << j) - 1u);
          // union of the run's element sets and its first record, wherever the run starts
          const u32 tot = hbLe ? pre[j] : (inTot | pre[j]);
          const u32 runStart = hbLe ? base + (31 - __clz(hbLe)) : inStart;
          const bool mine = hbLe || inMine;
          const bool flagged = (tot & 0x80000000u) != 0;
          // a run of reads that are their own group (NH <= 1 throughout) has nothing to close
          const bool irregular = flagged || (nh[j] > 1 && nh[j] != base + j + 1 - runStart);
          if (last && mine && irregular) sm.scratch[warp][lane * 4 + nWalk++] = runStart;
          if (last && mine && !irregular && nh[j] > 1) {
            ev[j] = tot & 0x7FFFFFFFu;
            closeBits |= 1u << j;
          }
        }
      }
    }
    // ---- counting: single-element sets into the lane's histogram column, the rest into the block table
    {
      u32 pend = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const u32 c = ev[j];
        const bool single = c != 0 && (c & (c - 1)) == 0;
        if (single && ((closeBits >> j) & 1u)) ++cResc;  // a multi-mapping read resolved to one element (mm:1691)
        if (HIST) {
          const u32 row = single ? (u32)(__ffs(c) - 1) : (u32)(HIST_ROWS - 1);  // E <= 30: the last row only ever receives zeros
          sm.hist[row][tid] += single ? 1 : 0;
          if (c != 0 && !single) pend |= 1u << j;
        } else {
          if (c != 0) pend |= 1u << j;
        }
      }
      cClosed += __popc(closeBits);
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
    if (STRAT == 0) {
      pWalks += nWalk;
#pragma unroll 1
      for (u32 q = 0; q < nWalk; ++q) { const u32 i0 = sm.scratch[warp][lane * 4 + q]; w.walk(i0, normKey(h.key[i0]), nullptr); }
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
      cKey = nextKey;
      cCont = tileEndsRun == 0;
    }
  }
  // ---- a run open at the end of the chunk continues in another warp's chunk: its remaining reads (GROUPS: from the group that is
  //      open there, or starts there) are finished by the serial walk.  (A run ending exactly at the chunk's last record was
  //      closed above, like the last run of the batch.)
  if (STRAT == 0 && cValid && cCont && t1 > t0 && lane == 0) {
    if (GROUPS) {
      const u32 next = t1 * WT_HITS, o = next - cStart;
      w.walk(next - ((cNh > 1) ? o % cNh : 0u), cKey, nullptr);
    } else {
      w.walk(cStart, cKey, nullptr);
    }
  }
  u32 cUnassigned = cHits - cAsg;
  u32 cReads = cHits - cMulti + cClosed + w.nReads, cRescued = cResc + w.nRescued;

  // ---- block epilogue: counters, the private histogram columns and the private table
  cHits = __reduce_add_sync(FULL, cHits); cUnassigned = __reduce_add_sync(FULL, cUnassigned);
  cAmbi = __reduce_add_sync(FULL, cAmbi); cUniq = __reduce_add_sync(FULL, cUniq);
  cMulti = __reduce_add_sync(FULL, cMulti); cReads = __reduce_add_sync(FULL, cReads);
  cRescued = __reduce_add_sync(FULL, cRescued); cMiss = __reduce_add_sync(FULL, cMiss);
  if (STRAT == 0 && !forceWalk) {
    pWalks = __reduce_add_sync(FULL, pWalks);
    if (lane == 0 && pWalks) atomicAdd(&ctl->walkCount, pWalks);
  }
  if (lane == 0) {
    atomicAdd(&sm.stat[ST_HITS], cHits); atomicAdd(&sm.stat[ST_UNASSIGNED], cUnassigned); atomicAdd(&sm.stat[ST_AMBIGUOUS], cAmbi);
    atomicAdd(&sm.stat[ST_UNIQUE], cUniq); atomicAdd(&sm.stat[ST_MULTIPLE], cMulti); atomicAdd(&sm.stat[ST_READS], cReads);
    atomicAdd(&sm.stat[ST_RESCUED], cRescued); atomicAdd(&sm.stat[7], cMiss);
  }
  __syncthreads();
  sm.bt.flush(table);
  if (HIST) {
    for (u32 e = warp; e < HIST_ROWS; e += LEAN_WARPS) {
      u32 v = 0;
#pragma unroll
      for (int q = 0; q < LEAN_WARPS; ++q) v += sm.hist[e][lane + 32 * q];
      v = __reduce_add_sync(FULL, v);
      if (lane == 0 && v) tableAdd(table, 1ull << e, v);
    }
  }
  if (tid < 7) {
    const int sv = (int)sm.stat[tid];  // reads / rescued can be negative within a block (unfinished reads)
    if (sv) atomicAdd(&ctl->stats[tid], (u64)(long long)sv);
  }
  if (tid == 7 && sm.stat[7]) atomicAdd(&ctl->fastMiss, sm.stat[7]);
}

}  // namespace mma
