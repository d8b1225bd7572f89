// k_batch_fast: the batch kernel (K2 + K3 + K4) for the common configuration -- segment answer table built (E <= 30),
// 32-bit element sets, -y default / unique / ratio.  Same decomposition as k_batch (mma_device.cuh): a warp owns a
// contiguous chunk of 128-hit warp tiles, 4 consecutive hits per lane, no block-wide barrier in the loop.  What differs
// is the shape of the per-tile code, written so that the warp stays converged and its memory requests overlap:
//
//   phase A   4 independent gathers per lane: the position-map entries {bitmap, rank} of the read starts (unconditional,
//             dummy address for hits that need no lookup) -> rank + popc = index of the segment holding the start
//   phase B   4 independent 256-bit gathers: the 32-byte records of those segments
//   phase C   pick the in-segment / cross-segment answer; whatever is left (upstream/downstream ties, other position-
//             dependent picks, reads over three or more segments, degenerate intervals: < 1 % of the hits) is COMPACTED
//             over the warp through shared memory and evaluated out of line, one hit per lane
//   counting  every hit slot yields at most one count event (a read that is its own group, or the read whose group ends
//             at this record); single-element sets go to the lane's private histogram column with one unconditional
//             read-modify-write, the others to the block table in a loop the whole warp runs in lockstep
//
// Two variants of the per-read step (template flag GROUPS): the plain one closes a run of n records carrying NH = n as one
// read; the GROUPS one cuts a run of k x n records into k reads inside the scan (paired-end data) and resolves a tile in
// which NH changes inside a run serially.  Whatever does not fit (-m rescue, batches following one that left unfinished read
// names, runs ending inside a group, runs cut by a chunk border, the read carried into the batch) takes the serial walk of
// RunWalker (mma_device.cuh) as in k_batch -- but in k_batch_walk (mma_batch_lean.cuh), launched behind this kernel: a lane
// only marks the run's first record in a bitmap.
#pragma once
#include "mma_device.cuh"

namespace mma {

#ifndef MMA_FAST_THREADS
#define MMA_FAST_THREADS 256
#endif
#ifndef MMA_FAST_BLOCKS_PER_SM
#define MMA_FAST_BLOCKS_PER_SM 3
#endif
#define FAST_THREADS MMA_FAST_THREADS
#define FAST_WARPS (FAST_THREADS / 32)

#ifdef MMA_DIAG
// tuning builds only: why hits leave the table fast path.  0 degenerate, 3 start beyond the looked-up segment (coarse granules),
// 4 read over 3+ segments (or 2 under an overlap mode), 6 answer GENERAL, 7 answered in-segment, 8 answered cross-segment,
// 9 hits looked up
__device__ unsigned long long g_diag[16];
#define DIAG(i, c) do { if (c) atomicAdd(&g_diag[i], 1ull); } while (0)
#else
#define DIAG(i, c) do { } while (0)
#endif

// one 256-bit read-only gather of a 32-byte segment record (LDG.E.256, sm_100+)
__device__ __forceinline__ void ldRecord(const uint4 *p, uint4 &lo, uint4 &hi) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
               : "l"(p));
}
#define CHR_SMEM 512  // the chromosome table lives in shared memory: annotations with more chromosomes take the general k_batch

template <int MODE>
__device__ __noinline__ u32 slowAnnotate(const FastView &fx, const IndexView &ix, u32 rs, u32 re, u32 meta, float ovl) {
  return fastAnnotate<MODE>(fx, ix, rs, re, meta, ovl);
}

// Block-private table for 32-bit element sets (multi-element sets only: a few thousand distinct keys per sample);
// flushed once per block.  Sized so that it does not fill up: a full neighbourhood falls through to the device table.
template <int SLOTS>
struct BlockTable32 {
  u32 keys[SLOTS];
  u32 cnt[SLOTS];
  __device__ void init() {
    for (int i = threadIdx.x; i < SLOTS; i += blockDim.x) { keys[i] = 0; cnt[i] = 0; }
  }
  __device__ __forceinline__ void add(u64 ckey64, u32 n, const TableView &g) {
    const u32 ckey = (u32)ckey64;
    u32 slot = (ckey * 0x9E3779B1u) >> (32 - __builtin_ctz(SLOTS));
#pragma unroll 1
    for (int probe = 0; probe < 16; ++probe) {
      u32 old = keys[slot];
      if (old != ckey) {
        if (old == 0) old = atomicCAS(&keys[slot], 0u, ckey);
        if (old != 0 && old != ckey) { slot = (slot + 1) & (SLOTS - 1); continue; }
      }
      atomicAdd(&cnt[slot], n);
      return;
    }
    tableAdd(g, ckey64, n);
  }
  __device__ void flush(const TableView &g) {
    for (int i = threadIdx.x; i < SLOTS; i += blockDim.x)
      if (cnt[i]) tableAdd(g, (u64)keys[i], cnt[i]);
  }
};

template <bool HIST, int SLOTS> struct BlockTableOf { typedef BlockTable<SLOTS> type; };
template <int SLOTS> struct BlockTableOf<true, SLOTS> { typedef BlockTable32<SLOTS> type; };

template <bool HIST, int SLOTS>
struct FastSmem {
  typename BlockTableOf<HIST, SLOTS>::type bt;
  unsigned short hist[HIST ? HIST_ROWS : 1][FAST_THREADS];
  u32 walkQ[4][FAST_THREADS];
  uint2 chrInfo[CHR_SMEM];
  u32 slowRes[FAST_WARPS][WT_HITS];
  unsigned char slowQ[FAST_WARPS][WT_HITS];
  u32 stat[ST_N];
};

template <bool HIST, int SLOTS>
struct FastCount {  // one read counted for an element set, from divergent code (unused since the serial walker left the kernel)
  FastSmem<HIST, SLOTS> &sm;
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

// GROUPS: runs of k x NH records (paired-end data) are resolved in parallel as k reads; costs registers, so the host only
// switches to this variant when a batch has sent many runs to the serial walker (walk counter of the control block)
template <int MODE, int STRAT, bool GROUPS>
__global__ void __launch_bounds__(FAST_THREADS, MMA_FAST_BLOCKS_PER_SM)
k_batch_fast(const __grid_constant__ IndexView ix, const __grid_constant__ FastView fx, const __grid_constant__ HitView h, const __grid_constant__ Rules r,
             const __grid_constant__ TableView table, SampleCtl *ctl, const __grid_constant__ SlowView slow, u32 *walkMap) {
  constexpr bool HIST = (STRAT != 3);
  constexpr int SLOTS = (STRAT == 3) ? 1024 : 2048;
  constexpr u32 FULL = 0xffffffffu;
  __shared__ FastSmem<HIST, SLOTS> sm;
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  sm.bt.init();
  if (HIST) {
#pragma unroll
    for (int e = 0; e < HIST_ROWS; ++e) sm.hist[e][tid] = 0;
  }
  if (tid < ST_N) sm.stat[tid] = 0;
  for (u32 c = tid; c < fx.nChr; c += FAST_THREADS) sm.chrInfo[c] = fx.chrInfo[c];  // launched only when nChr <= CHR_SMEM
  __syncthreads();
  const u32 seq = ctl->batchSeq;
  // every run takes the serial walker when rescue() needs multiplicities or some read name is known as unfinished
  const bool forceWalk = (STRAT == 0) && (r.rescue || __shfl_sync(FULL, ctl->openCount, 0) != 0);
  // per-thread counters, two 16-bit fields per register (a thread sees < 2^15 hits of one batch, see k_batch)
  u32 pAsgUniq = 0;   // assigned hits | hits with NH = 1 and exactly one element << 16        (mm:1666, 1668)
  u32 pMultAmbi = 0;  // hits joining the by-name countdown | ambiguous hits << 16               (mm:1670, 1667)
  u32 pHitsMiss = 0;  // hits looked at | segment-table misses << 16
  u32 pClosResc = 0;  // multi-mapping reads closed by the parallel countdown | of which rescued << 16
  u32 pWalks = 0;     // serial walks started (feeds the host's choice between the two variants of this kernel)

  // a run the scan cannot close, a run cut by the chunk border and the read carried into the batch are k_batch_walk's (launched
  // behind this kernel, see mma_batch_lean.cuh): bit i of walkMap = "walk the run from record i".  No call in the tile loop.
  auto queueWalk = [&](u32 i0) {
    if (i0 < h.n) atomicOr(&walkMap[i0 >> 5], 1u << (i0 & 31u));
  };

  const u32 nWT = (h.n + WT_HITS - 1) / WT_HITS;
  const u32 nWarps = gridDim.x * FAST_WARPS;
  const u32 per = (nWT + nWarps - 1) / nWarps;
  const u32 t0 = min(nWT, (blockIdx.x * FAST_WARPS + warp) * per), t1 = min(nWT, t0 + per);

  bool cValid = false, cCont = false;  // cCont: the run open at the end of the tile continues in the next tile
  u32 cStart = 0, cTot = 0, cNh = 0;
  u64 cKey = KEY_EMPTY;
  const u32 shift = fx.shift, gshift = fx.gshift;

  u64 peekKey = KEY_EMPTY;  // lane 31: first key of the tile after the current one
  if (STRAT == 0 && lane == 31u && t0 < t1 && (t0 + 1) * WT_HITS < h.n) peekKey = __ldg(&h.key[(t0 + 1) * WT_HITS]);

  for (u32 t = t0; t < t1; ++t) {
    const u32 base = t * WT_HITS + lane * 4;
    u32 rs[4], re[4], meta[4], nh[4];
    u64 key[4];
    u32 validBits;
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
        key[0] = k0.x; key[1] = k0.y; key[2] = k1.x; key[3] = k1.y;
        // the all-ones key is reserved (normKey): only a tile that holds a key with all-ones upper half needs the fix-up
        const u32 hiMax = max(max((u32)(k0.x >> 32), (u32)(k0.y >> 32)), max((u32)(k1.x >> 32), (u32)(k1.y >> 32)));
        if (__any_sync(FULL, hiMax == 0xFFFFFFFFu)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) key[j] = normKey(key[j]);
        }
      }
      validBits = 15u;
    } else {
      validBits = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const u32 i = base + j;
        rs[j] = 0; re[j] = 0; meta[j] = 0x00FFFFFFu; nh[j] = 1; key[j] = KEY_EMPTY;
        if (i < h.n) {
          validBits |= 1u << j;
          rs[j] = __ldcs(&h.start[i]); re[j] = __ldcs(&h.end[i]); meta[j] = __ldcs(&h.meta[i]); nh[j] = __ldcs(&h.nh[i]);
          if (STRAT == 0) key[j] = normKey(__ldcs(&h.key[i]));
        }
      }
    }
    // ---- run starts
    u32 hbits = 0, F = 0;
    u64 nextKey = KEY_EMPTY;
    u32 tileEndsRun = 1;  // lane 31: the record after the tile's last one starts another run (or the batch ends there)
    if (STRAT == 0) {
      // the first key of the next tile was requested one tile ago (peekKey); request the one after it now
      const u32 nextTile = (t + 1) * WT_HITS;
      if (lane == 31u && nextTile < h.n) tileEndsRun = (normKey(peekKey) != key[3]) ? 1u : 0u;
      if (lane == 31u && nextTile + WT_HITS < h.n) peekKey = __ldg(&h.key[nextTile + WT_HITS]);
      u64 prev = __shfl_up_sync(FULL, key[3], 1);
      if (lane == 0) {
        if (t != t0) prev = cKey;
        else if (base == 0) {
          const Carry &c = ctl->carry[seq & 1];
          prev = KEY_EMPTY;
          if (c.valid) prev = c.key;  // (the read carried into the batch is k_batch_walk's; here only: does its name continue?)
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
    // ---- phase A: position map entries of the read starts
    u32 m[4] = {0u, 0u, 0u, 0u};
    u32 lookBits = 0, slowBits = 0;
    uint2 en[4];
    u32 pm[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const u32 chr = meta[j] & 0x00FFFFFFu;
      const bool vis = ((visBits >> j) & 1u) && chr < fx.nChr;
      const bool degen = re[j] < rs[j] || re[j] >= 0xFFFFFFF0u;
      bool pass = true;
      if (MODE != 0) {  // no feature can overlap the read by more than end - start
        const u32 o = re[j] - rs[j];
        if (MODE == 1) pass = (o != 0) && (__fmul_rn((float)(o + 1u), r.overlap) <= (float)o);
        else pass = (o != 0) && ((float)o >= r.overlap);
      }
      if (vis && degen) slowBits |= 1u << j;
      DIAG(0, vis && degen);
      const bool look = vis && !degen && pass;
      if (look) lookBits |= 1u << j;
      const u32 chrSafe = look ? chr : 0u;
      const uint2 ci = sm.chrInfo[chrSafe];
      const u32 bRaw = rs[j] >> shift;
      en[j] = __ldg(&fx.bm[ci.x + min(bRaw, ci.y - 1u)]);
      const u32 p = (bRaw < ci.y) ? ((rs[j] >> gshift) & 31u) : 31u;
      pm[j] = (gshift == 0) ? (0xFFFFFFFFu >> (31u - p)) : ((1u << p) - 1u);
    }
    // ---- phase B: the 32-byte record of the segment holding the read start
    u32 tEnd[4], tLens[4], tAns[4], xAns[4], yAns[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool look = (lookBits >> j) & 1u;
      const bool fwd = (meta[j] >> 31) != 0;
      const u32 i = look ? en[j].y + __popc(en[j].x & pm[j]) : 0u;
      uint4 tt, xx;
      ldRecord(&fx.seg[2u * i], tt, xx);
      tEnd[j] = tt.x; tAns[j] = fwd ? tt.y : tt.z; tLens[j] = tt.w;
      xAns[j] = fwd ? xx.x : xx.y; yAns[j] = fwd ? xx.z : xx.w;
    }
    // ---- phase C: in-segment or cross-segment answer
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool look = (lookBits >> j) & 1u;
      const bool inT = re[j] <= tEnd[j];
      const int ahead = (MODE == 0 && !inT) ? segmentsAhead(re[j] - tEnd[j], tLens[j]) : 0;  // read over two / three segments
      const bool inX = ahead != 0;
      const u32 a = inT ? tAns[j] : (ahead == 2) ? yAns[j] : xAns[j];  // (an upstream/downstream tie, < 1 % of the hits, is settled out of line)
      const bool ok = rs[j] <= tEnd[j] && (inT || inX) && !(a & (ANS_VICPAIR | ANS_GENERAL));
      if (look && ok) m[j] = a;
      if (look && !ok) slowBits |= 1u << j;
      DIAG(9, look); DIAG(7, look && ok && inT); DIAG(8, look && ok && !inT); DIAG(3, look && rs[j] > tEnd[j]);
      DIAG(4, look && rs[j] <= tEnd[j] && !(inT || inX)); DIAG(6, look && rs[j] <= tEnd[j] && (inT || inX) && (a & ANS_GENERAL));
    }
    // ---- the rest, compacted over the warp
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
      for (u32 k = lane; k < total; k += 32) {
        const u32 id = sm.slowQ[warp][k];
        const u32 i = t * WT_HITS + id;
        sm.slowRes[warp][id] = slowAnnotate<MODE>(fx, ix, h.start[i], h.end[i], h.meta[i], r.overlap);
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if ((slowBits >> j) & 1u) {
          const u32 a = sm.slowRes[warp][lane * 4 + j];
          m[j] = a & ~FAST_MISS;
          pHitsMiss += (a >> 31) << 16;
        }
      __syncwarp();
    }
    // ---- per-hit counters (mm:1666-1668); a visited hit that does not join the by-name countdown is a read of its own
    //      Slots that are not visited (past the end of the batch, NH != 1 under -y unique) hold m = 0 and, past the end, NH = 1.
    //      Counted here: hits, ASSIGNED hits (unassigned = hits - assigned at the end), hits with NH = 1 and an element
    //      (unique, corrected below for the rare ambiguous hit), hits joining the by-name countdown.
    u32 ev[4];  // the element set counted at this hit slot (0 = none)
    u32 ambOr = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const u32 a1 = min(m[j], 1u);
      const bool multi = STRAT == 0 && nh[j] > 1;
      pAsgUniq += a1 + ((nh[j] == 1 ? a1 : 0u) << 16);
      pMultAmbi += multi ? 1u : 0u;
      ambOr |= m[j] & (m[j] - 1u);
      ev[j] = multi ? 0u : m[j];
      if (r.rescue) ev[j] = (u32)rescueSingle(r, (u64)ev[j]);
    }
    pHitsMiss += __popc(visBits);
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
    if (GROUPS) {
      if (STRAT == 0) {
        // ---- per-read countdown (mm:1669-1702)
        // A read opens at a record with NH = n > 1 and takes the n - 1 records of its name that follow, so a run of records
        // sharing a read key and carrying the same NH = n is a sequence of GROUPS of n records, one read each (one group for
        // single-end data, two -- the two mates -- for paired-end data, mm:1673-1681).  The element set of a read is the
        // union over its group: a segmented OR scan over the 128 hits of the warp tile, segments starting at run starts and
        // at every n-th record of a run, seeded with the state carried from the previous tile; the lane owning the LAST
        // record of a group counts the read.  A run that ends inside a group leaves an unfinished read: serial walk
        // (RunWalker) from the start of that group.  A tile in which NH changes inside a run -- or any tile while rescue() needs
        // multiplicities or some read name is known as unfinished -- is resolved serially: one walk per run (serialTile).
        u32 prevNh = __shfl_up_sync(FULL, nh[3], 1);
        if (lane == 0) prevNh = cNh;
        const u32 badBits = ((!(hbits & 1u) && nh[0] != prevNh) ? 1u : 0u) | ((!(hbits & 2u) && nh[1] != nh[0]) ? 2u : 0u) |
                            ((!(hbits & 4u) && nh[2] != nh[1]) ? 4u : 0u) | ((!(hbits & 8u) && nh[3] != nh[2]) ? 8u : 0u);
        const u32 before = F & ((1u << lane) - 1u);
        // (only in runs this warp owns: the run reaching into the chunk's first tile is the previous chunk's, NH carried or not)
        const u32 ownBits = (before || cValid) ? 0xFu : (hbits & 1u) ? 0xEu : (hbits & 2u) ? 0xCu : (hbits & 4u) ? 0x8u : 0u;
        serialTile = forceWalk || __any_sync(FULL, (badBits & validBits & ownBits) != 0);
        lastHeadPos = base + (31 - __clz(hbits | 1u));
        const u32 sPrev = __shfl_sync(FULL, lastHeadPos, before ? (31 - __clz(before)) : 0);
        // a run that starts before this lane's hits: its first record, and whether this warp owns it at all
        const u32 inStart = before ? sPrev : cStart;
        const bool inMine = before || cValid;  // else the run starts in another warp's chunk: that warp finishes it
        if (!serialTile) {
          const u32 nextHead0 = __shfl_down_sync(FULL, hbits & 1u, 1);
          // bit j: the next record starts another run, i.e. this record ends its run (the tile's last record: from the peek)
          const u32 lastBits = ((hbits >> 1) | ((lane < 31u ? nextHead0 : tileEndsRun) << 3)) & validBits;
          u32 off[4];           // position of the record inside its group
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
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const u32 tt = __shfl_up_sync(FULL, inc, d);
            if (lane >= (u32)d && ((F2 >> (lane - d + 1)) & ((1u << d) - 1u)) == 0) inc |= tt;
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
            if ((tailBits >> j) & 1u) sm.walkQ[nWalk++][tid] = base + j - off[j];
          }
        } else {
          // serial tile: the run carried into the tile (from the start of its open group) and every run that starts in it
          if (lane == 0 && cValid && !(hbits & 1u) && (validBits & 1u)) {
            const u32 o = base - cStart;
            sm.walkQ[nWalk++][tid] = base - ((cNh > 1) ? o % cNh : 0u);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (((hbits & validBits) >> j) & 1u) sm.walkQ[nWalk++][tid] = base + j;
        }
      }
    } else {
      if (STRAT == 0) {
        // ---- per-read countdown (mm:1669-1702)
        // A run of n records that all carry NH = n (> 1) is one read; its element set is the union over the run: a
        // segmented OR scan over the 128 hits of the warp tile (bit 31 of the scanned word = "irregular": NH changes inside
        // the run, or every run has to be walked), seeded with the state carried from the previous tile.  The lane owning
        // the run's LAST record closes it; irregular runs are marked for the serial walk (k_batch_walk) by the same lane.
        u32 prevNh = __shfl_up_sync(FULL, nh[3], 1);
        if (lane == 0) prevNh = cNh;
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
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const u32 tt = __shfl_up_sync(FULL, inc, d);
          if (lane >= (u32)d && ((F >> (lane - d + 1)) & ((1u << d) - 1u)) == 0) inc |= tt;
        }
        u32 X = __shfl_up_sync(FULL, inc, 1);
        if (lane == 0) X = 0;
        const u32 before = F & ((1u << lane) - 1u);
        lastHeadPos = base + (31 - __clz(hbits | 1u));
        const u32 sPrev = __shfl_sync(FULL, lastHeadPos, before ? (31 - __clz(before)) : 0);
        const u32 nextHead0 = __shfl_down_sync(FULL, hbits & 1u, 1);
        // bit j: the next record starts another run, i.e. this record ends its run (the tile's last record: from the peek)
        const u32 lastBits = ((hbits >> 1) | ((lane < 31u ? nextHead0 : tileEndsRun) << 3)) & validBits;
        // a run that starts before this lane's hits: union so far, first record, and whether this warp owns it at all
        const u32 inTot = before ? X : (cTot | X);
        const u32 inStart = before ? sPrev : cStart;
        const bool inMine = before || cValid;  // else the run starts in another warp's chunk: that warp finishes it
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
          if (last && mine && irregular) sm.walkQ[nWalk++][tid] = runStart;
          if (last && mine && !irregular && nh[j] > 1) {
            ev[j] = tot & 0x7FFFFFFFu;
            closeBits |= 1u << j;
          }
        }
      }
    }
    // ---- counting: single-element sets into the lane's histogram column, the rest into the block table
    {
      u32 pend = 0, rescBits = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const u32 c = ev[j];
        const bool single = c != 0 && (c & (c - 1)) == 0;
        if (single && ((closeBits >> j) & 1u)) rescBits |= 1u << j;  // a multi-mapping read resolved to one element (mm:1691)
        if (HIST) {
          const u32 row = single ? (u32)(__ffs(c) - 1) : (u32)(HIST_ROWS - 1);  // E <= 30: the last row only ever receives zeros
          sm.hist[row][tid] += single ? 1 : 0;
          if (c != 0 && !single) pend |= 1u << j;
        } else {
          if (c != 0) pend |= 1u << j;
        }
      }
      pClosResc += __popc(closeBits) | (__popc(rescBits) << 16);
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
      for (u32 q = 0; q < nWalk; ++q) queueWalk(sm.walkQ[q][tid]);
      if (GROUPS) {
        // the run (and, inside it, the group) still open at the end of the tile
        if (!serialTile) {
          const u32 incLast = __shfl_sync(FULL, inc, 31);
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
        const u32 incLast = __shfl_sync(FULL, inc, 31);
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
      cCont = __shfl_sync(FULL, tileEndsRun, 31) == 0;
    }
  }
  // ---- a run open at the end of the chunk continues in another warp's chunk (which leaves it alone: it does not own its start).
  //      This warp finishes it, as k_batch_lean does: the records of the run in the NEXT tile are annotated, 4 per lane, and lane 0
  //      takes the countdown through them (GROUPS: from the group open at the border through every later group of the run).  No
  //      per-hit counter moves: the hits belong to the other chunk.  A run that is irregular, that reaches beyond that tile or
  //      (GROUPS) that ends inside a group is marked for k_batch_walk.  (A run ending exactly at the chunk's last record was closed
  //      above, like the last run of the batch.)
  if (STRAT == 0 && cValid && cCont && t1 > t0) {
    const u32 next = t1 * WT_HITS;
    const u32 o = next - cStart;                               // records of the run inside this chunk
    const u32 inGroup = (GROUPS && cNh > 1) ? o % cNh : 0u;    // GROUPS: ... of which in the group open at the border
    const u32 walkFrom = GROUPS ? next - inGroup : cStart;
    const Annotator<MODE, true> annot{ix, fx, r.overlap};
    u32 tn[4], eq = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const size_t i = (size_t)next + lane * 4 + j;
      const bool v = i < h.n;
      const u64 kk = v ? normKey(__ldg(&h.key[i])) : ~cKey;
      tn[j] = v ? __ldg(&h.nh[i]) : 0u;
      if (kk == cKey) {
        eq |= 1u << j;
        sm.slowRes[warp][lane * 4 + j] = (u32)annot(__ldg(&h.start[i]), __ldg(&h.end[i]), __ldg(&h.meta[i]));
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
      if ((u32)j < mineN) bad |= tn[j] != (j ? tn[j > 0 ? j - 1 : 0] : prevNh);
    const bool anyBad = __any_sync(FULL, bad) || forceWalk || (int)cTot < 0 || fp == 32u;
    __syncwarp();
    if (lane == 0) {
      if (anyBad) {
        queueWalk(walkFrom);
      } else if (cNh > 1) {  // (a run of reads that are their own group has nothing to close)
        // one read closed by the countdown, counted like the closes of the loop
        auto closeRead = [&](u32 c) {
          pClosResc += 1u;
          if (c == 0) return;
          if ((c & (c - 1)) == 0) {
            pClosResc += 1u << 16;  // a multi-mapping read resolved to one element (mm:1691)
            if (HIST) sm.hist[__ffs(c) - 1][tid] += 1;
            else sm.bt.add((u64)c, 1, table);
          } else {
            sm.bt.add((u64)c, 1, table);
          }
        };
        if (!GROUPS) {
          if (o + tailLen == cNh) {
            u32 acc = cTot;
            for (u32 q = 0; q < tailLen; ++q) acc |= sm.slowRes[warp][q];
            closeRead(acc);
          } else {
            queueWalk(walkFrom);
          }
        } else if ((o + tailLen) % cNh != 0) {  // the run ends inside a group
          queueWalk(walkFrom);
        } else {
          u32 acc = inGroup ? cTot : 0u, pos = inGroup;
          for (u32 q = 0; q < tailLen; ++q) {
            acc |= sm.slowRes[warp][q];
            if (++pos == cNh) { closeRead(acc); acc = 0; pos = 0; }
          }
        }
      }
    }
  }
  u32 cHits = pHitsMiss & 0xFFFFu, cMiss = pHitsMiss >> 16, cUnassigned = cHits - (pAsgUniq & 0xFFFFu), cAmbiguous = pMultAmbi >> 16;
  u32 cUnique = pAsgUniq >> 16, cMultiple = pMultAmbi & 0xFFFFu;
  u32 cReads = cHits - cMultiple + (pClosResc & 0xFFFFu), cRescued = pClosResc >> 16;  // (what k_batch_walk opens and closes is added there)

  // ---- block epilogue: counters, the private histogram columns and the private table
  cHits = __reduce_add_sync(FULL, cHits); cUnassigned = __reduce_add_sync(FULL, cUnassigned);
  cAmbiguous = __reduce_add_sync(FULL, cAmbiguous); cUnique = __reduce_add_sync(FULL, cUnique);
  cMultiple = __reduce_add_sync(FULL, cMultiple); cReads = __reduce_add_sync(FULL, cReads);
  cRescued = __reduce_add_sync(FULL, cRescued); cMiss = __reduce_add_sync(FULL, cMiss);
  if (STRAT == 0 && !forceWalk) {
    pWalks = __reduce_add_sync(FULL, pWalks);
    if (lane == 0 && pWalks) atomicAdd(&ctl->walkCount, pWalks);
  }
  if (lane == 0) {
    atomicAdd(&sm.stat[ST_HITS], cHits); atomicAdd(&sm.stat[ST_UNASSIGNED], cUnassigned); atomicAdd(&sm.stat[ST_AMBIGUOUS], cAmbiguous);
    atomicAdd(&sm.stat[ST_UNIQUE], cUnique); atomicAdd(&sm.stat[ST_MULTIPLE], cMultiple); atomicAdd(&sm.stat[ST_READS], cReads);
    atomicAdd(&sm.stat[ST_RESCUED], cRescued); atomicAdd(&sm.stat[7], cMiss);
  }
  __syncthreads();
  sm.bt.flush(table);
  if (HIST) {
    for (u32 e = warp; e < HIST_ROWS; e += FAST_WARPS) {
      u32 v = 0;
#pragma unroll
      for (int q = 0; q < FAST_WARPS; ++q) v += sm.hist[e][lane + 32 * q];
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
